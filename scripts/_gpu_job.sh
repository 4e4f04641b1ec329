python -m pytest tests/test_gpu_cs.py -x -q 2>&1 | tail -3 > gpurun_out/r2g_tests.log
python scripts/cs_bench.py --reads 2000000 > gpurun_out/r2g_cs.json 2> gpurun_out/r2g_cs.err
python scripts/cs_bench.py --reads 500000 --read-len 250 > gpurun_out/r2g_cs250.json 2>> gpurun_out/r2g_cs.err
M=gpu__time_duration.sum,dram__bytes_read.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
ncu --metrics $M --clock-control none -k regex:cs_search_kernel -c 1 --csv --log-file gpurun_out/r2g_ncu.csv python scripts/cs_bench.py --reads 1000000 --reps 1 > /dev/null 2>&1
cat gpurun_out/r2g_tests.log gpurun_out/r2g_cs.json gpurun_out/r2g_cs250.json; tail -3 gpurun_out/r2g_cs.err; grep -v "^==" gpurun_out/r2g_ncu.csv | cut -d, -f 21- | tail -5
