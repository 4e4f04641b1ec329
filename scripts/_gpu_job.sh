python -m pytest tests/test_gpu_cs.py -x -q 2>&1 | tail -4 > gpurun_out/r1u_tests.log
python scripts/cs_bench.py --reads 2000000 > gpurun_out/r1u_cs.json 2> gpurun_out/r1u_cs.err
python scripts/cs_bench.py --reads 500000 --read-len 250 > gpurun_out/r1u_cs250.json 2>> gpurun_out/r1u_cs.err
python scripts/cs_bench.py --reads 2000000 --l2-fetch 32 > gpurun_out/r1u_cs_l2.json 2>> gpurun_out/r1u_cs.err
M=gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:cs_search_kernel -c 1 --csv --log-file gpurun_out/r1u_ncu.csv python scripts/cs_bench.py --reads 1000000 --reps 1 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:cs_search_kernel -c 1 --csv --log-file gpurun_out/r1u_ncu_l2.csv python scripts/cs_bench.py --reads 1000000 --reps 1 --l2-fetch 32 > /dev/null 2>&1
cat gpurun_out/r1u_tests.log gpurun_out/r1u_cs.json gpurun_out/r1u_cs250.json gpurun_out/r1u_cs_l2.json; tail -3 gpurun_out/r1u_cs.err; grep -v "^==" gpurun_out/r1u_ncu.csv | cut -d, -f 21- | tail -5; grep -v "^==" gpurun_out/r1u_ncu_l2.csv | cut -d, -f 21- | tail -5
