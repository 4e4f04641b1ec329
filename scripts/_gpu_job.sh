python -m pytest tests/test_gpu_cs.py -x -q 2>&1 | tail -3 > gpurun_out/r2b_tests.log
python scripts/cs_bench.py --reads 2000000 > gpurun_out/r2b_cs.json 2> gpurun_out/r2b_cs.err
python scripts/cs_bench.py --reads 500000 --read-len 250 > gpurun_out/r2b_cs250.json 2>> gpurun_out/r2b_cs.err
cat gpurun_out/r2b_tests.log gpurun_out/r2b_cs.json gpurun_out/r2b_cs250.json; tail -3 gpurun_out/r2b_cs.err
