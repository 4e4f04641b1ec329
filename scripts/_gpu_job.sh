python -m pytest tests/test_gpu_cs.py -x -q 2>&1 | tail -3 > gpurun_out/r2j_tests.log
python scripts/cs_bench.py --reads 4000000 > gpurun_out/r2j_cs.json 2> gpurun_out/r2j_cs.err
cat gpurun_out/r2j_tests.log gpurun_out/r2j_cs.json; tail -3 gpurun_out/r2j_cs.err
