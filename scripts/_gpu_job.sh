python -m pytest tests/test_gpu_select_pairs.py tests/test_gpu_pipeline_vs_ngm.py -x -q 2>&1 | tail -25 > gpurun_out/r1o_tests.log
cat gpurun_out/r1o_tests.log
