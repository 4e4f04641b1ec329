python -m pytest tests/test_gpu_pipeline_vs_ngm.py tests/test_gpu_select_pairs.py -x -q 2>&1 | tail -5 > gpurun_out/r2i_tests.log
cat gpurun_out/r2i_tests.log
