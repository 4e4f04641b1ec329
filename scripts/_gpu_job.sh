python -m pytest tests/test_gpu_select_pairs.py tests/test_gpu_cs.py tests/test_gpu_pipeline_vs_ngm.py -x -q 2>&1 | tail -8 > gpurun_out/r1q_tests.log
python scripts/cs_bench.py --reads 2000000 > gpurun_out/r1q_cs.json 2> gpurun_out/r1q_cs.err
python bench.py --steps 3 --no-e2e > gpurun_out/r1q_bench.json 2> gpurun_out/r1q_bench.err
tail -3 gpurun_out/r1q_tests.log; cat gpurun_out/r1q_cs.json; tail -c 600 gpurun_out/r1q_bench.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r1q_bench.json"))
print(d["value"], d["ms_per_step"])
print(json.dumps(d["candidate_search"].get("paired_end"), indent=1))
print(d["candidate_search"]["cs_ms"], d["candidate_search"]["pipeline_ms"], d["candidate_search"]["parity_sample"])
PY
