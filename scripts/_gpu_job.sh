python -m pytest tests/test_gpu_pipeline_vs_ngm.py -x -q 2>&1 | tail -5 > gpurun_out/r2f_tests.log
python bench.py --steps 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
cat gpurun_out/r2f_tests.log; tail -c 300 gpurun_out/r2f_bench.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2f_bench.json"))
cs = d["candidate_search"]
print(d["value"], d["e2e"]["value"], cs["cs_reads_per_s"], cs["pipeline_reads_per_s"], cs.get("sam_format"), cs.get("e2e_map_batch"))
PY
