python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2h_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1
python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python bench.py --impl reference > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ngm:: -c 4000 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2h_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cs_search_kernel -c 1 -o gpurun_out/r2h_cs_full -f python scripts/cs_bench.py --reads 1000000 --reps 1 > gpurun_out/r2h_ncu_cs.log 2>&1
cat gpurun_out/r2h_tests.log gpurun_out/r2h_smoke.log; tail -c 300 gpurun_out/r2h_bench.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2h_bench.json"))
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"], d["gpu_launches"])
cs = d["candidate_search"]
print({k: cs[k] for k in ("cs_ms", "cs_reads_per_s", "pipeline_ms", "pipeline_reads_per_s", "parity_sample", "index_build_seconds")})
print(cs.get("paired_end")); print(cs.get("e2e_map_batch")); print(cs.get("sam_format")); print(cs["roofline"])
print(open("gpurun_out/r2h_bench_ref.json").read()[:200])
PY
