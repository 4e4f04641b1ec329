python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
tail -c 400 gpurun_out/r2d_bench_n2.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2d_bench_n2.json"))
print(d["value"], d["n_gpus"], d["ms_per_step"], d["e2e"]["value"], d["scaling"])
cs = d["candidate_search"]
print({k: cs.get(k) for k in ("cs_reads_per_s", "pipeline_reads_per_s", "parity_sample", "error")})
pe = cs.get("paired_end")
print({k: pe.get(k) for k in ("pipeline_reads_per_s", "select_pairs_ms", "parity_sample", "error")} if pe else None)
PY
