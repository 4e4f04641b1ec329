#!/bin/bash
out=gpurun_out/sweep2.jsonl
: > $out
python bench.py --read-len 400 --reads 2000000 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>>gpurun_out/sweep2.err >> $out
python bench.py --read-len 250 --corridor 80 --sub-rate 0.12 --indel-rate 0.004 --reads 2000000 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>>gpurun_out/sweep2.err >> $out
python - <<'PY'
import json
for line in open("gpurun_out/sweep2.jsonl"):
    d = json.loads(line)
    print(d["config"]["workload"][:60], "| value %.1f M reads/s" % (d["value"] / 1e6), "| ms", {k: round(v, 2) for k, v in d["kernel_ms"].items() if not k.endswith("share")},
          "| score %.0f GCUPS align %.0f GCUPS" % (d["roofline_alu"]["score_gcups"], d["roofline_alu"]["align_gcups"]), "| parity", d["parity_sample"])
PY
