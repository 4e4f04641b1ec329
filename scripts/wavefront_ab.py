#!/usr/bin/env python
"""Design A/B: north_star's warp-per-tile skewed wavefront (shuffles) against the production
thread-per-pair(-pair) register-band kernel, local score, same pairs of the headline workload.

    python scripts/wavefront_ab.py [--pairs 6000000]
"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--contigs", type=int, default=24)
    ap.add_argument("--contig-len", type=int, default=125_000_000)
    args = ap.parse_args()
    import numpy as np
    import torch
    from nextgenmap_b200 import workload
    from nextgenmap_b200.host import CudaSW
    dev = torch.device("cuda", 0)
    L = 150
    qml, corridor = workload.shapes_for(L)
    ref = workload.make_reference(dev, args.contigs, args.contig_len, seed=20261017)
    batch = workload.make_reads(ref, args.reads, L, qml, corridor, seed=20261019)
    sw = CudaSW(qml, corridor)
    lib, ctx = sw.lib, sw.ctx
    st = torch.cuda.current_stream().cuda_stream
    lib.ngm_b200_exp_wavefront_score.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    assert lib.ngm_b200_dev_set_reference(ctx, ref.packed.data_ptr(), ref.concat_len, st) >= 0
    assert lib.ngm_b200_dev_set_reads(ctx, batch.reads.data_ptr(), batch.n_reads, qml, st) >= 0
    n = batch.n_pairs
    a = torch.empty(n, dtype=torch.float32, device=dev)
    b = torch.empty(n, dtype=torch.float32, device=dev)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def prod():
        assert lib.ngm_b200_dev_score_pairs(ctx, 0, n, batch.pairs.data_ptr(), a.data_ptr(), st) == n

    def wave():
        assert lib.ngm_b200_exp_wavefront_score(ctx, n, batch.pairs.data_ptr(), b.data_ptr(), st) == n, sw._err()

    ms_prod, ms_wave = timed(prod), timed(wave)
    same = bool(torch.equal(a, b))
    cells = n * L * corridor
    print(json.dumps({"pairs": n, "shape": f"{qml}/{corridor}", "production_thread_per_pair_s16x2_ms": ms_prod, "wavefront_warp_shuffle_ms": ms_wave,
                      "production_gcups": cells / ms_prod / 1e6, "wavefront_gcups": cells / ms_wave / 1e6, "speed_ratio_production_over_wavefront": ms_wave / ms_prod,
                      "scores_identical": same, "mismatches": int((a != b).sum().item())}))


if __name__ == "__main__":
    main()
