#!/usr/bin/env python
"""Candidate-search micro benchmark: index build + k-mer vote on the device, kernel level.

    python scripts/cs_bench.py --reads 2000000 [--contigs 24 --contig-len 125000000 --read-len 150 --sensitivity 0.5]
"""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--contigs", type=int, default=24)
    ap.add_argument("--contig-len", type=int, default=125_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--sensitivity", type=float, default=0.5)
    ap.add_argument("--sub-rate", type=float, default=0.01)
    ap.add_argument("--indel-rate", type=float, default=0.0005)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--mutate", type=int, default=0, help="1: --bs-mapping run (reads C>T converted, index of every position, mutated k-mers); 2: --slam-seq 4")
    ap.add_argument("--l2-fetch", type=int, default=0, help="experiment: cudaLimitMaxL2FetchGranularity (32 / 64 / 128)")
    args = ap.parse_args()
    import torch
    from nextgenmap_b200 import workload
    from nextgenmap_b200.host import CudaSW
    from nextgenmap_b200.host.cuda_sw import CsParams, _CContigRec
    dev = torch.device("cuda", 0)
    if args.l2_fetch:
        import glob
        torch.zeros(1, device=dev)
        rt = C.CDLL(sorted(glob.glob(str(Path(torch.__file__).parents[1] / "nvidia" / "cuda_runtime" / "lib" / "libcudart.so*")))[0])
        rc = rt.cudaDeviceSetLimit(5, C.c_size_t(args.l2_fetch))          # cudaLimitMaxL2FetchGranularity
        got = C.c_size_t(0)
        rt.cudaDeviceGetLimit(C.byref(got), 5)
        print(f"cudaLimitMaxL2FetchGranularity: rc {rc}, now {got.value}", file=sys.stderr)
    L = args.read_len
    qml, corridor = workload.shapes_for(L)
    ref = workload.make_reference(dev, args.contigs, args.contig_len, seed=20261017)
    batch = workload.make_reads(ref, args.reads, L, qml, corridor, seed=20261019, sub_rate=args.sub_rate, indel_rate=args.indel_rate)
    if args.mutate:                                             # the chemistry the mutated search undoes: C read as T (bisulfite) / T read as C (SLAMseq)
        g = torch.Generator(device=dev)
        g.manual_seed(7)
        draw = torch.rand(batch.reads.shape, device=dev, generator=g)
        frm, to, rate = (ord("C"), ord("T"), 0.9) if args.mutate == 1 else (ord("T"), ord("C"), 0.05)
        batch.reads[(batch.reads == frm) & (draw < rate)] = to
    sw = CudaSW(qml, min(corridor, 155), bs_mapping=1 if args.mutate == 1 else 0, slam_seq=4 if args.mutate == 2 else 0)
    lib, ctx = sw.lib, sw.ctx
    st = torch.cuda.current_stream().cuda_stream
    assert lib.ngm_b200_dev_set_reference(ctx, ref.packed.data_ptr(), ref.concat_len, st) >= 0
    arr = (_CContigRec * args.contigs)()
    for i, s0 in enumerate(ref.contig_start):
        arr[i].start, arr[i].length, arr[i].name_len = int(s0), int(ref.contig_len), 0
    csp = CsParams(13, 0 if args.mutate == 1 else 2, 2, 1, args.sensitivity, 0.0, 0, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc = lib.ngm_b200_cs_build_index(ctx, C.byref(csp), arr, args.contigs)
    assert rc >= 0, sw._err()
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    info = sw.cs_index_info()
    if args.mutate:
        sw.cs_configure_mutation(bs_mapping=1 if args.mutate == 1 else 0, slam_seq=4 if args.mutate == 2 else 0)
    n = args.reads
    cap = 4 * n + 1024
    d_cb = torch.empty(n + 1, dtype=torch.int32, device=dev)
    d_pairs = torch.empty((cap, 16), dtype=torch.uint8, device=dev)
    d_votes = torch.empty(cap, dtype=torch.float32, device=dev)

    def run():
        rc = lib.ngm_b200_dev_cs_search(ctx, batch.reads.data_ptr(), n, qml, 0, d_cb.data_ptr(), d_pairs.data_ptr(), d_votes.data_ptr(), cap, None, st)
        assert rc >= 0, sw._err()

    run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.reps):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.reps
    total = int(d_cb[n].item())
    # truth check: the true locus is among the candidates
    begin = d_cb[:-1].long()
    first_loc = d_pairs.view(torch.int64).view(cap, 2)[begin.clamp(max=cap - 1), 0] + (min(corridor, 155) >> 1)
    has = d_cb[1:] > d_cb[:-1]
    near = ((first_loc - batch.true_pos).abs() <= corridor + 8) & has
    print(json.dumps({"reads": n, "read_len": L, "mutate": args.mutate, "index_build_s": build_s, "index_positions": info["table_len"], "max_kfreq": info["max_kfreq"],
                      "cs_ms": ms, "cs_reads_per_s": n / (ms * 1e-3), "candidates_per_read": total / n, "exact_reads": sw.cs_exact_reads(), "exact_reasons": sw.cs_exact_reasons(),
                      "first_candidate_at_truth": float(near.float().mean().item())}))


if __name__ == "__main__":
    main()
