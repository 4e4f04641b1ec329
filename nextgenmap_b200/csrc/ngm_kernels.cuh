// ngm_kernels.cuh -- align kernel (forward + backtrace + format), packers, selection.
#pragma once

#include "ngm_common.cuh"
#include "ngm_dp_i32.cuh"
#include "ngm_format.cuh"
#include "../../include/ngm_b200.h"

namespace ngm {

// ---------------------------------------------------------------------------
// K3/K4 + K5 + computeCigarMD in one launch: one thread per alignment.
// ---------------------------------------------------------------------------
template <int W, int LO, int MODE>
__global__ void __launch_bounds__(128) align_i32_kernel(const __grid_constant__ DevParams P, const PairDesc *__restrict__ pairs, int n,
		const uint32_t *__restrict__ reads_fwd, const uint32_t *__restrict__ reads_rev, const uint16_t *__restrict__ rlen,
		const uint32_t *__restrict__ ref4, uint32_t *__restrict__ ptr_scratch, uint16_t *__restrict__ ops_scratch, int stride, int ops_cap,
		ngm_b200_align_rec *__restrict__ recs, char *__restrict__ strings, uint32_t str_cap, uint32_t *__restrict__ cursor, float *__restrict__ out_best) {
	__shared__ uint2 s_lut[16];
	if (threadIdx.x < 16) s_lut[threadIdx.x] = P.lut[threadIdx.x];
	__syncthreads();
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	const bool valid = idx < n;
	const int slot = idx;                      // launches never exceed `stride` pairs
	PairCtx c;
	c.rp = reads_fwd;
	c.wp = ref4;
	c.sub = 0;
	c.len = 0;
	c.dir = 0;
	TraceOut t;
	t.ok = 0;
	t.pos = MODE == 0 ? 0 : -1;
	t.qstart = t.qend = t.sp = 0;
	AlignScratch S;
	S.ptr = ptr_scratch;
	S.ops = ops_scratch;
	S.stride = stride;
	S.ops_cap = ops_cap;
	uint16_t *ops = ops_scratch + slot;
	if (valid) {
		uint32_t flags;
		float bs = MODE == 0 ? -1.0f : (float) kEndFreeMin;
		if (load_pair(P, pairs, idx, reads_fwd, reads_rev, rlen, ref4, c, flags)) {
			int best_read, best_ref, best_score, read_count;
			forward_i32<W, LO, MODE>(P, s_lut, c, S, slot, best_read, best_ref, best_score, read_count);
			bs = (float) best_score;
			t = backtrace_u16<W, MODE>(P, s_lut, c, S, slot, best_read, best_ref, best_score, read_count);
		}
		if (out_best != nullptr) out_best[idx] = bs;
	}
	FormatOut f;
	f.cigar_len = f.md_len = f.match = f.mismatch = f.total = f.read_index = 0;
	if (t.ok) f = format_cigar_md<false>(P, c, ops, stride, t, nullptr, nullptr);
	// warp-aggregated allocation in the string heap: one atomic per warp
	const uint32_t need = t.ok ? (uint32_t) (f.cigar_len + f.md_len) : 0u;
	const int lane = threadIdx.x & 31;
	uint32_t incl = need;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) incl += v;
	}
	uint32_t base = 0;
	if (lane == 31 && incl) base = atomicAdd(cursor, incl);
	base = __shfl_sync(0xffffffffu, base, 31);
	const uint32_t off = base + incl - need;
	if (!valid) return;
	ngm_b200_align_rec r;
	r.position_offset = t.pos;
	r.qstart = t.qstart;
	r.qend = t.qend > 0 ? t.qend : 0;
	r.str_off = off;
	if (t.ok) {
		if ((uint64_t) off + need <= (uint64_t) str_cap) format_cigar_md<true>(P, c, ops, stride, t, strings + off, strings + off + f.cigar_len);
		r.nm = f.mismatch;
		r.identity = (float) f.match * 1.0f / (float) f.total;       // SWOclCigar.cpp:609
		r.score = (float) f.read_index;                              // SWOclCigar.cpp:613
		r.cigar_len = (uint16_t) f.cigar_len;
		r.md_len = (uint16_t) f.md_len;
	} else {
		// the reference's backtracking kernel skips this lane and its host code then reads
		// uninitialised memory (oclSwCigar.cl:78, SWOclCigar.cpp:322-328); we report the
		// failure convention Score = -1 (SWOclCigar.cpp:326) -- see DESIGN.md
		r.nm = 0;
		r.identity = 0.0f;
		r.score = -1.0f;
		r.cigar_len = 0;
		r.md_len = 0;
	}
	reinterpret_cast<uint4 *>(recs)[2 * (size_t) idx] = *reinterpret_cast<const uint4 *>(&r);
	reinterpret_cast<uint4 *>(recs)[2 * (size_t) idx + 1] = *(reinterpret_cast<const uint4 *>(&r) + 1);
}

}  // namespace ngm
