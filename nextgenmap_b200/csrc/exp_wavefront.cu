// exp_wavefront.cu -- EXPERIMENT (SURVEY 7.2 / north_star): the score-only banded DP as a warp-level skewed wavefront.
//
// north_star sketches "one warp per read x candidate DP tile ... anti-diagonal wavefront driven by warp shuffles".  This file is
// that design, built so it can be timed against the production kernel (score_s16_kernel: one thread owns two whole pairs, the band lives
// in registers, no cross-lane dependency) on the same pairs -- bench.py `design_ab`.  It is not on any product path.
//
// Layout.  Half a warp (16 lanes) owns a DP tile; lane l owns band columns 2l and 2l + 1 (corridor <= 32).  Cell (r, j) depends on
// (r, j - 1) [left], (r - 1, j) [diag] and (r - 1, j + 1) [up], so lane l can take row r at step 2r + l at the earliest: its left neighbour
// must have finished row r (step 2r + l - 1) and its right neighbour row r - 1 (step 2(r - 1) + l + 1).  A lane is therefore busy
// every OTHER step for one tile; the idle steps are filled with a second tile whose schedule is shifted by one step, so a warp carries
// four pairs and every lane computes two cells per step.  Per step and lane: two SHFL (left value from lane l - 1, up value from lane
// l + 1: both were computed in the previous step and belong to the tile this lane works on now), two cells of the recurrence, the
// substitution lookups for this lane's OWN row (rows differ between lanes, so nothing about a row can be shared across the warp),
// and the sliding read / window registers.
#include "ngm_ctx.h"
#include "ngm_dp_i32.cuh"
#include "ngm_launch.h"

namespace ngm {

struct WfTile {
	const uint32_t *rp, *wp;
	int sub, len, dir, active;
	uint32_t rdw;              // current read word
	uint32_t w0, w1;           // window words holding this lane's two nibbles
	int d0, d1;                // H(r - 1, 2l), H(r - 1, 2l + 1)
	int best;
};

__device__ __forceinline__ void wf_load(const DevParams &P, const PairDesc *pairs, int idx, int n, const uint32_t *rf, const uint32_t *rr, const uint16_t *rlen,
		const uint32_t *ref4, WfTile &t) {
	PairCtx c;
	uint32_t flags;
	const bool ok = load_pair(P, pairs, min(idx, n - 1), rf, rr, rlen, ref4, c, flags);
	t.rp = c.rp;
	t.wp = c.wp;
	t.sub = c.sub;
	t.len = c.len;
	t.dir = c.dir;
	t.active = ok && idx < n;
	t.d0 = t.d1 = 0;
	t.best = 0;
	t.rdw = kNulWord;
	t.w0 = t.w1 = 0;
}

// one step of one tile on this lane: row r (may be out of range: then the lane only forwards zeros)
__device__ __forceinline__ void wf_step(const DevParams &P, const uint2 *s_lut, WfTile &t, int r, int l, int left_in, int up_in, int &h0, int &h1) {
	h0 = h1 = 0;
	if (r < 0 || r >= t.len) return;
	if ((r & 7) == 0) t.rdw = __ldg(t.rp + (r >> 3));
	const int rc = (t.rdw >> (4 * (r & 7))) & 7;
	const int nib = t.sub + r + 2 * l;                         // window nibble of column 2l
	if ((nib & 7) == 0 || r == 0) {
		t.w0 = __ldg(t.wp + (nib >> 3));
		t.w1 = __ldg(t.wp + (nib >> 3) + 1);
	}
	const uint32_t both = __funnelshift_r(t.w0, t.w1, 4 * (nib & 7));
	const uint2 tab = s_lut[t.dir * 8 + rc];
	const uint32_t sw = prmt(tab.x, tab.y, both);
	const int s0 = sbyte<0>(sw), s1 = sbyte<1>(sw);
	const int j0 = 2 * l, j1 = 2 * l + 1;
	int a = __viaddmax_s32(t.d1, P.gap_read, t.d0 + s0);       // up of column 2l is this lane's own column 2l + 1
	a = __viaddmax_s32_relu(left_in, P.gap_ref, a);
	h0 = j0 < P.corridor ? a : 0;
	int b = __viaddmax_s32(up_in, P.gap_read, t.d1 + s1);
	b = __viaddmax_s32_relu(h0, P.gap_ref, b);
	h1 = j1 < P.corridor ? b : 0;
	t.d0 = h0;
	t.d1 = h1;
	t.best = __vimax3_s32(t.best, h0, h1);
}

__global__ void __launch_bounds__(128) wf_score_kernel(const __grid_constant__ DevParams P, const PairDesc *__restrict__ pairs, int n,
		const uint32_t *__restrict__ reads_fwd, const uint32_t *__restrict__ reads_rev, const uint16_t *__restrict__ rlen,
		const uint32_t *__restrict__ ref4, float *__restrict__ out) {
	__shared__ uint2 s_lut[16];
	if (threadIdx.x < 16) s_lut[threadIdx.x] = P.lut[threadIdx.x];
	__syncthreads();
	const int lane = threadIdx.x & 31, l = lane & 15, g = lane >> 4;
	const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int base = warp_global * 4 + 2 * g;                  // this half-warp's two tiles
	if (warp_global * 4 >= n) return;
	WfTile tx, ty;
	wf_load(P, pairs, base, n, reads_fwd, reads_rev, rlen, ref4, tx);
	wf_load(P, pairs, base + 1, n, reads_fwd, reads_rev, rlen, ref4, ty);
	int rows = max(tx.len, ty.len);
	rows = max(rows, __shfl_xor_sync(0xffffffffu, rows, 16));   // both halves of the warp run the same number of steps
	const int steps = 2 * rows + 17;
	int last0 = 0, last1 = 0;                                    // what this lane computed in the previous step
	for (int t = 0; t < steps; ++t) {
		int left_in = __shfl_up_sync(0xffffffffu, last1, 1, 16);
		int up_in = __shfl_down_sync(0xffffffffu, last0, 1, 16);
		if (l == 0) left_in = 0;                                 // band border (local mode sentinel)
		if (l == 15) up_in = 0;
		const int which = (t + l) & 1;                           // tile X on (t + l) even
		const int r = (t - l - which) >> 1;
		int h0, h1;
		if (which == 0) wf_step(P, s_lut, tx, (t - l) >= 0 ? r : -1, l, left_in, up_in, h0, h1);
		else wf_step(P, s_lut, ty, (t - l - 1) >= 0 ? r : -1, l, left_in, up_in, h0, h1);
		last0 = h0;
		last1 = h1;
	}
	// tile maxima: reduce over the 16 lanes
	int bx = tx.best, by = ty.best;
#pragma unroll
	for (int d = 8; d > 0; d >>= 1) {
		bx = max(bx, __shfl_xor_sync(0xffffffffu, bx, d, 16));
		by = max(by, __shfl_xor_sync(0xffffffffu, by, d, 16));
	}
	if (l == 0) {
		if (base < n) out[base] = tx.active ? (float) bx : -1.0f;
		if (base + 1 < n) out[base + 1] = ty.active ? (float) by : -1.0f;
	}
}

}  // namespace ngm

using namespace ngm;

extern "C" int ngm_b200_exp_wavefront_score(ngm_b200_ctx *c, int n, const void *d_resolved_from_pairs, void *d_scores, void *stream) {
	if (c == nullptr || d_resolved_from_pairs == nullptr || d_scores == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	if (n <= 0) return 0;
	if (c->dp.corridor > 32) return fail(NGM_B200_EINVAL, "the wavefront experiment covers corridors up to 32");
	CU(cudaSetDevice(c->device));
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	int rc = resolve_pairs_for(c, d_resolved_from_pairs, n, st);
	if (rc) return rc;
	const int warps = (n + 3) / 4;
	wf_score_kernel<<<(warps + 3) / 4, 128, 0, st>>>(c->dp, c->d_rpairs.as<PairDesc>(), n, c->d_rfwd.as<uint32_t>(), c->d_rrev.as<uint32_t>(),
			c->d_rrlen.as<uint16_t>(), c->d_ref4.as<uint32_t>(), static_cast<float *>(d_scores));
	c->launches += 1;
	CU(cudaGetLastError());
	return n;
}
