// ngm_dp_i32.cuh -- banded DP kernels with one (read, window) pair per thread and
// int32 cells.  This is the general path: any integer scoring whose row LUT fits
// signed bytes.  The s16x2 two-pairs-per-thread kernels (ngm_dp_s16.cuh) are the
// fast path for the default parameter range.
//
// Reference semantics restated here (CPU-device variants):
//   score, local     oclSwScore.cl:111-154     (oclSW)
//   score, end-free  oclEndFreeScore.cl:5-55   (oclSW_Global)
//   forward, local   oclSwScore.cl:4-107       (oclSW_Score)
//   forward, end-free oclEndFreeScore.cl:58-146 (oclSW_ScoreGlobal)
//
// Layout of one thread's work.  The band (corridor cells of the previous row)
// lives in registers: line[0..W], W = compile-time capacity >= corridor, slots
// >= corridor pinned to the mode's sentinel (0 local / -16000 end-free) exactly
// like the never-written line[corridor] of the reference.  Rows are visited in
// "q" order, q = row + (window_start & 7): with that skew the window nibble of
// cell (row, j) is nibble q + j of the thread's word stream, so word refills and
// funnel-shift amounts are warp-uniform although every pair starts at its own
// nibble.  Rows q < (window_start & 7) and rows past the read see read code 6
// (NUL), whose score row is all zero: such rows leave the initial state, the
// running maximum and the first arg-max unchanged ("pad rows").
#pragma once

#include "ngm_common.cuh"

namespace ngm {

template <int W>
struct BandGeom {
	static constexpr int kGroups = W / 4;                 // PRMT groups of four cells
	static constexpr int kAligned = (W + 7) / 8;          // aligned window words per row
	static constexpr int kWin = (W + 7 + 7) / 8;          // window words held in registers
	static constexpr int kPtrWords = (W + 15) / 16;       // 2 bits per cell
};

// Rows per unrolled loop body of the int32 kernels: eight on narrow bands; wide bands (capacity > 48) would not fit the instruction cache
// (the s16x2 kernels measured 1.7x on this, see NGM_FWD1_WIDE_UNROLL).  -DNGM_I32_WIDE_UNROLL=8 restores the full unroll.
#ifndef NGM_I32_WIDE_UNROLL
#define NGM_I32_WIDE_UNROLL 2
#endif
template <int W>
constexpr int kUnrollI32 = W > 48 ? NGM_I32_WIDE_UNROLL : 8;

// Everything the row loop needs about one pair.
struct PairCtx {
	const uint32_t *rp;      // packed read row
	const uint32_t *wp;      // first window word
	int sub;                 // window_start & 7
	int len;                 // rows with data (index of last non-NUL read char + 1)
	int dir;                 // 0 / 1 -> FWD / REV score matrix
};

__device__ __forceinline__ bool load_pair(const DevParams &P, const PairDesc *__restrict__ pairs, int idx,
		const uint32_t *__restrict__ reads_fwd, const uint32_t *__restrict__ reads_rev, const uint16_t *__restrict__ rlen,
		const uint32_t *__restrict__ ref4, PairCtx &c, uint32_t &flags) {
	PairDesc d = pairs[idx];
	flags = d.flags;
	const uint32_t *rbase = (d.flags & PF_REVERSE) ? reads_rev : reads_fwd;
	c.rp = rbase + (size_t) d.read_idx * P.read_words;
	c.wp = ref4 + (d.win_nib >> 3);
	c.sub = (int) (d.win_nib & 7);
	c.len = rlen[d.read_idx];
	c.dir = (d.flags & PF_DIR) ? 1 : 0;
	return !(d.flags & PF_INACTIVE);
}

// ---------------------------------------------------------------------------
// score only
// ---------------------------------------------------------------------------
template <int W, int LO, int MODE>
__global__ void __launch_bounds__(128) score_i32_kernel(const __grid_constant__ DevParams P, const PairDesc *__restrict__ pairs, int n,
		const uint32_t *__restrict__ reads_fwd, const uint32_t *__restrict__ reads_rev, const uint16_t *__restrict__ rlen,
		const uint32_t *__restrict__ ref4, float *__restrict__ out, const int *__restrict__ sel, const int *__restrict__ n_dev) {
	using G = BandGeom<W>;
	__shared__ uint2 s_lut[16];
	if (threadIdx.x < 16) s_lut[threadIdx.x] = P.lut[threadIdx.x];
	__syncthreads();
	if (n_dev != nullptr) n = min(n, *n_dev);
	const int slot_i = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot_i >= n) return;
	const int idx = sel != nullptr ? sel[slot_i] : slot_i;
	constexpr int SENT = MODE == 0 ? 0 : kEndFreeMin;
	PairCtx c;
	uint32_t flags;
	if (!load_pair(P, pairs, idx, reads_fwd, reads_rev, rlen, ref4, c, flags)) {
		out[idx] = MODE == 0 ? -1.0f : (float) kEndFreeMin;      // quad skipped (oclSwScore.cl:123-124)
		return;
	}
	const int corridor = P.corridor, gap_read = P.gap_read, gap_ref = P.gap_ref;
	int line[W + 1];
#pragma unroll
	for (int j = 0; j <= W; ++j) line[j] = (j < corridor) ? 0 : SENT;
	int best = 0;
	uint32_t win[G::kWin];
#pragma unroll
	for (int k = 0; k < G::kWin; ++k) win[k] = __ldg(c.wp + k);
	const int nqw = (c.sub + c.len + 7) >> 3;
	const uint2 *lut = s_lut + c.dir * 8;
	uint32_t prev = kNulWord;
	for (int qw = 0; qw < nqw; ++qw) {
		const uint32_t cur = __ldg(c.rp + qw);
		const uint32_t rdw = __funnelshift_l(prev, cur, 4 * c.sub);
		prev = cur;
		const uint32_t nextw = __ldg(c.wp + qw + G::kWin);
#pragma unroll kUnrollI32<W>
		for (int t = 0; t < 8; ++t) {
			const uint2 tab = lut[(rdw >> (4 * t)) & 7];
			uint32_t al[G::kAligned];
#pragma unroll
			for (int k = 0; k < G::kAligned; ++k) al[k] = (kUnrollI32<W> == 8 && t == 0) ? win[k] : __funnelshift_r(win[k], win[k + 1], 4 * t);
			int left = SENT;
#pragma unroll
			for (int m = 0; m < G::kGroups; ++m) {
				const uint32_t sel = (m & 1) ? (al[m >> 1] >> 16) : al[m >> 1];
				const uint32_t sw = prmt(tab.x, tab.y, sel);
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					const int j = 4 * m + i;
					const int s = i == 0 ? sbyte<0>(sw) : i == 1 ? sbyte<1>(sw) : i == 2 ? sbyte<2>(sw) : sbyte<3>(sw);
					const int d = line[j] + s;
					const int u = __viaddmax_s32(line[j + 1], gap_read, d);
					int h = MODE == 0 ? __viaddmax_s32_relu(left, gap_ref, u) : __viaddmax_s32(left, gap_ref, u);
					if (j >= LO) h = (j < corridor) ? h : SENT;
					left = h;
					line[j] = h;
				}
			}
			if (MODE == 0) {
#pragma unroll
				for (int j = 0; j + 1 < W; j += 2) best = __vimax3_s32(best, line[j], line[j + 1]);
				if (W & 1) best = max(best, line[W - 1]);
			}
		}
#pragma unroll
		for (int k = 0; k + 1 < G::kWin; ++k) win[k] = win[k + 1];
		win[G::kWin - 1] = nextw;
	}
	if (MODE == 1) {
		best = kEndFreeMin;
#pragma unroll
		for (int j = 0; j < W; ++j) best = max(best, line[j]);
	}
	out[idx] = (float) best;
}

// ---------------------------------------------------------------------------
// forward pass with pointers + in-thread backtrace (K3/K4 + K5 of SURVEY 2a)
// ---------------------------------------------------------------------------
struct AlignScratch {
	uint32_t *ptr;           // [rows_cap][stride][kPtrWords] pointer matrix, 2 bits per cell
	uint16_t *ops;           // [ops_cap][stride] RLE op stack, (len << 4 | op) like the reference's shorts
	int stride;              // alignments per launch (padded)
	int ops_cap;
};

// Backtrace result handed to the formatter.
struct TraceOut {
	int ok;                  // 0 -> the reference's backtracking kernel skips this lane (oclSwCigar.cl:78)
	int pos;                 // ref_position (abs_ref_index + 1) or best_read_index when !ok
	int qstart;
	int qend;
	int sp;                  // number of entries on the op stack
};

template <int W, int LO, int MODE>
__device__ __forceinline__ void forward_i32(const DevParams &P, const uint2 *s_lut, const PairCtx &c, const AlignScratch &S, int slot,
		int &best_read, int &best_ref, int &best_score, int &read_count) {
	using G = BandGeom<W>;
	constexpr int SENT = MODE == 0 ? 0 : kEndFreeMin;
	const int corridor = P.corridor, gap_read = P.gap_read, gap_ref = P.gap_ref;
	int line[W + 1];
#pragma unroll
	for (int j = 0; j <= W; ++j) line[j] = (j < corridor) ? 0 : SENT;
	uint32_t win[G::kWin];
#pragma unroll
	for (int k = 0; k < G::kWin; ++k) win[k] = __ldg(c.wp + k);
	const int nqw = (c.sub + c.len + 7) >> 3;
	const uint2 *lut = s_lut + c.dir * 8;
	uint32_t prev = kNulWord;
	int curr_max = 0, rcount = 0;
	best_read = 0;
	best_ref = 0;
	uint32_t *prow = S.ptr + (size_t) slot * G::kPtrWords;
	const size_t row_stride = (size_t) S.stride * G::kPtrWords;
	for (int qw = 0; qw < nqw; ++qw) {
		const uint32_t cur = __ldg(c.rp + qw);
		const uint32_t rdw = __funnelshift_l(prev, cur, 4 * c.sub);
		prev = cur;
		const uint32_t nextw = __ldg(c.wp + qw + G::kWin);
#pragma unroll kUnrollI32<W>
		for (int t = 0; t < 8; ++t) {
			const int rc = (rdw >> (4 * t)) & 7;
			const uint2 tab = lut[rc];
			uint32_t al[G::kAligned];
#pragma unroll
			for (int k = 0; k < G::kAligned; ++k) al[k] = (kUnrollI32<W> == 8 && t == 0) ? win[k] : __funnelshift_r(win[k], win[k + 1], 4 * t);
			int left = SENT;
			int rowbest = 0;
			uint32_t pw[G::kPtrWords];
#pragma unroll
			for (int k = 0; k < G::kPtrWords; ++k) pw[k] = 0;
#pragma unroll
			for (int m = 0; m < G::kGroups; ++m) {
				const uint32_t sel = (m & 1) ? (al[m >> 1] >> 16) : al[m >> 1];
				const uint32_t sw = prmt(tab.x, tab.y, sel);
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					const int j = 4 * m + i;
					const int s = i == 0 ? sbyte<0>(sw) : i == 1 ? sbyte<1>(sw) : i == 2 ? sbyte<2>(sw) : sbyte<3>(sw);
					const int d = line[j] + s;
					const int ug = line[j + 1] + gap_read;
					const int lg = left + gap_ref;
					// The pointer is decided on the un-clamped maximum: cells clamped to 0 are STOP
					// cells whose pointer is never followed.  (Deciding it on the .RELU result lets
					// nvcc 12.9 fuse the compares into VIMNMX.RELU predicate outputs, which came out
					// wrong on sm_100a -- found by the parity tests, see DESIGN.md.)
					const int m0 = __vimax3_s32(lg, d, ug);
					int mx = MODE == 0 ? max(m0, 0) : m0;
					// priority diag > up (I) > left (D), oclSwScore.cl:80-83
					uint32_t p = (m0 == d) ? PTR_DIAG : ((m0 == ug) ? PTR_UP : PTR_LEFT);
					if (j >= LO) {
						const bool in = j < corridor;
						mx = in ? mx : SENT;
						p = in ? p : PTR_DIAG;
					}
					pw[j >> 4] |= p << (2 * (j & 15));
					if (MODE == 0) rowbest = max(rowbest, mx * 256 + (255 - j));      // capacity <= 160 < 256
					left = mx;
					line[j] = mx;
				}
			}
#pragma unroll
			for (int k = 0; k < G::kPtrWords; ++k) prow[k] = pw[k];
			prow += row_stride;
			if (MODE == 0) {
				const int rm = rowbest >> 8;
				if (rm > curr_max) {             // first strict maximum in row-major order (oclSwScore.cl:88-91)
					curr_max = rm;
					best_read = rcount;
					best_ref = 255 - (rowbest & 255);
				}
			}
			rcount += (rc != kCodeNul);
		}
#pragma unroll
		for (int k = 0; k + 1 < G::kWin; ++k) win[k] = win[k + 1];
		win[G::kWin - 1] = nextw;
	}
	read_count = rcount;
	if (MODE == 1) {
		int cm = kEndFreeMin;
		best_ref = 0;
#pragma unroll
		for (int j = 0; j < W; ++j) {                // first strict maximum of the final row (oclEndFreeScore.cl:135-140)
			const bool gt = (j < corridor) && line[j] > cm;
			cm = gt ? line[j] : cm;
			best_ref = gt ? j : best_ref;
		}
		best_read = rcount - 1;
		curr_max = cm;
	}
	best_score = curr_max;
}

// oclSW_Backtracking (oclSwCigar.cl:60-124): walk the pointers from the best cell,
// run-length encode the ops onto a stack.  STOP cells are not stored: in local mode a
// cell is STOP iff its score is <= 0 (oclSwScore.cl:84) and the score along the path is
// recomputed exactly from the moves; matrix row 0 and the band borders are STOP (local)
// or X (end-free, oclEndFreeScore.cl:94,129).
template <int W, int MODE>
__device__ __forceinline__ TraceOut backtrace_u16(const DevParams &P, const uint2 *s_lut, const PairCtx &c, const AlignScratch &S, int slot,
		int best_read, int best_ref, int best_score, int read_count) {
	using G = BandGeom<W>;
	TraceOut o;
	o.qstart = 0;
	o.qend = 0;
	o.sp = 0;
	if (best_read <= 0) {
		o.ok = 0;
		o.pos = best_read;
		return o;
	}
	o.ok = 1;
	const int corridor = P.corridor;
	int row = best_read, col = best_ref, abs_ref = best_ref + best_read;
	int h = best_score;
	const int qend = MODE == 0 ? read_count - best_read - 1 : 0;
	int elem = OP_S, len = qend, sp = 0;
	uint16_t *ops = S.ops + slot;
	const uint32_t *pbase = S.ptr + (size_t) slot * G::kPtrWords;
	const size_t row_stride = (size_t) S.stride * G::kPtrWords;
	while (true) {
		if (row < 0) break;                                         // matrix row 0
		const bool border = col < 0 || col >= corridor;
		int op;
		if (border) {
			if (MODE == 0) break;
			op = OP_X;
			row -= 1;
			abs_ref -= 1;
		} else {
			if (MODE == 0 && h <= 0) break;
			const uint32_t pwd = pbase[(size_t) (row + c.sub) * row_stride + (col >> 4)];
			const uint32_t p = (pwd >> (2 * (col & 15))) & 3u;
			if (p == PTR_DIAG) {
				const int rc = code_at(c.rp, row) & 7;
				const int fc = code_at(c.wp, (int64_t) c.sub + row + col) & 7;
				const int s = lut_score(s_lut, c.dir, rc, fc);
				op = (P.alt ? (rc == fc) : (s == P.match)) ? OP_EQ : OP_X;
				h -= s;
				row -= 1;
				abs_ref -= 1;
			} else if (p == PTR_UP) {
				op = OP_I;
				h -= P.gap_read;
				row -= 1;
				col += 1;
			} else {
				op = OP_D;
				h -= P.gap_ref;
				col -= 1;
				abs_ref -= 1;
			}
		}
		if (op == elem) {
			len += 1;
		} else {
			if (sp < S.ops_cap) ops[(size_t) sp * S.stride] = (uint16_t) (len << 4 | elem);
			sp += 1;
			elem = op;
			len = 1;
		}
	}
	if (sp < S.ops_cap) ops[(size_t) sp * S.stride] = (uint16_t) (len << 4 | elem);
	sp += 1;
	o.sp = sp;
	o.pos = abs_ref + 1;
	o.qstart = row + 1;
	o.qend = qend;
	return o;
}

}  // namespace ngm
