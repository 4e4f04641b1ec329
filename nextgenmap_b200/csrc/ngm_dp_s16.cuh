// ngm_dp_s16.cuh -- score-only banded DP with TWO (read, window) pairs per thread
// packed as s16x2 lanes of one 32-bit register (pair A = low half, pair B = high half).
//
// Blackwell executes the whole recurrence on native packed-halfword DPX instructions:
//     d = VIADD.16x2(diag, S)             diag + substitution score
//     u = VIADDMNMX.S16x2(up, gap, d)     max(up + gap_read, d)
//     h = VIADDMNMX.S16x2[.RELU](left, gap, u)
//     best = VIMNMX3.S16x2(best, h0, h1)
// and the substitution scores of both pairs come out of the same byte-LUT PRMT trick as
// the int32 kernels, merged and sign-extended to halfwords by one more PRMT.
// The two halves never interact; each sees NUL rows (score row of zeros) before its own
// first row and after its own last row, which do not change its result (see
// ngm_dp_i32.cuh).  Exactness needs every cell value to fit int16 -- checked on the host
// (build_params: use_s16); otherwise the int32 kernels are used.
#pragma once

#include "ngm_common.cuh"
#include "ngm_dp_i32.cuh"

// rows per unrolled loop body of the score kernel (8 = the whole read word; smaller bodies relieve the instruction cache)
#ifndef NGM_SCORE_ROW_UNROLL
#define NGM_SCORE_ROW_UNROLL 8
#endif
#define NGM_SPRAGMA_(x) _Pragma(#x)
#define NGM_SUNROLL_N(n) NGM_SPRAGMA_(unroll n)

namespace ngm {

// byte I of a (low half) and byte I of b (high half), each sign-extended to 16 bits
template <int I>
__device__ __forceinline__ uint32_t sbyte2(uint32_t a, uint32_t b) {
	constexpr uint32_t sel = (uint32_t) I | ((8u | I) << 4) | ((4u + I) << 8) | ((12u + I) << 12);
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "n"(sel));
	return d;
}

__device__ __forceinline__ uint32_t pack2(int lo, int hi) { return ((uint32_t) hi << 16) | ((uint32_t) lo & 0xFFFFu); }

// Shared selectors (-DNGM_SCORE_SHARED_SEL=0: every row aligns its own window).  With the eight rows of a window word unrolled, row t reads
// nibble j + t of the window as it sits in the registers: the rows share the sixteen-bit selector halves of the window words, no row needs a
// funnel shift, and slot j of row t takes byte (j + t) & 3 of group (j + t) >> 2.  EXACT: the corridor is LO exactly (compile-time): slots
// >= LO are never computed and no slot carries the run-time corridor select.
// rows per loop body on wide bands (capacity > 48): 250 bp / corridor 80 24.7 ms (8 rows) -> 22.4 ms (2 rows); 400 bp 66.5 -> 61.8 ms
#ifndef NGM_SCORE_WIDE_UNROLL
#define NGM_SCORE_WIDE_UNROLL 2
#endif
#ifndef NGM_SCORE_SHARED_SEL
#define NGM_SCORE_SHARED_SEL (NGM_SCORE_ROW_UNROLL == 8)
#endif

template <int W, int LO, int MODE, bool EXACT = false>
__global__ void __launch_bounds__(128) score_s16_kernel(const __grid_constant__ DevParams P, const PairDesc *__restrict__ pairs, int n,
		const uint32_t *__restrict__ reads_fwd, const uint32_t *__restrict__ reads_rev, const uint16_t *__restrict__ rlen,
		const uint32_t *__restrict__ ref4, float *__restrict__ out, const int *__restrict__ sel, const int *__restrict__ n_dev) {
	using G = BandGeom<W>;
	__shared__ uint2 s_lut[16];
	if (threadIdx.x < 16) s_lut[threadIdx.x] = P.lut[threadIdx.x];
	__syncthreads();
	if (n_dev != nullptr) n = min(n, *n_dev);
	const int t2 = blockIdx.x * blockDim.x + threadIdx.x;
	if (2 * t2 >= n) return;
	const bool has_b = 2 * t2 + 1 < n;
	const int ia = sel != nullptr ? sel[2 * t2] : 2 * t2;
	const int ib = has_b ? (sel != nullptr ? sel[2 * t2 + 1] : ia + 1) : ia;
	constexpr int SENT = MODE == 0 ? 0 : kEndFreeMin;
	const uint32_t SENT2 = pack2(SENT, SENT);
	PairCtx ca, cb;
	uint32_t fa, fb;
	const bool act_a = load_pair(P, pairs, ia, reads_fwd, reads_rev, rlen, ref4, ca, fa);
	const bool act_b = load_pair(P, pairs, ib, reads_fwd, reads_rev, rlen, ref4, cb, fb);
	const int corridor = P.corridor;
	const uint32_t gr2 = pack2(P.gap_read, P.gap_read), gf2 = pack2(P.gap_ref, P.gap_ref);
	uint32_t line[W + 1];
#pragma unroll
	for (int j = 0; j <= W; ++j) line[j] = (j < corridor) ? 0u : SENT2;
	uint32_t best = 0;
	uint32_t wa[G::kWin], wb[G::kWin];
#pragma unroll
	for (int k = 0; k < G::kWin; ++k) {
		wa[k] = __ldg(ca.wp + k);
		wb[k] = __ldg(cb.wp + k);
	}
	const int nqw = max((ca.sub + ca.len + 7) >> 3, (cb.sub + cb.len + 7) >> 3);
	const uint2 *luta = s_lut + ca.dir * 8, *lutb = s_lut + cb.dir * 8;
	uint32_t prev_a = kNulWord, prev_b = kNulWord;
	for (int qw = 0; qw < nqw; ++qw) {
		const uint32_t cur_a = __ldg(ca.rp + qw), cur_b = __ldg(cb.rp + qw);
		const uint32_t rda = __funnelshift_l(prev_a, cur_a, 4 * ca.sub);
		const uint32_t rdb = __funnelshift_l(prev_b, cur_b, 4 * cb.sub);
		prev_a = cur_a;
		prev_b = cur_b;
		const uint32_t next_a = __ldg(ca.wp + qw + G::kWin), next_b = __ldg(cb.wp + qw + G::kWin);
		// wide bands: the eight-row body no longer fits the instruction cache (W = 80: ~50 KB); NGM_SCORE_WIDE_UNROLL rows per body there
		constexpr int kUnrollS = W > 48 ? NGM_SCORE_WIDE_UNROLL : NGM_SCORE_ROW_UNROLL;
		if constexpr (NGM_SCORE_SHARED_SEL && kUnrollS == 8) {
		constexpr int NSLOT = EXACT ? LO : W;
		constexpr int NSEL = (NSLOT + 7 + 3) / 4;                  // selectors the eight rows touch (nibbles 0 .. NSLOT + 6)
		uint32_t sela[NSEL], selb[NSEL];
#pragma unroll
		for (int g = 0; g < NSEL; ++g) {
			sela[g] = (g & 1) ? (wa[g >> 1] >> 16) : wa[g >> 1];
			selb[g] = (g & 1) ? (wb[g >> 1] >> 16) : wb[g >> 1];
		}
#pragma unroll
		for (int t = 0; t < 8; ++t) {
			const uint2 ta = luta[(rda >> (4 * t)) & 7];
			const uint2 tb = lutb[(rdb >> (4 * t)) & 7];
			uint32_t left = SENT2;
#pragma unroll
			for (int g = t / 4; g <= (NSLOT - 1 + t) / 4; ++g) {
				const uint32_t sa = prmt(ta.x, ta.y, sela[g]);
				const uint32_t sb = prmt(tb.x, tb.y, selb[g]);
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					const int j = 4 * g + i - t;
					if (j < 0 || j >= NSLOT) continue;
					const uint32_t s2 = i == 0 ? sbyte2<0>(sa, sb) : i == 1 ? sbyte2<1>(sa, sb) : i == 2 ? sbyte2<2>(sa, sb) : sbyte2<3>(sa, sb);
					const uint32_t d = __vadd2(line[j], s2);
					const uint32_t u = __viaddmax_s16x2(line[j + 1], gr2, d);
					uint32_t h = MODE == 0 ? __viaddmax_s16x2_relu(left, gf2, u) : __viaddmax_s16x2(left, gf2, u);
					if (!EXACT && j >= LO) h = (j < corridor) ? h : SENT2;
					left = h;
					line[j] = h;
				}
			}
			if (MODE == 0) {
#pragma unroll
				for (int j = 0; j + 1 < NSLOT; j += 2) best = __vimax3_s16x2(best, line[j], line[j + 1]);
				if (NSLOT & 1) best = __vmaxs2(best, line[NSLOT - 1]);
			}
		}
		} else {
#pragma unroll kUnrollS
		for (int t = 0; t < 8; ++t) {
			const uint2 ta = luta[(rda >> (4 * t)) & 7];
			const uint2 tb = lutb[(rdb >> (4 * t)) & 7];
			uint32_t ala[G::kAligned], alb[G::kAligned];
#pragma unroll
			for (int k = 0; k < G::kAligned; ++k) {
				ala[k] = (kUnrollS == 8 && t == 0) ? wa[k] : __funnelshift_r(wa[k], wa[k + 1], 4 * t);
				alb[k] = (kUnrollS == 8 && t == 0) ? wb[k] : __funnelshift_r(wb[k], wb[k + 1], 4 * t);
			}
			uint32_t left = SENT2;
#pragma unroll
			for (int m = 0; m < G::kGroups; ++m) {
				const uint32_t sa = prmt(ta.x, ta.y, (m & 1) ? (ala[m >> 1] >> 16) : ala[m >> 1]);
				const uint32_t sb = prmt(tb.x, tb.y, (m & 1) ? (alb[m >> 1] >> 16) : alb[m >> 1]);
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					const int j = 4 * m + i;
					const uint32_t s2 = i == 0 ? sbyte2<0>(sa, sb) : i == 1 ? sbyte2<1>(sa, sb) : i == 2 ? sbyte2<2>(sa, sb) : sbyte2<3>(sa, sb);
					const uint32_t d = __vadd2(line[j], s2);
					const uint32_t u = __viaddmax_s16x2(line[j + 1], gr2, d);
					uint32_t h = MODE == 0 ? __viaddmax_s16x2_relu(left, gf2, u) : __viaddmax_s16x2(left, gf2, u);
					if (j >= LO) h = (j < corridor) ? h : SENT2;
					left = h;
					line[j] = h;
				}
			}
			if (MODE == 0) {
#pragma unroll
				for (int j = 0; j + 1 < W; j += 2) best = __vimax3_s16x2(best, line[j], line[j + 1]);
				if (W & 1) best = __vmaxs2(best, line[W - 1]);
			}
		}
		}
#pragma unroll
		for (int k = 0; k + 1 < G::kWin; ++k) {
			wa[k] = wa[k + 1];
			wb[k] = wb[k + 1];
		}
		wa[G::kWin - 1] = next_a;
		wb[G::kWin - 1] = next_b;
	}
	if (MODE == 1) {
		best = SENT2;
#pragma unroll
		for (int j = 0; j < W; ++j) best = __vmaxs2(best, line[j]);
	}
	const float inactive = MODE == 0 ? -1.0f : (float) kEndFreeMin;
	const float ra = act_a ? (float) (int) (short) (best & 0xFFFFu) : inactive;
	const float rb = act_b ? (float) (int) (short) (best >> 16) : inactive;
	if (has_b && sel == nullptr) {
		*reinterpret_cast<float2 *>(out + ia) = make_float2(ra, rb);
	} else {
		out[ia] = ra;
		if (has_b) out[ib] = rb;
	}
}

}  // namespace ngm
