// ngm_align_s16v2.cuh -- second-generation forward pass of the tagged s16x2 align path (narrow bands).
//
// Same arithmetic, pointer-matrix layout and results as align_s16_fwd_kernel (ngm_align_s16.cuh); what changed
// is where the instructions go.  The first-generation kernel is bound by the integer ALU pipe (64 lanes / clk / SM:
// PRMT, LOP3, VIADD, VIADDMNMX, VIMNMX3, SEL all share it) while the FMA pipe idles:
//
//   * the tag arithmetic runs on the FMA pipe: tag = h - clean and pw = 4 * pw + tag are issued as IMADs whose
//     multipliers (-1, 4) come from the kernel parameters, so ptxas cannot strength-reduce them back into
//     LOP3 / LEA on the ALU pipe;
//   * local mode no longer keeps a per-row snapshot of the band in registers (one PRMT per band register and row).
//     Instead the band is CHECKPOINTED to shared memory once per eight rows (one STS per band register and eight
//     rows, LSU pipe), the running maximum is tracked per block of eight rows only, and after the last row the
//     eight rows of the block in which each half's maximum first appeared are REPLAYED from that block's checkpoint
//     to find the best cell (first strict maximum in row-major order, oclSwScore.cl:88-91).  Three checkpoint
//     buffers rotate per thread: one owned by each half (the block of its last improvement) and one being written.
//
// Reference semantics: oclSW_Score / oclSW_ScoreGlobal (oclSwScore.cl:4-107, oclEndFreeScore.cl:58-146).
#pragma once

#include "ngm_align_s16.cuh"

// Rows per unrolled loop body.  Fully unrolled (8) the main loop alone is ~36 KB of SASS and the kernel 83 KB: ncu showed `no_instruction`
// (instruction fetch) as the top stall reason at 16 resident warps per SM.  One row per body is ~4.5 KB; the shift amounts and the
// read-code extraction simply take the row index from a register.
#ifndef NGM_FWD_ROW_UNROLL
#define NGM_FWD_ROW_UNROLL 2      /* measured on B200, 10 M x 150 bp: 8 rows 12.0 ms, 2 rows 10.1 ms, 1 row 10.6 ms */
#endif
#define NGM_PRAGMA_(x) _Pragma(#x)
#define NGM_UNROLL_N(n) NGM_PRAGMA_(unroll n)

namespace ngm {

// Window alignment on the FMA pipe (experiment, -DNGM_FUNNEL_FMA=1): (lo >> s) | (hi << (32 - s)) = mulhi(lo, 2^(32-s)) + hi * 2^(32-s), two
// FMA-pipe instructions instead of one ALU-pipe SHF; the multiplier comes from a run-time table so that ptxas keeps the multiplies.
#ifndef NGM_FUNNEL_FMA
#define NGM_FUNNEL_FMA 0
#endif
// The odd nibble groups' `>> 16` alone as IMAD.HI (x >> 16 = mulhi(x, 2^16)): -DNGM_HI16_FMA=1
#ifndef NGM_HI16_FMA
#define NGM_HI16_FMA NGM_FUNNEL_FMA
#endif
// Offset-binary band (EXPERIMENT, local mode, -DNGM_FWD_OFFSET=1): every band value is kept as true + 0x8000 per half and compared
// UNSIGNED.  A 32-bit IMAD that adds a negative per-half constant to such a word then ALWAYS carries out of the low half (the stored low
// half is >= 0x8000 > |4 * gap|), so the carry is a constant that the addend's high half absorbs (c - 0x10000): the up and left candidates
// become IMADs and the floor of local alignment the third operand of one VIMNMX3.U16x2 -- VIADDMNMX.U16x2 + VIMNMX3.U16x2 + 2 IMAD instead
// of VIADD.16x2 + 2 x VIADDMNMX.S16x2.  Bit-identical results (all GPU tests pass), but measured SLOWER on B200 (10 M x 150 bp: forward pass
// 10.42 vs 9.99 ms): scripts/issue_rates.py shows why -- VIADD.16x2 already issues on the FMA pipe, and VIMNMX3 costs the integer pipe the same
// slot as VIADDMNMX, so the variant only adds FMA-pipe work (profiles/r2_issue_rates.md).  Off by default.
#ifndef NGM_FWD_OFFSET
#define NGM_FWD_OFFSET 0
#endif
constexpr uint32_t kOfs2 = 0x80008000u;
// per-half constant c (two's complement) as the addend of an always-carrying 32-bit add
__device__ __forceinline__ uint32_t carry_addend(int c) { return pack2(c, c) - 0x10000u; }

__device__ __forceinline__ uint32_t funnel_fma(uint32_t lo, uint32_t hi, uint32_t mult) {
	uint32_t a, d;
	asm("mul.hi.u32 %0, %1, %2;" : "=r"(a) : "r"(lo), "r"(mult));
	asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(mult), "r"(a));
	return d;
}

// One DP row for both halves.  PTR: accumulate + return the pointer words of the row.  MAYBE0: t may be 0 (no shift).
// EXACT: the corridor is LO exactly (a compile-time constant): slots >= LO are never computed (they keep the sentinel; their pointer tags
// are zero shifts of the pointer word) and no slot needs the run-time corridor mask.
template <int W, int LO, int MODE, bool PTR, bool MAYBE0 = true, bool EXACT = false>
__device__ __forceinline__ void fwd2_row(uint32_t (&line)[W + 1], const uint32_t (&wa)[BandGeom<W>::kWin], const uint32_t (&wb)[BandGeom<W>::kWin],
		const int t, const uint2 ta, const uint2 tb, const uint32_t gr2, const uint32_t gf2, const uint32_t SENT2,
		const uint32_t (&keepm)[W - LO + 1], const uint32_t (&fillm)[W - LO + 1], const uint32_t c_four, const uint32_t c_neg1,
		uint32_t (&pw)[TagGeom<W>::kWords]) {
	using G = BandGeom<W>;
	constexpr bool OFS = MODE == 0 && NGM_FWD_OFFSET != 0;        // gr2 / gf2 then hold carry_addend()s and SENT2 = kOfs2 (see above)
	const uint32_t c_one = c_four >> 2;
	uint32_t ala[G::kAligned], alb[G::kAligned];
#pragma unroll
	for (int k = 0; k < G::kAligned; ++k) {
#if NGM_FUNNEL_FMA
		const uint32_t mult = c_four << ((30 - 4 * t) & 31);         // 2^(32 - 4t) for t >= 1 (c_four = 4, a run-time value)
		ala[k] = (MAYBE0 && t == 0) ? wa[k] : funnel_fma(wa[k], wa[k + 1], mult);
		alb[k] = (MAYBE0 && t == 0) ? wb[k] : funnel_fma(wb[k], wb[k + 1], mult);
#else
		ala[k] = __funnelshift_r(wa[k], wa[k + 1], 4 * t);          // t is a run-time value: the row loop is NOT fully unrolled (see NGM_FWD_ROW_UNROLL)
		alb[k] = __funnelshift_r(wb[k], wb[k + 1], 4 * t);
#endif
	}
	uint32_t left = SENT2;
	if (PTR) {
#pragma unroll
		for (int k = 0; k < TagGeom<W>::kWords; ++k) pw[k] = 0;
	}
#pragma unroll
	for (int m = 0; m < G::kGroups; ++m) {
#if NGM_HI16_FMA
		uint32_t hia = 0, hib = 0;
		if (m & 1) {                                                   // x >> 16 = mulhi(x, 2^16): FMA pipe
			asm("mul.hi.u32 %0, %1, %2;" : "=r"(hia) : "r"(ala[m >> 1]), "r"(c_four << 14));
			asm("mul.hi.u32 %0, %1, %2;" : "=r"(hib) : "r"(alb[m >> 1]), "r"(c_four << 14));
		}
		const uint32_t sa = prmt(ta.x, ta.y, (m & 1) ? hia : ala[m >> 1]);
		const uint32_t sb = prmt(tb.x, tb.y, (m & 1) ? hib : alb[m >> 1]);
#else
		const uint32_t sa = prmt(ta.x, ta.y, (m & 1) ? (ala[m >> 1] >> 16) : ala[m >> 1]);
		const uint32_t sb = prmt(tb.x, tb.y, (m & 1) ? (alb[m >> 1] >> 16) : alb[m >> 1]);
#endif
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const int j = 4 * m + i;
			if (EXACT && j >= LO) {
				if (PTR) pw[j >> 3] = imad_u32(pw[j >> 3], c_four, 0u);
				continue;
			}
			const uint32_t s2 = i == 0 ? sbyte2<0>(sa, sb) : i == 1 ? sbyte2<1>(sa, sb) : i == 2 ? sbyte2<2>(sa, sb) : sbyte2<3>(sa, sb);
			uint32_t h;
			if (OFS) {
				const uint32_t up = imad_u32(line[j + 1], c_one, gr2);    // FMA pipe
				const uint32_t lf = imad_u32(left, c_one, gf2);           // FMA pipe
				const uint32_t u = __viaddmax_u16x2(line[j], s2, up);
				h = __vimax3_u16x2(u, lf, kOfs2);
			} else {
				const uint32_t d = __vadd2(line[j], s2);
				const uint32_t u = __viaddmax_s16x2(line[j + 1], gr2, d);
				h = MODE == 0 ? __viaddmax_s16x2_relu(left, gf2, u) : __viaddmax_s16x2(left, gf2, u);
			}
			// slots at or beyond the corridor are pinned to the sentinel: one LOP3 with loop-invariant masks instead of compare + select
			if (!EXACT && j >= LO) h = (h & keepm[j - LO]) | ((MODE == 0 && !OFS) ? 0u : fillm[j - LO]);
			const uint32_t clean = h & 0xFFFCFFFCu;
			if (PTR) {
				const uint32_t tag = imad_u32(clean, c_neg1, h);          // h - clean, FMA pipe
				pw[j >> 3] = imad_u32(pw[j >> 3], c_four, tag);           // 4 * pw + tag, FMA pipe
			}
			left = clean;
			line[j] = clean;
		}
	}
}

// Row pairs on one window alignment (-DNGM_FWD_PAIR_ROWS=0: every row aligns its own window).  Rows t (even) and t + 1 read the same
// reference nibbles, one slot apart: slot j of row t + OFF is nibble j + OFF of the window aligned for row t.  The pair therefore shares the
// funnel shifts and the `>> 16` of the odd groups (12 ALU-pipe instructions per row at W = 28); the odd row pays with one more PRMT group.
#ifndef NGM_FWD_PAIR_ROWS
#define NGM_FWD_PAIR_ROWS 1
#endif
template <int W>
struct PairGeom {
	static constexpr int kNibbles = W + 1;                       // nibbles the two rows touch
	static constexpr int kWords = (kNibbles + 7) / 8;            // aligned words
	static constexpr int kSel = (kNibbles + 3) / 4;              // 16-bit PRMT selectors (groups of four nibbles)
};

// selectors of the window aligned for row t: sel[g] carries nibbles 4g .. 4g + 3 in its low 16 bits
template <int W>
__device__ __forceinline__ void fwd2_selectors(const uint32_t (&w)[BandGeom<W>::kWin], int t, uint32_t (&sel)[PairGeom<W>::kSel]) {
	using PG = PairGeom<W>;
#pragma unroll
	for (int k = 0; k < PG::kWords; ++k) {
		const uint32_t hi = k + 1 < BandGeom<W>::kWin ? w[k + 1] : 0u;      // (only the low nibbles of the last word are used)
		const uint32_t al = __funnelshift_r(w[k], hi, 4 * t);
		sel[2 * k] = al;
		if (2 * k + 1 < PG::kSel) sel[2 * k + 1] = al >> 16;
	}
}

// fwd2_row on shared selectors; OFF = 0 / 1: the row's distance from the row the window was aligned for
template <int W, int LO, int MODE, bool PTR, bool EXACT, int OFF>
__device__ __forceinline__ void fwd2_row_sel(uint32_t (&line)[W + 1], const uint32_t (&sela)[PairGeom<W>::kSel], const uint32_t (&selb)[PairGeom<W>::kSel],
		const uint2 ta, const uint2 tb, const uint32_t gr2, const uint32_t gf2, const uint32_t SENT2, const uint32_t (&keepm)[W - LO + 1],
		const uint32_t (&fillm)[W - LO + 1], const uint32_t c_four, const uint32_t c_neg1, uint32_t (&pw)[TagGeom<W>::kWords]) {
	constexpr int NSLOT = EXACT ? LO : W;
	uint32_t left = SENT2;
	if (PTR) {
#pragma unroll
		for (int k = 0; k < TagGeom<W>::kWords; ++k) pw[k] = 0;
	}
#pragma unroll
	for (int g = 0; g <= (NSLOT - 1 + OFF) / 4; ++g) {
		const uint32_t sa = prmt(ta.x, ta.y, sela[g]);
		const uint32_t sb = prmt(tb.x, tb.y, selb[g]);
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const int j = 4 * g + i - OFF;
			if (j < 0 || j >= NSLOT) continue;
			const uint32_t s2 = i == 0 ? sbyte2<0>(sa, sb) : i == 1 ? sbyte2<1>(sa, sb) : i == 2 ? sbyte2<2>(sa, sb) : sbyte2<3>(sa, sb);
			const uint32_t d = __vadd2(line[j], s2);
			const uint32_t u = __viaddmax_s16x2(line[j + 1], gr2, d);
			uint32_t h = MODE == 0 ? __viaddmax_s16x2_relu(left, gf2, u) : __viaddmax_s16x2(left, gf2, u);
			if (!EXACT && j >= LO) h = (h & keepm[j - LO]) | (MODE == 0 ? 0u : fillm[j - LO]);
			const uint32_t clean = h & 0xFFFCFFFCu;
			if (PTR) {
				const uint32_t tag = imad_u32(clean, c_neg1, h);          // h - clean, FMA pipe
				pw[j >> 3] = imad_u32(pw[j >> 3], c_four, tag);           // 4 * pw + tag, FMA pipe
			}
			left = clean;
			line[j] = clean;
		}
	}
	if (PTR && EXACT) {                                           // slots LO .. W - 1 are never computed: their tags are zero shifts
#pragma unroll
		for (int j = LO; j < W; ++j) pw[j >> 3] = imad_u32(pw[j >> 3], c_four, 0u);
	}
}

// resident blocks per SM the register allocation must allow (1 = no constraint beyond the 255-register limit)
#ifndef NGM_FWD2_MIN_BLOCKS
#define NGM_FWD2_MIN_BLOCKS(W) 1
#endif

template <int W, bool OFS = false, int N = W>             // N: slots that can hold a value
__device__ __forceinline__ uint32_t band_max(const uint32_t (&line)[W + 1], uint32_t acc) {
#pragma unroll
	for (int j = 0; j + 1 < N; j += 2) acc = OFS ? __vimax3_u16x2(acc, line[j], line[j + 1]) : __vimax3_s16x2(acc, line[j], line[j + 1]);
	if (N & 1) acc = OFS ? __vmaxu2(acc, line[N - 1]) : __vmaxs2(acc, line[N - 1]);
	return acc;
}

template <int W, int LO, int MODE, bool EXACT = false>
__global__ void __launch_bounds__(128, NGM_FWD2_MIN_BLOCKS(W)) align_s16_fwd2_kernel(const __grid_constant__ DevParams P, const PairDesc *__restrict__ pairs, int n,
		const uint32_t *__restrict__ reads_fwd, const uint32_t *__restrict__ reads_rev, const uint16_t *__restrict__ rlen,
		const uint32_t *__restrict__ ref4, uint32_t *__restrict__ ptr_scratch, int stride, int4 *__restrict__ best_out,
		const int *__restrict__ range, int range_m) {
	using G = BandGeom<W>;
	using T = TagGeom<W>;
	extern __shared__ uint32_t s_chk[];                           // local mode: [3][W][128] checkpoint words, one column per thread
	__shared__ uint2 s_lut4[16];
	if (threadIdx.x < 16) s_lut4[threadIdx.x] = P.lut4[threadIdx.x];
	__syncthreads();
	// range != nullptr: the launch covers pairs[range[0] .. range[range_m]) -- offsets that only the device knows (ngm_batch.cu: the
	// forward pass over every candidate of a chunk of reads); slots, pointer words and best_out stay relative to the chunk
	if (range != nullptr) {
		const int base = range[0];
		n = min(n, range[range_m] - base);
		pairs += base;
	}
	const int t2 = blockIdx.x * blockDim.x + threadIdx.x;        // thread slot: pairs 2*t2, 2*t2 + 1
	const int ia = 2 * t2;
	if (ia >= n) return;
	const bool vb = ia + 1 < n;
	const int ib = vb ? ia + 1 : ia;
	constexpr bool OFS = MODE == 0 && NGM_FWD_OFFSET != 0;
	constexpr int WE = EXACT ? LO : W;                            // slots that are ever computed
	constexpr int SENT = MODE == 0 ? 0 : 4 * kEndFreeMinS16;
	const uint32_t SENT2 = OFS ? kOfs2 : pack2(SENT, SENT);
	const uint32_t ZERO2 = OFS ? kOfs2 : 0u;                      // a cell value of 0 as stored
	auto half_lo = [](uint32_t v) { return OFS ? (int) (v & 0xFFFFu) - 0x8000 : (int) (short) (v & 0xFFFFu); };
	auto half_hi = [](uint32_t v) { return OFS ? (int) (v >> 16) - 0x8000 : (int) (short) (v >> 16); };
	PairCtx ca, cb;
	uint32_t fa, fb;
	const bool act_a = load_pair(P, pairs, ia, reads_fwd, reads_rev, rlen, ref4, ca, fa);
	const bool act_b = load_pair(P, pairs, ib, reads_fwd, reads_rev, rlen, ref4, cb, fb) && vb;
	const int corridor = P.corridor;
	const uint32_t gr2 = OFS ? carry_addend(4 * P.gap_read + 1) : pack2(4 * P.gap_read + 1, 4 * P.gap_read + 1);
	const uint32_t gf2 = OFS ? carry_addend(4 * P.gap_ref) : pack2(4 * P.gap_ref, 4 * P.gap_ref);
	const uint32_t c_four = P.c_four, c_neg1 = P.c_neg1;
	const int tstride = stride >> 1;
	const size_t row_stride = (size_t) tstride * T::kWords;
	uint32_t *prow = ptr_scratch + (size_t) t2;
	uint32_t *chk = s_chk + threadIdx.x;
	// wide local bands (capacity > 48): three checkpoints of W words per thread do not fit shared memory at a useful occupancy; they live in
	// local memory instead (L1 / L2 resident: 3 x W x 4 bytes per thread, written once per eight rows)
	constexpr bool CHK_LOCAL = MODE == 0 && W > kAlignS16MaxLocal;
	uint32_t chk_local[CHK_LOCAL ? 3 : 1][CHK_LOCAL ? W : 1];
	auto chk_at = [&](int buf, int j) -> uint32_t * { return CHK_LOCAL ? &chk_local[CHK_LOCAL ? buf : 0][CHK_LOCAL ? j : 0] : chk + (buf * W + j) * 128; };

	uint32_t line[W + 1];
#pragma unroll
	for (int j = 0; j <= W; ++j) line[j] = (j < corridor) ? ZERO2 : SENT2;
	uint32_t keepm[W - LO + 1], fillm[W - LO + 1];
#pragma unroll
	for (int i = 0; i <= W - LO; ++i) {
		keepm[i] = (LO + i < corridor) ? 0xFFFFFFFFu : 0u;
		fillm[i] = (LO + i < corridor) ? 0u : SENT2;
	}
	uint32_t best = ZERO2;
	int rc_a = 0, rc_b = 0;
	// checkpoint bookkeeping (local mode): per half the buffer, block and row count of the block of its last improvement
	int own_a = 0, own_b = 0, blk_a = 0, blk_b = 0, rc0_a = 0, rc0_b = 0;
	uint32_t wa[G::kWin], wb[G::kWin];
	const int wend_a = ca.sub + ca.len + corridor - 1, wend_b = cb.sub + cb.len + corridor - 1;
	auto nib_mask = [](int valid) { return valid >= 8 ? 0xFFFFFFFFu : (valid <= 0 ? 0u : ((1u << (4 * valid)) - 1u)); };
	uint32_t spec_a = 0, spec_b = 0;
#pragma unroll
	for (int k = 0; k < G::kWin; ++k) {
		wa[k] = __ldg(ca.wp + k);
		wb[k] = __ldg(cb.wp + k);
		spec_a |= wa[k] & nib_mask(wend_a - 8 * k);
		spec_b |= wb[k] & nib_mask(wend_b - 8 * k);
	}
	const int nqw = max(act_a ? (ca.sub + ca.len + 7) >> 3 : 0, act_b ? (cb.sub + cb.len + 7) >> 3 : 0);
	const uint2 *luta = s_lut4 + ca.dir * 8, *lutb = s_lut4 + cb.dir * 8;
	uint32_t prev_a = kNulWord, prev_b = kNulWord;
	for (int qw = 0; qw < nqw; ++qw) {
		const uint32_t cur_a = __ldg(ca.rp + qw), cur_b = __ldg(cb.rp + qw);
		const uint32_t rda = __funnelshift_l(prev_a, cur_a, 4 * ca.sub);
		const uint32_t rdb = __funnelshift_l(prev_b, cur_b, 4 * cb.sub);
		prev_a = cur_a;
		prev_b = cur_b;
		const uint32_t next_a = __ldg(ca.wp + qw + G::kWin), next_b = __ldg(cb.wp + qw + G::kWin);
		spec_a |= (next_a & nib_mask(wend_a - 8 * (qw + G::kWin))) | (cur_a & nib_mask(ca.len - 8 * qw));
		spec_b |= (next_b & nib_mask(wend_b - 8 * (qw + G::kWin))) | (cur_b & nib_mask(cb.len - 8 * qw));
		int cur_buf = 0;
		const uint32_t best_before = best;
		const int rcs_a = rc_a, rcs_b = rc_b;
		if (MODE == 0) {
			cur_buf = (own_a != 0 && own_b != 0) ? 0 : ((own_a != 1 && own_b != 1) ? 1 : 2);
#pragma unroll
			for (int j = 0; j < WE; ++j) *chk_at(cur_buf, j) = line[j];
		}
#if NGM_FUNNEL_FMA
#pragma unroll 1
		for (int tt = 0; tt < 8; tt += 2) {
#pragma unroll
			for (int h2 = 0; h2 < 2; ++h2) {
				const int t = tt + h2;
				const int rca = (rda >> (4 * t)) & 7, rcb = (rdb >> (4 * t)) & 7;
				uint32_t pw[T::kWords];
				if (h2 == 0) fwd2_row<W, LO, MODE, true, true, EXACT>(line, wa, wb, t, luta[rca], lutb[rcb], gr2, gf2, SENT2, keepm, fillm, c_four, c_neg1, pw);
				else fwd2_row<W, LO, MODE, true, false, EXACT>(line, wa, wb, t, luta[rca], lutb[rcb], gr2, gf2, SENT2, keepm, fillm, c_four, c_neg1, pw);
#pragma unroll
				for (int k = 0; k < T::kWords; ++k) prow[(size_t) k * tstride] = pw[k];
				prow += row_stride;
				if (MODE == 0) best = band_max<W, OFS, WE>(line, best);
				rc_a += (rca != kCodeNul);
				rc_b += (rcb != kCodeNul);
			}
		}
#elif NGM_FWD_PAIR_ROWS && !NGM_FWD_OFFSET
#pragma unroll 1
		for (int tt = 0; tt < 8; tt += 2) {
			uint32_t sela[PairGeom<W>::kSel], selb[PairGeom<W>::kSel];
			fwd2_selectors<W>(wa, tt, sela);
			fwd2_selectors<W>(wb, tt, selb);
#pragma unroll
			for (int h2 = 0; h2 < 2; ++h2) {
				const int t = tt + h2;
				const int rca = (rda >> (4 * t)) & 7, rcb = (rdb >> (4 * t)) & 7;
				uint32_t pw[T::kWords];
				if (h2 == 0) fwd2_row_sel<W, LO, MODE, true, EXACT, 0>(line, sela, selb, luta[rca], lutb[rcb], gr2, gf2, SENT2, keepm, fillm, c_four, c_neg1, pw);
				else fwd2_row_sel<W, LO, MODE, true, EXACT, 1>(line, sela, selb, luta[rca], lutb[rcb], gr2, gf2, SENT2, keepm, fillm, c_four, c_neg1, pw);
#pragma unroll
				for (int k = 0; k < T::kWords; ++k) prow[(size_t) k * tstride] = pw[k];
				prow += row_stride;
				if (MODE == 0) best = band_max<W, OFS, WE>(line, best);
				rc_a += (rca != kCodeNul);
				rc_b += (rcb != kCodeNul);
			}
		}
#else
NGM_UNROLL_N(NGM_FWD_ROW_UNROLL)
		for (int t = 0; t < 8; ++t) {
			const int rca = (rda >> (4 * t)) & 7, rcb = (rdb >> (4 * t)) & 7;
			uint32_t pw[T::kWords];
			fwd2_row<W, LO, MODE, true, true, EXACT>(line, wa, wb, t, luta[rca], lutb[rcb], gr2, gf2, SENT2, keepm, fillm, c_four, c_neg1, pw);
#pragma unroll
			for (int k = 0; k < T::kWords; ++k) prow[(size_t) k * tstride] = pw[k];
			prow += row_stride;
			if (MODE == 0) best = band_max<W, OFS, WE>(line, best);
			rc_a += (rca != kCodeNul);
			rc_b += (rcb != kCodeNul);
		}
#endif
		if (MODE == 0) {
			const uint32_t imp = best ^ best_before;
			if (imp & 0xFFFFu) {
				own_a = cur_buf;
				blk_a = qw;
				rc0_a = rcs_a;
			}
			if (imp >> 16) {
				own_b = cur_buf;
				blk_b = qw;
				rc0_b = rcs_b;
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
	HalfBest ba, bb;
	ba.read_count = rc_a | ((spec_a & 0x44444444u) ? kSpecialFlag : 0);
	bb.read_count = rc_b | ((spec_b & 0x44444444u) ? kSpecialFlag : 0);
	if (MODE == 0) {
		const int ma = half_lo(best), mb = half_hi(best);
		// ---- replay: the block of each half's last improvement, from its checkpoint, until the half reaches its maximum ----
#pragma unroll
		for (int j = 0; j < WE; ++j) line[j] = prmt(*chk_at(own_a, j), *chk_at(own_b, j), 0x7610u);
#pragma unroll
		for (int k = 0; k < G::kWin; ++k) {
			wa[k] = __ldg(ca.wp + blk_a + k);
			wb[k] = __ldg(cb.wp + blk_b + k);
		}
		const uint32_t pa = blk_a > 0 ? __ldg(ca.rp + blk_a - 1) : kNulWord, pb = blk_b > 0 ? __ldg(cb.rp + blk_b - 1) : kNulWord;
		const uint32_t rda = __funnelshift_l(pa, __ldg(ca.rp + blk_a), 4 * ca.sub);
		const uint32_t rdb = __funnelshift_l(pb, __ldg(cb.rp + blk_b), 4 * cb.sub);
		bool found_a = false, found_b = false;
		int brow_a = 0, brow_b = 0, rr_a = rc0_a, rr_b = rc0_b;
NGM_UNROLL_N(NGM_FWD_ROW_UNROLL)
		for (int t = 0; t < 8; ++t) {
			const int rca = (rda >> (4 * t)) & 7, rcb = (rdb >> (4 * t)) & 7;
			uint32_t pw[T::kWords];
			fwd2_row<W, LO, MODE, false, true, EXACT>(line, wa, wb, t, luta[rca], lutb[rcb], gr2, gf2, SENT2, keepm, fillm, c_four, c_neg1, pw);
			const uint32_t mx = band_max<W, OFS, WE>(line, ZERO2);
			const bool hit_a = !found_a && half_lo(mx) == ma;
			const bool hit_b = !found_b && half_hi(mx) == mb;
			// the band of the row in which a half first reaches its maximum goes to a (by now free) checkpoint buffer
			if (hit_a) {
#pragma unroll
				for (int j = 0; j < WE; ++j) *chk_at(0, j) = line[j];
				brow_a = rr_a;
			}
			if (hit_b) {
#pragma unroll
				for (int j = 0; j < WE; ++j) *chk_at(1, j) = line[j];
				brow_b = rr_b;
			}
			found_a = found_a || hit_a;
			found_b = found_b || hit_b;
			rr_a += (rca != kCodeNul);
			rr_b += (rcb != kCodeNul);
		}
		int ra = 0, rb = 0;
		bool fa_ = false, fb_ = false;
#pragma unroll
		for (int j = 0; j < WE; ++j) {
			const bool ha = !fa_ && j < corridor && half_lo(*chk_at(0, j)) == ma;
			const bool hb = !fb_ && j < corridor && half_hi(*chk_at(1, j)) == mb;
			ra = ha ? j : ra;
			rb = hb ? j : rb;
			fa_ = fa_ || ha;
			fb_ = fb_ || hb;
		}
		// nothing ever exceeded 0: the reference's first cell (0, 0) holds the maximum (oclSwScore.cl:48,88)
		ba.best_read = (ma > 0 && found_a) ? brow_a : 0;
		ba.best_ref = (ma > 0 && found_a) ? ra : 0;
		ba.best_score = ma >> 2;
		bb.best_read = (mb > 0 && found_b) ? brow_b : 0;
		bb.best_ref = (mb > 0 && found_b) ? rb : 0;
		bb.best_score = mb >> 2;
	} else {
		int cma = 4 * kEndFreeMinS16, cmb = 4 * kEndFreeMinS16, ra = 0, rb = 0;
#pragma unroll
		for (int j = 0; j < W; ++j) {                 // first strict maximum of the final row (oclEndFreeScore.cl:135-140)
			const int va_ = (int) (short) (line[j] & 0xFFFFu), vb_ = (int) (short) (line[j] >> 16);
			const bool ga = j < corridor && va_ > cma, gb = j < corridor && vb_ > cmb;
			cma = ga ? va_ : cma;
			ra = ga ? j : ra;
			cmb = gb ? vb_ : cmb;
			rb = gb ? j : rb;
		}
		ba.best_read = rc_a - 1;
		ba.best_ref = ra;
		ba.best_score = cma >> 2;
		bb.best_read = rc_b - 1;
		bb.best_ref = rb;
		bb.best_score = cmb >> 2;
	}
	// quad-skipped pairs: what the reference's forward kernel leaves behind (oclSwScore.cl:16-18,104-106)
	if (!act_a) { ba.best_read = MODE == 0 ? 0 : -1; ba.best_ref = 0; ba.best_score = 0; ba.read_count = 0; }
	if (!act_b) { bb.best_read = MODE == 0 ? 0 : -1; bb.best_ref = 0; bb.best_score = 0; bb.read_count = 0; }
	best_out[ia] = make_int4(ba.best_read, ba.best_ref, ba.best_score, ba.read_count);
	if (vb) best_out[ia + 1] = make_int4(bb.best_read, bb.best_ref, bb.best_score, bb.read_count);
}

}  // namespace ngm
