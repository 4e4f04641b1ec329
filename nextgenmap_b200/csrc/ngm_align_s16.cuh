// ngm_align_s16.cuh -- forward pass with pointers + backtrace + CIGAR/MD for TWO alignments per
// thread on s16x2 lanes (pair A = low half, pair B = high half).
//
// Direction tags.  All scores are kept scaled by 4; the three candidates of a cell carry their
// direction in the two low bits:
//     diag : 4*(H_diag + S)        + 2      (the LUT bytes already hold 4*S + 2)
//     up   : 4*(H_up   + gap_read) + 1
//     left : 4*(H_left + gap_ref)  + 0
// One packed max therefore yields score AND pointer, with the reference's tie order
// diag > I > D (oclSwScore.cl:80-83) for free, because true scores differ by >= 1, i.e. >= 4 scaled:
//     d = VIADD.16x2(line[j], s)   u = VIADDMNMX.S16x2(line[j+1], 4*gr+1, d)
//     h = VIADDMNMX.S16x2[.RELU](left, 4*gf, u)   clean = h & 0xFFFCFFFC   tag = h - clean
// and the 2-bit tags of eight cells are accumulated on the FMA pipe (pw = pw*4 + tag): one 32-bit
// pointer word = eight cells of A (low 16 bits) and of B (high 16 bits).
//
// Best cell (local mode, oclSwScore.cl:88-91: first strict maximum in row-major order): per row a
// VIMNMX3 tree gives the row maximum; when a half's running maximum strictly improves, that half of
// the band is copied into a snapshot (one PRMT per register, selector chosen per row).  After the
// last row the snapshot holds, per half, the row in which the final maximum first appeared; its
// first slot equal to the maximum is the best cell.
//
// Exactness needs 4*score to fit int16 and 4*S+2 to fit int8; build_params() decides
// (DevParams::align_s16), otherwise align_i32_kernel is used.  End-free mode uses the sentinel
// -7900 (scaled -31600) instead of -16000: under the same host-side range check the sentinel can
// never win a maximum, so its exact value is immaterial.
#pragma once

#include "ngm_common.cuh"
#include "ngm_dp_i32.cuh"
#include "ngm_dp_s16.cuh"
#include "ngm_format.cuh"
#include "../../include/ngm_b200.h"

namespace ngm {

constexpr int kEndFreeMinS16 = -7900;
// rows per loop body of the first-generation forward kernel on wide bands (capacity > 48).  Measured on B200: 250 bp / corridor 80 (5 M
// alignments) 8 rows 48.2 ms, 4 rows 39.4, 2 rows 28.3, 1 row 26.4; 400 bp / corridor 65 (capacity 72) 90.0 / 85.9 / 68.5 / 70.5 ms.
// ncu on the eight-row body: `no_instruction` 3.4 stalls per issue, integer pipe 40 % busy -- ~100 KB of SASS per body.
#ifndef NGM_FWD1_WIDE_UNROLL
#define NGM_FWD1_WIDE_UNROLL(W) ((W) >= 80 ? 1 : 2)
#endif

template <int W>
struct TagGeom {
	static constexpr int kWords = (W + 7) / 8;     // pointer words per row per thread (8 cells x 2 pairs each)
};

struct HalfBest {
	int best_read, best_ref, best_score, read_count;
};

// ---------------------------------------------------------------------------
// backtrace + CIGAR/MD: a separate, register-light kernel (one alignment per thread, high
// occupancy).  The walk is a chain of dependent loads (pointer word -> next cell); the first
// version ran it inside the forward kernel at 3 warps/SMSP and spent 2/3 of the kernel's time
// stalled on those loads (profiles/r1_v2_*).  Here the words of the next four rows are
// prefetched into a register ring (the walk moves up one row per step except on deletions), and
// match cells carry their own tag (3 = diag & EQ, 2 = diag & X), so the common step needs no
// sequence lookup at all.
// ---------------------------------------------------------------------------
// Runs of matches in one step (-DNGM_BT_RUNS=0: one cell per loop iteration, the first form): a diagonal run stays on one band slot, so
// the tags of the next eight rows sit at the same bit position of eight ring words; the walk takes all leading EQ cells of that window
// at once and spends a full iteration only on X / I / D cells.
#ifndef NGM_BT_RUNS
#define NGM_BT_RUNS 0
#endif
constexpr int kRing = NGM_BT_RUNS ? 32 : 16;       // rows in flight per thread (power of two)
constexpr int kRunRows = 8;                        // rows examined per iteration (NGM_BT_RUNS)
constexpr int kSpecialFlag = 1 << 30;   // in HalfBest::read_count: sequences hold codes other than A/C/G/T

__device__ __forceinline__ void cp_async4(uint32_t *smem_dst, const uint32_t *gmem_src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t) __cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct BandRt {
	int words;       // pointer words per row per thread slot = ceil(W / 8)
	int cap;         // band capacity W
	int tstride;     // thread slots per launch (distance between column groups of one row)
};

__device__ __forceinline__ uint32_t tagged_code(uint32_t w, int col, int cap, int half) {
	const int g = col >> 3;
	const int cnt = min(8, cap - 8 * g);
	return (w >> (2 * (cnt - 1 - (col & 7)) + 16 * half)) & 3u;
}

template <int MODE>
__device__ __forceinline__ TraceOut backtrace_tagged(const DevParams &P, const uint2 *s_lut, const PairCtx &c, const uint32_t *__restrict__ pbase,
		size_t row_stride, BandRt geo, int half, uint16_t *__restrict__ ops, int ops_stride, int ops_cap, const int4 b, uint32_t *ring,
		int ring_stride) {
	TraceOut o;
	o.qstart = 0;
	o.qend = 0;
	o.sp = 0;
	const int best_read = b.x, best_ref = b.y, best_score = b.z, read_count = b.w & (kSpecialFlag - 1);
	const bool plain = !P.alt && !(b.w & kSpecialFlag);            // every diag step scores `match` or `mismatch`
	if (best_read <= 0) {                                         // oclSwCigar.cl:78
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
	// Shared-memory ring fed by cp.async (LDGSTS): slot (r & 15) of this thread's column holds the pointer
	// word of row r at column group g; the copies for the next 15 rows are in flight while the current one
	// is consumed, so the dependent-load latency of the walk (the scratch matrix of a launch is larger
	// than L2) is overlapped instead of paid per step.  One loop iteration = one step for every lane; the
	// common step (tag 3: diag & EQ) touches registers + one LDS, the rare ones (X, I, D, band border)
	// take a short divergent branch.
	const long long rs = (long long) row_stride;
	int g = col >> 3;
	int sh = 2 * (min(8, geo.cap - 8 * g) - 1 - (col & 7)) + 16 * half;
	const uint32_t *pg = pbase + (size_t) (row + c.sub) * row_stride + (size_t) g * geo.tstride;   // word of the current row
	auto fill = [&](int r, const uint32_t *src) {                                 // async copy of row r's word (or an empty group)
		if (r >= 0) cp_async4(ring + (r & (kRing - 1)) * ring_stride, src);
		cp_async_commit();
	};
#pragma unroll
	for (int k = 0; k < kRing; ++k) fill(row - k, pg - k * rs);
	const uint32_t *pf = pg - (long long) kRing * rs;                             // next row to fetch: row - kRing
	cp_async_wait<kRing - 1>();
	uint32_t w0 = ring[(row & (kRing - 1)) * ring_stride];
	const bool fast_eq = !P.alt;                                  // tag 3 == EQ with score `match`
	bool border = col < 0 || col >= corridor;
	while (row >= 0) {
		int op;
		bool up_row = true;
#if NGM_BT_RUNS
		// every row consumed was matched by one (possibly empty) copy group, so kRing groups are outstanding here and all but the
		// kRing - kRunRows newest -- the rows row .. row - 7 -- have landed after this wait
		cp_async_wait<kRing - kRunRows>();
		w0 = ring[(row & (kRing - 1)) * ring_stride];
		if (!border && fast_eq) {
			uint32_t code = (w0 >> sh) & 3u;
#pragma unroll
			for (int k = 1; k < kRunRows; ++k) {
				const uint32_t wk = ring[((row - k) & (kRing - 1)) * ring_stride];       // (stale for rows < 0: clamped below)
				code |= ((wk >> sh) & 3u) << (2 * k);
			}
			const uint32_t eq = code & (code >> 1) & 0x5555u;                           // bit 2k: row - k carries tag 3
			int lead = (__ffs((int) ((~eq & 0x5555u) | 0x10000u)) - 1) >> 1;            // leading EQ cells of the window
			lead = min(lead, row + 1);
			// the same as `lead` iterations of the single-cell fast path: local mode stops at h <= 0, checked before every cell
			if (lead > 1 && (MODE != 0 || h - (lead - 1) * P.match > 0)) {
				h -= lead * P.match;
				abs_ref -= lead;
				if (elem == OP_EQ) {
					len += lead;
				} else {
					if (sp < ops_cap) ops[(size_t) sp * ops_stride] = (uint16_t) (len << 4 | elem);
					sp += 1;
					elem = OP_EQ;
					len = lead;
				}
#pragma unroll
				for (int k = 0; k < kRunRows; ++k) {
					if (k < lead) {
						fill(row - kRing - k, pf);                        // the window slides: row - 1 - k enters with row - 1 - k - (kRing - 1)
						pf -= rs;
					}
				}
				row -= lead;
				continue;
			}
		}
#endif
		const uint32_t p = (w0 >> sh) & 3u;
		if (!border && (MODE != 0 || h > 0) && p == 3u && fast_eq) {
			op = OP_EQ;
			h -= P.match;
			row -= 1;
			abs_ref -= 1;
		} else if (border) {
			if (MODE == 0) break;
			op = OP_X;
			row -= 1;
			abs_ref -= 1;
		} else {
			if (MODE == 0 && h <= 0) break;
			if (p >= 2u) {
				op = p == 3u ? OP_EQ : OP_X;
				if (plain) {
					h -= P.mismatch;                              // tag 2 between plain codes (tag 3 took the fast path)
				} else {
					const int rc = code_at(c.rp, row) & 7;
					const int fc = code_at(c.wp, (int64_t) c.sub + row + col) & 7;
					h -= lut_score(s_lut, c.dir, rc, fc);
				}
				row -= 1;
				abs_ref -= 1;
			} else {
				if (p == 1u) {
					op = OP_I;
					h -= P.gap_read;
					row -= 1;
					col += 1;
				} else {
					op = OP_D;
					h -= P.gap_ref;
					col -= 1;
					abs_ref -= 1;
					up_row = false;
				}
				border = col < 0 || col >= corridor;
				const int ng = (col < 0 ? 0 : col) >> 3;
				sh = 2 * (min(8, geo.cap - 8 * ng) - 1 - (col & 7)) + 16 * half;
				if (ng != g && !border && row >= 0) {             // column group changed: restart the ring there
					g = ng;
					cp_async_wait<0>();
					pg = pbase + (size_t) (row + c.sub) * row_stride + (size_t) g * geo.tstride;
#pragma unroll
					for (int k = 0; k < kRing; ++k) fill(row - k, pg - k * rs);
					pf = pg - (long long) kRing * rs;
					cp_async_wait<kRing - 1>();
					w0 = ring[(row & (kRing - 1)) * ring_stride];
					up_row = false;                               // ring already positioned on `row`
				}
			}
		}
		if (op == elem) {
			len += 1;
		} else {
			if (sp < ops_cap) ops[(size_t) sp * ops_stride] = (uint16_t) (len << 4 | elem);
			sp += 1;
			elem = op;
			len = 1;
		}
		if (up_row && row >= 0) {
			fill(row - (kRing - 1), pf);                          // reuses the slot of the row just left
			pf -= rs;
#if !NGM_BT_RUNS
			cp_async_wait<kRing - 1>();
			w0 = ring[(row & (kRing - 1)) * ring_stride];
#endif
		}
	}
	cp_async_wait<0>();
	if (sp < ops_cap) ops[(size_t) sp * ops_stride] = (uint16_t) (len << 4 | elem);
	sp += 1;
	o.sp = sp;
	o.pos = abs_ref + 1;
	o.qstart = row + 1;
	o.qend = qend;
	return o;
}

__device__ __forceinline__ void fill_record(ngm_b200_align_rec &r, const TraceOut &t, const FormatOut &f, uint32_t off) {
	r.position_offset = t.pos;
	r.qstart = t.qstart;
	r.qend = t.qend > 0 ? t.qend : 0;
	r.str_off = off;
	if (t.ok) {
		r.nm = f.mismatch;
		r.identity = (float) f.match * 1.0f / (float) f.total;       // SWOclCigar.cpp:609
		r.score = (float) f.read_index;                              // SWOclCigar.cpp:613
		r.cigar_len = (uint16_t) f.cigar_len;
		r.md_len = (uint16_t) f.md_len;
	} else {
		r.nm = 0;                                                    // failure convention, DESIGN.md section 5
		r.identity = 0.0f;
		r.score = -1.0f;
		r.cigar_len = 0;
		r.md_len = 0;
	}
}

__device__ __forceinline__ void store_record(ngm_b200_align_rec *recs, int idx, const ngm_b200_align_rec &r) {
	reinterpret_cast<uint4 *>(recs)[2 * (size_t) idx] = *reinterpret_cast<const uint4 *>(&r);
	reinterpret_cast<uint4 *>(recs)[2 * (size_t) idx + 1] = *(reinterpret_cast<const uint4 *>(&r) + 1);
}

// KNOWN = the local maximum of every pair is already known (`known`, the score kernel's output): the best
// cell is then found by comparing each row's maximum with it and scanning the band once, in the first row that
// reaches it; no band snapshot is kept, which halves the register footprint and lets wide bands (capacity
// 56..96: `-C 40`, 400 bp reads) stay on the s16x2 path.
template <int W, int LO, int MODE, bool KNOWN>
__global__ void __launch_bounds__(128) align_s16_fwd_kernel(const __grid_constant__ DevParams P, const PairDesc *__restrict__ pairs, int n,
		const uint32_t *__restrict__ reads_fwd, const uint32_t *__restrict__ reads_rev, const uint16_t *__restrict__ rlen,
		const uint32_t *__restrict__ ref4, uint32_t *__restrict__ ptr_scratch, int stride, int4 *__restrict__ best_out,
		const float *__restrict__ known) {
	using G = BandGeom<W>;
	using T = TagGeom<W>;
	__shared__ uint2 s_lut4[16];
	if (threadIdx.x < 16) s_lut4[threadIdx.x] = P.lut4[threadIdx.x];
	__syncthreads();
	const int t2 = blockIdx.x * blockDim.x + threadIdx.x;        // thread slot: pairs 2*t2, 2*t2 + 1
	const int ia = 2 * t2;
	if (ia >= n) return;
	const bool va = true, vb = ia + 1 < n;
	const int ib = vb ? ia + 1 : (va ? ia : 0);
	const int ja = va ? ia : 0;
	constexpr int SENT = MODE == 0 ? 0 : 4 * kEndFreeMinS16;
	const uint32_t SENT2 = pack2(SENT, SENT);
	PairCtx ca, cb;
	uint32_t fa, fb;
	bool act_a = load_pair(P, pairs, ja, reads_fwd, reads_rev, rlen, ref4, ca, fa) && va;
	bool act_b = load_pair(P, pairs, ib, reads_fwd, reads_rev, rlen, ref4, cb, fb) && vb;
	const int corridor = P.corridor;
	const uint32_t gr2 = pack2(4 * P.gap_read + 1, 4 * P.gap_read + 1), gf2 = pack2(4 * P.gap_ref, 4 * P.gap_ref);
	const uint32_t c_four = P.c_four, c_neg1 = P.c_neg1;
	const int tstride = stride >> 1;                             // thread slots per launch
	// pointer matrix layout [row q][column group][thread slot]: the words of one group and row are contiguous
	// over thread slots, so the forward stores and the backtrace loads of a warp are both fully coalesced
	const size_t row_stride = (size_t) tstride * T::kWords;
	uint32_t *pbase = ptr_scratch + (size_t) t2;
	HalfBest ba, bb;
	{
		uint32_t line[W + 1], snap[(MODE == 0 && !KNOWN) ? W : 1];
#pragma unroll
		for (int j = 0; j <= W; ++j) line[j] = (j < corridor) ? 0u : SENT2;
		if (MODE == 0 && !KNOWN) {
#pragma unroll
			for (int j = 0; j < W; ++j) snap[j] = 0u;
		}
		uint32_t best = 0;
		int brow_a = 0, brow_b = 0, rc_a = 0, rc_b = 0;
		int tgt_a = 0, tgt_b = 0, col_a = 0, col_b = 0;
		bool found_a = false, found_b = false;
		if (MODE == 0 && KNOWN) {
			tgt_a = 4 * (int) known[ja];
			tgt_b = 4 * (int) known[ib];
		}
		uint32_t wa[G::kWin], wb[G::kWin];
		// "special" = a code other than A/C/G/T inside the read or inside the part of the window the band can
		// touch: only then does a non-EQ diagonal step score something else than `mismatch`, and only then
		// does the backtrace have to look at the sequences (bit 2 of a nibble <=> code >= 4)
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
		uint32_t *prow = pbase;
		for (int qw = 0; qw < nqw; ++qw) {
			const uint32_t cur_a = __ldg(ca.rp + qw), cur_b = __ldg(cb.rp + qw);
			const uint32_t rda = __funnelshift_l(prev_a, cur_a, 4 * ca.sub);
			const uint32_t rdb = __funnelshift_l(prev_b, cur_b, 4 * cb.sub);
			prev_a = cur_a;
			prev_b = cur_b;
			const uint32_t next_a = __ldg(ca.wp + qw + G::kWin), next_b = __ldg(cb.wp + qw + G::kWin);
			spec_a |= (next_a & nib_mask(wend_a - 8 * (qw + G::kWin))) | (cur_a & nib_mask(ca.len - 8 * qw));
			spec_b |= (next_b & nib_mask(wend_b - 8 * (qw + G::kWin))) | (cur_b & nib_mask(cb.len - 8 * qw));
			// wide bands: the fully unrolled eight-row body is ~100 KB of SASS at W = 80 (instruction fetch)
			constexpr int kUnroll = W > kAlignS16MaxLocal ? NGM_FWD1_WIDE_UNROLL(W) : 8;
#pragma unroll kUnroll
			for (int t = 0; t < 8; ++t) {
				const int rca = (rda >> (4 * t)) & 7, rcb = (rdb >> (4 * t)) & 7;
				const uint2 ta = luta[rca];
				const uint2 tb = lutb[rcb];
				uint32_t ala[G::kAligned], alb[G::kAligned];
#pragma unroll
				for (int k = 0; k < G::kAligned; ++k) {
					ala[k] = (kUnroll == 8 && t == 0) ? wa[k] : __funnelshift_r(wa[k], wa[k + 1], 4 * t);
					alb[k] = (kUnroll == 8 && t == 0) ? wb[k] : __funnelshift_r(wb[k], wb[k + 1], 4 * t);
				}
				uint32_t left = SENT2;
				uint32_t pw[T::kWords];
#pragma unroll
				for (int k = 0; k < T::kWords; ++k) pw[k] = 0;
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
						const uint32_t clean = h & 0xFFFCFFFCu;
						pw[j >> 3] = imad_u32(pw[j >> 3], c_four, imad_u32(clean, c_neg1, h));      // 4 * pw + (h - clean), both on the FMA pipe
						left = clean;
						line[j] = clean;
					}
				}
#pragma unroll
				for (int k = 0; k < T::kWords; ++k) prow[(size_t) k * tstride] = pw[k];
				prow += row_stride;
				if (MODE == 0 && KNOWN) {
					uint32_t mx = line[0];
#pragma unroll
					for (int j = 1; j + 1 < W; j += 2) mx = __vimax3_s16x2(mx, line[j], line[j + 1]);
					if ((W & 1) == 0) mx = __vmaxs2(mx, line[W - 1]);
					const bool hit_a = !found_a && tgt_a > 0 && (int) (short) (mx & 0xFFFFu) == tgt_a;
					const bool hit_b = !found_b && tgt_b > 0 && (int) (short) (mx >> 16) == tgt_b;
					if (hit_a || hit_b) {                             // once per alignment: first slot holding the maximum
						bool fa_ = false, fb_ = false;
#pragma unroll
						for (int j = 0; j < W; ++j) {
							const bool ha = hit_a && !fa_ && j < corridor && (int) (short) (line[j] & 0xFFFFu) == tgt_a;
							const bool hb = hit_b && !fb_ && j < corridor && (int) (short) (line[j] >> 16) == tgt_b;
							col_a = ha ? j : col_a;
							col_b = hb ? j : col_b;
							fa_ = fa_ || ha;
							fb_ = fb_ || hb;
						}
						if (hit_a) { found_a = true; brow_a = rc_a; }
						if (hit_b) { found_b = true; brow_b = rc_b; }
					}
				}
				if (MODE == 0 && !KNOWN) {
					uint32_t mx = line[0];
#pragma unroll
					for (int j = 1; j + 1 < W; j += 2) mx = __vimax3_s16x2(mx, line[j], line[j + 1]);
					if ((W & 1) == 0) mx = __vmaxs2(mx, line[W - 1]);
					const uint32_t nb = __vmaxs2(best, mx);
					const uint32_t imp = nb ^ best;
					best = nb;
					const bool imp_a = (imp & 0xFFFFu) != 0, imp_b = (imp >> 16) != 0;
					const uint32_t sel = (imp_a ? 0x0054u : 0x0010u) | (imp_b ? 0x7600u : 0x3200u);
#pragma unroll
					for (int j = 0; j < W; ++j) snap[j] = prmt(snap[j], line[j], sel);
					brow_a = imp_a ? rc_a : brow_a;
					brow_b = imp_b ? rc_b : brow_b;
				}
				rc_a += (rca != kCodeNul);
				rc_b += (rcb != kCodeNul);
			}
#pragma unroll
			for (int k = 0; k + 1 < G::kWin; ++k) {
				wa[k] = wa[k + 1];
				wb[k] = wb[k + 1];
			}
			wa[G::kWin - 1] = next_a;
			wb[G::kWin - 1] = next_b;
		}
		ba.read_count = rc_a | ((spec_a & 0x44444444u) ? kSpecialFlag : 0);
		bb.read_count = rc_b | ((spec_b & 0x44444444u) ? kSpecialFlag : 0);
		if (MODE == 0 && KNOWN) {
			ba.best_read = found_a ? brow_a : 0;
			ba.best_ref = found_a ? col_a : 0;
			ba.best_score = found_a ? (tgt_a >> 2) : 0;
			bb.best_read = found_b ? brow_b : 0;
			bb.best_ref = found_b ? col_b : 0;
			bb.best_score = found_b ? (tgt_b >> 2) : 0;
		} else if (MODE == 0) {
			const int ma = (int) (short) (best & 0xFFFFu), mb = (int) (short) (best >> 16);
			int ra = 0, rb = 0;
			bool fa_ = false, fb_ = false;
#pragma unroll
			for (int j = 0; j < W; ++j) {
				const bool ha = !fa_ && j < corridor && (int) (short) (snap[j] & 0xFFFFu) == ma;
				const bool hb = !fb_ && j < corridor && (int) (short) (snap[j] >> 16) == mb;
				ra = ha ? j : ra;
				rb = hb ? j : rb;
				fa_ = fa_ || ha;
				fb_ = fb_ || hb;
			}
			// nothing ever exceeded 0: the reference's first cell (0, 0) holds the maximum (oclSwScore.cl:48,88)
			ba.best_read = ma > 0 ? brow_a : 0;
			ba.best_ref = ma > 0 ? ra : 0;
			ba.best_score = ma >> 2;
			bb.best_read = mb > 0 ? brow_b : 0;
			bb.best_ref = mb > 0 ? rb : 0;
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
	}
	// quad-skipped pairs: what the reference's forward kernel leaves behind (oclSwScore.cl:16-18,104-106)
	if (!act_a) { ba.best_read = MODE == 0 ? 0 : -1; ba.best_ref = 0; ba.best_score = 0; ba.read_count = 0; }
	if (!act_b) { bb.best_read = MODE == 0 ? 0 : -1; bb.best_ref = 0; bb.best_score = 0; bb.read_count = 0; }
	best_out[ia] = make_int4(ba.best_read, ba.best_ref, ba.best_score, ba.read_count);
	if (vb) best_out[ia + 1] = make_int4(bb.best_read, bb.best_ref, bb.best_score, bb.read_count);
}

// One alignment per thread: pointer walk (oclSW_Backtracking, oclSwCigar.cl:60-124), RLE op stack,
// CIGAR / MD / NM / identity (computeCigarMD, SWOclCigar.cpp:430-615), compact string heap.
template <int MODE>
__global__ void __launch_bounds__(256) backtrace_format_kernel(const __grid_constant__ DevParams P, const PairDesc *__restrict__ pairs, int n,
		const uint32_t *__restrict__ reads_fwd, const uint32_t *__restrict__ reads_rev, const uint16_t *__restrict__ rlen,
		const uint32_t *__restrict__ ref4, const uint32_t *__restrict__ ptr_scratch, int capacity, const int4 *__restrict__ best_in,
		uint16_t *__restrict__ ops_scratch, int stride, int ops_cap, ngm_b200_align_rec *__restrict__ recs, char *__restrict__ strings,
		uint32_t str_cap, uint32_t *__restrict__ cursor, float *__restrict__ out_best, const int *__restrict__ slot_of,
		const int *__restrict__ range, int ops_stride, const int *__restrict__ items_dev, const int *__restrict__ rec_of) {
	__shared__ uint2 s_lut[16];
	__shared__ uint32_t s_ring[kRing][256];
	if (threadIdx.x < 16) s_lut[threadIdx.x] = P.lut[threadIdx.x];
	__syncthreads();
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	// slot_of != nullptr: thread idx writes the record of item idx (a read of the chunk) from the forward pass's slot slot_of[idx]
	// (-1: nothing to align); the pair list starts at pairs[range[0]].  Otherwise item idx IS slot idx.
	// items_dev: the number of items is only known on the device; rec_of: item idx writes recs[rec_of[idx]].
	if (items_dev != nullptr) n = min(n, *items_dev);
	const bool in_range = idx < n;
	const int slot = in_range ? (slot_of != nullptr ? slot_of[idx] : idx) : -1;
	const bool valid = slot >= 0;
	const int id = valid ? slot : 0;
	if (range != nullptr) pairs += range[0];
	PairCtx c;
	uint32_t flags;
	const bool active = load_pair(P, pairs, id, reads_fwd, reads_rev, rlen, ref4, c, flags);
	// the forward pass's maximum is the pair's BatchScore result in this mode (oclSW / oclSW_Global compute the same recurrence)
	if (out_best != nullptr && in_range) out_best[idx] = active ? (float) best_in[id].z : (MODE == 0 ? -1.0f : (float) kEndFreeMin);
	BandRt geo;
	geo.cap = capacity;
	geo.words = (capacity + 7) / 8;
	geo.tstride = stride >> 1;
	const size_t row_stride = (size_t) geo.tstride * geo.words;
	const uint32_t *pbase = ptr_scratch + (size_t) (id >> 1);
	uint16_t *ops = ops_scratch + (in_range ? idx : 0);            // the op stack belongs to the item, not to the slot
	TraceOut t;
	t.ok = 0;
	t.pos = 0;
	t.qstart = t.qend = t.sp = 0;
	if (valid) t = backtrace_tagged<MODE>(P, s_lut, c, pbase, row_stride, geo, id & 1, ops, ops_stride, ops_cap, best_in[id], &s_ring[0][threadIdx.x], 256);
	FormatOut f;
	f.cigar_len = f.md_len = f.match = f.mismatch = f.total = f.read_index = 0;
	if (t.ok) f = format_cigar_md<false>(P, c, ops, ops_stride, t, nullptr, nullptr);
	const uint32_t need = t.ok ? (uint32_t) (f.cigar_len + f.md_len) : 0u;
	const int lane = threadIdx.x & 31;
	uint32_t incl = need;
#pragma unroll
	for (int dlt = 1; dlt < 32; dlt <<= 1) {
		const uint32_t v = __shfl_up_sync(0xffffffffu, incl, dlt);
		if (lane >= dlt) incl += v;
	}
	uint32_t base = 0;
	if (lane == 31 && incl) base = atomicAdd(cursor, incl);      // one atomic per warp
	base = __shfl_sync(0xffffffffu, base, 31);
	const uint32_t off = base + incl - need;
	if (!in_range) return;
	if (t.ok && (uint64_t) off + need <= (uint64_t) str_cap) format_cigar_md<true>(P, c, ops, ops_stride, t, strings + off, strings + off + f.cigar_len);
	ngm_b200_align_rec r;
	fill_record(r, t, f, off);
	store_record(recs, rec_of != nullptr ? rec_of[idx] : idx, r);
}

}  // namespace ngm
