// ngm_batch.cu -- one scored + selected + aligned read batch per call, pipelined over several lanes inside the library.
//
// What the host side of NGM does between candidate search and the writer (src/ScoreBuffer.cpp:80-277,365-502: window fetch,
// BatchScore, top1SE / top1PE, computeMQ; src/AlignmentBuffer.cpp:64-147: BatchAlign, computeCigarMD) for a whole batch, behind ONE
// C-ABI call that takes host buffers: ngm_b200_run_batch.  This is the call a re-plumbed ScoreBuffer / AlignmentBuffer submits its
// pinned staging buffers to (north_star), and the call bench.py's `e2e` times.
//
//   lanes       a lane is a child context: its own stream, read / pair / result staging and alignment scratch; the packed reference,
//               the k-mer index and the paired-end running sums are borrowed from the root context.  Sub-batches rotate over the
//               lanes, so the host->device copy of sub-batch i+1, the kernels of sub-batch i and the device->host copy of sub-batch
//               i-1 overlap (the reference overlaps its host-side packing with the kernels the same way, SWOcl.cpp:435-444).
//   fusion      a read with ONE candidate does not need BatchScore before its alignment: the forward pass of the alignment computes
//               the same recurrence and its maximum IS the pair's score (oclSW vs oclSW_Score, oclSwScore.cl:4-154).  Such pairs skip
//               the score kernel; their Score.f / MAPQ are filled in after the alignment.  Results are identical.
//   formats     reads as NUL-padded ASCII rows or 2-bit packed + exception list (42 instead of 152 bytes per 150 bp read over PCIe);
//               descriptors as ngm_b200_pair (16 bytes) or one 64-bit word (the read index is implicit in cand_begin).
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <cub/cub.cuh>

#include "ngm_ctx.h"
#include "ngm_launch.h"

using namespace ngm;

namespace ngm {

// ---------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread2to4(uint32_t x16) {      // 8 two-bit fields -> 8 nibbles (A0 C1 G2 T3 are the device codes already)
	uint32_t y = (x16 | (x16 << 8)) & 0x00FF00FFu;
	y = (y | (y << 4)) & 0x0F0F0F0Fu;
	y = (y | (y << 2)) & 0x33333333u;
	return y;
}

__device__ __forceinline__ uint32_t nib_keep(int valid) { return valid >= 8 ? 0xFFFFFFFFu : (valid <= 0 ? 0u : ((1u << (4 * valid)) - 1u)); }

// PACKED2 rows -> forward words, reverse-complement words (MappedRead::computeReverseSeq, MappedRead.cpp:36-67) and lengths in one
// pass.  One thread per output word index (row, w): it writes forward word w and reverse word w, so a warp's stores are two runs of 32
// consecutive words and the 40-byte input row stays in L1 for the ~21 threads that read it.  HBM-bound by design: 42 B in, 2 x 84 B out
// per 150 bp read.  (The first form -- 16 threads per row, two words each, bounds-checked 64-bit windows -- issued 116 warp instructions
// per row and was issue-bound at 1.29 ms per 10 M reads, profiles/r2_hot_path.md.)
constexpr int kExpandThreads = 256;

__global__ void __launch_bounds__(kExpandThreads) expand_packed2_kernel(const uint32_t *__restrict__ in, int rows, int in_words,
		const uint16_t *__restrict__ len_in, int max_len, uint32_t *__restrict__ fwd, uint32_t *__restrict__ rev, uint16_t *__restrict__ rlen, int words,
		uint32_t words_magic) {
	const uint32_t idx = blockIdx.x * kExpandThreads + threadIdx.x;
	const uint32_t row = __umulhi(idx, words_magic);                // idx / words (words_magic = ceil(2^32 / words); the launcher keeps idx in the exact range)
	const int w = (int) (idx - row * (uint32_t) words);
	if (row >= (uint32_t) rows) return;
	const int len = min((int) len_in[row], max_len);
	const uint32_t *r = in + (size_t) row * in_words;
	// forward word w: bases 8w .. 8w + 7 = half (w & 1) of input word w >> 1
	const int qf = w >> 1;
	const uint32_t xin = qf < in_words ? __ldg(r + qf) : 0u;
	const uint32_t keep = nib_keep(len - 8 * w);
	fwd[(size_t) idx] = (spread2to4((xin >> (16 * (w & 1))) & 0xFFFFu) & keep) | (kNulWord & ~keep);
	// reverse word w: nibble k = complement of base hi - k
	uint32_t out = kNulWord;
	const int hi = len - 1 - 8 * w;
	if (hi >= 0) {
		const int lo = hi - 7;                                    // may be negative: those nibbles are masked below
		const int q = lo >> 4;                                    // floor division (arithmetic shift)
		const uint32_t w0 = (q >= 0) ? __ldg(r + q) : 0u;
		const uint32_t w1 = (q + 1 < in_words) ? __ldg(r + q + 1) : 0u;
		const uint32_t v = __funnelshift_r(w0, w1, 2 * (lo & 15)) & 0xFFFFu;      // bases lo .. hi
		uint32_t y = __byte_perm(spread2to4(v), 0, 0x0123);
		y = ((y >> 4) & 0x0F0F0F0Fu) | ((y & 0x0F0F0F0Fu) << 4);                   // nibble k = base hi - k
		const uint32_t kp = nib_keep(hi + 1);
		out = ((y ^ 0x33333333u) & kp) | (kNulWord & ~kp);                         // A<->T, C<->G: code ^ 3
	}
	rev[(size_t) idx] = out;
	if (w == 0) rlen[row] = (uint16_t) len;
}

// bases that are not A/C/G/T: oclDefines.cl:64-80 classes (N 5, everything else 4; a NUL inside the row ends nothing here: lengths are explicit)
__global__ void patch_exceptions_kernel(const ngm_b200_read_exc *__restrict__ exc, uint32_t n_exc, uint32_t read_base, int rows, const uint16_t *__restrict__ rlen,
		uint32_t *__restrict__ fwd, uint32_t *__restrict__ rev, int words) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_exc) return;
	const ngm_b200_read_exc e = exc[i];
	const long long row = (long long) e.read_index - (long long) read_base;
	if (row < 0 || row >= rows) return;
	const int len = rlen[row];
	if ((int) e.pos >= len) return;
	const uint32_t u = e.ch & 0xDFu;
	const uint32_t code = u == 'A' ? 0u : u == 'C' ? 1u : u == 'G' ? 2u : u == 'T' ? 3u : u == 'N' ? 5u : (e.ch == 0 ? 6u : 4u);
	uint32_t *f = fwd + (size_t) row * words + (e.pos >> 3);
	const int sf = 4 * (e.pos & 7);
	atomicAnd(f, ~(0xFu << sf));
	atomicOr(f, code << sf);
	const int rp = len - 1 - (int) e.pos;
	uint32_t *rv = rev + (size_t) row * words + (rp >> 3);
	const int sr = 4 * (rp & 7);
	const uint32_t rcode = code < 4 ? 3u - code : code;
	atomicAnd(rv, ~(0xFu << sr));
	atomicOr(rv, rcode << sr);
}

// Per read: resolve its candidates' descriptors (window start -> nibble index of the resident reference; starts >= concat_len, incl. the
// unsigned underflow of loc - corridor/2, select the all-'N' region, ScoreBuffer.cpp:113-118) and, when single-candidate reads skip the
// score kernel, append the candidates of every other read to the score kernel's work list (one atomicAdd per warp).
template <int FMT>
__global__ void __launch_bounds__(256) batch_plan_kernel(int n_reads, const int *__restrict__ cb, int cb_base, const void *__restrict__ desc,
		PairDesc *__restrict__ rp, ngm_b200_pair *__restrict__ pairs16, unsigned long long concat_len, unsigned long long n_region_nib,
		const uint16_t *__restrict__ rlen, int fuse, int *__restrict__ sel, int *__restrict__ n_sel) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	const bool valid = r < n_reads;
	int b = 0, e = 0;
	if (valid) {
		b = cb[r] - cb_base;
		e = cb[r + 1] - cb_base;
	}
	const int cnt = e - b;
	const bool empty = valid && rlen[r] == 0;
	for (int i = b; i < e; ++i) {
		unsigned long long ws;
		uint32_t fl;
		if (FMT == NGM_B200_DESC_PAIR16) {
			const ngm_b200_pair p = static_cast<const ngm_b200_pair *>(desc)[i];
			ws = p.window_start;
			fl = p.flags;
		} else {
			const unsigned long long d = static_cast<const unsigned long long *>(desc)[i];
			ws = d & 0x00FFFFFFFFFFFFFFull;
			fl = (uint32_t) (d >> 56);
		}
		PairDesc d;
		d.win_nib = ws < concat_len ? ws : n_region_nib;
		d.read_idx = (uint32_t) r;
		d.flags = (fl & (PF_REVERSE | PF_DIR | PF_INACTIVE)) | (empty ? PF_INACTIVE : 0u);      // an empty read is its own quad leader
		rp[i] = d;
		if (pairs16 != nullptr) {
			ngm_b200_pair p;
			p.window_start = ws;
			p.read_index = (uint32_t) r;
			p.flags = fl;
			pairs16[i] = p;
		}
	}
	if (!fuse) return;
	const int want = cnt > 1 ? cnt : 0;
	const int lane = threadIdx.x & 31;
	int incl = want;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const int v = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) incl += v;
	}
	int base = 0;
	if (lane == 31 && incl) base = atomicAdd(n_sel, incl);
	base = __shfl_sync(0xffffffffu, base, 31);
	int at = base + incl - want;
	for (int i = 0; i < want; ++i) sel[at + i] = b + i;
}

// ScoreBuffer::top1SE + computeMQ (ScoreBuffer.cpp:34-40,228-277) incl. `strata`; single-candidate reads of a fused batch only get their
// winner here (score-dependent fields follow in batch_finalize_kernel)
__global__ void __launch_bounds__(256) batch_select_top1_kernel(int n_reads, const int *__restrict__ cb, int cb_base, const float *__restrict__ scores,
		int fuse, int strata, int *__restrict__ best_pair, int *__restrict__ mapq, int *__restrict__ num_top) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int b = cb[r] - cb_base, e = cb[r + 1] - cb_base;
	if (fuse && e - b == 1) {
		best_pair[r] = b;
		mapq[r] = 0;
		if (num_top != nullptr) num_top[r] = 0;
		return;
	}
	float best = 0.0f, second = 0.0f;
	int besti = 0, nbest = 0;
	for (int j = b; j < e; ++j) {
		const float s = scores[j];
		if (s > second) {
			if (s > best) {
				second = best;
				best = s;
				besti = j - b;
				nbest = 1;
			} else if (s == best) {
				++nbest;
				second = best;
			} else {
				second = s;
			}
		} else if (s == best) {
			++nbest;
		}
	}
	int mq = 0;
	if (best > 0.0f && second >= 0.0f) mq = (int) ceilf(60.0f * (best - second) / best);
	int bp = e > b ? b + besti : -1;
	if (strata && nbest != 1 && e > b) {                           // too many equally scoring positions (ScoreBuffer.cpp:270-276)
		bp = -1;
		mq = 0;
		nbest = 1;
	}
	best_pair[r] = bp;
	mapq[r] = mq;
	if (num_top != nullptr) num_top[r] = nbest;
}

__global__ void __launch_bounds__(256) batch_gather_kernel(int n_reads, const int *__restrict__ cb, int cb_base, const PairDesc *__restrict__ rp,
		const int *__restrict__ best_pair, const float *__restrict__ scores, int fuse, PairDesc *__restrict__ wp, float *__restrict__ wscores) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int bp = best_pair[r];
	PairDesc d;
	float s = 0.0f;
	if (bp >= 0) {
		d = rp[bp];
		const bool single = fuse && (cb[r + 1] - cb[r] == 1);
		if (!single) s = scores[bp];
	} else {
		d.win_nib = 0;
		d.read_idx = (uint32_t) r;
		d.flags = PF_INACTIVE;
	}
	wp[r] = d;
	wscores[r] = s;
}

// topn > 1 (ScoreBuffer::topNSE): row r * topn + j of the align batch = the j-th selected candidate of read r (inactive beyond n_sel[r])
__global__ void __launch_bounds__(256) batch_gather_topn_kernel(int n_reads, int topn, const PairDesc *__restrict__ rp, const int *__restrict__ sel,
		const float *__restrict__ scores, PairDesc *__restrict__ wp, float *__restrict__ wscores, int *__restrict__ best_pair) {
	const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (long long) n_reads * topn) return;
	const int r = (int) (i / topn), j = (int) (i - (long long) r * topn);
	const int s = sel[i];
	PairDesc d;
	float sc = 0.0f;
	if (s >= 0) {
		d = rp[s];
		sc = scores[s];
	} else {
		d.win_nib = 0;
		d.read_idx = (uint32_t) r;
		d.flags = PF_INACTIVE;
	}
	wp[i] = d;
	wscores[i] = sc;
	if (j == 0) best_pair[r] = s;
}

// fused batches: the score of a single-candidate read is the maximum its alignment's forward pass found; then top1SE over that one score
__global__ void __launch_bounds__(256) batch_finalize_kernel(int n_reads, const int *__restrict__ cb, int cb_base, const float *__restrict__ out_best,
		int strata, float *__restrict__ scores, int *__restrict__ best_pair, int *__restrict__ mapq, int *__restrict__ num_top) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int b = cb[r] - cb_base, e = cb[r + 1] - cb_base;
	if (e - b != 1) return;
	const float s = out_best[r];
	scores[b] = s;
	int nbest = s >= 0.0f ? 1 : 0;                                   // s > 0: new best; s == 0: equals the initial best (ScoreBuffer.cpp:238-251)
	int mq = s > 0.0f ? 60 : 0;                                      // computeMQ(best, 0)
	if (strata && nbest != 1) {
		best_pair[r] = -1;
		mq = 0;
		nbest = 1;
	}
	mapq[r] = mq;
	if (num_top != nullptr) num_top[r] = nbest;
}

// ---- forward pass over EVERY candidate of reads with few candidates ("fwd-all") ---------------------------------------------------
// Scoring a pair and then aligning the winner runs the DP twice for the winner.  The forward-with-pointers kernel costs ~1.3x the score
// kernel per pair, so for a read with up to kFwdAllMax candidates it is cheaper to run the forward pass on all of them, pick the winner
// from the forward maxima (the same numbers BatchScore returns) and backtrace only the winner: 1 candidate 104 instead of 104 (no score
// pass either way), 2 candidates 208 instead of 264, 4 candidates 416 instead of 424 ALU instructions per row.  Reads with more candidates
// keep score -> top-1 -> forward of the winner.
constexpr int kFwdAllMax = 4;

// per read: slots in the forward list (all candidates of a small read, one for the winner of a big one), and the score kernel's work list
template <int FMT>
__global__ void __launch_bounds__(256) batch_plan_all_kernel(int n_reads, const int *__restrict__ cb, const void *__restrict__ desc, PairDesc *__restrict__ rp,
		unsigned long long concat_len, unsigned long long n_region_nib, const uint16_t *__restrict__ rlen, int *__restrict__ fcnt, int *__restrict__ sel,
		int *__restrict__ n_sel, ngm_b200_pair *__restrict__ pairs16) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	const bool valid = r < n_reads;
	int b = 0, e = 0;
	if (valid) {
		b = cb[r];
		e = cb[r + 1];
	}
	const int cnt = e - b;
	const bool empty = valid && rlen[r] == 0;
	for (int i = b; i < e; ++i) {
		unsigned long long ws;
		uint32_t fl;
		if (FMT == NGM_B200_DESC_PAIR16) {
			const ngm_b200_pair p = static_cast<const ngm_b200_pair *>(desc)[i];
			ws = p.window_start;
			fl = p.flags;
		} else {
			const unsigned long long d = static_cast<const unsigned long long *>(desc)[i];
			ws = d & 0x00FFFFFFFFFFFFFFull;
			fl = (uint32_t) (d >> 56);
		}
		PairDesc d;
		d.win_nib = ws < concat_len ? ws : n_region_nib;
		d.read_idx = (uint32_t) r;
		d.flags = (fl & (PF_REVERSE | PF_DIR | PF_INACTIVE)) | (empty ? PF_INACTIVE : 0u);
		rp[i] = d;
		if (pairs16 != nullptr) {                                  // top1PE reads ngm_b200_pair records
			ngm_b200_pair p;
			p.window_start = ws;
			p.read_index = (uint32_t) r;
			p.flags = fl;
			pairs16[i] = p;
		}
	}
	if (valid) fcnt[r] = cnt <= kFwdAllMax ? cnt : 1;
	if (r == n_reads) fcnt[r] = 0;
	const int want = cnt > kFwdAllMax ? cnt : 0;
	const int lane = threadIdx.x & 31;
	int incl = want;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const int v = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) incl += v;
	}
	int base = 0;
	if (lane == 31 && incl) base = atomicAdd(n_sel, incl);
	base = __shfl_sync(0xffffffffu, base, 31);
	const int at = base + incl - want;
	for (int i = 0; i < want; ++i) sel[at + i] = b + i;
}

// top-1 of the reads with more than kFwdAllMax candidates (they were scored), and the forward list: F[fbegin[r] ..) = all candidates of a
// small read / the winner of a big one
__global__ void __launch_bounds__(256) batch_fill_fwd_kernel(int n_reads, const int *__restrict__ cb, const PairDesc *__restrict__ rp, const float *__restrict__ scores,
		const int *__restrict__ fbegin, int strata, PairDesc *__restrict__ F, int *__restrict__ best_pair, int *__restrict__ mapq, int *__restrict__ num_top,
		int paired) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int b = cb[r], e = cb[r + 1], f0 = fbegin[r];
	if (e - b <= kFwdAllMax) {
		for (int i = b; i < e; ++i) F[f0 + (i - b)] = rp[i];
		if (e == b && !paired) {
			best_pair[r] = -1;
			mapq[r] = 0;
			if (num_top != nullptr) num_top[r] = 0;
		}
		return;
	}
	if (paired) {                                                 // the winner is only known after top1PE: an idle slot for now
		PairDesc d;
		d.win_nib = 0;
		d.read_idx = (uint32_t) r;
		d.flags = PF_INACTIVE;
		F[f0] = d;
		return;
	}
	float best = 0.0f, second = 0.0f;
	int besti = 0, nbest = 0;
	for (int j = b; j < e; ++j) {
		const float s = scores[j];
		if (s > second) {
			if (s > best) {
				second = best;
				best = s;
				besti = j - b;
				nbest = 1;
			} else if (s == best) {
				++nbest;
				second = best;
			} else {
				second = s;
			}
		} else if (s == best) {
			++nbest;
		}
	}
	int mq = 0;
	if (best > 0.0f && second >= 0.0f) mq = (int) ceilf(60.0f * (best - second) / best);
	int bp = b + besti;
	if (strata && nbest != 1) {
		bp = -1;
		mq = 0;
		nbest = 1;
	}
	best_pair[r] = bp;
	mapq[r] = mq;
	if (num_top != nullptr) num_top[r] = nbest;
	PairDesc d;
	if (bp >= 0) {
		d = rp[bp];
	} else {
		d.win_nib = 0;
		d.read_idx = (uint32_t) r;
		d.flags = PF_INACTIVE;
	}
	F[f0] = d;
}

// after the forward pass of a chunk of reads: scores of the small reads' candidates (= forward maxima), their top-1 + MAPQ, and the
// forward slot every read's alignment is backtraced from
template <int MODE>
__global__ void __launch_bounds__(256) batch_pick_kernel(int m, int r0, const int *__restrict__ cb, const int *__restrict__ fbegin, const PairDesc *__restrict__ F,
		const int4 *__restrict__ best_in, int strata, float *__restrict__ scores, int *__restrict__ best_pair, int *__restrict__ mapq, int *__restrict__ num_top,
		int *__restrict__ slot_of, int paired) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	const int r = r0 + i;
	const int b = cb[r], e = cb[r + 1], base = fbegin[r0], f0 = fbegin[r] - base;
	if (paired) {                                                 // scores only: the selection is top1PE's (batch_slot_pairs_kernel follows it)
		if (e - b <= kFwdAllMax)
			for (int j = 0; j < e - b; ++j)
				scores[b + j] = (F[base + f0 + j].flags & PF_INACTIVE) ? (MODE == 0 ? -1.0f : (float) kEndFreeMin) : (float) best_in[f0 + j].z;
		return;
	}
	if (e - b > kFwdAllMax) {
		slot_of[i] = best_pair[r] >= 0 ? f0 : -1;
		return;
	}
	if (e == b) {
		slot_of[i] = -1;
		return;
	}
	float best = 0.0f, second = 0.0f;
	int besti = 0, nbest = 0;
	for (int j = 0; j < e - b; ++j) {
		const bool inactive = (F[base + f0 + j].flags & PF_INACTIVE) != 0;
		const float s = inactive ? (MODE == 0 ? -1.0f : (float) kEndFreeMin) : (float) best_in[f0 + j].z;
		scores[b + j] = s;
		if (s > second) {
			if (s > best) {
				second = best;
				best = s;
				besti = j;
				nbest = 1;
			} else if (s == best) {
				++nbest;
				second = best;
			} else {
				second = s;
			}
		} else if (s == best) {
			++nbest;
		}
	}
	int mq = 0;
	if (best > 0.0f && second >= 0.0f) mq = (int) ceilf(60.0f * (best - second) / best);
	int bp = b + besti, sl = f0 + besti;
	if (strata && nbest != 1) {
		bp = -1;
		sl = -1;
		mq = 0;
		nbest = 1;
	}
	best_pair[r] = bp;
	mapq[r] = mq;
	if (num_top != nullptr) num_top[r] = nbest;
	slot_of[i] = sl;
}

// paired batches after top1PE of a chunk: the forward slot of every small read's winner; the winners of big reads (scored, not yet
// forwarded) go to the late list and are aligned in a second, usually empty, pass
__global__ void __launch_bounds__(256) batch_slot_pairs_kernel(int m, int r0, const int *__restrict__ cb, const int *__restrict__ fbegin, const PairDesc *__restrict__ rp,
		const int *__restrict__ best_pair, int *__restrict__ slot_of, PairDesc *__restrict__ late_pairs, int *__restrict__ late_item, int *__restrict__ late_range) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	const int r = r0 + i;
	const int b = cb[r], e = cb[r + 1], bp = best_pair[r];
	int sl = -1;
	if (bp >= 0) {
		if (e - b <= kFwdAllMax) {
			sl = fbegin[r] - fbegin[r0] + (bp - b);
		} else {
			const int k = atomicAdd(late_range + 1, 1);            // late_range = {0, number of late winners}
			late_pairs[k] = rp[bp];
			late_item[k] = i;
		}
	}
	slot_of[i] = sl;
}

__global__ void batch_set_u32_kernel(uint32_t *p, uint32_t v, int *q) {
	if (p != nullptr) *p = v;
	if (q != nullptr) *q = 0;
}

__global__ void __launch_bounds__(256) batch_rebase_kernel(int n, int *__restrict__ cb, int base) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) cb[i] -= base;
}

__global__ void __launch_bounds__(256) batch_add_base_kernel(int n, int *__restrict__ best_pair, int base) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && best_pair[i] >= 0) best_pair[i] += base;
}

// ---------------------------------------------------------------------------------------------------------
// per-lane staging + the lanes of a root context
// ---------------------------------------------------------------------------------------------------------
struct LaneBuf {
	DevBuf d_in_reads, d_in_len, d_in_exc, d_cb, d_desc, d_rp, d_pairs16, d_sel, d_nsel, d_scores, d_best, d_mapq, d_ntop, d_pfail, d_wp, d_wscores, d_obest,
			d_recs, d_strings, d_cursor, d_maxhit, d_fcnt, d_fbegin, d_F, d_slot, d_scan_tmp, d_tsel, d_tnsel, d_late_pairs, d_late_item, d_late_range;
	cudaEvent_t searched = nullptr;                                // ngm_b200_map_batch: candidate search of the lane's sub-batch has finished
	cudaEvent_t done = nullptr;
	int pending = -1;                                              // sub-batch whose strings still have to be fetched
	void release() {
		DevBuf *all[] = { &d_in_reads, &d_in_len, &d_in_exc, &d_cb, &d_desc, &d_rp, &d_pairs16, &d_sel, &d_nsel, &d_scores, &d_best, &d_mapq, &d_ntop, &d_pfail,
				&d_wp, &d_wscores, &d_obest, &d_recs, &d_strings, &d_cursor, &d_maxhit, &d_fcnt, &d_fbegin, &d_F, &d_slot, &d_scan_tmp, &d_tsel, &d_tnsel, &d_late_pairs, &d_late_item, &d_late_range };
		for (DevBuf *b : all) b->release();
		if (done) cudaEventDestroy(done);
		if (searched) cudaEventDestroy(searched);
		done = searched = nullptr;
	}
};

struct BatchState {
	int n_lanes = 4;                                               // measured on B200 (10 M reads): 3 x 1 Mi 327 M reads/s, 4 x 512 Ki 417 M, 4 x 256 Ki 411 M, 6 x 512 Ki 386 M
	int sub_batch = 1 << 19;
	std::vector<ngm_b200_ctx *> lanes;
	std::vector<LaneBuf> bufs;
	LaneBuf own;                                                   // staging of ngm_b200_dev_run_batch on the root context itself
	uint64_t synced_epoch = ~0ull;
	HostBuf h_used;                                                // pinned: heap cursor of every sub-batch
	HostBuf h_count;                                               // pinned: candidates found in every sub-batch (ngm_b200_map_batch)
	cudaEvent_t pe_chain = nullptr;                                // orders the paired-end selections of consecutive sub-batches
};

void batch_release(BatchState *b) {
	if (b == nullptr) return;
	for (LaneBuf &l : b->bufs) l.release();
	b->own.release();
	for (ngm_b200_ctx *l : b->lanes) ngm_b200_destroy(l);
	b->h_used.release();
	b->h_count.release();
	if (b->pe_chain) cudaEventDestroy(b->pe_chain);
	delete b;
}

}  // namespace ngm

namespace {

struct DevIn {                      // one sub-batch with every pointer on the device
	int n_reads, mode, paired, desc_format;
	const int *cb;                  // n_reads + 1 offsets, relative to cb_base
	int cb_base;
	const void *desc;               // descriptors of this sub-batch (index 0 = candidate cb_base)
	int n_pairs;
};

struct DevOut {
	float *scores;                  // n_pairs (always present on the device)
	int *best_pair, *mapq, *num_top, *pair_fail;
	ngm_b200_align_rec *recs;
	char *strings;                  // heap pointer such that strings + offset is valid for offsets in [str_base, str_cap)
	uint32_t str_cap;
	uint32_t *cursor;               // already holds str_base
	int topn = 1;                   // > 1 (single-end): recs holds n_reads x topn records, sel n_reads x topn candidate indices, n_sel n_reads counts
	int *sel = nullptr, *n_sel = nullptr;
};

// single-end batches on the s16x2 second-generation kernels: forward pass over every candidate of reads with <= kFwdAllMax candidates
// Paired batches take the same route: BatchScore results of every candidate = forward maxima, top1PE per chunk of reads (the running
// insert-size sums see the fragments in input order either way), backtrace of the winners; winners of reads with more than kFwdAllMax
// candidates are aligned in a second, usually empty, pass over a late list.
int enqueue_fwd_all(ngm_b200_ctx *c, LaneBuf &L, const DevIn &in, const DevOut &out, int m0, int strata, cudaStream_t st, cudaEvent_t pe_wait, cudaEvent_t pe_signal) {
	const int n = in.n_reads, np = in.n_pairs;
	const int paired = in.paired ? 1 : 0;
	const bool own_p16 = paired && in.desc_format != NGM_B200_DESC_PAIR16;
	if (own_p16) CU(L.d_pairs16.ensure(std::max<size_t>(np, 1) * sizeof(ngm_b200_pair)));
	ngm_b200_pair *p16 = own_p16 ? L.d_pairs16.as<ngm_b200_pair>() : nullptr;
	const ngm_b200_pair *pe_pairs = own_p16 ? p16 : static_cast<const ngm_b200_pair *>(in.desc);
	CU(L.d_rp.ensure(std::max<size_t>(np, 1) * sizeof(PairDesc)));
	CU(L.d_F.ensure(std::max<size_t>(np, 1) * sizeof(PairDesc)));
	CU(L.d_sel.ensure(std::max<size_t>(np, 1) * 4));
	CU(L.d_nsel.ensure(4));
	CU(L.d_fcnt.ensure(((size_t) n + 1) * 4));
	CU(L.d_fbegin.ensure(((size_t) n + 1) * 4));
	batch_set_u32_kernel<<<1, 1, 0, st>>>(nullptr, 0, L.d_nsel.as<int>());
	const int blocks_r = (n + 255) / 256, blocks_r1 = (n + 1 + 255) / 256;
	if (in.desc_format == NGM_B200_DESC_PAIR16)
		batch_plan_all_kernel<NGM_B200_DESC_PAIR16><<<blocks_r1, 256, 0, st>>>(n, in.cb, in.desc, L.d_rp.as<PairDesc>(), (unsigned long long) c->concat_len,
				(unsigned long long) c->n_region_nib, c->d_rrlen.as<uint16_t>(), L.d_fcnt.as<int>(), L.d_sel.as<int>(), L.d_nsel.as<int>(), p16);
	else
		batch_plan_all_kernel<NGM_B200_DESC_U64><<<blocks_r1, 256, 0, st>>>(n, in.cb, in.desc, L.d_rp.as<PairDesc>(), (unsigned long long) c->concat_len,
				(unsigned long long) c->n_region_nib, c->d_rrlen.as<uint16_t>(), L.d_fcnt.as<int>(), L.d_sel.as<int>(), L.d_nsel.as<int>(), p16);
	size_t tmp_bytes = 0;
	CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, L.d_fcnt.as<int>(), L.d_fbegin.as<int>(), n + 1, st));
	CU(L.d_scan_tmp.ensure(tmp_bytes));
	CU(cub::DeviceScan::ExclusiveSum(L.d_scan_tmp.p, tmp_bytes, L.d_fcnt.as<int>(), L.d_fbegin.as<int>(), n + 1, st));
	c->launches += 3;
	CU(cudaGetLastError());
	const uint32_t *rf = c->d_rfwd.as<uint32_t>(), *rr = c->d_rrev.as<uint32_t>(), *ref4 = c->d_ref4.as<uint32_t>();
	const uint16_t *rl = c->d_rrlen.as<uint16_t>();
	if (np > 0) {                                                  // reads with many candidates: BatchScore first (the work list may be empty)
		ScoreArgs a;
		a.P = c->dp;
		a.pairs = L.d_rp.as<PairDesc>();
		a.n = np;
		a.reads_fwd = rf;
		a.reads_rev = rr;
		a.rlen = rl;
		a.ref4 = ref4;
		a.out = out.scores;
		a.sel = L.d_sel.as<int>();
		a.n_dev = L.d_nsel.as<int>();
		int rc = run_score(c, m0, a, st);
		if (rc) return rc;
	}
	batch_fill_fwd_kernel<<<blocks_r, 256, 0, st>>>(n, in.cb, L.d_rp.as<PairDesc>(), out.scores, L.d_fbegin.as<int>(), strata, L.d_F.as<PairDesc>(), out.best_pair,
			out.mapq, out.num_top, paired);
	c->launches += 1;
	CU(cudaGetLastError());
	// chunks of G reads: at most kFwdAllMax * G forward slots, whose pointer matrix is the launch set's scratch
	const size_t per_slot = (size_t) c->dp.rows_cap * ptr_words_for(c->capacity) * 4 + sizeof(int4);
	size_t slots_budget = std::max<size_t>((size_t) kFwdAllMax * 4096, std::min<size_t>((size_t) 4 << 20, ((size_t) 11 << 30) / per_slot));
	int G = (int) std::min<size_t>((size_t) n, slots_budget / kFwdAllMax);
	if (paired && (G & 1)) G = std::max(2, G - 1);                 // mates stay in one chunk
	const int C = (kFwdAllMax * G + 255) / 256 * 256;
	const int Gpad = (G + 127) / 128 * 128;
	const int ops_cap = 2 * c->dp.qml + c->dp.corridor + 2;
	CU(c->d_ptr.ensure((size_t) c->dp.rows_cap * C * ptr_words_for(c->capacity) * sizeof(uint32_t)));
	CU(c->d_best.ensure((size_t) C * sizeof(int4)));
	CU(c->d_ops.ensure((size_t) ops_cap * Gpad * sizeof(uint16_t)));
	CU(L.d_slot.ensure((size_t) Gpad * 4));
	if (paired) {
		CU(L.d_late_pairs.ensure((size_t) Gpad * sizeof(PairDesc)));
		CU(L.d_late_item.ensure((size_t) Gpad * 4));
		CU(L.d_late_range.ensure(8));
		if (out.num_top == nullptr) CU(L.d_ntop.ensure((size_t) n * 4));
		// The previous sub-batch's selections come first.  Waiting HERE (before this sub-batch's forward pass, not just before its first
		// top1PE) keeps the lanes staggered: lane k + 1 runs its forward pass while lane k traces back and copies out.  With the wait in
		// front of the selection only, all lanes ran their forward passes interleaved and finished together: e2e 370 -> 336 M reads/s.
		if (pe_wait) CU(cudaStreamWaitEvent(st, pe_wait, 0));
	} else if (pe_wait) {
		CU(cudaStreamWaitEvent(st, pe_wait, 0));                   // single-end: the same staggering of the lanes (pe_signal follows the forward passes)
	}
	for (int r0 = 0; r0 < n; r0 += G) {
		const int m = std::min(G, n - r0);
		AlignArgs a;
		a.P = c->dp;
		a.pairs = L.d_F.as<PairDesc>();
		a.n = std::min(C, kFwdAllMax * m);
		a.reads_fwd = rf;
		a.reads_rev = rr;
		a.rlen = rl;
		a.ref4 = ref4;
		a.ptr_scratch = c->d_ptr.as<uint32_t>();
		a.ops_scratch = c->d_ops.as<uint16_t>();
		a.best_scratch = c->d_best.as<int4>();
		a.known = nullptr;
		a.stride = C;
		a.ops_cap = ops_cap;
		a.recs = out.recs + r0;
		a.strings = out.strings;
		a.str_cap = out.str_cap;
		a.cursor = out.cursor;
		a.range = L.d_fbegin.as<int>() + r0;
		a.range_m = m;
		a.phase = 1;
		// ngm_b200_profile: forward kernel | pick + backtrace kernel of every chunk between events (bench.py's in-step kernel split)
		const bool prof = c->profile && c->pev_used + 3 <= 3 * 64;
		if (prof) {
			for (int k = 0; k < 3; ++k)
				if (c->pev[c->pev_used + k] == nullptr) CU(cudaEventCreate(&c->pev[c->pev_used + k]));
			CU(cudaEventRecord(c->pev[c->pev_used], st));
		}
		cudaError_t e = launch_align_s16(c->capacity, m0, a, st);
		if (e != cudaSuccess) return fail(NGM_B200_ECUDA, "forward kernel launch: %s", cudaGetErrorString(e));
		if (prof) CU(cudaEventRecord(c->pev[c->pev_used + 1], st));
		static const int stagger_mode = [] { const char *e = getenv("NGM_B200_STAGGER"); return e == nullptr ? 1 : atoi(e); }();
		if (!paired && pe_signal && r0 + G >= n && stagger_mode != 2) CU(cudaEventRecord(pe_signal, st));      // the next lane's forward pass may start
		if (m0 == 0)
			batch_pick_kernel<0><<<(m + 255) / 256, 256, 0, st>>>(m, r0, in.cb, L.d_fbegin.as<int>(), L.d_F.as<PairDesc>(), c->d_best.as<int4>(), strata, out.scores,
					out.best_pair, out.mapq, out.num_top, L.d_slot.as<int>(), paired);
		else
			batch_pick_kernel<1><<<(m + 255) / 256, 256, 0, st>>>(m, r0, in.cb, L.d_fbegin.as<int>(), L.d_F.as<PairDesc>(), c->d_best.as<int4>(), strata, out.scores,
					out.best_pair, out.mapq, out.num_top, L.d_slot.as<int>(), paired);
		if (paired) {
			// top1PE of the chunk's fragments (mates stay together: G is even), then the winners' slots
			int *ntop = out.num_top != nullptr ? out.num_top : L.d_ntop.as<int>();
			int rc = ngm_b200_dev_select_pairs(c, m, in.cb + r0, pe_pairs, out.scores, (uint32_t) np, out.best_pair + r0, out.mapq + r0, ntop + r0, out.pair_fail + r0, st);
			if (rc < 0) return rc;
			batch_set_u32_kernel<<<1, 1, 0, st>>>(L.d_late_range.as<uint32_t>(), 0u, L.d_late_range.as<int>() + 1);
			batch_slot_pairs_kernel<<<(m + 255) / 256, 256, 0, st>>>(m, r0, in.cb, L.d_fbegin.as<int>(), L.d_rp.as<PairDesc>(), out.best_pair, L.d_slot.as<int>(),
					L.d_late_pairs.as<PairDesc>(), L.d_late_item.as<int>(), L.d_late_range.as<int>());
			c->launches += 2;
		}
		a.phase = 2;
		a.n = m;
		a.n_items = m;
		a.slot_of = L.d_slot.as<int>();
		a.ops_stride = Gpad;
		e = launch_align_s16(c->capacity, m0, a, st);
		if (e != cudaSuccess) return fail(NGM_B200_ECUDA, "backtrace kernel launch: %s", cudaGetErrorString(e));
		if (prof) {
			CU(cudaEventRecord(c->pev[c->pev_used + 2], st));
			c->pev_used += 3;
		}
		c->launches += 3;
		if (paired) {
			// the late list: winners of reads with more than kFwdAllMax candidates; same scratch (the chunk's backtrace has been enqueued),
			// the number of entries is only known on the device
			AlignArgs b = a;
			b.pairs = L.d_late_pairs.as<PairDesc>();
			b.n = m;
			b.range = L.d_late_range.as<int>();
			b.range_m = 1;
			b.phase = 1;
			e = launch_align_s16(c->capacity, m0, b, st);
			if (e != cudaSuccess) return fail(NGM_B200_ECUDA, "forward kernel launch (late list): %s", cudaGetErrorString(e));
			b.phase = 2;
			b.n_items = m;
			b.items_dev = L.d_late_range.as<int>() + 1;
			b.slot_of = nullptr;
			b.rec_of = L.d_late_item.as<int>();
			e = launch_align_s16(c->capacity, m0, b, st);
			if (e != cudaSuccess) return fail(NGM_B200_ECUDA, "backtrace kernel launch (late list): %s", cudaGetErrorString(e));
			c->launches += 2;
		}
	}
	if (paired && pe_signal) CU(cudaEventRecord(pe_signal, st));
	{
		// NGM_B200_STAGGER=2 (experiment): the next lane's kernels start only after this lane's backtrace
		static const int stagger_mode2 = [] { const char *e = getenv("NGM_B200_STAGGER"); return e == nullptr ? 1 : atoi(e); }();
		if (!paired && pe_signal && stagger_mode2 == 2) CU(cudaEventRecord(pe_signal, st));
	}
	CU(cudaGetLastError());
	return NGM_B200_OK;
}

// single-end batches with topn > 1 (NGM -n / --strata; ScoreBuffer::topNSE, ScoreBuffer.cpp:279-330): BatchScore of every candidate, the
// sorted selection, BatchAlign of up to topn candidates per read
int enqueue_topn(ngm_b200_ctx *c, LaneBuf &L, const DevIn &in, const DevOut &out, int m0, int strata, cudaStream_t st) {
	const int n = in.n_reads, np = in.n_pairs, topn = out.topn;
	if (in.cb_base != 0) return fail(NGM_B200_EINVAL, "topn batches take candidate offsets relative to the sub-batch");
	if ((long long) n * topn > 0x7FFFFFFFll) return fail(NGM_B200_EINVAL, "n_reads x topn beyond 31 bits");
	const size_t items = (size_t) n * topn;
	CU(L.d_rp.ensure(std::max<size_t>(np, 1) * sizeof(PairDesc)));
	CU(L.d_wp.ensure(items * sizeof(PairDesc)));
	CU(L.d_wscores.ensure(items * 4));
	CU(L.d_nsel.ensure(4));
	const int blocks_r = (n + 255) / 256;
	if (in.desc_format == NGM_B200_DESC_PAIR16)
		batch_plan_kernel<NGM_B200_DESC_PAIR16><<<blocks_r, 256, 0, st>>>(n, in.cb, 0, in.desc, L.d_rp.as<PairDesc>(), nullptr, (unsigned long long) c->concat_len,
				(unsigned long long) c->n_region_nib, c->d_rrlen.as<uint16_t>(), 0, nullptr, nullptr);
	else
		batch_plan_kernel<NGM_B200_DESC_U64><<<blocks_r, 256, 0, st>>>(n, in.cb, 0, in.desc, L.d_rp.as<PairDesc>(), nullptr, (unsigned long long) c->concat_len,
				(unsigned long long) c->n_region_nib, c->d_rrlen.as<uint16_t>(), 0, nullptr, nullptr);
	c->launches += 1;
	CU(cudaGetLastError());
	const uint32_t *rf = c->d_rfwd.as<uint32_t>(), *rr = c->d_rrev.as<uint32_t>(), *ref4 = c->d_ref4.as<uint32_t>();
	const uint16_t *rl = c->d_rrlen.as<uint16_t>();
	if (np > 0) {
		ScoreArgs a = score_args(c, L.d_rp.as<PairDesc>(), np, rf, rr, rl, ref4, out.scores);
		int rc = run_score(c, m0, a, st);
		if (rc) return rc;
	}
	int rc = ngm_b200_dev_select_topn(c, n, in.cb, out.scores, (uint32_t) np, topn, strata, out.sel, out.n_sel, out.mapq, out.num_top != nullptr ? out.num_top : L.d_ntop.p, st);
	if (rc < 0) return rc;
	batch_gather_topn_kernel<<<(unsigned) ((items + 255) / 256), 256, 0, st>>>(n, topn, L.d_rp.as<PairDesc>(), out.sel, out.scores, L.d_wp.as<PairDesc>(),
			L.d_wscores.as<float>(), out.best_pair);
	c->launches += 1;
	CU(cudaGetLastError());
	return run_align(c, m0, L.d_wp.as<PairDesc>(), (int) items, rf, rr, rl, ref4, out.recs, out.strings, out.str_cap, out.cursor, st, L.d_wscores.as<float>(), nullptr);
}

// enqueue score -> select -> align of one sub-batch whose reads are installed in `c` (d_rfwd / d_rrev / d_rrlen)
int batch_enqueue(ngm_b200_ctx *c, LaneBuf &L, const DevIn &in, const DevOut &out, int strata, cudaStream_t st, cudaEvent_t pe_wait, cudaEvent_t pe_signal) {
	const int n = in.n_reads, np = in.n_pairs;
	const int m0 = mode_of(in.mode);
	if (m0 < 0) return fail(NGM_B200_EINVAL, "unsupported alignment mode %d", in.mode & 0xFF);
	if (!c->have_ref) return fail(NGM_B200_ESTATE, "set_reference must precede ngm_b200_run_batch");
	if (out.topn > 1) {
		if (in.paired) return fail(NGM_B200_EINVAL, "topn %d: paired runs report one alignment per mate (ScoreBuffer.cpp:365-502)", out.topn);
		if (out.sel == nullptr || out.n_sel == nullptr) return fail(NGM_B200_EINVAL, "topn %d needs the sel / n_sel arrays", out.topn);
		return enqueue_topn(c, L, in, out, m0, strata, st);
	}
	// single-candidate reads skip BatchScore when the alignment kernel of this configuration reports its forward maximum and does not
	// need the score beforehand (wide local bands locate their best cell by it)
	const bool wide_local = m0 == 0 && c->align_s16[0] && c->capacity > kAlignS16MaxLocal && !wide_v2_enabled();
	static const bool no_fuse = [] { const char *e = getenv("NGM_B200_NO_FUSE"); return e != nullptr && atoi(e) == 1; }();
	const int fuse = (!in.paired && !wide_local && !no_fuse) ? 1 : 0;
	// NGM_B200_FWD_ALL=0 keeps score -> top-1 -> forward of the winner for every multi-candidate read (A/B measurements)
	static const bool fwd_all_on = [] { const char *e = getenv("NGM_B200_FWD_ALL"); return e == nullptr || atoi(e) != 0; }();
	static const bool fwd_v1 = [] { const char *e = getenv("NGM_B200_FWD"); return e != nullptr && atoi(e) == 1; }();
	const bool s16_fwd2 = c->align_s16[m0] && c->capacity <= (m0 == 1 ? kAlignS16MaxEndFree : (wide_v2_enabled() ? kAlignS16MaxKnown : kAlignS16MaxLocal));
	static const bool pe_fwd_all = [] { const char *e = getenv("NGM_B200_PE_FWD_ALL"); return e == nullptr || atoi(e) != 0; }();
	if ((fuse || (in.paired && pe_fwd_all && !wide_local && !no_fuse)) && fwd_all_on && !fwd_v1 && s16_fwd2 && in.cb_base == 0)
		return enqueue_fwd_all(c, L, in, out, m0, strata, st, pe_wait, pe_signal);
	CU(L.d_rp.ensure(std::max<size_t>(np, 1) * sizeof(PairDesc)));
	CU(L.d_wp.ensure((size_t) n * sizeof(PairDesc)));
	CU(L.d_wscores.ensure((size_t) n * 4));
	CU(L.d_obest.ensure((size_t) n * 4));
	CU(L.d_nsel.ensure(4));
	if (fuse) CU(L.d_sel.ensure(std::max<size_t>(np, 1) * 4));
	const bool own_p16 = in.paired && in.desc_format != NGM_B200_DESC_PAIR16;      // top1PE reads ngm_b200_pair records: expand 64-bit descriptors
	if (own_p16) CU(L.d_pairs16.ensure(std::max<size_t>(np, 1) * sizeof(ngm_b200_pair)));
	batch_set_u32_kernel<<<1, 1, 0, st>>>(nullptr, 0, L.d_nsel.as<int>());
	const int blocks_r = (n + 255) / 256;
	ngm_b200_pair *p16 = own_p16 ? L.d_pairs16.as<ngm_b200_pair>() : nullptr;
	const ngm_b200_pair *pe_pairs = own_p16 ? p16 : static_cast<const ngm_b200_pair *>(in.desc);
	if (in.desc_format == NGM_B200_DESC_PAIR16)
		batch_plan_kernel<NGM_B200_DESC_PAIR16><<<blocks_r, 256, 0, st>>>(n, in.cb, in.cb_base, in.desc, L.d_rp.as<PairDesc>(), p16, (unsigned long long) c->concat_len,
				(unsigned long long) c->n_region_nib, c->d_rrlen.as<uint16_t>(), fuse, L.d_sel.as<int>(), L.d_nsel.as<int>());
	else
		batch_plan_kernel<NGM_B200_DESC_U64><<<blocks_r, 256, 0, st>>>(n, in.cb, in.cb_base, in.desc, L.d_rp.as<PairDesc>(), p16, (unsigned long long) c->concat_len,
				(unsigned long long) c->n_region_nib, c->d_rrlen.as<uint16_t>(), fuse, L.d_sel.as<int>(), L.d_nsel.as<int>());
	c->launches += 2;
	CU(cudaGetLastError());
	const uint32_t *rf = c->d_rfwd.as<uint32_t>(), *rr = c->d_rrev.as<uint32_t>(), *ref4 = c->d_ref4.as<uint32_t>();
	const uint16_t *rl = c->d_rrlen.as<uint16_t>();
	if (np > 0) {
		ScoreArgs a;
		a.P = c->dp;
		a.pairs = L.d_rp.as<PairDesc>();
		a.n = np;
		a.reads_fwd = rf;
		a.reads_rev = rr;
		a.rlen = rl;
		a.ref4 = ref4;
		a.out = out.scores;
		if (fuse) {
			a.sel = L.d_sel.as<int>();
			a.n_dev = L.d_nsel.as<int>();
		}
		int rc = run_score(c, m0, a, st);
		if (rc) return rc;
	}
	if (in.paired) {
		if (pe_wait) CU(cudaStreamWaitEvent(st, pe_wait, 0));
		int rc = ngm_b200_dev_select_pairs(c, n, in.cb, pe_pairs, out.scores, (uint32_t) np, out.best_pair, out.mapq, out.num_top != nullptr ? out.num_top : L.d_ntop.p,
				out.pair_fail, st);
		if (rc < 0) return rc;
		if (pe_signal) CU(cudaEventRecord(pe_signal, st));
	} else {
		batch_select_top1_kernel<<<blocks_r, 256, 0, st>>>(n, in.cb, in.cb_base, out.scores, fuse, strata, out.best_pair, out.mapq, out.num_top);
		c->launches += 1;
	}
	batch_gather_kernel<<<blocks_r, 256, 0, st>>>(n, in.cb, in.cb_base, L.d_rp.as<PairDesc>(), out.best_pair, out.scores, fuse, L.d_wp.as<PairDesc>(),
			L.d_wscores.as<float>());
	c->launches += 1;
	CU(cudaGetLastError());
	int rc = run_align(c, m0, L.d_wp.as<PairDesc>(), n, rf, rr, rl, ref4, out.recs, out.strings, out.str_cap, out.cursor, st,
			fuse ? nullptr : L.d_wscores.as<float>(), fuse ? L.d_obest.as<float>() : nullptr);
	if (rc) return rc;
	if (fuse) {
		batch_finalize_kernel<<<blocks_r, 256, 0, st>>>(n, in.cb, in.cb_base, L.d_obest.as<float>(), strata, out.scores, out.best_pair, out.mapq, out.num_top);
		c->launches += 1;
	}
	CU(cudaGetLastError());
	return NGM_B200_OK;
}

// install one sub-batch of reads (device pointers) in context c
int install_reads(ngm_b200_ctx *c, int format, const void *d_reads, int n, int stride, const uint16_t *d_len, const ngm_b200_read_exc *d_exc, uint32_t n_exc,
		uint32_t read_base, cudaStream_t st) {
	if (format == NGM_B200_READS_ASCII) return pack_reads_device(c, static_cast<const uint8_t *>(d_reads), n, stride, st);
	if (format != NGM_B200_READS_PACKED2) return fail(NGM_B200_EINVAL, "unknown read format %d", format);
	if (stride <= 0 || (stride & 3)) return fail(NGM_B200_EINVAL, "PACKED2 rows must be a multiple of 4 bytes (got %d)", stride);
	if (d_len == nullptr) return fail(NGM_B200_EINVAL, "PACKED2 reads need read_len");
	const int RW = c->dp.read_words;
	CU(c->d_rfwd.ensure((size_t) n * RW * 4));
	CU(c->d_rrev.ensure((size_t) n * RW * 4));
	CU(c->d_rrlen.ensure((size_t) n * 2));
	// idx / RW as mulhi(idx, ceil(2^32 / RW)) is exact while idx * e < 2^32, e = RW - 2^32 mod RW: launches of at most that many words
	const uint32_t magic = (uint32_t) ((0x100000000ull + (unsigned long long) RW - 1) / (unsigned long long) RW);
	const unsigned long long e = (unsigned long long) RW - (0x100000000ull % (unsigned long long) RW);
	const int rows_per = (int) std::min<unsigned long long>(0x7FFFFFFFull / RW, (0xFFFFFFFFull / e) / RW);
	for (int r0 = 0; r0 < n; r0 += rows_per) {
		const int m = std::min(rows_per, n - r0);
		expand_packed2_kernel<<<(unsigned) (((size_t) m * RW + kExpandThreads - 1) / kExpandThreads), kExpandThreads, 0, st>>>(
				static_cast<const uint32_t *>(d_reads) + (size_t) r0 * (stride / 4), m, stride / 4, d_len + r0, c->dp.qml, c->d_rfwd.as<uint32_t>() + (size_t) r0 * RW,
				c->d_rrev.as<uint32_t>() + (size_t) r0 * RW, c->d_rrlen.as<uint16_t>() + r0, RW, magic);
	}
	c->launches += 1;
	if (n_exc) {
		patch_exceptions_kernel<<<(n_exc + 255) / 256, 256, 0, st>>>(d_exc, n_exc, read_base, n, c->d_rrlen.as<uint16_t>(), c->d_rfwd.as<uint32_t>(),
				c->d_rrev.as<uint32_t>(), RW);
		c->launches += 1;
	}
	CU(cudaGetLastError());
	c->n_reads = n;
	c->reads_stride = 0;
	return NGM_B200_OK;
}

// a child context (lane / shared context) takes over the root's resident data: packed reference, k-mer index, selection parameters
int borrow_from_root(ngm_b200_ctx *l, ngm_b200_ctx *root) {
	l->d_ref4.borrow(root->d_ref4);
	l->concat_len = root->concat_len;
	l->n_region_nib = root->n_region_nib;
	l->have_ref = root->have_ref;
	l->se_strata = root->se_strata;
	l->se_topn = root->se_topn;
	int rc = cs_share_index(l, root);
	if (rc) return rc;
	rc = pe_share_state(l, root);
	if (rc) return rc;
	l->root_epoch = root->epoch;
	return NGM_B200_OK;
}

int sync_lanes(ngm_b200_ctx *c) {
	if (c->batch == nullptr) c->batch = new BatchState();
	BatchState *B = c->batch;
	while ((int) B->lanes.size() < B->n_lanes) {
		ngm_b200_params hp = c->hp;
		ngm_b200_ctx *l = ngm_b200_create(&hp);
		if (l == nullptr) return NGM_B200_ECUDA;
		l->root = c;
		B->lanes.push_back(l);
		B->bufs.emplace_back();
		CU(cudaEventCreateWithFlags(&B->bufs.back().done, cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&B->bufs.back().searched, cudaEventDisableTiming));
		B->synced_epoch = ~0ull;
	}
	if (B->pe_chain == nullptr) CU(cudaEventCreateWithFlags(&B->pe_chain, cudaEventDisableTiming));
	if (B->synced_epoch != c->epoch) {
		for (ngm_b200_ctx *l : B->lanes) {
			int rc = borrow_from_root(l, c);
			if (rc) return rc;
		}
		B->synced_epoch = c->epoch;
	}
	return NGM_B200_OK;
}

}  // namespace

namespace ngm {
// used by ngm_map.cu: the lanes of a root context, created / re-synchronised on demand
int batch_lanes(ngm_b200_ctx *c, ngm_b200_ctx ***lanes, int *n_lanes, int *sub_batch) {
	int rc = sync_lanes(c);
	if (rc) return rc;
	*lanes = c->batch->lanes.data();
	*n_lanes = (int) c->batch->lanes.size();
	*sub_batch = c->batch->sub_batch;
	return NGM_B200_OK;
}
uint64_t batch_lane_launches(const ngm_b200_ctx *c) {
	uint64_t n = 0;
	if (c->batch)
		for (const ngm_b200_ctx *l : c->batch->lanes) n += l->launches;
	return n;
}
}  // namespace ngm

extern "C" {

ngm_b200_ctx *ngm_b200_create_shared(ngm_b200_ctx *root) {
	if (root == nullptr || root->root != nullptr) {
		fail(NGM_B200_EINVAL, "ngm_b200_create_shared needs a root context");
		return nullptr;
	}
	ngm_b200_params hp = root->hp;
	ngm_b200_ctx *l = ngm_b200_create(&hp);
	if (l == nullptr) return nullptr;
	l->root = root;
	if (borrow_from_root(l, root) != NGM_B200_OK) {
		ngm_b200_destroy(l);
		return nullptr;
	}
	return l;
}

int ngm_b200_sync_shared(ngm_b200_ctx *c) {
	if (c == nullptr) return fail(NGM_B200_EINVAL, "ctx == NULL");
	if (c->root == nullptr || c->root_epoch == c->root->epoch) return NGM_B200_OK;
	return borrow_from_root(c, c->root);
}

int ngm_b200_set_pipeline(ngm_b200_ctx *c, int lanes, int sub_batch_reads) {
	if (c == nullptr) return fail(NGM_B200_EINVAL, "ctx == NULL");
	if (lanes < 1 || lanes > 8) return fail(NGM_B200_EINVAL, "lanes %d not in [1, 8]", lanes);
	if (sub_batch_reads < 2 || (sub_batch_reads & 1)) return fail(NGM_B200_EINVAL, "sub-batches hold an even number of reads (mates stay together), got %d", sub_batch_reads);
	if (c->batch == nullptr) c->batch = new BatchState();
	c->batch->n_lanes = lanes;
	c->batch->sub_batch = sub_batch_reads;
	return NGM_B200_OK;
}

int ngm_b200_se_configure(ngm_b200_ctx *c, int strata) {
	if (c == nullptr) return fail(NGM_B200_EINVAL, "ctx == NULL");
	c->se_strata = strata ? 1 : 0;
	c->epoch += 1;
	return NGM_B200_OK;
}

int ngm_b200_se_configure_topn(ngm_b200_ctx *c, int topn) {
	if (c == nullptr) return fail(NGM_B200_EINVAL, "ctx == NULL");
	if (topn < 1 || topn > 1000) return fail(NGM_B200_EINVAL, "topn %d not in [1, 1000]", topn);      // GenericReadWriter.h:79 MAX_PASSED
	c->se_topn = topn;
	c->epoch += 1;
	return NGM_B200_OK;
}

void *ngm_b200_host_alloc(size_t bytes) {
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
		fail(NGM_B200_ECUDA, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(cudaGetLastError()));
		return nullptr;
	}
	return p;
}

void ngm_b200_host_free(void *p) {
	if (p) cudaFreeHost(p);
}

int ngm_b200_host_register(void *p, size_t bytes) {
	if (p == nullptr || bytes == 0) return fail(NGM_B200_EINVAL, "NULL / empty range");
	CU(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
	return NGM_B200_OK;
}

int ngm_b200_host_unregister(void *p) {
	if (p == nullptr) return fail(NGM_B200_EINVAL, "NULL");
	CU(cudaHostUnregister(p));
	return NGM_B200_OK;
}

int ngm_b200_pack_reads(const char *ascii, int n_reads, int stride, void *packed, int row_bytes, uint16_t *read_len, ngm_b200_read_exc *exceptions,
		size_t exc_cap, size_t *n_exc, int threads) {
	if (ascii == nullptr || packed == nullptr || read_len == nullptr || n_reads < 0 || stride <= 0) return fail(NGM_B200_EINVAL, "bad argument");
	if ((row_bytes & 3) || row_bytes < 4 * ((stride + 15) / 16)) return fail(NGM_B200_EINVAL, "row_bytes %d too small for %d bases", row_bytes, stride);
	if (stride > 65535) return fail(NGM_B200_EINVAL, "rows longer than 65535 bases");
	int nt = threads > 0 ? threads : (int) std::thread::hardware_concurrency();
	nt = std::max(1, std::min(nt, std::max(1, n_reads / 4096)));
	// pass 1: pack + count the exceptions of every slice; pass 2 (only when there are any): write them in order, visiting the flagged reads only.
	// Eight bases at a time: the code is bits 1 and 2 of the letter ((c >> 1 ^ c >> 2) & 3 = A 0, C 1, G 2, T 3, either case), a group of eight
	// valid letters is folded into 16 bits with three shift-or steps; groups with any other byte and the tail of a read go base by base.
	std::vector<size_t> cnt((size_t) nt + 1, 0);
	std::vector<uint8_t> flagged;
	try {
		flagged.assign((size_t) n_reads, 0);
	} catch (...) {
		return fail(NGM_B200_EINVAL, "out of host memory");
	}
	constexpr uint64_t k01 = 0x0101010101010101ull, k7F = 0x7F7F7F7F7F7F7F7Full, k80 = 0x8080808080808080ull;
	auto is_byte = [&](uint64_t u, unsigned char ch) -> uint64_t {      // 0x80 in every byte of u that equals ch
		const uint64_t z = u ^ (k01 * ch);
		return ~(((z & k7F) + k7F) | z | k7F);
	};
	auto all_acgt = [&](uint64_t x) -> bool {
		const uint64_t u = x & 0xDFDFDFDFDFDFDFDFull;
		return (is_byte(u, 'A') | is_byte(u, 'C') | is_byte(u, 'G') | is_byte(u, 'T')) == k80;
	};
	auto pack_slice = [&](int t, bool emit, size_t off) {
		const int lo = (int) ((long long) n_reads * t / nt), hi = (int) ((long long) n_reads * (t + 1) / nt);
		size_t k = 0;
		for (int r = lo; r < hi; ++r) {
			const unsigned char *s = reinterpret_cast<const unsigned char *>(ascii) + (size_t) r * stride;
			if (!emit) {
				uint32_t *o = reinterpret_cast<uint32_t *>(static_cast<char *>(packed) + (size_t) r * row_bytes);
				int len = stride;
				while (len > 0 && s[len - 1] == 0) --len;              // MappedRead::length: index of the last non-NUL byte + 1
				size_t bad = 0;
				const int words = row_bytes / 4;
				for (int w = 0; w < words; ++w) {
					uint32_t v = 0;
					for (int half = 0; half < 2; ++half) {
						const int i0 = 16 * w + 8 * half;
						if (i0 >= len) break;
						uint32_t g = 0;
						uint64_t x;
						if (i0 + 8 <= len && (memcpy(&x, s + i0, 8), all_acgt(x))) {
							uint64_t c2 = ((x >> 1) ^ (x >> 2)) & 0x0303030303030303ull;
							c2 = (c2 | (c2 >> 6)) & 0x000F000F000F000Full;
							c2 = (c2 | (c2 >> 12)) & 0x000000FF000000FFull;
							g = (uint32_t) ((c2 | (c2 >> 24)) & 0xFFFFull);
						} else {
							for (int i = 0; i < 8 && i0 + i < len; ++i) {
								const unsigned char ch = s[i0 + i], u = ch & 0xDF;
								if (u == 'A' || u == 'C' || u == 'G' || u == 'T') g |= (uint32_t) (((ch >> 1) ^ (ch >> 2)) & 3u) << (2 * i);
								else ++bad;                            // code 0; the byte itself travels in the exception list
							}
						}
						v |= g << (16 * half);
					}
					o[w] = v;
				}
				read_len[r] = (uint16_t) len;
				if (bad) {
					flagged[(size_t) r] = 1;
					k += bad;
				}
			} else if (flagged[(size_t) r]) {
				const int len = (int) read_len[r];
				for (int i = 0; i < len; ++i) {
					const unsigned char u = s[i] & 0xDF;
					if (!(u == 'A' || u == 'C' || u == 'G' || u == 'T')) {
						if (off + k < exc_cap) {
							ngm_b200_read_exc x;
							x.read_index = (uint32_t) r;
							x.pos = (uint16_t) i;
							x.ch = s[i];
							x.pad = 0;
							exceptions[off + k] = x;
						}
						++k;
					}
				}
			}
		}
		if (!emit) cnt[(size_t) t + 1] = k;
	};
	{
		std::vector<std::thread> th;
		for (int t = 1; t < nt; ++t) th.emplace_back(pack_slice, t, false, (size_t) 0);
		pack_slice(0, false, 0);
		for (auto &x : th) x.join();
	}
	for (int t = 0; t < nt; ++t) cnt[(size_t) t + 1] += cnt[t];
	const size_t total = cnt[nt];
	if (n_exc) *n_exc = total;
	if (total == 0) return n_reads;
	if (exceptions == nullptr || total > exc_cap) return fail(NGM_B200_ERANGE, "exception list too small: %zu entries needed", total);
	{
		std::vector<std::thread> th;
		for (int t = 1; t < nt; ++t)
			if (cnt[(size_t) t + 1] != cnt[t]) th.emplace_back(pack_slice, t, true, cnt[t]);
		if (cnt[1] != cnt[0]) pack_slice(0, true, 0);
		for (auto &x : th) x.join();
	}
	return n_reads;
}

int ngm_b200_dev_set_reads_packed(ngm_b200_ctx *c, const void *d_packed, int n_reads, int stride, const void *d_read_len, const void *d_exceptions,
		uint32_t n_exceptions, void *stream) {
	if (c == nullptr || d_packed == nullptr || d_read_len == nullptr || n_reads <= 0) return fail(NGM_B200_EINVAL, "bad read batch");
	if (n_exceptions && d_exceptions == nullptr) return fail(NGM_B200_EINVAL, "NULL exception list");
	CU(cudaSetDevice(c->device));
	return install_reads(c, NGM_B200_READS_PACKED2, d_packed, n_reads, stride, static_cast<const uint16_t *>(d_read_len),
			static_cast<const ngm_b200_read_exc *>(d_exceptions), n_exceptions, 0, static_cast<cudaStream_t>(stream));
}

int ngm_b200_dev_run_batch(ngm_b200_ctx *c, const ngm_b200_batch_in *in, ngm_b200_batch_out *out, void *stream) {
	if (c == nullptr || in == nullptr || out == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	if (in->n_reads <= 0) return 0;
	if (in->reads == nullptr || in->cand_begin == nullptr || out->best_pair == nullptr || out->mapq == nullptr || out->recs == nullptr ||
			out->strings == nullptr || out->d_str_cursor == nullptr || (in->paired && out->pair_fail == nullptr))
		return fail(NGM_B200_EINVAL, "NULL array");
	if (in->paired && (in->n_reads & 1)) return fail(NGM_B200_EINVAL, "paired batches hold the mates in rows 2f and 2f + 1: %d rows", in->n_reads);
	if (out->str_capacity > 0xFFFFFFFFull) return fail(NGM_B200_EINVAL, "string heap beyond 32 bits");
	CU(cudaSetDevice(c->device));
	if (c->batch == nullptr) c->batch = new BatchState();
	LaneBuf &L = c->batch->own;
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	// the number of candidates sizes the staging: it is the caller's (reserved > 0) or read back once
	int np = (int) in->n_desc;
	if (np <= 0) {
		CU(cudaMemcpyAsync(&np, in->cand_begin + in->n_reads, 4, cudaMemcpyDeviceToHost, st));
		CU(cudaStreamSynchronize(st));
	}
	if (np > 0 && in->desc == nullptr) return fail(NGM_B200_EINVAL, "NULL descriptor array");
	int rc = install_reads(c, in->read_format, in->reads, in->n_reads, in->read_stride, in->read_len, in->exceptions, in->n_exceptions, 0, st);
	if (rc) return rc;
	float *scores = out->scores;
	if (scores == nullptr) {
		CU(L.d_scores.ensure(std::max<size_t>(np, 1) * 4));
		scores = L.d_scores.as<float>();
	}
	if (out->num_top == nullptr) CU(L.d_ntop.ensure((size_t) in->n_reads * 4));
	batch_set_u32_kernel<<<1, 1, 0, st>>>(out->d_str_cursor, 0u, nullptr);
	DevIn di = { in->n_reads, in->mode, in->paired, in->desc_format, in->cand_begin, 0, in->desc, np };
	DevOut dn = { scores, out->best_pair, out->mapq, out->num_top, out->pair_fail, out->recs, out->strings, (uint32_t) out->str_capacity, out->d_str_cursor };
	if (!in->paired && c->se_topn > 1) {
		dn.topn = c->se_topn;
		dn.sel = out->sel;
		dn.n_sel = out->n_sel;
	}
	rc = batch_enqueue(c, L, di, dn, c->se_strata, st, nullptr, nullptr);
	return rc ? rc : in->n_reads;
}

int ngm_b200_run_batch(ngm_b200_ctx *c, const ngm_b200_batch_in *in, ngm_b200_batch_out *out) {
	if (c == nullptr || in == nullptr || out == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	const int n = in->n_reads;
	if (n <= 0) return 0;
	if (in->reads == nullptr || in->cand_begin == nullptr || out->best_pair == nullptr || out->mapq == nullptr || out->recs == nullptr ||
			(out->str_capacity && out->strings == nullptr) || (in->paired && out->pair_fail == nullptr))
		return fail(NGM_B200_EINVAL, "NULL array");
	if (in->paired && (n & 1)) return fail(NGM_B200_EINVAL, "paired batches hold the mates in rows 2f and 2f + 1: %d rows", n);
	if (in->read_format == NGM_B200_READS_PACKED2 && in->read_len == nullptr) return fail(NGM_B200_EINVAL, "PACKED2 reads need read_len");
	if (in->read_stride <= 0) return fail(NGM_B200_EINVAL, "read_stride %d", in->read_stride);
	const int topn = in->paired ? 1 : c->se_topn;
	if (topn > 1 && (out->sel == nullptr || out->n_sel == nullptr)) return fail(NGM_B200_EINVAL, "topn %d needs the sel / n_sel arrays", topn);
	const int32_t *cb = in->cand_begin;
	const long long total_pairs = (long long) cb[n] - cb[0];
	if (total_pairs < 0 || (total_pairs > 0 && in->desc == nullptr)) return fail(NGM_B200_EINVAL, "bad candidate lists");
	CU(cudaSetDevice(c->device));
	int rc = sync_lanes(c);
	if (rc) return rc;
	BatchState *B = c->batch;
	const int SB = B->sub_batch;
	// Sub-batches: full ones in the middle, a ramp at both ends (SB/4, SB/2 ... SB/2, SB/4).  The first upload and the last download are
	// the only copies nothing overlaps with; small sub-batches there cut the pipeline's fill and drain (NGM_B200_RAMP=0: equal sizes).
	static const bool ramp = [] { const char *e = getenv("NGM_B200_RAMP"); return e == nullptr || atoi(e) != 0; }();
	std::vector<int> sub_r0;                                        // first read of every sub-batch, and n at the end
	{
		const int q = std::max(2, (SB / 4) & ~1), h = std::max(2, (SB / 2) & ~1);
		int at = 0;
		sub_r0.push_back(0);
		if (ramp && n > 2 * SB + 2 * (q + h)) {
			at += q;
			sub_r0.push_back(at);
			at += h;
			sub_r0.push_back(at);
			while (n - at > SB + h + q) {
				at += SB;
				sub_r0.push_back(at);
			}
			const int rest = n - at - h - q;                        // in (0, SB]
			at += (rest + 1) & ~1;
			sub_r0.push_back(at);
			at += h;
			sub_r0.push_back(at);
		} else {
			for (at = SB; at < n; at += SB) sub_r0.push_back(at);
		}
		if (sub_r0.back() != n) sub_r0.push_back(n);
	}
	const int n_sub = (int) sub_r0.size() - 1;
	// string heap: sub-batch k owns the part of the heap that corresponds to its share of the reads (16-byte aligned)
	if (out->str_capacity > 0xFFFFFFF0ull) return fail(NGM_B200_EINVAL, "string heap beyond 32 bits");
	std::vector<size_t> heap_at((size_t) n_sub + 1);
	for (int k = 0; k <= n_sub; ++k) heap_at[k] = (size_t) ((unsigned long long) out->str_capacity * (unsigned long long) sub_r0[k] / (unsigned long long) n) & ~(size_t) 15;
	size_t slot = 0;                                                // the largest part: sizes the lanes' staging
	for (int k = 0; k < n_sub; ++k) slot = std::max(slot, heap_at[k + 1] - heap_at[k]);
	CU(B->h_used.ensure((size_t) n_sub * 4));
	uint32_t *h_used = B->h_used.as<uint32_t>();
	const size_t desc_bytes = in->desc_format == NGM_B200_DESC_U64 ? 8 : sizeof(ngm_b200_pair);
	const int n_lanes = std::min(std::min(B->n_lanes, (int) B->lanes.size()), std::max(1, n_sub));
	int64_t pe_sum = 0, pe_count = 0;
	if (in->paired && (rc = ngm_b200_pe_insert_stats(c, &pe_sum, &pe_count)) < 0) return rc;
	size_t worst_used = 0, total_used = 0, need_capacity = 0;
	bool overflow = false;
	int err = NGM_B200_OK;
	auto finish = [&](int li) -> int {                              // fetch the strings of the lane's previous sub-batch
		LaneBuf &L = B->bufs[li];
		if (L.pending < 0) return NGM_B200_OK;
		const int k = L.pending;
		L.pending = -1;
		CU(cudaEventSynchronize(L.done));
		const size_t base = heap_at[k], part = heap_at[k + 1] - heap_at[k];
		const size_t used = (size_t) h_used[k] - base;
		worst_used = std::max(worst_used, used);
		total_used += used;
		// capacity with which this sub-batch's share would have held its strings
		need_capacity = std::max(need_capacity, (size_t) ((unsigned long long) (used + 32) * (unsigned long long) n / (unsigned long long) (sub_r0[k + 1] - sub_r0[k])) + 64);
		if (used > part) {
			overflow = true;
			return NGM_B200_OK;
		}
		if (used) CU(cudaMemcpyAsync(out->strings + base, L.d_strings.p, used, cudaMemcpyDeviceToHost, B->lanes[li]->stream));
		return NGM_B200_OK;
	};
	const ngm_b200_read_exc *exc = in->exceptions;
	size_t exc_at = 0;
	for (int k = 0; k < n_sub && err == NGM_B200_OK; ++k) {
		const int li = k % n_lanes;
		ngm_b200_ctx *l = B->lanes[li];
		LaneBuf &L = B->bufs[li];
		cudaStream_t st = l->stream;
		if ((err = finish(li)) != NGM_B200_OK) break;
		const int r0 = sub_r0[k], m = sub_r0[k + 1] - r0;
		const int p0 = cb[r0], mp = cb[r0 + m] - p0;
		// ---- host -> device
		const size_t rbytes = (size_t) m * in->read_stride;
		if (cudaSuccess != L.d_in_reads.ensure(rbytes) || cudaSuccess != L.d_cb.ensure(((size_t) m + 1) * 4) ||
				cudaSuccess != L.d_desc.ensure(std::max<size_t>(mp, 1) * desc_bytes) || cudaSuccess != L.d_scores.ensure(std::max<size_t>(mp, 1) * 4) ||
				cudaSuccess != L.d_best.ensure((size_t) m * 4) || cudaSuccess != L.d_mapq.ensure((size_t) m * 4) || cudaSuccess != L.d_ntop.ensure((size_t) m * 4) ||
				cudaSuccess != L.d_pfail.ensure((size_t) m * 4) || cudaSuccess != L.d_recs.ensure((size_t) m * topn * sizeof(ngm_b200_align_rec)) ||
				cudaSuccess != L.d_strings.ensure(std::max<size_t>(slot, 16)) || cudaSuccess != L.d_cursor.ensure(4) ||
				(topn > 1 && (cudaSuccess != L.d_tsel.ensure((size_t) m * topn * 4) || cudaSuccess != L.d_tnsel.ensure((size_t) m * 4)))) {
			err = fail(NGM_B200_ECUDA, "lane staging: %s", cudaGetErrorString(cudaGetLastError()));
			break;
		}
		cudaError_t e = cudaMemcpyAsync(L.d_in_reads.p, static_cast<const char *>(in->reads) + (size_t) r0 * in->read_stride, rbytes, cudaMemcpyHostToDevice, st);
		if (e == cudaSuccess) e = cudaMemcpyAsync(L.d_cb.p, cb + r0, ((size_t) m + 1) * 4, cudaMemcpyHostToDevice, st);
		if (e == cudaSuccess && mp) e = cudaMemcpyAsync(L.d_desc.p, static_cast<const char *>(in->desc) + (size_t) (p0 - cb[0]) * desc_bytes, (size_t) mp * desc_bytes, cudaMemcpyHostToDevice, st);
		uint32_t n_exc = 0;
		if (e == cudaSuccess && in->read_format == NGM_B200_READS_PACKED2) {
			if (cudaSuccess != L.d_in_len.ensure((size_t) m * 2)) e = cudaErrorMemoryAllocation;
			if (e == cudaSuccess) e = cudaMemcpyAsync(L.d_in_len.p, in->read_len + r0, (size_t) m * 2, cudaMemcpyHostToDevice, st);
			if (exc != nullptr) {
				while (exc_at < in->n_exceptions && exc[exc_at].read_index < (uint32_t) r0) ++exc_at;
				size_t hi = exc_at;
				while (hi < in->n_exceptions && exc[hi].read_index < (uint32_t) (r0 + m)) ++hi;
				n_exc = (uint32_t) (hi - exc_at);
				if (n_exc) {
					if (cudaSuccess != L.d_in_exc.ensure((size_t) n_exc * sizeof(ngm_b200_read_exc))) e = cudaErrorMemoryAllocation;
					if (e == cudaSuccess) e = cudaMemcpyAsync(L.d_in_exc.p, exc + exc_at, (size_t) n_exc * sizeof(ngm_b200_read_exc), cudaMemcpyHostToDevice, st);
				}
			}
		}
		if (e != cudaSuccess) {
			err = fail(NGM_B200_ECUDA, "host -> device copy: %s", cudaGetErrorString(e));
			break;
		}
		// ---- kernels
		if ((err = install_reads(l, in->read_format, L.d_in_reads.p, m, in->read_stride, L.d_in_len.as<uint16_t>(), L.d_in_exc.as<ngm_b200_read_exc>(), n_exc,
				(uint32_t) r0, st)) != NGM_B200_OK)
			break;
		const uint32_t base = (uint32_t) heap_at[k], part = (uint32_t) (heap_at[k + 1] - heap_at[k]);
		batch_set_u32_kernel<<<1, 1, 0, st>>>(L.d_cursor.as<uint32_t>(), base, nullptr);
		if (p0) batch_rebase_kernel<<<(m + 1 + 255) / 256, 256, 0, st>>>(m + 1, L.d_cb.as<int>(), p0);      // offsets into the sub-batch's own arrays
		DevIn di = { m, in->mode, in->paired, in->desc_format, L.d_cb.as<int>(), 0, L.d_desc.p, mp };
		DevOut dn = { L.d_scores.as<float>(), L.d_best.as<int>(), L.d_mapq.as<int>(), L.d_ntop.as<int>(), L.d_pfail.as<int>(), L.d_recs.as<ngm_b200_align_rec>(),
				L.d_strings.as<char>() - base, (uint32_t) (base + part), L.d_cursor.as<uint32_t>() };
		if (topn > 1) {
			dn.topn = topn;
			dn.sel = L.d_tsel.as<int>();
			dn.n_sel = L.d_tnsel.as<int>();
		}
		static const bool stagger = [] { const char *e = getenv("NGM_B200_STAGGER"); return e == nullptr || atoi(e) != 0; }();
		const bool chain = in->paired || stagger;
		if ((err = batch_enqueue(l, L, di, dn, c->se_strata, st, (chain && k > 0) ? B->pe_chain : nullptr, chain ? B->pe_chain : nullptr)) != NGM_B200_OK) break;
		// candidate indices of the caller's arrays, not of the sub-batch
		if (p0 - cb[0] != 0) {
			batch_add_base_kernel<<<(m + 255) / 256, 256, 0, st>>>(m, L.d_best.as<int>(), p0 - cb[0]);
			if (topn > 1) batch_add_base_kernel<<<(unsigned) (((size_t) m * topn + 255) / 256), 256, 0, st>>>(m * topn, L.d_tsel.as<int>(), p0 - cb[0]);
		}
		l->launches += 2;
		// ---- device -> host
		e = cudaMemcpyAsync(out->best_pair + r0, L.d_best.p, (size_t) m * 4, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess) e = cudaMemcpyAsync(out->mapq + r0, L.d_mapq.p, (size_t) m * 4, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess && out->num_top) e = cudaMemcpyAsync(out->num_top + r0, L.d_ntop.p, (size_t) m * 4, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess && in->paired) e = cudaMemcpyAsync(out->pair_fail + r0, L.d_pfail.p, (size_t) m * 4, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess) e = cudaMemcpyAsync(out->recs + (size_t) r0 * topn, L.d_recs.p, (size_t) m * topn * sizeof(ngm_b200_align_rec), cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess && topn > 1) e = cudaMemcpyAsync(out->sel + (size_t) r0 * topn, L.d_tsel.p, (size_t) m * topn * 4, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess && topn > 1) e = cudaMemcpyAsync(out->n_sel + r0, L.d_tnsel.p, (size_t) m * 4, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess && out->scores && mp) e = cudaMemcpyAsync(out->scores + (p0 - cb[0]), L.d_scores.p, (size_t) mp * 4, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess) e = cudaMemcpyAsync(h_used + k, L.d_cursor.p, 4, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess) e = cudaEventRecord(L.done, st);
		if (e != cudaSuccess) {
			err = fail(NGM_B200_ECUDA, "device -> host copy: %s", cudaGetErrorString(e));
			break;
		}
		L.pending = k;
	}
	for (int li = 0; li < n_lanes; ++li) {
		const int rc2 = finish(li);
		if (err == NGM_B200_OK) err = rc2;
	}
	for (int li = 0; li < n_lanes; ++li) {
		const cudaError_t e = cudaStreamSynchronize(B->lanes[li]->stream);
		if (e != cudaSuccess && err == NGM_B200_OK) err = fail(NGM_B200_ECUDA, "lane %d: %s", li, cudaGetErrorString(e));
	}
	if (err != NGM_B200_OK) {
		if (in->paired) ngm_b200_pe_set_insert_stats(c, pe_sum, pe_count);
		return err;
	}
	out->str_used = total_used;
	if (overflow) {
		if (in->paired) ngm_b200_pe_set_insert_stats(c, pe_sum, pe_count);
		out->str_used = need_capacity;
		return fail(NGM_B200_ERANGE, "string heap too small: a sub-batch needed %zu bytes, more than its share of the heap; %zu bytes would do", worst_used, need_capacity);
	}
	return n;
}

// CS::RunBatch -> ScoreBuffer::DoRun -> AlignmentBuffer::DoRun for a whole batch, host buffers in and out (CS.cpp:340-436,
// ScoreBuffer.cpp:80-277,365-502, AlignmentBuffer.cpp:64-147).  Sub-batches rotate over the lanes in two phases: A = upload + candidate
// search (the number of candidates comes back to the host), B = score + select + align + download at the sub-batch's offset in the
// caller's candidate arrays.  A(k) is enqueued before the host waits for the count of k - 1, so the upload of k overlaps the search of k - 1.
int ngm_b200_map_batch(ngm_b200_ctx *c, const char *reads, int n_reads, int stride, int mode, int paired, ngm_b200_map_result *res) {
	if (c == nullptr || reads == nullptr || res == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	if (n_reads <= 0) return 0;
	if (res->cand_begin == nullptr || res->best_pair == nullptr || res->mapq == nullptr || res->num_top == nullptr || res->max_hit == nullptr ||
			res->recs == nullptr || (res->capacity && (res->pairs == nullptr || res->scores == nullptr)) || (res->str_capacity && res->strings == nullptr) ||
			(paired && res->pair_fail == nullptr))
		return fail(NGM_B200_EINVAL, "NULL result array");
	if (paired && (n_reads & 1)) return fail(NGM_B200_EINVAL, "paired batches hold the mates in rows 2f and 2f + 1: %d rows", n_reads);
	if (res->capacity > 0x7FFFFFFFull || res->str_capacity > 0xFFFFFFFFull) return fail(NGM_B200_EINVAL, "capacity beyond 32 bits");
	if (c->cs == nullptr) return fail(NGM_B200_ESTATE, "cs_build_index / cs_load_index must precede ngm_b200_map_batch");
	const int m0 = mode_of(mode);
	if (m0 < 0) return fail(NGM_B200_EINVAL, "unsupported alignment mode %d", mode & 0xFF);
	const int topn = paired ? 1 : c->se_topn;
	if (topn > 1 && (res->sel == nullptr || res->n_sel == nullptr)) return fail(NGM_B200_EINVAL, "topn %d needs the sel / n_sel arrays", topn);
	CU(cudaSetDevice(c->device));
	int rc = sync_lanes(c);
	if (rc) return rc;
	BatchState *B = c->batch;
	const int n = n_reads, SB = B->sub_batch;
	const int n_sub = (n + SB - 1) / SB;
	const int n_lanes = std::min(std::min(B->n_lanes, (int) B->lanes.size()), std::max(1, n_sub));
	const size_t slot = (res->str_capacity / (size_t) n_sub) & ~(size_t) 15;
	const size_t cap = res->capacity;
	const uint32_t lane_cap = (uint32_t) std::max<size_t>(cap, 1);
	CU(B->h_used.ensure((size_t) n_sub * 4));
	CU(B->h_count.ensure((size_t) n_sub * 4));
	uint32_t *h_used = B->h_used.as<uint32_t>();
	int *h_count = B->h_count.as<int>();
	int64_t pe_sum = 0, pe_count = 0;
	if (paired && (rc = ngm_b200_pe_insert_stats(c, &pe_sum, &pe_count)) < 0) return rc;
	size_t total = 0, worst_used = 0, total_used = 0;
	bool cand_overflow = false, str_overflow = false;
	int err = NGM_B200_OK;
	res->n_candidates = 0;
	res->str_used = 0;
	auto finish = [&](int li) -> int {
		LaneBuf &L = B->bufs[li];
		if (L.pending < 0) return NGM_B200_OK;
		const int k = L.pending;
		L.pending = -1;
		CU(cudaEventSynchronize(L.done));
		const size_t base = slot * (size_t) k, used = (size_t) h_used[k] - base;
		worst_used = std::max(worst_used, used);
		total_used += used;
		if (used > slot) {
			str_overflow = true;
			return NGM_B200_OK;
		}
		if (used) CU(cudaMemcpyAsync(res->strings + base, L.d_strings.p, used, cudaMemcpyDeviceToHost, B->lanes[li]->stream));
		return NGM_B200_OK;
	};
	auto phase_a = [&](int k) -> int {
		const int li = k % n_lanes;
		ngm_b200_ctx *l = B->lanes[li];
		LaneBuf &L = B->bufs[li];
		cudaStream_t st = l->stream;
		const int r0 = k * SB, m = std::min(SB, n - r0);
		CU(L.d_in_reads.ensure((size_t) m * stride));
		CU(L.d_cb.ensure(((size_t) m + 1) * 4));
		CU(L.d_desc.ensure((size_t) lane_cap * sizeof(ngm_b200_pair)));
		CU(L.d_maxhit.ensure((size_t) m * 4));
		CU(cudaMemcpyAsync(L.d_in_reads.p, reads + (size_t) r0 * stride, (size_t) m * stride, cudaMemcpyHostToDevice, st));
		int rc2 = pack_reads_device(l, L.d_in_reads.as<uint8_t>(), m, stride, st);
		if (rc2) return rc2;
		l->reads_stride = stride;
		rc2 = ngm_b200_dev_cs_search(l, L.d_in_reads.p, m, stride, 0, L.d_cb.p, L.d_desc.p, nullptr, lane_cap, L.d_maxhit.p, st);
		if (rc2 < 0) return rc2;
		CU(cudaMemcpyAsync(h_count + k, L.d_cb.as<int>() + m, 4, cudaMemcpyDeviceToHost, st));
		CU(cudaEventRecord(L.searched, st));
		return NGM_B200_OK;
	};
	auto phase_b = [&](int k) -> int {
		const int li = k % n_lanes;
		ngm_b200_ctx *l = B->lanes[li];
		LaneBuf &L = B->bufs[li];
		cudaStream_t st = l->stream;
		const int r0 = k * SB, m = std::min(SB, n - r0);
		CU(cudaEventSynchronize(L.searched));
		const size_t mp = (size_t) h_count[k], p0 = total;
		total += mp;
		if (mp > lane_cap || total > cap) {
			cand_overflow = true;
			return NGM_B200_OK;
		}
		if (cand_overflow) return NGM_B200_OK;
		CU(L.d_scores.ensure(std::max<size_t>(mp, 1) * 4));
		CU(L.d_best.ensure((size_t) m * 4));
		CU(L.d_mapq.ensure((size_t) m * 4));
		CU(L.d_ntop.ensure((size_t) m * 4));
		CU(L.d_pfail.ensure((size_t) m * 4));
		CU(L.d_recs.ensure((size_t) m * topn * sizeof(ngm_b200_align_rec)));
		CU(L.d_strings.ensure(std::max<size_t>(slot, 16)));
		CU(L.d_cursor.ensure(4));
		const uint32_t base = (uint32_t) (slot * (size_t) k);
		batch_set_u32_kernel<<<1, 1, 0, st>>>(L.d_cursor.as<uint32_t>(), base, nullptr);
		DevIn di = { m, mode, paired, NGM_B200_DESC_PAIR16, L.d_cb.as<int>(), 0, L.d_desc.p, (int) mp };
		DevOut dn = { L.d_scores.as<float>(), L.d_best.as<int>(), L.d_mapq.as<int>(), L.d_ntop.as<int>(), L.d_pfail.as<int>(), L.d_recs.as<ngm_b200_align_rec>(),
				L.d_strings.as<char>() - base, (uint32_t) (base + slot), L.d_cursor.as<uint32_t>() };
		if (topn > 1) {
			CU(L.d_tsel.ensure((size_t) m * topn * 4));
			CU(L.d_tnsel.ensure((size_t) m * 4));
			dn.topn = topn;
			dn.sel = L.d_tsel.as<int>();
			dn.n_sel = L.d_tnsel.as<int>();
		}
		int rc2 = batch_enqueue(l, L, di, dn, c->se_strata, st, (paired && k > 0) ? B->pe_chain : nullptr, paired ? B->pe_chain : nullptr);
		if (rc2) return rc2;
		if (p0) {
			batch_add_base_kernel<<<(m + 255) / 256, 256, 0, st>>>(m, L.d_best.as<int>(), (int) p0);
			if (topn > 1) batch_add_base_kernel<<<(unsigned) (((size_t) m * topn + 255) / 256), 256, 0, st>>>(m * topn, L.d_tsel.as<int>(), (int) p0);
			batch_rebase_kernel<<<(m + 1 + 255) / 256, 256, 0, st>>>(m + 1, L.d_cb.as<int>(), -(int) p0);
		}
		l->launches += 3;
		CU(cudaMemcpyAsync(res->cand_begin + r0, L.d_cb.p, (size_t) m * 4, cudaMemcpyDeviceToHost, st));
		if (mp) {
			CU(cudaMemcpyAsync(res->pairs + p0, L.d_desc.p, mp * sizeof(ngm_b200_pair), cudaMemcpyDeviceToHost, st));
			CU(cudaMemcpyAsync(res->scores + p0, L.d_scores.p, mp * 4, cudaMemcpyDeviceToHost, st));
		}
		CU(cudaMemcpyAsync(res->best_pair + r0, L.d_best.p, (size_t) m * 4, cudaMemcpyDeviceToHost, st));
		CU(cudaMemcpyAsync(res->mapq + r0, L.d_mapq.p, (size_t) m * 4, cudaMemcpyDeviceToHost, st));
		CU(cudaMemcpyAsync(res->num_top + r0, L.d_ntop.p, (size_t) m * 4, cudaMemcpyDeviceToHost, st));
		if (paired) CU(cudaMemcpyAsync(res->pair_fail + r0, L.d_pfail.p, (size_t) m * 4, cudaMemcpyDeviceToHost, st));
		CU(cudaMemcpyAsync(res->max_hit + r0, L.d_maxhit.p, (size_t) m * 4, cudaMemcpyDeviceToHost, st));
		CU(cudaMemcpyAsync(res->recs + (size_t) r0 * topn, L.d_recs.p, (size_t) m * topn * sizeof(ngm_b200_align_rec), cudaMemcpyDeviceToHost, st));
		if (topn > 1) {
			CU(cudaMemcpyAsync(res->sel + (size_t) r0 * topn, L.d_tsel.p, (size_t) m * topn * 4, cudaMemcpyDeviceToHost, st));
			CU(cudaMemcpyAsync(res->n_sel + r0, L.d_tnsel.p, (size_t) m * 4, cudaMemcpyDeviceToHost, st));
		}
		CU(cudaMemcpyAsync(h_used + k, L.d_cursor.p, 4, cudaMemcpyDeviceToHost, st));
		CU(cudaEventRecord(L.done, st));
		L.pending = k;
		return NGM_B200_OK;
	};
	for (int k = 0; k <= n_sub && err == NGM_B200_OK; ++k) {
		if (n_lanes == 1 && k >= 1) err = phase_b(k - 1);             // one lane: its previous sub-batch must be complete before the lane is reused
		if (err == NGM_B200_OK && k < n_sub) {
			err = finish(k % n_lanes);
			if (err == NGM_B200_OK) err = phase_a(k);
		}
		if (err == NGM_B200_OK && n_lanes > 1 && k >= 1) err = phase_b(k - 1);
	}
	for (int li = 0; li < n_lanes; ++li) {
		const int rc2 = finish(li);
		if (err == NGM_B200_OK) err = rc2;
	}
	for (int li = 0; li < n_lanes; ++li) {
		const cudaError_t e = cudaStreamSynchronize(B->lanes[li]->stream);
		if (e != cudaSuccess && err == NGM_B200_OK) err = fail(NGM_B200_ECUDA, "lane %d: %s", li, cudaGetErrorString(e));
	}
	res->n_candidates = total;
	if (err == NGM_B200_OK && cand_overflow) err = fail(NGM_B200_ERANGE, "candidate arrays too small: %zu entries needed", total);
	if (err == NGM_B200_OK) {
		res->cand_begin[n] = (int32_t) total;
		res->str_used = total_used;
		if (str_overflow) {
			res->str_used = (worst_used + 64) * (size_t) n_sub;
			err = fail(NGM_B200_ERANGE, "string heap too small: a sub-batch needed %zu bytes, its slot holds %zu", worst_used, slot);
		}
	}
	if (err != NGM_B200_OK) {
		if (paired) ngm_b200_pe_set_insert_stats(c, pe_sum, pe_count);      // a batch that has to be repeated starts from the same insert-size sums
		return err;
	}
	return n_reads;
}

}  // extern "C"
