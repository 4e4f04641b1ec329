// ngm_select.cu -- paired-end selection between scoring and alignment.
//
// What ScoreBuffer does once both mates of a fragment are scored (src/ScoreBuffer.cpp:196-215): top1PE (:365-462) sorts both
// candidate lists by score (std::sort -- the order it leaves EQUAL scores in decides which candidate wins, so libstdc++'s
// introsort is restated step by step), derives both mapping qualities from the two best scores, keeps the candidates within
// pair_score_cutoff of the best, and searches the best-scoring (read, mate) combination whose insert size lies inside
// (min_insert_size, max_insert_size) (CheckPairs, :464-502).  Equal pair scores are broken by the distance to the running mean
// insert size of the pairs accepted SO FAR (pairDistSum / pairDistCount, ScoreBuffer.h:31-32): a sequential dependency through the
// whole run.  Here: one thread per fragment does everything that does not need the mean; fragments that reach the tie-break are
// deferred, a prefix sum over the accepted insert sizes gives the mean in front of every fragment, and the (few) deferred ones are
// replayed in input order by one thread.  The result equals the reference run with one CS thread (`-t 1`); with several threads the
// reference's own result depends on how reads were dealt to threads.
#include <cub/cub.cuh>

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstdint>

#include "ngm_ctx.h"

namespace ngm {

struct SelItem {           // LocationScore as far as selection needs it
	float score;
	int orig;              // index into the caller's pair / score arrays
};

struct PeDev {
	float cutoff;
	int min_insert, max_insert, strata, fast_pairing, half_corridor;
};

__device__ __forceinline__ bool sel_comp(const SelItem &a, const SelItem &b) { return a.score > b.score; }      // sortLocationScore, ScoreBuffer.cpp:30-32

__device__ __forceinline__ void sel_swap(SelItem &a, SelItem &b) {
	const SelItem t = a;
	a = b;
	b = t;
}

__device__ void sel_unguarded_linear_insert(SelItem *s, int last) {
	const SelItem val = s[last];
	int next = last - 1;
	while (sel_comp(val, s[next])) {
		s[last] = s[next];
		last = next;
		--next;
	}
	s[last] = val;
}

__device__ void sel_insertion_sort(SelItem *s, int first, int last) {
	for (int i = first + 1; i < last; ++i) {
		if (sel_comp(s[i], s[first])) {
			const SelItem val = s[i];
			for (int j = i; j > first; --j) s[j] = s[j - 1];
			s[first] = val;
		} else {
			sel_unguarded_linear_insert(s, i);
		}
	}
}

__device__ void sel_adjust_heap(SelItem *s, int hole, int len, SelItem value) {      // bits/stl_heap.h __adjust_heap + __push_heap
	const int top = hole;
	int second = hole;
	while (second < (len - 1) / 2) {
		second = 2 * (second + 1);
		if (sel_comp(s[second], s[second - 1])) second--;
		s[hole] = s[second];
		hole = second;
	}
	if ((len & 1) == 0 && second == (len - 2) / 2) {
		second = 2 * (second + 1);
		s[hole] = s[second - 1];
		hole = second - 1;
	}
	int parent = (hole - 1) / 2;
	while (hole > top && sel_comp(s[parent], value)) {
		s[hole] = s[parent];
		hole = parent;
		parent = (hole - 1) / 2;
	}
	s[hole] = value;
}

__device__ void sel_heap_sort(SelItem *s, int len) {           // __partial_sort(first, last, last)
	if (len >= 2) {
		for (int parent = (len - 2) / 2;; --parent) {
			sel_adjust_heap(s, parent, len, s[parent]);
			if (parent == 0) break;
		}
	}
	for (int last = len - 1; last >= 1; --last) {
		const SelItem value = s[last];
		s[last] = s[0];
		sel_adjust_heap(s, 0, last, value);
	}
}

// std::sort(s, s + n, sortLocationScore) of libstdc++ (bits/stl_algo.h: __sort, __introsort_loop, __final_insertion_sort).
// The recursion on the right part becomes an explicit stack: the sub-ranges are disjoint, so their order does not matter.
__device__ void sel_sort(SelItem *s, int n) {
	if (n <= 1) return;
	if (n > 16) {
		int lg = 0;
		for (int v = n; v > 1; v >>= 1) ++lg;
		int st_first[64], st_last[64], st_depth[64];
		int sp = 0;
		st_first[0] = 0;
		st_last[0] = n;
		st_depth[0] = 2 * lg;
		sp = 1;
		while (sp > 0) {
			--sp;
			int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
			while (last - first > 16) {
				if (depth == 0) {
					sel_heap_sort(s + first, last - first);
					break;
				}
				--depth;
				const int mid = first + (last - first) / 2;
				{                                                   // __move_median_to_first(first, first + 1, mid, last - 1)
					SelItem &res = s[first], &a = s[first + 1], &b = s[mid], &c = s[last - 1];
					if (sel_comp(a, b)) {
						if (sel_comp(b, c)) sel_swap(res, b);
						else if (sel_comp(a, c)) sel_swap(res, c);
						else sel_swap(res, a);
					} else if (sel_comp(a, c)) sel_swap(res, a);
					else if (sel_comp(b, c)) sel_swap(res, c);
					else sel_swap(res, b);
				}
				int lo = first + 1, hi = last;                      // __unguarded_partition(first + 1, last, first)
				for (;;) {
					while (sel_comp(s[lo], s[first])) ++lo;
					--hi;
					while (sel_comp(s[first], s[hi])) --hi;
					if (!(lo < hi)) break;
					sel_swap(s[lo], s[hi]);
					++lo;
				}
				if (sp < 64) {
					st_first[sp] = lo;
					st_last[sp] = last;
					st_depth[sp] = depth;
					++sp;
				}
				last = lo;
			}
		}
		sel_insertion_sort(s, 0, 16);
		for (int i = 16; i < n; ++i) sel_unguarded_linear_insert(s, i);
	} else {
		sel_insertion_sort(s, 0, n);
	}
}

__device__ __forceinline__ int sel_mq(float best, float second) {       // ScoreBuffer::computeMQ, ScoreBuffer.cpp:34-40
	int mq = 0;
	if (best > 0.0f && second >= 0.0f) mq = (int) ceilf(60.0f * (best - second) / best);
	return mq;
}

struct SelOut {
	int best, mapq, num_top;
};

__device__ void sel_top1_se(const SelItem *s, int n, int strata, SelOut &o) {    // top1SE, ScoreBuffer.cpp:228-277
	float best = 0.0f, second = 0.0f;
	int besti = 0, nbest = 0;
	for (int j = 0; j < n; ++j) {
		const float v = s[j].score;
		if (v > second) {
			if (v > best) {
				second = best;
				best = v;
				besti = j;
				nbest = 1;
			} else if (v == best) {
				++nbest;
				second = best;
			} else {
				second = v;
			}
		} else if (v == best) {
			++nbest;
		}
	}
	o.mapq = sel_mq(best, second);
	if (nbest == 1 || !strata) {
		o.best = n > 0 ? s[besti].orig : -1;
		o.num_top = nbest;
	} else {
		o.best = -1;
		o.mapq = 0;
		o.num_top = 1;
	}
}

enum : int { kFragNone = 0, kFragPair = 1, kFragDeferred = 2 };

// top1PE for fragment f (mates a = 2f, b = 2f + 1; the reference calls it for b, the mate whose scores complete last, with mate = a).
// The lists must be sorted already.  Returns false when the mean insert size is needed and not known.
__device__ bool sel_top1_pe(const SelItem *sb, int nb, int len_b, const SelItem *sa, int na, int len_a, const ngm_b200_pair *__restrict__ pairs,
		const PeDev &P, bool have_avg, int avg, SelOut &ob, SelOut &oa, int &paired_fail, int &distance) {
	ob.mapq = nb > 1 ? sel_mq(sb[0].score, sb[1].score) : 60;      // computeMQ(MappedRead *), :42-49
	oa.mapq = na > 1 ? sel_mq(sa[0].score, sa[1].score) : 60;
	const float min_b = __fmul_rn(sb[0].score, P.cutoff);
	int nr = 1;
	while (nr < nb && min_b <= sb[nr].score) ++nr;
	const float min_a = __fmul_rn(sa[0].score, P.cutoff);
	int nm = 1;
	while (nm < na && min_a <= sa[nm].score) ++nm;
	float top = 0.0f;
	int equal = 0, t1 = -1, t2 = -1;
	distance = 0;
	for (int i = 0; i < nr; ++i) {
		const uint64_t l1 = pairs[sb[i].orig].window_start + (uint64_t) P.half_corridor;
		for (int j = 0; j < nm; ++j) {                              // CheckPairs(&read->Scores[i], read->length, &mate->Scores[j], mate->length, ...)
			const uint64_t l2 = pairs[sa[j].orig].window_start + (uint64_t) P.half_corridor;
			const int cur = (l2 > l1) ? (int) (l2 - l1 + (uint64_t) (int64_t) len_a) : (int) (l1 - l2 + (uint64_t) (int64_t) len_b);
			if (cur > P.min_insert && cur < P.max_insert) {
				const float ps = __fadd_rn(sb[i].score, sa[j].score);
				if (ps > top) {
					top = ps;
					distance = cur;
					t1 = i;
					t2 = j;
				} else if (ps == top) {
					if (!have_avg) return false;
					if (abs(distance - avg) > abs(cur - avg)) {
						top = ps;
						distance = cur;
						t1 = i;
						t2 = j;
					} else if (abs(distance) == abs(cur)) {
						equal += 1;
					}
				}
			}
		}
	}
	paired_fail = 0;
	if (top > 0.0f) {
		if (equal <= 0 || !P.strata) {
			ob.best = sb[t1].orig;
			oa.best = sa[t2].orig;
			ob.num_top = oa.num_top = equal;
		} else {
			ob.best = oa.best = -1;
			ob.mapq = oa.mapq = 0;
			ob.num_top = oa.num_top = 1;
			distance = 0;
		}
	} else {                                                        // no proper pair: single-end selection (on the sorted lists)
		sel_top1_se(sb, nb, P.strata, ob);
		sel_top1_se(sa, na, P.strata, oa);
		paired_fail = 1;
		distance = 0;
	}
	return true;
}

__device__ __forceinline__ void sel_store(int r, const SelOut &o, int pf, int *best_pair, int *mapq, int *num_top, int *pflags) {
	best_pair[r] = o.best;
	mapq[r] = o.mapq;
	num_top[r] = o.num_top;
	pflags[r] = pf;
}

__global__ void select_pairs_kernel(int n_frag, const int *__restrict__ cand_begin, const ngm_b200_pair *__restrict__ pairs,
		const float *__restrict__ scores, const uint16_t *__restrict__ rlen, const PeDev P, SelItem *__restrict__ items,
		int *__restrict__ best_pair, int *__restrict__ mapq, int *__restrict__ num_top, int *__restrict__ pflags,
		longlong2 *__restrict__ contrib, uint8_t *__restrict__ deferred) {
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= n_frag) return;
	const int a = 2 * f, b = a + 1;
	const int ba = cand_begin[a], bb = cand_begin[b], eb = cand_begin[b + 1];
	const int na = bb - ba, nb = eb - bb;
	for (int j = ba; j < eb; ++j) {
		items[j].score = scores[j];
		items[j].orig = j;
	}
	SelOut oa = {-1, 0, 1}, ob = {-1, 0, 1};                        // MappedRead ctor: numTopScores 1
	int pf = 0, distance = 0, state = kFragNone;
	if (na > 0 && nb > 0 && !P.fast_pairing) {
		sel_sort(items + ba, na);
		sel_sort(items + bb, nb);
		if (sel_top1_pe(items + bb, nb, (int) rlen[b], items + ba, na, (int) rlen[a], pairs, P, false, 0, ob, oa, pf, distance)) {
			state = (!pf && ob.best >= 0) ? kFragPair : kFragNone;
		} else {
			state = kFragDeferred;
		}
	} else {
		if (na > 0) sel_top1_se(items + ba, na, P.strata, oa);
		if (nb > 0) sel_top1_se(items + bb, nb, P.strata, ob);
	}
	contrib[f] = state == kFragPair ? make_longlong2(distance, 1) : make_longlong2(0, 0);
	deferred[f] = state == kFragDeferred;
	if (state != kFragDeferred) {
		sel_store(a, oa, pf, best_pair, mapq, num_top, pflags);
		sel_store(b, ob, pf, best_pair, mapq, num_top, pflags);
	}
}

// ScoreBuffer::topNSE (ScoreBuffer.cpp:279-330), topn > 1; one thread per read.  sel[r * topn + j] = the candidates handed to alignment in
// the order of the sorted list (std::sort's order of equal scores included), -1 beyond n_sel[r].
__global__ void select_topn_kernel(int n_reads, const int *__restrict__ cand_begin, const float *__restrict__ scores, int topn, int strata,
		SelItem *__restrict__ items, int *__restrict__ sel, int *__restrict__ n_sel, int *__restrict__ mapq, int *__restrict__ num_top) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int b = cand_begin[r], e = cand_begin[r + 1], n = e - b;
	for (int j = 0; j < topn; ++j) sel[(size_t) r * topn + j] = -1;
	int ns = 0, mq = 0, nt = 1;                                 // MappedRead ctor: numTopScores 1
	if (n > 0) {
		SelItem *s = items + b;
		for (int j = 0; j < n; ++j) {
			s[j].score = scores[b + j];
			s[j].orig = b + j;
		}
		sel_sort(s, n);
		nt = 1;
		while (nt < n && s[0].score == s[nt].score) ++nt;
		if (nt <= topn || !strata) {
			ns = strata ? nt : min(n, topn);
			mq = n > 1 ? sel_mq(s[0].score, s[1].score) : 60;
			for (int j = 0; j < ns; ++j) sel[(size_t) r * topn + j] = s[j].orig;
		}
	}
	n_sel[r] = ns;
	mapq[r] = mq;
	num_top[r] = nt;
}

struct Sum2 {
	__device__ __forceinline__ longlong2 operator()(const longlong2 &x, const longlong2 &y) const { return make_longlong2(x.x + y.x, x.y + y.y); }
};

// One relaxation step over the deferred fragments, all at once: every deferred fragment is decided with the mean that the CURRENT
// contributions of all earlier fragments give (prefix = inclusive scan of contrib, which holds the deferred fragments' previous
// decisions, (0, 0) at first).  The dependency is strictly from earlier to later fragments, so a step that changes nothing has reached
// the sequential result; the mean moves by a fraction of a base per pair, so that is the case after two or three steps.
__global__ void relax_deferred_kernel(const int *__restrict__ list, const int *__restrict__ n_list, const int *__restrict__ cand_begin,
		const ngm_b200_pair *__restrict__ pairs, const uint16_t *__restrict__ rlen, const PeDev P, const SelItem *__restrict__ items,
		const longlong2 *__restrict__ prefix, longlong2 *__restrict__ contrib, int *__restrict__ best_pair, int *__restrict__ mapq,
		int *__restrict__ num_top, int *__restrict__ pflags, const long long *__restrict__ state, int *__restrict__ changed) {
	const int d = blockIdx.x * blockDim.x + threadIdx.x;
	if (d >= *n_list) return;
	const int f = list[d];
	const int a = 2 * f, b = a + 1;
	const int ba = cand_begin[a], bb = cand_begin[b], eb = cand_begin[b + 1];
	const longlong2 incl = prefix[f], own = contrib[f];
	const long long sum = state[0] + incl.x - own.x, cnt = state[1] + incl.y - own.y;
	SelOut oa, ob;
	int pf = 0, distance = 0;
	sel_top1_pe(items + bb, eb - bb, (int) rlen[b], items + ba, bb - ba, (int) rlen[a], pairs, P, true, (int) (sum / cnt), ob, oa, pf, distance);
	const longlong2 now = (!pf && ob.best >= 0) ? make_longlong2(distance, 1) : make_longlong2(0, 0);
	if (now.x != own.x || now.y != own.y) {
		contrib[f] = now;
		*changed = 1;
	}
	sel_store(a, oa, pf, best_pair, mapq, num_top, pflags);
	sel_store(b, ob, pf, best_pair, mapq, num_top, pflags);
}

// pairDistSum / pairDistCount after the batch, when the last relaxation step changed nothing (prefix is then exact).
__global__ void commit_state_kernel(int n_frag, const longlong2 *__restrict__ prefix, const int *__restrict__ changed, long long *__restrict__ state) {
	if (blockIdx.x != 0 || threadIdx.x != 0 || *changed || n_frag <= 0) return;
	const longlong2 all = prefix[n_frag - 1];
	state[0] += all.x;
	state[1] += all.y;
}

// Safety net, runs only if the relaxation has not settled (`changed` still set): the deferred fragments in input order, one after the
// other.  prefix0[f] = inclusive sums of the contributions of the fragments decided in parallel (deferred ones count (0, 0) there).
// state[0..1] = pairDistSum, pairDistCount carried between batches.
__global__ void resolve_deferred_kernel(int n_frag, const int *__restrict__ list, const int *__restrict__ n_list, const int *__restrict__ cand_begin,
		const ngm_b200_pair *__restrict__ pairs, const uint16_t *__restrict__ rlen, const PeDev P, const SelItem *__restrict__ items,
		const longlong2 *__restrict__ prefix, int *__restrict__ best_pair, int *__restrict__ mapq, int *__restrict__ num_top,
		int *__restrict__ pflags, long long *__restrict__ state, const int *__restrict__ changed) {
	if (blockIdx.x != 0 || threadIdx.x != 0 || !*changed) return;
	long long extra_sum = 0, extra_cnt = 0;
	const int n = *n_list;
	for (int d = 0; d < n; ++d) {
		const int f = list[d];
		const int a = 2 * f, b = a + 1;
		const int ba = cand_begin[a], bb = cand_begin[b], eb = cand_begin[b + 1];
		const longlong2 before = prefix[f];                         // a deferred fragment contributes (0, 0) to the scan
		const long long sum = state[0] + before.x + extra_sum, cnt = state[1] + before.y + extra_cnt;
		SelOut oa, ob;
		int pf = 0, distance = 0;
		sel_top1_pe(items + bb, eb - bb, (int) rlen[b], items + ba, bb - ba, (int) rlen[a], pairs, P, true, (int) (sum / cnt), ob, oa, pf, distance);
		if (!pf && ob.best >= 0) {
			extra_sum += distance;
			extra_cnt += 1;
		}
		sel_store(a, oa, pf, best_pair, mapq, num_top, pflags);
		sel_store(b, ob, pf, best_pair, mapq, num_top, pflags);
	}
	if (n_frag > 0) {
		const longlong2 all = prefix[n_frag - 1];
		state[0] += all.x + extra_sum;
		state[1] += all.y + extra_cnt;
	}
}

struct PeState {
	ngm_b200_pe_params hp;
	DevBuf d_items, d_contrib, d_prefix, d_prefix0, d_deferred, d_list, d_nlist, d_state, d_tmp, d_changed;
	bool have_state = false;
};

void pe_release(PeState *pe) {
	if (pe == nullptr) return;
	pe->d_items.release();
	pe->d_contrib.release();
	pe->d_prefix.release();
	pe->d_prefix0.release();
	pe->d_changed.release();
	pe->d_deferred.release();
	pe->d_list.release();
	pe->d_nlist.release();
	pe->d_state.release();
	pe->d_tmp.release();
	delete pe;
}

int pe_share_state(ngm_b200_ctx *lane, const ngm_b200_ctx *root) {
	const PeState *r = root->pe;
	if (r == nullptr || !r->have_state) {
		if (lane->pe) lane->pe->have_state = false;
		return NGM_B200_OK;
	}
	if (lane->pe == nullptr) lane->pe = new PeState();
	lane->pe->hp = r->hp;
	lane->pe->d_state.borrow(r->d_state);                          // the running insert-size sums are the root's: one sequence over all lanes
	lane->pe->have_state = true;
	return NGM_B200_OK;
}

}  // namespace ngm

using namespace ngm;

int ngm_b200_pe_configure(ngm_b200_ctx *c, const ngm_b200_pe_params *params) {
	if (c == nullptr || params == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	if (!(params->pair_score_cutoff >= 0.0f)) return fail(NGM_B200_EINVAL, "pair_score_cutoff %g", params->pair_score_cutoff);
	CU(cudaSetDevice(c->device));
	if (c->pe == nullptr) c->pe = new PeState();
	c->pe->hp = *params;
	if (c->pe->hp.max_insert_size <= 0) c->pe->hp.max_insert_size = INT_MAX;      // NGM.cpp:40-41
	CU(c->pe->d_state.ensure(16));
	const long long init[2] = {0, 1};                             // ScoreBuffer.h:90: pairDistCount(1), pairDistSum(0)
	CU(cudaMemcpy(c->pe->d_state.p, init, sizeof(init), cudaMemcpyHostToDevice));
	c->pe->have_state = true;
	c->epoch += 1;
	return NGM_B200_OK;
}

int ngm_b200_pe_insert_stats(ngm_b200_ctx *c, int64_t *dist_sum, int64_t *dist_count) {
	if (c == nullptr || c->pe == nullptr || !c->pe->have_state) return fail(NGM_B200_ESTATE, "ngm_b200_pe_configure must come first");
	CU(cudaSetDevice(c->device));
	CU(cudaDeviceSynchronize());
	long long st[2];
	CU(cudaMemcpy(st, c->pe->d_state.p, sizeof(st), cudaMemcpyDeviceToHost));
	if (dist_sum) *dist_sum = st[0];
	if (dist_count) *dist_count = st[1];
	return NGM_B200_OK;
}

int ngm_b200_dev_select_topn(ngm_b200_ctx *c, int n_reads, const void *d_cand_begin, const void *d_scores, uint32_t n_pairs, int topn, int strata,
		void *d_sel, void *d_n_sel, void *d_mapq, void *d_num_top, void *stream) {
	if (c == nullptr || d_cand_begin == nullptr || d_scores == nullptr || d_sel == nullptr || d_n_sel == nullptr || d_mapq == nullptr || d_num_top == nullptr)
		return fail(NGM_B200_EINVAL, "NULL argument");
	if (topn < 1 || topn > 1000) return fail(NGM_B200_EINVAL, "topn %d not in [1, 1000]", topn);      // GenericReadWriter.h:79 MAX_PASSED
	if (n_reads <= 0) return 0;
	CU(cudaSetDevice(c->device));
	if (c->pe == nullptr) c->pe = new PeState();
	CU(c->pe->d_items.ensure(std::max<size_t>(n_pairs, 1) * sizeof(SelItem)));
	select_topn_kernel<<<(n_reads + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(n_reads, static_cast<const int *>(d_cand_begin),
			static_cast<const float *>(d_scores), topn, strata, c->pe->d_items.as<SelItem>(), static_cast<int *>(d_sel), static_cast<int *>(d_n_sel),
			static_cast<int *>(d_mapq), static_cast<int *>(d_num_top));
	c->launches += 1;
	CU(cudaGetLastError());
	return n_reads;
}

int ngm_b200_pe_set_insert_stats(ngm_b200_ctx *c, int64_t dist_sum, int64_t dist_count) {
	if (c == nullptr || c->pe == nullptr || !c->pe->have_state) return fail(NGM_B200_ESTATE, "ngm_b200_pe_configure must come first");
	if (dist_count <= 0) return fail(NGM_B200_EINVAL, "pairDistCount starts at 1");
	CU(cudaSetDevice(c->device));
	CU(cudaDeviceSynchronize());
	const long long st[2] = { (long long) dist_sum, (long long) dist_count };
	CU(cudaMemcpy(c->pe->d_state.p, st, sizeof(st), cudaMemcpyHostToDevice));
	return NGM_B200_OK;
}

int64_t ngm_b200_pe_deferred_fragments(ngm_b200_ctx *c) {
	if (c == nullptr || c->pe == nullptr || c->pe->d_nlist.p == nullptr) return fail(NGM_B200_ESTATE, "no paired selection yet");
	CU(cudaSetDevice(c->device));
	CU(cudaDeviceSynchronize());
	int n = 0;
	CU(cudaMemcpy(&n, c->pe->d_nlist.p, 4, cudaMemcpyDeviceToHost));
	return n;
}

int ngm_b200_dev_select_pairs(ngm_b200_ctx *c, int n_reads, const void *d_cand_begin, const void *d_pairs, const void *d_scores, uint32_t n_pairs,
		void *d_best_pair, void *d_mapq, void *d_num_top, void *d_pair_fail, void *stream) {
	if (c == nullptr || d_cand_begin == nullptr || d_pairs == nullptr || d_scores == nullptr || d_best_pair == nullptr || d_mapq == nullptr ||
			d_num_top == nullptr || d_pair_fail == nullptr)
		return fail(NGM_B200_EINVAL, "NULL argument");
	if (c->pe == nullptr || !c->pe->have_state) return fail(NGM_B200_ESTATE, "ngm_b200_pe_configure must come first");
	if (n_reads <= 0) return 0;
	if (n_reads & 1) return fail(NGM_B200_EINVAL, "paired selection needs an even number of reads (mates are rows 2f and 2f + 1), got %d", n_reads);
	if (c->n_reads < n_reads || c->d_rrlen.p == nullptr) return fail(NGM_B200_ESTATE, "set_reads must hold the batch whose candidates are selected");
	CU(cudaSetDevice(c->device));
	PeState *pe = c->pe;
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	const int n_frag = n_reads / 2;
	CU(pe->d_items.ensure(std::max<size_t>(n_pairs, 1) * sizeof(SelItem)));
	CU(pe->d_contrib.ensure((size_t) n_frag * sizeof(longlong2)));
	CU(pe->d_prefix.ensure((size_t) n_frag * sizeof(longlong2)));
	CU(pe->d_prefix0.ensure((size_t) n_frag * sizeof(longlong2)));
	CU(pe->d_changed.ensure(4));
	CU(pe->d_deferred.ensure((size_t) n_frag));
	CU(pe->d_list.ensure((size_t) n_frag * 4));
	CU(pe->d_nlist.ensure(4));
	PeDev P;
	P.cutoff = pe->hp.pair_score_cutoff;
	P.min_insert = pe->hp.min_insert_size;
	P.max_insert = pe->hp.max_insert_size;
	P.strata = pe->hp.strata;
	P.fast_pairing = pe->hp.fast_pairing;
	P.half_corridor = c->dp.corridor >> 1;
	const int *begin = static_cast<const int *>(d_cand_begin);
	const ngm_b200_pair *pairs = static_cast<const ngm_b200_pair *>(d_pairs);
	select_pairs_kernel<<<(n_frag + 127) / 128, 128, 0, st>>>(n_frag, begin, pairs, static_cast<const float *>(d_scores), c->d_rrlen.as<uint16_t>(), P,
			pe->d_items.as<SelItem>(), static_cast<int *>(d_best_pair), static_cast<int *>(d_mapq), static_cast<int *>(d_num_top),
			static_cast<int *>(d_pair_fail), pe->d_contrib.as<longlong2>(), pe->d_deferred.as<uint8_t>());
	CU(cudaGetLastError());
	size_t t1 = 0, t2 = 0;
	cub::CountingInputIterator<int> ids(0);
	CU(cub::DeviceScan::InclusiveScan(nullptr, t1, pe->d_contrib.as<longlong2>(), pe->d_prefix.as<longlong2>(), Sum2(), n_frag, st));
	CU(cub::DeviceSelect::Flagged(nullptr, t2, ids, pe->d_deferred.as<uint8_t>(), pe->d_list.as<int>(), pe->d_nlist.as<int>(), n_frag, st));
	CU(pe->d_tmp.ensure(std::max(t1, t2)));
	CU(cub::DeviceScan::InclusiveScan(pe->d_tmp.p, t1, pe->d_contrib.as<longlong2>(), pe->d_prefix0.as<longlong2>(), Sum2(), n_frag, st));
	CU(cub::DeviceSelect::Flagged(pe->d_tmp.p, t2, ids, pe->d_deferred.as<uint8_t>(), pe->d_list.as<int>(), pe->d_nlist.as<int>(), n_frag, st));
	int n_deferred_bound = n_frag;                              // the count stays on the device: size the grid for the worst case
	int kRelaxSteps = 4;
	if (const char *e = getenv("NGM_B200_PE_RELAX_STEPS")) kRelaxSteps = std::max(0, std::min(16, atoi(e)));      // testing hook: 0 = sequential replay only
	const longlong2 *last_prefix = pe->d_prefix0.as<longlong2>();
	CU(cudaMemsetAsync(pe->d_changed.p, kRelaxSteps == 0 ? 1 : 0, 4, st));
	for (int it = 0; it < kRelaxSteps; ++it) {
		const longlong2 *prefix = it == 0 ? pe->d_prefix0.as<longlong2>() : pe->d_prefix.as<longlong2>();
		if (it > 0) CU(cub::DeviceScan::InclusiveScan(pe->d_tmp.p, t1, pe->d_contrib.as<longlong2>(), pe->d_prefix.as<longlong2>(), Sum2(), n_frag, st));
		last_prefix = prefix;
		CU(cudaMemsetAsync(pe->d_changed.p, 0, 4, st));
		relax_deferred_kernel<<<(n_deferred_bound + 127) / 128, 128, 0, st>>>(pe->d_list.as<int>(), pe->d_nlist.as<int>(), begin, pairs,
				c->d_rrlen.as<uint16_t>(), P, pe->d_items.as<SelItem>(), prefix, pe->d_contrib.as<longlong2>(), static_cast<int *>(d_best_pair),
				static_cast<int *>(d_mapq), static_cast<int *>(d_num_top), static_cast<int *>(d_pair_fail), pe->d_state.as<long long>(),
				pe->d_changed.as<int>());
	}
	// the last step changed nothing: its prefix is exact; otherwise replay the deferred fragments one by one (contributions of the
	// fragments decided in parallel: prefix0)
	resolve_deferred_kernel<<<1, 32, 0, st>>>(n_frag, pe->d_list.as<int>(), pe->d_nlist.as<int>(), begin, pairs, c->d_rrlen.as<uint16_t>(), P,
			pe->d_items.as<SelItem>(), pe->d_prefix0.as<longlong2>(), static_cast<int *>(d_best_pair), static_cast<int *>(d_mapq),
			static_cast<int *>(d_num_top), static_cast<int *>(d_pair_fail), pe->d_state.as<long long>(), pe->d_changed.as<int>());
	commit_state_kernel<<<1, 32, 0, st>>>(n_frag, last_prefix, pe->d_changed.as<int>(), pe->d_state.as<long long>());
	c->launches += 3 + 2 * kRelaxSteps + 2;
	CU(cudaGetLastError());
	return n_reads;
}
