/* oracle/select_oracle.c -- TEST INFRASTRUCTURE ONLY (see select_oracle.h).
 *
 * Plain-C restatement of NextGenMap's candidate selection after scoring: ScoreBuffer::top1SE, computeMQ, top1PE and
 * CheckPairs (src/ScoreBuffer.cpp:34-48,228-277,365-502), including the order std::sort leaves equal scores in
 * (libstdc++ introsort as shipped with the gcc that builds oracle/_ref/ngm/ngm_ref: bits/stl_algo.h __sort,
 * __introsort_loop, __move_median_to_first, __unguarded_partition, __final_insertion_sort; __partial_sort's heap path
 * from bits/stl_heap.h) and the running insert-size average that breaks ties between equally scoring pairs.
 */
#include "select_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef sel_oracle_cand cand;

static inline int comp(const cand *a, const cand *b) {          /* sortLocationScore, ScoreBuffer.cpp:30-32 */
	return a->score > b->score;
}

static inline void swap_c(cand *a, cand *b) {
	cand t = *a;
	*a = *b;
	*b = t;
}

/* ---- libstdc++ std::sort ---------------------------------------------------------------------------------- */
static void unguarded_linear_insert(cand *last) {
	cand val = *last;
	cand *next = last - 1;
	while (comp(&val, next)) {
		*last = *next;
		last = next;
		--next;
	}
	*last = val;
}

static void insertion_sort(cand *first, cand *last) {
	if (first == last) return;
	for (cand *i = first + 1; i != last; ++i) {
		if (comp(i, first)) {
			cand val = *i;
			memmove(first + 1, first, (size_t) (i - first) * sizeof(cand));
			*first = val;
		} else {
			unguarded_linear_insert(i);
		}
	}
}

static void move_median_to_first(cand *result, cand *a, cand *b, cand *c) {
	if (comp(a, b)) {
		if (comp(b, c)) swap_c(result, b);
		else if (comp(a, c)) swap_c(result, c);
		else swap_c(result, a);
	} else if (comp(a, c)) swap_c(result, a);
	else if (comp(b, c)) swap_c(result, c);
	else swap_c(result, b);
}

static cand *unguarded_partition(cand *first, cand *last, cand *pivot) {
	for (;;) {
		while (comp(first, pivot)) ++first;
		--last;
		while (comp(pivot, last)) --last;
		if (!(first < last)) return first;
		swap_c(first, last);
		++first;
	}
}

/* bits/stl_heap.h */
static void push_heap_(cand *first, long hole, long top, cand value) {
	long parent = (hole - 1) / 2;
	while (hole > top && comp(first + parent, &value)) {
		first[hole] = first[parent];
		hole = parent;
		parent = (hole - 1) / 2;
	}
	first[hole] = value;
}

static void adjust_heap(cand *first, long hole, long len, cand value) {
	const long top = hole;
	long second = hole;
	while (second < (len - 1) / 2) {
		second = 2 * (second + 1);
		if (comp(first + second, first + (second - 1))) second--;
		first[hole] = first[second];
		hole = second;
	}
	if ((len & 1) == 0 && second == (len - 2) / 2) {
		second = 2 * (second + 1);
		first[hole] = first[second - 1];
		hole = second - 1;
	}
	push_heap_(first, hole, top, value);
}

static void heap_sort_all(cand *first, cand *last) {           /* __partial_sort(first, last, last): __heap_select is make_heap only */
	const long len = last - first;
	if (len >= 2) {
		long parent = (len - 2) / 2;
		for (;;) {
			cand value = first[parent];
			adjust_heap(first, parent, len, value);
			if (parent == 0) break;
			parent--;
		}
	}
	while (last - first > 1) {                                  /* __sort_heap -> __pop_heap(first, last, last) */
		--last;
		cand value = *last;
		*last = *first;
		adjust_heap(first, 0, last - first, value);
	}
}

static void introsort_loop(cand *first, cand *last, long depth_limit) {
	while (last - first > 16) {
		if (depth_limit == 0) {
			heap_sort_all(first, last);
			return;
		}
		--depth_limit;
		cand *mid = first + (last - first) / 2;
		move_median_to_first(first, first + 1, mid, last - 1);
		cand *cut = unguarded_partition(first + 1, last, first);
		introsort_loop(cut, last, depth_limit);
		last = cut;
	}
}

void sel_oracle_sort(cand *first, int n) {
	if (n <= 0) return;
	cand *last = first + n;
	long lg = 0;
	for (long v = n; v > 1; v >>= 1) ++lg;
	introsort_loop(first, last, lg * 2);
	if (n > 16) {
		insertion_sort(first, first + 16);
		for (cand *i = first + 16; i != last; ++i) unguarded_linear_insert(i);
	} else {
		insertion_sort(first, last);
	}
}

/* ---- selection -------------------------------------------------------------------------------------------- */
static int mq_of(float best, float second) {                   /* ScoreBuffer::computeMQ(float, float), :34-40 */
	int mq = 0;
	if (best > 0 && second >= 0) mq = (int) ceil(60.0f * (best - second) / best);
	return mq;
}

static int mq_sorted(const cand *s, int n) {                    /* ScoreBuffer::computeMQ(MappedRead *), :42-49 */
	int mq = 60;
	if (n > 1) mq = mq_of(s[0].score, s[1].score);
	return mq;
}

void sel_oracle_top1_se(const cand *s, int n, const sel_oracle_params *p, sel_oracle_result *out) {     /* top1SE, :228-277 */
	float best = 0.0f, second = 0.0f;
	int best_i = 0, n_best = 0;
	for (int j = 0; j < n; ++j) {
		if (s[j].score > second) {
			if (s[j].score > best) {
				second = best;
				best = s[j].score;
				best_i = j;
				n_best = 1;
			} else if (s[j].score == best) {
				++n_best;
				second = best;
			} else {
				second = s[j].score;
			}
		} else if (s[j].score == best) {
			++n_best;
		}
	}
	out->mapq = mq_of(best, second);
	if (n_best == 1 || !p->strata) {
		out->best = n > 0 ? s[best_i].orig : -1;                /* the reference asserts hasCandidates() here */
		out->num_top = n_best;
	} else {
		out->best = -1;
		out->mapq = 0;
	}
}

static int check_pairs(const cand *ls1, int len1, const cand *ls2, int len2, float *top, int *insert, int *equal, const sel_oracle_params *p,
		const sel_oracle_state *st) {                              /* CheckPairs, :464-502 */
	const int cur = (ls2->location > ls1->location) ? (int) (ls2->location - ls1->location + (uint64_t) (int64_t) len2)
	                                                  : (int) (ls1->location - ls2->location + (uint64_t) (int64_t) len1);
	if (cur > p->min_insert && cur < p->max_insert) {
		const float pair_score = ls1->score + ls2->score;
		if (pair_score > *top * 1.00f) {
			*top = pair_score;
			*insert = cur;
			return 1;
		} else if (pair_score == *top) {
			const int avg = (int) (st->dist_sum / st->dist_count);
			if (abs(*insert - avg) > abs(cur - avg)) {
				*top = pair_score;
				*insert = cur;
				return 1;
			} else if (abs(*insert) == abs(cur)) {
				*equal += 1;
			}
		}
	}
	return 0;
}

/* One fragment whose mates both have candidates: top1PE (:365-462) called for `read` (the mate whose scores complete last -- the second
 * mate, odd ReadId) with mate = read->Paired.  s_read / s_mate are sorted in place. */
void sel_oracle_top1_pe(cand *s_read, int n_read, int len_read, cand *s_mate, int n_mate, int len_mate, const sel_oracle_params *p,
		sel_oracle_state *st, sel_oracle_result *r_read, sel_oracle_result *r_mate) {
	sel_oracle_sort(s_read, n_read);
	sel_oracle_sort(s_mate, n_mate);
	r_read->mapq = mq_sorted(s_read, n_read);
	r_mate->mapq = mq_sorted(s_mate, n_mate);
	r_read->paired_fail = r_mate->paired_fail = 0;
	const float min_read = s_read[0].score * p->pair_score_cutoff;
	int nr = 1;
	while (nr < n_read && min_read <= s_read[nr].score) nr += 1;
	const float min_mate = s_mate[0].score * p->pair_score_cutoff;
	int nm = 1;
	while (nm < n_mate && min_mate <= s_mate[nm].score) nm += 1;
	float top = 0.0f;
	int distance = 0, equal = 0, t1 = -1, t2 = -1;
	for (int i = 0; i < nr; ++i) {
		for (int j = 0; j < nm; ++j) {
			if (check_pairs(&s_read[i], len_read, &s_mate[j], len_mate, &top, &distance, &equal, p, st)) {
				t1 = i;
				t2 = j;
			}
		}
	}
	if (top > 0.0f) {
		if (equal <= 0 || !p->strata) {
			st->dist_sum += distance;
			st->dist_count += 1;
			r_read->num_top = r_mate->num_top = equal;
			r_read->best = s_read[t1].orig;
			r_mate->best = s_mate[t2].orig;
			r_read->insert = r_mate->insert = distance;
		} else {
			r_read->mapq = r_mate->mapq = 0;
			r_read->best = r_mate->best = -1;
		}
	} else {                                                    /* no proper pair: single-end selection on the SORTED lists */
		sel_oracle_top1_se(s_read, n_read, p, r_read);
		sel_oracle_top1_se(s_mate, n_mate, p, r_mate);
		r_read->paired_fail = r_mate->paired_fail = 1;
	}
}

/* A whole batch in input order, as one ScoreBuffer sees it with `-t 1` (ScoreBuffer::DoRun, :196-215).  Reads 2f and 2f+1 are the mates
 * of fragment f.  cands[cand_begin[r] .. cand_begin[r+1]) = read r's scored candidates in CS order (orig is overwritten). */
void sel_oracle_select_pairs(int n_reads, const int *cand_begin, cand *cands, const int *read_len, const sel_oracle_params *p,
		sel_oracle_state *st, sel_oracle_result *out) {
	for (int r = 0; r < n_reads; ++r) {
		for (int j = cand_begin[r]; j < cand_begin[r + 1]; ++j) cands[j].orig = j;
		out[r].best = -1;
		out[r].mapq = 0;
		out[r].num_top = 1;                                     /* MappedRead ctor */
		out[r].paired_fail = 0;
		out[r].insert = 0;
	}
	for (int f = 0; f + 1 < n_reads; f += 2) {
		const int a = f, b = f + 1;                             /* a: first mate, b: second mate */
		const int na = cand_begin[a + 1] - cand_begin[a], nb = cand_begin[b + 1] - cand_begin[b];
		cand *sa = cands + cand_begin[a], *sb = cands + cand_begin[b];
		if (na > 0 && nb > 0) {
			if (!p->fast_pairing) {
				sel_oracle_top1_pe(sb, nb, read_len[b], sa, na, read_len[a], p, st, &out[b], &out[a]);
			} else {
				sel_oracle_top1_se(sb, nb, p, &out[b]);
				sel_oracle_top1_se(sa, na, p, &out[a]);
			}
		} else if (na > 0) {
			sel_oracle_top1_se(sa, na, p, &out[a]);
		} else if (nb > 0) {
			sel_oracle_top1_se(sb, nb, p, &out[b]);
		}
	}
}

/* ScoreBuffer::topNSE (ScoreBuffer.cpp:279-330), topn > 1: s is sorted in place; the first *n_sel entries of the sorted list go to alignment
 * (sel[j] = their positions in the caller's array).  Returns 0 when the read is reported unmapped (too many equal top scores under strata). */
int sel_oracle_topn_se(cand *s, int n, int topn, const sel_oracle_params *p, int *sel, int *n_sel, int *mapq, int *num_top) {
	sel_oracle_sort(s, n);
	int n_scores = n, n_top = 1;
	while (n_top < n_scores && s[0].score == s[n_top].score) n_top += 1;
	*num_top = n_top;
	if (n_top <= topn || !p->strata) {
		if (p->strata) n_scores = n_top;
		else n_scores = n_scores < topn ? n_scores : topn;
		*mapq = mq_sorted(s, n);
		*n_sel = n_scores;
		for (int j = 0; j < n_scores; ++j) sel[j] = s[j].orig;
		return 1;
	}
	*mapq = 0;
	*n_sel = 0;
	return 0;
}
