/* oracle/select_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of NextGenMap's selection step between scoring and alignment: ScoreBuffer::top1SE, computeMQ,
 * top1PE and CheckPairs (src/ScoreBuffer.cpp:34-48,196-215,228-277,365-502).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * legs may load this; the product never does.
 *
 * Parity status: PINNED through whole runs -- tests/test_mapper_oracle.py strings the oracles together (candidate search, scores,
 * this selection, alignments) and requires the SAM of the unmodified NextGenMap (oracle/_ref/ngm/ngm_ref, `-p -t 1` for pairs), on
 * inputs with duplicated reference segments so that equal pair scores and the insert-size tie-break occur.
 */
#ifndef NGM_SELECT_ORACLE_H
#define NGM_SELECT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sel_oracle_cand {    /* LocationScore: Location.m_Location (CS: ResolveBin(bin)), Score.f (BatchScore) */
	uint64_t location;
	float score;
	int orig;                       /* position in the caller's array (filled by sel_oracle_select_pairs) */
} sel_oracle_cand;

typedef struct sel_oracle_params {  /* Config.cpp:393-406 */
	float pair_score_cutoff;        /* "pair_score_cutoff" 0.9 */
	int min_insert, max_insert;     /* "min_insert_size" 0, "max_insert_size" 1000 (<= 0: INT_MAX, NGM.cpp:38-41) */
	int strata;                     /* "strata" 0 */
	int fast_pairing;               /* "fast_pairing" 0 */
} sel_oracle_params;

typedef struct sel_oracle_state {   /* ScoreBuffer::pairDistSum / pairDistCount (ScoreBuffer.h:31-32,90: 0 and 1) */
	long long dist_sum, dist_count;
} sel_oracle_state;

typedef struct sel_oracle_result {
	int best;                       /* index of the candidate handed to AlignmentBuffer, -1 = none (read reported unmapped) */
	int mapq;                       /* MappedRead::mappingQlty */
	int num_top;                    /* MappedRead::numTopScores (SAM NH / X0) */
	int paired_fail;                /* NGMNames::PairedFail set by top1PE */
	int insert;                     /* insert size of the chosen pair (top1PE's `distance`), 0 otherwise */
} sel_oracle_result;

void sel_oracle_sort(sel_oracle_cand *first, int n);           /* std::sort(first, first + n, sortLocationScore) */
void sel_oracle_top1_se(const sel_oracle_cand *s, int n, const sel_oracle_params *p, sel_oracle_result *out);
void sel_oracle_top1_pe(sel_oracle_cand *s_read, int n_read, int len_read, sel_oracle_cand *s_mate, int n_mate, int len_mate,
		const sel_oracle_params *p, sel_oracle_state *st, sel_oracle_result *r_read, sel_oracle_result *r_mate);
void sel_oracle_select_pairs(int n_reads, const int *cand_begin, sel_oracle_cand *cands, const int *read_len, const sel_oracle_params *p,
		sel_oracle_state *st, sel_oracle_result *out);

/* ScoreBuffer::topNSE (ScoreBuffer.cpp:279-330): see select_oracle.c. */
int sel_oracle_topn_se(sel_oracle_cand *s, int n, int topn, const sel_oracle_params *p, int *sel, int *n_sel, int *mapq, int *num_top);

#ifdef __cplusplus
}
#endif
#endif
