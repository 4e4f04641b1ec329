/* oracle/cs_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's candidate search (CS.cpp, CSstatic.cpp, PrefixTable.cpp).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this; the product never does.
 *
 * Parity status: PINNED -- the index against the `<ref>-ht-<k>-<skip>.3.ngm` file the unmodified NextGenMap writes
 * (oracle/_ref/ngm/ngm_ref) and the candidate lists against oracle/_ref/ngm/ngm_cs_probe, which runs the reference's
 * own CS::PrefixIteration / PrefixSearch / AddLocationStd / CollectResultsStd (tests/test_cs_oracle_vs_reference.py).
 */
#ifndef NGM_CS_ORACLE_H
#define NGM_CS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cs_oracle_contig {   /* SequenceProvider RefIdx: SeqStart, SeqLen */
	uint64_t start;
	uint64_t length;
} cs_oracle_contig;

typedef struct cs_oracle_index {    /* one TableUnit (PrefixTable.h:62-80) */
	int k, ref_skip, bin_shift;
	uint32_t index_len;             /* 4^k + 1 (refIndexSize in the file) */
	uint32_t table_len;             /* cRefTableLen */
	uint32_t *tab;                  /* Index::m_TabIndex, index_len + 1 entries */
	signed char *weight;            /* Index::m_RevCompIndex; != 0 <=> used() */
	uint32_t *table;                /* Location::m_Location, table_len + 1 entries */
	int max_kfreq;                  /* stats(): ceil(max(100, avg + 5 sd)) */
} cs_oracle_index;

typedef struct cs_oracle_cand {     /* LocationScore as CollectResultsStd fills it */
	uint64_t location;              /* ResolveBin(bin) */
	float score;                    /* Score.f = k-mer votes */
	int reverse;
} cs_oracle_cand;

uint32_t cs_oracle_revcomp(uint32_t prefix, int k);
/* packed: NGM's 4-bit reference (A0 T1 G2 C3 N4, high nibble first).  skip_rep = 1 (CompactPrefixTable default). */
int cs_oracle_build_index(const unsigned char *packed, uint64_t concat_len, const cs_oracle_contig *contigs, int n_contigs, int k, int ref_skip,
		int bin_shift, int skip_rep, cs_oracle_index *ix);
void cs_oracle_free_index(cs_oracle_index *ix);
/* One read (CS::RunBatch body, non-bs path).  Returns the number of candidates (may exceed out_cap; only out_cap are written). */
int cs_oracle_search(const cs_oracle_index *ix, const char *read, int read_len, float sensitivity, float kmer_min, int max_kfreq, int max_cmrs,
		cs_oracle_cand *out, int out_cap, float *max_hit);
/* n_reads rows of `stride` bytes, NUL padded.  cand_begin: n_reads + 1 offsets.  Returns the total number of candidates. */
long long cs_oracle_search_batch(const cs_oracle_index *ix, const char *reads, int n_reads, int stride, float sensitivity, float kmer_min,
		int max_kfreq, int max_cmrs, int *cand_begin, cs_oracle_cand *out, long long out_cap, float *max_hit);

/* CS::RunBatch under bs_mapping (mutate_mode 1: every subset of the read k-mer's T -- second mate: A -- replaced by C -- G --, k-mers with more
 * than bs_cutoff such bases skipped, read k-mers taken every read_skip + 1 positions) or slam_seq & 4 (mutate_mode 2: the k-mer with weight 1, its
 * single replacements C -> T -- second mate: G -> A -- with weight 1 / (replaceable bases + 1)); CS.cpp:53-112,340-436 incl. the table-overflow
 * retries (table_bits = "search_table_length", 16).  paired: odd rows are second mates. */
long long cs_oracle_search_batch_mut(const cs_oracle_index *ix, const char *reads, int n_reads, int stride, float sensitivity, float kmer_min,
		int max_kfreq, int max_cmrs, int mutate_mode, int bs_cutoff, int paired, int read_skip, int table_bits, int *cand_begin,
		cs_oracle_cand *out, long long out_cap, float *max_hit);

/* diagnostics of the last cs_oracle_search_batch_mut call: searches repeated in a larger table, reads that overflowed every table */
long long cs_oracle_last_retries(void);
long long cs_oracle_last_dropped(void);

/* The sensitivity NGM estimates when -s is absent (ReadProvider::init, ReadProvider.cpp:236-251,310-325 with the static PrefixSearch
 * :81-123 and CollectResultsFallback :53-79).  `sampled`: the reads number 1000, 2000, ... of the input.  Returns the number of reads
 * that contributed; *sensitivity = min(max(0.3, mean), 0.9) (no --fast / --sensitive modifier). */
int cs_oracle_estimate_sensitivity(const cs_oracle_index *ix, const char *sampled, int n_reads, int stride, int max_kfreq, float *sensitivity);

#ifdef __cplusplus
}
#endif
#endif
