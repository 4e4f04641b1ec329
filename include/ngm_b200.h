/* ngm_b200.h -- C ABI of the B200 (sm_100a) alignment backend for NextGenMap.
 *
 * This is the drop-in boundary for NGM's seed-and-extend hot path: everything
 * NGM's ScoreBuffer / AlignmentBuffer reach through the IAlignment plugin
 * interface (reference: include/IAlignment.h:48-69) plus the descriptor fast
 * path that replaces their per-pair host-side window decoding
 * (reference: src/ScoreBuffer.cpp:87-127, src/AlignmentBuffer.cpp:73-115,
 * src/SequenceProvider.cpp:382-441).  Plain pointers and sizes only; no C++,
 * CUDA or torch types cross this boundary.  Every entry point returns >= 0 on
 * success and a negative NGM_B200_E* code on failure; ngm_b200_last_error()
 * gives the text.  There is NO CPU fallback behind any of these calls.
 *
 * C++ plugin exports layered on top (nextgenmap_b200/csrc/plugin.cpp) mirror the
 * reference's historical DLL surface (lib/mason/opencl/SWOcl_export.cpp:20-83):
 * Cookie, SetLog, SetConfig, IsAvailable, CreateAlignment, DeleteAlignment,
 * ExternalDeleteString.
 */
#ifndef NGM_B200_H
#define NGM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library itself is built with -fvisibility=hidden */
#endif

#define NGM_B200_ABI_VERSION 1   /* round 2 only adds entry points */

enum {
	NGM_B200_OK = 0,
	NGM_B200_EINVAL = -1,   /* bad argument / unsupported parameter combination */
	NGM_B200_ECUDA = -2,    /* CUDA runtime error (no device, launch failure, OOM) */
	NGM_B200_ERANGE = -3,   /* scoring parameters outside the exact-integer range of the kernels */
	NGM_B200_ESTATE = -4    /* call sequence error (e.g. descriptor call before set_reference) */
};

/* Alignment modes: `mode & 0xFF` of IAlignment::BatchScore/BatchAlign
 * (reference: SWOcl.cpp:85-102, SWOclCigar.cpp:196-211). */
enum { NGM_B200_MODE_LOCAL = 0, NGM_B200_MODE_ENDFREE = 1 };

/* The Config keys the reference backend reads (SWOcl.cpp:165-166,208-242,328-348;
 * SWOclCigar.cpp:450-454).  Penalties are positive, as in the config file. */
typedef struct ngm_b200_params {
	int32_t qry_max_len;        /* "qry_max_len"  -> read_length                     */
	int32_t corridor;           /* "corridor"     -> band width in cells per row     */
	float match_bonus;          /* "match_bonus"                                      */
	float mismatch_penalty;     /* "mismatch_penalty"                                 */
	float gap_read_penalty;     /* "gap_read_penalty"                                 */
	float gap_ref_penalty;      /* "gap_ref_penalty"                                  */
	float match_bonus_tt;       /* "match_bonus_tt" (bs-mapping / SLAMseq)            */
	float match_bonus_tc;       /* "match_bonus_tc"                                   */
	int32_t bs_mapping;         /* "bs_mapping"                                       */
	int32_t slam_seq;           /* "slam_seq"                                         */
	int32_t hard_clip;          /* "hard_clip"                                        */
	int32_t silent_clip;        /* "silent_clip"                                      */
	int32_t device;             /* CUDA device ordinal (low byte of CreateAlignment's mode, NGM.cpp:405) */
	int32_t score_batch;        /* 0 = default; value returned by GetScoreBatchSize() */
	int32_t align_batch;        /* 0 = default; value returned by GetAlignBatchSize() */
	int32_t lane_mode;          /* 0 = auto, 1 = force int32 lanes, 2 = force s16x2 lanes (testing) */
} ngm_b200_params;

/* Layout-compatible with the reference's `struct Align` (IAlignment.h:14-28):
 * caller-owned CIGAR / MD buffers of at least 4*max(1,qry_max_len) bytes each
 * (AlignmentBuffer.cpp:106-107). */
typedef struct ngm_b200_align {
	char *cigar;                /* Align::pBuffer1 */
	char *md;                   /* Align::pBuffer2 */
	void *extended;             /* Align::ExtendedData (left untouched) */
	int32_t position_offset;    /* Align::PositionOffset */
	int32_t qstart;             /* Align::QStart */
	int32_t qend;               /* Align::QEnd */
	float score;                /* Align::Score (read_index hack, SWOclCigar.cpp:612-613; -1 = failed) */
	float identity;             /* Align::Identity */
	int32_t nm;                 /* Align::NM */
} ngm_b200_align;

/* Fixed-size per-alignment record of the descriptor / device paths. */
typedef struct ngm_b200_align_rec {
	int32_t position_offset;
	int32_t qstart;
	int32_t qend;
	int32_t nm;
	float identity;
	float score;
	uint32_t str_off;           /* offset of the CIGAR bytes in the string heap; MD follows the CIGAR */
	uint16_t cigar_len;
	uint16_t md_len;            /* bytes written incl. embedded NULs (SURVEY 8a note 9) */
} ngm_b200_align_rec;

/* One (read, candidate window) pair of the descriptor fast path: what
 * ScoreBuffer::DoRun derives per pair before decoding the window on the host
 * (ScoreBuffer.cpp:92-118). */
typedef struct ngm_b200_pair {
	uint64_t window_start;      /* concatenated-reference position of the window: loc - corridor/2 */
	uint32_t read_index;        /* row in the uploaded read batch */
	uint32_t flags;             /* bit0: use the reverse-complement read (RevSeq); bit1: direction flag (extData) */
} ngm_b200_pair;

#define NGM_B200_PAIR_REVERSE 1u   /* use RevSeq */
#define NGM_B200_PAIR_DIR 2u       /* extData direction flag */
#define NGM_B200_PAIR_SKIP 4u      /* do not evaluate (score -1 / -16000, alignment record score -1) */

typedef struct ngm_b200_ctx ngm_b200_ctx;

/* -- lifetime ------------------------------------------------------------ */
int ngm_b200_abi_version(void);
int ngm_b200_device_count(void);
const char *ngm_b200_last_error(void);
ngm_b200_ctx *ngm_b200_create(const ngm_b200_params *params);
void ngm_b200_destroy(ngm_b200_ctx *ctx);
/* A context on the same device that borrows `root`'s resident data (packed reference, k-mer index, selection parameters) and owns
 * only its stream, read batch and scratch.  NGM creates one IAlignment per CS thread (CS.cpp:455-461; the reference shares its OpenCL
 * context the same way, OclHost.cpp:50,133): with this, sixteen threads hold ONE copy of a 3 Gbp reference.  Destroy it before the
 * root.  ngm_b200_sync_shared re-borrows after the root's reference / index changed; the caller serialises it against those changes. */
ngm_b200_ctx *ngm_b200_create_shared(ngm_b200_ctx *root);
int ngm_b200_sync_shared(ngm_b200_ctx *ctx);
/* IAlignment::GetScoreBatchSize / GetAlignBatchSize (IAlignment.h:53-54) */
int ngm_b200_score_batch_size(const ngm_b200_ctx *ctx);
int ngm_b200_align_batch_size(const ngm_b200_ctx *ctx);

/* -- strict IAlignment path (host ASCII buffers) ------------------------- */
/* IAlignment::BatchScore (IAlignment.h:56-60; reference impl SWOcl.cpp:33-162).
 * ref[i]: >= qry_max_len+corridor readable bytes; qry[i]: >= qry_max_len bytes,
 * NUL padded; dir: n direction flags or NULL.  Writes n floats; returns n. */
int ngm_b200_batch_score(ngm_b200_ctx *ctx, int mode, int n, const char *const *ref, const char *const *qry,
		float *scores, const char *dir);
/* IAlignment::BatchAlign (IAlignment.h:62-66; reference impl SWOclCigar.cpp:104-370). */
int ngm_b200_batch_align(ngm_b200_ctx *ctx, int mode, int n, const char *const *ref, const char *const *qry,
		const char *const *qal, ngm_b200_align *results, const char *dir);

/* -- descriptor fast path (reference resident in HBM) -------------------- */
/* Upload the concatenated reference in NGM's own packing: 4 bit/base, high nibble
 * first, A0 T1 G2 C3 N4 (SequenceProvider.cpp:72-109; body of <ref>-enc.2.ngm).
 * concat_len = number of bases (GetConcatRefLen()). */
int ngm_b200_set_reference(ngm_b200_ctx *ctx, const uint8_t *packed, uint64_t concat_len);
/* Upload a batch of reads: n_reads rows of `stride` ASCII bytes (NUL padded, like
 * MappedRead::Seq, MappedRead.cpp:25-29).  Reverse complements are derived on device
 * (MappedRead::computeReverseSeq, MappedRead.cpp:53-67). */
int ngm_b200_set_reads(ngm_b200_ctx *ctx, const char *reads, int n_reads, int stride);
/* Score n pairs (ScoreBuffer::DoRun + BatchScore). */
int ngm_b200_score_pairs(ngm_b200_ctx *ctx, int mode, int n, const ngm_b200_pair *pairs, float *scores);
/* Align n pairs (AlignmentBuffer::DoRun + BatchAlign).  recs: n records; strings:
 * heap of str_capacity bytes; *str_used receives the bytes needed (if it exceeds
 * str_capacity the call returns NGM_B200_ERANGE and must be repeated with a larger heap). */
int ngm_b200_align_pairs(ngm_b200_ctx *ctx, int mode, int n, const ngm_b200_pair *pairs, ngm_b200_align_rec *recs,
		char *strings, size_t str_capacity, size_t *str_used);

/* -- reference cache file of NGM ---------------------------- */
/* One contig of the concatenated reference: RefIdx of SequenceProvider.h:45-52. */
typedef struct ngm_b200_contig {
	uint64_t start;             /* SeqStart: concatenated coordinate of the first base */
	uint32_t length;            /* SeqLen */
	uint32_t name_len;
	char name[100];             /* not NUL terminated when name_len == 100 */
} ngm_b200_contig;

typedef struct ngm_b200_encref {
	uint64_t concat_len;        /* GetConcatRefLen() = binRefIndex - 1 (SequenceProvider.cpp:454-456) */
	uint64_t packed_bytes;      /* encRefSize */
	uint32_t n_contigs;
	uint8_t *packed;            /* binRef: 4 bit/base, high nibble first, A0 T1 G2 C3 N4 */
	ngm_b200_contig *contigs;
} ngm_b200_encref;

/* Read `<ref>-enc.2.ngm` as written by _SequenceProvider::writeEncRefToFile (SequenceProvider.cpp:189-208;
 * cookie 0x74656).  Host only (no CUDA call).  Free with ngm_b200_free_enc_ref. */
int ngm_b200_read_enc_ref(const char *path, ngm_b200_encref *out);
void ngm_b200_free_enc_ref(ngm_b200_encref *ref);
/* Write such a file (SequenceProvider.cpp:189-208): NGM started on `<ref>` then loads `<ref>-enc.2.ngm` instead of encoding the FASTA. */
int ngm_b200_write_enc_ref(const char *path, const ngm_b200_encref *ref);
/* _SequenceProvider::convert (SequenceProvider.cpp:111-141): concatenated position -> (contig, position).
 * Returns 1, or 0 when the position lies in the 1000-N spacer in front of the next contig (reported as
 * unmapped by NGM). */
int ngm_b200_convert(const ngm_b200_encref *ref, uint64_t concat_pos, uint32_t *contig, uint64_t *pos);

/* -- candidate search: k-mer index + per-read vote --------------------------------- */
/* The Config keys CS / CompactPrefixTable read (src/config/Config.cpp:383-390,512-518; CS.cpp:476-481,563-570). */
typedef struct ngm_b200_cs_params {
	int32_t kmer;               /* "kmer" (13); 8..14 */
	int32_t kmer_skip;          /* "kmer_skip" (2): every (kmer_skip+1)-th reference position is indexed */
	int32_t bin_size;           /* "bin_size" (2): votes are collected per 2^bin_size reference positions */
	int32_t skip_rep;           /* CompactPrefixTable(dualStrand, skip = true): drop repeated k-mers inside one bin (1) */
	float sensitivity;          /* "sensitivity" (-s; estimated or 0.5): threshold = sensitivity x best vote */
	float kmer_min;             /* "kmer_min" (0) */
	int32_t max_kfreq;          /* "max_kfreq": 0 = derive like CompactPrefixTable::stats (ceil(max(100, mean + 5 sd))) */
	int32_t max_cmrs;           /* "max_cmrs": 0 = INT_MAX; reads with that many candidates keep none */
} ngm_b200_cs_params;

/* Build the prefix table of CompactPrefixTable (PrefixTable.cpp:196-245,357-498,577-738) on the device from the
 * reference given to ngm_b200_set_reference / ngm_b200_dev_set_reference.  contigs: host array (SeqStart, SeqLen) as in
 * <ref>-enc.2.ngm.  Synchronous.  Single table unit: the concatenated reference must be shorter than 2^32 - 1. */
int ngm_b200_cs_build_index(ngm_b200_ctx *ctx, const ngm_b200_cs_params *params, const ngm_b200_contig *contigs, uint32_t n_contigs);
/* Upload a prefix table read from NGM's cache file `<ref>-ht-<kmer>-<kmer_skip>.3.ngm` (see ngm_b200_read_ht_file):
 * tab = Index::m_TabIndex, weight = Index::m_RevCompIndex (index_len = 4^kmer + 1 entries), table = Location::m_Location. */
int ngm_b200_cs_load_index(ngm_b200_ctx *ctx, const ngm_b200_cs_params *params, const uint32_t *tab, const int8_t *weight,
		uint32_t index_len, const uint32_t *table, uint32_t table_len);
int ngm_b200_cs_index_info(const ngm_b200_ctx *ctx, uint32_t *index_len, uint32_t *table_len, int32_t *max_kfreq);
/* Copy the device index back in the file's representation (tab: index_len entries, weight: index_len, table: table_len). */
int ngm_b200_cs_export_index(ngm_b200_ctx *ctx, uint32_t *tab, int8_t *weight, uint32_t *table);

/* Device-to-device forms for multi-GPU start-up (the index is broadcast over NCCL): the sender exports into device buffers
 * (index_len x uint32, index_len x int8, table_len x uint32 -- sizes from ngm_b200_cs_index_info), the receivers install what arrived.
 * params->max_kfreq must carry the sender's value.  Enqueued on `stream`; the load synchronises it. */
int ngm_b200_dev_cs_export_index(ngm_b200_ctx *ctx, void *d_tab, void *d_weight, void *d_table, void *stream);
int ngm_b200_dev_cs_load_index(ngm_b200_ctx *ctx, const ngm_b200_cs_params *params, const void *d_tab, const void *d_weight, uint32_t index_len,
		const void *d_table, uint32_t table_len, void *stream);
/* NGM's prefix-table cache file (CompactPrefixTable::saveToFile / readFromFile, PrefixTable.cpp:819-921). Host only. */
typedef struct ngm_b200_htfile {
	uint32_t kmer, kmer_skip, index_len, table_len;
	uint32_t *tab;              /* index_len entries */
	int8_t *weight;             /* index_len entries */
	uint32_t *table;            /* table_len entries */
	uint64_t unit_offset;
} ngm_b200_htfile;
int ngm_b200_read_ht_file(const char *path, ngm_b200_htfile *out);
void ngm_b200_free_ht_file(ngm_b200_htfile *ht);
int ngm_b200_write_ht_file(const char *path, const ngm_b200_htfile *ht);

/* Candidate search for n_reads rows of `stride` ASCII bytes (NUL padded, like MappedRead::Seq): CS::RunBatch without
 * the bs-mapping k-mer mutation unless ngm_b200_cs_configure_mutation switched it on (CS.cpp:340-436).  cand_begin: n_reads + 1 offsets; pairs[cand_begin[r] .. cand_begin[r+1])
 * are read r's candidates in the reference's order as descriptors for ngm_b200_score_pairs (window_start =
 * Location - corridor/2, flags = REVERSE|DIR for minus-strand candidates, ScoreBuffer.cpp:92-114); votes = LocationScore
 * Score.f; max_hit (optional) = MappedRead::s.  *total receives the number of candidates; if it exceeds `capacity` the
 * call returns NGM_B200_ERANGE and must be repeated with larger buffers.  mode_flags bit 0: use only the sequential
 * exact kernel (testing). */
int ngm_b200_cs_search(ngm_b200_ctx *ctx, const char *reads, int n_reads, int stride, int mode_flags, int32_t *cand_begin,
		ngm_b200_pair *pairs, float *votes, size_t capacity, size_t *total, float *max_hit);
/* The sensitivity NGM estimates when -s is not given (ReadProvider::init, ReadProvider.cpp:236-251,310-325,53-79): pass the
 * reads number 1000, 2000, ... (1-based) of the input as `sampled_reads` (ReadProvider samples every 1000th of the first 10 M, and
 * only estimates from inputs of >= 1000 reads); *sensitivity = clamp(mean(best vote / possible votes), 0.3, 0.9), without the
 * --fast / --sensitive modifiers.  Returns the number of reads that contributed.  ngm_b200_cs_set_sensitivity installs a value.
 * mode_flags bit 1 of the search calls selects the vote count used here (both strands added) for max_hit[]. */
int ngm_b200_cs_estimate_sensitivity(ngm_b200_ctx *ctx, const char *sampled_reads, int n, int stride, float *sensitivity);
int ngm_b200_cs_set_sensitivity(ngm_b200_ctx *ctx, float sensitivity);
/* bs-mapping / SLAMseq candidate search (CS::PrefixMutateSearch, CS.cpp:53-112; CS::RunBatch, CS.cpp:341-380).  bs_mapping = "bs_mapping" (1: every
 * read k-mer is also looked up with every subset of its T -- second mate: A -- replaced by C -- G --; k-mers with more than bs_cutoff = "bs_cutoff" (6)
 * such bases are skipped; read k-mers are taken every read_kmer_skip + 1 positions -- under bs_mapping the "kmer_skip" key (2) thins the READ's k-mers,
 * CS.cpp:556-560, and the index must have been built with ngm_b200_cs_params.kmer_skip 0 as CompactPrefixTable does, PrefixTable.cpp:204-207).  slam_seq = "slam_seq" (bit 2 set: the k-mer votes with weight 1, each single
 * C -> T -- second mate: G -> A -- replacement with weight 1 / (replaceable bases + 1), so votes become fractional).  paired = "paired": odd rows of
 * a read batch are second mates.  Both zero switches the mutation off.  With mutation on, every read is searched by the sequential kernel in the
 * table of CS::RunBatch's last overflow retry (2^20 slots, CS.cpp:404-428); a read that overflows it keeps no candidates, as in the reference.
 * Call after cs_build_index / cs_load_index (a new index starts with the mutation off); applies to cs_search, dev_cs_search and map_batch (sub-batches must hold whole pairs).  The sensitivity
 * estimate is not affected: ReadProvider::init estimates with plain k-mers under slam_seq and, under bs_mapping, uses 0.5 unless -s is given
 * (ReadProvider.cpp:194-197,326,378-384). */
int ngm_b200_cs_configure_mutation(ngm_b200_ctx *ctx, int bs_mapping, int slam_seq, int bs_cutoff, int paired, int read_kmer_skip);
/* Reads the last ngm_b200_cs_search call routed to the exact kernel. */
uint64_t ngm_b200_cs_exact_reads(const ngm_b200_ctx *ctx);
/* Why they left the block-per-read kernel (diagnostics): out[0..n) = counts per reason -- 0 more hits than the kernel's
 * hit buffer, 1 64-bit bin wrap-around, 2 repeat queue full, 3 table crowded, 4 too many two-vote entries, 5 zero threshold,
 * 6 more than 192 accepted entries, 7 too many hits to replay for the order, 8 order check failed.  Returns the number of reasons. */
int ngm_b200_cs_exact_reasons(const ngm_b200_ctx *ctx, uint32_t *out, int n);

/* -- device-pointer entry points (resident pipelines, bench.py `value`) --- */
/* All pointers are device pointers on ctx's device; work is enqueued on `stream`
 * (a cudaStream_t passed as void*) and NOT synchronised.
 * A CONTEXT IS SINGLE-STREAM AND SINGLE-THREAD: every entry point writes per-context scratch (resolved pairs, pointer matrices, op
 * stacks, candidate-search heaps, the paired selector's state) and reads the read batch installed last.  Calls on one context must all
 * use the same stream and must not overlap in time; concurrency comes from several contexts -- ngm_b200_create_shared gives each host
 * thread / stream its own context over one resident reference, which is what ngm_b200_run_batch does internally with its lanes.  d_pairs as above;
 * d_scores: n floats; d_recs: n records; d_strings/d_str_cursor: heap + 1 uint32 cursor
 * (zeroed by the caller before the call). */
int ngm_b200_dev_set_reference(ngm_b200_ctx *ctx, const void *d_packed, uint64_t concat_len, void *stream);
int ngm_b200_dev_set_reads(ngm_b200_ctx *ctx, const void *d_ascii_reads, int n_reads, int stride, void *stream);
/* Winner gather between scoring and alignment (what ScoreBuffer::top1SE hands to
 * AlignmentBuffer::addRead, ScoreBuffer.cpp:259-276): out[r] = pairs[best_pair[r]], or a pair
 * flagged NGM_B200_PAIR_SKIP when read r has no candidate (its record comes back with score -1). */
int ngm_b200_dev_gather_winners(ngm_b200_ctx *ctx, int n_reads, const void *d_pairs, const void *d_best_pair, void *d_out_pairs,
		void *stream);
/* Same, and out_scores[r] = scores[best_pair[r]] (0 when read r has no candidate): the winners' BatchScore results,
 * which ScoreBuffer keeps in LocationScore::Score.f (ScoreBuffer.cpp:151-153) and which
 * ngm_b200_dev_align_pairs_scored takes back. */
int ngm_b200_dev_gather_winners_scored(ngm_b200_ctx *ctx, int n_reads, const void *d_pairs, const void *d_scores,
		const void *d_best_pair, void *d_out_pairs, void *d_out_scores, void *stream);
int ngm_b200_dev_score_pairs(ngm_b200_ctx *ctx, int mode, int n, const void *d_pairs, void *d_scores, void *stream);
int ngm_b200_dev_align_pairs(ngm_b200_ctx *ctx, int mode, int n, const void *d_pairs, void *d_recs, void *d_strings,
		uint32_t str_capacity, void *d_str_cursor, void *stream);
/* BatchAlign of pairs whose BatchScore result in the SAME mode is already known (d_pair_scores[i] = the float
 * ngm_b200_dev_score_pairs returned for d_pairs[i]).  Wide bands (capacity > 48: `-C 40`, 400 bp reads) locate the
 * best cell by comparing against that maximum and so skip their internal score pass; narrow bands ignore it (their
 * snapshot kernel is faster).  Results are identical to ngm_b200_dev_align_pairs; on the wide-band path a score that
 * is NOT the pair's true local maximum yields a failed record (score -1). */
int ngm_b200_dev_align_pairs_scored(ngm_b200_ctx *ctx, int mode, int n, const void *d_pairs, const void *d_pair_scores,
		void *d_recs, void *d_strings, uint32_t str_capacity, void *d_str_cursor, void *stream);
/* Per-read top-1 selection over scored candidates (ScoreBuffer::top1SE + computeMQ,
 * ScoreBuffer.cpp:34-40,228-277).  cand_begin: n_reads+1 offsets into the pair/score arrays.
 * best_pair[r] = index of the winning pair or -1; mapq[r]. */
int ngm_b200_dev_select_top1(ngm_b200_ctx *ctx, int n_reads, const void *d_cand_begin, const void *d_scores,
		void *d_best_pair, void *d_mapq, void *stream);
/* Same, and num_top[r] = number of candidates sharing the best score (top1SE's numBestScore -> MappedRead::numTopScores,
 * ScoreBuffer.cpp:236-251; SAM tags NH / X0, SAMWriter.cpp:171,191). */
int ngm_b200_dev_select_top1_ex(ngm_b200_ctx *ctx, int n_reads, const void *d_cand_begin, const void *d_scores,
		void *d_best_pair, void *d_mapq, void *d_num_top, void *stream);
/* -- paired-end selection (BASELINE configs[2]) ------------------------------------------- */
/* The Config keys ScoreBuffer / NGM read for pairs (src/config/Config.cpp:393-406, ScoreBuffer.h:90-91, NGM.cpp:38-41). */
typedef struct ngm_b200_pe_params {
	float pair_score_cutoff;    /* "pair_score_cutoff" (0.9): candidates below cutoff x best score do not take part in pairing */
	int32_t min_insert_size;    /* "min_insert_size" (0), exclusive */
	int32_t max_insert_size;    /* "max_insert_size" (1000), exclusive; <= 0 = INT_MAX */
	int32_t strata;             /* "strata" (0) */
	int32_t fast_pairing;       /* "fast_pairing" (0): both mates are selected single-end */
} ngm_b200_pe_params;
/* Installs the parameters and resets the running insert-size sums (ScoreBuffer's pairDistSum = 0, pairDistCount = 1): call once per
 * mapping run, before the first batch. */
int ngm_b200_pe_configure(ngm_b200_ctx *ctx, const ngm_b200_pe_params *params);
/* pairDistSum / pairDistCount after the batches selected so far (synchronises the device). */
int ngm_b200_pe_insert_stats(ngm_b200_ctx *ctx, int64_t *dist_sum, int64_t *dist_count);
/* Installs sums read earlier (a batch that has to be repeated, a run that is resumed). */
int ngm_b200_pe_set_insert_stats(ngm_b200_ctx *ctx, int64_t dist_sum, int64_t dist_count);
/* Fragments of the last ngm_b200_dev_select_pairs call that met equal pair scores and were therefore decided one after the other
 * (diagnostics; synchronises the device). */
int64_t ngm_b200_pe_deferred_fragments(ngm_b200_ctx *ctx);
/* What ScoreBuffer::DoRun does once both mates of a fragment are scored (ScoreBuffer.cpp:196-215): top1PE + CheckPairs (:365-502), or
 * top1SE (:228-277) for a mate whose partner has no candidate, under fast_pairing, and as the fallback when no combination has an
 * insert size inside the limits (both mates then carry NGMNames::PairedFail).  Rows 2f and 2f + 1 of the read batch given to
 * ngm_b200_set_reads / ngm_b200_dev_set_reads are the mates of fragment f (ReadId & 1 = second mate); n_reads is even.
 * cand_begin / pairs / scores as for ngm_b200_dev_select_top1 (pairs: n_pairs descriptors as ngm_b200_cs_search emits them, the
 * candidate's Location is window_start + corridor / 2).  Per read: best_pair = index of the candidate handed to alignment or -1,
 * mapq, num_top = MappedRead::numTopScores (NH / X0: top1PE stores the number of equally good pairs there, 0 for a unique pair),
 * pair_fail.  Equal pair scores are broken by the distance to the mean insert size of the pairs accepted before, in input order over
 * all batches since ngm_b200_pe_configure -- the reference's result with one CS thread.  Enqueued on `stream`, not synchronised. */
int ngm_b200_dev_select_pairs(ngm_b200_ctx *ctx, int n_reads, const void *d_cand_begin, const void *d_pairs, const void *d_scores, uint32_t n_pairs,
		void *d_best_pair, void *d_mapq, void *d_num_top, void *d_pair_fail, void *stream);
/* -- a whole batch in one call ------------------------------------------------------------------------------ */
/* Host arrays the caller owns; n = rows of the batch. */
typedef struct ngm_b200_map_result {
	int32_t *cand_begin;        /* n + 1 offsets into pairs / scores */
	ngm_b200_pair *pairs;       /* `capacity` candidates in the reference's order (ngm_b200_cs_search) */
	float *scores;              /* `capacity` BatchScore results */
	size_t capacity;
	size_t n_candidates;        /* out: candidates found (> capacity: NGM_B200_ERANGE, repeat with larger arrays) */
	int32_t *best_pair;         /* n: the candidate handed to alignment or -1 */
	int32_t *mapq;              /* n */
	int32_t *num_top;           /* n: MappedRead::numTopScores */
	int32_t *pair_fail;         /* n, paired batches only */
	float *max_hit;             /* n: MappedRead::s */
	ngm_b200_align_rec *recs;   /* n: alignment of read r's selected candidate (score -1 when there is none) */
	char *strings;              /* string heap of `str_capacity` bytes */
	size_t str_capacity;
	size_t str_used;            /* out: bytes needed (> str_capacity: NGM_B200_ERANGE) */
	/* ngm_b200_se_configure_topn(topn > 1), single-end: recs then holds n x topn records (record r * topn + j = candidate sel[r * topn + j],
	 * j < n_sel[r]; score -1 beyond), best_pair[r] = sel[r * topn].  Not read for topn 1 (may be NULL). */
	int32_t *sel;               /* n x topn */
	int32_t *n_sel;             /* n */
} ngm_b200_map_result;
/* CS::RunBatch -> ScoreBuffer::DoRun (BatchScore, top1SE or top1PE) -> AlignmentBuffer::DoRun (BatchAlign) for one batch of reads, from
 * host buffers to host buffers (CS.cpp:340-436, ScoreBuffer.cpp:80-277,365-502, AlignmentBuffer.cpp:64-147).  Needs ngm_b200_set_reference
 * and a prefix table (ngm_b200_cs_build_index / cs_load_index); paired != 0 additionally ngm_b200_pe_configure, rows 2f / 2f + 1 = mates.
 * Synchronous.  The result feeds ngm_b200_format_sam (its ngm_b200_sam_batch takes the same arrays). */
int ngm_b200_map_batch(ngm_b200_ctx *ctx, const char *reads, int n_reads, int stride, int mode, int paired, ngm_b200_map_result *result);

/* -- SAM records of a batch ------------------------------------------------------------------ */
/* Output filters and pair limits (src/config/Config.cpp:405-410,430-431; GenericReadWriter.h:204-214,263-283). */
typedef struct ngm_b200_sam_opts {
	float min_identity;         /* "min_identity" (0.65) */
	float min_residues;         /* "min_residues" (0.5): <= 1 = fraction of the read length */
	int32_t min_insert_size;    /* "min_insert_size" (0), inclusive in the check of the aligned mates */
	int32_t max_insert_size;    /* "max_insert_size" (1000); <= 0 = INT_MAX */
	int32_t threads;            /* host threads; 0 = all */
	int32_t min_mq;             /* "min_mq" (0): reads below it are written as unmapped (AlignmentBuffer.cpp:46-49, GenericReadWriter.h:281-284) */
	int32_t clip_seq;           /* "hard_clip" or "silent_clip" set: SEQ / QUAL of a mapped read lose the clipped ends (SAMWriter.cpp:104,146-160);
	                             * the CIGAR's H ops (or none) come from the alignment itself (ngm_b200_params.hard_clip / silent_clip) */
	const char *read_group;     /* "rg_id": RG:Z:<id> on every record (SAMWriter.cpp:166-168,358-360); NULL = none */
	int32_t bs_mapping;         /* "bs_mapping" (1): ZS:Z:++ / -+ (first mate or single read, forward / reverse) or -- / +- (second mate) after NH:i
	                             * (SAMWriter.cpp:173-187) */
	int32_t slam_seq;           /* "slam_seq" (!= 0): TC:i (T>C conversions; A>G for reverse reads), RA:Z (the 25 reference-base x read-base counts over
	                             * the aligned columns) and MP:Z (type:read position:reference position of every mismatch) behind MD:Z -- what
	                             * computeSlaSeqTags makes of Align::ExtendedData (GenericReadWriter.h:87-181, SAMWriter.cpp:203-222), derived here from
	                             * CIGAR + MD + the read */
} ngm_b200_sam_opts;
/* One batch as the calls above leave it, all host pointers.  Paired runs: rows 2f / 2f + 1 are mates and pair_fail != NULL. */
typedef struct ngm_b200_sam_batch {
	int32_t n_reads, stride;
	const char *reads;              /* n_reads rows of `stride` bytes, NUL padded (MappedRead::Seq) */
	const char *quals;              /* same shape (MappedRead::qlty) */
	const char *const *names;       /* NUL-terminated read names, mate suffix stripped (ReadProvider.cpp:417-420) */
	const ngm_b200_pair *pairs;     /* candidates (ngm_b200_cs_search) */
	const float *scores;            /* BatchScore of every candidate -> AS:i */
	const int32_t *best_pair;       /* selection: index into pairs or -1 */
	const int32_t *mapq;
	const int32_t *num_top;         /* NH:i / X0:i */
	const int32_t *pair_fail;       /* NULL = single-end */
	const float *max_hit;           /* MappedRead::s -> XE:i */
	const ngm_b200_align_rec *recs; /* alignment of read r's selected candidate (ngm_b200_align_pairs with read_index = r) */
	const char *strings;            /* its string heap */
	/* "topn" > 1 (single-end only, ngm_b200_dev_select_topn): recs holds n_reads x topn records, record r * topn + j belongs to candidate
	 * sel[r * topn + j], j < n_sel[r]; best_pair is not read.  topn <= 1: leave these zero. */
	int32_t topn;
	const int32_t *sel;
	const int32_t *n_sel;
} ngm_b200_sam_batch;
/* SAM body lines (no header) of the batch in read order, what AlignmentBuffer::WriteRead (AlignmentBuffer.cpp:166-200),
 * GenericReadWriter::WriteRead / WritePair (GenericReadWriter.h:190-312) and SAMWriter::DoWriteReadGeneric / DoWriteUnmappedReadGeneric /
 * DoWritePair (SAMWriter.cpp:98-228,230-365) produce, incl. several alignments per read (topn), the clipping options, read group, and the
 * bs-mapping / SLAMseq tags (ngm_b200_sam_opts).  Host only, multi-threaded.  *out_used receives the bytes needed; NGM_B200_ERANGE if that
 * exceeds out_capacity (nothing is written then).  The library keeps its per-thread scratch (at most 1 GiB, about the size of the largest
 * batch's text) from call to call; concurrent calls are allowed, all but one of them work in scratch of their own. */
int ngm_b200_format_sam(const ngm_b200_encref *ref, const ngm_b200_sam_opts *opts, const ngm_b200_sam_batch *batch, char *out, size_t out_capacity,
		size_t *out_used);
/* Single-end selection with "topn" > 1 (ScoreBuffer::topNSE, ScoreBuffer.cpp:279-330; `-n`, `--strata`): the candidates are sorted by score
 * (std::sort's order of equal scores included); sel[r * topn + j], j < n_sel[r], = the candidates handed to alignment, best first (-1 beyond
 * n_sel[r]); mapq from the two best scores; num_top = candidates sharing the best score.  Under strata a read with more than topn equally
 * good candidates keeps none.  topn = 1 runs take ngm_b200_dev_select_top1 (the reference's top1SE differs: no sort, first of equals). */
int ngm_b200_dev_select_topn(ngm_b200_ctx *ctx, int n_reads, const void *d_cand_begin, const void *d_scores, uint32_t n_pairs, int topn, int strata,
		void *d_sel, void *d_n_sel, void *d_mapq, void *d_num_top, void *stream);
/* Device-pointer form of ngm_b200_cs_search: enqueued on `stream`, not synchronised; the caller checks
 * cand_begin[n_reads] <= capacity afterwards.  d_votes / d_max_hit may be NULL. */
int ngm_b200_dev_cs_search(ngm_b200_ctx *ctx, const void *d_ascii_reads, int n_reads, int stride, int mode_flags, void *d_cand_begin,
		void *d_pairs, void *d_votes, uint32_t capacity, void *d_max_hit, void *stream);
/* -- a whole scored + aligned batch in one call, pipelined inside the library -------------------------------------
 * ScoreBuffer::DoRun (window fetch, BatchScore, top1SE / top1PE + computeMQ; ScoreBuffer.cpp:80-277,365-502) followed by
 * AlignmentBuffer::DoRun (BatchAlign + computeCigarMD; AlignmentBuffer.cpp:64-147) for n_reads reads whose candidate lists the
 * caller already holds (CS::SendToBuffer, CS.cpp:320-335).  This is the entry point north_star's re-plumbed ScoreBuffer /
 * AlignmentBuffer submit to: reads and (read x candidate) descriptors packed into pinned staging buffers.
 *
 * Inside, the batch is cut into sub-batches that rotate over several lanes (stream + device staging each): the host->device copy of
 * sub-batch i+1, the kernels of sub-batch i and the device->host copy of sub-batch i-1 overlap, like the reference overlaps packing
 * with its kernels (SWOcl.cpp:435-444).  Reads with a single candidate skip BatchScore: their score is the maximum the alignment's
 * forward pass finds anyway (same recurrence), so results are identical.
 *
 * Read formats.  ASCII: n_reads rows of `read_stride` bytes, NUL padded (MappedRead::Seq).  PACKED2: rows of `read_stride` bytes
 * (a multiple of 4) holding 2 bits per base, base i in bits [2(i & 15), +2) of little-endian word i >> 4, A0 C1 G2 T3; read_len[r] =
 * bases in row r; every base that is not A/C/G/T is listed in `exceptions` (sorted by read, then position) with its ASCII byte.
 * 150 bp: 40 + 2 instead of 152 bytes per read cross PCIe.
 * Descriptor formats.  PAIR16: struct ngm_b200_pair -- its read_index is ignored, candidate lists are given by cand_begin.  U64: NGM_B200_DESC(). */
typedef struct ngm_b200_read_exc {
	uint32_t read_index;
	uint16_t pos;
	uint8_t ch;                 /* the ASCII byte ('N', IUPAC, ...) */
	uint8_t pad;
} ngm_b200_read_exc;

enum { NGM_B200_READS_ASCII = 0, NGM_B200_READS_PACKED2 = 1 };
enum { NGM_B200_DESC_PAIR16 = 0, NGM_B200_DESC_U64 = 1 };
/* window_start in bits 0..55 (saturated: a start that underflowed or lies behind the reference selects the all-'N' window either way,
 * ScoreBuffer.cpp:113-118), NGM_B200_PAIR_* flags in bits 56..63 */
#define NGM_B200_DESC(window_start, flags) \
	((((uint64_t) (window_start)) > 0x00FFFFFFFFFFFFFFull ? 0x00FFFFFFFFFFFFFFull : ((uint64_t) (window_start))) | ((uint64_t) (flags) << 56))

typedef struct ngm_b200_batch_in {
	int32_t n_reads;
	int32_t mode;               /* NGM_B200_MODE_* */
	int32_t paired;             /* rows 2f / 2f + 1 are mates; needs ngm_b200_pe_configure */
	int32_t read_format;        /* NGM_B200_READS_* */
	const void *reads;
	int32_t read_stride;        /* bytes per row */
	int32_t desc_format;        /* NGM_B200_DESC_* */
	const uint16_t *read_len;   /* PACKED2 only */
	const ngm_b200_read_exc *exceptions;   /* PACKED2 only; may be NULL when n_exceptions == 0 */
	uint32_t n_exceptions;
	uint32_t n_desc;            /* ngm_b200_dev_run_batch: number of descriptors (0 = cand_begin[n_reads] is read back: one stream synchronisation) */
	const int32_t *cand_begin;  /* n_reads + 1 offsets into desc */
	const void *desc;           /* cand_begin[n_reads] descriptors */
} ngm_b200_batch_in;

typedef struct ngm_b200_batch_out {
	float *scores;              /* optional (may be NULL): cand_begin[n_reads] BatchScore results (LocationScore::Score.f) */
	int32_t *best_pair;         /* n_reads: index of the candidate handed to alignment, or -1 */
	int32_t *mapq;              /* n_reads */
	int32_t *num_top;           /* n_reads, optional */
	int32_t *pair_fail;         /* n_reads, paired batches only */
	ngm_b200_align_rec *recs;   /* n_reads: alignment of the selected candidate (score -1 when there is none) */
	char *strings;              /* CIGAR / MD heap; recs[].str_off are offsets into it.  The heap is filled per sub-batch, so it is sparse:
	                             * every sub-batch owns the part of the heap that corresponds to its share of the reads (16-byte aligned) */
	size_t str_capacity;
	size_t str_used;            /* out: bytes of CIGAR / MD text.  NGM_B200_ERANGE: a sub-batch needed more than its part; str_used then holds a
	                             * sufficient total capacity for a repeat of the call */
	uint32_t *d_str_cursor;     /* ngm_b200_dev_run_batch only: device word that receives the heap bytes used (dense heap, one sub-batch) */
	/* ngm_b200_se_configure_topn(topn > 1), single-end batches (ScoreBuffer::topNSE): every candidate is scored, the list sorted, up to topn
	 * candidates per read aligned.  recs then holds n_reads x topn records (record r * topn + j = candidate sel[r * topn + j], j < n_sel[r];
	 * score -1 beyond), best_pair[r] = sel[r * topn], mapq / num_top as ngm_b200_dev_select_topn.  Not read for topn 1 (may be NULL). */
	int32_t *sel;               /* n_reads x topn */
	int32_t *n_sel;             /* n_reads */
} ngm_b200_batch_out;

/* Host buffers in, host buffers out; returns n_reads.  For copies that overlap with the kernels the buffers must be page-locked:
 * ngm_b200_host_alloc / ngm_b200_host_register (pageable memory works, but every copy then stalls its lane). */
int ngm_b200_run_batch(ngm_b200_ctx *ctx, const ngm_b200_batch_in *in, ngm_b200_batch_out *out);
/* The same with every pointer of `in` / `out` on the device: one sub-batch, enqueued on `stream`, not synchronised (resident
 * pipelines; bench.py `value`).  out->str_used is not written; out->d_str_cursor receives the bytes used. */
int ngm_b200_dev_run_batch(ngm_b200_ctx *ctx, const ngm_b200_batch_in *in, ngm_b200_batch_out *out, void *stream);
/* ngm_b200_dev_set_reads for PACKED2 rows (device pointers; see the formats above): installs the batch for the descriptor entry points. */
int ngm_b200_dev_set_reads_packed(ngm_b200_ctx *ctx, const void *d_packed, int n_reads, int stride, const void *d_read_len, const void *d_exceptions,
		uint32_t n_exceptions, void *stream);
/* Lanes (1..8, default 4) and reads per sub-batch (default 1 << 19) of ngm_b200_run_batch / ngm_b200_map_batch. */
int ngm_b200_set_pipeline(ngm_b200_ctx *ctx, int lanes, int sub_batch_reads);
/* "strata" for single-end runs (ScoreBuffer::top1SE, ScoreBuffer.cpp:259-276): a read with several equally best candidates is reported
 * unmapped (best_pair -1, mapq 0).  Applies to ngm_b200_dev_select_top1[_ex], ngm_b200_run_batch and ngm_b200_map_batch. */
int ngm_b200_se_configure(ngm_b200_ctx *ctx, int strata);
/* "topn" (NGM -n, 1..1000; default 1) of the single-end batches of ngm_b200_run_batch / ngm_b200_dev_run_batch / ngm_b200_map_batch:
 * > 1 switches them from top1SE to topNSE (ScoreBuffer.cpp:279-330; `strata` as configured above) -- see sel / n_sel of the result structs. */
int ngm_b200_se_configure_topn(ngm_b200_ctx *ctx, int topn);
/* Page-locked host memory for the staging buffers (cudaHostAlloc / cudaHostRegister). */
void *ngm_b200_host_alloc(size_t bytes);
void ngm_b200_host_free(void *p);
int ngm_b200_host_register(void *p, size_t bytes);
int ngm_b200_host_unregister(void *p);
/* ASCII rows -> PACKED2 on the host (what a re-plumbed ScoreBuffer does while it fills its staging buffer).  packed: n rows of
 * row_bytes (>= 4 * ceil(stride / 16)); exceptions: capacity exc_cap, *n_exc receives the number needed (NGM_B200_ERANGE if larger).
 * threads: host threads to use (0 = all). */
int ngm_b200_pack_reads(const char *ascii, int n_reads, int stride, void *packed, int row_bytes, uint16_t *read_len, ngm_b200_read_exc *exceptions,
		size_t exc_cap, size_t *n_exc, int threads);

/* -- measurement helpers (bench.py) ------------------------------------------------------------------------- */
/* Issue rates measured on this device, in thread-level instructions per second: VIADDMNMX.S16x2 (the integer ALU pipe the DP
 * recurrence runs on), IMAD (the FMA pipe the tag arithmetic is moved to) and both interleaved 1:1.  Denominators of bench.py's
 * roofline_alu instead of a nominal lanes x clock model.  Synchronous, ~50 ms. */
int ngm_b200_alu_peak(ngm_b200_ctx *ctx, double *viaddmnmx_per_s, double *imad_per_s, double *mixed_per_s);
/* enable != 0: the following align launch sets on this context are bracketed by events (forward kernel | backtrace kernel);
 * ngm_b200_profile_read synchronises, returns the number of launch sets and their summed durations in ms, and re-arms. */
int ngm_b200_profile(ngm_b200_ctx *ctx, int enable);
int ngm_b200_profile_read(ngm_b200_ctx *ctx, float *forward_ms, float *backtrace_ms);
/* EXPERIMENT, not a product path: north_star's warp-per-tile skewed wavefront (shuffles between lanes) for the local score, timed
 * against the production thread-per-pair kernel by bench.py (`design_ab`).  d_pairs: ngm_b200_pair descriptors as for
 * ngm_b200_dev_score_pairs (set_reference / set_reads first); corridor <= 32, local mode. */
int ngm_b200_exp_wavefront_score(ngm_b200_ctx *ctx, int n, const void *d_pairs, void *d_scores, void *stream);
/* MEASUREMENT, not a product path (csrc/exp_issue.cu, scripts/issue_rates.py): issue rates, in thread-level instructions per second, of the
 * instructions the DP kernels are made of, alone (op_b = -1) and interleaved 1:1 in pairs -- which of them share the integer ALU pipe.
 * Writes up to `cap` cases, returns their number. */
int ngm_b200_exp_issue_rates(ngm_b200_ctx *ctx, int cap, int *op_a, int *op_b, double *per_s);
/* Number of kernels this context has launched since creation (bench.py gpu_launches). */
uint64_t ngm_b200_launch_count(const ngm_b200_ctx *ctx);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
