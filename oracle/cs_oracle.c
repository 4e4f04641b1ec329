/* oracle/cs_oracle.c -- TEST INFRASTRUCTURE ONLY (see cs_oracle.h).
 *
 * Plain-C restatement of the reference's candidate search:
 *   k-mer iteration        src/CSstatic.cpp:20-76        (CS::PrefixIteration)
 *   index construction     src/PrefixTable.cpp:357-428, 500-575, 577-690 (CountKmerFreq, createRefTableIndex,
 *                                                          CreateTable, CountKmer, BuildPrefixTable, SaveToRefTable)
 *   index lookup           src/PrefixTable.cpp:750-817   (GetRefEntry)
 *   max k-mer frequency    src/PrefixTable.cpp:151-194   (stats)
 *   voting                 src/CS.cpp:114-213            (PrefixSearch, AddLocationStd)
 *   result collection      src/CS.cpp:263-313            (CollectResultsStd)
 *   bins                   src/CS.h:164-175              (GetBin, ResolveBin)
 * Single table unit only (concatenated reference < 2^32 - 1 bases, PrefixTable.cpp:24,228).
 */
#include "cs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

int ngm_oracle_decode_window(const unsigned char *packed, unsigned long long concat_len, unsigned long long offset,
		unsigned long long buffer_len, char *buffer);

/* CSstatic.cpp:19-22: A->0 C->1 T->2 G->3 (any other byte maps the same way) */
static inline unsigned enc2(char c) { return ((unsigned) c >> 1) & 3u; }

typedef void (*emit_fn)(uint64_t prefix, uint64_t pos, void *data);

/* CS::PrefixIteration (CSstatic.cpp:26-76); its tail recursion on every 'N' is written as a loop */
static void prefix_iteration(const char *seq, uint64_t length, emit_fn fn, void *data, unsigned prefixskip, uint64_t offset, unsigned k) {
	const uint64_t mask = ((uint64_t) 1 << (2 * k)) - 1;
	for (;;) {
		if (length < k) return;                                     /* :27-28 */
		if (*seq == 'N') {                                          /* :30-41 */
			unsigned n_skip = 1;
			while (seq[n_skip] == 'N') ++n_skip;
			seq += n_skip;
			if (n_skip >= length - k) return;
			length -= n_skip;
			offset += n_skip;
		}
		uint64_t prefix = 0;
		int restart = 0;
		uint64_t i;
		for (i = 0; i < k - 1; ++i) {                               /* :44-54 */
			const char c = seq[i];
			if (c == 'N') {
				restart = 1;
				break;
			}
			prefix = (prefix << 2) | enc2(c);
		}
		if (!restart) {
			unsigned skipcount = prefixskip;                        /* :56-75 */
			for (i = k - 1; i < length; ++i) {
				const char c = seq[i];
				if (c == 'N') {
					restart = 1;
					break;
				}
				prefix = ((prefix << 2) | enc2(c)) & mask;
				if (skipcount == prefixskip) {
					fn(prefix, offset + i + 1 - k, data);
					skipcount = 0;
				} else {
					++skipcount;
				}
			}
		}
		if (!restart) return;
		seq += i + 1;
		length -= i + 1;
		offset += i + 1;
	}
}

/* revComp (PrefixTable.cpp:93-108): complement by XOR 0xAAAAAAAA, then reverse the 2-bit groups of a 32-bit word */
uint32_t cs_oracle_revcomp(uint32_t prefix, int k) {
	const int shift = 32 - 2 * k;
	uint32_t c = (prefix ^ 0xAAAAAAAAu) << shift;
	c = (c & 0xFFFF0000u) >> 16 | (c & 0x0000FFFFu) << 16;
	c = (c & 0xFF00FF00u) >> 8 | (c & 0x00FF00FFu) << 8;
	c = (c & 0xF0F0F0F0u) >> 4 | (c & 0x0F0F0F0Fu) << 4;
	c = (c & 0xCCCCCCCCu) >> 2 | (c & 0x33333333u) << 2;
	return c;
}

/* ---- index construction ------------------------------------------------------------------------------- */
typedef struct {
	int bin_shift;
	int skip_rep;
	uint64_t last_prefix;
	int64_t last_bin;
	int *freq;                  /* count pass */
	cs_oracle_index *ix;        /* build pass */
	uint32_t *fill;             /* build pass: slots used so far per prefix (same result as the first-free-slot scan, :720-738) */
} build_state;

static void count_kmer(uint64_t prefix, uint64_t pos, void *data) {   /* CountKmer / CountKmerwoSkip, PrefixTable.cpp:632-665 */
	build_state *s = (build_state *) data;
	if (!s->skip_rep) {
		s->freq[prefix] += 1;
		return;
	}
	if (prefix == s->last_prefix) {
		const int64_t cur = (int64_t) (pos >> s->bin_shift);
		if (cur != s->last_bin || s->last_bin == -1) s->freq[prefix] += 1;
		s->last_bin = cur;
	} else {
		s->last_bin = -1;
		s->freq[prefix] += 1;
	}
	s->last_prefix = prefix;
}

static void save_location(build_state *s, uint64_t prefix, uint32_t pos) {     /* SaveToRefTable, PrefixTable.cpp:720-738 */
	const uint32_t start = s->ix->tab[prefix] - 1;
	s->ix->table[start + s->fill[prefix]] = pos;
	s->fill[prefix] += 1;
}

static void build_kmer(uint64_t prefix, uint64_t pos, void *data) {   /* BuildPrefixTable / ...woSkip, PrefixTable.cpp:667-718 */
	build_state *s = (build_state *) data;
	const int used = s->ix->weight[prefix] != 0;
	if (!s->skip_rep) {
		if (used) save_location(s, prefix, (uint32_t) pos);
		return;
	}
	if (prefix == s->last_prefix) {
		const int cur = (int) (pos >> s->bin_shift);
		if (cur != s->last_bin || s->last_bin == -1) {
			if (used) save_location(s, prefix, (uint32_t) pos);
		}
		s->last_bin = cur;
	} else {
		s->last_bin = -1;
		if (used) save_location(s, prefix, (uint32_t) pos);
	}
	s->last_prefix = prefix;
}

static void iterate_contigs(const unsigned char *packed, uint64_t concat_len, const cs_oracle_contig *contigs, int n_contigs, int k, int ref_skip,
		emit_fn fn, build_state *s) {
	for (int c = 0; c < n_contigs; ++c) {                           /* CountKmerFreq :362-386, Generate :407-434 */
		s->last_prefix = 111111;
		s->last_bin = -1;
		const uint64_t len = contigs[c].length;
		char *seq = (char *) calloc(len + 2, 1);
		/* DecodeRefSequence is handed the contig length as *buffer* length and therefore decodes two bases less
		 * (SequenceProvider.cpp:384); the tail of `seq` stays NUL and is iterated as code 0 */
		ngm_oracle_decode_window(packed, concat_len, contigs[c].start, len, seq);
		prefix_iteration(seq, len, fn, s, (unsigned) ref_skip, contigs[c].start, (unsigned) k);
		free(seq);
	}
}

int cs_oracle_build_index(const unsigned char *packed, uint64_t concat_len, const cs_oracle_contig *contigs, int n_contigs, int k, int ref_skip,
		int bin_shift, int skip_rep, cs_oracle_index *ix) {
	memset(ix, 0, sizeof(*ix));
	if (k < 4 || k > 15 || concat_len >= 4294967295ull) return -1;
	const uint32_t n_prefix = (uint32_t) 1 << (2 * k);              /* (int) pow(4, k) */
	const uint32_t length = n_prefix + 1;                          /* indexLength, PrefixTable.cpp:209 */
	ix->k = k;
	ix->ref_skip = ref_skip;
	ix->bin_shift = bin_shift;
	ix->index_len = length;
	build_state s;
	memset(&s, 0, sizeof(s));
	s.bin_shift = bin_shift;
	s.skip_rep = skip_rep;
	s.freq = (int *) calloc(length, sizeof(int));
	iterate_contigs(packed, concat_len, contigs, n_contigs, k, ref_skip, count_kmer, &s);
	/* createRefTableIndex, PrefixTable.cpp:436-498 */
	ix->tab = (uint32_t *) calloc((size_t) length + 1, sizeof(uint32_t));
	ix->weight = (signed char *) calloc((size_t) length + 1, 1);
	uint32_t next = 0;
	uint32_t i;
	for (i = 0; i < length - 1; ++i) {
		const uint32_t rc = cs_oracle_revcomp(i, k);
		const int freq = s.freq[i];
		const int total = freq + s.freq[rc];
		ix->tab[i] = next + 1;
		if (freq > 0) {
			const int dummy = 10000;
			ix->weight[i] = (signed char) ((dummy - (total < dummy ? total : dummy)) * 100.0f / dummy);
			next += (uint32_t) freq;
		}
	}
	ix->tab[i] = next + 1;
	ix->table_len = next;
	/* stats(), PrefixTable.cpp:151-194 */
	double sum = 0.0, sum2 = 0.0;
	for (uint32_t j = 0; j < n_prefix; ++j) {
		const double cnt = (double) (ix->tab[j + 1] - ix->tab[j]);
		sum += cnt;
		sum2 += pow(cnt, 2.0);
	}
	const double il = (double) n_prefix;
	const double avg = sum / il;
	const double stdev = sqrt(sum2 / (il - 1) - 2.0 * avg * (sum / (il - 1)) + ((il * pow(avg, 2.0)) / (il - 1)));
	const double m = avg + 5 * stdev;
	ix->max_kfreq = (int) ceil(m > 100.0 ? m : 100.0);
	/* CreateTable :599-606 + Generate */
	ix->table = (uint32_t *) calloc((size_t) next + 1, sizeof(uint32_t));
	s.ix = ix;
	s.fill = (uint32_t *) calloc(length, sizeof(uint32_t));
	iterate_contigs(packed, concat_len, contigs, n_contigs, k, ref_skip, build_kmer, &s);
	free(s.fill);
	free(s.freq);
	return 0;
}

void cs_oracle_free_index(cs_oracle_index *ix) {
	free(ix->tab);
	free(ix->weight);
	free(ix->table);
	memset(ix, 0, sizeof(*ix));
}

/* ---- search ------------------------------------------------------------------------------------------- */
typedef struct {               /* CSTableEntry (LocationScore.h:9-14): the bin is kept in 32 bits, compared against 64 */
	uint32_t loc;
	uint32_t state;
	float fscore, rscore;
} cs_entry;

typedef struct {
	const cs_oracle_index *ix;
	int read_len;
	int max_kfreq;
	float sensitivity;
	cs_entry *rtable;
	uint32_t table_len;         /* power of two; large enough never to overflow (the reference's overflow retry, CS.cpp:386-430,
	                               only changes the table size, not the result) */
	int table_bits;
	uint32_t cur_state;
	uint32_t *rlist;
	int rlist_len;
	float max_hit, cur_thresh;
	/* bs-mapping / SLAMseq (CS::PrefixMutateSearch, CS.cpp:53-112) */
	int mutate_mode;            /* 0 none, 1 bs_mapping, 2 slam_seq & 4 */
	uint64_t mutate_from, mutate_to;
	int bs_cutoff;
	float weight;               /* 1 / m_CurrentMutLocs (CS.cpp:136-138) */
	/* the reference's overflow test (hpoc, CS.cpp:176-178): only when probe_budget_on */
	int probe_budget_on;
	uint32_t hpoc;
	int overflow;
	unsigned read_skip;         /* m_PrefixBaseSkip: "kmer_skip" under bs_mapping, else 0 (CS.cpp:556-560) */
} search_state;

static inline uint32_t cs_hash(uint64_t n, int bits) {             /* CS::Hash, CS.h:94-102 */
	return (uint32_t) ((n * 11400714819323199488ull) >> (64 - bits));
}

static void add_location(search_state *s, uint64_t loc, int reverse, float freq) {     /* CS::AddLocationStd, CS.cpp:164-213 */
	if (s->overflow) return;                                        /* (the exception has left PrefixIteration) */
	uint32_t e = cs_hash(loc, s->table_bits);
	int found;
	while ((found = ((s->rtable[e].state & 0x7FFFFFFFu) == s->cur_state)) && !((uint64_t) s->rtable[e].loc == loc)) {
		++e;
		if (e >= s->table_len) e = 0;
		if (s->probe_budget_on && --s->hpoc == 0) {                 /* throw 1, CS.cpp:176-178 */
			s->overflow = 1;
			return;
		}
	}
	cs_entry *en = &s->rtable[e];
	float score = freq;
	if (!found) {
		en->loc = (uint32_t) loc;
		en->state = s->cur_state & 0x7FFFFFFFu;
		if (reverse) {
			en->fscore = 0.0f;
			en->rscore = score;
		} else {
			en->fscore = score;
			en->rscore = 0.0f;
		}
	} else if (reverse) {
		score = (en->rscore += freq);
	} else {
		score = (en->fscore += freq);
	}
	if (score > s->max_hit) {
		s->max_hit = score;
		s->cur_thresh = s->max_hit * s->sensitivity;
	}
	if (!(en->state & 0x80000000u) && score >= s->cur_thresh) {
		en->state |= 0x80000000u;
		s->rlist[s->rlist_len++] = e;
	}
}

static void prefix_search(uint64_t prefix, uint64_t pos, void *data) {     /* CS::PrefixSearch, CS.cpp:114-162 + GetRefEntry */
	search_state *s = (search_state *) data;
	const cs_oracle_index *ix = s->ix;
	const uint32_t rc = cs_oracle_revcomp((uint32_t) prefix, ix->k);
	uint32_t fstart = 0, fcount = 0, rstart = 0, rcount = 0;
	if (ix->weight[prefix] != 0) {
		fstart = ix->tab[prefix] - 1;
		fcount = ix->tab[prefix + 1] - 1 - fstart;
	}
	if (ix->weight[rc] != 0) {
		rstart = ix->tab[rc] - 1;
		rcount = ix->tab[rc + 1] - 1 - rstart;
	}
	if (!((int) (fcount + rcount) < s->max_kfreq)) return;         /* cur->refTotal < maxPrefixFreq, CS.cpp:122 */
	for (uint32_t i = 0; i < fcount; ++i) {
		const uint64_t loc = ix->table[fstart + i];
		add_location(s, (loc - pos) >> ix->bin_shift, 0, s->weight);
	}
	const uint64_t corr = (uint64_t) s->read_len - (pos + (uint64_t) ix->k);
	for (uint32_t i = 0; i < rcount; ++i) {
		const uint64_t loc = ix->table[rstart + i];
		add_location(s, (loc - corr) >> ix->bin_shift, 1, s->weight);
	}
}

/* CS::PrefixMutateSearchEx (CS.cpp:95-112): the k-mer itself, then, depth first, every k-mer that has a subset of its
 * `mutate_from` bases replaced by `mutate_to` (positions counted from the k-mer's LAST base, i = 0) */
static void prefix_mutate_search_ex(search_state *s, uint64_t prefix, uint64_t pos, int mpos) {
	prefix_search(prefix, pos, s);
	for (int i = mpos; i < s->ix->k; ++i) {
		if (((prefix >> (2 * i)) & 3u) == s->mutate_from) {
			const uint64_t p = (prefix & ~((uint64_t) 3 << (2 * i))) | (s->mutate_to << (2 * i));
			prefix_mutate_search_ex(s, p, pos, i + 1);
		}
	}
}

static void prefix_mutate_search(uint64_t prefix, uint64_t pos, void *data) {     /* CS::PrefixMutateSearch, CS.cpp:53-80 */
	search_state *s = (search_state *) data;
	int locs = 0;
	for (int i = 0; i < s->ix->k; ++i) locs += ((prefix >> (2 * i)) & 3u) == s->mutate_from;
	if (s->mutate_mode == 2) {
		s->weight = 1.0f;                                           /* m_CurrentMutLocs = 1 */
		prefix_search(prefix, pos, s);
		s->weight = 1.0f / (float) (locs + 1);                      /* 1.0f / cs->m_CurrentMutLocs, CS.cpp:137 */
		for (int i = 0; i < s->ix->k; ++i) {                        /* PrefixMutateSearchSlamSeq, CS.cpp:82-93: single replacements only */
			if (((prefix >> (2 * i)) & 3u) == s->mutate_from)
				prefix_search((prefix & ~((uint64_t) 3 << (2 * i))) | (s->mutate_to << (2 * i)), pos, s);
		}
	} else if (locs <= s->bs_cutoff) {
		s->weight = 1.0f;
		prefix_mutate_search_ex(s, prefix, pos, 0);
	}
}

static int table_bits_for(int read_len, int max_kfreq) {
	/* every k-mer contributes < max_kfreq hits */
	const uint64_t worst = (uint64_t) (read_len > 0 ? read_len : 1) * (uint64_t) (max_kfreq > 1 ? max_kfreq : 1) * 4 + 64;
	int bits = 8;
	while (((uint64_t) 1 << bits) < worst && bits < 30) ++bits;
	return bits;
}

static void state_alloc(search_state *s, int bits) {
	s->table_bits = bits;
	s->table_len = (uint32_t) 1 << bits;
	s->rtable = (cs_entry *) malloc((size_t) s->table_len * sizeof(cs_entry));
	s->rlist = (uint32_t *) malloc((size_t) s->table_len * sizeof(uint32_t));
	for (uint32_t i = 0; i < s->table_len; ++i) {               /* CS::DoRun, CS.cpp:470-474 */
		s->rtable[i].loc = 0;                                   /* (uint) 2^63 */
		s->rtable[i].state = 0xFFFFFFFFu;
	}
	s->cur_state = 0;
}

static int search_one(search_state *s, const char *read, int read_len, float kmer_min, int max_cmrs, cs_oracle_cand *out, int out_cap,
		float *max_hit) {
	const cs_oracle_index *ix = s->ix;
	s->read_len = read_len;
	s->cur_state += 1;                                              /* RunBatch, CS.cpp:351-356 */
	s->rlist_len = 0;
	s->max_hit = 0.0f;
	s->cur_thresh = 0.0f;
	s->overflow = 0;
	s->weight = 1.0f;
	prefix_iteration(read, (uint64_t) read_len, s->mutate_mode ? prefix_mutate_search : prefix_search, s, s->mutate_mode == 1 ? s->read_skip : 0, 0,
			(unsigned) ix->k);
	if (s->overflow) return -1;
	/* CollectResultsStd, CS.cpp:263-313 */
	const float thr = kmer_min > s->cur_thresh ? kmer_min : s->cur_thresh;
	int index = 0;
	const uint64_t off = ix->bin_shift > 0 ? (uint64_t) 1 << (ix->bin_shift - 1) : 0;      /* ResolveBin, CS.h:170-175 */
	for (int i = 0; i < s->rlist_len; ++i) {
		const cs_entry t = s->rtable[s->rlist[i]];
		if (t.fscore >= thr) {
			if (index < out_cap) {
				out[index].location = ((uint64_t) t.loc << ix->bin_shift) + off;
				out[index].reverse = 0;
				out[index].score = t.fscore;
			}
			++index;
		}
		if (t.rscore >= thr) {
			if (index < out_cap) {
				out[index].location = ((uint64_t) t.loc << ix->bin_shift) + off;
				out[index].reverse = 1;
				out[index].score = t.rscore;
			}
			++index;
		}
	}
	if (max_hit) *max_hit = s->max_hit;
	if (!(index < max_cmrs)) return 0;                              /* :308-310: the read keeps no scores */
	return index;
}

int cs_oracle_search(const cs_oracle_index *ix, const char *read, int read_len, float sensitivity, float kmer_min, int max_kfreq, int max_cmrs,
		cs_oracle_cand *out, int out_cap, float *max_hit) {
	search_state s;
	memset(&s, 0, sizeof(s));
	s.ix = ix;
	s.max_kfreq = max_kfreq;
	s.sensitivity = sensitivity;
	state_alloc(&s, table_bits_for(read_len, max_kfreq));
	const int n = search_one(&s, read, read_len, kmer_min, max_cmrs, out, out_cap, max_hit);
	free(s.rtable);
	free(s.rlist);
	return n;
}

long long cs_oracle_search_batch(const cs_oracle_index *ix, const char *reads, int n_reads, int stride, float sensitivity, float kmer_min,
		int max_kfreq, int max_cmrs, int *cand_begin, cs_oracle_cand *out, long long out_cap, float *max_hit) {
	search_state s;
	memset(&s, 0, sizeof(s));
	s.ix = ix;
	s.max_kfreq = max_kfreq;
	s.sensitivity = sensitivity;
	state_alloc(&s, table_bits_for(stride, max_kfreq));
	long long total = 0;
	for (int r = 0; r < n_reads; ++r) {
		const char *read = reads + (size_t) r * stride;
		int len = 0;
		while (len < stride && read[len] != '\0') ++len;           /* MappedRead::length */
		cand_begin[r] = (int) total;
		const long long room = out_cap - total;
		const int n = search_one(&s, read, len, kmer_min, max_cmrs, out + total, room > 0x7fffffff ? 0x7fffffff : (room < 0 ? 0 : (int) room), max_hit ? max_hit + r : 0);
		total += n;
	}
	cand_begin[n_reads] = (int) total;
	free(s.rtable);
	free(s.rlist);
	return total;
}

/* CS::RunBatch with bs_mapping / slam_seq (CS.cpp:340-436): the mutated base by mode and mate, the search in a table of 2^table_bits
 * slots with a budget of 0.333 x slots probe steps, after an overflow again with table_bits + 2, + 3, ... <= 20 and 0.777 x slots; a read
 * that overflows every table keeps no candidates. */
static long long g_mut_retries = 0, g_mut_dropped = 0;
/* diagnostics of the last cs_oracle_search_batch_mut call: searches repeated in a larger table, reads that overflowed every table */
long long cs_oracle_last_retries(void) { return g_mut_retries; }
long long cs_oracle_last_dropped(void) { return g_mut_dropped; }

long long cs_oracle_search_batch_mut(const cs_oracle_index *ix, const char *reads, int n_reads, int stride, float sensitivity, float kmer_min,
		int max_kfreq, int max_cmrs, int mutate_mode, int bs_cutoff, int paired, int read_skip, int table_bits, int *cand_begin,
		cs_oracle_cand *out, long long out_cap, float *max_hit) {
	search_state s;
	g_mut_retries = g_mut_dropped = 0;
	memset(&s, 0, sizeof(s));
	s.ix = ix;
	s.max_kfreq = max_kfreq;
	s.sensitivity = sensitivity;
	s.mutate_mode = mutate_mode;
	s.bs_cutoff = bs_cutoff;
	s.read_skip = (unsigned) read_skip;
	s.probe_budget_on = 1;
	state_alloc(&s, 20);
	long long total = 0;
	for (int r = 0; r < n_reads; ++r) {
		const char *read = reads + (size_t) r * stride;
		int len = 0;
		while (len < stride && read[len] != '\0') ++len;
		cand_begin[r] = (int) total;
		const int second = paired && (r & 1);                       /* ReadId & 1, CS.cpp:362-380 */
		if (mutate_mode == 2) {
			s.mutate_from = second ? 3 : 1;
			s.mutate_to = second ? 0 : 2;
		} else {
			s.mutate_from = second ? 0 : 2;
			s.mutate_to = second ? 3 : 1;
		}
		const long long room = out_cap - total;
		const int cap = room > 0x7fffffff ? 0x7fffffff : (room < 0 ? 0 : (int) room);
		int n = -1, x = 2, bits = table_bits, tries = 0;
		while (n < 0 && bits <= 20) {
			s.table_bits = bits;
			s.table_len = (uint32_t) 1 << bits;
			s.hpoc = (uint32_t) ((float) (int) s.table_len * (tries == 0 ? 0.333f : 0.777f));      /* CS.cpp:392,418 */
			n = search_one(&s, read, len, kmer_min, max_cmrs, out + total, cap, max_hit ? max_hit + r : 0);
			bits = table_bits + x;
			x += 1;
			++tries;
		}
		g_mut_retries += tries - 1;
		if (n < 0) {
			n = 0;
			g_mut_dropped += 1;
		}
		total += n;
	}
	cand_begin[n_reads] = (int) total;
	free(s.rtable);
	free(s.rlist);
	return total;
}

/* ---- sensitivity estimate (ReadProvider::init) -------------------------------------------------------- */
typedef struct {
	const cs_oracle_index *ix;
	int read_len, max_kfreq;
	uint64_t *bins;             /* stands in for iTable (std::map<uloc, float>, ReadProvider.cpp:34): every vote has weight 1 */
	size_t n, cap;
} estimate_state;

static void estimate_vote(estimate_state *s, uint64_t bin) {
	if (s->n == s->cap) {
		s->cap = s->cap ? 2 * s->cap : 4096;
		s->bins = (uint64_t *) realloc(s->bins, s->cap * sizeof(uint64_t));
	}
	s->bins[s->n++] = bin;
}

static void estimate_prefix_search(uint64_t prefix, uint64_t pos, void *data) {        /* ReadProvider.cpp:81-123 (static PrefixSearch) */
	estimate_state *s = (estimate_state *) data;
	const cs_oracle_index *ix = s->ix;
	const uint32_t rc = cs_oracle_revcomp((uint32_t) prefix, ix->k);
	uint32_t fstart = 0, fcount = 0, rstart = 0, rcount = 0;
	if (ix->weight[prefix] != 0) {
		fstart = ix->tab[prefix] - 1;
		fcount = ix->tab[prefix + 1] - 1 - fstart;
	}
	if (ix->weight[rc] != 0) {
		rstart = ix->tab[rc] - 1;
		rcount = ix->tab[rc + 1] - 1 - rstart;
	}
	if (!((int) (fcount + rcount) < s->max_kfreq)) return;         /* :87 */
	for (uint32_t i = 0; i < fcount; ++i) estimate_vote(s, ((uint64_t) ix->table[fstart + i] - pos) >> ix->bin_shift);        /* :111-112 */
	const uint64_t corr = (uint64_t) s->read_len - (pos + (uint64_t) ix->k);                                                   /* :97-100 */
	for (uint32_t i = 0; i < rcount; ++i) estimate_vote(s, ((uint64_t) ix->table[rstart + i] - corr) >> ix->bin_shift);
}

static int cmp_u64(const void *a, const void *b) {
	const uint64_t x = *(const uint64_t *) a, y = *(const uint64_t *) b;
	return x < y ? -1 : (x > y ? 1 : 0);
}

/* ReadProvider::init (ReadProvider.cpp:236-251,310-325) + CollectResultsFallback (:53-79): `sampled` are the reads number 1000, 2000, ...
 * of the input (the caller samples).  Both strands vote into ONE map keyed by the bin, so the best vote here is not MappedRead::s. */
int cs_oracle_estimate_sensitivity(const cs_oracle_index *ix, const char *sampled, int n_reads, int stride, int max_kfreq, float *sensitivity) {
	estimate_state s;
	memset(&s, 0, sizeof(s));
	s.ix = ix;
	s.max_kfreq = max_kfreq;
	const int skip = ix->ref_skip + 1;
	float sum = 0.0f;
	int count = 0;
	for (int r = 0; r < n_reads; ++r) {
		const char *read = sampled + (size_t) r * stride;
		int len = 0;
		while (len < stride && read[len] != '\0') ++len;
		s.read_len = len;
		s.n = 0;
		prefix_iteration(read, (uint64_t) len, estimate_prefix_search, &s, 0, 0, (unsigned) ix->k);
		float max_current = 0.0f;
		if (s.n > 0) {
			qsort(s.bins, s.n, sizeof(uint64_t), cmp_u64);
			size_t run = 1;
			for (size_t i = 1; i <= s.n; ++i) {
				if (i < s.n && s.bins[i] == s.bins[i - 1]) {
					++run;
					continue;
				}
				if ((float) run > max_current) max_current = (float) run;
				run = 1;
			}
		}
		const int max = (int) ceil((double) ((len - ix->k + 1) / skip) * 1.0);         /* :68: integer division, then ceil of an int */
		if ((float) max > 1.0f && max_current <= (float) max) {
			sum += max_current / (float) max;
			count += 1;
		}
	}
	free(s.bins);
	const float avg = sum / (float) count * 1.0f;                   /* :319; 0/0 when nothing qualified */
	const float lo = 0.3f > avg ? 0.3f : avg;                       /* std::max(0.3f, avg): NaN avg -> 0.3f?  std::max(a,b) = (a<b)?b:a */
	*sensitivity = lo < 0.9f ? lo : 0.9f;                           /* std::min(x, 0.9f) = (0.9f<x)?0.9f:x */
	return count;
}
