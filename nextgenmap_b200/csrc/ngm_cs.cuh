// ngm_cs.cuh -- candidate search on the device: NextGenMap's k-mer index ("prefix table") and
// the per-read vote that turns k-mer hits into candidate reference windows for BatchScore.
//
// Reference: src/CSstatic.cpp:20-76 (k-mer iteration), src/PrefixTable.cpp:357-498,577-738 (index construction),
// :750-817 (lookup), src/CS.cpp:114-213 (PrefixSearch + AddLocationStd), :263-313 (CollectResultsStd),
// src/CS.h:164-175 (bins).  The results -- candidate set, votes AND the order of every read's list (which decides
// ties in ScoreBuffer::top1SE) -- are identical to the reference's.
//
// Index in HBM:  tabu[p] = (exclusive prefix sum of the k-mer counts) | used(p) << 31 for p in [0, 4^k], and
// table[] = reference positions grouped by k-mer, ascending inside a group (the order the reference's sequential
// fill produces).  3 Gbp, k 13, every 3rd position: 268 MB + 4.1 GB.
//
// Search: one 256-thread block per read (details above cs_search_kernel): every k-mer of the read is looked up
// (forward and reverse-complement list) and every hit votes for bin (position - offset) >> bin_size; votes are
// counted exactly in shared memory behind a "seen" bitmap that filters the single-vote noise.  The reference's result only
// depends on the final votes, except for the ORDER of a read's candidates, which is the order in which entries first
// reached the running threshold sensitivity x max-votes-so-far while hits arrive in k-mer order.  That order is
// reconstructed exactly for the few reads that need it (more than one accepted entry): the hits of all entries that
// can influence the running maximum (>= 2 votes) or are accepted are collected with their sequence number, sorted,
// and replayed.  Reads that do not fit the fast path (table overflow, too many relevant hits, > 64 accepted entries,
// 64-bit bin arithmetic) go to cs_search_exact_kernel, a one-lane sequential restatement on a table in global memory.
#pragma once

#include "ngm_common.cuh"

// index / position-table loads of the search kernels: random, no reuse inside a block
#ifndef NGM_CS_LD
#define NGM_CS_LD(p) __ldg(p)          /* ld.global.cg measured 16 % slower: the 18 % L1 hits are worth keeping */
#endif

#ifndef NGM_CS_MIN_BLOCKS
#define NGM_CS_MIN_BLOCKS 5            /* resident blocks the register allocation must allow (5 x 256 threads: <= 51 registers) */
#endif

namespace ngm {

struct CsDev {
	const uint32_t *tabu;     // [4^k + 1]
	const uint4 *both;        // [4^k] {forward list start, length, reverse-complement list start, length} in `table`: one 16-byte load per k-mer
	                          // (`table` here is the search's pair-ordered copy, see cs_pair_copy_kernel)
	const uint32_t *table;
	int k, bin_shift, max_kfreq, max_cmrs;
	float sensitivity, kmer_min;
	int merged;               // max_hit[] receives the best vote with both strands added (ReadProvider's estimate) instead of MappedRead::s
	// bs-mapping / SLAMseq k-mer mutation (CS::PrefixMutateSearch, CS.cpp:53-112); served by cs_search_exact_kernel only
	int mut_mode;             // 0 off, 1 "bs_mapping", 2 "slam_seq" & 4
	int mut_cutoff;           // "bs_cutoff" (6): read k-mers with more replaceable bases are skipped (bs_mapping)
	int mut_paired;           // "paired": odd reads are second mates and mutate the complementary base (CS.cpp:362-380)
	int read_skip;            // m_PrefixBaseSkip: "kmer_skip" under bs_mapping, else 0 (CS.cpp:556-560)
	int ex_bits;              // log2 slots of the exact kernel's table (kCsExactBits; 20 with mutation = the largest table CS::RunBatch tries)
	float *heap_votes;        // mut_mode 2: the candidates' fractional votes, parallel to the heap (CsCand::votes holds integers)
};

struct CsRun {                // one N-free stretch of a contig as CS::PrefixIteration walks it
	uint64_t start;           // concatenated position of the first emitted k-mer
	uint64_t tail_start;      // positions >= tail_start read as code 0 (the two undecoded bases at a contig's end)
	uint64_t emit_base;       // emission index of this run's first k-mer
	uint64_t contig_base;     // emission index of the contig's first k-mer
	uint32_t n_emit;
	uint32_t pad;
};

struct CsMeta {               // per read: where its candidates sit in the heap
	uint32_t off;
	uint32_t count;           // kPending while the read waits for the exact kernel
};

struct CsCand {               // 8 bytes
	uint32_t bin;
	uint16_t votes;
	uint16_t rev;
};

constexpr uint32_t kCsPending = 0xFFFFFFFFu;
constexpr uint32_t kCsEmpty = 0xFFFFFFFFu;
constexpr int kCsMaxItems = 512;           // relevant hits replayed for the order
constexpr int kCsMaxAccepted = 192;        // noise bins collect two votes at a rate of ~1.5 % of the hits (a random 13-mer hit extends to the next indexed position with p = 1/64): ~100 per 250 bp read, all accepted when the best bin has <= 4 votes
constexpr int kCsMaxStride = 1024;         // NGM's maximum read length (ReadProvider.cpp:42)
constexpr int kCsExactBits = 17;           // slots of the exact kernel's table (reads <= 1000 bp x max_kfreq hits)

__host__ __device__ __forceinline__ uint32_t cs_revcomp(uint32_t prefix, int k) {      // PrefixTable.cpp:93-108
	uint32_t c = (prefix ^ 0xAAAAAAAAu) << (32 - 2 * k);
	c = (c & 0xFFFF0000u) >> 16 | (c & 0x0000FFFFu) << 16;
	c = (c & 0xFF00FF00u) >> 8 | (c & 0x00FF00FFu) << 8;
	c = (c & 0xF0F0F0F0u) >> 4 | (c & 0x0F0F0F0Fu) << 4;
	c = (c & 0xCCCCCCCCu) >> 2 | (c & 0x33333333u) << 2;
	return c;
}

// ---------------------------------------------------------------------------------------------------------
// index construction
// ---------------------------------------------------------------------------------------------------------
// boundaries of the N runs of the whole concatenated reference (code 5 in the device packing): entry = pos << 1 | is_end
__global__ void cs_find_n_kernel(const uint32_t *__restrict__ ref4, uint64_t n_words, uint64_t concat_len, unsigned long long *__restrict__ list,
		uint32_t cap, uint32_t *__restrict__ count) {
	const uint64_t wi = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (wi >= n_words) return;
	const uint32_t w = ref4[wi];
	// nibble == 5  <=>  (nibble ^ 5) == 0
	const uint32_t x = w ^ 0x55555555u;
	const uint32_t z = ~(x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u;     // bit 4i set iff nibble i is N
	if (z == 0) return;
	uint32_t nm = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) nm |= ((z >> (4 * i)) & 1u) << i;
	for (int i = 0; i < 8; ++i)
		if (wi * 8 + i >= concat_len) nm &= ~(1u << i);
	if (nm == 0) return;
	const uint32_t prev = wi > 0 ? ((ref4[wi - 1] >> 28) == 5u) : 0u;
	const uint32_t next = (wi + 1 < n_words && (wi + 1) * 8 < concat_len) ? ((ref4[wi + 1] & 0xFu) == 5u) : 0u;
	const uint32_t starts = nm & ~((nm << 1) | prev);
	const uint32_t ends = nm & ~((nm >> 1) | (next << 7));
	for (int i = 0; i < 8; ++i) {
		if ((starts >> i) & 1u) {
			const uint32_t at = atomicAdd(count, 1u);
			if (at < cap) list[at] = (unsigned long long) (wi * 8 + i) << 1;
		}
		if ((ends >> i) & 1u) {
			const uint32_t at = atomicAdd(count, 1u);
			if (at < cap) list[at] = ((unsigned long long) (wi * 8 + i) << 1) | 1ull;
		}
	}
}

__device__ __forceinline__ int cs_find_run(const CsRun *__restrict__ runs, int n_runs, uint64_t e) {
	int lo = 0, hi = n_runs - 1;
	while (lo < hi) {                                      // last run with emit_base <= e
		const int mid = (lo + hi + 1) >> 1;
		if (runs[mid].emit_base <= e) lo = mid; else hi = mid - 1;
	}
	return lo;
}

// k-mer code (A0 C1 T2 G3, CSstatic.cpp:19-22) of the k bases at concatenated position x
__device__ __forceinline__ uint32_t cs_ref_kmer(const uint32_t *__restrict__ ref4, uint64_t x, int k, uint64_t tail_start) {
	uint32_t p = 0;
	uint64_t wi = x >> 3;
	int sh = 4 * (int) (x & 7);
	uint32_t w = ref4[wi];
	for (int i = 0; i < k; ++i) {
		uint32_t c = (w >> sh) & 0xFu;
		if (x + i >= tail_start) c = 0;
		p = (p << 2) | ((c ^ (c >> 1)) & 3u);
		sh += 4;
		if (sh == 32) {
			sh = 0;
			w = ref4[++wi];
		}
	}
	return p;
}

// One thread per emitted reference k-mer: CountKmer's de-duplication (PrefixTable.cpp:632-655) written as a local rule,
//   counted(e) = !same(e) || !same(e-1) || bin(x_e) != bin(x_{e-1}),  same(e) = prefix(e) == prefix(e-1),
// with prefix(-1) = 111111 and same(-1) = false at the start of every contig (PrefixTable.cpp:363-365).
__global__ void cs_emit_kernel(const uint32_t *__restrict__ ref4, const CsRun *__restrict__ runs, int n_runs, uint64_t n_emit, int k, int step,
		int bin_shift, int skip_rep, uint32_t sentinel, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t *__restrict__ freq) {
	const uint64_t e = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n_emit) return;
	const int r = cs_find_run(runs, n_runs, e);
	const CsRun run = runs[r];
	const uint64_t x = run.start + (e - run.emit_base) * (uint64_t) step;
	const uint32_t p = cs_ref_kmer(ref4, x, k, run.tail_start);
	bool counted = true;
	if (skip_rep) {
		uint32_t p1 = 111111u, p2 = 111111u;
		uint64_t x1 = 0;
		bool have1 = e > run.contig_base, have2 = e > run.contig_base + 1;
		if (have1) {
			const int r1 = (e - 1 >= run.emit_base) ? r : cs_find_run(runs, n_runs, e - 1);
			const CsRun q = runs[r1];
			x1 = q.start + (e - 1 - q.emit_base) * (uint64_t) step;
			p1 = cs_ref_kmer(ref4, x1, k, q.tail_start);
		}
		if (have2) {
			const int r2 = (e - 2 >= run.emit_base) ? r : cs_find_run(runs, n_runs, e - 2);
			const CsRun q = runs[r2];
			p2 = cs_ref_kmer(ref4, q.start + (e - 2 - q.emit_base) * (uint64_t) step, k, q.tail_start);
		}
		const bool same0 = p == p1;
		const bool same1 = have1 && p1 == p2;
		counted = !same0 || !same1 || ((x >> bin_shift) != (x1 >> bin_shift));
	}
	if (counted) atomicAdd(&freq[p], 1u);
	keys[e] = counted ? p : sentinel;
	vals[e] = (uint32_t) x;
}

// createRefTableIndex (PrefixTable.cpp:436-498): weight / used flag per k-mer, and the sums stats() needs (:151-194)
__global__ void cs_weight_kernel(const uint32_t *__restrict__ freq, uint32_t n_prefix, int k, int8_t *__restrict__ weight,
		unsigned long long *__restrict__ sums) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long s1 = 0, s2 = 0;
	if (i < n_prefix) {
		const uint32_t f = freq[i];
		int8_t w = 0;
		if (f > 0) {
			const uint32_t tot = f + freq[cs_revcomp(i, k)];
			const int dummy = 10000;
			w = (int8_t) (int) ((float) (dummy - (int) min(tot, (uint32_t) dummy)) * 100.0f / (float) dummy);
		}
		weight[i] = w;
		s1 = f;
		s2 = (unsigned long long) f * f;
	}
	for (int d = 16; d > 0; d >>= 1) {
		s1 += __shfl_down_sync(0xffffffffu, s1, d);
		s2 += __shfl_down_sync(0xffffffffu, s2, d);
	}
	if ((threadIdx.x & 31) == 0 && (s1 | s2)) {
		atomicAdd(&sums[0], s1);
		atomicAdd(&sums[1], s2);
	}
}

__global__ void cs_tabu_kernel(const uint32_t *__restrict__ off, const int8_t *__restrict__ weight, uint32_t n_prefix, uint32_t *__restrict__ tabu) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > n_prefix) return;
	tabu[i] = off[i] | ((i < n_prefix && weight[i] != 0) ? 0x80000000u : 0u);
}

// k-mers that were counted but are not "used" (>= 9901 occurrences incl. reverse complement) keep zeroed slots
// (BuildPrefixTable only stores for used() k-mers, PrefixTable.cpp:677-690)
__global__ void cs_zero_unused_kernel(const uint32_t *__restrict__ tabu, uint32_t n_prefix, uint32_t *__restrict__ table) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_prefix) return;
	const uint32_t a = tabu[i];
	if (a >> 31) return;
	const uint32_t b = tabu[i + 1] & 0x7FFFFFFFu;
	for (uint32_t j = a; j < b; ++j) table[j] = 0;
}

// index loaded from a `<ref>-ht-<k>-<skip>.3.ngm` file (Index::m_TabIndex is 1-based)
__global__ void cs_tabu_from_file_kernel(const uint32_t *__restrict__ tab, const int8_t *__restrict__ weight, uint32_t n_prefix, uint32_t *__restrict__ tabu) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > n_prefix) return;
	tabu[i] = (tab[i] - 1u) | ((i < n_prefix && weight[i] != 0) ? 0x80000000u : 0u);
}

__global__ void cs_export_tab_kernel(const uint32_t *__restrict__ tabu, uint32_t n_prefix, uint32_t *__restrict__ tab) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > n_prefix) return;
	tab[i] = (tabu[i] & 0x7FFFFFFFu) + 1u;
}

// ---------------------------------------------------------------------------------------------------------
// search: shared helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cs_enc2(uint8_t c) { return ((uint32_t) c >> 1) & 3u; }      // CSstatic.cpp:19-22

// Is the k-mer starting at o one that CS::PrefixIteration emits?  No 'N' inside, and not the k-mer the N-skip path
// drops: the last k-mer of the read when it directly follows an N run entered through CSstatic.cpp:30-41.
__device__ __forceinline__ bool cs_read_kmer(const uint8_t *seq, int len, int o, int k, uint32_t &prefix) {
	uint32_t p = 0;
	bool ok = true;
	for (int i = 0; i < k; ++i) {
		const uint8_t c = seq[o + i];
		ok = ok && (c != 'N');
		p = (p << 2) | cs_enc2(c);
	}
	if (ok && o + k == len && o >= 1 && seq[o - 1] == 'N' && (o == 1 || seq[o - 2] == 'N')) ok = false;
	prefix = p;
	return ok;
}

struct CsLists {
	uint32_t fs, fc, rs, rc;
};

__device__ __forceinline__ bool cs_lookup(const CsDev &P, uint32_t prefix, CsLists &L) {      // GetRefEntry, PrefixTable.cpp:750-817
	const uint4 e = NGM_CS_LD(P.both + prefix);
	L.fs = e.x;
	L.fc = e.y;
	L.rs = e.z;
	L.rc = e.w;
	return (int) (L.fc + L.rc) < P.max_kfreq;              // cur->refTotal < maxPrefixFreq, CS.cpp:122
}

// The search's own copy of the index.  GetRefEntry assembles a k-mer's two lists from Index[p], Index[p + 1], Index[rc(p)], Index[rc(p) + 1]
// (two random 8-byte reads per k-mer of every read) and the lists of p and rc(p) lie wherever their prefixes put them.  Laid down once per
// index: table2 holds, for every pair {p, rc(p)}, the list of the smaller prefix followed by the list of the larger one, and both[p] =
// {start of p's list, length, start of rc(p)'s list, length} in table2 -- one aligned 16-byte entry and ONE contiguous range of positions
// per k-mer of a read.  tabu / table keep the reference's layout (export, file format).
__device__ __forceinline__ uint32_t cs_list_len(const uint32_t *__restrict__ tabu, uint32_t p) {
	const uint32_t a0 = tabu[p];
	return (a0 >> 31) ? (tabu[p + 1] & 0x7FFFFFFFu) - (a0 & 0x7FFFFFFFu) : 0u;
}

__global__ void cs_pair_size_kernel(const uint32_t *__restrict__ tabu, uint32_t n_prefix, int k, uint32_t *__restrict__ size) {
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n_prefix) return;
	const uint32_t c = cs_revcomp(p, k);
	const uint32_t n = p <= c ? cs_list_len(tabu, p) + (c != p ? cs_list_len(tabu, c) : 0u) : 0u;
	// every pair region starts on a 64-byte boundary (16 positions): a region then costs exactly ceil(bytes / 64) DRAM fetches of 64 bytes
	// instead of straddling one more, and it is a legal source of a bulk copy (cp.async.bulk: 16-byte aligned)
	size[p] = (n + 15u) & ~15u;
}

__global__ void cs_pair_copy_kernel(const uint32_t *__restrict__ tabu, uint32_t n_prefix, int k, const uint32_t *__restrict__ off2,
		const uint32_t *__restrict__ table, uint32_t *__restrict__ table2, uint4 *__restrict__ both) {
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n_prefix) return;
	const uint32_t c = cs_revcomp(p, k);
	const uint32_t n = cs_list_len(tabu, p), nc = cs_list_len(tabu, c);
	uint32_t dest_p, dest_c;
	if (p <= c) {
		dest_p = off2[p];
		dest_c = c != p ? dest_p + n : dest_p;
	} else {
		dest_c = off2[c];
		dest_p = dest_c + nc;
	}
	const uint32_t src = tabu[p] & 0x7FFFFFFFu;
	for (uint32_t i = 0; i < n; ++i) table2[dest_p + i] = table[src + i];
	both[p] = make_uint4(dest_p, n, dest_c, nc);
}

// ---------------------------------------------------------------------------------------------------------
// TMA (bulk asynchronous copy) helpers: cp.async.bulk global -> shared, completion on an mbarrier (SASS: UBLKCP + SYNCS)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cs_smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void cs_mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cs_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void cs_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cs_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool cs_mbar_try_wait(uint64_t *bar, uint32_t parity) {
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(ok) : "r"(cs_smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void cs_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			::"r"(cs_smem_u32(dst)), "l"(src), "r"(bytes), "r"(cs_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cs_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------
// search: fast path, one 256-thread block per read
//
// TMA variant (template parameter TMA, the default on sm_100a): a k-mer's two position lists are ONE 64-byte aligned region of `table`,
// so the thread that looked the k-mer up issues one bulk copy of that region straight into the k-mer's slots of the shared hit array;
// all copies of the read complete on one mbarrier.  Sweep B then walks the hit array k-mer by k-mer (one warp per k-mer: offset and
// strand are warp-uniform, no hit -> k-mer map, no per-lane global address arithmetic) and turns positions into bins in place.  Slots keep
// the table's layout order; the reference's hit order, which only phase E and path J need, is recovered per slot by cs_seq_of().
//
// Almost all hits of a read are noise: bins that are hit exactly once (~4100 of ~4150 at 150 bp on 3 Gbp).  Only
// bins hit at least twice can reach a threshold above one vote, so the hits are first run through a 64 Kbit
// "seen" bitmap (one ATOMS.OR each, no probing); only hits that find their bit already set (true repeats plus ~3 %
// false positives) enter a small exact table (double hashing, load < 0.2).  A second sweep over the hits -- kept
// in shared memory as 32-bit bins -- adds the one hit per repeated bin that set the bit and was therefore not counted.
//
//   A  every thread takes one k-mer of the read: validity, code, ONE 16-byte index entry (both strands' lists); block-wide scan over
//      (list length | "has hits" << 20) = number of every hit in the reference's order + rank of every k-mer that has hits;
//      km[rank] = {first hit number, list starts, forward length | k-mer}, rbase / bstart = chunk -> rank maps
//   B  the hits are swept in flat chunks of 32 consecutive hit numbers, eight chunks in flight per warp: a lane's k-mer is
//      km[rbase[chunk] + popc(bstart[chunk] & lanes_below)]; position load, bitmap test, bins[] <- hit; the hits that found their bit
//      set are queued (one slot reservation per warp and batch)
//   B2 the queue enters the exact table, one hit per thread
//   C  the vote that set the bit is missing from every table entry: looked up only for the entries that can still reach the
//      threshold (<= kCsMaxRes bins in registers), or for all of them where that is needed (phase E, path J, sensitivity estimate)
//   D  maximum, threshold, accepted entries
//   E  only if more than one entry is accepted: the order of the list (see the header of this file)
//   J  threshold <= 1 (every bin is a candidate): the bins in first-hit order
// Reads with more hits than MAXH, a crowded queue / table, a zero threshold or a 64-bit wrap-around go to the exact kernel.
// ---------------------------------------------------------------------------------------------------------
enum CsExactReason { kCsWhyHits = 0, kCsWhyWrap, kCsWhyQueue, kCsWhyTable, kCsWhyMulti, kCsWhyZeroThr, kCsWhyAccepted, kCsWhyItems, kCsWhyOrder, kCsWhyCount };
constexpr int kCsRepWords = 512;                               // 16384 "repeated" bits; the "seen" bitmap has max(65536, 32 x T2) bits
constexpr int kCsMaxMulti = 256;                               // (entry, strand) pairs with two or more votes
constexpr int kCsMaxRes = 4;                                   // entries that can still reach the threshold after the queue is counted
constexpr int kCsQueue = 512;                                  // hits per read that find their bit set (repeats + false positives)
constexpr uint32_t kHitInserted = 0x80000000u, kHitRev = 0x40000000u, kHitBin = 0x3FFFFFFFu;

template <int T2_LOG, int MAXK, int MAXH, int SEENW_ = 0>
struct CsSmem {
	static constexpr int T2 = 1 << T2_LOG;
	static constexpr int SEENW = SEENW_ ? SEENW_ : (T2 > 2048 ? T2 : 2048);        // words of the "seen" bitmap (reused as per-slot first-hit array, >= T2)
	static_assert(SEENW >= T2, "the first-hit array of path J lives in the bitmap");
	static constexpr size_t bytes_staged = (size_t) (SEENW + kCsRepWords) * 4 + (size_t) T2 * 8 + (size_t) MAXH * 4 + (size_t) MAXK * 16 + (size_t) MAXK + 32;
	static constexpr size_t bytes = bytes_staged + ((size_t) MAXH / 32 + 2) * 2 + ((size_t) MAXH / 32 + 1) * 4;      // + the chunk maps of the first-generation sweep
};

constexpr uint32_t kHitPad = 0xFFFFFFFFu;                       // TMA variant: padding slot behind a k-mer's hits (regions are multiples of four slots)

// TMA variant: km[rank] = {first slot, region start in `table`, first list's length | hits << 16, k-mer offset | palindrome << 30 | first list is the reverse one << 31}
__device__ __forceinline__ int cs_rank_of(const uint4 *km, int n_km, uint32_t slot) {
	int lo = 0, hi = n_km - 1;
	while (lo < hi) {                                           // last k-mer whose first slot is <= slot
		const int mid = (lo + hi + 1) >> 1;
		if (km[mid].x <= slot) lo = mid; else hi = mid - 1;
	}
	return lo;
}
// slot (layout order: the list of the smaller prefix first) <-> number of the hit in the reference's order (forward list first)
__device__ __forceinline__ uint32_t cs_seq_of(const uint4 *km, int n_km, uint32_t slot) {
	const uint4 e = km[cs_rank_of(km, n_km, slot)];
	if (!(e.w >> 31)) return slot;
	const uint32_t o = slot - e.x, first = e.z & 0xFFFFu, total = e.z >> 16;
	return o < first ? e.x + (total - first) + o : e.x + (o - first);
}
__device__ __forceinline__ uint32_t cs_slot_of(const uint4 *km, int n_km, uint32_t seq) {
	const uint4 e = km[cs_rank_of(km, n_km, seq)];
	if (!(e.w >> 31)) return seq;
	const uint32_t o = seq - e.x, first = e.z & 0xFFFFu, total = e.z >> 16, fwd = total - first;
	return o < fwd ? e.x + first + o : (o < total ? e.x + (o - fwd) : seq);      // (padding slots map to themselves)
}

template <int T2_LOG, int MAXK, int MAXH, int SEENW_ = 0, int STAGE = 0>
__global__ void __launch_bounds__(256, NGM_CS_MIN_BLOCKS) cs_search_kernel(const CsDev P, const uint8_t *__restrict__ reads, int n_reads, int stride,
		CsMeta *__restrict__ meta, CsCand *__restrict__ heap, uint32_t heap_cap, uint32_t *__restrict__ cursor, uint32_t *__restrict__ slow_list,
		uint32_t *__restrict__ slow_count, float *__restrict__ max_hit) {
	// slow_count[1 + reason]: why reads left the fast path (diagnostics, see CsExactReason)
	// STAGE: how the position lists reach the shared hit array -- 0: per-lane loads in sweep B (first generation); 1: one bulk copy
	// (cp.async.bulk, UBLKCP) per k-mer region, completion on an mbarrier; 2: 16-byte cp.async (LDGSTS) issued by the k-mer's own thread
	constexpr bool TMA = STAGE != 0;
	constexpr int T2 = 1 << T2_LOG;
	constexpr uint32_t MASK = T2 - 1;
	constexpr int NT = 256;
	constexpr int IPT = MAXK / NT;                             // list descriptors per thread in the scan
	static_assert(MAXK % NT == 0, "MAXK must be a multiple of the block size");
	extern __shared__ uint32_t s_dyn[];
	constexpr int SEENW = CsSmem<T2_LOG, MAXK, MAXH, SEENW_>::SEENW;
	constexpr int SEEN_SHIFT = SEENW == 2048 ? 16 : (SEENW == 4096 ? 15 : 14);      // 32 - log2(32 x SEENW)
	static_assert(SEENW == 2048 || SEENW == 4096 || SEENW == 8192, "seen bitmap size");
	uint32_t *seen = s_dyn, *rep = seen + SEENW, *keys = rep + kCsRepWords, *cnts = keys + T2, *bins = cnts + T2;
	// the k-mers that have hits, by rank: {number of the k-mer's first hit, forward list start, reverse list start, forward length | k-mer << 16}
	uint4 *km = reinterpret_cast<uint4 *>(bins + MAXH);
	uint8_t *s_read = reinterpret_cast<uint8_t *>(km + MAXK);             // MAXK + 32 bytes >= stride (stride - k + 1 <= MAXK, k <= 14)
	// hit -> k-mer map of sweep B: per chunk c of 32 hits, rbase[c] = rank of the k-mer that holds hit 32 c and bstart[c] = the hits of
	// the chunk that begin a k-mer (bit 0 left out: that is rbase's)
	uint16_t *rbase = reinterpret_cast<uint16_t *>(s_read + MAXK + 32);
	uint32_t *bstart = reinterpret_cast<uint32_t *>(rbase + (MAXH / 32 + 2));
	static_assert(MAXK % 4 == 0 && MAXH % 64 == 0 && MAXK < 4096 && MAXH < (1 << 20), "alignment of the chunk maps; packing of the scan");
	// dead while the "seen" bitmap is in use (sweep B), needed only from phase D on: they live in the bitmap's words
	uint32_t *s_items_s = seen;                                           // kCsMaxItems words
	uint16_t *s_items_c = reinterpret_cast<uint16_t *>(seen + kCsMaxItems);      // kCsMaxItems halves
	uint32_t *s_acc = seen + kCsMaxItems + kCsMaxItems / 2;               // kCsMaxAccepted words
	uint32_t *s_items_t = s_acc + kCsMaxAccepted;                         // kCsMaxItems words
	static_assert(2 * kCsMaxItems + kCsMaxItems / 2 + kCsMaxAccepted <= SEENW, "phase D/E scratch must fit into the seen bitmap");
	__shared__ uint32_t s_warp[NT / 32];
	__shared__ int s_len;
	__shared__ uint16_t s_queue[kCsQueue];                     // hit numbers (< MAXH)
	static_assert(MAXH <= 65536, "queue entries are 16 bits wide");
	__shared__ uint16_t s_multi[kCsMaxMulti];
	__shared__ uint32_t s_nhits, s_nkm, s_tx, s_npalin;
	__shared__ __align__(8) uint64_t s_mbar;
	__shared__ uint32_t s_max, s_maxm, s_slow, s_nmulti, s_nacc, s_ncand, s_nitems, s_nord, s_nq, s_nres;
	__shared__ uint32_t s_res[2 * kCsMaxRes];                  // bins whose first vote is looked for (see sweep C)
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int r = blockIdx.x;
	if (r >= n_reads) return;
	if (tid == 0) {
		s_len = stride;
		s_max = s_maxm = s_slow = s_nmulti = s_nacc = s_ncand = s_nitems = s_nord = s_nq = s_nres = 0;
		s_tx = s_npalin = 0;
		if (STAGE == 1) {
			cs_mbar_init(&s_mbar, 1);
			cs_fence_proxy_async();
		}
	}
	{
		uint4 *z4 = reinterpret_cast<uint4 *>(seen);               // both bitmaps
		for (int i = tid; i < (SEENW + kCsRepWords) / 4; i += NT) z4[i] = make_uint4(0, 0, 0, 0);
		if (!TMA)
			for (int i = tid; i < MAXH / 32 + 1; i += NT) bstart[i] = 0;
		uint4 *k4 = reinterpret_cast<uint4 *>(keys), *c4 = reinterpret_cast<uint4 *>(cnts);
		for (int i = tid; i < T2 / 4; i += NT) {
			k4[i] = make_uint4(kCsEmpty, kCsEmpty, kCsEmpty, kCsEmpty);
			c4[i] = make_uint4(0, 0, 0, 0);
		}
	}
	const uint8_t *src = reads + (size_t) r * stride;
	__syncthreads();
	for (int i = tid; i < stride; i += NT) {
		const uint8_t c = src[i];
		s_read[i] = c;
		if (c == 0) atomicMin(&s_len, i);                  // MappedRead::length
	}
	__syncthreads();
	const int len = s_len;
	const int k = P.k;
	const int n_kmers = max(0, len - k + 1);                   // <= MAXK (checked by the launcher: stride - k + 1 <= MAXK)

	// ---- A: list descriptors + sequence numbers -----------------------------------------------------------
	uint32_t mine[IPT], m_fs[IPT], m_rs[IPT], m_fc[IPT], m_rc[IPT];
#pragma unroll
	for (int q = 0; q < IPT; ++q) {
		const int o = tid * IPT + q;                           // blocked layout so that the scan below is a plain prefix sum
		uint32_t fc = 0, rc = 0, fs = 0, rs = 0;
		if (o < n_kmers) {
			uint32_t prefix;
			if (cs_read_kmer(s_read, len, o, k, prefix)) {
				CsLists L;
				if (cs_lookup(P, prefix, L)) {
					fc = L.fc;
					rc = L.rc;
					fs = L.fs;
					rs = L.rs;
				}
			}
		}
		m_fs[q] = fs;
		m_rs[q] = rs;
		m_fc[q] = fc;
		m_rc[q] = rc;
		// hits in bits 0..19 (TMA: slots, a multiple of four = 16 bytes, the granularity of a bulk copy), "has hits" counted in bits 20..31
		mine[q] = (TMA ? ((fc + rc + 3u) & ~3u) : (fc + rc)) | ((fc + rc) ? (1u << 20) : 0u);
	}
	uint32_t tsum = 0;
#pragma unroll
	for (int q = 0; q < IPT; ++q) tsum += mine[q];
	uint32_t incl = tsum;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) incl += v;
	}
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	uint32_t wbase = 0;
#pragma unroll
	for (int w = 0; w < NT / 32; ++w) wbase += (w < warp) ? s_warp[w] : 0u;
	uint32_t run = wbase + incl - tsum;
#pragma unroll
	for (int q = 0; q < IPT; ++q) {
		const uint32_t start = run & 0xFFFFFu, cnt = mine[q] & 0xFFFFFu, rank = run >> 20;
		if (cnt && start + cnt <= (uint32_t) MAXH) {            // (reads with more hits leave for the exact kernel below)
			if (TMA) {
				const uint32_t fc = m_fc[q], rc = m_rc[q], fs = m_fs[q], rs = m_rs[q];
				const bool palin = fc && rc && fs == rs;            // the k-mer is its own reverse complement: one list, hit on both strands
				const bool first_rev = rc && (!fc || rs < fs);      // layout: the list of the smaller prefix comes first
				const uint32_t rstart = first_rev ? rs : fs;
				km[rank] = make_uint4(start, rstart, (first_rev ? rc : fc) | ((fc + rc) << 16),
						(uint32_t) (tid * IPT + q) | (palin ? 0x40000000u : 0u) | (first_rev ? 0x80000000u : 0u));
				if (palin) {
					// (positions of a palindromic k-mer are read from the table in the sweep: its one list is hit on both strands)
				} else if (STAGE == 1) {
					cs_bulk_g2s(bins + start, P.table + rstart, cnt * 4u, &s_mbar);      // region start: 64-byte aligned; cnt * 4: a multiple of 16
					atomicAdd(&s_tx, cnt * 4u);
				} else {
					for (uint32_t e4 = 0; e4 < cnt; e4 += 4)                              // 16 bytes per copy, both sides 16-byte aligned
						asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(cs_smem_u32(bins + start + e4)), "l"(P.table + rstart + e4) : "memory");
				}
			} else {
				km[rank] = make_uint4(start, m_fs[q], m_rs[q], m_fc[q] | ((uint32_t) (tid * IPT + q) << 16));      // list lengths < max_kfreq <= 65535
				if (start & 31u) atomicOr(&bstart[start >> 5], 1u << (start & 31u));
				for (uint32_t ch = (start + 31u) >> 5; (ch << 5) < start + cnt; ++ch) rbase[ch] = (uint16_t) rank;
			}
		}
		run += mine[q];
	}
	if (tid == NT - 1) {
		s_nhits = run & 0xFFFFFu;
		s_nkm = run >> 20;
	}
	__syncthreads();
	// every copy issued above must have landed before this block may leave (or reuse) its shared memory, whatever path the read takes
	if (STAGE == 1) {
		if (tid == 0) {
			cs_mbar_expect_tx(&s_mbar, s_tx);
			uint32_t spin = 0;
			while (!cs_mbar_try_wait(&s_mbar, 0u))
				if (++spin > (1u << 24)) __trap();
		}
		__syncthreads();
	} else if (STAGE == 2) {
		asm volatile("cp.async.commit_group;" ::: "memory");
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncthreads();
	}
	const uint32_t n_hits = s_nhits;
	auto to_exact = [&](int reason) {
		if (tid == 0) {
			meta[r].off = 0;
			meta[r].count = kCsPending;
			slow_list[atomicAdd(slow_count, 1u)] = (uint32_t) r;
			atomicAdd(slow_count + 1 + reason, 1u);
		}
	};
	if (n_hits > (uint32_t) MAXH) {
		to_exact(kCsWhyHits);
		return;
	}
	if (n_hits == 0) {
		if (tid == 0) {
			meta[r].off = 0;
			meta[r].count = 0;
			if (max_hit != nullptr) max_hit[r] = 0.0f;
		}
		return;
	}

	// double hashing in the small table
	auto slot0 = [&](uint32_t bin) { return (bin * 2654435761u) >> (32 - T2_LOG); };
	auto step_of = [&](uint32_t bin) { return ((bin * 0x9E3779B1u) >> (32 - T2_LOG)) | 1u; };

	// ---- B: bitmap sweep; hits that find their bit set are queued for the table --------------------------------
	auto hash_bit = [&](uint32_t bin, uint32_t &word, uint32_t &bit) {
		const uint32_t hb = (bin * 0x85EBCA6Bu) >> SEEN_SHIFT; // one of the "seen" bits
		word = hb >> 5;
		bit = 1u << (hb & 31);
	};
	auto rep_bit = [&](uint32_t bin, uint32_t &word, uint32_t &bit) {
		const uint32_t hb = (bin * 0xC2B2AE35u) >> 18;         // one of 16384 "repeated" bits, independent of the hash above
		word = hb >> 5;
		bit = 1u << (hb & 31);
	};
	// -> true if the hit found its bit set (to be queued for the exact table)
	auto first_sweep = [&](bool valid, uint32_t loc, bool rev, uint32_t corr, uint32_t h) -> bool {
		if (!valid) return false;
		if (loc < corr) {                                      // the reference's 64-bit wrap-around: exact kernel
			s_slow = 1;
			loc = corr;
		}
		const uint32_t bin = (loc - corr) >> P.bin_shift;
		uint32_t word, bit;
		hash_bit(bin, word, bit);
		const uint32_t old = atomicOr(&seen[word], bit);
		const bool again = (old & bit) != 0;                   // seen before (or a colliding bin): to be counted exactly
		bins[h] = bin | (rev ? kHitRev : 0u) | (again ? kHitInserted : 0u);
		return again;
	};
	// The hits are numbered in the reference's order (k-mer by k-mer, forward list then reverse list); a warp takes 32 consecutive
	// hits whatever k-mers they belong to, so every lane carries a hit and no list has a tail.  U chunks in flight per warp; the
	// hits of a batch that have to be queued take one slot reservation per warp (the order inside the queue does not matter).
	constexpr int U = 8;
	const uint32_t n_chunks = (n_hits + 31u) >> 5;
	const uint32_t lane_le = (2u << lane) - 1u;
	const int n_km = (int) s_nkm;
	if (TMA) {
		// The positions are in bins[] already (layout order).  Every thread walks its own run of ceil(n_hits / 256) consecutive slots:
		// perfectly balanced, conflict-free shared-memory strides, and the k-mer a slot belongs to (offset, strand) changes only every
		// ~30 slots, so it is looked up once per thread (binary search over the region starts) and then advanced.  A hit costs about a
		// dozen instructions.  (Measured alternatives: a warp per k-mer -- 10.9 k warp instructions per read for this sweep; a thread
		// per list -- 7.5 k and a long wait for the last round at the barrier; the flat chunks of the first generation -- ~7 k.)
		const uint32_t corr_r_base = (uint32_t) (len - k);
		const uint32_t per = (n_hits + NT - 1) / NT;
		uint32_t h = tid * per;
		const uint32_t h_end = min(n_hits, h + per);
		if (h < h_end) {
			int rank = cs_rank_of(km, n_km, h);
			uint32_t R = 0, first = 0, total = 0, r_end = 0, corr1 = 0, corr2 = 0, flag1 = 0, flag2 = 0, src = 0;
			bool palin = false;
			auto load_region = [&]() {
				const uint4 e = km[rank];
				R = e.x;
				src = e.y;
				first = e.z & 0xFFFFu;
				total = e.z >> 16;
				r_end = R + ((total + 3u) & ~3u);
				const uint32_t j = e.w & 0xFFFFu;
				const bool first_rev = (e.w >> 31) != 0;
				palin = (e.w & 0x40000000u) != 0;
				const uint32_t c_f = j, c_r = corr_r_base - j;
				corr1 = first_rev ? c_r : c_f;
				corr2 = first_rev ? c_f : c_r;
				flag1 = first_rev ? kHitRev : 0u;
				flag2 = first_rev ? 0u : kHitRev;
			};
			load_region();
			for (; h < h_end; ++h) {
				if (h >= r_end) {                                          // next k-mer (regions are contiguous)
					++rank;
					load_region();
				}
				const uint32_t o = h - R;
				if (o >= total) {                                          // padding behind the k-mer's hits
					bins[h] = kHitPad;
					continue;
				}
				const bool in1 = o < first;
				uint32_t loc = palin ? NGM_CS_LD(P.table + src + (in1 ? o : o - first)) : bins[h];
				const uint32_t corr = in1 ? corr1 : corr2;
				if (loc < corr) {                                          // the reference's 64-bit wrap-around: exact kernel
					s_slow = 1;
					loc = corr;
				}
				const uint32_t bin = (loc - corr) >> P.bin_shift;
				const uint32_t hb = (bin * 0x85EBCA6Bu) >> SEEN_SHIFT;
				const uint32_t bit = 1u << (hb & 31);
				const uint32_t old = atomicOr(&seen[hb >> 5], bit);
				uint32_t v = bin | (in1 ? flag1 : flag2);
				if (old & bit) {
					v |= kHitInserted;
					const uint32_t at = atomicAdd(&s_nq, 1u);
					if (at < (uint32_t) kCsQueue) s_queue[at] = (uint16_t) h;
				}
				bins[h] = v;
			}
		}
	}
	for (uint32_t c0 = warp; !TMA && c0 < n_chunks; c0 += (NT / 32) * U) {
		uint32_t loc[U], meta_j[U];
#pragma unroll
		for (int u = 0; u < U; ++u) {
			const uint32_t ch = c0 + u * (NT / 32);
			const uint32_t h = (ch << 5) + lane;
			loc[u] = 0;
			meta_j[u] = 0xFFFFFFFFu;
			if (h < n_hits) {                                       // implies ch < n_chunks
				const uint4 e = km[rbase[ch] + __popc(bstart[ch] & lane_le)];
				const uint32_t o = h - e.x, fc = e.w & 0xFFFFu;
				const bool rv = o >= fc;
				loc[u] = NGM_CS_LD(P.table + (rv ? e.z + (o - fc) : e.y + o));
				meta_j[u] = (e.w >> 16) | (rv ? 0x80000000u : 0u);
			}
		}
		uint32_t again = 0;
#pragma unroll
		for (int u = 0; u < U; ++u) {
			const uint32_t ch = c0 + u * (NT / 32);
			const uint32_t m = meta_j[u];
			const bool rv = (m >> 31) != 0;
			const uint32_t j = m & 0x7FFFFFFFu;
			if (first_sweep(m != 0xFFFFFFFFu, loc[u], rv, rv ? (uint32_t) (len - ((int) j + k)) : j, (ch << 5) + lane)) again |= 1u << u;
		}
		if (__any_sync(0xffffffffu, again != 0)) {
			const uint32_t mine_q = __popc(again);
			uint32_t incl_q = mine_q;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t v = __shfl_up_sync(0xffffffffu, incl_q, d);
				if (lane >= d) incl_q += v;
			}
			uint32_t base_q = 0;
			if (lane == 31) base_q = atomicAdd(&s_nq, incl_q);
			base_q = __shfl_sync(0xffffffffu, base_q, 31);
			uint32_t at = base_q + incl_q - mine_q;
			while (again) {
				const int u = __ffs(again) - 1;
				again &= again - 1;
				if (at < (uint32_t) kCsQueue) s_queue[at] = (uint16_t) (((c0 + u * (NT / 32)) << 5) + lane);
				++at;
			}
		}
	}
	__syncthreads();
	if (s_slow || s_nq > (uint32_t) kCsQueue || s_nq > (uint32_t) (T2 / 2)) {
		to_exact(s_slow ? kCsWhyWrap : kCsWhyQueue);
		return;
	}
	// one more vote for (slot, strand); remembers the running maximum and every (entry, strand) that reaches two votes
	auto bump = [&](uint32_t slot, bool rev) {
		const uint32_t before = atomicAdd(&cnts[slot], rev ? 0x10000u : 1u);
		const uint32_t now = ((before >> (rev ? 16 : 0)) & 0x7FFFu) + 1u;
		if (P.merged) {
			const uint32_t both = (before & 0x7FFFu) + ((before >> 16) & 0x7FFFu) + 1u;
			if (both > *(volatile uint32_t *) &s_maxm) atomicMax(&s_maxm, both);
		}
		if (now >= 2u) {
			if (now > *(volatile uint32_t *) &s_max) atomicMax(&s_max, now);       // (a stale value only costs a redundant atomic)
			if (now == 2u) {
				const uint32_t at = atomicAdd(&s_nmulti, 1u);
				if (at < kCsMaxMulti) s_multi[at] = (uint16_t) slot;
			}
		}
	};
	// ---- B2: the queued hits enter the exact table (dense: one hit per thread) -----------------------------------
	for (uint32_t q = tid; q < s_nq; q += NT) {
		const uint32_t t = bins[s_queue[q]];
		const uint32_t bin = t & kHitBin;
		uint32_t slot = slot0(bin);
		const uint32_t st = step_of(bin);
		bool done = false;
		for (int probe = 0; probe < 64; ++probe) {
			const uint32_t was = atomicCAS(&keys[slot], kCsEmpty, bin);
			if (was == kCsEmpty || was == bin) {
				bump(slot, (t & kHitRev) != 0);
				done = true;
				break;
			}
			slot = (slot + st) & MASK;
		}
		if (!done) s_slow = 1;
		uint32_t word, bit;
		rep_bit(bin, word, bit);
		atomicOr(&rep[word], bit);
	}
	__syncthreads();
	if (s_slow) {
		to_exact(kCsWhyTable);
		return;
	}
	// read-only lookup (the table is complete as far as keys go)
	auto find = [&](uint32_t bin) -> int {
		uint32_t slot = slot0(bin);
		const uint32_t st = step_of(bin);
		for (int probe = 0; probe < 64; ++probe) {
			const uint32_t kk = keys[slot];
			if (kk == bin) return (int) slot;
			if (kk == kCsEmpty) return -1;
			slot = (slot + st) & MASK;
		}
		return -1;
	};
	// ---- C: the hit that set the bit of a repeated bin (the "repeated" bitmap spares the lookup for ~97 % of the hits)
	auto sweep_c = [&]() {
		for (uint32_t h4 = 4 * tid; h4 < n_hits; h4 += 4 * NT) {   // four hits per thread and step (bins[] is 16-byte aligned)
			const uint4 q4 = *reinterpret_cast<const uint4 *>(bins + h4);
			const uint32_t tt[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
			for (int e = 0; e < 4; ++e) {
				const uint32_t t = tt[e];
				if (h4 + e >= n_hits || (t & kHitInserted)) continue;      // (padding slots carry every flag)
				uint32_t word, bit;
				rep_bit(t & kHitBin, word, bit);
				if (!(rep[word] & bit)) continue;
				const int slot = find(t & kHitBin);
				if (slot >= 0) {
					bump((uint32_t) slot, (t & kHitRev) != 0);
					bins[h4 + e] = t | kHitInserted;
				}
			}
		}
	};
	// Only entries that can still reach the threshold need their missing first vote: with Mq the best count so far (queued hits
	// only), the final threshold is at least thr_lb = max(kmer_min, sensitivity x Mq), and an entry whose count + 1 stays below it can
	// neither pass nor be the maximum.  When thr_lb > 2 those entries are few (the true locus and its neighbours): the sweep then
	// compares every hit with up to kCsMaxRes bins held in registers instead of hashing it into the "repeated" bitmap.  The complete
	// sweep still runs when the order of the list has to be worked out (phase E), when every bin is a candidate (path J), for the
	// both-strand maximum of the sensitivity estimate, and when more entries qualify than the registers hold.
	bool full_c = P.merged != 0 || s_nmulti > (uint32_t) kCsMaxMulti;
	const float thr_lb = fmaxf(P.kmer_min, __fmul_rn((float) s_max, P.sensitivity));
	if (!(thr_lb > 2.0f)) full_c = true;
	if (!full_c) {
		for (uint32_t i = tid; i < s_nmulti; i += NT) {
			const uint32_t sl = s_multi[i];
			const uint32_t c = cnts[sl];
			if ((float) (max(c & 0x7FFFu, (c >> 16) & 0x7FFFu) + 1u) >= thr_lb) {
				const uint32_t at = atomicAdd(&s_nres, 1u);
				if (at < 2u * kCsMaxRes) s_res[at] = keys[sl];
			}
		}
		__syncthreads();
		if (tid == 0) {                                         // an entry is listed once per strand that reached two votes
			const uint32_t listed = min(s_nres, 2u * (uint32_t) kCsMaxRes);
			uint32_t n = 0;
			for (uint32_t i = 0; i < listed; ++i) {                 // compaction in place (n <= i)
				const uint32_t v = s_res[i];
				bool dup = false;
				for (uint32_t j = 0; j < n; ++j) dup = dup || s_res[j] == v;
				if (!dup) s_res[n++] = v;
			}
			const bool over = s_nres > 2u * (uint32_t) kCsMaxRes || n > (uint32_t) kCsMaxRes;
			for (uint32_t j = n; j < (uint32_t) kCsMaxRes; ++j) s_res[j] = 0xFFFFFFFFu;
			s_nres = over ? 0xFFFFFFFFu : n;
		}
		__syncthreads();
		if (s_nres == 0xFFFFFFFFu) full_c = true;
	}
	if (full_c) {
		sweep_c();
	} else if (s_nres > 0) {
		uint32_t key[kCsMaxRes];
#pragma unroll
		for (int j = 0; j < kCsMaxRes; ++j) key[j] = s_res[j];
		for (uint32_t h4 = 4 * tid; h4 < n_hits; h4 += 4 * NT) {
			const uint4 q4 = *reinterpret_cast<const uint4 *>(bins + h4);
			const uint32_t tt[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
			for (int e = 0; e < 4; ++e) {
				const uint32_t b = tt[e] & ~kHitRev;               // equals a key only if kHitInserted is clear: a first hit
				bool hit = false;
#pragma unroll
				for (int j = 0; j < kCsMaxRes; ++j) hit = hit || b == key[j];
				if (hit && h4 + e < n_hits) {
					const int slot = find(b);
					if (slot >= 0) {
						bump((uint32_t) slot, (tt[e] & kHitRev) != 0);
						bins[h4 + e] = tt[e] | kHitInserted;
					}
				}
			}
		}
	}
	__syncthreads();
	// ---- D: maximum, threshold, accepted entries -----------------------------------------------------------------
	const uint32_t M = max(s_max, 1u);                         // n_hits > 0: some bin has one vote
	if (max_hit != nullptr && tid == 0) max_hit[r] = P.merged ? (float) max(s_maxm, 1u) : (float) M;
	const float thr = fmaxf(P.kmer_min, __fmul_rn((float) M, P.sensitivity));      // CS.cpp:193-196,271
	if (!(thr > 0.0f) || s_nmulti > (uint32_t) kCsMaxMulti) {      // zero threshold (strands without a vote pass): exact kernel
		to_exact(thr > 0.0f ? kCsWhyMulti : kCsWhyZeroThr);
		return;
	}
	if (!(thr > 1.0f)) {
		// ---- J: nothing stands out (threshold <= 1): EVERY bin with a vote is a candidate (CS.cpp:286-305), and as the
		// running threshold never exceeded one vote, every entry entered rList at its first hit (CS.cpp:199-202): the
		// list is the bins in the order of their first hit, forward strand before reverse strand.
		uint32_t *firsth = seen;                               // the bitmap is no longer needed: first hit per table slot
		for (int sl = tid; sl < T2; sl += NT) firsth[sl] = 0xFFFFFFFFu;
		__syncthreads();
		// (TMA variant: h runs over the reference's hit numbers, cs_slot_of() gives the slot the hit is stored in)
		for (uint32_t h = tid; h < n_hits; h += NT) {
			const uint32_t t = bins[h];
			if (t != kHitPad && (t & kHitInserted)) atomicMin(&firsth[find(t & kHitBin)], TMA ? cs_seq_of(km, n_km, h) : h);
		}
		__syncthreads();
		const uint32_t chunk = (n_hits + NT - 1) / NT;
		const uint32_t h0 = min(n_hits, tid * chunk), h1 = min(n_hits, h0 + chunk);
		auto emits = [&](uint32_t h, uint32_t &f, uint32_t &rv, uint32_t &bin) -> uint32_t {
			const uint32_t t = bins[TMA ? cs_slot_of(km, n_km, h) : h];
			if (t == kHitPad) return 0u;
			bin = t & kHitBin;
			if (!(t & kHitInserted)) {                             // a bin with exactly one vote
				f = (t & kHitRev) ? 0u : 1u;
				rv = 1u - f;
				return 1u;
			}
			const int sl = find(bin);
			if (firsth[sl] != h) return 0u;
			const uint32_t c = cnts[sl];
			f = c & 0x7FFFu;
			rv = (c >> 16) & 0x7FFFu;
			return (f ? 1u : 0u) + (rv ? 1u : 0u);
		};
		uint32_t mine_n = 0;
		for (uint32_t h = h0; h < h1; ++h) {
			uint32_t f, rv, bin;
			mine_n += emits(h, f, rv, bin);
		}
		uint32_t inc2 = mine_n;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t v = __shfl_up_sync(0xffffffffu, inc2, d);
			if (lane >= d) inc2 += v;
		}
		if (lane == 31) s_warp[warp] = inc2;
		__syncthreads();
		uint32_t before = inc2 - mine_n, total_c = 0;
#pragma unroll
		for (int w = 0; w < NT / 32; ++w) {
			before += (w < warp) ? s_warp[w] : 0u;
			total_c += s_warp[w];
		}
		if (!((long long) total_c < (long long) P.max_cmrs)) total_c = 0;      // CS.cpp:308-310
		if (tid == 0) {
			s_nord = total_c ? atomicAdd(cursor, total_c) : 0u;
			meta[r].off = s_nord;
			meta[r].count = total_c;
		}
		__syncthreads();
		const uint32_t off = s_nord;
		if (total_c && (unsigned long long) off + total_c <= heap_cap) {
			uint32_t w = off + before;
			for (uint32_t h = h0; h < h1; ++h) {
				uint32_t f = 0, rv = 0, bin = 0;
				if (emits(h, f, rv, bin) == 0) continue;
				if (f) {
					CsCand cd;
					cd.bin = bin;
					cd.votes = (uint16_t) f;
					cd.rev = 0;
					heap[w++] = cd;
				}
				if (rv) {
					CsCand cd;
					cd.bin = bin;
					cd.votes = (uint16_t) rv;
					cd.rev = 1;
					heap[w++] = cd;
				}
			}
		}
		return;
	}
	for (uint32_t i = tid; i < s_nmulti; i += NT) {            // only entries with two votes on a strand can pass thr > 1
		const uint32_t sl = s_multi[i];
		const uint32_t c = cnts[sl];
		const uint32_t a = ((float) (c & 0x7FFFu) >= thr ? 1u : 0u) + ((float) ((c >> 16) & 0x7FFFu) >= thr ? 1u : 0u);
		if (a && !(atomicOr(&cnts[sl], 0x8000u) & 0x8000u)) {      // bit 15: accepted (an entry can be listed once per strand)
			const uint32_t at = atomicAdd(&s_nacc, 1u);
			if (at < kCsMaxAccepted) s_acc[at] = sl;
			atomicAdd(&s_ncand, a);
		}
	}
	__syncthreads();
	const uint32_t nacc = s_nacc, ncand = s_ncand;
	if (nacc > kCsMaxAccepted) {
		to_exact(kCsWhyAccepted);
		return;
	}
	if (!((long long) ncand < (long long) P.max_cmrs)) {      // CS.cpp:308-310
		if (tid == 0) {
			meta[r].off = 0;
			meta[r].count = 0;
		}
		return;
	}
	if (nacc > 1) {
		if (!full_c) {                                             // the replay needs every entry's votes and first hit
			sweep_c();
			__syncthreads();
		}
		// ---- E: order of the list: replay the relevant hits in sequence ---------------------------------------
		for (uint32_t h = tid; h < n_hits; h += NT) {
			const uint32_t t = bins[h];
			if (!(t & kHitInserted) || t == kHitPad) continue;     // not in the table: a single-vote bin
			const int slot = find(t & kHitBin);
			const uint32_t c = cnts[slot];
			if ((c & 0x8000u) || (c & 0x7FFFu) >= 2u || ((c >> 16) & 0x7FFFu) >= 2u) {
				const uint32_t at = atomicAdd(&s_nitems, 1u);
				if (at < kCsMaxItems) {
					s_items_t[at] = TMA ? cs_seq_of(km, n_km, h) : h;      // sequence number of the hit
					s_items_s[at] = (uint32_t) slot | ((t & kHitRev) ? 0x80000000u : 0u);
				}
			}
		}
		__syncthreads();
		const int n_items = (int) s_nitems;
		if (n_items > kCsMaxItems) {
			to_exact(kCsWhyItems);
			return;
		}
		int n2 = 1;
		while (n2 < n_items) n2 <<= 1;
		for (int i = n_items + tid; i < n2; i += NT) {
			s_items_t[i] = 0xFFFFFFFFu;
			s_items_s[i] = 0;
		}
		__syncthreads();
		for (int size = 2; size <= n2; size <<= 1) {           // bitonic sort by sequence number
			for (int st = size >> 1; st > 0; st >>= 1) {
				for (int i = tid; i < n2; i += NT) {
					const int j = i ^ st;
					if (j > i) {
						const bool up = (i & size) == 0;
						const uint32_t a = s_items_t[i], b = s_items_t[j];
						if ((a > b) == up) {
							s_items_t[i] = b;
							s_items_t[j] = a;
							const uint32_t sa = s_items_s[i];
							s_items_s[i] = s_items_s[j];
							s_items_s[j] = sa;
						}
					}
				}
				__syncthreads();
			}
		}
		// votes of the hit's (entry, strand) right after this hit
		for (int i = tid; i < n_items; i += NT) {
			const uint32_t me = s_items_s[i];
			uint32_t c = 1;
			for (int j = 0; j < i; ++j) c += (s_items_s[j] == me);
			s_items_c[i] = (uint16_t) c;
		}
		__syncthreads();
		if (tid == 0) {
			uint32_t run_max = 0, nord = 0;
			for (int i = 0; i < n_items; ++i) {
				const uint32_t c = s_items_c[i];
				run_max = max(run_max, c);                     // CS.cpp:193-196
				const uint32_t slot = s_items_s[i] & 0x7FFFFFFFu;
				const uint32_t cc = cnts[slot];
				// CS.cpp:199-202: enters rList the first time a hit lifts it to the running threshold
				if ((cc & 0x8000u) && !(cc & 0x80000000u) && (float) c >= __fmul_rn((float) run_max, P.sensitivity)) {
					cnts[slot] = cc | 0x80000000u;
					s_acc[nord++] = slot;
				}
			}
			s_nord = nord;
		}
		__syncthreads();
		if (s_nord != nacc) {                                  // cannot happen; be loud rather than wrong
			to_exact(kCsWhyOrder);
			return;
		}
	}
	// ---- emit (CollectResultsStd, CS.cpp:286-305) -------------------------------------------------------------
	if (tid == 0) {
		const uint32_t off = atomicAdd(cursor, ncand);
		meta[r].off = off;
		meta[r].count = ncand;
		if ((unsigned long long) off + ncand <= heap_cap) {
			uint32_t w = off;
			for (uint32_t a = 0; a < nacc; ++a) {
				const uint32_t slot = s_acc[a];
				const uint32_t c = cnts[slot];
				const uint32_t f = c & 0x7FFFu, rv = (c >> 16) & 0x7FFFu;
				if ((float) f >= thr) {
					CsCand cd;
					cd.bin = keys[slot];
					cd.votes = (uint16_t) f;
					cd.rev = 0;
					heap[w++] = cd;
				}
				if ((float) rv >= thr) {
					CsCand cd;
					cd.bin = keys[slot];
					cd.votes = (uint16_t) rv;
					cd.rev = 1;
					heap[w++] = cd;
				}
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------------------
// search: exact sequential path (one lane per read, table in global memory)
// ---------------------------------------------------------------------------------------------------------
struct CsExactEntry {         // CSTableEntry (LocationScore.h:9-14)
	uint32_t loc;
	uint32_t state;
	float fscore, rscore;
};

// One lane's search state (CS members rTable, rList, currentState, maxHitNumber, currentThresh, hpoc)
struct CsExactState {
	CsExactEntry *tab;
	uint32_t *rlist;
	uint32_t gen, rlen, tmask;
	int bits;
	float maxhit, thresh, maxmerged;
	uint32_t hpoc;            // probe steps left before the reference gives up (CS.cpp:176-178); 0 = unlimited
	bool overflow;
};

// CS::PrefixSearch + AddLocationStd (CS.cpp:114-213) for one k-mer at read offset o
__device__ __forceinline__ void cs_exact_vote(const CsDev &P, CsExactState &S, uint32_t prefix, int o, int len, float weight) {
	CsLists L;
	if (S.overflow || !cs_lookup(P, prefix, L)) return;
	const uint64_t corr_r = (uint64_t) len - ((uint64_t) o + (uint64_t) P.k);
	const uint32_t total = L.fc + L.rc;
	for (uint32_t i = 0; i < total; ++i) {
		const bool rev = i >= L.fc;
		const uint64_t loc = P.table[rev ? L.rs + (i - L.fc) : L.fs + i];
		const uint64_t bin = (loc - (rev ? corr_r : (uint64_t) o)) >> P.bin_shift;      // GetBin on 64 bits
		uint32_t e = (uint32_t) ((bin * 11400714819323199488ull) >> (64 - S.bits));      // CS::Hash
		bool found;
		uint32_t probes = 0;
		while ((found = ((S.tab[e].state & 0x7FFFFFFFu) == S.gen)) && !((uint64_t) S.tab[e].loc == bin)) {
			e = (e + 1) & S.tmask;
			if (++probes > S.tmask || (S.hpoc != 0 && --S.hpoc == 0)) {
				S.overflow = true;
				return;
			}
		}
		float score = weight;
		if (!found) {
			S.tab[e].loc = (uint32_t) bin;
			S.tab[e].state = S.gen;
			S.tab[e].fscore = rev ? 0.0f : weight;
			S.tab[e].rscore = rev ? weight : 0.0f;
		} else if (rev) {
			score = (S.tab[e].rscore = __fadd_rn(S.tab[e].rscore, weight));
		} else {
			score = (S.tab[e].fscore = __fadd_rn(S.tab[e].fscore, weight));
		}
		if (score > S.maxhit) {
			S.maxhit = score;
			S.thresh = __fmul_rn(S.maxhit, P.sensitivity);
		}
		S.maxmerged = fmaxf(S.maxmerged, S.tab[e].fscore + S.tab[e].rscore);
		if (!(S.tab[e].state & 0x80000000u) && score >= S.thresh) {
			S.tab[e].state |= 0x80000000u;
			S.rlist[S.rlen++] = e;
		}
	}
}

// CS::PrefixMutateSearch (CS.cpp:53-112) for one read k-mer.  bs_mapping: the k-mer and, depth first, every k-mer with a subset of its
// `from` bases replaced (PrefixMutateSearchEx: the subsets of the replaceable positions in lexicographic order of their sorted index lists,
// positions counted from the k-mer's last base); k-mers with more than `cutoff` such bases are skipped.  slam_seq: the k-mer with weight 1,
// then its single replacements with weight 1 / (replaceable bases + 1) (PrefixMutateSearchSlamSeq).
__device__ void cs_exact_mutate(const CsDev &P, CsExactState &S, uint32_t prefix, int o, int len, uint32_t from, uint32_t to) {
	int pos[16];
	int c = 0;
	for (int i = 0; i < P.k; ++i)
		if (((prefix >> (2 * i)) & 3u) == from) pos[c++] = i;
	if (P.mut_mode == 2) {
		cs_exact_vote(P, S, prefix, o, len, 1.0f);
		const float w = __fdiv_rn(1.0f, (float) (c + 1));
		for (int j = 0; j < c; ++j) cs_exact_vote(P, S, (prefix & ~(3u << (2 * pos[j]))) | (to << (2 * pos[j])), o, len, w);
		return;
	}
	if (c > P.mut_cutoff) return;
	int sub[16];
	int d = 0;
	for (;;) {
		uint32_t p = prefix;
		for (int j = 0; j < d; ++j) p = (p & ~(3u << (2 * pos[sub[j]]))) | (to << (2 * pos[sub[j]]));
		cs_exact_vote(P, S, p, o, len, 1.0f);
		if (S.overflow) return;
		const int last = d ? sub[d - 1] : -1;
		if (last + 1 < c) {
			sub[d++] = last + 1;
		} else {
			--d;                                               // (d >= 1 here unless c == 0)
			if (d <= 0) return;
			sub[d - 1] += 1;
		}
	}
}

// CS::PrefixIteration (CSstatic.cpp:26-76) as a loop, with the read-side k-mer skip: fn(prefix, offset) for every emitted k-mer
template <typename F>
__device__ void cs_exact_iterate(const uint8_t *seq, int length, int k, uint32_t prefixskip, F fn) {
	const uint32_t mask = k < 16 ? (1u << (2 * k)) - 1u : 0xFFFFFFFFu;
	int offset = 0;
	for (;;) {
		if (length < k) return;
		if (*seq == 'N') {
			int n_skip = 1;
			while (n_skip < length && seq[n_skip] == 'N') ++n_skip;
			seq += n_skip;
			if (n_skip >= length - k) return;
			length -= n_skip;
			offset += n_skip;
		}
		uint32_t prefix = 0;
		bool restart = false;
		int i;
		for (i = 0; i < k - 1; ++i) {
			if (seq[i] == 'N') {
				restart = true;
				break;
			}
			prefix = (prefix << 2) | cs_enc2(seq[i]);
		}
		if (!restart) {
			uint32_t skipcount = prefixskip;
			for (i = k - 1; i < length; ++i) {
				if (seq[i] == 'N') {
					restart = true;
					break;
				}
				prefix = ((prefix << 2) | cs_enc2(seq[i])) & mask;
				if (skipcount == prefixskip) {
					fn(prefix, offset + i + 1 - k);
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

__global__ void __launch_bounds__(32) cs_search_exact_kernel(const CsDev P, const uint8_t *__restrict__ reads, int n_reads, int stride,
		CsMeta *__restrict__ meta, CsCand *__restrict__ heap, uint32_t heap_cap, uint32_t *__restrict__ cursor, const uint32_t *__restrict__ work_list,
		const uint32_t *__restrict__ work_count, CsExactEntry *__restrict__ tables, uint32_t *__restrict__ rlists, uint32_t *__restrict__ gens,
		float *__restrict__ max_hit) {
	if (threadIdx.x != 0) return;
	const uint32_t TLEN = 1u << P.ex_bits;
	CsExactState S;
	S.tab = tables + (size_t) blockIdx.x * TLEN;
	S.rlist = rlists + (size_t) blockIdx.x * TLEN;
	S.tmask = TLEN - 1;
	S.bits = P.ex_bits;
	S.gen = gens[blockIdx.x];
	const uint32_t n_work = work_list != nullptr ? *work_count : (uint32_t) n_reads;
	for (uint32_t w = blockIdx.x; w < n_work; w += gridDim.x) {
		const int r = work_list != nullptr ? (int) work_list[w] : (int) w;
		const uint8_t *seq = reads + (size_t) r * stride;
		int len = 0;
		while (len < stride && seq[len] != 0) ++len;
		S.gen = (S.gen + 1) & 0x7FFFFFFFu;                     // CS::RunBatch, CS.cpp:351-356
		if (S.gen == 0x7FFFFFFFu) S.gen = 1;
		S.rlen = 0;
		S.maxhit = 0.0f;
		S.thresh = 0.0f;
		S.maxmerged = 0.0f;
		S.overflow = false;
		S.hpoc = 0;
		if (P.mut_mode == 0) {
			for (int o = 0; o + P.k <= len && !S.overflow; ++o) {
				uint32_t prefix;
				if (!cs_read_kmer(seq, len, o, P.k, prefix)) continue;
				cs_exact_vote(P, S, prefix, o, len, 1.0f);
			}
		} else {
			// the table of CS::RunBatch's last retry (2^20 slots, 0.777 x slots probe steps, CS.cpp:404-428): a read that fits a smaller
			// table gets the same list from this one, a read that overflows this one keeps no candidates
			S.hpoc = (uint32_t) __fmul_rn((float) (int) TLEN, 0.777f);
			const bool second = P.mut_paired && (r & 1);
			const uint32_t from = P.mut_mode == 2 ? (second ? 3u : 1u) : (second ? 0u : 2u);
			const uint32_t to = P.mut_mode == 2 ? (second ? 0u : 2u) : (second ? 3u : 1u);
			cs_exact_iterate(seq, len, P.k, P.mut_mode == 1 ? (uint32_t) P.read_skip : 0u,
					[&](uint32_t prefix, int o) { cs_exact_mutate(P, S, prefix, o, len, from, to); });
		}
		const bool overflow = S.overflow;
		const uint32_t rlen = S.rlen;
		if (max_hit != nullptr) max_hit[r] = P.merged ? S.maxmerged : S.maxhit;
		const float thr = fmaxf(P.kmer_min, S.thresh);
		uint32_t n = 0;
		for (uint32_t i = 0; i < rlen && !overflow; ++i) {
			const CsExactEntry t = S.tab[S.rlist[i]];
			n += (t.fscore >= thr) + (t.rscore >= thr);
		}
		if (overflow || !((long long) n < (long long) P.max_cmrs)) n = 0;
		uint32_t off = 0;
		if (n) {
			off = atomicAdd(cursor, n);
			if ((unsigned long long) off + n <= heap_cap) {
				uint32_t at = off;
				for (uint32_t i = 0; i < rlen; ++i) {
					const CsExactEntry t = S.tab[S.rlist[i]];
					if (t.fscore >= thr) {
						CsCand cd;
						cd.bin = t.loc;
						cd.votes = (uint16_t) t.fscore;
						cd.rev = 0;
						if (P.heap_votes != nullptr) P.heap_votes[at] = t.fscore;
						heap[at++] = cd;
					}
					if (t.rscore >= thr) {
						CsCand cd;
						cd.bin = t.loc;
						cd.votes = (uint16_t) t.rscore;
						cd.rev = 1;
						if (P.heap_votes != nullptr) P.heap_votes[at] = t.rscore;
						heap[at++] = cd;
					}
				}
			}
		}
		meta[r].off = off;
		meta[r].count = overflow ? 0u : n;
	}
	gens[blockIdx.x] = S.gen;
}

// ---------------------------------------------------------------------------------------------------------
// CSR assembly: counts -> (exclusive scan on the host side of this file's launcher) -> pairs
// ---------------------------------------------------------------------------------------------------------
__global__ void cs_counts_kernel(const CsMeta *__restrict__ meta, int n_reads, int *__restrict__ counts) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r > n_reads) return;
	counts[r] = r < n_reads ? (int) meta[r].count : 0;
}

// LocationScore -> (read, window) descriptor: Location = ResolveBin(bin) (CS.h:170-175); ScoreBuffer fetches the
// window at Location - corridor/2 and uses RevSeq for reverse candidates (ScoreBuffer.cpp:92-114)
__global__ void cs_gather_kernel(const CsMeta *__restrict__ meta, const CsCand *__restrict__ heap, uint32_t heap_cap, const int *__restrict__ begin,
		int n_reads, int bin_shift, int corridor, uint32_t out_cap, ngm_b200_pair *__restrict__ pairs, float *__restrict__ votes,
		const float *__restrict__ heap_votes, int second_mates) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const CsMeta m = meta[r];
	const uint32_t b = (uint32_t) begin[r];
	const unsigned long long half = bin_shift > 0 ? 1ull << (bin_shift - 1) : 0ull;
	for (uint32_t i = 0; i < m.count; ++i) {
		if ((unsigned long long) m.off + i >= heap_cap || b + i >= out_cap) break;
		const CsCand c = heap[m.off + i];
		ngm_b200_pair p;
		p.window_start = (((unsigned long long) c.bin << bin_shift) + half) - (unsigned long long) (corridor >> 1);
		p.read_index = (uint32_t) r;
		// the bs / SLAMseq direction flag ScoreBuffer hands to BatchScore: the strand, inverted for second mates (ScoreBuffer.cpp:92-110)
		const bool dir = (c.rev != 0) != (second_mates && (r & 1));
		p.flags = (c.rev ? NGM_B200_PAIR_REVERSE : 0u) | (dir ? NGM_B200_PAIR_DIR : 0u);
		pairs[b + i] = p;
		if (votes != nullptr) votes[b + i] = heap_votes != nullptr ? heap_votes[m.off + i] : (float) c.votes;
	}
}

}  // namespace ngm
