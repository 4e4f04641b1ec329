// ngm_cs.cu -- host side of the candidate-search entry points of include/ngm_b200.h.
//
// Index construction follows CompactPrefixTable (src/PrefixTable.cpp:196-245,357-498,577-690) for one table unit:
// the k-mer walk of CS::PrefixIteration (src/CSstatic.cpp:26-76) is resolved on the host into "runs" of emitted
// k-mers (it only depends on where the N's are), the k-mers themselves are extracted, de-duplicated, counted, scanned
// and sorted on the device.  CUB's radix sort / scan are used for this one-off start-up step only; the search path is
// made of the kernels in ngm_cs.cuh.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cub/cub.cuh>

#include "ngm_ctx.h"
#include "ngm_cs.cuh"

using namespace ngm;

namespace ngm {

struct CsState {
	ngm_b200_cs_params hp;
	int k = 0, step = 0, bin_shift = 0;
	uint32_t n_prefix = 0, table_len = 0;
	int max_kfreq = 0;
	bool ready = false;
	DevBuf d_tabu, d_table, d_weight, d_both, d_table2;
	// search scratch
	DevBuf d_meta, d_heap, d_cursor, d_slow_list, d_slow_count, d_counts, d_begin, d_scan_tmp, d_pairs, d_votes;
	DevBuf d_ex_tables, d_ex_rlists, d_ex_gens, d_heap_votes;
	HostBuf h_total;
	int ex_blocks = 0;
	int ex_bits = 0;               // table size the exact kernel's scratch was allocated for
	// bs-mapping / SLAMseq k-mer mutation (ngm_b200_cs_configure_mutation)
	int mut_mode = 0, mut_cutoff = 6, mut_paired = 0, mut_read_skip = 0;
	uint64_t exact_reads = 0;      // reads the last search sent to the exact kernel
	bool exact_all = false;
};

void cs_release(CsState *cs) {
	if (cs == nullptr) return;
	DevBuf *db[] = { &cs->d_tabu, &cs->d_table, &cs->d_weight, &cs->d_both, &cs->d_table2, &cs->d_meta, &cs->d_heap, &cs->d_cursor, &cs->d_slow_list, &cs->d_slow_count,
			&cs->d_counts, &cs->d_begin, &cs->d_scan_tmp, &cs->d_pairs, &cs->d_votes, &cs->d_ex_tables, &cs->d_ex_rlists, &cs->d_ex_gens, &cs->d_heap_votes };
	for (DevBuf *b : db) b->release();
	cs->h_total.release();
	delete cs;
}

int cs_share_index(ngm_b200_ctx *lane, const ngm_b200_ctx *root) {
	const CsState *r = root->cs;
	if (r == nullptr || !r->ready) {
		if (lane->cs) lane->cs->ready = false;
		return NGM_B200_OK;
	}
	if (lane->cs == nullptr) lane->cs = new CsState();
	CsState *s = lane->cs;
	s->hp = r->hp;
	s->k = r->k;
	s->step = r->step;
	s->bin_shift = r->bin_shift;
	s->n_prefix = r->n_prefix;
	s->table_len = r->table_len;
	s->max_kfreq = r->max_kfreq;
	s->mut_mode = r->mut_mode;
	s->mut_cutoff = r->mut_cutoff;
	s->mut_paired = r->mut_paired;
	s->mut_read_skip = r->mut_read_skip;
	s->d_tabu.borrow(r->d_tabu);
	s->d_table.borrow(r->d_table);
	s->d_weight.borrow(r->d_weight);
	s->d_both.borrow(r->d_both);
	s->d_table2.borrow(r->d_table2);
	s->ready = true;
	return NGM_B200_OK;
}

}  // namespace ngm

namespace {

int check_params(const ngm_b200_cs_params *p) {
	if (p == nullptr) return fail(NGM_B200_EINVAL, "cs params == NULL");
	if (p->kmer < 8 || p->kmer > 14) return fail(NGM_B200_EINVAL, "kmer %d not in [8, 14]", p->kmer);
	if (p->kmer_skip < 0 || p->kmer_skip > 64) return fail(NGM_B200_EINVAL, "kmer_skip %d out of range", p->kmer_skip);
	if (p->bin_size < 0 || p->bin_size > 8) return fail(NGM_B200_EINVAL, "bin_size %d out of range", p->bin_size);
	if (!(p->sensitivity >= 0.0f && p->sensitivity <= 1.0f)) return fail(NGM_B200_EINVAL, "sensitivity %g not in [0, 1]", p->sensitivity);
	return NGM_B200_OK;
}

CsState *fresh_state(ngm_b200_ctx *c, const ngm_b200_cs_params *p) {
	if (c->cs) cs_release(c->cs);
	c->cs = new CsState();
	c->cs->hp = *p;
	c->cs->k = p->kmer;
	c->cs->step = p->kmer_skip + 1;
	c->cs->bin_shift = p->bin_size;
	c->cs->n_prefix = 1u << (2 * p->kmer);
	c->epoch += 1;
	return c->cs;
}

// CompactPrefixTable::stats (PrefixTable.cpp:151-194): ceil(max(100, avg + 5 sd)) over the per-k-mer list lengths
int max_kfreq_from_sums(uint32_t n_prefix, unsigned long long s1, unsigned long long s2) {
	const double il = (double) n_prefix, sum = (double) s1, sum2 = (double) s2;
	const double avg = sum / il;
	const double sd = std::sqrt(sum2 / (il - 1) - 2.0 * avg * (sum / (il - 1)) + ((il * std::pow(avg, 2.0)) / (il - 1)));
	return (int) std::ceil(std::max(100.0, avg + 5 * sd));
}

struct NRun {
	uint64_t s, e;             // inclusive
};

// CS::PrefixIteration (CSstatic.cpp:26-76) over one contig at the level of N-free stretches.  The contig buffer the
// reference iterates has `len` characters of which only the first `real` are decoded bases (DecodeRefSequence is
// given the contig length as buffer length and decodes two less, SequenceProvider.cpp:384); the rest reads as code 0.
void contig_runs(uint64_t start, uint64_t len, uint64_t real, const std::vector<NRun> &nruns, int k, int step, std::vector<CsRun> &out,
		uint64_t &emit_total) {
	const uint64_t contig_base = emit_total;
	// N runs clipped to the decoded part of this contig, relative coordinates
	std::vector<NRun> nr;
	auto it = std::lower_bound(nruns.begin(), nruns.end(), start, [](const NRun &a, uint64_t v) { return a.e < v; });
	for (; it != nruns.end() && it->s < start + real; ++it) {
		const uint64_t s = std::max(it->s, start) - start, e = std::min(it->e, start + real - 1) - start;
		nr.push_back({s, e});
	}
	size_t ni = 0;
	uint64_t pos = 0, length = len;
	auto is_n = [&](uint64_t p) {
		while (ni < nr.size() && nr[ni].e < p) ++ni;
		return ni < nr.size() && nr[ni].s <= p;
	};
	for (;;) {
		if (length < (uint64_t) k) break;                          // :27-28
		if (is_n(pos)) {                                           // :30-41
			const uint64_t n_skip = nr[ni].e - pos + 1;
			pos += n_skip;
			if (n_skip >= length - (uint64_t) k) break;
			length -= n_skip;
		}
		while (ni < nr.size() && nr[ni].e < pos) ++ni;
		const uint64_t end = pos + length;                        // one past the last character of the buffer
		const uint64_t q = ni < nr.size() ? nr[ni].s : end;        // next N (or the end)
		if (q - pos >= (uint64_t) k) {
			CsRun r;
			r.start = start + pos;
			r.tail_start = start + real;
			r.emit_base = emit_total;
			r.contig_base = contig_base;
			r.n_emit = (uint32_t) ((q - pos - (uint64_t) k) / (uint64_t) step + 1);
			r.pad = 0;
			out.push_back(r);
			emit_total += r.n_emit;
		}
		if (q == end) break;
		length -= q - pos + 1;                                     // restart behind the N (:48-50,60-62)
		pos = q + 1;
	}
}

int finish_index(ngm_b200_ctx *c, CsState *cs, unsigned long long s1, unsigned long long s2) {
	cs->max_kfreq = cs->hp.max_kfreq > 0 ? cs->hp.max_kfreq : max_kfreq_from_sums(cs->n_prefix, s1, s2);
	// the search's copy: 16-byte entries + positions ordered by k-mer pair (cs_pair_copy_kernel)
	const uint32_t NP = cs->n_prefix;
	cudaStream_t st = c->stream;
	DevBuf d_size, d_off2, d_tmp;
	CU(cs->d_both.ensure((size_t) NP * sizeof(uint4)));
	CU(d_size.ensure((size_t) NP * 4));
	CU(d_off2.ensure((size_t) NP * 4));
	cs_pair_size_kernel<<<(NP + 255) / 256, 256, 0, st>>>(cs->d_tabu.as<uint32_t>(), NP, cs->k, d_size.as<uint32_t>());
	size_t tmp_bytes = 0;
	CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_size.as<uint32_t>(), d_off2.as<uint32_t>(), (int) NP, st));
	CU(d_tmp.ensure(tmp_bytes));
	CU(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, d_size.as<uint32_t>(), d_off2.as<uint32_t>(), (int) NP, st));
	uint32_t last_off = 0, last_size = 0;                      // the padded total sizes the search's copy of the positions
	CU(cudaMemcpyAsync(&last_off, d_off2.as<uint32_t>() + (NP - 1), 4, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(&last_size, d_size.as<uint32_t>() + (NP - 1), 4, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	const unsigned long long padded = (unsigned long long) last_off + last_size;
	if (padded >= 0xFFFFFFF0ull) return fail(NGM_B200_ERANGE, "position table too large after padding (%llu entries)", padded);
	CU(cs->d_table2.ensure((size_t) padded * 4 + 64));
	cs_pair_copy_kernel<<<(NP + 255) / 256, 256, 0, st>>>(cs->d_tabu.as<uint32_t>(), NP, cs->k, d_off2.as<uint32_t>(), cs->d_table.as<uint32_t>(),
			cs->d_table2.as<uint32_t>(), cs->d_both.as<uint4>());
	c->launches += 4;
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(st));
	d_size.release();
	d_off2.release();
	d_tmp.release();
	cs->ready = true;
	c->epoch += 1;
	return NGM_B200_OK;
}

}  // namespace

extern "C" {

int ngm_b200_cs_build_index(ngm_b200_ctx *c, const ngm_b200_cs_params *params, const ngm_b200_contig *contigs, uint32_t n_contigs) {
	if (c == nullptr || contigs == nullptr || n_contigs == 0) return fail(NGM_B200_EINVAL, "NULL argument");
	int rc = check_params(params);
	if (rc) return rc;
	if (!c->have_ref) return fail(NGM_B200_ESTATE, "set_reference must precede cs_build_index");
	if (c->concat_len >= 0xFFFFFFFFull) return fail(NGM_B200_ERANGE, "references of 2^32 - 1 bases or more need several table units (not supported)");
	CU(cudaSetDevice(c->device));
	CsState *cs = fresh_state(c, params);
	cudaStream_t st = c->stream;
	const int k = cs->k;
	const uint32_t *ref4 = c->d_ref4.as<uint32_t>();

	// 1. where are the N's (device) -> runs of emitted k-mers (host)
	const uint64_t n_words = (c->concat_len + 7) / 8;
	const uint32_t cap = 8u << 20;
	DevBuf d_list, d_cnt;
	CU(d_list.ensure((size_t) cap * 8));
	CU(d_cnt.ensure(4));
	CU(cudaMemsetAsync(d_cnt.p, 0, 4, st));
	cs_find_n_kernel<<<(unsigned) ((n_words + 255) / 256), 256, 0, st>>>(ref4, n_words, c->concat_len, d_list.as<unsigned long long>(), cap,
			d_cnt.as<uint32_t>());
	c->launches += 1;
	uint32_t n_bound = 0;
	CU(cudaMemcpyAsync(&n_bound, d_cnt.p, 4, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	if (n_bound > cap) {
		d_list.release();
		d_cnt.release();
		return fail(NGM_B200_ERANGE, "reference holds more than %u N-run boundaries", cap);
	}
	std::vector<unsigned long long> bounds(n_bound);
	if (n_bound) CU(cudaMemcpy(bounds.data(), d_list.p, (size_t) n_bound * 8, cudaMemcpyDeviceToHost));
	d_list.release();
	d_cnt.release();
	std::sort(bounds.begin(), bounds.end());
	std::vector<NRun> nruns;
	for (size_t i = 0; i < bounds.size();) {                       // (start, end) pairs; a single N is both
		const uint64_t s = bounds[i] >> 1;
		if (bounds[i] & 1ull) return fail(NGM_B200_ECUDA, "inconsistent N-run list");
		size_t j = i + 1;
		if (j >= bounds.size() || !(bounds[j] & 1ull)) return fail(NGM_B200_ECUDA, "inconsistent N-run list");
		nruns.push_back({s, (uint64_t) (bounds[j] >> 1)});
		i = j + 1;
	}
	std::vector<CsRun> runs;
	uint64_t n_emit = 0;
	for (uint32_t i = 0; i < n_contigs; ++i) {
		const uint64_t start = contigs[i].start, len = contigs[i].length;
		if (start + len > c->concat_len + 1) return fail(NGM_B200_EINVAL, "contig %u lies outside the reference", i);
		if (len < 2) continue;
		const uint64_t real = (start & 1) ? len - 1 : len - 2;     // SequenceProvider.cpp:384,416-426
		contig_runs(start, len, real, nruns, k, cs->step, runs, n_emit);
	}
	if (n_emit >= 0x7FFFFFF0ull) return fail(NGM_B200_ERANGE, "too many indexed positions (%llu)", (unsigned long long) n_emit);

	// 2. k-mers, counts
	const uint32_t NP = cs->n_prefix;
	DevBuf d_runs, d_freq, d_keys, d_vals, d_keys2, d_sums, d_tmp, d_off;
	CU(d_freq.ensure(((size_t) NP + 1) * 4));
	CU(cudaMemsetAsync(d_freq.p, 0, ((size_t) NP + 1) * 4, st));
	CU(d_sums.ensure(16));
	CU(cudaMemsetAsync(d_sums.p, 0, 16, st));
	const size_t ne = (size_t) std::max<uint64_t>(n_emit, 1);
	CU(d_keys.ensure(ne * 4));
	CU(d_vals.ensure(ne * 4));
	if (n_emit) {
		CU(d_runs.ensure(runs.size() * sizeof(CsRun)));
		CU(cudaMemcpyAsync(d_runs.p, runs.data(), runs.size() * sizeof(CsRun), cudaMemcpyHostToDevice, st));
		cs_emit_kernel<<<(unsigned) ((n_emit + 255) / 256), 256, 0, st>>>(ref4, d_runs.as<CsRun>(), (int) runs.size(), n_emit, k, cs->step, cs->bin_shift,
				params->skip_rep ? 1 : 0, NP, d_keys.as<uint32_t>(), d_vals.as<uint32_t>(), d_freq.as<uint32_t>());
		c->launches += 1;
	}
	CU(cs->d_weight.ensure((size_t) NP + 1));
	CU(cudaMemsetAsync(cs->d_weight.p, 0, (size_t) NP + 1, st));
	cs_weight_kernel<<<(NP + 255) / 256, 256, 0, st>>>(d_freq.as<uint32_t>(), NP, k, cs->d_weight.as<int8_t>(), d_sums.as<unsigned long long>());
	c->launches += 1;
	// 3. offsets
	CU(d_off.ensure(((size_t) NP + 1) * 4));
	size_t tmp_bytes = 0;
	CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_freq.as<uint32_t>(), d_off.as<uint32_t>(), (int) (NP + 1), st));
	CU(d_tmp.ensure(tmp_bytes));
	CU(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, d_freq.as<uint32_t>(), d_off.as<uint32_t>(), (int) (NP + 1), st));
	CU(cs->d_tabu.ensure(((size_t) NP + 1) * 4));
	cs_tabu_kernel<<<(NP + 1 + 255) / 256, 256, 0, st>>>(d_off.as<uint32_t>(), cs->d_weight.as<int8_t>(), NP, cs->d_tabu.as<uint32_t>());
	c->launches += 1;
	uint32_t total = 0;
	unsigned long long sums[2] = {0, 0};
	CU(cudaMemcpyAsync(&total, d_off.as<uint32_t>() + NP, 4, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(sums, d_sums.p, 16, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	d_freq.release();
	d_off.release();
	cs->table_len = total;
	// 4. positions grouped by k-mer: stable sort of (k-mer, position) in emission order keeps positions ascending
	CU(cs->d_table.ensure(ne * 4 + 4));
	if (n_emit) {
		CU(d_keys2.ensure(ne * 4));
		tmp_bytes = 0;
		CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys.as<uint32_t>(), d_keys2.as<uint32_t>(), d_vals.as<uint32_t>(),
				cs->d_table.as<uint32_t>(), (int) n_emit, 0, 2 * k + 1, st));
		CU(d_tmp.ensure(tmp_bytes));
		CU(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_keys.as<uint32_t>(), d_keys2.as<uint32_t>(), d_vals.as<uint32_t>(),
				cs->d_table.as<uint32_t>(), (int) n_emit, 0, 2 * k + 1, st));
		cs_zero_unused_kernel<<<(NP + 255) / 256, 256, 0, st>>>(cs->d_tabu.as<uint32_t>(), NP, cs->d_table.as<uint32_t>());
		c->launches += 1;
	}
	CU(cudaStreamSynchronize(st));
	CU(cudaGetLastError());
	d_keys.release();
	d_keys2.release();
	d_vals.release();
	d_tmp.release();
	d_runs.release();
	d_sums.release();
	return finish_index(c, cs, sums[0], sums[1]);
}

int ngm_b200_cs_load_index(ngm_b200_ctx *c, const ngm_b200_cs_params *params, const uint32_t *tab, const int8_t *weight, uint32_t index_len,
		const uint32_t *table, uint32_t table_len) {
	if (c == nullptr || tab == nullptr || weight == nullptr || (table == nullptr && table_len)) return fail(NGM_B200_EINVAL, "NULL argument");
	int rc = check_params(params);
	if (rc) return rc;
	const uint32_t NP = 1u << (2 * params->kmer);
	if (index_len != NP + 1) return fail(NGM_B200_EINVAL, "index length %u does not belong to kmer %d", index_len, params->kmer);
	// a stale or damaged cache file must fail here, not as out-of-bounds reads in the search kernels: offsets (1-based in the file,
	// PrefixTable.cpp:436-498) non-decreasing and ending at table_len, positions inside the reference this context holds
	if (tab[0] < 1) return fail(NGM_B200_EINVAL, "prefix table: the first offset must be 1 (got %u)", tab[0]);
	for (uint32_t j = 0; j < NP; ++j)
		if (tab[j + 1] < tab[j]) return fail(NGM_B200_EINVAL, "prefix table: offsets decrease at k-mer %u", j);
	if (tab[NP] - 1u != table_len) return fail(NGM_B200_EINVAL, "prefix table: the offsets end at %u, the table holds %u positions", tab[NP] - 1u, table_len);
	if (c->have_ref) {
		uint32_t top = 0;
		for (uint32_t i = 0; i < table_len; ++i) top = std::max(top, table[i]);
		if ((uint64_t) top >= c->concat_len) return fail(NGM_B200_EINVAL, "prefix table: position %u lies outside the reference (%llu bases): built for another reference?",
				top, (unsigned long long) c->concat_len);
	}
	CU(cudaSetDevice(c->device));
	CsState *cs = fresh_state(c, params);
	cudaStream_t st = c->stream;
	DevBuf d_tab;
	CU(d_tab.ensure((size_t) index_len * 4));
	CU(cs->d_weight.ensure((size_t) index_len));
	CU(cs->d_tabu.ensure((size_t) index_len * 4));
	CU(cs->d_table.ensure((size_t) table_len * 4 + 4));
	CU(cudaMemcpyAsync(d_tab.p, tab, (size_t) index_len * 4, cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(cs->d_weight.p, weight, (size_t) index_len, cudaMemcpyHostToDevice, st));
	if (table_len) CU(cudaMemcpyAsync(cs->d_table.p, table, (size_t) table_len * 4, cudaMemcpyHostToDevice, st));
	cs_tabu_from_file_kernel<<<(NP + 1 + 255) / 256, 256, 0, st>>>(d_tab.as<uint32_t>(), cs->d_weight.as<int8_t>(), NP, cs->d_tabu.as<uint32_t>());
	c->launches += 1;
	CU(cudaStreamSynchronize(st));
	CU(cudaGetLastError());
	d_tab.release();
	cs->table_len = table_len;
	// stats() over the file's index (PrefixTable.cpp:160-176)
	unsigned long long s1 = 0, s2 = 0;
	for (uint32_t j = 0; j < NP; ++j) {
		const unsigned long long cnt = tab[j + 1] - tab[j];
		s1 += cnt;
		s2 += cnt * cnt;
	}
	return finish_index(c, cs, s1, s2);
}

// device-to-device forms: what a rank does with the prefix table it received over NCCL (the index is broadcast at start-up)
int ngm_b200_dev_cs_export_index(ngm_b200_ctx *c, void *d_tab, void *d_weight, void *d_table, void *stream) {
	if (c == nullptr || c->cs == nullptr || !c->cs->ready) return fail(NGM_B200_ESTATE, "no candidate-search index");
	if (d_tab == nullptr || d_weight == nullptr || (d_table == nullptr && c->cs->table_len)) return fail(NGM_B200_EINVAL, "NULL argument");
	CU(cudaSetDevice(c->device));
	CsState *cs = c->cs;
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	const uint32_t NP = cs->n_prefix;
	cs_export_tab_kernel<<<(NP + 1 + 255) / 256, 256, 0, st>>>(cs->d_tabu.as<uint32_t>(), NP, static_cast<uint32_t *>(d_tab));
	c->launches += 1;
	CU(cudaMemcpyAsync(d_weight, cs->d_weight.p, (size_t) NP + 1, cudaMemcpyDeviceToDevice, st));
	if (cs->table_len) CU(cudaMemcpyAsync(d_table, cs->d_table.p, (size_t) cs->table_len * 4, cudaMemcpyDeviceToDevice, st));
	CU(cudaGetLastError());
	return NGM_B200_OK;
}

int ngm_b200_dev_cs_load_index(ngm_b200_ctx *c, const ngm_b200_cs_params *params, const void *d_tab, const void *d_weight, uint32_t index_len,
		const void *d_table, uint32_t table_len, void *stream) {
	if (c == nullptr || d_tab == nullptr || d_weight == nullptr || (d_table == nullptr && table_len)) return fail(NGM_B200_EINVAL, "NULL argument");
	int rc = check_params(params);
	if (rc) return rc;
	if (params->max_kfreq <= 0) return fail(NGM_B200_EINVAL, "the device form takes max_kfreq from the sender (ngm_b200_cs_index_info)");
	const uint32_t NP = 1u << (2 * params->kmer);
	if (index_len != NP + 1) return fail(NGM_B200_EINVAL, "index length %u does not belong to kmer %d", index_len, params->kmer);
	CU(cudaSetDevice(c->device));
	CsState *cs = fresh_state(c, params);
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	CU(cs->d_weight.ensure((size_t) index_len));
	CU(cs->d_tabu.ensure((size_t) index_len * 4));
	CU(cs->d_table.ensure((size_t) table_len * 4 + 4));
	CU(cudaMemcpyAsync(cs->d_weight.p, d_weight, (size_t) index_len, cudaMemcpyDeviceToDevice, st));
	if (table_len) CU(cudaMemcpyAsync(cs->d_table.p, d_table, (size_t) table_len * 4, cudaMemcpyDeviceToDevice, st));
	cs_tabu_from_file_kernel<<<(NP + 1 + 255) / 256, 256, 0, st>>>(static_cast<const uint32_t *>(d_tab), cs->d_weight.as<int8_t>(), NP, cs->d_tabu.as<uint32_t>());
	c->launches += 1;
	CU(cudaStreamSynchronize(st));
	CU(cudaGetLastError());
	cs->table_len = table_len;
	return finish_index(c, cs, 0, 0);                              // (max_kfreq is given: the sums are not needed)
}

int ngm_b200_cs_index_info(const ngm_b200_ctx *c, uint32_t *index_len, uint32_t *table_len, int32_t *max_kfreq) {
	if (c == nullptr || c->cs == nullptr || !c->cs->ready) return fail(NGM_B200_ESTATE, "no candidate-search index");
	if (index_len) *index_len = c->cs->n_prefix + 1;
	if (table_len) *table_len = c->cs->table_len;
	if (max_kfreq) *max_kfreq = c->cs->max_kfreq;
	return NGM_B200_OK;
}

int ngm_b200_cs_export_index(ngm_b200_ctx *c, uint32_t *tab, int8_t *weight, uint32_t *table) {
	if (c == nullptr || c->cs == nullptr || !c->cs->ready) return fail(NGM_B200_ESTATE, "no candidate-search index");
	CU(cudaSetDevice(c->device));
	CsState *cs = c->cs;
	const uint32_t NP = cs->n_prefix;
	if (tab) {
		DevBuf d_tab;
		CU(d_tab.ensure(((size_t) NP + 1) * 4));
		cs_export_tab_kernel<<<(NP + 1 + 255) / 256, 256, 0, c->stream>>>(cs->d_tabu.as<uint32_t>(), NP, d_tab.as<uint32_t>());
		c->launches += 1;
		CU(cudaMemcpyAsync(tab, d_tab.p, ((size_t) NP + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
		CU(cudaStreamSynchronize(c->stream));
		d_tab.release();
	}
	if (weight) CU(cudaMemcpy(weight, cs->d_weight.p, (size_t) NP + 1, cudaMemcpyDeviceToHost));
	if (table && cs->table_len) CU(cudaMemcpy(table, cs->d_table.p, (size_t) cs->table_len * 4, cudaMemcpyDeviceToHost));
	return NGM_B200_OK;
}

int ngm_b200_dev_cs_search(ngm_b200_ctx *c, const void *d_ascii_reads, int n_reads, int stride, int mode_flags, void *d_cand_begin, void *d_pairs,
		void *d_votes, uint32_t capacity, void *d_max_hit, void *stream) {
	if (c == nullptr || d_ascii_reads == nullptr || d_cand_begin == nullptr || d_pairs == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	if (c->cs == nullptr || !c->cs->ready) return fail(NGM_B200_ESTATE, "cs_build_index / cs_load_index must precede cs_search");
	if (n_reads <= 0) return 0;
	if (stride <= 0 || stride > kCsMaxStride) return fail(NGM_B200_EINVAL, "read stride %d not in [1, %d]", stride, kCsMaxStride);
	CU(cudaSetDevice(c->device));
	CsState *cs = c->cs;
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	CsDev P;
	P.tabu = cs->d_tabu.as<uint32_t>();
	P.both = cs->d_both.as<uint4>();
	P.table = cs->d_table2.as<uint32_t>();
	P.k = cs->k;
	P.bin_shift = cs->bin_shift;
	P.max_kfreq = cs->max_kfreq;
	P.max_cmrs = cs->hp.max_cmrs > 0 ? cs->hp.max_cmrs : 0x7FFFFFFF;
	P.sensitivity = cs->hp.sensitivity;
	P.kmer_min = cs->hp.kmer_min;
	P.merged = (mode_flags & 2) ? 1 : 0;
	P.mut_mode = cs->mut_mode;
	P.mut_cutoff = cs->mut_cutoff;
	P.mut_paired = cs->mut_paired;
	P.read_skip = cs->mut_mode == 1 ? cs->mut_read_skip : 0;
	P.ex_bits = cs->mut_mode != 0 ? 20 : kCsExactBits;
	P.heap_votes = nullptr;
	if (cs->mut_mode == 2) {
		CU(cs->d_heap_votes.ensure((size_t) capacity * sizeof(float) + 8));
		P.heap_votes = cs->d_heap_votes.as<float>();
	}
	CU(cs->d_meta.ensure((size_t) n_reads * sizeof(CsMeta)));
	CU(cs->d_heap.ensure((size_t) capacity * sizeof(CsCand) + 8));
	CU(cs->d_cursor.ensure(4));
	CU(cs->d_slow_count.ensure(64));
	CU(cs->d_slow_list.ensure((size_t) n_reads * 4));
	CU(cs->d_counts.ensure(((size_t) n_reads + 1) * 4));
	CU(cudaMemsetAsync(cs->d_cursor.p, 0, 4, st));
	CU(cudaMemsetAsync(cs->d_slow_count.p, 0, 64, st));
	{
		int sms = 0;
		CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
		int want = std::max(8, 2 * sms);
		const size_t tl = (size_t) 1 << P.ex_bits;
		if (cs->mut_mode != 0 && n_reads > want && cs->ex_blocks < std::min(n_reads, 16 * sms)) {
			// every read of a mutated run is searched by one lane, a chain of dependent memory accesses: throughput is the number of lanes in
			// flight.  One per read of the batch, up to 16 per SM, as many as a fifth of the free device memory holds (20 MB of table + list
			// per lane); grown when a larger batch arrives
			size_t free_b = 0, total_b = 0;
			CU(cudaMemGetInfo(&free_b, &total_b));
			const size_t per_lane = tl * (sizeof(CsExactEntry) + 4);
			const size_t have = cs->ex_bits == P.ex_bits ? (size_t) cs->ex_blocks : 0;
			const size_t fit = have + (free_b / 5) / per_lane;
			want = (int) std::max<size_t>((size_t) want, std::min<size_t>(fit, (size_t) std::min(n_reads, 16 * sms)));
		}
		if (cs->mut_mode != 0)
			if (const char *e = getenv("NGM_B200_CS_MUT_LANES")) want = std::max(1, atoi(e));
		if (cs->ex_blocks == 0 || cs->ex_bits != P.ex_bits || cs->ex_blocks < want) {
			cs->ex_blocks = want;
			cs->ex_bits = P.ex_bits;
			CU(cs->d_ex_tables.ensure(tl * sizeof(CsExactEntry) * cs->ex_blocks));
			CU(cs->d_ex_rlists.ensure(tl * 4 * cs->ex_blocks));
			CU(cs->d_ex_gens.ensure((size_t) cs->ex_blocks * 4));
			CU(cudaMemsetAsync(cs->d_ex_tables.p, 0xFF, tl * sizeof(CsExactEntry) * cs->ex_blocks, st));
			CU(cudaMemsetAsync(cs->d_ex_gens.p, 0, (size_t) cs->ex_blocks * 4, st));
		}
	}
	const uint8_t *reads = static_cast<const uint8_t *>(d_ascii_reads);
	CsMeta *meta = cs->d_meta.as<CsMeta>();
	CsCand *heap = cs->d_heap.as<CsCand>();
	uint32_t *cursor = cs->d_cursor.as<uint32_t>();
	float *max_hit = static_cast<float *>(d_max_hit);
	const bool exact_only = (mode_flags & 1) != 0 || cs->mut_mode != 0;      // the mutated k-mers are enumerated by the sequential kernel only
	bool exact_only_fallback = false;
	cs->exact_all = exact_only;
	cs->exact_reads = (uint64_t) n_reads;
	if (!exact_only) {
		// table size by the expected number of distinct bins per read: ~ (L - k + 1) k-mers x 2 lists x mean list length
		auto launch = [&](auto kern, size_t smem) -> cudaError_t {
			cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
			if (e != cudaSuccess) return e;
			e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
			if (e != cudaSuccess) return e;
			kern<<<n_reads, 256, smem, st>>>(P, reads, n_reads, stride, meta, heap, capacity, cursor, cs->d_slow_list.as<uint32_t>(),
					cs->d_slow_count.as<uint32_t>(), max_hit);
			return cudaGetLastError();
		};
		const double mean_list = (double) cs->table_len / (double) cs->n_prefix;
		const double expect = (double) std::max(1, stride - cs->k + 1) * 2.0 * std::max(mean_list, 0.5);      // hits per read
		const bool bins_fit = ((c->concat_len + 1024) >> cs->bin_shift) < (1ull << 30);      // two flag bits ride on every stored bin
		const bool scan_fits = (uint64_t) std::max(1, stride - cs->k + 1) * (uint64_t) P.max_kfreq < (1ull << 20) && P.max_kfreq <= 65535;      // packed scan: hits in 20 bits; list lengths in 16
		const int n_kmers_max = stride - cs->k + 1;
		// TMA variant (default): bulk copies into the hit array; its regions are padded to four slots, ~1.5 slots per k-mer
		// NGM_B200_CS_STAGE: how the position lists reach shared memory.  0 (default) = per-lane loads inside the flat sweep; 1 = one bulk copy
		// per k-mer region (cp.async.bulk / UBLKCP, completion on an mbarrier) + per-thread runs; 2 = 16-byte cp.async (LDGSTS) + per-thread runs.
		// Measured on B200, 4 M x 150 bp vs 3 Gbp (profiles/r2_cs_staging.md): 49.5 / 44.5 / 35.9 M reads/s -- staging adds a second dependent
		// DRAM round trip per read (index entry -> list) behind a block-wide barrier and ~1.4 k serialised UBLKCP issue instructions per read,
		// which the cheaper sweep (12 instructions per hit) does not buy back.  The staged variants stay selectable and are covered by the tests.
		static const int stage = [] { const char *e = getenv("NGM_B200_CS_STAGE"); return e == nullptr ? 0 : std::max(0, std::min(2, atoi(e))); }();
		const bool use_tma = stage != 0;
		const double pad = use_tma ? 1.5 * (double) std::max(1, stride - cs->k + 1) : 0.0;
		int config = (n_kmers_max <= 256 && expect + pad <= 4500.0) ? 1 : (n_kmers_max <= 256 && expect + pad <= 8000.0) ? 2 : (n_kmers_max <= 512 && expect + pad <= 14500.0) ? 3 : 4;
		if (!use_tma) config = (n_kmers_max <= 256 && expect <= 4250.0) ? 1 : (n_kmers_max <= 256 && expect <= 7700.0) ? 2 : (n_kmers_max <= 512 && expect <= 14000.0) ? 3 : 4;
		if (const char *e = getenv("NGM_B200_CS_CONFIG")) {        // testing hook: run a small case through a larger configuration
			const int want = atoi(e);
			if (want > config && want <= 4) config = want;
		}
		if (!bins_fit || !scan_fits) {
			cs->exact_all = true;                                  // (bin_size 0/1 on > 1 Gbp: sequential kernel only)
			exact_only_fallback = true;
		} else if (use_tma) {
			if (stage == 1) {
				if (config == 1) CU(launch(cs_search_kernel<10, 256, 4864, 0, 1>, CsSmem<10, 256, 4864>::bytes_staged));                  // 150 bp on 3 Gbp: 5 blocks / SM
				else if (config == 2) CU(launch(cs_search_kernel<11, 256, 8448, 4096, 1>, CsSmem<11, 256, 8448, 4096>::bytes_staged));      // 250 bp on 3 Gbp: 3 blocks / SM
				else if (config == 3) CU(launch(cs_search_kernel<12, 512, 16384, 0, 1>, CsSmem<12, 512, 16384>::bytes_staged));
				else CU(launch(cs_search_kernel<12, 1024, 36864, 0, 1>, CsSmem<12, 1024, 36864>::bytes_staged));
			} else {
				if (config == 1) CU(launch(cs_search_kernel<10, 256, 4864, 0, 2>, CsSmem<10, 256, 4864>::bytes_staged));
				else if (config == 2) CU(launch(cs_search_kernel<11, 256, 8448, 4096, 2>, CsSmem<11, 256, 8448, 4096>::bytes_staged));
				else if (config == 3) CU(launch(cs_search_kernel<12, 512, 16384, 0, 2>, CsSmem<12, 512, 16384>::bytes_staged));
				else CU(launch(cs_search_kernel<12, 1024, 36864, 0, 2>, CsSmem<12, 1024, 36864>::bytes_staged));
			}
		} else if (config == 1) {
			CU(launch(cs_search_kernel<10, 256, 4608>, CsSmem<10, 256, 4608>::bytes));
		} else if (config == 2) {
			CU(launch(cs_search_kernel<11, 256, 8192, 4096>, CsSmem<11, 256, 8192, 4096>::bytes));
		} else if (config == 3) {
			CU(launch(cs_search_kernel<12, 512, 16384>, CsSmem<12, 512, 16384>::bytes));
		} else {
			CU(launch(cs_search_kernel<12, 1024, 36864>, CsSmem<12, 1024, 36864>::bytes));
		}
		c->launches += 1;
	}
	cs_search_exact_kernel<<<cs->ex_blocks, 32, 0, st>>>(P, reads, n_reads, stride, meta, heap, capacity, cursor,
			(exact_only || exact_only_fallback) ? nullptr : cs->d_slow_list.as<uint32_t>(), cs->d_slow_count.as<uint32_t>(), cs->d_ex_tables.as<CsExactEntry>(),
			cs->d_ex_rlists.as<uint32_t>(), cs->d_ex_gens.as<uint32_t>(), max_hit);
	c->launches += 1;
	CU(cudaGetLastError());
	// CSR
	cs_counts_kernel<<<(n_reads + 1 + 255) / 256, 256, 0, st>>>(meta, n_reads, cs->d_counts.as<int>());
	size_t tmp_bytes = 0;
	CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cs->d_counts.as<int>(), static_cast<int *>(d_cand_begin), n_reads + 1, st));
	CU(cs->d_scan_tmp.ensure(tmp_bytes));
	CU(cub::DeviceScan::ExclusiveSum(cs->d_scan_tmp.p, tmp_bytes, cs->d_counts.as<int>(), static_cast<int *>(d_cand_begin), n_reads + 1, st));
	cs_gather_kernel<<<(n_reads + 255) / 256, 256, 0, st>>>(meta, heap, capacity, static_cast<const int *>(d_cand_begin), n_reads, cs->bin_shift,
			c->dp.corridor, capacity, static_cast<ngm_b200_pair *>(d_pairs), static_cast<float *>(d_votes), P.heap_votes,
			cs->mut_mode != 0 && cs->mut_paired);
	c->launches += 3;
	CU(cudaGetLastError());
	return n_reads;
}

int ngm_b200_cs_search(ngm_b200_ctx *c, const char *reads, int n_reads, int stride, int mode_flags, int32_t *cand_begin, ngm_b200_pair *pairs,
		float *votes, size_t capacity, size_t *total, float *max_hit) {
	if (c == nullptr || reads == nullptr || cand_begin == nullptr || (pairs == nullptr && capacity)) return fail(NGM_B200_EINVAL, "NULL argument");
	if (c->cs == nullptr || !c->cs->ready) return fail(NGM_B200_ESTATE, "cs_build_index / cs_load_index must precede cs_search");
	if (n_reads <= 0) return 0;
	if (capacity > 0x7FFFFFFFull) capacity = 0x7FFFFFFFull;
	CU(cudaSetDevice(c->device));
	CsState *cs = c->cs;
	cudaStream_t st = c->stream;
	DevBuf d_reads, d_mh;
	CU(d_reads.ensure((size_t) n_reads * stride));
	CU(cs->d_begin.ensure(((size_t) n_reads + 1) * 4));
	CU(cs->d_pairs.ensure(std::max<size_t>(capacity, 1) * sizeof(ngm_b200_pair)));
	CU(cs->d_votes.ensure(std::max<size_t>(capacity, 1) * 4));
	if (max_hit) CU(d_mh.ensure((size_t) n_reads * 4));
	CU(cudaMemcpyAsync(d_reads.p, reads, (size_t) n_reads * stride, cudaMemcpyHostToDevice, st));
	int rc = ngm_b200_dev_cs_search(c, d_reads.p, n_reads, stride, mode_flags, cs->d_begin.p, cs->d_pairs.p, cs->d_votes.p, (uint32_t) capacity,
			max_hit ? d_mh.p : nullptr, st);
	if (rc < 0) {
		d_reads.release();
		d_mh.release();
		return rc;
	}
	CU(cudaMemcpyAsync(cand_begin, cs->d_begin.p, ((size_t) n_reads + 1) * 4, cudaMemcpyDeviceToHost, st));
	if (max_hit) CU(cudaMemcpyAsync(max_hit, d_mh.p, (size_t) n_reads * 4, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	d_reads.release();
	d_mh.release();
	const size_t tot = (size_t) cand_begin[n_reads];
	if (total) *total = tot;
	if (tot > capacity) return fail(NGM_B200_ERANGE, "candidate buffer too small: %zu entries needed", tot);
	if (tot) {
		CU(cudaMemcpy(pairs, cs->d_pairs.p, tot * sizeof(ngm_b200_pair), cudaMemcpyDeviceToHost));
		if (votes) CU(cudaMemcpy(votes, cs->d_votes.p, tot * 4, cudaMemcpyDeviceToHost));
	}
	return n_reads;
}

uint64_t ngm_b200_cs_exact_reads(const ngm_b200_ctx *c) {
	if (c == nullptr || c->cs == nullptr) return 0;
	if (c->cs->exact_all) return c->cs->exact_reads;
	uint32_t slow = 0;                                         // counter of the last search (synchronises the device)
	if (c->cs->d_slow_count.p == nullptr || cudaSetDevice(c->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
			cudaMemcpy(&slow, c->cs->d_slow_count.p, 4, cudaMemcpyDeviceToHost) != cudaSuccess)
		return 0;
	return slow;
}

// CS::RunBatch's choice of k-mer callback (CS.cpp:341-343) and of the mutated base (:362-380)
int ngm_b200_cs_configure_mutation(ngm_b200_ctx *c, int bs_mapping, int slam_seq, int bs_cutoff, int paired, int read_kmer_skip) {
	if (c == nullptr || c->cs == nullptr || !c->cs->ready) return fail(NGM_B200_ESTATE, "no candidate-search index");
	if (bs_mapping == 1 && slam_seq != 0) return fail(NGM_B200_EINVAL, "bs_mapping and slam_seq exclude each other (Config.cpp:454-457)");
	if (bs_cutoff < 0 || read_kmer_skip < 0) return fail(NGM_B200_EINVAL, "bs_cutoff %d / read_kmer_skip %d < 0", bs_cutoff, read_kmer_skip);
	c->cs->mut_mode = bs_mapping == 1 ? 1 : ((slam_seq & 4) ? 2 : 0);
	c->cs->mut_cutoff = bs_cutoff;
	c->cs->mut_paired = paired ? 1 : 0;
	c->cs->mut_read_skip = read_kmer_skip;
	c->epoch += 1;
	return NGM_B200_OK;
}

int ngm_b200_cs_set_sensitivity(ngm_b200_ctx *c, float sensitivity) {
	if (c == nullptr || c->cs == nullptr || !c->cs->ready) return fail(NGM_B200_ESTATE, "no candidate-search index");
	if (!(sensitivity >= 0.0f && sensitivity <= 1.0f)) return fail(NGM_B200_EINVAL, "sensitivity %g not in [0, 1]", sensitivity);
	c->cs->hp.sensitivity = sensitivity;
	c->epoch += 1;
	return NGM_B200_OK;
}

// ReadProvider::init (ReadProvider.cpp:236-251,310-325) with CollectResultsFallback (:53-79): every sampled read contributes
// best-vote / possible-votes, votes counted with both strands added; the estimate is the mean clamped to [0.3, 0.9].
int ngm_b200_cs_estimate_sensitivity(ngm_b200_ctx *c, const char *sampled_reads, int n, int stride, float *sensitivity) {
	if (c == nullptr || sampled_reads == nullptr || sensitivity == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	if (c->cs == nullptr || !c->cs->ready) return fail(NGM_B200_ESTATE, "cs_build_index / cs_load_index must precede the estimate");
	if (n <= 0) return fail(NGM_B200_EINVAL, "no sampled reads");
	std::vector<int32_t> begin((size_t) n + 1);
	std::vector<float> best((size_t) n);
	ngm_b200_pair dummy;
	size_t total = 0;
	// ReadProvider::init votes with its own plain PrefixSearch unless bs_mapping is set, and under bs_mapping it does not use the estimate at
	// all (ReadProvider.cpp:194-197,326,378-384: 0.5 unless -s): the estimate never sees mutated k-mers
	const int mut_mode = c->cs->mut_mode;
	c->cs->mut_mode = 0;
	int rc = ngm_b200_cs_search(c, sampled_reads, n, stride, 2, begin.data(), &dummy, nullptr, 0, &total, best.data());
	c->cs->mut_mode = mut_mode;
	if (rc < 0 && rc != NGM_B200_ERANGE) return rc;                 // the candidates themselves are not wanted
	const int skip = c->cs->step;
	float sum = 0.0f;
	int count = 0;
	for (int r = 0; r < n; ++r) {
		const char *row = sampled_reads + (size_t) r * stride;
		int len = 0;
		while (len < stride && row[len] != 0) ++len;
		const int max = (int) std::ceil((double) ((len - c->cs->k + 1) / skip) * 1.0);
		const float cur = best[r];
		if ((float) max > 1.0f && cur <= (float) max) {
			sum += cur / (float) max;
			count += 1;
		}
	}
	float avg = sum / (float) count * 1.0f;                         // NaN when nothing qualified, like the reference
	*sensitivity = std::min(std::max(0.3f, avg), 0.9f);
	return count;
}

int ngm_b200_cs_exact_reasons(const ngm_b200_ctx *c, uint32_t *out, int n) {
	if (c == nullptr || c->cs == nullptr || out == nullptr || c->cs->d_slow_count.p == nullptr) return fail(NGM_B200_ESTATE, "no candidate search yet");
	uint32_t tmp[16] = {0};
	if (cudaSetDevice(c->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
			cudaMemcpy(tmp, c->cs->d_slow_count.p, 64, cudaMemcpyDeviceToHost) != cudaSuccess)
		return fail(NGM_B200_ECUDA, "cannot read the counters");
	for (int i = 0; i < n && i < (int) kCsWhyCount; ++i) out[i] = tmp[1 + i];
	return (int) kCsWhyCount;
}

}  // extern "C"
