// ngm_map.cu -- one call per read batch for a host that has handed the whole mapping loop to the device:
//
//   CS::RunBatch (src/CS.cpp:340-436)  ->  ScoreBuffer::DoRun (src/ScoreBuffer.cpp:80-226: BatchScore of every candidate, top1SE or
//   top1PE)  ->  AlignmentBuffer::DoRun (src/AlignmentBuffer.cpp:64-147: BatchAlign of the selected candidate)
//
// from host buffers to host buffers; everything in between stays in HBM (candidates, scores, winners, pointer matrices).  The records go
// to ngm_b200_format_sam or to the caller's own writer.  Built from the device-pointer entry points of this library; no CPU path.
#include <algorithm>
#include <cstdint>

#include "ngm_ctx.h"

namespace ngm {

struct MapState {
	DevBuf d_reads, d_begin, d_pairs, d_scores, d_maxhit, d_best, d_mapq, d_ntop, d_pfail, d_wpairs, d_wscores, d_recs, d_strings, d_cursor;
};

void map_release(MapState *m) {
	if (m == nullptr) return;
	DevBuf *all[] = { &m->d_reads, &m->d_begin, &m->d_pairs, &m->d_scores, &m->d_maxhit, &m->d_best, &m->d_mapq, &m->d_ntop, &m->d_pfail, &m->d_wpairs,
			&m->d_wscores, &m->d_recs, &m->d_strings, &m->d_cursor };
	for (DevBuf *b : all) b->release();
	delete m;
}

}  // namespace ngm

using namespace ngm;

int ngm_b200_map_batch(ngm_b200_ctx *c, const char *reads, int n_reads, int stride, int mode, int paired, ngm_b200_map_result *res) {
	if (c == nullptr || reads == nullptr || res == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	if (n_reads <= 0) return 0;
	if (res->cand_begin == nullptr || res->best_pair == nullptr || res->mapq == nullptr || res->num_top == nullptr || res->max_hit == nullptr ||
			res->recs == nullptr || (res->capacity && (res->pairs == nullptr || res->scores == nullptr)) || (res->str_capacity && res->strings == nullptr) ||
			(paired && res->pair_fail == nullptr))
		return fail(NGM_B200_EINVAL, "NULL result array");
	if (paired && (n_reads & 1)) return fail(NGM_B200_EINVAL, "paired batches hold the mates in rows 2f and 2f + 1: %d rows", n_reads);
	if (res->capacity > 0x7FFFFFFFull || res->str_capacity > 0xFFFFFFFFull) return fail(NGM_B200_EINVAL, "capacity beyond 32 bits");
	CU(cudaSetDevice(c->device));
	if (c->map == nullptr) c->map = new MapState();
	MapState *m = c->map;
	cudaStream_t st = c->stream;
	const size_t n = (size_t) n_reads, cap = std::max<size_t>(res->capacity, 1), scap = std::max<size_t>(res->str_capacity, 16);
	CU(m->d_reads.ensure(n * (size_t) stride));
	CU(m->d_begin.ensure((n + 1) * 4));
	CU(m->d_pairs.ensure(cap * sizeof(ngm_b200_pair)));
	CU(m->d_scores.ensure(cap * 4));
	CU(m->d_maxhit.ensure(n * 4));
	CU(m->d_best.ensure(n * 4));
	CU(m->d_mapq.ensure(n * 4));
	CU(m->d_ntop.ensure(n * 4));
	CU(m->d_pfail.ensure(n * 4));
	CU(m->d_wpairs.ensure(n * sizeof(ngm_b200_pair)));
	CU(m->d_wscores.ensure(n * 4));
	CU(m->d_recs.ensure(n * sizeof(ngm_b200_align_rec)));
	CU(m->d_strings.ensure(scap));
	CU(m->d_cursor.ensure(4));
	CU(cudaMemcpyAsync(m->d_reads.p, reads, n * (size_t) stride, cudaMemcpyHostToDevice, st));
	int rc = ngm_b200_dev_set_reads(c, m->d_reads.p, n_reads, stride, st);
	if (rc < 0) return rc;
	rc = ngm_b200_dev_cs_search(c, m->d_reads.p, n_reads, stride, 0, m->d_begin.p, m->d_pairs.p, nullptr, (uint32_t) res->capacity, m->d_maxhit.p, st);
	if (rc < 0) return rc;
	CU(cudaMemcpyAsync(res->cand_begin, m->d_begin.p, (n + 1) * 4, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	const size_t total = (size_t) res->cand_begin[n_reads];
	res->n_candidates = total;
	res->str_used = 0;
	if (total > res->capacity) return fail(NGM_B200_ERANGE, "candidate arrays too small: %zu entries needed", total);
	if (total) {
		rc = ngm_b200_dev_score_pairs(c, mode, (int) total, m->d_pairs.p, m->d_scores.p, st);
		if (rc < 0) return rc;
	}
	int64_t pe_sum = 0, pe_count = 0;                          // a batch that has to be repeated must start from the same insert-size sums
	if (paired && (rc = ngm_b200_pe_insert_stats(c, &pe_sum, &pe_count)) < 0) return rc;
	if (paired) rc = ngm_b200_dev_select_pairs(c, n_reads, m->d_begin.p, m->d_pairs.p, m->d_scores.p, (uint32_t) total, m->d_best.p, m->d_mapq.p, m->d_ntop.p, m->d_pfail.p, st);
	else rc = ngm_b200_dev_select_top1_ex(c, n_reads, m->d_begin.p, m->d_scores.p, m->d_best.p, m->d_mapq.p, m->d_ntop.p, st);
	if (rc < 0) return rc;
	rc = ngm_b200_dev_gather_winners_scored(c, n_reads, m->d_pairs.p, m->d_scores.p, m->d_best.p, m->d_wpairs.p, m->d_wscores.p, st);
	if (rc < 0) return rc;
	CU(cudaMemsetAsync(m->d_cursor.p, 0, 4, st));
	rc = ngm_b200_dev_align_pairs_scored(c, mode, n_reads, m->d_wpairs.p, m->d_wscores.p, m->d_recs.p, m->d_strings.p, (uint32_t) res->str_capacity, m->d_cursor.p, st);
	if (rc < 0) return rc;
	uint32_t used = 0;
	CU(cudaMemcpyAsync(&used, m->d_cursor.p, 4, cudaMemcpyDeviceToHost, st));
	if (total) {
		CU(cudaMemcpyAsync(res->pairs, m->d_pairs.p, total * sizeof(ngm_b200_pair), cudaMemcpyDeviceToHost, st));
		CU(cudaMemcpyAsync(res->scores, m->d_scores.p, total * 4, cudaMemcpyDeviceToHost, st));
	}
	CU(cudaMemcpyAsync(res->best_pair, m->d_best.p, n * 4, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(res->mapq, m->d_mapq.p, n * 4, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(res->num_top, m->d_ntop.p, n * 4, cudaMemcpyDeviceToHost, st));
	if (paired) CU(cudaMemcpyAsync(res->pair_fail, m->d_pfail.p, n * 4, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(res->max_hit, m->d_maxhit.p, n * 4, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(res->recs, m->d_recs.p, n * sizeof(ngm_b200_align_rec), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	res->str_used = used;
	if ((size_t) used > res->str_capacity) {
		if (paired) ngm_b200_pe_set_insert_stats(c, pe_sum, pe_count);
		return fail(NGM_B200_ERANGE, "string heap too small: %u bytes needed", used);
	}
	if (used) CU(cudaMemcpy(res->strings, m->d_strings.p, used, cudaMemcpyDeviceToHost));
	return n_reads;
}
