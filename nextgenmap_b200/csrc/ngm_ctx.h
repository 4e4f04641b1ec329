// ngm_ctx.h -- internal: the context behind the C ABI and small host helpers shared by the translation units
// that implement it (ngm_b200.cu: alignment path, ngm_cs.cu: candidate search).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "ngm_common.cuh"
#include "../../include/ngm_b200.h"

namespace ngm {

int fail(int code, const char *fmt, ...);      // records the message for ngm_b200_last_error() and returns `code`

#define CU(call) \
	do { \
		cudaError_t e_ = (call); \
		if (e_ != cudaSuccess) return ngm::fail(NGM_B200_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	bool borrowed = false;     // alias of another context's buffer (lanes share the reference and the index): never freed or resized here
	DevBuf() = default;
	DevBuf(const DevBuf &) = delete;
	DevBuf &operator=(const DevBuf &) = delete;
	DevBuf(DevBuf &&o) noexcept : p(o.p), cap(o.cap), borrowed(o.borrowed) { o.p = nullptr; o.cap = 0; o.borrowed = false; }
	DevBuf &operator=(DevBuf &&o) noexcept {
		if (this != &o) {
			release();
			p = o.p; cap = o.cap; borrowed = o.borrowed;
			o.p = nullptr; o.cap = 0; o.borrowed = false;
		}
		return *this;
	}
	~DevBuf() { release(); }   // temporaries of the entry points are freed on every early return (CU()), too
	void borrow(const DevBuf &o) {
		release();
		p = o.p;
		cap = o.cap;
		borrowed = true;
	}
	cudaError_t ensure(size_t bytes) {
		if (bytes <= cap) return cudaSuccess;
		if (borrowed) return cudaErrorInvalidValue;
		if (p) cudaFree(p);
		p = nullptr;
		cap = 0;
		size_t want = bytes + bytes / 8 + 256;
		cudaError_t e = cudaMalloc(&p, want);
		if (e == cudaSuccess) cap = want;
		return e;
	}
	void release() {
		if (p && !borrowed) cudaFree(p);
		p = nullptr;
		cap = 0;
		borrowed = false;
	}
	template <typename T> T *as() const { return static_cast<T *>(p); }
};

struct HostBuf {   // pinned
	void *p = nullptr;
	size_t cap = 0;
	HostBuf() = default;
	HostBuf(const HostBuf &) = delete;
	HostBuf &operator=(const HostBuf &) = delete;
	~HostBuf() { release(); }
	cudaError_t ensure(size_t bytes) {
		if (bytes <= cap) return cudaSuccess;
		if (p) cudaFreeHost(p);
		p = nullptr;
		cap = 0;
		size_t want = bytes + bytes / 8 + 256;
		cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
		if (e == cudaSuccess) cap = want;
		return e;
	}
	void release() {
		if (p) cudaFreeHost(p);
		p = nullptr;
		cap = 0;
	}
	template <typename T> T *as() const { return static_cast<T *>(p); }
};

struct ScoreArgs;
int mode_of(int mode);                          // mode & 0xFF -> 0 / 1, or -1
// (implemented in ngm_b200.cu; used by the batch engine in ngm_batch.cu)
int resolve_pairs_for(ngm_b200_ctx *c, const void *d_pairs_user, int n, cudaStream_t st);      // ngm_b200_pair -> c->d_rpairs (PairDesc)
int pack_reads_device(ngm_b200_ctx *c, const uint8_t *d_ascii, int n_reads, int stride, cudaStream_t st);
int run_score(ngm_b200_ctx *c, int mode, const ScoreArgs &a, cudaStream_t st);
bool wide_v2_enabled();                         // wide local bands without a score pass first (NGM_B200_WIDE_V2, default on)
ScoreArgs score_args(ngm_b200_ctx *c, const PairDesc *pairs, int n, const uint32_t *rf, const uint32_t *rr, const uint16_t *rl, const uint32_t *ref4, float *out);
int run_align(ngm_b200_ctx *c, int mode, const PairDesc *pairs, int n, const uint32_t *rf, const uint32_t *rr, const uint16_t *rl, const uint32_t *ref4,
		ngm_b200_align_rec *recs, char *strings, uint32_t str_cap, uint32_t *cursor, cudaStream_t st, const float *known_user, float *out_best);

struct CsState;                                 // candidate search: index + scratch (ngm_cs.cu)
void cs_release(CsState *cs);
struct PeState;                                 // paired-end selection: parameters, running insert-size sums, scratch (ngm_select.cu)
void pe_release(PeState *pe);
struct BatchState;                              // lanes + staging of ngm_b200_run_batch (ngm_batch.cu)
void batch_release(BatchState *b);
uint64_t batch_lane_launches(const ngm_b200_ctx *c);
int batch_lanes(ngm_b200_ctx *c, ngm_b200_ctx ***lanes, int *n_lanes, int *sub_batch);      // created / re-synchronised on demand
int cs_share_index(ngm_b200_ctx *lane, const ngm_b200_ctx *root);      // lane->cs aliases root's index (own search scratch)
int pe_share_state(ngm_b200_ctx *lane, const ngm_b200_ctx *root);      // lane->pe: root's parameters + an alias of its running insert-size sums

}  // namespace ngm

struct ngm_b200_ctx {
	ngm_b200_params hp;
	ngm::DevParams dp;
	int device = 0;
	int capacity = 0;          // band capacity W
	int use_s16 = 0;           // s16x2 lanes allowed for the score kernels
	int align_s16[2] = {0, 0}; // tagged s16x2 align kernel usable in local / end-free mode
	int score_batch = 0, align_batch = 0;
	int strict_chunk = 0;      // pairs per strict-path launch
	int align_chunk = 0;       // alignments per launch (bounded by scratch memory)
	int win_words = 0;         // strict path: packed words per window
	int ref_width = 0;         // qml + corridor bytes copied per window (SWOcl.cpp:546)
	cudaStream_t stream = nullptr;
	uint64_t launches = 0;
	// strict-path staging
	ngm::HostBuf h_reads, h_refs, h_flags, h_scores, h_recs, h_strings, h_cursor, h_noncanon;
	ngm::DevBuf d_areads, d_arefs, d_flags, d_reads4, d_rlen32, d_rlen, d_wins4, d_pairs, d_scores, d_recs, d_strings, d_cursor, d_noncanon;
	// align scratch
	ngm::DevBuf d_ptr, d_ops, d_best, d_known;
	// descriptor path
	ngm::DevBuf d_ref4, d_rfwd, d_rrev, d_rrlen32, d_rrlen, d_rascii, d_upairs, d_rpairs;
	uint64_t concat_len = 0, n_region_nib = 0;
	int n_reads = 0;
	int reads_stride = 0;      // bytes per row of d_rascii (set_reads)
	bool have_ref = false;
	ngm::CsState *cs = nullptr;
	ngm::PeState *pe = nullptr;
	ngm::BatchState *batch = nullptr;
	int profile = 0;           // ngm_b200_profile: time the forward / backtrace kernels of the align launch sets with events
	cudaEvent_t pev[3 * 64] = {};
	int pev_used = 0;
	int se_strata = 0;         // "strata" for single-end top-1 selection (ScoreBuffer.cpp:259)
	int se_topn = 1;           // "topn": alignments reported per single-end read (ScoreBuffer.cpp:279-330); batch entry points
	uint64_t epoch = 0;        // bumped whenever the reference / index / selection parameters change: lanes re-sync their aliases
	ngm_b200_ctx *root = nullptr;      // set in a lane / shared context: the context whose resident data it borrows
	uint64_t root_epoch = ~0ull;       // the root's epoch at the last borrow
};
