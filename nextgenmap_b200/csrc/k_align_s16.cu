// k_align_s16.cu -- instantiations + launcher of the tagged s16x2 align path:
// forward kernel (two alignments per thread) + backtrace/format kernel (one alignment per thread).
#include "ngm_launch.h"
#include <cstdlib>

#include "ngm_align_s16.cuh"
#include "ngm_align_s16v2.cuh"

namespace ngm {

cudaError_t launch_align_s16(int capacity, int mode, const AlignArgs &a, cudaStream_t st) {
	if (a.n <= 0) return cudaSuccess;
	const int threads = (a.n + 1) / 2;
	const dim3 block(128), grid((threads + 127) / 128);
	bool launched = a.phase == 2;
	// second-generation forward pass (ngm_align_s16v2.cuh): narrow local bands and every end-free band.  NGM_B200_FWD=1 keeps the
	// first-generation kernel (A/B measurements, tests of both).
	static const bool use_v1 = [] { const char *e = getenv("NGM_B200_FWD"); return e != nullptr && atoi(e) == 1; }();
	// corridors NGM derives from common read lengths (5 + 0.15 x length: 125 bp -> 23, 150 bp -> 27, 200 bp -> 35, 250 bp -> 42) get
	// instantiations that know the corridor at compile time: no slot beyond it is computed, no run-time corridor mask
	if (!use_v1 && !launched) {
#define XE(W, C) \
		if (capacity == W && a.P.corridor == C && !launched) { \
			if constexpr (W <= kAlignS16MaxLocal) { \
				if (mode == 0) { \
					constexpr int smem = 3 * W * 128 * 4; \
					static const cudaError_t attr = cudaFuncSetAttribute(align_s16_fwd2_kernel<W, C, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
					if (attr != cudaSuccess) return attr; \
					align_s16_fwd2_kernel<W, C, 0, true><<<grid, block, smem, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, \
							a.ptr_scratch, a.stride, a.best_scratch, a.range, a.range_m); \
					launched = true; \
				} \
			} \
			if (mode == 1) { \
				align_s16_fwd2_kernel<W, C, 1, true><<<grid, block, 0, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, \
						a.ptr_scratch, a.stride, a.best_scratch, a.range, a.range_m); \
				launched = true; \
			} \
		}
		static const bool no_exact = [] { const char *e = getenv("NGM_B200_FWD_EXACT"); return e != nullptr && atoi(e) == 0; }();
		if (!no_exact) {
			NGM_EXACT_LIST(XE)
		}
#undef XE
	}
	if (!use_v1 && !launched) {
#define X(W, LO) \
		if (capacity == W && !launched) { \
			if constexpr (W <= kAlignS16MaxLocal) { \
				if (mode == 0) { \
					constexpr int smem = 3 * W * 128 * 4; \
					static const cudaError_t attr = cudaFuncSetAttribute(align_s16_fwd2_kernel<W, LO, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
					if (attr != cudaSuccess) return attr; \
					align_s16_fwd2_kernel<W, LO, 0><<<grid, block, smem, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, \
							a.ptr_scratch, a.stride, a.best_scratch, a.range, a.range_m); \
					launched = true; \
				} \
			} else if constexpr (W <= kAlignS16MaxKnown) { \
				/* wide local band whose maxima are not known yet: the second-generation kernel with its checkpoints in local memory */ \
				if (mode == 0 && a.known == nullptr) { \
					align_s16_fwd2_kernel<W, LO, 0><<<grid, block, 0, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, \
							a.ptr_scratch, a.stride, a.best_scratch, a.range, a.range_m); \
					launched = true; \
				} \
			} \
			if constexpr (W <= kAlignS16MaxEndFree) { \
				if (mode == 1) { \
					align_s16_fwd2_kernel<W, LO, 1><<<grid, block, 0, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, \
							a.ptr_scratch, a.stride, a.best_scratch, a.range, a.range_m); \
					launched = true; \
				} \
			} \
		}
		NGM_BAND_LIST(X)
#undef X
	}
#define X(W, LO) \
	if (capacity == W && !launched) { \
		if constexpr (W <= kAlignS16MaxLocal) { \
			if (mode == 0) { \
				align_s16_fwd_kernel<W, LO, 0, false><<<grid, block, 0, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, \
						a.ptr_scratch, a.stride, a.best_scratch, nullptr); \
				launched = true; \
			} \
		} else if constexpr (W <= kAlignS16MaxKnown) { \
			if (mode == 0 && a.known != nullptr) { \
				align_s16_fwd_kernel<W, LO, 0, true><<<grid, block, 0, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, \
						a.ptr_scratch, a.stride, a.best_scratch, a.known); \
				launched = true; \
			} \
		} \
		if constexpr (W <= kAlignS16MaxEndFree) { \
			if (mode == 1) { \
				align_s16_fwd_kernel<W, LO, 1, false><<<grid, block, 0, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, \
						a.ptr_scratch, a.stride, a.best_scratch, nullptr); \
				launched = true; \
			} \
		} \
	}
	NGM_BAND_LIST(X)
#undef X
	if (!launched) return cudaErrorInvalidValue;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return e;
	if (a.phase == 1) return cudaSuccess;
	if (a.ev_mid != nullptr) cudaEventRecord(a.ev_mid, st);
	const int items = a.phase == 2 ? a.n_items : a.n;
	const int ops_stride = a.ops_stride > 0 ? a.ops_stride : a.stride;
	const dim3 b2(256), g2((items + 255) / 256);
	if (mode == 0)
		backtrace_format_kernel<0><<<g2, b2, 0, st>>>(a.P, a.pairs, items, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, a.ptr_scratch, capacity,
				a.best_scratch, a.ops_scratch, a.stride, a.ops_cap, a.recs, a.strings, a.str_cap, a.cursor, a.out_best, a.slot_of, a.range, ops_stride, a.items_dev, a.rec_of);
	else
		backtrace_format_kernel<1><<<g2, b2, 0, st>>>(a.P, a.pairs, items, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, a.ptr_scratch, capacity,
				a.best_scratch, a.ops_scratch, a.stride, a.ops_cap, a.recs, a.strings, a.str_cap, a.cursor, a.out_best, a.slot_of, a.range, ops_stride, a.items_dev, a.rec_of);
	return cudaGetLastError();
}

}  // namespace ngm
