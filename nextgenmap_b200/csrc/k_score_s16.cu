// k_score_s16.cu -- instantiations + launcher of the s16x2 two-pairs-per-thread score kernels.
#include "ngm_launch.h"
#include "ngm_dp_s16.cuh"

namespace ngm {

cudaError_t launch_score_s16(int capacity, int mode, const ScoreArgs &a, cudaStream_t st) {
	if (a.n <= 0) return cudaSuccess;
	const int threads = (a.n + 1) / 2;
	const dim3 block(128), grid((threads + 127) / 128);
	// corridors NGM derives from common read lengths get instantiations that know the corridor at compile time (see k_align_s16.cu)
#define XE(W, C) \
	if (capacity == W && a.P.corridor == C) { \
		if (mode == 0) score_s16_kernel<W, C, 0, true><<<grid, block, 0, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, a.out, a.sel, a.n_dev); \
		else score_s16_kernel<W, C, 1, true><<<grid, block, 0, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, a.out, a.sel, a.n_dev); \
		return cudaGetLastError(); \
	}
	NGM_EXACT_LIST(XE)
#undef XE
#define X(W, LO) \
	if (capacity == W) { \
		if (mode == 0) score_s16_kernel<W, LO, 0><<<grid, block, 0, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, a.out, a.sel, a.n_dev); \
		else score_s16_kernel<W, LO, 1><<<grid, block, 0, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, a.out, a.sel, a.n_dev); \
		return cudaGetLastError(); \
	}
	NGM_BAND_LIST(X)
#undef X
	return cudaErrorInvalidValue;
}

}  // namespace ngm
