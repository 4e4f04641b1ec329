// k_score_i32.cu -- instantiations + launcher of the int32-lane score kernels.
#include "ngm_launch.h"
#include "ngm_dp_i32.cuh"

namespace ngm {

int band_capacity(int corridor) {
	if (corridor < 1) return 0;
#define X(W, LO) if (corridor <= W) return W;
	NGM_BAND_LIST(X)
#undef X
	return 0;
}

int ptr_words_for(int capacity) { return (capacity + 7) / 8; }

cudaError_t launch_score_i32(int capacity, int mode, const ScoreArgs &a, cudaStream_t st) {
	if (a.n <= 0) return cudaSuccess;
	const dim3 block(128), grid((a.n + 127) / 128);
#define X(W, LO) \
	if (capacity == W) { \
		if (mode == 0) score_i32_kernel<W, LO, 0><<<grid, block, 0, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, a.out, a.sel, a.n_dev); \
		else score_i32_kernel<W, LO, 1><<<grid, block, 0, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, a.out, a.sel, a.n_dev); \
		return cudaGetLastError(); \
	}
	NGM_BAND_LIST(X)
#undef X
	return cudaErrorInvalidValue;
}

}  // namespace ngm
