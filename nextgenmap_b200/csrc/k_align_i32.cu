// k_align_i32.cu -- instantiations + launcher of the int32-lane align kernels.
#include "ngm_launch.h"
#include "ngm_kernels.cuh"

namespace ngm {

cudaError_t launch_align_i32(int capacity, int mode, const AlignArgs &a, cudaStream_t st) {
	if (a.n <= 0) return cudaSuccess;
	const dim3 block(128), grid((a.n + 127) / 128);
#define X(W, LO) \
	if (capacity == W) { \
		if (mode == 0) align_i32_kernel<W, LO, 0><<<grid, block, 0, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, \
				a.ptr_scratch, a.ops_scratch, a.stride, a.ops_cap, a.recs, a.strings, a.str_cap, a.cursor, a.out_best); \
		else align_i32_kernel<W, LO, 1><<<grid, block, 0, st>>>(a.P, a.pairs, a.n, a.reads_fwd, a.reads_rev, a.rlen, a.ref4, \
				a.ptr_scratch, a.ops_scratch, a.stride, a.ops_cap, a.recs, a.strings, a.str_cap, a.cursor, a.out_best); \
		return cudaGetLastError(); \
	}
	NGM_BAND_LIST(X)
#undef X
	return cudaErrorInvalidValue;
}

}  // namespace ngm
