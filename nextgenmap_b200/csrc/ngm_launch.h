// ngm_launch.h -- host-callable launchers of the templated DP kernels.  Each
// kernel family lives in its own translation unit so nvcc can build them in parallel.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "ngm_common.cuh"
#include "../../include/ngm_b200.h"

namespace ngm {

// Band capacities compiled into the library: X(capacity, smallest corridor served).
// A corridor c is served by the smallest capacity >= c; slots >= c are pinned to the
// sentinel at run time (only slots >= LO carry that select).
#define NGM_BAND_LIST(X) \
	X(12, 1) X(16, 13) X(20, 17) X(24, 21) X(28, 25) X(32, 29) X(36, 33) X(40, 37) X(44, 41) X(48, 45) \
	X(56, 49) X(64, 57) X(72, 65) X(80, 73) X(96, 81) X(112, 97) X(128, 113) X(160, 129)

// (capacity, corridor) pairs with kernels that know the corridor at compile time: the corridors NGM derives from common read lengths
// (5 + 0.15 x length: 125 bp -> 23, 150 bp -> 27, 200 bp -> 35, 250 bp -> 42)
#ifndef NGM_EXACT_LIST
#define NGM_EXACT_LIST(X) X(24, 23) X(28, 27) X(36, 35) X(44, 42)
#endif

constexpr int kMaxCorridor = 160;
// the tagged s16x2 align kernel keeps band + snapshot in registers: beyond these capacities it would spill
constexpr int kAlignS16MaxLocal = 48;
constexpr int kAlignS16MaxKnown = 96;     // local mode with the maximum known in advance (no snapshot)
constexpr int kAlignS16MaxEndFree = 96;

int band_capacity(int corridor);          // 0 if unsupported
int ptr_words_for(int capacity);          // 32-bit pointer words per DP row and alignment to allocate (covers both kernels' layouts)

struct ScoreArgs {
	DevParams P;
	const PairDesc *pairs;
	int n;
	const uint32_t *reads_fwd, *reads_rev;
	const uint16_t *rlen;
	const uint32_t *ref4;
	float *out;
	// optional work list: pair i of the launch is pairs[sel[i]] and its result goes to out[sel[i]]; the number of entries is read on the
	// device (*n_dev <= n) so that a list compacted on the device needs no host round trip.  sel == nullptr: pairs[0..n) in order.
	const int *sel = nullptr;
	const int *n_dev = nullptr;
};

struct AlignArgs {
	DevParams P;
	const PairDesc *pairs;
	int n;
	const uint32_t *reads_fwd, *reads_rev;
	const uint16_t *rlen;
	const uint32_t *ref4;
	uint32_t *ptr_scratch;
	uint16_t *ops_scratch;
	int4 *best_scratch;          // per alignment {best_read, best_ref, best_score, read_count} (s16 path)
	const float *known;          // local maxima of the pairs (score kernel output) or nullptr
	cudaEvent_t ev_mid = nullptr;   // optional: recorded between the forward and the backtrace kernel (ngm_b200_profile)
	// forward pass over every candidate of a chunk of reads (ngm_batch.cu), s16x2 second-generation kernels only:
	int phase = 0;                  // 0 = forward + backtrace of pairs[0..n); 1 = forward only; 2 = backtrace only
	const int *range = nullptr;     // device: the launch covers pairs[range[0] .. range[range_m]) (at most n of them)
	int range_m = 0;
	const int *slot_of = nullptr;   // phase 2: item i (one of n_items reads) is aligned from forward slot slot_of[i], -1 = none
	int n_items = 0;
	const int *items_dev = nullptr; // phase 2, optional (device): only the first min(n_items, *items_dev) items exist
	const int *rec_of = nullptr;    // phase 2, optional (device): item i writes recs[rec_of[i]] instead of recs[i]
	int ops_stride = 0;             // items per op-stack row (0: = stride)
	float *out_best = nullptr;   // optional, per alignment: the forward pass's maximum = what BatchScore returns for the pair in this mode
	int stride, ops_cap;
	ngm_b200_align_rec *recs;
	char *strings;
	uint32_t str_cap;
	uint32_t *cursor;
};

// int32 lanes, one pair per thread
cudaError_t launch_score_i32(int capacity, int mode, const ScoreArgs &a, cudaStream_t st);
cudaError_t launch_align_i32(int capacity, int mode, const AlignArgs &a, cudaStream_t st);
// s16x2 lanes, two pairs per thread
cudaError_t launch_score_s16(int capacity, int mode, const ScoreArgs &a, cudaStream_t st);
cudaError_t launch_align_s16(int capacity, int mode, const AlignArgs &a, cudaStream_t st);

}  // namespace ngm
