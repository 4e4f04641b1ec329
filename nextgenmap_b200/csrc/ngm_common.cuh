// ngm_common.cuh -- shared device-side definitions of the B200 alignment backend.
//
// Sequence representation on the device
// -------------------------------------
// Every base is a 4-bit code, eight codes per 32-bit word, code k of a word in
// bits [4k, 4k+4).  Codes follow the reference's `trans` table
// (lib/mason/opencl/opencl/oclDefines.cl:64-80):
//     A 0, C 1, G 2, T 3, "other" ('x', IUPAC, ...) 4, N 5, NUL 6.
// The substitution score of (read code rc, ref code fc) comes from a byte LUT:
// for one DP row rc is fixed, so the row's eight possible scores (fc = 0..7) sit in
// one uint2 and a single PRMT, using four window nibbles as its selector, yields
// the scores of four adjacent cells.  This reproduces the reference's 7x7 score
// matrices (oclDefines.cl:85-128) for ACGT, N, 'x' and NUL without branches.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace ngm {

constexpr int kCodeNul = 6;
constexpr uint32_t kNulWord = 0x66666666u;
constexpr int kEndFreeMin = -16000;  // oclDefines.cl:28 short_min

// pointer codes written by the forward pass (2 bits per cell)
enum : uint32_t { PTR_DIAG = 0, PTR_UP = 1, PTR_LEFT = 2 };
// CIGAR op codes of the reference (SWOclCigar.cpp:35)
enum : int { OP_I = 1, OP_D = 2, OP_S = 4, OP_EQ = 7, OP_X = 8 };

// pair flags (ngm_b200_pair::flags + internal bits)
enum : uint32_t { PF_REVERSE = 1u, PF_DIR = 2u, PF_INACTIVE = 4u };

struct PairDesc {          // == ngm_b200_pair, window_start already translated to a nibble index
	uint64_t win_nib;
	uint32_t read_idx;
	uint32_t flags;
};

struct DevParams {
	int qml;               // qry_max_len
	int corridor;
	int gap_read;          // negative
	int gap_ref;           // negative
	int match;             // match score, EQ test of the non-ALT kernels (oclSwScore.cl:64)
	int mismatch;          // -mismatch_penalty: score of every non-EQ diag step between plain A/C/G/T codes
	int alt;               // ALT scoring compiled in: EQ iff codes equal (oclSwScore.cl:69)
	int acct_alt;          // bs_mapping == 1 || slam_seq != 0 : X-op accounting (SWOclCigar.cpp:500-514)
	int acct_slam;         // slam_seq != 0 (bsFrom/bsTo choice, SWOclCigar.cpp:312-320)
	int hard_clip;
	int silent_clip;
	int read_words;        // words per packed read row
	int rows_cap;          // pointer rows per alignment in the scratch matrix
	uint32_t c_four;       // 4 and -1 as run-time values: multipliers of the tag IMADs, which ptxas must not fold back onto the ALU pipe
	uint32_t c_neg1;
	uint2 lut[16];         // [dir * 8 + rc] -> signed score bytes for fc = 0..7
	uint2 lut4[16];        // same rows holding 4*S + 2 (diag candidate of the tagged s16x2 align kernel)
};

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
	return d;
}

// a * b + c on the FMA pipe (IMAD); with b taken from the kernel parameters ptxas cannot turn it into LEA / LOP3 / IADD3 (ALU pipe)
__device__ __forceinline__ uint32_t imad_u32(uint32_t a, uint32_t b, uint32_t c) {
	uint32_t d;
	asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}

// byte I of `w`, sign-extended to 32 bits, in one PRMT (selector nibble msb = replicate sign)
template <int I>
__device__ __forceinline__ int sbyte(uint32_t w) {
	constexpr uint32_t sel = (uint32_t) I | ((8u | I) << 4) | ((8u | I) << 8) | ((8u | I) << 12);
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(0u), "n"(sel));
	return (int) d;
}

__device__ __forceinline__ int lut_score(const uint2 *lut, int dir, int rc, int fc) {
	uint2 t = lut[dir * 8 + rc];
	uint32_t w = fc < 4 ? t.x : t.y;
	return (int) (int8_t) (w >> (8 * (fc & 3)));
}

__device__ __forceinline__ int code_at(const uint32_t *__restrict__ words, int64_t nib) {
	return (int) ((words[nib >> 3] >> (4 * (int) (nib & 7))) & 0xF);
}

}  // namespace ngm
