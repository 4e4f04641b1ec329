// exp_issue.cu -- MEASUREMENT, not a product path: issue rates of the instructions the DP kernels are made of, alone and interleaved in
// pairs.  Two instructions that share an execution pipe interleave at the rate of one of them; two that do not, at (up to) the sum.  The
// table (scripts/issue_rates.py -> profiles/) is what DESIGN.md's instruction budget of the forward kernel is derived from: which of PRMT,
// LOP3, SHF, VIADD.16x2, VIMNMX, VIMNMX3, VIADDMNMX compete for the 64-lane integer pipe, and what a three-input maximum really costs.
#include "ngm_ctx.h"
#include "ngm_dp_s16.cuh"
#include "../../include/ngm_b200.h"

namespace ngm {

enum IssueOp : int { OP_NONE = -1, OP_VIADDMNMX = 0, OP_IMAD, OP_VIMNMX3, OP_PRMT, OP_VADD2, OP_LOP3, OP_SHF, OP_VIMNMX, OP_VIADDMNMX_U, OP_VIADDMNMX_RELU, OP_IADD3, OP_COUNT };

template <int OP>
__device__ __forceinline__ uint32_t issue_op(uint32_t a, uint32_t b, uint32_t g, uint32_t m) {
	if (OP == OP_VIADDMNMX) return __viaddmax_s16x2(a, g, b);
	if (OP == OP_IMAD) return imad_u32(a, m, b);
	if (OP == OP_VIMNMX3) return __vimax3_s16x2(a, b, g);
	if (OP == OP_PRMT) return prmt(a, b, g);
	if (OP == OP_VADD2) return __vadd2(a, b);
	if (OP == OP_LOP3) {                                            // (a & g) ^ b as one opaque instruction: chains of it must not be merged
		uint32_t d;
		asm volatile("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(d) : "r"(b), "r"(a), "r"(g));
		return d;
	}
	if (OP == OP_SHF) return __funnelshift_r(a, b, g);
	if (OP == OP_VIMNMX) {
		uint32_t d;
		asm volatile("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
		return d;
	}
	if (OP == OP_VIADDMNMX_U) return __viaddmax_u16x2(a, g, b);
	if (OP == OP_VIADDMNMX_RELU) return __viaddmax_s16x2_relu(a, g, b);
	if (OP == OP_IADD3) {
		uint32_t d;
		asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
		return d;
	}
	return a;
}

// eight independent chains per thread; A updates a[] from (a, b), B updates b[] from (b, a)
template <int A, int B, bool DEP = false>
__global__ void __launch_bounds__(256) issue_rate_kernel(uint32_t *__restrict__ out, int iters, uint32_t g, uint32_t m) {
	uint32_t a[8], b[8];
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		a[k] = threadIdx.x * 4u + k * 64u;
		b[k] = blockIdx.x + k;
	}
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int u = 0; u < 8; ++u) {
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				a[k] = issue_op<A>(a[k], b[k], g, m);
				// DEP: B reads the value A just wrote, so that two consecutive A's of a chain cannot be fused into one instruction
				if (B != OP_NONE) b[k] = issue_op<B>(b[k], a[DEP ? k : (k + 3) & 7], g, m);
			}
		}
	}
	uint32_t r = 0;
#pragma unroll
	for (int k = 0; k < 8; ++k) r ^= a[k] ^ b[k];
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// The forward kernel's per-slot instruction mix in three spellings of the same recurrence (independent chains, no memory traffic):
//   MIX 0  as built: VIADD.16x2 (d) + VIADDMNMX (u) + VIADDMNMX.RELU (h) + LOP3 (clean) + PRMT (substitution) + 2 IMAD (tag, pointer word)
//   MIX 1  u de-fused: VIADD.16x2 x2 + VIMNMX + VIADDMNMX.RELU + LOP3 + PRMT + 2 IMAD
//   MIX 2  all de-fused: VIADD.16x2 x3 + VIMNMX x3 + LOP3 + PRMT + 2 IMAD
//   MIX 3  MIX 0 with the band maximum (VIMNMX3 per two slots);  MIX 4  MIX 0 with the band maximum as two VIMNMX
// slots per second = thread-level "slot" evaluations; x2 cells.
__device__ __forceinline__ uint32_t vmax2_op(uint32_t a, uint32_t b) {
	uint32_t d;
	asm volatile("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
	return d;
}
template <int MIX>
__global__ void __launch_bounds__(128) issue_mix_kernel(uint32_t *__restrict__ out, int iters, uint32_t g, uint32_t m) {
	uint32_t line[9], pw = 0, left = threadIdx.x, acc = 0, sa = blockIdx.x * 0x01010101u + threadIdx.x;
#pragma unroll
	for (int k = 0; k < 9; ++k) line[k] = threadIdx.x * 4u + k * 64u;
	const uint32_t c_four = m & ~1u, c_neg1 = ~(m >> 3);
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int u8 = 0; u8 < 8; ++u8) {
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				const uint32_t s2 = prmt(sa, left, 0x9180u + j);                 // stands for sbyte2<i>(sa, sb)
				uint32_t h;
				if (MIX == 0 || MIX == 3 || MIX == 4) {
					const uint32_t d = __vadd2(line[j], s2);
					const uint32_t u = __viaddmax_s16x2(line[j + 1], g, d);
					h = __viaddmax_s16x2_relu(left, g, u);
				} else if (MIX == 1) {
					const uint32_t d = __vadd2(line[j], s2);
					const uint32_t up = __vadd2(line[j + 1], g);
					const uint32_t u = vmax2_op(up, d);
					h = __viaddmax_s16x2_relu(left, g, u);
				} else {
					const uint32_t d = __vadd2(line[j], s2);
					const uint32_t up = __vadd2(line[j + 1], g);
					const uint32_t lf = __vadd2(left, g);
					h = vmax2_op(vmax2_op(vmax2_op(up, d), lf), 0u);
				}
				const uint32_t clean = h & 0xFFFCFFFCu;
				pw = imad_u32(pw, c_four, imad_u32(clean, c_neg1, h));
				left = clean;
				line[j] = clean;
				if (MIX == 3 && (j & 1)) acc = __vimax3_s16x2(acc, line[j - 1], line[j]);
				if (MIX == 4) acc = vmax2_op(acc, clean);
			}
			sa = sa * 5u + pw;
		}
	}
	uint32_t r = pw ^ acc ^ sa;
#pragma unroll
	for (int k = 0; k < 9; ++k) r ^= line[k];
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

typedef void (*IssueKernel)(uint32_t *, int, uint32_t, uint32_t);
struct IssueCase {
	int a, b;
	IssueKernel k;
};
#define CASE(A, B) { A, B, issue_rate_kernel<A, B> }
static const IssueCase kIssueCases[] = {
	CASE(OP_VIADDMNMX, OP_NONE), CASE(OP_IMAD, OP_NONE), CASE(OP_VIMNMX3, OP_NONE), CASE(OP_PRMT, OP_NONE), CASE(OP_VADD2, OP_NONE), CASE(OP_LOP3, OP_NONE),
	CASE(OP_SHF, OP_NONE), CASE(OP_VIMNMX, OP_NONE), CASE(OP_VIADDMNMX_U, OP_NONE), CASE(OP_VIADDMNMX_RELU, OP_NONE), CASE(OP_IADD3, OP_NONE),
	CASE(OP_VIADDMNMX, OP_IMAD), CASE(OP_VIADDMNMX, OP_PRMT), CASE(OP_VIADDMNMX, OP_LOP3), CASE(OP_VIADDMNMX, OP_SHF), CASE(OP_VIADDMNMX, OP_VADD2),
	CASE(OP_VIADDMNMX, OP_VIMNMX3), CASE(OP_VIADDMNMX, OP_VIMNMX), CASE(OP_VIADDMNMX, OP_IADD3), CASE(OP_PRMT, OP_IMAD), CASE(OP_PRMT, OP_LOP3), CASE(OP_LOP3, OP_IMAD),
	CASE(OP_VIMNMX3, OP_IMAD), CASE(OP_SHF, OP_IMAD), CASE(OP_VADD2, OP_IMAD),
	{ OP_VIMNMX, OP_PRMT, issue_rate_kernel<OP_VIMNMX, OP_PRMT, true> }, { OP_VIMNMX, OP_IMAD, issue_rate_kernel<OP_VIMNMX, OP_IMAD, true> },
	{ OP_VIMNMX, OP_VADD2, issue_rate_kernel<OP_VIMNMX, OP_VADD2, true> }, { OP_VADD2, OP_PRMT, issue_rate_kernel<OP_VADD2, OP_PRMT, true> },
};
#undef CASE

}  // namespace ngm

using namespace ngm;

extern "C" {

// per case: op_a, op_b (-1 = none) and thread-level instructions per second of BOTH kinds together.  Returns the number of cases written.
int ngm_b200_exp_issue_rates(ngm_b200_ctx *c, int cap, int *op_a, int *op_b, double *per_s) {
	if (c == nullptr || op_a == nullptr || op_b == nullptr || per_s == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	CU(cudaSetDevice(c->device));
	int sms = 0;
	CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
	const int blocks = sms * 8, iters = 2048;
	DevBuf out;
	CU(out.ensure((size_t) blocks * 256 * 4));
	cudaEvent_t e0, e1;
	CU(cudaEventCreate(&e0));
	CU(cudaEventCreate(&e1));
	const int n_cases = (int) (sizeof(kIssueCases) / sizeof(kIssueCases[0]));
	int n = 0;
	for (; n < n_cases && n < cap; ++n) {
		float best = 1e30f;
		for (int rep = 0; rep < 4; ++rep) {
			CU(cudaEventRecord(e0, c->stream));
			kIssueCases[n].k<<<blocks, 256, 0, c->stream>>>(out.as<uint32_t>(), iters, 0x00050003u | (c->dp.c_four << 8), c->dp.c_four | 1u);
			CU(cudaEventRecord(e1, c->stream));
			CU(cudaEventSynchronize(e1));
			float ms = 0;
			CU(cudaEventElapsedTime(&ms, e0, e1));
			if (rep) best = std::min(best, ms);
		}
		c->launches += 4;
		op_a[n] = kIssueCases[n].a;
		op_b[n] = kIssueCases[n].b;
		per_s[n] = (double) blocks * 256.0 * iters * 64.0 * (kIssueCases[n].b == OP_NONE ? 1.0 : 2.0) / (best * 1e-3);
	}
	// the mixes: op_a = 100 + MIX, per_s = slot evaluations per second
	const IssueKernel mixes[] = { issue_mix_kernel<0>, issue_mix_kernel<1>, issue_mix_kernel<2>, issue_mix_kernel<3>, issue_mix_kernel<4> };
	for (int mx = 0; mx < 5 && n < cap; ++mx, ++n) {
		float best = 1e30f;
		const int mblocks = sms * 16, miters = 1024;
		CU(out.ensure((size_t) mblocks * 128 * 4));
		for (int rep = 0; rep < 4; ++rep) {
			CU(cudaEventRecord(e0, c->stream));
			mixes[mx]<<<mblocks, 128, 0, c->stream>>>(out.as<uint32_t>(), miters, 0xFF7DFF7Du, c->dp.c_four | 1u);
			CU(cudaEventRecord(e1, c->stream));
			CU(cudaEventSynchronize(e1));
			float ms = 0;
			CU(cudaEventElapsedTime(&ms, e0, e1));
			if (rep) best = std::min(best, ms);
		}
		c->launches += 4;
		op_a[n] = 100 + mx;
		op_b[n] = -1;
		per_s[n] = (double) mblocks * 128.0 * miters * 64.0 / (best * 1e-3);
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	return n;
}

}  // extern "C"
