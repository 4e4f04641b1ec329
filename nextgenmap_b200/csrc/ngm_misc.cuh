// ngm_misc.cuh -- small non-templated kernels: ASCII packers, reference transcoder,
// pair resolution, top-1 selection.  Included by exactly one translation unit (ngm_b200.cu).
#pragma once

#include "ngm_common.cuh"
#include "../../include/ngm_b200.h"

namespace ngm {

// ---------------------------------------------------------------------------
// packers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ascii_code(uint32_t ch) {
	// oclDefines.cl:64-80, branch-free (a switch on per-lane data diverges on every byte)
	const uint32_t u = ch & 0xDFu;                 // 'a'..'z' -> 'A'..'Z'; nothing else maps onto a letter
	uint32_t code = 4;
	code = u == 'A' ? 0u : code;
	code = u == 'C' ? 1u : code;
	code = u == 'G' ? 2u : code;
	code = u == 'T' ? 3u : code;
	code = u == 'N' ? 5u : code;
	code = ch == 0 ? 6u : code;
	return code;
}

// ASCII rows -> packed code words.  One thread per output word; bytes past `width`
// or rows >= src_rows read as NUL.  If rlen32 != nullptr, rlen32[row] = index of the last non-NUL byte + 1;
// if noncanon != nullptr, noncanon[row] = 1 when the row holds a byte outside "ACGTNx\\0".
__global__ void pack_ascii_kernel(const uint8_t *__restrict__ src, int src_rows, int rows, int width, int src_stride, uint32_t *__restrict__ dst,
		int dst_words, unsigned int *__restrict__ rlen32, uint8_t *__restrict__ noncanon) {
	const long long gid = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= (long long) rows * dst_words) return;
	const int row = (int) (gid / dst_words), w = (int) (gid % dst_words);
	const uint8_t *s = src + (size_t) row * src_stride;
	uint32_t word = 0;
	int last = 0;
	bool odd = false;
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		const int i = 8 * w + k;
		const uint32_t ch = (i < width && row < src_rows) ? s[i] : 0u;
		word |= ascii_code(ch) << (4 * k);
		if (ch != 0) last = i + 1;
		// bytes the code -> char table of the formatter cannot reproduce (lower case, IUPAC, ...)
		odd |= !(ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T' || ch == 'N' || ch == 'x' || ch == 0);
	}
	dst[(size_t) row * dst_words + w] = word;
	if (noncanon != nullptr && odd && row < src_rows) noncanon[row] = 1;
	if (rlen32 != nullptr && last > 0) atomicMax(rlen32 + row, (unsigned int) last);
}

__global__ void narrow_rlen_kernel(const unsigned int *__restrict__ rlen32, uint16_t *__restrict__ rlen, int rows) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < rows) rlen[i] = (uint16_t) rlen32[i];
}

// Read upload of the descriptor path (HBM-bound byte shuffling, three small kernels):
//   pack_words_kernel   one thread per 8 bytes: one 64-bit load, eight lookups in a 256-byte shared table
//                       of the reference's trans classes (oclDefines.cl:64-80), one word store
//   read_len_kernel     one thread per read: index of the last non-NUL code + 1
//   revcomp_words_kernel one thread per output word, nibble reversal + complement by bit tricks
//                       (MappedRead::computeReverseSeq, MappedRead.cpp:36-67: only A<->T, C<->G change)
__global__ void __launch_bounds__(256) pack_words_kernel(const uint8_t *__restrict__ src, int rows, int width, int src_stride,
		uint32_t *__restrict__ dst, int words) {
	__shared__ uint8_t s_code[256];
	s_code[threadIdx.x] = (uint8_t) ascii_code(threadIdx.x);
	__syncthreads();
	const long long gid = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= (long long) rows * words) return;
	const int row = (int) (gid / words), w = (int) (gid - (long long) row * words);
	const uint8_t *s = src + (size_t) row * src_stride;
	const int i0 = 8 * w;
	uint32_t word;
	if (((src_stride & 7) == 0) && ((reinterpret_cast<uintptr_t>(src) & 7) == 0) && i0 + 8 <= width) {
		const uint2 v = *reinterpret_cast<const uint2 *>(s + i0);
		word = (uint32_t) s_code[v.x & 0xFF] | (uint32_t) s_code[(v.x >> 8) & 0xFF] << 4 | (uint32_t) s_code[(v.x >> 16) & 0xFF] << 8 |
				(uint32_t) s_code[v.x >> 24] << 12 | (uint32_t) s_code[v.y & 0xFF] << 16 | (uint32_t) s_code[(v.y >> 8) & 0xFF] << 20 |
				(uint32_t) s_code[(v.y >> 16) & 0xFF] << 24 | (uint32_t) s_code[v.y >> 24] << 28;
	} else {
		word = 0;
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const int i = i0 + k;
			word |= (uint32_t) s_code[i < width ? s[i] : 0] << (4 * k);
		}
	}
	dst[gid] = word;
}

__global__ void read_len_kernel(const uint32_t *__restrict__ fwd, int rows, int words, uint16_t *__restrict__ rlen) {
	const int row = blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= rows) return;
	const uint32_t *f = fwd + (size_t) row * words;
	int len = 0;
	int w = words - 1;
	for (; w >= 0; --w) {
		const uint32_t x = f[w] ^ kNulWord;           // non-zero nibble <=> code != NUL
		if (x != 0) {
			len = 8 * w + (31 - __clz(x)) / 4 + 1;
			break;
		}
	}
	rlen[row] = (uint16_t) len;

}

__global__ void __launch_bounds__(256) revcomp_words_kernel(const uint32_t *__restrict__ fwd, const uint16_t *__restrict__ rlen, int rows, int words,
		uint32_t *__restrict__ rev) {
	const long long gid = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= (long long) rows * words) return;
	const int row = (int) (gid / words), w = (int) (gid - (long long) row * words);
	const int hi = (int) rlen[row] - 1 - 8 * w;        // source index of output nibble 0; nibble k reads hi - k
	uint32_t out = kNulWord;
	if (hi >= 0) {
		const uint32_t *f = fwd + (size_t) row * words;
		const int wi = hi >> 3;
		const uint32_t a = f[wi];
		const uint32_t b = wi > 0 ? f[wi - 1] : kNulWord;            // indices < 0 read as NUL
		// 16 nibbles: b = indices 8wi-8 .. 8wi-1, a = 8wi .. 8wi+7; take the 8 ending at `hi`
		const uint32_t x = __funnelshift_rc(b, a, 4 * ((hi & 7) + 1));   // nibble 7 = index hi ... nibble 0 = hi - 7
		uint32_t y = __byte_perm(x, 0, 0x0123);
		y = ((y >> 4) & 0x0F0F0F0Fu) | ((y & 0x0F0F0F0Fu) << 4);       // nibble order reversed: nibble k = index hi - k
		const uint32_t m = (~y >> 2) & 0x11111111u;                    // codes 0..3 (A C G T): complement = code ^ 3
		out = y ^ (m * 3u);
	}
	rev[gid] = out;
}

// The three kernels above in one pass (descriptor path, `set_reads`): one warp per read row.  Every lane packs eight bytes into
// a code word, the row's length falls out of a warp maximum, and the reverse complement is assembled from the row's words in
// shared memory -- the ASCII row is read once and nothing is re-read from HBM (the three-kernel form moves ~4 GB per 10 M reads,
// this one 3.2 GB: 152 B in, 2 x 84 B out).
__device__ __forceinline__ uint32_t revcomp_word(const uint32_t *f, int hi) {
	// output nibble k = complement of source index hi - k; indices < 0 read as NUL (only possible above the row's length: masked by the caller)
	const int wi = hi >> 3;
	const uint32_t a = f[wi];
	const uint32_t b = wi > 0 ? f[wi - 1] : kNulWord;
	const uint32_t x = __funnelshift_rc(b, a, 4 * ((hi & 7) + 1));   // nibble 7 = index hi ... nibble 0 = hi - 7
	uint32_t y = __byte_perm(x, 0, 0x0123);
	y = ((y >> 4) & 0x0F0F0F0Fu) | ((y & 0x0F0F0F0Fu) << 4);           // nibble k = index hi - k
	const uint32_t m = (~y >> 2) & 0x11111111u;                        // codes 0..3 (A C G T): complement = code ^ 3
	return y ^ (m * 3u);
}

// four ASCII bytes -> four code nibbles (bits 0..15).  Fast path for plain A/C/G/T (either case): code = ((b >> 1) ^ (b >> 2)) & 3
// for all four bytes at once, verified by mapping the codes back to letters with one PRMT; anything else takes the per-byte rule.
__device__ __forceinline__ uint32_t ascii4_codes(uint32_t b) {
	const uint32_t c = ((b >> 1) ^ (b >> 2)) & 0x03030303u;
	uint32_t t = (c | (c >> 4)) & 0x00FF00FFu;
	t = (t | (t >> 8)) & 0xFFFFu;
	if ((b & 0xDFDFDFDFu) == prmt(0x54474341u, 0u, t)) return t;             // "ACGT" indexed by code
	if (b == 0u) return 0x6666u;                                              // padding behind the read
	return ascii_code(b & 0xFFu) | ascii_code((b >> 8) & 0xFFu) << 4 | ascii_code((b >> 16) & 0xFFu) << 8 | ascii_code(b >> 24) << 12;
}

constexpr int kPackRowsPerBlock = 16;      // 16 threads per row, each packs 16 bytes into two code words

__global__ void __launch_bounds__(16 * kPackRowsPerBlock) pack_reads_fused_kernel(const uint8_t *__restrict__ src, int rows, int width, int src_stride,
		uint32_t *__restrict__ fwd, uint32_t *__restrict__ rev, uint16_t *__restrict__ rlen, int words) {
	extern __shared__ uint32_t s_rows[];                              // [kPackRowsPerBlock][words]
	const int x = threadIdx.x & 15, y = threadIdx.x >> 4;
	const int row = blockIdx.x * kPackRowsPerBlock + y;
	const bool live = row < rows;
	uint32_t *buf = s_rows + y * words;
	const uint8_t *s = src + (size_t) (live ? row : 0) * src_stride;
	const bool wide = ((src_stride & 15) == 0 || (src_stride & 7) == 0) && ((reinterpret_cast<uintptr_t>(src) & 7) == 0);
	int last = 0;
	for (int w2 = x; 2 * w2 < words; w2 += 16) {                      // code words 2 w2, 2 w2 + 1 = bytes 16 w2 .. 16 w2 + 15
#pragma unroll
		for (int h = 0; h < 2; ++h) {
			const int w = 2 * w2 + h, i0 = 8 * w;
			if (w >= words) break;
			uint32_t word;
			if (live && wide && i0 + 8 <= width) {
				const uint2 v = *reinterpret_cast<const uint2 *>(s + i0);
				word = ascii4_codes(v.x) | ascii4_codes(v.y) << 16;
			} else if (!live || i0 >= width) {
				word = kNulWord;
			} else {
				word = 0;
#pragma unroll
				for (int k = 0; k < 8; ++k) word |= ascii_code((live && i0 + k < width) ? s[i0 + k] : 0u) << (4 * k);
			}
			buf[w] = word;
			if (live) fwd[(size_t) row * words + w] = word;
			const uint32_t nz = word ^ kNulWord;                      // non-zero nibble <=> code != NUL
			if (nz != 0) last = max(last, i0 + (31 - __clz(nz)) / 4 + 1);
		}
	}
#pragma unroll
	for (int d = 8; d > 0; d >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, d, 16));
	__syncwarp();
	if (!live) return;
	if (x == 0) rlen[row] = (uint16_t) last;
	for (int w = x; w < words; w += 16) {
		const int hi = last - 1 - 8 * w;                             // source index of output nibble 0
		rev[(size_t) row * words + w] = hi >= 0 ? revcomp_word(buf, hi) : kNulWord;
	}
}

// strict path: pair i uses read row i and the window packed at word i * win_words
__global__ void strict_pairs_kernel(PairDesc *__restrict__ pairs, const uint8_t *__restrict__ flags, int n, int win_words) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	PairDesc d;
	d.win_nib = (uint64_t) i * (uint64_t) win_words * 8ull;
	d.read_idx = (uint32_t) i;
	d.flags = flags[i];
	pairs[i] = d;
}

// reverse complement of packed reads (MappedRead::computeReverseSeq, MappedRead.cpp:36-67):
// rev[k] = cpl(fwd[len-1-k]) for k < len, NUL beyond; only A<->T and C<->G are complemented.
__global__ void revcomp_kernel(const uint32_t *__restrict__ fwd, const uint16_t *__restrict__ rlen, int rows, int words,
		uint32_t *__restrict__ rev) {
	const long long gid = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= (long long) rows * words) return;
	const int row = (int) (gid / words), w = (int) (gid % words);
	const int len = rlen[row];
	const uint32_t *f = fwd + (size_t) row * words;
	uint32_t word = 0;
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		const int i = 8 * w + k;
		uint32_t code = kCodeNul;
		if (i < len) {
			code = code_at(f, len - 1 - i);
			if (code < 4) code = 3 - code;
		}
		word |= code << (4 * k);
	}
	rev[(size_t) row * words + w] = word;
}

// NGM reference packing (4 bit/base, high nibble first, A0 T1 G2 C3 N4,
// SequenceProvider.cpp:72-109) -> device code words.  Positions >= concat_len become
// 'x' (code 4) like DecodeRefSequence's overhang (SequenceProvider.cpp:427-429); the
// region [n_region, n_region + n_len) is filled with 'N' (ScoreBuffer.cpp:117).
__global__ void transcode_ref_kernel(const uint8_t *__restrict__ packed, unsigned long long concat_len, uint32_t *__restrict__ dst,
		unsigned long long total_words, unsigned long long n_region_word) {
	const unsigned long long w = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= total_words) return;
	uint32_t word = 0;
	if (w >= n_region_word) {
		word = 0x55555555u;
	} else {
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const unsigned long long pos = 8ull * w + k;
			uint32_t code = 4;
			if (pos < concat_len) {
				const uint32_t b = packed[pos >> 1];
				const uint32_t v = (pos & 1) ? (b & 0xF) : (b >> 4);
				code = v == 0 ? 0u : v == 1 ? 3u : v == 2 ? 2u : v == 3 ? 1u : v == 4 ? 5u : 4u;
			}
			word |= code << (4 * k);
		}
	}
	dst[w] = word;
}

// descriptor path: translate ngm_b200_pair::window_start into a nibble index of the
// resident reference; starts >= concat_len (incl. the unsigned underflow of
// loc - corridor/2) select the all-'N' region (ScoreBuffer.cpp:113-118).
__global__ void resolve_pairs_kernel(const ngm_b200_pair *__restrict__ in, PairDesc *__restrict__ out, int n, unsigned long long concat_len,
		unsigned long long n_region_nib, const uint16_t *__restrict__ rlen, unsigned int n_reads) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const ngm_b200_pair p = in[i];
	PairDesc d;
	d.win_nib = p.window_start < concat_len ? p.window_start : n_region_nib;
	d.read_idx = p.read_index < n_reads ? p.read_index : 0u;
	d.flags = p.flags & (PF_REVERSE | PF_DIR | PF_INACTIVE);
	if (p.read_index >= n_reads || rlen[p.read_index] == 0) d.flags |= PF_INACTIVE;      // an empty read is its own quad leader; a row outside the batch is skipped
	out[i] = d;
}

__global__ void gather_winners_kernel(int n_reads, const ngm_b200_pair *__restrict__ pairs, const int *__restrict__ best_pair,
		ngm_b200_pair *__restrict__ out, const float *__restrict__ scores, float *__restrict__ out_scores) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int b = best_pair[r];
	ngm_b200_pair p;
	float s = 0.0f;
	if (b >= 0) {
		p = pairs[b];
		if (scores != nullptr) s = scores[b];
	} else {
		p.window_start = 0;
		p.read_index = (uint32_t) r;
		p.flags = PF_INACTIVE;
	}
	out[r] = p;
	if (out_scores != nullptr) out_scores[r] = s;
}

// Issue-rate microbenchmarks behind ngm_b200_alu_peak: what the roofline of the DP kernels is measured against.
// KIND 0: VIADDMNMX.S16x2 (integer ALU pipe, the DP recurrence's instruction); 1: IMAD (FMA pipe); 2: both interleaved 1:1.
template <int KIND>
__global__ void __launch_bounds__(256) alu_peak_kernel(uint32_t *__restrict__ out, int iters, uint32_t g, uint32_t m) {
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
				if (KIND != 1) a[k] = __viaddmax_s16x2(a[k], g, b[k]);
				if (KIND != 0) b[k] = imad_u32(b[k], m, a[k]);
			}
		}
	}
	uint32_t r = 0;
#pragma unroll
	for (int k = 0; k < 8; ++k) r ^= a[k] ^ b[k];
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// ScoreBuffer::top1SE + computeMQ (ScoreBuffer.cpp:34-40,228-277); one thread per read.
__global__ void select_top1_kernel(int n_reads, const int *__restrict__ cand_begin, const float *__restrict__ scores,
		int *__restrict__ best_pair, int *__restrict__ mapq, int *__restrict__ num_top, int strata) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int b = cand_begin[r], e = cand_begin[r + 1];
	float best = 0.0f, second = 0.0f;
	int besti = 0, nbest = 0;                                  // numBestScore -> MappedRead::numTopScores (SAM NH / X0)
	for (int j = b; j < e; ++j) {
		const float s = scores[j];
		if (s > second) {
			if (s > best) {
				second = best;
				best = s;
				besti = j - b;
				nbest = 1;
			} else if (s == best) {
				++nbest;
				second = best;
			} else {
				second = s;
			}
		} else if (s == best) {
			++nbest;
		}
	}
	int mq = 0;
	if (best > 0.0f && second >= 0.0f) mq = (int) ceilf(60.0f * (best - second) / best);
	int bp = e > b ? b + besti : -1;
	if (strata && nbest != 1 && e > b) {                       // "too many equal scoring positions": unmapped (ScoreBuffer.cpp:259-276)
		bp = -1;
		mq = 0;
		nbest = 1;
	}
	if (num_top != nullptr) num_top[r] = nbest;
	best_pair[r] = bp;
	mapq[r] = mq;
}

}  // namespace ngm
