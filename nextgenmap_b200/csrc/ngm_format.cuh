// ngm_format.cuh -- device-side CIGAR / MD / NM / identity generation.
//
// Restates SWOclCigar::computeCigarMD (lib/mason/opencl/SWOclCigar.cpp:430-615),
// which the reference runs on the host with one sprintf per op.  Here the thread
// that did the backtrace pops its own RLE op stack (top = first op of the
// alignment) twice: once to measure, once to write into the compact string heap.
#pragma once

#include "ngm_common.cuh"
#include "ngm_dp_i32.cuh"

namespace ngm {

struct FormatOut {
	int cigar_len, md_len, match, mismatch, total, read_index;
};

template <bool WRITE>
__device__ __forceinline__ int put_num(char *dst, int pos, int v) {
	// sprintf("%d") for the non-negative run lengths that occur here
	int digits = v >= 1000 ? 4 : v >= 100 ? 3 : v >= 10 ? 2 : 1;
	if (v >= 10000) digits = v >= 100000 ? 6 : 5;
	if (WRITE) {
		int x = v;
		for (int k = digits - 1; k >= 0; --k) {
			dst[pos + k] = (char) ('0' + x % 10);
			x /= 10;
		}
	}
	return pos + digits;
}

template <bool WRITE>
__device__ __forceinline__ int put_chr(char *dst, int pos, char ch) {
	if (WRITE) dst[pos] = ch;
	return pos + 1;
}

__device__ __forceinline__ char code_char(int code) {
	// inverse of trans for the alphabet NGM's window decoder can produce
	// (SequenceProvider.cpp:87-101,424-439): ACGT, 'x', 'N', NUL
	const unsigned long long tab = 0x3f004e7854474341ull;   // "ACGTxN\0?"
	return (char) (tab >> (8 * code));
}

template <bool WRITE>
__device__ __forceinline__ FormatOut format_cigar_md(const DevParams &P, const PairCtx &c, const uint16_t *ops, int ops_stride,
		const TraceOut &t, char *cig, char *md) {
	FormatOut f;
	int co = 0, mo = 0;
	const char clip = P.hard_clip == 1 ? 'H' : (P.silent_clip != 1 ? 'S' : 0);
	if (t.qstart > 0 && clip) {
		co = put_num<WRITE>(cig, co, t.qstart);
		co = put_chr<WRITE>(cig, co, clip);
	}
	// bsFrom / bsTo of SWOclCigar.cpp:301-320 as codes
	int from = -1, to = -1;
	if (P.acct_alt) {
		if (P.acct_slam) { from = c.dir ? 2 : 1; to = c.dir ? 0 : 3; }
		else { from = c.dir ? 0 : 3; to = c.dir ? 2 : 1; }
	}
	int match = 0, mismatch = 0, total = 0, m_len = 0, eq_len = 0, ref_index = 0, read_index = t.qstart;
	const int64_t wbase = (int64_t) c.sub + t.pos;
	for (int k = t.sp - 1; k >= 1; --k) {
		const int e = ops[(size_t) k * ops_stride];
		const int op = e & 15, length = e >> 4;
		total += length;
		if (op == OP_X) {
			m_len += length;
			if (!P.acct_alt) mismatch += length;
			mo = put_num<WRITE>(md, mo, eq_len);
			for (int x = 0; x < length; ++x) {
				const int fc = code_at(c.wp, wbase + ref_index) & 7;
				if (P.acct_alt) {
					const int rc = code_at(c.rp, read_index) & 7;
					if (rc == from && fc == to) match += 1; else mismatch += 1;
				}
				mo = put_chr<WRITE>(md, mo, code_char(fc));
				ref_index += 1;
				read_index += 1;
			}
			eq_len = 0;
		} else if (op == OP_EQ) {
			match += length;
			m_len += length;
			eq_len += length;
			ref_index += length;
			read_index += length;
		} else if (op == OP_D) {
			if (m_len > 0) {
				co = put_num<WRITE>(cig, co, m_len);
				co = put_chr<WRITE>(cig, co, 'M');
				m_len = 0;
			}
			co = put_num<WRITE>(cig, co, length);
			co = put_chr<WRITE>(cig, co, 'D');
			mo = put_num<WRITE>(md, mo, eq_len);
			eq_len = 0;
			mo = put_chr<WRITE>(md, mo, '^');
			for (int x = 0; x < length; ++x) {
				mo = put_chr<WRITE>(md, mo, code_char(code_at(c.wp, wbase + ref_index) & 7));
				ref_index += 1;
			}
			mismatch += length;
		} else {  // OP_I (the backtrace emits nothing else)
			if (m_len > 0) {
				co = put_num<WRITE>(cig, co, m_len);
				co = put_chr<WRITE>(cig, co, 'M');
				m_len = 0;
			}
			co = put_num<WRITE>(cig, co, length);
			co = put_chr<WRITE>(cig, co, 'I');
			read_index += length;
			mismatch += length;
		}
	}
	mo = put_num<WRITE>(md, mo, eq_len);
	if (m_len > 0) {
		co = put_num<WRITE>(cig, co, m_len);
		co = put_chr<WRITE>(cig, co, 'M');
	}
	if (t.qend > 0 && clip) {
		co = put_num<WRITE>(cig, co, t.qend);
		co = put_chr<WRITE>(cig, co, clip);
	}
	f.cigar_len = co;
	f.md_len = mo;
	f.match = match;
	f.mismatch = mismatch;
	f.total = total;
	f.read_index = read_index;
	return f;
}

}  // namespace ngm
