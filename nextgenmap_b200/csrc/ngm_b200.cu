// ngm_b200.cu -- host side of the C ABI declared in include/ngm_b200.h.
//
// The strict entry points mirror what SWOcl::BatchScore / SWOclCigar::BatchAlign do on the
// host (lib/mason/opencl/SWOcl.cpp:33-162,409-452,534-556; SWOclCigar.cpp:52-102,104-370):
// gather the caller's char** rows into pinned staging, copy, launch, copy back.  The
// difference is what travels and where it is processed: ASCII is packed to 4-bit codes
// on the device, the pointer matrix never leaves the GPU, and CIGAR/MD come back as
// compact strings.  There is no CPU implementation of the DP in this library.
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "ngm_ctx.h"
#include "ngm_launch.h"
#include "ngm_misc.cuh"

using namespace ngm;

namespace ngm {

static thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}

const char *last_error_text() { return g_err; }

}  // namespace ngm

namespace {

bool is_int(float v) { return std::floor(v) == v && std::fabs(v) < 1e6f; }

}  // namespace

namespace {

// Build the row LUTs from the reference's 7x7 score matrices (oclDefines.cl:85-128).
int build_params(const ngm_b200_params &hp, DevParams &dp, int &use_s16, int *align_s16) {
	if (hp.qry_max_len < 1 || hp.qry_max_len > 4000) return fail(NGM_B200_EINVAL, "qry_max_len %d out of range", hp.qry_max_len);
	if (hp.corridor < 1 || hp.corridor > kMaxCorridor) return fail(NGM_B200_EINVAL, "corridor %d not in [1, %d]", hp.corridor, kMaxCorridor);
	const float fl[6] = { hp.match_bonus, hp.mismatch_penalty, hp.gap_read_penalty, hp.gap_ref_penalty, hp.match_bonus_tt, hp.match_bonus_tc };
	for (float v : fl)
		if (!is_int(v)) return fail(NGM_B200_ERANGE, "scoring parameters must be integer valued for bit-exact integer DP (got %g)", v);
	const int match = (int) hp.match_bonus, mism = -(int) hp.mismatch_penalty;
	const int alt = hp.bs_mapping == 1 ? 1 : ((hp.slam_seq & 2) ? 2 : 0);   // SWOcl.cpp:228-242
	const int match_alt = (int) hp.match_bonus_tt;
	const int mism_alt = alt == 1 ? (int) hp.match_bonus_tc : -(int) hp.match_bonus_tc;
	int tab[2][8][8];
	for (int d = 0; d < 2; ++d)
		for (int r = 0; r < 8; ++r)
			for (int c = 0; c < 8; ++c) {
				int v;
				if (r >= 6 || c >= 7) v = 0;                 // read NUL row; code 7 never occurs
				else if (r == 5) v = c < 4 ? 0 : mism;       // read N
				else if (c == 6) v = 0;                      // ref NUL column
				else v = (r == c && r < 4) ? match : mism;
				tab[d][r][c] = v;
			}
	if (alt == 1) {
		tab[0][3][1] = mism_alt; tab[0][3][3] = match_alt;
		tab[1][0][0] = match_alt; tab[1][0][2] = mism_alt;
	} else if (alt == 2) {
		tab[0][1][3] = mism_alt; tab[0][3][3] = match_alt;
		tab[1][0][0] = match_alt; tab[1][2][0] = mism_alt;
	}
	if (alt) for (int c = 0; c < 8; ++c) tab[1][5][c] = 0;   // REV matrices: read-N row is all zero
	int smin = 0, smax = 0;
	for (int d = 0; d < 2; ++d)
		for (int r = 0; r < 8; ++r)
			for (int c = 0; c < 8; ++c) {
				smin = std::min(smin, tab[d][r][c]);
				smax = std::max(smax, tab[d][r][c]);
			}
	if (smin < -127 || smax > 127) return fail(NGM_B200_ERANGE, "substitution scores must fit a signed byte (range %d..%d)", smin, smax);
	// The backtrace recomputes STOP cells from the path score, which needs the reference's
	// odd `max == line[j] + mismatch` clause (oclSwScore.cl:83) to be unreachable: true iff no
	// substitution score is below `mismatch`.
	if (smin < mism) return fail(NGM_B200_ERANGE, "a substitution score (%d) below -mismatch_penalty (%d) is not supported", smin, mism);
	const int gr = -(int) hp.gap_read_penalty, gf = -(int) hp.gap_ref_penalty;
	if (gr > 0 || gf > 0 || gr < -4000 || gf < -4000) return fail(NGM_B200_ERANGE, "gap penalties must be in [0, 4000]");
	if (match < 0) return fail(NGM_B200_ERANGE, "match_bonus must be >= 0");
	const long long span = (long long) hp.qry_max_len + 8;
	if (span * std::max(smax, 1) > 8000000LL || span * std::max(-smin, 1) > 8000000LL) return fail(NGM_B200_ERANGE, "score range too large");
	memset(&dp, 0, sizeof(dp));
	dp.qml = hp.qry_max_len;
	dp.corridor = hp.corridor;
	dp.gap_read = gr;
	dp.gap_ref = gf;
	dp.match = match;
	dp.mismatch = mism;
	dp.alt = alt != 0;
	dp.acct_alt = (hp.bs_mapping == 1 || hp.slam_seq != 0) ? 1 : 0;
	dp.acct_slam = hp.slam_seq != 0;
	dp.hard_clip = hp.hard_clip;
	dp.silent_clip = hp.silent_clip;
	dp.c_four = 4u;
	dp.c_neg1 = 0xFFFFFFFFu;
	dp.read_words = (hp.qry_max_len + 14) / 8 + 1;
	dp.rows_cap = 8 * ((hp.qry_max_len + 14) / 8);
	for (int d = 0; d < 2; ++d)
		for (int r = 0; r < 8; ++r) {
			uint32_t lo = 0, hi = 0;
			for (int c = 0; c < 4; ++c) {
				lo |= (uint32_t) (uint8_t) (int8_t) tab[d][r][c] << (8 * c);
				hi |= (uint32_t) (uint8_t) (int8_t) tab[d][r][c + 4] << (8 * c);
			}
			dp.lut[d * 8 + r] = make_uint2(lo, hi);
		}
	// s16x2 lanes are exact while every cell value stays inside int16: local scores are in
	// [0, smax*qml]; end-free values are bounded below by the -16000 sentinel plus one gap step
	// as long as a full-length mismatch path stays above the sentinel.
	use_s16 = (span * std::max(smax, 1) <= 30000) && (span * std::max(-smin, 1) <= 15000) && gr >= -700 && gf >= -700;
	// tagged align kernel: scores scaled by 4 must fit int16 and 4*S+2 must fit int8; in end-free mode a
	// full-length worst-case path must stay above its (smaller) sentinel so that the sentinel never wins
	const bool lut4_ok = 4 * smin + 2 >= -128 && 4 * smax + 3 <= 127 && gr >= -250 && gf >= -250;
	align_s16[0] = lut4_ok && 4 * span * std::max(smax, 1) <= 32000;
	align_s16[1] = lut4_ok && 4 * span * std::max(smax, 1) <= 30000 && span * std::max(-smin, 1) <= 7500;
	for (int d = 0; d < 2; ++d)
		for (int r = 0; r < 8; ++r) {
			uint32_t lo = 0, hi = 0;
			for (int c = 0; c < 4; ++c) {
				// diag candidate: 4*S + 2, +1 more when the cell is an EQ op (oclSwScore.cl:64,69)
				const int e0 = alt ? (r == c && r < 7) : (tab[d][r][c] == match);
				const int e1 = alt ? (r == c + 4 && r < 7) : (tab[d][r][c + 4] == match);
				lo |= (uint32_t) (uint8_t) (int8_t) (lut4_ok ? 4 * tab[d][r][c] + 2 + e0 : 0) << (8 * c);
				hi |= (uint32_t) (uint8_t) (int8_t) (lut4_ok ? 4 * tab[d][r][c + 4] + 2 + e1 : 0) << (8 * c);
			}
			dp.lut4[d * 8 + r] = make_uint2(lo, hi);
		}
	return NGM_B200_OK;
}

// SWOclCigar::computeCigarMD's X-op accounting under bs_mapping / slam_seq compares the
// caller's raw bytes (SWOclCigar.cpp:507-514).  The device works on codes, which is the same
// thing for the alphabet NGM produces; for rows flagged non-canonical redo the count on the host
// from the finished CIGAR / MD.  Returns match count; *mism and *total are updated.
void recount_alt(const char *cigar, const char *md, int md_len, const char *ref, const char *qry, int qstart, char bs_from, char bs_to,
		int *match_out, int *mism_out, int *total_out) {
	int match = 0, mism = 0, total = 0, read_index = qstart, ref_index = 0, mk = 0;
	int pending_eq = 0;      // unread part of the current MD number
	for (const char *p = cigar; *p;) {
		int len = 0;
		while (*p >= '0' && *p <= '9') len = len * 10 + (*p++ - '0');
		const char op = *p++;
		if (op == 'S' || op == 'H') continue;
		total += len;
		if (op == 'I') {
			read_index += len;
			mism += len;
		} else if (op == 'D') {
			// "<n>^<letters>": the number in front of '^' is the (possibly zero) rest of an EQ run
			while (mk < md_len && md[mk] != '^') ++mk;
			++mk;
			mk += len;
			pending_eq = 0;
			ref_index += len;
			mism += len;
		} else {  // 'M': EQ runs and X letters
			int left = len;
			while (left > 0) {
				if (pending_eq == 0 && mk < md_len && md[mk] >= '0' && md[mk] <= '9') {
					int v = 0;
					while (mk < md_len && md[mk] >= '0' && md[mk] <= '9') v = v * 10 + (md[mk++] - '0');
					pending_eq = v;
				}
				if (pending_eq > 0) {
					const int take = pending_eq < left ? pending_eq : left;
					match += take;
					read_index += take;
					ref_index += take;
					pending_eq -= take;
					left -= take;
				} else {   // an X column: one raw ref byte in the MD
					if (qry[read_index] == bs_from && ref[ref_index] == bs_to) match += 1; else mism += 1;
					++mk;
					read_index += 1;
					ref_index += 1;
					left -= 1;
				}
			}
		}
	}
	*match_out = match;
	*mism_out = mism;
	*total_out = total;
}

}  // namespace

namespace ngm {

int mode_of(int mode) {
	const int m = mode & 0xFF;
	return (m == 0 || m == 1) ? m : -1;
}

ScoreArgs score_args(ngm_b200_ctx *c, const PairDesc *pairs, int n, const uint32_t *rf, const uint32_t *rr, const uint16_t *rl,
		const uint32_t *ref4, float *out) {
	ScoreArgs a;
	a.P = c->dp;
	a.pairs = pairs;
	a.n = n;
	a.reads_fwd = rf;
	a.reads_rev = rr;
	a.rlen = rl;
	a.ref4 = ref4;
	a.out = out;
	return a;
}

int run_score(ngm_b200_ctx *c, int mode, const ScoreArgs &a, cudaStream_t st) {
	cudaError_t e;
#ifdef NGM_HAVE_S16
	if (c->use_s16) e = launch_score_s16(c->capacity, mode, a, st);
	else
#endif
	e = launch_score_i32(c->capacity, mode, a, st);
	c->launches += 1;
	if (e != cudaSuccess) return fail(NGM_B200_ECUDA, "score kernel launch: %s", cudaGetErrorString(e));
	return NGM_B200_OK;
}

int ensure_align_scratch(ngm_b200_ctx *c, int stride) {
	const size_t pw = (size_t) ptr_words_for(c->capacity);
	CU(c->d_ptr.ensure((size_t) c->dp.rows_cap * stride * pw * sizeof(uint32_t)));
	const size_t ops_cap = 2 * (size_t) c->dp.qml + c->dp.corridor + 2;
	CU(c->d_ops.ensure(ops_cap * stride * sizeof(uint16_t)));
	CU(c->d_best.ensure((size_t) stride * sizeof(int4)));
	CU(c->d_known.ensure((size_t) stride * sizeof(float)));
	return NGM_B200_OK;
}

// Launch the align kernel over n resolved pairs in slices of align_chunk.
bool wide_v2_enabled() {
	static const bool on = [] { const char *e = getenv("NGM_B200_WIDE_V2"); return e == nullptr || atoi(e) != 0; }();
	return on;
}

int run_align(ngm_b200_ctx *c, int mode, const PairDesc *pairs, int n, const uint32_t *rf, const uint32_t *rr, const uint16_t *rl,
		const uint32_t *ref4, ngm_b200_align_rec *recs, char *strings, uint32_t str_cap, uint32_t *cursor, cudaStream_t st,
		const float *known_user, float *out_best) {
	const int stride = std::min(std::max(n, 1), c->align_chunk);
	const int stride_pad = (stride + 127) / 128 * 128;
	int rc = ensure_align_scratch(c, stride_pad);
	if (rc) return rc;
	for (int s = 0; s < n; s += stride) {
		AlignArgs a;
		a.P = c->dp;
		a.pairs = pairs + s;
		a.n = std::min(stride, n - s);
		a.reads_fwd = rf;
		a.reads_rev = rr;
		a.rlen = rl;
		a.ref4 = ref4;
		a.ptr_scratch = c->d_ptr.as<uint32_t>();
		a.ops_scratch = c->d_ops.as<uint16_t>();
		a.best_scratch = c->d_best.as<int4>();
		a.known = nullptr;
		a.out_best = out_best != nullptr ? out_best + s : nullptr;
		if (mode == 0 && c->align_s16[0] && c->capacity > kAlignS16MaxLocal && known_user != nullptr) {
			// wide band and the caller already holds the pairs' local maxima (the resident pipeline: BatchScore ran first).
			// Narrow bands keep the snapshot kernel: measured faster there (19.5 vs 17.8 ms per 10 M x 150 bp, profiles/r1b).
			a.known = known_user + s;
		} else if (mode == 0 && c->align_s16[0] && c->capacity > kAlignS16MaxLocal && !wide_v2_enabled()) {
			// wide band (NGM_B200_WIDE_V2=0): get the local maxima from the (cheap) score kernel first, then run the snapshot-free forward pass;
			// the default is the second-generation kernel with local-memory checkpoints, which needs no score pass
			ScoreArgs sa = score_args(c, a.pairs, a.n, rf, rr, rl, ref4, c->d_known.as<float>());
			int rc2 = run_score(c, 0, sa, st);
			if (rc2) return rc2;
			a.known = c->d_known.as<float>();
		}
		const bool prof = c->profile && c->align_s16[mode] && c->pev_used + 3 <= 3 * 64;
		if (prof) {
			for (int k = 0; k < 3; ++k)
				if (c->pev[c->pev_used + k] == nullptr) CU(cudaEventCreate(&c->pev[c->pev_used + k]));
			CU(cudaEventRecord(c->pev[c->pev_used], st));
			a.ev_mid = c->pev[c->pev_used + 1];
		}
		a.stride = stride_pad;
		a.ops_cap = 2 * c->dp.qml + c->dp.corridor + 2;
		a.recs = recs + s;
		a.strings = strings;
		a.str_cap = str_cap;
		a.cursor = cursor;
		cudaError_t e;
#ifdef NGM_HAVE_S16
		if (c->align_s16[mode]) {
			e = launch_align_s16(c->capacity, mode, a, st);
			c->launches += 1;                                      // forward + backtrace/format
		} else
#endif
		e = launch_align_i32(c->capacity, mode, a, st);
		c->launches += 1;
		if (e != cudaSuccess) return fail(NGM_B200_ECUDA, "align kernel launch: %s", cudaGetErrorString(e));
		if (prof) {
			CU(cudaEventRecord(c->pev[c->pev_used + 2], st));
			c->pev_used += 3;
		}
	}
	return NGM_B200_OK;
}

}  // namespace ngm

namespace {

// Gather + upload + pack one strict-path chunk (copySeqDataToDevice, SWOcl.cpp:534-556).
int stage_strict(ngm_b200_ctx *c, int base, int m, const char *const *ref, const char *const *qry, const char *dir) {
	const int qml = c->dp.qml, rw = c->ref_width;
	CU(c->h_reads.ensure((size_t) m * qml));
	CU(c->h_refs.ensure((size_t) m * rw));
	CU(c->h_flags.ensure((size_t) m));
	char *hr = c->h_reads.as<char>(), *hf = c->h_refs.as<char>();
	uint8_t *fl = c->h_flags.as<uint8_t>();
	const int RW = c->dp.read_words, WW = c->win_words;
	CU(c->d_areads.ensure((size_t) m * qml));
	CU(c->d_arefs.ensure((size_t) m * rw));
	CU(c->d_flags.ensure((size_t) m));
	CU(c->d_reads4.ensure((size_t) m * RW * 4));
	CU(c->d_rlen32.ensure((size_t) m * 4));
	CU(c->d_rlen.ensure((size_t) m * 2));
	CU(c->d_wins4.ensure(((size_t) m + 2) * WW * 4 + 4096));
	CU(c->d_pairs.ensure((size_t) m * sizeof(PairDesc)));
	cudaStream_t st = c->stream;
	// the rows are gathered piece by piece and every piece is sent while the next one is gathered (the reference overlaps the packing of
	// one chunk with the kernel of the previous one the same way, SWOcl.cpp:435-444)
	constexpr int kPiece = 8192;                               // a multiple of 4: quad leaders stay inside their piece
	for (int p0 = 0; p0 < m; p0 += kPiece) {
		const int p1 = std::min(m, p0 + kPiece);
		for (int i = p0; i < p1; ++i) {
			memcpy(hr + (size_t) i * qml, qry[base + i], (size_t) qml);
			memcpy(hf + (size_t) i * rw, ref[base + i], (size_t) rw);
		}
		for (int i = p0; i < p1; ++i) {
			// CPU-device quirk: only lane 0 of every quad is tested for an empty read
			// (oclSwScore.cl:37,124; oclEndFreeScore.cl:20,74); base is a multiple of 4.
			const int leader = i & ~3;
			uint32_t f = hr[(size_t) leader * qml] == '\0' ? PF_INACTIVE : 0u;
			if (dir != nullptr && (c->dp.alt || c->dp.acct_alt) && dir[base + i] != 0) f |= PF_DIR;
			fl[i] = (uint8_t) f;
		}
		const size_t np = (size_t) (p1 - p0);
		CU(cudaMemcpyAsync(c->d_areads.as<char>() + (size_t) p0 * qml, hr + (size_t) p0 * qml, np * qml, cudaMemcpyHostToDevice, st));
		CU(cudaMemcpyAsync(c->d_arefs.as<char>() + (size_t) p0 * rw, hf + (size_t) p0 * rw, np * rw, cudaMemcpyHostToDevice, st));
		CU(cudaMemcpyAsync(c->d_flags.as<char>() + p0, fl + p0, np, cudaMemcpyHostToDevice, st));
	}
	CU(cudaMemsetAsync(c->d_rlen32.p, 0, (size_t) m * 4, st));
	CU(c->d_noncanon.ensure((size_t) m));
	CU(cudaMemsetAsync(c->d_noncanon.p, 0, (size_t) m, st));
	{
		const long long tot = (long long) m * RW;
		pack_ascii_kernel<<<(unsigned) ((tot + 255) / 256), 256, 0, st>>>(c->d_areads.as<uint8_t>(), m, m, qml, qml, c->d_reads4.as<uint32_t>(), RW,
				c->d_rlen32.as<unsigned int>(), c->d_noncanon.as<uint8_t>());
		narrow_rlen_kernel<<<(m + 255) / 256, 256, 0, st>>>(c->d_rlen32.as<unsigned int>(), c->d_rlen.as<uint16_t>(), m);
		const long long totw = ((long long) m + 2) * WW;   // two extra all-NUL windows as overrun padding
		pack_ascii_kernel<<<(unsigned) ((totw + 255) / 256), 256, 0, st>>>(c->d_arefs.as<uint8_t>(), m, m + 2, rw, rw, c->d_wins4.as<uint32_t>(), WW, nullptr, c->d_noncanon.as<uint8_t>());
		strict_pairs_kernel<<<(m + 255) / 256, 256, 0, st>>>(c->d_pairs.as<PairDesc>(), c->d_flags.as<uint8_t>(), m, WW);
		c->launches += 4;
	}
	CU(cudaGetLastError());
	return NGM_B200_OK;
}

}  // namespace

extern "C" {

int ngm_b200_abi_version(void) { return NGM_B200_ABI_VERSION; }

const char *ngm_b200_last_error(void) { return ngm::last_error_text(); }

int ngm_b200_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

ngm_b200_ctx *ngm_b200_create(const ngm_b200_params *params) {
	if (params == nullptr) {
		fail(NGM_B200_EINVAL, "params == NULL");
		return nullptr;
	}
	DevParams dp;
	int use_s16 = 0, align_s16[2] = {0, 0};
	if (build_params(*params, dp, use_s16, align_s16) != NGM_B200_OK) return nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) {
		cudaGetLastError();
		fail(NGM_B200_ECUDA, "no CUDA device available (%s); this backend has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
		return nullptr;
	}
	if (params->device < 0 || params->device >= ndev) {
		fail(NGM_B200_EINVAL, "device %d not in [0, %d)", params->device, ndev);
		return nullptr;
	}
	ngm_b200_ctx *c = new (std::nothrow) ngm_b200_ctx();
	if (c == nullptr) return nullptr;
	c->hp = *params;
	c->dp = dp;
	c->device = params->device;
	c->capacity = band_capacity(dp.corridor);
	c->use_s16 = params->lane_mode == 1 ? 0 : (params->lane_mode == 2 ? 1 : use_s16);
	for (int m = 0; m < 2; ++m) c->align_s16[m] = params->lane_mode == 1 ? 0 : align_s16[m];
	if (c->capacity > kAlignS16MaxKnown) c->align_s16[0] = 0;
	if (c->capacity > kAlignS16MaxEndFree) c->align_s16[1] = 0;
#ifndef NGM_HAVE_S16
	c->use_s16 = 0;
	c->align_s16[0] = c->align_s16[1] = 0;
#endif
	if (params->lane_mode == 2 && !use_s16) {
		fail(NGM_B200_ERANGE, "s16x2 lanes forced but the scoring range does not fit int16");
		delete c;
		return nullptr;
	}
	c->score_batch = params->score_batch > 0 ? params->score_batch : 262144;
	c->align_batch = params->align_batch > 0 ? params->align_batch : 262144;
	c->strict_chunk = 131072;
	c->ref_width = dp.qml + dp.corridor;
	// words per packed strict window: the kernels read up to rows_cap + capacity + 16 nibbles
	c->win_words = (dp.rows_cap + c->capacity + 16 + 7) / 8 + 2;
	const size_t per_aln = (size_t) dp.rows_cap * ptr_words_for(c->capacity) * 4 + (2 * (size_t) dp.qml + dp.corridor + 2) * 2;
	// alignments per launch set: the pointer matrix of a launch set is scratch (2.5 KB per 150 bp alignment); 1 M alignments per set keep the
	// grid at ~7 waves of 4 blocks per SM (a 262144-alignment set is 1.7 waves: the second wave runs 73 % full)
	c->align_chunk = (int) std::max<size_t>(4096, std::min<size_t>(1 << 20, (6ull << 30) / per_aln)) / 128 * 128;
	if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
		fail(NGM_B200_ECUDA, "cannot initialise device %d: %s", c->device, cudaGetErrorString(cudaGetLastError()));
		delete c;
		return nullptr;
	}
	return c;
}

void ngm_b200_destroy(ngm_b200_ctx *c) {
	if (c == nullptr) return;
	cudaSetDevice(c->device);
	for (cudaEvent_t &e : c->pev)
		if (e) cudaEventDestroy(e);
	if (c->batch) batch_release(c->batch);                 // the lanes first: they borrow this context's buffers
	c->batch = nullptr;
	if (c->stream) {
		cudaStreamSynchronize(c->stream);
		cudaStreamDestroy(c->stream);
	}
	HostBuf *hb[] = { &c->h_reads, &c->h_refs, &c->h_flags, &c->h_scores, &c->h_recs, &c->h_strings, &c->h_cursor, &c->h_noncanon };
	for (HostBuf *b : hb) b->release();
	DevBuf *db[] = { &c->d_areads, &c->d_arefs, &c->d_flags, &c->d_reads4, &c->d_rlen32, &c->d_rlen, &c->d_wins4, &c->d_pairs, &c->d_scores,
			&c->d_recs, &c->d_strings, &c->d_cursor, &c->d_ptr, &c->d_ops, &c->d_best, &c->d_known, &c->d_ref4, &c->d_rfwd, &c->d_rrev, &c->d_rrlen32, &c->d_rrlen,
			&c->d_rascii, &c->d_upairs, &c->d_rpairs, &c->d_noncanon };
	for (DevBuf *b : db) b->release();
	if (c->cs) cs_release(c->cs);
	if (c->pe) pe_release(c->pe);
	delete c;
}

int ngm_b200_alu_peak(ngm_b200_ctx *c, double *viaddmnmx_per_s, double *imad_per_s, double *mixed_per_s) {
	if (c == nullptr) return fail(NGM_B200_EINVAL, "ctx == NULL");
	CU(cudaSetDevice(c->device));
	int sms = 0;
	CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
	const int blocks = sms * 8, iters = 4096;
	DevBuf out;
	CU(out.ensure((size_t) blocks * 256 * 4));
	cudaEvent_t e0, e1;
	CU(cudaEventCreate(&e0));
	CU(cudaEventCreate(&e1));
	double res[3] = {0, 0, 0};
	for (int kind = 0; kind < 3; ++kind) {
		float best = 1e30f;
		for (int rep = 0; rep < 4; ++rep) {                    // the first repetition warms up
			CU(cudaEventRecord(e0, c->stream));
			if (kind == 0) alu_peak_kernel<0><<<blocks, 256, 0, c->stream>>>(out.as<uint32_t>(), iters, c->dp.c_neg1 << 2, c->dp.c_four | 1u);
			else if (kind == 1) alu_peak_kernel<1><<<blocks, 256, 0, c->stream>>>(out.as<uint32_t>(), iters, c->dp.c_neg1 << 2, c->dp.c_four | 1u);
			else alu_peak_kernel<2><<<blocks, 256, 0, c->stream>>>(out.as<uint32_t>(), iters, c->dp.c_neg1 << 2, c->dp.c_four | 1u);
			CU(cudaEventRecord(e1, c->stream));
			CU(cudaEventSynchronize(e1));
			float ms = 0;
			CU(cudaEventElapsedTime(&ms, e0, e1));
			if (rep) best = std::min(best, ms);
		}
		c->launches += 4;
		const double instr = (double) blocks * 256.0 * iters * 64.0 * (kind == 2 ? 2.0 : 1.0);      // thread-level instructions of the measured kind(s)
		res[kind] = instr / (best * 1e-3);
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	out.release();
	if (viaddmnmx_per_s) *viaddmnmx_per_s = res[0];
	if (imad_per_s) *imad_per_s = res[1];
	if (mixed_per_s) *mixed_per_s = res[2];
	return NGM_B200_OK;
}

int ngm_b200_profile(ngm_b200_ctx *c, int enable) {
	if (c == nullptr) return fail(NGM_B200_EINVAL, "ctx == NULL");
	c->profile = enable ? 1 : 0;
	c->pev_used = 0;
	return NGM_B200_OK;
}

int ngm_b200_profile_read(ngm_b200_ctx *c, float *forward_ms, float *backtrace_ms) {
	if (c == nullptr) return fail(NGM_B200_EINVAL, "ctx == NULL");
	CU(cudaSetDevice(c->device));
	float f = 0, b = 0;
	for (int i = 0; i + 3 <= c->pev_used; i += 3) {
		CU(cudaEventSynchronize(c->pev[i + 2]));
		float x = 0, y = 0;
		CU(cudaEventElapsedTime(&x, c->pev[i], c->pev[i + 1]));
		CU(cudaEventElapsedTime(&y, c->pev[i + 1], c->pev[i + 2]));
		f += x;
		b += y;
	}
	const int sets = c->pev_used / 3;
	c->pev_used = 0;
	if (forward_ms) *forward_ms = f;
	if (backtrace_ms) *backtrace_ms = b;
	return sets;
}

int ngm_b200_score_batch_size(const ngm_b200_ctx *c) { return c ? c->score_batch : 0; }
int ngm_b200_align_batch_size(const ngm_b200_ctx *c) { return c ? c->align_batch : 0; }
uint64_t ngm_b200_launch_count(const ngm_b200_ctx *c) { return c ? c->launches + batch_lane_launches(c) : 0; }

int ngm_b200_batch_score(ngm_b200_ctx *c, int mode, int n, const char *const *ref, const char *const *qry, float *scores, const char *dir) {
	if (c == nullptr) return fail(NGM_B200_EINVAL, "ctx == NULL");
	if (n <= 0) return 0;                                  // SWOcl.cpp:39-42
	const int m0 = mode_of(mode);
	if (m0 < 0) return fail(NGM_B200_EINVAL, "unsupported alignment mode %d", mode & 0xFF);
	if (ref == nullptr || qry == nullptr || scores == nullptr) return fail(NGM_B200_EINVAL, "NULL sequence list / result buffer");
	CU(cudaSetDevice(c->device));
	for (int base = 0; base < n; base += c->strict_chunk) {
		const int m = std::min(c->strict_chunk, n - base);
		int rc = stage_strict(c, base, m, ref, qry, dir);
		if (rc) return rc;
		CU(c->d_scores.ensure((size_t) m * 4));
		CU(c->h_scores.ensure((size_t) m * 4));
		ScoreArgs a = score_args(c, c->d_pairs.as<PairDesc>(), m, c->d_reads4.as<uint32_t>(), c->d_reads4.as<uint32_t>(), c->d_rlen.as<uint16_t>(),
				c->d_wins4.as<uint32_t>(), c->d_scores.as<float>());
		rc = run_score(c, m0, a, c->stream);
		if (rc) return rc;
		CU(cudaMemcpyAsync(c->h_scores.p, c->d_scores.p, (size_t) m * 4, cudaMemcpyDeviceToHost, c->stream));
		CU(cudaStreamSynchronize(c->stream));
		memcpy(scores + base, c->h_scores.p, (size_t) m * 4);
	}
	return n;
}

int ngm_b200_batch_align(ngm_b200_ctx *c, int mode, int n, const char *const *ref, const char *const *qry, const char *const *qal,
		ngm_b200_align *results, const char *dir) {
	(void) qal;                                            // only logged by the reference (SWOclCigar.cpp:486)
	if (c == nullptr) return fail(NGM_B200_EINVAL, "ctx == NULL");
	if (n <= 0) return 0;                                  // SWOclCigar.cpp:109-112
	const int m0 = mode_of(mode);
	if (m0 < 0) return fail(NGM_B200_EINVAL, "unsupported alignment mode %d", mode & 0xFF);
	if (ref == nullptr || qry == nullptr || results == nullptr) return fail(NGM_B200_EINVAL, "NULL sequence list / result buffer");
	CU(cudaSetDevice(c->device));
	const int chunk = std::min(c->strict_chunk, c->align_chunk);
	for (int base = 0; base < n; base += chunk) {
		const int m = std::min(chunk, n - base);
		int rc = stage_strict(c, base, m, ref, qry, dir);
		if (rc) return rc;
		// worst case per alignment: every base its own MD token -> < 4*qml+corridor bytes each for CIGAR and MD
		size_t str_cap = (size_t) m * 96;
		CU(c->d_recs.ensure((size_t) m * sizeof(ngm_b200_align_rec)));
		CU(c->h_recs.ensure((size_t) m * sizeof(ngm_b200_align_rec)));
		CU(c->d_cursor.ensure(4));
		CU(c->h_cursor.ensure(4));
		for (int attempt = 0; attempt < 2; ++attempt) {
			CU(c->d_strings.ensure(str_cap));
			CU(c->h_strings.ensure(str_cap));
			CU(cudaMemsetAsync(c->d_cursor.p, 0, 4, c->stream));
			rc = run_align(c, m0, c->d_pairs.as<PairDesc>(), m, c->d_reads4.as<uint32_t>(), c->d_reads4.as<uint32_t>(), c->d_rlen.as<uint16_t>(),
					c->d_wins4.as<uint32_t>(), c->d_recs.as<ngm_b200_align_rec>(), c->d_strings.as<char>(), (uint32_t) str_cap,
					c->d_cursor.as<uint32_t>(), c->stream, nullptr, nullptr);
			if (rc) return rc;
			CU(cudaMemcpyAsync(c->h_cursor.p, c->d_cursor.p, 4, cudaMemcpyDeviceToHost, c->stream));
			CU(cudaStreamSynchronize(c->stream));
			const uint32_t used = *c->h_cursor.as<uint32_t>();
			if (used <= str_cap) {
				CU(c->h_noncanon.ensure((size_t) m));
				CU(cudaMemcpyAsync(c->h_noncanon.p, c->d_noncanon.p, (size_t) m, cudaMemcpyDeviceToHost, c->stream));
				CU(cudaMemcpyAsync(c->h_recs.p, c->d_recs.p, (size_t) m * sizeof(ngm_b200_align_rec), cudaMemcpyDeviceToHost, c->stream));
				if (used) CU(cudaMemcpyAsync(c->h_strings.p, c->d_strings.p, used, cudaMemcpyDeviceToHost, c->stream));
				CU(cudaStreamSynchronize(c->stream));
				break;
			}
			if (attempt == 1) return fail(NGM_B200_ECUDA, "string heap overflow after resize");
			str_cap = (size_t) used + 64;                  // rare: long CIGAR/MD strings; repeat with the exact size
		}
		const ngm_b200_align_rec *recs = c->h_recs.as<ngm_b200_align_rec>();
		const char *heap = c->h_strings.as<char>();
		for (int i = 0; i < m; ++i) {
			const ngm_b200_align_rec &r = recs[i];
			ngm_b200_align &o = results[base + i];
			o.position_offset = r.position_offset;     // SWOclCigar.cpp:328
			o.qstart = r.qstart;
			o.qend = r.qend;
			o.score = r.score;
			if (r.score >= 0.0f) {
				o.identity = r.identity;
				o.nm = r.nm;
				if (o.cigar != nullptr) {
					memcpy(o.cigar, heap + r.str_off, r.cigar_len);
					o.cigar[r.cigar_len] = '\0';
				}
				if (o.md != nullptr) {
					memcpy(o.md, heap + r.str_off + r.cigar_len, r.md_len);
					o.md[r.md_len] = '\0';
					// The device prints MD letters from 4-bit codes.  The reference prints the caller's
					// raw window bytes (SWOclCigar.cpp:515-517,554-556); where a window holds bytes
					// outside "ACGTNx" (never produced by NGM's own decoder) put the originals back.
					if (c->h_noncanon.as<uint8_t>()[i]) {
						const char *w = ref[base + i] + r.position_offset;
						if (c->dp.acct_alt && o.cigar != nullptr) {
							const bool d1 = dir != nullptr && dir[base + i] == 1;
							char from, to;
							if (c->dp.acct_slam) { from = d1 ? 'G' : 'C'; to = d1 ? 'A' : 'T'; }      // SWOclCigar.cpp:312-320
							else { from = d1 ? 'A' : 'T'; to = d1 ? 'G' : 'C'; }                       // SWOclCigar.cpp:303-311
							int mt = 0, mm = 0, tot = 0;
							recount_alt(o.cigar, o.md, r.md_len, w, qry[base + i], r.qstart, from, to, &mt, &mm, &tot);
							o.nm = mm;
							o.identity = mt * 1.0f / tot;
						}
						int ri = 0;
						for (int k = 0; k < r.md_len; ++k) {
							const char ch = o.md[k];
							if (ch >= '0' && ch <= '9') {
								int v = 0;
								while (k < r.md_len && o.md[k] >= '0' && o.md[k] <= '9') v = v * 10 + (o.md[k++] - '0');
								ri += v;
								--k;
							} else if (ch != '^') {
								o.md[k] = w[ri++];
							}
						}
					}
				}
			}
		}
	}
	return n;
}

static int transcode_reference(ngm_b200_ctx *c, const uint8_t *d_packed, uint64_t concat_len, cudaStream_t st) {
	// layout in code words: [concat_len bases]['x' overhang >= 2 windows]['N' region >= 2 windows]
	const uint64_t over = (uint64_t) c->win_words * 8 * 2 + 64;
	const uint64_t n_region_word = (concat_len + over + 7) / 8;
	const uint64_t total_words = n_region_word + (over + 7) / 8;
	CU(c->d_ref4.ensure((size_t) total_words * 4));
	transcode_ref_kernel<<<(unsigned) ((total_words + 255) / 256), 256, 0, st>>>(d_packed, concat_len, c->d_ref4.as<uint32_t>(), total_words,
			n_region_word);
	c->launches += 1;
	CU(cudaGetLastError());
	c->concat_len = concat_len;
	c->n_region_nib = n_region_word * 8;
	c->have_ref = true;
	c->epoch += 1;
	return NGM_B200_OK;
}

int ngm_b200_dev_set_reference(ngm_b200_ctx *c, const void *d_packed, uint64_t concat_len, void *stream) {
	if (c == nullptr || d_packed == nullptr || concat_len == 0) return fail(NGM_B200_EINVAL, "bad reference");
	CU(cudaSetDevice(c->device));
	return transcode_reference(c, static_cast<const uint8_t *>(d_packed), concat_len, static_cast<cudaStream_t>(stream));
}

static int gather_winners(ngm_b200_ctx *c, int n_reads, const void *d_pairs, const void *d_scores, const void *d_best_pair, void *d_out_pairs,
		void *d_out_scores, void *stream) {
	if (c == nullptr || d_pairs == nullptr || d_best_pair == nullptr || d_out_pairs == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	if ((d_scores == nullptr) != (d_out_scores == nullptr)) return fail(NGM_B200_EINVAL, "d_scores and d_out_scores go together");
	if (n_reads <= 0) return 0;
	CU(cudaSetDevice(c->device));
	gather_winners_kernel<<<(n_reads + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(n_reads, static_cast<const ngm_b200_pair *>(d_pairs),
			static_cast<const int *>(d_best_pair), static_cast<ngm_b200_pair *>(d_out_pairs), static_cast<const float *>(d_scores),
			static_cast<float *>(d_out_scores));
	c->launches += 1;
	CU(cudaGetLastError());
	return n_reads;
}

int ngm_b200_dev_gather_winners(ngm_b200_ctx *c, int n_reads, const void *d_pairs, const void *d_best_pair, void *d_out_pairs, void *stream) {
	return gather_winners(c, n_reads, d_pairs, nullptr, d_best_pair, d_out_pairs, nullptr, stream);
}

int ngm_b200_dev_gather_winners_scored(ngm_b200_ctx *c, int n_reads, const void *d_pairs, const void *d_scores, const void *d_best_pair,
		void *d_out_pairs, void *d_out_scores, void *stream) {
	if (d_scores == nullptr || d_out_scores == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	return gather_winners(c, n_reads, d_pairs, d_scores, d_best_pair, d_out_pairs, d_out_scores, stream);
}

int ngm_b200_set_reference(ngm_b200_ctx *c, const uint8_t *packed, uint64_t concat_len) {
	if (c == nullptr || packed == nullptr || concat_len == 0) return fail(NGM_B200_EINVAL, "bad reference");
	CU(cudaSetDevice(c->device));
	DevBuf raw;
	const size_t raw_bytes = (size_t) ((concat_len + 1) / 2);
	CU(raw.ensure(raw_bytes));
	int rc = NGM_B200_OK;
	cudaError_t e = cudaMemcpyAsync(raw.p, packed, raw_bytes, cudaMemcpyHostToDevice, c->stream);
	if (e == cudaSuccess) rc = transcode_reference(c, raw.as<uint8_t>(), concat_len, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	raw.release();
	if (e != cudaSuccess) return fail(NGM_B200_ECUDA, "set_reference: %s", cudaGetErrorString(e));
	return rc;
}

static int pack_reads_device_impl(ngm_b200_ctx *c, const uint8_t *d_ascii, int n_reads, int stride, cudaStream_t st) {
	const int RW = c->dp.read_words;
	CU(c->d_rfwd.ensure((size_t) n_reads * RW * 4));
	CU(c->d_rrev.ensure((size_t) n_reads * RW * 4));
	CU(c->d_rrlen.ensure((size_t) n_reads * 2));
	const int width = std::min(stride, c->dp.qml);
	const long long tot = (long long) n_reads * RW;
	static const bool three_pass = [] { const char *e = getenv("NGM_B200_PACK3"); return e != nullptr && atoi(e) == 1; }();      // A/B measurements
	if (three_pass) {
		pack_words_kernel<<<(unsigned) ((tot + 255) / 256), 256, 0, st>>>(d_ascii, n_reads, width, stride, c->d_rfwd.as<uint32_t>(), RW);
		read_len_kernel<<<(n_reads + 255) / 256, 256, 0, st>>>(c->d_rfwd.as<uint32_t>(), n_reads, RW, c->d_rrlen.as<uint16_t>());
		revcomp_words_kernel<<<(unsigned) ((tot + 255) / 256), 256, 0, st>>>(c->d_rfwd.as<uint32_t>(), c->d_rrlen.as<uint16_t>(), n_reads, RW,
				c->d_rrev.as<uint32_t>());
		c->launches += 3;
	} else {
		pack_reads_fused_kernel<<<(n_reads + kPackRowsPerBlock - 1) / kPackRowsPerBlock, 16 * kPackRowsPerBlock, (size_t) kPackRowsPerBlock * RW * 4, st>>>(
				d_ascii, n_reads, width, stride, c->d_rfwd.as<uint32_t>(), c->d_rrev.as<uint32_t>(), c->d_rrlen.as<uint16_t>(), RW);
		c->launches += 1;
	}
	CU(cudaGetLastError());
	c->n_reads = n_reads;
	return NGM_B200_OK;
}

int ngm_b200_set_reads(ngm_b200_ctx *c, const char *reads, int n_reads, int stride) {
	if (c == nullptr || reads == nullptr || n_reads <= 0 || stride <= 0) return fail(NGM_B200_EINVAL, "bad read batch");
	CU(cudaSetDevice(c->device));
	CU(c->d_rascii.ensure((size_t) n_reads * stride));
	CU(cudaMemcpyAsync(c->d_rascii.p, reads, (size_t) n_reads * stride, cudaMemcpyHostToDevice, c->stream));
	int rc = pack_reads_device_impl(c, c->d_rascii.as<uint8_t>(), n_reads, stride, c->stream);
	if (rc) return rc;
	c->reads_stride = stride;
	CU(cudaStreamSynchronize(c->stream));
	return NGM_B200_OK;
}

int ngm_b200_dev_set_reads(ngm_b200_ctx *c, const void *d_ascii, int n_reads, int stride, void *stream) {
	if (c == nullptr || d_ascii == nullptr || n_reads <= 0 || stride <= 0) return fail(NGM_B200_EINVAL, "bad read batch");
	CU(cudaSetDevice(c->device));
	return pack_reads_device_impl(c, static_cast<const uint8_t *>(d_ascii), n_reads, stride, static_cast<cudaStream_t>(stream));
}

static int resolve_pairs(ngm_b200_ctx *c, const void *d_pairs_user, int n, cudaStream_t st) {
	if (!c->have_ref || c->n_reads == 0) return fail(NGM_B200_ESTATE, "set_reference and set_reads must precede descriptor calls");
	CU(c->d_rpairs.ensure((size_t) n * sizeof(PairDesc)));
	resolve_pairs_kernel<<<(n + 255) / 256, 256, 0, st>>>(static_cast<const ngm_b200_pair *>(d_pairs_user), c->d_rpairs.as<PairDesc>(), n,
			(unsigned long long) c->concat_len, (unsigned long long) c->n_region_nib, c->d_rrlen.as<uint16_t>(), (unsigned int) c->n_reads);
	c->launches += 1;
	CU(cudaGetLastError());
	return NGM_B200_OK;
}

int ngm_b200_dev_score_pairs(ngm_b200_ctx *c, int mode, int n, const void *d_pairs, void *d_scores, void *stream) {
	if (c == nullptr || d_pairs == nullptr || d_scores == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	if (n <= 0) return 0;
	const int m0 = mode_of(mode);
	if (m0 < 0) return fail(NGM_B200_EINVAL, "unsupported alignment mode %d", mode & 0xFF);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	int rc = resolve_pairs(c, d_pairs, n, st);
	if (rc) return rc;
	ScoreArgs a = score_args(c, c->d_rpairs.as<PairDesc>(), n, c->d_rfwd.as<uint32_t>(), c->d_rrev.as<uint32_t>(), c->d_rrlen.as<uint16_t>(),
			c->d_ref4.as<uint32_t>(), static_cast<float *>(d_scores));
	rc = run_score(c, m0, a, st);
	return rc ? rc : n;
}

static int dev_align_pairs(ngm_b200_ctx *c, int mode, int n, const void *d_pairs, const void *d_pair_scores, void *d_recs, void *d_strings,
		uint32_t str_capacity, void *d_str_cursor, void *stream) {
	if (c == nullptr || d_pairs == nullptr || d_recs == nullptr || d_strings == nullptr || d_str_cursor == nullptr)
		return fail(NGM_B200_EINVAL, "NULL argument");
	if (n <= 0) return 0;
	const int m0 = mode_of(mode);
	if (m0 < 0) return fail(NGM_B200_EINVAL, "unsupported alignment mode %d", mode & 0xFF);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	int rc = resolve_pairs(c, d_pairs, n, st);
	if (rc) return rc;
	rc = run_align(c, m0, c->d_rpairs.as<PairDesc>(), n, c->d_rfwd.as<uint32_t>(), c->d_rrev.as<uint32_t>(), c->d_rrlen.as<uint16_t>(),
			c->d_ref4.as<uint32_t>(), static_cast<ngm_b200_align_rec *>(d_recs), static_cast<char *>(d_strings), str_capacity,
			static_cast<uint32_t *>(d_str_cursor), st, static_cast<const float *>(d_pair_scores), nullptr);
	return rc ? rc : n;
}

int ngm_b200_dev_align_pairs(ngm_b200_ctx *c, int mode, int n, const void *d_pairs, void *d_recs, void *d_strings, uint32_t str_capacity,
		void *d_str_cursor, void *stream) {
	return dev_align_pairs(c, mode, n, d_pairs, nullptr, d_recs, d_strings, str_capacity, d_str_cursor, stream);
}

int ngm_b200_dev_align_pairs_scored(ngm_b200_ctx *c, int mode, int n, const void *d_pairs, const void *d_pair_scores, void *d_recs,
		void *d_strings, uint32_t str_capacity, void *d_str_cursor, void *stream) {
	if (d_pair_scores == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	return dev_align_pairs(c, mode, n, d_pairs, d_pair_scores, d_recs, d_strings, str_capacity, d_str_cursor, stream);
}

static int select_top1(ngm_b200_ctx *c, int n_reads, const void *d_cand_begin, const void *d_scores, void *d_best_pair, void *d_mapq,
		void *d_num_top, void *stream) {
	if (c == nullptr || d_cand_begin == nullptr || d_scores == nullptr || d_best_pair == nullptr || d_mapq == nullptr)
		return fail(NGM_B200_EINVAL, "NULL argument");
	if (n_reads <= 0) return 0;
	CU(cudaSetDevice(c->device));
	select_top1_kernel<<<(n_reads + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(n_reads, static_cast<const int *>(d_cand_begin),
			static_cast<const float *>(d_scores), static_cast<int *>(d_best_pair), static_cast<int *>(d_mapq), static_cast<int *>(d_num_top), c->se_strata);
	c->launches += 1;
	CU(cudaGetLastError());
	return n_reads;
}

int ngm_b200_dev_select_top1(ngm_b200_ctx *c, int n_reads, const void *d_cand_begin, const void *d_scores, void *d_best_pair, void *d_mapq,
		void *stream) {
	return select_top1(c, n_reads, d_cand_begin, d_scores, d_best_pair, d_mapq, nullptr, stream);
}

int ngm_b200_dev_select_top1_ex(ngm_b200_ctx *c, int n_reads, const void *d_cand_begin, const void *d_scores, void *d_best_pair, void *d_mapq,
		void *d_num_top, void *stream) {
	if (d_num_top == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	return select_top1(c, n_reads, d_cand_begin, d_scores, d_best_pair, d_mapq, d_num_top, stream);
}

int ngm_b200_score_pairs(ngm_b200_ctx *c, int mode, int n, const ngm_b200_pair *pairs, float *scores) {
	if (c == nullptr || pairs == nullptr || scores == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	if (n <= 0) return 0;
	CU(cudaSetDevice(c->device));
	CU(c->d_upairs.ensure((size_t) n * sizeof(ngm_b200_pair)));
	CU(c->d_scores.ensure((size_t) n * 4));
	CU(cudaMemcpyAsync(c->d_upairs.p, pairs, (size_t) n * sizeof(ngm_b200_pair), cudaMemcpyHostToDevice, c->stream));
	int rc = ngm_b200_dev_score_pairs(c, mode, n, c->d_upairs.p, c->d_scores.p, c->stream);
	if (rc < 0) return rc;
	CU(cudaMemcpyAsync(scores, c->d_scores.p, (size_t) n * 4, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	return n;
}

int ngm_b200_align_pairs(ngm_b200_ctx *c, int mode, int n, const ngm_b200_pair *pairs, ngm_b200_align_rec *recs, char *strings,
		size_t str_capacity, size_t *str_used) {
	if (c == nullptr || pairs == nullptr || recs == nullptr || strings == nullptr) return fail(NGM_B200_EINVAL, "NULL argument");
	if (n <= 0) return 0;
	if (str_capacity > 0xFFFFFFFFull) str_capacity = 0xFFFFFFFFull;
	CU(cudaSetDevice(c->device));
	CU(c->d_upairs.ensure((size_t) n * sizeof(ngm_b200_pair)));
	CU(c->d_recs.ensure((size_t) n * sizeof(ngm_b200_align_rec)));
	CU(c->d_strings.ensure(str_capacity));
	CU(c->d_cursor.ensure(4));
	CU(c->h_cursor.ensure(4));
	CU(cudaMemcpyAsync(c->d_upairs.p, pairs, (size_t) n * sizeof(ngm_b200_pair), cudaMemcpyHostToDevice, c->stream));
	CU(cudaMemsetAsync(c->d_cursor.p, 0, 4, c->stream));
	int rc = ngm_b200_dev_align_pairs(c, mode, n, c->d_upairs.p, c->d_recs.p, c->d_strings.p, (uint32_t) str_capacity, c->d_cursor.p, c->stream);
	if (rc < 0) return rc;
	CU(cudaMemcpyAsync(c->h_cursor.p, c->d_cursor.p, 4, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	const uint32_t used = *c->h_cursor.as<uint32_t>();
	if (str_used) *str_used = used;
	if (used > str_capacity) return fail(NGM_B200_ERANGE, "string heap too small: %u bytes needed", used);
	CU(cudaMemcpyAsync(recs, c->d_recs.p, (size_t) n * sizeof(ngm_b200_align_rec), cudaMemcpyDeviceToHost, c->stream));
	if (used) CU(cudaMemcpyAsync(strings, c->d_strings.p, used, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	return n;
}

}  // extern "C"

namespace ngm {
int pack_reads_device(ngm_b200_ctx *c, const uint8_t *d_ascii, int n_reads, int stride, cudaStream_t st) { return pack_reads_device_impl(c, d_ascii, n_reads, stride, st); }
}  // namespace ngm

namespace ngm {
int resolve_pairs_for(ngm_b200_ctx *c, const void *d_pairs_user, int n, cudaStream_t st) { return resolve_pairs(c, d_pairs_user, n, st); }
}  // namespace ngm
