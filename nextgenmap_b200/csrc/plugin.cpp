// plugin.cpp -- NextGenMap plugin surface on top of the C ABI.
//
// `CudaSW` is what NGM instantiates in place of `SWOclCigar` (reference NGM.cpp:388-420) and
// drives through the IAlignment vtable from ScoreBuffer / AlignmentBuffer.  The exported
// functions are the reference's historical DLL surface (lib/mason/opencl/SWOcl_export.cpp:20-83):
// SetLog, Cookie, SetConfig, IsAvailable, CreateAlignment, DeleteAlignment, ExternalDeleteString.
// Parameters are read through the host's IConfig exactly like SWOcl does
// (SWOcl.cpp:165-166,208-242; SWOclCigar.cpp:450-454).
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/ngm_b200.h"
#include "../../include/ngm_plugin_abi.h"
#include "slam_tags.h"

#define NGM_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

IConfig *g_config = 0;
ILog const *g_log = 0;

void log_msg(int lvl, char const *fmt, char const *arg) {
	if (g_log != 0) g_log->_Message(lvl, "Score (CUDA sm_100a)", fmt, arg);
	else fprintf(stderr, fmt, arg), fputc('\n', stderr);
}

// -- re-plumbed callers (link_seam/replumb_shim.cpp): a window buffer that starts with this magic holds the window's start in the
// concatenated reference instead of decoded bases
const char kWindowMagic[8] = { 0x01, 'N', 'G', 'M', 'B', '2', 'W', 0x02 };

constexpr int kMaxDevices = 64;
std::mutex g_mu;                                   // guards the roots and the offered reference
ngm_b200_ctx *g_root[kMaxDevices] = {};            // one root context per device: holds the resident reference for all CS threads
const void *g_ref_packed = 0;                      // NGM's own packed reference (SequenceProvider's binRef), offered by the shim
unsigned long long g_ref_len = 0;

inline bool is_descriptor(char const *window) { return window != 0 && memcmp(window, kWindowMagic, 8) == 0; }

class CudaSW : public IAlignment {
public:
	CudaSW(ngm_b200_ctx *c, int dev, int slam) : ctx(c), device(dev), slam_seq(slam), h_rows(0), h_rows_cap(0) {}
	virtual ~CudaSW() {
		ngm_b200_destroy(ctx);
		if (h_rows) ngm_b200_host_free(h_rows);
	}
	virtual int GetScoreBatchSize() const { return ngm_b200_score_batch_size(ctx); }
	virtual int GetAlignBatchSize() const { return ngm_b200_align_batch_size(ctx); }
	virtual int BatchScore(int const mode, int const n, char const *const *const ref, char const *const *const qry, char const *const *const,
			float *const results, void *extData) {
		int got;
		if (n > 0 && is_descriptor(ref[0])) got = score_descriptors(mode, n, ref, qry, results, static_cast<char const *>(extData));
		else got = ngm_b200_batch_score(ctx, mode, n, ref, qry, results, static_cast<char const *>(extData));
		if (got < 0) {
			// the reference has no error codes: Log.Error + exit (SWOcl.cpp:355-359); callers only
			// compare the return value with n (ScoreBuffer.cpp:131-132)
			log_msg(2, "BatchScore failed: %s", ngm_b200_last_error());
			return 0;
		}
		return got;
	}
	virtual int BatchAlign(int const mode, int const n, char const *const *const ref, char const *const *const qry, char const *const *const qal,
			Align *const results, void *extData) {
		static_assert(sizeof(Align) == sizeof(ngm_b200_align), "ngm_b200_align must mirror struct Align");
		int got;
		if (n > 0 && is_descriptor(ref[0])) got = align_descriptors(mode, n, ref, qry, results, static_cast<char const *>(extData));
		else got = ngm_b200_batch_align(ctx, mode, n, ref, qry, qal, reinterpret_cast<ngm_b200_align *>(results), static_cast<char const *>(extData));
		if (got < 0) {
			log_msg(2, "BatchAlign failed: %s", ngm_b200_last_error());
			return 0;
		}
		if (slam_seq != 0) keep_columns(got, qry, results);
		return got;
	}

private:
	// Align::ExtendedData under slam_seq: one AlignmentPosition per aligned column, closed by the default-constructed entry (type -1), which
	// SAMWriter / BAMWriter turn into the TC / RA / MP tags (SWOclCigar.cpp:443-447,484-497,523-535; GenericReadWriter.h:87-181).  The host
	// frees it with delete[] (MappedRead.cpp:97-99).  Rebuilt from the CIGAR / MD strings just written and the read as it was aligned.
	void keep_columns(int n, char const *const *qry, Align *results) {
		std::vector<ngm::SlamPos> pos;
		for (int i = 0; i < n; ++i) {
			Align &o = results[i];
			if (o.Score < 0.0f || o.pBuffer1 == 0 || o.pBuffer2 == 0 || qry[i] == 0) continue;
			if (!ngm::slam_positions(o.pBuffer1, strlen(o.pBuffer1), o.pBuffer2, strlen(o.pBuffer2), qry[i], (int) strlen(qry[i]), o.QStart, pos)) pos.clear();
			AlignmentPosition *ap = new AlignmentPosition[pos.size() + 1];
			for (size_t k = 0; k < pos.size(); ++k) {
				ap[k].type = pos[k].type;
				ap[k].readPosition = pos[k].read_pos;
				ap[k].refPosition = pos[k].ref_pos;
				ap[k].match = pos[k].match;
			}
			o.ExtendedData = ap;
		}
	}

	// The descriptor path behind the IAlignment calls of the re-plumbed ScoreBuffer / AlignmentBuffer: the read rows of the batch (each
	// distinct row once: consecutive candidates of a read share their qry pointer) go to the device as they are -- the caller already
	// chose Seq or RevSeq (ScoreBuffer.cpp:92-110) --, the windows as {start, row, direction} descriptors.
	int stage(int n, char const *const *ref, char const *const *qry, char const *dir) {
		{
			std::lock_guard<std::mutex> lock(g_mu);
			ngm_b200_ctx *root = g_root[device];
			if (root == 0 || g_ref_packed == 0) {
				fprintf(stderr, "libngm_b200: descriptor batch without a reference\n");
				return NGM_B200_ESTATE;
			}
			if (!root_has_ref[device]) {
				int rc = ngm_b200_set_reference(root, static_cast<const uint8_t *>(g_ref_packed), g_ref_len);
				if (rc < 0) return rc;
				root_has_ref[device] = true;
			}
			int rc = ngm_b200_sync_shared(ctx);
			if (rc < 0) return rc;
		}
		const int qml = g_config->GetInt("qry_max_len");
		if ((size_t) n * qml > h_rows_cap) {
			if (h_rows) ngm_b200_host_free(h_rows);
			h_rows_cap = (size_t) n * qml + (size_t) n * qml / 4 + 4096;
			h_rows = static_cast<char *>(ngm_b200_host_alloc(h_rows_cap));
			if (h_rows == 0) {
				h_rows_cap = 0;
				return NGM_B200_ECUDA;
			}
		}
		desc.resize((size_t) n);
		int rows = 0;
		char const *prev = 0;
		for (int i = 0; i < n; ++i) {
			if (!is_descriptor(ref[i])) return NGM_B200_EINVAL;           // a batch is decoded by one caller: all windows or none
			if (qry[i] != prev) {
				memcpy(h_rows + (size_t) rows * qml, qry[i], (size_t) qml);
				prev = qry[i];
				++rows;
			}
			unsigned long long start;
			memcpy(&start, ref[i] + 8, 8);
			desc[i].window_start = start;
			desc[i].read_index = (uint32_t) (rows - 1);
			desc[i].flags = (dir != 0 && dir[i] != 0) ? NGM_B200_PAIR_DIR : 0u;
		}
		return ngm_b200_set_reads(ctx, h_rows, rows, qml);
	}
	int score_descriptors(int mode, int n, char const *const *ref, char const *const *qry, float *results, char const *dir) {
		int rc = stage(n, ref, qry, dir);
		if (rc < 0) return rc;
		rc = ngm_b200_score_pairs(ctx, mode, n, desc.data(), results);
		return rc < 0 ? rc : n;
	}
	int align_descriptors(int mode, int n, char const *const *ref, char const *const *qry, Align *results, char const *dir) {
		int rc = stage(n, ref, qry, dir);
		if (rc < 0) return rc;
		recs.resize((size_t) n);
		size_t cap = heap.size() < (size_t) n * 96 ? (size_t) n * 96 : heap.size(), used = 0;
		for (int attempt = 0; attempt < 2; ++attempt) {
			heap.resize(cap);
			rc = ngm_b200_align_pairs(ctx, mode, n, desc.data(), recs.data(), heap.data(), cap, &used);
			if (rc == NGM_B200_ERANGE && used > cap && attempt == 0) {
				cap = used + 64;
				continue;
			}
			break;
		}
		if (rc < 0) return rc;
		for (int i = 0; i < n; ++i) {                                     // what SWOclCigar::BatchAlign leaves in struct Align (SWOclCigar.cpp:322-328)
			const ngm_b200_align_rec &r = recs[i];
			Align &o = results[i];
			o.PositionOffset = r.position_offset;
			o.QStart = r.qstart;
			o.QEnd = r.qend;
			o.Score = r.score;
			if (r.score >= 0.0f) {
				o.Identity = r.identity;
				o.NM = r.nm;
				if (o.pBuffer1 != 0) {
					memcpy(o.pBuffer1, heap.data() + r.str_off, r.cigar_len);
					o.pBuffer1[r.cigar_len] = '\0';
				}
				if (o.pBuffer2 != 0) {
					memcpy(o.pBuffer2, heap.data() + r.str_off + r.cigar_len, r.md_len);
					o.pBuffer2[r.md_len] = '\0';
				}
			}
		}
		return n;
	}

	ngm_b200_ctx *ctx;
	int device;
	int slam_seq;                       // "slam_seq" != 0: BatchAlign also leaves Align::ExtendedData
	char *h_rows;                       // pinned staging of the batch's read rows
	size_t h_rows_cap;
	std::vector<ngm_b200_pair> desc;
	std::vector<ngm_b200_align_rec> recs;
	std::vector<char> heap;

public:
	static bool root_has_ref[kMaxDevices];
};

bool CudaSW::root_has_ref[kMaxDevices] = {};

float cfg_float(char const *key, float dflt) { return (g_config != 0 && g_config->Exists(key)) ? g_config->GetFloat(key) : dflt; }
int cfg_int(char const *key, int dflt) { return (g_config != 0 && g_config->Exists(key)) ? g_config->GetInt(key) : dflt; }

}  // namespace

NGM_EXPORT void SetLog(ILog *log) { g_log = log; }

NGM_EXPORT int Cookie() { return cCookie; }

NGM_EXPORT void SetConfig(IConfig *config) { g_config = config; }

NGM_EXPORT bool IsAvailable() { return ngm_b200_device_count() > 0; }

// mode: low byte = device ordinal, (mode >> 8) & 0xFF = report type, must be 1 (CIGAR + MD) like
// the only live branch of the reference's factory (NGM.cpp:407-416).
NGM_EXPORT IAlignment *CreateAlignment(int const mode) {
	if (g_config == 0) {
		log_msg(2, "CreateAlignment: %s", "SetConfig was not called");
		return 0;
	}
	int const report = (mode >> 8) & 0xFF;
	if (report != 1) {
		char buf[32];
		snprintf(buf, sizeof(buf), "%d", mode);
		log_msg(2, "Unsupported report type %s", buf);
		return 0;
	}
	ngm_b200_params p;
	memset(&p, 0, sizeof(p));
	p.qry_max_len = g_config->GetInt("qry_max_len");
	p.corridor = g_config->GetInt("corridor");
	p.match_bonus = g_config->GetFloat("match_bonus");
	p.mismatch_penalty = g_config->GetFloat("mismatch_penalty");
	p.gap_read_penalty = g_config->GetFloat("gap_read_penalty");
	p.gap_ref_penalty = g_config->GetFloat("gap_ref_penalty");
	p.match_bonus_tt = cfg_float("match_bonus_tt", 0.0f);
	p.match_bonus_tc = cfg_float("match_bonus_tc", 0.0f);
	p.bs_mapping = cfg_int("bs_mapping", 0);
	p.slam_seq = cfg_int("slam_seq", 0);
	p.hard_clip = cfg_int("hard_clip", 0);
	p.silent_clip = cfg_int("silent_clip", 0);
	p.device = mode & 0xFF;
	if (p.device < 0 || p.device >= kMaxDevices) {
		log_msg(2, "CreateAlignment: %s", "device ordinal out of range");
		return 0;
	}
	// One root context per device owns what every CS thread shares (the resident reference); the instance NGM gets is a context that
	// borrows it (the reference backend shares its OpenCL context between instances the same way, OclHost.cpp:50,133,156-160).
	std::lock_guard<std::mutex> lock(g_mu);
	if (g_root[p.device] == 0) {
		g_root[p.device] = ngm_b200_create(&p);
		if (g_root[p.device] == 0) {
			log_msg(2, "CreateAlignment: %s", ngm_b200_last_error());
			return 0;
		}
	}
	ngm_b200_ctx *ctx = ngm_b200_create_shared(g_root[p.device]);
	if (ctx == 0) {
		log_msg(2, "CreateAlignment: %s", ngm_b200_last_error());
		return 0;
	}
	return new CudaSW(ctx, p.device, p.slam_seq);
}

// -- the two hooks of link_seam/replumb_shim.cpp ---------------------------------------------------------------
NGM_EXPORT const char *ngm_b200_plugin_window_magic(void) { return kWindowMagic; }

NGM_EXPORT void ngm_b200_plugin_offer_reference(const void *packed, unsigned long long concat_len) {
	if (g_ref_packed == packed && g_ref_len == concat_len) return;             // (unsynchronised fast path: the values never change once set)
	std::lock_guard<std::mutex> lock(g_mu);
	if (g_ref_packed == 0) {
		g_ref_packed = packed;
		g_ref_len = concat_len;
	}
}

NGM_EXPORT void DeleteAlignment(IAlignment *instance) { delete instance; }

NGM_EXPORT void ExternalDeleteString(char *mem) { delete[] mem; }
