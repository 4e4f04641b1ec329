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

#include "../../include/ngm_b200.h"
#include "../../include/ngm_plugin_abi.h"

#define NGM_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

IConfig *g_config = 0;
ILog const *g_log = 0;

void log_msg(int lvl, char const *fmt, char const *arg) {
	if (g_log != 0) g_log->_Message(lvl, "Score (CUDA sm_100a)", fmt, arg);
	else fprintf(stderr, fmt, arg), fputc('\n', stderr);
}

class CudaSW : public IAlignment {
public:
	explicit CudaSW(ngm_b200_ctx *c) : ctx(c) {}
	virtual ~CudaSW() { ngm_b200_destroy(ctx); }
	virtual int GetScoreBatchSize() const { return ngm_b200_score_batch_size(ctx); }
	virtual int GetAlignBatchSize() const { return ngm_b200_align_batch_size(ctx); }
	virtual int BatchScore(int const mode, int const n, char const *const *const ref, char const *const *const qry, char const *const *const,
			float *const results, void *extData) {
		int const got = ngm_b200_batch_score(ctx, mode, n, ref, qry, results, static_cast<char const *>(extData));
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
		int const got = ngm_b200_batch_align(ctx, mode, n, ref, qry, qal, reinterpret_cast<ngm_b200_align *>(results), static_cast<char const *>(extData));
		if (got < 0) {
			log_msg(2, "BatchAlign failed: %s", ngm_b200_last_error());
			return 0;
		}
		return got;
	}

private:
	ngm_b200_ctx *ctx;
};

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
	ngm_b200_ctx *ctx = ngm_b200_create(&p);
	if (ctx == 0) {
		log_msg(2, "CreateAlignment: %s", ngm_b200_last_error());
		return 0;
	}
	return new CudaSW(ctx);
}

NGM_EXPORT void DeleteAlignment(IAlignment *instance) { delete instance; }

NGM_EXPORT void ExternalDeleteString(char *mem) { delete[] mem; }
