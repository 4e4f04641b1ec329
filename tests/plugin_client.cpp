// tests/plugin_client.cpp -- a stand-in for NGM's side of the plugin boundary (test infrastructure).
//
// Built twice by tests/test_plugin_abi.py: against include/ngm_plugin_abi.h, and -- where
// /root/reference exists -- against the reference's own include/IAlignment.h, IConfig.h, ILog.h
// (-DUSE_REFERENCE_HEADERS), to show the library is driven correctly by a host compiled with NGM's
// real headers.  dlopens the backend, checks Cookie(), feeds it an IConfig / ILog, and, when a GPU is
// present ("run" mode), creates the aligner the way _NGM::CreateAlignment does and calls BatchScore /
// BatchAlign through the vtable the way ScoreBuffer::DoRun / AlignmentBuffer::DoRun do.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <string>
#include <vector>

#ifdef USE_REFERENCE_HEADERS
#include "IAlignment.h"
#include "IConfig.h"
#include "ILog.h"
#undef Message
#undef Warning
#undef Error
#else
#include "ngm_plugin_abi.h"
#endif

struct Cfg : public IConfig {
	std::map<std::string, float> kv;
	char const *GetString(char const *const) const { return ""; }
	int GetInt(char const *const n) const { return (int) kv.at(n); }
	int GetInt(char const *const n, int, int) const { return (int) kv.at(n); }
	int GetParameter(char const *const n) const { return (int) kv.at(n); }
	float GetFloat(char const *const n) const { return kv.at(n); }
	float GetFloat(char const *const n, float, float) const { return kv.at(n); }
	int GetIntArray(char const *const, int *, int) const { return 0; }
	int GetFloatArray(char const *const, float *, int) const { return 0; }
	int GetDoubleArray(char const *const, double *, int) const { return 0; }
	bool Exists(char const *const n) const { return kv.count(n) != 0; }
	bool HasArray(char const *const) const { return false; }
};

struct Lg : public ILog {
	void _Message(int const lvl, char const *const title, char const *const msg, ...) const {
		va_list ap;
		va_start(ap, msg);
		fprintf(stderr, "[log %d %s] ", lvl, title ? title : "");
		vfprintf(stderr, msg, ap);
		fputc('\n', stderr);
		va_end(ap);
	}
	void _Debug(int const, char const *const, char const *const, ...) const {}
};

int main(int argc, char **argv) {
	if (argc < 3) return 64;
	void *h = dlopen(argv[1], RTLD_NOW);
	if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
	int (*cookie)() = (int (*)()) dlsym(h, "Cookie");
	void (*setLog)(ILog *) = (void (*)(ILog *)) dlsym(h, "SetLog");
	void (*setConfig)(IConfig *) = (void (*)(IConfig *)) dlsym(h, "SetConfig");
	bool (*isAvailable)() = (bool (*)()) dlsym(h, "IsAvailable");
	pfCreateAlignment create = (pfCreateAlignment) dlsym(h, "CreateAlignment");
	pfDeleteAlignment destroy = (pfDeleteAlignment) dlsym(h, "DeleteAlignment");
	void (*delString)(char *) = (void (*)(char *)) dlsym(h, "ExternalDeleteString");
	if (!cookie || !setLog || !setConfig || !isAvailable || !create || !destroy || !delString) { fprintf(stderr, "missing export\n"); return 3; }
	if (cookie() != cCookie) { fprintf(stderr, "cookie mismatch\n"); return 4; }
	static Cfg cfg;
	static Lg lg;
	int qml = 32, corridor = 10;
	cfg.kv["qry_max_len"] = qml; cfg.kv["corridor"] = corridor; cfg.kv["match_bonus"] = 10; cfg.kv["mismatch_penalty"] = 15;
	cfg.kv["gap_read_penalty"] = 20; cfg.kv["gap_ref_penalty"] = 20; cfg.kv["bs_mapping"] = 0; cfg.kv["slam_seq"] = 0;
	setLog(&lg);
	setConfig(&cfg);
	printf("cookie ok available %d\n", isAvailable() ? 1 : 0);
	if (strcmp(argv[2], "probe") == 0) return 0;
	// "run": SURVEY Appendix A rows 0, 2, 6 through the vtable
	IAlignment *al = create(0 | (1 << 8));
	if (!al) { fprintf(stderr, "CreateAlignment failed\n"); return 5; }
	const char *W = "GCCCAGTGTGAATCGCTTAAGGGTTAAGTAAGTGTGATGCAT";
	const char *reads[3] = { "GTGTGAATCGCTTAAGGGTTAAGTAAGTGT", "GTGTGAATCGCTAAGGGTTAAGTAAGTGTG", "GACACTCGCTATGAATCTCTGATTTACCCA" };
	int const n = 3;
	std::vector<std::vector<char> > rb(n, std::vector<char>(((qml + corridor) | 1) + 1, 0)), qb(n, std::vector<char>(qml, 0));
	std::vector<char const *> refs(n), qrys(n);
	for (int i = 0; i < n; ++i) {
		memcpy(rb[i].data(), W, strlen(W));
		memcpy(qb[i].data(), reads[i], strlen(reads[i]));
		refs[i] = rb[i].data();
		qrys[i] = qb[i].data();
	}
	printf("batch sizes %d %d\n", al->GetScoreBatchSize(), al->GetAlignBatchSize());
	for (int mode = 0; mode < 2; ++mode) {
		std::vector<float> sc(n, -7);
		if (al->BatchScore(mode, n, refs.data(), qrys.data(), 0, sc.data(), 0) != n) return 6;
		std::vector<Align> res(n);
		std::vector<std::vector<char> > cg(n, std::vector<char>(4 * qml, 0)), md(n, std::vector<char>(4 * qml, 0));
		for (int i = 0; i < n; ++i) { res[i].pBuffer1 = cg[i].data(); res[i].pBuffer2 = md[i].data(); }
		if (al->BatchAlign(mode | (1 << 8), n, refs.data(), qrys.data(), qrys.data(), res.data(), 0) != n) return 7;
		for (int i = 0; i < n; ++i)
			printf("mode %d pair %d score %g off %d qs %d qe %d nm %d id %g cigar %s md %s as %g\n", mode, i, sc[i], res[i].PositionOffset, res[i].QStart,
					res[i].QEnd, res[i].NM, res[i].Identity, res[i].pBuffer1, res[i].pBuffer2, res[i].Score);
	}
	destroy(al);
	return 0;
}
