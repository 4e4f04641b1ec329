// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Drives the UNMODIFIED reference backend (SWOclCigar on the CPU OpenCL device)
// through its IAlignment interface (include/IAlignment.h:48-69), the same way
// ScoreBuffer::DoRun (src/ScoreBuffer.cpp:127) and AlignmentBuffer::DoRun
// (src/AlignmentBuffer.cpp:114) do.  Built by oracle/Makefile into
// oracle/_ref/ngm_ref_harness; it supplies only the two globals the library
// expects (_config / _log, cf. lib/mason/opencl/SWOcl_export.cpp:19,28).
//
//   ngm_ref_harness run   <in.bin> <out.bin>
//   ngm_ref_harness bench <in.bin> <threads> <warm_passes> <timed_passes>
//
// in.bin  : int32 hdr[16] = {magic, qml, corridor, mode, n, match, mismatch,
//           gap_read, gap_ref, bs_mapping, slam_seq, match_tt, match_tc,
//           ref_buf_len, n_align, has_dir}; n*ref_buf_len ref bytes;
//           n*qml read bytes; [n dir bytes].  Pairs [0, n_align) are aligned.
// out.bin : n float scores; then per aligned pair
//           {i32 pos, qstart, qend, nm; f32 identity, ascore; u16 clen, mlen;
//            clen cigar bytes; mlen md bytes}.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <string>
#include <vector>
#include <map>
#include <pthread.h>
#include <sys/time.h>

#include "IAlignment.h"
#include "IConfig.h"
#include "ILog.h"
#include "OclHost.h"
#include "SWOclCigar.h"

namespace {

struct MapConfig : public IConfig {
	std::map<std::string, float> kv;
	float get(char const * name) const {
		std::map<std::string, float>::const_iterator it = kv.find(name);
		if (it == kv.end()) {
			fprintf(stderr, "[harness] config key '%s' not set -> 0\n", name);
			return 0.0f;
		}
		return it->second;
	}
	char const * GetString(char const * const) const { return ""; }
	int GetInt(char const * const name) const { return (int) get(name); }
	int GetInt(char const * const name, int, int) const { return (int) get(name); }
	int GetParameter(char const * const name) const { return (int) get(name); }
	float GetFloat(char const * const name) const { return get(name); }
	float GetFloat(char const * const name, float, float) const { return get(name); }
	int GetIntArray(char const * const, int *, int) const { return 0; }
	int GetFloatArray(char const * const, float *, int) const { return 0; }
	int GetDoubleArray(char const * const, double *, int) const { return 0; }
	bool Exists(char const * const name) const { return kv.count(name) != 0; }
	bool HasArray(char const * const) const { return false; }
};

struct StderrLog : public ILog {
	int quiet;
	void _Message(int const lvl, char const * const title, char const * const msg, ...) const {
		if (quiet && lvl < 1) return;
		va_list ap;
		va_start(ap, msg);
		fprintf(stderr, "[ref:%d:%s] ", lvl, title ? title : "");
		vfprintf(stderr, msg, ap);
		fprintf(stderr, "\n");
		va_end(ap);
	}
	void _Debug(int const, char const * const, char const * const, ...) const {}
};

struct Job {
	int hdr[16];
	std::vector<char> ref, qry, dir;
};

bool loadJob(char const * path, Job & j) {
	FILE * f = fopen(path, "rb");
	if (!f) return false;
	if (fread(j.hdr, sizeof(int), 16, f) != 16 || j.hdr[0] != 0x4e474d31) { fclose(f); return false; }
	size_t n = j.hdr[4], qml = j.hdr[1], rbl = j.hdr[13];
	j.ref.resize(n * rbl);
	j.qry.resize(n * qml);
	bool ok = fread(j.ref.data(), 1, n * rbl, f) == n * rbl && fread(j.qry.data(), 1, n * qml, f) == n * qml;
	if (ok && j.hdr[15]) {
		j.dir.resize(n);
		ok = fread(j.dir.data(), 1, n, f) == n;
	}
	fclose(f);
	return ok;
}

void configure(MapConfig & c, Job const & j, int threads) {
	c.kv["qry_max_len"] = j.hdr[1];
	c.kv["corridor"] = j.hdr[2];
	c.kv["match_bonus"] = j.hdr[5];
	c.kv["mismatch_penalty"] = j.hdr[6];
	c.kv["gap_read_penalty"] = j.hdr[7];
	c.kv["gap_ref_penalty"] = j.hdr[8];
	c.kv["bs_mapping"] = j.hdr[9];
	c.kv["slam_seq"] = j.hdr[10];
	c.kv["match_bonus_tt"] = j.hdr[11];
	c.kv["match_bonus_tc"] = j.hdr[12];
	c.kv["block_multiplier"] = 2;
	c.kv["step_count"] = 4;
	c.kv["cpu_threads"] = threads;
	c.kv["hard_clip"] = 0;
	c.kv["silent_clip"] = 0;
}

double now() {
	timeval tv;
	gettimeofday(&tv, 0);
	return tv.tv_sec + tv.tv_usec * 1e-6;
}

pthread_mutex_t ctorLock = PTHREAD_MUTEX_INITIALIZER;

pthread_barrier_t phaseBarrier;

struct Worker {
	Job const * job;
	int tid, threads;
	int warmPasses, timedPasses;
	double scoreSeconds, alignSeconds, wallSeconds;
	long scored, aligned;
	pthread_t th;
};

// One CS-thread equivalent: own OclHost sub-device + own SWOclCigar, batches of
// GetScoreBatchSize() pairs for scoring and GetAlignBatchSize()/2 for alignment
// (src/ScoreBuffer.h:92, src/AlignmentBuffer.h:63).
void * benchThread(void * arg) {
	Worker * w = (Worker *) arg;
	Job const & j = *w->job;
	int const n = j.hdr[4], qml = j.hdr[1], rbl = j.hdr[13], mode = j.hdr[3], nAlign = j.hdr[14];
	pthread_mutex_lock(&ctorLock);
	OclHost * host = new OclHost(CL_DEVICE_TYPE_CPU, w->tid, w->threads);
	SWOclCigar * sw = new SWOclCigar(host);
	pthread_mutex_unlock(&ctorLock);
	int const sb = sw->GetScoreBatchSize(), ab = sw->GetAlignBatchSize() / 2;
	int lo = (long) n * w->tid / w->threads, hi = (long) n * (w->tid + 1) / w->threads;
	int alo = (long) nAlign * w->tid / w->threads, ahi = (long) nAlign * (w->tid + 1) / w->threads;
	std::vector<char const *> refs(sb), qrys(sb), qals(sb);
	std::vector<float> scores(sb);
	std::vector<Align> aligns(ab);
	std::vector<char> strbuf((size_t) ab * 8 * qml);
	for (int i = 0; i < ab; ++i) {
		aligns[i].pBuffer1 = &strbuf[(size_t) i * 8 * qml];
		aligns[i].pBuffer2 = &strbuf[(size_t) i * 8 * qml + 4 * qml];
	}
	w->scoreSeconds = w->alignSeconds = 0;
	w->scored = w->aligned = 0;
	pthread_barrier_wait(&phaseBarrier);          // every instance is constructed (JIT done)
	double t0 = 0;
	for (int pass = 0; pass < w->warmPasses + w->timedPasses; ++pass) {
		bool const timed = pass >= w->warmPasses;
		if (pass == w->warmPasses) {
			pthread_barrier_wait(&phaseBarrier);  // all threads start the timed passes together
			t0 = now();
		}
		double a = now();
		for (int s = lo; s < hi; s += sb) {
			int m = (hi - s < sb) ? hi - s : sb;
			for (int i = 0; i < m; ++i) {
				refs[i] = &j.ref[(size_t) (s + i) * rbl];
				qrys[i] = &j.qry[(size_t) (s + i) * qml];
			}
			sw->BatchScore(mode, m, refs.data(), qrys.data(), 0, scores.data(), j.hdr[15] ? (void *) &j.dir[s] : 0);
			if (timed) w->scored += m;
		}
		double b = now();
		for (int s = alo; s < ahi; s += ab) {
			int m = (ahi - s < ab) ? ahi - s : ab;
			for (int i = 0; i < m; ++i) {
				refs[i] = &j.ref[(size_t) (s + i) * rbl];
				qrys[i] = &j.qry[(size_t) (s + i) * qml];
				qals[i] = qrys[i];
			}
			sw->BatchAlign(mode | (1 << 8), m, refs.data(), qrys.data(), qals.data(), aligns.data(), j.hdr[15] ? (void *) &j.dir[s] : 0);
			if (timed) w->aligned += m;
		}
		double c = now();
		if (timed) {
			w->scoreSeconds += b - a;
			w->alignSeconds += c - b;
		}
	}
	w->wallSeconds = now() - t0;
	pthread_mutex_lock(&ctorLock);
	delete sw;
	delete host;
	pthread_mutex_unlock(&ctorLock);
	return 0;
}

int runOnce(Job const & j, char const * outPath) {
	int const n = j.hdr[4], qml = j.hdr[1], rbl = j.hdr[13], mode = j.hdr[3], nAlign = j.hdr[14];
	OclHost * host = new OclHost(CL_DEVICE_TYPE_CPU, 0, 1);
	SWOclCigar * sw = new SWOclCigar(host);
	int const sb = sw->GetScoreBatchSize(), ab = sw->GetAlignBatchSize() / 2;
	std::vector<char const *> refs(n), qrys(n);
	for (int i = 0; i < n; ++i) {
		refs[i] = &j.ref[(size_t) i * rbl];
		qrys[i] = &j.qry[(size_t) i * qml];
	}
	std::vector<float> scores(n, -12345.0f);
	for (int s = 0; s < n; s += sb) {
		int m = (n - s < sb) ? n - s : sb;
		int got = sw->BatchScore(mode, m, &refs[s], &qrys[s], 0, &scores[s], j.hdr[15] ? (void *) &j.dir[s] : 0);
		if (got != m) fprintf(stderr, "[harness] BatchScore returned %d of %d\n", got, m);
	}
	FILE * f = fopen(outPath, "wb");
	if (!f) return 2;
	fwrite(scores.data(), sizeof(float), n, f);
	std::vector<Align> aligns(ab);
	std::vector<char> strbuf((size_t) ab * 8 * qml);
	for (int s = 0; s < nAlign; s += ab) {
		int m = (nAlign - s < ab) ? nAlign - s : ab;
		for (int i = 0; i < m; ++i) {
			aligns[i] = Align();
			aligns[i].pBuffer1 = &strbuf[(size_t) i * 8 * qml];
			aligns[i].pBuffer2 = &strbuf[(size_t) i * 8 * qml + 4 * qml];
			// AlignmentBuffer.cpp:108-109 pre-fills both with "!!!\0"
			memset(aligns[i].pBuffer1, 0, 4 * qml);
			memset(aligns[i].pBuffer2, 0, 4 * qml);
			memcpy(aligns[i].pBuffer1, "!!!", 4);
			memcpy(aligns[i].pBuffer2, "!!!", 4);
		}
		int got = sw->BatchAlign(mode | (1 << 8), m, &refs[s], &qrys[s], &qrys[s], aligns.data(), j.hdr[15] ? (void *) &j.dir[s] : 0);
		if (got != m) fprintf(stderr, "[harness] BatchAlign returned %d of %d\n", got, m);
		for (int i = 0; i < m; ++i) {
			Align const & a = aligns[i];
			int iv[4] = { a.PositionOffset, a.QStart, a.QEnd, a.NM };
			float fv[2] = { a.Identity, a.Score };
			unsigned short lens[2] = { (unsigned short) strlen(a.pBuffer1), (unsigned short) strlen(a.pBuffer2) };
			fwrite(iv, sizeof(int), 4, f);
			fwrite(fv, sizeof(float), 2, f);
			fwrite(lens, sizeof(unsigned short), 2, f);
			fwrite(a.pBuffer1, 1, lens[0], f);
			fwrite(a.pBuffer2, 1, lens[1], f);
		}
	}
	fclose(f);
	delete sw;
	delete host;
	return 0;
}

}  // namespace

IConfig * _config = 0;
ILog const * _log = 0;

int main(int argc, char ** argv) {
	if (argc < 4) {
		fprintf(stderr, "usage: %s run <in.bin> <out.bin> | bench <in.bin> <threads> <warm_passes> <timed_passes>\n", argv[0]);
		return 64;
	}
	static MapConfig cfg;
	static StderrLog log;
	log.quiet = getenv("NGM_REF_VERBOSE") == 0;
	_config = &cfg;
	_log = &log;
	Job job;
	if (!loadJob(argv[2], job)) {
		fprintf(stderr, "[harness] cannot read %s\n", argv[2]);
		return 65;
	}
	if (strcmp(argv[1], "run") == 0) {
		configure(cfg, job, 1);
		return runOnce(job, argv[3]);
	}
	if (strcmp(argv[1], "bench") == 0 && argc >= 5) {
		int threads = atoi(argv[3]);
		if (threads < 1) threads = 1;
		configure(cfg, job, threads);
		int warm = atoi(argv[4]), timed = argc >= 6 ? atoi(argv[5]) : 1;
		if (timed < 1) timed = 1;
		std::vector<Worker> ws(threads);
		pthread_barrier_init(&phaseBarrier, 0, threads);
		for (int t = 0; t < threads; ++t) {
			ws[t].job = &job;
			ws[t].tid = t;
			ws[t].threads = threads;
			ws[t].warmPasses = warm;
			ws[t].timedPasses = timed;
			pthread_create(&ws[t].th, 0, benchThread, &ws[t]);
		}
		long scored = 0, aligned = 0;
		double sSec = 0, aSec = 0, wall = 0;
		for (int t = 0; t < threads; ++t) {
			pthread_join(ws[t].th, 0);
			scored += ws[t].scored;
			aligned += ws[t].aligned;
			if (ws[t].scoreSeconds > sSec) sSec = ws[t].scoreSeconds;
			if (ws[t].alignSeconds > aSec) aSec = ws[t].alignSeconds;
			if (ws[t].wallSeconds > wall) wall = ws[t].wallSeconds;
		}
		printf("{\"threads\": %d, \"passes\": %d, \"scored\": %ld, \"aligned\": %ld, \"score_seconds\": %.6f, \"align_seconds\": %.6f, \"wall_seconds\": %.6f}\n",
				threads, timed, scored, aligned, sSec, aSec, wall);
		return 0;
	}
	return 64;
}
