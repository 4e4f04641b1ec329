// link_seam.cpp -- implementation of the two classes NGM's factory names (see OclHost.h).
// Compiled against NGM's own headers (IAlignment.h, IConfig.h, ILog.h) and linked with libngm_b200.so.
#include "SWOclCigar.h"

#include <cstdio>
#include <cstdlib>

#include "IConfig.h"
#include "ILog.h"

// plugin surface of libngm_b200.so (nextgenmap_b200/csrc/plugin.cpp)
extern "C" {
int Cookie();
void SetLog(ILog * log);
void SetConfig(IConfig * config);
bool IsAvailable();
IAlignment * CreateAlignment(int const mode);
void DeleteAlignment(IAlignment * instance);
}

OclHost::OclHost(int const device_type, int gpu_id, int const) {
	// with -g/--gpu NGM hands over the configured device ids (CS.cpp:441-453); without it the id is the
	// CS thread number, which means "CPU sub-device" to the OpenCL backend: all threads share CUDA device 0
	device = (device_type == CL_DEVICE_TYPE_GPU) ? gpu_id : 0;
}

OclHost::~OclHost() {
}

SWOclCigar::SWOclCigar(OclHost * phost) : SWOcl(phost), impl(0) {
	if (Cookie() != cCookie) {
		fprintf(stderr, "libngm_b200: interface cookie mismatch\n");
		exit(1);
	}
	SetLog(const_cast<ILog *>(_log));
	SetConfig(_config);
	if (!IsAvailable()) {
		Log.Error("libngm_b200: no CUDA device available (this backend has no CPU fallback)");
		exit(1);
	}
	impl = CreateAlignment(phost->cudaDevice() | (1 << 8));
	if (impl == 0) {
		Log.Error("libngm_b200: could not create the CUDA alignment backend");
		exit(1);
	}
}

SWOclCigar::~SWOclCigar() {
	DeleteAlignment(impl);
}

int SWOclCigar::GetScoreBatchSize() const {
	return impl->GetScoreBatchSize();
}

int SWOclCigar::GetAlignBatchSize() const {
	return impl->GetAlignBatchSize();
}

int SWOclCigar::BatchScore(int const mode, int const batchSize, char const * const * const refSeqList, char const * const * const qrySeqList,
		char const * const * const qalSeqList, float * const results, void * extData) {
	return impl->BatchScore(mode, batchSize, refSeqList, qrySeqList, qalSeqList, results, extData);
}

int SWOclCigar::BatchAlign(int const mode, int const batchSize, char const * const * const refSeqList, char const * const * const qrySeqList,
		char const * const * const qalSeqList, Align * const results, void * extData) {
	return impl->BatchAlign(mode, batchSize, refSeqList, qrySeqList, qalSeqList, results, extData);
}
