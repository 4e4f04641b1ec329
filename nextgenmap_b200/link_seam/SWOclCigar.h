// link_seam/SWOclCigar.h -- see OclHost.h.  `SWOclCigar` here is a thin IAlignment that owns the
// CUDA backend instance obtained from libngm_b200.so's plugin exports.
#ifndef NGM_B200_LINK_SEAM_SWOCLCIGAR_H
#define NGM_B200_LINK_SEAM_SWOCLCIGAR_H

#include "IAlignment.h"
#include "OclHost.h"

class SWOcl : public IAlignment {
public:
	explicit SWOcl(OclHost * phost) : host(phost) {}
	OclHost * getHost() { return host; }

protected:
	OclHost * host;
};

class SWOclCigar : public SWOcl {
public:
	explicit SWOclCigar(OclHost * host);
	virtual ~SWOclCigar();
	virtual int GetScoreBatchSize() const;
	virtual int GetAlignBatchSize() const;
	virtual int BatchScore(int const mode, int const batchSize, char const * const * const refSeqList, char const * const * const qrySeqList,
			char const * const * const qalSeqList, float * const results, void * extData);
	virtual int BatchAlign(int const mode, int const batchSize, char const * const * const refSeqList, char const * const * const qrySeqList,
			char const * const * const qalSeqList, Align * const results, void * extData);

private:
	IAlignment * impl;
};

#endif
