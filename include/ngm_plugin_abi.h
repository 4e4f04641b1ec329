/* ngm_plugin_abi.h -- the C++ plugin surface of NextGenMap that an alignment backend implements,
 * declared layout-compatibly so that libngm_b200.so can be built without NGM's source tree.
 *
 * These are interface declarations only (no behaviour).  The binary contract -- member order,
 * vtable slot order, struct layout -- is that of the reference headers:
 *     struct Align, class IAlignment, pfCreateAlignment / pfDeleteAlignment, cCookie
 *         -> reference include/IAlignment.h:14-28, 30, 48-69, 71-72
 *     class IConfig  -> reference include/IConfig.h:59-80   (12 virtuals, virtual dtor LAST)
 *     class ILog     -> reference include/ILog.h:4-11       (2 variadic virtuals, dtor, one data member)
 * A backend compiled against the reference's own headers and one compiled against this file are
 * interchangeable; tests/plugin_client.cpp is built both ways to prove it.
 */
#ifndef NGM_PLUGIN_ABI_H
#define NGM_PLUGIN_ABI_H

#ifndef __IALIGNMENT_H__   /* the reference's include guards: never declare the types twice */
#define __IALIGNMENT_H__

struct AlignmentPosition {
	AlignmentPosition() : type(-1), readPosition(0), refPosition(0), match(true) {}
	int type, readPosition, refPosition;
	bool match;
};

struct Align {
	Align() : pBuffer1(0), pBuffer2(0), ExtendedData(0), PositionOffset(0), QStart(0), QEnd(0), Score(0.0f), Identity(0.0f), NM(0) {}
	char *pBuffer1;       /* CIGAR, caller-allocated 4*qry_max_len bytes */
	char *pBuffer2;       /* MD,    caller-allocated 4*qry_max_len bytes */
	void *ExtendedData;
	int PositionOffset, QStart, QEnd;
	float Score, Identity;
	int NM;
};

static int const cCookie = 0x10201130;

class IAlignment {
public:
	virtual ~IAlignment() {}
	virtual int GetScoreBatchSize() const = 0;
	virtual int GetAlignBatchSize() const = 0;
	virtual int BatchScore(int const mode, int const batchSize, char const *const *const refSeqList, char const *const *const qrySeqList,
			char const *const *const qalSeqList, float *const results, void *extData) = 0;
	virtual int BatchAlign(int const mode, int const batchSize, char const *const *const refSeqList, char const *const *const qrySeqList,
			char const *const *const qalSeqList, Align *const results, void *extData) = 0;
};

typedef IAlignment *(*pfCreateAlignment)(int const gpu_id);
typedef void (*pfDeleteAlignment)(IAlignment *);
#endif /* __IALIGNMENT_H__ */

#ifndef __ICONFIG_H__
#define __ICONFIG_H__
class IConfig {
public:
	virtual char const *GetString(char const *const name) const = 0;
	virtual int GetInt(char const *const name) const = 0;
	virtual int GetInt(char const *const name, int min, int max) const = 0;
	virtual int GetParameter(char const *const name) const = 0;
	virtual float GetFloat(char const *const name) const = 0;
	virtual float GetFloat(char const *const name, float min, float max) const = 0;
	virtual int GetIntArray(char const *const name, int *pData, int len) const = 0;
	virtual int GetFloatArray(char const *const name, float *pData, int len) const = 0;
	virtual int GetDoubleArray(char const *const name, double *pData, int len) const = 0;
	virtual bool Exists(char const *const name) const = 0;
	virtual bool HasArray(char const *const name) const = 0;
	virtual ~IConfig() {}
};
#endif /* __ICONFIG_H__ */

#ifndef __ILOG_H__
#define __ILOG_H__
class ILog {
public:
	virtual void _Message(int const lvl, char const *const title, char const *const msg, ...) const = 0;
	virtual void _Debug(int const lvl, char const *const title, char const *const msg, ...) const = 0;
	virtual ~ILog() {}
	void *null;
};
#endif /* __ILOG_H__ */

#endif
