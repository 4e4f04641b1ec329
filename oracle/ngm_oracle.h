/* oracle/ngm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's CPU-device alignment path
 * (lib/mason/opencl/opencl/*.cl __CPU__ variants + the host code in
 * lib/mason/opencl/SWOcl.cpp / SWOclCigar.cpp).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this; the product
 * (nextgenmap_b200/) never does.
 *
 * Parity status: PINNED -- checked against outputs of the reference itself
 * (oracle/_ref/ngm_ref_harness, built from the reference's own sources) on the
 * committed fixtures under tests/golden/ and by differential fuzzing
 * (tests/test_oracle_vs_reference.py).
 */
#ifndef NGM_ORACLE_H
#define NGM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ngm_oracle_params {
	float match;          /* match_bonus                      (SWOcl.cpp:210)  */
	float mismatch;       /* -mismatch_penalty                (SWOcl.cpp:211)  */
	float gap_read;       /* -gap_read_penalty                (SWOcl.cpp:212)  */
	float gap_ref;        /* -gap_ref_penalty                 (SWOcl.cpp:213)  */
	float match_alt;      /* match_bonus_tt                   (SWOcl.cpp:230)  */
	float mismatch_alt;   /* +match_bonus_tc (bs) / -match_bonus_tc (slam) (SWOcl.cpp:231,237) */
	int alt_scoring;      /* 0 none, 1 bs_mapping matrices, 2 slamSeq matrices (SWOcl.cpp:228-242) */
	int bs_mapping;       /* host-side X-op accounting        (SWOclCigar.cpp:303-311,500-514) */
	int slam_seq;         /* any bit set => same accounting   (SWOclCigar.cpp:312-320) */
	int hard_clip;        /* SWOclCigar.cpp:450 */
	int silent_clip;      /* SWOclCigar.cpp:454 */
} ngm_oracle_params;

/* SWOcl::BatchScore on the CPU device (SWOcl.cpp:33-162): quad padding with pair 0,
 * zero-filled tail, kernels oclSW (mode 0) / oclSW_Global (mode 1).
 * refs: n rows of ref_stride bytes (>= qml+corridor readable), qrys: n rows of
 * qry_stride bytes (>= qml).  dir may be NULL.  Returns n. */
int ngm_oracle_batch_score(const ngm_oracle_params *p, int mode, int qml, int corridor, int n,
		const char *refs, long ref_stride, const char *qrys, long qry_stride,
		const char *dir, float *scores);

/* SWOclCigar::BatchAlign on the CPU device (SWOclCigar.cpp:104-370): forward kernel
 * oclSW_Score / oclSW_ScoreGlobal, oclSW_Backtracking, computeCigarMD.
 * cigar / md: n rows of str_stride bytes (caller pre-fills, reference callers use
 * 4*qml, AlignmentBuffer.cpp:106-109).  md_len receives the number of MD bytes
 * written before the terminating NUL (may exceed strlen when the alignment ran
 * into NUL padding, SURVEY 8a note 9); -1 where the strings were left untouched.
 * Returns n. */
int ngm_oracle_batch_align(const ngm_oracle_params *p, int mode, int qml, int corridor, int n,
		const char *refs, long ref_stride, const char *qrys, long qry_stride,
		const char *dir,
		int *position_offset, int *qstart, int *qend, int *nm,
		float *identity, float *ascore,
		char *cigar, char *md, long str_stride, int *cigar_len, int *md_len);

/* _SequenceProvider::DecodeRefSequence (SequenceProvider.cpp:382-441) over a
 * reference packed 4 bit/base, high nibble first, A0 T1 G2 C3 N4
 * (SequenceProvider.cpp:72-109).  Returns 0 when offset >= concat_len (buffer
 * untouched), else 1. */
int ngm_oracle_decode_window(const unsigned char *packed, unsigned long long concat_len,
		unsigned long long offset, unsigned long long buffer_len, char *buffer);

/* enc4 packing of an ASCII reference (SequenceProvider.cpp:72-85, 2 bases/byte). */
void ngm_oracle_pack_ref(const char *ascii, unsigned long long len, unsigned char *packed);

#ifdef __cplusplus
}
#endif
#endif
