/* oracle/ngm_oracle.c -- TEST INFRASTRUCTURE ONLY (see ngm_oracle.h).
 *
 * A scalar restatement of what the reference computes on its CPU OpenCL device.
 * The reference evaluates four alignments per work-item in float4 lanes; the
 * lanes never interact except through the quad-granular "empty read" test
 * (oclSwScore.cl:37,124; oclEndFreeScore.cl:20,74), which is modelled by the
 * `active` flag below.  Arithmetic is kept in float like the reference.
 */
#include "ngm_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { OP_I = 1, OP_D = 2, OP_S = 4, OP_EQ = 7, OP_X = 8, OP_STOP = 10 }; /* SWOclCigar.cpp:35, oclDefines.cl:21 */
#define END_FREE_MIN (-16000.0f) /* oclDefines.cl:28 short_min */
#define CPU_BATCH 2048           /* SWOcl.cpp:588, SWOclCigar.cpp:690 */

/* oclDefines.cl:64-80: A/a 0, C/c 1, G/g 2, T/t 3, N/n 5, NUL 6, everything else 4 */
static int code_of(char c) {
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': return 3;
	case 'N': case 'n': return 5;
	case '\0': return 6;
	default: return 4;
	}
}

typedef struct { float fwd[7][7], rev[7][7]; } score_tables;

/* oclDefines.cl:85-128, rows = read code, columns = ref code */
static void build_tables(const ngm_oracle_params *p, score_tables *t) {
	const float M = p->match, X = p->mismatch;
	for (int r = 0; r < 7; ++r)
		for (int c = 0; c < 7; ++c) {
			float v;
			if (r == 6) v = 0.0f;                       /* read NUL: zero row */
			else if (r == 5) v = (c < 4) ? 0.0f : X;    /* read N: 0 vs ACGT, mismatch vs X/N/NUL */
			else if (c == 6) v = 0.0f;                  /* ref NUL column */
			else v = (r == c && r < 4) ? M : X;
			t->fwd[r][c] = t->rev[r][c] = v;
		}
	if (p->alt_scoring == 1) {           /* scoresBsFWD / scoresBsREV */
		t->fwd[3][1] = p->mismatch_alt;  /* read T vs ref C */
		t->fwd[3][3] = p->match_alt;     /* read T vs ref T */
		t->rev[0][0] = p->match_alt;     /* read A vs ref A */
		t->rev[0][2] = p->mismatch_alt;  /* read A vs ref G */
		for (int c = 0; c < 7; ++c) t->rev[5][c] = 0.0f; /* oclDefines.cl:108: read-N row all zero in REV */
	} else if (p->alt_scoring == 2) {    /* scoresSlamSeqFWD / scoresSlamSeqREV */
		t->fwd[1][3] = p->mismatch_alt;  /* read C vs ref T */
		t->fwd[3][3] = p->match_alt;
		t->rev[0][0] = p->match_alt;
		t->rev[2][0] = p->mismatch_alt;  /* read G vs ref A */
		for (int c = 0; c < 7; ++c) t->rev[5][c] = 0.0f; /* oclDefines.cl:127 */
	}
}

static inline float fmax2(float a, float b) { return a > b ? a : b; }

/* oclSwScore.cl:111-154 (oclSW), one lane */
static float score_local(const score_tables *t, int rev, int qml, int corridor, const char *ref, const char *read, float gap_read,
		float gap_ref, float *line) {
	const float (*S)[7] = rev ? t->rev : t->fwd;
	float best = -1.0f;
	for (int i = 0; i <= corridor; ++i) line[i] = 0.0f;
	for (int r = 0; r < qml; ++r) {
		int rc = code_of(read[r]);
		float left = 0.0f;
		for (int j = 0; j < corridor; ++j) {
			left = fmax2(0.0f, left + gap_ref);
			left = fmax2(line[j + 1] + gap_read, left);
			left = fmax2(line[j] + S[rc][code_of(ref[r + j])], left);
			best = fmax2(best, left);
			line[j] = left;
		}
	}
	return best;
}

/* oclEndFreeScore.cl:5-55 (oclSW_Global), one lane */
static float score_endfree(const score_tables *t, int rev, int qml, int corridor, const char *ref, const char *read,
		float gap_read, float gap_ref, float *line) {
	const float (*S)[7] = rev ? t->rev : t->fwd;
	for (int i = 0; i <= corridor; ++i) line[i] = 0.0f;
	line[corridor] = END_FREE_MIN;
	for (int r = 0; r < qml; ++r) {
		int rc = code_of(read[r]);
		float left = END_FREE_MIN;
		for (int j = 0; j < corridor; ++j) {
			left = fmax2(line[j + 1] + gap_read, left + gap_ref);
			left = fmax2(line[j] + S[rc][code_of(ref[r + j])], left);
			line[j] = left;
		}
	}
	float best = END_FREE_MIN;
	for (int i = 0; i <= corridor; ++i) best = fmax2(best, line[i]);
	return best;
}

int ngm_oracle_batch_score(const ngm_oracle_params *p, int mode, int qml, int corridor, int n, const char *refs, long ref_stride,
		const char *qrys, long qry_stride, const char *dir, float *scores) {
	if (n <= 0) return 0;                                   /* SWOcl.cpp:39-42 */
	if ((mode & 0xFF) != 0 && (mode & 0xFF) != 1) return 0;
	score_tables t;
	build_tables(p, &t);
	float *line = (float *) malloc(sizeof(float) * (size_t) (corridor + 1));
	for (int base = 0; base < n; base += CPU_BATCH) {       /* runSwScoreKernel chunking, SWOcl.cpp:421 */
		int m = n - base < CPU_BATCH ? n - base : CPU_BATCH;
		for (int i = 0; i < m; ++i) {
			/* lane 0 of the quad decides whether the quad is evaluated at all; a ragged last
			 * quad is filled with copies of pair 0 (SWOcl.cpp:64-80) which only matters for
			 * lanes we drop, and the quad leader is always a real pair. */
			int leader = base + (i & ~3);
			int active = qrys[(long) leader * qry_stride] != '\0';
			const char *ref = refs + (long) (base + i) * ref_stride;
			const char *read = qrys + (long) (base + i) * qry_stride;
			int rev = (dir != NULL && p->alt_scoring) ? (dir[base + i] != 0) : 0;
			float s;
			if ((mode & 0xFF) == 0)
				s = active ? score_local(&t, rev, qml, corridor, ref, read, p->gap_read, p->gap_ref, line) : -1.0f;
			else
				s = active ? score_endfree(&t, rev, qml, corridor, ref, read, p->gap_read, p->gap_ref, line) : END_FREE_MIN;
			scores[base + i] = s;
		}
	}
	free(line);
	return n;
}

typedef struct { short best_read, best_ref, qend; } fwd_result;

#define MAT(m, c) matrix[(size_t) (m) * (size_t) (corridor + 2) + (size_t) (c)]

/* oclSwScore.cl:4-107 (oclSW_Score), one lane; matrix is (qml+1) x (corridor+2) */
static fwd_result forward_local(const ngm_oracle_params *p, const score_tables *t, int rev, int qml, int corridor, const char *ref,
		const char *read, float *line, unsigned char *matrix) {
	const float (*S)[7] = rev ? t->rev : t->fwd;
	float best_read = 0.0f, best_ref = 0.0f, curr_max = -1.0f, read_index = 0.0f;
	for (int i = 0; i <= corridor; ++i) line[i] = 0.0f;
	for (int c = 0; c <= corridor + 1; ++c) MAT(0, c) = OP_STOP;
	for (int r = 0; r < qml; ++r) {
		int rc = code_of(read[r]);
		int m = r + 1;
		float left = 0.0f;
		MAT(m, 0) = OP_STOP;
		for (int j = 0; j < corridor; ++j) {
			int fc = code_of(ref[r + j]);
			left += p->gap_ref;
			float score = S[rc][fc];
			float diag = line[j] + score;
			int pointerX;
			if (p->alt_scoring) pointerX = (rc == fc) ? OP_EQ : OP_X;   /* oclSwScore.cl:69 */
			else pointerX = (score == p->match) ? OP_EQ : OP_X;          /* oclSwScore.cl:64 */
			float up = line[j + 1] + p->gap_read;
			float mx = 0.0f;
			mx = fmax2(left, mx);
			mx = fmax2(diag, mx);
			mx = fmax2(up, mx);
			int pointer = 0;
			if (mx == left) pointer = OP_D;
			if (mx == up) pointer = OP_I;
			if (mx == diag || mx == line[j] + p->mismatch) pointer = pointerX;
			if (pointer == 0 || mx <= 0.0f) pointer = OP_STOP;
			MAT(m, j + 1) = (unsigned char) pointer;
			if (mx > curr_max) {
				best_read = read_index;
				best_ref = (float) j;
			}
			curr_max = fmax2(curr_max, mx);
			left = mx;
			line[j] = mx;
		}
		MAT(m, corridor + 1) = OP_STOP;
		if (rc != 6) read_index += 1.0f;
	}
	fwd_result fr = { (short) best_read, (short) best_ref, (short) (read_index - best_read - 1.0f) };
	return fr;
}

/* oclEndFreeScore.cl:58-146 (oclSW_ScoreGlobal), one lane */
static fwd_result forward_endfree(const ngm_oracle_params *p, const score_tables *t, int rev, int qml, int corridor, const char *ref,
		const char *read, float *line, unsigned char *matrix) {
	const float (*S)[7] = rev ? t->rev : t->fwd;
	float best_ref = 0.0f, read_index = 0.0f;
	for (int i = 0; i <= corridor; ++i) line[i] = 0.0f;
	for (int c = 0; c <= corridor + 1; ++c) MAT(0, c) = OP_STOP;
	line[corridor] = END_FREE_MIN;
	for (int r = 0; r < qml; ++r) {
		int rc = code_of(read[r]);
		int m = r + 1;
		float left = END_FREE_MIN;
		MAT(m, 0) = OP_X;
		for (int j = 0; j < corridor; ++j) {
			int fc = code_of(ref[r + j]);
			left += p->gap_ref;
			float score = S[rc][fc];
			float diag = line[j] + score;
			int pointerX;
			if (p->alt_scoring) pointerX = (rc == fc) ? OP_EQ : OP_X;
			else pointerX = (score == p->match) ? OP_EQ : OP_X;
			float up = line[j + 1] + p->gap_read;
			float mx = fmax2(left, diag);
			mx = fmax2(up, mx);
			int pointer = 0;
			if (mx == left) pointer = OP_D;
			if (mx == up) pointer = OP_I;
			if (mx == diag || mx == line[j] + p->mismatch) pointer = pointerX;
			MAT(m, j + 1) = (unsigned char) pointer;
			left = mx;
			line[j] = mx;
		}
		MAT(m, corridor + 1) = OP_X;
		if (rc != 6) read_index += 1.0f;
	}
	float curr_max = END_FREE_MIN;
	for (int i = 0; i <= corridor; ++i)
		if (line[i] > curr_max) {
			curr_max = line[i];
			best_ref = (float) i;
		}
	fwd_result fr = { (short) (read_index - 1.0f), (short) best_ref, 0 };
	return fr;
}

/* oclSwCigar.cl:60-124 (oclSW_Backtracking), one lane.  result[4] follows the kernel's slot
 * reuse: in {best_read, best_ref, qend, -}, out {ref_position, qstart, qend, alignment_offset}.
 * Returns 0 when the kernel skips the lane (best_read_index <= 0). */
static int backtrack(int corridor, int alignment_length, const unsigned char *matrix, short *result, short *alignments) {
	short best_read = result[0];
	short best_ref = result[1];
	if (best_read <= 0) return 0;
	int m = best_read + 1;
	short abs_ref = (short) (best_ref + best_read);
	short ai = (short) (alignment_length - 1);
	int pointer, elem = OP_S, len = result[2];
	while ((pointer = MAT(m, best_ref + 1)) != OP_STOP) {
		if (pointer == OP_X || pointer == OP_EQ) {
			m -= 1;
			best_read -= 1;
			abs_ref -= 1;
		} else if (pointer == OP_I) {
			m -= 1;
			best_read -= 1;
			best_ref += 1;
		} else {
			best_ref -= 1;
			abs_ref -= 1;
		}
		if (pointer == elem) {
			len += 1;
		} else {
			alignments[ai--] = (short) (len << 4 | elem);
			elem = pointer;
			len = 1;
		}
	}
	alignments[ai--] = (short) (len << 4 | elem);
	alignments[ai] = (short) ((best_read + 1) << 4 | OP_S);
	result[0] = (short) (abs_ref + 1);
	result[1] = (short) (best_read + 1);
	result[3] = ai;
	return 1;
}

typedef struct {
	int qstart, qend, nm;
	float identity, ascore;
	int cigar_len, md_len;
} cigar_out;

static int put_num(char *dst, int v) { return sprintf(dst, "%d", v); }

/* SWOclCigar::computeCigarMD (SWOclCigar.cpp:430-615).  Returns 0 where the reference returns false. */
static int cigar_md(const ngm_oracle_params *p, int alignment_length, int cig_off, const short *ops, const char *refSeq, const char *qrySeq,
		char bsFrom, char bsTo, char *cigar, char *md, cigar_out *o) {
	int co = 0, mo = 0;
	int alt = p->bs_mapping == 1 || p->slam_seq != 0;
	o->qstart = 0;
	o->qend = 0;
	if ((ops[cig_off] >> 4) > 0) {
		if (p->hard_clip == 1) co += sprintf(cigar + co, "%d%c", ops[cig_off] >> 4, 'H');
		else if (p->silent_clip != 1) co += sprintf(cigar + co, "%d%c", ops[cig_off] >> 4, 'S');
		o->qstart = ops[cig_off] >> 4;
	}
	int match = 0, mismatch = 0, total = 0, m_len = 0, eq_len = 0, ref_index = 0, read_index = o->qstart;
	for (int j = cig_off + 1; j < alignment_length - 1; ++j) {
		int op = ops[j] & 15;
		int length = ops[j] >> 4;
		total += length;
		switch (op) {
		case OP_X:
			m_len += length;
			if (!alt) mismatch += length;
			mo += put_num(md + mo, eq_len);
			for (int k = 0; k < length; ++k) {
				if (alt) {
					/* read_index is advanced in this loop, ref_index too */
					if (qrySeq[read_index] == bsFrom && refSeq[ref_index] == bsTo) match += 1;
					else mismatch += 1;
				}
				md[mo++] = refSeq[ref_index++];
				read_index += 1;
			}
			eq_len = 0;
			break;
		case OP_EQ:
			match += length;
			m_len += length;
			eq_len += length;
			ref_index += length;
			read_index += length;
			break;
		case OP_D:
			if (m_len > 0) {
				co += sprintf(cigar + co, "%d%c", m_len, 'M');
				m_len = 0;
			}
			co += sprintf(cigar + co, "%d%c", length, 'D');
			mo += put_num(md + mo, eq_len);
			eq_len = 0;
			md[mo++] = '^';
			for (int k = 0; k < length; ++k) md[mo++] = refSeq[ref_index++];
			mismatch += length;
			break;
		case OP_I:
			if (m_len > 0) {
				co += sprintf(cigar + co, "%d%c", m_len, 'M');
				m_len = 0;
			}
			co += sprintf(cigar + co, "%d%c", length, 'I');
			read_index += length;
			mismatch += length;
			break;
		default:
			return 0;
		}
	}
	mo += put_num(md + mo, eq_len);
	if (m_len > 0) co += sprintf(cigar + co, "%d%c", m_len, 'M');
	if ((ops[alignment_length - 1] >> 4) > 0) {
		if (p->hard_clip == 1) co += sprintf(cigar + co, "%d%c", ops[alignment_length - 1] >> 4, 'H');
		else if (p->silent_clip != 1) co += sprintf(cigar + co, "%d%c", ops[alignment_length - 1] >> 4, 'S');
		o->qend = ops[alignment_length - 1] >> 4;
	}
	cigar[co] = '\0';
	md[mo] = '\0';
	o->identity = match * 1.0f / total;
	o->nm = mismatch;
	o->ascore = (float) read_index;
	o->cigar_len = co;
	o->md_len = mo;
	return 1;
}

int ngm_oracle_batch_align(const ngm_oracle_params *p, int mode, int qml, int corridor, int n, const char *refs, long ref_stride,
		const char *qrys, long qry_stride, const char *dir, int *position_offset, int *qstart, int *qend, int *nm, float *identity,
		float *ascore, char *cigar, char *md, long str_stride, int *cigar_len, int *md_len) {
	if (n <= 0) return 0;                                   /* SWOclCigar.cpp:109-112 */
	if ((mode & 0xFF) != 0 && (mode & 0xFF) != 1) return 0;  /* reference exit(-1)s, SWOclCigar.cpp:208-210 */
	score_tables t;
	build_tables(p, &t);
	const int alignment_length = 2 * qml + corridor + 1;    /* SWOcl.cpp:342 */
	float *line = (float *) malloc(sizeof(float) * (size_t) (corridor + 1));
	unsigned char *matrix = (unsigned char *) malloc((size_t) (qml + 1) * (size_t) (corridor + 2));
	short *ops = (short *) malloc(sizeof(short) * (size_t) alignment_length * 2);
	for (int i = 0; i < n; ++i) {
		int chunk = i / CPU_BATCH * CPU_BATCH;
		int leader = chunk + ((i - chunk) & ~3);
		int active = qrys[(long) leader * qry_stride] != '\0';
		const char *ref = refs + (long) i * ref_stride;
		const char *read = qrys + (long) i * qry_stride;
		int rev = (dir != NULL && p->alt_scoring) ? (dir[i] != 0) : 0;
		fwd_result fr;
		if (!active) {
			/* quad skipped: oclSwScore.cl:16-18,104-106 / oclEndFreeScore.cl:71-72,143-145 */
			fr.best_read = (short) (((mode & 0xFF) == 0) ? 0 : -1);
			fr.best_ref = 0;
			fr.qend = 0;
		} else if ((mode & 0xFF) == 0) {
			fr = forward_local(p, &t, rev, qml, corridor, ref, read, line, matrix);
		} else {
			fr = forward_endfree(p, &t, rev, qml, corridor, ref, read, line, matrix);
		}
		/* Freshly allocated result / alignment buffers are modelled as zero-filled.  The
		 * reference leaves result[3] and the op array uninitialised when the backtracking
		 * kernel skips a lane (oclSwCigar.cl:78); see DESIGN.md "reference UB". */
		short result[4] = { fr.best_read, fr.best_ref, fr.qend, 0 };
		memset(ops, 0, sizeof(short) * (size_t) alignment_length * 2);
		backtrack(corridor, alignment_length, matrix, result, ops);
		char bsFrom = '0', bsTo = '0';                       /* SWOclCigar.cpp:301-320 */
		if (p->bs_mapping == 1) {
			if (dir != NULL && dir[i] == 1) { bsFrom = 'A'; bsTo = 'G'; } else { bsFrom = 'T'; bsTo = 'C'; }
		}
		if (p->slam_seq) {
			if (dir != NULL && dir[i] == 1) { bsFrom = 'G'; bsTo = 'A'; } else { bsFrom = 'C'; bsTo = 'T'; }
		}
		cigar_out o;
		memset(&o, 0, sizeof(o));
		char *cg = cigar + (long) i * str_stride, *mdp = md + (long) i * str_stride;
		if (cigar_md(p, alignment_length, result[3], ops, ref + result[0], read, bsFrom, bsTo, cg, mdp, &o)) {
			ascore[i] = o.ascore;
			identity[i] = o.identity;
			nm[i] = o.nm;
			cigar_len[i] = o.cigar_len;
			md_len[i] = o.md_len;
		} else {
			ascore[i] = -1.0f;                               /* SWOclCigar.cpp:326 */
			identity[i] = 0.0f;
			nm[i] = 0;
			cigar_len[i] = -1;
			md_len[i] = -1;
		}
		qstart[i] = o.qstart;
		qend[i] = o.qend;
		position_offset[i] = result[0];                      /* SWOclCigar.cpp:328 */
	}
	free(ops);
	free(matrix);
	free(line);
	return n;
}

static char dec4(int v) {
	switch (v) {
	case 0: return 'A';
	case 1: return 'T';
	case 2: return 'G';
	case 3: return 'C';
	default: return 'N';
	}
}

int ngm_oracle_decode_window(const unsigned char *packed, unsigned long long concat_len, unsigned long long offset,
		unsigned long long buffer_len, char *buffer) {
	unsigned long long len = buffer_len - 2;
	if (offset >= concat_len) return 0;
	unsigned long long end = 0;
	if (offset + len > concat_len) {
		end = offset + len - concat_len;
		len -= end;
	}
	unsigned long long start = (offset + 1) / 2;
	unsigned long long k = 0;
	if (offset & 1) buffer[k++] = dec4(packed[start - 1] & 0xF);
	for (unsigned long long i = 0; i < (len + 1) / 2; ++i) {
		buffer[k++] = dec4(packed[start + i] >> 4);
		buffer[k++] = dec4(packed[start + i] & 0xF);
	}
	if (len & 1) buffer[k - 1] = 'x';
	for (unsigned long long i = 0; i < end; ++i) buffer[k++] = 'x';
	for (unsigned long long i = k; i < buffer_len; ++i) buffer[i] = '\0';
	return 1;
}

static int enc4(char c) {
	switch (c) {
	case 'A': case 'a': return 0;
	case 'T': case 't': return 1;
	case 'G': case 'g': return 2;
	case 'C': case 'c': return 3;
	default: return 4;
	}
}

void ngm_oracle_pack_ref(const char *ascii, unsigned long long len, unsigned char *packed) {
	for (unsigned long long i = 0; i + 1 < len; i += 2) packed[i / 2] = (unsigned char) (enc4(ascii[i]) << 4 | enc4(ascii[i + 1]));
	if (len & 1) packed[len / 2] = (unsigned char) (enc4(ascii[len - 1]) << 4 | 4);
}
