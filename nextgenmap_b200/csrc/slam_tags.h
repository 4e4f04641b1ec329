// slam_tags.h -- the per-column record SWOclCigar::computeCigarMD keeps under slam_seq (Align::ExtendedData, an AlignmentPosition per
// aligned column: SWOclCigar.cpp:443-447,484-497,523-535) and the SAM tags GenericReadWriter::computeSlaSeqTags derives from it
// (GenericReadWriter.h:87-181; SAMWriter.cpp:203-222: TC:i, RA:Z, MP:Z).
//
// The device leaves CIGAR and MD; both determine the columns completely: a CIGAR M run is a run of = / X columns, MD names the reference
// base of every X column (X <=> the raw bases differ, oclSwScore.cl:275) and counts the = columns, where the reference base is the read's.
// Host only.
#ifndef NGM_SLAM_TAGS_H
#define NGM_SLAM_TAGS_H

#include <cstddef>
#include <cstdio>
#include <string>
#include <vector>

namespace ngm {

struct SlamPos {                 // AlignmentPosition (include/IAlignment.h:4-12)
	int type;                    // 5 * trans[reference base] + trans[read base]
	int read_pos, ref_pos;       // 0-based; the read position counts the clipped start (read_index starts at QStart, SWOclCigar.cpp:478)
	bool match;
};

inline int slam_trans(unsigned char c) {      // trans[] of SWOclCigar.cpp:412-426: A 0, C 1, G 2, T 3 (either case), anything else 4
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': return 3;
	default: return 4;
	}
}

// cigar / md: what BatchAlign returned; qry: the read as it was aligned (RevSeq for reverse candidates).  false when the two strings do not
// describe one alignment.
inline bool slam_positions(const char *cigar, size_t cigar_len, const char *md, size_t md_len, const char *qry, int qry_len, int qstart,
		std::vector<SlamPos> &out) {
	out.clear();
	size_t mp = 0;
	long eq = 0;                 // = columns left before the next MD letter
	int read_pos = qstart, ref_pos = 0;
	auto md_number = [&]() {
		long v = 0;
		while (mp < md_len && md[mp] >= '0' && md[mp] <= '9') v = v * 10 + (md[mp++] - '0');
		return v;
	};
	size_t cp = 0;
	while (cp < cigar_len && cigar[cp] != '\0') {
		long len = 0;
		const size_t digits_at = cp;
		while (cp < cigar_len && cigar[cp] >= '0' && cigar[cp] <= '9') len = len * 10 + (cigar[cp++] - '0');
		if (cp >= cigar_len || cp == digits_at || len > (1L << 24)) return false;      // an op needs a count
		const char op = cigar[cp++];
		if (op == 'S' || op == 'H') continue;                     // QStart / QEnd: already in read_pos
		if (op == 'I') {
			read_pos += (int) len;
		} else if (op == 'D') {
			if (eq != 0) return false;
			if (mp < md_len && md[mp] >= '0' && md[mp] <= '9' && md_number() != 0) return false;
			if (mp >= md_len || md[mp] != '^') return false;
			mp += 1 + (size_t) len;
			if (mp > md_len) return false;
			ref_pos += (int) len;
		} else if (op == 'M' || op == '=' || op == 'X') {
			for (long i = 0; i < len; ++i) {
				if (read_pos >= qry_len) return false;
				const int q = slam_trans((unsigned char) qry[read_pos]);
				while (eq == 0 && mp < md_len && md[mp] >= '0' && md[mp] <= '9') eq = md_number();
				SlamPos p;
				p.read_pos = read_pos;
				p.ref_pos = ref_pos;
				if (eq > 0) {
					--eq;
					p.type = 5 * q + q;
					p.match = true;
				} else {
					if (mp >= md_len || md[mp] == '^' || md[mp] == '\0') return false;
					p.type = 5 * slam_trans((unsigned char) md[mp++]) + q;
					p.match = false;
				}
				out.push_back(p);
				++read_pos;
				++ref_pos;
			}
		} else {
			return false;
		}
	}
	return true;
}

// "\tTC:i:<n>\tRA:Z:<25 counts>[\tMP:Z:<type:readPos:refPos,...>]" as SAMWriter.cpp:203-222 prints it behind MD:Z
inline void slam_sam_tags(const std::vector<SlamPos> &pos, bool reverse, std::string &out) {
	int rates[25] = {};
	std::string mp;
	char buf[48];
	for (const SlamPos &p : pos) {
		rates[p.type] += 1;
		if (!p.match) {
			snprintf(buf, sizeof buf, "%d:%d:%d,", p.type, p.read_pos + 1, p.ref_pos + 1);
			mp += buf;
		}
	}
	snprintf(buf, sizeof buf, "\tTC:i:%d\tRA:Z:", reverse ? rates[5 * 0 + 2] : rates[5 * 3 + 1]);      // A>G on the reverse strand, T>C on the forward
	out += buf;
	for (int i = 0; i < 25; ++i) {
		snprintf(buf, sizeof buf, i ? ",%d" : "%d", rates[i]);
		out += buf;
	}
	if (!mp.empty()) {
		mp.pop_back();
		out += "\tMP:Z:";
		out += mp;
	}
}

}  // namespace ngm
#endif
