// sam_format.cpp -- SAM body lines from the records of the device path ("multi-thread SAM formatting so the host keeps up").
// Host only, no CUDA call.
//
// What NextGenMap does per read after BatchAlign: AlignmentBuffer::DoRun's `Location += PositionOffset - corridor / 2`
// (src/AlignmentBuffer.cpp:129), AlignmentBuffer::WriteRead (convert() to contig coordinates :166-176; for pairs the check of the
// aligned positions :176-200), the output filters of GenericReadWriter::WriteRead / WritePair (src/writer/GenericReadWriter.h:190-312:
// min_identity, min_residues) and SAMWriter::DoWriteReadGeneric / DoWriteUnmappedReadGeneric / DoWritePair
// (src/writer/SAMWriter.cpp:98-228,230-310,312-365).  bs-mapping adds ZS:Z (ngm_b200_sam_opts.bs_mapping).
// Lines come out in read order (mates: second mate's line first, like DoWritePair); NGM's own order depends on its thread timing.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ngm_b200.h"
#include "slam_tags.h"

namespace {

struct ReadView {          // one read after AlignmentBuffer::DoRun: what WriteRead / WritePair look at
	int r = 0;
	const char *name = nullptr;
	const char *seq = nullptr;
	const char *qual = nullptr;
	int length = 0;
	bool has = false;      // hasCandidates() with a usable alignment
	bool reverse = false;
	uint64_t loc = 0;      // Location.m_Location (contig-relative once convert() succeeded)
	uint32_t contig = 0;
	bool converted = false;
	int bp = -1;
	const ngm_b200_align_rec *rec = nullptr;
};

struct Job {
	const ngm_b200_encref *ref;
	const ngm_b200_sam_opts *o;
	const ngm_b200_sam_batch *b;
};

struct CompTable {             // MappedRead.cpp:36-47: A <-> T, C <-> G, everything else unchanged
	unsigned char t[256];
	CompTable() {
		for (int i = 0; i < 256; ++i) t[i] = (unsigned char) i;
		t['A'] = 'T';
		t['T'] = 'A';
		t['C'] = 'G';
		t['G'] = 'C';
	}
};
const CompTable kComp;

void collect_at(const Job &j, int r, int bp, const ngm_b200_align_rec *rec, ReadView &v);

void collect(const Job &j, int r, ReadView &v) { collect_at(j, r, j.b->best_pair[r], &j.b->recs[r], v); }

void collect_at(const Job &j, int r, int bp, const ngm_b200_align_rec *rec, ReadView &v) {
	const ngm_b200_sam_batch &b = *j.b;
	v.r = r;
	v.rec = rec;
	v.name = b.names[r];
	v.seq = b.reads + (size_t) r * b.stride;
	v.qual = b.quals + (size_t) r * b.stride;
	v.length = (int) strnlen(v.seq, (size_t) b.stride);
	if (bp >= 0 && rec->score >= 0.0f && b.mapq[r] >= j.o->min_mq) {          // AlignmentBuffer.cpp:46-49: below min_mq = unmapped, never aligned
		const ngm_b200_pair &p = b.pairs[bp];
		v.has = true;
		v.bp = bp;
		v.reverse = (p.flags & NGM_B200_PAIR_REVERSE) != 0;
		v.loc = p.window_start + (uint64_t) (int64_t) rec->position_offset;      // window_start is Location - corridor / 2 already
		uint32_t contig = 0;
		uint64_t pos = 0;
		if (ngm_b200_convert(j.ref, v.loc, &contig, &pos)) {
			v.contig = contig;
			v.loc = pos;
			v.converted = true;
		}
	}
}

bool passes(const Job &j, const ReadView &v) {     // GenericReadWriter.h:206-214,273-283
	const ngm_b200_align_rec &rec = *v.rec;
	const float mres = j.o->min_residues <= 1.0f ? (float) v.length * j.o->min_residues : j.o->min_residues;
	return rec.identity >= j.o->min_identity && (float) (v.length - rec.qstart - rec.qend) >= mres;
}

struct Line {              // one output line is assembled in a scratch buffer that is large enough by construction (line_bound)
	char *p;
	void push_back(char c) { *p++ = c; }
	void append(const char *s, size_t n) {
		memcpy(p, s, n);
		p += n;
	}
	void append(const char *s) { append(s, strlen(s)); }
};

size_t line_bound_at(const Job &j, int r, bool aligned, const ngm_b200_align_rec &rec) {       // name + FLAG..TLEN + SEQ + QUAL + tags + CIGAR + MD
	return strlen(j.b->names[r]) + 2 * (size_t) j.b->stride + (aligned ? (size_t) rec.cigar_len + rec.md_len : 0) + 2 * 100 + 320 +
			(j.o->read_group != nullptr ? strlen(j.o->read_group) + 8 : 0) + (j.o->slam_seq != 0 ? 20 * (size_t) j.b->stride + 192 : 0);
}

size_t line_bound(const Job &j, int r) { return line_bound_at(j, r, j.b->best_pair[r] >= 0, j.b->recs[r]); }

void put_name(Line &out, const ngm_b200_contig &c) { out.append(c.name, c.name_len); }

inline void put_u64(Line &out, unsigned long long v) {       // what Print("%u") / Print("%d") produce, without the printf machinery
	char buf[24];
	int n = 0;
	do {
		buf[n++] = (char) ('0' + v % 10);
		v /= 10;
	} while (v);
	while (n) out.push_back(buf[--n]);
}

inline void put_int(Line &out, long long v) {
	if (v < 0) {
		out.push_back('-');
		put_u64(out, (unsigned long long) (-v));
	} else {
		put_u64(out, (unsigned long long) v);
	}
}

inline void tag_int(Line &out, const char *tag, long long v) {
	out.append(tag);
	put_int(out, v);
}

// Print("XI:f:%g", round(Identity * 10000.0f) / 10000.0f) (SAMWriter.cpp:188-189): the value is k / 10000 with k = 0 .. 10000, which %g
// prints as 0, 1 or 0.dddd without trailing zeros; anything else (NaN of an empty alignment) goes through printf itself.
inline void put_xi(Line &out, float raw_identity) {
	const float scaled = roundf(raw_identity * 10000.0f);
	if (scaled >= 0.0f && scaled <= 10000.0f) {
		int kk = (int) scaled;
		if (kk == 0 || kk == 10000) {
			out.push_back(kk ? '1' : '0');
			return;
		}
		char d[4] = { (char) ('0' + kk / 1000), (char) ('0' + kk / 100 % 10), (char) ('0' + kk / 10 % 10), (char) ('0' + kk % 10) };
		int n = 4;
		while (d[n - 1] == '0') --n;
		out.push_back('0');
		out.push_back('.');
		out.append(d, (size_t) n);
		return;
	}
	char num[48];
	snprintf(num, sizeof num, "%g", (float) ((double) scaled / 10000.0));
	out.append(num);
}

void mapped_line(const Job &j, const ReadView &v, int flags, const char *rnext, const ngm_b200_contig *rnext_contig, int64_t pnext, int tlen,
		Line &out) {                                        // SAMWriter::DoWriteReadGeneric
	const ngm_b200_sam_batch &b = *j.b;
	const ngm_b200_align_rec &rec = *v.rec;
	if (v.reverse) flags |= 0x10;
	out.append(v.name);
	out.push_back('\t');
	put_int(out, flags);
	out.push_back('\t');
	put_name(out, j.ref->contigs[v.contig]);
	out.push_back('\t');
	put_u64(out, (unsigned) (v.loc + 1));
	out.push_back('\t');
	put_int(out, b.mapq[v.r]);
	out.push_back('\t');
	out.append(b.strings + rec.str_off, rec.cigar_len);
	out.push_back('\t');
	if (rnext_contig) put_name(out, *rnext_contig);
	else out.append(rnext);
	out.push_back('\t');
	put_u64(out, (unsigned) (pnext + 1));
	out.push_back('\t');
	put_int(out, tlen);
	out.push_back('\t');
	// "hard_clip" / "silent_clip": SEQ and QUAL lose the clipped ends of the oriented read (SAMWriter.cpp:104,146-160)
	const int skip = j.o->clip_seq ? rec.qstart : 0;
	const int keep = std::max(0, j.o->clip_seq ? v.length - rec.qstart - rec.qend : v.length);
	char *dst = out.p;
	out.p += 2 * (size_t) keep + 1;
	if (v.reverse) {                                            // RevSeq, reversed qualities (SAMWriter.cpp:120-126)
		const unsigned char *sq = reinterpret_cast<const unsigned char *>(v.seq) + v.length - skip;
		for (int i = 0; i < keep; ++i) dst[i] = (char) kComp.t[*--sq];
		dst[keep] = '\t';
		// (a quality string that starts with '*' counts as "no qualities" and stays as it is, SAMWriter.cpp:122)
		if (v.qual[0] == '*') memcpy(dst + keep + 1, v.qual + skip, (size_t) keep);
		else std::reverse_copy(v.qual + v.length - skip - keep, v.qual + v.length - skip, dst + keep + 1);
	} else {
		memcpy(dst, v.seq + skip, (size_t) keep);
		dst[keep] = '\t';
		memcpy(dst + keep + 1, v.qual + skip, (size_t) keep);
	}
	if (j.o->read_group != nullptr) {                            // SAMWriter.cpp:166-168
		out.append("\tRG:Z:");
		out.append(j.o->read_group);
	}
	const int ntop = b.num_top[v.r];
	tag_int(out, "\tAS:i:", (int) b.scores[v.bp]);
	tag_int(out, "\tNM:i:", rec.nm);
	tag_int(out, "\tNH:i:", ntop);
	if (j.o->bs_mapping == 1) {                                  // SAMWriter.cpp:173-187: which converted strand the read matches
		const bool second = b.pair_fail != nullptr && (v.r & 1); // !(ReadId & 1) || Paired == 0
		out.append(second ? (v.reverse ? "\tZS:Z:+-" : "\tZS:Z:--") : (v.reverse ? "\tZS:Z:-+" : "\tZS:Z:++"));
	}
	out.append("\tXI:f:");
	put_xi(out, rec.identity);
	tag_int(out, "\tX0:i:", ntop);
	tag_int(out, "\tXE:i:", (int) b.max_hit[v.r]);
	tag_int(out, "\tXR:i:", v.length - rec.qstart - rec.qend);
	out.append("\tMD:Z:");
	const char *md = b.strings + rec.str_off + rec.cigar_len;
	const size_t md_len = strnlen(md, rec.md_len);
	out.append(md, md_len);                                     // printed with %s: stops at an embedded NUL (SURVEY 8a note 9)
	if (j.o->slam_seq != 0) {                                   // SAMWriter.cpp:203-222: TC:i / RA:Z / MP:Z from Align::ExtendedData
		std::string oriented(v.seq, (size_t) v.length), tags;
		if (v.reverse) {
			for (int i = 0; i < v.length; ++i) oriented[(size_t) i] = (char) kComp.t[(unsigned char) v.seq[v.length - 1 - i]];
		}
		std::vector<ngm::SlamPos> pos;
		if (ngm::slam_positions(b.strings + rec.str_off, rec.cigar_len, md, md_len, oriented.data(), v.length, rec.qstart, pos)) {
			ngm::slam_sam_tags(pos, v.reverse, tags);
			out.append(tags.data(), tags.size());
		}
	}
	out.push_back('\n');
}

void unmapped_line(const Job &j, const ReadView &v, int flags, const ngm_b200_contig *rname, int64_t loc, char rnext, int64_t pnext, Line &out) {
	out.append(v.name);                                         // SAMWriter::DoWriteUnmappedReadGeneric
	out.push_back('\t');
	put_int(out, flags | 0x4);
	out.push_back('\t');
	if (rname) put_name(out, *rname);
	else out.push_back('*');
	out.push_back('\t');
	put_int(out, (int) (loc + 1));
	out.append("\t0\t*\t");
	out.push_back(rnext);
	out.push_back('\t');
	put_int(out, (int) (pnext + 1));
	out.append("\t0\t");
	out.append(v.seq, (size_t) v.length);
	out.push_back('\t');
	out.append(v.qual, (size_t) v.length);
	if (j.o->read_group != nullptr) {                            // SAMWriter.cpp:358-360
		out.append("\tRG:Z:");
		out.append(j.o->read_group);
	}
	out.push_back('\n');
}

void single(const Job &j, int r, std::string &text, std::vector<char> &scratch) {
	ReadView v;
	collect(j, r, v);
	scratch.resize(std::max(scratch.size(), line_bound(j, r)));
	Line out = { scratch.data() };
	if (v.has && v.converted && passes(j, v)) mapped_line(j, v, 0, "*", nullptr, -1, 0, out);
	else unmapped_line(j, v, 0, nullptr, -1, '*', -1, out);
	text.append(scratch.data(), (size_t) (out.p - scratch.data()));
}

// topn > 1: up to topn lines per read.  AlignmentBuffer::WriteRead converts every location and keeps the LAST result
// (AlignmentBuffer.cpp:169-175); GenericReadWriter::WriteRead (GenericReadWriter.h:200-248) stops accepting alignments after the first one
// that fails a filter and drops repeated locations; lines after the best candidate's carry 0x100 (SAMWriter.cpp:116-118).
void single_topn(const Job &j, int r, std::string &text, std::vector<char> &scratch) {
	const ngm_b200_sam_batch &b = *j.b;
	const int topn = b.topn, ns = b.n_sel[r];
	std::vector<ReadView> views((size_t) std::max(ns, 1));
	size_t bound = line_bound_at(j, r, false, b.recs[(size_t) r * topn]);
	bool mapped = ns > 0;
	for (int k = 0; k < ns; ++k) {
		const ngm_b200_align_rec *rec = &b.recs[(size_t) r * topn + k];
		collect_at(j, r, b.sel[(size_t) r * topn + k], rec, views[(size_t) k]);
		bound += line_bound_at(j, r, true, *rec);
		mapped = views[(size_t) k].converted;                  // the last convert() decides
	}
	for (int k = 0; k < ns; ++k) if (!views[(size_t) k].has) mapped = false;       // a failed alignment: nothing of this read is written as mapped
	scratch.resize(std::max(scratch.size(), bound));
	Line out = { scratch.data() };
	int written = 0;
	for (int k = 0; k < ns; ++k) {
		const ReadView &v = views[(size_t) k];
		mapped = mapped && passes(j, v);
		if (!mapped) continue;
		bool dup = false;
		for (int q = 0; q < k && !dup; ++q) {
			const ReadView &w = views[(size_t) q];
			dup = w.bp == -2 && w.loc == v.loc && w.contig == v.contig && w.reverse == v.reverse;      // bp == -2 marks the locations already written
		}
		if (dup) continue;
		mapped_line(j, v, k ? 0x100 : 0, "*", nullptr, -1, 0, out);
		views[(size_t) k].bp = -2;
		++written;
	}
	if (written == 0) {
		if (ns == 0) collect_at(j, r, -1, &b.recs[(size_t) r * topn], views[0]);
		unmapped_line(j, views[0], 0, nullptr, -1, '*', -1, out);
	}
	text.append(scratch.data(), (size_t) (out.p - scratch.data()));
}

void fragment(const Job &j, int f, std::string &text, std::vector<char> &scratch) {
	const ngm_b200_sam_batch &bt = *j.b;
	ReadView a, b;                                              // a: first mate (ReadId even), b: second mate
	collect(j, 2 * f, a);
	collect(j, 2 * f + 1, b);
	scratch.resize(std::max(scratch.size(), line_bound(j, 2 * f) + line_bound(j, 2 * f + 1)));
	Line out = { scratch.data() };
	const int max_insert = j.o->max_insert_size > 0 ? j.o->max_insert_size : 0x7FFFFFFF;
	bool fail = bt.pair_fail[a.r] != 0 || bt.pair_fail[b.r] != 0;
	if (a.has && b.has) {                                       // AlignmentBuffer::WriteRead: read = first mate (it arrives second)
		const int d = (int) (b.loc > a.loc ? b.loc - a.loc + (uint64_t) a.length : a.loc - b.loc + (uint64_t) b.length);
		if (a.contig != b.contig || d < j.o->min_insert_size || d > max_insert || a.reverse == b.reverse) fail = true;
	}
	if (a.has && !passes(j, a)) a.has = false;                  // WritePair: mapped1 / mapped2, clearScores() otherwise
	if (b.has && !passes(j, b)) b.has = false;
	int fa = 0x1 | 0x40, fb = 0x1 | 0x80;
	const ngm_b200_contig *ca = &j.ref->contigs[a.contig], *cb = &j.ref->contigs[b.contig];
	if (!a.has && !b.has) {
		unmapped_line(j, b, fb | 0x8, nullptr, -1, '*', -1, out);
		unmapped_line(j, a, fa | 0x8, nullptr, -1, '*', -1, out);
	} else if (!a.has) {
		mapped_line(j, b, fb | 0x8, "=", nullptr, (int64_t) b.loc, 0, out);
		unmapped_line(j, a, fa, cb, (int64_t) b.loc, '=', (int64_t) b.loc, out);
	} else if (!b.has) {
		unmapped_line(j, b, fb, ca, (int64_t) a.loc, '=', (int64_t) a.loc, out);
		mapped_line(j, a, fa | 0x8, "=", nullptr, (int64_t) a.loc, 0, out);
	} else if (!fail) {
		fa |= 0x2;
		fb |= 0x2;
		const ReadView &fwd = a.reverse ? b : a, &rev = a.reverse ? a : b;      // exactly one mate is on the minus strand here
		const ngm_b200_align_rec &rec = bt.recs[rev.r];
		const int dist = (int) ((rev.loc + (uint64_t) rev.length - (uint64_t) rec.qstart - (uint64_t) rec.qend) - fwd.loc);
		const int ffwd = (&fwd == &a) ? fa : fb, frev = (&fwd == &a) ? fb : fa;
		mapped_line(j, rev, frev, "=", nullptr, (int64_t) fwd.loc, -dist, out);
		mapped_line(j, fwd, ffwd | 0x20, "=", nullptr, (int64_t) rev.loc, dist, out);
	} else {
		if (a.reverse) fb |= 0x20;
		if (b.reverse) fa |= 0x20;
		mapped_line(j, b, fb, nullptr, ca, (int64_t) a.loc, 0, out);
		mapped_line(j, a, fa, nullptr, cb, (int64_t) b.loc, 0, out);
	}
	text.append(scratch.data(), (size_t) (out.p - scratch.data()));
}

}  // namespace

namespace {
std::mutex g_part_cache_mu;
std::vector<std::string> g_part_cache;
constexpr size_t kPartCacheBytes = (size_t) 1 << 30;
}  // namespace

extern "C" __attribute__((visibility("default"))) int ngm_b200_format_sam(const ngm_b200_encref *ref, const ngm_b200_sam_opts *opts,
		const ngm_b200_sam_batch *batch, char *out, size_t out_capacity, size_t *out_used) {
	if (ref == nullptr || opts == nullptr || batch == nullptr || out_used == nullptr || (out == nullptr && out_capacity)) return NGM_B200_EINVAL;
	const ngm_b200_sam_batch &b = *batch;
	if (b.n_reads < 0 || (b.n_reads && (b.reads == nullptr || b.quals == nullptr || b.names == nullptr || (b.best_pair == nullptr && b.topn <= 1) || b.mapq == nullptr ||
			b.num_top == nullptr || b.max_hit == nullptr || b.recs == nullptr))) return NGM_B200_EINVAL;
	const bool paired = b.pair_fail != nullptr;
	const bool multi = b.topn > 1;
	if (multi && (paired || b.sel == nullptr || b.n_sel == nullptr)) return NGM_B200_EINVAL;      // several alignments per read: single-end only, like the reference
	if (paired && (b.n_reads & 1)) return NGM_B200_EINVAL;
	// a mapped read dereferences its candidate, its score and its strings: refuse NULL arrays as soon as any read selects one
	if (b.pairs == nullptr || b.scores == nullptr || b.strings == nullptr) {
		for (int r = 0; r < b.n_reads; ++r)
			if (multi ? b.n_sel[r] > 0 : b.best_pair[r] >= 0) return NGM_B200_EINVAL;
	}
	try {
	const int units = paired ? b.n_reads / 2 : b.n_reads;
	int threads = opts->threads > 0 ? opts->threads : (int) std::thread::hardware_concurrency();
	threads = std::max(1, std::min(threads, std::max(1, units / 256)));
	// The per-thread output parts are kept from call to call (grow-only, at most kPartCacheBytes in total): a fresh allocation of this size
	// is unmapped memory whose first touch costs more than formatting the lines (measured: 1.2 -> 3.0 M reads/s per thread).  A call that
	// finds the cache in use by another thread formats into parts of its own.
	std::unique_lock<std::mutex> cache_lock(g_part_cache_mu, std::try_to_lock);
	std::vector<std::string> own_parts;
	std::vector<std::string> &parts = cache_lock.owns_lock() ? g_part_cache : own_parts;
	if (parts.size() < (size_t) threads) parts.resize((size_t) threads);
	for (int t = 0; t < threads; ++t) parts[(size_t) t].clear();
	struct CacheTrim {                                             // runs on every way out, before the lock is released
		std::vector<std::string> &p;
		bool cached;
		~CacheTrim() {
			if (!cached) return;
			size_t held = 0;
			for (const std::string &x : p) held += x.capacity();
			if (held > kPartCacheBytes) std::vector<std::string>().swap(p);
		}
	} trim = { parts, cache_lock.owns_lock() };
	const Job job = { ref, opts, batch };
	std::atomic<int> failed(0);                                    // an allocation failure inside a worker must not reach std::terminate
	auto work = [&](int t) {
		try {
			const int lo = (int) ((long long) units * t / threads), hi = (int) ((long long) units * (t + 1) / threads);
			std::string &s = parts[(size_t) t];
			s.reserve((size_t) (hi - lo) * (paired ? 2 : 1) * (size_t) (2 * b.stride + 160));
			std::vector<char> scratch(4096);
			for (int u = lo; u < hi; ++u) {
				if (paired) fragment(job, u, s, scratch);
				else if (multi) single_topn(job, u, s, scratch);
				else single(job, u, s, scratch);
			}
		} catch (...) {
			failed.store(1);
		}
	};
	auto run_all = [&](auto fn) {
		if (threads == 1) {
			fn(0);
		} else {
			std::vector<std::thread> pool;
			for (int t = 0; t < threads; ++t) pool.emplace_back(fn, t);
			for (std::thread &t : pool) t.join();
		}
	};
	run_all(work);
	if (failed.load()) return NGM_B200_EINVAL;
	size_t total = 0;
	std::vector<size_t> at((size_t) threads);
	for (int t = 0; t < threads; ++t) {
		at[(size_t) t] = total;
		total += parts[(size_t) t].size();
	}
	*out_used = total;
	if (total > out_capacity) return NGM_B200_ERANGE;
	run_all([&](int t) { memcpy(out + at[(size_t) t], parts[(size_t) t].data(), parts[(size_t) t].size()); });      // every thread places (and first-touches) its own part
	return b.n_reads;
	} catch (...) {                                                 // std::bad_alloc / std::system_error: no C++ exception crosses the C ABI
		return NGM_B200_EINVAL;
	}
}
