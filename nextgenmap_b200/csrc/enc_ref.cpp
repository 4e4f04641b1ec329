// enc_ref.cpp -- reader of NextGenMap's encoded-reference cache file and the concat -> contig mapping
// (reference src/SequenceProvider.cpp:111-141,143-208).  Host only.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/ngm_b200.h"

namespace {

// on-disk RefIdx (SequenceProvider.h:45-52) as laid out by the x86-64 ABI the reference is built with
struct DiskRefIdx {
	uint32_t SeqId;
	uint32_t Flags;
	uint64_t SeqStart;
	uint32_t SeqLen;
	uint32_t NameLen;
	char name[100];
	char pad[4];
};
static_assert(sizeof(DiskRefIdx) == 128, "RefIdx is 128 bytes on LP64");

const uint32_t kRefEncCookie = 0x74656;   // SequenceProvider.cpp:41

}  // namespace

extern "C" __attribute__((visibility("default"))) int ngm_b200_read_enc_ref(const char *path, ngm_b200_encref *out) {
	if (path == nullptr || out == nullptr) return NGM_B200_EINVAL;
	memset(out, 0, sizeof(*out));
	FILE *fp = fopen(path, "rb");
	if (fp == nullptr) return NGM_B200_EINVAL;
	uint32_t cookie = 0, ref_count = 0;
	uint64_t bin_ref_index = 0, enc_size = 0;
	bool ok = fread(&cookie, 4, 1, fp) == 1 && fread(&ref_count, 4, 1, fp) == 1 && fread(&bin_ref_index, 8, 1, fp) == 1 &&
			fread(&enc_size, 8, 1, fp) == 1 && cookie == kRefEncCookie && bin_ref_index >= 1 && enc_size * 2 >= bin_ref_index;
	if (ok) {
		out->contigs = static_cast<ngm_b200_contig *>(calloc(ref_count ? ref_count : 1, sizeof(ngm_b200_contig)));
		out->packed = static_cast<uint8_t *>(malloc(enc_size ? enc_size : 1));
		ok = out->contigs != nullptr && out->packed != nullptr;
	}
	for (uint32_t i = 0; ok && i < ref_count; ++i) {
		DiskRefIdx d;
		ok = fread(&d, sizeof(d), 1, fp) == 1;
		if (ok) {
			out->contigs[i].start = d.SeqStart;
			out->contigs[i].length = d.SeqLen;
			out->contigs[i].name_len = d.NameLen > 100 ? 100 : d.NameLen;
			memcpy(out->contigs[i].name, d.name, 100);
		}
	}
	if (ok) ok = fread(out->packed, 1, enc_size, fp) == enc_size;
	fclose(fp);
	if (!ok) {
		ngm_b200_free_enc_ref(out);
		return NGM_B200_EINVAL;
	}
	out->n_contigs = ref_count;
	out->packed_bytes = enc_size;
	out->concat_len = bin_ref_index - 1;
	return NGM_B200_OK;
}

extern "C" __attribute__((visibility("default"))) void ngm_b200_free_enc_ref(ngm_b200_encref *ref) {
	if (ref == nullptr) return;
	free(ref->packed);
	free(ref->contigs);
	memset(ref, 0, sizeof(*ref));
}

extern "C" __attribute__((visibility("default"))) int ngm_b200_convert(const ngm_b200_encref *ref, uint64_t concat_pos, uint32_t *contig, uint64_t *pos) {
	if (ref == nullptr || ref->n_contigs == 0) return 0;
	// refStartPos[] = contig starts + one artificial upper bound (SequenceProvider.cpp:369-378)
	const uint32_t n = ref->n_contigs;
	const uint64_t last_bound = ref->contigs[n - 1].start + ref->contigs[n - 1].length + 1000;
	// upper_bound over [starts..., last_bound]
	uint32_t lo = 0, hi = n + 1;
	while (lo < hi) {
		const uint32_t mid = (lo + hi) / 2;
		const uint64_t v = mid < n ? ref->contigs[mid].start : last_bound;
		if (v <= concat_pos) lo = mid + 1; else hi = mid;
	}
	const uint32_t upper = lo;                         // first entry > concat_pos
	if (upper == 0 || upper > n) return 0;             // before the first contig / beyond the bound: not a valid mapping position
	const uint64_t upper_val = upper < n ? ref->contigs[upper].start : last_bound;
	if (upper_val - concat_pos < 1000) return 0;       // inside the spacer in front of the next contig
	if (contig) *contig = upper - 1;
	if (pos) *pos = concat_pos - ref->contigs[upper - 1].start;
	return 1;
}
