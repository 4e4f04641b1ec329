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

// _SequenceProvider::writeEncRefToFile (SequenceProvider.cpp:189-208): the file NGM reads back instead of encoding the FASTA again
extern "C" __attribute__((visibility("default"))) int ngm_b200_write_enc_ref(const char *path, const ngm_b200_encref *ref) {
	if (path == nullptr || ref == nullptr || ref->packed == nullptr || (ref->n_contigs && ref->contigs == nullptr)) return NGM_B200_EINVAL;
	FILE *fp = fopen(path, "wb");
	if (fp == nullptr) return NGM_B200_EINVAL;
	const uint32_t cookie = kRefEncCookie, ref_count = ref->n_contigs;
	const uint64_t bin_ref_index = ref->concat_len + 1, enc_size = ref->packed_bytes;
	bool ok = fwrite(&cookie, 4, 1, fp) == 1 && fwrite(&ref_count, 4, 1, fp) == 1 && fwrite(&bin_ref_index, 8, 1, fp) == 1 && fwrite(&enc_size, 8, 1, fp) == 1;
	for (uint32_t i = 0; ok && i < ref_count; ++i) {
		DiskRefIdx d;
		memset(&d, 0, sizeof(d));
		d.SeqId = i;
		d.SeqStart = ref->contigs[i].start;
		d.SeqLen = ref->contigs[i].length;
		d.NameLen = ref->contigs[i].name_len > 100 ? 100 : ref->contigs[i].name_len;
		memcpy(d.name, ref->contigs[i].name, 100);
		ok = fwrite(&d, sizeof(d), 1, fp) == 1;
	}
	if (ok) ok = fwrite(ref->packed, 1, enc_size, fp) == enc_size;
	ok = fclose(fp) == 0 && ok;
	return ok ? NGM_B200_OK : NGM_B200_EINVAL;
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

// ---- prefix-table cache file `<ref>-ht-<k>-<skip>.3.ngm` (PrefixTable.cpp:819-921) --------------------
// layout: cookie 0x74656, kmer, kmer_skip, unit count, index size; per unit: cRefTableLen, Index[index size] (packed
// {uint32 m_TabIndex; char m_RevCompIndex}, PrefixTable.h:17-33), Location[cRefTableLen], uint64 Offset; then the
// signature cookie + kmer + kmer_skip + units + index size.
extern "C" __attribute__((visibility("default"))) int ngm_b200_read_ht_file(const char *path, ngm_b200_htfile *out) {
	if (path == nullptr || out == nullptr) return NGM_B200_EINVAL;
	memset(out, 0, sizeof(*out));
	FILE *fp = fopen(path, "rb");
	if (fp == nullptr) return NGM_B200_EINVAL;
	uint32_t hdr[5];
	int rc = NGM_B200_EINVAL;
	uint8_t *raw = nullptr;
	do {
		if (fread(hdr, 4, 5, fp) != 5 || hdr[0] != kRefEncCookie || hdr[3] != 1) break;      // one table unit (< 4 Gbp)
		out->kmer = hdr[1];
		out->kmer_skip = hdr[2];
		out->index_len = hdr[4];
		if (fread(&out->table_len, 4, 1, fp) != 1) break;
		const size_t n = out->index_len;
		raw = (uint8_t *) malloc(n * 5);
		out->tab = (uint32_t *) malloc(n * 4);
		out->weight = (int8_t *) malloc(n);
		out->table = (uint32_t *) malloc((size_t) out->table_len * 4 + 4);
		if (!raw || !out->tab || !out->weight || !out->table) break;
		if (fread(raw, 5, n, fp) != n) break;
		for (size_t i = 0; i < n; ++i) {
			memcpy(&out->tab[i], raw + 5 * i, 4);
			out->weight[i] = (int8_t) raw[5 * i + 4];
		}
		if (fread(out->table, 4, out->table_len, fp) != out->table_len) break;
		uint32_t sig = 0;
		if (fread(&out->unit_offset, 8, 1, fp) != 1 || fread(&sig, 4, 1, fp) != 1) break;
		if (sig != hdr[0] + hdr[1] + hdr[2] + hdr[3] + hdr[4]) break;
		rc = NGM_B200_OK;
	} while (false);
	free(raw);
	fclose(fp);
	if (rc != NGM_B200_OK) ngm_b200_free_ht_file(out);
	return rc;
}

extern "C" __attribute__((visibility("default"))) void ngm_b200_free_ht_file(ngm_b200_htfile *ht) {
	if (ht == nullptr) return;
	free(ht->tab);
	free(ht->weight);
	free(ht->table);
	memset(ht, 0, sizeof(*ht));
}

extern "C" __attribute__((visibility("default"))) int ngm_b200_write_ht_file(const char *path, const ngm_b200_htfile *ht) {
	if (path == nullptr || ht == nullptr || ht->tab == nullptr || ht->weight == nullptr) return NGM_B200_EINVAL;
	FILE *fp = fopen(path, "wb");
	if (fp == nullptr) return NGM_B200_EINVAL;
	const uint32_t hdr[5] = { kRefEncCookie, ht->kmer, ht->kmer_skip, 1u, ht->index_len };
	bool ok = fwrite(hdr, 4, 5, fp) == 5 && fwrite(&ht->table_len, 4, 1, fp) == 1;
	const size_t n = ht->index_len;
	uint8_t *raw = (uint8_t *) malloc(n * 5);
	ok = ok && raw != nullptr;
	if (ok) {
		for (size_t i = 0; i < n; ++i) {
			memcpy(raw + 5 * i, &ht->tab[i], 4);
			raw[5 * i + 4] = (uint8_t) ht->weight[i];
		}
		ok = fwrite(raw, 5, n, fp) == n;
	}
	free(raw);
	ok = ok && (ht->table_len == 0 || fwrite(ht->table, 4, ht->table_len, fp) == ht->table_len);
	const uint32_t sig = hdr[0] + hdr[1] + hdr[2] + hdr[3] + hdr[4];
	ok = ok && fwrite(&ht->unit_offset, 8, 1, fp) == 1 && fwrite(&sig, 4, 1, fp) == 1;
	fclose(fp);
	return ok ? NGM_B200_OK : NGM_B200_EINVAL;
}
