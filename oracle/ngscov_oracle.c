/* oracle/ngscov_oracle.c -- TEST INFRASTRUCTURE ONLY (see ngscov_oracle.h).
 *
 * CPU restatement of the reference coverage path.  It follows the reference's control flow
 * literally where the flow carries quirks (the sorted-key sweep of hash2BedGraph and both `overlap`
 * variants, cursor and back-off included) and replaces only the container: the two string-keyed hash
 * tables Start/End (reference hashtbl.c) become two sorted (key,count) arrays.  Keys keep the
 * reference's 28-bit aliasing (hashtbl.c:243-259 int2char/char2int keep 4x7 bits).
 *
 * Reference lines restated by each function are cited at the function.
 */
#include "ngscov_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <zlib.h>

/* ---------------------------------------------------------------- growable text buffer */
typedef struct { char *p; size_t n, cap; } tbuf;
static void tb_printf(tbuf *t, const char *fmt, ...)
{
	va_list ap;
	if (t->cap - t->n < 512) {
		t->cap = t->cap ? t->cap * 2 : 1 << 16;
		t->p = (char *)realloc(t->p, t->cap);
	}
	va_start(ap, fmt);
	int k = vsnprintf(t->p + t->n, t->cap - t->n, fmt, ap);
	va_end(ap);
	if ((size_t)k >= t->cap - t->n) {           /* very long contig name: grow and retry */
		t->cap = t->cap * 2 + (size_t)k;
		t->p = (char *)realloc(t->p, t->cap);
		va_start(ap, fmt);
		k = vsnprintf(t->p + t->n, t->cap - t->n, fmt, ap);
		va_end(ap);
	}
	t->n += (size_t)k;
}

/* ---------------------------------------------------------------- little-endian field access */
static inline uint32_t rd_u32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline int32_t  rd_i32(const uint8_t *p) { int32_t v;  memcpy(&v, p, 4); return v; }
static inline uint16_t rd_u16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

/* ---------------------------------------------------------------- the "hash tables" */
typedef struct { uint32_t *key; size_t n, cap; } keylist;      /* raw increments, later sorted */
typedef struct { int *key; int *cnt; size_t n; } counted;       /* distinct keys + counts (= HASHTBL) */

static void kl_push(keylist *k, uint32_t v)
{
	if (k->n == k->cap) { k->cap = k->cap ? k->cap * 2 : 1 << 16; k->key = (uint32_t *)realloc(k->key, k->cap * 4); }
	k->key[k->n++] = v;
}
static int cmp_u32(const void *a, const void *b)
{
	uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
	return x < y ? -1 : x > y;
}
static void kl_count(keylist *k, counted *c)
{
	qsort(k->key, k->n, 4, cmp_u32);
	c->key = (int *)malloc((k->n + 1) * sizeof(int));
	c->cnt = (int *)malloc((k->n + 1) * sizeof(int));
	c->n = 0;
	for (size_t i = 0; i < k->n; ++i) {
		if (c->n && (uint32_t)c->key[c->n - 1] == k->key[i]) c->cnt[c->n - 1]++;
		else { c->key[c->n] = (int)k->key[i]; c->cnt[c->n++] = 1; }
	}
}
/* hashtbl_get (reference hashtbl.c:167-180): NULL when the key is absent */
static const int *ct_get(const counted *c, int key)
{
	size_t lo = 0, hi = c->n;
	while (lo < hi) {
		size_t mid = (lo + hi) >> 1;
		if (c->key[mid] < key) lo = mid + 1; else hi = mid;
	}
	return (lo < c->n && c->key[lo] == key) ? &c->cnt[lo] : NULL;
}

/* int2char followed by char2int (reference hashtbl.c:243-259): only 4x7 = 28 bits of the position
 * survive as the hash key. */
static inline uint32_t key28(uint32_t pos) { return pos & 0x0FFFFFFFu; }

/* ---------------------------------------------------------------- per-run window accumulation */
typedef struct {
	int mode; uint32_t window; uint32_t target_len;
	int j, subject_count;             /* the reference's persistent cursor + back-off counter */
	int windows; double *bins;
} binstate;

/* `overlap`, depth variant (reference bam2depth.c:132-176) and wig variant (bam2wig.c:131-175),
 * restated literally including the persistent cursor j and the subject_count back-off. */
static void overlap(binstate *s, uint32_t last_start, uint32_t last_end, double last_depth)
{
	if (s->subject_count > 1) {
		if (s->j - s->subject_count >= 0) s->j -= s->subject_count; else s->j = 0;
	}
	uint32_t window_start, window_end;
	s->subject_count = 0;
	while (s->j <= s->windows) {
		window_start = s->window * (uint32_t)s->j;
		if (s->mode == ORA_MODE_DEPTH) {
			window_end = (uint32_t)(s->j + 1) * s->window;
			if (window_end > s->target_len) window_end = s->target_len;
			if (last_end < window_start) break;
			if (last_start <= window_start) {
				if (last_end <= window_end) { s->bins[s->j] += (last_end - window_start) * last_depth; s->subject_count++; break; }
				else { s->bins[s->j++] += (window_end - window_start) * last_depth; s->subject_count++; }
			} else if (last_start < window_end) {
				if (last_end <= window_end) { s->bins[s->j] += (last_end - last_start) * last_depth; s->subject_count++; break; }
				else { s->bins[s->j++] += (window_end - last_start) * last_depth; s->subject_count++; }
			} else s->j++;
		} else {
			window_end = (uint32_t)(s->j + 1) * s->window - 1;
			if (window_end > s->target_len) window_end = s->target_len;
			if (last_end < window_start) break;
			if (last_start < window_start) {
				if (last_end < window_end) { s->bins[s->j] += (last_end - window_start) * last_depth; s->subject_count++; break; }
				else { s->bins[s->j++] += (window_end - window_start + 1) * last_depth; s->subject_count++; }
			} else if (last_start <= window_end) {
				if (last_end <= window_end) { s->bins[s->j] += (last_end - last_start) * last_depth; s->subject_count++; break; }
				else { s->bins[s->j++] += (window_end - last_start) * last_depth; s->subject_count++; }
			} else s->j++;
		}
	}
}

/* hash2BedGraph (reference bam2depth.c:203-236, bam2wig.c:202-233): sweep the sorted union of the
 * Start and End keys (duplicates kept, as union_hashed_keys hashtbl.c:275-297 keeps them). */
static void hash2bedgraph(const counted *Start, const counted *End, const char *name, binstate *bs, tbuf *bedgraph)
{
	size_t all_keys_count = Start->n + End->n;
	if (!all_keys_count) return;
	int *all_keys = (int *)malloc(all_keys_count * sizeof(int));
	{ /* sorted concatenation == qsort(cmp_twoBitInt) of the concatenated key pointers */
		size_t a = 0, b = 0, o = 0;
		while (a < Start->n || b < End->n) {
			if (b >= End->n || (a < Start->n && Start->key[a] <= End->key[b])) all_keys[o++] = Start->key[a++];
			else all_keys[o++] = End->key[b++];
		}
	}
	int prevkey = 0, last_start = 0, last_end = 0, last_depth = 0, Count = 0, pos = 0;
	for (size_t i = 0; i < all_keys_count; ++i) {
		pos = all_keys[i];
		if (pos && pos == prevkey) continue;
		if (prevkey) {
			if (last_depth == Count) prevkey = last_start;
			else if (last_depth) {
				if (bedgraph) tb_printf(bedgraph, "%s\t%d\t%d\t%d\n", name, last_start, last_end, last_depth);
				overlap(bs, (uint32_t)last_start, (uint32_t)last_end, (double)last_depth);
			}
		}
		last_start = prevkey;
		last_end = pos;
		last_depth = Count;
		const int *d = ct_get(Start, pos);
		if (d) Count += *d;
		d = ct_get(End, pos);
		if (d) Count -= *d;
		prevkey = pos;
	}
	if (last_depth) {
		overlap(bs, (uint32_t)last_start, (uint32_t)last_end, (double)last_depth);
		if (bedgraph) tb_printf(bedgraph, "%s\t%d\t%d\t%d\n", name, last_start, last_end, last_depth);
	}
	free(all_keys);
}

/* output_bins (reference bam2depth.c:238-246) */
static void output_bins(tbuf *t, const char *name, uint32_t window, uint32_t target_len, const double *bins, int windows)
{
	for (int i = 0; i < windows; i++) {
		int window_start = (int)(window * (uint32_t)i);
		int window_end = window * (uint32_t)(i + 1) > target_len ? (int)target_len : (int)(window * (uint32_t)(i + 1));
		tb_printf(t, "%s\t%d\t%d\t%.2f\n", name, window_start, window_end, bins[i] / window);
	}
}
/* output_bins_wig (reference bam2depth.c:248-255, bam2wig.c:245-252) */
static void output_bins_wig(tbuf *t, const char *name, uint32_t window, const double *bins, int windows)
{
	tb_printf(t, "variableStep chrom=%s span=%d\n", name, (int)window);
	for (int i = 0; i < windows; i++) {
		int window_start = (int)(window * (uint32_t)i);
		if (bins[i]) tb_printf(t, "%d\t%.2f\n", window_start, bins[i] / window);
	}
}

int ngscov_oracle_run(const uint8_t *rec, size_t nbytes, int n_ref, const uint32_t *ref_len,
                      const char *const *ref_name, int mode, uint32_t window, int emit_wig,
                      ora_result *out)
{
	memset(out, 0, sizeof(*out));
	if (window == 0 || n_ref < 0) return -1;
	tbuf bed = {0}, dep = {0}, wig = {0}, chs = {0};
	/* one pass over the stream: per-contig Start/End key lists.  fetch_func (reference
	 * bam2depth.c:86-110, bam2wig.c:84-109) applied to every record whose refID is that contig
	 * (bam_fetch over [0,2^29): samtools bam_index.c:682-726 delivers each record of the contig whose
	 * pos < 2^29). */
	keylist *S = (keylist *)calloc((size_t)n_ref + 1, sizeof(keylist));
	keylist *E = (keylist *)calloc((size_t)n_ref + 1, sizeof(keylist));
	const uint32_t mask = mode == ORA_MODE_DEPTH ? 0x704u : 0x4u;
	size_t off = 0;
	int rc = 0;
	while (off + 4 <= nbytes) {
		int32_t block_size = rd_i32(rec + off);
		if (block_size < 32 || off + 4 + (size_t)block_size > nbytes) { rc = -2; break; }
		const uint8_t *c = rec + off + 4;
		int32_t tid = rd_i32(c), pos = rd_i32(c + 4);
		uint32_t l_qname = c[8], n_cigar = rd_u16(c + 12), flag = rd_u16(c + 14);
		out->n_records++;
		out->cigar_words += n_cigar;
		if (32 + l_qname + 4 * (size_t)n_cigar > (size_t)block_size) { rc = -3; break; }
		if (!(flag & mask) && tid >= 0 && tid < n_ref && pos < (1 << 29)) {
			out->n_kept++;
			const uint8_t *cig = c + 32 + l_qname;
			unsigned int temp_start = (unsigned int)pos;
			for (uint32_t i = 0; i < n_cigar; ++i) {
				uint32_t w = rd_u32(cig + 4 * i), op = w & 0xf, l = w >> 4;
				if (op == 1) continue;                       /* BAM_CINS */
				if (op == 2 || op == 3) temp_start += l;       /* BAM_CDEL, BAM_CREF_SKIP */
				else if (op == 0) {                           /* BAM_CMATCH */
					kl_push(&S[tid], key28(temp_start));
					temp_start += l;
					kl_push(&E[tid], key28(temp_start));
					out->n_mops++;
				}                                             /* S H P = X B: ignored, no advance */
			}
		}
		off += 4 + (size_t)block_size;
	}
	if (rc == 0 && off != nbytes) rc = -4;
	for (int j = 0; j < n_ref && rc == 0; ++j) {                /* main loop, reference bam2depth.c:325-339 */
		counted Start, End;
		kl_count(&S[j], &Start);
		kl_count(&E[j], &End);
		binstate bs;
		memset(&bs, 0, sizeof(bs));
		bs.mode = mode; bs.window = window; bs.target_len = ref_len[j];
		bs.windows = (int)(ref_len[j] / window + 1);
		/* the reference callocs exactly `windows` doubles and can touch bins[windows] for reads
		 * overhanging far beyond the contig end (undefined there); keep one spare slot so the
		 * oracle stays defined, it is never printed. */
		bs.bins = (double *)calloc((size_t)bs.windows + 2, sizeof(double));
		hash2bedgraph(&Start, &End, ref_name[j], &bs, mode == ORA_MODE_DEPTH ? &bed : NULL);
		if (mode == ORA_MODE_DEPTH) output_bins(&dep, ref_name[j], window, ref_len[j], bs.bins, bs.windows);
		if (mode == ORA_MODE_WIG || emit_wig) {
			output_bins_wig(&wig, ref_name[j], window, bs.bins, bs.windows);
			tb_printf(&chs, "%s\t%d\n", ref_name[j], (int)ref_len[j]);
		}
		free(bs.bins); free(Start.key); free(Start.cnt); free(End.key); free(End.cnt);
	}
	for (int j = 0; j <= n_ref; ++j) { free(S[j].key); free(E[j].key); }
	free(S); free(E);
	out->bedgraph = bed.p; out->bedgraph_len = bed.n;
	out->depth = dep.p; out->depth_len = dep.n;
	out->wig = wig.p; out->wig_len = wig.n;
	out->chromsize = chs.p; out->chromsize_len = chs.n;
	return rc;
}

void ngscov_oracle_free(ora_result *r)
{
	free(r->bedgraph); free(r->depth); free(r->wig); free(r->chromsize);
	memset(r, 0, sizeof(*r));
}

/* ---------------------------------------------------------------- BGZF/BAM loading
 * Block layout per samtools bgzf.c:240-247,307-340: 18-byte gzip header with the 'B','C' extra field
 * carrying BSIZE = block_length-1 at byte 16, raw DEFLATE payload, 8-byte CRC32/ISIZE footer (not
 * verified by the reference reader either). */
int ngscov_oracle_load_bam(const char *path, ora_bam *bam)
{
	memset(bam, 0, sizeof(*bam));
	FILE *f = fopen(path, "rb");
	if (!f) return -1;
	fseek(f, 0, SEEK_END);
	long fsz = ftell(f);
	fseek(f, 0, SEEK_SET);
	uint8_t *comp = (uint8_t *)malloc((size_t)fsz + 1);
	if (fread(comp, 1, (size_t)fsz, f) != (size_t)fsz) { fclose(f); free(comp); return -1; }
	fclose(f);
	size_t cap = (size_t)fsz * 4 + (1 << 16), n = 0, off = 0;
	uint8_t *out = (uint8_t *)malloc(cap);
	while (off + 18 <= (size_t)fsz) {
		const uint8_t *h = comp + off;
		if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4) || h[12] != 'B' || h[13] != 'C') { free(comp); free(out); return -2; }
		size_t bsize = (size_t)rd_u16(h + 16) + 1;
		if (off + bsize > (size_t)fsz) { free(comp); free(out); return -2; }
		if (cap - n < 65536) { cap *= 2; out = (uint8_t *)realloc(out, cap); }
		z_stream zs;
		memset(&zs, 0, sizeof(zs));
		zs.next_in = (Bytef *)(h + 18); zs.avail_in = (uInt)(bsize - 18 - 8);
		zs.next_out = out + n; zs.avail_out = 65536;
		if (inflateInit2(&zs, -15) != Z_OK) { free(comp); free(out); return -3; }
		int zr = inflate(&zs, Z_FINISH);
		inflateEnd(&zs);
		if (zr != Z_STREAM_END) { free(comp); free(out); return -3; }
		n += zs.total_out;
		off += bsize;
	}
	free(comp);
	bam->inflated = out; bam->inflated_len = n;
	/* header, samtools bam.c:81-121 */
	if (n < 12 || memcmp(out, "BAM\1", 4)) return -4;
	size_t p = 4;
	uint32_t l_text = rd_u32(out + p); p += 4 + l_text;
	if (p + 4 > n) return -4;
	bam->n_ref = (int)rd_u32(out + p); p += 4;
	bam->ref_len = (uint32_t *)calloc((size_t)bam->n_ref + 1, 4);
	bam->ref_name = (char **)calloc((size_t)bam->n_ref + 1, sizeof(char *));
	for (int i = 0; i < bam->n_ref; ++i) {
		if (p + 4 > n) return -4;
		uint32_t l_name = rd_u32(out + p); p += 4;
		if (p + l_name + 4 > n) return -4;
		bam->ref_name[i] = strndup((const char *)out + p, l_name); p += l_name;
		bam->ref_len[i] = rd_u32(out + p); p += 4;
	}
	bam->records = out + p; bam->records_len = n - p;
	return 0;
}

void ngscov_oracle_free_bam(ora_bam *bam)
{
	for (int i = 0; i < bam->n_ref; ++i) free(bam->ref_name[i]);
	free(bam->ref_name); free(bam->ref_len); free(bam->inflated);
	memset(bam, 0, sizeof(*bam));
}

#ifdef NGSCOV_ORACLE_MAIN
/* ngscov_oracle depth|wig [-o PFX] [-w W] [-W] a.bam [b.bam ..]  -- writes the same file names as the
 * reference tools (reference bam2depth.c:307-321, bam2wig.c:300-307) so outputs can be cmp'ed. */
#include <libgen.h>
#include <time.h>
static void dump(const char *pfx, const char *suffix, const char *buf, size_t len)
{
	char name[4096];
	snprintf(name, sizeof name, "%s%s", pfx, suffix);
	FILE *f = fopen(name, "wb");
	if (!f) { perror(name); exit(1); }
	fwrite(buf, 1, len, f);
	fclose(f);
}
int main(int argc, char **argv)
{
	if (argc < 3) { fprintf(stderr, "usage: ngscov_oracle depth|wig [-o PFX] [-w W] [-W] a.bam ..\n"); return 1; }
	int mode = !strcmp(argv[1], "wig") ? ORA_MODE_WIG : ORA_MODE_DEPTH, emit_wig = 0, a = 2, nfile = 0;
	const char *pfx = "out"; uint32_t window = 20000;
	for (; a < argc; ++a) {
		if (!strcmp(argv[a], "-o") && a + 1 < argc) pfx = argv[++a];
		else if (!strcmp(argv[a], "-w") && a + 1 < argc) window = (uint32_t)atoi(argv[++a]);
		else if (!strcmp(argv[a], "-W")) emit_wig = 1;
		else {
			ora_bam bam; ora_result r; char suffix[64];
			struct timespec t0, t1;
			if (ngscov_oracle_load_bam(argv[a], &bam)) { fprintf(stderr, "cannot load %s\n", argv[a]); return 1; }
			clock_gettime(CLOCK_MONOTONIC, &t0);
			int rc = ngscov_oracle_run(bam.records, bam.records_len, bam.n_ref, bam.ref_len,
			                           (const char *const *)bam.ref_name, mode, window, emit_wig, &r);
			clock_gettime(CLOCK_MONOTONIC, &t1);
			if (rc) { fprintf(stderr, "oracle failed on %s: %d\n", argv[a], rc); return 1; }
			++nfile;
			if (mode == ORA_MODE_DEPTH) {
				char *tmp = strdup(argv[a]);
				snprintf(suffix, sizeof suffix, ".%d.bedGraph", nfile); dump(basename(tmp), suffix, r.bedgraph, r.bedgraph_len);
				free(tmp);
				snprintf(suffix, sizeof suffix, ".%d.depth", nfile); dump(pfx, suffix, r.depth, r.depth_len);
			}
			if (mode == ORA_MODE_WIG || emit_wig) {
				snprintf(suffix, sizeof suffix, ".%d.wig", nfile); dump(pfx, suffix, r.wig, r.wig_len);
				snprintf(suffix, sizeof suffix, ".%d.chromSize.txt", nfile); dump(pfx, suffix, r.chromsize, r.chromsize_len);
			}
			fprintf(stderr, "oracle: %s: %llu records, %llu kept, %llu M ops, %.3f s\n", argv[a],
			        (unsigned long long)r.n_records, (unsigned long long)r.n_kept, (unsigned long long)r.n_mops,
			        (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec));
			ngscov_oracle_free(&r);
			ngscov_oracle_free_bam(&bam);
		}
	}
	return 0;
}
#endif
