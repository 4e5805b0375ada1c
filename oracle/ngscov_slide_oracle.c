/* oracle/ngscov_slide_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's bam_sliding_count text path: per window
 * of every contig the number of mapped reads STARTING in it, the sum of their l_qseq and the GC content of their SEQ;
 * the PNG the reference draws with libgd is out of scope.  Follows /root/reference/bam_sliding_count.c:
 *   fetch_func        :93-124   (tid < 0 or BAM_FUNMAP dropped; window = (unsigned short)(pos / window); GC from the
 *                                4-bit SEQ codes 2 (C) and 4 (G) in an unsigned short; float accumulation of GC)
 *   cal_GC            :84-91
 *   calc_winGC        :126-138  (float arithmetic, sequential float sum over the windows of a contig)
 *   output_count_GC   :140-165  (header with max window count, "%s\t%d\t%u\t%lu\t%f\t%f" + "\t%d\t%u\t%f" per window)
 *   max               :198-205
 * Parity status: PINNED against the unmodified reference binary (oracle/_ref/bam_sliding_count, built with the libgd
 * stubs of oracle/gdstub/): tests/test_oracle_vs_reference.py runs both here, tests/golden/slide_outputs.json travels.
 * Only tests/ may call this; the product path never links it. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>

typedef struct { char *p; size_t len, cap; } sbuf;
static void sb_printf(sbuf *b, const char *fmt, ...)
{
	char tmp[512];
	va_list ap; va_start(ap, fmt);
	int n = vsnprintf(tmp, sizeof tmp, fmt, ap);
	va_end(ap);
	if (n < 0) return;
	char *src = tmp, *big = NULL;
	if ((size_t)n >= sizeof tmp) { big = (char *)malloc((size_t)n + 1); va_start(ap, fmt); vsnprintf(big, (size_t)n + 1, fmt, ap); va_end(ap); src = big; }
	if (b->len + (size_t)n + 1 > b->cap) { b->cap = (b->len + (size_t)n + 1) * 2; b->p = (char *)realloc(b->p, b->cap); }
	memcpy(b->p + b->len, src, (size_t)n); b->len += (size_t)n; b->p[b->len] = 0;
	free(big);
}

/* records: the inflated BAM record stream.  out: malloc'ed text of <outfile>.txt.  n_count: reads counted (:107).
 * Returns 0, -1 on a malformed stream. */
int ngscov_oracle_slide_region(const uint8_t *rec, size_t nbytes, int n_ref, const uint32_t *ref_len, const char *const *ref_name,
                               int window, int reg_ref, int reg_beg, int reg_end, char **out, size_t *out_len, uint64_t *n_count);
int ngscov_oracle_slide(const uint8_t *rec, size_t nbytes, int n_ref, const uint32_t *ref_len, const char *const *ref_name,
                        int window, char **out, size_t *out_len, uint64_t *n_count)
{
	return ngscov_oracle_slide_region(rec, nbytes, n_ref, ref_len, ref_name, window, -1, 0, 0, out, out_len, n_count);
}

/* The same with the `-r` mode (bam_sliding_count.c:393-409): only what bam_fetch(reg_ref, reg_beg, reg_end) hands to
 * fetch_func is counted.  On a coordinate-sorted file that is every record with tid == reg_ref, (uint32) pos < end and
 * (uint32) rend > beg, rend = bam_calend (reference-consuming CIGAR operations) or pos + 1 without a CIGAR
 * (samtools-0.1.19 bam_index.c bam_iter_read / is_overlap).  reg_ref < 0: no region. */
int ngscov_oracle_slide_region(const uint8_t *rec, size_t nbytes, int n_ref, const uint32_t *ref_len, const char *const *ref_name,
                               int window, int reg_ref, int reg_beg, int reg_end, char **out, size_t *out_len, uint64_t *n_count)
{
	uint32_t *windows = (uint32_t *)calloc((size_t)n_ref + 1, sizeof(uint32_t));
	unsigned int **bins = (unsigned int **)calloc((size_t)n_ref + 1, sizeof(*bins)), **len = (unsigned int **)calloc((size_t)n_ref + 1, sizeof(*len));
	float **GC = (float **)calloc((size_t)n_ref + 1, sizeof(*GC));
	unsigned int *sum_count = (unsigned int *)calloc((size_t)n_ref + 1, sizeof(unsigned int));
	unsigned long *sum_base = (unsigned long *)calloc((size_t)n_ref + 1, sizeof(unsigned long));
	float *sum_GC = (float *)calloc((size_t)n_ref + 1, sizeof(float));
	uint64_t count = 0;
	int rc = 0;
	size_t o = 0;
	while (o + 4 <= nbytes) {
		int32_t bs; memcpy(&bs, rec + o, 4);
		if (bs < 32 || o + 4 + (size_t)bs > nbytes) { rc = -1; break; }
		const uint8_t *r = rec + o + 4;
		int32_t tid, pos, l_qseq; uint32_t bin_mq_nl, flag_nc;
		memcpy(&tid, r, 4); memcpy(&pos, r + 4, 4); memcpy(&bin_mq_nl, r + 8, 4); memcpy(&flag_nc, r + 12, 4); memcpy(&l_qseq, r + 16, 4);
		o += 4 + (size_t)bs;
		const uint32_t l_qname = bin_mq_nl & 0xff, n_cigar = flag_nc & 0xffff, flag = flag_nc >> 16;
		if (reg_ref >= 0) {
			if (tid != reg_ref || (uint32_t)pos >= (uint32_t)reg_end) continue;
			uint32_t rend = (uint32_t)pos + 1;
			if (n_cigar) {
				int end = pos;
				for (uint32_t k = 0; k < n_cigar; ++k) {
					uint32_t w; memcpy(&w, r + 32 + l_qname + 4 * (size_t)k, 4);
					if ((0x3C1A7 >> ((w & 0xf) << 1)) & 2) end += (int)(w >> 4);
				}
				rend = (uint32_t)end;
			}
			if (!(rend > (uint32_t)reg_beg)) continue;
		}
		if (tid < 0) continue;                                        /* :96 */
		if (flag & 4) continue;                                       /* :97 BAM_FUNMAP */
		if (tid >= n_ref) { rc = -1; break; }
		if (windows[tid] == 0) {                                      /* :98-103 */
			windows[tid] = (unsigned int)ref_len[tid] / window + 1;
			bins[tid] = (unsigned int *)calloc(windows[tid], sizeof(unsigned int));
			GC[tid] = (float *)calloc(windows[tid], sizeof(float));
			len[tid] = (unsigned int *)calloc(windows[tid], sizeof(unsigned int));
		}
		++count;                                                      /* :104 */
		const unsigned short current_window = (unsigned short)(pos / window);      /* :116 */
		unsigned short current_GC = 0;
		const uint8_t *seq = r + 32 + l_qname + 4 * (size_t)n_cigar;
		for (int32_t i = 0; i < l_qseq; ++i) {                        /* cal_GC :84-91, bam1_seqi */
			const int t = seq[i >> 1] >> ((~i & 1) << 2) & 0xf;
			if (t == 2 || t == 4) ++current_GC;
		}
		if (current_window >= windows[tid]) continue;                 /* the reference writes out of bounds here (POS beyond the contig end) */
		bins[tid][current_window]++;                                  /* :119-121 */
		GC[tid][current_window] += current_GC;
		len[tid][current_window] += (unsigned int)l_qseq;
	}
	sbuf b = { NULL, 0, 0 };
	if (!rc) {
		for (int j = 0; j < n_ref; ++j) {                             /* calc_winGC :126-138 */
			for (uint32_t k = 0; k < windows[j]; ++k) {
				sum_count[j] += bins[j][k];
				sum_GC[j] += GC[j][k];
				sum_base[j] += len[j][k];
				GC[j][k] = GC[j][k] != 0 ? GC[j][k] / len[j][k] * 100 : GC[j][k];
			}
			sum_GC[j] = sum_GC[j] / sum_base[j] * 100;
		}
		sb_printf(&b, "#chr\tchr_len\tchr_sum_read_count\tchr_sum_base\tchr_mean_cov\tchr_mean_GC%%");      /* output_count_GC :140-165 */
		uint32_t max_len = 0;
		for (unsigned short i = 0; i < n_ref; ++i) max_len = ref_len[i] > max_len ? ref_len[i] : max_len;
		const uint32_t max_target_window = max_len / window + 1;
		for (uint32_t k = 0; k < max_target_window; ++k) sb_printf(&b, "\t%u\tcount\tGC%%", k + 1);
		sb_printf(&b, "\n");
		for (int j = 0; j < n_ref; ++j) {
			if (windows[j] == 0) continue;
			sb_printf(&b, "%s\t%d\t%u\t%lu\t%f\t%f", ref_name[j], ref_len[j], sum_count[j], sum_base[j], (double)sum_base[j] / ref_len[j], sum_GC[j]);
			for (uint32_t k = 0; k < windows[j]; ++k) sb_printf(&b, "\t%d\t%u\t%f", k + 1, bins[j][k], GC[j][k]);
			sb_printf(&b, "\n");
		}
	}
	for (int j = 0; j < n_ref; ++j) { free(bins[j]); free(GC[j]); free(len[j]); }
	free(windows); free(bins); free(len); free(GC); free(sum_count); free(sum_base); free(sum_GC);
	if (rc) { free(b.p); return rc; }
	*out = b.p ? b.p : (char *)calloc(1, 1); *out_len = b.len; *n_count = count;
	return 0;
}

void ngscov_oracle_slide_free(char *p) { free(p); }
