/* host/ngscov_shard.h -- host-side logic of the region-sharded drivers (NGSCOV_GPUS=N), kept apart from the I/O so
 * that it can be unit-tested without a GPU (tests/shard_check.c against highperformancengs_b200/shard.py):
 * BAI loading (samtools-0.1.19 bam_index.c:326-387 layout), slice planning, virtual-offset lookup, halo stitch.
 * No counterpart in the single-process reference. */
#ifndef NGSCOV_SHARD_H
#define NGSCOV_SHARD_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "../../include/ngscov.h"

#define NGS_TILE_BP 16384u
#define NGS_END_MARGIN_TILES 4u

typedef struct {
	int n_ref;
	uint64_t **lin; int *n_lin;          /* linear index per contig: virtual offsets */
	uint64_t *first;                     /* smallest chunk begin of a contig (0: no records) */
} bai_t;

static int bai_load(const char *bam, bai_t *b)
{
	size_t l = strlen(bam);
	char *a = (char *)malloc(l + 8);
	sprintf(a, "%s.bai", bam);
	FILE *f = fopen(a, "rb");
	if (!f && l > 4 && !strcmp(bam + l - 4, ".bam")) { strcpy(a, bam); strcpy(a + l - 3, "bai"); f = fopen(a, "rb"); }
	free(a);
	if (!f) return -1;
	char magic[4]; int32_t n_ref;
	if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "BAI\1", 4) || fread(&n_ref, 4, 1, f) != 1 || n_ref < 0) { fclose(f); return -1; }
	b->n_ref = n_ref;
	b->lin = (uint64_t **)calloc((size_t)n_ref + 1, sizeof(uint64_t *)); b->n_lin = (int *)calloc((size_t)n_ref + 1, sizeof(int));
	b->first = (uint64_t *)calloc((size_t)n_ref + 1, 8);
	for (int i = 0; i < n_ref; ++i) {
		int32_t n_bin;
		if (fread(&n_bin, 4, 1, f) != 1) { fclose(f); return -1; }
		for (int k = 0; k < n_bin; ++k) {
			uint32_t bin; int32_t n_chunk;
			if (fread(&bin, 4, 1, f) != 1 || fread(&n_chunk, 4, 1, f) != 1) { fclose(f); return -1; }
			for (int c = 0; c < n_chunk; ++c) {
				uint64_t ch[2];
				if (fread(ch, 8, 2, f) != 2) { fclose(f); return -1; }
				if (bin != 37450u && (b->first[i] == 0 || ch[0] < b->first[i])) b->first[i] = ch[0];   /* 37450: samtools' metadata pseudo-bin */
			}
		}
		int32_t n_intv;
		if (fread(&n_intv, 4, 1, f) != 1 || n_intv < 0) { fclose(f); return -1; }
		b->n_lin[i] = n_intv;
		b->lin[i] = (uint64_t *)calloc((size_t)n_intv + 1, 8);
		if (n_intv && fread(b->lin[i], 8, (size_t)n_intv, f) != (size_t)n_intv) { fclose(f); return -1; }
	}
	fclose(f);
	return 0;
}

/* virtual offset of the first record that can matter from contig c, position p on (0: nothing follows) */
static uint64_t bai_seek(const bai_t *b, int c, uint32_t p)
{
	if (c >= b->n_ref) return 0;
	for (int w = (int)(p >> 14); w < b->n_lin[c]; ++w) if (b->lin[c][w]) return b->lin[c][w];
	for (int c2 = c + 1; c2 < b->n_ref; ++c2) if (b->first[c2]) return b->first[c2];
	return 0;
}

typedef struct { uint64_t beg, end; } slice_t;

/* same cuts as highperformancengs_b200/shard.py plan_slices */
static void plan_slices(const uint32_t *len, int n_ref, int n, slice_t *out)
{
	uint64_t G = 0, prev = 0;
	for (int i = 0; i < n_ref; ++i) G += len[i];
	for (int r = 1; r <= n; ++r) {
		uint64_t cut = G;
		if (r < n) {
			const uint64_t g = (uint64_t)(((unsigned __int128)G * (unsigned)r) / (unsigned)n);
			uint64_t acc = 0; int c = 0;
			while (c + 1 < n_ref && acc + len[c] <= g) { acc += len[c]; ++c; }
			cut = acc + (g - acc) / NGS_TILE_BP * NGS_TILE_BP;
			if (cut < prev) cut = prev;
		}
		out[r - 1].beg = prev; out[r - 1].end = cut; prev = cut;
	}
}

static void locate(const uint32_t *len, int n_ref, uint64_t g, int *c, uint32_t *p)
{
	uint64_t acc = 0;
	for (int i = 0; i < n_ref; ++i) { if (g < acc + len[i]) { *c = i; *p = (uint32_t)(g - acc); return; } acc += len[i]; }
	*c = n_ref; *p = 0;
}


/* decisions for ngscov_halo_put and the window-bin corrections; same rule as shard.py stitch() */
typedef struct { ngscov_run run; int sign; } corr_t;
static size_t stitch(const ngscov_halo *halo, int n, int *drop_first, uint32_t *new_last_end, corr_t *corr)
{
	size_t nc = 0; int have = 0, owner = 0; ngscov_run carry; memset(&carry, 0, sizeof carry);
	for (int r = 0; r < n; ++r) { drop_first[r] = 0; new_last_end[r] = 0; }
	for (int r = 0; r < n; ++r) {
		if (halo[r].n_runs == 0) { have = 0; continue; }
		const ngscov_run f = halo[r].first;
		int merged = 0;
		if (have && carry.contig == f.contig && carry.end == f.start && carry.depth == f.depth) {
			merged = 1;
			corr[nc].run = carry; corr[nc++].sign = -1;
			corr[nc].run = f; corr[nc++].sign = -1;
			carry.end = f.end;
			corr[nc].run = carry; corr[nc++].sign = +1;
			drop_first[r] = 1; new_last_end[owner] = carry.end;
		}
		if (merged && halo[r].n_runs == 1) continue;       /* the whole shard was one run: it stays open into the next */
		carry = halo[r].last; owner = r; have = 1;
	}
	return nc;
}

#endif
