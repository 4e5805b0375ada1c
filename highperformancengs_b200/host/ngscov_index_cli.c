/* host/ngscov_index_cli.c -- `bam_index in.bam [out.bai]`: writes the BAI index the reference tools insist on
 * (bam2depth.c:112-119 "BAM indexing file is not available") for a coordinate-sorted BAM.
 *
 * SURVEY.md section 8(f) rank 3.  The reference repository has no indexer of its own: its users run `samtools index`,
 * i.e. bam_index_build -> bam_index_core / bam_index_save of the vendored samtools-0.1.19 (bam_index.c:151-324).  This
 * is a restatement of that logic over the inflated record stream: per contig the binning index (a chunk = a maximal run
 * of consecutive records in one bin, chunks of a bin that touch inside one BGZF block merged, :113-139), the meta
 * pseudo-bin 37450 (file range and mapped / unmapped counts of the contig), the 16 kb linear index with fill_missing
 * (:141-149), and the count of reads without coordinates.  Host-only code: the BGZF blocks are inflated by one zlib
 * worker per core, a batch of blocks at a time; the sequential part only touches the 36-byte core and the CIGAR of each
 * record.  Bins are written in ascending order (samtools writes them in hash-table order; readers do not care).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>
#include <fcntl.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <zlib.h>

#define DIE(...) do { fprintf(stderr, __VA_ARGS__); exit(1); } while (0)
#define BATCH 4096                       /* BGZF blocks inflated per round (<= 256 MiB) */

typedef struct { const uint8_t *src; uint32_t clen, ulen; uint8_t *dst; uint64_t foff, uoff; } blk_t;
typedef struct { blk_t *b; size_t n; volatile size_t next; int err; } job_t;
static void *inflate_worker(void *arg)
{
	job_t *j = (job_t *)arg;
	z_stream zs; memset(&zs, 0, sizeof zs);
	if (inflateInit2(&zs, -15) != Z_OK) { j->err = 1; return NULL; }
	for (;;) {
		const size_t i = __sync_fetch_and_add(&j->next, 1);
		if (i >= j->n) break;
		blk_t *b = &j->b[i];
		if (!b->ulen) continue;
		inflateReset(&zs);
		zs.next_in = (Bytef *)b->src; zs.avail_in = b->clen; zs.next_out = b->dst; zs.avail_out = b->ulen;
		if (inflate(&zs, Z_FINISH) != Z_STREAM_END || zs.total_out != b->ulen) j->err = 1;
	}
	inflateEnd(&zs);
	return NULL;
}
static int inflate_blocks(blk_t *b, size_t n, int nthreads)
{
	job_t j; memset(&j, 0, sizeof j); j.b = b; j.n = n;
	pthread_t th[64]; int started = 0;
	if (nthreads > 64) nthreads = 64;
	for (int i = 0; i < nthreads - 1 && (size_t)i + 1 < n; ++i) if (pthread_create(&th[started], NULL, inflate_worker, &j) == 0) ++started;
	inflate_worker(&j);
	for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
	return j.err;
}

typedef struct { uint32_t bin; uint64_t beg, end; } chunk_t;
typedef struct { chunk_t *c; size_t n, cap; uint64_t *lin; size_t nlin, caplin; } refidx_t;
static int cmp_chunk(const void *a, const void *b)
{
	const chunk_t *x = (const chunk_t *)a, *y = (const chunk_t *)b;
	if (x->bin != y->bin) return x->bin < y->bin ? -1 : 1;
	return x->beg < y->beg ? -1 : x->beg > y->beg;
}
static void idx_chunk(refidx_t *r, uint32_t bin, uint64_t beg, uint64_t end)
{
	if (r->n == r->cap) { r->cap = r->cap ? r->cap * 2 : 1024; r->c = (chunk_t *)realloc(r->c, r->cap * sizeof(chunk_t)); }
	r->c[r->n].bin = bin; r->c[r->n].beg = beg; r->c[r->n++].end = end;
}

/* virtual offset of inflated position s: (file offset of its block) << 16 | offset inside the block.  A position at the
 * very end of a block belongs to the next block (bgzf_read moves on when a block is used up, samtools bgzf.c), a position
 * past the data is the end of the file.  Queried in ascending order: `cur` is a cursor. */
static uint64_t voff(const blk_t *blk, size_t nblk, size_t *cur, uint64_t s, uint64_t total, uint64_t file_end)
{
	if (s >= total || nblk == 0) return file_end << 16;
	while (*cur + 1 < nblk && blk[*cur + 1].uoff <= s) ++*cur;
	while (*cur > 0 && blk[*cur].uoff > s) --*cur;
	return (blk[*cur].foff << 16) | (s - blk[*cur].uoff);
}

int main(int argc, char **argv)
{
	if (argc < 2 || argc > 3) DIE("usage: %s in.bam [out.bai]   (writes in.bam.bai: the index bam2depth / bam2wig require)\n", argv[0]);
	const char *fn = argv[1];
	int fd = open(fn, O_RDONLY);
	struct stat sb;
	if (fd < 0 || fstat(fd, &sb) != 0 || sb.st_size < 28) DIE("bam_index: cannot open %s\n", fn);
	const size_t fsz = (size_t)sb.st_size;
	const uint8_t *file = (const uint8_t *)mmap(NULL, fsz, PROT_READ, MAP_PRIVATE, fd, 0);
	if (file == MAP_FAILED) DIE("bam_index: cannot map %s\n", fn);
	int nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
	if (getenv("NGSCOV_THREADS")) nthreads = atoi(getenv("NGSCOV_THREADS"));
	if (nthreads < 1) nthreads = 1;
	/* block table */
	size_t nblk = 0, cap = 1 << 16, off = 0; uint64_t uo = 0;
	blk_t *blk = (blk_t *)malloc(cap * sizeof(blk_t));
	while (off + 18 <= fsz) {
		const uint8_t *p = file + off;
		if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4) || p[12] != 'B' || p[13] != 'C') DIE("bam_index: %s is not a BGZF file\n", fn);
		const size_t bsize = (size_t)(p[16] | p[17] << 8) + 1;
		if (bsize < 26 || off + bsize > fsz) DIE("bam_index: truncated BGZF block in %s\n", fn);
		if (nblk == cap) { cap *= 2; blk = (blk_t *)realloc(blk, cap * sizeof(blk_t)); }
		blk[nblk].src = p + 18; blk[nblk].clen = (uint32_t)(bsize - 26); blk[nblk].foff = off; blk[nblk].uoff = uo;
		memcpy(&blk[nblk].ulen, p + bsize - 4, 4);
		if (blk[nblk].ulen > 65536) DIE("bam_index: bad BGZF block size in %s\n", fn);
		uo += blk[nblk].ulen; ++nblk; off += bsize;
	}
	const uint64_t total = uo, file_end = off;
	/* virtual offset of inflated position s: (file offset of its block) << 16 | offset inside the block; a position at
	 * the very end of a block belongs to the next block (bam_tell after a completed read), past the data: end of file */
	size_t vb = 0;                        /* cursor: positions are queried in ascending order */
#define VOFF(s) voff(blk, nblk, &vb, (s), total, file_end)
	/* rolling window of inflated bytes: [win0, win1) in buf */
	uint8_t *buf = (uint8_t *)malloc((size_t)BATCH * 65536 + (64u << 20));
	size_t next_blk = 0; uint64_t win0 = 0, win1 = 0;
	int n_ref = -1; refidx_t *R = NULL;
	uint64_t s = 0, n_no_coor = 0;
	uint32_t last_bin = 0xffffffffu, save_bin = 0xffffffffu;
	int32_t last_tid = -1, save_tid = -1; int first = 1, stopped = 0, have_hdr = 0;
	uint64_t save_off = 0, last_off = 0, off_beg = 0, n_mapped = 0, n_unmapped = 0;
	for (;;) {
		/* make sure [s, s + need) is in the window */
		uint64_t need = 4;
		if (!have_hdr) need = 12;
		for (int pass = 0; pass < 2; ++pass) {
			while (s + need > win1 && next_blk < nblk) {
				/* drop what lies before s, append a batch of blocks */
				if (s > win0) { memmove(buf, buf + (s - win0), (size_t)(win1 - s)); win0 = s; }
				const size_t b0 = next_blk; size_t fill = (size_t)(win1 - win0);
				while (next_blk < nblk && next_blk - b0 < BATCH) { blk[next_blk].dst = buf + fill; fill += blk[next_blk].ulen; ++next_blk; }
				if (fill > (size_t)BATCH * 65536 + (64u << 20)) DIE("bam_index: record larger than 64 MiB in %s\n", fn);
				if (inflate_blocks(&blk[b0], next_blk - b0, nthreads)) DIE("bam_index: inflate failed in %s\n", fn);
				win1 = win0 + fill;
			}
			if (pass == 0 && have_hdr && s + 4 <= win1) { int32_t bsz; memcpy(&bsz, buf + (s - win0), 4); if (bsz < 32) DIE("bam_index: malformed record in %s\n", fn); need = 4 + (uint64_t)bsz; }
		}
		if (!have_hdr) {
			/* BAM header (samtools bam.c:81-121): grow the window until it is complete */
			if (win1 < 12 || memcmp(buf, "BAM\1", 4)) DIE("bam_index: %s is not a BAM file\n", fn);
			uint32_t l_text; memcpy(&l_text, buf + 4, 4);
			uint64_t q = 8 + (uint64_t)l_text;
			if (q + 4 > win1) { if (next_blk >= nblk) DIE("bam_index: truncated header in %s\n", fn); need = q + 4; s = 0; goto grow; }
			int32_t nr; memcpy(&nr, buf + q, 4); q += 4;
			for (int i = 0; i < nr; ++i) {
				if (q + 4 > win1) { if (next_blk >= nblk) DIE("bam_index: truncated header in %s\n", fn); need = q + 4; goto grow; }
				uint32_t ln; memcpy(&ln, buf + q, 4);
				q += 4 + (uint64_t)ln + 4;
			}
			if (q > win1) { if (next_blk >= nblk) DIE("bam_index: truncated header in %s\n", fn); need = q; goto grow; }
			n_ref = nr; R = (refidx_t *)calloc((size_t)nr + 1, sizeof(refidx_t));
			s = q; have_hdr = 1;
			save_off = last_off = off_beg = VOFF(s);
			continue;
		grow:
			while (need > win1 && next_blk < nblk) {
				const size_t b0 = next_blk; size_t fill = (size_t)win1;
				while (next_blk < nblk && next_blk - b0 < 64) { blk[next_blk].dst = buf + fill; fill += blk[next_blk].ulen; ++next_blk; }
				if (inflate_blocks(&blk[b0], next_blk - b0, nthreads)) DIE("bam_index: inflate failed in %s\n", fn);
				win1 = fill;
			}
			if (need > win1) DIE("bam_index: truncated header in %s\n", fn);
			continue;
		}
		if (s >= total) break;
		if (s + need > win1) DIE("bam_index: truncated record in %s\n", fn);
		/* ---- one record: bam_index_core (samtools bam_index.c:151-243) */
		const uint8_t *r0 = buf + (s - win0);
		int32_t bsz, tid, pos; memcpy(&bsz, r0, 4); memcpy(&tid, r0 + 4, 4); memcpy(&pos, r0 + 8, 4);
		const uint32_t bin = r0[14] | r0[15] << 8, n_cigar = r0[16] | r0[17] << 8, flag = r0[18] | r0[19] << 8, l_qname = r0[12];
		const uint64_t next = s + 4 + (uint64_t)bsz;
		if (tid >= n_ref) DIE("bam_index: bad refID in %s\n", fn);
		if (tid < 0) ++n_no_coor;
		if (!stopped) {
			if (last_tid < tid || (last_tid >= 0 && tid < 0)) { last_tid = tid; last_bin = 0xffffffffu; }
			else if (tid >= 0 && last_tid > tid) DIE("bam_index: %s is not sorted by coordinate\n", fn);
			if (tid >= 0 && !(flag & 4)) {
				uint32_t end = (uint32_t)pos;
				for (uint32_t k = 0; k < n_cigar; ++k) { uint32_t w; memcpy(&w, r0 + 36 + l_qname + 4 * k, 4); if ((0x3C1A7 >> ((w & 0xf) << 1)) & 2) end += w >> 4; }
				/* insert_offset2 (bam_index.c:92-111), literally: int arithmetic, and n is SET to the last record's end + 1 */
				refidx_t *r = &R[tid];
				const int b = pos >> 14, e = ((int)end - 1) >> 14;
				if (e >= 0 && r->caplin < (size_t)e + 1) { const size_t oc = r->caplin; r->caplin = ((size_t)e + 1) * 2; r->lin = (uint64_t *)realloc(r->lin, r->caplin * 8); memset(r->lin + oc, 0, (r->caplin - oc) * 8); }
				for (int i = b; i <= e; ++i) if (i >= 0 && r->lin[i] == 0) r->lin[i] = last_off;
				r->nlin = e + 1 > 0 ? (size_t)(e + 1) : 0;
			}
			if (bin != last_bin) {
				if (save_bin != 0xffffffffu) idx_chunk(&R[save_tid], save_bin, save_off, last_off);
				if (last_bin == 0xffffffffu && !first) {
					idx_chunk(&R[save_tid], 37450, off_beg, last_off);
					idx_chunk(&R[save_tid], 37450, n_mapped, n_unmapped);
					n_mapped = n_unmapped = 0; off_beg = last_off;
				}
				save_off = last_off; save_bin = last_bin = bin; save_tid = tid; first = 0;
				if (save_tid < 0) stopped = 1;
			}
			if (!stopped) {
				if (flag & 4) ++n_unmapped; else ++n_mapped;
				last_off = VOFF(next);
			}
		}
		s = next;
	}
	if (!have_hdr) DIE("bam_index: %s has no BAM header\n", fn);
	if (save_tid >= 0) {
		/* samtools takes bam_tell() after the read that hit end-of-file: past the last block of the file (the EOF marker) */
		const uint64_t vend = file_end << 16;
		idx_chunk(&R[save_tid], save_bin, save_off, vend);
		idx_chunk(&R[save_tid], 37450, off_beg, vend);
		idx_chunk(&R[save_tid], 37450, n_mapped, n_unmapped);
	}
	char *ip;
	if (argc == 3) ip = strdup(argv[2]); else { ip = (char *)malloc(strlen(fn) + 8); sprintf(ip, "%s.bai", fn); }
	FILE *fi = fopen(ip, "wb");
	if (!fi) DIE("bam_index: cannot create %s\n", ip);
	fwrite("BAI\1", 1, 4, fi);
	int32_t n32 = n_ref; fwrite(&n32, 4, 1, fi);
	for (int t = 0; t < n_ref; ++t) {
		refidx_t *r = &R[t];
		chunk_t meta[2]; int nmeta = 0; size_t w = 0;
		for (size_t i = 0; i < r->n; ++i) { if (r->c[i].bin == 37450) { if (nmeta < 2) meta[nmeta++] = r->c[i]; } else r->c[w++] = r->c[i]; }
		qsort(r->c, w, sizeof(chunk_t), cmp_chunk);
		size_t m = 0;                                 /* merge_chunks (bam_index.c:113-139) */
		for (size_t i = 0; i < w; ++i) {
			if (m && r->c[m - 1].bin == r->c[i].bin && (r->c[m - 1].end >> 16) == (r->c[i].beg >> 16)) r->c[m - 1].end = r->c[i].end;
			else r->c[m++] = r->c[i];
		}
		int32_t nbin = 0;
		for (size_t i = 0; i < m; ++i) if (i == 0 || r->c[i].bin != r->c[i - 1].bin) ++nbin;
		if (nmeta) ++nbin;
		fwrite(&nbin, 4, 1, fi);
		for (size_t i = 0; i < m; ) {
			size_t j = i; while (j < m && r->c[j].bin == r->c[i].bin) ++j;
			const int32_t nch = (int32_t)(j - i);
			fwrite(&r->c[i].bin, 4, 1, fi); fwrite(&nch, 4, 1, fi);
			for (size_t k = i; k < j; ++k) { fwrite(&r->c[k].beg, 8, 1, fi); fwrite(&r->c[k].end, 8, 1, fi); }
			i = j;
		}
		if (nmeta) {
			const uint32_t mb = 37450; const int32_t two = nmeta;
			fwrite(&mb, 4, 1, fi); fwrite(&two, 4, 1, fi);
			for (int k = 0; k < nmeta; ++k) { fwrite(&meta[k].beg, 8, 1, fi); fwrite(&meta[k].end, 8, 1, fi); }
		}
		for (size_t i = 1; i < r->nlin; ++i) if (r->lin[i] == 0) r->lin[i] = r->lin[i - 1];   /* fill_missing (bam_index.c:141-149) */
		const int32_t nl = (int32_t)r->nlin; fwrite(&nl, 4, 1, fi);
		if (nl) fwrite(r->lin, 8, (size_t)nl, fi);
		free(r->c); free(r->lin);
	}
	fwrite(&n_no_coor, 8, 1, fi);
	fclose(fi);
	free(ip); free(R); free(buf); free(blk);
	munmap((void *)file, fsz); close(fd);
	return 0;
}
