/* host/ngscov_cli.c -- the drop-in bam2depth / bam2wig command lines on top of libngscov_cuda.
 *
 * Compiled twice: plain -> bam2depth, with -DNGSCOV_TOOL_WIG -> bam2wig.  Behaviour kept from the
 * reference mains (reference bam2depth.c:257-355, bam2wig.c:254-336): the getopt strings
 * ("o:w:r:s:Wh?" / "o:w:r:s:h?"), defaults (-o "-", -w 20000), atoi parsing, -r/-s parsed and never
 * used, usage on stderr + exit(1), output names <basename(bam)>.N.bedGraph (cwd) / <PFX>.N.depth /
 * <PFX>.N.wig / <PFX>.N.chromSize.txt created with O_TRUNC before the index is looked up, names that
 * begin with '-' meaning stdout (IO_stream.h:69-97), the hard requirement of a .bai next to the BAM
 * (bam2depth.c:112-119), and the "<chr> at %.3f s" / "Converted %s to wig format at %.3f s" progress
 * lines on stderr.
 *
 * What is different is everything in between: the host only inflates BGZF blocks (zlib, one worker per
 * core, straight into the library's pinned staging buffers) and hands the inflated record stream to
 * the GPU engine; record parsing, depth accumulation, run detection, window bins and text formatting
 * all happen on the device.  I/O (read + inflate) and engine time are reported separately when
 * NGSCOV_VERBOSE is set.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>
#include <fcntl.h>
#include <getopt.h>
#include <libgen.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <zlib.h>
#include "../../include/ngscov.h"

#ifdef NGSCOV_TOOL_WIG
#define TOOL_DEPTH 0
#define OPTSTRING "o:w:r:s:h?"
#else
#define TOOL_DEPTH 1
#define OPTSTRING "o:w:r:s:Wh?"
#endif

static long long usec(void) { struct timeval tv; gettimeofday(&tv, NULL); return (long long)tv.tv_sec * 1000000 + tv.tv_usec; }

static void usage(const char *prog)
{
	fprintf(stderr,
	        "\nUsage: %s [-o OUTFILE] [-w WINDOW_SIZE] [-r chr1:1-2000000]%s [-s 0] [-h] bamFile1 bamFile2 ..\n"
	        "  B200-native drop-in for the HighPerformanceNGS %s tool.\n"
	        "   [-o OUTPUT_FILE]  = output prefix.                                  [required]\n"
	        "   [-w WINDOW_SIZE]  = window size, default 20000.                     [option]\n"
	        "%s"
	        "   [-r]              = accepted and ignored, like the reference.       [option]\n"
	        "   [-s]              = accepted and ignored, like the reference.       [option]\n"
	        "   [-h]              = this help.                                      [option]\n"
	        "   Infiles           = coordinate-sorted, indexed BAM file(s).         [required]\n\n",
	        prog, TOOL_DEPTH ? " [-W is wig]" : "", TOOL_DEPTH ? "bam2depth" : "bam2wig",
	        TOOL_DEPTH ? "   [-W is wig]       = also write wig + chromSize, default not.        [option]\n" : "");
	exit(1);
}

/* fopen_output_stream (IO_stream.h:69-83): "-"-prefixed or empty names go to stdout */
static FILE *open_out(const char *prefix, const char *suffix)
{
	char *name = (char *)calloc(strlen(prefix) + strlen(suffix) + 2, 1);
	sprintf(name, "%s%s", prefix, suffix);
	int fd;
	if (strncmp(name, "-", 1) == 0 || !strcmp(name, "")) fd = STDOUT_FILENO;
	else {
		fd = open(name, O_CREAT | O_WRONLY | O_TRUNC, 0666);
		if (fd == -1) { fprintf(stderr, "Failed to create output file (%s)\n", name); exit(1); }   /* the reference carries on and crashes later */
	}
	free(name);
	FILE *f = fdopen(fd, "wb");
	if (!f) { fprintf(stderr, "Failed to open output stream\n"); exit(1); }
	return f;
}

static int have_bai(const char *fn)
{
	size_t l = strlen(fn);
	char *a = (char *)malloc(l + 8);
	sprintf(a, "%s.bai", fn);
	int ok = access(a, R_OK) == 0;
	if (!ok && l > 4 && !strcmp(fn + l - 4, ".bam")) { strcpy(a, fn); strcpy(a + l - 3, "bai"); ok = access(a, R_OK) == 0; }
	free(a);
	return ok;
}

/* ------------------------------------------------------------------ parallel BGZF inflate */
typedef struct { const uint8_t *src; uint32_t clen, ulen; uint8_t *dst; uint64_t foff; } blk_t;
typedef struct { blk_t *b; size_t n; volatile size_t next; int err; pthread_mutex_t mu; } job_t;

static void *inflate_worker(void *arg)
{
	job_t *j = (job_t *)arg;
	z_stream zs;
	memset(&zs, 0, sizeof zs);
	if (inflateInit2(&zs, -15) != Z_OK) { j->err = 1; return NULL; }
	for (;;) {
		size_t i = __sync_fetch_and_add(&j->next, 1);
		if (i >= j->n) break;
		blk_t *b = &j->b[i];
		if (b->ulen == 0) continue;
		inflateReset(&zs);
		zs.next_in = (Bytef *)b->src; zs.avail_in = b->clen;
		zs.next_out = b->dst; zs.avail_out = b->ulen;
		if (inflate(&zs, Z_FINISH) != Z_STREAM_END || zs.total_out != b->ulen) j->err = 1;
	}
	inflateEnd(&zs);
	return NULL;
}

static int inflate_blocks(blk_t *b, size_t n, int nthreads)
{
	job_t j; memset(&j, 0, sizeof j);
	j.b = b; j.n = n;
	if (nthreads > 64) nthreads = 64;
	pthread_t th[64];
	int started = 0;
	for (int i = 0; i < nthreads - 1 && (size_t)i + 1 < n; ++i) if (pthread_create(&th[started], NULL, inflate_worker, &j) == 0) ++started;
	inflate_worker(&j);
	for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
	return j.err;
}

#define DIE(...) do { fprintf(stderr, __VA_ARGS__); exit(1); } while (0)

/* ------------------------------------------------------------------ output
 * The reference fprintf()s one line at a time; here a finished text sits in pinned host memory.  Regular files are
 * written by several threads with pwrite() on disjoint ranges (a single fwrite of the 25 GB bedGraph of a 30x genome
 * takes several times longer than the whole GPU job); pipes and terminals get a plain fwrite. */
typedef struct { int fd; const char *buf; size_t off, len; int err; } wjob_t;
static void *write_worker(void *arg)
{
	wjob_t *w = (wjob_t *)arg;
	size_t done = 0;
	while (done < w->len) {
		const ssize_t k = pwrite(w->fd, w->buf + done, w->len - done > (64u << 20) ? (64u << 20) : w->len - done, (off_t)(w->off + done));
		if (k <= 0) { w->err = 1; break; }
		done += (size_t)k;
	}
	return NULL;
}
static int write_out(FILE *f, const char *buf, size_t len, size_t *pos, int nthreads)
{
	if (!len) return 0;
	struct stat sb;
	const int fd = fileno(f);
	if (len < (32u << 20) || fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) { *pos += len; return fwrite(buf, 1, len, f) == len ? 0 : 1; }
	fflush(f);
	if (nthreads > 16) nthreads = 16;
	if (nthreads < 1) nthreads = 1;
	pthread_t th[16]; wjob_t job[16];
	const size_t per = (len + (size_t)nthreads - 1) / (size_t)nthreads;
	int n = 0, err = 0;
	for (size_t o = 0; o < len; o += per, ++n) {
		job[n].fd = fd; job[n].buf = buf + o; job[n].off = *pos + o; job[n].len = len - o < per ? len - o : per; job[n].err = 0;
		if (pthread_create(&th[n], NULL, write_worker, &job[n])) { write_worker(&job[n]); th[n] = 0; }
	}
	for (int i = 0; i < n; ++i) { if (th[i]) pthread_join(th[i], NULL); err |= job[i].err; }
	*pos += len;
	if (fseeko(f, (off_t)*pos, SEEK_SET)) err = 1;
	return err;
}
/* "<contig> at <t> s" (reference bam2depth.c:338): one line per contig, printed when the engine reports it finalized */
static void report_contigs(ngscov_t *h, char **ref_name, int n_ref, int *printed, long long begin, int all)
{
	uint32_t done = 0;
	if (all) done = (uint32_t)n_ref; else if (ngscov_contigs_done(h, &done)) return;
	for (; *printed < (int)done && *printed < n_ref; ++*printed) fprintf(stderr, "%s at %.3f s\n", ref_name[*printed], (double)(usec() - begin) / 1e6);
}

/* ------------------------------------------------------------------ region shards (NGSCOV_GPUS=N)
 * No counterpart in the single-process reference.  The genome is cut at multiples of 16 384 bp (the BAI linear
 * index granularity, samtools bam_index.c:43); shard g reads the file from the linear-index offset of its first
 * tile -- the first record OVERLAPPING that tile (bam_index.c:92-111,605-615), so reads hanging in from the left
 * are included -- up to the offset of a tile a little behind its end, and then keeps reading until the engine has
 * seen a record that starts behind the slice (ngscov_slice_done; coordinate-sorted input).  Each shard owns an
 * engine handle on its own GPU; deltas are clipped to the slice on the device.  Afterwards the runs cut by shard
 * boundaries are stitched (halo exchange through the host: 72 bytes per shard), window bins are corrected for the
 * merged runs and summed, and the bedGraph texts are concatenated in shard order. */
#include "ngscov_shard.h"

static size_t block_of(const blk_t *blk, size_t nblk, uint64_t coff)
{
	size_t lo = 0, hi = nblk;                  /* first block with foff >= coff */
	while (lo < hi) { size_t mid = (lo + hi) / 2; if (blk[mid].foff < coff) lo = mid + 1; else hi = mid; }
	return lo;
}

typedef struct {
	/* in */
	ngscov_cfg cfg; blk_t *blk; size_t nblk; size_t b_first, b_bound; uint32_t uoff; int to_eof, empty; int nthreads; const char *fn;
	int gpu_inflate; const uint8_t *file_end;
	/* out */
	ngscov_t *h; int rc; char err[256]; long long t_io;
} shard_t;

static void *shard_worker(void *arg)
{
	shard_t *s = (shard_t *)arg;
	int rc;
	if ((rc = ngscov_create(&s->h, &s->cfg))) { snprintf(s->err, sizeof s->err, "GPU engine unavailable on device %d: %s", s->cfg.device, ngscov_strerror(rc)); s->rc = rc; return NULL; }
	void *stage[2]; size_t scap[2];
	for (int i = 0; i < 2; ++i) if ((rc = ngscov_staging(s->h, i, &stage[i], &scap[i]))) { snprintf(s->err, sizeof s->err, "%s", ngscov_strerror(rc)); s->rc = rc; return NULL; }
	/* blocks are inflated by a private copy of the table entries (dst differs per shard; shards may overlap at their ends) */
	size_t b = s->b_first, limit = s->to_eof ? s->nblk : (s->b_bound + 1 < s->nblk ? s->b_bound + 1 : s->nblk);
	int idx = 0, first = 1, last = s->empty;
	if (limit < b) limit = b;            /* the first record at or behind the slice start already lies behind its end bound */
	if (s->gpu_inflate && !s->empty) {
		/* compressed blocks [b, limit) go to the device as they lie in the file; extend while the slice is not done */
		size_t skip = s->uoff;
		while (!last) {
			const uint8_t *p0 = s->blk[b].src - 18;
			const uint8_t *p1 = limit < s->nblk ? s->blk[limit].src - 18 : s->file_end;
			const int eof = limit >= s->nblk;
			const size_t nb = (size_t)(p1 - p0);
			if ((rc = ngscov_push_bgzf(s->h, p0, nb, nb ? skip : 0, eof ? NGSCOV_PUSH_LAST : 0))) break;
			if (nb) skip = 0;
			b = limit;
			if (eof) { last = 1; break; }
			int done = 0;
			if ((rc = ngscov_slice_done(s->h, &done))) break;
			if (done) { rc = ngscov_push(s->h, 0, 0, NGSCOV_PUSH_LAST | NGSCOV_PUSH_TRUNCATED_OK); last = 1; }
			else limit = limit + 64 < s->nblk ? limit + 64 : s->nblk;
		}
		if (!rc) rc = ngscov_finish_runs(s->h);
		if (rc && !s->err[0]) snprintf(s->err, sizeof s->err, "%s: %s (%s)", s->fn, ngscov_strerror(rc), ngscov_last_cuda_error(s->h));
		s->rc = rc;
		return NULL;
	}
	if (s->empty) { rc = ngscov_push(s->h, 0, 0, NGSCOV_PUSH_LAST); if (rc) { snprintf(s->err, sizeof s->err, "%s", ngscov_strerror(rc)); s->rc = rc; return NULL; } }
	uint8_t *tmp = (uint8_t *)malloc(65536);        /* (freed on every path below: the loop only leaves through `break`) */
	blk_t *mine = (blk_t *)malloc(4096 * sizeof(blk_t));
	while (!last) {
		const long long t0 = usec();
		if ((rc = ngscov_staging_wait(s->h, idx))) break;
		uint8_t *dst = (uint8_t *)stage[idx];
		size_t fill = 0;
		if (first && b < limit) {                     /* the slice starts uoff bytes into its first block */
			blk_t one = s->blk[b]; one.dst = tmp;
			if (inflate_blocks(&one, 1, 1) || s->uoff > one.ulen) { snprintf(s->err, sizeof s->err, "inflate failed in %s", s->fn); rc = NGSCOV_EFORMAT; break; }
			memcpy(dst, tmp + s->uoff, one.ulen - s->uoff); fill = one.ulen - s->uoff; ++b;
		}
		first = 0;
		int bad = 0;
		while (b < limit && fill + s->blk[b].ulen <= scap[idx]) {
			size_t m = 0;
			while (m < 4096 && b < limit && fill + s->blk[b].ulen <= scap[idx]) { mine[m] = s->blk[b]; mine[m].dst = dst + fill; fill += s->blk[b].ulen; ++m; ++b; }
			if (m && inflate_blocks(mine, m, s->nthreads)) { bad = 1; break; }
		}
		if (bad) { snprintf(s->err, sizeof s->err, "inflate failed in %s", s->fn); rc = NGSCOV_EFORMAT; break; }
		s->t_io += usec() - t0;
		const int at_limit = b >= limit;
		if (at_limit && limit >= s->nblk) {            /* end of file: records are complete */
			rc = ngscov_push(s->h, idx, fill, NGSCOV_PUSH_LAST); last = 1;
		} else {
			if ((rc = ngscov_push(s->h, idx, fill, 0))) break;
			if (at_limit) {
				int done = 0;
				if ((rc = ngscov_slice_done(s->h, &done))) break;
				if (done) { rc = ngscov_push(s->h, idx ^ 1, 0, NGSCOV_PUSH_LAST | NGSCOV_PUSH_TRUNCATED_OK); last = 1; }
				else limit = limit + 64 < s->nblk ? limit + 64 : s->nblk;      /* a long read still spans the cut: read on */
			}
		}
		if (rc) break;
		idx ^= 1;
	}
	free(tmp); free(mine);
	if (!rc) rc = ngscov_finish_runs(s->h);
	if (rc && !s->err[0]) snprintf(s->err, sizeof s->err, "%s: %s (%s)", s->fn, ngscov_strerror(rc), ngscov_last_cuda_error(s->h));
	s->rc = rc;
	return NULL;
}


int main(int argc, char *argv[])
{
	const char *outfile = "-"; uint32_t window = 20000; int wig = 0, opt;
	if (argc < 2) usage(argv[0]);
	while ((opt = getopt(argc, argv, OPTSTRING)) != -1) {
		switch (opt) {
		case 'o': outfile = optarg; break;
		case 'w': window = (uint32_t)atoi(optarg); break;
		case 'W': wig++; break;
		case 'r': case 's': break;                 /* parsed, never used (reference bam2depth.c:281-285) */
		case '?': case 'h': usage(argv[0]); break;
		default: fprintf(stderr, "error parameter!\n"); break;
		}
	}
	const long long begin = usec();
	const int verbose = getenv("NGSCOV_VERBOSE") != NULL;
	int nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
	if (getenv("NGSCOV_THREADS")) nthreads = atoi(getenv("NGSCOV_THREADS"));
	if (nthreads < 1) nthreads = 1;
	const int device = getenv("NGSCOV_DEVICE") ? atoi(getenv("NGSCOV_DEVICE")) : 0;

	for (int fi = optind; fi < argc; ++fi) {
		const char *fn = argv[fi];
		const int n = fi - optind + 1;
		int fd = open(fn, O_RDONLY);
		if (fd < 0) DIE("%s: bam2bed: Fail to open BAM file %s\n", basename(argv[0]), fn);
		char suffix[64];
		FILE *f_bed = NULL, *f_depth = NULL, *f_wig = NULL, *f_chr = NULL;
		if (TOOL_DEPTH) {
			char *tmp = strdup(fn);
			sprintf(suffix, ".%u.bedGraph", (unsigned)n); f_bed = open_out(basename(tmp), suffix);
			free(tmp);
			sprintf(suffix, ".%u.depth", (unsigned)n); f_depth = open_out(outfile, suffix);
		}
		if (wig || !TOOL_DEPTH) {
			sprintf(suffix, ".%u.wig", (unsigned)n); f_wig = open_out(outfile, suffix);
			sprintf(suffix, ".%u.chromSize.txt", (unsigned)n); f_chr = open_out(outfile, suffix);
		}
		/* the reference refuses to run without a .bai (bam2depth.c:112-119) although only bam_fetch needs it; the engine
		 * walks the whole stream, so NGSCOV_NO_INDEX=1 lifts the requirement for single-GPU runs (region shards still take
		 * their file offsets from the linear index) */
		const int no_index = getenv("NGSCOV_NO_INDEX") && atoi(getenv("NGSCOV_NO_INDEX")) != 0;
		if (!have_bai(fn) && !(no_index && !(getenv("NGSCOV_GPUS") && atoi(getenv("NGSCOV_GPUS")) > 1))) DIE("bam2bed: BAM indexing file is not available.\n");

		struct stat sb;
		if (fstat(fd, &sb) != 0 || sb.st_size < 28) DIE("%s: bam2bed: Fail to open BAM file %s\n", basename(argv[0]), fn);
		const size_t fsz = (size_t)sb.st_size;
		const uint8_t *file = (const uint8_t *)mmap(NULL, fsz, PROT_READ, MAP_PRIVATE, fd, 0);
		if (file == MAP_FAILED) DIE("bam2bed: cannot map %s\n", fn);
		madvise((void *)file, fsz, MADV_SEQUENTIAL);

		/* block table: offsets, compressed and inflated sizes (18-byte header with BSIZE, 8-byte footer with ISIZE) */
		size_t nblk = 0, cap = 1 << 16, off = 0;
		blk_t *blk = (blk_t *)malloc(cap * sizeof(blk_t));
		while (off + 18 <= fsz) {
			const uint8_t *p = file + off;
			if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4) || p[12] != 'B' || p[13] != 'C') DIE("bam2bed: %s is not a BGZF file\n", fn);
			const size_t bsize = (size_t)(p[16] | p[17] << 8) + 1;
			if (bsize < 26 || off + bsize > fsz) DIE("bam2bed: truncated BGZF block in %s\n", fn);
			if (nblk == cap) { cap *= 2; blk = (blk_t *)realloc(blk, cap * sizeof(blk_t)); }
			blk[nblk].src = p + 18; blk[nblk].clen = (uint32_t)(bsize - 26); blk[nblk].foff = off;
			memcpy(&blk[nblk].ulen, p + bsize - 4, 4);
			if (blk[nblk].ulen > 65536) DIE("bam2bed: bad BGZF block size in %s\n", fn);
			++nblk; off += bsize;
		}

		/* header: inflate leading blocks until it is complete (samtools bam.c:81-121 layout) */
		size_t hb = 0, hbytes = 0, hcap = 1 << 20, hdr_len = 0;
		uint8_t *hbuf = (uint8_t *)malloc(hcap);
		int n_ref = -1; uint32_t *ref_len = NULL; char **ref_name = NULL;
		while (n_ref < 0) {
			if (hb >= nblk) DIE("bam2bed: %s has no BAM header\n", fn);
			if (hbytes + 65536 > hcap) { hcap *= 2; hbuf = (uint8_t *)realloc(hbuf, hcap); }
			blk[hb].dst = hbuf + hbytes;
			if (inflate_blocks(&blk[hb], 1, 1)) DIE("bam2bed: inflate failed in %s\n", fn);
			hbytes += blk[hb].ulen; ++hb;
			if (hbytes < 12) continue;
			if (memcmp(hbuf, "BAM\1", 4)) DIE("bam2bed: %s is not a BAM file\n", fn);
			uint32_t l_text; memcpy(&l_text, hbuf + 4, 4);
			size_t p = 8 + (size_t)l_text;
			if (p + 4 > hbytes) continue;
			int32_t nr; memcpy(&nr, hbuf + p, 4); p += 4;
			int complete = 1;
			size_t q = p;
			for (int i = 0; i < nr; ++i) {
				if (q + 4 > hbytes) { complete = 0; break; }
				uint32_t ln; memcpy(&ln, hbuf + q, 4);
				if (q + 4 + ln + 4 > hbytes) { complete = 0; break; }
				q += 4 + ln + 4;
			}
			if (!complete) continue;
			n_ref = nr; hdr_len = q;
			ref_len = (uint32_t *)calloc((size_t)nr + 1, 4); ref_name = (char **)calloc((size_t)nr + 1, sizeof(char *));
			for (int i = 0; i < nr; ++i) {
				uint32_t ln; memcpy(&ln, hbuf + p, 4); p += 4;
				ref_name[i] = strndup((const char *)hbuf + p, ln); p += ln;
				memcpy(&ref_len[i], hbuf + p, 4); p += 4;
			}
		}

		ngscov_cfg cfg; memset(&cfg, 0, sizeof cfg);
		cfg.mode = TOOL_DEPTH ? NGSCOV_MODE_DEPTH : NGSCOV_MODE_WIG;
		cfg.window = window; cfg.emit_wig = wig ? 1 : 0; cfg.n_ref = n_ref; cfg.ref_len = ref_len;
		cfg.ref_name = (const char *const *)ref_name; cfg.device = device;
		{	/* two pinned staging buffers: large enough to keep PCIe busy, no larger than the file needs */
			uint64_t total = 0;
			for (size_t i = 0; i < nblk; ++i) total += blk[i].ulen;
			uint64_t want = total / 2 + (1u << 20);
			if (want > (256ull << 20)) want = 256ull << 20;
			cfg.staging_bytes = want;
			if (getenv("NGSCOV_STAGING_MB")) cfg.staging_bytes = (uint64_t)atoll(getenv("NGSCOV_STAGING_MB")) << 20;
		}
		int n_shards = getenv("NGSCOV_GPUS") ? atoi(getenv("NGSCOV_GPUS")) : 1;
		if (n_shards > 64) n_shards = 64;
		if (n_shards > 1) {
			/* ---- region-sharded run: one engine per GPU (NGSCOV_DEVICES="0,1,.." overrides the default 0..N-1) */
			int devs[64];
			for (int g = 0; g < n_shards; ++g) devs[g] = g;
			if (getenv("NGSCOV_DEVICES")) { const char *q = getenv("NGSCOV_DEVICES"); for (int g = 0; g < n_shards && *q; ++g) { devs[g] = atoi(q); while (*q && *q != ',') ++q; if (*q) ++q; } }
			bai_t bai; memset(&bai, 0, sizeof bai);
			if (bai_load(fn, &bai) || bai.n_ref != n_ref) DIE("bam2bed: BAM indexing file is not available.\n");
			slice_t sl[64]; plan_slices(ref_len, n_ref, n_shards, sl);
			shard_t sh[64]; memset(sh, 0, sizeof sh);
			pthread_t th[64];
			int ns = 0;
			for (int g = 0; g < n_shards; ++g) {
				if (sl[g].end <= sl[g].beg) continue;          /* more shards than 16 kb tiles */
				shard_t *s = &sh[ns];
				s->cfg = cfg; s->cfg.device = devs[g]; s->cfg.slice_beg = sl[g].beg; s->cfg.slice_end = sl[g].end;
				s->blk = blk; s->nblk = nblk; s->fn = fn; s->gpu_inflate = getenv("NGSCOV_GPU_INFLATE") && atoi(getenv("NGSCOV_GPU_INFLATE")) != 0; s->file_end = file + off; s->nthreads = nthreads / n_shards > 0 ? nthreads / n_shards : 1;
				int cb, ce; uint32_t pb, pe;
				locate(ref_len, n_ref, sl[g].beg, &cb, &pb); locate(ref_len, n_ref, sl[g].end, &ce, &pe);
				const uint64_t v0 = bai_seek(&bai, cb, pb);
				if (v0 == 0) s->empty = 1;                     /* no mapped record from here on */
				else { s->b_first = block_of(blk, nblk, v0 >> 16); s->uoff = (uint32_t)(v0 & 0xffff); if (s->b_first >= nblk) s->empty = 1; }
				if (ce >= n_ref) s->to_eof = 1;                /* the last shard also walks the unmapped tail */
				else {
					const uint64_t v1 = bai_seek(&bai, ce, pe + NGS_END_MARGIN_TILES * NGS_TILE_BP);
					if (v1 == 0) s->to_eof = 1; else s->b_bound = block_of(blk, nblk, v1 >> 16);
				}
				{	uint64_t want = (64u << 20); if (getenv("NGSCOV_STAGING_MB")) want = (uint64_t)atoll(getenv("NGSCOV_STAGING_MB")) << 20; s->cfg.staging_bytes = want; }
				++ns;
			}
			if (ns == 0) DIE("bam2bed: %s: nothing to shard (empty genome)\n", fn);
			for (int g = 0; g < ns; ++g) if (pthread_create(&th[g], NULL, shard_worker, &sh[g])) DIE("bam2bed: cannot start shard thread\n");
			for (int g = 0; g < ns; ++g) pthread_join(th[g], NULL);
			for (int g = 0; g < ns; ++g) if (sh[g].rc) DIE("bam2bed: shard %d: %s\n", g, sh[g].err);
			/* halo stitch + window bins */
			ngscov_halo halo[64]; int drop[64]; uint32_t nle[64]; corr_t corr[3 * 64];
			int rc2;
			for (int g = 0; g < ns; ++g) if ((rc2 = ngscov_halo_get(sh[g].h, &halo[g]))) DIE("bam2bed: %s\n", ngscov_strerror(rc2));
			const size_t ncorr = stitch(halo, ns, drop, nle, corr);
			for (int g = 0; g < ns; ++g) if ((rc2 = ngscov_halo_put(sh[g].h, drop[g], nle[g]))) DIE("bam2bed: %s\n", ngscov_strerror(rc2));
			uint64_t nb = 0; ngscov_bins_count(sh[0].h, &nb);
			uint64_t *sum = (uint64_t *)calloc(nb + 1, 8), *one = (uint64_t *)calloc(nb + 1, 8);
			for (int g = 0; g < ns; ++g) {
				if ((rc2 = ngscov_bins_get(sh[g].h, one, nb))) DIE("bam2bed: %s\n", ngscov_strerror(rc2));
				for (uint64_t k = 0; k < nb; ++k) sum[k] += one[k];
			}
			for (size_t k = 0; k < ncorr; ++k) if ((rc2 = ngscov_bins_apply_run(sh[0].h, &corr[k].run, corr[k].sign, sum, nb))) DIE("bam2bed: %s\n", ngscov_strerror(rc2));
			if ((rc2 = ngscov_bins_put(sh[0].h, sum, nb))) DIE("bam2bed: %s\n", ngscov_strerror(rc2));
			free(sum); free(one);
			const uint32_t all = (TOOL_DEPTH ? (1u << NGSCOV_OUT_BEDGRAPH) | (1u << NGSCOV_OUT_DEPTH) : 0u) |
			                     ((wig || !TOOL_DEPTH) ? (1u << NGSCOV_OUT_WIG) | (1u << NGSCOV_OUT_CHROMSIZE) : 0u);
			for (int g = 0; g < ns; ++g) {
				const uint32_t mask = g == 0 ? all : (TOOL_DEPTH ? (1u << NGSCOV_OUT_BEDGRAPH) : 0u);
				if (mask && (rc2 = ngscov_format(sh[g].h, mask))) DIE("bam2bed: %s: %s\n", fn, ngscov_strerror(rc2));
			}
			const int which2[4] = { NGSCOV_OUT_BEDGRAPH, NGSCOV_OUT_DEPTH, NGSCOV_OUT_WIG, NGSCOV_OUT_CHROMSIZE };
			FILE *outs2[4] = { f_bed, f_depth, f_wig, f_chr };
			for (int k = 0; k < 4; ++k) {
				if (!outs2[k]) continue;
				size_t pos = 0;
				for (int g = 0; g < (k == 0 ? ns : 1); ++g) {
					const char *buf; size_t len;
					if ((rc2 = ngscov_output(sh[g].h, which2[k], &buf, &len))) DIE("bam2bed: %s\n", ngscov_strerror(rc2));
					if (write_out(outs2[k], buf, len, &pos, nthreads)) DIE("bam2bed: write failed\n");
				}
			}
			for (int j = 0; j < n_ref; ++j) fprintf(stderr, "%s at %.3f s\n", ref_name[j], (double)(usec() - begin) / 1e6);
			if (verbose) {
				for (int g = 0; g < ns; ++g) {
					ngscov_stats_t st; ngscov_stats(sh[g].h, &st);
					fprintf(stderr, "ngscov: shard %d (device %d, slice [%llu,%llu)): %llu records, %llu runs; host read+inflate %.3f s (%d threads); "
					        "device: ingest %.3f ms scan %.3f ms format %.3f ms\n", g, sh[g].cfg.device, (unsigned long long)sh[g].cfg.slice_beg,
					        (unsigned long long)sh[g].cfg.slice_end, (unsigned long long)st.records, (unsigned long long)st.runs, sh[g].t_io / 1e6,
					        sh[g].nthreads, st.ms_ingest, st.ms_scan, st.ms_format);
				}
			}
			for (int g = 0; g < ns; ++g) ngscov_destroy(sh[g].h);
			for (int k = 0; k < 4; ++k) if (outs2[k]) fclose(outs2[k]);
			for (int j = 0; j < n_ref; ++j) { free(ref_name[j]); free(bai.lin[j]); }
			free(bai.lin); free(bai.n_lin); free(bai.first);
			free(ref_name); free(ref_len); free(hbuf); free(blk);
			munmap((void *)file, fsz); close(fd);
			fprintf(stderr, "Converted %s to wig format at %.3f s\n", fn, (double)(usec() - begin) / 1e6);
			continue;
		}
		ngscov_t *h = NULL;
		int rc = ngscov_create(&h, &cfg);
		if (rc) DIE("bam2bed: GPU engine unavailable: %s\n", ngscov_strerror(rc));
		void *stage[2]; size_t scap[2];
		for (int i = 0; i < 2; ++i) if ((rc = ngscov_staging(h, i, &stage[i], &scap[i]))) DIE("bam2bed: %s\n", ngscov_strerror(rc));

		const int gpu_inflate = getenv("NGSCOV_GPU_INFLATE") && atoi(getenv("NGSCOV_GPU_INFLATE")) != 0;
		long long t_io = 0;
		int printed = 0;
		if (gpu_inflate) {
			/* the file goes to the device compressed, a few hundred blocks at a time so that finished contigs can be reported;
			 * the inflated BAM header (hdr_len bytes) is skipped on the device */
			size_t b0 = 0, skip = hdr_len;
			const size_t step = 4096;                              /* blocks per call: ~64 MB of BAM */
			while (b0 < nblk) {
				const size_t b1 = b0 + step < nblk ? b0 + step : nblk;
				const uint8_t *p0 = blk[b0].src - 18, *p1 = b1 < nblk ? blk[b1].src - 18 : file + off;
				if ((rc = ngscov_push_bgzf(h, p0, (size_t)(p1 - p0), skip, b1 >= nblk ? NGSCOV_PUSH_LAST : 0))) DIE("bam2bed: %s: %s (%s)\n", fn, ngscov_strerror(rc), ngscov_last_cuda_error(h));
				skip = 0; b0 = b1;
				report_contigs(h, ref_name, n_ref, &printed, begin, 0);
			}
			if (nblk == 0 && (rc = ngscov_push_bgzf(h, file, 0, 0, NGSCOV_PUSH_LAST))) DIE("bam2bed: %s: %s\n", fn, ngscov_strerror(rc));
		}
		/* stream: [tail of the header blocks] then the remaining blocks, inflated in parallel into staging */
		int idx = 0; size_t b = hb;
		size_t pending_hdr_tail = hbytes - hdr_len;            /* record bytes that came with the header blocks */
		int first = 1, pushed_last = gpu_inflate;
		while (!pushed_last) {
			const long long t0 = usec();
			if ((rc = ngscov_staging_wait(h, idx))) DIE("bam2bed: %s\n", ngscov_strerror(rc));
			uint8_t *dst = (uint8_t *)stage[idx];
			size_t fill = 0;
			if (first && pending_hdr_tail) {
				if (pending_hdr_tail > scap[idx]) DIE("bam2bed: staging buffer too small\n");
				memcpy(dst, hbuf + hdr_len, pending_hdr_tail); fill = pending_hdr_tail;
			}
			first = 0;
			const size_t b0 = b;
			while (b < nblk && fill + blk[b].ulen <= scap[idx]) { blk[b].dst = dst + fill; fill += blk[b].ulen; ++b; }
			if (b > b0 && inflate_blocks(&blk[b0], b - b0, nthreads)) DIE("bam2bed: inflate failed in %s\n", fn);
			t_io += usec() - t0;
			pushed_last = b >= nblk;
			if ((rc = ngscov_push(h, idx, fill, pushed_last ? NGSCOV_PUSH_LAST : 0))) DIE("bam2bed: %s: %s\n", fn, ngscov_strerror(rc));
			report_contigs(h, ref_name, n_ref, &printed, begin, 0);
			idx ^= 1;
		}
		if ((rc = ngscov_finish(h))) DIE("bam2bed: %s: %s (%s)\n", fn, ngscov_strerror(rc), ngscov_last_cuda_error(h));

		const int which[4] = { NGSCOV_OUT_BEDGRAPH, NGSCOV_OUT_DEPTH, NGSCOV_OUT_WIG, NGSCOV_OUT_CHROMSIZE };
		FILE *outs[4] = { f_bed, f_depth, f_wig, f_chr };
		report_contigs(h, ref_name, n_ref, &printed, begin, 1);
		for (int k = 0; k < 4; ++k) {
			if (!outs[k]) continue;
			const char *buf; size_t len, pos = 0;
			if ((rc = ngscov_output(h, which[k], &buf, &len))) DIE("bam2bed: %s\n", ngscov_strerror(rc));
			if (write_out(outs[k], buf, len, &pos, nthreads)) DIE("bam2bed: write failed\n");
		}
		if (verbose) {
			ngscov_stats_t st; ngscov_stats(h, &st);
			fprintf(stderr, "ngscov: %llu records (%llu kept), %llu runs; host read+inflate %.3f s (%d threads); device: h2d %.3f ms ingest %.3f ms "
			        "scan %.3f ms format %.3f ms d2h %.3f ms, device inflate %.3f ms, %u kernel launches\n", (unsigned long long)st.records, (unsigned long long)st.kept,
			        (unsigned long long)st.runs, t_io / 1e6, nthreads, st.ms_h2d, st.ms_ingest, st.ms_scan, st.ms_format, st.ms_d2h, st.ms_inflate, st.launches);
		}
		ngscov_destroy(h);
		for (int k = 0; k < 4; ++k) if (outs[k]) fclose(outs[k]);
		for (int j = 0; j < n_ref; ++j) free(ref_name[j]);
		free(ref_name); free(ref_len); free(hbuf); free(blk);
		munmap((void *)file, fsz); close(fd);
		fprintf(stderr, "Converted %s to wig format at %.3f s\n", fn, (double)(usec() - begin) / 1e6);
	}
	return 0;
}
