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
		if (fd == -1) fprintf(stderr, "Failed to create output file (%s)", name);
	}
	free(name);
	return fdopen(fd, "wb");
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
typedef struct { const uint8_t *src; uint32_t clen, ulen; uint8_t *dst; } blk_t;
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
		if (!have_bai(fn)) DIE("bam2bed: BAM indexing file is not available.\n");

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
			blk[nblk].src = p + 18; blk[nblk].clen = (uint32_t)(bsize - 26);
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
		ngscov_t *h = NULL;
		int rc = ngscov_create(&h, &cfg);
		if (rc) DIE("bam2bed: GPU engine unavailable: %s\n", ngscov_strerror(rc));
		void *stage[2]; size_t scap[2];
		for (int i = 0; i < 2; ++i) if ((rc = ngscov_staging(h, i, &stage[i], &scap[i]))) DIE("bam2bed: %s\n", ngscov_strerror(rc));

		/* stream: [tail of the header blocks] then the remaining blocks, inflated in parallel into staging */
		long long t_io = 0;
		int idx = 0; size_t b = hb;
		size_t pending_hdr_tail = hbytes - hdr_len;            /* record bytes that came with the header blocks */
		int first = 1, pushed_last = 0;
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
			idx ^= 1;
		}
		if ((rc = ngscov_finish(h))) DIE("bam2bed: %s: %s (%s)\n", fn, ngscov_strerror(rc), ngscov_last_cuda_error(h));

		const int which[4] = { NGSCOV_OUT_BEDGRAPH, NGSCOV_OUT_DEPTH, NGSCOV_OUT_WIG, NGSCOV_OUT_CHROMSIZE };
		FILE *outs[4] = { f_bed, f_depth, f_wig, f_chr };
		for (int k = 0; k < 4; ++k) {
			if (!outs[k]) continue;
			const char *buf; size_t len;
			if ((rc = ngscov_output(h, which[k], &buf, &len))) DIE("bam2bed: %s\n", ngscov_strerror(rc));
			if (len && fwrite(buf, 1, len, outs[k]) != len) DIE("bam2bed: write failed\n");
		}
		for (int j = 0; j < n_ref; ++j) fprintf(stderr, "%s at %.3f s\n", ref_name[j], (double)(usec() - begin) / 1e6);
		if (verbose) {
			ngscov_stats_t st; ngscov_stats(h, &st);
			fprintf(stderr, "ngscov: %llu records (%llu kept), %llu runs; host read+inflate %.3f s (%d threads); device: h2d %.3f ms ingest %.3f ms "
			        "scan %.3f ms format %.3f ms d2h %.3f ms, %u kernel launches\n", (unsigned long long)st.records, (unsigned long long)st.kept,
			        (unsigned long long)st.runs, t_io / 1e6, nthreads, st.ms_h2d, st.ms_ingest, st.ms_scan, st.ms_format, st.ms_d2h, st.launches);
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
