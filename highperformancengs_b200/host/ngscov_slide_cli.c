/* host/ngscov_slide_cli.c -- the drop-in bam_sliding_count command line on top of libngscov_cuda (NGSCOV_MODE_SLIDE).
 *
 * Kept from the reference main (reference bam_sliding_count.c:333-435): getopt string "o:w:r:s:h?", defaults
 * (-o "out", -w 20000), atoi parsing, the table <outfile>.txt (output_count_GC, :140-165: first input file only, as the
 * reference prints `databuf`, not `databuf + i`), the progress lines on stderr including the read count, and the -r
 * region mode (:393-409: the region line on stdout, the .bai must exist, only what bam_fetch hands out is counted --
 * here the whole file still goes through the GPU and ngscov_set_region filters on the device).  Not produced:
 * <bam>_hits.png (draw_hits needs libgd, :276-331).
 * The BAM file goes to the device compressed (ngscov_push_bgzf); record framing, SEQ GC counting and the per-window
 * sums run on the GPU, the O(windows) float post-processing of calc_winGC and the printing on the host. */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>
#include <fcntl.h>
#include <getopt.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <zlib.h>
#include "../../include/ngscov.h"

#define DIE(...) do { fprintf(stderr, __VA_ARGS__); exit(1); } while (0)
static long long usec(void) { struct timeval tv; gettimeofday(&tv, NULL); return (long long)tv.tv_sec * 1000000 + tv.tv_usec; }

static void usage(const char *prog)
{
	fprintf(stderr, "\nUsage: %s [-o OUTFILE] [-w WINDOW_SIZE] [-s 0] [-h] bamFile1 bamFile2 ..\n"
	        "  B200-native drop-in for the HighPerformanceNGS bam_sliding_count tool: reads count and GC content per window.\n"
	        "   [-o OUTPUT_FILE]  = output prefix (<prefix>.txt), default out.      [option]\n"
	        "   [-w WINDOW_SIZE]  = window size, default 20000.                     [option]\n"
	        "   [-r]              = region, default is whole genome.(chr1:1-20000) [option]\n"
	        "   [-s]              = accepted and ignored, like the reference.       [option]\n"
	        "   Infiles           = BAM file(s); the table is that of the first.    [required]\n\n", prog);
	exit(1);
}

/* Region strings as bam_parse_region reads them (samtools-0.1.19 bam_aux.c, called at bam_sliding_count.c:401):
 * white space is dropped; what follows the LAST colon is an interval if it consists of digits, commas and at most one
 * hyphen, otherwise the colon belongs to the contig name; a name the header does not know is retried as the whole
 * string; "beg-end" is 1-based inclusive and becomes [beg - 1, end), a lone "beg" runs to 1 << 29, no interval means
 * the whole contig [0, 1 << 29).  Returns the refID or -1. */
static int find_ref(char **ref_name, int n_ref, const char *s)
{
	for (int i = 0; i < n_ref; ++i) if (!strcmp(ref_name[i], s)) return i;
	return -1;
}
static int interval_like(const char *t)
{
	int hyphens = 0;
	for (; *t; ++t) {
		if (*t == '-') { if (++hyphens > 1) return 0; }
		else if (*t != ',' && (*t < '0' || *t > '9')) return 0;
	}
	return 1;
}
static int parse_region(char **ref_name, int n_ref, const char *str, int *beg, int *end)
{
	char *s = (char *)malloc(strlen(str) + 1), *w = s, *colon, *num = NULL;
	int ref;
	for (const char *p = str; *p; ++p) if (*p != ' ' && !(*p >= 9 && *p <= 13)) *w++ = *p;
	*w = 0;
	*beg = 0; *end = 1 << 29;
	colon = strrchr(s, ':');
	if (colon && interval_like(colon + 1)) {
		*colon = 0;
		ref = find_ref(ref_name, n_ref, s);
		if (ref >= 0) num = colon + 1;
		else ref = find_ref(ref_name, n_ref, str);               /* the name itself may end in ":digits" */
	} else ref = find_ref(ref_name, n_ref, str);
	if (ref >= 0 && num) {
		char *d = num, *hy;
		for (char *p = num; *p; ++p) if (*p != ',') *d++ = *p;      /* thousands separators */
		*d = 0;
		*beg = atoi(num);
		hy = strchr(num, '-');
		if (hy) *end = atoi(hy + 1);
		if (*beg > 0) --*beg;
	}
	free(s);
	return ref;
}
static int have_index(const char *fn)
{
	/* bam_index_load_local (samtools bam_index.c): fn.bai, else fn with its "bam" ending turned into "bai" */
	char *p = (char *)malloc(strlen(fn) + 8);
	int ok;
	sprintf(p, "%s.bai", fn);
	ok = access(p, R_OK) == 0;
	if (!ok) { const size_t l = strlen(fn); if (l > 3 && !strcmp(fn + l - 3, "bam")) { strcpy(p, fn); strcpy(p + l - 3, "bai"); ok = access(p, R_OK) == 0; } }
	free(p);
	return ok;
}

int main(int argc, char *argv[])
{
	const char *outfile = "out", *region = "-"; int window = 20000, opt;
	if (argc < 2) usage(argv[0]);
	while ((opt = getopt(argc, argv, "o:w:r:s:h?")) != -1) {
		switch (opt) {
		case 'o': outfile = optarg; break;
		case 'w': window = atoi(optarg); break;
		case 'r': region = optarg; break;          /* (the reference falls through into -s here) */
		case 's': break;
		case '?': case 'h': usage(argv[0]); break;
		default: fprintf(stderr, "error parameter!\n"); break;
		}
	}
	if (optind >= argc) usage(argv[0]);
	if (window < 1) DIE("bam_sliding_count: bad window size\n");
	const long long begin = usec();
	const char *fn = argv[optind];
	int fd = open(fn, O_RDONLY);
	struct stat sb;
	if (fd < 0 || fstat(fd, &sb) != 0 || sb.st_size < 28) DIE("bam2bed: Fail to open BAM file %s\n", fn);
	const size_t fsz = (size_t)sb.st_size;
	const uint8_t *file = (const uint8_t *)mmap(NULL, fsz, PROT_READ, MAP_PRIVATE, fd, 0);
	if (file == MAP_FAILED) DIE("bam2bed: cannot map %s\n", fn);
	/* header: inflate leading blocks until it is complete (samtools bam.c:81-121 layout) */
	size_t off = 0, hbytes = 0, hcap = 1 << 20, hdr_len = 0;
	uint8_t *hbuf = (uint8_t *)malloc(hcap);
	int n_ref = -1; uint32_t *ref_len = NULL; char **ref_name = NULL;
	while (n_ref < 0) {
		if (off + 18 > fsz) DIE("bam2bed: %s has no BAM header\n", fn);
		const uint8_t *p = file + off;
		if (p[0] != 0x1f || p[1] != 0x8b || p[12] != 'B' || p[13] != 'C') DIE("bam2bed: %s is not a BGZF file\n", fn);
		const size_t bsize = (size_t)(p[16] | p[17] << 8) + 1;
		uint32_t ulen; memcpy(&ulen, p + bsize - 4, 4);
		if (hbytes + ulen > hcap) { hcap = (hbytes + ulen) * 2; hbuf = (uint8_t *)realloc(hbuf, hcap); }
		z_stream zs; memset(&zs, 0, sizeof zs);
		if (inflateInit2(&zs, -15) != Z_OK) DIE("bam2bed: zlib\n");
		zs.next_in = (Bytef *)(p + 18); zs.avail_in = (uInt)(bsize - 26); zs.next_out = hbuf + hbytes; zs.avail_out = ulen;
		if (ulen && inflate(&zs, Z_FINISH) != Z_STREAM_END) DIE("bam2bed: inflate failed in %s\n", fn);
		inflateEnd(&zs);
		hbytes += ulen; off += bsize;
		if (hbytes < 12) continue;
		if (memcmp(hbuf, "BAM\1", 4)) DIE("bam2bed: %s is not a BAM file\n", fn);
		uint32_t l_text; memcpy(&l_text, hbuf + 4, 4);
		size_t q = 8 + (size_t)l_text;
		if (q + 4 > hbytes) continue;
		int32_t nr; memcpy(&nr, hbuf + q, 4); q += 4;
		size_t q0 = q; int complete = 1;
		for (int i = 0; i < nr; ++i) {
			if (q + 4 > hbytes) { complete = 0; break; }
			uint32_t ln; memcpy(&ln, hbuf + q, 4);
			if (q + 4 + ln + 4 > hbytes) { complete = 0; break; }
			q += 4 + ln + 4;
		}
		if (!complete) continue;
		n_ref = nr; hdr_len = q;
		ref_len = (uint32_t *)calloc((size_t)nr + 1, 4); ref_name = (char **)calloc((size_t)nr + 1, sizeof(char *));
		q = q0;
		for (int i = 0; i < nr; ++i) {
			uint32_t ln; memcpy(&ln, hbuf + q, 4); q += 4;
			ref_name[i] = strndup((const char *)hbuf + q, ln); q += ln;
			memcpy(&ref_len[i], hbuf + q, 4); q += 4;
		}
	}
	/* whole blocks of the file */
	size_t end = 0;
	while (end + 18 <= fsz) { const size_t bs = (size_t)(file[end + 16] | file[end + 17] << 8) + 1; if (end + bs > fsz) break; end += bs; }

	ngscov_cfg cfg; memset(&cfg, 0, sizeof cfg);
	cfg.mode = NGSCOV_MODE_SLIDE; cfg.window = (uint32_t)window; cfg.n_ref = n_ref; cfg.ref_len = ref_len;
	cfg.ref_name = (const char *const *)ref_name; cfg.device = getenv("NGSCOV_DEVICE") ? atoi(getenv("NGSCOV_DEVICE")) : 0;
	ngscov_t *h = NULL;
	int rc = ngscov_create(&h, &cfg);
	if (rc) DIE("bam2bed: GPU engine unavailable: %s\n", ngscov_strerror(rc));
	if (strncmp(region, "-", 6) != 0) {                              /* the reference's test for "no region" (:389) */
		int ref, rbeg, rend;
		if (!have_index(fn)) { fprintf(stderr, "bam2bed: BAM indexing file is not available.\n"); return 1; }
		ref = parse_region(ref_name, n_ref, region, &rbeg, &rend);
		if (ref < 0) { fprintf(stderr, "bam2bed: Invalid region %s\n", region); return 1; }
		fprintf(stdout, "%s\t%d\t%d\n", ref_name[ref], rbeg, rend);
		fflush(stdout);
		if ((rc = ngscov_set_region(h, ref, rbeg, rend))) DIE("bam2bed: %s\n", ngscov_strerror(rc));
	}
	if ((rc = ngscov_push_bgzf(h, file, end, hdr_len, NGSCOV_PUSH_LAST))) DIE("bam2bed: %s: %s (%s)\n", fn, ngscov_strerror(rc), ngscov_last_cuda_error(h));
	if ((rc = ngscov_finish(h))) DIE("bam2bed: %s: %s (%s)\n", fn, ngscov_strerror(rc), ngscov_last_cuda_error(h));
	fprintf(stderr, "Done load bam file %s at %.3f s\n", fn, (double)(usec() - begin) / 1e6);
	ngscov_stats_t st; ngscov_stats(h, &st);
	fprintf(stderr, "%lu\n", (unsigned long)st.kept);                    /* n_count: reads counted (:104, :421) */
	const char *txt; size_t len;
	if ((rc = ngscov_output(h, NGSCOV_OUT_SLIDE, &txt, &len))) DIE("bam2bed: %s\n", ngscov_strerror(rc));
	char *name = (char *)malloc(strlen(outfile) + 8);
	sprintf(name, "%s.txt", outfile);
	FILE *f = fopen(name, "wb");
	if (!f) DIE("Failed to open file (%s)\n", name);
	if (len && fwrite(txt, 1, len, f) != len) DIE("bam2bed: write failed\n");
	fclose(f);
	fprintf(stderr, "Done output %s.txt at %.3f s\n", outfile, (double)(usec() - begin) / 1e6);
	ngscov_destroy(h);
	for (int j = 0; j < n_ref; ++j) free(ref_name[j]);
	free(ref_name); free(ref_len); free(hbuf); free(name);
	munmap((void *)file, fsz); close(fd);
	return 0;
}
