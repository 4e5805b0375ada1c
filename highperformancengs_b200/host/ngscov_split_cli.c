/* host/ngscov_split_cli.c -- `bamSplitChr [-o PREFIX] [-1 x | -u x] in.bam ...`: one BAM per contig.
 *
 * SURVEY.md section 8(f) rank 4.  Replaces reference bamSplitChr.c:82-146: for every contig j of the header, in header
 * order, `<prefix>_<name>.bam` receives the header and every record bam_fetch(j, 0, 1<<29) hands to samwrite
 * (bamSplitChr.c:55-61, 72-80).  On a coordinate-sorted file that is the run of records with refID == j and
 * 0 <= pos < 2^29 (bam_iter_read / is_overlap, samtools-0.1.19 bam_index.c), copied bit for bit.
 *
 * The output is byte-identical to the reference's when both link the same zlib, because the BGZF writer's blocking is
 * restated exactly (samtools-0.1.19 bgzf.c / bam.c):
 *   - the header stream is cut every 0xff00 bytes and flushed (bam_header_write ends with bgzf_flush);
 *   - before a record of 4 + block_size bytes is appended, the open block is flushed if the record would not fit in
 *     0xff00 bytes (bam_write1_core -> bgzf_flush_try); a block that reaches 0xff00 bytes is flushed at once;
 *   - each block is one raw deflate stream, deflateInit2(level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY), 18-byte
 *     header with BSIZE, CRC32 + ISIZE footer (bgzf_compress); closing writes an empty block at the default level.
 * What is different is the schedule: the input blocks are inflated and the output blocks deflated by one zlib worker
 * per core, a window of the file at a time; the sequential part only reads the first 12 bytes of each record.  Like the
 * reference, the tool insists that the .bai exists (bamSplitChr.c:63-70) although it does not need it.
 * Host-only code: there is no arithmetic here for a GPU to do.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>
#include <getopt.h>
#include <fcntl.h>
#include <limits.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <sys/uio.h>
#include <zlib.h>

#define DIE(...) do { fprintf(stderr, __VA_ARGS__); exit(1); } while (0)
#define BATCH 4096                       /* BGZF blocks inflated per round (<= 256 MiB) */
#define SLACK (64u << 20)                /* carried-over tail: an open output block and a partial record */
#define BLOCK 0xff00u                    /* BGZF_BLOCK_SIZE of samtools-0.1.19 bgzf.h:36 */

typedef struct { const uint8_t *src; uint32_t clen, ulen; uint8_t *dst; } iblk_t;
typedef struct { const uint8_t *src; uint32_t len; uint8_t *dst; uint32_t dlen; } oblk_t;
typedef struct { iblk_t *in; size_t n_in; oblk_t *out; size_t n_out; int level; volatile size_t next; int err; } job_t;

/* bgzf_compress (bgzf.c:175-200) */
static int bgzf_block(uint8_t *dst, uint32_t *dlen, const uint8_t *src, uint32_t slen, int level)
{
	static const uint8_t magic[16] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0 };
	z_stream zs; memset(&zs, 0, sizeof zs);
	zs.next_in = (Bytef *)src; zs.avail_in = slen; zs.next_out = dst + 18; zs.avail_out = 0x10000 - 18 - 8;
	if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
	if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); return -1; }
	deflateEnd(&zs);
	const uint32_t n = (uint32_t)zs.total_out + 26, crc = (uint32_t)crc32(crc32(0L, NULL, 0L), src, slen);
	memcpy(dst, magic, 16);
	dst[16] = (uint8_t)((n - 1) & 0xff); dst[17] = (uint8_t)((n - 1) >> 8);
	memcpy(dst + n - 8, &crc, 4); memcpy(dst + n - 4, &slen, 4);
	*dlen = n;
	return 0;
}

static void *worker(void *arg)
{
	job_t *j = (job_t *)arg;
	if (j->n_in) {
		z_stream zs; memset(&zs, 0, sizeof zs);
		if (inflateInit2(&zs, -15) != Z_OK) { j->err = 1; return NULL; }
		for (;;) {
			const size_t i = __sync_fetch_and_add(&j->next, 1);
			if (i >= j->n_in) break;
			iblk_t *b = &j->in[i];
			if (!b->ulen) continue;
			inflateReset(&zs);
			zs.next_in = (Bytef *)b->src; zs.avail_in = b->clen; zs.next_out = b->dst; zs.avail_out = b->ulen;
			if (inflate(&zs, Z_FINISH) != Z_STREAM_END || zs.total_out != b->ulen) j->err = 1;
		}
		inflateEnd(&zs);
	} else {
		for (;;) {
			const size_t i = __sync_fetch_and_add(&j->next, 1);
			if (i >= j->n_out) break;
			oblk_t *b = &j->out[i];
			if (bgzf_block(b->dst, &b->dlen, b->src, b->len, j->level)) j->err = 1;
		}
	}
	return NULL;
}
static int run_pool(job_t *j, int nthreads)
{
	pthread_t th[256]; int started = 0;
	const size_t n = j->n_in ? j->n_in : j->n_out;
	if (nthreads > 256) nthreads = 256;
	j->next = 0; j->err = 0;
	for (int i = 0; i < nthreads - 1 && (size_t)i + 1 < n; ++i) if (pthread_create(&th[started], NULL, worker, j) == 0) ++started;
	worker(j);
	for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
	return j->err;
}

static long long usec(void) { struct timeval tv; gettimeofday(&tv, NULL); return (long long)tv.tv_sec * 1000000 + tv.tv_usec; }

/* what the sequential walk asks the writer to do, in order */
enum { OP_OPEN, OP_BLOCK, OP_CLOSE };
typedef struct { int op; int32_t tid; size_t blk; uint32_t count; } op_t;
typedef struct { op_t *v; size_t n, cap; } ops_t;
static void op_push(ops_t *o, int op, int32_t tid, size_t blk, uint32_t count)
{
	if (o->n == o->cap) { o->cap = o->cap ? o->cap * 2 : 4096; o->v = (op_t *)realloc(o->v, o->cap * sizeof(op_t)); }
	o->v[o->n].op = op; o->v[o->n].tid = tid; o->v[o->n].blk = blk; o->v[o->n++].count = count;
}
typedef struct { oblk_t *v; size_t n, cap; } oblks_t;
static size_t oblk_push(oblks_t *o, const uint8_t *src, uint32_t len)
{
	if (o->n == o->cap) { o->cap = o->cap ? o->cap * 2 : 8192; o->v = (oblk_t *)realloc(o->v, o->cap * sizeof(oblk_t)); }
	o->v[o->n].src = src; o->v[o->n].len = len; o->v[o->n].dst = NULL; o->v[o->n].dlen = 0;
	return o->n++;
}

static void write_all(int fd, struct iovec *iov, int n, const char *fn)
{
	while (n > 0) {
		ssize_t w = writev(fd, iov, n > IOV_MAX ? IOV_MAX : n);
		if (w < 0) DIE("bamSplitChr: write to %s failed\n", fn);
		while (n > 0 && (size_t)w >= iov->iov_len) { w -= (ssize_t)iov->iov_len; ++iov; --n; }
		if (n > 0 && w > 0) { iov->iov_base = (char *)iov->iov_base + w; iov->iov_len -= (size_t)w; }
	}
}

static int have_index(const char *fn)
{
	/* bam_index_load_local (samtools bam_index.c): fn.bai, else fn with its "bam" ending turned into "bai" */
	char p[PATH_MAX];
	snprintf(p, sizeof p, "%s.bai", fn);
	if (access(p, R_OK) == 0) return 1;
	const size_t l = strlen(fn);
	if (l > 3 && l < sizeof p && strcmp(fn + l - 3, "bam") == 0) { strcpy(p, fn); strcpy(p + l - 3, "bai"); if (access(p, R_OK) == 0) return 1; }
	return 0;
}

static void split_file(const char *fn, const char *prefix, int level, int nthreads, long long begin)
{
	int fd = open(fn, O_RDONLY);
	struct stat sb;
	if (fd < 0 || fstat(fd, &sb) != 0 || sb.st_size < 28) DIE("bam2bed: Fail to open BAM file %s\n", fn);
	if (!have_index(fn)) DIE("bam2bed: BAM indexing file is not available.\n");
	const size_t fsz = (size_t)sb.st_size;
	const uint8_t *file = (const uint8_t *)mmap(NULL, fsz, PROT_READ, MAP_PRIVATE, fd, 0);
	if (file == MAP_FAILED) DIE("bamSplitChr: cannot map %s\n", fn);
	size_t nblk = 0, cap = 1 << 16, off = 0;
	iblk_t *blk = (iblk_t *)malloc(cap * sizeof(iblk_t));
	while (off + 18 <= fsz) {
		const uint8_t *p = file + off;
		if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4) || p[12] != 'B' || p[13] != 'C') DIE("bamSplitChr: %s is not a BGZF file\n", fn);
		const size_t bsize = (size_t)(p[16] | p[17] << 8) + 1;
		if (bsize < 26 || off + bsize > fsz) DIE("bamSplitChr: truncated BGZF block in %s\n", fn);
		if (nblk == cap) { cap *= 2; blk = (iblk_t *)realloc(blk, cap * sizeof(iblk_t)); }
		blk[nblk].src = p + 18; blk[nblk].clen = (uint32_t)(bsize - 26);
		memcpy(&blk[nblk].ulen, p + bsize - 4, 4);
		if (blk[nblk].ulen > 65536) DIE("bamSplitChr: bad BGZF block size in %s\n", fn);
		++nblk; off += bsize;
	}
	const size_t bufcap = (size_t)BATCH * 65536 + SLACK;
	uint8_t *buf = (uint8_t *)malloc(bufcap);
	size_t next_blk = 0, fill = 0;               /* buf[0, fill) = inflated bytes not yet let go */
	job_t job; memset(&job, 0, sizeof job);

	/* ---- header: inflate until it is complete (bam_header_read, samtools bam.c:81-121) */
	uint64_t hdr_need = 12, hdr_len = 0; int32_t n_ref = 0;
	for (;;) {
		while (fill < hdr_need && next_blk < nblk) {
			const size_t b0 = next_blk;
			while (next_blk < nblk && next_blk - b0 < 64 && fill + 65536 <= bufcap) { blk[next_blk].dst = buf + fill; fill += blk[next_blk].ulen; ++next_blk; }
			if (next_blk == b0) DIE("bamSplitChr: header of %s is too large\n", fn);
			job.in = &blk[b0]; job.n_in = next_blk - b0;
			if (run_pool(&job, nthreads)) DIE("bamSplitChr: inflate failed in %s\n", fn);
		}
		if (fill < hdr_need) DIE("bamSplitChr: truncated header in %s\n", fn);
		if (memcmp(buf, "BAM\1", 4)) DIE("bamSplitChr: %s is not a BAM file\n", fn);
		uint32_t l_text; memcpy(&l_text, buf + 4, 4);
		uint64_t q = 8 + (uint64_t)l_text;
		if (q + 4 > fill) { hdr_need = q + 4; continue; }
		memcpy(&n_ref, buf + q, 4); q += 4;
		int again = 0;
		for (int i = 0; i < n_ref; ++i) {
			if (q + 4 > fill) { hdr_need = q + 4; again = 1; break; }
			uint32_t ln; memcpy(&ln, buf + q, 4);
			q += 4 + (uint64_t)ln + 4;
		}
		if (again) continue;
		if (q > fill) { hdr_need = q; continue; }
		hdr_len = q;
		break;
	}
	/* the header stream every output starts with (bam_header_write, bam.c:123-158: names are written as C strings) */
	uint32_t l_text; memcpy(&l_text, buf + 4, 4);
	char **names = (char **)calloc((size_t)n_ref + 1, sizeof(char *));
	int32_t *lens = (int32_t *)calloc((size_t)n_ref + 1, sizeof(int32_t));
	uint8_t *hs = (uint8_t *)malloc((size_t)hdr_len + 16); size_t hn = 0;
	memcpy(hs, buf, 8 + (size_t)l_text + 4); hn = 8 + (size_t)l_text + 4;
	{
		uint64_t q = 8 + (uint64_t)l_text + 4;
		for (int i = 0; i < n_ref; ++i) {
			uint32_t ln; memcpy(&ln, buf + q, 4);
			names[i] = strndup((const char *)buf + q + 4, ln);
			memcpy(&lens[i], buf + q + 4 + ln, 4);
			const int32_t nl = (int32_t)strlen(names[i]) + 1;
			memcpy(hs + hn, &nl, 4); memcpy(hs + hn + 4, names[i], (size_t)nl); memcpy(hs + hn + 4 + nl, &lens[i], 4);
			hn += 4 + (size_t)nl + 4; q += 4 + (uint64_t)ln + 4;
		}
	}
	const size_t n_hb = (hn + BLOCK - 1) / BLOCK;
	oblk_t *hb = (oblk_t *)calloc(n_hb + 1, sizeof(oblk_t));
	uint8_t *hdst = (uint8_t *)malloc((n_hb + 1) * 65536);
	for (size_t i = 0; i < n_hb; ++i) { hb[i].src = hs + i * BLOCK; hb[i].len = (uint32_t)(hn - i * BLOCK < BLOCK ? hn - i * BLOCK : BLOCK); hb[i].dst = hdst + i * 65536; }
	job.in = NULL; job.n_in = 0; job.out = hb; job.n_out = n_hb; job.level = level;
	if (run_pool(&job, nthreads)) DIE("bamSplitChr: deflate failed\n");
	uint8_t eof_blk[64]; uint32_t eof_len = 0;
	if (bgzf_block(eof_blk, &eof_len, (const uint8_t *)"", 0, Z_DEFAULT_COMPRESSION)) DIE("bamSplitChr: deflate failed\n");

	/* ---- records */
	size_t s = (size_t)hdr_len;                 /* next record, as an offset into buf */
	size_t open_at = s; uint32_t open_len = 0;   /* the output block being filled: buf[open_at, open_at + open_len) */
	int32_t cur = -1;                            /* contig whose file is open */
	uint32_t count = 0;
	int out_fd = -1; char out_fn[PATH_MAX] = "";
	ops_t ops; memset(&ops, 0, sizeof ops);
	oblks_t ob; memset(&ob, 0, sizeof ob);
	uint8_t *arena = NULL; size_t arena_cap = 0;
	struct iovec *iov = NULL; size_t iov_cap = 0;
	int input_done = 0, tail = 0;                /* tail: records without a contig (refID -1) reached: nothing left to copy */
	while (!input_done) {
		/* let go of what is done, append a batch of blocks */
		const size_t keep = open_len ? open_at : s;
		if (keep) { memmove(buf, buf + keep, fill - keep); fill -= keep; s -= keep; open_at -= keep; }
		const size_t b0 = next_blk;
		while (next_blk < nblk && next_blk - b0 < BATCH && fill + 65536 <= bufcap) { blk[next_blk].dst = buf + fill; fill += blk[next_blk].ulen; ++next_blk; }
		if (next_blk > b0) {
			job.in = &blk[b0]; job.n_in = next_blk - b0;
			if (run_pool(&job, nthreads)) DIE("bamSplitChr: inflate failed in %s\n", fn);
		} else if (next_blk < nblk) DIE("bamSplitChr: record larger than 64 MiB in %s\n", fn);
		input_done = next_blk >= nblk;
		ops.n = 0; ob.n = 0;
		/* walk the complete records of the window */
		while (!tail && s + 4 <= fill) {
			int32_t bsz; memcpy(&bsz, buf + s, 4);
			if (bsz < 32) DIE("bamSplitChr: malformed record in %s\n", fn);
			const size_t rec = 4 + (size_t)bsz;
			if (s + rec > fill) break;
			int32_t tid, pos; memcpy(&tid, buf + s + 4, 4); memcpy(&pos, buf + s + 8, 4);
			if (tid >= n_ref) DIE("bamSplitChr: bad refID in %s\n", fn);
			if (tid < 0) { tail = 1; break; }
			if (tid < cur) DIE("bamSplitChr: %s is not sorted by coordinate\n", fn);
			if (pos < 0 || pos >= (1 << 29)) {
				/* not handed out by bam_fetch(0, 1<<29): take it out of the stream so that blocks stay contiguous */
				memmove(buf + s, buf + s + rec, fill - s - rec); fill -= rec;
				continue;
			}
			if (tid != cur) {
				if (cur >= 0) {
					if (open_len) { op_push(&ops, OP_BLOCK, cur, oblk_push(&ob, buf + open_at, open_len), 0); open_len = 0; }
					op_push(&ops, OP_CLOSE, cur, 0, count);
				}
				for (int32_t t = cur + 1; t < tid; ++t) { op_push(&ops, OP_OPEN, t, 0, 0); op_push(&ops, OP_CLOSE, t, 0, 0); }
				op_push(&ops, OP_OPEN, tid, 0, 0);
				cur = tid; count = 0; open_at = s;
			}
			/* bam_write1_core -> bgzf_flush_try(4 + block_len), then bgzf_write */
			if (open_len && open_len + rec > BLOCK) { op_push(&ops, OP_BLOCK, cur, oblk_push(&ob, buf + open_at, open_len), 0); open_at += open_len; open_len = 0; }
			if (!open_len) open_at = s;
			size_t add = rec;
			while (open_len + add >= BLOCK) {
				const uint32_t take = BLOCK - open_len;
				op_push(&ops, OP_BLOCK, cur, oblk_push(&ob, buf + open_at, BLOCK), 0);
				open_at += BLOCK; open_len = 0; add -= take;
			}
			open_len += (uint32_t)add;
			++count; s += rec;
		}
		if (input_done || tail) {
			if (!tail && s != fill) DIE("bamSplitChr: truncated record in %s\n", fn);
			if (cur >= 0) {
				if (open_len) { op_push(&ops, OP_BLOCK, cur, oblk_push(&ob, buf + open_at, open_len), 0); open_len = 0; }
				op_push(&ops, OP_CLOSE, cur, 0, count);
			}
			for (int32_t t = cur + 1; t < n_ref; ++t) { op_push(&ops, OP_OPEN, t, 0, 0); op_push(&ops, OP_CLOSE, t, 0, 0); }
			cur = n_ref; input_done = 1;
		}
		/* deflate this round's blocks, one worker per core */
		if (ob.n * 65536 > arena_cap) { arena_cap = ob.n * 65536; free(arena); arena = (uint8_t *)malloc(arena_cap); if (!arena) DIE("bamSplitChr: out of memory\n"); }
		for (size_t i = 0; i < ob.n; ++i) ob.v[i].dst = arena + i * 65536;
		if (ob.n) {
			job.in = NULL; job.n_in = 0; job.out = ob.v; job.n_out = ob.n; job.level = level;
			if (run_pool(&job, nthreads)) DIE("bamSplitChr: deflate failed\n");
		}
		/* write in order */
		if (ops.n + n_hb + 2 > iov_cap) { iov_cap = ops.n + n_hb + 2; free(iov); iov = (struct iovec *)malloc(iov_cap * sizeof(struct iovec)); }
		int niov = 0;
		for (size_t i = 0; i < ops.n; ++i) {
			const op_t *o = &ops.v[i];
			if (o->op == OP_OPEN) {
				/* bamSplitChr.c:128 builds the name in a 256-byte buffer; longer names are an error here, not an overflow */
				if (snprintf(out_fn, sizeof out_fn, "%s_%s.bam", prefix, names[o->tid]) >= (int)sizeof out_fn) DIE("bamSplitChr: output name too long\n");
				out_fd = open(out_fn, O_WRONLY | O_CREAT | O_TRUNC, 0666);
				if (out_fd < 0) DIE("[main_samview] fail to open \"%s\" for writing.\n", out_fn);
				niov = 0;
				for (size_t k = 0; k < n_hb; ++k) { iov[niov].iov_base = hb[k].dst; iov[niov++].iov_len = hb[k].dlen; }
			} else if (o->op == OP_BLOCK) {
				iov[niov].iov_base = ob.v[o->blk].dst; iov[niov++].iov_len = ob.v[o->blk].dlen;
			} else {
				iov[niov].iov_base = eof_blk; iov[niov++].iov_len = eof_len;
				write_all(out_fd, iov, niov, out_fn); niov = 0;
				if (close(out_fd) != 0) DIE("bamSplitChr: write to %s failed\n", out_fn);
				out_fd = -1;
				fprintf(stderr, "chr: %s\tchr_len: %d\treads_count: %u at %.3f s\n", names[o->tid], lens[o->tid], o->count, (double)(usec() - begin) / 1e6);
			}
		}
		if (niov) { write_all(out_fd, iov, niov, out_fn); niov = 0; }
	}
	fprintf(stderr, "splited %s into each chromosome at %.3f s\n", fn, (double)(usec() - begin) / 1e6);
	for (int i = 0; i < n_ref; ++i) free(names[i]);
	free(names); free(lens); free(hs); free(hb); free(hdst); free(arena); free(iov); free(ops.v); free(ob.v); free(buf); free(blk);
	munmap((void *)file, fsz); close(fd);
}

static void usage(const char *a0)
{
	fprintf(stderr, "\nUsage: %s [-o OUTFILE_PREFIX] [-h] bamFile1 bamFile2 ..\n"
		"  Splits coordinate-sorted, indexed BAM files into one BAM per contig: <prefix>_<contig>.bam\n"
		"   [-o OUTPUT_FILE]  = output prefix (default: the input file name)\n"
		"   -1 x                 fast BAM compression (takes an argument, as in the reference's option string)\n"
		"   -u x                 uncompressed BAM output (takes an argument, as in the reference's option string)\n"
		"   [-h]              = this help\n\n", a0);
	exit(1);
}

int main(int argc, char **argv)
{
	/* the reference's option string, bamSplitChr.c:89 ("u:" and "1:" consume an argument there, so they do here) */
	const char *optstr = "o:w:r:s:u:1:h?";
	const char *prefix = NULL; int level = -1, opt;
	if (argc < 2) usage(argv[0]);
	while ((opt = getopt(argc, argv, optstr)) != -1) {
		switch (opt) {
		case 'o': prefix = optarg; break;
		case 'u': level = 0; break;
		case '1': level = 1; break;
		case '?': case 'h': usage(argv[0]); break;
		default: fprintf(stderr, "error parameter!\n"); break;
		}
	}
	int nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
	if (getenv("NGSCOV_THREADS")) nthreads = atoi(getenv("NGSCOV_THREADS"));
	if (nthreads < 1) nthreads = 1;
	const long long begin = usec();
	for (int i = optind; i < argc; ++i) {
		/* bamSplitChr.c:120,142: -o applies to the first input only; afterwards each input is its own prefix */
		split_file(argv[i], prefix ? prefix : argv[i], level < 0 ? Z_DEFAULT_COMPRESSION : level, nthreads, begin);
		prefix = NULL;
	}
	return 0;
}
