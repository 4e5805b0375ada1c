/* tools/bamgen.c -- synthetic coordinate-sorted BAM (+ .bai) generator for the five benchmark/parity
 * configurations of BASELINE.json (spec: SURVEY.md section 8(d) "Synthetic inputs").
 *
 * Test/bench infrastructure: it produces INPUTS only (it computes no coverage).  Everything is
 * derived from a counter-based splitmix64 hash of (seed, contig, unit, read), so output is identical
 * for any thread count and on any machine.
 *
 * A contig is generated as an ordered list of independent "units" whose read start positions do not
 * interleave (16 kbp tiles for WGS, genes for RNA-seq, amplicon ends for the panel), so units are
 * generated in parallel and concatenated: the stream is coordinate sorted by construction.
 *
 * BAM record layout written here follows the SAM/BAM specification as read by samtools-0.1.19
 * bam_read1 (reference tarball, bam.c:191-223) and bam1_core_t (bam.h:178-187); `bin` is
 * bam_reg2bin(pos, calend) (bam.h:723-732) because samtools' indexer trusts it (bam_index.c:197).
 * The .bai follows bam_index_core / bam_index_save (bam_index.c:151-324): binning index with the
 * 37450 meta bin, 16 kbp linear index with fill_missing, trailing n_no_coor.
 *
 * Build:  gcc -O3 -fopenmp -shared -fPIC tools/bamgen.c -o tools/libbamgen.so -lz -lm
 *         gcc -O3 -fopenmp -DBAMGEN_MAIN tools/bamgen.c -o tools/bamgen -lz -lm
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <zlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ PRNG (counter based) */
static inline uint64_t mix64(uint64_t z)
{
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
static inline uint64_t h4(uint64_t seed, uint64_t a, uint64_t b, uint64_t c)
{
	return mix64(mix64(mix64(mix64(seed) ^ a) ^ (b * 0xD6E8FEB86659FD93ull)) ^ (c * 0xA0761D6478BD642Full));
}
static inline double u01(uint64_t x) { return (double)(x >> 11) * (1.0 / 9007199254740992.0); }
static inline double gauss(uint64_t x)          /* Box-Muller from one 64-bit hash */
{
	double a = u01(mix64(x)) + 1e-300, b = u01(mix64(x ^ 0x5555555555555555ull));
	return sqrt(-2.0 * log(a)) * cos(6.283185307179586 * b);
}

/* ------------------------------------------------------------------ plan */
enum { K_TILE = 0, K_GENE = 1, K_AMPLICON = 2, K_UNMAPPED = 3 };
#define TILE_BP 16384u
#define MAX_EXON 12

typedef struct {
	uint32_t start;                 /* first possible read start of the unit */
	uint32_t width;                 /* K_TILE: tile width; K_AMPLICON: jitter width */
	uint64_t n_reads, read0;        /* reads in this unit; global index of its first read */
	uint64_t byte0, nbytes;         /* filled by sizing pass: offset within contig stream */
	int n_exon; uint32_t ex_start[MAX_EXON], ex_len[MAX_EXON]; uint32_t tlen;   /* K_GENE */
} unit_t;

typedef struct {
	char name[64]; uint32_t len; int kind;
	uint64_t n_reads, read0;
	unit_t *unit; size_t n_unit;
	uint64_t nbytes; int sized;
} contig_t;

typedef struct bamgen {
	int config; uint64_t seed; double scale; int seq_mode; uint32_t read_len;
	contig_t *ctg; int n_ref;        /* ctg[n_ref] is the unmapped tail pseudo-contig */
	uint64_t n_reads;
} bamgen_t;

static const uint32_t HG38_LEN[24] = {
	248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
	138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
	83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415 };

static void plan_tiles(bamgen_t *g, contig_t *c, uint32_t maxreflen)
{
	uint32_t last_start = c->len - maxreflen;        /* reads never overhang the contig end */
	size_t nt = (size_t)(last_start / TILE_BP) + 1;
	c->unit = (unit_t *)calloc(nt, sizeof(unit_t));
	c->n_unit = nt;
	uint64_t span = (uint64_t)last_start + 1, done = 0;
	for (size_t t = 0; t < nt; ++t) {
		unit_t *u = &c->unit[t];
		u->start = (uint32_t)(t * TILE_BP);
		uint64_t hi = (t + 1) * (uint64_t)TILE_BP;
		if (hi > span) hi = span;
		u->width = (uint32_t)(hi - u->start);
		/* reads proportional to covered start positions, exact total */
		uint64_t upto = (uint64_t)((long double)c->n_reads * (long double)hi / (long double)span + 0.5L);
		if (t == nt - 1) upto = c->n_reads;
		u->n_reads = upto - done; u->read0 = c->read0 + done; done = upto;
	}
}

static void plan_genes(bamgen_t *g, contig_t *c, int ci)
{
	size_t cap = 4096, n = 0;
	unit_t *U = (unit_t *)calloc(cap, sizeof(unit_t));
	double *w = (double *)calloc(cap, sizeof(double)), wsum = 0;
	uint32_t cursor = 5000;
	for (uint64_t gi = 0; ; ++gi) {
		unit_t u; memset(&u, 0, sizeof u);
		u.n_exon = 2 + (int)(h4(g->seed, ci, gi, 1) % 11);
		uint32_t p = cursor, tl = 0;
		for (int e = 0; e < u.n_exon; ++e) {
			u.ex_start[e] = p;
			u.ex_len[e] = 100 + (uint32_t)(h4(g->seed, ci, gi, 10 + e) % 301);
			tl += u.ex_len[e];
			p += u.ex_len[e];
			if (e + 1 < u.n_exon) p += (uint32_t)(200.0 * exp(u01(h4(g->seed, ci, gi, 40 + e)) * log(100.0))); /* 0.2-20 kb */
		}
		if ((uint64_t)p + 2000 >= c->len) break;
		u.tlen = tl; u.start = cursor;
		if (n == cap) { cap *= 2; U = (unit_t *)realloc(U, cap * sizeof(unit_t)); w = (double *)realloc(w, cap * sizeof(double)); }
		w[n] = exp(2.0 * gauss(h4(g->seed, ci, gi, 2))) * tl; wsum += w[n];
		U[n++] = u;
		cursor = p + 1000 + (uint32_t)(h4(g->seed, ci, gi, 3) % 20000);
	}
	uint64_t done = 0; double acc = 0;
	for (size_t i = 0; i < n; ++i) {
		acc += w[i];
		uint64_t upto = (i == n - 1) ? c->n_reads : (uint64_t)((double)c->n_reads * (acc / wsum));
		if (upto < done) upto = done;
		U[i].n_reads = upto - done; U[i].read0 = c->read0 + done; done = upto;
	}
	free(w);
	c->unit = U; c->n_unit = n;
	if (n == 0) c->n_reads = 0;
}

static void plan_amplicons(bamgen_t *g, contig_t *c, int n_amp)
{
	c->unit = (unit_t *)calloc((size_t)n_amp * 2, sizeof(unit_t));
	c->n_unit = (size_t)n_amp * 2;
	uint32_t pitch = (c->len - 200000) / (uint32_t)n_amp;
	uint64_t done = 0;
	for (size_t i = 0; i < c->n_unit; ++i) {
		unit_t *u = &c->unit[i];
		uint32_t a = 100000 + (uint32_t)(i / 2) * pitch;
		u->start = (i & 1) ? a + 330 : a - 20;       /* forward pile-up at a+-20, reverse at a+350+-20 */
		u->width = 41;
		uint64_t upto = (uint64_t)((long double)c->n_reads * (long double)(i + 1) / (long double)c->n_unit);
		u->n_reads = upto - done; u->read0 = c->read0 + done; done = upto;
	}
}

bamgen_t *bamgen_create(int config, uint64_t seed, double scale, int seq_mode)
{
	bamgen_t *g = (bamgen_t *)calloc(1, sizeof(bamgen_t));
	g->config = config; g->seed = seed ? seed : (uint64_t)config; g->scale = scale > 0 ? scale : 1.0; g->seq_mode = seq_mode;
	uint64_t read0 = 0;
#define SCL(x) ((uint32_t)fmax(2000.0, floor((double)(x) * g->scale + 0.5)))
	if (config == 1) {
		g->n_ref = 1; g->read_len = 100;
		g->ctg = (contig_t *)calloc(2, sizeof(contig_t));
		strcpy(g->ctg[0].name, "chr1"); g->ctg[0].len = SCL(10000000); g->ctg[0].kind = K_TILE;
		g->ctg[0].n_reads = (uint64_t)(1000000 * g->scale + 0.5);
	} else if (config == 2 || config == 3) {
		g->n_ref = 24; g->read_len = 150;
		g->ctg = (contig_t *)calloc(25, sizeof(contig_t));
		for (int i = 0; i < 24; ++i) {
			if (i < 22) sprintf(g->ctg[i].name, "chr%d", i + 1); else strcpy(g->ctg[i].name, i == 22 ? "chrX" : "chrY");
			g->ctg[i].len = SCL(HG38_LEN[i]); g->ctg[i].kind = K_TILE;
			g->ctg[i].n_reads = (uint64_t)(30.0 * g->ctg[i].len / 150.0 + 0.5);
		}
	} else if (config == 4) {
		static const uint32_t L4[3] = { 50000000, 30000000, 10000000 };
		g->n_ref = 3; g->read_len = 100;
		g->ctg = (contig_t *)calloc(4, sizeof(contig_t));
		for (int i = 0; i < 3; ++i) {
			sprintf(g->ctg[i].name, "chr%d", i + 1); g->ctg[i].len = SCL(L4[i]); g->ctg[i].kind = K_GENE;
			g->ctg[i].n_reads = (uint64_t)(20000000.0 * g->scale * L4[i] / 90000000.0 + 0.5);
		}
	} else if (config == 5) {
		g->n_ref = 1; g->read_len = 150;
		g->ctg = (contig_t *)calloc(2, sizeof(contig_t));
		strcpy(g->ctg[0].name, "chr1"); g->ctg[0].len = 50000000; g->ctg[0].kind = K_AMPLICON;
		g->ctg[0].n_reads = (uint64_t)(50000000.0 * g->scale + 0.5);
	} else { free(g); return NULL; }
	for (int i = 0; i < g->n_ref; ++i) {
		contig_t *c = &g->ctg[i];
		c->read0 = read0;
		if (c->kind == K_TILE) plan_tiles(g, c, g->read_len + 8);
		else if (c->kind == K_GENE) plan_genes(g, c, i);
		else plan_amplicons(g, c, 400);
		read0 += c->n_reads;
	}
	/* unmapped tail (tid = -1) for the WGS config */
	contig_t *t = &g->ctg[g->n_ref];
	t->kind = K_UNMAPPED; t->read0 = read0;
	t->n_reads = (config == 2 || config == 3) ? read0 / 1000 : 0;
	t->unit = (unit_t *)calloc(1, sizeof(unit_t)); t->n_unit = 1;
	t->unit[0].n_reads = t->n_reads; t->unit[0].read0 = read0;
	g->n_reads = read0 + t->n_reads;
	return g;
}

void bamgen_destroy(bamgen_t *g)
{
	if (!g) return;
	for (int i = 0; i <= g->n_ref; ++i) free(g->ctg[i].unit);
	free(g->ctg); free(g);
}
int bamgen_n_ref(const bamgen_t *g) { return g->n_ref; }
const char *bamgen_ref_name(const bamgen_t *g, int i) { return g->ctg[i].name; }
uint32_t bamgen_ref_len(const bamgen_t *g, int i) { return g->ctg[i].len; }
uint64_t bamgen_n_reads(const bamgen_t *g) { return g->n_reads; }
uint64_t bamgen_contig_reads(const bamgen_t *g, int i) { return g->ctg[i].n_reads; }

/* ------------------------------------------------------------------ record emission */
typedef struct { uint32_t op[4 * MAX_EXON + 8]; int n; } cigar_t;
static inline void cg_push(cigar_t *c, uint32_t op, uint32_t len)
{
	if (len == 0 && op != 0) return;
	if (c->n && (c->op[c->n - 1] & 0xf) == op) { c->op[c->n - 1] += len << 4; return; }
	c->op[c->n++] = len << 4 | op;
}
static inline uint32_t reg2bin(uint32_t beg, uint32_t end)
{
	--end;
	if (beg >> 14 == end >> 14) return 4681 + (beg >> 14);
	if (beg >> 17 == end >> 17) return 585 + (beg >> 17);
	if (beg >> 20 == end >> 20) return 73 + (beg >> 20);
	if (beg >> 23 == end >> 23) return 9 + (beg >> 23);
	if (beg >> 26 == end >> 26) return 1 + (beg >> 26);
	return 0;
}
static inline uint32_t cigar_reflen(const cigar_t *c)
{
	uint32_t l = 0;
	for (int i = 0; i < c->n; ++i) if ((0x3C1A7 >> ((c->op[i] & 0xf) << 1)) & 2) l += c->op[i] >> 4;
	return l;
}
static inline int u64_digits(uint64_t v) { int d = 1; while (v >= 10) { v /= 10; ++d; } return d; }

/* size of one record in the stream (4-byte block_size prefix included) */
static inline size_t rec_size(const bamgen_t *g, uint64_t ridx, int n_cigar, uint32_t l_seq)
{
	return 4 + 32 + (size_t)(1 + u64_digits(ridx) + 1) + 4 * (size_t)n_cigar + (l_seq + 1) / 2 + l_seq;
}
/* writes the record if out != NULL; returns its size */
static size_t put_rec(const bamgen_t *g, uint8_t *out, uint64_t ridx, int32_t tid, int32_t pos, uint32_t flag,
                      const cigar_t *cg, uint32_t l_seq)
{
	size_t sz = rec_size(g, ridx, cg->n, l_seq);
	if (!out) return sz;
	int nd = u64_digits(ridx);
	const uint64_t ridx0 = ridx;
	uint32_t l_qname = (uint32_t)nd + 2, bin;
	if (tid < 0) bin = 4680;
	else { uint32_t rl = cigar_reflen(cg); bin = reg2bin((uint32_t)pos, (uint32_t)pos + (rl ? rl : 1)); }
	int32_t v;
	uint8_t *p = out;
	v = (int32_t)(sz - 4); memcpy(p, &v, 4); p += 4;
	memcpy(p, &tid, 4); memcpy(p + 4, &pos, 4);
	p[8] = (uint8_t)l_qname; p[9] = tid < 0 ? 0 : 60; p[10] = bin & 0xff; p[11] = bin >> 8;
	p[12] = cg->n & 0xff; p[13] = cg->n >> 8; p[14] = flag & 0xff; p[15] = flag >> 8;
	memcpy(p + 16, &l_seq, 4);
	v = -1; memcpy(p + 20, &v, 4); memcpy(p + 24, &v, 4);
	v = 0; memcpy(p + 28, &v, 4);
	p += 32;
	*p++ = 'r';
	for (int i = nd - 1; i >= 0; --i) { p[i] = '0' + (uint8_t)(ridx % 10); ridx /= 10; }
	p += nd; *p++ = 0;
	memcpy(p, cg->op, 4 * (size_t)cg->n); p += 4 * (size_t)cg->n;
	if (l_seq) {
		uint64_t r = mix64(g->seed ^ (ridx0 * 31 + (uint64_t)(uint32_t)pos));
		size_t nb = (l_seq + 1) / 2;
		for (size_t i = 0; i < nb; ++i) {
			if ((i & 15) == 0) r = mix64(r);
			p[i] = (uint8_t)((1u << ((r >> ((i & 15) * 4)) & 3)) << 4 | (1u << ((r >> ((i & 15) * 4 + 2)) & 3)));
		}
		if (l_seq & 1) p[nb - 1] &= 0xf0;
		p += nb;
		memset(p, 40, l_seq); p += l_seq;
	}
	return sz;
}

static inline uint32_t wgs_flag(uint64_t r)
{
	uint32_t f = (r & 1) ? 16 : 0;
	uint32_t k = (uint32_t)((r >> 8) % 1000);
	if (k < 10) f |= 1024; else if (k < 15) f |= 256; else if (k < 17) f |= 512;
	return f;
}

/* generate one unit; out may be NULL (sizing pass).  tile_first (optional) receives, for K_TILE, nothing. */
static size_t gen_unit(const bamgen_t *g, int ci, const contig_t *c, size_t ui, uint8_t *out)
{
	const unit_t *u = &c->unit[ui];
	size_t total = 0;
	uint32_t l_seq_full = g->read_len;
	cigar_t cg;
	if (c->kind == K_UNMAPPED) {
		cg.n = 0;
		for (uint64_t i = 0; i < u->n_reads; ++i) {
			size_t s = put_rec(g, out ? out + total : NULL, u->read0 + i, -1, -1, 4, &cg, g->seq_mode ? l_seq_full : 0);
			total += s;
		}
		return total;
	}
	if (u->n_reads == 0) return 0;
	uint32_t *pos = (uint32_t *)malloc(u->n_reads * sizeof(uint32_t));
	if (c->kind == K_TILE || c->kind == K_AMPLICON) {
		/* counting sort of uniform offsets within the unit */
		uint32_t W = u->width;
		uint32_t *cnt = (uint32_t *)calloc((size_t)W + 1, sizeof(uint32_t));
		for (uint64_t i = 0; i < u->n_reads; ++i) cnt[h4(g->seed, ci, ui, i) % W]++;
		uint64_t k = 0;
		for (uint32_t x = 0; x < W; ++x) for (uint32_t j = 0; j < cnt[x]; ++j) pos[k++] = u->start + x;
		free(cnt);
	} else { /* K_GENE: transcript offsets, sorted */
		uint32_t maxspan = g->read_len + 4;
		uint32_t W = u->tlen > maxspan ? u->tlen - maxspan : 1;
		uint32_t *cnt = (uint32_t *)calloc((size_t)W + 1, sizeof(uint32_t));
		for (uint64_t i = 0; i < u->n_reads; ++i) cnt[h4(g->seed, ci, ui, i) % W]++;
		uint64_t k = 0;
		for (uint32_t x = 0; x < W; ++x) for (uint32_t j = 0; j < cnt[x]; ++j) pos[k++] = x;
		free(cnt);
	}
	for (uint64_t i = 0; i < u->n_reads; ++i) {
		uint64_t ridx = u->read0 + i, r = h4(g->seed ^ 0xABCDEF, ci, ui, i);
		uint32_t flag = 0, p = pos[i], L = g->read_len;
		cg.n = 0;
		if (g->config == 1) {
			cg_push(&cg, 0, L);
		} else if (g->config == 2 || g->config == 3) {
			flag = wgs_flag(r);
			uint32_t k = (uint32_t)((r >> 20) % 100);
			if (k == 0) { cg_push(&cg, 0, 70); cg_push(&cg, 2, 2); cg_push(&cg, 0, 80); }
			else if (k == 1) { cg_push(&cg, 4, 5); cg_push(&cg, 0, 145); }
			else if (k == 2) { cg_push(&cg, 0, 60); cg_push(&cg, 1, 3); cg_push(&cg, 0, 87); }
			else cg_push(&cg, 0, L);
		} else if (g->config == 5) {
			flag = (ui & 1) ? 16 : 0;
			if ((r >> 8) % 100 < 2) flag |= 1024;
			uint32_t k = (uint32_t)((r >> 20) % 10);
			if (k < 7) cg_push(&cg, 0, L);
			else if (ui & 1) { cg_push(&cg, 4, 30); cg_push(&cg, 0, L - 30); }
			else { cg_push(&cg, 0, L - 30); cg_push(&cg, 4, 30); }
		} else { /* config 4: build query ops in transcript space then splice over introns */
			uint32_t k = (uint32_t)((r >> 20) % 100), fk = (uint32_t)((r >> 8) % 100);
			flag = (r & 1) ? 16 : 0;
			if (fk < 3) flag |= 256; else if (fk < 6) flag |= 1024; else if (fk < 7) flag |= 2048;
			uint32_t qop[6], qlen[6]; int nq = 0;
			uint32_t a = 10 + (uint32_t)((r >> 32) % 60);
			if (k < 5)       { qop[0] = 0; qlen[0] = a; qop[1] = 1; qlen[1] = 3; qop[2] = 0; qlen[2] = L - 3 - a; nq = 3; }
			else if (k < 10) { qop[0] = 0; qlen[0] = a; qop[1] = 2; qlen[1] = 2; qop[2] = 0; qlen[2] = L - a; nq = 3; }
			else if (k < 15) { qop[0] = 4; qlen[0] = 5; qop[1] = 0; qlen[1] = L - 5; nq = 2; }
			else if (k < 20) { qop[0] = 0; qlen[0] = L - 8; qop[1] = 4; qlen[1] = 8; nq = 2; }
			else if (k < 21) { qop[0] = 0; qlen[0] = 20; qop[1] = 7; qlen[1] = 5; qop[2] = 8; qlen[2] = 5; qop[3] = 0; qlen[3] = L - 30; nq = 4; }
			else             { qop[0] = 0; qlen[0] = L; nq = 1; }
			/* locate transcript offset p in the exon list */
			int e = 0; uint32_t toff = p;
			while (e < u->n_exon && toff >= u->ex_len[e]) { toff -= u->ex_len[e]; ++e; }
			uint32_t gpos = u->ex_start[e] + toff, room = u->ex_len[e] - toff;
			p = gpos;
			for (int q = 0; q < nq; ++q) {
				uint32_t op = qop[q], len = qlen[q];
				if (!((0x3C1A7 >> (op << 1)) & 2)) { cg_push(&cg, op, len); continue; }   /* I, S: no reference */
				while (len) {
					if (room == 0) {              /* cross the intron */
						cg_push(&cg, 3, u->ex_start[e + 1] - (u->ex_start[e] + u->ex_len[e]));
						++e; room = u->ex_len[e];
					}
					uint32_t t = len < room ? len : room;
					cg_push(&cg, op, t);
					len -= t; room -= t;
				}
			}
		}
		size_t s = put_rec(g, out ? out + total : NULL, ridx, ci, (int32_t)p, flag, &cg, g->seq_mode ? L : 0);
		total += s;
	}
	free(pos);
	return total;
}

/* sizing pass for contig ci (idempotent) */
static void size_contig(bamgen_t *g, int ci)
{
	contig_t *c = &g->ctg[ci];
	if (c->sized) return;
	#pragma omp parallel for schedule(dynamic, 16)
	for (long ui = 0; ui < (long)c->n_unit; ++ui) c->unit[ui].nbytes = gen_unit(g, ci, c, (size_t)ui, NULL);
	uint64_t off = 0;
	for (size_t ui = 0; ui < c->n_unit; ++ui) { c->unit[ui].byte0 = off; off += c->unit[ui].nbytes; }
	c->nbytes = off; c->sized = 1;
}

/* bytes of the record stream of contigs [c0, c1); index n_ref is the unmapped tail */
uint64_t bamgen_records_size(bamgen_t *g, int c0, int c1)
{
	uint64_t t = 0;
	for (int ci = c0; ci < c1 && ci <= g->n_ref; ++ci) { size_contig(g, ci); t += g->ctg[ci].nbytes; }
	return t;
}

/* generate the record stream of contigs [c0,c1) into buf (capacity cap).  Returns bytes written, 0 on
 * overflow. */
uint64_t bamgen_records(bamgen_t *g, int c0, int c1, uint8_t *buf, uint64_t cap)
{
	uint64_t need = bamgen_records_size(g, c0, c1), base = 0;
	if (need > cap) return 0;
	for (int ci = c0; ci < c1 && ci <= g->n_ref; ++ci) {
		contig_t *c = &g->ctg[ci];
		#pragma omp parallel for schedule(dynamic, 16)
		for (long ui = 0; ui < (long)c->n_unit; ++ui)
			gen_unit(g, ci, c, (size_t)ui, buf + base + c->unit[ui].byte0);
		base += c->nbytes;
	}
	return need;
}

/* Per-16-kbp-tile index in STREAM coordinates for K_TILE contigs (what a BAI linear index gives in
 * virtual-offset coordinates): off[t] = byte offset, within the stream of contigs [c0,c1), of the first
 * record STARTING in tile t of contig ci.  Returns the number of tiles written (<= cap). */
uint64_t bamgen_tile_offsets(bamgen_t *g, int c0, int ci, uint64_t *off, uint64_t cap)
{
	uint64_t base = 0;
	for (int k = c0; k < ci; ++k) { size_contig(g, k); base += g->ctg[k].nbytes; }
	size_contig(g, ci);
	contig_t *c = &g->ctg[ci];
	if (c->kind != K_TILE) return 0;
	uint64_t n = c->n_unit < cap ? c->n_unit : cap;
	for (uint64_t t = 0; t < n; ++t) off[t] = base + c->unit[t].byte0;
	return n;
}

uint64_t bamgen_header(const bamgen_t *g, uint8_t *buf, uint64_t cap)
{
	char text[8192]; size_t tl = 0;
	tl += (size_t)snprintf(text + tl, sizeof text - tl, "@HD\tVN:1.0\tSO:coordinate\n");
	for (int i = 0; i < g->n_ref; ++i) tl += (size_t)snprintf(text + tl, sizeof text - tl, "@SQ\tSN:%s\tLN:%u\n", g->ctg[i].name, g->ctg[i].len);
	uint64_t need = 4 + 4 + tl + 4;
	for (int i = 0; i < g->n_ref; ++i) need += 4 + strlen(g->ctg[i].name) + 1 + 4;
	if (!buf || cap < need) return need;
	uint8_t *p = buf; uint32_t v;
	memcpy(p, "BAM\1", 4); p += 4;
	v = (uint32_t)tl; memcpy(p, &v, 4); p += 4; memcpy(p, text, tl); p += tl;
	v = (uint32_t)g->n_ref; memcpy(p, &v, 4); p += 4;
	for (int i = 0; i < g->n_ref; ++i) {
		v = (uint32_t)strlen(g->ctg[i].name) + 1; memcpy(p, &v, 4); p += 4;
		memcpy(p, g->ctg[i].name, v); p += v;
		memcpy(p, &g->ctg[i].len, 4); p += 4;
	}
	return need;
}

/* ------------------------------------------------------------------ BGZF + BAI writer */
#define BGZF_IN 0xff00u
static size_t bgzf_block(const uint8_t *in, size_t n, uint8_t *out, int level)
{
	static const uint8_t hdr[16] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0 };
	z_stream zs; memset(&zs, 0, sizeof zs);
	deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
	zs.next_in = (Bytef *)in; zs.avail_in = (uInt)n;
	zs.next_out = out + 18; zs.avail_out = 65536 - 26;
	int zr = deflate(&zs, Z_FINISH);
	size_t clen = zs.total_out;
	deflateEnd(&zs);
	if (zr != Z_STREAM_END) { fprintf(stderr, "bamgen: deflate overflow\n"); exit(1); }
	memcpy(out, hdr, 16);
	uint16_t bs = (uint16_t)(clen + 25); memcpy(out + 16, &bs, 2);
	uint32_t crc = (uint32_t)crc32(crc32(0, NULL, 0), in, (uInt)n), isz = (uint32_t)n;
	memcpy(out + 18 + clen, &crc, 4); memcpy(out + 22 + clen, &isz, 4);
	return clen + 26;
}

typedef struct { uint32_t bin; uint64_t beg, end; } chunk_t;
typedef struct { chunk_t *c; size_t n, cap; uint64_t *lin; size_t nlin, caplin; uint64_t off_beg, off_end, n_mapped, n_unmapped; int seen; } refidx_t;
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

/* Writes path (BAM) and path.bai.  The whole record stream is generated in memory contig by contig. */
int bamgen_write_bam(bamgen_t *g, const char *path, int level)
{
	FILE *f = fopen(path, "wb");
	if (!f) return -1;
	uint64_t hl = bamgen_header(g, NULL, 0), total = hl + bamgen_records_size(g, 0, g->n_ref + 1);
	uint8_t *raw = (uint8_t *)malloc(total + 1);
	if (!raw) { fclose(f); return -2; }
	bamgen_header(g, raw, hl);
	bamgen_records(g, 0, g->n_ref + 1, raw + hl, total - hl);
	size_t nblk = (size_t)((total + BGZF_IN - 1) / BGZF_IN);
	uint64_t *coff = (uint64_t *)calloc(nblk + 1, sizeof(uint64_t));
	uint8_t *comp = (uint8_t *)malloc((size_t)65536 * 4096);
	uint32_t *clen = (uint32_t *)calloc(4096, sizeof(uint32_t));
	uint64_t fpos = 0;
	for (size_t b0 = 0; b0 < nblk; b0 += 4096) {                 /* compress 4096 blocks at a time */
		size_t nb = nblk - b0 < 4096 ? nblk - b0 : 4096;
		#pragma omp parallel for schedule(dynamic, 8)
		for (long i = 0; i < (long)nb; ++i) {
			uint64_t s = (b0 + (size_t)i) * (uint64_t)BGZF_IN, n = total - s < BGZF_IN ? total - s : BGZF_IN;
			clen[i] = (uint32_t)bgzf_block(raw + s, (size_t)n, comp + (size_t)i * 65536, level);
		}
		for (size_t i = 0; i < nb; ++i) { coff[b0 + i] = fpos; fwrite(comp + i * 65536, 1, clen[i], f); fpos += clen[i]; }
	}
	coff[nblk] = fpos;
	static const uint8_t eof_blk[28] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
	fwrite(eof_blk, 1, 28, f);
	fclose(f);
	free(comp); free(clen);
#define VOFF(s) ((s) >= total ? coff[nblk] << 16 : ((coff[(s) / BGZF_IN] << 16) | ((s) % BGZF_IN)))
	/* ---- index, restating bam_index_core (samtools bam_index.c:151-243) over the in-memory stream */
	refidx_t *R = (refidx_t *)calloc((size_t)g->n_ref + 1, sizeof(refidx_t));
	uint64_t s = hl, n_no_coor = 0;
	uint32_t last_bin = 0xffffffffu, save_bin = 0xffffffffu;
	int32_t last_tid = -1, save_tid = -1; int first = 1;
	uint64_t save_off = VOFF(s), last_off = VOFF(s), off_beg = VOFF(s), n_mapped = 0, n_unmapped = 0;
	int stopped = 0;
	while (s < total) {
		int32_t bsz, tid, pos; memcpy(&bsz, raw + s, 4); memcpy(&tid, raw + s + 4, 4); memcpy(&pos, raw + s + 8, 4);
		uint32_t bin = raw[s + 14] | raw[s + 15] << 8, n_cigar = raw[s + 16] | raw[s + 17] << 8, flag = raw[s + 18] | raw[s + 19] << 8, l_qname = raw[s + 12];
		uint64_t next = s + 4 + (uint64_t)bsz;
		if (tid < 0) ++n_no_coor;
		if (stopped) { s = next; continue; }
		if (last_tid < tid || (last_tid >= 0 && tid < 0)) { last_tid = tid; last_bin = 0xffffffffu; }
		if (tid >= 0 && !(flag & 4)) {
			uint32_t end = (uint32_t)pos;
			for (uint32_t k = 0; k < n_cigar; ++k) { uint32_t w; memcpy(&w, raw + s + 36 + l_qname + 4 * k, 4); if ((0x3C1A7 >> ((w & 0xf) << 1)) & 2) end += w >> 4; }
			refidx_t *r = &R[tid];
			size_t b = (uint32_t)pos >> 14, e = (end - 1) >> 14;
			if (r->caplin < e + 1) { size_t oc = r->caplin; r->caplin = (e + 1) * 2; r->lin = (uint64_t *)realloc(r->lin, r->caplin * 8); memset(r->lin + oc, 0, (r->caplin - oc) * 8); }
			for (size_t i = b; i <= e; ++i) if (r->lin[i] == 0) r->lin[i] = last_off;
			r->nlin = e + 1;
		}
		if (bin != last_bin) {
			if (save_bin != 0xffffffffu) idx_chunk(&R[save_tid], save_bin, save_off, last_off);
			if (last_bin == 0xffffffffu && !first) {
				idx_chunk(&R[save_tid], 37450, off_beg, last_off);
				idx_chunk(&R[save_tid], 37450, n_mapped, n_unmapped);
				n_mapped = n_unmapped = 0; off_beg = last_off;
			}
			save_off = last_off; save_bin = last_bin = bin; save_tid = tid; first = 0;
			if (save_tid < 0) { stopped = 1; s = next; continue; }
		}
		if (flag & 4) ++n_unmapped; else ++n_mapped;
		last_off = VOFF(next);
		s = next;
	}
	if (save_tid >= 0) {
		/* samtools takes bam_tell() after the read that hit end-of-file, i.e. past the 28-byte EOF block */
		const uint64_t vend = (coff[nblk] + 28) << 16;
		idx_chunk(&R[save_tid], save_bin, save_off, vend);
		idx_chunk(&R[save_tid], 37450, off_beg, vend);
		idx_chunk(&R[save_tid], 37450, n_mapped, n_unmapped);
	}
	if (stopped) n_no_coor = n_no_coor;   /* counted above for every tid<0 record */
	char *ip = (char *)malloc(strlen(path) + 8); sprintf(ip, "%s.bai", path);
	FILE *fi = fopen(ip, "wb"); free(ip);
	if (!fi) { free(raw); free(coff); return -3; }
	fwrite("BAI\1", 1, 4, fi);
	int32_t n32 = g->n_ref; fwrite(&n32, 4, 1, fi);
	for (int t = 0; t < g->n_ref; ++t) {
		refidx_t *r = &R[t];
		/* chunks were appended in file order: stable-sort by bin keeps per-bin file order; the meta bin's
		 * two pseudo-chunks must keep their order, so sort on (bin) only for non-meta and append meta last */
		chunk_t meta[2]; int nmeta = 0; size_t w = 0;
		for (size_t i = 0; i < r->n; ++i) { if (r->c[i].bin == 37450) { if (nmeta < 2) meta[nmeta++] = r->c[i]; } else r->c[w++] = r->c[i]; }
		qsort(r->c, w, sizeof(chunk_t), cmp_chunk);
		/* merge_chunks (bam_index.c:113-139): adjacent chunks of a bin in the same BGZF block coalesce */
		size_t m = 0;
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
			int32_t nch = (int32_t)(j - i);
			fwrite(&r->c[i].bin, 4, 1, fi); fwrite(&nch, 4, 1, fi);
			for (size_t k = i; k < j; ++k) { fwrite(&r->c[k].beg, 8, 1, fi); fwrite(&r->c[k].end, 8, 1, fi); }
			i = j;
		}
		if (nmeta) {
			uint32_t mb = 37450; int32_t two = nmeta;
			fwrite(&mb, 4, 1, fi); fwrite(&two, 4, 1, fi);
			for (int k = 0; k < nmeta; ++k) { fwrite(&meta[k].beg, 8, 1, fi); fwrite(&meta[k].end, 8, 1, fi); }
		}
		for (size_t i = 1; i < r->nlin; ++i) if (r->lin[i] == 0) r->lin[i] = r->lin[i - 1];   /* fill_missing */
		int32_t nl = (int32_t)r->nlin; fwrite(&nl, 4, 1, fi);
		if (nl) fwrite(r->lin, 8, (size_t)nl, fi);
		free(r->c); free(r->lin);
	}
	fwrite(&n_no_coor, 8, 1, fi);
	fclose(fi);
	free(R); free(raw); free(coff);
	return 0;
}

#ifdef BAMGEN_MAIN
int main(int argc, char **argv)
{
	int config = 1, seq = -1, level = 1; double scale = 1.0; uint64_t seed = 0; const char *out = NULL;
	for (int a = 1; a < argc; ++a) {
		if (!strcmp(argv[a], "--config") && a + 1 < argc) config = atoi(argv[++a]);
		else if (!strcmp(argv[a], "--scale") && a + 1 < argc) scale = atof(argv[++a]);
		else if (!strcmp(argv[a], "--seed") && a + 1 < argc) seed = strtoull(argv[++a], 0, 10);
		else if (!strcmp(argv[a], "--seq") && a + 1 < argc) seq = !strcmp(argv[++a], "full");
		else if (!strcmp(argv[a], "--level") && a + 1 < argc) level = atoi(argv[++a]);
		else if (!strcmp(argv[a], "-o") && a + 1 < argc) out = argv[++a];
		else { fprintf(stderr, "usage: bamgen --config 1..5 [--scale S] [--seed N] [--seq none|full] [--level Z] -o out.bam\n"); return 1; }
	}
	if (!out) { fprintf(stderr, "bamgen: -o required\n"); return 1; }
	if (seq < 0) seq = (config == 1);          /* config 1 carries SEQ/QUAL like the survey probe */
	bamgen_t *g = bamgen_create(config, seed, scale, seq);
	if (!g) { fprintf(stderr, "bamgen: bad config\n"); return 1; }
	int rc = bamgen_write_bam(g, out, level);
	fprintf(stderr, "bamgen: config %d scale %g: %llu reads, %d contigs -> %s (rc %d)\n", config, scale,
	        (unsigned long long)bamgen_n_reads(g), g->n_ref, out, rc);
	bamgen_destroy(g);
	return rc ? 1 : 0;
}
#endif
