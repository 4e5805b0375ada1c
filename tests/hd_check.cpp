/* tests/hd_check.cpp -- CPU-only check of the record-level logic the kernels run (csrc/ngscov_hd.h).
 *
 * TEST INFRASTRUCTURE.  Compiles the NGS_HD functions with g++ and drives them with a sequential model
 * of the kernel orchestration (speculative sub-chunk walk, running-max verification, zipper repair,
 * prefix scan, run emission, window bins, decimal formatting), then compares every output byte with
 * the oracle (oracle/ngscov_oracle.c).  It lets the semantics be validated on a machine without a GPU;
 * the GPU tests then only have to prove that the parallel orchestration matches this model.
 *
 *   hd_check <bam> depth|wig <window> <subchunk_bytes> <chunk_bytes>
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include "../highperformancengs_b200/csrc/ngscov_hd.h"
#include "../include/ngscov.h"
#include "../oracle/ngscov_oracle.h"

struct HostMem {
	const uint8_t *base; int64_t n;
	int64_t size() const { return n; }
	uint32_t u32(int64_t o) const { uint32_t v; memcpy(&v, base + o, 4); return v; }
	uint32_t u8(int64_t o) const { return base[o]; }
	void hdr(int64_t o, ngs_hdr &h) const
	{
		h.block_size = (int32_t)u32(o); h.ref_id = (int32_t)u32(o + 4); h.pos = (int32_t)u32(o + 8);
		uint32_t a = u32(o + 12), b = u32(o + 16);
		h.l_qname = a & 0xff; h.n_cigar = b & 0xffff; h.flag = b >> 16;
		h.l_seq = (int32_t)u32(o + 20); h.next_ref_id = (int32_t)u32(o + 24); h.next_pos = (int32_t)u32(o + 28);
	}
	void hdr_lite(int64_t o, ngs_hdr &h) const { hdr(o, h); h.next_ref_id = h.next_pos = 0; }
	int64_t next_candidate(int64_t o, int64_t end, uint32_t n_ref) const
	{
		for (; o < end; ++o) if (u32(o) - 32u <= NGS_MAX_BS - 32u && u32(o + 4) + 1u <= n_ref) return o;
		return NGS_NONE;
	}
};
struct HostSink {
	std::vector<int> *diff; int *range_err;
	void add2(uint64_t a, uint64_t b, int sign) { (*diff)[a] += sign; (*diff)[b] -= sign; }
	void flag_range_error() { *range_err = 1; }
};
static long long exit_num(long long x) { return x == NGS_EXIT_INVALID ? 0 : x; }

/* one bedGraph line exactly as k_format_bed composes it: name piece, two cached-H position pieces, depth piece */
static uint32_t model_line(ngs_lw<uint32_t *> &L, ngs_hcache &hc, uint32_t n_lo, uint32_t n_hi, uint32_t nl, uint32_t s, uint32_t e, uint32_t d)
{
	const uint32_t k0 = L.k; uint32_t *w0 = L.w;
	L.bed_line_cached(hc, n_lo, n_hi, nl, s, ngs_ascii4(s % 10000u), e, ngs_ascii4(e % 10000u), d);
	return (uint32_t)((L.w - w0) * 4 + L.k - k0);
}

/* known-answer check of the register line assembly against snprintf: every digit count (1..9 for positions,
 * 1..9 for depths), every name length 1..8, every stream alignment */
static int selftest_lines()
{
	unsigned long long x = 88172645463325252ull;
	auto rnd = [&x]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
	static const uint32_t p10[10] = { 1, 10, 100, 1000, 10000, 100000, 1000000, 10000000, 100000000, 1000000000 };
	for (int it = 0; it < 200000; ++it) {
		const uint32_t nl = 1 + (uint32_t)(rnd() % 8);
		char name[16]; for (uint32_t k = 0; k + 1 < nl; ++k) name[k] = (char)('A' + rnd() % 26); name[nl - 1] = '\t'; name[nl] = 0;
		unsigned long long n8 = 0; for (uint32_t k = 0; k < nl; ++k) n8 |= (unsigned long long)(unsigned char)name[k] << (8 * k);
		const int n_lines = 1 + (int)(rnd() % 5);
		std::vector<uint32_t> words(64, 0xA5A5A5A5u);
		const uint32_t mis = (uint32_t)(rnd() % 16);
		std::string exp;
		ngs_lw<uint32_t *> L; L.begin(words.data(), mis);
		ngs_hcache hc; memset(&hc, 0, sizeof hc); hc.H = 0xffffffffu;
		for (int l = 0; l < n_lines; ++l) {
			uint32_t v[3];
			for (int k = 0; k < 3; ++k) { const int dg = 1 + (int)(rnd() % 9); v[k] = dg == 1 ? (uint32_t)(rnd() % 10) : p10[dg - 1] + (uint32_t)(rnd() % (p10[dg] - p10[dg - 1])); if (rnd() % 16 == 0) v[k] = p10[dg] - 1; }
			if (v[2] == 0) v[2] = 1;
			if (rnd() % 2) { v[1] = v[0] + (uint32_t)(rnd() % 300); if (v[1] >= 1000000000u) v[1] = v[0]; if (rnd() % 2) v[2] = 1 + (uint32_t)(rnd() % 9999); }
			ngs_num a, b, c; ngs_num_conv(v[0], a); ngs_num_conv(v[1], b); ngs_num_conv(v[2], c);
			char tmp[96]; const int len = snprintf(tmp, sizeof tmp, "%s%u\t%u\t%u\n", name, v[0], v[1], v[2]);
			if ((int)(nl + (a.tn & 0xff) + (b.tn & 0xff) + (c.tn & 0xff) + 3) != len) { fprintf(stderr, "selftest: length of %s", tmp); return 1; }
			exp.append(tmp, len);
			if (it & 1) L.bed_line((uint32_t)n8, (uint32_t)(n8 >> 32), nl, a, b, c);
			else if ((int)model_line(L, hc, (uint32_t)n8, (uint32_t)(n8 >> 32), nl, v[0], v[1], v[2]) != len) { fprintf(stderr, "selftest: model length of %s", tmp); return 1; }
		}
		if (L.k) { words[(mis + exp.size()) >> 2] &= 0; *L.w |= L.carry; }
		if (memcmp((const char *)words.data() + mis, exp.data(), exp.size())) { fprintf(stderr, "selftest: bytes differ for %s", exp.c_str()); return 1; }
	}
	/* flag values: 10 digits */
	ngs_num f; ngs_num_conv(4000000000u, f); if ((f.tn & 0xff) != 10) return 1;
	ngs_num_conv(0, f); if (f.tn != 1 || f.lo != '0' || f.hi != 0) return 1;
	printf("hd_check selftest-lines ok\n");
	return 0;
}

int main(int argc, char **argv)
{
	if (argc >= 2 && !strcmp(argv[1], "--selftest-lines")) return selftest_lines();
	if (argc < 6) { fprintf(stderr, "usage: hd_check bam depth|wig window subchunk chunk\n"); return 2; }
	const int mode = !strcmp(argv[2], "wig") ? 1 : 0;
	const uint32_t W = (uint32_t)atoi(argv[3]);
	int shift = 0; while ((1 << (shift + 1)) <= atoi(argv[4])) ++shift;
	const int64_t chunk = atoll(argv[5]);
	ora_bam bam;
	if (ngscov_oracle_load_bam(argv[1], &bam)) { fprintf(stderr, "cannot load %s\n", argv[1]); return 2; }
	ora_result ora;
	if (ngscov_oracle_run(bam.records, bam.records_len, bam.n_ref, bam.ref_len, (const char *const *)bam.ref_name, mode, W, 1, &ora)) { fprintf(stderr, "oracle failed\n"); return 2; }

	/* geometry (whole genome) */
	std::vector<ngs_seg> seg; std::vector<int32_t> c2s(bam.n_ref); std::vector<uint64_t> wb(bam.n_ref + 1, 0);
	uint64_t elems = 0;
	for (int i = 0; i < bam.n_ref; ++i) {
		ngs_seg s; memset(&s, 0, sizeof s);
		s.contig = i; s.clip_lo = 0; s.ref_len = bam.ref_len[i]; s.n_win = bam.ref_len[i] / W + 1; s.hard_end = 1;
		s.lim = (uint32_t)std::min<uint64_t>((uint64_t)bam.ref_len[i] + NGSCOV_OVERHANG_PAD, NGS_POS_MASK);
		s.base = elems; s.bins0 = wb[i]; wb[i + 1] = wb[i] + s.n_win;
		elems += (uint64_t)s.lim + 1;
		c2s[i] = i; seg.push_back(s);
	}
	std::vector<int> diff(elems, 0);
	int range_err = 0;
	ngs_walk_cfg cfg = { bam.n_ref, mode ? 0x4u : 0x704u, c2s.data(), seg.data(), -2, 0 };
	ngs_counts cnt = { 0, 0, 0, 0 };
	long long mismatches = 0, walks = 0;

	/* streamed pushes with carry, exactly as the library does it */
	std::vector<uint8_t> buf(NGS_CARRY_MAX + chunk + 64);
	uint8_t *data = buf.data() + NGS_CARRY_MAX;
	long long entry_next = 0; std::vector<uint8_t> carry;
	for (int64_t off = 0; off < (int64_t)bam.records_len || off == 0; off += chunk) {
		const int64_t n = std::min<int64_t>(chunk, (int64_t)bam.records_len - off);
		if (n <= 0) break;
		memcpy(data, bam.records + off, n);
		memcpy(data - carry.size(), carry.data(), carry.size());
		HostMem mem = { data, n };
		HostSink sink = { &diff, &range_err };
		const int64_t n_sub = (n + (1ll << shift) - 1) >> shift;
		std::vector<long long> entry(n_sub), ex(n_sub, 0);
		for (int64_t c = 0; c < n_sub; ++c) {
			const int64_t lo = c << shift, hi = std::min<int64_t>(lo + (1ll << shift), n);
			entry[c] = c == 0 ? (entry_next < hi ? entry_next : NGS_NONE) : ngs_find_entry(cfg, mem, bam.ref_len, lo, hi);
			if (entry[c] != NGS_NONE) { ex[c] = ngs_walk(cfg, mem, entry[c], hi, sink, +1, cnt); ++walks; }
		}
		/* verify + zipper repair */
		for (;;) {
			long long run = entry_next, bad = -1, expected = 0; int kind = 0;
			for (int64_t c = 0; c < n_sub; ++c) {
				const int64_t lo = c << shift, hi = std::min<int64_t>(lo + (1ll << shift), n);
				const bool exp_in = run < hi;
				int k = -1;
				if (exp_in && c > 0 && run < lo) k = 1;
				else if (exp_in ? entry[c] != run : entry[c] != NGS_NONE) k = 0;
				else if (entry[c] != NGS_NONE && ex[c] == NGS_EXIT_INVALID) k = 1;
				if (k >= 0) { bad = c; kind = k; expected = run; break; }
				run = std::max(run, exit_num(ex[c]));
			}
			if (bad < 0) { entry_next = run; break; }
			if (kind == 1) { fprintf(stderr, "hd_check: broken chain at sub-chunk %lld (chunk off %lld n %lld expected %lld entry %lld exit %lld e0 %lld)\n", bad, (long long)off, (long long)n, expected, entry[bad], ex[bad], entry_next); for (int64_t c=0;c<std::min<int64_t>(n_sub,6);++c) fprintf(stderr,"  c=%lld entry=%lld exit=%lld\n",(long long)c,entry[c],ex[c]); return 1; }
			++mismatches;
			ngs_counts sub = { 0, 0, 0, 0 };
			long long c = bad, cur = expected;
			for (;;) {
				long long cc = cur >= n ? n_sub : (cur < 0 ? 0 : cur >> shift);
				for (long long k = c; k < cc; ++k) if (entry[k] != NGS_NONE) {
					ngs_walk(cfg, mem, entry[k], std::min<int64_t>((k + 1) << shift, n), sink, -1, sub); entry[k] = NGS_NONE; ex[k] = 0; }
				if (cc >= n_sub || entry[cc] == cur) break;
				const int64_t hi = std::min<int64_t>((cc + 1) << shift, n);
				if (entry[cc] != NGS_NONE) ngs_walk(cfg, mem, entry[cc], hi, sink, -1, sub);
				entry[cc] = cur; ex[cc] = ngs_walk(cfg, mem, cur, hi, sink, +1, cnt);
				c = cc + 1; cur = (ex[cc] == NGS_EXIT_INVALID || ex[cc] >= NGS_EXIT_INF) ? n : ex[cc];
			}
			cnt.records -= sub.records; cnt.kept -= sub.kept; cnt.m_ops -= sub.m_ops; cnt.cigar_words -= sub.cigar_words;
		}
		/* finalize */
		carry.clear();
		if (entry_next >= NGS_EXIT_INF) {
			const long long o = entry_next - NGS_EXIT_INF - (1ll << 40);
			carry.assign(data + o, data + n);
			entry_next = -(long long)carry.size();
		} else entry_next -= n;
	}
	if (entry_next != 0 || range_err) { fprintf(stderr, "hd_check: truncated (%lld) or range error (%d)\n", entry_next, range_err); return 1; }

	/* scan + runs + bins + text */
	std::string bed, dep, wig, chs, bed_fast;
	struct Rec { uint32_t s, e; int d; };
	char tmp[512];
	std::vector<uint64_t> bins(wb[bam.n_ref], 0);
	for (int i = 0; i < bam.n_ref; ++i) {
		const ngs_seg &s = seg[i];
		const std::string nm = bam.ref_name[i];
		int depth = 0; uint32_t last = 0;
		uint64_t *b = bins.data() + s.bins0;
		auto add = [b](uint32_t k, int64_t v) { b[k] += (uint64_t)v; };
		std::vector<Rec> recs;                       /* boundary records of the contig, gaps (depth 0) included, as K3+K4 emits them */
		for (uint64_t p = 0; p <= s.lim; ++p) {
			const int d = diff[s.base + p];
			if (!d) continue;
			recs.push_back(Rec{ last, (uint32_t)p, depth });
			if (depth != 0) {
				char *q = tmp; memcpy(q, nm.data(), nm.size()); q += nm.size(); *q++ = '\t';
				q += ngs_put_u32(q, last); *q++ = '\t'; q += ngs_put_u32(q, (uint32_t)p); *q++ = '\t'; q += ngs_put_i32(q, depth); *q++ = '\n';
				bed.append(tmp, q - tmp);
				ngs_bin_run(mode, last, (uint32_t)p, depth, W, s.ref_len, s.n_win, add);
			}
			depth += d; last = (uint32_t)p;
		}
		if (depth != 0) { fprintf(stderr, "hd_check: depth does not return to zero on %s\n", nm.c_str()); return 1; }
		/* model of k_format_bed: tiles of 256 threads x 4 consecutive records, lines assembled by ngs_lw as whole
		 * words at a varying 16-byte misalignment, tails OR-ed in after the (modelled) barrier */
		if (nm.size() + 1 <= 8) {
			uint32_t n_lo = 0, n_hi = 0;
			{ std::string nt = nm + "\t"; unsigned long long n8 = 0; for (size_t k = 0; k < nt.size(); ++k) n8 |= (unsigned long long)(unsigned char)nt[k] << (8 * k); n_lo = (uint32_t)n8; n_hi = (uint32_t)(n8 >> 32); }
			const uint32_t nl = (uint32_t)nm.size() + 1;
			for (size_t t0 = 0; t0 < recs.size(); t0 += 1024) {
				const uint32_t mis = (uint32_t)(bed_fast.size() & 15);
				std::vector<uint32_t> words(1024 * 48 / 4 + 16, 0xDEADBEEFu);   /* garbage: nothing may rely on zeroed staging */
				std::vector<uint32_t> off(256), tl(256);
				std::vector<ngs_lw<uint32_t *>> L(256);
				std::vector<ngs_hcache> hcs(256); for (auto &h : hcs) { memset(&h, 0, sizeof h); h.H = 0xffffffffu; }
				uint32_t total = 0;
				std::vector<ngs_num> cs(1024), ce(1024), cd(1024);
				for (int t = 0; t < 256; ++t) {
					off[t] = total; tl[t] = 0;
					for (int i = 0; i < 4; ++i) {
						const size_t r = t0 + t * 4 + i;
						if (r >= recs.size() || recs[r].d == 0) continue;
						if (recs[r].d < 0) { fprintf(stderr, "hd_check: negative depth in model\n"); return 1; }
						ngs_num_conv(recs[r].s, cs[t * 4 + i]); ngs_num_conv(recs[r].e, ce[t * 4 + i]); ngs_num_conv((uint32_t)recs[r].d, cd[t * 4 + i]);
						tl[t] += nl + (cs[t * 4 + i].tn & 0xff) + (ce[t * 4 + i].tn & 0xff) + (cd[t * 4 + i].tn & 0xff) + 3;
					}
					total += tl[t];
				}
				words[(mis + total) >> 2] = 0;
				for (int t = 0; t < 256; ++t) {              /* phase 1: plain word stores */
					L[t].begin(words.data(), mis + off[t]);
					for (int i = 0; i < 4; ++i) {
						const size_t r = t0 + t * 4 + i;
						if (r >= recs.size() || recs[r].d == 0) continue;
						model_line(L[t], hcs[t], n_lo, n_hi, nl, recs[r].s, recs[r].e, (uint32_t)recs[r].d);
					}
				}
				for (int t = 0; t < 256; ++t) if (tl[t] && L[t].k) *L[t].w |= L[t].carry;   /* phase 2 */
				bed_fast.append((const char *)words.data() + mis, total);
			}
		} else {
			for (const Rec &rc : recs) if (rc.d) {
				char *q = tmp; memcpy(q, nm.data(), nm.size()); q += nm.size(); *q++ = '\t';
				q += ngs_put_u32(q, rc.s); *q++ = '\t'; q += ngs_put_u32(q, rc.e); *q++ = '\t'; q += ngs_put_i32(q, rc.d); *q++ = '\n';
				bed_fast.append(tmp, q - tmp);
			}
		}
		wig += "variableStep chrom=" + nm + " span=" + std::to_string((int)W) + "\n";
		for (uint32_t k = 0; k < s.n_win; ++k) {
			const uint64_t h = ngs_hundredths((double)(long long)b[k] / (double)W);
			const uint64_t ws = (uint64_t)k * W, we = std::min<uint64_t>(ws + W, s.ref_len);
			char *q = tmp; memcpy(q, nm.data(), nm.size()); q += nm.size(); *q++ = '\t';
			q += ngs_put_u32(q, (uint32_t)ws); *q++ = '\t'; q += ngs_put_u32(q, (uint32_t)we); *q++ = '\t'; q += ngs_put_f2(q, h); *q++ = '\n';
			dep.append(tmp, q - tmp);
			if (b[k]) { q = tmp; q += ngs_put_u32(q, (uint32_t)ws); *q++ = '\t'; q += ngs_put_f2(q, h); *q++ = '\n'; wig.append(tmp, q - tmp); }
		}
		chs += nm + "\t" + std::to_string((int)s.ref_len) + "\n";
	}
	int fail = 0;
	auto cmp = [&](const char *what, const std::string &got, const char *ref, size_t len) {
		if (got.size() != len || memcmp(got.data(), ref, len)) { fprintf(stderr, "hd_check: %s differs (%zu vs %zu bytes)\n", what, got.size(), len); fail = 1; }
	};
	cmp("bedGraph (register line assembly)", bed_fast, bed.data(), bed.size());
	if (mode == 0) { cmp("bedGraph", bed, ora.bedgraph, ora.bedgraph_len); cmp("depth", dep, ora.depth, ora.depth_len); }
	cmp("wig", wig, ora.wig, ora.wig_len); cmp("chromSize", chs, ora.chromsize, ora.chromsize_len);
	if (cnt.records != ora.n_records || cnt.kept != ora.n_kept || cnt.m_ops != ora.n_mops) {
		fprintf(stderr, "hd_check: counters differ: records %llu/%llu kept %llu/%llu mops %llu/%llu\n",
		        (unsigned long long)cnt.records, (unsigned long long)ora.n_records, (unsigned long long)cnt.kept, (unsigned long long)ora.n_kept,
		        (unsigned long long)cnt.m_ops, (unsigned long long)ora.n_mops);
		fail = 1;
	}
	printf("hd_check %s mode=%d W=%u S=%d chunk=%lld: records=%llu walks=%lld repaired=%lld %s\n", argv[1], mode, W, 1 << shift,
	       (long long)chunk, (unsigned long long)cnt.records, walks, mismatches, fail ? "FAIL" : "ok");
	return fail;
}
