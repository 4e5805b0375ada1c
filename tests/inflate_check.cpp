/* tests/inflate_check.cpp -- CPU-only check of the DEFLATE decode core the GPU inflate kernel runs
 * (csrc/ngscov_inflate.h) against zlib.  TEST INFRASTRUCTURE.
 *
 *   inflate_check                 synthetic streams: every block type (stored / fixed / dynamic), all zlib levels and
 *                                 strategies, record-like, random and degenerate inputs, truncated / corrupt streams
 *   inflate_check file.bam        every BGZF block of a real file against zlib's inflate of the same block
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <zlib.h>
#include "../highperformancengs_b200/csrc/ngscov_inflate.h"

struct SeqOut {
	uint8_t *o;
	void literal(uint32_t pos, uint32_t b) { o[pos] = (uint8_t)b; }
	void match(uint32_t pos, uint32_t len, uint32_t dist) { for (uint32_t i = 0; i < len; ++i) o[pos + i] = o[pos - dist + i]; }
	void stored(uint32_t pos, const uint8_t *src, uint32_t len) { memcpy(o + pos, src, len); }
	void finish() {}
	int tables(ngs_inf_bits &b, uint32_t btype, ngs_inf_tables &T) { return ngs_inf_block_tables(b, btype, T); }
	uint32_t lentry(uint32_t i) { return ngs_inf_len_entry(i); }
	uint32_t dentry(uint32_t j) { return ngs_inf_dist_entry(j); }
};
/* the lane-parallel copy rule of the kernel: byte i of a match comes from pos - dist + (i % dist), all old bytes */
struct ParOut {
	uint8_t *o;
	void literal(uint32_t pos, uint32_t b) { o[pos] = (uint8_t)b; }
	void match(uint32_t pos, uint32_t len, uint32_t dist)
	{
		std::vector<uint8_t> tmp(len);
		for (uint32_t i = 0; i < len; ++i) tmp[i] = o[pos - dist + (i % dist)];      /* all reads before any write */
		memcpy(o + pos, tmp.data(), len);
	}
	void stored(uint32_t pos, const uint8_t *src, uint32_t len) { memcpy(o + pos, src, len); }
	void finish() {}
	int tables(ngs_inf_bits &b, uint32_t btype, ngs_inf_tables &T) { return ngs_inf_block_tables(b, btype, T); }
	uint32_t lentry(uint32_t i) { return ngs_inf_len_entry(i); }
	uint32_t dentry(uint32_t j) { return ngs_inf_dist_entry(j); }
};

static std::vector<uint8_t> deflate_raw(const std::vector<uint8_t> &in, int level, int strategy)
{
	z_stream zs; memset(&zs, 0, sizeof zs);
	deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy);
	std::vector<uint8_t> out(deflateBound(&zs, in.size()) + 64);
	zs.next_in = (Bytef *)in.data(); zs.avail_in = (uInt)in.size(); zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
	if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { fprintf(stderr, "deflate failed\n"); exit(2); }
	out.resize(zs.total_out);
	deflateEnd(&zs);
	return out;
}

static int check_stream(const uint8_t *comp, size_t clen, const uint8_t *exp, size_t ulen, const char *what)
{
	static ngs_inf_tables T;
	std::vector<uint8_t> a(ulen + 8, 0xEE), b(ulen + 8, 0xEE);
	SeqOut so = { a.data() }; ParOut po = { b.data() };
	std::vector<uint8_t> padded(clen + 20, 0xA7);              /* the decoder may read up to 15 bytes past the stream */
	uint8_t *cp = padded.data() + (clen % 3);                  /* and must cope with any alignment */
	memcpy(cp, comp, clen);
	const long ra = ngs_inflate_stream(cp, (uint32_t)clen, (uint32_t)ulen, T, so);
	const long rb = ngs_inflate_stream(cp, (uint32_t)clen, (uint32_t)ulen, T, po);
	if (ra != (long)ulen || rb != (long)ulen || memcmp(a.data(), exp, ulen) || memcmp(b.data(), exp, ulen)) {
		fprintf(stderr, "inflate_check: %s: mismatch (got %ld / %ld bytes, expected %zu)\n", what, ra, rb, ulen);
		return 1;
	}
	return 0;
}

int main(int argc, char **argv)
{
	int fail = 0; size_t n_streams = 0;
	if (argc >= 2) {
		FILE *f = fopen(argv[1], "rb");
		if (!f) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
		std::vector<uint8_t> file; uint8_t buf[1 << 16]; size_t n;
		while ((n = fread(buf, 1, sizeof buf, f)) > 0) file.insert(file.end(), buf, buf + n);
		fclose(f);
		size_t off = 0;
		std::vector<uint8_t> exp(65536);
		while (off + 18 <= file.size()) {
			const uint8_t *p = file.data() + off;
			const size_t bsize = (size_t)(p[16] | p[17] << 8) + 1;
			uint32_t isize; memcpy(&isize, p + bsize - 4, 4);
			z_stream zs; memset(&zs, 0, sizeof zs); inflateInit2(&zs, -15);
			zs.next_in = (Bytef *)p + 18; zs.avail_in = (uInt)(bsize - 26); zs.next_out = exp.data(); zs.avail_out = 65536;
			if (inflate(&zs, Z_FINISH) != Z_STREAM_END || zs.total_out != isize) { fprintf(stderr, "zlib failed at %zu\n", off); return 2; }
			inflateEnd(&zs);
			char what[64]; snprintf(what, sizeof what, "block at %zu", off);
			fail |= check_stream(p + 18, bsize - 26, exp.data(), isize, what);
			++n_streams; off += bsize;
		}
	} else {
		unsigned long long x = 0x9E3779B97F4A7C15ull;
		auto rnd = [&x]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
		std::vector<std::vector<uint8_t>> inputs;
		inputs.push_back({});                                                   /* empty (BGZF EOF block) */
		inputs.push_back({ 'a' });
		inputs.push_back(std::vector<uint8_t>(65280, 0));                        /* one long run: dist 1, len 258 chains */
		{ std::vector<uint8_t> v(65280); for (auto &c : v) c = (uint8_t)rnd(); inputs.push_back(v); }           /* incompressible */
		{ std::vector<uint8_t> v(40000); for (size_t i = 0; i < v.size(); ++i) v[i] = (uint8_t)("ACGT"[rnd() & 3]); inputs.push_back(v); }
		{ std::vector<uint8_t> v(65280); for (size_t i = 0; i < v.size(); ++i) v[i] = (uint8_t)(i % 251 < 200 ? i % 51 : rnd()); inputs.push_back(v); }
		{	/* BAM-record-like: fixed-size records with slowly increasing positions and repeated fields */
			std::vector<uint8_t> v; uint32_t pos = 1000;
			while (v.size() + 64 < 65280) {
				uint32_t rec[9] = { 47, 0, pos, 0x49123c05u, 0x00000001u, 0, 0xffffffffu, 0xffffffffu, 0 };
				pos += (uint32_t)(rnd() % 7);
				const uint8_t *q = (const uint8_t *)rec; v.insert(v.end(), q, q + 36);
				char nm[8]; const int l = snprintf(nm, sizeof nm, "r%u", (unsigned)(rnd() % 100000)); v.insert(v.end(), nm, nm + l + 1);
				uint32_t cig = 150u << 4; q = (const uint8_t *)&cig; v.insert(v.end(), q, q + 4);
			}
			inputs.push_back(v);
		}
		for (int k = 0; k < 40; ++k) {                                          /* random mixtures of literals and copies */
			std::vector<uint8_t> v; const size_t want = 1 + rnd() % 65000;
			while (v.size() < want) {
				if (v.empty() || rnd() % 3 == 0) { const size_t m = 1 + rnd() % 40; for (size_t i = 0; i < m; ++i) v.push_back((uint8_t)(rnd() % (k % 2 ? 256 : 4))); }
				else { const size_t d = 1 + rnd() % (v.size() < 32768 ? v.size() : 32768), m = 3 + rnd() % 300; for (size_t i = 0; i < m; ++i) v.push_back(v[v.size() - d]); }
			}
			v.resize(want);
			inputs.push_back(v);
		}
		const int strategies[] = { Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE, Z_FILTERED };
		for (size_t i = 0; i < inputs.size(); ++i)
			for (int level = 0; level <= 9; level += (i < 8 ? 1 : 3))
				for (int st : strategies) {
					const std::vector<uint8_t> c = deflate_raw(inputs[i], level, st);
					char what[96]; snprintf(what, sizeof what, "input %zu level %d strategy %d", i, level, st);
					fail |= check_stream(c.data(), c.size(), inputs[i].data(), inputs[i].size(), what);
					++n_streams;
					if (i >= 3 && i < 8 && level == 6 && !c.empty()) {               /* malformed streams must fail cleanly, never run away */
						static ngs_inf_tables T; std::vector<uint8_t> o(inputs[i].size() + 8); SeqOut so = { o.data() };
						std::vector<uint8_t> t(c.begin(), c.begin() + c.size() / 2); t.resize(t.size() + 16, 0); t.resize(t.size() - 16);
						if (ngs_inflate_stream(t.data(), (uint32_t)t.size(), (uint32_t)inputs[i].size(), T, so) == (long)inputs[i].size() && inputs[i].size() > 64) { fprintf(stderr, "inflate_check: truncated stream accepted (%s)\n", what); fail = 1; }
						std::vector<uint8_t> c2 = c; c2.reserve(c2.size() + 16);
						if (ngs_inflate_stream(c2.data(), (uint32_t)c2.size(), (uint32_t)(inputs[i].size() / 2), T, so) >= 0 && inputs[i].size() > 2) { fprintf(stderr, "inflate_check: output overrun not detected (%s)\n", what); fail = 1; }
						for (int z = 0; z < 20; ++z) { std::vector<uint8_t> g = c; g[rnd() % g.size()] ^= (uint8_t)(1u << (rnd() % 8)); g.reserve(g.size() + 16); ngs_inflate_stream(g.data(), (uint32_t)g.size(), (uint32_t)inputs[i].size(), T, so); }
					}
				}
	}
	printf("inflate_check: %zu streams %s\n", n_streams, fail ? "FAIL" : "ok");
	return fail;
}
