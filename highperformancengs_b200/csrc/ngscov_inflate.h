/* csrc/ngscov_inflate.h -- DEFLATE (RFC 1951) decode core for BGZF blocks, shared by the device kernel
 * (k_bgzf_inflate, one warp per BGZF block) and the CPU model test (tests/inflate_check.cpp).
 *
 * Replaces, on the device, what the reference does on the host through samtools-0.1.19 bgzf.c:214-238
 * (inflate_block -> zlib inflate with -15 window bits) -- SURVEY.md 8(f) rank 1.  A BGZF block is an independent
 * raw-deflate stream of at most 64 KiB of output (bgzf.c:38-40), so blocks decode in parallel and every back
 * reference stays inside the block's own output.
 *
 * Everything here is NGS_HD and sequential: one decoder = one instruction stream.  On the device the 32 lanes of a
 * warp run this stream in lock step on identical values (table loads are broadcasts) and only the byte copies of
 * stored blocks and matches are spread over the lanes.
 *
 * Not a CPU fallback: the library calls it from device code only.
 */
#ifndef NGSCOV_INFLATE_H
#define NGSCOV_INFLATE_H
#include "ngscov_hd.h"

#ifndef NGS_INF_LBITS
#define NGS_INF_LBITS 10                 /* primary lookup width, literal/length code */
#endif
#ifndef NGS_INF_DBITS
#define NGS_INF_DBITS 8                  /* primary lookup width, distance code */
#endif

/* Primary-table entry (16 bit): bits 3..0 code length (0: the code is longer than the primary width), 15..4 the
 * symbol.  Base value and extra-bit count of a length / distance symbol (RFC 1951 3.2.5) come from two 32-entry
 * tables the output policy hands out (lentry / dentry: value in bits 23..0, number of extra bits in 31..24) -- on the
 * device they sit in shared memory once per CTA.  Two-byte entries keep one decoder's tables at 3.5 KB: 48 decoders per SM
 * (k_bgzf_inflate), or two per warp (k_bgzf_inflate_ls). */
typedef struct {
	uint16_t lfast[1 << NGS_INF_LBITS];
	uint16_t dfast[1 << NGS_INF_DBITS];
	uint16_t lsym[288], dsym[32];        /* symbols in canonical order, for the bit-by-bit slow path */
	uint16_t lcount[16], dcount[16];     /* codes per length */
	uint32_t lslow[2], dslow[2];         /* slow path restart behind the primary width: first code, symbol index */
	uint8_t  lens[320];                  /* code lengths of the block header (literal/length then distance) */
} ngs_inf_tables;

/* Bit reader, LSB first.  The stream is read as ALIGNED 32-bit words starting at the word that holds its first byte
 * (`lead` = the bits of that word that lie before the stream): w0 holds the next unread bit at position bp, w1 and w2
 * are the two words after it.  A field of up to 32 bits is one funnel shift of (w0, w1); when bp passes 32 the words
 * move down and the word after w2 is fetched -- two word-advances before anything needs it, so the load latency is
 * off the decode chain.  The reader may touch up to 15 bytes past the end of the stream and up to 3 before its start
 * (the caller guarantees that much readable memory); what it reads there only matters to a malformed stream, which
 * the consumed-bytes check at the end of every block rejects. */
typedef struct {
	const uint32_t *wbase;               /* aligned word holding the first byte read by this reader */
	uint32_t w0, w1, w2, bp, nw;         /* nw: words left behind so far */
	uint32_t lead, start;                /* start: stream offset (bytes) this reader was opened at */
} ngs_inf_bits;

NGS_HD uint32_t ngs_inf_word(const uint32_t *p)
{
#if defined(__CUDA_ARCH__)
	return __ldg(p);
#else
	uint32_t v; memcpy(&v, p, 4); return v;
#endif
}
NGS_HD void ngs_inf_open(ngs_inf_bits &b, const uint8_t *in, uint32_t start)
{
	const uintptr_t a = (uintptr_t)(in + start);
	b.wbase = (const uint32_t *)(a & ~(uintptr_t)3);
	b.lead = (uint32_t)(a & 3u) * 8u; b.bp = b.lead; b.nw = 0; b.start = start;
	b.w0 = ngs_inf_word(b.wbase); b.w1 = ngs_inf_word(b.wbase + 1); b.w2 = ngs_inf_word(b.wbase + 2);
}
/* the next 32 bits of the stream */
NGS_HD uint32_t ngs_inf_peek(const ngs_inf_bits &b)
{
#if defined(__CUDA_ARCH__)
	return __funnelshift_r(b.w0, b.w1, b.bp);
#else
	return b.bp ? (b.w0 >> b.bp) | (b.w1 << (32u - b.bp)) : b.w0;
#endif
}
NGS_HD void ngs_inf_skip(ngs_inf_bits &b, uint32_t n)        /* n <= 32 */
{
	b.bp += n;
	if (b.bp >= 32u) { b.bp -= 32u; ++b.nw; b.w0 = b.w1; b.w1 = b.w2; b.w2 = ngs_inf_word(b.wbase + b.nw + 2u); }
}
NGS_HD uint32_t ngs_inf_get(ngs_inf_bits &b, uint32_t n)     /* n <= 16 */
{
	const uint32_t v = ngs_inf_peek(b) & ((1u << n) - 1u);
	ngs_inf_skip(b, n);
	return v;
}
/* bytes consumed so far, counted from the start of the stream (a partly used byte counts) */
NGS_HD uint32_t ngs_inf_consumed(const ngs_inf_bits &b) { return b.start + ((b.nw * 32u + b.bp - b.lead + 7u) >> 3); }

/* length symbol 257 + i and distance symbol j -> base value | extra bits << 24 (RFC 1951 3.2.5).  Symbols that must
 * not occur (length 286/287, distance 30/31) get a value no block can satisfy, so the range check of the match
 * rejects them without a test of their own. */
NGS_HD uint32_t ngs_inf_len_entry(uint32_t i)
{
	if (i > 28u) return 0xffffffu;
	if (i < 8u) return 3u + i;
	if (i == 28u) return 258u;
	const uint32_t ext = (i >> 2) - 1u;
	return (3u + ((4u + (i & 3u)) << ext)) | (ext << 24);
}
NGS_HD uint32_t ngs_inf_dist_entry(uint32_t j)
{
	if (j > 29u) return 0xffffffu;
	if (j < 4u) return 1u + j;
	const uint32_t ext = (j >> 1) - 1u;
	return (1u + ((2u + (j & 1u)) << ext)) | (ext << 24);
}

/* canonical Huffman tables from code lengths: symbols in canonical order, codes per length, and the primary lookup
 * (indexed by the next `bits` stream bits = the code reversed).  Returns 0, or -1 for an over-subscribed code. */
NGS_HD int ngs_inf_build(const uint8_t *lens, uint32_t n, uint16_t *count, uint16_t *sym, uint16_t *fast, uint32_t bits, uint32_t *slow)
{
	uint16_t offs[16];
	for (uint32_t l = 0; l < 16u; ++l) count[l] = 0;
	for (uint32_t s = 0; s < n; ++s) count[lens[s]] = (uint16_t)(count[lens[s]] + 1);
	count[0] = 0;
	int left = 1;
	for (uint32_t l = 1; l < 16u; ++l) { left = (left << 1) - (int)count[l]; if (left < 0) return -1; }
	offs[1] = 0;
	for (uint32_t l = 1; l < 15u; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
	for (uint32_t s = 0; s < n; ++s) if (lens[s]) { sym[offs[lens[s]]] = (uint16_t)s; offs[lens[s]] = (uint16_t)(offs[lens[s]] + 1); }
	for (uint32_t i = 0; i < (1u << bits); ++i) fast[i] = 0;
	uint32_t code = 0, idx = 0;
	for (uint32_t l = 1; l <= bits; ++l) {
		code <<= 1;
		for (uint32_t k = 0; k < count[l]; ++k, ++code, ++idx) {
#if defined(__CUDA_ARCH__)
			const uint32_t r = __brev(code) >> (32u - l);
#else
			uint32_t r = 0;
			for (uint32_t j = 0; j < l; ++j) r |= ((code >> j) & 1u) << (l - 1u - j);
#endif
			const uint16_t e = (uint16_t)(l | ((uint32_t)sym[idx] << 4));
			for (uint32_t f = r; f < (1u << bits); f += 1u << l) fast[f] = e;
		}
	}
	/* state of the bit-by-bit decode after `bits` levels (it does not depend on the bits read): a long code restarts here */
	uint32_t first = 0, index = 0;
	for (uint32_t l = 1; l <= bits; ++l) { index += count[l]; first = (first + count[l]) << 1; }
	slow[0] = first; slow[1] = index;
	return 0;
}

/* code longer than the primary width, decoded bit by bit over the canonical order (RFC 1951 3.2.2) from the 32 peeked
 * bits x.  Returns the symbol and its code length in *clen (nothing is consumed), or -1. */
NGS_HD int ngs_inf_decode_slow(uint32_t x, const uint16_t *count, const uint16_t *sym, uint32_t bits, const uint32_t *slow, uint32_t *clen)
{
	/* the first `bits` stream bits are the code prefix, most significant bit first */
	const uint32_t pre = x & ((1u << bits) - 1u);
#if defined(__CUDA_ARCH__)
	int code = (int)((__brev(pre) >> (32u - bits)) << 1);
#else
	uint32_t r = 0;
	for (uint32_t j = 0; j < bits; ++j) r |= ((pre >> j) & 1u) << (bits - 1u - j);
	int code = (int)(r << 1);
#endif
	int first = (int)slow[0], index = (int)slow[1];
	uint32_t bb = x >> bits;
	for (uint32_t l = bits + 1u; l < 16u; ++l) {
		code |= (int)(bb & 1u); bb >>= 1;
		const int c = (int)count[l];
		if (code - c < first) { *clen = l; return (int)sym[index + (code - first)]; }
		index += c; first += c; first <<= 1; code <<= 1;
	}
	return -1;
}
/* one symbol of a plain table */
NGS_HD int ngs_inf_decode(ngs_inf_bits &b, const uint16_t *count, const uint16_t *sym, const uint16_t *fast, uint32_t bits, const uint32_t *slow)
{
	const uint32_t x = ngs_inf_peek(b);
	const uint32_t e = fast[x & ((1u << bits) - 1u)];
	if (e & 15u) { ngs_inf_skip(b, e & 15u); return (int)(e >> 4); }
	uint32_t cl = 0;
	const int s = ngs_inf_decode_slow(x, count, sym, bits, slow, &cl);
	if (s >= 0) ngs_inf_skip(b, cl);
	return s;
}

/* header of a dynamic block (RFC 1951 3.2.7) or the fixed code (3.2.6) -> tables.  0 / -1. */
NGS_HD int ngs_inf_block_tables(ngs_inf_bits &b, uint32_t btype, ngs_inf_tables &T)
{
	uint32_t nlen, ndist;
	if (btype == 1u) {
		for (uint32_t s = 0; s < 144u; ++s) T.lens[s] = 8;
		for (uint32_t s = 144u; s < 256u; ++s) T.lens[s] = 9;
		for (uint32_t s = 256u; s < 280u; ++s) T.lens[s] = 7;
		for (uint32_t s = 280u; s < 288u; ++s) T.lens[s] = 8;
		for (uint32_t s = 0; s < 30u; ++s) T.lens[288u + s] = 5;
		nlen = 288u; ndist = 30u;
	} else {
		nlen = ngs_inf_get(b, 5) + 257u; ndist = ngs_inf_get(b, 5) + 1u;
		const uint32_t ncode = ngs_inf_get(b, 4) + 4u;
		if (nlen > 286u || ndist > 30u) return -1;
		/* code-length code, permuted order 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15 packed five bits each */
		const uint64_t ord_lo = 16ull | (17ull << 5) | (18ull << 10) | (0ull << 15) | (8ull << 20) | (7ull << 25) | (9ull << 30) | (6ull << 35) | (10ull << 40) | (5ull << 45) | (11ull << 50) | (4ull << 55);
		const uint64_t ord_hi = 12ull | (3ull << 5) | (13ull << 10) | (2ull << 15) | (14ull << 20) | (1ull << 25) | (15ull << 30);
		for (uint32_t i = 0; i < 19u; ++i) T.lens[i] = 0;
		for (uint32_t i = 0; i < ncode; ++i) {
			const uint32_t o = i < 12u ? (uint32_t)(ord_lo >> (5u * i)) & 31u : (uint32_t)(ord_hi >> (5u * (i - 12u))) & 31u;
			T.lens[o] = (uint8_t)ngs_inf_get(b, 3);
		}
		/* the 19-symbol code borrows the distance tables for a moment */
		if (ngs_inf_build(T.lens, 19u, T.dcount, T.dsym, T.dfast, 7u, T.dslow)) return -1;
		uint32_t i = 0;
		while (i < nlen + ndist) {
			const int s = ngs_inf_decode(b, T.dcount, T.dsym, T.dfast, 7u, T.dslow);
			if (s < 0) return -1;
			if (s < 16) { T.lens[i++] = (uint8_t)s; continue; }
			uint32_t rep, val = 0;
			if (s == 16) { if (i == 0) return -1; val = T.lens[i - 1]; rep = 3u + ngs_inf_get(b, 2); }
			else if (s == 17) rep = 3u + ngs_inf_get(b, 3);
			else rep = 11u + ngs_inf_get(b, 7);
			if (i + rep > nlen + ndist) return -1;
			while (rep--) T.lens[i++] = (uint8_t)val;
		}
		if (T.lens[256] == 0) return -1;
	}
	/* the lengths sit at lens[0..nlen) and lens[nlen..nlen+ndist); the 19-symbol code that lived in the distance
	 * tables is dead by now */
	if (ngs_inf_build(T.lens + nlen, ndist, T.dcount, T.dsym, T.dfast, NGS_INF_DBITS, T.dslow)) return -1;
	if (ngs_inf_build(T.lens, nlen, T.lcount, T.lsym, T.lfast, NGS_INF_LBITS, T.lslow)) return -1;
	return 0;
}

/* Decoding loop with the output left to a policy:
 *   void literal(uint32_t pos, uint32_t byte); void match(uint32_t pos, uint32_t len, uint32_t dist);
 *   void stored(uint32_t pos, const uint8_t *src, uint32_t len); void finish();   (finish: nothing may stay pending)
 *   uint32_t lentry(uint32_t i); uint32_t dentry(uint32_t j);   ngs_inf_len_entry / ngs_inf_dist_entry, or a table of them
 *   int tables(ngs_inf_bits &b, uint32_t btype, ngs_inf_tables &T);   builds the block's tables, normally by calling
 *       ngs_inf_block_tables; the device policy lets ONE lane build them in the warp's shared-memory slice and hands the
 *       advanced bit reader to the other lanes
 * `in` must be readable from 3 bytes before it up to in_len + 15.  Returns the number of bytes produced, or -1 (malformed stream / output
 * larger than out_len). */
template <class Out>
NGS_HD long ngs_inflate_stream(const uint8_t *in, uint32_t in_len, uint32_t out_len, ngs_inf_tables &T, Out &out)
{
	ngs_inf_bits b; ngs_inf_open(b, in, 0u);
	uint32_t pos = 0, last, bad = 0;
	do {
		const uint32_t hdr = ngs_inf_get(b, 3);
		last = hdr & 1u;
		const uint32_t btype = hdr >> 1;
		if (btype == 0u) {
			ngs_inf_skip(b, (8u - (b.bp & 7u)) & 7u);            /* to the next byte boundary (lead is a multiple of 8) */
			const uint32_t len = ngs_inf_get(b, 16), nlen = ngs_inf_get(b, 16);
			if ((len ^ nlen) != 0xffffu) return -1;
			const uint32_t at = ngs_inf_consumed(b);
			if (at + len > in_len || pos + len > out_len) return -1;
			out.finish();
			out.stored(pos, in + at, len);
			pos += len;
			ngs_inf_open(b, in, at + len);
		} else if (btype == 3u) return -1;
		else {
			if (out.tables(b, btype, T)) return -1;
			for (;;) {
				/* literal/length: at most 15 code bits + 5 extra bits, all inside one 32-bit peek */
				uint32_t x = ngs_inf_peek(b);
				const uint32_t e = T.lfast[x & ((1u << NGS_INF_LBITS) - 1u)];
				uint32_t cl = e & 15u, sym = e >> 4;
				if (!cl) { const int s = ngs_inf_decode_slow(x, T.lcount, T.lsym, NGS_INF_LBITS, T.lslow, &cl); if (s < 0) { bad = 1; break; } sym = (uint32_t)s; }
				if (sym < 256u) {
					ngs_inf_skip(b, cl);
					if (pos >= out_len) { bad = 1; break; }
					out.literal(pos, sym); ++pos;
					continue;
				}
				if (sym == 256u) { ngs_inf_skip(b, cl); break; }
				const uint32_t le = out.lentry(sym - 257u);
				const uint32_t len = (le & 0xffffffu) + ((x >> cl) & ((1u << (le >> 24)) - 1u));
				ngs_inf_skip(b, cl + (le >> 24));
				/* distance: at most 15 code bits + 13 extra bits */
				x = ngs_inf_peek(b);
				const uint32_t d = T.dfast[x & ((1u << NGS_INF_DBITS) - 1u)];
				uint32_t cd = d & 15u, ds = d >> 4;
				if (!cd) { const int s = ngs_inf_decode_slow(x, T.dcount, T.dsym, NGS_INF_DBITS, T.dslow, &cd); if (s < 0) { bad = 1; break; } ds = (uint32_t)s; }
				const uint32_t de = out.dentry(ds);
				const uint32_t dist = (de & 0xffffffu) + ((x >> cd) & ((1u << (de >> 24)) - 1u));
				ngs_inf_skip(b, cd + (de >> 24));
				if (dist > pos || pos + len > out_len) { bad = 1; break; }      /* also: symbols that must not occur */
				out.match(pos, len, dist);
				pos += len;
			}
			if (bad) return -1;
		}
		if (ngs_inf_consumed(b) > in_len) return -1;               /* ran past the compressed data */
	} while (!last);
	out.finish();
	return (long)pos;
}

#endif
