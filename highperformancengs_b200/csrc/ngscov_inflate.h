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

#define NGS_INF_LBITS 10                 /* primary lookup width, literal/length code */
#define NGS_INF_DBITS 8                  /* primary lookup width, distance code */

/* Primary-table entry (32 bit): bits 3..0 code length (0: the code is longer than the primary width), 7..4 number of
 * extra bits, 9..8 kind (0 literal, 1 length or distance, 2 end of block), 31..16 value: the literal, or the BASE of
 * the length / distance (RFC 1951 3.2.5), so the hot loop does no arithmetic on symbols.  Plain tables (the
 * code-length code) carry the symbol in 31..16. */
#define NGS_INF_KIND_LIT 0u
#define NGS_INF_KIND_VAL 1u
#define NGS_INF_KIND_EOB 2u
#define NGS_INF_KIND_BAD 3u

/* table memory of one decoder (device: one slice of shared memory per warp) */
typedef struct {
	uint32_t lfast[1 << NGS_INF_LBITS];
	uint32_t dfast[1 << NGS_INF_DBITS];
	uint16_t lsym[288], dsym[32];        /* symbols in canonical order, for the bit-by-bit slow path */
	uint16_t lcount[16], dcount[16];     /* codes per length */
	uint32_t lslow[2], dslow[2];         /* slow path restart behind the primary width: first code, symbol index */
	uint8_t  lens[320];                  /* code lengths of the block header (literal/length then distance) */
} ngs_inf_tables;

typedef struct {
	const uint8_t *in; uint32_t in_len, in_pos;
	uint64_t buf; uint32_t cnt;          /* bit buffer, LSB first */
} ngs_inf_bits;

/* 32 more bits into the buffer (requires cnt <= 32).  Away from the end of the input the four bytes come from two
 * aligned 32-bit loads and a funnel shift (they may read up to seven bytes past in_len: the caller guarantees that
 * much readable memory); near the end byte by byte, zero bits past the end. */
NGS_HD void ngs_inf_refill32(ngs_inf_bits &b)
{
	uint32_t v = 0;
	if (b.in_pos + 8u <= b.in_len) {
		const uintptr_t p = (uintptr_t)(b.in + b.in_pos);
		const uint32_t *w = (const uint32_t *)(p & ~(uintptr_t)3);
		const uint32_t sh = (uint32_t)(p & 3u) * 8u, lo = w[0], hi = w[1];
#if defined(__CUDA_ARCH__)
		v = __funnelshift_r(lo, hi, sh);
#else
		v = sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
	} else {
		for (uint32_t k = 0; k < 4u; ++k) if (b.in_pos + k < b.in_len) v |= (uint32_t)b.in[b.in_pos + k] << (8u * k);
	}
	b.in_pos += 4u;                      /* may run past in_len: only zero bits come from there, and that is an error below */
	b.buf |= (uint64_t)v << b.cnt; b.cnt += 32u;
}
NGS_HD uint32_t ngs_inf_get(ngs_inf_bits &b, uint32_t n)     /* n <= 16 */
{
	if (b.cnt < n) ngs_inf_refill32(b);
	const uint32_t v = (uint32_t)b.buf & ((1u << n) - 1u);
	b.buf >>= n; b.cnt -= n;
	return v;
}
/* bytes really consumed so far (the buffer may hold look-ahead) */
NGS_HD uint32_t ngs_inf_consumed(const ngs_inf_bits &b) { return b.in_pos - (b.cnt >> 3); }

/* symbol -> entry bits above the code length */
NGS_HD uint32_t ngs_inf_ll_entry(uint32_t s)       /* literal/length alphabet */
{
	if (s < 256u) return (s << 16) | (NGS_INF_KIND_LIT << 8);
	if (s == 256u) return NGS_INF_KIND_EOB << 8;
	if (s > 285u) return NGS_INF_KIND_BAD << 8;
	const uint32_t i = s - 257u;
	uint32_t base, ext;
	if (i < 8u) { base = 3u + i; ext = 0; }
	else if (i >= 28u) { base = 258u; ext = 0; }
	else { ext = (i >> 2) - 1u; base = 3u + ((4u + (i & 3u)) << ext); }
	return (base << 16) | (NGS_INF_KIND_VAL << 8) | (ext << 4);
}
NGS_HD uint32_t ngs_inf_d_entry(uint32_t j)        /* distance alphabet */
{
	if (j > 29u) return NGS_INF_KIND_BAD << 8;
	uint32_t base, ext;
	if (j < 4u) { base = 1u + j; ext = 0; }
	else { ext = (j >> 1) - 1u; base = 1u + ((2u + (j & 1u)) << ext); }
	return (base << 16) | (NGS_INF_KIND_VAL << 8) | (ext << 4);
}

/* canonical Huffman tables from code lengths: symbols in canonical order, codes per length, and the primary lookup
 * (indexed by the next `bits` stream bits = the code reversed).  alphabet: 0 plain, 1 literal/length, 2 distance.
 * Returns 0, or -1 for an over-subscribed code. */
NGS_HD int ngs_inf_build(const uint8_t *lens, uint32_t n, uint16_t *count, uint16_t *sym, uint32_t *fast, uint32_t bits, int alphabet, uint32_t *slow)
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
			const uint32_t sy = sym[idx];
			const uint32_t e = l | (alphabet == 1 ? ngs_inf_ll_entry(sy) : alphabet == 2 ? ngs_inf_d_entry(sy) : (sy << 16));
			for (uint32_t f = r; f < (1u << bits); f += 1u << l) fast[f] = e;
		}
	}
	/* state of the bit-by-bit decode after `bits` levels (it does not depend on the bits read): a long code restarts here */
	uint32_t first = 0, index = 0;
	for (uint32_t l = 1; l <= bits; ++l) { index += count[l]; first = (first + count[l]) << 1; }
	slow[0] = first; slow[1] = index;
	return 0;
}

/* code longer than the primary width: bit by bit over the canonical order (RFC 1951 3.2.2); needs >= 15 bits in the
 * buffer.  Returns the symbol (bits consumed) or -1. */
NGS_HD int ngs_inf_decode_slow(ngs_inf_bits &b, const uint16_t *count, const uint16_t *sym, uint32_t bits, const uint32_t *slow)
{
	/* the first `bits` stream bits are the code prefix, most significant bit first */
	const uint32_t pre = (uint32_t)b.buf & ((1u << bits) - 1u);
#if defined(__CUDA_ARCH__)
	int code = (int)((__brev(pre) >> (32u - bits)) << 1);
#else
	uint32_t r = 0;
	for (uint32_t j = 0; j < bits; ++j) r |= ((pre >> j) & 1u) << (bits - 1u - j);
	int code = (int)(r << 1);
#endif
	int first = (int)slow[0], index = (int)slow[1];
	uint64_t bb = b.buf >> bits;
	for (uint32_t l = bits + 1u; l < 16u; ++l) {
		code |= (int)(bb & 1u); bb >>= 1;
		const int c = (int)count[l];
		if (code - c < first) { b.buf >>= l; b.cnt -= l; return (int)sym[index + (code - first)]; }
		index += c; first += c; first <<= 1; code <<= 1;
	}
	return -1;
}
/* one symbol of a plain table */
NGS_HD int ngs_inf_decode(ngs_inf_bits &b, const uint16_t *count, const uint16_t *sym, const uint32_t *fast, uint32_t bits, const uint32_t *slow)
{
	if (b.cnt < 16u) ngs_inf_refill32(b);
	const uint32_t e = fast[(uint32_t)b.buf & ((1u << bits) - 1u)];
	if (e & 15u) { b.buf >>= (e & 15u); b.cnt -= (e & 15u); return (int)(e >> 16); }
	return ngs_inf_decode_slow(b, count, sym, bits, slow);
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
		if (ngs_inf_build(T.lens, 19u, T.dcount, T.dsym, T.dfast, 7u, 0, T.dslow)) return -1;
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
	if (ngs_inf_build(T.lens + nlen, ndist, T.dcount, T.dsym, T.dfast, NGS_INF_DBITS, 2, T.dslow)) return -1;
	if (ngs_inf_build(T.lens, nlen, T.lcount, T.lsym, T.lfast, NGS_INF_LBITS, 1, T.lslow)) return -1;
	return 0;
}

/* Decoding loop with the output left to a policy:
 *   void literal(uint32_t pos, uint32_t byte); void match(uint32_t pos, uint32_t len, uint32_t dist);
 *   void stored(uint32_t pos, const uint8_t *src, uint32_t len); void finish();   (finish: nothing may stay pending)
 *   int tables(ngs_inf_bits &b, uint32_t btype, ngs_inf_tables &T);   builds the block's tables, normally by calling
 *       ngs_inf_block_tables; the device policy lets ONE lane build them in the warp's shared-memory slice and hands the
 *       advanced bit reader to the other lanes
 * `in` must be readable up to in_len + 7.  Returns the number of bytes produced, or -1 (malformed stream / output
 * larger than out_len). */
template <class Out>
NGS_HD long ngs_inflate_stream(const uint8_t *in, uint32_t in_len, uint32_t out_len, ngs_inf_tables &T, Out &out)
{
	ngs_inf_bits b; b.in = in; b.in_len = in_len; b.in_pos = 0; b.buf = 0; b.cnt = 0;
	uint32_t pos = 0, last;
	do {
		last = ngs_inf_get(b, 1);
		const uint32_t btype = ngs_inf_get(b, 2);
		if (btype == 0u) {
			const uint32_t drop = b.cnt & 7u;                      /* to the next byte boundary */
			b.buf >>= drop; b.cnt -= drop;
			const uint32_t len = ngs_inf_get(b, 16), nlen = ngs_inf_get(b, 16);
			if ((len ^ nlen) != 0xffffu) return -1;
			const uint32_t at = ngs_inf_consumed(b);
			if (at + len > in_len || pos + len > out_len) return -1;
			out.finish();
			out.stored(pos, in + at, len);
			pos += len;
			b.in_pos = at + len; b.buf = 0; b.cnt = 0;
		} else if (btype == 3u) return -1;
		else {
			if (out.tables(b, btype, T)) return -1;
			for (;;) {
				/* literal/length: at most 15 code bits + 5 extra bits */
				if (b.cnt < 32u) ngs_inf_refill32(b);
				uint32_t e = T.lfast[(uint32_t)b.buf & ((1u << NGS_INF_LBITS) - 1u)];
				if (e & 15u) { b.buf >>= (e & 15u); b.cnt -= (e & 15u); }
				else { const int s = ngs_inf_decode_slow(b, T.lcount, T.lsym, NGS_INF_LBITS, T.lslow); if (s < 0) return -1; e = ngs_inf_ll_entry((uint32_t)s); }
				const uint32_t kind = (e >> 8) & 3u;
				if (kind == NGS_INF_KIND_LIT) { if (pos >= out_len) return -1; out.literal(pos, e >> 16); ++pos; continue; }
				if (kind != NGS_INF_KIND_VAL) { if (kind == NGS_INF_KIND_EOB) break; return -1; }
				const uint32_t le = (e >> 4) & 15u;
				const uint32_t len = (e >> 16) + ((uint32_t)b.buf & ((1u << le) - 1u));
				b.buf >>= le; b.cnt -= le;
				/* distance: at most 15 code bits + 13 extra bits */
				if (b.cnt < 32u) ngs_inf_refill32(b);
				uint32_t d = T.dfast[(uint32_t)b.buf & ((1u << NGS_INF_DBITS) - 1u)];
				if (d & 15u) { b.buf >>= (d & 15u); b.cnt -= (d & 15u); }
				else { const int s = ngs_inf_decode_slow(b, T.dcount, T.dsym, NGS_INF_DBITS, T.dslow); if (s < 0) return -1; d = ngs_inf_d_entry((uint32_t)s); }
				if (((d >> 8) & 3u) != NGS_INF_KIND_VAL) return -1;
				const uint32_t de = (d >> 4) & 15u;
				const uint32_t dist = (d >> 16) + ((uint32_t)b.buf & ((1u << de) - 1u));
				b.buf >>= de; b.cnt -= de;
				if (dist > pos || pos + len > out_len) return -1;
				out.match(pos, len, dist);
				pos += len;
			}
		}
		if (ngs_inf_consumed(b) > in_len) return -1;               /* ran past the compressed data */
	} while (!last);
	out.finish();
	return (long)pos;
}

#endif
