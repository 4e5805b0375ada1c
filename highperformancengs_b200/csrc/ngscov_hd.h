/* csrc/ngscov_hd.h -- record-level logic shared by every kernel of libngscov_cuda.
 *
 * Everything here is `NGS_HD` (__host__ __device__) and free of CUDA intrinsics so that the exact code
 * the kernels run can also be compiled with g++ and unit-tested on a CPU-only machine
 * (tests/hd_check.cpp).  It is NOT a CPU fallback: the library itself only ever calls it from device
 * code, plus ngscov_bins_apply_run for the few boundary runs of the shard stitcher.
 *
 * Semantics restated (reference file:line):
 *   record layout            samtools-0.1.19 bam.h:178-187 (bam1_core_t), bam.c:191-223 (bam_read1)
 *   flag filter, CIGAR walk  bam2depth.c:86-110 (mask 0x704), bam2wig.c:84-109 (mask 0x4)
 *   28-bit position keys     hashtbl.c:243-259 (int2char / char2int)
 *   window bins, variant A   bam2depth.c:132-176 ; variant B bam2wig.c:131-175
 *   "%.2f" of bins/window    bam2depth.c:243,253 (glibc printf, round-half-even on the exact double)
 */
#ifndef NGSCOV_HD_H
#define NGSCOV_HD_H
#include <stdint.h>

#if defined(__CUDACC__)
#define NGS_HD __host__ __device__ __forceinline__
#define NGS_HDM __host__ __device__ __forceinline__      /* member functions */
#else
#define NGS_HD static inline
#define NGS_HDM inline
#endif

#define NGS_POS_MASK   0x0FFFFFFFu        /* 4 x 7 bit key of the reference */
#define NGS_MAX_BS     (1u << 26)         /* block_size plausibility bound used only by speculation */
#define NGS_CARRY_MAX  (512u * 1024u)     /* > 36 + 255 + 4*65535: any header+CIGAR fits */

/* exit codes of a sub-chunk walk, stored in the per-sub-chunk exit slot */
#define NGS_NONE       (-(1ll << 62))                   /* entry: no record start found / none expected */
#define NGS_EXIT_INF   (1ll << 60)                  /* exit >= INF: chain ended inside this chunk */
#define NGS_EXIT_PARTIAL(o) (NGS_EXIT_INF + (o) + (1ll << 40))   /* header of record at o is cut by chunk end */
#define NGS_EXIT_INVALID (-(1ll << 62) - 1)                  /* a walked record failed structural validation */

typedef struct {
	int32_t  block_size;
	int32_t  ref_id;
	int32_t  pos;
	uint32_t l_qname;      /* byte  */
	uint32_t n_cigar;      /* u16   */
	uint32_t flag;         /* u16   */
	int32_t  l_seq;
	int32_t  next_ref_id;
	int32_t  next_pos;
} ngs_hdr;

/* Per-segment geometry.  A segment is the tracked part of one contig on this shard. */
typedef struct {
	uint32_t contig;       /* index into the header's contig list */
	uint32_t clip_lo;      /* first tracked position (contig coordinates) */
	uint32_t lim;          /* last index that may receive a delta (the -1 of an interval ending here) */
	uint32_t hard_end;     /* 1: lim is contig end + pad (going beyond is ERANGE); 0: lim is a shard cut (clip) */
	uint64_t base;         /* element offset of clip_lo in the difference array */
	uint64_t tile0;        /* first scan tile of the segment */
	uint64_t bins0;        /* first bin of the contig in the concatenated bins array */
	uint32_t ref_len;      /* contig length */
	uint32_t n_win;        /* ref_len / window + 1 */
} ngs_seg;

/* structural validation applied to EVERY walked record (a real chain never fails it on a well-formed
 * BAM; a mis-speculated chain almost always does, quickly) */
NGS_HD int ngs_hdr_valid(const ngs_hdr &h, int32_t n_ref)
{
	if (h.block_size < 32) return 0;
	if (h.ref_id < -1 || h.ref_id >= n_ref) return 0;
	if (h.l_seq < 0) return 0;
	uint64_t need = 32ull + h.l_qname + 4ull * h.n_cigar + ((uint64_t)h.l_seq + 1) / 2 + (uint64_t)h.l_seq;
	return need <= (uint64_t)(uint32_t)h.block_size;
}

/* stricter test used only to pick a speculative entry point */
NGS_HD int ngs_hdr_plausible(const ngs_hdr &h, int32_t n_ref, const uint32_t *ref_len)
{
	if (!ngs_hdr_valid(h, n_ref)) return 0;
	if ((uint32_t)h.block_size > NGS_MAX_BS) return 0;
	if (h.l_qname < 1) return 0;
	if (h.pos < -1) return 0;
	if (h.ref_id >= 0 && (uint32_t)h.pos != 0xFFFFFFFFu && (uint32_t)h.pos > ref_len[h.ref_id]) return 0;
	if (h.next_ref_id < -1 || h.next_ref_id >= n_ref) return 0;
	if (h.next_pos < -1) return 0;
	return 1;
}

/* ---- window bins ---------------------------------------------------------------------------------
 * Adds the contribution of run [s,e) with depth d to bins[0..n_win) through add(k, value).
 * variant 0 = bam2depth (half-open windows): exact sum of depth over the bases of each window.
 * variant 1 = bam2wig  (inclusive window ends): restates the reference loop, off-by-one terms and all,
 *             starting at the first window whose end is >= s (the reference's persistent cursor only
 *             skips windows that cannot contribute). */
template <class Add>
NGS_HD void ngs_bin_run(int variant, uint32_t s, uint32_t e, int64_t d, uint32_t W, uint32_t L, uint32_t n_win, Add &add)
{
	if (e <= s) return;
	uint64_t k = s / W;
	if (variant == 0) {
		for (; k < n_win; ++k) {
			uint64_t ws = k * (uint64_t)W, we = ws + W;
			if (we > L) we = L;
			if (e <= ws) break;
			uint64_t lo = s > ws ? s : ws, hi = e < we ? e : we;
			if (hi > lo) add((uint32_t)k, (int64_t)(hi - lo) * d);
			if (e <= we) break;
		}
	} else {
		for (; k <= n_win; ++k) {
			uint64_t ws = k * (uint64_t)W, we = ws + W - 1;
			if (we > L) we = L;
			if (e < ws) break;
			int64_t c; int stop;
			if (s < ws) {
				if (e < we) { c = (int64_t)(e - ws); stop = 1; }
				else { c = (int64_t)(we - ws + 1); stop = 0; }
			} else if (s <= we) {
				if (e <= we) { c = (int64_t)(e - s); stop = 1; }
				else { c = (int64_t)(we - s); stop = 0; }
			} else continue;
			if (k < n_win && c != 0) add((uint32_t)k, c * d);   /* k == n_win is out of bounds in the reference too */
			if (stop) break;
		}
	}
}

/* ---- decimal formatting --------------------------------------------------------------------------*/
NGS_HD uint32_t ngs_ndigits(uint32_t v)
{
	/* balanced decision tree: four compares instead of a chain of nine */
	return v < 10000u ? (v < 100u ? (v < 10u ? 1u : 2u) : (v < 1000u ? 3u : 4u))
	                  : (v < 1000000u ? (v < 100000u ? 5u : 6u)
	                                  : (v < 100000000u ? (v < 10000000u ? 7u : 8u) : (v < 1000000000u ? 9u : 10u)));
}
NGS_HD uint32_t ngs_ndigits64(uint64_t v) { uint32_t n = 1; while (v >= 10) { v /= 10; ++n; } return n; }
/* printf("%d") length */
NGS_HD uint32_t ngs_len_i32(int32_t v) { return v < 0 ? 1 + ngs_ndigits((uint32_t)(-(int64_t)v)) : ngs_ndigits((uint32_t)v); }
/* writes printf("%d") of v at dst, returns bytes written */
template <class Ptr>
NGS_HD uint32_t ngs_put_u64(Ptr dst, uint64_t v)
{
	uint32_t n = ngs_ndigits64(v);
	for (uint32_t i = n; i-- > 0;) { dst[i] = (char)('0' + (uint32_t)(v % 10)); v /= 10; }
	return n;
}
template <class Ptr>
NGS_HD uint32_t ngs_put_u32(Ptr dst, uint32_t v)
{
	uint32_t n = ngs_ndigits(v);
	for (uint32_t i = n; i-- > 0;) { uint32_t q = v / 10u; dst[i] = (char)('0' + (v - q * 10u)); v = q; }
	return n;
}
/* eight decimal digits of v (< 10^8) as ASCII, most significant digit in the LOWEST byte, computed with
 * three SWAR divide steps on packed lanes instead of eight dependent divisions by ten */
NGS_HD uint64_t ngs_digits8(uint32_t v)
{
	const uint32_t hi4 = v / 10000u, lo4 = v - hi4 * 10000u;
	uint64_t x = (uint64_t)hi4 | ((uint64_t)lo4 << 32);                 /* two 4-digit numbers in 32-bit lanes */
	uint64_t q = ((x * 5243ull) >> 19) & 0x0000007F0000007Full;          /* /100 per lane (exact below 10^4) */
	x = q | ((x - q * 100ull) << 16);                                    /* four 2-digit numbers in 16-bit lanes */
	q = ((x * 103ull) >> 10) & 0x000F000F000F000Full;                    /* /10 per lane (exact below 100) */
	x = q | ((x - q * 10ull) << 8);                                      /* eight digits in byte lanes */
	return x + 0x3030303030303030ull;
}
/* writes the n (1..10) digits of v at dst */
template <class Ptr>
NGS_HD void ngs_put_u32_swar(Ptr dst, uint32_t v, uint32_t n)
{
	const uint32_t top = v / 100000000u;                                 /* 0..42 */
	uint64_t d = ngs_digits8(v - top * 100000000u);
	if (n > 8) {
		if (n == 10) { dst[0] = (char)('0' + top / 10u); dst[1] = (char)('0' + top % 10u); dst += 2; }
		else { dst[0] = (char)('0' + top); dst += 1; }
		n = 8;
	}
	d >>= 8 * (8 - n);
	const uint32_t lo = (uint32_t)d, hi = (uint32_t)(d >> 32);
	dst[0] = (char)(lo & 0xff);                       /* n >= 1 */
	if (n > 1) dst[1] = (char)((lo >> 8) & 0xff);
	if (n > 2) dst[2] = (char)((lo >> 16) & 0xff);
	if (n > 3) dst[3] = (char)(lo >> 24);
	if (n > 4) dst[4] = (char)(hi & 0xff);
	if (n > 5) dst[5] = (char)((hi >> 8) & 0xff);
	if (n > 6) dst[6] = (char)((hi >> 16) & 0xff);
	if (n > 7) dst[7] = (char)(hi >> 24);
}
/* same with the digit count already known */
template <class Ptr>
NGS_HD void ngs_put_u32n(Ptr dst, uint32_t v, uint32_t n)
{
	for (uint32_t i = n; i-- > 0;) { uint32_t q = v / 10u; dst[i] = (char)('0' + (v - q * 10u)); v = q; }
}
template <class Ptr>
NGS_HD uint32_t ngs_put_i32(Ptr dst, int32_t v)
{
	if (v < 0) { dst[0] = '-'; return 1 + ngs_put_u32(dst + 1, (uint32_t)(-(int64_t)v)); }
	return ngs_put_u32(dst, (uint32_t)v);
}

/* value of printf("%.2f", (double)bins / (double)window) as an integer number of hundredths, rounded
 * half-to-even on the EXACT binary value of the IEEE quotient (what glibc does).  `q` must be that
 * correctly rounded quotient (device: __ddiv_rn; host: plain division).  q >= 0. */
NGS_HD uint64_t ngs_hundredths(double q)
{
	union { double d; uint64_t u; } cv; cv.d = q;
	uint64_t bits = cv.u;
	int ex = (int)((bits >> 52) & 0x7ff);
	uint64_t m = bits & 0xFFFFFFFFFFFFFull;
	if (ex == 0) { if (m == 0) return 0; ex = 1; } else m |= 1ull << 52;
	int sh = 1075 - ex;                          /* q = m * 2^-sh */
	if (sh <= 0) {                               /* integer >= 2^52 (not reachable for bins < 2^53) */
		uint64_t ip = m << (-sh > 10 ? 10 : -sh);
		return ip * 100;
	}
	if (sh >= 64) return 0;                      /* m*100 < 2^60 <= 2^(sh-1): below one half */
	uint64_t N = m * 100ull;
	uint64_t t = N >> sh, rem = N & ((1ull << sh) - 1), half = 1ull << (sh - 1);
	if (rem > half || (rem == half && (t & 1))) ++t;
	return t;
}
NGS_HD uint32_t ngs_len_f2(uint64_t hundredths) { return ngs_ndigits64(hundredths / 100) + 3; }
template <class Ptr>
NGS_HD uint32_t ngs_put_f2(Ptr dst, uint64_t hundredths)
{
	uint32_t n = ngs_put_u64(dst, hundredths / 100);
	uint32_t f = (uint32_t)(hundredths % 100);
	dst[n] = '.'; dst[n + 1] = (char)('0' + f / 10); dst[n + 2] = (char)('0' + f % 10);
	return n + 3;
}

/* ---- register line assembly (fast bedGraph path of K5, k_format_bed) ------------------------------
 * A line is appended to a byte stream in shared memory as a few "pieces" of at most 12 bytes held in
 * registers.  The writer keeps the stream's current partial word in a register (`carry`, `k` valid low
 * bytes) and only ever stores whole 32-bit words, so a line costs about ten word stores instead of one
 * store (plus extraction and a predicate) per byte.  The first word of a thread's range may be shared
 * with the previous thread: the owner of the range's HEAD stores it (low bytes zero), the previous
 * thread ORs its tail in after a barrier (ngs_lw::carry / k after the last piece). */
NGS_HD uint32_t ngs_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
	return __byte_perm(a, b, sel);
#else
	const uint64_t pair = (uint64_t)a | ((uint64_t)b << 32);
	uint32_t r = 0;
	for (int j = 0; j < 4; ++j) r |= (uint32_t)((pair >> (8 * ((sel >> (4 * j)) & 7))) & 0xff) << (8 * j);
	return r;
#endif
}
/* index of the lowest non-zero byte of x (x != 0) */
NGS_HD uint32_t ngs_low_byte_index(uint32_t x)
{
#if defined(__CUDA_ARCH__)
	return (uint32_t)(__ffs((int)x) - 1) >> 3;
#else
	return (uint32_t)__builtin_ctz(x) >> 3;
#endif
}
/* four decimal digits of y (< 10^4), one per byte, most significant in the LOWEST byte (no '0' offset):
 * the 32-bit-lane form of ngs_digits8 (64-bit SWAR costs three to four IMADs per multiply on the GPU) */
NGS_HD uint32_t ngs_digits4_raw(uint32_t y)
{
	const uint32_t q = (y * 5243u) >> 19;                        /* y / 100 (exact below 10^4) */
	const uint32_t z = q + ((y - q * 100u) << 16);               /* two 2-digit numbers in 16-bit lanes */
	const uint32_t t = ((z * 103u) >> 10) & 0x000F000Fu;         /* / 10 per lane (exact below 100) */
	return t + ((z - t * 10u) << 8);
}
/* decimal ASCII of v.  n = tn & 0xff digits.  n <= 8: digits left-aligned in (lo, hi) -- first significant digit in
 * byte 0 of lo, bytes beyond n are zero.  n == 9: the eight low digits in (lo, hi), the leading digit (ASCII) in
 * tn >> 8.  n == 10 is only a flag (v >= 10^9): the caller takes its slow path. */
typedef struct { uint32_t lo, hi, tn; } ngs_num;
NGS_HD void ngs_num_conv(uint32_t v, ngs_num &o)
{
	const uint32_t top = v / 100000000u;
	v -= top * 100000000u;
	const uint32_t h4 = v / 10000u, l4 = v - h4 * 10000u;
	const uint32_t a = ngs_digits4_raw(h4), b = ngs_digits4_raw(l4);
	if (top) {
		o.lo = a + 0x30303030u; o.hi = b + 0x30303030u; o.tn = ((top + 0x30u) << 8) | (top < 10u ? 9u : 10u);
		return;
	}
	const uint32_t lz = a ? ngs_low_byte_index(a) : (b ? 4u + ngs_low_byte_index(b) : 7u);
	const uint64_t d = (((uint64_t)(b + 0x30303030u) << 32) | (uint64_t)(a + 0x30303030u)) >> (8u * lz);
	o.lo = (uint32_t)d; o.hi = (uint32_t)(d >> 32); o.tn = 8u - lz;
}
/* the up to ten bytes "digits, separator" of x as a 12-byte piece */
NGS_HD void ngs_num_piece(const ngs_num &x, uint32_t sep, uint32_t &q0, uint32_t &q1, uint32_t &q2, uint32_t &cnt)
{
	const uint32_t n = x.tn & 0xffu;
	if (n <= 8u) {
		const uint32_t sv = sep << ((n & 3u) * 8u);
		q0 = x.lo; q1 = x.hi; q2 = 0;
		if (n < 4u) q0 |= sv; else if (n < 8u) q1 |= sv; else q2 = sep;
		cnt = n + 1u;
	} else {
		q0 = (x.tn >> 8) | (x.lo << 8); q1 = (x.lo >> 24) | (x.hi << 8); q2 = (x.hi >> 24) | (sep << 8);
		cnt = 10u;
	}
}
/* Positions are written as H = p / 10^4 (thread-cached: the records of a thread lie within a few bases of each
 * other, so H changes about once in a thousand threads) followed by the four digits of p % 10^4.  With the cache
 * valid a position costs one 4-digit conversion and three byte permutes whose selectors only depend on the
 * number of digits of H.  p < 10^4 (H == 0) and depths go through ngs_small_piece. */
typedef struct {
	uint32_t H, hlo, sh, nH, selA, selB, selC;      /* sh = '\t' | fifth digit of H << 8 */
	uint32_t nh0, nh1, nh2, nh3;                     /* "name\tHHHHH": the constant head of every line of the thread */
	uint32_t th0, th1;                               /* "\tHHHHH": what separates the low digits of start and end */
} ngs_hcache;
NGS_HD uint32_t ngs_ascii4(uint32_t y) { return ngs_digits4_raw(y) + 0x30303030u; }
NGS_HD void ngs_hcache_fill(ngs_hcache &c, uint32_t H, uint32_t name_lo, uint32_t name_hi, uint32_t nl)   /* 1 <= H < 10^5, 1 <= nl <= 8 */
{
	/* byte selectors of the 12-byte piece "H digits, 4 digits, tab": sources (hlo, L4) for bytes 0-3 and (L4, sh) for
	 * bytes 4-11; in sh byte 0 is the tab, byte 1 the fifth digit of H, byte 2 zero */
	const uint32_t top = H / 10000u, raw = ngs_digits4_raw(H - top * 10000u), a = raw + 0x30303030u;
	uint32_t hhi = 0, n;
	if (top) { c.hlo = (top + 0x30u) | (a << 8); hhi = a >> 24; n = 5u; }
	else { const uint32_t lz = ngs_low_byte_index(raw); c.hlo = a >> (8u * lz); n = 4u - lz; }
	c.sh = 0x09u | (hhi << 8);
	/* selectors for n = 1..5 digits of H, 16 bits each, packed into constants (a local table would live on the stack):
	 * A: 6540 5410 4210 3210 3210   B: 6643 6432 4321 3210 2105   C: 6666 6666 6666 6664 6643 */
	const uint32_t sh16 = 16u * (n - 1u);
	c.H = H; c.nH = n;
	c.selA = n < 4u ? (uint32_t)(0x421054106540ull >> sh16) & 0xffffu : 0x3210u;
	c.selB = n < 5u ? (uint32_t)(0x3210432164326643ull >> sh16) & 0xffffu : 0x2105u;
	c.selC = n < 4u ? 0x6666u : (n == 4u ? 0x6664u : 0x6643u);
	const uint64_t hs = (uint64_t)c.hlo | ((uint64_t)hhi << 32), nm = (uint64_t)name_lo | ((uint64_t)name_hi << 32);
	const uint64_t lo = nm | (nl < 8u ? hs << (8u * nl) : 0ull), hi = hs >> (64u - 8u * nl);
	const uint64_t th = 0x09ull | (hs << 8);
	c.nh0 = (uint32_t)lo; c.nh1 = (uint32_t)(lo >> 32); c.nh2 = (uint32_t)hi; c.nh3 = (uint32_t)(hi >> 32);
	c.th0 = (uint32_t)th; c.th1 = (uint32_t)(th >> 32);
}
/* v < 10^4 followed by sep, left-aligned in (q0, q1); cnt = digits + 1 */
NGS_HD void ngs_small_piece(uint32_t v, uint32_t sep, uint32_t &q0, uint32_t &q1, uint32_t &cnt)
{
	const uint32_t raw = ngs_digits4_raw(v), lz = raw ? ngs_low_byte_index(raw) : 3u;
	const uint32_t a = (raw + 0x30303030u) >> (8u * lz);
	q0 = lz ? (a | (sep << (32u - 8u * lz))) : a;
	q1 = lz ? 0u : sep;
	cnt = 5u - lz;
}
/* p = H * 10^4 + L with the cache holding H (H >= 1), L4 = ngs_ascii4(L) */
NGS_HD void ngs_pos_piece(const ngs_hcache &c, uint32_t L4, uint32_t &q0, uint32_t &q1, uint32_t &q2, uint32_t &cnt)
{
	q0 = ngs_prmt(c.hlo, L4, c.selA); q1 = ngs_prmt(L4, c.sh, c.selB); q2 = ngs_prmt(L4, c.sh, c.selC);
	cnt = c.nH + 5u;
}

/* stream writer: W is a pointer-like type to 32-bit words (shared memory on the device) */
template <class W>
struct ngs_lw {
	W w;                 /* word the next byte goes to */
	uint32_t carry, k;   /* the k (0..3) low bytes of that word already produced */
	NGS_HDM void begin(W base_words, uint32_t byte_off) { w = base_words + (byte_off >> 2); k = byte_off & 3u; carry = 0; }
	NGS_HDM void put8(uint32_t q0, uint32_t q1, uint32_t cnt)        /* cnt in 1..8, bytes beyond cnt zero */
	{
		const uint32_t sel = 0x7654u - 0x1111u * k;
		const uint32_t x0 = ngs_prmt(0u, q0, sel) | carry, x1 = ngs_prmt(q0, q1, sel), x2 = ngs_prmt(q1, 0u, sel);
		const uint32_t tot = k + cnt, nw = tot >> 2;
		if (nw > 0u) w[0] = x0;
		if (nw > 1u) w[1] = x1;
		carry = nw == 0u ? x0 : (nw == 1u ? x1 : x2);
		w += nw; k = tot & 3u;
	}
	NGS_HDM void put12(uint32_t q0, uint32_t q1, uint32_t q2, uint32_t cnt)   /* cnt in 1..12 */
	{
		const uint32_t sel = 0x7654u - 0x1111u * k;
		const uint32_t x0 = ngs_prmt(0u, q0, sel) | carry, x1 = ngs_prmt(q0, q1, sel), x2 = ngs_prmt(q1, q2, sel), x3 = ngs_prmt(q2, 0u, sel);
		const uint32_t tot = k + cnt, nw = tot >> 2;
		if (nw > 0u) w[0] = x0;
		if (nw > 1u) w[1] = x1;
		if (nw > 2u) w[2] = x2;
		carry = nw < 2u ? (nw == 0u ? x0 : x1) : (nw == 2u ? x2 : x3);
		w += nw; k = tot & 3u;
	}
	NGS_HDM void put16(uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3, uint32_t cnt)   /* cnt in 1..16 */
	{
		const uint32_t sel = 0x7654u - 0x1111u * k;
		const uint32_t x0 = ngs_prmt(0u, q0, sel) | carry, x1 = ngs_prmt(q0, q1, sel), x2 = ngs_prmt(q1, q2, sel), x3 = ngs_prmt(q2, q3, sel), x4 = ngs_prmt(q3, 0u, sel);
		const uint32_t tot = k + cnt, nw = tot >> 2;
		if (nw > 0u) w[0] = x0;
		if (nw > 1u) w[1] = x1;
		if (nw > 2u) w[2] = x2;
		if (nw > 3u) w[3] = x3;
		carry = nw < 2u ? (nw == 0u ? x0 : x1) : (nw == 2u ? x2 : (nw == 3u ? x3 : x4));
		w += nw; k = tot & 3u;
	}
	/* One bedGraph line.  L4s / L4e = ngs_ascii4(s % 10^4), ngs_ascii4(e % 10^4); d >= 1.  When start and end share
	 * the cached H and the depth has at most four digits the line is three pieces: the thread-constant head
	 * "name\tH", "ssss\tH" and "eeee\tdepth\n"; everything else goes piece by piece. */
	NGS_HDM void bed_line_cached(ngs_hcache &c, uint32_t name_lo, uint32_t name_hi, uint32_t nl,
	                             uint32_t s, uint32_t L4s, uint32_t e, uint32_t L4e, uint32_t d)
	{
		const uint32_t Hs = s / 10000u, He = e / 10000u;
		if (Hs != 0u && Hs == He && d < 10000u) {
			if (Hs != c.H) ngs_hcache_fill(c, Hs, name_lo, name_hi, nl);
			uint32_t dq0, dq1, dc;
			ngs_small_piece(d, 0x0Au, dq0, dq1, dc);
			put16(c.nh0, c.nh1, c.nh2, c.nh3, nl + c.nH);
			put12(L4s, c.th0, c.th1, 5u + c.nH);
			put12(L4e, 0x09u | (dq0 << 8), (dq0 >> 24) | (dq1 << 8), 5u + dc);
			return;
		}
		put8(name_lo, name_hi, nl);
		put_pos(c, s, name_lo, name_hi, nl);
		put_pos(c, e, name_lo, name_hi, nl);
		if (d < 10000u) { uint32_t q0, q1, cnt; ngs_small_piece(d, 0x0Au, q0, q1, cnt); put8(q0, q1, cnt); }
		else { ngs_num x; uint32_t q0, q1, q2, cnt; ngs_num_conv(d, x); ngs_num_piece(x, 0x0Au, q0, q1, q2, cnt); put12(q0, q1, q2, cnt); }
	}
	/* position p followed by a tab */
	NGS_HDM void put_pos(ngs_hcache &c, uint32_t p, uint32_t name_lo, uint32_t name_hi, uint32_t nl)
	{
		const uint32_t H = p / 10000u, Lo = p - H * 10000u;
		uint32_t q0, q1, q2 = 0, cnt;
		if (H) {
			if (H != c.H) ngs_hcache_fill(c, H, name_lo, name_hi, nl);
			ngs_pos_piece(c, ngs_ascii4(Lo), q0, q1, q2, cnt);
		} else ngs_small_piece(Lo, 0x09u, q0, q1, cnt);
		put12(q0, q1, q2, cnt);
	}
	/* "name\tstart\tend\tdepth\n": name8 = the first eight bytes of "name\t" (nl <= 8 of them valid, rest zero) */
	NGS_HDM void bed_line(uint32_t name_lo, uint32_t name_hi, uint32_t nl, const ngs_num &s, const ngs_num &e, const ngs_num &d)
	{
		uint32_t q0, q1, q2, cnt;
		put8(name_lo, name_hi, nl);
		ngs_num_piece(s, 0x09u, q0, q1, q2, cnt); put12(q0, q1, q2, cnt);
		ngs_num_piece(e, 0x09u, q0, q1, q2, cnt); put12(q0, q1, q2, cnt);
		ngs_num_piece(d, 0x0Au, q0, q1, q2, cnt); put12(q0, q1, q2, cnt);
	}
};

/* ---- the record walk -----------------------------------------------------------------------------
 * Policy `Mem` supplies the loads:  void hdr_lite(int64 o, ngs_hdr&)  first 24 bytes of the record,
 *                                   void hdr(int64 o, ngs_hdr&)       all 36 bytes,
 *                                   uint32 u32(int64 o), uint8 u8(int64 o), int64 size().
 * Policy `Sink` receives deltas:    void add2(uint64 idx_plus, uint64 idx_minus, int sign),
 *                                   void flag_range_error().
 */
typedef struct {
	uint64_t records, kept, m_ops, cigar_words;
	uint32_t max_ref_p1;         /* 1 + highest refID seen (0: none): sorted input => every lower contig is complete */
} ngs_counts;
typedef struct { uint32_t records, kept, m_ops, cigar_words, max_ref_p1; } ngs_counts32;   /* one row's worth */

typedef struct {
	int32_t  n_ref;
	uint32_t flag_mask;
	const int32_t *ctg2seg;      /* [n_ref] segment index of a contig or -1 (outside the shard) */
	const ngs_seg *seg;
	int32_t  end_ctg;            /* region shard: a record of this contig at or beyond end_pos lies behind the slice */
	uint32_t end_pos;            /* (end_ctg = -2: no slice).  Such a record counts as "the next contig has begun" */
	int32_t  reg_ref;            /* NGSCOV_MODE_SLIDE, `-r` of bam_sliding_count: only records bam_fetch(reg_ref, reg_beg, reg_end) */
	int32_t  reg_beg, reg_end;   /* hands out are counted; reg_ref < 0: no region */
} ngs_walk_cfg;

/* The record test of bam_fetch (samtools-0.1.19 bam_index.c: bam_iter_read stops at another refID or POS >= end,
 * is_overlap wants the record's end > beg in unsigned arithmetic; the end is bam_calend over the reference-consuming
 * CIGAR operations M D N = X, or POS + 1 without a CIGAR).  cig(i) = CIGAR word i.  The 'B' operation of bam_calend is
 * not restated (no aligner writes it). */
template <class Cig>
NGS_HD bool ngs_region_pass(const ngs_walk_cfg &cfg, int32_t ref_id, int32_t pos, uint32_t n_cigar, const Cig &cig)
{
	if (cfg.reg_ref < 0) return true;
	if (ref_id != cfg.reg_ref) return false;
	const uint32_t rbeg = (uint32_t)pos, beg = (uint32_t)cfg.reg_beg;
	if (rbeg >= (uint32_t)cfg.reg_end) return false;
	if (rbeg > beg) return true;                    /* its end is at least POS */
	uint32_t rend = rbeg + 1u;
	if (n_cigar) {
		rend = rbeg;
		for (uint32_t k = 0; k < n_cigar; ++k) { const uint32_t w = cig(k); if ((0x3C1A7u >> ((w & 15u) << 1)) & 2u) rend += w >> 4; }
	}
	return rend > beg;
}

/* ---- compact record table (K1 -> owner tiles) ---------------------------------------------------
 * K1 (k_chain_classify) frames AND decodes: every record that contributes at least one M interval leaves one
 * 8-byte entry in the ROW of the sub-chunk it starts in (at most NGS_ROW_SLOTS entries per row), so that the
 * stream is never read a second time for ordinary records:
 *   x = pos (29 bits, the record's POS) | cdelta << 29   (contig = row.ctg + cdelta, cdelta 0..6)
 *   y = kind << 30 | payload
 *       SIMPLE  payload = len               one interval [pos, pos+len)  (abutting M ops are merged: the
 *                                           zero-net deltas between them are transparent to the sweep)
 *       PAIR    payload = len1 | gap << 12 | len2 << 18   [pos, pos+len1) and [pos+len1+gap, .. +len2)
 *       CHAIN   payload = byte offset of the record from the row's entry point: decoded from the stream by
 *                         k_special_reds while the chunk is still resident (long CIGARs, leading D/N, 28-bit
 *                         key aliasing, records reaching outside the shard's clip range, more than seven
 *                         contigs in one row)
 * SIMPLE and PAIR records lie entirely inside [clip_lo, lim] of their segment and below 2^28, so the owner of a
 * tile applies them without any clipping.
 * A delta at position q of a record starting at pos is applied by the OWNER of q's tile (k_cover: shared-memory
 * accumulation, never materialised in HBM) when the record is visible to it: P0(tile(q)) - NGS_NEAR_H <= pos.
 * Every other delta is "far" and is added to the global difference array with a RED by k_special_reds, which
 * also raises the tile's far flag so that the owner merges that tile of the array.  Both kernels evaluate
 * ngs_delta_is_near, so each delta is applied exactly once. */
#define NGS_ENT_SIMPLE  0u
#define NGS_ENT_PAIR    1u
#define NGS_ENT_CHAIN   2u
#define NGS_ENT_CDELTA_MAX 6u
#define NGS_ROW_SLOTS   256            /* entries per row = threads of an owner CTA */
#define NGS_NEAR_H      512            /* positions before a tile from which records are visible to its owner */
#ifndef NGS_CT
#define NGS_CT          4096           /* positions per owner tile (k_cover) */
#endif
#define NGS_CPT         (4096 / NGS_CT) /* owner tiles per difference-array tile (NGS_TILE / NGS_CT) */
#define NGS_KEY_NONE    0xFFFFFFFFFFFFFFFFull
#define NGS_ROW_UNSORTED 0x80000000u   /* row.special: record keys decrease inside the row */
#define NGS_ROW_OVERFLOW 0x40000000u   /* row.special: more entries than slots -- the whole row goes the RED way */

/* per-row (= per sub-chunk) metadata written by K1 together with the entries */
typedef struct {
	uint64_t fk, lk;             /* sort key of the first / last record starting in the row (fk = NGS_KEY_NONE: none) */
	uint32_t count;              /* entries stored (0 when the row overflowed) */
	uint32_t special;            /* CHAIN + possibly-far entries | NGS_ROW_* flags */
	int32_t  ctg;                /* contig of cdelta 0 (-1: no entry) */
	uint32_t nrec;               /* records starting in the row */
	uint32_t nkept, nmops, ncig; /* as ngs_counts, for this row */
	uint32_t max_ref_p1;
	uint64_t pad0, pad1;
} ngs_row;

/* sort key of a record: samtools sorts by (refID as unsigned, pos); unmapped (refID -1) sort last */
NGS_HD uint64_t ngs_sort_key(int32_t ref_id, int32_t pos) { return ((uint64_t)(uint32_t)ref_id << 32) | (uint32_t)(pos + 1); }

/* Is the delta at contig position q (inside its segment) of a record starting at rec_pos applied by the owner of
 * q's tile?  The tile grid of a segment starts at clip_lo. */
NGS_HD int ngs_delta_is_near(uint32_t clip_lo, uint32_t rec_pos, uint32_t q)
{
	const int64_t p0 = (int64_t)clip_lo + (int64_t)((q - clip_lo) / NGS_CT) * NGS_CT;
	return (int64_t)rec_pos >= p0 - NGS_NEAR_H;
}

/* Intervals of a SIMPLE / PAIR entry in contig coordinates: n = 1 or 2 */
NGS_HD int ngs_entry_intervals(uint32_t pos, uint32_t y, uint32_t a[2], uint32_t b[2])
{
	const uint32_t kind = y >> 30, pl = y & 0x3FFFFFFFu;
	a[0] = pos;
	if (kind == NGS_ENT_SIMPLE) { b[0] = pos + pl; a[1] = b[1] = 0; return 1; }
	b[0] = pos + (pl & 0xFFFu);
	a[1] = b[0] + ((pl >> 12) & 0x3Fu);
	b[1] = a[1] + (pl >> 18);
	return 2;
}

/* What K1 learns about one record (cig(i) returns CIGAR word i).  Returns 0 = no entry (filtered / nothing to count);
 * 1 = entry word y to store (its x is pos | cdelta << 29, made by the caller); special = 1 when the entry is CHAIN or
 * may own a far delta.  Counters as ngs_decode_record except m_ops, which counts MERGED intervals here. */
template <class Cig, class Cnt>
NGS_HD int ngs_classify_record(const ngs_walk_cfg &cfg, const ngs_hdr &h, uint32_t slot_off, Cig &cig,
                               uint32_t &y, int &special, Cnt &cnt)
{
	cnt.records += 1; cnt.cigar_words += h.n_cigar;
	{
		uint32_t r1 = (uint32_t)(h.ref_id + 1);
		if (h.ref_id == cfg.end_ctg && (uint32_t)h.pos >= cfg.end_pos) r1 += 1;
		if (r1 > cnt.max_ref_p1) cnt.max_ref_p1 = r1;
	}
	if ((h.flag & cfg.flag_mask) || h.ref_id < 0 || (uint32_t)h.pos >= (1u << 29)) return 0;
	cnt.kept += 1;
	const int32_t si = cfg.ctg2seg[h.ref_id];
	if (si < 0 || h.n_cigar == 0) return 0;
	/* walk the CIGAR: M intervals with abutting ones merged */
	const uint32_t pos = (uint32_t)h.pos;
	uint32_t na = 0, a0 = 0, b0 = 0, a1 = 0, b1 = 0, cb = 0;
	int wrapped = 0;
	const uint32_t w0 = cig(0);
	if (h.n_cigar == 1) {                                /* the common case, "<len>M": no loop, no divergence among the lanes of a walk */
		if ((w0 & 0xf) != 0 || (w0 >> 4) == 0) return 0;
		a0 = pos; b0 = cb = pos + (w0 >> 4); na = 1;     /* pos < 2^29, len < 2^28: no 32-bit wrap */
	} else {
		uint32_t p = pos, ca = 0; int open = 0;
		for (uint32_t i = 0; i < h.n_cigar; ++i) {
			const uint32_t w = i ? cig(i) : w0, op = w & 0xf, len = w >> 4;
			if (op == 0) {
				if (len == 0) continue;
				if (open && cb == p) cb = p + len;
				else {
					if (open) { if (na == 0) { a0 = ca; b0 = cb; } else if (na == 1) { a1 = ca; b1 = cb; } ++na; }
					ca = p; cb = p + len; open = 1;
				}
				if (p + len < p) wrapped = 1;
				p += len;
			} else if (op == 2 || op == 3) { if (p + len < p) wrapped = 1; p += len; }
		}
		if (!open) return 0;                             /* no M at all */
		if (na == 0) { a0 = ca; b0 = cb; } else if (na == 1) { a1 = ca; b1 = cb; }
		++na;
	}
	cnt.m_ops += na;
	const ngs_seg &sg = cfg.seg[si];
	/* entirely inside the tracked range of the segment and below the 28-bit key wrap? */
	const int inside = !wrapped && cb <= NGS_POS_MASK && pos >= sg.clip_lo && cb <= sg.lim;
	special = 1;
	if (inside && a0 == pos) {
		const int far = cb - pos > NGS_NEAR_H;            /* a delta further than H from POS may fall outside its owner's view */
		if (na == 1 && b0 - a0 < (1u << 30)) { y = (NGS_ENT_SIMPLE << 30) | (b0 - a0); special = far; return 1; }
		if (na == 2 && b0 - a0 < 4096u && a1 - b0 < 64u && b1 - a1 < 4096u) {
			y = (NGS_ENT_PAIR << 30) | (b0 - a0) | ((a1 - b0) << 12) | ((b1 - a1) << 18);
			special = far; return 1;
		}
	}
	y = (NGS_ENT_CHAIN << 30) | slot_off;
	return 1;
}

/* decode one (already validated) record at offset o and emit its deltas with weight `sign` */
template <class Mem, class Sink>
NGS_HD void ngs_decode_record(const ngs_walk_cfg &cfg, const Mem &mem, int64_t o, const ngs_hdr &h,
                              Sink &sink, int sign, ngs_counts &cnt)
{
	cnt.records += 1; cnt.cigar_words += h.n_cigar;
	{	/* sorted input: once refID (or, for a shard, the position inside the slice's last contig) has moved on,
		 * everything before is complete -- streamed finalize and ngscov_slice_done rely on this mark */
		uint32_t r1 = (uint32_t)(h.ref_id + 1);
		if (h.ref_id == cfg.end_ctg && (uint32_t)h.pos >= cfg.end_pos) r1 += 1;
		if (r1 > cnt.max_ref_p1) cnt.max_ref_p1 = r1;
	}
	if ((h.flag & cfg.flag_mask) || h.ref_id < 0 || (uint32_t)h.pos >= (1u << 29)) return;   /* bam_fetch's is_overlap compares as uint32 (samtools bam_index.c:571-576) */
	cnt.kept += 1;
	const int32_t si = cfg.ctg2seg[h.ref_id];
	if (si < 0 || h.n_cigar == 0) return;
	const ngs_seg &sg = cfg.seg[si];                     /* only base / clip_lo / lim / hard_end are touched */
	uint32_t p = (uint32_t)h.pos;
	int64_t co = o + 36 + (int64_t)h.l_qname;
	for (uint32_t i = 0; i < h.n_cigar; ++i, co += 4) {
		const uint32_t w = mem.u32(co), op = w & 0xf, len = w >> 4;
		if (op == 0) {                                   /* M: Start[p]++, End[p+len]++ */
			uint32_t a = p & NGS_POS_MASK;
			p += len;
			uint32_t b = p & NGS_POS_MASK;
			if (a == b) continue;                        /* net-zero pair of keys: transparent to the sweep */
			if (b < a) { sink.flag_range_error(); continue; }
			if (b > sg.lim) {
				if (sg.hard_end) { sink.flag_range_error(); continue; }
				b = sg.lim;
			}
			if (a < sg.clip_lo) a = sg.clip_lo;
			if (a >= b) continue;
			cnt.m_ops += 1;
			sink.add2(sg.base + (a - sg.clip_lo), sg.base + (b - sg.clip_lo), sign);
		} else if (op == 2 || op == 3) p += len;         /* D, N advance; I S H P = X B do not */
	}
}

/* Walk the record chain from offset o while records START before `hi`.  The exit code is
 *   >= hi            offset of the first record start at or beyond hi (may exceed the chunk size)
 *   NGS_EXIT_PARTIAL the record at that offset has its header/CIGAR cut by the chunk end
 *   NGS_EXIT_INVALID a record failed validation
 * The walk is split into begin/step so that a warp can advance all of its chains in lock step (one
 * record per lane per iteration) and stay converged; ngs_walk is the plain sequential composition. */
typedef struct {
	int64_t o;          /* offset of the record whose header is in h */
	int64_t result;     /* exit code once active == 0 */
	ngs_hdr h;
	int active;
} ngs_walk_state;

template <class Mem>
NGS_HD void ngs_walk_begin(const Mem &mem, int64_t o, int64_t hi, ngs_walk_state &st)
{
	st.o = o; st.active = 0; st.result = o;
	if (o >= hi) return;
	if (o + 36 > mem.size()) { st.result = NGS_EXIT_PARTIAL(o); return; }
	mem.hdr_lite(o, st.h);
	st.active = 1;
}

template <class Mem, class Sink>
NGS_HD void ngs_walk_step(const ngs_walk_cfg &cfg, const Mem &mem, int64_t hi, Sink &sink, int sign, ngs_counts &cnt, ngs_walk_state &st)
{
	const int64_t n = mem.size(), o = st.o;
	const ngs_hdr h = st.h;
	if (!ngs_hdr_valid(h, cfg.n_ref)) { st.active = 0; st.result = NGS_EXIT_INVALID; return; }
	if (o + 36 + (int64_t)h.l_qname + 4ll * h.n_cigar > n) { st.active = 0; st.result = NGS_EXIT_PARTIAL(o); return; }
	/* the next header is requested before this record is decoded: two loads in flight per chain */
	const int64_t nx = o + 4 + (int64_t)h.block_size;
	const bool more = nx < hi, have = more && nx + 36 <= n;
	if (have) mem.hdr_lite(nx, st.h);
	ngs_decode_record(cfg, mem, o, h, sink, sign, cnt);
	st.o = nx;
	if (!more) { st.active = 0; st.result = nx; }
	else if (!have) { st.active = 0; st.result = NGS_EXIT_PARTIAL(nx); }
}

template <class Mem, class Sink>
NGS_HD int64_t ngs_walk(const ngs_walk_cfg &cfg, const Mem &mem, int64_t o, int64_t hi,
                        Sink &sink, int sign, ngs_counts &cnt)
{
	ngs_walk_state st;
	ngs_walk_begin(mem, o, hi, st);
	while (st.active) ngs_walk_step(cfg, mem, hi, sink, sign, cnt, st);
	return st.result;
}

/* Speculation: does a record chain of `depth` plausible records start at offset o? */
template <class Mem>
NGS_HD int ngs_candidate_ok(const ngs_walk_cfg &cfg, const Mem &mem, const uint32_t *ref_len, int64_t o)
{
	const int64_t n = mem.size();
	int64_t c = o;
	for (int depth = 0; depth < 3; ++depth) {
		ngs_hdr h;
		if (c + 36 > n) break;                                   /* chain leaves the chunk: accept */
		mem.hdr(c, h);
		if (!ngs_hdr_plausible(h, cfg.n_ref, ref_len)) return 0;
		const int64_t nul = c + 36 + (int64_t)h.l_qname - 1;
		if (nul < n && mem.u8(nul) != 0) return 0;
		c += 4 + (int64_t)h.block_size;
	}
	return 1;
}

/* Speculative entry search in [lo, hi): first offset whose record and the two records chained after
 * it all look plausible.  NGS_NONE if there is none (or it cannot be checked inside the chunk).
 * mem.next_candidate(o, end, n_ref) returns the first offset in [o, end) whose block_size and refID
 * fields are in range (a tight byte scan), or NGS_NONE. */
template <class Mem>
NGS_HD int64_t ngs_find_entry(const ngs_walk_cfg &cfg, const Mem &mem, const uint32_t *ref_len, int64_t lo, int64_t hi)
{
	const int64_t n = mem.size();
	const int64_t end = hi < n - 35 ? hi : n - 35;            /* a candidate needs its 36 header bytes inside the chunk */
	int64_t o = lo;
	for (;;) {
		o = mem.next_candidate(o, end, (uint32_t)cfg.n_ref);
		if (o == NGS_NONE) return NGS_NONE;
		if (ngs_candidate_ok(cfg, mem, ref_len, o)) return o;
		++o;
	}
}

#endif
