/* csrc/ngscov_cover.cuh -- the "cover" pipeline of libngscov_cuda (round 2): the difference array is never
 * materialised in HBM for ordinary records.
 *
 *   K1  k_chain_classify   speculative per-sub-chunk block_size chain walk (as k_chain_walk) that also decodes core
 *                          fields + CIGAR of every record while its cache line is there, and leaves one 8-byte entry
 *                          per contributing record in the row of its sub-chunk (ngs_row / entry format: ngscov_hd.h).
 *                          Replaces bam_read1 framing (samtools bam.c:191-223) + fetch_func (bam2depth.c:86-110,
 *                          bam2wig.c:84-109).
 *       k_verify_*_k       the boundary check of k_verify_*, plus a running max of the record sort keys: gives every
 *                          row the largest key seen up to it (rkmax, the search array of k_tile_rows) and detects
 *                          input that is not coordinate sorted.
 *       k_special_reds     the rare records: CHAIN entries (decoded from the stream while the chunk is resident) and
 *                          deltas further than NGS_NEAR_H from their record's POS -> REDs into the global int32
 *                          array + far flag of the tile; also folds the per-row counters.
 *   K2-K5  k_cover         one CTA owns a tile of NGS_CT positions: gathers the entries of the rows that can reach it
 *                          (k_tile_rows: two binary searches per tile), accumulates their +1/-1 in shared memory,
 *                          merges the tile of the far array when flagged, scans (decoupled look-back over one 64-bit
 *                          word: sum | last boundary), detects run boundaries, adds the window bins (both `overlap`
 *                          variants, bam2depth.c:132-176 / bam2wig.c:131-175) and writes the bedGraph lines
 *                          (bam2depth.c:208-234) straight into the output text (second look-back over byte counts).
 *                          Replaces insertInt2Hash + hashtbl (bam2depth.c:72-84, hashtbl.c) + union_hashed_keys +
 *                          the sweep of hash2BedGraph + its fprintf.
 */
#pragma once
#include "ngscov_kernels.cuh"

#ifndef NGS_COVER_THREADS
#define NGS_COVER_THREADS 256
#endif
#ifndef NGS_COVER_MINB
#define NGS_COVER_MINB 3
#endif
#define NGS_WATCHDOG_POLLS (1u << 26)      /* look-back polls of one entry before the wait is abandoned (seconds) */

template <class Mem>
struct CigReader {
	const Mem &mem; int64_t co;
	__device__ __forceinline__ uint32_t operator()(uint32_t i) const { return mem.u32(co + 4ll * i); }
};

/* ---- bam_sliding_count (reference bam_sliding_count.c:84-124) --------------------------------
 * The same table and the same verification serve a second tool: per window the number of mapped reads STARTING in it,
 * the sum of their l_qseq and the number of C / G bases in their SEQ.  An entry is x = POS | contig-delta << 29,
 * y = GC count (16 bits, the reference's unsigned short) << 16 | l_qseq (16 bits).  Records that do not fit (l_qseq >=
 * 65535, POS < 0, more than seven contigs in a row) flag their row for the sequential path of k_slide_accumulate. */
template <class Mem>
__device__ __forceinline__ uint32_t ngs_seq_gc(const Mem &mem, int64_t so, int32_t l_seq)
{
	/* 4-bit codes, two per byte (bam1_seqi): count the nibbles equal to 2 (C) or 4 (G), eight per 32-bit word */
	uint32_t gc = 0;
	const int64_t nb = ((int64_t)l_seq + 1) >> 1;
	for (int64_t i = 0; i < nb; i += 4) {
		uint32_t w = mem.u32(so + i);
		const int64_t left = nb - i;
		if (left < 4) w &= (1u << (8 * (uint32_t)left)) - 1u;      /* bytes past the sequence: code 0, never counted */
		uint32_t a = w ^ 0x22222222u, b = w ^ 0x44444444u;          /* a nibble is zero where the code is 2 / 4 */
		a |= a >> 1; a |= a >> 2; b |= b >> 1; b |= b >> 2;
		gc += 16u - (uint32_t)__popc(a & 0x11111111u) - (uint32_t)__popc(b & 0x11111111u);
	}
	return gc;
}
/* returns 1 = entry y, 2 = record needs the sequential path, 0 = dropped (tid < 0 or BAM_FUNMAP, :96-97) */
template <class Mem, class Cnt>
__device__ __forceinline__ int ngs_slide_record(const ngs_walk_cfg &cfg, const Mem &mem, int64_t o, const ngs_hdr &h, uint32_t &y, Cnt &cnt)
{
	cnt.records += 1; cnt.cigar_words += h.n_cigar;
	const uint32_t r1 = (uint32_t)(h.ref_id + 1);
	if (r1 > cnt.max_ref_p1) cnt.max_ref_p1 = r1;
	if (h.ref_id < 0 || (h.flag & 4u)) return 0;
	if (cfg.reg_ref >= 0) {                          /* -r: what bam_fetch would not hand out is not counted (:393-409) */
		const CigReader<Mem> cig = { mem, o + 36 + (int64_t)h.l_qname };
		if (!ngs_region_pass(cfg, h.ref_id, h.pos, h.n_cigar, cig)) return 0;
	}
	cnt.kept += 1;
	if (h.pos < 0 || h.l_seq >= 0xFFFF) return 2;
	const uint32_t gc = ngs_seq_gc(mem, o + 36 + (int64_t)h.l_qname + 4ll * h.n_cigar, h.l_seq) & 0xFFFFu;
	y = (gc << 16) | (uint32_t)h.l_seq;
	return 1;
}

/* walk + classify the records starting in [entry, hi): entries -> row[], metadata -> m */
template <bool SLIDE, class Mem>
__device__ __forceinline__ long long chain_classify(const ngs_walk_cfg &cfg, const Mem &mem, long long entry, int64_t hi,
                                                    uint2 *row, ngs_row &m)
{
	const int64_t n = mem.size();
	int64_t o = entry;
	ngs_counts32 cnt = { 0, 0, 0, 0, 0 };
	uint64_t fk = NGS_KEY_NONE, prev = 0;
	uint32_t count = 0, special = 0; int32_t ctg0 = -1;
	long long result;
	if (o >= hi) result = o;
	else if (o + 36 > n) result = NGS_EXIT_PARTIAL(o);
	else {
		ngs_hdr h;
		mem.hdr_lite(o, h);
		for (;;) {
			if (!ngs_hdr_valid(h, cfg.n_ref)) { result = NGS_EXIT_INVALID; break; }
			if (o + 36 + (int64_t)h.l_qname + 4ll * h.n_cigar + (SLIDE ? ((int64_t)h.l_seq + 1) / 2 : 0) > n) { result = NGS_EXIT_PARTIAL(o); break; }
			/* the next header is requested before this record is classified: the chain of dependent loads stays as
			 * short as in the framing-only walk, the CIGAR loads ride along */
			const int64_t nx = o + 4 + (int64_t)h.block_size;
			const bool more = nx < hi, have = more && nx + 36 <= n;
			ngs_hdr hn;
			if (have) mem.hdr_lite(nx, hn);
			const uint64_t key = ngs_sort_key(h.ref_id, h.pos);
			if (fk == NGS_KEY_NONE) fk = key;
			if (key < prev) special |= NGS_ROW_UNSORTED;
			prev = key;
			uint32_t y = 0; int sp = 0;
			CigReader<Mem> cig = { mem, o + 36 + (int64_t)h.l_qname };
			int got;
			if (SLIDE) { got = ngs_slide_record(cfg, mem, o, h, y, cnt); if (got == 2) { special |= NGS_ROW_OVERFLOW; got = 0; } }
			else got = ngs_classify_record(cfg, h, (uint32_t)(o - entry), cig, y, sp, cnt);
			if (got) {
				if (ctg0 < 0) ctg0 = h.ref_id;
				uint32_t cd = (uint32_t)(h.ref_id - ctg0);
				if (cd > NGS_ENT_CDELTA_MAX) {
					if (SLIDE) special |= NGS_ROW_OVERFLOW;
					else { cd = 7u; y = (NGS_ENT_CHAIN << 30) | (uint32_t)(o - entry); sp = 1; }
				}
				if (count < NGS_ROW_SLOTS) row[count] = make_uint2((uint32_t)h.pos | (cd << 29), y);
				++count; special += (uint32_t)sp;
			}
			if (!more) { result = nx; break; }
			if (!have) { result = NGS_EXIT_PARTIAL(nx); break; }
			h = hn; o = nx;
		}
	}
	if (count > NGS_ROW_SLOTS || (special & NGS_ROW_OVERFLOW)) { count = 0; special = (special & NGS_ROW_UNSORTED) | NGS_ROW_OVERFLOW; }
	m.fk = fk; m.lk = prev; m.count = count; m.special = special; m.ctg = ctg0;
	m.nrec = cnt.records; m.nkept = cnt.kept; m.nmops = cnt.m_ops; m.ncig = cnt.cigar_words; m.max_ref_p1 = cnt.max_ref_p1;
	m.pad0 = 0; m.pad1 = 0;
	return result;
}

__device__ __forceinline__ void store_row(ngs_row *dst, const ngs_row &m)
{
	const uint4 *s = reinterpret_cast<const uint4 *>(&m);
	uint4 *d = reinterpret_cast<uint4 *>(dst);
	d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
}

#ifndef NGS_K1_MINB
#define NGS_K1_MINB 4
#endif
#ifndef NGS_K1_BUF
#define NGS_K1_BUF 8
#endif                                 /* entries a lane buffers in shared memory between two flushes: 8 x 8 B = two full 32-byte sectors
                                          (measured on the 30x workload: 4 -> 11.57, 8 -> 11.29, 16 -> 11.71 ms) */

/* K1.  One thread per sub-chunk; the 32 walks of a warp advance in lock step (one record per lane per iteration) so that
 * the table entries can leave the SM coalesced: every lane appends its entries to a small shared-memory buffer and every
 * NGS_K1_BUF iterations the warp writes the 32 buffers out, 32 / NGS_K1_BUF rows per store instruction, whole sectors.
 * Written straight from the lanes (8 bytes per lane into 32 different rows per store) the entries cost 4.1 of the
 * kernel's 12.9 ms on the 30x workload (618 M partial-sector write transactions; 8.8 ms with the stores removed, 5.9 ms
 * for framing alone): staged, the kernel takes 11.3 ms. */
template <bool SLIDE>
__global__ void __launch_bounds__(256, NGS_K1_MINB) k_chain_classify(const __grid_constant__ IngestArgs A)
{
	__shared__ uint2 s_buf[8][32][NGS_K1_BUF + 1];           /* + 1: lanes hit different banks */
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const bool live = c < A.n_sub;
	DevMem mem = { A.base, A.n };
	const int64_t n = A.n;
	const int64_t lo = live ? (c << A.shift) : 0;
	int64_t hi = live ? lo + ((int64_t)1 << A.shift) : 0;
	if (hi > A.n) hi = A.n;
	long long entry = NGS_NONE;
	bool searching = live && c != 0;
	if (live && c == 0) {
		const long long e0 = A.entry0_from_state ? A.st->entry_next : A.fixed_entry0;
		entry = e0 < hi ? e0 : NGS_NONE;
	}
	{
		const int64_t end = hi < A.n - 35 ? hi : A.n - 35;
		int64_t o = lo;
		while (__any_sync(full, searching)) {
			if (searching) { o = mem.next_candidate(o, end, (uint32_t)A.cfg.n_ref); if (o == NGS_NONE) searching = false; }
			__syncwarp();
			if (searching) {
				if (ngs_candidate_ok(A.cfg, mem, A.ref_len, o)) { entry = o; searching = false; } else ++o;
			}
		}
	}
	/* ---- the walk */
	ngs_counts32 cnt = { 0, 0, 0, 0, 0 };
	uint64_t fk = NGS_KEY_NONE, prev = 0;
	uint32_t count = 0, nbuf = 0, special = 0; int32_t ctg0 = -1;
	long long result = 0;
	int64_t o = entry;
	ngs_hdr h;
	bool active = false;
	if (entry != NGS_NONE) {
		if (o >= hi) result = o;
		else if (o + 36 > n) result = NGS_EXIT_PARTIAL(o);
		else { mem.hdr_lite(o, h); active = true; }
	}
	uint2 *const out0 = A.ent + (size_t)(c - lane) * NGS_ROW_SLOTS;      /* row of lane 0 of this warp */
	auto flush = [&]() {
		__syncwarp();
		#pragma unroll 4
		for (int j = 0; j < 32; j += 32 / NGS_K1_BUF) {
			const int src = j + lane / NGS_K1_BUF, k = lane % NGS_K1_BUF;
			const uint32_t nb = __shfl_sync(full, nbuf, src), base = __shfl_sync(full, count - nbuf, src);
			if ((uint32_t)k < nb && base + (uint32_t)k < NGS_ROW_SLOTS) out0[(size_t)src * NGS_ROW_SLOTS + base + k] = s_buf[warp][src][k];
		}
		nbuf = 0;
		__syncwarp();
	};
	int it = 0;
	while (__any_sync(full, active)) {
		if (active) {
			if (!ngs_hdr_valid(h, A.cfg.n_ref)) { result = NGS_EXIT_INVALID; active = false; }
			else if (o + 36 + (int64_t)h.l_qname + 4ll * h.n_cigar + (SLIDE ? ((int64_t)h.l_seq + 1) / 2 : 0) > n) { result = NGS_EXIT_PARTIAL(o); active = false; }
			else {
				/* the next header is requested before this record is classified: the chain of dependent loads stays as
				 * short as in a framing-only walk, the CIGAR loads ride along */
				const int64_t nx = o + 4 + (int64_t)h.block_size;
				const bool more = nx < hi, have = more && nx + 36 <= n;
				ngs_hdr hn;
				if (have) mem.hdr_lite(nx, hn);
				const uint64_t key = ngs_sort_key(h.ref_id, h.pos);
				if (fk == NGS_KEY_NONE) fk = key;
				if (key < prev) special |= NGS_ROW_UNSORTED;
				prev = key;
				uint32_t y = 0; int sp = 0;
				CigReader<DevMem> cig = { mem, o + 36 + (int64_t)h.l_qname };
				int got;
				if (SLIDE) { got = ngs_slide_record(A.cfg, mem, o, h, y, cnt); if (got == 2) { special |= NGS_ROW_OVERFLOW; got = 0; } }
				else got = ngs_classify_record(A.cfg, h, (uint32_t)(o - entry), cig, y, sp, cnt);
				if (got) {
					if (ctg0 < 0) ctg0 = h.ref_id;
					uint32_t cd = (uint32_t)(h.ref_id - ctg0);
					if (cd > NGS_ENT_CDELTA_MAX) {
						if (SLIDE) special |= NGS_ROW_OVERFLOW;
						else { cd = 7u; y = (NGS_ENT_CHAIN << 30) | (uint32_t)(o - entry); sp = 1; }
					}
					s_buf[warp][lane][nbuf++] = make_uint2((uint32_t)h.pos | (cd << 29), y);
					++count; special += (uint32_t)sp;
				}
				if (!more) { result = nx; active = false; }
				else if (!have) { result = NGS_EXIT_PARTIAL(nx); active = false; }
				else { h = hn; o = nx; }
			}
		}
		if (++it == NGS_K1_BUF) { flush(); it = 0; }
	}
	flush();
	if (count > NGS_ROW_SLOTS || (special & NGS_ROW_OVERFLOW)) { count = 0; special = (special & NGS_ROW_UNSORTED) | NGS_ROW_OVERFLOW; }
	ngs_row m;
	m.fk = fk; m.lk = prev; m.count = count; m.special = special; m.ctg = ctg0;
	m.nrec = cnt.records; m.nkept = cnt.kept; m.nmops = cnt.m_ops; m.ncig = cnt.cigar_words; m.max_ref_p1 = cnt.max_ref_p1;
	m.pad0 = 0; m.pad1 = 0;
	if (live) { A.entry[c] = entry; A.exit[c] = entry != NGS_NONE ? result : 0; store_row(A.rows + c, m); }
	const unsigned w = __reduce_add_sync(full, entry != NGS_NONE ? 1u : 0u);
	if (lane == 0 && w) atomicAdd(&A.st->spec_walks, (unsigned long long)w);
}

/* Rare path, single thread: zip the true chain forward from the first mismatching sub-chunk until it meets an
 * already-framed entry, rewriting entry / exit / row on the way (k_repair_chain for the table) */
template <bool SLIDE>
__global__ void k_repair_classify(IngestArgs A, long long first_bad, const long long *bexcl)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	DevMem mem = { A.base, A.n };
	long long cur = bexcl[first_bad / NGS_VER_PER_BLOCK];
	for (long long k = (first_bad / NGS_VER_PER_BLOCK) * NGS_VER_PER_BLOCK; k < first_bad; ++k) {
		const long long x = exit_num(A.exit[k]);
		cur = x > cur ? x : cur;
	}
	ngs_row empty;
	empty.fk = NGS_KEY_NONE; empty.lk = 0; empty.count = 0; empty.special = 0; empty.ctg = -1;
	empty.nrec = empty.nkept = empty.nmops = empty.ncig = empty.max_ref_p1 = 0; empty.pad0 = empty.pad1 = 0;
	long long c = first_bad, repaired = 0;
	for (;;) {
		long long cc = cur >= A.n ? A.n_sub : (cur < 0 ? 0 : (cur >> A.shift));
		if (cc > A.n_sub) cc = A.n_sub;
		for (long long k = c; k < cc; ++k)              /* sub-chunks the true chain jumps over hold no record start */
			if (A.entry[k] != NGS_NONE) { A.entry[k] = NGS_NONE; A.exit[k] = 0; A.rows[k] = empty; ++repaired; }
		if (cc >= A.n_sub) break;
		if (A.entry[cc] == cur) break;                  /* converged */
		long long hi = (cc + 1) << A.shift; if (hi > A.n) hi = A.n;
		ngs_row m;
		const long long x = chain_classify<SLIDE>(A.cfg, mem, cur, hi, A.ent + (size_t)cc * NGS_ROW_SLOTS, m);
		A.entry[cc] = cur; A.exit[cc] = x; A.rows[cc] = m; ++repaired;
		c = cc + 1;
		cur = (x == NGS_EXIT_INVALID || x >= NGS_EXIT_INF) ? A.n : x;
	}
	atomicAdd(&A.st->spec_repaired, (unsigned long long)repaired);
}

/* ---- verification with keys: the three kernels of k_verify_* carrying a second running max ------------------- */
__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long *sm)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
	const int w = threadIdx.x >> 5;
	__syncthreads();
	if ((threadIdx.x & 31) == 0) sm[w] = v;
	__syncthreads();
	unsigned long long r = sm[0];
	for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = sm[i] > r ? sm[i] : r;
	return r;
}

__global__ void __launch_bounds__(256) k_verify_blockmax_k(const long long *exit, const ngs_row *rows, int64_t n_sub, long long *bmax, unsigned long long *kbmax)
{
	__shared__ long long sm[8];
	__shared__ unsigned long long smk[8];
	const int64_t b0 = (int64_t)blockIdx.x * NGS_VER_PER_BLOCK;
	long long m = 0; unsigned long long km = 0;
	#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const int64_t c = b0 + i * 256 + threadIdx.x;
		if (c < n_sub) {
			long long x = exit_num(exit[c]); m = x > m ? x : m;
			const unsigned long long lk = rows[c].lk; km = lk > km ? lk : km;
		}
	}
	m = block_max_ll(m, sm);
	km = block_max_u64(km, smk);
	if (threadIdx.x == 0) { bmax[blockIdx.x] = m; kbmax[blockIdx.x] = km; }
}

/* single block: exclusive running max over the block maxima (exits seeded with the chunk's first entry, keys with
 * the committed key of the previous chunks) */
__global__ void __launch_bounds__(1024) k_verify_scan_k(long long *bmax, unsigned long long *kbmax, int64_t n_blocks, ngs_dev_state *st,
                                                        long long fixed_entry0, int entry0_from_state)
{
	__shared__ long long sm[32];
	__shared__ unsigned long long smk[32];
	__shared__ long long carry_s;
	__shared__ unsigned long long kcarry_s;
	if (threadIdx.x == 0) {
		carry_s = entry0_from_state ? st->entry_next : fixed_entry0;
		kcarry_s = st->key_seed;
		st->n_bad = 0; st->first_bad_key = ~0ull; st->unsorted = 0;
	}
	__syncthreads();
	for (int64_t b0 = 0; b0 < n_blocks; b0 += 1024) {
		const int64_t b = b0 + threadIdx.x;
		long long v = b < n_blocks ? bmax[b] : 0;
		unsigned long long kv = b < n_blocks ? kbmax[b] : 0;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			long long t = __shfl_up_sync(0xffffffffu, v, o); unsigned long long kt = __shfl_up_sync(0xffffffffu, kv, o);
			if ((threadIdx.x & 31) >= o) { if (t > v) v = t; if (kt > kv) kv = kt; }
		}
		if ((threadIdx.x & 31) == 31) { sm[threadIdx.x >> 5] = v; smk[threadIdx.x >> 5] = kv; }
		__syncthreads();
		long long pre = carry_s; unsigned long long kpre = kcarry_s;
		for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) { pre = sm[w] > pre ? sm[w] : pre; kpre = smk[w] > kpre ? smk[w] : kpre; }
		const long long incl = v > pre ? v : pre; const unsigned long long kincl = kv > kpre ? kv : kpre;
		const long long prev = __shfl_up_sync(0xffffffffu, v, 1); const unsigned long long kprev = __shfl_up_sync(0xffffffffu, kv, 1);
		const long long excl = (threadIdx.x & 31) ? (prev > pre ? prev : pre) : pre;
		const unsigned long long kexcl = (threadIdx.x & 31) ? (kprev > kpre ? kprev : kpre) : kpre;
		if (b < n_blocks) { bmax[b] = excl; kbmax[b] = kexcl; }
		__syncthreads();
		if (threadIdx.x == 1023) { carry_s = incl; kcarry_s = kincl; }
		__syncthreads();
	}
	if (threadIdx.x == 0) { st->chunk_exit = carry_s; st->key_chunk_max = kcarry_s; }
}

__global__ void __launch_bounds__(256) k_verify_check_k(const long long *entry, const long long *exit, const ngs_row *rows, int64_t n_sub,
                                                        const long long *bexcl, const unsigned long long *kbexcl, unsigned long long *rkmax,
                                                        int shift, int64_t n, ngs_dev_state *st)
{
	__shared__ long long sm[8];
	__shared__ unsigned long long smk[8];
	const int64_t c0 = (int64_t)blockIdx.x * NGS_VER_PER_BLOCK + (int64_t)threadIdx.x * 4;
	long long ex[4], en[4]; unsigned long long fk[4], lk[4]; unsigned int sp[4];
	#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const int64_t c = c0 + i;
		ex[i] = c < n_sub ? exit_num(exit[c]) : 0;
		en[i] = c < n_sub ? entry[c] : NGS_NONE;
		fk[i] = NGS_KEY_NONE; lk[i] = 0; sp[i] = 0;
		if (c < n_sub) { const uint4 a = reinterpret_cast<const uint4 *>(rows + c)[0], b = reinterpret_cast<const uint4 *>(rows + c)[1];
			fk[i] = (unsigned long long)a.x | ((unsigned long long)a.y << 32); lk[i] = (unsigned long long)a.z | ((unsigned long long)a.w << 32); sp[i] = b.y; }
	}
	long long tmax = ex[0];
	tmax = ex[1] > tmax ? ex[1] : tmax; tmax = ex[2] > tmax ? ex[2] : tmax; tmax = ex[3] > tmax ? ex[3] : tmax;
	unsigned long long ktmax = lk[0];
	ktmax = lk[1] > ktmax ? lk[1] : ktmax; ktmax = lk[2] > ktmax ? lk[2] : ktmax; ktmax = lk[3] > ktmax ? lk[3] : ktmax;
	long long v = tmax; unsigned long long kv = ktmax;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		long long t = __shfl_up_sync(0xffffffffu, v, o); unsigned long long kt = __shfl_up_sync(0xffffffffu, kv, o);
		if ((threadIdx.x & 31) >= o) { if (t > v) v = t; if (kt > kv) kv = kt; }
	}
	if ((threadIdx.x & 31) == 31) { sm[threadIdx.x >> 5] = v; smk[threadIdx.x >> 5] = kv; }
	__syncthreads();
	long long pre = bexcl[blockIdx.x]; unsigned long long kpre = kbexcl[blockIdx.x];
	for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) { pre = sm[w] > pre ? sm[w] : pre; kpre = smk[w] > kpre ? smk[w] : kpre; }
	const long long prev = __shfl_up_sync(0xffffffffu, v, 1); const unsigned long long kprev = __shfl_up_sync(0xffffffffu, kv, 1);
	long long run = (threadIdx.x & 31) ? (prev > pre ? prev : pre) : pre;
	unsigned long long krun = (threadIdx.x & 31) ? (kprev > kpre ? kprev : kpre) : kpre;
	unsigned int unsorted = 0;
	#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const int64_t c = c0 + i;
		if (c < n_sub) {
			const long long expected = run;
			const int64_t lo = c << shift;
			int64_t hi = lo + ((int64_t)1 << shift);
			if (hi > n) hi = n;
			const bool exp_in = expected < hi;
			int kind = -1;
			if (exp_in && c > 0 && expected < lo) kind = 1;                  /* a verified chain ended early */
			else if (exp_in ? (en[i] != expected) : (en[i] != NGS_NONE)) kind = 0;
			else if (en[i] != NGS_NONE && exit[c] == NGS_EXIT_INVALID) kind = 1;
			if (kind >= 0) {
				atomicAdd(&st->n_bad, 1ull);
				atomicMin(&st->first_bad_key, (unsigned long long)c * 2 + (unsigned)kind);
			}
			run = ex[i] > run ? ex[i] : run;
			if ((sp[i] & NGS_ROW_UNSORTED) || (fk[i] != NGS_KEY_NONE && fk[i] < krun)) unsorted = 1;
			krun = lk[i] > krun ? lk[i] : krun;
			rkmax[c] = krun;
		}
	}
	if (__any_sync(0xffffffffu, unsorted) && (threadIdx.x & 31) == 0) st->unsorted = 1;
}

/* ---- the rare records ------------------------------------------------------------------------------------- */
struct FarSink {                     /* RED into the global array + far flag of the element's tile */
	int *diff; uint32_t *far_flag; unsigned int *flags; int mark;
	__device__ __forceinline__ void one(uint64_t idx, int v) { atomicAdd(diff + idx, v); if (mark) far_flag[idx / NGS_CT] = 1u; }
	__device__ __forceinline__ void add2(uint64_t ia, uint64_t ib, int sign) { one(ia, sign); one(ib, -sign); }
	__device__ __forceinline__ void flag_range_error() { atomicOr(flags, NGS_FLAG_ERANGE); }
};

/* mode 0: CHAIN entries (all their deltas) and the far deltas of SIMPLE / PAIR entries, rows that overflowed; folds
 *         the per-row counters into the striped counters (cover pipeline, once per verified chunk)
 * mode 1: every delta of every entry (input that is not coordinate sorted: the classic scan follows) + counters
 * mode 2: only the near deltas of SIMPLE / PAIR entries (catch-up for rows that already had a mode-0 pass when the
 *         input turned out to be unsorted; needs no stream) */
__global__ void __launch_bounds__(256) k_special_reds(const __grid_constant__ IngestArgs A, int mode)
{
	const unsigned full = 0xffffffffu;
	if (mode != 2) {
		const ngs_dev_state *st = A.st;
		if (st->n_bad != 0 || (st->flags & NGS_FLAG_EFORMAT)) return;      /* launched optimistically behind the verification */
	}
	const int lane = threadIdx.x & 31;
	const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5, wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	DevMem mem = { A.base, A.n };
	FarSink sink = { A.diff, A.far_flag, &A.st->flags, mode == 0 ? 1 : 0 };
	unsigned int s_rec = 0, s_kept = 0, s_mops = 0, s_cig = 0, s_max = 0, s_sp = 0;
	/* rows are dealt out round robin (lane l of warp w takes rows w + l * n_warps of a stripe): special rows come in
	 * bursts of consecutive rows -- the records hanging into a region shard from the left -- and a burst inside one
	 * warp's 32 rows was processed row after row (0.19 ms per shard at N > 1) */
	for (int64_t base = 0; base < A.n_sub; base += n_warps * 32) {
		const int64_t c = base + (int64_t)lane * n_warps + wid;
		uint4 b = make_uint4(0, 0, 0xffffffffu, 0), d = make_uint4(0, 0, 0, 0);
		if (c < A.n_sub) { b = reinterpret_cast<const uint4 *>(A.rows + c)[1]; d = reinterpret_cast<const uint4 *>(A.rows + c)[2]; }
		/* b = count, special, ctg, nrec ; d = nkept, nmops, ncig, max_ref_p1 */
		s_rec += b.w; s_kept += d.x; s_mops += d.y; s_cig += d.z; s_max = d.w > s_max ? d.w : s_max; s_sp += b.y & 0xffffu;
		const bool need = mode == 0 ? (b.y & ~NGS_ROW_UNSORTED) != 0 : (mode == 1 ? (b.x != 0 || (b.y & NGS_ROW_OVERFLOW)) : b.x != 0);
		unsigned todo = __ballot_sync(full, need);
		while (todo) {
			const int src = __ffs(todo) - 1; todo &= todo - 1;
			const unsigned int count = __shfl_sync(full, b.x, src), special = __shfl_sync(full, b.y, src);
			const int ctg0 = (int)__shfl_sync(full, b.z, src);
			const int64_t row = base + (int64_t)src * n_warps + wid;
			if (special & NGS_ROW_OVERFLOW) {
				if (mode != 2 && lane == 0) {         /* the whole row goes the classic way: sequential walk, run-length combined REDs */
					long long hi = (row + 1) << A.shift; if (hi > A.n) hi = A.n;
					ngs_counts dummy = { 0, 0, 0, 0, 0 };
					ngs_walk(A.cfg, mem, A.entry[row], hi, sink, +1, dummy);
				}
				continue;
			}
			const long long ent0 = mode != 2 ? A.entry[row] : 0;
			for (unsigned int k = lane; k < count; k += 32) {
				const uint2 e = A.ent[(size_t)row * NGS_ROW_SLOTS + k];
				const uint32_t kind = e.y >> 30;
				if (kind == NGS_ENT_CHAIN) {
					if (mode == 2) continue;
					const int64_t o = ent0 + (int64_t)(e.y & 0x3FFFFFFFu);
					ngs_hdr h; mem.hdr_lite(o, h);
					ngs_counts dummy = { 0, 0, 0, 0, 0 };
					ngs_decode_record(A.cfg, mem, o, h, sink, +1, dummy);
					continue;
				}
				const uint32_t pos = e.x & 0x1FFFFFFFu;
				uint32_t ia[2], ib[2];
				const int ni = ngs_entry_intervals(pos, e.y, ia, ib);
				if (mode == 0 && ib[ni - 1] - pos <= NGS_NEAR_H) continue;       /* nothing of it can be far */
				const ngs_seg &sg = A.cfg.seg[A.cfg.ctg2seg[ctg0 + (int)(e.x >> 29)]];
				for (int j = 0; j < ni; ++j) {
					const int na = ngs_delta_is_near(sg.clip_lo, pos, ia[j]), nb = ngs_delta_is_near(sg.clip_lo, pos, ib[j]);
					if (mode == 1 || (mode == 0 ? !na : na)) sink.one(sg.base + (ia[j] - sg.clip_lo), +1);
					if (mode == 1 || (mode == 0 ? !nb : nb)) sink.one(sg.base + (ib[j] - sg.clip_lo), -1);
				}
			}
		}
	}
	if (mode == 2) return;
	const unsigned v0 = __reduce_add_sync(full, s_rec), v1 = __reduce_add_sync(full, s_kept), v2 = __reduce_add_sync(full, s_mops),
	               v3 = __reduce_add_sync(full, s_cig), v4 = __reduce_max_sync(full, s_max), v5 = __reduce_add_sync(full, s_sp);
	if (lane == 0 && v0) {
		unsigned long long *c4 = A.ctr + (size_t)(wid % NGS_CTR_STRIPES) * 16;
		atomicAdd(c4 + 0, (unsigned long long)v0);
		if (v1) atomicAdd(c4 + 1, (unsigned long long)v1);
		if (v2) atomicAdd(c4 + 2, (unsigned long long)v2);
		if (v3) atomicAdd(c4 + 3, (unsigned long long)v3);
		if (v4 > A.st->max_ref_p1) atomicMax(&A.st->max_ref_p1, v4);
		if (v5) atomicAdd(&A.st->n_special, v5);
	}
}

/* ---- tile -> rows ------------------------------------------------------------------------------------------ */
/* For every tile of the launch: the rows whose records can reach it.  rkmax is non-decreasing by construction, so
 *   row_a = first row with rkmax >= key(contig, P0 - H)   (every earlier row only holds smaller keys)
 *   row_b = 1 + first row with rkmax >= key(contig, P0 + CT)  (sorted input: every later row starts at or behind it) */
__global__ void __launch_bounds__(256) k_tile_rows(const unsigned long long *rkmax, uint32_t n_rows, const uint32_t *tile2seg, const ngs_seg *seg,
                                                   uint64_t tile_off, uint64_t n_tiles, uint32_t *row_a, uint32_t *row_b)
{
	const uint64_t tile = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (tile >= n_tiles) return;
	const ngs_seg &sg = seg[tile2seg[(tile + tile_off) / NGS_CPT]];       /* tile2seg: one entry per difference-array tile, not pre-offset */
	const uint64_t p0 = (uint64_t)sg.clip_lo + (tile + tile_off - sg.tile0 * NGS_CPT) * NGS_CT;
	const uint64_t lo_pos = p0 > NGS_NEAR_H ? p0 - NGS_NEAR_H : 0;
	const unsigned long long ka = ((unsigned long long)sg.contig << 32) | (lo_pos + 1), kb = ((unsigned long long)sg.contig << 32) | (p0 + NGS_CT + 1);
	uint32_t lo = 0, hi = n_rows;                       /* first row with rkmax >= ka */
	while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (rkmax[mid] < ka) lo = mid + 1; else hi = mid; }
	const uint32_t a = lo;
	hi = n_rows;
	while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (rkmax[mid] < kb) lo = mid + 1; else hi = mid; }
	row_a[tile] = a;
	row_b[tile] = lo < n_rows ? lo + 1 : n_rows;
}

/* ---- heavy tiles (targeted panels: thousands of reads starting at one base) --------------------------------------
 * An owner CTA gathers its rows four at a time; a tile fed by thousands of rows (config 5: a million reads in one tile)
 * would keep one CTA busy for milliseconds while the rest of the GPU idles.  Tiles with more than NGS_HEAVY_ROWS rows
 * are therefore taken apart: k_heavy_plan cuts their row range into work items of NGS_HEAVY_SLICE rows and empties the
 * tile's own range; k_heavy_gather accumulates each item in shared memory like the owner would and adds the non-zero
 * sums to the tile of the far array, whose flag makes the owner merge it. */
#define NGS_HEAVY_ROWS  96
#define NGS_HEAVY_SLICE 32
__global__ void __launch_bounds__(256) k_heavy_plan(uint32_t *row_a, uint32_t *row_b, uint64_t n_tiles, uint4 *items, unsigned int *n_items, unsigned int cap)
{
	const uint64_t tile = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (tile >= n_tiles) return;
	const uint32_t a = row_a[tile], b = row_b[tile];
	if (b - a <= NGS_HEAVY_ROWS) return;
	const uint32_t n = (b - a + NGS_HEAVY_SLICE - 1) / NGS_HEAVY_SLICE;
	const unsigned int at = atomicAdd(n_items, n);
	if (at + n > cap) { atomicSub(n_items, n); return; }          /* no room (cannot happen: the list is sized from the row count): the owner keeps the tile */
	for (uint32_t i = 0; i < n; ++i) items[at + i] = make_uint4((uint32_t)tile, a + i * NGS_HEAVY_SLICE, min(b, a + (i + 1) * NGS_HEAVY_SLICE), 0u);
	row_b[tile] = a;                                              /* the owner gathers nothing itself */
}
__global__ void __launch_bounds__(256) k_heavy_gather(const uint4 *items, const unsigned int *n_items, const uint2 *ent, const ngs_row *rows, const uint32_t *tile2seg,
                                                     const ngs_seg *seg, uint64_t tile_off, int *diff, uint32_t *far_flag)
{
	__shared__ int s_acc[NGS_CT];
	const int t = threadIdx.x;
	const unsigned int n = *n_items;
	for (unsigned int it = blockIdx.x; it < n; it += gridDim.x) {
		const uint4 w = items[it];
		const uint64_t tile = w.x, tile_g = tile + tile_off;
		const ngs_seg &sg = seg[tile2seg[tile_g / NGS_CPT]];
		const uint32_t pos0 = sg.clip_lo + (uint32_t)((tile_g - sg.tile0 * NGS_CPT) * NGS_CT), ctg = sg.contig;
		const int64_t vis_lo = (int64_t)pos0 - NGS_NEAR_H;
		for (int i = t; i < NGS_CT; i += 256) s_acc[i] = 0;
		__syncthreads();
		for (uint32_t r = w.y; r < w.z; ++r) {
			const uint4 m = __ldg(reinterpret_cast<const uint4 *>(rows + r) + 1);          /* count, special, ctg, nrec */
			if ((uint32_t)t >= m.x) continue;
			const uint2 e = __ldg(ent + (size_t)r * NGS_ROW_SLOTS + t);
			const uint32_t kind = e.y >> 30, pos = e.x & 0x1FFFFFFFu;
			if (kind == NGS_ENT_CHAIN || (int)m.z + (int)(e.x >> 29) != (int)ctg || (int64_t)pos < vis_lo) continue;
			const uint32_t pl = e.y & 0x3FFFFFFFu, a0 = pos - pos0;
			uint32_t b0;
			if (kind == NGS_ENT_SIMPLE) b0 = a0 + pl;
			else {
				b0 = a0 + (pl & 0xFFFu);
				const uint32_t a1 = b0 + ((pl >> 12) & 0x3Fu), b1 = a1 + (pl >> 18);
				if (a1 < NGS_CT) atomicAdd(&s_acc[a1], 1);
				if (b1 < NGS_CT) atomicAdd(&s_acc[b1], -1);
			}
			if (a0 < NGS_CT) atomicAdd(&s_acc[a0], 1);
			if (b0 < NGS_CT) atomicAdd(&s_acc[b0], -1);
		}
		__syncthreads();
		int *g = diff + tile * (uint64_t)NGS_CT;
		for (int i = t; i < NGS_CT; i += 256) { const int v = s_acc[i]; if (v) atomicAdd(g + i, v); }
		if (t == 0) far_flag[tile] = 1u;
		__syncthreads();
	}
}

/* ---- K2..K5 fused: the tile owner -------------------------------------------------------------------------- */
struct CoverArgs {
	uint64_t n_tiles, tile_off;          /* tiles of this launch / global index of the first one; per-tile pointers are pre-offset */
	const uint32_t *tile2seg; const ngs_seg *seg;
	const uint2 *ent; const ngs_row *rows;              /* the whole table */
	const uint32_t *row_a, *row_b;
	int *diff; uint32_t *far_flag;       /* far array */
	unsigned long long *stB, *stF;       /* look-back words: status:2 | lastb:30 | sum:32  and  status:2 | bytes:62 */
	const unsigned long long *prev;      /* launch that continues inside a contig: the inclusive (sum, last boundary) word of the
	                                        tile before its first one, published by the previous launch; else null */
	unsigned long long *bins; uint32_t window; int variant;
	uint4 *tile_first, *tile_last;       /* first / last non-zero run of each tile (halo candidates) */
	NameTable nt;
	char *out; unsigned long long out_cap;
	uint32_t stage_bytes;
	int debug;                           /* bit 0: skip the look-back waits (timing experiments only, wrong output) */
	ngs_dev_state *st;
};

/* padded accumulator index: NGS_CITEMS consecutive elements per thread + 4 ints of padding -> conflict-free 128-bit reads */
#define NGS_CITEMS (NGS_CT / NGS_COVER_THREADS)          /* elements per thread: 8 or 16 */
__device__ __forceinline__ int acc_idx(int i) { return i + (i / NGS_CITEMS) * 4; }
#define NGS_ACC_INTS (NGS_CT * 3 / 2)                    /* >= padded accumulators; = u16 [NGS_CT] + i32 [NGS_CT] of the boundary list */
#define NGS_ACC_USED (NGS_CT / NGS_CITEMS * (NGS_CITEMS + 4))
#ifndef NGS_CIPT
#define NGS_CIPT 5                                       /* consecutive boundaries per thread per text batch */
#endif
#define NGS_CBATCH (NGS_COVER_THREADS * NGS_CIPT)
#define NGS_CNW (NGS_COVER_THREADS / 32)                 /* warps per owner CTA */
#define NGS_ROW_PASSES (NGS_ROW_SLOTS / NGS_COVER_THREADS)

/* generic bedGraph line, byte by byte (tiles with a long contig name or positions below 10^4, staging overflow) */
__device__ __noinline__ uint32_t bed_line_bytes(char *dst, const NameTable nt, uint32_t ctg, uint32_t s, uint32_t e, int d)
{
	const uint32_t a = nt.name_off[ctg], nl = nt.name_off[ctg + 1] - a;
	char *p = dst;
	for (uint32_t i = 0; i < nl; ++i) p[i] = nt.names[a + i];
	p += nl;
	p += ngs_put_u32(p, s); *p++ = '\t';
	p += ngs_put_u32(p, e); *p++ = '\t';
	p += ngs_put_i32(p, d); *p++ = '\n';
	return (uint32_t)(p - dst);
}
/* one line through the register writer with full conversions (a start before the tile, depth of five digits or more) */
/* The register line writer of ngscov_hd.h (ngs_lw) on a 32-bit shared-memory address, with the word stores predicated
 * in PTX: written as C `if`s the three conditional stores of a piece become a ladder of divergent branches (BSSY / BRA /
 * BSYNC per piece, a fifth of the instructions of a line); a predicated st.shared is one instruction. */
__device__ __forceinline__ void sts_if(uint32_t addr, uint32_t v, bool p)
{
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.u32 [%0], %1;\n\t}" :: "r"(addr), "r"(v), "r"((uint32_t)p) : "memory");
}
struct ngs_lws {
	uint32_t a;          /* shared address of the word the next byte goes to */
	uint32_t carry, k;   /* the k (0..3) low bytes of that word already produced */
	__device__ __forceinline__ void begin(uint32_t base, uint32_t byte_off) { a = base + (byte_off & ~3u); k = byte_off & 3u; carry = 0; }
	__device__ __forceinline__ void put12(uint32_t q0, uint32_t q1, uint32_t q2, uint32_t cnt)   /* cnt in 1..12 */
	{
		const uint32_t sel = 0x7654u - 0x1111u * k;
		const uint32_t x0 = __byte_perm(0u, q0, sel) | carry, x1 = __byte_perm(q0, q1, sel), x2 = __byte_perm(q1, q2, sel), x3 = __byte_perm(q2, 0u, sel);
		const uint32_t tot = k + cnt, nw = tot >> 2;
		sts_if(a, x0, nw > 0u); sts_if(a + 4u, x1, nw > 1u); sts_if(a + 8u, x2, nw > 2u);
		carry = nw < 2u ? (nw == 0u ? x0 : x1) : (nw == 2u ? x2 : x3);
		a += nw << 2; k = tot & 3u;
	}
	__device__ __forceinline__ void put16(uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3, uint32_t cnt)   /* cnt in 1..16 */
	{
		const uint32_t sel = 0x7654u - 0x1111u * k;
		const uint32_t x0 = __byte_perm(0u, q0, sel) | carry, x1 = __byte_perm(q0, q1, sel), x2 = __byte_perm(q1, q2, sel), x3 = __byte_perm(q2, q3, sel), x4 = __byte_perm(q3, 0u, sel);
		const uint32_t tot = k + cnt, nw = tot >> 2;
		sts_if(a, x0, nw > 0u); sts_if(a + 4u, x1, nw > 1u); sts_if(a + 8u, x2, nw > 2u); sts_if(a + 12u, x3, nw > 3u);
		carry = nw < 2u ? (nw == 0u ? x0 : x1) : (nw == 2u ? x2 : (nw == 3u ? x3 : x4));
		a += nw << 2; k = tot & 3u;
	}
};

/* (the writer state travels by value: a reference would pin it in local memory for the whole text loop) */
__device__ __noinline__ uint4 lw_generic_line(uint32_t *w, uint32_t carry, uint32_t k, uint32_t n_lo, uint32_t n_hi, uint32_t nl, uint32_t s, uint32_t e, uint32_t d)
{
	ngs_lw<uint32_t *> L; L.w = w; L.carry = carry; L.k = k;
	ngs_num xs, xe, xd;
	ngs_num_conv(s, xs); ngs_num_conv(e, xe); ngs_num_conv(d, xd);
	L.bed_line(n_lo, n_hi, nl, xs, xe, xd);
	return make_uint4((uint32_t)(L.w - w), L.carry, L.k, 0u);
}
/* cold paths kept out of line: the hot loop of k_cover must stay small enough for the instruction cache */
struct AtomicBinAdd { unsigned long long *bins; __device__ void operator()(uint32_t k, int64_t v) const { atomicAdd(bins + k, (unsigned long long)v); } };
__device__ __noinline__ void bin_run_slow(int variant, uint32_t s, uint32_t e, int d, uint32_t W, uint32_t L, uint32_t n_win, unsigned long long *bins)
{
	AtomicBinAdd add = { bins };
	ngs_bin_run(variant, s, e, (int64_t)d, W, L, n_win, add);
}
__device__ __noinline__ uint32_t line_len_generic(uint32_t nl, uint32_t s, uint32_t e, int d)
{
	return nl + ngs_ndigits(s) + ngs_ndigits(e) + ngs_len_i32(d) + 3u;
}
/* thread 0 / 1: the cached prefixes "name\tH" and "\tH" for H0 and H0 + 1 (positions of a tile span at most two H) */
__device__ __noinline__ void fill_tile_hcache(ngs_hcache *dst, uint32_t H, uint32_t n_lo, uint32_t n_hi, uint32_t nl)
{
	ngs_hcache c;
	ngs_hcache_fill(c, H, n_lo, n_hi, nl);
	*dst = c;
}

/* block-wide look-back over one status word per tile (status:2 | payload:62, payload added up), thread i inspects
 * predecessor tile-1-i.  Returns the exclusive prefix; publishes the inclusive one. */
__device__ __forceinline__ unsigned long long lookback_bytes(unsigned long long *status, unsigned long long tile, unsigned long long total,
                                                             unsigned int *s_li, unsigned long long *s_lb, ngs_dev_state *st)
{
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	unsigned long long run = 0;
	long long base = (long long)tile - 1;
	for (;;) {
		const long long idx = base - t;
		unsigned long long a = NGS_ST_INC << 62;
		if (idx >= 0) {
			unsigned int polls = 0;
			do { a = ld_relaxed_u64(&status[idx]); } while ((a >> 62) == NGS_ST_EMPTY && ++polls < NGS_WATCHDOG_POLLS);
			if ((a >> 62) == NGS_ST_EMPTY) { atomicOr(&st->flags, NGS_FLAG_STALL); a = NGS_ST_INC << 62; }
		}
		const unsigned int im = __ballot_sync(0xffffffffu, (a >> 62) == NGS_ST_INC);
		if (lane == 0) s_li[warp] = im ? (unsigned)(warp * 32 + __ffs(im) - 1) : 0xffffffffu;
		__syncthreads();
		unsigned int first_inc = 0xffffffffu;
		#pragma unroll
		for (int w = 0; w < NGS_CNW; ++w) first_inc = s_li[w] < first_inc ? s_li[w] : first_inc;
		unsigned long long v = (unsigned)t <= first_inc ? (a & 0x3FFFFFFFFFFFFFFFull) : 0;
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
		if (lane == 0) s_lb[warp] = v;
		__syncthreads();
		#pragma unroll
		for (int w = 0; w < NGS_CNW; ++w) run += s_lb[w];
		if (first_inc != 0xffffffffu) break;
		base -= NGS_COVER_THREADS;
		__syncthreads();
	}
	if (t == 0) st_relaxed_u64(&status[tile], (NGS_ST_INC << 62) | (run + total));
	return run;
}

/* rare: the tile's share of the far array */
__device__ __noinline__ void merge_far_tile(int *g, int *s_tile, uint32_t *flag)
{
	const int t = threadIdx.x;
	for (int j = 0; j < NGS_CITEMS; ++j) {
		const int i = j * NGS_COVER_THREADS + t;
		const int v = g[i];
		if (v) { s_tile[acc_idx(i)] += v; g[i] = 0; }
	}
	if (t == 0) *flag = 0;
}

template <bool TEXT>
__global__ void __launch_bounds__(NGS_COVER_THREADS, NGS_COVER_MINB) k_cover(const __grid_constant__ CoverArgs A)
{
	extern __shared__ __align__(16) char s_stage[];
	__shared__ __align__(16) int s_tile[NGS_ACC_INTS];        /* phase 1: padded accumulators; phase 2: compact boundary list
	                                                             (u16 position [NGS_CT] + i32 local depth [NGS_CT]) */
	__shared__ ngs_hcache s_hc[2];
	__shared__ unsigned int s_rcnt[32]; __shared__ int s_rctg[32];
	__shared__ long long s_wsum[NGS_CNW]; __shared__ unsigned int s_wcnt[NGS_CNW], s_wlast[NGS_CNW];
	__shared__ unsigned int s_lb_inc[NGS_CNW], s_lb_last[NGS_CNW]; __shared__ long long s_lb_sum[NGS_CNW];
	__shared__ long long s_bk[NGS_CNW], s_bv[NGS_CNW];
	__shared__ unsigned int s_wr[NGS_CNW]; __shared__ int s_wmax[NGS_CNW], s_wmin[NGS_CNW];
	__shared__ unsigned int s_rfirst[NGS_CNW], s_rlast[NGS_CNW];
	__shared__ unsigned int s_w[NGS_CNW], s_li[NGS_CNW];
	__shared__ unsigned long long s_lb[NGS_CNW];

	const unsigned full = 0xffffffffu;
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	const uint64_t tile = blockIdx.x;                    /* launch-local cover tile; dispatch order = tile order (see k_scan_runs) */
	if (tile >= A.n_tiles) return;
	const uint64_t tile_g = tile + A.tile_off;
	const uint32_t si = A.tile2seg[tile_g / NGS_CPT];    /* tile2seg is NOT pre-offset: one entry per difference-array tile */
	const uint32_t ra = A.row_a[tile], rb = A.row_b[tile], far = A.far_flag[tile];
	#pragma unroll
	for (int j = 0; j < NGS_ACC_USED / 4 / NGS_COVER_THREADS; ++j) reinterpret_cast<int4 *>(s_tile)[j * NGS_COVER_THREADS + t] = make_int4(0, 0, 0, 0);
	const ngs_seg sg = A.seg[si];
	const uint32_t pos0 = sg.clip_lo + (uint32_t)((tile_g - sg.tile0 * NGS_CPT) * NGS_CT);
	const uint32_t ctg = sg.contig;
	const uint32_t nl = A.nt.name_off[ctg + 1] - A.nt.name_off[ctg];
	const unsigned long long name8 = A.nt.name8[ctg];
	const uint32_t n_lo = (uint32_t)name8, n_hi = (uint32_t)(name8 >> 32);
	const uint32_t H0 = pos0 / 10000u;
	if (TEXT && t < 2 && H0 != 0 && nl <= 8u) fill_tile_hcache(&s_hc[t], H0 + (uint32_t)t, n_lo, n_hi, nl);

	/* ---- gather: +1/-1 of every visible record into the shared accumulators.  The row headers of up to 32 rows go
	 * through shared memory first, so the entry loads of all rows are independent of each other */
	{
		const int64_t vis_lo = (int64_t)pos0 - NGS_NEAR_H;
		for (uint32_t rbase = ra; rbase < rb; rbase += 32) {
			const uint32_t nr = rb - rbase < 32u ? rb - rbase : 32u;
			__syncthreads();                                 /* accumulators zeroed / previous row headers consumed */
			if ((uint32_t)t < nr) { const uint4 m = __ldg(reinterpret_cast<const uint4 *>(A.rows + rbase + t) + 1); s_rcnt[t] = m.x; s_rctg[t] = (int)m.z; }
			__syncthreads();
			/* a row has NGS_ROW_SLOTS slots: NGS_ROW_PASSES blocks of one slot per thread; four blocks in flight */
			for (uint32_t b0 = 0; b0 < nr * NGS_ROW_PASSES; b0 += 4) {
				uint2 e[4];
				#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const uint32_t bk = b0 + j, rw = bk / NGS_ROW_PASSES, slot = (bk % NGS_ROW_PASSES) * NGS_COVER_THREADS + (uint32_t)t;
					e[j] = make_uint2(0, NGS_ENT_CHAIN << 30);
					if (rw < nr && slot < s_rcnt[rw]) e[j] = __ldg(A.ent + (size_t)(rbase + rw) * NGS_ROW_SLOTS + slot);
				}
				#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const uint32_t kind = e[j].y >> 30, pos = e[j].x & 0x1FFFFFFFu;
					if (kind == NGS_ENT_CHAIN) continue;
					if (s_rctg[((b0 + j) / NGS_ROW_PASSES) & 31] + (int)(e[j].x >> 29) != (int)ctg || (int64_t)pos < vis_lo) continue;
					const uint32_t pl = e[j].y & 0x3FFFFFFFu;
					const uint32_t a0 = pos - pos0;                     /* wraps to a huge value for pos < pos0: fails the < NGS_CT test */
					uint32_t b0e;
					if (kind == NGS_ENT_SIMPLE) b0e = a0 + pl;
					else {
						b0e = a0 + (pl & 0xFFFu);
						const uint32_t a1 = b0e + ((pl >> 12) & 0x3Fu), b1 = a1 + (pl >> 18);
						if (a1 < NGS_CT) atomicAdd(&s_tile[acc_idx((int)a1)], 1);
						if (b1 < NGS_CT) atomicAdd(&s_tile[acc_idx((int)b1)], -1);
					}
					if (a0 < NGS_CT) atomicAdd(&s_tile[acc_idx((int)a0)], 1);
					if (b0e < NGS_CT) atomicAdd(&s_tile[acc_idx((int)b0e)], -1);
				}
			}
		}
	}
	__syncthreads();
	if (far) { merge_far_tile(A.diff + tile * (uint64_t)NGS_CT, s_tile, &A.far_flag[tile]); __syncthreads(); }

	/* ---- scan: NGS_CITEMS consecutive elements per thread */
	int d[NGS_CITEMS];
	#pragma unroll
	for (int j = 0; j < NGS_CITEMS / 4; ++j) {
		const int4 v = *reinterpret_cast<const int4 *>(&s_tile[t * (NGS_CITEMS + 4) + j * 4]);
		d[4 * j] = v.x; d[4 * j + 1] = v.y; d[4 * j + 2] = v.z; d[4 * j + 3] = v.w;
	}
	int tsum = 0; unsigned int tcnt = 0, tlast = 0;
	#pragma unroll
	for (int i = 0; i < NGS_CITEMS; ++i) { tsum += d[i]; if (d[i] != 0) { ++tcnt; tlast = (unsigned)(t * NGS_CITEMS + i + 1); } }
	int isum = tsum; unsigned int icnt = tcnt, ilast = tlast;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const int a = __shfl_up_sync(full, isum, o);
		const unsigned int b = __shfl_up_sync(full, icnt, o);
		const unsigned int c = __shfl_up_sync(full, ilast, o);
		if (lane >= o) { isum += a; icnt += b; ilast = ilast > c ? ilast : c; }
	}
	if (lane == 31) { s_wsum[warp] = isum; s_wcnt[warp] = icnt; s_wlast[warp] = ilast; }
	__syncthreads();                                     /* also: every thread has its elements, s_tile may be reused */
	long long wpre_sum = 0; unsigned int wpre_cnt = 0;
	long long tot_sum = 0; unsigned int tot_cnt = 0, tot_last = 0;
	#pragma unroll
	for (int w = 0; w < NGS_CNW; ++w) {
		if (w < warp) { wpre_sum += s_wsum[w]; wpre_cnt += s_wcnt[w]; }
		tot_sum += s_wsum[w]; tot_cnt += s_wcnt[w]; tot_last = s_wlast[w] > tot_last ? s_wlast[w] : tot_last;
	}
	const int psum = __shfl_up_sync(full, isum, 1);
	const unsigned int pcnt = __shfl_up_sync(full, icnt, 1);
	const uint32_t agg_lastb = tot_last ? pos0 + tot_last : 0;     /* 1 + contig position, 0 = none */
	const bool has_pred = tile > 0 || A.prev != nullptr;
	if (t == 0) st_relaxed_u64(&A.stB[tile], ((has_pred ? NGS_ST_AGG : NGS_ST_INC) << 62) | ((unsigned long long)agg_lastb << 32) | (uint32_t)(int)tot_sum);

	/* ---- compact the boundaries (tile-relative position, depth before it WITHOUT the carry) while the predecessors
	 * publish their aggregates */
	unsigned short *c_pos = reinterpret_cast<unsigned short *>(s_tile);
	int *c_dep = s_tile + NGS_CT / 2;
	if (tcnt) {
		int depth = (int)wpre_sum + (lane ? psum : 0);
		unsigned int r = wpre_cnt + (lane ? pcnt : 0);
		#pragma unroll
		for (int i = 0; i < NGS_CITEMS; ++i) {
			if (d[i] != 0) { c_pos[r] = (unsigned short)(t * NGS_CITEMS + i); c_dep[r] = depth; depth += d[i]; ++r; }
		}
	}

	/* ---- decoupled look-back over (sum, last boundary) */
	long long carry = 0; unsigned int tile_last = 0;
	if (has_pred && !(A.debug & 1)) {
		long long base = (long long)tile - 1;
		for (;;) {
			const long long idx = base - t;
			unsigned long long b = NGS_ST_INC << 62;             /* virtual tiles before the launch's first: inclusive zero ... */
			if (idx >= 0) {
				unsigned int polls = 0;
				do { b = ld_relaxed_u64(&A.stB[idx]); } while ((b >> 62) == NGS_ST_EMPTY && ++polls < NGS_WATCHDOG_POLLS);
				if ((b >> 62) == NGS_ST_EMPTY) { atomicOr(&A.st->flags, NGS_FLAG_STALL); b = NGS_ST_INC << 62; }
			} else if (idx == -1 && A.prev) b = ld_relaxed_u64(A.prev);   /* ... or the carry of the launch that finalized the tiles before */
			const unsigned int im = __ballot_sync(full, (b >> 62) == NGS_ST_INC);
			if (lane == 0) s_lb_inc[warp] = im ? (unsigned)(warp * 32 + __ffs(im) - 1) : 0xffffffffu;
			__syncthreads();
			unsigned int first_inc = 0xffffffffu;
			#pragma unroll
			for (int w = 0; w < NGS_CNW; ++w) first_inc = s_lb_inc[w] < first_inc ? s_lb_inc[w] : first_inc;
			int vs = 0; unsigned int vl = 0;
			if ((unsigned)t <= first_inc) { vs = (int)(uint32_t)b; vl = (unsigned int)((b >> 32) & 0x3FFFFFFFu); }
			vs = __reduce_add_sync(full, vs);
			const unsigned int has = __ballot_sync(full, vl != 0);
			unsigned int near_l = 0;
			if (has) near_l = __shfl_sync(full, vl, __ffs(has) - 1);      /* lowest lane = nearest tile */
			if (lane == 0) { s_lb_sum[warp] = vs; s_lb_last[warp] = near_l; }
			__syncthreads();
			unsigned int rl = 0;
			#pragma unroll
			for (int w = 0; w < NGS_CNW; ++w) { carry += s_lb_sum[w]; if (rl == 0) rl = s_lb_last[w]; }
			if (tile_last == 0) tile_last = rl;
			if (first_inc != 0xffffffffu) break;
			base -= NGS_COVER_THREADS;
			__syncthreads();
		}
		if (t == 0) {
			const unsigned int inc_last = agg_lastb ? agg_lastb : tile_last;
			st_relaxed_u64(&A.stB[tile], (NGS_ST_INC << 62) | ((unsigned long long)inc_last << 32) | (uint32_t)(int)(carry + tot_sum));
		}
	} else __syncthreads();                              /* c_pos / c_dep complete */
	if (tot_cnt == 0) {
		/* an empty tile still resolves and publishes its inclusive byte count: a successor must never have to look back
		 * through a long stretch of aggregate-only words (sparse coverage: most tiles of the genome are empty) */
		if (TEXT) {
			if (t == 0) st_relaxed_u64(&A.stF[tile], ((tile == 0 ? NGS_ST_INC : NGS_ST_AGG) << 62));
			const unsigned long long run = (tile > 0 && !(A.debug & 1)) ? lookback_bytes(A.stF, tile, 0, s_li, s_lb, A.st) : 0;
			if (t == 0 && tile == A.n_tiles - 1) A.st->fmt_bytes = run;
		}
		return;
	}
	const int icarry = (int)carry;
	const uint32_t first_start = tile_last ? tile_last - 1 : 0;       /* start of the run that ends at the tile's first boundary */
	const bool tile_slow = TEXT && (H0 == 0 || nl > 8u);               /* byte-wise lines for the whole tile */
	const bool multi = tot_cnt > NGS_CBATCH;

	/* ---- more boundaries than one batch holds (coverage above ~45x): the tile's text length is needed before the first
	 * batch is assembled, so the lengths get a pass of their own */
	unsigned int total = 0;
	if (TEXT && multi) {
		unsigned int my_len = 0;
		for (unsigned int r = t; r < tot_cnt; r += NGS_COVER_THREADS) {
			const int depth = c_dep[r] + icarry;
			if (depth != 0) my_len += line_len_generic(nl, r ? pos0 + c_pos[r - 1] : first_start, pos0 + c_pos[r], depth);
		}
		my_len = __reduce_add_sync(full, my_len);
		if (lane == 0) s_w[warp] = my_len;
		__syncthreads();
		#pragma unroll
		for (int w = 0; w < NGS_CNW; ++w) total += s_w[w];
		if (t == 0) st_relaxed_u64(&A.stF[tile], ((tile == 0 ? NGS_ST_INC : NGS_ST_AGG) << 62) | total);
		__syncthreads();
	}

	/* ---- batches of NGS_CBATCH boundaries, NGS_CIPT consecutive ones per thread: window bins, statistics and the
	 * bedGraph lines (assembled in registers, staged in shared memory as in k_format_bed; the byte-offset look-back runs
	 * after the first batch is assembled) */
	unsigned long long *const seg_bins = A.bins + sg.bins0;
	unsigned int my_runs = 0, my_first = 0xffffffffu, my_last = 0; int my_max = 0, my_min = 0;
	const uint32_t W = A.window;
	const uint32_t kt = pos0 / W;
	/* fast path of the bins: run inside the window of the tile's first position and inside the contig (positions and
	 * windows are below 2^30, so the bounds fit 32 bits) */
	const uint32_t wlo = kt * W, whi = min(wlo + W - 1u, sg.ref_len);
	const uint32_t wlo1 = wlo + W, whi1 = min(wlo1 + W - 1u, sg.ref_len);   /* a tile of 4096 positions reaches into window kt + 1 at most when W >= 4096 */
	long long facc = 0, facc1 = 0;                                   /* contributions to windows kt and kt + 1 */
	const uint32_t Hb = (H0 + 1u) * 10000u;                          /* positions >= Hb have H = H0 + 1 */
	const uint32_t nH0 = (TEXT && !tile_slow) ? s_hc[0].nH : 0u, nH1 = (TEXT && !tile_slow) ? s_hc[1].nH : 0u;
	uint32_t *sw = reinterpret_cast<uint32_t *>(s_stage);
	unsigned long long goff = 0; uint32_t done_bytes = 0;
	const unsigned int n_batches = (tot_cnt + NGS_CBATCH - 1) / NGS_CBATCH;
	for (unsigned int bt = 0; bt < n_batches; ++bt) {
		const unsigned int r0 = bt * NGS_CBATCH + NGS_CIPT * (unsigned)t;
		uint32_t S0 = 0, E[NGS_CIPT]; int D[NGS_CIPT];
		#pragma unroll
		for (int i = 0; i < NGS_CIPT; ++i) {
			const bool in = r0 + i < tot_cnt;
			E[i] = in ? pos0 + c_pos[r0 + i] : 0; D[i] = in ? c_dep[r0 + i] + icarry : 0;
		}
		if (r0 < tot_cnt) S0 = r0 ? pos0 + c_pos[r0 - 1] : first_start;
		uint32_t len[NGS_CIPT], tlen = 0;
		#pragma unroll
		for (int i = 0; i < NGS_CIPT; ++i) {
			len[i] = 0;
			if (D[i] != 0) {
				const uint32_t start = i ? E[i - 1] : S0, pos = E[i];
				if (start >= wlo && pos <= whi) facc += (long long)(pos - start) * D[i];
				else if (start >= wlo1 && pos <= whi1) facc1 += (long long)(pos - start) * D[i];
				else bin_run_slow(A.variant, start, pos, D[i], W, sg.ref_len, sg.n_win, seg_bins);   /* rare: a run leaving the tile's first window */
				++my_runs; my_max = D[i] > my_max ? D[i] : my_max; my_min = D[i] < my_min ? D[i] : my_min;
				my_first = r0 + i < my_first ? r0 + i : my_first; my_last = r0 + i;
				if (TEXT) {
					/* positions inside the tile have the digits of H0 or H0 + 1 in front of four more */
					if (tile_slow || start < pos0 || (uint32_t)D[i] >= 10000u) len[i] = line_len_generic(nl, start, pos, D[i]);
					else {
						const uint32_t nd = 1u + ((uint32_t)D[i] >= 10u) + ((uint32_t)D[i] >= 100u) + ((uint32_t)D[i] >= 1000u);
						len[i] = nl + (start >= Hb ? nH1 : nH0) + (pos >= Hb ? nH1 : nH0) + 11u + nd;
					}
					tlen += len[i];
				}
			}
		}
		if (!TEXT) continue;
		uint32_t inc = tlen;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(full, inc, o); if (lane >= o) inc += v; }
		if (lane == 31) s_w[warp] = inc;
		__syncthreads();
		uint32_t wpre = 0, btotal = 0;
		#pragma unroll
		for (int w = 0; w < NGS_CNW; ++w) { if (w < warp) wpre += s_w[w]; btotal += s_w[w]; }
		const uint32_t off = wpre + inc - tlen;
		if (!multi) { total = btotal; if (t == 0) st_relaxed_u64(&A.stF[tile], ((tile == 0 ? NGS_ST_INC : NGS_ST_AGG) << 62) | total); }
		const bool staged = btotal + 32u <= A.stage_bytes;
		if (t == 0 && staged && !tile_slow) sw[btotal >> 2] = 0;    /* the batch's last, partial word is only ever OR-ed into */
		const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sw);
		ngs_lws L;
		L.begin(sbase, off);
		if (staged && tlen) {
			if (tile_slow) {
				char *p = s_stage + off;
				#pragma unroll
				for (int i = 0; i < NGS_CIPT; ++i) if (len[i]) p += bed_line_bytes(p, A.nt, ctg, i ? E[i - 1] : S0, E[i], D[i]);
			} else {
				/* the prefixes of H0 / H0 + 1 live in shared memory; a thread keeps those of its current H in registers */
				uint32_t hcur = 2u, nH = 0, nh0 = 0, nh1 = 0, nh2 = 0, nh3 = 0, th0 = 0, th1 = 0, hbase = 0;
				uint32_t hlast = 2u, L4last = 0;                    /* four digits of the previous line's end, valid under H index hlast */
				#pragma unroll
				for (int i = 0; i < NGS_CIPT; ++i) {
					if (!len[i]) { hlast = 2u; continue; }
					const uint32_t start = i ? E[i - 1] : S0, pos = E[i];
					const uint32_t he = pos >= Hb, hs = start >= Hb;
					if (start < pos0 || (uint32_t)D[i] >= 10000u || hs != he) {
						const uint4 q = lw_generic_line(sw + ((L.a - sbase) >> 2), L.carry, L.k, n_lo, n_hi, nl, start, pos, (uint32_t)D[i]);
						L.a += q.x << 2; L.carry = q.y; L.k = q.z; hlast = 2u; continue;
					}
					if (he != hcur) {
						const ngs_hcache &c = s_hc[he];
						hcur = he; nH = c.nH; nh0 = c.nh0; nh1 = c.nh1; nh2 = c.nh2; nh3 = c.nh3; th0 = c.th0; th1 = c.th1; hbase = c.H * 10000u;
					}
					const uint32_t L4s = hlast == he ? L4last : ngs_ascii4(start - hbase);
					const uint32_t L4e = ngs_ascii4(pos - hbase);
					uint32_t dq0, dq1, dc;
					ngs_small_piece((uint32_t)D[i], 0x0Au, dq0, dq1, dc);
					L.put16(nh0, nh1, nh2, nh3, nl + nH);
					L.put12(L4s, th0, th1, 5u + nH);
					L.put12(L4e, 0x09u | (dq0 << 8), (dq0 >> 24) | (dq1 << 8), 5u + dc);
					L4last = L4e; hlast = he;
				}
			}
		}
		__syncthreads();                                        /* every head word stored */
		if (staged && !tile_slow && tlen && L.k) atomicOr(sw + ((L.a - sbase) >> 2), L.carry);
		if (bt == 0) {
			if (tile > 0 && !(A.debug & 1)) goff = lookback_bytes(A.stF, tile, total, s_li, s_lb, A.st);
			else __syncthreads();
			if (A.debug & 1) goff = tile * 12000ull;
			if (t == 0 && tile == A.n_tiles - 1) A.st->fmt_bytes = goff + total;
		} else __syncthreads();
		if (goff + total <= A.out_cap) {                        /* otherwise the host sees fmt_bytes > capacity */
			char *g = A.out + goff + done_bytes;
			if (!staged) {                                      /* lines too long for the staging area: straight to global */
				char *p = g + off;
				#pragma unroll
				for (int i = 0; i < NGS_CIPT; ++i) if (len[i]) p += bed_line_bytes(p, A.nt, ctg, i ? E[i - 1] : S0, E[i], D[i]);
			} else if (btotal) {
				/* copy-out: head bytes up to the first 16-byte boundary of the output, aligned 128-bit stores whose source
				 * starts (16 - mis) bytes into the staging area, tail bytes */
				const uint32_t mis = (uint32_t)((uintptr_t)g & 15);
				uint32_t hb = (16u - mis) & 15u;
				if (hb > btotal) hb = btotal;
				if ((uint32_t)t < hb) g[t] = s_stage[t];
				const uint32_t nv = (btotal - hb) >> 4;
				int4 *dst = reinterpret_cast<int4 *>(g + hb);
				const int4 *src = reinterpret_cast<const int4 *>(s_stage);
				if (hb == 0) {
					for (uint32_t c = t; c < nv; c += NGS_COVER_THREADS) dst[c] = src[c];
				} else {
					const uint32_t wsh = hb >> 2, sel = 0x3210u + 0x1111u * (hb & 3u);
					for (uint32_t c = t; c < nv; c += NGS_COVER_THREADS) {
						const int4 a = src[c], b = src[c + 1];
						uint32_t v0, v1, v2, v3, v4;
						if (wsh == 0) { v0 = a.x; v1 = a.y; v2 = a.z; v3 = a.w; v4 = b.x; }
						else if (wsh == 1) { v0 = a.y; v1 = a.z; v2 = a.w; v3 = b.x; v4 = b.y; }
						else if (wsh == 2) { v0 = a.z; v1 = a.w; v2 = b.x; v3 = b.y; v4 = b.z; }
						else { v0 = a.w; v1 = b.x; v2 = b.y; v3 = b.z; v4 = b.w; }
						int4 o;
						o.x = (int)__byte_perm(v0, v1, sel); o.y = (int)__byte_perm(v1, v2, sel);
						o.z = (int)__byte_perm(v2, v3, sel); o.w = (int)__byte_perm(v3, v4, sel);
						dst[c] = o;
					}
				}
				const uint32_t tb = hb + (nv << 4);
				if (tb + (uint32_t)t < btotal) g[tb + t] = s_stage[tb + t];
			}
		}
		done_bytes += btotal;
		if (bt + 1 < n_batches) __syncthreads();                /* staging area free for the next batch */
	}

	/* ---- warp- then block-level aggregation of the pending window contribution and the statistics
	 * (one global atomic per distinct window per tile instead of one per thread) */
	{
		long long wk = (long long)kt, wv = facc;
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) wv += __shfl_xor_sync(full, wv, o);
		if (__any_sync(full, facc1 != 0)) {
			long long w1 = facc1;
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) w1 += __shfl_xor_sync(full, w1, o);
			if (lane == 0 && w1) atomicAdd(seg_bins + kt + 1, (unsigned long long)w1);
		}
		const unsigned wr = __reduce_add_sync(full, my_runs);
		const int wmax = __reduce_max_sync(full, my_max), wmin = __reduce_min_sync(full, my_min);
		const unsigned wfirst = __reduce_min_sync(full, my_first), wlast = __reduce_max_sync(full, my_last);
		if (lane == 0) {
			s_bk[warp] = wk; s_bv[warp] = wv;
			s_wr[warp] = wr; s_wmax[warp] = wmax; s_wmin[warp] = wmin; s_rfirst[warp] = wfirst; s_rlast[warp] = wlast;
		}
	}
	__syncthreads();
	if (t == 0) {
		long long ck = -1, cv = 0;
		unsigned long long nr = 0; unsigned int fr = 0xffffffffu, lr = 0; int mx = 0, mn = 0;
		for (int w = 0; w < NGS_CNW; ++w) {
			if (s_bk[w] >= 0) {
				if (s_bk[w] == ck) cv += s_bv[w];
				else { if (ck >= 0 && cv) atomicAdd(seg_bins + ck, (unsigned long long)cv); ck = s_bk[w]; cv = s_bv[w]; }
			}
			nr += s_wr[w]; mx = s_wmax[w] > mx ? s_wmax[w] : mx; mn = s_wmin[w] < mn ? s_wmin[w] : mn;
			fr = s_rfirst[w] < fr ? s_rfirst[w] : fr; lr = s_rlast[w] > lr ? s_rlast[w] : lr;
		}
		if (ck >= 0 && cv) atomicAdd(seg_bins + ck, (unsigned long long)cv);
		atomicAdd(&A.st->n_boundaries, (unsigned long long)tot_cnt);
		if (nr) {
			atomicAdd(&A.st->n_runs, nr);
			atomicMax(&A.st->max_depth, mx); atomicMin(&A.st->min_depth, mn);
			A.tile_first[tile] = make_uint4(fr ? pos0 + c_pos[fr - 1] : first_start, pos0 + c_pos[fr], (uint32_t)(c_dep[fr] + icarry), ctg);
			A.tile_last[tile] = make_uint4(lr ? pos0 + c_pos[lr - 1] : first_start, pos0 + c_pos[lr], (uint32_t)(c_dep[lr] + icarry), ctg);
			atomicMin(&A.st->first_tile, (unsigned long long)tile_g); atomicMax(&A.st->last_tile_p1, (unsigned long long)tile_g + 1);
		}
	}
}

/* ---- bam_sliding_count: accumulation from the verified table ---------------------------------------------------- */
struct SlideArgs {
	unsigned int *cnt, *len;             /* [n_bins] reads starting in the window, sum of l_qseq (the reference's unsigned int: wraps) */
	unsigned long long *gc;              /* [n_bins] sum of the per-read GC counts, exact integer */
	unsigned int *ctg_reads;             /* [n_ref] mapped reads seen per contig (the reference allocates a contig's windows on its first read) */
	const unsigned long long *win_base;  /* [n_ref + 1] */
	const uint32_t *n_win;               /* [n_ref] target_len / window + 1 */
	int window;
	unsigned long long *n_count;
};
__device__ __forceinline__ void slide_add(const SlideArgs &S, uint32_t ctg, int32_t pos, uint32_t gc, uint32_t lq)
{
	atomicAdd(&S.ctg_reads[ctg], 1u);
	const uint32_t w = (uint32_t)(unsigned short)(pos / S.window);            /* (unsigned short)(c->pos / window), :116 */
	if (w >= S.n_win[ctg]) return;                                             /* the reference writes out of bounds (POS beyond the contig) */
	const unsigned long long k = S.win_base[ctg] + w;
	atomicAdd(&S.cnt[k], 1u); atomicAdd(&S.len[k], lq); atomicAdd(&S.gc[k], (unsigned long long)gc);
}
__global__ void __launch_bounds__(256) k_slide_accumulate(const __grid_constant__ IngestArgs A, const __grid_constant__ SlideArgs S)
{
	const unsigned full = 0xffffffffu;
	{
		const ngs_dev_state *st = A.st;
		if (st->n_bad != 0 || (st->flags & NGS_FLAG_EFORMAT)) return;      /* launched optimistically behind the verification */
	}
	const int lane = threadIdx.x & 31;
	const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5, wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	DevMem mem = { A.base, A.n };
	unsigned int s_rec = 0, s_kept = 0, s_cig = 0, s_max = 0;
	unsigned long long n_count = 0;
	for (int64_t base = 0; base < A.n_sub; base += n_warps * 32) {
		const int64_t c = base + (int64_t)lane * n_warps + wid;
		uint4 b = make_uint4(0, 0, 0xffffffffu, 0), d = make_uint4(0, 0, 0, 0);
		if (c < A.n_sub) { b = reinterpret_cast<const uint4 *>(A.rows + c)[1]; d = reinterpret_cast<const uint4 *>(A.rows + c)[2]; }
		s_rec += b.w; s_kept += d.x; s_cig += d.z; s_max = d.w > s_max ? d.w : s_max;
		unsigned todo = __ballot_sync(full, b.x != 0 || (b.y & NGS_ROW_OVERFLOW));
		while (todo) {
			const int src = __ffs(todo) - 1; todo &= todo - 1;
			const unsigned int count = __shfl_sync(full, b.x, src), special = __shfl_sync(full, b.y, src);
			const int ctg0 = (int)__shfl_sync(full, b.z, src);
			const int64_t row = base + (int64_t)src * n_warps + wid;
			if (special & NGS_ROW_OVERFLOW) {
				if (lane == 0) {                      /* a record that does not fit an entry: the row is walked again, record by record */
					long long hi = (row + 1) << A.shift; if (hi > A.n) hi = A.n;
					int64_t o = A.entry[row];
					while (o < hi) {
						ngs_hdr h; mem.hdr_lite(o, h);
						const CigReader<DevMem> cig = { mem, o + 36 + (int64_t)h.l_qname };
						if (h.ref_id >= 0 && !(h.flag & 4u) && ngs_region_pass(A.cfg, h.ref_id, h.pos, h.n_cigar, cig)) {
							const uint32_t gc = ngs_seq_gc(mem, o + 36 + (int64_t)h.l_qname + 4ll * h.n_cigar, h.l_seq) & 0xFFFFu;
							slide_add(S, (uint32_t)h.ref_id, h.pos, gc, (uint32_t)h.l_seq);
							++n_count;
						}
						o += 4 + (int64_t)h.block_size;
					}
				}
				continue;
			}
			for (unsigned int k0 = 0; k0 < count; k0 += 32) {
				const unsigned int k = k0 + lane;
				const bool on = k < count;
				uint2 e = make_uint2(0, 0);
				if (on) e = A.ent[(size_t)row * NGS_ROW_SLOTS + k];
				const uint32_t ctg = (uint32_t)(ctg0 + (int)(e.x >> 29)), pos = e.x & 0x1FFFFFFFu;
				const uint32_t w = (uint32_t)(unsigned short)((int32_t)pos / S.window);
				/* sorted input: the lanes of a warp mostly share one window -- one atomic per distinct window and field */
				const unsigned long long key = on ? (((unsigned long long)ctg << 32) | w) : ~0ull;
				const unsigned grp = __match_any_sync(full, key);
				const unsigned int gl = __reduce_add_sync(grp, on ? (e.y & 0xFFFFu) : 0u), gg = __reduce_add_sync(grp, on ? (e.y >> 16) : 0u);
				if (on && lane == __ffs(grp) - 1) {
					const unsigned int nrd = (unsigned int)__popc(grp);
					atomicAdd(&S.ctg_reads[ctg], nrd);
					if (w < S.n_win[ctg]) {
						const unsigned long long kk = S.win_base[ctg] + w;
						atomicAdd(&S.cnt[kk], nrd); atomicAdd(&S.len[kk], gl); atomicAdd(&S.gc[kk], (unsigned long long)gg);
					}
					n_count += nrd;
				}
			}
		}
	}
	const unsigned v0 = __reduce_add_sync(full, s_rec), v1 = __reduce_add_sync(full, s_kept), v3 = __reduce_add_sync(full, s_cig), v4 = __reduce_max_sync(full, s_max);
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) n_count += __shfl_xor_sync(full, n_count, o);
	if (lane == 0 && v0) {
		unsigned long long *c4 = A.ctr + (size_t)(wid % NGS_CTR_STRIPES) * 16;
		atomicAdd(c4 + 0, (unsigned long long)v0);
		if (v1) atomicAdd(c4 + 1, (unsigned long long)v1);
		if (v3) atomicAdd(c4 + 3, (unsigned long long)v3);
		if (v4 > A.st->max_ref_p1) atomicMax(&A.st->max_ref_p1, v4);
		if (n_count) atomicAdd(S.n_count, n_count);
	}
}

/* The reference accumulates the GC count of a window in a FLOAT, read after read in file order (:120): exact while the
 * sum stays below 2^24, order dependent beyond.  Windows whose exact integer sum reaches 2^24 are redone here the
 * reference's way: one thread per window walks the window's table entries in file order and adds in float. */
__global__ void k_slide_seqfloat(const uint32_t *list, uint32_t n_list, const uint2 *ent, const ngs_row *rows, const unsigned long long *rkmax, uint32_t n_rows,
                                 const unsigned long long *win_base, uint32_t n_ref, int window, float *out, unsigned int *err)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_list) return;
	const uint32_t bin = list[i];
	uint32_t lo = 0, hi = n_ref;                        /* contig of the bin */
	while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (win_base[mid] <= bin) lo = mid; else hi = mid; }
	const uint32_t ctg = lo, w = bin - (uint32_t)win_base[ctg];
	const unsigned long long p0 = (unsigned long long)w * (unsigned)window, p1 = p0 + (unsigned)window;
	const unsigned long long ka = ((unsigned long long)ctg << 32) | (p0 + 1), kb = ((unsigned long long)ctg << 32) | (p1 + 1);
	uint32_t a = 0, b = n_rows;
	while (a < b) { const uint32_t mid = (a + b) >> 1; if (rkmax[mid] < ka) a = mid + 1; else b = mid; }
	float f = 0.0f;
	for (uint32_t r = a; r < n_rows; ++r) {
		const ngs_row &rw = rows[r];
		if (rw.special & NGS_ROW_OVERFLOW) { atomicOr(err, 1u); break; }     /* records outside the table: cannot be replayed in order */
		for (uint32_t k = 0; k < rw.count; ++k) {
			const uint2 e = ent[(size_t)r * NGS_ROW_SLOTS + k];
			if ((uint32_t)(rw.ctg + (int)(e.x >> 29)) != ctg) continue;
			const uint32_t pos = e.x & 0x1FFFFFFFu;
			if (pos >= p0 && pos < p1) f += (float)(e.y >> 16);
		}
		if (rkmax[r] >= kb) break;
	}
	out[i] = f;
}

/* after a cover launch: the halo candidates of the first / last tile that produced a run -> the state block, and the
 * shard's running halo record (9 x int64: n_runs, first{contig,start,end,depth}, last{..}) that the device-side stitch
 * of region shards exchanges */
__global__ void k_halo_gather(ngs_dev_state *st, const uint4 *tile_first, const uint4 *tile_last, long long *halo)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	if (st->last_tile_p1 == 0) return;
	const uint4 a = tile_first[st->first_tile], b = tile_last[st->last_tile_p1 - 1];
	st->halo_first[0] = a.x; st->halo_first[1] = a.y; st->halo_first[2] = a.z; st->halo_first[3] = a.w;
	st->halo_last[0] = b.x; st->halo_last[1] = b.y; st->halo_last[2] = b.z; st->halo_last[3] = b.w;
	if (halo[0] == 0) { halo[1] = a.w; halo[2] = a.x; halo[3] = a.y; halo[4] = (int)a.z; }
	halo[5] = b.w; halo[6] = b.x; halo[7] = b.y; halo[8] = (int)b.z;
	halo[0] += (long long)st->n_runs;
}

/* ---- device-side stitch of region shards (shard.py stitch() / host/ngscov_shard.h stitch() restated) --
 * Every rank runs the same sequential rule over all ranks' halo records (gathered over NVLink by the caller) and applies
 * what concerns it: its first line is skipped when that run continues the previous shard's last run, its last line is
 * rewritten when a later shard extends it, and the window bins of a merged run are corrected by the rank whose first run
 * merged (the bam2wig window arithmetic is not additive over the pieces of a run); the bins are summed over ranks after. */
struct DevBinAdd { unsigned long long *bins; long long sign; __device__ void operator()(uint32_t k, int64_t v) const { bins[k] += (unsigned long long)(sign * v); } };

__global__ void k_stitch(const long long *halos, int world, int rank, char *text, unsigned long long bed_off, NameTable nt,
                         unsigned long long *bins, const unsigned long long *win_base, const uint32_t *ref_len, uint32_t window, int variant,
                         int do_text, unsigned long long *patch, unsigned long long *patch_host)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	int have = 0, owner = 0, drop_me = 0; uint32_t new_end_me = 0;
	long long cc = 0, cs = 0, ce = 0, cd = 0;
	for (int r = 0; r < world; ++r) {
		const long long *h = halos + 9 * r;
		if (h[0] == 0) { have = 0; continue; }
		int merged = 0;
		if (have && cc == h[1] && ce == h[2] && cd == h[4]) {
			merged = 1;
			if (r == rank) {
				const uint32_t L = ref_len[cc], nw = L / window + 1;
				DevBinAdd sub = { bins + win_base[cc], -1 }, add = { bins + win_base[cc], +1 };
				ngs_bin_run(variant, (uint32_t)cs, (uint32_t)ce, cd, window, L, nw, sub);
				ngs_bin_run(variant, (uint32_t)h[2], (uint32_t)h[3], h[4], window, L, nw, sub);
				ngs_bin_run(variant, (uint32_t)cs, (uint32_t)h[3], cd, window, L, nw, add);
				drop_me = 1;
			}
			ce = h[3];
			if (owner == rank) new_end_me = (uint32_t)ce;
		}
		if (merged && h[0] == 1) continue;               /* the whole shard was one run: it stays open into the next shard */
		cc = h[5]; cs = h[6]; ce = h[7]; cd = h[8]; owner = r; have = 1;
	}
	unsigned long long begin = 0, end = bed_off;
	const long long *me = halos + 9 * rank;
	if (do_text && me[0]) {
		char tmp[16];
		auto line_len = [&](uint32_t c, uint32_t s, uint32_t e, int d) -> uint32_t {
			return (nt.name_off[c + 1] - nt.name_off[c]) + ngs_put_u32(tmp, s) + ngs_put_u32(tmp, e) + ngs_put_i32(tmp, d) + 3u;
		};
		if (drop_me) begin = me[0] == 1 ? bed_off : line_len((uint32_t)me[1], (uint32_t)me[2], (uint32_t)me[3], (int)me[4]);
		if (new_end_me && !(drop_me && me[0] == 1)) {
			const uint32_t old = line_len((uint32_t)me[5], (uint32_t)me[6], (uint32_t)me[7], (int)me[8]);
			end = bed_off - old + bed_line_bytes(text + bed_off - old, nt, (uint32_t)me[5], (uint32_t)me[6], new_end_me, (int)me[8]);
		}
	}
	patch[0] = begin; patch[1] = end;
	patch_host[0] = begin; patch_host[1] = end;
	__threadfence_system();
}
