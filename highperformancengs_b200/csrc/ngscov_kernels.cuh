/* csrc/ngscov_kernels.cuh -- the sm_100a kernels of libngscov_cuda.
 *
 * Stage map (SURVEY.md section 7 / BASELINE.json north_star):
 *   K1+K2  k_ingest_walk     record-offset discovery (speculative per-sub-chunk block_size chain walk)
 *                            fused with core/CIGAR decode and +1/-1 scatter into the int32 difference
 *                            array (REDs, run-length combined per thread)
 *          k_verify_*        boundary check: each sub-chunk's entry must equal the running max of the
 *                            exits before it;  k_repair re-walks (undo + redo) the rare misses
 *          k_finalize_chunk  carry of a record cut by the chunk end
 *   K3+K4  k_scan_runs       single-pass decoupled-look-back prefix scan over the difference array,
 *                            run-boundary detection, stream compaction into (start,end,depth) records,
 *                            window bins; also re-zeroes the array for the next input
 *   K5     k_format_lines<>  device-side integer -> ASCII formatting with a second look-back scan for
 *                            byte offsets, shared-memory line staging and 128-bit coalesced stores
 *
 * None of this is a dense contraction, so there is no tensor-core path; every stage is bound by HBM
 * bandwidth or by L2 atomic throughput.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ngscov_hd.h"
#include "ngscov_inflate.h"

#ifndef NGS_SCAN_THREADS
#define NGS_SCAN_THREADS 256
#endif
#define NGS_SCAN_WARPS   (NGS_SCAN_THREADS / 32)
#ifndef NGS_LB_PER
#define NGS_LB_PER 1                 /* predecessors inspected per thread per look-back round (4: 30.1 ms, 1: 17.8 ms) */
#endif
#ifndef NGS_LB_THREADS
#define NGS_LB_THREADS NGS_SCAN_THREADS   /* threads taking part in the look-back (status polling is L2 traffic:
                                            256 pollers x 16 B per round is a quarter of the tile's own 16 KB) */
#endif
#define NGS_SCAN_SMEM    ((NGS_SCAN_THREADS * NGS_SCAN_ITEMS / 2 + NGS_SCAN_THREADS * NGS_SCAN_ITEMS) * 4)
#define NGS_SCAN_ITEMS   16
#define NGS_TILE         (NGS_SCAN_THREADS * NGS_SCAN_ITEMS)   /* 4096 difference-array elements */
#define NGS_FMT_THREADS  256
#define NGS_BAD_CAP      4096
#ifndef NGS_SCAN_MINB
#define NGS_SCAN_MINB 8
#endif
#ifndef NGS_ING_MINB
#define NGS_ING_MINB 1
#endif
#ifndef NGS_FMT_MINB
#define NGS_FMT_MINB 4
#endif

#define NGS_CTR_STRIPES  256          /* counter stripes, one 128-byte line each: same-address atomics from ~10^6
                                         warps otherwise serialise in L2 and stall every load of the state line */
#define NGS_FLAG_ERANGE  1u
#define NGS_FLAG_EFORMAT 2u

/* ------------------------------------------------------------------------------------------------ */
/* device-side global state of one handle (lives in device memory, mirrored to pinned host memory) */
struct ngs_dev_state {
	unsigned long long records, kept, m_ops, cigar_words;   /* ingest counters */
	unsigned long long spec_walks, spec_repaired;
	unsigned long long n_bad;          /* verify: mismatching sub-chunks found this pass */
	unsigned long long first_bad_key;  /* verify: min over (sub-chunk*2 + kind), kind 1 = broken chain */
	long long first_bad_expected;
	long long chunk_exit;              /* verify: max exit of the chunk (entry of the next chunk + n) */
	long long entry_next;              /* entry offset of the next chunk relative to its data start */
	unsigned int carry_len;
	unsigned int flags;                /* NGS_FLAG_* */
	unsigned int max_ref_p1;           /* 1 + highest refID decoded so far */
	unsigned int pad0;
	/* scan */
	unsigned long long scan_ticket;
	unsigned long long n_boundaries, n_runs;
	unsigned long long first_run, last_run;   /* record indices of the first / last non-zero run */
	int max_depth, min_depth;
	/* format */
	unsigned long long fmt_ticket;
	unsigned long long fmt_bytes;
	/* cover pipeline */
	unsigned long long key_seed;       /* running max of the record sort keys over the chunks already committed */
	unsigned long long key_chunk_max;  /* ... including the chunk being verified (committed by k_finalize_chunk) */
	unsigned int unsorted;             /* verify pass: a record key decreases (only meaningful in a clean pass) */
	unsigned int n_special;            /* CHAIN / possibly-far entries seen */
	unsigned long long first_tile, last_tile_p1;   /* cover: first / 1 + last tile holding a non-zero run */
	unsigned int halo_first[4], halo_last[4];      /* start, end, depth, contig of the first / last non-zero run */
};
#define NGS_FLAG_STALL   4u           /* a look-back wait hit its watchdog bound (predecessor tile never published) */

/* numeric value of an exit slot for the running max */
__device__ __forceinline__ long long exit_num(long long x) { return x == NGS_EXIT_INVALID ? 0 : x; }
#define NGS_VER_PER_BLOCK 1024   /* sub-chunks per verify block: 256 threads x 4 */

/* ------------------------------------------------------------------------------------------------ */
/* memory policy: unaligned little-endian field access on the inflated stream in HBM               */
/* 128-bit / 32-bit read-only loads that ask L2 to fetch the whole 128-byte line: the next record of a
 * chain usually starts in the same line, so its header then comes from L2 instead of DRAM */
__device__ __forceinline__ uint4 ldg_line_v4(const uint4 *p)
{
	uint4 r;
	asm volatile("ld.global.nc.L2::128B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
__device__ __forceinline__ uint32_t ldg_line_u32(const uint32_t *p)
{
	uint32_t r;
	asm volatile("ld.global.nc.L2::128B.u32 %0, [%1];" : "=r"(r) : "l"(p));
	return r;
}

struct DevMem {
	const uint8_t *base;     /* offset 0 of the chunk (carry bytes live just before it) */
	int64_t n;
	__device__ __forceinline__ int64_t size() const { return n; }
	__device__ __forceinline__ uint32_t u32(int64_t o) const
	{
		const uintptr_t a = (uintptr_t)(base + o);
		const uint32_t *p = (const uint32_t *)(a & ~(uintptr_t)3);
		const uint32_t sh = (uint32_t)(a & 3) * 8;
		const uint32_t lo = __ldg(p);
		if (sh == 0) return lo;
		return __funnelshift_r(lo, __ldg(p + 1), sh);
	}
	__device__ __forceinline__ uint32_t u8(int64_t o) const
	{
		const uintptr_t a = (uintptr_t)(base + o);
		return (__ldg((const uint32_t *)(a & ~(uintptr_t)3)) >> ((a & 3) * 8)) & 0xff;
	}
	/* first offset in [o, end) with 32 <= block_size <= NGS_MAX_BS and -1 <= refID < n_ref: a sliding
	 * three-word window, one aligned load per four candidate bytes */
	__device__ __forceinline__ int64_t next_candidate(int64_t o, int64_t end, uint32_t n_ref) const
	{
		if (o >= end) return NGS_NONE;
		const uintptr_t a = (uintptr_t)(base + o);
		const uint32_t *p = (const uint32_t *)(a & ~(uintptr_t)3);
		uint32_t sh = (uint32_t)(a & 3) * 8;
		uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
		p += 3;
		for (;;) {
			const uint32_t bs = __funnelshift_r(w0, w1, sh), rid = __funnelshift_r(w1, w2, sh);
			if (bs - 32u <= NGS_MAX_BS - 32u && rid + 1u <= n_ref) return o;
			if (++o >= end) return NGS_NONE;
			sh += 8;
			if (sh == 32) { sh = 0; w0 = w1; w1 = w2; w2 = __ldg(p++); }
		}
	}
	/* first 24 bytes of a record with two (sometimes three) 128-bit loads */
	__device__ __forceinline__ void hdr_lite(int64_t o, ngs_hdr &h) const
	{
		const uintptr_t a = (uintptr_t)(base + o);
		const uint4 *p = (const uint4 *)(a & ~(uintptr_t)15);
		const uint32_t bo = (uint32_t)(a & 15);
		const uint4 q0 = ldg_line_v4(p), q1 = ldg_line_v4(p + 1);
		uint4 q2 = make_uint4(0, 0, 0, 0);
		if (bo > 8) q2 = ldg_line_v4(p + 2);
		uint32_t w[10] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y };
		if (bo & 8) {
			#pragma unroll
			for (int i = 0; i < 8; ++i) w[i] = w[i + 2];
		}
		if (bo & 4) {
			#pragma unroll
			for (int i = 0; i < 7; ++i) w[i] = w[i + 1];
		}
		const uint32_t sh = (bo & 3) * 8;
		uint32_t f[6];
		#pragma unroll
		for (int i = 0; i < 6; ++i) f[i] = __funnelshift_r(w[i], w[i + 1], sh);
		h.block_size = (int32_t)f[0]; h.ref_id = (int32_t)f[1]; h.pos = (int32_t)f[2];
		h.l_qname = f[3] & 0xff; h.n_cigar = f[4] & 0xffff; h.flag = f[4] >> 16; h.l_seq = (int32_t)f[5];
		h.next_ref_id = 0; h.next_pos = 0;
	}
	__device__ __forceinline__ void hdr(int64_t o, ngs_hdr &h) const
	{
		h.block_size = (int32_t)u32(o); h.ref_id = (int32_t)u32(o + 4); h.pos = (int32_t)u32(o + 8);
		const uint32_t a = u32(o + 12), b = u32(o + 16);
		h.l_qname = a & 0xff; h.n_cigar = b & 0xffff; h.flag = b >> 16;
		h.l_seq = (int32_t)u32(o + 20); h.next_ref_id = (int32_t)u32(o + 24); h.next_pos = (int32_t)u32(o + 28);
	}
};

/* delta sink: one pending (address,count) pair for the +1 stream and one for the -1 stream, so that
 * pile-ups (identical starts / ends in consecutive records) collapse into a single RED */
struct DevSink {
	int *diff; unsigned int *flags; int dbg;
	unsigned long long pa, pb; int ca, cb;
	__device__ __forceinline__ void init(int *d, unsigned int *f, int dg = 0) { diff = d; flags = f; pa = pb = ~0ull; ca = cb = 0; dbg = dg; }
	__device__ __forceinline__ void add2(uint64_t ia, uint64_t ib, int sign)
	{
		if (dbg & 1) { ca += (int)ia; cb += (int)ib; return; }     /* timing experiment: no atomics */
		if (ia == pa) ca += sign; else { if (ca) atomicAdd(diff + pa, ca); pa = ia; ca = sign; }
		if (ib == pb) cb -= sign; else { if (cb) atomicAdd(diff + pb, cb); pb = ib; cb = -sign; }
	}
	__device__ __forceinline__ void flush()
	{
		if (dbg & 1) { if (ca + cb == 0x12345677) atomicAdd(diff, 1); return; }
		if (ca) atomicAdd(diff + pa, ca);
		if (cb) atomicAdd(diff + pb, cb);
		ca = cb = 0; pa = pb = ~0ull;
	}
	__device__ __forceinline__ void flag_range_error() { atomicOr(flags, NGS_FLAG_ERANGE); }
};

struct IngestArgs {
	const uint8_t *base; int64_t n;        /* chunk */
	int shift;                             /* log2(sub-chunk bytes) */
	int64_t n_sub;
	long long *entry, *exit;               /* [n_sub] */
	unsigned int *count;                   /* [n_sub] records starting in the sub-chunk */
	unsigned short *slots;                 /* [n_sub << slot_shift] record offsets relative to the sub-chunk's entry */
	int slot_shift;                        /* log2(slots per sub-chunk) = shift - 5 (a record is >= 36 bytes) */
	ngs_walk_cfg cfg;
	const uint32_t *ref_len;
	int *diff;
	ngs_dev_state *st;
	int debug;                             /* timing experiments only: bit0 no atomics */
	unsigned long long *ctr;               /* [NGS_CTR_STRIPES][16] striped counters: records, kept, m_ops, cigar_words */
	long long fixed_entry0;                /* used when entry0_from_state == 0 */
	int entry0_from_state;                 /* 1: first entry = st->entry_next (streamed pushes) */
	/* compact record table (cover pipeline, ngscov_cover.cuh): this chunk's rows, one per sub-chunk */
	ngs_row *rows; uint2 *ent; unsigned long long *rkmax;
	uint32_t *far_flag;                    /* [n_tiles] raised when a far delta was added to a tile of the difference array */
};

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

/* K1+K2: one thread per sub-chunk.  Both phases run warp-synchronously so that the 32 independent
 * chains of a warp stay converged: (1) the entry search alternates a cheap byte scan with a converged
 * deep validation round; (2) the walk advances every live chain by one record per iteration. */
__global__ void __launch_bounds__(256, NGS_ING_MINB) k_ingest_walk(IngestArgs A)
{
	const unsigned full = 0xffffffffu;
	const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	ngs_counts cnt = { 0, 0, 0, 0 };
	unsigned long long walks = 0;
	const bool live = c < A.n_sub;
	DevMem mem = { A.base, A.n };
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
	ngs_walk_state ws;
	ws.active = 0; ws.result = 0; ws.o = 0;
	DevSink sink; sink.init(A.diff, &A.st->flags, A.debug);
	if (entry != NGS_NONE) { ngs_walk_begin(mem, entry, hi, ws); walks = 1; }
	while (__any_sync(full, ws.active != 0)) {
		if (ws.active) ngs_walk_step(A.cfg, mem, hi, sink, +1, cnt, ws);
	}
	sink.flush();
	if (live) { A.entry[c] = entry; A.exit[c] = entry != NGS_NONE ? ws.result : 0; }
	{
		const unsigned mr = __reduce_max_sync(full, cnt.max_ref_p1);
		if ((threadIdx.x & 31) == 0 && mr > A.st->max_ref_p1) atomicMax(&A.st->max_ref_p1, mr);
	}
	/* block-aggregate the counters */
	unsigned long long v0 = warp_sum_u64(cnt.records), v1 = warp_sum_u64(cnt.kept),
	                   v2 = warp_sum_u64(cnt.m_ops), v3 = warp_sum_u64(cnt.cigar_words), v4 = warp_sum_u64(walks);
	if ((threadIdx.x & 31) == 0) {
		if (v0) atomicAdd(&A.st->records, v0);
		if (v1) atomicAdd(&A.st->kept, v1);
		if (v2) atomicAdd(&A.st->m_ops, v2);
		if (v3) atomicAdd(&A.st->cigar_words, v3);
		if (v4) atomicAdd(&A.st->spec_walks, v4);
	}
}

/* ------------------------------------------------------------------------------------------------ */
/* K1: record-offset discovery.  Same speculative sub-chunk chain walk as k_ingest_walk, but it only
 * frames records: every record start found is written to the sub-chunk's slot row (16-bit offset from
 * the sub-chunk's entry).  It needs few registers (full occupancy hides the dependent-load latency of
 * the chain) and issues no atomics; after k_verify_* / k_repair_chain have made the chains consistent
 * the decode runs as K2 with one thread per RECORD, where consecutive lanes hold consecutive records:
 * their loads and their +1/-1 REDs then fall into a handful of cache lines per warp instead of 32. */
template <class Mem>
__device__ __forceinline__ long long chain_walk(const ngs_walk_cfg &cfg, const Mem &mem, long long entry, int64_t hi,
                                                unsigned short *row, unsigned int n_slots, unsigned int &count)
{
	const int64_t n = mem.size();
	int64_t o = entry;
	count = 0;
	if (o >= hi) return o;
	if (o + 36 > n) return NGS_EXIT_PARTIAL(o);
	ngs_hdr h;
	mem.hdr_lite(o, h);
	for (;;) {
		if (!ngs_hdr_valid(h, cfg.n_ref)) return NGS_EXIT_INVALID;
		if (o + 36 + (int64_t)h.l_qname + 4ll * h.n_cigar > n) return NGS_EXIT_PARTIAL(o);
		if (count < n_slots) row[count] = (unsigned short)(o - entry);
		++count;
		const int64_t nx = o + 4 + (int64_t)h.block_size;
		if (nx >= hi) return nx;
		if (nx + 36 > n) return NGS_EXIT_PARTIAL(nx);
		mem.hdr_lite(nx, h);
		o = nx;
	}
}

__global__ void __launch_bounds__(256, 8) k_chain_walk(IngestArgs A)
{
	const unsigned full = 0xffffffffu;
	const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const bool live = c < A.n_sub;
	DevMem mem = { A.base, A.n };
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
	unsigned int count = 0; long long ex = 0;
	if (entry != NGS_NONE) ex = chain_walk(A.cfg, mem, entry, hi, A.slots + ((size_t)c << A.slot_shift), 1u << A.slot_shift, count);
	if (live) { A.entry[c] = entry; A.exit[c] = ex; A.count[c] = count; }
	const unsigned long long w = warp_sum_u64(entry != NGS_NONE ? 1ull : 0ull);
	if ((threadIdx.x & 31) == 0 && w) atomicAdd(&A.st->spec_walks, w);
}

/* Rare path, single thread: zip the true chain forward from the first mismatching sub-chunk until it
 * meets an already-framed entry, rewriting entry / exit / count / slots on the way. */
__global__ void k_repair_chain(IngestArgs A, long long first_bad, const long long *bexcl)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	DevMem mem = { A.base, A.n };
	long long cur = bexcl[first_bad / NGS_VER_PER_BLOCK];
	for (long long k = (first_bad / NGS_VER_PER_BLOCK) * NGS_VER_PER_BLOCK; k < first_bad; ++k) {
		const long long x = exit_num(A.exit[k]);
		cur = x > cur ? x : cur;
	}
	long long c = first_bad, repaired = 0;
	for (;;) {
		long long cc = cur >= A.n ? A.n_sub : (cur < 0 ? 0 : (cur >> A.shift));
		if (cc > A.n_sub) cc = A.n_sub;
		for (long long k = c; k < cc; ++k)              /* sub-chunks the true chain jumps over hold no record start */
			if (A.entry[k] != NGS_NONE) { A.entry[k] = NGS_NONE; A.exit[k] = 0; A.count[k] = 0; ++repaired; }
		if (cc >= A.n_sub) break;
		if (A.entry[cc] == cur) break;                  /* converged */
		long long hi = (cc + 1) << A.shift; if (hi > A.n) hi = A.n;
		unsigned int cnt = 0;
		const long long x = chain_walk(A.cfg, mem, cur, hi, A.slots + ((size_t)cc << A.slot_shift), 1u << A.slot_shift, cnt);
		A.entry[cc] = cur; A.exit[cc] = x; A.count[cc] = cnt; ++repaired;
		c = cc + 1;
		cur = (x == NGS_EXIT_INVALID || x >= NGS_EXIT_INF) ? A.n : x;
	}
	atomicAdd(&A.st->spec_repaired, (unsigned long long)repaired);
}

/* K2 sink: one record per lane.  Sorted input puts equal positions in adjacent lanes, so pile-ups
 * (targeted panels: thousands of reads starting at one base) are folded with a warp match before the
 * RED; `fold` is decided once per record batch from the headers, plain REDs otherwise. */
struct WarpSink {
	int *diff; unsigned int *flags; int dbg; bool fold;
	__device__ __forceinline__ void one(uint64_t idx, int v)
	{
		if (dbg & 1) return;                                   /* timing experiment: no atomics */
		if (fold) {
			const unsigned m = __match_any_sync(__activemask(), idx);
			if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(diff + idx, v * __popc(m));
		} else atomicAdd(diff + idx, v);
	}
	__device__ __forceinline__ void add2(uint64_t ia, uint64_t ib, int sign) { one(ia, sign); one(ib, -sign); }
	__device__ __forceinline__ void flag_range_error() { atomicOr(flags, NGS_FLAG_ERANGE); }
};

/* K2: core/CIGAR decode + scatter, one thread per framed record (slot).  Consecutive lanes hold
 * consecutive records, so a warp's header loads and its +1/-1 REDs fall into a handful of neighbouring
 * cache lines: measured 200 G RED/s for this pattern against 28-47 G RED/s when every lane walks its own
 * sub-chunk (tools/microbench/red_bench.cu).  The chain of dependent loads is kept to three levels
 * (slot -> header -> CIGAR + segment geometry, the latter L1-resident). */
__global__ void __launch_bounds__(256, 8) k_decode_scatter(IngestArgs A)
{
	const unsigned full = 0xffffffffu;
	const ngs_dev_state *st = A.st;
	if (st->n_bad != 0 || (st->flags & NGS_FLAG_EFORMAT)) return;      /* launched optimistically behind the verification */
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t c = (int64_t)(tid >> A.slot_shift);
	const unsigned int n_slots = 1u << A.slot_shift, k = (unsigned int)(tid & (n_slots - 1));
	ngs_counts cnt = { 0, 0, 0, 0 };
	unsigned int todo = 0;
	int64_t o = 0;
	if (c < A.n_sub) {
		const unsigned int have = A.count[c];
		if (k < have) {
			o = A.entry[c] + A.slots[((size_t)c << A.slot_shift) + k];
			todo = (k == n_slots - 1) ? have - k : 1;                    /* the last slot also owns any overflow records */
		}
	}
	DevMem mem = { A.base, A.n };
	ngs_hdr h;
	h.pos = -2 - (int)(threadIdx.x & 31); h.ref_id = -1; h.block_size = 0;
	if (todo) mem.hdr_lite(o, h);
	const ngs_walk_cfg &cfg = A.cfg;                                    /* segment table: a few L1-resident lines */
	/* pile-up detection: sorted input puts equal starts in adjacent lanes */
	const int pn = __shfl_down_sync(full, h.pos, 1);
	WarpSink sink = { A.diff, &A.st->flags, A.debug, false };
	/* fold only real pile-ups (half the warp shares its start with a neighbour); at WGS depth one or two equal
	 * neighbours per warp are normal and the warp match would cost more than the few duplicate REDs */
	sink.fold = __popc(__ballot_sync(full, todo && (threadIdx.x & 31) < 31 && pn == h.pos)) >= 16;
	if (todo) {
		ngs_decode_record(cfg, mem, o, h, sink, +1, cnt);
		for (--todo; todo; --todo) {                                      /* overflow records of a row (pathologically small records) */
			o += 4 + (int64_t)h.block_size;
			mem.hdr_lite(o, h);
			sink.fold = false;
			ngs_decode_record(cfg, mem, o, h, sink, +1, cnt);
		}
	}
	{
		const unsigned mr = __reduce_max_sync(0xffffffffu, cnt.max_ref_p1);
		if ((threadIdx.x & 31) == 0 && mr > A.st->max_ref_p1) atomicMax(&A.st->max_ref_p1, mr);
	}
	/* warp-level reduction, then one atomic per counter per warp into a stripe picked by the warp index (no
	 * block barrier: this kernel does so little per thread that the barrier was 19 % of its stall samples) */
	/* (32-bit REDUX: a thread decodes at most a sub-chunk's worth of records, the warp total stays far below 2^32;
	 * four 64-bit shuffle trees here were 23 % of the kernel's instructions) */
	const unsigned long long v0 = __reduce_add_sync(full, (unsigned)cnt.records), v1 = __reduce_add_sync(full, (unsigned)cnt.kept),
	                         v2 = __reduce_add_sync(full, (unsigned)cnt.m_ops), v3 = __reduce_add_sync(full, (unsigned)cnt.cigar_words);
	if ((threadIdx.x & 31) == 0 && v0) {
		unsigned long long *c4 = A.ctr + (size_t)((tid >> 5) % NGS_CTR_STRIPES) * 16;
		atomicAdd(c4 + 0, v0);
		if (v1) atomicAdd(c4 + 1, v1);
		if (v2) atomicAdd(c4 + 2, v2);
		if (v3) atomicAdd(c4 + 3, v3);
	}
}

/* ------------------------------------------------------------------------------------------------ */
/* K1+K2, shared-memory staged variant (small records).  The per-thread chain walk is a chain of
 * dependent loads; issued against HBM each hop costs a DRAM round trip and the kernel ends up bound
 * by memory latency at ~40 % of the bandwidth (ncu: long-scoreboard stalls on the header loads).
 * Here a CTA first pulls its whole stream tile (NGS_TILE_THREADS sub-chunks + a margin for the record
 * that overhangs the tile end) into shared memory with ONE bulk async copy (cp.async.bulk, TMA
 * engine, mbarrier completion), so HBM sees large sequential requests issued far ahead, and the hops
 * of the walk then cost shared-memory latency.  Bytes outside the staged window (the carry before
 * offset 0, or a header + CIGAR longer than the margin) fall back to the global-memory path. */
#define NGS_TILE_THREADS 64
#define NGS_TILE_MARGIN  1024

struct TileMem {
	DevMem g;                       /* whole chunk in global memory */
	const uint8_t *sm;              /* staged bytes: chunk offsets [t0, t1) */
	int64_t t0, t1;
	__device__ __forceinline__ int64_t size() const { return g.n; }
	__device__ __forceinline__ uint32_t u32(int64_t o) const
	{
		if (o >= t0 && o + 8 <= t1) {
			const uint32_t off = (uint32_t)(o - t0);
			const uint32_t *p = (const uint32_t *)(sm + (off & ~3u));
			return __funnelshift_r(p[0], p[1], (off & 3) * 8);
		}
		return g.u32(o);
	}
	__device__ __forceinline__ uint32_t u8(int64_t o) const
	{
		if (o >= t0 && o < t1) return sm[o - t0];
		return g.u8(o);
	}
	__device__ __forceinline__ void hdr_lite(int64_t o, ngs_hdr &h) const
	{
		if (o >= t0 && o + 48 <= t1) {
			const uint32_t off = (uint32_t)(o - t0);
			const uint32_t *p = (const uint32_t *)(sm + (off & ~3u));
			const uint32_t sh = (off & 3) * 8;
			uint32_t w[7];
			#pragma unroll
			for (int i = 0; i < 7; ++i) w[i] = p[i];
			uint32_t f[6];
			#pragma unroll
			for (int i = 0; i < 6; ++i) f[i] = __funnelshift_r(w[i], w[i + 1], sh);
			h.block_size = (int32_t)f[0]; h.ref_id = (int32_t)f[1]; h.pos = (int32_t)f[2];
			h.l_qname = f[3] & 0xff; h.n_cigar = f[4] & 0xffff; h.flag = f[4] >> 16; h.l_seq = (int32_t)f[5];
			h.next_ref_id = 0; h.next_pos = 0;
		} else g.hdr_lite(o, h);
	}
	__device__ __forceinline__ void hdr(int64_t o, ngs_hdr &h) const
	{
		h.block_size = (int32_t)u32(o); h.ref_id = (int32_t)u32(o + 4); h.pos = (int32_t)u32(o + 8);
		const uint32_t a = u32(o + 12), b = u32(o + 16);
		h.l_qname = a & 0xff; h.n_cigar = b & 0xffff; h.flag = b >> 16;
		h.l_seq = (int32_t)u32(o + 20); h.next_ref_id = (int32_t)u32(o + 24); h.next_pos = (int32_t)u32(o + 28);
	}
	__device__ __forceinline__ int64_t next_candidate(int64_t o, int64_t end, uint32_t n_ref) const
	{
		if (o >= end) return NGS_NONE;
		if (o < t0 || end + 12 > t1) return g.next_candidate(o, end, n_ref);
		uint32_t off = (uint32_t)(o - t0);
		const uint32_t *p = (const uint32_t *)(sm + (off & ~3u));
		uint32_t sh = (off & 3) * 8;
		uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
		p += 3;
		for (;;) {
			const uint32_t bs = __funnelshift_r(w0, w1, sh), rid = __funnelshift_r(w1, w2, sh);
			if (bs - 32u <= NGS_MAX_BS - 32u && rid + 1u <= n_ref) return o;
			if (++o >= end) return NGS_NONE;
			sh += 8;
			if (sh == 32) { sh = 0; w0 = w1; w1 = w2; w2 = *p++; }
		}
	}
};

__global__ void __launch_bounds__(NGS_TILE_THREADS) k_ingest_tiles(IngestArgs A)
{
	extern __shared__ __align__(128) uint8_t s_stream[];
	__shared__ __align__(8) unsigned long long s_bar;
	const unsigned full = 0xffffffffu;
	const int t = threadIdx.x;
	const int64_t tile_bytes = (int64_t)NGS_TILE_THREADS << A.shift;
	const int64_t t0 = (int64_t)blockIdx.x * tile_bytes;
	int64_t t1 = t0 + tile_bytes + NGS_TILE_MARGIN;
	if (t1 > A.n) t1 = A.n;
	t1 = t0 + ((t1 - t0) & ~(int64_t)15);                     /* bulk copies move multiples of 16 bytes */
	const uint32_t nbytes = (uint32_t)(t1 - t0);
	const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
	if (t == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (t == 0 && nbytes) {
		const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_stream);
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(nbytes) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		             :: "r"(dst), "l"(A.base + t0), "r"(nbytes), "r"(bar) : "memory");
	}
	if (nbytes) {
		uint32_t done = 0;
		while (!done) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			             : "=r"(done) : "r"(bar) : "memory");
		}
	}
	const int64_t c = (int64_t)blockIdx.x * NGS_TILE_THREADS + t;
	ngs_counts cnt = { 0, 0, 0, 0 };
	unsigned long long walks = 0;
	const bool live = c < A.n_sub;
	TileMem mem; mem.g.base = A.base; mem.g.n = A.n; mem.sm = s_stream; mem.t0 = t0; mem.t1 = t1;
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
	ngs_walk_state ws;
	ws.active = 0; ws.result = 0; ws.o = 0;
	DevSink sink; sink.init(A.diff, &A.st->flags, A.debug);
	if (entry != NGS_NONE) { ngs_walk_begin(mem, entry, hi, ws); walks = 1; }
	while (__any_sync(full, ws.active != 0)) {
		if (ws.active) ngs_walk_step(A.cfg, mem, hi, sink, +1, cnt, ws);
	}
	sink.flush();
	if (live) { A.entry[c] = entry; A.exit[c] = entry != NGS_NONE ? ws.result : 0; }
	{
		const unsigned mr = __reduce_max_sync(full, cnt.max_ref_p1);
		if ((threadIdx.x & 31) == 0 && mr > A.st->max_ref_p1) atomicMax(&A.st->max_ref_p1, mr);
	}
	unsigned long long v0 = warp_sum_u64(cnt.records), v1 = warp_sum_u64(cnt.kept),
	                   v2 = warp_sum_u64(cnt.m_ops), v3 = warp_sum_u64(cnt.cigar_words), v4 = warp_sum_u64(walks);
	if ((threadIdx.x & 31) == 0) {
		if (v0) atomicAdd(&A.st->records, v0);
		if (v1) atomicAdd(&A.st->kept, v1);
		if (v2) atomicAdd(&A.st->m_ops, v2);
		if (v3) atomicAdd(&A.st->cigar_words, v3);
		if (v4) atomicAdd(&A.st->spec_walks, v4);
	}
}


/* ------------------------------------------------------------------------------------------------ */
/* K1+K2 in one pass over the stream ("tiles2"): a CTA stages a tile of NGS_T2_SUB sub-chunks in shared
 * memory with one bulk async copy (TMA engine), frames its records there (speculative per-sub-chunk
 * chain walk, hops at shared-memory latency, offsets into a shared slot table), then decodes them one
 * thread per RECORD straight from shared memory, so the REDs are warp-coalesced (consecutive lanes =
 * consecutive records) and HBM sees the stream exactly once.  Deltas are applied before the chunk-level
 * verification, so a mis-speculated sub-chunk is repaired with the undo/redo walk of k_repair. */
#define NGS_T2_SUB     64           /* sub-chunks (walker threads) per tile */
#define NGS_T2_THREADS 256

__global__ void __launch_bounds__(NGS_T2_THREADS) k_ingest_tiles2(IngestArgs A)
{
	extern __shared__ __align__(128) uint8_t s_stream[];
	__shared__ __align__(8) unsigned long long s_bar;
	__shared__ long long s_entry[NGS_T2_SUB];
	__shared__ unsigned int s_count[NGS_T2_SUB];
	const unsigned full = 0xffffffffu;
	const int t = threadIdx.x, lane = t & 31;
	const int64_t tile_bytes = (int64_t)NGS_T2_SUB << A.shift;
	const int64_t t0 = (int64_t)blockIdx.x * tile_bytes;
	int64_t t1 = t0 + tile_bytes + NGS_TILE_MARGIN;
	if (t1 > A.n) t1 = A.n;
	t1 = t0 + ((t1 - t0) & ~(int64_t)15);                     /* bulk copies move multiples of 16 bytes */
	const uint32_t nbytes = (uint32_t)(t1 - t0);
	const uint32_t n_slots = 1u << A.slot_shift;
	unsigned short *s_slots = reinterpret_cast<unsigned short *>(s_stream + tile_bytes + NGS_TILE_MARGIN);   /* [NGS_T2_SUB][n_slots] */
	const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
	if (t == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (t == 0 && nbytes) {
		const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_stream);
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(nbytes) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		             :: "r"(dst), "l"(A.base + t0), "r"(nbytes), "r"(bar) : "memory");
	}
	if (nbytes) {
		uint32_t done = 0;
		while (!done) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			             : "=r"(done) : "r"(bar) : "memory");
		}
	}
	TileMem mem; mem.g.base = A.base; mem.g.n = A.n; mem.sm = s_stream; mem.t0 = t0; mem.t1 = t1;
	/* ---- phase 1 (first NGS_T2_SUB threads = two warps): frame the records of each sub-chunk */
	if (t < NGS_T2_SUB) {
		const int64_t c = (int64_t)blockIdx.x * NGS_T2_SUB + t;
		const bool live = c < A.n_sub;
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
		unsigned int count = 0; long long ex = 0;
		if (entry != NGS_NONE) ex = chain_walk(A.cfg, mem, entry, hi, s_slots + (size_t)t * n_slots, n_slots, count);
		s_entry[t] = entry; s_count[t] = count;
		if (live) { A.entry[c] = entry; A.exit[c] = ex; }
		const unsigned long long w = warp_sum_u64(entry != NGS_NONE ? 1ull : 0ull);
		if (lane == 0 && w) atomicAdd(&A.st->spec_walks, w);
	}
	__syncthreads();
	/* ---- phase 2 (all threads): decode + scatter, one thread per framed record, straight from shared memory */
	ngs_counts cnt = { 0, 0, 0, 0 };
	WarpSink sink = { A.diff, &A.st->flags, A.debug, false };
	const uint32_t total = NGS_T2_SUB * n_slots;
	const uint32_t total_up = (total + NGS_T2_THREADS - 1) / NGS_T2_THREADS * NGS_T2_THREADS;   /* warp-uniform trip count */
	for (uint32_t p = t; p < total_up; p += NGS_T2_THREADS) {
		const uint32_t r = p >> A.slot_shift, k = p & (n_slots - 1);
		unsigned int todo = 0; int64_t o = 0;
		if (p < total) {
			const unsigned int have = s_count[r];
			if (k < have) { o = s_entry[r] + s_slots[(size_t)r * n_slots + k]; todo = (k == n_slots - 1) ? have - k : 1; }
		}
		ngs_hdr h;
		h.pos = -2 - lane; h.ref_id = -1; h.block_size = 0;
		if (todo) mem.hdr_lite(o, h);
		const int pn = __shfl_down_sync(full, h.pos, 1);
		sink.fold = __popc(__ballot_sync(full, todo && lane < 31 && pn == h.pos)) >= 16;
		if (todo) {
			ngs_decode_record(A.cfg, mem, o, h, sink, +1, cnt);
			for (--todo; todo; --todo) {
				o += 4 + (int64_t)h.block_size;
				mem.hdr_lite(o, h);
				sink.fold = false;
				ngs_decode_record(A.cfg, mem, o, h, sink, +1, cnt);
			}
		}
		__syncwarp();
	}
	{
		const unsigned mr = __reduce_max_sync(full, cnt.max_ref_p1);
		if (lane == 0 && mr > A.st->max_ref_p1) atomicMax(&A.st->max_ref_p1, mr);
	}
	unsigned long long v0 = warp_sum_u64(cnt.records), v1 = warp_sum_u64(cnt.kept),
	                   v2 = warp_sum_u64(cnt.m_ops), v3 = warp_sum_u64(cnt.cigar_words);
	if (lane == 0 && v0) {
		unsigned long long *c4 = A.ctr + (size_t)((blockIdx.x * (NGS_T2_THREADS / 32) + (t >> 5)) % NGS_CTR_STRIPES) * 16;
		atomicAdd(c4 + 0, v0);
		if (v1) atomicAdd(c4 + 1, v1);
		if (v2) atomicAdd(c4 + 2, v2);
		if (v3) atomicAdd(c4 + 3, v3);
	}
}

__device__ __forceinline__ long long block_max_ll(long long v, long long *sm)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) { long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
	const int w = threadIdx.x >> 5;
	__syncthreads();
	if ((threadIdx.x & 31) == 0) sm[w] = v;
	__syncthreads();
	long long r = sm[0];
	for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = sm[i] > r ? sm[i] : r;
	return r;
}


__global__ void __launch_bounds__(256) k_verify_blockmax(const long long *exit, int64_t n_sub, long long *bmax)
{
	__shared__ long long sm[8];
	const int64_t b0 = (int64_t)blockIdx.x * NGS_VER_PER_BLOCK;
	long long m = 0;
	#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const int64_t c = b0 + i * 256 + threadIdx.x;
		if (c < n_sub) { long long x = exit_num(exit[c]); m = x > m ? x : m; }
	}
	m = block_max_ll(m, sm);
	if (threadIdx.x == 0) bmax[blockIdx.x] = m;
}

/* single block: exclusive running max over the block maxima, seeded with the chunk's first entry */
__global__ void __launch_bounds__(1024) k_verify_scan(long long *bmax, int64_t n_blocks, ngs_dev_state *st,
                                                      long long fixed_entry0, int entry0_from_state)
{
	__shared__ long long sm[32];
	__shared__ long long carry_s;
	if (threadIdx.x == 0) {
		carry_s = entry0_from_state ? st->entry_next : fixed_entry0;
		st->n_bad = 0; st->first_bad_key = ~0ull;
	}
	__syncthreads();
	for (int64_t b0 = 0; b0 < n_blocks; b0 += 1024) {
		const int64_t b = b0 + threadIdx.x;
		const long long own = b < n_blocks ? bmax[b] : 0;
		/* inclusive max-scan within the block of 1024 */
		long long v = own;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, v, o); if ((threadIdx.x & 31) >= o && t > v) v = t; }
		if ((threadIdx.x & 31) == 31) sm[threadIdx.x >> 5] = v;
		__syncthreads();
		long long pre = carry_s;
		for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) pre = sm[w] > pre ? sm[w] : pre;
		long long incl = v > pre ? v : pre;
		/* exclusive value for this block = max(pre, values of earlier lanes) */
		long long prev = __shfl_up_sync(0xffffffffu, v, 1);
		long long excl = (threadIdx.x & 31) ? (prev > pre ? prev : pre) : pre;
		if (b < n_blocks) bmax[b] = excl;
		__syncthreads();
		if (threadIdx.x == 1023) carry_s = incl;
		__syncthreads();
	}
	if (threadIdx.x == 0) st->chunk_exit = carry_s;
}

__global__ void __launch_bounds__(256) k_verify_check(const long long *entry, const long long *exit, int64_t n_sub,
                                                      const long long *bexcl, int shift, int64_t n, ngs_dev_state *st)
{
	__shared__ long long sm[8];
	const int64_t c0 = (int64_t)blockIdx.x * NGS_VER_PER_BLOCK + (int64_t)threadIdx.x * 4;
	long long ex[4], en[4];
	#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const int64_t c = c0 + i;
		ex[i] = c < n_sub ? exit_num(exit[c]) : 0;
		en[i] = c < n_sub ? entry[c] : NGS_NONE;
	}
	long long tmax = ex[0];
	tmax = ex[1] > tmax ? ex[1] : tmax; tmax = ex[2] > tmax ? ex[2] : tmax; tmax = ex[3] > tmax ? ex[3] : tmax;
	/* exclusive max-scan of per-thread maxima across the block */
	long long v = tmax;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, v, o); if ((threadIdx.x & 31) >= o && t > v) v = t; }
	if ((threadIdx.x & 31) == 31) sm[threadIdx.x >> 5] = v;
	__syncthreads();
	long long pre = bexcl[blockIdx.x];
	for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) pre = sm[w] > pre ? sm[w] : pre;
	long long prev = __shfl_up_sync(0xffffffffu, v, 1);
	long long run = (threadIdx.x & 31) ? (prev > pre ? prev : pre) : pre;
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
				const unsigned long long key = (unsigned long long)c * 2 + (unsigned)kind;
				const unsigned long long old = atomicMin(&st->first_bad_key, key);
				(void)old;
			}
			run = ex[i] > run ? ex[i] : run;
		}
	}
}

/* Rare path: single thread zips the true chain forward from the first mismatching sub-chunk until it
 * re-converges with an already-walked entry, undoing (sign -1) every mis-speculated walk it crosses. */
__global__ void k_repair(IngestArgs A, long long first_bad, const long long *bexcl)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	/* expected entry of the first bad sub-chunk = running max of the exits before it */
	long long expected0 = bexcl[first_bad / NGS_VER_PER_BLOCK];
	for (long long k = (first_bad / NGS_VER_PER_BLOCK) * NGS_VER_PER_BLOCK; k < first_bad; ++k) {
		const long long x = exit_num(A.exit[k]);
		expected0 = x > expected0 ? x : expected0;
	}
	DevMem mem = { A.base, A.n };
	DevSink sink; sink.init(A.diff, &A.st->flags);
	ngs_counts add = { 0, 0, 0, 0 }, sub = { 0, 0, 0, 0 };
	long long c = first_bad, cur = expected0, repaired = 0;
	for (;;) {
		long long cc = cur >= A.n ? A.n_sub : (cur < 0 ? 0 : (cur >> A.shift));
		if (cc > A.n_sub) cc = A.n_sub;
		for (long long k = c; k < cc; ++k) {            /* sub-chunks the true chain jumps over */
			if (A.entry[k] != NGS_NONE) {
				long long hi = (k + 1) << A.shift; if (hi > A.n) hi = A.n;
				ngs_walk(A.cfg, mem, A.entry[k], hi, sink, -1, sub);
				A.entry[k] = NGS_NONE; A.exit[k] = 0; ++repaired;
			}
		}
		if (cc >= A.n_sub) break;
		if (A.entry[cc] == cur) break;                  /* converged */
		long long hi = (cc + 1) << A.shift; if (hi > A.n) hi = A.n;
		if (A.entry[cc] != NGS_NONE) ngs_walk(A.cfg, mem, A.entry[cc], hi, sink, -1, sub);
		A.entry[cc] = cur;
		const long long x = ngs_walk(A.cfg, mem, cur, hi, sink, +1, add);
		A.exit[cc] = x; ++repaired;
		c = cc + 1;
		if (x == NGS_EXIT_INVALID || x >= NGS_EXIT_INF) cur = A.n;   /* chain ends in this chunk */
		else cur = x;
	}
	sink.flush();
	atomicAdd(&A.st->records, add.records - sub.records);
	atomicAdd(&A.st->kept, add.kept - sub.kept);
	atomicAdd(&A.st->m_ops, add.m_ops - sub.m_ops);
	atomicAdd(&A.st->cigar_words, add.cigar_words - sub.cigar_words);
	atomicAdd(&A.st->spec_repaired, (unsigned long long)repaired);
	if (add.max_ref_p1) atomicMax(&A.st->max_ref_p1, add.max_ref_p1);
}

/* after a clean verify: derive the next chunk's entry and save the bytes of a cut record */
__global__ void __launch_bounds__(256) k_finalize_chunk(const uint8_t *base, int64_t n, ngs_dev_state *st, uint8_t *carry)
{
	const long long ce = st->chunk_exit;
	__shared__ long long s_o;
	if (threadIdx.x == 0) {
		if (ce >= NGS_EXIT_INF) {
			const long long o = ce - NGS_EXIT_INF - (1ll << 40);
			if (n - o > (long long)NGS_CARRY_MAX || n - o < 0) { atomicOr(&st->flags, NGS_FLAG_EFORMAT); s_o = n; }
			else s_o = o;
		} else s_o = n;
	}
	__syncthreads();
	const long long o = s_o;
	const unsigned int len = (unsigned int)(n - o);
	for (unsigned int i = threadIdx.x; i < len; i += blockDim.x) carry[i] = base[o + i];
	if (threadIdx.x == 0) {
		st->carry_len = len;
		st->entry_next = len ? -(long long)len : (ce >= NGS_EXIT_INF ? 0 : ce - n);
		if (st->key_chunk_max > st->key_seed) st->key_seed = st->key_chunk_max;
	}
}

/* place the saved carry bytes right before the data start of the next chunk buffer */
__global__ void __launch_bounds__(256) k_place_carry(uint8_t *data_start, const ngs_dev_state *st, const uint8_t *carry)
{
	const unsigned int len = st->carry_len;
	for (unsigned int i = threadIdx.x; i < len; i += blockDim.x) data_start[(long long)i - (long long)len] = carry[i];
}

/* ------------------------------------------------------------------------------------------------ */
/* K3+K4: decoupled look-back scan + run compaction + window bins                                    */

#define NGS_ST_EMPTY 0ull
#define NGS_ST_AGG   1ull
#define NGS_ST_INC   2ull

struct ScanArgs {
	int *diff;                       /* difference array, zeroed again on the way through */
	uint64_t n_tiles;                /* tiles of this launch (the whole array, or the segments of finished contigs) */
	uint64_t tile_off;               /* global index of the launch's first tile; diff, tile2seg, stA, stB are passed
	                                    pre-offset so the look-back spin loop only sees launch-local indices (a runtime
	                                    first-tile bound inside that loop cost 47 % of the kernel) */
	const uint32_t *tile2seg;        /* [n_tiles] */
	const ngs_seg *seg;
	unsigned long long *stA, *stB;   /* [n_tiles] look-back words: A = status:2|count:62, B = status:2|lastb:30|sum:32 */
	uint32_t *run_start, *run_end; int *run_depth;   /* boundary records */
	uint64_t run_cap;
	unsigned long long *rec_base;    /* indexed by GLOBAL segment (pointer pre-offset by the launch's first segment) */
	uint32_t n_seg;
	unsigned long long *bins;        /* window bins (concatenated contigs) */
	uint32_t window; int variant;
	int rezero;
	int debug;                       /* bit 0: skip the look-back wait (timing experiments only, wrong results) */
	ngs_dev_state *st;
};

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
	unsigned long long v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v)
{
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

struct BinAcc {   /* per-thread pending (bin, value): consecutive runs of a thread share a window */
	unsigned long long *bins; long long k; long long acc;
	__device__ __forceinline__ void operator()(uint32_t kk, int64_t v)
	{
		if ((long long)kk == k) acc += v;
		else { if (k >= 0 && acc) atomicAdd(bins + k, (unsigned long long)acc); k = kk; acc = v; }
	}
};

__global__ void __launch_bounds__(NGS_SCAN_THREADS, NGS_SCAN_MINB) k_scan_runs(ScanArgs A)
{
	__shared__ __align__(16) int s_tile[NGS_SCAN_SMEM / 4];   /* phase 1: NGS_SCAN_THREADS x (16 + 4 pad) ints for the transposition
	                                                             (conflict-free 128-bit reads); phase 2: compact boundary list
	                                                             (u16 position [NGS_TILE] + i32 depth [NGS_TILE]).  Static on purpose:
	                                                             the dynamic-shared variant measured 45 % slower (2.49 vs 1.70 ms) */
	__shared__ long long s_wsum[NGS_SCAN_WARPS]; __shared__ unsigned int s_wcnt[NGS_SCAN_WARPS]; __shared__ unsigned int s_wlast[NGS_SCAN_WARPS];
	__shared__ unsigned int s_lb_inc[NGS_SCAN_WARPS], s_lb_last[NGS_SCAN_WARPS]; __shared__ long long s_lb_sum[NGS_SCAN_WARPS]; __shared__ unsigned long long s_lb_cnt[NGS_SCAN_WARPS];
	__shared__ long long s_bk[NGS_SCAN_WARPS], s_bv[NGS_SCAN_WARPS];
	__shared__ unsigned int s_wr[NGS_SCAN_WARPS]; __shared__ int s_wmax[NGS_SCAN_WARPS], s_wmin[NGS_SCAN_WARPS];
	__shared__ unsigned int s_rfirst[NGS_SCAN_WARPS], s_rlast[NGS_SCAN_WARPS];

	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	/* tile id = blockIdx.x: CTAs are dispatched in increasing block order, so every predecessor of a
	 * resident tile is resident or finished and the look-back cannot starve (no ticket round trip) */
	const uint64_t tile = blockIdx.x;                    /* launch-local */
	if (tile >= A.n_tiles) return;

	/* coalesced 128-bit loads (striped) issued first; the segment lookup overlaps their latency */
	int4 *gp = reinterpret_cast<int4 *>(A.diff + tile * (uint64_t)NGS_TILE);
	int4 v4[4];
	#pragma unroll
	for (int j = 0; j < 4; ++j) v4[j] = gp[j * NGS_SCAN_THREADS + t];
	const uint32_t si = A.tile2seg[tile];
	const ngs_seg sg = A.seg[si];
	/* transpose through shared memory to 16 consecutive elements per thread */
	#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const int q = j * NGS_SCAN_THREADS + t;              /* int4 index within the tile */
		*reinterpret_cast<int4 *>(&s_tile[(q >> 2) * 20 + (q & 3) * 4]) = v4[j];
		if (A.rezero) gp[q] = make_int4(0, 0, 0, 0);
	}
	__syncthreads();
	int d[16];
	#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const int4 v = *reinterpret_cast<const int4 *>(&s_tile[t * 20 + j * 4]);
		d[4 * j] = v.x; d[4 * j + 1] = v.y; d[4 * j + 2] = v.z; d[4 * j + 3] = v.w;
	}
	/* thread-local aggregate */
	int tsum = 0; unsigned int tcnt = 0, tlast = 0;         /* tlast: 1 + tile-relative index of last boundary */
	#pragma unroll
	for (int i = 0; i < 16; ++i) { tsum += d[i]; if (d[i] != 0) { ++tcnt; tlast = (unsigned)(t * 16 + i + 1); } }
	/* warp inclusive scans */
	int isum = tsum; unsigned int icnt = tcnt, ilast = tlast;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const int a = __shfl_up_sync(0xffffffffu, isum, o);
		const unsigned int b = __shfl_up_sync(0xffffffffu, icnt, o);
		const unsigned int c = __shfl_up_sync(0xffffffffu, ilast, o);
		if (lane >= o) { isum += a; icnt += b; ilast = ilast > c ? ilast : c; }
	}
	if (lane == 31) { s_wsum[warp] = isum; s_wcnt[warp] = icnt; s_wlast[warp] = ilast; }
	__syncthreads();
	long long wpre_sum = 0; unsigned int wpre_cnt = 0, wpre_last = 0;
	long long tot_sum = 0; unsigned int tot_cnt = 0, tot_last = 0;
	#pragma unroll
	for (int w = 0; w < NGS_SCAN_WARPS; ++w) {
		if (w < warp) { wpre_sum += s_wsum[w]; wpre_cnt += s_wcnt[w]; wpre_last = s_wlast[w] > wpre_last ? s_wlast[w] : wpre_last; }
		tot_sum += s_wsum[w]; tot_cnt += s_wcnt[w]; tot_last = s_wlast[w] > tot_last ? s_wlast[w] : tot_last;
	}
	/* exclusive prefix of this thread within the tile */
	const int psum = __shfl_up_sync(0xffffffffu, isum, 1);
	const unsigned int pcnt = __shfl_up_sync(0xffffffffu, icnt, 1);
	long long ex_sum = wpre_sum + (lane ? psum : 0);
	unsigned int ex_cnt = wpre_cnt + (lane ? pcnt : 0);

	/* tile-relative -> contig position of element 0 of this tile */
	const uint64_t tile_g = tile + A.tile_off;          /* global tile index, only for the segment geometry */
	const uint32_t pos0 = sg.clip_lo + (uint32_t)((tile_g - sg.tile0) * NGS_TILE);
	const uint32_t agg_lastb = tot_last ? pos0 + tot_last : 0;     /* 1 + contig position, 0 = none */

	/* ---- decoupled look-back, block-wide: thread i inspects predecessor tile-1-i, so one round covers
	 * 256 predecessors (with ~600 tiles in flight a warp-wide window needs ~19 dependent L2 round
	 * trips per tile, which is what bounded this kernel; 256-wide needs 2-3) */
	long long tile_sum = 0; unsigned long long tile_cnt = 0; unsigned int tile_last = 0;
	if (t == 0) {
		const unsigned long long stt = tile == 0 ? NGS_ST_INC : NGS_ST_AGG;
		st_relaxed_u64(&A.stA[tile], (stt << 62) | (unsigned long long)tot_cnt);
		st_relaxed_u64(&A.stB[tile], (stt << 62) | ((unsigned long long)agg_lastb << 32) | (uint32_t)(int)tot_sum);
	}
	if (tile > 0 && !(A.debug & 1)) {
		/* each thread inspects NGS_LB_PER consecutive predecessors, so one round covers
		 * NGS_SCAN_THREADS * NGS_LB_PER tiles: with ~1200 tiles in flight the nearest inclusive prefix is almost
		 * always inside the first round (each extra round costs two block barriers plus an L2 round trip) */
		long long base = (long long)tile - 1;
		for (;;) {
			unsigned long long a[NGS_LB_PER], b[NGS_LB_PER];
			#pragma unroll
			for (int j = 0; j < NGS_LB_PER; ++j) {
				const long long idx = base - ((long long)t * NGS_LB_PER + j);
				a[j] = NGS_ST_INC << 62; b[j] = NGS_ST_INC << 62;        /* virtual tiles before the launch's first: inclusive zero */
				if (idx >= 0 && t < NGS_LB_THREADS) { a[j] = ld_relaxed_u64(&A.stA[idx]); b[j] = ld_relaxed_u64(&A.stB[idx]); }
			}
			unsigned int my_inc = 0xffffffffu;                           /* position (0 = nearest tile) of this thread's first inclusive entry */
			#pragma unroll
			for (int j = 0; j < NGS_LB_PER; ++j) {
				const long long idx = base - ((long long)t * NGS_LB_PER + j);
				if (idx >= 0 && t < NGS_LB_THREADS) {
					while ((a[j] >> 62) == NGS_ST_EMPTY || (a[j] >> 62) != (b[j] >> 62)) {
						a[j] = ld_relaxed_u64(&A.stA[idx]); b[j] = ld_relaxed_u64(&A.stB[idx]);
					}
				}
				if (t < NGS_LB_THREADS && my_inc == 0xffffffffu && (a[j] >> 62) == NGS_ST_INC) my_inc = (unsigned)(t * NGS_LB_PER + j);
			}
			const unsigned int wmin = __reduce_min_sync(0xffffffffu, my_inc);
			if (lane == 0) s_lb_inc[warp] = wmin;
			__syncthreads();
			unsigned int first_inc = 0xffffffffu;                        /* nearest inclusive predecessor (entry position) */
			#pragma unroll
			for (int w = 0; w < NGS_SCAN_WARPS; ++w) first_inc = s_lb_inc[w] < first_inc ? s_lb_inc[w] : first_inc;
			long long vs = 0; unsigned long long vc = 0; unsigned int vl = 0;
			#pragma unroll
			for (int j = 0; j < NGS_LB_PER; ++j) {
				if (t < NGS_LB_THREADS && (unsigned)(t * NGS_LB_PER + j) <= first_inc) {
					vs += (long long)(int)(uint32_t)b[j];
					vc += a[j] & 0x3FFFFFFFFFFFFFFFull;
					if (vl == 0) vl = (unsigned int)((b[j] >> 32) & 0x3FFFFFFFu);   /* nearest non-zero boundary wins */
				}
			}
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) { vs += __shfl_xor_sync(0xffffffffu, vs, o); vc += __shfl_xor_sync(0xffffffffu, vc, o); }
			const unsigned int has = __ballot_sync(0xffffffffu, vl != 0);
			unsigned int near_l = 0;
			if (has) near_l = __shfl_sync(0xffffffffu, vl, __ffs(has) - 1);      /* lowest lane = nearest tiles */
			if (lane == 0) { s_lb_sum[warp] = vs; s_lb_cnt[warp] = vc; s_lb_last[warp] = near_l; }
			__syncthreads();
			unsigned int rl = 0;
			#pragma unroll
			for (int w = 0; w < NGS_SCAN_WARPS; ++w) { tile_sum += s_lb_sum[w]; tile_cnt += s_lb_cnt[w]; if (rl == 0) rl = s_lb_last[w]; }
			if (tile_last == 0) tile_last = rl;
			if (first_inc != 0xffffffffu) break;
			base -= NGS_LB_THREADS * NGS_LB_PER;
			__syncthreads();
		}
		if (t == 0) {
			const unsigned int inc_last = agg_lastb ? agg_lastb : tile_last;
			st_relaxed_u64(&A.stA[tile], (NGS_ST_INC << 62) | (tile_cnt + tot_cnt));
			st_relaxed_u64(&A.stB[tile], (NGS_ST_INC << 62) | ((unsigned long long)inc_last << 32) | (uint32_t)(int)(tile_sum + tot_sum));
		}
	}
	if (t == 0) {
		if (tile_g == sg.tile0) A.rec_base[si] = tile_cnt;
		if (tile == A.n_tiles - 1) {
			for (uint32_t s2 = si + 1; s2 <= A.n_seg; ++s2) A.rec_base[s2] = tile_cnt + tot_cnt;
			A.st->n_boundaries = tile_cnt + tot_cnt;
		}
	}
	if (tot_cnt == 0) return;

	/* ---- compact the tile's boundaries into shared memory (sparse pass, tiny body) ...
	 * the transposition buffer is dead by now (every thread holds its 16 elements in registers and two
	 * block barriers have passed), so it is reused for the compact (position, depth-before) list */
	unsigned short *c_pos = reinterpret_cast<unsigned short *>(s_tile);              /* [4096] tile-relative position */
	int *c_dep = s_tile + NGS_TILE / 2;                                               /* [4096] depth before the boundary */
	if (tcnt) {
		int depth = (int)(tile_sum + ex_sum);
		unsigned int r = ex_cnt;
		#pragma unroll
		for (int i = 0; i < 16; ++i) {
			if (d[i] != 0) { c_pos[r] = (unsigned short)(t * 16 + i); c_dep[r] = depth; depth += d[i]; ++r; }
		}
	}
	__syncthreads();
	/* ---- ... then emit them densely: thread r handles record r, r+256, .. (coalesced stores) */
	BinAcc bacc; bacc.bins = A.bins + sg.bins0; bacc.k = -1; bacc.acc = 0;
	unsigned int my_runs = 0, my_first = 0xffffffffu, my_last = 0; int my_max = 0, my_min = 0;
	const uint32_t W = A.window;
	const uint32_t kt = pos0 / W;                                   /* window of the tile's first position */
	const uint64_t wlo = (uint64_t)kt * W, whi = wlo + W - 1;        /* fast path: run inside [wlo, whi) and inside the contig */
	for (unsigned int r = t; r < tot_cnt; r += NGS_SCAN_THREADS) {
		const uint32_t pos = pos0 + c_pos[r];
		const int depth = c_dep[r];
		const uint32_t start = r ? pos0 + c_pos[r - 1] : (tile_last ? tile_last - 1 : 0);
		const unsigned long long rec = tile_cnt + r;
		if (rec < A.run_cap) { A.run_start[rec] = start; A.run_end[rec] = pos; A.run_depth[rec] = depth; }
		if (depth != 0) {
			if (start >= wlo && pos <= whi && pos <= sg.ref_len) bacc(kt, (int64_t)(pos - start) * depth);
			else ngs_bin_run(A.variant, start, pos, (int64_t)depth, W, sg.ref_len, sg.n_win, bacc);
			++my_runs; my_max = depth > my_max ? depth : my_max; my_min = depth < my_min ? depth : my_min;
			my_first = r < my_first ? r : my_first; my_last = r;
		}
	}
	/* ---- warp- then block-level aggregation of the pending window contribution and the statistics
	 * (one global atomic per distinct window per tile instead of one per thread) */
	{
		const unsigned full = 0xffffffffu;
		const bool has = bacc.k >= 0 && bacc.acc != 0;
		const unsigned hm = __ballot_sync(full, has);
		long long wk = -1, wv = 0;
		if (hm) {
			const int leader = __ffs(hm) - 1;
			const long long k0 = __shfl_sync(full, bacc.k, leader);
			if (__all_sync(full, !has || bacc.k == k0)) {
				long long v = has ? bacc.acc : 0;
				#pragma unroll
				for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
				wk = k0; wv = v;
			} else if (has) atomicAdd(bacc.bins + bacc.k, (unsigned long long)bacc.acc);
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
		for (int w = 0; w < NGS_SCAN_WARPS; ++w) {
			if (s_bk[w] >= 0) {
				if (s_bk[w] == ck) cv += s_bv[w];
				else { if (ck >= 0 && cv) atomicAdd(bacc.bins + ck, (unsigned long long)cv); ck = s_bk[w]; cv = s_bv[w]; }
			}
			nr += s_wr[w]; mx = s_wmax[w] > mx ? s_wmax[w] : mx; mn = s_wmin[w] < mn ? s_wmin[w] : mn;
			fr = s_rfirst[w] < fr ? s_rfirst[w] : fr; lr = s_rlast[w] > lr ? s_rlast[w] : lr;
		}
		if (ck >= 0 && cv) atomicAdd(bacc.bins + ck, (unsigned long long)cv);
		if (nr) {
			atomicAdd(&A.st->n_runs, nr);
			atomicMax(&A.st->max_depth, mx); atomicMin(&A.st->min_depth, mn);
			atomicMin(&A.st->first_run, tile_cnt + fr); atomicMax(&A.st->last_run, tile_cnt + lr);
		}
	}
}

/* ------------------------------------------------------------------------------------------------ */
/* K5: line formatting.  A "producer" maps an item index to (length, bytes).  One look-back scan over
 * tile byte counts gives each tile its output offset; lines are assembled in shared memory and copied
 * out with aligned 128-bit stores.                                                                   */

struct NameTable {
	const char *names;               /* concatenated "name\t" strings */
	const uint32_t *name_off;        /* [n_ref+1] */
	const unsigned long long *name8; /* [n_ref] first eight bytes of "name\t", little endian */
};

/* bedGraph: item = boundary record; gap records (depth 0) and the dropped halo run have length 0 */
struct BedProducer {
	const uint32_t *run_start, *run_end; const int *run_depth;
	const unsigned long long *rec_base; const ngs_seg *seg; uint32_t n_seg;
	NameTable nt;
	unsigned long long drop_idx, mod_idx; uint32_t mod_end;
	struct Item { uint32_t s, e; int d; uint32_t ctg; uint32_t nd; };     /* nd: digit counts packed 8 bits each */
	struct Cache { unsigned long long lo, hi; uint32_t ctg, name_len; unsigned long long name8; };
	__device__ __forceinline__ void init(Cache &c) const { c.lo = 1; c.hi = 0; c.ctg = 0; c.name_len = 0; c.name8 = 0; }
	__device__ __forceinline__ void locate(unsigned long long r, Cache &c) const
	{
		uint32_t lo = 0, hi = n_seg;                          /* last segment with rec_base <= r */
		while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (rec_base[mid] <= r) lo = mid; else hi = mid; }
		c.lo = rec_base[lo]; c.hi = rec_base[lo + 1]; c.ctg = seg[lo].contig;
		const uint32_t a = nt.name_off[c.ctg];
		c.name_len = nt.name_off[c.ctg + 1] - a;
		c.name8 = nt.name8[c.ctg];
	}
	__device__ __forceinline__ uint32_t length(unsigned long long r, Item &it, Cache &c) const
	{
		it.d = run_depth[r];
		if (it.d == 0 || r == drop_idx) { it.d = 0; return 0; }
		it.s = run_start[r]; it.e = (r == mod_idx) ? mod_end : run_end[r];
		if (r < c.lo || r >= c.hi) locate(r, c);
		it.ctg = c.ctg;
		const uint32_t ns = ngs_ndigits(it.s), ne = ngs_ndigits(it.e), nd = ngs_len_i32(it.d);
		it.nd = ns | ne << 8 | nd << 16;
		return c.name_len + ns + 1 + ne + 1 + nd + 1;
	}
	__device__ __forceinline__ void write(char *dst, const Item &it, const Cache &c) const
	{
		uint32_t nl;
		if (it.ctg == c.ctg && c.name_len <= 8) {            /* the common case: name bytes cached in a register */
			const uint32_t lo = (uint32_t)c.name8, hi = (uint32_t)(c.name8 >> 32);
			nl = c.name_len;                                   /* >= 1: at least the tab */
			dst[0] = (char)(lo & 0xff);
			if (nl > 1) dst[1] = (char)((lo >> 8) & 0xff);
			if (nl > 2) dst[2] = (char)((lo >> 16) & 0xff);
			if (nl > 3) dst[3] = (char)(lo >> 24);
			if (nl > 4) dst[4] = (char)(hi & 0xff);
			if (nl > 5) dst[5] = (char)((hi >> 8) & 0xff);
			if (nl > 6) dst[6] = (char)((hi >> 16) & 0xff);
			if (nl > 7) dst[7] = (char)(hi >> 24);
		} else {                                              /* long name, or a thread whose items straddle two contigs */
			const uint32_t a = nt.name_off[it.ctg];
			nl = nt.name_off[it.ctg + 1] - a;
			for (uint32_t i = 0; i < nl; ++i) dst[i] = nt.names[a + i];
		}
		dst += nl;
		const uint32_t ns = it.nd & 0xff, ne = (it.nd >> 8) & 0xff;
		ngs_put_u32_swar(dst, it.s, ns); dst += ns; *dst++ = '\t';
		ngs_put_u32_swar(dst, it.e, ne); dst += ne; *dst++ = '\t';
		dst += ngs_put_i32(dst, it.d); *dst = '\n';
	}
};

/* .depth: item = window (every window of every contig) */
struct DepthProducer {
	const unsigned long long *bins; const unsigned long long *win_base;   /* [n_ref+1] first item of contig */
	const uint32_t *ref_len; uint32_t n_ref; uint32_t window; NameTable nt;
	struct Item { uint32_t ws, we, ctg; unsigned long long h; };
	struct Cache { int unused; };
	__device__ __forceinline__ void init(Cache &) const {}
	__device__ __forceinline__ uint32_t ctg_of(unsigned long long r) const
	{
		uint32_t lo = 0, hi = n_ref;
		while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (win_base[mid] <= r) lo = mid; else hi = mid; }
		return lo;
	}
	__device__ __forceinline__ uint32_t length(unsigned long long r, Item &it, Cache &) const
	{
		it.ctg = ctg_of(r);
		const unsigned long long k = r - win_base[it.ctg];
		const unsigned long long ws = k * window, we = ws + window;
		it.ws = (uint32_t)ws; it.we = we > ref_len[it.ctg] ? ref_len[it.ctg] : (uint32_t)we;
		it.h = ngs_hundredths(__ddiv_rn((double)(long long)bins[r], (double)window));
		return (nt.name_off[it.ctg + 1] - nt.name_off[it.ctg]) + ngs_ndigits(it.ws) + 1 + ngs_ndigits(it.we) + 1 + ngs_len_f2(it.h) + 1;
	}
	__device__ __forceinline__ void write(char *dst, const Item &it, const Cache &) const
	{
		const uint32_t a = nt.name_off[it.ctg], nl = nt.name_off[it.ctg + 1] - a;
		for (uint32_t i = 0; i < nl; ++i) dst[i] = nt.names[a + i];
		dst += nl;
		dst += ngs_put_u32(dst, it.ws); *dst++ = '\t';
		dst += ngs_put_u32(dst, it.we); *dst++ = '\t';
		dst += ngs_put_f2(dst, it.h); *dst = '\n';
	}
};

/* wig: per contig one header item followed by one item per window (length 0 when the bin is empty) */
struct WigProducer {
	const unsigned long long *bins; const unsigned long long *win_base;   /* [n_ref+1] first BIN of contig */
	uint32_t n_ref; uint32_t window;
	const char *hdr; const uint32_t *hdr_off;                             /* "variableStep chrom=%s span=%d\n" per contig */
	struct Item { uint32_t ws, ctg; int is_hdr; unsigned long long h; };
	struct Cache { int unused; };
	__device__ __forceinline__ void init(Cache &) const {}
	__device__ __forceinline__ uint32_t length(unsigned long long r, Item &it, Cache &) const
	{
		/* item space: contig c owns items [win_base[c] + c, win_base[c+1] + c + 1) */
		uint32_t lo = 0, hi = n_ref;
		while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (win_base[mid] + mid <= r) lo = mid; else hi = mid; }
		it.ctg = lo;
		const unsigned long long k = r - (win_base[lo] + lo);
		if (k == 0) { it.is_hdr = 1; return hdr_off[lo + 1] - hdr_off[lo]; }
		it.is_hdr = 0;
		const unsigned long long b = bins[win_base[lo] + k - 1];
		if (b == 0) { it.h = ~0ull; return 0; }
		it.ws = (uint32_t)((k - 1) * window);
		it.h = ngs_hundredths(__ddiv_rn((double)(long long)b, (double)window));
		return ngs_ndigits(it.ws) + 1 + ngs_len_f2(it.h) + 1;
	}
	__device__ __forceinline__ void write(char *dst, const Item &it, const Cache &) const
	{
		if (it.is_hdr) {
			const uint32_t a = hdr_off[it.ctg], nl = hdr_off[it.ctg + 1] - a;
			for (uint32_t i = 0; i < nl; ++i) dst[i] = hdr[a + i];
			return;
		}
		dst += ngs_put_u32(dst, it.ws); *dst++ = '\t';
		dst += ngs_put_f2(dst, it.h); *dst = '\n';
	}
};

struct FmtArgs {
	unsigned long long n_items;
	unsigned long long n_tiles;
	unsigned long long *status;        /* [n_tiles] status:2 | bytes:62 */
	char *out; unsigned long long out_cap;
	ngs_dev_state *st;
	uint32_t stage_bytes;              /* dynamic shared memory for line staging */
	int debug;                         /* bit 0: skip the look-back wait (timing experiments only, wrong output) */
};

template <class Producer, int IPT>
__global__ void __launch_bounds__(NGS_FMT_THREADS, NGS_FMT_MINB) k_format_lines(FmtArgs A, Producer P)
{
	extern __shared__ __align__(16) char s_stage[];
	__shared__ unsigned int s_w[8], s_li[8];
	__shared__ unsigned long long s_lb[8];
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	const unsigned long long tile = blockIdx.x;          /* dispatch order = tile order, see k_scan_runs */
	if (tile >= A.n_tiles) return;
	const unsigned long long r0 = tile * (unsigned long long)(NGS_FMT_THREADS * IPT) + (unsigned long long)t * IPT;
	typename Producer::Item it[IPT];
	typename Producer::Cache cache;
	P.init(cache);
	uint32_t len[IPT], tlen = 0;
	#pragma unroll
	for (int i = 0; i < IPT; ++i) {
		len[i] = (r0 + i < A.n_items) ? P.length(r0 + i, it[i], cache) : 0;
		tlen += len[i];
	}
	/* block exclusive scan of tlen */
	uint32_t inc = tlen;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
	if (lane == 31) s_w[warp] = inc;
	__syncthreads();
	uint32_t wpre = 0, total = 0;
	#pragma unroll
	for (int w = 0; w < 8; ++w) { if (w < warp) wpre += s_w[w]; total += s_w[w]; }
	uint32_t off = wpre + inc - tlen;
	/* look-back over tile byte totals, block-wide (thread i inspects tile-1-i) */
	unsigned long long run = 0;
	if (t == 0) st_relaxed_u64(&A.status[tile], ((tile == 0 ? NGS_ST_INC : NGS_ST_AGG) << 62) | total);
	if (tile > 0) {
		long long base = (long long)tile - 1;
		for (;;) {
			const long long idx = base - t;
			unsigned long long a = NGS_ST_INC << 62;
			if (idx >= 0) { do { a = ld_relaxed_u64(&A.status[idx]); } while ((a >> 62) == NGS_ST_EMPTY); }
			const unsigned int im = __ballot_sync(0xffffffffu, (a >> 62) == NGS_ST_INC);
			if (lane == 0) s_li[warp] = im ? (unsigned)(warp * 32 + __ffs(im) - 1) : 0xffffffffu;
			__syncthreads();
			unsigned int first_inc = 0xffffffffu;
			#pragma unroll
			for (int w = 0; w < 8; ++w) first_inc = s_li[w] < first_inc ? s_li[w] : first_inc;
			unsigned long long v = (unsigned)t <= first_inc ? (a & 0x3FFFFFFFFFFFFFFFull) : 0;
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
			if (lane == 0) s_lb[warp] = v;
			__syncthreads();
			#pragma unroll
			for (int w = 0; w < 8; ++w) run += s_lb[w];
			if (first_inc != 0xffffffffu) break;
			base -= NGS_FMT_THREADS;
			__syncthreads();
		}
		if (t == 0) st_relaxed_u64(&A.status[tile], (NGS_ST_INC << 62) | (run + total));
	}
	if (t == 0 && tile == A.n_tiles - 1) A.st->fmt_bytes = run + total;
	const unsigned long long goff = run;
	if (total == 0 || goff + total > A.out_cap) return;
	/* assemble lines in shared memory, shifted so that 16-byte alignment matches the global address */
	const uint32_t mis = (uint32_t)((uintptr_t)(A.out + goff) & 15);
	if (total + mis <= A.stage_bytes) {
		#pragma unroll
		for (int i = 0; i < IPT; ++i) { if (len[i]) P.write(s_stage + mis + off, it[i], cache); off += len[i]; }
		__syncthreads();
		char *g = A.out + goff - mis;                       /* 16-byte aligned */
		const uint32_t end = mis + total;
		uint32_t a0 = (mis + 15u) & ~15u;                   /* head bytes [mis,a0), 128-bit body [a0,a1), tail [a1,end) */
		if (a0 > end) a0 = end;
		uint32_t a1 = end & ~15u;
		if (a1 < a0) a1 = a0;
		for (uint32_t i = mis + t; i < a0; i += NGS_FMT_THREADS) g[i] = s_stage[i];
		for (uint32_t i = a0 + (uint32_t)t * 16; i < a1; i += NGS_FMT_THREADS * 16)
			*reinterpret_cast<int4 *>(g + i) = *reinterpret_cast<const int4 *>(s_stage + i);
		for (uint32_t i = a1 + t; i < end; i += NGS_FMT_THREADS) g[i] = s_stage[i];
	} else {
		/* lines too long for the staging area (very long contig names): write straight to global */
		#pragma unroll
		for (int i = 0; i < IPT; ++i) { if (len[i]) P.write(A.out + goff + off, it[i], cache); off += len[i]; }
	}
}

/* K5, bedGraph fast path.  Same tiling, byte-offset look-back and aligned copy-out as k_format_lines, but the
 * lines are assembled in registers and leave the thread as whole 32-bit words (ngs_lw in ngscov_hd.h):
 *   - the IPT consecutive boundary records of a thread are loaded with 128-bit loads; record r starts where
 *     record r-1 ends, so each position is converted to ASCII once and used for both lines;
 *   - the digit count falls out of the conversion (lowest non-zero byte), no separate compare tree;
 *   - the contig is found once per warp (a tile inside one contig is the only case on this path).
 * Tiles that straddle two contigs, hold the halo-dropped / halo-extended run, have a "name\t" longer than
 * eight bytes or a depth outside [0, 10^9) take the generic BedProducer code (bed_slow_*), which produces the
 * same lengths by construction. */
__device__ __noinline__ uint4 bed_slow_len4(const BedProducer &P, unsigned long long r0, unsigned long long n_items)
{
	BedProducer::Cache c; P.init(c);
	BedProducer::Item it;
	uint4 l;
	l.x = (r0 + 0 < n_items) ? P.length(r0 + 0, it, c) : 0;
	l.y = (r0 + 1 < n_items) ? P.length(r0 + 1, it, c) : 0;
	l.z = (r0 + 2 < n_items) ? P.length(r0 + 2, it, c) : 0;
	l.w = (r0 + 3 < n_items) ? P.length(r0 + 3, it, c) : 0;
	return l;
}
__device__ __noinline__ void bed_slow_write(const BedProducer &P, unsigned long long r0, unsigned long long n_items, int ipt, char *dst)
{
	BedProducer::Cache c; P.init(c);
	for (int i = 0; i < ipt; ++i) {
		BedProducer::Item it;
		const uint32_t l = (r0 + i < n_items) ? P.length(r0 + i, it, c) : 0;
		if (l) P.write(dst, it, c);
		dst += l;
	}
}

template <int IPT, int MINB>
__global__ void __launch_bounds__(NGS_FMT_THREADS, MINB) k_format_bed(const __grid_constant__ FmtArgs A, const __grid_constant__ BedProducer P)
{
	extern __shared__ __align__(16) char s_stage[];
	__shared__ unsigned int s_w[8], s_li[8];
	__shared__ unsigned long long s_lb[8];
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	const unsigned long long tile = blockIdx.x;
	if (tile >= A.n_tiles) return;
	const unsigned long long tr0 = tile * (unsigned long long)(NGS_FMT_THREADS * IPT), r0 = tr0 + (unsigned long long)t * IPT;
	uint32_t S[IPT], E[IPT]; int D[IPT];
	if (r0 + IPT <= A.n_items) {
		#pragma unroll
		for (int j = 0; j < IPT / 4; ++j) {
			const uint4 a = reinterpret_cast<const uint4 *>(P.run_start + r0)[j], b = reinterpret_cast<const uint4 *>(P.run_end + r0)[j];
			const int4 c = reinterpret_cast<const int4 *>(P.run_depth + r0)[j];
			S[4 * j] = a.x; S[4 * j + 1] = a.y; S[4 * j + 2] = a.z; S[4 * j + 3] = a.w;
			E[4 * j] = b.x; E[4 * j + 1] = b.y; E[4 * j + 2] = b.z; E[4 * j + 3] = b.w;
			D[4 * j] = c.x; D[4 * j + 1] = c.y; D[4 * j + 2] = c.z; D[4 * j + 3] = c.w;
		}
	} else {
		#pragma unroll
		for (int i = 0; i < IPT; ++i) {
			const bool in = r0 + i < A.n_items;
			S[i] = in ? P.run_start[r0 + i] : 0; E[i] = in ? P.run_end[r0 + i] : 0; D[i] = in ? P.run_depth[r0 + i] : 0;
		}
	}
	/* contig of the tile: last segment whose first record is <= the tile's first (rec_base is non-decreasing) */
	uint32_t sgi;
	if (P.n_seg <= 32u) {
		const bool le = (uint32_t)lane < P.n_seg && P.rec_base[lane] <= tr0;          /* rec_base[0] = 0: lane 0 always votes */
		sgi = 31u - (uint32_t)__clz((int)__ballot_sync(0xffffffffu, le));
	} else {
		sgi = 0;
		for (uint32_t i = lane; i < P.n_seg; i += 32) if (P.rec_base[i] <= tr0) sgi = i;
		sgi = __reduce_max_sync(0xffffffffu, sgi);
	}
	unsigned long long tr1 = tr0 + (unsigned long long)(NGS_FMT_THREADS * IPT);
	if (tr1 > A.n_items) tr1 = A.n_items;
	const uint32_t ctg = P.seg[sgi].contig;
	const uint32_t nl = P.nt.name_off[ctg + 1] - P.nt.name_off[ctg];
	const unsigned long long name8 = P.nt.name8[ctg];
	const bool slow_u = tr1 > P.rec_base[sgi + 1] || nl > 8u || (P.drop_idx >= tr0 && P.drop_idx < tr1) || (P.mod_idx >= tr0 && P.mod_idx < tr1);

	/* ---- lengths.  Positions: p = H * 10^4 + L, H cached per thread (ngs_hcache), L converted once per distinct
	 * position (record i starts where record i-1 ends).  Depths below 10^4 are converted here and kept as a
	 * ready piece; larger ones are converted again when written. */
	ngs_hcache hc; hc.H = 0xffffffffu; hc.hlo = hc.sh = hc.nH = hc.selA = hc.selB = hc.selC = 0;
	hc.nh0 = hc.nh1 = hc.nh2 = hc.nh3 = hc.th0 = hc.th1 = 0;
	const uint32_t n_lo = (uint32_t)name8, n_hi = (uint32_t)(name8 >> 32);
	uint32_t L4e[IPT], L4s0 = 0;
	uint32_t len[IPT], tlen = 0;
	int bad = 0;
	auto pos_digits = [&](uint32_t p, uint32_t &L4) -> uint32_t {
		const uint32_t H = p / 10000u, raw = ngs_digits4_raw(p - H * 10000u);
		L4 = raw + 0x30303030u;
		if (H == 0) return raw ? 4u - ngs_low_byte_index(raw) : 1u;
		if (H >= 100000u) { bad = 1; return 10u; }
		if (H != hc.H) ngs_hcache_fill(hc, H, n_lo, n_hi, nl);
		return hc.nH + 4u;
	};
	if (!slow_u) {
		bool any = false;
		#pragma unroll
		for (int i = 0; i < IPT; ++i) { any |= D[i] != 0; len[i] = 0; L4e[i] = 0; }
		if (any) {
			uint32_t n_prev = pos_digits(S[0], L4s0);
			#pragma unroll
			for (int i = 0; i < IPT; ++i) {
				uint32_t ns = n_prev;
				if (i > 0 && S[i] != E[i - 1] && D[i] != 0) { uint32_t tmp; ns = pos_digits(S[i], tmp); }
				const uint32_t ne = pos_digits(E[i], L4e[i]);
				n_prev = ne;
				if (D[i] != 0) {
					uint32_t nd;
					if ((uint32_t)D[i] < 10000u) nd = 1u + ((uint32_t)D[i] >= 10u) + ((uint32_t)D[i] >= 100u) + ((uint32_t)D[i] >= 1000u);
					else if ((uint32_t)D[i] < 1000000000u) nd = ngs_ndigits((uint32_t)D[i]);
					else { bad = 1; nd = ngs_len_i32(D[i]); }
					len[i] = nl + ns + ne + nd + 3u;
				}
				tlen += len[i];
			}
		}
	} else {
		#pragma unroll
		for (int j = 0; j < IPT / 4; ++j) {
			const uint4 l = bed_slow_len4(P, r0 + 4 * j, A.n_items);
			len[4 * j] = l.x; len[4 * j + 1] = l.y; len[4 * j + 2] = l.z; len[4 * j + 3] = l.w;
			tlen += l.x + l.y + l.z + l.w;
		}
	}

	/* block exclusive scan of tlen; the barrier also tells every thread whether the tile must go the slow way */
	uint32_t inc = tlen;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
	if (lane == 31) s_w[warp] = inc;
	const bool slow = __syncthreads_or(bad) != 0 || slow_u;
	uint32_t wpre = 0, total = 0;
	#pragma unroll
	for (int w = 0; w < 8; ++w) { if (w < warp) wpre += s_w[w]; total += s_w[w]; }
	const uint32_t off = wpre + inc - tlen;
	/* ---- order of the remaining work: publish the tile's byte count, ASSEMBLE the lines, and only then look back.
	 * Looking back right after the lengths (as k_format_lines does) makes every tile wait for neighbours that were
	 * dispatched a moment earlier and are still computing their own lengths: 37 % of the stall samples, and 1.86 ms
	 * instead of 1.19 ms with the wait removed (62 M reads).  After the assembly the predecessors' counts exist and a
	 * tile dispatched a microsecond earlier already holds an inclusive prefix, so one round almost always suffices.
	 * The price: the tile's global offset -- its 16-byte misalignment -- is not known while the lines are laid out,
	 * so they start at byte 0 of the staging area and the copy-out realigns them with byte permutes.
	 * (Also measured: look-back by warp 0 alone while warps 1..7 assemble -- 2.18 ms, the look-back stays on warp 0's
	 * critical path and the others wait for it at the barrier.) */
	uint32_t *sw = reinterpret_cast<uint32_t *>(s_stage);
	const bool staged = total + 32u <= A.stage_bytes;
	if (t == 0) {
		st_relaxed_u64(&A.status[tile], ((tile == 0 ? NGS_ST_INC : NGS_ST_AGG) << 62) | total);
		if (staged && !slow) sw[total >> 2] = 0;         /* the tile's last, partial word is only ever OR-ed into */
	}
	ngs_lw<uint32_t *> L;
	L.begin(sw, off);
	if (staged && total) {
		if (slow) bed_slow_write(P, r0, A.n_items, IPT, s_stage + off);
		else if (tlen) {
			#pragma unroll
			for (int i = 0; i < IPT; ++i) {
				if (len[i]) {
					if (i == 0) L.bed_line_cached(hc, n_lo, n_hi, nl, S[0], L4s0, E[0], L4e[0], (uint32_t)D[0]);
					else if (S[i] == E[i - 1]) L.bed_line_cached(hc, n_lo, n_hi, nl, E[i - 1], L4e[i - 1], E[i], L4e[i], (uint32_t)D[i]);
					else L.bed_line_cached(hc, n_lo, n_hi, nl, S[i], ngs_ascii4(S[i] % 10000u), E[i], L4e[i], (uint32_t)D[i]);
				}
			}
		}
	}
	__syncthreads();                                        /* every head word stored */
	if (staged && !slow && tlen && L.k) atomicOr(L.w, L.carry);   /* tail bytes into the word whose owner stored its head part */
	/* look-back over tile byte totals, block-wide (thread i inspects tile-1-i); its barriers also order the ORs
	 * above before the copy-out (tile 0 gets an explicit one) */
	unsigned long long run = 0;
	if (tile > 0 && !(A.debug & 1)) {
		long long base = (long long)tile - 1;
		for (;;) {
			const long long idx = base - t;
			unsigned long long a = NGS_ST_INC << 62;
			if (idx >= 0) { do { a = ld_relaxed_u64(&A.status[idx]); } while ((a >> 62) == NGS_ST_EMPTY); }
			const unsigned int im = __ballot_sync(0xffffffffu, (a >> 62) == NGS_ST_INC);
			if (lane == 0) s_li[warp] = im ? (unsigned)(warp * 32 + __ffs(im) - 1) : 0xffffffffu;
			__syncthreads();
			unsigned int first_inc = 0xffffffffu;
			#pragma unroll
			for (int w = 0; w < 8; ++w) first_inc = s_li[w] < first_inc ? s_li[w] : first_inc;
			unsigned long long v = (unsigned)t <= first_inc ? (a & 0x3FFFFFFFFFFFFFFFull) : 0;
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
			if (lane == 0) s_lb[warp] = v;
			__syncthreads();
			#pragma unroll
			for (int w = 0; w < 8; ++w) run += s_lb[w];
			if (first_inc != 0xffffffffu) break;
			base -= NGS_FMT_THREADS;
			__syncthreads();
		}
		if (t == 0) st_relaxed_u64(&A.status[tile], (NGS_ST_INC << 62) | (run + total));
	} else __syncthreads();
	if (A.debug & 1) run = tile * 24000ull;              /* timing experiment only: no wait, wrong offsets */
	if (t == 0 && tile == A.n_tiles - 1) A.st->fmt_bytes = run + total;
	const unsigned long long goff = run;
	if (total == 0 || goff + total > A.out_cap) return;
	if (!staged) {                                      /* lines too long for the staging area: straight to global */
		bed_slow_write(P, r0, A.n_items, IPT, A.out + goff + off);
		return;
	}
	/* ---- copy-out: head bytes up to the first 16-byte boundary of the output, then aligned 128-bit stores whose
	 * source starts (16 - mis) bytes into the staging area, then the tail bytes */
	char *g = A.out + goff;
	const uint32_t mis = (uint32_t)((uintptr_t)g & 15);
	uint32_t hb = (16u - mis) & 15u;
	if (hb > total) hb = total;
	if ((uint32_t)t < hb) g[t] = s_stage[t];
	const uint32_t nv = (total - hb) >> 4;
	int4 *dst = reinterpret_cast<int4 *>(g + hb);
	const int4 *src = reinterpret_cast<const int4 *>(s_stage);
	if (hb == 0) {
		for (uint32_t c = t; c < nv; c += NGS_FMT_THREADS) dst[c] = src[c];
	} else {
		const uint32_t wsh = hb >> 2, sel = 0x3210u + 0x1111u * (hb & 3u);
		for (uint32_t c = t; c < nv; c += NGS_FMT_THREADS) {
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
	if (tb + (uint32_t)t < total) g[tb + t] = s_stage[tb + t];
}

/* ------------------------------------------------------------------------------------------------ */
/* K0: BGZF inflate (replaces samtools-0.1.19 bgzf.c:214-238 inflate_block on the host).
 * One warp per BGZF block.  The 32 lanes run the sequential DEFLATE decoder of ngscov_inflate.h in lock step on
 * identical values (Huffman tables in the warp's slice of shared memory, table loads are broadcasts), so there is
 * no divergence and nothing to exchange; only the bytes of a match or of a stored block are spread over the
 * lanes.  A match of length len at distance dist copies byte i from pos - dist + (i mod dist): every source byte
 * lies before pos, so the lanes need no ordering among themselves within one copy; __syncwarp() orders a copy
 * after the literals and copies that produced its source. */
struct ngs_bgzf_desc { unsigned long long src, dst; uint32_t clen, ulen; };

/* Output policy of one decoder = a group of W lanes (W = 32, 16 or 8: one, two or four decoders per warp).  The lanes
 * of a group run the decode in lock step on identical values; the groups of a warp work on different BGZF blocks and
 * share the warp's instruction stream wherever they happen to be at the same point (they reconverge at the head of
 * the symbol loop), which is where the other lanes of the warp pay off: the decode is a chain of dependent
 * instructions, and the table size, not the arithmetic, limits how many chains an SM can keep in flight. */
template <int W>
struct GroupInflateOut {
	static constexpr int K = 32 / W;     /* bytes per lane of a deferred match */
	uint8_t *o; int lane; unsigned gmask; uint32_t ltab, dtab;   /* ltab, dtab: shared-window addresses of the length / distance tables */
	/* a match of up to 32 bytes is stored one symbol late: its loads are issued, the next symbol is decoded while
	 * the bytes travel, and only then does the group touch the loaded values */
	uint32_t pend_pos, pend_len, pend_val;
	__device__ __forceinline__ uint32_t lentry(uint32_t i) const { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ltab + 4u * i)); return v; }
	__device__ __forceinline__ uint32_t dentry(uint32_t j) const { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(dtab + 4u * j)); return v; }
	__device__ __forceinline__ void finish()
	{
		if (pend_len) {
			#pragma unroll
			for (int k = 0; k < K; ++k) { const uint32_t i = (uint32_t)(lane + k * W); if (i < pend_len) o[pend_pos + i] = (uint8_t)(pend_val >> (8 * k)); }
			pend_len = 0;
		}
	}
	/* the block's Huffman tables are built by the group's first lane alone (all lanes writing the same values to the
	 * same shared words is harmless but is what compute-sanitizer's racecheck reports); the other lanes take over
	 * the bit reader it advanced */
	__device__ __forceinline__ int tables(ngs_inf_bits &b, uint32_t btype, ngs_inf_tables &T)
	{
		int r = 0;
		__syncwarp(gmask);
		if (lane == 0) r = ngs_inf_block_tables(b, btype, T);
		__syncwarp(gmask);
		r = __shfl_sync(gmask, r, 0, W);
		b.w0 = __shfl_sync(gmask, b.w0, 0, W); b.w1 = __shfl_sync(gmask, b.w1, 0, W); b.w2 = __shfl_sync(gmask, b.w2, 0, W);
		b.bp = __shfl_sync(gmask, b.bp, 0, W); b.nw = __shfl_sync(gmask, b.nw, 0, W);
		return r;
	}
	__device__ __forceinline__ void literal(uint32_t pos, uint32_t b) { if (lane == 0) o[pos] = (uint8_t)b; }
	__device__ __forceinline__ void match(uint32_t pos, uint32_t len, uint32_t dist)
	{
		finish();                       /* the previous match may be this one's source */
		__syncwarp(gmask);
		const uint32_t s0 = pos - dist;
		if (dist >= len) {
			if (len <= 32u) {
				uint32_t v = 0;
				#pragma unroll
				for (int k = 0; k < K; ++k) { const uint32_t i = (uint32_t)(lane + k * W); if (i < len) v |= (uint32_t)o[s0 + i] << (8 * k); }
				pend_val = v; pend_pos = pos; pend_len = len;
				return;
			}
			for (uint32_t i = lane; i < len; i += W) o[pos + i] = o[s0 + i];
		}
		else if (dist == 1u) { const uint8_t v = o[s0]; for (uint32_t i = lane; i < len; i += W) o[pos + i] = v; }
		else { for (uint32_t i = lane; i < len; i += W) o[pos + i] = o[s0 + i % dist]; }
	}
	__device__ __forceinline__ void stored(uint32_t pos, const uint8_t *src, uint32_t len)
	{
		for (uint32_t i = lane; i < len; i += W) o[pos + i] = src[i];
	}
};

#define NGS_INF_WARPS 4
#ifndef NGS_INF_MINB
#define NGS_INF_MINB 12                  /* 40 registers: 48 warps per SM (measured best of 1 / 12 / 16) */
#endif
#ifndef NGS_INF_LS_MINB
#define NGS_INF_LS_MINB(S) ((S) == 2 ? 8 : 4)
#endif
/* one decoder per group of 32 / S lanes (the library uses S = 1: a warp per BGZF block), NGS_INF_WARPS warps per CTA */
template <int S>
__global__ void __launch_bounds__(NGS_INF_WARPS * 32, NGS_INF_MINB) k_bgzf_inflate(const uint8_t *comp, const ngs_bgzf_desc *desc, uint32_t n_blk, uint8_t *out, ngs_dev_state *st)
{
	constexpr int W = 32 / S;
	__shared__ ngs_inf_tables s_T[NGS_INF_WARPS * S];
	__shared__ uint32_t s_ltab[32], s_dtab[32];
	if (threadIdx.x < 32) { s_ltab[threadIdx.x] = ngs_inf_len_entry(threadIdx.x); s_dtab[threadIdx.x] = ngs_inf_dist_entry(threadIdx.x); }
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / W;
	const unsigned gmask = S == 1 ? 0xffffffffu : ((1u << W) - 1u) << (grp * W);
	const uint32_t first = (blockIdx.x * NGS_INF_WARPS + warp) * S + grp, stride = gridDim.x * NGS_INF_WARPS * S;
	const uint32_t ltab = (uint32_t)__cvta_generic_to_shared(s_ltab), dtab = (uint32_t)__cvta_generic_to_shared(s_dtab);
	for (uint32_t b = first; b < n_blk; b += stride) {
		const ngs_bgzf_desc d = desc[b];
		GroupInflateOut<W> o = { out + d.dst, lane % W, gmask, ltab, dtab, 0u, 0u, 0u };
		const long r = ngs_inflate_stream(comp + d.src, d.clen, d.ulen, s_T[warp * S + grp], o);
		if (r != (long)d.ulen && o.lane == 0) atomicOr(&st->flags, NGS_FLAG_EFORMAT);
		__syncwarp(gmask);
	}
}

/* The same decoders as a warp-wide state machine: S groups of 32/S lanes, one BGZF block each, and ONE symbol per group
 * and pass of the loop.  The __syncwarp at the head of the loop brings the groups back together, so the table lookup
 * and whichever of the two paths (literal / match) the groups share are issued once for all of them: the decode is
 * bound by instruction issue, and this is how the other lanes of the warp take part in it.  Block headers, table
 * builds and block changes are passes of their own (rare: one per ~5000 symbols). */
enum { NGS_LS_OPEN = 0, NGS_LS_HDR = 1, NGS_LS_SYM = 2, NGS_LS_END = 3, NGS_LS_FAIL = 4, NGS_LS_IDLE = 5 };
template <int S>
__global__ void __launch_bounds__(NGS_INF_WARPS * 32, NGS_INF_LS_MINB(S)) k_bgzf_inflate_ls(const uint8_t *comp, const ngs_bgzf_desc *desc, uint32_t n_blk, uint8_t *out, ngs_dev_state *st)
{
	constexpr int W = 32 / S;
	__shared__ ngs_inf_tables s_T[NGS_INF_WARPS * S];
	__shared__ uint32_t s_ltab[32], s_dtab[32];
	if (threadIdx.x < 32) { s_ltab[threadIdx.x] = ngs_inf_len_entry(threadIdx.x); s_dtab[threadIdx.x] = ngs_inf_dist_entry(threadIdx.x); }
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / W;
	const unsigned gmask = ((1u << (W - 1)) | ((1u << (W - 1)) - 1u)) << (grp * W);
	const uint32_t stride = gridDim.x * NGS_INF_WARPS * S;
	uint32_t blk = (blockIdx.x * NGS_INF_WARPS + warp) * S + grp;
	ngs_inf_tables &T = s_T[warp * S + grp];
	GroupInflateOut<W> o = { out, lane % W, gmask, (uint32_t)__cvta_generic_to_shared(s_ltab), (uint32_t)__cvta_generic_to_shared(s_dtab), 0u, 0u, 0u };
	ngs_inf_bits b; b.wbase = nullptr; b.w0 = b.w1 = b.w2 = b.bp = b.nw = b.lead = b.start = 0;
	const uint8_t *in = comp;
	uint32_t in_len = 0, out_len = 0, pos = 0, last = 0;
	int state = blk < n_blk ? NGS_LS_OPEN : NGS_LS_IDLE;
	for (;;) {
		__syncwarp();
		if (__all_sync(0xffffffffu, state == NGS_LS_IDLE)) break;
		if (state == NGS_LS_SYM) {
			uint32_t x = ngs_inf_peek(b);
			const uint32_t e = T.lfast[x & ((1u << NGS_INF_LBITS) - 1u)];
			uint32_t cl = e & 15u, sym = e >> 4;
			if (!cl) { const int s = ngs_inf_decode_slow(x, T.lcount, T.lsym, NGS_INF_LBITS, T.lslow, &cl); if (s < 0) { sym = 256u; cl = 0; state = NGS_LS_FAIL; } else sym = (uint32_t)s; }
			if (sym < 256u) {
				ngs_inf_skip(b, cl);
				if (pos >= out_len) state = NGS_LS_FAIL;
				else { o.literal(pos, sym); ++pos; }
			} else if (sym == 256u) {
				if (state != NGS_LS_FAIL) {
					ngs_inf_skip(b, cl);
					state = ngs_inf_consumed(b) > in_len ? NGS_LS_FAIL : (last ? NGS_LS_END : NGS_LS_HDR);
				}
			} else {
				const uint32_t le = o.lentry(sym - 257u);
				const uint32_t len = (le & 0xffffffu) + ((x >> cl) & ((1u << (le >> 24)) - 1u));
				ngs_inf_skip(b, cl + (le >> 24));
				x = ngs_inf_peek(b);
				const uint32_t d = T.dfast[x & ((1u << NGS_INF_DBITS) - 1u)];
				uint32_t cd = d & 15u, ds = d >> 4;
				if (!cd) { const int s = ngs_inf_decode_slow(x, T.dcount, T.dsym, NGS_INF_DBITS, T.dslow, &cd); if (s < 0) { ds = 31u; cd = 0; } else ds = (uint32_t)s; }
				const uint32_t de = o.dentry(ds);
				const uint32_t dist = (de & 0xffffffu) + ((x >> cd) & ((1u << (de >> 24)) - 1u));
				ngs_inf_skip(b, cd + (de >> 24));
				if (dist > pos || pos + len > out_len) state = NGS_LS_FAIL;    /* also: symbols that must not occur */
				else { o.match(pos, len, dist); pos += len; }
			}
		} else if (state == NGS_LS_HDR) {
			const uint32_t hdr = ngs_inf_get(b, 3);
			last = hdr & 1u;
			const uint32_t btype = hdr >> 1;
			if (btype == 0u) {
				ngs_inf_skip(b, (8u - (b.bp & 7u)) & 7u);
				const uint32_t len = ngs_inf_get(b, 16), nlen = ngs_inf_get(b, 16);
				const uint32_t at = ngs_inf_consumed(b);
				if ((len ^ nlen) != 0xffffu || at + len > in_len || pos + len > out_len) state = NGS_LS_FAIL;
				else {
					o.finish();
					__syncwarp(gmask);
					o.stored(pos, in + at, len);
					__syncwarp(gmask);
					pos += len;
					ngs_inf_open(b, in, at + len);
					state = last ? NGS_LS_END : NGS_LS_HDR;
				}
			} else if (btype == 3u) state = NGS_LS_FAIL;
			else state = o.tables(b, btype, T) ? NGS_LS_FAIL : NGS_LS_SYM;
		} else if (state == NGS_LS_OPEN) {
			const ngs_bgzf_desc d = desc[blk];
			in = comp + d.src; in_len = d.clen; out_len = d.ulen; o.o = out + d.dst; o.pend_len = 0; pos = 0; last = 0;
			ngs_inf_open(b, in, 0u);
			state = NGS_LS_HDR;
		} else if (state != NGS_LS_IDLE) {                       /* NGS_LS_END, NGS_LS_FAIL */
			o.finish();
			if ((state == NGS_LS_FAIL || pos != out_len) && o.lane == 0) atomicOr(&st->flags, NGS_FLAG_EFORMAT);
			blk += stride;
			state = blk < n_blk ? NGS_LS_OPEN : NGS_LS_IDLE;
		}
	}
}

/* small device -> pinned host copies done by a kernel (stores to host-mapped memory) instead of the copy engine: the
 * engine's device-to-host queue is busy for milliseconds at a time with the text of finished contigs, and a 200-byte
 * state snapshot queued behind it stalled the whole engine stream (measured: 4 ms per 256 MiB chunk) */
__global__ void k_to_host(uint32_t *dst, const uint32_t *src, unsigned int n_words)
{
	for (unsigned int i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = src[i];
	__threadfence_system();
}

/* reset of look-back status words and tickets before a scan / format launch */
__global__ void k_clear_u64(unsigned long long *p, unsigned long long n)
{
	const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = 0;
}
