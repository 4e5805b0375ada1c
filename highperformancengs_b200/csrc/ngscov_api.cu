/* csrc/ngscov_api.cu -- C ABI of libngscov_cuda (see include/ngscov.h for the contract and the
 * reference code each entry point replaces).  Host-side orchestration only: memory, streams, launch
 * order, the (rare) verify/repair loop and the per-stage CUDA-event timers.  All arithmetic on the
 * data runs in the kernels of ngscov_kernels.cuh.
 */
#include "../../include/ngscov.h"
#include "ngscov_cover.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <type_traits>
#include <algorithm>
#include <chrono>

#define NGS_HEADROOM (NGS_CARRY_MAX + 256)

struct ngscov {
	ngscov_cfg cfg;
	int n_ref;
	std::vector<uint32_t> ref_len;
	std::vector<std::string> ref_name;
	int device;
	cudaStream_t s_main, s_copy;
	cudaEvent_t ev_h2d[2], ev_verified, ev_finalized, ev_t0, ev_t1, ev_cov0, ev_cov1;
	char cuda_err[256];

	/* geometry */
	std::vector<ngs_seg> seg;
	std::vector<int32_t> ctg2seg;
	uint64_t n_tiles, n_elems, n_bins;
	std::vector<uint64_t> win_base;          /* [n_ref+1] */
	ngs_seg *d_seg; int32_t *d_ctg2seg; uint32_t *d_tile2seg; uint32_t *d_ref_len;
	unsigned long long *d_win_base;
	int *d_diff; bool diff_dirty;
	unsigned long long *d_bins;

	/* names */
	char *d_names; uint32_t *d_name_off; uint32_t max_name_tab;   /* "name\t" */
	unsigned long long *d_name8;
	char *d_wighdr; uint32_t *d_wighdr_off; uint32_t max_wighdr;

	/* staging + chunk buffers */
	size_t stage_cap;
	uint8_t *h_stage[2]; uint8_t *d_chunk[2]; uint8_t *d_carry;
	bool stage_busy[2];

	/* sub-chunk walk */
	int shift;
	long long *d_entry, *d_exit, *d_bmax; size_t sub_cap, bmax_cap;
	unsigned int *d_count; unsigned short *d_slots; size_t slots_cap;
	unsigned long long *d_ctr, *h_ctr;        /* striped K2 counters */
	bool fused;                               /* deltas are applied before the verification (single-pass variants) */
	int ingest_mode;                          /* 0 K1+K2 split, 1 fused walk, 2 tiles (fused decode), 3 tiles2 (staged K1+K2) */
	bool p_k2_done;
	ngs_dev_state *d_st, *h_st;

	/* pending (unverified) chunk */
	bool pending; const uint8_t *p_base; int64_t p_n; int64_t p_nsub; bool p_from_state; long long p_fixed0;
	bool last_pushed; int last_flags;
	int sticky_err;

	/* runs */
	uint32_t *d_run_start, *d_run_end; int *d_run_depth; size_t run_cap;
	unsigned long long *d_rec_base, *d_stA, *d_stB, *d_fmt_status; size_t fmt_status_cap;
	std::vector<unsigned long long> h_rec_base; unsigned long long *h_recb;   /* pinned landing area of the per-segment record offsets */
	bool runs_done, formatted;
	unsigned long long drop_idx, mod_idx; uint32_t mod_end;
	ngscov_halo halo;

	/* text */
	char *d_text[4]; size_t d_text_cap[4]; size_t text_len[4]; bool text_ready[4];
	char *h_text[4]; size_t h_text_cap[4]; bool h_text_valid[4];
	std::string chromsize;

	/* streaming per-contig finalize (single shard, sorted input): scan + format + D2H of a contig start as soon
	 * as a higher refID has been decoded, so the device->host copy of the text overlaps the host->device copy
	 * of the rest of the stream */
	bool stream_fin; size_t next_fin_seg; unsigned long long mops_fin; bool bins_zeroed;
	int32_t end_ctg; uint32_t end_pos;          /* first position behind the slice (end_ctg = -2: whole genome) */
	double fin_wall; int fin_calls;             /* host time spent in streamed finalizes (NGSCOV_DEBUG) */
	size_t bed_off;                          /* bedGraph bytes produced so far */
	size_t bed_d2h;                          /* ... of which already queued for D2H */
	ngs_dev_state *h_st2;                    /* pinned read-back of the state block after scan / format launches */
	ngs_dev_state *h_tmpl;                   /* pinned, constant: initial values of the scan fields */
	cudaStream_t s_out; cudaEvent_t ev_fmt;
	unsigned long long acc_boundaries, acc_runs; int acc_max, acc_min;

	/* cover pipeline (ngscov_cover.cuh): compact record table, rows appended chunk after chunk */
	bool cover;                              /* default; false: the classic RED + scan + format kernels (NGSCOV_PIPELINE=classic or any NGSCOV_INGEST) */
	bool red_mode;                           /* cover pipeline fell back to REDs + classic scan: the input is not coordinate sorted */
	uint2 *d_ent; ngs_row *d_rows; unsigned long long *d_rkmax; size_t rows_cap, rows_used;
	size_t p_row0;                           /* first global row of the pending chunk */
	unsigned long long *d_kbmax;
	uint32_t *d_far_flag, *d_row_a, *d_row_b;
	uint4 *d_heavy; unsigned int *d_heavy_n; size_t heavy_cap;   /* work items of the heavy tiles (k_heavy_plan / k_heavy_gather) */
	unsigned long long *d_stF, *d_stC;       /* look-back words of the owner tiles: bytes / (sum, last boundary) */
	uint4 *d_tile_first, *d_tile_last;
	uint64_t fin_tile;                       /* owner tiles already finalized (streamed finalize works at tile granularity) */
	bool halo_have; ngscov_run cov_first, cov_last;   /* first / last non-zero run over the cover launches so far */
	size_t bed_begin;                        /* halo: bytes of the dropped first line */
	std::string bed_tail; size_t bed_tail_old; /* halo: replacement of the last line, and how many bytes it replaces */
	bool bed_patched;
	/* NGSCOV_MODE_SLIDE (bam_sliding_count) */
	int reg_ref, reg_beg, reg_end;     /* ngscov_set_region (slide mode); reg_ref < 0: none */
	bool cov_pending; uint64_t cov_t0, cov_t1;   /* a cover launch whose counters have not been read back yet (cover_harvest) */
	bool slide; unsigned int *d_sl_cnt, *d_sl_len, *d_sl_ctg; unsigned long long *d_sl_gc, *d_sl_ncount; uint32_t *d_sl_nwin;
	std::string slide_text; unsigned long long slide_ncount;
	long long *d_halo;                       /* the shard's running halo record (device), see k_halo_gather */
	unsigned long long *d_patch, *h_patch;   /* device-side stitch: [0] first byte, [1] end of the shard's bedGraph text */
	bool dev_stitched;

	ngscov_stats_t stats;
	int sm_count;
	/* per-push timing events (start/end pairs), reused across inputs */
	std::vector<cudaEvent_t> ev_ing, ev_cp, ev_inf, ev_k2; size_t n_ing, n_cp, n_inf, n_k2;
	uint8_t *d_comp[2]; size_t comp_cap[2]; ngs_bgzf_desc *d_desc[2], *h_desc[2];   /* BGZF blocks awaiting the device inflate (h_desc: pinned) */
};

static cudaEvent_t next_event(std::vector<cudaEvent_t> &pool, size_t &used)
{
	if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
	return pool[used++];
}
static float sum_events(std::vector<cudaEvent_t> &pool, size_t &used)
{
	float tot = 0;
	for (size_t i = 0; i + 1 < used; i += 2) { float ms = 0; if (cudaEventElapsedTime(&ms, pool[i], pool[i + 1]) == cudaSuccess) tot += ms; }
	used = 0;
	return tot;
}

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
	snprintf(h->cuda_err, sizeof h->cuda_err, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
	return e__ == cudaErrorMemoryAllocation ? NGSCOV_ENOMEM : NGSCOV_ECUDA; } } while (0)

static int small_d2h(ngscov *h, void *pinned, const void *dev, size_t nbytes);
static void tic(ngscov *h) { cudaEventRecord(h->ev_t0, h->s_main); }
static float toc(ngscov *h)
{
	float ms = 0;
	cudaEventRecord(h->ev_t1, h->s_main);
	cudaEventSynchronize(h->ev_t1);
	cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1);
	return ms;
}

extern "C" int ngscov_abi_version(void) { return NGSCOV_ABI_VERSION; }

extern "C" const char *ngscov_strerror(int code)
{
	switch (code) {
	case NGSCOV_OK: return "ok";
	case NGSCOV_EINVAL: return "invalid argument";
	case NGSCOV_ECUDA: return "CUDA error";
	case NGSCOV_ENOMEM: return "out of device or pinned memory";
	case NGSCOV_EFORMAT: return "malformed or truncated BAM record stream";
	case NGSCOV_ERANGE: return "alignment reaches beyond contig end + overhang pad";
	case NGSCOV_ESTATE: return "call sequence violated";
	case NGSCOV_ETOOBIG: return "record header larger than the carry area";
	case NGSCOV_EUNSORTED: return "input is not coordinate sorted";
	default: return "unknown error";
	}
}

extern "C" const char *ngscov_last_cuda_error(const ngscov_t *h) { return h ? h->cuda_err : ""; }

template <class T>
static int dev_upload(ngscov *h, T **dst, const std::vector<T> &src)
{
	CK(cudaMalloc((void **)dst, std::max<size_t>(1, src.size()) * sizeof(T)));
	if (!src.empty()) CK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
	return 0;
}

static int build_geometry(ngscov *h)
{
	const ngscov_cfg &c = h->cfg;
	const uint32_t W = c.window;
	h->ctg2seg.assign(h->n_ref, -1);
	h->win_base.assign((size_t)h->n_ref + 1, 0);
	uint64_t g0 = 0, elems = 0, tiles = 0;
	for (int i = 0; i < h->n_ref; ++i) {
		const uint32_t L = h->ref_len[i];
		const uint64_t g1 = g0 + L;
		h->win_base[i + 1] = h->win_base[i] + (uint64_t)(L / W) + 1;
		bool in; uint32_t lo = 0, hi = L;
		if (c.slice_end == 0) in = true;
		else {
			in = (c.slice_beg < g1 && c.slice_end > g0) || (L == 0 && c.slice_beg <= g0 && g0 < c.slice_end);
			if (in) {
				lo = c.slice_beg > g0 ? (uint32_t)(c.slice_beg - g0) : 0;
				hi = c.slice_end < g1 ? (uint32_t)(c.slice_end - g0) : L;
			}
		}
		if (in) {
			ngs_seg s; memset(&s, 0, sizeof s);
			s.contig = (uint32_t)i; s.clip_lo = lo; s.ref_len = L; s.n_win = L / W + 1;
			s.hard_end = hi == L;
			if (s.hard_end) { uint64_t lim = (uint64_t)L + NGSCOV_OVERHANG_PAD; s.lim = (uint32_t)std::min<uint64_t>(lim, NGS_POS_MASK); }
			else s.lim = hi;
			if (s.lim < lo) s.lim = lo;
			const uint64_t n = (uint64_t)s.lim - lo + 1, nt = (n + NGS_TILE - 1) / NGS_TILE;
			s.base = elems; s.tile0 = tiles; s.bins0 = h->win_base[i];
			elems += nt * NGS_TILE; tiles += nt;
			h->ctg2seg[i] = (int32_t)h->seg.size();
			h->seg.push_back(s);
		}
		g0 = g1;
	}
	h->n_elems = elems; h->n_tiles = tiles; h->n_bins = h->win_base[h->n_ref];
	return 0;
}

static int create_impl(ngscov_t **out, const ngscov_cfg *cfg);
extern "C" int ngscov_create(ngscov_t **out, const ngscov_cfg *cfg)
{
	if (!out) return NGSCOV_EINVAL;
	*out = nullptr;
	const int r = create_impl(out, cfg);
	if (r && *out) {
		static char last_err[256];
		snprintf(last_err, sizeof last_err, "%s", (*out)->cuda_err);
		if (getenv("NGSCOV_DEBUG")) fprintf(stderr, "ngscov_create: %s: %s\n", ngscov_strerror(r), last_err);
		ngscov_destroy(*out); *out = nullptr;
	}
	return r;
}
static int create_impl(ngscov_t **out, const ngscov_cfg *cfg)
{
	if (!out || !cfg || cfg->n_ref < 0 || (cfg->n_ref && (!cfg->ref_len || !cfg->ref_name))) return NGSCOV_EINVAL;
	if (cfg->window < 1 || cfg->window > (1u << 30)) return NGSCOV_EINVAL;
	if (cfg->mode != NGSCOV_MODE_DEPTH && cfg->mode != NGSCOV_MODE_WIG && cfg->mode != NGSCOV_MODE_SLIDE) return NGSCOV_EINVAL;
	if (cfg->mode == NGSCOV_MODE_SLIDE && (cfg->slice_end || cfg->window > 0x7fffffffu)) return NGSCOV_EINVAL;
	if (cfg->slice_end && cfg->slice_end <= cfg->slice_beg) return NGSCOV_EINVAL;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev) return NGSCOV_ECUDA;
	ngscov *h = new ngscov();
	h->cfg = *cfg; h->n_ref = cfg->n_ref; h->device = cfg->device; h->cuda_err[0] = 0;
	h->ref_len.assign(cfg->ref_len, cfg->ref_len + cfg->n_ref);
	for (int i = 0; i < cfg->n_ref; ++i) h->ref_name.push_back(cfg->ref_name[i] ? cfg->ref_name[i] : "");
	h->cfg.ref_len = nullptr; h->cfg.ref_name = nullptr;
	*out = h;
#define CKC(call) do { int r__ = (call); if (r__) { return r__; } } while (0)
	CK(cudaSetDevice(h->device));
	cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, h->device));
	h->sm_count = prop.multiProcessorCount;
	CK(cudaStreamCreateWithFlags(&h->s_main, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
	CK(cudaEventCreateWithFlags(&h->ev_fmt, cudaEventDisableTiming));
	for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&h->ev_verified, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&h->ev_finalized, cudaEventDisableTiming));
	CK(cudaEventCreate(&h->ev_t0)); CK(cudaEventCreate(&h->ev_t1)); CK(cudaEventCreate(&h->ev_cov0)); CK(cudaEventCreate(&h->ev_cov1));
	CK(cudaFuncSetAttribute(k_ingest_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)NGS_TILE_THREADS << 10) + NGS_TILE_MARGIN)));
	CK(cudaFuncSetAttribute(k_ingest_tiles2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)NGS_T2_SUB << 10) + NGS_TILE_MARGIN + ((size_t)NGS_T2_SUB << 5) * 2)));
	build_geometry(h);
	CKC(dev_upload(h, &h->d_seg, h->seg));
	CKC(dev_upload(h, &h->d_ctg2seg, h->ctg2seg));
	CKC(dev_upload(h, &h->d_ref_len, h->ref_len));
	{
		std::vector<unsigned long long> wb(h->win_base.begin(), h->win_base.end());
		CKC(dev_upload(h, &h->d_win_base, wb));
		std::vector<uint32_t> t2s(h->n_tiles);
		for (size_t s = 0; s < h->seg.size(); ++s) {
			const uint64_t t1 = s + 1 < h->seg.size() ? h->seg[s + 1].tile0 : h->n_tiles;
			for (uint64_t t = h->seg[s].tile0; t < t1; ++t) t2s[t] = (uint32_t)s;
		}
		CKC(dev_upload(h, &h->d_tile2seg, t2s));
		/* name tables */
		std::vector<char> names, wh; std::vector<uint32_t> noff(1, 0), whoff(1, 0);
		h->max_name_tab = 0; h->max_wighdr = 0;
		for (int i = 0; i < h->n_ref; ++i) {
			const std::string &nm = h->ref_name[i];
			names.insert(names.end(), nm.begin(), nm.end()); names.push_back('\t');
			noff.push_back((uint32_t)names.size());
			h->max_name_tab = std::max<uint32_t>(h->max_name_tab, (uint32_t)nm.size() + 1);
			char tmp[64]; snprintf(tmp, sizeof tmp, " span=%d\n", (int)h->cfg.window);
			std::string hd = "variableStep chrom=" + nm + tmp;
			wh.insert(wh.end(), hd.begin(), hd.end());
			whoff.push_back((uint32_t)wh.size());
			h->max_wighdr = std::max<uint32_t>(h->max_wighdr, (uint32_t)hd.size());
			char cs[64]; snprintf(cs, sizeof cs, "\t%d\n", (int)h->ref_len[i]);
			h->chromsize += nm + cs;
		}
		CKC(dev_upload(h, &h->d_names, names)); CKC(dev_upload(h, &h->d_name_off, noff));
		{
			std::vector<unsigned long long> n8(h->n_ref, 0);
			for (int i = 0; i < h->n_ref; ++i)
				for (uint32_t k = 0; k < 8 && noff[i] + k < noff[i + 1]; ++k) n8[i] |= (unsigned long long)(unsigned char)names[noff[i] + k] << (8 * k);
			CKC(dev_upload(h, &h->d_name8, n8));
		}
		CKC(dev_upload(h, &h->d_wighdr, wh)); CKC(dev_upload(h, &h->d_wighdr_off, whoff));
	}
	CK(cudaMalloc((void **)&h->d_diff, std::max<uint64_t>(1, h->n_elems) * sizeof(int)));
	tic(h);
	CK(cudaMemsetAsync(h->d_diff, 0, h->n_elems * sizeof(int), h->s_main));
	h->stats.ms_zero += toc(h);
	CK(cudaMalloc((void **)&h->d_bins, std::max<uint64_t>(1, h->n_bins) * 8));
	CK(cudaMalloc((void **)&h->d_stA, std::max<uint64_t>(1, h->n_tiles) * 8));
	CK(cudaMalloc((void **)&h->d_stB, std::max<uint64_t>(1, h->n_tiles) * 8));
	CK(cudaMalloc((void **)&h->d_rec_base, (h->seg.size() + 2) * 8));
	CK(cudaMalloc((void **)&h->d_carry, NGS_CARRY_MAX));
	CK(cudaMalloc((void **)&h->d_st, sizeof(ngs_dev_state)));
	CK(cudaHostAlloc((void **)&h->h_st, sizeof(ngs_dev_state), cudaHostAllocDefault));
	CK(cudaHostAlloc((void **)&h->h_st2, sizeof(ngs_dev_state), cudaHostAllocDefault));
	CK(cudaHostAlloc((void **)&h->h_tmpl, sizeof(ngs_dev_state), cudaHostAllocDefault));
	memset(h->h_tmpl, 0, sizeof(ngs_dev_state)); h->h_tmpl->first_run = ~0ull;
	CK(cudaMalloc((void **)&h->d_ctr, NGS_CTR_STRIPES * 16 * 8));
	CK(cudaHostAlloc((void **)&h->h_ctr, NGS_CTR_STRIPES * 16 * 8, cudaHostAllocDefault));
	CK(cudaHostAlloc((void **)&h->h_recb, ((size_t)cfg->n_ref + 4) * 8, cudaHostAllocDefault));
	h->stage_cap = cfg->staging_bytes ? (size_t)cfg->staging_bytes : (size_t)256 << 20;
	/* Two ingest organisations are kept, both parity-tested: K1 record framing + K2 per-record decode (default:
	 * 22.7 ms on config 2) and the single fused walk+decode kernel (NGSCOV_FUSED=1: 27.1 ms), DESIGN.md section 4. */
	h->end_ctg = -2; h->end_pos = 0;
	if (cfg->slice_end) {
		uint64_t acc = 0;
		h->end_ctg = cfg->n_ref;                  /* slice reaches the end of the genome: only the unmapped tail lies behind */
		for (int i = 0; i < cfg->n_ref; ++i) {
			if (cfg->slice_end < acc + cfg->ref_len[i]) { h->end_ctg = i; h->end_pos = (uint32_t)(cfg->slice_end - acc); break; }
			acc += cfg->ref_len[i];
		}
	}
	h->ingest_mode = 0;
	if (getenv("NGSCOV_FUSED") && atoi(getenv("NGSCOV_FUSED")) == 1) h->ingest_mode = 1;
	if (h->ingest_mode == 1 && getenv("NGSCOV_TILES") && atoi(getenv("NGSCOV_TILES")) == 1) h->ingest_mode = 2;
	if (getenv("NGSCOV_INGEST")) {
		const char *m = getenv("NGSCOV_INGEST");
		h->ingest_mode = !strcmp(m, "fused") ? 1 : !strcmp(m, "tiles") ? 2 : !strcmp(m, "tiles2") ? 3 : 0;
	}
	h->fused = h->ingest_mode != 0;
	h->cover = !(getenv("NGSCOV_INGEST") || getenv("NGSCOV_FUSED") || (getenv("NGSCOV_PIPELINE") && !strcmp(getenv("NGSCOV_PIPELINE"), "classic")));
	/* streamed per-contig finalize: the cover pipeline also streams region shards (their halo only patches the first and
	 * last line afterwards); the classic kernels need the whole shard before the halo is known */
	h->stream_fin = (cfg->slice_end == 0 || h->cover) && !(getenv("NGSCOV_STREAM_FINALIZE") && atoi(getenv("NGSCOV_STREAM_FINALIZE")) == 0);
	h->slide = cfg->mode == NGSCOV_MODE_SLIDE;
	h->reg_ref = -1; h->reg_beg = 0; h->reg_end = 0;
	if (h->slide) {
		if (!h->cover) return NGSCOV_EINVAL;                              /* the table pipeline only */
		h->stream_fin = false;
		const size_t nb = std::max<uint64_t>(1, h->n_bins);
		CK(cudaMalloc((void **)&h->d_sl_cnt, nb * 4)); CK(cudaMalloc((void **)&h->d_sl_len, nb * 4)); CK(cudaMalloc((void **)&h->d_sl_gc, nb * 8));
		CK(cudaMalloc((void **)&h->d_sl_ctg, ((size_t)h->n_ref + 1) * 4)); CK(cudaMalloc((void **)&h->d_sl_ncount, 8));
		std::vector<uint32_t> nw((size_t)h->n_ref + 1, 0);
		for (int i = 0; i < h->n_ref; ++i) nw[i] = h->ref_len[i] / h->cfg.window + 1;
		CKC(dev_upload(h, &h->d_sl_nwin, nw));
	}
	if (h->cover) {
		const size_t nt = std::max<uint64_t>(1, h->n_tiles) * NGS_CPT;      /* owner tiles */
		CK(cudaMalloc((void **)&h->d_far_flag, nt * 4)); CK(cudaMemset(h->d_far_flag, 0, nt * 4));
		CK(cudaMalloc((void **)&h->d_stC, nt * 8));
		CK(cudaMalloc((void **)&h->d_row_a, nt * 4)); CK(cudaMalloc((void **)&h->d_row_b, nt * 4));
		CK(cudaMalloc((void **)&h->d_stF, nt * 8));
		CK(cudaMalloc((void **)&h->d_tile_first, nt * 16)); CK(cudaMalloc((void **)&h->d_tile_last, nt * 16));
		h->h_tmpl->first_tile = ~0ull;
		CK(cudaMalloc((void **)&h->d_halo, 9 * 8));
		CK(cudaMalloc((void **)&h->d_patch, 2 * 8));
		CK(cudaHostAlloc((void **)&h->h_patch, 2 * 8, cudaHostAllocDefault));
	}
	h->shift = 0;
	const char *env = getenv("NGSCOV_SUBCHUNK");
	if (env && atoi(env) >= 64) { int v = atoi(env), s = 6; while ((1 << (s + 1)) <= v) ++s; h->shift = s > 15 ? 15 : s; }   /* 16-bit record slots */
	return ngscov_reset(h);
}

extern "C" void ngscov_destroy(ngscov_t *h)
{
	if (!h) return;
	cudaSetDevice(h->device);
	cudaDeviceSynchronize();
	void *dev[] = { h->d_seg, h->d_ctg2seg, h->d_tile2seg, h->d_ref_len, h->d_win_base, h->d_diff, h->d_bins, h->d_names, h->d_name8,
	                h->d_name_off, h->d_wighdr, h->d_wighdr_off, h->d_chunk[0], h->d_chunk[1], h->d_carry, h->d_entry, h->d_exit,
	                h->d_bmax, h->d_count, h->d_slots, h->d_ctr, h->d_st, h->d_run_start, h->d_run_end, h->d_run_depth, h->d_rec_base, h->d_stA, h->d_stB,
	                h->d_fmt_status, h->d_text[0], h->d_text[1], h->d_text[2], h->d_text[3],
	                h->d_ent, h->d_rows, h->d_rkmax, h->d_kbmax, h->d_far_flag, h->d_row_a, h->d_row_b, h->d_stF, h->d_stC, h->d_tile_first, h->d_tile_last, h->d_halo, h->d_patch,
	                h->d_sl_cnt, h->d_sl_len, h->d_sl_gc, h->d_sl_ctg, h->d_sl_ncount, h->d_sl_nwin, h->d_heavy, h->d_heavy_n };
	for (void *p : dev) if (p) cudaFree(p);
	for (int i = 0; i < 2; ++i) if (h->h_stage[i]) cudaFreeHost(h->h_stage[i]);
	for (int i = 0; i < 4; ++i) if (h->h_text[i]) cudaFreeHost(h->h_text[i]);
	if (h->h_st) cudaFreeHost(h->h_st);
	if (h->h_ctr) cudaFreeHost(h->h_ctr);
	if (h->h_recb) cudaFreeHost(h->h_recb);
	if (h->h_st2) cudaFreeHost(h->h_st2);
	if (h->h_tmpl) cudaFreeHost(h->h_tmpl);
	if (h->h_patch) cudaFreeHost(h->h_patch);
	if (h->s_out) cudaStreamDestroy(h->s_out);
	if (h->ev_fmt) cudaEventDestroy(h->ev_fmt);
	if (h->s_main) cudaStreamDestroy(h->s_main);
	if (h->s_copy) cudaStreamDestroy(h->s_copy);
	for (int i = 0; i < 2; ++i) if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
	if (h->ev_verified) cudaEventDestroy(h->ev_verified);
	if (h->ev_finalized) cudaEventDestroy(h->ev_finalized);
	if (h->ev_t0) cudaEventDestroy(h->ev_t0);
	if (h->ev_t1) cudaEventDestroy(h->ev_t1);
	if (h->ev_cov0) cudaEventDestroy(h->ev_cov0);
	if (h->ev_cov1) cudaEventDestroy(h->ev_cov1);
	for (cudaEvent_t e : h->ev_ing) cudaEventDestroy(e);
	for (cudaEvent_t e : h->ev_inf) cudaEventDestroy(e);
	for (cudaEvent_t e : h->ev_k2) cudaEventDestroy(e);
	for (int i = 0; i < 2; ++i) { if (h->d_comp[i]) cudaFree(h->d_comp[i]); if (h->d_desc[i]) cudaFree(h->d_desc[i]); if (h->h_desc[i]) cudaFreeHost(h->h_desc[i]); }
	for (cudaEvent_t e : h->ev_cp) cudaEventDestroy(e);
	delete h;
}

extern "C" int ngscov_reset(ngscov_t *h)
{
	if (!h) return NGSCOV_EINVAL;
	CK(cudaSetDevice(h->device));
	CK(cudaStreamSynchronize(h->s_main)); CK(cudaStreamSynchronize(h->s_copy));
	h->cov_pending = false;
	if (h->diff_dirty) {            /* pushes were made but the scan (which re-zeroes) never ran */
		tic(h);
		if (h->d_far_flag) CK(cudaMemsetAsync(h->d_far_flag, 0, std::max<uint64_t>(1, h->n_tiles) * NGS_CPT * 4, h->s_main));
		CK(cudaMemsetAsync(h->d_diff, 0, h->n_elems * sizeof(int), h->s_main));
		h->stats.ms_zero = toc(h);
		h->diff_dirty = false;
	}
	const float keep_zero = h->stats.ms_zero;
	memset(&h->stats, 0, sizeof h->stats);
	h->stats.ms_zero = keep_zero;
	CK(cudaMemsetAsync(h->d_st, 0, sizeof(ngs_dev_state), h->s_main));
	CK(cudaMemsetAsync(h->d_ctr, 0, NGS_CTR_STRIPES * 16 * 8, h->s_main));
	CK(cudaStreamSynchronize(h->s_main));
	h->pending = false; h->last_pushed = false; h->last_flags = 0; h->sticky_err = 0;
	h->runs_done = false; h->formatted = false;
	h->drop_idx = ~0ull; h->mod_idx = ~0ull; h->mod_end = 0;
	memset(&h->halo, 0, sizeof h->halo); h->halo.first.contig = h->halo.last.contig = -1;
	for (int i = 0; i < 4; ++i) { h->text_len[i] = 0; h->text_ready[i] = false; h->h_text_valid[i] = false; }
	h->stage_busy[0] = h->stage_busy[1] = false;
	h->n_ing = 0; h->n_cp = 0; h->n_inf = 0; h->n_k2 = 0;
	CK(cudaStreamSynchronize(h->s_out));
	h->next_fin_seg = 0; h->mops_fin = 0; h->bins_zeroed = false; h->bed_off = 0; h->bed_d2h = 0;
	h->acc_boundaries = h->acc_runs = 0; h->acc_max = h->acc_min = 0;
	h->rows_used = 0; h->red_mode = false; h->halo_have = false; h->fin_tile = 0;
	h->bed_begin = 0; h->bed_tail.clear(); h->bed_tail_old = 0; h->bed_patched = false; h->dev_stitched = false;
	if (h->d_halo) { CK(cudaMemsetAsync(h->d_halo, 0, 9 * 8, h->s_main)); CK(cudaStreamSynchronize(h->s_main)); }
	if (h->slide) {
		const size_t nb = std::max<uint64_t>(1, h->n_bins);
		CK(cudaMemsetAsync(h->d_sl_cnt, 0, nb * 4, h->s_main)); CK(cudaMemsetAsync(h->d_sl_len, 0, nb * 4, h->s_main)); CK(cudaMemsetAsync(h->d_sl_gc, 0, nb * 8, h->s_main));
		CK(cudaMemsetAsync(h->d_sl_ctg, 0, ((size_t)h->n_ref + 1) * 4, h->s_main)); CK(cudaMemsetAsync(h->d_sl_ncount, 0, 8, h->s_main));
		CK(cudaStreamSynchronize(h->s_main));
		h->slide_text.clear(); h->slide_ncount = 0;
	}
	return 0;
}

extern "C" int ngscov_sync(ngscov_t *h)
{
	if (!h) return NGSCOV_EINVAL;
	CK(cudaSetDevice(h->device));
	CK(cudaStreamSynchronize(h->s_copy)); CK(cudaStreamSynchronize(h->s_main)); CK(cudaStreamSynchronize(h->s_out));
	return 0;
}

extern "C" void *ngscov_stream(ngscov_t *h) { return h ? (void *)h->s_main : nullptr; }

extern "C" int ngscov_staging(ngscov_t *h, int idx, void **pinned, size_t *cap)
{
	if (!h || idx < 0 || idx > 1 || !pinned || !cap) return NGSCOV_EINVAL;
	CK(cudaSetDevice(h->device));
	if (!h->h_stage[idx]) {
		CK(cudaHostAlloc((void **)&h->h_stage[idx], h->stage_cap, cudaHostAllocDefault));
		if (!h->d_chunk[idx]) CK(cudaMalloc((void **)&h->d_chunk[idx], NGS_HEADROOM + h->stage_cap + 256));
	}
	*pinned = h->h_stage[idx]; *cap = h->stage_cap;
	return 0;
}

extern "C" int ngscov_staging_wait(ngscov_t *h, int idx)
{
	if (!h || idx < 0 || idx > 1) return NGSCOV_EINVAL;
	if (h->stage_busy[idx]) { CK(cudaEventSynchronize(h->ev_h2d[idx])); h->stage_busy[idx] = false; }
	return 0;
}

static int ensure_sub(ngscov *h, int64_t n_sub)
{
	if ((size_t)n_sub > h->sub_cap) {
		CK(cudaStreamSynchronize(h->s_main));
		if (h->d_entry) cudaFree(h->d_entry);
		if (h->d_exit) cudaFree(h->d_exit);
		h->d_entry = h->d_exit = nullptr; h->sub_cap = 0;
		const size_t cap = (size_t)n_sub + (size_t)n_sub / 8 + 1024;
		CK(cudaMalloc((void **)&h->d_entry, cap * 8));
		CK(cudaMalloc((void **)&h->d_exit, cap * 8));
		if (h->d_count) cudaFree(h->d_count);
		h->d_count = nullptr;
		CK(cudaMalloc((void **)&h->d_count, cap * 4));
		h->sub_cap = cap;
	}
	if (h->ingest_mode == 0 && !h->cover) {
		const size_t need = ((size_t)n_sub << (h->shift - 5)) + 64;
		if (need > h->slots_cap) {
			CK(cudaStreamSynchronize(h->s_main));
			if (h->d_slots) cudaFree(h->d_slots);
			h->d_slots = nullptr; h->slots_cap = 0;
			CK(cudaMalloc((void **)&h->d_slots, (need + need / 8) * 2));
			h->slots_cap = need + need / 8;
		}
	}
	const size_t nb = ((size_t)n_sub + NGS_VER_PER_BLOCK - 1) / NGS_VER_PER_BLOCK;
	if (nb > h->bmax_cap) {
		CK(cudaStreamSynchronize(h->s_main));
		if (h->d_bmax) cudaFree(h->d_bmax);
		if (h->d_kbmax) cudaFree(h->d_kbmax);
		h->d_bmax = nullptr; h->d_kbmax = nullptr; h->bmax_cap = 0;
		CK(cudaMalloc((void **)&h->d_bmax, (nb + 1024) * 8));
		CK(cudaMalloc((void **)&h->d_kbmax, (nb + 1024) * 8));
		h->bmax_cap = nb + 1024;
	}
	if (h->cover && h->rows_used + (size_t)n_sub > h->rows_cap) {
		/* the table grows by doubling; what is there is copied device to device (a few GB at most, ~1 ms per GB) */
		CK(cudaStreamSynchronize(h->s_main));
		const size_t need = h->rows_used + (size_t)n_sub;
		const size_t cap = std::max(need + need / 8 + 64, h->rows_cap * 2);
		uint2 *ne = nullptr; ngs_row *nr = nullptr; unsigned long long *nk = nullptr;
		CK(cudaMalloc((void **)&ne, cap * NGS_ROW_SLOTS * sizeof(uint2)));
		CK(cudaMalloc((void **)&nr, cap * sizeof(ngs_row)));
		CK(cudaMalloc((void **)&nk, cap * 8));
		if (h->rows_used) {
			CK(cudaMemcpyAsync(ne, h->d_ent, h->rows_used * NGS_ROW_SLOTS * sizeof(uint2), cudaMemcpyDeviceToDevice, h->s_main));
			CK(cudaMemcpyAsync(nr, h->d_rows, h->rows_used * sizeof(ngs_row), cudaMemcpyDeviceToDevice, h->s_main));
			CK(cudaMemcpyAsync(nk, h->d_rkmax, h->rows_used * 8, cudaMemcpyDeviceToDevice, h->s_main));
			CK(cudaStreamSynchronize(h->s_main));
		}
		if (h->d_ent) cudaFree(h->d_ent);
		if (h->d_rows) cudaFree(h->d_rows);
		if (h->d_rkmax) cudaFree(h->d_rkmax);
		h->d_ent = ne; h->d_rows = nr; h->d_rkmax = nk; h->rows_cap = cap;
	}
	return 0;
}

static IngestArgs make_ingest_args(ngscov *h, const uint8_t *base, int64_t n, int64_t n_sub, bool from_state, long long fixed0)
{
	IngestArgs A;
	A.base = base; A.n = n; A.shift = h->shift; A.n_sub = n_sub; A.entry = h->d_entry; A.exit = h->d_exit;
	A.cfg.n_ref = h->n_ref; A.cfg.flag_mask = h->cfg.mode == NGSCOV_MODE_DEPTH ? 0x704u : 0x4u;
	A.cfg.ctg2seg = h->d_ctg2seg; A.cfg.seg = h->d_seg; A.cfg.end_ctg = h->end_ctg; A.cfg.end_pos = h->end_pos;
	A.cfg.reg_ref = h->reg_ref; A.cfg.reg_beg = h->reg_beg; A.cfg.reg_end = h->reg_end;
	A.ref_len = h->d_ref_len; A.diff = h->d_diff; A.st = h->d_st;
	A.count = h->d_count; A.slots = h->d_slots; A.slot_shift = h->shift - 5; A.ctr = h->d_ctr;
	A.fixed_entry0 = fixed0; A.entry0_from_state = from_state ? 1 : 0;
	A.debug = getenv("NGSCOV_DEBUG_INGEST") ? atoi(getenv("NGSCOV_DEBUG_INGEST")) : 0;
	A.rows = h->d_rows ? h->d_rows + h->p_row0 : nullptr; A.ent = h->d_ent ? h->d_ent + h->p_row0 * NGS_ROW_SLOTS : nullptr;
	A.rkmax = h->d_rkmax ? h->d_rkmax + h->p_row0 : nullptr; A.far_flag = h->d_far_flag;
	return A;
}

static int small_d2h(ngscov *h, void *pinned, const void *dev, size_t nbytes)
{
	k_to_host<<<1, 256, 0, h->s_main>>>((uint32_t *)pinned, (const uint32_t *)dev, (unsigned int)(nbytes / 4));
	h->stats.launches += 1;
	CK(cudaGetLastError());
	return 0;
}

static int launch_verify(ngscov *h)
{
	const int64_t n_sub = h->p_nsub;
	const unsigned nb = (unsigned)((n_sub + NGS_VER_PER_BLOCK - 1) / NGS_VER_PER_BLOCK);
	if (h->cover) {
		k_verify_blockmax_k<<<nb, 256, 0, h->s_main>>>(h->d_exit, h->d_rows + h->p_row0, n_sub, h->d_bmax, h->d_kbmax);
		k_verify_scan_k<<<1, 1024, 0, h->s_main>>>(h->d_bmax, h->d_kbmax, nb, h->d_st, h->p_fixed0, h->p_from_state ? 1 : 0);
		k_verify_check_k<<<nb, 256, 0, h->s_main>>>(h->d_entry, h->d_exit, h->d_rows + h->p_row0, n_sub, h->d_bmax, h->d_kbmax, h->d_rkmax + h->p_row0,
		                                           h->shift, h->p_n, h->d_st);
	} else {
		k_verify_blockmax<<<nb, 256, 0, h->s_main>>>(h->d_exit, n_sub, h->d_bmax);
		k_verify_scan<<<1, 1024, 0, h->s_main>>>(h->d_bmax, nb, h->d_st, h->p_fixed0, h->p_from_state ? 1 : 0);
		k_verify_check<<<nb, 256, 0, h->s_main>>>(h->d_entry, h->d_exit, n_sub, h->d_bmax, h->shift, h->p_n, h->d_st);
	}
	h->stats.launches += 3;
	{ int r_ = small_d2h(h, h->h_st, h->d_st, sizeof(ngs_dev_state)); if (r_) return r_; }
	/* the striped K2 counters travel with the same snapshot: both describe the chunks BEFORE this one (this chunk's K2
	 * runs behind the verification), which is all the streamed finalize needs -- and it spares a second host
	 * synchronisation per chunk */
	{ int r_ = small_d2h(h, h->h_ctr, h->d_ctr, NGS_CTR_STRIPES * 16 * 8); if (r_) return r_; }
	CK(cudaEventRecord(h->ev_verified, h->s_main));
	CK(cudaGetLastError());
	return 0;
}

static int launch_special(ngscov *h, const IngestArgs &A, int mode)
{
	/* a warp takes 32 rows per step */
	const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((A.n_sub + 255) / 256, (int64_t)h->sm_count * 8));
	if (h->slide) {
		SlideArgs S;
		S.cnt = h->d_sl_cnt; S.len = h->d_sl_len; S.gc = h->d_sl_gc; S.ctg_reads = h->d_sl_ctg; S.win_base = h->d_win_base; S.n_win = h->d_sl_nwin;
		S.window = (int)h->cfg.window; S.n_count = h->d_sl_ncount;
		k_slide_accumulate<<<grid, 256, 0, h->s_main>>>(A, S);
		h->stats.launches += 1;
		CK(cudaGetLastError());
		return 0;
	}
	k_special_reds<<<grid, 256, 0, h->s_main>>>(A, mode);
	h->stats.launches += 1;
	CK(cudaGetLastError());
	return 0;
}

/* wait for the verification of the chunk in flight; repair mis-speculated sub-chunks; fix the carry */
static int stream_finalize(ngscov *h, size_t seg_end);
static int stream_finalize_cover(ngscov *h);

static int resolve_pending(ngscov *h, bool allow_stream = false)
{
	if (!h->pending) return 0;
	for (int iter = 0; ; ++iter) {
		CK(cudaEventSynchronize(h->ev_verified));
		if (h->h_st->n_bad == 0) break;
		const unsigned long long key = h->h_st->first_bad_key;
		if ((key & 1) || iter > 100000) { h->sticky_err = NGSCOV_EFORMAT; h->pending = false; return NGSCOV_EFORMAT; }
		IngestArgs A = make_ingest_args(h, h->p_base, h->p_n, h->p_nsub, h->p_from_state, h->p_fixed0);
		if (h->cover && h->slide) k_repair_classify<true><<<1, 32, 0, h->s_main>>>(A, (long long)(key >> 1), h->d_bmax);
		else if (h->cover) k_repair_classify<false><<<1, 32, 0, h->s_main>>>(A, (long long)(key >> 1), h->d_bmax);
		else if (h->fused) k_repair<<<1, 32, 0, h->s_main>>>(A, (long long)(key >> 1), h->d_bmax);
		else k_repair_chain<<<1, 32, 0, h->s_main>>>(A, (long long)(key >> 1), h->d_bmax);
		h->stats.launches += 1;
		h->p_k2_done = false;                    /* the optimistic K2 saw n_bad != 0 and did nothing */
		int r = launch_verify(h);
		if (r) return r;
	}
	if (!h->fused && !h->p_k2_done) {
		IngestArgs A = make_ingest_args(h, h->p_base, h->p_n, h->p_nsub, h->p_from_state, h->p_fixed0);
		if (h->cover) { int r_ = launch_special(h, A, h->red_mode ? 1 : 0); if (r_) return r_; }
		else {
			k_decode_scatter<<<(unsigned)((((uint64_t)h->p_nsub << A.slot_shift) + 255) / 256), 256, 0, h->s_main>>>(A);
			h->stats.launches += 1;
		}
		h->p_k2_done = true;
	}
	if (h->cover && h->h_st->unsorted && !h->red_mode && !h->slide) {
		/* the input is not coordinate sorted: the owner tiles cannot find their records by key.  Everything pushed so far
		 * already had its rare deltas added to the global array; add the ordinary ones too (needs only the table) and go
		 * on with REDs + the classic scan.  Contigs that were already finalized cannot be taken back. */
		if (h->next_fin_seg > 0 || h->fin_tile > 0) { h->sticky_err = NGSCOV_EUNSORTED; h->pending = false; return NGSCOV_EUNSORTED; }
		h->red_mode = true; h->stream_fin = false;
		IngestArgs A = make_ingest_args(h, h->p_base, h->p_n, h->p_nsub, h->p_from_state, h->p_fixed0);
		A.rows = h->d_rows; A.ent = h->d_ent; A.rkmax = h->d_rkmax; A.n_sub = (int64_t)h->rows_used;
		int r_ = launch_special(h, A, 2);
		if (r_) return r_;
	}
	k_finalize_chunk<<<1, 256, 0, h->s_main>>>(h->p_base, h->p_n, h->d_st, h->d_carry);
	h->stats.launches += 1;
	CK(cudaEventRecord(h->ev_finalized, h->s_main));
	CK(cudaGetLastError());
	h->pending = false;
	/* sorted input: every contig below the highest refID decoded so far is complete */
	if (allow_stream && h->stream_fin && (h->ingest_mode == 0 || h->ingest_mode == 3)) {
		/* these variants count in striped counters: fold the stripes of the verification-time snapshot (K2 of the
		 * chunks before this one; the contigs it reports complete were decoded entirely by those) */
		for (int i = 0; i < NGS_CTR_STRIPES; ++i) h->h_st->m_ops += h->h_ctr[i * 16 + 2];
	}
	if (allow_stream && h->stream_fin && h->cover && !h->red_mode) {
		if (h->h_st->flags & (NGS_FLAG_EFORMAT | NGS_FLAG_ERANGE)) return 0;
		return stream_finalize_cover(h);
	}
	if (allow_stream && h->stream_fin && !(h->h_st->flags & (NGS_FLAG_EFORMAT | NGS_FLAG_ERANGE)) && h->h_st->max_ref_p1 > 1) {
		/* complete contigs -> complete segments (a region shard only holds the contigs that intersect its slice) */
		const uint32_t done_ctg = h->h_st->max_ref_p1 - 1;
		size_t upto = h->next_fin_seg;
		while (upto < h->seg.size() && h->seg[upto].contig < done_ctg) ++upto;
		if (upto > h->next_fin_seg) return stream_finalize(h, upto);
	}
	return 0;
}

static int pick_shift(ngscov *h, int32_t first_block_size)
{
	if (h->shift) return 0;
	int64_t rec = (int64_t)first_block_size + 4;
	if (rec < 36) rec = 36;
	if (rec > 4096) rec = 4096;
	int64_t want = rec * 128; int s = 9;                      /* ~128-256 records per sub-chunk, 512 B .. 32 KiB */
	int smax = 15;
	if (h->ingest_mode >= 2) { want = rec * 16; smax = 10; }   /* staged variants: a tile of 64 sub-chunks must fit shared memory */
	while (((int64_t)1 << s) < want && s < smax) ++s;
	h->shift = s;
	return 0;
}

static int submit_chunk(ngscov *h, const uint8_t *base, int64_t n, bool from_state, long long fixed0)
{
	const int64_t n_sub = (n + ((int64_t)1 << h->shift) - 1) >> h->shift;
	int r = ensure_sub(h, n_sub);
	if (r) return r;
	h->p_row0 = h->rows_used;
	IngestArgs A = make_ingest_args(h, base, n, n_sub, from_state, fixed0);
	CK(cudaEventRecord(next_event(h->ev_ing, h->n_ing), h->s_main));
	if (h->cover) {
		/* K1: frame AND classify (speculative chain walk, one 8-byte table entry per contributing record) */
		if (h->slide) k_chain_classify<true><<<(unsigned)((n_sub + 255) / 256), 256, 0, h->s_main>>>(A);
		else k_chain_classify<false><<<(unsigned)((n_sub + 255) / 256), 256, 0, h->s_main>>>(A);
		h->rows_used += (size_t)n_sub;
	} else if (!h->fused) {
		/* K1: frame the records (speculative chain walk, no atomics, full occupancy) */
		const unsigned nb = (unsigned)((n_sub + 255) / 256);
		k_chain_walk<<<nb, 256, 0, h->s_main>>>(A);
	} else {
		/* single-kernel variants (NGSCOV_FUSED=1; NGSCOV_TILES=1 stages tiles in shared memory): kept for A/B
		 * measurements, see DESIGN.md section 4 */
		const bool staged_ok = h->shift <= 10 && (((uintptr_t)base) & 15) == 0;
		if (h->ingest_mode == 3 && staged_ok) {
			const size_t smem = ((size_t)NGS_T2_SUB << h->shift) + NGS_TILE_MARGIN + ((size_t)NGS_T2_SUB << A.slot_shift) * 2;
			const unsigned nb = (unsigned)((n_sub + NGS_T2_SUB - 1) / NGS_T2_SUB);
			k_ingest_tiles2<<<nb, NGS_T2_THREADS, smem, h->s_main>>>(A);
		} else if (h->ingest_mode == 2 && staged_ok) {
			const size_t smem = ((size_t)NGS_TILE_THREADS << h->shift) + NGS_TILE_MARGIN;
			const unsigned nb = (unsigned)((n_sub + NGS_TILE_THREADS - 1) / NGS_TILE_THREADS);
			k_ingest_tiles<<<nb, NGS_TILE_THREADS, smem, h->s_main>>>(A);
		} else {
			const unsigned nb = (unsigned)((n_sub + 255) / 256);
			k_ingest_walk<<<nb, 256, 0, h->s_main>>>(A);
		}
	}
	h->stats.launches += 1;
	CK(cudaGetLastError());
	h->pending = true; h->p_base = base; h->p_n = n; h->p_nsub = n_sub; h->p_from_state = from_state; h->p_fixed0 = fixed0;
	h->diff_dirty = true;
	r = launch_verify(h);
	if (!r && !h->fused) {
		/* K2 queued optimistically behind the verification: it returns at once if a mismatch was found,
		 * and is then launched again by resolve_pending after the repair */
		CK(cudaEventRecord(next_event(h->ev_k2, h->n_k2), h->s_main));
		if (h->cover) { if ((r = launch_special(h, A, h->red_mode ? 1 : 0))) return r; }
		else {
			k_decode_scatter<<<(unsigned)((((uint64_t)n_sub << A.slot_shift) + 255) / 256), 256, 0, h->s_main>>>(A);
			h->stats.launches += 1;
		}
		CK(cudaEventRecord(next_event(h->ev_k2, h->n_k2), h->s_main));
		h->p_k2_done = true;
		CK(cudaGetLastError());
	}
	CK(cudaEventRecord(next_event(h->ev_ing, h->n_ing), h->s_main));
	return r;
}

extern "C" int ngscov_push(ngscov_t *h, int idx, size_t nbytes, int flags)
{
	if (!h || idx < 0 || idx > 1 || !h->h_stage[idx] || nbytes > h->stage_cap) return NGSCOV_EINVAL;
	if (h->runs_done || h->last_pushed) return NGSCOV_ESTATE;
	if (h->sticky_err) return h->sticky_err;
	CK(cudaSetDevice(h->device));
	int r = resolve_pending(h, true);
	if (r) return r;
	if (flags & NGSCOV_PUSH_LAST) { h->last_pushed = true; h->last_flags = flags; }
	h->stats.stream_bytes += nbytes;
	if (nbytes == 0) return 0;
	if (!h->shift) pick_shift(h, *(const int32_t *)h->h_stage[idx]);
	uint8_t *data = h->d_chunk[idx] + NGS_HEADROOM;
	CK(cudaStreamWaitEvent(h->s_copy, h->ev_finalized, 0));
	CK(cudaEventRecord(next_event(h->ev_cp, h->n_cp), h->s_copy));
	CK(cudaMemcpyAsync(data, h->h_stage[idx], nbytes, cudaMemcpyHostToDevice, h->s_copy));
	CK(cudaEventRecord(next_event(h->ev_cp, h->n_cp), h->s_copy));
	CK(cudaEventRecord(h->ev_h2d[idx], h->s_copy));
	h->stage_busy[idx] = true;
	CK(cudaStreamWaitEvent(h->s_main, h->ev_h2d[idx], 0));
	k_place_carry<<<1, 256, 0, h->s_main>>>(data, h->d_st, h->d_carry);
	h->stats.launches += 1;
	return submit_chunk(h, data, (int64_t)nbytes, true, 0);
}

static int ensure_chunk(ngscov *h, int idx)
{
	if (!h->d_chunk[idx]) CK(cudaMalloc((void **)&h->d_chunk[idx], NGS_HEADROOM + h->stage_cap + 256));
	return 0;
}

static int issue_h2d(ngscov *h, int idx, const uint8_t *src, size_t m)
{
	CK(cudaStreamWaitEvent(h->s_copy, h->ev_finalized, 0));
	CK(cudaEventRecord(next_event(h->ev_cp, h->n_cp), h->s_copy));
	CK(cudaMemcpyAsync(h->d_chunk[idx] + NGS_HEADROOM, src, m, cudaMemcpyHostToDevice, h->s_copy));
	CK(cudaEventRecord(next_event(h->ev_cp, h->n_cp), h->s_copy));
	CK(cudaEventRecord(h->ev_h2d[idx], h->s_copy));
	return 0;
}

/* Stream a caller-owned host buffer (ideally pinned) through the two device chunk buffers: the copy of
 * chunk k+1 overlaps the ingest of chunk k.  Records may straddle chunks and calls. */
extern "C" int ngscov_push_host(ngscov_t *h, const void *host, size_t nbytes, int flags)
{
	if (!h || (!host && nbytes)) return NGSCOV_EINVAL;
	if (h->runs_done || h->last_pushed) return NGSCOV_ESTATE;
	if (h->sticky_err) return h->sticky_err;
	CK(cudaSetDevice(h->device));
	int r;
	if ((r = ensure_chunk(h, 0)) || (r = ensure_chunk(h, 1))) return r;
	if (flags & NGSCOV_PUSH_LAST) { h->last_pushed = true; h->last_flags = flags; }
	h->stats.stream_bytes += nbytes;
	if (nbytes == 0) return resolve_pending(h);
	const uint8_t *src = (const uint8_t *)host;
	if (!h->shift) pick_shift(h, *(const int32_t *)src);
	if ((r = resolve_pending(h))) return r;
	const size_t cap = h->stage_cap;
	const size_t nchunks = (nbytes + cap - 1) / cap;
	int idx = 0;
	if ((r = issue_h2d(h, idx, src, std::min(cap, nbytes)))) return r;
	for (size_t k = 0; k < nchunks; ++k, idx ^= 1) {
		const size_t off = k * cap, m = std::min(cap, nbytes - off);
		if ((r = resolve_pending(h, true))) return r;                /* chunk k-1 verified, its buffer is free */
		if (k + 1 < nchunks && (r = issue_h2d(h, idx ^ 1, src + off + m, std::min(cap, nbytes - off - m)))) return r;
		uint8_t *data = h->d_chunk[idx] + NGS_HEADROOM;
		CK(cudaStreamWaitEvent(h->s_main, h->ev_h2d[idx], 0));
		k_place_carry<<<1, 256, 0, h->s_main>>>(data, h->d_st, h->d_carry);
		h->stats.launches += 1;
		if ((r = submit_chunk(h, data, (int64_t)m, true, 0))) return r;
	}
	return 0;
}

/* the device inflate.  Default: one decoder per warp (k_bgzf_inflate).  NGSCOV_INF_S=2 runs two decoders per warp as a
 * warp-wide state machine (k_bgzf_inflate_ls): 13 % fewer warp instructions, but half the warps per block of input --
 * +12 % when a launch covers several waves of blocks, -30 % at the default chunk size.  The
 * grid fills every SM to its occupancy and the decoders stride over the blocks */
template <class K>
static void launch_inflate_k(ngscov_t *h, K kern, int S, int &per_sm, int idx, uint32_t nb, uint8_t *data, int sm)
{
	if (!per_sm) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NGS_INF_WARPS * 32, 0); if (per_sm < 1) per_sm = 1; }
	const size_t per_cta = (size_t)NGS_INF_WARPS * S;
	const unsigned grid = (unsigned)std::min<size_t>((nb + per_cta - 1) / per_cta, (size_t)sm * per_sm);
	kern<<<grid, NGS_INF_WARPS * 32, 0, h->s_main>>>(h->d_comp[idx], h->d_desc[idx], nb, data, h->d_st);
}
static void launch_inflate(ngscov_t *h, int idx, uint32_t nb, uint8_t *data, int sm)
{
	static int occ[2] = { 0, 0 };
	const char *e = getenv("NGSCOV_INF_S");
	if (e && atoi(e) == 2) launch_inflate_k(h, k_bgzf_inflate_ls<2>, 2, occ[1], idx, nb, data, sm);
	else launch_inflate_k(h, k_bgzf_inflate<1>, 1, occ[0], idx, nb, data, sm);
}

/* BGZF blocks as they lie in the file -> device inflate -> the same chunk pipeline as ngscov_push_host.  The host only
 * walks the 18-byte block headers (BSIZE) and 8-byte footers (ISIZE); samtools bgzf.c:214-238 (inflate_block) runs on
 * the GPU (k_bgzf_inflate).  Compressed bytes of chunk k+1 are copied while chunk k is inflated and ingested. */
extern "C" int ngscov_push_bgzf(ngscov_t *h, const void *blocks, size_t nbytes, size_t skip, int flags)
{
	if (!h || (!blocks && nbytes)) return NGSCOV_EINVAL;
	if (h->runs_done || h->last_pushed) return NGSCOV_ESTATE;
	if (h->sticky_err) return h->sticky_err;
	CK(cudaSetDevice(h->device));
	int r;
	if ((r = ensure_chunk(h, 0)) || (r = ensure_chunk(h, 1))) return r;
	/* Chunk plan: whole blocks, at most stage_cap inflated bytes (and desc_blocks blocks) per chunk.  The host only reads
	 * each block's BSIZE and ISIZE, but that is one dependent cache miss per block over the whole file: the plan of chunk
	 * k+1 is therefore made while the GPU inflates chunk k (it used to be made for the whole file first: 60-100 ms
	 * before the first kernel on the 30x BAM).  Descriptors are written straight into pinned memory. */
	const uint8_t *src = (const uint8_t *)blocks;
	const bool dbg = getenv("NGSCOV_DEBUG") != nullptr;
	auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double t_parse = 0, t_res = 0, t_stage = 0, t_launch = 0;
	struct Chunk { size_t c0, c1, nb, ulen; };
	const size_t desc_blocks = h->stage_cap / 16384 + 4096;
	for (int i = 0; i < 2; ++i) if (!h->h_desc[i]) {
		CK(cudaHostAlloc((void **)&h->h_desc[i], desc_blocks * sizeof(ngs_bgzf_desc), cudaHostAllocDefault));
		CK(cudaMalloc((void **)&h->d_desc[i], desc_blocks * sizeof(ngs_bgzf_desc)));
	}
	size_t off = 0, n_blocks = 0, n_chunks = 0, total_u = 0;
	auto plan = [&](int idx, Chunk &c) -> int {                   /* next chunk from `off`; c.nb == 0: end of input */
		const double t0 = now();
		CK(cudaEventSynchronize(h->ev_h2d[idx]));                 /* the copy that last read h_desc[idx] */
		ngs_bgzf_desc *dd = h->h_desc[idx];
		c.c0 = off; c.nb = 0; c.ulen = 0;
		while (off < nbytes && c.nb < desc_blocks) {
			if (off + 18 > nbytes) return NGSCOV_EFORMAT;
			const uint8_t *p = src + off;
			if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4) || p[12] != 'B' || p[13] != 'C') return NGSCOV_EFORMAT;
			const size_t bsize = (size_t)(p[16] | p[17] << 8) + 1;
			if (bsize < 26 || off + bsize > nbytes) return NGSCOV_EFORMAT;
			uint32_t isize; memcpy(&isize, p + bsize - 4, 4);
			if (isize > 65536 || isize > h->stage_cap) return NGSCOV_EFORMAT;
			if (c.ulen + isize > h->stage_cap) break;
			ngs_bgzf_desc d; d.src = off + 18 - c.c0; d.dst = c.ulen; d.clen = (uint32_t)(bsize - 26); d.ulen = isize;
			dd[c.nb++] = d; c.ulen += isize; off += bsize;
		}
		c.c1 = off; n_blocks += c.nb; total_u += c.ulen;
		t_parse += now() - t0;
		return 0;
	};
	auto stage_in = [&](const Chunk &c, int idx) -> int {         /* compressed bytes + descriptors of a chunk -> device */
		const size_t cb = c.c1 - c.c0;
		if (cb + 16 > h->comp_cap[idx]) {
			CK(cudaStreamSynchronize(h->s_main)); CK(cudaStreamSynchronize(h->s_copy));
			if (h->d_comp[idx]) cudaFree(h->d_comp[idx]);
			h->d_comp[idx] = nullptr; h->comp_cap[idx] = 0;
			const size_t want = cb + cb / 8 + 4096;
			CK(cudaMalloc((void **)&h->d_comp[idx], want)); h->comp_cap[idx] = want;
		}
		/* no device-side wait: the inflate that last read these two buffers precedes a verification the host has
		 * already waited for */
		CK(cudaEventRecord(next_event(h->ev_cp, h->n_cp), h->s_copy));
		CK(cudaMemcpyAsync(h->d_comp[idx], src + c.c0, cb, cudaMemcpyHostToDevice, h->s_copy));
		CK(cudaMemcpyAsync(h->d_desc[idx], h->h_desc[idx], c.nb * sizeof(ngs_bgzf_desc), cudaMemcpyHostToDevice, h->s_copy));
		CK(cudaEventRecord(next_event(h->ev_cp, h->n_cp), h->s_copy));
		CK(cudaEventRecord(h->ev_h2d[idx], h->s_copy));
		return 0;
	};
	if (flags & NGSCOV_PUSH_LAST) { h->last_pushed = true; h->last_flags = flags; }
	if ((r = resolve_pending(h))) return r;
	Chunk cur, nxt;
	int idx = 0;
	if ((r = plan(idx, cur))) return h->sticky_err = r;
	if (cur.nb && (r = stage_in(cur, idx))) return r;
	int sm = 148;
	cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, h->device);
	for (; cur.nb; cur = nxt, idx ^= 1, ++n_chunks) {
		const Chunk &c = cur;
		/* the inflate of chunk k is queued BEFORE the host waits for the verification of chunk k-1: it depends on
		 * nothing of it (other chunk buffer, no engine state), and it keeps the GPU busy while the host round trip --
		 * plan k+1, wait, finalize k-1, copy k+1, frame k -- happens (measured: 4.5 ms of idle GPU per chunk otherwise) */
		double t0 = now();
		uint8_t *data = h->d_chunk[idx] + NGS_HEADROOM;
		CK(cudaStreamWaitEvent(h->s_main, h->ev_h2d[idx], 0));
		const uint32_t nb = (uint32_t)c.nb;
		CK(cudaEventRecord(next_event(h->ev_inf, h->n_inf), h->s_main));
		launch_inflate(h, idx, nb, data, sm);
		CK(cudaEventRecord(next_event(h->ev_inf, h->n_inf), h->s_main));
		h->stats.launches += 1;
		CK(cudaGetLastError());
		t_launch += now() - t0;
		if ((r = plan(idx ^ 1, nxt))) return h->sticky_err = r;
		t0 = now();
		if ((r = resolve_pending(h, true))) return r;                /* chunk k-1 verified, its buffers are free */
		t_res += now() - t0; t0 = now();
		if (nxt.nb && (r = stage_in(nxt, idx ^ 1))) return r;
		t_stage += now() - t0; t0 = now();
		const size_t sk = std::min(skip, c.ulen);
		skip -= sk;
		h->stats.stream_bytes += c.ulen - sk;
		if (c.ulen == sk) {                                          /* header-only chunk: nothing is submitted, so nothing orders the
		                                                                inflate that reads d_comp / d_desc[idx] before their reuse */
			CK(cudaStreamSynchronize(h->s_main));
			continue;
		}
		if (!h->shift) {
			int32_t bs = 0;
			CK(cudaMemcpyAsync(&bs, data + sk, 4, cudaMemcpyDeviceToHost, h->s_main));
			CK(cudaStreamSynchronize(h->s_main));
			pick_shift(h, bs);
		}
		k_place_carry<<<1, 256, 0, h->s_main>>>(data + sk, h->d_st, h->d_carry);
		h->stats.launches += 1;
		if ((r = submit_chunk(h, data + sk, (int64_t)(c.ulen - sk), true, 0))) return r;
		t_launch += now() - t0;
	}
	if (dbg) fprintf(stderr, "ngscov_push_bgzf: %zu blocks, %zu chunks; host seconds: plan %.4f resolve %.4f stage_in %.4f launch %.4f\n",
	                 n_blocks, n_chunks, t_parse, t_res, t_stage, t_launch);
	if (skip) return h->sticky_err = NGSCOV_EINVAL;                  /* skip beyond the inflated size of the blocks */
	return 0;
}

extern "C" int ngscov_push_device(ngscov_t *h, const void *dev_records, size_t nbytes, int flags)
{
	if (!h || (!dev_records && nbytes)) return NGSCOV_EINVAL;
	if (h->runs_done || h->last_pushed) return NGSCOV_ESTATE;
	if (h->sticky_err) return h->sticky_err;
	CK(cudaSetDevice(h->device));
	int r = resolve_pending(h);
	if (r) return r;
	/* whole records only: the previous push must have ended on a record boundary */
	if (h->stats.stream_bytes) {
		CK(cudaMemcpyAsync(h->h_st, h->d_st, sizeof(ngs_dev_state), cudaMemcpyDeviceToHost, h->s_main));
		CK(cudaStreamSynchronize(h->s_main));
		if (h->h_st->entry_next != 0 || h->h_st->carry_len != 0) { h->sticky_err = NGSCOV_EFORMAT; return NGSCOV_EFORMAT; }
	}
	if (flags & NGSCOV_PUSH_LAST) { h->last_pushed = true; h->last_flags = flags; }
	h->stats.stream_bytes += nbytes;
	if (nbytes == 0) return 0;
	if (!h->shift) {
		int32_t bs = 0;
		CK(cudaMemcpy(&bs, dev_records, 4, cudaMemcpyDeviceToHost));
		pick_shift(h, bs);
	}
	return submit_chunk(h, (const uint8_t *)dev_records, (int64_t)nbytes, false, 0);
}

template <class T>
static int ensure_buf(ngscov *h, T **p, size_t *cap, size_t need)
{
	if (need <= *cap && *p) return 0;
	CK(cudaStreamSynchronize(h->s_main));
	if (*p) cudaFree(*p);
	*p = nullptr; *cap = 0;
	const size_t want = need + need / 16 + 1024;
	CK(cudaMalloc((void **)p, want * sizeof(T)));
	*cap = want;
	return 0;
}

/* Scan + run compaction + window bins over segments [s0, s1).  Records are written from index 0 of the
 * run buffers; h_rec_base[0..s1-s0] receives the per-segment record offsets, h_st2 the counters. */
static int scan_range(ngscov *h, size_t s0, size_t s1, unsigned long long mops_bound)
{
	int r;
	const uint64_t t0 = h->seg[s0].tile0, t1 = s1 < h->seg.size() ? h->seg[s1].tile0 : h->n_tiles;
	const uint64_t elems = (t1 - t0) * (uint64_t)NGS_TILE;
	/* boundary records: at most two per accumulated M interval, never more than tracked positions */
	const size_t need = (size_t)std::min<unsigned long long>(2ull * mops_bound, elems) + 16;
	{
		size_t cap0 = h->run_cap, cap1 = h->run_cap, cap2 = h->run_cap;
		if ((r = ensure_buf(h, &h->d_run_start, &cap0, need))) return r;
		if ((r = ensure_buf(h, &h->d_run_end, &cap1, need))) return r;
		if ((r = ensure_buf(h, &h->d_run_depth, &cap2, need))) return r;
		h->run_cap = std::min(cap0, std::min(cap1, cap2));
	}
	tic(h);
	if (t1 > t0) {
		CK(cudaMemsetAsync(h->d_stA + t0, 0, (t1 - t0) * 8, h->s_main));
		CK(cudaMemsetAsync(h->d_stB + t0, 0, (t1 - t0) * 8, h->s_main));
	}
	if (!h->bins_zeroed) { CK(cudaMemsetAsync(h->d_bins, 0, h->n_bins * 8, h->s_main)); h->bins_zeroed = true; }
	{	/* reset only the scan fields of the device state block (the ingest fields stay live) */
		const size_t o0 = offsetof(ngs_dev_state, scan_ticket), o1 = offsetof(ngs_dev_state, fmt_ticket);
		CK(cudaMemcpyAsync((char *)h->d_st + o0, (const char *)h->h_tmpl + o0, o1 - o0, cudaMemcpyHostToDevice, h->s_main));
	}
	if (t1 > t0) {
		ScanArgs A;
		A.diff = h->d_diff + t0 * (uint64_t)NGS_TILE; A.n_tiles = t1 - t0; A.tile_off = t0;
		A.tile2seg = h->d_tile2seg + t0; A.seg = h->d_seg;
		A.stA = h->d_stA + t0; A.stB = h->d_stB + t0;
		A.run_start = h->d_run_start; A.run_end = h->d_run_end; A.run_depth = h->d_run_depth; A.run_cap = h->run_cap;
		A.rec_base = h->d_rec_base - s0; A.n_seg = (uint32_t)s1;      /* the kernel indexes by global segment */
		A.bins = h->d_bins; A.window = h->cfg.window; A.variant = h->cfg.mode == NGSCOV_MODE_WIG ? 1 : 0;
		A.rezero = 1; A.st = h->d_st;
		A.debug = getenv("NGSCOV_DEBUG_SCAN") ? atoi(getenv("NGSCOV_DEBUG_SCAN")) : 0;
		k_scan_runs<<<(unsigned)(t1 - t0), NGS_SCAN_THREADS, 0, h->s_main>>>(A);
		h->stats.launches += 1;
		CK(cudaGetLastError());
	}
	if ((r = small_d2h(h, h->h_st2, h->d_st, sizeof(ngs_dev_state)))) return r;
	if (t1 > t0 && (r = small_d2h(h, h->h_recb, h->d_rec_base, (s1 - s0 + 1) * 8))) return r;
	h->stats.ms_scan += toc(h);
	h->h_rec_base.assign(h->seg.size() + 1, 0);
	if (t1 > t0) memcpy(h->h_rec_base.data(), h->h_recb, (s1 - s0 + 1) * 8);
	const ngs_dev_state &s2 = *h->h_st2;
	h->acc_boundaries += s2.n_boundaries; h->acc_runs += s2.n_runs;
	h->acc_max = std::max(h->acc_max, s2.max_depth); h->acc_min = std::min(h->acc_min, s2.min_depth);
	h->stats.boundaries = h->acc_boundaries; h->stats.runs = h->acc_runs; h->stats.positions = h->n_elems;
	if (s2.n_boundaries > h->run_cap) return h->sticky_err = NGSCOV_ENOMEM;
	return 0;
}

static int format_bed_range(ngscov *h, size_t s0, size_t s1);
static int grow_host_text(ngscov *h, int which, size_t need, size_t keep);
static int cover_tiles(ngscov *h, uint64_t t0, uint64_t t1, unsigned long long mops_bound, bool defer = false);
static int cover_harvest(ngscov *h);
static int stream_finalize_cover(ngscov *h);

/* Finalize segments [next_fin_seg, seg_end) while the rest of the stream is still arriving. */
static int stream_finalize_impl(ngscov *h, size_t seg_end);
static int stream_finalize(ngscov *h, size_t seg_end)
{
	const auto t0 = std::chrono::steady_clock::now();
	const int r = stream_finalize_impl(h, seg_end);
	h->fin_wall += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); ++h->fin_calls;
	return r;
}
static int stream_finalize_impl(ngscov *h, size_t seg_end)
{
	int r;
	const unsigned long long mops_now = h->h_st->m_ops;
	const size_t s0 = h->next_fin_seg;
	if ((r = scan_range(h, s0, seg_end, mops_now - h->mops_fin))) return r;
	/* how many of the intervals seen so far belong to the finished contigs is not known exactly; every interval
	 * makes at most two boundaries, so boundaries/2 is a safe lower bound of what they consumed */
	h->mops_fin += h->h_st2->n_boundaries / 2;
	if (h->cfg.mode == NGSCOV_MODE_DEPTH && (r = format_bed_range(h, s0, seg_end))) return r;
	h->next_fin_seg = seg_end;
	return 0;
}

/* bam_sliding_count: the per-window sums come back (a few MB), calc_winGC (bam_sliding_count.c:126-138) and
 * output_count_GC (:140-165) run on the host in the reference's own types -- float arithmetic included -- over
 * O(windows) values. */
static int slide_finish(ngscov *h)
{
	const size_t nb = (size_t)h->n_bins;
	std::vector<unsigned int> cnt(nb + 1), len(nb + 1), ctg((size_t)h->n_ref + 1);
	std::vector<unsigned long long> gc(nb + 1);
	CK(cudaMemcpyAsync(cnt.data(), h->d_sl_cnt, nb * 4, cudaMemcpyDeviceToHost, h->s_main));
	CK(cudaMemcpyAsync(len.data(), h->d_sl_len, nb * 4, cudaMemcpyDeviceToHost, h->s_main));
	CK(cudaMemcpyAsync(gc.data(), h->d_sl_gc, nb * 8, cudaMemcpyDeviceToHost, h->s_main));
	CK(cudaMemcpyAsync(ctg.data(), h->d_sl_ctg, (size_t)h->n_ref * 4, cudaMemcpyDeviceToHost, h->s_main));
	CK(cudaMemcpyAsync(&h->slide_ncount, h->d_sl_ncount, 8, cudaMemcpyDeviceToHost, h->s_main));
	CK(cudaStreamSynchronize(h->s_main));
	std::vector<float> GC(nb + 1);
	/* the reference adds the per-read counts in a float, in file order: exact below 2^24, otherwise redone its way */
	std::vector<uint32_t> redo;
	for (size_t k = 0; k < nb; ++k) { GC[k] = (float)gc[k]; if (gc[k] >= (1ull << 24)) redo.push_back((uint32_t)k); }
	if (!redo.empty()) {
		if (h->h_st->unsorted) return h->sticky_err = NGSCOV_EUNSORTED;            /* file order is not table order */
		for (uint32_t k : redo) {
			const size_t c = std::upper_bound(h->win_base.begin(), h->win_base.begin() + h->n_ref, (uint64_t)k) - h->win_base.begin() - 1;
			if (h->ref_len[c] / h->cfg.window + 1 > 65536u) return h->sticky_err = NGSCOV_ERANGE;   /* window index wraps in the reference's unsigned short */
		}
		uint32_t *d_list = nullptr; float *d_out = nullptr; unsigned int *d_err = nullptr;
		CK(cudaMalloc((void **)&d_list, redo.size() * 4)); CK(cudaMalloc((void **)&d_out, redo.size() * 4)); CK(cudaMalloc((void **)&d_err, 4));
		CK(cudaMemcpyAsync(d_list, redo.data(), redo.size() * 4, cudaMemcpyHostToDevice, h->s_main));
		CK(cudaMemsetAsync(d_err, 0, 4, h->s_main));
		k_slide_seqfloat<<<(unsigned)((redo.size() + 63) / 64), 64, 0, h->s_main>>>(d_list, (uint32_t)redo.size(), h->d_ent, h->d_rows, h->d_rkmax, (uint32_t)h->rows_used,
		                                                                       h->d_win_base, (uint32_t)h->n_ref, (int)h->cfg.window, d_out, d_err);
		h->stats.launches += 1;
		std::vector<float> out(redo.size()); unsigned int err = 0;
		CK(cudaMemcpyAsync(out.data(), d_out, redo.size() * 4, cudaMemcpyDeviceToHost, h->s_main));
		CK(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, h->s_main));
		CK(cudaStreamSynchronize(h->s_main));
		cudaFree(d_list); cudaFree(d_out); cudaFree(d_err);
		if (err) return h->sticky_err = NGSCOV_ERANGE;
		for (size_t i = 0; i < redo.size(); ++i) GC[redo[i]] = out[i];
	}
	const int window = (int)h->cfg.window;
	std::string &t = h->slide_text;
	t.clear();
	char buf[512];
	t += "#chr\tchr_len\tchr_sum_read_count\tchr_sum_base\tchr_mean_cov\tchr_mean_GC%";
	uint32_t max_len = 0;
	for (unsigned short i = 0; i < h->n_ref; ++i) max_len = h->ref_len[i] > max_len ? h->ref_len[i] : max_len;      /* max(), :198-205 */
	const uint32_t max_target_window = max_len / window + 1;
	for (uint32_t k = 0; k < max_target_window; ++k) { snprintf(buf, sizeof buf, "\t%u\tcount\tGC%%", k + 1); t += buf; }
	t += "\n";
	for (int j = 0; j < h->n_ref; ++j) {
		if (ctg[j] == 0) continue;                                  /* windows[j] == 0: no mapped read on this contig */
		const uint32_t windows = (unsigned int)h->ref_len[j] / window + 1;
		const size_t b0 = (size_t)h->win_base[j];
		unsigned int sum_count = 0; unsigned long sum_base = 0; float sum_GC = 0;
		for (uint32_t k = 0; k < windows; ++k) {                    /* calc_winGC */
			sum_count += cnt[b0 + k];
			sum_GC += GC[b0 + k];
			sum_base += len[b0 + k];
			GC[b0 + k] = GC[b0 + k] != 0 ? GC[b0 + k] / len[b0 + k] * 100 : GC[b0 + k];
		}
		sum_GC = sum_GC / sum_base * 100;
		snprintf(buf, sizeof buf, "%s\t%d\t%u\t%lu\t%f\t%f", h->ref_name[j].c_str(), h->ref_len[j], sum_count, sum_base, (double)sum_base / h->ref_len[j], sum_GC);
		t += buf;
		for (uint32_t k = 0; k < windows; ++k) { snprintf(buf, sizeof buf, "\t%d\t%u\t%f", k + 1, cnt[b0 + k], GC[b0 + k]); t += buf; }
		t += "\n";
	}
	h->stats.out_bytes[0] = t.size();
	return 0;
}

extern "C" int ngscov_finish_runs(ngscov_t *h)
{
	if (!h) return NGSCOV_EINVAL;
	if (h->runs_done) return NGSCOV_ESTATE;
	if (h->sticky_err) return h->sticky_err;
	CK(cudaSetDevice(h->device));
	int r = resolve_pending(h);
	if (r) return r;
	CK(cudaMemcpyAsync(h->h_st, h->d_st, sizeof(ngs_dev_state), cudaMemcpyDeviceToHost, h->s_main));
	CK(cudaMemcpyAsync(h->h_ctr, h->d_ctr, NGS_CTR_STRIPES * 16 * 8, cudaMemcpyDeviceToHost, h->s_main));
	CK(cudaStreamSynchronize(h->s_main));
	CK(cudaStreamSynchronize(h->s_copy));
	for (int i = 0; i < NGS_CTR_STRIPES; ++i) {     /* K2 counts in stripes; the fused variant counts in the state block */
		h->h_st->records += h->h_ctr[i * 16 + 0]; h->h_st->kept += h->h_ctr[i * 16 + 1];
		h->h_st->m_ops += h->h_ctr[i * 16 + 2]; h->h_st->cigar_words += h->h_ctr[i * 16 + 3];
	}
	h->stats.ms_ingest += sum_events(h->ev_ing, h->n_ing);
	h->stats.ms_h2d += sum_events(h->ev_cp, h->n_cp);
	if (getenv("NGSCOV_DEBUG") && h->n_inf >= 4) {
		/* device timeline of the BGZF path: inflate kernels and what lies between them */
		cudaEventSynchronize(h->ev_inf[h->n_inf - 1]);
		float inf = 0, gap = 0, span = 0, g;
		for (size_t i = 0; i + 1 < h->n_inf; i += 2) { cudaEventElapsedTime(&g, h->ev_inf[i], h->ev_inf[i + 1]); inf += g; }
		for (size_t i = 1; i + 1 < h->n_inf; i += 2) { cudaEventElapsedTime(&g, h->ev_inf[i], h->ev_inf[i + 1]); gap += g; }
		cudaEventElapsedTime(&span, h->ev_inf[0], h->ev_inf[h->n_inf - 1]);
		fprintf(stderr, "ngscov: bgzf timeline: %zu inflate launches %.2f ms, between them %.2f ms, first..last %.2f ms\n", h->n_inf / 2, inf, gap, span);
	}
	h->stats.ms_inflate += sum_events(h->ev_inf, h->n_inf);
	h->stats.ms_decode += sum_events(h->ev_k2, h->n_k2);
	const ngs_dev_state st = *h->h_st;
	h->stats.records = st.records; h->stats.kept = st.kept; h->stats.m_ops = st.m_ops; h->stats.cigar_words = st.cigar_words;
	h->stats.spec_subchunks = st.spec_walks; h->stats.spec_repaired = st.spec_repaired;
	if (st.flags & NGS_FLAG_EFORMAT) return h->sticky_err = NGSCOV_EFORMAT;
	if ((st.entry_next != 0 || st.carry_len != 0) && !(h->last_flags & NGSCOV_PUSH_TRUNCATED_OK)) return h->sticky_err = NGSCOV_EFORMAT;
	if (st.flags & NGS_FLAG_ERANGE) return h->sticky_err = NGSCOV_ERANGE;

	if (h->slide) {
		if ((r = slide_finish(h))) return r;
		h->diff_dirty = false; h->runs_done = true;
		return 0;
	}
	if (h->cover && !h->red_mode) {
		if (!h->seg.empty() && h->fin_tile < h->n_tiles * NGS_CPT && (r = cover_tiles(h, h->fin_tile, h->n_tiles * NGS_CPT, st.m_ops - h->mops_fin))) return r;
		if ((r = cover_harvest(h))) return r;
		if (h->cfg.mode == NGSCOV_MODE_DEPTH) { h->text_ready[NGSCOV_OUT_BEDGRAPH] = true; h->text_len[NGSCOV_OUT_BEDGRAPH] = h->bed_off; }
		h->next_fin_seg = h->seg.size();
		h->diff_dirty = false;
		memset(&h->halo, 0, sizeof h->halo); h->halo.first.contig = h->halo.last.contig = -1;
		h->halo.n_runs = h->acc_runs;
		if (h->halo_have) { h->halo.first = h->cov_first; h->halo.last = h->cov_last; }
		h->stats.boundaries = h->acc_boundaries; h->stats.runs = h->acc_runs; h->stats.positions = h->n_elems;
		if (getenv("NGSCOV_DEBUG")) { fprintf(stderr, "ngscov: streamed finalize: %d calls, %.3f s host wall\n", h->fin_calls, h->fin_wall); h->fin_wall = 0; h->fin_calls = 0; }
		h->runs_done = true;
		return 0;
	}
	const bool streamed = h->next_fin_seg > 0;
	if (h->next_fin_seg < h->seg.size() || h->seg.empty()) {
		if (h->seg.empty()) { h->h_rec_base.assign(1, 0); memset(h->h_st2, 0, sizeof *h->h_st2); }
		else if ((r = scan_range(h, h->next_fin_seg, h->seg.size(), st.m_ops - h->mops_fin))) return r;
	}
	h->diff_dirty = false;
	memset(&h->halo, 0, sizeof h->halo); h->halo.first.contig = h->halo.last.contig = -1;
	if (streamed) {
		/* the remaining contigs' text is produced here too so that everything is in flight to the host */
		if (h->cfg.mode == NGSCOV_MODE_DEPTH && h->next_fin_seg < h->seg.size() && (r = format_bed_range(h, h->next_fin_seg, h->seg.size()))) return r;
		h->next_fin_seg = h->seg.size();
		h->halo.n_runs = h->acc_runs;                    /* first / last are not tracked in streamed mode (single shard) */
	} else {
		const ngs_dev_state &s2 = *h->h_st2;
		h->halo.n_runs = s2.n_runs;
		if (s2.n_runs) {
			const unsigned long long idx[2] = { s2.first_run, s2.last_run };
			ngscov_run *dst[2] = { &h->halo.first, &h->halo.last };
			for (int k = 0; k < 2; ++k) {
				CK(cudaMemcpy(&dst[k]->start, h->d_run_start + idx[k], 4, cudaMemcpyDeviceToHost));
				CK(cudaMemcpy(&dst[k]->end, h->d_run_end + idx[k], 4, cudaMemcpyDeviceToHost));
				CK(cudaMemcpy(&dst[k]->depth, h->d_run_depth + idx[k], 4, cudaMemcpyDeviceToHost));
				size_t sg = std::upper_bound(h->h_rec_base.begin(), h->h_rec_base.begin() + h->seg.size(), idx[k]) - h->h_rec_base.begin() - 1;
				dst[k]->contig = (int32_t)h->seg[sg].contig;
			}
		}
	}
	if (getenv("NGSCOV_DEBUG")) { fprintf(stderr, "ngscov: streamed finalize: %d calls, %.3f s host wall\n", h->fin_calls, h->fin_wall); h->fin_wall = 0; h->fin_calls = 0; }
	h->runs_done = true;
	return 0;
}

extern "C" int ngscov_slice_done(ngscov_t *h, int *done)
{
	if (!h || !done) return NGSCOV_EINVAL;
	*done = 0;
	if (h->sticky_err) return h->sticky_err;
	if (h->end_ctg < 0) return 0;                                  /* no slice: never done before the last push */
	CK(cudaSetDevice(h->device));
	int r = resolve_pending(h);
	if (r) return r;
	unsigned int mr = 0;
	if ((r = small_d2h(h, &h->h_st->max_ref_p1, &h->d_st->max_ref_p1, sizeof mr))) return r;
	CK(cudaStreamSynchronize(h->s_main));
	mr = h->h_st->max_ref_p1;
	*done = mr >= (unsigned int)h->end_ctg + 2u ? 1 : 0;
	return 0;
}

/* NGSCOV_MODE_SLIDE: count only the records bam_fetch(ref, beg, end) hands out -- the `-r` mode of the reference
 * (bam_sliding_count.c:393-409).  Before the first push; ref < 0 clears the region. */
extern "C" int ngscov_set_region(ngscov_t *h, int ref, int beg, int end)
{
	if (!h) return NGSCOV_EINVAL;
	if (!h->slide || ref >= h->n_ref) return NGSCOV_EINVAL;
	if (h->stats.stream_bytes || h->runs_done || h->last_pushed) return NGSCOV_ESTATE;
	h->reg_ref = ref < 0 ? -1 : ref; h->reg_beg = beg; h->reg_end = end;
	return 0;
}

extern "C" int ngscov_contigs_done(ngscov_t *h, uint32_t *n)
{
	if (!h || !n) return NGSCOV_EINVAL;
	if (h->runs_done) { *n = (uint32_t)h->n_ref; return 0; }
	*n = h->next_fin_seg < h->seg.size() ? h->seg[h->next_fin_seg].contig : (h->next_fin_seg ? (uint32_t)h->n_ref : 0u);
	if (h->next_fin_seg == 0) *n = 0;
	return 0;
}

extern "C" int ngscov_halo_get(ngscov_t *h, ngscov_halo *halo)
{
	if (!h || !halo) return NGSCOV_EINVAL;
	if (!h->runs_done) return NGSCOV_ESTATE;
	*halo = h->halo;
	return 0;
}

extern "C" int ngscov_halo_put(ngscov_t *h, int drop_first, uint32_t new_last_end)
{
	if (!h) return NGSCOV_EINVAL;
	if (!h->runs_done || h->formatted) return NGSCOV_ESTATE;
	if (h->halo.n_runs == 0) return drop_first || new_last_end ? NGSCOV_EINVAL : 0;
	if (h->cover && !h->red_mode) {
		/* the text exists already: dropping the first run skips its line, a new end of the last run rewrites the last line */
		auto line = [&](const ngscov_run &rn) { char b[64]; snprintf(b, sizeof b, "\t%u\t%u\t%d\n", rn.start, rn.end, rn.depth); return h->ref_name[rn.contig] + b; };
		h->bed_begin = 0; h->bed_tail.clear(); h->bed_tail_old = 0; h->bed_patched = false;
		if (h->cfg.mode != NGSCOV_MODE_DEPTH) return 0;
		if (drop_first) h->bed_begin = h->halo.n_runs == 1 ? h->bed_off : line(h->halo.first).size();
		if (new_last_end && !(drop_first && h->halo.n_runs == 1)) {
			ngscov_run l = h->halo.last;
			h->bed_tail_old = line(l).size();
			l.end = new_last_end;
			h->bed_tail = line(l);
		}
		return 0;
	}
	if (h->next_fin_seg > 0) return NGSCOV_ESTATE;           /* streamed finalize: the text is already made */
	h->drop_idx = drop_first ? h->h_st2->first_run : ~0ull;
	if (new_last_end) { h->mod_idx = h->h_st2->last_run; h->mod_end = new_last_end; } else h->mod_idx = ~0ull;
	return 0;
}

extern "C" int ngscov_halo_device(ngscov_t *h, const void **dev_halo)
{
	if (!h || !dev_halo) return NGSCOV_EINVAL;
	if (!h->runs_done || !h->cover || h->red_mode) return NGSCOV_ESTATE;
	*dev_halo = h->d_halo;
	return 0;
}

extern "C" int ngscov_bins_device(ngscov_t *h, void **dev_bins, uint64_t *n)
{
	if (!h || !dev_bins || !n) return NGSCOV_EINVAL;
	if (!h->runs_done) return NGSCOV_ESTATE;
	*dev_bins = h->d_bins; *n = h->n_bins;
	return 0;
}

extern "C" int ngscov_stitch_device(ngscov_t *h, const void *dev_all_halos, int world, int rank)
{
	if (!h || !dev_all_halos || world < 1 || rank < 0 || rank >= world) return NGSCOV_EINVAL;
	if (!h->runs_done || h->formatted || !h->cover || h->red_mode) return NGSCOV_ESTATE;
	CK(cudaSetDevice(h->device));
	const int do_text = h->cfg.mode == NGSCOV_MODE_DEPTH ? 1 : 0;
	NameTable nt = { h->d_names, h->d_name_off, h->d_name8 };
	k_stitch<<<1, 32, 0, h->s_main>>>((const long long *)dev_all_halos, world, rank, do_text ? h->d_text[NGSCOV_OUT_BEDGRAPH] : nullptr, h->bed_off, nt,
	                                  h->d_bins, h->d_win_base, h->d_ref_len, h->cfg.window, h->cfg.mode == NGSCOV_MODE_WIG ? 1 : 0, do_text, h->d_patch, h->h_patch);
	h->stats.launches += 1;
	CK(cudaGetLastError());
	h->dev_stitched = true;
	if (do_text && h->bed_d2h) {
		/* streamed finalize already sent the text to the host: send the (possibly rewritten) last line again */
		const size_t win = (size_t)h->max_name_tab + 64, lo = h->bed_off > win ? h->bed_off - win : 0, hi = h->bed_off + 32;
		int r = grow_host_text(h, NGSCOV_OUT_BEDGRAPH, hi, h->bed_d2h);
		if (r) return r;
		CK(cudaStreamSynchronize(h->s_out));
		CK(cudaMemcpyAsync(h->h_text[NGSCOV_OUT_BEDGRAPH] + lo, h->d_text[NGSCOV_OUT_BEDGRAPH] + lo, hi - lo, cudaMemcpyDeviceToHost, h->s_main));
	}
	return 0;
}

extern "C" int ngscov_bins_count(ngscov_t *h, uint64_t *n)
{
	if (!h || !n) return NGSCOV_EINVAL;
	*n = h->n_bins;
	return 0;
}

extern "C" int ngscov_bins_get(ngscov_t *h, uint64_t *bins, uint64_t n)
{
	if (!h || !bins || n != h->n_bins) return NGSCOV_EINVAL;
	if (!h->runs_done) return NGSCOV_ESTATE;
	CK(cudaSetDevice(h->device));
	CK(cudaMemcpyAsync(bins, h->d_bins, n * 8, cudaMemcpyDeviceToHost, h->s_main));
	CK(cudaStreamSynchronize(h->s_main));
	return 0;
}

extern "C" int ngscov_bins_put(ngscov_t *h, const uint64_t *bins, uint64_t n)
{
	if (!h || !bins || n != h->n_bins) return NGSCOV_EINVAL;
	if (!h->runs_done || h->formatted) return NGSCOV_ESTATE;
	CK(cudaSetDevice(h->device));
	CK(cudaMemcpyAsync(h->d_bins, bins, n * 8, cudaMemcpyHostToDevice, h->s_main));
	CK(cudaStreamSynchronize(h->s_main));
	return 0;
}

struct HostBinAdd { uint64_t *bins; int sign; __host__ __device__ void operator()(uint32_t k, int64_t v) const { bins[k] += (uint64_t)(sign * v); } };

extern "C" int ngscov_bins_apply_run(ngscov_t *h, const ngscov_run *run, int sign, uint64_t *bins, uint64_t n)
{
	if (!h || !run || !bins || n != h->n_bins || run->contig < 0 || run->contig >= h->n_ref) return NGSCOV_EINVAL;
	const uint32_t L = h->ref_len[run->contig];
	HostBinAdd add = { bins + h->win_base[run->contig], sign < 0 ? -1 : 1 };
	ngs_bin_run(h->cfg.mode == NGSCOV_MODE_WIG ? 1 : 0, run->start, run->end, (int64_t)run->depth, h->cfg.window, L, L / h->cfg.window + 1, add);
	return 0;
}

/* make room for `need` bytes in a text buffer, keeping its first `keep` bytes */
static int grow_text(ngscov *h, int which, size_t need, size_t keep)
{
	if (need <= h->d_text_cap[which] && h->d_text[which]) return 0;
	char *nb = nullptr;
	const size_t want = need + need / 4 + 4096;
	CK(cudaMalloc((void **)&nb, want));
	if (keep && h->d_text[which]) CK(cudaMemcpyAsync(nb, h->d_text[which], keep, cudaMemcpyDeviceToDevice, h->s_main));
	CK(cudaStreamSynchronize(h->s_main));
	CK(cudaStreamSynchronize(h->s_out));
	if (h->d_text[which]) cudaFree(h->d_text[which]);
	h->d_text[which] = nb; h->d_text_cap[which] = want;
	return 0;
}
static int grow_host_text(ngscov *h, int which, size_t need, size_t keep)
{
	if (need <= h->h_text_cap[which] && h->h_text[which]) return 0;
	char *nb = nullptr;
	const size_t want = need + need / 4 + 4096;
	CK(cudaHostAlloc((void **)&nb, want, cudaHostAllocDefault));
	CK(cudaStreamSynchronize(h->s_out));
	if (keep && h->h_text[which]) memcpy(nb, h->h_text[which], keep);
	if (h->h_text[which]) cudaFreeHost(h->h_text[which]);
	h->h_text[which] = nb; h->h_text_cap[which] = want;
	return 0;
}

template <class Producer>
static int run_format(ngscov *h, int which, unsigned long long n_items, const Producer &P, uint32_t max_line, size_t bound, size_t out_off = 0)
{
	int r;
	if ((r = grow_text(h, which, out_off + bound + 64, out_off))) return r;
	int ipt = 4;
	if (getenv("NGSCOV_FMT_IPT")) { const int v = atoi(getenv("NGSCOV_FMT_IPT")); if (v == 1 || v == 2 || v == 4) ipt = v; }
	const uint32_t smem_max = 96 * 1024;
	while (ipt > 1 && (size_t)NGS_FMT_THREADS * ipt * max_line + 32 > smem_max) ipt >>= 1;
	uint32_t stage = (uint32_t)std::min<size_t>((size_t)NGS_FMT_THREADS * ipt * max_line + 32, smem_max);
	const unsigned long long per = (unsigned long long)NGS_FMT_THREADS * ipt;
	const unsigned long long n_tiles = (n_items + per - 1) / per;
	h->text_len[which] = out_off; h->text_ready[which] = true; h->h_text_valid[which] = false;
	if (n_tiles == 0) return 0;
	size_t scap = h->fmt_status_cap;
	if ((r = ensure_buf(h, &h->d_fmt_status, &scap, (size_t)n_tiles))) return r;
	h->fmt_status_cap = scap;
	tic(h);
	CK(cudaMemsetAsync(h->d_fmt_status, 0, n_tiles * 8, h->s_main));
	CK(cudaMemsetAsync(&h->d_st->fmt_ticket, 0, 16, h->s_main));      /* fmt_ticket + fmt_bytes */
	FmtArgs A;
	A.n_items = n_items; A.n_tiles = n_tiles; A.status = h->d_fmt_status; A.out = h->d_text[which] + out_off;
	A.out_cap = h->d_text_cap[which] - out_off; A.st = h->d_st; A.stage_bytes = stage;
	A.debug = getenv("NGSCOV_DEBUG_FMT") ? atoi(getenv("NGSCOV_DEBUG_FMT")) : 0;
	bool fast_bed = false;
	if constexpr (std::is_same<Producer, BedProducer>::value) {
		/* bedGraph: register-assembled lines (k_format_bed); NGSCOV_FMT=generic keeps the byte-wise producer path */
		const char *e = getenv("NGSCOV_FMT");
		fast_bed = ipt == 4 && !(e && !strcmp(e, "generic"));
		if (fast_bed) {
			const int minb = getenv("NGSCOV_FMT_MINB") ? atoi(getenv("NGSCOV_FMT_MINB")) : 6;   /* measured at 62 M reads: 1.436 / 1.457 / 1.343 ms for 4 / 5 / 6 */
			if (minb == 6) {
				CK(cudaFuncSetAttribute(k_format_bed<4, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage));
				k_format_bed<4, 6><<<(unsigned)n_tiles, NGS_FMT_THREADS, stage, h->s_main>>>(A, P);
			} else if (minb == 5) {
				CK(cudaFuncSetAttribute(k_format_bed<4, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage));
				k_format_bed<4, 5><<<(unsigned)n_tiles, NGS_FMT_THREADS, stage, h->s_main>>>(A, P);
			} else {
				CK(cudaFuncSetAttribute(k_format_bed<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage));
				k_format_bed<4, 4><<<(unsigned)n_tiles, NGS_FMT_THREADS, stage, h->s_main>>>(A, P);
			}
		}
	}
	if (fast_bed) {
	} else if (ipt == 4) {
		CK(cudaFuncSetAttribute(k_format_lines<Producer, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage));
		k_format_lines<Producer, 4><<<(unsigned)n_tiles, NGS_FMT_THREADS, stage, h->s_main>>>(A, P);
	} else if (ipt == 2) {
		CK(cudaFuncSetAttribute(k_format_lines<Producer, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage));
		k_format_lines<Producer, 2><<<(unsigned)n_tiles, NGS_FMT_THREADS, stage, h->s_main>>>(A, P);
	} else {
		CK(cudaFuncSetAttribute(k_format_lines<Producer, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage));
		k_format_lines<Producer, 1><<<(unsigned)n_tiles, NGS_FMT_THREADS, stage, h->s_main>>>(A, P);
	}
	h->stats.launches += 1;
	CK(cudaGetLastError());
	unsigned long long *fb = &h->h_st2->fmt_bytes;
	if ((r = small_d2h(h, fb, &h->d_st->fmt_bytes, 8))) return r;
	h->stats.ms_format += toc(h);
	if (*fb > h->d_text_cap[which] - out_off) return NGSCOV_ENOMEM;
	h->text_len[which] = out_off + (size_t)*fb;
	h->stats.out_bytes[which] = h->text_len[which];
	return 0;
}


static uint32_t dig(uint64_t v) { uint32_t n = 1; while (v >= 10) { v /= 10; ++n; } return n; }

/* Cover pipeline: owner tiles [t0, t1) (global owner-tile indices, any range: a launch may start and end inside a
 * contig) -- accumulate, scan, window bins and (DEPTH mode) the bedGraph text appended at bed_off, one kernel.
 * h_st2 receives the counters. */
/* A streamed finalize (defer) does not wait for its kernels: the counters they leave (bytes of text, runs, halo) are
 * read back by cover_harvest at the next launch, one chunk later, when they have long arrived -- the host never blocks
 * on a finalize and the next chunk's kernels are queued behind it at once (the wait used to idle the GPU for about a
 * millisecond per 256 MiB chunk).  At most one launch is pending. */
static int cover_tiles(ngscov *h, uint64_t t0, uint64_t t1, unsigned long long mops_bound, bool defer)
{
	int r;
	if ((r = cover_harvest(h))) return r;
	const bool text = h->cfg.mode == NGSCOV_MODE_DEPTH;
	const uint64_t elems = (t1 - t0) * (uint64_t)NGS_CT;
	/* segments the range touches */
	size_t s0 = 0, s1 = h->seg.size();
	while (s0 + 1 < h->seg.size() && h->seg[s0 + 1].tile0 * NGS_CPT <= t0) ++s0;
	while (s1 > s0 + 1 && h->seg[s1 - 1].tile0 * NGS_CPT >= t1) --s1;
	const bool continues = t0 > h->seg[s0].tile0 * NGS_CPT;     /* the range starts inside a contig: carry from the tile before */
	size_t bound = 0; uint32_t stage = 0;
	if (text) {
		/* boundaries: at most two per accumulated interval, never more than tracked positions; depth <= intervals */
		const size_t nb = (size_t)std::min<unsigned long long>(2ull * mops_bound, elems);
		const uint32_t dd = std::min<uint32_t>(11, dig(mops_bound) + 1);
		size_t max_line = 8, typ_line = 8;
		for (size_t sg = s0; sg < s1; ++sg) {
			const uint32_t pd = dig(h->seg[sg].lim);
			const size_t nm = h->ref_name[h->seg[sg].contig].size() + 1;
			max_line = std::max(max_line, nm + 2 * pd + 2 + dd + 1);
			typ_line = std::max(typ_line, nm + 2 * pd + 2 + std::min<uint32_t>(dd, 5) + 1);
		}
		bound = nb * max_line;
		/* the staging area of a batch is sized for depths below 10^5; a batch that needs more writes its lines straight
		 * to global memory */
		stage = (uint32_t)std::min<size_t>(NGS_CBATCH * typ_line + 32, 64 * 1024);
		if (getenv("NGSCOV_COVER_STAGE")) stage = (uint32_t)atoi(getenv("NGSCOV_COVER_STAGE"));
		if ((r = grow_text(h, NGSCOV_OUT_BEDGRAPH, h->bed_off + bound + 64, h->bed_off))) return r;
		h->text_ready[NGSCOV_OUT_BEDGRAPH] = true; h->h_text_valid[NGSCOV_OUT_BEDGRAPH] = false;
	}
	CK(cudaEventRecord(h->ev_cov0, h->s_main));
	if (!h->bins_zeroed) { CK(cudaMemsetAsync(h->d_bins, 0, h->n_bins * 8, h->s_main)); h->bins_zeroed = true; }
	{	/* reset the per-launch fields of the device state block (the ingest fields stay live) */
		const size_t o0 = offsetof(ngs_dev_state, scan_ticket), o1 = offsetof(ngs_dev_state, key_seed);
		CK(cudaMemcpyAsync((char *)h->d_st + o0, (const char *)h->h_tmpl + o0, o1 - o0, cudaMemcpyHostToDevice, h->s_main));
		const size_t o2 = offsetof(ngs_dev_state, first_tile), o3 = sizeof(ngs_dev_state);
		CK(cudaMemcpyAsync((char *)h->d_st + o2, (const char *)h->h_tmpl + o2, o3 - o2, cudaMemcpyHostToDevice, h->s_main));
	}
	if (t1 > t0) {
		CK(cudaMemsetAsync(h->d_stC + t0, 0, (t1 - t0) * 8, h->s_main));
		if (text) CK(cudaMemsetAsync(h->d_stF + t0, 0, (t1 - t0) * 8, h->s_main));
		k_tile_rows<<<(unsigned)((t1 - t0 + 255) / 256), 256, 0, h->s_main>>>(h->d_rkmax, (uint32_t)h->rows_used, h->d_tile2seg, h->d_seg, t0, t1 - t0,
		                                                                   h->d_row_a + t0, h->d_row_b + t0);
		{	/* tiles fed by very many rows are gathered by many CTAs through the far array */
			const size_t need = h->rows_used / (NGS_HEAVY_SLICE / 2) + 1024;
			if (need > h->heavy_cap) {
				CK(cudaStreamSynchronize(h->s_main));
				if (h->d_heavy) cudaFree(h->d_heavy);
				h->d_heavy = nullptr; h->heavy_cap = 0;
				CK(cudaMalloc((void **)&h->d_heavy, (need + need / 4) * sizeof(uint4))); h->heavy_cap = need + need / 4;
			}
			if (!h->d_heavy_n) CK(cudaMalloc((void **)&h->d_heavy_n, 4));
			CK(cudaMemsetAsync(h->d_heavy_n, 0, 4, h->s_main));
			k_heavy_plan<<<(unsigned)((t1 - t0 + 255) / 256), 256, 0, h->s_main>>>(h->d_row_a + t0, h->d_row_b + t0, t1 - t0, h->d_heavy, h->d_heavy_n, (unsigned int)h->heavy_cap);
			k_heavy_gather<<<(unsigned)h->sm_count * 4, 256, 0, h->s_main>>>(h->d_heavy, h->d_heavy_n, h->d_ent, h->d_rows, h->d_tile2seg, h->d_seg, t0,
			                                                               h->d_diff + t0 * (uint64_t)NGS_CT, h->d_far_flag + t0);
			h->stats.launches += 2;
		}
		CoverArgs A;
		A.n_tiles = t1 - t0; A.tile_off = t0; A.tile2seg = h->d_tile2seg; A.seg = h->d_seg;
		A.ent = h->d_ent; A.rows = h->d_rows; A.row_a = h->d_row_a + t0; A.row_b = h->d_row_b + t0;
		A.diff = h->d_diff + t0 * (uint64_t)NGS_CT; A.far_flag = h->d_far_flag + t0;
		A.stB = h->d_stC + t0; A.stF = h->d_stF + t0;
		A.prev = continues ? h->d_stC + t0 - 1 : nullptr;
		A.bins = h->d_bins; A.window = h->cfg.window; A.variant = h->cfg.mode == NGSCOV_MODE_WIG ? 1 : 0;
		A.tile_first = h->d_tile_first + t0; A.tile_last = h->d_tile_last + t0;
		A.nt.names = h->d_names; A.nt.name_off = h->d_name_off; A.nt.name8 = h->d_name8;
		A.out = text ? h->d_text[NGSCOV_OUT_BEDGRAPH] + h->bed_off : nullptr;
		A.out_cap = text ? h->d_text_cap[NGSCOV_OUT_BEDGRAPH] - h->bed_off : 0;
		A.stage_bytes = stage; A.st = h->d_st;
		A.debug = getenv("NGSCOV_DEBUG_SCAN") ? atoi(getenv("NGSCOV_DEBUG_SCAN")) : 0;
		if (text) {
			CK(cudaFuncSetAttribute(k_cover<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage));
			k_cover<true><<<(unsigned)(t1 - t0), NGS_COVER_THREADS, stage, h->s_main>>>(A);
		} else k_cover<false><<<(unsigned)(t1 - t0), NGS_COVER_THREADS, 0, h->s_main>>>(A);
		k_halo_gather<<<1, 32, 0, h->s_main>>>(h->d_st, h->d_tile_first, h->d_tile_last, h->d_halo);
		h->stats.launches += 3;
		CK(cudaGetLastError());
	}
	if ((r = small_d2h(h, h->h_st2, h->d_st, sizeof(ngs_dev_state)))) return r;
	CK(cudaEventRecord(h->ev_cov1, h->s_main));
	h->cov_pending = true; h->cov_t0 = t0; h->cov_t1 = t1;
	h->fin_tile = t1;
	return defer ? 0 : cover_harvest(h);
}

static int cover_harvest(ngscov *h)
{
	if (!h->cov_pending) return 0;
	h->cov_pending = false;
	int r;
	const bool text = h->cfg.mode == NGSCOV_MODE_DEPTH;
	const uint64_t total_tiles = h->n_tiles * NGS_CPT, t0 = h->cov_t0, t1 = h->cov_t1;
	float ms = 0;
	CK(cudaEventSynchronize(h->ev_cov1));
	cudaEventElapsedTime(&ms, h->ev_cov0, h->ev_cov1);
	h->stats.ms_scan += ms;
	const ngs_dev_state &s2 = *h->h_st2;
	if (s2.flags & NGS_FLAG_STALL) { snprintf(h->cuda_err, sizeof h->cuda_err, "look-back watchdog: a predecessor tile never published"); return h->sticky_err = NGSCOV_ECUDA; }
	h->acc_boundaries += s2.n_boundaries; h->acc_runs += s2.n_runs;
	h->mops_fin += s2.n_boundaries / 2;                         /* every interval makes at most two boundaries */
	h->acc_max = std::max(h->acc_max, s2.max_depth); h->acc_min = std::min(h->acc_min, s2.min_depth);
	h->stats.boundaries = h->acc_boundaries; h->stats.runs = h->acc_runs; h->stats.positions = h->n_elems;
	if (s2.n_runs) {
		ngscov_run f = { (int32_t)s2.halo_first[3], s2.halo_first[0], s2.halo_first[1], (int32_t)s2.halo_first[2] };
		ngscov_run l = { (int32_t)s2.halo_last[3], s2.halo_last[0], s2.halo_last[1], (int32_t)s2.halo_last[2] };
		if (!h->halo_have) { h->cov_first = f; h->halo_have = true; }
		h->cov_last = l;
	}
	if (text) {
		if (s2.fmt_bytes > h->d_text_cap[NGSCOV_OUT_BEDGRAPH] - h->bed_off) return h->sticky_err = NGSCOV_ENOMEM;
		h->bed_off += (size_t)s2.fmt_bytes;
		h->text_len[NGSCOV_OUT_BEDGRAPH] = h->bed_off;
		h->stats.out_bytes[NGSCOV_OUT_BEDGRAPH] = h->bed_off;
		const bool streamed = h->stream_fin && (t0 > 0 || t1 < total_tiles);
		if (streamed && h->bed_off > h->bed_d2h && !getenv("NGSCOV_DEBUG_NOD2H")) {
			if ((r = grow_host_text(h, NGSCOV_OUT_BEDGRAPH, h->bed_off, h->bed_d2h))) return r;
			CK(cudaMemcpyAsync(h->h_text[NGSCOV_OUT_BEDGRAPH] + h->bed_d2h, h->d_text[NGSCOV_OUT_BEDGRAPH] + h->bed_d2h, h->bed_off - h->bed_d2h,
			                   cudaMemcpyDeviceToHost, h->s_out));
			h->bed_d2h = h->bed_off;
		}
	}
	size_t ns = 0;                                              /* segments entirely below the finalized tiles */
	while (ns < h->seg.size() && (ns + 1 < h->seg.size() ? h->seg[ns + 1].tile0 : h->n_tiles) * NGS_CPT <= t1) ++ns;
	if (t1 >= total_tiles) ns = h->seg.size();
	h->next_fin_seg = ns;
	return 0;
}

/* Streamed finalize of the cover pipeline, at tile granularity.  The input is coordinate sorted (checked), so with
 * (contig kc, position kp) the largest record key verified so far every contig before kc is complete and so is every
 * position of kc below kp: no later record can add a delta there.  Finalizing whole tiles below that point while the
 * stream is still arriving lets the device -> host copy of the text overlap the input from the first chunks on, also for
 * a region shard that holds one or two contigs. */
static int stream_finalize_cover(ngscov *h)
{
	const unsigned long long key = h->h_st->key_chunk_max;
	if (key == 0 || h->seg.empty()) return 0;
	const uint32_t kc = (uint32_t)(key >> 32), kp = (uint32_t)key ? (uint32_t)key - 1u : 0u;
	uint64_t upto = h->fin_tile;
	for (size_t s = 0; s < h->seg.size(); ++s) {
		const ngs_seg &sg = h->seg[s];
		const uint64_t t_beg = sg.tile0 * NGS_CPT, t_end = (s + 1 < h->seg.size() ? h->seg[s + 1].tile0 : h->n_tiles) * NGS_CPT;
		if (kc == 0xFFFFFFFFu || sg.contig < kc || (h->end_ctg == (int32_t)kc && sg.contig == kc && kp >= h->end_pos)) upto = std::max(upto, t_end);
		else if (sg.contig == kc && kp > sg.clip_lo) upto = std::max(upto, std::min<uint64_t>(t_end, t_beg + (kp - sg.clip_lo) / NGS_CT));
	}
	const uint64_t min_tiles = getenv("NGSCOV_STREAM_MIN_TILES") ? (uint64_t)atoll(getenv("NGSCOV_STREAM_MIN_TILES")) : 2048;
	if (upto < h->fin_tile + min_tiles) return 0;               /* not worth a launch yet */
	const auto t0 = std::chrono::steady_clock::now();
	/* the interval counters lag one chunk behind the keys: allow for every record the pending chunk can hold */
	const unsigned long long mops = h->h_st->m_ops - h->mops_fin + (unsigned long long)(h->p_n / 36 + 1) * 2;
	int r = cover_tiles(h, h->fin_tile, upto, mops, true);
	h->fin_wall += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); ++h->fin_calls;
	return r;
}

/* bedGraph text of segments [s0, s1) from the records of the scan_range that just ran over them; appended at
 * bed_off.  In streamed mode the new bytes are queued for D2H at once (stream s_out). */
static int format_bed_range(ngscov *h, size_t s0, size_t s1)
{
	int r;
	const bool streamed = h->stream_fin && (h->next_fin_seg > 0 || s1 < h->seg.size() || s0 > 0);
	const ngs_dev_state st = *h->h_st2;
	BedProducer P;
	P.run_start = h->d_run_start; P.run_end = h->d_run_end; P.run_depth = h->d_run_depth;
	P.rec_base = h->d_rec_base; P.seg = h->d_seg + s0; P.n_seg = (uint32_t)(s1 - s0);
	P.nt.names = h->d_names; P.nt.name_off = h->d_name_off; P.nt.name8 = h->d_name8;
	P.drop_idx = h->drop_idx; P.mod_idx = h->mod_idx; P.mod_end = h->mod_end;
	const uint32_t dd = std::max(dig((uint64_t)std::max(st.max_depth, 0)), 1 + dig((uint64_t)(-(int64_t)std::min(st.min_depth, 0))));
	size_t bound = 0, max_line = 8;
	for (size_t sg = s0; sg < s1; ++sg) {
		const unsigned long long k = h->h_rec_base[sg - s0 + 1] - h->h_rec_base[sg - s0];
		const uint32_t pd = dig(std::max<uint64_t>(h->seg[sg].lim, h->mod_end));
		const size_t line = h->ref_name[h->seg[sg].contig].size() + 1 + 2 * pd + 2 + dd + 1;     /* longest possible line of the contig */
		bound += (size_t)k * line;
		max_line = std::max(max_line, line);
	}
	/* the staging tile is sized from the longest line that can occur (known depth range, contig lengths), not from
	 * the type limits: 29 instead of 40 bytes per line on the WGS workload lets more tiles be resident per SM */
	r = run_format(h, NGSCOV_OUT_BEDGRAPH, st.n_boundaries, P, (uint32_t)max_line, bound, h->bed_off);
	if (r) return r;
	h->bed_off = h->text_len[NGSCOV_OUT_BEDGRAPH];
	if (streamed && h->bed_off > h->bed_d2h && !getenv("NGSCOV_DEBUG_NOD2H")) {
		if ((r = grow_host_text(h, NGSCOV_OUT_BEDGRAPH, h->bed_off, h->bed_d2h))) return r;
		CK(cudaMemcpyAsync(h->h_text[NGSCOV_OUT_BEDGRAPH] + h->bed_d2h, h->d_text[NGSCOV_OUT_BEDGRAPH] + h->bed_d2h, h->bed_off - h->bed_d2h,
		                   cudaMemcpyDeviceToHost, h->s_out));
		h->bed_d2h = h->bed_off;
	}
	return 0;
}

extern "C" int ngscov_format(ngscov_t *h, uint32_t which_mask)
{
	if (!h) return NGSCOV_EINVAL;
	if (!h->runs_done) return NGSCOV_ESTATE;
	if (h->slide) { h->formatted = true; return 0; }
	CK(cudaSetDevice(h->device));
	int r = 0;
	NameTable nt = { h->d_names, h->d_name_off, h->d_name8 };
	if ((which_mask & (1u << NGSCOV_OUT_BEDGRAPH)) && h->cover && !h->red_mode) {
		h->stats.out_bytes[NGSCOV_OUT_BEDGRAPH] = h->bed_off - h->bed_begin - h->bed_tail_old + h->bed_tail.size();
	} else if ((which_mask & (1u << NGSCOV_OUT_BEDGRAPH)) && h->next_fin_seg == 0) {      /* streamed inputs already have their text */
		if ((r = format_bed_range(h, 0, h->seg.size()))) return r;
	}
	if (which_mask & (1u << NGSCOV_OUT_DEPTH)) {
		DepthProducer P;
		P.bins = h->d_bins; P.win_base = h->d_win_base; P.ref_len = h->d_ref_len; P.n_ref = (uint32_t)h->n_ref;
		P.window = h->cfg.window; P.nt = nt;
		const size_t bound = (size_t)h->n_bins * (h->max_name_tab + 10 + 1 + 10 + 1 + 24);
		r = run_format(h, NGSCOV_OUT_DEPTH, h->n_bins, P, h->max_name_tab + 10 + 1 + 10 + 1 + 24, bound);
		if (r) return r;
	}
	if (which_mask & (1u << NGSCOV_OUT_WIG)) {
		WigProducer P;
		P.bins = h->d_bins; P.win_base = h->d_win_base; P.n_ref = (uint32_t)h->n_ref; P.window = h->cfg.window;
		P.hdr = h->d_wighdr; P.hdr_off = h->d_wighdr_off;
		const uint32_t ml = std::max<uint32_t>(h->max_wighdr, 10 + 1 + 24);
		size_t bound = (size_t)h->n_bins * (10 + 1 + 24);
		for (int i = 0; i < h->n_ref; ++i) bound += h->ref_name[i].size() + 48;
		r = run_format(h, NGSCOV_OUT_WIG, h->n_bins + (unsigned long long)h->n_ref, P, ml, bound);
		if (r) return r;
	}
	if (which_mask & (1u << NGSCOV_OUT_CHROMSIZE)) {
		/* static header information only (reference bam2depth.c:337): no computed data, kept on the host */
		h->text_len[NGSCOV_OUT_CHROMSIZE] = h->chromsize.size();
		h->text_ready[NGSCOV_OUT_CHROMSIZE] = true;
		h->stats.out_bytes[NGSCOV_OUT_CHROMSIZE] = h->chromsize.size();
	}
	h->formatted = true;
	return 0;
}

extern "C" int ngscov_finish(ngscov_t *h)
{
	if (!h) return NGSCOV_EINVAL;
	int r = ngscov_finish_runs(h);
	if (r) return r;
	uint32_t mask;
	if (h->slide) return ngscov_format(h, 1u << NGSCOV_OUT_SLIDE);
	if (h->cfg.mode == NGSCOV_MODE_DEPTH) {
		mask = (1u << NGSCOV_OUT_BEDGRAPH) | (1u << NGSCOV_OUT_DEPTH);
		if (h->cfg.emit_wig) mask |= (1u << NGSCOV_OUT_WIG) | (1u << NGSCOV_OUT_CHROMSIZE);
	} else mask = (1u << NGSCOV_OUT_WIG) | (1u << NGSCOV_OUT_CHROMSIZE);
	return ngscov_format(h, mask);
}

extern "C" int ngscov_output_device(ngscov_t *h, int which, const void **dev_buf, size_t *len)
{
	if (!h || which < 0 || which > 2 || !dev_buf || !len) return NGSCOV_EINVAL;
	if (!h->text_ready[which]) return NGSCOV_ESTATE;
	{ const int rh = cover_harvest(h); if (rh) return rh; }
	if (which == NGSCOV_OUT_BEDGRAPH && h->dev_stitched) {
		CK(cudaSetDevice(h->device));
		CK(cudaStreamSynchronize(h->s_main));
		const size_t b = (size_t)h->h_patch[0], e = (size_t)h->h_patch[1];
		*dev_buf = h->d_text[which] + (b < e ? b : 0); *len = b < e ? e - b : 0;
		return 0;
	}
	if (which == NGSCOV_OUT_BEDGRAPH && h->cover && !h->red_mode && (h->bed_begin || h->bed_tail_old)) {
		/* halo decisions: skip the dropped first line, overwrite the last line in place (the buffer has slack) */
		CK(cudaSetDevice(h->device));
		const size_t body = h->bed_off - h->bed_tail_old;
		if (h->bed_begin >= h->bed_off) { *dev_buf = h->d_text[which]; *len = 0; return 0; }
		if (!h->bed_tail.empty() && !h->bed_patched) {
			CK(cudaMemcpyAsync(h->d_text[which] + body, h->bed_tail.data(), h->bed_tail.size(), cudaMemcpyHostToDevice, h->s_main));
			CK(cudaStreamSynchronize(h->s_main));
			h->bed_patched = true;
		}
		*dev_buf = h->d_text[which] + h->bed_begin; *len = body + h->bed_tail.size() - h->bed_begin;
		return 0;
	}
	*dev_buf = h->d_text[which]; *len = h->text_len[which];
	return 0;
}

extern "C" int ngscov_output(ngscov_t *h, int which, const char **buf, size_t *len)
{
	if (h && which == NGSCOV_OUT_SLIDE && buf && len) {
		if (!h->slide || !h->runs_done) return NGSCOV_ESTATE;
		*buf = h->slide_text.data(); *len = h->slide_text.size();
		return 0;
	}
	if (!h || which < 0 || which > 3 || !buf || !len) return NGSCOV_EINVAL;
	if (!h->text_ready[which]) return NGSCOV_ESTATE;
	if (which == NGSCOV_OUT_CHROMSIZE) { *buf = h->chromsize.data(); *len = h->chromsize.size(); return 0; }
	CK(cudaSetDevice(h->device));
	{ const int rh = cover_harvest(h); if (rh) return rh; }      /* a streamed finalize still in flight (only before ngscov_finish) */
	if (which == NGSCOV_OUT_BEDGRAPH && h->bed_d2h && h->bed_d2h == h->text_len[which] && !h->h_text_valid[which]) {
		tic(h);
		CK(cudaStreamSynchronize(h->s_out));               /* streamed finalize: the copy has been in flight since the scan */
		h->stats.ms_d2h += toc(h);
		h->h_text_valid[which] = true;
	}
	if (!h->h_text_valid[which]) {
		const size_t n = h->text_len[which] + ((which == NGSCOV_OUT_BEDGRAPH && h->dev_stitched && h->text_len[which]) ? 32 : 0);   /* a rewritten last line may be a few bytes longer */
		if (n > h->h_text_cap[which] || !h->h_text[which]) {
			if (h->h_text[which]) cudaFreeHost(h->h_text[which]);
			h->h_text[which] = nullptr; h->h_text_cap[which] = 0;
			const size_t want = n + n / 16 + 4096;
			CK(cudaHostAlloc((void **)&h->h_text[which], want, cudaHostAllocDefault));
			h->h_text_cap[which] = want;
		}
		tic(h);
		if (n) CK(cudaMemcpyAsync(h->h_text[which], h->d_text[which], n, cudaMemcpyDeviceToHost, h->s_main));
		h->stats.ms_d2h += toc(h);
		h->h_text_valid[which] = true;
	}
	*buf = h->h_text[which]; *len = h->text_len[which];
	if (which == NGSCOV_OUT_BEDGRAPH && h->dev_stitched) {
		CK(cudaStreamSynchronize(h->s_main));
		const size_t b = (size_t)h->h_patch[0], e = (size_t)h->h_patch[1];
		*buf = h->h_text[which] + (b < e ? b : 0); *len = b < e ? e - b : 0;
		return 0;
	}
	if (which == NGSCOV_OUT_BEDGRAPH && h->cover && !h->red_mode && (h->bed_begin || h->bed_tail_old)) {
		if (h->bed_begin >= h->bed_off) { *len = 0; return 0; }
		const size_t body = h->bed_off - h->bed_tail_old;
		if (body + h->bed_tail.size() > h->h_text_cap[which]) return NGSCOV_ENOMEM;     /* cannot happen: the buffers keep slack */
		if (!h->bed_tail.empty()) memcpy(h->h_text[which] + body, h->bed_tail.data(), h->bed_tail.size());
		*buf = h->h_text[which] + h->bed_begin; *len = body + h->bed_tail.size() - h->bed_begin;
	}
	return 0;
}

extern "C" int ngscov_runs_get(ngscov_t *h, ngscov_run *runs, uint64_t cap, uint64_t *n)
{
	if (!h || !n) return NGSCOV_EINVAL;
	if (h->runs_done && h->cover && !h->red_mode) {
		/* the cover pipeline keeps no binary run records: they are read back from the bedGraph text (DEPTH mode only) */
		if (h->cfg.mode != NGSCOV_MODE_DEPTH) return NGSCOV_ESTATE;
		const char *buf; size_t len;
		int r = ngscov_output(h, NGSCOV_OUT_BEDGRAPH, &buf, &len);
		if (r) return r;
		CK(cudaStreamSynchronize(h->s_main));
		uint64_t k = 0; size_t i = 0; int ctg = 0;
		while (i < len) {
			size_t j = i; while (j < len && buf[j] != '\t') ++j;
			if (ctg >= h->n_ref || h->ref_name[ctg].size() != j - i || memcmp(h->ref_name[ctg].data(), buf + i, j - i)) {
				for (ctg = 0; ctg < h->n_ref; ++ctg) if (h->ref_name[ctg].size() == j - i && !memcmp(h->ref_name[ctg].data(), buf + i, j - i)) break;
			}
			uint64_t v[3] = { 0, 0, 0 };
			i = j + 1;
			for (int f = 0; f < 3; ++f) { while (i < len && buf[i] >= '0' && buf[i] <= '9') v[f] = v[f] * 10 + (uint64_t)(buf[i++] - '0'); ++i; }
			if (runs && k < cap) { runs[k].contig = ctg < h->n_ref ? ctg : -1; runs[k].start = (uint32_t)v[0]; runs[k].end = (uint32_t)v[1]; runs[k].depth = (int32_t)v[2]; }
			++k;
		}
		*n = k;
		return 0;
	}
	if (!h->runs_done || h->next_fin_seg > 0) return NGSCOV_ESTATE;      /* streamed inputs keep only the last contig's records */
	CK(cudaSetDevice(h->device));
	const uint64_t nb = h->stats.boundaries;
	*n = h->stats.runs;
	if (!runs || cap == 0 || nb == 0) return 0;
	std::vector<uint32_t> s(nb), e(nb); std::vector<int> d(nb);
	CK(cudaMemcpy(s.data(), h->d_run_start, nb * 4, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(e.data(), h->d_run_end, nb * 4, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(d.data(), h->d_run_depth, nb * 4, cudaMemcpyDeviceToHost));
	uint64_t k = 0; size_t sg = 0;
	for (uint64_t i = 0; i < nb && k < cap; ++i) {
		while (sg + 1 < h->seg.size() && h->h_rec_base[sg + 1] <= i) ++sg;
		if (d[i] == 0) continue;
		runs[k].contig = (int32_t)h->seg[sg].contig; runs[k].start = s[i]; runs[k].end = e[i]; runs[k].depth = d[i];
		++k;
	}
	return 0;
}

extern "C" int ngscov_stats(ngscov_t *h, ngscov_stats_t *st)
{
	if (!h || !st) return NGSCOV_EINVAL;
	*st = h->stats;
	return 0;
}
