/* include/ngscov.h -- C ABI of libngscov_cuda, the B200 (sm_100a) coverage engine that replaces the
 * depth-accumulation path of the reference tools bam2depth / bam2wig.
 *
 * The reference has no plugin/FFI layer for this path; its seams are the samtools record callback and
 * a handful of static functions.  Each entry point below names the reference code it replaces:
 *
 *   ngscov_create        DataBuf + hashtbl_create x2 + calloc(Bins)   bam2depth.c:15-19,326-330
 *                        globalArgs (window, wig)                     bam2depth.c:21-29,259-297
 *   ngscov_staging/push/push_host  bam_fetch -> bam_iter_read -> bam_read1 -> fetch_func callback
 *                        samtools bam_index.c:682-726, bam.c:191-223; bam2depth.c:86-110, bam2wig.c:84-109
 *                        (type bam_fetch_f, samtools bam.h:627) + insertInt2Hash bam2depth.c:72-84
 *   ngscov_finish_runs   union_hashed_keys hashtbl.c:275-297 + the sweep of hash2BedGraph
 *                        bam2depth.c:203-236 / bam2wig.c:202-233 + overlap bam2depth.c:132-176 /
 *                        bam2wig.c:131-175
 *   ngscov_format        fprintf sites bam2depth.c:217,233 ; output_bins :238-246 ; output_bins_wig
 *                        :248-255 (bam2wig.c:245-252) ; chromSize line :337 (bam2wig.c:322)
 *   ngscov_output        the FILE* streams of fcreat_outfile, IO_stream.h:92-97 (the caller writes the
 *                        bytes to the reference's file names)
 *   ngscov_reset         hashtbl_destroy x2 + free(Bins)              bam2depth.c:332-336
 *   NGSCOV_MODE_SLIDE    fetch_func + cal_GC + calc_winGC + output_count_GC      bam_sliding_count.c:84-165
 *   ngscov_halo_*, ngscov_bins_*   no reference counterpart (the reference is single-process): region
 *                        sharding across GPUs, SURVEY.md section 8(e)
 *
 * Conventions: every function returns 0 on success or a negative NGSCOV_E* code; nothing throws or
 * longjmps across the boundary; the library never writes files or stderr; a handle is bound to one
 * GPU and must be driven from one host thread at a time; handles are independent of each other.
 * There is NO CPU fallback: without a CUDA device ngscov_create fails with NGSCOV_ECUDA.
 */
#ifndef NGSCOV_H
#define NGSCOV_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGSCOV_ABI_VERSION 1

typedef struct ngscov ngscov_t;

enum {                       /* cfg.mode */
	NGSCOV_MODE_DEPTH = 0,   /* bam2depth: drop flag & 0x704, half-open windows, bedGraph + .depth */
	NGSCOV_MODE_WIG   = 1,   /* bam2wig:   drop flag & 0x4,   inclusive-end windows, wig + chromSize */
	NGSCOV_MODE_SLIDE = 2    /* bam_sliding_count: per window the mapped reads STARTING in it, the sum of their l_qseq and the
	                            GC content of their SEQ -> the <outfile>.txt table (reference bam_sliding_count.c:84-165; the PNG
	                            it also draws with libgd is not produced).  One GPU, whole genome. */
};

enum {                       /* ngscov_output `which` */
	NGSCOV_OUT_BEDGRAPH  = 0,
	NGSCOV_OUT_DEPTH     = 1,
	NGSCOV_OUT_WIG       = 2,
	NGSCOV_OUT_CHROMSIZE = 3,
	NGSCOV_OUT_SLIDE     = 4     /* NGSCOV_MODE_SLIDE: the read-count / GC table */
};

enum {                       /* ngscov_push flags */
	NGSCOV_PUSH_LAST          = 1,  /* last push of this input */
	NGSCOV_PUSH_TRUNCATED_OK  = 2   /* with LAST: a trailing partial record is ignored (slice reads) */
};

enum {                       /* error codes */
	NGSCOV_OK        = 0,
	NGSCOV_EINVAL    = -1,   /* bad argument / configuration */
	NGSCOV_ECUDA     = -2,   /* CUDA runtime error (ngscov_last_cuda_error gives the text) */
	NGSCOV_ENOMEM    = -3,   /* device or pinned allocation failed */
	NGSCOV_EFORMAT   = -4,   /* malformed BAM record stream (block_size < 32, record cut short, ...) */
	NGSCOV_ERANGE    = -5,   /* an alignment reaches beyond contig end + NGSCOV_OVERHANG_PAD (the
	                            reference's own behaviour is undefined there) */
	NGSCOV_ESTATE    = -6,   /* call sequence violated */
	NGSCOV_ETOOBIG   = -7,   /* record header+CIGAR larger than the carry area */
	NGSCOV_EUNSORTED = -8    /* the input is not coordinate sorted and contigs had already been finalized while streaming
	                            (unsorted input that is detected earlier is still processed, on a slower path) */
};

#define NGSCOV_OVERHANG_PAD 65536u   /* positions past a contig's end that are still tracked */

typedef struct {
	int32_t  mode;                 /* NGSCOV_MODE_* */
	uint32_t window;               /* -w, 1 .. 2^30 */
	int32_t  emit_wig;             /* bam2depth -W: also produce wig + chromSize in DEPTH mode */
	int32_t  n_ref;                /* contigs in the BAM header */
	const uint32_t *ref_len;       /* [n_ref] target_len */
	const char *const *ref_name;   /* [n_ref] target_name (NUL terminated, copied) */
	int32_t  device;               /* CUDA device ordinal */
	int32_t  reserved0;
	/* Region shard in concatenated-genome coordinates (sum of ref_len of earlier contigs + position).
	 * slice_end == 0 means the whole genome.  Deltas are clipped to the slice. */
	uint64_t slice_beg, slice_end;
	uint64_t staging_bytes;        /* capacity of each of the two pinned staging buffers; 0 = 256 MiB */
	uint32_t flags;                /* reserved, 0 */
	uint32_t reserved1;
} ngscov_cfg;

/* One maximal constant-depth run, contig coordinates, half-open. */
typedef struct {
	int32_t  contig;               /* -1: none */
	uint32_t start, end;
	int32_t  depth;
} ngscov_run;

/* Boundary information of a shard (see ngscov_halo_get / ngscov_halo_put). */
typedef struct {
	uint64_t   n_runs;             /* non-zero-depth runs in this shard */
	ngscov_run first, last;        /* valid when n_runs > 0 (equal when n_runs == 1) */
} ngscov_halo;

typedef struct {
	uint64_t records;              /* records walked */
	uint64_t kept;                 /* records that passed the flag / refID filter */
	uint64_t m_ops;                /* CIGAR M operations accumulated (each = one +1 and one -1) */
	uint64_t cigar_words;          /* sum of n_cigar over walked records */
	uint64_t stream_bytes;         /* inflated bytes pushed */
	uint64_t boundaries;           /* non-zero entries of the difference array (run records incl. gaps) */
	uint64_t runs;                 /* non-zero-depth runs */
	uint64_t positions;            /* difference-array elements scanned */
	uint64_t out_bytes[4];         /* text bytes per NGSCOV_OUT_* */
	uint64_t spec_subchunks;       /* speculative sub-chunk walks */
	uint64_t spec_repaired;        /* sub-chunks whose speculative entry was wrong and was re-walked */
	/* device time per stage, CUDA events on the engine stream, summed over launches of this input */
	float ms_zero, ms_ingest, ms_scan, ms_format, ms_h2d, ms_d2h;
	uint32_t launches;             /* kernels launched for this input */
	float ms_inflate;              /* device BGZF inflate (ngscov_push_bgzf), same clock as the stage times above */
	float ms_decode;               /* the K2 (record decode + delta scatter) part of ms_ingest; the rest is K1 framing + verification */
	float reserved2;
} ngscov_stats_t;

int  ngscov_abi_version(void);
int  ngscov_create(ngscov_t **h, const ngscov_cfg *cfg);
void ngscov_destroy(ngscov_t *h);
const char *ngscov_strerror(int code);
const char *ngscov_last_cuda_error(const ngscov_t *h);

/* The two pinned host staging buffers owned by the library (idx 0/1).  The caller inflates BGZF blocks
 * into them; the bytes are the BAM record stream that follows the header. */
int ngscov_staging(ngscov_t *h, int idx, void **pinned, size_t *cap);

/* Submit the first nbytes of staging buffer idx.  Asynchronous: returns once the copy and the kernels
 * are queued.  Records may straddle pushes.  The buffer may be refilled after ngscov_staging_wait. */
int ngscov_push(ngscov_t *h, int idx, size_t nbytes, int flags);
int ngscov_staging_wait(ngscov_t *h, int idx);

/* Submit a caller-owned host buffer (pinned memory gives full PCIe speed).  The library streams it through
 * its two device chunk buffers, overlapping the copy of one chunk with the ingest of the previous one.
 * Records may straddle chunks and calls.  Returns when every chunk has been queued. */
int ngscov_push_host(ngscov_t *h, const void *host_records, size_t nbytes, int flags);

/* BGZF blocks exactly as they lie in the BAM file (host memory, whole blocks): the library walks the block headers,
 * copies the COMPRESSED bytes to the device and inflates them there (one warp per block) into the chunk buffers the
 * ingest kernels read -- the device-side replacement of samtools-0.1.19 bgzf.c:214-238 (inflate_block) /
 * bgzf_read_block (bgzf.c:307-368).  `skip` inflated bytes at the start are dropped (the BAM header, or the in-block
 * part of a BAI virtual offset).  Records may straddle calls like with ngscov_push_host. */
int ngscov_push_bgzf(ngscov_t *h, const void *blocks, size_t nbytes, size_t skip, int flags);
/* Submit records that already live in device memory (whole records only: the range must start and
 * end on record boundaries).  Used when the inflated stream is resident in HBM. */
int ngscov_push_device(ngscov_t *h, const void *dev_records, size_t nbytes, int flags);

/* Prefix scan of the difference array, run detection, window bins.  After it the shard's runs and bins
 * exist on the device; ngscov_halo_get / ngscov_bins_get may be called. */
int ngscov_finish_runs(ngscov_t *h);

/* Shard stitching.  halo_get returns the first / last run of this shard.  halo_put tells the shard that
 * its first run was merged into the previous shard's last run (drop_first) and/or that its last run now
 * ends at new_last_end (0 = unchanged) before text is formatted. */
/* Region shards read the BAM from the BAI linear-index offset of their first 16 kb tile (samtools bam_index.c:605-615)
 * and must consume every record that starts before the end of the slice.  *done becomes 1 once a record at or beyond
 * the slice end (or of a later contig) has been decoded: the input is coordinate sorted, so the caller can stop
 * pushing and close the input with NGSCOV_PUSH_LAST | NGSCOV_PUSH_TRUNCATED_OK.  Completes the pending push first
 * (synchronises); call it only near the expected end.  Always 0 for a handle without a slice. */
int ngscov_slice_done(ngscov_t *h, int *done);
/* Streamed finalize: how many leading contigs of the header are finished by now (their runs, bins and bedGraph text
 * exist and the text is on its way to the host).  Non-blocking; used by the drivers for the reference's per-contig
 * progress line (bam2depth.c:338).  Equals n_ref after ngscov_finish_runs. */
int ngscov_contigs_done(ngscov_t *h, uint32_t *n);

/* NGSCOV_MODE_SLIDE only: the `-r chr:beg-end` mode of bam_sliding_count (bam_sliding_count.c:393-409).  Only the records
 * bam_fetch(ref, beg, end) would hand to fetch_func are counted: refID == ref, POS < end and the alignment's end > beg
 * (samtools-0.1.19 bam_index.c: bam_iter_read / is_overlap; beg is 0-based, end exclusive, as bam_parse_region yields
 * them).  Call before the first push; ref < 0 clears the region; it survives ngscov_reset.  NGSCOV_EINVAL in the other
 * modes or for a refID outside the header, NGSCOV_ESTATE once data has been pushed. */
int ngscov_set_region(ngscov_t *h, int ref, int beg, int end);
int ngscov_halo_get(ngscov_t *h, ngscov_halo *halo);
int ngscov_halo_put(ngscov_t *h, int drop_first, uint32_t new_last_end);

/* Device-side stitch (no host round trip; the cover pipeline, i.e. the default kernels).  After ngscov_finish_runs:
 *   ngscov_halo_device   device address of this shard's halo record, 9 x int64 {n_runs, first.contig, first.start,
 *                        first.end, first.depth, last.contig, last.start, last.end, last.depth}; the caller gathers the
 *                        records of all shards in rank order (NCCL all_gather over NVLink, cudaMemcpyPeerAsync, ...)
 *   ngscov_stitch_device queues ONE kernel on ngscov_stream that applies the stitch rule to this shard from the gathered
 *                        records: skips its first line / rewrites its last line in the device text and corrects its
 *                        window bins for the merged runs it owns.  Replaces ngscov_halo_get + host logic +
 *                        ngscov_halo_put + ngscov_bins_apply_run.
 *   ngscov_bins_device   device address of the window bins (uint64 x n), to be summed over shards in place
 *                        (NCCL all_reduce) before rank 0 formats the .depth / wig text. */
int ngscov_halo_device(ngscov_t *h, const void **dev_halo);
int ngscov_stitch_device(ngscov_t *h, const void *dev_all_halos, int world, int rank);
int ngscov_bins_device(ngscov_t *h, void **dev_bins, uint64_t *n);

/* Window bins (sum over bases of depth, per window, all contigs concatenated: contig c owns
 * ref_len[c]/window + 1 consecutive entries).  get copies device -> host, put host -> device. */
int ngscov_bins_count(ngscov_t *h, uint64_t *n);
int ngscov_bins_get(ngscov_t *h, uint64_t *bins, uint64_t n);
int ngscov_bins_put(ngscov_t *h, const uint64_t *bins, uint64_t n);
/* Contribution of one run to the bins with the handle's mode/window (host-side arithmetic used by the
 * shard stitcher): bins[k] += sign * contribution. */
int ngscov_bins_apply_run(ngscov_t *h, const ngscov_run *run, int sign, uint64_t *bins, uint64_t n);

/* Device-side formatting of every requested text output.  which_mask: bit NGSCOV_OUT_* set = produce. */
int ngscov_format(ngscov_t *h, uint32_t which_mask);

/* finish = finish_runs + format(all outputs of the mode) for the unsharded case. */
int ngscov_finish(ngscov_t *h);

/* Formatted text in pinned host memory (copied on first request), valid until reset/destroy. */
int ngscov_output(ngscov_t *h, int which, const char **buf, size_t *len);
/* Same text, device pointer (no copy). */
int ngscov_output_device(ngscov_t *h, int which, const void **dev_buf, size_t *len);
/* The shard's non-zero-depth runs, copied to caller memory (cap entries); *n = total. */
int ngscov_runs_get(ngscov_t *h, ngscov_run *runs, uint64_t cap, uint64_t *n);

int ngscov_stats(ngscov_t *h, ngscov_stats_t *st);
/* Forget the current input, keep every allocation (next BAM / next benchmark step). */
int ngscov_reset(ngscov_t *h);
/* Block until all queued work of the handle is done. */
int ngscov_sync(ngscov_t *h);
/* The CUDA stream (cudaStream_t) the engine kernels are launched on, for external event timing. */
void *ngscov_stream(ngscov_t *h);

#ifdef __cplusplus
}
#endif
#endif
