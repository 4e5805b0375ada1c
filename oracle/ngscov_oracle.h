/* oracle/ngscov_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference bam2depth / bam2wig coverage path
 * (reference bam2depth.c:86-110,132-176,203-255 and bam2wig.c:84-109,131-175,202-252).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may call this; the
 * product path (libngscov_cuda + drivers) never links or imports it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors for this path, so the oracle is pinned
 * against outputs of the unmodified reference binaries built by oracle/Makefile into oracle/_ref/
 * (tests/test_oracle_vs_reference.py runs both here; tests/golden/ holds committed outputs of the
 * reference for the GPU box where /root/reference is absent).
 */
#ifndef NGSCOV_ORACLE_H
#define NGSCOV_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORA_MODE_DEPTH = 0, ORA_MODE_WIG = 1 };

typedef struct {
	char *bedgraph;  size_t bedgraph_len;   /* <basename(bam)>.N.bedGraph (depth mode only) */
	char *depth;     size_t depth_len;      /* <PFX>.N.depth            (depth mode only) */
	char *wig;       size_t wig_len;        /* <PFX>.N.wig              (wig mode, or depth mode with -W) */
	char *chromsize; size_t chromsize_len;  /* <PFX>.N.chromSize.txt    (same condition as wig) */
	uint64_t n_records;                     /* records seen in the stream */
	uint64_t n_kept;                        /* records that passed the flag/tid filter */
	uint64_t n_mops;                        /* CIGAR M operations counted */
	uint64_t cigar_words;                   /* sum of n_cigar over all records */
} ora_result;

/* Run the restated path over an inflated BAM record stream (the bytes that follow the BAM header:
 * repeated {int32 block_size; 32-byte core; qname; cigar; seq; qual; aux}).
 * mode: ORA_MODE_DEPTH (flag mask 0x704, half-open windows) or ORA_MODE_WIG (mask 0x4, inclusive windows).
 * emit_wig: bam2depth's -W.  Returns 0, or negative on malformed input. */
int ngscov_oracle_run(const uint8_t *rec, size_t nbytes, int n_ref, const uint32_t *ref_len,
                      const char *const *ref_name, int mode, uint32_t window, int emit_wig,
                      ora_result *out);
void ngscov_oracle_free(ora_result *r);

/* Inflate a whole BGZF/BAM file into memory and parse the header.  *records points INTO *inflated. */
typedef struct {
	uint8_t *inflated; size_t inflated_len;
	const uint8_t *records; size_t records_len;
	int n_ref; uint32_t *ref_len; char **ref_name;
} ora_bam;
int ngscov_oracle_load_bam(const char *path, ora_bam *bam);
void ngscov_oracle_free_bam(ora_bam *bam);

#ifdef __cplusplus
}
#endif
#endif
