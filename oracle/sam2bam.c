/* oracle/sam2bam.c -- TEST INFRASTRUCTURE ONLY.
 * Our own ~40-line helper linked against the reference's vendored samtools-0.1.19 libbam
 * (built into oracle/_ref/ by oracle/Makefile): converts a SAM text file (with @SQ header) to BAM
 * and builds the .bai with samtools' own indexer, so tests can (a) feed hand-written SAM cases to the
 * reference bam2depth/bam2wig and (b) cross-check our BAM/BAI writer against samtools'.
 *   sam2bam in.sam out.bam        -> out.bam + out.bam.bai
 *   sam2bam --index-only x.bam    -> x.bam.bai
 */
#include <stdio.h>
#include <string.h>
#include "sam.h"

int main(int argc, char **argv)
{
	if (argc == 3 && !strcmp(argv[1], "--index-only"))
		return bam_index_build(argv[2]) == 0 ? 0 : 1;
	if (argc != 3) { fprintf(stderr, "usage: sam2bam in.sam out.bam | --index-only x.bam\n"); return 2; }
	samfile_t *in = samopen(argv[1], "r", 0);
	if (!in) { fprintf(stderr, "sam2bam: cannot open %s\n", argv[1]); return 1; }
	samfile_t *out = samopen(argv[2], "wb", in->header);
	if (!out) { fprintf(stderr, "sam2bam: cannot create %s\n", argv[2]); return 1; }
	bam1_t *b = bam_init1();
	long n = 0;
	while (samread(in, b) >= 0) { samwrite(out, b); ++n; }
	bam_destroy1(b);
	samclose(out);
	samclose(in);
	if (bam_index_build(argv[2]) != 0) return 1;
	fprintf(stderr, "sam2bam: %ld records\n", n);
	return 0;
}
