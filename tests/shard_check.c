/* tests/shard_check.c -- CPU-only driver for host/ngscov_shard.h (TEST INFRASTRUCTURE; no GPU, no libngscov_cuda).
 *   shard_check plan  n  len0 len1 ...                      -> "beg end" per shard
 *   shard_check seek  bam  contig pos                        -> virtual offset chosen for that position
 *   shard_check stitch  n  then per shard: n_runs fc fs fe fd lc ls le ld   (first / last run)
 *                                                             -> "drop new_last_end" per shard, then "corr contig start end depth sign" lines
 */
#include "../highperformancengs_b200/host/ngscov_shard.h"

int main(int argc, char **argv)
{
	if (argc < 3) return 2;
	if (!strcmp(argv[1], "plan")) {
		const int n = atoi(argv[2]), n_ref = argc - 3;
		uint32_t *len = (uint32_t *)calloc((size_t)n_ref + 1, 4);
		for (int i = 0; i < n_ref; ++i) len[i] = (uint32_t)strtoul(argv[3 + i], NULL, 10);
		slice_t *sl = (slice_t *)calloc((size_t)n + 1, sizeof(slice_t));
		plan_slices(len, n_ref, n, sl);
		for (int g = 0; g < n; ++g) {
			int c; uint32_t p; locate(len, n_ref, sl[g].beg, &c, &p);
			printf("%llu %llu %d %u\n", (unsigned long long)sl[g].beg, (unsigned long long)sl[g].end, c, p);
		}
		return 0;
	}
	if (!strcmp(argv[1], "seek") && argc >= 5) {
		bai_t b; memset(&b, 0, sizeof b);
		if (bai_load(argv[2], &b)) { fprintf(stderr, "no index\n"); return 1; }
		printf("%d %llu\n", b.n_ref, (unsigned long long)bai_seek(&b, atoi(argv[3]), (uint32_t)strtoul(argv[4], NULL, 10)));
		return 0;
	}
	if (!strcmp(argv[1], "stitch")) {
		const int n = atoi(argv[2]);
		if (argc != 3 + 9 * n) return 2;
		ngscov_halo *h = (ngscov_halo *)calloc((size_t)n + 1, sizeof(ngscov_halo));
		for (int g = 0; g < n; ++g) {
			char **a = argv + 3 + 9 * g;
			h[g].n_runs = strtoull(a[0], NULL, 10);
			h[g].first.contig = atoi(a[1]); h[g].first.start = (uint32_t)strtoul(a[2], NULL, 10); h[g].first.end = (uint32_t)strtoul(a[3], NULL, 10); h[g].first.depth = atoi(a[4]);
			h[g].last.contig = atoi(a[5]); h[g].last.start = (uint32_t)strtoul(a[6], NULL, 10); h[g].last.end = (uint32_t)strtoul(a[7], NULL, 10); h[g].last.depth = atoi(a[8]);
		}
		int *drop = (int *)calloc((size_t)n + 1, sizeof(int)); uint32_t *nle = (uint32_t *)calloc((size_t)n + 1, 4);
		corr_t *corr = (corr_t *)calloc((size_t)3 * n + 3, sizeof(corr_t));
		const size_t nc = stitch(h, n, drop, nle, corr);
		for (int g = 0; g < n; ++g) printf("%d %u\n", drop[g], nle[g]);
		for (size_t k = 0; k < nc; ++k) printf("corr %d %u %u %d %d\n", corr[k].run.contig, corr[k].run.start, corr[k].run.end, corr[k].run.depth, corr[k].sign);
		return 0;
	}
	return 2;
}
