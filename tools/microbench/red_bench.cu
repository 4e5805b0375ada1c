// RED (no-return atomic add) throughput on B200 as a function of the active window and of per-warp locality.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 red_bench.cu -o red_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t z){ z += 0x9E3779B97F4A7C15ull; z=(z^(z>>30))*0xBF58476D1CE4E5B9ull; z=(z^(z>>27))*0x94D049BB133111EBull; return z^(z>>31);}
// mode 0: every lane random int in a moving window of `win` ints (window start advances with the thread index)
// mode 1: lanes of a warp hit consecutive-ish ints (stride 5) -> a handful of lines per warp instruction
// mode 2: like the fused walk: lane L of a warp works at base + L*lane_stride, advancing 5 ints per step
__global__ void k(int *a, uint64_t n_ints, uint64_t win, int mode, int steps, uint64_t lane_stride)
{
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
	const uint64_t warp = tid >> 5, lane = tid & 31;
	for (int s = 0; s < steps; ++s) {
		uint64_t idx;
		if (mode == 0) { const uint64_t base = (uint64_t)((double)tid / nthr * (n_ints - win)); idx = base + mix(tid * 7919 + s) % win; }
		else if (mode == 1) { idx = ((warp * steps + s) * 160 + lane * 5) % n_ints; }
		else { idx = (warp * 32 * lane_stride + lane * lane_stride + (uint64_t)s * 5) % n_ints; }
		atomicAdd(a + idx, 1);
		atomicAdd(a + (idx + 150) % n_ints, -1);
	}
}
int main()
{
	const uint64_t n_ints = 3ull << 30;   // 12 GB
	int *a; cudaMalloc(&a, n_ints * 4); cudaMemset(a, 0, n_ints * 4);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int blocks = 148 * 4 * 64, threads = 256, steps = 20;   // 9.7M threads x 20 steps x 2 REDs = 388M REDs
	struct { int mode; uint64_t win; uint64_t ls; const char *name; } cfg[] = {
		{0, 1ull << 20, 0, "random, window 4 MB"}, {0, 1ull << 24, 0, "random, window 64 MB"}, {0, 1ull << 28, 0, "random, window 1 GB"},
		{1, 0, 0, "warp-coalesced (stride 5 ints)"}, {2, 0, 50, "walk-like, lanes 50 ints apart (S=512)"},
		{2, 0, 200, "walk-like, lanes 200 ints apart (S=2048)"}, {2, 0, 800, "walk-like, lanes 800 ints apart (S=8192)"} };
	for (auto &c : cfg) {
		k<<<blocks, threads>>>(a, n_ints, c.win ? c.win : 1, c.mode, 2, c.ls);   // warm
		cudaEventRecord(e0);
		k<<<blocks, threads>>>(a, n_ints, c.win ? c.win : 1, c.mode, steps, c.ls);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		printf("%-45s %8.3f ms  %7.1f G RED/s\n", c.name, ms, (double)blocks * threads * steps * 2 / ms / 1e6);
	}
	return 0;
}
