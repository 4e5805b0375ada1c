// Prototype of the "tile owner" delta accumulation (K2 redesign): a CTA owns a window of T positions, gathers the
// deltas of its records from a compact 8-byte-per-record table in shared memory and writes the tile with plain
// 128-bit stores.  Measures shared-memory atomic throughput against the plain-store lower bound.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 owner_bench.cu -o owner_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define T 4096
#define H 512
#define SLOTS 256
__device__ __forceinline__ uint32_t mixh(uint32_t z){ z ^= z >> 16; z *= 0x7feb352du; z ^= z >> 15; z *= 0x846ca68bu; z ^= z >> 16; return z; }
// records: i -> pos = i*5 + jitter(0..4); rows of `per_row` records in SLOTS-slot rows
__global__ void k_fill(uint2 *ent, uint32_t *count, uint32_t *fk, uint64_t n_rec, uint32_t per_row)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_rec) return;
	const uint64_t row = i / per_row; const uint32_t k = (uint32_t)(i % per_row);
	const uint32_t pos = (uint32_t)(i * 5 + mixh((uint32_t)i) % 5);
	ent[row * SLOTS + k] = make_uint2(pos, 150u);
	if (k == 0) { fk[row] = pos; count[row] = (uint32_t)min((uint64_t)per_row, n_rec - i); }
}
__global__ void k_index(const uint32_t *fk, uint32_t n_rows, uint32_t *row0, uint32_t n_tiles)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_tiles) return;
	const int64_t thr = (int64_t)t * T - H;
	uint32_t lo = 0, hi = n_rows;                 // last row with fk < thr, or 0
	while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if ((int64_t)fk[mid] < thr) lo = mid; else hi = mid; }
	row0[t] = lo;
}
template <int MODE>   // 0: smem atomics for both deltas; 1: no accumulation at all (bound); 2: starts by plain store, ends atomic
__global__ void __launch_bounds__(256, 8) k_owner(const uint2 *ent, const uint32_t *count, const uint32_t *fk, const uint32_t *row0,
                                                   uint32_t n_rows, int *diff, int4 *agg, uint32_t n_tiles)
{
	__shared__ __align__(16) int s_acc[T];
	__shared__ int s_red[8][3];
	const int t = threadIdx.x;
	const uint32_t tile = blockIdx.x;
	#pragma unroll
	for (int j = 0; j < T / 4 / 256; ++j) reinterpret_cast<int4 *>(s_acc)[j * 256 + t] = make_int4(0, 0, 0, 0);
	__syncthreads();
	const int64_t p0 = (int64_t)tile * T, p1 = p0 + T;
	uint32_t sink = 0;
	for (uint32_t row = row0[tile]; row < n_rows && (int64_t)fk[row] < p1; ++row) {
		const uint32_t c = count[row];
		if ((uint32_t)t < c) {
			const uint2 e = ent[(size_t)row * SLOTS + t];
			const int64_t s = e.x;
			if (s >= p0 - H && s < p1) {
				const int64_t a = s - p0, b = s + e.y - p0;
				if (MODE == 0) {
					if (a >= 0) atomicAdd(&s_acc[a], 1);
					if (b >= 0 && b < T) atomicAdd(&s_acc[b], -1);
				} else if (MODE == 2) {
					if (a >= 0) s_acc[a] = 1;
				} else sink += (uint32_t)(a ^ b);
			}
		}
		if (MODE == 2) {
			__syncthreads();
			if ((uint32_t)t < c) {
				const uint2 e = ent[(size_t)row * SLOTS + t];
				const int64_t s = e.x;
				if (s >= p0 - H && s < p1) { const int64_t b = s + e.y - p0; if (b >= 0 && b < T) atomicAdd(&s_acc[b], -1); }
			}
		}
	}
	if (MODE == 1 && sink == 0x12345678u) s_acc[t] = 1;
	__syncthreads();
	int sum = 0, nz = 0, last = 0;
	int4 *g = reinterpret_cast<int4 *>(diff + (size_t)tile * T);
	#pragma unroll
	for (int j = 0; j < T / 4 / 256; ++j) {
		const int q = j * 256 + t;
		const int4 v = reinterpret_cast<const int4 *>(s_acc)[q];
		g[q] = v;
		sum += v.x + v.y + v.z + v.w;
		nz += (v.x != 0) + (v.y != 0) + (v.z != 0) + (v.w != 0);
		if (v.x) last = max(last, 4 * q + 1); if (v.y) last = max(last, 4 * q + 2); if (v.z) last = max(last, 4 * q + 3); if (v.w) last = max(last, 4 * q + 4);
	}
	sum = __reduce_add_sync(0xffffffffu, sum); nz = __reduce_add_sync(0xffffffffu, nz); last = __reduce_max_sync(0xffffffffu, last);
	if ((t & 31) == 0) { s_red[t >> 5][0] = sum; s_red[t >> 5][1] = nz; s_red[t >> 5][2] = last; }
	__syncthreads();
	if (t == 0) {
		int a = 0, b = 0, c = 0;
		for (int w = 0; w < 8; ++w) { a += s_red[w][0]; b += s_red[w][1]; c = max(c, s_red[w][2]); }
		agg[tile] = make_int4(a, b, c, 0);
	}
}
// raw shared-memory atomic throughput: every lane its own bank (stride 5 words), many iterations
__global__ void __launch_bounds__(256) k_atoms(int *out, int iters, int mode)
{
	__shared__ int s[T];
	for (int i = threadIdx.x; i < T; i += 256) s[i] = 0;
	__syncthreads();
	uint32_t a = threadIdx.x * 5u;
	for (int i = 0; i < iters; ++i) {
		if (mode == 0) atomicAdd(&s[a & (T - 1)], 1);
		else if (mode == 1) s[a & (T - 1)] += 1;
		else atomicAdd(&s[(a * 32u) & (T - 1)], 1);      // all lanes of a warp in one bank
		a += 1283u;
	}
	__syncthreads();
	if (threadIdx.x == 0) out[blockIdx.x] = s[5];
}
int main(int argc, char **argv)
{
	const double scale = argc > 1 ? atof(argv[1]) : 0.25;
	const uint64_t n_rec = (uint64_t)(618e6 * scale);
	const uint32_t per_row = 160;
	const uint32_t n_rows = (uint32_t)((n_rec + per_row - 1) / per_row);
	const uint32_t n_tiles = (uint32_t)((n_rec * 5 + 200 + T - 1) / T);
	uint2 *ent; uint32_t *count, *fk, *row0; int *diff; int4 *agg; int *out;
	cudaMalloc(&ent, (size_t)n_rows * SLOTS * 8); cudaMalloc(&count, n_rows * 4); cudaMalloc(&fk, n_rows * 4);
	cudaMalloc(&row0, n_tiles * 4); cudaMalloc(&diff, (size_t)n_tiles * T * 4); cudaMalloc(&agg, (size_t)n_tiles * 16); cudaMalloc(&out, 1 << 20);
	k_fill<<<(unsigned)((n_rec + 255) / 256), 256>>>(ent, count, fk, n_rec, per_row);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	float ms;
	cudaEventRecord(e0);
	k_index<<<(n_tiles + 255) / 256, 256>>>(fk, n_rows, row0, n_tiles);
	cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
	printf("records %llu rows %u tiles %u; index build %.3f ms\n", (unsigned long long)n_rec, n_rows, n_tiles, ms);
	for (int rep = 0; rep < 2; ++rep) {
		cudaEventRecord(e0); k_owner<0><<<n_tiles, 256>>>(ent, count, fk, row0, n_rows, diff, agg, n_tiles); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
		printf("owner, smem atomics (+1 and -1)        %8.3f ms  (%.1f GB table + %.1f GB tile -> %.0f GB/s)\n", ms, n_rec * 8e-9, n_tiles * (double)T * 4e-9, (n_rec * 8.0 + n_tiles * (double)T * 4) / ms / 1e6);
		cudaEventRecord(e0); k_owner<1><<<n_tiles, 256>>>(ent, count, fk, row0, n_rows, diff, agg, n_tiles); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
		printf("owner, no accumulation (bound)         %8.3f ms\n", ms);
		cudaEventRecord(e0); k_owner<2><<<n_tiles, 256>>>(ent, count, fk, row0, n_rows, diff, agg, n_tiles); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
		printf("owner, starts plain store, ends atomic %8.3f ms\n", ms);
	}
	const char *nm[3] = { "ATOMS, one bank per lane", "LDS+STS (non-atomic), one bank per lane", "ATOMS, all lanes one bank" };
	for (int mode = 0; mode < 3; ++mode) {
		const int iters = 4096, blocks = 148 * 8;
		k_atoms<<<blocks, 256>>>(out, 16, mode);
		cudaEventRecord(e0); k_atoms<<<blocks, 256>>>(out, iters, mode); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
		printf("%-42s %8.3f ms  %.2f lane-ops/clk/SM (at 1.965 GHz)\n", nm[mode], ms, (double)blocks * 256 * iters / 148 / (ms * 1e-3 * 1.965e9));
	}
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
