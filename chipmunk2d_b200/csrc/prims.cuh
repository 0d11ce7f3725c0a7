// prims.cuh -- device-wide primitives written for this engine: exclusive scan and a
// stable LSD radix sort of (u64 key, int value) pairs (8-bit digits, per-block
// histograms + warp match ranking).  Used by the Morton-code LBVH broadphase (K3).
// In the CPB_EMU developer build they are replaced by trivial host loops.
#pragma once
#include "cpb_rt.h"

#define CPB_SCAN_BLOCK 256
#define CPB_SCAN_ITEMS 4
#define CPB_SCAN_TILE (CPB_SCAN_BLOCK*CPB_SCAN_ITEMS)

#define CPB_SORT_BLOCK 256
#define CPB_SORT_ITEMS 8
#define CPB_SORT_TILE (CPB_SORT_BLOCK*CPB_SORT_ITEMS)

#ifndef CPB_EMU

// ---------------------------------------------------------------- scan
// Block-level exclusive scan of a tile; writes the tile total to sums[blockIdx.x].
__global__ void __launch_bounds__(CPB_SCAN_BLOCK) k_scan_tiles(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t *__restrict__ sums, int n)
{
	__shared__ uint32_t warp_tot[CPB_SCAN_BLOCK/32];
	int base = blockIdx.x*CPB_SCAN_TILE + threadIdx.x*CPB_SCAN_ITEMS;
	uint32_t v[CPB_SCAN_ITEMS];
	uint32_t tsum = 0;
#pragma unroll
	for(int k = 0; k < CPB_SCAN_ITEMS; k++){
		int i = base + k;
		v[k] = (i < n ? in[i] : 0u);
		tsum += v[k];
	}
	// inclusive warp scan of thread sums
	uint32_t inc = tsum;
	int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for(int d = 1; d < 32; d <<= 1){
		uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
		if(lane >= d) inc += t;
	}
	if(lane == 31) warp_tot[warp] = inc;
	__syncthreads();
	if(warp == 0){
		uint32_t w = (lane < CPB_SCAN_BLOCK/32 ? warp_tot[lane] : 0u);
		uint32_t winc = w;
#pragma unroll
		for(int d = 1; d < 32; d <<= 1){
			uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
			if(lane >= d) winc += t;
		}
		if(lane < CPB_SCAN_BLOCK/32) warp_tot[lane] = winc - w; // exclusive
		if(lane == CPB_SCAN_BLOCK/32 - 1 && sums) sums[blockIdx.x] = winc;
	}
	__syncthreads();
	uint32_t run = warp_tot[warp] + inc - tsum;
#pragma unroll
	for(int k = 0; k < CPB_SCAN_ITEMS; k++){
		int i = base + k;
		if(i < n) out[i] = run;
		run += v[k];
	}
}

__global__ void __launch_bounds__(CPB_SCAN_BLOCK) k_scan_add(uint32_t *__restrict__ out, const uint32_t *__restrict__ offs, int n)
{
	int base = blockIdx.x*CPB_SCAN_TILE + threadIdx.x*CPB_SCAN_ITEMS;
	uint32_t o = offs[blockIdx.x];
#pragma unroll
	for(int k = 0; k < CPB_SCAN_ITEMS; k++){
		int i = base + k;
		if(i < n) out[i] += o;
	}
}

// tmp must hold at least cpb_scan_tmp_elems(n) uint32_t.  in may alias out.
static inline size_t cpb_scan_tmp_elems(int n)
{
	size_t total = 0;
	while(n > CPB_SCAN_TILE){ n = cpb_div_up(n, CPB_SCAN_TILE); total += (size_t)n; }
	return total + 1;
}

static void cpb_exclusive_scan(const uint32_t *in, uint32_t *out, int n, uint32_t *tmp, cudaStream_t st)
{
	if(n <= 0) return;
	int blocks = cpb_div_up(n, CPB_SCAN_TILE);
	if(blocks == 1){
		LAUNCH(k_scan_tiles, 1, CPB_SCAN_BLOCK, st, in, out, (uint32_t *)NULL, n);
		return;
	}
	LAUNCH(k_scan_tiles, blocks, CPB_SCAN_BLOCK, st, in, out, tmp, n);
	cpb_exclusive_scan(tmp, tmp, blocks, tmp + blocks, st);
	LAUNCH(k_scan_add, blocks, CPB_SCAN_BLOCK, st, out, (const uint32_t *)tmp, n);
}

// ---------------------------------------------------------------- radix sort
// hist[d*nblocks + block] = number of keys of the block's tile whose digit is d
__global__ void __launch_bounds__(CPB_SORT_BLOCK) k_sort_hist(const uint64_t *__restrict__ keys, uint32_t *__restrict__ hist, int n, int shift, int nblocks)
{
	__shared__ uint32_t h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	int base = blockIdx.x*CPB_SORT_TILE;
#pragma unroll
	for(int k = 0; k < CPB_SORT_ITEMS; k++){
		int i = base + k*CPB_SORT_BLOCK + threadIdx.x;
		if(i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
	}
	__syncthreads();
	hist[threadIdx.x*nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(CPB_SORT_BLOCK) k_sort_scatter(
	const uint64_t *__restrict__ keys_in, const int *__restrict__ vals_in,
	uint64_t *__restrict__ keys_out, int *__restrict__ vals_out,
	const uint32_t *__restrict__ offs, int n, int shift, int nblocks)
{
	__shared__ uint32_t warp_cnt[CPB_SORT_BLOCK/32][256];
	__shared__ uint32_t basep[256];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	basep[threadIdx.x] = offs[threadIdx.x*nblocks + blockIdx.x];
	int tile = blockIdx.x*CPB_SORT_TILE;
	for(int k = 0; k < CPB_SORT_ITEMS; k++){
#pragma unroll
		for(int w = 0; w < CPB_SORT_BLOCK/32; w++) warp_cnt[w][threadIdx.x] = 0;
		__syncthreads();
		int i = tile + k*CPB_SORT_BLOCK + threadIdx.x;
		bool valid = (i < n);
		uint64_t key = valid ? keys_in[i] : 0ull;
		int val = valid ? vals_in[i] : 0;
		uint32_t d = valid ? ((uint32_t)(key >> shift) & 255u) : 256u;
		uint32_t peers = __match_any_sync(0xffffffffu, d);
		uint32_t rank = __popc(peers & ((1u << lane) - 1u));
		if(valid && rank == 0) warp_cnt[warp][d] = __popc(peers);
		__syncthreads();
		{
			// thread d: turn per-warp counts of digit d into running offsets
			uint32_t run = basep[threadIdx.x];
#pragma unroll
			for(int w = 0; w < CPB_SORT_BLOCK/32; w++){
				uint32_t c = warp_cnt[w][threadIdx.x];
				warp_cnt[w][threadIdx.x] = run;
				run += c;
			}
			basep[threadIdx.x] = run;
		}
		__syncthreads();
		if(valid){
			uint32_t pos = warp_cnt[warp][d] + rank;
			keys_out[pos] = key;
			vals_out[pos] = val;
		}
		__syncthreads();
	}
}

static inline size_t cpb_sort_tmp_elems(int n)
{
	int nblocks = cpb_div_up(n > 0 ? n : 1, CPB_SORT_TILE);
	return (size_t)256*nblocks + cpb_scan_tmp_elems(256*nblocks);
}

// Sorts (keys, vals) by bits [first_bit, bits) of the key (first_bit a multiple of 8; lower bits must be equal); ping-pongs between the two buffer
// pairs and returns 0 if the result is in (keys_a, vals_a), 1 if in (keys_b, vals_b).
static int cpb_radix_sort(uint64_t *keys_a, int *vals_a, uint64_t *keys_b, int *vals_b, int n, int bits, uint32_t *tmp, cudaStream_t st, int first_bit = 0)
{
	if(n <= 1) return 0;
	int nblocks = cpb_div_up(n, CPB_SORT_TILE);
	uint32_t *hist = tmp;
	uint32_t *scan_tmp = tmp + (size_t)256*nblocks;
	int cur = 0;
	for(int shift = first_bit; shift < bits; shift += 8){
		const uint64_t *ki = cur ? keys_b : keys_a; const int *vi = cur ? vals_b : vals_a;
		uint64_t *ko = cur ? keys_a : keys_b; int *vo = cur ? vals_a : vals_b;
		LAUNCH(k_sort_hist, nblocks, CPB_SORT_BLOCK, st, ki, hist, n, shift, nblocks);
		cpb_exclusive_scan(hist, hist, 256*nblocks, scan_tmp, st);
		LAUNCH(k_sort_scatter, nblocks, CPB_SORT_BLOCK, st, ki, vi, ko, vo, (const uint32_t *)hist, n, shift, nblocks);
		cur ^= 1;
	}
	return cur;
}

// warp-aggregated append: returns this thread's slot in a list guarded by *counter
// (one atomic per warp; the ballot compaction named by the north star for pair emission).
__device__ __forceinline__ int cpb_warp_append(int *counter, bool pred)
{
	unsigned active = __activemask();
	unsigned votes = __ballot_sync(active, pred);
	if(!pred) return -1;
	int lane = threadIdx.x & 31;
	int leader = __ffs(votes) - 1;
	int base = 0;
	if(lane == leader) base = atomicAdd(counter, __popc(votes));
	base = __shfl_sync(votes, base, leader);
	return base + __popc(votes & ((1u << lane) - 1u));
}

#else
// ---------------------------------------------------------------- emulation
static inline size_t cpb_scan_tmp_elems(int){ return 1; }
static inline size_t cpb_sort_tmp_elems(int){ return 1; }
static void cpb_exclusive_scan(const uint32_t *in, uint32_t *out, int n, uint32_t *, cudaStream_t)
{
	uint32_t run = 0;
	for(int i = 0; i < n; i++){ uint32_t v = in[i]; out[i] = run; run += v; }
}
static int cpb_radix_sort(uint64_t *keys_a, int *vals_a, uint64_t *, int *, int n, int bits, uint32_t *, cudaStream_t, int = 0)
{
	uint64_t mask = (bits >= 64 ? ~0ull : ((1ull << bits) - 1ull));
	int *idx = (int *)malloc(sizeof(int)*(size_t)(n > 0 ? n : 1));
	for(int i = 0; i < n; i++) idx[i] = i;
	std::stable_sort(idx, idx + n, [&](int x, int y){ return (keys_a[x] & mask) < (keys_a[y] & mask); });
	uint64_t *k2 = (uint64_t *)malloc(8*(size_t)(n > 0 ? n : 1)); int *v2_ = (int *)malloc(4*(size_t)(n > 0 ? n : 1));
	for(int i = 0; i < n; i++){ k2[i] = keys_a[idx[i]]; v2_[i] = vals_a[idx[i]]; }
	memcpy(keys_a, k2, 8*(size_t)n); memcpy(vals_a, v2_, 4*(size_t)n);
	free(idx); free(k2); free(v2_);
	return 0;
}
static inline int cpb_warp_append(int *counter, bool pred)
{
	if(!pred) return -1;
	return atomicAdd(counter, 1);
}
#endif
