// Developer micro-benchmark (GPU box): cost of one grid-wide barrier of the persistent solver, several variants.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_bench barrier_bench.cu && ./barrier_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void bar_v0(unsigned *bar, unsigned nblocks)
{
	__syncthreads();
	if(threadIdx.x == 0){
		__threadfence();
		unsigned gen = *((volatile unsigned *)&bar[1]);
		unsigned arrived = atomicAdd(&bar[0], 1u) + 1u;
		if(arrived == nblocks){ bar[0] = 0u; __threadfence(); atomicAdd(&bar[1], 1u); }
		else { while(*((volatile unsigned *)&bar[1]) == gen){ } }
		__threadfence();
	}
	__syncthreads();
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p){ unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p){ unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned atom_add_acq_rel(unsigned *p, unsigned v){ unsigned o; asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory"); return o; }
__device__ __forceinline__ void st_release(unsigned *p, unsigned v){ asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

// monotonically increasing counter: arrival k of generation g brings the count to g*nblocks + k; no reset, no generation word
__device__ __forceinline__ void bar_v1(unsigned *bar, unsigned nblocks, unsigned &target)
{
	__syncthreads();
	if(threadIdx.x == 0){
		target += nblocks;
		atom_add_acq_rel(&bar[0], 1u);
		while((int)(ld_acquire(&bar[0]) - target) < 0){ }
	}
	__syncthreads();
}

// two levels: groups of 16 CTAs count on their own word (different cache lines), the last of a group bumps the top word
__device__ __forceinline__ void bar_v2(unsigned *bar, unsigned nblocks, unsigned &gen)
{
	__syncthreads();
	if(threadIdx.x == 0){
		gen++;
		const unsigned group = blockIdx.x >> 4, ngroups = (nblocks + 15) >> 4;
		const unsigned gsize = min(16u, nblocks - (group << 4));
		unsigned *gw = bar + 32 + group*32;
		if(atom_add_acq_rel(gw, 1u) + 1u == gen*gsize){
			if(atom_add_acq_rel(&bar[0], 1u) + 1u == gen*ngroups) st_release(&bar[1], gen);
		}
		while((int)(ld_acquire(&bar[1]) - gen) < 0){ }
	}
	__syncthreads();
}

template<int V> __global__ void __launch_bounds__(256, 2) k_bar(unsigned *bar, int n, double *sink)
{
	unsigned target = 0, gen = 0;
	double acc = threadIdx.x;
	for(int i = 0; i < n; i++){
		if(V == 0) bar_v0(bar, gridDim.x);
		if(V == 1) bar_v1(bar, gridDim.x, target);
		if(V == 2) bar_v2(bar, gridDim.x, gen);
		acc = acc*1.0000001 + 1.0;
	}
	if(acc == 12345.678) sink[0] = acc;
}

template<int V> static void run(int blocks, int n)
{
	unsigned *bar; double *sink;
	cudaMalloc(&bar, 4*32*64); cudaMemset(bar, 0, 4*32*64); cudaMalloc(&sink, 8);
	void *args[] = {&bar, &n, &sink};
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for(int rep = 0; rep < 3; rep++){
		cudaMemset(bar, 0, 4*32*64);
		cudaEventRecord(e0);
		cudaError_t e = cudaLaunchCooperativeKernel((void *)k_bar<V>, dim3(blocks), dim3(256), args, 0, 0);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
		if(rep == 2) printf("variant %d  blocks %3d  %d barriers: %.3f us each  (%s)\n", V, blocks, n, 1000.0*ms/n, cudaGetErrorString(e));
	}
	cudaFree(bar); cudaFree(sink);
}

int main()
{
	for(int blocks : {148, 296}){
		run<0>(blocks, 2000); run<1>(blocks, 2000); run<2>(blocks, 2000);
	}
	return 0;
}
