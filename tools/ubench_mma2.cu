// ubench_mma2.cu -- NEXT-ROUND EXPERIMENT (compiled here, never run yet: no GPU minutes were left).
// Question: does a CTA PAIR (tcgen05.mma.cta_group::2, M = 256 = 128 rows per CTA) halve the per-SM MMA issue cost that bounds
// the sparse conv?  tools/ubench_mma.cu measured ~46 clk per cta_group::1 instruction for N <= 64 whatever its size; if one
// cta_group::2 instruction (issued by the leader CTA only) also costs ~46 clk, the two SMs of a TPC get their 128-row tiles
// for half the issue time each.  Prints clk per MMA for N = 32 / 64 / 128 next to the cta_group::1 numbers of ubench_mma.
//
// RUN UNDER A SHORT TIMEOUT (first contact with cluster-scope barriers: a protocol mistake hangs the GPU):
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I detzero_b200/csrc -o tools/ubench_mma2 tools/ubench_mma2.cu -lcuda
//     timeout 20 ./tools/ubench_mma2
#include <cstdio>
#include <cstdlib>
#include "tc.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int NCOLS> __device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst) {          // one full warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(smem_dst)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS> __device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void mma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma2_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"l"((uint64_t)__cvta_generic_to_shared(bar)), "h"(cta_mask) : "memory");
}

// A: 128 rows x 32 fp32 per CTA (its half of M = 256); B: N/2 rows x 32 fp32 per CTA (its half of N); both K-major SW128
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) k_mma2(long long* out, int iters) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    constexpr int COLS = N < 32 ? 32 : N;
    for (int i = threadIdx.x; i < (16384 + (N / 2) * 128) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) { tc::mbar_init(&bar_done, 1); tc::fence_barrier_init(); }
    tc::fence_proxy_async();
    __syncthreads();
    cluster_sync_all();                                   // both CTAs initialised before anything cluster-scoped
    if (threadIdx.x < 32) tmem_alloc2<COLS>(&tmem_slot);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    cluster_sync_all();
    const uint32_t tmem = tmem_slot;
    const uint32_t rank = cluster_ctarank();
    if (rank == 0 && threadIdx.x == 0) {
        constexpr uint32_t idesc = tc::instr_desc(2, 256, N);
        const uint32_t sa = tc::smem_u32(smem);
        const uint64_t adesc = tc::smem_desc_sw128(sa), bdesc = tc::smem_desc_sw128(sa + 16384);
        long long t0 = clock64();
        for (int g = 0; g < iters; ++g) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma2_tf32(tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (g | kk) ? 1u : 0u);
        }
        long long t1 = clock64();
        mma2_commit_multicast(&bar_done, 0b11);           // arrives on bar_done of BOTH CTAs when every MMA has completed
        tc::mbar_wait(&bar_done, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else if (rank == 1 && threadIdx.x == 0) {
        tc::mbar_wait(&bar_done, 0);                      // the peer keeps its shared memory / TMEM alive until the pair is done
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x < 32) tmem_dealloc2<COLS>(tmem);
}

template <int N>
void run(long long* d_out, int sms) {
    const int iters = 2000;
    const size_t smem = 16384 + (N / 2) * 128 + 1024;
    CK(cudaFuncSetAttribute(k_mma2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(100 * 1024)));
    k_mma2<N><<<(sms / 2) * 2, 128, 100 * 1024>>>(d_out, iters);          // one CTA per SM (shared-memory footprint), pairs per TPC
    CK(cudaDeviceSynchronize());
    long long h[2]; CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
    printf("cta_group::2 M=256 N=%3d | issue %.1f clk/MMA, complete %.1f clk/MMA per PAIR (two SMs' 128-row tiles); pipe floor 128*N/(256*2) = %d\n",
           N, (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), 128 * N / 512);
    (void)smem;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    long long* d_out; CK(cudaMalloc(&d_out, 64));
    run<32>(d_out, prop.multiProcessorCount);
    run<64>(d_out, prop.multiProcessorCount);
    run<128>(d_out, prop.multiProcessorCount);
    return 0;
}
