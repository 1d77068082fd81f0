// ubench_mma.cu -- how fast can ONE thread issue tcgen05.mma (kind::tf32, M=128, K=8) as a function of N, of the
// per-4-MMA bookkeeping the conv kernels do (fence.proxy.async, tcgen05.commit, mbarrier wait) and of CTAs per SM?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I detzero_b200/csrc -o tools/ubench_mma tools/ubench_mma.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include "tc.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// instruction descriptor as in the conv kernels (tc::instr_desc is defined further down in tc.cuh)
template <int N, int MODE>      // MODE bit0: fence.proxy.async per group; bit1: commit per group; bit2: wait for the commit of group g-2 (pipelined)
__global__ void __launch_bounds__(128) k_mma(long long* out, int iters) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars[4];
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) tc::mbar_init(bars + i, 1); tc::fence_barrier_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc<(N < 32 ? 32 : N)>(&tmem_slot);
    tc::fence_proxy_async();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = tc::instr_desc(2, 128, N);
        const uint32_t sa = tc::smem_u32(smem);
        const uint64_t adesc = tc::smem_desc_sw128(sa), bdesc = tc::smem_desc_sw128(sa + 16384);
        long long t0 = clock64();
        for (int g = 0; g < iters; ++g) {
            if (MODE & 4) { if (g >= 2) tc::mbar_wait(bars + (g & 1), ((g - 2) >> 1) & 1); }
            if (MODE & 1) tc::fence_proxy_async();
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) tc::mma_tf32(tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (g | kk) ? 1u : 0u);
            if (MODE & 2) tc::mma_commit(bars + (g & 1));
        }
        long long t1 = clock64();
        tc::mma_commit(bars + 2);
        tc::mbar_wait(bars + 2, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc<(N < 32 ? 32 : N)>(tmem);
}

template <int N, int MODE>
void run(int ctas_per_sm, long long* d_out, int sms) {
    const int iters = 2000;
    size_t smem = 16384 + N * 128 + 1024;
    if (ctas_per_sm == 1) smem = 120 * 1024;
    CK(cudaFuncSetAttribute(k_mma<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_mma<N, MODE><<<sms * ctas_per_sm, 128, smem>>>(d_out, iters);
    CK(cudaDeviceSynchronize());
    long long h[2]; CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
    printf("N=%3d fence=%d commit/4=%d wait=%d ctas/SM=%d | issue %.1f clk/MMA, complete %.1f clk/MMA (floor 128*N/256 = %d)\n", N, MODE & 1, (MODE >> 1) & 1,
           (MODE >> 2) & 1, ctas_per_sm, (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), 128 * N / 256);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    long long* d_out; CK(cudaMalloc(&d_out, 64));
    const int sms = prop.multiProcessorCount;
    for (int c = 1; c <= 2; ++c) {
        run<16, 0>(c, d_out, sms); run<32, 0>(c, d_out, sms); run<64, 0>(c, d_out, sms); run<128, 0>(c, d_out, sms); run<256, 0>(c, d_out, sms);
        run<64, 2>(c, d_out, sms); run<64, 3>(c, d_out, sms); run<64, 7>(c, d_out, sms); run<64, 6>(c, d_out, sms);
        run<128, 2>(c, d_out, sms); run<128, 3>(c, d_out, sms); run<128, 7>(c, d_out, sms); run<128, 6>(c, d_out, sms);
        run<32, 7>(c, d_out, sms); run<16, 7>(c, d_out, sms);
    }
    return 0;
}
