// ubench_halo.cu -- can tcgen05.mma (kind::tf32, SW128 K-major A operand) read a SHIFTED WINDOW of a halo tile in shared memory?
// A 3x3 stride-1 conv as implicit GEMM re-reads its 128-pixel A tile once per tap; if the A descriptor may start at any 128-byte pixel
// of a (rows+2) x (cols+2) halo tile and step between 8-pixel row groups with SBO = halo row pitch, one halo load serves all 9 taps.
// The test: B = 32x32 identity, halo value = pixel index (even channels) / channel (odd channels), so D[m][n] names the pixel row m read.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I detzero_b200/csrc -o tools/ubench_halo tools/ubench_halo.cu
#include <cstdio>
#include <cstdlib>
#include "tc.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(128) k_halo(float* out, int pitch, int rows, int dy, int dx, int base_mode) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int npix = pitch * rows;
    unsigned char* bsm = smem + ((npix * 128 + 1023) & ~1023);
    for (int i = threadIdx.x; i < npix * 32; i += blockDim.x) {
        const int p = i >> 5, c = i & 31;
        const float v = (c & 1) ? (float)c : (float)p;
        *reinterpret_cast<float*>(smem + p * 128 + (((c >> 2) ^ (p & 7)) << 4) + (c & 3) * 4) = v;
    }
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
        const int n = i >> 5, c = i & 31;
        *reinterpret_cast<float*>(bsm + n * 128 + (((c >> 2) ^ (n & 7)) << 4) + (c & 3) * 4) = (n == c) ? 1.f : 0.f;
    }
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc<32>(&tmem_slot);
    tc::fence_proxy_async();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = tc::instr_desc(2, 128, 32);
        const uint32_t a0 = tc::smem_u32(smem) + (uint32_t)((dy * pitch + dx) * 128);
        const uint32_t boff = base_mode ? ((a0 >> 7) & 7u) : 0u;
        const uint64_t adesc = (uint64_t)((a0 >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((pitch * 128) >> 4) << 32) | ((uint64_t)1 << 46) |
                               ((uint64_t)boff << 49) | ((uint64_t)2 << 61);
        const uint64_t bdesc = tc::smem_desc_sw128(tc::smem_u32(bsm));
        for (int kk = 0; kk < 4; ++kk) tc::mma_tf32(tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, kk ? 1u : 0u);
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
    }
    __syncthreads();
    tc::tcgen05_fence_after();
    float v[32];
    const int warp = threadIdx.x >> 5;
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
    for (int n = 0; n < 32; ++n) out[threadIdx.x * 32 + n] = v[n];
    tc::tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc<32>(tmem);
}

int main() {
    float* d; CK(cudaMalloc(&d, 128 * 32 * 4));
    float h[128 * 32];
    CK(cudaFuncSetAttribute(k_halo, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    const int pitches[3] = {10, 16, 8};
    for (int pi = 0; pi < 3; ++pi)
        for (int base_mode = 0; base_mode < 2; ++base_mode)
            for (int dy = 0; dy < 3; dy += 1)
                for (int dx = 0; dx < 3; ++dx) {
                    const int pitch = pitches[pi], rows = 18;
                    if (pitch == 8 && dx) continue;
                    k_halo<<<1, 128, 100 * 1024>>>(d, pitch, rows, dy, dx, base_mode);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("pitch %d base_mode %d dy %d dx %d: CUDA error %s\n", pitch, base_mode, dy, dx, cudaGetErrorString(e)); return 1; }
                    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
                    int bad_pix = 0, bad_ch = 0, first = -1;
                    for (int m = 0; m < 128; ++m) {
                        const int y = m >> 3, x = m & 7, want = (y + dy) * pitch + x + dx;
                        for (int n = 0; n < 32; ++n) {
                            const float w = (n & 1) ? (float)n : (float)want;
                            if (h[m * 32 + n] != w) { if (n & 1) ++bad_ch; else ++bad_pix; if (first < 0) first = m; }
                        }
                    }
                    printf("pitch %2d base_mode %d dy %d dx %d: bad pixel entries %4d, bad channel entries %4d", pitch, base_mode, dy, dx, bad_pix, bad_ch);
                    if (first >= 0) {
                        printf(" | first bad row %d; rows 0..17 read pixel:", first);
                        for (int m = 0; m < 18; ++m) printf(" %g", h[m * 32 + 0]);
                        printf(" | ch0..7 of row %d:", first);
                        for (int n = 0; n < 8; ++n) printf(" %g", h[first * 32 + n]);
                    }
                    printf("\n");
                }
    return 0;
}
