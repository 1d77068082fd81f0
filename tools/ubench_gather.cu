// ubench_gather.cu -- micro-benchmark of the sparse-conv gather (global rows -> swizzled shared tile), sm_100a.
// Answers: what does a cp.async (LDGSTS.128) row gather cost per SM as a function of (a) how many of the rows are real
// vs zero-filled, (b) row locality, (c) producer warps per SM, (d) cp.async vs predicated cp.async + st.shared zero vs
// ld.global + st.shared.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_gather tools/ubench_gather.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// PATH 0: cp.async zfill for every row; 1: predicated cp.async (invalid lanes off) + st.shared zeros; 2: ld.global.v4 + st.shared.v4
template <int PATH>
__global__ void __launch_bounds__(256) k_gather(const float* __restrict__ tab, const int* __restrict__ idx, int pool, int steps,
                                                float* __restrict__ sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    unsigned char* sa = smem + warp * 16384;
    const uint32_t sa_u = smem_u32(sa);
    const int j = lane & 7, rg = lane >> 3;
    float acc = 0.f;
    for (int it = 0; it < steps; ++it) {
        const int* rows = idx + (size_t)(((blockIdx.x * nw + warp) * 131 + it) % pool) * 128 + rg * 32;
#pragma unroll
        for (int i0 = 0; i0 < 32; i0 += 4) {
            const int4 nq = *reinterpret_cast<const int4*>(rows + i0);
            const int r4[4] = {nq.x, nq.y, nq.z, nq.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int src = r4[u];
                const int rr = (i0 + u) & 7;
                const uint32_t dst = sa_u + (uint32_t)((rg * 4 + ((i0 + u) >> 3)) * 1024 + rr * 128 + ((j ^ rr) << 4));
                const bool ok = src >= 0;
                const float* p = tab + (size_t)(ok ? src : 0) * 32 + j * 4;
                if (PATH == 0) {
                    int sz = ok ? 16 : 0;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(p), "r"(sz) : "memory");
                } else if (PATH == 1) {
                    if (ok) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p) : "memory");
                    else asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "f"(0.f) : "memory");
                } else {
                    float4 v = ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        acc += *reinterpret_cast<float*>(sa + lane * 4);
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main() {
    const int R = 140000, POOL = 4096;
    float* tab; CK(cudaMalloc(&tab, (size_t)R * 128)); CK(cudaMemset(tab, 0, (size_t)R * 128));
    float* sink; CK(cudaMalloc(&sink, 4));
    int* d_idx; CK(cudaMalloc(&d_idx, (size_t)POOL * 128 * 4));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("device %s, %d SMs, nominal %d MHz\n", prop.name, prop.multiProcessorCount, clk_khz / 1000);
    const char* pname[] = {"all real, random rows", "50% real (random rows)", "50% real (groups of 4 rows)", "0% real", "all real, consecutive rows",
                           "17% real (random rows)", "50% real (groups of 32 rows)"};
    auto gen = [&](int pat) {
        std::vector<int> h((size_t)POOL * 128);
        uint64_t s = 88172645463325252ull;
        auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); };
        for (int t = 0; t < POOL; ++t) {
            int base = rnd() % (R - 128);
            for (int i = 0; i < 128; ++i) {
                int row = rnd() % R; bool ok = true;
                if (pat == 1) ok = rnd() & 1;
                if (pat == 2) { static bool g; if (i % 4 == 0) g = rnd() & 1; ok = g; }
                if (pat == 3) ok = false;
                if (pat == 4) row = base + i;
                if (pat == 5) ok = (rnd() % 100) < 17;
                if (pat == 6) { static bool g2; if (i % 32 == 0) g2 = rnd() & 1; ok = g2; }
                h[(size_t)t * 128 + i] = ok ? row : -1;
            }
        }
        CK(cudaMemcpy(d_idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        size_t real = 0; for (int v : h) real += v >= 0;
        return (double)real / h.size();
    };
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(k_gather<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384));
    CK(cudaFuncSetAttribute(k_gather<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384));
    CK(cudaFuncSetAttribute(k_gather<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384));
    const int steps = 400;
    printf("%-32s %-6s %5s %5s | %9s %12s %14s %16s\n", "pattern", "path", "warps", "cta/sm", "us", "real GB/s", "B/clk/SM(real)", "clk per LDGSTS/SM");
    for (int pat = 0; pat < 7; ++pat) {
        double frac = gen(pat);
        for (int path = 0; path < 3; ++path) {
            for (int cfg = 0; cfg < 3; ++cfg) {
                if (path != 0 && cfg != 0) continue;
                const int warps = cfg == 1 ? 8 : 4, cps = cfg == 2 ? 1 : 2;
                if (pat >= 3 && cfg != 0) continue;
                const int grid = prop.multiProcessorCount * cps;
                const size_t smem = (size_t)warps * 16384;
                auto run = [&]() {
                    const size_t sm = cps == 1 ? 120 * 1024 : smem;      // force 1 CTA/SM by shared-memory footprint
                    if (path == 0) k_gather<0><<<grid, warps * 32, sm>>>(tab, d_idx, POOL, steps, sink);
                    if (path == 1) k_gather<1><<<grid, warps * 32, sm>>>(tab, d_idx, POOL, steps, sink);
                    if (path == 2) k_gather<2><<<grid, warps * 32, sm>>>(tab, d_idx, POOL, steps, sink);
                };
                run(); CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0)); run(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                const double ksteps = (double)grid * warps * steps;
                const double real_bytes = ksteps * 16384.0 * frac;
                const double clk = ms * 1e-3 * 1.965e9;          // B200 boost clock under load (bench.py reports 1965 MHz)
                printf("%-32s %-6s %5d %5d | %9.1f %12.1f %14.2f %16.2f\n", pname[pat], path == 0 ? "zfill" : path == 1 ? "pred" : "ldg",
                       warps, cps, ms * 1e3, real_bytes / (ms * 1e-3) / 1e9, real_bytes / prop.multiProcessorCount / clk,
                       clk / (ksteps * 32.0 / prop.multiProcessorCount));
            }
        }
    }
    return 0;
}
