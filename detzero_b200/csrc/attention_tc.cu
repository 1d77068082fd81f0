// attention_tc.cu -- multi-head attention core on the tensor cores (tcgen05, TF32 operands, fp32 TMEM accumulators), sm_100a.
//
// Replaces bmm(q,k^T) + key_padding masked_fill(-inf) + softmax + bmm(.,v) of multi_head_attention_forward
// (refining/detzero_refine/models/modules/transformer/multi_head_attention.py:266-286), head_dim 32, without ever
// materialising the (B*H, Pq, Pk) score tensor (15.7 GB at the BASELINE config-4 size).
//
// One CTA = one (batch, head) x one tile of 128 queries.  TWO passes over the keys instead of the usual online-softmax
// rescaling: pass 1 computes S = Q K^T block by block on the tensor cores and only tracks the row maxima; pass 2 recomputes
// S, forms P = exp(S - max) and accumulates O += P V in TMEM.  With head_dim 32 the QK^T MMAs are 4 of the 20 MMAs per
// key block, so recomputing them costs 25 % more tensor work and removes the accumulator-correction path entirely.
//   warps 4-7: loaders -- warp 4 lane 0: TMA for the Q tile and the K blocks (rows of 128 B = one head slice, SWIZZLE_128B =
//             K-major UMMA operand); all four: register transpose of one 32-key k-block of V each into V^T (the B operand of
//             P V must be K-major in keys).  One warp doing all four k-blocks was the bottleneck of round 1's kernel: its
//             32 L2-latency-exposed row loads + 128 scalar shared stores per lane took ~5 K clk per 128-key block
//   warp 8  : single-thread MMA issuer (S = Q K^T : M128 N128 K32 ; O += P V^T : M128 N32 K128)
//   warps 0-15: softmax.  Thread = (query row, 32-key column group): tcgen05.ld of its 32 scores, masking, max / exp / sum, P
//             written to shared memory in the swizzled A-operand layout; the four column groups of a row combine their maxima
//             after pass 1 and their sums at the end through shared memory.  (Round 1 used ONE thread per row, 128 scores each:
//             the clock trace showed those 4 warps busy 817 K of the CTA's 870 K clocks while the MMA thread waited for P.)
// S is double-buffered in TMEM so S(j+1) is computed while the softmax of block j runs.
#include "common.cuh"
#include "tc.cuh"

static constexpr int AT_M = 128;            // queries per CTA
static constexpr int AT_N = 128;            // keys per block
static constexpr int AT_D = 32;             // head dim
static constexpr int AT_TILE = AT_M * 128;  // 16 KB: 128 rows x 128 B
static constexpr int AT_SM_WARPS = 16;          // softmax warps: 4 TMEM lane quarters x 4 column groups of 32 keys
static constexpr int AT_THREADS = 32 * (AT_SM_WARPS + 5);   // warps 0-15 softmax, 16-19 loaders (V^T: one 32-key k-block each), 20 MMA

// validity of 32 consecutive keys as a bit mask (1 = masked); 4 mask bytes per load when the row is 4-byte aligned
__device__ __forceinline__ uint32_t masked_bits32(const unsigned char* mrow, int key0, int Pk, bool fast) {
    uint32_t bits = 0;
    if (fast) {
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int key = key0 + 4 * w;
            uint32_t m4 = key < Pk ? (mrow ? __ldg(reinterpret_cast<const uint32_t*>(mrow + key)) : 0u) : 0x01010101u;
            bits |= ((m4 & 0xffu) ? 1u : 0u) << (4 * w);
            bits |= ((m4 & 0xff00u) ? 1u : 0u) << (4 * w + 1);
            bits |= ((m4 & 0xff0000u) ? 1u : 0u) << (4 * w + 2);
            bits |= ((m4 & 0xff000000u) ? 1u : 0u) << (4 * w + 3);
        }
    } else {
        for (int c = 0; c < 32; ++c) {
            const int key = key0 + c;
            if (key >= Pk || (mrow && mrow[key])) bits |= 1u << c;
        }
    }
    return bits;
}

__device__ long long g_at_trace[64];         // clock accounting of CTA (0,0) (tools/trace_attention.py); written only when g_at_on != 0
__device__ int g_at_on = 0;
extern "C" int dz_debug_attention_trace(long long* host, int on) {
    cudaDeviceSynchronize();
    if (cudaMemcpyToSymbol(g_at_on, &on, sizeof(int)) != cudaSuccess) return -1;
    return host ? (cudaMemcpyFromSymbol(host, g_at_trace, sizeof(long long) * 64) == cudaSuccess ? 0 : -1) : 0;
}
#define AT_T(slot, stmt) do { if (tr) { const long long c0__ = clock64(); stmt; acc[slot] += clock64() - c0__; } else { stmt; } } while (0)

// 2^x on the MUFU pipe, one instruction (exp2f() is a ~6-instruction sequence; its arguments here are <= 0 and results land in
// (0, 1], where ex2.approx is good to 2 ulp -- far inside the TF32 tolerance of this mode)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AtSmem {
    static constexpr int Q = 0;
    static constexpr int K = AT_TILE;                    // 2 stages
    static constexpr int VT = K + 2 * AT_TILE;           // 2 stages of 4 k-blocks x (32 rows x 128 B) = 16 KB each
    static constexpr int P = VT + 2 * AT_TILE;           // 2 buffers x 4 k-blocks x 16 KB: the softmax of block j+1 writes P while the
    static constexpr int BARS = P + 8 * AT_TILE;         // tensor core still reads block j's (single-buffered they took turns)
    static constexpr int RED = BARS + 256;               // [2][4][128] floats: row max / row sum partials of the column groups
    static constexpr int TOTAL = RED + 2 * 4 * 128 * 4 + 1024;
};

__global__ void __launch_bounds__(AT_THREADS, 1)
k_attention_tf32(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, int q_col0, int k_col0,
                 const float* __restrict__ v, int ldv, const unsigned char* __restrict__ kpm, int Pq, int Pk, int H,
                 float* __restrict__ out, int ldo) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AtSmem::BARS);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;      // [2]
    uint64_t* k_empty = bars + 3;     // [2]
    uint64_t* v_full = bars + 5;      // [2]
    uint64_t* v_empty = bars + 7;     // [2]
    uint64_t* s_full = bars + 9;      // [2]
    uint64_t* s_empty = bars + 11;    // [2]
    uint64_t* p_full = bars + 13;     // [2]
    uint64_t* p_empty = bars + 15;    // [2]
    uint64_t* o_full = bars + 17;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool tr = g_at_on && blockIdx.x == 0 && blockIdx.y == 0;
    long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_start = tr ? clock64() : 0;
    const int bh = blockIdx.y, b = bh / H, h = bh % H;
    const int q0 = blockIdx.x * AT_M;
    const int nblk = (Pk + AT_N - 1) / AT_N;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmQ);
        tc::prefetch_tmap(&tmK);
        tc::mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(k_full + i, 1); tc::mbar_init(k_empty + i, 1);
            tc::mbar_init(v_full + i, 4); tc::mbar_init(v_empty + i, 1);
            tc::mbar_init(s_full + i, 1); tc::mbar_init(s_empty + i, AT_SM_WARPS);     // one arrival per WARP: 512 per-thread arrivals on one
                                                                                        // mbarrier serialise in shared memory (~1.5 K clk per block)
        }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(p_full + i, AT_SM_WARPS); tc::mbar_init(p_empty + i, 1); }
        tc::mbar_init(o_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == AT_SM_WARPS + 4) tc::tmem_alloc<512>(tmem_slot);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_s[2] = {tmem, tmem + 128};
    const uint32_t tmem_o = tmem + 256;

    if (warp >= AT_SM_WARPS && warp < AT_SM_WARPS + 4) {
        // ============================== loaders ==============================
        const int kk = warp - AT_SM_WARPS;                          // this warp's k-block (32 keys) of every V block
        if (kk == 0 && lane == 0) {
            tc::mbar_arrive_expect_tx(q_full, AT_TILE);
            tc::tma_load_2d(smem + AtSmem::Q, &tmQ, q_full, q_col0 + h * AT_D, b * Pq + q0);
        }
        for (int it = 0; it < 2 * nblk; ++it) {                     // K blocks: pass 1 then pass 2
            const int j = it % nblk, st = it & 1;
            if (kk == 0 && lane == 0) {
                AT_T(0, tc::mbar_wait(k_empty + st, ((it >> 1) & 1) ^ 1));
                tc::mbar_arrive_expect_tx(k_full + st, AT_TILE);
                tc::tma_load_2d(smem + AtSmem::K + st * AT_TILE, &tmK, k_full + st, k_col0 + h * AT_D, b * Pk + j * AT_N);
            }
            if (it >= nblk) {                                       // pass 2 also needs V_j^T
                const int jv = it - nblk, sv = jv & 1;
                AT_T(1, tc::mbar_wait(v_empty + sv, ((jv >> 1) & 1) ^ 1));
                const long long tv0 = tr ? clock64() : 0;
                unsigned char* vt = smem + AtSmem::VT + sv * AT_TILE;
                // lane handles key kk*32 + lane ; 8 float4 per key ; writes V^T[d][key] swizzled
                {
                    const int key = kk * 32 + lane, kg = jv * AT_N + key;
                    const float* vp = v + ((size_t)b * Pk + kg) * ldv + h * AT_D;
                    float4 x[8];
#pragma unroll
                    for (int d4 = 0; d4 < 8; ++d4) x[d4] = (kg < Pk) ? __ldg(reinterpret_cast<const float4*>(vp) + d4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int d4 = 0; d4 < 8; ++d4) {
                        const float e[4] = {x[d4].x, x[d4].y, x[d4].z, x[d4].w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int d = d4 * 4 + i;               // row of V^T
                            // k-block kk (32 keys = 128 B), row d: (d>>3)*1024 + (d&7)*128, 16-byte chunk (lane>>2) ^ (d&7)
                            const uint32_t off = (uint32_t)(kk * 4096 + (d >> 3) * 1024 + (d & 7) * 128 + (((lane >> 2) ^ (d & 7)) << 4) + (lane & 3) * 4);
                            *reinterpret_cast<float*>(vt + off) = e[i];
                        }
                    }
                }
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(v_full + sv);
                if (tr) acc[2] += clock64() - tv0;
            }
        }
        if (tr && kk == 0 && lane == 0) { g_at_trace[0] = acc[0]; g_at_trace[1] = acc[1]; g_at_trace[2] = acc[2]; g_at_trace[3] = clock64() - t_start; }
    } else if (warp == AT_SM_WARPS + 4) {
        // ============================== MMA issuer ==============================
        if (lane == 0) {
            constexpr uint32_t idesc_s = tc::instr_desc(2, 128, AT_N);
            constexpr uint32_t idesc_o = tc::instr_desc(2, 128, AT_D);
            const uint64_t qdesc = tc::smem_desc_sw128(tc::smem_u32(smem + AtSmem::Q));
            tc::mbar_wait(q_full, 0);
            auto issue_s = [&](int it) {                            // S[it&1] = Q K_it^T
                const int st = it & 1;
                AT_T(0, tc::mbar_wait(k_full + st, (it >> 1) & 1));
                AT_T(1, tc::mbar_wait(s_empty + st, ((it >> 1) & 1) ^ 1));
                const long long ti0 = tr ? clock64() : 0;
                tc::tcgen05_fence_after();
                const uint64_t kdesc = tc::smem_desc_sw128(tc::smem_u32(smem + AtSmem::K + st * AT_TILE));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) tc::mma_tf32(tmem_s[st], qdesc + (uint64_t)(kk * 2), kdesc + (uint64_t)(kk * 2), idesc_s, kk ? 1u : 0u);
                tc::mma_commit(k_empty + st);
                tc::mma_commit(s_full + st);
                if (tr) acc[4] += clock64() - ti0;
            };
            for (int it = 0; it < nblk; ++it) issue_s(it);          // pass 1: scores only
            if (tr) g_at_trace[15] = clock64() - t_start;
            issue_s(nblk);                                          // first block of pass 2
            for (int j = 0; j < nblk; ++j) {
                if (j + 1 < nblk) issue_s(nblk + j + 1);            // S of the next block overlaps this block's softmax
                const int sv = j & 1;
                AT_T(2, tc::mbar_wait(v_full + sv, (j >> 1) & 1));
                AT_T(3, tc::mbar_wait(p_full + (j & 1), (j >> 1) & 1));
                const long long tp0 = tr ? clock64() : 0;
                tc::tcgen05_fence_after();
                const uint32_t pa = tc::smem_u32(smem + AtSmem::P + (j & 1) * 4 * AT_TILE), va = tc::smem_u32(smem + AtSmem::VT + sv * AT_TILE);
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) {
                    const uint64_t pdesc = tc::smem_desc_sw128(pa + (kk >> 2) * AT_TILE) + (uint64_t)((kk & 3) * 2);
                    const uint64_t vdesc = tc::smem_desc_sw128(va + (kk >> 2) * 4096) + (uint64_t)((kk & 3) * 2);
                    tc::mma_tf32(tmem_o, pdesc, vdesc, idesc_o, (j | kk) ? 1u : 0u);
                }
                tc::mma_commit(v_empty + sv);
                tc::mma_commit(p_empty + (j & 1));
                if (tr) acc[5] += clock64() - tp0;
            }
            tc::mma_commit(o_full);
            if (tr) { for (int i = 0; i < 6; ++i) g_at_trace[8 + i] = acc[i]; g_at_trace[14] = clock64() - t_start; }
        }
    } else {
        // ============================== softmax / epilogue: thread = (query row, 32-key column group) ==============================
        const int qq = warp & 3, cg = warp >> 2;                    // TMEM lane quarter (== warp % 4), column group
        const int row = qq * 32 + lane;
        const int qi = q0 + row;
        const uint32_t lane_base = (uint32_t)(qq * 32) << 16;
        const unsigned char* mrow = kpm ? kpm + (size_t)b * Pk : nullptr;
        const bool fast = (Pk & 3) == 0 && ((reinterpret_cast<uintptr_t>(mrow) & 3) == 0);
        float* red = reinterpret_cast<float*>(smem + AtSmem::RED);
        const int c0 = cg * 32;
        float rmax = -INFINITY;
        for (int it = 0; it < nblk; ++it) {                         // ---- pass 1: row maxima
            const int st = it & 1;
            AT_T(0, tc::mbar_wait(s_full + st, (it >> 1) & 1));
            const long long tb0 = tr ? clock64() : 0;
            tc::tcgen05_fence_after();
            float sc[32];
            tc::tmem_ld32(tmem_s[st] + lane_base + (uint32_t)c0, sc);
            const uint32_t mb = masked_bits32(mrow, it * AT_N + c0, Pk, fast);
            if (mb == 0u) {                                         // the mask depends on the key only: uniform over the CTA
#pragma unroll
                for (int c = 0; c < 32; ++c) rmax = fmaxf(rmax, sc[c]);
            } else if (mb != 0xffffffffu) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (!((mb >> c) & 1u)) rmax = fmaxf(rmax, sc[c]);
            }
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(s_empty + st);
            if (tr) acc[1] += clock64() - tb0;
        }
        red[cg * 128 + row] = rmax;                                 // combine the four column groups of the row
        asm volatile("bar.sync 1, %0;" ::"n"(32 * AT_SM_WARPS) : "memory");
        rmax = fmaxf(fmaxf(red[row], red[128 + row]), fmaxf(red[256 + row], red[384 + row]));
        if (tr && threadIdx.x == 0) g_at_trace[23] = clock64() - t_start;
        float lsum = 0.f;
        const float LOG2E = 1.4426950408889634f;
        const float mscaled = rmax * LOG2E;
        unsigned char* pbase0 = smem + AtSmem::P + cg * AT_TILE + (row >> 3) * 1024 + (row & 7) * 128;     // k-block cg of P
        for (int j = 0; j < nblk; ++j) {                            // ---- pass 2: P = exp(S - max), O += P V
            const int it = nblk + j, st = it & 1;
            AT_T(2, tc::mbar_wait(s_full + st, (it >> 1) & 1));
            tc::tcgen05_fence_after();
            AT_T(3, tc::mbar_wait(p_empty + (j & 1), ((j >> 1) & 1) ^ 1));   // the P buffer of block j-2 has been consumed by the tensor core
            const long long tb0 = tr ? clock64() : 0;
            unsigned char* pk = pbase0 + (j & 1) * 4 * AT_TILE;
            float sc[32];
            tc::tmem_ld32(tmem_s[st] + lane_base + (uint32_t)c0, sc);
            const uint32_t mb = masked_bits32(mrow, j * AT_N + c0, Pk, fast);
            if (mb == 0u) {
#pragma unroll
                for (int c = 0; c < 32; ++c) { sc[c] = ex2_approx(fmaf(sc[c], LOG2E, -mscaled)); lsum += sc[c]; }
            } else if (mb == 0xffffffffu) {
#pragma unroll
                for (int c = 0; c < 32; ++c) sc[c] = 0.f;
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float pv = ((mb >> c) & 1u) ? 0.f : ex2_approx(fmaf(sc[c], LOG2E, -mscaled));
                    lsum += pv;
                    sc[c] = pv;
                }
            }
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
                *reinterpret_cast<float4*>(pk + ((ch ^ (row & 7)) << 4)) = make_float4(sc[4 * ch], sc[4 * ch + 1], sc[4 * ch + 2], sc[4 * ch + 3]);
            tc::fence_proxy_async();
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) { tc::mbar_arrive(p_full + (j & 1)); tc::mbar_arrive(s_empty + st); }
            if (tr) acc[4] += clock64() - tb0;
        }
        if (tr && threadIdx.x == 0) { for (int i = 0; i < 5; ++i) g_at_trace[16 + i] = acc[i]; g_at_trace[22] = clock64() - t_start; g_at_trace[24] = nblk; }
        red[512 + cg * 128 + row] = lsum;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * AT_SM_WARPS) : "memory");
        if (cg == 0) {
            lsum = (red[512 + row] + red[512 + 128 + row]) + (red[512 + 256 + row] + red[512 + 384 + row]);
            tc::mbar_wait(o_full, 0);
            tc::tcgen05_fence_after();
            float o[32];
            tc::tmem_ld32(tmem_o + lane_base, o);
            if (qi < Pq) {
                float* op = out + ((size_t)b * Pq + qi) * ldo + h * AT_D;
                const float inv = 1.f / lsum;                       // lsum == 0 (every key masked) -> NaN like softmax of all -inf
#pragma unroll
                for (int d = 0; d < AT_D; d += 4) {
                    float4 r = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
                    if (lsum == 0.f) r = make_float4(NAN, NAN, NAN, NAN);
                    *reinterpret_cast<float4*>(op + d) = r;
                }
            }
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == AT_SM_WARPS + 4) tc::tmem_dealloc<512>(tmem);
}

int dz_attention_fwd_tc(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, const unsigned char* kpm, int B,
                        int Pq, int Pk, int H, int dh, float* out, int ldo, int mode, cudaStream_t st) {
    if (mode != DZ_TF32) { dz_set_error("dz_attention_fwd: tensor-core mode %d not built (tf32 only)", mode); return DZ_ERR_UNSUPPORTED; }
    if (dh != AT_D) { dz_set_error("dz_attention_fwd(tf32): head_dim %d unsupported (32)", dh); return DZ_ERR_UNSUPPORTED; }
    if ((ldq & 3) || (ldk & 3) || (ldv & 3) || (ldo & 3) || (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15)) {
        dz_set_error("dz_attention_fwd(tf32): rows must be 16-byte aligned"); return DZ_ERR_ARG;
    }
    tc::EncodeTiledFn enc = tc::get_encode_tiled();
    if (!enc) { dz_set_error("cuTensorMapEncodeTiled unavailable"); return DZ_ERR_CUDA; }
    CUtensorMap tmQ, tmK;
    const float* ptrs[2] = {q, k};
    const int lds[2] = {ldq, ldk};
    const long long rows[2] = {(long long)B * Pq, (long long)B * Pk};
    CUtensorMap* maps[2] = {&tmQ, &tmK};
    for (int i = 0; i < 2; ++i) {
        // the head slice starts at column h*32 of a row of ld floats; the map covers the H*32 columns reachable from the base pointer
        cuuint64_t dims[2] = {(cuuint64_t)(H * AT_D), (cuuint64_t)rows[i]};
        cuuint64_t strides[1] = {(cuuint64_t)lds[i] * 4};
        cuuint32_t box[2] = {AT_D, AT_M};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptrs[i], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { dz_set_error("cuTensorMapEncodeTiled(attention %d) failed: %d", i, (int)r); return DZ_ERR_CUDA; }
    }
    static bool configured = false;
    if (!configured) {
        DZ_CUDA(cudaFuncSetAttribute(k_attention_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, AtSmem::TOTAL));
        configured = true;
    }
    dim3 grid(dz_cdiv(Pq, AT_M), B * H);
    k_attention_tf32<<<grid, AT_THREADS, AtSmem::TOTAL, st>>>(tmQ, tmK, 0, 0, v, ldv, kpm, Pq, Pk, H, out, ldo);
    DZ_LAUNCH_CHECK();
    return DZ_OK;
}
