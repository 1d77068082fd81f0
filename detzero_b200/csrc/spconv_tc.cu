// spconv_tc.cu -- sparse convolution forward on the 5th-gen tensor cores (tcgen05, TF32 operands, fp32 TMEM accumulators).
//
// Same contract as the exact-fp32 kernel in spconv.cu (SubMConv3d / SparseConv3d + folded BatchNorm1d + bias + residual
// + ReLU, backbone3d.py:64-83,105-121) for DZ_TF32.  spconv's own default keeps TF32 off, so this mode
// is opt-in (COMPUTE_MODE: tf32) with a stated 2e-3 tolerance.
//
// Design: output-stationary implicit GEMM.  A CTA owns 128 output rows -- the rows order[p] of one tile of the rulebook's
// tile schedule (rows with alike neighbour masks, csrc/rulebook.cu), tiles taken heaviest-first -- and reads their
// neighbour lists from the row-major table (one 128-byte line per row).  The reduction dimension is the concatenation
// (kernel offset k, input channel c) of length K*Cin, cut into 32-float (128-byte) blocks -- for Cin = 16 one block spans
// two offsets, for Cin = 64 an offset spans two blocks -- so narrow layers still fill the MMA K dimension.  Blocks whose
// offsets no row of the tile uses are skipped (that is what the schedule maximises).  Per block, the producer warps (PW
// per pipeline stage) gather the neighbour rows straight from global/L2 into shared memory with cp.async (zero-fill for
// missing neighbours) in the canonical K-major 128B-swizzled UMMA layout, one thread TMA-loads the matching
// W[cout][k-block] tile, and one thread issues tcgen05.mma (M=128, N=Cout, K=8) into TMEM.  Warps 0-3 then run the
// epilogue: tcgen05.ld -> scale/shift (+residual) -> ReLU -> 128-bit stores to row order[p].  No atomics: deterministic,
// and bit-identical with or without a schedule (a row's accumulation order over k never changes).
#include <stdlib.h>
#include "common.cuh"
#include "tc.cuh"

static constexpr int ST_ROWS = 128;
static constexpr int ST_A_BYTES = ST_ROWS * 128;        // one 32-float block for 128 rows
static constexpr int ST_KMAX = 27;

template <int COUT, int PW>
struct StCfg {
    static constexpr int W_BYTES = COUT * 128;
    static constexpr int STAGE_BYTES = ST_A_BYTES + W_BYTES;
    // 2 CTAs/SM (the other CTA's prologue/epilogue overlaps this one's main loop)
    // Cout = 128 only occurs with K = 3 (conv_out): few k-steps, so 2 CTAs/SM beat depth.  (3 stages / 3 CTAs per SM for
    // Cout <= 32 was measured: no gain.)
    static constexpr int STAGES = COUT >= 128 ? 3 : 4;
    // PW producer warps per stage: warp w gathers its 128/PW rows of the steps it = s, s+STAGES, ... into stage s = w % STAGES
    // (a single warp sustains only one LDGSTS per ~70-130 cycles, tools/ubench_gather.cu; the SM's LSU takes one per ~8.5);
    // warp STAGES*PW issues the MMAs and owns TMEM; warps 0-3 double as the epilogue
    static constexpr int PRODUCERS = STAGES * PW;
    static constexpr int THREADS = 32 * (PRODUCERS + 1);
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + ST_KMAX * ST_ROWS * 4 + 128 * 4 + 256;
    static constexpr int TMEM_COLS = COUT < 32 ? 32 : COUT;
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(uint32_t dst, const void* src, bool valid) {
    int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

template <int CIN_PAD, int COUT, int PW>
__global__ void __launch_bounds__(StCfg<COUT, PW>::THREADS)
k_spconv_tf32(const __grid_constant__ CUtensorMap tmW, const float* __restrict__ in, int cin, const int32_t* __restrict__ nbr, int K,
              int nbr_cap, const int32_t* __restrict__ order, const int* __restrict__ d_n_out, int out_cap, const float* __restrict__ scale,
              const float* __restrict__ shift, const float* __restrict__ residual, int relu, float* __restrict__ out, long long* __restrict__ dbgbuf) {
    using Cfg = StCfg<COUT, PW>;
    const bool tr = dbgbuf && blockIdx.x == 1;      // clock trace of CTA 1 (tools/trace_spconv.py); dbgbuf is NULL in normal runs
    if (tr && threadIdx.x == 0) dbgbuf[0] = clock64();
    const int n = min(*d_n_out, out_cap);
    // under a tile schedule CTAs take the tiles heaviest-first (launch order ~ blockIdx order): order[nbr_cap + i]
    const int row0 = (order ? __ldg(order + nbr_cap + blockIdx.x) : (int)blockIdx.x) * ST_ROWS;
    if (row0 >= n) return;

    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    int* s_nbr = reinterpret_cast<int*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);          // [K][128]
    int* s_blocks = s_nbr + ST_KMAX * ST_ROWS;                                            // active block list (<= 108)
    uint64_t* full = reinterpret_cast<uint64_t*>(s_blocks + 124);
    uint64_t* empty = full + Cfg::STAGES;
    uint64_t* tmem_full = empty + Cfg::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    int* s_nb = reinterpret_cast<int*>(tmem_slot + 1);
    unsigned* s_mask = reinterpret_cast<unsigned*>(s_nb + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ktot = K * CIN_PAD;
    const int nb_tot = (ktot + 31) / 32;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmW);
        for (int s = 0; s < Cfg::STAGES; ++s) { tc::mbar_init(full + s, 32 * PW + 1); tc::mbar_init(empty + s, 1); }
        tc::mbar_init(tmem_full, 1);
        tc::fence_barrier_init();
        *s_mask = 0u;
    }
    if (warp == Cfg::PRODUCERS) tc::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    __syncthreads();
    // ---- neighbour rows of this tile + which offsets are populated
    if (warp < 4) {
        const int r = row0 + threadIdx.x;
        // table row of this tile position (row-major table, one 128-byte line per output row; order[pos] under a schedule)
        const int src = r < n ? (order ? __ldg(order + r) : r) : -1;
        int v[ST_KMAX];
        {
            int w[28];
            const int4* rowp = reinterpret_cast<const int4*>(nbr) + (size_t)(src < 0 ? 0 : src) * 8;
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                const int4 t4 = src >= 0 ? __ldg(rowp + q) : make_int4(-1, -1, -1, -1);
                w[4 * q] = t4.x; w[4 * q + 1] = t4.y; w[4 * q + 2] = t4.z; w[4 * q + 3] = t4.w;
            }
#pragma unroll
            for (int k = 0; k < ST_KMAX; ++k) v[k] = k < K ? w[k] : -1;
        }
        unsigned mine = 0;
#pragma unroll
        for (int k = 0; k < ST_KMAX; ++k) {
            if (k < K) s_nbr[k * ST_ROWS + threadIdx.x] = v[k];
            mine |= (v[k] >= 0 ? 1u : 0u) << k;
        }
        mine = __reduce_or_sync(0xffffffffu, mine);
        if (lane == 0 && mine) atomicOr(s_mask, mine);
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    if (threadIdx.x == 0) {
        const unsigned mask = *s_mask;
        int nb = 0;
        for (int kb = 0; kb < nb_tot; ++kb) {
            int k_lo = (kb * 32) / CIN_PAD, k_hi = min(K - 1, (kb * 32 + 31) / CIN_PAD);
            bool act = false;
            for (int k = k_lo; k <= k_hi; ++k) act |= (mask >> k) & 1u;
            if (act) s_blocks[nb++] = kb;
        }
        *s_nb = nb;
    }
    __syncthreads();
    const int nb = *s_nb;
    const uint32_t tmem_base = *tmem_slot;
    if (tr && threadIdx.x == 0) { dbgbuf[1] = clock64(); dbgbuf[2] = nb; }

    if (warp < Cfg::PRODUCERS) {
        // ================= producers: warp w fills rows [part*RPW, (part+1)*RPW) of stage s for the steps s, s+STAGES, ... ===
        constexpr int RPW = ST_ROWS / PW;           // rows per producer warp
        constexpr int RPL = RPW / 4;                // rows per 8-lane group
        const int j = lane & 7;                     // 16-byte chunk inside the 128-byte row
        const int rg = lane >> 3;                   // lane covers chunk j of rows rbase .. rbase+RPL-1
        const int s = warp % Cfg::STAGES, part = warp / Cfg::STAGES;
        const int rbase = part * RPW + rg * RPL;
        unsigned char* sa = smem + s * Cfg::STAGE_BYTES;
        const uint32_t sa_u = tc::smem_u32(sa);
        for (int it = s, round = 0; it < nb; it += Cfg::STAGES, ++round) {
            const int kb = s_blocks[it];
            if (tr && lane == 0 && part == 0 && it < 120) dbgbuf[8 + it * 8 + 0] = clock64();
            tc::mbar_wait(empty + s, (round & 1) ^ 1);
            if (tr && lane == 0 && part == 0 && it < 120) dbgbuf[8 + it * 8 + 1] = clock64();
            if (part == 0 && lane == 0) {
                tc::mbar_arrive_expect_tx(full + s, Cfg::W_BYTES);
                tc::tma_load_2d(sa + ST_A_BYTES, &tmW, full + s, kb * 32, 0);
            }
            const int kidx = kb * 32 + j * 4;       // position in the concatenated (offset, channel) dimension
            const int k = kidx / CIN_PAD, c = kidx % CIN_PAD;
            const bool k_ok = k < K;
            const int4* nb_k = reinterpret_cast<const int4*>(s_nbr + (k_ok ? k : 0) * ST_ROWS + rbase);
            const float* in_c = in + c;
#pragma unroll
            for (int i0 = 0; i0 < RPL; i0 += 4) {
                const int4 nq = nb_k[i0 / 4];
                const int rows4[4] = {nq.x, nq.y, nq.z, nq.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int src_row = k_ok ? rows4[u] : -1;
                    const int r = rbase + i0 + u;   // tile row: swizzle atom r/8, row-in-atom r%8
                    const int rr = r & 7;
                    const uint32_t dst = sa_u + (uint32_t)((r >> 3) * 1024 + rr * 128 + ((j ^ rr) << 4));
                    if (CIN_PAD >= 16) {
                        const bool ok = src_row >= 0;
                        cp_async16_zfill(dst, ok ? (const void*)(in_c + (size_t)src_row * cin) : (const void*)in, ok);
                    } else {
                        // first layer (cin = 5 padded to 8): rows are not 16-byte aligned -> 4-byte copies
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const bool ok = src_row >= 0 && (c + e) < cin;
                            cp_async4_zfill(dst + 4 * e, ok ? (const void*)(in_c + (size_t)src_row * cin + e) : (const void*)in, ok);
                        }
                    }
                }
            }
            cp_async_mbar_arrive_noinc(full + s);
            if (tr && lane == 0 && part == 0 && it < 120) dbgbuf[8 + it * 8 + 2] = clock64();
        }
    }
    if (warp < 4) {
        // ================= epilogue =================
        const int q = warp;                         // TMEM lane quarter == warp id for warps 0..3
        const int pos = row0 + q * 32 + lane;       // tile position; the output row is order[pos] under a tile schedule
        const int r = pos < n ? (order ? __ldg(order + pos) : pos) : n;
        tc::mbar_wait(tmem_full, 0);
        tc::tcgen05_fence_after();
        if (tr && threadIdx.x == 0) dbgbuf[3] = clock64();
#pragma unroll 1
        for (int c0 = 0; c0 < COUT; c0 += 32) {
            float v[32];
            tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (r < n) {
#pragma unroll
                for (int jj = 0; jj < 32; jj += 4) {
                    const int ch = c0 + jj;
                    if (ch >= COUT) break;
                    float4 o = nb > 0 ? make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (scale) { float4 sc = __ldg(reinterpret_cast<const float4*>(scale + ch)); o.x *= sc.x; o.y *= sc.y; o.z *= sc.z; o.w *= sc.w; }
                    if (shift) { float4 sh = __ldg(reinterpret_cast<const float4*>(shift + ch)); o.x += sh.x; o.y += sh.y; o.z += sh.z; o.w += sh.w; }
                    if (residual) {
                        float4 rr = __ldg(reinterpret_cast<const float4*>(residual + (size_t)r * COUT + ch));
                        o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                    }
                    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    // round-to-nearest to TF32: the next layer's truncating tensor-core read is then exact (no toward-zero bias)
                    o.x = tc::rna_tf32(o.x); o.y = tc::rna_tf32(o.y); o.z = tc::rna_tf32(o.z); o.w = tc::rna_tf32(o.w);
                    *reinterpret_cast<float4*>(out + (size_t)r * COUT + ch) = o;
                }
            }
        }
    }
    if (warp == Cfg::PRODUCERS) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = tc::instr_desc(2, 128, COUT);
            for (int it = 0; it < nb; ++it) {
                const int s = it % Cfg::STAGES;
                if (tr && it < 120) dbgbuf[8 + it * 8 + 3] = clock64();
                tc::mbar_wait(full + s, (it / Cfg::STAGES) & 1);
                if (tr && it < 120) dbgbuf[8 + it * 8 + 4] = clock64();
                tc::fence_proxy_async();            // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
                tc::tcgen05_fence_after();
                const uint32_t sa = tc::smem_u32(smem + s * Cfg::STAGE_BYTES);
                const uint64_t adesc = tc::smem_desc_sw128(sa), bdesc = tc::smem_desc_sw128(sa + ST_A_BYTES);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    tc::mma_tf32(tmem_base, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (it | kk) ? 1u : 0u);
                tc::mma_commit(empty + s);
                if (tr && it < 120) dbgbuf[8 + it * 8 + 5] = clock64();
            }
            tc::mma_commit(tmem_full);
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (tr && threadIdx.x == 0) dbgbuf[4] = clock64();
    if (warp == Cfg::PRODUCERS) tc::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

long long* g_dbgbuf = nullptr;          // shared with spconv_bf16.cu (clock traces, tools/trace_*.py)
bool g_trace_on = false;
extern "C" int dz_debug_trace(long long* host, int n) {      // clock-trace helper for tools/trace_spconv.py (not part of the public header)
    if (!g_dbgbuf) { if (cudaMalloc(&g_dbgbuf, 4096 * 8) != cudaSuccess) return -1; cudaMemset(g_dbgbuf, 0, 4096 * 8); g_trace_on = true; return 1; }
    cudaDeviceSynchronize();
    return cudaMemcpy(host, g_dbgbuf, (size_t)n * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}

template <int CIN_PAD, int COUT, int PW>
static int launch_pw(const CUtensorMap& tmW, const float* in, int cin, const int32_t* nbr, int K, int nbr_cap, const int32_t* order, const int* d_n_out,
                     int out_cap, const float* scale, const float* shift, const float* residual, int relu, float* out, cudaStream_t st) {
    using Cfg = StCfg<COUT, PW>;
    static bool configured = false;
    if (!configured) {
        DZ_CUDA(cudaFuncSetAttribute(k_spconv_tf32<CIN_PAD, COUT, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        configured = true;
    }
    k_spconv_tf32<CIN_PAD, COUT, PW><<<dz_cdiv(out_cap, ST_ROWS), Cfg::THREADS, Cfg::SMEM, st>>>(tmW, in, cin, nbr, K, nbr_cap, order, d_n_out, out_cap,
                                                                                             scale, shift, residual, relu, out, g_trace_on ? g_dbgbuf : nullptr);
    DZ_LAUNCH_CHECK();
    return DZ_OK;
}

template <int CIN_PAD, int COUT>
static int launch(const CUtensorMap& tmW, const float* in, int cin, int in_rows, const int32_t* nbr, int K, int nbr_cap, const int32_t* order, const int* d_n_out,
                  int out_cap, const float* scale, const float* shift, const float* residual, int relu, float* out, cudaStream_t st) {
    (void)in_rows;
    static const int pw = getenv("DZ_SPCONV_PW") ? atoi(getenv("DZ_SPCONV_PW")) : 2;
    if (pw == 1) return launch_pw<CIN_PAD, COUT, 1>(tmW, in, cin, nbr, K, nbr_cap, order, d_n_out, out_cap, scale, shift, residual, relu, out, st);
    return launch_pw<CIN_PAD, COUT, 2>(tmW, in, cin, nbr, K, nbr_cap, order, d_n_out, out_cap, scale, shift, residual, relu, out, st);
}

// weight layout for this path: (cout, K * cin_pad) row-major, cin_pad = 8 for cin <= 8 else cin.
// (A TMA tile::gather4 producer was built and measured 1.6x slower than the cp.async gather for 128-byte rows -- the TMA unit
// has a per-row cost -- and a hybrid LSU+TMA split gave no gain: profiles/r01_spconv_notes.md.  Both were removed.)
int dz_spconv_fwd_tc3(const float* in, int cin, const int32_t* nbr, int K, int nbr_cap, const int32_t* order, const int* d_n_out, int out_cap,
                      const float* weight, const float* scale, const float* shift, const float* residual, int relu, float* out,
                      int cout, cudaStream_t st);

int dz_spconv_fwd_tc(const float* in, int cin, int in_rows, const int32_t* nbr, int K, int nbr_cap, const int32_t* order, const int* d_n_out, int out_cap,
                     const float* weight, const float* scale, const float* shift, const float* residual, int relu, float* out,
                     int cout, int mode, cudaStream_t st) {
    if (mode == DZ_TF32X3) return dz_spconv_fwd_tc3(in, cin, nbr, K, nbr_cap, order, d_n_out, out_cap, weight, scale, shift, residual, relu, out, cout, st);
    if (mode != DZ_TF32) { dz_set_error("dz_spconv_fwd: tensor-core mode %d not built (tf32 only)", mode); return DZ_ERR_UNSUPPORTED; }
    if (K > ST_KMAX) { dz_set_error("dz_spconv_fwd(tf32): K=%d > 27", K); return DZ_ERR_UNSUPPORTED; }
    const int cin_pad = cin <= 8 ? 8 : cin;
    tc::EncodeTiledFn enc = tc::get_encode_tiled();
    if (!enc) { dz_set_error("cuTensorMapEncodeTiled unavailable"); return DZ_ERR_CUDA; }
    CUtensorMap tmW;
    {
        cuuint64_t ktot = (cuuint64_t)K * cin_pad;
        cuuint64_t dims[2] = {ktot, (cuuint64_t)cout};
        cuuint64_t strides[1] = {ktot * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)cout};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)weight, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { dz_set_error("cuTensorMapEncodeTiled(W) failed: %d", (int)r); return DZ_ERR_CUDA; }
    }
#define DZ_ST(CP, CO) return launch<CP, CO>(tmW, in, cin, in_rows, nbr, K, nbr_cap, order, d_n_out, out_cap, scale, shift, residual, relu, out, st)
    if (cin_pad == 8 && cout == 16) DZ_ST(8, 16);
    if (cin_pad == 16 && cout == 16) DZ_ST(16, 16);
    if (cin_pad == 16 && cout == 32) DZ_ST(16, 32);
    if (cin_pad == 32 && cout == 32) DZ_ST(32, 32);
    if (cin_pad == 32 && cout == 64) DZ_ST(32, 64);
    if (cin_pad == 64 && cout == 64) DZ_ST(64, 64);
    if (cin_pad == 64 && cout == 128) DZ_ST(64, 128);
    if (cin_pad == 128 && cout == 128) DZ_ST(128, 128);
#undef DZ_ST
    dz_set_error("dz_spconv_fwd(tf32): (cin=%d, cout=%d) not instantiated", cin, cout);
    return DZ_ERR_UNSUPPORTED;
}
