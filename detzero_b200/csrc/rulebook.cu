// rulebook.cu -- rulebook (neighbour-table) build for submanifold and strided sparse convolution, sm_100a.
//
// Replaces spconv's generate_subm_conv_inds / generate_conv_inds_stage{1,2} (GPU hash table) used through
// SubMConv3d / SparseConv3d at detection/detzero_det/models/centerpoint_modules/backbone3d.py:68-71,93-100.
//
// Design: coordinates are looked up in the L2-resident grid index (bitmap + popcount prefix) -- two dependent
// loads whose cache lines are shared by x-adjacent probes -- instead of probing a hash table.  The output is the
// output-stationary neighbour table nbr[k][o] (k-major => coalesced writes here and coalesced reads in the conv
// kernel); strided-conv output sites come out of a bitmap rank scan already sorted by (b,z,y,x), so no sort.
#include "common.cuh"

struct ConvGeom {
    int k[3], s[3], p[3];
    int in_dhw[3], out_dhw[3];
};

__global__ void __launch_bounds__(256) k_subm_nbr(const int32_t* __restrict__ coords, const int* __restrict__ d_n, int cap,
                                                  GridIndex g, int KD, int KH, int KW, int32_t* __restrict__ nbr) {
    int n = min(*d_n, cap);
    int K = KD * KH * KW;
    int hz = (KD - 1) / 2, hy = (KH - 1) / 2, hx = (KW - 1) / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 c = __ldg(reinterpret_cast<const int4*>(coords) + i);    // b,z,y,x
        int k = 0;
        for (int kz = 0; kz < KD; ++kz)
            for (int ky = 0; ky < KH; ++ky)
                for (int kx = 0; kx < KW; ++kx, ++k) {
                    int j;
                    if (kz == hz && ky == hy && kx == hx) j = i;        // centre tap is the site itself
                    else j = grid_lookup(g, c.x, c.y + kz - hz, c.z + ky - hy, c.w + kx - hx);
                    nbr[(size_t)k * cap + i] = j;
                }
        (void)K;
    }
}

extern "C" int dz_rulebook_subm(const int32_t* coords, const int* d_n, int cap, int B, int D, int H, int W,
                                const int* ks, const uint32_t* bitmap, const uint32_t* prefix, const int32_t* perm,
                                int32_t* nbr, dz_stream_t stream) {
    DZ_CHECK_ARG(coords && d_n && bitmap && prefix && nbr && cap >= 1);
    DZ_CHECK_ARG(ks[0] % 2 == 1 && ks[1] % 2 == 1 && ks[2] % 2 == 1);
    GridIndex g{bitmap, prefix, perm, B, D, H, W, dz_cells_pad(D, H, W)};
    int blocks = max(1, min(dz_cdiv(cap, 256), DZ_NUM_SMS * 8));
    k_subm_nbr<<<blocks, 256, 0, (cudaStream_t)stream>>>(coords, d_n, cap, g, ks[0], ks[1], ks[2], nbr);
    DZ_LAUNCH_CHECK();
    return DZ_OK;
}

// mark every output cell reachable from an active input: o = (i + p - k) / s when divisible and in range
__global__ void __launch_bounds__(256) k_conv_mark(const int32_t* __restrict__ coords, const int* __restrict__ d_n, int cap,
                                                   ConvGeom cg, long long out_cells_pad, uint32_t* __restrict__ out_bitmap) {
    int n = min(*d_n, cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 c = __ldg(reinterpret_cast<const int4*>(coords) + i);
        int iz[3] = {c.y, c.z, c.w};
        // per-dimension candidate outputs (at most ceil(k/s) each)
        int oz[3][3], cnt[3];
        for (int d = 0; d < 3; ++d) {
            cnt[d] = 0;
            for (int k = 0; k < cg.k[d]; ++k) {
                int num = iz[d] + cg.p[d] - k;
                if (num < 0 || num % cg.s[d]) continue;
                int o = num / cg.s[d];
                if (o >= cg.out_dhw[d]) continue;
                oz[d][cnt[d]++] = o;
            }
        }
        for (int a = 0; a < cnt[0]; ++a)
            for (int b = 0; b < cnt[1]; ++b)
                for (int e = 0; e < cnt[2]; ++e) {
                    long long cell = (long long)c.x * out_cells_pad +
                                     ((long long)oz[0][a] * cg.out_dhw[1] + oz[1][b]) * cg.out_dhw[2] + oz[2][e];
                    uint32_t bit = 1u << (cell & 31);
                    uint32_t* wp = out_bitmap + (cell >> 5);
                    if (!(*wp & bit)) atomicOr(wp, bit);        // plain read first: most cells are already set
                }
    }
}

// bitmap words -> sorted coordinate list
__global__ void __launch_bounds__(256) k_index_to_coords(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ prefix,
                                                         size_t n_words, int D, int H, int W, long long cells_pad, int cap,
                                                         int32_t* __restrict__ coords) {
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < n_words; w += (size_t)gridDim.x * blockDim.x) {
        uint32_t word = __ldg(bitmap + w);
        if (!word) continue;
        int rank = (int)__ldg(prefix + w);
        long long cell0 = (long long)w << 5;
        while (word) {
            int bit = __ffs(word) - 1;
            word &= word - 1;
            if (rank < cap) {
                long long cell = cell0 + bit;
                int b = (int)(cell / cells_pad);
                long long r = cell % cells_pad;
                int x = (int)(r % W); r /= W;
                int y = (int)(r % H);
                int z = (int)(r / H);
                reinterpret_cast<int4*>(coords)[rank] = make_int4(b, z, y, x);
            }
            ++rank;
        }
    }
}

__global__ void __launch_bounds__(256) k_conv_nbr(const int32_t* __restrict__ out_coords, const int* __restrict__ d_n_out, int out_cap,
                                                  ConvGeom cg, GridIndex gin, int32_t* __restrict__ nbr) {
    int n = min(*d_n_out, out_cap);
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
        int4 c = __ldg(reinterpret_cast<const int4*>(out_coords) + o);
        int k = 0;
        for (int kz = 0; kz < cg.k[0]; ++kz)
            for (int ky = 0; ky < cg.k[1]; ++ky)
                for (int kx = 0; kx < cg.k[2]; ++kx, ++k) {
                    int z = c.y * cg.s[0] - cg.p[0] + kz;
                    int y = c.z * cg.s[1] - cg.p[1] + ky;
                    int x = c.w * cg.s[2] - cg.p[2] + kx;
                    nbr[(size_t)k * out_cap + o] = grid_lookup(gin, c.x, z, y, x);
                }
    }
}

__global__ void k_clamp_count(int* d_n, int cap) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && *d_n > cap) *d_n = cap;   // overflow is reported by the host wrapper
}

extern "C" int dz_rulebook_conv(const int32_t* in_coords, const int* d_n_in, int in_cap, int B, const int* in_dhw,
                                const int* ks, const int* st_, const int* pd, const uint32_t* in_bitmap,
                                const uint32_t* in_prefix, const int32_t* in_perm, int32_t* out_coords, int* d_n_out,
                                int out_cap, uint32_t* out_bitmap, uint32_t* out_prefix, int32_t* nbr, void* ws,
                                size_t ws_bytes, dz_stream_t stream) {
    DZ_CHECK_ARG(in_coords && d_n_in && in_bitmap && in_prefix && out_coords && d_n_out && out_bitmap && out_prefix && nbr);
    DZ_CHECK_ARG(in_cap >= 1 && out_cap >= 1 && B >= 1);
    ConvGeom cg;
    for (int d = 0; d < 3; ++d) {
        DZ_CHECK_ARG(ks[d] >= 1 && ks[d] <= 3 && st_[d] >= 1 && pd[d] >= 0);
        cg.k[d] = ks[d]; cg.s[d] = st_[d]; cg.p[d] = pd[d]; cg.in_dhw[d] = in_dhw[d];
        cg.out_dhw[d] = (in_dhw[d] + 2 * pd[d] - (ks[d] - 1) - 1) / st_[d] + 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    size_t out_words = dz_grid_index_words(B, cg.out_dhw[0], cg.out_dhw[1], cg.out_dhw[2]);
    if (ws_bytes < dz_scan_ws_bytes(out_words)) { dz_set_error("dz_rulebook_conv: workspace too small"); return DZ_ERR_WORKSPACE; }
    long long out_cp = dz_cells_pad(cg.out_dhw[0], cg.out_dhw[1], cg.out_dhw[2]);
    int blocks_in = max(1, min(dz_cdiv(in_cap, 256), DZ_NUM_SMS * 8));
    k_conv_mark<<<blocks_in, 256, 0, st>>>(in_coords, d_n_in, in_cap, cg, out_cp, out_bitmap);
    int rc = dz_grid_index_scan(out_bitmap, out_prefix, out_words, nullptr, d_n_out, ws, ws_bytes, stream);
    if (rc) return rc;
    int blocks_w = max(1, min(dz_cdiv((long long)out_words, 256), DZ_NUM_SMS * 8));
    k_index_to_coords<<<blocks_w, 256, 0, st>>>(out_bitmap, out_prefix, out_words, cg.out_dhw[0], cg.out_dhw[1], cg.out_dhw[2],
                                                out_cp, out_cap, out_coords);
    GridIndex gin{in_bitmap, in_prefix, in_perm, B, in_dhw[0], in_dhw[1], in_dhw[2], dz_cells_pad(in_dhw[0], in_dhw[1], in_dhw[2])};
    int blocks_out = max(1, min(dz_cdiv(out_cap, 256), DZ_NUM_SMS * 8));
    k_conv_nbr<<<blocks_out, 256, 0, st>>>(out_coords, d_n_out, out_cap, cg, gin, nbr);
    DZ_LAUNCH_CHECK();
    return DZ_OK;
}
