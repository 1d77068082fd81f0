// tc_stubs.cu -- tensor-core (tcgen05) entry points that are not implemented yet fail loudly: there is no silent
// fallback to the fp32 path.
#include "common.cuh"

int dz_attention_fwd_tc(const float*, int, const float*, int, const float*, int, const unsigned char*, int, int, int, int, int,
                        float*, int, int mode, cudaStream_t) {
    dz_set_error("dz_attention_fwd: tensor-core mode %d not built", mode); return DZ_ERR_UNSUPPORTED;
}
