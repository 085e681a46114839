// tcgen05 tensor-core convolution path (placeholder: not taking any op yet).
#include "common.cuh"
#include "conv_params.cuh"
bool dh_tc_supported(const ConvParams& p, const dh_packed_w* packed, bool separable) {
    (void)p; (void)packed; (void)separable;
    return false;
}
int dh_launch_conv_tc(dh_ctx* ctx, const ConvParams& p, const dh_packed_w* packed, bool separable,
                      int precision, cudaStream_t s) {
    (void)ctx; (void)p; (void)packed; (void)separable; (void)precision; (void)s;
    dh_set_error("dh_launch_conv_tc: not built");
    return -1;
}
