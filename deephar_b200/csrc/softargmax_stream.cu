// Streaming soft-argmax for large dense heat-maps (placeholder until the TMA-bulk pipeline lands).
#include "common.cuh"
bool dh_sam_stream_supported(const dh_view* h, int conf_on_prob, float alpha, bool has_d, bool has_prob) {
    (void)h; (void)conf_on_prob; (void)alpha; (void)has_d; (void)has_prob;
    return false;
}
int dh_sam_stream_launch(dh_ctx* ctx, const dh_view* h, int nj, int n_ctx, float alpha_mix,
                         float* out_pose, float* out_conf, void* stream) {
    (void)ctx; (void)h; (void)nj; (void)n_ctx; (void)alpha_mix; (void)out_pose; (void)out_conf; (void)stream;
    dh_set_error("dh_sam_stream_launch: not built");
    return -1;
}
