// Evaluation-time input pipeline on the GPU: the step BEFORE the forward path.
//   deephar/utils/transform.py:60-134   T.rotate_crop(angle 0) -> crop(integer box, zeros outside) ->
//                                       T.resize(crop_resolution, Image.BILINEAR) [-> horizontal_flip] -> asarray
//   deephar/utils/transform.py:212-231  normalize_channels: x / 255 [** chpower], (x - 0.5) * 2      (float32)
//   driven by deephar/data/mpii.py:91-122 (fixed evaluation config).
// `Image.resize(BILINEAR)` is Pillow's two-pass fixed-point resampler (libImaging/Resample.c): per output index a
// window of source pixels weighted by a triangle filter widened by the down-scaling factor, weights as 22-bit fixed
// point, horizontal pass -> uint8 -> vertical pass -> uint8.  The weight tables are computed on the host
// (deephar_b200/preprocess.py, double precision exactly as Pillow does); the kernels do the pixel work, bit-exact.
// Batched: one launch pair per batch of decoded uint8 frames of arbitrary sizes, output written straight into the
// (N, H, W, 3) fp32 NHWC input tensor of the network.
#include "common.cuh"

namespace {

constexpr int PREC = 22;

__device__ __forceinline__ int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// horizontal pass: tmp[n][r][xx][c], r over the crop rows
__global__ void resize_h_kernel(const dh_frame_src* __restrict__ frames, const int32_t* __restrict__ bounds,
                                const int32_t* __restrict__ coefs, int out_w, uint8_t* __restrict__ tmp,
                                int64_t tmp_stride) {
    const dh_frame_src f = frames[blockIdx.z];
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (xx >= out_w || r >= f.ch) return;
    const int first = bounds[f.kx_off + 2 * xx], n = bounds[f.kx_off + 2 * xx + 1];
    const int32_t* k = coefs + (int64_t)f.kx_coef_off + (int64_t)xx * f.ksx;
    int a0 = 1 << (PREC - 1), a1 = a0, a2 = a0;
    const int sy = f.y0 + r;
    if (sy >= 0 && sy < f.h) {
        const uint8_t* row = f.data + (int64_t)sy * f.stride;
        for (int t = 0; t < n; ++t) {
            const int sx = f.x0 + first + t;
            if (sx >= 0 && sx < f.w) {                      // Image.crop fills the outside with zeros
                const int kk = k[t];
                a0 += row[sx * 3 + 0] * kk;
                a1 += row[sx * 3 + 1] * kk;
                a2 += row[sx * 3 + 2] * kk;
            }
        }
    }
    uint8_t* o = tmp + blockIdx.z * tmp_stride + ((int64_t)r * out_w + xx) * 3;
    o[0] = (uint8_t)clip8(a0 >> PREC);
    o[1] = (uint8_t)clip8(a1 >> PREC);
    o[2] = (uint8_t)clip8(a2 >> PREC);
}

// vertical pass + flip + normalize_channels -> out[n][yy][xo][c] fp32
__global__ void resize_v_norm_kernel(const dh_frame_src* __restrict__ frames, const int32_t* __restrict__ bounds,
                                     const int32_t* __restrict__ coefs, int out_h, int out_w,
                                     const uint8_t* __restrict__ tmp, int64_t tmp_stride, float p0, float p1, float p2,
                                     float* __restrict__ out) {
    const dh_frame_src f = frames[blockIdx.z];
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy = blockIdx.y * blockDim.y + threadIdx.y;
    if (xx >= out_w || yy >= out_h) return;
    const int first = bounds[f.ky_off + 2 * yy], n = bounds[f.ky_off + 2 * yy + 1];
    const int32_t* k = coefs + (int64_t)f.ky_coef_off + (int64_t)yy * f.ksy;
    const uint8_t* col = tmp + blockIdx.z * tmp_stride + (int64_t)xx * 3;
    int a0 = 1 << (PREC - 1), a1 = a0, a2 = a0;
    for (int t = 0; t < n; ++t) {
        const uint8_t* px = col + (int64_t)(first + t) * out_w * 3;
        const int kk = k[t];
        a0 += px[0] * kk;
        a1 += px[1] * kk;
        a2 += px[2] * kk;
    }
    const int xo = f.hflip ? out_w - 1 - xx : xx;            // Image.transpose(FLIP_LEFT_RIGHT)
    float v[3] = {(float)clip8(a0 >> PREC), (float)clip8(a1 >> PREC), (float)clip8(a2 >> PREC)};
    const float pw[3] = {p0, p1, p2};
    float* o = out + (((int64_t)blockIdx.z * out_h + yy) * out_w + xo) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float t = __fdiv_rn(v[c], 255.f);                    // frame /= 255.   (float32, correctly rounded)
        if (pw[c] != 1.f) t = powf(t, pw[c]);
        o[c] = __fmul_rn(__fsub_rn(t, 0.5f), 2.f);           // frame -= .5 ; frame *= 2.
    }
}

}  // namespace

extern "C" int dh_crop_resize_norm_u8(dh_ctx* ctx, const dh_frame_src* frames_dev, int n, int max_crop_h,
                                      const int32_t* bounds_dev, const int32_t* coefs_dev, int out_h, int out_w,
                                      const float* chpower3, uint8_t* tmp_dev, int64_t tmp_stride, float* out_dev,
                                      void* stream) {
    DH_CHECK_ARG(ctx && frames_dev && bounds_dev && coefs_dev && tmp_dev && out_dev, "dh_crop_resize_norm_u8: NULL argument");
    DH_CHECK_ARG(n >= 0 && max_crop_h >= 1 && out_h >= 1 && out_w >= 1, "dh_crop_resize_norm_u8: bad sizes");
    DH_CHECK_ARG(tmp_stride >= (int64_t)max_crop_h * out_w * 3, "dh_crop_resize_norm_u8: tmp_stride too small");
    DH_CHECK_ARG(n <= 65535, "dh_crop_resize_norm_u8: at most 65535 frames per call");
    if (n == 0) return 0;
    const float p0 = chpower3 ? chpower3[0] : 1.f, p1 = chpower3 ? chpower3[1] : 1.f, p2 = chpower3 ? chpower3[2] : 1.f;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 block(32, 8);
    dim3 gh((out_w + 31) / 32, (max_crop_h + 7) / 8, n);
    resize_h_kernel<<<gh, block, 0, s>>>(frames_dev, bounds_dev, coefs_dev, out_w, tmp_dev, tmp_stride);
    dim3 gv((out_w + 31) / 32, (out_h + 7) / 8, n);
    resize_v_norm_kernel<<<gv, block, 0, s>>>(frames_dev, bounds_dev, coefs_dev, out_h, out_w, tmp_dev, tmp_stride, p0, p1, p2,
                                              out_dev);
    DH_LAUNCH_EPILOGUE(ctx, 2);
}
