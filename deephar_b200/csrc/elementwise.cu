// Pooling / resampling / n-ary add kernels (HBM-bound, coalesced over the NHWC channel axis).
// replaces: keras MaxPooling2D / UpSampling2D / add / Lambda layers of
// deephar/models/reception.py:74-127, models/common.py:70-108, models/spnet.py:98-146.
#include <float.h>
#include "common.cuh"

namespace {

struct PoolParams {
    const float* x; int N, H, W, C, ldx;
    float* out; int Ho, Wo, ldo;
    int kh, kw, sh, sw, pt, pl;
};

// mode 0: max ; mode 1: max + min  (layers.py:411-425 max_min_pooling = max(x) - max(-x))
template <int MODE>
__global__ void __launch_bounds__(256) pool_kernel(PoolParams p) {
    const int64_t total = (int64_t)p.N * p.Ho * p.Wo * p.C;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(idx % p.C);
        int64_t m = idx / p.C;
        int ox = (int)(m % p.Wo);
        int64_t t = m / p.Wo;
        int oy = (int)(t % p.Ho);
        int n = (int)(t / p.Ho);
        float mx = -FLT_MAX, mn = FLT_MAX;
        for (int ky = 0; ky < p.kh; ++ky) {
            int iy = oy * p.sh - p.pt + ky;
            if (iy < 0 || iy >= p.H) continue;
            for (int kx = 0; kx < p.kw; ++kx) {
                int ix = ox * p.sw - p.pl + kx;
                if (ix < 0 || ix >= p.W) continue;
                float v = __ldg(p.x + ((size_t)(n * p.H + iy) * p.W + ix) * p.ldx + c);
                mx = fmaxf(mx, v);
                mn = fminf(mn, v);
            }
        }
        p.out[(size_t)m * p.ldo + c] = MODE == 0 ? mx : mx + mn;
    }
}

struct UpParams {
    const float* a; int lda;
    const float* b; int ldb;
    float* out; int ldo;
    int N, H, W, C;  // output dims
};

__global__ void __launch_bounds__(256) upsample2x_add_kernel(UpParams p) {
    const int64_t total = (int64_t)p.N * p.H * p.W * p.C;
    const int Hb = p.H / 2, Wb = p.W / 2;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(idx % p.C);
        int64_t m = idx / p.C;
        int x = (int)(m % p.W);
        int64_t t = m / p.W;
        int y = (int)(t % p.H);
        int n = (int)(t / p.H);
        float v = __ldg(p.b + ((size_t)(n * Hb + (y >> 1)) * Wb + (x >> 1)) * p.ldb + c);
        if (p.a) v += __ldg(p.a + (size_t)m * p.lda + c);
        p.out[(size_t)m * p.ldo + c] = v;
    }
}

struct AddParams {
    const float* in[4]; int ld[4];
    int n_in;
    const float* scale; const float* shift; int relu;
    float* out; int ldo;
    int64_t M; int C;
};

__global__ void __launch_bounds__(256) add_n_kernel(AddParams p) {
    const int64_t total = p.M * p.C;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(idx % p.C);
        int64_t m = idx / p.C;
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < p.n_in) v += __ldg(p.in[i] + (size_t)m * p.ld[i] + c);
        if (p.scale) v = fmaf(v, __ldg(p.scale + c), __ldg(p.shift + c));
        if (p.relu) v = fmaxf(v, 0.f);
        p.out[(size_t)m * p.ldo + c] = v;
    }
}

// global (max + min) over (H,W) then softmax over channels: one CTA per batch item.
__global__ void __launch_bounds__(128) global_maxmin_softmax_kernel(const float* x, int H, int W, int C,
                                                                    int ld, float* out) {
    extern __shared__ float s[];  // C logits
    const int b = blockIdx.x;
    const float* xb = x + (size_t)b * H * W * ld;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float mx = -FLT_MAX, mn = FLT_MAX;
        for (int i = 0; i < H * W; ++i) {
            float v = __ldg(xb + (size_t)i * ld + c);
            mx = fmaxf(mx, v);
            mn = fminf(mn, v);
        }
        s[c] = mx + mn;
    }
    __syncthreads();
    float m = -FLT_MAX;
    for (int c = 0; c < C; ++c) m = fmaxf(m, s[c]);
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(s[c] - m);
    for (int c = threadIdx.x; c < C; c += blockDim.x) out[(size_t)b * C + c] = expf(s[c] - m) / sum;
}

__global__ void __launch_bounds__(256) mask_mul_kernel(const float* p, const float* c, int64_t rows,
                                                       int dim, float* out) {
    int64_t total = rows * dim;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x)
        out[idx] = __ldg(p + idx) * __ldg(c + idx / dim);
}

inline int grid_for(int64_t total, int num_sms) {
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)num_sms * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

static int pool_common(dh_ctx* ctx, const dh_view* x, int kh, int kw, int sh, int sw, int pad_same,
                       const dh_view* out, void* stream, int mode, const char* who) {
    DH_CHECK_ARG(ctx && x && out && x->p && out->p, "%s: NULL argument", who);
    PoolParams p;
    int pt, pl;
    int ho = dh_out_size(x->h, kh, sh, pad_same, &pt);
    int wo = dh_out_size(x->w, kw, sw, pad_same, &pl);
    DH_CHECK_ARG(out->n == x->n && out->h == ho && out->w == wo && out->c == x->c,
                 "%s: output view is (%d,%d,%d,%d), expected (%d,%d,%d,%d)", who, out->n, out->h, out->w,
                 out->c, x->n, ho, wo, x->c);
    p.x = x->p; p.N = x->n; p.H = x->h; p.W = x->w; p.C = x->c; p.ldx = x->ld;
    p.out = out->p; p.Ho = ho; p.Wo = wo; p.ldo = out->ld;
    p.kh = kh; p.kw = kw; p.sh = sh; p.sw = sw; p.pt = pt; p.pl = pl;
    int64_t total = (int64_t)p.N * ho * wo * p.C;
    if (mode == 0)
        pool_kernel<0><<<grid_for(total, ctx->num_sms), 256, 0, (cudaStream_t)stream>>>(p);
    else
        pool_kernel<1><<<grid_for(total, ctx->num_sms), 256, 0, (cudaStream_t)stream>>>(p);
    DH_LAUNCH_EPILOGUE(ctx, 1);
}

extern "C" int dh_maxpool2d_f32(dh_ctx* ctx, const dh_view* x, int kh, int kw, int sh, int sw,
                                int pad_same, const dh_view* out, void* stream) {
    return pool_common(ctx, x, kh, kw, sh, sw, pad_same, out, stream, 0, "dh_maxpool2d_f32");
}

extern "C" int dh_maxmin_pool2d_f32(dh_ctx* ctx, const dh_view* x, const dh_view* out, void* stream) {
    return pool_common(ctx, x, 2, 2, 2, 2, 1, out, stream, 1, "dh_maxmin_pool2d_f32");
}

extern "C" int dh_upsample2x_add_f32(dh_ctx* ctx, const dh_view* a, const dh_view* b,
                                     const dh_view* out, void* stream) {
    DH_CHECK_ARG(ctx && b && out && b->p && out->p, "dh_upsample2x_add_f32: NULL argument");
    DH_CHECK_ARG(out->n == b->n && out->h == 2 * b->h && out->w == 2 * b->w && out->c == b->c,
                 "dh_upsample2x_add_f32: output must be 2x the low-resolution input");
    UpParams p;
    p.a = nullptr; p.lda = 0;
    if (a && a->p) {
        DH_CHECK_ARG(a->n == out->n && a->h == out->h && a->w == out->w && a->c == out->c,
                     "dh_upsample2x_add_f32: `a` shape mismatch");
        p.a = a->p; p.lda = a->ld;
    }
    p.b = b->p; p.ldb = b->ld; p.out = out->p; p.ldo = out->ld;
    p.N = out->n; p.H = out->h; p.W = out->w; p.C = out->c;
    int64_t total = (int64_t)p.N * p.H * p.W * p.C;
    upsample2x_add_kernel<<<grid_for(total, ctx->num_sms), 256, 0, (cudaStream_t)stream>>>(p);
    DH_LAUNCH_EPILOGUE(ctx, 1);
}

extern "C" int dh_add_n_f32(dh_ctx* ctx, const dh_view* in, int n_in, const float* scale,
                            const float* shift, int relu, const dh_view* out, void* stream) {
    DH_CHECK_ARG(ctx && in && out && out->p, "dh_add_n_f32: NULL argument");
    DH_CHECK_ARG(n_in >= 1 && n_in <= 4, "dh_add_n_f32: n_in must be 1..4");
    DH_CHECK_ARG((scale == nullptr) == (shift == nullptr), "dh_add_n_f32: scale/shift must come together");
    AddParams p;
    for (int i = 0; i < 4; ++i) { p.in[i] = nullptr; p.ld[i] = 0; }
    for (int i = 0; i < n_in; ++i) {
        DH_CHECK_ARG(in[i].p && in[i].n == out->n && in[i].h == out->h && in[i].w == out->w &&
                         in[i].c == out->c,
                     "dh_add_n_f32: input %d shape mismatch", i);
        p.in[i] = in[i].p; p.ld[i] = in[i].ld;
    }
    p.n_in = n_in; p.scale = scale; p.shift = shift; p.relu = relu;
    p.out = out->p; p.ldo = out->ld;
    p.M = (int64_t)out->n * out->h * out->w; p.C = out->c;
    add_n_kernel<<<grid_for(p.M * p.C, ctx->num_sms), 256, 0, (cudaStream_t)stream>>>(p);
    DH_LAUNCH_EPILOGUE(ctx, 1);
}

extern "C" int dh_global_maxmin_softmax_f32(dh_ctx* ctx, const dh_view* x, float* out, void* stream) {
    DH_CHECK_ARG(ctx && x && x->p && out, "dh_global_maxmin_softmax_f32: NULL argument");
    DH_CHECK_ARG(x->c <= 4096, "dh_global_maxmin_softmax_f32: too many classes");
    global_maxmin_softmax_kernel<<<x->n, 128, x->c * sizeof(float), (cudaStream_t)stream>>>(
        x->p, x->h, x->w, x->c, x->ld, out);
    DH_LAUNCH_EPILOGUE(ctx, 1);
}

namespace {
__global__ void __launch_bounds__(256) zeropad_kernel(const float* x, int H, int W, int C, int ldx, int pt, int pl,
                                                      float* out, int Ho, int Wo, int ldo, int64_t total) {
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(idx % C);
        int64_t m = idx / C;
        int ox = (int)(m % Wo);
        int64_t t = m / Wo;
        int oy = (int)(t % Ho);
        int n = (int)(t / Ho);
        int iy = oy - pt, ix = ox - pl;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + ((size_t)(n * H + iy) * W + ix) * ldx + c);
        out[(size_t)m * ldo + c] = v;
    }
}
}  // namespace

// keras ZeroPadding2D(((top, bottom), (left, right))) (spnet.py:124-125,131-132)
extern "C" int dh_zeropad2d_f32(dh_ctx* ctx, const dh_view* x, int top, int left, const dh_view* out,
                                void* stream) {
    DH_CHECK_ARG(ctx && x && out && x->p && out->p, "dh_zeropad2d_f32: NULL argument");
    DH_CHECK_ARG(out->n == x->n && out->c == x->c && out->h >= x->h + top && out->w >= x->w + left && top >= 0 &&
                     left >= 0,
                 "dh_zeropad2d_f32: output (%d,%d) too small for input (%d,%d) + pad (%d,%d)", out->h, out->w, x->h,
                 x->w, top, left);
    int64_t total = (int64_t)out->n * out->h * out->w * out->c;
    zeropad_kernel<<<grid_for(total, ctx->num_sms), 256, 0, (cudaStream_t)stream>>>(
        x->p, x->h, x->w, x->c, x->ld, top, left, out->p, out->h, out->w, out->ld, total);
    DH_LAUNCH_EPILOGUE(ctx, 1);
}

extern "C" int dh_mask_mul_f32(dh_ctx* ctx, const float* p, const float* c, int64_t rows, int dim,
                               float* out, void* stream) {
    DH_CHECK_ARG(ctx && p && c && out, "dh_mask_mul_f32: NULL argument");
    mask_mul_kernel<<<grid_for(rows * dim, ctx->num_sms), 256, 0, (cudaStream_t)stream>>>(p, c, rows, dim, out);
    DH_LAUNCH_EPILOGUE(ctx, 1);
}
