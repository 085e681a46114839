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
// VEC = 4: one thread handles 4 consecutive channels with 16-byte loads/stores (C, ld % 4 == 0).
template <int MODE, int VEC>
__global__ void __launch_bounds__(256) pool_kernel(PoolParams p) {
    const int CV = p.C / VEC;
    const int64_t total = (int64_t)p.N * p.Ho * p.Wo * CV;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(idx % CV) * VEC;
        int64_t m = idx / CV;
        int ox = (int)(m % p.Wo);
        int64_t t = m / p.Wo;
        int oy = (int)(t % p.Ho);
        int n = (int)(t / p.Ho);
        float mx[VEC], mn[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) { mx[e] = -FLT_MAX; mn[e] = FLT_MAX; }
        for (int ky = 0; ky < p.kh; ++ky) {
            int iy = oy * p.sh - p.pt + ky;
            if (iy < 0 || iy >= p.H) continue;
            for (int kx = 0; kx < p.kw; ++kx) {
                int ix = ox * p.sw - p.pl + kx;
                if (ix < 0 || ix >= p.W) continue;
                const float* src = p.x + ((size_t)(n * p.H + iy) * p.W + ix) * p.ldx + c;
                float v[VEC];
                if (VEC == 4) {
                    float4 q = __ldg(reinterpret_cast<const float4*>(src));
                    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                } else {
                    v[0] = __ldg(src);
                }
#pragma unroll
                for (int e = 0; e < VEC; ++e) { mx[e] = fmaxf(mx[e], v[e]); mn[e] = fminf(mn[e], v[e]); }
            }
        }
        float* dst = p.out + (size_t)m * p.ldo + c;
        if (VEC == 4) {
            float4 o = MODE == 0 ? make_float4(mx[0], mx[1], mx[2], mx[3])
                                 : make_float4(mx[0] + mn[0], mx[1] + mn[1], mx[2] + mn[2], mx[3] + mn[3]);
            *reinterpret_cast<float4*>(dst) = o;
        } else {
            dst[0] = MODE == 0 ? mx[0] : mx[0] + mn[0];
        }
    }
}

struct UpParams {
    const float* a; int lda;
    const float* b; int ldb;
    float* out; int ldo;
    int N, H, W, C;  // output dims
};

template <int VEC>
__global__ void __launch_bounds__(256) upsample2x_add_kernel(UpParams p) {
    const int CV = p.C / VEC;
    const int64_t total = (int64_t)p.N * p.H * p.W * CV;
    const int Hb = p.H / 2, Wb = p.W / 2;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(idx % CV) * VEC;
        int64_t m = idx / CV;
        int x = (int)(m % p.W);
        int64_t t = m / p.W;
        int y = (int)(t % p.H);
        int n = (int)(t / p.H);
        const float* bp = p.b + ((size_t)(n * Hb + (y >> 1)) * Wb + (x >> 1)) * p.ldb + c;
        float* dst = p.out + (size_t)m * p.ldo + c;
        if (VEC == 4) {
            float4 v = __ldg(reinterpret_cast<const float4*>(bp));
            if (p.a) {
                float4 a = __ldg(reinterpret_cast<const float4*>(p.a + (size_t)m * p.lda + c));
                v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
            }
            *reinterpret_cast<float4*>(dst) = v;
        } else {
            float v = __ldg(bp);
            if (p.a) v += __ldg(p.a + (size_t)m * p.lda + c);
            dst[0] = v;
        }
    }
}

struct AddParams {
    const float* in[4]; int ld[4];
    int n_in;
    const float* scale; const float* shift; int relu;
    float* out; int ldo;
    int64_t M; int C;
};

template <int VEC>
__global__ void __launch_bounds__(256) add_n_kernel(AddParams p) {
    const int CV = p.C / VEC;
    const int64_t total = p.M * CV;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(idx % CV) * VEC;
        int64_t m = idx / CV;
        float v[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[e] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < p.n_in) {
                const float* src = p.in[i] + (size_t)m * p.ld[i] + c;
                if (VEC == 4) {
                    float4 q = __ldg(reinterpret_cast<const float4*>(src));
                    v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w;
                } else {
                    v[0] += __ldg(src);
                }
            }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            if (p.scale) v[e] = fmaf(v[e], __ldg(p.scale + c + e), __ldg(p.shift + c + e));
            if (p.relu) v[e] = fmaxf(v[e], 0.f);
        }
        float* dst = p.out + (size_t)m * p.ldo + c;
        if (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        else dst[0] = v[0];
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
    const bool vec = (p.C % 4 == 0) && (p.ldx % 4 == 0) && (p.ldo % 4 == 0) && dh_aligned16(p.x) && dh_aligned16(p.out);
    int64_t total = (int64_t)p.N * ho * wo * (vec ? p.C / 4 : p.C);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0 && vec) pool_kernel<0, 4><<<grid_for(total, ctx->num_sms), 256, 0, st>>>(p);
    else if (mode == 0) pool_kernel<0, 1><<<grid_for(total, ctx->num_sms), 256, 0, st>>>(p);
    else if (vec) pool_kernel<1, 4><<<grid_for(total, ctx->num_sms), 256, 0, st>>>(p);
    else pool_kernel<1, 1><<<grid_for(total, ctx->num_sms), 256, 0, st>>>(p);
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
    const bool vec = (p.C % 4 == 0) && (p.ldb % 4 == 0) && (p.ldo % 4 == 0) && dh_aligned16(p.b) && dh_aligned16(p.out) &&
                     (!p.a || ((p.lda % 4 == 0) && dh_aligned16(p.a)));
    int64_t total = (int64_t)p.N * p.H * p.W * (vec ? p.C / 4 : p.C);
    if (vec) upsample2x_add_kernel<4><<<grid_for(total, ctx->num_sms), 256, 0, (cudaStream_t)stream>>>(p);
    else upsample2x_add_kernel<1><<<grid_for(total, ctx->num_sms), 256, 0, (cudaStream_t)stream>>>(p);
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
    bool vec = (p.C % 4 == 0) && (p.ldo % 4 == 0) && dh_aligned16(p.out) &&
               (!scale || (dh_aligned16(scale) && dh_aligned16(shift)));
    for (int i = 0; i < n_in; ++i) vec = vec && (p.ld[i] % 4 == 0) && dh_aligned16(p.in[i]);
    if (vec) add_n_kernel<4><<<grid_for(p.M * (p.C / 4), ctx->num_sms), 256, 0, (cudaStream_t)stream>>>(p);
    else add_n_kernel<1><<<grid_for(p.M * p.C, ctx->num_sms), 256, 0, (cudaStream_t)stream>>>(p);
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
