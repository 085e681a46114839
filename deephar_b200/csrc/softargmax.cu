// Soft-argmax heads: spatial softmax + coordinate expectation + joint confidence in one
// pass over the heat-maps; the probability maps are never written back to HBM.
//
// replaces (reference, per prediction block): channel_softmax_2d (activations.py:3-16),
// the two fixed-weight SeparableConv2D(R x R, valid) "grid" convolutions + squeezes +
// concat of lin_interpolation_2d / softargmax2d (layers.py:122-129,160-200, grid from
// utils/math.py:6-19), AveragePooling2D*4 + GlobalMaxPooling2D of keypoint_confidence /
// build_joints_probability (layers.py:107-119, blocks.py:328-343), the context
// aggregation model (blocks.py:217-285), the depth expectation (spnet.py:201-205) and the
// volumetric marginal regression (reception.py:193-222).
//
// This file holds the general kernels (any H, W, C; one CTA per frame, frame staged in
// shared memory).  The streaming variant used for the large 32x32x48 reception maps is
// in softargmax_stream.cu.
#include <float.h>
#include "common.cuh"

namespace {

constexpr float K_EPSILON = 1e-7f;  // keras.backend.epsilon(), activations.py:12

struct SamParams {
    const float* h; int ldh;
    const float* d; int ldd;
    int N, H, W, C;
    float alpha;
    int conf_on_prob;
    float* out_pose; int pose_dim;
    float* out_conf;
    float* prob; int ldp;
    int nj, n_ctx; float alpha_mix;  // context aggregation when n_ctx > 0
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// np.linspace(0, 1, n) cast to float32 (utils/math.py:8-19): i * (1/(n-1)) in double, last = 1.
__device__ __forceinline__ void fill_grid(float* g, int n, int tid, int nthreads) {
    double step = n > 1 ? 1.0 / (double)(n - 1) : 0.0;
    for (int i = tid; i < n; i += nthreads) g[i] = (n > 1 && i == n - 1) ? 1.0f : (float)(i * step);
}

// Per-channel soft-argmax statistics of an smem-resident frame s_h[P][C].
// On return (after the trailing __syncthreads) s_res holds, per channel c:
//   s_res[0*C+c] = E[x], [1*C+c] = E[y], [2*C+c] = conf, [3*C+c] = z, [4*C+c] = clipped denominator
// and s_h holds exp(alpha*h - max).  d (global, may be NULL) is the depth map of this frame.
__device__ void sam_stats(float* s_h, const float* s_gx, const float* s_gy, float* s_red, float* s_res,
                          int H, int W, int C, float alpha, int conf_on_prob, const float* d, int ldd) {
    const int tid = threadIdx.x, T = blockDim.x;
    const int P = H * W;
    const int parts = T / C;
    const bool active = tid < parts * C;
    const int c = tid % C, part = tid / C;

    float m = -FLT_MAX, wv = -FLT_MAX;
    if (active) {
        for (int pix = part; pix < P; pix += parts) {
            float v = s_h[pix * C + c];
            m = fmaxf(m, alpha * v);
            if (!conf_on_prob) {
                int row = pix / W, col = pix - row * W;
                if (row + 1 < H && col + 1 < W)
                    wv = fmaxf(wv, (v + s_h[(pix + 1) * C + c]) + (s_h[(pix + W) * C + c] + s_h[(pix + W + 1) * C + c]));
            }
        }
        s_red[tid] = m;
    }
    __syncthreads();
    if (active) {
        m = -FLT_MAX;
        for (int q = 0; q < parts; ++q) m = fmaxf(m, s_red[q * C + c]);
    }
    __syncthreads();

    float s = 0.f, sx = 0.f, sy = 0.f, z = 0.f;
    if (active) {
        for (int pix = part; pix < P; pix += parts) {
            int row = pix / W, col = pix - row * W;
            float e = expf(alpha * s_h[pix * C + c] - m);
            s_h[pix * C + c] = e;
            s += e;
            sx = fmaf(e, s_gx[col], sx);
            sy = fmaf(e, s_gy[row], sy);
            if (d) z = fmaf(e, sigmoidf_(__ldg(d + (size_t)pix * ldd + c)), z);
        }
    }
    __syncthreads();
    if (active && conf_on_prob) {
        for (int pix = part; pix < P; pix += parts) {
            int row = pix / W, col = pix - row * W;
            if (row + 1 < H && col + 1 < W)
                wv = fmaxf(wv, (s_h[pix * C + c] + s_h[(pix + 1) * C + c]) +
                                   (s_h[(pix + W) * C + c] + s_h[(pix + W + 1) * C + c]));
        }
    }
    if (active) {
        s_red[0 * T + tid] = s;
        s_red[1 * T + tid] = sx;
        s_red[2 * T + tid] = sy;
        s_red[3 * T + tid] = z;
        s_red[4 * T + tid] = wv;
    }
    __syncthreads();
    if (tid < C) {
        float S = 0.f, SX = 0.f, SY = 0.f, Z = 0.f, Wm = -FLT_MAX;
        for (int q = 0; q < parts; ++q) {
            int i = q * C + tid;
            S += s_red[0 * T + i];
            SX += s_red[1 * T + i];
            SY += s_red[2 * T + i];
            Z += s_red[3 * T + i];
            Wm = fmaxf(Wm, s_red[4 * T + i]);
        }
        float den = fmaxf(S, K_EPSILON);
        s_res[0 * C + tid] = SX / den;
        s_res[1 * C + tid] = SY / den;
        s_res[2 * C + tid] = conf_on_prob ? Wm / den : Wm;
        s_res[3 * C + tid] = Z / den;
        s_res[4 * C + tid] = den;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(512) softargmax2d_kernel(SamParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, T = blockDim.x;
    const int P = p.H * p.W, C = p.C;
    const int PC4 = (P * C + 3) & ~3;
    float* s_h = smem;
    float* s_gx = s_h + PC4;
    float* s_gy = s_gx + p.W;
    float* s_red = s_gy + p.H;
    float* s_res = s_red + 5 * T;
    const int n = blockIdx.x;

    fill_grid(s_gx, p.W, tid, T);
    fill_grid(s_gy, p.H, tid, T);
    const float* hb = p.h + (size_t)n * P * p.ldh;
    if ((C & 3) == 0 && (p.ldh & 3) == 0 && ((uintptr_t)p.h & 15) == 0) {
        const int C4 = C >> 2;
        for (int i = tid; i < P * C4; i += T) {
            int pix = i / C4, q = i - pix * C4;
            float4 v = __ldg(reinterpret_cast<const float4*>(hb + (size_t)pix * p.ldh) + q);
            reinterpret_cast<float4*>(s_h)[i] = v;
        }
    } else {
        for (int i = tid; i < P * C; i += T) {
            int pix = i / C, c = i - pix * C;
            s_h[i] = __ldg(hb + (size_t)pix * p.ldh + c);
        }
    }
    __syncthreads();

    const float* db = p.d ? p.d + (size_t)n * P * p.ldd : nullptr;
    sam_stats(s_h, s_gx, s_gy, s_red, s_res, p.H, p.W, C, p.alpha, p.conf_on_prob, db, p.ldd);

    if (p.n_ctx > 0) {
        // blocks.py:217-285: channels [0,nj) specialised, then nj groups of n_ctx context maps.
        if (tid < p.nj) {
            float pcs = 0.f, px = 0.f, py = 0.f;
            for (int i = 0; i < p.n_ctx; ++i) {
                int cc = p.nj + tid * p.n_ctx + i;
                float pc = s_res[2 * C + cc];
                pcs += pc;
                px = fmaf(s_res[0 * C + cc], pc, px);
                py = fmaf(s_res[1 * C + cc], pc, py);
            }
            float a = p.alpha_mix;
            p.out_pose[((size_t)n * p.nj + tid) * 2 + 0] = a * s_res[0 * C + tid] + (1.f - a) * (px / pcs);
            p.out_pose[((size_t)n * p.nj + tid) * 2 + 1] = a * s_res[1 * C + tid] + (1.f - a) * (py / pcs);
            p.out_conf[(size_t)n * p.nj + tid] = s_res[2 * C + tid];
        }
        return;
    }
    if (tid < C) {
        float* o = p.out_pose + ((size_t)n * C + tid) * p.pose_dim;
        o[0] = s_res[0 * C + tid];
        o[1] = s_res[1 * C + tid];
        if (p.pose_dim == 3) o[2] = s_res[3 * C + tid];
        p.out_conf[(size_t)n * C + tid] = s_res[2 * C + tid];
    }
    if (p.prob) {
        float* pb = p.prob + (size_t)n * P * p.ldp;
        for (int i = tid; i < P * C; i += T) {
            int pix = i / C, c = i - pix * C;
            pb[(size_t)pix * p.ldp + c] = s_h[i] / s_res[4 * C + c];
        }
    }
}

// ---------------------------------------------------------------------------
// Volumetric (reception 3-D) head: reception.py:193-222.
// h (N,H,W,D*nj) with channel = d*nj + j is streamed ONCE through shared memory in
// chunks of PCH pixels; both marginals (mean over d -> hxy, mean over hw -> hz) are
// accumulated on the fly, then the 2-D / 1-D soft-argmax run on the smem-resident
// marginals.
// ---------------------------------------------------------------------------
constexpr int PCH = 32;

struct Sam3dParams {
    const float* h; int ldh;
    int N, H, W, nj, D;
    float* out_pose; float* out_vis;
    float vis_scale;            // visible = sigmoid(vis_scale * (max hxy + max hz)): 1 reception.py:217-220, 2 action.py:291-292
    float* prob; int ldp;       // optional: channel_softmax_2d(hxy) (N,H,W,nj) for the kronecker product (action.py:294-295)
};

__global__ void __launch_bounds__(512) softargmax3d_kernel(Sam3dParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, T = blockDim.x;
    const int P = p.H * p.W, nj = p.nj, D = p.D, C = nj * D;
    const int PJ4 = (P * nj + 3) & ~3;
    float* s_hxy = smem;                 // [P][nj]
    float* s_chunk = s_hxy + PJ4;        // [PCH][C]
    float* s_hz = s_chunk + PCH * C;     // [C]
    float* s_gx = s_hz + ((C + 3) & ~3);
    float* s_gy = s_gx + p.W;
    float* s_red = s_gy + p.H;
    float* s_res = s_red + 5 * T;
    const int n = blockIdx.x;

    fill_grid(s_gx, p.W, tid, T);
    fill_grid(s_gy, p.H, tid, T);
    const float* hb = p.h + (size_t)n * P * p.ldh;
    const bool vec = (C & 3) == 0 && (p.ldh & 3) == 0 && ((uintptr_t)p.h & 15) == 0;
    const int nch_slots = (C + T - 1) / T;  // channels per thread for the hz accumulation (<= 2)
    float hz_acc[2] = {0.f, 0.f};

    for (int p0 = 0; p0 < P; p0 += PCH) {
        const int np = min(PCH, P - p0);
        if (vec) {
            const int C4 = C >> 2;
            for (int i = tid; i < np * C4; i += T) {
                int pl = i / C4, q = i - pl * C4;
                reinterpret_cast<float4*>(s_chunk)[i] =
                    __ldg(reinterpret_cast<const float4*>(hb + (size_t)(p0 + pl) * p.ldh) + q);
            }
        } else {
            for (int i = tid; i < np * C; i += T) {
                int pl = i / C, c = i - pl * C;
                s_chunk[i] = __ldg(hb + (size_t)(p0 + pl) * p.ldh + c);
            }
        }
        __syncthreads();
        for (int sl = 0; sl < nch_slots; ++sl) {
            int ch = tid + sl * T;
            if (ch < C) {
                float a = 0.f;
                for (int pl = 0; pl < np; ++pl) a += s_chunk[pl * C + ch];
                hz_acc[sl] += a;
            }
        }
        for (int i = tid; i < np * nj; i += T) {
            int pl = i / nj, j = i - pl * nj;
            float a = 0.f;
            for (int dd = 0; dd < D; ++dd) a += s_chunk[pl * C + dd * nj + j];
            s_hxy[(p0 + pl) * nj + j] = a / (float)D;
        }
        __syncthreads();
    }
    for (int sl = 0; sl < nch_slots; ++sl) {
        int ch = tid + sl * T;
        if (ch < C) s_hz[ch] = hz_acc[sl] / (float)P;
    }
    __syncthreads();

    // vxy = max over pixels of hxy, before sam_stats overwrites s_hxy with exponentials:
    // sam_stats' "raw" confidence slot is not used here, so compute the max directly.
    float vmax = -FLT_MAX;
    {
        const int parts = T / nj;
        if (tid < parts * nj) {
            int c = tid % nj, part = tid / nj;
            for (int pix = part; pix < P; pix += parts) vmax = fmaxf(vmax, s_hxy[pix * nj + c]);
            s_red[tid] = vmax;
        }
        __syncthreads();
        if (tid < nj) {
            vmax = -FLT_MAX;
            for (int q = 0; q < parts; ++q) vmax = fmaxf(vmax, s_red[q * nj + tid]);
        }
        __syncthreads();
    }
    sam_stats(s_hxy, s_gx, s_gy, s_red, s_res, p.H, p.W, nj, 1.0f, 1, nullptr, 0);

    if (tid < nj) {
        // zSAM: blocks.py:288-303 -- softmax over depth, grid (k+0.5)/D (layers.py:141-146)
        float zm = -FLT_MAX;
        for (int dd = 0; dd < D; ++dd) zm = fmaxf(zm, s_hz[dd * nj + tid]);
        double start = 1.0 / (2.0 * D), step = D > 1 ? ((1.0 - start) - start) / (double)(D - 1) : 0.0;
        float zs = 0.f, ze = 0.f;
        for (int dd = 0; dd < D; ++dd) {
            float e = expf(s_hz[dd * nj + tid] - zm);
            float g = (D > 1 && dd == D - 1) ? (float)(1.0 - start) : (float)(dd * step + start);
            zs += e;
            ze = fmaf(e, g, ze);
        }
        float* o = p.out_pose + ((size_t)n * nj + tid) * 3;
        o[0] = s_res[0 * nj + tid];
        o[1] = s_res[1 * nj + tid];
        o[2] = ze / zs;
        p.out_vis[(size_t)n * nj + tid] = sigmoidf_(p.vis_scale * (vmax + zm));
    }
    if (p.prob) {
        float* pb = p.prob + (size_t)n * P * p.ldp;
        for (int i = tid; i < P * nj; i += T) {
            int pix = i / nj, c = i - pix * nj;
            pb[(size_t)pix * p.ldp + c] = s_hxy[i] / s_res[4 * nj + c];
        }
    }
}

// layers.py:478-508: out[n,j,f] = sum_p P[n,p,j] * Z[n,p,f].  grid (N, ceil(F/128)), 128 threads;
// P is staged in shared memory in chunks of 64 pixels, each thread owns one feature f.
constexpr int KR_MAXJ = 32, KR_PCH = 64;
__global__ void __launch_bounds__(128) kron_kernel(const float* pm, int ldpm, const float* z, int ldz,
                                                   int P, int nj, int F, float* out) {
    __shared__ float s_p[KR_PCH * KR_MAXJ];
    const int n = blockIdx.x;
    const int f = blockIdx.y * 128 + threadIdx.x;
    float acc[KR_MAXJ];
#pragma unroll
    for (int j = 0; j < KR_MAXJ; ++j) acc[j] = 0.f;
    const float* pb = pm + (size_t)n * P * ldpm;
    const float* zb = z + (size_t)n * P * ldz;
    for (int p0 = 0; p0 < P; p0 += KR_PCH) {
        int np = min(KR_PCH, P - p0);
        __syncthreads();
        for (int i = threadIdx.x; i < np * nj; i += 128) {
            int pl = i / nj, j = i - pl * nj;
            s_p[pl * KR_MAXJ + j] = __ldg(pb + (size_t)(p0 + pl) * ldpm + j);
        }
        __syncthreads();
        if (f < F) {
            for (int pl = 0; pl < np; ++pl) {
                float zv = __ldg(zb + (size_t)(p0 + pl) * ldz + f);
#pragma unroll
                for (int j = 0; j < KR_MAXJ; ++j)
                    if (j < nj) acc[j] = fmaf(s_p[pl * KR_MAXJ + j], zv, acc[j]);
            }
        }
    }
    if (f < F) {
#pragma unroll
        for (int j = 0; j < KR_MAXJ; ++j)
            if (j < nj) out[((size_t)n * nj + j) * F + f] = acc[j];
    }
}

int launch_sam(dh_ctx* ctx, SamParams& p, void* stream, const char* who) {
    const int P = p.H * p.W;
    DH_CHECK_ARG(p.C >= 1 && p.C <= 512, "%s: C=%d not in 1..512", who, p.C);
    // the confidence is a max over 2x2 windows (AveragePooling2D((2,2), valid) in the reference raises on smaller maps)
    DH_CHECK_ARG(p.H >= 2 && p.W >= 2, "%s: maps must be at least 2x2 (got %dx%d)", who, p.H, p.W);
    int T = (P * p.C <= 4096) ? 256 : 512;
    if (T < p.C) T = 512;
    size_t smem = (size_t)(((P * p.C + 3) & ~3) + p.W + p.H + 5 * T + 5 * p.C) * sizeof(float);
    DH_CHECK_ARG(smem <= 227 * 1024, "%s: frame of %d x %d x %d floats does not fit shared memory", who, p.H, p.W, p.C);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(softargmax2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { dh_set_error("%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e)); return (int)e; }
    }
    softargmax2d_kernel<<<p.N, T, smem, (cudaStream_t)stream>>>(p);
    DH_LAUNCH_EPILOGUE(ctx, 1);
}

}  // namespace

// streaming kernel for large dense maps (softargmax_stream.cu); returns false if it does not apply
bool dh_sam_stream_supported(const dh_view* h, int conf_on_prob, float alpha, bool has_d, bool has_prob);
int dh_sam_stream_launch(dh_ctx* ctx, const dh_view* h, int nj, int n_ctx, float alpha_mix,
                         float* out_pose, float* out_conf, void* stream);

extern "C" int dh_softargmax2d_f32(dh_ctx* ctx, const dh_view* h, const dh_view* d, float alpha,
                                   int conf_on_prob, float* out_pose, float* out_conf,
                                   const dh_view* prob_out, void* stream) {
    DH_CHECK_ARG(ctx && h && h->p && out_pose && out_conf, "dh_softargmax2d_f32: NULL argument");
    SamParams p;
    p.h = h->p; p.ldh = h->ld; p.N = h->n; p.H = h->h; p.W = h->w; p.C = h->c;
    p.d = nullptr; p.ldd = 0;
    if (d && d->p) {
        DH_CHECK_ARG(d->n == h->n && d->h == h->h && d->w == h->w && d->c == h->c,
                     "dh_softargmax2d_f32: depth map shape mismatch");
        p.d = d->p; p.ldd = d->ld;
    }
    p.alpha = alpha; p.conf_on_prob = conf_on_prob;
    p.out_pose = out_pose; p.pose_dim = p.d ? 3 : 2; p.out_conf = out_conf;
    p.prob = nullptr; p.ldp = 0;
    if (prob_out && prob_out->p) {
        DH_CHECK_ARG(prob_out->n == h->n && prob_out->h == h->h && prob_out->w == h->w && prob_out->c == h->c,
                     "dh_softargmax2d_f32: prob_out shape mismatch");
        p.prob = prob_out->p; p.ldp = prob_out->ld;
    }
    p.nj = 0; p.n_ctx = 0; p.alpha_mix = 0.f;
    if (dh_sam_stream_supported(h, conf_on_prob, alpha, p.d != nullptr, p.prob != nullptr))
        return dh_sam_stream_launch(ctx, h, 0, 0, 0.f, out_pose, out_conf, stream);
    return launch_sam(ctx, p, stream, "dh_softargmax2d_f32");
}

extern "C" int dh_softargmax2d_ctx_f32(dh_ctx* ctx, const dh_view* h, int nj, int n_ctx,
                                       float alpha_mix, float* out_pose, float* out_vis, void* stream) {
    DH_CHECK_ARG(ctx && h && h->p && out_pose && out_vis, "dh_softargmax2d_ctx_f32: NULL argument");
    DH_CHECK_ARG(nj >= 1 && n_ctx >= 1 && h->c == nj * (1 + n_ctx),
                 "dh_softargmax2d_ctx_f32: C=%d is not nj*(1+n_ctx) = %d*(1+%d)", h->c, nj, n_ctx);
    if (dh_sam_stream_supported(h, 0, 1.0f, false, false))
        return dh_sam_stream_launch(ctx, h, nj, n_ctx, alpha_mix, out_pose, out_vis, stream);
    SamParams p;
    p.h = h->p; p.ldh = h->ld; p.N = h->n; p.H = h->h; p.W = h->w; p.C = h->c;
    p.d = nullptr; p.ldd = 0; p.alpha = 1.0f; p.conf_on_prob = 0;
    p.out_pose = out_pose; p.pose_dim = 2; p.out_conf = out_vis; p.prob = nullptr; p.ldp = 0;
    p.nj = nj; p.n_ctx = n_ctx; p.alpha_mix = alpha_mix;
    return launch_sam(ctx, p, stream, "dh_softargmax2d_ctx_f32");
}

bool dh_sam3d_stream_supported(const dh_view* h, int nj, int depth_maps);
int dh_sam3d_stream_launch(dh_ctx* ctx, const dh_view* h, int nj, int depth_maps, float vis_scale, float* out_pose,
                           float* out_vis, void* stream);

static int launch_sam3d(dh_ctx* ctx, const dh_view* h, int nj, int depth_maps, float vis_scale, float* out_pose,
                        float* out_vis, const dh_view* prob_out, void* stream, const char* who) {
    DH_CHECK_ARG(ctx && h && h->p && out_pose && out_vis, "%s: NULL argument", who);
    DH_CHECK_ARG(nj >= 1 && depth_maps >= 1 && h->c == nj * depth_maps,
                 "%s: C=%d is not depth_maps*nj = %d*%d", who, h->c, depth_maps, nj);
    DH_CHECK_ARG(h->h >= 2 && h->w >= 2, "%s: maps must be at least 2x2 (got %dx%d)", who, h->h, h->w);
    // large dense volumes: the cluster-split streaming kernel (softargmax_stream.cu); the probability export of the
    // merge model and odd shapes stay on the staged kernel below
    if (!(prob_out && prob_out->p) && ctx->sam3d_stream && dh_sam3d_stream_supported(h, nj, depth_maps))
        return dh_sam3d_stream_launch(ctx, h, nj, depth_maps, vis_scale, out_pose, out_vis, stream);
    const int T = 512, C = h->c, P = h->h * h->w;
    DH_CHECK_ARG(C <= 2 * T && nj <= T, "%s: too many channels", who);
    Sam3dParams p;
    p.h = h->p; p.ldh = h->ld; p.N = h->n; p.H = h->h; p.W = h->w; p.nj = nj; p.D = depth_maps;
    p.out_pose = out_pose; p.out_vis = out_vis;
    p.vis_scale = vis_scale;
    p.prob = nullptr; p.ldp = 0;
    if (prob_out && prob_out->p) {
        DH_CHECK_ARG(prob_out->n == h->n && prob_out->h == h->h && prob_out->w == h->w && prob_out->c == nj,
                     "%s: prob_out must be (N,H,W,nj)", who);
        p.prob = prob_out->p; p.ldp = prob_out->ld;
    }
    size_t smem = (size_t)(((P * nj + 3) & ~3) + PCH * C + ((C + 3) & ~3) + h->w + h->h + 5 * T + 5 * nj) * sizeof(float);
    DH_CHECK_ARG(smem <= 227 * 1024, "%s: marginal maps do not fit shared memory", who);
    if (smem > 48 * 1024) {
        static size_t cur = 0;
        if (smem > cur) {
            cudaError_t e = cudaFuncSetAttribute(softargmax3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { dh_set_error("%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e)); return (int)e; }
            cur = smem;
        }
    }
    softargmax3d_kernel<<<p.N, T, smem, (cudaStream_t)stream>>>(p);
    DH_LAUNCH_EPILOGUE(ctx, 1);
}

extern "C" int dh_softargmax3d_f32(dh_ctx* ctx, const dh_view* h, int nj, int depth_maps,
                                   float* out_pose, float* out_vis, void* stream) {
    return launch_sam3d(ctx, h, nj, depth_maps, 1.0f, out_pose, out_vis, nullptr, stream, "dh_softargmax3d_f32");
}

extern "C" int dh_softargmax3d_ex_f32(dh_ctx* ctx, const dh_view* h, int nj, int depth_maps, float vis_scale,
                                      float* out_pose, float* out_vis, const dh_view* prob_out, void* stream) {
    return launch_sam3d(ctx, h, nj, depth_maps, vis_scale, out_pose, out_vis, prob_out, stream, "dh_softargmax3d_ex_f32");
}

extern "C" int dh_kron_pool_f32(dh_ctx* ctx, const dh_view* pm, const dh_view* z, float* out, void* stream) {
    DH_CHECK_ARG(ctx && pm && z && pm->p && z->p && out, "dh_kron_pool_f32: NULL argument");
    DH_CHECK_ARG(pm->n == z->n && pm->h == z->h && pm->w == z->w, "dh_kron_pool_f32: P and Z spatial shapes differ");
    DH_CHECK_ARG(pm->c <= KR_MAXJ, "dh_kron_pool_f32: more than %d joints", KR_MAXJ);
    dim3 grid(pm->n, (z->c + 127) / 128);
    kron_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(pm->p, pm->ld, z->p, z->ld, pm->h * pm->w, pm->c, z->c, out);
    DH_LAUNCH_EPILOGUE(ctx, 1);
}
