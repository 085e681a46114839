// Micro-benchmark: tcgen05.mma kind::f16 issue/execute rate for M=128, K=16 and various N, SS mode,
// operands in (garbage) shared memory.  One CTA per SM, one issuing thread.  Prints cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, int sbo, int layout) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
           ((uint64_t)layout << 61);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: 128B swizzle (BK=64), mode 1: 64B swizzle (BK=32).  reps K-blocks; per K-block nsub * ksteps * passes MMAs
__global__ void __launch_bounds__(128, 1) k(int N, int nsub, int mode, int passes, int reps, long long* out, int same_ab) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tslot;
    const int rowb = mode == 0 ? 128 : 64;
    const int sbo = 8 * rowb, layout = mode == 0 ? 2 : 4;
    const int ksteps = rowb / 32;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_hi = smem_u32(smem), a_lo = a_hi + 128 * rowb;
    const uint32_t b_hi = a_lo + 128 * rowb, b_lo = b_hi + nsub * N * rowb;
    long long t0 = 0, t1 = 0;
    if (warp == 1 && lane == 0) {
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            for (int sub = 0; sub < nsub; ++sub) {
                const uint32_t d = tb + sub * ((N + 31) / 32 * 32);
                const uint32_t bo = sub * N * rowb;
                for (int kk = 0; kk < ksteps; ++kk) {
                    const uint32_t ko = kk * 32;
                    umma(d, desc(a_hi + ko, sbo, layout), desc(b_hi + bo + ko, sbo, layout), idesc, (r | kk) ? 1u : 0u);
                    if (passes == 3) {
                        umma(d, desc(a_lo + ko, sbo, layout), desc((same_ab ? b_hi : b_hi) + bo + ko, sbo, layout), idesc, 1u);
                        umma(d, desc(a_hi + ko, sbo, layout), desc(b_lo + bo + ko, sbo, layout), idesc, 1u);
                    }
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int ns[] = {64, 96, 128, 144, 160, 192, 256};
    for (int mode = 0; mode < 2; ++mode)
        for (int passes = 1; passes <= 3; passes += 2)
            for (int N : ns)
                for (int nsub = 1; nsub <= 2; ++nsub) {
                    if (nsub * ((N + 31) / 32 * 32) > 512) continue;
                    const int reps = 200;
                    k<<<148, 128, 200 * 1024>>>(N, nsub, mode, passes, reps, d, 0);
                    long long c = 0;
                    cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    const int ksteps = mode == 0 ? 4 : 2;
                    const double n_mma = (double)reps * nsub * ksteps * passes;
                    printf("swizzle%-3d passes %d N %3d nsub %d: %7.1f cyc/MMA (floor %5.1f)  eff %.2f\n", mode == 0 ? 128 : 64,
                           passes, N, nsub, c / n_mma, 128.0 * N / 256, (128.0 * N / 256) / (c / n_mma));
                }
    return 0;
}
