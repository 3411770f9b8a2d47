// tcgen05.mma issue-rate probe (timing only, operand contents are irrelevant): cycles per instruction for the shapes,
// operand sources and shared-memory layouts the loss/grad kernel can choose from.  One CTA per SM on `grid` SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_probe3 tools/tc_probe3.cu && tools/tc_probe3
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}
// kind: 0 tf32 (a/b format 2), 1 bf16 (a/b format 1, kind::f16)
__device__ __forceinline__ uint32_t make_idesc(int kind, int M, int N, int a_mn, int b_mn) {
    const uint32_t fmt = kind == 0 ? 2u : 1u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
template <int KIND, bool TS>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0) {
        if (TS)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                         "r"((uint32_t)a), "l"(b), "r"(idesc), "r"(acc) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                         "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    } else {
        if (TS)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                         "r"((uint32_t)a), "l"(b), "r"(idesc), "r"(acc) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                         "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    }
}

// LAYOUT: 0 = K-major no swizzle (core matrices, LBO 128 / SBO 2048), 1 = MN-major SWIZZLE_128B_BASE32B (the dW images),
//         2 = K-major SWIZZLE_128B (128-byte rows, SBO 1024, K step = +32 B)
// NACC: number of distinct accumulators cycled through (dependent-chain test)
template <int KIND, bool TS, int M, int N, int LAYOUT, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    for (int i = tid; i < 48 * 1024; i += 128) reinterpret_cast<float*>(raw)[i] = 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(KIND, M, N, LAYOUT == 1, LAYOUT == 1);
        const uint32_t aimg = base, bimg = base + 64 * 1024;
        uint64_t da[8], db[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            if (LAYOUT == 0) {
                da[kk] = TS ? (uint64_t)(tb + 256 + kk * 8) : make_desc(aimg + kk * 256, 128, 2048, 0);
                db[kk] = make_desc(bimg + kk * 256, 128, 2048, 0);
            } else if (LAYOUT == 1) {
                da[kk] = make_desc(aimg + kk * 2048, 512, 1024, 1);
                db[kk] = make_desc(bimg + kk * 2048, 512, 1024, 1);
            } else {
                da[kk] = TS ? (uint64_t)(tb + 256 + kk * 8) : make_desc(aimg + kk * 32, 16, 1024, 2);
                db[kk] = make_desc(bimg + kk * 32, 16, 1024, 2);
            }
        }
        const long long t0 = clock64();
        for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) mma<KIND, TS>(tb + (NACC > 1 ? (kk % NACC) * (N < 128 ? N : 128) % 256 : 0), da[kk], db[kk], idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0));
        if (blockIdx.x == 0) *cycles = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

template <int KIND, bool TS, int M, int N, int LAYOUT, int NACC>
static void run(const char* name, int grid) {
    long long* d;
    cudaMalloc(&d, 8);
    auto k = rate_kernel<KIND, TS, M, N, LAYOUT, NACC>;
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 256;
    k<<<grid, 128, smem>>>(reps, d);
    k<<<grid, 128, smem>>>(reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / (reps * 8);
    const int K = KIND == 0 ? 8 : 16;
    printf("%-58s grid %3d: %7.1f cycles/MMA  %7.1f MAC/clk/SM  (%s)\n", name, grid, per, (double)M * N * K / per, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148}) {
        run<0, false, 128, 64, 0, 1>("tf32 SS M128 N64  K-major noswz", grid);
        run<0, true, 128, 64, 0, 1>("tf32 TS M128 N64  K-major noswz", grid);
        run<0, true, 128, 64, 0, 2>("tf32 TS M128 N64  K-major noswz 2 accumulators", grid);
        run<0, true, 128, 128, 0, 1>("tf32 TS M128 N128 K-major noswz", grid);
        run<0, true, 128, 256, 0, 1>("tf32 TS M128 N256 K-major noswz", grid);
        run<0, false, 128, 128, 0, 1>("tf32 SS M128 N128 K-major noswz", grid);
        run<0, false, 128, 64, 2, 1>("tf32 SS M128 N64  K-major SW128", grid);
        run<0, true, 128, 64, 2, 1>("tf32 TS M128 N64  K-major SW128", grid);
        run<0, true, 128, 128, 2, 1>("tf32 TS M128 N128 K-major SW128", grid);
        run<0, false, 128, 256, 2, 1>("tf32 SS M128 N256 K-major SW128", grid);
        run<0, false, 64, 64, 1, 1>("tf32 SS M64  N64  MN-major SW128_32B", grid);
        run<0, false, 128, 64, 1, 1>("tf32 SS M128 N64  MN-major SW128_32B", grid);
        run<0, false, 128, 128, 1, 1>("tf32 SS M128 N128 MN-major SW128_32B", grid);
        run<0, false, 64, 64, 1, 2>("tf32 SS M64  N64  MN-major SW128_32B 2 accumulators", grid);
        run<1, false, 128, 64, 2, 1>("bf16 SS M128 N64  K-major SW128", grid);
        run<1, true, 128, 64, 2, 1>("bf16 TS M128 N64  K-major SW128", grid);
        run<1, false, 128, 256, 2, 1>("bf16 SS M128 N256 K-major SW128", grid);
    }
    return 0;
}
