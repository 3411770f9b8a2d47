// Issue-rate probe for the kind::f16 tcgen05.mma forms of update_ft.cuh (timing only): cycles per instruction for the three
// GEMM orientations in the no-swizzle core layout and in SWIZZLE_128B, M = 64 / 128, N = 64 / 128, one or two accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ft_rate tools/ft_rate.cu && tools/ft_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
struct Cfg { int M, N, a_mn, b_mn, layout, a_step, a_lbo, a_sbo, b_step, b_lbo, b_sbo, nacc, ksteps, load_warps, load_kind; };
__global__ void __launch_bounds__(512, 1) rate_kernel(Cfg c, int reps, long long* cycles) {
    __shared__ volatile int stop_flag;
    extern __shared__ __align__(1024) unsigned char raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    for (int i = tid; i < 36 * 1024; i += 512) reinterpret_cast<float*>(raw)[i] = 0.f;
    if (tid == 0) stop_flag = 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
        const uint32_t aimg = base, bimg = base + 48 * 1024;
        uint64_t da[8], db[8];
        for (int kk = 0; kk < 8; ++kk) {
            da[kk] = make_desc(aimg + (kk % c.ksteps) * c.a_step, c.a_lbo, c.a_sbo, c.layout);
            db[kk] = make_desc(bimg + (kk % c.ksteps) * c.b_step, c.b_lbo, c.b_sbo, c.layout);
        }
        const long long t0 = clock64();
        for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                // nacc 2: alternate between the two interleaved M = 64 accumulators (lane offset 0 / 16)
                const uint32_t d = tb + ((c.nacc == 2 && (kk & 1)) ? (16u << 16) : 0u);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da[kk]),
                             "l"(db[kk]), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0));
        if (blockIdx.x == 0) *cycles = clock64() - t0;
        stop_flag = 1;
    } else if (warp >= 1 && warp <= c.load_warps) {
        // shared-memory traffic next to the operand images: kind 1 = LDS.128, 2 = STS.128, 3 = both (like the kernel's phases)
        float4* area = reinterpret_cast<float4*>(raw + (base - smem_u32(raw)) + 112 * 1024) + (warp - 1) * 64;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int it = 0;
        while (!stop_flag) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (c.load_kind & 1) { const float4 v = area[(tid & 31) + ((u & 1) << 5)]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
                if (c.load_kind & 2) area[(tid & 31) + ((u & 1) << 5)] = acc;
            }
            ++it;
        }
        if (acc.x == 123.f && it == -1) cycles[1] = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}
static void run(const char* name, Cfg c) {
    long long* d;
    cudaMalloc(&d, 16);
    const int smem = 200 * 1024, reps = 256;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_kernel<<<148, 512, smem>>>(c, reps, d);
    rate_kernel<<<148, 512, smem>>>(c, reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long cy = 0;
    cudaMemcpy(&cy, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)cy / (reps * 8);
    printf("%-64s %7.1f cycles/MMA  %7.1f MAC/clk/SM  (%s)\n", name, per, (double)c.M * c.N * 16 / per, cudaGetErrorString(e));
    cudaFree(d);
}
int main() {
    // no-swizzle core layout, 64 x 64 images: K-major: step 256, LBO 128, SBO 1024; MN-major: step 2048, LBO 1024, SBO 128
    //                         M, N, a_mn, b_mn, layout, a_step, a_lbo, a_sbo, b_step, b_lbo, b_sbo, nacc, ksteps
    run("noswz M64 N64 G1 (A K-major, B MN-major)",            {64, 64, 0, 1, 0, 256, 128, 1024, 2048, 1024, 128, 1, 4, 0, 0});
    run("noswz M64 N64 G2 (A MN-major, B MN-major)",           {64, 64, 1, 1, 0, 2048, 1024, 128, 2048, 1024, 128, 1, 4, 0, 0});
    run("noswz M64 N64 G3 (A K-major, B K-major)",             {64, 64, 0, 0, 0, 256, 128, 1024, 256, 128, 1024, 1, 4, 0, 0});
    run("noswz M64 N64 G1, two interleaved accumulators",      {64, 64, 0, 1, 0, 256, 128, 1024, 2048, 1024, 128, 2, 4, 0, 0});
    run("noswz M64 N64 G3, two interleaved accumulators",      {64, 64, 0, 0, 0, 256, 128, 1024, 256, 128, 1024, 2, 4, 0, 0});
    // 128-sample tiles (N = 128 for G1/G2: B images [64][128]: MN-major LBO 2048 (K dir), SBO 128; G3: K = 128 samples)
    run("noswz M64 N128 G1 (B MN-major [64][128])",            {64, 128, 0, 1, 0, 256, 128, 1024, 4096, 2048, 128, 1, 4, 0, 0});
    run("noswz M64 N128 G2 (A MN, B MN)",                      {64, 128, 1, 1, 0, 2048, 1024, 128, 4096, 2048, 128, 1, 4, 0, 0});
    run("noswz M128 N64 G1 form",                              {128, 64, 0, 1, 0, 256, 128, 1024, 2048, 1024, 128, 1, 4, 0, 0});
    run("noswz M128 N128 G1 form",                             {128, 128, 0, 1, 0, 256, 128, 1024, 4096, 2048, 128, 1, 4, 0, 0});
    // SWIZZLE_128B (layout code 2), rows of 128 B = 64 halfs: K-major: step 32, SBO 1024; MN-major: step 2048 (16 K rows), SBO 1024
    run("SW128 M64 N64 G1 (A K-major, B MN-major)",            {64, 64, 0, 1, 2, 32, 16, 1024, 2048, 8192, 1024, 1, 4, 0, 0});
    run("SW128 M64 N64 G2 (A MN-major, B MN-major)",           {64, 64, 1, 1, 2, 2048, 8192, 1024, 2048, 8192, 1024, 1, 4, 0, 0});
    run("SW128 M64 N64 G3 (A K-major, B K-major)",             {64, 64, 0, 0, 2, 32, 16, 1024, 32, 16, 1024, 1, 4, 0, 0});
    run("SW128 M64 N64 G1, two interleaved accumulators",      {64, 64, 0, 1, 2, 32, 16, 1024, 2048, 8192, 1024, 2, 4, 0, 0});
    run("SW128 M64 N128 G1 (B: two 64-sample blocks, LBO 8192)", {64, 128, 0, 1, 2, 32, 16, 1024, 2048, 8192, 1024, 1, 4, 0, 0});
    run("SW128 M128 N64 G3 form",                              {128, 64, 0, 0, 2, 32, 16, 1024, 32, 16, 1024, 1, 4, 0, 0});
    run("SW128 M128 N128 G3 form",                             {128, 128, 0, 0, 2, 32, 16, 1024, 32, 16, 1024, 1, 4, 0, 0});
    run("SW128 M128 N256 G3 form",                             {128, 256, 0, 0, 2, 32, 16, 1024, 32, 16, 1024, 1, 4, 0, 0});
    // the same instruction forms with shared-memory traffic from the other warps of the CTA
    for (int kind = 1; kind <= 3; ++kind)
        for (int lw : {3, 7, 15}) {
            char name[128];
            snprintf(name, sizeof(name), "noswz M64 N64 G1 + %d warps of %s", lw, kind == 1 ? "LDS.128" : (kind == 2 ? "STS.128" : "LDS.128 + STS.128"));
            run(name, {64, 64, 0, 1, 0, 256, 128, 1024, 2048, 1024, 128, 1, 4, lw, kind});
        }
    run("noswz M64 N64 G3 + 15 warps of LDS.128 + STS.128", {64, 64, 0, 0, 0, 256, 128, 1024, 256, 128, 1024, 1, 4, 15, 3});
    run("SW128 M64 N64 G1 + 15 warps of LDS.128 + STS.128", {64, 64, 0, 1, 2, 32, 16, 1024, 2048, 8192, 1024, 1, 4, 15, 3});
    run("SW128 M64 N64 G3 + 15 warps of LDS.128 + STS.128", {64, 64, 0, 0, 2, 32, 16, 1024, 32, 16, 1024, 1, 4, 15, 3});
    return 0;
}
