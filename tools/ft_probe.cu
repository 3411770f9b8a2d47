// Probe for the building blocks of the features-on-lanes loss/grad kernel (update_ft.cuh), all checked bit-exact on
// integer-valued fp16 data (one CTA, 128 threads):
//   * kind::f16 operands in the no-swizzle core layout (8 x 8 halfs = 128 B per core) used BOTH ways: one image
//     [row][col] (col contiguous) as a K-major operand (row = M/N, col = K) and as an MN-major operand (row = K,
//     col = M/N); variant 0: LBO = stride along K, SBO = stride along MN (CUTLASS make_umma_desc), variant 1: swapped;
//   * two M = 64 accumulators interleaved in the same TMEM columns (lane offset 0 and 16);
//   * 1-D bulk async copy global -> shared with mbarrier complete_tx.
//   T1  D[n][m] = sum_k W[n][k] H[k][m]   A = W K-major,  B = H MN-major      (forward GEMM of the kernel)
//   T2  D[k][m] = sum_n W[n][k] Z[n][m]   A = W MN-major, B = Z MN-major      (dH GEMM)
//   T3  D[k][n] = sum_m H[k][m] Z[n][m]   A = H K-major,  B = Z K-major       (dW GEMM)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ft_probe tools/ft_probe.cu && tools/ft_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, float* r) {
    uint32_t u[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]),
          "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]),
          "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]),
          "=r"(u[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(u[j]);
}
// element (r, c) of a [R][64] fp16 matrix in the no-swizzle core layout, cores ordered [r/8][c/8]
__host__ __device__ inline int core_idx(int r, int c) { return ((r >> 3) * 8 + (c >> 3)) * 64 + (r & 7) * 8 + (c & 7); }

#define IMG 8192      // bytes of one 64 x 64 fp16 image
#define OFF_W 0
#define OFF_W2 (1 * IMG)
#define OFF_H 2 * IMG
#define OFF_Z 3 * IMG
#define OFF_BULK 4 * IMG
#define BULK_BYTES 2304

// out: [6 slots][128 lanes][64 cols]; slots: T1 v0, T1 v1, T2 v0, T2 v1, T3, bulk check (first floats)
__global__ void __launch_bounds__(128, 1) probe_kernel(const __half* W, const __half* W2, const __half* H, const __half* Z, const float* bulk_src,
                                                       float* out) {
    extern __shared__ __align__(1024) unsigned char raw[];
    __shared__ __align__(8) uint64_t bar, bar_bulk;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    unsigned char* sm = raw + (base - smem_u32(raw));
    for (int i = tid; i < 64 * 64; i += 128) {
        const int r = i >> 6, c = i & 63;
        reinterpret_cast<__half*>(sm + OFF_W)[core_idx(r, c)] = W[i];
        reinterpret_cast<__half*>(sm + OFF_W2)[core_idx(r, c)] = W2[i];
        reinterpret_cast<__half*>(sm + OFF_H)[core_idx(r, c)] = H[i];
        reinterpret_cast<__half*>(sm + OFF_Z)[core_idx(r, c)] = Z[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_bulk)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        // bulk copy
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_bulk)), "r"(BULK_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + OFF_BULK),
                     "l"(bulk_src), "r"(BULK_BYTES), "r"(smem_u32(&bar_bulk)) : "memory");
        const uint32_t lane16 = 16u << 16;
        for (int v = 0; v < 2; ++v) {
            const uint32_t mlbo = v ? 128 : 1024, msbo = v ? 1024 : 128;      // MN-major operand strides
            // T1: A = W (second accumulator: W2) K-major, B = H MN-major
            const uint32_t i1 = make_idesc(64, 64, 0, 1);
            for (int kk = 0; kk < 4; ++kk) {
                mma_ss(tb + (uint32_t)v * 64, make_desc(base + OFF_W + kk * 256, 128, 1024), make_desc(base + OFF_H + kk * 2048, mlbo, msbo), i1, kk ? 1u : 0u);
                mma_ss(tb + lane16 + (uint32_t)v * 64, make_desc(base + OFF_W2 + kk * 256, 128, 1024), make_desc(base + OFF_H + kk * 2048, mlbo, msbo), i1, kk ? 1u : 0u);
            }
            // T2: A = W MN-major (M = k, K = n), B = Z MN-major
            const uint32_t i2 = make_idesc(64, 64, 1, 1);
            for (int kk = 0; kk < 4; ++kk) {
                mma_ss(tb + 128 + (uint32_t)v * 64, make_desc(base + OFF_W + kk * 2048, mlbo, msbo), make_desc(base + OFF_Z + kk * 2048, mlbo, msbo), i2, kk ? 1u : 0u);
                mma_ss(tb + lane16 + 128 + (uint32_t)v * 64, make_desc(base + OFF_W2 + kk * 2048, mlbo, msbo), make_desc(base + OFF_Z + kk * 2048, mlbo, msbo), i2, kk ? 1u : 0u);
            }
        }
        // T3: A = H K-major (M = k, K = m), B = Z K-major (N = n, K = m); second accumulator: A = Z, B = H
        const uint32_t i3 = make_idesc(64, 64, 0, 0);
        for (int kk = 0; kk < 4; ++kk) {
            mma_ss(tb + 256, make_desc(base + OFF_H + kk * 256, 128, 1024), make_desc(base + OFF_Z + kk * 256, 128, 1024), i3, kk ? 1u : 0u);
            mma_ss(tb + lane16 + 256, make_desc(base + OFF_Z + kk * 256, 128, 1024), make_desc(base + OFF_H + kk * 256, 128, 1024), i3, kk ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar_bulk)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t my = tb + ((uint32_t)(warp * 32) << 16);
    for (int s = 0; s < 5; ++s) {
        float r[32];
        for (int h = 0; h < 2; ++h) {
            ld32(my + s * 64 + h * 32, r);
            for (int j = 0; j < 32; ++j) out[((size_t)s * 128 + tid) * 64 + h * 32 + j] = r[j];
        }
    }
    for (int i = tid; i < BULK_BYTES / 4; i += 128) out[(size_t)5 * 128 * 64 + i] = reinterpret_cast<const float*>(sm + OFF_BULK)[i];
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
    (void)lane;
}

int main() {
    std::vector<__half> W(4096), W2(4096), H(4096), Z(4096);
    std::vector<float> fW(4096), fW2(4096), fH(4096), fZ(4096), bulk(BULK_BYTES / 4);
    srand(1);
    auto fill = [](std::vector<__half>& h, std::vector<float>& f) {
        for (int i = 0; i < 4096; ++i) { f[i] = (float)(rand() % 7 - 3); h[i] = __float2half(f[i]); }
    };
    fill(W, fW); fill(W2, fW2); fill(H, fH); fill(Z, fZ);
    for (size_t i = 0; i < bulk.size(); ++i) bulk[i] = (float)i * 0.5f;
    __half *dW, *dW2, *dH, *dZ;
    float *dB, *dO;
    const size_t out_n = (size_t)6 * 128 * 64;
    cudaMalloc(&dW, 8192); cudaMalloc(&dW2, 8192); cudaMalloc(&dH, 8192); cudaMalloc(&dZ, 8192);
    cudaMalloc(&dB, BULK_BYTES); cudaMalloc(&dO, out_n * 4);
    cudaMemcpy(dW, W.data(), 8192, cudaMemcpyHostToDevice); cudaMemcpy(dW2, W2.data(), 8192, cudaMemcpyHostToDevice);
    cudaMemcpy(dH, H.data(), 8192, cudaMemcpyHostToDevice); cudaMemcpy(dZ, Z.data(), 8192, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, bulk.data(), BULK_BYTES, cudaMemcpyHostToDevice);
    cudaMemset(dO, 0, out_n * 4);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    probe_kernel<<<1, 128, 64 * 1024>>>(dW, dW2, dH, dZ, dB, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> out(out_n);
    cudaMemcpy(out.data(), dO, out_n * 4, cudaMemcpyDeviceToHost);
    // references; matrices are [row][col] row-major: W[n][k], H[k][m], Z[n][m]
    auto ref = [&](int test, int acc, int r, int c) {
        float s = 0.f;
        const std::vector<float>& w = acc ? fW2 : fW;
        for (int t = 0; t < 64; ++t) {
            if (test == 1) s += w[r * 64 + t] * fH[t * 64 + c];               // D[n][m] = sum_k W[n][k] H[k][m]
            else if (test == 2) s += w[t * 64 + r] * fZ[t * 64 + c];          // D[k][m] = sum_n W[n][k] Z[n][m]
            else s += acc ? fZ[r * 64 + t] * fH[c * 64 + t] : fH[r * 64 + t] * fZ[c * 64 + t];   // D[k][n] = sum_m H[k][m] Z[n][m]
        }
        return s;
    };
    const char* names[5] = {"T1 fwd (A K-major, B MN-major) v0 LBO=K-stride", "T1 v1 (LBO/SBO swapped)", "T2 dH (A, B MN-major) v0", "T2 v1 (swapped)",
                            "T3 dW (A, B K-major)"};
    const int tests[5] = {1, 1, 2, 2, 3};
    for (int s = 0; s < 5; ++s)
        for (int acc = 0; acc < 2; ++acc) {
            int bad = 0;
            for (int i = 0; i < 64; ++i) {
                const int lane = 32 * (i / 16) + 16 * acc + i % 16;        // M = 64: row i -> lane 32 (i / 16) + i % 16 (+16: second accumulator)
                for (int c = 0; c < 64; ++c)
                    if (out[((size_t)s * 128 + lane) * 64 + c] != ref(tests[s], acc, i, c)) ++bad;
            }
            printf("%-50s accumulator at lane offset %2d: %s (%d mismatches)\n", names[s], 16 * acc, bad ? "FAIL" : "OK", bad);
        }
    int bad = 0;
    for (size_t i = 0; i < bulk.size(); ++i) if (out[(size_t)5 * 128 * 64 + i] != bulk[i]) ++bad;
    printf("bulk async copy (cp.async.bulk + mbarrier complete_tx, %d bytes): %s\n", BULK_BYTES, bad ? "FAIL" : "OK");
    return 0;
}
