// tc_probe2.cu — second tcgen05 probe (groundwork for the tensor-core PPO update of the next round):
//   P1  dW shape  D[k][n] = sum_m H[m][k] Z[m][n]  with MN-major tf32 operands in the SWIZZLE_128B_BASE32B
//       layout (the only MN-major layout tf32 accepts), M = 128 and M = 64 (discovers the TMEM row mapping)
//   P2  TS mode   D[m][k] = sum_n Z[m][n] W[k][n]  with the A operand written to TMEM by tcgen05.st
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe2 tools/tc_probe2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | ((uint64_t)layout_type << 61);
}
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                 "l"(da), "l"(db), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                 "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity));
}

// mode 0: P1 (operands P, Q are prebuilt shared-memory images, 32 KB each); mode 1: P2 (P = Z row-major [128][64] for
// tcgen05.st, Q = W in the no-swizzle K-major core layout)
__global__ void probe2_kernel(const float* __restrict__ P, const float* __restrict__ Q, float* __restrict__ D, int mode, int Mdim,
                              uint32_t lbo, uint32_t sbo, uint32_t ltype, uint32_t kstep_bytes) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* sP = reinterpret_cast<float*>(smem_raw);
    float* sQ = sP + 128 * 64;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (mode == 0) for (int i = tid; i < 128 * 64; i += blockDim.x) sP[i] = P[i];
    for (int i = tid; i < (mode == 0 ? 128 : 64) * 64; i += blockDim.x) sQ[i] = Q[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
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
    // clear the accumulator columns so untouched lanes read as a sentinel
    for (int c0 = 0; c0 < 64; c0 += 8) {
        const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + c0;
        const uint32_t s = __float_as_uint(-12345.0f);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(s));
    }
    if (mode == 1) {   // A operand Z[m][0..63] -> TMEM columns 64..127 of lane m
        for (int c0 = 0; c0 < 64; c0 += 8) {
            const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + 64 + c0;
            const float* z = P + (size_t)tid * 64 + c0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                         "r"(__float_as_uint(z[0])), "r"(__float_as_uint(z[1])), "r"(__float_as_uint(z[2])), "r"(__float_as_uint(z[3])),
                         "r"(__float_as_uint(z[4])), "r"(__float_as_uint(z[5])), "r"(__float_as_uint(z[6])), "r"(__float_as_uint(z[7])));
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (tid == 0) {
        const uint32_t aP = smem_u32(sP), aQ = smem_u32(sQ);
        if (mode == 0) {
            const uint32_t idesc = make_idesc(Mdim, 64, 1, 1);
            for (int kk = 0; kk < 16; ++kk)
                mma_ss(tb, make_desc(aP + kk * kstep_bytes, lbo, sbo, ltype), make_desc(aQ + kk * kstep_bytes, lbo, sbo, ltype), idesc, kk > 0);
        } else {
            const uint32_t idesc = make_idesc(128, 64, 0, 0);
            for (int kk = 0; kk < 8; ++kk)
                mma_ts(tb, tb + 64 + kk * 8, make_desc(aQ + kk * 2 * 128, 128, 2048, 0), idesc, kk > 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    wait_bar(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t r[8];
        const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int j = 0; j < 8; ++j) D[(size_t)tid * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256));
}

static inline int core_index(int r, int c, int C) { return ((r >> 3) * (C >> 2) + (c >> 2)) * 32 + (r & 7) * 4 + (c & 3); }

// MN-major SWIZZLE_128B_BASE32B image of X[m][f] (m = K index, f = MN index, F = 64 = two 32-element blocks):
//   [m/4][f/32][m%4][pos(f%32)], pos = e ^ ((e >> 2) & 3) when swz, else e
static void mn_image(const std::vector<float>& X, std::vector<float>& img, int swz, int block_outer) {
    img.assign(128 * 64, 0.f);
    for (int m = 0; m < 128; ++m)
        for (int f = 0; f < 64; ++f) {
            int e = f & 31, pos = e;
            if (swz == 1) pos = e ^ ((e >> 2) & 3);
            if (swz == 2) pos = (((e >> 2) ^ (m & 3)) << 2) | (e & 3);       // 16 B chunk index xor k-row (alternative guess)
            if (swz == 3) pos = (((e >> 3) ^ (m & 3)) << 3) | (e & 7);       // Swizzle<2,5,2> on BYTE addresses: 32 B chunk ^= k-row
            int idx;
            if (!block_outer) idx = (((m >> 2) * 2 + (f >> 5)) * 4 + (m & 3)) * 32 + pos;   // [m/4][f/32][m%4][32]
            else idx = (((f >> 5) * 32 + (m >> 2)) * 4 + (m & 3)) * 32 + pos;             // [f/32][m/4][m%4][32]
            img[idx] = X[(size_t)m * 64 + f];
        }
}

int main() {
    srand(2);
    auto rnd_int = [] { return (float)((rand() % 7) - 3); };
    std::vector<float> H(128 * 64), Z(128 * 64), W(64 * 64), dump(128 * 64), Hi, Zi, Wc(64 * 64);
    for (auto& v : H) v = rnd_int();
    for (auto& v : Z) v = rnd_int();
    for (auto& v : W) v = rnd_int();
    float *dP, *dQ, *dD;
    CK(cudaMalloc(&dP, 128 * 64 * 4)); CK(cudaMalloc(&dQ, 128 * 64 * 4)); CK(cudaMalloc(&dD, 128 * 64 * 4));
    const size_t smem = 2 * 128 * 64 * 4 + 2048;
    CK(cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> ref(64 * 64);
    for (int k = 0; k < 64; ++k) for (int n = 0; n < 64; ++n) {
        float r = 0; for (int m = 0; m < 128; ++m) r += H[m * 64 + k] * Z[m * 64 + n];
        ref[k * 64 + n] = r;
    }
    auto launch = [&](const std::vector<float>& p, const std::vector<float>& q, int mode, int Mdim, uint32_t lbo, uint32_t sbo, uint32_t lt,
                      uint32_t kstep) {
        CK(cudaMemcpy(dP, p.data(), 128 * 64 * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dQ, q.data(), q.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(dD, 0, 128 * 64 * 4));
        probe2_kernel<<<1, 128, smem>>>(dP, dQ, dD, mode, Mdim, lbo, sbo, lt, kstep);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  kernel failed: %s\n", cudaGetErrorString(e)); exit(2); }
        CK(cudaMemcpy(dump.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost));
    };
    auto score = [&](int* lane_of) {
        int found = 0;
        for (int k = 0; k < 64; ++k) {
            lane_of[k] = -1;
            for (int lane = 0; lane < 128 && lane_of[k] < 0; ++lane) {
                bool eq = true;
                for (int n = 0; n < 64 && eq; ++n) eq = dump[lane * 64 + n] == ref[k * 64 + n];
                if (eq) lane_of[k] = lane;
            }
            found += lane_of[k] >= 0;
        }
        return found;
    };
    // P1: try the layout hypotheses
    struct Var { int swz, block_outer; uint32_t lbo, sbo, kstep; const char* name; };
    Var vars[] = {
        {1, 0, 512, 1024, 2048, "[m/4][f/32][m%4][32] swz e^((e>>2)&3) LBO=512 (MN block) SBO=1024 (K group)"},
        {1, 0, 1024, 512, 2048, "same image, LBO/SBO swapped"},
        {0, 0, 512, 1024, 2048, "[m/4][f/32][m%4][32] no swizzle, LBO=512 SBO=1024"},
        {2, 0, 512, 1024, 2048, "[m/4][f/32][m%4][32] swz chunk^(m&3), LBO=512 SBO=1024"},
        {1, 1, 16384, 512, 1024, "[f/32][m/4][m%4][32] swz e^((e>>2)&3) LBO=16384 (MN block) SBO=512 (K group)"},
        {1, 1, 512, 16384, 1024, "same image, LBO/SBO swapped"},
        {2, 1, 16384, 512, 1024, "[f/32][m/4][m%4][32] swz chunk^(m&3) LBO=16384 SBO=512"},
        {0, 1, 16384, 512, 1024, "[f/32][m/4][m%4][32] no swizzle LBO=16384 SBO=512"},
        {3, 0, 512, 1024, 2048, "[m/4][f/32][m%4][32] swz 32B-chunk^(m&3), LBO=512 (MN block) SBO=1024 (K group)"},
        {3, 0, 1024, 512, 2048, "same image, LBO/SBO swapped"},
        {3, 1, 16384, 512, 1024, "[f/32][m/4][m%4][32] swz 32B-chunk^(m&3), LBO=16384 SBO=512"},
    };
    int best = -1;
    for (int v = 0; v < (int)(sizeof(vars) / sizeof(vars[0])); ++v) {
        mn_image(H, Hi, vars[v].swz, vars[v].block_outer);
        mn_image(Z, Zi, vars[v].swz, vars[v].block_outer);
        launch(Hi, Zi, 0, 128, vars[v].lbo, vars[v].sbo, 1, vars[v].kstep);
        int lane_of[64];
        int f = score(lane_of);
        int nz = 0; for (float x : dump) nz += (x != 0.f && x != -12345.0f);
        printf("P1 variant %d (%s): rows matched %d/64, nonzero outputs %d, D[0][0..3] = %g %g %g %g (ref %g %g %g %g)\n", v, vars[v].name, f,
               nz, dump[0], dump[1], dump[2], dump[3], ref[0], ref[1], ref[2], ref[3]);
        if (f == 64 && best < 0) best = v;
    }
    if (best >= 0) {
        mn_image(H, Hi, vars[best].swz, vars[best].block_outer);
        mn_image(Z, Zi, vars[best].swz, vars[best].block_outer);
        for (int Mdim : {128, 64}) {
            launch(Hi, Zi, 0, Mdim, vars[best].lbo, vars[best].sbo, 1, vars[best].kstep);
            int lane_of[64];
            int f = score(lane_of);
            printf("P1 M=%d with variant %d: rows matched %d/64; lane_of[0,1,15,16,17,31,32,33,47,48,63] = %d %d %d %d %d %d %d %d %d %d %d\n", Mdim,
                   best, f, lane_of[0], lane_of[1], lane_of[15], lane_of[16], lane_of[17], lane_of[31], lane_of[32], lane_of[33], lane_of[47],
                   lane_of[48], lane_of[63]);
        }
    } else printf("P1: no MN-major hypothesis matched\n");
    // P2: TS mode
    for (int k = 0; k < 64; ++k) for (int n = 0; n < 64; ++n) Wc[core_index(k, n, 64)] = W[k * 64 + n];
    launch(Z, Wc, 1, 128, 0, 0, 0, 0);
    {
        int bad = 0;
        for (int m = 0; m < 128; ++m) for (int k = 0; k < 64; ++k) {
            float r = 0; for (int n = 0; n < 64; ++n) r += Z[m * 64 + n] * W[k * 64 + n];
            if (dump[m * 64 + k] != r) { if (bad < 4) printf("P2 mismatch m=%d k=%d got %g ref %g\n", m, k, dump[m * 64 + k], r); ++bad; }
        }
        printf("P2 TS mode (A from TMEM via tcgen05.st, B K-major smem): %s (%d mismatches)\n", bad ? "FAIL" : "OK", bad);
    }
    return 0;
}
