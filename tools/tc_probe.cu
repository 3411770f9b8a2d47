// tc_probe.cu — stand-alone probe of tcgen05.mma.kind::tf32 with hand-built shared-memory
// descriptors (no-swizzle canonical layouts), used to pin down the layouts for the PPO update path:
//   T1 forward   D[m][n] = sum_k A[m][k] W[k][n]      A: K-major,  B: MN-major view of W
//   T2 dH        D[m][k] = sum_n Z[m][n] W[k][n]      A: K-major,  B: K-major view of W
//   T3 dW        D[k][n] = sum_m H[m][k] Z[m][n]      A: MN-major view of H, B: MN-major view of Z (M = 64 / 128)
//   T4 3xTF32    forward with hi/lo splits, error vs fp64
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// no-swizzle descriptor: start, LBO, SBO in bytes
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46);
}
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc));
}

struct Op {      // one operand: canonical core-matrix storage [R/8][C/4][8][4] floats for a logical [R][C] matrix
    int R, C;
};
// element (r, c) of a [R][C] matrix stored as core matrices of 8 rows x 4 cols (16 B rows), cores ordered
// [r/8][c/4]
__host__ __device__ inline int core_index(int r, int c, int C) { return ((r >> 3) * (C >> 2) + (c >> 2)) * 32 + (r & 7) * 4 + (c & 3); }

// mode: 0 = T1 fwd, 1 = T2 dH, 2 = T3 dW.  P [RP][CP] and Q [RQ][CQ] are stored in core layout.
//   fwd: P = A[128][64] (m,k), Q = W[64][64] (k,n)
//   dH : P = Z[128][64] (m,n), Q = W[64][64] (k,n)
//   dW : P = H[128][64] (m,k), Q = Z[128][64] (m,n), M = Mdim (64 or 128)
__global__ void probe_kernel(const float* __restrict__ P, const float* __restrict__ Q, const float* __restrict__ P2,
                             const float* __restrict__ Q2, float* __restrict__ D /* [128][ncols] raw TMEM dump */, int mode,
                             int Mdim, int ncols, int split, int reps, long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sP = reinterpret_cast<float*>(smem_raw);            // 128*64
    float* sQ = sP + 128 * 64;                                 // up to 128*64
    float* sP2 = sQ + 128 * 64;                                // lo parts
    float* sQ2 = sP2 + 128 * 64;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int RQ = mode == 2 ? 128 : 64;
    for (int i = tid; i < 128 * 64; i += blockDim.x) { sP[i] = P[i]; if (split) sP2[i] = P2[i]; }
    for (int i = tid; i < RQ * 64; i += blockDim.x) { sQ[i] = Q[i]; if (split) sQ2[i] = Q2[i]; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128));
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
        const uint32_t aP = smem_u32(sP), aQ = smem_u32(sQ), aP2 = smem_u32(sP2), aQ2 = smem_u32(sQ2);
        uint32_t idesc;
        int nk;
        if (mode == 0) { idesc = make_idesc(128, 64, 0, 1); nk = 8; }
        else if (mode == 1) { idesc = make_idesc(128, 64, 0, 0); nk = 8; }
        else { idesc = make_idesc(Mdim, 64, 1, 1); nk = 16; }
        const int passes = split ? 3 : 1;
        uint32_t acc = 0;
        const long long t0 = clock64();
        for (int rep = 0; rep < reps; ++rep)
        for (int ps = 0; ps < passes; ++ps) {
            // pass 0: hi*hi, pass 1: lo*hi, pass 2: hi*lo
            const uint32_t bp = (ps == 1) ? aP2 : aP;
            const uint32_t bq = (ps == 2) ? aQ2 : aQ;
            for (int kk = 0; kk < nk; ++kk) {
                uint64_t da, db;
                if (mode == 0) {
                    // A (m,k) K-major: rows m: 16 B apart within a core, m/8 groups 2048 B apart (SBO), 16 B k-chunks 128 B apart (LBO)
                    da = make_desc(bp + kk * 2 * 128, 128, 2048);
                    // B = W (k,n) viewed (n,k) MN-major: n/4 blocks 128 B apart (SBO), k/8 groups 2048 B apart (LBO)
                    db = make_desc(bq + kk * 2048, 2048, 128);
                } else if (mode == 1) {
                    // A = Z (m,n) K-major with K = n
                    da = make_desc(bp + kk * 2 * 128, 128, 2048);
                    // B = W (k,n) viewed (N'=k, K'=n) K-major: k rows 16 B apart, k/8 groups 2048 B (SBO), n/4 chunks 128 B (LBO)
                    db = make_desc(bq + kk * 2 * 128, 128, 2048);
                } else {
                    // A' = H (m,k) viewed (M'=k, K'=m) MN-major: k/4 blocks 128 B apart (SBO), m/8 groups 2048 B apart (LBO)
                    da = make_desc(bp + kk * 2048, 2048, 128);
                    // B' = Z (m,n) viewed (N'=n, K'=m) MN-major
                    db = make_desc(bq + kk * 2048, 2048, 128);
                }
                mma_tf32(tb, da, db, idesc, acc);
                acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(smem_u32(&bar)), "r"(0));
        }
        if (cycles) *cycles = clock64() - t0;
    }
    // wait for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(smem_u32(&bar)), "r"(0));
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // dump: warp w reads lanes 32w..32w+31
    for (int c0 = 0; c0 < ncols; c0 += 8) {
        uint32_t r[8];
        const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int j = 0; j < 8; ++j) D[(size_t)tid * ncols + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(128));
}

static void to_core(const std::vector<float>& src, int R, int C, std::vector<float>& dst) {
    dst.assign((size_t)R * C, 0.f);
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) dst[core_index(r, c, C)] = src[(size_t)r * C + c];
}
static int g_reps = 1;
static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; }

int main() {
    const int M = 128, K = 64, N = 64;
    srand(1);
    auto rnd_int = [] { return (float)((rand() % 7) - 3); };
    auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    std::vector<float> A(M * K), W(K * N), Z(M * N), Ac, Wc, Zc, dump(128 * 64);
    for (auto& v : A) v = rnd_int();
    for (auto& v : W) v = rnd_int();
    for (auto& v : Z) v = rnd_int();
    to_core(A, M, K, Ac); to_core(W, K, N, Wc); to_core(Z, M, N, Zc);
    float *dP, *dQ, *dP2, *dQ2, *dD;
    long long* dCyc;
    CK(cudaMalloc(&dCyc, 8));
    CK(cudaMalloc(&dP, 128 * 64 * 4)); CK(cudaMalloc(&dQ, 128 * 64 * 4)); CK(cudaMalloc(&dP2, 128 * 64 * 4));
    CK(cudaMalloc(&dQ2, 128 * 64 * 4)); CK(cudaMalloc(&dD, 128 * 64 * 4));
    const size_t smem = 4 * 128 * 64 * 4 + 256;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto run = [&](const std::vector<float>& p, const std::vector<float>& q, const std::vector<float>* p2, const std::vector<float>* q2,
                   int mode, int Mdim) {
        CK(cudaMemcpy(dP, p.data(), p.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dQ, q.data(), q.size() * 4, cudaMemcpyHostToDevice));
        if (p2) CK(cudaMemcpy(dP2, p2->data(), p2->size() * 4, cudaMemcpyHostToDevice));
        if (q2) CK(cudaMemcpy(dQ2, q2->data(), q2->size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(dD, 0, 128 * 64 * 4));
        probe_kernel<<<1, 128, smem>>>(dP, dQ, dP2, dQ2, dD, mode, Mdim, 64, p2 ? 1 : 0, g_reps, dCyc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(dump.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost));
    };
    // T1
    run(Ac, Wc, nullptr, nullptr, 0, 128);
    {
        int bad = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            float ref = 0; for (int k = 0; k < K; ++k) ref += A[m * K + k] * W[k * N + n];
            if (dump[m * 64 + n] != ref) { if (bad < 5) printf("T1 mismatch m=%d n=%d got %g ref %g\n", m, n, dump[m * 64 + n], ref); ++bad; }
        }
        printf("T1 forward (A K-major, B MN-major): %s (%d mismatches)\n", bad ? "FAIL" : "OK", bad);
    }
    // T2
    run(Zc, Wc, nullptr, nullptr, 1, 128);
    {
        int bad = 0;
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
            float ref = 0; for (int n = 0; n < N; ++n) ref += Z[m * N + n] * W[k * N + n];
            if (dump[m * 64 + k] != ref) { if (bad < 5) printf("T2 mismatch m=%d k=%d got %g ref %g\n", m, k, dump[m * 64 + k], ref); ++bad; }
        }
        printf("T2 dH (A K-major, B K-major): %s (%d mismatches)\n", bad ? "FAIL" : "OK", bad);
    }
    // T3 with M=128 (rows 64..127 are don't-care) and M=64 (discover the TMEM row mapping)
    for (int Mdim : {128, 64}) {
        run(Ac, Zc, nullptr, nullptr, 2, Mdim);
        std::vector<float> ref(64 * 64);
        for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) {
            float r = 0; for (int m = 0; m < M; ++m) r += A[m * K + k] * Z[m * N + n];
            ref[k * 64 + n] = r;
        }
        // find for each logical row k the TMEM lane holding it
        int found = 0, ident = 0;
        int lane_of[64];
        for (int k = 0; k < 64; ++k) {
            lane_of[k] = -1;
            for (int lane = 0; lane < 128; ++lane) {
                bool eq = true;
                for (int n = 0; n < 64 && eq; ++n) eq = dump[lane * 64 + n] == ref[k * 64 + n];
                if (eq) { lane_of[k] = lane; break; }
            }
            if (lane_of[k] >= 0) ++found;
            if (lane_of[k] == k) ++ident;
        }
        printf("T3 dW (A,B MN-major) M=%d: rows found %d/64, identity-mapped %d; lane_of[0,1,15,16,17,31,32,33,63] = %d %d %d %d %d %d %d %d %d\n",
               Mdim, found, ident, lane_of[0], lane_of[1], lane_of[15], lane_of[16], lane_of[17], lane_of[31], lane_of[32], lane_of[33], lane_of[63]);
    }
    // T4 3xTF32 accuracy on random fp32 data, through the verified K-major path (dH shape)
    {
        std::vector<float> Zr(M * N), Wr(K * N), Zh(M * N), Zl(M * N), Wh(K * N), Wl(K * N), Zhc, Zlc, Whc, Wlc, Zrc, Wrc;
        for (auto& v : Zr) v = rnd();
        for (auto& v : Wr) v = rnd();
        for (size_t i = 0; i < Zr.size(); ++i) { Zh[i] = tf32_trunc(Zr[i]); Zl[i] = Zr[i] - Zh[i]; }
        for (size_t i = 0; i < Wr.size(); ++i) { Wh[i] = tf32_trunc(Wr[i]); Wl[i] = Wr[i] - Wh[i]; }
        to_core(Zh, M, N, Zhc); to_core(Zl, M, N, Zlc); to_core(Wh, K, N, Whc); to_core(Wl, K, N, Wlc);
        to_core(Zr, M, N, Zrc); to_core(Wr, K, N, Wrc);
        std::vector<double> ref(M * K);
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
            double r = 0; for (int n = 0; n < N; ++n) r += (double)Zr[m * N + n] * Wr[k * N + n];
            ref[m * K + k] = r;
        }
        auto err = [&](const char* name) {
            double e = 0, nrm = 0;
            for (int i = 0; i < M * K; ++i) { e = fmax(e, fabs(dump[i] - ref[i])); nrm = fmax(nrm, fabs(ref[i])); }
            printf("T4 %s: max abs err %.3e (max |ref| %.3f)\n", name, e, nrm);
        };
        run(Zrc, Wrc, nullptr, nullptr, 1, 128); err("1xTF32 (raw fp32 operands, hardware truncation)");
        run(Zhc, Whc, &Zlc, &Wlc, 1, 128); err("3xTF32 (hi*hi + lo*hi + hi*lo)");
        double e = 0;
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
            float r = 0; for (int n = 0; n < N; ++n) r = fmaf(Zr[m * N + n], Wr[k * N + n], r);
            e = fmax(e, fabs(r - ref[m * K + k]));
        }
        printf("T4 fp32 FMA chain: max abs err %.3e\n", e);
        // T5 throughput of the issue stream: 200 x (3 passes x 8 MMAs of 128x64x8)
        g_reps = 200;
        run(Zhc, Whc, &Zlc, &Wlc, 1, 128);
        long long cyc = 0;
        CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
        double macs = 200.0 * 3 * 8 * 128 * 64 * 8;
        printf("T5 one SM, 4800 tcgen05.mma (M128 N64 K8 tf32) back to back: %lld cycles -> %.1f MAC/clk/SM (fp32-equivalent 3xTF32: %.1f MAC/clk/SM; FFMA peak 128)\n",
               cyc, macs / cyc, macs / 3 / cyc);
        g_reps = 1;
    }
    return 0;
}
