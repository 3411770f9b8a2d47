// Tensor-core variant of the fused PPO loss forward/backward (ppo_loss_grad_kernel) for the reference's default
// network shape: two tanh MLPs with hidden_dims = [64, 64] (layers/layer_constructors.jl:3-11,53-57), obs_dim <= 4,
// Discrete(n <= 2) actions — i.e. the CartPole configurations of BASELINE.json.
//
// The three 64x64 GEMMs per net that hold 94 % of the FLOPs run on tcgen05.mma.kind::tf32 with the 3xTF32 split
// (hi*hi + lo*hi + hi*lo, hi = x with 13 mantissa bits cleared, lo = x - hi; 2e-6 relative error measured in
// tools/tc_probe.cu, so north_star check (c) at 1e-4 holds), accumulators in TMEM:
//   G1  H1pre[m][n] = sum_k H0[m][k] W1[k][n]   A = H0 in TMEM (TS mode), B = W1^T K-major no-swizzle smem image
//   G2  dH0[m][k]   = sum_n dZ1[m][n] W1[k][n]   A = dZ1 in TMEM,          B = W1   K-major no-swizzle smem image
//   G3  dW1[k][n]   = sum_m H0[m][k] dZ1[m][n]   A = [H0_hi | H0_lo] (M = 128) and B = dZ1_hi, dZ1_lo: MN-major
//                                                 SWIZZLE_128B_BASE32B smem images; the accumulator stays in TMEM
//                                                 across all tiles of the pass (rows 0..63 hi part, 64..127 lo part)
// Everything else is CUDA-core code with two threads per sample (thread (m, half) <-> TMEM lane m, hidden features
// [32 half, 32 half + 32)): layer 0 (K = obs_dim), the output layers, the loss head and the deltas.  The thin-layer
// gradients are sums over samples: dW2 and db1 by a register shuffle transpose-reduce (outside the critical section
// below), dW0 and db0 through a [sample][68] shared-memory staging buffer where thread (feature f, sample quarter)
// accumulates its 32 samples per tile in registers; the partial sums are combined once at the end of the pass, where
// the hi and lo rows of the dW1 accumulator are also added so that every CTA writes ONE partial plane.
//
// Two groups of 8 warps work on alternate 128-sample tiles of the CTA (ping-pong): while one group's MMAs run or
// it waits, the other group is in a CUDA-core phase.  Each group owns 192 TMEM columns (D | A/Z hi | A/Z lo; the dZ1
// operand overwrites the H0 operand once G1 has consumed it) and its own mbarriers (G1 done, G2 done, G3 done,
// region free); the dW1 accumulator (64 columns) and the 128 KB image region are shared.  A group takes the image
// region when its deltas are ready (images -> G2, G3 -> layer-0 delta, which re-reads H0 from the image while G3 still
// runs -> staging over the consumed dZ1 images -> layer-0 sums) and then releases it through an mbarrier, so use
// strictly alternates.  MMAs are issued by one elected lane of a warp whose index is warp-uniform for the compiler;
// the next tile's samples are gathered while G2/G3 run.
//
// The kernel ends with a cooperative tail (tc_fused_tail): reduction of the per-CTA partials, optional peer-memory
// push exchange between GPUs, gradient norm, clip, KL stop, statistics and Adam.
// tcgen05 rates measured on B200 by tools/tc_probe3.cu: max(44, N/2) cycles per tf32 instruction at M = 128.
// Descriptor recipes are the ones verified on hardware by tools/tc_probe*.cu (profiles/r01_tcgen05_probe.txt).
// One pass over the minibatch per net (actor, then critic).
#pragma once
#include <cooperative_groups.h>
#include "update.cuh"

#define TC_M 128
#define TC_THREADS 512
#define TC_GROUP_THREADS 256
#define TC_COL_GROUP 192      // TMEM columns per group
#define TC_COL_D 0            // G1 output, later G2 output
#define TC_COL_HI 64          // H0 hi, later dZ1 hi
#define TC_COL_LO 128
#define TC_COL_D3 384
#define TC_TMEM_COLS 512
#define TC_OFF_W1C_HI 0
#define TC_OFF_W1C_LO 16384
#define TC_OFF_WT1C_HI 32768
#define TC_OFF_WT1C_LO 49152
#define TC_OFF_H0CAT 65536    // [m/4][hi f<32 | hi f>=32 | lo f<32 | lo f>=32][m%4][32 floats], 64 KB
#define TC_OFF_Z1_HI 131072
#define TC_OFF_Z1_LO 163840
#define TC_OFF_SMALL 196608
#define TC_SMALL_FLOATS 3328
#define TC_STAGE_LD 68      // floats per sample row of a staging buffer (272 B: conflict-free 16-byte stores by 32 samples)
#define TC_SMEM_BYTES (TC_OFF_SMALL + TC_SMALL_FLOATS * 4 + 1024)

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t tc_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ uint32_t tc_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                 "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                 "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(tc_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const float* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(__float_as_uint(r[0])),
                 "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])), "r"(__float_as_uint(r[3])), "r"(__float_as_uint(r[4])),
                 "r"(__float_as_uint(r[5])), "r"(__float_as_uint(r[6])), "r"(__float_as_uint(r[7])) : "memory");
}
__device__ __forceinline__ float tc_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// [R][C] matrix in the no-swizzle K-major core layout: cores of 8 rows x 4 columns (16-byte rows), cores ordered [r/8][c/4]
__device__ __forceinline__ int tc_core_index(int r, int c, int C) { return ((r >> 3) * (C >> 2) + (c >> 2)) * 32 + (r & 7) * 4 + (c & 3); }

// warp transpose-reduce: every lane contributes v[0..N); afterwards lane l holds in v[0..N/32) the warp sums of the
// original indices l*(N/32) + i
template <int N, int OFF>
struct TcTR {
    static __device__ __forceinline__ void run(float* v, int lane) {
        const bool up = lane & OFF;
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            const float send = up ? v[i] : v[i + N / 2];
            const float keep = up ? v[i + N / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        TcTR<N / 2, OFF / 2>::run(v, lane);
    }
};
template <int N>
struct TcTR<N, 0> {
    static __device__ __forceinline__ void run(float*, int) {}
};

// 32 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* r) {
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
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(u[j]);
}

__device__ __forceinline__ void tc_ld8x2(uint32_t ta, uint32_t tb2, float* ra, float* rb) {
    uint32_t u[8], v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(ta));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(tb2));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) { ra[j] = __uint_as_float(u[j]); rb[j] = __uint_as_float(v[j]); }
}
// -DTC_TRACE: phase timestamps (clock64) of CTA 0, written by thread 0 of each group: [group][tile slot][point]
#ifdef TC_TRACE
__device__ long long g_tc_trace[2][32][16];
#define TC_MARK(pt) do { if (t == 0 && blockIdx.x == 0) g_tc_trace[g][(ACTOR ? 0 : 16) + (it & 15)][pt] = clock64(); } while (0)
#else
#define TC_MARK(pt) do { } while (0)
#endif
#ifdef TC_TRACE
#define TAIL_MARK(pt) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_tc_trace[1][31][pt] = clock64(); } while (0)
#else
#define TAIL_MARK(pt) do { } while (0)
#endif
__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_group_sync(int g) { asm volatile("bar.sync %0, 256;" ::"r"(g + 1) : "memory"); }

// per-sample minibatch scalars of one tile row (both threads of a sample load the same values)
struct TcSample {
    float x0, x1, x2, x3;
    float p, q;        // actor: advantage (raw), old log-prob; critic: return, old value
    int aidx;
    bool valid;
};
template <int NOUT, bool ACTOR>
__device__ __forceinline__ TcSample tc_gather(const LossArgs& a, long long tile, int m) {
    const BufDev& buf = a.buf;
    const int D = a.pd.obs_dim;
    TcSample s;
    s.x0 = s.x1 = s.x2 = s.x3 = 0.f; s.p = 0.f; s.q = 0.f; s.aidx = 0;
    const long long pos = a.mb.start + tile * TC_M + m;
    s.valid = pos < a.mb.start + a.mb.count;
    if (s.valid) {
        const long long sidx = a.mb.identity ? pos : feistel_permute(pos, a.mb.n_total, a.mb.fk);
        const float* xo = buf.obs + sidx * D;
        if (D == 4) {
            const float4 v = *reinterpret_cast<const float4*>(xo);
            s.x0 = v.x; s.x1 = v.y; s.x2 = v.z; s.x3 = v.w;
        } else {
            s.x0 = xo[0];
            if (D > 1) s.x1 = xo[1];
            if (D > 2) s.x2 = xo[2];
        }
        if (ACTOR) {
            s.p = buf.advantages[sidx];
            s.q = buf.logprobs[sidx];
            int ai = reinterpret_cast<const int*>(buf.actions)[sidx] - a.pd.act_start;
            s.aidx = ai < 0 ? 0 : (ai >= NOUT ? NOUT - 1 : ai);
        } else {
            s.p = buf.returns[sidx];
            s.q = buf.values[sidx];
        }
    }
    return s;
}

// one pass (one net) over all tiles of this CTA.  512 threads = 2 groups x 256; within a group thread
// (m = t & 127, half = t >> 7) owns sample m (TMEM lane m; warps w and w+4 share lane quadrant w) and the 32 hidden
// features [32*half, 32*half + 32); for the sums over samples the same thread is (f = t & 63, quarter = t >> 6).
// n_own: tiles of this pass handled by this group; base_own / base_other: mbarrier phases the groups completed in
// earlier passes.
template <int NOUT, bool ACTOR>
__device__ __forceinline__ void tc_pass(const LossArgs& a, unsigned char* sm, uint32_t sm_base, uint32_t tb, uint64_t* bars,
                                        int n_own, uint32_t base_own, uint32_t base_other, float adv_mean, float adv_den,
                                        float invB, float* stats, float* gp) {
    const PolicyDesc& pd = a.pd;
    const int net = ACTOR ? 0 : 1;
    const LayerDesc& L0 = pd.L[net][0];
    const LayerDesc& L1 = pd.L[net][1];
    const LayerDesc& L2 = pd.L[net][2];
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler (uniform datapath, no waterfalls)
    const int g = warp >> 3, t = tid & 255, lane = tid & 31;
    const int m = t & 127, half = (warp >> 2) & 1, f0 = half * 32;
    const int rf = t & 63, rq = t >> 6;            // reducer role: feature, sample quarter
    const bool issuer = (warp & 7) == 0;           // warp that issues this group's MMAs (one elected lane)
    float* sW1c_hi = reinterpret_cast<float*>(sm + TC_OFF_W1C_HI);
    float* sW1c_lo = reinterpret_cast<float*>(sm + TC_OFF_W1C_LO);
    float* sWt1c_hi = reinterpret_cast<float*>(sm + TC_OFF_WT1C_HI);
    float* sWt1c_lo = reinterpret_cast<float*>(sm + TC_OFF_WT1C_LO);
    float* sSmall = reinterpret_cast<float*>(sm + TC_OFF_SMALL);
    float* sW0 = sSmall;            // [4][64]
    float* sb0 = sW0 + 256;         // [64]
    float* sb1 = sb0 + 64;          // [64]
    float* sW2 = sb1 + 64;          // [64][4]
    float* sb2 = sW2 + 256;         // [4]
    float* sGrp = sb2 + 8 + g * 1280;
    float* sOutP = sGrp;            // per group: [2 halves][NOUT<=2][128] partial output-layer dot products
    float* sX = sGrp + 768;         // per group: [128][4] observations
    float* sStageB = reinterpret_cast<float*>(sm + TC_OFF_Z1_HI);   // [128][68] staging of dZ0 once G3 has consumed the dZ1 images
    float* sRed = reinterpret_cast<float*>(sm + TC_OFF_H0CAT);      // end-of-pass scratch [8 slots][8 sums][64]
    uint64_t* bar1 = bars + g * 4;
    uint64_t* bar2 = bars + g * 4 + 1;      // G2 complete
    uint64_t* bfree = bars + g * 4 + 2;
    uint64_t* bar3 = bars + g * 4 + 3;      // G3 complete
    uint64_t* obfree = bars + (g ^ 1) * 4 + 2;
    const uint32_t gcol = tb + (uint32_t)g * TC_COL_GROUP;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t my = gcol + lane_base;       // this thread's lane, this group's columns
    // ---- stage this net's weights; clear the dW1 accumulator ---------------------------------------------------
    __syncthreads();
    for (int i = tid; i < 64 * 64; i += TC_THREADS) {
        const int k = i >> 6, n = i & 63;
        const float w = a.pack[L1.pw_off + i];
        const float hi = tc_hi(w), lo = w - hi;
        sW1c_hi[tc_core_index(k, n, 64)] = hi; sW1c_lo[tc_core_index(k, n, 64)] = lo;
        sWt1c_hi[tc_core_index(n, k, 64)] = hi; sWt1c_lo[tc_core_index(n, k, 64)] = lo;
    }
    for (int i = tid; i < 256; i += TC_THREADS) {
        sW0[i] = a.pack[L0.pw_off + i];
        sW2[i] = a.pack[L2.pw_off + i];
    }
    if (tid < 64) { sb0[tid] = a.pack[L0.pb_off + tid]; sb1[tid] = a.pack[L1.pb_off + tid]; }
    if (tid < 4) sb2[tid] = a.pack[L2.pb_off + tid];
    if (g == 0) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 8) tc_st8(tb + lane_base + TC_COL_D3 + f0 + c0, z);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // persistent thin-layer gradient accumulators of reducer (rf, rq): sums over the samples of quarter rq of this
    // group's tiles
    float accW0_0 = 0.f, accW0_1 = 0.f, accW0_2 = 0.f, accW0_3 = 0.f, accb0 = 0.f, accb1 = 0.f, accW2_0 = 0.f, accW2_1 = 0.f;
    float accb2[NOUT];
#pragma unroll
    for (int j = 0; j < NOUT; ++j) accb2[j] = 0.f;

    const uint32_t idesc_k = tc_idesc(128, 64, 0, 0);
    const uint32_t idesc_mn = tc_idesc(128, 64, 1, 1);
    const uint32_t r4 = m & 3;
    const uint32_t img_h0 = ((m >> 2) * 4 + half) * 512 + r4 * 128;   // this thread's 128-byte row of the H0 image (hi block)
    const uint32_t img_z1 = ((m >> 2) * 2 + half) * 512 + r4 * 128;   // ... of a dZ1 image
    TcSample cur;
    cur.valid = false; cur.x0 = cur.x1 = cur.x2 = cur.x3 = cur.p = cur.q = 0.f; cur.aidx = 0;
    if (n_own > 0) cur = tc_gather<NOUT, ACTOR>(a, (long long)blockIdx.x + (long long)g * gridDim.x, m);
    for (int it = 0; it < n_own; ++it) {
        const uint32_t phase = (base_own + (uint32_t)it) & 1u;
        TC_MARK(0);
        // ---- this tile's samples were gathered during the previous tile's G2/G3 (or before the loop) ------------------
        const bool valid = cur.valid;
        const float x0 = cur.x0, x1 = cur.x1, x2 = cur.x2, x3 = cur.x3;
        float adv = 0.f, ret = 0.f, olp = 0.f, ov = 0.f;
        const int aidx = cur.aidx;
        if (ACTOR) {
            adv = cur.p;
            if (valid && a.hp.normalize_advantage) adv = (adv - adv_mean) / adv_den;
            olp = cur.q;
        } else {
            ret = cur.p;
            ov = cur.q;
        }
        // sX of the previous tile was last read before the group barrier that ended that tile
        if (half == 0) *reinterpret_cast<float4*>(sX + m * 4) = make_float4(x0, x1, x2, x3);
        // ---- layer 0 on CUDA cores, H0 hi/lo -> TMEM (A operand of G1) ----------------------------------------------
#pragma unroll 2
        for (int c0 = 0; c0 < 32; c0 += 8) {
            float h[8], hi[8], lo[8];
            {
                const float4 ba = *reinterpret_cast<const float4*>(sb0 + f0 + c0);
                const float4 bb = *reinterpret_cast<const float4*>(sb0 + f0 + c0 + 4);
                h[0] = ba.x; h[1] = ba.y; h[2] = ba.z; h[3] = ba.w; h[4] = bb.x; h[5] = bb.y; h[6] = bb.z; h[7] = bb.w;
            }
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const float xd = d == 0 ? x0 : (d == 1 ? x1 : (d == 2 ? x2 : x3));
                const float4 w0 = *reinterpret_cast<const float4*>(sW0 + d * 64 + f0 + c0);
                const float4 w1 = *reinterpret_cast<const float4*>(sW0 + d * 64 + f0 + c0 + 4);
                h[0] = fmaf(xd, w0.x, h[0]); h[1] = fmaf(xd, w0.y, h[1]); h[2] = fmaf(xd, w0.z, h[2]); h[3] = fmaf(xd, w0.w, h[3]);
                h[4] = fmaf(xd, w1.x, h[4]); h[5] = fmaf(xd, w1.y, h[5]); h[6] = fmaf(xd, w1.z, h[6]); h[7] = fmaf(xd, w1.w, h[7]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { h[j] = fast_tanh(h[j]); hi[j] = tc_hi(h[j]); lo[j] = h[j] - hi[j]; }
            tc_st8(my + TC_COL_HI + f0 + c0, hi);
            tc_st8(my + TC_COL_LO + f0 + c0, lo);
        }
        TC_MARK(1);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        tc_group_sync(g);
        // ---- G1 -------------------------------------------------------------------------------------------
        if (issuer && tc_elect_one()) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ps = 0; ps < 3; ++ps) {
                const uint32_t acol = gcol + (ps == 1 ? TC_COL_LO : TC_COL_HI);
                const uint32_t bimg = sm_base + (ps == 2 ? TC_OFF_WT1C_LO : TC_OFF_WT1C_HI);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    tc_mma_ts(gcol + TC_COL_D, acol + kk * 8, tc_desc(bimg + kk * 256, 128, 2048, 0), idesc_k, (ps | kk) ? 1u : 0u);
            }
            tc_commit(bar1);
        }
        TC_MARK(2);
        tc_wait(bar1, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        TC_MARK(3);
        // ---- H1 = tanh(D + b1) (own 32 features), output layer (partials exchanged through smem), loss head ----------
        float h1[32];
        tc_ld32(my + TC_COL_D + f0, h1);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(sb1 + f0 + j);
            h1[j] = fast_tanh(h1[j] + b.x); h1[j + 1] = fast_tanh(h1[j + 1] + b.y);
            h1[j + 2] = fast_tanh(h1[j + 2] + b.z); h1[j + 3] = fast_tanh(h1[j + 3] + b.w);
        }
        {
            float po[NOUT];
#pragma unroll
            for (int j = 0; j < NOUT; ++j) po[j] = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const float4 w = *reinterpret_cast<const float4*>(sW2 + (f0 + k) * 4);
                po[0] = fmaf(h1[k], w.x, po[0]);
                if (NOUT > 1) po[NOUT > 1 ? 1 : 0] = fmaf(h1[k], w.y, po[NOUT > 1 ? 1 : 0]);
            }
#pragma unroll
            for (int j = 0; j < NOUT; ++j) sOutP[(half * 2 + j) * 128 + m] = po[j];
        }
        tc_group_sync(g);
        float out[NOUT];
#pragma unroll
        for (int j = 0; j < NOUT; ++j) out[j] = sb2[j] + (sOutP[j * 128 + m] + sOutP[(2 + j) * 128 + m]);
        float dout[NOUT];
        if (ACTOR) {
            float mx = out[0];
#pragma unroll
            for (int j = 1; j < NOUT; ++j) mx = fmaxf(mx, out[j]);
            float ex[NOUT], s = 0.f;
#pragma unroll
            for (int j = 0; j < NOUT; ++j) { ex[j] = expf(out[j] - mx); s += ex[j]; }
            float pj[NOUT], lpj[NOUT], hsum = 0.f, p_a = 0.f;
#pragma unroll
            for (int j = 0; j < NOUT; ++j) {
                pj[j] = ex[j] / s; lpj[j] = logf(pj[j]); hsum += pj[j] * lpj[j];
                if (j == aidx) p_a = pj[j];
            }
            const float ent = -hsum;
            const float logp = logf(p_a);
            const float log_ratio = logp - olp;
            const float ratio = expf(log_ratio);
            const float rc = fminf(fmaxf(ratio, 1.0f - a.hp.clip_range), 1.0f + a.hp.clip_range);
            const float s1 = ratio * adv, s2 = rc * adv;
            const float g_logp = (!valid || s2 < s1) ? 0.f : -invB * adv * ratio;
            const float g_ent = valid ? -a.hp.ent_coef * invB : 0.f;
#pragma unroll
            for (int j = 0; j < NOUT; ++j) dout[j] = g_logp * ((j == aidx ? 1.0f : 0.0f) - pj[j]) + g_ent * (-pj[j] * (lpj[j] + ent));
            if (valid && half == 0) {
                stats[0] += -fminf(s1, s2);
                stats[2] += ent;
                stats[3] += (ratio != rc) ? 1.0f : 0.0f;
                stats[4] += expf(log_ratio) - 1.0f - log_ratio;
                stats[5] += ratio;
            }
        } else {
            const float v_raw = out[0];
            float v = v_raw;
            bool v_pass = true;
            if (a.hp.clip_range_vf >= 0.f) {
                const float dlt = v_raw - ov;
                v_pass = dlt >= -a.hp.clip_range_vf && dlt <= a.hp.clip_range_vf;
                v = ov + fminf(fmaxf(dlt, -a.hp.clip_range_vf), a.hp.clip_range_vf);
            }
            const float verr = v - ret;
            dout[0] = (valid && v_pass) ? a.hp.vf_coef * 2.0f * verr * invB : 0.f;
            if (valid && half == 0) stats[1] += verr * verr;
        }
        if (half == 0) {
#pragma unroll
            for (int j = 0; j < NOUT; ++j) accb2[j] += warp_sum(dout[j]);
        }
        // ---- dW2[k][j] += sum_m H1[m][k] dout[m][j] and db1 += sum_m dZ1[m][.] over the warp's 32 samples (shuffle
        //      transpose-reduce: lane l ends up with feature f0 + l); dZ1 = (dout W2^T) .* (1 - H1^2) in place over h1.
        //      Done on registers so that it stays outside the image-region critical section. --------------------------------
        TC_MARK(4);
        {
            float tt[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) tt[k] = h1[k] * dout[0];
            TcTR<32, 16>::run(tt, lane);
            accW2_0 += tt[0];
            if (NOUT > 1) {
#pragma unroll
                for (int k = 0; k < 32; ++k) tt[k] = h1[k] * dout[NOUT > 1 ? 1 : 0];
                TcTR<32, 16>::run(tt, lane);
                accW2_1 += tt[0];
            }
        }
#pragma unroll
        for (int n = 0; n < 32; ++n) {
            const float4 w = *reinterpret_cast<const float4*>(sW2 + (f0 + n) * 4);
            float s = dout[0] * w.x;
            if (NOUT > 1) s = fmaf(dout[NOUT > 1 ? 1 : 0], w.y, s);
            h1[n] = s * (1.0f - h1[n] * h1[n]);
        }
        {
            float tt[32];
#pragma unroll
            for (int n = 0; n < 32; ++n) tt[n] = h1[n];
            TcTR<32, 16>::run(tt, lane);
            accb1 += tt[0];
        }
        TC_MARK(5);
        // ---- take over the image region: the other group's previous tile must have released it ------------------------
        {
            const int k = g ? it : it - 1;
            if (k >= 0) tc_wait(obfree, (base_other + (uint32_t)k) & 1u);
        }
        TC_MARK(6);
        TC_MARK(7);
        // ---- H0 hi/lo back from TMEM -> H0 image; dZ1 hi/lo -> TMEM (over H0) + dZ1 images ---------------------------
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 8) {
            float ahi[8], alo[8], hi[8], lo[8];
            tc_ld8x2(my + TC_COL_HI + f0 + c0, my + TC_COL_LO + f0 + c0, ahi, alo);
            const uint32_t sw = (((c0 >> 3) ^ r4) * 32);
            *reinterpret_cast<float4*>(sm + TC_OFF_H0CAT + img_h0 + sw) = make_float4(ahi[0], ahi[1], ahi[2], ahi[3]);
            *reinterpret_cast<float4*>(sm + TC_OFF_H0CAT + img_h0 + sw + 16) = make_float4(ahi[4], ahi[5], ahi[6], ahi[7]);
            *reinterpret_cast<float4*>(sm + TC_OFF_H0CAT + img_h0 + 1024 + sw) = make_float4(alo[0], alo[1], alo[2], alo[3]);
            *reinterpret_cast<float4*>(sm + TC_OFF_H0CAT + img_h0 + 1024 + sw + 16) = make_float4(alo[4], alo[5], alo[6], alo[7]);
#pragma unroll
            for (int j = 0; j < 8; ++j) { hi[j] = tc_hi(h1[c0 + j]); lo[j] = h1[c0 + j] - hi[j]; }
            tc_st8(my + TC_COL_HI + f0 + c0, hi);
            tc_st8(my + TC_COL_LO + f0 + c0, lo);
            *reinterpret_cast<float4*>(sm + TC_OFF_Z1_HI + img_z1 + sw) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(sm + TC_OFF_Z1_HI + img_z1 + sw + 16) = make_float4(hi[4], hi[5], hi[6], hi[7]);
            *reinterpret_cast<float4*>(sm + TC_OFF_Z1_LO + img_z1 + sw) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<float4*>(sm + TC_OFF_Z1_LO + img_z1 + sw + 16) = make_float4(lo[4], lo[5], lo[6], lo[7]);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        tc_group_sync(g);
        TC_MARK(8);
        // ---- G2 (dH0) and G3 (dW1, accumulated in TMEM over the whole pass by both groups) ----------------------------
        if (issuer && tc_elect_one()) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ps = 0; ps < 3; ++ps) {
                const uint32_t acol = gcol + (ps == 1 ? TC_COL_LO : TC_COL_HI);
                const uint32_t bimg = sm_base + (ps == 2 ? TC_OFF_W1C_LO : TC_OFF_W1C_HI);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    tc_mma_ts(gcol + TC_COL_D, acol + kk * 8, tc_desc(bimg + kk * 256, 128, 2048, 0), idesc_k, (ps | kk) ? 1u : 0u);
            }
            tc_commit(bar2);                 // the layer-0 delta only needs G2; G3 keeps running underneath it
#pragma unroll
            for (int ps = 0; ps < 2; ++ps) {
                const uint32_t bimg = sm_base + (ps ? TC_OFF_Z1_LO : TC_OFF_Z1_HI);
#pragma unroll 4
                for (int kk = 0; kk < 16; ++kk)
                    tc_mma_ss(tb + TC_COL_D3, tc_desc(sm_base + TC_OFF_H0CAT + kk * 4096, 512, 2048, 1),
                              tc_desc(bimg + kk * 2048, 512, 1024, 1), idesc_mn, 1u);
            }
            tc_commit(bar3);
        }
        TC_MARK(9);
        // ---- the next tile of this group (j = 2 (it + 1) + g) is gathered while G2 / G3 run ---------------------------
        if (it + 1 < n_own) cur = tc_gather<NOUT, ACTOR>(a, (long long)blockIdx.x + (long long)(2 * (it + 1) + g) * gridDim.x, m);
        tc_wait(bar2, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        TC_MARK(10);
        // ---- dZ0 = dH0 .* (1 - H0^2), H0 = hi + lo re-read from this thread's own rows of the H0 image; staged over the
        //      dZ1 images (G3 has consumed them) -------------------------------------------------------------------------
        {
            float dz0[32];
            tc_ld32(my + TC_COL_D + f0, dz0);
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 8) {
                const uint32_t sw = (((c0 >> 3) ^ r4) * 32);
                const float4 ha = *reinterpret_cast<const float4*>(sm + TC_OFF_H0CAT + img_h0 + sw);
                const float4 hb = *reinterpret_cast<const float4*>(sm + TC_OFF_H0CAT + img_h0 + sw + 16);
                const float4 la = *reinterpret_cast<const float4*>(sm + TC_OFF_H0CAT + img_h0 + 1024 + sw);
                const float4 lb = *reinterpret_cast<const float4*>(sm + TC_OFF_H0CAT + img_h0 + 1024 + sw + 16);
                const float h0[8] = {ha.x + la.x, ha.y + la.y, ha.z + la.z, ha.w + la.w, hb.x + lb.x, hb.y + lb.y, hb.z + lb.z, hb.w + lb.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) dz0[c0 + j] *= (1.0f - h0[j] * h0[j]);
            }
            tc_wait(bar3, phase);            // G3 has consumed the dZ1 images: their space becomes the staging buffer
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float* rb = sStageB + m * TC_STAGE_LD + f0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(rb + j) = make_float4(dz0[j], dz0[j + 1], dz0[j + 2], dz0[j + 3]);
        }
        tc_group_sync(g);
        TC_MARK(11);
        // ---- dW0[d][f] += sum_m x[m][d] dZ0[m][f], db0[f] += sum_m dZ0[m][f] over this reducer's 32 samples ----------------
        {
            const float* pb = sStageB + (rq * 32) * TC_STAGE_LD + rf;
            const float* px = sX + (rq * 32) * 4;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, sb = 0.f;
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
                const float z = pb[i * TC_STAGE_LD];
                const float4 xv = *reinterpret_cast<const float4*>(px + i * 4);
                s0 = fmaf(xv.x, z, s0); s1 = fmaf(xv.y, z, s1); s2 = fmaf(xv.z, z, s2); s3 = fmaf(xv.w, z, s3);
                sb += z;
            }
            accW0_0 += s0; accW0_1 += s1; accW0_2 += s2; accW0_3 += s3; accb0 += sb;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        tc_group_sync(g);
        TC_MARK(12);
        if (t == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bfree)) : "memory");
    }
    // ---- end of pass: dW1 from TMEM (M = 128: row i <-> lane i; rows 0..63 hold the hi part of H0, rows 64..127 the lo
    //      part: the two are added through shared memory so that one partial plane per CTA is enough) ---------------------
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float dw1[32];
    float* sTmp = sRed + 8192;                                       // [64][68] lo-part rows (behind the sRed scratch)
    if (g == 0) {
        tc_ld32(tb + lane_base + TC_COL_D3 + f0, dw1);
        if (m >= 64) {
            float* d = sTmp + (m - 64) * TC_STAGE_LD + f0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(d + j) = make_float4(dw1[j], dw1[j + 1], dw1[j + 2], dw1[j + 3]);
        }
    }
    // reducer partial sums -> sRed[slot = g*4 + rq][sum][f] -> fixed-order sum over the 8 slots
    {
        float* r = sRed + ((g * 4 + rq) * 8) * 64 + rf;
        r[0 * 64] = accW0_0; r[1 * 64] = accW0_1; r[2 * 64] = accW0_2; r[3 * 64] = accW0_3;
        r[4 * 64] = accb0;
        // the shuffle-reduced sums live in (sample quadrant = warp & 3, feature = f0 + lane)
        float* r2 = sRed + ((g * 4 + (warp & 3)) * 8) * 64 + f0 + lane;
        r2[5 * 64] = accb1; r2[6 * 64] = accW2_0; r2[7 * 64] = accW2_1;
        if ((tid & 31) == 0 && half == 0) {
#pragma unroll
            for (int j = 0; j < NOUT; ++j) sRed[8 * 8 * 64 + warp * 2 + j] = accb2[j];
        }
    }
    __syncthreads();
    if (g == 0 && m < 64) {
        const float* d = sTmp + m * TC_STAGE_LD + f0;
        float* gw = gp + L1.pw_off + m * 64 + f0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 lo4 = *reinterpret_cast<const float4*>(d + j);
            *reinterpret_cast<float4*>(gw + j) = make_float4(dw1[j] + lo4.x, dw1[j + 1] + lo4.y, dw1[j + 2] + lo4.z, dw1[j + 3] + lo4.w);
        }
    }
    for (int i = tid; i < (6 + NOUT) * 64 + NOUT; i += TC_THREADS) {
        if (i < (6 + NOUT) * 64) {
            const int q = i >> 6, f = i & 63;
            float s = 0.f;
#pragma unroll
            for (int sl = 0; sl < 8; ++sl) s += sRed[(sl * 8 + q) * 64 + f];
            if (q < 4) gp[L0.pw_off + q * 64 + f] = s;                    // W0 packed [4][64]
            else if (q == 4) gp[L0.pb_off + f] = s;
            else if (q == 5) gp[L1.pb_off + f] = s;
            else gp[L2.pw_off + f * 4 + (q - 6)] = s;                     // W2 packed [64][4]
        } else {
            const int j = i - (6 + NOUT) * 64;
            const float* r = sRed + 8 * 8 * 64 + j;                      // warps 0..3 and 8..11 hold half 0
            gp[L2.pb_off + j] = ((r[0] + r[2]) + (r[4] + r[6])) + ((r[16] + r[18]) + (r[20] + r[22]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// Fused tail (cooperative launch): reduction of the per-CTA partials, (multi-GPU: one-shot exchange over NVLink peer
// memory,) global gradient norm, clip, KL stop, statistics and the Adam step, inside the loss/grad kernel.  Every CTA
// owns a contiguous slice of the flat gradient; grid barriers separate "partials written" / "slices reduced (and
// published to the peers)" / "sums of squares written".  Replaces reduce_adam_kernel (+ p2p_sum_adam_kernel).
// ---------------------------------------------------------------------------------------------------------
struct TailArgs {
    int mode;                       // 0: none (caller reduces), 1: single GPU, 2: peer-memory allreduce
    const int* flat2g;
    const unsigned char* f2planes;
    int stats_off;
    double* sq_part;                // [>= gridDim.x]
    AdamArgs adam;
    P2PDev pp;
};

// Returns the target-KL stop flag (identical in every CTA).  global_count < 0: the minibatch size over all ranks is tl.adam's.
__device__ __forceinline__ bool tc_fused_tail(const LossArgs& a, const TailArgs& tl, float* s_red /* [8][64] */, double* scratch,
                                              float* s_f, double global_count = -1.0) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const AdamArgs& ad = tl.adam;
    const double gcount = global_count > 0.0 ? global_count : ad.global_count;
    const int tid = threadIdx.x, nb = (int)gridDim.x;
    const int n = ad.n_params + 6;
    const int per = (n + nb - 1) / nb;
    const int p_lo = min(n, (int)blockIdx.x * per), p_hi = min(n, p_lo + per);
    const int gpack = a.pd.gpack;
    TAIL_MARK(0);
    // the packed index / plane count of this thread's first parameter do not depend on the other CTAs: fetched before the barrier
    int idx_first = 0, planes_first = 1;
    {
        const int p = p_lo + (tid & 63);
        if (p < p_hi) {
            idx_first = p < ad.n_params ? tl.flat2g[p] : tl.stats_off + (p - ad.n_params);
            planes_first = p < ad.n_params ? tl.f2planes[p] : 1;
        }
    }
    __threadfence();
    grid.sync();                                             // every CTA's partial planes are visible
    TAIL_MARK(1);
    const double bp1 = ad.iter_acc[12], bp2 = ad.iter_acc[13];   // running beta^t BEFORE this step (CTA 0 updates them at the end)
    const unsigned long long seq = tl.mode == 2 ? *tl.pp.local_seq + 1ull : 0ull;
    const size_t par_off = (size_t)(seq & 1ull) * tl.pp.nranks * tl.pp.n_slots;      // parity half of recv[2][nranks][n_slots]
    double sq = 0.0;
    float g_own = 0.f;                                         // reduced gradient of parameter p_lo + tid (first 64 of the slice)
    for (int base = p_lo; base < p_hi; base += 64) {
        const int lp = tid & 63, grp = tid >> 6, p = base + lp;
        float part = 0.f;
        if (p < p_hi) {
            const int idx = base == p_lo ? idx_first : (p < ad.n_params ? tl.flat2g[p] : tl.stats_off + (p - ad.n_params));
            const int planes = base == p_lo ? planes_first : (p < ad.n_params ? tl.f2planes[p] : 1);
            // all loads of this thread's share (every 8th CTA of up to 2 planes) are issued before the first add: one L2
            // round trip instead of one per batch of 8
            TAIL_MARK(5);
            const float* src0 = a.gpart + idx;
            const float* src1 = a.gpart + (size_t)a.half_stride * gpack + idx;
            for (int c0 = grp; c0 < nb; c0 += 8 * 20) {
                float v0[20], v1[20];
#pragma unroll
                for (int j = 0; j < 20; ++j) {
                    const int c = c0 + j * 8;
                    v0[j] = c < nb ? __ldcg(src0 + (size_t)c * gpack) : 0.f;
                    v1[j] = (planes > 1 && c < nb) ? __ldcg(src1 + (size_t)c * gpack) : 0.f;
                }
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int j = 0; j < 20; ++j) { s0 += v0[j]; s1 += v1[j]; }
                part += s0 + s1;
            }
            for (int pl = 2; pl < planes; ++pl) {                  // (the tensor-core kernel writes at most 2 planes)
                const float* src = a.gpart + (size_t)pl * a.half_stride * gpack + idx;
                for (int c = grp; c < nb; c += 8) part += __ldcg(src + (size_t)c * gpack);
            }
        }
        TAIL_MARK(6);
        s_red[grp * 64 + lp] = part;
        __syncthreads();
        TAIL_MARK(7);
        if (grp == 0 && p < p_hi) {
            float gs = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) gs += s_red[j * 64 + lp];
            if (tl.mode == 2) {
                // push this rank's reduced slice into every rank's receive buffer (own one included)
                const size_t off = par_off + (size_t)tl.pp.rank * tl.pp.n_slots + p;
                for (int r = 0; r < tl.pp.nranks; ++r) tl.pp.peer_recv[r][off] = gs;
            } else {
                ad.g[p] = gs;
                if (base == p_lo) g_own = gs;
                if (p < ad.n_params) sq += (double)gs * (double)gs;
            }
        }
        __syncthreads();
    }
    if (tl.mode == 2) {
        // raise this CTA's arrival flag at every rank, then wait for every rank's CTA with the same slice
        if (tid == 0) {
            __threadfence_system();
            for (int r = 0; r < tl.pp.nranks; ++r)
                *reinterpret_cast<volatile unsigned long long*>(tl.pp.peer_cflag[r] + (size_t)tl.pp.rank * tl.pp.max_cta + blockIdx.x) = seq;
        }
        if (tid < tl.pp.nranks) {
            // acquire loads at system scope: the flag and the slices it guards were written by the peers' SMs over NVLink
            const unsigned long long* f = const_cast<const unsigned long long*>(tl.pp.local_cflag) + (size_t)tid * tl.pp.max_cta + blockIdx.x;
            long long spins = 0;
            unsigned long long seen;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(f) : "memory");
                if (seen < seq && ++spins > (1ll << 26)) { atomicExch(tl.pp.err, 1); break; }
            } while (seen < seq);
        }
        __syncthreads();
        // sum of all ranks' slices in rank order (identical on every rank => bit-identical parameters)
        for (int p = p_lo + tid; p < p_hi; p += blockDim.x) {
            float gs = 0.f;
            for (int r = 0; r < tl.pp.nranks; ++r) gs += __ldcv(tl.pp.local_recv + par_off + (size_t)r * tl.pp.n_slots + p);
            ad.g[p] = gs;
            if (p == p_lo + tid && tid < 64) g_own = gs;
            if (p < ad.n_params) sq += (double)gs * (double)gs;
        }
    }
    TAIL_MARK(2);
    // Adam state of this thread's parameter is fetched before the barrier (it does not depend on the other CTAs)
    const int p_own = p_lo + tid;
    const bool own = tid < 64 && p_own < min(p_hi, ad.n_params);
    float m_own = 0.f, v_own = 0.f, w_own = 0.f;
    int ip_own = -1, it_own = -1;
    if (own) { m_own = ad.m[p_own]; v_own = ad.v[p_own]; w_own = ad.flat[p_own]; ip_own = ad.flat2pack[p_own]; it_own = ad.flat2packT[p_own]; }
    // CTA 0's statistic accumulators do not depend on the other CTAs either
    double acc_old = 0.0;
    float acc_invB = 0.f;
    if (blockIdx.x == 0 && tid < 16) { acc_old = ad.iter_acc[tid]; acc_invB = (float)(1.0 / gcount); }
    sq = block_sum(sq, scratch);
    if (tid == 0) tl.sq_part[blockIdx.x] = sq;
    __threadfence();
    grid.sync();                                             // all slices of g and all sums of squares are visible
    TAIL_MARK(3);
    double q = 0.0;
    for (int b = tid; b < nb; b += blockDim.x) q += __ldcg(tl.sq_part + b);
    q = block_sum(q, scratch);
    TAIL_MARK(8);
    const float norm = (float)sqrt(q);
    const int stop = adam_stop(ad, gcount);
    TAIL_MARK(9);
    if (blockIdx.x == 0) adam_accumulate_pre(ad, norm, stop, tid, s_f, acc_old, acc_invB);
    TAIL_MARK(10);
    if (tl.mode == 2 && blockIdx.x == 0 && tid == 32) *tl.pp.local_seq = seq;       // every CTA read the old value before the barrier
    if (stop) return true;
    float scale = 1.f;
    if (ad.hp.max_grad_norm >= 0.f && norm > ad.hp.max_grad_norm) scale = ad.hp.max_grad_norm / norm;
    const float c1 = (float)(1.0 - bp1 * (double)ad.hp.beta1), c2 = (float)(1.0 - bp2 * (double)ad.hp.beta2);
    if (own) {                                               // same arithmetic as adam_param, operands already in registers
        const float b1 = ad.hp.beta1, b2 = ad.hp.beta2;
        const float gg = __fmul_rn(g_own, scale);
        const float mm = __fadd_rn(__fmul_rn(b1, m_own), __fmul_rn(__fsub_rn(1.0f, b1), gg));
        const float vv = __fadd_rn(__fmul_rn(b2, v_own), __fmul_rn(__fmul_rn(__fsub_rn(1.0f, b2), gg), gg));
        const float upd = __fmul_rn(__fdiv_rn(__fdiv_rn(mm, c1), __fadd_rn(__fsqrt_rn(__fdiv_rn(vv, c2)), ad.hp.adam_eps)), ad.hp.lr);
        const float w = __fsub_rn(w_own, upd);
        ad.m[p_own] = mm; ad.v[p_own] = vv; ad.flat[p_own] = w;
        if (ip_own >= 0) ad.pack[ip_own] = w;
        if (it_own >= 0) ad.pack[it_own] = w;
    }
    for (int p = p_lo + 64 + tid; p < min(p_hi, ad.n_params); p += blockDim.x) adam_param(ad, p, scale, c1, c2);   // slices wider than 64
    TAIL_MARK(4);
    return false;
}

__global__ void __launch_bounds__(TC_THREADS, 1) ppo_loss_grad_tc_kernel(const __grid_constant__ LossArgs a, const __grid_constant__ TailArgs tl) {
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    __shared__ __align__(8) uint64_t bars[8];
    __shared__ uint32_t tmem_base_s;
    __shared__ double scratch[32];
    __shared__ float s_f2[2];
    if (*a.stop_flag) return;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), g = warp >> 3;
    const uint32_t raw = tc_smem_u32(tc_smem_raw);
    const uint32_t sm_base = (raw + 1023u) & ~1023u;
    unsigned char* sm = tc_smem_raw + (sm_base - raw);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_base_s)), "r"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base_s, 0);

    float adv_mean = 0.f, adv_den = 1.f;
    if (a.hp.normalize_advantage) {
        const double n = a.mb.global_count;
        const double mean = a.mbstats[0] / n;
        double var = (a.mbstats[1] - n * mean * mean) / (n - 1.0);
        if (var < 0.0) var = 0.0;
        adv_mean = (float)mean;
        adv_den = (float)sqrt(var) + 1e-8f;
    }
    const float invB = (float)(1.0 / a.mb.global_count);
    float stats[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float* gp = a.gpart + (size_t)blockIdx.x * a.pd.gpack;
    // tiles of this CTA: blockIdx.x + j * gridDim.x, j = 0 .. nt-1; group g takes j = g, g+2, ...
    const long long n_tiles = (a.mb.count + TC_M - 1) / TC_M;
    const int nt = (long long)blockIdx.x < n_tiles ? (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    const int n_own = g ? nt / 2 : (nt + 1) / 2;
    const int n_other = nt - n_own;
    if (a.pd.act_n == 1) tc_pass<1, true>(a, sm, sm_base, tb, bars, n_own, 0u, 0u, adv_mean, adv_den, invB, stats, gp);
    else tc_pass<2, true>(a, sm, sm_base, tb, bars, n_own, 0u, 0u, adv_mean, adv_den, invB, stats, gp);
    tc_pass<1, false>(a, sm, sm_base, tb, bars, n_own, (uint32_t)n_own, (uint32_t)n_other, adv_mean, adv_den, invB, stats, gp);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double s = block_sum((double)stats[i], scratch);
        if (tid == 0) gp[a.pd.pack_fwd + a.pd.act_n + i] = (float)s;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(TC_TMEM_COLS));
    if (tl.mode) tc_fused_tail(a, tl, reinterpret_cast<float*>(sm + TC_OFF_H0CAT), scratch, s_f2);
}
