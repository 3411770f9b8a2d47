// PPO update: minibatch advantage moments, fused loss forward/backward, gradient reduction,
// grad-norm clip + Adam.
//
// Replaces, per minibatch, Lux.Training.compute_gradients(AutoZygote(), alg, batch, train_state)
// (algorithms/ppo.jl:207) i.e. the loss functor ppo.jl:365-407 and its reverse pass, the
// post-processing of ppo.jl:209-238 (utils/optimization_utils.jl:74-107) and
// Lux.Training.apply_gradients! with Optimisers.Adam (ppo.jl:239, 64-66).  Minibatches are the
// DataLoader batches of ppo.jl:188-195 with the shuffle replaced by a keyed Feistel bijection
// evaluated on the fly (no permutation array, no gather pass).
#pragma once
#include "mlp.cuh"
#include "mma_tiles.cuh"

struct UpdateHyper {
    float clip_range, clip_range_vf, ent_coef, vf_coef, max_grad_norm, target_kl;
    int normalize_advantage;
    float lr, beta1, beta2, adam_eps;
};

struct Minibatch {
    long long n_total;   // samples in the (local) buffer
    long long start;     // first permuted position of this minibatch
    long long count;     // local samples in this minibatch
    double global_count; // samples of this minibatch over all ranks (1/B of the loss means)
    FeistelKey fk;
    int identity;        // parity entry: no permutation
};

// ---- minibatch advantage moments (normalize!, ppo.jl:350-356) ---------------------------
// grid (blocks_per_mb, n_minibatches); partial[(mb*gridDim.x + blk)*2 + {0,1}] = sum, sum of squares
#define DRIL_MAX_EPOCHS_BATCHED 16
struct FeistelKeys { FeistelKey k[DRIL_MAX_EPOCHS_BATCHED]; };
// grid (blocks_per_mb, n_minibatches, epochs): the moments of every minibatch of every epoch of an update in
// one launch (advantages and permutations are fixed during the update) => one allreduce per update
__global__ void __launch_bounds__(256) adv_stats_kernel(const float* __restrict__ adv, long long n_total,
                                                        long long batch_size, const __grid_constant__ FeistelKeys fks,
                                                        int identity, double* __restrict__ partial) {
    __shared__ double scratch[32];
    const FeistelKey& fk = fks.k[blockIdx.z];
    partial += (size_t)blockIdx.z * gridDim.y * gridDim.x * 2;
    long long start = (long long)blockIdx.y * batch_size;
    long long end = min(start + batch_size, n_total);
    double s = 0, q = 0;
    for (long long i = start + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (long long)gridDim.x * blockDim.x) {
        long long idx = identity ? i : feistel_permute(i, n_total, fk);
        double a = adv[idx];
        s += a; q += a * a;
    }
    s = block_sum(s, scratch); q = block_sum(q, scratch);
    if (threadIdx.x == 0) {
        partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2] = s;
        partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 + 1] = q;
    }
}
// mbstats[mb*2 + {0,1}] = fixed-order sum of the partials (then allreduced across ranks)
__global__ void adv_stats_finalize_kernel(const double* __restrict__ partial, int blocks_per_mb, int n_mb,
                                          double* __restrict__ mbstats) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mb * 2) return;
    int mb = i >> 1, c = i & 1;
    double s = 0;
    for (int b = 0; b < blocks_per_mb; ++b) s += partial[((size_t)mb * blocks_per_mb + b) * 2 + c];
    mbstats[i] = s;
}

// ---- fused loss forward + backward ------------------------------------------------------
// Shared-memory plan per tile of M4 samples (feature-major rows of ld = M4+4 floats):
//   X [obs_dim_p] | per net: H_0 .. H_{L-1} (layer outputs; the deltas dZ_l overwrite H_l in place)
//   | per net: dOut [Np_last] | per-sample rows: adv, ret, old_logp, old_val, actions (+ log_std terms)
// Phases per tile: gather | L forward layers | loss head | for l = L-1..0: {dW_l, db_l} then
// {dZ_{l-1} = (dZ_l W_l^T) .* (1 - H_{l-1}^2) in place}.  dW partial sums go to this CTA's packed
// gradient partial in global memory (L2-resident read-modify-write by the owning thread).
// Variants chosen per policy by plan_loss (api.cu):
//   weights_smem  all packed weights staged in shared memory (small nets), else streamed from L2;
//   single_net    wide nets: two passes over the minibatch (actor, then critic) sharing one set of H rows -> 128-sample tile;
//   use_mma       (tile width multiple of 16) layers with Kp, Np multiples of 16 run forward / dH / dW on the tensor-core tiles of
//                 mma_tiles.cuh (3xTF32) instead of the FMA tiles; dW then goes to plane 0 only;
//   stage_thin    with use_mma and streamed weights: the remaining layers' weights staged in shared memory.
struct LossSmem {
    int ld;
    size_t w, x, h[2][DRIL_MAX_LAYERS], dout[2], samp, thin, dbl_bytes_off, total;
};

// Layers that stay on FMA tiles while the wide ones run on MMA tiles with weights streamed from L2 ("thin" layers: the
// input layer, K = obs_dim, and the output layer, N = n_actions | 1): their W | bias and Wt blocks are staged in shared
// memory, compacted in (net, layer) order, so that the few threads working on them do not wait for L2 on every k step.
__host__ __device__ inline bool loss_layer_mma(const LayerDesc& L) { return mma_layer_ok(L.Kp, L.Np); }
__host__ __device__ inline int loss_thin_floats(const PolicyDesc& pd) {
    int n = 0;
    for (int net = 0; net < 2; ++net)
        for (int l = 0; l < pd.n_layers; ++l)
            if (!loss_layer_mma(pd.L[net][l])) n += 2 * pd.L[net][l].Kp * pd.L[net][l].Np + pd.L[net][l].Np;
    return n;
}
// offset (floats) of the [W | bias] block of a thin layer inside the staged region; its Wt block follows directly
__host__ __device__ inline int loss_thin_offset(const PolicyDesc& pd, int net, int l) {
    int n = 0;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < pd.n_layers; ++b) {
            if (a == net && b == l) return n;
            if (!loss_layer_mma(pd.L[a][b])) n += 2 * pd.L[a][b].Kp * pd.L[a][b].Np + pd.L[a][b].Np;
        }
    return n;
}

// single_net: the nets are processed in two passes over the minibatch (actor, then critic) and share one set of
// activation rows, so a wide network keeps a twice as wide sample tile (weights streamed from L2 are then read half
// as often per sample).
__host__ __device__ inline LossSmem loss_smem_layout(const PolicyDesc& pd, int M4, bool weights_smem, bool single_net = false,
                                                     bool stage_thin = false) {
    LossSmem s;
    s.ld = M4 + 4;
    size_t o = 0;
    s.w = o; o += weights_smem ? (size_t)pd.pack_total : 0;
    s.x = o; o += (size_t)pd.obs_dim_p * s.ld;
    if (single_net) {
        for (int l = 0; l < pd.n_layers; ++l) {
            const int np = pd.L[0][l].Np > pd.L[1][l].Np ? pd.L[0][l].Np : pd.L[1][l].Np;
            s.h[0][l] = s.h[1][l] = o; o += (size_t)np * s.ld;
        }
        const int npl = pd.L[0][pd.n_layers - 1].Np > pd.L[1][pd.n_layers - 1].Np ? pd.L[0][pd.n_layers - 1].Np : pd.L[1][pd.n_layers - 1].Np;
        s.dout[0] = s.dout[1] = o; o += (size_t)npl * s.ld;
    } else {
        for (int net = 0; net < 2; ++net)
            for (int l = 0; l < pd.n_layers; ++l) { s.h[net][l] = o; o += (size_t)pd.L[net][l].Np * s.ld; }
        for (int net = 0; net < 2; ++net) { s.dout[net] = o; o += (size_t)pd.L[net][pd.n_layers - 1].Np * s.ld; }
    }
    int arows = pd.act_kind == DRIL_ACT_CONTINUOUS ? 2 * pd.act_n : 1;   // actions (+ log_std grad contributions)
    s.samp = o; o += (size_t)(4 + arows) * s.ld;   // adv, ret, old_logp, old_val, actions...
    o = (o + 3) & ~(size_t)3;
    s.thin = o; o += stage_thin ? (size_t)loss_thin_floats(pd) : 0;
    o = (o + 3) & ~(size_t)3;
    s.dbl_bytes_off = o * sizeof(float);
    s.total = s.dbl_bytes_off + 40 * sizeof(double);
    return s;
}

struct LossArgs {
    PolicyDesc pd;
    BufDev buf;            // obs, actions, advantages, returns, logprobs (old), values (old)
    const float* pack;     // packed W + bias + Wt
    const float* flat;     // log_std
    const double* mbstats; // [2] sum / sum of squares of advantages over the global minibatch
    float* gpart;          // [2 halves][half_stride][pd.gpack] per-CTA packed gradient partials (+ log_std + stats)
    const int* stop_flag;  // target_kl stop already fired: do nothing
    Minibatch mb;
    UpdateHyper hp;
    int M4, weights_smem, half_stride;   // half_stride: CTAs per partial plane
    int single_net;                      // 1: two passes (actor, critic) over the minibatch with shared activation rows
    int use_mma;                         // 1: layers with Kp, Np multiples of 16 run on mma.sync 3xTF32 tiles (needs M4 % 16 == 0)
    int stage_thin;                      // 1: the other layers' weights are staged in shared memory (weights_smem == 0 only)
    int small_splits;                    // sample-range splits of the 4x4 dW tiles (planes 0..small_splits-1)
};

// fire-and-forget vector reduction into this CTA's own gradient partial (no load on the critical path;
// the partial is private to the CTA, so there is no contention and the order of the adds per address is
// the program order of one thread => deterministic)
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// dW[k][n] (+)= sum_m Ain[k][m] * dZ[n][m] for the 4 interleaved rows k = kt + kstride*i and the
// 4 columns n0..n0+3; accumulated into this CTA's packed partial.
__device__ __forceinline__ void dense_tile_dw(const float* __restrict__ Ain, const float* __restrict__ dZ, int m_begin,
                                              int m_end, int ld, int kt, int kstride, int n0, float* __restrict__ gW,
                                              int Np, bool first) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* hp0 = Ain + (size_t)kt * ld;
    const float* dp0 = dZ + (size_t)n0 * ld;
#pragma unroll 2
    for (int m = m_begin; m < m_end; m += 4) {
        float4 h[4], d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = *reinterpret_cast<const float4*>(hp0 + (size_t)i * kstride * ld + m);
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = *reinterpret_cast<const float4*>(dp0 + (size_t)j * ld + m);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[i][j] = fmaf(h[i].x, d[j].x, acc[i][j]);
                acc[i][j] = fmaf(h[i].y, d[j].y, acc[i][j]);
                acc[i][j] = fmaf(h[i].z, d[j].z, acc[i][j]);
                acc[i][j] = fmaf(h[i].w, d[j].w, acc[i][j]);
            }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4* g = reinterpret_cast<float4*>(gW + (size_t)(kt + i * kstride) * Np + n0);
        float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (first) *g = v; else red_add_v4(reinterpret_cast<float*>(g), v);
    }
}

// 8x8 version over the sample range [m_begin, m_end): rows k = kt + 8*i (i<8, consecutive lanes ->
// consecutive rows, 4 banks apart: conflict-free), columns n = 8*nt + j.
__device__ __forceinline__ void dense_tile_dw8(const float* __restrict__ Ain, const float* __restrict__ dZ, int m_begin,
                                               int m_end, int ld, int kt, int kstride, int n0, float* __restrict__ gW,
                                               int Np, bool first) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const float* hp0 = Ain + (size_t)kt * ld;
    const float* dp0 = dZ + (size_t)n0 * ld;
    for (int m = m_begin; m < m_end; m += 4) {
        float4 h[8], d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = *reinterpret_cast<const float4*>(hp0 + (size_t)i * kstride * ld + m);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = *reinterpret_cast<const float4*>(dp0 + (size_t)j * ld + m);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[i][j] = fmaf(h[i].x, d[j].x, acc[i][j]);
                acc[i][j] = fmaf(h[i].y, d[j].y, acc[i][j]);
                acc[i][j] = fmaf(h[i].z, d[j].z, acc[i][j]);
                acc[i][j] = fmaf(h[i].w, d[j].w, acc[i][j]);
            }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float4* g = reinterpret_cast<float4*>(gW + (size_t)(kt + i * kstride) * Np + n0);
        float4 v0 = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        float4 v1 = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        if (first) { g[0] = v0; g[1] = v1; }
        else { red_add_v4(reinterpret_cast<float*>(g), v0); red_add_v4(reinterpret_cast<float*>(g + 1), v1); }
    }
}

// H[k0..k0+3][m0..m0+3] <- (sum_n Wt[n][k0..] * dZ[n][m0..]) * (1 - H[k][m]^2)      (in place)
__device__ __forceinline__ void dense_tile_dh(const float* __restrict__ Wt, int Nred, int Kp, const float* __restrict__ dZ,
                                              float* __restrict__ H, int ld, int k0, int m0) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* wp = Wt + k0;
    const float* dp = dZ + m0;
#pragma unroll 4
    for (int n = 0; n < Nred; ++n, wp += Kp, dp += ld) {
        float4 w = *reinterpret_cast<const float4*>(wp);
        float4 d = *reinterpret_cast<const float4*>(dp);
        acc[0][0] = fmaf(w.x, d.x, acc[0][0]); acc[0][1] = fmaf(w.x, d.y, acc[0][1]);
        acc[0][2] = fmaf(w.x, d.z, acc[0][2]); acc[0][3] = fmaf(w.x, d.w, acc[0][3]);
        acc[1][0] = fmaf(w.y, d.x, acc[1][0]); acc[1][1] = fmaf(w.y, d.y, acc[1][1]);
        acc[1][2] = fmaf(w.y, d.z, acc[1][2]); acc[1][3] = fmaf(w.y, d.w, acc[1][3]);
        acc[2][0] = fmaf(w.z, d.x, acc[2][0]); acc[2][1] = fmaf(w.z, d.y, acc[2][1]);
        acc[2][2] = fmaf(w.z, d.z, acc[2][2]); acc[2][3] = fmaf(w.z, d.w, acc[2][3]);
        acc[3][0] = fmaf(w.w, d.x, acc[3][0]); acc[3][1] = fmaf(w.w, d.y, acc[3][1]);
        acc[3][2] = fmaf(w.w, d.z, acc[3][2]); acc[3][3] = fmaf(w.w, d.w, acc[3][3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4* hp = reinterpret_cast<float4*>(H + (size_t)(k0 + i) * ld + m0);
        float4 h = *hp, o;
        o.x = acc[i][0] * (1.0f - h.x * h.x); o.y = acc[i][1] * (1.0f - h.y * h.y);
        o.z = acc[i][2] * (1.0f - h.z * h.z); o.w = acc[i][3] * (1.0f - h.w * h.w);
        *hp = o;
    }
}

// 8x8 version: rows k in {4kt..} U {kh+4kt..}, cols m in {4mt..} U {mh+4mt..}
__device__ __forceinline__ void dense_tile_dh8(const float* __restrict__ Wt, int Nred, int Kp, const float* __restrict__ dZ,
                                               float* __restrict__ H, int ld, int kt, int mt, int kh, int mh) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const float* w0p = Wt + 4 * kt;
    const float* w1p = Wt + kh + 4 * kt;
    const float* d0p = dZ + 4 * mt;
    const float* d1p = dZ + mh + 4 * mt;
#pragma unroll 2
    for (int n = 0; n < Nred; ++n, w0p += Kp, w1p += Kp, d0p += ld, d1p += ld) {
        float4 w0 = *reinterpret_cast<const float4*>(w0p);
        float4 w1 = *reinterpret_cast<const float4*>(w1p);
        float4 d0 = *reinterpret_cast<const float4*>(d0p);
        float4 d1 = *reinterpret_cast<const float4*>(d1p);
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(w[i], d[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = (i < 4) ? 4 * kt + i : kh + 4 * kt + (i - 4);
        float4* hp0 = reinterpret_cast<float4*>(H + (size_t)k * ld + 4 * mt);
        float4* hp1 = reinterpret_cast<float4*>(H + (size_t)k * ld + mh + 4 * mt);
        float4 h0 = *hp0, h1 = *hp1, o0, o1;
        o0.x = acc[i][0] * (1.0f - h0.x * h0.x); o0.y = acc[i][1] * (1.0f - h0.y * h0.y);
        o0.z = acc[i][2] * (1.0f - h0.z * h0.z); o0.w = acc[i][3] * (1.0f - h0.w * h0.w);
        o1.x = acc[i][4] * (1.0f - h1.x * h1.x); o1.y = acc[i][5] * (1.0f - h1.y * h1.y);
        o1.z = acc[i][6] * (1.0f - h1.z * h1.z); o1.w = acc[i][7] * (1.0f - h1.w * h1.w);
        *hp0 = o0; *hp1 = o1;
    }
}

template <bool WS>
__global__ void __launch_bounds__(DRIL_THREADS) ppo_loss_grad_kernel(const __grid_constant__ LossArgs a) {
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);
    const PolicyDesc& pd = a.pd;
    const BufDev& buf = a.buf;
    const int M4 = a.M4, D = pd.obs_dim, Dp = pd.obs_dim_p, NL = pd.n_layers;
    const LossSmem S = loss_smem_layout(pd, M4, WS, a.single_net != 0, !WS && a.stage_thin);
    const int ld = S.ld;
    const int tid = threadIdx.x;
    float* sX = smem + S.x;
    float* sAdv = smem + S.samp;
    float* sRet = sAdv + ld;
    float* sOldLp = sRet + ld;
    float* sOldV = sOldLp + ld;
    float* sActn = sOldV + ld;                       // discrete: 1 row (int bits); continuous: act_n rows + act_n rows of log_std contributions
    double* sDbl = reinterpret_cast<double*>(reinterpret_cast<char*>(smem) + S.dbl_bytes_off);
    float* gp = a.gpart + (size_t)blockIdx.x * pd.gpack;                       // half 0 partial of this CTA
    float* gp1 = a.gpart + ((size_t)a.half_stride + blockIdx.x) * pd.gpack;    // half 1 (8x8 dW only)

    if (*a.stop_flag) return;

    // WS: weights staged in shared memory (address space known at compile time -> LDS);
    // otherwise streamed from global/L2 through L1
    const float* __restrict__ Wbase = WS ? smem : a.pack;   // S.w == 0
    if (WS) {
        const float4* src = reinterpret_cast<const float4*>(a.pack);
        float4* dst = reinterpret_cast<float4*>(smem);
        for (int i = tid; i < pd.pack_total / 4; i += blockDim.x) dst[i] = src[i];
    }
    const bool thin = !WS && a.stage_thin;
    float* sThin = smem + S.thin;
    if (thin) {
        for (int net = 0; net < 2; ++net)
            for (int l = 0; l < NL; ++l) {
                const LayerDesc& Ld = pd.L[net][l];
                if (loss_layer_mma(Ld)) continue;
                float* dst = sThin + loss_thin_offset(pd, net, l);
                const int nf = Ld.Kp * Ld.Np + Ld.Np, nt = Ld.Kp * Ld.Np;
                for (int i = tid; i < nf; i += blockDim.x) dst[i] = a.pack[Ld.pw_off + i];
                for (int i = tid; i < nt; i += blockDim.x) dst[nf + i] = a.pack[Ld.pwt_off + i];
            }
    }
    // advantage normalisation constants of this (global) minibatch
    float adv_mean = 0.f, adv_den = 1.f;
    if (a.hp.normalize_advantage) {
        double n = a.mb.global_count;
        double mean = a.mbstats[0] / n;
        double var = (a.mbstats[1] - n * mean * mean) / (n - 1.0);   // Bessel-corrected (Julia std)
        if (var < 0.0) var = 0.0;
        adv_mean = (float)mean;
        adv_den = (float)sqrt(var) + 1e-8f;
    }
    const float invB = (float)(1.0 / a.mb.global_count);
    double st_p = 0, st_v = 0, st_e = 0, st_clip = 0, st_kl = 0, st_ratio = 0;   // per-thread stat sums
    double ls_acc = 0;                                                          // thread j < act_n: log_std gradient
    const long long n_tiles = (a.mb.count + M4 - 1) / M4;
    const bool m8 = (M4 & 7) == 0;
    __syncthreads();

    for (int pass = 0; pass < (a.single_net ? 2 : 1); ++pass) {
    const int nmask = a.single_net ? (1 << pass) : 3;        // bit0 actor, bit1 critic: the nets this pass works on
    bool first = true;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, first = false) {
        const long long p0 = a.mb.start + tile * M4;
        const int nvalid = (int)min((long long)M4, a.mb.start + a.mb.count - p0);
        // ---- gather the tile through the Feistel bijection --------------------------------
        if (tid < M4) {
            long long sidx = -1;
            if (tid < nvalid) sidx = a.mb.identity ? (p0 + tid) : feistel_permute(p0 + tid, a.mb.n_total, a.mb.fk);
            float adv = 0.f, ret = 0.f, olp = 0.f, ov = 0.f;
            if (sidx >= 0) {
                adv = buf.advantages[sidx]; ret = buf.returns[sidx]; olp = buf.logprobs[sidx]; ov = buf.values[sidx];
                if (a.hp.normalize_advantage) adv = (adv - adv_mean) / adv_den;
            }
            sAdv[tid] = adv; sRet[tid] = ret; sOldLp[tid] = olp; sOldV[tid] = ov;
            for (int d = 0; d < Dp; ++d) sX[(size_t)d * ld + tid] = (sidx >= 0 && d < D) ? buf.obs[sidx * D + d] : 0.f;
            if (pd.act_kind == DRIL_ACT_DISCRETE) {
                reinterpret_cast<int*>(sActn)[tid] = sidx >= 0 ? reinterpret_cast<const int*>(buf.actions)[sidx] : pd.act_start;
            } else {
                for (int j = 0; j < pd.act_n; ++j)
                    sActn[(size_t)j * ld + tid] = sidx >= 0 ? reinterpret_cast<const float*>(buf.actions)[sidx * pd.act_n + j] : 0.f;
            }
        }
        __syncthreads();
        // ---- forward, all activations kept ------------------------------------------------
        for (int l = 0; l < NL; ++l) {
            const float* ia = l == 0 ? sX : smem + S.h[0][l - 1];
            const float* ic = l == 0 ? sX : smem + S.h[1][l - 1];
            int fmask = nmask;             // nets of this layer left to the FMA tiles
            for (int net = 0; net < 2; ++net) {
                const LayerDesc& Ld = pd.L[net][l];
                if ((nmask >> net & 1) && a.use_mma && mma_layer_ok(Ld.Kp, Ld.Np)) {
                    mma_rows_layer<0>(Wbase + Ld.pw_off, Ld.Np, Ld.Np, Ld.Kp, net ? ic : ia, smem + S.h[net][l], ld, Wbase + Ld.pb_off,
                                      l < NL - 1, M4 >> 3);
                    fmask &= ~(1 << net);
                }
            }
            if (fmask && thin) {           // staged copies: one call per net with the base shifted to the staged [W | bias] block
                for (int net = 0; net < 2; ++net)
                    if (fmask >> net & 1)
                        dense_layer_auto(pd, sThin + loss_thin_offset(pd, net, l) - pd.L[net][l].pw_off, l, ia, ic, smem + S.h[0][l],
                                         smem + S.h[1][l], M4, ld, 1 << net);
            } else if (fmask) {
                dense_layer_auto(pd, Wbase, l, ia, ic, smem + S.h[0][l], smem + S.h[1][l], M4, ld, fmask);
            }
            __syncthreads();
        }
        // ---- loss head: dL/dlogits (or dL/dmean), dL/dvalue ---------------------------------
        float* gA = smem + S.dout[0];
        float* gC = smem + S.dout[1];
        if (tid < M4) {
            const bool valid = tid < nvalid;
            const float* z = smem + S.h[0][NL - 1] + tid;
            const float adv = sAdv[tid];
            float logp = 0.f, ent = 0.f;
            const int A = pd.act_n;
            const int Ap = pd.L[0][NL - 1].Np;
            float g_logp = 0.f;
            const float g_ent = valid ? -a.hp.ent_coef * invB : 0.f;
            float ratio = 1.f, s1 = 0.f, s2 = 0.f, log_ratio = 0.f, rc = 1.f;
            if (!(nmask & 1)) {
                // critic-only pass
            } else if (pd.act_kind == DRIL_ACT_DISCRETE) {
                int aidx = reinterpret_cast<const int*>(sActn)[tid] - pd.act_start;
                aidx = aidx < 0 ? 0 : (aidx >= A ? A - 1 : aidx);
                float m = z[0];
                for (int j = 1; j < A; ++j) m = fmaxf(m, z[(size_t)j * ld]);
                float s = 0.f;
                for (int j = 0; j < A; ++j) s += expf(z[(size_t)j * ld] - m);
                float h = 0.f;
                for (int j = 0; j < A; ++j) { float p = expf(z[(size_t)j * ld] - m) / s; h += p * logf(p); }
                ent = -h;
                logp = logf(expf(z[(size_t)aidx * ld] - m) / s);
                log_ratio = logp - sOldLp[tid];
                ratio = expf(log_ratio);
                rc = fminf(fmaxf(ratio, 1.0f - a.hp.clip_range), 1.0f + a.hp.clip_range);
                s1 = ratio * adv; s2 = rc * adv;
                g_logp = (!valid || s2 < s1) ? 0.f : -invB * adv * ratio;   // min(s1,s2): ties -> s1
                for (int j = 0; j < Ap; ++j) {
                    float dz = 0.f;
                    if (j < A) {
                        float p = expf(z[(size_t)j * ld] - m) / s;
                        dz = g_logp * ((j == aidx ? 1.0f : 0.0f) - p) + g_ent * (-p * (logf(p) + ent));
                    }
                    gA[(size_t)j * ld + tid] = dz;
                }
            } else {
                float ls_sum = 0.f, dss = 0.f;
                for (int j = 0; j < A; ++j) {
                    float ls = a.flat[pd.log_std_off + j];
                    float diff = sActn[(size_t)j * ld + tid] - z[(size_t)j * ld];
                    dss += diff * diff * expf(-2.0f * ls);
                    ls_sum += ls;
                }
                logp = -0.5f * (2.0f * ls_sum + dss + (float)A * DRIL_LOG2PI);
                ent = 0.5f * (float)A * (1.0f + DRIL_LOG2PI) + ls_sum;
                log_ratio = logp - sOldLp[tid];
                ratio = expf(log_ratio);
                rc = fminf(fmaxf(ratio, 1.0f - a.hp.clip_range), 1.0f + a.hp.clip_range);
                s1 = ratio * adv; s2 = rc * adv;
                g_logp = (!valid || s2 < s1) ? 0.f : -invB * adv * ratio;
                for (int j = 0; j < Ap; ++j) {
                    float dz = 0.f;
                    if (j < A) {
                        float ls = a.flat[pd.log_std_off + j];
                        float vi = expf(-2.0f * ls);
                        float diff = sActn[(size_t)j * ld + tid] - z[(size_t)j * ld];
                        dz = g_logp * diff * vi;
                        sActn[(size_t)(A + j) * ld + tid] = g_logp * (-1.0f + diff * diff * vi) + g_ent;   // d/dlog_std_j
                    }
                    gA[(size_t)j * ld + tid] = dz;
                }
            }
            // critic
            if (nmask & 2) {
                const float v_raw = smem[S.h[1][NL - 1] + tid];
                float v = v_raw;
                bool v_pass = true;
                if (a.hp.clip_range_vf >= 0.f) {
                    float dlt = v_raw - sOldV[tid];
                    v_pass = dlt >= -a.hp.clip_range_vf && dlt <= a.hp.clip_range_vf;
                    v = sOldV[tid] + fminf(fmaxf(dlt, -a.hp.clip_range_vf), a.hp.clip_range_vf);
                }
                const float verr = v - sRet[tid];
                const float g_val = (valid && v_pass) ? a.hp.vf_coef * 2.0f * verr * invB : 0.f;
                const int Cp = pd.L[1][NL - 1].Np;
                for (int j = 0; j < Cp; ++j) gC[(size_t)j * ld + tid] = j == 0 ? g_val : 0.f;
                if (valid) st_v += (double)(verr * verr);
            }
            if (valid && (nmask & 1)) {
                st_p += (double)(-fminf(s1, s2));
                st_e += (double)ent;
                st_clip += (ratio != rc) ? 1.0 : 0.0;
                st_kl += (double)(expf(log_ratio) - 1.0f - log_ratio);
                st_ratio += (double)ratio;
            }
        }
        __syncthreads();
        if ((nmask & 1) && pd.act_kind == DRIL_ACT_CONTINUOUS && tid < pd.act_n) {
            double s = 0;
            for (int e = 0; e < nvalid; ++e) s += (double)sActn[(size_t)(pd.act_n + tid) * ld + e];
            ls_acc += s;
        }
        // ---- backward -----------------------------------------------------------------------
        for (int l = NL - 1; l >= 0; --l) {
            const LayerDesc& La = pd.L[0][l];
            const LayerDesc& Lc = pd.L[1][l];
            const float* dZa = l == NL - 1 ? gA : smem + S.h[0][l];
            const float* dZc = l == NL - 1 ? gC : smem + S.h[1][l];
            // phase 1: dW_l, db_l (reads the layer input and dZ_l)
            {
                const bool a8 = m8 && (M4 & 15) == 0 && (La.Kp & 7) == 0 && (La.Np & 7) == 0;
                const bool c8 = m8 && (M4 & 15) == 0 && (Lc.Kp & 7) == 0 && (Lc.Np & 7) == 0;
                const int nsplit = a.small_splits;
                const bool amma = a.use_mma && mma_layer_ok(La.Kp, La.Np), cmma = a.use_mma && mma_layer_ok(Lc.Kp, Lc.Np);
                if ((nmask & 1) && amma) mma_dw_layer(l == 0 ? sX : smem + S.h[0][l - 1], dZa, ld, La.Kp, La.Np, gp + La.pw_off, first, M4);
                if ((nmask & 2) && cmma) mma_dw_layer(l == 0 ? sX : smem + S.h[1][l - 1], dZc, ld, Lc.Kp, Lc.Np, gp + Lc.pw_off, first, M4);
                const int ca = (!(nmask & 1) || amma) ? 0 : (a8 ? (La.Kp >> 3) * (La.Np >> 3) * 2 : (La.Kp >> 2) * (La.Np >> 2) * nsplit);
                const int cc = (!(nmask & 2) || cmma) ? 0 : (c8 ? (Lc.Kp >> 3) * (Lc.Np >> 3) * 2 : (Lc.Kp >> 2) * (Lc.Np >> 2) * nsplit);
                const int nba = (nmask & 1) ? La.N : 0;
                const int cb = nba + ((nmask & 2) ? Lc.N : 0);
                for (int t = tid; t < ca + cc + cb; t += blockDim.x) {
                    if (t < ca + cc) {
                        const bool crit = t >= ca;
                        int u = crit ? t - ca : t;
                        const LayerDesc& Ld = crit ? Lc : La;
                        const float* Ain = l == 0 ? sX : smem + S.h[crit][l - 1];
                        const float* dZ = crit ? dZc : dZa;
                        if (crit ? c8 : a8) {
                            const int kq = Ld.Kp >> 3, per_half = kq * (Ld.Np >> 3);
                            const int half = u / per_half;
                            u -= half * per_half;
                            const int nt = u / kq, kt = u - nt * kq;
                            const int mh = M4 >> 1;
                            dense_tile_dw8(Ain, dZ, half * mh, half * mh + mh, ld, kt, kq, nt << 3,
                                           (half ? gp1 : gp) + Ld.pw_off, Ld.Np, first);
                        } else {
                            // small layer: the sample range is split so that all threads get a tile; split s
                            // accumulates into partial plane s of this CTA
                            const int kq = Ld.Kp >> 2, per_split = kq * (Ld.Np >> 2);
                            const int sp = u / per_split;
                            u -= sp * per_split;
                            const int nt = u / kq, kt = u - nt * kq;
                            const int ms = M4 / nsplit;
                            float* gq = a.gpart + ((size_t)sp * a.half_stride + blockIdx.x) * pd.gpack;
                            dense_tile_dw(Ain, dZ, sp * ms, sp * ms + ms, ld, kt, kq, nt << 2, gq + Ld.pw_off, Ld.Np, first);
                        }
                    } else {
                        int u = t - ca - cc;
                        const bool crit = u >= nba;
                        if (crit) u -= nba;
                        const LayerDesc& Ld = crit ? Lc : La;
                        const float* dZ = (crit ? dZc : dZa) + (size_t)u * ld;
                        float s = 0.f;
                        for (int m = 0; m < M4; m += 4) {
                            float4 d = *reinterpret_cast<const float4*>(dZ + m);
                            s += (d.x + d.y) + (d.z + d.w);
                        }
                        float* g = gp + Ld.pb_off + u;
                        if (first) *g = s; else atomicAdd(g, s);
                    }
                }
            }
            __syncthreads();
            // phase 2: dZ_{l-1} = (dZ_l W_l^T) .* (1 - H_{l-1}^2), in place over H_{l-1}
            if (l > 0) {
                const bool a8 = m8 && (La.Kp & 7) == 0 && La.N >= 8;
                const bool c8 = m8 && (Lc.Kp & 7) == 0 && Lc.N >= 8;
                const int mt4 = M4 >> 2, mt8 = M4 >> 3;
                const bool amma = a.use_mma && mma_layer_ok(La.Kp, La.Np), cmma = a.use_mma && mma_layer_ok(Lc.Kp, Lc.Np);
                if ((nmask & 1) && amma) mma_rows_layer<1>(Wbase + La.pwt_off, La.Kp, La.Kp, La.Np, dZa, smem + S.h[0][l - 1], ld, nullptr, false, M4 >> 3);
                if ((nmask & 2) && cmma) mma_rows_layer<1>(Wbase + Lc.pwt_off, Lc.Kp, Lc.Kp, Lc.Np, dZc, smem + S.h[1][l - 1], ld, nullptr, false, M4 >> 3);
                const int ca = (!(nmask & 1) || amma) ? 0 : (a8 ? (La.Kp >> 3) * mt8 : (La.Kp >> 2) * mt4);
                const int cc = (!(nmask & 2) || cmma) ? 0 : (c8 ? (Lc.Kp >> 3) * mt8 : (Lc.Kp >> 2) * mt4);
                // staged Wt blocks of this layer (only read when the layer is not on MMA tiles)
                const float* sWtA = sThin + loss_thin_offset(pd, 0, l) + La.Kp * La.Np + La.Np;
                const float* sWtC = sThin + loss_thin_offset(pd, 1, l) + Lc.Kp * Lc.Np + Lc.Np;
                for (int t = tid; t < ca + cc; t += blockDim.x) {
                    const bool crit = t >= ca;
                    const int u = crit ? t - ca : t;
                    const LayerDesc& Ld = crit ? Lc : La;
                    const float* dZ = crit ? dZc : dZa;
                    float* H = smem + S.h[crit][l - 1];
                    const float* Wt = thin ? (crit ? sWtC : sWtA) : Wbase + Ld.pwt_off;
                    if (crit ? c8 : a8) {
                        const int kt = u / mt8, m = u - kt * mt8;
                        dense_tile_dh8(Wt, Ld.N, Ld.Kp, dZ, H, ld, kt, m, Ld.Kp >> 1, M4 >> 1);
                    } else {
                        const int kt = u / mt4, m = u - kt * mt4;
                        dense_tile_dh(Wt, Ld.N, Ld.Kp, dZ, H, ld, kt << 2, m << 2);
                    }
                }
                __syncthreads();
            }
        }
    }
    __syncthreads();       // the next pass re-stages the sample rows the last backward phase may still read
    }
    // ---- per-CTA tail: log_std gradient + statistic sums --------------------------------------
    if (pd.act_kind == DRIL_ACT_CONTINUOUS && tid < pd.act_n) gp[pd.pack_fwd + tid] = (float)ls_acc;
    double sums[6] = {st_p, st_v, st_e, st_clip, st_kl, st_ratio};
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = block_sum(sums[i], sDbl);
        if (tid == 0) gp[pd.pack_fwd + pd.act_n + i] = (float)s;
    }
}

// iteration accumulators (device): [0..6] sums over applied minibatches of policy_loss, value_loss,
// entropy_loss, clip_fraction, approx_kl, entropy, ratio; [7] loss; [8] grad_norm sum; [9] applied
// count; [10] grad_norm count; [12],[13] running beta1^t, beta2^t
#define ITER_ACC_N 16

struct AdamArgs {
    float* g;            // [n_params + 6]: gradient then the six stat sums (already summed over ranks)
    float* flat;
    float* m;
    float* v;
    float* pack;
    const int* flat2pack;
    const int* flat2packT;   // -1 for biases / log_std
    long long* step;
    double* iter_acc;
    int* stop_flag;
    double global_count;
    UpdateHyper hp;
    int n_params;
    int apply_stats;     // 0 for the dril_optimizer_step parity entry (no stat bookkeeping)
};

// KL early stop of ppo.jl:235-238, decided BEFORE the step is applied, from the reduced statistics in g[n_params..]
__device__ __forceinline__ int adam_stop(const AdamArgs& a, double global_count) {
    const float kl = a.g[a.n_params + 4] * (float)(1.0 / global_count);
    return (a.apply_stats && a.hp.target_kl >= 0.f && kl > 1.5f * a.hp.target_kl) ? 1 : 0;
}
__device__ __forceinline__ int adam_stop(const AdamArgs& a) { return adam_stop(a, a.global_count); }
// The per-iteration accumulators (learn_stats sums, applied-step count, running beta^t, step counter, stop flag):
// threads 0..15 of ONE CTA, independent global round trips.  s_f[0..1] receive the Adam bias corrections 1 - beta^t.
// `old` = iter_acc[tid] and invB = 1 / global_count may be fetched / computed by the caller ahead of time (they do not depend on
// the reduced gradient): the fused tail does so before its second grid barrier.
__device__ __forceinline__ void adam_accumulate_pre(const AdamArgs& a, float norm, int stop, int tid, float* s_f, double old, float invB);
__device__ __forceinline__ void adam_accumulate(const AdamArgs& a, float norm, int stop, int tid, float* s_f) {
    if (tid >= 16) return;
    adam_accumulate_pre(a, norm, stop, tid, s_f, a.iter_acc[tid], (float)(1.0 / a.global_count));
}
__device__ __forceinline__ void adam_accumulate_pre(const AdamArgs& a, float norm, int stop, int tid, float* s_f, double old, float invB) {
    if (tid >= 16) return;
    // all loads first and unconditionally (one round trip): the per-thread cases below are pure arithmetic, so the
    // divergent switch does not serialise a global load per case
    const float* st = a.g + a.n_params;
    const float st0 = __ldcg(st), st1 = __ldcg(st + 1), st2 = __ldcg(st + 2), st3 = __ldcg(st + 3), st4 = __ldcg(st + 4), st5 = __ldcg(st + 5);
    if (a.apply_stats) {
        const float p_loss = st0 * invB, v_loss = st1 * invB, ent = st2 * invB;
        const float ent_loss = -ent;
        double add = 0.0;
        bool on_apply = true;
        switch (tid) {
            case 0: add = p_loss; break;
            case 1: add = v_loss; break;
            case 2: add = ent_loss; break;
            case 3: add = st3 * invB; break;
            case 4: add = st4 * invB; break;
            case 5: add = ent; break;
            case 6: add = st5 * invB; break;
            case 7: add = p_loss + a.hp.ent_coef * ent_loss + a.hp.vf_coef * v_loss; break;
            case 8: add = norm; on_apply = false; break;     // grad_norms gets the PRE-clip norm before the KL check (ppo.jl:216-223)
            case 9: add = 1.0; break;
            case 10: add = 1.0; on_apply = false; break;
            default: add = 0.0; on_apply = false; break;
        }
        if (tid <= 10 && (!on_apply || !stop)) a.iter_acc[tid] = old + add;
    } else if (tid == 8) {
        a.iter_acc[8] = (double)norm;
    }
    if (!stop && (tid == 12 || tid == 13)) {
        double b = tid == 12 ? (double)a.hp.beta1 : (double)a.hp.beta2;
        double pw = old * b;
        a.iter_acc[tid] = pw;
        if (s_f) s_f[tid - 12] = (float)(1.0 - pw);
    }
    if (tid == 14) {
        if (stop) *a.stop_flag = 1; else *a.step += 1;
    }
}
// Adam (Optimisers.Adam, eps 1e-5: ppo.jl:64-66) on one parameter + refresh of its packed copies
__device__ __forceinline__ void adam_param(const AdamArgs& a, int p, float scale, float c1, float c2) {
    const float b1 = a.hp.beta1, b2 = a.hp.beta2;
    float g = __fmul_rn(a.g[p], scale);
    float m = __fadd_rn(__fmul_rn(b1, a.m[p]), __fmul_rn(__fsub_rn(1.0f, b1), g));
    float v = __fadd_rn(__fmul_rn(b2, a.v[p]), __fmul_rn(__fmul_rn(__fsub_rn(1.0f, b2), g), g));
    float upd = __fmul_rn(__fdiv_rn(__fdiv_rn(m, c1), __fadd_rn(__fsqrt_rn(__fdiv_rn(v, c2)), a.hp.adam_eps)), a.hp.lr);
    float w = __fsub_rn(a.flat[p], upd);
    int ip = a.flat2pack[p];
    int it = a.flat2packT[p];
    a.m[p] = m; a.v[p] = v; a.flat[p] = w;
    if (ip >= 0) a.pack[ip] = w;
    if (it >= 0) a.pack[it] = w;
}

// One CTA: (given sum of squares q) clip scale -> KL stop -> statistics -> Adam -> refresh packed layouts.
__device__ __forceinline__ void adam_apply(const AdamArgs& a, double q, float* s_f, int* s_i) {
    const int tid = threadIdx.x;
    const float norm = (float)sqrt(q);
    const int stop = adam_stop(a);
    adam_accumulate(a, norm, stop, tid, s_f);
    if (tid == 14) *s_i = stop;
    __syncthreads();
    if (stop) return;
    float scale = 1.f;
    if (a.hp.max_grad_norm >= 0.f && norm > a.hp.max_grad_norm) scale = a.hp.max_grad_norm / norm;   // no epsilon (optimization_utils.jl:99-107)
    const float c1 = s_f[0], c2 = s_f[1];
#pragma unroll 4
    for (int p = tid; p < a.n_params; p += blockDim.x) adam_param(a, p, scale, c1, c2);
}

// stand-alone (multi-GPU path after the NCCL allreduce, and the dril_optimizer_step parity entry)
__global__ void __launch_bounds__(1024) adam_finalize_kernel(AdamArgs a) {
    __shared__ double scratch[32];
    __shared__ float s_f[2];
    __shared__ int s_i;
    if (*a.stop_flag) return;
    double q = 0;
    for (int p = threadIdx.x; p < a.n_params; p += blockDim.x) { double g = a.g[p]; q += g * g; }
    q = block_sum(q, scratch);
    adam_apply(a, q, s_f, &s_i);
}

// single-GPU path: partial reduction over CTAs/planes by many CTAs, then the last CTA to finish
// (ticket) computes the global norm from the per-CTA sums of squares and applies clip + Adam.
// Block = 32 parameters x 32 CTA-groups: lanes read 32 consecutive packed entries (coalesced), the 32
// warps split the (plane, CTA) list 32 ways; partial sums meet in shared memory in a fixed order.
#define RA_PARAMS_PER_BLOCK 32
#define RA_GROUPS 32
__global__ void __launch_bounds__(1024) reduce_adam_kernel(const float* __restrict__ gpart, int n_cta, int half_stride,
                                                          int gpack, const int* __restrict__ flat2g,
                                                          const unsigned char* __restrict__ f2planes, int stats_off,
                                                          double* __restrict__ sq_part, unsigned int* __restrict__ ticket,
                                                          AdamArgs a, int do_adam, float* p2p_gbuf, int p2p_slots,
                                                          unsigned long long* p2p_seq, volatile unsigned long long* p2p_flag) {
    __shared__ float s_red[RA_GROUPS][RA_PARAMS_PER_BLOCK + 1];
    __shared__ double scratch[32];
    __shared__ float s_f[2];
    __shared__ int s_i;
    __shared__ unsigned int s_ticket;
    if (*a.stop_flag) return;
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int p = blockIdx.x * RA_PARAMS_PER_BLOCK + lane;
    float part = 0.f;
    if (p < a.n_params + 6) {
        const int idx = p < a.n_params ? flat2g[p] : stats_off + (p - a.n_params);
        const int planes = p < a.n_params ? f2planes[p] : 1;
        for (int pl = 0; pl < planes; ++pl) {
            const float* base = gpart + (size_t)pl * half_stride * gpack + idx;
            // up to 8 independent loads in flight per thread (n_cta <= 8 * RA_GROUPS)
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = grp + j * RA_GROUPS;
                v[j] = c < n_cta ? base[(size_t)c * gpack] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) part += v[j];
            for (int c = grp + 8 * RA_GROUPS; c < n_cta; c += RA_GROUPS) part += base[(size_t)c * gpack];
        }
    }
    s_red[grp][lane] = part;
    __syncthreads();
    if (grp == 0) {
        double sq = 0.0;
        if (p < a.n_params + 6) {
            float g = 0.f;
#pragma unroll
            for (int j = 0; j < RA_GROUPS; ++j) g += s_red[j][lane];
            if (do_adam == 2) {
                // P2P path: this rank's contribution goes to the peer-visible buffer of the NEXT sequence number
                p2p_gbuf[(size_t)((*p2p_seq + 1ull) & 1ull) * p2p_slots + p] = g;
                __threadfence();          // gpu scope; the last CTA's system-scope fence below is cumulative over these
            } else {
                a.g[p] = g;
                if (p < a.n_params) sq = (double)g * (double)g;
                __threadfence();          // publish this thread's slice before the ticket
            }
        }
        if (do_adam) {
            sq = warp_sum(sq);
            if (lane == 0) {
                sq_part[blockIdx.x] = sq;
                __threadfence();
                s_ticket = atomicAdd(ticket, 1u);
            }
        }
    }
    if (!do_adam) return;                 // multi-GPU without peer access: NCCL allreduce of g, then adam_finalize_kernel
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    if (do_adam == 2) {                   // last CTA: everything of this rank is written -> publish the sequence number
        if (threadIdx.x == 0) {
            *ticket = 0u;
            __threadfence_system();
            const unsigned long long sq_ = *p2p_seq + 1ull;
            *p2p_seq = sq_;
            __threadfence_system();
            *p2p_flag = sq_;
            __threadfence_system();
        }
        return;
    }
    __threadfence();
    if (threadIdx.x == 0) *ticket = 0u;
    // global sum of squares: one load per thread + fixed-order block reduction (a per-thread serial
    // loop over gridDim.x L2 round trips was the critical path of this kernel)
    double q = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) q += sq_part[b];
    q = block_sum(q, scratch);
    adam_apply(a, q, s_f, &s_i);
}

// ---------------------------------------------------------------------------------------
// Multi-GPU: one-shot gradient allreduce over NVLink peer memory, fused into the reduce / Adam kernels
// (replaces reduce -> ncclAllReduce -> Adam for the 36.6 KB latency-bound message).
// Every rank owns a region {gbuf[2][n_slots], seq, flag} mapped into all peers with CUDA IPC.
//   kernel R (reduce_adam_kernel, do_adam = 2): partial planes -> LOCAL gbuf[(seq+1)&1]; the last CTA
//            (ticket) bumps seq and publishes flag = seq with a system-scope fence.
//   kernel X (p2p_sum_adam_kernel): waits until every peer's flag >= seq, sums the peers' gbuf in rank
//            order (identical on all ranks => bit-identical parameters without broadcast), per-CTA
//            sums of squares, ticket, last CTA: clip + KL stop + Adam.
// Double buffering is enough: a rank can only start step s+2 after every peer published s+1, i.e. after
// every peer finished reading step s.  Spins are bounded; a timeout raises err (checked on the host).
// ---------------------------------------------------------------------------------------
#define DRIL_MAX_RANKS 16
struct P2PDev {
    float* local_gbuf;                         // [2][n_slots]
    unsigned long long* local_seq;             // device-side step sequence of this rank
    volatile unsigned long long* local_flag;   // published sequence (read by peers)
    const float* peer_gbuf[DRIL_MAX_RANKS];
    const volatile unsigned long long* peer_flag[DRIL_MAX_RANKS];
    int* err;
    int n_slots, nranks, rank;
    // push exchange (fused tail of the tensor-core kernel): every rank owns recv[2][nranks][n_slots] and per-(source rank,
    // CTA) arrival flags; rank r's CTA b stores its reduced slice into EVERY rank's recv[parity][r] and then raises
    // that rank's cflag[r][b], so a receiver only polls and reads its own memory (one NVLink one-way latency)
    float* peer_recv[DRIL_MAX_RANKS];
    unsigned long long* peer_cflag[DRIL_MAX_RANKS];
    const float* local_recv;
    const volatile unsigned long long* local_cflag;
    int max_cta;
    // small fp64 allreduce (minibatch advantage moments, explained-variance moments): srecv[2][nranks][P2P_SMALL_MAX]
    double* peer_srecv[DRIL_MAX_RANKS];
    unsigned long long* peer_sflag[DRIL_MAX_RANKS];     // [nranks]
    const double* local_srecv;
    const volatile unsigned long long* local_sflag;
    unsigned long long* small_seq;
};
#define P2P_SMALL_MAX 256

// In-place sum over ranks of up to P2P_SMALL_MAX doubles (two arrays back to back), as a push over peer memory:
// one CTA; every rank stores its values into every rank's srecv[parity][rank], raises sflag[rank], waits for all
// ranks' flags in its own memory and sums in rank order (bit-identical on every rank).  ~5 us instead of ~30 us NCCL.
__global__ void __launch_bounds__(P2P_SMALL_MAX) p2p_small_allreduce_kernel(P2PDev pp, double* __restrict__ a, int na,
                                                                            double* __restrict__ b, int nb) {
    const int tid = threadIdx.x, n = na + nb;
    const unsigned long long seq = *pp.small_seq + 1ull;
    const size_t par_off = (size_t)(seq & 1ull) * pp.nranks * P2P_SMALL_MAX;
    if (tid < n) {
        const double v = tid < na ? a[tid] : b[tid - na];
        for (int r = 0; r < pp.nranks; ++r) pp.peer_srecv[r][par_off + (size_t)pp.rank * P2P_SMALL_MAX + tid] = v;
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        for (int r = 0; r < pp.nranks; ++r) *reinterpret_cast<volatile unsigned long long*>(pp.peer_sflag[r] + pp.rank) = seq;
    }
    if (tid < pp.nranks) {
        const volatile unsigned long long* f = pp.local_sflag + tid;
        long long spins = 0;
        while (*f < seq) {
            __nanosleep(32);
            if (++spins > (1ll << 24)) { atomicExch(pp.err, 1); break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (tid < n) {
        double s = 0.0;
        for (int r = 0; r < pp.nranks; ++r) s += __ldcv(pp.local_srecv + par_off + (size_t)r * P2P_SMALL_MAX + tid);
        if (tid < na) a[tid] = s; else b[tid - na] = s;
    }
    if (tid == 0) *pp.small_seq = seq;
}

__global__ void __launch_bounds__(1024) p2p_sum_adam_kernel(P2PDev pp, double* __restrict__ sq_part,
                                                           unsigned int* __restrict__ ticket, AdamArgs a) {
    __shared__ double scratch[32];
    __shared__ float s_f[2];
    __shared__ int s_i;
    __shared__ unsigned int s_ticket;
    if (*a.stop_flag) return;
    const unsigned long long seq = *pp.local_seq;
    if (threadIdx.x < pp.nranks) {
        const volatile unsigned long long* f = pp.peer_flag[threadIdx.x];
        long long spins = 0;
        while (*f < seq) {
            __nanosleep(64);
            if (++spins > (1ll << 24)) { atomicExch(pp.err, 1); break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = a.n_params + 6;
    double sq = 0.0;
    if (p < n) {
        const size_t off = (size_t)(seq & 1ull) * pp.n_slots + p;
        float v[DRIL_MAX_RANKS];
#pragma unroll
        for (int r = 0; r < DRIL_MAX_RANKS; ++r) v[r] = r < pp.nranks ? __ldcv(pp.peer_gbuf[r] + off) : 0.f;
        float g = 0.f;
#pragma unroll
        for (int r = 0; r < DRIL_MAX_RANKS; ++r) g += v[r];
        a.g[p] = g;
        if (p < a.n_params) sq = (double)g * (double)g;
        __threadfence();
    }
    sq = block_sum(sq, scratch);
    if (threadIdx.x == 0) {
        sq_part[blockIdx.x] = sq;
        __threadfence();
        s_ticket = atomicAdd(ticket, 1u);
    }
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    if (threadIdx.x == 0) *ticket = 0u;
    double q = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) q += sq_part[b];
    q = block_sum(q, scratch);
    adam_apply(a, q, s_f, &s_i);
}

// (re)build the packed layouts from the flat vector (after set_params)
__global__ void repack_kernel(const float* __restrict__ flat, float* __restrict__ pack, const int* __restrict__ flat2pack,
                              const int* __restrict__ flat2packT, int n_params) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_params) return;
    float w = flat[p];
    int ip = flat2pack[p];
    if (ip >= 0) pack[ip] = w;
    int it = flat2packT[p];
    if (it >= 0) pack[it] = w;
}
