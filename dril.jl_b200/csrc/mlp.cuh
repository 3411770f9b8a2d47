// Tile-level small-MLP primitives (fp32 CUDA cores, register-tiled, operands in shared memory).
//
// Data layout inside a CTA: activations are FEATURE-MAJOR, act[f * ld + m] (m = sample within
// the tile, ld = tile width + 4 so that rows start 4 banks apart), weights use the packed
// layout of PolicyDesc: W[k * Np + n] (== Lux Dense weight (out,in) column-major, zero padded
// to multiples of 4) and Wt[n * Kp + k].  With these, every operand of the three GEMM shapes
// (forward, dH = dZ·Wᵀ, dW = Hᵀ·dZ) is read as a float4.
//
// Actor and critic (two separate MLPs of the same depth, layers/layer_helpers.jl:27-57) are
// processed together: one pass over the thread-tiles of both nets per layer.
#pragma once
#include "common.cuh"
#include "mma_tiles.cuh"

// C[n0..n0+3][m0..m0+3] = f( sum_k W[k][n0..] * A[k][m0..] + b[n0..] )
__device__ __forceinline__ void dense_tile_fwd(const float* __restrict__ W, const float* __restrict__ bias,
                                               int K, int Np, const float* __restrict__ A, float* __restrict__ C,
                                               int ld, int n0, int m0, bool apply_tanh) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* wp = W + n0;
    const float* ap = A + m0;
#pragma unroll 4
    for (int k = 0; k < K; ++k, wp += Np, ap += ld) {
        float4 w = *reinterpret_cast<const float4*>(wp);
        float4 a = *reinterpret_cast<const float4*>(ap);
        acc[0][0] = fmaf(w.x, a.x, acc[0][0]); acc[0][1] = fmaf(w.x, a.y, acc[0][1]);
        acc[0][2] = fmaf(w.x, a.z, acc[0][2]); acc[0][3] = fmaf(w.x, a.w, acc[0][3]);
        acc[1][0] = fmaf(w.y, a.x, acc[1][0]); acc[1][1] = fmaf(w.y, a.y, acc[1][1]);
        acc[1][2] = fmaf(w.y, a.z, acc[1][2]); acc[1][3] = fmaf(w.y, a.w, acc[1][3]);
        acc[2][0] = fmaf(w.z, a.x, acc[2][0]); acc[2][1] = fmaf(w.z, a.y, acc[2][1]);
        acc[2][2] = fmaf(w.z, a.z, acc[2][2]); acc[2][3] = fmaf(w.z, a.w, acc[2][3]);
        acc[3][0] = fmaf(w.w, a.x, acc[3][0]); acc[3][1] = fmaf(w.w, a.y, acc[3][1]);
        acc[3][2] = fmaf(w.w, a.z, acc[3][2]); acc[3][3] = fmaf(w.w, a.w, acc[3][3]);
    }
    float4 b = *reinterpret_cast<const float4*>(bias + n0);
    float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 o;
        o.x = acc[i][0] + bb[i]; o.y = acc[i][1] + bb[i]; o.z = acc[i][2] + bb[i]; o.w = acc[i][3] + bb[i];
        if (apply_tanh) { o.x = fast_tanh(o.x); o.y = fast_tanh(o.y); o.z = fast_tanh(o.z); o.w = fast_tanh(o.w); }
        *reinterpret_cast<float4*>(C + (size_t)(n0 + i) * ld + m0) = o;
    }
}

// 8x8 register tile: rows n in {4nt..4nt+3} U {nh+4nt..}, cols m in {4mt..4mt+3} U {mh+4mt..}.
// 64 FMA per 4 LDS.128 (1 B of shared-memory traffic per FMA: the LDS pipe moves 128 B/clk/SM, the
// FMA pipes 128 FMA/clk/SM, so anything smaller than 8x8 is LDS-bound — profiles/r01).
// Lane mapping (mt fastest) makes the A loads 8 consecutive float4 per quarter-warp and the W
// loads a broadcast: conflict-free.
__device__ __forceinline__ void dense_tile_fwd8(const float* __restrict__ W, const float* __restrict__ bias,
                                                int K, int Np, const float* __restrict__ A, float* __restrict__ C,
                                                int ld, int nt, int mt, int nh, int mh, bool apply_tanh) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const float* w0p = W + 4 * nt;
    const float* w1p = W + nh + 4 * nt;
    const float* a0p = A + 4 * mt;
    const float* a1p = A + mh + 4 * mt;
#pragma unroll 2
    for (int k = 0; k < K; ++k, w0p += Np, w1p += Np, a0p += ld, a1p += ld) {
        float4 w0 = *reinterpret_cast<const float4*>(w0p);
        float4 w1 = *reinterpret_cast<const float4*>(w1p);
        float4 a0 = *reinterpret_cast<const float4*>(a0p);
        float4 a1 = *reinterpret_cast<const float4*>(a1p);
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(w[i], a[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = (i < 4) ? 4 * nt + i : nh + 4 * nt + (i - 4);
        const float b = bias[n];
        float4 o0, o1;
        o0.x = acc[i][0] + b; o0.y = acc[i][1] + b; o0.z = acc[i][2] + b; o0.w = acc[i][3] + b;
        o1.x = acc[i][4] + b; o1.y = acc[i][5] + b; o1.z = acc[i][6] + b; o1.w = acc[i][7] + b;
        if (apply_tanh) {
            o0.x = fast_tanh(o0.x); o0.y = fast_tanh(o0.y); o0.z = fast_tanh(o0.z); o0.w = fast_tanh(o0.w);
            o1.x = fast_tanh(o1.x); o1.y = fast_tanh(o1.y); o1.z = fast_tanh(o1.z); o1.w = fast_tanh(o1.w);
        }
        *reinterpret_cast<float4*>(C + (size_t)n * ld + 4 * mt) = o0;
        *reinterpret_cast<float4*>(C + (size_t)n * ld + mh + 4 * mt) = o1;
    }
}

// layer forward with 8x8 tiles where the shapes allow it (Np, M4 multiples of 8 and enough tiles to
// occupy the block), 4x4 tiles otherwise.
__device__ __forceinline__ void dense_layer_auto(const PolicyDesc& pd, const float* __restrict__ Wbase, int layer,
                                                 const float* in_a, const float* in_c, float* out_a, float* out_c,
                                                 int M4, int ld, int net_mask) {
    const LayerDesc& La = pd.L[0][layer];
    const LayerDesc& Lc = pd.L[1][layer];
    const bool act = layer < pd.n_layers - 1;
    const bool m8 = (M4 & 7) == 0;
    const int mt4 = M4 >> 2, mt8 = M4 >> 3;
    const bool a8 = (net_mask & 1) && m8 && (La.Np & 7) == 0 && La.K >= 8;
    const bool c8 = (net_mask & 2) && m8 && (Lc.Np & 7) == 0 && Lc.K >= 8;
    const int ta = (net_mask & 1) ? (a8 ? (La.Np >> 3) * mt8 : (La.Np >> 2) * mt4) : 0;
    const int tc = (net_mask & 2) ? (c8 ? (Lc.Np >> 3) * mt8 : (Lc.Np >> 2) * mt4) : 0;
    for (int t = threadIdx.x; t < ta + tc; t += blockDim.x) {
        const bool crit = t >= ta;
        const int u = crit ? t - ta : t;
        const LayerDesc& Ld = crit ? Lc : La;
        const float* in = crit ? in_c : in_a;
        float* out = crit ? out_c : out_a;
        if (crit ? c8 : a8) {
            const int nt = u / mt8, m = u - nt * mt8;
            dense_tile_fwd8(Wbase + Ld.pw_off, Wbase + Ld.pb_off, Ld.K, Ld.Np, in, out, ld, nt, m, Ld.Np >> 1, M4 >> 1, act);
        } else {
            const int nt = u / mt4, m = u - nt * mt4;
            dense_tile_fwd(Wbase + Ld.pw_off, Wbase + Ld.pb_off, Ld.K, Ld.Np, in, out, ld, nt << 2, m << 2, act);
        }
    }
}

// One dense layer of one or both nets over a tile of M4 (multiple of 4) samples.
//   net_mask: bit0 actor, bit1 critic.  in_a/in_c, out_a/out_c: feature-major activations.
// Caller synchronises afterwards.
__device__ __forceinline__ void dense_layer(const PolicyDesc& pd, const float* __restrict__ Wbase, int layer,
                                            const float* in_a, const float* in_c, float* out_a, float* out_c,
                                            int M4, int ld, int net_mask) {
    const int mt = M4 >> 2;
    const LayerDesc& La = pd.L[0][layer];
    const LayerDesc& Lc = pd.L[1][layer];
    const int ta = (net_mask & 1) ? (La.Np >> 2) * mt : 0;
    const int tc = (net_mask & 2) ? (Lc.Np >> 2) * mt : 0;
    const bool act = layer < pd.n_layers - 1;
    for (int t = threadIdx.x; t < ta + tc; t += blockDim.x) {
        if (t < ta) {
            int nt = t / mt, m = t - nt * mt;
            dense_tile_fwd(Wbase + La.pw_off, Wbase + La.pb_off, La.K, La.Np, in_a, out_a, ld, nt << 2, m << 2, act);
        } else {
            int u = t - ta;
            int nt = u / mt, m = u - nt * mt;
            dense_tile_fwd(Wbase + Lc.pw_off, Wbase + Lc.pb_off, Lc.K, Lc.Np, in_c, out_c, ld, nt << 2, m << 2, act);
        }
    }
}

// Whole forward of the selected nets; ping-pong buffers act[net][2][max_np*ld]; returns the
// parity index of the buffer holding the final layer output. X is the shared input.
// Contains a __syncthreads after every layer (so outputs are visible on return).
// use_mma (tile width multiple of 16): layers whose padded dims are multiples of 16 run on the warp-level tensor-core
// tiles of mma_tiles.cuh (3xTF32), the others on the FMA tiles.
__device__ __forceinline__ int mlp_forward_pingpong(const PolicyDesc& pd, const float* __restrict__ Wbase,
                                                    const float* X, float* act_a, float* act_c, int M4, int ld,
                                                    int net_mask, bool use_mma = false) {
    const size_t bufsz = (size_t)pd.max_np * ld;
    for (int l = 0; l < pd.n_layers; ++l) {
        const float* ia = l == 0 ? X : act_a + ((l - 1) & 1) * bufsz;
        const float* ic = l == 0 ? X : act_c + ((l - 1) & 1) * bufsz;
        int fmask = net_mask;
        if (use_mma) {
            for (int net = 0; net < 2; ++net) {
                const LayerDesc& Ld = pd.L[net][l];
                if ((net_mask >> net & 1) && mma_layer_ok(Ld.Kp, Ld.Np)) {
                    mma_rows_layer<0>(Wbase + Ld.pw_off, Ld.Np, Ld.Np, Ld.Kp, net ? ic : ia, (net ? act_c : act_a) + (l & 1) * bufsz, ld,
                                      Wbase + Ld.pb_off, l < pd.n_layers - 1, M4 >> 3);
                    fmask &= ~(1 << net);
                }
            }
        }
        if (fmask) dense_layer(pd, Wbase, l, ia, ic, act_a + (l & 1) * bufsz, act_c + (l & 1) * bufsz, M4, ld, fmask);
        __syncthreads();
    }
    return (pd.n_layers - 1) & 1;
}

// ---------------------------------------------------------------------------------------
// Distribution heads (thread per sample). z points at the actor output column of this sample:
// z[j*ld] is logit/mean j.
// ---------------------------------------------------------------------------------------
struct HeadOut {
    float logp;
    float entropy;
    int action_idx;  // discrete: env-space action value
};

// Categorical (DRiLDistributions/categorical.jl:20-52, layers/layer_forward.jl:141-149).
// mode: 0 sample with u, 1 deterministic (argmax), 2 forced (action_value given).
__device__ __forceinline__ HeadOut categorical_head(const float* z, int ld, int A, int start, int mode, double u,
                                                    int forced_value, bool want_entropy) {
    float m = z[0];
    for (int j = 1; j < A; ++j) m = fmaxf(m, z[(size_t)j * ld]);
    float s = 0.f;
    for (int j = 0; j < A; ++j) s += expf(z[(size_t)j * ld] - m);
    int idx = A - 1;
    if (mode == 2) {
        idx = forced_value - start;
        idx = idx < 0 ? 0 : (idx >= A ? A - 1 : idx);
    } else if (mode == 1) {
        float best = -1.f;
        for (int j = 0; j < A; ++j) {
            float p = expf(z[(size_t)j * ld] - m) / s;
            if (p > best) { best = p; idx = j; }
        }
    } else {
        float cum = 0.f;
        bool found = false;
        for (int j = 0; j < A; ++j) {
            cum += expf(z[(size_t)j * ld] - m) / s;       // fp32 cumsum vs Float64 u (categorical.jl:45-47)
            if (!found && (double)cum >= u) { idx = j; found = true; }
        }
    }
    HeadOut o;
    o.action_idx = idx + start;
    o.logp = logf(expf(z[(size_t)idx * ld] - m) / s);      // log(p[a]) — log of softmax, not log-softmax
    o.entropy = 0.f;
    if (want_entropy) {
        float h = 0.f;
        for (int j = 0; j < A; ++j) {
            float p = expf(z[(size_t)j * ld] - m) / s;
            h += p * logf(p);
        }
        o.entropy = -h;
    }
    return o;
}

// standard normal draw j of the SAMPLE stream (oracle/philox.py normals())
__device__ __forceinline__ float sample_normal(uint32_t gid, uint32_t step, int j, unsigned long long seed) {
    uint32_t x[4];
    philox4x32(gid, step, (uint32_t)(j >> 2), DRIL_TAG_SAMPLE, seed, x);
    int pair = (j >> 1) & 1;
    float u1 = u01_f32_open(x[2 * pair]);
    float u2 = u01_f32(x[2 * pair + 1]);
    float r = sqrtf(-2.0f * logf(u1));
    float th = 6.2831853071795864f * u2;
    return (j & 1) ? r * sinf(th) : r * cosf(th);
}

#define DRIL_LOG2PI 1.8378770664093453f
