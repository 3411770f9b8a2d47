// Register-resident rollout for the synthetic env with a small policy (the rollout-only sweep C5):
// ONE THREAD PER ENV for all n_steps.  Same semantics and citations as rollout_kernel (rollout.cuh): collect_trajectories
// (buffers/trajectory.jl:33-76) over MonitorWrapperEnv.act! (monitorWrapperEnv.jl:44-60) and MultiThreadedParallelEnv.act!
// (multithreadedParallelEnv.jl:47-74) with the layer call (layers/layer_forward.jl:3-39) and the bootstrap predict_values
// calls (trajectory.jl:57-70) fused in.
//
// Why a kernel of its own: with a policy this small the rollout is a stream of 4·D + 17 bytes per env-step into the buffer
// and the general kernel's shared-memory staging (raw tile -> normalised tile -> activations -> head) and per-step barriers
// cost 20x more than the arithmetic.  Here
//   * the observation is never materialised: each Philox block of 4 values is generated, stored to the buffer row as one
//     float4 and consumed by layer 0 (weights read from shared memory as warp-uniform LDS.128 broadcasts) on the spot;
//   * hidden activations of both nets (<= 2 x 16), step / lifetime / episode counters and the Monitor accumulators live in
//     registers; there is no barrier in the step loop;
//   * V(terminal_obs) of a truncated step is a rare divergent critic evaluation; Monitor totals are accumulated per thread
//     and reduced once per CTA at the end (one atomic per CTA instead of one per finished episode).
// Eligibility (host: syn_rollout_eligible): synthetic env (two actions), no NormalizeWrapperEnv, one or two hidden layers of
// width <= 16, obs_dim <= 256.  Narrow nets (<= 8) run two envs per thread: every weight read feeds both, and the two Philox /
// tanh chains interleave.
#pragma once
#include "rollout.cuh"

#define SYN_MAX_A 4

struct SynSmem {
    int w0, b0, w1, b1, wo, bo, wc, total;   // float offsets; per-net strides below
};
// layout: W0[net][Dp][HP] | b0[net][HP] | W1[net][HP][HP] | b1[net][HP] | Wo[HP][4] | bo[4] | Wc[HP] | bc
__host__ __device__ inline SynSmem syn_smem_layout(int Dp, int HP, int NH) {
    SynSmem s;
    int o = 0;
    s.w0 = o; o += 2 * Dp * HP;
    s.b0 = o; o += 2 * HP;
    s.w1 = o; o += NH == 2 ? 2 * HP * HP : 0;
    s.b1 = o; o += NH == 2 ? 2 * HP : 0;
    s.wo = o; o += HP * SYN_MAX_A;
    s.bo = o; o += SYN_MAX_A;
    s.wc = o; o += HP;
    o += 4;                                  // bc (+ padding)
    s.total = o;
    return s;
}

inline bool syn_rollout_eligible(const PolicyDesc& pd, const EnvDev& d) {
    if (d.kind != DRIL_ENV_SYNTHETIC || d.normalize) return false;
    if (pd.act_kind != DRIL_ACT_DISCRETE || pd.act_n != 2) return false;    // the synthetic env's Discrete(2)
    if (pd.n_layers < 2 || pd.n_layers > 3 || d.obs_dim > 256) return false;
    for (int net = 0; net < 2; ++net)
        for (int l = 0; l + 1 < pd.n_layers; ++l)
            if (pd.L[net][l].Np > 16) return false;
    return true;
}

// tanh of a pre-activation that already carries the factor 2·log2(e) (folded into the staged weights and biases):
// tanh(x) = 1 - 2 / (1 + 2^(x · 2 log2 e)), the same evaluation as fast_tanh (common.cuh) without its multiply
#define SYN_TANH_SCALE 2.8853900817779268f
__device__ __forceinline__ float syn_tanh_prescaled(float z) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}

// observation block b of lifetime step `life`: the values of synthetic_obs_block (env.cuh) — (float)(x >> 8) · 2^-24 is exact and
// so is its double, hence -1 + 2u rounds once, exactly like this single fma
__device__ __forceinline__ void syn_obs_block(uint32_t gid, uint32_t life, int b, unsigned long long seed, float o[4]) {
    uint32_t x[4];
    philox4x32(gid, life, (uint32_t)b, DRIL_TAG_SYN_OBS, seed, x);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = fmaf((float)(x[j] >> 8), 1.1920928955078125e-07f, -1.0f);
}

// hidden layers of the nets in MASK (1 actor, 2 critic) for E envs at once (every weight read from shared memory feeds E envs);
// the observation rows are stored to obs_row[e] when non-null.  h[e][net][*] returns the last hidden activation.
template <int HP, int NH, int MASK, int E>
__device__ __forceinline__ void syn_hidden(const float* __restrict__ sW, const SynSmem& L, int D, int Dp, const uint32_t (&gid)[E],
                                           const uint32_t (&life)[E], unsigned long long seed, float* const (&obs_row)[E],
                                           float (&h)[E][2][HP]) {
#pragma unroll
    for (int net = 0; net < 2; ++net)
        if (MASK & (1 << net)) {
#pragma unroll
            for (int n = 0; n < HP; ++n) {
                const float bv = sW[L.b0 + net * HP + n];
#pragma unroll
                for (int e = 0; e < E; ++e) h[e][net][n] = bv;
            }
        }
    const int nb = Dp >> 2;
    for (int b = 0; b < nb; ++b) {
        float o[E][4];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            syn_obs_block(gid[e], life[e], b, seed, o[e]);
            if (obs_row[e]) {
                if ((D & 3) == 0) *reinterpret_cast<float4*>(obs_row[e] + 4 * b) = make_float4(o[e][0], o[e][1], o[e][2], o[e][3]);
                else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (4 * b + j < D) obs_row[e][4 * b + j] = o[e][j];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int net = 0; net < 2; ++net)
                if (MASK & (1 << net)) {
                    const float4* w = reinterpret_cast<const float4*>(sW + L.w0 + ((size_t)net * Dp + 4 * b + j) * HP);
#pragma unroll
                    for (int q = 0; q < HP / 4; ++q) {
                        const float4 v = w[q];
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            h[e][net][4 * q + 0] = fmaf(o[e][j], v.x, h[e][net][4 * q + 0]);
                            h[e][net][4 * q + 1] = fmaf(o[e][j], v.y, h[e][net][4 * q + 1]);
                            h[e][net][4 * q + 2] = fmaf(o[e][j], v.z, h[e][net][4 * q + 2]);
                            h[e][net][4 * q + 3] = fmaf(o[e][j], v.w, h[e][net][4 * q + 3]);
                        }
                    }
                }
        }
    }
#pragma unroll
    for (int e = 0; e < E; ++e)
#pragma unroll
        for (int net = 0; net < 2; ++net)
            if (MASK & (1 << net)) {
#pragma unroll
                for (int n = 0; n < HP; ++n) h[e][net][n] = syn_tanh_prescaled(h[e][net][n]);
            }
    if (NH == 2) {
#pragma unroll
        for (int net = 0; net < 2; ++net)
            if (MASK & (1 << net)) {
                float g[E][HP];
#pragma unroll
                for (int n = 0; n < HP; ++n) {
                    const float bv = sW[L.b1 + net * HP + n];
#pragma unroll
                    for (int e = 0; e < E; ++e) g[e][n] = bv;
                }
#pragma unroll
                for (int k = 0; k < HP; ++k) {
                    const float4* w = reinterpret_cast<const float4*>(sW + L.w1 + ((size_t)net * HP + k) * HP);
#pragma unroll
                    for (int q = 0; q < HP / 4; ++q) {
                        const float4 v = w[q];
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            g[e][4 * q + 0] = fmaf(h[e][net][k], v.x, g[e][4 * q + 0]);
                            g[e][4 * q + 1] = fmaf(h[e][net][k], v.y, g[e][4 * q + 1]);
                            g[e][4 * q + 2] = fmaf(h[e][net][k], v.z, g[e][4 * q + 2]);
                            g[e][4 * q + 3] = fmaf(h[e][net][k], v.w, g[e][4 * q + 3]);
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < E; ++e)
#pragma unroll
                    for (int n = 0; n < HP; ++n) h[e][net][n] = syn_tanh_prescaled(g[e][n]);
            }
    }
}

template <int HP>
__device__ __forceinline__ float syn_value(const float* __restrict__ sW, const SynSmem& L, const float (&hc)[HP]) {
    float v = sW[L.wc + HP];
#pragma unroll
    for (int k = 0; k < HP; ++k) v = fmaf(hc[k], sW[L.wc + k], v);
    return v;
}

// critic value of the observation of (gid, life): V(terminal_obs) / V(new_obs); rare or once per rollout -> not inlined
template <int HP, int NH>
__device__ __noinline__ float syn_critic_only(const float* __restrict__ sW, SynSmem L, int D, int Dp, uint32_t gid, uint32_t life,
                                              unsigned long long seed) {
    float h[1][2][HP];
    const uint32_t g1[1] = {gid}, l1[1] = {life};
    float* const none[1] = {nullptr};
    syn_hidden<HP, NH, 2, 1>(sW, L, D, Dp, g1, l1, seed, none, h);
    return syn_value<HP>(sW, L, h[0][1]);
}

// E envs per thread: env e of thread tid in CTA c is c·E·blockDim + e·blockDim + tid (per-step rows stay coalesced per e)
template <int HP, int NH, int E>
__global__ void __launch_bounds__(128) rollout_syn_kernel(const __grid_constant__ RolloutArgs a) {
    extern __shared__ float4 smem4[];
    float* sW = reinterpret_cast<float*>(smem4);
    __shared__ double s_red[3][4];
    const EnvDev& env = a.env;
    const BufDev& buf = a.buf;
    const PolicyDesc& pd = a.pd;
    const int D = env.obs_dim, Dp = (D + 3) & ~3;
    constexpr int A = 2;                                       // the synthetic env's Discrete(2)
    const long long N = env.n_envs;
    const SynSmem L = syn_smem_layout(Dp, HP, NH);
    const int tid = threadIdx.x;
    // ---- weights: packed [Kp][Np] (zero padded) -> own layout padded to HP (padding columns zero: tanh(0) = 0 feeds nothing);
    //      hidden layers carry tanh's 2·log2(e) ------------------------------------------------------------------------------
    for (int i = tid; i < L.total; i += blockDim.x) sW[i] = 0.f;
    __syncthreads();
    for (int net = 0; net < 2; ++net) {
        const LayerDesc& l0 = pd.L[net][0];
        for (int i = tid; i < l0.K * l0.N; i += blockDim.x) {
            const int k = i / l0.N, n = i - k * l0.N;
            sW[L.w0 + ((size_t)net * Dp + k) * HP + n] = a.pack[l0.pw_off + k * l0.Np + n] * SYN_TANH_SCALE;
        }
        for (int n = tid; n < l0.N; n += blockDim.x) sW[L.b0 + net * HP + n] = a.pack[l0.pb_off + n] * SYN_TANH_SCALE;
        if (NH == 2) {
            const LayerDesc& l1 = pd.L[net][1];
            for (int i = tid; i < l1.K * l1.N; i += blockDim.x) {
                const int k = i / l1.N, n = i - k * l1.N;
                sW[L.w1 + ((size_t)net * HP + k) * HP + n] = a.pack[l1.pw_off + k * l1.Np + n] * SYN_TANH_SCALE;
            }
            for (int n = tid; n < l1.N; n += blockDim.x) sW[L.b1 + net * HP + n] = a.pack[l1.pb_off + n] * SYN_TANH_SCALE;
        }
    }
    {
        const LayerDesc& lo = pd.L[0][NH];
        for (int i = tid; i < lo.K * A; i += blockDim.x) {
            const int k = i / A, j = i - k * A;
            sW[L.wo + k * SYN_MAX_A + j] = a.pack[lo.pw_off + k * lo.Np + j];
        }
        for (int j = tid; j < A; j += blockDim.x) sW[L.bo + j] = a.pack[lo.pb_off + j];
        const LayerDesc& lc = pd.L[1][NH];
        for (int k = tid; k < lc.K; k += blockDim.x) sW[L.wc + k] = a.pack[lc.pw_off + k * lc.Np];
        if (tid == 0) sW[L.wc + HP] = a.pack[lc.pb_off];
    }
    __syncthreads();

    long long n[E];
    bool mine[E];
    uint32_t gid[E], life[E], episode[E];
    int steps[E], ep_len[E];
    float ep_ret[E];
    double sum_r = 0.0, sum_l = 0.0, sum_e = 0.0;
    const int monitor = env.monitor, max_steps = env.max_steps, act_start = pd.act_start;
    const unsigned long long eseed = env.seed, pseed = a.pseed;
    const int* __restrict__ forced = reinterpret_cast<const int*>(a.forced);
    const bool deterministic = (a.flags & RO_DETERMINISTIC) != 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        n[e] = ((long long)blockIdx.x * E + e) * blockDim.x + tid;
        mine[e] = n[e] < N;
        gid[e] = (uint32_t)(env.gid_offset + n[e]);
        life[e] = 0; episode[e] = 0; steps[e] = 0; ep_len[e] = 0; ep_ret[e] = 0.f;
        if (mine[e]) {
            life[e] = env.life[n[e]]; steps[e] = env.steps[n[e]]; episode[e] = env.episode[n[e]];
            if (monitor) { ep_ret[e] = env.ep_ret[n[e]]; ep_len[e] = env.ep_len[n[e]]; }
        }
    }
    for (int t = 0; t < a.T; ++t) {
        const size_t row = (size_t)t * N;
        float h[E][2][HP];
        float* orow[E];
#pragma unroll
        for (int e = 0; e < E; ++e) orow[e] = mine[e] ? buf.obs + (row + n[e]) * D : nullptr;
        // (threads past the end compute on gid garbage and store nothing)
        syn_hidden<HP, NH, 3, E>(sW, L, D, Dp, gid, life, eseed, orow, h);
        int done_any = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const float value = syn_value<HP>(sW, L, h[e][1]);
            float z0, z1;
            {
                const float2 bo = *reinterpret_cast<const float2*>(sW + L.bo);
                z0 = bo.x; z1 = bo.y;
#pragma unroll
                for (int k = 0; k < HP; ++k) {
                    const float2 w = *reinterpret_cast<const float2*>(sW + L.wo + k * SYN_MAX_A);
                    z0 = fmaf(h[e][0][k], w.x, z0); z1 = fmaf(h[e][0][k], w.y, z1);
                }
            }
            // categorical head in registers: the operations of categorical_head (mlp.cuh; categorical.jl:20-52) for two logits
            const float m = fmaxf(z0, z1);
            const float e0 = expf(z0 - m), e1 = expf(z1 - m);
            const float s = e0 + e1;
            int idx = A - 1;
            if (forced) {
                idx = mine[e] ? forced[row + n[e]] - act_start : 0;
                idx = idx < 0 ? 0 : (idx >= A ? A - 1 : idx);
            } else if (deterministic) {
                idx = (e1 / s > e0 / s) ? 1 : 0;
            } else {
                uint32_t x[4];
                philox4x32(gid[e], a.step0 + (uint32_t)t, 0u, DRIL_TAG_SAMPLE, pseed, x);
                const double u = u01_f64(x[0], x[1]);
                const float c0 = 0.f + e0 / s;                       // fp32 cumsum vs Float64 u (categorical.jl:45-47)
                const float c1 = c0 + e1 / s;
                idx = (double)c0 >= u ? 0 : ((double)c1 >= u ? 1 : A - 1);
            }
            const float logp = logf((idx == 0 ? e0 : e1) / s);
            // env step (the action is ignored by this env), Monitor, auto-reset
            bool term = false;
            const float r = synthetic_step(gid[e], life[e], eseed, &term);
            life[e] += 1;
            steps[e] += 1;
            const bool trunc = steps[e] >= max_steps;
            const bool done = term || trunc;
            if (mine[e]) {
                const size_t sidx = row + n[e];
                reinterpret_cast<int*>(buf.actions)[sidx] = idx + act_start;
                buf.values[sidx] = value;
                buf.logprobs[sidx] = logp;
                buf.flags[sidx] = (unsigned char)((term ? 1 : 0) | (trunc ? 2 : 0));
                buf.rewards[sidx] = r;
                if (monitor) {
                    ep_ret[e] = __fadd_rn(ep_ret[e], r);
                    ep_len[e] += 1;
                    if (done) {
                        buf.episode_r[sidx] = ep_ret[e];
                        buf.episode_l[sidx] = ep_len[e];
                        sum_r += (double)ep_ret[e]; sum_l += (double)ep_len[e]; sum_e += 1.0;
                        done_any += 1;
                        ep_ret[e] = 0.f; ep_len[e] = 0;
                    }
                }
                // terminal_observation = observe() of the stepped env (lifetime counter already advanced), trajectory.jl:57-61
                if (trunc) buf.boot[sidx] = syn_critic_only<HP, NH>(sW, L, D, Dp, gid[e], life[e], eseed);
            }
            if (done) { episode[e] += 1; steps[e] = 0; }
        }
        if (monitor) {
            int c = done_any;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (c && (tid & 31) == 0) atomicAdd(&buf.done_count[t], c);
        }
    }
#pragma unroll
    for (int e = 0; e < E; ++e)
        if (mine[e]) {
            if (a.T > 0) buf.last_values[n[e]] = syn_critic_only<HP, NH>(sW, L, D, Dp, gid[e], life[e], eseed);   // trajectory.jl:65-70
            env.life[n[e]] = life[e]; env.steps[n[e]] = steps[e]; env.episode[n[e]] = episode[e];
            if (monitor) { env.ep_ret[n[e]] = ep_ret[e]; env.ep_len[n[e]] = ep_len[e]; }
        }
    if (monitor) {
        sum_r = warp_sum(sum_r); sum_l = warp_sum(sum_l); sum_e = warp_sum(sum_e);
        if ((tid & 31) == 0) { s_red[0][tid >> 5] = sum_r; s_red[1][tid >> 5] = sum_l; s_red[2][tid >> 5] = sum_e; }
        __syncthreads();
        if (tid == 0) {
            double r = 0.0, l = 0.0, e = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { r += s_red[0][w]; l += s_red[1][w]; e += s_red[2][w]; }
            if (e > 0.0) {
                atomicAdd(&env.roll_sums[0], r);
                atomicAdd(&env.roll_sums[1], l);
                atomicAdd(env.roll_eps, (unsigned long long)(e + 0.5));
            }
        }
    }
}
