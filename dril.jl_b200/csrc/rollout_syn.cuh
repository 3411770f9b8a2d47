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
// Eligibility (host: syn_rollout_eligible): synthetic env, no NormalizeWrapperEnv, discrete head with <= 4 actions, one or
// two hidden layers of width <= 16, obs_dim <= 256.
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
    if (pd.act_kind != DRIL_ACT_DISCRETE || pd.act_n > SYN_MAX_A) return false;
    if (pd.n_layers < 2 || pd.n_layers > 3 || d.obs_dim > 256) return false;
    for (int net = 0; net < 2; ++net)
        for (int l = 0; l + 1 < pd.n_layers; ++l)
            if (pd.L[net][l].Np > 16) return false;
    return true;
}

// hidden layers of the nets in MASK (1 actor, 2 critic) on the observation of (gid, life); the observation row is stored to
// `obs_row` when non-null.  h[net][*] returns the last hidden activation.
template <int HP, int NH, int MASK>
__device__ __forceinline__ void syn_hidden(const float* __restrict__ sW, const SynSmem& L, int D, int Dp, uint32_t gid, uint32_t life,
                                           unsigned long long seed, float* __restrict__ obs_row, float (&h)[2][HP]) {
#pragma unroll
    for (int net = 0; net < 2; ++net)
        if (MASK & (1 << net)) {
#pragma unroll
            for (int n = 0; n < HP; ++n) h[net][n] = sW[L.b0 + net * HP + n];
        }
    const int nb = Dp >> 2;
    for (int b = 0; b < nb; ++b) {
        float o[4];
        synthetic_obs_block(gid, life, b, seed, o);
        if (obs_row) {
            if ((D & 3) == 0) *reinterpret_cast<float4*>(obs_row + 4 * b) = make_float4(o[0], o[1], o[2], o[3]);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (4 * b + j < D) obs_row[4 * b + j] = o[j];
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
                        h[net][4 * q + 0] = fmaf(o[j], v.x, h[net][4 * q + 0]);
                        h[net][4 * q + 1] = fmaf(o[j], v.y, h[net][4 * q + 1]);
                        h[net][4 * q + 2] = fmaf(o[j], v.z, h[net][4 * q + 2]);
                        h[net][4 * q + 3] = fmaf(o[j], v.w, h[net][4 * q + 3]);
                    }
                }
        }
    }
#pragma unroll
    for (int net = 0; net < 2; ++net)
        if (MASK & (1 << net)) {
#pragma unroll
            for (int n = 0; n < HP; ++n) h[net][n] = fast_tanh(h[net][n]);
        }
    if (NH == 2) {
#pragma unroll
        for (int net = 0; net < 2; ++net)
            if (MASK & (1 << net)) {
                float g[HP];
#pragma unroll
                for (int n = 0; n < HP; ++n) g[n] = sW[L.b1 + net * HP + n];
#pragma unroll
                for (int k = 0; k < HP; ++k) {
                    const float4* w = reinterpret_cast<const float4*>(sW + L.w1 + ((size_t)net * HP + k) * HP);
#pragma unroll
                    for (int q = 0; q < HP / 4; ++q) {
                        const float4 v = w[q];
                        g[4 * q + 0] = fmaf(h[net][k], v.x, g[4 * q + 0]);
                        g[4 * q + 1] = fmaf(h[net][k], v.y, g[4 * q + 1]);
                        g[4 * q + 2] = fmaf(h[net][k], v.z, g[4 * q + 2]);
                        g[4 * q + 3] = fmaf(h[net][k], v.w, g[4 * q + 3]);
                    }
                }
#pragma unroll
                for (int n = 0; n < HP; ++n) h[net][n] = fast_tanh(g[n]);
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
    float h[2][HP];
    syn_hidden<HP, NH, 2>(sW, L, D, Dp, gid, life, seed, nullptr, h);
    return syn_value<HP>(sW, L, h[1]);
}

template <int HP, int NH>
__global__ void __launch_bounds__(128) rollout_syn_kernel(const __grid_constant__ RolloutArgs a) {
    extern __shared__ float4 smem4[];
    float* sW = reinterpret_cast<float*>(smem4);
    __shared__ double s_red[3][4];
    const EnvDev& env = a.env;
    const BufDev& buf = a.buf;
    const PolicyDesc& pd = a.pd;
    const int D = env.obs_dim, Dp = (D + 3) & ~3, A = pd.act_n;
    const long long N = env.n_envs;
    const SynSmem L = syn_smem_layout(Dp, HP, NH);
    const int tid = threadIdx.x;
    // ---- weights: packed [Kp][Np] (zero padded) -> own layout padded to HP (padding columns zero: tanh(0) = 0 feeds nothing) ----
    for (int i = tid; i < L.total; i += blockDim.x) sW[i] = 0.f;
    __syncthreads();
    for (int net = 0; net < 2; ++net) {
        const LayerDesc& l0 = pd.L[net][0];
        for (int i = tid; i < l0.K * l0.N; i += blockDim.x) {
            const int k = i / l0.N, n = i - k * l0.N;
            sW[L.w0 + ((size_t)net * Dp + k) * HP + n] = a.pack[l0.pw_off + k * l0.Np + n];
        }
        for (int n = tid; n < l0.N; n += blockDim.x) sW[L.b0 + net * HP + n] = a.pack[l0.pb_off + n];
        if (NH == 2) {
            const LayerDesc& l1 = pd.L[net][1];
            for (int i = tid; i < l1.K * l1.N; i += blockDim.x) {
                const int k = i / l1.N, n = i - k * l1.N;
                sW[L.w1 + ((size_t)net * HP + k) * HP + n] = a.pack[l1.pw_off + k * l1.Np + n];
            }
            for (int n = tid; n < l1.N; n += blockDim.x) sW[L.b1 + net * HP + n] = a.pack[l1.pb_off + n];
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

    const long long n = (long long)blockIdx.x * blockDim.x + tid;
    const bool mine = n < N;
    const uint32_t gid = (uint32_t)(env.gid_offset + n);
    uint32_t life = 0, episode = 0;
    int steps = 0, ep_len = 0;
    float ep_ret = 0.f;
    double sum_r = 0.0, sum_l = 0.0, sum_e = 0.0;
    if (mine) {
        life = env.life[n]; steps = env.steps[n]; episode = env.episode[n];
        if (env.monitor) { ep_ret = env.ep_ret[n]; ep_len = env.ep_len[n]; }
    }
    const bool deterministic = (a.flags & RO_DETERMINISTIC) != 0;
    for (int t = 0; t < a.T; ++t) {
        const size_t row = (size_t)t * N;
        int done_i = 0;
        if (mine) {
            float h[2][HP];
            syn_hidden<HP, NH, 3>(sW, L, D, Dp, gid, life, env.seed, buf.obs + (row + n) * D, h);
            const float value = syn_value<HP>(sW, L, h[1]);
            float z[SYN_MAX_A];
            {
                const float4 bo = *reinterpret_cast<const float4*>(sW + L.bo);
                z[0] = bo.x; z[1] = bo.y; z[2] = bo.z; z[3] = bo.w;
#pragma unroll
                for (int k = 0; k < HP; ++k) {
                    const float4 w = *reinterpret_cast<const float4*>(sW + L.wo + k * SYN_MAX_A);
                    z[0] = fmaf(h[0][k], w.x, z[0]); z[1] = fmaf(h[0][k], w.y, z[1]);
                    z[2] = fmaf(h[0][k], w.z, z[2]); z[3] = fmaf(h[0][k], w.w, z[3]);
                }
            }
            // categorical head in registers: the operations of categorical_head (mlp.cuh; categorical.jl:20-52)
            float m = z[0];
#pragma unroll
            for (int j = 1; j < SYN_MAX_A; ++j) if (j < A) m = fmaxf(m, z[j]);
            float ex[SYN_MAX_A], s = 0.f;
#pragma unroll
            for (int j = 0; j < SYN_MAX_A; ++j) { ex[j] = j < A ? expf(z[j] - m) : 0.f; if (j < A) s += ex[j]; }
            int idx = A - 1;
            if (a.forced) {
                idx = reinterpret_cast<const int*>(a.forced)[row + n] - pd.act_start;
                idx = idx < 0 ? 0 : (idx >= A ? A - 1 : idx);
            } else if (deterministic) {
                float best = -1.f;
#pragma unroll
                for (int j = 0; j < SYN_MAX_A; ++j) if (j < A) { const float p = ex[j] / s; if (p > best) { best = p; idx = j; } }
            } else {
                uint32_t x[4];
                philox4x32(gid, a.step0 + (uint32_t)t, 0u, DRIL_TAG_SAMPLE, a.pseed, x);
                const double u = u01_f64(x[0], x[1]);
                float cum = 0.f;
                bool found = false;
#pragma unroll
                for (int j = 0; j < SYN_MAX_A; ++j)
                    if (j < A) {
                        cum += ex[j] / s;                            // fp32 cumsum vs Float64 u (categorical.jl:45-47)
                        if (!found && (double)cum >= u) { idx = j; found = true; }
                    }
            }
            float pe = ex[0];
#pragma unroll
            for (int j = 1; j < SYN_MAX_A; ++j) if (j == idx) pe = ex[j];
            reinterpret_cast<int*>(buf.actions)[row + n] = idx + pd.act_start;
            buf.values[row + n] = value;
            buf.logprobs[row + n] = logf(pe / s);
            // env step (the action is ignored by this env), Monitor, auto-reset
            bool term = false;
            const float r = synthetic_step(gid, life, env.seed, &term);
            life += 1;
            steps += 1;
            const bool trunc = steps >= env.max_steps;
            const bool done = term || trunc;
            buf.flags[row + n] = (unsigned char)((term ? 1 : 0) | (trunc ? 2 : 0));
            buf.rewards[row + n] = r;
            if (env.monitor) {
                ep_ret = __fadd_rn(ep_ret, r);
                ep_len += 1;
                if (done) {
                    buf.episode_r[row + n] = ep_ret;
                    buf.episode_l[row + n] = ep_len;
                    sum_r += (double)ep_ret; sum_l += (double)ep_len; sum_e += 1.0;
                    done_i = 1;
                    ep_ret = 0.f; ep_len = 0;
                }
            }
            // terminal_observation = observe() of the stepped env (lifetime counter already advanced), trajectory.jl:57-61
            if (trunc) buf.boot[row + n] = syn_critic_only<HP, NH>(sW, L, D, Dp, gid, life, env.seed);
            if (done) { episode += 1; steps = 0; }
        }
        if (env.monitor) {
            const unsigned int bal = __ballot_sync(0xffffffffu, done_i);
            if (bal && (tid & 31) == 0) atomicAdd(&buf.done_count[t], __popc(bal));
        }
    }
    if (mine && a.T > 0) buf.last_values[n] = syn_critic_only<HP, NH>(sW, L, D, Dp, gid, life, env.seed);   // trajectory.jl:65-70
    if (mine) {
        env.life[n] = life; env.steps[n] = steps; env.episode[n] = episode;
        if (env.monitor) { env.ep_ret[n] = ep_ret; env.ep_len[n] = ep_len; }
    }
    if (env.monitor) {
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
