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
__global__ void __launch_bounds__(256) adv_stats_kernel(const float* __restrict__ adv, long long n_total,
                                                        long long batch_size, FeistelKey fk, int identity,
                                                        double* __restrict__ partial) {
    __shared__ double scratch[32];
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
struct LossSmem {
    int ld;
    size_t w, x, h[2][DRIL_MAX_LAYERS], g[2], samp, total_floats, dbl_bytes_off, total;
};

__host__ __device__ inline LossSmem loss_smem_layout(const PolicyDesc& pd, int M4, bool weights_smem) {
    LossSmem s;
    s.ld = M4 + 4;
    size_t o = 0;
    s.w = o; o += weights_smem ? (size_t)pd.pack_total : 0;
    s.x = o; o += (size_t)pd.obs_dim_p * s.ld;
    for (int net = 0; net < 2; ++net)
        for (int l = 0; l < pd.n_layers; ++l) { s.h[net][l] = o; o += (size_t)pd.L[net][l].Np * s.ld; }
    for (int net = 0; net < 2; ++net) { s.g[net] = o; o += (size_t)2 * pd.max_np * s.ld; }
    int arows = pd.act_kind == DRIL_ACT_CONTINUOUS ? 2 * pd.act_n : 1;   // actions (+ log_std grad contributions)
    s.samp = o; o += (size_t)(4 + arows) * s.ld;   // adv, ret, old_logp, old_val, actions...
    o = (o + 3) & ~(size_t)3;
    s.total_floats = o;
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
    float* gpart;          // [gridDim.x][pd.gpack] per-CTA packed gradient partials (+ log_std + stats)
    const int* stop_flag;  // target_kl stop already fired: do nothing
    Minibatch mb;
    UpdateHyper hp;
    int M4, weights_smem;
};

// dW[k][n] (+)= sum_m Ain[k][m] * dZ[n][m] for the 4 interleaved rows k = kt + kstride*i and the
// 4 columns n0..n0+3; accumulated into this CTA's packed partial.
__device__ __forceinline__ void dense_tile_dw(const float* __restrict__ Ain, const float* __restrict__ dZ, int M4, int ld,
                                              int kt, int kstride, int n0, float* __restrict__ gW, int Np, bool first) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* hp0 = Ain + (size_t)kt * ld;
    const float* dp0 = dZ + (size_t)n0 * ld;
#pragma unroll 2
    for (int m = 0; m < M4; m += 4) {
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
        if (!first) { float4 o = *g; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
        *g = v;
    }
}

// dZprev[k0..k0+3][m0..m0+3] = (sum_n Wt[n][k0..] * dZ[n][m0..]) * (1 - H[k][m]^2)
__device__ __forceinline__ void dense_tile_dh(const float* __restrict__ Wt, int Nred, int Kp, const float* __restrict__ dZ,
                                              const float* __restrict__ H, float* __restrict__ dZprev, int ld, int k0, int m0) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* wp = Wt + k0;
    const float* dp = dZ + m0;
#pragma unroll 4
    for (int n = 0; n < Nred; ++n) {
        float4 w = *reinterpret_cast<const float4*>(wp + (size_t)n * Kp);
        float4 d = *reinterpret_cast<const float4*>(dp + (size_t)n * ld);
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
        float4 h = *reinterpret_cast<const float4*>(H + (size_t)(k0 + i) * ld + m0);
        float4 o;
        o.x = acc[i][0] * (1.0f - h.x * h.x); o.y = acc[i][1] * (1.0f - h.y * h.y);
        o.z = acc[i][2] * (1.0f - h.z * h.z); o.w = acc[i][3] * (1.0f - h.w * h.w);
        *reinterpret_cast<float4*>(dZprev + (size_t)(k0 + i) * ld + m0) = o;
    }
}

__global__ void __launch_bounds__(DRIL_THREADS) ppo_loss_grad_kernel(const __grid_constant__ LossArgs a) {
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);
    const PolicyDesc& pd = a.pd;
    const BufDev& buf = a.buf;
    const int M4 = a.M4, D = pd.obs_dim, Dp = pd.obs_dim_p, NL = pd.n_layers;
    const LossSmem S = loss_smem_layout(pd, M4, a.weights_smem);
    const int ld = S.ld;
    const int tid = threadIdx.x;
    float* sX = smem + S.x;
    float* sAdv = smem + S.samp;
    float* sRet = sAdv + ld;
    float* sOldLp = sRet + ld;
    float* sOldV = sOldLp + ld;
    float* sActn = sOldV + ld;                       // discrete: 1 row (int bits); continuous: act_n rows + act_n rows of log_std contributions
    double* sDbl = reinterpret_cast<double*>(reinterpret_cast<char*>(smem) + S.dbl_bytes_off);
    float* gp = a.gpart + (size_t)blockIdx.x * pd.gpack;

    if (*a.stop_flag) return;

    const float* Wbase = a.pack;
    if (a.weights_smem) {
        const float4* src = reinterpret_cast<const float4*>(a.pack);
        float4* dst = reinterpret_cast<float4*>(smem + S.w);
        for (int i = tid; i < pd.pack_total / 4; i += blockDim.x) dst[i] = src[i];
        Wbase = smem + S.w;
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
    bool first = true;
    __syncthreads();

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, first = false) {
        const long long p0 = a.mb.start + tile * M4;
        const int nvalid = (int)min((long long)M4, a.mb.start + a.mb.count - p0);
        // ---- gather the tile through the Feistel bijection --------------------------------
        long long sidx = -1;
        if (tid < M4) {
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
            dense_layer(pd, Wbase, l, ia, ic, smem + S.h[0][l], smem + S.h[1][l], M4, ld, 3);
            __syncthreads();
        }
        // ---- loss head: dL/dlogits (or dL/dmean), dL/dvalue ---------------------------------
        float* gA = smem + S.g[0] + (size_t)((NL - 1) & 1) * pd.max_np * ld;
        float* gC = smem + S.g[1] + (size_t)((NL - 1) & 1) * pd.max_np * ld;
        if (tid < M4) {
            const bool valid = tid < nvalid;
            const float* z = smem + S.h[0][NL - 1] + tid;
            const float adv = sAdv[tid];
            float logp, ent;
            const int A = pd.act_n;
            const int Ap = pd.L[0][NL - 1].Np;
            float g_logp = 0.f;
            const float g_ent = valid ? -a.hp.ent_coef * invB : 0.f;
            float ratio = 1.f, s1 = 0.f, s2 = 0.f, log_ratio = 0.f, rc = 1.f;
            if (pd.act_kind == DRIL_ACT_DISCRETE) {
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
            if (valid) {
                st_p += (double)(-fminf(s1, s2));
                st_v += (double)(verr * verr);
                st_e += (double)ent;
                st_clip += (ratio != rc) ? 1.0 : 0.0;
                st_kl += (double)(expf(log_ratio) - 1.0f - log_ratio);
                st_ratio += (double)ratio;
            }
        }
        __syncthreads();
        if (pd.act_kind == DRIL_ACT_CONTINUOUS && tid < pd.act_n) {
            double s = 0;
            for (int e = 0; e < nvalid; ++e) s += (double)sActn[(size_t)(pd.act_n + tid) * ld + e];
            ls_acc += s;
        }
        // ---- backward -----------------------------------------------------------------------
        for (int l = NL - 1; l >= 0; --l) {
            const int mt = M4 >> 2;
            int counts[5];   // dW actor, dW critic, dH actor, dH critic, db (both nets)
            const LayerDesc& La = pd.L[0][l];
            const LayerDesc& Lc = pd.L[1][l];
            counts[0] = (La.Kp >> 2) * (La.Np >> 2);
            counts[1] = (Lc.Kp >> 2) * (Lc.Np >> 2);
            counts[2] = l > 0 ? (La.Kp >> 2) * mt : 0;
            counts[3] = l > 0 ? (Lc.Kp >> 2) * mt : 0;
            counts[4] = La.N + Lc.N;
            const int total = counts[0] + counts[1] + counts[2] + counts[3] + counts[4];
            for (int t = tid; t < total; t += blockDim.x) {
                int u = t;
                if (u < counts[0] + counts[1]) {
                    const int net = u >= counts[0];
                    if (net) u -= counts[0];
                    const LayerDesc& Ld = net ? Lc : La;
                    const int kq = Ld.Kp >> 2;
                    const int nt = u / kq, kt = u - nt * kq;
                    const float* Ain = l == 0 ? sX : smem + S.h[net][l - 1];
                    const float* dZ = smem + S.g[net] + (size_t)(l & 1) * pd.max_np * ld;
                    dense_tile_dw(Ain, dZ, M4, ld, kt, kq, nt << 2, gp + Ld.pw_off, Ld.Np, first);
                    continue;
                }
                u -= counts[0] + counts[1];
                if (u < counts[2] + counts[3]) {
                    const int net = u >= counts[2];
                    if (net) u -= counts[2];
                    const LayerDesc& Ld = net ? Lc : La;
                    const int kq_t = u / mt, m = u - kq_t * mt;
                    const float* dZ = smem + S.g[net] + (size_t)(l & 1) * pd.max_np * ld;
                    float* dZprev = smem + S.g[net] + (size_t)((l - 1) & 1) * pd.max_np * ld;
                    dense_tile_dh(Wbase + Ld.pwt_off, Ld.N, Ld.Kp, dZ, smem + S.h[net][l - 1], dZprev, ld, kq_t << 2, m << 2);
                    continue;
                }
                u -= counts[2] + counts[3];
                {
                    const int net = u >= La.N;
                    if (net) u -= La.N;
                    const LayerDesc& Ld = net ? Lc : La;
                    const float* dZ = smem + S.g[net] + (size_t)(l & 1) * pd.max_np * ld + (size_t)u * ld;
                    float s = 0.f;
                    for (int m = 0; m < M4; ++m) s += dZ[m];
                    float* g = gp + Ld.pb_off + u;
                    *g = first ? s : *g + s;
                }
            }
            __syncthreads();
        }
    }
    // ---- per-CTA tail: log_std gradient + statistic sums --------------------------------------
    if (n_tiles <= blockIdx.x) {
        // this CTA had no tile: its partial must still read as zero
        for (int i = tid; i < pd.pack_fwd; i += blockDim.x) gp[i] = 0.f;
    }
    if (pd.act_kind == DRIL_ACT_CONTINUOUS && tid < pd.act_n) gp[pd.pack_fwd + tid] = (float)ls_acc;
    double sums[6] = {st_p, st_v, st_e, st_clip, st_kl, st_ratio};
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = block_sum(sums[i], sDbl);
        if (tid == 0) gp[pd.pack_fwd + pd.act_n + i] = (float)s;
    }
}

// g_flat[p] = sum over CTAs of the packed partials; stats6 likewise (fixed order => deterministic)
__global__ void __launch_bounds__(256) grad_reduce_kernel(const float* __restrict__ gpart, int n_cta, int gpack,
                                                          const int* __restrict__ flat2g, int n_params,
                                                          int stats_off, float* __restrict__ g_flat /* [n_params + 8] */,
                                                          const int* stop_flag) {
    if (*stop_flag) return;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_params + 6) return;
    int idx = p < n_params ? flat2g[p] : stats_off + (p - n_params);
    float s = 0.f;
    for (int c = 0; c < n_cta; ++c) s += gpart[(size_t)c * gpack + idx];
    g_flat[p] = s;
}

// iteration accumulators (device): [0..6] sums over applied minibatches of policy_loss, value_loss,
// entropy_loss, clip_fraction, approx_kl, entropy, ratio; [7] loss; [8] grad_norm sum; [9] applied
// count; [10] grad_norm count
#define ITER_ACC_N 12

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

// single CTA: global grad norm -> clip -> KL stop -> Adam -> refresh packed layouts
__global__ void __launch_bounds__(1024) adam_finalize_kernel(AdamArgs a) {
    __shared__ double scratch[32];
    __shared__ float s_scale;
    __shared__ int s_stop;
    const int tid = threadIdx.x;
    if (*a.stop_flag) return;
    double q = 0;
    for (int p = tid; p < a.n_params; p += blockDim.x) { double g = a.g[p]; q += g * g; }
    q = block_sum(q, scratch);
    if (tid == 0) {
        float norm = (float)sqrt(q);
        float scale = 1.f;
        if (a.hp.max_grad_norm >= 0.f && norm > a.hp.max_grad_norm) scale = a.hp.max_grad_norm / norm;   // no epsilon (optimization_utils.jl:99-107)
        s_scale = scale;
        int stop = 0;
        if (a.apply_stats) {
            const float* st = a.g + a.n_params;
            float invB = (float)(1.0 / a.global_count);
            float p_loss = st[0] * invB, v_loss = st[1] * invB, ent = st[2] * invB;
            float clipf = st[3] * invB, kl = st[4] * invB, ratio = st[5] * invB;
            a.iter_acc[8] += (double)norm;                 // grad_norms gets the PRE-clip norm, before the KL check (ppo.jl:216-223)
            a.iter_acc[10] += 1.0;
            if (a.hp.target_kl >= 0.f && kl > 1.5f * a.hp.target_kl) stop = 1;   // ppo.jl:235-238: stop BEFORE applying
            if (!stop) {
                float ent_loss = -ent;
                float loss = p_loss + a.hp.ent_coef * ent_loss + a.hp.vf_coef * v_loss;
                a.iter_acc[0] += p_loss; a.iter_acc[1] += v_loss; a.iter_acc[2] += ent_loss;
                a.iter_acc[3] += clipf; a.iter_acc[4] += kl; a.iter_acc[5] += ent; a.iter_acc[6] += ratio;
                a.iter_acc[7] += loss; a.iter_acc[9] += 1.0;
            } else {
                *a.stop_flag = 1;
            }
        } else {
            a.iter_acc[8] = (double)norm;
        }
        s_stop = stop;
        if (!stop) *a.step += 1;
    }
    __syncthreads();
    if (s_stop) return;
    const long long t = *a.step;
    const float c1 = (float)(1.0 - pow((double)a.hp.beta1, (double)t));
    const float c2 = (float)(1.0 - pow((double)a.hp.beta2, (double)t));
    const float scale = s_scale, b1 = a.hp.beta1, b2 = a.hp.beta2;
    for (int p = tid; p < a.n_params; p += blockDim.x) {
        float g = __fmul_rn(a.g[p], scale);
        float m = __fadd_rn(__fmul_rn(b1, a.m[p]), __fmul_rn(__fsub_rn(1.0f, b1), g));
        float v = __fadd_rn(__fmul_rn(b2, a.v[p]), __fmul_rn(__fmul_rn(__fsub_rn(1.0f, b2), g), g));
        float upd = __fmul_rn(__fdiv_rn(__fdiv_rn(m, c1), __fadd_rn(__fsqrt_rn(__fdiv_rn(v, c2)), a.hp.adam_eps)), a.hp.lr);
        float w = __fsub_rn(a.flat[p], upd);
        a.m[p] = m; a.v[p] = v; a.flat[p] = w;
        int ip = a.flat2pack[p];
        if (ip >= 0) a.pack[ip] = w;
        int it = a.flat2packT[p];
        if (it >= 0) a.pack[it] = w;
    }
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
