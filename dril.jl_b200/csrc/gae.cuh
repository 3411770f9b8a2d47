// GAE as a per-env reverse scan over the time-major buffer.
// Replaces compute_advantages! (buffers/trajectory.jl:80-102) applied to every trajectory and
// `returns = advantages + values` (buffers/rollout_buffer.jl:83-87).  The carry
// (A_next, V_next) is reset wherever a trajectory closes (terminated | truncated | last step):
//   terminated           -> delta = r - V
//   truncated            -> delta = r + gamma * V(terminal_obs) - V      (boot[t][n])
//   rollout end, running -> delta = r + gamma * V(new_obs)      - V      (last_values[n])
// HBM-bound: 9 B read + 8 B written per env-step (+ sparse bootstrap reads), thread per env,
// fully coalesced rows; the t-loop is unrolled so 8 rows of loads are in flight per thread.
#pragma once
#include "common.cuh"

#define GAE_UNROLL 8

__global__ void __launch_bounds__(256) gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                  const unsigned char* __restrict__ flags,
                                                  const float* __restrict__ boot, const float* __restrict__ last_values,
                                                  float* __restrict__ adv, float* __restrict__ ret, long long T,
                                                  long long N, float gamma, float lambda, double* ev_acc4) {
    __shared__ double scratch[32];
    long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = n < N;
    if (!live) n = N - 1;          // keep the thread for the block sums of the fused explained-variance moments (no stores)
    double sd = 0, sdd = 0, sr = 0, srr = 0;
    const float gl = __fmul_rn(gamma, lambda);
    float a_next = 0.f, v_next = 0.f;
    long long t = T - 1;
    // peel so that the main loop runs in blocks of GAE_UNROLL
    auto step = [&](long long tt, float r, float v, unsigned char f) {
        const bool term = f & 1, trunc = f & 2;
        const bool last = term || trunc || (tt == T - 1);
        float a;
        if (last) {
            if (term) a = __fsub_rn(r, v);
            else {
                float b = trunc ? boot[tt * N + n] : last_values[n];
                a = __fsub_rn(__fadd_rn(r, __fmul_rn(gamma, b)), v);
            }
        } else {
            float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(gamma, v_next)), v);
            a = __fadd_rn(delta, __fmul_rn(gl, a_next));
        }
        const float rt = __fadd_rn(a, v);
        if (live) {
            adv[tt * N + n] = a;
            ret[tt * N + n] = rt;
            const double dv = (double)v - (double)rt, dr = (double)rt;
            sd += dv; sdd += dv * dv; sr += dr; srr += dr * dr;
        }
        a_next = a; v_next = v;
    };
    while (t >= GAE_UNROLL - 1) {
        float r[GAE_UNROLL], v[GAE_UNROLL];
        unsigned char f[GAE_UNROLL];
#pragma unroll
        for (int i = 0; i < GAE_UNROLL; ++i) {
            long long idx = (t - i) * N + n;
            r[i] = rewards[idx]; v[i] = values[idx]; f[i] = flags[idx];
        }
#pragma unroll
        for (int i = 0; i < GAE_UNROLL; ++i) step(t - i, r[i], v[i], f[i]);
        t -= GAE_UNROLL;
    }
    for (; t >= 0; --t) step(t, rewards[t * N + n], values[t * N + n], flags[t * N + n]);
    // explained_variance moments of the same rollout (algorithms/ppo.jl:256) without a second pass over the buffer
    if (ev_acc4) {
        sd = block_sum(sd, scratch); sdd = block_sum(sdd, scratch);
        sr = block_sum(sr, scratch); srr = block_sum(srr, scratch);
        if (threadIdx.x == 0) {
            atomicAdd(&ev_acc4[0], sd); atomicAdd(&ev_acc4[1], sdd); atomicAdd(&ev_acc4[2], sr); atomicAdd(&ev_acc4[3], srr);
        }
    }
}

// explained_variance = 1 - var(values - returns) / var(returns) (algorithms/ppo.jl:256):
// four double sums over the whole buffer; finalised on the host side of the ABI.
__global__ void __launch_bounds__(256) explained_variance_kernel(const float* __restrict__ values,
                                                                 const float* __restrict__ returns, long long n_total,
                                                                 double* acc4) {
    __shared__ double scratch[32];
    double sd = 0, sdd = 0, sr = 0, srr = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += (long long)gridDim.x * blockDim.x) {
        double v = values[i], r = returns[i], d = v - r;
        sd += d; sdd += d * d; sr += r; srr += r * r;
    }
    sd = block_sum(sd, scratch); sdd = block_sum(sdd, scratch);
    sr = block_sum(sr, scratch); srr = block_sum(srr, scratch);
    if (threadIdx.x == 0) {
        atomicAdd(&acc4[0], sd); atomicAdd(&acc4[1], sdd); atomicAdd(&acc4[2], sr); atomicAdd(&acc4[3], srr);
    }
}

// MonitorWrapperEnv ring buffers (monitorWrapperEnv.jl:1-7,52-53): push this rollout's finished
// episodes in the reference's order (step ascending, env ascending) into the `window`-deep ring.
// Only rows t >= t0 can still be inside the window at the end, so the scan starts there.
// Single CTA; rows are scanned in chunks with an ordered block-level compaction.
struct MonitorRing {
    float* ret;        // [window]
    int* len;          // [window]
    long long* head;   // [0] total pushes so far
    int window;
};

__global__ void __launch_bounds__(1024) monitor_finalize_kernel(MonitorRing ring, const unsigned char* __restrict__ flags,
                                                                const float* __restrict__ ep_r, const int* __restrict__ ep_l,
                                                                const int* __restrict__ done_count, long long T, long long N) {
    __shared__ int warp_off[33];
    __shared__ long long s_head;
    __shared__ long long s_t0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        long long cum = 0, t0 = T;
        while (t0 > 0 && cum < ring.window) { --t0; cum += done_count[t0]; }
        s_t0 = t0;
        s_head = *ring.head;
    }
    __syncthreads();
    for (long long t = s_t0; t < T; ++t) {
        if (done_count[t] == 0) continue;
        for (long long base = 0; base < N; base += blockDim.x) {
            long long n = base + tid;
            bool done = n < N && (flags[t * N + n] & 3);
            unsigned ball = __ballot_sync(0xffffffffu, done);
            if (lane == 0) warp_off[warp + 1] = __popc(ball);
            __syncthreads();
            if (tid == 0) {
                warp_off[0] = 0;
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w) warp_off[w + 1] += warp_off[w];
            }
            __syncthreads();
            if (done) {
                long long pos = s_head + warp_off[warp] + __popc(ball & ((1u << lane) - 1u));
                int slot = (int)(pos % ring.window);
                // later pushes of the same chunk may wrap onto the same slot only if the chunk holds
                // more than `window` episodes; keep the last one (highest pos) in that case
                long long chunk_total = warp_off[blockDim.x >> 5];
                if (pos >= s_head + chunk_total - ring.window) {
                    ring.ret[slot] = ep_r[t * N + n];
                    ring.len[slot] = ep_l[t * N + n];
                }
            }
            __syncthreads();
            if (tid == 0) s_head += warp_off[blockDim.x >> 5];
            __syncthreads();
        }
    }
    if (tid == 0) *ring.head = s_head;
}

// ---------------------------------------------------------------------------------------------------------
// Everything dril_iteration_result reports, gathered into one record at the end of an iteration so that the host
// needs a single pinned device->host copy per iteration and may enqueue the next iteration before reading it.
// ---------------------------------------------------------------------------------------------------------
struct IterRecord {
    double acc[16];             // iter_acc (update.cuh ITER_ACC_N)
    double ev[4];               // explained-variance moments (already summed over ranks)
    double roll_sums[2];
    unsigned long long roll_eps;
    long long ring_count;       // min(pushes, window)
    float ring_rew_mean, ring_len_mean;   // mean(::CircularBuffer{Float32}) order: sequential over the ring slots
    int stop, p2p_err;
};
struct IterRecordSrc {
    const double* acc; const double* ev; const double* roll_sums; const unsigned long long* roll_eps;
    const int* stop; const int* p2p_err; MonitorRing ring; int has_ring;
};
__global__ void iter_record_kernel(IterRecordSrc s, IterRecord* out) {
    const int t = threadIdx.x;
    if (t < 16) out->acc[t] = s.acc[t];
    if (t < 4) out->ev[t] = s.ev[t];
    if (t < 2) out->roll_sums[t] = s.roll_sums ? s.roll_sums[t] : 0.0;
    if (t == 0) {
        out->roll_eps = s.roll_eps ? *s.roll_eps : 0ull;
        out->stop = *s.stop;
        out->p2p_err = s.p2p_err ? *s.p2p_err : 0;
    }
    {   // window means: lane-strided partial sums, then a fixed-order butterfly (deterministic)
        long long cnt = 0;
        float sr = 0.f;
        double sl = 0.0;
        if (s.has_ring) {
            const long long head = *s.ring.head;
            cnt = head < s.ring.window ? head : s.ring.window;
            for (long long i = t; i < cnt; i += 32) { sr += s.ring.ret[i]; sl += (double)s.ring.len[i]; }
        }
        sr = warp_sum(sr); sl = warp_sum(sl);
        if (t == 31) {
            out->ring_count = cnt;
            out->ring_rew_mean = cnt ? sr / (float)cnt : nanf("");
            out->ring_len_mean = cnt ? (float)(sl / (double)cnt) : nanf("");
        }
    }
}
