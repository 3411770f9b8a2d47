// Batched classic-control dynamics (device functions).
//
// The reference takes CartPole / Pendulum from the un-vendored ClassicControlEnvironments.jl
// (call sites README.md:50,76; benchmark/bench_utils.jl:14,20); these follow the Gymnasium
// CartPole-v1 / Pendulum-v1 equations exactly as restated in oracle/envs.py.  Every fp32
// operation uses a round-to-nearest intrinsic (never contracted into FMA) and sin/cos are the
// correctly rounded fp32 values (fp64 evaluation, one rounding), so replayed action sequences
// reproduce the oracle bit for bit.
#pragma once
#include "common.cuh"

#define ENV_MAX_STATE 4

__device__ __forceinline__ void sincos_rn(float th, float* s, float* c) {
    double sd, cd;
    sincos((double)th, &sd, &cd);
    *s = (float)sd;
    *c = (float)cd;
}

// raw observation d-th component from state (CartPole / Pendulum only)
__device__ __forceinline__ void env_raw_obs(int kind, const float* st, float* o) {
    if (kind == DRIL_ENV_CARTPOLE) {
        o[0] = st[0]; o[1] = st[1]; o[2] = st[2]; o[3] = st[3];
    } else {  // pendulum: (cos th, sin th, thdot)
        float s, c;
        sincos_rn(st[0], &s, &c);
        o[0] = c; o[1] = s; o[2] = st[1]; o[3] = 0.f;
    }
}

// observation as the wrappers above the env see it: raw, or mapped to [-1, 1] by ScalingWrapperEnv.observe
// (scalingWrapperEnv.jl:72-75,94-99: `(x - offset) * scale - 1`, three separate fp32 roundings)
__device__ __forceinline__ void env_obs(const EnvDev& env, const float* st, float* o) {
    env_raw_obs(env.kind, st, o);
    if (env.scaling) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < env.obs_dim) o[j] = __fsub_rn(__fmul_rn(__fsub_rn(o[j], env.sc_obs_o[j]), env.sc_obs_f[j]), 1.0f);
    }
}
// ScalingWrapperEnv.act! (scalingWrapperEnv.jl:77-80,112-115): `(a + 1) / scale + offset` before the wrapped env's act!
__device__ __forceinline__ float env_unscale_action(const EnvDev& env, float a) {
    return env.scaling ? __fadd_rn(__fdiv_rn(__fadd_rn(a, 1.0f), env.sc_act_f), env.sc_act_o) : a;
}

__device__ __forceinline__ void env_reset_state(int kind, uint32_t gid, uint32_t episode, unsigned long long seed,
                                                float* st) {
    uint32_t x[4];
    philox4x32(gid, episode, 0u, DRIL_TAG_RESET, seed, x);
    if (kind == DRIL_ENV_CARTPOLE) {
#pragma unroll
        for (int k = 0; k < 4; ++k) st[k] = __fadd_rn(-0.05f, __fmul_rn(0.1f, u01_f32(x[k])));
    } else if (kind == DRIL_ENV_PENDULUM) {
        st[0] = __fadd_rn(-3.14159274101257324f, __fmul_rn(6.28318548202514648f, u01_f32(x[0])));
        st[1] = __fadd_rn(-1.0f, __fmul_rn(2.0f, u01_f32(x[1])));
        st[2] = 0.f; st[3] = 0.f;
    }
}

// CartPole-v1 Euler step with the (correctly rounded) sin/cos of the pole angle supplied by the caller.
// a01: 0 = push left, 1 = push right. Returns reward; sets terminated.
__device__ __forceinline__ float cartpole_step_sc(float* st, int a01, float sinth, float costh, bool* terminated) {
    const float GRAVITY = 9.8f, MASSPOLE = 0.1f, LENGTH = 0.5f, FORCE_MAG = 10.0f, TAU = 0.02f;
    const float TOTAL_MASS = __fadd_rn(0.1f, 1.0f);
    const float PML = __fmul_rn(0.1f, 0.5f);
    const float THETA_THR = 0.20943951023931953f, X_THR = 2.4f, FOUR_THIRDS = 1.3333333333333333f;
    float x = st[0], x_dot = st[1], th = st[2], th_dot = st[3];
    float force = (a01 == 1) ? FORCE_MAG : -FORCE_MAG;
    float temp = __fdiv_rn(__fadd_rn(force, __fmul_rn(__fmul_rn(PML, __fmul_rn(th_dot, th_dot)), sinth)), TOTAL_MASS);
    float den = __fmul_rn(LENGTH, __fsub_rn(FOUR_THIRDS, __fdiv_rn(__fmul_rn(MASSPOLE, __fmul_rn(costh, costh)), TOTAL_MASS)));
    float thetaacc = __fdiv_rn(__fsub_rn(__fmul_rn(GRAVITY, sinth), __fmul_rn(costh, temp)), den);
    float xacc = __fsub_rn(temp, __fdiv_rn(__fmul_rn(__fmul_rn(PML, thetaacc), costh), TOTAL_MASS));
    float xn = __fadd_rn(x, __fmul_rn(TAU, x_dot));
    float xdn = __fadd_rn(x_dot, __fmul_rn(TAU, xacc));
    float thn = __fadd_rn(th, __fmul_rn(TAU, th_dot));
    float thdn = __fadd_rn(th_dot, __fmul_rn(TAU, thetaacc));
    st[0] = xn; st[1] = xdn; st[2] = thn; st[3] = thdn;
    *terminated = (xn < -X_THR) || (xn > X_THR) || (thn < -THETA_THR) || (thn > THETA_THR);
    return 1.0f;
}
__device__ __forceinline__ float cartpole_step(float* st, int a01, bool* terminated) {
    float sinth, costh;
    sincos_rn(st[2], &sinth, &costh);
    return cartpole_step_sc(st, a01, sinth, costh, terminated);
}

// Pendulum-v1 step (g = 10, m = l = 1). u: env-space torque (clipped again like Gymnasium).
__device__ __forceinline__ float pendulum_step(float* st, float u) {
    const float MAX_SPEED = 8.0f, MAX_TORQUE = 2.0f, DT = 0.05f;
    const float PI = 3.14159274101257324f, TWO_PI = 6.28318548202514648f;
    u = fminf(fmaxf(u, -MAX_TORQUE), MAX_TORQUE);
    float th = st[0], thdot = st[1];
    float xp = __fadd_rn(th, PI);
    float an = __fsub_rn(__fsub_rn(xp, __fmul_rn(TWO_PI, floorf(__fdiv_rn(xp, TWO_PI)))), PI);
    float cost = __fadd_rn(__fadd_rn(__fmul_rn(an, an), __fmul_rn(0.1f, __fmul_rn(thdot, thdot))),
                           __fmul_rn(0.001f, __fmul_rn(u, u)));
    float sinth, costh;
    sincos_rn(th, &sinth, &costh);
    float nthdot = __fadd_rn(thdot, __fmul_rn(__fadd_rn(__fmul_rn(15.0f, sinth), __fmul_rn(3.0f, u)), DT));
    nthdot = fminf(fmaxf(nthdot, -MAX_SPEED), MAX_SPEED);
    float nth = __fadd_rn(th, __fmul_rn(nthdot, DT));
    st[0] = nth; st[1] = nthdot;
    return -cost;
}

// Synthetic env: obs block b of lifetime step `life`
__device__ __forceinline__ void synthetic_obs_block(uint32_t gid, uint32_t life, int b, unsigned long long seed, float o[4]) {
    uint32_t x[4];
    philox4x32(gid, life, (uint32_t)b, DRIL_TAG_SYN_OBS, seed, x);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = __fadd_rn(-1.0f, __fmul_rn(2.0f, u01_f32(x[j])));
}
__device__ __forceinline__ float synthetic_step(uint32_t gid, uint32_t life, unsigned long long seed, bool* terminated) {
    uint32_t x[4];
    philox4x32(gid, life, 0u, DRIL_TAG_SYN_DYN, seed, x);
    *terminated = u01_f32(x[1]) < 0.005f;
    return u01_f32(x[0]);
}
