// Warp-level tensor-core tiles (mma.sync.m16n8k8 TF32, 3xTF32 split) for the wide layers of the general-shape
// loss/grad kernel (update.cuh) and the general rollout kernel (rollout.cuh): hidden layers whose padded dims are multiples
// of 16, on sample tiles whose width is a multiple of 16 (128 in the update when shared memory allows, 64 in the rollout).
//
// The tcgen05 kernel (update_tc.cuh) covers the reference's default [64,64] discrete policy; every other shape ran on
// fp32 FMA tiles at ~31 % of the FMA pipe (40 MAC/clk/SM).  The legacy warp-level tensor path issues 512 TF32 MAC/clk/SM
// on sm_100a (tools/mma_probe.cu), i.e. 170 MAC/clk/SM after the 3-way split that keeps fp32-level accuracy
// (hi = x with 13 mantissa bits cleared, lo = x - hi; hi*hi + lo*hi + hi*lo, fp32 accumulation; relative error ~2^-21).
//
// The tensor core reads only the upper 19 bits of a TF32 operand, so only `lo` costs instructions: ptxas drops the mask of
// every `hi` that feeds nothing but an mma (seen in the SASS: HMMA.1688.F32.TF32 takes the raw LDS result for the hi*hi and
// lo*hi products), leaving one LOP3 + one FADD per operand element.
//
// Operand layouts are the ones the FMA tiles use, so the two kinds mix freely inside a layer loop:
//   activations feature-major in shared memory  act[f * ld + m]   (m = sample in the tile, ld = tile width + 4)
//   weights in the packed global layout          W[k * Np + n], Wt[n * Kp + k]   (read once per tile and CTA, L2)
// Fragment coordinates (g = lane >> 2, t = lane & 3), PTX ISA m16n8k8 .tf32:
//   A (16x8, row): a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)
//   B (8x8,  col): b0 (k = t, n = g)       b1 (k = t+4, n = g)
//   C (16x8):      c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
#pragma once
#include "common.cuh"

#define MMA_TILE_M 128    // samples per tile of the loss kernel when shared memory allows (any multiple of 16 works)

// a layer (padded dims Kp x Np) runs on these tiles; the one predicate host planning (api.cu) and the kernels share
__host__ __device__ inline bool mma_layer_ok(int Kp, int Np) { return (Kp & 15) == 0 && (Np & 15) == 0 && Kp >= 16 && Np >= 16; }

// not volatile: pure function of its operands, so the compiler may interleave independent accumulators
__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_red_add_v2(float* addr, float x, float y) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(x), "f"(y) : "memory");
}

// out[r][m] = sum_c Wm[c * ldw + r] * In[c * ld + m]   for the RB row blocks r in [r0, r0 + 16 RB), samples m in
// [8*mb0, 8*(mb0+MB)), contraction over c in [0, C) (C multiple of 8).  Wm global (or shared), row stride ldw; In shared.
// A warp keeps RB x MB accumulator tiles: every In fragment is split once and used by RB row blocks, every weight
// fragment by MB sample blocks; the three products of a split go to RB*2 different accumulators in turn so that
// dependent MMAs are 2 RB issues apart.
//   EPI 0 (forward):   Out[r][m] = f(acc + bias[r])           f = tanh if apply_tanh
//   EPI 1 (dH):        Out[r][m] = acc * (1 - Out[r][m]^2)     in place over the activation H
template <int RB, int MB, int EPI>
__device__ __forceinline__ void mma_rows_unit(const float* __restrict__ Wm, int ldw, int C, const float* __restrict__ In,
                                              float* __restrict__ Out, int ld, int r0, int mb0, const float* __restrict__ bias,
                                              bool apply_tanh) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float acc[RB][MB][4];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int i = 0; i < MB; ++i) { acc[r][i][0] = acc[r][i][1] = acc[r][i][2] = acc[r][i][3] = 0.f; }
    // A fragment rows = output features r (from the weights), columns = contraction index
    const float* wp = Wm + (size_t)t * ldw + r0 + g;
    const float* ip = In + (size_t)t * ld + mb0 * 8 + g;
    float w[RB][4];
#pragma unroll
    for (int r = 0; r < RB; ++r) { w[r][0] = wp[r * 16]; w[r][1] = wp[r * 16 + 8]; w[r][2] = wp[(size_t)4 * ldw + r * 16]; w[r][3] = wp[(size_t)4 * ldw + r * 16 + 8]; }
    for (int c0 = 0; c0 < C; c0 += 8) {
        uint32_t ahi[RB][4], alo[RB][4];
#pragma unroll
        for (int r = 0; r < RB; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_split(w[r][q], ahi[r][q], alo[r][q]);
        if (c0 + 8 < C) {          // prefetch the next weight fragments (L2 latency) under this step's MMAs
            const float* wn = wp + (size_t)(c0 + 8) * ldw;
#pragma unroll
            for (int r = 0; r < RB; ++r) { w[r][0] = wn[r * 16]; w[r][1] = wn[r * 16 + 8]; w[r][2] = wn[(size_t)4 * ldw + r * 16]; w[r][3] = wn[(size_t)4 * ldw + r * 16 + 8]; }
        }
        const float* ic = ip + (size_t)c0 * ld;
#pragma unroll
        for (int i = 0; i < MB; i += 2) {
            uint32_t bh[2][2], bl[2][2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                mma_split(ic[(i + j) * 8], bh[j][0], bl[j][0]);
                mma_split(ic[(size_t)4 * ld + (i + j) * 8], bh[j][1], bl[j][1]);
            }
#pragma unroll
            for (int r = 0; r < RB; ++r)
#pragma unroll
                for (int j = 0; j < 2; ++j) mma_tf32(acc[r][i + j], alo[r], bh[j][0], bh[j][1]);
#pragma unroll
            for (int r = 0; r < RB; ++r)
#pragma unroll
                for (int j = 0; j < 2; ++j) mma_tf32(acc[r][i + j], ahi[r], bl[j][0], bl[j][1]);
#pragma unroll
            for (int r = 0; r < RB; ++r)
#pragma unroll
                for (int j = 0; j < 2; ++j) mma_tf32(acc[r][i + j], ahi[r], bh[j][0], bh[j][1]);
        }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        const int ra = r0 + r * 16 + g, rb = ra + 8;
        float ba = 0.f, bb = 0.f;
        if (EPI == 0) { ba = bias[ra]; bb = bias[rb]; }
#pragma unroll
        for (int i = 0; i < MB; ++i) {
            const int m = (mb0 + i) * 8 + 2 * t;
            float2* pa = reinterpret_cast<float2*>(Out + (size_t)ra * ld + m);
            float2* pb = reinterpret_cast<float2*>(Out + (size_t)rb * ld + m);
            float2 va, vb;
            if (EPI == 0) {
                va = make_float2(acc[r][i][0] + ba, acc[r][i][1] + ba);
                vb = make_float2(acc[r][i][2] + bb, acc[r][i][3] + bb);
                if (apply_tanh) { va.x = fast_tanh(va.x); va.y = fast_tanh(va.y); vb.x = fast_tanh(vb.x); vb.y = fast_tanh(vb.y); }
            } else {
                const float2 ha = *pa, hb = *pb;
                va = make_float2(acc[r][i][0] * (1.0f - ha.x * ha.x), acc[r][i][1] * (1.0f - ha.y * ha.y));
                vb = make_float2(acc[r][i][2] * (1.0f - hb.x * hb.x), acc[r][i][3] * (1.0f - hb.y * hb.y));
            }
            *pa = va; *pb = vb;
        }
    }
}

// All (row-block group, sample chunk) units of one GEMM of this shape over the CTA's warps.  R = number of output rows
// (multiple of 16), mblocks = sample blocks of 8 in the tile (even: tile width multiple of 16; 16 for the loss kernel's
// 128-sample tiles, less in the rollout).  Caller synchronises afterwards.
template <int RB, int MB, int EPI>
__device__ __forceinline__ void mma_rows_units(const float* __restrict__ Wm, int ldw, int R, int C, const float* __restrict__ In,
                                               float* __restrict__ Out, int ld, const float* __restrict__ bias, bool apply_tanh,
                                               int mblocks) {
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int chunks = mblocks / MB, units = (R / (16 * RB)) * chunks;
    for (int u = warp; u < units; u += nwarps)
        mma_rows_unit<RB, MB, EPI>(Wm, ldw, C, In, Out, ld, (u / chunks) * 16 * RB, (u % chunks) * MB, bias, apply_tanh);
}
template <int EPI>
__device__ __forceinline__ void mma_rows_layer(const float* __restrict__ Wm, int ldw, int R, int C, const float* __restrict__ In,
                                               float* __restrict__ Out, int ld, const float* __restrict__ bias, bool apply_tanh,
                                               int mblocks = MMA_TILE_M / 8) {
    const int nwarps = blockDim.x >> 5;
    const int rblocks = R >> 4;
    const bool r2 = (rblocks & 1) == 0;
    // largest accumulator tile (most operand reuse) whose unit count divides over the warps; two row blocks preferred
    if (r2 && (mblocks & 7) == 0 && ((rblocks >> 1) * (mblocks >> 3)) % nwarps == 0)
        mma_rows_units<2, 8, EPI>(Wm, ldw, R, C, In, Out, ld, bias, apply_tanh, mblocks);
    else if ((mblocks & 7) == 0 && (rblocks * (mblocks >> 3)) % nwarps == 0)
        mma_rows_units<1, 8, EPI>(Wm, ldw, R, C, In, Out, ld, bias, apply_tanh, mblocks);
    else if (r2 && (mblocks & 3) == 0 && ((rblocks >> 1) * (mblocks >> 2)) % nwarps == 0)
        mma_rows_units<2, 4, EPI>(Wm, ldw, R, C, In, Out, ld, bias, apply_tanh, mblocks);
    else if ((mblocks & 3) == 0 && (rblocks * (mblocks >> 2)) % nwarps == 0)
        mma_rows_units<1, 4, EPI>(Wm, ldw, R, C, In, Out, ld, bias, apply_tanh, mblocks);
    else if (r2)
        mma_rows_units<2, 2, EPI>(Wm, ldw, R, C, In, Out, ld, bias, apply_tanh, mblocks);
    else
        mma_rows_units<1, 2, EPI>(Wm, ldw, R, C, In, Out, ld, bias, apply_tanh, mblocks);
}

// dW[k][n] (+)= sum_m Ain[k * ld + m] * dZ[n * ld + m] over the M samples of the tile (multiple of 8), the KB row blocks k in
// [k0, k0 + 16 KB), columns n in [8*nb0, 8*(nb0+NB)); written (first tile of the pass) or red-added into this CTA's
// packed partial.  Same accumulator interleaving as mma_rows_unit.
template <int KB, int NB>
__device__ __forceinline__ void mma_dw_unit(const float* __restrict__ Ain, const float* __restrict__ dZ, int ld, int k0, int nb0,
                                            float* __restrict__ gW, int Np, bool first, int M) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float acc[KB][NB][4];
#pragma unroll
    for (int r = 0; r < KB; ++r)
#pragma unroll
        for (int i = 0; i < NB; ++i) { acc[r][i][0] = acc[r][i][1] = acc[r][i][2] = acc[r][i][3] = 0.f; }
    const float* ap = Ain + (size_t)(k0 + g) * ld + t;          // A (row k, col m): conflict-free (bank 4g + t)
    const float* zp = dZ + (size_t)(nb0 * 8 + g) * ld + t;      // B (m, col n)
    for (int m0 = 0; m0 < M; m0 += 8) {
        uint32_t ahi[KB][4], alo[KB][4];
#pragma unroll
        for (int r = 0; r < KB; ++r) {
            mma_split(ap[(size_t)(r * 16) * ld + m0], ahi[r][0], alo[r][0]);
            mma_split(ap[(size_t)(r * 16 + 8) * ld + m0], ahi[r][1], alo[r][1]);
            mma_split(ap[(size_t)(r * 16) * ld + m0 + 4], ahi[r][2], alo[r][2]);
            mma_split(ap[(size_t)(r * 16 + 8) * ld + m0 + 4], ahi[r][3], alo[r][3]);
        }
#pragma unroll
        for (int i = 0; i < NB; i += 2) {
            uint32_t bh[2][2], bl[2][2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                mma_split(zp[(size_t)(i + j) * 8 * ld + m0], bh[j][0], bl[j][0]);
                mma_split(zp[(size_t)(i + j) * 8 * ld + m0 + 4], bh[j][1], bl[j][1]);
            }
#pragma unroll
            for (int r = 0; r < KB; ++r)
#pragma unroll
                for (int j = 0; j < 2; ++j) mma_tf32(acc[r][i + j], alo[r], bh[j][0], bh[j][1]);
#pragma unroll
            for (int r = 0; r < KB; ++r)
#pragma unroll
                for (int j = 0; j < 2; ++j) mma_tf32(acc[r][i + j], ahi[r], bl[j][0], bl[j][1]);
#pragma unroll
            for (int r = 0; r < KB; ++r)
#pragma unroll
                for (int j = 0; j < 2; ++j) mma_tf32(acc[r][i + j], ahi[r], bh[j][0], bh[j][1]);
        }
    }
#pragma unroll
    for (int r = 0; r < KB; ++r)
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int n = (nb0 + i) * 8 + 2 * t;
            float* pa = gW + (size_t)(k0 + r * 16 + g) * Np + n;
            float* pb = gW + (size_t)(k0 + r * 16 + g + 8) * Np + n;
            if (first) {
                *reinterpret_cast<float2*>(pa) = make_float2(acc[r][i][0], acc[r][i][1]);
                *reinterpret_cast<float2*>(pb) = make_float2(acc[r][i][2], acc[r][i][3]);
            } else {
                mma_red_add_v2(pa, acc[r][i][0], acc[r][i][1]);
                mma_red_add_v2(pb, acc[r][i][2], acc[r][i][3]);
            }
        }
}

// dW of one layer (Kp x Np, both multiples of 16) over the CTA's warps.  Caller synchronises afterwards.
__device__ __forceinline__ void mma_dw_layer(const float* __restrict__ Ain, const float* __restrict__ dZ, int ld, int Kp, int Np,
                                             float* __restrict__ gW, bool first, int M = MMA_TILE_M) {
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int kblocks = Kp >> 4, nblocks = Np >> 3;      // nblocks is even
    if ((kblocks & 1) == 0 && (nblocks & 7) == 0 && ((kblocks >> 1) * (nblocks >> 3)) % nwarps == 0) {          // KB 2 x NB 8
        const int per = nblocks >> 3;
        for (int u = warp; u < (kblocks >> 1) * per; u += nwarps) mma_dw_unit<2, 8>(Ain, dZ, ld, (u / per) << 5, (u % per) * 8, gW, Np, first, M);
    } else if ((kblocks & 1) == 0 && (nblocks & 3) == 0 && ((kblocks >> 1) * (nblocks >> 2)) % nwarps == 0) {   // KB 2 x NB 4
        const int per = nblocks >> 2;
        for (int u = warp; u < (kblocks >> 1) * per; u += nwarps) mma_dw_unit<2, 4>(Ain, dZ, ld, (u / per) << 5, (u % per) * 4, gW, Np, first, M);
    } else if ((kblocks & 1) == 0) {                                                                            // KB 2 x NB 2
        const int per = nblocks >> 1;
        for (int u = warp; u < (kblocks >> 1) * per; u += nwarps) mma_dw_unit<2, 2>(Ain, dZ, ld, (u / per) << 5, (u % per) * 2, gW, Np, first, M);
    } else {                                                                                                    // KB 1 x NB 2
        const int per = nblocks >> 1;
        for (int u = warp; u < kblocks * per; u += nwarps) mma_dw_unit<1, 2>(Ain, dZ, ld, (u / per) << 4, (u % per) * 2, gW, Np, first, M);
    }
}
