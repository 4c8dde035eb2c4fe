// GRU gate arithmetic shared by the tensor-core kernels (fp32, MUFU ex2/rcp).
//
// Follows torch rnn.py:1221-1224 as used by code/model.py:81, :412:
//     r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)      z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
//     n = tanh(W_in x + b_in + r * (W_hn h + b_hn))   h' = (1 - z) * n + z * h
// with every pre-activation pre-scaled on the way in (weights and constants by -log2 e for r, z and by 2 log2 e for n):
//     sigmoid(a) = 1 / (1 + 2^(-a log2 e))            tanh(a) = 1 - 2 / (1 + 2^(2 a log2 e))
// MUFU is the scarce pipe (16 lanes/clk/SM): reciprocals are shared -- 1/(d_r d_z) serves r and z of one pair,
// 1/(d_n0 d_n1) serves the n gates of two pairs -- which brings a (unit, stream) pair from 6 to 4 MUFU operations.
// Arguments of ex2 are clamped to 60 so that a product of two denominators stays below 2^121.
#pragma once
#include "ntm_common.cuh"

namespace ntm {

struct UnitConst {
    float cr_w, cr_b, cz_w, cz_b, cn_w, cn_b, ch_b, wo;
};

__device__ __forceinline__ UnitConst load_unit_const(const float* __restrict__ blob, int j)
{
    constexpr float L = 1.4426950408889634f;
    UnitConst c;
    c.cr_w = -L * blob[BlobLayout::W_IH + j];
    c.cr_b = -L * (blob[BlobLayout::B_IH + j] + blob[BlobLayout::B_HH + j]);
    c.cz_w = -L * blob[BlobLayout::W_IH + 64 + j];
    c.cz_b = -L * (blob[BlobLayout::B_IH + 64 + j] + blob[BlobLayout::B_HH + 64 + j]);
    c.cn_w = 2.0f * L * blob[BlobLayout::W_IH + 128 + j];
    c.cn_b = 2.0f * L * blob[BlobLayout::B_IH + 128 + j];
    c.ch_b = 2.0f * L * blob[BlobLayout::B_HH + 128 + j];
    c.wo = blob[BlobLayout::W_OUT + j];
    return c;
}

constexpr float EX2_CLAMP = 60.0f;

// r, z and the denominator of the n gate of one pair.  ar, az, an: scaled W_h* h products from the tensor core.
__device__ __forceinline__ void gates_rz_dn(const UnitConst& c, float ar, float az, float an, float x, float& z, float& dn)
{
    const float dr = 1.0f + ex2_approx(fminf(ar + fmaf(c.cr_w, x, c.cr_b), EX2_CLAMP));
    const float dz = 1.0f + ex2_approx(fminf(az + fmaf(c.cz_w, x, c.cz_b), EX2_CLAMP));
    const float rinv = rcp_approx(dr * dz);
    const float r = dz * rinv;
    z = dr * rinv;
    dn = 1.0f + ex2_approx(fminf(fmaf(r, an + c.ch_b, fmaf(c.cn_w, x, c.cn_b)), EX2_CLAMP));
}

// the same with complete pre-activations: ar, az already contain W_i x + b (K-augmented MMA), an = W_hn h + b_hn,
// gin = W_in x + b_in (all scaled)
__device__ __forceinline__ void gates_rz_dn_pre(float ar, float az, float an, float gin, float& z, float& dn)
{
    const float dr = 1.0f + ex2_approx(fminf(ar, EX2_CLAMP));
    const float dz = 1.0f + ex2_approx(fminf(az, EX2_CLAMP));
    const float rinv = rcp_approx(dr * dz);
    const float r = dz * rinv;
    z = dr * rinv;
    dn = 1.0f + ex2_approx(fminf(fmaf(r, an, gin), EX2_CLAMP));
}

// new states of two pairs (their n-gate denominators share one reciprocal).  Only pair up the SAME stream (two hidden
// units): a reciprocal shared between two streams would make a stream's rounding depend on its neighbour.
__device__ __forceinline__ void gates_blend2(float z0, float dn0, float h0, float z1, float dn1, float h1, float& hn0, float& hn1)
{
    const float qinv = rcp_approx(dn0 * dn1);
    const float n0 = fmaf(-2.0f, dn1 * qinv, 1.0f);
    const float n1 = fmaf(-2.0f, dn0 * qinv, 1.0f);
    hn0 = fmaf(z0, h0 - n0, n0);
    hn1 = fmaf(z1, h1 - n1, n1);
}

// new state of one pair (own reciprocal: keeps a stream's arithmetic independent of its neighbours)
__device__ __forceinline__ float gates_blend1(float z, float dn, float h)
{
    const float n = fmaf(-2.0f, rcp_approx(dn), 1.0f);
    return fmaf(z, h - n, n);
}

}  // namespace ntm
