// tfm_jacobian_common.cuh -- device helpers shared by the two translation units of K2
// (tfm_jacobian.cu: parity mode + general fast path; tfm_jacobian_fast.cu: the lean Gaussian kernels).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "tfm_chain.cuh"
#include "tfm_common.cuh"

namespace tfm {

// ---- 2x2 SVD / recompose (helpers.pyx:1845-1916) ----
__device__ __forceinline__ void svd2x2_dev(const double *M, double *U, double *s, double *V)
{
    const double a = M[0], b = M[1], c = M[2], d = M[3];
    const double au = __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b));
    const double bu = __dadd_rn(__dmul_rn(a, c), __dmul_rn(b, d));
    const double du = __dadd_rn(__dmul_rn(c, c), __dmul_rn(d, d));
    const double theta = __dmul_rn(0.5, atan2(__dmul_rn(2.0, bu), __dsub_rn(au, du)));
    double St, Ct;
    sincos(theta, &St, &Ct);
    U[0] = Ct; U[1] = -St; U[2] = St; U[3] = Ct;
    const double aw = __dadd_rn(__dmul_rn(a, a), __dmul_rn(c, c));
    const double bw = __dadd_rn(__dmul_rn(a, b), __dmul_rn(c, d));
    const double dw = __dadd_rn(__dmul_rn(b, b), __dmul_rn(d, d));
    const double phi = __dmul_rn(0.5, atan2(__dmul_rn(2.0, bw), __dsub_rn(aw, dw)));
    double Sp, Cp;
    sincos(phi, &Sp, &Cp);
    const double cs00 = __dadd_rn(
        __dmul_rn(__dadd_rn(__dmul_rn(Ct, a), __dmul_rn(St, c)), Cp),
        __dmul_rn(__dadd_rn(__dmul_rn(Ct, b), __dmul_rn(St, d)), Sp));
    const double cs11 = __dsub_rn(
        __dmul_rn(__dsub_rn(__dmul_rn(St, a), __dmul_rn(Ct, c)), Sp),
        __dmul_rn(__dsub_rn(__dmul_rn(St, b), __dmul_rn(Ct, d)), Cp));
    s[0] = fabs(cs00);
    s[1] = fabs(cs11);
    const double c00 = cs00 >= 0 ? 1.0 : -1.0, c11 = cs11 >= 0 ? 1.0 : -1.0;
    V[0] = __dmul_rn(Cp, c00); V[1] = __dmul_rn(-Sp, c11);
    V[2] = __dmul_rn(Sp, c00); V[3] = __dmul_rn(Cp, c11);
}

__device__ __forceinline__ void svc2x2_dev(const double *U, const double *s, const double *V, double *M)
{
    const double U00S0 = __dmul_rn(U[0], s[0]), U10S0 = __dmul_rn(U[2], s[0]);
    const double U01S1 = __dmul_rn(U[1], s[1]), U11S1 = __dmul_rn(U[3], s[1]);
    M[0] = __dadd_rn(__dmul_rn(U00S0, V[0]), __dmul_rn(U01S1, V[1]));
    M[1] = __dadd_rn(__dmul_rn(U00S0, V[2]), __dmul_rn(U01S1, V[3]));
    M[2] = __dadd_rn(__dmul_rn(U10S0, V[0]), __dmul_rn(U11S1, V[1]));
    M[3] = __dadd_rn(__dmul_rn(U10S0, V[2]), __dmul_rn(U11S1, V[3]));
}

// generic boundary on one axis; same rules as K1 / helpers.pyx:1750-1790 with R9
__device__ __forceinline__ int jbound(long long i, int size, int mode, int &state)
{
    if ((unsigned long long)i < (unsigned long long)size) return (int)i;
    switch (mode) {
    case TFM_B_EXTEND: return i < 0 ? 0 : size - 1;
    case TFM_B_PERIODIC: { long long r = i % (long long)size; if (r < 0) r += size; return (int)r; }
    case TFM_B_MIRROR: {
        long long s2 = 2LL * size, r = i % s2;
        if (r < 0) r += s2;
        if (r >= size) r = s2 - 1 - r;
        return (int)r; }
    case TFM_B_FORBID: state |= 2; return 0;
    default: state |= 1; return 0;           // truncate
    }
}

// single-instruction MUFU approximations (fast mode only; 1-2 ulp, flush-to-zero)
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// f(lambda) = (padval + lambda^(p/2))^(-2/p): the squared padded inverse singular value as a function
// of the squared singular value lambda (helpers.pyx:1661-1662 applied to s = sqrt(lambda)).
__device__ __forceinline__ float pad_inv_sq(float lam, float padval, float half_pow, float m2_over_pow)
{
    // lam = 0 -> lg2 = -inf -> 0.  Above 2^60 the padding constant is below fp32 resolution and the
    // power may overflow fp32 (pad_pow = 64 with s >= 4): the logarithm of the sum is the exponent itself
    const float e = half_pow * fast_lg2(lam);
    const float lg = (e > 60.0f) ? e : fast_lg2(padval + fast_ex2(e));
    return fast_ex2(m2_over_pow * lg);
}

// Gaussian footprint weights need only Q = J^T J with J = V diag(s') U^T, i.e. Q = U diag(s'^2) U^T
// = f(M M^T) for M = Ji = U diag(s) V^T: a function of a symmetric 2x2 matrix, which is always
// alpha*I + beta*(M M^T).  No angles, no sincos -- eigenvalues by one square root.
// Returns 1/smin (the footprint half-width is ceil(oblur/smin)).
__device__ __forceinline__ void gaussian_quadratic_f32(const double *Ji, float padval, float pad_pow,
                                                       float &q00, float &q01, float &q11, float &inv_smin)
{
    const float a = (float)Ji[0], b = (float)Ji[1], c = (float)Ji[2], d = (float)Ji[3];
    const float g00 = a * a + b * b, g01 = a * c + b * d, g11 = c * c + d * d;
    const float mean = 0.5f * (g00 + g11), dif = 0.5f * (g00 - g11);
    const float r2 = dif * dif + g01 * g01;
    const float rad = (r2 > 0.0f) ? r2 * fast_rsqrt(r2) : 0.0f;
    const float lp = mean + rad, lm = fmaxf(mean - rad, 0.0f);
    const float half_pow = 0.5f * pad_pow, m2p = -2.0f * fast_rcp(pad_pow);
    const float fp = pad_inv_sq(lp, padval, half_pow, m2p), fm = pad_inv_sq(lm, padval, half_pow, m2p);
    // f(G) = fp I + beta (G - lp I), beta the divided difference of f over the two eigenvalues.  For a
    // nearly isotropic Jacobian beta is ill-conditioned (error ~ eps f / (lp - lm)), but it multiplies
    // G - lp I, whose entries are O(lp - lm) when formed from rad and dif directly: the error of Q
    // stays ~ eps f.  (alpha + beta g with alpha = fp - beta lp cancels catastrophically there.)
    const float dl = lp - lm;
    const float beta = (dl > 0.0f) ? (fp - fm) * fast_rcp(dl) : 0.0f;
    q00 = fp - beta * (rad - dif);         // lp - g00 = rad - dif
    q01 = beta * g01;
    q11 = fp - beta * (rad + dif);         // lp - g11 = rad + dif
    inv_smin = fast_rsqrt(fp);             // the larger singular value has the smaller padded inverse
}

// ---------------------------------------------------------------------------------------------
// Gaussian footprint through the texture unit (fast mode): shared by the lean kernels of
// tfm_jacobian_fast.cu and the general kernel of tfm_jacobian.cu
// ---------------------------------------------------------------------------------------------
// Weight state of the texture form, all of it packed over the two columns of a gather (lane .x = column c,
// lane .y = column c + 1), so that no instruction is spent moving scalars in and out of register pairs (ncu
// source view of the first form: 54 MOV-type instructions per pixel, 11 % of the kernel):
//   W  = (w(c, b), w(c+1, b))               weights of the current column pair
//   R  = (r(c)^2 cA, r(c+1)^2 cA)           W(c+2) = W(c) R(c),  R(c+2) = R(c) cA^4
//   RB = (rb(0, b), rb(1, b))               first pair of the next row: W(0, b+1) = W(0, b) RB(b),
//                                           RB(b+1) = RB(b) cC;  R(0, b+1) = R(0, b) cB^2
struct TexRows {
    float2 W, R, RB;
    float cA4, cB2, cC;
};

__device__ __forceinline__ float2 bc2(float v) { return make_float2(v, v); }

// one row pair (r, r+1) of the texture form, texels already fetched; BOTH = false: row r+1 lies outside
// the square.  Leaves the row state on row r + 2.
template <int HALF, bool BOTH>
__device__ __forceinline__ void tex_row_pair(const float4 (&g)[HALF + 1], TexRows &S,
                                             float2 &val2, float2 &wgt2, float &val, float &wgt)
{
    constexpr int NP = HALF + 1;
    const float2 CA4 = bc2(S.cA4), CB2 = bc2(S.cB2), CC = bc2(S.cC);
    float2 Wa = S.W, Ra = S.R;
    float2 Wb = __fmul2_rn(S.W, S.RB), Rb = __fmul2_rn(S.R, CB2);
    if (BOTH) {
        const float2 RB1 = __fmul2_rn(S.RB, CC);
        S.W = __fmul2_rn(Wb, RB1);
        S.R = __fmul2_rn(Rb, CB2);
        S.RB = __fmul2_rn(RB1, CC);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        // gather order: x = (c, r+1), y = (c+1, r+1), z = (c+1, r), w = (c, r)
        if (i < HALF) {
            val2 = __ffma2_rn(make_float2(g[i].w, g[i].z), Wa, val2);
            wgt2 = __fadd2_rn(wgt2, Wa);
            Wa = __fmul2_rn(Wa, Ra);
            if (i + 1 < HALF) Ra = __fmul2_rn(Ra, CA4);
            if (BOTH) {
                val2 = __ffma2_rn(make_float2(g[i].x, g[i].y), Wb, val2);
                wgt2 = __fadd2_rn(wgt2, Wb);
                Wb = __fmul2_rn(Wb, Rb);
                if (i + 1 < HALF) Rb = __fmul2_rn(Rb, CA4);
            }
        } else {                                             // last pair: column N is outside the square
            val = fmaf(g[i].w, Wa.x, val);
            wgt += Wa.x;
            if (BOTH) {
                val = fmaf(g[i].x, Wb.x, val);
                wgt += Wb.x;
            }
        }
    }
}

// The texture unit answers after several hundred cycles and a warp consumes its texels in order: what
// hides the latency is the number of fetches in flight.  ncu on the first form (fetch, use, fetch, ...,
// 2-3 in flight): 56 % of the stall samples on the first FFMA2 after a fetch.  Here the fetches of the
// WHOLE footprint (HALF <= 2: 4 or 9 TLD4, 16 or 36 registers) or of a whole row pair (HALF + 1 TLD4) are
// issued before the first weight is formed.
template <int HALF>
__device__ __forceinline__ void gauss_taps_tex(cudaTextureObject_t tex, int xs, int ys, float w0, float r0, float rb,
                                               float cA, float cB, float cC, float &val, float &wgt)
{
    constexpr int NP = HALF + 1;
    const float fx0 = (float)(xs + 1);
    float fy = (float)(ys + 1);
    float2 val2 = make_float2(0.0f, 0.0f), wgt2 = make_float2(0.0f, 0.0f);
    TexRows S;
    if constexpr (HALF <= 2) {
        float4 g[NP][NP];
#pragma unroll
        for (int j = 0; j < NP; ++j)
#pragma unroll
            for (int i = 0; i < NP; ++i)
                g[j][i] = tex2Dgather<float4>(tex, fx0 + (float)(2 * i), fy + (float)(2 * j), 0);
        const float cA2 = cA * cA, r02 = r0 * r0;
        S.W = __fmul2_rn(bc2(w0), make_float2(1.0f, r0));
        S.R = __fmul2_rn(bc2(r02), make_float2(cA, cA * cA2));
        S.RB = __fmul2_rn(bc2(rb), make_float2(1.0f, cB));
        S.cA4 = cA2 * cA2; S.cB2 = cB * cB; S.cC = cC;
#pragma unroll
        for (int j = 0; j < HALF; ++j) tex_row_pair<HALF, true>(g[j], S, val2, wgt2, val, wgt);
        tex_row_pair<HALF, false>(g[HALF], S, val2, wgt2, val, wgt);
    } else {
        float4 g[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) g[i] = tex2Dgather<float4>(tex, fx0 + (float)(2 * i), fy, 0);
        const float cA2 = cA * cA, r02 = r0 * r0;
        S.W = __fmul2_rn(bc2(w0), make_float2(1.0f, r0));
        S.R = __fmul2_rn(bc2(r02), make_float2(cA, cA * cA2));
        S.RB = __fmul2_rn(bc2(rb), make_float2(1.0f, cB));
        S.cA4 = cA2 * cA2; S.cB2 = cB * cB; S.cC = cC;
#pragma unroll 1
        for (int j = 0; j < HALF; ++j) {
            fy += 2.0f;
            tex_row_pair<HALF, true>(g, S, val2, wgt2, val, wgt);
#pragma unroll
            for (int i = 0; i < NP; ++i) g[i] = tex2Dgather<float4>(tex, fx0 + (float)(2 * i), fy, 0);
        }
        tex_row_pair<HALF, false>(g, S, val2, wgt2, val, wgt);
    }
    val += val2.x + val2.y;
    wgt += wgt2.x + wgt2.y;
}

}  // namespace tfm
