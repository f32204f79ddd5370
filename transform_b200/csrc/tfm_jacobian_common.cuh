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

}  // namespace tfm
