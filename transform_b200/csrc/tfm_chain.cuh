// tfm_chain.cuh -- in-register evaluation of a flattened, direction-resolved transform chain.
//
// Replaces the whole-array temporaries of Composition._forward/_reverse (core.py:1551-1559),
// Linear (basic.py:141-176), Radial (basic.py:670-711) and the linear/TAN/CAR part of WCS
// (core.py:1719-1747): one (x,y) pair lives in two fp64 registers from pixel index to tap.
//
// Two arithmetic policies:
//   Exact  every operation is an explicitly rounded intrinsic in the reference's order, so nvcc
//          cannot contract mul+add into fma; the only fma is the gemv one NumPy itself performs
//          (see include/tfm_b200.h, TFM_LIN_MAT_FORDER).  Affine chains are bit-identical to the
//          reference; transcendental ops differ from glibc by <= 2 ulp (CUDA libdevice).
//   Fast   same formulas, contraction allowed (the host has already pre-composed affine runs).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tfm_b200.h"

namespace tfm {

struct ChainDev {
    int32_t n;
    int32_t pad;
    tfm_op ops[TFM_MAX_OPS];
};

struct Exact {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static constexpr bool exact = true;
};

struct Fast {
    static __device__ __forceinline__ double add(double a, double b) { return a + b; }
    static __device__ __forceinline__ double sub(double a, double b) { return a - b; }
    static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
    static __device__ __forceinline__ double div(double a, double b) { return a / b; }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
    static constexpr bool exact = false;
};

// np.remainder for floats (npy_divmod): result takes the sign of the divisor; a zero result is
// copysign(0, b).  Used for `out[...,0] %= 2*np.pi` (basic.py:683-684).
__device__ __forceinline__ double np_remainder(double a, double b)
{
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0.0) != (m < 0.0)) m = __dadd_rn(m, b);
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

template <class A>
__device__ __forceinline__ void matvec2(const double *m, bool f_order, double &x, double &y)
{
    double ox, oy;
    if (f_order) {          // transposed view in the reference: fma(m[i][1], y, fl(m[i][0]*x))
        ox = A::fma(m[1], y, A::mul(m[0], x));
        oy = A::fma(m[3], y, A::mul(m[2], x));
    } else {                // C-contiguous: fma(m[i][0], x, fl(m[i][1]*y))
        ox = A::fma(m[0], x, A::mul(m[1], y));
        oy = A::fma(m[2], x, A::mul(m[3], y));
    }
    x = ox;
    y = oy;
}

template <class A>
__device__ __forceinline__ void op_linear(const tfm_op &op, double &x, double &y)
{
    const int f = op.flags;
    if (op.dir == TFM_FWD) {                                   // basic.py:141-158
        if (f & TFM_LIN_HAS_PRE) { x = A::add(x, op.p[4]); y = A::add(y, op.p[5]); }
        if (f & TFM_LIN_HAS_MATRIX) matvec2<A>(op.p, (f & TFM_LIN_MAT_FORDER) != 0, x, y);
        if (f & TFM_LIN_HAS_POST) { x = A::add(x, op.p[6]); y = A::add(y, op.p[7]); }
    } else {                                                   // basic.py:160-176
        if (f & TFM_LIN_HAS_POST) { x = A::sub(x, op.p[6]); y = A::sub(y, op.p[7]); }
        if (f & TFM_LIN_HAS_MATRIX) matvec2<A>(op.p, (f & TFM_LIN_MAT_FORDER) != 0, x, y);
        if (f & TFM_LIN_HAS_PRE) { x = A::sub(x, op.p[4]); y = A::sub(y, op.p[5]); }
    }
}

template <class A>
__device__ __forceinline__ void op_radial(const tfm_op &op, double &x, double &y)
{
    const int f = op.flags;
    if (op.dir == TFM_FWD) {                                   // basic.py:670-690
        if (f & TFM_RAD_ORIGIN_NONZERO) { x = A::sub(x, op.p[0]); y = A::sub(y, op.p[1]); }
        double th = atan2((f & TFM_RAD_CCW) ? y : -y, x);
        th = A::div(th, op.p[3]);
        if (f & TFM_RAD_POS_ONLY) th = np_remainder(th, 6.283185307179586);
        double r2 = A::add(A::mul(x, x), A::mul(y, y));
        double r;
        if (f & TFM_RAD_R0_TRUTHY)
            r = A::mul(0.5, log(A::div(r2, A::mul(op.p[2], op.p[2]))));
        else
            r = __dsqrt_rn(r2);
        x = th;
        y = r;
    } else {                                                   // basic.py:692-711
        double d0 = A::mul(x, op.p[3]);
        double s, c;
        sincos(d0, &s, &c);
        if (!(f & TFM_RAD_CCW)) s = -s;
        double r = (f & TFM_RAD_R0_NOT_NONE) ? A::mul(op.p[2], exp(y)) : y;
        x = A::mul(c, r);
        y = A::mul(s, r);
        if (f & TFM_RAD_ORIGIN_NONZERO) { x = A::add(x, op.p[0]); y = A::add(y, op.p[1]); }
    }
}

// ---- per-axis pincushion family (basic.py:846-1462); formulas and constants: include/tfm_b200.h ----
template <class A>
__device__ __forceinline__ double pin_root3(double x)          // (1 - 2*(x<0)) * abs(x)**(1/3)
{
    const double r = pow(fabs(x), 1.0 / 3.0);
    return (x < 0.0) ? -r : r;
}

template <class A>
__device__ __forceinline__ double pin_axis(const tfm_op &op, const double *c, double d)
{
    const int kind = op.flags & 3;
    const bool shifted = (op.flags & TFM_PIN_SHIFTED) != 0;
    const double o = c[0];
    double u = shifted ? A::sub(d, o) : d;
    if (kind == TFM_PIN_QUAD) {
        if (op.dir == TFM_FWD) {
            u = A::add(u, A::div(A::mul(c[1], A::mul(u, fabs(u))), c[2]));
            u = A::div(u, c[3]);
        } else {
            double t = __dsqrt_rn(A::add(1.0, A::mul(A::mul(c[1], fabs(u)), c[2])));
            t = A::add(-1.0, t);
            t = A::div(A::mul(t, c[3]), c[4]);
            u = (d < (shifted ? o : 0.0)) ? -t : t;
        }
    } else if (kind == TFM_PIN_CUBIC) {
        if (op.dir == TFM_FWD) {
            const double q = A::div(u, c[2]);
            u = A::add(u, A::mul(A::mul(A::mul(c[1], u), q), q));
            u = A::div(u, c[3]);
        } else {
            u = A::mul(u, c[1]);
            const double al = A::mul(c[2], -u);
            const double in = __dsqrt_rn(A::add(A::mul(al, al), c[3]));
            const double rn = pin_root3<A>(A::mul(0.5, A::sub(al, in)));
            const double rp = pin_root3<A>(A::mul(0.5, A::add(al, in)));
            u = A::mul(c[4], A::add(rp, rn));
        }
    } else {
        double a = u;
        const int degree = op.reserved;
        for (int k = 2; k <= degree; ++k) {
            const double pk = (k == 2) ? A::mul(u, u) : pow(u, (double)k);
            a = A::add(a, A::div(A::mul(c[1], pk), c[5 + (k - 2)]));
        }
        u = A::mul(c[2], a);
    }
    return shifted ? A::add(u, o) : u;
}

template <class A>
__device__ __forceinline__ void op_pin(const tfm_op &op, double &x, double &y)
{
    x = pin_axis<A>(op, op.p, x);
    if (!(op.flags & TFM_PIN_DIM1)) y = pin_axis<A>(op, op.p + 12, y);
    if (op.flags & TFM_PIN_TRUNC) {                             // float64 -> int64 write-back of the reference
        x = (fabs(x) < 9.2e18) ? trunc(x) : -9223372036854775808.0;
        if (!(op.flags & TFM_PIN_DIM1)) y = (fabs(y) < 9.2e18) ? trunc(y) : -9223372036854775808.0;
    }
}

// ---- FITS WCS (Greisen & Calabretta 2002 paper I; Calabretta & Greisen 2002 paper II) ----
// PARITY UNPINNED beyond the linear part: the reference delegates to astropy/wcslib
// (core.py:1727, 1742), which is absent and unpinned; restated from the papers.
#define TFM_D2R 0.017453292519943295
#define TFM_R2D 57.29577951308232

// Spherical part in unit-vector form.  m = (cos t cos p, cos t sin p, sin t) for native (p, t);
// l = R m is the celestial unit vector (R built on the host from CRVAL / LONPOLE / LATPOLE).
//   TAN (gnomonic): intermediate (x, y) = R2D (m_y, -m_x) / m_z  <=>  m ~ (-y, x, R2D)   [eq. 54-55]
//   CAR           : (x, y) = (p, t) in degrees                                          [eq. 83-84]
// so TAN needs no trigonometry at all and CAR needs one sincos pair or one atan2 pair.
// intermediate world coordinates (degrees) -> native direction m (any positive scale); NaN when the
// point is outside the projection's image of the sphere
// LVL 1 kernels know TAN and CAR only (the projections of the benchmark configurations); the full
// table is compiled into the LVL 2 instantiations -- inlined everywhere it cost the spherical kernel
// variants 16 registers and the TAN -> CAR batch 11 %; out of line it cost stack traffic instead.
__device__ __forceinline__ void wcs_deproject_all(int proj, double ix, double iy, double &mx, double &my, double &mz);
__device__ __forceinline__ void wcs_project_all(int proj, double mx, double my, double mz, double &ix, double &iy);

template <int LVL>
__device__ __forceinline__ void wcs_deproject(int proj, double ix, double iy, double &mx, double &my, double &mz)
{
    if (LVL >= 2) { wcs_deproject_all(proj, ix, iy, mx, my, mz); return; }
    if (proj == TFM_WCS_PROJ_TAN) { mx = -iy; my = ix; mz = TFM_R2D; return; }
    double sp, cp, st, ct;
    sincos(ix * TFM_D2R, &sp, &cp);
    sincos(iy * TFM_D2R, &st, &ct);
    mx = ct * cp; my = ct * sp; mz = st;
}

template <int LVL>
__device__ __forceinline__ void wcs_project(int proj, double mx, double my, double mz, double &ix, double &iy)
{
    if (LVL >= 2) { wcs_project_all(proj, mx, my, mz, ix, iy); return; }
    if (proj == TFM_WCS_PROJ_TAN) {
        const double k = (mz > 0.0) ? TFM_R2D / mz : nan("");      // theta <= 0: the far side of the sky
        ix = my * k; iy = -mx * k;
        return;
    }
    ix = atan2(my, mx) * TFM_R2D;
    iy = atan2(mz, sqrt(mx * mx + my * my)) * TFM_R2D;
}

__device__ __forceinline__ void wcs_deproject_all(int proj, double ix, double iy, double &mx, double &my, double &mz)
{
    const double R0 = TFM_R2D;
    switch (proj) {
    case TFM_WCS_PROJ_TAN: mx = -iy; my = ix; mz = R0; return;
    case TFM_WCS_PROJ_STG: {                    // R = 2 r0 tan((90-theta)/2)
        const double t2 = (ix * ix + iy * iy) / (4.0 * R0 * R0);
        mx = -iy / R0; my = ix / R0; mz = 1.0 - t2;          // scaled by (1 + t^2)
        return;
    }
    case TFM_WCS_PROJ_SIN: {                    // R = r0 cos(theta)
        const double u = ix / R0, v = iy / R0, r2 = u * u + v * v;
        mx = -v; my = u; mz = (r2 <= 1.0) ? sqrt(1.0 - r2) : nan("");
        return;
    }
    case TFM_WCS_PROJ_ARC: {                    // R = 90 - theta
        const double R = sqrt(ix * ix + iy * iy), g = R * TFM_D2R;
        double sg, cg;
        sincos(g, &sg, &cg);
        const double k = (R > 0.0) ? sg / R : TFM_D2R;
        mx = -iy * k; my = ix * k; mz = (R <= 180.0) ? cg : nan("");
        return;
    }
    case TFM_WCS_PROJ_ZEA: {                    // R = 2 r0 sin((90-theta)/2)
        const double r2 = (ix * ix + iy * iy) / (R0 * R0);
        mz = (r2 <= 4.0) ? 1.0 - 0.5 * r2 : nan("");
        const double q = sqrt(fmax(0.5 * (1.0 + mz), 0.0)) / R0;
        mx = -iy * q; my = ix * q;
        return;
    }
    default: break;
    }
    // cylindrical and pseudo-cylindrical: native (phi, theta) first
    double phi = ix * TFM_D2R, st, ct;
    switch (proj) {
    case TFM_WCS_PROJ_CEA: st = iy / R0; ct = (fabs(st) <= 1.0) ? sqrt(1.0 - st * st) : nan(""); break;
    case TFM_WCS_PROJ_MER: { const double u = iy / R0; st = tanh(u); ct = 1.0 / cosh(u); break; }
    case TFM_WCS_PROJ_SFL:
        sincos(iy * TFM_D2R, &st, &ct);
        phi = (ct > 0.0) ? phi / ct : ((ix == 0.0) ? 0.0 : nan(""));
        if (!(fabs(iy) <= 90.0) || !(fabs(phi) <= 3.141592653589793 * (1.0 + 1e-12))) phi = nan("");
        break;
    case TFM_WCS_PROJ_AIT: {
        const double u = ix / (4.0 * R0), v = iy / (2.0 * R0), z2 = 1.0 - u * u - v * v;
        if (!(z2 >= 0.5 * (1.0 - 1e-12))) { phi = nan(""); st = nan(""); ct = nan(""); break; }
        const double Z = sqrt(z2);
        phi = 2.0 * atan2(2.0 * u * Z, 2.0 * z2 - 1.0);
        st = 2.0 * v * Z;                       // (y / r0) Z
        ct = sqrt(fmax(1.0 - st * st, 0.0));
        break;
    }
    default: sincos(iy * TFM_D2R, &st, &ct); break;              // CAR
    }
    double sp, cp;
    sincos(phi, &sp, &cp);
    mx = ct * cp; my = ct * sp; mz = st;
}

// native UNIT vector m -> intermediate world coordinates (degrees); NaN where the projection diverges
// or the point is on the hidden side
__device__ __forceinline__ void wcs_project_all(int proj, double mx, double my, double mz, double &ix, double &iy)
{
    const double R0 = TFM_R2D;
    double k;
    switch (proj) {
    case TFM_WCS_PROJ_TAN: k = (mz > 0.0) ? R0 / mz : nan(""); ix = my * k; iy = -mx * k; return;
    case TFM_WCS_PROJ_STG: k = (mz > -1.0) ? 2.0 * R0 / (1.0 + mz) : nan(""); ix = my * k; iy = -mx * k; return;
    case TFM_WCS_PROJ_SIN: k = (mz >= 0.0) ? R0 : nan(""); ix = my * k; iy = -mx * k; return;
    case TFM_WCS_PROJ_ARC: {
        const double rho = sqrt(mx * mx + my * my);
        k = (rho > 0.0) ? R0 * atan2(rho, mz) / rho : R0;
        ix = my * k; iy = -mx * k;
        return;
    }
    case TFM_WCS_PROJ_ZEA: k = (mz > -1.0) ? R0 * sqrt(2.0 / (1.0 + mz)) : nan(""); ix = my * k; iy = -mx * k; return;
    default: break;
    }
    const double rho = sqrt(mx * mx + my * my);
    const double phi = atan2(my, mx);                              // radians, (-pi, pi]
    switch (proj) {
    case TFM_WCS_PROJ_CEA: ix = phi * R0; iy = R0 * mz; return;
    case TFM_WCS_PROJ_MER: ix = phi * R0; iy = (fabs(mz) < 1.0) ? R0 * atanh(mz) : nan(""); return;
    case TFM_WCS_PROJ_SFL: ix = phi * R0 * rho; iy = atan2(mz, rho) * R0; return;
    case TFM_WCS_PROJ_AIT: {
        double sh, ch;
        sincos(0.5 * phi, &sh, &ch);
        const double g = sqrt(2.0 / (1.0 + rho * ch));
        ix = 2.0 * R0 * g * rho * sh; iy = R0 * g * mz;
        return;
    }
    default: ix = phi * R0; iy = atan2(mz, rho) * R0; return;      // CAR
    }
}

// ---- slant orthographic projection: SIN with (xi, eta) = (PV2_1, PV2_2), paper II eq. 61-62 ----
// x = r0 [cos t sin p + xi (1 - sin t)], y = -r0 [cos t cos p - eta (1 - sin t)], i.e. in native unit-vector form
// ix = r0 (m_y + xi s), iy = -r0 (m_x - eta s) with s = 1 - m_z; the inverse is the smaller root of
// (xi^2 + eta^2 + 1) s^2 - 2 (X xi + Y eta + 1) s + X^2 + Y^2 = 0 (the solution nearer the native pole).
// Out of line (cold): inlined 8-wide into the texture kernels they took the SPH = 2 instantiation from 70 to
// 200 registers.
struct SlantVec { double x, y, z; };

static __device__ __noinline__ SlantVec sin_slant_deproject_cold(double xi, double eta, double ix, double iy)
{
    const double X = ix / TFM_R2D, Y = iy / TFM_R2D;
    const double a = xi * xi + eta * eta + 1.0, b = X * xi + Y * eta + 1.0, c = X * X + Y * Y;
    const double disc = b * b - a * c;
    const double s = (disc >= 0.0) ? c / (b + sqrt(disc)) : nan("");      // = (b - sqrt(disc)) / a, stable near 0
    return SlantVec{-Y + eta * s, X - xi * s, 1.0 - s};
}

static __device__ __noinline__ SlantVec sin_slant_project_cold(double xi, double eta, double mx, double my, double mz)
{
    const double s = 1.0 - mz;
    const bool visible = mz + xi * my - eta * mx >= 0.0;                  // theta >= -atan(xi sin p - eta cos p)
    return SlantVec{visible ? TFM_R2D * (my + xi * s) : nan(""), visible ? -TFM_R2D * (mx - eta * s) : nan(""), 0.0};
}

__device__ __forceinline__ void wcs_sin_slant_deproject(double xi, double eta, double ix, double iy,
                                                        double &mx, double &my, double &mz)
{
    const SlantVec v = sin_slant_deproject_cold(xi, eta, ix, iy);
    mx = v.x; my = v.y; mz = v.z;
}

__device__ __forceinline__ void wcs_sin_slant_project(double xi, double eta, double mx, double my, double mz,
                                                      double &ix, double &iy)
{
    const SlantVec v = sin_slant_project_cold(xi, eta, mx, my, mz);
    ix = v.x; iy = v.y;
}

template <class A, int LVL>
__device__ __forceinline__ void op_wcs(const tfm_op &op, double &x, double &y)
{
    const int proj = op.flags & 0xff;
    const double *R = op.p + 12;
    if (op.dir == TFM_FWD) {            // pixel (origin 0) -> world: all_pix2world(d, 0), core.py:1727
        double dx = A::sub(A::add(x, 1.0), op.p[0]);       // FITS pixels are 1-based
        double dy = A::sub(A::add(y, 1.0), op.p[1]);
        double ix = A::add(A::mul(op.p[2], dx), A::mul(op.p[3], dy));
        double iy = A::add(A::mul(op.p[4], dx), A::mul(op.p[5], dy));
        if (proj == TFM_WCS_PROJ_NONE) {
            x = A::add(ix, op.p[10]);
            y = A::add(iy, op.p[11]);
            return;
        }
        if constexpr (LVL >= 2)                            // CEA: y = r0 sin(theta) / lambda (eq. 80), PV2_1 in p[21]
            if (proj == TFM_WCS_PROJ_CEA && op.p[21] != 0.0) iy *= op.p[21];
        double mx, my, mz;
        if constexpr (LVL >= 2) {
            if (proj == TFM_WCS_PROJ_SIN && (op.p[21] != 0.0 || op.p[22] != 0.0))
                wcs_sin_slant_deproject(op.p[21], op.p[22], ix, iy, mx, my, mz);
            else
                wcs_deproject<LVL>(proj, ix, iy, mx, my, mz);
        } else {
            wcs_deproject<LVL>(proj, ix, iy, mx, my, mz);  // un-normalised: only angles are taken below
        }
        const double lx = R[0] * mx + R[1] * my + R[2] * mz;
        const double ly = R[3] * mx + R[4] * my + R[5] * mz;
        const double lz = R[6] * mx + R[7] * my + R[8] * mz;
        double lon = atan2(ly, lx) * TFM_R2D;
        const double lat = atan2(lz, sqrt(lx * lx + ly * ly)) * TFM_R2D;
        if (lon < 0.0) lon += 360.0;                       // wcslib normalises into [0, 360)
        x = lon;
        y = lat;
    } else {                            // world -> pixel: all_world2pix(d, 0), core.py:1742
        double ix, iy;
        if (proj == TFM_WCS_PROJ_NONE) {
            ix = A::sub(x, op.p[10]);
            iy = A::sub(y, op.p[11]);
        } else {
            double sl, cl, sb, cb;
            sincos(x * TFM_D2R, &sl, &cl);
            sincos(y * TFM_D2R, &sb, &cb);
            const double lx = cb * cl, ly = cb * sl, lz = sb;
            const double mx = R[0] * lx + R[3] * ly + R[6] * lz;       // R^T l
            const double my = R[1] * lx + R[4] * ly + R[7] * lz;
            const double mz = R[2] * lx + R[5] * ly + R[8] * lz;
            if constexpr (LVL >= 2) {
                if (proj == TFM_WCS_PROJ_SIN && (op.p[21] != 0.0 || op.p[22] != 0.0))
                    wcs_sin_slant_project(op.p[21], op.p[22], mx, my, mz, ix, iy);
                else
                    wcs_project<LVL>(proj, mx, my, mz, ix, iy);
                if (proj == TFM_WCS_PROJ_CEA && op.p[21] != 0.0) iy /= op.p[21];
            } else {
                wcs_project<LVL>(proj, mx, my, mz, ix, iy);    // NaN on the far side of the sky (TAN, SIN)
            }
        }
        double px = A::add(A::mul(op.p[6], ix), A::mul(op.p[7], iy));
        double py = A::add(A::mul(op.p[8], ix), A::mul(op.p[9], iy));
        x = A::sub(A::add(px, op.p[0]), 1.0);
        y = A::sub(A::add(py, op.p[1]), 1.0);
    }
}

// ---- SIP distortion (Shupe et al. 2005): the pix2foc stage of all_pix2world / all_world2pix ----
// PARITY UNPINNED (astropy absent).  f(u, v) = sum_{p+q <= order} C_p_q u^p v^q as a Horner form in u of
// Horner forms in v; the coefficient table is read through the read-only path at warp-uniform addresses.
__device__ __forceinline__ double sip_poly(const double *c, int order, double u, double v)
{
    double s = 0.0;
    for (int p = order; p >= 0; --p) {
        const double *row = c + p * TFM_SIP_DIM;
        double t = 0.0;
        for (int q = order - p; q >= 0; --q) t = fma(t, v, __ldg(row + q));
        s = fma(s, u, t);
    }
    return s;
}

// Cold ops are kept OUT of line: inlined into the N-wide interpreter their loops were unrolled per pixel and took
// the SPH = 2 texture kernel from 70 to 200 registers (cuobjdump --dump-resource-usage); results by value, a
// reference argument would live on the stack.
struct ChainXY { double x, y; };

static __device__ __noinline__ ChainXY sip_cold(const tfm_op &op, double x, double y)
{
    const double *coef = reinterpret_cast<const double *>(__double_as_longlong(op.p[0]));
    const int ao = (int)op.p[1], bo = (int)op.p[2];
    const double *ca = coef, *cb = coef + TFM_SIP_DIM * TFM_SIP_DIM;
    if (op.dir == TFM_FWD) {            // pixel -> focal plane (still in pixel units, origin 0)
        const double u = (x + 1.0) - op.p[3], v = (y + 1.0) - op.p[4];
        x += sip_poly(ca, ao, u, v);
        y += sip_poly(cb, bo, u, v);
    } else {                            // focal plane -> pixel: pix <- pix0 - f(pix), as all_world2pix iterates
        const double x0 = x, y0 = y;
        for (int k = 0; k < TFM_SIP_MAXITER; ++k) {
            const double u = (x + 1.0) - op.p[3], v = (y + 1.0) - op.p[4];
            const double nx = x0 - sip_poly(ca, ao, u, v), ny = y0 - sip_poly(cb, bo, u, v);
            const double dx = nx - x, dy = ny - y;
            x = nx; y = ny;
            if (!(dx * dx + dy * dy >= 1e-24)) break;       // converged (or NaN: nothing left to refine)
        }
    }
    return ChainXY{x, y};
}

__device__ __forceinline__ void op_sip(const tfm_op &op, double &x, double &y)
{
    const ChainXY r = sip_cold(op, x, y);
    x = r.x; y = r.y;
}

// ---- TPV distortion: polynomial in the intermediate world coordinates (xi, eta) and r = hypot(xi, eta) ----
// PARITY UNPINNED (wcslib absent).  One axis: c[0..39] in the order of the convention (include/tfm_b200.h);
// x is the axis' own coordinate, y the other one.  Degree d contributes sum_k c_k x^(d-k) y^k -- accumulated as
// acc <- acc x + c_k y^k -- plus c r^d for odd d.
__device__ __forceinline__ double tpv_poly(const double *c, double x, double y)
{
    const double r = sqrt(x * x + y * y), r2 = r * r;
    double s = __ldg(c), rp = r;
    int m = 1;
#pragma unroll 1                        // rolled: fully unrolled, the 80 coefficient loads took 200 registers
    for (int d = 1; d <= 7; ++d) {
        double acc = __ldg(c + m), yp = 1.0;
#pragma unroll 1
        for (int k = 1; k <= d; ++k) {
            yp *= y;
            acc = fma(acc, x, __ldg(c + m + k) * yp);
        }
        s += acc;
        m += d + 1;
        if (d & 1) {
            s = fma(__ldg(c + m), rp, s);
            rp *= r2;
            ++m;
        }
    }
    return s;
}

static __device__ __noinline__ ChainXY tpv_cold(const tfm_op &op, double x, double y)
{
    const double *c1 = reinterpret_cast<const double *>(__double_as_longlong(op.p[0]));
    const double *c2 = c1 + TFM_TPV_NCOEF;
    if (op.dir == TFM_FWD) {
        const double xi = x, eta = y;
        x = tpv_poly(c1, xi, eta);
        y = tpv_poly(c2, eta, xi);
    } else {                            // quasi-Newton on the constant linear part (p[1..4] = its inverse)
        const double tx = x, ty = y;
        const double ax = tx - __ldg(c1), ay = ty - __ldg(c2);
        double vx = fma(op.p[1], ax, op.p[2] * ay), vy = fma(op.p[3], ax, op.p[4] * ay);
        for (int k = 0; k < TFM_TPV_MAXITER; ++k) {
            const double ex = tpv_poly(c1, vx, vy) - tx, ey = tpv_poly(c2, vy, vx) - ty;
            const double dx = fma(op.p[1], ex, op.p[2] * ey), dy = fma(op.p[3], ex, op.p[4] * ey);
            vx -= dx; vy -= dy;
            if (!(dx * dx + dy * dy >= 1e-26)) break;
        }
        x = vx; y = vy;
    }
    return ChainXY{x, y};
}

__device__ __forceinline__ void op_tpv(const tfm_op &op, double &x, double &y)
{
    const ChainXY r = tpv_cold(op, x, y);
    x = r.x; y = r.y;
}

// Internal opcode (never crosses the C ABI): a WCS pixel->world op directly followed by a WCS
// world->pixel op, both with a spherical projection -- what remap() builds for a WCS-to-WCS
// resampling (core.py:920-923).  The fast-mode host pass (fuse_chain_fast) replaces the pair by one
// op that keeps the celestial direction as a unit vector: no angles, no atan2 -> sincos round trip.
//   p[0..1] CRPIX_A  p[2..5] CD_A  p[6..14] R = R_B^T R_A  p[15..18] CDinv_B  p[19..20] CRPIX_B
//   flags = projA | projB << 8
#define TFM_OP_WCS_PAIR 100

__device__ __forceinline__ void op_wcs_pair(const tfm_op &op, double &x, double &y)
{
    const int pa = op.flags & 0xff, pb = (op.flags >> 8) & 0xff;
    const double dx = (x + 1.0) - op.p[0], dy = (y + 1.0) - op.p[1];
    const double ix = op.p[2] * dx + op.p[3] * dy, iy = op.p[4] * dx + op.p[5] * dy;
    double mx, my, mz;
    if (pa == TFM_WCS_PROJ_TAN) {
        const double inv = rsqrt(ix * ix + iy * iy + TFM_R2D * TFM_R2D);
        mx = -iy * inv; my = ix * inv; mz = TFM_R2D * inv;
    } else {
        double sp, cp, st, ct;
        sincos(ix * TFM_D2R, &sp, &cp);
        sincos(iy * TFM_D2R, &st, &ct);
        mx = ct * cp; my = ct * sp; mz = st;
    }
    const double *R = op.p + 6;
    const double nx = R[0] * mx + R[1] * my + R[2] * mz;
    const double ny = R[3] * mx + R[4] * my + R[5] * mz;
    const double nz = R[6] * mx + R[7] * my + R[8] * mz;
    double ox, oy;
    if (pb == TFM_WCS_PROJ_TAN) {
        if (nz > 0.0) { const double sc = TFM_R2D / nz; ox = ny * sc; oy = -nx * sc; }
        else { ox = nan(""); oy = nan(""); }
    } else {
        ox = atan2(ny, nx) * TFM_R2D;
        oy = atan2(nz, sqrt(nx * nx + ny * ny)) * TFM_R2D;
    }
    x = (op.p[15] * ox + op.p[16] * oy + op.p[19]) - 1.0;
    y = (op.p[17] * ox + op.p[18] * oy + op.p[20]) - 1.0;
}

// ---- per-tile tables for a fused WCS pair whose FIRST projection is CAR (the output frame of a
// TAN -> CAR remap, BASELINE config 4).  Native longitude and latitude of a CAR pixel are AFFINE in the
// pixel indices, phi = D2R (cd00 dx + cd01 dy), theta = D2R (cd10 dx + cd11 dy), so inside a tile their
// cosines and sines follow from four small tables by the angle-addition formulas: 4 x (tile edge) sincos
// per TILE instead of two per PIXEL (fp64 sincos is ~45 instructions; the pair op was 12 % of the HBM
// roofline with them).  The celestial rotation and the second projection's CD^-1 are folded into three
// 3-vectors (u, v, w) = (CDinv_B R)_{x,y} and R_z rows, so a TAN target costs 9 DFMA, one division and
// two DFMA per pixel.
struct WcsCarCoef {            // per-thread constants of the second half (registers)
    double ux, uy, uz, vx, vy, vz, wx, wy, wz, cx, cy;
};

__device__ __forceinline__ void wcs_car_tables(const tfm_op &op, double TX0, double TY0, int ni, int nj,
                                               double2 *aphi, double2 *bphi, double2 *ath, double2 *bth,
                                               int tid, int nthreads)
{
    const double dx0 = (TX0 + 1.0) - op.p[0], dy0 = (TY0 + 1.0) - op.p[1];
    for (int k = tid; k < 2 * (ni + nj); k += nthreads) {
        double s, c;
        if (k < ni) {
            sincos(TFM_D2R * op.p[2] * (double)k, &s, &c);
            aphi[k] = make_double2(c, s);
        } else if (k < ni + nj) {
            const int j = k - ni;
            sincos(TFM_D2R * (op.p[2] * dx0 + op.p[3] * (dy0 + (double)j)), &s, &c);
            bphi[j] = make_double2(c, s);
        } else if (k < 2 * ni + nj) {
            const int i = k - ni - nj;
            sincos(TFM_D2R * op.p[4] * (double)i, &s, &c);
            ath[i] = make_double2(c, s);
        } else {
            const int j = k - 2 * ni - nj;
            sincos(TFM_D2R * (op.p[4] * dx0 + op.p[5] * (dy0 + (double)j)), &s, &c);
            bth[j] = make_double2(c, s);
        }
    }
}

__device__ __forceinline__ WcsCarCoef wcs_car_coef(const tfm_op &op)
{
    // x = (c15 ox + c16 oy + crpix1_B) - 1 with (ox, oy) = (ny, -nx) R2D / nz for a TAN target, n = R m
    const double *R = op.p + 6;
    WcsCarCoef k;
    k.ux = op.p[15] * R[3] - op.p[16] * R[0]; k.uy = op.p[15] * R[4] - op.p[16] * R[1]; k.uz = op.p[15] * R[5] - op.p[16] * R[2];
    k.vx = op.p[17] * R[3] - op.p[18] * R[0]; k.vy = op.p[17] * R[4] - op.p[18] * R[1]; k.vz = op.p[17] * R[5] - op.p[18] * R[2];
    k.wx = R[6]; k.wy = R[7]; k.wz = R[8];
    k.cx = op.p[19] - 1.0; k.cy = op.p[20] - 1.0;
    return k;
}

// pixel (i, j) of the tile whose tables were built for origin (TX0, TY0); target projection TAN
__device__ __forceinline__ void wcs_car_tan_from_tables(const WcsCarCoef &k, int i, int j, const double2 *aphi,
                                                        const double2 *bphi, const double2 *ath, const double2 *bth,
                                                        double &x, double &y)
{
    const double2 a = aphi[i], b = bphi[j], c = ath[i], d = bth[j];
    const double cp = a.x * b.x - a.y * b.y, sp = a.y * b.x + a.x * b.y;
    const double ct = c.x * d.x - c.y * d.y, st = c.y * d.x + c.x * d.y;
    const double mx = ct * cp, my = ct * sp, mz = st;
    const double w = k.wx * mx + k.wy * my + k.wz * mz;
    const double u = k.ux * mx + k.uy * my + k.uz * mz;
    const double v = k.vx * mx + k.vy * my + k.vz * mz;
    const double sc = (w > 0.0) ? TFM_R2D / w : nan("");           // theta <= 0: the far side of the sky
    x = fma(u, sc, k.cx);
    y = fma(v, sc, k.cy);
}

// Axis-aligned CD matrix of the CAR frame (cd01 = cd10 = 0: every FITS header without rotation, BASELINE
// config 4): longitude depends on the COLUMN only, latitude on the ROW only, and with m = (ct cp, ct sp,
// st) each of the three dot products is  ct (k_x cp + k_y sp) + k_z st = fma(ct[row], A[col], B[row]).
// Per tile: one sincos and 6 flops per column and per row; per PIXEL: 3 DFMA, one reciprocal (hardware
// seed + one Newton step: 2^-40, i.e. 4e-9 pixels at 4096), 3 more.  10 fp64 instructions instead of 35.
__device__ __forceinline__ void wcs_car_sep_tables(const tfm_op &op, const WcsCarCoef &k, double TX0, double TY0,
                                                   int ni, int nj, double4 *col, double4 *row, int tid, int nthreads)
{
    const double dx0 = (TX0 + 1.0) - op.p[0], dy0 = (TY0 + 1.0) - op.p[1];
    for (int q = tid; q < ni + nj; q += nthreads) {
        double s, c;
        if (q < ni) {
            sincos(TFM_D2R * op.p[2] * (dx0 + (double)q), &s, &c);
            col[q] = make_double4(k.ux * c + k.uy * s, k.vx * c + k.vy * s, k.wx * c + k.wy * s, 0.0);
        } else {
            const int j = q - ni;
            sincos(TFM_D2R * op.p[5] * (dy0 + (double)j), &s, &c);
            row[j] = make_double4(c, k.uz * s, k.vz * s, k.wz * s);
        }
    }
}

__device__ __forceinline__ void wcs_car_tan_sep(const WcsCarCoef &k, const double4 a, const double4 b, double &x, double &y)
{
    const double w = fma(b.x, a.z, b.w), u = fma(b.x, a.x, b.y), v = fma(b.x, a.y, b.z);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(w));
    r = fma(r, fma(-w, r, 1.0), r);
    const double sc = (w > 0.0) ? TFM_R2D * r : nan("");           // theta <= 0: the far side of the sky
    x = fma(u, sc, k.cx);
    y = fma(v, sc, k.cy);
}

// "Heavy" ops are compiled only into dedicated kernel instantiations (template parameter SPH):
//   0  none of them        1  WCS with linear axes, TAN or CAR (+ the fused WCS pair)
//   2  every WCS projection, the SIP / TPV distortions and the pow()-based pincushion family
static inline int chain_sph_level(const ChainDev &ch)
{
    int lvl = 0;
    for (int i = 0; i < ch.n; ++i) {
        const tfm_op &op = ch.ops[i];
        if (op.kind == TFM_OP_PIN || op.kind == TFM_OP_SIP || op.kind == TFM_OP_TPV) lvl = 2;
        else if (op.kind == TFM_OP_WCS_PAIR) lvl = lvl < 1 ? 1 : lvl;
        else if (op.kind == TFM_OP_WCS) {
            const int need = ((op.flags & 0xff) > TFM_WCS_PROJ_CAR) ? 2 : 1;   // (slant SIN is > CAR)
            lvl = lvl < need ? need : lvl;
        }
    }
    return lvl;
}

// Internal opcode (never crosses the C ABI): [affine..., Radial reverse, affine...] at the START of a
// chain, i.e. a polar / log-polar unwrap seen from the output pixel grid.  Angle and radius argument
// are affine in the pixel indices, theta = tX X + tY Y + t0, rho = a10 X + a11 Y + b1, so inside a
// tile cos/sin(theta) follow from two small tables by the angle-addition formulas (and exp(rho)
// from two tables by exp(a+b) = exp(a) exp(b)): the resample kernels evaluate 2 x (tile edge)
// sincos per TILE instead of one per grid point.  op_polar is the direct form used elsewhere.
//   p[0..3] A1  p[4..5] b1  p[6] ang_coeff  p[7] r0  p[8..9] origin  p[10..13] A2  p[14..15] b2
//   flags: TFM_RAD_CCW, TFM_RAD_R0_NOT_NONE, TFM_POLAR_HAS_A2
#define TFM_OP_POLAR 101
#define TFM_POLAR_HAS_A2 32

__device__ __forceinline__ void polar_finish(const tfm_op &op, double c, double s, double r, double &x, double &y)
{
    if (!(op.flags & TFM_RAD_CCW)) s = -s;
    const double x1 = c * r + op.p[8], y1 = s * r + op.p[9];
    if (op.flags & TFM_POLAR_HAS_A2) {
        x = fma(op.p[10], x1, fma(op.p[11], y1, op.p[14]));
        y = fma(op.p[12], x1, fma(op.p[13], y1, op.p[15]));
    } else {
        x = x1;
        y = y1;
    }
}

__device__ __forceinline__ void op_polar(const tfm_op &op, double &x, double &y)
{
    const double u = fma(op.p[0], x, fma(op.p[1], y, op.p[4]));
    const double v = fma(op.p[2], x, fma(op.p[3], y, op.p[5]));
    double s, c;
    sincos(u * op.p[6], &s, &c);
    const double r = (op.flags & TFM_RAD_R0_NOT_NONE) ? op.p[7] * exp(v) : v;
    polar_finish(op, c, s, r, x, y);
}

// Per-tile tables for TFM_OP_POLAR.  ta[i] = (cos, sin)(i tX), tb[j] = (cos, sin)(theta(X0,Y0) + j tY);
// ea[i] = exp(i a10), eb[j] = r0 exp(rho(X0,Y0) + j a11) for the log-polar form.  Call with all
// threads of the CTA, then __syncthreads().
__device__ __forceinline__ void polar_tables(const tfm_op &op, double X0, double Y0, int ni, int nj,
                                             double2 *ta, double2 *tb, double *ea, double *eb,
                                             int tid, int nthreads)
{
    const double coeff = op.p[6];
    const double tX = op.p[0] * coeff, tY = op.p[1] * coeff;
    const double t0 = fma(op.p[0], X0, fma(op.p[1], Y0, op.p[4])) * coeff;
    const bool lg = (op.flags & TFM_RAD_R0_NOT_NONE) != 0;
    const double rho0 = fma(op.p[2], X0, fma(op.p[3], Y0, op.p[5]));
    for (int k = tid; k < ni + nj; k += nthreads) {
        double s, c;
        if (k < ni) {
            sincos((double)k * tX, &s, &c);
            ta[k] = make_double2(c, s);
            if (lg) ea[k] = exp((double)k * op.p[2]);
        } else {
            const int j = k - ni;
            sincos(fma((double)j, tY, t0), &s, &c);
            tb[j] = make_double2(c, s);
            if (lg) eb[j] = op.p[7] * exp(fma((double)j, op.p[3], rho0));
        }
    }
}

// coordinates of local grid point (i, j) of the tile whose tables were built for origin (X0, Y0)
__device__ __forceinline__ void polar_from_tables(const tfm_op &op, double X0, double Y0, int i, int j,
                                                  const double2 *ta, const double2 *tb, const double *ea,
                                                  const double *eb, double &x, double &y)
{
    const double2 a = ta[i], b = tb[j];
    const double c = a.x * b.x - a.y * b.y, s = a.y * b.x + a.x * b.y;
    double r;
    if (op.flags & TFM_RAD_R0_NOT_NONE) r = ea[i] * eb[j];
    else r = fma(op.p[2], X0 + (double)i, fma(op.p[3], Y0 + (double)j, op.p[5]));
    polar_finish(op, c, s, r, x, y);
}

// Compose LINEAR / IDENTITY ops [i0, i1) into out = A.in + b (fp64, host).  Parameter preparation,
// a handful of flops per call -- not part of the data path.  The fp64-exact mode never pre-fuses.
static inline void compose_linear_run(const tfm_op *ops, int i0, int i1, double *A4, double *b2)
{
    double a00 = 1, a01 = 0, a10 = 0, a11 = 1, b0 = 0, b1 = 0;
    for (int i = i0; i < i1; ++i) {
        const tfm_op &op = ops[i];
        if (op.kind != TFM_OP_LINEAR) continue;
        const bool fwd = (op.dir == TFM_FWD);
        const int f = op.flags;
        double m00 = 1, m01 = 0, m10 = 0, m11 = 1;
        if (f & TFM_LIN_HAS_MATRIX) { m00 = op.p[0]; m01 = op.p[1]; m10 = op.p[2]; m11 = op.p[3]; }
        double pa0 = 0, pa1 = 0, pb0 = 0, pb1 = 0;      // out = M.(in + pa) + pb
        if (fwd) {
            if (f & TFM_LIN_HAS_PRE) { pa0 = op.p[4]; pa1 = op.p[5]; }
            if (f & TFM_LIN_HAS_POST) { pb0 = op.p[6]; pb1 = op.p[7]; }
        } else {
            if (f & TFM_LIN_HAS_POST) { pa0 = -op.p[6]; pa1 = -op.p[7]; }
            if (f & TFM_LIN_HAS_PRE) { pb0 = -op.p[4]; pb1 = -op.p[5]; }
        }
        const double c0 = b0 + pa0, c1 = b1 + pa1;
        const double n00 = m00 * a00 + m01 * a10, n01 = m00 * a01 + m01 * a11;
        const double n10 = m10 * a00 + m11 * a10, n11 = m10 * a01 + m11 * a11;
        b0 = m00 * c0 + m01 * c1 + pb0;
        b1 = m10 * c0 + m11 * c1 + pb1;
        a00 = n00; a01 = n01; a10 = n10; a11 = n11;
    }
    A4[0] = a00; A4[1] = a01; A4[2] = a10; A4[3] = a11; b2[0] = b0; b2[1] = b1;
}

static inline bool is_linear_like(const tfm_op &op) { return op.kind == TFM_OP_LINEAR || op.kind == TFM_OP_IDENTITY; }

// Host-side fast-mode peephole pass: copy `ops` into `out`, fusing (a) a leading
// [affine..., Radial reverse, affine...] run into one TFM_OP_POLAR and (b) adjacent spherical WCS pairs.
static inline int fuse_chain_fast(const tfm_op *ops, int n, tfm_op *out)
{
    int m = 0, i = 0;
    {   // (a) polar prefix
        int k = 0;
        while (k < n && is_linear_like(ops[k])) ++k;
        if (k < n && ops[k].kind == TFM_OP_RADIAL && ops[k].dir == TFM_REV) {
            int e = k + 1;
            while (e < n && is_linear_like(ops[e])) ++e;
            tfm_op f;
            f.kind = TFM_OP_POLAR; f.dir = TFM_FWD; f.reserved = 0;
            for (int q = 0; q < TFM_OP_NPARAM; ++q) f.p[q] = 0.0;
            compose_linear_run(ops, 0, k, f.p + 0, f.p + 4);
            const tfm_op &r = ops[k];
            f.p[6] = r.p[3]; f.p[7] = r.p[2]; f.p[8] = r.p[0]; f.p[9] = r.p[1];
            f.flags = r.flags & (TFM_RAD_CCW | TFM_RAD_R0_NOT_NONE);
            if (e > k + 1) {
                compose_linear_run(ops, k + 1, e, f.p + 10, f.p + 14);
                f.flags |= TFM_POLAR_HAS_A2;
            }
            out[m++] = f;
            i = e;
        }
    }
    for (; i < n; ++i) {
        const tfm_op &a = ops[i];
        const auto pairable = [](int f) { return (f & 0xff) == TFM_WCS_PROJ_TAN || (f & 0xff) == TFM_WCS_PROJ_CAR; };
        if (i + 1 < n && a.kind == TFM_OP_WCS && a.dir == TFM_FWD && pairable(a.flags) &&
            ops[i + 1].kind == TFM_OP_WCS && ops[i + 1].dir == TFM_REV && pairable(ops[i + 1].flags)) {
            const tfm_op &b = ops[i + 1];
            tfm_op f;
            f.kind = TFM_OP_WCS_PAIR; f.dir = TFM_FWD; f.reserved = 0;
            f.flags = (a.flags & 0xff) | ((b.flags & 0xff) << 8);
            for (int k = 0; k < TFM_OP_NPARAM; ++k) f.p[k] = 0.0;
            f.p[0] = a.p[0]; f.p[1] = a.p[1];
            for (int k = 0; k < 4; ++k) { f.p[2 + k] = a.p[2 + k]; f.p[15 + k] = b.p[6 + k]; }
            const double *RA = a.p + 12, *RB = b.p + 12;
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    double acc = 0.0;
                    for (int k = 0; k < 3; ++k) acc += RB[k * 3 + r] * RA[k * 3 + c];   // (R_B^T R_A)[r][c]
                    f.p[6 + r * 3 + c] = acc;
                }
            f.p[19] = b.p[0]; f.p[20] = b.p[1];
            out[m++] = f;
            ++i;
        } else {
            out[m++] = a;
        }
    }
    return m;
}

// Host-side validation of the opcodes that cross the C ABI.
static inline bool chain_ops_valid(const tfm_op *ops, int n)
{
    for (int i = 0; i < n; ++i) {
        const tfm_op &op = ops[i];
        if (op.kind < TFM_OP_IDENTITY || op.kind > TFM_OP_TPV) return false;
        if (op.dir != TFM_FWD && op.dir != TFM_REV) return false;
        if (op.kind == TFM_OP_WCS && ((op.flags & 0xff) < 0 || (op.flags & 0xff) > TFM_WCS_PROJ_MAX)) return false;
        if (op.kind == TFM_OP_TPV) {
            long long bits;
            memcpy(&bits, &op.p[0], sizeof bits);
            if (bits == 0) return false;                    // no coefficient table
        }
        if (op.kind == TFM_OP_SIP) {
            long long bits;
            memcpy(&bits, &op.p[0], sizeof bits);
            if (bits == 0) return false;                    // no coefficient table
            if (!(op.p[1] >= 0 && op.p[1] <= TFM_SIP_MAX_ORDER && op.p[2] >= 0 && op.p[2] <= TFM_SIP_MAX_ORDER)) return false;
        }
        if (op.kind == TFM_OP_PIN) {
            const int k = op.flags & 3;
            if (k > TFM_PIN_POLY) return false;
            if (k == TFM_PIN_POLY && (op.dir != TFM_FWD || op.reserved < 0 || op.reserved > TFM_PIN_MAX_DEGREE))
                return false;                               // Poly has no inverse (basic.py:1296)
        }
    }
    return true;
}

// interpND / interpND_grid take a precomputed coordinate array (helpers.pyx:379, 1293)
__device__ __forceinline__ void op_index(const tfm_op &op, double &x, double &y)
{
    const double2 *idx = reinterpret_cast<const double2 *>(__double_as_longlong(op.p[0]));
    const long long nx = (long long)op.p[1], ny = (long long)op.p[2];
    const long long ix = (long long)x, iy = (long long)y;
    if (ix < 0 || ix >= nx || iy < 0 || iy >= ny) {
        x = nan("");
        y = nan("");
        return;
    }
    const double2 v = __ldg(idx + iy * nx + ix);
    x = v.x;
    y = v.y;
}

// SPH: the chain contains spherical WCS ops.  Their code (sincos/atan2 in fp64) is compiled only into
// the kernel instantiations that the host selects for such chains -- inlined into every kernel it
// costs the others registers (measured +7% on the Jacobian kernel), out of line it spills the
// interpreter's state to the stack (measured 2x on the parity-mode linear kernel).
template <class A, int SPH>
__device__ __forceinline__ void chain_eval(const ChainDev &ch, double &x, double &y, int first = 0)
{
    for (int i = first; i < ch.n; ++i) {
        const tfm_op &op = ch.ops[i];
        switch (op.kind) {
        case TFM_OP_LINEAR: op_linear<A>(op, x, y); break;
        case TFM_OP_RADIAL: op_radial<A>(op, x, y); break;
        case TFM_OP_WCS: if constexpr (SPH >= 1) op_wcs<A, SPH>(op, x, y); break;
        case TFM_OP_WCS_PAIR: if constexpr (SPH >= 1) op_wcs_pair(op, x, y); break;
        case TFM_OP_PIN: if constexpr (SPH >= 2) op_pin<A>(op, x, y); break;
        case TFM_OP_SIP: if constexpr (SPH >= 2) op_sip(op, x, y); break;
        case TFM_OP_TPV: if constexpr (SPH >= 2) op_tpv(op, x, y); break;
        case TFM_OP_POLAR: op_polar(op, x, y); break;
        case TFM_OP_INDEX: op_index(op, x, y); break;
        default: break;                                     // identity
        }
    }
}

// ---- N-wide evaluation: the interpreter's dispatch (constant loads, flag tests, branches) is paid
// once per op per THREAD instead of once per op per pixel; the per-pixel work is the arithmetic only.
template <class A, int SPH, int N>
__device__ __forceinline__ void chain_eval_n(const ChainDev &ch, double (&x)[N], double (&y)[N], int first = 0)
{
    for (int i = first; i < ch.n; ++i) {
        const tfm_op &op = ch.ops[i];
        const int f = op.flags;
        switch (op.kind) {
        case TFM_OP_LINEAR: {
            const bool fwd = (op.dir == TFM_FWD);
            // fwd: +pre, M, +post ; rev: -post, Minv, -pre   (basic.py:141-176)
            // x - c == x + (-c) exactly in IEEE arithmetic, so one add form serves both directions
            const double a0 = fwd ? op.p[4] : -op.p[6], a1 = fwd ? op.p[5] : -op.p[7];
            const double b0 = fwd ? op.p[6] : -op.p[4], b1 = fwd ? op.p[7] : -op.p[5];
            const bool has_a = fwd ? (f & TFM_LIN_HAS_PRE) : (f & TFM_LIN_HAS_POST);
            const bool has_b = fwd ? (f & TFM_LIN_HAS_POST) : (f & TFM_LIN_HAS_PRE);
            if (has_a) {
#pragma unroll
                for (int k = 0; k < N; ++k) { x[k] = A::add(x[k], a0); y[k] = A::add(y[k], a1); }
            }
            if (f & TFM_LIN_HAS_MATRIX) {
                const double m0 = op.p[0], m1 = op.p[1], m2 = op.p[2], m3 = op.p[3];
                if (f & TFM_LIN_MAT_FORDER) {
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        const double ox = A::fma(m1, y[k], A::mul(m0, x[k]));
                        const double oy = A::fma(m3, y[k], A::mul(m2, x[k]));
                        x[k] = ox; y[k] = oy;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        const double ox = A::fma(m0, x[k], A::mul(m1, y[k]));
                        const double oy = A::fma(m2, x[k], A::mul(m3, y[k]));
                        x[k] = ox; y[k] = oy;
                    }
                }
            }
            if (has_b) {
#pragma unroll
                for (int k = 0; k < N; ++k) { x[k] = A::add(x[k], b0); y[k] = A::add(y[k], b1); }
            }
            break;
        }
        case TFM_OP_RADIAL:
#pragma unroll
            for (int k = 0; k < N; ++k) op_radial<A>(op, x[k], y[k]);
            break;
        case TFM_OP_WCS:
            if constexpr (SPH >= 1) {
#pragma unroll
                for (int k = 0; k < N; ++k) op_wcs<A, SPH>(op, x[k], y[k]);
            }
            break;
        case TFM_OP_WCS_PAIR:
            if constexpr (SPH >= 1) {
#pragma unroll
                for (int k = 0; k < N; ++k) op_wcs_pair(op, x[k], y[k]);
            }
            break;
        case TFM_OP_PIN:
            if constexpr (SPH >= 2) {
#pragma unroll
                for (int k = 0; k < N; ++k) op_pin<A>(op, x[k], y[k]);
            }
            break;
        case TFM_OP_SIP:
            if constexpr (SPH >= 2) {
#pragma unroll
                for (int k = 0; k < N; ++k) op_sip(op, x[k], y[k]);
            }
            break;
        case TFM_OP_TPV:
            if constexpr (SPH >= 2) {
#pragma unroll
                for (int k = 0; k < N; ++k) op_tpv(op, x[k], y[k]);
            }
            break;
        case TFM_OP_POLAR:
#pragma unroll
            for (int k = 0; k < N; ++k) op_polar(op, x[k], y[k]);
            break;
        case TFM_OP_INDEX:
#pragma unroll
            for (int k = 0; k < N; ++k) op_index(op, x[k], y[k]);
            break;
        default: break;
        }
    }
}

}  // namespace tfm
