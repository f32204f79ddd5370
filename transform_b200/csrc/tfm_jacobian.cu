// tfm_jacobian.cu -- K2: Jacobian-adaptive (DeForest 2004) anti-aliased resampler.
//
// One launch replaces interpND_grid(..., antialias=True) (helpers.pyx:1293-1530):
//   coordinates      self.invert(grid)                      -> chain_eval on a (32+4) x (8+4) halo tile
//   simple_jacobian  helpers.pyx:896-906   [NY-1,NX-1,2,2]  -> cell differences from the smem tile
//   jump_detect      helpers.pyx:1012-1141 (9-/6-lag sort)  -> selection network per cell, smem flags
//   jacobian         helpers.pyx:1225-1247 (grid-centring)  -> 4-cell flagged mean per pixel
//   interpND_jacobian helpers.pyx:1542-1803                 -> svd2x2, padding, footprint loop
// None of the reference's [NY,NX,2], [NY,NX,2,2] float64 temporaries (1 + 2 GB at 8192^2) exists:
// HBM sees the source pixels and the output pixels only.  The footprint loop (25..121+ taps per
// pixel, each an exp) makes this kernel instruction-bound, not HBM-bound -- see DESIGN.md.
//
// Repairs relative to the shipped (non-functional) reference text are those enumerated in
// oracle/oracle_np.py::interp_grid_jacobian (R1-R12).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "tfm_jacobian_common.cuh"

namespace tfm {

constexpr int JTX = 32;                     // output tile width; height JTY is a template parameter
constexpr int PW = JTX + 4;                 // coordinate points per row (halo 2)
constexpr int CW = JTX + 3;                 // cells per row (lower-left point index), halo 2 / 1

struct JacParams {
    const void *src;
    void *dst;
    long long src_pitch, dst_pitch;
    long long src_plane, dst_plane;
    int n_planes;
    int src_h, src_w, src_x0, src_y0, full_h, full_w;
    int dst_h, dst_w;
    long long dst_x0, dst_y0;
    int grid_h, grid_w;                      // full output grid (edge rules)
    int bound_x, bound_y;
    int filter;
    int flags;
    int src_f64, dst_f64;
    double oblur, pad_pow, padval, jump_thresh;
    long long max_reg;
    int *status;
    // constant-Jacobian path (fast mode, Gaussian, the whole chain is affine): everything that depends
    // on J alone is computed once on the host, in fp64
    int lw;                                  // log2 of the warp patch width in the footprint phase
    cudaTextureObject_t tex;                 // fast mode, Gaussian: point-sampled texture over the whole source
    int tex_ok;                              // 0 none; 1 footprints inside the image only; 2 both bounds truncate:
                                             // the texture's border IS the reference's truncated tap
    int aff_const;
    int aff_half;                            // footprint half-width
    double aff[6];                           // source = A . grid + b
    float aff_q[3];                          // Gaussian quadratic form (q00, q01, q11)
    ChainDev chain;
};

// ---- k-th smallest of a handful of doubles (np.sort(...)[k], helpers.pyx:1063-1064, 1109-1110) ----
__device__ __forceinline__ void cswap(double &a, double &b)
{
    const double lo = fmin(a, b), hi = fmax(a, b);
    a = lo;
    b = hi;
}

__device__ __forceinline__ double kth3_of9(double *v)     // element 3 of the sorted 9
{
    // 25-exchange sorting network for 9 inputs
    cswap(v[0], v[1]); cswap(v[3], v[4]); cswap(v[6], v[7]);
    cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[0], v[1]); cswap(v[3], v[4]); cswap(v[6], v[7]);
    cswap(v[0], v[3]); cswap(v[3], v[6]); cswap(v[0], v[3]);
    cswap(v[1], v[4]); cswap(v[4], v[7]); cswap(v[1], v[4]);
    cswap(v[2], v[5]); cswap(v[5], v[8]); cswap(v[2], v[5]);
    cswap(v[1], v[3]); cswap(v[5], v[7]);
    cswap(v[2], v[6]); cswap(v[4], v[6]); cswap(v[2], v[4]);
    cswap(v[2], v[3]); cswap(v[5], v[6]);
    return v[3];
}

__device__ __forceinline__ double kth2_of6(double *v)     // element 2 of the sorted 6
{
    // 12-exchange sorting network for 6 inputs
    cswap(v[0], v[5]); cswap(v[1], v[3]); cswap(v[2], v[4]);
    cswap(v[1], v[2]); cswap(v[3], v[4]);
    cswap(v[0], v[3]); cswap(v[2], v[5]);
    cswap(v[0], v[1]); cswap(v[2], v[3]); cswap(v[4], v[5]);
    cswap(v[1], v[2]); cswap(v[3], v[4]);
    return v[2];
}

template <typename RegT>
__device__ __forceinline__ RegT filt_eval(int filter, RegT xf, RegT yf);

template <>
__device__ __forceinline__ double filt_eval<double>(int filter, double xf, double yf)
{
    const double pi = 3.141592653589793;
    double filt, a, b;
    switch (filter) {
    case TFM_FILT_TENT:                                          // helpers.pyx:1693-1695
        filt = (xf < 1) ? __dsub_rn(1.0, xf) : 0.0;
        filt = __dmul_rn(filt, (yf < 1) ? __dsub_rn(1.0, yf) : 0.0);
        return filt;
    case TFM_FILT_LANCZOS:                                       // helpers.pyx:1697-1706
        if (xf > 3 || yf > 3) return 0.0;
        a = __dmul_rn(pi, xf); b = __ddiv_rn(a, 3.0);
        filt = __ddiv_rn(__dmul_rn(__ddiv_rn(sin(a), a), sin(b)), b);
        a = __dmul_rn(pi, yf); b = __ddiv_rn(a, 3.0);
        return __dmul_rn(filt, __ddiv_rn(__dmul_rn(__ddiv_rn(sin(a), a), sin(b)), b));
    case TFM_FILT_HANNING:                                       // helpers.pyx:1708-1717
        if (xf > 1 || yf > 1) return 0.0;
        b = cos(__dmul_rn(pi, xf)); filt = __dmul_rn(b, b);
        b = cos(__dmul_rn(pi, yf)); return __dmul_rn(filt, __dmul_rn(b, b));
    case TFM_FILT_TUKEY:                                         // helpers.pyx:1719-1732
        if (xf > 1 || yf > 1) return 0.0;
        filt = 1.0;
        if (xf > 0.5) { b = cos(__dmul_rn(__dmul_rn(pi, 2.0), __dsub_rn(xf, 0.5))); filt = __dmul_rn(b, b); }
        if (yf > 0.5) { b = cos(__dmul_rn(__dmul_rn(pi, 2.0), __dsub_rn(yf, 0.5))); filt = __dmul_rn(filt, __dmul_rn(b, b)); }
        return filt;
    case TFM_FILT_HANN_WINDOW:                                   // SVD.pyx:110-111
        a = __dadd_rn(cos(__dmul_rn(xf < 1 ? xf : 1.0, pi)), 1.0);
        b = __dadd_rn(cos(__dmul_rn(yf < 1 ? yf : 1.0, pi)), 1.0);
        return __ddiv_rn(__dmul_rn(a, b), 2.0);
    default:                                                     // Gaussian, helpers.pyx:1734-1735
        return exp(-__dadd_rn(__dmul_rn(xf, xf), __dmul_rn(yf, yf)));
    }
}

template <>
__device__ __forceinline__ float filt_eval<float>(int filter, float xf, float yf)
{
    const float pi = 3.14159265358979f;
    float filt, a, b;
    switch (filter) {
    case TFM_FILT_TENT:
        return ((xf < 1) ? 1.0f - xf : 0.0f) * ((yf < 1) ? 1.0f - yf : 0.0f);
    case TFM_FILT_LANCZOS:
        if (xf > 3 || yf > 3) return 0.0f;
        a = pi * xf; b = a / 3.0f; filt = sinf(a) / a * sinf(b) / b;
        a = pi * yf; b = a / 3.0f; return filt * (sinf(a) / a * sinf(b) / b);
    // arguments of the cosines are within [0, pi]: the MUFU approximation (abs. error 2^-21) is enough
    case TFM_FILT_HANNING:
        if (xf > 1 || yf > 1) return 0.0f;
        b = __cosf(pi * xf); filt = b * b;
        b = __cosf(pi * yf); return filt * b * b;
    case TFM_FILT_TUKEY:
        if (xf > 1 || yf > 1) return 0.0f;
        filt = 1.0f;
        if (xf > 0.5f) { b = __cosf(pi * 2.0f * (xf - 0.5f)); filt = b * b; }
        if (yf > 0.5f) { b = __cosf(pi * 2.0f * (yf - 0.5f)); filt *= b * b; }
        return filt;
    case TFM_FILT_HANN_WINDOW:
        return 0.5f * (__cosf(pi * fminf(xf, 1.0f)) + 1.0f) * (__cosf(pi * fminf(yf, 1.0f)) + 1.0f);
    default:
        return __expf(-(xf * xf + yf * yf));
    }
}

// Exact jump flag of one cell (helpers.pyx:1045-1139): 3x3 lags inside the cell grid, 2x3 / 3x2 on a
// grid edge (the edge line and the next one), corners untouched (1).  Cold path: reached only for
// cells near a jump, on grid edges, or when the fp32 quick-accept cannot decide.
__device__ __noinline__ unsigned char jump_flag_exact(const double *m2, int cw, int ci, int cj,
                                                      int gcx, int gcy, int NCX, int NCY, double thresh)
{
    const bool ix_in = (gcx >= 1 && gcx <= NCX - 2), iy_in = (gcy >= 1 && gcy <= NCY - 2);
    const double me = m2[cj * cw + ci];
    double v[9];
    if (ix_in && iy_in) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) v[a * 3 + b] = m2[(cj - 1 + a) * cw + ci - 1 + b];
        return me <= __dmul_rn(kth3_of9(v), thresh);
    }
    if (ix_in && NCY >= 2) {                       // bottom / top row of cells: 2 x 3 lags
        const int r0 = (gcy == 0) ? cj : cj - 1;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) v[a * 3 + b] = m2[(r0 + a) * cw + ci - 1 + b];
        return me <= __dmul_rn(kth2_of6(v), thresh);
    }
    if (iy_in && NCX >= 2) {                       // left / right column of cells: 3 x 2 lags
        const int c0 = (gcx == 0) ? ci : ci - 1;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) v[a * 2 + b] = m2[(cj - 1 + a) * cw + c0 + b];
        return me <= __dmul_rn(kth2_of6(v), thresh);
    }
    return 1;                                      // corners (helpers.pyx:986-989)
}

// ---- fast-mode (fp32) pieces -------------------------------------------------------------------
// padded inverse Jacobian in fp32: same closed form as svd2x2_dev/svc2x2_dev, MUFU-based
// transcendentals.  Returns the three coefficients of the quadratic form q(a,b) = A a^2 + 2B ab + C b^2
// with q = xf^2 + yf^2, (xf,yf) = J.(a,b)/oblur, pre-multiplied by log2(e) so that the Gaussian
// weight exp(-q) is exp2(-q'), plus J itself for the non-Gaussian filters.
__device__ __forceinline__ float pad_inv_sv(float s, float padval, float pad_pow)
{
    // (padval + s^p)^(-1/p)   (helpers.pyx:1661-1662)
    // s = 0 -> log2 = -inf -> 0; above 2^60 the sum IS the power (and the power may overflow fp32)
    const float e = pad_pow * __log2f(s);
    const float lg = (e > 60.0f) ? e : __log2f(padval + exp2f(e));
    return exp2f(-lg / pad_pow);
}

__device__ __forceinline__ void padded_inverse_f32(const double *Ji, float padval, float pad_pow,
                                                   float *J, float &smin)
{
    const float a = (float)Ji[0], b = (float)Ji[1], c = (float)Ji[2], d = (float)Ji[3];
    const float theta = 0.5f * atan2f(2.0f * (a * c + b * d), (a * a + b * b) - (c * c + d * d));
    const float phi = 0.5f * atan2f(2.0f * (a * b + c * d), (a * a + c * c) - (b * b + d * d));
    float St, Ct, Sp, Cp;
    sincosf(theta, &St, &Ct);
    sincosf(phi, &Sp, &Cp);
    const float cs00 = (Ct * a + St * c) * Cp + (Ct * b + St * d) * Sp;
    const float cs11 = (St * a - Ct * c) * Sp - (St * b - Ct * d) * Cp;
    const float s0 = pad_inv_sv(fabsf(cs00), padval, pad_pow);
    const float s1 = pad_inv_sv(fabsf(cs11), padval, pad_pow);
    const float c00 = cs00 >= 0 ? 1.0f : -1.0f, c11 = cs11 >= 0 ? 1.0f : -1.0f;
    // J = V diag(s') U^T with V = W diag(sign), U = R(theta)   (svc2x2_fast(V, s, U, J), 1663)
    const float V00 = Cp * c00, V01 = -Sp * c11, V10 = Sp * c00, V11 = Cp * c11;
    const float a00 = V00 * s0, a10 = V10 * s0, a01 = V01 * s1, a11 = V11 * s1;
    J[0] = a00 * Ct + a01 * (-St);
    J[1] = a00 * St + a01 * Ct;
    J[2] = a10 * Ct + a11 * (-St);
    J[3] = a10 * St + a11 * Ct;
    smin = fminf(s0, s1);
}

// (2*HALF+1)^2 Gaussian taps.  Weights by recurrence: along a row w(a+1) = w(a) r(a),
// r(a+1) = r(a) cA; from row to row w0 *= rb, rb *= cC, r0 *= cB.  Rows are processed in PAIRS with
// Blackwell's packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2, sm_100+): lane .x carries row j, lane
// .y row j+1, so a pair of taps costs 2 loads + 4 packed instructions instead of 2 + 8.  The odd last
// row runs in scalar form.
template <int HALF, typename SrcT>
__device__ __forceinline__ void gauss_taps(const SrcT *p, long long pe, float w0, float r0, float rb,
                                           float cA, float cB, float cC, float &val, float &wgt)
{
    constexpr int N = 2 * HALF + 1;
    const float2 cA2 = make_float2(cA, cA);
    const float cB2s = cB * cB;                        // two-row steps of the row-to-row recurrences
    const float cC2s = cC * cC;
    float2 val2 = make_float2(0.0f, 0.0f), wgt2 = make_float2(0.0f, 0.0f);
#pragma unroll 1
    for (int j = 0; j + 1 < N; j += 2) {
        // rows j and j+1: first weights w0, w0*rb; first ratios r0, r0*cB
        float2 w = make_float2(w0, w0 * rb);
        float2 r = make_float2(r0, r0 * cB);
        const SrcT *q = p + pe;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const float2 v = make_float2((float)__ldg(p + k), (float)__ldg(q + k));
            val2 = __ffma2_rn(v, w, val2);
            wgt2 = __fadd2_rn(wgt2, w);
            w = __fmul2_rn(w, r);
            r = __fmul2_rn(r, cA2);
        }
        p += 2 * pe;
        // advance two rows: w0 <- w0 rb (rb cC) ; rb <- rb cC^2 ; r0 <- r0 cB^2
        w0 *= rb * (rb * cC);
        rb *= cC2s;
        r0 *= cB2s;
    }
    {   // last (odd) row
        float w = w0, r = r0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            val = fmaf((float)__ldg(p + k), w, val);
            wgt += w;
            w *= r;
            r *= cA;
        }
    }
    val += val2.x + val2.y;
    wgt += wgt2.x + wgt2.y;
}

#ifndef TFM_K2_MINBLOCKS
#define TFM_K2_MINBLOCKS 4
#endif
// A: chain arithmetic; RegT: footprint arithmetic (double = parity mode, float = fast mode)
template <class A, typename SrcT, typename DstT, typename RegT, int JTY, int SPH>
__global__ void __launch_bounds__(JTX * 8, TFM_K2_MINBLOCKS)
k_jacobian(const __grid_constant__ JacParams P)
{
    // 256 threads (32 x 8) cover a 32 x JTY tile, JTY in {8, 16, 32}: thread (tx,ty) owns pixels
    // (tx, ty + 8j).  Taller tiles amortise the 2-point coordinate halo: evaluations per pixel are
    // (36 x 12)/256 = 1.69 at JTY = 8, 1.41 at 16, 1.27 at 32.
    constexpr bool FASTI = (sizeof(RegT) == 4);
    constexpr int NTH = JTX * 8;
    constexpr int PH = JTY + 4, CH = JTY + 3;
    __shared__ double2 s_p[PH][PW];                   // inverse-mapped (x,y) of grid points, interleaved
    __shared__ double s_m2[CH][CW];                   // Jmag2 per cell (helpers.pyx:1020)
    __shared__ float s_m2f[CH][CW];                   // the same, rounded to fp32 (quick-accept test)
    __shared__ unsigned char s_flag[CH][CW];          // jump flag per cell (1 = ok)

    const int tid = threadIdx.y * JTX + threadIdx.x;
    const char *const src_base = static_cast<const char *>(P.src) + (long long)blockIdx.z * P.src_plane;
    char *const dst_base = static_cast<char *>(P.dst) + (long long)blockIdx.z * P.dst_plane;
    // global grid coordinates of the tile origin; the host guarantees |origin| + extent < 2^31
    const int GX0 = (int)P.dst_x0 + (int)blockIdx.x * JTX;
    const int GY0 = (int)P.dst_y0 + (int)blockIdx.y * JTY;
    const int GW = P.grid_w, GH = P.grid_h;
    // smem window origin: 2 points of halo; when the tile's first column (row) is the LAST grid
    // column (row), the edge-copy clamp below reaches one further back, so slide the window by 1
    // (such a tile has a single valid column (row), the far side of the window is unused).
    const int shx = (GX0 == GW - 1) ? 1 : 0, shy = (GY0 == GH - 1) ? 1 : 0;
    const int WX0 = GX0 - 2 - shx, WY0 = GY0 - 2 - shy;

    // ---- 1. coordinates of the halo tile (points outside the grid are never used) ----
    __shared__ double2 s_ta[PW], s_tb[PH];            // polar chains: per-tile angle tables
    __shared__ double s_ea[PW], s_eb[PH];             //               and radius (log-polar) tables
    const bool polar = FASTI && P.chain.n > 0 && P.chain.ops[0].kind == TFM_OP_POLAR;
    // Analytic path (fast mode, the chain is ONE fused polar op): coordinates AND Jacobian of a pixel
    // follow in closed form from the per-tile angle / radius tables -- no halo tile, no cell
    // differences, no flag pass, one barrier.  The reference's Jacobian is a +-1-pixel central
    // difference; against the derivative it is off by O(step^2) of the angle (and log-radius) step,
    // so the path is taken only when those steps are < 1e-2 (relative error < 1e-4 / 6), and only when
    // the closed form of |J|^2 = g2 + r^2 t2 (x r^2 for log-polar) proves that no cell of the tile can
    // be flagged as a jump: max / min over the tile (corners; r = 0 inside counts as min) times the
    // squared condition number of the trailing affine map stays below the threshold.
    bool analytic = false;
    if (polar && P.chain.n == 1 && !(P.flags & 0x100)) {          // 0x100: tuning bit 512, force the difference path
        const tfm_op &op0 = P.chain.ops[0];
        const double coeff = op0.p[6];
        const double t2 = (op0.p[0] * op0.p[0] + op0.p[1] * op0.p[1]) * coeff * coeff;
        const double g2 = op0.p[2] * op0.p[2] + op0.p[3] * op0.p[3];
        const bool lg = (op0.flags & TFM_RAD_R0_NOT_NONE) != 0;
        analytic = (t2 < 1.0e-4) && (!lg || g2 < 1.0e-4);
        if (analytic && P.jump_thresh != 0.0) {
            const double xa = (double)(GX0 - 2), xb = (double)(GX0 + JTX + 1);
            const double ya = (double)(GY0 - 2), yb = (double)(GY0 + JTY + 1);
            const double v0 = fma(op0.p[2], xa, fma(op0.p[3], ya, op0.p[5])), v1 = fma(op0.p[2], xb, fma(op0.p[3], ya, op0.p[5]));
            const double v2 = fma(op0.p[2], xa, fma(op0.p[3], yb, op0.p[5])), v3 = fma(op0.p[2], xb, fma(op0.p[3], yb, op0.p[5]));
            const double vmin = fmin(fmin(v0, v1), fmin(v2, v3)), vmax = fmax(fmax(v0, v1), fmax(v2, v3));
            double ratio;
            if (lg) {
                ratio = exp(2.0 * (vmax - vmin));
            } else {
                const double hi = fmax(vmin * vmin, vmax * vmax);
                const double lo = (vmin <= 0.0 && vmax >= 0.0) ? 0.0 : fmin(vmin * vmin, vmax * vmax);
                ratio = (g2 + t2 * hi) / (g2 + t2 * lo);
            }
            if (op0.flags & TFM_POLAR_HAS_A2) {
                const double E = op0.p[10] * op0.p[10] + op0.p[11] * op0.p[11] + op0.p[12] * op0.p[12] + op0.p[13] * op0.p[13];
                const double det = op0.p[10] * op0.p[13] - op0.p[11] * op0.p[12];
                const double q = sqrt(fmax(E * E - 4.0 * det * det, 0.0));
                ratio *= (E + q) / (E - q);
            }
            analytic = (P.jump_thresh > 0.0) && (ratio * 1.01 <= P.jump_thresh);      // NaN compares false
        }
    }
    const bool affc = FASTI && P.aff_const;
    bool tile_clean = true;
    if (affc) {
        // nothing to prepare: coordinates are affine in the grid indices, the Jacobian is a constant
    } else if (analytic) {
        // tables for grid points (GX0-1 .. GX0+JTX) x (GY0-1 .. GY0+JTY): the tile plus the one-point
        // reach of the edge-copy clamp
        polar_tables(P.chain.ops[0], (double)(GX0 - 1), (double)(GY0 - 1), JTX + 2, JTY + 2, s_ta, s_tb, s_ea, s_eb, tid, NTH);
        __syncthreads();
    } else {
    if (polar) {
        // [affine, Radial reverse, affine] prefix: angle and radius are affine in the grid indices, so
        // PW + PH sincos per tile (angle addition) replace PW x PH of them
        const tfm_op &op0 = P.chain.ops[0];
        polar_tables(op0, (double)WX0, (double)WY0, PW, PH, s_ta, s_tb, s_ea, s_eb, tid, NTH);
        __syncthreads();
        for (int i = tid; i < PW * PH; i += NTH) {
            const int pj = i / PW, pi = i - pj * PW;
            double x, y;
            polar_from_tables(op0, (double)WX0, (double)WY0, pi, pj, s_ta, s_tb, s_ea, s_eb, x, y);
            chain_eval<A, SPH>(P.chain, x, y, 1);
            s_p[pj][pi] = make_double2(x, y);
        }
    } else {
        // two points per interpreter dispatch (chain_eval_n): i and i + nthreads
        for (int i = tid; i < PW * PH; i += 2 * NTH) {
            double x[2], y[2];
            int pi[2], pj[2];
            bool in[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int ii = i + k * NTH;
                pj[k] = ii / PW;
                pi[k] = ii - pj[k] * PW;
                const int gx = WX0 + pi[k], gy = WY0 + pj[k];
                x[k] = (double)gx;
                y[k] = (double)gy;
                in[k] = ii < PW * PH;
            }
            chain_eval_n<A, SPH, 2>(P.chain, x, y);
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (in[k]) s_p[pj[k]][pi[k]] = make_double2(x[k], y[k]);
        }
    }
    __syncthreads();

    // ---- 2. Jmag2 per cell: cell (ci,cj) spans points (ci..ci+1, cj..cj+1) ----
    // Alongside: min and max of Jmag2 over the tile's in-grid cells.  If max <= thresh * min no cell
    // of the tile can be flagged (its k-th smallest neighbour is >= min), and phase 3 is skipped.
    // Non-negative floats order like their bit patterns, so the warp reduction is one REDUX each.
    __shared__ unsigned int s_mm[2];
    if (tid == 0) { s_mm[0] = 0xffffffffu; s_mm[1] = 0u; }      // ordered by the barrier after phase 1
    unsigned int mn_bits = 0xffffffffu, mx_bits = 0u;
    const int NCX = GW - 1, NCY = GH - 1;              // cell grid extent
    const int wx0 = WX0, wy0 = WY0;
    for (int i = tid; i < CW * CH; i += NTH) {
        const int cj = i / CW, ci = i - cj * CW;
        const double2 pa = s_p[cj + 1][ci + 1], pb = s_p[cj + 1][ci], pc = s_p[cj][ci + 1], pd = s_p[cj][ci];
        const double ax = pa.x, bx = pb.x, cx = pc.x, dx = pd.x, ay = pa.y, by = pb.y, cy = pc.y, dy = pd.y;
        // helpers.pyx:897-906: ((a-b)+c)-d over 2 and ((a+b)-c)-d over 2
        const double j00 = __dmul_rn(__dsub_rn(__dadd_rn(__dsub_rn(ax, bx), cx), dx), 0.5);
        const double j10 = __dmul_rn(__dsub_rn(__dadd_rn(__dsub_rn(ay, by), cy), dy), 0.5);
        const double j01 = __dmul_rn(__dsub_rn(__dsub_rn(__dadd_rn(ax, bx), cx), dx), 0.5);
        const double j11 = __dmul_rn(__dsub_rn(__dsub_rn(__dadd_rn(ay, by), cy), dy), 0.5);
        const double m2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(j00, j00), __dmul_rn(j01, j01)),
                                              __dmul_rn(j10, j10)), __dmul_rn(j11, j11));
        s_m2[cj][ci] = m2;
        const float m2f = (float)m2;
        s_m2f[cj][ci] = m2f;
        if ((unsigned)(wx0 + ci) < (unsigned)NCX && (unsigned)(wy0 + cj) < (unsigned)NCY) {
            const unsigned int b = __float_as_uint(m2f);         // NaN sorts above inf: fails the test below
            mn_bits = min(mn_bits, b);
            mx_bits = max(mx_bits, b);
        }
    }
    mn_bits = __reduce_min_sync(0xffffffffu, mn_bits);
    mx_bits = __reduce_max_sync(0xffffffffu, mx_bits);
    if ((tid & 31) == 0) { atomicMin(&s_mm[0], mn_bits); atomicMax(&s_mm[1], mx_bits); }
    __syncthreads();

    // ---- 3. jump flags for the (JTX+1) x (JTY+1) cells the tile's pixels average over ----
    const float thr_f = (float)P.jump_thresh;
    const float tile_mn = __uint_as_float(s_mm[0]), tile_mx = __uint_as_float(s_mm[1]);
    // CTA-uniform: every cell of the tile passes (or detection is off): no flags to compute or read
    tile_clean = (P.jump_thresh == 0.0) ||
                 (thr_f > 0.0f && tile_mx < 1.0e30f && tile_mx * 1.00001f <= thr_f * tile_mn * 0.99999f);
    if (!tile_clean)
    for (int i = tid; i < (JTX + 1) * (JTY + 1); i += NTH) {
        const int fj = i / (JTX + 1), fi = i - fj * (JTX + 1);
        const int ci = fi + 1, cj = fj + 1;            // index into s_m2 (cell halo of 1)
        const int gcx = wx0 + ci, gcy = wy0 + cj;      // global cell index
        unsigned char flag = 1;
        if (P.jump_thresh != 0.0 && (unsigned)gcx < (unsigned)NCX && (unsigned)gcy < (unsigned)NCY) {
            const bool interior = (gcx >= 1 && gcx <= NCX - 2 && gcy >= 1 && gcy <= NCY - 2);
            bool decided = false;
            if (interior) {
                // quick accept in fp32 (fully unrolled): the k-th smallest is >= the minimum, so for
                // a positive threshold me <= thr*min implies me <= thr*kth; the margins cover the
                // fp32 rounding.  Only cells near a jump (or non-finite ones) take the exact sort.
                const float *r0 = &s_m2f[cj - 1][ci - 1], *r1 = &s_m2f[cj][ci - 1], *r2 = &s_m2f[cj + 1][ci - 1];
                const float mn = fminf(fminf(fminf(r0[0], r0[1]), fminf(r0[2], r1[0])),
                                       fminf(fminf(r1[1], r1[2]), fminf(fminf(r2[0], r2[1]), r2[2])));
                const float mef = r1[1];
                decided = (thr_f > 0.0f) && (mef < 1.0e30f) && (mef * 1.00001f <= thr_f * mn * 0.99999f);
            }
            if (!decided) flag = jump_flag_exact(&s_m2[0][0], CW, ci, cj, gcx, gcy, NCX, NCY, P.jump_thresh);
        }
        s_flag[cj][ci] = flag;
    }
    if (!tile_clean) __syncthreads();
    }   // !analytic

    // ---- 4. per pixel: grid-centred Jacobian, SVD padding, footprint ----
    int state = 0;
    // Thread -> pixel map of the footprint phase: a warp covers a K2_WX x (32 / K2_WX) patch of the
    // tile instead of one 32-pixel row, so that the 32 footprints a tap load gathers from sit in a
    // compact source patch (an arc of 32 pixels of a polar unwrap spans up to 140 source pixels: one
    // tap load then touches ~8 cache lines, and the L1 tag stage, one line per clock, was the limit).
    // Measured at 8192^2 (polar unwrap / rotated affine down-scale): 32x1 2.70 / 0.70 ms, 16x2 2.20 /
    // 0.69, 8x4 2.00 / 0.72, 4x8 1.98 / 0.82 -- the host picks log2(width) per launch (P.lw).
    const int lw = P.lw;                               // 5, 4, 3 or 2: patch is (1 << lw) x (32 >> lw)
    const int lane = tid & 31, wrp = tid >> 5;
    const int tx = ((wrp & ((32 >> lw) - 1)) << lw) + (lane & ((1 << lw) - 1));
    const int ty = (wrp >> (5 - lw)) * (32 >> lw) + (lane >> lw);
    const int X = blockIdx.x * JTX + tx;
#pragma unroll 1
    for (int jrow = 0; jrow < JTY / 8; ++jrow) {
    const int ly = ty + 8 * jrow;                       // row inside the tile
    const int Y = blockIdx.y * JTY + ly;
    if (X >= P.dst_w || Y >= P.dst_h) continue;
    const int gx = (int)P.dst_x0 + X, gy = (int)P.dst_y0 + Y;
    // edge rows/columns copy their inner neighbour (helpers.pyx:1244-1247): clamp to [1, G-2]
    const int jx = gx < 1 ? 1 : (gx > GW - 2 ? GW - 2 : gx);
    const int jy = gy < 1 ? 1 : (gy > GH - 2 ? GH - 2 : gy);
    // point index of (jx,jy) inside the smem tile; the clamp moves by at most 1, halo is 2
    const int qi = jx - WX0, qj = jy - WY0;

    double Ji[4] = {0.0, 0.0, 0.0, 0.0};
    double px, py;
    if (affc) {
        px = fma(P.aff[0], (double)gx, fma(P.aff[1], (double)gy, P.aff[4]));
        py = fma(P.aff[2], (double)gx, fma(P.aff[3], (double)gy, P.aff[5]));
        Ji[0] = P.aff[0]; Ji[1] = P.aff[1]; Ji[2] = P.aff[2]; Ji[3] = P.aff[3];
    } else if (analytic) {
        const tfm_op &op0 = P.chain.ops[0];
        const double TX0 = (double)(GX0 - 1), TY0 = (double)(GY0 - 1);
        polar_from_tables(op0, TX0, TY0, tx + 1, ly + 1, s_ta, s_tb, s_ea, s_eb, px, py);
        // closed-form Jacobian at the (edge-clamped) grid point (jx, jy)
        const int ti = jx - (GX0 - 1), tj = jy - (GY0 - 1);
        const double2 ca = s_ta[ti], cb = s_tb[tj];
        const bool lg = (op0.flags & TFM_RAD_R0_NOT_NONE) != 0;
        const float sgn = (op0.flags & TFM_RAD_CCW) ? 1.0f : -1.0f;
        const float c = (float)(ca.x * cb.x - ca.y * cb.y), sn = (float)(ca.y * cb.x + ca.x * cb.y);    // cos, sin(theta)
        const float r = lg ? (float)(s_ea[ti] * s_eb[tj])
                           : (float)fma(op0.p[2], TX0 + (double)ti, fma(op0.p[3], TY0 + (double)tj, op0.p[5]));
        const float cf = (float)op0.p[6];
        const float tX = (float)op0.p[0] * cf, tY = (float)op0.p[1] * cf;
        const float drX = lg ? r * (float)op0.p[2] : (float)op0.p[2], drY = lg ? r * (float)op0.p[3] : (float)op0.p[3];
        // x1 = r cos + ox, y1 = sgn r sin + oy
        float j00 = c * drX - r * sn * tX, j01 = c * drY - r * sn * tY;
        float j10 = sgn * (sn * drX + r * c * tX), j11 = sgn * (sn * drY + r * c * tY);
        if (op0.flags & TFM_POLAR_HAS_A2) {
            const float a = (float)op0.p[10], b = (float)op0.p[11], cc = (float)op0.p[12], d = (float)op0.p[13];
            const float k00 = a * j00 + b * j10, k01 = a * j01 + b * j11;
            const float k10 = cc * j00 + d * j10, k11 = cc * j01 + d * j11;
            j00 = k00; j01 = k01; j10 = k10; j11 = k11;
        }
        Ji[0] = j00; Ji[1] = j01; Ji[2] = j10; Ji[3] = j11;
    } else {
    const bool all_ok = tile_clean ||
                        (s_flag[qj - 1][qi - 1] & s_flag[qj][qi - 1] & s_flag[qj - 1][qi] & s_flag[qj][qi]) != 0;
    if (FASTI && all_ok) {
        // no flagged cell: the four cell Jacobians add up to the Sobel-like stencil
        //   d/dx = (I[y+1][x+1]-I[y+1][x-1] + 2(I[y][x+1]-I[y][x-1]) + I[y-1][x+1]-I[y-1][x-1]) / 8
        // (identical up to fp64 rounding of the regrouped sum)
        const double2 *r0 = &s_p[qj - 1][qi - 1], *r1 = &s_p[qj][qi - 1], *r2 = &s_p[qj + 1][qi - 1];
        const double2 a0 = r0[0], a1 = r0[1], a2 = r0[2], b0 = r1[0], b2 = r1[2], c0 = r2[0], c1 = r2[1], c2 = r2[2];
        Ji[0] = 0.125 * ((c2.x - c0.x) + 2.0 * (b2.x - b0.x) + (a2.x - a0.x));
        Ji[2] = 0.125 * ((c2.y - c0.y) + 2.0 * (b2.y - b0.y) + (a2.y - a0.y));
        Ji[1] = 0.125 * ((c0.x - a0.x) + 2.0 * (c1.x - a1.x) + (c2.x - a2.x));
        Ji[3] = 0.125 * ((c0.y - a0.y) + 2.0 * (c1.y - a1.y) + (c2.y - a2.y));
    } else {
        int nvalid = 0;
        // helpers.pyx:1227-1231: cells (y-1,x-1) + (y,x-1) + (y-1,x) + (y,x), in that order
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ci = qi - 1 + (k >> 1), cj = qj - 1 + (k & 1);
            const bool ok = tile_clean || s_flag[cj][ci] != 0;
            nvalid += ok ? 1 : 0;
            const double2 pa = s_p[cj + 1][ci + 1], pb = s_p[cj + 1][ci], pc = s_p[cj][ci + 1], pd = s_p[cj][ci];
            const double ax = pa.x, bx = pb.x, cx = pc.x, dx = pd.x, ay = pa.y, by = pb.y, cy = pc.y, dy = pd.y;
            double j00 = __dmul_rn(__dsub_rn(__dadd_rn(__dsub_rn(ax, bx), cx), dx), 0.5);
            double j10 = __dmul_rn(__dsub_rn(__dadd_rn(__dsub_rn(ay, by), cy), dy), 0.5);
            double j01 = __dmul_rn(__dsub_rn(__dsub_rn(__dadd_rn(ax, bx), cx), dx), 0.5);
            double j11 = __dmul_rn(__dsub_rn(__dsub_rn(__dadd_rn(ay, by), cy), dy), 0.5);
            if (!ok) { j00 = 0.0; j01 = 0.0; j10 = 0.0; j11 = 0.0; }          // helpers.pyx:1207
            if (k == 0) { Ji[0] = j00; Ji[1] = j01; Ji[2] = j10; Ji[3] = j11; }
            else {
                Ji[0] = __dadd_rn(Ji[0], j00); Ji[1] = __dadd_rn(Ji[1], j01);
                Ji[2] = __dadd_rn(Ji[2], j10); Ji[3] = __dadd_rn(Ji[3], j11);
            }
        }
        const double denom = (P.jump_thresh != 0.0) ? (double)(nvalid < 1 ? 1 : nvalid) : 4.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) Ji[k] = __ddiv_rn(Ji[k], denom);
    }

    const double2 pme = s_p[ly + 2 + shy][tx + 2 + shx];
    px = pme.x;
    py = pme.y;
    }   // !analytic
    DstT *drow = reinterpret_cast<DstT *>(dst_base + (long long)Y * P.dst_pitch);
    if (!(fabs(px) < 4.0e18) || !(fabs(py) < 4.0e18)) { drow[X] = (DstT)0; continue; }
    const double fxs = floor(px), fys = floor(py);
    const long long xs = (long long)fxs, ys = (long long)fys;

    if constexpr (!FASTI) {
        // ------------------------- parity mode: the reference's loop, in its order -------------
        double U[4], V[4], s[2], J[4];
        svd2x2_dev(Ji, U, s, V);                                                     // helpers.pyx:1659
        s[0] = exp(__ddiv_rn(-log(__dadd_rn(P.padval, exp(__dmul_rn(log(s[0]), P.pad_pow)))), P.pad_pow));
        s[1] = exp(__ddiv_rn(-log(__dadd_rn(P.padval, exp(__dmul_rn(log(s[1]), P.pad_pow)))), P.pad_pow));
        svc2x2_dev(V, s, U, J);                                                      // helpers.pyx:1663
        const double smin = fmin(s[0], s[1]);
        const double rs = __dmul_rn(2.0, ceil(__ddiv_rn(P.oblur, smin)));            // helpers.pyx:1666
        long long reg = (rs < 1.0e15) ? (long long)rs : 1000000000000000LL;
        if (P.max_reg > 0 && reg > P.max_reg) reg = P.max_reg;
        const int half = (int)(reg / 2 > 0x3fffffff ? 0x3fffffff : reg / 2);
        const double xf_of = __dsub_rn(px, fxs), yf_of = __dsub_rn(py, fys);
        const double ob = P.oblur;
        double val = 0, wgt = 0;
        // Footprint wholly inside the image and the staged rectangle (nearly every pixel): no boundary rule can
        // fire, so the taps are consecutive elements of consecutive rows.  Same operations on the same operands in
        // the same order as the general loop below -- bit-identical -- minus what does not depend on the tap:
        // ncu on the general loop (profiles/r2_kernels.md, jacobian_parity): 8,500 instructions per pixel,
        // ~190 per tap, of which the weight itself (two dot products, exp) is under 50.  Row terms J01 y0 and
        // J11 y0 are formed once per row; tap offsets count in doubles (exact integers) instead of converting;
        // |.| before squaring is dropped ((-v)(-v) is v v exactly).  (flags 0x200: tuning bit 2097152 disables it.)
        const long long wx0 = xs - half, wy0 = ys - half;
        // A footprint wholly beyond a truncated edge: every tap sets tflag and none adds its value
        // (helpers.pyx:1754-1755, 1774-1775, 1793-1796), so the pixel is 0 / wgt = 0 whatever the weights are --
        // unless the other axis forbids, where the reference raises from inside the same loop.  These pixels carry
        // the largest footprints of an unwrap (its corners): in the profile of the general loop they were 3/4 of
        // all taps evaluated.
        const bool all_out = !(P.flags & 0x200) && P.bound_x != TFM_B_FORBID && P.bound_y != TFM_B_FORBID &&
                             ((P.bound_x == TFM_B_TRUNCATE && (xs + half < 0 || wx0 >= P.full_w)) ||
                              (P.bound_y == TFM_B_TRUNCATE && (ys + half < 0 || wy0 >= P.full_h)));
        if (all_out) { drow[X] = (DstT)0; continue; }
        const bool inside = !(P.flags & 0x200) && half <= 16384 &&
                            wx0 >= P.src_x0 && wx0 >= 0 && wx0 + 2LL * half < min((long long)P.full_w, (long long)P.src_x0 + P.src_w) &&
                            wy0 >= P.src_y0 && wy0 >= 0 && wy0 + 2LL * half < min((long long)P.full_h, (long long)P.src_y0 + P.src_h);
        if (inside) {
            const char *prow = src_base + (wy0 - P.src_y0) * P.src_pitch + (wx0 - P.src_x0) * (long long)sizeof(SrcT);
            const bool lean = (P.filter == TFM_FILT_GAUSSIAN) && (ob == 1.0);
            const int n = 2 * half + 1;
            const double first = -(double)half;
            // Gaussian: exp(-(xf^2 + yf^2)) = exp(-(A a^2 + 2 B a b + C b^2)) in the tap offsets (a, b), A, B, C
            // from J and oblur.  The recurrence is taken when no weight of the window can leave the normal
            // range (flags 0x400: tuning bit 4194304 keeps one exp per tap, bit-identical with the general loop).
            double qA = 0, qB = 0, qC = 0, cA = 0;
            bool recur = false;
            if (P.filter == TFM_FILT_GAUSSIAN && !(P.flags & 0x400)) {
                const double io2 = 1.0 / (ob * ob);
                qA = (J[0] * J[0] + J[2] * J[2]) * io2;
                qB = (J[0] * J[1] + J[2] * J[3]) * io2;
                qC = (J[1] * J[1] + J[3] * J[3]) * io2;
                const double hh = (double)(half + 1);
                recur = (qA + 2.0 * fabs(qB) + qC) * (hh * hh) < 600.0;
                cA = exp(-2.0 * qA);
            }
            double yrd = first;
            for (int j = 0; j < n; ++j, yrd += 1.0, prow += P.src_pitch) {
                const SrcT *const p = reinterpret_cast<const SrcT *>(prow);
                const double y0 = __dadd_rn(yrd, yf_of);
                const double a1 = __dmul_rn(J[1], y0), a3 = __dmul_rn(J[3], y0);
                double xrd = first;
                if (recur) {
                    // w(a + 1) = w(a) r(a), r(a + 1) = r(a) exp(-2A): two multiplications per tap instead of an
                    // exp (~40 instructions).  Both are restarted from exp() every row and every 32 taps, so a
                    // weight carries at most ~32^2 / 2 roundings: 1e-13 relative, inside the stated 1e-11.
                    for (int i0 = 0; i0 < n; i0 += 32) {
                        const double x0 = __dadd_rn(first + (double)i0, xf_of);
                        double w = exp(-((qA * x0 + 2.0 * (qB * y0)) * x0 + qC * (y0 * y0)));
                        double r = exp(-(qA * (2.0 * x0 + 1.0) + 2.0 * (qB * y0)));
                        const int i1 = min(n, i0 + 32);
#pragma unroll 4
                        for (int i = i0; i < i1; ++i) {
                            wgt = __dadd_rn(wgt, w);
                            val = __dadd_rn(val, __dmul_rn((double)__ldg(p + i), w));
                            w *= r;
                            r *= cA;
                        }
                    }
                } else if (lean) {
#pragma unroll 2
                    for (int i = 0; i < n; ++i, xrd += 1.0) {
                        const double x0 = __dadd_rn(xrd, xf_of);
                        const double xf = __dadd_rn(__dmul_rn(J[0], x0), a1), yf = __dadd_rn(__dmul_rn(J[2], x0), a3);
                        const double filt = exp(-__dadd_rn(__dmul_rn(xf, xf), __dmul_rn(yf, yf)));
                        wgt = __dadd_rn(wgt, filt);
                        val = __dadd_rn(val, __dmul_rn((double)__ldg(p + i), filt));
                    }
                } else {
                    for (int i = 0; i < n; ++i, xrd += 1.0) {
                        const double x0 = __dadd_rn(xrd, xf_of);
                        double xf = fabs(__dadd_rn(__dmul_rn(J[0], x0), a1)), yf = fabs(__dadd_rn(__dmul_rn(J[2], x0), a3));
                        if (ob != 1.0) { xf = __ddiv_rn(xf, ob); yf = __ddiv_rn(yf, ob); }
                        const double filt = filt_eval<double>(P.filter, xf, yf);
                        wgt = __dadd_rn(wgt, filt);
                        val = __dadd_rn(val, __dmul_rn((double)__ldg(p + i), filt));
                    }
                }
            }
        } else
        for (int yr = -half; yr <= half; ++yr) {
            const double y0 = __dadd_rn((double)yr, yf_of);
            for (int xr = -half; xr <= half; ++xr) {
                const double x0 = __dadd_rn((double)xr, xf_of);
                // helpers.pyx:1685-1690; x / 1.0 is x exactly, and the default oblur is 1: skip two
                // IEEE divisions (~25 instructions each) per tap in that case
                double xf = fabs(__dadd_rn(__dmul_rn(J[0], x0), __dmul_rn(J[1], y0)));
                double yf = fabs(__dadd_rn(__dmul_rn(J[2], x0), __dmul_rn(J[3], y0)));
                if (ob != 1.0) { xf = __ddiv_rn(xf, ob); yf = __ddiv_rn(yf, ob); }
                const double filt = filt_eval<double>(P.filter, xf, yf);
                wgt = __dadd_rn(wgt, filt);                                          // helpers.pyx:1740
                int st = 0;
                int iy = jbound(ys + yr, P.full_h, P.bound_y, st);
                int ix = jbound(xs + xr, P.full_w, P.bound_x, st);
                state |= st & 2;
                if (st == 0) {
                    ix -= P.src_x0;
                    iy -= P.src_y0;
                    if ((unsigned)ix < (unsigned)P.src_w && (unsigned)iy < (unsigned)P.src_h) {
                        const double sv = (double)__ldg(reinterpret_cast<const SrcT *>(
                            src_base + (long long)iy * P.src_pitch) + ix);
                        val = __dadd_rn(val, __dmul_rn(sv, filt));                   // helpers.pyx:1796
                    } else {
                        state |= 4;
                    }
                }
            }
        }
        double res = (wgt > 0) ? val / wgt : 0.0;                                    // helpers.pyx:1800
        if (P.flags & TFM_F_FLUX)                                                    // SVD.pyx:285-286
            res = __dmul_rn(res, fabs(__dsub_rn(__dmul_rn(Ji[0], Ji[3]), __dmul_rn(Ji[1], Ji[2]))));
        drow[X] = (DstT)res;
    } else {
        // ------------------------- fast mode (fp32) --------------------------------------------
        const float ob = (float)P.oblur, iob = fast_rcp(ob);
        const bool gauss = (P.filter == TFM_FILT_GAUSSIAN);
        float J[4] = {0.f, 0.f, 0.f, 0.f}, smin, inv_smin, q00 = 0.f, q01 = 0.f, q11 = 0.f;
        int half_const = -1;
        if (affc) { q00 = P.aff_q[0]; q01 = P.aff_q[1]; q11 = P.aff_q[2]; inv_smin = 1.0f; half_const = P.aff_half; }
        else if (gauss) gaussian_quadratic_f32(Ji, (float)P.padval, (float)P.pad_pow, q00, q01, q11, inv_smin);
        else { padded_inverse_f32(Ji, (float)P.padval, (float)P.pad_pow, J, smin); inv_smin = fast_rcp(smin); }
        float ratio = ob * inv_smin;
        if (half_const < 0 && fabsf(ratio - rintf(ratio)) < 4e-5f * ratio) {
            // the footprint size is a ceil(): when fp32 cannot decide it, redo that one number in fp64
            double U[4], V[4], sd[2];
            svd2x2_dev(Ji, U, sd, V);
            const double s0 = exp(-log(P.padval + exp(log(sd[0]) * P.pad_pow)) / P.pad_pow);
            const double s1 = exp(-log(P.padval + exp(log(sd[1]) * P.pad_pow)) / P.pad_pow);
            ratio = (float)ceil(P.oblur / fmin(s0, s1));
        }
        const float rsf = 2.0f * ceilf(ratio);
        int reg = (rsf < 1.0e9f) ? (int)rsf : 1000000000;
        if (P.max_reg > 0 && (long long)reg > P.max_reg) reg = (int)P.max_reg;
        const int half = (half_const >= 0) ? half_const : reg / 2;
        const float xf_of = (float)(px - fxs), yf_of = (float)(py - fys);
        float val = 0.0f, wgt = 0.0f;
        const bool whole = (P.src_x0 == 0 && P.src_y0 == 0 && P.src_w == P.full_w && P.src_h == P.full_h);
        // 32-bit tap origin for every coordinate an image can actually contain; absurd ones keep the
        // 64-bit tests (and can never be `inside`)
        const bool sane32 = (fabs(px) < 1.0e9) && (fabs(py) < 1.0e9);
        const int xi = sane32 ? (int)fxs : 0, yi = sane32 ? (int)fys : 0;
        const bool inside = whole && sane32 && xi - half >= 0 && xi + half < P.full_w &&
                            yi - half >= 0 && yi + half < P.full_h;
        // a footprint entirely beyond a truncated edge: every tap is skipped, val stays 0 and the
        // pixel is 0 whatever the weights are (helpers.pyx:1793-1800)
        const bool all_out = sane32
            ? ((P.bound_x == TFM_B_TRUNCATE && (xi + half < 0 || xi - half >= P.full_w)) ||
               (P.bound_y == TFM_B_TRUNCATE && (yi + half < 0 || yi - half >= P.full_h)))
            : ((P.bound_x == TFM_B_TRUNCATE && (xs + half < 0 || xs - half >= P.full_w)) ||
               (P.bound_y == TFM_B_TRUNCATE && (ys + half < 0 || ys - half >= P.full_h)));
        if (all_out) { drow[X] = (DstT)0; continue; }
        const long long pe = P.src_pitch / (long long)sizeof(SrcT);
        const float L2E = 1.4426950408889634f;
        const float Aq = q00 * iob * iob * L2E, Bq = q01 * iob * iob * L2E, Cq = q11 * iob * iob * L2E;
        const float h2 = (float)half * (float)half + 2.0f * (float)half + 1.0f;
        const bool recur_fit = gauss && half >= 1 && half <= 6 && (Aq + 2.0f * fabsf(Bq) + Cq) * h2 < 100.0f;
        const bool recur_ok = recur_fit && inside;
        bool tex_done = false;
        if constexpr (sizeof(SrcT) == 4) {
            // texture-gather taps (as the lean kernels of tfm_jacobian_fast.cu): 2 x 2 taps per fetch, and with
            // truncate boundaries the border texel is the truncated tap, so edge footprints need no slow path
            if (recur_fit && sane32 && P.tex_ok && (inside || P.tex_ok == 2)) {
                const float a0 = (float)(-half) + xf_of, b0 = (float)(-half) + yf_of;
                const float w0 = fast_ex2(-((Aq * a0 + 2.0f * Bq * b0) * a0 + Cq * b0 * b0));
                const float r0 = fast_ex2(-(Aq * (2.0f * a0 + 1.0f) + 2.0f * Bq * b0));
                const float rb = fast_ex2(-(Cq * (2.0f * b0 + 1.0f) + 2.0f * Bq * a0));
                const float cA = fast_ex2(-2.0f * Aq), cB = fast_ex2(-2.0f * Bq), cC = fast_ex2(-2.0f * Cq);
                const int tx0 = xi - half, ty0 = yi - half;
                switch (half) {
                case 1: gauss_taps_tex<1>(P.tex, tx0, ty0, w0, r0, rb, cA, cB, cC, val, wgt); break;
                case 2: gauss_taps_tex<2>(P.tex, tx0, ty0, w0, r0, rb, cA, cB, cC, val, wgt); break;
                case 3: gauss_taps_tex<3>(P.tex, tx0, ty0, w0, r0, rb, cA, cB, cC, val, wgt); break;
                case 4: gauss_taps_tex<4>(P.tex, tx0, ty0, w0, r0, rb, cA, cB, cC, val, wgt); break;
                case 5: gauss_taps_tex<5>(P.tex, tx0, ty0, w0, r0, rb, cA, cB, cC, val, wgt); break;
                default: gauss_taps_tex<6>(P.tex, tx0, ty0, w0, r0, rb, cA, cB, cC, val, wgt); break;
                }
                tex_done = true;
            }
        }
        if (tex_done) {
        } else if (recur_ok) {
            // exp(-(xf^2+yf^2)) = exp2(-(A a^2 + 2B ab + C b^2)): three exp2 per pixel row set, two
            // multiplies per tap
            const float a0 = (float)(-half) + xf_of, b0 = (float)(-half) + yf_of;
            const float w0 = fast_ex2(-((Aq * a0 + 2.0f * Bq * b0) * a0 + Cq * b0 * b0));
            const float r0 = fast_ex2(-(Aq * (2.0f * a0 + 1.0f) + 2.0f * Bq * b0));
            const float rb = fast_ex2(-(Cq * (2.0f * b0 + 1.0f) + 2.0f * Bq * a0));
            const float cA = fast_ex2(-2.0f * Aq), cB = fast_ex2(-2.0f * Bq), cC = fast_ex2(-2.0f * Cq);
            const SrcT *p0 = reinterpret_cast<const SrcT *>(src_base) + (long long)(yi - half) * pe + (xi - half);
            switch (half) {
            case 1: gauss_taps<1, SrcT>(p0, pe, w0, r0, rb, cA, cB, cC, val, wgt); break;
            case 2: gauss_taps<2, SrcT>(p0, pe, w0, r0, rb, cA, cB, cC, val, wgt); break;
            case 3: gauss_taps<3, SrcT>(p0, pe, w0, r0, rb, cA, cB, cC, val, wgt); break;
            case 4: gauss_taps<4, SrcT>(p0, pe, w0, r0, rb, cA, cB, cC, val, wgt); break;
            case 5: gauss_taps<5, SrcT>(p0, pe, w0, r0, rb, cA, cB, cC, val, wgt); break;
            default: gauss_taps<6, SrcT>(p0, pe, w0, r0, rb, cA, cB, cC, val, wgt); break;
            }
        } else if (!gauss && inside) {
            // tent / Lanczos / "Hanning" / Tukey with the whole footprint inside the image: plain loads
            const float J0 = J[0] * iob, J1 = J[1] * iob, J2 = J[2] * iob, J3 = J[3] * iob;
            const SrcT *prow = reinterpret_cast<const SrcT *>(src_base) + (long long)(yi - half) * pe + (xi - half);
            const int nt = 2 * half + 1;
            for (int j = 0; j < nt; ++j) {
                const float y0 = (float)(j - half) + yf_of;
                const float ux = J1 * y0, uy = J3 * y0;
#pragma unroll 4
                for (int k = 0; k < nt; ++k) {
                    const float x0 = (float)(k - half) + xf_of;
                    const float filt = filt_eval<float>(P.filter, fabsf(fmaf(J0, x0, ux)), fabsf(fmaf(J2, x0, uy)));
                    wgt += filt;
                    val = fmaf((float)__ldg(prow + k), filt, val);
                }
                prow += pe;
            }
        } else {
            const float J0 = J[0] * iob, J1 = J[1] * iob, J2 = J[2] * iob, J3 = J[3] * iob;
            for (int yr = -half; yr <= half; ++yr) {
                const float y0 = (float)yr + yf_of;
                int sty = 0;
                const int iy0 = jbound(ys + yr, P.full_h, P.bound_y, sty);
                for (int xr = -half; xr <= half; ++xr) {
                    const float x0 = (float)xr + xf_of;
                    float filt;
                    if (gauss) filt = fast_ex2(-((Aq * x0 + 2.0f * Bq * y0) * x0 + Cq * y0 * y0));
                    else filt = filt_eval<float>(P.filter, fabsf(J0 * x0 + J1 * y0), fabsf(J2 * x0 + J3 * y0));
                    wgt += filt;
                    int st = sty;
                    int ix = jbound(xs + xr, P.full_w, P.bound_x, st);
                    state |= st & 2;
                    if (st == 0) {
                        ix -= P.src_x0;
                        const int iy = iy0 - P.src_y0;
                        if ((unsigned)ix < (unsigned)P.src_w && (unsigned)iy < (unsigned)P.src_h)
                            val = fmaf((float)__ldg(reinterpret_cast<const SrcT *>(src_base) + iy * pe + ix), filt, val);
                        else
                            state |= 4;
                    }
                }
            }
        }
        float res = (wgt > 0) ? val * fast_rcp(wgt) : 0.0f;
        if (P.flags & TFM_F_FLUX) res *= fabsf((float)Ji[0] * (float)Ji[3] - (float)Ji[1] * (float)Ji[2]);
        drow[X] = (DstT)res;
    }
    }   // jrow
    if (state && P.status) atomicOr(P.status, (state & 2 ? 1 : 0) | (state & 4 ? 2 : 0));
}

// ---------------------------------------------------------------------------------------------
// Stand-alone entry points for the three Jacobian helpers the reference exposes as module functions
// (helpers.pyx:846-958 simple_jacobian, 960-1141 jump_detect, 1145-1290 jacobian), 2-D case.  They
// share the arithmetic of the fused kernel above (same differences, same selection networks, same
// summation order) but work on arrays in global memory.
// ---------------------------------------------------------------------------------------------
__global__ void k_simple_jacobian(const double2 *idx, int ny, int nx, double *Js)
{
    const int cx = blockIdx.x * blockDim.x + threadIdx.x, cy = blockIdx.y * blockDim.y + threadIdx.y;
    if (cx >= nx - 1 || cy >= ny - 1) return;
    const double2 pa = idx[(long long)(cy + 1) * nx + cx + 1], pb = idx[(long long)(cy + 1) * nx + cx];
    const double2 pc = idx[(long long)cy * nx + cx + 1], pd = idx[(long long)cy * nx + cx];
    double *o = Js + ((long long)cy * (nx - 1) + cx) * 4;
    o[0] = __dmul_rn(__dsub_rn(__dadd_rn(__dsub_rn(pa.x, pb.x), pc.x), pd.x), 0.5);    // d x / d pix_x
    o[1] = __dmul_rn(__dsub_rn(__dsub_rn(__dadd_rn(pa.x, pb.x), pc.x), pd.x), 0.5);    // d x / d pix_y
    o[2] = __dmul_rn(__dsub_rn(__dadd_rn(__dsub_rn(pa.y, pb.y), pc.y), pd.y), 0.5);    // d y / d pix_x
    o[3] = __dmul_rn(__dsub_rn(__dsub_rn(__dadd_rn(pa.y, pb.y), pc.y), pd.y), 0.5);    // d y / d pix_y
}

__global__ void k_jmag2(const double *Js, long long n, double *m2)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *j = Js + i * 4;
    m2[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(j[0], j[0]), __dmul_rn(j[1], j[1])),
                                __dmul_rn(j[2], j[2])), __dmul_rn(j[3], j[3]));
}

__global__ void k_jump_detect(const double *m2, int ncy, int ncx, double thresh, long long *flags)
{
    const int cx = blockIdx.x * blockDim.x + threadIdx.x, cy = blockIdx.y * blockDim.y + threadIdx.y;
    if (cx >= ncx || cy >= ncy) return;
    flags[(long long)cy * ncx + cx] = jump_flag_exact(m2, ncx, cx, cy, cx, cy, ncx, ncy, thresh);
}

__global__ void k_jacobian_grid(const double *Js, const long long *flags, int ny, int nx, double *J)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nx || y >= ny) return;
    // edge rows / columns copy their inner neighbour (helpers.pyx:1244-1247)
    const int jx = x < 1 ? 1 : (x > nx - 2 ? nx - 2 : x), jy = y < 1 ? 1 : (y > ny - 2 ? ny - 2 : y);
    const int ncx = nx - 1;
    double acc[4] = {0, 0, 0, 0};
    int nvalid = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {          // cells (y-1,x-1) + (y,x-1) + (y-1,x) + (y,x), helpers.pyx:1227-1231
        const int cx = jx - 1 + (k >> 1), cy = jy - 1 + (k & 1);
        const long long c = (long long)cy * ncx + cx;
        const bool ok = flags ? (flags[c] != 0) : true;
        nvalid += ok ? 1 : 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double v = ok ? Js[c * 4 + q] : 0.0;
            acc[q] = (k == 0) ? v : __dadd_rn(acc[q], v);
        }
    }
    const double denom = flags ? (double)(nvalid < 1 ? 1 : nvalid) : 4.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) J[((long long)y * nx + x) * 4 + q] = __ddiv_rn(acc[q], denom);
}

int jacobian_helpers(int which, const void *in0, const void *in1, long long ny, long long nx, double thresh,
                     void *out, void *scratch, cudaStream_t st)
{
    if (ny < 2 || nx < 2 || ny > 0x3fffffff || nx > 0x3fffffff || !in0 || !out) return TFM_EVALUE;
    dim3 block(32, 8);
    if (which == 0) {                       // simple_jacobian: index[ny][nx][2] -> Js[ny-1][nx-1][2][2]
        dim3 grid((unsigned)((nx - 1 + 31) / 32), (unsigned)((ny - 1 + 7) / 8));
        k_simple_jacobian<<<grid, block, 0, st>>>(static_cast<const double2 *>(in0), (int)ny, (int)nx,
                                                  static_cast<double *>(out));
    } else if (which == 1) {                // jump_detect: Js[ny][nx][2][2] (cell grid) -> flags[ny][nx] (int64)
        if (!scratch) return TFM_EVALUE;
        const long long n = ny * nx;
        k_jmag2<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(static_cast<const double *>(in0), n,
                                                              static_cast<double *>(scratch));
        count_launch();
        dim3 grid((unsigned)((nx + 31) / 32), (unsigned)((ny + 7) / 8));
        k_jump_detect<<<grid, block, 0, st>>>(static_cast<const double *>(scratch), (int)ny, (int)nx, thresh,
                                              static_cast<long long *>(out));
    } else {                                // jacobian: Js (cells), flags|NULL -> J[ny][nx][2][2] (grid points)
        if (ny < 3 || nx < 3) return TFM_EVALUE;
        dim3 grid((unsigned)((nx + 31) / 32), (unsigned)((ny + 7) / 8));
        k_jacobian_grid<<<grid, block, 0, st>>>(static_cast<const double *>(in0), static_cast<const long long *>(in1),
                                                (int)ny, (int)nx, static_cast<double *>(out));
    }
    count_launch();
    set_last_kernel(which == 0 ? "k_simple_jacobian" : (which == 1 ? "k_jump_detect" : "k_jacobian_grid"));
    return cudaGetLastError() == cudaSuccess ? TFM_OK : TFM_ECUDA;
}

__global__ void k_svd_hook(int which, const double *in, double *out)
{
    if (which == 0) svd2x2_dev(in, out, out + 4, out + 6);          // M -> U, s, V
    else svc2x2_dev(in, in + 4, in + 6, out);                       // U, s, V -> M
}

int svd_hooks(int which, const double *in0, const double *in1, const double *in2,
              double *out0, double *out1, double *out2)
{
    double h_in[10] = {0}, h_out[10] = {0};
    if (which == 0) { for (int i = 0; i < 4; ++i) h_in[i] = in0[i]; }
    else {
        for (int i = 0; i < 4; ++i) { h_in[i] = in0[i]; h_in[6 + i] = in2[i]; }
        h_in[4] = in1[0]; h_in[5] = in1[1];
    }
    double *d = nullptr;
    if (cudaMalloc(&d, 20 * sizeof(double)) != cudaSuccess) return TFM_ECUDA;
    cudaError_t e = cudaMemcpy(d, h_in, 10 * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_svd_hook<<<1, 1>>>(which, d, d + 10);
        count_launch();
        e = cudaMemcpy(h_out, d + 10, 10 * sizeof(double), cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    if (e != cudaSuccess) return TFM_ECUDA;
    if (which == 0) {
        for (int i = 0; i < 4; ++i) { out0[i] = h_out[i]; out2[i] = h_out[6 + i]; }
        out1[0] = h_out[4]; out1[1] = h_out[5];
    } else {
        for (int i = 0; i < 4; ++i) out0[i] = h_out[i];
    }
    return TFM_OK;
}

int jacobian_dispatch(const tfm_jacobian_args *a, cudaStream_t st)
{
    if (!a || a->abi_version != TFM_ABI_VERSION) return TFM_EVALUE;
    if (a->n_ops < 0 || a->n_ops > TFM_MAX_OPS || (a->n_ops && !a->chain)) return TFM_EVALUE;
    if (a->filter < TFM_FILT_TENT || a->filter > TFM_FILT_HANN_WINDOW) return TFM_EVALUE;   // helpers.pyx:1505
    if (a->bound_x < TFM_B_FORBID || a->bound_x > TFM_B_MIRROR ||
        a->bound_y < TFM_B_FORBID || a->bound_y > TFM_B_MIRROR) return TFM_EVALUE;       // helpers.pyx:1516
    const tfm_image &s = a->src, &d = a->dst;
    if (s.h < 0 || s.w < 0 || d.h < 0 || d.w < 0) return TFM_EVALUE;
    if (s.n_planes < 1 || s.n_planes != d.n_planes || s.n_planes > 65535) return TFM_EVALUE;
    if (s.full_h < s.h || s.full_w < s.w || s.full_h > 0x3fffffff || s.full_w > 0x3fffffff) return TFM_EVALUE;
    if (d.full_h < 3 || d.full_w < 3 || d.full_h > 0x3fffffff || d.full_w > 0x3fffffff) return TFM_EVALUE;
    if (d.x0 < 0 || d.y0 < 0 || d.x0 + d.w > d.full_w || d.y0 + d.h > d.full_h) return TFM_EVALUE;
    if (d.h == 0 || d.w == 0) return TFM_OK;
    if (s.full_h == 0 || s.full_w == 0 || !s.data || !d.data) return TFM_EVALUE;
    if (!(a->oblur > 0) || !(a->pad_pow > 0)) return TFM_EVALUE;
    if (!chain_ops_valid(a->chain, a->n_ops)) return TFM_EVALUE;
    const bool subrect = (s.x0 != 0 || s.y0 != 0 || s.full_h != s.h || s.full_w != s.w);
    if ((a->bound_x == TFM_B_FORBID || a->bound_y == TFM_B_FORBID || subrect) && !a->status_dev)
        return TFM_EVALUE;
    // the kernel indexes the output grid with 32-bit integers
    if (d.x0 < -0x3fffffff || d.x0 > 0x3fffffff || d.y0 < -0x3fffffff || d.y0 > 0x3fffffff) return TFM_EUNSUPPORTED;

    // fast mode, Gaussian, a map with a closed-form Jacobian: the lean kernels of tfm_jacobian_fast.cu
    {
        int rc = TFM_OK;
        if (jacobian_fast_dispatch(a, st, &rc)) return rc;
    }

    JacParams P;
    P.tex = 0; P.tex_ok = 0;
    P.src = s.data; P.dst = d.data; P.src_pitch = s.pitch; P.dst_pitch = d.pitch;
    P.src_plane = s.plane_pitch; P.dst_plane = d.plane_pitch; P.n_planes = (int)s.n_planes;
    P.src_h = (int)s.h; P.src_w = (int)s.w; P.src_x0 = (int)s.x0; P.src_y0 = (int)s.y0;
    P.full_h = (int)s.full_h; P.full_w = (int)s.full_w;
    P.dst_h = (int)d.h; P.dst_w = (int)d.w; P.dst_x0 = d.x0; P.dst_y0 = d.y0;
    P.grid_h = (int)d.full_h; P.grid_w = (int)d.full_w;
    P.bound_x = a->bound_x; P.bound_y = a->bound_y; P.filter = a->filter;
    P.flags = (a->flags & 0xff) | ((a->tuning & 512) ? 0x100 : 0) | ((a->tuning & 2097152) ? 0x200 : 0) |
              ((a->tuning & 4194304) ? 0x400 : 0);
    P.src_f64 = s.dtype == TFM_F64; P.dst_f64 = d.dtype == TFM_F64;
    P.oblur = a->oblur; P.pad_pow = a->pad_pow; P.jump_thresh = a->jump_thresh;
    // helpers.pyx:1612-1615: padval 1 for tent/Hanning/Tukey, 3 for Lanczos/Gaussian
    P.padval = (a->filter == TFM_FILT_LANCZOS || a->filter == TFM_FILT_GAUSSIAN) ? 3.0 : 1.0;
    P.max_reg = a->max_reg;
    P.status = a->status_dev;
    P.chain.pad = 0;
    if (a->precision == TFM_PREC_FAST32) {
        P.chain.n = fuse_chain_fast(a->chain, a->n_ops, P.chain.ops);       // fast mode: peephole fusion
    } else {
        P.chain.n = a->n_ops;
        for (int i = 0; i < a->n_ops; ++i) P.chain.ops[i] = a->chain[i];
    }

    // constant-Jacobian path: fast mode, Gaussian, an all-affine chain, and a jump threshold under
    // which a constant |J|^2 can never be flagged (ratio of neighbouring cells is 1 up to rounding)
    P.aff_const = 0;
    P.aff_half = 0;
    if (a->precision == TFM_PREC_FAST32 && a->filter == TFM_FILT_GAUSSIAN && !(a->tuning & 512) &&
        (a->jump_thresh == 0.0 || a->jump_thresh >= 1.001)) {
        bool lin = true;
        for (int i = 0; i < a->n_ops; ++i) lin = lin && is_linear_like(a->chain[i]);
        if (lin) {
            compose_linear_run(a->chain, 0, a->n_ops, P.aff, P.aff + 4);
            const double ja = P.aff[0], jb = P.aff[1], jc = P.aff[2], jd = P.aff[3];
            // Q = f(J J^T), f(lambda) = (padval + lambda^(p/2))^(-2/p): the fp64 twin of gaussian_quadratic_f32
            const double g00 = ja * ja + jb * jb, g01 = ja * jc + jb * jd, g11 = jc * jc + jd * jd;
            const double mean = 0.5 * (g00 + g11), dif = 0.5 * (g00 - g11);
            const double rad = sqrt(dif * dif + g01 * g01);
            const double lp = mean + rad, lm = fmax(mean - rad, 0.0);
            auto f = [&](double lam) { return pow(P.padval + pow(lam, 0.5 * P.pad_pow), -2.0 / P.pad_pow); };
            const double fp = f(lp), fm = f(lm);
            const double beta = (lp - lm > 0.0) ? (fp - fm) / (lp - lm) : 0.0;
            const double q00 = fp - beta * (rad - dif), q01 = beta * g01, q11 = fp - beta * (rad + dif);
            const double smin = sqrt(fp);                       // the smaller padded inverse singular value
            double reg = 2.0 * ceil(P.oblur / smin);            // helpers.pyx:1666
            if (P.max_reg > 0 && reg > (double)P.max_reg) reg = (double)P.max_reg;
            if (isfinite(q00) && isfinite(q01) && isfinite(q11) && isfinite(reg) && reg >= 0.0 && reg < 2.0e9 &&
                isfinite(P.aff[4]) && isfinite(P.aff[5])) {
                P.aff_const = 1;
                P.aff_half = (int)(reg / 2.0);
                P.aff_q[0] = (float)q00; P.aff_q[1] = (float)q01; P.aff_q[2] = (float)q11;
            }
        }
    }

    // warp patch of the footprint phase: 8 x 4 unless the map is affine (16 x 2); tuning bits 1024 /
    // 2048 force 32 x 1 / 4 x 8 for A/B runs
    P.lw = P.aff_const ? 4 : 3;
    if (a->tuning & 1024) P.lw = 5;
    if (a->tuning & 2048) P.lw = 2;

    const int sph = chain_sph_level(P.chain);
    // tile height: 32 by default (4 pixels per thread); tuning bits 8 / 32 select 16 / 8 for A/B runs
    const int jty = sph ? 8 : ((a->tuning & 32) ? 8 : ((a->tuning & 8) ? 16 : 32));
    dim3 grid((unsigned)((P.dst_w + JTX - 1) / JTX), (unsigned)((P.dst_h + jty - 1) / jty),
              (unsigned)P.n_planes);
    if (grid.y > 65535u) return TFM_EUNSUPPORTED;
    dim3 block(JTX, 8);
    const char *name;
#define TFM_JLAUNCH(AA, ST, DT, RT)                                                          \
    do {                                                                                     \
        if (sph == 2) k_jacobian<AA, ST, DT, RT, 8, 2><<<grid, block, 0, st>>>(P);            \
        else if (sph == 1) k_jacobian<AA, ST, DT, RT, 8, 1><<<grid, block, 0, st>>>(P);       \
        else if (jty == 8) k_jacobian<AA, ST, DT, RT, 8, 0><<<grid, block, 0, st>>>(P);   \
        else if (jty == 16) k_jacobian<AA, ST, DT, RT, 16, 0><<<grid, block, 0, st>>>(P); \
        else k_jacobian<AA, ST, DT, RT, 32, 0><<<grid, block, 0, st>>>(P);                \
    } while (0)
    if (a->precision == TFM_PREC_EXACT64) {
        if (s.dtype == TFM_F32 && d.dtype == TFM_F64) TFM_JLAUNCH(Exact, float, double, double);
        else if (s.dtype == TFM_F64 && d.dtype == TFM_F64) TFM_JLAUNCH(Exact, double, double, double);
        else return TFM_EUNSUPPORTED;
        name = "k_jacobian<exact64>";
    } else if (a->precision == TFM_PREC_FAST32) {
        if (s.dtype != TFM_F32 || d.dtype != TFM_F32) return TFM_EUNSUPPORTED;
        // Gaussian taps through the texture unit where a pitch-linear texture can be laid over the WHOLE source
        // (one plane, geometry within the 2-D texture limits); tuning bits 4096 / 32768 keep plain loads
        int tex_slot = -1;
        if (a->filter == TFM_FILT_GAUSSIAN && !(a->tuning & (4096 | 32768)) && s.n_planes == 1 && s.x0 == 0 && s.y0 == 0 &&
            s.full_h == s.h && s.full_w == s.w && s.w <= 131072 && s.h <= 65000 && (s.pitch % 32) == 0 &&
            (reinterpret_cast<uintptr_t>(s.data) % 512) == 0 &&
            get_texture(s.data, (int)s.w, (int)s.h, s.pitch, &P.tex, &tex_slot))
            P.tex_ok = (a->bound_x == TFM_B_TRUNCATE && a->bound_y == TFM_B_TRUNCATE) ? 2 : 1;
        TFM_JLAUNCH(Fast, float, float, float);
        texture_used(tex_slot, st);
        name = P.tex_ok ? "k_jacobian<fast32,tex>" : "k_jacobian<fast32>";
    } else {
        return TFM_EVALUE;
    }
#undef TFM_JLAUNCH
    count_launch();
    set_last_kernel(name);
    return cudaGetLastError() == cudaSuccess ? TFM_OK : TFM_ECUDA;
}

}  // namespace tfm
