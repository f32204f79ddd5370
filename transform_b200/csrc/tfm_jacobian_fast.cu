// tfm_jacobian_fast.cu -- K2, fast mode (fp32), Gaussian filter: the lean closed-form kernels.
//
// Same algorithm as tfm_jacobian.cu (interpND_grid(..., antialias=True), helpers.pyx:1293-1530:
// jacobian() + interpND_jacobian() with the Gaussian of helpers.pyx:1734-1735), for the two map
// families whose Jacobian has a closed form, so that nothing of the reference's finite-difference
// machinery (coordinate halo tile, cell differences, jump flags) has to exist:
//   k_jac_gauss_const<HALF>       the whole chain is affine: J, the quadratic form of the Gaussian, the
//                                 footprint size and the recurrence constants are launch constants
//   k_jac_gauss_polar             the chain is one fused polar op [affine, Radial^-1, affine] (the polar /
//                                 log-polar unwrap): coordinates and J per pixel from per-tile angle tables
//   k_jac_gauss_polar_rows        ... whose radius is constant along output rows: eigenvalues, padding and
//                                 footprint size per tile ROW, a pixel only rotates the row's quadratic form
//   k_jac_gauss_polar_rows_pair   ... and whose rows map <= ~1 source pixel apart: two output rows share each
//                                 set of texture fetches (the kernel of bench.py's Jacobian call)
// All are chosen by the HOST (jacobian_fast_dispatch): the conditions under which the closed form
// stands for the reference's +-1-pixel central difference and under which jump_detect cannot flag a
// cell are properties of the map and the output grid, not of a tile -- and every row block of a sharded
// job must take the same kernel.
//
// Why a second translation unit: ncu on k_jacobian<Fast> (8192^2 polar unwrap, profiles/r1_kernels.md)
// showed 646 instructions per pixel at 61 % issue utilisation, of which only ~200 were footprint taps;
// ~350 were per-pixel prologue (fp64 <-> fp32 conversions on the XU pipe, 64-bit index arithmetic,
// paths for every filter / boundary / precision sharing one register allocation) and ~60 per-tile
// tests executed by every thread.  Here every constant arrives as the type it is used in, indices
// are 32-bit, floor/fraction is fixed point (floor_frac_fast), and pixels whose footprint is not
// wholly inside the staged source take one out-of-line slow path.  The measured history of the polar
// unwrap at 8192^2 (2.05 -> 1.04 ms) is in DESIGN.md section 4.
#include "tfm_jacobian_common.cuh"

namespace tfm {

constexpr int FT = 32;                      // output tile edge (256 threads, 4 pixels per thread)

struct JacFastParams {
    const float *src;
    float *dst;
    long long src_plane, dst_plane;          // elements between planes
    int src_pe, dst_pe;                      // row pitch in elements
    int src_h, src_w, src_x0, src_y0, full_h, full_w;
    int dst_h, dst_w, dst_x0, dst_y0;
    int grid_h, grid_w;                      // full output grid (edge rule of helpers.pyx:1244-1247)
    int bound_x, bound_y;
    int flux;                                // phot='flux': scale by |det J| (SVD.pyx:285-286)
    int lw;                                  // log2 of the warp patch width
    int quads;                               // lanes 4k..4k+3 cover a 2 x 2 block (texture requests are per quad)
    int max_half;                            // cap on the footprint half-width (max_reg / 2)
    int *status;
    cudaTextureObject_t tex;                 // point-sampled, border = 0 (tap mode 2)
    int tex_edges;                           // both bounds truncate: the texture's border IS the reference's truncated
                                             // tap wherever the staged rectangle ends where the image ends (1);
                                             // 2: the whole image is staged, so that is everywhere
    float qscale;                            // log2(e) / oblur^2: exp(-q/oblur^2) = exp2(-q * qscale)
    float ob;                                // oblur
    float padval, half_pow, m2_over_pow;     // padding (helpers.pyx:1661-1662): pad_pow / 2, -2 / pad_pow
    double oblur_d, padval_d, pad_pow_d;     // for the exact footprint-size decision
    // ---- constant-Jacobian kernel ----
    double aff[6];                           // source = A . grid + b
    float Aq, Bq, Cq;                        // quadratic form, already times qscale
    float cA, cB, cC;                        // exp2(-2 Aq), exp2(-2 Bq), exp2(-2 Cq)
    float det;                               // |det A|
    int half;
    // ---- polar kernel ----
    double pa[6];                            // A1 (row-major), b1: (theta arg, rho) = A1 . grid + b1
    double coeff, r0, ox, oy;                // ang_coeff, r0 (log-polar), origin
    double a2[6];                            // trailing affine map A2, b2
    int ccw, lg, has_a2;
    float tXf, tYf, p2f, p3f, a2f[4];
};

// ---------------------------------------------------------------------------------------------
// footprint accumulation
// ---------------------------------------------------------------------------------------------
// (2*HALF+1)^2 Gaussian taps of a footprint wholly inside the staged source.  Weights by recurrence:
// along a row w(a+1) = w(a) r(a), r(a+1) = r(a) cA; from row to row w0 *= rb, rb *= cC, r0 *= cB.
// Rows are processed in PAIRS with Blackwell's packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2): lane
// .x carries row j, lane .y row j+1, and the row-to-row state (first weights W, first ratios R, and the
// two-row weight step RB2) is packed the same way -- a pair of taps costs 2 loads + 4 packed
// instructions, a pair of rows 3 packed instructions of bookkeeping.  The odd last row is scalar.
template <int HALF>
__device__ __forceinline__ void gauss_taps2(const float *__restrict__ p0, int pe, float w0, float r0, float rb,
                                            float cA, float cB, float cC, float &val, float &wgt)
{
    constexpr int N = 2 * HALF + 1;
    const float2 cA2 = make_float2(cA, cA);
    const float cC2 = cC * cC, cB2 = cB * cB;
    const float2 CB2 = make_float2(cB2, cB2), C4 = make_float2(cC2 * cC2, cC2 * cC2);
    const float rb1 = rb * cC;
    float2 W = make_float2(w0, w0 * rb);                     // first weights of rows j, j+1
    float2 R = make_float2(r0, r0 * cB);                     // first ratios of rows j, j+1
    float2 RB2 = make_float2(rb * rb * cC, rb1 * rb1 * cC);  // w0(j+2) / w0(j), w0(j+3) / w0(j+1)
    float2 val2 = make_float2(0.0f, 0.0f), wgt2 = make_float2(0.0f, 0.0f);
#pragma unroll(HALF <= 2 ? HALF : 1)
    for (int j = 0; j < HALF; ++j) {
        const float *p = p0 + (2 * j) * pe, *q = p + pe;
        float2 w = W, r = R;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const float2 v = make_float2(__ldg(p + k), __ldg(q + k));
            val2 = __ffma2_rn(v, w, val2);
            wgt2 = __fadd2_rn(wgt2, w);
            w = __fmul2_rn(w, r);
            r = __fmul2_rn(r, cA2);
        }
        W = __fmul2_rn(W, RB2);
        RB2 = __fmul2_rn(RB2, C4);
        R = __fmul2_rn(R, CB2);
    }
    {   // last (odd) row: its first weight and ratio are lane .x of the advanced state
        const float *p = p0 + (2 * HALF) * pe;
        float w = W.x, r = R.x;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            val = fmaf(__ldg(p + k), w, val);
            wgt += w;
            w *= r;
            r *= cA;
        }
    }
    val += val2.x + val2.y;
    wgt += wgt2.x + wgt2.y;
}

// The same footprint with 8-byte loads.  ncu on the row-pair form (profiles/r2_kernels.md): the kernel is
// bound by the L1TEX sector stage, not by instruction issue -- one 4-byte tap load of a warp touches
// ~13 sectors (32 footprints spread along an arc), and there are (2*HALF+1)^2 of them per pixel.  An
// 8-byte load of two horizontally adjacent taps touches the SAME sectors, so loading column PAIRS
// nearly halves the sector traffic.  8-byte alignment: the window starts on the even column at or
// below the footprint's first column and spans N+1 columns; the one column outside the reference's
// square (the first when the start was odd, else the last) is selected away (value AND weight: a NaN
// there must not leak).  Packed lanes are the two columns of a pair: w(c+2) = w(c) r(c) r(c+1), so
// with R2 = (r(c)^2 cA, r(c+1)^2 cA) the pair state advances as W2 *= R2, R2 *= cA^4.
template <int HALF>
__device__ __forceinline__ void gauss_taps_cp(const float *__restrict__ p0, int pe, bool first_out, float w0, float r0,
                                              float rb, float cA, float cB, float cC, float &val, float &wgt)
{
    constexpr int N = 2 * HALF + 1, NP = HALF + 1;
    const float cA2 = cA * cA, cA4s = cA2 * cA2;
    const float2 CA4 = make_float2(cA4s, cA4s);
    const float2 CA13 = make_float2(cA, cA * cA2);           // R2 = r0^2 * (cA, cA^3)
    float2 WR = make_float2(w0, r0);                         // (first weight, first ratio) of the row
    float2 STEP = make_float2(rb, cB);                       // row-to-row: w0 *= rb, r0 *= cB; rb *= cC
    const float2 CC1 = make_float2(cC, 1.0f);
    float2 val2 = make_float2(0.0f, 0.0f), wgt2 = make_float2(0.0f, 0.0f);
#pragma unroll(HALF <= 2 ? N : 2)
    for (int j = 0; j < N; ++j) {
        const float2 *p = reinterpret_cast<const float2 *>(p0 + j * pe);
        const float2 t = __fmul2_rn(WR, make_float2(WR.y, WR.y));        // (w0 r0, r0^2)
        float2 W2 = make_float2(WR.x, t.x);
        float2 R2 = __fmul2_rn(make_float2(t.y, t.y), CA13);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            float2 v = __ldg(p + i);
            float2 W = W2;
            if (i == 0) { v.x = first_out ? 0.0f : v.x; W.x = first_out ? 0.0f : W.x; }
            if (i == NP - 1) { v.y = first_out ? v.y : 0.0f; W.y = first_out ? W.y : 0.0f; }
            val2 = __ffma2_rn(v, W, val2);
            wgt2 = __fadd2_rn(wgt2, W);
            if (i + 1 < NP) {
                W2 = __fmul2_rn(W2, R2);
                R2 = __fmul2_rn(R2, CA4);
            }
        }
        WR = __fmul2_rn(WR, STEP);
        STEP = __fmul2_rn(STEP, CC1);
    }
    val += val2.x + val2.y;
    wgt += wgt2.x + wgt2.y;
}

// The same footprint through the texture unit (tap mode 2).  One TLD4 returns the 2 x 2 block
// (c, c+1) x (r, r+1) of every lane with no alignment rule, so the window starts exactly on the
// footprint; it spans (N+1) x (N+1) texels in (HALF+1)^2 fetches, and the column and row beyond the
// reference's square (always the LAST ones) are simply not accumulated -- no selects.  Out-of-range
// texels read 0 (border mode), which is the reference's truncated tap (weight counted, value skipped,
// helpers.pyx:1740 vs 1793-1796): with truncate boundaries NO pixel needs the slow path.
// ncu, 8192^2 polar unwrap (profiles/r2_kernels.md): the 8-byte LDG form is bound by the LSU data pipe --
// 3.5 wavefronts per request (one per cache line the 32 footprints of a warp touch), 2 taps per lane;
// the texture path returns 4 taps per lane per request through its own filter/write-back stage.
// (TexRows, tex_row_pair, gauss_taps_tex: tfm_jacobian_common.cuh -- the general kernel of tfm_jacobian.cu uses them too)

template <int TM>
__device__ __forceinline__ void gauss_taps_any(cudaTextureObject_t tex, int xs, int ys, int half, const float *p0, int pe, bool first_out,
                                               float w0, float r0, float rb, float cA, float cB, float cC,
                                               float &val, float &wgt)
{
#define TFM_TAPS(H)                                                                           \
    if (TM == 2) gauss_taps_tex<H>(tex, xs, ys, w0, r0, rb, cA, cB, cC, val, wgt);             \
    else if (TM == 1) gauss_taps_cp<H>(p0, pe, first_out, w0, r0, rb, cA, cB, cC, val, wgt);   \
    else gauss_taps2<H>(p0, pe, w0, r0, rb, cA, cB, cC, val, wgt)
    switch (half) {
    case 1: TFM_TAPS(1); break;
    case 2: TFM_TAPS(2); break;
    case 3: TFM_TAPS(3); break;
    case 4: TFM_TAPS(4); break;
    case 5: TFM_TAPS(5); break;
    default: TFM_TAPS(6); break;
    }
#undef TFM_TAPS
}

// Out-of-line path for the pixels the recurrence cannot serve: footprint crossing an image (or staged
// rectangle) edge, half-width 0 or > 6, weights too small for the recurrence, absurd coordinates.
// The reference's loop as coded (helpers.pyx:1678-1800, repairs R7/R9): a truncated tap adds its
// weight but not its value; 'forbid' raises.  One exp2 per tap.
struct SlowResult { float res; int state; };       // by value: reference arguments would live on the stack

__device__ __noinline__ SlowResult gauss_slow(const JacFastParams &P, const float *src_base, double px, double py,
                                              float Aq, float Bq, float Cq, int half)
{
    long long xs, ys;
    float xf_of, yf_of;
    if ((fabs(px) < TFM_SANE_COORD) && (fabs(py) < TFM_SANE_COORD)) {
        // the same (2^-23-quantised) floor and fraction as the recurrence path: which pixels take this
        // path must not move a footprint window
        int xi, yi;
        floor_frac_fast(px, xi, xf_of);
        floor_frac_fast(py, yi, yf_of);
        xs = xi; ys = yi;
    } else {
        const double fxs = floor(px), fys = floor(py);
        xs = (long long)fxs; ys = (long long)fys;
        xf_of = (float)(px - fxs); yf_of = (float)(py - fys);
    }
    float val = 0.0f, wgt = 0.0f;
    int state = 0;
    for (int yr = -half; yr <= half; ++yr) {
        const float y0 = (float)yr + yf_of;
        int sty = 0;
        const int iy0 = jbound(ys + yr, P.full_h, P.bound_y, sty);
        for (int xr = -half; xr <= half; ++xr) {
            const float x0 = (float)xr + xf_of;
            const float filt = fast_ex2(-((Aq * x0 + 2.0f * Bq * y0) * x0 + Cq * y0 * y0));
            wgt += filt;
            int st = sty;
            int ix = jbound(xs + xr, P.full_w, P.bound_x, st);
            state |= st & 2;
            if (st == 0) {
                ix -= P.src_x0;
                const int iy = iy0 - P.src_y0;
                if ((unsigned)ix < (unsigned)P.src_w && (unsigned)iy < (unsigned)P.src_h)
                    val = fmaf(__ldg(src_base + (long long)iy * P.src_pe + ix), filt, val);
                else
                    state |= 4;
            }
        }
    }
    SlowResult out;
    out.res = (wgt > 0.0f) ? val * fast_rcp(wgt) : 0.0f;
    out.state = state;
    return out;
}

// The footprint half-width is ceil(oblur / smin'): when fp32 cannot decide the ceil(), that one number
// is redone in fp64 from the same Jacobian with the reference's own SVD and padding formulas.
__device__ __noinline__ float exact_half_ratio(const JacFastParams &P, float j00, float j01, float j10, float j11)
{
    const double Ji[4] = {(double)j00, (double)j01, (double)j10, (double)j11};
    double U[4], V[4], sd[2];
    svd2x2_dev(Ji, U, sd, V);
    const double s0 = exp(-log(P.padval_d + exp(log(sd[0]) * P.pad_pow_d)) / P.pad_pow_d);
    const double s1 = exp(-log(P.padval_d + exp(log(sd[1]) * P.pad_pow_d)) / P.pad_pow_d);
    return (float)ceil(P.oblur_d / fmin(s0, s1));
}

// fp32 twin of gaussian_quadratic_f32 (tfm_jacobian_common.cuh) taking the Jacobian as floats
__device__ __forceinline__ void gaussian_quadratic(float a, float b, float c, float d, float padval, float half_pow,
                                                   float m2p, float &q00, float &q01, float &q11, float &inv_smin)
{
    const float g00 = a * a + b * b, g01 = a * c + b * d, g11 = c * c + d * d;
    const float mean = 0.5f * (g00 + g11), dif = 0.5f * (g00 - g11);
    const float r2 = dif * dif + g01 * g01;
    const float rad = (r2 > 0.0f) ? r2 * fast_rsqrt(r2) : 0.0f;
    const float lp = mean + rad, lm = fmaxf(mean - rad, 0.0f);
    const float fp = pad_inv_sq(lp, padval, half_pow, m2p), fm = pad_inv_sq(lm, padval, half_pow, m2p);
    const float dl = lp - lm;
    const float beta = (dl > 0.0f) ? (fp - fm) * fast_rcp(dl) : 0.0f;
    q00 = fp - beta * (rad - dif);
    q01 = beta * g01;
    q11 = fp - beta * (rad + dif);
    inv_smin = fast_rsqrt(fp);
}

// Footprint set-up of one pixel: what the tap loop needs, 9 words.  off < 0: nothing left to do (the
// pixel was all-out, or went through the slow path; `res` holds its value).
struct TapRec {
    int off;               // element offset of the window's first (aligned) tap inside the plane; tap mode 2:
                           // x of the window's first texel + 2^20 (kept non-negative: < 0 means "no taps")
    int ys;                // tap mode 2: y of the window's first texel
    int half;              // half-width | first_out << 8
    float w0, r0, rb;      // weight at the window's first tap, its ratio along the row, and to the next row
    float res;             // value of a pixel that needs no taps
};

// Classifies the pixel: 1 = the recurrence can serve it (floor / fraction of the source position in xi, yi,
// xf, yf); 0 = finished here (`res` holds its value: all-out, or the out-of-line slow path).
template <int TM>
__device__ __forceinline__ int gauss_classify(const JacFastParams &P, const float *__restrict__ src_base, double px, double py,
                                              float Aq, float Bq, float Cq, int half, int &xi, int &yi, float &xf, float &yf,
                                              float &res, int &state, int fit = -1)
{
    constexpr bool CP = (TM == 1);
    res = 0.0f;
    const bool sane = (fabs(px) < TFM_SANE_COORD) && (fabs(py) < TFM_SANE_COORD);
    bool fast = false;
    xi = 0; yi = 0; xf = 0.0f; yf = 0.0f;
    if (sane) {
        floor_frac_fast(px, xi, xf);
        floor_frac_fast(py, yi, yf);
        // a footprint entirely beyond a truncated edge: every tap is skipped, the pixel is 0 whatever the
        // weights are (helpers.pyx:1793-1800)
        // (not when the other axis forbids: the reference raises from inside the same loop)
        if (P.bound_x != TFM_B_FORBID && P.bound_y != TFM_B_FORBID &&
            ((P.bound_x == TFM_B_TRUNCATE && (xi + half < 0 || xi - half >= P.full_w)) ||
             (P.bound_y == TFM_B_TRUNCATE && (yi + half < 0 || yi - half >= P.full_h))))
            return 0;
        const int ox = xi - half - P.src_x0, oy = yi - half - P.src_y0;
        const float hf1 = (float)(half + 1);
        bool reach_ok;
        if (TM == 2 && P.tex_edges == 2) {
            // texture taps, truncate boundaries, the whole image staged: the texture's border IS the
            // reference's truncated tap -- nothing to classify
            reach_ok = true;
        } else {
            // whole footprint (plus the alignment column of the 8-byte form) inside the staged source
            const int limx = P.src_w - 2 * half - (CP ? 1 : 0), limy = P.src_h - 2 * half;
            reach_ok = limx > 0 && limy > 0 && (unsigned)ox < (unsigned)limx && (unsigned)oy < (unsigned)limy;
            // texture taps on a staged rectangle: beyond it the texture reads 0, which is right exactly where
            // "beyond the rectangle" means "beyond the image" -- i.e. when the part of the footprint that
            // lies inside the image lies inside the rectangle
            if (TM == 2 && P.tex_edges && !reach_ok) {
                const int wx0 = max(xi - half, 0), wx1 = min(xi + half, P.full_w - 1);
                const int wy0 = max(yi - half, 0), wy1 = min(yi + half, P.full_h - 1);
                reach_ok = wx0 >= P.src_x0 && wx1 < P.src_x0 + P.src_w && wy0 >= P.src_y0 && wy1 < P.src_y0 + P.src_h;
            }
        }
        // fit: the caller already knows whether the recurrence may run (size 1..6, exponents in range)
        fast = reach_ok && (fit >= 0 ? fit != 0
                                     : (half >= 1 && half <= 6 && (Aq + 2.0f * fabsf(Bq) + Cq) * (hf1 * hf1) < 100.0f));
    } else if (!(fabs(px) < 4.0e18) || !(fabs(py) < 4.0e18)) {
        return 0;                                                // NaN / inf (as tfm_jacobian.cu)
    }
    if (!fast) {
        const SlowResult sr = gauss_slow(P, src_base, px, py, Aq, Bq, Cq, half);
        state |= sr.state;
        res = sr.res;
        return 0;
    }
    return 1;
}

// Classifies the pixel and prepares the recurrence.  Pixels the recurrence cannot serve are finished
// here, out of line.
template <int TM>
__device__ __forceinline__ TapRec gauss_setup(const JacFastParams &P, const float *__restrict__ src_base, double px, double py,
                                              float Aq, float Bq, float Cq, int half, int &state, int fit = -1)
{
    TapRec rec;
    constexpr bool CP = (TM == 1);
    rec.off = -1; rec.ys = 0; rec.half = half; rec.w0 = 0.0f; rec.r0 = 0.0f; rec.rb = 0.0f; rec.res = 0.0f;
    int xi, yi;
    float xf, yf;
    if (!gauss_classify<TM>(P, src_base, px, py, Aq, Bq, Cq, half, xi, yi, xf, yf, rec.res, state, fit)) return rec;
    const int ox = xi - half - P.src_x0, oy = yi - half - P.src_y0;
    const bool first_out = CP && (ox & 1);
    const float hf = (float)half;
    const float a0 = xf - hf - (first_out ? 1.0f : 0.0f), b0 = yf - hf;
    rec.w0 = fast_ex2(-((Aq * a0 + 2.0f * Bq * b0) * a0 + Cq * b0 * b0));
    rec.r0 = fast_ex2(-(Aq * (2.0f * a0 + 1.0f) + 2.0f * Bq * b0));
    rec.rb = fast_ex2(-(Cq * (2.0f * b0 + 1.0f) + 2.0f * Bq * a0));
    if (TM == 2) { rec.off = ox + (1 << 20); rec.ys = oy; }
    else rec.off = oy * P.src_pe + (CP ? (ox & ~1) : ox);
    rec.half = half | (first_out ? 256 : 0);
    return rec;
}

// HALF_CT > 0: compile-time half-width (constant-Jacobian kernel); 0: run time.
// TM: tap mode -- 0 four-byte loads (row pairs), 1 eight-byte loads (column pairs), 2 texture gather.
template <int HALF_CT, int TM>
__device__ __forceinline__ float gauss_run(const JacFastParams &P, const float *__restrict__ src_base, int off, int ys,
                                           int half_fo, float w0, float r0, float rb, float cA, float cB, float cC)
{
    constexpr int H = (HALF_CT > 0 ? HALF_CT : 1);
    const float *p0 = src_base + (TM == 2 ? 0 : off);
    const int xs = off - (1 << 20);
    const int pe = P.src_pe;
    const bool first_out = (half_fo & 256) != 0;
    float val = 0.0f, wgt = 0.0f;
    if (HALF_CT > 0) {
        if (TM == 2) gauss_taps_tex<H>(P.tex, xs, ys, w0, r0, rb, cA, cB, cC, val, wgt);
        else if (TM == 1) gauss_taps_cp<H>(p0, pe, first_out, w0, r0, rb, cA, cB, cC, val, wgt);
        else gauss_taps2<H>(p0, pe, w0, r0, rb, cA, cB, cC, val, wgt);
    } else {
        gauss_taps_any<TM>(P.tex, xs, ys, half_fo & 255, p0, pe, first_out, w0, r0, rb, cA, cB, cC, val, wgt);
    }
    return (wgt > 0.0f) ? val * fast_rcp(wgt) : 0.0f;
}

// thread -> pixel map of a 32 x 32 tile: a warp covers a (1 << lw) x (32 >> lw) patch, repeated 4 times
// 8 rows apart, so that the 32 footprints one tap load gathers from sit in a compact source patch
// With `quads`, lanes 4k .. 4k+3 -- the group the texture unit serves as one request -- form a 2 x 2
// block (1) or a 1 x 4 column (2) of output pixels instead of 4 pixels of one row (0), whichever keeps the
// four footprints closest in the source: the texture path is bound by sectors per request.
__device__ __forceinline__ void patch_xy(int lw, int quads, int tid, int &tx, int &ty)
{
    const int lane = tid & 31, wrp = tid >> 5;
    int lx, ly;
    if (quads == 2) {                                                // 1 x 4 column: patch height >= 4 (lw <= 3)
        const int q = lane & 3, rest = lane >> 2;
        lx = rest & ((1 << lw) - 1);
        ly = ((rest >> lw) << 2) + q;
    } else if (quads) {
        const int q = lane & 3, rest = lane >> 2;                    // rest: 8 quads of the patch
        const int qw = (1 << lw) >> 1;                               // quads per patch row
        lx = ((rest & (qw - 1)) << 1) + (q & 1);
        ly = ((rest / qw) << 1) + (q >> 1);
    } else {
        lx = lane & ((1 << lw) - 1);
        ly = lane >> lw;
    }
    tx = ((wrp & ((32 >> lw) - 1)) << lw) + lx;
    ty = (wrp >> (5 - lw)) * (32 >> lw) + ly;
}

// ---------------------------------------------------------------------------------------------
// constant-Jacobian kernel
// ---------------------------------------------------------------------------------------------
#ifndef TFM_K2F_MINBLOCKS
#define TFM_K2F_MINBLOCKS 4
#endif
#ifndef TFM_K2F_TWOPHASE
#define TFM_K2F_TWOPHASE 0
#endif
template <int HALF, int TM>
__global__ void __launch_bounds__(256, TFM_K2F_MINBLOCKS)
k_jac_gauss_const(const __grid_constant__ JacFastParams P)
{
    const float *const src_base = P.src + (long long)blockIdx.z * P.src_plane;
    float *const dst_base = P.dst + (long long)blockIdx.z * P.dst_plane;
    int tx, ty;
    patch_xy(P.lw, P.quads, threadIdx.x, tx, ty);
    const int X = blockIdx.x * FT + tx;
    if (X >= P.dst_w) return;
    const int Y0 = blockIdx.y * FT + ty;
    const double gxd = (double)(P.dst_x0 + X);
    double gyd = (double)(P.dst_y0 + Y0);
    int state = 0;
    float *drow = dst_base + (long long)Y0 * P.dst_pe + X;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
        if (Y0 + 8 * j >= P.dst_h) break;
        const double px = fma(P.aff[0], gxd, fma(P.aff[1], gyd, P.aff[4]));
        const double py = fma(P.aff[2], gxd, fma(P.aff[3], gyd, P.aff[5]));
        const TapRec rec = gauss_setup<TM>(P, src_base, px, py, P.Aq, P.Bq, P.Cq, HALF > 0 ? HALF : P.half, state);
        float res = rec.res;
        if (rec.off >= 0)
            res = gauss_run<HALF, TM>(P, src_base, rec.off, rec.ys, rec.half, rec.w0, rec.r0, rec.rb, P.cA, P.cB, P.cC);
        if (P.flux) res *= P.det;
        *drow = res;
        drow += 8 * P.dst_pe;
        gyd += 8.0;
    }
    if (state && P.status) atomicOr(P.status, (state & 2 ? 1 : 0) | (state & 4 ? 2 : 0));
}

// ---------------------------------------------------------------------------------------------
// analytic polar map: one pixel
// ---------------------------------------------------------------------------------------------
constexpr int PTN = FT + 2;                 // table entries: grid points G0-1 .. G0+FT (edge-copy clamp reach)

struct PolarTables {
    double2 ta[PTN], tb[PTN];               // (cos, sin)(i tX), (cos, sin)(theta(X0, Y0) + j tY)
    double ea[PTN], eb[PTN];                // log-polar: exp(i a10), r0 exp(rho(X0, Y0) + j a11)
};

// tables of the tile whose first grid point is (GX0, GY0); same formulas as polar_tables (tfm_chain.cuh)
__device__ __forceinline__ void polar_fill_tables(const JacFastParams &P, PolarTables &T, int GX0, int GY0, int tid)
{
    const double X0 = (double)(GX0 - 1), Y0d = (double)(GY0 - 1);
    const bool lg = P.lg != 0;
    if (tid < 2 * PTN) {
        const double tX = P.pa[0] * P.coeff, tY = P.pa[1] * P.coeff;
        double s, c;
        if (tid < PTN) {
            sincos((double)tid * tX, &s, &c);
            T.ta[tid] = make_double2(c, s);
            if (lg) T.ea[tid] = exp((double)tid * P.pa[2]);
        } else {
            const int j = tid - PTN;
            const double t0 = fma(P.pa[0], X0, fma(P.pa[1], Y0d, P.pa[4])) * P.coeff;
            sincos(fma((double)j, tY, t0), &s, &c);
            T.tb[j] = make_double2(c, s);
            if (lg) {
                const double rho0 = fma(P.pa[2], X0, fma(P.pa[3], Y0d, P.pa[5]));
                T.eb[j] = P.r0 * exp(fma((double)j, P.pa[3], rho0));
            }
        }
    }
}

struct PolarPix {
    double px, py;         // source position
    float Aq, Bq, Cq;      // padded quadratic form of the Gaussian, times qscale
    float fl;              // phot='flux' factor (1 otherwise)
    int half;              // footprint half-width
};

// Source position (fp64: cos/sin by angle addition from the tables), closed-form Jacobian (fp32), padded
// quadratic form and footprint size of grid point (gx, gy) = tile point (tx, ly).
__device__ __forceinline__ PolarPix polar_pixel(const JacFastParams &P, const PolarTables &T, int GX0, int GY0,
                                                int tx, int ly, int gx, int gy)
{
    PolarPix o;
    const bool lg = P.lg != 0;
    const int GW = P.grid_w, GH = P.grid_h;
    const double gxd = (double)gx, gyd = (double)gy;
    const float sgn = P.ccw ? 1.0f : -1.0f;
    const double2 ta = T.ta[tx + 1];
    const double2 tb = T.tb[ly + 1];
    double c = ta.x * tb.x - ta.y * tb.y, s = ta.y * tb.x + ta.x * tb.y;
    double r = lg ? T.ea[tx + 1] * T.eb[ly + 1] : fma(P.pa[2], gxd, fma(P.pa[3], gyd, P.pa[5]));
    const double x1 = fma(c, r, P.ox), y1 = fma(P.ccw ? s : -s, r, P.oy);
    o.px = x1; o.py = y1;
    if (P.has_a2) {
        o.px = fma(P.a2[0], x1, fma(P.a2[1], y1, P.a2[4]));
        o.py = fma(P.a2[2], x1, fma(P.a2[3], y1, P.a2[5]));
    }
    // ---- closed-form Jacobian, fp32.  Grid-edge pixels copy their inner neighbour
    // (helpers.pyx:1244-1247): there (only) cos, sin and radius are re-derived at the clamped point.
    // Inputs are rounded from the fp64 values of the pixel, which do not depend on how the grid
    // is tiled: row-block shards reproduce the full frame.
    const int jx = gx < 1 ? 1 : (gx > GW - 2 ? GW - 2 : gx);
    const int jy = gy < 1 ? 1 : (gy > GH - 2 ? GH - 2 : gy);
    if (jx != gx || jy != gy) {
        const double2 ca = T.ta[jx - (GX0 - 1)], cb = T.tb[jy - (GY0 - 1)];
        c = ca.x * cb.x - ca.y * cb.y;
        s = ca.y * cb.x + ca.x * cb.y;
        r = lg ? T.ea[jx - (GX0 - 1)] * T.eb[jy - (GY0 - 1)]
               : fma(P.pa[2], (double)jx, fma(P.pa[3], (double)jy, P.pa[5]));
    }
    const float cf = (float)c, sf = (float)s, rf = (float)r;
    const float drX = lg ? rf * P.p2f : P.p2f, drY = lg ? rf * P.p3f : P.p3f;
    const float rs = rf * sf, rc = rf * cf;
    // x1 = r cos + ox, y1 = sgn r sin + oy
    float j00 = fmaf(cf, drX, -rs * P.tXf), j01 = fmaf(cf, drY, -rs * P.tYf);
    float j10 = sgn * fmaf(sf, drX, rc * P.tXf), j11 = sgn * fmaf(sf, drY, rc * P.tYf);
    if (P.has_a2) {
        const float k00 = P.a2f[0] * j00 + P.a2f[1] * j10, k01 = P.a2f[0] * j01 + P.a2f[1] * j11;
        const float k10 = P.a2f[2] * j00 + P.a2f[3] * j10, k11 = P.a2f[2] * j01 + P.a2f[3] * j11;
        j00 = k00; j01 = k01; j10 = k10; j11 = k11;
    }
    // ---- padded quadratic form and footprint size ----
    float q00, q01, q11, inv_smin;
    gaussian_quadratic(j00, j01, j10, j11, P.padval, P.half_pow, P.m2_over_pow, q00, q01, q11, inv_smin);
    float ratio = P.ob * inv_smin;
    if (fabsf(ratio - rintf(ratio)) < 4e-5f * ratio) ratio = exact_half_ratio(P, j00, j01, j10, j11);
    int half = (ratio < 1.0e9f) ? __float2int_ru(ratio) : 1000000000;
    o.half = min(half, P.max_half);
    o.Aq = q00 * P.qscale; o.Bq = q01 * P.qscale; o.Cq = q11 * P.qscale;
    o.fl = P.flux ? fabsf(j00 * j11 - j01 * j10) : 1.0f;
    return o;
}

// ---------------------------------------------------------------------------------------------
// analytic polar kernel
// ---------------------------------------------------------------------------------------------
// One thread = 4 pixels of a 32 x 32 tile, 8 rows apart.  (A two-phase form -- prologue records of the 4
// pixels parked in shared memory, then the tap loops -- was built and measured no faster; removed.)
template <int TM>
__global__ void __launch_bounds__(256, TFM_K2F_MINBLOCKS)
k_jac_gauss_polar(const __grid_constant__ JacFastParams P)
{
    __shared__ PolarTables T;
    const int tid = threadIdx.x;
    const float *const src_base = P.src + (long long)blockIdx.z * P.src_plane;
    const int GX0 = P.dst_x0 + (int)blockIdx.x * FT, GY0 = P.dst_y0 + (int)blockIdx.y * FT;
    polar_fill_tables(P, T, GX0, GY0, tid);
    __syncthreads();

    int tx, ty;
    patch_xy(P.lw, P.quads, tid, tx, ty);
    const int X = blockIdx.x * FT + tx;
    if (X >= P.dst_w) return;
    const int Yt = blockIdx.y * FT + ty;
    int state = 0;
    const int gx = P.dst_x0 + X;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
        const int ly = ty + 8 * j;
        if (Yt + 8 * j >= P.dst_h) break;
        const PolarPix q = polar_pixel(P, T, GX0, GY0, tx, ly, gx, P.dst_y0 + Yt + 8 * j);
        const TapRec rec = gauss_setup<TM>(P, src_base, q.px, q.py, q.Aq, q.Bq, q.Cq, q.half, state);
        float res = rec.res * q.fl;
        if (rec.off >= 0)
            res = q.fl * gauss_run<0, TM>(P, src_base, rec.off, rec.ys, rec.half, rec.w0, rec.r0, rec.rb,
                                          fast_ex2(-2.0f * q.Aq), fast_ex2(-2.0f * q.Bq), fast_ex2(-2.0f * q.Cq));
        P.dst[(long long)blockIdx.z * P.dst_plane + (long long)(Yt + 8 * j) * P.dst_pe + X] = res;
    }
    if (state && P.status) atomicOr(P.status, (state & 2 ? 1 : 0) | (state & 4 ? 2 : 0));
}

// ---------------------------------------------------------------------------------------------
// analytic polar kernel, radius constant along output rows
// ---------------------------------------------------------------------------------------------
// ncu source view of k_jac_gauss_polar at 8192^2 (profiles/r2_kernels.md): 478 instructions per pixel, 261 of
// them BEFORE the first tap -- Jacobian, eigenvalues of J J^T, padding (8 MUFU), footprint size, and the launch
// constants all of that re-reads per pixel.  In the common unwrap (no radius term in X: p2 = 0, no trailing
// affine map) the radius depends on the output row only, and J = B(theta) J0(row) with B orthogonal, so
//     Q = f(J J^T) = B f(J0 J0^T) B^T:
// eigenvalues, padding, footprint size, flux factor and the fitness of the recurrence are per-ROW numbers
// (one table entry per tile row, 34 evaluations per tile instead of 1024) and a pixel only rotates the row's
// quadratic form by its angle: 11 fp32 operations.  Sines are stored with the sense of rotation folded in
// (sin(sgn theta) = sgn sin(theta)), which also removes the sign handling from the source position.
struct PolarRowTables {
    double2 ta[PTN], tb[PTN];               // (cos, sgn sin)(i tX), (cos, sgn sin)(theta(X0, row))
    double r[PTN];                          // radius of the row
    float4 q[PTN];                          // f(J0 J0^T) x qscale: m00, sgn m01, m11; |det J0| (flux)
    int half[PTN];                          // footprint half-width; -1 - half: the recurrence is unfit for the row
};

__device__ __forceinline__ void polar_fill_row_tables(const JacFastParams &P, PolarRowTables &T, int GX0, int GY0, int tid)
{
    if (tid < 3 * PTN) {
        // same formulas as polar_fill_tables / polar_tables (tfm_chain.cuh): the source position of a pixel is
        // bit-identical with the other kernels'
        const double X0 = (double)(GX0 - 1), Y0d = (double)(GY0 - 1);
        const double sg = P.ccw ? 1.0 : -1.0;
        double sn, cs;
        if (tid < PTN) {
            sincos((double)tid * (P.pa[0] * P.coeff), &sn, &cs);
            T.ta[tid] = make_double2(cs, sg * sn);
        } else if (tid < 2 * PTN) {
            const int j = tid - PTN;
            const double t0 = fma(P.pa[0], X0, fma(P.pa[1], Y0d, P.pa[4])) * P.coeff;
            sincos(fma((double)j, P.pa[1] * P.coeff, t0), &sn, &cs);
            T.tb[j] = make_double2(cs, sg * sn);
        } else {
            const int j = tid - 2 * PTN;
            const double gyd = (double)(GY0 - 1 + j);
            const double r = P.lg ? P.r0 * exp(fma((double)j, P.pa[3], fma(P.pa[3], Y0d, P.pa[5])))
                                  : fma(P.pa[3], gyd, P.pa[5]);
            T.r[j] = r;
            // J0 = [[0, drY], [r tX, r tY]]: the Jacobian at angle 0 without the sense of rotation
            const float rf = (float)r;
            const float drY = P.lg ? rf * P.p3f : P.p3f;
            const float j10 = rf * P.tXf, j11 = rf * P.tYf;
            float q00, q01, q11, inv_smin;
            gaussian_quadratic(0.0f, drY, j10, j11, P.padval, P.half_pow, P.m2_over_pow, q00, q01, q11, inv_smin);
            float ratio = P.ob * inv_smin;
            if (fabsf(ratio - rintf(ratio)) < 4e-5f * ratio) ratio = exact_half_ratio(P, 0.0f, drY, j10, j11);
            int half = (ratio < 1.0e9f) ? __float2int_ru(ratio) : 1000000000;
            half = min(half, P.max_half);
            const float m00 = q00 * P.qscale, m01 = q01 * P.qscale, m11 = q11 * P.qscale;
            // exponents of the recurrence over the window, whatever the angle: Aq + Cq is the trace,
            // |Bq| at most the radius of the form's circle
            const float hd = 0.5f * (m00 - m11), hf1 = (float)(half + 1);
            const float bmax = sqrtf(hd * hd + m01 * m01);
            const bool fit = half >= 1 && half <= 6 && (m00 + m11 + 2.0f * bmax) * (hf1 * hf1) < 120.0f;   // weights stay normal fp32
            T.q[j] = make_float4(m00, P.ccw ? m01 : -m01, m11, P.flux ? fabsf(drY * j10) : 1.0f);
            T.half[j] = fit ? half : -1 - half;
        }
    }
}

template <int TM>
__global__ void __launch_bounds__(256, TFM_K2F_MINBLOCKS)
k_jac_gauss_polar_rows(const __grid_constant__ JacFastParams P)
{
    __shared__ PolarRowTables T;
    const int tid = threadIdx.x;
    const float *const src_base = P.src + (long long)blockIdx.z * P.src_plane;
    const int GX0 = P.dst_x0 + (int)blockIdx.x * FT, GY0 = P.dst_y0 + (int)blockIdx.y * FT;
    polar_fill_row_tables(P, T, GX0, GY0, tid);
    __syncthreads();

    int tx, ty;
    patch_xy(P.lw, P.quads, tid, tx, ty);
    const int X = blockIdx.x * FT + tx;
    if (X >= P.dst_w) return;
    const int Yt = blockIdx.y * FT + ty;
    int state = 0;
    const int gx = P.dst_x0 + X;
    const int GW = P.grid_w, GH = P.grid_h;
    const double2 ta = T.ta[tx + 1];
    const int jx = gx < 1 ? 1 : (gx > GW - 2 ? GW - 2 : gx);
    float *drow = P.dst + (long long)blockIdx.z * P.dst_plane + (long long)Yt * P.dst_pe + X;
#pragma unroll 1
    for (int j = 0; j < 4; ++j, drow += 8 * P.dst_pe) {
        const int ly = ty + 8 * j;
        if (Yt + 8 * j >= P.dst_h) break;
        const int gy = P.dst_y0 + Yt + 8 * j;
        const double2 tb = T.tb[ly + 1];
        double c = ta.x * tb.x - ta.y * tb.y, s = ta.y * tb.x + ta.x * tb.y;
        const double r = T.r[ly + 1];
        const double px = fma(c, r, P.ox), py = fma(s, r, P.oy);
        // grid-edge pixels copy the Jacobian of their inner neighbour (helpers.pyx:1244-1247)
        const int jy = gy < 1 ? 1 : (gy > GH - 2 ? GH - 2 : gy);
        int jt = ly + 1;
        if (jx != gx || jy != gy) {
            jt = jy - (GY0 - 1);
            const double2 ca = T.ta[jx - (GX0 - 1)], cb = T.tb[jt];
            c = ca.x * cb.x - ca.y * cb.y;
            s = ca.y * cb.x + ca.x * cb.y;
        }
        const float4 m = T.q[jt];
        const int ht = T.half[jt];
        const int half = ht < 0 ? -1 - ht : ht;
        const float cf = (float)c, sf = (float)s;
        const float c2 = cf * cf, s2 = sf * sf, cs2 = 2.0f * cf * sf;
        const float Aq = fmaf(c2, m.x, fmaf(s2, m.z, -cs2 * m.y));
        const float Cq = fmaf(s2, m.x, fmaf(c2, m.z, cs2 * m.y));
        const float Bq = fmaf(0.5f * cs2, m.x - m.z, (c2 - s2) * m.y);
        const TapRec rec = gauss_setup<TM>(P, src_base, px, py, Aq, Bq, Cq, half, state, ht >= 0 ? 1 : 0);
        float res = rec.res;
        if (rec.off >= 0)
            res = gauss_run<0, TM>(P, src_base, rec.off, rec.ys, rec.half, rec.w0, rec.r0, rec.rb,
                                   fast_ex2(-2.0f * Aq), fast_ex2(-2.0f * Bq), fast_ex2(-2.0f * Cq));
        *drow = res * m.w;
    }
    if (state && P.status) atomicOr(P.status, (state & 2 ? 1 : 0) | (state & 4 ? 2 : 0));
}

// ---------------------------------------------------------------------------------------------
// analytic polar kernel, per-row tables, two output rows per set of texture fetches
// ---------------------------------------------------------------------------------------------
// ncu on k_jac_gauss_polar_rows (profiles/r2_kernels.md): 339 instructions per pixel at 49 % issue
// utilisation, texture write-back 92 % busy -- a gather returns 512 bytes per warp and the path moves 32 bytes
// per clock, so the fetches ARE the run time.  Most texels are fetched twice: where one output row maps less
// than a source pixel away from the next (the radial direction of an unwrap), the (2 HALF + 1)^2 windows of the
// output pixels (X, Y) and (X, Y + 1) start at most one texel apart in each axis, so BOTH lie inside the
// (2 HALF + 2)^2 texels that (HALF + 1)^2 gathers return anyway: 28-40 % fewer texels written back per pixel.
// The shared window starts at the smaller of the two window origins per axis.  A pixel whose own window
// starts one column (row) later runs its recurrence from the shared origin all the same and leaves the first
// column (row) out; one whose window starts on the shared origin leaves the last out.  Left-out taps are
// selected away value AND weight: a NaN outside a pixel's own square must not reach it (the reference's window
// is the square, helpers.pyx:1678-1679).
// Pairs are the GLOBAL grid rows (2m, 2m + 1): the tile grid starts on the even row at or below the row block's
// first, and a pixel whose partner lies outside the block still computes the partner's window -- every pixel's
// arithmetic depends on its own global pair only, so any split into row blocks reproduces the full frame bit
// for bit.  A pair whose windows do not fit one shared window (different footprint sizes, one of the two not
// served by the recurrence) takes the one-pixel form.
struct PairPix {
    TexRows S;
    float2 val2, wgt2;     // packed accumulators: inner column pairs
    float val, wgt;        // scalar accumulators: first and last column pair
    bool mx, my;           // own window starts one column / row after the shared origin
};

template <int HALF>
__device__ __forceinline__ void pair_rows(const float4 (&g)[HALF + 1], PairPix &A, const bool use_a, const bool use_b)
{
    constexpr int NP = HALF + 1;
    TexRows &S = A.S;
    const float2 CA4 = bc2(S.cA4), CB2 = bc2(S.cB2), CC = bc2(S.cC);
    float2 Wa = S.W, Ra = S.R;
    float2 Wb = __fmul2_rn(S.W, S.RB), Rb = __fmul2_rn(S.R, CB2);
    {
        const float2 RB1 = __fmul2_rn(S.RB, CC);
        S.W = __fmul2_rn(Wb, RB1);
        S.R = __fmul2_rn(Rb, CB2);
        S.RB = __fmul2_rn(RB1, CC);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        // gather order: x = (c, r+1), y = (c+1, r+1), z = (c+1, r), w = (c, r)
        if (i == 0 || i == NP - 1) {
            // first / last column pair: one column always belongs to the pixel, the other only if its window starts
            // on the shared origin (first pair) or one later (last pair) -- scalar, the doubtful lane under a predicate
            // (selecting lanes of a packed pair costs a select AND a move per operand: 44 of the 173 instructions of
            // a row pair in the first form of this kernel; measured 1.079 -> 1.066 ms)
            const bool xin = (i == 0) ? !A.mx : true, yin = (i == NP - 1) ? A.mx : true;
            if (use_a) {
                if (xin) { A.val = fmaf(g[i].w, Wa.x, A.val); A.wgt += Wa.x; }
                if (yin) { A.val = fmaf(g[i].z, Wa.y, A.val); A.wgt += Wa.y; }
            }
            if (use_b) {
                if (xin) { A.val = fmaf(g[i].x, Wb.x, A.val); A.wgt += Wb.x; }
                if (yin) { A.val = fmaf(g[i].y, Wb.y, A.val); A.wgt += Wb.y; }
            }
        } else {
            if (use_a) { A.val2 = __ffma2_rn(make_float2(g[i].w, g[i].z), Wa, A.val2); A.wgt2 = __fadd2_rn(A.wgt2, Wa); }
            if (use_b) { A.val2 = __ffma2_rn(make_float2(g[i].x, g[i].y), Wb, A.val2); A.wgt2 = __fadd2_rn(A.wgt2, Wb); }
        }
        if (i + 1 < NP) {
            Wa = __fmul2_rn(Wa, Ra); Wb = __fmul2_rn(Wb, Rb);
            if (i + 2 < NP) { Ra = __fmul2_rn(Ra, CA4); Rb = __fmul2_rn(Rb, CA4); }
        }
    }
}

#ifndef TFM_K2P_DB
#define TFM_K2P_DB 4             // largest half-width served by the double-buffered form
#endif
template <int HALF>
__device__ __forceinline__ void pair_taps_tex(cudaTextureObject_t tex, int ux, int uy, PairPix &A, PairPix &B)
{
    constexpr int NP = HALF + 1;
    const float fx0 = (float)(ux + 1);
    float fy = (float)(uy + 1);
    if constexpr (HALF <= 1) {
        float4 g[NP][NP];
#pragma unroll
        for (int j = 0; j < NP; ++j)
#pragma unroll
            for (int i = 0; i < NP; ++i)
                g[j][i] = tex2Dgather<float4>(tex, fx0 + (float)(2 * i), fy + (float)(2 * j), 0);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            pair_rows<HALF>(g[j], A, j == 0 ? !A.my : true, j == NP - 1 ? A.my : true);
            pair_rows<HALF>(g[j], B, j == 0 ? !B.my : true, j == NP - 1 ? B.my : true);
        }
    } else if constexpr (HALF <= TFM_K2P_DB) {
        // double-buffered: the next row pair's gathers are issued before the current one is consumed (rows fully
        // unrolled).  With 80 registers (3 resident CTAs) 1.064 -> 1.037 ms on the 8192^2 unwrap, bit-identical
        // output; at 64 registers it spills (1.136 ms); windows of half-width 5 and 6 (TFM_K2P_DB = 5, 6) change
        // nothing on that map and spill more, so they keep the rolled loop below.
        float4 g[NP], gn[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) g[i] = tex2Dgather<float4>(tex, fx0 + (float)(2 * i), fy, 0);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            if (j + 1 < NP) {
                fy += 2.0f;
#pragma unroll
                for (int i = 0; i < NP; ++i) gn[i] = tex2Dgather<float4>(tex, fx0 + (float)(2 * i), fy, 0);
            }
            pair_rows<HALF>(g, A, j == 0 ? !A.my : true, j == NP - 1 ? A.my : true);
            pair_rows<HALF>(g, B, j == 0 ? !B.my : true, j == NP - 1 ? B.my : true);
            if (j + 1 < NP) {
#pragma unroll
                for (int i = 0; i < NP; ++i) g[i] = gn[i];
            }
        }
    } else {
        float4 g[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) g[i] = tex2Dgather<float4>(tex, fx0 + (float)(2 * i), fy, 0);
        pair_rows<HALF>(g, A, !A.my, true);
        pair_rows<HALF>(g, B, !B.my, true);
#pragma unroll 1
        for (int j = 1; j < NP - 1; ++j) {
            fy += 2.0f;
#pragma unroll
            for (int i = 0; i < NP; ++i) g[i] = tex2Dgather<float4>(tex, fx0 + (float)(2 * i), fy, 0);
            pair_rows<HALF>(g, A, true, true);
            pair_rows<HALF>(g, B, true, true);
        }
        fy += 2.0f;
#pragma unroll
        for (int i = 0; i < NP; ++i) g[i] = tex2Dgather<float4>(tex, fx0 + (float)(2 * i), fy, 0);
        pair_rows<HALF>(g, A, true, A.my);
        pair_rows<HALF>(g, B, true, B.my);
    }
}

// recurrence of one pixel of a pair, started at the shared window's origin (dx, dy in {0, 1} before its own)
__device__ __forceinline__ void pair_begin(PairPix &A, float Aq, float Bq, float Cq, float xf, float yf, int half, int dx, int dy)
{
    const float a0 = xf - (float)(half + dx), b0 = yf - (float)(half + dy);
    const float w0 = fast_ex2(-((Aq * a0 + 2.0f * Bq * b0) * a0 + Cq * b0 * b0));
    const float r0 = fast_ex2(-(Aq * (2.0f * a0 + 1.0f) + 2.0f * Bq * b0));
    const float rb = fast_ex2(-(Cq * (2.0f * b0 + 1.0f) + 2.0f * Bq * a0));
    const float cA = fast_ex2(-2.0f * Aq), cB = fast_ex2(-2.0f * Bq);
    const float cA2 = cA * cA, r02 = r0 * r0;
    A.S.W = __fmul2_rn(bc2(w0), make_float2(1.0f, r0));
    A.S.R = __fmul2_rn(bc2(r02), make_float2(cA, cA * cA2));
    A.S.RB = __fmul2_rn(bc2(rb), make_float2(1.0f, cB));
    A.S.cA4 = cA2 * cA2; A.S.cB2 = cB * cB; A.S.cC = fast_ex2(-2.0f * Cq);
    A.val2 = make_float2(0.0f, 0.0f); A.wgt2 = make_float2(0.0f, 0.0f);
    A.val = 0.0f; A.wgt = 0.0f;
    A.mx = dx != 0; A.my = dy != 0;
}

__device__ __forceinline__ float pair_end(const PairPix &A)
{
    const float val = A.val + (A.val2.x + A.val2.y), wgt = A.wgt + (A.wgt2.x + A.wgt2.y);
    return (wgt > 0.0f) ? val * fast_rcp(wgt) : 0.0f;
}

// one pixel of a pair on its own (its partner's window does not fit): the one-pixel texture form
__device__ __noinline__ float pair_single(const JacFastParams &P, float Aq, float Bq, float Cq, float xf, float yf, int half,
                                          int xs, int ys)
{
    const float a0 = xf - (float)half, b0 = yf - (float)half;
    const float w0 = fast_ex2(-((Aq * a0 + 2.0f * Bq * b0) * a0 + Cq * b0 * b0));
    const float r0 = fast_ex2(-(Aq * (2.0f * a0 + 1.0f) + 2.0f * Bq * b0));
    const float rb = fast_ex2(-(Cq * (2.0f * b0 + 1.0f) + 2.0f * Bq * a0));
    float val = 0.0f, wgt = 0.0f;
    gauss_taps_any<2>(P.tex, xs, ys, half, nullptr, 0, false, w0, r0, rb, fast_ex2(-2.0f * Aq), fast_ex2(-2.0f * Bq),
                      fast_ex2(-2.0f * Cq), val, wgt);
    return (wgt > 0.0f) ? val * fast_rcp(wgt) : 0.0f;
}

#ifndef TFM_K2P_MINBLOCKS
#define TFM_K2P_MINBLOCKS 3    // 80 registers: what the double-buffered row pairs need (4 CTAs / 64 registers: spills)
#endif
__global__ void __launch_bounds__(256, TFM_K2P_MINBLOCKS)
k_jac_gauss_polar_rows_pair(const __grid_constant__ JacFastParams P)
{
    __shared__ PolarRowTables T;
    const int tid = threadIdx.x;
    const float *const src_base = P.src + (long long)blockIdx.z * P.src_plane;
    const int ybase = P.dst_y0 & ~1;                         // global row of the tile grid's first row (even)
    const int GX0 = P.dst_x0 + (int)blockIdx.x * FT, GY0 = ybase + (int)blockIdx.y * FT;
    polar_fill_row_tables(P, T, GX0, GY0, tid);
    __syncthreads();

    const int lane = tid & 31, wrp = tid >> 5;
    int lx, pr;                                              // column and pair row inside the 8 x 8 patch
    if (P.quads == 2) { pr = lane & 3; lx = lane >> 2; }     // lanes 4k .. 4k+3: 4 pairs of one column
    else { lx = lane & 7; pr = lane >> 3; }
    const int GW = P.grid_w, GH = P.grid_h;
    int state = 0;
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
        const int patch = wrp + 8 * k;                       // 4 x 4 patches of 8 x 8 pixels
        const int tx = ((patch & 3) << 3) + lx, ly = ((patch >> 2) << 3) + 2 * pr;
        const int X = blockIdx.x * FT + tx;
        const int gx = P.dst_x0 + X, gy = GY0 + ly;          // global grid point of the pair's first pixel
        const int Y = gy - P.dst_y0;                         // its row inside the block (-1: the partner of row 0)
        if (X >= P.dst_w || Y >= P.dst_h) continue;
        const double2 ta = T.ta[tx + 1];
        const int jx = gx < 1 ? 1 : (gx > GW - 2 ? GW - 2 : gx);
        double pxy[2][2];
        float Q[2][3], fl[2];
        int hf[2], fit[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const double2 tb = T.tb[ly + e + 1];
            double c = ta.x * tb.x - ta.y * tb.y, s = ta.y * tb.x + ta.x * tb.y;
            const double r = T.r[ly + e + 1];
            pxy[e][0] = fma(c, r, P.ox); pxy[e][1] = fma(s, r, P.oy);
            const int gye = gy + e;
            const int jy = gye < 1 ? 1 : (gye > GH - 2 ? GH - 2 : gye);
            int jt = ly + e + 1;
            if (jx != gx || jy != gye) {
                jt = jy - (GY0 - 1);
                const double2 ca = T.ta[jx - (GX0 - 1)], cb = T.tb[jt];
                c = ca.x * cb.x - ca.y * cb.y;
                s = ca.y * cb.x + ca.x * cb.y;
            }
            const float4 m = T.q[jt];
            const int ht = T.half[jt];
            hf[e] = ht < 0 ? -1 - ht : ht;
            fit[e] = ht >= 0 ? 1 : 0;
            const float cf = (float)c, sf = (float)s;
            const float c2 = cf * cf, s2 = sf * sf, cs2 = 2.0f * cf * sf;
            Q[e][0] = fmaf(c2, m.x, fmaf(s2, m.z, -cs2 * m.y));
            Q[e][2] = fmaf(s2, m.x, fmaf(c2, m.z, cs2 * m.y));
            Q[e][1] = fmaf(0.5f * cs2, m.x - m.z, (c2 - s2) * m.y);
            fl[e] = m.w;
        }
        const bool in_a = Y >= 0, in_b = Y + 1 < P.dst_h;
        int xia, yia, xib, yib, sta = 0, stb = 0;
        float xfa, yfa, xfb, yfb, resa, resb;
        const int ka = gauss_classify<2>(P, src_base, pxy[0][0], pxy[0][1], Q[0][0], Q[0][1], Q[0][2], hf[0], xia, yia, xfa, yfa, resa, sta, fit[0]);
        const int kb = gauss_classify<2>(P, src_base, pxy[1][0], pxy[1][1], Q[1][0], Q[1][1], Q[1][2], hf[1], xib, yib, xfb, yfb, resb, stb, fit[1]);
        if (in_a) state |= sta;
        if (in_b) state |= stb;
        const int h = hf[0];
        const int oxa = xia - h, oya = yia - h, oxb = xib - hf[1], oyb = yib - hf[1];
        const int ux = min(oxa, oxb), uy = min(oya, oyb);
        const bool joint = ka && kb && hf[1] == h && max(oxa, oxb) - ux <= 1 && max(oya, oyb) - uy <= 1;
        if (joint) {
            PairPix A, B;
            pair_begin(A, Q[0][0], Q[0][1], Q[0][2], xfa, yfa, h, oxa - ux, oya - uy);
            pair_begin(B, Q[1][0], Q[1][1], Q[1][2], xfb, yfb, h, oxb - ux, oyb - uy);
            const int txs = ux - P.src_x0, tys = uy - P.src_y0;
            switch (h) {
            case 1: pair_taps_tex<1>(P.tex, txs, tys, A, B); break;
            case 2: pair_taps_tex<2>(P.tex, txs, tys, A, B); break;
            case 3: pair_taps_tex<3>(P.tex, txs, tys, A, B); break;
            case 4: pair_taps_tex<4>(P.tex, txs, tys, A, B); break;
            case 5: pair_taps_tex<5>(P.tex, txs, tys, A, B); break;
            default: pair_taps_tex<6>(P.tex, txs, tys, A, B); break;
            }
            resa = pair_end(A);
            resb = pair_end(B);
        } else {
            if (ka) resa = pair_single(P, Q[0][0], Q[0][1], Q[0][2], xfa, yfa, h, oxa - P.src_x0, oya - P.src_y0);
            if (kb) resb = pair_single(P, Q[1][0], Q[1][1], Q[1][2], xfb, yfb, hf[1], oxb - P.src_x0, oyb - P.src_y0);
        }
        float *d = P.dst + (long long)blockIdx.z * P.dst_plane + (long long)Y * P.dst_pe + X;
        if (in_a) d[0] = resa * fl[0];
        if (in_b) d[P.dst_pe] = resb * fl[1];
    }
    if (state && P.status) atomicOr(P.status, (state & 2 ? 1 : 0) | (state & 4 ? 2 : 0));
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// Largest ratio max|J|^2 / min|J|^2 any (FT+4)^2 neighbourhood of grid points can show for the polar
// map, from the closed form |J|^2 = g2 + t2 v^2 (x v^2-free exp form for log-polar), v = radius
// argument, affine in the grid indices: f(w) = (g2 + t2 (w + D)^2) / (g2 + t2 w^2) over the |v| the
// output region contains, D = the span of v across one neighbourhood.  f rises to
// w* = (-D + sqrt(D^2 + 4 g2 / t2)) / 2 and falls after it.
static double polar_jump_ratio_bound(const tfm_op &op, const tfm_image &d)
{
    const double coeff = op.p[6];
    const double t2 = (op.p[0] * op.p[0] + op.p[1] * op.p[1]) * coeff * coeff;
    const double g2 = op.p[2] * op.p[2] + op.p[3] * op.p[3];
    const bool lg = (op.flags & TFM_RAD_R0_NOT_NONE) != 0;
    const double D = (FT + 3) * (fabs(op.p[2]) + fabs(op.p[3]));
    double ratio;
    if (lg) {
        ratio = exp(2.0 * D);
    } else {
        const double xa = (double)d.x0 - 2.0, xb = (double)(d.x0 + d.w) + 1.0;
        const double ya = (double)d.y0 - 2.0, yb = (double)(d.y0 + d.h) + 1.0;
        const double v[4] = {op.p[2] * xa + op.p[3] * ya + op.p[5], op.p[2] * xb + op.p[3] * ya + op.p[5],
                             op.p[2] * xa + op.p[3] * yb + op.p[5], op.p[2] * xb + op.p[3] * yb + op.p[5]};
        double vmin = v[0], vmax = v[0];
        for (int k = 1; k < 4; ++k) { vmin = fmin(vmin, v[k]); vmax = fmax(vmax, v[k]); }
        const double wmax = fmax(fabs(vmin), fabs(vmax));
        const double wmin = (vmin <= 0.0 && vmax >= 0.0) ? 0.0 : fmin(fabs(vmin), fabs(vmax));
        if (!(g2 > 0.0)) return INFINITY;
        if (!(t2 > 0.0)) ratio = 1.0;
        else {
            double w = 0.5 * (-D + sqrt(D * D + 4.0 * g2 / t2));
            w = fmin(fmax(w, wmin), wmax);
            ratio = (g2 + t2 * (w + D) * (w + D)) / (g2 + t2 * w * w);
        }
    }
    if (op.flags & TFM_POLAR_HAS_A2) {
        const double E = op.p[10] * op.p[10] + op.p[11] * op.p[11] + op.p[12] * op.p[12] + op.p[13] * op.p[13];
        const double det = op.p[10] * op.p[13] - op.p[11] * op.p[12];
        const double q = sqrt(fmax(E * E - 4.0 * det * det, 0.0));
        ratio *= (E + q) / (E - q);
    }
    return ratio;
}

// Returns 1 when one of the lean kernels was launched (status in *rc), 0 when the call is outside
// their scope and the general kernel must take it.
int jacobian_fast_dispatch(const tfm_jacobian_args *a, cudaStream_t st, int *rc)
{
    const tfm_image &s = a->src, &d = a->dst;
    if (a->precision != TFM_PREC_FAST32 || a->filter != TFM_FILT_GAUSSIAN) return 0;
    if (s.dtype != TFM_F32 || d.dtype != TFM_F32) return 0;
    if (a->tuning & (512 | 8 | 32 | 8192)) return 0;              // A/B switches of the general kernel
    if ((s.pitch % 4) || (d.pitch % 4) || (s.plane_pitch % 4) || (d.plane_pitch % 4)) return 0;
    if (s.pitch / 4 > 0x3fffffff || d.pitch / 4 > 0x3fffffff || s.h * (s.pitch / 4) >= 0x7fffffffLL) return 0;

    JacFastParams P;
    memset(&P, 0, sizeof(P));
    P.src = static_cast<const float *>(s.data); P.dst = static_cast<float *>(d.data);
    P.src_plane = s.plane_pitch / 4; P.dst_plane = d.plane_pitch / 4;
    P.src_pe = (int)(s.pitch / 4); P.dst_pe = (int)(d.pitch / 4);
    P.src_h = (int)s.h; P.src_w = (int)s.w; P.src_x0 = (int)s.x0; P.src_y0 = (int)s.y0;
    P.full_h = (int)s.full_h; P.full_w = (int)s.full_w;
    P.dst_h = (int)d.h; P.dst_w = (int)d.w; P.dst_x0 = (int)d.x0; P.dst_y0 = (int)d.y0;
    P.grid_h = (int)d.full_h; P.grid_w = (int)d.full_w;
    P.bound_x = a->bound_x; P.bound_y = a->bound_y;
    P.flux = (a->flags & TFM_F_FLUX) ? 1 : 0;
    P.status = a->status_dev;
    P.max_half = (a->max_reg > 0 && a->max_reg / 2 < 0x3fffffff) ? (int)(a->max_reg / 2) : 0x3fffffff;
    const double padval = 3.0;                                  // Gaussian (helpers.pyx:1612-1615)
    P.qscale = (float)(1.4426950408889634 / (a->oblur * a->oblur));
    P.ob = (float)a->oblur;
    P.padval = (float)padval; P.half_pow = (float)(0.5 * a->pad_pow); P.m2_over_pow = (float)(-2.0 / a->pad_pow);
    P.oblur_d = a->oblur; P.padval_d = padval; P.pad_pow_d = a->pad_pow;

    dim3 grid((unsigned)((P.dst_w + FT - 1) / FT), (unsigned)((P.dst_h + FT - 1) / FT), (unsigned)s.n_planes);
    if (grid.y > 65535u) return 0;
    // tap mode: texture gather where a pitch-linear texture can be laid over the source (one plane,
    // geometry within the 2-D texture limits), else 8-byte loads where rows are 8-byte aligned, else
    // 4-byte loads.  Tuning bits 4096 / 32768 force the 4-byte / 8-byte form for A/B runs.
    int tm = 0, tex_slot = -1;
    if (!(a->tuning & 4096) && (reinterpret_cast<uintptr_t>(s.data) % 8) == 0 && (s.pitch % 8) == 0 &&
        (s.plane_pitch % 8) == 0)
        tm = 1;
    if (!(a->tuning & (4096 | 32768)) && s.n_planes == 1 && s.w <= 131072 && s.h <= 65000 && (s.pitch % 32) == 0 &&
        (reinterpret_cast<uintptr_t>(s.data) % 512) == 0 && get_texture(s.data, (int)s.w, (int)s.h, s.pitch, &P.tex, &tex_slot))
        tm = 2;
    P.tex_edges = (tm == 2 && a->bound_x == TFM_B_TRUNCATE && a->bound_y == TFM_B_TRUNCATE) ? 1 : 0;
    if (P.tex_edges && s.x0 == 0 && s.y0 == 0 && s.full_h == s.h && s.full_w == s.w) P.tex_edges = 2;   // whole image
    const bool want_quads = (tm == 2) != ((a->tuning & 65536) != 0);   // default for texture taps; bit 65536 toggles

    bool lin = true;
    for (int i = 0; i < a->n_ops; ++i) lin = lin && is_linear_like(a->chain[i]);
    if (lin) {
        // constant-Jacobian kernel: a constant |J|^2 cannot be flagged by jump_detect when the threshold
        // exceeds 1 (ratio of neighbouring cells is 1 up to rounding)
        if (!(a->jump_thresh == 0.0 || a->jump_thresh >= 1.001)) return 0;
        compose_linear_run(a->chain, 0, a->n_ops, P.aff, P.aff + 4);
        const double ja = P.aff[0], jb = P.aff[1], jc = P.aff[2], jd = P.aff[3];
        // Q = f(J J^T), f(lambda) = (padval + lambda^(p/2))^(-2/p): the fp64 twin of gaussian_quadratic
        const double g00 = ja * ja + jb * jb, g01 = ja * jc + jb * jd, g11 = jc * jc + jd * jd;
        const double mean = 0.5 * (g00 + g11), dif = 0.5 * (g00 - g11);
        const double rad = sqrt(dif * dif + g01 * g01);
        const double lp = mean + rad, lm = fmax(mean - rad, 0.0);
        auto f = [&](double lam) { return pow(padval + pow(lam, 0.5 * a->pad_pow), -2.0 / a->pad_pow); };
        const double fp = f(lp), fm = f(lm);
        const double beta = (lp - lm > 0.0) ? (fp - fm) / (lp - lm) : 0.0;
        const double q00 = fp - beta * (rad - dif), q01 = beta * g01, q11 = fp - beta * (rad + dif);
        double reg = 2.0 * ceil(a->oblur / sqrt(fp));            // helpers.pyx:1666
        if (a->max_reg > 0 && reg > (double)a->max_reg) reg = (double)a->max_reg;
        if (!(isfinite(q00) && isfinite(q01) && isfinite(q11) && isfinite(reg) && reg >= 0.0 && reg < 2.0e9 &&
              isfinite(P.aff[4]) && isfinite(P.aff[5])))
            return 0;
        const double sc = 1.4426950408889634 / (a->oblur * a->oblur);
        P.half = (int)(reg / 2.0);
        P.Aq = (float)(q00 * sc); P.Bq = (float)(q01 * sc); P.Cq = (float)(q11 * sc);
        P.cA = (float)exp2(-2.0 * q00 * sc); P.cB = (float)exp2(-2.0 * q01 * sc); P.cC = (float)exp2(-2.0 * q11 * sc);
        P.det = (float)fabs(ja * jd - jb * jc);
        P.lw = 4;
        if (a->tuning & 1024) P.lw = 5;
        if (a->tuning & 2048) P.lw = 3;
#define TFM_CJ(H)                                                              \
        if (tm == 2) k_jac_gauss_const<H, 2><<<grid, 256, 0, st>>>(P);          \
        else if (tm == 1) k_jac_gauss_const<H, 1><<<grid, 256, 0, st>>>(P);     \
        else k_jac_gauss_const<H, 0><<<grid, 256, 0, st>>>(P)
        switch (P.half) {
        case 1: TFM_CJ(1); break;
        case 2: TFM_CJ(2); break;
        case 3: TFM_CJ(3); break;
        case 4: TFM_CJ(4); break;
        case 5: TFM_CJ(5); break;
        case 6: TFM_CJ(6); break;
        default: TFM_CJ(0); break;                               // every pixel takes the slow path
        }
#undef TFM_CJ
        count_launch();
        if (tm == 2) texture_used(tex_slot, st);
        set_last_kernel(tm == 2 ? "k_jacobian<fast32,const-J,tex>"
                                : (tm == 1 ? "k_jacobian<fast32,const-J,ld64>" : "k_jacobian<fast32,const-J,ld32>"));
        *rc = cudaGetLastError() == cudaSuccess ? TFM_OK : TFM_ECUDA;
        return 1;
    }

    tfm_op fused[TFM_MAX_OPS];
    const int nf = fuse_chain_fast(a->chain, a->n_ops, fused);
    if (nf != 1 || fused[0].kind != TFM_OP_POLAR) return 0;
    const tfm_op &op = fused[0];
    {
        // The reference's Jacobian is a +-1-pixel central difference; against the derivative it is off by
        // O(step^2) of the angle (and log-radius) step: closed form only when those are < 1e-2
        const double coeff = op.p[6];
        const double t2 = (op.p[0] * op.p[0] + op.p[1] * op.p[1]) * coeff * coeff;
        const double g2 = op.p[2] * op.p[2] + op.p[3] * op.p[3];
        const bool lg = (op.flags & TFM_RAD_R0_NOT_NONE) != 0;
        if (!(t2 < 1.0e-4) || (lg && !(g2 < 1.0e-4))) return 0;
        if (a->jump_thresh != 0.0) {
            const double ratio = polar_jump_ratio_bound(op, d);
            if (!(a->jump_thresh > 0.0) || !(ratio * 1.01 <= a->jump_thresh)) return 0;   // NaN compares false
        }
    }
    for (int k = 0; k < 6; ++k) P.pa[k] = op.p[k];
    P.coeff = op.p[6]; P.r0 = op.p[7]; P.ox = op.p[8]; P.oy = op.p[9];
    P.a2[0] = op.p[10]; P.a2[1] = op.p[11]; P.a2[2] = op.p[12]; P.a2[3] = op.p[13]; P.a2[4] = op.p[14]; P.a2[5] = op.p[15];
    P.ccw = (op.flags & TFM_RAD_CCW) ? 1 : 0;
    P.lg = (op.flags & TFM_RAD_R0_NOT_NONE) ? 1 : 0;
    P.has_a2 = (op.flags & TFM_POLAR_HAS_A2) ? 1 : 0;
    P.tXf = (float)(op.p[0] * op.p[6]); P.tYf = (float)(op.p[1] * op.p[6]);
    P.p2f = (float)op.p[2]; P.p3f = (float)op.p[3];
    for (int k = 0; k < 4; ++k) P.a2f[k] = (float)op.p[10 + k];
    P.lw = 3;
    if (a->tuning & 1024) P.lw = 5;
    if (a->tuning & 2048) P.lw = 2;
    if (a->tuning & 16384) P.lw = 4;
    P.quads = (want_quads && P.lw >= 1 && P.lw <= 4) ? 1 : 0;
    if (P.quads && P.lw <= 3) {
        // source displacement per output step: along x the arc (radius x angular step), along y the radial
        // step; a column of 4 pixels is the compact quad when the arc step dominates over the region
        const double t1 = sqrt(op.p[0] * op.p[0] * op.p[6] * op.p[6]), g1 = fabs(op.p[3]);
        const double t2 = sqrt(op.p[1] * op.p[1] * op.p[6] * op.p[6]), g0 = fabs(op.p[2]);
        const double xm = (double)d.x0 + 0.5 * (double)d.w, ym = (double)d.y0 + 0.5 * (double)d.h;
        const double rmid = fabs(op.p[2] * xm + op.p[3] * ym + op.p[5]);
        const double rr = (op.flags & TFM_RAD_R0_NOT_NONE) ? 1.0 : rmid;       // log-polar: steps scale together
        const double step_x = hypot(rr * t1, g0 * ((op.flags & TFM_RAD_R0_NOT_NONE) ? 1.0 : 1.0));
        const double step_y = hypot(rr * t2, g1);
        if (step_x > 2.0 * step_y) P.quads = 2;
    }
    if (a->tuning & 131072) P.quads = (P.lw <= 3) ? 2 : P.quads;
    // radius constant along output rows (no radius term in X, no trailing affine map): per-row spectral tables
    const bool rows = op.p[2] == 0.0 && !P.has_a2 && !(a->tuning & 262144);
    // two output rows per set of fetches: worth it where one output row maps at most about one source pixel from
    // the next (bound from the closed form over the whole output GRID, not the block: row blocks of one job must
    // all take the same kernel)
    bool pairs = false;
    if (rows && tm == 2 && !(a->tuning & 524288)) {
        const bool lgp = P.lg != 0;
        double vmax = 0.0;
        for (int k = 0; k < 2; ++k) {
            const double v = op.p[3] * (k ? (double)d.full_h : 0.0) + op.p[5];
            vmax = fmax(vmax, lgp ? v : fabs(v));
        }
        const double rmax = lgp ? op.p[7] * exp(vmax) : vmax;
        const double dsy = lgp ? rmax * hypot(op.p[1] * op.p[6], op.p[3]) : hypot(rmax * op.p[1] * op.p[6], op.p[3]);
        pairs = dsy <= 1.05;
    }
    if (pairs) {
        dim3 pgrid(grid.x, (unsigned)((P.dst_h + (P.dst_y0 & 1) + FT - 1) / FT), grid.z);
        if (P.quads) {
            // lanes 4k .. 4k+3: 4 pairs of one column (8 output rows) or of one row (4 output columns), whichever
            // spans less of the source at the middle of the grid
            const double rmid = P.lg ? 1.0 : fabs(op.p[3] * 0.5 * (double)d.full_h + op.p[5]);
            const double ax = hypot(rmid * op.p[0] * op.p[6], op.p[2]), ay = hypot(rmid * op.p[1] * op.p[6], op.p[3]);
            P.quads = (3.0 * ax > 7.0 * ay) ? 2 : 1;
            if (a->tuning & 131072) P.quads = 3 - P.quads;
        }
        k_jac_gauss_polar_rows_pair<<<pgrid, 256, 0, st>>>(P);
    } else if (rows) {
        if (tm == 2) k_jac_gauss_polar_rows<2><<<grid, 256, 0, st>>>(P);
        else if (tm == 1) k_jac_gauss_polar_rows<1><<<grid, 256, 0, st>>>(P);
        else k_jac_gauss_polar_rows<0><<<grid, 256, 0, st>>>(P);
    } else {
        if (tm == 2) k_jac_gauss_polar<2><<<grid, 256, 0, st>>>(P);
        else if (tm == 1) k_jac_gauss_polar<1><<<grid, 256, 0, st>>>(P);
        else k_jac_gauss_polar<0><<<grid, 256, 0, st>>>(P);
    }
    count_launch();
    if (tm == 2) texture_used(tex_slot, st);
    static const char *const names[2][3] = {
        {"k_jacobian<fast32,analytic-polar,ld32>", "k_jacobian<fast32,analytic-polar,ld64>", "k_jacobian<fast32,analytic-polar,tex>"},
        {"k_jacobian<fast32,analytic-polar-rows,ld32>", "k_jacobian<fast32,analytic-polar-rows,ld64>",
         "k_jacobian<fast32,analytic-polar-rows,tex>"}};
    set_last_kernel(pairs ? "k_jacobian<fast32,analytic-polar-rows,tex,row-pairs>" : names[rows ? 1 : 0][tm]);
    *rc = cudaGetLastError() == cudaSuccess ? TFM_OK : TFM_ECUDA;
    return 1;
}

}  // namespace tfm
