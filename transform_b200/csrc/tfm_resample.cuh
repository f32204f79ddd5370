// tfm_resample.cuh -- K1 kernels and their launch templates (see tfm_resample.cu for the overview).
// Included by the dispatch unit (extern templates) and by the instantiation units
// tfm_resample_exact_f32.cu / tfm_resample_exact_f64.cu / tfm_resample_fast.cu, which exist only to
// compile the kernel families in parallel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <mutex>
#include "tfm_chain.cuh"
#include "tfm_common.cuh"

namespace tfm {

struct ResampleParams {
    const void *src;
    void *dst;
    long long src_pitch, dst_pitch;          // bytes
    long long src_plane, dst_plane;          // bytes between planes (blockIdx.z)
    int src_h, src_w;                        // staged rows / columns at `src`
    int src_x0, src_y0;                      // global origin of src[0][0]
    int full_h, full_w;                      // full source extent (boundary conditions)
    int dst_h, dst_w;
    long long dst_x0, dst_y0;                // global grid origin of dst[0][0]
    int bound_x, bound_y;
    int flags;
    int lw;                                  // log2 of warp patch width (5, 4 or 3)
    int n_planes;
    int int_w, int_h;                        // #columns / #rows at which a whole NT x NT footprint fits
    double aff[6];                           // pre-composed affine map a00 a01 a10 a11 b0 b1 (AFF kernels)
    int aff_bounded;                         // host proved |coordinate| < 2^30 over the whole output grid
    float aff_f[8];                          // fp32 copy of aff[0..5] + (dst_x0, dst_y0) for the empty-tile test
    float aff_margin;                        // absolute fp32 rounding allowance for that test
    float aff_c[4];                          // tile-centre form of the same test (32 x 32 tiles): source coordinate of
                                             // the centre of tile (0,0) relative to the output origin, and the
                                             // half-extents (footprint, safety margin and rounding allowance included)
    float aff_lim[2];                        // src_w + aff_c[2], src_h + aff_c[3]
    double fill;
    int *status;
    int win_method;                          // TFM_WIN2 kernels: which window (TFM_HANN / TFM_TUKEY / TFM_TENT ...)
    double win_oblur;
    ChainDev chain;
};

// ---------------------------------------------------------------------------------------------
// tap index helpers
// ---------------------------------------------------------------------------------------------
// floor value (integer-valued double, or NaN/inf) -> int64 tap base, x86-64 `.astype(int)` rule:
// NaN / inf / |f| >= 2^63 give INT64_MIN (helpers.pyx:127; oracle_np._cast_i64).
__device__ __forceinline__ long long cast_i64(double f)
{
    return (fabs(f) < 9.2e18) ? __double2ll_rz(f) : LLONG_MIN;
}

// Generic boundary on one axis (helpers.pyx:136-176).  Returns the in-range index, or -1 with
// `trunc` set for a truncated tap; 'forbid' violations raise bit 1 in *status and read as truncated.
static __device__ __noinline__ int bound_slow(long long i, int size, int mode, bool &trunc, int *status)
{
    switch (mode) {
    case TFM_B_EXTEND:
        return i < 0 ? 0 : size - 1;
    case TFM_B_PERIODIC: {
        long long r = i % (long long)size;
        if (r < 0) r += size;
        return (int)r;
    }
    case TFM_B_MIRROR: {
        long long s2 = 2LL * size;
        long long r = i % s2;
        if (r < 0) r += s2;
        if (r >= size) r = s2 - 1 - r;
        return (int)r;
    }
    case TFM_B_FORBID:
        if (status) atomicOr(status, 1);
        trunc = true;
        return -1;
    default:  // truncate
        trunc = true;
        return -1;
    }
}

__device__ __forceinline__ int bound_index(long long i, int size, int mode, bool &trunc, int *status)
{
    if ((unsigned long long)i < (unsigned long long)size) return (int)i;
    return bound_slow(i, size, mode, trunc, status);
}

template <typename T>
__device__ __forceinline__ T ldg_px(const void *base, long long pitch, int y, int x)
{
    return __ldg(reinterpret_cast<const T *>(static_cast<const char *>(base) + (long long)y * pitch) + x);
}

// One tap, generic path: global index -> boundary -> staged sub-rectangle -> value.
template <typename SrcT, typename RegT>
__device__ __forceinline__ RegT tap_generic(const ResampleParams &P, const void *src, long long gx, long long gy)
{
    bool trunc = false;
    int jx = bound_index(gx, P.full_w, P.bound_x, trunc, P.status);
    int jy = bound_index(gy, P.full_h, P.bound_y, trunc, P.status);
    if (trunc) return (RegT)P.fill;
    jx -= P.src_x0;
    jy -= P.src_y0;
    if ((unsigned)jx >= (unsigned)P.src_w || (unsigned)jy >= (unsigned)P.src_h) {
        if (P.status) atomicOr(P.status, 2);             // host staged too small a sub-rectangle
        return (RegT)P.fill;
    }
    return (RegT)ldg_px<SrcT>(src, P.src_pitch, jy, jx);
}

// One tap, truncate-on-both-axes whole-image path (the reference default bound='t').
template <typename SrcT, typename RegT>
__device__ __forceinline__ RegT tap_trunc(const ResampleParams &P, const void *src, int ix, int iy)
{
    bool inb = ((unsigned)ix < (unsigned)P.src_w) & ((unsigned)iy < (unsigned)P.src_h);
    RegT v = (RegT)P.fill;
    if (inb) v = (RegT)ldg_px<SrcT>(src, P.src_pitch, iy, ix);
    return v;
}

// Tap base for the truncate fast path: any coordinate that is NaN or absurdly large maps to an
// index whose whole footprint is out of range (the reference's INT64_MIN is out of range too).
__device__ __forceinline__ int base_trunc(double f)
{
    return (fabs(f) < TFM_SANE_COORD) ? __double2int_rz(f) : -1000000;
}

// ---------------------------------------------------------------------------------------------
// interpolation arithmetic
// ---------------------------------------------------------------------------------------------
// Cubic Hermite with central-difference slopes in the reference's operation order
// (helpers.pyx:755-778).  T = float reproduces NumPy's float32 first pass (region not promoted).
__device__ __forceinline__ double hermite_exact(double r0, double r1, double r2, double r3, double b)
{
    double d = __dsub_rn(r2, r1);
    double s0 = __dmul_rn(__dsub_rn(r2, r0), 0.5);
    double s1 = __dmul_rn(__dsub_rn(r3, r1), 0.5);
    double c2 = __dsub_rn(__dsub_rn(__dmul_rn(3.0, d), __dmul_rn(2.0, s0)), s1);
    double c3 = __dsub_rn(__dadd_rn(s1, s0), __dmul_rn(2.0, d));
    double t = __dadd_rn(c2, __dmul_rn(b, c3));
    t = __dadd_rn(s0, __dmul_rn(b, t));
    return __dadd_rn(r1, __dmul_rn(b, t));
}

__device__ __forceinline__ double hermite_exact_f32(float r0, float r1, float r2, float r3, double b)
{
    float d = __fsub_rn(r2, r1);
    float s0 = __fmul_rn(__fsub_rn(r2, r0), 0.5f);
    float s1 = __fmul_rn(__fsub_rn(r3, r1), 0.5f);
    float c2 = __fsub_rn(__fsub_rn(__fmul_rn(3.0f, d), __fmul_rn(2.0f, s0)), s1);
    float c3 = __fsub_rn(__fadd_rn(s1, s0), __fmul_rn(2.0f, d));
    double t = __dadd_rn((double)c2, __dmul_rn(b, (double)c3));
    t = __dadd_rn((double)s0, __dmul_rn(b, t));
    return __dadd_rn((double)r1, __dmul_rn(b, t));
}

__device__ __forceinline__ float hermite_fast(float r0, float r1, float r2, float r3, float b)
{
    float d = r2 - r1;
    float s0 = (r2 - r0) * 0.5f;
    float s1 = (r3 - r1) * 0.5f;
    float c2 = 3.0f * d - 2.0f * s0 - s1;
    float c3 = s1 + s0 - 2.0f * d;
    return r1 + b * (s0 + b * (c2 + b * c3));
}

// ---------------------------------------------------------------------------------------------
// window functions of the collapsible filters (helpers.pyx:792-835; see k_resample_filter below)
// ---------------------------------------------------------------------------------------------
struct FilterParams {
    int method;                              // TFM_SINC .. TFM_TENT
    int size;                                // region edge S
    int offset;                              // S/2 - 1
    double oblur;                            // 1.0 when None
};

__device__ __forceinline__ double clip_keepnan(double x, double lo, double hi)
{
    return x < lo ? lo : (x > hi ? hi : x);             // NaN falls through, like np.clip
}

__device__ __forceinline__ double np_sinc(double x)
{
    const double y = __dmul_rn(3.141592653589793, x == 0.0 ? 1.0e-20 : x);     // numpy.sinc
    return __ddiv_rn(sin(y), y);
}

// window value at offset `of`, operation order of helpers.pyx:809-826
__device__ __forceinline__ double window_exact(int m, double of)
{
    switch (m) {
    case TFM_SINC: return np_sinc(of);
    case TFM_LANCZOS: {
        const double bb = clip_keepnan(of, -3.0, 3.0);
        return __dmul_rn(__dmul_rn(3.0, np_sinc(bb)), np_sinc(__ddiv_rn(bb, 3.0)));
    }
    case TFM_GAUSSIAN: return exp(__dmul_rn(__dmul_rn(__dmul_rn(-of, of), 2.0), 2.0));   // -bb*bb/0.5/0.5
    case TFM_HANN: return __dadd_rn(1.0, cos(__dmul_rn(3.141592653589793, clip_keepnan(of, -1.0, 1.0))));
    case TFM_TUKEY:
        return __dadd_rn(1.0, cos(__dmul_rn(3.141592653589793,
                                            clip_keepnan(__dsub_rn(__dmul_rn(fabs(of), 2.0), 0.5), 0.0, 1.0))));
    default: return __dsub_rn(1.0, clip_keepnan(fabs(of), 0.0, 1.0));
    }
}

__device__ __forceinline__ float window_fast(int m, float of)
{
    // sinpif / cospif reduce the argument exactly (the sinc and Lanczos weights alternate in sign, so
    // absolute errors of the plain fast intrinsics would not cancel); 2S evaluations per pixel only
    const float PI = 3.14159265358979f;
    switch (m) {
    case TFM_SINC: {
        const float y = PI * of;
        return fabsf(of) < 1e-4f ? 1.0f : __fdividef(sinpif(of), y);
    }
    case TFM_LANCZOS: {
        const float bb = fminf(fmaxf(of, -3.0f), 3.0f);
        const float b3 = bb * (1.0f / 3.0f);
        return fabsf(bb) < 1e-3f ? 3.0f : __fdividef(3.0f * sinpif(bb) * sinpif(b3), (PI * PI) * bb * b3);
    }
    case TFM_GAUSSIAN: return exp2f(-5.770780163555854f * of * of);              // exp(-4 of^2)
    case TFM_HANN: return 1.0f + cospif(fminf(fmaxf(of, -1.0f), 1.0f));
    case TFM_TUKEY: return 1.0f + cospif(fminf(fmaxf(fabsf(of) * 2.0f - 0.5f, 0.0f), 1.0f));
    default: return 1.0f - fminf(fabsf(of), 1.0f);
    }
}

// Fast-mode window values at the consecutive offsets of_i = o + i (oblur = 1): the sinc and Lanczos
// windows need sin(pi of_i) = (-1)^i sin(pi o) and sin(pi of_i / 3), which advances by a fixed
// rotation of pi/3 -- one sinpif / sincospif per axis instead of one or two per tap position (window
// evaluation was 870 of the 1,516 instructions per pixel of the Lanczos kernel).  Other windows, and
// any oblur != 1, evaluate window_fast per position.
struct WinSeq {
    int method;
    float o, iob;
    bool recur;
    float sg, s3, c3;

    __device__ __forceinline__ void init(int m, float o_, float iob_)
    {
        method = m; o = o_; iob = iob_;
        recur = (iob_ == 1.0f) && (m == TFM_SINC || m == TFM_LANCZOS);
        sg = 0.f; s3 = 0.f; c3 = 1.f;
        if (recur) {
            sg = sinpif(o_);
            if (m == TFM_LANCZOS) sincospif(o_ * (1.0f / 3.0f), &s3, &c3);
        }
    }
    // value at position i; must be called for i = 0, 1, 2, ... in order
    __device__ __forceinline__ float next(int i)
    {
        if (!recur) return window_fast(method, (o + (float)i) * iob);
        const float PI = 3.14159265358979f;
        const float x = o + (float)i;
        float k;
        if (method == TFM_SINC) {
            k = fabsf(x) < 1e-4f ? 1.0f : __fdividef(sg, PI * x);
        } else {
            k = fabsf(x) < 1e-3f ? 3.0f : (fabsf(x) >= 3.0f ? 0.0f : __fdividef(9.0f * sg * s3, (PI * PI) * x * x));
            const float ns = fmaf(c3, 0.8660254037844386f, 0.5f * s3), nc = fmaf(-s3, 0.8660254037844386f, 0.5f * c3);
            s3 = ns; c3 = nc;
        }
        sg = -sg;
        return k;
    }
};

// of_i = ((1 - b) - (offset + 1) + i) / oblur      (helpers.pyx:800-802)
__device__ __forceinline__ double filt_of(double one_minus_b, const FilterParams &F, int i)
{
    return __ddiv_rn(__dadd_rn(__dsub_rn(one_minus_b, (double)(F.offset + 1)), (double)i), F.oblur);
}

#define TFM_WIN2 9   // internal METHOD: any filter whose region is 2 x 2 (Hann, Tukey, tent with oblur <= 1)

template <int METHOD> struct MethodTraits;
template <> struct MethodTraits<TFM_NEAREST> { static constexpr int NPX = 4, NT = 1; };
template <> struct MethodTraits<TFM_LINEAR>  { static constexpr int NPX = 4, NT = 2; };
template <> struct MethodTraits<TFM_CUBIC>   { static constexpr int NPX = 2, NT = 4; };
template <> struct MethodTraits<TFM_WIN2>    { static constexpr int NPX = 4, NT = 2; };

// phase 3: weights, reduction, store.  CHECK = false: every pixel of the warp is a valid output
// pixel (decided by the warp vote), so the stores are unconditional.
template <int METHOD, typename SrcT, typename DstT, typename RegT, int NPX, int NT, bool CHECK>
__device__ __forceinline__ void finish(const ResampleParams &P, RegT (&v)[NPX][NT][NT],
                                       const double (&fracx)[NPX], const double (&fracy)[NPX],
                                       const float (&ffx)[NPX], const float (&ffy)[NPX],
                                       const bool (&valid)[NPX], char *dst_base, int X, int Y0)
{
    constexpr bool FASTI = (sizeof(RegT) == 4);
#pragma unroll
    for (int j = 0; j < NPX; ++j) {
        if (CHECK && !valid[j]) continue;
        const int Y = Y0 + 8 * j;
        RegT r;
        if (METHOD == TFM_NEAREST) {
            r = v[j][0][0];
        } else if (METHOD == TFM_LINEAR) {
            if (!FASTI) {
                // helpers.pyx:584-610: weight = prod(alpha | 1-alpha); sum(region*weight) runs
                // sequentially over the contiguous 2x2 block (x fastest).
                const double ax = fracx[j], ay = fracy[j];
                const double bx = __dsub_rn(1.0, ax), by = __dsub_rn(1.0, ay);
                double acc = __dmul_rn((double)v[j][0][0], __dmul_rn(bx, by));
                acc = __dadd_rn(acc, __dmul_rn((double)v[j][0][1], __dmul_rn(ax, by)));
                acc = __dadd_rn(acc, __dmul_rn((double)v[j][1][0], __dmul_rn(bx, ay)));
                acc = __dadd_rn(acc, __dmul_rn((double)v[j][1][1], __dmul_rn(ax, ay)));
                r = (RegT)acc;
            } else {
                const float ax = ffx[j], ay = ffy[j];
                const float top = fmaf(ax, (float)v[j][0][1] - (float)v[j][0][0], (float)v[j][0][0]);
                const float bot = fmaf(ax, (float)v[j][1][1] - (float)v[j][1][0], (float)v[j][1][0]);
                r = (RegT)fmaf(ay, bot - top, top);
            }
        } else if (METHOD == TFM_WIN2) {
            // 2 x 2 region of a collapsible filter (helpers.pyx:792-835), offset = 0:
            //   of_i = ((1 - b) - 1 + i) / oblur;  row_j = (kx0 v_j0 + kx1 v_j1) / (kx0 + kx1);
            //   out = (ky0 row_0 + ky1 row_1) / (ky0 + ky1)
            if (!FASTI) {
                FilterParams F;
                F.method = P.win_method; F.size = 2; F.offset = 0; F.oblur = P.win_oblur;
                const double bx1 = __dsub_rn(1.0, fracx[j]), by1 = __dsub_rn(1.0, fracy[j]);
                const double kx0 = window_exact(F.method, filt_of(bx1, F, 0)), kx1 = window_exact(F.method, filt_of(bx1, F, 1));
                const double ky0 = window_exact(F.method, filt_of(by1, F, 0)), ky1 = window_exact(F.method, filt_of(by1, F, 1));
                const double kxs = __dadd_rn(kx0, kx1), kys = __dadd_rn(ky0, ky1);
                const double r0 = __ddiv_rn(__dadd_rn(__dmul_rn(kx0, (double)v[j][0][0]), __dmul_rn(kx1, (double)v[j][0][1])), kxs);
                const double r1 = __ddiv_rn(__dadd_rn(__dmul_rn(kx0, (double)v[j][1][0]), __dmul_rn(kx1, (double)v[j][1][1])), kxs);
                r = (RegT)__ddiv_rn(__dadd_rn(__dmul_rn(ky0, r0), __dmul_rn(ky1, r1)), kys);
            } else {
                const float iob = 1.0f / (float)P.win_oblur;
                const float bx1 = 1.0f - ffx[j], by1 = 1.0f - ffy[j];
                const float kx0 = window_fast(P.win_method, (bx1 - 1.0f) * iob), kx1 = window_fast(P.win_method, bx1 * iob);
                const float ky0 = window_fast(P.win_method, (by1 - 1.0f) * iob), ky1 = window_fast(P.win_method, by1 * iob);
                const float r0 = fmaf(kx1, (float)v[j][0][1], kx0 * (float)v[j][0][0]);
                const float r1 = fmaf(kx1, (float)v[j][1][1], kx0 * (float)v[j][1][0]);
                r = (RegT)__fdividef(fmaf(ky1, r1, ky0 * r0), (kx0 + kx1) * (ky0 + ky1));
            }
        } else {
            if (!FASTI) {
                double rows[4];
                const bool f32pass = (sizeof(SrcT) == 4) && (P.flags & TFM_F_REGION_SRC_DTYPE);
#pragma unroll
                for (int ty = 0; ty < 4; ++ty) {
                    if (f32pass)
                        rows[ty] = hermite_exact_f32((float)v[j][ty][0], (float)v[j][ty][1],
                                                     (float)v[j][ty][2], (float)v[j][ty][3], fracx[j]);
                    else
                        rows[ty] = hermite_exact((double)v[j][ty][0], (double)v[j][ty][1],
                                                 (double)v[j][ty][2], (double)v[j][ty][3], fracx[j]);
                }
                r = (RegT)hermite_exact(rows[0], rows[1], rows[2], rows[3], fracy[j]);
            } else {
                float rows[4];
#pragma unroll
                for (int ty = 0; ty < 4; ++ty)
                    rows[ty] = hermite_fast((float)v[j][ty][0], (float)v[j][ty][1],
                                            (float)v[j][ty][2], (float)v[j][ty][3], ffx[j]);
                r = (RegT)hermite_fast(rows[0], rows[1], rows[2], rows[3], ffy[j]);
            }
        }
        DstT *row = reinterpret_cast<DstT *>(dst_base + (long long)Y * P.dst_pitch);
        row[X] = (DstT)r;
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// METHOD   tfm_method
// SrcT     source element type; DstT output element type
// A        chain arithmetic policy (Exact / Fast)
// RegT     arithmetic type of the interpolation (double in exact mode, float in fast mode)
// GENERIC  false: both axes 'truncate', source is the whole image (fast path);
//          true : per-axis runtime boundary modes and/or staged source sub-rectangle
// AFF      the whole chain was pre-composed by the host into one affine map P.aff (fast mode only)
template <int METHOD, typename SrcT, typename DstT, class A, typename RegT, bool GENERIC, bool AFF, int SPH>
__global__ void __launch_bounds__(256)
k_resample(const __grid_constant__ ResampleParams P)
{
    constexpr int NPX = MethodTraits<METHOD>::NPX;
    constexpr int NT = MethodTraits<METHOD>::NT;
    constexpr int T0 = (METHOD == TFM_CUBIC) ? -1 : 0;          // first tap offset
    constexpr bool FASTI = (sizeof(RegT) == 4);                 // fp32 interpolation arithmetic

    // a warp covers a 16 x 2 patch of the tile (compile-time shape: measured within 8% of the best
    // of 32x1 / 16x2 / 8x4 on rotated maps, and no worse on axis-aligned ones)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const char *const src_base = static_cast<const char *>(P.src) + (long long)blockIdx.z * P.src_plane;
    char *const dst_base = static_cast<char *>(P.dst) + (long long)blockIdx.z * P.dst_plane;
    const int px = ((warp & 1) << 4) + (lane & 15);
    const int py = ((warp >> 1) << 1) + (lane >> 4);
    const int X = blockIdx.x * 32 + px;
    const int Y0 = blockIdx.y * (8 * NPX) + py;
    const bool xin = X < P.dst_w;

    // ---- phase 0 (host-composed affine maps only): CTA-uniform rejection of empty tiles ----
    // The tile maps to a parallelogram whose bounding box is spanned by its 4 corners.  If that box,
    // padded by the footprint and a 2-pixel safety margin, misses a truncated image entirely, every
    // pixel of the tile is the fill value: store it and retire the CTA before any per-pixel work.
    if (AFF && !GENERIC && P.aff_bounded) {
        // fp32 is plenty here: |coordinate| < 2.5e8 is host-proven and the test carries a margin of
        // 2 pixels + 1e-6 relative (P.aff_margin) on top of the footprint
        const float tx0 = (float)(int)(blockIdx.x * 32) + P.aff_f[6], tx1 = tx0 + 31.0f;
        const float ty0 = (float)(int)(blockIdx.y * (8 * NPX)) + P.aff_f[7], ty1 = ty0 + (float)(8 * NPX - 1);
        const float ux0 = P.aff_f[0] * tx0, ux1 = P.aff_f[0] * tx1, uy0 = P.aff_f[1] * ty0, uy1 = P.aff_f[1] * ty1;
        const float vx0 = P.aff_f[2] * tx0, vx1 = P.aff_f[2] * tx1, vy0 = P.aff_f[3] * ty0, vy1 = P.aff_f[3] * ty1;
        const float xmin = fminf(ux0, ux1) + fminf(uy0, uy1) + P.aff_f[4], xmax = fmaxf(ux0, ux1) + fmaxf(uy0, uy1) + P.aff_f[4];
        const float ymin = fminf(vx0, vx1) + fminf(vy0, vy1) + P.aff_f[5], ymax = fmaxf(vx0, vx1) + fmaxf(vy0, vy1) + P.aff_f[5];
        const float pad = (float)(NT + 2) + P.aff_margin;
        if (xmax < -pad || ymax < -pad || xmin > (float)P.src_w + pad || ymin > (float)P.src_h + pad) {
            const DstT f = (DstT)(RegT)P.fill;
            if (xin) {
#pragma unroll
                for (int j = 0; j < NPX; ++j) {
                    const int Y = Y0 + 8 * j;
                    if (Y < P.dst_h)
                        reinterpret_cast<DstT *>(dst_base + (long long)Y * P.dst_pitch)[X] = f;
                }
            }
            return;
        }
    }

    // ---- phase 1: coordinates (fp64, in registers) ----
    double cx[NPX], cy[NPX];
    if (AFF) {
        // canonical form of every fast-mode kernel: c = fma(a0, X, fma(a1, Y, b)) on exact doubles X, Y
        const double Xd = (double)(P.dst_x0 + X), Yd = (double)(P.dst_y0 + Y0);
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
            const double Yj = Yd + (double)(8 * j);
            cx[j] = fma(P.aff[0], Xd, fma(P.aff[1], Yj, P.aff[4]));
            cy[j] = fma(P.aff[2], Xd, fma(P.aff[3], Yj, P.aff[5]));
        }
    } else if (SPH == 1 && !A::exact && P.chain.n > 0 && P.chain.ops[0].kind == TFM_OP_WCS_PAIR &&
               (P.chain.ops[0].flags & 0xffff) == (TFM_WCS_PROJ_CAR | (TFM_WCS_PROJ_TAN << 8))) {
        // fused WCS pair, CAR pixel -> sky -> TAN pixel (a TAN -> CAR remap seen from its output grid):
        // longitude / latitude are affine in the pixel, their sincos come from per-tile tables
        __shared__ double4 s_t0[32], s_t1[8 * NPX];         // separable form: column / row tables;
        const tfm_op &op0 = P.chain.ops[0];                   // general form: (aphi, ath) / (bphi, bth) pairs
        const double TX0 = (double)(P.dst_x0 + (long long)blockIdx.x * 32);
        const double TY0 = (double)(P.dst_y0 + (long long)blockIdx.y * (8 * NPX));
        const WcsCarCoef kc = wcs_car_coef(op0);
        if (op0.p[3] == 0.0 && op0.p[4] == 0.0) {              // axis-aligned CD (CTA-uniform)
            wcs_car_sep_tables(op0, kc, TX0, TY0, 32, 8 * NPX, s_t0, s_t1, threadIdx.x, 256);
            __syncthreads();
            const double4 a = s_t0[px];
#pragma unroll
            for (int j = 0; j < NPX; ++j) wcs_car_tan_sep(kc, a, s_t1[py + 8 * j], cx[j], cy[j]);
        } else {
            double2 *const s_aphi = reinterpret_cast<double2 *>(s_t0), *const s_ath = s_aphi + 32;
            double2 *const s_bphi = reinterpret_cast<double2 *>(s_t1), *const s_bth = s_bphi + 8 * NPX;
            wcs_car_tables(op0, TX0, TY0, 32, 8 * NPX, s_aphi, s_bphi, s_ath, s_bth, threadIdx.x, 256);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < NPX; ++j)
                wcs_car_tan_from_tables(kc, px, py + 8 * j, s_aphi, s_bphi, s_ath, s_bth, cx[j], cy[j]);
        }
        chain_eval_n<A, SPH, NPX>(P.chain, cx, cy, 1);
    } else {
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
            cx[j] = (double)(P.dst_x0 + X);
            cy[j] = (double)(P.dst_y0 + Y0 + 8 * j);
        }
        chain_eval_n<A, SPH, NPX>(P.chain, cx, cy);
    }

    // ---- phase 2: taps (all loads of the NPX pixels issued before any is consumed) ----
    RegT v[NPX][NT][NT];
    double fracx[NPX], fracy[NPX];          // exact mode
    float ffx[NPX], ffy[NPX];               // fast mode
    bool valid[NPX];

    if (!GENERIC) {
        int ix[NPX], iy[NPX];
        bool sane[NPX];
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
            double x = cx[j], y = cy[j];
            if (METHOD == TFM_NEAREST) { x = __dadd_rn(x, 0.5); y = __dadd_rn(y, 0.5); }   // helpers.pyx:127
            if (FASTI) {
                floor_frac_fast(x, ix[j], ffx[j]);
                floor_frac_fast(y, iy[j], ffy[j]);
            } else {
                const double fx = floor(x), fy = floor(y);
                fracx[j] = __dsub_rn(x, fx);                    // alpha = index - floor(index)
                fracy[j] = __dsub_rn(y, fy);
                ix[j] = __double2int_rz(fx);
                iy[j] = __double2int_rz(fy);
            }
            // false for NaN / inf / absurd.  For host-composed affine maps the host has bounded
            // |coordinate| over the whole output grid (P.aff_bounded): nothing to test per pixel.
            if (AFF) sane[j] = true;
            else {
                sane[j] = (fabs(x) < TFM_SANE_COORD) && (fabs(y) < TFM_SANE_COORD);
                // the fixed-point fraction is meaningless out there: NaN for a non-finite coordinate (the
                // reference's alpha = v - floor(v)), 0 for an absurd finite one
                if (FASTI && !sane[j]) { ffx[j] = (float)(x - x); ffy[j] = (float)(y - y); }
            }
        }
        const int pe = (int)(P.src_pitch / (long long)sizeof(SrcT));       // host guarantees h*pe < 2^31
        // CTA-uniform: the whole 32 x (8*NPX) tile lies inside the output and (AFF) coordinates are
        // known to be sane -> the per-warp classification below needs range compares only.
        const bool tile_ok = (blockIdx.x * 32 + 32 <= P.dst_w) && (blockIdx.y * (8 * NPX) + 8 * NPX <= P.dst_h) &&
                             (!AFF || P.aff_bounded);
        if (tile_ok) {
            bool interior = true;
#pragma unroll
            for (int j = 0; j < NPX; ++j)
                interior = interior && sane[j] && ((unsigned)(ix[j] + T0) < (unsigned)P.int_w) &&
                           ((unsigned)(iy[j] + T0) < (unsigned)P.int_h);
            if (__all_sync(0xffffffffu, interior)) {
                // every tap of every pixel of this warp lies inside the image: no predicates
                const SrcT *const base = reinterpret_cast<const SrcT *>(src_base);
#pragma unroll
                for (int j = 0; j < NPX; ++j) {
                    const int o = (iy[j] + T0) * pe + (ix[j] + T0);
#pragma unroll
                    for (int ty = 0; ty < NT; ++ty) {
                        const SrcT *p = base + (o + ty * pe);
#pragma unroll
                        for (int tx = 0; tx < NT; ++tx) v[j][ty][tx] = (RegT)__ldg(p + tx);
                    }
                }
#pragma unroll
                for (int j = 0; j < NPX; ++j) valid[j] = true;
                finish<METHOD, SrcT, DstT, RegT, NPX, NT, false>(P, v, fracx, fracy, ffx, ffy, valid, dst_base, X, Y0);
                return;
            }
            // whole footprint of every pixel outside the image: the result is the fill value
            bool outside = true;
#pragma unroll
            for (int j = 0; j < NPX; ++j)
                outside = outside && (!sane[j] ||
                                      (unsigned)(ix[j] + T0 + NT - 1) >= (unsigned)(P.src_w + NT - 1) ||
                                      (unsigned)(iy[j] + T0 + NT - 1) >= (unsigned)(P.src_h + NT - 1));
            if (__all_sync(0xffffffffu, outside)) {
#pragma unroll
                for (int j = 0; j < NPX; ++j) {
                    valid[j] = true;
#pragma unroll
                    for (int ty = 0; ty < NT; ++ty)
#pragma unroll
                        for (int tx = 0; tx < NT; ++tx) v[j][ty][tx] = (RegT)P.fill;
                }
                finish<METHOD, SrcT, DstT, RegT, NPX, NT, false>(P, v, fracx, fracy, ffx, ffy, valid, dst_base, X, Y0);
                return;
            }
        }
        // mixed warp or ragged tile: per-tap predicates
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
            valid[j] = xin && (Y0 + 8 * j < P.dst_h);
            const bool ok = sane[j] && (!AFF || P.aff_bounded || ((fabs(cx[j]) < TFM_SANE_COORD) && (fabs(cy[j]) < TFM_SANE_COORD)));
            // NaN / absurd coordinate: whole footprint out of range -> fill (INT64_MIN in the reference)
            const int bx = ok ? ix[j] : -1000000, by = ok ? iy[j] : -1000000;
#pragma unroll
            for (int ty = 0; ty < NT; ++ty)
#pragma unroll
                for (int tx = 0; tx < NT; ++tx)
                    v[j][ty][tx] = valid[j] ? tap_trunc<SrcT, RegT>(P, src_base, bx + T0 + tx, by + T0 + ty)
                                            : (RegT)0;
        }
    } else {
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
            valid[j] = xin && (Y0 + 8 * j < P.dst_h);
            double x = cx[j], y = cy[j];
            if (METHOD == TFM_NEAREST) { x = __dadd_rn(x, 0.5); y = __dadd_rn(y, 0.5); }
            const double fx = floor(x), fy = floor(y);
            fracx[j] = __dsub_rn(x, fx);
            fracy[j] = __dsub_rn(y, fy);
            ffx[j] = (float)fracx[j];
            ffy[j] = (float)fracy[j];
            // |f| >= 2^51 (or NaN): the reference's float tap offsets are absorbed / invalid; every
            // tap shares the base index ("absurd coordinates").
            const bool okx = fabs(fx) < 2.0e15, oky = fabs(fy) < 2.0e15;
            long long bx = cast_i64(fx), by = cast_i64(fy);
            if (FASTI && fabs(x) < TFM_SANE_COORD && fabs(y) < TFM_SANE_COORD) {
                // every fast-mode variant (LDG gather, texture gather, TMA pipeline, this generic one) takes
                // tap index AND fraction from the one fixed-point floor, so that a sharded run reproduces the
                // unsharded one bit for bit.  (Index from floor() with the fraction from floor_frac_fast was
                // a bug: a coordinate a hair below an integer k floors to k-1 but quantises to (k, 0.0).)
                int iu, iv;
                floor_frac_fast(x, iu, ffx[j]);
                floor_frac_fast(y, iv, ffy[j]);
                bx = iu;
                by = iv;
            }
#pragma unroll
            for (int ty = 0; ty < NT; ++ty)
#pragma unroll
                for (int tx = 0; tx < NT; ++tx)
                    v[j][ty][tx] = valid[j] ? tap_generic<SrcT, RegT>(P, src_base, okx ? bx + T0 + tx : bx,
                                                                      oky ? by + T0 + ty : by)
                                            : (RegT)0;
        }
    }

    finish<METHOD, SrcT, DstT, RegT, NPX, NT, true>(P, v, fracx, fracy, ffx, ffy, valid, dst_base, X, Y0);
}

// ---------------------------------------------------------------------------------------------
// windowed ("collapsible") filters: sinc, Lanczos, Gaussian, Hann, Tukey, tent-with-oblur
// ---------------------------------------------------------------------------------------------
// helpers.pyx:703-735, 792-835.  A size x size region at floor(v) - (size/2 - 1) is weighted by a
// separable window of the sub-pixel offset and collapsed one axis at a time, X first:
//     row_j  = sum_i(kx_i * tap_ji) / sum_i(kx_i)          out = sum_j(ky_j * row_j) / sum_j(ky_j)
// One thread per output pixel.  The S x-weights of a pixel are evaluated once and parked in shared
// memory ([i][thread] layout, conflict-free), the y-weight of a row is evaluated when the row is
// consumed: 2S window evaluations for S^2 taps.
//
// Exact mode reproduces NumPy's reduction ORDER (pinned against the compiled reference, see
// oracle_np.interp_filter): the x normaliser is a sequential sum (strided axis), the other three
// reductions use NumPy's pairwise scheme -- 8 interleaved partial sums over the first 8*floor(S/8)
// terms, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the remainder sequentially.
// NumPy's pairwise reduction over n streamed terms; term(i) is called exactly once per i, in order.
template <class TermFn>
__device__ __forceinline__ double np_pairwise_sum(int n, TermFn term)
{
    const int nb = n >> 3;
    double res;
    int i = 0;
    if (nb == 0) {
        res = term(0);
        i = 1;
    } else {
        double r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = term(k);
        for (int b = 1; b < nb; ++b) {
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], term(b * 8 + k));
        }
        res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                        __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        i = nb * 8;
    }
    for (; i < n; ++i) res = __dadd_rn(res, term(i));
    return res;
}

// two reductions of the same length in one pass (term(i, a, b) yields both addends)
template <class TermFn>
__device__ __forceinline__ void np_pairwise_sum2(int n, TermFn term, double &sa, double &sb)
{
    const int nb = n >> 3;
    int i = 0;
    if (nb == 0) {
        term(0, sa, sb);
        i = 1;
    } else {
        double r[8], q[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) term(k, r[k], q[k]);
        for (int b = 1; b < nb; ++b) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                double ta, tb;
                term(b * 8 + k, ta, tb);
                r[k] = __dadd_rn(r[k], ta);
                q[k] = __dadd_rn(q[k], tb);
            }
        }
        sa = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                       __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        sb = __dadd_rn(__dadd_rn(__dadd_rn(q[0], q[1]), __dadd_rn(q[2], q[3])),
                       __dadd_rn(__dadd_rn(q[4], q[5]), __dadd_rn(q[6], q[7])));
        i = nb * 8;
    }
    for (; i < n; ++i) {
        double ta, tb;
        term(i, ta, tb);
        sa = __dadd_rn(sa, ta);
        sb = __dadd_rn(sb, tb);
    }
}

// one collapsed row (exact mode): sum_i kx_i * tap(i), pairwise order.  Kept out of line: the y
// reduction instantiates it once per unrolled slot.
template <typename SrcT, bool GENERIC>
__device__ __noinline__ double filt_row_exact(const ResampleParams &P, const char *src_base, const double *kx,
                                              int S, long long gx0, long long gy, bool okx, bool interior)
{
    if (interior) {
        const SrcT *row = reinterpret_cast<const SrcT *>(src_base + (gy - P.src_y0) * P.src_pitch) + (gx0 - P.src_x0);
        return np_pairwise_sum(S, [&](int i) { return __dmul_rn(kx[i * 256], (double)__ldg(row + i)); });
    }
    return np_pairwise_sum(S, [&](int i) {
        return __dmul_rn(kx[i * 256], tap_generic<SrcT, double>(P, src_base, okx ? gx0 + i : gx0, gy));
    });
}

template <typename SrcT, typename DstT, class A, typename RegT, int SPH>
__global__ void __launch_bounds__(256)
k_resample_filter(const __grid_constant__ ResampleParams P, const __grid_constant__ FilterParams F)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool FASTI = (sizeof(RegT) == 4);
    RegT *const kx = reinterpret_cast<RegT *>(smem_raw) + threadIdx.x;          // kx[i * 256]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const char *const src_base = static_cast<const char *>(P.src) + (long long)blockIdx.z * P.src_plane;
    char *const dst_base = static_cast<char *>(P.dst) + (long long)blockIdx.z * P.dst_plane;
    const int X = blockIdx.x * 32 + ((warp & 1) << 4) + (lane & 15);
    const int Y = blockIdx.y * 8 + ((warp >> 1) << 1) + (lane >> 4);
    if (X >= P.dst_w || Y >= P.dst_h) return;

    double cx[1] = {(double)(P.dst_x0 + X)}, cy[1] = {(double)(P.dst_y0 + Y)};
    chain_eval_n<A, SPH, 1>(P.chain, cx, cy);
    const double x = cx[0], y = cy[0];
    const double fx = floor(x), fy = floor(y);
    const int S = F.size;
    // tap bases; |f| >= 2^51 or NaN: all taps share the (invalid) base index, as in k_resample
    const bool okx = fabs(fx) < 2.0e15, oky = fabs(fy) < 2.0e15;
    const long long gx0 = okx ? cast_i64(fx) - F.offset : cast_i64(fx);
    const long long gy0 = oky ? cast_i64(fy) - F.offset : cast_i64(fy);
    // whole region inside the staged data: plain loads, no boundary logic
    const bool interior = okx && oky && gx0 >= P.src_x0 && gy0 >= P.src_y0 &&
                          gx0 + S <= (long long)P.src_x0 + P.src_w && gy0 + S <= (long long)P.src_y0 + P.src_h;
    DstT *const out = reinterpret_cast<DstT *>(dst_base + (long long)Y * P.dst_pitch) + X;
    // truncated axis with the whole region out of range (or an invalid coordinate): every tap is
    // the fill value -- one collapsed row serves all S rows, and nothing is loaded
    const bool allfill =
        (P.bound_x == TFM_B_TRUNCATE && (!okx || gx0 >= (long long)P.full_w || gx0 + S <= 0)) ||
        (P.bound_y == TFM_B_TRUNCATE && (!oky || gy0 >= (long long)P.full_h || gy0 + S <= 0));

    if (!FASTI) {
        const double bx1 = __dsub_rn(1.0, __dsub_rn(x, fx)), by1 = __dsub_rn(1.0, __dsub_rn(y, fy));
        double kxs = 0.0;                                                        // sequential (strided axis)
        for (int i = 0; i < S; ++i) {
            const double k = window_exact(F.method, filt_of(bx1, F, i));
            kx[i * 256] = (RegT)k;
            kxs = (i == 0) ? k : __dadd_rn(kxs, k);
        }
        const double *kxd = reinterpret_cast<const double *>(kx);
        double row_fill = 0.0;
        if (allfill) row_fill = np_pairwise_sum(S, [&](int i) { return __dmul_rn(kxd[i * 256], P.fill); });
        double acc, kys;
        np_pairwise_sum2(S, [&](int j, double &ta, double &tb) {
            const double ky = window_exact(F.method, filt_of(by1, F, j));
            const double row = allfill ? row_fill
                                       : filt_row_exact<SrcT, true>(P, src_base, kxd, S, gx0, oky ? gy0 + j : gy0, okx, interior);
            ta = __dmul_rn(ky, __ddiv_rn(row, kxs));
            tb = ky;
        }, acc, kys);
        *out = (DstT)__ddiv_rn(acc, kys);
    } else {
        if (allfill) {
            *out = (DstT)((float)P.fill + (float)((x - x) + (y - y)));          // NaN / inf coordinates stay NaN
            return;
        }
        const float bx1 = 1.0f - (float)(x - fx), by1 = 1.0f - (float)(y - fy);
        const float iob = 1.0f / (float)F.oblur, o1 = (float)(F.offset + 1);
        float kxs = 0.0f;
        WinSeq wx, wy;
        wx.init(F.method, bx1 - o1, iob);
        wy.init(F.method, by1 - o1, iob);
        for (int i = 0; i < S; ++i) {
            const float k = wx.next(i);
            kx[i * 256] = (RegT)k;
            kxs += k;
        }
        const float *kxf = reinterpret_cast<const float *>(kx);
        float acc = 0.0f, kys = 0.0f;
        for (int j = 0; j < S; ++j) {
            const float ky = wy.next(j);
            kys += ky;
            float row = 0.0f;
            if (interior) {
                // S is even: two independent chains, several taps in flight per iteration
                const SrcT *r = reinterpret_cast<const SrcT *>(src_base + (gy0 + j - P.src_y0) * P.src_pitch) + (gx0 - P.src_x0);
                float ra = 0.0f, rb = 0.0f;
#pragma unroll 4
                for (int i = 0; i < S; i += 2) {
                    ra = fmaf(kxf[i * 256], (float)__ldg(r + i), ra);
                    rb = fmaf(kxf[(i + 1) * 256], (float)__ldg(r + i + 1), rb);
                }
                row = ra + rb;
            } else {
                for (int i = 0; i < S; ++i)
                    row = fmaf(kxf[i * 256], tap_generic<SrcT, float>(P, src_base, okx ? gx0 + i : gx0, oky ? gy0 + j : gy0), row);
            }
            acc = fmaf(ky, row, acc);
        }
        *out = (DstT)(acc / (kxs * kys));
    }
}

template <typename SrcT, typename DstT, class A, typename RegT>
cudaError_t launch_filter(const ResampleParams &P, const FilterParams &F, cudaStream_t st, const char **name)
{
    dim3 grid((unsigned)((P.dst_w + 31) / 32), (unsigned)((P.dst_h + 7) / 8), (unsigned)P.n_planes);
    if (grid.y > 65535u || grid.z > 65535u) return cudaErrorInvalidConfiguration;
    const size_t smem = (size_t)256 * (size_t)F.size * sizeof(RegT);
    const int sph = chain_sph_level(P.chain);
    auto go = [&](auto kern) -> cudaError_t {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<grid, 256, smem, st>>>(P, F);
        return cudaGetLastError();
    };
    cudaError_t e = sph == 2 ? go(k_resample_filter<SrcT, DstT, A, RegT, 2>)
                    : sph == 1 ? go(k_resample_filter<SrcT, DstT, A, RegT, 1>)
                               : go(k_resample_filter<SrcT, DstT, A, RegT, 0>);
    count_launch();
    *name = "k_resample_filter";
    return e;
}

// ---------------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------------
template <int METHOD, typename SrcT, typename DstT, class A, typename RegT>
cudaError_t launch2(ResampleParams &P, bool generic, bool aff, cudaStream_t st, const char **name)
{
    constexpr int NPX = MethodTraits<METHOD>::NPX;
    constexpr int NT = MethodTraits<METHOD>::NT;
    dim3 grid((unsigned)((P.dst_w + 31) / 32), (unsigned)((P.dst_h + 8 * NPX - 1) / (8 * NPX)),
              (unsigned)P.n_planes);
    if (grid.y > 65535u || grid.z > 65535u) return cudaErrorInvalidConfiguration;
    P.int_w = P.src_w - NT + 1 > 0 ? P.src_w - NT + 1 : 0;
    P.int_h = P.src_h - NT + 1 > 0 ? P.src_h - NT + 1 : 0;
    const int sph = chain_sph_level(P.chain);
    if (generic) {
        if (aff && !A::exact) {
            // same host-composed coordinates as the non-generic fast kernels: a staged shard must
            // reproduce the unsharded fast-mode result bit for bit
            if constexpr (!A::exact) {
                k_resample<METHOD, SrcT, DstT, Fast, RegT, true, true, 0><<<grid, 256, 0, st>>>(P);
            }
        } else if (sph == 2) k_resample<METHOD, SrcT, DstT, A, RegT, true, false, 2><<<grid, 256, 0, st>>>(P);
        else if (sph == 1) k_resample<METHOD, SrcT, DstT, A, RegT, true, false, 1><<<grid, 256, 0, st>>>(P);
        else k_resample<METHOD, SrcT, DstT, A, RegT, true, false, 0><<<grid, 256, 0, st>>>(P);
        *name = "k_resample<generic>";
    } else if (aff && !A::exact) {
        if constexpr (!A::exact) {
            k_resample<METHOD, SrcT, DstT, Fast, RegT, false, true, 0><<<grid, 256, 0, st>>>(P);
        }
        *name = "k_resample<trunc,affine>";
    } else {
        if (sph == 2) k_resample<METHOD, SrcT, DstT, A, RegT, false, false, 2><<<grid, 256, 0, st>>>(P);
        else if (sph == 1) k_resample<METHOD, SrcT, DstT, A, RegT, false, false, 1><<<grid, 256, 0, st>>>(P);
        else k_resample<METHOD, SrcT, DstT, A, RegT, false, false, 0><<<grid, 256, 0, st>>>(P);
        *name = "k_resample<trunc>";
    }
    count_launch();
    return cudaGetLastError();
}

template <typename SrcT, typename DstT, class A, typename RegT>
cudaError_t launch1(int method, ResampleParams &P, bool generic, bool aff, cudaStream_t st,
                           const char **name)
{
    switch (method) {
    case TFM_NEAREST: return launch2<TFM_NEAREST, SrcT, DstT, A, RegT>(P, generic, aff, st, name);
    case TFM_LINEAR:  return launch2<TFM_LINEAR, SrcT, DstT, A, RegT>(P, generic, aff, st, name);
    case TFM_WIN2:    return launch2<TFM_WIN2, SrcT, DstT, A, RegT>(P, generic, aff, st, name);
    default:          return launch2<TFM_CUBIC, SrcT, DstT, A, RegT>(P, generic, aff, st, name);
    }
}


}  // namespace tfm
