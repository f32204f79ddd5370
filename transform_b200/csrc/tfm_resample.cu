// tfm_resample.cu -- K1: fused inverse-chain evaluation + nearest / linear / cubic gather.
//
// One launch replaces Transform.resample's three whole-array stages (core.py:573-584):
//   grid enumeration (core.py:574-578)      -> pixel index from block/thread index, zero bytes
//   self.invert(grid) (core.py:1556-1559)   -> tfm::chain_eval in registers
//   interpND(...) (helpers.pyx:379-844)     -> boundary-conditioned taps + weights, fused
// HBM traffic is the compulsory one: every source pixel read once (L1/L2 absorb the 4x/16x tap
// reuse), every output pixel written once.  No tensor cores: nothing here is a contraction.
//
// Thread mapping ("direct" variant): a CTA of 256 threads covers a 32 x (8*NPX) output tile.  A warp
// covers a WW x (32/WW) patch (WW = 32, 16 or 8, chosen per launch from the local rotation of the
// map) so that one warp-wide tap load touches as few 128-byte lines as possible, and repeats it
// NPX times 8 rows apart with all loads of the NPX pixels in flight together.
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
__device__ __noinline__ int bound_slow(long long i, int size, int mode, bool &trunc, int *status)
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
    return (fabs(f) < 1.0e9) ? __double2int_rz(f) : -1000000;
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

// floor + fractional part without the conversion (XU) pipe: t = x + 1.5*2^52 rounds x to the nearest
// integer in the low mantissa bits; one compare-and-correct turns round-to-nearest into floor.
// Valid for |x| < 2^31; callers route anything else to an out-of-range index first.
__device__ __forceinline__ void floor_frac_fast(double x, int &i, float &frac)
{
    const double MAGIC = 6755399441055744.0;
    const double t = x + MAGIC;
    i = __double2loint(t);                // rn(x)
    float f = (float)(x - (t - MAGIC));   // x - rn(x) in [-0.5, 0.5], exact before the conversion
    if (f < 0.0f) { f += 1.0f; i -= 1; }  // round-to-nearest -> floor (fp32/ALU pipes, not fp64)
    frac = f;
}

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
        // fp32 is plenty here: |coordinate| < 1e9 is host-proven and the test carries a margin of
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
        const double Xd = (double)(P.dst_x0 + X), Yd = (double)(P.dst_y0 + Y0);
        const double bx = fma(P.aff[0], Xd, fma(P.aff[1], Yd, P.aff[4]));
        const double by = fma(P.aff[2], Xd, fma(P.aff[3], Yd, P.aff[5]));
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
            cx[j] = fma((double)(8 * j), P.aff[1], bx);
            cy[j] = fma((double)(8 * j), P.aff[3], by);
        }
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
            else sane[j] = (fabs(x) < 1.0e9) && (fabs(y) < 1.0e9);
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
            const bool ok = sane[j] && (!AFF || P.aff_bounded || ((fabs(cx[j]) < 1.0e9) && (fabs(cy[j]) < 1.0e9)));
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
            if (FASTI && fabs(x) < 1.0e9 && fabs(y) < 1.0e9) {
                // every fast-mode variant (LDG gather, texture gather, this generic one) must round the
                // fp32 fraction identically, so that a sharded run reproduces the unsharded one bit for bit
                int iu;
                floor_frac_fast(x, iu, ffx[j]);
                floor_frac_fast(y, iu, ffy[j]);
            }
            // |f| >= 2^51 (or NaN): the reference's float tap offsets are absorbed / invalid; every
            // tap shares the base index ("absurd coordinates").
            const bool okx = fabs(fx) < 2.0e15, oky = fabs(fy) < 2.0e15;
            const long long bx = cast_i64(fx), by = cast_i64(fy);
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
// texture-gather variant (fp32 fast path, linear, bound = truncate with fill 0, whole image)
// ---------------------------------------------------------------------------------------------
// The texture unit does what the ALU pipe was saturated with in k_resample: address generation,
// the out-of-range test of every tap (border mode returns 0 = the reference's truncate boundary with
// its default fillvalue) and the fetch of the whole 2x2 footprint in ONE instruction (TLD4).  What is
// left per pixel: 2 DFMA (coordinates), the XU-free floor/fraction, 6 fp32 ops of interpolation in
// the same order as k_resample, one store.  No interior/outside classification at all.
template <class A, bool AFF, int SPH>
__global__ void __launch_bounds__(256)
k_resample_tex(const __grid_constant__ ResampleParams P, cudaTextureObject_t tex)
{
    constexpr int NPX = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    char *const dst_base = static_cast<char *>(P.dst);
    const int px = ((warp & 1) << 4) + (lane & 15);
    const int py = ((warp >> 1) << 1) + (lane >> 4);
    const int X = blockIdx.x * 32 + px;
    const int Y0 = blockIdx.y * (8 * NPX) + py;
    const bool xin = X < P.dst_w;

    if (AFF && P.aff_bounded) {            // CTA-uniform rejection of empty tiles (see k_resample)
        const float tx0 = (float)(int)(blockIdx.x * 32) + P.aff_f[6], tx1 = tx0 + 31.0f;
        const float ty0 = (float)(int)(blockIdx.y * (8 * NPX)) + P.aff_f[7], ty1 = ty0 + (float)(8 * NPX - 1);
        const float ux0 = P.aff_f[0] * tx0, ux1 = P.aff_f[0] * tx1, uy0 = P.aff_f[1] * ty0, uy1 = P.aff_f[1] * ty1;
        const float vx0 = P.aff_f[2] * tx0, vx1 = P.aff_f[2] * tx1, vy0 = P.aff_f[3] * ty0, vy1 = P.aff_f[3] * ty1;
        const float xmin = fminf(ux0, ux1) + fminf(uy0, uy1) + P.aff_f[4], xmax = fmaxf(ux0, ux1) + fmaxf(uy0, uy1) + P.aff_f[4];
        const float ymin = fminf(vx0, vx1) + fminf(vy0, vy1) + P.aff_f[5], ymax = fmaxf(vx0, vx1) + fmaxf(vy0, vy1) + P.aff_f[5];
        const float pad = 4.0f + P.aff_margin;
        if (xmax < -pad || ymax < -pad || xmin > (float)P.src_w + pad || ymin > (float)P.src_h + pad) {
            if (xin) {
#pragma unroll
                for (int j = 0; j < NPX; ++j) {
                    const int Y = Y0 + 8 * j;
                    if (Y < P.dst_h) reinterpret_cast<float *>(dst_base + (long long)Y * P.dst_pitch)[X] = 0.0f;
                }
            }
            return;
        }
    }

    double cx[NPX], cy[NPX];
    if (AFF) {
        const double Xd = (double)(P.dst_x0 + X), Yd = (double)(P.dst_y0 + Y0);
        const double bx = fma(P.aff[0], Xd, fma(P.aff[1], Yd, P.aff[4]));
        const double by = fma(P.aff[2], Xd, fma(P.aff[3], Yd, P.aff[5]));
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
            cx[j] = fma((double)(8 * j), P.aff[1], bx);
            cy[j] = fma((double)(8 * j), P.aff[3], by);
        }
    } else if (P.chain.n > 0 && P.chain.ops[0].kind == TFM_OP_POLAR) {
        // polar prefix: 32 + 32 sincos per tile through the angle-addition tables (tfm_chain.cuh)
        __shared__ double2 s_ta[32], s_tb[32];
        __shared__ double s_ea[32], s_eb[32];
        const tfm_op &op0 = P.chain.ops[0];
        const double TX0 = (double)(P.dst_x0 + (long long)blockIdx.x * 32);
        const double TY0 = (double)(P.dst_y0 + (long long)blockIdx.y * 32);
        polar_tables(op0, TX0, TY0, 32, 32, s_ta, s_tb, s_ea, s_eb, threadIdx.x, 256);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NPX; ++j)
            polar_from_tables(op0, TX0, TY0, px, py + 8 * j, s_ta, s_tb, s_ea, s_eb, cx[j], cy[j]);
        chain_eval_n<A, SPH, NPX>(P.chain, cx, cy, 1);
    } else {
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
            cx[j] = (double)(P.dst_x0 + X);
            cy[j] = (double)(P.dst_y0 + Y0 + 8 * j);
        }
        chain_eval_n<A, SPH, NPX>(P.chain, cx, cy);
    }

    float4 g[NPX];
    float fx[NPX], fy[NPX];
#pragma unroll
    for (int j = 0; j < NPX; ++j) {
        int ix, iy;
        floor_frac_fast(cx[j], ix, fx[j]);
        floor_frac_fast(cy[j], iy, fy[j]);
        // NaN / inf / absurd coordinate -> far outside the texture (border = 0); the compare is only
        // needed when the host could not bound the coordinates
        if (!(AFF && P.aff_bounded)) {
            if (!((fabs(cx[j]) < 1.0e9) && (fabs(cy[j]) < 1.0e9))) { ix = -100000; iy = -100000; fx[j] = 0.f; fy[j] = 0.f; }
        }
        // unnormalised coordinates: the bilinear footprint of (u,v) is texels floor(u-0.5)+{0,1},
        // so (ix+1, iy+1) selects columns ix, ix+1 and rows iy, iy+1
        g[j] = tex2Dgather<float4>(tex, (float)ix + 1.0f, (float)iy + 1.0f, 0);
    }
#pragma unroll
    for (int j = 0; j < NPX; ++j) {
        const int Y = Y0 + 8 * j;
        if (!xin || Y >= P.dst_h) continue;
        // gather order: x = (i0,j1), y = (i1,j1), z = (i1,j0), w = (i0,j0)
        const float v00 = g[j].w, v01 = g[j].z, v10 = g[j].x, v11 = g[j].y;
        const float top = fmaf(fx[j], v01 - v00, v00);
        const float bot = fmaf(fx[j], v11 - v10, v10);
        reinterpret_cast<float *>(dst_base + (long long)Y * P.dst_pitch)[X] = fmaf(fy[j], bot - top, top);
    }
}

// Texture objects over caller memory are cached per (device, pointer, geometry): creating one costs
// microseconds, a resample launch ~100.  A texture object refers to an address range, not to its
// contents, so reuse after the caller rewrote (or re-allocated) the buffer is safe.
struct TexKey {
    int dev; const void *ptr; int w, h; long long pitch;
    bool operator==(const TexKey &o) const { return dev == o.dev && ptr == o.ptr && w == o.w && h == o.h && pitch == o.pitch; }
};
static std::mutex g_tex_mu;
static TexKey g_tex_keys[16];
static cudaTextureObject_t g_tex_objs[16];
static int g_tex_n = 0, g_tex_next = 0;

static bool get_texture(const void *ptr, int w, int h, long long pitch, cudaTextureObject_t *out)
{
    int dev = 0;
    cudaGetDevice(&dev);
    TexKey key{dev, ptr, w, h, pitch};
    std::lock_guard<std::mutex> lock(g_tex_mu);
    for (int i = 0; i < g_tex_n; ++i)
        if (g_tex_keys[i] == key) { *out = g_tex_objs[i]; return true; }
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = const_cast<void *>(ptr);
    rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
    rd.res.pitch2D.width = (size_t)w;
    rd.res.pitch2D.height = (size_t)h;
    rd.res.pitch2D.pitchInBytes = (size_t)pitch;
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = cudaAddressModeBorder;
    td.addressMode[1] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    cudaTextureObject_t t = 0;
    if (cudaCreateTextureObject(&t, &rd, &td, nullptr) != cudaSuccess) { cudaGetLastError(); return false; }
    int slot;
    if (g_tex_n < 16) slot = g_tex_n++;
    else { slot = g_tex_next; g_tex_next = (g_tex_next + 1) & 15; cudaDestroyTextureObject(g_tex_objs[slot]); }
    g_tex_keys[slot] = key;
    g_tex_objs[slot] = t;
    *out = t;
    return true;
}

// ---------------------------------------------------------------------------------------------
// TMA-staged variant (fp32 fast path, linear, affine map, bound = truncate with fill 0, whole image)
// ---------------------------------------------------------------------------------------------
// An affine map sends a 32 x 32 output tile to a parallelogram whose bounding box has a size the host
// knows in advance.  One elected thread asks the Tensor Memory Accelerator for exactly that box
// (cp.async.bulk.tensor.2d, completion on an mbarrier); out-of-range parts of the box are zero-filled
// by the TMA unit, which IS the reference's truncate boundary with its default fillvalue -- no
// per-tap classification anywhere.  While the box is in flight all threads evaluate their
// coordinates; then every tap is a shared-memory load.  HBM sees each source line once per tile
// (halo overlaps are L2 hits), registers are not staged through, and neither the LSU address path
// nor the texture write-back (the two limits of the other variants) is on the critical path.
// Status: correct (bit-identical to the other variants, tests), but measured slower than the texture
// gather in this one-box-per-CTA form (see the dispatch below); kept behind a tuning bit.
struct TmaBox {
    int bw, bh;                              // box columns / rows (bw * 4 bytes is a multiple of 16)
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int DUMMY = 0>
__global__ void __launch_bounds__(256)
k_resample_tma(const __grid_constant__ ResampleParams P, const __grid_constant__ CUtensorMap tmap,
               const __grid_constant__ TmaBox B)
{
    constexpr int NPX = 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // the box first (TMA needs a 128-byte aligned destination: no static shared memory in this kernel,
    // so the dynamic window starts at offset 0), the mbarrier behind it
    float *const tile = reinterpret_cast<float *>(smem_raw);
    unsigned long long *const s_mbar_p = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)B.bw * B.bh * 4);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    char *const dst_base = static_cast<char *>(P.dst);
    const int px = ((warp & 1) << 4) + (lane & 15);
    const int py = ((warp >> 1) << 1) + (lane >> 4);
    const int X = blockIdx.x * 32 + px;
    const int Y0 = blockIdx.y * (8 * NPX) + py;
    const bool xin = X < P.dst_w;

    // CTA-uniform rejection of empty tiles (see k_resample)
    {
        const float tx0 = (float)(int)(blockIdx.x * 32) + P.aff_f[6], tx1 = tx0 + 31.0f;
        const float ty0 = (float)(int)(blockIdx.y * 32) + P.aff_f[7], ty1 = ty0 + 31.0f;
        const float ux0 = P.aff_f[0] * tx0, ux1 = P.aff_f[0] * tx1, uy0 = P.aff_f[1] * ty0, uy1 = P.aff_f[1] * ty1;
        const float vx0 = P.aff_f[2] * tx0, vx1 = P.aff_f[2] * tx1, vy0 = P.aff_f[3] * ty0, vy1 = P.aff_f[3] * ty1;
        const float xmin = fminf(ux0, ux1) + fminf(uy0, uy1) + P.aff_f[4], xmax = fmaxf(ux0, ux1) + fmaxf(uy0, uy1) + P.aff_f[4];
        const float ymin = fminf(vx0, vx1) + fminf(vy0, vy1) + P.aff_f[5], ymax = fmaxf(vx0, vx1) + fmaxf(vy0, vy1) + P.aff_f[5];
        const float pad = 4.0f + P.aff_margin;
        if (xmax < -pad || ymax < -pad || xmin > (float)P.src_w + pad || ymin > (float)P.src_h + pad) {
            if (xin) {
#pragma unroll
                for (int j = 0; j < NPX; ++j) {
                    const int Y = Y0 + 8 * j;
                    if (Y < P.dst_h) reinterpret_cast<float *>(dst_base + (long long)Y * P.dst_pitch)[X] = 0.0f;
                }
            }
            return;
        }
    }

    // source box of this tile: the map is affine, so the extremes are at the tile corners; one pixel
    // of margin on the low side, the host-chosen box size covers the high side
    const double TX = (double)(P.dst_x0 + (long long)blockIdx.x * 32), TY = (double)(P.dst_y0 + (long long)blockIdx.y * 32);
    const double cx0 = fma(P.aff[0], TX, fma(P.aff[1], TY, P.aff[4]));
    const double cy0 = fma(P.aff[2], TX, fma(P.aff[3], TY, P.aff[5]));
    const double xlo = cx0 + fmin(31.0 * P.aff[0], 0.0) + fmin(31.0 * P.aff[1], 0.0);
    const double ylo = cy0 + fmin(31.0 * P.aff[2], 0.0) + fmin(31.0 * P.aff[3], 0.0);
    // the box must start on a 16-byte boundary of the innermost axis (measured: an unaligned start
    // coordinate raises an illegal-instruction fault), hence the round-down to a multiple of 4 pixels
    const int bx0 = ((int)floor(xlo) - 1) & ~3, by0 = (int)floor(ylo) - 1;     // |coordinate| < 1e9: host-proven

    const uint32_t mbar = smem_addr(s_mbar_p);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = (uint32_t)(B.bw * B.bh * 4);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
            ::"r"(smem_addr(tile)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(bx0), "r"(by0), "r"(mbar)
            : "memory");
    }

    // coordinates while the box is in flight
    float fx[NPX], fy[NPX];
    int lx[NPX], ly[NPX];
    {
        const double Xd = (double)(P.dst_x0 + X), Yd = (double)(P.dst_y0 + Y0);
        const double bx = fma(P.aff[0], Xd, fma(P.aff[1], Yd, P.aff[4]));
        const double by = fma(P.aff[2], Xd, fma(P.aff[3], Yd, P.aff[5]));
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
            int ix, iy;
            floor_frac_fast(fma((double)(8 * j), P.aff[1], bx), ix, fx[j]);
            floor_frac_fast(fma((double)(8 * j), P.aff[3], by), iy, fy[j]);
            lx[j] = ix - bx0;
            ly[j] = iy - by0;
        }
    }
    __syncthreads();                                   // the barrier's initialisation is visible to all
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(mbar), "r"(0u) : "memory");
        }
    }

    const int bw = B.bw;
#pragma unroll
    for (int j = 0; j < NPX; ++j) {
        const int Y = Y0 + 8 * j;
        if (!xin || Y >= P.dst_h) continue;
        float v00, v01, v10, v11;
        if ((unsigned)lx[j] < (unsigned)(bw - 1) && (unsigned)ly[j] < (unsigned)(B.bh - 1)) {
            const float *t = tile + ly[j] * bw + lx[j];
            v00 = t[0]; v01 = t[1]; v10 = t[bw]; v11 = t[bw + 1];
        } else {
            // cannot happen for in-range arithmetic; kept so that a rounding surprise reads the right
            // texel from global memory instead of a wrong one from the box
            const int ix = lx[j] + bx0, iy = ly[j] + by0;
            v00 = tap_trunc<float, float>(P, P.src, ix, iy);
            v01 = tap_trunc<float, float>(P, P.src, ix + 1, iy);
            v10 = tap_trunc<float, float>(P, P.src, ix, iy + 1);
            v11 = tap_trunc<float, float>(P, P.src, ix + 1, iy + 1);
        }
        const float top = fmaf(fx[j], v01 - v00, v00);
        const float bot = fmaf(fx[j], v11 - v10, v10);
        reinterpret_cast<float *>(dst_base + (long long)Y * P.dst_pitch)[X] = fmaf(fy[j], bot - top, top);
    }
}

// Tensor maps over caller memory, cached like the texture objects (the map describes an address range
// and a box shape, not contents).  cuTensorMapEncodeTiled comes from the driver through the runtime's
// entry-point query: no link-time dependency on libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct TmapKey {
    int dev; const void *ptr; int w, h; long long pitch; int bw, bh;
    bool operator==(const TmapKey &o) const
    { return dev == o.dev && ptr == o.ptr && w == o.w && h == o.h && pitch == o.pitch && bw == o.bw && bh == o.bh; }
};
static std::mutex g_tmap_mu;
static TmapKey g_tmap_keys[16];
static CUtensorMap g_tmap_maps[16];
static int g_tmap_n = 0, g_tmap_next = 0;

static bool get_tensor_map(const void *ptr, int w, int h, long long pitch, int bw, int bh, CUtensorMap *out)
{
    static EncodeTiledFn encode = nullptr;
    static bool looked_up = false;
    int dev = 0;
    cudaGetDevice(&dev);
    TmapKey key{dev, ptr, w, h, pitch, bw, bh};
    std::lock_guard<std::mutex> lock(g_tmap_mu);
    if (!looked_up) {
        looked_up = true;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<EncodeTiledFn>(fn);
        else
            cudaGetLastError();
    }
    if (!encode) return false;
    for (int i = 0; i < g_tmap_n; ++i)
        if (g_tmap_keys[i] == key) { *out = g_tmap_maps[i]; return true; }
    CUtensorMap m;
    const cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)h};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(ptr), gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    int slot;
    if (g_tmap_n < 16) slot = g_tmap_n++;
    else { slot = g_tmap_next; g_tmap_next = (g_tmap_next + 1) & 15; }
    g_tmap_keys[slot] = key;
    g_tmap_maps[slot] = m;
    *out = m;
    return true;
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
        for (int i = 0; i < S; ++i) {
            const float k = window_fast(F.method, (bx1 - o1 + (float)i) * iob);
            kx[i * 256] = (RegT)k;
            kxs += k;
        }
        const float *kxf = reinterpret_cast<const float *>(kx);
        float acc = 0.0f, kys = 0.0f;
        for (int j = 0; j < S; ++j) {
            const float ky = window_fast(F.method, (by1 - o1 + (float)j) * iob);
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
static cudaError_t launch_filter(const ResampleParams &P, const FilterParams &F, cudaStream_t st, const char **name)
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
static cudaError_t launch2(ResampleParams &P, bool generic, bool aff, cudaStream_t st, const char **name)
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
static cudaError_t launch1(int method, ResampleParams &P, bool generic, bool aff, cudaStream_t st,
                           const char **name)
{
    switch (method) {
    case TFM_NEAREST: return launch2<TFM_NEAREST, SrcT, DstT, A, RegT>(P, generic, aff, st, name);
    case TFM_LINEAR:  return launch2<TFM_LINEAR, SrcT, DstT, A, RegT>(P, generic, aff, st, name);
    case TFM_WIN2:    return launch2<TFM_WIN2, SrcT, DstT, A, RegT>(P, generic, aff, st, name);
    default:          return launch2<TFM_CUBIC, SrcT, DstT, A, RegT>(P, generic, aff, st, name);
    }
}

// Pre-compose a chain made only of LINEAR / IDENTITY ops into one affine map (fast mode only; the
// fp64-exact mode never pre-fuses, SURVEY.md section 7 "hard parts").  Parameter preparation on the
// host, a handful of flops per call -- not part of the data path.
static bool compose_affine(const tfm_op *ops, int n, double *aff)
{
    for (int i = 0; i < n; ++i)
        if (!is_linear_like(ops[i])) return false;
    compose_linear_run(ops, 0, n, aff, aff + 4);
    return true;
}

int resample_dispatch(const tfm_resample_args *a, cudaStream_t st)
{
    if (!a || a->abi_version != TFM_ABI_VERSION) return TFM_EVALUE;
    if (a->n_ops < 0 || a->n_ops > TFM_MAX_OPS || (a->n_ops && !a->chain)) return TFM_EVALUE;
    if (a->method < TFM_NEAREST || a->method > TFM_TENT) return TFM_EVALUE;            // helpers.pyx:840-844
    const bool is_filter = a->method >= TFM_SINC;
    // helpers.pyx:543-544, 748-749: oblur is rejected by nearest and cubic; 'l' with oblur is TFM_TENT
    if (!is_filter && a->oblur != 0.0) return TFM_EVALUE;
    if (a->method == TFM_TENT && !(a->oblur != 0.0)) return TFM_EVALUE;
    if (is_filter && !(a->oblur >= 0.0)) return TFM_EVALUE;
    if (a->bound_x < TFM_B_FORBID || a->bound_x > TFM_B_MIRROR ||
        a->bound_y < TFM_B_FORBID || a->bound_y > TFM_B_MIRROR) return TFM_EVALUE;     // helpers.pyx:174-176
    const tfm_image &s = a->src, &d = a->dst;
    if (s.h < 0 || s.w < 0 || d.h < 0 || d.w < 0) return TFM_EVALUE;
    if (s.n_planes < 1 || s.n_planes != d.n_planes || s.n_planes > 65535) return TFM_EVALUE;
    if (s.full_h < s.h || s.full_w < s.w || s.full_h > 0x3fffffff || s.full_w > 0x3fffffff ||
        d.h > 0x3fffffff || d.w > 0x3fffffff) return TFM_EVALUE;
    if (d.h == 0 || d.w == 0) return TFM_OK;                                            // ragged / empty output
    if (s.full_h == 0 || s.full_w == 0) return TFM_EVALUE;
    if (!s.data || !d.data) return TFM_EVALUE;
    if (!chain_ops_valid(a->chain, a->n_ops)) return TFM_EVALUE;

    ResampleParams P;
    P.src = s.data; P.dst = d.data;
    P.src_pitch = s.pitch; P.dst_pitch = d.pitch;
    P.src_plane = s.plane_pitch; P.dst_plane = d.plane_pitch; P.n_planes = (int)s.n_planes;
    P.src_h = (int)s.h; P.src_w = (int)s.w;
    P.src_x0 = (int)s.x0; P.src_y0 = (int)s.y0;
    P.full_h = (int)s.full_h; P.full_w = (int)s.full_w;
    P.dst_h = (int)d.h; P.dst_w = (int)d.w;
    P.dst_x0 = d.x0; P.dst_y0 = d.y0;
    P.bound_x = a->bound_x; P.bound_y = a->bound_y;
    P.flags = a->flags;
    P.fill = a->fillvalue;
    P.status = a->status_dev;
    P.chain.pad = 0;
    if (a->precision == TFM_PREC_FAST32) {
        P.chain.n = fuse_chain_fast(a->chain, a->n_ops, P.chain.ops);       // fast mode: peephole fusion
    } else {
        P.chain.n = a->n_ops;
        for (int i = 0; i < a->n_ops; ++i) P.chain.ops[i] = a->chain[i];
    }
    int lw = 5;
    if ((a->tuning & 3) == 2) lw = 4;
    if ((a->tuning & 3) == 3) lw = 3;
    P.lw = lw;

    const bool subrect = (s.x0 != 0 || s.y0 != 0 || s.full_h != s.h || s.full_w != s.w);
    const long long src_elems = s.h * (s.pitch / (s.dtype == TFM_F64 ? 8 : 4));
    const bool generic = subrect || a->bound_x != TFM_B_TRUNCATE || a->bound_y != TFM_B_TRUNCATE ||
                         src_elems >= 0x7fffffffLL;      // the fast path indexes taps with 32-bit offsets
    if ((a->bound_x == TFM_B_FORBID || a->bound_y == TFM_B_FORBID || subrect) && !a->status_dev)
        return TFM_EVALUE;

    if (s.pitch % (s.dtype == TFM_F64 ? 8 : 4) != 0) return TFM_EVALUE;
    const char *name = "";
    cudaError_t e;
    const bool exact = (a->precision == TFM_PREC_EXACT64);
    const bool aff = !exact && !(a->tuning & 4) && compose_affine(a->chain, a->n_ops, P.aff);
    P.aff_bounded = 0;
    if (aff) {
        const double xm = fabs((double)d.x0) + (double)d.w, ym = fabs((double)d.y0) + (double)d.h;
        const double bx = fabs(P.aff[0]) * xm + fabs(P.aff[1]) * ym + fabs(P.aff[4]) + 1.0;
        const double by = fabs(P.aff[2]) * xm + fabs(P.aff[3]) * ym + fabs(P.aff[5]) + 1.0;
        P.aff_bounded = (bx < 1.0e9 && by < 1.0e9) ? 1 : 0;      // NaN parameters compare false
        for (int k = 0; k < 6; ++k) P.aff_f[k] = (float)P.aff[k];
        P.aff_f[6] = (float)d.x0;
        P.aff_f[7] = (float)d.y0;
        P.aff_margin = (float)(4e-6 * (bx > by ? bx : by));      // a few fp32 ulps of the largest term
        if (!(P.aff_margin < 64.0f)) P.aff_bounded = 0;          // huge coordinates: skip the shortcut
    }
    if (is_filter) {
        static const int base_size[] = {16, 6, 8, 2, 2, 2};                           // helpers.pyx:708
        FilterParams F;
        F.method = a->method;
        F.oblur = a->oblur != 0.0 ? a->oblur : 1.0;
        double sz = base_size[a->method - TFM_SINC];
        if (a->oblur != 0.0) sz = 2.0 * ceil(sz / 2.0 * a->oblur);                      // helpers.pyx:711
        if (!(sz >= 2.0)) return TFM_EVALUE;
        if (sz > (double)TFM_FILTER_MAX_SIZE) return TFM_EUNSUPPORTED;
        F.size = (int)sz;
        F.offset = F.size / 2 - 1;
        P.win_method = F.method;
        P.win_oblur = F.oblur;
        if (F.size == 2 && !(a->tuning & 64)) {
            // 2 x 2 regions share k_resample's tiling, warp classification and tap code
            if (exact && s.dtype == TFM_F32 && d.dtype == TFM_F64)
                e = launch1<float, double, Exact, double>(TFM_WIN2, P, generic, false, st, &name);
            else if (exact && s.dtype == TFM_F64 && d.dtype == TFM_F64)
                e = launch1<double, double, Exact, double>(TFM_WIN2, P, generic, false, st, &name);
            else if (!exact && s.dtype == TFM_F32 && d.dtype == TFM_F32)
                e = launch1<float, float, Fast, float>(TFM_WIN2, P, generic, aff, st, &name);
            else
                return TFM_EUNSUPPORTED;
        } else if (exact && s.dtype == TFM_F32 && d.dtype == TFM_F64)
            e = launch_filter<float, double, Exact, double>(P, F, st, &name);
        else if (exact && s.dtype == TFM_F64 && d.dtype == TFM_F64)
            e = launch_filter<double, double, Exact, double>(P, F, st, &name);
        else if (!exact && s.dtype == TFM_F32 && d.dtype == TFM_F32)
            e = launch_filter<float, float, Fast, float>(P, F, st, &name);
        else
            return TFM_EUNSUPPORTED;
    } else if (exact) {
        if (s.dtype == TFM_F32 && d.dtype == TFM_F64)
            e = launch1<float, double, Exact, double>(a->method, P, generic, false, st, &name);
        else if (s.dtype == TFM_F64 && d.dtype == TFM_F64)
            e = launch1<double, double, Exact, double>(a->method, P, generic, false, st, &name);
        else if (s.dtype == TFM_F32 && d.dtype == TFM_F32 && a->method == TFM_NEAREST)
            e = launch2<TFM_NEAREST, float, float, Exact, double>(P, generic, false, st, &name);
        else
            return TFM_EUNSUPPORTED;
    } else if (a->precision == TFM_PREC_FAST32) {
        // texture-gather variant: linear, truncate on both axes with the default fill 0, one plane,
        // geometry within the limits of pitch-linear 2-D textures
        cudaTextureObject_t tex = 0;
        const bool tex_ok = !generic && a->method == TFM_LINEAR && s.dtype == TFM_F32 && d.dtype == TFM_F32 &&
                            a->fillvalue == 0.0 && s.n_planes == 1 && !(a->tuning & 16) &&
                            s.w <= 131072 && s.h <= 65000 && (s.pitch % 32) == 0 &&
                            (reinterpret_cast<uintptr_t>(s.data) % 512) == 0 &&
                            get_texture(s.data, (int)s.w, (int)s.h, s.pitch, &tex);
        // TMA-staged variant: same conditions, host-composed affine map with a bounded source box
        bool tma_done = false;
        // Opt-in (tuning bit 256): measured SLOWER than the texture gather on B200 (8192^2: named chain
        // 0.138 vs 0.090 ms, pure shift 0.227 vs 0.141 ms) -- per-CTA load latency is exposed in this
        // one-box-per-CTA form and the shared-memory taps cost more issue slots than one TLD4.
        if (tex_ok && aff && P.aff_bounded && (a->tuning & 256) && (s.pitch % 16) == 0) {
            const double ex = 31.0 * (fabs(P.aff[0]) + fabs(P.aff[1])), ey = 31.0 * (fabs(P.aff[2]) + fabs(P.aff[3]));
            if (ex < 200.0 && ey < 200.0) {
                TmaBox B;
                // +1 low margin, +2 for floor/+1 tap, +1 slack, +3 for the aligned start; multiple of 4
                B.bw = (((int)ceil(ex) + 7) + 3) & ~3;
                B.bh = (int)ceil(ey) + 4;
                const size_t smem = (size_t)B.bw * B.bh * 4 + 16;          // box + mbarrier
                CUtensorMap tm;
                if (B.bw <= 256 && B.bh <= 256 && smem <= 96 * 1024 &&
                    get_tensor_map(s.data, (int)s.w, (int)s.h, s.pitch, B.bw, B.bh, &tm)) {
                    dim3 grid((unsigned)((P.dst_w + 31) / 32), (unsigned)((P.dst_h + 31) / 32), 1);
                    if (grid.y > 65535u) return TFM_EUNSUPPORTED;
                    cudaError_t ea = cudaSuccess;
                    if (smem > 48 * 1024)
                        ea = cudaFuncSetAttribute(k_resample_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    if (ea == cudaSuccess) {
                        k_resample_tma<0><<<grid, 256, smem, st>>>(P, tm, B);
                        count_launch();
                        name = "k_resample_tma<affine>";
                        e = cudaGetLastError();
                        tma_done = true;
                    }
                }
            }
        }
        if (tma_done) {
        } else if (tex_ok) {
            dim3 grid((unsigned)((P.dst_w + 31) / 32), (unsigned)((P.dst_h + 31) / 32), 1);
            if (grid.y > 65535u) return TFM_EUNSUPPORTED;
            const int sph = chain_sph_level(P.chain);
            if (aff) k_resample_tex<Fast, true, 0><<<grid, 256, 0, st>>>(P, tex);
            else if (sph == 2) k_resample_tex<Fast, false, 2><<<grid, 256, 0, st>>>(P, tex);
            else if (sph == 1) k_resample_tex<Fast, false, 1><<<grid, 256, 0, st>>>(P, tex);
            else k_resample_tex<Fast, false, 0><<<grid, 256, 0, st>>>(P, tex);
            count_launch();
            name = aff ? "k_resample_tex<affine>" : "k_resample_tex";
            e = cudaGetLastError();
        } else if (s.dtype == TFM_F32 && d.dtype == TFM_F32)
            e = launch1<float, float, Fast, float>(a->method, P, generic, aff, st, &name);
        else
            return TFM_EUNSUPPORTED;
    } else {
        return TFM_EVALUE;
    }
    set_last_kernel(name);
    return e == cudaSuccess ? TFM_OK : TFM_ECUDA;
}

}  // namespace tfm
