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
#include "tfm_resample.cuh"

// Which of the two fast affine kernels runs is a performance choice only (bit-identical results): the
// persistent TMA pipeline where most of the output maps inside the source (measured 8192^2, ms: shift
// 0.114 vs 0.137, rot 5 0.114 vs 0.141, rot 45 0.114 vs 0.132, scale 2 0.112 vs 0.127, scale 0.8 [64 %
// coverage] 0.097 vs 0.114), the texture gather where most output tiles are empty (named chain, 42 %
// coverage: 0.082 vs 0.104 -- its CTAs retire on a 10-instruction test, the pipeline's consumers still
// walk every tile).
// Tuning bit 8192 forces the pipeline, 128 / 16 / 256 force the other variants.
#define TFM_PIPE_COVERAGE 0.55

namespace tfm {

// the kernel families are compiled in their own translation units
extern template cudaError_t launch1<float, double, Exact, double>(int, ResampleParams &, bool, bool, cudaStream_t, const char **);
extern template cudaError_t launch1<double, double, Exact, double>(int, ResampleParams &, bool, bool, cudaStream_t, const char **);
extern template cudaError_t launch1<float, float, Fast, float>(int, ResampleParams &, bool, bool, cudaStream_t, const char **);
extern template cudaError_t launch2<TFM_NEAREST, float, float, Exact, double>(ResampleParams &, bool, bool, cudaStream_t, const char **);
extern template cudaError_t launch_filter<float, double, Exact, double>(const ResampleParams &, const FilterParams &, cudaStream_t, const char **);
extern template cudaError_t launch_filter<double, double, Exact, double>(const ResampleParams &, const FilterParams &, cudaStream_t, const char **);
extern template cudaError_t launch_filter<float, float, Fast, float>(const ResampleParams &, const FilterParams &, cudaStream_t, const char **);

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
        // canonical form of every fast-mode kernel: c = fma(a0, X, fma(a1, Y, b)) on exact doubles X, Y
        const double Xd = (double)(P.dst_x0 + X), Yd = (double)(P.dst_y0 + Y0);
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
            const double Yj = Yd + (double)(8 * j);
            cx[j] = fma(P.aff[0], Xd, fma(P.aff[1], Yj, P.aff[4]));
            cy[j] = fma(P.aff[2], Xd, fma(P.aff[3], Yj, P.aff[5]));
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
            if (!((fabs(cx[j]) < TFM_SANE_COORD) && (fabs(cy[j]) < TFM_SANE_COORD))) {
                // far outside the texture (border = 0); a non-finite coordinate makes the weights NaN, as in the
                // reference (alpha = v - floor(v)), an absurd finite one leaves them 0
                ix = -100000; iy = -100000;
                fx[j] = (float)(cx[j] - cx[j]); fy[j] = (float)(cy[j] - cy[j]);
            }
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

// ---------------------------------------------------------------------------------------------
// texture-gather variant, second form: 4 pixels of ONE output row per thread
// ---------------------------------------------------------------------------------------------
// ncu on k_resample_tex (8192^2, pure shift, profiles/r2_kernels.md): 60.8 instructions per pixel at 83 %
// issue utilisation -- the kernel was issue-bound before it was texture-bound.  Of the 60.8, 16 were
// the per-thread share of the empty-tile test, 14 the four stores (each with its own bounds check and
// 64-bit row address), 20 the two floor/fraction extractions.  Here:
//   * a thread owns pixels (X, X+8, X+16, X+24) of row Y: one row pointer, four stores at immediate
//     offsets, the row part fma(a1, Y, b) of the coordinates shared by the four pixels;
//   * a warp covers an 8 x 4 patch per TLD4, so the 32 footprints of one instruction stay compact.
//     (Four CONSECUTIVE pixels per thread and one st.global.v4 were built first: 35.8 instructions
//     per pixel, but 0.192 ms instead of 0.143 -- lanes 4 pixels apart make every TLD4 touch 2.7x the
//     sectors and the L1TEX sector stage saturates at 92 %.  Scalar stores of 8 consecutive pixels
//     x 4 rows fill whole 32-byte sectors anyway.)
//   * the empty-tile test works from the tile centre and a host-computed half-extent;
//   * floor/fraction in fixed point (floor_frac_fast).
// WCSONLY: the chain is exactly one fused CAR -> TAN pair (a TAN -> CAR remap): an instantiation without
// the chain interpreter -- 40 registers instead of 64, twice the resident warps for a latency-bound kernel
// Batches of planes (frames resampled through the same chain): the coordinates, the per-tile tables and the
// floor/fraction of a pixel do not depend on the plane, only the gather, the three lerps and the store do.
// NZ > 1: a CTA serves its tile for up to NZ consecutive planes (texof(z) = texture object of plane z of the
// launch, blockIdx.z selects the group of NZ) -- 73 instructions per pixel for the first plane of a TAN -> CAR
// remap, ~15 for each further one.
constexpr int TEX_BATCH_MAX = 16;          // texture objects per launch
struct TexBatch { cudaTextureObject_t tex[TEX_BATCH_MAX]; int n; };

template <class A, bool AFF, int SPH, bool WCSONLY, int NZ, class TexOf>
__device__ __forceinline__ void resample_tex4_body(const ResampleParams &P, const TexOf texof, const int nz_total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = lane & 7;
    const int py = (warp << 2) + (lane >> 3);
    const int X = blockIdx.x * 32 + px;
    const int Y = blockIdx.y * 32 + py;
    const int z0 = NZ > 1 ? (int)blockIdx.z * NZ : 0;
    const int nz = NZ > 1 ? min(NZ, nz_total - z0) : 1;
    float *drow = reinterpret_cast<float *>(static_cast<char *>(P.dst) + (long long)z0 * P.dst_plane + (long long)Y * P.dst_pitch) + X;
    const bool row_in = Y < P.dst_h;
    const bool full = row_in && (blockIdx.x * 32 + 32 <= (unsigned)P.dst_w);

    // AFF instantiations are launched only when the host has bounded the coordinates (P.aff_bounded)
    if (AFF) {                             // CTA-uniform rejection of empty tiles: centre +- half-extent
        const float xc = (float)(int)(blockIdx.x * 32) + P.aff_f[6], yc = (float)(int)(blockIdx.y * 32) + P.aff_f[7];
        const float cxc = fmaf(P.aff_f[0], xc, fmaf(P.aff_f[1], yc, P.aff_c[0]));
        const float cyc = fmaf(P.aff_f[2], xc, fmaf(P.aff_f[3], yc, P.aff_c[1]));
        if (cxc < -P.aff_c[2] || cxc > P.aff_lim[0] || cyc < -P.aff_c[3] || cyc > P.aff_lim[1]) {
            for (int z = 0; z < nz; ++z, drow = reinterpret_cast<float *>(reinterpret_cast<char *>(drow) + P.dst_plane)) {
                if (full) { drow[0] = 0.f; drow[8] = 0.f; drow[16] = 0.f; drow[24] = 0.f; }
                else if (row_in) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (X + 8 * j < P.dst_w) drow[8 * j] = 0.0f;
                }
            }
            return;
        }
    }

    double cx[4], cy[4];
    if (AFF) {
        // canonical form of every fast-mode kernel: c = fma(a0, X, fma(a1, Y, b)) on exact doubles X, Y
        const double Xd = (double)(P.dst_x0 + X), Yd = (double)(P.dst_y0 + Y);
        const double rx = fma(P.aff[1], Yd, P.aff[4]), ry = fma(P.aff[3], Yd, P.aff[5]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double Xj = Xd + (double)(8 * j);
            cx[j] = fma(P.aff[0], Xj, rx);
            cy[j] = fma(P.aff[2], Xj, ry);
        }
    } else if (!WCSONLY && P.chain.n > 0 && P.chain.ops[0].kind == TFM_OP_POLAR) {
        __shared__ double2 s_ta[32], s_tb[32];
        __shared__ double s_ea[32], s_eb[32];
        const tfm_op &op0 = P.chain.ops[0];
        const double TX0 = (double)(P.dst_x0 + (long long)blockIdx.x * 32);
        const double TY0 = (double)(P.dst_y0 + (long long)blockIdx.y * 32);
        polar_tables(op0, TX0, TY0, 32, 32, s_ta, s_tb, s_ea, s_eb, threadIdx.x, 256);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j)
            polar_from_tables(op0, TX0, TY0, px + 8 * j, py, s_ta, s_tb, s_ea, s_eb, cx[j], cy[j]);
        chain_eval_n<A, SPH, 4>(P.chain, cx, cy, 1);
    } else if (WCSONLY || (SPH == 1 && P.chain.n > 0 && P.chain.ops[0].kind == TFM_OP_WCS_PAIR &&
                           (P.chain.ops[0].flags & 0xffff) == (TFM_WCS_PROJ_CAR | (TFM_WCS_PROJ_TAN << 8)))) {
        // CAR pixel -> sky -> TAN pixel: per-tile sincos tables (see wcs_car_tables, tfm_chain.cuh)
        __shared__ double4 s_t0[32], s_t1[32];
        const tfm_op &op0 = P.chain.ops[0];
        const double TX0 = (double)(P.dst_x0 + (long long)blockIdx.x * 32);
        const double TY0 = (double)(P.dst_y0 + (long long)blockIdx.y * 32);
        const WcsCarCoef kc = wcs_car_coef(op0);
        if (op0.p[3] == 0.0 && op0.p[4] == 0.0) {              // axis-aligned CD (CTA-uniform): separable form
            wcs_car_sep_tables(op0, kc, TX0, TY0, 32, 32, s_t0, s_t1, threadIdx.x, 256);
            __syncthreads();
            const double4 b = s_t1[py];
#pragma unroll
            for (int j = 0; j < 4; ++j) wcs_car_tan_sep(kc, s_t0[px + 8 * j], b, cx[j], cy[j]);
        } else {
            double2 *const s_aphi = reinterpret_cast<double2 *>(s_t0), *const s_ath = s_aphi + 32;
            double2 *const s_bphi = reinterpret_cast<double2 *>(s_t1), *const s_bth = s_bphi + 32;
            wcs_car_tables(op0, TX0, TY0, 32, 32, s_aphi, s_bphi, s_ath, s_bth, threadIdx.x, 256);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j)
                wcs_car_tan_from_tables(kc, px + 8 * j, py, s_aphi, s_bphi, s_ath, s_bth, cx[j], cy[j]);
        }
        if (!WCSONLY) chain_eval_n<A, SPH, 4>(P.chain, cx, cy, 1);
    } else if (!WCSONLY) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            cx[j] = (double)(P.dst_x0 + X + 8 * j);
            cy[j] = (double)(P.dst_y0 + Y);
        }
        chain_eval_n<A, SPH, 4>(P.chain, cx, cy);
    }

    float tcx[4], tcy[4];
    float fx[4], fy[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int ix, iy;
        floor_frac_fast(cx[j], ix, fx[j]);
        floor_frac_fast(cy[j], iy, fy[j]);
        if (!AFF) {                          // NaN / inf / absurd coordinate -> far outside the texture (border = 0)
            if (!((fabs(cx[j]) < TFM_SANE_COORD) && (fabs(cy[j]) < TFM_SANE_COORD))) {
                // far outside the texture (border = 0); a non-finite coordinate makes the weights NaN, as in the
                // reference (alpha = v - floor(v)), an absurd finite one leaves them 0
                ix = -100000; iy = -100000;
                fx[j] = (float)(cx[j] - cx[j]); fy[j] = (float)(cy[j] - cy[j]);
            }
        }
        // texel index inside the staged rectangle (src_x0 = src_y0 = 0 for a whole image)
        tcx[j] = (float)(ix + 1 - P.src_x0);
        tcy[j] = (float)(iy + 1 - P.src_y0);
    }
#pragma unroll(NZ > 1 ? 2 : 1)
    for (int z = 0; z < nz; ++z, drow = reinterpret_cast<float *>(reinterpret_cast<char *>(drow) + P.dst_plane)) {
        const cudaTextureObject_t tex = texof(z0 + z);
        float4 g[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) g[j] = tex2Dgather<float4>(tex, tcx[j], tcy[j], 0);
        float r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // gather order: x = (i0,j1), y = (i1,j1), z = (i1,j0), w = (i0,j0)
            const float v00 = g[j].w, v01 = g[j].z, v10 = g[j].x, v11 = g[j].y;
            const float top = fmaf(fx[j], v01 - v00, v00);
            const float bot = fmaf(fx[j], v11 - v10, v10);
            r[j] = fmaf(fy[j], bot - top, top);
        }
        if (full) { drow[0] = r[0]; drow[8] = r[1]; drow[16] = r[2]; drow[24] = r[3]; }
        else if (row_in) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (X + 8 * j < P.dst_w) drow[8 * j] = r[j];
        }
    }
}

// Resident CTAs: the chain-interpreter instantiation without spherical ops (polar unwrap, pincushion, ...) is bound
// by the latency of its gathers, not by registers it needs -- 40 registers / 6 CTAs per SM instead of 51 / 4:
// polar unwrap, linear, 8192^2: 0.237 -> 0.194 ms (42 % of the HBM roofline).
template <class A, bool AFF, int SPH, bool WCSONLY = false>
__global__ void __launch_bounds__(256, (WCSONLY || (SPH == 0 && !AFF)) ? 6 : 1)
k_resample_tex4(const __grid_constant__ ResampleParams P, cudaTextureObject_t tex)
{
    resample_tex4_body<A, AFF, SPH, WCSONLY, 1>(P, [tex](int) { return tex; }, 1);
}

constexpr int TEX_PLANES_PER_CTA = 8;
template <class A, bool AFF, int SPH, bool WCSONLY = false>
__global__ void __launch_bounds__(256, (WCSONLY || (SPH == 0 && !AFF)) ? 6 : 1)
k_resample_tex4_batch(const __grid_constant__ ResampleParams P, const __grid_constant__ TexBatch TB)
{
    resample_tex4_body<A, AFF, SPH, WCSONLY, TEX_PLANES_PER_CTA>(P, [&TB](int z) { return TB.tex[z]; }, TB.n);
}

// Texture objects over caller memory are cached per (device, pointer, geometry): creating one costs
// microseconds, a resample launch ~100.  A texture object refers to an address range, not to its
// contents, so reuse after the caller rewrote (or re-allocated) the buffer is safe.
//
// This cache and the tensor-map cache below are the library's only process-wide state besides the launch
// counter (include/tfm_b200.h says so).  Both are mutex-protected, LRU, and never block a stream: every
// launch that reads through a texture records an event behind itself on its stream; an evicted object
// whose last reader has not finished is parked and destroyed by a later call once that event has
// completed (round 1 drained the whole device with cudaDeviceSynchronize() instead).
struct TexKey {
    int dev; const void *ptr; int w, h; long long pitch;
    bool operator==(const TexKey &o) const { return dev == o.dev && ptr == o.ptr && w == o.w && h == o.h && pitch == o.pitch; }
};
struct TexEntry {
    TexKey key;
    cudaTextureObject_t obj = 0;
    cudaEvent_t last_use = nullptr;          // recorded behind the latest launch that reads through obj
    bool used = false;
    unsigned long long stamp = 0;            // LRU clock
};
constexpr int TEX_CACHE = 256;
static std::mutex g_tex_mu;
static TexEntry g_tex[TEX_CACHE];
static int g_tex_n = 0;
static unsigned long long g_tex_clock = 0;
struct TexParked { cudaTextureObject_t obj; cudaEvent_t ev; int dev; };
static TexParked g_tex_parked[TEX_CACHE];
static int g_tex_nparked = 0;

static void tex_reap_parked(int dev)         // destroy parked objects whose last reader has finished
{
    int k = 0;
    for (int i = 0; i < g_tex_nparked; ++i) {
        TexParked &p = g_tex_parked[i];
        if (p.dev == dev && cudaEventQuery(p.ev) == cudaSuccess) {
            cudaDestroyTextureObject(p.obj);
            cudaEventDestroy(p.ev);
        } else {
            g_tex_parked[k++] = p;
        }
    }
    cudaGetLastError();                      // cudaErrorNotReady from the queries is not an error
    g_tex_nparked = k;
}

bool get_texture(const void *ptr, int w, int h, long long pitch, cudaTextureObject_t *out, int *slot_out)
{
    int dev = 0;
    cudaGetDevice(&dev);
    TexKey key{dev, ptr, w, h, pitch};
    std::lock_guard<std::mutex> lock(g_tex_mu);
    for (int i = 0; i < g_tex_n; ++i)
        if (g_tex[i].key == key) {
            g_tex[i].stamp = ++g_tex_clock;
            *out = g_tex[i].obj;
            if (slot_out) *slot_out = i;
            return true;
        }
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
    if (g_tex_nparked) tex_reap_parked(dev);
    int slot;
    if (g_tex_n < TEX_CACHE) slot = g_tex_n++;
    else {
        // evict the least recently used entry; a kernel on some stream may still be reading through it
        slot = 0;
        for (int i = 1; i < TEX_CACHE; ++i)
            if (g_tex[i].stamp < g_tex[slot].stamp) slot = i;
        TexEntry &v = g_tex[slot];
        const bool busy = v.used && v.last_use && cudaEventQuery(v.last_use) != cudaSuccess;
        cudaGetLastError();
        if (busy && g_tex_nparked < TEX_CACHE) {
            g_tex_parked[g_tex_nparked++] = TexParked{v.obj, v.last_use, v.key.dev};
            v.last_use = nullptr;            // the parked record owns the event now
        } else {
            if (busy) cudaEventSynchronize(v.last_use);       // parking lot full: wait for this ONE reader
            cudaDestroyTextureObject(v.obj);
        }
        v.used = false;
    }
    g_tex[slot].key = key;
    g_tex[slot].obj = t;
    g_tex[slot].stamp = ++g_tex_clock;
    *out = t;
    if (slot_out) *slot_out = slot;
    return true;
}

// call right after a launch on `st` that reads through the texture of `slot`
void texture_used(int slot, cudaStream_t st)
{
    if (slot < 0) return;
    std::lock_guard<std::mutex> lock(g_tex_mu);
    TexEntry &e = g_tex[slot];
    if (!e.last_use && cudaEventCreateWithFlags(&e.last_use, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        e.last_use = nullptr;
        return;
    }
    cudaEventRecord(e.last_use, st);
    e.used = true;
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
    const int bx0 = ((int)floor(xlo) - 1) & ~3, by0 = (int)floor(ylo) - 1;     // |coordinate| < 2.5e8: host-proven

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
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
            int ix, iy;
            const double Yj = Yd + (double)(8 * j);
            floor_frac_fast(fma(P.aff[0], Xd, fma(P.aff[1], Yj, P.aff[4])), ix, fx[j]);
            floor_frac_fast(fma(P.aff[2], Xd, fma(P.aff[3], Yj, P.aff[5])), iy, fy[j]);
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
// persistent, multi-stage TMA pipeline (fp32 fast path, linear, affine map, truncate / fill 0)
// ---------------------------------------------------------------------------------------------
// The texture-gather kernels stop at the texture unit's write-back (88 % busy at 0.135 ms for a pure
// shift, profiles/r2_kernels.md): every source pixel crosses it four times.  Here HBM -> shared memory is one
// bulk tensor copy per tile -- the source bounding box of a 32 x 32 output tile under an affine map has
// a size the host knows -- and the four taps are LDS.  What made the one-box-per-CTA form of round 1
// slow (exposed TMA latency, per-CTA set-up, mbarrier polling) is removed by making the CTAs persistent:
//   * grid = resident CTAs only; CTA b walks tiles b, b + G, b + 2G, ...
//   * STAGES boxes in flight per CTA: while tile i is being interpolated, the boxes of tiles i+1 ..
//     i+STAGES-1 are already on their way (one elected thread issues cp.async.bulk.tensor.2d and arms
//     the stage's mbarrier with the byte count; everybody waits on the barrier's phase parity);
//   * one __syncthreads per tile (1024 pixels) releases the stage for the next box;
//   * boxes that miss the image entirely are never fetched: the producer arrives on the barrier without
//     a transaction and flags the stage, the consumers store zeros.
// Out-of-range parts of a box are zero-filled by the TMA unit = the reference's truncate boundary with
// fill 0.  Coordinates, floor/fraction and interpolation order are those of every other fast-mode
// kernel (bit-identical results, tests/test_gpu_headline.py).
struct PipeParams {
    int bw, bh;                              // box columns / rows (bw * 4 bytes is a multiple of 16)
    int stage_bytes;                         // box bytes rounded up to 128
    int tiles_x, tiles_y;
    int step_x, step_y;                      // gridDim.x % tiles_x, gridDim.x / tiles_x: the tile walk without divisions
};

constexpr int PIPE_STAGES = 4;
constexpr int PIPE_TW = 64, PIPE_TH = 32;   // output tile: 8 consumer warps x 4 rows, 8 pixels per thread
constexpr int PIPE_THREADS = 288;           // 8 consumer warps + 1 producer warp

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

// out-of-line: the 8 pixels of one consumer thread straight from global memory (never taken unless the
// box arithmetic of host and device disagree)
template <int CUBIC>
static __device__ __noinline__ void pipe_redo_row(const ResampleParams &P, double Xd, double rx, double ry, float *drow, int X)
{
    for (int k = 0; k < 8; ++k) {
        if (X + 8 * k >= P.dst_w) break;
        const double Xj = Xd + (double)(8 * k);
        int ix, iy;
        float fx, fy;
        floor_frac_fast(fma(P.aff[0], Xj, rx), ix, fx);
        floor_frac_fast(fma(P.aff[2], Xj, ry), iy, fy);
        if (CUBIC) {
            float rows[4];
            for (int t = 0; t < 4; ++t)
                rows[t] = hermite_fast(tap_trunc<float, float>(P, P.src, ix - 1, iy - 1 + t), tap_trunc<float, float>(P, P.src, ix, iy - 1 + t),
                                       tap_trunc<float, float>(P, P.src, ix + 1, iy - 1 + t), tap_trunc<float, float>(P, P.src, ix + 2, iy - 1 + t), fx);
            drow[8 * k] = hermite_fast(rows[0], rows[1], rows[2], rows[3], fy);
            continue;
        }
        const float v00 = tap_trunc<float, float>(P, P.src, ix, iy), v01 = tap_trunc<float, float>(P, P.src, ix + 1, iy);
        const float v10 = tap_trunc<float, float>(P, P.src, ix, iy + 1), v11 = tap_trunc<float, float>(P, P.src, ix + 1, iy + 1);
        const float top = fmaf(fx, v01 - v00, v00);
        const float bot = fmaf(fx, v11 - v10, v10);
        drow[8 * k] = fmaf(fy, bot - top, top);
    }
}

__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(256);
    }
}

// CUBIC = 1: the 4 x 4 Hermite taps of interpND's cubic method (helpers.pyx:741-789, fast-mode arithmetic of
// k_resample: hermite_fast along x, then along y) from the same staged boxes, one texel wider on every side.
template <int CUBIC = 0>
__global__ void __launch_bounds__(PIPE_THREADS)
k_resample_pipe(const __grid_constant__ ResampleParams P, const __grid_constant__ CUtensorMap tmap,
                const __grid_constant__ PipeParams B)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // [STAGES boxes][STAGES x int4 box info][STAGES full barriers][STAGES empty barriers]
    int4 *const s_info = reinterpret_cast<int4 *>(smem_raw + (size_t)PIPE_STAGES * B.stage_bytes);   // (bx0, by0, empty, -)
    unsigned long long *const s_full = reinterpret_cast<unsigned long long *>(s_info + PIPE_STAGES);
    unsigned long long *const s_empty = s_full + PIPE_STAGES;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = gridDim.x;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < PIPE_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&s_full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(smem_addr(&s_empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();                                   // barrier initialisation visible to all

    // tile walk of this CTA: t = blockIdx.x, + G, + 2G, ... as (tx, ty) without divisions in the loop
    int tx = blockIdx.x % B.tiles_x, ty = blockIdx.x / B.tiles_x;

    if (warp == 8) {
        // ------------------------------ producer warp (one lane) ------------------------------
        if (lane != 0) return;
        for (int it = 0; ty < B.tiles_y; ++it) {
            const int s = it % PIPE_STAGES;
            // the producer runs ahead and then waits most of the time: back off between polls instead of
            // spinning (ncu: its try_wait loop was 12 % of all issued instructions)
            if (it >= PIPE_STAGES) mbar_wait_sleep(smem_addr(&s_empty[s]), (uint32_t)(((it / PIPE_STAGES) - 1) & 1));
            // source box of the tile: the map is affine, so the extremes are at the tile corners; one pixel
            // of margin on the low side, the host-chosen box size covers the high side
            const double TX = (double)(P.dst_x0 + (long long)tx * PIPE_TW), TY = (double)(P.dst_y0 + (long long)ty * PIPE_TH);
            const double cx0 = fma(P.aff[0], TX, fma(P.aff[1], TY, P.aff[4]));
            const double cy0 = fma(P.aff[2], TX, fma(P.aff[3], TY, P.aff[5]));
            const double xlo = cx0 + fmin((PIPE_TW - 1.0) * P.aff[0], 0.0) + fmin((PIPE_TH - 1.0) * P.aff[1], 0.0);
            const double ylo = cy0 + fmin((PIPE_TW - 1.0) * P.aff[2], 0.0) + fmin((PIPE_TH - 1.0) * P.aff[3], 0.0);
            // the box must start on a 16-byte boundary of the innermost axis: round down to a multiple of 4
            const int bx0 = ((int)floor(xlo) - 1 - CUBIC) & ~3, by0 = (int)floor(ylo) - 1 - CUBIC;   // |coordinate| < 2.5e8: host-proven
            // box position inside the staged rectangle (the whole image: src_x0 = src_y0 = 0)
            const int tx0 = bx0 - P.src_x0, ty0 = by0 - P.src_y0;
            const bool empty = tx0 >= P.src_w || ty0 >= P.src_h || tx0 + B.bw <= 0 || ty0 + B.bh <= 0;
            s_info[s] = make_int4(bx0, by0, empty ? 1 : 0, 0);
            const uint32_t bar = smem_addr(&s_full[s]);
            if (empty) {
                // nothing to fetch: the consumers store zeros
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
            } else {
                const uint32_t bytes = (uint32_t)(B.bw * B.bh * 4);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                    ::"r"(smem_addr(smem_raw + (size_t)s * B.stage_bytes)), "l"(reinterpret_cast<uint64_t>(&tmap)),
                      "r"(tx0), "r"(ty0), "r"(bar)
                    : "memory");
            }
            tx += B.step_x; ty += B.step_y;
            if (tx >= B.tiles_x) { tx -= B.tiles_x; ++ty; }
        }
        return;
    }

    // ---------------------------------- consumer warps ----------------------------------
    const int px = lane & 7;
    const int py = (warp << 2) + (lane >> 3);
    const int bw = B.bw;
    for (int it = 0; ty < B.tiles_y; ++it) {
        const int s = it % PIPE_STAGES;
        const int X = tx * PIPE_TW + px, Y = ty * PIPE_TH + py;
        // row part of the coordinates while the box may still be in flight (canonical form of every
        // fast-mode kernel: c = fma(a0, X, fma(a1, Y, b)) on exact doubles)
        const double Xd = (double)(P.dst_x0 + X), Yd = (double)(P.dst_y0 + Y);
        const double rx = fma(P.aff[1], Yd, P.aff[4]), ry = fma(P.aff[3], Yd, P.aff[5]);
        mbar_wait(smem_addr(&s_full[s]), (uint32_t)((it / PIPE_STAGES) & 1));
        const int4 info = s_info[s];
        const float *const tile = reinterpret_cast<const float *>(smem_raw + (size_t)s * B.stage_bytes);
        float *const drow = reinterpret_cast<float *>(static_cast<char *>(P.dst) + (long long)Y * P.dst_pitch) + X;
        const bool row_in = Y < P.dst_h;
        const bool full = row_in && (tx * PIPE_TW + PIPE_TW <= P.dst_w);
        // local tap index = floor(c) - box origin; the bias of the raw floor is folded into the origin
        const unsigned ox = (unsigned)info.x - TFM_FLOOR_RAW_BIAS, oy = (unsigned)info.y - TFM_FLOOR_RAW_BIAS;
        bool bad = false;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float r[4];
            if (info.z) {
                r[0] = r[1] = r[2] = r[3] = 0.0f;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double Xj = Xd + (double)(8 * (4 * h + j));
                    unsigned ixr, iyr;
                    float fx, fy;
                    floor_frac_raw(fma(P.aff[0], Xj, rx), ixr, fx);
                    floor_frac_raw(fma(P.aff[2], Xj, ry), iyr, fy);
                    const unsigned lx = ixr - ox, ly = iyr - oy;
                    // the host sized the box so that every tap of the tile lies inside it; should rounding ever
                    // disagree, the tap index is neutralised here and the thread redoes its pixels from global
                    // memory below
                    if (CUBIC) {
                        // taps lx - 1 .. lx + 2, ly - 1 .. ly + 2 must lie inside the box
                        const bool out = (lx - 1u) >= (unsigned)(bw - 3) || (ly - 1u) >= (unsigned)(B.bh - 3);
                        bad = bad || out;
                        const float *tp = tile + (out ? 0 : (int)((ly - 1u) * (unsigned)bw + (lx - 1u)));
                        float rows[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            rows[t] = hermite_fast(tp[t * bw], tp[t * bw + 1], tp[t * bw + 2], tp[t * bw + 3], fx);
                        r[j] = hermite_fast(rows[0], rows[1], rows[2], rows[3], fy);
                        continue;
                    }
                    const bool out = lx >= (unsigned)(bw - 1) || ly >= (unsigned)(B.bh - 1);
                    bad = bad || out;
                    const float *tp = tile + (out ? 0 : (int)(ly * (unsigned)bw + lx));
                    const float v00 = tp[0], v01 = tp[1], v10 = tp[bw], v11 = tp[bw + 1];
                    const float top = fmaf(fx, v01 - v00, v00);
                    const float bot = fmaf(fx, v11 - v10, v10);
                    r[j] = fmaf(fy, bot - top, top);
                }
            }
            float *const d = drow + 32 * h;
            if (full) { d[0] = r[0]; d[8] = r[1]; d[16] = r[2]; d[24] = r[3]; }
            else if (row_in) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (X + 32 * h + 8 * j < P.dst_w) d[8 * j] = r[j];
            }
        }
        if (bad && row_in) pipe_redo_row<CUBIC>(P, Xd, rx, ry, drow, X);
        // this warp is done with stage s: one arrival per consumer warp releases it to the producer
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&s_empty[s])) : "memory");
        tx += B.step_x; ty += B.step_y;
        if (tx >= B.tiles_x) { tx -= B.tiles_x; ++ty; }
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

// Fraction of a 16 x 16 lattice of output grid points that an affine map sends inside the source.
static double affine_coverage(const double *aff, const tfm_image &d, const tfm_image &s)
{
    int inside = 0;
    for (int j = 0; j < 16; ++j)
        for (int i = 0; i < 16; ++i) {
            const double X = (double)d.x0 + (i + 0.5) * (double)d.w / 16.0, Y = (double)d.y0 + (j + 0.5) * (double)d.h / 16.0;
            const double x = aff[0] * X + aff[1] * Y + aff[4] - (double)s.x0, y = aff[2] * X + aff[3] * Y + aff[5] - (double)s.y0;
            inside += (x >= -1.0 && x <= (double)s.w && y >= -1.0 && y <= (double)s.h) ? 1 : 0;
        }
    return inside / 256.0;
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
    bool generic = subrect || a->bound_x != TFM_B_TRUNCATE || a->bound_y != TFM_B_TRUNCATE ||
                   src_elems >= 0x7fffffffLL;            // the fast path indexes taps with 32-bit offsets
    bool staged_fast = false;
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
        P.aff_bounded = (bx < TFM_SANE_COORD && by < TFM_SANE_COORD) ? 1 : 0;      // NaN parameters compare false
        for (int k = 0; k < 6; ++k) P.aff_f[k] = (float)P.aff[k];
        P.aff_f[6] = (float)d.x0;
        P.aff_f[7] = (float)d.y0;
        P.aff_margin = (float)(4e-6 * (bx > by ? bx : by));      // a few fp32 ulps of the largest term
        if (!(P.aff_margin < 64.0f)) P.aff_bounded = 0;          // huge coordinates: skip the shortcut
        P.aff_c[0] = (float)(P.aff[4] + 15.5 * (P.aff[0] + P.aff[1]));
        P.aff_c[1] = (float)(P.aff[5] + 15.5 * (P.aff[2] + P.aff[3]));
        P.aff_c[2] = (float)(15.5 * (fabs(P.aff[0]) + fabs(P.aff[1])) + 3.0) + P.aff_margin;
        P.aff_c[3] = (float)(15.5 * (fabs(P.aff[2]) + fabs(P.aff[3])) + 3.0) + P.aff_margin;
        P.aff_lim[0] = (float)s.w + P.aff_c[2];
        P.aff_lim[1] = (float)s.h + P.aff_c[3];
    }
    // A staged source rectangle (multi-GPU shard) under an affine map with truncate boundaries: when the host
    // can PROVE that every in-image tap of this output rectangle lies inside the staged rectangle (an affine
    // map takes its extremes at the corners), the texture-gather / TMA-pipeline kernels run on the staged
    // rectangle as if it were the image -- same coordinates, tap indices shifted by the staged origin -- and
    // what they read as 0 beyond it is beyond the image.  Anything unproven stays with the generic kernel,
    // which checks every tap and raises the status word.
    if (subrect && aff && P.aff_bounded && a->method == TFM_LINEAR && a->bound_x == TFM_B_TRUNCATE &&
        a->bound_y == TFM_B_TRUNCATE && a->fillvalue == 0.0 && src_elems < 0x7fffffffLL && (s.x0 % 4) == 0) {
        double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
        for (int k = 0; k < 4; ++k) {
            const double X = (double)d.x0 + ((k & 1) ? (double)(d.w - 1) : 0.0), Y = (double)d.y0 + ((k & 2) ? (double)(d.h - 1) : 0.0);
            const double x = P.aff[0] * X + P.aff[1] * Y + P.aff[4], y = P.aff[2] * X + P.aff[3] * Y + P.aff[5];
            xmin = fmin(xmin, x); xmax = fmax(xmax, x); ymin = fmin(ymin, y); ymax = fmax(ymax, y);
        }
        // taps floor(c) .. floor(c) + 1, one more on each side for the 2^-23 quantisation of the floor
        const double cx0 = fmax(floor(xmin) - 1.0, 0.0), cx1 = fmin(floor(xmax) + 2.0, (double)s.full_w - 1.0);
        const double cy0 = fmax(floor(ymin) - 1.0, 0.0), cy1 = fmin(floor(ymax) + 2.0, (double)s.full_h - 1.0);
        const bool none = cx1 < cx0 || cy1 < cy0;                    // the rectangle maps wholly outside the image
        staged_fast = none || (cx0 >= (double)s.x0 && cx1 <= (double)(s.x0 + s.w - 1) &&
                               cy0 >= (double)s.y0 && cy1 <= (double)(s.y0 + s.h - 1));
        if (staged_fast) {
            generic = false;
            // the empty-tile tests work in staged-local coordinates
            P.aff_c[0] -= (float)s.x0; P.aff_c[1] -= (float)s.y0;
        }
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
        // (a batch of planes runs the texture kernel once per plane, each plane through its own object)
        cudaTextureObject_t tex = 0;
        int tex_slot = -1;
        const bool multi = s.n_planes > 1;
        // the TMA pipeline also serves the cubic method (4 x 4 taps from the same staged boxes); it needs no texture
        const bool cubic_pipe = !generic && a->method == TFM_CUBIC && s.dtype == TFM_F32 && d.dtype == TFM_F32 &&
                                a->fillvalue == 0.0 && s.n_planes == 1 && !(a->tuning & (16 | 128 | 256));
        const bool tex_ok = !generic && a->method == TFM_LINEAR && s.dtype == TFM_F32 && d.dtype == TFM_F32 &&
                            a->fillvalue == 0.0 && !(a->tuning & 16) &&
                            (!multi || ((s.plane_pitch % 512) == 0 && s.n_planes <= 192 && !(a->tuning & (128 | 256 | 8192)))) &&
                            s.w <= 131072 && s.h <= 65000 && (s.pitch % 32) == 0 &&
                            (reinterpret_cast<uintptr_t>(s.data) % 512) == 0 &&
                            get_texture(s.data, (int)s.w, (int)s.h, s.pitch, &tex, &tex_slot);
        // TMA-staged variant: same conditions, host-composed affine map with a bounded source box
        bool tma_done = false;
        // Opt-in (tuning bit 256): measured SLOWER than the texture gather on B200 (8192^2: named chain
        // 0.138 vs 0.090 ms, pure shift 0.227 vs 0.141 ms) -- per-CTA load latency is exposed in this
        // one-box-per-CTA form and the shared-memory taps cost more issue slots than one TLD4.
        if (tex_ok && !multi && aff && P.aff_bounded && (a->tuning & 256) && (s.pitch % 16) == 0 && !staged_fast) {
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
        // persistent multi-stage TMA pipeline: same conditions
        if (!tma_done && (tex_ok || cubic_pipe) && !multi && aff && P.aff_bounded && (s.pitch % 16) == 0 && !(a->tuning & (128 | 16 | 256)) &&
            ((a->tuning & 8192) || cubic_pipe || affine_coverage(P.aff, d, s) >= TFM_PIPE_COVERAGE)) {
            // (cubic: the pipeline wins whatever the coverage -- named chain, 42 % covered: 0.28 vs 0.33 ms)
            const int cub = cubic_pipe ? 1 : 0;
            const double ex = (PIPE_TW - 1) * fabs(P.aff[0]) + (PIPE_TH - 1) * fabs(P.aff[1]);
            const double ey = (PIPE_TW - 1) * fabs(P.aff[2]) + (PIPE_TH - 1) * fabs(P.aff[3]);
            if (ex < 240.0 && ey < 240.0) {
                PipeParams B;
                // +1 low margin, +2 for floor/+1 tap, +2 slack (the 2^-23 quantisation of floor_frac_fast can round
                // a coordinate up to the next integer), +3 for the aligned start; multiple of 4
                // cubic: one more texel on every side
                B.bw = (((int)ceil(ex) + 8 + 2 * cub) + 3) & ~3;
                B.bh = (int)ceil(ey) + 5 + 2 * cub;
                B.stage_bytes = (B.bw * B.bh * 4 + 127) & ~127;
                B.tiles_x = (P.dst_w + PIPE_TW - 1) / PIPE_TW;
                B.tiles_y = (P.dst_h + PIPE_TH - 1) / PIPE_TH;
                const size_t smem = (size_t)PIPE_STAGES * B.stage_bytes + PIPE_STAGES * (16 + 8 + 8) + 16;
                CUtensorMap tm;
                if (B.bw <= 256 && B.bh <= 256 && smem <= 200 * 1024 &&
                    (long long)B.tiles_x * B.tiles_y < 0x3fffffffLL &&
                    get_tensor_map(s.data, (int)s.w, (int)s.h, s.pitch, B.bw, B.bh, &tm)) {
                    static int sm_count = 0;
                    if (!sm_count) {
                        int dev = 0;
                        cudaGetDevice(&dev);
                        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
                    }
                    cudaError_t ea = cudaSuccess;
                    const void *kfn = cub ? (const void *)k_resample_pipe<1> : (const void *)k_resample_pipe<0>;
                    if (smem > 48 * 1024)
                        ea = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    int per_sm = 0;
                    if (ea == cudaSuccess)
                        ea = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, PIPE_THREADS, smem);
                    if (ea == cudaSuccess && per_sm > 0) {
                        long long g = (long long)sm_count * per_sm;
                        const long long nt = (long long)B.tiles_x * B.tiles_y;
                        if (g > nt) g = nt;
                        B.step_x = (int)(g % B.tiles_x);
                        B.step_y = (int)(g / B.tiles_x);
                        if (cub) k_resample_pipe<1><<<(unsigned)g, PIPE_THREADS, smem, st>>>(P, tm, B);
                        else k_resample_pipe<0><<<(unsigned)g, PIPE_THREADS, smem, st>>>(P, tm, B);
                        count_launch();
                        name = cub ? "k_resample_pipe<affine,cubic>" : "k_resample_pipe<affine>";
                        e = cudaGetLastError();
                        tma_done = true;
                    } else {
                        cudaGetLastError();
                    }
                }
            }
        }
        if (tma_done) {
        } else if (tex_ok) {
            dim3 grid((unsigned)((P.dst_w + 31) / 32), (unsigned)((P.dst_h + 31) / 32), 1);
            if (grid.y > 65535u) return TFM_EUNSUPPORTED;
            const int sph = chain_sph_level(P.chain);
            if ((a->tuning & 128) && !staged_fast) { // first form (4 pixels 8 rows apart, scalar stores): A/B runs only
                if (aff) k_resample_tex<Fast, true, 0><<<grid, 256, 0, st>>>(P, tex);
                else if (sph == 2) k_resample_tex<Fast, false, 2><<<grid, 256, 0, st>>>(P, tex);
                else if (sph == 1) k_resample_tex<Fast, false, 1><<<grid, 256, 0, st>>>(P, tex);
                else k_resample_tex<Fast, false, 0><<<grid, 256, 0, st>>>(P, tex);
                name = aff ? "k_resample_tex<affine>" : "k_resample_tex";
            } else {
                const bool aff4 = aff && P.aff_bounded;
                const bool wcs_only = !aff4 && P.chain.n == 1 && P.chain.ops[0].kind == TFM_OP_WCS_PAIR &&
                                      (P.chain.ops[0].flags & 0xffff) == (TFM_WCS_PROJ_CAR | (TFM_WCS_PROJ_TAN << 8));
                e = cudaSuccess;
                if (s.n_planes == 1 || (a->tuning & 1048576)) {    // bit 1048576: one launch per plane (A/B runs)
                    for (long long k = 0; k < s.n_planes && e == cudaSuccess; ++k) {
                        if (k > 0) {
                            P.src = static_cast<const char *>(s.data) + k * s.plane_pitch;
                            P.dst = static_cast<char *>(d.data) + k * d.plane_pitch;
                            if (!get_texture(P.src, (int)s.w, (int)s.h, s.pitch, &tex, &tex_slot)) { e = cudaErrorUnknown; break; }
                        }
                        if (aff4) k_resample_tex4<Fast, true, 0><<<grid, 256, 0, st>>>(P, tex);
                        else if (wcs_only) k_resample_tex4<Fast, false, 1, true><<<grid, 256, 0, st>>>(P, tex);
                        else if (sph == 2) k_resample_tex4<Fast, false, 2><<<grid, 256, 0, st>>>(P, tex);
                        else if (sph == 1) k_resample_tex4<Fast, false, 1><<<grid, 256, 0, st>>>(P, tex);
                        else k_resample_tex4<Fast, false, 0><<<grid, 256, 0, st>>>(P, tex);
                        count_launch();
                        texture_used(tex_slot, st);
                        tex_slot = -1;
                        e = cudaGetLastError();
                    }
                    name = aff4 ? "k_resample_tex4<affine>" : (wcs_only ? "k_resample_tex4<car-tan>" : "k_resample_tex4");
                } else {
                    // a batch of planes: up to TEX_BATCH_MAX texture objects per launch, a CTA computes the
                    // coordinates of its tile once for TEX_PLANES_PER_CTA planes
                    for (long long k0 = 0; k0 < s.n_planes && e == cudaSuccess; k0 += TEX_BATCH_MAX) {
                        TexBatch TB;
                        int slots[TEX_BATCH_MAX];
                        TB.n = (int)std::min<long long>(TEX_BATCH_MAX, s.n_planes - k0);
                        for (int z = 0; z < TEX_BATCH_MAX; ++z) { TB.tex[z] = 0; slots[z] = -1; }
                        for (int z = 0; z < TB.n; ++z) {
                            if (k0 + z == 0) { TB.tex[0] = tex; slots[0] = tex_slot; tex_slot = -1; continue; }
                            const char *pz = static_cast<const char *>(s.data) + (k0 + z) * s.plane_pitch;
                            if (!get_texture(pz, (int)s.w, (int)s.h, s.pitch, &TB.tex[z], &slots[z])) { e = cudaErrorUnknown; break; }
                        }
                        if (e != cudaSuccess) break;
                        P.src = static_cast<const char *>(s.data) + k0 * s.plane_pitch;
                        P.dst = static_cast<char *>(d.data) + k0 * d.plane_pitch;
                        dim3 bgrid(grid.x, grid.y, (unsigned)((TB.n + TEX_PLANES_PER_CTA - 1) / TEX_PLANES_PER_CTA));
                        if (aff4) k_resample_tex4_batch<Fast, true, 0><<<bgrid, 256, 0, st>>>(P, TB);
                        else if (wcs_only) k_resample_tex4_batch<Fast, false, 1, true><<<bgrid, 256, 0, st>>>(P, TB);
                        else if (sph == 2) k_resample_tex4_batch<Fast, false, 2><<<bgrid, 256, 0, st>>>(P, TB);
                        else if (sph == 1) k_resample_tex4_batch<Fast, false, 1><<<bgrid, 256, 0, st>>>(P, TB);
                        else k_resample_tex4_batch<Fast, false, 0><<<bgrid, 256, 0, st>>>(P, TB);
                        count_launch();
                        for (int z = 0; z < TB.n; ++z) texture_used(slots[z], st);
                        e = cudaGetLastError();
                    }
                    name = aff4 ? "k_resample_tex4_batch<affine>" : (wcs_only ? "k_resample_tex4_batch<car-tan>" : "k_resample_tex4_batch");
                }
                set_last_kernel(name);
                return e == cudaSuccess ? TFM_OK : TFM_ECUDA;
            }
            count_launch();
            texture_used(tex_slot, st);
            e = cudaGetLastError();
        } else if (s.dtype == TFM_F32 && d.dtype == TFM_F32) {
            if (staged_fast) { P.aff_c[0] += (float)s.x0; P.aff_c[1] += (float)s.y0; }
            e = launch1<float, float, Fast, float>(a->method, P, generic || staged_fast, aff, st, &name);
        } else
            return TFM_EUNSUPPORTED;
    } else {
        return TFM_EVALUE;
    }
    set_last_kernel(name);
    return e == cudaSuccess ? TFM_OK : TFM_ECUDA;
}

}  // namespace tfm
