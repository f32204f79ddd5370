// tfm_common.cuh -- bookkeeping shared by the translation units of libtfm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tfm_b200.h"

namespace tfm {

void count_launch();                       // bumps the process-wide launch counter (tfm_launch_count)
void set_last_kernel(const char *name);    // thread-local: variant chosen by the last call

int resample_dispatch(const tfm_resample_args *a, cudaStream_t st);
int jacobian_dispatch(const tfm_jacobian_args *a, cudaStream_t st);
int jacobian_fast_dispatch(const tfm_jacobian_args *a, cudaStream_t st, int *rc);   // tfm_jacobian_fast.cu
int apply_dispatch(const tfm_apply_args *a, cudaStream_t st);
int apply_linear_nd_dispatch(const tfm_apply_nd_args *a, cudaStream_t st);
int jacobian_helpers(int which, const void *in0, const void *in1, long long ny, long long nx, double thresh,
                     void *out, void *scratch, cudaStream_t st);
int fits_decode_dispatch(const void *raw, int bitpix, long long n, double bscale, double bzero,
                         int out_dtype, void *out, cudaStream_t st);
int svd_hooks(int which, const double *in0, const double *in1, const double *in2,
              double *out0, double *out1, double *out2);


// point-sampled, border-addressed texture object over a pitch-linear float image (cached; tfm_resample.cu)
bool get_texture(const void *ptr, int w, int h, long long pitch, cudaTextureObject_t *out, int *slot_out = nullptr);
void texture_used(int slot, cudaStream_t st);      // record "this stream reads through the texture of `slot`"

#ifdef __CUDACC__
#define TFM_SANE_COORD 2.5e8   /* |coordinate| bound of the fast-mode floor/fraction (floor_frac_fast) */

// floor + fractional part without the conversion (XU) pipe, in fixed point: t = x + 1.5*2^29 has an
// ulp of 2^-23 for |x| < 2^28, so the low 52 bits of t hold 2^51 + round(x * 2^23).  An arithmetic
// right shift by 23 of that integer IS floor(x) (of the rounded value), and its low 23 bits are the
// fraction -- turned into a float by planting them in the mantissa of 1.0f.  5 ALU/FMA-pipe
// instructions per axis, none on the conversion pipe; the fraction is quantised to 2^-23 (1.2e-7 px,
// far inside the fp32 path's 1e-5 tolerance) and is exactly 0 at integer coordinates, so an identity
// resample returns the image bit for bit.  EVERY fast-mode kernel takes index and fraction from this
// one function: variants (texture, LDG, TMA-staged, generic/sharded) agree bit for bit.
// Valid for |x| < TFM_SANE_COORD; callers route anything else to an out-of-range index first.
__device__ __forceinline__ unsigned lop3_and_or(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));      // (a & b) | c
    return d;
}

// the same, with the index left as (floor(x) - TFM_FLOOR_RAW_BIAS) mod 2^32: a caller that subtracts an origin
// anyway folds the bias into it
#define TFM_FLOOR_RAW_BIAS 0x70000000u
__device__ __forceinline__ void floor_frac_raw(double x, unsigned &i_raw, float &frac)
{
    const double t = x + 805306368.0;
    const unsigned lo = (unsigned)__double2loint(t), hi = (unsigned)__double2hiint(t);
    i_raw = __funnelshift_r(lo, hi, 23);
    frac = __uint_as_float(lop3_and_or(lo, 0x007fffffu, 0x3f800000u)) - 1.0f;
}

__device__ __forceinline__ void floor_frac_fast(double x, int &i, float &frac)
{
    const double t = x + 805306368.0;                          // 1.5 * 2^29 = 0x41C8000000000000
    const unsigned lo = (unsigned)__double2loint(t), hi = (unsigned)__double2hiint(t);
    // (hi:lo - 0x41C80000:0) >> 23, low 32 bits: the constant's contribution is (0x41C80000 << 9) mod 2^32
    i = (int)(__funnelshift_r(lo, hi, 23) + 0x70000000u);
    frac = __uint_as_float(lop3_and_or(lo, 0x007fffffu, 0x3f800000u)) - 1.0f;
}

#endif  // __CUDACC__

}  // namespace tfm
