// tfm_fits.cu -- K0: FITS pixel decode on the device (the step before
// the path, file -> device image).
//
// A FITS data unit is big-endian, BITPIX in {8, 16, 32, 64, -32, -64}, optionally scaled
// (physical = BZERO + BSCALE * stored).  The host reader used to do this with NumPy
// (`frombuffer(...).astype(native)`, `data * bscale + bzero`, fitswcs.read_fits); here the raw bytes
// are uploaded as they are and one streaming kernel byte-swaps, converts and scales.  Pure HBM
// traffic: bytes in + bytes out, every byte touched once, 16-byte packets per thread, persistent
// grid-stride loop.  Bit-exact with the NumPy path (integer -> float64 conversions are exact for
// |v| < 2^53; the scaling is one rounded multiply and one rounded add, as NumPy does it).
#include <cuda_runtime.h>
#include <stdint.h>
#include "tfm_common.cuh"

namespace tfm {

__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

struct FitsParams {
    const unsigned char *raw;
    void *out;
    long long n;                             // pixels
    double bscale, bzero;
    int scaled;
};

// physical value of one stored sample, NumPy's `data * bscale + bzero` (two roundings)
__device__ __forceinline__ double fits_scale(const FitsParams &P, double v)
{
    return P.scaled ? __dadd_rn(__dmul_rn(v, P.bscale), P.bzero) : v;
}

// BITPIX = -32: float32 in, float32 out; 4 pixels (16 bytes) per packet.  Scaling stays in float32,
// as NumPy does for a float32 array and Python-scalar BSCALE / BZERO.
__device__ __forceinline__ uint32_t fits_f32_word(const FitsParams &P, uint32_t raw)
{
    const uint32_t u = bswap32(raw);
    if (!P.scaled) return u;
    return __float_as_uint(__fadd_rn(__fmul_rn(__uint_as_float(u), (float)P.bscale), (float)P.bzero));
}

__global__ void __launch_bounds__(256) k_fits_f32(const __grid_constant__ FitsParams P)
{
    const long long npk = P.n >> 2;
    const uint4 *in = reinterpret_cast<const uint4 *>(P.raw);
    uint4 *out = reinterpret_cast<uint4 *>(P.out);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npk; i += stride) {
        uint4 v = __ldcs(in + i);
        v.x = fits_f32_word(P, v.x); v.y = fits_f32_word(P, v.y);
        v.z = fits_f32_word(P, v.z); v.w = fits_f32_word(P, v.w);
        __stcs(out + i, v);
    }
    if (blockIdx.x == 0 && threadIdx.x < (P.n & 3)) {                 // ragged tail
        const long long k = (npk << 2) + threadIdx.x;
        reinterpret_cast<uint32_t *>(P.out)[k] = fits_f32_word(P, reinterpret_cast<const uint32_t *>(P.raw)[k]);
    }
}

// every other combination: float64 out.  BYTES = bytes per stored sample; FLT: stored type is IEEE
template <int BYTES, bool FLT>
__device__ __forceinline__ double fits_sample(const unsigned char *p)
{
    if (BYTES == 1) return (double)p[0];                              // BITPIX 8 is unsigned
    if (BYTES == 2) return (double)(short)((p[0] << 8) | p[1]);
    if (BYTES == 4) {
        const uint32_t u = bswap32(*reinterpret_cast<const uint32_t *>(p));
        return FLT ? (double)__uint_as_float(u) : (double)(int)u;
    }
    const uint32_t hi = bswap32(reinterpret_cast<const uint32_t *>(p)[0]);
    const uint32_t lo = bswap32(reinterpret_cast<const uint32_t *>(p)[1]);
    const unsigned long long u = ((unsigned long long)hi << 32) | lo;
    return FLT ? __longlong_as_double((long long)u) : (double)(long long)u;
}

// One 16-byte OUTPUT packet (two float64) per thread and iteration: the stores of a warp are one
// contiguous 512-byte run and its loads one contiguous 2*BYTES*32-byte run, whatever BITPIX is
// (16 bytes of INPUT per thread made the 8x-expanding BITPIX = 16 case store with a 64-byte lane
// stride: measured 43 % of the HBM roofline against 75-79 % for the others).
template <int BYTES> struct FitsPair;
template <> struct FitsPair<1> { typedef unsigned short type; };
template <> struct FitsPair<2> { typedef unsigned int type; };
template <> struct FitsPair<4> { typedef uint2 type; };
template <> struct FitsPair<8> { typedef uint4 type; };

template <int BYTES, bool FLT>
__global__ void __launch_bounds__(256) k_fits_f64(const __grid_constant__ FitsParams P)
{
    typedef typename FitsPair<BYTES>::type Pair;
    const long long npk = P.n >> 1;
    const Pair *in = reinterpret_cast<const Pair *>(P.raw);
    double2 *out = static_cast<double2 *>(P.out);
    const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll 4
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npk; i += stride) {
        alignas(16) Pair pk = __ldcs(in + i);
        const unsigned char *b = reinterpret_cast<const unsigned char *>(&pk);
        const double v0 = fits_scale(P, fits_sample<BYTES, FLT>(b));
        const double v1 = fits_scale(P, fits_sample<BYTES, FLT>(b + BYTES));
        __stcs(out + i, make_double2(v0, v1));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && (P.n & 1)) {           // odd last sample
        const long long k = P.n - 1;
        alignas(8) unsigned char al[8];               // fits_sample reads 4- and 8-byte words
        for (int q = 0; q < BYTES; ++q) al[q] = P.raw[k * BYTES + q];
        static_cast<double *>(P.out)[k] = fits_scale(P, fits_sample<BYTES, FLT>(al));
    }
}

int fits_decode_dispatch(const void *raw, int bitpix, long long n, double bscale, double bzero,
                         int out_dtype, void *out, cudaStream_t st)
{
    if (n < 0) return TFM_EVALUE;
    if (n == 0) return TFM_OK;
    if (!raw || !out) return TFM_EVALUE;
    if ((reinterpret_cast<uintptr_t>(raw) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return TFM_EVALUE;
    const bool scaled = !(bscale == 1.0 && bzero == 0.0);
    FitsParams P;
    P.raw = static_cast<const unsigned char *>(raw);
    P.out = out;
    P.n = n;
    P.bscale = bscale;
    P.bzero = bzero;
    P.scaled = scaled ? 1 : 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (n + 1023) / 1024;
    const int grid = (int)(want < (long long)sms * 8 ? (want > 0 ? want : 1) : (long long)sms * 8);
    if (bitpix == -32 && out_dtype == TFM_F32) {
        k_fits_f32<<<grid, 256, 0, st>>>(P);
        set_last_kernel("k_fits_f32");
    } else {
        if (out_dtype != TFM_F64) return TFM_EVALUE;
        switch (bitpix) {
        case 8: k_fits_f64<1, false><<<grid, 256, 0, st>>>(P); break;
        case 16: k_fits_f64<2, false><<<grid, 256, 0, st>>>(P); break;
        case 32: k_fits_f64<4, false><<<grid, 256, 0, st>>>(P); break;
        case 64: k_fits_f64<8, false><<<grid, 256, 0, st>>>(P); break;
        case -32: k_fits_f64<4, true><<<grid, 256, 0, st>>>(P); break;
        case -64: k_fits_f64<8, true><<<grid, 256, 0, st>>>(P); break;
        default: return TFM_EVALUE;
        }
        set_last_kernel("k_fits_f64");
    }
    count_launch();
    return cudaGetLastError() == cudaSuccess ? TFM_OK : TFM_ECUDA;
}

}  // namespace tfm
