// tfm_capi.cu -- the extern "C" surface declared in include/tfm_b200.h.
// Plain pointers and sizes only; no torch types, no exceptions across the boundary.
#include <atomic>
#include <cuda_runtime.h>
#include "tfm_common.cuh"

namespace tfm {

static std::atomic<long long> g_launches{0};
static thread_local const char *t_last_kernel = "";

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void set_last_kernel(const char *name) { t_last_kernel = name; }

}  // namespace tfm

extern "C" {

int tfm_abi_version(void) { return TFM_ABI_VERSION; }

const char *tfm_strerror(int status)
{
    switch (status) {
    case TFM_OK: return "ok";
    case TFM_EVALUE: return "invalid argument (ValueError in the reference)";
    case TFM_EDIRECTION: return "transform is invalid in the requested direction (AssertionError in the reference)";
    case TFM_EFORBID: return "boundary violation with 'forbid' condition (ValueError in the reference)";
    case TFM_ECUDA: return "CUDA runtime error";
    case TFM_EUNSUPPORTED: return "combination not supported by libtfm_b200 (no CPU fallback exists)";
    case TFM_ESTAGED: return "a tap fell outside the staged source sub-rectangle";
    default: return "unknown status";
    }
}

int64_t tfm_launch_count(void) { return (int64_t)tfm::g_launches.load(std::memory_order_relaxed); }

const char *tfm_last_kernel(void) { return tfm::t_last_kernel; }

int tfm_resample(const tfm_resample_args *args, void *stream)
{
    return tfm::resample_dispatch(args, static_cast<cudaStream_t>(stream));
}

int tfm_resample_jacobian(const tfm_jacobian_args *args, void *stream)
{
    return tfm::jacobian_dispatch(args, static_cast<cudaStream_t>(stream));
}

int tfm_apply(const tfm_apply_args *args, void *stream)
{
    return tfm::apply_dispatch(args, static_cast<cudaStream_t>(stream));
}

int tfm_apply_linear_nd(const tfm_apply_nd_args *args, void *stream)
{
    return tfm::apply_linear_nd_dispatch(args, static_cast<cudaStream_t>(stream));
}

int tfm_simple_jacobian(const double *index, int64_t ny, int64_t nx, double *Js, void *stream)
{
    return tfm::jacobian_helpers(0, index, nullptr, ny, nx, 0.0, Js, nullptr, static_cast<cudaStream_t>(stream));
}

int tfm_jump_detect(const double *Js, int64_t ncy, int64_t ncx, double jump_thresh, int64_t *flags,
                    double *scratch, void *stream)
{
    return tfm::jacobian_helpers(1, Js, nullptr, ncy, ncx, jump_thresh, flags, scratch,
                                 static_cast<cudaStream_t>(stream));
}

int tfm_jacobian_grid(const double *Js, const int64_t *flags, int64_t ny, int64_t nx, double *J, void *stream)
{
    return tfm::jacobian_helpers(2, Js, flags, ny, nx, 0.0, J, nullptr, static_cast<cudaStream_t>(stream));
}

int tfm_fits_decode(const void *raw, int32_t bitpix, int64_t n, double bscale, double bzero,
                    int32_t out_dtype, void *out, void *stream)
{
    return tfm::fits_decode_dispatch(raw, bitpix, n, bscale, bzero, out_dtype, out,
                                     static_cast<cudaStream_t>(stream));
}

int tfm_svd2x2(const double *M, double *U, double *s, double *V)
{
    if (!M || !U || !s || !V) return TFM_EVALUE;
    return tfm::svd_hooks(0, M, nullptr, nullptr, U, s, V);
}

int tfm_svc2x2(const double *U, const double *s, const double *V, double *M)
{
    if (!M || !U || !s || !V) return TFM_EVALUE;
    return tfm::svd_hooks(1, U, s, V, M, nullptr, nullptr);
}

}  // extern "C"
