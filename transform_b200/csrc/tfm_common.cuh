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
int apply_dispatch(const tfm_apply_args *a, cudaStream_t st);
int jacobian_helpers(int which, const void *in0, const void *in1, long long ny, long long nx, double thresh,
                     void *out, void *scratch, cudaStream_t st);
int fits_decode_dispatch(const void *raw, int bitpix, long long n, double bscale, double bzero,
                         int out_dtype, void *out, cudaStream_t st);
int svd_hooks(int which, const double *in0, const double *in1, const double *in2,
              double *out0, double *out1, double *out2);

}  // namespace tfm
