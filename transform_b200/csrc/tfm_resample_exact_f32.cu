// tfm_resample_exact_f32.cu -- explicit instantiations of one K1 kernel family (parallel compilation only).
#include "tfm_resample.cuh"

namespace tfm {

template cudaError_t launch1<float, double, Exact, double>(int, ResampleParams &, bool, bool, cudaStream_t, const char **);
template cudaError_t launch2<TFM_NEAREST, float, float, Exact, double>(ResampleParams &, bool, bool, cudaStream_t, const char **);
template cudaError_t launch_filter<float, double, Exact, double>(const ResampleParams &, const FilterParams &, cudaStream_t, const char **);

}  // namespace tfm
