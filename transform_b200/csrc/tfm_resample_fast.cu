// tfm_resample_fast.cu -- explicit instantiations of one K1 kernel family (parallel compilation only).
#include "tfm_resample.cuh"

namespace tfm {

template cudaError_t launch1<float, float, Fast, float>(int, ResampleParams &, bool, bool, cudaStream_t, const char **);
template cudaError_t launch_filter<float, float, Fast, float>(const ResampleParams &, const FilterParams &, cudaStream_t, const char **);

}  // namespace tfm
