// tfm_resample_exact_f64.cu -- explicit instantiations of one K1 kernel family (parallel compilation only).
#include "tfm_resample.cuh"

namespace tfm {

template cudaError_t launch1<double, double, Exact, double>(int, ResampleParams &, bool, bool, cudaStream_t, const char **);
template cudaError_t launch_filter<double, double, Exact, double>(const ResampleParams &, const FilterParams &, cudaStream_t, const char **);

}  // namespace tfm
