// tfm_apply.cu -- K3: streaming Transform.apply over [n_vec, vec_len] arrays.
//
// Replaces Transform.apply / invert (core.py:218-315) for 2-D transforms: the first two
// components go through the chain in registers, components 2.. are copied through
// (core.py:271-276, 287-292).  The reference materialises one full float64 temporary per chain
// stage; here every vector is read once and written once (HBM-bound for affine chains,
// fp64-pipe-bound once a chain contains sincos/atan2).
//
// Layout: interleaved (x, y[, extras]) rows, densely packed.  vec_len == 2 is the hot case:
// one thread moves whole 16-byte packets (one fp64 vector, or two fp32 vectors) so that every
// warp access is a contiguous 512-byte run; a persistent grid-stride loop (grid = SMs x resident
// CTAs) keeps UNROLL packets in flight per thread.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "tfm_chain.cuh"
#include "tfm_common.cuh"

namespace tfm {

struct ApplyParams {
    const void *in;
    void *out;
    long long n_vec;
    int vec_len;
    ChainDev chain;
};

template <class A, int UNROLL, int SPH>
__global__ void __launch_bounds__(256)
k_apply_f64x2(const __grid_constant__ ApplyParams P)
{
    const double2 *in = static_cast<const double2 *>(P.in);
    double2 *out = static_cast<double2 *>(P.out);
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < P.n_vec; i += UNROLL * stride) {
        double2 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ldcs(in + i + u * stride);
        double x[UNROLL], y[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { x[u] = v[u].x; y[u] = v[u].y; }
        chain_eval_n<A, SPH, UNROLL>(P.chain, x, y);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) __stcs(out + i + u * stride, make_double2(x[u], y[u]));
    }
    for (; i < P.n_vec; i += stride) {
        double2 v = __ldcs(in + i);
        chain_eval<A, SPH>(P.chain, v.x, v.y);
        __stcs(out + i, v);
    }
}

// two fp32 vectors per 16-byte packet
template <class A, int UNROLL, int SPH>
__global__ void __launch_bounds__(256)
k_apply_f32x2(const __grid_constant__ ApplyParams P)
{
    const long long n_pk = P.n_vec >> 1;
    const float4 *in = static_cast<const float4 *>(P.in);
    float4 *out = static_cast<float4 *>(P.out);
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n_pk; i += UNROLL * stride) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ldcs(in + i + u * stride);
        double x[2 * UNROLL], y[2 * UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            x[2 * u] = v[u].x; y[2 * u] = v[u].y; x[2 * u + 1] = v[u].z; y[2 * u + 1] = v[u].w;
        }
        chain_eval_n<A, SPH, 2 * UNROLL>(P.chain, x, y);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            __stcs(out + i + u * stride, make_float4((float)x[2 * u], (float)y[2 * u],
                                                     (float)x[2 * u + 1], (float)y[2 * u + 1]));
    }
    for (; i < n_pk; i += stride) {
        float4 v = __ldcs(in + i);
        double x0 = v.x, y0 = v.y, x1 = v.z, y1 = v.w;
        chain_eval<A, SPH>(P.chain, x0, y0);
        chain_eval<A, SPH>(P.chain, x1, y1);
        __stcs(out + i, make_float4((float)x0, (float)y0, (float)x1, (float)y1));
    }
    if ((P.n_vec & 1) && blockIdx.x == 0 && threadIdx.x == 0) {       // odd tail vector
        const float *fi = static_cast<const float *>(P.in) + 2 * (P.n_vec - 1);
        float *fo = static_cast<float *>(P.out) + 2 * (P.n_vec - 1);
        double x = fi[0], y = fi[1];
        chain_eval<A, SPH>(P.chain, x, y);
        fo[0] = (float)x;
        fo[1] = (float)y;
    }
}

// general vec_len (>= 2): one thread per vector, extras copied through
template <class A, typename T, int SPH>
__global__ void __launch_bounds__(256)
k_apply_generic(const __grid_constant__ ApplyParams P)
{
    const T *in = static_cast<const T *>(P.in);
    T *out = static_cast<T *>(P.out);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int L = P.vec_len;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n_vec; i += stride) {
        const T *vi = in + i * L;
        T *vo = out + i * L;
        double x = (double)vi[0], y = (double)vi[1];
        chain_eval<A, SPH>(P.chain, x, y);
        for (int k = 2; k < L; ++k) vo[k] = vi[k];
        vo[0] = (T)x;
        vo[1] = (T)y;
    }
}

static int resident_grid(const void *kernel, long long work_items)
{
    int dev = 0, sms = 148, per_sm = 4;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1)
        per_sm = 4;
    long long want = (work_items + 255) / 256;
    long long grid = (long long)sms * per_sm;              // a whole number of waves, persistent
    if (want < grid) grid = want;
    return (int)(grid < 1 ? 1 : grid);
}

int apply_dispatch(const tfm_apply_args *a, cudaStream_t st)
{
    if (!a || a->abi_version != TFM_ABI_VERSION) return TFM_EVALUE;
    if (a->n_ops < 0 || a->n_ops > TFM_MAX_OPS || (a->n_ops && !a->chain)) return TFM_EVALUE;
    if (a->n_vec < 0 || a->vec_len < 2) return TFM_EVALUE;          // core.py:268-269, 284-285
    if (a->dtype != TFM_F32 && a->dtype != TFM_F64) return TFM_EVALUE;
    if (a->n_vec == 0) return TFM_OK;
    if (!a->in || !a->out) return TFM_EVALUE;
    if (!chain_ops_valid(a->chain, a->n_ops)) return TFM_EVALUE;

    ApplyParams P;
    P.in = a->in; P.out = a->out; P.n_vec = a->n_vec; P.vec_len = a->vec_len;
    P.chain.pad = 0;
    if (a->precision == TFM_PREC_FAST32) {
        P.chain.n = fuse_chain_fast(a->chain, a->n_ops, P.chain.ops);       // fast mode: peephole fusion
    } else {
        P.chain.n = a->n_ops;
        for (int i = 0; i < a->n_ops; ++i) P.chain.ops[i] = a->chain[i];
    }

    const bool exact = (a->precision == TFM_PREC_EXACT64);
    const bool aligned = ((reinterpret_cast<uintptr_t>(a->in) | reinterpret_cast<uintptr_t>(a->out)) & 15) == 0;
    const char *name;
#define TFM_LAUNCH(KERNEL, ITEMS)                                                       \
    do {                                                                                \
        int grid = resident_grid((const void *)KERNEL, (ITEMS));                        \
        KERNEL<<<grid, 256, 0, st>>>(P);                                                \
    } while (0)
    const int sph = chain_sph_level(P.chain);
    // K<A, ..., SPH>: one instantiation per arithmetic policy and "heavy op" level (tfm_chain.cuh)
#define TFM_L2(K, PRE, ITEMS)                                                                       \
    do {                                                                                             \
        if (exact) {                                                                                 \
            if (sph == 2) TFM_LAUNCH((K<Exact, PRE, 2>), ITEMS);                                     \
            else if (sph == 1) TFM_LAUNCH((K<Exact, PRE, 1>), ITEMS);                                \
            else TFM_LAUNCH((K<Exact, PRE, 0>), ITEMS);                                              \
        } else {                                                                                     \
            if (sph == 2) TFM_LAUNCH((K<Fast, PRE, 2>), ITEMS);                                      \
            else if (sph == 1) TFM_LAUNCH((K<Fast, PRE, 1>), ITEMS);                                 \
            else TFM_LAUNCH((K<Fast, PRE, 0>), ITEMS);                                               \
        }                                                                                            \
    } while (0)
    if (a->vec_len == 2 && aligned && a->dtype == TFM_F64) {
        TFM_L2(k_apply_f64x2, 4, a->n_vec);
        name = "k_apply_f64x2";
    } else if (a->vec_len == 2 && aligned && a->dtype == TFM_F32) {
        TFM_L2(k_apply_f32x2, 4, a->n_vec / 2 + 1);
        name = "k_apply_f32x2";
    } else if (a->dtype == TFM_F64) {
        TFM_L2(k_apply_generic, double, a->n_vec);
        name = "k_apply_generic<f64>";
    } else {
        TFM_L2(k_apply_generic, float, a->n_vec);
        name = "k_apply_generic<f32>";
    }
#undef TFM_L2
#undef TFM_LAUNCH
    count_launch();
    set_last_kernel(name);
    return cudaGetLastError() == cudaSuccess ? TFM_OK : TFM_ECUDA;
}

// ---------------------------------------------------------------------------------------------
// K3-ND: Linear-family chains of dimension 3..8 (basic.py:141-176 under core.py:255-292)
// ---------------------------------------------------------------------------------------------
// Streaming, grid-stride, one vector per thread and iteration.  Row i of the matrix product is rounded as
// NumPy's np.matmul(m, v[..., None]) rounds it (measured against NumPy 2.3 / OpenBLAS 0.3.30 on x86-64,
// the way DESIGN.md section 2 describes the 2 x 2 case): C-contiguous matrix s = m_i1 v_1, s = fma(m_i0, v_0, s),
// then s = fma(m_ik, v_k, s) for k = 2..; a transposed view (Rotation's matinv) of dimension <= 3 runs the
// chain from k = 0.  dim = 2 reduces to the two orders of the 2-D opcode.
struct LinNdParams {
    const void *in;
    void *out;
    long long n_vec;
    int vec_len, dim, n_ops, pad;
    tfm_linear_nd_op ops[TFM_MAX_OPS];
};

template <typename T>
__global__ void __launch_bounds__(256)
k_apply_linear_nd(const __grid_constant__ LinNdParams P)
{
    const T *in = static_cast<const T *>(P.in);
    T *out = static_cast<T *>(P.out);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int L = P.vec_len, D = P.dim;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n_vec; i += stride) {
        const T *vi = in + i * L;
        T *vo = out + i * L;
        double v[TFM_LIN_ND_MAX], w[TFM_LIN_ND_MAX];
#pragma unroll
        for (int k = 0; k < TFM_LIN_ND_MAX; ++k) v[k] = (k < D) ? (double)vi[k] : 0.0;
        for (int o = 0; o < P.n_ops; ++o) {
            const tfm_linear_nd_op &op = P.ops[o];
            if (op.flags & TFM_LIN_HAS_PRE) {
#pragma unroll
                for (int k = 0; k < TFM_LIN_ND_MAX; ++k) if (k < D) v[k] = __dadd_rn(v[k], op.add_before[k]);
            }
            if (op.flags & TFM_LIN_HAS_MATRIX) {
                const bool from0 = (op.flags & TFM_LIN_MAT_FORDER) && D <= 3;
#pragma unroll
                for (int r = 0; r < TFM_LIN_ND_MAX; ++r) {
                    if (r >= D) { w[r] = 0.0; continue; }
                    const double *m = op.matrix + r * D;
                    double s;
                    if (D == 1) s = __dmul_rn(m[0], v[0]);
                    else if (from0) s = __fma_rn(m[1], v[1], __dmul_rn(m[0], v[0]));
                    else s = __fma_rn(m[0], v[0], __dmul_rn(m[1], v[1]));
#pragma unroll
                    for (int k = 2; k < TFM_LIN_ND_MAX; ++k) if (k < D) s = __fma_rn(m[k], v[k], s);
                    w[r] = s;
                }
#pragma unroll
                for (int k = 0; k < TFM_LIN_ND_MAX; ++k) v[k] = w[k];
            }
            if (op.flags & TFM_LIN_HAS_POST) {
#pragma unroll
                for (int k = 0; k < TFM_LIN_ND_MAX; ++k) if (k < D) v[k] = __dadd_rn(v[k], op.add_after[k]);
            }
        }
        for (int k = D; k < L; ++k) vo[k] = vi[k];
#pragma unroll
        for (int k = 0; k < TFM_LIN_ND_MAX; ++k) if (k < D) vo[k] = (T)v[k];
    }
}

int apply_linear_nd_dispatch(const tfm_apply_nd_args *a, cudaStream_t st)
{
    if (!a || a->abi_version != TFM_ABI_VERSION) return TFM_EVALUE;
    if (a->n_ops < 0 || a->n_ops > TFM_MAX_OPS || (a->n_ops && !a->chain)) return TFM_EVALUE;
    if (a->dim < 1 || a->dim > TFM_LIN_ND_MAX || a->vec_len < a->dim || a->n_vec < 0) return TFM_EVALUE;
    if (a->dtype != TFM_F32 && a->dtype != TFM_F64) return TFM_EVALUE;
    if (a->n_vec == 0) return TFM_OK;
    if (!a->in || !a->out) return TFM_EVALUE;
    LinNdParams P;
    memset(&P, 0, sizeof(P));
    P.in = a->in; P.out = a->out; P.n_vec = a->n_vec; P.vec_len = a->vec_len; P.dim = a->dim; P.n_ops = a->n_ops;
    for (int i = 0; i < a->n_ops; ++i) P.ops[i] = a->chain[i];
    if (a->dtype == TFM_F64) {
        const int grid = resident_grid((const void *)k_apply_linear_nd<double>, a->n_vec);
        k_apply_linear_nd<double><<<grid, 256, 0, st>>>(P);
        set_last_kernel("k_apply_linear_nd<f64>");
    } else {
        const int grid = resident_grid((const void *)k_apply_linear_nd<float>, a->n_vec);
        k_apply_linear_nd<float><<<grid, 256, 0, st>>>(P);
        set_last_kernel("k_apply_linear_nd<f32>");
    }
    count_launch();
    return cudaGetLastError() == cudaSuccess ? TFM_OK : TFM_ECUDA;
}

}  // namespace tfm
