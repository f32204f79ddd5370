/* examples/c_abi_example.c -- the C ABI of libtfm_b200.so from plain C (no torch, no Python).
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_example.c -Ltransform_b200/lib -ltfm_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -o c_abi_example
 *
 * Resamples a small float32 frame through Scale(1.3) o Rotation(30 deg) o Offset with the linear
 * method in parity mode (float64 out), the way Transform.resample lowers that composition:
 * the inverse chain in execution order is {Scale rev, Rotation rev, Offset rev} (core.py:1556-1559).
 * Device memory comes from the CUDA runtime; the library itself only needs device pointers. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "tfm_b200.h"

/* the four CUDA runtime calls this example needs, declared by hand to stay free of CUDA headers */
extern int cudaMalloc(void **p, size_t n);
extern int cudaMemcpy(void *dst, const void *src, size_t n, int kind);
extern int cudaDeviceSynchronize(void);
extern int cudaFree(void *p);

static tfm_op linear_rev(const double *matinv, const double *pre)
{
    tfm_op op;
    memset(&op, 0, sizeof op);
    op.kind = TFM_OP_LINEAR;
    op.dir = TFM_REV;
    if (matinv) {
        op.flags |= TFM_LIN_HAS_MATRIX;
        memcpy(op.p, matinv, 4 * sizeof(double));
    }
    if (pre) {
        op.flags |= TFM_LIN_HAS_PRE;
        op.p[4] = pre[0];
        op.p[5] = pre[1];
    }
    return op;
}

int main(void)
{
    enum { H = 96, W = 128 };
    const double c = cos(30.0 * 3.14159265358979323846 / 180.0), s = sin(30.0 * 3.14159265358979323846 / 180.0);
    const double scale_inv[4] = {1.0 / 1.3, 0.0, 0.0, 1.0 / 1.3};
    const double rot_inv[4] = {c, s, -s, c};            /* transpose of the rotation matrix */
    const double offset[2] = {-W / 2.0, -H / 2.0};
    tfm_op chain[3];
    tfm_resample_args a;
    float *h_src = malloc(sizeof(float) * H * W);
    double *h_dst = malloc(sizeof(double) * H * W);
    void *d_src = NULL, *d_dst = NULL;
    int i, rc;

    chain[0] = linear_rev(scale_inv, NULL);
    chain[1] = linear_rev(rot_inv, NULL);
    chain[1].flags |= TFM_LIN_MAT_FORDER;               /* Rotation.matinv is a transposed view in the reference */
    chain[2] = linear_rev(NULL, offset);
    for (i = 0; i < H * W; ++i) h_src[i] = (float)(i % 251) / 251.0f;
    if (cudaMalloc(&d_src, sizeof(float) * H * W) || cudaMalloc(&d_dst, sizeof(double) * H * W)) return 2;
    cudaMemcpy(d_src, h_src, sizeof(float) * H * W, 1 /* host to device */);

    memset(&a, 0, sizeof a);
    a.abi_version = TFM_ABI_VERSION;
    a.n_ops = 3;
    a.chain = chain;
    a.src.data = d_src; a.src.dtype = TFM_F32; a.src.h = H; a.src.w = W; a.src.pitch = W * sizeof(float);
    a.src.full_h = H; a.src.full_w = W; a.src.n_planes = 1; a.src.plane_pitch = (int64_t)H * W * sizeof(float);
    a.dst.data = d_dst; a.dst.dtype = TFM_F64; a.dst.h = H; a.dst.w = W; a.dst.pitch = W * sizeof(double);
    a.dst.full_h = H; a.dst.full_w = W; a.dst.n_planes = 1; a.dst.plane_pitch = (int64_t)H * W * sizeof(double);
    a.method = TFM_LINEAR;
    a.bound_x = a.bound_y = TFM_B_TRUNCATE;
    a.precision = TFM_PREC_EXACT64;
    rc = tfm_resample(&a, NULL /* default stream */);
    if (rc != TFM_OK) { fprintf(stderr, "tfm_resample: %s\n", tfm_strerror(rc)); return 1; }
    cudaDeviceSynchronize();
    cudaMemcpy(h_dst, d_dst, sizeof(double) * H * W, 2 /* device to host */);
    printf("kernel %s, out[%d][%d] = %.17g\n", tfm_last_kernel(), H / 2, W / 2, h_dst[(H / 2) * W + W / 2]);
    cudaFree(d_src); cudaFree(d_dst); free(h_src); free(h_dst);
    return 0;
}
