/* oracle/oracle_c.c -- C restatement of the compiled parts of the reference's hot path.
 *
 * TEST INFRASTRUCTURE ONLY: built into oracle/_build/liboracle_c.so by oracle/build_oracle.py
 * and loaded (ctypes) only by oracle/oracle_np.py, i.e. by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  The product (transform_b200) never links it.
 *
 * What is restated here and where it comes from (paths under /root/reference/transform/):
 *   orc_matvec_fma      basic.py:150-152 / 168-170 -- np.matmul of an NxM matrix with a stack of
 *                       M-vectors.  NumPy 2.3.5 + OpenBLAS 0.3.30 on x86-64 evaluates the 2x2 case
 *                       as fma(m[i][0], x, fl(m[i][1]*y)) for a C-contiguous matrix and as
 *                       fma(m[i][1], y, fl(m[i][0]*x)) for a transposed view (gemv N vs T
 *                       kernels; measured bit-exact in the authoring container), and continues a
 *                       3x3 row with one more fma in either case.  This is the arithmetic
 *                       convention of the CUDA fp64-exact mode as well.
 *   orc_svd2x2          helpers.pyx:1845-1894 (svd2x2_fast)
 *   orc_svc2x2          helpers.pyx:1901-1916 (svc2x2_fast)
 *   orc_interp_jacobian helpers.pyx:1542-1803 (interpND_jacobian) with the repairs listed in
 *                       oracle_np.interp_grid_jacobian (R4 assert(0) removed, R5 integer reg_size,
 *                       R6 bound as C int, R7 weight-counted/value-skipped truncation kept,
 *                       R8 padding formula as coded, R9 mirror = 2s-1-i as apply_boundary:172).
 *
 * PARITY UNPINNED for orc_interp_jacobian: the reference ships it unreachable (assert(0) at
 * helpers.pyx:1640, result dropped at 1520-1530) and has no test for it.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define ORC_OK 0
#define ORC_EVALUE 1        /* -> ValueError (bad method / boundary) */
#define ORC_EFORBID 2       /* -> ValueError: 'forbid' boundary violated (helpers.pyx:1753,1774) */

/* out[n][rows] = M[rows][cols] . v[n][cols]   (v may have a wider stride: extra components)
 * The order in which np.matmul(m, v[..., None]) accumulates a row, measured in the authoring container
 * (NumPy 2.3.5, OpenBLAS 0.3.30, x86-64; tests/test_oracle_kat.py::test_matvec_rounding_matches_numpy
 * re-measures it wherever the CPU suite runs):
 * f_order = 0: M was C-contiguous in the reference  -> acc = fl(m[i][1]*v[1]); fma(m[i][0], v[0], acc);
 *              then fma(m[i][k], v[k], acc) for k = 2..  (bit-exact for 2 x 2: 6000/6000, 3 x 3: 6000/6000)
 * f_order = 1: M was a transposed (Fortran) view, e.g. Rotation's matinv = out.transpose()
 *              (basic.py:501)                      -> cols <= 3: acc = fl(m[i][0]*v[0]), fma upwards
 *              (2 x 2 and 3 x 3 bit-exact); cols = 4: the f_order = 0 order (bit-exact, 720/720).
 * Anything larger is not pinned (OpenBLAS' vectorised kernels sum partial accumulators): a few ulp. */
void orc_matvec_fma(const double *M, int rows, int cols, int f_order,
                    const double *v, int64_t n, int64_t v_stride,
                    double *out, int64_t out_stride)
{
    for (int64_t p = 0; p < n; ++p) {
        const double *vp = v + p * v_stride;
        double *op = out + p * out_stride;
        for (int i = 0; i < rows; ++i) {
            const double *mi = M + (size_t)i * cols;
            double acc;
            if (cols == 1) {
                acc = mi[0] * vp[0];
            } else if (f_order && cols <= 3) {
                acc = mi[0] * vp[0];
                for (int k = 1; k < cols; ++k)
                    acc = fma(mi[k], vp[k], acc);
            } else {
                acc = fma(mi[0], vp[0], mi[1] * vp[1]);
                for (int k = 2; k < cols; ++k)
                    acc = fma(mi[k], vp[k], acc);
            }
            op[i] = acc;
        }
    }
}

/* M = U diag(s) V^T; U = R(theta), W = R(phi), V = W diag(sign)   (helpers.pyx:1845-1894) */
void orc_svd2x2(const double *M, double *U, double *s, double *V)
{
    double a = M[0], b = M[1], c = M[2], d = M[3];
    double au = a * a + b * b, bu = a * c + b * d, du = c * c + d * d;
    double theta = 0.5 * atan2(2 * bu, au - du);
    double St = sin(theta), Ct = cos(theta);
    U[0] = Ct; U[1] = -St; U[2] = St; U[3] = Ct;
    double aw = a * a + c * c, bw = a * b + c * d, dw = b * b + d * d;
    double phi = 0.5 * atan2(2 * bw, aw - dw);
    double Sp = sin(phi), Cp = cos(phi);
    double cs00 = (Ct * a + St * c) * Cp + (Ct * b + St * d) * Sp;
    double cs11 = (St * a - Ct * c) * Sp - (St * b - Ct * d) * Cp;
    s[0] = fabs(cs00);
    s[1] = fabs(cs11);
    double c00 = cs00 >= 0 ? 1.0 : -1.0;
    double c11 = cs11 >= 0 ? 1.0 : -1.0;
    V[0] = Cp * c00; V[1] = -Sp * c11; V[2] = Sp * c00; V[3] = Cp * c11;
}

/* helpers.pyx:1901-1916 */
void orc_svc2x2(const double *U, const double *s, const double *V, double *M)
{
    double U00S0 = U[0] * s[0], U10S0 = U[2] * s[0];
    double U01S1 = U[1] * s[1], U11S1 = U[3] * s[1];
    M[0] = U00S0 * V[0] + U01S1 * V[1];
    M[1] = U00S0 * V[2] + U01S1 * V[3];
    M[2] = U10S0 * V[0] + U11S1 * V[1];
    M[3] = U10S0 * V[2] + U11S1 * V[3];
}

/* One axis of helpers.pyx:1750-1790.  Returns 0 ok, 1 truncated tap, 2 forbid violation. */
static int orc_bound1(int64_t *pi, int64_t size, int bound)
{
    int64_t i = *pi;
    if (i >= 0 && i < size) return 0;
    switch (bound) {
    case 1: return 2;                                   /* forbid */
    case 2: return 1;                                   /* truncate */
    case 3: *pi = i < 0 ? 0 : size - 1; return 0;       /* extend */
    case 4: {                                           /* periodic: C-truncating divide, then fix-up */
        int64_t ia = i / size;
        i -= ia * size;
        if (i < 0) i += size;
        *pi = i; return 0; }
    case 5: {                                           /* mirror (R9: 2s-1-i, helpers.pyx:172) */
        int64_t ia = i / (size * 2);
        i -= ia * size * 2;
        if (i < 0) i += size * 2;
        if (i >= size) i = 2 * size - 1 - i;
        *pi = i; return 0; }
    default: return 2;
    }
}

/* source[sh][sw] f64, index[ny][nx][2], jac[ny][nx][2][2] -> dest[ny][nx]
 * method: l=1 z=2 h=3 t=4 g=5 (helpers.pyx:1501), w=6 (Hann window of SVD.pyx:110-111); bound: f=1 t=2 e=3 p=4 m=5 (helpers.pyx:1511).
 * max_reg: 0 = unlimited (reference behaviour); >0 caps reg_size (the documented-but-unused
 *          sv_limit guard, helpers.pyx:1441-1445, made explicit; the CUDA path takes the same cap). */
int orc_interp_jacobian(const double *source, int64_t sh, int64_t sw,
                        const double *index, const double *jac, int64_t ny, int64_t nx,
                        int method, int bound_x, int bound_y,
                        double fillvalue, double oblur, double iblur, double pad_pow,
                        int64_t max_reg, double *dest)
{
    (void)fillvalue; (void)iblur;                       /* unused by the reference too (R7) */
    double padval;
    const double pi = 3.141592653589793;
    if (method == 1 || method == 3 || method == 4 || method == 6) padval = 1; /* helpers.pyx:1612-1615; 6 as 3 */
    else if (method == 2 || method == 5) padval = 3;
    else return ORC_EVALUE;
    if (bound_x < 1 || bound_x > 5 || bound_y < 1 || bound_y > 5) return ORC_EVALUE;

    for (int64_t yd = 0; yd < ny; ++yd) {
        for (int64_t xd = 0; xd < nx; ++xd) {
            const double *ip = index + (yd * nx + xd) * 2;
            const double *Ji = jac + (yd * nx + xd) * 4;
            double U[4], V[4], s[2], J[4];
            orc_svd2x2(Ji, U, s, V);                                             /* 1659 */
            s[0] = exp(-log(padval + exp(log(s[0]) * pad_pow)) / pad_pow);       /* 1661 */
            s[1] = exp(-log(padval + exp(log(s[1]) * pad_pow)) / pad_pow);       /* 1662 */
            orc_svc2x2(V, s, U, J);                                              /* 1663 */
            double smin = s[0] < s[1] ? s[0] : s[1];
            double rs = 2 * ceil(oblur / smin);                                  /* 1666 */
            int64_t reg_size = (rs < 1e15) ? (int64_t)rs : (int64_t)1e15;
            if (max_reg > 0 && reg_size > max_reg) reg_size = max_reg;
            int64_t half = reg_size / 2;
            if (!(fabs(ip[0]) < 4.0e18) || !(fabs(ip[1]) < 4.0e18)) {
                /* NaN / inf / absurd coordinate: in the reference every filter value is NaN, so
                 * wgt is NaN, `wgt > 0` is false and the pixel is 0 (helpers.pyx:1800). */
                dest[yd * nx + xd] = 0;
                continue;
            }
            double xsf = floor(ip[0]), ysf = floor(ip[1]);
            /* <int>floor(): out-of-range / NaN doubles give INT64_MIN on x86-64 */
            int64_t xs = (fabs(xsf) < 9.2e18) ? (int64_t)xsf : INT64_MIN;
            int64_t ys = (fabs(ysf) < 9.2e18) ? (int64_t)ysf : INT64_MIN;
            double xf_of = ip[0] - (double)xs, yf_of = ip[1] - (double)ys;
            double val = 0.0, wgt = 0.0;
            for (int64_t yr = -half; yr <= half; ++yr) {
                for (int64_t xr = -half; xr <= half; ++xr) {
                    double x0 = (double)xr + xf_of, y0 = (double)yr + yf_of;
                    double xf = fabs(J[0] * x0 + J[1] * y0) / oblur;             /* 1689 */
                    double yf = fabs(J[2] * x0 + J[3] * y0) / oblur;             /* 1690 */
                    double filt, a, b;
                    switch (method) {
                    case 1:
                        filt = (xf < 1) ? 1 - xf : 0;
                        filt *= (yf < 1) ? 1 - yf : 0;
                        break;
                    case 2:
                        if (xf > 3 || yf > 3) filt = 0;
                        else {
                            a = pi * xf; b = a / 3.0; filt = (sin(a) / a * sin(b) / b);
                            a = pi * yf; b = a / 3.0; filt *= (sin(a) / a * sin(b) / b);
                        }
                        break;
                    case 3:
                        if (xf > 1 || yf > 1) filt = 0;
                        else {
                            b = cos(pi * xf); filt = b * b;
                            b = cos(pi * yf); filt *= b * b;
                        }
                        break;
                    case 4:
                        if (xf > 1 || yf > 1) filt = 0;
                        else {
                            filt = 1;
                            if (xf > 0.5) { b = cos(pi * 2 * (xf - 0.5)); filt = b * b; }
                            if (yf > 0.5) { b = cos(pi * 2 * (yf - 0.5)); filt *= b * b; }
                        }
                        break;
                    case 6:
                        /* the Hann window the reference's OTHER anti-aliaser uses (SVD.pyx:110-111):
                         * (cos(pi min(|x|,1)) + 1)(cos(pi min(|y|,1)) + 1) / 2 -- helpers.pyx:1708-1717
                         * (case 3) squares cos(pi x) instead, which is not a window */
                        filt = (cos((xf < 1 ? xf : 1) * pi) + 1.0) * (cos((yf < 1 ? yf : 1) * pi) + 1.0) / 2.0;
                        break;
                    default:
                        filt = exp(-(xf * xf + yf * yf));                        /* 1735 */
                    }
                    wgt += filt;                                                 /* 1740 */
                    int64_t yrs = ys + yr, xrs = xs + xr;
                    int fy = orc_bound1(&yrs, sh, bound_y);
                    int fx = orc_bound1(&xrs, sw, bound_x);
                    if (fy == 2 || fx == 2) return ORC_EFORBID;
                    if (fy == 0 && fx == 0)
                        val += source[yrs * sw + xrs] * filt;                    /* 1793-1796 */
                }
            }
            dest[yd * nx + xd] = (wgt > 0) ? val / wgt : 0;                      /* 1800 */
        }
    }
    return ORC_OK;
}
