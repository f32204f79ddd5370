/* tfm_b200.h -- C ABI of libtfm_b200.so: the B200 (sm_100a) replacement for the image-resampling
 * hot path of drzowie/transform.
 *
 * The reference has no FFI: its seam for this path is Python-level.  Each
 * entry point below names the reference interface it stands in for (paths under
 * /root/reference/transform/).  The Python host side (transform_b200/) mirrors the reference's
 * classes and binds these symbols with ctypes; INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions (same as the reference, core.py:48-59): vectors are (X,Y) in the last axis, images
 * are row-major [Y][X]; pixel centres sit on integers; (0,0) is the centre of the first pixel.
 * All data pointers are DEVICE pointers unless a name ends in _host.  Calls are stream-ordered on
 * the cudaStream_t passed as `stream` (a void* here so the header needs no CUDA include) and may be
 * issued concurrently on different streams / devices.  Nothing throws across the boundary: every
 * function returns a tfm_status.
 * Process-wide state, all of it: the launch counter behind tfm_launch_count(), the thread-local name
 * behind tfm_last_kernel(), and two mutex-protected LRU caches of driver objects laid over CALLER memory
 * (256 texture objects, 16 tensor maps, keyed by device / pointer / geometry -- they describe address
 * ranges, not contents).  Neither cache ever blocks a stream: a texture object evicted while a launch
 * may still read through it is destroyed later, once an event recorded behind that launch has completed.
 */
#ifndef TFM_B200_H
#define TFM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TFM_API __attribute__((visibility("default")))
#else
#define TFM_API
#endif

#define TFM_ABI_VERSION 2
#define TFM_MAX_OPS 16          /* flattened chain length (Composition flattening, core.py:1534-1537) */
#define TFM_OP_NPARAM 24

/* ---- status codes: 1:1 with the exceptions of the reference (SURVEY.md section 8b "Errors") ---- */
typedef enum tfm_status {
    TFM_OK = 0,
    TFM_EVALUE = 1,        /* ValueError: bad dims / shape / method / boundary (core.py:566-571,
                              helpers.pyx:174-176, 840-844, 1505, 1516) */
    TFM_EDIRECTION = 2,    /* AssertionError: transform invalid in that direction (core.py:265-266) */
    TFM_EFORBID = 3,       /* ValueError: 'forbid' boundary violated (helpers.pyx:141-144, 1753) */
    TFM_ECUDA = 4,         /* CUDA runtime error (no reference counterpart) */
    TFM_EUNSUPPORTED = 5,  /* valid in the reference but outside this library's scope (fails loudly) */
    TFM_ESTAGED = 6        /* a tap fell outside the staged source sub-rectangle of a shard */
} tfm_status;

/* ---- transform chain ----
 * A chain is the reference's flattened Composition AFTER direction resolution: element i is
 * applied i-th; `dir` says whether that element runs its _forward or its _reverse.
 * Transform.invert(grid) on Composition([A,B,C]) lowers to {A rev, B rev, C rev}
 * (core.py:1556-1559); Transform.apply lowers to {C fwd, B fwd, A fwd} (core.py:1551-1554). */
typedef enum tfm_op_kind {
    TFM_OP_IDENTITY = 0,   /* core.py:1345-1401 */
    TFM_OP_LINEAR = 1,     /* basic.py:12-218 (+Scale 224-336, Rotation 338-506, Offset 508-551) */
    TFM_OP_RADIAL = 2,     /* basic.py:553-716 */
    TFM_OP_WCS = 3,        /* core.py:1650-1747 (astropy.wcs all_pix2world/all_world2pix, origin 0) */
    TFM_OP_INDEX = 4,      /* coordinate lookup: (x,y) = index[y][x][0..1].  Not a Transform of the
                              reference: it is how interpND(source, index, ...) / interpND_grid
                              (helpers.pyx:379, 1293), which take a precomputed coordinate array,
                              run through the same kernels.  Must be the first element of a chain. */
    TFM_OP_PIN = 5,        /* per-axis pincushion family: Quadpin (basic.py:846-1017), Cubicpin
                              (1020-1175), Poly / Quarticpin (1178-1462, forward only) */
    TFM_OP_SIP = 6,        /* SIP polynomial distortion of a FITS header (CTYPE '...-SIP'): the
                              pix2foc stage of astropy's all_pix2world / all_world2pix
                              (core.py:1727, 1742).  Precedes a TFM_OP_WCS in the pixel -> world
                              direction and follows it in the world -> pixel direction. */
    TFM_OP_TPV = 7         /* TPV polynomial distortion (CTYPE '...-TPV', PV1_m / PV2_m, m = 0..39): acts on
                              the intermediate world coordinates between the linear part of a header and
                              its gnomonic projection -- wcslib's "sequent distortion" inside wcsp2s / wcss2p,
                              which all_pix2world / all_world2pix call (core.py:1727, 1742).  The host lowers
                              such a header to {WCS linear part, TPV, WCS projection part}. */
} tfm_op_kind;

typedef enum tfm_dir { TFM_FWD = 0, TFM_REV = 1 } tfm_dir;

/* TFM_OP_LINEAR  fwd: out = post + M.(in + pre)   (basic.py:141-158)
 *                rev: out = Minv.(in - post) - pre (basic.py:160-176)
 *   p[0..3]  2x2 matrix used in THIS direction, row-major (matrix for fwd, matinv for rev)
 *   p[4..5]  pre     p[6..7] post
 *   Matrix-vector rounding in the fp64-exact mode follows NumPy's gemv as measured:
 *   C-contiguous matrix: fma(m[i][0], x, fl(m[i][1]*y)); transposed view (TFM_LIN_MAT_FORDER,
 *   e.g. Rotation's matinv = out.transpose(), basic.py:501): fma(m[i][1], y, fl(m[i][0]*x)). */
#define TFM_LIN_HAS_PRE 1
#define TFM_LIN_HAS_POST 2
#define TFM_LIN_HAS_MATRIX 4
#define TFM_LIN_MAT_FORDER 8

/* TFM_OP_RADIAL  (basic.py:670-711)
 *   p[0..1] origin   p[2] r0   p[3] ang_coeff (radians per output angle unit) */
#define TFM_RAD_CCW 1
#define TFM_RAD_POS_ONLY 2
#define TFM_RAD_R0_NOT_NONE 4   /* reverse tests `r0 is not None` (basic.py:703) */
#define TFM_RAD_R0_TRUTHY 8     /* forward tests truthiness of r0   (basic.py:686) */
#define TFM_RAD_ORIGIN_NONZERO 16 /* `if not np.all(origin == 0)` (basic.py:675, 709) */

/* TFM_OP_WCS  fwd = pixel(origin 0) -> world, rev = world -> pixel (core.py:1719-1747).
 *   p[0..1]  CRPIX (1-based, FITS)          p[2..5]  CD matrix (PC*CDELT folded), row-major
 *   p[6..9]  inverse of the CD matrix        p[10..11] CRVAL
 *   flags low byte = projection: 0 none (purely linear axes, e.g. 'Solar-X'), 1 TAN, 2 CAR
 *   p[12..20] 3x3 native->celestial rotation matrix (Calabretta & Greisen 2002 eq. 2/5, from
 *             CRVAL, LONPOLE, LATPOLE), row-major; the transpose is used for celestial->native */
#define TFM_WCS_PROJ_NONE 0
#define TFM_WCS_PROJ_TAN 1
#define TFM_WCS_PROJ_CAR 2
/* further projections of paper II (Calabretta & Greisen 2002), default projection parameters only */
#define TFM_WCS_PROJ_STG 3   /* stereographic            eq. 56-57 */
#define TFM_WCS_PROJ_SIN 4   /* orthographic              eq. 59-60; slant (xi, eta) = p[21], p[22] (PV2_1, PV2_2), eq. 61-62 */
#define TFM_WCS_PROJ_ARC 5   /* zenithal equidistant     eq. 67-68 */
#define TFM_WCS_PROJ_ZEA 6   /* zenithal equal-area      eq. 69-70 */
#define TFM_WCS_PROJ_CEA 7   /* cylindrical equal-area    eq. 80-81; lambda = p[21] (PV2_1; 0 reads as the default 1) */
#define TFM_WCS_PROJ_MER 8   /* Mercator                 eq. 85-86 */
#define TFM_WCS_PROJ_SFL 9   /* Sanson-Flamsteed         eq. 87-88 */
#define TFM_WCS_PROJ_AIT 10  /* Hammer-Aitoff            eq. 108-110 */
#define TFM_WCS_PROJ_MAX 10

/* TFM_OP_SIP  (Shupe et al. 2005; PARITY UNPINNED like the spherical projections: astropy is absent)
 *   p[0] the BIT PATTERN of a device pointer to float64 coef[2][TFM_SIP_DIM][TFM_SIP_DIM]:
 *        coef[0][p][q] = A_p_q, coef[1][p][q] = B_p_q (absent keywords are 0); the table must stay
 *        allocated while launches that use it are in flight
 *   p[1] A_ORDER   p[2] B_ORDER  (0..TFM_SIP_MAX_ORDER)     p[3..4] CRPIX (1-based, FITS)
 *   fwd: u = x + 1 - CRPIX1, v = y + 1 - CRPIX2;  x += sum A_p_q u^p v^q,  y += sum B_p_q u^p v^q
 *   rev: the fixed-point iteration of all_world2pix, pix <- pix0 - f(pix), which uses the FORWARD
 *        polynomials only (AP_p_q / BP_p_q are ignored there too); iterated per point until the
 *        step is below 1e-12 pixel or TFM_SIP_MAXITER evaluations (astropy's maxiter = 20; its
 *        stopping rule is 1e-4 pixel on the worst point of the batch, so its result depends on the
 *        batch -- this one does not, and agrees with it to within that tolerance) */
#define TFM_SIP_MAX_ORDER 9
#define TFM_SIP_DIM 10
#define TFM_SIP_MAXITER 20

/* TFM_OP_TPV  (the TPV convention of the FITS WCS registry, as written by SCAMP; PARITY UNPINNED)
 *   p[0] the BIT PATTERN of a device pointer to float64 coef[2][TFM_TPV_NCOEF]: coef[0][m] = PV1_m,
 *        coef[1][m] = PV2_m (defaults: PV1_1 = PV2_1 = 1, everything else 0)
 *   p[1..4] inverse of the linear part [[PV1_1, PV1_2], [PV2_2, PV2_1]], row-major (the reverse direction's
 *        quasi-Newton step)
 *   fwd: xi' = P1(xi, eta), eta' = P2(eta, xi) with r = sqrt(xi^2 + eta^2) and
 *        P(x, y) = c0 + c1 x + c2 y + c3 r + c4 x^2 + c5 xy + c6 y^2 + c7 x^3 + c8 x^2 y + c9 x y^2 + c10 y^3
 *                + c11 r^3 + c12 x^4 + ... + c16 y^4 + c17 x^5 + ... + c22 y^5 + c23 r^5 + c24 x^6 + ...
 *                + c30 y^6 + c31 x^7 + ... + c38 y^7 + c39 r^7       (degrees, like the CD matrix's output)
 *   rev: v <- v - Minv.(T(v) - target) from v = Minv.(target - c0), per point, until the step is below
 *        1e-13 degree or TFM_TPV_MAXITER evaluations (wcslib inverts the same map iteratively to its own
 *        tolerance; both converge to the root) */
#define TFM_TPV_NCOEF 40
#define TFM_TPV_MAXITER 30

/* TFM_OP_INDEX
 *   p[0] the BIT PATTERN of a device pointer to float64 index[ny][nx][2] (dense, (X,Y) last axis)
 *   p[1] nx   p[2] ny   (grid points outside [0,nx) x [0,ny) yield NaN coordinates) */

/* TFM_OP_PIN  -- elementwise per axis a (0 = X, 1 = Y); parameter block of axis a at p[12a ..]:
 *   [0] origin   [1..4] c0..c3   [5..11] length^(k-1) for k = 2..8 (Poly only)
 * The constants are the reference's own Python-scalar sub-expressions, evaluated by the host shim
 * exactly as the reference evaluates them (d = input component, o = origin):
 *   Quadpin  fwd (basic.py:932-967):  u=d-o; u = u + ((c0*(u*|u|))/c1); u = u/c2; +o
 *            c0 = strength, c1 = length, c2 = |strength|+1
 *   Quadpin  rev (970-1010): u = -1 + sqrt(1 + ((c0*|d-o|)*c1)); u = (u*c2)/c3; u *= (d<o ? -1 : 1); +o
 *            c0 = 4*strength/length, c1 = 1+|strength|, c2 = length, c3 = 2*strength
 *   Cubicpin fwd (1097-1128): u=d-o; q=u/c1; u = u + (((c0*u)*q)*q); u = u/c2; +o
 *            c0 = strength, c1 = length, c2 = strength^2+1
 *   Cubicpin rev (1130-1170), A = strength/length/length: u=(d-o)*c0; al = c1*(-u);
 *            in = sqrt(al*al + c2); root(x) = sign(x)*|x|^(1/3);
 *            u = c3*(root(0.5*(al+in)) + root(0.5*(al-in))); +o
 *            c0 = strength+1, c1 = 27*A*A, c2 = 4*(3A)^3, c3 = -1/(3A)
 *   Poly     fwd (1318-1358): u=d-o; a=u; for k=2..degree: a = a + ((c0*u^k)/length^(k-1)); u = c1*a; +o
 *            c0 = strength, c1 = 1/(|strength|^(degree-1)+1) (den=1) or 1/(|strength|+1);
 *            degree in tfm_op.reserved (2..8)
 * flags: bits 0-1 TFM_PIN_QUAD / CUBIC / POLY; TFM_PIN_SHIFTED: origin is not all-zero (the
 * subtraction and re-addition happen only then); TFM_PIN_DIM1: only the X component (dim=1). */
#define TFM_PIN_QUAD 0
#define TFM_PIN_CUBIC 1
#define TFM_PIN_POLY 2
#define TFM_PIN_SHIFTED 4
#define TFM_PIN_DIM1 8
#define TFM_PIN_TRUNC 16   /* cast the result to int64 (C truncation; NaN/inf -> INT64_MIN): what the
                              reference does when a pincushion op constructed with `dim=` is the first
                              op applied to INTEGER data -- e.g. the np.mgrid pixel grid of resample()
                              (core.py:574-578) -- because it writes its result back into the copy of
                              its input (basic.py:962-965, 1005-1008, 1123-1126, 1165-1168, 1353-1356) */
#define TFM_PIN_MAX_DEGREE 8

typedef struct tfm_op {
    int32_t kind;          /* tfm_op_kind */
    int32_t dir;           /* tfm_dir */
    int32_t flags;         /* per-kind bits above */
    int32_t reserved;
    double p[TFM_OP_NPARAM];
} tfm_op;

/* ---- enumerations shared by the resamplers ---- */
typedef enum tfm_dtype { TFM_F32 = 0, TFM_F64 = 1 } tfm_dtype;

/* interpND methods (helpers.pyx:541-611, 703-835); first letter in the reference.  The last six are
 * the "collapsible filters" (helpers.pyx:792-835): a size x size region around floor(v) weighted by
 * a separable window, size = 16 / 6 / 8 / 2 / 2 / 2 scaled by oblur (helpers.pyx:708-711).
 * TFM_TENT is the reference's method 'l' WITH oblur (without, 'l' is TFM_LINEAR). */
typedef enum tfm_method {
    TFM_NEAREST = 0, TFM_LINEAR = 1, TFM_CUBIC = 2,
    TFM_SINC = 3, TFM_LANCZOS = 4, TFM_GAUSSIAN = 5, TFM_HANN = 6, TFM_TUKEY = 7, TFM_TENT = 8
} tfm_method;
#define TFM_FILTER_MAX_SIZE 112  /* region edge limit of the filter kernel (shared-memory weights) */

/* boundary codes: the reference's own table (helpers.pyx:1511) */
typedef enum tfm_bound {
    TFM_B_FORBID = 1, TFM_B_TRUNCATE = 2, TFM_B_EXTEND = 3, TFM_B_PERIODIC = 4, TFM_B_MIRROR = 5
} tfm_bound;

/* Jacobian-filter shapes: the reference's own table (helpers.pyx:1501) */
typedef enum tfm_filter {
    TFM_FILT_TENT = 1, TFM_FILT_LANCZOS = 2, TFM_FILT_HANNING = 3, TFM_FILT_TUKEY = 4,
    TFM_FILT_GAUSSIAN = 5,
    /* not in helpers.pyx's table: the Hann window (cos(pi x) + 1)/2 per axis that the reference's other,
     * working anti-aliaser uses (SVD.pyx:110-111).  TFM_FILT_HANNING is helpers.pyx:1708-1717 as
     * coded -- cos^2(pi x), which has weight 1 at the EDGE of its support. */
    TFM_FILT_HANN_WINDOW = 6
} tfm_filter;

typedef enum tfm_precision {
    TFM_PREC_EXACT64 = 0,  /* parity mode: chain evaluated op by op in fp64 in the reference's
                              operation order (no FMA contraction except the documented gemv fma);
                              interpolation arithmetic in fp64 in the reference's order */
    TFM_PREC_FAST32 = 1    /* throughput mode: coordinates still fp64 (adjacent affine ops are
                              pre-composed by the host), interpolation weights/accumulation fp32 */
} tfm_precision;

/* tfm_image.flags / tfm_resample_args.flags */
#define TFM_F_REGION_SRC_DTYPE 1 /* cubic, exact mode: first Hermite pass in the SOURCE dtype (float32)
                                    because the reference's region was not promoted by the truncation
                                    pass (helpers.pyx:373-374, 755-762) */
#define TFM_F_FLUX 2 /* K2 only: phot='flux' (core.py:521-533, documented but not implemented there;
                        SVD.pyx:285-286 has it): scale each output by |det J| of the unpadded Jacobian */

/* One 2-D image plane in device memory.  (x0,y0) is the position of element [0][0] in the
 * coordinate system of the FULL image of size full_w x full_h: a whole image has x0=y0=0 and
 * full_* == w,h; a shard (output row block, or staged input sub-rectangle) has its own origin. */
typedef struct tfm_image {
    void *data;
    int32_t dtype;         /* tfm_dtype */
    int32_t reserved;
    int64_t h, w;          /* rows, columns held at `data` */
    int64_t pitch;         /* bytes between rows */
    int64_t x0, y0;        /* global origin of element [0][0] */
    int64_t full_h, full_w;/* extent of the full image (boundary conditions act on this) */
    int64_t n_planes;      /* >= 1: leading (broadcast) axes flattened -- image cubes, frame batches
                              (core.py:566-571); src and dst must agree; plane k of dst is resampled
                              from plane k of src with the same chain */
    int64_t plane_pitch;   /* bytes between planes */
} tfm_image;

/* ---- K1: fused inverse-chain evaluation + nearest / linear / cubic / windowed-filter gather ----
 * Replaces, for one call of Transform.resample (core.py:573-584):
 *   output-grid enumeration  np.mgrid[...].transpose()        core.py:574-578   (zero bytes here)
 *   icoords = self.invert(grid)                               core.py:574, 1556-1559
 *   interpND(data, icoords, method, bound)                    core.py:584, helpers.pyx:379-844
 * `chain` (host memory) is the inverse chain in execution order.  dst pixel (x,y) of the shard maps
 * to grid point (dst.x0 + x, dst.y0 + y).  `status_dev` (device int32, may be NULL unless a
 * boundary is TFM_B_FORBID or src is a sub-rectangle) receives OR-ed bits: 1 = forbid violated,
 * 2 = tap outside staged sub-rectangle. */
typedef struct tfm_resample_args {
    int32_t abi_version;   /* TFM_ABI_VERSION */
    int32_t n_ops;
    const tfm_op *chain;   /* host pointer */
    tfm_image src;
    tfm_image dst;
    int32_t method;        /* tfm_method */
    int32_t bound_x, bound_y; /* tfm_bound, (X,Y) order like the reference's per-axis list */
    int32_t precision;     /* tfm_precision */
    int32_t flags;
    int32_t tuning;        /* 0 = auto; otherwise a kernel-variant selector (bench / ncu only) */
    double fillvalue;      /* helpers.pyx:373-374 */
    double oblur;          /* helpers.pyx:509-516; 0 = None.  Only with the filter methods */
    int32_t *status_dev;
} tfm_resample_args;

TFM_API int tfm_resample(const tfm_resample_args *args, void *stream);

/* ---- K2: Jacobian-adaptive (DeForest 2004) anti-aliased resampler ----
 * Replaces interpND_grid(..., antialias=True) (helpers.pyx:1293-1530):
 *   jacobian(index, jump_detect)  = simple_jacobian + jump_detect + grid-centring
 *                                   (helpers.pyx:846-958, 960-1141, 1145-1290)
 *   interpND_jacobian(...)        = per-pixel svd2x2_fast, singular-value padding, footprint loop
 *                                   (helpers.pyx:1542-1803, 1845-1916)
 * with the coordinates produced in-kernel from `chain` instead of read from an index array, and
 * with the repair list of oracle/oracle_np.py::interp_grid_jacobian.  grid_h/grid_w (via dst.full_*)
 * give the FULL output grid so that the reference's grid-edge rules fire on true edges only when
 * the output is sharded by rows. */
typedef struct tfm_jacobian_args {
    int32_t abi_version;
    int32_t n_ops;
    const tfm_op *chain;   /* host pointer */
    tfm_image src;
    tfm_image dst;
    int32_t filter;        /* tfm_filter */
    int32_t bound_x, bound_y;
    int32_t precision;
    int32_t flags;
    int32_t tuning;
    double fillvalue;      /* accepted, unused -- as in the reference (helpers.pyx:1547) */
    double oblur;          /* helpers.pyx:1417-1422 */
    double iblur;          /* accepted, unused -- as in the reference (helpers.pyx:1549) */
    double pad_pow;        /* helpers.pyx:1431-1439, applied as coded at 1661-1662 */
    double jump_thresh;    /* interpND_grid's jump_detect; 0 disables (helpers.pyx:1447-1459) */
    int64_t max_reg;       /* cap on the footprint size reg_size (0 = none); stands in for the
                              reference's accepted-but-unused sv_limit (helpers.pyx:1441-1445) */
    int32_t *status_dev;
} tfm_jacobian_args;

TFM_API int tfm_resample_jacobian(const tfm_jacobian_args *args, void *stream);

/* ---- `tuning`: kernel-variant selectors for A/B measurements (bench.py breakdown, ncu captures, parity tests
 * that pin one variant against another).  0 = the library's own choice; every bit only ever selects another
 * CORRECT variant of the same computation.  Values are stable within an ABI version.
 *   tfm_resample (K1)
 *     1..3      warp patch of the load/store kernel (2: 16 x 2, 3: 8 x 4)
 *     4         do not pre-compose affine chains
 *     16        no texture-gather variants (plain loads)
 *     64        2 x 2 windowed filters through the general filter kernel
 *     128       texture gather, first form (4 pixels 8 rows apart)
 *     256       TMA-staged, one box per CTA
 *     8192      force the persistent TMA pipeline whatever the coverage of the map
 *     1048576   batches of planes: one launch per plane instead of the plane-batch kernel
 *   tfm_resample_jacobian (K2)
 *     8, 32     tile height 16 / 8 of the general kernel
 *     512       general (finite-difference) kernel even where a closed-form kernel applies
 *     1024, 2048, 16384   warp patch widths
 *     4096, 32768         4-byte / 8-byte tap loads instead of texture gathers
 *     65536, 131072       lane-to-pixel maps of a texture request group
 *     262144    analytic polar: per-pixel closed form instead of per-row tables
 *     524288    analytic polar with per-row tables: one pixel per set of fetches instead of row pairs
 *     2097152   parity mode: the general tap loop only (no interior loop, no all-out shortcut)
 *     4194304   parity mode: one exp per tap instead of the fp64 weight recurrence */

/* ---- K3: streaming Transform.apply over vector arrays ----
 * Replaces Transform.apply / invert on [..., vec_len] arrays (core.py:218-315) including the
 * pass-through of components beyond the transform's dimension (core.py:271-276, 287-292).
 * in/out: n_vec rows of vec_len elements of `dtype`, densely packed; may alias (in-place). */
typedef struct tfm_apply_args {
    int32_t abi_version;
    int32_t n_ops;
    const tfm_op *chain;   /* host pointer, execution order */
    const void *in;
    void *out;
    int64_t n_vec;
    int32_t vec_len;       /* >= 2; components 2.. are copied through */
    int32_t dtype;         /* tfm_dtype of in and out (arithmetic is fp64 either way) */
    int32_t precision;
    int32_t tuning;
} tfm_apply_args;

TFM_API int tfm_apply(const tfm_apply_args *args, void *stream);

/* ---- K3-ND: Transform.apply for Linear-family chains of dimension 3..8 ----
 * Replaces Linear._forward / _reverse (basic.py:141-176) under Transform.apply (core.py:255-292) for the
 * transforms the reference builds in more than two dimensions: Linear with an N x N matrix, Scale, Offset
 * and Rotation from axis pairs or Euler angles (basic.py:338-506; tests/test_02_basic.py:170-186).
 * One op = (vector + add_before) -> matrix . vector -> (+ add_after); the host resolves the direction
 * (forward: +pre, matrix, +post; reverse: -post, matinv, -pre -- x - p is x + (-p) exactly).
 * The matrix product is rounded as NumPy rounds np.matmul(m, v[..., None]) for a [dim, dim] matrix and
 * [..., dim, 1] stacks (NumPy 2.3 / OpenBLAS 0.3.30, x86-64): pinned bit for bit for dim = 3 in both memory
 * orders and for dim = 4 transposed views; dim >= 4 otherwise holds 4 ulp of sum |m_ik v_k|.
 * Components dim.. of each vector are copied through (core.py:271-276, 287-292). */
#define TFM_LIN_ND_MAX 8
typedef struct tfm_linear_nd_op {
    int32_t flags;                                   /* TFM_LIN_HAS_PRE (add_before), TFM_LIN_HAS_POST (add_after),
                                                        TFM_LIN_HAS_MATRIX, TFM_LIN_MAT_FORDER */
    int32_t reserved;
    double matrix[TFM_LIN_ND_MAX * TFM_LIN_ND_MAX];  /* row-major [dim][dim] */
    double add_before[TFM_LIN_ND_MAX];
    double add_after[TFM_LIN_ND_MAX];
} tfm_linear_nd_op;

typedef struct tfm_apply_nd_args {
    int32_t abi_version;
    int32_t n_ops;                                   /* <= TFM_MAX_OPS */
    const tfm_linear_nd_op *chain;                   /* host pointer, execution order */
    int32_t dim;                                     /* 1..TFM_LIN_ND_MAX, the same for every op */
    int32_t vec_len;                                 /* >= dim; components dim.. are copied through */
    const void *in;
    void *out;
    int64_t n_vec;
    int32_t dtype;                                   /* tfm_dtype of in and out (arithmetic is fp64) */
    int32_t reserved;
} tfm_apply_nd_args;

TFM_API int tfm_apply_linear_nd(const tfm_apply_nd_args *args, void *stream);

/* ---- K0: FITS pixel decode (the step before the path: file -> device image) ----
 * Replaces what astropy.io.fits does for the reference when it opens the file (core.py:991-1023,
 * DataWrapper on a file name / HDU): big-endian BITPIX samples -> native floating point, with
 * physical = BZERO + BSCALE * stored when either keyword is non-trivial.
 * raw: device pointer to the data unit as it sits in the file (16-byte aligned), n samples.
 * out: TFM_F32 for BITPIX = -32 (a byte swap; scaling in float32), TFM_F64 for every BITPIX
 *      (integers convert exactly; scaling is one rounded multiply and one rounded add). */
TFM_API int tfm_fits_decode(const void *raw, int32_t bitpix, int64_t n, double bscale, double bzero,
                            int32_t out_dtype, void *out, void *stream);

/* ---- the Jacobian helpers the reference exposes as module functions (2-D case) ----
 * All pointers are device pointers, arrays dense and row-major, float64 (flags int64 like NumPy's int).
 * tfm_simple_jacobian  replaces simple_jacobian(index)            helpers.pyx:846-958 (2-D: 896-906)
 *     index[ny][nx][2] -> Js[ny-1][nx-1][2][2],  Js[...,i,j] = d coord_i / d pix_j
 * tfm_jump_detect      replaces jump_detect(Js, jump_thresh)      helpers.pyx:960-1141 (N-D branch, N=2)
 *     Js[ncy][ncx][2][2] -> flags[ncy][ncx] (1 = ok, 0 = jump); scratch: ncy*ncx doubles.
 *     Repair R12: upper-edge results land on the true upper edge.
 * tfm_jacobian_grid    replaces the 2-D branch of jacobian(index) helpers.pyx:1225-1247, with the
 *     repairs that make it return (R1, R2, R11): cells Js[ny-1][nx-1][2][2] and flags (or NULL for
 *     no jump detection) -> grid-centred J[ny][nx][2][2]; edge rows/columns copy their neighbour. */
TFM_API int tfm_simple_jacobian(const double *index, int64_t ny, int64_t nx, double *Js, void *stream);
TFM_API int tfm_jump_detect(const double *Js, int64_t ncy, int64_t ncx, double jump_thresh, int64_t *flags,
                    double *scratch, void *stream);
TFM_API int tfm_jacobian_grid(const double *Js, const int64_t *flags, int64_t ny, int64_t nx, double *J,
                      void *stream);

/* ---- host-side test hooks mirroring helpers.pyx:1807-1810 (svd2x2 / svc2x2) ----
 * M, U, V: 4 doubles row-major; s: 2 doubles.  Run the device routines on one matrix on the
 * current device and copy the result back (used by parity tests only). */
TFM_API int tfm_svd2x2(const double *M_host, double *U_host, double *s_host, double *V_host);
TFM_API int tfm_svc2x2(const double *U_host, const double *s_host, const double *V_host, double *M_host);

/* ---- misc ---- */
TFM_API int tfm_abi_version(void);
TFM_API const char *tfm_strerror(int status);
/* Number of kernel launches issued by this library in this process (bench.py's gpu_launches). */
TFM_API int64_t tfm_launch_count(void);
/* Name of the kernel variant chosen by the last tfm_resample / tfm_resample_jacobian / tfm_apply
 * call on this thread (for bench.py / profiles). */
TFM_API const char *tfm_last_kernel(void);

#ifdef __cplusplus
}
#endif
#endif /* TFM_B200_H */
