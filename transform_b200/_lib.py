"""ctypes binding of libtfm_b200.so (include/tfm_b200.h).

The product path has NO CPU fallback: if the shared library is missing, or there is no CUDA
device, every compute entry point raises -- loudly -- instead of computing something else.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TFM_B200_LIB selects another build of the SAME library (A/B kernel experiments, tools/ab_variants.py)
LIB_PATH = os.environ.get("TFM_B200_LIB") or os.path.join(_HERE, "lib", "libtfm_b200.so")

ABI_VERSION = 2
MAX_OPS = 16
OP_NPARAM = 24

# status codes
OK, EVALUE, EDIRECTION, EFORBID, ECUDA, EUNSUPPORTED, ESTAGED = range(7)
# op kinds / directions
OP_IDENTITY, OP_LINEAR, OP_RADIAL, OP_WCS, OP_INDEX, OP_PIN, OP_SIP, OP_TPV = range(8)
SIP_MAX_ORDER, SIP_DIM, SIP_MAXITER = 9, 10, 20
TPV_NCOEF, TPV_MAXITER = 40, 30
PIN_QUAD, PIN_CUBIC, PIN_POLY, PIN_SHIFTED, PIN_DIM1, PIN_TRUNC, PIN_MAX_DEGREE = 0, 1, 2, 4, 8, 16, 8
FWD, REV = 0, 1
# flags
LIN_HAS_PRE, LIN_HAS_POST, LIN_HAS_MATRIX, LIN_MAT_FORDER = 1, 2, 4, 8
RAD_CCW, RAD_POS_ONLY, RAD_R0_NOT_NONE, RAD_R0_TRUTHY, RAD_ORIGIN_NONZERO = 1, 2, 4, 8, 16
WCS_PROJ_NONE, WCS_PROJ_TAN, WCS_PROJ_CAR = 0, 1, 2
(WCS_PROJ_STG, WCS_PROJ_SIN, WCS_PROJ_ARC, WCS_PROJ_ZEA, WCS_PROJ_CEA, WCS_PROJ_MER, WCS_PROJ_SFL,
 WCS_PROJ_AIT) = range(3, 11)
F32, F64 = 0, 1
NEAREST, LINEAR, CUBIC = 0, 1, 2
SINC, LANCZOS, GAUSSIAN, HANN, TUKEY, TENT = 3, 4, 5, 6, 7, 8
FILTER_MAX_SIZE = 112
BOUND_CODE = {"f": 1, "t": 2, "e": 3, "p": 4, "m": 5}            # helpers.pyx:1511
FILTER_CODE = {"l": 1, "z": 2, "h": 3, "t": 4, "g": 5,           # helpers.pyx:1501
               "w": 6}                                           # Hann window of SVD.pyx:110-111 (TFM_FILT_HANN_WINDOW)
PREC_EXACT64, PREC_FAST32 = 0, 1
F_REGION_SRC_DTYPE = 1
F_FLUX = 2


class TfmOp(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("dir", ctypes.c_int32), ("flags", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("p", ctypes.c_double * OP_NPARAM)]


class TfmImage(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("dtype", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("h", ctypes.c_int64), ("w", ctypes.c_int64), ("pitch", ctypes.c_int64),
                ("x0", ctypes.c_int64), ("y0", ctypes.c_int64),
                ("full_h", ctypes.c_int64), ("full_w", ctypes.c_int64),
                ("n_planes", ctypes.c_int64), ("plane_pitch", ctypes.c_int64)]


class TfmResampleArgs(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int32), ("n_ops", ctypes.c_int32),
                ("chain", ctypes.POINTER(TfmOp)),
                ("src", TfmImage), ("dst", TfmImage),
                ("method", ctypes.c_int32), ("bound_x", ctypes.c_int32), ("bound_y", ctypes.c_int32),
                ("precision", ctypes.c_int32), ("flags", ctypes.c_int32), ("tuning", ctypes.c_int32),
                ("fillvalue", ctypes.c_double), ("oblur", ctypes.c_double), ("status_dev", ctypes.c_void_p)]


class TfmJacobianArgs(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int32), ("n_ops", ctypes.c_int32),
                ("chain", ctypes.POINTER(TfmOp)),
                ("src", TfmImage), ("dst", TfmImage),
                ("filter", ctypes.c_int32), ("bound_x", ctypes.c_int32), ("bound_y", ctypes.c_int32),
                ("precision", ctypes.c_int32), ("flags", ctypes.c_int32), ("tuning", ctypes.c_int32),
                ("fillvalue", ctypes.c_double), ("oblur", ctypes.c_double), ("iblur", ctypes.c_double),
                ("pad_pow", ctypes.c_double), ("jump_thresh", ctypes.c_double),
                ("max_reg", ctypes.c_int64), ("status_dev", ctypes.c_void_p)]


class TfmApplyArgs(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int32), ("n_ops", ctypes.c_int32),
                ("chain", ctypes.POINTER(TfmOp)),
                ("inp", ctypes.c_void_p), ("out", ctypes.c_void_p),
                ("n_vec", ctypes.c_int64), ("vec_len", ctypes.c_int32), ("dtype", ctypes.c_int32),
                ("precision", ctypes.c_int32), ("tuning", ctypes.c_int32)]


LIN_ND_MAX = 8


class TfmLinearNdOp(ctypes.Structure):
    """One Linear-family op of dimension 1..8 (include/tfm_b200.h: tfm_linear_nd_op)."""
    _fields_ = [("flags", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("matrix", ctypes.c_double * (LIN_ND_MAX * LIN_ND_MAX)),
                ("add_before", ctypes.c_double * LIN_ND_MAX), ("add_after", ctypes.c_double * LIN_ND_MAX)]


class TfmApplyNdArgs(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int32), ("n_ops", ctypes.c_int32),
                ("chain", ctypes.POINTER(TfmLinearNdOp)),
                ("dim", ctypes.c_int32), ("vec_len", ctypes.c_int32),
                ("inp", ctypes.c_void_p), ("out", ctypes.c_void_p),
                ("n_vec", ctypes.c_int64), ("dtype", ctypes.c_int32), ("reserved", ctypes.c_int32)]


EXPORTED_SYMBOLS = ("tfm_resample", "tfm_resample_jacobian", "tfm_apply", "tfm_apply_linear_nd", "tfm_svd2x2", "tfm_svc2x2",
                    "tfm_simple_jacobian", "tfm_jump_detect", "tfm_jacobian_grid", "tfm_fits_decode",
                    "tfm_abi_version", "tfm_strerror", "tfm_launch_count", "tfm_last_kernel")

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load():
    """Load libtfm_b200.so.  Raises LibraryMissing (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            "transform_b200: %s is missing -- build it with `python transform_b200/csrc/build.py` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.tfm_abi_version.restype = ctypes.c_int
    lib.tfm_strerror.restype = ctypes.c_char_p
    lib.tfm_strerror.argtypes = [ctypes.c_int]
    lib.tfm_launch_count.restype = ctypes.c_int64
    lib.tfm_last_kernel.restype = ctypes.c_char_p
    lib.tfm_resample.restype = ctypes.c_int
    lib.tfm_resample.argtypes = [ctypes.POINTER(TfmResampleArgs), ctypes.c_void_p]
    lib.tfm_resample_jacobian.restype = ctypes.c_int
    lib.tfm_resample_jacobian.argtypes = [ctypes.POINTER(TfmJacobianArgs), ctypes.c_void_p]
    lib.tfm_apply.restype = ctypes.c_int
    lib.tfm_apply.argtypes = [ctypes.POINTER(TfmApplyArgs), ctypes.c_void_p]
    lib.tfm_apply_linear_nd.restype = ctypes.c_int
    lib.tfm_apply_linear_nd.argtypes = [ctypes.POINTER(TfmApplyNdArgs), ctypes.c_void_p]
    vp, i64 = ctypes.c_void_p, ctypes.c_int64
    lib.tfm_fits_decode.restype = ctypes.c_int
    lib.tfm_fits_decode.argtypes = [vp, ctypes.c_int32, i64, ctypes.c_double, ctypes.c_double, ctypes.c_int32, vp, vp]
    lib.tfm_simple_jacobian.restype = ctypes.c_int
    lib.tfm_simple_jacobian.argtypes = [vp, i64, i64, vp, vp]
    lib.tfm_jump_detect.restype = ctypes.c_int
    lib.tfm_jump_detect.argtypes = [vp, i64, i64, ctypes.c_double, vp, vp, vp]
    lib.tfm_jacobian_grid.restype = ctypes.c_int
    lib.tfm_jacobian_grid.argtypes = [vp, vp, i64, i64, vp, vp]
    lib.tfm_svd2x2.restype = ctypes.c_int
    lib.tfm_svd2x2.argtypes = [dp, dp, dp, dp]
    lib.tfm_svc2x2.restype = ctypes.c_int
    lib.tfm_svc2x2.argtypes = [dp, dp, dp, dp]
    if lib.tfm_abi_version() != ABI_VERSION:
        raise LibraryMissing("transform_b200: ABI mismatch between _lib.py and %s" % LIB_PATH)
    _lib = lib
    return lib


def strerror(status):
    return load().tfm_strerror(int(status)).decode()


def raise_for_status(status, what):
    """Map a tfm_status to the exception the reference raises for the same condition
    (SURVEY.md section 8b "Errors")."""
    if status == OK:
        return
    msg = "%s: %s" % (what, strerror(status))
    if status in (EVALUE, EFORBID):
        raise ValueError(msg)
    if status == EDIRECTION:
        raise AssertionError(msg)
    if status == EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def make_chain(ops):
    """list of TfmOp -> (ctypes array, n)."""
    n = len(ops)
    if n > MAX_OPS:
        raise NotImplementedError("transform_b200: chains longer than %d elements are not supported" % MAX_OPS)
    arr = (TfmOp * max(n, 1))()
    for i, o in enumerate(ops):
        if not isinstance(o, TfmOp):
            raise NotImplementedError("transform_b200: a transform of dimension > 2 inside a 2-D chain")
        arr[i] = o
    return arr, n


def launch_count():
    return int(load().tfm_launch_count())


def last_kernel():
    return load().tfm_last_kernel().decode()
