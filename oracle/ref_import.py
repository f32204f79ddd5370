"""oracle/ref_import.py -- TEST INFRASTRUCTURE ONLY; works only where /root/reference exists.

Imports the UNMODIFIED reference package (`transform`) from /root/reference in the
authoring container so that golden vectors can be generated from the real thing
(tests/golden/make_golden.py) and the numpy/C restatement in this directory can be pinned
against it.  Nothing here is reachable from transform_b200, bench.py's product arm, or the
`-m gpu` tests (the GPU box has no /root/reference).

How: (1) oracle/build_ref.py compiles helpers.pyx / SVD.pyx where they lie into
oracle/_ref/*.so; (2) that helpers module is registered as `transform.helpers` before the
package is imported (the package directory is read-only, so the .so cannot sit next to
core.py); (3) oracle/astropy_stub provides the `astropy.units` surface core.py/basic.py
need at import time (astropy itself is not installed).
"""
import importlib
import importlib.machinery
import importlib.util
import os
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TFM_REFERENCE_ROOT", "/root/reference")
_REFDIR = os.path.join(HERE, "_ref")


def have_reference_sources():
    return os.path.exists(os.path.join(REF, "transform", "core.py"))


def _load_ext(modname, fullname):
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    path = os.path.join(_REFDIR, modname + suffix)
    if not os.path.exists(path) and modname in ("helpers", "SVD", "core", "basic") and have_reference_sources():
        import build_ref                       # a fresh checkout in the authoring container: build on first use
        build_ref.build()
    if not os.path.exists(path):
        raise ImportError("oracle/_ref/%s%s not built (run oracle/build_ref.py)" % (modname, suffix))
    loader = importlib.machinery.ExtensionFileLoader(modname, path)
    spec = importlib.util.spec_from_file_location(modname, path, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules[fullname] = mod
    return mod


def load_ref_helpers():
    """The reference's compiled helpers module alone (travels to the GPU box as a .so)."""
    if "tfm_ref_helpers" in sys.modules:
        return sys.modules["tfm_ref_helpers"]
    return _load_ext("helpers", "tfm_ref_helpers")


def load_ref_helpers_repaired():
    """The reference's helpers module built with the named repairs of oracle/build_ref_repaired.py:
    its anti-aliased path (interpND_grid -> jacobian -> interpND_jacobian) runs."""
    if "tfm_ref_helpers_repaired" in sys.modules:
        return sys.modules["tfm_ref_helpers_repaired"]
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    if not os.path.exists(os.path.join(_REFDIR, "helpers_repaired" + suffix)):
        import build_ref_repaired
        if not build_ref_repaired.build():
            raise ImportError("oracle/_ref/helpers_repaired%s not built and the reference is absent" % suffix)
    return _load_ext("helpers_repaired", "tfm_ref_helpers_repaired")


def load_ref_svd():
    if "tfm_ref_SVD" in sys.modules:
        return sys.modules["tfm_ref_SVD"]
    return _load_ext("SVD", "tfm_ref_SVD")


def load_reference_compiled():
    """The reference's own classes (core.py, basic.py) and helpers, all as compiled by oracle/build_ref.py into
    oracle/_ref -- importable on the GPU box, where /root/reference does not exist.  Returns a module-like
    namespace `transform` with the reference's public names (Composition, Scale, Rotation, Offset, Radial, ...)."""
    if "transform" in sys.modules and getattr(sys.modules["transform"], "_tfm_is_reference", False):
        return sys.modules["transform"]
    import types
    try:
        import astropy  # noqa: F401
    except ImportError:
        sys.path.insert(0, os.path.join(HERE, "astropy_stub"))
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    if have_reference_sources() and not all(os.path.exists(os.path.join(_REFDIR, nm + suffix)) for nm in ("helpers", "core", "basic")):
        import build_ref
        build_ref.build()
    for name in ("helpers", "core", "basic"):
        if not os.path.exists(os.path.join(_REFDIR, name + suffix)):
            raise ImportError("oracle/_ref/%s%s not built (run oracle/build_ref.py where the reference is present)" % (name, suffix))
    pkg = types.ModuleType("transform")
    pkg.__path__ = []                          # a package with no directory: its members are registered below
    pkg.__package__ = "transform"
    saved = {k: sys.modules.get(k) for k in ("transform", "transform.helpers", "transform.core", "transform.basic")}
    try:
        sys.modules["transform"] = pkg
        helpers = load_ref_helpers()
        sys.modules["transform.helpers"] = helpers
        pkg.helpers = helpers
        for name in ("core", "basic"):
            full = "transform." + name
            path = os.path.join(_REFDIR, name + suffix)
            loader = importlib.machinery.ExtensionFileLoader(full, path)
            spec = importlib.util.spec_from_file_location(full, path, loader=loader)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            loader.exec_module(mod)
            setattr(pkg, name, mod)
            for k, v in vars(mod).items():
                if not k.startswith("_"):
                    setattr(pkg, k, v)
        for k, v in vars(helpers).items():
            if not k.startswith("_") and not hasattr(pkg, k):
                setattr(pkg, k, v)
    except Exception:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        raise
    pkg._tfm_is_reference = True
    pkg._tfm_compiled = True
    return pkg


def load_reference():
    """Import the unmodified reference package; container only."""
    if not have_reference_sources():
        raise ImportError("reference sources not present at %s" % REF)
    if "transform" in sys.modules and getattr(sys.modules["transform"], "_tfm_is_reference", False):
        return sys.modules["transform"]
    try:
        import astropy  # noqa: F401  (a real astropy, if one ever appears, wins)
    except ImportError:
        sys.path.insert(0, os.path.join(HERE, "astropy_stub"))
    helpers = load_ref_helpers()
    sys.modules["transform.helpers"] = helpers
    sys.path.insert(0, REF)
    try:
        mod = importlib.import_module("transform")
    finally:
        sys.path.remove(REF)
    mod._tfm_is_reference = True
    return mod
