"""Import the UNMODIFIED reference package (read-only at /root/reference) in THIS container.

TEST INFRASTRUCTURE ONLY (golden-vector generation and oracle pinning).  Never imported by the
product package; never used on the GPU box (the reference tree does not exist there).

The reference is 2020-era code; the shims below are applied at import time, none of its files
is edited or copied:
  1. ``np.int`` alias (removed in numpy 1.24; used by sample_labels.py:31, case_control_likelihood.py:46,
     hdp_lpcm.py:107, sample_auxillary.py:11, datasets/samples_generator.py).
  2. stub ``statsmodels.regression.linear_model.yule_walker`` (trace_utils.py:6; post-sampling only).
  3. ``sklearn.utils.check_array(force_all_finite=...)`` -> ``ensure_all_finite`` (lsm.py:341, hdp_lpcm.py:663).
  4. the package ``__init__`` is not executed (it star-imports plots' dependencies); a synthetic
     package object with ``__path__ = [reference/dynetlsm, oracle/_ref]`` lets the pure-Python
     modules resolve from the reference tree and the Cython kernels from oracle/_ref.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DYNETLSM_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REF, "dynetlsm"))


def load_reference():
    """Return the ``dynetlsm`` package object backed by the reference sources."""
    if "dynetlsm" in sys.modules and getattr(sys.modules["dynetlsm"], "_b200_shimmed", False):
        return sys.modules["dynetlsm"]
    if not reference_available():
        raise RuntimeError("reference sources not present at %s" % REF)
    sys.path.insert(0, HERE)
    import build_ref
    if not build_ref.build():
        raise RuntimeError("could not build oracle/_ref")

    import numpy as np
    if not hasattr(np, "int"):
        np.int = int  # shim 1

    # shim 2
    if "statsmodels" not in sys.modules:
        sm = types.ModuleType("statsmodels")
        smr = types.ModuleType("statsmodels.regression")
        sml = types.ModuleType("statsmodels.regression.linear_model")

        def yule_walker(*a, **k):
            raise NotImplementedError("statsmodels is not installed (stub)")
        sml.yule_walker = yule_walker
        sm.regression = smr
        smr.linear_model = sml
        sys.modules["statsmodels"] = sm
        sys.modules["statsmodels.regression"] = smr
        sys.modules["statsmodels.regression.linear_model"] = sml

    # shim 3
    import sklearn.utils
    import sklearn.utils.validation as skv
    if not getattr(skv.check_array, "_b200_wrapped", False):
        _orig = skv.check_array

        def check_array(*a, **k):
            if "force_all_finite" in k:
                k["ensure_all_finite"] = k.pop("force_all_finite")
            return _orig(*a, **k)
        check_array._b200_wrapped = True
        skv.check_array = check_array
        sklearn.utils.check_array = check_array

    # shim 4
    pkg = types.ModuleType("dynetlsm")
    pkg.__path__ = [os.path.join(REF, "dynetlsm"), os.path.join(HERE, "_ref")]
    pkg._b200_shimmed = True
    sys.modules["dynetlsm"] = pkg
    import warnings
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    warnings.filterwarnings("ignore", category=FutureWarning)
    import dynetlsm.lsm  # noqa: F401
    import dynetlsm.hdp_lpcm  # noqa: F401
    import dynetlsm.datasets  # noqa: F401
    # geweke needs statsmodels; make the post-sampling diagnostic a no-op
    import dynetlsm.trace_utils as tu
    tu.geweke_diag = lambda *a, **k: float("nan")
    dynetlsm.hdp_lpcm.geweke_diag = tu.geweke_diag
    return pkg
