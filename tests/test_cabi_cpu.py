"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports exactly
what include/dlsm.h declares; without a GPU it fails loudly instead of falling back."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from dynetlsm_b200 import _lib
    _lib.build()
    return _lib.load()


def test_header_and_exports_agree(lib):
    from dynetlsm_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dlsm.h")).read()
    declared = set(re.findall(r"\b(dlsm_[a-z_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.dlsm_abi_version() == 2


def test_every_entry_point_cites_the_reference():
    hdr = open(os.path.join(ROOT, "include", "dlsm.h")).read()
    for ref in ("sample_latent_positions.py:92", "sample_coefficients.py:12", "sample_coefficients.py:91",
                "sample_labels.py:134", "case_control_likelihood.py:37", "lsm.py:501",
                "directed_likelihoods_fast.pyx:185", "metropolis.py:40"):
        assert ref in hdr, ref


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dynetlsm_b200 import _lib
    assert lib.dlsm_device_count() == 0
    with pytest.raises(_lib.DlsmError) as ei:
        _lib.Engine(T=2, n=5, d=2)
    assert ei.value.code == -2
    assert "no CPU fallback" in str(ei.value)


def test_invalid_configs_are_rejected(lib):
    from dynetlsm_b200 import _lib
    with pytest.raises(_lib.DlsmError) as ei:
        _lib.Engine(T=2, n=5, d=2, case_control=True, is_directed=False)
    assert ei.value.code == -1  # lsm.py:425-427
    assert "only supported for directed" in str(ei.value)
    with pytest.raises(_lib.DlsmError):
        _lib.Engine(T=2, n=5, d=9)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "dynetlsm_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in src and "liboracle" not in src and "oracle/" not in src, f


def test_field_ids_match_the_header():
    """The Python field constants mirror include/dlsm.h's dlsm_field enum one to one."""
    import re
    from dynetlsm_b200 import _lib
    src = open(os.path.join(ROOT, "include", "dlsm.h")).read()
    enum = src[src.index("DLSM_F_X = 0"):src.index("DLSM_F_COUNT_")]
    ids = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"DLSM_F_(\w+)\s*=\s*(\d+)", enum))
    assert sorted(ids.values()) == list(range(len(ids)))
    assert _lib.N_FIELDS == len(ids)
    for name, val in ids.items():
        assert getattr(_lib, "F_" + name) == val, name
