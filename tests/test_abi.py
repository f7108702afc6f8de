"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declare; compute entry points fail loudly
without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from deepreinforcementlearningcontrolofquantumcartpoles_b200 import _lib as L, make_config, configs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for header in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not header.endswith(".h"):
            continue
        txt = open(os.path.join(ROOT, "include", header)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names.update(re.findall(r"\b(qc_[a-z0-9_]+)\s*\(", txt))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(n for n, _, _ in L.SYMBOLS) == names


def test_version_and_philox_host_function():
    lib = L.load()
    assert lib.qc_version().decode().startswith("qcart")


def test_bad_config_is_rejected_before_touching_the_device():
    import torch
    lib = L.load()
    cfg = make_config(configs.quartic())
    cfg.n_levels = 20                      # must be odd
    h = C.c_void_p()
    rc = lib.qc_create(C.byref(cfg), C.byref(h))
    assert rc != 0
    if not torch.cuda.is_available():
        assert rc == L.QC_ERR_CUDA         # device check comes first on a CPU box: there is no CPU fallback
        assert b"no usable CUDA device" in lib.qc_last_error()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-box check")
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import BatchedSim, QcartError
    with pytest.raises(QcartError):
        BatchedSim(configs.quartic(), batch=4)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deepreinforcementlearningcontrolofquantumcartpoles_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), os.path.join(dirpath, f)
