"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/lpi_b200.h declares (no compute calls without a GPU), and the product path fails loudly
without a device instead of falling back."""
import ctypes
import os

import pytest
import torch

from lpi_b200 import _lib


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as G

    G.build()
    names = _lib.declared_symbols()
    assert len(names) >= 10 and "lpi_sim_topk_bf16" in names and "lpi_gemm_bf16" in names
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"liblpi_b200.so does not export {n}"
    assert lib.lpi_version() >= 100


def test_product_never_imports_the_oracle():
    """Only tests/, smoke and bench.py's cpu_baseline leg may touch oracle/ (it is the checker)."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for dirpath, _, files in os.walk(os.path.join(root, "lpi_b200")):
        for f in files:
            if f.endswith(".py") and f != "smoke.py":     # smoke.py is the driver's checker entry
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f"{f} imports the oracle"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from lpi_b200 import ops, retrieval as R

    x = torch.randn(4, 512)
    with pytest.raises(_lib.LpiError):
        ops.l2_normalize(x)
    with pytest.raises(_lib.LpiError):
        ops.sim_topk(x.bfloat16(), x.bfloat16(), 2)
    with pytest.raises((_lib.LpiError, RuntimeError, AssertionError)):
        R.itm_eval(x.numpy(), x.numpy().T.copy(), {i: i for i in range(4)}, {i: [i] for i in range(4)}, [0] * 4, [0] * 4, 1)
