"""ctypes binding of liblpi_b200.so (the C ABI declared in include/lpi_b200.h).

There is deliberately no fallback: if the shared library is missing or the device is not sm_100a
every op raises.  Build it with `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LPI_LIB_PATH") or os.path.join(_HERE, "liblpi_b200.so")      # override: A/B runs of two builds on one box
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "lpi_b200.h")

_lib = None


class LpiError(RuntimeError):
    pass


def declared_symbols(header_path: str = HEADER_PATH):
    """Every `lpi_*` function the public header declares (used by the CPU-side ABI test)."""
    with open(header_path) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(lpi_[a-z0-9_]+)\s*\(", src)))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise LpiError(f"{LIB_PATH} not built -- run __graft_entry__.build(); lpi_b200 has no CPU / eager fallback")
        l = C.CDLL(LIB_PATH)
        l.lpi_last_error.restype = C.c_char_p
        for name in declared_symbols():
            fn = getattr(l, name, None)
            if fn is None:
                raise LpiError(f"{LIB_PATH} does not export {name}; rebuild it")
            if name != "lpi_last_error":
                fn.restype = C.c_int
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise LpiError(f"{what or 'lpi'} failed ({rc}): {lib().lpi_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def call(name: str, *args):
    """Invoke `lpi_<name>` with ctypes-converted args and raise on a non-zero status."""
    fn = getattr(lib(), "lpi_" + name)
    check(fn(*args), "lpi_" + name)


_device_ok = False


def require_device():
    """Fail loudly unless a B200-class (sm_100) device is current."""
    global _device_ok
    if _device_ok:
        return
    import torch

    if not torch.cuda.is_available():
        raise LpiError("lpi_b200 needs a CUDA device (sm_100a); there is no CPU path")
    sms, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    check(lib().lpi_device_check(C.byref(sms), C.byref(maj), C.byref(mnr)), "lpi_device_check")
    _device_ok = True
