"""ctypes binding of libspgnn_b200.so — the reference-side stub of INTEGRATION.md, as shipped.

The prototypes are read from ``include/spgnn_b200.h`` so the binding cannot drift from the header.  There is no
fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPGNN_B200_LIB") or os.path.join(_HERE, "libspgnn_b200.so")   # override: A/B of two builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "spgnn_b200.h")

_CT = {
    "int": ctypes.c_int, "float": ctypes.c_float, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64,
    "uint32_t": ctypes.c_uint32, "int32_t": ctypes.c_int32, "double": ctypes.c_double,
}


def parse_header(path=HEADER_PATH):
    """[(name, restype, [argtypes])] for every prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = []
    for m in re.finditer(r"(?:^|\n)\s*(const\s+char\s*\*|int64_t|int|void)\s+(spgnn_\w+)\s*\(([^;{]*?)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = ctypes.c_char_p if "char" in ret else (None if ret == "void" else _CT[ret])
        argtypes = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.replace("const", "").split()[0]
                    argtypes.append(_CT[ty])
        protos.append((name, restype, argtypes))
    return protos


class SpgnnError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise SpgnnError(
                f"{LIB_PATH} not found — build it with `python -m spgnn_b200.build` (there is no CPU fallback)")
        self._dll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, restype, argtypes in self.protos:
            fn = getattr(self._dll, name)          # AttributeError if the header declares something not exported
            fn.restype = restype
            fn.argtypes = argtypes
        self._status = {n for n, r, _ in self.protos if r is ctypes.c_int and n not in
                        ("spgnn_abi_version", "spgnn_seed_salt_units")}
        self.profile = None          # when a list: (name, key, start_event, stop_event) per C-ABI call

    def last_error(self):
        return self._dll.spgnn_last_error().decode()

    def __getattr__(self, name):
        fn = getattr(self._dll, "spgnn_" + name)
        if "spgnn_" + name not in self._status:
            return fn

        def call(*args, _key=None, _name=None):
            prof = self.profile
            if prof is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            rc = fn(*args)
            if rc != 0:
                raise SpgnnError(f"spgnn_{name} failed ({rc}): {self.last_error()}")
            if prof is not None:
                e1.record()
                prof.append((_name or name, _key, e0, e1))
        return call


_lib = None


def lib() -> _Lib:
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def ptr(t):
    """Device pointer of a tensor (None → NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SpgnnError("spgnn_b200 runs on CUDA tensors only (no CPU fallback); got a tensor on " + str(t.device))
