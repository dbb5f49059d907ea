"""ctypes binding of libaccx.so (the C ABI declared in include/accx.h).

The prototypes are read from the header itself, so the binding cannot drift from the ABI and
tests can check that every declared symbol is exported.  There is NO fallback: if the shared
library is missing or a CUDA device is absent, the ops raise -- nothing in the product path
computes on the CPU.
"""
from __future__ import annotations

import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
HEADER = os.path.join(ROOT, "include", "accx.h")
LIB_PATH = os.path.join(HERE, "libaccx.so")

F32, BF16 = 0, 1
MAX_OPERANDS, MAX_ADDENDS = 9, 4


class Operand(ctypes.Structure):
    """mirror of accx_operand_t"""
    _fields_ = [
        ("data", ctypes.c_void_p), ("ld", ctypes.c_int64), ("K", ctypes.c_int32), ("act", ctypes.c_int32),
        ("scale", ctypes.c_void_p), ("shift", ctypes.c_void_p), ("w", ctypes.c_void_p),
        ("w_ld", ctypes.c_int64), ("w_ks", ctypes.c_int64), ("dy", ctypes.c_int32), ("dx", ctypes.c_int32),
    ]


_SCALARS = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float, "double": ctypes.c_double,
            "unsigned int": ctypes.c_uint}


def parse_header(path: str = HEADER):
    """-> {name: (restype, [argtype, ...], [argname, ...])} for every accx_* prototype."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|int64_t|int)\s+(accx_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        types, names = [], []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                ty, nm = mm.group(1).strip(), mm.group(2)
                names.append(nm)
                if "*" in ty:
                    types.append(ctypes.POINTER(Operand) if "accx_operand_t" in ty else ctypes.c_void_p)
                else:
                    types.append(_SCALARS[ty.replace("const ", "")])
        rt = ctypes.c_char_p if "char" in ret else (ctypes.c_int64 if ret == "int64_t" else ctypes.c_int)
        protos[name] = (rt, types, names)
    return protos


class AccxError(RuntimeError):
    pass


_lib = None
_protos = None


def load():
    """Load libaccx.so and attach prototypes.  Raises (never falls back) when it is absent."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AccxError(f"{LIB_PATH} not found: build it with `make -C acc-unet-unext_b200/csrc` "
                        "(or __graft_entry__.build()); accx has no CPU/eager fallback")
    lib = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    for name, (ret, types, _) in _protos.items():
        fn = getattr(lib, name)
        fn.restype = ret
        fn.argtypes = types
    _lib = lib
    # whole-step experiments: ACCX_KNOBS="index=value,index=value" (see KNOB_* in csrc/common.cuh)
    for kv in filter(None, os.environ.get("ACCX_KNOBS", "").split(",")):
        k, v = kv.split("=")
        lib.accx_set_knob(int(k), int(v))
    return lib


def call(name: str, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise AccxError(f"{name} failed ({rc}): {lib.accx_last_error().decode()}")
