"""ctypes binding of the C-ABI in include/ganlab_b200.h.

The signatures are parsed from the header itself, so the Python side can never drift from the ABI the
header declares.  There is no fallback: if the shared library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes
import os
import re
from pathlib import Path

_PKG = Path(__file__).resolve().parent
HEADER = _PKG.parent / "include" / "ganlab_b200.h"
LIB_PATH = _PKG / "libganlab_b200.so"

_CTYPES = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "float": ctypes.c_float,
    "glb_stream_t": ctypes.c_void_p,
    "void": None,
}


def parse_header(path: Path = HEADER):
    """-> {name: (restype, [argtypes])} for every function the header declares."""
    text = re.sub(r"/\*.*?\*/", " ", path.read_text(), flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    text = re.sub(r"#[^\n]*", " ", text)
    out = {}
    for m in re.finditer(r"(const\s+char\s*\*|int64_t|int)\s+(glb_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "char" in ret else _CTYPES[ret]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.split()[-2] if len(a.split()) > 1 else a
                    argtypes.append(_CTYPES[ty])
        out[name] = (restype, argtypes)
    return out


class GlbError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        self._dll = None
        self.signatures = parse_header()

    def load(self):
        if self._dll is not None:
            return self._dll
        if not LIB_PATH.exists():
            raise GlbError(
                f"{LIB_PATH} not found: build the sm_100a extension first (python -c 'import __graft_entry__ as g; g.build()'). "
                "gan_lab_b200 has no CPU or PyTorch fallback.")
        dll = ctypes.CDLL(str(LIB_PATH), mode=os.RTLD_NOW | os.RTLD_LOCAL)
        for name, (restype, argtypes) in self.signatures.items():
            fn = getattr(dll, name)            # AttributeError if the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        self._dll = dll
        return dll

    def call(self, name: str, *args):
        fn = getattr(self.load(), name)
        rc = fn(*args)
        if rc != 0:
            msg = self._dll.glb_last_error()
            raise GlbError(f"{name} failed (rc={rc}): {msg.decode() if msg else '?'}")

    def fn(self, name: str):
        return getattr(self.load(), name)


LIB = _Lib()
