"""ctypes binding of the C-ABI library (include/xmodal_b200.h).

The signatures are read from the header itself, so the header stays the single source of truth
for the boundary.  There is no CPU fallback: if the library is missing `lib()` raises, and every
call that returns a non-zero status raises `XmodalError`.
"""
from __future__ import annotations

import ctypes
import re
from functools import lru_cache
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
HEADER = PKG_DIR.parent / "include" / "xmodal_b200.h"
LIB_PATH = PKG_DIR / "_C" / "libxmodal_b200.so"

_CTYPES = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "uint64_t": ctypes.c_uint64,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
}


class XmodalError(RuntimeError):
    """status: the XM_ERR_* code of the failed call (-2 = XM_ERR_UNSUPPORTED), None when raised by the Python layer."""

    def __init__(self, msg, status=None):
        super().__init__(msg)
        self.status = status


def _parse_type(tok: str):
    tok = tok.strip()
    if "*" in tok:
        return ctypes.c_char_p if tok.replace(" ", "") == "constchar*" else ctypes.c_void_p
    tok = tok.replace("const", "").strip()
    return _CTYPES[tok]


@lru_cache(maxsize=1)
def declared_functions():
    """{name: (restype, [argtypes], [argnames])} parsed from the header."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(xm_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                argtypes.append(_parse_type(mm.group(1)))
                argnames.append(mm.group(2))
        out[name] = (_parse_type(ret), argtypes, argnames)
    return out


@lru_cache(maxsize=1)
def lib() -> ctypes.CDLL:
    if not LIB_PATH.exists():
        raise XmodalError(
            f"{LIB_PATH} is missing: build it with `python -m multimodal_eeg_fmri_b200.build` "
            "(there is no CPU fallback for this path)"
        )
    dll = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes, _) in declared_functions().items():
        fn = getattr(dll, name)  # AttributeError here == header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    return dll


def check(status: int, what: str) -> None:
    if status != 0:
        dll = lib()
        msg = dll.xm_strerror(status).decode()
        # (raised without binding it to a local: `err = ...; raise err` would tie the exception to this frame and the
        #  frame to the exception -- a reference cycle that keeps the caller's tensors alive until a garbage collection)
        raise XmodalError(f"{what} failed: {msg} (status {status}, cuda error {dll.xm_last_cuda_error()})", int(status))


def call(name: str, *args) -> None:
    """Invoke an int-status entry point and raise on failure."""
    check(getattr(lib(), name)(*args), name)
