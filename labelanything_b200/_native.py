"""ctypes binding of liblabelanything_b200.so (the C ABI declared in include/labelanything_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  Allocation failures keep the substring "out of memory" because the reference's callers
string-match it (label_anything/experiment/run.py:339-355, experiment/utils.py:230-238).
"""
from __future__ import annotations

import ctypes
import os
import re
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "liblabelanything_b200.so"
HEADER_PATH = PKG_DIR.parent / "include" / "labelanything_b200.h"

_lib = None

_CTYPE = {
    "void*": ctypes.c_void_p,
    "const void*": ctypes.c_void_p,
    "const float*": ctypes.c_void_p,
    "float*": ctypes.c_void_p,
    "const int*": ctypes.c_void_p,
    "int*": ctypes.c_void_p,
    "const long long*": ctypes.c_void_p,
    "long long*": ctypes.c_void_p,
    "const unsigned char*": ctypes.c_void_p,
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
}


def declared_functions(header: Path = HEADER_PATH) -> dict[str, tuple[str, list[str]]]:
    """Parse `ret name(args);` prototypes out of the public header -> {name: (ret, [arg types])}."""
    text = header.read_text()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    out: dict[str, tuple[str, list[str]]] = {}
    for m in re.finditer(r"\b(int|long long|const char\*)\s+(la_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        types: list[str] = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                # drop the parameter name (last identifier), keep the type incl. '*'
                mm = re.match(r"^(.*?)(\w+)$", a)
                t = mm.group(1).strip() if mm else a
                t = t.replace(" *", "*")
                types.append(t)
        out[name] = (ret, types)
    return out


def lib() -> ctypes.CDLL:
    """Load the library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = LIB_PATH
    if os.environ.get("LA_B200_LIB"):   # experiment builds of tools/ (build.build_variant); never set by the product
        path = Path(os.environ["LA_B200_LIB"])
    if not path.exists():
        raise RuntimeError(
            f"labelanything_b200: native library {path} is missing. Build it with "
            "`python -m labelanything_b200.build` (or __graft_entry__.build()). There is no CPU/eager fallback."
        )
    cdll = ctypes.CDLL(str(path))
    for name, (ret, types) in declared_functions().items():
        fn = getattr(cdll, name)  # AttributeError if the header and the library disagree
        fn.restype = {"const char*": ctypes.c_char_p, "long long": ctypes.c_longlong}.get(ret, ctypes.c_int)
        fn.argtypes = [_CTYPE[t] for t in types]
    _lib = cdll
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().la_last_error().decode(errors="replace")
        raise RuntimeError(f"labelanything_b200.{what} failed (code {rc}): {msg}")
