"""ctypes binding of libmfnerf_b200.so (C ABI declared in include/mfnerf_b200.h).

The product path has NO fallback: if the shared library is missing or a symbol does not resolve, importing
this module raises.  Nothing here imports or executes anything under oracle/.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "lib", "libmfnerf_b200.so"))
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "include", "mfnerf_b200.h"))

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C mf-nerf_b200/csrc`). mfnerf_b200 has no CPU / eager fallback.")

lib = ctypes.CDLL(LIB_PATH)

_C = {
    "int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float, "uint64_t": ctypes.c_uint64,
    "double": ctypes.c_double,
}


def _ctype(decl: str):
    """C type text of one parameter / return value -> ctypes type (pointers are opaque void*)"""
    decl = decl.replace("const", " ").strip()
    if "*" in decl:
        return ctypes.c_char_p if decl.split("*")[0].strip() == "char" else ctypes.c_void_p
    if decl == "void":
        return None
    return _C[decl.split()[0]]


def declared_functions(header_path: str = HEADER_PATH):
    """Parse `rettype mfn_xxx(args);` prototypes out of the public header -> {name: (restype, [argtypes])}."""
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"^([A-Za-z_][\w \*]*?)\s*\b(mfn_\w+)\s*\(([^)]*)\)\s*;", text, flags=re.M):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                a_type = a[: a.rindex("*") + 1] if "*" in a else a.rsplit(" ", 1)[0]
                argtypes.append(_ctype(a_type))
        out[name] = (_ctype(ret), argtypes)
    return out


FUNCS = declared_functions()
for _name, (_ret, _args) in FUNCS.items():
    _f = getattr(lib, _name)  # AttributeError here = header / library mismatch: fail loudly
    _f.restype = _ret
    _f.argtypes = _args


class MfnError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib.mfn_last_error()
        raise MfnError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def call(name: str, *args):
    rc = getattr(lib, name)(*args)
    check(rc, name)
