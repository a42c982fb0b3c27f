"""ctypes binding of libndjir_b200.so.  Signatures are read from include/ndjir_b200.h so the header stays
the single source of truth.  There is NO fallback: if the shared library is missing or a symbol is absent
this raises - the product path never routes around the CUDA kernels.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libndjir_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ndjir_b200.h")

_CTYPE = {
    "long long": ctypes.c_longlong,
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "cudaStream_t": ctypes.c_void_p,
    "const char*": ctypes.c_char_p,
}


class NdjirError(RuntimeError):
    pass


def parse_header(path=HEADER_PATH):
    """Returns {name: (restype, [(ctype_key, argname), ...])} for every ndjir_* prototype."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    text = re.sub(r"^\s*#.*$", " ", text, flags=re.M)
    protos = {}
    for m in re.finditer(r"\b(int|long long|void)\s+(ndjir_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        parsed = []
        for a in args.split(","):
            a = " ".join(a.split())
            if not a or a == "void":
                continue
            mm = re.match(r"^(.*?)(\w+)$", a)
            ty, an = mm.group(1).strip(), mm.group(2)
            ty = ty.replace(" *", "*")
            parsed.append((ty, an))
        protos[name] = (ret, parsed)
    return protos


def _to_ctype(ty):
    if ty in _CTYPE:
        return _CTYPE[ty]
    if ty.endswith("*"):
        return ctypes.c_void_p
    raise NdjirError(f"unknown C type in header: {ty!r}")


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise NdjirError(
                f"{LIB_PATH} is missing: build it with `python -m ndjir_b200.build` "
                "(there is no CPU or PyTorch fallback for the ndjir_b200 hot path)")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        self._fn = {}
        for name, (ret, args) in self.protos.items():
            try:
                f = getattr(self.cdll, name)
            except AttributeError as e:
                raise NdjirError(f"symbol {name} declared in {HEADER_PATH} is not exported by {LIB_PATH}") from e
            f.restype = ctypes.c_longlong if ret == "long long" else (None if ret == "void" else ctypes.c_int)
            f.argtypes = [_to_ctype(t) for t, _ in args]
            self._fn[name] = f

    def call(self, name, *args):
        """Calls an entry point.  Device pointers may be ints or objects with .data_ptr(); host float/int
        triples may be Python sequences.  Raises NdjirError on a non-zero status."""
        ret, proto = self.protos[name]
        if len(args) != len(proto):
            raise TypeError(f"{name} takes {len(proto)} arguments ({[a for _, a in proto]}), got {len(args)}")
        conv, keep = [], []
        for (ty, an), v in zip(proto, args):
            if ty.endswith("*") and ty != "const char*":
                if v is None:
                    conv.append(None)
                elif hasattr(v, "data_ptr"):
                    conv.append(ctypes.c_void_p(v.data_ptr()))
                elif isinstance(v, int):
                    conv.append(ctypes.c_void_p(v))
                elif isinstance(v, ctypes.Array):
                    conv.append(ctypes.cast(v, ctypes.c_void_p))
                elif isinstance(v, ctypes.Structure):
                    conv.append(ctypes.cast(ctypes.pointer(v), ctypes.c_void_p))
                else:  # host sequence -> temporary C array
                    base = ty.replace("const ", "").rstrip("*").strip()
                    cty = {"float": ctypes.c_float, "int": ctypes.c_int, "long long": ctypes.c_longlong}[base]
                    arr = (cty * len(v))(*[cty(x).value for x in v])
                    keep.append(arr)
                    conv.append(ctypes.cast(arr, ctypes.c_void_p))
            elif ty == "const char*":
                conv.append(v.encode() if isinstance(v, str) else v)
            elif ty == "cudaStream_t":
                conv.append(ctypes.c_void_p(int(v) if v else 0))
            elif ty == "float":
                conv.append(float(v))
            else:
                conv.append(int(v))
        r = self._fn[name](*conv)
        if ret == "int" and r != 0:
            raise NdjirError(f"{name} failed with status {r}" + (" (invalid argument)" if r == -1 else " (CUDA error)"))
        return r


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB


def call(name, *args):
    return lib().call(name, *args)
