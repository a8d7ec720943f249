"""Loader for libcelda_cuda.so (the product).  Fails loudly if the CUDA library is missing: there is no CPU fallback."""
import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libcelda_cuda.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "celda_cuda.h")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_llp = C.POINTER(C.c_longlong)
_TYPES = {
    "int": C.c_int, "double": C.c_double, "long long": C.c_longlong, "unsigned long long": C.c_ulonglong,
    "const int *": _ip, "int *": _ip, "const double *": _dp, "double *": _dp, "long long *": _llp,
    "float *": C.POINTER(C.c_float), "celda_chain *": C.c_void_p, "celda_chain **": C.POINTER(C.c_void_p),
    "const char *": C.c_char_p, "char *": C.c_char_p, "void": None, "celda_comm *": C.c_void_p,
    "celda_comm **": C.POINTER(C.c_void_p),
}


class CeldaCudaError(RuntimeError):
    """Raised when a libcelda_cuda entry point returns a non-zero status (the R shim would call Rf_error)."""


def build(verbose=False):
    """Compile libcelda_cuda.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libcelda_cuda.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return SO_PATH


def declared_functions(header=HEADER):
    """Parse include/celda_cuda.h -> {name: (restype, [argtypes])}; CELDA_CUDA_PLANNED blocks are skipped."""
    src = open(header).read()
    src = re.sub(r"#ifdef CELDA_CUDA_PLANNED.*?#endif /\* CELDA_CUDA_PLANNED \*/", "", src, flags=re.S)
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    out = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(celda_\w+)\s*\(([^;{}]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3)
        if "typedef" in ret or "struct" in ret:
            continue

        def norm(t):
            t = re.sub(r"\s+", " ", t.strip())
            t = re.sub(r"\s*\*\s*", " *", t)
            t = t.replace("* *", "**").replace(" **", " **")
            return t

        argt = []
        for a in args.split(","):
            a = a.strip()
            if a in ("void", ""):
                continue
            m2 = re.match(r"(.*?)(\w+)$", a)  # strip the parameter name
            argt.append(norm(m2.group(1)))
        out[name] = (norm(ret), argt)
    return out


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise CeldaCudaError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(celda_b200 has no CPU fallback)")
        L = C.CDLL(SO_PATH)
        for name, (ret, args) in declared_functions().items():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch
            fn.restype = _TYPES[ret]
            fn.argtypes = [_TYPES[a] for a in args]
        _lib = L
    return _lib


def check(rc):
    if rc:
        raise CeldaCudaError(lib().celda_cuda_last_error().decode())
