"""ctypes binding of the C-ABI library ``libgcc_b200.so`` (declared in ``include/gcc_b200.h``).

The prototypes are parsed from the header, so the header is the single source of truth for the
boundary.  There is no fallback: if the library is missing or a call fails, an exception is
raised (``GccB200Error``).
"""
import ctypes
import os
import re
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "gcc_b200.h")
LIBPATH = os.path.join(HERE, "libgcc_b200.so")


class GccB200Error(RuntimeError):
    pass


_CTYPES = {
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "long long": ctypes.c_longlong,
    "int64_t": ctypes.c_int64,
    "uint64_t": ctypes.c_uint64,
    "void": None,
}


def _ctype(decl):
    decl = decl.strip()
    if decl.endswith("*") or "*" in decl:
        base = decl.replace("const", "").replace("*", "").strip()
        if base == "char":
            return ctypes.c_char_p
        return ctypes.c_void_p
    decl = decl.replace("const", "").strip()
    return _CTYPES[decl]


def parse_header(path=HEADER):
    """Return {name: (restype, [argtypes], [argnames])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(gcc_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a)
                argtypes.append(_ctype(mm.group(1)))
                argnames.append(mm.group(2))
        protos[name] = (_ctype(ret) if ret != "void" else None, argtypes, argnames)
    return protos


_lib = None
_protos = None


def lib():
    """Load the shared library (building nothing: see gcc_b200._build / __graft_entry__.build)."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise GccB200Error(
            "libgcc_b200.so is missing (%s); run `python -m gcc_b200._build`. There is no CPU fallback." % LIBPATH)
    l = ctypes.CDLL(LIBPATH)
    _protos = parse_header()
    for name, (ret, argtypes, _) in _protos.items():
        fn = getattr(l, name)  # AttributeError if the header declares a symbol the library lacks
        fn.restype = ret
        fn.argtypes = argtypes
    _lib = l
    return l


_tls = threading.local()


def _bind_thread(l):
    """Once per host thread (torch's autograd engine runs backward on its own threads)."""
    import torch
    dev = torch.cuda.current_device()
    if getattr(_tls, "device", None) != dev:
        if l.gcc_bind_thread(dev) != 0:
            raise GccB200Error("gcc_bind_thread(%d) failed: %s" % (dev, l.gcc_last_error().decode()))
        _tls.device = dev


def call(name, *args):
    """Call an int-returning entry point and raise on a non-zero status."""
    l = lib()
    _bind_thread(l)
    rc = getattr(l, name)(*args)
    if rc != 0:
        msg = l.gcc_last_error()
        raise GccB200Error("%s failed (%d): %s" % (name, rc, msg.decode() if msg else "?"))


def ptr(t):
    """Device pointer of a torch tensor (or None) as an int for ctypes."""
    return None if t is None else t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
