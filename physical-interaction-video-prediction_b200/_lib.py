"""ctypes binding of libpivp.so.  Prototypes are parsed from include/pivp.h so the header stays the single
source of truth for the C-ABI.  There is NO fallback: if the library is missing or a call fails, we raise."""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "pivp.h")
LIBPATH = os.path.join(HERE, "libpivp.so")

_CTYPES = {
    "int": ctypes.c_int, "long": ctypes.c_long, "float": ctypes.c_float, "size_t": ctypes.c_size_t,
    "void": None,
}


class PivpError(RuntimeError):
    pass


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes], [argnames])} for every `pivp_*` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(pivp_\w+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()

        def conv(decl):
            decl = decl.replace("const", " ").strip()
            if "*" in decl:
                return ctypes.c_char_p if decl.startswith("char") else ctypes.c_void_p
            return _CTYPES[decl.split()[0]]
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                nm = re.findall(r"(\w+)$", a)[0]
                argtypes.append(conv(a[: a.rfind(nm)]))
                argnames.append(nm)
        protos[name] = (conv(ret + " "), argtypes, argnames)
    return protos


class _Lib(object):
    def __init__(self):
        if not os.path.exists(LIBPATH):
            raise PivpError("libpivp.so is not built (%s). Run `python __graft_entry__.py` / build.build_lib(); "
                            "there is no CPU or PyTorch fallback for this path." % LIBPATH)
        self.cdll = ctypes.CDLL(LIBPATH)
        self.protos = parse_header()
        self.launches = 0
        for name, (res, argtypes, _) in self.protos.items():
            fn = getattr(self.cdll, name)
            fn.restype = res
            fn.argtypes = argtypes
        self.cdll.pivp_last_error.restype = ctypes.c_char_p

    def call(self, name, *args):
        """Call an int-returning entry point; raise PivpError with pivp_last_error() on failure."""
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            raise PivpError("%s failed (%d): %s" % (name, rc, self.cdll.pivp_last_error().decode()))
        self.launches += 1

    def query(self, name, *args):
        return getattr(self.cdll, name)(*args)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
