"""ctypes binding of the Spirit C API, generated from the prototypes in include/.

The same binder serves both libraries a parity test drives:
  * spirit_b200/libSpirit.so     -- this package's CUDA library (the product)
  * oracle/_ref/libSpirit_ref.so -- the unmodified reference built by oracle/Makefile (test infrastructure)
Both export the reference's C API (core/include/Spirit/*.h). The double-precision probes are
`SpiritB200_*` in the product (include/spirit_b200.h) and `refshim_*` in the oracle (oracle/ref_shim.cpp) with
identical signatures; `Library.probe(name)` resolves whichever the loaded library has.

The prototypes are parsed from the headers, so a symbol declared in include/ but missing from the library is
an error at load time (tests/test_abi.py relies on that).
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
INCLUDE = os.path.join(ROOT, "include")
# SPIRIT_B200_LIB selects an experimental build variant (spirit_b200/build.py); the default is the product library
PRODUCT_LIB = os.path.join(HERE, os.environ.get("SPIRIT_B200_LIB", "libSpirit.so"))
ORACLE_LIB = os.path.join(ROOT, "oracle", "_ref", "libSpirit_ref.so")
# the same reference built with -DSPIRIT_ENABLE_PINNING -DSPIRIT_ENABLE_DEFECTS (`make -C oracle pd`)
ORACLE_PD_LIB = os.path.join(ROOT, "oracle", "_ref", "libSpirit_ref_pd.so")


class Simulation_Run_Info(ctypes.Structure):
    """core/include/Spirit/Simulation.h:58-71"""
    _fields_ = [
        ("total_iterations", ctypes.c_int),
        ("total_walltime", ctypes.c_int),
        ("total_ips", ctypes.c_float),
        ("max_torque", ctypes.c_float),
        ("n_history_iteration", ctypes.c_int),
        ("history_iteration", ctypes.POINTER(ctypes.c_int)),
        ("n_history_max_torque", ctypes.c_int),
        ("history_max_torque", ctypes.POINTER(ctypes.c_float)),
        ("n_history_energy", ctypes.c_int),
        ("history_energy", ctypes.POINTER(ctypes.c_float)),
    ]


_SCALARS = {
    "void": None,
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "scalar": ctypes.c_double,
    "bool": ctypes.c_bool,
    "unsigned long long": ctypes.c_ulonglong,
    "Bravais_Lattice_Type": ctypes.c_int,
    "Spirit_Log_Level": ctypes.c_int,
    "Spirit_Log_Sender": ctypes.c_int,
    "Simulation_Run_Info": Simulation_Run_Info,
}


def _ctype(decl, is_return=False):
    """C declarator (type + optional name + optional []) -> ctypes type"""
    d = re.sub(r"\bconst\b", " ", decl).strip()
    n_ptr = d.count("*") + d.count("[")
    d = re.sub(r"\[[^\]]*\]", " ", d).replace("*", " ")
    words = d.split()
    # drop the parameter name (last word) unless the whole declarator is the type
    for take in (len(words), len(words) - 1):
        base = " ".join(words[:take])
        if base in _SCALARS or base in ("State", "char"):
            break
    else:
        raise ValueError("cannot parse C declarator: %r" % decl)
    if base == "State":
        return ctypes.c_void_p
    if base == "char":
        return ctypes.c_char_p if n_ptr == 1 else ctypes.c_void_p
    t = _SCALARS[base]
    if n_ptr == 0:
        return t
    if n_ptr > 1 or t is None:
        return ctypes.c_void_p
    if base == "Simulation_Run_Info":
        return ctypes.POINTER(Simulation_Run_Info)
    return ctypes.POINTER(t)


_PROTO = re.compile(r"SPIRIT_API\s+(.*?)\s*\b(\w+)\s*\(([^;]*)\)\s*SPIRIT_NOEXCEPT\s*;", re.S)


def _strip_defaults(params):
    out, depth, i = "", 0, 0
    params = re.sub(r"SPIRIT_DEFAULT\s*\(", "\x00(", params)
    while i < len(params):
        c = params[i]
        if c == "\x00":
            depth, i = 1, i + 2
            while depth:
                depth += {"(": 1, ")": -1}.get(params[i], 0)
                i += 1
            continue
        out += c
        i += 1
    return out


def parse_header(path):
    """-> {name: (restype, [argtypes])} for every SPIRIT_API prototype in the header"""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    text = "\n".join(l for l in text.split("\n") if not l.lstrip().startswith("#"))
    protos = {}
    for ret, name, params in _PROTO.findall(text):
        params = _strip_defaults(params).strip()
        args = [] if params in ("", "void") else [_ctype(p) for p in params.split(",")]
        protos[name] = (_ctype(ret + " ", True) if ret.strip() != "void" else None, args)
    return protos


def declared_prototypes(with_extensions=True):
    protos = {}
    d = os.path.join(INCLUDE, "Spirit")
    for f in sorted(os.listdir(d)):
        if f.endswith(".h"):
            protos.update(parse_header(os.path.join(d, f)))
    ext = parse_header(os.path.join(INCLUDE, "spirit_b200.h")) if with_extensions else {}
    return protos, ext


class Library:
    """A loaded Spirit-API library with typed functions as attributes."""

    def __init__(self, path, kind):
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s not found. Build it first: python -c 'import __graft_entry__ as g; g.build()'" % path)
        self.path, self.kind = path, kind
        self.cdll = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        api, ext = declared_prototypes()
        self.missing = []
        for name, (res, args) in api.items():
            self._bind(name, name, res, args)
        for name, (res, args) in ext.items():
            if kind == "product":
                self._bind(name, name, res, args)
            else:
                twin = name.replace("SpiritB200_", "refshim_")
                if hasattr(self.cdll, twin):
                    self._bind(name, twin, res, args)
        if kind == "oracle":
            for n, res, args in (("refshim_num_threads", ctypes.c_int, []), ("refshim_set_num_threads", None, [ctypes.c_int])):
                self._bind(n, n, res, args)

    def _bind(self, attr, symbol, res, args):
        try:
            f = getattr(self.cdll, symbol)
        except AttributeError:
            self.missing.append(symbol)
            return
        f.restype, f.argtypes = res, args
        setattr(self, attr, f)

    def probe(self, name):
        """The double-precision probe `name` (without prefix) of this library"""
        return getattr(self, "SpiritB200_" + name)


_cache = {}


def load_product():
    """The CUDA library. There is no fallback: if it is not built this raises."""
    if "product" not in _cache:
        _cache["product"] = Library(PRODUCT_LIB, "product")
    return _cache["product"]


def load_oracle():
    """TEST INFRASTRUCTURE ONLY: the reference CPU build (oracle/_ref). Never used by the product path."""
    if "oracle" not in _cache:
        _cache["oracle"] = Library(ORACLE_LIB, "oracle")
    return _cache["oracle"]


def load_oracle_pd():
    """TEST INFRASTRUCTURE ONLY: the reference CPU build with pinning and defects compiled in. Never used by the product path."""
    if "oracle_pd" not in _cache:
        _cache["oracle_pd"] = Library(ORACLE_PD_LIB, "oracle")
    return _cache["oracle_pd"]
