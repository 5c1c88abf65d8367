"""The ctypes signature table (rocnrdma_b200/_native.py) against the C prototypes it binds: every RN_API entry
point it names must exist in the sources with the same number of parameters, the same parameter WIDTHS
(pointer / 64-bit / 32-bit) in the same order, and a return type of the same width.  An argument added on
one side only is undefined behaviour at call time, not an error -- this test is what catches it."""
import ctypes as C
import re
from pathlib import Path

from rocnrdma_b200 import _native as N

ROOT = Path(__file__).resolve().parent.parent / "rocnrdma_b200" / "csrc"
PROTO = re.compile(r'(?:RN_API|extern\s+"C"\s+__attribute__\(\(visibility\("default"\)\)\))\s+([\w\s\*]+?)\s*\b(rn_\w+)\s*\(([^)]*)\)\s*\{', re.S)


def _c_width(t: str) -> str:
    t = re.sub(r"\bconst\b|\bvolatile\b|\bstruct\b", "", t).strip()
    if "*" in t:
        return "ptr"
    base = t.split()[0] if t.split() else "void"
    if t.startswith("unsigned long long") or t.startswith("long long") or base in ("uint64_t", "int64_t", "size_t", "uintptr_t"):
        return "64"
    if base in ("void",):
        return "void"
    if base in ("uint32_t", "int32_t", "int", "unsigned", "float", "uint16_t", "uint8_t", "bool", "char", "short"):
        return "32"          # small integers are promoted to a register-wide slot in the SysV x86-64 ABI
    raise AssertionError(f"unknown C type {t!r}")


def _ct_width(t) -> str:
    if t is None:
        return "void"
    if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, (C._Pointer, C._CFuncPtr))):
        return "ptr"
    return "64" if C.sizeof(t) == 8 else "32"


def _prototypes():
    out = {}
    for path in list(ROOT.rglob("*.cu")) + list(ROOT.rglob("*.cc")):
        text = re.sub(r"//[^\n]*", "", path.read_text())
        for ret, name, args in PROTO.findall(text):
            params = [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"]
            widths = []
            for p in params:
                p = re.sub(r"\s*=\s*[^,]+$", "", p)
                ty = p.rsplit(" ", 1)[0] if not p.endswith("*") and " " in p else p
                if "*" in p:
                    ty = p[:p.rindex("*") + 1]
                widths.append(_c_width(ty))
            out[name] = (_c_width(ret), widths, path.name)
    return out


def test_every_bound_symbol_matches_its_c_prototype():
    protos = _prototypes()
    assert len(protos) > 60, "prototype scan found suspiciously few entry points"
    checked = 0
    optional_missing = []
    for table, optional in ((N._SIGS, False), (N._OPTIONAL_SIGS, True)):
        for name, (res, args) in table.items():
            if name not in protos:
                if optional:
                    optional_missing.append(name)       # compile-gated sources (verbs backend) may be absent
                    continue
                raise AssertionError(f"{name} is bound in _native.py but has no RN_API definition in csrc/")
            c_ret, c_args, where = protos[name]
            py_args = [_ct_width(a) for a in args]
            assert py_args == c_args, f"{name} ({where}): ctypes {py_args} vs C {c_args}"
            py_ret = _ct_width(res)
            assert py_ret == c_ret, f"{name} ({where}): return {py_ret} vs {c_ret}"
            checked += 1
    assert checked > 60
