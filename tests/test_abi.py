"""CPU: the C-ABI library loads, exports every symbol include/dm_b200.h declares, and the ctypes
signatures in diffusionmodel_b200/_lib.py agree with the header prototypes (no compute calls)."""
import os
import re

from diffusionmodel_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _prototypes():
    src = open(os.path.join(ROOT, "include", "dm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|void|long long|const char\*)\s+(dm_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        sig = ""
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    sig += "p"
                elif a.startswith("long long"):
                    sig += "l"
                elif a.startswith("double"):
                    sig += "d"
                elif a.startswith("float"):
                    sig += "f"
                elif a.startswith("int"):
                    sig += "i"
                else:
                    raise AssertionError(f"unparsed argument {a!r} in {name}")
        out[name] = sig
    return out


def test_library_builds_and_exports_every_declared_symbol():
    from diffusionmodel_b200 import build
    build.build()
    lib = _lib.lib()
    declared = _lib.declared_symbols()
    assert len(declared) >= 40
    for s in declared:
        assert hasattr(lib, s), f"{s} declared in dm_b200.h but not exported by libdm_b200.so"
    assert lib.dm_version() == 100


def test_ctypes_signatures_match_header():
    protos = _prototypes()
    for name, sig in _lib._SIGS.items():
        assert name in protos, name
        assert sig.replace(" ", "") == protos[name], f"{name}: binding {sig.replace(' ', '')} != header {protos[name]}"
    unbound = set(protos) - set(_lib._SIGS) - {"dm_last_error", "dm_version", "dm_debug_set", "dm_launch_count", "dm_kernel_count", "dm_last_kernel", "dm_set_sm_limit",
                                                     "dm_set_workspace"}
    assert not unbound, f"header entry points without a Python binding: {sorted(unbound)}"


def test_sass_is_blackwell_native():
    """The conv kernels must be tcgen05/TMA (UTCHMMA, UTMALDG, LDTM in SASS), not legacy mma.sync (HMMA)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        return
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    # the only warp-level MMA in the library is the 16-row split-K product of the up0 data gradient (skinny.cu: one
    # m16n8k16 fragment IS the whole M extent, the kernel is bound by the 302 MB weight stream, not by the tensor pipe)
    for fn in sass.split("Function : ")[1:]:
        name = fn.split("\n", 1)[0]
        if " HMMA" in fn:
            assert "skinny_gemm_kernel" in name, name
        if "conv_gemm_cu" in name:                       # every GEMM kernel of conv_gemm.cu issues tcgen05.mma
            assert "UTCHMMA" in fn, name
