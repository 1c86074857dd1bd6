"""CPU checks of the drop-in boundary: libpartmanip_b200.so loads without a GPU, exports every entry point
include/partmanip_b200.h declares (and the ctypes table binds exactly that set), the host-side mirrors of the reference
interface import, and the product never reaches into oracle/.  No kernel is launched here."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "partmanip_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol_and_binding_table_matches():
    from partmanip_b200 import _lib
    names = _declared()
    assert len(names) >= 40
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported by the library"
    assert sorted(_lib.SIGNATURES) == names, (sorted(set(names) ^ set(_lib.SIGNATURES)))
    assert _lib.lib.pm_version() >= 100 and _lib.lib.pm_has_tcgen05() == 1
    # pure host-side queries work without a device
    assert _lib.lib.pm_pointnet_encode_backward_ws_bytes(2048, 1024, 3, 0, 1) > 0
    assert _lib.lib.pm_adam_ws_bytes(1000) > 0 and _lib.lib.pm_pointnet_head_backward_ws_bytes(2048, 512) > 0


def test_binding_table_arity_and_pointer_slots_match_the_header():
    """Every ctypes signature has exactly the header's parameter count, and pointer / scalar slots line up (a pointer passed in an int slot
    would be truncated silently)."""
    import ctypes as C
    from partmanip_b200 import _lib
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    decls = re.findall(r"\b(int|size_t|const char\*)\s+(pm_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    assert len(decls) >= 60
    seen = set()
    for ret, name, params in decls:
        seen.add(name)
        assert _lib.SIGNATURES[name][0] is {"int": C.c_int, "size_t": C.c_size_t, "const char*": C.c_char_p}[ret], (name, ret)
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [q.strip() for q in params.split(",")]
        _, argtypes = _lib.SIGNATURES[name]
        assert len(plist) == len(argtypes), (name, len(plist), len(argtypes), plist)
        for decl, ct in zip(plist, argtypes):
            is_ptr_decl = "*" in decl or decl.startswith("pm_stream_t")
            is_ptr_ct = ct in (C.c_void_p, C.c_char_p) or hasattr(ct, "contents") or (isinstance(ct, type) and issubclass(ct, C._Pointer))
            assert is_ptr_decl == is_ptr_ct, (name, decl, ct)
            if not is_ptr_decl:                     # scalar widths: int64_t <-> c_int64, float <-> c_float, ...
                ctype = decl.rsplit(" ", 1)[0].replace("const ", "").strip()
                want = {"int": C.c_int, "int32_t": C.c_int, "int64_t": C.c_int64, "uint64_t": C.c_uint64, "float": C.c_float,
                        "size_t": C.c_size_t, "uint32_t": C.c_uint32}.get(ctype)
                assert want is not None and ct is want, (name, decl, ct)
    assert seen == set(_lib.SIGNATURES), sorted(seen ^ set(_lib.SIGNATURES))


def test_sm100a_tensor_core_and_no_legacy_mma_in_the_library():
    """The shipped .so carries sm_100a SASS with tcgen05 MMAs (UTCHMMA) and TMEM loads (LDTM) and no legacy HMMA path."""
    from partmanip_b200 import _lib
    try:
        sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    except (FileNotFoundError, subprocess.TimeoutExpired):
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass
    assert sass.count("UTCHMMA") >= 40 and "LDTM" in sass
    assert not re.search(r"\bHMMA\b", sass) and "HGMMA" not in sass


def test_host_mirrors_import_and_reject_cpu_devices():
    import torch
    from partmanip_b200.algorithms import dagger, ppo          # noqa: F401  (names train.py dispatches on)
    from partmanip_b200.algorithms.algo_utils import ActorCritic, Normalization, RolloutStorage   # noqa: F401
    from partmanip_b200 import ops
    with pytest.raises(ValueError, match="CUDA tensor"):
        ops.gae(torch.zeros(2, 2, 1), torch.zeros(2, 2, 1), torch.zeros(2, 2, 1, dtype=torch.bool), None, torch.zeros(2), 0.99, 0.95, None)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "partmanip_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(d, f)
            if f.endswith((".cu", ".cuh", ".h")):                 # comments may cite the oracle; code may not include it
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r"#include\s+[<\"][^>\"]*oracle", txt), os.path.join(d, f)
    code = "import sys; import partmanip_b200, partmanip_b200.algorithms; assert not any(m.startswith('oracle') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
