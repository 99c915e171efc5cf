"""CPU tests: the C-ABI library loads and exports every symbol declared in include/metamorph_b200.h;
argument validation fails loudly without touching a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "metamorph_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mm_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_declared_symbols():
    from metamorph_b200 import _build
    from metamorph_b200._lib import lib
    _build.build(verbose=False)
    l = lib()
    names = _declared()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(l, n)]
    assert not missing, f"declared but not exported: {missing}"
    assert l.mm_abi_version() == 1


def test_header_compiles_as_c():
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write('#include "metamorph_b200.h"\nint main(void){return 0;}\n')
        subprocess.run(["gcc", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", c],
                       check=True)


def test_bad_arguments_fail_loudly_without_gpu():
    from metamorph_b200._lib import MetaMorphB200Error, call, ll
    from ctypes import c_float, c_int, c_void_p
    with pytest.raises(MetaMorphB200Error, match="H%8"):
        call("mm_rmsnorm_fwd", c_void_p(0), c_void_p(0), c_void_p(0), ll(4), ll(7), c_float(1e-5), c_void_p(0))
    with pytest.raises(MetaMorphB200Error, match="batch"):
        call("mm_skinny_gemm", c_void_p(0), c_void_p(0), c_void_p(0), c_void_p(0), c_void_p(0), ll(64), ll(64),
             ll(64), ll(0), c_int(9), c_int(64), c_int(64), c_int(0), c_int(0), c_void_p(0))


def test_product_never_imports_oracle():
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "metamorph_b200")):
        for f in files:
            if f.endswith(".py") and re.search(r"^\s*(from|import)\s+oracle", open(os.path.join(dp, f)).read(), flags=re.M):
                bad.append(os.path.join(dp, f))
    assert not bad, f"product code must not import the oracle: {bad}"
