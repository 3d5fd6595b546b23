"""CPU: the C-ABI library loads and exports every symbol include/supernova_b200.h declares;
without a GPU the product refuses to run (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "supernova_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    L = ctypes.CDLL(built)
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"


def test_no_cpu_fallback(built):
    import torch
    import supernova_b200 as sb
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sb.SnError, match="no CUDA device"):
        sb.Context(0)


def test_product_does_not_touch_the_oracle():
    """Nothing under supernova_b200/ may import, link or execute oracle/."""
    for dp, _, files in os.walk(os.path.join(ROOT, "supernova_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "sn_oracle" not in txt and "oracle." not in txt and "from oracle" not in txt, os.path.join(dp, f)


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/supernova_b200.h must compile as C99 (no C++ in the signatures, no torch types), and a C
    caller can name every struct and constant it declares."""
    import subprocess
    src = tmp_path / "c_abi.c"
    src.write_text('#include "supernova_b200.h"\n'
                   "int main(void) { sn_params p; sn_counts c; sn_synth s; sn_dup_stats d; sn_kmer_rec k; sn_ctx* x = 0;\n"
                   "  (void)p; (void)c; (void)s; (void)d; (void)k; (void)x; return SN_OK + SN_ERR_CUDA + SN_SEM_TADA; }\n")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
