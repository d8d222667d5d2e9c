"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/tdm_b200.h declares,
the ctypes table in tdm_b200/_lib.py covers exactly those symbols, and argument errors come back as
negative codes + a message (no compute call needs a device for that)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = (ROOT / "include" / "tdm_b200.h").read_text()


def _declared():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    body = re.sub(r"typedef struct \w+ \{.*?\} \w+;", "", body, flags=re.S)
    return sorted(set(re.findall(r"\b(tdm_[a-z0-9_]+)\s*\(", body)))


@pytest.fixture(scope="module")
def lib():
    from tdm_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_header_symbols_are_exported_and_bound(lib):
    from tdm_b200 import _lib
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tdm_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, set(_lib.SIGNATURES) ^ set(names)


def test_struct_layouts_match_header():
    """Field order of the ctypes mirrors == field order in the header structs."""
    from tdm_b200 import _lib
    for cname, cls in (("TdmGemm", _lib.TdmGemm), ("TdmAttn", _lib.TdmAttn), ("TdmDitShape", _lib.TdmDitShape),
                       ("TdmBlockWeights", _lib.TdmBlockWeights), ("TdmBlockGrads", _lib.TdmBlockGrads)):
        m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), HEADER, flags=re.S)
        body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(void|float|int32_t|int64_t)\s*\*?\s*", "", decl)
            fields += [f.strip().lstrip("*").strip() for f in decl.split(",")]
        assert fields == [f[0] for f in cls._fields_], cname


def test_errors_are_codes_not_exceptions(lib):
    assert lib.tdm_version() >= 2
    assert lib.tdm_gemm_bf16(None, None) == -1
    assert b"null descriptor" in lib.tdm_last_error()
    from tdm_b200 import _lib
    g = _lib.TdmGemm()
    g.M, g.N, g.K = 128, 12, 64  # N not a multiple of 8
    g.A = g.B = 256
    g.out_f32 = 256
    assert lib.tdm_gemm_bf16(C.byref(g), None) == -1
    assert b"multiple of 8" in lib.tdm_last_error()
    assert lib.tdm_solver_step(256, None, 1.0, 256, 0, 256, None, 256, 256, 0, 1000, None, 7, None, 256, None, 0, 1, 4, 8, 16, None) == -1
    assert b"currently not supported" in lib.tdm_last_error()  # the reference's ValueError text (utils.py:56-57)
    a = _lib.TdmAttn()
    a.Q = a.K = a.V = a.O = 256
    a.B = a.H = 1
    a.Sq = a.Sk = 128
    a.d = 70
    assert lib.tdm_attn_fwd(C.byref(a), None) == -1
    assert b"head_dim" in lib.tdm_last_error()
    assert lib.tdm_adamw_step(256, 256, 256, 256, None, 6, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, None, None) == -1


def test_product_path_has_no_cpu_fallback_and_never_imports_the_oracle():
    import torch
    from tdm_b200 import _lib, ops
    with pytest.raises(_lib.TdmError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16),
                 out_f32=torch.zeros(8, 8))
    for py in (ROOT / "tdm_b200").glob("*.py"):
        src = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), py
