"""CPU: the C-ABI library loads, exports every symbol include/segofa_b200.h declares, and the
ctypes struct layouts agree with what a C compiler makes of the header (no GPU calls)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from helpers import ROOT

HEADER = os.path.join(ROOT, "include", "segofa_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from ifseg_b200 import _lib

    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from ifseg_b200 import _lib

    src = open(HEADER).read()
    declared = set(re.findall(r"\b(sgf_[a-z0-9_]+)\s*\(", src))
    assert len(declared) >= 13
    bound = {n for n, _, _ in _lib.EXPORTS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.sgf_abi_version() == 3
    assert lib.sgf_launch_count() == 0


def test_struct_layouts_match_the_header(tmp_path):
    from ifseg_b200 import _lib

    structs = {"sgf_gemm_args": _lib.GemmArgs, "sgf_conv3x3_args": _lib.Conv3x3Args, "sgf_conv2d_args": _lib.Conv2dArgs, "sgf_rowln_args": _lib.RowLnArgs,
               "sgf_relblock": _lib.RelBlock, "sgf_bias_args": _lib.BiasArgs, "sgf_attention_args": _lib.AttentionArgs,
               "sgf_segmask_args": _lib.SegmaskArgs, "sgf_segloss_args": _lib.SeglossArgs,
               "sgf_segloss_bwd_args": _lib.SeglossBwdArgs, "sgf_rowln_bwd_args": _lib.RowLnBwdArgs,
               "sgf_attention_bwd_args": _lib.AttentionBwdArgs, "sgf_bias_bwd_args": _lib.BiasBwdArgs, "sgf_artsample_args": _lib.ArtSampleArgs}
    src = open(HEADER).read()
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    expect = {}
    for cname, ct in structs.items():
        body = re.search(r"typedef struct \{([^{}]*)\}\s*" + cname + ";", src, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                nm = re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[\d+\])?\s*$", part.strip())[0]
                fields.append(nm)
        assert fields == [f[0] for f in ct._fields_], (cname, fields)
        prog.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        expect[cname] = C.sizeof(ct)
        for f in fields:
            prog.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
            expect[f"{cname}.{f}"] = getattr(ct, f).offset
    prog.append("return 0;}")
    cfile = tmp_path / "layout.c"
    cfile.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(cfile)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = {l.split()[0]: int(l.split()[1]) for l in out.strip().splitlines()}
    assert got == expect


def test_invalid_arguments_are_reported_not_crashed(lib):
    from ifseg_b200 import _lib

    args = _lib.GemmArgs()
    rc = lib.sgf_gemm_bf16(C.byref(args), None)
    assert rc == 1 and b"bad shape" in lib.sgf_last_error()
    with pytest.raises(ValueError, match="bad shape"):
        _lib.check(rc, "sgf_gemm_bf16")
    a = _lib.AttentionArgs()
    assert lib.sgf_attention_bf16(C.byref(a), None) == 1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from ifseg_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        _lib.load()
