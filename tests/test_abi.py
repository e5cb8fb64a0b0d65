"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and exports exactly
the symbols include/u2b200.h declares; the ctypes table mirrors the header."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "u2b200.h")).read() + open(os.path.join(ROOT, "include", "u2b200_train.h")).read()
    return sorted(set(re.findall(r"U2_API\s+[\w\s\*]+?\b(u2_\w+)\s*\(", src)))


def test_header_declares_symbols():
    syms = header_symbols()
    assert "u2_gemm_bf16" in syms and "u2_gemv_bf16" in syms and len(syms) >= 15


def test_library_exports_every_declared_symbol():
    from u2tokenizer_b200 import _lib
    if not _lib.lib_path().exists():
        from u2tokenizer_b200 import build
        build.build()
    lib = ctypes.CDLL(str(_lib.lib_path()))
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in u2b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == header_symbols()
    assert _lib.load().u2_version() == 1


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: the product path fails loudly off-GPU."""
    import torch
    from u2tokenizer_b200 import ops
    with pytest.raises(RuntimeError):
        ops.linear(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    if not torch.cuda.is_available():
        from u2tokenizer_b200.engine import U2Engine
        with pytest.raises(RuntimeError):
            U2Engine(None, {})


def test_struct_mirrors_match_header_field_order():
    """Every ctypes Structure in _lib.py lists the fields of the C struct it mirrors in the header's order (a renamed or
    reordered field - e.g. the v_sb / v_ss / v_sh strides of u2_fa_desc - would silently shift every later argument)."""
    from u2tokenizer_b200 import _lib
    src = open(os.path.join(ROOT, "include", "u2b200.h")).read() + open(os.path.join(ROOT, "include", "u2b200_train.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    structs = dict(re.findall(r"typedef struct (u2_\w+) \{(.*?)\} \1;", src, flags=re.S))
    checked = 0
    for name, cls in vars(_lib).items():
        doc = getattr(cls, "__doc__", None) or ""
        m = re.search(r"Mirror of ``(u2_\w+)``", doc)
        if not (isinstance(cls, type) and issubclass(cls, ctypes.Structure) and m):
            continue
        body = structs[m.group(1)]
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.sub(r"\[.*?\]", "", part).replace("*", " ").split()[-1])
        assert [f[0] for f in cls._fields_] == names, f"{name} does not mirror {m.group(1)}: {names}"
        checked += 1
    assert checked >= 5


def test_integration_doc_lists_every_entry_point():
    """INTEGRATION.md is the maintainer-facing map 'entry point -> reference call site': no exported symbol may be missing."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [s for s in header_symbols() if not re.search(re.escape(s) + r"\b", doc) and not re.search(
        re.escape(s.rsplit("_", 1)[0]) + r"_\*", doc)]
    assert not missing, missing


def test_struct_mirrors_have_the_c_layout(tmp_path):
    """sizeof and every offsetof of the descriptor structs as gcc lays them out (the headers are plain C) against the
    ctypes mirrors the Python host side passes across the ABI."""
    import shutil
    import subprocess
    from u2tokenizer_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    mirrors = {}
    for cls in vars(_lib).values():
        m = re.search(r"Mirror of ``(u2_\w+)``", getattr(cls, "__doc__", None) or "")
        if isinstance(cls, type) and issubclass(cls, ctypes.Structure) and m:
            mirrors[m.group(1)] = cls
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "u2b200.h"', '#include "u2b200_train.h"', "int main(void) {"]
    for cname, cls in mirrors.items():
        lines.append(f'  printf("{cname} sizeof %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    n = 0
    for line in out.splitlines():
        cname, field, val = line.split()
        cls = mirrors[cname]
        got = ctypes.sizeof(cls) if field == "sizeof" else getattr(cls, field).offset
        assert got == int(val), f"{cname}.{field}: C {val}, ctypes {got}"
        n += 1
    assert n > 50
