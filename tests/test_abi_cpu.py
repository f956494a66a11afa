"""CPU-side checks of the boundary: the library loads, exports every symbol include/bpt.h declares, struct
layouts match the header, host-only helpers work, and the product fails loudly without a GPU."""
import ctypes as C
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bpt = importlib.import_module("single-file-vulkan-pathtracing_b200")


def header_symbols():
    text = open(os.path.join(ROOT, "include", "bpt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bpt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = bpt.load_library()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), s
    assert sorted(bpt.ABI) == syms, set(syms) ^ set(bpt.ABI)
    assert L.bpt_abi_version() == 2


def test_struct_layouts_match_header():
    assert C.sizeof(bpt.Params) == 7 * 4 + 9 * 4 + 2 * 4 + 2 * 4 + 3 * 4 + 2 * 4
    assert C.sizeof(bpt.Stats) == 13 * 8
    assert C.sizeof(bpt.AccelInfo) == 6 * 4 + 2 * 8 + 2 * 4
    assert bpt.NODE8_DTYPE.itemsize == 64 and bpt.WOOP_DTYPE.itemsize == 64 and bpt.HIT_DTYPE.itemsize == 16


def test_python_constants_match_the_header():
    """Every BPT_OPT_* / BPT_ACCUM_* / BPT_SAMPLER_* the header defines has the same value in the ctypes mirror."""
    text = open(os.path.join(ROOT, "include", "bpt.h")).read()
    defs = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+BPT_((?:OPT|ACCUM|SAMPLER)_[A-Z0-9_]+)\s+(-?\w+)", text)}
    assert len([k for k in defs if k.startswith("OPT_")]) >= 12 and "OPT_FUSED_PATHS" in defs
    for name, value in defs.items():
        assert getattr(bpt, name) == value, name
    assert len({v for k, v in defs.items() if k.startswith("OPT_")}) == len([k for k in defs if k.startswith("OPT_")])  # ids unique


def test_default_params_are_the_reference_constants():
    p = bpt.default_params()
    assert (p.width, p.height, p.spp_per_frame, p.max_depth, p.frame) == (1024, 1024, 32, 8, 0)  # main.cpp:16-17, raygen.rgen:43,62
    assert list(p.cam_origin) == [0.0, -1.0, 5.0] and list(p.cam_target) == [0.0, -1.0, 2.0]    # raygen.rgen:55-56
    assert [round(x, 6) for x in p.sky] == [0.7, 0.6, 0.5]                                        # miss.rmiss:10
    assert abs(p.tmin - 0.001) < 1e-9 and p.tmax == 10000.0                                       # raygen.rgen:71,73
    assert p.accum_mode == bpt.ACCUM_FLOAT4 and p.sampler == bpt.SAMPLER_UNIFORM


def test_tile_rows_partition():
    for h, n in ((4096, 8), (1024, 4), (256, 2), (1080, 1), (1080, 7), (5, 3)):  # incl. heights the ranks do not divide
        rows = [bpt.tile_rows(h, r, n) for r in range(n)]
        assert rows[0][0] == 0 and sum(r[1] for r in rows) == h
        for a, b in zip(rows, rows[1:]):
            assert a[0] + a[1] == b[0]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bpt.BptError) as e:
        bpt.PathTracer(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "single-file-vulkan-pathtracing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower(), (dirpath, f)


def test_nccl_missing_is_an_error_code_not_a_crash():
    """when libnccl cannot be loaded, bpt_nccl_unique_id must return BPT_E_NCCL with a message (it used to
    call dlerror() twice and build a std::string from NULL). The loader is a process-wide once: run in a subprocess."""
    import subprocess
    import sys
    code = (
        "import importlib, ctypes as C\n"
        f"import sys; sys.path.insert(0, {ROOT!r})\n"
        "bpt = importlib.import_module('single-file-vulkan-pathtracing_b200')\n"
        "L = bpt.load_library()\n"
        "buf = (C.c_uint8 * 128)()\n"
        "rc = L.bpt_nccl_unique_id(buf)\n"
        "msg = L.bpt_last_error(None).decode()\n"
        "assert rc == -3, rc\n"
        "assert 'cannot dlopen /nonexistent/libnccl-missing.so' in msg, msg\n"
        "rc2 = L.bpt_nccl_unique_id(buf)\n"
        "assert rc2 == -3\n"
        "print('OK', msg)\n")
    env = dict(os.environ, BPT_NCCL_LIB="/nonexistent/libnccl-missing.so")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "OK" in r.stdout, (r.returncode, r.stdout, r.stderr)
