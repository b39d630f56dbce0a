"""The C-ABI library loads and exports exactly what include/gpjax_b200.h declares (no compute calls)."""
import ctypes
import os
import re
import subprocess

import pytest

from gpjax_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gpjax_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpb_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_python_prototypes():
    assert header_functions() == sorted(_abi.PROTOTYPES)


def test_header_argument_counts_match_prototypes():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _abi.PROTOTYPES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("void", "") else len(params.split(","))
        assert n == len(args), f"{name}: header has {n} parameters, _abi.py declares {len(args)}"


@pytest.fixture(scope="module")
def cuda_lib_path():
    from gpjax_b200.build import build

    return build()


def test_cuda_library_exports_every_declared_symbol(cuda_lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", cuda_lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (gpb_[a-z0-9_]+)", out))
    assert exported == set(header_functions())


def test_cuda_library_loads_and_answers_queries(cuda_lib_path):
    lib = _abi.declare(ctypes.CDLL(cuda_lib_path))
    assert b"sm_100a" in lib.gpb_version()
    assert lib.gpb_block_size() in (256, 512, 1024)
    assert lib.gpb_block_size_for(50_000) in (256, 512, 1024, 2048) and lib.gpb_block_size_for(100) == lib.gpb_block_size()
    assert lib.gpb_max_input_dim() >= 16
    assert lib.gpb_mll_workspace_bytes(50000, 8) > 0
    assert lib.gpb_factor_workspace_bytes(1000, 1, 0) < lib.gpb_factor_workspace_bytes(1000, 1, 1)


def test_cuda_library_is_sm100a_with_dmma(cuda_lib_path):
    """The shipped kernels are sm_100a SASS and the GEMM really issues FP64 tensor-core MMAs."""
    out = subprocess.run(["cuobjdump", "-lelf", cuda_lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    obj = os.path.join(os.path.dirname(cuda_lib_path), "obj", "gemm_f64.o")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    assert sass.count("DMMA.8x8x4") >= 64 and "LDGSTS" in sass
    # both tile shapes are in the binary, for all four operand layouts: the full 128 x 64 tile and the quarter tile that few-tile
    # launches use (gemm_f64.cu: gemm())
    funcs = re.findall(r"Function : (\S*gemm_f64_kernel\S*)", sass)
    assert sum("ILi128ELi64ELi32ELi64ELi3ELi2ELb0ELb0E" in f for f in funcs) == 4
    assert sum("ILi64ELi32ELi32ELi32ELi3ELi4ELb0ELb0E" in f for f in funcs) == 4


def test_int8_kernel_is_tcgen05_with_tma_and_tmem(cuda_lib_path):
    """The Ozaki kernel really issues 5th-generation tensor-core int8 MMAs fed by TMA with TMEM accumulators, and the role
    loops are warp-uniform: the four UTCIMMA of a K-block are issued back to back (no R2UR waterfall in between)."""
    obj = os.path.join(os.path.dirname(cuda_lib_path), "obj", "ozaki_i8.o")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for mnemonic in ("UTCIMMA", "UTMALDG.2D", "UTCBAR", "LDTM", "UTCATOMSWS"):
        assert mnemonic in sass, mnemonic
    lines = [ln for ln in sass.splitlines() if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", ln)]
    idx = [i for i, ln in enumerate(lines) if "UTCIMMA" in ln]
    assert any(idx[j + 3] - idx[j] == 3 for j in range(len(idx) - 3)), "no run of 4 consecutive UTCIMMA instructions"


SHIM = os.path.join(ROOT, "gpjax_b200", "csrc", "xla_ffi_shim.cc")
XLA_STUB = os.path.join(ROOT, "tests", "xla_stub")
# entry points a jax.ffi binding of the hot path needs a handler for (queries, measurement hooks, raw int8 building blocks and
# the NCCL helpers -- XLA owns its collectives: jax.lax.psum -- are host-side / not part of the reference-facing surface)
SHIM_MUST_BIND = ["gpb_gram", "gpb_gram_bwd", "gpb_potrf_lower", "gpb_trsv_lower", "gpb_trsm_lower_left", "gpb_sum_log_diag",
                  "gpb_potri_lower", "gpb_mll_forward", "gpb_mll_backward", "gpb_sgpr_stats", "gpb_sgpr_stats_raw",
                  "gpb_sgpr_finish", "gpb_sgpr_grad_local", "gpb_sgpr_grad_finish", "gpb_svgp_finish", "gpb_svgp_grad_finish"]


def _compile_shim(path):
    return subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", XLA_STUB, "-I", "/usr/local/cuda/include", path],
                          capture_output=True, text=True)


def test_xla_ffi_shim_compiles_against_the_stub_and_binds_the_whole_path(tmp_path):
    """jaxlib is not installable here, so the jax.ffi handlers are type-checked against a stand-in of xla/ffi/api/ffi.h whose
    binder performs the real header's check (handler parameters == Bind() chain).  A deliberately wrong chain must fail."""
    src = open(SHIM).read()
    for fn in SHIM_MUST_BIND:
        assert re.search(r"\b" + fn + r"\b", src), f"no handler forwards to {fn}"
    assert src.count("XLA_FFI_DEFINE_HANDLER_SYMBOL(") >= 14
    ok = _compile_shim(SHIM)
    assert ok.returncode == 0, ok.stderr[-3000:]
    # the in-place backward declares its residuals as results (aliased), never mutates a read-only operand
    assert "ffi::Result<F64> sigma_out, ffi::Result<F64> ws_out" in src and "const_cast" not in src
    broken = tmp_path / "broken_shim.cc"
    bad = src.replace('#include "../../include/gpjax_b200.h"', f'#include "{HEADER}"')
    bad = bad.replace("ffi::Ffi::Bind().Ctx<Stream>().Arg<F64>().Ret<F64>());", "ffi::Ffi::Bind().Ctx<Stream>().Arg<F64>().Arg<F64>().Ret<F64>());")
    assert bad != src
    broken.write_text(bad)
    r = _compile_shim(str(broken))
    assert r.returncode != 0 and "does not match its Ffi::Bind() chain" in r.stderr


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gpjax_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "hostsim" not in txt or f in ("primitives.h", "algorithms.h"), f


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    from gpjax_b200 import _lib

    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.ExtensionMissingError):
        _lib.lib()
