"""CPU: the C-ABI library builds, loads, and exports every symbol include/cusrl_b200.h declares.
No compute calls are made here (there is no GPU in the build container)."""

from __future__ import annotations

import ctypes

import pytest

from cusrl_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    declared = _lib.declared_symbols()
    assert len(declared) >= 15
    missing = [name for name in declared if not hasattr(lib, name)]
    assert not missing, f"declared in include/cusrl_b200.h but not exported: {missing}"


def test_every_bound_signature_is_declared(lib):
    declared = set(_lib.declared_symbols())
    extra = [name for name in _lib._SIGNATURES if name not in declared]
    assert not extra, f"bound in _lib.py but not declared in the header: {extra}"


def test_abi_version(lib):
    assert lib.cusrl_b200_abi_version() == 1


def test_argument_validation_without_gpu(lib):
    # argument errors are detected before any CUDA call, so these are safe on a CPU-only box
    code = lib.cusrl_b200_gae_f32(None, None, None, None, None, None, 24, 8, 1, 0.99, 0.95, -1.0, None)
    assert code == -1
    assert b"null pointer" in lib.cusrl_b200_last_error()
    buf = (ctypes.c_float * 8)()
    flags = (ctypes.c_uint8 * 8)()
    p, f = ctypes.addressof(buf), ctypes.addressof(flags)
    # gamma outside [0, 1): same rule as GeneralizedAdvantageEstimation.__init__ (reference gae.py:59-60)
    code = lib.cusrl_b200_gae_f32(p, f, p, p, p, p, 2, 4, 1, 1.0, 0.95, -1.0, None)
    assert code == -1 and b"gamma" in lib.cusrl_b200_last_error()
    code = lib.cusrl_b200_gae_f32(p, f, p, p, p, p, 2, 4, 1, 0.99, 1.5, -1.0, None)
    assert code == -1 and b"lamda" in lib.cusrl_b200_last_error()


def test_ops_refuse_cpu_tensors():
    import torch

    from cusrl_b200 import ops

    x = torch.zeros(2, 3, 1)
    flags = torch.zeros(2, 3, 1, dtype=torch.bool)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gae(x, flags, x, x, 0.99, 0.95)


def test_every_compute_entry_point_rejects_null_arguments_before_any_cuda_call():
    """Every launch entry point of the header, called with null pointers and zero sizes on a box WITHOUT a GPU, must come
    back with a negative argument-error code and a message: no crash, no CUDA call (a positive cudaError_t would mean the
    validation let the call through).  Run in a child process so that a crashing entry point fails this test instead of
    taking the test session down."""
    import subprocess
    import sys
    import textwrap

    script = textwrap.dedent("""
        import ctypes
        from cusrl_b200 import _lib
        lib = _lib.load()
        skip = ("_set_", "_supported", "_bytes", "abi_version", "last_error", "sm_count")
        bad = []
        names = [n for n in _lib._SIGNATURES if not any(s in n for s in skip)]
        for name in names:
            _, argtypes = _lib._SIGNATURES[name]
            args = [0.0 if a in (ctypes.c_float, ctypes.c_double)
                    else 0 if a in (ctypes.c_int, ctypes.c_int64, ctypes.c_size_t) else None for a in argtypes]
            code = getattr(lib, name)(*args)
            if code >= 0 or not lib.cusrl_b200_last_error():
                bad.append((name, code))
        print(len(names), bad)
    """)
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=300,
                         cwd=str(_lib.PKG_DIR.parent))
    assert out.returncode == 0, out.stderr[-2000:]
    count, bad = out.stdout.strip().split(" ", 1)
    assert int(count) >= 40 and bad == "[]", out.stdout
