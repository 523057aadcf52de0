"""CPU-side checks of the C ABI: the shared library builds for sm_100a, loads, and exports exactly the symbols
include/otk.h declares (no compute call is made - there is no GPU here)."""
import os
import subprocess

import pytest

from ot_vae_lightning_b200 import _native as N


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(N.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return N.load(require_gpu=False)


def test_every_declared_symbol_is_bound_and_exported(lib):
    declared = set(N.declared_symbols())
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
    exported = subprocess.run(["nm", "-D", "--defined-only", N.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in exported.splitlines() if " T " in line}
    assert declared <= exported, declared - exported
    assert not {s for s in exported if s.startswith("otk_")} - declared, "undeclared otk_* exports"


def test_abi_version_and_status_strings(lib):
    assert lib.otk_abi_version() == 1
    assert lib.otk_status_string(0) == b"ok"
    assert b"workspace" in lib.otk_status_string(-2)
    assert lib.otk_stats_update_workspace_bytes(1, 100, 128) >= 128 * 128 * 8
    assert lib.otk_sqrtm_workspace_bytes(2, 64) > 2 * 5 * 64 * 64 * 4


def test_library_targets_sm100a():
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", N.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_compute_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ot_vae_lightning_b200.ot import sqrtm
    with pytest.raises(N.NativeError):
        sqrtm(torch.eye(4))
