"""Kernel-level GPU parity against the oracle: every entry point of include/dpsttc.h through the C ABI
(tools/gpu_check.py holds the sweep so it can also be run stand-alone under gpurun)."""
import importlib.util
import os

import pytest

from helpers import REPO

pytestmark = pytest.mark.gpu


def test_kernel_parity_sweep(capsys):
    spec = importlib.util.spec_from_file_location("gpu_check", os.path.join(REPO, "tools", "gpu_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert len(mod.results) >= 100, "sweep did not run"
    assert not mod.failures, "\n".join(mod.failures)


def test_library_really_ran_on_the_gpu():
    from dps_ttc_b200 import _lib
    import ctypes
    sm = ctypes.c_int(0)
    _lib.check(_lib.lib().dps_device_sm(ctypes.byref(sm)))
    assert sm.value >= 100, f"expected a Blackwell device, got sm_{sm.value}"
    assert _lib.launch_count() > 0
