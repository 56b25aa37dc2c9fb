"""GPU parity suite (run on the B200 box: pytest -m gpu).  Every case calls the CUDA path
through the C ABI and checks it against the CPU oracle / torch fp32 math / reference goldens.
The case list is shared with tools/gpu_diag.py (crash-isolated bring-up runner)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import gpu_diag  # noqa: E402

_CASES = gpu_diag.registry()


@pytest.mark.gpu
@pytest.mark.parametrize("name,fn,kw", _CASES, ids=[c[0] for c in _CASES])
def test_gpu_case(name, fn, kw):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (no CPU fallback exists)"
    fn(**kw)


@pytest.mark.gpu
def test_native_library_is_loaded_and_counts_launches():
    import torch
    from leanyolo_b200 import _native as N
    import gpu_checks as G
    before = N.lib().ly_launch_count()
    G.check_up()
    assert N.lib().ly_launch_count() == before + 1
    loaded = [l for l in open("/proc/self/maps").read().splitlines() if "libleanyolo_b200.so" in l]
    assert loaded, "libleanyolo_b200.so is not mapped into the test process"
