"""Per-kernel parity vs fp32 PyTorch on the GPU, through the C ABI (ctypes -> libtcow_b200.so)."""
import pytest

from gpu_checks import ALL_CHECKS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,fn', ALL_CHECKS, ids=[c[0] for c in ALL_CHECKS])
def test_kernel(name, fn):
    err, tol, detail = fn()
    assert err <= tol, f'{detail}: max abs err {err:.3e} > tol {tol:.3e}'
