"""Per-kernel parity of the training-step kernels vs fp32 PyTorch autograd on the GPU, through the C ABI."""
import pytest

from gpu_checks_train import TRAIN_CHECKS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,fn', TRAIN_CHECKS, ids=[c[0] for c in TRAIN_CHECKS])
def test_train_kernel(name, fn):
    err, tol, detail = fn()
    assert err <= tol, f'{detail}: err {err:.3e} > tol {tol:.3e}'
