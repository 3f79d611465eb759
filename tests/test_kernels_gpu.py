"""Per-kernel parity vs fp32 PyTorch on the GPU, through the C ABI (ctypes -> libtcow_b200.so)."""
import pytest

from gpu_checks import ALL_CHECKS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,fn', ALL_CHECKS, ids=[c[0] for c in ALL_CHECKS])
def test_kernel(name, fn):
    err, tol, detail = fn()
    assert err <= tol, f'{detail}: max abs err {err:.3e} > tol {tol:.3e}'


def test_streamed_spatial_attention_serves_every_shape():
    """TCOW_SPATIAL_IMPL=stream routes EVERY frame size through the streamed kernel (attn_spatial_rs.cu), which normally only
    sees S > 304: all attn_spatial checks — one to ~1200 tokens per frame, one and several work units per worker — must pass
    on it too.  The switch is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = (
        "import sys\n"
        f"sys.path.insert(0, {here!r}); sys.path.insert(0, {os.path.dirname(here)!r})\n"
        "from gpu_checks import ALL_CHECKS\n"
        "bad = []\n"
        "for name, fn in ALL_CHECKS:\n"
        "    if name.startswith('attn_spatial'):\n"
        "        err, tol, detail = fn()\n"
        "        if not err <= tol:\n"
        "            bad.append((name, err, tol, detail))\n"
        "print('BAD', bad) if bad else print('OK')\n"
        "sys.exit(1 if bad else 0)\n")
    env = dict(os.environ, TCOW_SPATIAL_IMPL='stream')
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
