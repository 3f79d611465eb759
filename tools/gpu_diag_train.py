"""Run every per-kernel check and print a table; does not stop at the first failure.  For gpurun sessions."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import torch  # noqa: E402

from gpu_checks_train import TRAIN_CHECKS as ALL_CHECKS  # noqa: E402

only = sys.argv[1:]
fails = 0
for name, fn in ALL_CHECKS:
    if only and not any(o in name for o in only):
        continue
    t0 = time.time()
    try:
        err, tol, detail = fn()
        ok = err <= tol
        print(f"{'ok  ' if ok else 'FAIL'} {name:28s} err {err:.3e} tol {tol:.3e}  {detail}  [{time.time() - t0:.1f}s]", flush=True)
        fails += (not ok)
    except Exception as e:  # a CUDA error poisons the context: stop
        print(f'EXC  {name}: {e!r}', flush=True)
        traceback.print_exc()
        fails += 1
        if 'CUDA' in repr(e) or 'cuda' in repr(e):
            break
print('failures:', fails)
