timeout 300 python tools/gpu_diag_train.py ln_bwd_scaled 2>&1 | tail -4 | cut -c1-170
