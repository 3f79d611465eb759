timeout 300 python tools/gpu_diag_train.py attn_spatial_bwd > gpurun_out/diag_bwd_tc.txt 2>&1; tail -9 gpurun_out/diag_bwd_tc.txt | cut -c1-150
TCOW_SPATIAL_BWD_IMPL=mma timeout 300 python tools/gpu_diag_train.py attn_spatial_bwd_7 2>&1 | tail -2 | cut -c1-150
timeout 300 python tools/gpu_diag_train.py wgrad 2>&1 | tail -13 | cut -c1-150
timeout 600 python tools/train_bench.py --profile 2>&1 | grep -E "train step|attn_spatial|profiled|wgrad" | cut -c1-110
TCOW_WGRAD_CLUSTER=1 timeout 600 python tools/train_bench.py --profile 2>&1 | grep -E "train step|wgrad" | cut -c1-110
