timeout 300 python tools/gpu_diag_train.py attn_spatial_bwd > gpurun_out/diag_bwd_tc.txt 2>&1
TCOW_SPATIAL_BWD_IMPL=mma timeout 300 python tools/gpu_diag_train.py attn_spatial_bwd_301 > gpurun_out/diag_bwd_mma.txt 2>&1
tail -12 gpurun_out/diag_bwd_tc.txt; tail -3 gpurun_out/diag_bwd_mma.txt
timeout 600 python tools/train_bench.py --profile 2>&1 | head -6
