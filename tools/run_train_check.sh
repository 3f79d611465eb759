TCOW_B200_LIB=$PWD/tcow_b200/libtcow_b200_w8.so timeout 300 python tools/gpu_diag_train.py attn_spatial_bwd_301 2>&1 | tail -4
TCOW_B200_LIB=$PWD/tcow_b200/libtcow_b200_w8.so timeout 600 python tools/train_bench.py --profile 2>&1 | head -5
