timeout 300 python tools/gpu_diag_train.py attn_temporal_bwd > gpurun_out/diag_tbwd.txt 2>&1; tail -8 gpurun_out/diag_tbwd.txt
timeout 1200 python -m pytest tests/test_train_gpu.py -x -q 2>&1 | tail -4
timeout 600 python tools/train_bench.py --profile 2>&1 | grep -E "train step|attn_temporal_bwd|profiled"
