for grp in gemm_add_scaled scale_rows; do timeout 300 python tools/gpu_diag_train.py $grp 2>&1 | tail -12; done > gpurun_out/diag_train3.txt 2>&1
timeout 900 python -m pytest tests/test_train_gpu.py -q 2>&1 | tail -40 > gpurun_out/train_gpu3.txt
cat gpurun_out/diag_train3.txt gpurun_out/train_gpu3.txt
