for grp in gemm_dgelu ln_ colsum; do timeout 300 python tools/gpu_diag_train.py $grp 2>&1 | tail -12; done > gpurun_out/diag_train2.txt 2>&1
timeout 900 python -m pytest tests/test_train_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/train_gpu2.txt
timeout 600 python tools/train_bench.py --profile > gpurun_out/train_bench2.txt 2>&1
cat gpurun_out/diag_train2.txt gpurun_out/train_gpu2.txt; head -12 gpurun_out/train_bench2.txt
