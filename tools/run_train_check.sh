timeout 900 python -m pytest tests/test_train_gpu.py -x -q 2>&1 | tail -30 > gpurun_out/train_gpu1.txt
timeout 600 python tools/train_bench.py --profile > gpurun_out/train_bench1.txt 2>&1
tail -3 gpurun_out/train_gpu1.txt; cat gpurun_out/train_bench1.txt
