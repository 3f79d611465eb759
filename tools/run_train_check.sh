timeout 900 python -m pytest tests/test_train_gpu.py -x -q -k "edge" 2>&1 | tail -15
