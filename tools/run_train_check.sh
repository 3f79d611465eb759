timeout 600 python -m pytest tests/test_train_gpu.py -x -q -k "fused_mask_loss or sgd or optimizer" 2>&1 | tail -12
python bench.py --workload train --steps 8 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train', round(d['value'],1), round(d['ms_per_step'],2))"
