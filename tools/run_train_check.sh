python tools/train_phases.py 2>&1 | tail -16
python bench.py --workload train --steps 8 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train', round(d['value'],1), round(d['ms_per_step'],2))"
