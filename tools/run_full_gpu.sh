timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -2
python bench.py --steps 15 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], d['breakdown']['ln'])"
