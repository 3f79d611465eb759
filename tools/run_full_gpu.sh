timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu8.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke8.txt 2>&1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/pytest_gpu8.txt; tail -2 gpurun_out/smoke8.txt; python -c "
import json; d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ['value','ms_per_step','steps','warmup','gpu_launches']}, d['e2e']['value'], d['cpu_baseline'], d['roofline']['frac'], d['roofline']['traffic'])"
