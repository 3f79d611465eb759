python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --workload train --steps 8 --warmup 3 > gpurun_out/train_final_1gpu.json 2> gpurun_out/train_final.err
python bench.py --workload train --steps 8 --warmup 3 --drop-path 0 > gpurun_out/train_final_nodp_1gpu.json 2>> gpurun_out/train_final.err
python - <<'PY'
import json
for f in ['bench_final','train_final_1gpu','train_final_nodp_1gpu']:
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, round(d['value'],1), round(d['ms_per_step'],2), d.get('e2e',{}).get('value'), d['roofline']['achieved'], d['roofline'].get('frac'), d.get('cpu_baseline',{}).get('value'), d['clocks'])
PY
