python bench.py --workload train --steps 6 --warmup 3 > gpurun_out/train_1gpu.json 2> gpurun_out/train_1gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --workload train --gpus 2 --steps 6 --warmup 3 > gpurun_out/train_2gpu.json 2> gpurun_out/train_2gpu.err
tail -3 gpurun_out/train_1gpu.err gpurun_out/train_2gpu.err
python - <<'PY'
import json
for f in ['gpurun_out/train_1gpu.json','gpurun_out/train_2gpu.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d[k] for k in ['value','ms_per_step','n_gpus','allreduce','loss','clocks']})
        print({k:v['ms_per_step'] for k,v in list(d['breakdown'].items())[:8]})
    except Exception as e: print(f, 'ERR', e)
PY
