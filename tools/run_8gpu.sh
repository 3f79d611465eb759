TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29541 bench.py --workload train --gpus 8 --steps 8 --warmup 3 > gpurun_out/train_8gpu_c.json 2> gpurun_out/train_8gpu_c.err
$TR --master-port 29542 bench.py --workload train --gpus 8 --steps 8 --warmup 3 --nccl-ctas 0 > gpurun_out/train_8gpu_d.json 2> gpurun_out/train_8gpu_d.err
$TR --master-port 29543 bench.py --workload hires --gpus 8 --steps 10 --warmup 3 > gpurun_out/hires_8gpu.json 2> gpurun_out/hires_8gpu.err
for f in train_8gpu_c train_8gpu_d hires_8gpu; do tail -n 1 gpurun_out/$f.err | cut -c1-160; python - $f <<'PY'
import json,sys
f='gpurun_out/%s.json'%sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(sys.argv[1], {k:d.get(k) for k in ['metric','value','ms_per_step','n_gpus','allreduce']}, d.get('e2e',{}).get('value'))
except Exception as e: print(f,'ERR',e)
PY
done
