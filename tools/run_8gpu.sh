TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/infer_8gpu.json 2> gpurun_out/infer_8gpu.err
$TR --master-port 29522 bench.py --workload train --gpus 8 --steps 6 --warmup 3 > gpurun_out/train_8gpu.json 2> gpurun_out/train_8gpu.err
python bench.py --workload sweep --videos-total 16 > gpurun_out/sweep16_1gpu.json 2> gpurun_out/sweep16_1gpu.err
$TR --master-port 29523 bench.py --workload sweep --gpus 8 --videos-total 16 > gpurun_out/sweep16_8gpu.json 2> gpurun_out/sweep16_8gpu.err
for f in infer_8gpu train_8gpu sweep16_1gpu sweep16_8gpu; do tail -n 2 gpurun_out/$f.err | cut -c1-200; python - $f <<'PY'
import json,sys
f='gpurun_out/%s.json'%sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(sys.argv[1], {k:d.get(k) for k in ['metric','value','ms_per_step','n_gpus','allreduce','clocks']}, d.get('e2e',{}).get('value'))
except Exception as e: print(f,'ERR',e)
PY
done
