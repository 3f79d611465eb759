# The driver's launch of the default bench line on N GPUs of one box (inference + the `train` sub-record with real NCCL
# ranks) and the hires workload: usage  gpurun --gpus 8 -- 'bash tools/run_8gpu.sh 8 r02z'
N=${1:-8}
T=${2:-r02z}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
$TR --master-port 29543 bench.py --workload hires --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_hires_${N}gpu.json 2> gpurun_out/${T}_hires_${N}gpu.err
for f in ${T}_bench_${N}gpu ${T}_hires_${N}gpu; do tail -n 1 gpurun_out/$f.err | cut -c1-160; python - $f <<'PY'
import json,sys
f='gpurun_out/%s.json'%sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:d.get(k) for k in ['metric','value','ms_per_step','n_gpus']}, 'e2e', d.get('e2e',{}).get('value'))
    t=d.get('train')
    if t: print('   train', t['value'], t['ms_per_step'], t['allreduce'], {k:v['ms_per_step'] for k,v in t['breakdown'].items() if k.startswith('wgrad')})
except Exception as e: print(f,'ERR',e)
PY
done
