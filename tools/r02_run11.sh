# round 2, call 11: ticket-scheduled weight-gradient GEMM — checks, 1-GPU A/B (static vs tickets), gradient parity
set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag_train.py wgrad > gpurun_out/r02k_diag.txt 2>&1
grep -E "tickets|failures|EXC|timed" gpurun_out/r02k_diag.txt
for v in 0 1; do
TCOW_WGRAD_SCHED=$v timeout 600 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/r02k_train_$v.json 2> gpurun_out/r02k_train_$v.err
tail -2 gpurun_out/r02k_train_$v.err
python -c "
import json;d=json.loads(open('gpurun_out/r02k_train_$v.json').read().strip().splitlines()[-1])
print('SCHED=$v train', d['value'], d['ms_per_step'], {k:(v['ms_per_step'],v['tflops']) for k,v in d['breakdown'].items() if k.startswith('wgrad')})
"
done
TCOW_WGRAD_SCHED=1 timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -4
