timeout 300 python tools/gpu_diag.py gelu fc1 2>&1 | tail -4 | cut -c1-160
for L in d8 d10 d8 d10; do
  if [ $L = d8 ]; then unset TCOW_B200_LIB; else export TCOW_B200_LIB=$PWD/tcow_b200/libtcow_b200_d10.so; fi
  python bench.py --steps 15 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$L', round(d['value'],1), d['breakdown']['gemm_fc1'])"
done
unset TCOW_B200_LIB
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -2
