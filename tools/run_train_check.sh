for H in base 200 2000 20000; do
  if [ $H = base ]; then unset TCOW_B200_LIB; else export TCOW_B200_LIB=$PWD/tcow_b200/libtcow_b200_h$H.so; fi
  echo "== hint $H"
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('infer', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:v['ms_per_step'] for k,v in d['breakdown'].items() if k in ('attn_spatial','gemm_qkv','gemm_fc1','gemm_fc2','gemm_proj','attn_temporal')})"
  timeout 600 python tools/train_bench.py --profile 2>&1 | grep -E "train step|attn_spatial" | cut -c1-110
done
