set -x
export TCOW_CUDA_GRAPH=0
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01e_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|attn_spatial_tc|attn_temporal|layernorm' -s 150 -c 12 -o gpurun_out/r01e_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r01e_ncu_full.log 2>&1
ls -la gpurun_out/r01e_launches.csv gpurun_out/r01e_full.ncu-rep
