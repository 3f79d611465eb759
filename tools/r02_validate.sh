# Round-2 validation on one B200: the whole GPU suite, smoke(), the default bench line, the ncu launch list with DRAM bytes
# (-> profiles/kernel_traffic.json via tools/ncu_traffic.py) and one `ncu --set full` capture of the kernels new this round.
set -x
mkdir -p gpurun_out
T=${1:-r02m}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/${T}_pytest.txt
cat gpurun_out/${T}_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; tail -2 gpurun_out/${T}_smoke.txt
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
python bench.py --workload hires --steps 10 > gpurun_out/${T}_hires.json 2> gpurun_out/${T}_hires.err
TCOW_CUDA_GRAPH=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-train \
  > gpurun_out/${T}_ncu_bench.log 2>&1
TCOW_CUDA_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:'attn_spatial_r1|patch_embed_fused|gemm_bf16_tn|attn_temporal|layernorm' \
  -s 160 -c 14 -o gpurun_out/${T}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-train \
  > gpurun_out/${T}_ncu_full.log 2>&1
ls -la gpurun_out/${T}_*
