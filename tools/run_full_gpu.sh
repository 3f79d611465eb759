timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu6.txt
python bench.py --workload sweep > gpurun_out/sweep_1gpu.json 2> gpurun_out/sweep_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 900 --csv --log-file gpurun_out/r01c_train_launches.csv python tools/train_bench.py --steps 1 --warmup 1 > gpurun_out/r01c_train_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'wgrad|attn_spatial_bwd|attn_temporal_bwd|layernorm_bwd|colsum' -s 40 -c 14 -o gpurun_out/r01c_train_full python tools/train_bench.py --steps 1 --warmup 1 > gpurun_out/r01c_train_ncu_full.log 2>&1
cat gpurun_out/pytest_gpu6.txt; tail -2 gpurun_out/sweep_1gpu.err; cut -c1-160 gpurun_out/sweep_1gpu.json; ls -la gpurun_out/r01c*
