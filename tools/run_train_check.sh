ncu --set full --clock-control none --import-source on -k regex:'attn_spatial_stream' -s 14 -c 1 -o gpurun_out/r01e_stream python tools/hires_bench.py 2 > gpurun_out/r01e_stream.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 160 --csv --log-file gpurun_out/r01e_hires_launches.csv python tools/hires_bench.py 2 > gpurun_out/r01e_hires_ncu.log 2>&1
ls -la gpurun_out/r01e_*
