timeout 300 python tools/gpu_diag.py attn_spatial > gpurun_out/diag_stream1.txt 2>&1
TCOW_SPATIAL_IMPL=stream timeout 300 python tools/gpu_diag.py attn_spatial > gpurun_out/diag_stream2.txt 2>&1
timeout 600 python -m pytest tests/test_parity_gpu.py -q -k "long_ or sweep or errors" 2>&1 | tail -5 > gpurun_out/pytest_stream.txt
timeout 600 python tools/hires_bench.py > gpurun_out/hires2.json 2> gpurun_out/hires2.err
python bench.py --workload sweep > gpurun_out/sweep_1gpu.json 2> gpurun_out/sweep_1gpu.err
grep -v "^ok" gpurun_out/diag_stream1.txt | tail -5; grep -c "^ok" gpurun_out/diag_stream1.txt; grep -v "^ok" gpurun_out/diag_stream2.txt | tail -5; grep -c "^ok" gpurun_out/diag_stream2.txt
cat gpurun_out/pytest_stream.txt; tail -2 gpurun_out/hires2.err; cat gpurun_out/hires2.json; cut -c1-160 gpurun_out/sweep_1gpu.json
