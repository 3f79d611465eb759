timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu5.txt
python bench.py --workload sweep > gpurun_out/sweep_1gpu.json 2> gpurun_out/sweep_1gpu.err
timeout 600 python tools/hires_bench.py > gpurun_out/hires1.json 2> gpurun_out/hires1.err
cat gpurun_out/pytest_gpu5.txt; tail -2 gpurun_out/sweep_1gpu.err; cut -c1-200 gpurun_out/sweep_1gpu.json; tail -2 gpurun_out/hires1.err; cat gpurun_out/hires1.json
