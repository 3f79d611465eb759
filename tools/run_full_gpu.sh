timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu10.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke10.txt 2>&1
cat gpurun_out/pytest_gpu10.txt; tail -2 gpurun_out/smoke10.txt
