# round 2, call 1: GPU tests + default bench line + hires bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_pytest.txt
python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -5 gpurun_out/r02a_bench.err
python bench.py --workload hires --steps 10 > gpurun_out/r02a_hires.json 2> gpurun_out/r02a_hires.err
tail -3 gpurun_out/r02a_hires.err
cat gpurun_out/r02a_pytest.txt
