# round 2, call 2: new GPU tests (input path, loss, parity additions) then the whole GPU suite
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_input_path.py tests/test_loss_full.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r02b_new_tests.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02b_pytest.txt
cat gpurun_out/r02b_new_tests.txt gpurun_out/r02b_pytest.txt
