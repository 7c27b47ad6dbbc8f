set -x
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c14_pytest.log
timeout 300 python tools/gpu_probe_type1.py > gpurun_out/c14_type1.log 2>&1
MG_PROFILE=1 timeout 300 python tools/gpu_probe_type1.py 2>&1 | tail -36 > gpurun_out/c14_type1_prof.log
tail -3 gpurun_out/c14_pytest.log; cat gpurun_out/c14_type1.log; cat gpurun_out/c14_type1_prof.log
