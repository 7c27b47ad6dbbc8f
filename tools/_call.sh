set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c11_pytest.log
timeout 300 python tools/gpu_probe_type1.py > gpurun_out/c11_type1.log 2>&1
MG_CHOL_OUTER=1 timeout 300 python tools/gpu_probe_type1.py > gpurun_out/c11_type1_g1.log 2>&1
MG_CHOL_OUTER=2 timeout 300 python tools/gpu_probe_type1.py > gpurun_out/c11_type1_g2.log 2>&1
MG_PROFILE=1 timeout 300 python tools/gpu_probe_type1.py 2>&1 | tail -32 > gpurun_out/c11_type1_prof.log
timeout 300 python tools/gpu_profile_type3.py > gpurun_out/c11_type3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_pair -c 6 -o gpurun_out/c11_trailing_full -f python tools/gpu_profile_type1.py > gpurun_out/c11_ncu_trailing.log 2>&1
tail -3 gpurun_out/c11_pytest.log; cat gpurun_out/c11_type1.log gpurun_out/c11_type1_g1.log gpurun_out/c11_type1_g2.log; cat gpurun_out/c11_type3.log; cat gpurun_out/c11_type1_prof.log
