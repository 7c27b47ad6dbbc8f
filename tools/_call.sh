set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c3_pytest.log
timeout 300 python tools/gpu_probe_type1.py > gpurun_out/c3_type1.log 2>&1
MG_PROFILE=1 timeout 300 python tools/gpu_probe_type1.py 2>&1 | tail -45 > gpurun_out/c3_type1_prof.log
timeout 120 python tools/gpu_probe_clocks.py > gpurun_out/c3_clocks.log 2>&1
timeout 600 python -m modegpt_b200.run_modegpt --model synthetic:llama-2-7b --order mlp,qk,vo --compression_ratio 0.25 --calib_size 128 --calibs_batch_size 16 --nystrom_ridge 1e-4 --ridge_vo 1e-5 --ridge_qk 1e-2 --sparsity_smoothing 0.04948 --max_sparsity 0.95 --dataset synthetic --output_dir /tmp/e2e_out --temp_storage_dir /tmp/e2e_out/layers/ > gpurun_out/c3_e2e_7b.log 2>&1
ls -la /tmp/e2e_out/layers | head -5; ls /tmp/e2e_out/layers | wc -l
cat gpurun_out/c3_pytest.log; cat gpurun_out/c3_type1.log; cat gpurun_out/c3_clocks.log; grep -v "Compressed layer\|compressed to" gpurun_out/c3_e2e_7b.log | tail -12
