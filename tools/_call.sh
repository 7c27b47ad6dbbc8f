set -x
timeout 300 python tools/gpu_profile_type3.py > gpurun_out/c5_type3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c5_type3_launches.csv python tools/gpu_profile_type3.py > gpurun_out/c5_ncu_type3.log 2>&1
timeout 600 python -m modegpt_b200.run_modegpt --model synthetic:llama-2-7b --order mlp,qk,vo --compression_ratio 0.25 --calib_size 128 --calibs_batch_size 16 --nystrom_ridge 1e-4 --ridge_vo 1e-5 --ridge_qk 1e-2 --sparsity_smoothing 0.04948 --max_sparsity 0.95 --dataset synthetic --output_dir /tmp/e2e_out --temp_storage_dir /tmp/e2e_out/layers/ > gpurun_out/c5_e2e_7b.log 2>&1
cat gpurun_out/c5_type3.log; grep -v "Compressed layer\|compressed to" gpurun_out/c5_e2e_7b.log | tail -8
grep -n "MLP\] Layer 0 \|MLP\] Layer 31 \|QK\] Layer 0\|VO\] Compressed layer 0 \|VO\] Compressed layer 31" gpurun_out/c5_e2e_7b.log
