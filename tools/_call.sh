set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vo_factor -c 1 -o gpurun_out/c18_vo_full -f python tools/gpu_profile_type3.py > gpurun_out/c18_ncu_vo.log 2>&1
timeout 600 python -m modegpt_b200.run_modegpt --model synthetic:llama-2-7b --order mlp,qk,vo --compression_ratio 0.25 --calib_size 128 --calibs_batch_size 16 --nystrom_ridge 1e-4 --ridge_vo 1e-5 --ridge_qk 1e-2 --sparsity_smoothing 0.04948 --max_sparsity 0.95 --dataset synthetic --output_dir /tmp/e2e_out --temp_storage_dir /tmp/e2e_out/layers/ > gpurun_out/c18_e2e_7b.log 2>&1
grep "stages:\|calibration " gpurun_out/c18_e2e_7b.log; tail -2 gpurun_out/c18_ncu_vo.log
