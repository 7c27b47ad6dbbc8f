set -x
for mode in "" "--serial_stages"; do
timeout 600 python -m modegpt_b200.run_modegpt --model synthetic:llama-2-7b --order mlp,qk,vo --compression_ratio 0.25 --calib_size 128 --calibs_batch_size 16 --nystrom_ridge 1e-4 --ridge_vo 1e-5 --ridge_qk 1e-2 --sparsity_smoothing 0.04948 --max_sparsity 0.95 --dataset synthetic --output_dir /tmp/e2e_out --temp_storage_dir /tmp/e2e_out/layers/ $mode > gpurun_out/c7_e2e_7b$mode.log 2>&1
grep "stages:\|calibration " gpurun_out/c7_e2e_7b$mode.log
rm -rf /tmp/e2e_out
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_pair -c 4 -o gpurun_out/c7_trailing_full -f python tools/gpu_profile_type1.py > gpurun_out/c7_ncu_trailing.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/c7_type1_launches.csv python tools/gpu_profile_type1.py > gpurun_out/c7_ncu_type1.log 2>&1
tail -2 gpurun_out/c7_ncu_trailing.log
