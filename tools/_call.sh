set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c1_pytest.log
python tools/gpu_probe_band.py > gpurun_out/c1_band.log 2>&1
python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/c1_launches.csv python bench.py --steps 2 --warmup 1 --compress-layers 0 > gpurun_out/c1_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_pair -c 2 -o gpurun_out/c1_syrk_full -f python tools/gpu_profile_syrk.py > gpurun_out/c1_ncu_syrk.log 2>&1
python -m modegpt_b200.run_modegpt --model synthetic:llama-2-7b --order mlp,qk,vo --compression_ratio 0.25 --calib_size 128 --calibs_batch_size 16 --nystrom_ridge 1e-4 --ridge_vo 1e-5 --ridge_qk 1e-2 --sparsity_smoothing 0.04948 --max_sparsity 0.95 --dataset synthetic --output_dir /tmp/e2e_out --temp_storage_dir /tmp/e2e_out/layers/ > gpurun_out/c1_e2e_7b.log 2>&1
tail -3 gpurun_out/c1_pytest.log; cat gpurun_out/c1_band.log; cat gpurun_out/c1_bench.json; tail -5 gpurun_out/c1_e2e_7b.log
