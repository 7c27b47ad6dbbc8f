set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c15_pytest.log
timeout 900 python bench.py > gpurun_out/c15_bench.json 2> gpurun_out/c15_bench.err
timeout 600 python -m modegpt_b200.run_modegpt --model synthetic:llama-2-7b --order mlp,qk,vo --compression_ratio 0.25 --calib_size 128 --calibs_batch_size 16 --nystrom_ridge 1e-4 --ridge_vo 1e-5 --ridge_qk 1e-2 --sparsity_smoothing 0.04948 --max_sparsity 0.95 --dataset synthetic --output_dir /tmp/e2e_out --temp_storage_dir /tmp/e2e_out/layers/ > gpurun_out/c15_e2e_7b.log 2>&1
rm -rf /tmp/e2e_out
timeout 600 python -m modegpt_b200.run_modegpt --model synthetic:llama-2-7b --order mlp,qk,vo --compression_ratio 0.25 --calib_size 128 --calibs_batch_size 16 --nystrom_ridge 1e-4 --ridge_vo 1e-5 --ridge_qk 1e-2 --sparsity_smoothing 0.04948 --max_sparsity 0.95 --dataset synthetic --output_dir /tmp/e2e_out --temp_storage_dir /tmp/e2e_out/layers/ --mlp_workers 1 > gpurun_out/c15_e2e_7b_w1.log 2>&1
tail -3 gpurun_out/c15_pytest.log; cat gpurun_out/c15_bench.json; tail -3 gpurun_out/c15_bench.err; grep "stages:\|calibration " gpurun_out/c15_e2e_7b.log gpurun_out/c15_e2e_7b_w1.log
