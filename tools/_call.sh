set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c17_pytest.log
timeout 900 python bench.py > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c17_bench_ref.json 2> gpurun_out/c17_bench_ref.err
MG_PROFILE=1 timeout 300 python tools/gpu_probe_type1.py 2>&1 | tail -40 > gpurun_out/c17_type1_prof.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/c17_type1_launches.csv python tools/gpu_profile_type1.py > gpurun_out/c17_ncu_type1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c17_type3_launches.csv python tools/gpu_profile_type3.py > gpurun_out/c17_ncu_type3.log 2>&1
timeout 600 python -m modegpt_b200.run_modegpt --model synthetic:llama-2-7b --order mlp,qk,vo --compression_ratio 0.25 --calib_size 128 --calibs_batch_size 16 --nystrom_ridge 1e-4 --ridge_vo 1e-5 --ridge_qk 1e-2 --sparsity_smoothing 0.04948 --max_sparsity 0.95 --dataset synthetic --output_dir /tmp/e2e_out --temp_storage_dir /tmp/e2e_out/layers/ > gpurun_out/c17_e2e_7b.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c17_smoke.log 2>&1
tail -3 gpurun_out/c17_pytest.log; cat gpurun_out/c17_bench.json; cat gpurun_out/c17_bench_ref.json; tail -2 gpurun_out/c17_smoke.log; grep "stages:\|calibration " gpurun_out/c17_e2e_7b.log
