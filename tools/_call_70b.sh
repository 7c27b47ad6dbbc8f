set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571"
rm -rf /dev/shm/mg70; mkdir -p /dev/shm/mg70
( while true; do nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | tr '\n' ' ' >> gpurun_out/r2c34_mem.log; echo >> gpurun_out/r2c34_mem.log; sleep 10; done ) &
MON=$!
timeout 900 $TR -m modegpt_b200.run_modegpt --model synthetic:llama-2-70b --order mlp,qk,vo --compression_ratio 0.30 --calib_size 256 --calibs_batch_size 16 --nystrom_ridge 1e-4 --ridge_vo 1e-5 --ridge_qk 1e-2 --sparsity_smoothing 0.04948 --max_sparsity 0.95 --dataset synthetic --stream_layers --skip_rebuild --skip_baseline_ppl --output_dir /dev/shm/mg70 --temp_storage_dir /dev/shm/mg70/layers/ > gpurun_out/r2c34_e2e_llama2_70b_n8.log 2>&1
echo "exit $?" >> gpurun_out/r2c34_e2e_llama2_70b_n8.log
kill $MON
ls /dev/shm/mg70/layers | wc -l >> gpurun_out/r2c34_e2e_llama2_70b_n8.log; du -sh /dev/shm/mg70/layers >> gpurun_out/r2c34_e2e_llama2_70b_n8.log
rm -rf /dev/shm/mg70
grep -v "Layer\|target_layers" gpurun_out/r2c34_e2e_llama2_70b_n8.log | tail -16; tail -2 gpurun_out/r2c34_mem.log
