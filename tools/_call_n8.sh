set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561"
COMMON="--order mlp,qk,vo --calibs_batch_size 16 --nystrom_ridge 1e-4 --ridge_vo 1e-5 --ridge_qk 1e-2 --sparsity_smoothing 0.04948 --max_sparsity 0.95 --dataset synthetic"
timeout 600 $TR bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/r2c33_bench_n8.json 2> gpurun_out/r2c33_bench_n8.err
timeout 600 $TR -m modegpt_b200.run_modegpt --model synthetic:qwen3-8b --compression_ratio 0.30 --calib_size 128 $COMMON --output_dir /tmp/o4 --temp_storage_dir /tmp/o4/layers/ > gpurun_out/r2c33_e2e_qwen3_8b_n8.log 2>&1
rm -rf /tmp/o4
python - <<'PY'
import json
raw=open('gpurun_out/r2c33_bench_n8.json').read()
d=json.loads(raw[raw.index('{"metric"'):])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['strong_scaling'], d['roofline']['frac'], d['timed_region']); print(json.dumps(d['compress']))
PY
tail -3 gpurun_out/r2c33_bench_n8.err
grep "start-up\|Baseline\|Compressed (PPL)" gpurun_out/r2c33_e2e_qwen3_8b_n8.log | tail -3
