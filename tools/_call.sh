set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c22_pytest.log
timeout 900 python bench.py > gpurun_out/c22_bench.json 2> gpurun_out/c22_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c22_smoke.log 2>&1
tail -3 gpurun_out/c22_pytest.log; cat gpurun_out/c22_bench.json; tail -2 gpurun_out/c22_smoke.log
