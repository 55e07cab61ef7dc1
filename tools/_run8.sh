timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "pipelined or sampled" 2>&1 | tail -12
for pl in 1 0; do echo "== PIPELINE=$pl"; OSRL_PIPELINE=$pl timeout 200 python tools/quick_bench.py bcql 256 1000 2>&1 | cut -c1-120; done
