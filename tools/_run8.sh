timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "pipelined" 2>&1 | tail -12
for algo in cpq bearl; do for pl in 1 0; do echo "== $algo PIPELINE=$pl"; OSRL_PIPELINE=$pl timeout 200 python tools/quick_bench.py $algo 256 500 2>&1 | cut -c1-100; done; done
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
