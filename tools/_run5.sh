timeout 200 python tools/gemm_bench.py tc5 2>&1 | grep debug_gemm | awk '{print $3, $4, $5, $6, $7}' | grep -v " 6.1"
timeout 300 python -m pytest tests/test_gpu_gemm.py -q 2>&1 | tail -2
timeout 200 python tools/quick_bench.py bcql 256 1000 2>&1 | cut -c1-120
