timeout 200 python tools/gemm_bench.py tc5 2>&1 | grep debug_gemm
OPS=1 timeout 200 python tools/quick_bench_cdt.py 2>&1 | head -10 | cut -c1-140
timeout 200 python tools/quick_bench.py bcql 256 1000 2>&1 | cut -c1-120
