timeout 300 python -m pytest tests/test_gpu_gemm.py -q 2>&1 | tail -3
timeout 200 python tools/quick_bench.py bcql 256 1000 2>&1 | cut -c1-120
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/launches_r01_thin4.csv python tools/quick_bench.py bcql 256 3 > gpurun_out/l.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 200 python tools/quick_bench_cdt.py 2>&1 | tail -9 | cut -c1-160
