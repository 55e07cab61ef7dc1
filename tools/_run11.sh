timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/launches_r01_final.csv python tools/quick_bench.py bcql 256 1 > gpurun_out/l.log 2>&1
tail -1 gpurun_out/l.log | cut -c1-100
timeout 900 ncu --set full --clock-control none -k regex:"k_gemm_thin|k_gemm_tc5|k_gemm_mma" -s 40 -c 40 -o /tmp/final_gemm -f python tools/quick_bench.py bcql 256 1 > gpurun_out/n.log 2>&1
ncu -i /tmp/final_gemm.ncu-rep --page raw --csv > gpurun_out/final_gemm_raw.csv
ls -la gpurun_out/final_gemm_raw.csv
