timeout 200 python tools/quick_bench.py bcql 256 1000 2>&1 | cut -c1-120
timeout 200 python tools/profile_step.py bcql 256 2>&1 | grep -E "tc5|total" | head -5
