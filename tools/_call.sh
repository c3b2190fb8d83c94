timeout 200 python tools/c1_bench.py 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-220
