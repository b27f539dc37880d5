python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --workload cfg2 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
