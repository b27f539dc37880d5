set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 3 --warmup 3 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --workload cfg2 --no-cpu-baseline 2>&1 | tail -1
python bench.py --steps 2 --warmup 3 --workload cfg5 --no-cpu-baseline 2>&1 | tail -1
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1
