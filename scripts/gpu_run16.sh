timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/h2d_probe.py | head -2
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-others 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(round(d['value'],1), d['e2e'])"
