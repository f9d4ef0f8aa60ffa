#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu --steps 5 --unique 48 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
k = d.get('kernel_ms_per_step', {})
print(round(d['value'],1), round(d['e2e']['value'],1), {n: round(t,2) for n, t in k.items() if t > 3})
"
