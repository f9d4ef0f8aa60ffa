#!/bin/bash
# Round 2, call 49: adopted occupancy settings — whole GPU suite, default bench line
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 400 python bench.py --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); k = d['kernel_ms_per_step']
print('cfg2 default', round(d['value'],1), round(d['e2e']['value'],1), d['host_ms_each_step']['value'])"
