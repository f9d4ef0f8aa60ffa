#!/bin/bash
# Round 2, call 54: last check of the committed build — smoke(), whole GPU suite, default bench line
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2_last_bench_cfg2_n1.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/r2_last_bench_cfg2_n1.json').readline()); print(round(d['value']), round(d['e2e']['value']), d['clocks'], d['gpu_launches'], d['cpu_baseline']['value'])"
