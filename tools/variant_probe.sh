#!/bin/bash
# Probe: times bench.py (device-resident value, e2e, MLE kernel ms) with each liblsl variant under gpurun_variants/.
# Runs on the GPU box copy only; the in-tree library is restored afterwards.
cp lineslam_b200/liblsl_b200.so /tmp/lib_base.so
for v in base "$@"; do
  if [ "$v" = base ]; then cp /tmp/lib_base.so lineslam_b200/liblsl_b200.so; else cp gpurun_variants/lib_$v.so lineslam_b200/liblsl_b200.so; fi
  LSL_BENCH_NOCLOCKS=1 python bench.py --no-cpu --no-pipeline --unique 148 --steps 3 --warmup 3 $PROBE_FLAGS 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
k = d.get('kernel_ms_per_step', {})
print('$v', round(d['value'],1), round(d['e2e']['value'],1), {n: round(t,2) for n, t in k.items() if t > 3})
"
done
cp /tmp/lib_base.so lineslam_b200/liblsl_b200.so
