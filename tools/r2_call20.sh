#!/bin/bash
# Round 2, call 20: pose_kernel thread/occupancy variants, MLE size classes, batch 1184 (probe; output in gpurun_out/r2i_variants.log)
{
timeout 400 bash tools/variant_probe.sh pose128x4 pose128x3
echo "== LSL_MLE_SIZE_CLASSES=1"
LSL_MLE_SIZE_CLASSES=1 timeout 120 bash tools/variant_probe.sh
echo "== batch 1184"
timeout 200 python bench.py --no-cpu --steps 3 --warmup 3 --batch 1184 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
k = d.get('kernel_ms_per_step', {})
print('b1184', round(d['value'],1), round(d['e2e']['value'],1), {n: round(t,2) for n, t in k.items() if t > 3})
"
} 2>&1 | tee gpurun_out/r2i_variants.log
