#!/bin/bash
# Round 2, call 3: new parity tests (real images, border clip), pose_kernel thread/occupancy variants.
python -m pytest tests -m gpu -x -q -k "reference or border or golden" 2>&1 | tail -5
bash tools/variant_probe.sh pose128x4 pose192x2 pose128x3 2>&1 | tail -5
