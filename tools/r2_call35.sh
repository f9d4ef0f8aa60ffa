#!/bin/bash
# Round 2, call 35: source-level ncu capture of png_inflate_kernel (148 colour + 148 depth streams)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"png_inflate_kernel" -c 2 -o gpurun_out/r2_inflate -f python tools/tum_probe.py 148 > gpurun_out/r2_inflate_ncu.log 2>&1
tail -3 gpurun_out/r2_inflate_ncu.log | cut -c1-200
