#!/bin/bash
# Round 2, call 28: probe — extraction of step s+1 (context A) overlapped with the pair stage of step s (context B)
timeout 300 python tools/pipe2.py 592 8 2>&1 | tail -2
timeout 300 python tools/pipe2.py 1184 6 2>&1 | tail -2
