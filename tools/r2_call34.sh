#!/bin/bash
# Round 2, call 34: smoke() with the SIFT leg, SIFT edge cases, full GPU suite
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
