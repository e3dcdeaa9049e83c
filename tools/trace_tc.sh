#!/bin/bash
# per-block timeline of the tensor-core stack kernel (CTA 0, blocks 1 and 2) for a few cases
export VQ3D_TC_TRACE=1 VQ3D_TC_DEBUG=1
for c in 3 4 2 1; do timeout 100 python tools/check_tc_stack.py $c --tc-only --no-graph 2>&1 | grep -v "regs=" | head -21; done
