#!/bin/bash
# sweep (th,tw) tile shapes of the tensor-core stack kernel: case index (tools/check_tc_stack.py) x tiles
export VQ3D_TC_DEBUG=1
run() { VQ3D_TC_TILE=$2 timeout 120 python tools/check_tc_stack.py $1 --tc-only 2>&1 | grep -v "^$" | tail -2 | cut -c1-300 | tr '\n' ' '; echo; }
echo "== 18ch @128x128x32 n=4"; for t in 2,8 4,8 8,8 2,16 4,4 4,16 1,16 2,32 1,32; do run 1 $t; done
echo "== 72ch @32x32x8 n=4"; for t in 1,8 2,8 4,8 4,4 2,4 2,16 1,16 1,32; do run 2 $t; done
echo "== 32ch @8x8x2 n=8"; for t in 8,8 4,8 4,4 2,4 2,2 1,8 1,2; do run 3 $t; done
echo "== 8ch @32x32x8 n=8"; for t in 2,8 4,8 8,8 2,4 1,8 4,16 1,4; do run 4 $t; done
