#!/bin/bash
# sweep (th,tw) tile shapes of the tensor-core stack kernel: case index (tools/check_tc_stack.py) x tiles
export VQ3D_TC_DEBUG=1
run() { VQ3D_TC_TILE=$2 timeout 120 python tools/check_tc_stack.py $1 --tc-only 2>&1 | grep -v "^$\|regs=" | tail -2 | cut -c1-300 | tr '\n' ' '; echo; }
echo "== 18ch @128x128x32 n=4"; for t in 2,8 4,8 3,8 2,16 4,4 2,4 1,16 1,8; do run 1 $t; done
echo "== 72ch @32x32x8 n=4"; for t in 1,8 2,8 2,4 4,4 1,4; do run 2 $t; done
echo "== 8ch @256x256x64 n=3"; for t in 1,8 2,8 1,4 2,4 1,16 4,4; do run 6 $t; done
