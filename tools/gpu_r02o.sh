#!/bin/bash
# r02: full GPU suite with fused2p as the default + ncu source capture of the Y pass (16 KiB pitch)
TAG=${1:-r02o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --import-source on --clock-control none -k regex:fused2p -s 1 -c 1 -o $OUT/fused2p_y python tools/axis_time.py --dims 64,2048,2048 --axes 2 --steps 2 > $OUT/ncu_y.log 2>&1
tail -2 $OUT/ncu_y.log
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest.txt
echo "== cfg5 per-axis"; timeout 300 python tools/axis_time.py --size 2048 --steps 3 --axes 1,2,4,7 2>&1 | cut -c1-260 | tee $OUT/axis_2048.txt
