#!/bin/bash
# default-mode A/B over library builds (_ab/libfmgpu_<name>.so) x chunk widths: usage tools/ab_resize_libs.sh "name:cx:dxu" ...
O=gpurun_out; out=$O/ab_resize_libs.log; : > $out
cp find_motion_b200/libfmgpu.so /tmp/libfmgpu_orig.so
B="python bench.py --mode default --steps 20 --warmup 3 --min-seconds 1 --no-e2e --no-cpu-baseline --no-extras"
for r in 1 2; do
for v in "$@"; do
  IFS=: read name cx dxu <<< "$v"
  cp _ab/libfmgpu_$name.so find_motion_b200/libfmgpu.so; touch find_motion_b200/libfmgpu.so
  FM_K0_CX=$cx FM_K0_DXU=$dxu timeout 200 $B 2>>$O/ab.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$v', $r, d['value'], d['roofline']['groups_ms_per_step'], d['clocks']['sm_mhz'])
" >> $out
done
done
cp /tmp/libfmgpu_orig.so find_motion_b200/libfmgpu.so
cat $out
