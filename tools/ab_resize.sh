#!/bin/bash
# default-mode front end A/B on one B200: row-per-lane kernel (chunk widths / segment lengths) against the warp-per-row kernels
O=gpurun_out; out=$O/ab_resize.log; : > $out
[ -n "$SKIPTEST" ] || timeout 600 python -m pytest tests/test_gpu_resize.py tests/test_gpu_parity.py::test_golden_traces tests/test_gpu_benchmarked.py -x -q --timeout 300 -k "resize or golden or default" 2>&1 | tail -5 | tee -a $out
B="python bench.py --mode default --steps 20 --warmup 3 --min-seconds 1 --no-e2e --no-cpu-baseline --no-extras"
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$1', d['value'], d['roofline']['groups_ms_per_step'], d['clocks']['sm_mhz'])
"; }
timeout 200 $B --front-end warp-resize 2>>$O/ab.err | line "warp" >> $out
for v in ${VARIANTS:-1:20 2:20 3:21 2:10 2:40 4:20}; do
  cx=${v%%:*}; dxu=${v##*:}
  FM_K0_CX=$cx FM_K0_DXU=$dxu timeout 200 $B 2>>$O/ab.err | line "rows-cx$cx-dxu$dxu" >> $out
done
cat $out
