#!/bin/bash
# A/B of k_fused builds on one B200: _ab/libfmgpu_<name>.so variants (built here with nvcc), each
# parity-checked (all, or those named in $PARITY) and then timed in interleaved rounds.  usage: tools/ab_k1.sh name1 name2 ...
O=gpurun_out
cp find_motion_b200/libfmgpu.so /tmp/libfmgpu_orig.so
for v in ${PARITY-"$@"}; do
  cp _ab/libfmgpu_$v.so find_motion_b200/libfmgpu.so; touch find_motion_b200/libfmgpu.so
  echo "== $v parity"; timeout 400 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_benchmarked.py -q -x --timeout 120 2>&1 | tail -2
done
for r in 1 2 3; do
  for v in "$@"; do
    cp _ab/libfmgpu_$v.so find_motion_b200/libfmgpu.so; touch find_motion_b200/libfmgpu.so
    timeout 120 python bench.py --steps 20 --warmup 3 --min-seconds 1 --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', $r, d['value'], d['roofline']['groups_ms_per_step'], d['clocks']['sm_mhz'])"
  done
done | tee $O/ab_k1.log
cp /tmp/libfmgpu_orig.so find_motion_b200/libfmgpu.so
