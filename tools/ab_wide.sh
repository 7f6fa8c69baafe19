#!/bin/bash
# A/B of builds of the wide-kernel path on one B200: _ab/libfmgpu_<name>.so variants; parity (tests/test_gpu_wide.py and the
# benchmarked configurations) for those named in $PARITY, then interleaved timing at k = 97 (1080p) and k = 193 (4K).
# A variant may carry an environment setting: name:VAR=value (e.g. htma:FM_WIDE_GP=3).
O=gpurun_out
cp find_motion_b200/libfmgpu.so /tmp/libfmgpu_orig.so
use() { local v=$1; local name=${v%%:*}; cp _ab/libfmgpu_$name.so find_motion_b200/libfmgpu.so; touch find_motion_b200/libfmgpu.so
        unset FM_WIDE_GP FM_WIDE_TMA_KB; if [[ $v == *:* ]]; then export "${v#*:}"; fi; }
for v in ${PARITY-"$@"}; do
  use $v
  echo "== $v parity"; timeout 600 python -m pytest tests/test_gpu_wide.py tests/test_gpu_benchmarked.py tests/test_gpu_fullsize.py -q -x --timeout 200 2>&1 | tail -2
done
for r in 1 2; do
  for v in "$@"; do
    use $v
    timeout 120 python bench.py --steps 10 --warmup 3 --min-seconds 1 --no-e2e --no-cpu-baseline --no-extras --blur-scale 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', $r, 'k97', d['value'], d['roofline']['groups_ms_per_step'])"
    [ -n "${NO4K-}" ] || timeout 120 python bench.py --steps 5 --warmup 3 --min-seconds 1 --no-e2e --no-cpu-baseline --no-extras --size 3840x2160 --streams 1 --frames 16 --ring 16 --blur-scale 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', $r, '4k-k193', d['value'], d['roofline']['groups_ms_per_step'])"
  done
done | tee $O/ab_wide.log
cp /tmp/libfmgpu_orig.so find_motion_b200/libfmgpu.so
