#!/bin/bash
# Runs every BASELINE.json configuration on one B200 and appends the bench lines to
# gpurun_out/configs.jsonl.  Usage (on the GPU box): bash profiles/run_configs.sh
set -u
OUT=gpurun_out/configs.jsonl
: > $OUT
run() { echo "# $*" >> $OUT; python bench.py "$@" --no-cpu-baseline --no-extras >> $OUT 2>> gpurun_out/configs.err; }
# cfg2: 1080p with masks: full-res k=5 (HBM regime), full-res k=97 (reference default blur scale), default mode
run --steps 20 --warmup 3
run --steps 10 --warmup 3 --blur-scale 20
run --steps 20 --warmup 3 --mode default
# cfg2 literally: ONE 1080p stream, T = 32
run --steps 20 --warmup 3 --streams 1 --frames 32 --ring 32
run --steps 20 --warmup 3 --streams 1 --frames 32 --ring 32 --mode default
# cfg4: one 4K stream, k = 193 and k = 385 (wide-halo stencil stress), and 4K default mode
run --steps 5 --warmup 3 --size 3840x2160 --streams 1 --frames 16 --ring 16 --blur-scale 20
run --steps 5 --warmup 3 --size 3840x2160 --streams 1 --frames 16 --ring 16 --blur-scale 10
run --steps 10 --warmup 3 --size 3840x2160 --streams 1 --frames 16 --ring 16 --blur-scale 768
run --steps 10 --warmup 3 --size 3840x2160 --streams 1 --frames 16 --ring 16 --mode default
# cfg5: 256 x 720p streams, k = 5, time-block sweep
for T in 1 2 4 8 16 32; do
  run --steps 10 --warmup 3 --size 1280x720 --streams 256 --distinct 8 --frames $T --ring 32 --blur-scale 256 --no-e2e
done
grep -c '^{' $OUT
