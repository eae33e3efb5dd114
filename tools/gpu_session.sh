#!/bin/bash
# One gpurun call: GPU parity tests, bench (both arms), ncu launch lists and ncu --set full captures.
# Usage (from the repo root on the GPU box): bash tools/gpu_session.sh <tag> [steps...]
tag=${1:-rX}; shift
steps=${@:-tests bench ref launches full}
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/box.txt; nproc >> $out/box.txt
for s in $steps; do case $s in
tests) timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -5 $out/pytest_gpu.log;;
smoke) timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; tail -2 $out/smoke.log;;
bench) timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; tail -c 600 $out/bench_n1.json;;
ref)   timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2>&1; cat $out/bench_reference.json;;
launches)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/cast_launches.csv python bench.py --steps 2 --warmup 3 --no-render > $out/ncu_cast.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/render_cfg1_launches.csv python tools/render_once.py cfg1 16 > $out/ncu_render.log 2>&1;;
full)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:closestHitKernel -s 3 -c 1 -o $out/closestHit python bench.py --steps 1 --warmup 3 --no-render > $out/ncu_full_cast.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 1 -c 2 -o $out/shade python tools/render_once.py cfg1 16 > $out/ncu_full_shade.log 2>&1;;
esac; done
ls -la $out
