#!/bin/bash
# One gpurun call: GPU parity tests, bench (both arms), ncu launch lists and ncu --set full captures.
# Usage (from the repo root on the GPU box): bash tools/gpu_session.sh <tag> [steps...]
tag=${1:-rX}; shift
steps=${@:-tests bench ref launches full}
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/box.txt; nproc >> $out/box.txt
for s in $steps; do case $s in
tests) timeout 1800 python -m pytest tests -m gpu -q -s --durations=15 > $out/pytest_gpu.log 2>&1; tail -25 $out/pytest_gpu.log;;
smoke) timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; tail -2 $out/smoke.log;;
bench) timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; tail -c 600 $out/bench_n1.json;;
ref)   timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2>&1; cat $out/bench_reference.json;;
launches)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/cast_launches.csv python bench.py --workload cfg2 --legs none --no-cpu-baseline --steps 2 --warmup 3 > $out/ncu_cast.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/render_cfg1_launches.csv python tools/render_once.py cfg1 16 > $out/ncu_render.log 2>&1;;
full)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:closestHitKernel -s 3 -c 1 -o $out/closestHit python bench.py --workload cfg2 --legs none --no-cpu-baseline --steps 1 --warmup 3 > $out/ncu_full_cast.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 1 -c 2 -o $out/shade python tools/render_once.py cfg1 16 > $out/ncu_full_shade.log 2>&1;;
terrain)
  timeout 900 python bench.py --workload terrain --legs none --no-cpu-baseline > $out/bench_terrain_n1.json 2> $out/bench_terrain.err; tail -c 900 $out/bench_terrain_n1.json
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:closestHitKernel -s 3 -c 1 -o $out/closestHit_terrain python bench.py --workload terrain --legs none --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_full_terrain.log 2>&1;;
render35)
  timeout 900 python bench_render.py --workload cfg3 --cpu-spp 4 > $out/render_cfg3_n1.json 2> $out/render_cfg3.err; tail -c 700 $out/render_cfg3_n1.json
  timeout 1200 python bench_render.py --workload cfg5 --cpu-spp 1 > $out/render_cfg5_n1.json 2> $out/render_cfg5.err; tail -c 700 $out/render_cfg5_n1.json;;
launch35)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/render_cfg3_launches.csv python tools/render_once.py cfg3 4 1 > $out/ncu_render3.log 2>&1
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/render_cfg5_launches.csv python tools/render_once.py cfg5 1 1 > $out/ncu_render5.log 2>&1;;
shade3)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 1 -c 2 -o $out/shade_cfg3 python tools/render_once.py cfg3 4 1 > $out/ncu_full_shade3.log 2>&1;;
rfull)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'shadowKernel|misKernel|intersectKernel' -s 3 -c 6 -o $out/render_casts python tools/render_once.py cfg1 16 > $out/ncu_full_rcasts.log 2>&1;;
multi)
  n=${NGPU:-2}
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n > $out/bench_n$n.json 2> $out/bench_n$n.err; tail -c 400 $out/bench_n$n.json
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench_render.py --workload cfg5 --gpus $n > $out/render_cfg5_n$n.json 2> $out/render_cfg5_n$n.err; tail -c 400 $out/render_cfg5_n$n.json
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench_render.py --workload cfg3 --gpus $n > $out/render_cfg3_n$n.json 2> $out/render_cfg3_n$n.err; tail -c 400 $out/render_cfg3_n$n.json;;
full5)
  # one lane, so that the captured launch is a whole batch's (8.3 M paths): the first-bounce cast and the first shade
  IILE_RENDER_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:castKernel -s 1 -c 1 -o $out/cast_cfg5 python tools/render_once.py cfg5 1 1 > $out/ncu_full_cast5.log 2>&1
  IILE_RENDER_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 0 -c 1 -o $out/shade_cfg5 python tools/render_once.py cfg5 1 1 > $out/ncu_full_shade5.log 2>&1;;
dram5)
  # DRAM bytes of every castKernel launch of a 1-spp cfg5 render (profiles/traffic.json "castKernel@cfg5")
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:castKernel -c 40 --csv --log-file $out/cast_cfg5_dram.csv python tools/render_once.py cfg5 1 1 > $out/ncu_dram5.log 2>&1;;
iispt)
  timeout 600 python tools/time_iispt_task.py 256 8 > $out/iispt_task_256_t8.json 2> $out/iispt_task.err; cat $out/iispt_task_256_t8.json;;
hlbvh)
  timeout 600 python -m pytest tests/test_hlbvh_device_gpu.py -m gpu -q -s -k cfg5 > $out/hlbvh_cfg5.log 2>&1; grep "cfg5 hlbvh\|hlbvh(device)\|passed\|failed" $out/hlbvh_cfg5.log;;
ab)
  bash tools/ab_cast.sh $tag "${AB_LIBS:-lib}" "${AB_CAST:-cfg2 terrain}"
  bash tools/ab_render.sh $tag "${AB_LIBS:-lib}" "${AB_RENDER:-cfg1:16 cfg3:8 cfg5:4}";;
esac; done
ls -la $out
