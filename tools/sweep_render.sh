#!/bin/bash
# Sweep of run-time render tunables on one box: tools/sweep_render.sh <tag> "<cfg:spp ...>" "<ENV=VAL,ENV=VAL ...>"
# Every setting is a comma-separated list of environment assignments ("-" = defaults).
tag=${1:-sweep}; cfgs=${2:-"cfg1:16 cfg3:8 cfg5:4"}; settings=${3:-"-"}
out=gpurun_out/$tag; mkdir -p $out
for c in $cfgs; do for s in $settings; do
  name=${c%%:*}; spp=${c##*:}
  envs=$(echo "$s" | tr ',' ' '); [ "$s" = "-" ] && envs=""
  r=$(env $envs timeout 600 python tools/render_once.py $name $spp 3 2>&1 | tail -1)
  echo "$s $c $r" | tee -a $out/sweep.txt
done; done
