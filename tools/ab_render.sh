#!/bin/bash
# A/B of kernel-tuning builds on one box: tools/ab_render.sh <tag> "<lib dirs>" "<cfg:spp ...>"
# Each lib dir is pbrt-v3-iile_b200/<dir>/libiile_b200.so (built with csrc/Makefile OUT=../<dir> EXTRA=-D...).
tag=${1:-ab}; libs=${2:-"lib"}; cfgs=${3:-"cfg1:16 cfg1:256 cfg3:8 cfg5:2"}
out=gpurun_out/$tag; mkdir -p $out
for c in $cfgs; do for l in $libs; do
  name=${c%%:*}; spp=${c##*:}
  r=$(IILE_B200_LIB=$PWD/pbrt-v3-iile_b200/$l/libiile_b200.so timeout 600 python tools/render_once.py $name $spp 3 2>&1 | tail -1)
  echo "$l $c $r" | tee -a $out/ab.txt
done; done
