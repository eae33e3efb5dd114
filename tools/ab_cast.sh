#!/bin/bash
# A/B of traversal-kernel builds on one box: tools/ab_cast.sh <tag> "<lib dirs>" ["cfg2 terrain"]
# Each lib dir is pbrt-v3-iile_b200/<dir>/libiile_b200.so.  Prints Mrays/s, per-kernel ms and the
# bit-exactness flag of bench.py's oracle sample for every build.
tag=${1:-abc}; libs=${2:-"lib"}; wls=${3:-"cfg2 terrain"}
out=gpurun_out/$tag; mkdir -p $out
for w in $wls; do for l in $libs; do
  IILE_B200_LIB=$PWD/pbrt-v3-iile_b200/$l/libiile_b200.so timeout 600 python bench.py --workload $w --legs none --no-cpu-baseline --steps 10 > $out/cast_${w}_$l.json 2> $out/cast_${w}_$l.err
  python - "$out/cast_${w}_$l.json" "$l" "$w" <<'PY' | tee -a $out/ab_cast.txt
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d["roofline"]
    print(f"{sys.argv[2]:12s} {sys.argv[3]:8s} {d['value']:9.1f} Mrays/s closest {r['ms_per_launch']:.4f} ms any {r['any_hit']['ms_per_launch']:.4f} ms frac {r['frac']:.3f} parity {d['config']['parity_vs_oracle_on_sample']}")
except Exception as e:
    print(sys.argv[2], sys.argv[3], "FAILED", e)
PY
done; done
