# Every ray of 12 variant scenes and of cfg1 against the reference's (tools/compare_rays.py), one line per scene; the last
# command lists the differing rays of cfg1, if any, next to the nearest reference ray.  Run on the GPU box from the repo root
# (profiles/r10_compare_rays_all_identical.log is its output).
set -e
cd $GRAFT_REPO_ROOT
python - <<'PY'
import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from pathlib import Path
import test_variants_gpu as tv
out = Path('/tmp/crv'); out.mkdir(exist_ok=True)
for name in ['uber_full','substrate_wall','translucent_sheet','glass_rough','plastic_noremap','oren_nayar_rough','sphere_light','spot_light','infinite_light','strategy_power','disk_light','cylinder_instead_of_sphere']:
    (out / f'{name}.pbrt').write_text(tv.variant(name))
PY
for f in /tmp/crv/*.pbrt; do timeout 120 python tools/compare_rays.py $f 3 4 2>&1 | grep -v "^api.cpp\|FILMSTAT" | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['scene'], {k:(v['not_traced_by_the_reference'], v['rays']) for k,v in d.items() if isinstance(v,dict)})"; done
timeout 300 python tools/compare_rays.py oracle/_ref/scenes/cfg1.pbrt 3 4 2>&1 | grep -v "^api.cpp\|FILMSTAT" | tail -1 | cut -c1-500
COMPARE_RAYS_EXPLAIN=40 timeout 300 python tools/compare_rays.py oracle/_ref/scenes/cfg1.pbrt 3 1 2>&1 | grep "^DIFF" | cut -c1-330
