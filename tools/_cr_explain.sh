set -e
cd $GRAFT_REPO_ROOT
python - <<'PY'
import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from pathlib import Path
import test_variants_gpu as tv
out = Path('/tmp/crv'); out.mkdir(exist_ok=True)
(out / 'substrate_wall.pbrt').write_text(tv.variant('substrate_wall'))
PY
export COMPARE_RAYS_EXPLAIN=60
timeout 120 python tools/compare_rays.py /tmp/crv/substrate_wall.pbrt 3 4 2>&1 | grep -v "^api.cpp\|FILMSTAT"
timeout 300 python tools/compare_rays.py oracle/_ref/scenes/cfg1.pbrt 3 1 2>&1 | grep -v "^api.cpp\|FILMSTAT"
