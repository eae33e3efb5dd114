"""Shared fixtures.  `-m "not gpu"` covers the oracle, the host front-end and the C-ABI
surface; `-m gpu` holds the parity tests proper (CUDA path vs oracle / reference)."""
import importlib.util
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running statistical parity test")


def load_product():
    name = "pbrt_v3_iile_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, ROOT / "pbrt-v3-iile_b200" / "__init__.py",
                                                  submodule_search_locations=[str(ROOT / "pbrt-v3-iile_b200")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def iile():
    import __graft_entry__ as ge
    ge.build()
    return load_product()


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def ref_scenes(oracle):
    """The reference's scene files (copied next to oracle/_ref by oracle/Makefile)."""
    d = oracle.REF_SCENES
    if not (d / "cfg1.pbrt").exists():
        pytest.skip("oracle/_ref/scenes missing (needs one `make -C oracle ref` where /root/reference exists)")
    return d


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    return tmp_path_factory.mktemp("iile")


@pytest.fixture(scope="session")
def cfg2_flat(iile, ref_scenes, workdir):
    """cfg-2 geometry (66,532 triangles), flattened by the product front-end (host only)."""
    s = iile.Scene.from_pbrt(ref_scenes / "cfg2.pbrt", upload=False)
    p = workdir / "cfg2.flat"
    s.save_flat(p)
    s.close()
    return p


@pytest.fixture(scope="session")
def cfg1_flat(iile, ref_scenes, workdir):
    s = iile.Scene.from_pbrt(ref_scenes / "cfg1.pbrt", upload=False)
    p = workdir / "cfg1.flat"
    s.save_flat(p)
    s.close()
    return p
