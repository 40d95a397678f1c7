import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def root():
    return ROOT


@pytest.fixture(scope="session")
def ptb():
    import ptb200
    return ptb200


@pytest.fixture(scope="session")
def orc():
    import oracle
    return oracle


@pytest.fixture(scope="session")
def rtweekend1(ptb, root):
    return ptb.load_file(os.path.join(root, "scenes", "rtweekend1.ssml"))


@pytest.fixture(scope="session")
def overshadowed(ptb, root):
    return ptb.load_file(os.path.join(root, "scenes", "overshadowed.ssml"))


@pytest.fixture(scope="session")
def gpu_ctx(ptb):
    ctx = ptb.Context(0)
    yield ctx
    ctx.close()


def random_rays(ptb, n, seed, centre=(0, 0, 0), radius=2.0):
    return ptb.meshgen.philox_rays(n, seed=seed, centre=centre, radius=radius)


def furnace_scene(ptb, sampler_res=(0, 0)):
    """implementations/tests/sampling.rs:31-63 (furnace_test)."""
    f = ptb.HostScene()
    tw = f.add_texture(ptb.TEX_SOLID, (1, 1, 1))
    tg = f.add_texture(ptb.TEX_SOLID, (0.5, 0.5, 0.5))
    tm = f.add_texture(ptb.TEX_SOLID, (1, 0, 1))
    sky_t = f.add_texture(ptb.TEX_LERP, (0, 0, 0), (0.5, 1.0, 0.2))
    ml = f.add_material(ptb.MAT_EMIT, tw, 1.0)
    mg = f.add_material(ptb.MAT_LAMBERTIAN, tg, 0.5)
    mh = f.add_material(ptb.MAT_EMIT, tm, 15.0)
    f.add_sphere((0, 0, 0), 0.5, mg)
    f.add_sphere((0, 0, 0), 1000.0, ml)
    f.add_sphere((0, 0, -5), 0.45, mh)
    f.set_camera((0, 0, 3), (0, 0, 0), (0, 1, 0), 40)
    f.set_sky(sky_t, sampler_res)
    return f
