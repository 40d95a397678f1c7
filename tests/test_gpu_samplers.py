"""The reference's chi-squared sampler tests on the DEVICE samplers: ptb_sample_only draws
directions with the functions k_shade calls, ptb_sampler_pdf evaluates the matching device pdf, and the harness is the
one the oracle tests use (tests/test_oracle_kats.py::_chi2_sphere — spherical_sampling.rs:64-226 + the cell pooling of
chi_squared.rs:6-70). Thresholds are the reference's: p > 0.01 / 10 (bxdfs/lambertian.rs:30-48,
trowbridge_reitz_vndf.rs:156-218), pdf mass within 5e-3 of 1.
"""
import numpy as np
import pytest

from test_oracle_kats import _chi2_sphere
from test_materials_textures import _to_world, _wi

pytestmark = pytest.mark.gpu


def test_query_layout(ptb):
    import ctypes
    assert ctypes.sizeof(ptb._lib.SamplerQuery) == 48 and ptb._lib.SamplerQuery.seed.offset == 40


def test_device_lambertian_chi_squared(ptb, orc, gpu_ctx):
    L = ptb._lib
    normal = orc.random_unit_vectors(1, seed=11)[0]
    dirs, pdf = gpu_ctx.sample_only(L.SAMPLER_LAMBERTIAN, 400_000, normal=normal, seed=21)
    assert np.allclose(np.linalg.norm(dirs, axis=1), 1, atol=1e-5)
    assert np.allclose(pdf, gpu_ctx.sampler_pdf(L.SAMPLER_LAMBERTIAN, dirs, normal=normal))          # sampler's pdf == pdf entry
    assert np.allclose(pdf, orc.lambertian_pdf(normal, dirs), rtol=1e-5, atol=1e-6)                   # == the oracle's pdf
    p, mass_err = _chi2_sphere(dirs, lambda d: gpu_ctx.sampler_pdf(L.SAMPLER_LAMBERTIAN, d, normal=normal))
    assert mass_err < 2e-3 and p > 0.01 / 10


def test_device_uniform_sphere_chi_squared(ptb, gpu_ctx):
    L = ptb._lib
    dirs, pdf = gpu_ctx.sample_only(L.SAMPLER_UNIFORM_SPHERE, 300_000, seed=5)
    assert np.allclose(np.linalg.norm(dirs, axis=1), 1, atol=1e-5) and np.allclose(pdf, 1 / (4 * np.pi))
    p, _ = _chi2_sphere(dirs, lambda d: gpu_ctx.sampler_pdf(L.SAMPLER_UNIFORM_SPHERE, d))
    assert p > 1e-3


@pytest.mark.parametrize("alpha,cos_i", [(0.15, 0.9), (0.5, 0.6), (0.9, 0.25)])
def test_device_tr_vndf_chi_squared(ptb, orc, gpu_ctx, alpha, cos_i):
    """trowbridge_reitz_vndf.rs:156-184 (isotropic_non_local): sample about a random normal, pdf integrates to 1."""
    L = ptb._lib
    normal = orc.random_unit_vectors(1, seed=17)[0]
    incoming = _to_world(normal, _wi(cos_i, 1.1))
    dirs, pdf = gpu_ctx.sample_only(L.SAMPLER_TR_VNDF, 300_000, alpha=alpha, normal=normal, aux=incoming, seed=5)
    assert np.allclose(np.linalg.norm(dirs, axis=1), 1, atol=1e-4)
    want = orc.tr_pdf(alpha, incoming, dirs, which=orc.TR_WORLD, normal=normal)
    ok = want > 1e-3
    assert np.allclose(pdf[ok], want[ok], rtol=2e-3)
    p, mass_err = _chi2_sphere(dirs, lambda d: gpu_ctx.sampler_pdf(L.SAMPLER_TR_VNDF, d, alpha=alpha, normal=normal, aux=incoming))
    assert mass_err < 5e-3 and p > 0.01 / 10


def test_device_sky_sampler_chi_squared(ptb, orc, gpu_ctx, rtweekend1):
    """sky.rs:43-78 on the device table (100 x 100 Distribution2D built at commit)."""
    L = ptb._lib
    gpu_ctx.upload(rtweekend1)
    gpu_ctx.commit()
    dirs, pdf = gpu_ctx.sample_only(L.SAMPLER_SKY, 400_000, seed=7)
    assert np.allclose(np.linalg.norm(dirs, axis=1), 1, atol=1e-5)
    o = orc.OracleScene(rtweekend1)
    # same table, same lookup; a direction within an ulp of a cell boundary may land in the neighbouring cell (acosf / atan2f)
    assert np.mean(~np.isclose(pdf, o.sky_pdf(dirs), rtol=1e-3, atol=1e-6)) < 1e-3
    p, mass_err = _chi2_sphere(dirs, lambda d: gpu_ctx.sampler_pdf(L.SAMPLER_SKY, d), n_theta=25, n_phi=50)
    assert mass_err < 5e-3 and p > 1e-3


def test_device_light_sampler_chi_squared(ptb, gpu_ctx, overshadowed):
    """Sphere light of scenes/overshadowed.ssml seen from a point outside it: cone sampling (sphere.rs:118-154) against
    the solid-angle pdf 1 / (2 pi (1 - cos theta_max)) inside the cone (sphere.rs:155-170), 0 outside."""
    L = ptb._lib
    gpu_ctx.upload(overshadowed)
    gpu_ctx.commit()
    light = overshadowed.spheres[overshadowed.materials["kind"][overshadowed.spheres["material"]] == ptb.MAT_EMIT][0]
    c, r = light["center"].astype(np.float64), float(light["radius"])
    point = (c + np.array([2.5, -1.5, 1.0]) * r).astype(np.float32)       # ~3 radii away: the cone is wide enough to bin
    normal = (c - point) / np.linalg.norm(c - point)
    dirs, pdf = gpu_ctx.sample_only(L.SAMPLER_LIGHT, 300_000, normal=normal, aux=point, light_index=0, seed=3)
    assert np.allclose(np.linalg.norm(dirs, axis=1), 1, atol=1e-5)
    dist = np.linalg.norm(c - point)
    cos_max = np.sqrt(1 - (r / dist) ** 2)
    inside = dirs @ normal >= cos_max - 1e-4
    assert inside.mean() > 0.999
    hit = pdf > 0
    assert hit.mean() > 0.99 and np.allclose(pdf[hit], 1 / (2 * np.pi * (1 - cos_max)), rtol=1e-3)
    # the pdf is constant on the cone and 0 outside (a step: the spherical harness's quadrature is wrong in the boundary
    # cells), so test uniformity in the cone's own coordinates: cos(theta) uniform on [cos_max, 1], phi uniform on [0, 2 pi)
    from scipy import stats
    x = np.cross(normal, [0.0, 0.0, 1.0]); x /= np.linalg.norm(x)
    y = np.cross(normal, x)
    d64 = dirs.astype(np.float64)
    u = np.clip((d64 @ normal - cos_max) / (1 - cos_max), 0, 1 - 1e-12)
    v = np.mod(np.arctan2(d64 @ y, d64 @ x), 2 * np.pi) / (2 * np.pi)
    counts, _, _ = np.histogram2d(u, np.clip(v, 0, 1 - 1e-12), bins=(20, 20), range=((0, 1), (0, 1)))
    assert stats.chisquare(counts.ravel()).pvalue > 1e-3
    # and the pdf entry point integrates to 1 over the sphere
    _, mass_err = _chi2_sphere(dirs, lambda dd: gpu_ctx.sampler_pdf(L.SAMPLER_LIGHT, dd, normal=normal, aux=point, light_index=0),
                               n_theta=120, n_phi=240)
    assert mass_err < 2e-2
    with pytest.raises(ptb.PtbError):
        gpu_ctx.sample_only(L.SAMPLER_LIGHT, 10, light_index=99)
