"""SURVEY.md §8(f) N3 on the device: TrowbridgeReitz material, Checkered / Perlin / Image textures (k_shade<.., FULL>)
through the C ABI vs the oracle's restatement (materials/trowbridge_reitz.rs, bxdfs/trowbridge_reitz*.rs,
textures/mod.rs). Same counter-based RNG on both sides -> same paths; differences are libm rounding only.
Tolerances are per-channel RMSE on linear radiance."""
import numpy as np
import pytest

from test_gpu_render import render_both, rmse

pytestmark = pytest.mark.gpu


def showcase_scene(ptb, with_light=True, sky="lerp", sampler_res=(16, 8)):
    s = ptb.HostScene()
    t_chk = s.add_texture(ptb.TEX_CHECKERED, (0.8, 0.8, 0.8), (0.2, 0.3, 0.2))
    t_gold = s.add_texture(ptb.TEX_SOLID, (1.0, 0.78, 0.34))
    t_white = s.add_texture(ptb.TEX_SOLID, (1, 1, 1))
    t_noise = s.add_perlin_texture(seed=3)
    yy, xx = np.mgrid[0:32, 0:64]
    env = np.stack([0.3 + 0.7 * (xx / 63.0), 0.4 + 0.3 * np.sin(yy / 5.0) ** 2, 1.0 - 0.8 * (yy / 31.0)], -1).astype(np.float32)
    env[4:8, 10:14] = 20.0   # a "sun"
    t_env = s.add_image_texture(env)
    t_lerp = s.add_texture(ptb.TEX_LERP, (0.5, 0.7, 1.0), (1, 1, 1))
    m_ground = s.add_material(ptb.MAT_LAMBERTIAN, t_chk, 0.5)
    m_rough = s.add_material(ptb.MAT_TROWBRIDGE_REITZ, t_gold, 0.5 * 0.5, ior=(1.5, 1.5, 1.5), metallic=1.0)
    m_smooth = s.add_material(ptb.MAT_TROWBRIDGE_REITZ, t_white, 0.15 * 0.15, ior=(1.8, 1.5, 1.3), metallic=0.0)
    m_marble = s.add_material(ptb.MAT_LAMBERTIAN, t_noise, 0.7)
    m_light = s.add_material(ptb.MAT_EMIT, t_white, 6.0)
    s.add_sphere((0, 1, -100.5), 100.0, m_ground)
    s.add_sphere((-1.1, 1, 0), 0.5, m_rough)
    s.add_sphere((0, 1, 0), 0.5, m_smooth)
    s.add_sphere((1.1, 1, 0), 0.5, m_marble)
    if with_light:
        s.add_sphere((0, 0.2, 1.6), 0.3, m_light)
    s.set_camera((0, -2.5, 0.6), (0, 1, 0), (0, 0, 1), 50)
    s.set_sky(t_env if sky == "image" else t_lerp, sampler_res)
    return s


@pytest.mark.parametrize("method", [0, 1])
@pytest.mark.parametrize("sky", ["lerp", "image"])
def test_showcase_matches_oracle(ptb, orc, gpu_ctx, method, sky):
    s = showcase_scene(ptb, sky=sky)
    gpu_ctx.stats_reset()
    g, o, st, counts = render_both(ptb, orc, gpu_ctx, s, 160, 90, 32, method)
    assert np.all(np.isfinite(g)) and g.mean() > 0.05
    assert rmse(g, o) < 3e-2, rmse(g, o)     # emitter + sun texels: high-variance samples, a few decision flips move energy
    assert abs(g.mean() - o.mean()) < 5e-3 + 0.02 * o.mean()
    assert st.rays_camera == counts["camera"]
    assert abs(int(st.rays_bounce) - counts["bounce"]) <= 3e-3 * counts["bounce"]


def test_trowbridge_reitz_only_scene_is_tight(ptb, orc, gpu_ctx):
    """No emitter, smooth sky: the only differences are last-ulp libm roundings inside the GGX terms."""
    s = showcase_scene(ptb, with_light=False, sky="lerp")
    for method in (0, 1):
        g, o, st, counts = render_both(ptb, orc, gpu_ctx, s, 128, 72, 32, method)
        assert rmse(g, o) < 5e-3, (method, rmse(g, o))
        assert abs(g.mean() - o.mean()) < 1e-3


def test_weak_white_furnace_on_device(ptb, gpu_ctx):
    """A metallic = 1, white-F0 GGX sphere inside a uniform white sky loses energy only through G2/G1 <= 1
    (trowbridge_reitz.rs:188-230 g2_test) and never gains any: radiance in (0, 1]."""
    s = ptb.HostScene()
    tw = s.add_texture(ptb.TEX_SOLID, (1, 1, 1))
    m = s.add_material(ptb.MAT_TROWBRIDGE_REITZ, tw, 0.4 * 0.4, ior=(1.5, 1.5, 1.5), metallic=1.0)
    s.add_sphere((0, 0, 0), 0.5, m)
    s.set_camera((0, 0, 3), (0, 0, 0), (0, 1, 0), 1e-4)
    s.set_sky(tw, (0, 0))
    sc = ptb.Scene(s, ctx=gpu_ctx)
    img = sc.render(ptb.RenderOptions(samples_per_pixel=256, render_method=0, width=64, height=36, seed=4))
    v = img.reshape(-1, 3).mean(axis=0)
    assert np.all(v <= 1.0 + 1e-3) and np.all(v > 0.85), v


def test_texture_data_errors(ptb, gpu_ctx):
    c = ptb.Context(0)
    s = showcase_scene(ptb)
    data = dict(s.texture_data)
    s.texture_data = {}
    c.upload(s)
    with pytest.raises(ptb.PtbError) as e:
        c.commit()                                   # image / perlin texture without its data
    assert e.value.code == 6
    s.texture_data = data
    c.upload(s)
    c.commit()
    c.close()


@pytest.mark.parametrize("method", [0, 1])
def test_material_sorted_shading_is_the_same_image(ptb, gpu_ctx, method, monkeypatch):
    """Window mode: a scene with more than two material kinds has each block of k_shade sort its hits by the kind of the
    surface hit before shading them (north_star: "shading on sorted ray queues"). Only the ORDER inside a block changes:
    same image, same counters with the sort forced off and on (the showcase scene enables it by itself: three kinds)."""
    s = showcase_scene(ptb, sky="image")
    sc = ptb.Scene(s, ctx=gpu_ctx)
    o = ptb.RenderOptions(samples_per_pixel=16, render_method=method, width=200, height=120, seed=12)
    out = []
    for flag in ("0", "1", None):
        if flag is None:
            monkeypatch.delenv("PTB_SHADE_SORT")
        else:
            monkeypatch.setenv("PTB_SHADE_SORT", flag)
        gpu_ctx.stats_reset()
        img = sc.render(o)
        st = gpu_ctx.stats()
        out.append((img, (st.rays_camera, st.rays_bounce, st.rays_shadow_light, st.rays_shadow_sky, st.rays_reference, st.paths)))
    assert out[0][1] == out[1][1] == out[2][1]
    assert np.allclose(out[0][0], out[1][0], rtol=1e-5, atol=1e-5) and np.allclose(out[0][0], out[2][0], rtol=1e-5, atol=1e-5)
