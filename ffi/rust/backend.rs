//! `--backend cuda` for the frontend (un-compiled here). Lives next to src/scene.rs; `Scene::render` dispatches to it.
//! Flattens the concrete types named in src/parameters.rs:7-12 to the POD arrays of include/ptb200.h.
use crate::ptb200_sys::*;
use implementations::{rt_core::*, *};
use std::{collections::HashMap, ffi::CStr, ptr};

type Tex = AllTextures;
type Mat<'a> = AllMaterials<'a, Tex>;
type Prim<'a> = AllPrimitives<'a, Mat<'a>>;

fn v(p: Vec3) -> ptb_vec3 { ptb_vec3 { x: p.x, y: p.y, z: p.z } }

pub struct CudaScene { ctx: *mut ptb_ctx }

impl CudaScene {
    /// `primitives` in loader order (spheres first, then mesh triangles: loader/src/lib.rs:234-240).
    pub fn new(primitives: &[Prim], camera: &SimpleCamera, sky_tex: &Tex, sampler_res: (usize, usize), device: i32) -> Result<Self, String> {
        let mut ctx = ptr::null_mut();
        check(ptr::null_mut(), unsafe { ptb_create(device, &mut ctx) })?;
        // intern textures / materials by address: the arena (crates/region) keeps them alive and unique
        let (mut texs, mut mats) = (Vec::<ptb_texture>::new(), Vec::<ptb_material>::new());
        let (mut tex_ids, mut mat_ids) = (HashMap::<*const Tex, u32>::new(), HashMap::<*const Mat, u32>::new());
        let mut tex_data = Vec::<(u32, u32, u32, Vec<f32>)>::new(); // (texture, width, height, words) of image / perlin textures
        let mut tex_id = |t: &Tex| *tex_ids.entry(t as *const _).or_insert_with(|| {
            texs.push(flatten_texture(t));
            let id = texs.len() as u32 - 1;
            if let Some((w, h, words)) = texture_words(t) { tex_data.push((id, w, h, words)); }
            id
        });
        let sky_id = tex_id(sky_tex);
        let (mut spheres, mut tris) = (Vec::new(), Vec::new());
        for p in primitives {
            let m = match p { Prim::Sphere(s) => s.material, Prim::Triangle(t) => t.material, Prim::MeshTriangle(t) => t.material };
            let mid = *mat_ids.entry(m as *const _).or_insert_with(|| { mats.push(flatten_material(m, &mut tex_id)); mats.len() as u32 - 1 });
            match p {
                Prim::Sphere(s) => spheres.push(ptb_sphere { center: v(s.center), radius: s.radius, material: mid }),
                Prim::Triangle(t) => tris.push(ptb_triangle { p: t.points.map(v), n: t.normals.map(v), material: mid }),
                Prim::MeshTriangle(t) => tris.push(ptb_triangle {
                    p: t.point_indices.map(|i| v(t.mesh.vertices[i])), n: t.normal_indices.map(|i| v(t.mesh.normals[i])), material: mid }),
            }
        }
        let cam = ptb_camera { origin: v(camera.origin), lower_left: v(camera.lower_left), horizontal: v(camera.horizontal), vertical: v(camera.vertical) };
        let sky = ptb_sky { texture: sky_id, sampler_res_x: sampler_res.0 as u32, sampler_res_y: sampler_res.1 as u32 };
        unsafe {
            check(ctx, ptb_scene_set_textures(ctx, texs.as_ptr(), texs.len()))?;
            for (id, w, h, words) in &tex_data {
                check(ctx, ptb_scene_set_texture_data(ctx, *id, *w, *h, words.as_ptr(), words.len()))?;
            }
            check(ctx, ptb_scene_set_materials(ctx, mats.as_ptr(), mats.len()))?;
            check(ctx, ptb_scene_set_spheres(ctx, spheres.as_ptr(), spheres.len()))?;
            check(ctx, ptb_scene_set_triangles(ctx, tris.as_ptr(), tris.len()))?;
            check(ctx, ptb_scene_set_camera(ctx, &cam))?;
            check(ctx, ptb_scene_set_sky(ctx, &sky))?;
            check(ctx, ptb_scene_commit(ctx, 0))?; // Bvh::new
        }
        Ok(Self { ctx })
    }

    /// Scene::render (src/scene.rs:35-42) for the cuda backend: the accumulator comes back as the running-mean image the
    /// TUI closure of src/main.rs:175-191 would hold, ready for output::save_data_to_image.
    pub fn render(&self, opts: RenderOptions, seed: u64) -> Result<(Vec<Float>, u64), String> {
        let o = self.opts(opts, seed);
        let mut image = vec![0.0 as Float; (opts.width * opts.height * 3) as usize];
        let mut st = ptb_stats::default();
        unsafe {
            check(self.ctx, ptb_accum_clear(self.ctx))?;
            check(self.ctx, ptb_render(self.ctx, &o, None, ptr::null_mut()))?;
            check(self.ctx, ptb_accum_read(self.ctx, image.as_mut_ptr(), image.len(), 1))?;
            check(self.ctx, ptb_stats_get(self.ctx, &mut st))?;
        }
        Ok((image, st.rays_reference))
    }
}

impl CudaScene {
    /// The reference's own contract: `presentation_update` sees the single-sample image of every pass
    /// (Sampler::sample_image, samplers/mod.rs:7-20). `f` returns true to stop, like the Rust closure.
    pub fn render_with_update<T, F: Fn(&mut T, &SamplerProgress, u64) -> bool>(&self, opts: RenderOptions, seed: u64, data: &mut T, f: F) -> Result<(), String> {
        struct Thunk<'a, T, F> { data: &'a mut T, f: F, progress: SamplerProgress }
        unsafe extern "C" fn call<T, F: Fn(&mut T, &SamplerProgress, u64) -> bool>(user: *mut std::ffi::c_void, img: *const f32, n: usize, i: u64, rays: u64) -> i32 {
            let t = &mut *(user as *mut Thunk<T, F>);
            t.progress.current_image.copy_from_slice(std::slice::from_raw_parts(img, n));
            t.progress.rays_shot = rays;
            (t.f)(t.data, &t.progress, i) as i32
        }
        let mut thunk = Thunk { data, f, progress: SamplerProgress::new(opts.width * opts.height, 3) };
        let o = self.opts(opts, seed);
        unsafe {
            check(self.ctx, ptb_accum_clear(self.ctx))?;
            match ptb_render_passes(self.ctx, &o, Some(call::<T, F>), &mut thunk as *mut _ as *mut _) {
                7 /* PTB_ERR_ABORTED: the closure asked to stop */ => Ok(()),
                rc => check(self.ctx, rc),
            }
        }
    }
    fn opts(&self, opts: RenderOptions, seed: u64) -> ptb_render_opts {
        ptb_render_opts {
            width: opts.width as u32, height: opts.height as u32, samples_per_pixel: opts.samples_per_pixel as u32, sample_offset: 0,
            method: match opts.render_method { RenderMethod::Naive => PTB_METHOD_NAIVE, RenderMethod::MIS => PTB_METHOD_MIS },
            max_depth: 50, rr_threshold: PTB_RR_DEFAULT, flags: 0, seed, row_begin: 0, row_count: 0,
        }
    }
}

impl Drop for CudaScene { fn drop(&mut self) { unsafe { ptb_destroy(self.ctx); } } }

fn check(ctx: *mut ptb_ctx, rc: i32) -> Result<(), String> {
    if rc == PTB_OK { return Ok(()); }
    Err(unsafe { CStr::from_ptr(ptb_last_error(ctx)) }.to_string_lossy().into_owned())
}

fn flatten_texture(t: &Tex) -> ptb_texture {
    match t { // tag == enum order (textures/mod.rs:17-24)
        AllTextures::CheckeredTexture(c) => ptb_texture { kind: 0, a: v(c.colour_one), b: v(c.colour_two) },
        AllTextures::SolidColour(s) => ptb_texture { kind: 1, a: v(s.colour), b: v(Vec3::zero()) },
        AllTextures::Lerp(l) => ptb_texture { kind: 3, a: v(l.colour_one), b: v(l.colour_two) },
        AllTextures::ImageTexture(_) => ptb_texture { kind: 2, a: v(Vec3::zero()), b: v(Vec3::zero()) }, // pixels: texture_words
        AllTextures::Perlin(_) => ptb_texture { kind: 4, a: v(Vec3::zero()), b: v(Vec3::zero()) },       // tables: texture_words
    }
}

/// Bulk data for ptb_scene_set_texture_data. ImageTexture: `data` as width*height RGB f32 (dim holds width-1, height-1:
/// textures/mod.rs:232). Perlin: 256 ran_vecs[i].x, then perm_x | perm_y | perm_z as u32 bit patterns (needs the four
/// private fields of `Perlin` made pub(crate), textures/mod.rs:76-81).
fn texture_words(t: &Tex) -> Option<(u32, u32, Vec<f32>)> {
    match t {
        AllTextures::ImageTexture(i) => Some((i.dim.0 as u32 + 1, i.dim.1 as u32 + 1, i.data.iter().flat_map(|c| [c.x, c.y, c.z]).collect())),
        AllTextures::Perlin(p) => {
            let mut w: Vec<f32> = p.ran_vecs.iter().map(|r| r.x).collect();
            for perm in [&p.perm_x, &p.perm_y, &p.perm_z] { w.extend(perm.iter().map(|&i| f32::from_bits(i))); }
            Some((0, 0, w))
        }
        _ => None,
    }
}

fn flatten_material(m: &Mat, tex_id: &mut impl FnMut(&Tex) -> u32) -> ptb_material {
    let one = v(Vec3::one());
    match m { // tag == enum order (materials/mod.rs:19-25)
        AllMaterials::Emit(e) => ptb_material { kind: 0, texture: tex_id(e.texture), param: e.strength, ior: one, metallic: 0.0 },
        AllMaterials::Lambertian(l) => ptb_material { kind: 1, texture: tex_id(l.texture), param: l.albedo, ior: one, metallic: 0.0 },
        AllMaterials::TrowbridgeReitz(t) => ptb_material { kind: 2, texture: tex_id(t.texture), param: t.alpha, ior: v(t.ior), metallic: t.metallic },
        AllMaterials::Reflect(r) => ptb_material { kind: 3, texture: tex_id(r.texture), param: r.fuzz, ior: one, metallic: 0.0 },
        AllMaterials::Refract(r) => ptb_material { kind: 4, texture: tex_id(r.texture), param: r.eta, ior: one, metallic: 0.0 },
    }
}
