/*
 * ptb200.h — C ABI of the B200-native path-tracing backend ("ptb200") for
 * nonl4331/raytracing-rust.
 *
 * This header is the drop-in boundary: it is exactly what a Rust `build.rs` +
 * `extern "C"` shim (see ffi/rust/) binds when the frontend is
 * started with `--backend cuda`.  All citations are file:line into the reference
 * tree (crates/<name>/src/... abbreviated as <name>/...).
 *
 * Conventions
 *   - every entry point returns an int32 status (PTB_OK == 0); no C++ exception
 *     or Rust panic ever crosses the boundary; ptb_last_error() gives the text;
 *   - the caller owns all host buffers; the library copies on upload;
 *   - one ptb_ctx per GPU; a ctx is not thread-safe, distinct ctxs are independent;
 *   - plain pointers and sizes only; every struct below is #[repr(C)]-mirrorable POD;
 *   - primitive ids are indices into the reference's primitive array order:
 *     all spheres first, then all mesh triangles (loader/lib.rs:234-240);
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry
 *     point fails with PTB_ERR_CUDA.
 */
#ifndef PTB200_H
#define PTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTB_ABI_VERSION 2u  /* 2: ptb_render_opts gained row_begin / row_count (image-tile axis of the multi-GPU split) */

/* ---------------------------------------------------------------- status -- */
enum {
  PTB_OK = 0,
  PTB_ERR_INVALID = 1,   /* bad argument / bad state (e.g. render before commit)   */
  PTB_ERR_CUDA = 2,      /* CUDA runtime error or no device                          */
  PTB_ERR_OOM = 3,       /* host or device allocation failed                         */
  PTB_ERR_PARSE = 4,     /* .ssml / .obj syntax error     (loader LoadErr::ParseError, lib.rs:180-194) */
  PTB_ERR_IO = 5,        /* file could not be read/written (LoadErr::FileNotRead)    */
  PTB_ERR_MISSING = 6,   /* required key/object missing   (LoadErr::MissingRequired, MissingCamera) */
  PTB_ERR_ABORTED = 7,   /* progress callback asked to stop (random_sampler.rs:84-86) */
  PTB_ERR_UNSUPPORTED = 8
};

/* ------------------------------------------------------------------- POD -- */
/* rt_core/vec.rs:108-114 — `#[repr(C)] struct Vec3 { x, y, z }` with Float = f32 (rt_core/lib.rs:23-29) */
typedef struct ptb_vec3 { float x, y, z; } ptb_vec3;

/* implementations/primitives/sphere.rs:8-13 (material reference -> index into the material array) */
typedef struct ptb_sphere {
  ptb_vec3 center;
  float    radius;
  uint32_t material;
} ptb_sphere;

/* implementations/primitives/triangle.rs:9-14 / :30-36 — Triangle and MeshTriangle both flatten to this:
 * three positions, three vertex normals (MeshData indirection resolved by the host), one material. */
typedef struct ptb_triangle {
  ptb_vec3 p[3];
  ptb_vec3 n[3];
  uint32_t material;
} ptb_triangle;

/* enum order == tag order: implementations/materials/mod.rs:19-25 */
enum {
  PTB_MAT_EMIT = 0,             /* emissive.rs:7-10      param = strength */
  PTB_MAT_LAMBERTIAN = 1,       /* lambertian.rs:6-9     param = albedo   */
  PTB_MAT_TROWBRIDGE_REITZ = 2, /* trowbridge_reitz.rs:6-11  param = alpha (already squared), ior, metallic */
  PTB_MAT_REFLECT = 3,          /* reflect.rs:8-11       param = fuzz     */
  PTB_MAT_REFRACT = 4           /* refract.rs:9-12       param = eta      */
};
typedef struct ptb_material {
  uint32_t kind;
  uint32_t texture;   /* index into the texture array */
  float    param;
  ptb_vec3 ior;       /* TrowbridgeReitz only */
  float    metallic;  /* TrowbridgeReitz only */
} ptb_material;

/* enum order == tag order: implementations/textures/mod.rs:17-24 */
enum {
  PTB_TEX_CHECKERED = 0, /* textures/mod.rs:26-30,61-73   a = colour_one, b = colour_two */
  PTB_TEX_SOLID = 1,     /* textures/mod.rs:182-200       a = colour                     */
  PTB_TEX_IMAGE = 2,     /* textures/mod.rs:202-266       pixels via ptb_scene_set_texture_data  */
  PTB_TEX_LERP = 3,      /* textures/mod.rs:268-291       a = colour_one, b = colour_two */
  PTB_TEX_PERLIN = 4     /* textures/mod.rs:75-180        tables via ptb_scene_set_texture_data  */
};
#define PTB_PERLIN_TABLE_WORDS 1024u /* 256 f32 ran_vecs[i].x (every ran_vec is r*(1,1,1), textures/mod.rs:96-99)
                                        then perm_x, perm_y, perm_z as 3 x 256 uint32 bit patterns */
typedef struct ptb_texture {
  uint32_t kind;
  ptb_vec3 a;
  ptb_vec3 b;
} ptb_texture;

/* implementations/camera.rs:6-17 — only the four vectors get_ray() reads (camera.rs:57-63) */
typedef struct ptb_camera {
  ptb_vec3 origin;
  ptb_vec3 lower_left;
  ptb_vec3 horizontal;
  ptb_vec3 vertical;
} ptb_camera;

/* implementations/sky.rs:12-18 — material is always Emit(texture, 1.0) (loader/misc.rs:26);
 * sampler_res (0,0) disables importance sampling (sky.rs:61-63). The library builds the
 * Distribution2D table itself (sky.rs:20-37, textures/mod.rs:32-50, distributions.rs:82-99). */
typedef struct ptb_sky {
  uint32_t texture;
  uint32_t sampler_res_x;
  uint32_t sampler_res_y;
} ptb_sky;

/* rt_core/ray.rs:4-11 — the caller passes origin + (un-normalised) direction; the library derives
 * direction/|direction|, d_inverse and shear exactly as Ray::new does (ray.rs:13-46). 2 x 16 B. */
typedef struct ptb_ray {
  float ox, oy, oz, _pad0;
  float dx, dy, dz, _pad1;
} ptb_ray;

/* what AccelerationStructure::check_hit returns (implementations/acceleration/mod.rs:265-298), reduced
 * to the comparable part: t, primitive id in the ORIGINAL (pre-BVH) order, barycentrics b1,b2 of
 * triangle.rs:149-151 (0 for spheres). Miss: prim = PTB_MISS, t = 0 (sky.rs:79-91). 16 B. */
#define PTB_MISS 0xFFFFFFFFu
typedef struct ptb_hit {
  float    t;
  uint32_t prim;
  float    u, v;
} ptb_hit;

/* implementations/samplers/mod.rs:43-47 */
enum { PTB_METHOD_NAIVE = 0, PTB_METHOD_MIS = 1 };

/* implementations/samplers/mod.rs:22-29 + the constants of integrators/mod.rs:7-8 made explicit,
 * plus what a deterministic, shardable device renderer needs (seed, sample_offset). */
typedef struct ptb_render_opts {
  uint32_t width;
  uint32_t height;
  uint32_t samples_per_pixel; /* samples rendered by THIS call                                  */
  uint32_t sample_offset;     /* absolute index of the first sample (rank r renders [off, off+spp)) */
  uint32_t method;            /* PTB_METHOD_*; reference default is MIS (parameters.rs:34-35)   */
  uint32_t max_depth;         /* 0 -> 50  (integrators/mod.rs:7)                                */
  uint32_t rr_threshold;      /* 0xFFFFFFFF -> 3 (integrators/mod.rs:8); RR applies when depth > threshold */
  uint32_t flags;             /* reserved, 0                                                    */
  uint64_t seed;              /* counter-based RNG key; same seed + same sample range == same image */
  uint32_t row_begin;         /* image tile: only the pixel rows [row_begin, row_begin + row_count) are rendered;   */
  uint32_t row_count;         /* 0 -> every row from row_begin down. Pixels keep their full-image coordinates (camera,
                                 RNG key, accumulator position), so the tiles of one image add up to the whole image. */
} ptb_render_opts;
#define PTB_RR_DEFAULT 0xFFFFFFFFu

/* BVH2 node as exported for bit-exact tests: both children's boxes live in the parent (64 B).
 * child ref: bit31 set -> leaf, low bits = position in the Morton-sorted primitive order. */
#define PTB_LEAF_BIT 0x80000000u
typedef struct ptb_bvh_node {
  float    lmin[3], lmax[3];
  float    rmin[3], rmax[3];
  uint32_t left, right;
  uint32_t parent;  /* 0xFFFFFFFF for the root */
  uint32_t _pad;
} ptb_bvh_node;

typedef struct ptb_stats {
  uint64_t rays_camera;        /* traversals launched for camera rays                         */
  uint64_t rays_bounce;        /* ... for BSDF-sampled bounce rays                            */
  uint64_t rays_shadow_light;  /* ... for NEE rays towards a light primitive (check_hit_index) */
  uint64_t rays_shadow_sky;    /* ... for NEE rays towards the sky                            */
  uint64_t rays_reference;     /* the reference's own counter: naive = 1 per check_hit (integrators/mod.rs:34),
                                  MIS = 1 per bounce iteration (mis.rs:38)                     */
  uint64_t paths;              /* camera paths finished                                       */
  uint64_t wavefront_iterations;
  uint64_t kernel_launches;    /* kernels of this library launched since the last ptb_stats_reset */
  uint64_t nodes_fetched;      /* PTB_OPT_COUNT_TRAVERSAL: 64-byte BVH nodes fetched by closest-hit traversals */
  uint64_t prims_tested;       /* PTB_OPT_COUNT_TRAVERSAL: primitives tested by closest-hit traversals          */
  uint64_t rays_counted;       /* PTB_OPT_COUNT_TRAVERSAL: closest-hit traversals the two counters cover        */
  uint64_t trace_launches;     /* closest-hit kernel launches (k_trace / k_closest_hit_api)                     */
  double   build_ms;           /* last ptb_scene_commit: device LBVH build time               */
  double   render_ms;          /* last ptb_render: device time (CUDA events)                  */
  double   ms_generate;        /* PTB_OPT_TIME_KERNELS: summed CUDA-event time per kernel class */
  double   ms_trace;
  double   ms_shade;
  double   ms_shadow;
  double   ms_tail;            /* fused tail launches (k_tail: closest hit + shade + NEE of a chunk's last paths); they
                                  run on a side stream under the next chunk, so the classes may add up to > render_ms */
} ptb_stats;

typedef struct ptb_ctx ptb_ctx;

/* called between wavefront iterations; return non-zero to abort (random_sampler.rs:82-88 contract,
 * carrying counters only — the per-pass image is replaced by the device accumulator). */
typedef int32_t (*ptb_progress_fn)(void* user, uint64_t samples_completed, uint64_t rays_shot);

/* ------------------------------------------------------------ lifecycle -- */
uint32_t    ptb_abi_version(void);
int32_t     ptb_device_count(int32_t* count);
int32_t     ptb_create(int32_t device, ptb_ctx** out);
int32_t     ptb_destroy(ptb_ctx* ctx);
const char* ptb_last_error(const ptb_ctx* ctx);           /* ctx may be NULL: last error of ptb_create */
int32_t     ptb_set_stream(ptb_ctx* ctx, void* cuda_stream); /* optional: run on the caller's cudaStream_t */
int32_t     ptb_synchronize(ptb_ctx* ctx);
/* Measurement switches (off by default; they add event records / atomic counters to the hot path). */
enum { PTB_OPT_TIME_KERNELS = 1, PTB_OPT_COUNT_TRAVERSAL = 2 };
int32_t     ptb_set_option(ptb_ctx* ctx, uint32_t option, uint32_t value);

/* ---------------------------------------------------------------- scene -- */
/* Replaces what loader::load_file_full returns (loader/lib.rs:196-243) + Bvh::new's input
 * (implementations/acceleration/mod.rs:58-93). Host pointers; copied. n may be 0. */
int32_t ptb_scene_set_spheres(ptb_ctx* ctx, const ptb_sphere* spheres, size_t n);
int32_t ptb_scene_set_triangles(ptb_ctx* ctx, const ptb_triangle* triangles, size_t n);
int32_t ptb_scene_set_materials(ptb_ctx* ctx, const ptb_material* materials, size_t n);
int32_t ptb_scene_set_textures(ptb_ctx* ctx, const ptb_texture* textures, size_t n);
/* Bulk data of one texture (copied). PTB_TEX_IMAGE: ImageTexture.data (textures/mod.rs:202-245) = width*height RGB f32,
 * row-major, n_floats = 3*width*height. PTB_TEX_PERLIN: Perlin's tables (textures/mod.rs:75-112), width = height = 0,
 * n_floats = PTB_PERLIN_TABLE_WORDS. Call after ptb_scene_set_textures; commit fails with PTB_ERR_MISSING without it. */
int32_t ptb_scene_set_texture_data(ptb_ctx* ctx, uint32_t texture, uint32_t width, uint32_t height, const float* data,
                                   size_t n_floats);
int32_t ptb_scene_set_camera(ptb_ctx* ctx, const ptb_camera* camera);
int32_t ptb_scene_set_sky(ptb_ctx* ctx, const ptb_sky* sky);

/* Replaces Bvh::new (acceleration/mod.rs:58-93): builds the acceleration structure on the device — always the LBVH
 * (Morton sort, Karras hierarchy, refit), and on top of it, when the wide tree is selected, its SAH-driven collapse into
 * a compressed 8-wide BVH (quantised 96-byte nodes, leaf groups of up to 3 primitives) which the traversal kernels then
 * walk instead. PTB_BUILD_SAH makes the binary tree with the device's top-down SAH builder instead of the Karras
 * hierarchy (8 bins on each axis of a node's box, exact sweep for subtrees of <= 32 primitives, one primitive per leaf;
 * the replacement for the QUALITY of build_bvh + Split::Sah, acceleration/mod.rs:97-160, split.rs:78-187; CPU definition
 * oracle/sah_ref.hpp; 2..2^24 primitives, otherwise the LBVH is built). Same hits, fewer nodes per ray, a slower commit.
 * PTB_BUILD_DEFAULT follows the environment (PTB_BVH=binary|lbvh|sah|wide) and otherwise the library's default. */
enum { PTB_BUILD_DEFAULT = 0, PTB_BUILD_BINARY = 1, PTB_BUILD_WIDE = 2, PTB_BUILD_SAH = 4 };
int32_t ptb_scene_commit(ptb_ctx* ctx, uint32_t build_flags);

/* Bit-exact test hooks: n_prims Morton codes and primitive ids in sorted order, n_prims-1 nodes
 * (1 node when n_prims == 1). Any pointer may be NULL. After an SAH build prim_sorted is the SAH tree's own primitive
 * order (what its leaf references index) while morton_sorted still holds the sorted codes of the order it started from. */
int32_t ptb_bvh_info(ptb_ctx* ctx, uint64_t* n_prims, uint64_t* n_nodes);
int32_t ptb_bvh_export(ptb_ctx* ctx, uint32_t* morton_sorted, uint32_t* prim_sorted, ptb_bvh_node* nodes);
/* The same binary tree as the traversal kernels read it (bit-exact test hook): 32 bytes per node, both children's boxes
 * snapped outwards onto a 65536^3 grid over the scene box — eight 32-bit words {left box x, y, z, right box x, y, z (each:
 * low half = min, high half = max grid coordinate), left, right}; frame = grid origin xyz, grid step xyz
 * (plane = origin + q * step). nodes32 receives nothing when the committed scene uses the wide tree. CPU definition:
 * oracle/lbvh_ref.hpp. Any pointer may be NULL. The 32-byte nodes are a compile-time option of the library
 * (-DPTB_QNODES=1): the default build answers PTB_ERR_UNSUPPORTED. */
int32_t ptb_bvh_export_quantised(ptb_ctx* ctx, float frame[6], void* nodes32);

/* The compressed 8-wide tree, for bit-exact tests: *n_nodes = 0 when the committed scene uses the binary tree. Nodes are
 * 96 bytes each (layout: raytracing-rust_b200/csrc/ptb_common.cuh CwNode == oracle/cwbvh_ref.hpp CwNode), slot_prim maps
 * the tree's primitive order to original primitive ids (n_prims entries). Any pointer may be NULL. */
int32_t ptb_bvh_wide_info(ptb_ctx* ctx, uint64_t* n_nodes, uint32_t* max_leaf);
/* Which builder made the committed tree: *builder = PTB_BUILD_BINARY (Karras LBVH), PTB_BUILD_SAH or PTB_BUILD_WIDE;
 * *sah_levels = levels of large tasks the SAH build took (0 otherwise). Either pointer may be NULL. */
int32_t ptb_bvh_builder(ptb_ctx* ctx, uint32_t* builder, uint32_t* sah_levels);
int32_t ptb_bvh_wide_export(ptb_ctx* ctx, void* nodes96, uint32_t* slot_prim);

/* ---------------------------------------------------------- closest hit -- */
/* Replaces AccelerationStructure::check_hit (acceleration/mod.rs:265-298) for a batch of rays.
 * Host buffers; H2D + kernel + D2H. hits[i] answers rays[i]; internally batches of >= 65 536 rays are traced in a
 * spatially sorted order (PTB_HIT_SORT=0 disables), which needs 16 B of device scratch per ray. */
int32_t ptb_closest_hit(ptb_ctx* ctx, const ptb_ray* rays, size_t n, ptb_hit* hits);
/* Same, buffers already resident in device memory (kernel only; used for roofline timing). */
int32_t ptb_closest_hit_device(ptb_ctx* ctx, const void* d_rays, size_t n, void* d_hits);

/* --------------------------------------------------------------- render -- */
/* Replaces Sampler::sample_image (samplers/mod.rs:7-20, random_sampler.rs:10-99) + the running mean of
 * src/main.rs:175-191: adds opts->samples_per_pixel samples per pixel into the device accumulator (sums).
 * Device memory: the path state of a call takes at most 8 GiB (PTB_POOL_BYTES=<bytes> changes the budget; 69 B per path,
 * 133 B with MIS), never more than half of the memory that was free at the context's first large render. A call that fits
 * (124 Mi paths, 64 Mi with MIS) runs as one chunk; a larger one runs chunk by chunk in two slots of half the budget each,
 * the wide iterations of chunk k+1 overlapping the latency-bound tail of chunk k. PTB_POOL_PATHS=<paths> sets the slot size directly,
 * PTB_TAIL_PATHS=<paths> the live-path count (default 65 536) below which a chunk is handed to the fused tail kernel (0: never),
 * PTB_WAVEFRONT=queue selects the small-pool regenerating mode. The
 * image is a pure function of (scene, opts): pool size, chunking and GPU count only change the f32 summation order
 * (finished paths are added with float atomics, so two runs of the same call agree to ~1e-6 relative, not bit for bit).
 * A non-zero return of `progress` stops the call with PTB_ERR_ABORTED and CLEARS the accumulator (samples of earlier
 * calls included): a partially rendered call has no sample count to normalise by. */
int32_t ptb_render(ptb_ctx* ctx, const ptb_render_opts* opts, ptb_progress_fn progress, void* user);
/* The reference's per-pass presentation contract (random_sampler.rs:31-98, SamplerProgress samplers/mod.rs:49-63): one
 * pass = one sample of every pixel. `update` receives the SINGLE-SAMPLE image of a finished pass (row-major, top row
 * first, RGB f32, n_floats = width*height*3; valid only during the call), the 1-based number of that pass and the
 * pass's rays_shot (the reference's own ray counter). As in the reference the image of pass k is handed over after pass
 * k+1 has been rendered, and the last one after the loop; a non-zero return stops the render (PTB_ERR_ABORTED; the
 * reference returns without the final call). Every pass is also added to the device accumulator, so ptb_accum_read
 * afterwards yields the running mean the TUI closure of src/main.rs:175-191 computes. Pass k uses absolute sample index
 * opts->sample_offset + k: the sum of the passes equals one ptb_render call of the same options. */
typedef int32_t (*ptb_pass_fn)(void* user, const float* pass_image, size_t n_floats, uint64_t pass_number, uint64_t rays_shot);
int32_t ptb_render_passes(ptb_ctx* ctx, const ptb_render_opts* opts, ptb_pass_fn update, void* user);
int32_t ptb_accum_clear(ptb_ctx* ctx);
/* Reads back width*height*3 floats, row-major, top row first, RGB (the layout of
 * SamplerProgress.current_image, samplers/mod.rs:49-63). normalise != 0 divides by the samples accumulated. */
int32_t ptb_accum_read(ptb_ctx* ctx, float* rgb, size_t n_floats, int32_t normalise);
/* Device pointer of the width*height*3 float sums (for the NCCL reduce) and its element count. */
int32_t ptb_accum_device_ptr(ptb_ctx* ctx, void** d_ptr, size_t* n_floats);
/* After an external reduce wrote sums of `total_samples` samples into the accumulator. */
int32_t ptb_accum_set_samples(ptb_ctx* ctx, uint64_t total_samples);

/* ------------------------------------------------------------ multi-GPU -- */
/* The path shards by samples (independent), scene and BVH replicated per GPU: rank r of `world` renders the absolute
 * samples [first, first + count) of every pixel. */
void    ptb_shard_samples(uint32_t samples_per_pixel, uint32_t sample_offset, int32_t rank, int32_t world,
                          uint32_t* first, uint32_t* count);
/* Second axis, for requests with fewer samples than GPUs: rank r renders every sample of the pixel rows
 * [first, first + count) of the `rows` rows starting at row_begin (bands of whole rows; sizes differ by at most 1). */
void    ptb_shard_rows(uint32_t rows, uint32_t row_begin, int32_t rank, int32_t world, uint32_t* first, uint32_t* count);
/* Scene::render across n GPUs of one box: ctxs[r] (one per GPU, scene already committed on each) renders its share on its
 * own host thread — a sample range of every pixel when samples_per_pixel >= n, else a band of pixel rows at every sample —
 * then ONE ncclReduce(sum) of the accumulators (width*height*3 f32) to ctxs[0] — the path's only collective; the
 * communicators are checked with ncclCommGetAsyncError after the reduce has completed. Afterwards ptb_accum_read(ctxs[0], ..) returns the image of all opts->samples_per_pixel samples. NCCL is
 * bound at run time (libnccl.so.2, or $PTB_NCCL_LIB); n == 1 never touches it. */
int32_t ptb_render_multi(ptb_ctx* const* ctxs, int32_t n, const ptb_render_opts* opts);

/* ----------------------------------------------------- sampler test hook -- */
/* The reference's best-pinned tests are chi-squared tests of its direction samplers against their pdfs
 * (implementations/statistics/spherical_sampling.rs:64-226, bxdfs/lambertian.rs:30-48, bxdfs/trowbridge_reitz_vndf.rs:156-218,
 * sky.rs:43-78). These two entry points run the DEVICE samplers — the same functions the shade kernel calls — outside a
 * render so the same harness can be pointed at them. Host buffers. */
enum {
  PTB_SAMPLER_LAMBERTIAN = 0,     /* lambertian::sample about `normal`                         (bxdfs/lambertian.rs:5-22)   */
  PTB_SAMPLER_TR_VNDF = 1,        /* trowbridge_reitz_vndf::sample(alpha, incoming = aux, normal) (..._vndf.rs:35-52, 84-113) */
  PTB_SAMPLER_SKY = 2,            /* Sky::sample / Sky::pdf of the committed scene             (sky.rs:43-78)               */
  PTB_SAMPLER_LIGHT = 3,          /* sample_visible_from_point / scattering_pdf of light `light_index` seen from the shading
                                     point aux with surface normal `normal`                   (sphere.rs:112-170, triangle.rs:249-280) */
  PTB_SAMPLER_UNIFORM_SPHERE = 4  /* the fixed-budget stand-in for random_unit_vector          (utility/mod.rs:15-25)       */
};
typedef struct ptb_sampler_query {
  uint32_t kind;         /* PTB_SAMPLER_*                                                                           */
  float    alpha;        /* TR_VNDF                                                                                  */
  ptb_vec3 normal;       /* LAMBERTIAN, TR_VNDF: surface normal; LIGHT: normal at the shading point                  */
  ptb_vec3 aux;          /* TR_VNDF: unit direction from the surface towards the viewer; LIGHT: the shading point     */
  uint32_t light_index;  /* LIGHT: index into the scene's light list (original primitive order)                       */
  uint64_t seed;         /* sample k draws from Philox counter (k, 0, test stream), key = seed                        */
} ptb_sampler_query;
/* n sampled directions (3 floats each) and, when pdf != NULL, the sampler's own pdf of each. */
int32_t ptb_sample_only(ptb_ctx* ctx, const ptb_sampler_query* query, size_t n, float* dirs, float* pdf);
/* pdf of n given unit directions under the same sampler. */
int32_t ptb_sampler_pdf(ptb_ctx* ctx, const ptb_sampler_query* query, const float* dirs, size_t n, float* pdf);

int32_t ptb_stats_get(ptb_ctx* ctx, ptb_stats* out);
int32_t ptb_stats_reset(ptb_ctx* ctx);

/* ------------------------------------------------------------- host side -- */
/* C++ restatement of the `loader` crate (loader/lib.rs:196-243, parser.rs:112-197, misc.rs, materials.rs,
 * primitives.rs, meshes.rs, obj.rs, textures.rs): parses a .ssml file (and the OBJ files it names) into the
 * POD arrays above. No GPU needed. */
typedef struct ptb_host_scene ptb_host_scene;
int32_t ptb_ssml_load_file(const char* path, ptb_host_scene** out);
int32_t ptb_ssml_load_str(const char* text, const char* base_dir, ptb_host_scene** out);
void    ptb_host_scene_free(ptb_host_scene* s);
const char* ptb_host_last_error(void);
size_t  ptb_host_scene_spheres(const ptb_host_scene* s, const ptb_sphere** out);
size_t  ptb_host_scene_triangles(const ptb_host_scene* s, const ptb_triangle** out);
size_t  ptb_host_scene_materials(const ptb_host_scene* s, const ptb_material** out);
size_t  ptb_host_scene_textures(const ptb_host_scene* s, const ptb_texture** out);
/* bulk data of texture `texture` (see ptb_scene_set_texture_data); returns n_floats, 0 if the texture has none */
size_t  ptb_host_scene_texture_data(const ptb_host_scene* s, uint32_t texture, uint32_t* width, uint32_t* height,
                                    const float** data);
int32_t ptb_host_scene_camera(const ptb_host_scene* s, ptb_camera* out);
int32_t ptb_host_scene_sky(const ptb_host_scene* s, ptb_sky* out);
/* SimpleCamera::new (camera.rs:20-53); aspect is 16/9 in the loader (loader/misc.rs:15). */
int32_t ptb_camera_make(ptb_vec3 origin, ptb_vec3 lookat, ptb_vec3 vup, float hfov_deg, float aspect,
                        float aperture, float focus_dist, ptb_camera* out);
/* Perlin::new (textures/mod.rs:90-112, 141-159): 256 gen_range(-1..1) scalars + three Fisher-Yates permutations
 * (`target = gen_range(0..i)` for i = 255..1). The reference seeds them from OS entropy (unpinned); here they are a pure
 * function of `seed` (Philox4x32-10). out = PTB_PERLIN_TABLE_WORDS words. */
int32_t ptb_perlin_tables(uint64_t seed, float* out);
/* ImageTexture::new's decode (textures/mod.rs:208-245, `image` crate to_rgb32f): ppm/pgm (P2,P3,P5,P6), pfm, bmp
 * (24/32 bit) and png (8/16 bit, non-interlaced). *rgb = width*height*3 f32, released with ptb_image_free. */
int32_t ptb_image_load(const char* filename, uint32_t* width, uint32_t* height, float** rgb);
void    ptb_image_free(float* rgb);
const char* ptb_image_last_error(void);
/* ptb_scene_set_* for every array of a loaded scene. */
int32_t ptb_scene_upload(ptb_ctx* ctx, const ptb_host_scene* s);

/* output::save_data_to_image (output/lib.rs:74-113): gamma + 8-bit for ppm/bmp/png, linear f32 for pfm. */
int32_t ptb_image_save(const char* filename, uint32_t width, uint32_t height, const float* rgb, float gamma);

#ifdef __cplusplus
}
#endif
#endif /* PTB200_H */
