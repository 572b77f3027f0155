/*
 * b2rt.h -- C ABI of the B200-native wide-BVH path-tracing hot path.
 *
 * This is the drop-in boundary for ONE path of saipraveenb25/cuda-raytracer: breadth-first
 * wide-BVH traversal with dynamic ray scheduling, plus the per-bounce shading that feeds it.
 * The reference has no FFI; the path sits behind three C++ classes.  Every entry point below
 * names the reference interface it replaces (paths relative to the reference checkout).
 *
 *   PathTracer  (CPU renderer shell)        src/pathtracer.h:51-257, src/pathtracer.cpp
 *   BVHAccel    (acceleration structure)    src/bvh.h:99-191,        src/bvh.cpp
 *   CudaRenderer(GPU renderer, the caller   src/cudaRenderer.h:173-272, src/cudaRenderer.cu
 *                of the hot path today)
 *
 * Conventions: plain pointers and sizes only; all inputs are copied (caller keeps ownership);
 * all outputs are copied into caller buffers; every call returns 0 on success or a negative
 * b2rt_status (the reference calls exit()/printf instead, e.g. src/cudaRenderer.cu:1683-1687);
 * b2rt_last_error() returns a thread-local message.  Handles are single-owner, not thread-safe.
 * There is no CPU fallback: without a CUDA device every device entry point fails with
 * B2RT_ERR_NO_DEVICE.
 */
#ifndef B2RT_H
#define B2RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2RT_ABI_VERSION 2   /* 2: b2rt_stats grew the level-0 fields, the overflow counters and graph_replays; b2rt_comm_*, b2rt_write_*, b2rt_bench_fp32 */

typedef enum b2rt_status {
  B2RT_OK = 0,
  B2RT_ERR_INVALID = -1,   /* bad argument / wrong state (reference: silent no-op, pathtracer.cpp:72-74,184) */
  B2RT_ERR_NO_DEVICE = -2, /* no CUDA device; there is no CPU path */
  B2RT_ERR_CUDA = -3,      /* a CUDA runtime call failed; see b2rt_last_error() */
  B2RT_ERR_OOM = -4,
  B2RT_ERR_IO = -5,
  B2RT_ERR_OVERFLOW = -6   /* a ray queue overflowed even after splitting the batch */
} b2rt_status;

/* ---- scene description (flat arrays) ------------------------------------------------------
 * Replaces StaticScene::Scene / Mesh / Triangle / Sphere / AreaLight object graphs
 * (src/static_scene/{scene,object,triangle,sphere,light}.h) and the CuTriangle/CuBSDF/CuEmitter
 * upload of src/cudaRenderer.cu:1694-1792.  Primitive ids reported by the intersect calls are
 * indices into THIS order: triangles 0..n_tris-1, then spheres n_tris..n_tris+n_spheres-1.
 */
typedef enum b2rt_material_kind {
  B2RT_MAT_DIFFUSE = 0,    /* DiffuseBSDF   src/bsdf.h, f = albedo/pi (src/bsdf.cpp:37-39) */
  B2RT_MAT_MIRROR = 1,     /* MirrorBSDF    (delta) */
  B2RT_MAT_GLASS = 2,      /* GlassBSDF     (delta, Fresnel reflect/refract) */
  B2RT_MAT_EMISSION = 3,   /* EmissionBSDF  (f = 0, get_emission() = radiance) */
  B2RT_MAT_REFRACTION = 4, /* RefractionBSDF(delta, always refract) */
  B2RT_MAT_GLOSSY = 5      /* GlossyBSDF(reflectance, roughness) (src/bsdf.h:143-162, commented out in the checkout):
                              normalised Phong lobe about the mirror direction, f = albedo (n+2)/(2 pi) cos^n, with the
                              INTEGER exponent n = clamp(int(2 / roughness^2) - 2, 1, 4096); not delta */
} b2rt_material_kind;

typedef struct b2rt_material {
  int32_t kind;            /* b2rt_material_kind */
  float albedo[3];         /* diffuse albedo, or mirror/glass reflectance */
  float transmittance[3];  /* glass / refraction */
  float emission[3];       /* EmissionBSDF radiance */
  float ior;               /* glass / refraction */
  float roughness;         /* glossy: lobe width (see B2RT_MAT_GLOSSY); parsed but unused for glass / refraction */
} b2rt_material;

typedef enum b2rt_light_kind {
  B2RT_LIGHT_AREA = 0,        /* StaticScene::AreaLight  src/static_scene/light.cpp:71-92 */
  B2RT_LIGHT_POINT = 1,       /* PointLight   (delta) */
  B2RT_LIGHT_DIRECTIONAL = 2  /* DirectionalLight (delta) */
} b2rt_light_kind;

typedef struct b2rt_light {
  int32_t kind;
  float radiance[3];
  float position[3];
  float direction[3];      /* unit; area light emits where dot(d, direction) < 0 */
  float dim_x[3];          /* area light edge vectors (full extent), light.cpp:74-79 */
  float dim_y[3];
} b2rt_light;

typedef struct b2rt_scene_desc {
  uint32_t n_tris;
  const float* tri_verts;        /* n_tris * 9 : p1 p2 p3 */
  const float* tri_normals;      /* n_tris * 9 : per-vertex shading normals, or NULL (geometric) */
  const uint32_t* tri_material;  /* n_tris, or NULL (all material 0) */
  uint32_t n_spheres;
  const float* spheres;          /* n_spheres * 4 : cx cy cz r */
  const uint32_t* sphere_material;
  uint32_t n_materials;
  const b2rt_material* materials;
  uint32_t n_lights;
  const b2rt_light* lights;
} b2rt_scene_desc;

/* Camera: replaces CMU462::Camera (src/camera.h) as consumed by PathTracer::set_camera
 * (src/pathtracer.h:83).  generate_ray contract: src/camera.h:71-81.  c2w columns are the
 * camera x, y and z (= direction TO the camera) axes, src/camera.cpp:87-109. */
typedef struct b2rt_camera {
  float pos[3];
  float c2w[9];   /* column-major: c2w[0..2] = x axis, [3..5] = y axis, [6..8] = z axis */
  float hfov_deg;
  float vfov_deg;
} b2rt_camera;

/* Knobs: the PathTracer constructor arguments (src/pathtracer.h:57-60) and the CLI flags
 * -s -m -l (src/main.cpp:78-105).  The CUDA reference hard-codes these
 * (src/cudaRenderer.h:58-83). */
typedef struct b2rt_config {
  uint32_t ns_aa;            /* samples per pixel */
  uint32_t max_ray_depth;    /* surface interactions per path (1 = direct lighting only) */
  uint32_t ns_area_light;    /* shadow rays per area light per interaction */
  uint64_t seed;             /* counter-based RNG key; stream = (pixel, sample, bounce) */
  float ray_eps;             /* t_min offset of secondary rays; 0 -> 1e-4 */
  uint32_t bvh_width;        /* 2, 4, 8 or 16 (2 and 16: host builder only; they exist for the width sweep); 0 -> default (4) */
  uint32_t max_leaf_size;    /* BVHAccel max_leaf_size (src/bvh.h:111); 0 -> default (host builder: 4, ended earlier by SAH cost; device builder: 3) */
  uint32_t treelet_bytes;    /* shared-memory subtree budget; 0 -> default */
  uint32_t max_wave_paths;   /* paths in flight per wave; 0 -> default */
  uint32_t median_threshold; /* 3x3 median when total spp < this (POST_PROCESS_THRESHOLD,
                                src/cudaRenderer.h:70); 0 = never */
  int32_t device;            /* CUDA device ordinal; -1 -> current */
  /* sample sharding for multi-GPU: this handle renders samples
     s = sample_first + k*sample_stride, k = 0..ns_aa_local-1 (ns_aa is the LOCAL count). */
  uint32_t sample_first;
  uint32_t sample_stride;    /* 0 -> 1 */
  uint32_t bvh_builder;      /* 0 -> automatic: host binned-SAH build below 2^14 primitives, device build
                                (b2rt_bvh_build_device: Morton order + PLOC clustering, trees within a few percent of
                                the SAH builder's) from there on; 1 -> host; 2 -> device */
  /* Reconstruction filter applied while total spp < median_threshold.  0 = the reference's 3x3
     per-channel median (kernelMedianFilter, src/cudaRenderer.cu:773-842); 1 = 3x3 binomial Gaussian (the reference
     has one commented out, :755-771; weights (1,2,1)x(1,2,1), taps outside the image dropped and the rest
     renormalised); 2 = 5x5 joint bilateral: binomial (1,4,6,4,1)^2 spatial weights times the range weight
     1 / (1 + |c_q - c_p|^2 / sigma_r^2) on the RGB difference to the centre pixel. */
  uint32_t filter_kind;
  float filter_sigma_r;      /* bilateral range scale; 0 -> 0.25 */
} b2rt_config;

typedef struct b2rt_stats {
  uint64_t rays_camera;
  uint64_t rays_bounce;
  uint64_t rays_shadow;
  uint64_t node_visits;      /* (ray, wide node) pairs tested */
  uint64_t leaf_prim_tests;  /* ray-primitive tests */
  uint64_t subtree_visits;   /* (ray, subtree) queue entries processed */
  uint64_t queue_pushes;     /* ray ids pushed to child-subtree queues */
  uint64_t staged_bytes;     /* subtree bytes staged global->shared by TMA bulk copies */
  uint64_t hit_updates;      /* packed (t, prim) 64-bit atomicMin operations */
  uint64_t kernel_launches;  /* kernels launched by the last render/intersect call */
  uint64_t traverse_launches;/* of which launches of the traversal kernel */
  double ms_total;           /* device time of the last render/intersect call (CUDA events) */
  double ms_traverse;        /* device time inside traversal + scheduling kernels */
  double ms_build;           /* host BVH build + upload of the current scene */
  uint32_t bvh_nodes;
  uint32_t bvh_subtrees;
  uint32_t bvh_levels;       /* subtree levels = traversal passes per ray batch */
  uint32_t bvh_width;
  uint64_t bvh_bytes;
  /* the same traversal counters for LEVEL 0 alone (the root subtree: every ray, streamed from the dense ray list);
     the deeper levels (rays gathered by id, 64-bit atomicMin merges) are the difference to the totals above */
  uint64_t node_visits_l0;
  uint64_t leaf_prim_tests_l0;
  uint64_t queue_pushes_l0;
  uint64_t staged_bytes_l0;
  uint64_t hit_updates_l0;
  uint64_t traverse_launches_l0;
  double ms_traverse_l0;
  /* ray-queue overflows of the last render call (b2rt_wait recovers from them, the frame is complete either way):
     waves rendered again, and how often the queues were enlarged (x2 per step, kept for later frames) */
  uint64_t waves_retried;
  uint64_t queues_grown;
  /* frames of this renderer that were replayed from a captured CUDA graph so far (a frame whose launch sequence equals
     the previous one's is captured on its second occurrence; B2RT_GRAPH=0 disables it) */
  uint64_t graph_replays;
} b2rt_stats;

const char* b2rt_last_error(void);
int b2rt_abi_version(void);
int b2rt_device_count(void);

/* ---- BVHAccel -------------------------------------------------------------------------------
 * b2rt_bvh_build        replaces BVHAccel::BVHAccel(prims, max_leaf_size) + compactedTree() +
 *                       BVHSubTree::compress()        src/bvh.cpp:339-365, 275-337, 234-273
 * b2rt_bvh_intersect    replaces BVHAccel::intersect(const Ray&, Intersection*) (closest hit)
 *                       src/bvh.h:150-163; batch form of kernelRayIntersectSingle/Level +
 *                       kernelMergeIntersections   src/cudaRenderer.cu:1304-1310,1435-1489,515-540
 * b2rt_bvh_occluded     replaces BVHAccel::intersect(const Ray&) (any hit)  src/bvh.h:139-148
 * Rays are SoA host arrays: org[3n], dir[3n], tmin[n], tmax[n].  Results: hit_t[n] (= tmax
 * sentinel INFINITY when missed) and hit_prim[n] (0xFFFFFFFF when missed).  Ties in t resolve
 * to the lowest primitive id.
 */
typedef struct b2rt_bvh b2rt_bvh;

int b2rt_bvh_build(const b2rt_scene_desc* scene, uint32_t max_leaf_size, uint32_t width,
                   uint32_t treelet_bytes, int32_t device, b2rt_bvh** out);
/* Same result type, built on the device (Morton codes, radix sort, binary radix
 * tree, bottom-up boxes, wide collapse, subtree packing and serialisation all as CUDA kernels).
 * Replaces the same reference functions as b2rt_bvh_build; meant for scenes whose host build
 * dominates set-up.  The binary tree is built by parallel locally-ordered clustering (PLOC) over the Morton
 * order (search radius 8, leaves of at most 3 primitives by default); rays trace as fast as on the host builder's
 * SAH tree (cfg2 equal, cfg3 stand-in 2-3 % faster, 10 M soup equal / 10 % faster). */
int b2rt_bvh_build_device(const b2rt_scene_desc* scene, uint32_t max_leaf_size, uint32_t width,
                          uint32_t treelet_bytes, int32_t device, b2rt_bvh** out);
/* Structural check of the BVH a handle holds (either builder): the subtree blobs are read back
 * and walked the way the traversal kernel decodes them (every primitive stored once, child boxes
 * contain their contents, exits point to the next level, byte / stack budgets).  out8 as in
 * b2rt_bvh_validate_host.  Mirrors the invariants of SURVEY 8c(iii). */
int b2rt_bvh_validate(b2rt_bvh* bvh, const b2rt_scene_desc* scene, uint64_t out8[8]);
int b2rt_bvh_intersect(b2rt_bvh* bvh, const float* org, const float* dir, const float* tmin,
                       const float* tmax, uint64_t n, float* hit_t, uint32_t* hit_prim);
int b2rt_bvh_occluded(b2rt_bvh* bvh, const float* org, const float* dir, const float* tmin,
                      const float* tmax, uint64_t n, uint8_t* occluded);
/* Device-resident variant used for kernel-only timing: rays are generated on the device
 * (mode 0: camera-like coherent rays toward the scene box, mode 1: incoherent), traversed
 * `repeats` times; reports device ms per repeat.  No reference equivalent (the reference
 * times whole frames, src/cudaRenderer.cu:2558). */
int b2rt_bvh_bench_rays(b2rt_bvh* bvh, uint64_t n, int mode, uint64_t seed, int repeats,
                        int any_hit, double* ms_per_repeat, uint64_t* hits);
/* Distance slicing of the batch traversal (no reference equivalent; the reference's level-synchronous
 * traversal, src/cudaRenderer.cu:2304-2331, has the same lack of front-to-back order across
 * nodes).  Pass p traces only [lo, lo + first_slice * growth^p] of every ray still without a
 * hit, the last of `passes` takes the remainder; results are identical to the unsliced trace.
 * first_slice > 0: explicit length; 0: off; < 0: automatic (2 mean free paths of the scene, and
 * only when the subtree graph has >= 3 levels and that length is below 1/8 of the scene
 * diagonal -- the default after b2rt_bvh_build).  growth <= 1 or passes < 2 select 4 / 4. */
int b2rt_bvh_set_slicing(b2rt_bvh* bvh, float first_slice, float growth, int32_t passes);
int b2rt_bvh_get_stats(b2rt_bvh* bvh, b2rt_stats* out);
/* Measured FP32 peak of the device (an FFMA-saturating kernel: 16 independent chains per thread, every SM full),
 * the denominator of the roofline's arithmetic term.  No reference equivalent. */
int b2rt_bench_fp32(int32_t device, double* tflops);
/* get_bbox(): src/bvh.h:120-126.  out[6] = min xyz, max xyz */
int b2rt_bvh_get_bbox(b2rt_bvh* bvh, float* out6);
void b2rt_bvh_destroy(b2rt_bvh* bvh);

/* ---- PathTracer / CudaRenderer ---------------------------------------------------------------
 * b2rt_create            PathTracer::PathTracer(...)        src/pathtracer.cpp:23-63
 *                        CudaRenderer::CudaRenderer/setup   src/cudaRenderer.cu:1496,1872-2113
 * b2rt_set_scene         PathTracer::set_scene+build_accel  src/pathtracer.cpp:71-92,215-239
 *                        CudaRenderer::loadScene            src/cudaRenderer.cu:1679-1842
 * b2rt_set_camera        PathTracer::set_camera             src/pathtracer.cpp:94-103
 *                        CudaRenderer::setViewpoint         src/cudaRenderer.cu:1845-1870
 * b2rt_set_frame_size    PathTracer::set_frame_size         src/pathtracer.cpp:105-114
 *                        CudaRenderer::allocOutputImage     src/cudaRenderer.cu:2119
 * b2rt_start             PathTracer::start_raytracing       src/pathtracer.cpp:183-213 (async)
 * b2rt_is_done / b2rt_wait  PathTracer::is_done             src/pathtracer.cpp:572-575
 * b2rt_stop              PathTracer::stop                   src/pathtracer.cpp:116-139
 * b2rt_clear             PathTracer::clear / CudaRenderer::clearImage
 * b2rt_render            CudaRenderer::render (blocking: start + wait, accumulates ns_aa more
 *                        samples onto the running sum like renderAccumulate, :2419-2457)
 * b2rt_read_hdr          HDRImageBuffer data, Spectrum RGB32F, index x + y*w (src/image.h:114-118)
 * b2rt_read_ldr          ImageBuffer RGBA8 via toColor (src/image.h:49-58,173-188)
 * b2rt_read_rgba32f      CudaRenderer::getImage float4 RGBA (src/cudaRenderer.cu:1539-1570) but
 *                        row-major x + y*w (the reference's transposed x*H+y is NOT preserved)
 */
typedef struct b2rt_renderer b2rt_renderer;

int b2rt_create(const b2rt_config* cfg, b2rt_renderer** out);
int b2rt_set_config(b2rt_renderer* r, const b2rt_config* cfg); /* ns_aa/max_ray_depth/... knobs */
int b2rt_set_scene(b2rt_renderer* r, const b2rt_scene_desc* scene);
int b2rt_set_camera(b2rt_renderer* r, const b2rt_camera* cam);
/* Environment light: the `HDRImageBuffer* envmap` argument of PathTracer::PathTracer (src/pathtracer.h:57-60; CLI -e,
 * src/main.cpp:78-105) + EnvironmentLight (src/static_scene/environment_light.h; sample_L / sample_dir are stubs in the
 * checkout).  rgb = width*height RGB fp32 triples, index x + y*width, row 0 = the +y pole, x = azimuth
 * atan2(d.z, d.x) / 2 pi; copied.  NULL removes it.  Rays that leave the scene see the map (bilinear look-up; counted
 * like emitted radiance: camera rays and after delta bounces); every diffuse / glossy interaction samples it as one
 * more light, uniformly over the sphere.  Restarts accumulation. */
int b2rt_set_envmap(b2rt_renderer* r, const float* rgb, uint32_t width, uint32_t height);
int b2rt_set_frame_size(b2rt_renderer* r, uint32_t width, uint32_t height);
int b2rt_start(b2rt_renderer* r);
int b2rt_is_done(b2rt_renderer* r);     /* 1 done, 0 running, <0 error */
int b2rt_wait(b2rt_renderer* r);
int b2rt_stop(b2rt_renderer* r);
int b2rt_clear(b2rt_renderer* r);
int b2rt_render(b2rt_renderer* r);
int b2rt_read_hdr(b2rt_renderer* r, float* rgb, size_t n_floats);
int b2rt_read_ldr(b2rt_renderer* r, uint32_t* rgba8, size_t n_pixels);
int b2rt_read_rgba32f(b2rt_renderer* r, float* rgba, size_t n_floats);
/* CudaRenderer::getImage (src/cudaRenderer.cu:1539-1570, src/cudaRenderer.h:196): resolves the current mean
 * (+ median filter below the threshold), copies it into a renderer-owned page-locked host buffer and returns that
 * buffer: float4 RGBA, row-major x + y*w, valid until the next call on this handle or b2rt_destroy.  No copy into a
 * caller buffer, like the reference (`const Image* getImage()`). */
int b2rt_get_image(b2rt_renderer* r, const float** rgba, size_t* n_floats);
int b2rt_get_stats(b2rt_renderer* r, b2rt_stats* out);
/* The per-GPU accumulation buffer (float4 per pixel: rgb SUM + sample count), for callers that combine
 * the GPUs with their own collective (b2rt_reduce_accum below is the library's).  The pointer is device
 * memory owned by the handle, valid until set_frame_size/destroy. */
int b2rt_accum_device_ptr(b2rt_renderer* r, void** dev_ptr, size_t* n_floats);
/* ---- multi-GPU combine (no reference equivalent: src/cudaRenderer.cu is single-GPU, it never calls
 * cudaSetDevice; it replaces nothing and extends CudaRenderer::renderAccumulate, :2419-2457, across GPUs) ------
 * Samples are sharded by index (b2rt_config.sample_first / sample_stride), the scene is replicated, and ONE
 * ncclReduce (fp32 sum) of the float4 accumulation buffers combines the frame on `root`; root < 0 = all-reduce.
 * The reduce is enqueued on the renderer's stream behind the frame (it first completes b2rt_wait's bookkeeping,
 * incl. re-rendering overflowed waves); b2rt_read_* / b2rt_get_image on the root then resolve the combined frame.
 * One process per GPU: rank 0 calls b2rt_comm_unique_id, ships the 128 bytes to the other ranks by any means
 * (MPI, torch.distributed, a file) and every rank calls b2rt_comm_create.  One process, n GPUs:
 * b2rt_comm_create_all (ncclCommInitAll) + b2rt_reduce_accum_all (one NCCL group).  NCCL is bound at run time
 * (dlopen; the copy already loaded in the process wins); b2rt_comm_version() = 0 when none is available. */
#define B2RT_COMM_ID_BYTES 128
typedef struct b2rt_comm b2rt_comm;
int b2rt_comm_version(void);
int b2rt_comm_unique_id(uint8_t id[B2RT_COMM_ID_BYTES]);
int b2rt_comm_create(int32_t n_ranks, int32_t rank, const uint8_t id[B2RT_COMM_ID_BYTES], int32_t device, b2rt_comm** out);
int b2rt_comm_create_all(int32_t n, const int32_t* devices /* NULL = 0..n-1 */, b2rt_comm** out /* [n] */);
int b2rt_reduce_accum(b2rt_renderer* r, b2rt_comm* comm, int32_t root);
int b2rt_reduce_accum_all(b2rt_renderer** r, b2rt_comm** comm, int32_t n, int32_t root);
void b2rt_comm_destroy(b2rt_comm* comm);
int b2rt_stream_handle(b2rt_renderer* r, void** cuda_stream);
/* Run on a caller-owned CUDA stream (e.g. the framework stream that also carries the NCCL reduce).
 * cuda_stream = NULL restores the handle's own stream. */
int b2rt_set_stream(b2rt_renderer* r, void* cuda_stream);
/* collect_counters: traversal statistics (node visits, pushes, ...) cost a few percent; off by default.
 * time_kernels: CUDA-event pair around every traversal-kernel launch -> b2rt_stats.ms_traverse. */
int b2rt_set_profiling(b2rt_renderer* r, int collect_counters, int time_kernels);
void b2rt_destroy(b2rt_renderer* r);

/* ---- host helpers (no device needed) -----------------------------------------------------------
 * b2rt_scene_load / b2rt_scene_free: flat binary scene files (.b2s) written by tools/dae2scene.py
 * or b2rt_scene_save; b2rt_load_dae parses the COLLADA subset + CMU462 <extra> profile that
 * Collada::ColladaParser::load handles (src/collada/collada.cpp:117-951) and flattens it the way
 * Application::load / DynamicScene::Mesh / StaticScene::Mesh do (src/application.cpp:347-435,
 * src/dynamic_scene/mesh.cpp:21-46, src/static_scene/object.cpp:17-72).
 * b2rt_camera_place reproduces Application::load's camera placement (src/application.cpp:395-408,
 * src/camera.cpp:15-33,87-109) for a given frame size.
 */
typedef struct b2rt_scene_file {
  b2rt_scene_desc desc;      /* pointers into storage owned by this object */
  b2rt_camera camera;        /* placed for `aspect` = width/height given at load */
  float cam_dir[3];          /* COLLADA camera view direction (world) */
  float cam_hfov_deg, cam_vfov_deg;
  float bbox[6];
  void* storage;
} b2rt_scene_file;

int b2rt_scene_load(const char* path, b2rt_scene_file** out);
int b2rt_scene_save(const char* path, const b2rt_scene_file* scene);
int b2rt_load_dae(const char* path, b2rt_scene_file** out);
void b2rt_scene_free(b2rt_scene_file* s);
/* Builds the wide BVH on the host and checks the serialised subtree blobs structurally (every primitive
 * in exactly one leaf, child boxes contain their contents, exits point one level down, byte budget and
 * per-ray stack bound respected).  out[8] = subtrees, levels, wide nodes, leaves, blob bytes, largest
 * subtree bytes, stack bound, exits.  Test hook; no reference equivalent. */
int b2rt_bvh_validate_host(const b2rt_scene_desc* scene, uint32_t max_leaf_size, uint32_t width,
                           uint32_t treelet_bytes, uint64_t out[8]);
int b2rt_camera_place(const float bbox[6], const float view_dir[3], float hfov_deg,
                      float vfov_deg, uint32_t width, uint32_t height, b2rt_camera* out);
/* Camera of CudaRenderer::setViewpoint(origin, lookAt) (src/cudaRenderer.cu:1845-1870): eye at `origin`, viewing
 * along `look_at` (a direction), the reference's basis left = (0,1,0) x -lookAt, up = left x -lookAt
 * (src/cudaRenderer.cu:1592-1599; the screen's right axis is `left`, its up axis is -`up`) and, with fov_deg = 0, its
 * fixed frustum k = (u - .5, -(v - .5), 1) (src/cudaRenderer.cu:347), i.e. 2 atan(.5) = 53.13 degrees on both axes.
 * Fails for a look_at parallel to the y axis (the reference divides by zero there). */
int b2rt_camera_look_at(const float origin[3], const float look_at[3], float fov_deg, b2rt_camera* out);
/* Image files.  b2rt_save_png / b2rt_save_exr are plain host helpers (no device): RGBA8 words as b2rt_read_ldr
 * returns them / RGB fp32 triples as b2rt_read_hdr returns them, row 0 = bottom row of the image; the files store the
 * top row first.  b2rt_write_png = PathTracer::save_image (src/pathtracer.cpp:577-591: tone-mapped frame, flipped,
 * PNG); b2rt_write_exr writes the HDR frame as an uncompressed fp32 OpenEXR scanline file (the reference reads EXR
 * for its environment maps, src/main.cpp:38-70, and the C++ report tooling around it writes HDR frames). */
int b2rt_save_png(const char* path, const uint32_t* rgba8, uint32_t width, uint32_t height);
int b2rt_save_exr(const char* path, const float* rgb, uint32_t width, uint32_t height);
int b2rt_write_png(b2rt_renderer* r, const char* path);
int b2rt_write_exr(b2rt_renderer* r, const char* path);

#ifdef __cplusplus
}
#endif
#endif /* B2RT_H */
