/* drt.h — C ABI of libdartray_gpu.so, the B200 (sm_100a) ray-tracing core that replaces the
 * data-parallel hot path of DartRay's SamplerRenderer.
 *
 * The reference (brendan-duncan/dartray, 100 % Dart) has NO native boundary.  Its extension
 * mechanism is the Plugin registry (lib/core/plugin.dart:62-180) and the coarse seam is
 * Renderer.render(scene) (lib/core/renderer.dart:27-35, called from lib/dartray/dartray.dart:574).
 * Every entry point below replaces the work that one reference interface does behind that seam and
 * is what a `dart:ffi` binding in a Dart `GpuSamplerRenderer` would look up.
 *
 * Conventions
 *   - all functions are extern "C", plain pointers and sizes, blocking;
 *   - return 0 on success, a negative DRT_E_* code on failure; drt_last_error() gives the text
 *     (reference analogue: LogWarning + null / LogSevere exception, lib/core/log.dart:44-46);
 *   - the caller owns every input and output buffer; the library copies inputs during the call
 *     and never retains host pointers;
 *   - a drt_ctx is bound to ONE CUDA device and may be used from one host thread at a time
 *     (reference analogue: one isolate with a private scene, lib/dartray_web/render_isolate.dart:31-41);
 *   - there is NO CPU fallback: creating a context without a usable CUDA device fails.
 *
 * Primitive ids: the position in upload order — triangle k of drt_set_triangles is
 * id k, sphere j of drt_set_spheres is id ntris + j, disk j of drt_set_disks is id ntris + nspheres + j.  Every hit
 * record reports this id.
 */
#ifndef DRT_H_
#define DRT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRT_VERSION 100

enum {
  DRT_OK = 0,
  DRT_E_INVALID = -1,   /* bad argument */
  DRT_E_STATE = -2,     /* call order (e.g. trace before build) */
  DRT_E_CUDA = -3,      /* CUDA runtime error, text in drt_last_error */
  DRT_E_NODEVICE = -4,  /* no CUDA device: there is no CPU fallback */
  DRT_E_NOMEM = -5,
  DRT_E_UNSUPPORTED = -6 /* a reference feature that is not on the GPU path: the caller keeps the Dart renderer for it */
};

/* BVHAccel split methods, lib/accelerators/bvh_accel.dart:37-39 */
enum { DRT_SPLIT_MIDDLE = 0, DRT_SPLIT_EQUAL_COUNTS = 1, DRT_SPLIT_SAH = 2 };

typedef struct drt_ctx drt_ctx;

/* One closest-hit result: 16 bytes.  t is the reference's f64 tHit rounded to f32; b1/b2 are the
 * triangle barycentrics of lib/shapes/triangle.dart:77,87 (sphere: u,v of sphere.dart:119-121);
 * prim is the upload-order primitive id, -1 = miss (then t = +inf, b1 = b2 = 0). */
typedef struct drt_hit {
  float t, b1, b2;
  int32_t prim;
} drt_hit;

/* Traversal statistics of the last trace call (reference: the Stats counters of
 * lib/core/stats.dart:541-555 and the probes in bvh_accel.dart:106-135). */
typedef struct drt_counters {
  uint64_t rays;
  uint64_t nodes_visited; /* slab tests, bvh_accel.dart:125/187 */
  uint64_t prims_tested;  /* primitive tests, bvh_accel.dart:131/193 */
  uint64_t hits;
} drt_counters;

typedef struct drt_bvh_info {
  uint32_t n_nodes;      /* reference _LinearBVHNode count (bvh_accel.dart:82) */
  uint32_t n_prims;
  uint32_t n_leaves;
  uint32_t max_leaf_prims;
  uint32_t max_depth;
  uint64_t device_bytes; /* node + primitive records resident in HBM */
  double build_seconds;
} drt_bvh_info;

int drt_version(void);

/* Context on CUDA device `device_id`.  Returns NULL when no device is usable (message via
 * drt_last_error(NULL)).  DRT_DEVICE_NONE creates a host-only staging context that can hold a scene
 * and run the host side of drt_build_bvh / drt_bvh_export (BVH construction is host code, like the
 * reference's constructor); every query on it fails with DRT_E_NODEVICE. */
#define DRT_DEVICE_NONE (-1)
drt_ctx* drt_create(int device_id);
/* One context over several GPUs of the box: what lib/dartray_web/render_manager.dart:100-141 does with one isolate (and one
 * private copy of the scene) per image region, and what `taskNum / taskCount` (lib/dartray/dartray.dart:1009-1023) express.
 * Every drt_set_* call and drt_film_clear is applied to all devices; drt_build_bvh builds the tree ONCE on the host and uploads
 * its arrays to every device (one host thread per device); drt_render / drt_render_shard split their window over the devices
 * in interleaved 1024-pixel blocks (keyed sample streams: the union is exactly the one-device sample set), each device driven
 * by its own host thread, and the per-device films are then summed into the first device's over NVLink — one kernel reading
 * the peers' films through peer-mapped pointers (a staged cudaMemcpyPeer where peer mapping is unavailable) — so that
 * drt_film_read / drt_film_device return the whole image and drt_render_stats_get the rays of all devices.  Ray queries
 * (drt_trace_*) and drt_pixel_samples run on the first device.  SURVEY 8b proposed this signature for drt_create. */
drt_ctx* drt_create_multi(const int* device_ids, int n_devices);
int drt_device_count(const drt_ctx* ctx);
void drt_destroy(drt_ctx* ctx);
const char* drt_last_error(const drt_ctx* ctx);

/* Replaces TriangleMesh storage + refine (lib/shapes/triangle_mesh.dart:24-60,83-89): world-space
 * float32 positions and uint32 indices.  material_of_tri / light_of_tri (-1 = none) /
 * reverse_orientation_of_tri may be NULL (0 / -1 / 0). */
int drt_set_triangles(drt_ctx* ctx, const float* P, uint32_t nverts, const uint32_t* idx, uint32_t ntris,
                      const int32_t* material_of_tri, const int32_t* light_of_tri,
                      const uint8_t* reverse_orientation_of_tri);

/* Per-vertex shading attributes of the TriangleMeshes the soup was merged from: what Triangle.getUVs and
 * Triangle.getShadingGeometry read (lib/shapes/triangle.dart:246-262, 271-364; storage lib/shapes/triangle_mesh.dart:24-28).
 * N / S: nverts x 3 float32 in OBJECT space as the scene file gives them (the reference transforms them per hit with the
 * mesh's objectToWorld); uv: nverts x 2 float32.  Any of the three may be NULL.  mesh_of_tri: ntris mesh indices;
 * o2w / w2o: nmeshes x 16 row-major float32; mesh_flags bit 0 / 1 / 2: the mesh has N / S / uv (vertices of a mesh without
 * an attribute may hold anything in that array).  Call after drt_set_triangles (which clears it); nmeshes == 0 clears.
 * With uv present the triangle's dpdu / dpdv — and through them dg.nn — follow the file's parameterisation. */
int drt_set_mesh_shading(drt_ctx* ctx, const float* N, const float* S, const float* uv, const uint32_t* mesh_of_tri,
                         uint32_t nmeshes, const float* o2w, const float* w2o, const uint8_t* mesh_flags);

/* Replaces Sphere construction (lib/shapes/sphere.dart:24-32,314-323).  o2w / w2o: n x 16 row-major
 * float32 (Matrix4x4, lib/core/matrix4x4.dart:26); params: n x 4 doubles radius, zmin, zmax,
 * phimax(degrees) exactly as ParamSet hands them to Sphere.Create. */
int drt_set_spheres(drt_ctx* ctx, uint32_t n, const float* o2w, const float* w2o, const double* radius_zmin_zmax_phimax,
                    const int32_t* material_of_sphere, const int32_t* light_of_sphere,
                    const uint8_t* reverse_orientation_of_sphere);

/* Replaces Disk construction (lib/shapes/disk.dart:24-31,157-166).  Disks join the spheres in one quadric id
 * range: disk j is primitive ntris + nspheres + j, so call this AFTER drt_set_spheres (which resets the range).
 * params: n x 4 doubles height, radius, innerradius, phimax(degrees) as ParamSet hands them to Disk.Create. */
int drt_set_disks(drt_ctx* ctx, uint32_t n, const float* o2w, const float* w2o, const double* height_radius_inner_phimax,
                  const int32_t* material_of_disk, const int32_t* light_of_disk, const uint8_t* reverse_orientation_of_disk);

/* Replaces Cylinder / Cone / Paraboloid / Hyperboloid construction (lib/shapes/cylinder.dart:24-31,239-247,
 * cone.dart:23-27,216-222, paraboloid.dart:23-29,220-228, hyperboloid.dart:23-49,263-268).  They are appended to the
 * same quadric id range in call order (after drt_set_spheres / drt_set_disks).  params: n x 8 doubles, the ParamSet
 * values in the order of each Create():
 *   kind 2 cylinder    radius, zmin, zmax, phimax(degrees)
 *   kind 3 cone        height, radius, phimax
 *   kind 4 paraboloid  radius, zmin, zmax, phimax
 *   kind 5 hyperboloid p1.x, p1.y, p1.z, p2.x, p2.y, p2.z, phimax
 * Of these only the cylinder implements Shape.sample (cylinder.dart:230-240), so only it may be an area light's shape;
 * drt_set_lights rejects the others (the reference logs 'Unimplemented Shape.sample' and returns garbage). */
#define DRT_QUADRIC_CYLINDER 2
#define DRT_QUADRIC_CONE 3
#define DRT_QUADRIC_PARABOLOID 4
#define DRT_QUADRIC_HYPERBOLOID 5
int drt_set_quadrics(drt_ctx* ctx, int kind, uint32_t n, const float* o2w, const float* w2o, const double* params8,
                     const int32_t* material_of_quadric, const int32_t* light_of_quadric,
                     const uint8_t* reverse_orientation_of_quadric);

/* TransformedPrimitives (lib/core/primitive/transformed_primitive.dart:26-82): what DartRay.shape builds for a shape under an
 * animated CTM (lib/dartray/dartray.dart:404-452) and DartRay.objectInstance for an ObjectInstance (:505-546).
 *   objects    the `primitive` each one wraps: object i owns the geometric primitives object_prims[object_offsets[i] ..
 *              object_offsets[i + 1]) in the refined order its accelerator sees (one entry: the GeometricPrimitive itself, no
 *              accelerator); more than one: a BVHAccel with object_split / object_max_node_prims (NULL: sah, 1 — the constructor
 *              defaults an animated shape gets, bvh_accel.dart:41; an ObjectInstance uses the scene's accelerator parameters)
 *   instances  instance j wraps object instance_object[j] with AnimatedTransform(start, times[2j], end, times[2j + 1])
 *              (lib/core/animated_transform.dart:35-43): start / end are WORLD-TO-PRIMITIVE Transforms (m and mInv, row-major) —
 *              Inverse(CTM[0]), Inverse(CTM[1]); times NULL = (0, 1).  The library runs the constructor's Decompose and, per ray,
 *              worldToPrimitive.interpolate(ray.time) (:61-136) with the reference's float32 / binary64 arithmetic, and
 *              motionBounds (:183-200) for the world bound.
 * Primitive ids: instance j is top-level primitive n_triangles + n_quadrics + j in drt_set_build_order; object primitives are
 * geometric primitives that the build order leaves out (their shapes keep the transforms they were built with: identity for an
 * animated shape).  Hits report the geometric primitive.  The ray's time: the camera sample's time inside drt_render (every ray
 * a sample spawns inherits it, ray.dart:59); drt_set_ray_times for the drt_trace_* calls.  Scenes with instances are traced by
 * the literal-walk kernel (one thread per ray descends into the object with the transformed ray).  Area lights cannot be
 * instanced (dartray.dart:407-410,455-458 drop them with a warning); meshes inside objects carry no per-vertex N / S here.
 * Call before drt_set_build_order / drt_build_bvh.  n_instances == 0 removes them. */
int drt_set_instances(drt_ctx* ctx, uint32_t n_objects, const uint32_t* object_offsets, const uint32_t* object_prims,
                      const int32_t* object_split, const int32_t* object_max_node_prims, uint32_t n_instances,
                      const uint32_t* instance_object, const float* start_m, const float* start_minv, const float* end_m,
                      const float* end_minv, const double* times);

/* Ray i of the following drt_trace_* calls travels at times[i] (Ray.time, lib/core/ray.dart:37-38; read by TransformedPrimitives
 * only).  NULL: every ray at time 0. */
int drt_set_ray_times(drt_ctx* ctx, const double* times, uint64_t n);

/* Order in which BVHAccel sees the refined primitives (a permutation of primitive ids).  The
 * reference's Primitive.fullyRefine is LIFO (lib/core/primitive.dart:71-84), so a mesh's triangles
 * reach the builder in reverse order; the build's partition steps depend on it.  NULL = identity. */
int drt_set_build_order(drt_ctx* ctx, const uint32_t* prim_ids, uint32_t n);

/* Replaces BVHAccel(p, maxPrims, splitMethod) (lib/accelerators/bvh_accel.dart:41-91, 228-437,
 * Create :474-482): same tree, same leaf contents and in-leaf order, laid out for the GPU. */
int drt_build_bvh(drt_ctx* ctx, int split_method, int max_node_prims);
int drt_bvh_info_get(const drt_ctx* ctx, drt_bvh_info* out);

/* Export the tree in the REFERENCE's linear layout (bvh_accel.dart:419-437, 533-538) for topology
 * checks: bounds n_nodes x 6 (pMin xyz, pMax xyz), offset, n_primitives, axis per node, and the
 * reordered primitive list (upload ids).  Any pointer may be NULL. */
int drt_bvh_export(const drt_ctx* ctx, float* bounds, int32_t* offset, int32_t* n_primitives, int32_t* axis,
                   uint32_t* ordered_prim_ids);

/* Replaces Scene.intersect -> BVHAccel.intersect (lib/core/scene.dart:51-56,
 * lib/accelerators/bvh_accel.dart:101-165) with Triangle.intersect (lib/shapes/triangle.dart:44-98)
 * and Sphere.intersect (lib/shapes/sphere.dart:39-116) for a batch of n rays held in HOST memory.
 * ray_o_tmin: n x 4 float (origin xyz, minDistance); ray_d_tmax: n x 4 float (direction xyz,
 * maxDistance; +inf allowed).  hits: n records. */
int drt_trace_closest(drt_ctx* ctx, const float* ray_o_tmin, const float* ray_d_tmax, uint64_t n, drt_hit* hits);

/* Replaces Scene.intersectP -> BVHAccel.intersectP (bvh_accel.dart:167-226) with
 * Triangle.intersectP (triangle.dart:162-194) / Sphere.intersectP (sphere.dart:169-241).
 * occluded: n bytes, 1 = some primitive hit. */
int drt_trace_any(drt_ctx* ctx, const float* ray_o_tmin, const float* ray_d_tmax, uint64_t n, uint8_t* occluded);

/* Same two queries on buffers already resident in this context's device memory, launched on
 * `cuda_stream` (a cudaStream_t, NULL = default stream); asynchronous — the caller synchronises. */
int drt_trace_closest_device(drt_ctx* ctx, const void* d_ray_o_tmin, const void* d_ray_d_tmax, uint64_t n,
                             void* d_hits, void* cuda_stream);
int drt_trace_any_device(drt_ctx* ctx, const void* d_ray_o_tmin, const void* d_ray_d_tmax, uint64_t n,
                         void* d_occluded, void* cuda_stream);

/* All kernel variants make every decision with the reference's arithmetic and return identical
 * results.  FAST (default): the library picks FAST_Q or FAST_V1 by scene size (FAST_Q from 65536 primitives up).
 * FAST_Q: persistent warps over 64-byte quantised 4-wide nodes, conservative float32 box tests, postponed
 * leaves, every leaf box decided in float64 before its primitives are tested (falls back to FAST_V1 for
 * scenes whose coordinates cannot be quantised: non-finite or beyond 2^62).
 * FAST_V1: the first-generation kernel, 128-byte float32 nodes with a float32-filtered slab test.
 * EXACT_WALK: one thread per ray, float64 slab test at every node — the literal reference walk, kept as a
 * cross-check and as the counting kernel. */
enum { DRT_KERNEL_FAST = 0, DRT_KERNEL_EXACT_WALK = 1, DRT_KERNEL_FAST_V1 = 2, DRT_KERNEL_FAST_Q = 3 };
int drt_set_kernel_variant(drt_ctx* ctx, int variant);

/* When enabled the trace kernels also count slab and primitive tests (uses the EXACT_WALK kernel;
 * slower; off by default). */
int drt_set_counting(drt_ctx* ctx, int enabled);
int drt_get_counters(drt_ctx* ctx, drt_counters* out);

/* Device time of the kernels launched by the last host-buffer trace call (CUDA events), ms. */
double drt_last_kernel_ms(const drt_ctx* ctx);
/* Number of kernels this context has launched so far. */
uint64_t drt_kernel_launches(const drt_ctx* ctx);

/* ---------------------------------------------------------------------------------------------------
 * Render path: what Renderer.render(scene) does for a SamplerRenderer (lib/renderers/
 * sampler_renderer.dart:36-65,118-218) with the objects DartRay.worldEnd built
 * (lib/dartray/dartray.dart:549-764).  The caller flattens those objects into the arrays below.
 * ------------------------------------------------------------------------------------------------- */

/* Replaces MatteMaterial (lib/materials/matte_material.dart:41-65) with constant Kd / sigma
 * textures (lib/core/texture/constant_texture.dart:23-39).  kind: 0 = matte (the only kind on the
 * path, NULL = all 0); kd_rgb: n x 3; sigma: n (degrees, NULL = 0).  Material index = position. */
int drt_set_materials(drt_ctx* ctx, uint32_t n, const int32_t* kind, const float* kd_rgb, const float* sigma);

/* Materials as ordered BxDF lists: what Material.getBSDF builds with constant textures, for the materials whose
 * getBSDF only adds these BxDFs: matte (lib/materials/matte_material.dart:41-65), mirror
 * (mirror_material.dart:26-43), glass (glass_material.dart:26-52), plastic (plastic_material.dart:26-53), metal
 * (metal_material.dart:26-46), uber (uber_material.dart:27-75), shinymetal (shiny_metal_material.dart:42-76), substrate, and —
 * with drt_set_lobe_wrappers — translucent and mix.  Material i owns lobes [lobe_offsets[i],
 * lobe_offsets[i + 1]) in the order of its bsdf.add calls (BSDF.sample_f picks by position, bsdf.dart:68-79; at most 8,
 * bsdf.dart:253).  Per lobe:
 *   lobe_kind     0 Lambertian (lambertian.dart), 1 OrenNayar (oren_nayar.dart), 2 Microfacet with a Blinn distribution
 *                 (microfacet.dart, blinn.dart), 3 SpecularReflection, 4 SpecularTransmission, 5 FresnelBlend over an
 *                 Anisotropic distribution (fresnel_blend.dart, anisotropic.dart: SubstrateMaterial, substrate_material.dart:46-68)
 *                 with Rd in lobe_rgb, Rs in fresnel_eta and the exponents 1 / uroughness, 1 / vroughness (clamped to 10000 as
 *                 anisotropic.dart:30-37 clamps them) in lobe_scalars[0] and [1]
 *   lobe_rgb      R / T of the BxDF, already clamped and multiplied as the material does (n_lobes x 3)
 *   fresnel_kind  0 FresnelNoOp, 1 FresnelDielectric(ei, et), 2 FresnelConductor(eta, k) (fresnel_*.dart); NULL = all 0
 *   fresnel_eta, fresnel_k   conductor spectra as RGB (n_lobes x 3; NULL when no lobe uses a conductor)
 *   lobe_scalars  n_lobes x 3 doubles: {Blinn exponent after blinn.dart:24-28 | OrenNayar sigma in degrees, ei, et}
 *   lobe_kind 6 RegularHalfangleBRDF (regular_halfangle_brdf.dart) / 7 IrregularIsotropicBRDF (irregular_isotropic_brdf.dart), the
 *                 BxDFs of MeasuredMaterial: lobe_scalars[0] = index of the table given to drt_set_measured
 * Every integrator takes every combination; directlighting evaluates its SpecularReflect / SpecularTransmit recursion
 * (lib/core/integrator.dart:187-290) chain by chain up to maxdepth 17 (DRT_E_UNSUPPORTED beyond).  Replaces a previous
 * drt_set_materials and vice versa. */
int drt_set_material_lobes(drt_ctx* ctx, uint32_t n, const uint32_t* lobe_offsets, const int32_t* lobe_kind, const float* lobe_rgb,
                           const int32_t* fresnel_kind, const float* fresnel_eta, const float* fresnel_k,
                           const double* lobe_scalars);

/* The data a MeasuredMaterial holds once its file is loaded (lib/materials/measured_material.dart:76-205; loading and the
 * spectrum -> RGB conversion of .brdf files stay with the caller).  Table i = data[offsets[i] ...]:
 *   kind 0  regularHalfangleData of a .merl file: dims = (nThetaH, nThetaD, nPhiD) = (90, 90, 180) in the reference, 3 floats (RGB)
 *           per cell with the phi difference as the minor index (regular_halfangle_brdf.dart:66-71)
 *   kind 1  the IrregIsotropicBRDFSamples of a .brdf file: dims[0] samples of 6 floats, the BRDFRemap point (brdf_remap.dart:23-47)
 *           then the sample's RGB value; the reference finds the samples around a query through a KdTree (kdtree.dart:86-112),
 *           which visits exactly those closer than the search radius — here they are visited in the order given
 * Referenced by lobe kinds 6 / 7 of drt_set_material_lobes and by material program kind 11.  n_tables == 0 removes them. */
int drt_set_measured(drt_ctx* ctx, uint32_t n_tables, const int32_t* kind, const int32_t* dims, const uint64_t* offsets,
                     const float* data, uint64_t n_floats);

/* The two BxDF adapters of the reference around the lobes of the last drt_set_material_lobes, one entry per lobe in the same
 * order: wrap bit 0 = BRDFToBTDF(bxdf) (lib/core/reflection/brdf_to_btdf.dart: TranslucentMaterial's transmissive Lambertian and
 * Microfacet, translucent_material.dart:62-86), bit 1 = ScaledBxDF(.., scale) (scaled_bxdf.dart: MixMaterial scales the BxDFs of
 * its first material by `amount` and those of the second by 1 - amount, mix_material.dart:36-50); scale_rgb: n_lobes x 3
 * float32 (read where bit 1 is set).  As in the reference a ScaledBxDF answers pdf() with BxDF's cosine density
 * (scaled_bxdf.dart has no pdf override, bxdf.dart:84-88).  One level: a mix of mixes is not representable. */
int drt_set_lobe_wrappers(drt_ctx* ctx, uint32_t n_lobes, const int32_t* wrap, const float* scale_rgb);

/* Textures that read the hit point.  One node per Texture object of the scene (lib/core/texture.dart); children are
 * referenced by node index.  Replaces what Texture.evaluate(dg) does for:
 *   kind 0 ConstantTexture (lib/core/texture/constant_texture.dart)            value
 *   kind 1 ScaleTexture    (lib/textures/scale_texture.dart:26-33)              tex1, tex2
 *   kind 2 MixTexture      (lib/textures/mix_texture.dart:26-31)                tex1, tex2, amount (a float texture)
 *   kind 3 ImageTexture    (lib/textures/image_texture.dart:76-86)              mapping + image_*: MIPMap.lookup2, i.e. the
 *          trilinear lookup (lib/core/mipmap.dart:206-222) or the EWA filter (:224-339) over the box pyramid (:142-166)
 *   kind 4 CheckerboardTexture, dimension 2 (checkerboard_texture.dart:29-75)   mapping, tex1, tex2, aa_method (0 none, 1 closedform)
 *   kind 5 UVTexture       (uv_texture.dart:26-37)                              mapping
 *   kind 6 BilerpTexture   (bilerp_texture.dart:26-40)                          mapping, value = v00, value2 = v01, v10, v11
 *   kind 7 FBmTexture / 8 WrinkledTexture (fbm_texture.dart, wrinkled_texture.dart)   world_to_texture, aa_method = octaves, value[0] = roughness
 *   kind 9 WindyTexture    (windy_texture.dart:26-38)                           world_to_texture
 *   kind 10 MarbleTexture  (marble_texture.dart:27-66; spectrum only)           world_to_texture, aa_method = octaves, value = roughness, scale, variation
 *   kind 11 DotsTexture    (dots_texture.dart:26-52)                            mapping, tex1 = outside, tex2 = inside
 *   kind 12 CheckerboardTexture, dimension 3 (checkerboard_3d_texture.dart)     world_to_texture, tex1, tex2
 *   (7-10, 12 go through IdentityMapping3D, identity_mapping_3d.dart: world_to_texture is the Transform the mapping holds — the
 *   plugins hand it tex2world as it is; Noise / FBm / Turbulence: lib/core/texture.dart:40-140)
 * spectrum: 0 = Texture<double> (values are Dart doubles; value[0]), 1 = Texture<Spectrum> (float32 RGB per operation).
 * mapping: 0 UVMapping2D (lib/core/texture/uv_mapping_2d.dart; su, sv, du, dv), 1 SphericalMapping2D, 2 CylindricalMapping2D
 * (world_to_texture, row-major), 3 PlanarMapping2D (v1, v2, du = ds, dv = dt).
 * Images: `texels` holds every image's level 0 AS THE REFERENCE'S MIPMap CONSTRUCTOR LEFT IT (pyramid[0]: scale / gamma applied,
 * float images converted, resampled to power-of-two resolution, mipmap.dart:72-139) — image_channels (1 or 3) floats per texel
 * from float offset image_offset; the library rebuilds the pyramid, including what SpectrumImage's shared return object does to
 * the spectrum levels (see oracle/ref_texture.cpp).  image_wrap: 0 repeat, 1 black, 2 clamp. */
typedef struct drt_texture {
  int32_t kind, spectrum;
  int32_t tex1, tex2, amount; /* child node indices, -1 = none */
  int32_t mapping;
  int32_t image_width, image_height, image_channels, image_wrap, image_trilinear;
  int32_t aa_method;
  uint64_t image_offset;
  double value[3];
  double value2[9];
  double su, sv, du, dv;
  double max_anisotropy;
  float world_to_texture[16];
  float v1[3], v2[3];
} drt_texture;
int drt_set_textures(drt_ctx* ctx, uint32_t n, const drt_texture* nodes, const float* texels, uint64_t n_texel_floats);

/* Materials whose parameters are textures, or that carry a bump map (Material.Bump, lib/core/material.dart:35-88): one entry per
 * material of the last drt_set_material_lobes / drt_set_materials.  kind -1 keeps the material's flattened lobe list.  Otherwise
 * the library evaluates the textures at every hit (after DifferentialGeometry.computeDifferentials with the camera ray's
 * differentials, differential_geometry.dart:122-205, perspective_camera.dart:122-128, sampler_renderer.dart:166) and builds
 * the BSDF the material's getBSDF builds; tex[] by kind (node indices into drt_set_textures; every parameter is a node,
 * constants included):
 *   0 matte        Kd, sigma                         5 shinymetal   Ks, Kr, roughness
 *   1 mirror       Kr                                6 substrate    Kd, Ks, uroughness, vroughness
 *   2 glass        Kr, Kt, index                     7 translucent  Kd, Ks, reflect, transmit, roughness
 *   3 plastic      Kd, Ks, roughness                 8 uber         Kd, Ks, Kr, Kt, roughness, opacity, index
 *   4 metal        eta, k, roughness                 9 mix          amount; m1 / m2 = material indices (one level)
 *   10 subsurface / kdsubsurface   Kr, index  (their BSDF is SpecularReflection(Kr, FresnelDielectric(1, index)),
 *      subsurface_material.dart:52-69; the BSSRDF belongs to the dipole integrator, which is not on the path)
 *   11 measured    no textures; m1 = table of drt_set_measured (measured_material.dart:219-238: the table's kind picks
 *      RegularHalfangleBRDF or IrregularIsotropicBRDF)
 * bump: a float texture node or -1.  n == 0 removes the programs. */
typedef struct drt_material_program {
  int32_t kind;
  int32_t tex[8];
  int32_t bump;
  int32_t m1, m2;
} drt_material_program;
int drt_set_material_programs(drt_ctx* ctx, uint32_t n, const drt_material_program* programs);

/* Replaces scene.lights: DiffuseAreaLight (kind 0, lib/lights/diffuse_area_light.dart:44-70; L = Lemit
 * x scale) and PointLight (kind 1, lib/lights/point_light.dart:41-47; L = intensity, pos = world
 * position); kinds 2 / 3: see drt_set_spot_params below.  nsamples: per light (NULL = 1).  The ShapeSet of light i (lib/core/light/
 * shape_set.dart:26-50) is shape_prims[shape_offsets[i] .. shape_offsets[i+1]) — primitive ids in the
 * order the reference's refine loop leaves them.  Light index = position; drt_set_triangles /
 * drt_set_spheres refer to it through light_of_*. */
int drt_set_lights(drt_ctx* ctx, uint32_t n, const int32_t* kind, const float* L_rgb, const float* pos,
                   const int32_t* nsamples, const uint32_t* shape_offsets, const uint32_t* shape_prims);

/* The other delta lights of scene.lights go through the same call: DistantLight (kind 2,
 * lib/lights/distant_light.dart:24-48; L = radiance x scale, pos = lightDir = normalize(lightToWorld(from - to)), the
 * float32 Vector the light holds) and SpotLight (kind 3, lib/lights/spot_light.dart:24-70; L = intensity x scale,
 * pos = lightPos).  Spot lights additionally need, after drt_set_lights, their worldToLight matrix (n x 16 float32,
 * row-major; rows of other lights are ignored) and {cosTotalWidth, cosFalloffStart} (n x 2 doubles,
 * spot_light.dart:29-30). */
int drt_set_spot_params(drt_ctx* ctx, uint32_t n, const float* world_to_light, const double* cos_total_falloff);

/* Replaces InfiniteAreaLight construction (lib/lights/infinite_area_light.dart:37-69,276-316) for light `index`, which the last
 * drt_set_lights declared with kind 4 (L = the light's L * scale, nsamples as given).  rgb: width x height RGB float32 texels,
 * row-major — level 0 of the light's radiance MIPMap as the reference holds it (`radianceMap.pyramid[0]`): power-of-two
 * resolution (the reference's constructor has resampled the image, mipmap.dart:72-139) and already multiplied by L (:50-53);
 * a scene without "mapname" has the 1 x 1 white map (:66-68).  Radiance is lookup * L as the reference computes it (:240-242).
 * The library builds the box pyramid and the Distribution2D the light is importance-sampled from. */
int drt_set_infinite_light(drt_ctx* ctx, uint32_t index, int width, int height, const float* rgb, const float* light_to_world,
                           const float* world_to_light);

/* Replaces ProjectionLight (kind 5 of drt_set_lights; lib/lights/projection_light.dart:38-139) and GoniometricLight (kind 6;
 * lib/lights/goniometric_light.dart:37-86) construction for light `index`: pos / L of drt_set_lights are lightPos / intensity.
 * rgb: level 0 of the light's MIPMap (power-of-two resolution, row-major RGB float32) or NULL when the scene names no map
 * (projection: 1 inside the screen window; goniometric: 1).  world_to_light: 16 floats.  Projection lights also pass
 * light_projection (Transform.Perspective(fov, hither, yon), 16 floats), screen_window = screenX0, screenX1, screenY0,
 * screenY1 and hither as the constructor left them; NULL / ignored for goniometric lights.  Both are delta lights: intensity *
 * map lookup / distance^2. */
int drt_set_light_map(drt_ctx* ctx, uint32_t index, int width, int height, const float* rgb, const float* world_to_light,
                      const float* light_projection, const double* screen_window, double hither);

/* Participating media.  Replaces the VolumeRegion plugins (lib/volume_regions/homogenous_volume_region.dart:24-95,
 * exponential_density_region.dart:22-72, volume_grid.dart:22-109 over lib/core/volume/density_region.dart:22-86; several regions
 * = AggregateVolume, lib/core/volume/aggregate_volume.dart:23-103, as DartRay.worldEnd builds it, lib/dartray/dartray.dart:604-612).
 * kind: 0 homogeneous, 1 exponential, 2 volumegrid.  sigma_a / sigma_s / le: n x 3; g: n; p0_p1: n x 6 (the extent's two corners
 * in volume space); volume_to_world / world_to_volume: n x 16 row-major; exp_a_b: n x 2 and up_dir: n x 3 (exponential: density =
 * a * exp(-b * height along normalize(up)); may be NULL without such a region); grid_dims: n x 3 (nx, ny, nz) and density values
 * [density_offsets[i], density_offsets[i+1]) of region i, z-major as the scene file lists them (volumegrid; may be NULL without one).
 * n = 0 removes the volume: every transmittance is 1 and no random number is drawn for it, as in the reference. */
int drt_set_volumes(drt_ctx* ctx, uint32_t n, const int32_t* kind, const float* sigma_a_rgb, const float* sigma_s_rgb, const float* le_rgb,
                    const double* g, const float* p0_p1, const float* volume_to_world, const float* world_to_volume,
                    const double* exp_a_b, const float* up_dir, const int32_t* grid_dims, const uint64_t* density_offsets,
                    const double* density);
/* Replaces the VolumeIntegrator plugins: 0 = emission (the default of a scene without a VolumeIntegrator statement,
 * lib/dartray/render_options.dart:24-39; lib/volume_integrators/emission_integrator.dart:22-112), 1 = single
 * (single_scatter_integrator.dart:24-140).  Both march the camera ray through the regions in steps of `step_size`
 * (SamplerRenderer.Li returns T * Li + Lvi, lib/renderers/sampler_renderer.dart:67-98) and give the surface integrators their
 * transmittance() (integrator.dart:137,178, path_integrator.dart:116, whitted_integrator.dart:58).  Not on the GPU path (the call
 * that starts the render returns DRT_E_UNSUPPORTED): media together with the specular recursion of directlighting / whitted, or with
 * the halton / adaptive / bestcandidate samplers. */
int drt_set_volume_integrator(drt_ctx* ctx, int32_t kind, double step_size);

/* Arithmetic of the path integrator's shading kernels.  DRT_PRECISION_F64 (the default) evaluates every expression the way the Dart VM
 * does — float32 objects, binary64 expressions (lib/core/vector.dart:26-74, rgb_color.dart:23-169) — and reproduces the reference's
 * radiance per camera sample to ~1e-6.  DRT_PRECISION_F32 runs PathIntegrator.Li's vertex code (lib/surface_integrators/
 * path_integrator.dart:44-119, lib/core/integrator.dart:79-185) in float32 throughout: the same samples, queues and binary64
 * camera rays, but a sample's radiance now agrees with the reference's only within float32 rounding (the integrator's rays are
 * intersected in float32 too, so a hit within rounding of an edge may go the other way), which is inside what the Monte Carlo
 * estimate itself promises — per-pixel means within 3 sigma.  drt_trace_* are not affected.  It applies to the path integrator on
 * scenes without texture programs, media or object instances; every other render keeps the binary64 kernels whatever this is set
 * to. */
#define DRT_PRECISION_F64 0
#define DRT_PRECISION_F32 1
int drt_set_shading_precision(drt_ctx* ctx, int32_t precision);

/* Replaces PerspectiveCamera (lib/cameras/perspective_camera.dart:46-57 + lib/core/
 * projective_camera.dart:34-53): the two float32 row-major matrices the camera holds
 * (rasterToCamera, cameraToWorld.startTransform) and its lens / shutter scalars. */
int drt_set_camera(drt_ctx* ctx, const float raster_to_camera[16], const float camera_to_world[16], double lens_radius,
                   double focal_distance, double shutter_open, double shutter_close);

/* An animated camera: Camera.cameraToWorld is AnimatedTransform(cam2world[0], start, cam2world[1], end) (lib/core/camera.dart:27,
 * lib/dartray/dartray.dart:971-975); every camera ray goes through cameraToWorld.interpolate(sample.time)
 * (lib/core/animated_transform.dart:107-169).  camera_to_world_end: the end-time matrix (row-major; NULL = a static camera, the
 * default); the start-time matrix is drt_set_camera's.  Call after drt_set_camera. */
int drt_set_camera_motion(drt_ctx* ctx, const float* camera_to_world_end, double start_time, double end_time);

/* Which Camera plugin the matrices of drt_set_camera belong to: 0 = perspective (default), 1 = orthographic
 * (lib/cameras/orthographic_camera.dart:52-80: origin = rasterToCamera(Pras), direction +z; same lens model),
 * 2 = environment (lib/cameras/environment_camera.dart:42-52: direction from (theta, phi) of the raster
 * position; raster_to_camera is ignored). */
int drt_set_camera_kind(drt_ctx* ctx, int32_t kind);

/* Replaces ImageFilm's constructor (lib/film/image_film.dart:51-97): resolution, crop window
 * (x0, x1, y0, y1; NULL = 0,1,0,1), filter widths and the 16x16 table of filter.evaluate values
 * (image_film.dart:74-82) computed by the caller's Filter object.  Resets the film. */
int drt_set_film(drt_ctx* ctx, int32_t xres, int32_t yres, const double crop[4], double xwidth, double ywidth,
                 const float filter_table[256]);

/* Replaces the Sampler plugins lowdiscrepancy (kind 0, lib/samplers/low_discrepancy_sampler.dart;
 * spp rounded up to a power of two), stratified (kind 1, lib/samplers/stratified_sampler.dart; xs x ys
 * strata, jitter), random (kind 2, lib/samplers/random_sampler.dart) and halton (kind 3,
 * lib/samplers/halton_sampler.dart: spp * max(w, h)^2 indices of one global sequence per sample window, samples outside the
 * window rejected as the reference rejects them; streams keyed by the index) and adaptive (kind 4,
 * lib/samplers/adaptive_sampler.dart: xs = minsamples, ys = maxsamples, jitter = method (0 shape ids, 1 contrast); every pixel is
 * rendered with minsamples lowdiscrepancy samples and, where reportResults asks for it, again with maxsamples, the first visit's
 * samples being dropped) and bestcandidate (kind 5, needs drt_set_sample_table).  pixel_order / tile_size name
 * the PixelSampler (0 linear, 1 tile: lib/pixel_samplers/{linear,tile}_pixel_sampler.dart); with per-pixel keyed streams the
 * visiting order does not change any sample, so they only document the request.  seed keys every
 * stream (the reference seeds its single RNG with the task number, sampler_renderer.dart:137). */
int drt_set_sampler(drt_ctx* ctx, int32_t kind, int32_t xs, int32_t ys, int32_t spp, int32_t jitter, int32_t pixel_order,
                    int32_t tile_size, uint64_t seed);

/* The pattern of the bestcandidate sampler (kind 5 of drt_set_sampler; lib/samplers/best_candidate_sampler.dart:32-132): the
 * 4096 x 5 doubles of its _SAMPLE_TABLE (imageX, imageY, time, lensU, lensV per entry, :163-4258), which is data of the
 * reference and therefore travels through the ABI instead of living in this library.  The sampler walks every entry of the
 * pattern in every table tile the sample window touches (tile width 64 / sqrt(pixelsamples) pixels), shifts time / lens by the
 * tile's three random offsets — drawn, as in the reference, from a dart:math Random seeded with xTile + (yTile << 8) — rejects
 * samples outside the window exactly as :117-118 does, and draws the integrator arrays with LDShuffleScrambled1D / 2D. */
int drt_set_sample_table(drt_ctx* ctx, const double* table_4096x5, uint32_t n_entries);

/* Replaces the SurfaceIntegrator plugins path (kind 0, lib/surface_integrators/path_integrator.dart;
 * maxdepth), ambientocclusion (kind 1, ambient_occlusion_integrator.dart; nsamples rounded up to a
 * power of two, mindist, maxdist) and directlighting (kind 2, direct_lighting_integrator.dart;
 * strategy 0 = all, 1 = one) and whitted (kind 3, whitted_integrator.dart; maxdepth).  directlighting and whitted
 * evaluate their SpecularReflect / SpecularTransmit recursion (lib/core/integrator.dart:187-290) up to maxdepth 17. */
int drt_set_integrator(drt_ctx* ctx, int32_t kind, int32_t maxdepth, int32_t strategy, int32_t ao_nsamples,
                       double ao_mindist, double ao_maxdist);

/* Replaces _SamplerRendererTask.run (sampler_renderer.dart:118-218) for task task_num of task_count:
 * the task's sub-window of the sample extent (lib/dartray/dartray.dart:1009-1023, GetSubWindow
 * lib/core/common.dart:52-73).  Adds into the context's film; blocking. */
int drt_render(drt_ctx* ctx, int32_t task_num, int32_t task_count);

/* Same work, split for load balance instead of by sub-window: the sample extent's pixels are cut into
 * 1024-pixel blocks and shard s renders blocks s, s + n_shards, ...  The union over all shards is
 * exactly the sample set of drt_render(0, 1). */
int drt_render_shard(drt_ctx* ctx, int32_t shard, int32_t n_shards);

/* Camera samples in flight per wavefront batch (0 = default 16 Mi, about 10 GB of device memory for the largest renders; a
 * render with fewer samples allocates only what it needs). */
int drt_set_batch_slots(drt_ctx* ctx, uint64_t slots);

int drt_film_clear(drt_ctx* ctx);
/* left, top, width, height of the film's pixel window (image_film.dart:67-70) */
int drt_film_size(const drt_ctx* ctx, int32_t out[4]);
/* Replaces ImageFilm.writeImage (image_film.dart:268-299): rgb = width*height*3 floats (OutputImage.rgb);
 * xyz (width*height*3) / weight (width*height) are the raw accumulators; any pointer may be NULL. */
int drt_film_read(drt_ctx* ctx, float* rgb, float* xyz, float* weight);
/* The film accumulators in device memory: width*height x (X, Y, Z, weight) doubles — what a multi-GPU
 * caller sums across ranks (NCCL) before drt_film_read. */
int drt_film_device(drt_ctx* ctx, void** d_film, uint64_t* n_doubles);

/* The camera samples the sampler generates for pixel (x, y) (Sampler.getMoreSamples, lib/core/
 * sampler.dart:56): per sample imageX - x, imageY - y, lensU, lensV, time, then the integrator's 1D
 * arrays and 2D arrays in request order.  For parity tests of the sequences. */
int drt_pixel_samples(drt_ctx* ctx, int32_t x, int32_t y, float* out, int32_t cap, int32_t* n_samples,
                      int32_t* floats_per_sample);

/* Ray counters of the renders since the last drt_film_clear (lib/core/stats.dart:541-555). */
typedef struct drt_render_stats {
  uint64_t camera_samples;
  uint64_t closest_rays; /* camera + MIS + path-extension rays (Scene.intersect) */
  uint64_t shadow_rays;  /* Scene.intersectP */
  uint64_t zeroed_samples; /* NaN / negative / infinite radiance set to black, sampler_renderer.dart:181-193 */
} drt_render_stats;
int drt_render_stats_get(drt_ctx* ctx, drt_render_stats* out);

/* Where a render's GPU time and algorithmic work go (SURVEY 8d "per-sample work"; the reference keeps the same kind of
 * tallies in lib/core/stats.dart:527-640).  drt_set_render_profiling flags: DRT_PROFILE_TIME = record a CUDA event after
 * every kernel the following renders launch on the context's stream and accumulate the spans per kernel class;
 * DRT_PROFILE_WORK = also run, for every ray queue the renders trace, the counting walk of the reference's tree
 * (bvh_accel.dart:125/131/187/193: every slab test and every primitive test; slow, results unchanged).
 * drt_render_profile_get returns what accumulated since the last drt_film_clear. */
enum { DRT_PROFILE_TIME = 1, DRT_PROFILE_WORK = 2 };
enum { DRT_PK_TRACE_CLOSEST = 0, DRT_PK_TRACE_ANY = 1, DRT_PK_INTEGRATOR = 2, DRT_PK_SAMPLER = 3, DRT_PK_RESOLVE = 4,
       DRT_PK_FILM = 5, DRT_PK_OTHER = 6, DRT_PK_COUNT = 7 };
typedef struct drt_render_profile {
  double ms[DRT_PK_COUNT];         /* device time per kernel class (launch-to-launch spans on the render stream) */
  uint64_t launches[DRT_PK_COUNT];
  drt_counters closest;            /* reference work of the closest-hit rays (Scene.intersect) */
  drt_counters any;                /* reference work of the shadow rays (Scene.intersectP) */
} drt_render_profile;
int drt_set_render_profiling(drt_ctx* ctx, int flags);
int drt_render_profile_get(drt_ctx* ctx, drt_render_profile* out);

#ifdef __cplusplus
}
#endif
#endif /* DRT_H_ */
