/* A plain-C client of include/drt.h (no Python, no CUDA headers): drives libdartray_gpu.so in the call order of the Dart shim
 * (dart/lib/gpu/gpu_sampler_renderer.dart: create, scene arrays, build order, drt_build_bvh, materials, lights, camera, film,
 * sampler, integrator, drt_render, drt_film_size, drt_film_read, drt_destroy) on a scene read from a flat blob file, and writes
 * the film.  The image has no Dart SDK, so this harness is what stands in for the shim at run time: it shows the header
 * compiles as C and that a non-Python client gets the same film (tests/test_abi_harness.py compares it with the ctypes path).
 *
 *   gcc -std=c99 -O1 -I include -o abi_harness tools/abi_harness.c -L dartray_b200 -ldartray_gpu -Wl,-rpath,$PWD/dartray_b200
 *   abi_harness scene.blob film.out
 *
 * Blob file: int32 count, then per entry: char name[32], int64 nbytes, data. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "drt.h"

/* the plain-data records of the texture ABI, as a C compiler lays them out (host.TEX_DTYPE / PROG_DTYPE assert the same in Python) */
typedef char drt_texture_is_280_bytes[sizeof(drt_texture) == 280 ? 1 : -1];
typedef char drt_material_program_is_48_bytes[sizeof(drt_material_program) == 48 ? 1 : -1];

typedef struct { char name[32]; int64_t nbytes; void* data; } blob;
static blob* g_blobs;
static int g_nblobs;

static const blob* find(const char* name) {
  for (int i = 0; i < g_nblobs; ++i)
    if (strcmp(g_blobs[i].name, name) == 0) return &g_blobs[i];
  return NULL;
}
static const void* ptr(const char* name) { const blob* b = find(name); return (b && b->nbytes) ? b->data : NULL; }
static int64_t count(const char* name, int64_t item) { const blob* b = find(name); return b ? b->nbytes / item : 0; }
static double num(const char* name, int i) { return ((const double*)find(name)->data)[i]; }

#define CK(call)                                                                  \
  do {                                                                            \
    int rc_ = (call);                                                             \
    if (rc_ != DRT_OK) {                                                          \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, drt_last_error(ctx));          \
      return 2;                                                                   \
    }                                                                             \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s scene.blob film.out\n", argv[0]); return 1; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  int32_t n = 0;
  if (fread(&n, 4, 1, f) != 1) return 1;
  g_blobs = (blob*)calloc((size_t)n, sizeof(blob));
  g_nblobs = n;
  for (int i = 0; i < n; ++i) {
    if (fread(g_blobs[i].name, 32, 1, f) != 1 || fread(&g_blobs[i].nbytes, 8, 1, f) != 1) return 1;
    g_blobs[i].data = malloc((size_t)g_blobs[i].nbytes + 1);
    if (g_blobs[i].nbytes && fread(g_blobs[i].data, (size_t)g_blobs[i].nbytes, 1, f) != 1) return 1;
  }
  fclose(f);

  if (drt_version() != DRT_VERSION) { fprintf(stderr, "header / library version mismatch\n"); return 1; }
  drt_ctx* ctx = drt_create((int)num("device", 0));
  if (!ctx) { fprintf(stderr, "drt_create: %s\n", drt_last_error(NULL)); return 3; }

  /* _flattenScene: triangles, spheres, then disks — ids follow in that order (gpu_sampler_renderer.dart:376-445) */
  CK(drt_set_triangles(ctx, (const float*)ptr("P"), (uint32_t)count("P", 12), (const uint32_t*)ptr("idx"), (uint32_t)count("idx", 12),
                       (const int32_t*)ptr("tri_mat"), (const int32_t*)ptr("tri_light"), (const uint8_t*)ptr("tri_rev")));
  CK(drt_set_spheres(ctx, (uint32_t)count("sph_params", 32), (const float*)ptr("sph_o2w"), (const float*)ptr("sph_w2o"),
                     (const double*)ptr("sph_params"), (const int32_t*)ptr("sph_mat"), (const int32_t*)ptr("sph_light"),
                     (const uint8_t*)ptr("sph_rev")));
  if (count("dsk_params", 32))
    CK(drt_set_disks(ctx, (uint32_t)count("dsk_params", 32), (const float*)ptr("dsk_o2w"), (const float*)ptr("dsk_w2o"),
                     (const double*)ptr("dsk_params"), (const int32_t*)ptr("dsk_mat"), (const int32_t*)ptr("dsk_light"),
                     (const uint8_t*)ptr("dsk_rev")));
  CK(drt_set_build_order(ctx, (const uint32_t*)ptr("order"), (uint32_t)count("order", 4)));
  CK(drt_build_bvh(ctx, (int)num("bvh", 0), (int)num("bvh", 1)));
  drt_bvh_info info;
  CK(drt_bvh_info_get(ctx, &info));

  /* materials: the matte table (drt_set_materials), or BxDF lists plus the texture nodes and material programs of a scene whose
   * materials read the hit point (gpu_sampler_renderer.dart: the materials block; gpu_textures.dart) */
  if (find("mat_lobe_offsets")) {
    CK(drt_set_material_lobes(ctx, (uint32_t)count("mat_lobe_offsets", 4) - 1, (const uint32_t*)ptr("mat_lobe_offsets"),
                              (const int32_t*)ptr("lobe_kind"), (const float*)ptr("lobe_rgb"), (const int32_t*)ptr("lobe_fresnel"),
                              (const float*)ptr("lobe_eta"), (const float*)ptr("lobe_k"), (const double*)ptr("lobe_scalars")));
    if (find("tex_nodes")) {
      CK(drt_set_textures(ctx, (uint32_t)count("tex_nodes", sizeof(drt_texture)), (const drt_texture*)ptr("tex_nodes"),
                          (const float*)ptr("tex_texels"), (uint64_t)count("tex_texels", 4)));
      CK(drt_set_material_programs(ctx, (uint32_t)count("mat_programs", sizeof(drt_material_program)),
                                   (const drt_material_program*)ptr("mat_programs")));
    }
  } else {
    CK(drt_set_materials(ctx, (uint32_t)count("mat_kind", 4), (const int32_t*)ptr("mat_kind"), (const float*)ptr("mat_kd"),
                         (const float*)ptr("mat_sigma")));
  }
  CK(drt_set_lights(ctx, (uint32_t)count("light_kind", 4), (const int32_t*)ptr("light_kind"), (const float*)ptr("light_L"),
                    (const float*)ptr("light_pos"), (const int32_t*)ptr("light_nsamples"), (const uint32_t*)ptr("light_shape_offsets"),
                    (const uint32_t*)ptr("light_shape_prims")));
  CK(drt_set_camera(ctx, (const float*)ptr("raster_to_camera"), (const float*)ptr("camera_to_world"), num("camera", 0), num("camera", 1),
                    num("camera", 2), num("camera", 3)));
  CK(drt_set_camera_kind(ctx, (int32_t)num("camera", 4)));
  CK(drt_set_film(ctx, (int32_t)num("film", 0), (int32_t)num("film", 1), (const double*)ptr("crop"), num("film", 2), num("film", 3),
                  (const float*)ptr("filter_table")));
  CK(drt_set_sampler(ctx, (int32_t)num("sampler", 0), (int32_t)num("sampler", 1), (int32_t)num("sampler", 2), (int32_t)num("sampler", 3),
                     (int32_t)num("sampler", 4), (int32_t)num("sampler", 5), (int32_t)num("sampler", 6), (uint64_t)num("sampler", 7)));
  CK(drt_set_integrator(ctx, (int32_t)num("integrator", 0), (int32_t)num("integrator", 1), (int32_t)num("integrator", 2),
                        (int32_t)num("integrator", 3), num("integrator", 4), num("integrator", 5)));

  /* GpuSamplerRenderer.render: one task, then the film (gpu_sampler_renderer.dart:274-290) */
  CK(drt_render(ctx, 0, 1));
  int32_t sz[4];
  CK(drt_film_size(ctx, sz));
  const size_t px = (size_t)sz[2] * (size_t)sz[3];
  float* rgb = (float*)malloc(px * 3 * sizeof(float));
  float* wt = (float*)malloc(px * sizeof(float));
  CK(drt_film_read(ctx, rgb, NULL, wt));
  drt_render_stats st;
  CK(drt_render_stats_get(ctx, &st));
  FILE* o = fopen(argv[2], "wb");
  if (!o) { perror(argv[2]); return 1; }
  fwrite(sz, 4, 4, o);
  fwrite(rgb, sizeof(float), px * 3, o);
  fwrite(wt, sizeof(float), px, o);
  fclose(o);
  printf("abi_harness: %u primitives, %u BVH nodes, film %dx%d, %llu camera samples, %llu closest + %llu shadow rays, %llu kernel launches\n",
         info.n_prims, info.n_nodes, sz[2], sz[3], (unsigned long long)st.camera_samples, (unsigned long long)st.closest_rays,
         (unsigned long long)st.shadow_rays, (unsigned long long)drt_kernel_launches(ctx));
  drt_destroy(ctx);
  return 0;
}
