/*
 * oracle.h -- C API of the CPU oracle (test infrastructure; see oracle.cpp).
 * Scene descriptors and record layouts are the ones of include/rl_b200.h.
 */
#ifndef ORC_ORACLE_H
#define ORC_ORACLE_H

#include <stdint.h>
#include "../include/rl_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MATH_LIBM 0 /* glibc sinf/cosf/expf/...: what the Rust binary calls  */
#define ORC_MATH_SPEC 1 /* the specified polynomial math the CUDA path also uses */

typedef struct orc_counters {
    uint64_t photons;
    uint64_t rays;            /* Scene::intersect calls (scene.rs:39)            */
    uint64_t primitive_tests; /* leaf Surface::intersect calls, if requested     */
    uint64_t emissive_hits;
    uint64_t max_bounces;
} orc_counters;

/* TraceUnit::render over photon ids [first, first+n) (trace_unit.rs:151-168). */
int orc_trace(const rl_scene_desc *desc, uint64_t seed, uint32_t width, uint32_t height,
              uint64_t first_photon, uint64_t n, int math_mode, int count_tests,
              rl_mapped_photon *out, orc_counters *counters);
/* PlotUnit::plot into a packed w*h*3 buffer (plot_unit.rs:87-95). */
int orc_plot(uint32_t width, uint32_t height, const rl_mapped_photon *photons, uint64_t n,
             float *xyz);
/* GatherUnit::accumulate (gather_unit.rs:49-64). */
int orc_gather_accumulate(float *acc, float *comp, const float *px, uint64_t n_pixels);
/* TonemapUnit::find_exposure / tonemap (tonemap_unit.rs:55-100).  Pass NaN as
 * exposure to have it computed the reference's way. */
int orc_find_exposure(uint32_t width, uint32_t height, const float *xyz, float *out);
int orc_tonemap(uint32_t width, uint32_t height, const float *xyz, int math_mode,
                float exposure_or_nan, uint8_t *rgb);
/* Scene::intersect (scene.rs:39-60) for n rays. */
int orc_intersect(const rl_scene_desc *desc, const rl_ray *rays, uint64_t n, rl_hit *out);
/* fn: 0 sin 1 cos 2 exp 3 acos 4 boltzmann(in=nm, in2=K) 5 SF10 ior 6 ln 7 pow 8 tan */
int orc_math(int fn, int math_mode, const float *in, const float *in2, uint64_t n, float *out);
int orc_blackbody_intensity(float temperature, float normalisation, int math_mode,
                            const float *wavelengths, uint64_t n, float *out);
int orc_tristimulus(const float *wavelengths, uint64_t n, float *out_xyz);
int orc_camera_rays(const rl_scene_desc *desc, uint64_t seed, uint32_t width, uint32_t height,
                    uint64_t first_photon, uint64_t n, int math_mode, rl_ray *out_rays,
                    rl_mapped_photon *out_xy);
int orc_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
               uint32_t *out4);
int orc_draws(uint64_t seed, uint64_t photon, uint32_t n, const uint8_t *half_open, float *out);
/* Multi-threaded trace+plot+gather in the reference's pipeline shape. */
int orc_render_mt(const rl_scene_desc *desc, uint64_t seed, uint32_t width, uint32_t height,
                  uint64_t first_photon, uint64_t n_photons, uint64_t batch, int threads,
                  int math_mode, float *xyz_out, orc_counters *counters,
                  double *seconds_trace_plot);
int orc_dump_rays(const rl_scene_desc *desc, uint64_t seed, uint32_t width, uint32_t height,
                  uint64_t first_photon, uint64_t n, uint64_t cap, rl_ray *out, uint64_t *n_out);
int orc_hardware_threads(void);

#ifdef __cplusplus
}
#endif
#endif
