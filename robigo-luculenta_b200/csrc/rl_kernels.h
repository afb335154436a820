// rl_kernels.h -- host-callable launchers of the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rl_b200.h"
#include "rl_scene_layout.h"

namespace rl {

// One batch of photons: the ids [first_photon, first_photon + n_photons) of the stream `seed`
// (one TraceUnit::render call, trace_unit.rs:151-168).
struct TraceSegment {
    uint64_t seed;
    uint64_t first_photon;
    uint64_t n_photons;
    rl_mapped_photon *records;        // device, n_photons entries, or nullptr
    unsigned long long *ray_counter;  // device, or nullptr
};
// Batches traced by one launch share the canvas and the accumulator.  Several segments: each
// fewer than 2^RL_SEGMENT_INDEX_BITS photons; one segment: any size (split into launches of 2^31).
#define RL_MAX_SEGMENTS 32
#define RL_SEGMENT_INDEX_BITS 27
struct TraceLaunch {
    uint32_t width, height;
    float4 *accum;                    // device, width*height float4, or nullptr (fused splat)
    uint32_t n_segments;
    TraceSegment seg[RL_MAX_SEGMENTS];
};

// Dynamic shared memory the trace kernels need for a scene.
size_t trace_smem_bytes(const DevScene &sc, int threads);
// K1: TraceUnit::render (+ PlotUnit::plot when accum != nullptr).
cudaError_t launch_trace(const DevScene &sc, const TraceLaunch &p, int sm_count, cudaStream_t st);
// K2: PlotUnit::plot over device records.
cudaError_t launch_splat(const rl_mapped_photon *records, uint64_t n, float4 *accum, uint32_t width,
                         uint32_t height, int sm_count, cudaStream_t st);
// xyzw (float4 per pixel) -> packed xyz (3 floats per pixel)
cudaError_t launch_pack_xyz(const float4 *accum, float *xyz, uint64_t n_pixels, cudaStream_t st);
// K3: GatherUnit::accumulate.  Sources: n_src padded buffers (float4/pixel),
// or one packed buffer (3 floats/pixel) when packed_src != nullptr.
cudaError_t launch_gather(float *acc, float *comp, const float4 *const *srcs, uint32_t n_src,
                          const float *packed_src, float4 *clear_or_null, uint64_t n_pixels,
                          int sm_count, cudaStream_t st);
// K4: TonemapUnit::tonemap.  moments = 2 doubles of scratch; exposure_out = 1 float.
// reference_fold: find_exposure as the reference's two sequential f32 folds (one thread) instead
// of the parallel f64 reduction.
cudaError_t launch_tonemap(const float *xyz, uint32_t width, uint32_t height, double *moments,
                           float *exposure_out, uint8_t *rgb, int sm_count, bool reference_fold, cudaStream_t st);

// probes
cudaError_t launch_debug_intersect(const DevScene &sc, const rl_ray *rays, uint64_t n, rl_hit *out,
                                   cudaStream_t st);
cudaError_t launch_debug_math(int fn, const float *in, const float *in2, uint64_t n, float *out,
                              cudaStream_t st);
cudaError_t launch_debug_tristimulus(const float *wl, uint64_t n, float *out, cudaStream_t st);
cudaError_t launch_debug_camera(const DevScene &sc, uint64_t seed, uint32_t width, uint32_t height,
                                uint64_t first, uint64_t n, rl_ray *rays, rl_mapped_photon *xy,
                                cudaStream_t st);

cudaError_t launch_debug_cull_check(const DevScene &sc, uint64_t seed, uint32_t width, uint32_t height,
                                    uint64_t first, uint64_t n, unsigned long long *rays,
                                    unsigned long long *mismatches, cudaStream_t st);

uint64_t kernel_launches();
void kernel_launches_reset();

}  // namespace rl
