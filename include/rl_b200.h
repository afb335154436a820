/*
 * rl_b200.h -- C ABI of the B200-native trace/plot/gather/tonemap engine.
 *
 * This is the drop-in boundary for the hot path of ruuda/robigo-luculenta:
 * the four pipeline "units" that app.rs / task_scheduler.rs drive.  The
 * reference has no FFI of its own (one binary crate of private modules), so
 * each entry point below names the Rust inherent method it replaces; thin
 * Rust shims with the same type names, methods and pub fields call these
 * (see INTEGRATION.md for the shim source a maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary;
 *   - every function returns int: RL_OK (0) or a negative rl_status; the
 *     message for the last failure on the calling thread is rl_last_error().
 *     The reference panics on failure (app.rs:107,163; gather_unit.rs:69-70),
 *     the Rust shim turns non-zero into `expect`;
 *   - a handle is bound to the CUDA device current at creation time and owns
 *     one stream; distinct handles may be driven from distinct host threads
 *     concurrently (the scheduler's contract, app.rs:104-109), one handle is
 *     never used from two threads at once;
 *   - host-facing byte layouts equal the reference's: MappedPhoton is
 *     4 x f32 (trace_unit.rs:23-37), a tristimulus buffer is packed
 *     3 x f32 per pixel, row-major (vector3.rs:20-25, plot_unit.rs:80-83),
 *     an image is packed RGB8 (tonemap_unit.rs:30,95-97);
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails with RL_ERR_CUDA.
 */
#ifndef RL_B200_H
#define RL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RL_ABI_VERSION 3

/* trace_unit.rs:67 -- photons per TraceUnit batch (1024 under cfg(test), :70) */
#define RL_BATCH_PHOTONS (1024u * 512u)
#define RL_TEST_BATCH_PHOTONS 1024u

typedef enum rl_status {
    RL_OK = 0,
    RL_ERR_INVALID = -1,     /* bad argument / malformed descriptor          */
    RL_ERR_UNSUPPORTED = -2, /* descriptor uses a shape the engine lacks     */
    RL_ERR_CUDA = -3,        /* CUDA runtime failure, or no device           */
    RL_ERR_IO = -4,          /* buffer.raw could not be read / written       */
    RL_ERR_NOMEM = -5
} rl_status;

/* ------------------------------------------------------------------ PODs */

typedef struct rl_vec3 { float x, y, z; } rl_vec3;        /* vector3.rs:20-25   */
typedef struct rl_quat { float x, y, z, w; } rl_quat;     /* quaternion.rs:19-25 */

/* trace_unit.rs:23-37 (add #[repr(C)] in the shim) */
typedef struct rl_mapped_photon {
    float x;            /* screen x in [-1, 1]                       */
    float y;            /* screen y in [-1/aspect, 1/aspect]         */
    float probability;  /* path contribution                         */
    float wavelength;   /* nm, [380, 780]                            */
} rl_mapped_photon;

/*
 * A surface node is the state of one reference geometry struct after its
 * constructor ran (geometry.rs); compound nodes reference two children by
 * index into the same node array (geometry.rs:361-407).  Leaves of a compound
 * are the reference's two Surface + Volume types: half-spaces (geometry.rs:
 * 90-128) and spheres (geometry.rs:186-267).
 */
typedef enum rl_surface_kind {
    RL_SURFACE_PLANE = 1,      /* geometry.rs:35-87    a = normal, b = offset                    */
    RL_SURFACE_HALFSPACE = 2,  /* geometry.rs:90-128   a = normal (points outside), b = offset   */
    RL_SURFACE_CIRCLE = 3,     /* geometry.rs:130-184  a = normal, b = position, s = radius^2    */
    RL_SURFACE_SPHERE = 4,     /* geometry.rs:186-267  a = position, s = radius^2                */
    RL_SURFACE_PARABOLOID = 5, /* geometry.rs:269-358  a = offset, b = normal, c = focal_point   */
    RL_SURFACE_COMPOUND = 6    /* geometry.rs:361-407  child[0] = surface1, child[1] = surface2  */
} rl_surface_kind;

typedef struct rl_surface {
    uint32_t kind;
    rl_vec3 a;
    rl_vec3 b;
    rl_vec3 c;
    float s;
    uint32_t child[2];
} rl_surface;

typedef enum rl_material_kind {
    RL_MATERIAL_BLACKBODY = 1,        /* emissive, material.rs:77-105   p0 = temperature, p1 = normalisation_factor */
    RL_MATERIAL_DIFFUSE_GREY = 2,     /* material.rs:109-130            p0 = reflectance                            */
    RL_MATERIAL_DIFFUSE_COLOURED = 3, /* material.rs:134-168            p0 = reflectance, p1 = wavelength, p2 = deviation */
    RL_MATERIAL_GLOSSY_MIRROR = 4,    /* material.rs:171-196            p0 = glossiness                             */
    RL_MATERIAL_SF10_GLASS = 5,       /* material.rs:199-261                                                         */
    RL_MATERIAL_SOAP_BUBBLE = 6       /* material.rs:265-306                                                         */
} rl_material_kind;

typedef struct rl_material {
    uint32_t kind;
    float p0, p1, p2;
} rl_material;

/* object.rs:20-31; list order is significant (scene.rs:51: first wins ties) */
typedef struct rl_object {
    uint32_t surface;  /* index of the object's root node in surfaces[] */
    rl_material material;
} rl_object;

/* camera.rs:21-44 */
typedef struct rl_camera {
    rl_vec3 position;
    float field_of_view;
    float focal_distance;
    float depth_of_field;
    float chromatic_abberation;
    rl_quat orientation;
} rl_camera;

/*
 * scene.rs:34 holds `fn(f32) -> Camera`, which cannot cross an FFI.  The three
 * camera models the engine evaluates per photon on the device:
 *   STATIC : `fixed` for every t.
 *   ORBIT  : the closed form of make_camera (app.rs:327-357):
 *              phi      = PI * (phi_base   + phi_rate   * t)
 *              alpha    = PI * (alpha_base + alpha_rate * t)
 *              distance = distance_base + distance_rate * t
 *              position = (cos a * sin p, cos a * cos p, sin a) * distance
 *              orientation = rot(0,0,-1, phi + PI) * rot(1,0,0, -alpha)
 *              focal_distance = distance * focal_factor
 *            field_of_view / depth_of_field / chromatic_abberation from `fixed`.
 *   KEYFRAMES : any other `fn(f32) -> Camera`, tabulated by the host: the
 *            camera of time t (a draw in [0, 1], trace_unit.rs:141) is
 *              keyframes[min(floor(t * n_keyframes), n_keyframes - 1)],
 *            every field taken from the keyframe.  Exact for a camera function
 *            that is piecewise constant on that grid, a discretisation of
 *            anything else (the host picks n_keyframes; 11 floats per frame).
 */
typedef enum rl_camera_kind { RL_CAMERA_STATIC = 1, RL_CAMERA_ORBIT = 2, RL_CAMERA_KEYFRAMES = 3 } rl_camera_kind;

typedef struct rl_camera_model {
    uint32_t kind;
    rl_camera fixed;
    float phi_base, phi_rate;
    float alpha_base, alpha_rate;
    float distance_base, distance_rate;
    float focal_factor;
    const rl_camera *keyframes;  /* KEYFRAMES only; copied by rl_scene_create */
    uint32_t n_keyframes;
} rl_camera_model;

/* scene.rs:23-36 flattened: what a `describe()` pass over Scene emits */
typedef struct rl_scene_desc {
    const rl_surface *surfaces;
    uint32_t n_surfaces;
    const rl_object *objects;
    uint32_t n_objects;
    rl_camera_model camera;
} rl_scene_desc;

/* --------------------------------------------------------------- handles */

typedef struct rl_scene rl_scene;
typedef struct rl_trace_unit rl_trace_unit;
typedef struct rl_plot_unit rl_plot_unit;
typedef struct rl_gather_unit rl_gather_unit;
typedef struct rl_tonemap_unit rl_tonemap_unit;

/* ------------------------------------------------------------- library   */

int rl_abi_version(void);
/* Message of the last failure on this thread ("" if none). */
const char *rl_last_error(void);
/* Number of visible CUDA devices (0 when there is none); never fails. */
int rl_device_count(void);
/* Kernels launched by this library since load / since the last reset. */
uint64_t rl_kernel_launch_count(void);
void rl_kernel_launch_count_reset(void);

/* ---------------------------------------------------------------- scene  */

/* Replaces Arc<Scene> (app.rs:63): validates the descriptor, uploads the
 * primitive / material tables and the camera model to the current device. */
int rl_scene_create(const rl_scene_desc *desc, rl_scene **out);
int rl_scene_destroy(rl_scene *scene);

/* ----------------------------------------------------------- TraceUnit   */

/* TraceUnit::new (trace_unit.rs:64-78).  `seed` keys the counter-based RNG
 * that stands in for rand::random (monte_carlo.rs:22-28 is unseeded). */
int rl_trace_unit_create(uint64_t id, uint32_t width, uint32_t height, uint64_t seed,
                         rl_trace_unit **out);
int rl_trace_unit_destroy(rl_trace_unit *unit);
/* Photons per render() call; default RL_BATCH_PHOTONS (trace_unit.rs:67). */
int rl_trace_unit_set_batch_size(rl_trace_unit *unit, uint64_t n_photons);
/* Bind the unit to a caller-owned cudaStream_t (NULL = its own stream). */
int rl_trace_unit_set_stream(rl_trace_unit *unit, void *cuda_stream);
/* TraceUnit::render (trace_unit.rs:151-168).  Photon ids of the batch are
 * taken from the scene's batch counter (one scene = one App, app.rs:63), so the
 * union of photons over any schedule of B render() calls on a scene is ids
 * [0, B * batch).  A batch of the reference's size (4.6 photons per thread of
 * a full grid) is launched as a few dozen small blocks that take only the
 * launch's share of each SM's block slots, by the number of other units'
 * batches in flight (task_scheduler.rs:95-96,127-182 keeps 3C units going):
 * the batches of the worker threads run side by side on every SM and cover
 * each other's tails ("small launches").  With RL_TRACE_GROUPS=1
 * in the environment the batches are instead queued to a per-scene dispatcher
 * thread that traces whatever is queued with one multi-segment launch
 * (measured slower, kept for experiments).  If `out` is non-NULL
 * it receives batch_size records (the shim's `mapped_photons` Vec) and the
 * call blocks; with NULL the records stay on the device for plot_device. */
int rl_trace_unit_render(rl_trace_unit *unit, const rl_scene *scene, rl_mapped_photon *out);
/* The same call without the wait: the kernel and the copy into `out` are
 * queued on the unit's stream and the call returns; `out` (page-locked by
 * rl_host_register, or the copy degrades to a blocking one) is valid after
 * rl_trace_unit_sync.  The shim makes `mapped_photons` a buffer that waits when
 * it is first read, so a worker thread hands the GPU a batch
 * and goes on to its next task: the GPU's queue is then as deep as the 3C trace
 * units of task_scheduler.rs:100, not as the C worker threads. */
int rl_trace_unit_render_async(rl_trace_unit *unit, const rl_scene *scene, rl_mapped_photon *out);
/* Copy the records of the last render out of the device (blocking); lets the
 * shim leave them there (`out` = NULL above) until host code really reads
 * `mapped_photons`.  `out` has room for `capacity` records: the call fails with
 * RL_ERR_INVALID if the last render left more; `out_count` (optional) receives
 * the number copied (0 after a fused render: nothing is copied). */
int rl_trace_unit_download(rl_trace_unit *unit, rl_mapped_photon *out, uint64_t capacity,
                           uint64_t *out_count);
/* Same, for an explicit photon-id range [first_photon, first_photon + n). */
int rl_trace_unit_render_range(rl_trace_unit *unit, const rl_scene *scene,
                               uint64_t first_photon, uint64_t n_photons,
                               rl_mapped_photon *out);
/* Fused path: trace [first, first+n) and splat straight into `plot`'s device
 * accumulator; no MappedPhoton record round-trip (trace_unit.rs:151-168 +
 * plot_unit.rs:87-95 in one kernel).  Asynchronous on the trace unit's stream;
 * the plot unit observes it through an event. */
int rl_trace_unit_render_fused(rl_trace_unit *unit, const rl_scene *scene, rl_plot_unit *plot,
                               uint64_t first_photon, uint64_t n_photons);
/* Rays traced (= Scene::intersect calls, scene.rs:39) by this unit so far. */
int rl_trace_unit_ray_count(rl_trace_unit *unit, uint64_t *out_rays);
int rl_trace_unit_sync(rl_trace_unit *unit);
/* Set the scene's batch counter used by rl_trace_unit_render: the next batch
 * is ids [next_batch * batch, (next_batch + 1) * batch). */
int rl_scene_batch_counter_reset(const rl_scene *scene, uint64_t next_batch);
/* What the scene's trace dispatcher has done so far: launches, and the
 * TraceUnit::render batches they carried (batches / launches = mean group). */
int rl_scene_dispatch_stats(const rl_scene *scene, uint64_t *out_launches, uint64_t *out_batches);

/* Bytes copied host -> device and device -> host by the entry points of this
 * ABI since the last reset (process-wide): scene tables, MappedPhoton batches,
 * tristimulus buffers, buffer.raw snapshots, RGB images.  What bench.py reports
 * as h2d/d2h bytes per step. */
void rl_transfer_counters(uint64_t *h2d_bytes, uint64_t *d2h_bytes);
void rl_transfer_counters_reset(void);

/* Host buffers of the units.  The reference's units own plain `Vec`s
 * (`mapped_photons`, trace_unit.rs:56,75-77; `tristimulus_buffer`,
 * plot_unit.rs:34,47 and gather_unit.rs:26,40) that the host passes around as
 * slices (app.rs:139,146,157).  They are allocated once per unit and never
 * resized, so the shim page-locks them once after allocation: copies to and
 * from a registered buffer are true DMA transfers that overlap the kernels of
 * the other units instead of staged, blocking copies through pageable memory.
 * Registration is optional (an unregistered buffer works, slower) and
 * idempotent per address; unregister before the Vec is dropped. */
int rl_host_register(void *ptr, size_t bytes);
int rl_host_unregister(void *ptr);

/* ------------------------------------------------------------ PlotUnit   */

/* PlotUnit::new (plot_unit.rs:43-53) */
int rl_plot_unit_create(uint64_t id, uint32_t width, uint32_t height, rl_plot_unit **out);
int rl_plot_unit_destroy(rl_plot_unit *unit);
int rl_plot_unit_set_stream(rl_plot_unit *unit, void *cuda_stream);
/* PlotUnit::plot (plot_unit.rs:87-95) on a host slice of photons. */
int rl_plot_unit_plot(rl_plot_unit *unit, const rl_mapped_photon *photons, uint64_t n);
/* PlotUnit::plot on the records a trace unit left on the device. */
int rl_plot_unit_plot_device(rl_plot_unit *unit, rl_trace_unit *trace);
/* PlotUnit::clear (plot_unit.rs:98-102) */
int rl_plot_unit_clear(rl_plot_unit *unit);
/* Copy out `tristimulus_buffer` (plot_unit.rs:34): width*height*3 floats. */
int rl_plot_unit_download(rl_plot_unit *unit, float *xyz);
/* Device address of the accumulator: width*height pixels of 4 floats
 * (X, Y, Z, 0), for zero-copy views (NCCL reduce, peer access). */
int rl_plot_unit_device_buffer(rl_plot_unit *unit, void **out_ptr, size_t *out_bytes);
int rl_plot_unit_sync(rl_plot_unit *unit);
/* Cross-process sharing of the accumulator (one process per GPU): export a
 * 64-byte handle (cudaIpcMemHandle_t) in the owning process, open it in the
 * gathering process to get a device pointer usable as a source of
 * rl_gather_unit_accumulate_device -- the cross-GPU sum is then fused into the
 * gather kernel as peer loads over NVLink. */
#define RL_IPC_HANDLE_BYTES 64
int rl_plot_unit_ipc_export(rl_plot_unit *unit, void *handle_out);
int rl_ipc_open(const void *handle, void **out_ptr);
int rl_ipc_close(void *ptr);

/* ---------------------------------------------------------- GatherUnit   */

/* GatherUnit::new (gather_unit.rs:35-46).  `resume_path` NULL = start from
 * zero; otherwise the file is read like gather_unit.rs:81-92 if it exists
 * (the reference's fixed name is "buffer.raw"). */
int rl_gather_unit_create(uint32_t width, uint32_t height, const char *resume_path,
                          rl_gather_unit **out);
int rl_gather_unit_destroy(rl_gather_unit *unit);
int rl_gather_unit_set_stream(rl_gather_unit *unit, void *cuda_stream);
/* GatherUnit::accumulate (gather_unit.rs:49-64) from a host tristimulus
 * slice of width*height*3 floats. */
int rl_gather_unit_accumulate(rl_gather_unit *unit, const float *xyz);
/* accumulate(&plot.tristimulus_buffer) followed by plot.clear()
 * (app.rs:145-148), fused in one pass over the plot unit's device buffer. */
int rl_gather_unit_accumulate_plot(rl_gather_unit *unit, rl_plot_unit *plot, int clear_plot);
/* Same Kahan step from `n_buffers` device accumulators in PlotUnit layout
 * (4 floats per pixel); pointers may be peer-device memory mapped into this
 * process -- the cross-GPU reduce fused into the gather. */
int rl_gather_unit_accumulate_device(rl_gather_unit *unit, const void *const *xyzw_buffers,
                                     uint32_t n_buffers);
/* GatherUnit::save (gather_unit.rs:68-78): accumulator then compensation,
 * 12 raw bytes per pixel each, no header -> 24*w*h bytes.
 * The host calls this after every gather (app.rs:151).  The call snapshots
 * the two buffers on the device, in the unit's stream order, and returns; a
 * writer thread of the unit copies the newest snapshot to page-locked host
 * memory, puts it into `path.tmp` and renames it over `path`, so the file
 * always holds one complete snapshot (the reference truncates and rewrites
 * in place).  A newer snapshot replaces one that has not been written yet,
 * and the writer starts at most one file per save interval (default 0.1 s,
 * rl_gather_unit_set_save_interval; 0 = as fast as the files can be
 * written): a GPU gathers a hundred times a second where the reference's
 * CPU host gathers every few seconds, and checkpoints that close together
 * are only host traffic.  The file is never further behind the latest save
 * than that interval plus one write.  The first save to a path is written
 * before the call returns, so an unwritable path fails here ("failed to
 * open file", gather_unit.rs:69); a later write failure is returned by the
 * next save or flush.  rl_gather_unit_flush waits for the file to be
 * current (without waiting for the interval); rl_gather_unit_load and
 * rl_gather_unit_destroy flush first. */
int rl_gather_unit_save(rl_gather_unit *unit, const char *path);
int rl_gather_unit_set_save_interval(rl_gather_unit *unit, double seconds);
int rl_gather_unit_flush(rl_gather_unit *unit);
/* GatherUnit::read (gather_unit.rs:81-92); a short file is not an error
 * (read.rs:20-32), a missing file is RL_ERR_IO here. */
int rl_gather_unit_load(rl_gather_unit *unit, const char *path);
/* Copy out `tristimulus_buffer` (gather_unit.rs:26); compensation optional. */
int rl_gather_unit_download(rl_gather_unit *unit, float *xyz, float *compensation_or_null);
int rl_gather_unit_sync(rl_gather_unit *unit);

/* --------------------------------------------------------- TonemapUnit   */

/* TonemapUnit::new (tonemap_unit.rs:43-51) */
int rl_tonemap_unit_create(uint32_t width, uint32_t height, rl_tonemap_unit **out);
int rl_tonemap_unit_destroy(rl_tonemap_unit *unit);
int rl_tonemap_unit_set_stream(rl_tonemap_unit *unit, void *cuda_stream);
/* TonemapUnit::tonemap (tonemap_unit.rs:73-100) on a host slice; writes
 * width*height*3 bytes to `rgb` (the shim's `rgb_buffer`). */
int rl_tonemap_unit_tonemap(rl_tonemap_unit *unit, const float *xyz, uint8_t *rgb);
/* tonemap(&gather.tristimulus_buffer) (app.rs:157) without the host trip. */
int rl_tonemap_unit_tonemap_gather(rl_tonemap_unit *unit, rl_gather_unit *gather, uint8_t *rgb);
/* The exposure (`max_intensity`, tonemap_unit.rs:55-69) of the last call. */
int rl_tonemap_unit_last_exposure(rl_tonemap_unit *unit, float *out);
/* How find_exposure (tonemap_unit.rs:55-69) sums the pixels.
 * RL_EXPOSURE_F64_REDUCTION (default): parallel, in f64 -- the exposure agrees with the reference's
 *   to ~1e-4 relative, and a near-constant image gets a finite exposure.
 * RL_EXPOSURE_REFERENCE_FOLD: the reference's two sequential f32 folds in pixel order, by one
 *   thread (a few ms per megapixel; the reference tone-maps once per 30 s): the exposure is
 *   bit-equal to the reference's arithmetic, including the NaN it yields when the f32 sums make
 *   the variance round negative (near-constant images go black, as they do in the reference). */
typedef enum rl_exposure_mode { RL_EXPOSURE_F64_REDUCTION = 0, RL_EXPOSURE_REFERENCE_FOLD = 1 } rl_exposure_mode;
int rl_tonemap_unit_set_exposure_mode(rl_tonemap_unit *unit, int mode);

/* ------------------------------------------------------------- debugging */
/*
 * Probes that run single device functions of the path over arrays, so the
 * parity tests can compare them with the oracle one function at a time.
 */
typedef struct rl_ray { rl_vec3 origin; rl_vec3 direction; float wavelength; float probability; } rl_ray; /* ray.rs:19-33 */
typedef struct rl_hit {                 /* intersection.rs:19-32 + object index */
    int32_t object;                     /* -1 = miss */
    float distance;
    rl_vec3 position, normal, tangent;
} rl_hit;
/* Scene::intersect (scene.rs:39-60) for n host rays. */
int rl_debug_intersect(const rl_scene *scene, const rl_ray *rays, uint64_t n, rl_hit *out);
/* Device math of the path: fn = 0 sin, 1 cos, 2 exp, 3 acos (f32);
 * 4 = Planck `boltzmann` (material.rs:61-74) with in2 = temperature,
 * 5 = SF10 index of refraction (material.rs:203-213). */
int rl_debug_math(int fn, const float *in, const float *in2_or_null, uint64_t n, float *out);
/* cie1931::get_tristimulus (cie1931.rs:20-48): out = 3 floats per input. */
int rl_debug_tristimulus(const float *wavelengths, uint64_t n, float *out_xyz);
/* Camera ray of photon ids [first, first+n): TraceUnit::render's draws +
 * render_camera_ray up to camera.get_ray (trace_unit.rs:136-158). */
int rl_debug_camera_rays(const rl_scene *scene, uint64_t seed, uint32_t width, uint32_t height,
                         uint64_t first_photon, uint64_t n, rl_ray *out_rays,
                         rl_mapped_photon *out_xy);

/* Self-check of the result-preserving culls: traces photon ids [first, first+n)
 * with the brute-force Scene::intersect (every primitive, reference
 * arithmetic) and evaluates the culled intersect beside it on every ray;
 * reports the rays compared and how many differed in object, distance bits or
 * primitive (must be 0). */
int rl_debug_cull_check(const rl_scene *scene, uint64_t seed, uint32_t width, uint32_t height,
                        uint64_t first_photon, uint64_t n, uint64_t *out_rays,
                        uint64_t *out_mismatches);

#ifdef __cplusplus
}
#endif
#endif /* RL_B200_H */
