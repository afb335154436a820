/*
 * rl_host.h -- host-only scene builders (librl_host.so).
 *
 * Host-side mirrors of the reference's geometry / material constructors and of
 * App::set_up_scene (app.rs:166-363) that emit the flattened descriptors
 * include/rl_b200.h takes.  They are INPUT GENERATORS for tests and benchmarks:
 * no CUDA, no device code, not part of the device path -- which is why they
 * live in a library of their own (a process that only needs a scene
 * description, such as the CPU reference arm of bench.py, never maps the
 * product library).  In the integration the Rust host emits the descriptor
 * itself (`describe()`).
 */
#ifndef RL_HOST_H
#define RL_HOST_H

#include "rl_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rl_scene_builder rl_scene_builder;

typedef enum rl_builtin_scene {
    RL_SCENE_C1_SPHERE_PLANE = 1,  /* 1 diffuse sphere + emissive plane, static camera     */
    RL_SCENE_C2_BUILTIN = 2,       /* app.rs:166-363, 339 objects, orbit camera            */
    RL_SCENE_C3_PRISM = 3,         /* SF10 prism + emissive circle + grey floor            */
    RL_SCENE_C4_SPHERES = 4,       /* 4096 random spheres (param = sphere count, 0 = 4096) */
    RL_SCENE_C6_LENSES = 6         /* compounds over spheres and half-spaces (lens, dome, ...) */
} rl_builtin_scene;

int rl_scene_builder_create(rl_scene_builder **out);
int rl_scene_builder_destroy(rl_scene_builder *b);
int rl_scene_builder_builtin(rl_scene_builder *b, int which, uint32_t param);
/* Primitive constructors; each returns the new surface node index (>= 0). */
int rl_scene_builder_plane(rl_scene_builder *b, rl_vec3 normal, rl_vec3 offset);
int rl_scene_builder_circle(rl_scene_builder *b, rl_vec3 normal, rl_vec3 position, float radius);
int rl_scene_builder_sphere(rl_scene_builder *b, rl_vec3 position, float radius);
/* SpacePartitioning::new (geometry.rs:99-106) and Compound::new (geometry.rs:369-378): the
 * children of a compound are Volumes -- half-spaces, spheres or compounds of those. */
int rl_scene_builder_halfspace(rl_scene_builder *b, rl_vec3 normal, rl_vec3 offset);
int rl_scene_builder_compound(rl_scene_builder *b, uint32_t surface1, uint32_t surface2);
int rl_scene_builder_paraboloid(rl_scene_builder *b, rl_vec3 normal, rl_vec3 offset,
                                float focal_distance);
int rl_scene_builder_prism(rl_scene_builder *b, rl_vec3 axis, rl_vec3 offset, float edge_length,
                           float angle, float height);
int rl_scene_builder_hexagonal_prism(rl_scene_builder *b, rl_vec3 axis, rl_vec3 offset,
                                     float edge_length, float bevel_size, float angle,
                                     float height);
/* BlackBodyMaterial::new (material.rs:92-97) -> material record. */
int rl_material_blackbody(float kelvins, float intensity, rl_material *out);
/* Object::new (object.rs:35-42); returns the object index. */
int rl_scene_builder_object(rl_scene_builder *b, uint32_t surface, rl_material material);
int rl_scene_builder_camera(rl_scene_builder *b, const rl_camera_model *camera);
/* Borrow the descriptor; valid until the builder changes or is destroyed. */
int rl_scene_builder_desc(rl_scene_builder *b, rl_scene_desc *out);

#ifdef __cplusplus
}
#endif
#endif /* RL_HOST_H */
