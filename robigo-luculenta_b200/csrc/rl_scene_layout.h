// rl_scene_layout.h -- the flattened scene as the kernels receive it (plain
// host/device structs; the device functions over it are in rl_device.cuh).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define RL_MAX_COMPOUND_STACK 5   // deepest evaluation stack of a compound-surface program
// the kernels pack (lane, table index) pairs into 16 bits: 5 bits of lane, 11 bits of cluster or
// compound index; rl_scene_create rejects scenes with more clusters or compounds than that
#define RL_PAIR_INDEX_BITS 11
#define RL_PAIR_INDEX_MAX ((1u << RL_PAIR_INDEX_BITS) - 1u)

namespace rl {

struct DevCamera {
    uint32_t kind;
    float px, py, pz;
    float field_of_view, focal_distance, depth_of_field, chromatic_abberation;
    float screen_distance;    // 1 / tan(field_of_view / 2) (camera.rs:60), specified arithmetic
    float qx, qy, qz, qw;
    float phi_base, phi_rate, alpha_base, alpha_rate, distance_base, distance_rate, focal_factor;
    // KEYFRAMES: three float4 per frame {position, focal_distance} {orientation} {depth_of_field,
    // chromatic_abberation, screen_distance, 0}, in global memory
    const float4 *keyframes;
    uint32_t n_keyframes;
};

struct DevScene {
    const float4 *blob;       // global copy of the primitive blob
    uint32_t blob_vec4;       // its size in float4 units
    uint32_t smem_vec4;       // leading part that the kernels copy into shared memory
    // offsets into the blob, in float4 units
    uint32_t off_spheres, n_spheres;
    uint32_t off_planes, n_planes;
    uint32_t off_paraboloids, n_paraboloids;
    uint32_t off_leaves, n_leaves;
    uint32_t off_compounds, n_compounds;
    uint32_t off_ops, n_ops;
    uint32_t off_sphere_obj, off_plane_obj, off_paraboloid_obj, off_compound_obj, off_sphere_k;
    uint32_t off_clusters, n_clusters, off_cluster_range;   // clusters padded to a multiple of 8 records
    // scenes with a thousand spheres or more: a level of bounds above the clusters -- super k bounds
    // the members of clusters [8k, 8k + 8) (padded to a multiple of 8 records; 0: no such level)
    uint32_t off_supers, n_supers;
    uint32_t off_body_bounds, off_body_always;              // bounding spheres of the compounds in the pre-test's form
                                                            // (padded to 8), and per 64 bodies the mask of unbounded ones
    const float4 *materials;  // per object
    uint32_t n_objects;
    float sphere_cmax2;       // max (|centre|^2 + r^2) over spheres and clusters (error bound of the pre-test)
    float cluster_rmax;       // largest cluster bounding radius
    float super_rmax;         // largest bounding radius of a group of eight clusters
    float leaf_off_max;       // largest |offset| over the half-spaces of compound surfaces (slab-test inflation)
    float body_rmax;          // largest bounding radius of a compound
    uint32_t sphere_leaves;   // 1: some compound has a sphere leaf (geometry.rs:263-267)
    uint32_t sphere_k_global; // 1: the spheres' pre-test records (off_sphere_k) are not in the shared-memory part
    DevCamera camera;
};

}  // namespace rl
