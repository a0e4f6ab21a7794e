/*
 * quasi_oracle_mesh.h - CPU oracle for the triangle-mesh scene (BASELINE config 3).
 *
 * TEST INFRASTRUCTURE ONLY (see quasi_oracle.h).
 *
 * The reference has no triangles, no mesh and no BVH:
 * scene intersection there is a linear scan over <= 32 quads
 * (/root/reference/backends/metal/plugin.mm:199-212). BASELINE.json's third
 * configuration asks for a 1 M-triangle procedural scene behind a SAH BVH, so
 * this file DEFINES that scene and its arithmetic; everything around the
 * intersection (seed, camera ray, path loop, sampling, accumulation) is the
 * reference shader's, shared with quasi_oracle.c. "Parity unpinned" by the
 * reference for this configuration: there is nothing in it to pin against. The
 * pins are internal: BVH traversal must equal the brute-force scan bit for bit,
 * and the CUDA backend must equal both.
 *
 * Scene ("terrain box"): the six quads of the Cornell room as two triangles each
 * (12 triangles), then a height-field terrain of grid x grid cells, two triangles
 * per cell, wall to wall. Heights are three octaves of value noise from the
 * shader's own pcg_hash, smoothstep-interpolated: only + - * / on binary32, so
 * the mesh is the same bits everywhere.
 *
 * Intersection: Moeller-Trumbore in the operation order of qm_intersect below;
 * closest hit, t in [0.001, closest]; among equal t the higher triangle index
 * wins (the rule the shader's loop implies, plugin.mm:177,205-207), so the
 * result does not depend on visiting order.
 *
 * BVH: binary, built top-down with 16-bin SAH on centroids; children are
 * allocated in pairs when their parent is split, left subtree built first;
 * leaves hold <= 2 triangles. One rule precedes the SAH, at the root only:
 * triangles whose box area exceeds 1/64 of the scene box's (the room's walls)
 * are split off, in order, into the right subtree, which is then built the same
 * way with leaves of one triangle. Node boxes are padded by 1e-4*extent + 1e-5 so
 * that the slab test can only ever prune a node the ray provably misses. The node
 * array (32-byte nodes) and the leaf-order triangle permutation are the
 * bit-exact contract with the backend's own builder.
 */
#ifndef QUASI_ORACLE_MESH_H
#define QUASI_ORACLE_MESH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    float v0[3];
    uint32_t material;
    float e1[3];
    uint32_t pad0;
    float e2[3];
    uint32_t pad1;
} qm_triangle; /* 48 bytes */

typedef struct {
    float bmin[3];
    uint32_t left_first; /* interior: index of the left child (right = +1); leaf: first slot in leaf order */
    float bmax[3];
    uint32_t count;      /* 0 = interior, else number of triangles in the leaf */
} qm_node; /* 32 bytes */

typedef struct {
    float albedo[3];
    float emission[3];
} qm_material;

typedef struct qm_scene qm_scene;

/* Builds mesh + BVH. grid = terrain cells per side (707 gives 999 710 triangles). */
qm_scene* qm_scene_create(uint32_t grid, uint32_t seed);
void qm_scene_destroy(qm_scene* scene);

uint32_t qm_triangle_count(const qm_scene* scene);
uint32_t qm_node_count(const qm_scene* scene);
const qm_triangle* qm_triangles(const qm_scene* scene);   /* original order */
const qm_node* qm_nodes(const qm_scene* scene);
const uint32_t* qm_leaf_order(const qm_scene* scene);     /* leaf slot -> original triangle index */

typedef struct {
    uint64_t rays;
    uint64_t nodes_visited;    /* box tests */
    uint64_t triangles_tested;
} qm_stats;

/* n_frames render calls from a reset (same contract as qo_render_frames).
 * use_bvh = 0 scans every triangle for every ray (the oracle's oracle). */
void qm_render_frames(const qm_scene* scene, const float position[3], const float target[3],
                      const float up[3], float fov, uint32_t width, uint32_t height,
                      uint32_t frame_first, uint32_t frame_stride, uint32_t n_frames,
                      uint32_t max_bounces, int use_bvh, float* const layers[4], int threads,
                      qm_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
