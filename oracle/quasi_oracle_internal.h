/*
 * quasi_oracle_internal.h - vector helpers and the two shader functions shared
 * by the quad oracle (quasi_oracle.c) and the mesh oracle (quasi_oracle_mesh.c).
 * TEST INFRASTRUCTURE ONLY (see quasi_oracle.h).
 */
#ifndef QUASI_ORACLE_INTERNAL_H
#define QUASI_ORACLE_INTERNAL_H

#include "quasi_oracle.h"

#include <math.h>

typedef struct {
    float x, y, z;
} v3;

/* vec.hpp:40-50 (component-wise operators) */
static inline v3 v3_make(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_mul(v3 a, v3 b) { return v3_make(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3_scale(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_div(v3 a, float s) { return v3_make(a.x / s, a.y / s, a.z / s); }
static inline v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
static inline v3 v3_adds(v3 a, float s) { return v3_make(a.x + s, a.y + s, a.z + s); }
/* vec.hpp:92-94 */
static inline float v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* vec.hpp:100-106 */
static inline v3 v3_cross(v3 a, v3 b) {
    return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* vec.hpp:108-114: v / length(v) */
static inline v3 v3_normalize(v3 a) { return v3_div(a, sqrtf(v3_dot(a, a))); }
static inline v3 v3_load(const float* p) { return v3_make(p[0], p[1], p[2]); }


typedef struct {
    v3 origin;
    v3 direction;
} ray_t;

/* plugin.mm:111-125 */
v3 qo_cosine_sample_hemisphere(v3 normal, uint32_t* rng);
/* plugin.mm:129-153 */
ray_t qo_get_camera_ray(const qo_camera* cam, float uv_x, float uv_y, uint32_t* rng);

#endif
