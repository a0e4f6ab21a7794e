/*
 * quasi_oracle.c - see quasi_oracle.h. TEST INFRASTRUCTURE ONLY.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off -fno-fast-math -fPIC -shared
 *        (oracle/Makefile). x86-64 SSE2 arithmetic: every float expression
 *        below is evaluated in binary32 with no excess precision.
 */
#define _GNU_SOURCE
#include "quasi_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "quasi_oracle_internal.h"

static inline float q_max(float a, float b) { return a > b ? a : b; }

/* Correctly rounded binary32 transcendental results (header comment). */
float qo_sin(float x) { return (float)sin((double)x); }
float qo_cos(float x) { return (float)cos((double)x); }
float qo_tan(float x) { return (float)tan((double)x); }
void qo_sincos_array(const float* phi, uint64_t n, float* sin_out, float* cos_out) {
    for (uint64_t i = 0; i < n; i++) {
        sin_out[i] = qo_sin(phi[i]);
        cos_out[i] = qo_cos(phi[i]);
    }
}
static inline float q_pow(float x, float y) { return (float)pow((double)x, (double)y); }

#define Q_PI_F 3.14159265358979323846264338327950288f /* MSL M_PI_F */

/* plugin.mm:98-102 */
uint32_t qo_pcg_hash(uint32_t input) {
    uint32_t state = input * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}

/* plugin.mm:104-107. float(0xFFFFFFFFu) rounds to 2^32. */
float qo_random_float(uint32_t* state) {
    *state = qo_pcg_hash(*state);
    return (float)(*state) / (float)0xFFFFFFFFu;
}

/* plugin.mm:295 - the 1920x1080 stride is hard-coded in the reference. */
uint32_t qo_seed(uint32_t px, uint32_t py, uint32_t frame_count) {
    return qo_pcg_hash(px + py * 1920u + frame_count * 1920u * 1080u);
}

typedef struct {
    float t;
    v3 point;
    v3 normal;
    uint32_t material_index;
    int hit;
} hit_t;

/* plugin.mm:111-125 */
v3 qo_cosine_sample_hemisphere(v3 normal, uint32_t* rng) {
    float r1 = qo_random_float(rng);
    float r2 = qo_random_float(rng);

    float phi = 2.0f * Q_PI_F * r1;
    float cos_theta = sqrtf(1.0f - r2);
    float sin_theta = sqrtf(r2);

    v3 w = v3_normalize(normal);
    v3 a = (fabsf(w.x) > 0.9f) ? v3_make(0, 1, 0) : v3_make(1, 0, 0);
    v3 v = v3_normalize(v3_cross(w, a));
    v3 u = v3_cross(w, v);

    v3 du = v3_scale(v3_scale(u, qo_cos(phi)), sin_theta);
    v3 dv = v3_scale(v3_scale(v, qo_sin(phi)), sin_theta);
    v3 dw = v3_scale(w, cos_theta);
    return v3_normalize(v3_add(v3_add(du, dv), dw));
}

/* plugin.mm:129-153. Jitter draws x first, then y. */
ray_t qo_get_camera_ray(const qo_camera* cam, float uv_x, float uv_y, uint32_t* rng) {
    float jx = qo_random_float(rng);
    float jy = qo_random_float(rng);
    jx = jx - 0.5f;
    jy = jy - 0.5f;
    uv_x += jx * 0.001f;
    uv_y += jy * 0.001f;

    float theta = cam->fov * Q_PI_F / 180.0f;
    float h = qo_tan(theta / 2.0f);
    float viewport_height = 2.0f * h;
    float viewport_width = cam->aspect * viewport_height;

    v3 position = v3_load(cam->position);
    v3 w = v3_normalize(v3_neg(v3_load(cam->direction)));
    v3 right = v3_normalize(v3_cross(v3_load(cam->up), w));
    v3 up = v3_cross(w, right);

    v3 horizontal = v3_scale(right, viewport_width);
    v3 vertical = v3_scale(up, viewport_height);
    v3 lower_left =
        v3_sub(v3_sub(v3_sub(position, v3_scale(horizontal, 0.5f)), v3_scale(vertical, 0.5f)), w);

    v3 target = v3_add(v3_add(lower_left, v3_scale(horizontal, uv_x)), v3_scale(vertical, uv_y));

    ray_t ray;
    ray.origin = position;
    ray.direction = v3_normalize(v3_sub(target, ray.origin));
    return ray;
}

/* plugin.mm:157-195 */
static hit_t intersect_quad(ray_t ray, const qo_quad* q, uint32_t mat_idx, float t_min,
                            float t_max) {
    hit_t rec;
    memset(&rec, 0, sizeof rec);

    v3 origin = v3_load(q->origin);
    v3 u = v3_load(q->u);
    v3 v = v3_load(q->v);

    v3 n = v3_cross(u, v);
    float area_sq = v3_dot(n, n);
    if (area_sq < 1e-8f) return rec;

    v3 normal = v3_scale(n, 1.0f / sqrtf(area_sq)); /* n * rsqrt(area_sq) */
    float d = v3_dot(normal, origin);

    float denom = v3_dot(normal, ray.direction);
    if (fabsf(denom) < 1e-8f) return rec;

    float t = (d - v3_dot(normal, ray.origin)) / denom;
    if (t < t_min || t > t_max) return rec;

    v3 p = v3_add(ray.origin, v3_scale(ray.direction, t));
    v3 planar = v3_sub(p, origin);

    v3 w = v3_div(n, area_sq);
    float alpha = v3_dot(w, v3_cross(planar, v));
    float beta = v3_dot(w, v3_cross(u, planar));
    if (alpha < 0.0f || alpha > 1.0f || beta < 0.0f || beta > 1.0f) return rec;

    rec.hit = 1;
    rec.t = t;
    rec.point = p;
    rec.material_index = mat_idx;
    rec.normal = (denom < 0.0f) ? normal : v3_neg(normal);
    return rec;
}

/* plugin.mm:199-212. t == closest.t is accepted, so the later index wins. */
static hit_t trace_scene(ray_t ray, const qo_scene_uniforms* scene) {
    hit_t closest;
    memset(&closest, 0, sizeof closest);
    closest.t = 1e30f;
    for (uint32_t i = 0; i < scene->quad_count && i < QO_MAX_QUADS; i++) {
        hit_t hit = intersect_quad(ray, &scene->quads[i], i, 0.001f, closest.t);
        if (hit.hit) closest = hit;
    }
    return closest;
}

typedef struct {
    v3 color;
    v3 first_hit_albedo;
    v3 first_hit_normal;
    float first_hit_depth;
    uint32_t rays;
} path_result_t;

/* plugin.mm:223-276 */
static path_result_t path_trace(ray_t ray, const qo_scene_uniforms* scene, uint32_t* rng) {
    path_result_t result;
    memset(&result, 0, sizeof result);
    v3 throughput = v3_make(1.0f, 1.0f, 1.0f);

    for (int bounce = 0; bounce < QO_MAX_BOUNCES; bounce++) {
        hit_t hit = trace_scene(ray, scene);
        result.rays++;
        if (!hit.hit) break;

        const qo_material* mat = &scene->materials[hit.material_index];
        v3 emission = v3_load(mat->emission);
        v3 albedo = v3_load(mat->albedo);

        if (bounce == 0) {
            float emit_str = q_max(emission.x, q_max(emission.y, emission.z));
            result.first_hit_albedo = (emit_str > 0.1f) ? emission : albedo;
            result.first_hit_normal = v3_adds(v3_scale(hit.normal, 0.5f), 0.5f);
            result.first_hit_depth = hit.t;
        }

        result.color = v3_add(result.color, v3_mul(throughput, emission));

        float emit_strength = q_max(emission.x, q_max(emission.y, emission.z));
        if (emit_strength > 0.1f) break;

        throughput = v3_mul(throughput, albedo);

        if (bounce > 2) {
            float p = q_max(0.05f, q_max(throughput.x, q_max(throughput.y, throughput.z)));
            if (qo_random_float(rng) > p) break;
            throughput = v3_div(throughput, p);
        }

        ray.origin = v3_add(hit.point, v3_scale(hit.normal, 0.001f));
        ray.direction = qo_cosine_sample_hemisphere(hit.normal, rng);
    }
    return result;
}

/* plugin.mm:287-309; uv closed form: SURVEY.md section 8a row A1. */
uint32_t qo_pathtrace_pixel(const qo_scene_uniforms* scene, uint32_t px, uint32_t py,
                            uint32_t width, uint32_t height, qo_fragment* out) {
    float uv_x = ((float)px + 0.5f) / (float)width;
    float uv_y = 1.0f - ((float)py + 0.5f) / (float)height;

    uint32_t rng = qo_seed(px, py, scene->frame_count);

    ray_t ray = qo_get_camera_ray(&scene->camera, uv_x, uv_y, &rng);
    path_result_t pt = path_trace(ray, scene, &rng);

    float norm_depth = (pt.first_hit_depth - 1.0f) / 4.0f;
    norm_depth = fminf(fmaxf(norm_depth, 0.0f), 1.0f);

    out->beauty[0] = pt.color.x;
    out->beauty[1] = pt.color.y;
    out->beauty[2] = pt.color.z;
    out->beauty[3] = 1.0f;
    out->albedo[0] = pt.first_hit_albedo.x;
    out->albedo[1] = pt.first_hit_albedo.y;
    out->albedo[2] = pt.first_hit_albedo.z;
    out->albedo[3] = 1.0f;
    out->normal[0] = pt.first_hit_normal.x;
    out->normal[1] = pt.first_hit_normal.y;
    out->normal[2] = pt.first_hit_normal.z;
    out->normal[3] = 1.0f;
    out->depth[0] = norm_depth;
    out->depth[1] = norm_depth;
    out->depth[2] = norm_depth;
    out->depth[3] = 1.0f;
    return pt.rays;
}

int qo_online_cores(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) {
        int c = CPU_COUNT(&set);
        if (c > 0 && c < n) n = c;
    }
    return n > 0 ? (int)n : 1;
}

typedef struct {
    const qo_scene_uniforms* scene;
    uint32_t width, height;
    float* layer[4];
    int tid, nthreads;
    uint64_t rays;
} sample_job_t;

/* Rows are interleaved over threads; pixels are independent, so the thread
 * count cannot change a single bit of the result. */
static void* sample_worker(void* arg) {
    sample_job_t* job = (sample_job_t*)arg;
    uint64_t rays = 0;
    for (uint32_t py = (uint32_t)job->tid; py < job->height; py += (uint32_t)job->nthreads) {
        for (uint32_t px = 0; px < job->width; px++) {
            qo_fragment f;
            rays += qo_pathtrace_pixel(job->scene, px, py, job->width, job->height, &f);
            size_t o = ((size_t)py * job->width + px) * 4;
            memcpy(job->layer[0] + o, f.beauty, 16);
            memcpy(job->layer[1] + o, f.albedo, 16);
            memcpy(job->layer[2] + o, f.normal, 16);
            memcpy(job->layer[3] + o, f.depth, 16);
        }
    }
    job->rays = rays;
    return NULL;
}

uint64_t qo_render_sample(const qo_scene_uniforms* scene, uint32_t width, uint32_t height,
                          float* beauty, float* albedo, float* normal, float* depth,
                          int threads) {
    if (threads <= 0) threads = qo_online_cores();
    if ((uint32_t)threads > height) threads = (int)height;
    if (threads < 1) threads = 1;

    pthread_t* tids = (pthread_t*)calloc((size_t)threads, sizeof *tids);
    sample_job_t* jobs = (sample_job_t*)calloc((size_t)threads, sizeof *jobs);
    for (int i = 0; i < threads; i++) {
        jobs[i].scene = scene;
        jobs[i].width = width;
        jobs[i].height = height;
        jobs[i].layer[0] = beauty;
        jobs[i].layer[1] = albedo;
        jobs[i].layer[2] = normal;
        jobs[i].layer[3] = depth;
        jobs[i].tid = i;
        jobs[i].nthreads = threads;
        if (i > 0) pthread_create(&tids[i], NULL, sample_worker, &jobs[i]);
    }
    sample_worker(&jobs[0]);
    uint64_t rays = jobs[0].rays;
    for (int i = 1; i < threads; i++) {
        pthread_join(tids[i], NULL);
        rays += jobs[i].rays;
    }
    free(tids);
    free(jobs);
    return rays;
}

/* plugin.mm:313-330: mix(accumulated, new_sample, 1/(frame_count+1)). */
void qo_accumulate(const float* current, const float* history, float* out, uint64_t n_px,
                   uint32_t frame_count) {
    float weight = 1.0f / (float)(frame_count + 1u);
    for (uint64_t i = 0; i < n_px * 4; i++) {
        float accumulated = history[i];
        out[i] = accumulated + (current[i] - accumulated) * weight;
    }
}

uint64_t qo_render_frames(const qo_scene_uniforms* scene, uint32_t width, uint32_t height,
                          uint32_t frame_first, uint32_t frame_stride, uint32_t n_frames,
                          float* const layers[4], int threads) {
    uint64_t n_px = (uint64_t)width * height;
    float* cur[4];
    for (int l = 0; l < 4; l++) {
        cur[l] = (float*)malloc(n_px * 16);
        memset(layers[l], 0, n_px * 16); /* history before call 0 := 0 (H9) */
    }
    qo_scene_uniforms s = *scene;
    uint64_t rays = 0;
    for (uint32_t i = 0; i < n_frames; i++) {
        s.frame_count = frame_first + i * frame_stride;
        rays += qo_render_sample(&s, width, height, cur[0], cur[1], cur[2], cur[3], threads);
        for (int l = 0; l < 4; l++) qo_accumulate(cur[l], layers[l], layers[l], n_px, i);
    }
    for (int l = 0; l < 4; l++) free(cur[l]);
    return rays;
}

static inline uint8_t unorm8(float c) {
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    return (uint8_t)rintf(c * 255.0f);
}

/* plugin.mm:334-350. Output texel (px,py) reads accumulated texel (px,py). */
void qo_tonemap_bgra8(const float* accumulated, uint32_t width, uint32_t height,
                      uint8_t* bgra) {
    const float inv_gamma = 1.0f / 2.2f;
    uint64_t n_px = (uint64_t)width * height;
    for (uint64_t i = 0; i < n_px; i++) {
        float c[3];
        for (int k = 0; k < 3; k++) {
            float v = accumulated[i * 4 + k];
            v = v / (v + 1.0f);
            c[k] = q_pow(v, inv_gamma);
        }
        bgra[i * 4 + 0] = unorm8(c[2]);
        bgra[i * 4 + 1] = unorm8(c[1]);
        bgra[i * 4 + 2] = unorm8(c[0]);
        bgra[i * 4 + 3] = 255;
    }
}

void qo_combine_shards(const float* const* shards, uint32_t n_shards, uint64_t n_floats,
                       float* out) {
    float scale = 1.0f / (float)n_shards;
    for (uint64_t i = 0; i < n_floats; i++) {
        float s = shards[0][i];
        for (uint32_t g = 1; g < n_shards; g++) s = s + shards[g][i];
        out[i] = s * scale;
    }
}

/* ---- demo scene, cornell_box.hpp ---- */

/* cornell_box.hpp:32-37 (std::cos/std::sin on float = cosf/sinf). */
static v3 rotate_y(v3 p, float angle_deg) {
    float rad = angle_deg * 3.14159265359f / 180.0f;
    float c = cosf(rad);
    float s = sinf(rad);
    return v3_make(p.x * c + p.z * s, p.y, -p.x * s + p.z * c);
}

static void push_quad(qo_scene_uniforms* scene, v3 origin, v3 u, v3 v, const qo_material* m) {
    uint32_t i = scene->quad_count++;
    qo_quad* q = &scene->quads[i];
    memset(q, 0, sizeof *q);
    q->origin[0] = origin.x; q->origin[1] = origin.y; q->origin[2] = origin.z;
    q->u[0] = u.x; q->u[1] = u.y; q->u[2] = u.z;
    q->v[0] = v.x; q->v[1] = v.y; q->v[2] = v.z;
    scene->materials[i] = *m;
}

/* cornell_box.hpp:45-90 */
static void add_box(qo_scene_uniforms* scene, v3 center, v3 size, float angle_y,
                    const qo_material* m) {
    float hw = size.x * 0.5f;
    float h = size.y;
    float hd = size.z * 0.5f;
    v3 local[8] = {
        {-hw, 0, -hd}, {hw, 0, -hd}, {-hw, 0, hd}, {hw, 0, hd},
        {-hw, h, -hd}, {hw, h, -hd}, {-hw, h, hd}, {hw, h, hd},
    };
    v3 p[8];
    for (int i = 0; i < 8; i++) p[i] = v3_add(rotate_y(local[i], angle_y), center);

    push_quad(scene, p[4], v3_sub(p[5], p[4]), v3_sub(p[6], p[4]), m); /* top */
    push_quad(scene, p[0], v3_sub(p[1], p[0]), v3_sub(p[4], p[0]), m); /* front */
    push_quad(scene, p[3], v3_sub(p[2], p[3]), v3_sub(p[7], p[3]), m); /* back */
    push_quad(scene, p[2], v3_sub(p[0], p[2]), v3_sub(p[6], p[2]), m); /* left */
    push_quad(scene, p[1], v3_sub(p[3], p[1]), v3_sub(p[5], p[1]), m); /* right */
}

/* cornell_box.hpp:95-134 */
void qo_make_cornell_box(qo_scene_uniforms* scene) {
    qo_material white = {{0.73f, 0.73f, 0.73f}, 1.0f, {0, 0, 0}, 0.0f};
    qo_material red = {{0.65f, 0.05f, 0.05f}, 1.0f, {0, 0, 0}, 0.0f};
    qo_material green = {{0.12f, 0.45f, 0.15f}, 1.0f, {0, 0, 0}, 0.0f};
    qo_material light = {{0, 0, 0}, 1.0f, {15.0f, 15.0f, 15.0f}, 0.0f};

    scene->quad_count = 0;
    memset(scene->quads, 0, sizeof scene->quads);
    memset(scene->materials, 0, sizeof scene->materials);

    push_quad(scene, v3_make(-1, 0, -1), v3_make(2, 0, 0), v3_make(0, 0, 2), &white);  /* floor */
    push_quad(scene, v3_make(-1, 2, 1), v3_make(2, 0, 0), v3_make(0, 0, -2), &white);  /* ceiling */
    push_quad(scene, v3_make(-1, 0, -1), v3_make(2, 0, 0), v3_make(0, 2, 0), &white);  /* back */
    push_quad(scene, v3_make(-1, 0, 1), v3_make(0, 0, -2), v3_make(0, 2, 0), &red);    /* left */
    push_quad(scene, v3_make(1, 0, -1), v3_make(0, 0, 2), v3_make(0, 2, 0), &green);   /* right */

    scene->light_index = scene->quad_count;
    push_quad(scene, v3_make(-0.25f, 1.99f, -0.25f), v3_make(0.5f, 0, 0), v3_make(0, 0, 0.5f),
              &light);

    add_box(scene, v3_make(-0.35f, 0.0f, 0.3f), v3_make(0.5f, 1.2f, 0.5f), 15.0f, &white);
    add_box(scene, v3_make(0.35f, 0.0f, -0.3f), v3_make(0.55f, 0.55f, 0.55f), -18.0f, &white);
}

/* plugin.mm:648-674 */
void qo_pack_camera(qo_scene_uniforms* scene, const float position[3], const float target[3],
                    const float up[3], float fov, uint32_t width, uint32_t height) {
    float aspect = (float)width / (float)height;
    float dx = target[0] - position[0];
    float dy = target[1] - position[1];
    float dz = target[2] - position[2];
    float len = sqrtf(dx * dx + dy * dy + dz * dz);
    if (len > 0.0f) {
        dx /= len;
        dy /= len;
        dz /= len;
    }
    qo_camera* c = &scene->camera;
    memset(c, 0, sizeof *c);
    c->position[0] = position[0]; c->position[1] = position[1]; c->position[2] = position[2];
    c->direction[0] = dx; c->direction[1] = dy; c->direction[2] = dz;
    c->up[0] = up[0]; c->up[1] = up[1]; c->up[2] = up[2];
    c->fov = fov;
    c->aspect = aspect;
}
