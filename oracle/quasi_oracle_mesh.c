/*
 * quasi_oracle_mesh.c - see quasi_oracle_mesh.h. TEST INFRASTRUCTURE ONLY.
 */
#define _GNU_SOURCE
#include "quasi_oracle_mesh.h"

#include "quasi_oracle_internal.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

struct qm_scene {
    qm_triangle* tris;
    uint32_t n_tris;
    qm_node* nodes;
    uint32_t n_nodes;
    uint32_t* leaf_order;
    qm_material materials[8];
};

/* ------------------------------------------------------------------ mesh -- */

static float lattice(uint32_t a, uint32_t b, uint32_t octave, uint32_t seed) {
    uint32_t h = qo_pcg_hash(a + b * 64u + octave * 4096u + seed * 65536u);
    return (float)h * 2.3283064365386963e-10f; /* * 2^-32, as random_float */
}

static float terrain_height(uint32_t i, uint32_t j, uint32_t grid, uint32_t seed) {
    static const float amplitude[3] = {0.35f, 0.15f, 0.07f};
    float u = (float)i / (float)grid;
    float v = (float)j / (float)grid;
    float h = 0.03f;
    for (uint32_t o = 0; o < 3; o++) {
        uint32_t freq = 4u << o;
        float fx = u * (float)freq;
        float fz = v * (float)freq;
        uint32_t ix = (uint32_t)fx;
        uint32_t iz = (uint32_t)fz;
        if (ix >= freq) ix = freq - 1u;
        if (iz >= freq) iz = freq - 1u;
        float tx = fx - (float)ix;
        float tz = fz - (float)iz;
        float sx = tx * tx * (3.0f - 2.0f * tx);
        float sz = tz * tz * (3.0f - 2.0f * tz);
        float l00 = lattice(ix, iz, o, seed), l10 = lattice(ix + 1u, iz, o, seed);
        float l01 = lattice(ix, iz + 1u, o, seed), l11 = lattice(ix + 1u, iz + 1u, o, seed);
        float n0 = l00 + (l10 - l00) * sx;
        float n1 = l01 + (l11 - l01) * sx;
        float n = n0 + (n1 - n0) * sz;
        h = h + amplitude[o] * n;
    }
    return h;
}

static v3 terrain_vertex(uint32_t i, uint32_t j, uint32_t grid, uint32_t seed) {
    float x = -1.0f + 2.0f * ((float)i / (float)grid);
    float z = -1.0f + 2.0f * ((float)j / (float)grid);
    return v3_make(x, terrain_height(i, j, grid, seed), z);
}

static void set_triangle(qm_triangle* t, v3 v0, v3 e1, v3 e2, uint32_t material) {
    memset(t, 0, sizeof *t);
    t->v0[0] = v0.x; t->v0[1] = v0.y; t->v0[2] = v0.z;
    t->e1[0] = e1.x; t->e1[1] = e1.y; t->e1[2] = e1.z;
    t->e2[0] = e2.x; t->e2[1] = e2.y; t->e2[2] = e2.z;
    t->material = material;
}

static void build_mesh(qm_scene* s, uint32_t grid, uint32_t seed) {
    /* cornell_box.hpp:103-121: floor, ceiling, back, left (red), right (green), light */
    static const float room[6][10] = {
        {-1, 0, -1, 2, 0, 0, 0, 0, 2, 0},        {-1, 2, 1, 2, 0, 0, 0, 0, -2, 0},
        {-1, 0, -1, 2, 0, 0, 0, 2, 0, 0},        {-1, 0, 1, 0, 0, -2, 0, 2, 0, 1},
        {1, 0, -1, 0, 0, 2, 0, 2, 0, 2},         {-0.25f, 1.99f, -0.25f, 0.5f, 0, 0, 0, 0, 0.5f, 3},
    };
    const qm_material mats[5] = {
        {{0.73f, 0.73f, 0.73f}, {0, 0, 0}}, {{0.65f, 0.05f, 0.05f}, {0, 0, 0}},
        {{0.12f, 0.45f, 0.15f}, {0, 0, 0}}, {{0, 0, 0}, {15.0f, 15.0f, 15.0f}},
        {{0.6f, 0.55f, 0.4f}, {0, 0, 0}},
    };
    memcpy(s->materials, mats, sizeof mats);

    s->n_tris = 12u + 2u * grid * grid;
    s->tris = (qm_triangle*)malloc((size_t)s->n_tris * sizeof(qm_triangle));
    uint32_t k = 0;
    for (int q = 0; q < 6; q++) {
        v3 o = v3_make(room[q][0], room[q][1], room[q][2]);
        v3 u = v3_make(room[q][3], room[q][4], room[q][5]);
        v3 v = v3_make(room[q][6], room[q][7], room[q][8]);
        uint32_t m = (uint32_t)room[q][9];
        set_triangle(&s->tris[k++], o, u, v3_add(u, v), m);
        set_triangle(&s->tris[k++], o, v3_add(u, v), v, m);
    }
    for (uint32_t j = 0; j < grid; j++) {
        for (uint32_t i = 0; i < grid; i++) {
            v3 p00 = terrain_vertex(i, j, grid, seed), p10 = terrain_vertex(i + 1u, j, grid, seed);
            v3 p01 = terrain_vertex(i, j + 1u, grid, seed), p11 = terrain_vertex(i + 1u, j + 1u, grid, seed);
            set_triangle(&s->tris[k++], p00, v3_sub(p01, p00), v3_sub(p11, p00), 4u);
            set_triangle(&s->tris[k++], p00, v3_sub(p11, p00), v3_sub(p10, p00), 4u);
        }
    }
}

/* ------------------------------------------------------------------- BVH -- */

typedef struct {
    float lo[3], hi[3];
} box_t;

static void box_empty(box_t* b) {
    for (int a = 0; a < 3; a++) {
        b->lo[a] = 1e30f;
        b->hi[a] = -1e30f;
    }
}
static void box_grow_point(box_t* b, const float p[3]) {
    for (int a = 0; a < 3; a++) {
        if (p[a] < b->lo[a]) b->lo[a] = p[a];
        if (p[a] > b->hi[a]) b->hi[a] = p[a];
    }
}
static void box_grow_box(box_t* b, const box_t* o) {
    box_grow_point(b, o->lo);
    box_grow_point(b, o->hi);
}
static float box_area(const box_t* b) {
    float dx = b->hi[0] - b->lo[0], dy = b->hi[1] - b->lo[1], dz = b->hi[2] - b->lo[2];
    return 2.0f * ((dx * dy + dy * dz) + dz * dx);
}

typedef struct {
    const qm_triangle* tris;
    box_t* prim_box;      /* per original triangle */
    float (*centroid)[3]; /* per original triangle */
    uint32_t* order;      /* leaf order under construction */
    uint32_t* scratch;
    qm_node* nodes;
    uint32_t n_nodes;
    uint32_t outlier_first; /* leaf slots from here on hold the outliers split off at the root */
} builder_t;

#define QM_BINS 16
#define QM_LEAF_MAX 2
#define QM_OUTLIER_FRACTION (1.0f / 64.0f)

static void build_node(builder_t* b, uint32_t node_index, uint32_t first, uint32_t count) {
    box_t bounds, cbounds;
    box_empty(&bounds);
    box_empty(&cbounds);
    for (uint32_t k = first; k < first + count; k++) {
        uint32_t t = b->order[k];
        box_grow_box(&bounds, &b->prim_box[t]);
        box_grow_point(&cbounds, b->centroid[t]);
    }
    qm_node* node = &b->nodes[node_index];
    for (int a = 0; a < 3; a++) {
        float pad = (bounds.hi[a] - bounds.lo[a]) * 1e-4f + 1e-5f;
        node->bmin[a] = bounds.lo[a] - pad;
        node->bmax[a] = bounds.hi[a] + pad;
    }
    if (count <= (first >= b->outlier_first ? 1u : QM_LEAF_MAX)) {
        node->left_first = first;
        node->count = count;
        return;
    }

    /* Root only: triangles whose box is larger than 1/64 of the scene's box (walls, floor)
     * go to the right subtree as they are, in leaves of one; everything else goes left. A
     * greedy SAH never isolates them by itself and each would inflate one box per level. */
    if (node_index == 0) {
        float limit = box_area(&bounds) * QM_OUTLIER_FRACTION;
        uint32_t nl = 0, nr = 0;
        for (uint32_t k = first; k < first + count; k++) { /* stable partition */
            uint32_t t = b->order[k];
            if (box_area(&b->prim_box[t]) > limit) {
                b->scratch[nr++] = t;
            } else {
                b->order[first + nl++] = t;
            }
        }
        memcpy(&b->order[first + nl], b->scratch, (size_t)nr * sizeof(uint32_t));
        if (nl > 0 && nr > 0) {
            b->outlier_first = first + nl;
            uint32_t left = b->n_nodes;
            b->n_nodes += 2;
            node->left_first = left;
            node->count = 0;
            build_node(b, left, first, nl);
            build_node(b, left + 1u, first + nl, nr);
            return;
        }
    }

    /* 16-bin SAH over the centroid bounds; first strictly smaller cost wins. */
    int best_axis = -1;
    uint32_t best_split = 0;
    float best_cost = 1e30f;
    for (int axis = 0; axis < 3; axis++) {
        float lo = cbounds.lo[axis];
        float extent = cbounds.hi[axis] - lo;
        if (!(extent > 0.0f)) continue;
        box_t bin_box[QM_BINS];
        uint32_t bin_count[QM_BINS];
        for (int i = 0; i < QM_BINS; i++) {
            box_empty(&bin_box[i]);
            bin_count[i] = 0;
        }
        for (uint32_t k = first; k < first + count; k++) {
            uint32_t t = b->order[k];
            int bin = (int)(((b->centroid[t][axis] - lo) / extent) * (float)QM_BINS);
            if (bin > QM_BINS - 1) bin = QM_BINS - 1;
            bin_count[bin]++;
            box_grow_box(&bin_box[bin], &b->prim_box[t]);
        }
        box_t right_box[QM_BINS];
        uint32_t right_count[QM_BINS];
        box_t acc;
        box_empty(&acc);
        uint32_t n = 0;
        for (int i = QM_BINS - 1; i >= 1; i--) {
            box_grow_box(&acc, &bin_box[i]);
            n += bin_count[i];
            right_box[i] = acc;
            right_count[i] = n;
        }
        box_empty(&acc);
        n = 0;
        for (int split = 1; split < QM_BINS; split++) {
            box_grow_box(&acc, &bin_box[split - 1]);
            n += bin_count[split - 1];
            if (n == 0 || right_count[split] == 0) continue;
            float cost = box_area(&acc) * (float)n + box_area(&right_box[split]) * (float)right_count[split];
            if (cost < best_cost) {
                best_cost = cost;
                best_axis = axis;
                best_split = (uint32_t)split;
            }
        }
    }

    uint32_t left_count;
    if (best_axis < 0) {
        left_count = count / 2u; /* all centroids coincide: split by position */
    } else {
        float lo = cbounds.lo[best_axis];
        float extent = cbounds.hi[best_axis] - lo;
        uint32_t nl = 0, nr = 0;
        for (uint32_t k = first; k < first + count; k++) { /* stable partition */
            uint32_t t = b->order[k];
            int bin = (int)(((b->centroid[t][best_axis] - lo) / extent) * (float)QM_BINS);
            if (bin > QM_BINS - 1) bin = QM_BINS - 1;
            if ((uint32_t)bin < best_split) {
                b->order[first + nl++] = t;
            } else {
                b->scratch[nr++] = t;
            }
        }
        memcpy(&b->order[first + nl], b->scratch, (size_t)nr * sizeof(uint32_t));
        left_count = nl;
    }

    uint32_t left = b->n_nodes;
    b->n_nodes += 2;
    node->left_first = left;
    node->count = 0;
    build_node(b, left, first, left_count);
    build_node(b, left + 1u, first + left_count, count - left_count);
}

static void build_bvh(qm_scene* s) {
    builder_t b;
    uint32_t n = s->n_tris;
    b.tris = s->tris;
    b.prim_box = (box_t*)malloc((size_t)n * sizeof(box_t));
    b.centroid = (float(*)[3])malloc((size_t)n * sizeof(float[3]));
    b.order = (uint32_t*)malloc((size_t)n * sizeof(uint32_t));
    b.scratch = (uint32_t*)malloc((size_t)n * sizeof(uint32_t));
    b.nodes = (qm_node*)calloc((size_t)2 * n, sizeof(qm_node));
    b.n_nodes = 1;
    b.outlier_first = 0xffffffffu;
    for (uint32_t t = 0; t < n; t++) {
        const qm_triangle* tri = &s->tris[t];
        float p1[3], p2[3];
        for (int a = 0; a < 3; a++) {
            p1[a] = tri->v0[a] + tri->e1[a];
            p2[a] = tri->v0[a] + tri->e2[a];
            b.centroid[t][a] = ((tri->v0[a] + p1[a]) + p2[a]) * (1.0f / 3.0f);
        }
        box_empty(&b.prim_box[t]);
        box_grow_point(&b.prim_box[t], tri->v0);
        box_grow_point(&b.prim_box[t], p1);
        box_grow_point(&b.prim_box[t], p2);
        b.order[t] = t;
    }
    build_node(&b, 0, 0, n);
    s->nodes = (qm_node*)realloc(b.nodes, (size_t)b.n_nodes * sizeof(qm_node));
    s->n_nodes = b.n_nodes;
    s->leaf_order = b.order;
    free(b.prim_box);
    free(b.centroid);
    free(b.scratch);
}

qm_scene* qm_scene_create(uint32_t grid, uint32_t seed) {
    qm_scene* s = (qm_scene*)calloc(1, sizeof *s);
    build_mesh(s, grid, seed);
    build_bvh(s);
    return s;
}

void qm_scene_destroy(qm_scene* s) {
    if (!s) return;
    free(s->tris);
    free(s->nodes);
    free(s->leaf_order);
    free(s);
}

uint32_t qm_triangle_count(const qm_scene* s) { return s->n_tris; }
uint32_t qm_node_count(const qm_scene* s) { return s->n_nodes; }
const qm_triangle* qm_triangles(const qm_scene* s) { return s->tris; }
const qm_node* qm_nodes(const qm_scene* s) { return s->nodes; }
const uint32_t* qm_leaf_order(const qm_scene* s) { return s->leaf_order; }

/* ------------------------------------------------------------ intersection -- */

typedef struct {
    float t;
    uint32_t prim; /* original triangle index; 0xffffffff = miss */
} mesh_hit_t;

/* Moeller-Trumbore. Accepts t in [0.001, best->t]; an equal t wins only with a
 * higher triangle index, so the outcome is independent of visiting order. */
static void qm_intersect(const qm_triangle* tri, uint32_t prim, v3 o, v3 d, mesh_hit_t* best) {
    v3 e1 = v3_load(tri->e1), e2 = v3_load(tri->e2);
    v3 pvec = v3_cross(d, e2);
    float det = v3_dot(e1, pvec);
    if (fabsf(det) < 1e-12f) return;
    float inv_det = 1.0f / det;
    v3 tvec = v3_sub(o, v3_load(tri->v0));
    float u = v3_dot(tvec, pvec) * inv_det;
    if (u < 0.0f || u > 1.0f) return;
    v3 qvec = v3_cross(tvec, e1);
    float v = v3_dot(d, qvec) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return;
    float t = v3_dot(e2, qvec) * inv_det;
    if (t < 0.001f || t > best->t) return;
    if (t == best->t && best->prim != 0xffffffffu && prim < best->prim) return;
    best->t = t;
    best->prim = prim;
}

static int box_hit(const qm_node* n, v3 o, v3 idir, float closest) {
    float t0x = (n->bmin[0] - o.x) * idir.x, t1x = (n->bmax[0] - o.x) * idir.x;
    float t0y = (n->bmin[1] - o.y) * idir.y, t1y = (n->bmax[1] - o.y) * idir.y;
    float t0z = (n->bmin[2] - o.z) * idir.z, t1z = (n->bmax[2] - o.z) * idir.z;
    float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
    float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), closest));
    return tn <= tf;
}

static mesh_hit_t trace_mesh(const qm_scene* s, v3 o, v3 d, int use_bvh, qm_stats* st) {
    mesh_hit_t best;
    best.t = 1e30f;
    best.prim = 0xffffffffu;
    if (!use_bvh) {
        for (uint32_t t = 0; t < s->n_tris; t++) qm_intersect(&s->tris[t], t, o, d, &best);
        st->triangles_tested += s->n_tris;
        return best;
    }
    v3 idir = v3_make(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    uint32_t stack[64];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const qm_node* n = &s->nodes[stack[--sp]];
        st->nodes_visited++;
        if (!box_hit(n, o, idir, best.t)) continue;
        if (n->count != 0) {
            for (uint32_t k = 0; k < n->count; k++) {
                uint32_t prim = s->leaf_order[n->left_first + k];
                qm_intersect(&s->tris[prim], prim, o, d, &best);
            }
            st->triangles_tested += n->count;
        } else {
            stack[sp++] = n->left_first + 1u;
            stack[sp++] = n->left_first;
        }
    }
    return best;
}

/* --------------------------------------------------------------- rendering -- */

typedef struct {
    const qm_scene* scene;
    qo_camera camera;
    uint32_t width, height, frame_index, max_bounces;
    int use_bvh;
    float* layer[4]; /* current-sample layers */
    int tid, nthreads;
    qm_stats stats;
} mesh_job_t;

/* path_trace (plugin.mm:223-276) with trace_scene replaced by the mesh query. */
static void mesh_pixel(mesh_job_t* job, uint32_t px, uint32_t py, float out[16]) {
    const qm_scene* s = job->scene;
    float uv_x = ((float)px + 0.5f) / (float)job->width;
    float uv_y = 1.0f - ((float)py + 0.5f) / (float)job->height;
    uint32_t rng = qo_seed(px, py, job->frame_index);
    ray_t ray = qo_get_camera_ray(&job->camera, uv_x, uv_y, &rng);

    v3 color = v3_make(0, 0, 0), albedo0 = color, normal0 = color, throughput = v3_make(1, 1, 1);
    float depth0 = 0.0f;
    for (uint32_t bounce = 0; bounce < job->max_bounces; bounce++) {
        mesh_hit_t hit = trace_mesh(s, ray.origin, ray.direction, job->use_bvh, &job->stats);
        job->stats.rays++;
        if (hit.prim == 0xffffffffu) break;
        const qm_triangle* tri = &s->tris[hit.prim];
        const qm_material* mat = &s->materials[tri->material];
        v3 emission = v3_load(mat->emission), albedo = v3_load(mat->albedo);
        v3 ng = v3_normalize(v3_cross(v3_load(tri->e1), v3_load(tri->e2)));
        v3 n = (v3_dot(ng, ray.direction) < 0.0f) ? ng : v3_neg(ng);
        v3 point = v3_add(ray.origin, v3_scale(ray.direction, hit.t));
        float emit = fmaxf(emission.x, fmaxf(emission.y, emission.z));
        if (bounce == 0) {
            albedo0 = (emit > 0.1f) ? emission : albedo;
            normal0 = v3_adds(v3_scale(n, 0.5f), 0.5f);
            depth0 = hit.t;
        }
        color = v3_add(color, v3_mul(throughput, emission));
        if (emit > 0.1f) break;
        throughput = v3_mul(throughput, albedo);
        if (bounce > 2) {
            float p = fmaxf(0.05f, fmaxf(throughput.x, fmaxf(throughput.y, throughput.z)));
            if (qo_random_float(&rng) > p) break;
            throughput = v3_div(throughput, p);
        }
        ray.origin = v3_add(point, v3_scale(n, 0.001f));
        ray.direction = qo_cosine_sample_hemisphere(n, &rng);
    }
    float nd = fminf(fmaxf((depth0 - 1.0f) / 4.0f, 0.0f), 1.0f);
    float f[16] = {color.x, color.y, color.z, 1, albedo0.x, albedo0.y, albedo0.z, 1,
                   normal0.x, normal0.y, normal0.z, 1, nd, nd, nd, 1};
    memcpy(out, f, sizeof f);
}

static void* mesh_worker(void* arg) {
    mesh_job_t* job = (mesh_job_t*)arg;
    for (uint32_t py = (uint32_t)job->tid; py < job->height; py += (uint32_t)job->nthreads) {
        for (uint32_t px = 0; px < job->width; px++) {
            float f[16];
            mesh_pixel(job, px, py, f);
            size_t o = ((size_t)py * job->width + px) * 4;
            for (int l = 0; l < 4; l++) memcpy(job->layer[l] + o, f + l * 4, 16);
        }
    }
    return NULL;
}

void qm_render_frames(const qm_scene* scene, const float position[3], const float target[3],
                      const float up[3], float fov, uint32_t width, uint32_t height,
                      uint32_t frame_first, uint32_t frame_stride, uint32_t n_frames,
                      uint32_t max_bounces, int use_bvh, float* const layers[4], int threads,
                      qm_stats* stats) {
    if (threads <= 0) threads = qo_online_cores();
    if ((uint32_t)threads > height) threads = (int)height;
    if (threads < 1) threads = 1;
    uint64_t n_px = (uint64_t)width * height;
    qo_scene_uniforms cam_block;
    memset(&cam_block, 0, sizeof cam_block);
    qo_pack_camera(&cam_block, position, target, up, fov, width, height);

    float* cur[4];
    for (int l = 0; l < 4; l++) {
        cur[l] = (float*)malloc(n_px * 16);
        memset(layers[l], 0, n_px * 16);
    }
    qm_stats total = {0, 0, 0};
    pthread_t* tids = (pthread_t*)calloc((size_t)threads, sizeof *tids);
    mesh_job_t* jobs = (mesh_job_t*)calloc((size_t)threads, sizeof *jobs);
    for (uint32_t i = 0; i < n_frames; i++) {
        for (int t = 0; t < threads; t++) {
            mesh_job_t* j = &jobs[t];
            memset(j, 0, sizeof *j);
            j->scene = scene;
            j->camera = cam_block.camera;
            j->width = width;
            j->height = height;
            j->frame_index = frame_first + i * frame_stride;
            j->max_bounces = max_bounces;
            j->use_bvh = use_bvh;
            for (int l = 0; l < 4; l++) j->layer[l] = cur[l];
            j->tid = t;
            j->nthreads = threads;
            if (t > 0) pthread_create(&tids[t], NULL, mesh_worker, j);
        }
        mesh_worker(&jobs[0]);
        for (int t = 1; t < threads; t++) pthread_join(tids[t], NULL);
        for (int t = 0; t < threads; t++) {
            total.rays += jobs[t].stats.rays;
            total.nodes_visited += jobs[t].stats.nodes_visited;
            total.triangles_tested += jobs[t].stats.triangles_tested;
        }
        for (int l = 0; l < 4; l++) qo_accumulate(cur[l], layers[l], layers[l], n_px, i);
    }
    for (int l = 0; l < 4; l++) free(cur[l]);
    free(tids);
    free(jobs);
    if (stats) *stats = total;
}
