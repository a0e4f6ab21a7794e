// Host check of the quad scene's fast scan (quasi_b200/csrc/scan_decide.hpp) against the
// literal loop of /root/reference/backends/metal/plugin.mm:157-212 (trace_scene +
// intersect_quad, restated below in the shader's operation order), on the demo scene.
// Whenever the fast scan does not ask for the literal loop, its winner must be the
// literal winner. Rays: camera rays, bounce rays from random points of every quad, and
// rays aimed at what is hard: the edges of the boxes' faces, the floor within t_min of a
// box, directions almost parallel to a face.
//   usage: scan_decide_check [rays] [quads]   -> prints counts, exit code 1 on any mismatch
//   "quads": every quad on its own (the backend's default); otherwise the blocks are grouped
#include "scan_decide.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>

using namespace Q::cuda;

namespace {

float dot3(const float* a, v3 b) { return a[0] * b.x + a[1] * b.y + a[2] * b.z; }

int literal_winner(const scene_params& scene, v3 ro, v3 rd) {
    float closest = 1e30f;
    int index = -1;
    for (uint32_t k = 0; k < scene.quad_count; k++) {
        const uint32_t i = scene.scene_order[k];
        const quad_params& q = scene.quads[i];
        if (q.valid == 0.0f) continue;
        const float denom = dot3(q.normal, rd);
        if (std::fabs(denom) < 1e-8f) continue;
        const float t = (q.plane_d - dot3(q.normal, ro)) / denom;
        if (t < 0.001f || t > closest) continue;
        const v3 planar{ro.x + rd.x * t - q.origin[0], ro.y + rd.y * t - q.origin[1], ro.z + rd.z * t - q.origin[2]};
        const v3 pv{planar.y * q.v[2] - planar.z * q.v[1], planar.z * q.v[0] - planar.x * q.v[2],
                    planar.x * q.v[1] - planar.y * q.v[0]};
        const v3 up{q.u[1] * planar.z - q.u[2] * planar.y, q.u[2] * planar.x - q.u[0] * planar.z,
                    q.u[0] * planar.y - q.u[1] * planar.x};
        const float alpha = dot3(q.w, pv), beta = dot3(q.w, up);
        if (alpha < 0.0f || alpha > 1.0f || beta < 0.0f || beta > 1.0f) continue;
        closest = t;
        index = static_cast<int>(i);
    }
    return index;
}

bool inside_a_box_or_outside_the_room(const scene_params& scene, v3 p) {  // `scene`: the grouped form
    if (std::fabs(p.x) > 1.0f || p.y < 0.0f || p.y > 2.0f || p.z < -1.0f) return true;
    for (uint32_t n = 0; n < scene.box_count; n++) {
        const box_params& b = scene.boxes[n];
        const float c0 = b.axis[0][0] * p.x + b.axis[0][1] * p.z, c1 = b.axis[1][0] * p.x + b.axis[1][1] * p.z;
        if (std::fabs(c0 - b.mid[0]) < b.half[0] && std::fabs(c1 - b.mid[1]) < b.half[1] &&
            std::fabs(p.y - b.mid[2]) < b.half[2]) {
            return true;
        }
    }
    return false;
}

}  // namespace

int main(int argc, char** argv) {
    const long rays = argc > 1 ? std::atol(argv[1]) : 3000000;
    const bool group_boxes = !(argc > 2 && std::string(argv[2]) == "quads");
    scene_params scene;
    const scene_desc desc = make_cornell_box();
    prepare_scene(desc, 5, scene, group_boxes);
    const float camera[3] = {0.0f, 1.0f, 3.5f};
    set_decision_margins(scene, camera);
    if (scene.box_count != (group_boxes ? 2u : 0u)) {
        std::printf("expected the two blocks of the demo scene to be detected, found %u\n", scene.box_count);
        return 1;
    }
    scene_params grouped;  // the blocks' geometry, for aiming rays, whichever form is checked
    prepare_scene(desc, 5, grouped, true);
    set_decision_margins(grouped, camera);
    const uint32_t block_faces = 10;  // quads 6..15 of make_cornell_box
    uint32_t floor_position = 0;  // the floor is quad 0 of make_cornell_box
    floor_position = scene.scene_order[0];

    std::mt19937 rng(7);
    std::uniform_real_distribution<float> uniform(0.0f, 1.0f);
    auto coin = [&] { return (rng() & 1u) != 0u; };
    long tested = 0, decided = 0, literal = 0, mismatches = 0;
    for (long it = 0; it < rays; it++) {
        v3 ro{}, rd{};
        const int kind = static_cast<int>(it % 5);
        if (kind == 4 && coin()) {
            // A ray that meets a quad at t = t_min give or take 1e-5 of it: the band (k_t_min_lo, k_t_min_hi] in
            // which t' cannot tell which side of plugin.mm:177 the shader's t falls on (once per ray, on the winner).
            const quad_params& q = scene.quads[rng() % scene.quad_count];
            const float a = 0.05f + 0.9f * uniform(rng), b = 0.05f + 0.9f * uniform(rng);
            const v3 target{q.origin[0] + q.u[0] * a + q.v[0] * b, q.origin[1] + q.u[1] * a + q.v[1] * b,
                            q.origin[2] + q.u[2] * a + q.v[2] * b};
            float len2;
            do {
                rd = {2.0f * uniform(rng) - 1.0f, 2.0f * uniform(rng) - 1.0f, 2.0f * uniform(rng) - 1.0f};
                len2 = rd.x * rd.x + rd.y * rd.y + rd.z * rd.z;
            } while (len2 > 1.0f || len2 < 1e-2f);
            const float len = std::sqrt(len2);
            rd = {rd.x / len, rd.y / len, rd.z / len};
            const float t0 = 0.001f * (1.0f + (uniform(rng) - 0.5f) * (coin() ? 2e-5f : 1e-6f));
            ro = {target.x - rd.x * t0, target.y - rd.y * t0, target.z - rd.z * t0};
            if (inside_a_box_or_outside_the_room(grouped, ro)) continue;
        } else if (kind == 4) {
            // A ray almost IN the plane of a quad, starting within a few 1e-8 of that plane and aimed at a
            // point of the quad: its |denom| straddles the shader's 1e-8 rejection (plugin.mm:174) while
            // t' = num / denom is an ordinary distance. The scan takes that rejection once per ray, on the winner.
            const quad_params& q = scene.quads[rng() % scene.quad_count];
            const float a = 0.05f + 0.9f * uniform(rng), b = 0.05f + 0.9f * uniform(rng);
            const v3 target{q.origin[0] + q.u[0] * a + q.v[0] * b, q.origin[1] + q.u[1] * a + q.v[1] * b,
                            q.origin[2] + q.u[2] * a + q.v[2] * b};
            const float angle = 6.2831853f * uniform(rng), reach = 0.05f + 1.5f * uniform(rng);
            const float ulen = std::sqrt(q.u[0] * q.u[0] + q.u[1] * q.u[1] + q.u[2] * q.u[2]);
            const float vlen = std::sqrt(q.v[0] * q.v[0] + q.v[1] * q.v[1] + q.v[2] * q.v[2]);
            const float cu = std::cos(angle) * reach / ulen, cv = std::sin(angle) * reach / vlen;
            const float slope = (coin() ? 1.0f : -1.0f) * std::pow(10.0f, -9.5f + 3.0f * uniform(rng));  // 3e-10 .. 3e-7
            const float lift = slope * reach;  // so that the ray comes down to the plane at the target
            ro = {target.x - q.u[0] * cu - q.v[0] * cv + q.normal[0] * lift,
                  target.y - q.u[1] * cu - q.v[1] * cv + q.normal[1] * lift,
                  target.z - q.u[2] * cu - q.v[2] * cv + q.normal[2] * lift};
            rd = {target.x - ro.x, target.y - ro.y, target.z - ro.z};
            if (inside_a_box_or_outside_the_room(grouped, ro)) continue;
        } else if (kind == 0 && coin()) {  // camera rays of the default camera
            ro = {0.0f, 1.0f, 3.5f};
            rd = {-0.6f + 1.2f * uniform(rng), -0.5f + uniform(rng), -1.0f};
        } else if (kind == 0) {  // a camera anywhere around the room, looking at a point in or near it
            float len2;
            do {
                ro = {2.0f * uniform(rng) - 1.0f, 2.0f * uniform(rng) - 1.0f, 2.0f * uniform(rng) - 1.0f};
                len2 = ro.x * ro.x + ro.y * ro.y + ro.z * ro.z;
            } while (len2 > 1.0f || len2 < 0.01f);
            const float radius = (2.0f + 3.0f * uniform(rng)) / std::sqrt(len2);
            ro = {ro.x * radius, 1.0f + ro.y * radius, ro.z * radius};
            const v3 target{2.4f * uniform(rng) - 1.2f, 2.4f * uniform(rng) - 0.2f, 2.4f * uniform(rng) - 1.2f};
            rd = {target.x - ro.x, target.y - ro.y, target.z - ro.z};
        } else {
            uint32_t qi = rng() % scene.quad_count;
            if (kind >= 2) qi = coin() ? floor_position : scene.scene_order[6u + rng() % block_faces];
            const quad_params& q = scene.quads[qi];
            float a = uniform(rng), b = uniform(rng), side = coin() ? 1.0f : -1.0f;
            int grazed_box = -1, grazed_slab = 0;  // kind 3: a floor point in the plane of a box side
            if (kind >= 2 && qi != floor_position) {  // hug the edges of a face of a box or of the room
                if (coin()) a = coin() ? uniform(rng) * 0.01f : 1.0f - uniform(rng) * 0.01f;
                if (coin()) b = coin() ? uniform(rng) * 0.01f : 1.0f - uniform(rng) * 0.01f;
            }
            if (kind >= 2 && qi == floor_position) {  // the floor around the foot of a box
                const uint32_t which_box = rng() % grouped.box_count;
                const box_params& box = grouped.boxes[which_box];
                float c0 = box.mid[0] + (2.0f * uniform(rng) - 1.0f) * (box.half[0] + 0.004f);
                float c1 = box.mid[1] + (2.0f * uniform(rng) - 1.0f) * (box.half[1] + 0.004f);
                float off = uniform(rng) * 0.004f - 0.0005f;
                const bool first_slab = coin();
                if (kind == 3 && coin()) {  // within a few 1e-8 of the plane of a side, to be grazed below
                    off = (uniform(rng) - 0.5f) * 2e-7f;
                    grazed_box = static_cast<int>(which_box);
                    grazed_slab = first_slab ? 0 : 1;
                }
                if (first_slab) {
                    c0 = box.mid[0] + (coin() ? 1.0f : -1.0f) * (box.half[0] + off);
                } else {
                    c1 = box.mid[1] + (coin() ? 1.0f : -1.0f) * (box.half[1] + off);
                }
                const float x = box.axis[0][0] * c0 + box.axis[1][0] * c1, z = box.axis[0][1] * c0 + box.axis[1][1] * c1;
                a = (x + 1.0f) * 0.5f;
                b = (z + 1.0f) * 0.5f;
                side = -1.0f;  // the floor's stored normal points down
            }
            ro = {q.origin[0] + q.u[0] * a + q.v[0] * b + q.normal[0] * 0.001f * side,
                  q.origin[1] + q.u[1] * a + q.v[1] * b + q.normal[1] * 0.001f * side,
                  q.origin[2] + q.u[2] * a + q.v[2] * b + q.normal[2] * 0.001f * side};
            float len2;
            do {
                rd = {2.0f * uniform(rng) - 1.0f, 2.0f * uniform(rng) - 1.0f, 2.0f * uniform(rng) - 1.0f};
                len2 = rd.x * rd.x + rd.y * rd.y + rd.z * rd.z;
            } while (len2 > 1.0f || len2 < 1e-4f);
            if (kind == 3 && (grazed_box >= 0 || coin())) {  // almost parallel to a face of a box or to the floor
                const box_params& box = grouped.boxes[grazed_box >= 0 ? static_cast<uint32_t>(grazed_box) : rng() % grouped.box_count];
                const int which = grazed_box >= 0 ? grazed_slab : static_cast<int>(rng() % 3u);
                const float tiny = (uniform(rng) - 0.5f) * (coin() ? 4e-6f : 4e-4f);
                if (which == 2) {
                    rd.y = tiny;
                } else {  // remove the component along the slab direction, leave a tiny one
                    const float along = box.axis[which][0] * rd.x + box.axis[which][1] * rd.z;
                    rd.x += (tiny - along) * box.axis[which][0];
                    rd.z += (tiny - along) * box.axis[which][1];
                }
            }
            if ((rd.x * q.normal[0] + rd.y * q.normal[1] + rd.z * q.normal[2]) * side < 0.0f) {
                rd = {-rd.x, -rd.y, -rd.z};
            }
            if (inside_a_box_or_outside_the_room(grouped, ro)) continue;  // no path ever starts there
        }
        const float len = std::sqrt(rd.x * rd.x + rd.y * rd.y + rd.z * rd.z);
        rd = {rd.x / len, rd.y / len, rd.z / len};

        scan::scan_state state;
        scan::scan_scene(state, scene, ro, rd);
        tested++;
        bool to_literal = scan::needs_literal_scan(state);
        if (!to_literal && state.best_i >= 0) {  // the kernel's finalisation: the winner's own denominator
            to_literal = !scan::accept_winner(dot3(scene.quads[state.best_i].normal, rd));
        }
        if (to_literal) {
            literal++;
            continue;
        }
        decided++;
        const int want = literal_winner(scene, ro, rd);
        if (state.best_i != want) {
            if (mismatches < 10) {
                std::printf("MISMATCH fast %d literal %d ro %.9g %.9g %.9g rd %.9g %.9g %.9g\n", state.best_i, want, ro.x,
                            ro.y, ro.z, rd.x, rd.y, rd.z);
            }
            mismatches++;
        }
    }
    std::printf("rays %ld decided %ld literal %ld mismatches %ld\n", tested, decided, literal, mismatches);
    return mismatches == 0 ? 0 : 1;
}
