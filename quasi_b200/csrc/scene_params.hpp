/// @file scene_params.hpp
/// @brief Launch-parameter blocks of the CUDA path tracer and the host code that fills them.
///
/// The Metal backend rebuilds a 2 624-byte GpuSceneUniforms block on the CPU every
/// frame and pushes it with setFragmentBytes
/// (/root/reference/backends/metal/plugin.mm:354-387, 662-700, 735); its shader
/// then re-derives every quad's plane, every quad's sampling frame and the
/// camera basis once per ray or per pixel. Here the same quantities are derived
/// ONCE on the host, with the same IEEE binary32 operations in the same order, and
/// handed to the kernel as __grid_constant__ parameters, which live in the
/// constant bank and are read with warp-uniform loads. Hoisting an exactly
/// rounded operation out of a loop cannot change its result, so the images stay
/// bit-identical to the literal shader (tests/test_cuda_parity.py).
///
/// This header is plain C++ (no CUDA types) so that host-only code can use it.

#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>

namespace Q::cuda {

/// Three floats; the vector type of the kernels (device_common.cuh) and of scan_decide.hpp.
struct v3 {
    float x, y, z;
};

inline constexpr uint32_t k_max_quads = 32;   ///< plugin.mm:21, 30
inline constexpr uint32_t k_max_bounces = 5;  ///< plugin.mm:22, 31

/// Which zero pattern a quad has, found on the host by looking for exact zeros.
/// A pattern only selects a cheaper form of the same DECISION arithmetic in the
/// scan (terms that are exactly zero are not computed); every pattern falls back
/// to the same literal test, so a misclassification could cost speed, not bits.
enum quad_shape : uint32_t {
    k_shape_general = 0,
    k_shape_flat_x = 1,    ///< normal along x; alpha, beta live in (y, z)   - side walls
    k_shape_flat_y = 2,    ///< normal along y; alpha, beta live in (x, z)   - floor, ceiling, light, box tops
    k_shape_flat_z = 3,    ///< normal along z; alpha, beta live in (x, y)   - back wall
    k_shape_upright = 4,   ///< normal.y == 0, one edge vertical: alpha in (x, z), beta along y - box sides
    k_shape_boxed = 5,     ///< a face of a detected box (box_params): scanned with its box, not on its own
};

/// One quad, pre-digested: 56 floats (224 bytes), 16-byte rows so constant-bank
/// loads vectorise. Three groups: what the fast scan reads for every ray, what
/// only the exact (literal) test reads, and what shading reads after a hit.
struct alignas(16) quad_params {
    // -- fast scan --
    float normal[3];        ///< n * rsqrt(dot(n,n)), n = cross(u,v)   (plugin.mm:165-170)
    float plane_d;          ///< dot(normal, origin)                   (plugin.mm:171)
    float fast_a[3];        ///< cross(v, w): alpha - 1/2 ~= dot(p - center, fast_a)
    float half_plus;        ///< 1/2 + delta: beyond it the point is surely outside
    float fast_b[3];        ///< cross(w, u): beta  - 1/2 ~= dot(p - center, fast_b)
    float half_minus;       ///< 1/2 - delta: within it the point is surely inside
    float center[3];        ///< origin + u/2 + v/2
    uint32_t shape;         ///< quad_shape: which specialised test applies
    // -- exact test only --
    float w[3];             ///< n / dot(n,n)                          (plugin.mm:182)
    float valid;            ///< 0 when dot(n,n) < 1e-8                (plugin.mm:168)
    float u[3];
    float fast_len;         ///< max(|fast_a|, |fast_b|), scales delta
    float v[3];
    float pad1;
    float origin[3];
    float pad0;
    // -- shading (copied to shared memory, 24 floats) --
    float shade_normal[3];  ///< == normal
    float is_light;         ///< max(emission) > 0.1                   (plugin.mm:255-256)
    /// Cosine-sampling frame for a hit on the side the normal points to
    /// (plugin.mm:119-122). The other side is exactly (frame_u, -frame_v, -frame_w).
    float frame_u[3];
    float shade_plane_d;    ///< == plane_d
    float frame_v[3];
    float pad3;
    float frame_w[3];
    float pad4;
    float albedo[3];
    float pad5;
    float emission[3];
    float pad6;
};
static_assert(sizeof(quad_params) == 224);
inline constexpr uint32_t k_shade_offset = 32;  ///< first float of the shading group
inline constexpr uint32_t k_shade_floats = 24;

/// What the fast scan reads of one flat or upright quad, packed in the order it is used so
/// that the constant bank is read with three 16-byte loads per quad instead of eight loads
/// of 4 and 8 bytes (sm_100 has no constant operands for FP32 instructions: every constant
/// costs an LDCU issue slot, and the scan of the demo scene is 126 of them per ray).
/// The values are copies of quad_params fields, bit for bit (fill_scan_records):
///   flat<K> (I, J the other two axes):
///     {normal[K], plane_d, center[I], center[J]} {fast_a[I], fast_a[J], fast_b[I], fast_b[J]} {half_plus, half_minus}
///   upright:
///     {normal[0], normal[2], plane_d, center[0]} {center[1], center[2], fast_a[0], fast_a[2]} {fast_b[1], half_plus, half_minus}
/// General quads are scanned from quad_params.
struct alignas(16) scan_record {
    float r[12];
};

inline constexpr uint32_t k_shape_count = 5;  ///< shapes the scan walks quad by quad (boxed faces are not among them)
inline constexpr uint32_t k_max_boxes = 4;

/// Four upright quads in two parallel pairs closed by a flat_y quad on top, standing on
/// something (no bottom face): the two blocks of the demo scene. A convex solid is entered
/// at most once, so the closest hit among its five faces is where the ray enters the
/// intersection of the three slabs, and ONE slab test (3 plane distances, one point, 3
/// interval tests) DECIDES what the five separate quad tests decide; pathtrace.cu:scan_box.
/// Like everything in the fast scan this only decides: distances use each face's own plane
/// constants in the shader's form, an entry within `delta` of an edge, a ray parallel to a
/// slab or an entry through the open bottom hand the ray to the literal test, and the
/// accepted hit is recomputed with the literal operations of its face.
struct alignas(16) box_params {
    float axis[2][2];         ///< (x, z) of the two horizontal slab directions (stored normal of one side each)
    float side_normal[2][2][2];  ///< [slab][low, high side][x, z]: that face's own stored normal
    float side_d[2][2];       ///< [slab][low, high side]: that face's own plane_d
    uint32_t side_index[2][2];   ///< [slab][low, high side]: position of the face in scene_params::quads
    float top_normal_y;       ///< stored normal of the top face is (0, top_normal_y, 0)
    float top_d;
    uint32_t top_index;
    float y_bottom;
    float mid[3];             ///< centre of the box in (slab 0, slab 1, y) coordinates
    float half[3];            ///< half extents in the same coordinates
    float delta;              ///< decision margin, in length units (set_decision_margins)
    float edge_scale;         ///< longest face edge: turns a face's alpha margin into a length
};
static_assert(sizeof(box_params) % 16 == 0);

/// The scene as the kernel sees it. `quads` is sorted by shape (stable), so the
/// scan runs one tight loop per shape with that shape's arithmetic;
/// shape_begin[s] .. shape_begin[s+1] is the range of shape s; the faces of detected
/// boxes (shape_begin[5] .. shape_begin[6]) and then the degenerate quads sit behind
/// the ranges the scan walks. The scan order does not matter for the result
/// (closest hit; exact ties go to the literal test). The literal test walks
/// `scene_order`, i.e. the ORIGINAL order, because plugin.mm:205-207 lets the
/// later index win a tie. Hit indices on the device are positions in `quads`.
struct alignas(16) scene_params {
    quad_params quads[k_max_quads];
    uint32_t scene_order[k_max_quads];       ///< original index -> position in quads
    uint32_t shape_begin[k_shape_count + 3]; ///< prefix ranges, padded to 8 entries
    uint32_t quad_count;
    uint32_t max_bounces;
    float corner_radius;    ///< farthest quad corner from the world origin
    uint32_t box_count;
    box_params boxes[k_max_boxes];
    scan_record records[k_max_quads];        ///< per position in `quads` (fill_scan_records): what the scan reads
};

/// Per-call camera and sample bookkeeping.
struct alignas(16) frame_params {
    float position[3];     ///< ray origin
    float weight;          ///< 1 / (sample_index + 1)            (plugin.mm:326)
    float lower_left[3];   ///< plugin.mm:145
    uint32_t width;
    float horizontal[3];   ///< plugin.mm:143
    uint32_t height;
    float vertical[3];     ///< plugin.mm:144
    uint32_t frame_index;  ///< RNG frame index                   (plugin.mm:295)
    uint32_t sample_index; ///< render calls since the last restart
    uint32_t tiles_x_magic; ///< floor(2^32 / ceil(width / 8)), for the kernel's tile -> (x, y)
    uint32_t batch_samples; ///< samples per pixel in this launch (0 or 1: one; QUASI_CUDA_BATCH)
    uint32_t frame_stride;  ///< RNG frame index of sample s of the launch: frame_index + s * frame_stride
};

/// Where a launch of several samples per pixel (frame.batch_samples > 1) sends finished tiles.
/// When a tile has received the LAST sample of the launch, the warp that rendered it copies the
/// tile's pixels of all four layers out of the accumulation block, rows of 8 pixels (128 bytes) at a
/// time: into pinned host memory (the frame is on the host when the launch ends: plugin.cu, readback) or
/// into other GPUs' memory over NVLink (multi-GPU: each rank pushes every slice of its layers to the
/// slice's owner while it renders). The block of 4 * width * height float4 is cut into `export_slices`
/// slices, slice s being the float4 [slice_lo[s], slice_lo[s + 1]); float4 i of slice s goes to
/// export_base[s] + (i - slice_lo[s]). 0 slices: nothing is exported.
///
/// Fused reduce (`reduce_ranks` = G > 1; multi-GPU): export_base[s] is this rank's slot in the staging area of
/// the GPU that owns slice s, and the slices are cut along rows of tiles, so that the 32 pixels of a tile in one
/// layer (a unit) have one owner. After storing a unit a warp adds one to the unit's arrival counter at the owner
/// (system-scope atomic on peer memory, after a system-scope fence); the warp whose arrival is the G-th, on
/// whichever GPU it runs, loads the G staged copies of the unit, sums them IN RANK ORDER, scales by 1/G and stores
/// the unit of the finished frame to reduce_out[s] + (i - slice_lo[s]) (s the unit's slice: pinned host memory
/// every GPU has mapped, or the owner's device memory). The reduce and the frame's way to the host are then part
/// of the render as well: what is left after the last rank's launch is its last tiles. A unit of one staging buffer
/// receives exactly G arrivals per frame that uses the buffer; the G-th puts the counter back to zero (nobody else
/// touches it before the buffer's next frame, which the device-side barrier between frames keeps away).
inline constexpr uint32_t k_max_export_slices = 16;
struct alignas(16) batch_params {
    void* export_base[k_max_export_slices];
    uint32_t slice_lo[k_max_export_slices + 1];
    uint32_t export_slices;
    uint32_t reduce_ranks;         ///< G; 0 or 1: no fused reduce
    uint32_t reduce_slot_stride;   ///< float4 between the slots of two ranks in a staging area
    float reduce_scale;            ///< 1 / G
    void* reduce_staging[k_max_export_slices];             ///< slice s: slot 0 of its owner's staging area
    unsigned int* reduce_counters[k_max_export_slices];    ///< slice s: its owner's arrival counters, [layer * n_tiles + tile]
    void* reduce_out[k_max_export_slices];                 ///< slice s: where float4 slice_lo[s] of the finished frame goes
};

/// The scene as the host describes it: the same 12+8 floats per quad the Metal
/// plugin uploads (GpuQuad / GpuMaterial, plugin.mm:363-377).
struct quad_desc {
    float origin[3];
    float u[3];
    float v[3];
    float albedo[3];
    float roughness;
    float emission[3];
    float metallic;
};

struct scene_desc {
    quad_desc quads[k_max_quads];
    uint32_t quad_count = 0;
    uint32_t light_index = 0;
};

namespace detail {

struct f3 {
    float x, y, z;
};
inline f3 load3(const float* p) { return {p[0], p[1], p[2]}; }
inline void store3(float* p, f3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
inline f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline f3 operator/(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
/// (a.x*b.x + a.y*b.y) + a.z*b.z, /root/reference/src/quasi/math/vec.hpp:92-94
inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/// vec.hpp:100-106
inline f3 cross(f3 a, f3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
/// v / length(v), vec.hpp:108-114
inline f3 normalize(f3 a) { return a / std::sqrt(dot(a, a)); }

/// cornell_box.hpp:32-37
inline f3 rotate_y(f3 p, float angle_deg) {
    float rad = angle_deg * 3.14159265359f / 180.0f;
    float c = std::cos(rad);
    float s = std::sin(rad);
    return {p.x * c + p.z * s, p.y, -p.x * s + p.z * c};
}

}  // namespace detail

/// @brief The demo scene: a Cornell box of 16 quads, light at index 5.
///
/// Same geometry, order and materials as Q::scene::make_cornell_box
/// (/root/reference/src/quasi/scene/cornell_box.hpp:95-134, add_box :45-90);
/// checked bit-for-bit against that function in tests/test_scene.py.
inline scene_desc make_cornell_box() {
    using namespace detail;
    scene_desc s;

    struct mat {
        float albedo[3];
        float emission[3];
    };
    const mat white{{0.73f, 0.73f, 0.73f}, {0, 0, 0}};
    const mat red{{0.65f, 0.05f, 0.05f}, {0, 0, 0}};
    const mat green{{0.12f, 0.45f, 0.15f}, {0, 0, 0}};
    const mat light{{0, 0, 0}, {15.0f, 15.0f, 15.0f}};

    auto push = [&s](f3 origin, f3 u, f3 v, const mat& m) {
        quad_desc& q = s.quads[s.quad_count++];
        store3(q.origin, origin);
        store3(q.u, u);
        store3(q.v, v);
        std::memcpy(q.albedo, m.albedo, sizeof q.albedo);
        std::memcpy(q.emission, m.emission, sizeof q.emission);
        q.roughness = 1.0f;
        q.metallic = 0.0f;
    };

    push({-1, 0, -1}, {2, 0, 0}, {0, 0, 2}, white);   // floor
    push({-1, 2, 1}, {2, 0, 0}, {0, 0, -2}, white);   // ceiling
    push({-1, 0, -1}, {2, 0, 0}, {0, 2, 0}, white);   // back wall
    push({-1, 0, 1}, {0, 0, -2}, {0, 2, 0}, red);     // left wall
    push({1, 0, -1}, {0, 0, 2}, {0, 2, 0}, green);    // right wall
    s.light_index = s.quad_count;
    push({-0.25f, 1.99f, -0.25f}, {0.5f, 0, 0}, {0, 0, 0.5f}, light);

    auto box = [&](f3 center, f3 size, float angle, const mat& m) {
        float hw = size.x * 0.5f, h = size.y, hd = size.z * 0.5f;
        const f3 local[8] = {{-hw, 0, -hd}, {hw, 0, -hd}, {-hw, 0, hd}, {hw, 0, hd},
                             {-hw, h, -hd}, {hw, h, -hd}, {-hw, h, hd}, {hw, h, hd}};
        f3 p[8];
        for (int i = 0; i < 8; i++) p[i] = rotate_y(local[i], angle) + center;
        push(p[4], p[5] - p[4], p[6] - p[4], m);  // top
        push(p[0], p[1] - p[0], p[4] - p[0], m);  // front
        push(p[3], p[2] - p[3], p[7] - p[3], m);  // back
        push(p[2], p[0] - p[2], p[6] - p[2], m);  // left
        push(p[1], p[3] - p[1], p[5] - p[1], m);  // right
    };
    box({-0.35f, 0.0f, 0.3f}, {0.5f, 1.2f, 0.5f}, 15.0f, white);
    box({0.35f, 0.0f, -0.3f}, {0.55f, 0.55f, 0.55f}, -18.0f, white);
    return s;
}

/// @brief Derives the per-quad constants the shader recomputes per ray.
/// @param group_boxes Detect blocks (box_params) and scan each with one slab test
/// (QUASI_CUDA_SCAN=boxes); false keeps every quad on its own, the default.
inline void prepare_scene(const scene_desc& desc, uint32_t max_bounces, scene_params& out, bool group_boxes = false) {
    using namespace detail;
    std::memset(&out, 0, sizeof out);
    out.quad_count = desc.quad_count < k_max_quads ? desc.quad_count : k_max_quads;
    out.max_bounces = max_bounces;

    float corner_radius = 0.0f;
    for (uint32_t i = 0; i < out.quad_count; i++) {
        const quad_desc& q = desc.quads[i];
        for (int c = 0; c < 4; c++) {
            f3 corner = load3(q.origin) + load3(q.u) * float(c & 1) + load3(q.v) * float(c >> 1);
            corner_radius = std::fmax(corner_radius, std::sqrt(dot(corner, corner)));
        }
    }
    out.corner_radius = corner_radius;
    quad_params digested[k_max_quads];
    std::memset(digested, 0, sizeof digested);
    for (uint32_t i = 0; i < out.quad_count; i++) {
        const quad_desc& q = desc.quads[i];
        quad_params& g = digested[i];
        f3 origin = load3(q.origin), u = load3(q.u), v = load3(q.v);
        store3(g.origin, origin);
        store3(g.u, u);
        store3(g.v, v);
        std::memcpy(g.albedo, q.albedo, sizeof g.albedo);
        std::memcpy(g.emission, q.emission, sizeof g.emission);
        float emit = std::fmax(q.emission[0], std::fmax(q.emission[1], q.emission[2]));
        g.is_light = emit > 0.1f ? 1.0f : 0.0f;

        f3 n = cross(u, v);
        float area_sq = dot(n, n);
        if (area_sq < 1e-8f) {
            g.valid = 0.0f;
            continue;
        }
        g.valid = 1.0f;
        f3 normal = n * (1.0f / std::sqrt(area_sq));
        store3(g.normal, normal);
        store3(g.shade_normal, normal);
        g.plane_d = dot(normal, origin);
        g.shade_plane_d = g.plane_d;
        f3 wv = n / area_sq;
        store3(g.w, wv);

        // Fast-scan helpers. alpha = dot(w, cross(planar, v)) equals
        // dot(planar, cross(v, w)) in exact arithmetic; in binary32 the two
        // differ by a few ulps of |planar||a|. See set_decision_margins().
        f3 fa = cross(v, wv);
        f3 fb = cross(wv, u);
        g.fast_len = std::fmax(std::sqrt(dot(fa, fa)), std::sqrt(dot(fb, fb)));
        store3(g.center, origin + u * 0.5f + v * 0.5f);

        // Zero patterns (exact comparisons on purpose, see quad_shape).
        auto zero = [](float x) { return x == 0.0f; };
        const float nn[3] = {normal.x, normal.y, normal.z};
        const float aa[3] = {fa.x, fa.y, fa.z};
        const float bb[3] = {fb.x, fb.y, fb.z};
        g.shape = k_shape_general;
        for (int k = 0; k < 3; k++) {
            const int i = (k + 1) % 3, j = (k + 2) % 3;
            if (zero(nn[i]) && zero(nn[j]) && zero(aa[k]) && zero(bb[k])) g.shape = k_shape_flat_x + k;
        }
        if (g.shape == k_shape_general && zero(nn[1])) {
            // Only max(|alpha - 1/2|, |beta - 1/2|) is ever used, so the two
            // barycentrics may trade places: put the vertical one in fast_b.
            if (zero(aa[0]) && zero(aa[2]) && zero(bb[1])) std::swap(fa, fb);
            if (zero(fb.x) && zero(fb.z) && zero(fa.y)) g.shape = k_shape_upright;
        }
        store3(g.fast_a, fa);
        store3(g.fast_b, fb);

        // cosine_sample_hemisphere's basis for hit.normal == normal.
        f3 w = normalize(normal);
        f3 a = (std::fabs(w.x) > 0.9f) ? f3{0, 1, 0} : f3{1, 0, 0};
        f3 fv = normalize(cross(w, a));
        f3 fu = cross(w, fv);
        store3(g.frame_u, fu);
        store3(g.frame_v, fv);
        store3(g.frame_w, w);
    }

    // Boxes: a flat_y quad whose four corners are the top ends of four upright quads that
    // form two parallel pairs at right angles and share one bottom height.
    struct found_box {
        uint32_t top, side[2][2];
    };
    found_box found[k_max_boxes];
    uint32_t n_found = 0;
    auto corner_of = [&](uint32_t i, int c) {
        const quad_params& g = digested[i];
        return load3(g.origin) + load3(g.u) * float(c & 1) + load3(g.v) * float(c >> 1);
    };
    auto near_xz = [](f3 a, f3 b) { return std::fabs(a.x - b.x) < 1e-5f && std::fabs(a.z - b.z) < 1e-5f; };
    for (uint32_t t = 0; group_boxes && t < out.quad_count && n_found < k_max_boxes; t++) {
        if (digested[t].valid == 0.0f || digested[t].shape != k_shape_flat_y) continue;
        const float y_top = digested[t].origin[1];
        uint32_t sides[4];
        uint32_t n_sides = 0;
        float y_bottom = 0.0f;
        bool consistent = true;
        for (uint32_t i = 0; i < out.quad_count; i++) {
            if (digested[i].valid == 0.0f || digested[i].shape != k_shape_upright) continue;
            float lo = 1e30f, hi = -1e30f;
            for (int c = 0; c < 4; c++) {
                lo = std::fmin(lo, corner_of(i, c).y);
                hi = std::fmax(hi, corner_of(i, c).y);
            }
            if (std::fabs(hi - y_top) > 1e-5f) continue;
            int matched = 0;  // top corners of the side that are corners of the top quad
            for (int c = 0; c < 4; c++) {
                const f3 p = corner_of(i, c);
                if (std::fabs(p.y - hi) > 1e-5f) continue;
                for (int k = 0; k < 4; k++) matched += near_xz(p, corner_of(t, k)) ? 1 : 0;
            }
            if (matched != 2) continue;
            if (n_sides == 4) {
                consistent = false;
                break;
            }
            if (n_sides == 0) y_bottom = lo;
            consistent = consistent && std::fabs(lo - y_bottom) < 1e-5f;
            sides[n_sides++] = i;
        }
        if (!consistent || n_sides != 4) continue;
        // pair the sides: sides[0] with the one parallel to it, the other two with each other
        auto hdot = [&](uint32_t a, uint32_t b) {
            return digested[a].normal[0] * digested[b].normal[0] + digested[a].normal[2] * digested[b].normal[2];
        };
        int partner = -1;
        for (int k = 1; k < 4; k++) {
            if (std::fabs(hdot(sides[0], sides[k])) > 1.0f - 1e-5f) partner = partner < 0 ? k : 4;
        }
        if (partner < 1 || partner > 3) continue;
        uint32_t other[2];
        int n_other = 0;
        for (int k = 1; k < 4; k++) {
            if (k != partner) other[n_other++] = sides[k];
        }
        if (std::fabs(hdot(other[0], other[1])) <= 1.0f - 1e-5f || std::fabs(hdot(sides[0], other[0])) > 1e-4f) continue;
        found_box& b = found[n_found++];
        b.top = t;
        b.side[0][0] = sides[0];
        b.side[0][1] = sides[partner];
        b.side[1][0] = other[0];
        b.side[1][1] = other[1];
        digested[t].shape = k_shape_boxed;
        for (uint32_t k = 0; k < 4; k++) digested[sides[k]].shape = k_shape_boxed;
    }

    // Stable sort by shape; boxed faces, then degenerate quads (never hit), go last, outside
    // every range the scan walks.
    uint32_t position = 0;
    for (uint32_t shape = 0; shape <= k_shape_boxed + 1; shape++) {
        if (shape <= k_shape_boxed) out.shape_begin[shape] = position;
        for (uint32_t i = 0; i < out.quad_count; i++) {
            const bool degenerate = digested[i].valid == 0.0f;
            const bool take = shape <= k_shape_boxed ? (!degenerate && digested[i].shape == shape) : degenerate;
            if (take) {
                out.quads[position] = digested[i];
                out.scene_order[i] = position;
                position++;
            }
        }
        if (shape == k_shape_boxed) {
            for (uint32_t k = k_shape_boxed + 1; k < k_shape_count + 3; k++) out.shape_begin[k] = position;
        }
    }

    out.box_count = n_found;
    for (uint32_t n = 0; n < n_found; n++) {
        box_params& b = out.boxes[n];
        const quad_params& top = digested[found[n].top];
        b.top_normal_y = top.normal[1];
        b.top_d = top.plane_d;
        b.top_index = out.scene_order[found[n].top];
        float y_bottom = 1e30f, longest = 0.0f, shortest = 1e30f;
        for (int slab = 0; slab < 2; slab++) {
            const quad_params& first = digested[found[n].side[slab][0]];
            b.axis[slab][0] = first.normal[0];
            b.axis[slab][1] = first.normal[2];
            float offset[2];
            for (int k = 0; k < 2; k++) {
                const quad_params& g = digested[found[n].side[slab][k]];
                offset[k] = b.axis[slab][0] * g.center[0] + b.axis[slab][1] * g.center[2];
                for (int c = 0; c < 4; c++) y_bottom = std::fmin(y_bottom, corner_of(found[n].side[slab][k], c).y);
                const float lu = std::sqrt(dot(load3(g.u), load3(g.u))), lv = std::sqrt(dot(load3(g.v), load3(g.v)));
                longest = std::fmax(longest, std::fmax(lu, lv));
                shortest = std::fmin(shortest, std::fmin(lu, lv));
            }
            const int low = offset[0] <= offset[1] ? 0 : 1;
            for (int k = 0; k < 2; k++) {
                const uint32_t original = found[n].side[slab][k == 0 ? low : 1 - low];
                const quad_params& g = digested[original];
                b.side_normal[slab][k][0] = g.normal[0];
                b.side_normal[slab][k][1] = g.normal[2];
                b.side_d[slab][k] = g.plane_d;
                b.side_index[slab][k] = out.scene_order[original];
            }
            b.mid[slab] = (offset[0] + offset[1]) * 0.5f;
            b.half[slab] = std::fabs(offset[1] - offset[0]) * 0.5f;
        }
        const float y_top = top.origin[1];
        b.y_bottom = y_bottom;
        b.mid[2] = (y_bottom + y_top) * 0.5f;
        b.half[2] = (y_top - y_bottom) * 0.5f;
        b.edge_scale = longest;
        (void)shortest;
        b.delta = 1e-4f;  // set_decision_margins
    }
}

/// @brief Copies what the scan reads of every flat and upright quad into scene.records.
inline void fill_scan_records(scene_params& scene) {
    for (uint32_t i = 0; i < scene.quad_count; i++) {
        const quad_params& g = scene.quads[i];
        float* r = scene.records[i].r;
        for (int k = 0; k < 12; k++) r[k] = 0.0f;
        int axis = -1;
        if (g.shape == k_shape_flat_x) axis = 0;
        if (g.shape == k_shape_flat_y) axis = 1;
        if (g.shape == k_shape_flat_z) axis = 2;
        if (axis >= 0) {
            const int a = (axis + 1) % 3, b = (axis + 2) % 3;
            const float packed[10] = {g.normal[axis], g.plane_d, g.center[a], g.center[b], g.fast_a[a],
                                      g.fast_a[b],    g.fast_b[a], g.fast_b[b], g.half_plus, g.half_minus};
            for (int k = 0; k < 10; k++) r[k] = packed[k];
        } else if (g.shape == k_shape_upright || g.shape == k_shape_boxed) {
            const float packed[11] = {g.normal[0], g.normal[2], g.plane_d,   g.center[0], g.center[1],  g.center[2],
                                      g.fast_a[0], g.fast_a[2], g.fast_b[1], g.half_plus, g.half_minus};
            for (int k = 0; k < 11; k++) r[k] = packed[k];
        }
    }
}

/// @brief Sets the per-quad decision margin of the fast scan for a camera position.
///
/// The fast scan decides "is the plane hit inside the quad" from
/// alpha' - 1/2 = dot(p' - center, fast_a) with p' = ro + rd * t', t' = num * rcp(denom)
/// (center = origin + u/2 + v/2, so |alpha' - 1/2| <= 1/2 is the inside test).
/// Against the literal alpha of plugin.mm:179-184 that differs by at most about
/// |fast_a| * R * (16 * 2^-24 + 2 * 2^-21) ~= 1.9e-6 * |fast_a| * R, where R
/// bounds the distance of ray origins and hit points from the world origin
/// (rounding of both forms plus the error of the approximate t). delta is 13x
/// that bound. A candidate within delta of an edge is handed to the literal
/// test, so delta trades speed for nothing else: it never reaches a pixel.
inline void set_decision_margins(scene_params& scene, const float camera_position[3]) {
    using namespace detail;
    f3 c = load3(camera_position);
    float radius = std::fmax(scene.corner_radius, std::sqrt(dot(c, c)));
    for (uint32_t i = 0; i < scene.quad_count; i++) {
        quad_params& g = scene.quads[i];
        float delta = std::fmax(1e-5f, 2.5e-5f * g.fast_len * radius);
        g.half_plus = 0.5f + delta;
        g.half_minus = 0.5f - delta;
    }
    // A box decides in length units: the widest alpha margin of its faces times its longest
    // edge, i.e. the widest band any of its faces keeps around its own edges.
    for (uint32_t n = 0; n < scene.box_count; n++) {
        box_params& b = scene.boxes[n];
        float widest = 0.0f;
        const uint32_t faces[5] = {b.top_index, b.side_index[0][0], b.side_index[0][1], b.side_index[1][0],
                                   b.side_index[1][1]};
        for (uint32_t f : faces) widest = std::fmax(widest, scene.quads[f].half_plus - 0.5f);
        b.delta = widest * b.edge_scale;
    }
    fill_scan_records(scene);
}

/// @brief Camera block for one render call: plugin.mm:648-674 (host) followed by
/// the pixel-invariant half of get_camera_ray, plugin.mm:134-145.
///
/// tan() is the correctly rounded binary32 tangent (binary64 libm, rounded
/// once), the same definition the oracle uses.
inline void prepare_frame(const float position[3], const float target[3], const float up_in[3],
                          float fov, uint32_t width, uint32_t height, uint32_t frame_index,
                          uint32_t sample_index, frame_params& out) {
    using namespace detail;
    std::memset(&out, 0, sizeof out);
    float aspect = static_cast<float>(width) / static_cast<float>(height);
    float dx = target[0] - position[0];
    float dy = target[1] - position[1];
    float dz = target[2] - position[2];
    float len = std::sqrt(dx * dx + dy * dy + dz * dz);
    if (len > 0.0f) {
        dx /= len;
        dy /= len;
        dz /= len;
    }
    constexpr float pi_f = 3.14159265358979323846264338327950288f;
    float theta = fov * pi_f / 180.0f;
    float h = static_cast<float>(std::tan(static_cast<double>(theta / 2.0f)));
    float viewport_height = 2.0f * h;
    float viewport_width = aspect * viewport_height;

    f3 pos = load3(position);
    f3 w = normalize(-f3{dx, dy, dz});
    f3 right = normalize(cross(load3(up_in), w));
    f3 up = cross(w, right);
    f3 horizontal = right * viewport_width;
    f3 vertical = up * viewport_height;
    f3 lower_left = pos - horizontal * 0.5f - vertical * 0.5f - w;

    store3(out.position, pos);
    store3(out.lower_left, lower_left);
    store3(out.horizontal, horizontal);
    store3(out.vertical, vertical);
    out.width = width;
    out.height = height;
    out.frame_index = frame_index;
    out.sample_index = sample_index;
    out.weight = 1.0f / static_cast<float>(sample_index + 1u);
    const uint64_t tiles_x = (static_cast<uint64_t>(width) + 7u) / 8u;
    out.tiles_x_magic = tiles_x > 1 ? static_cast<uint32_t>((uint64_t{1} << 32) / tiles_x) : 0xffffffffu;
}

}  // namespace Q::cuda
