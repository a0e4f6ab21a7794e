/// @file lbvh_build.cu
/// @brief BVH construction on the GPU for the mesh scene, sm_100a.
///
/// Nothing in the reference builds an acceleration structure (its scene is 16
/// quads in a uniform block, /root/reference/backends/metal/plugin.mm:199-212),
/// so there is no reference behaviour to match. The contract is the one of
/// mesh_trace.cu: ANY tree whose boxes contain their triangles with the padding of
/// mesh_scene.hpp gives the image of oracle/quasi_oracle_mesh.c bit for bit,
/// because the closest hit does not depend on the order of the visits. That is
/// what tests/test_cuda_mesh.py checks for the tree built here.
///
/// Linear BVH (Lauterbach et al. 2009) with the fully parallel hierarchy of
/// Karras 2012:
///   1. primitive boxes + centroids, scene centroid bounds (block reduction, one
///      ordered-integer atomic per block and axis)
///   2. 30-bit Morton code of the centroid; key = outlier << 62 | code << 32 |
///      triangle index, so keys are unique and ties need no special case. An
///      outlier is a triangle whose box exceeds 1/64 of the scene's box (the rule
///      of mesh_scene.hpp:build_bvh): the top bit makes the root split outliers |
///      everything else, instead of letting each wall inflate one box per level
///   3. 64-bit radix sort of the keys (cub::DeviceRadixSort, library code)
///   4. one thread per internal node: range and split from the common-prefix
///      lengths of neighbouring keys
///   5. boxes bottom-up: one thread per leaf walks to the root, the second thread
///      to arrive at a node merges its children
///   6. emission into the traversal layout: 32-byte nodes, the two children of
///      internal node i at 1 + 2i and 2 + 2i (siblings share a 64-byte line), a
///      subtree of at most 2 triangles that are not outliers becomes one leaf (its
///      range of the sorted order is contiguous), boxes padded by 1e-4 of their
///      extent + 1e-5
///   7. triangles and shading records gathered into leaf (= sorted) order
/// The build of the 999 710-triangle scene takes 0.57 ms of kernels against 2.3 s for the
/// binned-SAH build on one host core; rays visit 42 boxes on this tree against 37
/// on the SAH tree (frame 28.5 ms against 26.2 ms), the usual LBVH trade.

#include "kernels.hpp"

#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

#include <cfloat>

namespace Q::cuda {

namespace {

constexpr int k_threads = 256;
constexpr uint32_t k_leaf_flag = 0x80000000u;  ///< child reference: leaf k (position in sorted order)
constexpr uint32_t k_leaf_max = 2;
constexpr uint64_t k_outlier_bit = 1ull << 62;
constexpr float k_outlier_fraction = 1.0f / 64.0f;  // mesh_scene.hpp:k_outlier_fraction

inline unsigned blocks_for(uint32_t n) { return (n + k_threads - 1) / k_threads; }

/// float <-> unsigned with the same order, for atomicMin / atomicMax.
__device__ __forceinline__ uint32_t ordered(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unordered(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

struct build_arrays {
    const float4* tri_v0;  ///< original order: {v0, material}
    const float4* tri_e1;  ///< {e1, -}
    const float4* tri_e2;  ///< {e2, -}
    float4* prim_lo;       ///< [n] box of triangle t, w = centroid.x
    float4* prim_hi;       ///< [n] w = centroid.y; centroid.z in cz
    float* prim_cz;        ///< [n]
    uint32_t* scene;       ///< [12] ordered-uint: centroid bounds lo xyz, hi xyz, then scene box lo xyz, hi xyz
    uint64_t* keys;        ///< [n] sorted keys (after the sort)
    uint32_t* child;       ///< [2 (n-1)] children of internal node i at 2i, 2i+1
    uint32_t* range;       ///< [2 (n-1)] first, last sorted position covered by internal node i
    uint32_t* parent;      ///< [n-1 internal][n leaf]: parent internal node (root: itself)
    uint32_t* visits;      ///< [n-1] arrival counters of the bottom-up pass
    float4* box_lo;        ///< [n-1]
    float4* box_hi;        ///< [n-1]
};

__global__ void __launch_bounds__(k_threads)
prim_bounds_kernel(build_arrays a, uint32_t n) {
    const uint32_t t = blockIdx.x * k_threads + threadIdx.x;
    float c[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, d[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    float b_lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, b_hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (t < n) {
        const float4 v0 = a.tri_v0[t], e1 = a.tri_e1[t], e2 = a.tri_e2[t];
        const float p0[3] = {v0.x, v0.y, v0.z};
        const float p1[3] = {v0.x + e1.x, v0.y + e1.y, v0.z + e1.z};  // mesh_scene.hpp:build_bvh
        const float p2[3] = {v0.x + e2.x, v0.y + e2.y, v0.z + e2.z};
        float lo[3], hi[3];
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(p0[k], fminf(p1[k], p2[k]));
            hi[k] = fmaxf(p0[k], fmaxf(p1[k], p2[k]));
            c[k] = d[k] = ((p0[k] + p1[k]) + p2[k]) * (1.0f / 3.0f);
            b_lo[k] = lo[k];
            b_hi[k] = hi[k];
        }
        a.prim_lo[t] = make_float4(lo[0], lo[1], lo[2], c[0]);
        a.prim_hi[t] = make_float4(hi[0], hi[1], hi[2], c[1]);
        a.prim_cz[t] = c[2];
    }
    // block reduction of the centroid bounds and of the scene box, then one atomic per block and axis
    __shared__ float s_lo[6][k_threads / 32], s_hi[6][k_threads / 32];
    for (int k = 0; k < 6; k++) {
        float lo = k < 3 ? c[k] : b_lo[k - 3], hi = k < 3 ? d[k] : b_hi[k - 3];
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) {
            s_lo[k][threadIdx.x >> 5] = lo;
            s_hi[k][threadIdx.x >> 5] = hi;
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float lo = s_lo[threadIdx.x][0], hi = s_hi[threadIdx.x][0];
        for (int w = 1; w < k_threads / 32; w++) {
            lo = fminf(lo, s_lo[threadIdx.x][w]);
            hi = fmaxf(hi, s_hi[threadIdx.x][w]);
        }
        const uint32_t group = threadIdx.x < 3 ? 0u : 6u, axis = threadIdx.x % 3u;
        atomicMin(a.scene + group + axis, ordered(lo));
        atomicMax(a.scene + group + 3 + axis, ordered(hi));
    }
}

/// 10 bits -> every third bit.
__device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void __launch_bounds__(k_threads)
morton_kernel(build_arrays a, uint64_t* keys_out, uint32_t n) {
    const uint32_t t = blockIdx.x * k_threads + threadIdx.x;
    if (t >= n) return;
    const float c[3] = {a.prim_lo[t].w, a.prim_hi[t].w, a.prim_cz[t]};
    uint32_t q[3];
    for (int k = 0; k < 3; k++) {
        const float lo = unordered(a.scene[k]), hi = unordered(a.scene[3 + k]);
        const float extent = hi - lo;
        const float u = extent > 0.0f ? (c[k] - lo) / extent : 0.0f;
        q[k] = static_cast<uint32_t>(fminf(fmaxf(u * 1024.0f, 0.0f), 1023.0f));
    }
    const uint32_t code = (expand_bits(q[0]) << 2) | (expand_bits(q[1]) << 1) | expand_bits(q[2]);
    // outlier: box area above 1/64 of the scene box's (both 2 (dx dy + dy dz + dz dx), mesh_scene.hpp)
    float ext[3], own[3];
    const float4 lo4 = a.prim_lo[t], hi4 = a.prim_hi[t];
    const float own_lo[3] = {lo4.x, lo4.y, lo4.z}, own_hi[3] = {hi4.x, hi4.y, hi4.z};
    for (int k = 0; k < 3; k++) {
        ext[k] = unordered(a.scene[9 + k]) - unordered(a.scene[6 + k]);
        own[k] = own_hi[k] - own_lo[k];
    }
    const float scene_area = 2.0f * ((ext[0] * ext[1] + ext[1] * ext[2]) + ext[2] * ext[0]);
    const float own_area = 2.0f * ((own[0] * own[1] + own[1] * own[2]) + own[2] * own[0]);
    const uint64_t outlier = own_area > scene_area * k_outlier_fraction ? k_outlier_bit : 0ull;
    keys_out[t] = outlier | (static_cast<uint64_t>(code) << 32) | t;
}

/// Length of the common prefix of keys i and j; -1 outside the array (Karras 2012, section 4).
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll(static_cast<long long>(keys[i] ^ keys[j]));
}

__global__ void __launch_bounds__(k_threads)
hierarchy_kernel(build_arrays a, uint32_t n) {
    const int i = static_cast<int>(blockIdx.x * k_threads + threadIdx.x);
    const int count = static_cast<int>(n);
    if (i >= count - 1) return;
    const uint64_t* keys = a.keys;
    const int d = delta(keys, count, i, i + 1) - delta(keys, count, i, i - 1) >= 0 ? 1 : -1;
    const int d_min = delta(keys, count, i, i - d);
    int l_max = 2;
    while (delta(keys, count, i, i + l_max * d) > d_min) l_max *= 2;
    int l = 0;
    for (int t = l_max / 2; t >= 1; t /= 2) {
        if (delta(keys, count, i, i + (l + t) * d) > d_min) l += t;
    }
    const int j = i + l * d;
    const int d_node = delta(keys, count, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, count, i, i + (s + t) * d) > d_node) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const bool left_leaf = first == gamma, right_leaf = last == gamma + 1;
    a.child[2 * i] = left_leaf ? (k_leaf_flag | static_cast<uint32_t>(gamma)) : static_cast<uint32_t>(gamma);
    a.child[2 * i + 1] =
        right_leaf ? (k_leaf_flag | static_cast<uint32_t>(gamma + 1)) : static_cast<uint32_t>(gamma + 1);
    a.range[2 * i] = static_cast<uint32_t>(first);
    a.range[2 * i + 1] = static_cast<uint32_t>(last);
    uint32_t* parent_leaf = a.parent + (n - 1);
    if (left_leaf) parent_leaf[gamma] = static_cast<uint32_t>(i); else a.parent[gamma] = static_cast<uint32_t>(i);
    if (right_leaf) parent_leaf[gamma + 1] = static_cast<uint32_t>(i); else a.parent[gamma + 1] = static_cast<uint32_t>(i);
    if (i == 0) a.parent[0] = 0u;
}

__device__ __forceinline__ void child_box(const build_arrays& a, uint32_t ref, float4& lo, float4& hi) {
    if (ref & k_leaf_flag) {
        const uint32_t t = static_cast<uint32_t>(a.keys[ref & ~k_leaf_flag]);  // low word = triangle index
        lo = a.prim_lo[t];
        hi = a.prim_hi[t];
    } else {
        lo = __ldcg(a.box_lo + ref);  // written by another thread of this launch: bypass L1
        hi = __ldcg(a.box_hi + ref);
    }
}

__global__ void __launch_bounds__(k_threads)
fit_kernel(build_arrays a, uint32_t n) {
    const uint32_t k = blockIdx.x * k_threads + threadIdx.x;
    if (k >= n) return;
    uint32_t node = a.parent[(n - 1) + k];
    for (;;) {
        __threadfence();
        if (atomicAdd(a.visits + node, 1u) == 0u) return;  // the sibling subtree is not finished yet
        float4 l_lo, l_hi, r_lo, r_hi;
        child_box(a, a.child[2 * node], l_lo, l_hi);
        child_box(a, a.child[2 * node + 1], r_lo, r_hi);
        __stcg(a.box_lo + node, make_float4(fminf(l_lo.x, r_lo.x), fminf(l_lo.y, r_lo.y), fminf(l_lo.z, r_lo.z), 0.0f));
        __stcg(a.box_hi + node, make_float4(fmaxf(l_hi.x, r_hi.x), fmaxf(l_hi.y, r_hi.y), fmaxf(l_hi.z, r_hi.z), 0.0f));
        if (node == 0u) return;
        node = a.parent[node];
    }
}

/// mesh_scene.hpp:build_bvh: pad = extent * 1e-4 + 1e-5 on every axis.
__device__ __forceinline__ void write_node(float4* nodes, uint32_t at, float4 lo, float4 hi, uint32_t left_first,
                                           uint32_t count) {
    const float px = (hi.x - lo.x) * 1e-4f + 1e-5f, py = (hi.y - lo.y) * 1e-4f + 1e-5f,
                pz = (hi.z - lo.z) * 1e-4f + 1e-5f;
    nodes[2 * at] = make_float4(lo.x - px, lo.y - py, lo.z - pz, __uint_as_float(left_first));
    nodes[2 * at + 1] = make_float4(hi.x + px, hi.y + py, hi.z + pz, __uint_as_float(count));
}

/// An internal node becomes one leaf if it covers at most k_leaf_max triangles, none of them an outlier.
__device__ __forceinline__ bool collapses(const build_arrays& a, uint32_t node) {
    const uint32_t first = a.range[2 * node], size = a.range[2 * node + 1] - first + 1u;
    return size <= k_leaf_max && (a.keys[a.range[2 * node + 1]] & k_outlier_bit) == 0ull;
}

__device__ __forceinline__ void emit_child(const build_arrays& a, float4* nodes, uint32_t at, uint32_t ref) {
    float4 lo, hi;
    if (ref & k_leaf_flag) {
        const uint32_t k = ref & ~k_leaf_flag;
        const uint32_t t = static_cast<uint32_t>(a.keys[k]);
        write_node(nodes, at, a.prim_lo[t], a.prim_hi[t], k, 1u);
        return;
    }
    lo = a.box_lo[ref];
    hi = a.box_hi[ref];
    const uint32_t first = a.range[2 * ref], size = a.range[2 * ref + 1] - first + 1u;
    if (collapses(a, ref)) {
        write_node(nodes, at, lo, hi, first, size);  // the whole subtree as one leaf
    } else {
        write_node(nodes, at, lo, hi, 1u + 2u * ref, 0u);
    }
}

__global__ void __launch_bounds__(k_threads)
emit_kernel(build_arrays a, float4* nodes, uint32_t n) {
    const uint32_t i = blockIdx.x * k_threads + threadIdx.x;
    if (i >= n - 1) return;
    if (i == 0) write_node(nodes, 0u, a.box_lo[0], a.box_hi[0], 1u, 0u);
    emit_child(a, nodes, 1u + 2u * i, a.child[2 * i]);
    emit_child(a, nodes, 2u + 2u * i, a.child[2 * i + 1]);
}

/// Longest root-to-leaf path in nodes of the emitted tree, and the leaf order.
__global__ void __launch_bounds__(k_threads)
depth_and_order_kernel(build_arrays a, uint32_t* leaf_order, uint32_t* depth, uint32_t n) {
    const uint32_t k = blockIdx.x * k_threads + threadIdx.x;
    // Every lane of the warp reaches the reduction below (no early return): the mask is explicit
    // and lanes past the end contribute zero, so the depth that guards the traversal stack
    // (plugin.cu: mesh.depth > k_mesh_stack_entries) never depends on where the warp reconverges
    // after the walks of different length.
    uint32_t levels = 0u;
    if (k < n) {
        leaf_order[k] = static_cast<uint32_t>(a.keys[k]);
        uint32_t node = a.parent[(n - 1) + k];
        levels = 1u;  // the leaf node itself
        for (;;) {
            if (!collapses(a, node)) levels++;  // an interior node of the emitted tree
            if (node == 0u) break;
            node = a.parent[node];
        }
    }
    __syncwarp();
    levels = __reduce_max_sync(0xffffffffu, levels);
    if ((threadIdx.x & 31) == 0 && levels != 0u) atomicMax(depth, levels);
}

__global__ void __launch_bounds__(k_threads)
gather_kernel(const uint32_t* __restrict__ leaf_order, const float4* __restrict__ v0, const float4* __restrict__ e1,
              const float4* __restrict__ e2, const float4* __restrict__ shade, float4* __restrict__ out_v0,
              float4* __restrict__ out_e1, float4* __restrict__ out_e2, float4* __restrict__ out_shade, uint32_t n) {
    const uint32_t k = blockIdx.x * k_threads + threadIdx.x;
    if (k >= n) return;
    const uint32_t t = leaf_order[k];
    out_v0[k] = v0[t];
    float4 edge = e1[t];
    edge.w = __uint_as_float(t);  // original index: the tie rule of the closest hit
    out_e1[k] = edge;
    float4 other = e2[t];
    other.w = __uint_as_float(k | (__float_as_uint(v0[t].w) << k_mesh_slot_bits));  // what a hit is recorded as (mesh_scene.hpp)
    out_e2[k] = other;
    for (int r = 0; r < 3; r++) out_shade[3u * k + r] = shade[3u * t + r];
}

}  // namespace

cudaError_t build_lbvh(const lbvh_input& in, const lbvh_output& out, cudaStream_t stream, float* build_ms) {
    const uint32_t n = in.n_triangles;
    if (n <= k_leaf_max || n >= k_leaf_flag) return cudaErrorInvalidValue;

    // One scratch allocation, carved up (256-byte aligned pieces).
    size_t offset = 0;
    auto carve = [&](size_t bytes) {
        const size_t at = offset;
        offset += (bytes + 255) & ~static_cast<size_t>(255);
        return at;
    };
    const size_t o_prim_lo = carve(sizeof(float4) * n), o_prim_hi = carve(sizeof(float4) * n);
    const size_t o_prim_cz = carve(sizeof(float) * n), o_scene = carve(sizeof(uint32_t) * 12);
    const size_t o_keys_a = carve(sizeof(uint64_t) * n), o_keys_b = carve(sizeof(uint64_t) * n);
    const size_t o_child = carve(sizeof(uint32_t) * 2 * n), o_range = carve(sizeof(uint32_t) * 2 * n);
    const size_t o_parent = carve(sizeof(uint32_t) * 2 * n), o_visits = carve(sizeof(uint32_t) * n);
    const size_t o_box_lo = carve(sizeof(float4) * n), o_box_hi = carve(sizeof(float4) * n);
    size_t sort_bytes = 0;
    cudaError_t err = cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, static_cast<uint64_t*>(nullptr),
                                                     static_cast<uint64_t*>(nullptr), static_cast<int>(n), 0, 64, stream);
    if (err != cudaSuccess) return err;
    const size_t o_sort = carve(sort_bytes);

    char* scratch = nullptr;
    err = cudaMalloc(reinterpret_cast<void**>(&scratch), offset);
    if (err != cudaSuccess) return err;
    build_arrays a{};
    a.tri_v0 = in.tri_v0;
    a.tri_e1 = in.tri_e1;
    a.tri_e2 = in.tri_e2;
    a.prim_lo = reinterpret_cast<float4*>(scratch + o_prim_lo);
    a.prim_hi = reinterpret_cast<float4*>(scratch + o_prim_hi);
    a.prim_cz = reinterpret_cast<float*>(scratch + o_prim_cz);
    a.scene = reinterpret_cast<uint32_t*>(scratch + o_scene);
    auto* keys_unsorted = reinterpret_cast<uint64_t*>(scratch + o_keys_a);
    a.keys = reinterpret_cast<uint64_t*>(scratch + o_keys_b);
    a.child = reinterpret_cast<uint32_t*>(scratch + o_child);
    a.range = reinterpret_cast<uint32_t*>(scratch + o_range);
    a.parent = reinterpret_cast<uint32_t*>(scratch + o_parent);
    a.visits = reinterpret_cast<uint32_t*>(scratch + o_visits);
    a.box_lo = reinterpret_cast<float4*>(scratch + o_box_lo);
    a.box_hi = reinterpret_cast<float4*>(scratch + o_box_hi);

    auto finish = [&](cudaError_t e) {
        cudaFree(scratch);  // synchronises with the stream's work on it
        return e;
    };
    // ordered(+FLT_MAX) for the minima, ordered(-FLT_MAX) for the maxima
    const uint32_t scene_init[12] = {0xff7fffffu, 0xff7fffffu, 0xff7fffffu, 0x00800000u, 0x00800000u, 0x00800000u,
                                     0xff7fffffu, 0xff7fffffu, 0xff7fffffu, 0x00800000u, 0x00800000u, 0x00800000u};
    if ((err = cudaMemcpyAsync(a.scene, scene_init, sizeof scene_init, cudaMemcpyHostToDevice, stream)) != cudaSuccess ||
        (err = cudaMemsetAsync(a.visits, 0, sizeof(uint32_t) * n, stream)) != cudaSuccess ||
        (err = cudaMemsetAsync(out.depth, 0, sizeof(uint32_t), stream)) != cudaSuccess ||
        (err = cudaMemsetAsync(out.nodes, 0, sizeof(float4) * 2 * (2 * static_cast<size_t>(n) - 1), stream)) !=
            cudaSuccess) {
        return finish(err);
    }
    // Timed: the kernels of the build, not the allocation of its scratch space above.
    cudaEvent_t begin = nullptr, end = nullptr;
    if (build_ms != nullptr) {
        *build_ms = 0.0f;
        if (cudaEventCreate(&begin) != cudaSuccess || cudaEventCreate(&end) != cudaSuccess) {
            (void)cudaGetLastError();
            if (begin) cudaEventDestroy(begin);
            begin = end = nullptr;
        } else {
            cudaStreamSynchronize(stream);  // the memsets above are not part of the measurement either
            cudaEventRecord(begin, stream);
        }
    }
    prim_bounds_kernel<<<blocks_for(n), k_threads, 0, stream>>>(a, n);
    morton_kernel<<<blocks_for(n), k_threads, 0, stream>>>(a, keys_unsorted, n);
    err = cub::DeviceRadixSort::SortKeys(scratch + o_sort, sort_bytes, keys_unsorted, a.keys, static_cast<int>(n), 0, 64,
                                         stream);
    if (err != cudaSuccess) return finish(err);
    hierarchy_kernel<<<blocks_for(n - 1), k_threads, 0, stream>>>(a, n);
    fit_kernel<<<blocks_for(n), k_threads, 0, stream>>>(a, n);
    emit_kernel<<<blocks_for(n - 1), k_threads, 0, stream>>>(a, out.nodes, n);
    depth_and_order_kernel<<<blocks_for(n), k_threads, 0, stream>>>(a, out.leaf_order, out.depth, n);
    gather_kernel<<<blocks_for(n), k_threads, 0, stream>>>(out.leaf_order, in.tri_v0, in.tri_e1, in.tri_e2, in.tri_shade,
                                                           out.tri_v0, out.tri_e1, out.tri_e2, out.tri_shade, n);
    if (end != nullptr) cudaEventRecord(end, stream);
    if ((err = cudaGetLastError()) == cudaSuccess) err = cudaStreamSynchronize(stream);
    if (end != nullptr) {
        if (err == cudaSuccess) cudaEventElapsedTime(build_ms, begin, end);
        cudaEventDestroy(begin);
        cudaEventDestroy(end);
    }
    return finish(err);
}

}  // namespace Q::cuda
