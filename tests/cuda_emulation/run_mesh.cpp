// Runs the SOURCE of the mesh path-tracing kernel (quasi_b200/csrc/mesh_trace.cu) on the host under
// tests/cuda_emulation/cuda_on_host.hpp, on the backend's own scene and tree (mesh_scene.hpp), and
// writes the four accumulated layers as raw binary32.
//   usage: run_mesh GRID WIDTH HEIGHT FRAMES WIDE OUT.bin [BATCH]
//     WIDE: 0 binary tree, 1 the 4-wide tree; BATCH: samples per launch (default 1)
//   prints "rays N nodes N triangles N"
#define QUASI_HOST_EMULATION 1
#include "cuda_on_host.hpp"

#include "../../quasi_b200/csrc/mesh_trace.cu"

#include <cstdio>
#include <cstdlib>

using namespace Q::cuda;

int main(int argc, char** argv) {
    if (argc < 7) return 2;
    const uint32_t grid_cells = static_cast<uint32_t>(std::atoi(argv[1]));
    const uint32_t width = static_cast<uint32_t>(std::atoi(argv[2])), height = static_cast<uint32_t>(std::atoi(argv[3]));
    const uint32_t frames = static_cast<uint32_t>(std::atoi(argv[4]));
    const bool wide = std::atoi(argv[5]) != 0;

    mesh::scene s;
    mesh::generate_triangles(grid_cells, 1, s);
    mesh::build_bvh(s);
    std::vector<mesh::float4_t> packed[5];
    mesh::pack_for_device(s, packed[0], packed[1], packed[2], packed[3], packed[4]);
    mesh::wide_tree w;
    if (wide) w = mesh::collapse_wide(s, k_mesh_stack_entries - 1);

    static mesh_params scene{};
    static_assert(sizeof(mesh::float4_t) == sizeof(float4));
    scene.nodes = reinterpret_cast<const float4*>(wide ? w.nodes.data() : packed[0].data());
    scene.tri_v0 = reinterpret_cast<const float4*>(packed[1].data());
    scene.tri_e1 = reinterpret_cast<const float4*>(packed[2].data());
    scene.tri_e2 = reinterpret_cast<const float4*>(packed[3].data());
    scene.tri_shade = reinterpret_cast<const float4*>(packed[4].data());
    scene.n_nodes = wide ? w.node_count : static_cast<uint32_t>(s.nodes.size());
    scene.n_triangles = static_cast<uint32_t>(s.triangles.size());
    scene.max_bounces = 4;
    scene.wide = wide ? 1u : 0u;
    for (uint32_t m = 0; m < mesh::k_material_count; m++) {
        for (int c = 0; c < 3; c++) {
            scene.materials[m * 8 + c] = s.materials[m].albedo[c];
            scene.materials[m * 8 + 4 + c] = s.materials[m].emission[c];
        }
    }

    const float position[3] = {0.0f, 1.0f, 3.5f}, target[3] = {0.0f, 1.0f, 0.0f}, up[3] = {0.0f, 1.0f, 0.0f};
    const size_t n = static_cast<size_t>(width) * height;
    std::vector<float4> layers[4];
    for (auto& l : layers) l.assign(n, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    unsigned long long counters[3] = {0, 0, 0};
    unsigned int tile_counters[2] = {0, 0};
    const uint32_t n_tiles = ((width + 7u) / 8u) * ((height + 3u) / 4u);
    const uint32_t needed = (n_tiles + k_warps_per_block - 1u) / k_warps_per_block;
    const uint32_t grid = needed < 2u ? needed : 2u;
    for (uint32_t f = 0; f < frames; f++) {  // one launch per sample: the mesh kernel has no batched form
        frame_params frame;
        prepare_frame(position, target, up, 40.0f, width, height, f, f, frame);
        auto run = [&](auto kernel) {
            emu::launch_with_shared(kernel, grid, k_block_threads, 1u, k_stack_bytes, scene, frame, layers[0].data(),
                                    layers[1].data(), layers[2].data(), layers[3].data(), counters, tile_counters);
        };
        wide ? run(mesh_pathtrace_kernel<true>) : run(mesh_pathtrace_kernel<false>);
    }
    FILE* out = std::fopen(argv[6], "wb");
    if (!out) return 5;
    for (auto& l : layers) std::fwrite(l.data(), sizeof(float4), n, out);
    std::fclose(out);
    std::printf("rays %llu nodes %llu triangles %llu\n", counters[0], counters[1], counters[2]);
#if defined(QUASI_EMU_STATS)
    const char* names[5] = {"shading", "new_rays", "scheduler", "descent", "triangle"};
    for (int p = 0; p < 5; p++) {
        const double passes = static_cast<double>(emu::g_phase_passes[p]), lanes = static_cast<double>(emu::g_phase_lanes[p]);
        std::printf("phase %-9s passes_per_ray %.3f lanes_per_pass %.2f\n", names[p], passes / static_cast<double>(counters[0]),
                    passes > 0 ? lanes / passes : 0.0);
    }
#endif
    return 0;
}
