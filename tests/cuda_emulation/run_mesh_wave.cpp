// Runs the SOURCE of the wavefront form of the mesh path tracer (quasi_b200/csrc/mesh_wavefront.cu: ray generation,
// trace, shade, resolve) on the host under tests/cuda_emulation/cuda_on_host.hpp, on the backend's own scene and
// tree (mesh_scene.hpp), with the launch sequence of launch_mesh_wavefront, and writes the four accumulated layers
// as raw binary32.
//   usage: run_mesh_wave GRID WIDTH HEIGHT FRAMES WIDE OUT.bin BATCH
//     WIDE: 0 binary tree, 1 the 4-wide tree, 2 the binary tree in its quantized form (quantized_bvh.hpp); BATCH: samples per batch (the last batch may be shorter)
//   prints "rays N nodes N triangles N"
#define QUASI_HOST_EMULATION 1
#include "cuda_on_host.hpp"

#include "../../quasi_b200/csrc/mesh_wavefront.cu"

#include <cstdio>
#include <cstdlib>

using namespace Q::cuda;

int main(int argc, char** argv) {
    if (argc < 8) return 2;
    const uint32_t grid_cells = static_cast<uint32_t>(std::atoi(argv[1]));
    const uint32_t width = static_cast<uint32_t>(std::atoi(argv[2])), height = static_cast<uint32_t>(std::atoi(argv[3]));
    const uint32_t frames = static_cast<uint32_t>(std::atoi(argv[4]));
    const bool wide = std::atoi(argv[5]) == 1, quant = std::atoi(argv[5]) == 2;
    const uint32_t batch = static_cast<uint32_t>(std::atoi(argv[7]));
    if (batch == 0) return 2;

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
    mesh::quantized_tree grid;
    if (quant) {
        grid = mesh::quantize_pairs(s);
        if (grid.pairs.empty()) return 8;
        scene.pairs = reinterpret_cast<const uint4*>(grid.pairs.data());
        for (int a = 0; a < 3; a++) scene.grid_origin[a] = grid.origin[a], scene.grid_step[a] = grid.step[a];
    }
    for (uint32_t m = 0; m < mesh::k_material_count; m++) {
        for (int c = 0; c < 3; c++) {
            scene.materials[m * 8 + c] = s.materials[m].albedo[c];
            scene.materials[m * 8 + 4 + c] = s.materials[m].emission[c];
        }
    }

    const float position[3] = {0.0f, 1.0f, 3.5f}, target[3] = {0.0f, 1.0f, 0.0f}, up[3] = {0.0f, 1.0f, 0.0f};
    const size_t n = static_cast<size_t>(width) * height;
    std::vector<float4> layers[4];
    for (auto& l : layers) l.assign(n, make_float4(-3.0f, -3.0f, -3.0f, -3.0f));  // (the first sample must not read it)
    unsigned long long counters[3] = {0, 0, 0};

    const size_t paths = static_cast<size_t>((width + 7u) / 8u) * ((height + 3u) / 4u) * 32u * batch;  // queue entries
    const float4 junk = make_float4(-9.0f, -9.0f, -9.0f, -9.0f);
    std::vector<float4> arrays[12];
    for (auto& a : arrays) a.assign(paths, junk);
    std::vector<uint32_t> counts(2 * k_mesh_wave_counters, 0u);
    mesh_wave_buffers q{};
    for (int i = 0; i < 2; i++) {
        q.origin[i] = arrays[4 * i].data();
        q.direction[i] = arrays[4 * i + 1].data();
        q.throughput[i] = arrays[4 * i + 2].data();
        q.colour[i] = arrays[4 * i + 3].data();
    }
    std::vector<float2> hits(paths, make_float2(-9.0f, -9.0f));
    q.hit = hits.data();
    q.out_beauty = arrays[9].data();
    q.out_albedo_depth = arrays[10].data();
    q.out_normal = arrays[11].data();
    q.counts = counts.data();
    q.capacity = paths;

    for (uint32_t first = 0; first < frames; first += batch) {
        const uint32_t samples = frames - first < batch ? frames - first : batch;
        frame_params frame;
        prepare_frame(position, target, up, 40.0f, width, height, /*frame_index*/ first, /*sample_index*/ first, frame);
        frame.batch_samples = samples;
        frame.frame_stride = 1;
        std::fill(counts.begin(), counts.end(), 0u);
        emu::launch(mesh_wave_raygen_kernel, 2u, k_block_threads, 1u, frame, q, samples);
        for (uint32_t wave = 0; wave < scene.max_bounces; wave++) {
            const uint32_t in = wave & 1u;
            auto run = [&](auto kernel) {
                emu::launch_with_shared(kernel, 2u, k_block_threads, 1u, k_stack_bytes, scene, q.origin[in], q.direction[in],
                                        q.hit, q.counts + wave, q.counts + k_mesh_wave_counters + wave, counters);
            };
            wide ? run(mesh_wave_trace_kernel<true, false>)
                 : (quant ? run(mesh_wave_trace_kernel<false, true>) : run(mesh_wave_trace_kernel<false, false>));
            emu::launch(mesh_wave_shade_kernel, 3u, k_block_threads, 1u, scene, q, wave);
            if (wave == 0u) {  // the first-hit layers, as soon as they are final
                emu::launch(mesh_wave_resolve_kernel<false>, 2u, k_block_threads, 1u, frame, q, samples, layers[0].data(),
                            layers[1].data(), layers[2].data(), layers[3].data());
            }
        }
        emu::launch(mesh_wave_resolve_kernel<true>, 2u, k_block_threads, 1u, frame, q, samples, layers[0].data(),
                    layers[1].data(), layers[2].data(), layers[3].data());
    }
    FILE* out = std::fopen(argv[6], "wb");
    if (!out) return 5;
    for (auto& l : layers) std::fwrite(l.data(), sizeof(float4), n, out);
    std::fclose(out);
    std::printf("rays %llu nodes %llu triangles %llu\n", counters[0], counters[1], counters[2]);
#if defined(QUASI_EMU_STATS)
    const char* names[5] = {"refill", "-", "scheduler", "descent", "triangle"};
    for (int p = 0; p < 5; p++) {
        const double passes = static_cast<double>(emu::g_phase_passes[p]), lanes = static_cast<double>(emu::g_phase_lanes[p]);
        std::printf("phase %-9s passes_per_ray %.3f lanes_per_pass %.2f\n", names[p], passes / static_cast<double>(counters[0]),
                    passes > 0 ? lanes / passes : 0.0);
    }
#endif
    return 0;
}
