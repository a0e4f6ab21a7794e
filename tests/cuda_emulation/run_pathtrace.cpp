// Runs the SOURCE of the quad path-tracing kernel (quasi_b200/csrc/pathtrace.cu) on the host under
// tests/cuda_emulation/cuda_on_host.hpp and writes the four accumulated layers as raw binary32.
//   usage: run_pathtrace WIDTH HEIGHT FRAMES BATCH SCAN OUT.bin [RANKS]
//     BATCH  samples per launch (1 = one launch per render call; k > 1 = the batched form: every pixel gets k
//            samples in one launch, every tile is exported when its last pixel has had its last sample)
//     SCAN   "quads" (kernel unrolled for the demo scene), "boxes" (blocks as slab tests), "generic"
//     RANKS  G > 1: the multi-GPU form on one host. G shards are rendered one after the other (shard g: FRAMES
//            samples with RNG frame indices g, g + G, ...; BATCH must equal FRAMES: one batched launch each) with
//            the export and the fused reduce of batch_params: staging areas, arrival counters and the finished
//            frame are plain arrays here. OUT.bin is then the finished frame (the last shard's launch reduces
//            every unit, being the last to arrive), which the caller compares with the oracle's combine.
//   prints "rays N launches M"
#define QUASI_HOST_EMULATION 1
#include "cuda_on_host.hpp"

#include "../../quasi_b200/csrc/pathtrace.cu"

#include <cstdio>
#include <cstdlib>
#include <string>

using namespace Q::cuda;

int main(int argc, char** argv) {
    if (argc < 7) return 2;
    const uint32_t width = static_cast<uint32_t>(std::atoi(argv[1])), height = static_cast<uint32_t>(std::atoi(argv[2]));
    const uint32_t frames = static_cast<uint32_t>(std::atoi(argv[3])), batch = static_cast<uint32_t>(std::atoi(argv[4]));
    const std::string scan = argv[5];

    scene_desc desc = make_cornell_box();
    static scene_params scene;
    prepare_scene(desc, k_max_bounces, scene, scan == "boxes");
    // the camera of the headless hosts (quasi_b200/cuda_context.py:orbit_camera, host/src/main.cpp)
    const float position[3] = {0.0f, 1.0f, 3.5f}, target[3] = {0.0f, 1.0f, 0.0f}, up[3] = {0.0f, 1.0f, 0.0f};
    set_decision_margins(scene, position);

    const size_t n = static_cast<size_t>(width) * height;
    std::vector<float4> layers[4];
    for (auto& l : layers) l.assign(n, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    unsigned long long rays = 0;
    unsigned int tile_counters[2] = {0, 0};
    // batched launches export every finished tile: here into one buffer per layer (what the plugin's readback mirror is)
    std::vector<float4> mirror[4];
    for (auto& m : mirror) m.assign(n, make_float4(-1.0f, -1.0f, -1.0f, -1.0f));

    const uint32_t n_tiles = ((width + 7u) / 8u) * ((height + 3u) / 4u);
    const uint32_t ranks = argc > 7 ? static_cast<uint32_t>(std::atoi(argv[7])) : 1u;
    if (ranks > 1u) {
        if (batch != frames || frames < 2u || ranks > k_max_export_slices) return 7;
        // slices along rows of tiles (quasi_b200/multi_gpu.py:tile_row_slices)
        const uint32_t tiles_y = (height + 3u) / 4u, pairs = 4u * tiles_y;
        batch_params launch{};
        uint32_t longest = 0;
        for (uint32_t s = 0; s <= ranks; s++) {
            const uint32_t pair = (s * pairs + ranks - 1u) / ranks;  // ceil(s * pairs / ranks)
            launch.slice_lo[s] = s == ranks ? 4u * static_cast<uint32_t>(n) : ((pair / tiles_y) * height + (pair % tiles_y) * 4u) * width;
            if (s > 0) longest = std::max(longest, launch.slice_lo[s] - launch.slice_lo[s - 1]);
        }
        std::vector<std::vector<float4>> staging(ranks, std::vector<float4>(static_cast<size_t>(ranks) * longest,
                                                                             make_float4(-7.0f, -7.0f, -7.0f, -7.0f)));
        std::vector<std::vector<unsigned int>> counters(ranks, std::vector<unsigned int>(4u * n_tiles, 0u));
        std::vector<float4> finished(4u * n, make_float4(-1.0f, -1.0f, -1.0f, -1.0f));
        launch.export_slices = launch.reduce_ranks = ranks;
        launch.reduce_slot_stride = longest;
        launch.reduce_scale = 1.0f / static_cast<float>(ranks);
        for (uint32_t s = 0; s < ranks; s++) {
            launch.reduce_out[s] = finished.data() + launch.slice_lo[s];
            launch.reduce_staging[s] = staging[s].data();
            launch.reduce_counters[s] = counters[s].data();
        }
        const uint32_t needed_blocks = (n_tiles + k_warps_per_block - 1u) / k_warps_per_block;
        const uint32_t grid_size = needed_blocks < 3u ? needed_blocks : 3u;
        for (uint32_t g = 0; g < ranks; g++) {
            for (auto& l : layers) l.assign(n, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
            for (uint32_t s = 0; s < ranks; s++) launch.export_base[s] = staging[s].data() + static_cast<size_t>(g) * longest;
            frame_params frame;
            prepare_frame(position, target, up, 40.0f, width, height, /*frame_index*/ g, /*sample_index*/ 0, frame);
            frame.batch_samples = frames;
            frame.frame_stride = ranks;
            emu::launch_with_shared(pathtrace_accumulate_kernel<16, cornell_quads_layout, true, true>, grid_size, k_block_threads, 1u,
                                    k_wavefront_smem, scene, frame, launch, layers[0].data(), layers[1].data(), layers[2].data(),
                                    layers[3].data(), &rays, tile_counters);
        }
        FILE* out = std::fopen(argv[6], "wb");
        if (!out) return 5;
        std::fwrite(finished.data(), sizeof(float4), finished.size(), out);
        std::fclose(out);
        for (const auto& c : counters) {  // every unit has been completed, and its counter put back
            for (unsigned int v : c) {
                if (v != 0u) return 9;
            }
        }
        std::printf("rays %llu launches %u\n", rays, ranks);
        return 0;
    }
    const uint32_t needed = (n_tiles + k_warps_per_block - 1u) / k_warps_per_block;
    const uint32_t grid = needed < 3u ? needed : 3u;  // a persistent grid smaller than the work, as on the device
    uint32_t launches = 0;
    for (uint32_t first = 0; first < frames; first += batch) {
        frame_params frame;
        prepare_frame(position, target, up, 40.0f, width, height, /*frame_index*/ first, /*sample_index*/ first, frame);
        frame.batch_samples = std::min(batch, frames - first);
        frame.frame_stride = 1;
        const bool batched = frame.batch_samples > 1u;
        batch_params launch{};
        if (batched) {
            launch.export_slices = 4;
            for (int i = 0; i < 4; i++) {
                launch.export_base[i] = mirror[i].data();
                launch.slice_lo[i] = static_cast<uint32_t>(i * n);
            }
            launch.slice_lo[4] = static_cast<uint32_t>(4 * n);
        }
        auto run = [&](auto kernel) {
            emu::launch_with_shared(kernel, grid, k_block_threads, 1u, k_wavefront_smem, scene, frame, launch,
                                    layers[0].data(), layers[1].data(), layers[2].data(), layers[3].data(), &rays, tile_counters);
        };
        if (scan == "generic") {
            batched ? run(pathtrace_accumulate_kernel<32, dynamic_layout, true>) : run(pathtrace_accumulate_kernel<32, dynamic_layout, false>);
        } else if (scan == "boxes" && cornell_layout::matches(scene)) {
            batched ? run(pathtrace_accumulate_kernel<16, cornell_layout, true>) : run(pathtrace_accumulate_kernel<16, cornell_layout, false>);
        } else if (scan == "quads" && cornell_quads_layout::matches(scene)) {
            batched ? run(pathtrace_accumulate_kernel<16, cornell_quads_layout, true>)
                    : run(pathtrace_accumulate_kernel<16, cornell_quads_layout, false>);
        } else {
            std::fprintf(stderr, "scene does not match the layout of scan form %s\n", scan.c_str());
            return 3;
        }
        launches++;
        if (tile_counters[0] != 0u || tile_counters[1] != 0u) {
            std::fprintf(stderr, "work queue not reset after launch %u\n", launches);
            return 4;
        }
    }
    if (batch > 1u && frames % batch != 1u) {  // (a last launch of one sample exports nothing)
        for (int i = 0; i < 4; i++) {
            if (std::memcmp(mirror[i].data(), layers[i].data(), n * sizeof(float4)) != 0) {
                std::fprintf(stderr, "export of layer %d differs from the layer\n", i);
                return 6;
            }
        }
    }
    FILE* out = std::fopen(argv[6], "wb");
    if (!out) return 5;
    for (auto& l : layers) std::fwrite(l.data(), sizeof(float4), n, out);
    std::fclose(out);
    std::printf("rays %llu launches %u\n", rays, launches);
    return 0;
}
