/// @file mesh_dump_tool.cpp
/// @brief Test helper: builds the backend's mesh scene + BVH on the CPU and dumps it.
///
/// Usage: mesh_dump_tool <grid> <seed> <out.bin> [task_limit]
/// out.bin = uint32 n_triangles, n_nodes, depth, 0; triangles (48 B each, original
/// order); nodes (32 B each); leaf_order (uint32 each). tests/test_mesh_cpu.py
/// compares it bit for bit with oracle/quasi_oracle_mesh.c's independent build.

#include <mesh_scene.hpp>

#include <cstdio>
#include <cstdlib>

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    Q::cuda::mesh::scene s;
    Q::cuda::mesh::generate_triangles(static_cast<uint32_t>(std::atoi(argv[1])),
                                      static_cast<uint32_t>(std::atoi(argv[2])), s);
    Q::cuda::mesh::build_bvh(s, argc > 4 ? static_cast<uint32_t>(std::atoi(argv[4])) : 32768u);
    std::FILE* f = std::fopen(argv[3], "wb");
    if (!f) return 1;
    const uint32_t hdr[4] = {static_cast<uint32_t>(s.triangles.size()), static_cast<uint32_t>(s.nodes.size()),
                             s.depth, 0};
    std::fwrite(hdr, sizeof hdr, 1, f);
    std::fwrite(s.triangles.data(), sizeof(Q::cuda::mesh::triangle), s.triangles.size(), f);
    std::fwrite(s.nodes.data(), sizeof(Q::cuda::mesh::node), s.nodes.size(), f);
    std::fwrite(s.leaf_order.data(), sizeof(uint32_t), s.leaf_order.size(), f);
    std::fclose(f);
    return 0;
}
