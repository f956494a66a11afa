// build.cuh — host-side handles of the acceleration-structure build (build.cu, radix_sort.cu).
#pragma once
#include "common.cuh"

// One BVH8 over n primitives (triangles of the mesh, or instances of it).
struct Bvh8 {
    uint32_t n = 0;
    // primitive boxes (K1) and encoded scene bounds
    float4 *plo = nullptr, *phi = nullptr;
    uint32_t* bounds = nullptr;
    // LBVH (K2-K5); kept after the build for bpt_download_morton / bpt_download_lbvh
    uint64_t *keys = nullptr, *keys_tmp = nullptr;
    uint32_t *left = nullptr, *right = nullptr, *parent = nullptr, *first = nullptr, *last = nullptr, *arrive = nullptr;
    float4 *nlo = nullptr, *nhi = nullptr;  // 2n-1 boxes: internal nodes, then leaves (sorted order)
    // BVH8 (K6)
    Node8* nodes = nullptr;
    uint32_t nodes_cap = 0;
    uint32_t* wide_src = nullptr;
    uint32_t* prim_index = nullptr;  // leaf slot -> primitive id
    uint32_t* counters = nullptr;
    void* sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    // host copies, valid after bvh8_build
    uint32_t num_nodes = 0, num_leaf_slots = 0, depth = 0;
    float scene_lo[3] = {0, 0, 0}, scene_hi[3] = {0, 0, 0};
};

cudaError_t bvh8_alloc(Bvh8& b, uint32_t n);
void bvh8_free(Bvh8& b);
void bvh8_launch_tri_bounds(Bvh8& b, const float* verts, const uint32_t* idx, cudaStream_t st);
void bvh8_launch_instance_bounds(Bvh8& b, const float* xforms, const float mesh_lo[3], const float mesh_hi[3],
                                 cudaStream_t st);
cudaError_t bvh8_build(Bvh8& b, cudaStream_t st);
void bvh8_launch_woop(const Bvh8& b, const float* verts, const uint32_t* idx, WoopTri* out, cudaStream_t st);

// two-level scenes: instance leaf records (inverse transforms) and the merged node array [mesh | instances]
void bvh8_launch_instance_records(const Bvh8& tlas, const float* inv_xforms, WoopTri* out, cudaStream_t st);
void bvh8_launch_append_nodes(const Node8* src, uint32_t n, uint32_t node_off, uint32_t rec_off, Node8* dst, cudaStream_t st);

// radix_sort.cu — stable LSD radix sort of 64-bit keys on bits [begin_bit, end_bit).
// Returns the buffer (keys or tmp) that holds the sorted result.
size_t radix_sort_u64_temp_bytes(uint32_t n);
uint64_t* radix_sort_u64(uint64_t* keys, uint64_t* tmp, uint32_t n, int begin_bit, int end_bit, void* temp,
                         size_t temp_bytes, cudaStream_t st);
