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
    float* dp_cost = nullptr;               // optimal-collapse costs c(node, 1..7)
    uint8_t* dp_dec = nullptr;              // and choices, 8 bytes per binary node (build.cu)
    bool optimal_collapse = true;           // BPT_OPT_BVH_OPTIMAL_COLLAPSE
    uint32_t* leaf_prim = nullptr;          // leaf position -> primitive: the sorted order, permuted inside the subtrees
                                            // the SAH stage rebuilds (sah.cu)
    uint32_t sah_max_leaves = 32;           // BPT_OPT_BVH_SAH_SUBTREE: 0 = plain LBVH
    // BVH8 (K6): one array of 64-byte records, nodes and triangle (instance) records interleaved
    Node8* recs = nullptr;
    uint32_t nodes_cap = 0, recs_cap = 0;
    uint32_t* wide_src = nullptr;    // node record -> binary node it expands
    uint32_t* rec_prim = nullptr;    // record -> primitive id, BPT_MISS for node records
    uint32_t* list[2] = {nullptr, nullptr};  // node records of the current / next level
    uint32_t* counters = nullptr;
    void* sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    // host copies, valid after bvh8_build
    uint32_t num_recs = 0, num_nodes = 0, num_leaf_slots = 0, depth = 0;
    float scene_lo[3] = {0, 0, 0}, scene_hi[3] = {0, 0, 0};
    // scene grid of the node origins: origin = fmaf(float(2^23 + c), grid_step, grid_bias), c < 2^BPT_GRID_BITS
    float grid_lo[3] = {0, 0, 0}, grid_step[3] = {1, 1, 1}, grid_bias[3] = {0, 0, 0};
};

cudaError_t bvh8_alloc(Bvh8& b, uint32_t n);
void bvh8_free(Bvh8& b);
void bvh8_launch_tri_bounds(Bvh8& b, const float* verts, const uint32_t* idx, cudaStream_t st);
void bvh8_launch_instance_bounds(Bvh8& b, const float* xforms, const float mesh_lo[3], const float mesh_hi[3],
                                 cudaStream_t st);
cudaError_t bvh8_build(Bvh8& b, cudaStream_t st);
// writes the triangle records at the positions the collapse reserved for them
void bvh8_launch_woop(const Bvh8& b, const float* verts, const uint32_t* idx, cudaStream_t st);

// two-level scenes: instance records (inverse transforms) at the instance-level leaf positions, and the merged
// record array [mesh | instances] (child pointers of the instance-level nodes rebased by rec_off)
void bvh8_launch_instance_records(const Bvh8& tlas, const float* inv_xforms, cudaStream_t st);
void bvh8_launch_append_recs(const Bvh8& tlas, uint32_t rec_off, Node8* dst, cudaStream_t st);

// sah.cu — SAH rebuild of the small subtrees of the LBVH (between K4 and K5)
void sah_launch_leaf_order(const uint64_t* keys, uint32_t n, uint32_t* leaf_prim, cudaStream_t st);
void sah_launch_rebuild(uint32_t n, uint32_t max_leaves, const float4* plo, const float4* phi, uint32_t* leaf_prim,
                        uint32_t* left, uint32_t* right, uint32_t* parent, uint32_t* first, uint32_t* last, uint32_t* roots,
                        uint32_t* nroots, cudaStream_t st);

// radix_sort.cu — stable LSD radix sort of 64-bit keys on bits [begin_bit, end_bit).
// Returns the buffer (keys or tmp) that holds the sorted result.
size_t radix_sort_u64_temp_bytes(uint32_t n);
uint64_t* radix_sort_u64(uint64_t* keys, uint64_t* tmp, uint32_t n, int begin_bit, int end_bit, void* temp,
                         size_t temp_bytes, cudaStream_t st);
