// radix_sort.cu — K3: stable LSD radix sort of the 64-bit (morton << 32 | prim) keys.
//
// Hand-written for sm_100a (no CUB). One pass per 8-bit digit:
//   hist    : every CTA histograms its 4096-key tile into shared memory, stores the 256 counts
//             digit-major ([digit][tile]) so a flat exclusive scan yields the scatter bases
//   scan    : one CTA per digit scans its row of tile counts; row totals go to digit_total[]
//   scatter : every CTA re-reads its tile warp-contiguously, ranks keys stably with
//             __match_any_sync + popc inside a warp and a cross-warp prefix in shared memory,
//             then writes them to their final position of this pass
// The builder only sorts the 30 Morton bits [32,62): the primitive id in the low word starts
// out ascending, and a stable sort keeps it ascending inside equal Morton codes, which is the
// duplicate tie-break the LBVH needs (the Cornell asset has exact duplicate triangles).
#include "build.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 keys per CTA
constexpr int RS_WARPS = RS_THREADS / 32;

__global__ __launch_bounds__(RS_THREADS) void k_rs_hist(const uint64_t* __restrict__ keys, uint32_t n, int shift,
                                                        uint32_t mask, uint32_t* __restrict__ tile_hist,
                                                        uint32_t ntiles) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint32_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    tile_hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of row `digit` (ntiles counts) in place; total to digit_total[digit]
__global__ __launch_bounds__(RS_THREADS) void k_rs_scan_rows(uint32_t* __restrict__ tile_hist, uint32_t ntiles,
                                                             uint32_t* __restrict__ digit_total) {
    __shared__ uint32_t warp_sum[RS_WARPS];
    __shared__ uint32_t carry_s;
    uint32_t* row = tile_hist + (size_t)blockIdx.x * ntiles;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t b = 0; b < ntiles; b += RS_THREADS) {
        uint32_t i = b + threadIdx.x;
        uint32_t v = i < ntiles ? row[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[wid] = x;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < wid; ++w) woff += warp_sum[w];
        uint32_t carry = carry_s;
        if (i < ntiles) row[i] = carry + woff + x - v;
        __syncthreads();
        if (threadIdx.x == RS_THREADS - 1) carry_s = carry + woff + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) digit_total[blockIdx.x] = carry_s;
}

__global__ __launch_bounds__(RS_THREADS) void k_rs_scatter(const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                                                           uint32_t n, int shift, uint32_t mask,
                                                           const uint32_t* __restrict__ tile_hist, uint32_t ntiles,
                                                           const uint32_t* __restrict__ digit_total) {
    __shared__ uint32_t wcount[RS_WARPS][256];
    __shared__ uint32_t dbase[256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int w = 0; w < RS_WARPS; ++w) wcount[w][threadIdx.x] = 0;
    // exclusive scan of the 256 digit totals -> global base of each digit
    {
        uint32_t v = digit_total[threadIdx.x];
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        __shared__ uint32_t ws[RS_WARPS];
        if (lane == 31) ws[wid] = x;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < wid; ++w) woff += ws[w];
        dbase[threadIdx.x] = woff + x - v + tile_hist[(size_t)threadIdx.x * ntiles + blockIdx.x];
    }
    __syncthreads();

    // phase A: warp-contiguous, round by round; rank inside the warp's 512-key strip
    const uint32_t strip = blockIdx.x * RS_TILE + wid * (32 * RS_ITEMS);
    uint64_t key[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint32_t i = strip + r * 32 + lane;
        bool valid = i < n;
        key[r] = valid ? in[i] : ~0ull;
        uint32_t d = valid ? ((uint32_t)(key[r] >> shift) & mask) : 256u;  // 256: padding lanes group together
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t before = 0;
        if (valid) before = wcount[wid][d];
        __syncwarp();
        if (valid && (peers & lt) == 0) wcount[wid][d] = before + __popc(peers);  // group leader
        __syncwarp();
        rank[r] = before + __popc(peers & lt);
    }
    __syncthreads();
    // cross-warp exclusive prefix per digit, seeded with the digit's global base for this tile
    {
        uint32_t run = dbase[threadIdx.x];
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t t = wcount[w][threadIdx.x];
            wcount[w][threadIdx.x] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint32_t i = strip + r * 32 + lane;
        if (i < n) {
            uint32_t d = (uint32_t)(key[r] >> shift) & mask;
            out[wcount[wid][d] + rank[r]] = key[r];
        }
    }
}

}  // namespace

size_t radix_sort_u64_temp_bytes(uint32_t n) {
    size_t ntiles = (n + RS_TILE - 1) / RS_TILE;
    if (ntiles == 0) ntiles = 1;
    return (256 * ntiles + 256) * sizeof(uint32_t);
}

uint64_t* radix_sort_u64(uint64_t* keys, uint64_t* tmp, uint32_t n, int begin_bit, int end_bit, void* temp,
                         size_t temp_bytes, cudaStream_t st) {
    (void)temp_bytes;
    if (n == 0) return keys;
    const uint32_t ntiles = (n + RS_TILE - 1) / RS_TILE;
    uint32_t* tile_hist = static_cast<uint32_t*>(temp);
    uint32_t* digit_total = tile_hist + (size_t)256 * ntiles;
    uint64_t *src = keys, *dst = tmp;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        int bits = end_bit - shift < 8 ? end_bit - shift : 8;
        uint32_t mask = (1u << bits) - 1u;
        k_rs_hist<<<ntiles, RS_THREADS, 0, st>>>(src, n, shift, mask, tile_hist, ntiles);
        k_rs_scan_rows<<<256, RS_THREADS, 0, st>>>(tile_hist, ntiles, digit_total);
        k_rs_scatter<<<ntiles, RS_THREADS, 0, st>>>(src, dst, n, shift, mask, tile_hist, ntiles, digit_total);
        uint64_t* t = src; src = dst; dst = t;
    }
    return src;
}
