// Library-backed binning primitives: the prefix sum over tiles-touched and the stable radix
// sort of (tile | depth) keys.  The reference calls the same two CUB algorithms
// (RAST/cuda_rasterizer/rasterizer_impl.cu:278 DeviceScan::InclusiveSum,
//  rasterizer_impl.cu:304-309 DeviceRadixSort::SortPairs); stability of the radix sort is what
// gives equal-depth Gaussians their index-order tie-break.
#include <cub/cub.cuh>
#include "raster_kernels.cuh"

namespace agr {

struct U32ToU64 {
    __host__ __device__ __forceinline__ uint64_t operator()(uint32_t v) const { return (uint64_t)v; }
};
using WideIter = cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*>;

// 64-bit running sum: V*P*tiles can exceed 2^32 for degenerate (screen-filling) inputs and must be
// reported as a capacity error, not wrap around.
size_t scan_temp_bytes(size_t n) {
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, bytes, WideIter((const uint32_t*)nullptr, U32ToU64()), (uint64_t*)nullptr, (int)n);
    return bytes;
}

size_t sort_temp_bytes(size_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return bytes;
}

cudaError_t inclusive_scan_u32(void* tmp, size_t tmp_bytes, const uint32_t* in, uint64_t* out, size_t n, cudaStream_t s) {
    return cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, WideIter(in, U32ToU64()), out, (int)n, s);
}

cudaError_t sort_pairs_u64_u32(void* tmp, size_t tmp_bytes, const uint64_t* kin, uint64_t* kout,
                               const uint32_t* vin, uint32_t* vout, size_t n, int end_bit, cudaStream_t s) {
    return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kin, kout, vin, vout, (int)n, 0, end_bit, s);
}

// ------------------------------------------------------------------ workspace carving
template <typename T>
static void take(char*& cur, T*& ptr, size_t count) {
    size_t off = align_up((size_t)cur, 128);
    ptr = reinterpret_cast<T*>(off);
    cur = reinterpret_cast<char*>(ptr + count);
}

GeomWs carve_geom(void* base, size_t P, size_t V, size_t M) {
    GeomWs w{};
    char* cur = static_cast<char*>(base);
    const size_t n = P * V;
    take(cur, w.rec, n);
    take(cur, w.tiles, n);
    take(cur, w.offsets, n);
    if (M > 0) {
        take(cur, w.rgb, n * 3);
        take(cur, w.clamped, n * 3);
    }
    w.scan_tmp_bytes = scan_temp_bytes(n);
    char* tmp; take(cur, tmp, w.scan_tmp_bytes);
    w.scan_tmp = tmp;
    w.total = (size_t)(cur - static_cast<char*>(base)) + 128;
    return w;
}

ImageWs carve_image(void* base, size_t V, size_t W, size_t H) {
    ImageWs w{};
    char* cur = static_cast<char*>(base);
    const size_t tiles = ((W + AGR_TILE_X - 1) / AGR_TILE_X) * ((H + AGR_TILE_Y - 1) / AGR_TILE_Y);
    take(cur, w.ranges, V * tiles);
    take(cur, w.tile_last, V * tiles);
    take(cur, w.n_contrib, V * W * H);
    take(cur, w.status, 2);
    w.total = (size_t)(cur - static_cast<char*>(base)) + 128;
    return w;
}

BinWs carve_binning(void* base, size_t capacity) {
    BinWs w{};
    char* cur = static_cast<char*>(base);
    take(cur, w.keys_in, capacity);
    take(cur, w.keys_out, capacity);
    take(cur, w.vals_in, capacity);
    take(cur, w.vals_out, capacity);
    take(cur, w.stream, capacity);
    w.sort_tmp_bytes = sort_temp_bytes(capacity);
    char* tmp; take(cur, tmp, w.sort_tmp_bytes);
    w.sort_tmp = tmp;
    w.total = (size_t)(cur - static_cast<char*>(base)) + 128;
    return w;
}

}  // namespace agr
