// g4d_raster.cu -- tile binning (scan, key emission, radix sort, tile ranges) and the per-tile front-to-back
// alpha compositing forward / back-to-front backward.   SURVEY.md Appendix A.2-A.4.
// Reference stage replaced: the CUDA rasterizer behind /root/reference/gaussian_renderer/__init__.py:120-128.
#include <cstdlib>

#include <cub/cub.cuh>

#include "g4d_internal.h"

namespace g4d {

// ------------------------------------------------------------------------------------------------------
size_t scan_temp_bytes(int64_t n) {
    size_t b = 0;
    cub::DeviceScan::InclusiveSum(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return b;
}
size_t sort_temp_bytes(int64_t r) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)r);
    return b;
}
cudaError_t launch_scan(const uint32_t* in, uint32_t* out, int64_t n, void* temp, size_t temp_bytes, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    return cub::DeviceScan::InclusiveSum(temp, temp_bytes, in, out, (int)n, st);
}

// ---- depth order -------------------------------------------------------------------------------------------------
// The reference sorts R (tile | depth) 64-bit keys on 32 + log2(tiles) bits (A.2).  A stable sort of the N Gaussians by
// depth bits followed by key emission IN THAT ORDER and a stable sort of the R keys on the tile bits alone gives the
// identical sequence (ties: depth-equal entries keep Gaussian-index order in both) with 2 instead of 6 passes over R.
struct PermutedCount {
    const uint32_t* tt; const uint32_t* perm;
    __host__ __device__ uint32_t operator()(int k) const { return tt[perm[k]]; }
};
__global__ void __launch_bounds__(256) depth_keys_kernel(int64_t n, GeomBuffers g) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    g.dkeys[i] = g.tiles_touched[i] ? __float_as_uint(g.rec2[i].y) : 0xFFFFFFFFu;   // depth > 0.2: sign bit clear, order-preserving
    g.dkeys[2 * n + i] = (uint32_t)i;
}
size_t depth_order_temp_bytes(int64_t n) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)n);
    cub::CountingInputIterator<int> cnt(0);
    cub::TransformInputIterator<uint32_t, PermutedCount, cub::CountingInputIterator<int>> it(cnt, PermutedCount{nullptr, nullptr});
    cub::DeviceScan::InclusiveSum(nullptr, b, it, (uint32_t*)nullptr, (int)n);
    return a > b ? a : b;
}
cudaError_t launch_depth_order(int64_t n, GeomBuffers g, void* temp, size_t temp_bytes, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    depth_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, g);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, g.dkeys, g.dkeys + n, g.dkeys + 2 * n, g.perm, (int)n, 0, 32, st);
    if (e != cudaSuccess) return e;
    cub::CountingInputIterator<int> cnt(0);
    cub::TransformInputIterator<uint32_t, PermutedCount, cub::CountingInputIterator<int>> it(cnt, PermutedCount{g.tiles_touched, g.perm});
    return cub::DeviceScan::InclusiveSum(temp, temp_bytes, it, g.offsets, (int)n, st);
}

// ---- exact-image tile culling (G4D_OPT_TIGHT_CULL) -------------------------------------------------------------
// A (Gaussian, tile) pair can be dropped without changing a single pixel when even the best-placed point of the
// tile's pixel rectangle has alpha = opacity * exp(-q/2) < 1/255 (the blend stage skips such contributions, A.3).
// q is a convex quadratic, so its minimum over the rectangle is 0 (centre inside) or lies on one of the 4 edges.
// The 1e-4 margin covers the different rounding of the per-pixel evaluation in the blend kernel.
G4D_D float edge_min(float a, float b, float c, float fixed, float lo, float hi) {
    // min over t in [lo,hi] of a*fixed^2 + 2*b*fixed*t + c*t^2
    float t = -b * fixed / c;
    t = fminf(fmaxf(t, lo), hi);
    return a * fixed * fixed + 2.f * b * fixed * t + c * t * t;
}
// can the Gaussian reach alpha >= 1/255 anywhere in the pixel rectangle [x0, x1] x [y0, y1] (inclusive pixel centres)?
G4D_D bool rect_contributes(float4 r0, float4 r1, float x0, float x1, float y0, float y1) {
    const float A = r0.z, B = r0.w, C = r1.x, op = r1.y;
    const float dx0 = r0.x - x1, dx1 = r0.x - x0;
    const float dy0 = r0.y - y1, dy1 = r0.y - y0;
    float qmin;
    if (dx0 <= 0.f && dx1 >= 0.f && dy0 <= 0.f && dy1 >= 0.f) qmin = 0.f;
    else {
        qmin = fminf(fminf(edge_min(A, B, C, dx0, dy0, dy1), edge_min(A, B, C, dx1, dy0, dy1)),
                     fminf(edge_min(C, B, A, dy0, dx0, dx1), edge_min(C, B, A, dy1, dx0, dx1)));
        qmin = fmaxf(qmin, 0.f);
    }
    return op * __expf(-0.5f * qmin) * 1.0001f >= kAlphaMin;
}
G4D_D bool tile_contributes(float4 r0, float4 r1, int tx, int ty) {
    const float A = r0.z, B = r0.w, C = r1.x, op = r1.y;
    const float dx0 = r0.x - (float)(tx * kTile + kTile - 1), dx1 = r0.x - (float)(tx * kTile);
    const float dy0 = r0.y - (float)(ty * kTile + kTile - 1), dy1 = r0.y - (float)(ty * kTile);
    float qmin;
    if (dx0 <= 0.f && dx1 >= 0.f && dy0 <= 0.f && dy1 >= 0.f) qmin = 0.f;
    else {
        qmin = fminf(fminf(edge_min(A, B, C, dx0, dy0, dy1), edge_min(A, B, C, dx1, dy0, dy1)),
                     fminf(edge_min(C, B, A, dy0, dx0, dx1), edge_min(C, B, A, dy1, dx0, dx1)));
        qmin = fmaxf(qmin, 0.f);
    }
    return op * __expf(-0.5f * qmin) * 1.0001f >= kAlphaMin;
}

// tight mode: tiles_touched := number of tiles of the rect that can contribute (same predicate as emit_keys_kernel,
// same translation unit, hence bit-identical decisions).
// Warp-cooperative: a warp owns 32 consecutive Gaussians; for each visible one (broadcast by shuffle) the 32 lanes test
// 32 tiles of its rect at a time -- a thread-per-Gaussian loop serialises on the largest rect of the warp.
struct TileJob { float4 r0, r1; int minx, miny, w, ntiles; };
G4D_D TileJob bcast_job(const float4& r0, const float4& r1, uint2 rc, int src) {
    TileJob j;
    j.r0.x = __shfl_sync(0xffffffffu, r0.x, src); j.r0.y = __shfl_sync(0xffffffffu, r0.y, src);
    j.r0.z = __shfl_sync(0xffffffffu, r0.z, src); j.r0.w = __shfl_sync(0xffffffffu, r0.w, src);
    j.r1.x = __shfl_sync(0xffffffffu, r1.x, src); j.r1.y = __shfl_sync(0xffffffffu, r1.y, src);
    j.r1.z = 0.f; j.r1.w = 0.f;
    const uint32_t rx = __shfl_sync(0xffffffffu, rc.x, src), ry = __shfl_sync(0xffffffffu, rc.y, src);
    j.minx = rx & 0xFFFF; j.miny = rx >> 16;
    const int maxx = ry & 0xFFFF, maxy = ry >> 16;
    j.w = maxx - j.minx;
    j.ntiles = j.w * (maxy - j.miny);
    return j;
}

__global__ void __launch_bounds__(256) cull_count_kernel(int64_t n, GeomBuffers g) {
    const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool vis = gi < n && g.tiles_touched[gi] != 0;
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
    uint2 rc = make_uint2(0u, 0u);
    if (vis) { r0 = g.rec0[gi]; r1 = g.rec1[gi]; rc = g.rect[gi]; }
    uint32_t todo = __ballot_sync(0xffffffffu, vis);
    uint32_t mine = 0;
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const TileJob j = bcast_job(r0, r1, rc, src);
        uint32_t cnt = 0;
        for (int base = 0; base < j.ntiles; base += 32) {
            const int t = base + lane;
            const int ty = t / j.w, tx = t - ty * j.w;
            const bool c = t < j.ntiles && tile_contributes(j.r0, j.r1, j.minx + tx, j.miny + ty);
            cnt += __popc(__ballot_sync(0xffffffffu, c));
        }
        if (lane == src) mine = cnt;
    }
    if (vis) g.tiles_touched[gi] = mine;
}

cudaError_t launch_cull_count(int64_t n, GeomBuffers g, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    cull_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, g);
    return cudaGetLastError();
}

// A.2: one (tile | depth bits) key and the Gaussian index per touched tile, at consecutive slots from offsets[i-1],
// tiles in row-major order of the rect (warp-cooperative like cull_count_kernel)
__global__ void __launch_bounds__(256) emit_keys_kernel(const CameraDev* __restrict__ cam, int64_t n, GeomBuffers g,
                                                        BinBuffers b, int64_t capacity, int tight) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // position in depth order
    const int lane = threadIdx.x & 31;
    const uint32_t gi = k < n ? g.perm[k] : 0u;
    const bool vis = k < n && g.tiles_touched[gi] != 0;
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
    uint2 rc = make_uint2(0u, 0u);
    uint32_t dbits = 0, off0 = 0;
    if (vis) {
        rc = g.rect[gi];
        dbits = __float_as_uint(g.rec2[gi].y);
        off0 = k == 0 ? 0u : g.offsets[k - 1];
        if (tight) { r0 = g.rec0[gi]; r1 = g.rec1[gi]; }
    }
    const int gx = cam->grid_x;
    uint32_t todo = __ballot_sync(0xffffffffu, vis);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const TileJob j = bcast_job(r0, r1, rc, src);
        const uint32_t db = __shfl_sync(0xffffffffu, dbits, src);
        int64_t off = (int64_t)__shfl_sync(0xffffffffu, off0, src);
        const uint32_t id = __shfl_sync(0xffffffffu, gi, src);
        for (int base = 0; base < j.ntiles; base += 32) {
            const int t = base + lane;
            const int ty = t / j.w, tx = t - ty * j.w;
            const int x = j.minx + tx, y = j.miny + ty;
            const bool c = t < j.ntiles && (!tight || tile_contributes(j.r0, j.r1, x, y));
            const uint32_t m = __ballot_sync(0xffffffffu, c);
            if (c) {
                const int64_t slot = off + __popc(m & ((1u << lane) - 1u));
                if (slot < capacity) {
                    b.keys_unsorted[slot] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | db;
                    b.ids_unsorted[slot] = id;
                }
            }
            off += __popc(m);
        }
    }
}

cudaError_t launch_emit_keys(const CameraDev* cam, int64_t n, GeomBuffers g, BinBuffers b, int64_t capacity, int tight,
                             cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    emit_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cam, n, g, b, capacity, tight);
    return cudaGetLastError();
}

cudaError_t launch_sort(BinBuffers b, int64_t r, int begin_bit, int end_bit, void* temp, size_t temp_bytes, cudaStream_t st) {
    if (r == 0) return cudaSuccess;
    return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, b.keys_unsorted, b.keys_sorted, b.ids_unsorted, b.ids_sorted,
                                           (int)r, begin_bit, end_bit, st);
}

__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint64_t* __restrict__ keys, int64_t r, uint2* ranges,
                                                          uint32_t num_tiles) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r) return;
    const uint32_t t = (uint32_t)(keys[i] >> 32);
    if (t >= num_tiles) return;   // padding slots of the capacity-bounded (no-sync) mode
    if (i == 0 || (uint32_t)(keys[i - 1] >> 32) != t) ranges[t].x = (uint32_t)i;
    if (i == r - 1 || (uint32_t)(keys[i + 1] >> 32) != t) ranges[t].y = (uint32_t)(i + 1);
}

cudaError_t launch_tile_ranges(BinBuffers b, int64_t r, int num_tiles, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(b.ranges, 0, sizeof(uint2) * (size_t)num_tiles, st);
    if (e != cudaSuccess || r == 0) return e;
    tile_ranges_kernel<<<(unsigned)((r + 255) / 256), 256, 0, st>>>(b.keys_sorted, r, b.ranges, (uint32_t)num_tiles);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------
// A.3 blend forward: one 16x16 CTA per tile, instances staged through shared memory in batches of 256.
template <int PPT>
__global__ void __launch_bounds__(kTilePixels / PPT)
blend_forward_kernel(const CameraDev* __restrict__ cam, GeomBuffers g, const uint32_t* __restrict__ ids,
                     const uint2* __restrict__ ranges, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                     float* __restrict__ out_color, float* __restrict__ out_depth, int warp_cull) {
    constexpr int NT = kTilePixels / PPT;
    __shared__ float4 s0[kTilePixels];
    __shared__ float4 s1[kTilePixels];
    __shared__ float2 s2[kTilePixels];
    const int H = cam->H, W = cam->W;
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const uint2 range = ranges[tile];
    const int rounds = (int)((range.y - range.x + kTilePixels - 1) / kTilePixels);
    int todo = (int)(range.y - range.x);
    // a warp owns a compact 16 x (2 PPT) pixel strip of the tile: lane l, pixel k -> (l & 15, 2 PPT warp + (l >> 4) + 2 k)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float pxf[PPT], pyf[PPT], T[PPT], C0[PPT], C1[PPT], C2[PPT], D[PPT];
    uint32_t last_contributor[PPT];
    bool done[PPT], inside[PPT];
    bool all_done = true;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int px = blockIdx.x * kTile + (lane & 15), py = blockIdx.y * kTile + 2 * PPT * warp + (lane >> 4) + 2 * k;
        inside[k] = px < W && py < H;
        pxf[k] = (float)px; pyf[k] = (float)py;
        T[k] = 1.f; C0[k] = C1[k] = C2[k] = D[k] = 0.f;
        last_contributor[k] = 0;
        done[k] = !inside[k];
        all_done = all_done && done[k];
    }
    // strip rectangle for the per-warp exact cull (same predicate as G4D_OPT_TIGHT_CULL, on 16 x 2 PPT pixels)
    const float sx0 = (float)(blockIdx.x * kTile), sx1 = sx0 + (float)(kTile - 1);
    const float sy0 = (float)(blockIdx.y * kTile + 2 * PPT * warp), sy1 = sy0 + (float)(2 * PPT - 1);
    for (int i = 0; i < rounds; ++i, todo -= kTilePixels) {
        if (__syncthreads_count(all_done) == NT) break;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int slot = threadIdx.x + k * NT;
            const int progress = i * kTilePixels + slot;
            if (range.x + progress < range.y) {
                const uint32_t id = ids[range.x + progress];
                s0[slot] = g.rec0[id];
                s1[slot] = g.rec1[id];
                s2[slot] = g.rec2[id];
            }
        }
        __syncthreads();
        const int cnt = min(kTilePixels, todo);
        for (int base = 0; base < cnt; base += 32) {
            if (__all_sync(0xffffffffu, all_done)) break;
            // lane l tests instance base + l against the warp's strip: instances that cannot reach alpha >= 1/255 on any
            // of its pixels are skipped by the whole warp (they would be skipped pixel by pixel anyway)
            const int jt = base + lane;
            bool hit = jt < cnt;
            if (hit && warp_cull) hit = rect_contributes(s0[jt], s1[jt], sx0, sx1, sy0, sy1);
            uint32_t mask = __ballot_sync(0xffffffffu, hit);
            while (mask) {
                const int j = base + __ffs(mask) - 1;
                mask &= mask - 1;
                const uint32_t contributor = (uint32_t)(i * kTilePixels + j + 1);
                const float4 a = s0[j];
                const float4 b = s1[j];
                const float2 c = s2[j];
#pragma unroll
                for (int k = 0; k < PPT; ++k) {
                    if (!done[k]) {
                        const float dx = a.x - pxf[k], dy = a.y - pyf[k];
                        const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
                        const float alpha = fminf(kAlphaMax, b.y * __expf(power));
                        if (power <= 0.f && alpha >= kAlphaMin) {
                            const float test_T = T[k] * (1.f - alpha);
                            if (test_T < kTransmittanceStop) done[k] = true;
                            else {
                                const float w = alpha * T[k];
                                C0[k] = fmaf(b.z, w, C0[k]); C1[k] = fmaf(b.w, w, C1[k]); C2[k] = fmaf(c.x, w, C2[k]);
                                D[k] = fmaf(c.y, w, D[k]);
                                T[k] = test_T;
                                last_contributor[k] = contributor;
                            }
                        }
                    }
                }
            }
            all_done = true;
#pragma unroll
            for (int k = 0; k < PPT; ++k) all_done = all_done && done[k];
        }
    }
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        if (!inside[k]) continue;
        const size_t pix = (size_t)pyf[k] * W + (size_t)pxf[k];
        const size_t hw = (size_t)H * W;
        final_T[pix] = T[k];
        n_contrib[pix] = last_contributor[k];
        out_color[pix] = fmaf(T[k], cam->bg[0], C0[k]);
        out_color[hw + pix] = fmaf(T[k], cam->bg[1], C1[k]);
        out_color[2 * hw + pix] = fmaf(T[k], cam->bg[2], C2[k]);
        out_depth[pix] = D[k];
    }
}

cudaError_t launch_blend_forward(const CameraDev* cam, int grid_x, int grid_y, GeomBuffers g, BinBuffers b, ImageBuffers im,
                                 float* out_color, float* out_depth, int warp_cull, cudaStream_t st) {
    if (grid_x * grid_y == 0) return cudaSuccess;
    static int ppt = []() { const char* e = getenv("G4D_BLEND_FWD_PPT"); const int v = e ? atoi(e) : 2; return (v == 1 || v == 4) ? v : 2; }();
#define G4D_LAUNCH_BF(P)                                                                                                  \
    blend_forward_kernel<P><<<dim3(grid_x, grid_y), kTilePixels / P, 0, st>>>(cam, g, b.ids_sorted, b.ranges, im.final_T, \
                                                                               im.n_contrib, out_color, out_depth, warp_cull)
    if (ppt == 1) G4D_LAUNCH_BF(1);
    else if (ppt == 4) G4D_LAUNCH_BF(4);
    else G4D_LAUNCH_BF(2);
#undef G4D_LAUNCH_BF
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------
// A.4 blend backward: same tiling, instances traversed back to front.  Per instance the 9 partial gradients
// of a warp's 32 pixels are reduced with shuffles (only when some lane contributes) and lane 0 issues the
// atomics, so global RED traffic is <= 8 x 9 per (tile, instance) instead of 256 x 9.
G4D_D float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// PPT pixels per thread (256 / PPT threads per tile): the per-Gaussian warp reduction -- half of the instruction count with
// one pixel per thread -- is shared by PPT x 32 pixels, and the staged record is read once per thread instead of per pixel.
template <int PPT>
__global__ void __launch_bounds__(kTilePixels / PPT)
blend_backward_kernel(const CameraDev* __restrict__ cam, GeomBuffers g, const uint32_t* __restrict__ ids,
                      const uint2* __restrict__ ranges, const float* __restrict__ final_T,
                      const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dcolor,
                      float* __restrict__ g_mean2D, float* __restrict__ g_conic, float* __restrict__ g_opacity,
                      float* __restrict__ g_rgb, int warp_cull) {
    constexpr int NT = kTilePixels / PPT;
    __shared__ float4 s0[kTilePixels];
    __shared__ float4 s1[kTilePixels];
    __shared__ float s2[kTilePixels];
    __shared__ uint32_t sid[kTilePixels];
    const int H = cam->H, W = cam->W;
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const uint2 range = ranges[tile];
    const int total = (int)(range.y - range.x);
    const int rounds = (total + kTilePixels - 1) / kTilePixels;
    const size_t hw = (size_t)H * W;
    // a warp owns a compact 16 x (2 PPT) pixel strip of the tile: lane l, pixel k -> (l & 15, 2 PPT warp + (l >> 4) + 2 k)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float sx0 = (float)(blockIdx.x * kTile), sx1 = sx0 + (float)(kTile - 1);
    const float sy0 = (float)(blockIdx.y * kTile + 2 * PPT * warp), sy1 = sy0 + (float)(2 * PPT - 1);
    float pxf[PPT], pyf[PPT], Tfin[PPT], T[PPT], dp0[PPT], dp1[PPT], dp2[PPT], bgdot[PPT];
    float ac0[PPT], ac1[PPT], ac2[PPT], lc0[PPT], lc1[PPT], lc2[PPT], last_alpha[PPT];
    int last[PPT];
    int my_last = 0;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int px = blockIdx.x * kTile + (lane & 15), py = blockIdx.y * kTile + 2 * PPT * warp + (lane >> 4) + 2 * k;
        const bool inside = px < W && py < H;
        const size_t pix = (size_t)py * W + px;
        pxf[k] = (float)px; pyf[k] = (float)py;
        Tfin[k] = inside ? final_T[pix] : 0.f;
        T[k] = Tfin[k];
        last[k] = inside ? (int)n_contrib[pix] : 0;
        dp0[k] = dp1[k] = dp2[k] = 0.f;
        if (inside) { dp0[k] = dL_dcolor[pix]; dp1[k] = dL_dcolor[hw + pix]; dp2[k] = dL_dcolor[2 * hw + pix]; }
        bgdot[k] = cam->bg[0] * dp0[k] + cam->bg[1] * dp1[k] + cam->bg[2] * dp2[k];
        ac0[k] = ac1[k] = ac2[k] = lc0[k] = lc1[k] = lc2[k] = last_alpha[k] = 0.f;
        my_last = max(my_last, last[k]);
    }
    const float ddx = 0.5f * (float)W, ddy = 0.5f * (float)H;
    // block-wide maximum of n_contrib: entries beyond it contribute to no pixel of the tile
    const int max_last = __reduce_max_sync(0xffffffffu, my_last);
    __shared__ int s_max[NT / 32];
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = max_last;
    __syncthreads();
    int tile_last = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) tile_last = max(tile_last, s_max[w]);
    // warp-level reduction plan: after the butterfly lane l holds value (l >> 2) of {mean2D.xy, conic.xyz, rgb}; lanes
    // 0,4,..,28 add one value each, lane 1 adds the opacity gradient -- one predicated RED instruction per Gaussian
    const bool hi16 = lane & 16, hi8 = lane & 8, hi4 = lane & 4;
    const int vidx = lane >> 2;
    const bool red_lane = (lane & 3) == 0 || lane == 1;
    float* red_base = lane == 1 ? g_opacity : vidx < 2 ? g_mean2D + vidx : vidx < 5 ? g_conic + (vidx - 2) : g_rgb + (vidx - 5);
    const uint32_t red_stride = lane == 1 ? 1u : vidx < 2 ? 2u : 3u;

    for (int i = 0; i < rounds; ++i) {
        // batch i covers list positions [total - (i+1)*256, total - i*256) traversed from the back
        const int hi = total - i * kTilePixels;           // exclusive upper position of this batch
        if (hi - kTilePixels >= tile_last) continue;      // whole batch behind every pixel's last contributor
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int slot = threadIdx.x + k * NT;
            const int pos = hi - 1 - slot;                // slot s stages list position hi-1-s
            if (pos >= 0) {
                const uint32_t id = ids[range.x + pos];
                sid[slot] = id;
                s0[slot] = g.rec0[id];
                s1[slot] = g.rec1[id];
                s2[slot] = g.rec2[id].x;
            }
        }
        __syncthreads();
        const int cnt = min(kTilePixels, hi);
        for (int base = 0; base < cnt; base += 32) {
          // lane l tests instance base + l: behind every pixel of this warp's strip, or unable to reach alpha >= 1/255
          // anywhere on the strip (same predicate as the exact tile cull) -> skipped by the whole warp
          const int jt = base + lane;
          bool hit = jt < cnt && (hi - 1 - jt) < max_last;
          if (hit && warp_cull) hit = rect_contributes(s0[jt], s1[jt], sx0, sx1, sy0, sy1);
          uint32_t mask = __ballot_sync(0xffffffffu, hit);
          while (mask) {
            const int j = base + __ffs(mask) - 1;
            mask &= mask - 1;
            const int lpos = hi - 1 - j;                  // position in the tile list (0-based)
            const float4 a = s0[j];
            const float4 b = s1[j];
            float Gk[PPT], alphak[PPT], dxk[PPT], dyk[PPT];
            bool validk[PPT];
            bool any_valid = false;
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                dxk[k] = a.x - pxf[k]; dyk[k] = a.y - pyf[k];
                const float power = -0.5f * (a.z * dxk[k] * dxk[k] + b.x * dyk[k] * dyk[k]) - a.w * dxk[k] * dyk[k];
                Gk[k] = __expf(power);
                alphak[k] = fminf(kAlphaMax, b.y * Gk[k]);
                validk[k] = lpos < last[k] && power <= 0.f && alphak[k] >= kAlphaMin;
                any_valid = any_valid || validk[k];
            }
            if (!__any_sync(0xffffffffu, any_valid)) continue;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // mean2D.xy, conic.xyz, rgb
            float v_op = 0.f;
            const float c0 = b.z, c1 = b.w, c2 = s2[j];
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                if (!validk[k]) continue;
                const float alpha = alphak[k], G = Gk[k], dx = dxk[k], dy = dyk[k];
                const float ra = __fdividef(1.f, 1.f - alpha);   // alpha <= 0.99: well inside the fast-division range
                T[k] = T[k] * ra;
                const float w = alpha * T[k];
                ac0[k] = last_alpha[k] * lc0[k] + (1.f - last_alpha[k]) * ac0[k]; lc0[k] = c0;
                ac1[k] = last_alpha[k] * lc1[k] + (1.f - last_alpha[k]) * ac1[k]; lc1[k] = c1;
                ac2[k] = last_alpha[k] * lc2[k] + (1.f - last_alpha[k]) * ac2[k]; lc2[k] = c2;
                float dL_dalpha = (c0 - ac0[k]) * dp0[k] + (c1 - ac1[k]) * dp1[k] + (c2 - ac2[k]) * dp2[k];
                v[5] += w * dp0[k]; v[6] += w * dp1[k]; v[7] += w * dp2[k];
                dL_dalpha *= T[k];
                last_alpha[k] = alpha;
                dL_dalpha += (-Tfin[k] * ra) * bgdot[k];
                const float dL_dG = b.y * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                v[0] += dL_dG * (-gdx * a.z - gdy * a.w) * ddx;
                v[1] += dL_dG * (-gdy * b.x - gdx * a.w) * ddy;
                v[2] += -0.5f * gdx * dx * dL_dG;
                v[3] += -gdx * dy * dL_dG;
                v[4] += -0.5f * gdy * dy * dL_dG;
                v_op += G * dL_dalpha;
            }
            // transposing butterfly: 8 values x 32 lanes -> value (lane >> 2) summed over the warp in 9 shuffles
            float u[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float send = hi16 ? v[k] : v[k + 4];
                u[k] = (hi16 ? v[k + 4] : v[k]) + __shfl_xor_sync(0xffffffffu, send, 16);
            }
            float x2[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float send = hi8 ? u[k] : u[k + 2];
                x2[k] = (hi8 ? u[k + 2] : u[k]) + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            float x = (hi4 ? x2[1] : x2[0]) + __shfl_xor_sync(0xffffffffu, hi4 ? x2[0] : x2[1], 4);
            x += __shfl_xor_sync(0xffffffffu, x, 2);
            x += __shfl_xor_sync(0xffffffffu, x, 1);
            v_op = warp_sum(v_op);
            if (red_lane) {
                const uint32_t id = sid[j];
                atomicAdd(red_base + red_stride * id, lane == 1 ? v_op : x);
            }
          }
        }
    }
}

cudaError_t launch_blend_backward(const CameraDev* cam, int grid_x, int grid_y, GeomBuffers g, BinBuffers b, ImageBuffers im,
                                  const float* dL_dcolor, float* g_mean2D, float* g_conic, float* g_opacity, float* g_rgb,
                                  int warp_cull, cudaStream_t st) {
    if (grid_x * grid_y == 0) return cudaSuccess;
    // pixels per thread: 2 measured best at C3 (tools/profile_step.py; G4D_BLEND_BWD_PPT overrides for experiments)
    static int ppt = []() { const char* e = getenv("G4D_BLEND_BWD_PPT"); const int v = e ? atoi(e) : 2; return (v == 1 || v == 4) ? v : 2; }();
#define G4D_LAUNCH_BB(P)                                                                                               \
    blend_backward_kernel<P><<<dim3(grid_x, grid_y), kTilePixels / P, 0, st>>>(cam, g, b.ids_sorted, b.ranges, im.final_T, \
                                                                                im.n_contrib, dL_dcolor, g_mean2D, g_conic, \
                                                                                g_opacity, g_rgb, warp_cull)
    if (ppt == 1) G4D_LAUNCH_BB(1);
    else if (ppt == 4) G4D_LAUNCH_BB(4);
    else G4D_LAUNCH_BB(2);
#undef G4D_LAUNCH_BB
    return cudaGetLastError();
}

}  // namespace g4d
