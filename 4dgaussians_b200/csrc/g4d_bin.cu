// g4d_bin.cu -- tile binning as TWO hand-written launches, no library sort, no host round trip:
//
//   bin_sort_kernel  (cooperative, persistent, one 1024-thread CTA per SM, grid-wide syncs between phases)
//       1. LSD radix sort (4 x 8 bits, stable) of the VISIBLE Gaussians by the bits of their view-space depth; the first
//          pass compacts away the invisible ones while it scatters.  Ties keep Gaussian-index order.   -> perm[n_visible]
//       2. the depth-ordered list is cut into one chunk per CTA with equal numbers of tile instances (near Gaussians cover
//          many more tiles than far ones)                                                                  -> chunk_start
//       3. every chunk counts its instances per tile (shared-memory histogram)                            -> M[chunk][tile]
//       4. per tile: exclusive scan over the chunks, then an exclusive scan over the tiles             -> ranges[tile], R
//   bin_place_kernel (one CTA per chunk)
//       every warp walks its share of the chunk IN DEPTH ORDER and drops each (Gaussian, tile) instance at
//       tile_start[tile] + M[chunk][tile] + (instances of earlier warps of the chunk) + (its own running count): a stable
//       counting placement.  Because the Gaussians arrive depth-sorted, every tile's segment comes out depth-sorted: the
//       reference's sort of R 64-bit (tile | depth) keys (6 radix passes over 12 B x R) is replaced by ONE 4-byte write per
//       instance, and the sorted list / tile ranges are bit-identical to the reference's (tests: sorted ids, ranges, keys).
//
// The instance count R never has to visit the host: the placement clamps to the buffer capacity and the overflow is
// reported through the context (G4D_OPT_SYNC_MODE = 0); in the default exact mode the host reads R between the two
// launches only to size the buffer.
//
// Reference stage replaced: duplicateWithKeys + cub::DeviceRadixSort + identifyTileRanges of the CUDA rasterizer behind
// /root/reference/gaussian_renderer/__init__.py:120-128.
#include <cooperative_groups.h>

#include "g4d_internal.h"
#include "raster_cull.cuh"

namespace cg = cooperative_groups;

namespace g4d {

namespace {

constexpr int kBinThreads = 1024;
constexpr int kRadix = 256;
constexpr uint32_t kSortSmemBytes = 32u * kRadix * 4u;   // per-warp digit counters of one scatter round

// inclusive scan of v over the 1024 threads of the block; total = block sum.  s_w: 33 words of shared scratch.
__device__ __forceinline__ uint32_t block_scan_incl(uint32_t v, uint32_t* s_w, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();   // previous users of s_w are done
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_w[lane];
        uint32_t ws = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += y;
        }
        s_w[lane] = ws - w;
        if (lane == 31) s_w[32] = ws;
    }
    __syncthreads();
    total = s_w[32];
    return x + s_w[warp];
}

struct SortShared {
    uint32_t hist[kRadix];
    uint32_t part[8 * kRadix];   // [0,4): totals of a quarter of the CTAs per digit, [4,8): totals of the CTAs before mine
    uint32_t base[kRadix];       // running output position per digit for my slice
    uint32_t sw[33];
};

// One stable LSD pass.  FIRST: input = the N raw Gaussians (key = depth bits, value = index), invisible ones are dropped.
template <bool FIRST, bool LAST>
__device__ __forceinline__ uint32_t radix_pass(const BinSortArgs& a, cg::grid_group& grid, SortShared& s, uint32_t* wc, int shift,
                                               uint32_t count, const uint32_t* __restrict__ kin, const uint32_t* __restrict__ vin,
                                               uint32_t* __restrict__ kout, uint32_t* __restrict__ vout) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t G = gridDim.x, c = blockIdx.x;
    const uint32_t per = (count + G - 1) / G;
    const uint32_t lo = min(c * per, count), hi = min(lo + per, count);
    auto load = [&](uint32_t i, uint32_t& key, uint32_t& val) -> bool {
        if (FIRST) {
            if (a.tiles_touched[i] == 0) return false;
            key = __float_as_uint(a.rec2[i].y);   // depth > 0.2: sign bit clear, integer order = float order
            val = i;
            return true;
        }
        key = kin[i]; val = vin[i];
        return true;
    };
    // ---- (a) digit histogram of my slice
    if (tid < kRadix) s.hist[tid] = 0;
    __syncthreads();
    for (uint32_t i = lo + tid; i < hi; i += kBinThreads) {
        uint32_t key, val;
        if (load(i, key, val)) atomicAdd(&s.hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid < kRadix) a.H[c * kRadix + tid] = s.hist[tid];
    grid.sync();
    // ---- (b) my output base per digit = (all smaller digits of every CTA) + (same digit of the CTAs before me)
    {
        const uint32_t d = tid & 255u, q = tid >> 8;
        const uint32_t qs = (G + 3) / 4, c0 = q * qs, c1 = min(G, c0 + qs);
        uint32_t tot = 0, bef = 0;
        for (uint32_t cc = c0; cc < c1; ++cc) {
            const uint32_t v = __ldcg(a.H + cc * kRadix + d);
            tot += v;
            if (cc < c) bef += v;
        }
        s.part[q * kRadix + d] = tot;
        s.part[(4 + q) * kRadix + d] = bef;
    }
    __syncthreads();
    uint32_t tot = 0, bef = 0;
    if (tid < kRadix) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { tot += s.part[q * kRadix + tid]; bef += s.part[(4 + q) * kRadix + tid]; }
    }
    uint32_t total;
    const uint32_t incl = block_scan_incl(tot, s.sw, total);
    if (tid < kRadix) s.base[tid] = incl - tot + bef;
    __syncthreads();
    // ---- (c) stable scatter, 1024 elements per round: rank inside the warp by match, then across the warps by a scan
    for (uint32_t b0 = lo; b0 < hi; b0 += kBinThreads) {
        for (int j = tid; j < 32 * kRadix; j += kBinThreads) wc[j] = 0;
        __syncthreads();
        const uint32_t i = b0 + tid;
        uint32_t key = 0, val = 0;
        const bool valid = i < hi && load(i, key, val);
        const uint32_t digit = (key >> shift) & 255u;
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : (256u + (uint32_t)lane));
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) wc[warp * kRadix + digit] = __popc(peers);
        __syncthreads();
        if (tid < kRadix) {
            uint32_t run = s.base[tid];
#pragma unroll 8
            for (int w = 0; w < 32; ++w) {
                const uint32_t t = wc[w * kRadix + tid];
                wc[w * kRadix + tid] = run;
                run += t;
            }
            s.base[tid] = run;
        }
        __syncthreads();
        if (valid) {
            const uint32_t pos = wc[warp * kRadix + digit] + rank;
            if (!LAST) kout[pos] = key;
            vout[pos] = val;
        }
        __syncthreads();
    }
    grid.sync();
    return total;
}

// number of tile instances of the Gaussians of one chunk, per tile of the band [y0, y1): shared-memory histogram
__device__ __forceinline__ void count_chunk_band(const BinSortArgs& a, uint32_t cs, uint32_t ce, int y0, int y1, uint32_t* cnt) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t k0 = cs + warp * 32u; k0 < ce; k0 += kBinThreads) {
        const uint32_t k = k0 + lane;
        const bool valid = k < ce;
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
        uint2 rc = make_uint2(0u, 0u);
        if (valid) {
            const uint32_t gi = a.perm[k];
            rc = a.rect[gi];
            if (a.tight) { r0 = a.rec0[gi]; r1 = a.rec1[gi]; }
        }
        uint32_t todo = __ballot_sync(0xffffffffu, valid);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const TileJob j = bcast_job(r0, r1, rc, src, y0, y1);
            for (int b = 0; b < j.ntiles; b += 32) {
                const int t = b + lane;
                const int ty = t / j.w, tx = t - ty * j.w;
                if (t < j.ntiles && (!a.tight || tile_contributes(j.r0, j.r1, j.minx + tx, j.miny + ty)))
                    atomicAdd(&cnt[(j.miny + ty - y0) * a.grid_x + j.minx + tx], 1u);
            }
        }
    }
}

}  // namespace

__global__ void __launch_bounds__(kBinThreads, 1) bin_sort_kernel(BinSortArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) uint32_t dyn[];   // scatter: per-warp digit counters; count: tile histogram of a band
    __shared__ SortShared s;
    const int tid = threadIdx.x;
    const uint32_t G = gridDim.x, c = blockIdx.x;
    const uint32_t N = (uint32_t)a.n;

    // ---- 1. depth order of the visible Gaussians
    const uint32_t nvis = radix_pass<true, false>(a, grid, s, dyn, 0, N, nullptr, nullptr, a.kA, a.vA);
    radix_pass<false, false>(a, grid, s, dyn, 8, nvis, a.kA, a.vA, a.kB, a.vB);
    radix_pass<false, false>(a, grid, s, dyn, 16, nvis, a.kB, a.vB, a.kA, a.vA);
    radix_pass<false, true>(a, grid, s, dyn, 24, nvis, a.kA, a.vA, nullptr, a.perm);

    // ---- 2. chunks of (nearly) equal instance counts along the depth order
    const uint32_t per = (nvis + G - 1) / G;
    const uint32_t lo = min(c * per, nvis), hi = min(lo + per, nvis);
    {
        uint32_t sum = 0;
        for (uint32_t k = lo + tid; k < hi; k += kBinThreads) sum += a.tiles_touched[a.perm[k]];
        uint32_t total;
        block_scan_incl(sum, s.sw, total);
        if (tid == 0) a.S[c] = total;
        if (c == 0) {
            for (uint32_t j = tid; j <= G; j += kBinThreads) a.chunk_start[j] = j == 0 ? 0u : nvis;
        }
    }
    grid.sync();
    {
        uint32_t v = 0, before = 0;
        if ((uint32_t)tid < G) { v = __ldcg(a.S + tid); before = (uint32_t)tid < c ? v : 0u; }
        uint32_t est, my_excl;
        block_scan_incl(v, s.sw, est);
        block_scan_incl(before, s.sw, my_excl);
        const uint32_t T = max(1u, (est + G - 1) / G);
        uint32_t run = my_excl;
        for (uint32_t b0 = lo; b0 < hi; b0 += kBinThreads) {
            const uint32_t k = b0 + tid;
            const uint32_t tt = k < hi ? a.tiles_touched[a.perm[k]] : 0u;
            uint32_t tot;
            const uint32_t incl = block_scan_incl(tt, s.sw, tot);
            if (k < hi) {
                const uint32_t e0 = run + incl - tt, e1 = run + incl;      // exclusive / inclusive prefix of element k
                const uint32_t ja = e0 / T, jb = min(e1 / T, G - 1);       // element k + 1 opens chunks (ja, jb]
                for (uint32_t j = ja + 1; j <= jb; ++j) a.chunk_start[j] = k + 1;
            }
            run += tot;
        }
    }
    grid.sync();

    // ---- 3. instances per (chunk, tile)
    const uint32_t cs = __ldcg(a.chunk_start + c), ce = __ldcg(a.chunk_start + c + 1);
    for (int y0 = 0; y0 < a.grid_y; y0 += a.count_band_rows) {
        const int y1 = min(a.grid_y, y0 + a.count_band_rows);
        const int bn = (y1 - y0) * a.grid_x;
        for (int j = tid; j < bn; j += kBinThreads) dyn[j] = 0;
        __syncthreads();
        count_chunk_band(a, cs, ce, y0, y1, dyn);
        __syncthreads();
        uint32_t* row = a.M + (size_t)c * a.num_tiles + (size_t)y0 * a.grid_x;
        for (int j = tid; j < bn; j += kBinThreads) row[j] = dyn[j];
        __syncthreads();
    }
    grid.sync();

    // ---- 4. per tile: exclusive scan over the chunks (in place); then exclusive scan over the tiles
    const uint32_t tps = ((uint32_t)a.num_tiles + G - 1) / G;           // tiles per CTA
    const uint32_t t_lo = min(c * tps, (uint32_t)a.num_tiles), t_hi = min(t_lo + tps, (uint32_t)a.num_tiles);
    for (uint32_t t = t_lo + tid; t < t_hi; t += kBinThreads) {
        uint32_t run = 0;
        uint32_t* col = a.M + t;
        uint32_t j = 0;
        for (; j + 8 <= G; j += 8) {
            uint32_t v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(col + (size_t)(j + u) * a.num_tiles);
#pragma unroll
            for (int u = 0; u < 8; ++u) { col[(size_t)(j + u) * a.num_tiles] = run; run += v[u]; }
        }
        for (; j < G; ++j) { const uint32_t v = __ldcg(col + (size_t)j * a.num_tiles); col[(size_t)j * a.num_tiles] = run; run += v; }
        a.tile_total[t] = run;
    }
    grid.sync();
    {
        // start of my tile slice = sum of every tile before it
        uint32_t sum = 0;
        for (uint32_t t = tid; t < t_lo; t += kBinThreads) sum += __ldcg(a.tile_total + t);
        uint32_t before;
        block_scan_incl(sum, s.sw, before);
        uint32_t run = before;
        for (uint32_t b0 = t_lo; b0 < t_hi; b0 += kBinThreads) {
            const uint32_t t = b0 + tid;
            const uint32_t v = t < t_hi ? __ldcg(a.tile_total + t) : 0u;
            uint32_t tot;
            const uint32_t incl = block_scan_incl(v, s.sw, tot);
            if (t < t_hi) {
                const uint32_t start = run + incl - v;
                a.tile_start[t] = start;
                a.ranges[t] = make_uint2(min(start, a.capacity), min(start + v, a.capacity));
            }
            run += tot;
        }
        if (c == G - 1 && tid == 0) {     // the last slice ends at R (empty trailing slices all live in the last CTAs: run = R there too)
            a.ctl->n_visible = nvis;
            a.ctl->R = run;
            a.ctl->overflow = run > a.capacity ? 1u : 0u;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// stable counting placement: CTA = chunk, warp = contiguous share of the chunk, lanes = tiles of one Gaussian's rect
template <bool COUNT>
__device__ __forceinline__ void place_walk(const BinPlaceArgs& a, uint32_t ws, uint32_t we, int y0, int y1, uint32_t* row) {
    const int lane = threadIdx.x & 31;
    for (uint32_t k0 = ws; k0 < we; k0 += 32u) {
        const uint32_t k = k0 + lane;
        const bool valid = k < we;
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
        uint2 rc = make_uint2(0u, 0u);
        uint32_t gi = 0;
        if (valid) {
            gi = a.perm[k];
            rc = a.rect[gi];
            if (a.tight) { r0 = a.rec0[gi]; r1 = a.rec1[gi]; }
        }
        uint32_t todo = __ballot_sync(0xffffffffu, valid);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const TileJob j = bcast_job(r0, r1, rc, src, y0, y1);
            const uint32_t id = __shfl_sync(0xffffffffu, gi, src);
            for (int b = 0; b < j.ntiles; b += 32) {
                const int t = b + lane;
                const int ty = t / j.w, tx = t - ty * j.w;
                if (t < j.ntiles && (!a.tight || tile_contributes(j.r0, j.r1, j.minx + tx, j.miny + ty))) {
                    uint32_t* p = row + (j.miny + ty - y0) * a.grid_x + j.minx + tx;   // lanes hold distinct tiles
                    const uint32_t slot = *p;
                    *p = slot + 1;
                    if (!COUNT && slot < a.capacity) a.ids[slot] = id;
                }
                __syncwarp();   // the next step (possibly another Gaussian on the same tile) must see these counters
            }
        }
    }
}

__global__ void __launch_bounds__(512) bin_place_kernel(BinPlaceArgs a) {
    extern __shared__ __align__(16) uint32_t rows[];   // [warps][tiles of the band]
    const int tid = threadIdx.x, warp = tid >> 5, nw = blockDim.x >> 5;
    const uint32_t c = blockIdx.x;
    const uint32_t cs = a.chunk_start[c], ce = a.chunk_start[c + 1];
    if (cs >= ce) return;
    const uint32_t len = ce - cs, per = (len + nw - 1) / nw;
    const uint32_t ws = cs + min((uint32_t)warp * per, len), we = cs + min((uint32_t)(warp + 1) * per, len);
    for (int y0 = 0; y0 < a.grid_y; y0 += a.band_rows) {
        const int y1 = min(a.grid_y, y0 + a.band_rows);
        const int bn = (y1 - y0) * a.grid_x;
        for (int j = tid; j < nw * bn; j += blockDim.x) rows[j] = 0;
        __syncthreads();
        place_walk<true>(a, ws, we, y0, y1, rows + warp * bn);
        __syncthreads();
        const uint32_t* mrow = a.M + (size_t)c * a.num_tiles + (size_t)y0 * a.grid_x;
        const uint32_t* tstart = a.tile_start + (size_t)y0 * a.grid_x;
        for (int t = tid; t < bn; t += blockDim.x) {
            uint32_t run = mrow[t] + tstart[t];
            for (int w = 0; w < nw; ++w) {
                const uint32_t v = rows[w * bn + t];
                rows[w * bn + t] = run;
                run += v;
            }
        }
        __syncthreads();
        place_walk<false>(a, ws, we, y0, y1, rows + warp * bn);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------------
namespace {
constexpr size_t kCountSmemBudget = 160 * 1024;
constexpr size_t kPlaceSmemBudget = 200 * 1024;
size_t align256(size_t v) { return (v + 255) / 256 * 256; }
}  // namespace

size_t bin_aux_bytes(int64_t n, int num_tiles, int sm_count) {
    const size_t N = (size_t)(n > 0 ? n : 1), G = (size_t)sm_count;
    return 4 * align256(N * 4) + align256(G * kRadix * 4) + align256(G * 4) + align256((G + 1) * 4) +
           align256(G * (size_t)num_tiles * 4) + 2 * align256((size_t)num_tiles * 4) + align256(sizeof(BinCtl));
}

cudaError_t launch_bin_sort(int64_t n, int grid_x, int grid_y, const GeomBuffers& g, void* aux, uint2* ranges, uint32_t capacity,
                            int tight, int sm_count, BinLayout* out, cudaStream_t st) {
    const int num_tiles = grid_x * grid_y;
    const size_t N = (size_t)(n > 0 ? n : 1), G = (size_t)sm_count;
    char* p = (char*)aux;
    auto take = [&](size_t bytes) { char* r = p; p += align256(bytes); return r; };
    BinSortArgs a{};
    a.n = n; a.rec2 = g.rec2; a.tiles_touched = g.tiles_touched; a.rect = g.rect; a.rec0 = g.rec0; a.rec1 = g.rec1;
    a.kA = (uint32_t*)take(N * 4); a.vA = (uint32_t*)take(N * 4); a.kB = (uint32_t*)take(N * 4); a.vB = (uint32_t*)take(N * 4);
    a.perm = g.perm;
    a.H = (uint32_t*)take(G * kRadix * 4); a.S = (uint32_t*)take(G * 4); a.chunk_start = (uint32_t*)take((G + 1) * 4);
    a.M = (uint32_t*)take(G * (size_t)num_tiles * 4);
    a.tile_total = (uint32_t*)take((size_t)num_tiles * 4); a.tile_start = (uint32_t*)take((size_t)num_tiles * 4);
    a.ctl = (BinCtl*)take(sizeof(BinCtl));
    a.ranges = ranges; a.grid_x = grid_x; a.grid_y = grid_y; a.num_tiles = num_tiles; a.capacity = capacity; a.tight = tight;
    int rows = (int)(kCountSmemBudget / ((size_t)grid_x * 4));
    if (rows < 1) return cudaErrorInvalidValue;
    a.count_band_rows = rows < grid_y ? rows : grid_y;
    size_t smem = (size_t)a.count_band_rows * grid_x * 4;
    if (smem < kSortSmemBytes) smem = kSortSmemBytes;
    out->chunk_start = a.chunk_start; out->M = a.M; out->tile_start = a.tile_start; out->ctl = a.ctl; out->chunks = sm_count;
    cudaError_t e = cudaFuncSetAttribute(bin_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    void* args[] = {&a};
    return cudaLaunchCooperativeKernel((const void*)bin_sort_kernel, dim3((unsigned)sm_count), dim3(kBinThreads), args, smem, st);
}

cudaError_t launch_bin_place(int grid_x, int grid_y, const GeomBuffers& g, const BinLayout& lay, uint32_t* ids, uint32_t capacity,
                             int tight, cudaStream_t st) {
    const int num_tiles = grid_x * grid_y;
    BinPlaceArgs a{};
    a.perm = g.perm; a.rect = g.rect; a.rec0 = g.rec0; a.rec1 = g.rec1; a.chunk_start = lay.chunk_start; a.M = lay.M;
    a.tile_start = lay.tile_start; a.ids = ids; a.capacity = capacity; a.grid_x = grid_x; a.grid_y = grid_y;
    a.num_tiles = num_tiles; a.tight = tight;
    int warps;
    if ((size_t)num_tiles * 4 * 4 <= kPlaceSmemBudget) {
        warps = (int)(kPlaceSmemBudget / ((size_t)num_tiles * 4));
        if (warps > 16) warps = 16;
        a.band_rows = grid_y;
    } else {
        warps = 4;
        a.band_rows = (int)(kPlaceSmemBudget / ((size_t)warps * grid_x * 4));
        if (a.band_rows < 1) return cudaErrorInvalidValue;
        if (a.band_rows > grid_y) a.band_rows = grid_y;
    }
    const size_t smem = (size_t)warps * a.band_rows * grid_x * 4;
    cudaError_t e = cudaFuncSetAttribute(bin_place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    bin_place_kernel<<<lay.chunks, warps * 32, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace g4d
