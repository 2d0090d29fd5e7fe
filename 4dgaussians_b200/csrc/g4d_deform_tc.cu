// g4d_deform_tc.cu -- tensor-core version of the fused deform + activate + project forward (net_width 128).
//
// The deformation MLP (scene/deformation.py:67-148: 93.6 kMAC per Gaussian at the DyNeRF config) runs on the 5th-gen
// tensor cores with hand-written tcgen05 PTX:
//   * a CTA owns a tile of 128 Gaussians = the 128 TMEM lanes; every activation matrix lives in TMEM and is the
//     A operand of the next GEMM (tcgen05.mma A-from-TMEM), so activations never touch shared memory or HBM;
//   * weights are the B operand: pre-packed once per parameter version into the canonical K-major smem image
//     (hi | lo TF32 parts) and streamed head by head with TMA bulk copies + mbarriers;
//   * fp32 accuracy through 3xTF32 (lo*hi + hi*lo + hi*hi, fp32 accumulate in TMEM): measured error below a plain
//     fp32 SGEMM's, which the 1e-4 image tolerance needs -- single-pass TF32 is 1000x worse;
//   * epilogues (bias, ReLU, sign bits for the backward, hi/lo split) are tcgen05.ld -> registers -> tcgen05.st, one
//     thread per TMEM lane, two warps per lane quadrant splitting the hidden columns;
//   * the HexPlane gather runs BEFORE this kernel at full occupancy (deform_features_kernel, feat [N][F] stays in L2):
//     inside a 1-CTA/SM persistent kernel four gather warps were latency-bound and slowed the epilogue warps;
//   * the tile's 128 result rows stay in registers of their owner threads, which finish the Gaussian
//     (activations, EWA projection, SH colour) exactly like the SIMT path (geom_finish.cuh).
//
// TMEM map (512 columns): [0,128) A1 hi, [128,256) A1 lo, [256,384) D0, [384,512) D1 = X.
//   D0 / D1: the layer-1 accumulators of consecutive heads ALTERNATE between the two, so that the GEMM of head k+1 runs on
//            the tensor pipe while the 256 epilogue threads drain head k (the SH head always gets D0; its layer-2 output D2
//            reuses D0[0,48));
//   X (= D1): feature operand of layer 0 (hi [384,384+F), lo [448,448+F)) at the start of a tile, the SH head's hidden half
//            (hi [384,448), lo [448,512)) at its end -- both at times when no head GEMM owns D1.
// W1 streaming: the 128 KB (hi | lo) image of a head is split into two K-halves of 64 KB, each with its own "landed" and
// "consumed" mbarrier: while the MMAs of K-half 1 of head k run, the TMA of K-half 0 of head k+1 is already in flight, so
// the tensor pipe never waits for a whole-image copy (measured before: 960 cycles of exposed TMA wait per head + the whole
// epilogue serialised behind every GEMM).
// Roles (288 threads): warps 0-7 = the two epilogue groups (below); warp 8, one lane = the MMA issuer.  tcgen05.mma issue is
// NOT free for the issuing thread -- it blocks while the pipe's queue is full (measured: an epilogue thread that also issued
// the next head's 48 MMAs tripled every epilogue) -- so it gets its own warp and plain blocking waits on mbarriers:
// features staged / A1 written / accumulator drained / hidden half written are signalled by the epilogue threads with
// mbarrier.arrive, weights-landed by the TMA transaction count.  The TMA copies themselves are issued by thread 128, which
// polls "K-half consumed" between the chunks of its own epilogue work (a copy is one non-blocking instruction).
// Compiled with -fmad=false (it contains the projection, see g4d_math.cuh).
#include "geom_finish.cuh"
#include "tc_umma.cuh"

namespace g4d {

constexpr int kW1Parts = 4;                       // K-quarters of a head's W1 image: 4 x (hi 16 KB | lo 16 KB)
constexpr uint32_t kW1PartBytes = 2u * 128u * (128u / kW1Parts) * 4u;
constexpr uint32_t kColA1Hi = 0, kColA1Lo = 128, kColD = 256, kColD1 = 384, kColX = 384, kColXLo = 448;

struct TcSmem { uint32_t w1, w2, w0, bias, w2s, bars, acc, total; };

inline TcSmem tc_smem_layout(int F, int max_kp16) {
    TcSmem s{};
    uint32_t off = 0;
    auto take = [&](uint32_t bytes) { uint32_t o = off; off += (bytes + 127u) & ~127u; return o; };
    s.w1 = take(2u * 128 * 128 * 4);
    s.w2 = take(max_kp16 == 48 ? 2u * 48 * 128 * 4 : 128u);
    s.w0 = take(2u * 128 * F * 4);
    s.bias = take((128 + G4D_NUM_HEADS * 128 + 64) * 4);
    s.w2s = take(4 * 128 * 16);
    s.bars = take(256);     // 18 mbarriers
    s.acc = take(2 * 128 * 16);
    s.total = off;
    return s;
}

// ---- weight packing (once per parameter version) --------------------------------------------------------------
struct TcPackDesc {
    const float* src[1 + 2 * G4D_NUM_HEADS];
    float* dst[1 + 2 * G4D_NUM_HEADS];
    int rows_src[1 + 2 * G4D_NUM_HEADS], rows_dst[1 + 2 * G4D_NUM_HEADS], K[1 + 2 * G4D_NUM_HEADS];
    int khalves[1 + 2 * G4D_NUM_HEADS];   // P > 0: image = P K-parts (hi_k0 | lo_k0 | hi_k1 | lo_k1 | ...), each canonical with K / P columns
    int start[2 + 2 * G4D_NUM_HEADS];
    int count;
};

__global__ void tc_pack_weights_kernel(TcPackDesc p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.start[p.count]) return;
    int m = 0;
    while (i >= p.start[m + 1]) ++m;
    const int e = i - p.start[m];
    const int K = p.K[m];
    const uint32_t n = e / K, k = e % K;
    const float v = (int)n < p.rows_src[m] ? __ldg(p.src[m] + n * K + k) : 0.f;
    uint32_t hi, lo;
    tc::tf32_split(v, hi, lo);
    if (p.khalves[m]) {
        const uint32_t Kh = K / p.khalves[m], j = k / Kh, blk = p.rows_dst[m] * Kh;   // words per (hi or lo, K-part) block
        const uint32_t off = tc::canon_off(n, k - j * Kh, Kh) >> 2;
        reinterpret_cast<uint32_t*>(p.dst[m])[(2 * j) * blk + off] = hi;
        reinterpret_cast<uint32_t*>(p.dst[m])[(2 * j + 1) * blk + off] = lo;
        return;
    }
    const uint32_t off = tc::canon_off(n, k, K) >> 2;
    reinterpret_cast<uint32_t*>(p.dst[m])[off] = hi;
    reinterpret_cast<uint32_t*>(p.dst[m])[p.rows_dst[m] * K + off] = lo;
}

size_t tc_packed_floats(const G4DDeformParams& prm) {
    const int F = prm.levels * prm.channels;
    size_t f = 2 * (size_t)128 * F;
    for (int h = 0; h < G4D_NUM_HEADS; ++h) f += 2 * (size_t)128 * 128 + 2 * (size_t)(h == 4 ? 48 : 16) * 128;
    return f + 64;
}

cudaError_t launch_tc_pack_weights(const G4DDeformParams& prm, float* blob, TcWeights* out, cudaStream_t st) {
    TcPackDesc p{};
    const int F = prm.levels * prm.channels;
    int m = 0, total = 0;
    float* q = blob;
    auto add = [&](const float* src, int rows_src, int rows_dst, int K, int khalves = 0) {
        p.src[m] = src; p.dst[m] = q; p.rows_src[m] = rows_src; p.rows_dst[m] = rows_dst; p.K[m] = K; p.khalves[m] = khalves;
        p.start[m] = total;
        total += rows_dst * K; float* r = q; q += 2 * (size_t)rows_dst * K; ++m; return r;
    };
    out->w0 = add(prm.w0, 128, 128, F);
    for (int h = 0; h < G4D_NUM_HEADS; ++h) {
        out->kp16[h] = h == 4 ? 48 : 16;
        out->w1[h] = nullptr; out->w2[h] = nullptr;
        if (!(prm.head_mask & (1 << h))) continue;
        out->w1[h] = add(prm.w1[h], 128, 128, 128, kW1Parts);
        out->w2[h] = add(prm.w2[h], head_out(h), out->kp16[h], 128);
    }
    p.start[m] = total; p.count = m;
    tc_pack_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(p);
    return cudaGetLastError();
}

// ---- the kernel -------------------------------------------------------------------------------------------------
// Thread groups (256 threads; thread 128+m and thread m both own TMEM lane m = Gaussian m of the tile):
//   M = warps 4-7: MMA issue + TMA weight streaming (thread 128), hidden columns [64,128) of every epilogue, head outputs,
//       activations + projection of the Gaussian.  (The hardware scheduler favours the higher warp ids of a sub-partition.)
//   G = warps 0-3: stages the next tile's features into TMEM, hidden columns [0,64) of every epilogue (its partial layer-2
//       sums of the small heads go to M through shared memory), and the SH colour of the current tile.
// Hand-offs use named barriers (ids 1-4 and 6, 256 threads) + tcgen05 fences; the M group syncs internally on id 5.
// Each MMA kind has its own mbarrier (layer 0 / layer 1 / layer-2 partials) so that a thread that only waits for some of
// them can never fall a full phase behind.
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__device__ __forceinline__ int b2off_of(int mask, int h) {
    int o = 0;
    for (int i = 0; i < h; ++i)
        if (mask & (1 << i)) o += head_out(i);
    return o;
}

constexpr int kBarScratch = 3, kBarScratchFree = 4, kBarE = 6;   // named barriers between the two epilogue groups
constexpr uint32_t kColScratch = kColA1Hi;   // p(3) + dsh(48) handed from M to G after the last head (A1 is dead then)

// One channel vector (4 channels) of one level: product over the 6 planes of the bilinear samples.
__device__ __forceinline__ float4 sample_vector(const DeformDesc* __restrict__ dp, int l, int v, int C4, float px, float py, float pz) {
    const DeformDesc& d = *dp;   // lives in shared memory (a by-reference kernel parameter would be spilled to the stack)
    const float pcs[3] = {px, py, pz};
    Tap1D tx[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) tx[a] = make_tap(pcs[a], d.res[l][a]);
    float4 prod = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int c0 = plane_axis0(k), c1 = plane_axis1(k);
        float4 s;
        if (c1 == 3) {
            const float4* rowp = reinterpret_cast<const float4*>(d.trow[l][c0]);
            const float4 r0 = __ldg(rowp + tx[c0].i0 * C4 + v), r1 = __ldg(rowp + tx[c0].i1 * C4 + v);
            const float w0 = tx[c0].w0, w1 = tx[c0].w1;
            s.x = fmaf(r1.x, w1, r0.x * w0); s.y = fmaf(r1.y, w1, r0.y * w0);
            s.z = fmaf(r1.z, w1, r0.z * w0); s.w = fmaf(r1.w, w1, r0.w * w0);
        } else {
            const int W = d.res[l][c0];
            const float4* pl = reinterpret_cast<const float4*>(d.planes[l][k]);
            const Tap1D &X = tx[c0], &Y = tx[c1];
            const float4 nw = __ldg(pl + (Y.i0 * W + X.i0) * C4 + v), ne = __ldg(pl + (Y.i0 * W + X.i1) * C4 + v);
            const float4 sw = __ldg(pl + (Y.i1 * W + X.i0) * C4 + v), se = __ldg(pl + (Y.i1 * W + X.i1) * C4 + v);
            const float wnw = X.w0 * Y.w0, wne = X.w1 * Y.w0, wsw = X.w0 * Y.w1, wse = X.w1 * Y.w1;
            s.x = fmaf(se.x, wse, fmaf(sw.x, wsw, fmaf(ne.x, wne, nw.x * wnw)));
            s.y = fmaf(se.y, wse, fmaf(sw.y, wsw, fmaf(ne.y, wne, nw.y * wnw)));
            s.z = fmaf(se.z, wse, fmaf(sw.z, wsw, fmaf(ne.z, wne, nw.z * wnw)));
            s.w = fmaf(se.w, wse, fmaf(sw.w, wsw, fmaf(ne.w, wne, nw.w * wnw)));
        }
        prod.x *= s.x; prod.y *= s.y; prod.z *= s.z; prod.w *= s.w;
    }
    return prod;
}

template <int C, int L>
__device__ __forceinline__ void sample_features_regs(const DeformDesc* d, const float pcs[3], float (&feat)[C * L]) {
    constexpr int C4 = C / 4;
#pragma unroll
    for (int l = 0; l < L; ++l) {
#pragma unroll
        for (int v = 0; v < C4; ++v) {
            const float4 prod = sample_vector(d, l, v, C4, pcs[0], pcs[1], pcs[2]);
            feat[l * C + 4 * v + 0] = prod.x; feat[l * C + 4 * v + 1] = prod.y;
            feat[l * C + 4 * v + 2] = prod.z; feat[l * C + 4 * v + 3] = prod.w;
        }
    }
}

// ReLU sign bits of one Gaussian for one layer (slot 0 = hidden, 1 + h = head h): [slot][N] x 128 bits
__device__ __forceinline__ void store_relu_bits(uint32_t* base, int slot, int64_t n, int64_t gi, uint64_t lo, uint64_t hi) {
    *reinterpret_cast<uint4*>(base + ((size_t)slot * (size_t)n + (size_t)gi) * 4) =
        make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}

// one 64-bit half of a row's sign bits (word0 = 0 for hidden units 0..63, 2 for 64..127)
__device__ __forceinline__ void store_relu_half(uint32_t* base, int slot, int64_t n, int64_t gi, int word0, uint32_t lo, uint32_t hi) {
    *reinterpret_cast<uint2*>(base + ((size_t)slot * (size_t)n + (size_t)gi) * 4 + word0) = make_uint2(lo, hi);
}

// non-blocking probe of an mbarrier phase (mbar_wait spins on the blocking form)
__device__ __forceinline__ bool mbar_test(void* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kTcThreads = 288;   // 8 epilogue warps + the MMA warp

template <int MODE, int C, int L, bool SAVE>
__global__ void __launch_bounds__(kTcThreads, 1)
deform_tc_kernel(DeformDesc d, TcWeights tw, TcSmem Ls, const CameraDev* __restrict__ camp, float time_arg, int use_cam_time,
                 int64_t n, DeformIO io) {
    constexpr int F = C * L;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ CameraDev cam;
    __shared__ DeformDesc sd;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & 127;
    for (int i = tid; i < (int)(sizeof(DeformDesc) / 4); i += kTcThreads)
        reinterpret_cast<uint32_t*>(&sd)[i] = reinterpret_cast<const uint32_t*>(&d)[i];
    // the hardware scheduler favours the higher warp ids of an SM sub-partition: the latency-critical M group takes them
    const bool is_mma_warp = warp == 8;
    const bool is_m = tid >= 128 && !is_mma_warp;
    const bool tma_thread = tid == 128;
    const int64_t ntiles = (n + 127) / 128;
    if (MODE == 1 || use_cam_time) {
        for (int i = tid; i < (int)(sizeof(CameraDev) / 4); i += kTcThreads)
            reinterpret_cast<uint32_t*>(&cam)[i] = reinterpret_cast<const uint32_t*>(camp)[i];
    }
    float* sBias = reinterpret_cast<float*>(smem + Ls.bias);     // b0[128] | b1[5][128] | b2[64]
    float4* sW2s = reinterpret_cast<float4*>(smem + Ls.w2s);     // [4 small heads][128 hidden]: (W2[0][j], W2[1][j], W2[2][j], W2[3][j])
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Ls.bars);
    // mbarriers: 0 w0 | 1-4 full[j] (W1 K-quarter j landed, tx count) | 5-8 hfree[j] (its MMAs retired) | 9,10 dfull[b] (head
    // GEMM into D[b] complete) | 11 l0 | 12 l2 | 13 featready (128 G threads) | 14 a1ready (256) | 15,16 dfree[b] (256:
    // accumulator drained by its epilogue) | 17 xready (256: SH hidden half written)
    uint64_t *bar_w0 = bars, *bar_full = bars + 1, *bar_hfree = bars + 5, *bar_dfull = bars + 9, *bar_l0 = bars + 11, *bar_l2 = bars + 12;
    uint64_t *bar_feat = bars + 13, *bar_a1 = bars + 14, *bar_dfree = bars + 15, *bar_x = bars + 17;
    int b2off[G4D_NUM_HEADS];
    {
        int o = 0;
#pragma unroll
        for (int h = 0; h < G4D_NUM_HEADS; ++h) { b2off[h] = o; if (d.head_mask & (1 << h)) o += head_out(h); }
    }
    for (int i = tid; i < 128; i += kTcThreads) sBias[i] = __ldg(d.b0 + i);
    for (int h = 0; h < G4D_NUM_HEADS; ++h) {
        if (!(d.head_mask & (1 << h))) continue;
        for (int i = tid; i < 128; i += kTcThreads) sBias[128 + h * 128 + i] = __ldg(d.b1[h] + i);
        for (int i = tid; i < head_out(h); i += kTcThreads) sBias[128 + G4D_NUM_HEADS * 128 + b2off[h] + i] = __ldg(d.b2[h] + i);
        if (h < 4) {
            const int ko = head_out(h);
            for (int j = tid; j < 128; j += kTcThreads)
                sW2s[h * 128 + j] = make_float4(__ldg(d.w2[h] + j), ko > 1 ? __ldg(d.w2[h] + 128 + j) : 0.f,
                                               ko > 2 ? __ldg(d.w2[h] + 256 + j) : 0.f, ko > 3 ? __ldg(d.w2[h] + 384 + j) : 0.f);
        }
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, tc::kTmemCols);
    if (tid == 0) {
        for (int i = 0; i < 13; ++i) mbar_init(bars + i, 1);
        mbar_init(bar_feat, 128); mbar_init(bar_a1, 256); mbar_init(bar_dfree, 256); mbar_init(bar_dfree + 1, 256); mbar_init(bar_x, 256);
        fence_barrier_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base_s, 0);    // warp-uniform for the compiler (MMA operands are uniform registers)
    const uint32_t lane_base = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t sW1 = tc::smem_addr(smem + Ls.w1), sW2 = tc::smem_addr(smem + Ls.w2), sW0 = tc::smem_addr(smem + Ls.w0);
    const bool hsh = d.head_mask & G4D_HEAD_SHS;

    // ---- the W1 stream of this CTA: loads / GEMMs are numbered q = 0, 1, ... over (tile, active head) in program order,
    //      each made of two K-halves j
    const uint32_t m_heads = (uint32_t)__popc(d.head_mask & 31);
    const uint32_t par0 = hsh ? ((m_heads - 1u) & 1u) : 0u;      // accumulator of head k of a tile = (k + par0) & 1: SH gets D0
    const int64_t my_tiles = (int64_t)blockIdx.x < ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const uint32_t total_q = (uint32_t)my_tiles * m_heads;
    // head id of the k-th head this CTA processes: the small heads in an order ROTATED by the CTA index (at any moment the
    // 148 CTAs then stream four different W1 images instead of hammering the L2 slices of one), the SH head always last
    uint32_t hseq_pack = 0;               // 3 bits per position
    {
        const int ns = __popc(d.head_mask & 15);
        const int rot = ns ? (int)(blockIdx.x % (unsigned)ns) : 0;
        int pos = 0;
        for (int h = 0; h < 4; ++h) {
            if (!(d.head_mask & (1 << h))) continue;
            const int k = pos >= rot ? pos - rot : pos - rot + ns;      // active small head #pos runs at position k
            hseq_pack |= (uint32_t)h << (3 * k);
            ++pos;
        }
        if (hsh) hseq_pack |= 4u << (3 * ns);
    }
    auto head_at = [&](uint32_t k) { return (int)((hseq_pack >> (3u * k)) & 7u); };

    // ================================================================================================================
    // MMA warp: one lane issues every tcgen05.mma of the CTA, in order, with blocking waits
    // ================================================================================================================
    if (is_mma_warp) {
        // the warp walks the loop converged, one elected lane issues (uniform-register operands, tc_umma.cuh)
        const bool leader = tc::elect_one_sync();
        {
            mbar_wait(bar_w0, 0);
            uint32_t used[2] = {0u, 0u}, seen[2] = {0u, 0u};       // GEMMs issued into / epilogues observed on D[b]
            uint32_t ph_feat = 0, ph_a1 = 0, ph_x = 0, q = 0;
            auto drain = [&](uint32_t b) {                          // every earlier user of D[b] has been read out
                while (seen[b] < used[b]) { mbar_wait(bar_dfree + b, seen[b] & 1u); ++seen[b]; }
            };
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                // ---- layer 0: D0 = feat * W0^T
                mbar_wait(bar_feat, ph_feat); ph_feat ^= 1u;
                drain(0); drain(1);
                tc::fence_after_sync();
                if (leader) {
                    tc::gemm_3xtf32<F>(tbase + kColD, tbase + kColX, tbase + kColXLo, sW0, sW0 + 128u * F * 4, 128, F, 0, false);
                    tc::umma_commit(bar_l0);
                }
                __syncwarp();
                mbar_wait(bar_a1, ph_a1); ph_a1 ^= 1u;              // epilogue 0 has written A1 and drained D0
                for (uint32_t k = 0; k < m_heads; ++k, ++q) {
                    const uint32_t b = (k + par0) & 1u;
                    drain(b);
#pragma unroll 1
                    for (uint32_t j = 0; j < (uint32_t)kW1Parts; ++j) {
                        constexpr uint32_t KP = 128 / kW1Parts;
                        mbar_wait(bar_full + j, q & 1u);
                        tc::fence_after_sync();
                        if (leader) {
                            tc::gemm_3xtf32<(int)KP>(tbase + (b ? kColD1 : kColD), tbase + kColA1Hi + j * KP, tbase + kColA1Lo + j * KP,
                                                     sW1 + j * kW1PartBytes, sW1 + j * kW1PartBytes + kW1PartBytes / 2, 128, KP, 0, j != 0);
                            tc::umma_commit(bar_hfree + j);
                        }
                        __syncwarp();
                    }
                    if (leader) tc::umma_commit(bar_dfull + b);
                    __syncwarp();
                    ++used[b];
                    if (hsh && k + 1 == m_heads) {
                        // SH head: two layer-2 partials D2 (+)= a2_half * W2[:, 64hh : 64hh+64]^T (D2 = D0 columns [0, 48))
                        for (int hh = 0; hh < 2; ++hh) {
                            mbar_wait(bar_x, ph_x); ph_x ^= 1u;
                            tc::fence_after_sync();
                            if (leader) {
                                tc::gemm_3xtf32<64>(tbase + kColD, tbase + kColX, tbase + kColXLo, sW2, sW2 + 48u * 128 * 4, 48, 128,
                                                    (uint32_t)(hh * 16), hh == 1);
                                tc::umma_commit(bar_l2);
                            }
                            __syncwarp();
                        }
                    }
                }
            }
        }
    } else {
    // ================================================================================================================
    // epilogue groups
    // ================================================================================================================
    // next W1 K-half copy (thread 128 only): its buffer must have been consumed by the previous GEMM; never blocks
    uint32_t tq = 0, tj = 0, tk = 0;      // next copy: GEMM index, K-part, head position inside the tile
    auto pump_tma = [&]() {
        for (;;) {
            if (tq >= total_q || (tq != 0 && !mbar_test(bar_hfree + tj, (tq - 1u) & 1u))) break;
            const int h = head_at(tk);
            mbar_expect_tx(bar_full + tj, kW1PartBytes);
            tma_bulk_g2s(smem + Ls.w1 + tj * kW1PartBytes, tw.w1[h] + tj * (kW1PartBytes / 4u), kW1PartBytes, bar_full + tj);
            if (++tj == (uint32_t)kW1Parts) { tj = 0; ++tq; if (++tk == m_heads) tk = 0; }
        }
    };
    // wait for an mbarrier phase; thread 128 keeps the weight stream moving meanwhile
    auto wait_pumping = [&](uint64_t* bar, uint32_t parity) {
        if (tma_thread) {
            while (!mbar_test(bar, parity)) pump_tma();
        } else {
            mbar_wait(bar, parity);
        }
    };

    // ---- resident operands: W0 and (when active) the SH head's W2 image; the first two W1 K-halves
    if (tma_thread) {
        const uint32_t w2b = hsh ? 2u * 48 * 128 * 4 : 0u;
        mbar_expect_tx(bar_w0, 2u * 128 * F * 4 + w2b);
        tma_bulk_g2s(smem + Ls.w0, tw.w0, 2u * 128 * F * 4, bar_w0);
        if (hsh) tma_bulk_g2s(smem + Ls.w2, tw.w2[4], w2b, bar_w0);
        pump_tma();
    }
    uint32_t ph_l0 = 0, ph_l2 = 0, ph_d0 = 0, ph_d1 = 0;
    // column chunks (16 hidden units each) of every epilogue are split between the two thread groups
    const int ch_lo = is_m ? 4 : 0, ch_hi = is_m ? 8 : 4;        // full-width epilogues (8 chunks)
    const int hc_lo = is_m ? 2 : 0, hc_hi = is_m ? 4 : 2;        // half-width epilogues of the SH head (4 chunks)
    float4* sAcc = reinterpret_cast<float4*>(smem + Ls.acc);     // [2][128]: G group's partial layer-2 sums of a small head

    // features of a tile ([N][F] fp32, written by deform_features_kernel at full occupancy) -> A operand X (hi | lo).
    // X (= D1) is free here: every GEMM that wrote or read it has been waited for by this thread.
    auto stage_features = [&](int64_t tl) {
        const int64_t gs = tl * 128 + row;
        float feat[F];
        if (gs < n) {
            const float4* src = reinterpret_cast<const float4*>(tw.feat + gs * F);
#pragma unroll
            for (int c = 0; c < F; c += 4) {
                const float4 t4 = __ldg(src + (c >> 2));
                feat[c] = t4.x; feat[c + 1] = t4.y; feat[c + 2] = t4.z; feat[c + 3] = t4.w;
            }
        } else {
#pragma unroll
            for (int c = 0; c < F; ++c) feat[c] = 0.f;
        }
#pragma unroll
        for (int c0 = 0; c0 < F; c0 += 8) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) tc::tf32_split(feat[c0 + j], hi[j], lo[j]);
            tc::tmem_st8(lane_base + kColX + c0, hi);
            tc::tmem_st8(lane_base + kColXLo + c0, lo);
        }
        tc::wait_st();
        tc::fence_before_sync();
        mbar_arrive(bar_feat);
    };
    if (!is_m && blockIdx.x < ntiles) stage_features(blockIdx.x);

    // optional per-phase cycle accounting (thread 128; G4D debug only)
    long long cyc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
#define G4D_CYC(i) do { if (tw.dbg && tma_thread) { const long long tn_ = clock64(); cyc[i] += tn_ - tprev; tprev = tn_; } } while (0)
    bool first = true;
    int hseq = 0;   // running small-head counter -> sAcc buffer
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, first = false) {
        const int64_t gi = tile * 128 + row;
        const bool valid = gi < n;
        const bool has_next = tile + gridDim.x < ntiles;
        Vec3 p{0.f, 0.f, 0.f};
        float sl[3] = {0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f}, ol = 0.f;
        if (is_m && valid) {
            p = Vec3{io.xyz[3 * gi], io.xyz[3 * gi + 1], io.xyz[3 * gi + 2]};
            if (io.scaling) { sl[0] = io.scaling[3 * gi]; sl[1] = io.scaling[3 * gi + 1]; sl[2] = io.scaling[3 * gi + 2]; }
            if (io.rotation) { const float4 r4 = *reinterpret_cast<const float4*>(io.rotation + 4 * gi); q[0] = r4.x; q[1] = r4.y; q[2] = r4.z; q[3] = r4.w; }
            if (io.opacity) ol = io.opacity[gi];
        }
        G4D_CYC(0);   // input loads
        // ---- layer 0 (MMA warp): D0 = feat * W0^T
        wait_pumping(bar_l0, ph_l0); ph_l0 ^= 1u;
        tc::fence_after_sync();
        G4D_CYC(2);   // features + layer-0 MMA
        if (is_m && !first) { bar_sync(kBarScratchFree, 256); tc::fence_after_sync(); }   // G has read the previous tile's scratch (A1 region)
        G4D_CYC(3);   // wait scratch free
        // ---- epilogue 0: a1 = relu(D0 + b0) -> A1 (hi | lo); ReLU sign bits saved for the backward
        uint32_t rb[4] = {0u, 0u, 0u, 0u};   // my chunks' sign bits: 2 chunks per word
#pragma unroll 1
        for (int ch = ch_lo; ch < ch_hi; ++ch) {
            const uint32_t c0 = (uint32_t)(ch * 16);
            uint32_t v[16], hi[16], lo[16];
            tc::tmem_ld16(lane_base + kColD + c0, v);
            if (tma_thread) pump_tma();
            tc::wait_ld();
            float bb[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + j);
                bb[j] = b4.x; bb[j + 1] = b4.y; bb[j + 2] = b4.z; bb[j + 3] = b4.w;
            }
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float x = __uint_as_float(v[j]) + bb[j];
                if (SAVE && x > 0.f) bits |= 1u << j;
                tc::tf32_split(fmaxf(x, 0.f), hi[j], lo[j]);
            }
            if (SAVE) { const int k = ch - ch_lo; if (k < 2) rb[0] |= bits << (k * 16); else rb[1] |= bits << ((k - 2) * 16); }
            tc::tmem_st16(lane_base + kColA1Hi + c0, hi);
            tc::tmem_st16(lane_base + kColA1Lo + c0, lo);
        }
        if (SAVE && valid) store_relu_half(tw.relu_bits, 0, n, gi, is_m ? 2 : 0, rb[0], rb[1]);
        tc::wait_st();
        tc::fence_before_sync();
        mbar_arrive(bar_a1);          // A1 complete, D0 drained: the MMA warp may start the head GEMMs of this tile
        G4D_CYC(4);   // epilogue 0

        float dl[11];
#pragma unroll
        for (int j = 0; j < 11; ++j) dl[j] = 0.f;
        float dsh[48];
#pragma unroll
        for (int j = 0; j < 48; ++j) dsh[j] = 0.f;

#pragma unroll 1
        for (uint32_t kact = 0; kact < m_heads; ++kact) {     // heads in this CTA's (rotated) order
            const int h = head_at(kact);
            const float* b1 = sBias + 128 + h * 128;
            const float* b2 = sBias + 128 + G4D_NUM_HEADS * 128 + b2off_of(d.head_mask, h);
            const uint32_t db = (kact + par0) & 1u;                     // accumulator of this head
            const uint32_t dcol = db ? kColD1 : kColD;
            // ---- layer 1 (issued ahead by the MMA warp): wait for D[db] = a1 * W1^T
            wait_pumping(bar_dfull + db, db ? ph_d1 : ph_d0);
            if (db) ph_d1 ^= 1u; else ph_d0 ^= 1u;
            tc::fence_after_sync();
            G4D_CYC(6);   // layer-1 MMA not hidden behind the previous epilogue
            rb[0] = rb[1] = rb[2] = rb[3] = 0u;
            if (h < 4) {
                // ---- small head: layer 2 (k <= 4 outputs) fused into the epilogue in exact fp32 -- cheaper than two
                //      tensor-core round trips for a 128 x 4 x 128 GEMM.  Each group sums its own hidden units.
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4* w2 = sW2s + h * 128;
#pragma unroll 1
                for (int ch = ch_lo; ch < ch_hi; ++ch) {
                    uint32_t v[16];
                    tc::tmem_ld16(lane_base + dcol + ch * 16, v);
                    if (tma_thread) pump_tma();
                    tc::wait_ld();
                    float bb[16];
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(b1 + ch * 16 + j);
                        bb[j] = b4.x; bb[j + 1] = b4.y; bb[j + 2] = b4.z; bb[j + 3] = b4.w;
                    }
                    uint32_t bits = 0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float x = __uint_as_float(v[j]) + bb[j];
                        if (SAVE && x > 0.f) bits |= 1u << j;
                        const float a2 = fmaxf(x, 0.f);
                        const float4 w = w2[ch * 16 + j];
                        acc.x = fmaf(a2, w.x, acc.x); acc.y = fmaf(a2, w.y, acc.y);
                        acc.z = fmaf(a2, w.z, acc.z); acc.w = fmaf(a2, w.w, acc.w);
                    }
                    if (SAVE) { const int k = ch - ch_lo; if (k < 2) rb[0] |= bits << (k * 16); else rb[1] |= bits << ((k - 2) * 16); }
                }
                tc::fence_before_sync();
                mbar_arrive(bar_dfree + db);   // my part of D[db] is read: once all 256 have arrived the MMA warp may reuse it
                if (SAVE && valid) store_relu_half(tw.relu_bits, 1 + h, n, gi, is_m ? 2 : 0, rb[0], rb[1]);
                float4* slot = sAcc + (hseq & 1) * 128 + row;
                if (!is_m) *slot = acc;
                bar_sync(kBarE, 256);   // G's partial sums are visible to M
                if (is_m) {
                    // fixed summation order: (G's hidden units 0..63) + (M's hidden units 64..127)
                    const float4 ga = *slot;
                    acc = make_float4(ga.x + acc.x, ga.y + acc.y, ga.z + acc.z, ga.w + acc.w);
                    if (h == 0) { dl[0] = acc.x + b2[0]; dl[1] = acc.y + b2[1]; dl[2] = acc.z + b2[2]; }
                    else if (h == 1) { dl[3] = acc.x + b2[0]; dl[4] = acc.y + b2[1]; dl[5] = acc.z + b2[2]; }
                    else if (h == 2) { dl[6] = acc.x + b2[0]; dl[7] = acc.y + b2[1]; dl[8] = acc.z + b2[2]; dl[9] = acc.w + b2[3]; }
                    else { dl[10] = acc.x + b2[0]; }
                }
                ++hseq;
                G4D_CYC(8);   // small-head epilogue + fp32 layer 2
            } else {
                // the SH head is the last one and owns D0; D1 (= X) was drained by the previous head's epilogue
#pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {
                    // hidden half hh: a2 = relu(D0[:, 64hh : 64hh+64] + b1) -> X (hi | lo)
#pragma unroll 1
                    for (int ch = hc_lo; ch < hc_hi; ++ch) {
                        const uint32_t cl = (uint32_t)(ch * 16);
                        const uint32_t cg = (uint32_t)(hh * 64) + cl;
                        uint32_t v[16], hi[16], lo[16];
                        tc::tmem_ld16(lane_base + kColD + cg, v);
                        if (tma_thread) pump_tma();
                        tc::wait_ld();
                        float bb[16];
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(b1 + cg + j);
                            bb[j] = b4.x; bb[j + 1] = b4.y; bb[j + 2] = b4.z; bb[j + 3] = b4.w;
                        }
                        uint32_t bits = 0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float x = __uint_as_float(v[j]) + bb[j];
                            if (SAVE && x > 0.f) bits |= 1u << j;
                            tc::tf32_split(fmaxf(x, 0.f), hi[j], lo[j]);
                        }
                        if (SAVE) { if (hh == 0) rb[0] |= bits << ((ch - hc_lo) * 16); else rb[1] |= bits << ((ch - hc_lo) * 16); }
                        tc::tmem_st16(lane_base + kColX + cl, hi);
                        tc::tmem_st16(lane_base + kColXLo + cl, lo);
                    }
                    tc::wait_st();
                    tc::fence_before_sync();
                    mbar_arrive(bar_x);        // hidden half written: the MMA warp issues the layer-2 partial
                    G4D_CYC(8);   // hidden-half epilogue
                    wait_pumping(bar_l2, ph_l2); ph_l2 ^= 1u;
                    tc::fence_after_sync();
                    G4D_CYC(9);   // layer-2 partial MMA
                }
                // sign bits of the SH head: word w of the 128-bit row = hidden units [32w, 32w+32); G owns words 0 and 2
                if (SAVE && valid) {
                    uint32_t* dst = tw.relu_bits + ((size_t)(1 + h) * (size_t)n + (size_t)gi) * 4 + (is_m ? 1 : 0);
                    dst[0] = rb[0]; dst[2] = rb[1];
                }
                if (is_m) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        uint32_t v[16];
                        tc::tmem_ld16(lane_base + kColD + ch * 16, v);
                        tc::wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; ++j) dsh[ch * 16 + j] = __uint_as_float(v[j]) + b2[ch * 16 + j];
                    }
                }
                tc::fence_before_sync();
                mbar_arrive(bar_dfree + db);   // D0 (and D2 inside it) read out: the next tile's layer-0 GEMM may overwrite it
            }
            G4D_CYC(10);  // head output / hand-over
        }
        if (is_m) {
            p.x += dl[0]; p.y += dl[1]; p.z += dl[2];
            // ---- hand the deformed position and the SH deltas to the G thread of this lane (A1 is dead now)
            {
                uint32_t s8[8] = {__float_as_uint(p.x), __float_as_uint(p.y), __float_as_uint(p.z), 0u, 0u, 0u, 0u, 0u};
                tc::tmem_st8(lane_base + kColScratch, s8);
                if (hsh) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        uint32_t v[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(dsh[ch * 16 + j]);
                        tc::tmem_st16(lane_base + kColScratch + 8 + ch * 16, v);
                    }
                }
                tc::wait_st();
                tc::fence_before_sync();
                bar_arrive(kBarScratch, 256);
            }
            if (tma_thread) pump_tma();   // (prefetch of the next tile's first W1 halves once the last GEMM has retired)
            // ---- finish the geometric half of my Gaussian
            if (valid) {
                sl[0] += dl[3]; sl[1] += dl[4]; sl[2] += dl[5];
                q[0] += dl[6]; q[1] += dl[7]; q[2] += dl[8]; q[3] += dl[9];
                ol += dl[10];
                if (MODE == 0) {
                    io.out_xyz[3 * gi] = p.x; io.out_xyz[3 * gi + 1] = p.y; io.out_xyz[3 * gi + 2] = p.z;
                    if (io.out_scaling) { io.out_scaling[3 * gi] = sl[0]; io.out_scaling[3 * gi + 1] = sl[1]; io.out_scaling[3 * gi + 2] = sl[2]; }
                    if (io.out_rotation) *reinterpret_cast<float4*>(io.out_rotation + 4 * gi) = make_float4(q[0], q[1], q[2], q[3]);
                    if (io.out_opacity) io.out_opacity[gi] = ol;
                } else {
                    fused_finish_geometry(cam, io, gi, p, sl, q, ol);
                }
            }
            if (tma_thread) pump_tma();
            G4D_CYC(11);  // scratch hand-off + geometry tail
        } else {
            // ---- G tail: next tile's features, then the SH colour of my Gaussian once M has published p and the SH deltas
            if (has_next) stage_features(tile + gridDim.x);
            bar_sync(kBarScratch, 256);
            tc::fence_after_sync();
            uint32_t s8[8];
            tc::tmem_ld8(lane_base + kColScratch, s8);
            float gsh[48];
            if (hsh) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    uint32_t v[16];
                    tc::tmem_ld16(lane_base + kColScratch + 8 + ch * 16, v);
                    tc::wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) gsh[ch * 16 + j] = __uint_as_float(v[j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 48; ++j) gsh[j] = 0.f;
            }
            tc::wait_ld();
            tc::fence_before_sync();
            bar_arrive(kBarScratchFree, 256);
            if (valid) {
                const Vec3 pg{__uint_as_float(s8[0]), __uint_as_float(s8[1]), __uint_as_float(s8[2])};
                if (MODE == 0) {
                    if (io.out_shs && hsh) {
#pragma unroll
                        for (int j = 0; j < 48; j += 4) {
                            const float4 b = *reinterpret_cast<const float4*>(io.shs + gi * 48 + j);
                            *reinterpret_cast<float4*>(io.out_shs + gi * 48 + j) = make_float4(b.x + gsh[j], b.y + gsh[j + 1], b.z + gsh[j + 2], b.w + gsh[j + 3]);
                        }
                    }
                } else {
                    fused_finish_colour(cam, io, gi, pg, hsh, gsh);
                }
            }
        }
    }
    if (is_m && !first) { bar_sync(kBarScratchFree, 256); }   // pair the G group's last arrive
    if (tw.dbg && tma_thread) {
        for (int i = 0; i < 12; ++i) tw.dbg[blockIdx.x * 12 + i] = cyc[i];
    }
#undef G4D_CYC
    }   // epilogue groups
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, tc::kTmemCols);
}

// ---- HexPlane gather at full occupancy: C/4 threads per Gaussian, one channel vector each -> feat [N][F] fp32 -----------
template <int C4>
__global__ void __launch_bounds__(256) deform_features_kernel(DeformDesc d, int64_t n, const float* __restrict__ xyz, float* __restrict__ feat) {
    pdl_wait();         // the collapsed time rows come from the previous kernel of the stream
    pdl_trigger();
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t g = t / C4;
    const int v = (int)(t % C4);
    if (g >= n) return;
    float pcs[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float amax = __ldg(d.aabb + a), ascale = 2.0f / (__ldg(d.aabb + 3 + a) - amax);
        pcs[a] = (xyz[3 * g + a] - amax) * ascale - 1.0f;
    }
    for (int l = 0; l < d.levels; ++l)
        *reinterpret_cast<float4*>(feat + g * d.F + l * d.C + 4 * v) = sample_vector(&d, l, v, C4, pcs[0], pcs[1], pcs[2]);
}

cudaError_t launch_deform_features(const DeformDesc& d, int64_t n, const float* xyz, float* feat, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const int C4 = d.C / 4;
    const unsigned grid = (unsigned)((n * C4 + 255) / 256);
    if (C4 == 4) return launch_k(deform_features_kernel<4>, dim3(grid), dim3(256), 0, st, true, d, n, xyz, feat);
    if (C4 == 8) return launch_k(deform_features_kernel<8>, dim3(grid), dim3(256), 0, st, true, d, n, xyz, feat);
    return cudaErrorInvalidValue;
}

bool tc_deform_supported(const DeformDesc& d) {
    if (d.WD != 128) return false;
    if (!((d.C == 16 && (d.levels == 2 || d.levels == 3)) || (d.C == 32 && d.levels == 2))) return false;
    const int max_kp = (d.head_mask & G4D_HEAD_SHS) ? 48 : 16;
    return tc_smem_layout(d.F, max_kp).total + 1024 <= 227 * 1024;
}

template <int MODE, int C, int L>
static cudaError_t launch_deform_tc_t(const DeformDesc& d, const TcWeights& tw, const TcSmem& Ls, size_t bytes, int grid,
                                      const CameraDev* cam, float time, bool use_cam_time, int64_t n, const DeformIO& io,
                                      cudaStream_t st) {
    cudaError_t e;
    if (tw.relu_bits) {
        e = cudaFuncSetAttribute(deform_tc_kernel<MODE, C, L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        deform_tc_kernel<MODE, C, L, true><<<grid, kTcThreads, bytes, st>>>(d, tw, Ls, cam, time, use_cam_time ? 1 : 0, n, io);
    } else {
        e = cudaFuncSetAttribute(deform_tc_kernel<MODE, C, L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        deform_tc_kernel<MODE, C, L, false><<<grid, kTcThreads, bytes, st>>>(d, tw, Ls, cam, time, use_cam_time ? 1 : 0, n, io);
    }
    return cudaGetLastError();
}

cudaError_t launch_deform_tc(const DeformDesc& d, const TcWeights& tw, int mode, const CameraDev* cam, float time,
                             bool use_cam_time, int64_t n, const DeformIO& io, int sm_count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    if (!tw.feat) return cudaErrorInvalidValue;
    {
        cudaError_t e = launch_deform_features(d, n, io.xyz, tw.feat, st);
        if (e != cudaSuccess) return e;
    }
    const int max_kp = (d.head_mask & G4D_HEAD_SHS) ? 48 : 16;
    const TcSmem Ls = tc_smem_layout(d.F, max_kp);
    const size_t bytes = Ls.total + 1024;
    const int64_t ntiles = (n + 127) / 128;
    const int grid = (int)(ntiles < sm_count ? ntiles : sm_count);
#define G4D_TC_CASE(CC, LL)                                                                                                   \
    if (d.C == CC && d.levels == LL)                                                                                          \
        return mode == 0 ? launch_deform_tc_t<0, CC, LL>(d, tw, Ls, bytes, grid, cam, time, use_cam_time, n, io, st)          \
                         : launch_deform_tc_t<1, CC, LL>(d, tw, Ls, bytes, grid, cam, time, use_cam_time, n, io, st);
    G4D_TC_CASE(16, 2)
    G4D_TC_CASE(16, 3)
    G4D_TC_CASE(32, 2)
#undef G4D_TC_CASE
    return cudaErrorInvalidValue;
}

}  // namespace g4d
