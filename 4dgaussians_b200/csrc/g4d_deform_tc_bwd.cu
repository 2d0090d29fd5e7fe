// g4d_deform_tc_bwd.cu -- tensor-core backward of the deformation network (net_width 128), two kernels:
//
//  A  "dgrad":  tile of 128 Gaussians per pass (TMEM lanes).  Recomputes the forward activations with BF16x2 tcgen05
//               MMAs (hi + lo parts, 3 products, fp32 accumulate: ~16 mantissa bits, measured gradient error 1.5e-5),
//               forms dz per head in the epilogue (layer 2 is tiny: its dgrad is done in fp32 by the epilogue threads),
//               accumulates d(a1) over the heads IN TMEM with one MMA chain per head, then d(feat) = dh W0 -> [N][F] fp32.
//               Both 128-thread groups own half of the hidden columns of every epilogue.  The ReLU signs come from the
//               bits the forward saved (G4D_RELU_BITS_WORDS), so this is the gradient of the forward that ran.
//               Everything the weight gradients need is written ONCE as ready-to-use MMA operand images
//               (8x8 bf16 core-matrix layout, tc_umma.cuh) so that kernel B is pure TMA + MMA.
//               HexPlane gather (bwd_features_kernel, skipped when the forward's staging buffer was kept) and scatter
//               (bwd_scatter_kernel: plane REDs, d(xyz), residual paths) run around it at full occupancy.
//  B  "wgrad":  dW1_h = DZ_h^T A1, dW2_h^T = A2_h^T DOUT_h, db1_h = DZ_h^T 1, dW0 = DH^T FEAT, db0 = DH^T 1 as split-K
//               tcgen05 GEMMs over the Gaussian index (both operands MN-major straight from the images), accumulators
//               persistent in TMEM across the CTA's tiles, one atomic flush per CTA; two half-tile smem stages pipeline
//               the TMA loads against the MMAs.
//
// Replaces the autograd backward of scene/deformation.py:67-148 + scene/hexplane.py:73-106 (loss.backward(), train.py:219).
// Compiled with the default FMA contraction (no index-producing math here).
#include "deform_bwd_common.cuh"
#include "tc_umma.cuh"

namespace g4d {

namespace {

constexpr uint32_t kA1 = 0, kA1Lo = 64, kD = 128, kDZ = 256, kDZLo = 320, kDA1 = 384;   // TMEM columns (kernel A)
constexpr int kBarFeat = 1, kBarXFree = 2, kBarScratch = 3, kBarScratchFree = 4, kBarM = 5, kBarE = 6;
constexpr uint32_t kImg128 = 128u * 128u * 2u;   // bytes of one 128x128 bf16 part (hi or lo)

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// ---- MMA issue helpers (one thread) -----------------------------------------------------------------------------
// D[128 x N] (+)= A[128 x K] * B^T with A = (hi, lo) bf16 packed two-per-column in TMEM, B = (hi, lo) images in smem.
//   b_mn = false: B image rows = N, cols = K (K-major);  b_mn = true: B image rows = K, cols = N (MN-major)
template <int K>
__device__ __forceinline__ void gemm_bf16x2_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t N,
                                               uint32_t b_ncols, bool b_mn, bool accumulate) {
    const uint32_t idesc = tc::make_idesc_bf16(128, N, false, b_mn);
    const uint32_t rowgrp = (b_ncols >> 3) * 128u;
    const uint32_t lbo = b_mn ? rowgrp : 128u, sbo = b_mn ? 128u : rowgrp;
    const uint32_t step16 = (b_mn ? 2u * rowgrp : 256u) >> 4;   // one K = 16 step, in 16-byte units
    const uint64_t bd_hi = tc::make_smem_desc(b_hi, lbo, sbo), bd_lo = tc::make_smem_desc(b_lo, lbo, sbo);
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const uint32_t a = (p == 0) ? a_lo : a_hi;
        const uint64_t bd = (p == 1) ? bd_lo : bd_hi;
#pragma unroll
        for (int ks = 0; ks < K; ks += 16)
            tc::umma_bf16_ts(d_tmem, a + (ks >> 1), bd + (uint64_t)((ks >> 4) * step16), idesc, accumulate || p > 0 || ks > 0);
    }
}

// D[128 x N] (+)= A^T B with both operands MN-major images (rows = K = Gaussian, cols = M resp. N), K rows
template <int K>
__device__ __forceinline__ void gemm_bf16x2_ss_mn(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_ncols, uint32_t b_hi,
                                                  uint32_t b_lo, uint32_t b_ncols, uint32_t N, bool accumulate, bool b_single) {
    const uint32_t idesc = tc::make_idesc_bf16(128, N, true, true);
    const uint32_t a_grp = (a_ncols >> 3) * 128u, b_grp = (b_ncols >> 3) * 128u;
    const uint64_t ad_hi = tc::make_smem_desc(a_hi, a_grp, 128u), ad_lo = tc::make_smem_desc(a_lo, a_grp, 128u);
    const uint64_t bd_hi = tc::make_smem_desc(b_hi, b_grp, 128u), bd_lo = tc::make_smem_desc(b_lo, b_grp, 128u);
    const uint32_t a_step = (2u * a_grp) >> 4, b_step = (2u * b_grp) >> 4;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        if (b_single && p == 1) continue;   // B has no lo part (e.g. the all-ones operand)
        const uint64_t ad = (p == 0) ? ad_lo : ad_hi;
        const uint64_t bd = (p == 1) ? bd_lo : bd_hi;
#pragma unroll
        for (int ks = 0; ks < K; ks += 16)
            tc::umma_bf16_ss(d_tmem, ad + (uint64_t)((ks >> 4) * a_step), bd + (uint64_t)((ks >> 4) * b_step), idesc,
                             accumulate || p > 0 || ks > 0);
    }
}

// one stage of a transposing butterfly: N values per lane -> N/2 (lanes with `hi` keep the upper half), summed with the partner lane
template <int N, int CAP>
__device__ __forceinline__ void halve(float (&v)[CAP], bool hi, int lane_mask) {
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
        const float send = hi ? v[k] : v[k + N / 2];
        const float keep = hi ? v[k + N / 2] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, lane_mask);
    }
}

__device__ __forceinline__ uint32_t pack2(uint16_t even_k, uint16_t odd_k) { return (uint32_t)even_k | ((uint32_t)odd_k << 16); }

// 16 consecutive values of row `g` -> (hi | lo) bf16: packed TMEM words (8 + 8) and two 16-byte image core rows each
struct Split16 {
    uint32_t hi[8], lo[8];
};
__device__ __forceinline__ void split16(const float (&x)[16], Split16& o) {
    // two elements per cvt: cvt.rn.bf16x2.f32 d, a, b packs a into the upper and b into the lower half (lower = even k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint32_t hi, lo;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x[2 * j + 1]), "f"(x[2 * j]));
        const float he = __uint_as_float(hi << 16), ho = __uint_as_float(hi & 0xffff0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x[2 * j + 1] - ho), "f"(x[2 * j] - he));
        o.hi[j] = hi;
        o.lo[j] = lo;
    }
}
// write 16 columns [c0, c0+16) of row g of a (hi | lo) image with `ncols` columns (lo part at +part_bytes)
__device__ __forceinline__ void store_img16(uint8_t* img, uint32_t part_bytes, uint32_t ncols, uint32_t g, uint32_t c0, const Split16& s) {
    uint8_t* p = img + tc::img16_off(g, c0, ncols);
    *reinterpret_cast<uint4*>(p) = make_uint4(s.hi[0], s.hi[1], s.hi[2], s.hi[3]);
    *reinterpret_cast<uint4*>(p + 128) = make_uint4(s.hi[4], s.hi[5], s.hi[6], s.hi[7]);
    *reinterpret_cast<uint4*>(p + part_bytes) = make_uint4(s.lo[0], s.lo[1], s.lo[2], s.lo[3]);
    *reinterpret_cast<uint4*>(p + part_bytes + 128) = make_uint4(s.lo[4], s.lo[5], s.lo[6], s.lo[7]);
}

}  // namespace

// ---- global-memory layout of the operand images ------------------------------------------------------------------
struct BwdImages {
    uint8_t* a1;                    // [ntiles][2 * kImg128]
    uint8_t* dh;                    // [ntiles][2 * kImg128]
    uint8_t* feat;                  // [ntiles][2 * 128*F*2]
    uint8_t* dz[G4D_NUM_HEADS];     // [ntiles][2 * kImg128]
    uint8_t* a2[G4D_NUM_HEADS];
    uint8_t* dout[G4D_NUM_HEADS];   // [ntiles][2 * 128*kp16*2]
    uint32_t feat_bytes;            // bytes of one (hi | lo) feature image
};

struct BwdSmemA { uint32_t w0, w1, w2s, w2sh, bias, bars, total; };
static BwdSmemA bwd_smem_a(int F, bool sh) {
    BwdSmemA s{};
    uint32_t off = 0;
    auto take = [&](uint32_t b) { uint32_t o = off; off += (b + 127u) & ~127u; return o; };
    s.w0 = take(2u * 128 * F * 2);
    s.w1 = take(2u * 2u * kImg128);          // two buffers of (hi | lo)
    s.w2s = take(4 * 128 * 16);
    s.w2sh = take(sh ? 48 * 128 * 4 : 128);
    s.bias = take((128 + G4D_NUM_HEADS * 128) * 4);
    s.bars = take(64);
    s.total = off;
    return s;
}

// ---- weight images (bf16 hi | lo), rebuilt when the parameter version changes --------------------------------------
struct BwdPackDesc {
    const float* src[1 + G4D_NUM_HEADS];
    uint8_t* dst[1 + G4D_NUM_HEADS];
    int K[1 + G4D_NUM_HEADS];
    int start[2 + G4D_NUM_HEADS];
    int count;
};
__global__ void bwd_pack_weights_kernel(BwdPackDesc p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.start[p.count]) return;
    int m = 0;
    while (i >= p.start[m + 1]) ++m;
    const int e = i - p.start[m], K = p.K[m];
    const uint32_t n = e / K, k = e % K;
    uint16_t hi, lo;
    tc::bf16_split(__ldg(p.src[m] + e), hi, lo);
    const uint32_t off = tc::img16_off(n, k, K);
    *reinterpret_cast<uint16_t*>(p.dst[m] + off) = hi;
    *reinterpret_cast<uint16_t*>(p.dst[m] + 128u * K * 2u + off) = lo;
}

size_t tc_bwd_weight_bytes(const G4DDeformParams& prm) {
    return (size_t)2 * 128 * (prm.levels * prm.channels) * 2 + (size_t)G4D_NUM_HEADS * 2 * kImg128 + 256;
}

cudaError_t launch_tc_bwd_pack_weights(const G4DDeformParams& prm, uint8_t* blob, TcBwdWeights* out, cudaStream_t st) {
    BwdPackDesc p{};
    const int F = prm.levels * prm.channels;
    int m = 0, total = 0;
    uint8_t* q = blob;
    p.src[m] = prm.w0; p.dst[m] = q; p.K[m] = F; p.start[m] = total; total += 128 * F; out->w0 = q; q += (size_t)2 * 128 * F * 2; ++m;
    for (int h = 0; h < G4D_NUM_HEADS; ++h) {
        out->w1[h] = nullptr;
        if (!(prm.head_mask & (1 << h))) continue;
        p.src[m] = prm.w1[h]; p.dst[m] = q; p.K[m] = 128; p.start[m] = total; total += 128 * 128; out->w1[h] = q; q += 2 * kImg128; ++m;
    }
    p.start[m] = total; p.count = m;
    bwd_pack_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(p);
    return cudaGetLastError();
}

// ======================================================================================================
// Kernel A
// ======================================================================================================
struct BwdADesc {
    DeformDesc d;
    TcBwdWeights w;
    BwdImages img;
    const float* go[G4D_NUM_HEADS];
    float* gi[G4D_NUM_HEADS];
    float* g_b2[G4D_NUM_HEADS];
    const float* feat;           // [N][F] fp32 (bwd_features_kernel)
    float* dfeat;                // [N][F] fp32 -> bwd_scatter_kernel
    const uint32_t* relu_bits;   // [6][N][4] + tag (g4d.h G4D_RELU_BITS_WORDS) or NULL
    long long* dbg;              // optional [grid][12] per-phase cycle counters (G4D_OPT_TC_DEBUG)
};

template <int C, int L>
__global__ void __launch_bounds__(256, 1)
deform_tc_bwd_dgrad_kernel(BwdADesc bd, BwdSmemA Ls, float time, int64_t n, const float* __restrict__ xyz) {
    constexpr int F = C * L;
    constexpr int C4 = C / 4;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ DeformDesc sd;
    __shared__ uint32_t tmem_base_s;
    const DeformDesc& d = bd.d;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & 127;
    const bool is_m = tid >= 128;
    const bool issuer = tid == 128;          // TMA copies
    const bool issue_warp = warp == 4;       // tcgen05.mma: the whole warp arrives converged, one elected lane issues (tc_umma.cuh)
    const int64_t ntiles = (n + 127) / 128;
    for (int i = tid; i < (int)(sizeof(DeformDesc) / 4); i += 256)
        reinterpret_cast<uint32_t*>(&sd)[i] = reinterpret_cast<const uint32_t*>(&d)[i];
    float* sBias = reinterpret_cast<float*>(smem + Ls.bias);           // b0[128] | b1[5][128]
    float4* sW2s = reinterpret_cast<float4*>(smem + Ls.w2s);           // small heads: (W2[0][j], W2[1][j], W2[2][j], W2[3][j])
    float* sW2sh = reinterpret_cast<float*>(smem + Ls.w2sh);           // SH head: W2 [48][128] fp32
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Ls.bars);
    uint64_t *bar_w0 = bars, *bar_w1 = bars + 1 /* [2] */, *bar_l0 = bars + 3, *bar_da1 = bars + 4, *bar_l1 = bars + 5, *bar_g6 = bars + 6;
    const bool hsh = d.head_mask & G4D_HEAD_SHS;
    for (int i = tid; i < 128; i += 256) sBias[i] = __ldg(d.b0 + i);
    for (int h = 0; h < G4D_NUM_HEADS; ++h) {
        if (!(d.head_mask & (1 << h))) continue;
        for (int i = tid; i < 128; i += 256) sBias[128 + h * 128 + i] = __ldg(d.b1[h] + i);
        if (h < 4) {
            const int ko = head_out(h);
            for (int j = tid; j < 128; j += 256)
                sW2s[h * 128 + j] = make_float4(__ldg(d.w2[h] + j), ko > 1 ? __ldg(d.w2[h] + 128 + j) : 0.f,
                                               ko > 2 ? __ldg(d.w2[h] + 256 + j) : 0.f, ko > 3 ? __ldg(d.w2[h] + 384 + j) : 0.f);
        } else {
            for (int i = tid; i < 48 * 128; i += 256) sW2sh[i] = __ldg(d.w2[4] + i);
        }
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, tc::kTmemCols);
    if (tid == 0) {
        mbar_init(bar_w0, 1); mbar_init(bar_w1, 1); mbar_init(bar_w1 + 1, 1); mbar_init(bar_l0, 1); mbar_init(bar_da1, 1); mbar_init(bar_l1, 1); mbar_init(bar_g6, 1);
        fence_barrier_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base_s, 0);    // warp-uniform for the compiler (MMA operands are uniform registers)
    const uint32_t lane_base = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t sW0 = tc::smem_addr(smem + Ls.w0), sW1 = tc::smem_addr(smem + Ls.w1);
    float amax[3], ascale[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { amax[a] = __ldg(d.aabb + a); ascale[a] = 2.0f / (__ldg(d.aabb + 3 + a) - amax[a]); }
    int nheads = 0;
    for (int h = 0; h < G4D_NUM_HEADS; ++h) nheads += (d.head_mask >> h) & 1;
    // saved ReLU signs are used when the forward that produced them says so (tag word behind the bits)
    const bool use_bits = bd.relu_bits && __ldg(bd.relu_bits + (size_t)24 * (size_t)n) == 0x5A5A5A5Au;
    auto load_bits = [&](int slot, int64_t gi, bool valid, uint64_t& lo, uint64_t& hi) {
        uint4 m = make_uint4(0u, 0u, 0u, 0u);
        if (valid) m = __ldg(reinterpret_cast<const uint4*>(bd.relu_bits + ((size_t)slot * (size_t)n + (size_t)gi) * 4));
        lo = (uint64_t)m.x | ((uint64_t)m.y << 32);
        hi = (uint64_t)m.z | ((uint64_t)m.w << 32);
    };

    // W1 images are double-buffered: the k-th (tile, head) pair this CTA processes uses buffer k & 1; pair k + 2 is loaded
    // when pair k retires
    auto load_w1 = [&](int h, int buf) {
        mbar_expect_tx(bar_w1 + buf, 2u * kImg128);
        tma_bulk_g2s(smem + Ls.w1 + buf * 2u * kImg128, bd.w.w1[h], 2u * kImg128, bar_w1 + buf);
    };
    auto nth_head = [&](int idx) {
        int m = d.head_mask;
        for (int i = 0; i < idx; ++i) m &= m - 1;
        return __ffs(m) - 1;
    };
    const int first_h = __ffs(d.head_mask) - 1;
    const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int64_t total_uses = my_tiles * nheads;
    if (issuer) {
        mbar_expect_tx(bar_w0, 2u * 128 * F * 2);
        tma_bulk_g2s(smem + Ls.w0, bd.w.w0, 2u * 128 * F * 2, bar_w0);
        load_w1(first_h, 0);
        if (total_uses > 1) load_w1(nth_head(1 % nheads), 1);
    }
    if (is_m) mbar_wait(bar_w0, 0);
    uint32_t ph_w1_0 = 0, ph_w1_1 = 0, ph_l0 = 0, ph_g6 = 0, ph_l1 = 0, ph_da1 = 0;
    int64_t seq = 0;      // running (tile, head) counter -> W1 buffer
    // head epilogues are shared: the G group takes column chunks [0, kGChunks), the M group the rest
    constexpr int kGChunks = 4;
    const int ch_lo = is_m ? kGChunks : 0, ch_hi = is_m ? 8 : kGChunks;

    // ---- features of a tile: [N][F] fp32 from bwd_features_kernel -> A operand of layer 0 (DZ region) + FEAT image
    float feat[F];
    auto load_features = [&](int64_t tl) {
        const int64_t gs = tl * 128 + row;
        if (gs < n) {
            const float4* src = reinterpret_cast<const float4*>(bd.feat + gs * F);
#pragma unroll
            for (int c = 0; c < F; c += 4) {
                const float4 t4 = __ldg(src + (c >> 2));
                feat[c] = t4.x; feat[c + 1] = t4.y; feat[c + 2] = t4.z; feat[c + 3] = t4.w;
            }
        } else {
#pragma unroll
            for (int c = 0; c < F; ++c) feat[c] = 0.f;
        }
    };
    auto write_features = [&](int64_t tl) {
        uint8_t* img_f = bd.img.feat + (size_t)tl * bd.img.feat_bytes;
#pragma unroll
        for (int c0 = 0; c0 < F; c0 += 16) {
            float x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = feat[c0 + j];
            Split16 s;
            split16(x, s);
            tc::tmem_st8(lane_base + kDZ + (c0 >> 1), s.hi);
            tc::tmem_st8(lane_base + kDZLo + (c0 >> 1), s.lo);
            store_img16(img_f, 128u * F * 2u, (uint32_t)F, (uint32_t)row, (uint32_t)c0, s);
        }
        tc::wait_st();
        tc::fence_before_sync();
        bar_arrive(kBarFeat, 256);
    };
    if (!is_m && blockIdx.x < ntiles) {
        load_features(blockIdx.x);
        write_features(blockIdx.x);
    }

    float acc_b2s[4] = {0.f, 0.f, 0.f, 0.f};   // (M) db2 partial sums of the small heads (butterfly-distributed over lanes)
    float acc_b2sh[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int lane = tid & 31;
    long long cyc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
    const bool prober = bd.dbg && (tid == 128 || tid == 0);
#define G4D_CYC(i) do { if (prober) { const long long tn_ = clock64(); cyc[(i) + (is_m ? 0 : 6)] += tn_ - tprev; tprev = tn_; } } while (0)
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t gi = tile * 128 + row;
        const bool valid = gi < n;
        const bool has_next = tile + gridDim.x < ntiles;
        auto load_dout = [&](int hh, float (&dst)[48]) {
#pragma unroll
            for (int o = 0; o < 48; ++o) dst[o] = 0.f;
            if (valid && bd.go[hh]) {
                const int kk = head_out(hh);
                const float* gp = bd.go[hh] + gi * kk;
                if (hh == 4) {
#pragma unroll
                    for (int o = 0; o < 48; o += 4) {
                        const float4 t4 = __ldg(reinterpret_cast<const float4*>(gp + o));
                        dst[o] = t4.x; dst[o + 1] = t4.y; dst[o + 2] = t4.z; dst[o + 3] = t4.w;
                    }
                } else {
#pragma unroll
                    for (int o = 0; o < 4; ++o)
                        if (o < kk) dst[o] = __ldg(gp + o);
                }
            }
        };
        float dn[48];
        load_dout(first_h, dn);
        uint64_t hm0 = 0ull, hm1 = 0ull;   // bit j set <=> hidden[j] > 0 (only this thread's own chunks when recomputed)
        if (use_bits) load_bits(0, gi, valid, hm0, hm1);
        // ---- layer 0 recompute: D = feat W0^T
        if (is_m) {
            bar_sync(kBarFeat, 256);
            if (issue_warp) {
                const bool leader = tc::elect_one_sync();
                tc::fence_after_sync();
                if (leader) {
                    gemm_bf16x2_ts<F>(tbase + kD, tbase + kDZ, tbase + kDZLo, sW0, sW0 + 128u * F * 2, 128, F, false, false);
                    tc::umma_commit(bar_l0);
                }
                __syncwarp();
            }
        }
        mbar_wait(bar_l0, ph_l0); ph_l0 ^= 1u;
        tc::fence_after_sync();
        G4D_CYC(0);   // wait features + layer 0
        {
            uint8_t* img_a1 = bd.img.a1 + (size_t)tile * 2 * kImg128;
#pragma unroll 1
            for (int ch = ch_lo; ch < ch_hi; ++ch) {
                uint32_t v[16];
                tc::tmem_ld16(lane_base + kD + ch * 16, v);
                tc::wait_ld();
                float a1v[16];
                uint32_t bits = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float hpre = __uint_as_float(v[j]) + sBias[ch * 16 + j];
                    if (hpre > 0.f) bits |= 1u << j;
                    a1v[j] = fmaxf(hpre, 0.f);
                }
                if (use_bits) {
                    const uint32_t sb = (uint32_t)((ch < 4 ? hm0 : hm1) >> ((ch & 3) * 16)) & 0xffffu;
#pragma unroll
                    for (int j = 0; j < 16; ++j) a1v[j] = ((sb >> j) & 1u) ? a1v[j] : 0.f;
                } else if (ch < 4) hm0 |= (uint64_t)bits << (ch * 16);
                else hm1 |= (uint64_t)bits << ((ch - 4) * 16);
                Split16 s;
                split16(a1v, s);
                tc::tmem_st8(lane_base + kA1 + ch * 8, s.hi);
                tc::tmem_st8(lane_base + kA1Lo + ch * 8, s.lo);
                store_img16(img_a1, kImg128, 128, (uint32_t)row, (uint32_t)(ch * 16), s);
            }
            tc::wait_st();
            tc::fence_before_sync();
            G4D_CYC(1);   // epilogue 0 (+ dh epilogue below)
            bar_sync(kBarE, 256);
            G4D_CYC(4);   // wait for the other group
        }
        if (is_m) {
            // ---- first layer-1 GEMM of the tile
            const int buf = (int)(seq & 1);
            if (buf == 0) { mbar_wait(bar_w1, ph_w1_0); ph_w1_0 ^= 1u; } else { mbar_wait(bar_w1 + 1, ph_w1_1); ph_w1_1 ^= 1u; }
            if (issue_warp) {
                const bool leader = tc::elect_one_sync();
                tc::fence_after_sync();
                const uint32_t w = sW1 + buf * 2u * kImg128;
                if (leader) {
                    gemm_bf16x2_ts<128>(tbase + kD, tbase + kA1, tbase + kA1Lo, w, w + kImg128, 128, 128, false, false);
                    tc::umma_commit(bar_l1);
                }
                __syncwarp();
            }
        }

        // ================= heads: both groups, column chunks split =================
        bool da1_started = false;
        int h = first_h;
#pragma unroll 1
        for (int hc = 0; hc < nheads; ++hc) {
            const int buf = (int)(seq & 1);
            const float* b1 = sBias + 128 + h * 128;
            int next_h = -1;
            {
                const int later = d.head_mask >> (h + 1);
                if (later) next_h = h + 1 + (__ffs(later) - 1);
            }
            // my row of dL/d(out_h): fetched one head ahead (the loads fly during the previous head's epilogue)
            float dout[48];
#pragma unroll
            for (int o = 0; o < 48; ++o) dout[o] = dn[o];
            if (next_h >= 0) load_dout(next_h, dn);
            if (is_m) {
                // DOUT image (operand of dW2)
                const uint32_t kp16 = (h == 4) ? 48u : 16u;
                uint8_t* img = bd.img.dout[h] + (size_t)tile * 2 * 128 * kp16 * 2;
#pragma unroll
                for (int c0 = 0; c0 < 48; c0 += 16) {
                    if (c0 > 0 && h != 4) break;
                    float x[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) x[j] = dout[c0 + j];
                    Split16 s;
                    split16(x, s);
                    store_img16(img, 128u * kp16 * 2u, kp16, (uint32_t)row, (uint32_t)c0, s);
                }
                // db2: transposing butterfly over the warp, partial sums stay distributed over the lanes until the end
                if (h == 4) {
                    float t[48];
#pragma unroll
                    for (int o = 0; o < 48; ++o) t[o] = dout[o];
                    halve<48>(t, lane & 16, 16); halve<24>(t, lane & 8, 8); halve<12>(t, lane & 4, 4);
#pragma unroll
                    for (int k = 0; k < 6; ++k) acc_b2sh[k] += t[k];
                } else {
                    float t[4] = {dout[0], dout[1], dout[2], dout[3]};
                    halve<4>(t, lane & 16, 16); halve<2>(t, lane & 8, 8);
                    if (h == 0) acc_b2s[0] += t[0]; else if (h == 1) acc_b2s[1] += t[0];
                    else if (h == 2) acc_b2s[2] += t[0]; else acc_b2s[3] += t[0];
                }
            }
            uint64_t zm0 = 0ull, zm1 = 0ull;
            if (use_bits) load_bits(1 + h, gi, valid, zm0, zm1);
            G4D_CYC(2);   // dout / DOUT image / db2
            // ---- wait for z = a1 W1^T
            mbar_wait(bar_l1, ph_l1); ph_l1 ^= 1u;
            tc::fence_after_sync();
            G4D_CYC(5);   // MMA waits
            uint8_t* img_a2 = bd.img.a2[h] + (size_t)tile * 2 * kImg128;
            uint8_t* img_dz = bd.img.dz[h] + (size_t)tile * 2 * kImg128;
#pragma unroll 1
            for (int ch = ch_lo; ch < ch_hi; ++ch) {
                uint32_t v[16];
                tc::tmem_ld16(lane_base + kD + ch * 16, v);
                tc::wait_ld();
                float a2v[16], dzv[16];
                const uint32_t sb = (uint32_t)((ch < 4 ? zm0 : zm1) >> ((ch & 3) * 16)) & 0xffffu;
                if (h < 4) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float zc = __uint_as_float(v[j]) + b1[ch * 16 + j];
                        const float4 w = sW2s[h * 128 + ch * 16 + j];
                        const float da2 = dout[0] * w.x + dout[1] * w.y + dout[2] * w.z + dout[3] * w.w;
                        const bool on = use_bits ? ((sb >> j) & 1u) != 0u : zc > 0.f;
                        a2v[j] = on ? fmaxf(zc, 0.f) : 0.f;
                        dzv[j] = on ? da2 : 0.f;
                    }
                } else {
                    float da2[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) da2[j] = 0.f;
#pragma unroll 4
                    for (int o = 0; o < 48; ++o) {
                        const float* wr = sW2sh + o * 128 + ch * 16;
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 w = *reinterpret_cast<const float4*>(wr + j);
                            da2[j] = fmaf(dout[o], w.x, da2[j]); da2[j + 1] = fmaf(dout[o], w.y, da2[j + 1]);
                            da2[j + 2] = fmaf(dout[o], w.z, da2[j + 2]); da2[j + 3] = fmaf(dout[o], w.w, da2[j + 3]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float zc = __uint_as_float(v[j]) + b1[ch * 16 + j];
                        const bool on = use_bits ? ((sb >> j) & 1u) != 0u : zc > 0.f;
                        a2v[j] = on ? fmaxf(zc, 0.f) : 0.f;
                        dzv[j] = on ? da2[j] : 0.f;
                    }
                }
                Split16 s;
                split16(a2v, s);
                store_img16(img_a2, kImg128, 128, (uint32_t)row, (uint32_t)(ch * 16), s);
                split16(dzv, s);
                store_img16(img_dz, kImg128, 128, (uint32_t)row, (uint32_t)(ch * 16), s);
                tc::tmem_st8(lane_base + kDZ + ch * 8, s.hi);
                tc::tmem_st8(lane_base + kDZLo + ch * 8, s.lo);
            }
            tc::wait_st();
            tc::fence_before_sync();
            G4D_CYC(3);   // head epilogue (my chunks)
            bar_sync(kBarE, 256);
            G4D_CYC(4);   // wait for the other group
            // ---- d(a1) += dz W1 (W1 image read MN-major), then the next head's layer 1 straight behind it
            if (issue_warp) {
                const bool leader = tc::elect_one_sync();
                tc::fence_after_sync();
                const uint32_t w = sW1 + buf * 2u * kImg128;
                if (leader) {
                    gemm_bf16x2_ts<128>(tbase + kDA1, tbase + kDZ, tbase + kDZLo, w, w + kImg128, 128, 128, true, da1_started);
                    tc::umma_commit(bar_da1);
                }
                __syncwarp();
            }
            da1_started = true;
            ++seq;
            if (next_h >= 0 && is_m) {
                const int nbuf = (int)(seq & 1);
                if (nbuf == 0) { mbar_wait(bar_w1, ph_w1_0); ph_w1_0 ^= 1u; } else { mbar_wait(bar_w1 + 1, ph_w1_1); ph_w1_1 ^= 1u; }
                if (issue_warp) {
                    const bool leader = tc::elect_one_sync();
                    const uint32_t w = sW1 + nbuf * 2u * kImg128;
                    if (leader) {
                        gemm_bf16x2_ts<128>(tbase + kD, tbase + kA1, tbase + kA1Lo, w, w + kImg128, 128, 128, false, false);
                        tc::umma_commit(bar_l1);
                    }
                    __syncwarp();
                }
            }
            // the dz W1 chain has retired: DZ may be overwritten, and W1 buffer `buf` takes the pair after next
            mbar_wait(bar_da1, ph_da1); ph_da1 ^= 1u;
            tc::fence_after_sync();
            if (issuer && seq + 1 < total_uses) load_w1(nth_head((int)((seq + 1) % nheads)), buf);
            h = next_h;
            G4D_CYC(5);   // MMA waits
        }

        // ---- dh = d(a1) * (hidden > 0) -> A operand (DZ region) + DH image
        {
            uint8_t* img_dh = bd.img.dh + (size_t)tile * 2 * kImg128;
#pragma unroll 1
            for (int ch = ch_lo; ch < ch_hi; ++ch) {
                uint32_t v[16];
                tc::tmem_ld16(lane_base + kDA1 + ch * 16, v);
                tc::wait_ld();
                float dh[16];
                const uint32_t bits = (uint32_t)((ch < 4 ? hm0 : hm1) >> ((ch & 3) * 16)) & 0xffffu;
#pragma unroll
                for (int j = 0; j < 16; ++j) dh[j] = ((bits >> j) & 1u) ? __uint_as_float(v[j]) : 0.f;
                Split16 s;
                split16(dh, s);
                store_img16(img_dh, kImg128, 128, (uint32_t)row, (uint32_t)(ch * 16), s);
                tc::tmem_st8(lane_base + kDZ + ch * 8, s.hi);
                tc::tmem_st8(lane_base + kDZLo + ch * 8, s.lo);
            }
            tc::wait_st();
            tc::fence_before_sync();
            G4D_CYC(1);
            bar_sync(kBarE, 256);
            G4D_CYC(4);
        }
        if (is_m) {
            // ---- d(feat) = dh W0  (W0 image read MN-major: rows = hidden j = K, cols = feature f = N) -> [N][F] fp32
            if (issue_warp) {
                const bool leader = tc::elect_one_sync();
                tc::fence_after_sync();
                if (leader) {
                    gemm_bf16x2_ts<128>(tbase + kD, tbase + kDZ, tbase + kDZLo, sW0, sW0 + 128u * F * 2, F, F, true, false);
                    tc::umma_commit(bar_g6);
                }
                __syncwarp();
            }
            mbar_wait(bar_g6, ph_g6); ph_g6 ^= 1u;
            tc::fence_after_sync();
            if (has_next) bar_arrive(kBarXFree, 256);   // the DZ region (next tile's feature operand) is free
#pragma unroll
            for (int c0 = 0; c0 < F; c0 += 16) {
                uint32_t v[16];
                tc::tmem_ld16(lane_base + kD + c0, v);
                tc::wait_ld();
                if (valid) {
                    float4* dst = reinterpret_cast<float4*>(bd.dfeat + gi * F + c0);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                             __uint_as_float(v[4 * q + 3]));
                }
            }
            tc::fence_before_sync();
            G4D_CYC(5);
        } else if (has_next) {
            load_features(tile + gridDim.x);
            bar_sync(kBarXFree, 256);
            tc::fence_after_sync();
            write_features(tile + gridDim.x);
            G4D_CYC(5);
        }
    }
    if (is_m) {
        // db2 flush: finish the butterflies (remaining lane bits) and add once per warp
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
            float x = acc_b2s[hh];
            x += __shfl_xor_sync(0xffffffffu, x, 4); x += __shfl_xor_sync(0xffffffffu, x, 2); x += __shfl_xor_sync(0xffffffffu, x, 1);
            const int idx = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
            if ((d.head_mask & (1 << hh)) && (lane & 7) == 0 && idx < head_out(hh)) atomicAdd(bd.g_b2[hh] + idx, x);
        }
        if (d.head_mask & G4D_HEAD_SHS) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                float x = acc_b2sh[k];
                x += __shfl_xor_sync(0xffffffffu, x, 2); x += __shfl_xor_sync(0xffffffffu, x, 1);
                const int idx = ((lane >> 4) & 1) * 24 + ((lane >> 3) & 1) * 12 + ((lane >> 2) & 1) * 6 + k;
                if ((lane & 3) == 0) atomicAdd(bd.g_b2[4] + idx, x);
            }
        }
    }
    if (prober) {
        // slots (M: 0..5, G: 6..11): 0 wait features + layer 0, 1 epilogue 0 + dh epilogue, 2 dout/DOUT/db2, 3 head epilogues,
        //                           4 waiting for the other group, 5 MMA waits + tile tail
        for (int i = 0; i < 6; ++i) bd.dbg[blockIdx.x * 12 + i + (is_m ? 0 : 6)] = cyc[i + (is_m ? 0 : 6)];
    }
#undef G4D_CYC
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, tc::kTmemCols);
}

// ---- HexPlane gather / scatter at full occupancy (C/4 threads per Gaussian, one channel vector each) ---------------
G4D_D float4 hexplane_vector(const DeformDesc& d, int l, int v, int C4, const float pcs[3]) {
    Tap1D tx[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) tx[a] = make_tap(pcs[a], d.res[l][a]);
    float4 prod = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int c0 = plane_axis0(k), c1 = plane_axis1(k);
        float4 sv;
        if (c1 == 3) {
            const float4* rowp = reinterpret_cast<const float4*>(d.trow[l][c0]);
            const float4 r0 = __ldg(rowp + tx[c0].i0 * C4 + v), r1 = __ldg(rowp + tx[c0].i1 * C4 + v);
            const float w0 = tx[c0].w0, w1 = tx[c0].w1;
            sv = make_float4(fmaf(r1.x, w1, r0.x * w0), fmaf(r1.y, w1, r0.y * w0), fmaf(r1.z, w1, r0.z * w0), fmaf(r1.w, w1, r0.w * w0));
        } else {
            const int W = d.res[l][c0];
            const float4* pl = reinterpret_cast<const float4*>(d.planes[l][k]);
            const Tap1D &X = tx[c0], &Y = tx[c1];
            const float4 nw = __ldg(pl + (Y.i0 * W + X.i0) * C4 + v), ne = __ldg(pl + (Y.i0 * W + X.i1) * C4 + v);
            const float4 sw = __ldg(pl + (Y.i1 * W + X.i0) * C4 + v), se = __ldg(pl + (Y.i1 * W + X.i1) * C4 + v);
            const float wnw = X.w0 * Y.w0, wne = X.w1 * Y.w0, wsw = X.w0 * Y.w1, wse = X.w1 * Y.w1;
            sv = make_float4(fmaf(se.x, wse, fmaf(sw.x, wsw, fmaf(ne.x, wne, nw.x * wnw))),
                             fmaf(se.y, wse, fmaf(sw.y, wsw, fmaf(ne.y, wne, nw.y * wnw))),
                             fmaf(se.z, wse, fmaf(sw.z, wsw, fmaf(ne.z, wne, nw.z * wnw))),
                             fmaf(se.w, wse, fmaf(sw.w, wsw, fmaf(ne.w, wne, nw.w * wnw))));
        }
        prod.x *= sv.x; prod.y *= sv.y; prod.z *= sv.z; prod.w *= sv.w;
    }
    return prod;
}

template <int C4>
__global__ void __launch_bounds__(256) bwd_features_kernel(DeformDesc d, int64_t n, const float* __restrict__ xyz, float* __restrict__ feat) {
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
        *reinterpret_cast<float4*>(feat + g * d.F + l * d.C + 4 * v) = hexplane_vector(d, l, v, C4, pcs);
}

struct BwdScatterDesc {
    DeformDesc d;
    const float* dfeat;                     // [N][F]
    float* g_planes[G4D_MAX_LEVELS][6];
    float* trow_grad[G4D_MAX_LEVELS][3];
    const float* go[G4D_NUM_HEADS];
    float* gi[G4D_NUM_HEADS];
};

template <int C4>
__global__ void __launch_bounds__(256, 2) bwd_scatter_kernel(BwdScatterDesc sd, int64_t n, const float* __restrict__ xyz) {
    const DeformDesc& d = sd.d;
    constexpr int GPB = 256 / C4;           // Gaussians per block
    const int64_t g0 = (int64_t)blockIdx.x * GPB;
    const int64_t g = g0 + threadIdx.x / C4;
    const int v = threadIdx.x % C4;
    // residual path: d(out)/d(in) = identity for scaling / rotation / opacity / shs (contiguous ranges, coalesced)
    for (int hh = 1; hh < G4D_NUM_HEADS; ++hh) {
        if (!sd.gi[hh]) continue;
        const int ko = head_out(hh);
        const int64_t lo = g0 * ko, hi = (g0 + GPB < n ? g0 + GPB : n) * ko;
        for (int64_t i = lo + threadIdx.x; i < hi; i += 256) sd.gi[hh][i] = sd.go[hh] ? sd.go[hh][i] : 0.f;
    }
    float gpix[3] = {0.f, 0.f, 0.f};
    float ascale[3] = {0.f, 0.f, 0.f};
    if (g < n) {
        float pcs[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float amax = __ldg(d.aabb + a);
            ascale[a] = 2.0f / (__ldg(d.aabb + 3 + a) - amax);
            pcs[a] = (xyz[3 * g + a] - amax) * ascale[a] - 1.0f;
        }
        for (int l = 0; l < d.levels; ++l) {
            const float4 df = __ldg(reinterpret_cast<const float4*>(sd.dfeat + g * d.F + l * d.C + 4 * v));
            scatter_vector(d, sd.g_planes, sd.trow_grad, l, v, C4, pcs, df, gpix);
        }
    }
    // the C4 threads of a Gaussian are adjacent lanes
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int o = 1; o < C4; o <<= 1) gpix[a] += __shfl_xor_sync(0xffffffffu, gpix[a], o);
    if (g < n && v == 0 && sd.gi[0]) {
#pragma unroll
        for (int a = 0; a < 3; ++a) sd.gi[0][g * 3 + a] = (sd.go[0] ? sd.go[0][g * 3 + a] : 0.f) + gpix[a] * ascale[a];
    }
}

// ======================================================================================================
// Kernel B: weight gradients from the operand images.
//   CTA (group, chunk): group 0..4 = head, group 5 = layer 0.  128 threads; thread 0 streams + issues, all read out.
// ======================================================================================================
struct BwdBDesc {
    BwdImages img;
    int head_mask, F, nheads, chunks_head, chunks_l0;
    int64_t ntiles;
    float* g_w0; float* g_b0;
    float* g_w1[G4D_NUM_HEADS]; float* g_b1[G4D_NUM_HEADS]; float* g_w2[G4D_NUM_HEADS];
};

__global__ void __launch_bounds__(128, 1) deform_tc_bwd_wgrad_kernel(BwdBDesc b) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar_ld[2], bar_mma[2], bar_done;
    const int tid = threadIdx.x, warp = tid >> 5;
    // CTAs [0, nheads * chunks_head) serve the active heads, the rest layer 0 (CTA counts proportional to the bytes per tile)
    const bool layer0 = (int)blockIdx.x >= b.nheads * b.chunks_head;
    const int chunk = layer0 ? (int)blockIdx.x - b.nheads * b.chunks_head : (int)blockIdx.x % b.chunks_head;
    const int nch = layer0 ? b.chunks_l0 : b.chunks_head;
    int h = 5;
    if (!layer0) {
        int m = b.head_mask;
        for (int i = 0; i < (int)blockIdx.x / b.chunks_head; ++i) m &= m - 1;
        h = __ffs(m) - 1;
    }
    const int kp16 = (h == 4) ? 48 : 16;
    // Two pipeline stages of HALF a tile each (64 Gaussians = the K extent of one MMA chain): the images are row-group major,
    // so rows [64 hf, 64 hf + 64) of a part are one contiguous half of its bytes.  While the tensor core chews on one stage the
    // TMA fills the other.  Per stage: X (DZ_h or DH) | Y (A1 or FEAT) | A2_h | DOUT_h, each as (hi half | lo half).
    constexpr uint32_t kHalf128 = kImg128 / 2;                       // 16 KB: 64 rows x 128 cols bf16
    constexpr uint32_t kStage = 3u * 2u * kHalf128 + 2u * 64 * 48 * 2;   // 108 KB
    uint8_t* sOnes = smem + 2 * kStage;
    // ones image [64 g][16]: bf16 1.0 = 0x3F80
    for (int i = tid; i < 64 * 16; i += 128) reinterpret_cast<uint16_t*>(sOnes)[i] = 0x3F80;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, tc::kTmemCols);
    if (tid == 0) {
        mbar_init(&bar_ld[0], 1); mbar_init(&bar_ld[1], 1); mbar_init(&bar_mma[0], 1); mbar_init(&bar_mma[1], 1); mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base_s, 0);    // warp-uniform for the compiler
    const uint32_t lane_base = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t colW = 0, colW2 = 128, colB = 192;   // accumulators: dW (128 or F cols) | dW2^T (kp16) | bias (16)
    const int64_t per = (b.ntiles + nch - 1) / nch;
    const int64_t t0 = (int64_t)chunk * per, t1 = (t0 + per < b.ntiles) ? t0 + per : b.ntiles;
    const uint32_t yp = layer0 ? 128u * b.F * 2u : kImg128;            // bytes of one part (hi or lo) of Y
    const uint32_t dp = 128u * kp16 * 2u;                              // ... of DOUT
    const int64_t nwork = t1 > t0 ? 2 * (t1 - t0) : 0;
    bool started = false;
    if (warp == 0 && nwork > 0) {
        // warp 0 walks the work list converged; one elected lane issues the TMA copies and the MMAs (tc_umma.cuh)
        const bool leader = tc::elect_one_sync();
        auto load = [&](int64_t w) {
            const int st = (int)(w & 1), hf = (int)(w & 1);             // work item w = 2 (t - t0) + hf uses stage w & 1
            const int64_t t = t0 + (w >> 1);
            uint8_t* base = smem + st * kStage;
            uint8_t *dX = base, *dY = base + 2 * kHalf128, *dA2 = base + 4 * kHalf128, *dDO = base + 6 * kHalf128;
            const uint8_t* srcX = (layer0 ? b.img.dh : b.img.dz[h]) + (size_t)t * 2 * kImg128;
            const uint8_t* srcY = layer0 ? b.img.feat + (size_t)t * b.img.feat_bytes : b.img.a1 + (size_t)t * 2 * kImg128;
            const uint32_t bytes = 2u * kHalf128 + yp + (layer0 ? 0u : 2u * kHalf128 + dp);
            mbar_expect_tx(&bar_ld[st], bytes);
            tma_bulk_g2s(dX, srcX + hf * kHalf128, kHalf128, &bar_ld[st]);
            tma_bulk_g2s(dX + kHalf128, srcX + kImg128 + hf * kHalf128, kHalf128, &bar_ld[st]);
            tma_bulk_g2s(dY, srcY + hf * (yp / 2), yp / 2, &bar_ld[st]);
            tma_bulk_g2s(dY + yp / 2, srcY + yp + hf * (yp / 2), yp / 2, &bar_ld[st]);
            if (!layer0) {
                const uint8_t* srcA = b.img.a2[h] + (size_t)t * 2 * kImg128;
                const uint8_t* srcD = b.img.dout[h] + (size_t)t * 2 * dp;
                tma_bulk_g2s(dA2, srcA + hf * kHalf128, kHalf128, &bar_ld[st]);
                tma_bulk_g2s(dA2 + kHalf128, srcA + kImg128 + hf * kHalf128, kHalf128, &bar_ld[st]);
                tma_bulk_g2s(dDO, srcD + hf * (dp / 2), dp / 2, &bar_ld[st]);
                tma_bulk_g2s(dDO + dp / 2, srcD + dp + hf * (dp / 2), dp / 2, &bar_ld[st]);
            }
        };
        uint32_t ph_ld[2] = {0u, 0u}, ph_mma[2] = {0u, 0u};
        if (leader) load(0);
        __syncwarp();
        for (int64_t w = 0; w < nwork; ++w) {
            const int st = (int)(w & 1);
            if (w + 1 < nwork) {
                const int ns = st ^ 1;
                if (w >= 1) { mbar_wait(&bar_mma[ns], ph_mma[ns]); ph_mma[ns] ^= 1u; }   // MMAs of work item w-1 have released stage ns
                if (leader) load(w + 1);
                __syncwarp();
            }
            mbar_wait(&bar_ld[st], ph_ld[st]); ph_ld[st] ^= 1u;
            tc::fence_after_sync();
            const uint32_t base = tc::smem_addr(smem + st * kStage), ones = tc::smem_addr(sOnes);
            const uint32_t x = base, y = base + 2 * kHalf128, a2 = base + 4 * kHalf128, dd = base + 6 * kHalf128;
            if (leader) {
                if (layer0) {
                    gemm_bf16x2_ss_mn<64>(tbase + colW, x, x + kHalf128, 128, y, y + yp / 2, (uint32_t)b.F, (uint32_t)b.F, started, false);
                } else {
                    gemm_bf16x2_ss_mn<64>(tbase + colW, x, x + kHalf128, 128, y, y + kHalf128, 128, 128, started, false);
                    gemm_bf16x2_ss_mn<64>(tbase + colW2, a2, a2 + kHalf128, 128, dd, dd + dp / 2, (uint32_t)kp16, (uint32_t)kp16, started, false);
                }
                gemm_bf16x2_ss_mn<64>(tbase + colB, x, x + kHalf128, 128, ones, ones, 16, 16, started, true);
                tc::umma_commit(&bar_mma[st]);
            }
            __syncwarp();
            started = true;
        }
        if (leader) tc::umma_commit(&bar_done);        // covers every MMA issued above
        __syncwarp();
        mbar_wait(&bar_done, 0);
    }
    started = nwork > 0;
    __syncthreads();
    tc::fence_after_sync();
    if (started || true) {
        // every thread owns one output row j (its TMEM lane)
        const int j = tid;
        const bool any = t1 > t0;
        if (any) {
            const int ncolW = layer0 ? b.F : 128;
            float* gw = layer0 ? b.g_w0 : b.g_w1[h];
            for (int c0 = 0; c0 < ncolW; c0 += 16) {
                uint32_t v[16];
                tc::tmem_ld16(lane_base + colW + c0, v);
                tc::wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) atomicAdd(gw + (size_t)j * ncolW + c0 + i, __uint_as_float(v[i]));
            }
            {
                uint32_t v[16];
                tc::tmem_ld16(lane_base + colB, v);
                tc::wait_ld();
                atomicAdd((layer0 ? b.g_b0 : b.g_b1[h]) + j, __uint_as_float(v[0]));
            }
            if (!layer0) {
                const int ko = head_out(h);
                for (int c0 = 0; c0 < kp16; c0 += 16) {
                    uint32_t v[16];
                    tc::tmem_ld16(lane_base + colW2 + c0, v);
                    tc::wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < ko) atomicAdd(b.g_w2[h] + (size_t)(c0 + i) * 128 + j, __uint_as_float(v[i]));
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, tc::kTmemCols);
}

// ======================================================================================================
// host side
// ======================================================================================================
size_t tc_deform_backward_scratch_bytes(const DeformDesc& d, int64_t n) {
    const size_t ntiles = (size_t)((n + 127) / 128);
    size_t per_tile = 2 * (size_t)2 * kImg128 + (size_t)2 * 128 * d.F * 2;   // a1, dh, feat
    for (int h = 0; h < G4D_NUM_HEADS; ++h)
        if (d.head_mask & (1 << h)) per_tile += 2 * (size_t)2 * kImg128 + (size_t)2 * 128 * (h == 4 ? 48 : 16) * 2;
    size_t rows = 0;
    for (int l = 0; l < d.levels; ++l)
        for (int a = 0; a < 3; ++a) rows += (size_t)d.res[l][a] * d.C;
    return ntiles * per_tile + rows * 4 + (size_t)2 * ntiles * 128 * d.F * 4 + 16384;
}

template <int C, int L>
static cudaError_t launch_a(const BwdADesc& bd, float time, int64_t n, const float* xyz, int sm_count, cudaStream_t st) {
    const BwdSmemA Ls = bwd_smem_a(C * L, bd.d.head_mask & G4D_HEAD_SHS);
    const size_t bytes = Ls.total + 1024;
    cudaError_t e = cudaFuncSetAttribute(deform_tc_bwd_dgrad_kernel<C, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    const int64_t ntiles = (n + 127) / 128;
    const int grid = (int)(ntiles < sm_count ? ntiles : sm_count);
    deform_tc_bwd_dgrad_kernel<C, L><<<grid, 256, bytes, st>>>(bd, Ls, time, n, xyz);
    return cudaGetLastError();
}

cudaError_t launch_deform_backward_tc(const DeformDesc& d, const G4DDeformParams& prm, const G4DDeformGrads& grads,
                                      const TcBwdWeights& w, float time, int64_t n, const float* xyz,
                                      const float* const go[G4D_NUM_HEADS], float* const gi[G4D_NUM_HEADS],
                                      const uint32_t* relu_bits, const float* saved_feat, long long* dbg, uint8_t* scratch,
                                      int sm_count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const int64_t ntiles = (n + 127) / 128;
    BwdADesc a{};
    a.d = d; a.w = w; a.relu_bits = relu_bits; a.dbg = dbg;
    uint8_t* p = scratch;
    auto take = [&](size_t bytes) { uint8_t* o = p; p += (bytes + 255) & ~(size_t)255; return o; };
    a.img.feat_bytes = 2u * 128 * d.F * 2;
    a.img.a1 = take((size_t)ntiles * 2 * kImg128);
    a.img.dh = take((size_t)ntiles * 2 * kImg128);
    a.img.feat = take((size_t)ntiles * a.img.feat_bytes);
    for (int h = 0; h < G4D_NUM_HEADS; ++h) {
        a.go[h] = go[h]; a.gi[h] = gi[h]; a.g_b2[h] = grads.b2[h];
        if (!(d.head_mask & (1 << h))) continue;
        a.img.dz[h] = take((size_t)ntiles * 2 * kImg128);
        a.img.a2[h] = take((size_t)ntiles * 2 * kImg128);
        a.img.dout[h] = take((size_t)ntiles * 2 * 128 * (h == 4 ? 48 : 16) * 2);
    }
    float* feat = reinterpret_cast<float*>(take((size_t)ntiles * 128 * d.F * 4));
    float* dfeat = reinterpret_cast<float*>(take((size_t)ntiles * 128 * d.F * 4));
    a.feat = saved_feat ? saved_feat : feat; a.dfeat = dfeat;   // saved_feat: the features the forward of this view staged
    size_t row_floats = 0;
    for (int l = 0; l < d.levels; ++l)
        for (int k = 0; k < 3; ++k) row_floats += (size_t)d.res[l][k] * d.C;
    float* rows = reinterpret_cast<float*>(take(row_floats * 4));
    BwdScatterDesc sc{};
    sc.d = d; sc.dfeat = dfeat;
    {
        float* q = rows;
        for (int l = 0; l < d.levels; ++l) {
            for (int k = 0; k < 6; ++k) sc.g_planes[l][k] = grads.planes[l][k];
            for (int k = 0; k < 3; ++k) { sc.trow_grad[l][k] = q; q += (size_t)d.res[l][k] * d.C; }
        }
    }
    for (int h = 0; h < G4D_NUM_HEADS; ++h) { sc.go[h] = go[h]; sc.gi[h] = gi[h]; }
    cudaError_t e = cudaMemsetAsync(rows, 0, row_floats * 4, st);
    if (e != cudaSuccess) return e;
    const int C4 = d.C / 4;
    // gather at full occupancy (skipped when the forward's feature staging buffer was kept for this backward)
    if (!saved_feat) {
        const int64_t threads = n * C4;
        const unsigned grid = (unsigned)((threads + 255) / 256);
        if (C4 == 4) bwd_features_kernel<4><<<grid, 256, 0, st>>>(d, n, xyz, feat);
        else if (C4 == 8) bwd_features_kernel<8><<<grid, 256, 0, st>>>(d, n, xyz, feat);
        else return cudaErrorInvalidValue;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (d.C == 16 && d.levels == 2) e = launch_a<16, 2>(a, time, n, xyz, sm_count, st);
    else if (d.C == 16 && d.levels == 3) e = launch_a<16, 3>(a, time, n, xyz, sm_count, st);
    else if (d.C == 32 && d.levels == 2) e = launch_a<32, 2>(a, time, n, xyz, sm_count, st);
    else return cudaErrorInvalidValue;
    if (e != cudaSuccess) return e;
    // scatter d(feat) into the planes, d(xyz), residual copies
    {
        const int gpb = 256 / C4;
        const unsigned grid = (unsigned)((n + gpb - 1) / gpb);
        if (C4 == 4) bwd_scatter_kernel<4><<<grid, 256, 0, st>>>(sc, n, xyz);
        else bwd_scatter_kernel<8><<<grid, 256, 0, st>>>(sc, n, xyz);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    // kernel B
    BwdBDesc b{};
    b.img = a.img; b.head_mask = d.head_mask; b.F = d.F; b.ntiles = ntiles;
    {
        int nheads = 0;
        for (int h = 0; h < G4D_NUM_HEADS; ++h) nheads += (d.head_mask >> h) & 1;
        // bytes per tile: a head streams DZ + A1 + A2 (+ DOUT), layer 0 streams DH + FEAT
        const double wh = 3.0 * 64 + 16, w0 = 64 + (double)(2 * 128 * d.F * 2) / 1024.0, tot = nheads * wh + w0;
        int ch = nheads ? (int)(sm_count * wh / tot) : 0;
        if (ch < 1) ch = 1;
        int c0 = sm_count - nheads * ch;
        if (c0 < 1) c0 = 1;
        if ((int64_t)ch > ntiles) ch = (int)ntiles;
        if ((int64_t)c0 > ntiles) c0 = (int)ntiles;
        b.nheads = nheads; b.chunks_head = ch; b.chunks_l0 = c0;
    }
    b.g_w0 = grads.w0; b.g_b0 = grads.b0;
    for (int h = 0; h < G4D_NUM_HEADS; ++h) { b.g_w1[h] = grads.w1[h]; b.g_b1[h] = grads.b1[h]; b.g_w2[h] = grads.w2[h]; }
    const size_t smem_b = (size_t)2 * (3 * kImg128 + 2 * 64 * 48 * 2) + 64 * 16 * 2 + 1024;   // two half-tile stages + ones
    e = cudaFuncSetAttribute(deform_tc_bwd_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
    if (e != cudaSuccess) return e;
    deform_tc_bwd_wgrad_kernel<<<b.nheads * b.chunks_head + b.chunks_l0, 128, smem_b, st>>>(b);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    return launch_distribute_time_grad(d, sc.trow_grad, sc.g_planes, time, st);
}

}  // namespace g4d
