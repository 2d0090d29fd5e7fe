// g4d_deform_f16.cu -- round-2 tensor-core forward of the fused deform + activate + project stage (net_width 128):
// FP16x2 operands, two tiles in flight per SM.
//
// Why another arithmetic.  The 3xTF32 kernel (g4d_deform_tc.cu) spends 48 tcgen05.mma dispatches of 64 cycles per head and
// 256 TMEM columns on the (hi | lo) activation operand, which leaves room for only ONE tile per SM: its epilogue threads and
// the tensor pipe take turns.  An FP16 (hi, lo) pair carries the same 11 + 11 significand bits as a TF32 pair, kind::f16
// retires K = 16 per dispatch (half the dispatches, half the A-operand reads) and the pair packs into HALF the TMEM columns:
//   * x = hi + lo with hi = rn_f16(s x), lo = rn_f16(s x - hi), s a fixed power of two per operand class (features 2^6,
//     activations 2^3, weights 2^8; undone exactly by the epilogue's FMA).  |x - (hi + lo)/s| <= max(2^-22 |x|, 2^-25 / s):
//     the relative term is the 3xTF32 one; the absolute floor (3.7e-9 for activations, 4.7e-10 for features, 1.2e-10 for
//     weights) is far below the fp32 accumulation noise of a K = 128 dot product.  Three products (lo*hi, hi*lo, hi*hi),
//     fp32 accumulation in TMEM.
//   * range: s |x| must stay below 65504 (activations < 8188, features < 1023, weights < 255).  Conversions saturate and
//     the kernel raises a flag in host-mapped memory that the next call reports (G4D_ERR_OVERFLOW; the 3xTF32 kernel,
//     G4D_OPT_TENSOR_CORES = 1, has no such limit).
//
// Structure (512 threads = 4 warpgroups, 1 CTA / SM, persistent over tiles of 128 Gaussians = 128 TMEM lanes):
//   TMEM: two SLOTS of 256 columns: A [0,128) = activation operand (hi [0,64) | lo [64,128), two f16 per column; the feature
//         operand of layer 0 and the SH head's hidden layer overlay it) and D [128,256) = accumulator.
//   warps 0-3 / 4-7: the epilogue group of slot 0 / slot 1 -- ONE thread owns one Gaussian through every layer of the network
//         (bias, ReLU, sign bits, hi/lo split; layer 2 of the <=4-wide heads in exact fp32 with packed FFMA2).  No partial sums
//         between threads.
//   warp 8: MMA issuer.  The warp walks its loop CONVERGED and one elected lane issues, so that every tcgen05.mma operand is a
//         uniform register (measured, tools/umma_rate.py: issued from a single-lane divergent branch each UTCHMMA is wrapped
//         in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop and costs ~150 cycles; uniform issue reaches the pipe's
//         76 cycles per 128 x 128 x 16 f16 dispatch, 54 for N = 64, so heads are issued whole: N = 128).  Order per head:
//         slot 0, slot 1 -- one slot's GEMM is on the pipe while the other slot's threads run their epilogue.
//   warp 9: TMA producer (same converged / elected scheme): W1 images stream through a ring of two 64 KB parts; an image is
//         used by both slots before it is recycled, so the L2 -> smem weight traffic is one image per TWO tiles.
//   warps 12-15: finisher group -- the per-Gaussian tail of every tile (residual adds, activations, EWA projection, SH colour,
//         record stores) runs here, off the network's critical path: the epilogue threads hand over 11 deltas through shared
//         memory, the 48 SH deltas are read straight out of the layer-2 accumulator in TMEM.
// Replaces scene/deformation.py:67-148 + gaussian_renderer/__init__.py:97-99 + the rasterizer's preprocess (same contract
// as g4d_deform_tc.cu).  Compiled with -fmad=false (projection maths, g4d_math.cuh); the FMAs below are explicit.
#include "geom_finish.cuh"
#include "tc_umma.cuh"

namespace g4d {

namespace {

constexpr float kSF = 64.f, kSA = 8.f, kSW = 256.f;        // operand scales (powers of two)
constexpr float kInvL0 = kSA / (kSF * kSW);                 // layer-0 accumulator -> activation already scaled by kSA
constexpr float kInvH = 1.f / (kSA * kSW);                  // head accumulators -> value
constexpr float kInvHS = 1.f / kSW;                         // SH head layer 1 -> hidden activation already scaled by kSA
constexpr float kF16Max = 65504.f;
constexpr int kF16Threads = 512;   // 4 warpgroups: epilogue slot 0 | epilogue slot 1 | MMA, TMA (+2 idle warps) | finisher
constexpr uint32_t kSlotCols = 256, kColA = 0, kColALo = 64, kColDd = 128;
constexpr uint32_t kPartBytes = 2u * 128u * 128u * 2u;     // one head's W1 image: (hi | lo) x 128 rows x 128 K x f16 = 64 KB
constexpr int kRing = 2;

struct F16Smem { uint32_t w1, w0, w2, bias, w2s, out, bars, total; };

F16Smem f16_smem_layout(int F, bool sh) {
    F16Smem s{};
    uint32_t off = 0;
    auto take = [&](uint32_t bytes) { uint32_t o = off; off += (bytes + 127u) & ~127u; return o; };
    s.w1 = take(kRing * kPartBytes);
    s.w0 = take(2u * 128 * F * 2);
    s.w2 = take(sh ? 2u * 48 * 128 * 2 : 128u);
    s.bias = take((128 + G4D_NUM_HEADS * 128 + 64) * 4);
    s.w2s = take(4 * 64 * 32);
    s.out = take(2 * 11 * 128 * 4);
    s.bars = take(256);
    s.total = off;
    return s;
}

__device__ __forceinline__ void mbar_arrive_(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// two values -> packed f16x2 (hi) and packed f16x2 of the residuals (lo); lower half = first value
__device__ __forceinline__ void f16_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    float ha, hb;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(ha), "=f"(hb) : "r"(hi));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - hb), "f"(a - ha));
}

__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);   // D = F32, A = B = F16 (format 0), both K-major
}

// D[128 x N] (+)= A[128 x K] * B[N x K]^T as three f16 products; A (hi, lo) packed two-per-column in TMEM, B (hi, lo)
// images in shared memory ([N][b_ncols] 8x8-core layout of tc_umma.cuh, K-major).  ONE thread.
template <int K>
__device__ __forceinline__ void gemm_f16x2(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t N,
                                           uint32_t b_ncols) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t sbo = (b_ncols >> 3) * 128u;
    const uint64_t bd_hi = tc::make_smem_desc(b_hi, 128u, sbo), bd_lo = tc::make_smem_desc(b_lo, 128u, sbo);
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const uint32_t a = (p == 0) ? a_lo : a_hi;
        const uint64_t bd = (p == 1) ? bd_lo : bd_hi;
#pragma unroll
        for (int ks = 0; ks < K; ks += 16)
            tc::umma_bf16_ts(d_tmem, a + (ks >> 1), bd + (uint64_t)((ks >> 4) * 16), idesc, p > 0 || ks > 0);   // kind::f16
    }
}

// ---- weight images -------------------------------------------------------------------------------------------------
struct F16PackDesc {
    const float* src[1 + 2 * G4D_NUM_HEADS];
    uint8_t* dst[1 + 2 * G4D_NUM_HEADS];
    int rows_src[1 + 2 * G4D_NUM_HEADS], rows_dst[1 + 2 * G4D_NUM_HEADS], K[1 + 2 * G4D_NUM_HEADS], halves[1 + 2 * G4D_NUM_HEADS];
    int start[2 + 2 * G4D_NUM_HEADS];
    int count;
};

__global__ void f16_pack_weights_kernel(F16PackDesc p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.start[p.count]) return;
    int m = 0;
    while (i >= p.start[m + 1]) ++m;
    const int e = i - p.start[m];
    const int K = p.K[m];
    const uint32_t n = e / K, k = e % K;
    const float v = ((int)n < p.rows_src[m] ? __ldg(p.src[m] + n * K + k) : 0.f) * kSW;
    uint32_t hi, lo;
    f16_split2(v, 0.f, hi, lo);
    uint8_t* base = p.dst[m];
    const uint32_t rows = p.rows_dst[m], nn = n;
    const uint32_t off = tc::img16_off(nn, k, K);
    *reinterpret_cast<uint16_t*>(base + off) = (uint16_t)hi;
    *reinterpret_cast<uint16_t*>(base + (size_t)rows * K * 2 + off) = (uint16_t)lo;
}


}  // namespace

cudaError_t launch_f16_pack_weights(const G4DDeformParams& prm, float* blob, TcWeights* out, cudaStream_t st) {
    F16PackDesc p{};
    const int F = prm.levels * prm.channels;
    int m = 0, total = 0;
    uint8_t* q = reinterpret_cast<uint8_t*>(blob);
    auto add = [&](const float* src, int rows_src, int rows_dst, int K, int halves) {
        p.src[m] = src; p.dst[m] = q; p.rows_src[m] = rows_src; p.rows_dst[m] = rows_dst; p.K[m] = K; p.halves[m] = halves;
        p.start[m] = total;
        total += rows_dst * K;
        uint8_t* r = q; q += 2 * (size_t)rows_dst * K * 2; ++m;
        return reinterpret_cast<const float*>(r);
    };
    out->w0 = add(prm.w0, 128, 128, F, 0);
    for (int h = 0; h < G4D_NUM_HEADS; ++h) {
        out->kp16[h] = h == 4 ? 48 : 16;
        out->w1[h] = nullptr; out->w2[h] = nullptr;
        if (!(prm.head_mask & (1 << h))) continue;
        out->w1[h] = add(prm.w1[h], 128, 128, 128, 0);
        if (h == 4) out->w2[h] = add(prm.w2[h], 48, 48, 128, 0);   // the small heads' layer 2 runs in fp32 from the caller's tensors
    }
    p.start[m] = total; p.count = m;
    f16_pack_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(p);
    return cudaGetLastError();
}

namespace {

// ---- epilogue helpers (one thread = one TMEM lane) --------------------------------------------------------------------
// Walks NCH chunks of 16 accumulator columns starting at `col`, double-buffered: the tcgen05.ld of chunk c + 1 is in
// flight while chunk c is processed.  f(c, v) with c a compile-time-resolvable index after unrolling.
template <int NCH, class Fn>
__device__ __forceinline__ void for_chunks(uint32_t col, Fn&& f) {
    uint32_t va[16], vb[16];
    tc::tmem_ld16(col, va);
    tc::wait_ld();
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        if (c + 1 < NCH) {
            if (c & 1) tc::tmem_ld16(col + (uint32_t)(c + 1) * 16u, va); else tc::tmem_ld16(col + (uint32_t)(c + 1) * 16u, vb);
        }
        if (c & 1) f(c, vb); else f(c, va);
        if (c + 1 < NCH) tc::wait_ld();
    }
}

// ReLU + split of 16 pre-scaled pre-activations -> 8 + 8 packed operand words; sign bits; running maximum
template <bool SAVE>
__device__ __forceinline__ void relu_split16(const uint32_t (&v)[16], float inv, const float* __restrict__ bias, uint32_t (&hi)[8],
                                             uint32_t (&lo)[8], uint32_t& bits, float& mx) {
    bits = 0;
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias + j);
        const float2 x01 = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), make_float2(inv, inv), make_float2(b4.x, b4.y));
        const float2 x23 = __ffma2_rn(make_float2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), make_float2(inv, inv), make_float2(b4.z, b4.w));
        if (SAVE) {
            bits |= (x01.x > 0.f ? 1u : 0u) << j; bits |= (x01.y > 0.f ? 1u : 0u) << (j + 1);
            bits |= (x23.x > 0.f ? 1u : 0u) << (j + 2); bits |= (x23.y > 0.f ? 1u : 0u) << (j + 3);
        }
        const float a0 = fmaxf(x01.x, 0.f), a1 = fmaxf(x01.y, 0.f), a2 = fmaxf(x23.x, 0.f), a3 = fmaxf(x23.y, 0.f);
        mx = fmaxf(mx, fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)));
        f16_split2(a0, a1, hi[j >> 1], lo[j >> 1]);
        f16_split2(a2, a3, hi[(j >> 1) + 1], lo[(j >> 1) + 1]);
    }
}

// One small head for one Gaussian: a2 = relu(acc * inv + b1), layer 2 (KO <= 4 outputs) in exact fp32 on hidden-unit PAIRS with
// packed FMAs: acc[o] = (sum over even units, sum over odd units) of a2 * W2[o][unit]; two accumulator sets alternate.
// w2p: [64 unit pairs][2] float4 = (W2[0][e], W2[0][o], W2[1][e], W2[1][o]), (W2[2][e], W2[2][o], W2[3][e], W2[3][o]).
template <int KO, bool SAVE>
__device__ __forceinline__ void small_head(uint32_t col, const float* __restrict__ b1, const float4* __restrict__ w2p,
                                           float2 (&accA)[4], float2 (&accB)[4], uint32_t (&rb)[4]) {
    for_chunks<8>(col, [&](int c, const uint32_t (&v)[16]) {
        uint32_t bits = 0;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(b1 + c * 16 + j);
            float2 x01 = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), make_float2(kInvH, kInvH), make_float2(b4.x, b4.y));
            float2 x23 = __ffma2_rn(make_float2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), make_float2(kInvH, kInvH), make_float2(b4.z, b4.w));
            if (SAVE) {
                bits |= (x01.x > 0.f ? 1u : 0u) << j; bits |= (x01.y > 0.f ? 1u : 0u) << (j + 1);
                bits |= (x23.x > 0.f ? 1u : 0u) << (j + 2); bits |= (x23.y > 0.f ? 1u : 0u) << (j + 3);
            }
            x01.x = fmaxf(x01.x, 0.f); x01.y = fmaxf(x01.y, 0.f); x23.x = fmaxf(x23.x, 0.f); x23.y = fmaxf(x23.y, 0.f);
            const float4* w = w2p + (c * 8 + (j >> 1)) * 2;       // pair index (c*16 + j) / 2, two float4 per pair
            {
                const float4 q0 = w[0];
                accA[0] = __ffma2_rn(x01, make_float2(q0.x, q0.y), accA[0]);
                if (KO > 1) accA[1] = __ffma2_rn(x01, make_float2(q0.z, q0.w), accA[1]);
                if (KO > 2) {
                    const float4 q1 = w[1];
                    accA[2] = __ffma2_rn(x01, make_float2(q1.x, q1.y), accA[2]);
                    if (KO > 3) accA[3] = __ffma2_rn(x01, make_float2(q1.z, q1.w), accA[3]);
                }
            }
            {
                const float4 q0 = w[2];
                accB[0] = __ffma2_rn(x23, make_float2(q0.x, q0.y), accB[0]);
                if (KO > 1) accB[1] = __ffma2_rn(x23, make_float2(q0.z, q0.w), accB[1]);
                if (KO > 2) {
                    const float4 q1 = w[3];
                    accB[2] = __ffma2_rn(x23, make_float2(q1.x, q1.y), accB[2]);
                    if (KO > 3) accB[3] = __ffma2_rn(x23, make_float2(q1.z, q1.w), accB[3]);
                }
            }
        }
        if (SAVE) rb[c >> 1] |= bits << ((c & 1) * 16);
    });
}

}  // namespace

// mbarrier indices
//   0 w0 | 1,2 full[part] | 3,4 hfree[part] | per slot s (base 5 + 9 s): +0 feat (128) | +1 a1 (128) | +2 x (128) | +3 l0full
//   | +4 l2full | +5 dfull | +6 dfree (128) | +7 ofull (128: deltas handed to the finisher) | +8 ofree (128: read by it)
constexpr int kBarW0 = 0, kBarFull = 1, kBarHFree = 3, kBarSlot = 5, kBarPerSlot = 9;

template <int MODE, int C, int L, bool SAVE>
__global__ void __launch_bounds__(kF16Threads, 1)
deform_f16_kernel(DeformDesc d, TcWeights tw, F16Smem Ls, const CameraDev* __restrict__ camp, int use_cam, int64_t n, DeformIO io) {
    constexpr int F = C * L;
    static_assert(F % 16 == 0 && F <= 64, "feature width must be 32, 48 or 64");
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ CameraDev cam;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & 127;
    const int64_t ntiles = (n + 127) / 128;
    // (programmatic dependent launch: everything up to pdl_wait() below reads only the network's own parameters -- written long
    //  before the previous two kernels of the stream -- and sets up shared memory, TMEM and the barriers while the HexPlane
    //  gather in front of this kernel drains)
    float* sBias = reinterpret_cast<float*>(smem + Ls.bias);     // b0 * kSA [128] | b1[5][128] (SH head * kSA) | b2s[4][4] | b2sh[48]
    float4* sW2p = reinterpret_cast<float4*>(smem + Ls.w2s);     // [4 small heads][64 unit pairs][2] (small_head)
    float* sOut = reinterpret_cast<float*>(smem + Ls.out);       // [2 slots][11 deltas][128 Gaussians]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Ls.bars);
    for (int i = tid; i < 128; i += kF16Threads) sBias[i] = __ldg(d.b0 + i) * kSA;
    for (int i = tid; i < 64; i += kF16Threads) sBias[128 + G4D_NUM_HEADS * 128 + i] = 0.f;
    for (int h = 0; h < G4D_NUM_HEADS; ++h) {
        if (!(d.head_mask & (1 << h))) continue;
        for (int i = tid; i < 128; i += kF16Threads) sBias[128 + h * 128 + i] = __ldg(d.b1[h] + i) * (h == 4 ? kSA : 1.f);
        if (h < 4) {
            const int ko = head_out(h);
            for (int pq = tid; pq < 64; pq += kF16Threads) {
                const int e = 2 * pq, o = 2 * pq + 1;
                const float* w = d.w2[h];
                sW2p[(h * 64 + pq) * 2] = make_float4(__ldg(w + e), __ldg(w + o), ko > 1 ? __ldg(w + 128 + e) : 0.f, ko > 1 ? __ldg(w + 128 + o) : 0.f);
                sW2p[(h * 64 + pq) * 2 + 1] = make_float4(ko > 2 ? __ldg(w + 256 + e) : 0.f, ko > 2 ? __ldg(w + 256 + o) : 0.f,
                                                         ko > 3 ? __ldg(w + 384 + e) : 0.f, ko > 3 ? __ldg(w + 384 + o) : 0.f);
            }
        }
    }
    __syncthreads();      // (the zero fill of the b2 block above must land before the per-head values)
    for (int h = 0; h < G4D_NUM_HEADS; ++h) {
        if (!(d.head_mask & (1 << h))) continue;
        for (int i = tid; i < head_out(h); i += kF16Threads) sBias[128 + G4D_NUM_HEADS * 128 + (h < 4 ? 4 * h : 16) + i] = __ldg(d.b2[h] + i);
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, tc::kTmemCols);
    if (tid == 0) {
        for (int i = 0; i < kBarSlot; ++i) mbar_init(bars + i, 1);
        for (int s = 0; s < 2; ++s) {
            uint64_t* b = bars + kBarSlot + s * kBarPerSlot;
            mbar_init(b + 0, 128); mbar_init(b + 1, 128); mbar_init(b + 2, 128);
            mbar_init(b + 3, 1); mbar_init(b + 4, 1); mbar_init(b + 5, 1);
            mbar_init(b + 6, 128); mbar_init(b + 7, 128); mbar_init(b + 8, 128);
        }
        fence_barrier_init();
    }
    pdl_wait();           // from here on: the camera, the staged features, the weight images
    pdl_trigger();
    if (use_cam) {
        for (int i = tid; i < (int)(sizeof(CameraDev) / 4); i += kF16Threads)
            reinterpret_cast<uint32_t*>(&cam)[i] = reinterpret_cast<const uint32_t*>(camp)[i];
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    // warp-uniform for the compiler (CUTLASS's canonical-warp-idx trick): tcgen05.mma takes its operands from UNIFORM registers;
    // a per-thread value (an LDS result) would make every issue a waterfall loop (ELECT / R2UR.BROADCAST / BRA.U.ANY)
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base_s, 0);
    const bool hsh = d.head_mask & G4D_HEAD_SHS;
    const uint32_t m_heads = (uint32_t)__popc(d.head_mask & 31);
    const int64_t my_tiles = (int64_t)blockIdx.x < ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const uint32_t npairs = (uint32_t)((my_tiles + 1) / 2);
    // head id of the k-th head: the small heads in an order ROTATED by the CTA index (the 148 CTAs then stream four different
    // W1 images out of L2 at any moment), the SH head always last (its hidden layer reuses the dead A operand)
    uint32_t hseq_pack = 0;
    {
        const int ns = __popc(d.head_mask & 15);
        const int rot = ns ? (int)(blockIdx.x % (unsigned)ns) : 0;
        int pos = 0;
        for (int h = 0; h < 4; ++h) {
            if (!(d.head_mask & (1 << h))) continue;
            const int k = pos >= rot ? pos - rot : pos - rot + ns;
            hseq_pack |= (uint32_t)h << (3 * k);
            ++pos;
        }
        if (hsh) hseq_pack |= 4u << (3 * ns);
    }
    auto head_at = [&](uint32_t k) { return (int)((hseq_pack >> (3u * k)) & 7u); };
    const uint32_t sW1 = tc::smem_addr(smem + Ls.w1), sW2 = tc::smem_addr(smem + Ls.w2), sW0 = tc::smem_addr(smem + Ls.w0);
    auto slot_bar = [&](int s, int i) { return bars + kBarSlot + s * kBarPerSlot + i; };

    if (warp == 9) {
        // ============================================================================================================
        // TMA producer.  The whole warp walks the loop CONVERGED and one elected lane issues: operands then live in uniform
        // registers (a single-lane divergent branch turns every issue into an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop)
        // ============================================================================================================
        const bool leader = tc::elect_one_sync();
        const uint32_t w2b = hsh ? 2u * 48 * 128 * 2 : 0u;
        if (leader) {
            mbar_expect_tx(bars + kBarW0, 2u * 128 * F * 2 + w2b);
            tma_bulk_g2s(smem + Ls.w0, tw.w0, 2u * 128 * F * 2, bars + kBarW0);
            if (hsh) tma_bulk_g2s(smem + Ls.w2, tw.w2[4], w2b, bars + kBarW0);
        }
        __syncwarp();
        const uint32_t total_q = npairs * m_heads;
        uint32_t k = 0;
        for (uint32_t q = 0; q < total_q; ++q) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(tw.w1[head_at(k)]);
            const uint32_t part = q & 1u;
            if (q >= 2) mbar_wait(bars + kBarHFree + part, ((q >> 1) - 1u) & 1u);
            if (leader) {
                mbar_expect_tx(bars + kBarFull + part, kPartBytes);
                tma_bulk_g2s(smem + Ls.w1 + part * kPartBytes, src, kPartBytes, bars + kBarFull + part);
            }
            __syncwarp();
            if (++k == m_heads) k = 0;
        }
    } else if (warp == 8) {
        // ============================================================================================================
        // MMA issuer (converged walk, one elected issuer).  Order per head: slot 0, slot 1 -- while one slot's GEMM is on the
        // tensor pipe the other slot's threads run the epilogue of theirs.
        // ============================================================================================================
        const bool leader = tc::elect_one_sync();
        long long t_wait = 0, t0 = clock64();
        auto wait = [&](uint64_t* bar, uint32_t parity) {
            if (tw.dbg) { const long long a = clock64(); mbar_wait(bar, parity); t_wait += clock64() - a; }
            else mbar_wait(bar, parity);
        };
        wait(bars + kBarW0, 0);
        uint32_t nfree[2] = {0u, 0u};
        uint32_t q = 0;
        for (uint32_t pr = 0; pr < npairs; ++pr) {
            const int nact = (int64_t)(2 * pr + 1) < my_tiles ? 2 : 1;
            const uint32_t par = pr & 1u;
            // ---- layer 0: D = X W0^T
            for (int s = 0; s < nact; ++s) {
                wait(slot_bar(s, 0), par);
                if (pr > 0) {
                    if (m_heads) { wait(slot_bar(s, 6), nfree[s] & 1u); ++nfree[s]; }
                    else wait(slot_bar(s, 1), par ^ 1u);
                }
                tc::fence_after_sync();
                const uint32_t sb = tbase + (uint32_t)s * kSlotCols;
                if (leader) {
                    gemm_f16x2<F>(sb + kColDd, sb + kColA, sb + kColALo, sW0, sW0 + 128u * F * 2u, 128, F);
                    tc::umma_commit(slot_bar(s, 3));
                }
                __syncwarp();
            }
            for (uint32_t k = 0; k < m_heads; ++k, ++q) {
                const uint32_t part = q & 1u;
                wait(bars + kBarFull + part, (q >> 1) & 1u);
                const uint32_t bw = sW1 + part * kPartBytes;
                for (int s = 0; s < nact; ++s) {
                    if (k == 0) wait(slot_bar(s, 1), par);            // a1 written, D drained
                    else { wait(slot_bar(s, 6), nfree[s] & 1u); ++nfree[s]; }
                    tc::fence_after_sync();
                    const uint32_t sb = tbase + (uint32_t)s * kSlotCols;
                    if (leader) {
                        gemm_f16x2<128>(sb + kColDd, sb + kColA, sb + kColALo, bw, bw + kPartBytes / 2, 128, 128);
                        tc::umma_commit(slot_bar(s, 5));
                    }
                    __syncwarp();
                }
                if (leader) tc::umma_commit(bars + kBarHFree + part);
                __syncwarp();
                if (hsh && k + 1 == m_heads) {
                    // SH head, layer 2: D[:, 0:48) = a2 W2^T
                    for (int s = 0; s < nact; ++s) {
                        wait(slot_bar(s, 2), par);
                        tc::fence_after_sync();
                        const uint32_t sb = tbase + (uint32_t)s * kSlotCols;
                        if (leader) {
                            gemm_f16x2<128>(sb + kColDd, sb + kColA, sb + kColALo, sW2, sW2 + 48u * 128 * 2, 48, 128);
                            tc::umma_commit(slot_bar(s, 4));
                        }
                        __syncwarp();
                    }
                }
            }
        }
        if (tw.dbg && leader) { tw.dbg[blockIdx.x * 12 + 10] = t_wait; tw.dbg[blockIdx.x * 12 + 11] = clock64() - t0; }
    } else if (warp < 8) {
        // ============================================================================================================
        // epilogue groups: slot = warp / 4, one thread per Gaussian of the slot's tile, through every layer of the network
        // ============================================================================================================
        const int slot = warp >> 2;
        const uint32_t lane_base = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)slot * kSlotCols;
        uint64_t *bar_feat = slot_bar(slot, 0), *bar_a1 = slot_bar(slot, 1), *bar_x = slot_bar(slot, 2), *bar_l0 = slot_bar(slot, 3),
                 *bar_l2 = slot_bar(slot, 4), *bar_dfull = slot_bar(slot, 5), *bar_dfree = slot_bar(slot, 6),
                 *bar_ofull = slot_bar(slot, 7), *bar_ofree = slot_bar(slot, 8);
        float* myOut = sOut + slot * 11 * 128 + row;
        float mx = 0.f;     // running maximum of every value converted to f16 (range check)
        const bool dbgt = tw.dbg && tid == 0;
        long long cyc[7] = {0, 0, 0, 0, 0, 0, 0};
        long long tprev = clock64();
#define G4D_CYC(i) do { if (dbgt) { const long long tn_ = clock64(); cyc[i] += tn_ - tprev; tprev = tn_; } } while (0)

        // features of a tile ([N][F] fp32, deform_features_kernel) -> layer-0 A operand (overlays A: no MMA reads A now)
        auto stage_features = [&](int64_t tl) {
            const int64_t gs = tl * 128 + row;
            const float4* src = reinterpret_cast<const float4*>(tw.feat + gs * F);
#pragma unroll
            for (int c0 = 0; c0 < F; c0 += 16) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (gs < n) t4 = __ldg(src + ((c0 + j) >> 2));
                    t4.x *= kSF; t4.y *= kSF; t4.z *= kSF; t4.w *= kSF;
                    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(t4.x), fabsf(t4.y)), fmaxf(fabsf(t4.z), fabsf(t4.w))));
                    f16_split2(t4.x, t4.y, hi[j >> 1], lo[j >> 1]);
                    f16_split2(t4.z, t4.w, hi[(j >> 1) + 1], lo[(j >> 1) + 1]);
                }
                tc::tmem_st8(lane_base + kColA + (uint32_t)(c0 >> 1), hi);
                tc::tmem_st8(lane_base + kColALo + (uint32_t)(c0 >> 1), lo);
            }
            tc::wait_st();
            tc::fence_before_sync();
            mbar_arrive_(bar_feat);
        };
        if (slot < my_tiles) stage_features((int64_t)blockIdx.x + (int64_t)slot * gridDim.x);

        uint32_t nd = 0;      // head GEMMs consumed by this thread (parity of dfull)
        uint32_t it = 0;
        for (int64_t i = slot; i < my_tiles; i += 2, ++it) {
            const int64_t tile = (int64_t)blockIdx.x + i * gridDim.x;
            const int64_t gi = tile * 128 + row;
            const bool valid = gi < n;
            const uint32_t par = it & 1u;
            // ---- epilogue 0: a1 = relu(D / s + b0) -> A (hi | lo), pre-scaled by kSA
            mbar_wait(bar_l0, par);
            tc::fence_after_sync();
            G4D_CYC(0);   // wait layer-0 GEMM
            {
                uint32_t rb[4] = {0u, 0u, 0u, 0u};
                for_chunks<8>(lane_base + kColDd, [&](int c, const uint32_t (&v)[16]) {
                    uint32_t hi[8], lo[8], bits;
                    relu_split16<SAVE>(v, kInvL0, sBias + c * 16, hi, lo, bits, mx);
                    if (SAVE) rb[c >> 1] |= bits << ((c & 1) * 16);
                    tc::tmem_st8(lane_base + kColA + (uint32_t)c * 8u, hi);
                    tc::tmem_st8(lane_base + kColALo + (uint32_t)c * 8u, lo);
                });
                if (SAVE && valid) *reinterpret_cast<uint4*>(tw.relu_bits + (size_t)gi * 4) = make_uint4(rb[0], rb[1], rb[2], rb[3]);
            }
            tc::wait_st();
            tc::fence_before_sync();
            mbar_arrive_(bar_a1);          // A complete, D drained
            G4D_CYC(1);   // epilogue 0

            float dl[11];
#pragma unroll
            for (int j = 0; j < 11; ++j) dl[j] = 0.f;
#pragma unroll 1
            for (uint32_t kact = 0; kact < m_heads; ++kact, ++nd) {
                const int h = head_at(kact);
                mbar_wait(bar_dfull, nd & 1u);
                tc::fence_after_sync();
                G4D_CYC(2);   // wait head GEMM
                if (h < 4) {
                    float2 accA[4], accB[4];
#pragma unroll
                    for (int o = 0; o < 4; ++o) { accA[o] = make_float2(0.f, 0.f); accB[o] = make_float2(0.f, 0.f); }
                    uint32_t rb[4] = {0u, 0u, 0u, 0u};
                    const uint32_t col = lane_base + kColDd;
                    const float* b1 = sBias + 128 + h * 128;
                    const float4* w2p = sW2p + h * 128;
                    if (h == 2) small_head<4, SAVE>(col, b1, w2p, accA, accB, rb);
                    else if (h == 3) small_head<1, SAVE>(col, b1, w2p, accA, accB, rb);
                    else small_head<3, SAVE>(col, b1, w2p, accA, accB, rb);
                    tc::fence_before_sync();
                    mbar_arrive_(bar_dfree);   // D is read: the MMA warp may overwrite it
                    if (SAVE && valid) *reinterpret_cast<uint4*>(tw.relu_bits + ((size_t)(1 + h) * (size_t)n + (size_t)gi) * 4) = make_uint4(rb[0], rb[1], rb[2], rb[3]);
                    const float4 b2 = *reinterpret_cast<const float4*>(sBias + 128 + G4D_NUM_HEADS * 128 + 4 * h);
                    float o4[4];
#pragma unroll
                    for (int o = 0; o < 4; ++o) o4[o] = (accA[o].x + accA[o].y) + (accB[o].x + accB[o].y);
                    if (h == 0) { dl[0] = o4[0] + b2.x; dl[1] = o4[1] + b2.y; dl[2] = o4[2] + b2.z; }
                    else if (h == 1) { dl[3] = o4[0] + b2.x; dl[4] = o4[1] + b2.y; dl[5] = o4[2] + b2.z; }
                    else if (h == 2) { dl[6] = o4[0] + b2.x; dl[7] = o4[1] + b2.y; dl[8] = o4[2] + b2.z; dl[9] = o4[3] + b2.w; }
                    else { dl[10] = o4[0] + b2.x; }
                    G4D_CYC(3);   // small-head epilogue
                } else {
                    // ---- SH head (always last): hidden layer -> A (a1 is dead once this head's GEMM has retired); the 48-wide
                    //      output is read out of D by the finisher warps
                    uint32_t rb[4] = {0u, 0u, 0u, 0u};
                    for_chunks<8>(lane_base + kColDd, [&](int c, const uint32_t (&v)[16]) {
                        uint32_t hi[8], lo[8], bits;
                        relu_split16<SAVE>(v, kInvHS, sBias + 128 + 4 * 128 + c * 16, hi, lo, bits, mx);
                        if (SAVE) rb[c >> 1] |= bits << ((c & 1) * 16);
                        tc::tmem_st8(lane_base + kColA + (uint32_t)c * 8u, hi);
                        tc::tmem_st8(lane_base + kColALo + (uint32_t)c * 8u, lo);
                    });
                    if (SAVE && valid) *reinterpret_cast<uint4*>(tw.relu_bits + ((size_t)(1 + h) * (size_t)n + (size_t)gi) * 4) = make_uint4(rb[0], rb[1], rb[2], rb[3]);
                    tc::wait_st();
                    tc::fence_before_sync();
                    mbar_arrive_(bar_x);            // hidden layer written (and D read): layer 2 may go
                    G4D_CYC(4);   // SH hidden epilogue
                }
            }
            // ---- hand the 11 small-head deltas of this Gaussian to the finisher warps
            if (it > 0) mbar_wait(bar_ofree, (it - 1u) & 1u);     // (the previous tile's have been read)
#pragma unroll
            for (int j = 0; j < 11; ++j) myOut[j * 128] = dl[j];
            mbar_arrive_(bar_ofull);
            G4D_CYC(5);   // hand-over
            // ---- A is free once every MMA that read it has retired: stage the features of this slot's next tile
            if (i + 2 < my_tiles) {
                if (hsh) { mbar_wait(bar_l2, par); tc::fence_after_sync(); }
                stage_features(tile + 2 * (int64_t)gridDim.x);
            }
            G4D_CYC(6);   // wait SH layer 2 + feature staging
        }
        // range check of everything that went through an f16 conversion (already in operand units)
        if (mx >= kF16Max && tw.status) *reinterpret_cast<volatile uint32_t*>(tw.status) = 1u;
        if (dbgt) {
            for (int k = 0; k < 7; ++k) tw.dbg[blockIdx.x * 12 + k] = cyc[k];
        }
#undef G4D_CYC
    } else if (warp >= 12) {
        // ============================================================================================================
        // finisher warps: the per-Gaussian tail of every tile (both slots) off the network's critical path -- residual adds,
        // activations, EWA projection, SH colour, record stores.  Thread r finishes Gaussian r of the tile.
        // ============================================================================================================
        const uint32_t lane_q = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        const float* b2sh = sBias + 128 + G4D_NUM_HEADS * 128 + 16;
        const bool dbgt = tw.dbg && tid == 384;
        long long cyc[3] = {0, 0, 0};
        long long tprev = clock64();
#define G4D_CYC(i) do { if (dbgt) { const long long tn_ = clock64(); cyc[i] += tn_ - tprev; tprev = tn_; } } while (0)
        for (int64_t i = 0; i < my_tiles; ++i) {
            const int s = (int)(i & 1);
            const uint32_t par = (uint32_t)(i >> 1) & 1u;
            const int64_t tile = (int64_t)blockIdx.x + i * gridDim.x;
            const int64_t gi = tile * 128 + row;
            const bool valid = gi < n;
            Vec3 p{0.f, 0.f, 0.f};
            float sl[3] = {0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f}, ol = 0.f;
            if (valid) {
                p = Vec3{io.xyz[3 * gi], io.xyz[3 * gi + 1], io.xyz[3 * gi + 2]};
                if (io.scaling) { sl[0] = io.scaling[3 * gi]; sl[1] = io.scaling[3 * gi + 1]; sl[2] = io.scaling[3 * gi + 2]; }
                if (io.rotation) { const float4 r4 = *reinterpret_cast<const float4*>(io.rotation + 4 * gi); q[0] = r4.x; q[1] = r4.y; q[2] = r4.z; q[3] = r4.w; }
                if (io.opacity) ol = io.opacity[gi];
            }
            mbar_wait(slot_bar(s, 7), par);
            G4D_CYC(0);   // wait deltas
            {
                const float* o = sOut + s * 11 * 128 + row;
                p.x += o[0]; p.y += o[128]; p.z += o[256];
                sl[0] += o[384]; sl[1] += o[512]; sl[2] += o[640];
                q[0] += o[768]; q[1] += o[896]; q[2] += o[1024]; q[3] += o[1152];
                ol += o[1280];
            }
            mbar_arrive_(slot_bar(s, 8));
            float dsh[48];
            if (hsh) {
                mbar_wait(slot_bar(s, 4), par);
                tc::fence_after_sync();
                for_chunks<3>(lane_q + (uint32_t)s * kSlotCols + kColDd, [&](int c, const uint32_t (&v)[16]) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) dsh[c * 16 + j] = fmaf(__uint_as_float(v[j]), kInvH, b2sh[c * 16 + j]);
                });
                tc::fence_before_sync();
                mbar_arrive_(slot_bar(s, 6));        // D[0,48) read out: the slot's next GEMM may overwrite it
            } else {
#pragma unroll
                for (int j = 0; j < 48; ++j) dsh[j] = 0.f;
            }
            G4D_CYC(1);   // SH output
            if (valid) {
                if (MODE == 0) {
                    io.out_xyz[3 * gi] = p.x; io.out_xyz[3 * gi + 1] = p.y; io.out_xyz[3 * gi + 2] = p.z;
                    if (io.out_scaling) { io.out_scaling[3 * gi] = sl[0]; io.out_scaling[3 * gi + 1] = sl[1]; io.out_scaling[3 * gi + 2] = sl[2]; }
                    if (io.out_rotation) *reinterpret_cast<float4*>(io.out_rotation + 4 * gi) = make_float4(q[0], q[1], q[2], q[3]);
                    if (io.out_opacity) io.out_opacity[gi] = ol;
                    if (io.out_shs && hsh) {
#pragma unroll
                        for (int j = 0; j < 48; j += 4) {
                            const float4 b = *reinterpret_cast<const float4*>(io.shs + gi * 48 + j);
                            *reinterpret_cast<float4*>(io.out_shs + gi * 48 + j) = make_float4(b.x + dsh[j], b.y + dsh[j + 1], b.z + dsh[j + 2], b.w + dsh[j + 3]);
                        }
                    }
                } else {
                    fused_finish_geometry(cam, io, gi, p, sl, q, ol);
                    fused_finish_colour(cam, io, gi, p, hsh, dsh);
                }
            }
            G4D_CYC(2);   // activations + projection + colour
        }
        if (dbgt) {
            for (int k = 0; k < 3; ++k) tw.dbg[blockIdx.x * 12 + 7 + k] = cyc[k];
        }
#undef G4D_CYC
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, tc::kTmemCols);
}

bool f16_deform_supported(const DeformDesc& d) {
    if (d.WD != 128) return false;
    if (!((d.C == 16 && (d.levels == 2 || d.levels == 3)) || (d.C == 32 && d.levels == 2))) return false;
    return f16_smem_layout(d.F, (d.head_mask & G4D_HEAD_SHS) != 0).total + 1024 <= 227 * 1024;
}

template <int MODE, int C, int L>
static cudaError_t launch_deform_f16_t(const DeformDesc& d, const TcWeights& tw, const F16Smem& Ls, size_t bytes, int grid,
                                       const CameraDev* cam, bool use_cam, int64_t n, const DeformIO& io, cudaStream_t st) {
    cudaError_t e;
    if (tw.relu_bits) {
        e = cudaFuncSetAttribute(deform_f16_kernel<MODE, C, L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        return launch_k(deform_f16_kernel<MODE, C, L, true>, dim3(grid), dim3(kF16Threads), bytes, st, true, d, tw, Ls, cam, use_cam ? 1 : 0, n, io);
    } else {
        e = cudaFuncSetAttribute(deform_f16_kernel<MODE, C, L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        return launch_k(deform_f16_kernel<MODE, C, L, false>, dim3(grid), dim3(kF16Threads), bytes, st, true, d, tw, Ls, cam, use_cam ? 1 : 0, n, io);
    }
}

cudaError_t launch_deform_features(const DeformDesc& d, int64_t n, const float* xyz, float* feat, cudaStream_t st);

cudaError_t launch_deform_f16(const DeformDesc& d, const TcWeights& tw, int mode, const CameraDev* cam, bool use_cam_time,
                              int64_t n, const DeformIO& io, int sm_count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    if (!tw.feat) return cudaErrorInvalidValue;
    {
        cudaError_t e = launch_deform_features(d, n, io.xyz, tw.feat, st);
        if (e != cudaSuccess) return e;
    }
    const F16Smem Ls = f16_smem_layout(d.F, (d.head_mask & G4D_HEAD_SHS) != 0);
    const size_t bytes = Ls.total + 1024;
    const int64_t ntiles = (n + 127) / 128;
    const int64_t want = (ntiles + 1) / 2;                       // two tiles in flight per CTA
    const int grid = (int)(want < sm_count ? want : sm_count);
    const bool use_cam = mode == 1 || use_cam_time;
    cudaError_t rc = cudaErrorInvalidValue;
#define G4D_F16_CASE(CC, LL)                                                                                       \
    if (d.C == CC && d.levels == LL)                                                                               \
        rc = mode == 0 ? launch_deform_f16_t<0, CC, LL>(d, tw, Ls, bytes, grid, cam, use_cam, n, io, st)           \
                       : launch_deform_f16_t<1, CC, LL>(d, tw, Ls, bytes, grid, cam, use_cam, n, io, st);
    G4D_F16_CASE(16, 2)
    G4D_F16_CASE(16, 3)
    G4D_F16_CASE(32, 2)
#undef G4D_F16_CASE
    return rc;
}

}  // namespace g4d
