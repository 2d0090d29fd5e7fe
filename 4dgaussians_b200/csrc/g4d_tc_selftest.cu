// g4d_tc_selftest.cu -- single-CTA self test of the tcgen05 building blocks in tc_umma.cuh:
//   D[128 x N] = A[128 x K] * B[N x K]^T   with 3xTF32, A through TMEM, B through shared memory.
// The layout / descriptor hypotheses are runtime parameters so that one GPU session can tell which encoding the
// hardware accepts (there is no GPU in the build container).  Debug entry point g4d_debug_umma (include/g4d.h).
#include "g4d_internal.h"
#include "tc_umma.cuh"

namespace g4d {

struct UmmaTestCfg {
    int N, K;
    int layout_mode;   // 0: K cores contiguous (LBO=128, SBO=(K/4)*128); 1: 8-row groups contiguous (SBO=128, LBO=(N/8)*128)
    int swap_desc;     // 1: put LBO in the SBO field and vice versa
    int a_cols_per_k;  // TMEM columns the A operand advances per element of K (1 for 32-bit types)
    int use_tma;       // 1: B arrives pre-packed (hi | lo) in global memory and is staged with cp.async.bulk
    int single_pass;   // 1: plain TF32 (hi*hi only)
    int version_bit;   // descriptor version field (1 on sm_100)
};

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(UmmaTestCfg c, const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ Bpacked,
                     float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar_mma, bar_tma;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t N = c.N, K = c.K;
    uint8_t* b_hi = smem_raw;
    uint8_t* b_lo = smem_raw + N * K * 4;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, tc::kTmemCols);
    if (tid == 0) { mbar_init(&bar_mma, 1); mbar_init(&bar_tma, 1); fence_barrier_init(); }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t lbo = c.layout_mode == 0 ? 128u : (N >> 3) * 128u;
    const uint32_t sbo = c.layout_mode == 0 ? (K >> 2) * 128u : 128u;
    // ---- B -> shared memory (canonical layout), hi and lo parts
    if (c.use_tma) {
        if (tid == 0) {
            mbar_expect_tx(&bar_tma, 2 * N * K * 4);
            tma_bulk_g2s(b_hi, Bpacked, N * K * 4, &bar_tma);
            tma_bulk_g2s(b_lo, Bpacked + N * K, N * K * 4, &bar_tma);
        }
        mbar_wait(&bar_tma, 0);
    } else {
        for (uint32_t i = tid; i < N * K; i += blockDim.x) {
            const uint32_t n = i / K, k = i % K;
            uint32_t hi, lo;
            tc::tf32_split(B[i], hi, lo);
            const uint32_t off = (n >> 3) * sbo + (k >> 2) * lbo + (n & 7) * 16 + (k & 3) * 4;
            *reinterpret_cast<uint32_t*>(b_hi + off) = hi;
            *reinterpret_cast<uint32_t*>(b_lo + off) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core (async proxy) reads
    }
    // ---- A -> TMEM: thread r owns lane r; hi at columns [0, K*cpk), lo right after
    {
        const int r = tid;
        const uint32_t cpk = c.a_cols_per_k;
        for (uint32_t k0 = 0; k0 < K; k0 += 8) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) tc::tf32_split(A[r * K + k0 + j], hi[j], lo[j]);
            tc::tmem_st8(tbase + lane_base + k0 * cpk, hi);
            tc::tmem_st8(tbase + lane_base + K * cpk + k0 * cpk, lo);
        }
        tc::wait_st();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t d_col = 256;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(128, N);
        bool acc = false;
        const int passes = c.single_pass ? 1 : 3;
        for (int p = 0; p < passes; ++p) {
            const uint32_t a = tbase + ((c.single_pass || p != 0) ? 0u : K * c.a_cols_per_k);
            const uint8_t* b = (!c.single_pass && p == 1) ? b_lo : b_hi;
            for (uint32_t ks = 0; ks < K; ks += 8) {
                const uint32_t saddr = tc::smem_addr(b) + (ks >> 2) * lbo;
                uint64_t bd = c.swap_desc ? tc::make_smem_desc(saddr, sbo, lbo) : tc::make_smem_desc(saddr, lbo, sbo);
                if (!c.version_bit) bd &= ~(1ull << 46);
                tc::umma_tf32_ts(tbase + d_col, a + ks * c.a_cols_per_k, bd, idesc, acc);
                acc = true;
            }
        }
        tc::umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc::fence_after_sync();
    {
        const int r = tid;
        for (uint32_t n0 = 0; n0 < N; n0 += 8) {
            uint32_t v[8];
            tc::tmem_ld8(tbase + lane_base + d_col + n0, v);
            tc::wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) D[r * N + n0 + j] = __uint_as_float(v[j]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, tc::kTmemCols);
}

// pack an [N][K] fp32 matrix into (hi | lo) canonical tf32 images (layout_mode 0)
__global__ void pack_canonical_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * K) return;
    const uint32_t n = i / K, k = i % K;
    uint32_t hi, lo;
    tc::tf32_split(src[i], hi, lo);
    const uint32_t off = tc::canon_off(n, k, K) >> 2;
    reinterpret_cast<uint32_t*>(dst)[off] = hi;
    reinterpret_cast<uint32_t*>(dst)[N * K + off] = lo;
}

cudaError_t launch_umma_selftest(const int cfg[8], const float* A, const float* B, float* scratch_packed, float* D, cudaStream_t st) {
    UmmaTestCfg c{cfg[0], cfg[1], cfg[2], cfg[3], cfg[4], cfg[5], cfg[6], cfg[7]};
    if (c.N < 16 || c.N > 128 || (c.N % 16) || c.K < 8 || c.K > 128 || (c.K % 8) || c.a_cols_per_k < 1 || c.a_cols_per_k > 2)
        return cudaErrorInvalidValue;
    if (c.use_tma) {
        pack_canonical_kernel<<<(c.N * c.K + 255) / 256, 256, 0, st>>>(B, scratch_packed, c.N, c.K);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    const size_t smem = (size_t)2 * c.N * c.K * 4 + 1024;
    cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    umma_selftest_kernel<<<1, 128, smem, st>>>(c, A, B, scratch_packed, D);
    return cudaGetLastError();
}

}  // namespace g4d

// ------------------------------------------------------------------------------------------------------
// Self test #2: kind::f16 (BF16 hi+lo, 3 products) with every operand role the backward kernels use.
//   D[128 x N] = A[128 x K] * B[N x K]^T
// a_mode: 0 = TMEM (two bf16 per 32-bit column), 1 = smem K-major image of A[M][K], 2 = smem MN-major image built from
//         the TRANSPOSED matrix At[K][M] (the layout a [g][j] activation image has when g is the contraction index)
// b_mode: 0 = smem K-major image of B[N][K], 1 = smem MN-major image built from Bt[K][N]
// pack_hi_first: TMEM packing order of the two K elements in a column (0: even k in the low half)
// ------------------------------------------------------------------------------------------------------
namespace g4d {

struct Umma16Cfg { int N, K, a_mode, b_mode, pack_hi_first, single_pass, r0, r1; };

__global__ void __launch_bounds__(128, 1)
umma16_selftest_kernel(Umma16Cfg c, const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar_mma;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t N = c.N, K = c.K, M = 128;
    // images: A hi, A lo (128 x K or K x 128), B hi, B lo
    uint8_t* a_img[2] = {smem_raw, smem_raw + M * K * 2};
    uint8_t* b_img[2] = {smem_raw + 2 * M * K * 2, smem_raw + 2 * M * K * 2 + N * K * 2};
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, tc::kTmemCols);
    if (tid == 0) { mbar_init(&bar_mma, 1); fence_barrier_init(); }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    // ---- B images
    for (uint32_t i = tid; i < N * K; i += blockDim.x) {
        const uint32_t n = i / K, k = i % K;
        uint16_t hi, lo;
        tc::bf16_split(B[i], hi, lo);
        const uint32_t off = c.b_mode == 0 ? tc::img16_off(n, k, K) : tc::img16_off(k, n, N);
        *reinterpret_cast<uint16_t*>(b_img[0] + off) = hi;
        *reinterpret_cast<uint16_t*>(b_img[1] + off) = lo;
    }
    // ---- A: images or TMEM
    if (c.a_mode != 0) {
        for (uint32_t i = tid; i < M * K; i += blockDim.x) {
            const uint32_t m = i / K, k = i % K;
            uint16_t hi, lo;
            tc::bf16_split(A[i], hi, lo);
            const uint32_t off = c.a_mode == 1 ? tc::img16_off(m, k, K) : tc::img16_off(k, m, M);
            *reinterpret_cast<uint16_t*>(a_img[0] + off) = hi;
            *reinterpret_cast<uint16_t*>(a_img[1] + off) = lo;
        }
    } else {
        const int r = tid;
        for (uint32_t k0 = 0; k0 < K; k0 += 16) {   // 16 K elements = 8 columns
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint16_t h0, l0, h1, l1;
                tc::bf16_split(A[r * K + k0 + 2 * j], h0, l0);
                tc::bf16_split(A[r * K + k0 + 2 * j + 1], h1, l1);
                hi[j] = c.pack_hi_first ? ((uint32_t)h0 << 16 | h1) : ((uint32_t)h1 << 16 | h0);
                lo[j] = c.pack_hi_first ? ((uint32_t)l0 << 16 | l1) : ((uint32_t)l1 << 16 | l0);
            }
            tc::tmem_st8(tbase + lane_base + (k0 >> 1), hi);
            tc::tmem_st8(tbase + lane_base + 128 + (k0 >> 1), lo);
        }
        tc::wait_st();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t d_col = 256;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(128, N, c.a_mode == 2, c.b_mode == 1);
        // per-operand descriptor strides and the byte step of one K = 16 MMA
        const uint32_t a_lbo = c.a_mode == 1 ? 128u : (M >> 3) * 128u, a_sbo = c.a_mode == 1 ? (K >> 3) * 128u : 128u;
        const uint32_t a_step = c.a_mode == 1 ? 2u * 128u : 2u * (M >> 3) * 128u;
        const uint32_t b_lbo = c.b_mode == 0 ? 128u : (N >> 3) * 128u, b_sbo = c.b_mode == 0 ? (K >> 3) * 128u : 128u;
        const uint32_t b_step = c.b_mode == 0 ? 2u * 128u : 2u * (N >> 3) * 128u;
        bool acc = false;
        const int passes = c.single_pass ? 1 : 3;
        for (int p = 0; p < passes; ++p) {
            const int ai = (!c.single_pass && p == 0) ? 1 : 0, bi = (!c.single_pass && p == 1) ? 1 : 0;
            for (uint32_t ks = 0; ks < K; ks += 16) {
                const uint64_t bd = tc::make_smem_desc(tc::smem_addr(b_img[bi]) + (ks >> 4) * b_step, b_lbo, b_sbo);
                if (c.a_mode == 0) {
                    tc::umma_bf16_ts(tbase + d_col, tbase + (ai ? 128u : 0u) + (ks >> 1), bd, idesc, acc);
                } else {
                    const uint64_t ad = tc::make_smem_desc(tc::smem_addr(a_img[ai]) + (ks >> 4) * a_step, a_lbo, a_sbo);
                    tc::umma_bf16_ss(tbase + d_col, ad, bd, idesc, acc);
                }
                acc = true;
            }
        }
        tc::umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc::fence_after_sync();
    {
        const int r = tid;
        for (uint32_t n0 = 0; n0 < N; n0 += 8) {
            uint32_t v[8];
            tc::tmem_ld8(tbase + lane_base + d_col + n0, v);
            tc::wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) D[r * N + n0 + j] = __uint_as_float(v[j]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, tc::kTmemCols);
}

cudaError_t launch_umma16_selftest(const int cfg[8], const float* A, const float* B, float* D, cudaStream_t st) {
    Umma16Cfg c{cfg[0], cfg[1], cfg[2], cfg[3], cfg[4], cfg[5], cfg[6], cfg[7]};
    if (c.N < 16 || c.N > 128 || (c.N % 16) || c.K < 16 || c.K > 128 || (c.K % 16)) return cudaErrorInvalidValue;
    const size_t smem = (size_t)2 * 128 * c.K * 2 + (size_t)2 * c.N * c.K * 2 + 1024;
    cudaError_t e = cudaFuncSetAttribute(umma16_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    umma16_selftest_kernel<<<1, 128, smem, st>>>(c, A, B, D);
    return cudaGetLastError();
}

}  // namespace g4d

// ---- C entry of the stand-alone self-test library (libg4d_selftest.so; NOT part of libg4d.so / include/g4d.h) -------------
// cfg = {N, K, layout_mode, swap_desc, a_cols_per_k, use_tma, single_pass | 16 for the bf16 variant, version_bit}
extern "C" int g4d_selftest_umma(const int* cfg, const float* A, const float* B, float* D, void* stream) {
    if (!cfg || !A || !B || !D) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    if (cfg[6] == 16) return g4d::launch_umma16_selftest(cfg, A, B, D, st) == cudaSuccess ? 0 : -1;
    static float* scratch = nullptr;
    if (!scratch && cudaMalloc(&scratch, (size_t)2 * 128 * 128 * 4 + 256) != cudaSuccess) return -3;
    return g4d::launch_umma_selftest(cfg, A, B, scratch, D, st) == cudaSuccess ? 0 : -1;
}
