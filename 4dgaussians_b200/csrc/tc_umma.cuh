// tc_umma.cuh -- hand-written Blackwell tensor-core building blocks (inline PTX, sm_100a only):
// TMEM allocation, tcgen05.ld / tcgen05.st, tcgen05.mma kind::tf32 with the A operand in TMEM and the B operand in
// shared memory (K-major, SWIZZLE_NONE canonical layout), tcgen05.commit -> mbarrier, and the 3xTF32 split that keeps
// fp32-level accuracy (the reference MLP is plain fp32: scene/deformation.py:80-142 through cuBLAS SGEMM).
//
// Canonical B layout used everywhere here (units of bytes), for an [N][K] tf32 matrix (K contiguous per row in the
// source): core matrix = 8 rows x 16 B (= 4 tf32) stored as 128 contiguous bytes;
//     off(n, k) = (n / 8) * SBO + (k / 4) * LBO + (n % 8) * 16 + (k % 4) * 4,     LBO = 128, SBO = (K / 4) * 128
// (CUTLASS "Major-K INTERLEAVE": ((8,n),2):((1,SBO),LBO) in 16-byte units).  One MMA consumes K = 8 (two core
// matrices): advancing one k-step adds 2*LBO bytes to the descriptor's start address.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace g4d {
namespace tc {

constexpr uint32_t kTmemCols = 512;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM management (one full warp executes alloc / dealloc) ----------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(smem_result)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- TMEM <-> registers: the warp's 32 lanes of its quadrant, N consecutive 32-bit columns -------------------
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

// ---- issuing tcgen05.mma ---------------------------------------------------------------------------------------------------
// UTCHMMA takes every operand from UNIFORM registers.  Issued from a single-lane divergent branch (`if (tid == X)`) the compiler
// wraps each instruction in a waterfall loop (ELECT / R2UR.BROADCAST / BRA.U.ANY, ~150 cycles per MMA, measured with
// tools/umma_rate.py); issued by an elected lane of a CONVERGED warp the operands stay in uniform registers and the issue rate
// is the tensor pipe's (64 cycles per 128x128x8 tf32, 76 per 128x128x16 f16 dispatch).  elect.sync with the full member mask
// also tells the compiler that the warp is converged here.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- 3xTF32 split ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_rna(x);
    lo = tf32_rna(x - __uint_as_float(hi));
}

// ---- descriptors ----------------------------------------------------------------------------------------------------
// instruction descriptor, kind::tf32, D = F32, A and B K-major (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N) {
    return (1u << 4)            // c_format = F32
           | (2u << 7)          // a_format = TF32
           | (2u << 10)         // b_format = TF32
           | ((N >> 3) << 17)   // n_dim
           | ((M >> 4) << 24);  // m_dim
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, LBO, SBO in 16-byte units, version = 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// D[tmem] (+)= A[tmem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(void* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// byte offset of element (n, k) of an [N][K] tf32 matrix in the canonical layout described at the top
__host__ __device__ constexpr uint32_t canon_off(uint32_t n, uint32_t k, uint32_t K) {
    return (n >> 3) * ((K >> 2) * 128u) + (k >> 2) * 128u + (n & 7u) * 16u + (k & 3u) * 4u;
}

// One GEMM  D[128 x N] (+)= A[128 x K] * B[N x K]^T  as 3 TF32 products (lo*hi, hi*lo, hi*hi), A hi/lo in TMEM at
// a_hi / a_lo (K columns each), B hi/lo in shared memory (canonical layout, kcore0 = first 16-byte K core to use,
// Kfull = K extent the layout was packed with).  Issued by ONE thread; fully unrolled so that the descriptor
// arithmetic is a chain of constant adds and the UTCHMMA issue rate is not limited by address computation.
template <int K>
__device__ __forceinline__ void gemm_3xtf32(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi_saddr, uint32_t b_lo_saddr,
                                            uint32_t N, uint32_t Kfull, uint32_t kcore0, bool accumulate) {
    const uint32_t idesc = make_idesc_tf32(128, N);
    const uint32_t sbo = (Kfull >> 2) * 128u;
    const uint64_t bd_hi = make_smem_desc(b_hi_saddr + kcore0 * 128u, 128u, sbo);
    const uint64_t bd_lo = make_smem_desc(b_lo_saddr + kcore0 * 128u, 128u, sbo);
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const uint32_t a = (p == 0) ? a_lo : a_hi;
        const uint64_t bd = (p == 1) ? bd_lo : bd_hi;
#pragma unroll
        for (int ks = 0; ks < K; ks += 8) {
            // one k-step = two 128-byte K cores = +16 in the descriptor's 16-byte address field
            umma_tf32_ts(d_tmem, a + ks, bd + (uint64_t)((ks >> 2) * 8), idesc, accumulate || p > 0 || ks > 0);
        }
    }
}

}  // namespace tc
}  // namespace g4d

// ======================================================================================================
// kind::f16 with BF16 operands, fp32 accumulate; "bf16x2" split (hi + lo, 3 products) gives ~16 mantissa bits
// (measured gradient error 1.5e-5 of max).  Used by the backward kernels.
//
// One image format serves every operand role: 8 x 8 bf16 core matrices (8 rows x 16 bytes = 128 contiguous bytes)
// stored as image[row / 8][col / 8][row % 8][col % 8]:
//     off(row, col) = (row / 8) * (ncols / 8) * 128 + (col / 8) * 128 + (row % 8) * 16 + (col % 8) * 2
//   * read as a K-major operand  (rows = M or N, cols = K):  LBO = 128 (K-adjacent cores), SBO = (ncols/8)*128
//   * read as an MN-major operand (rows = K, cols = M or N): SBO = 128 (MN-adjacent cores), LBO = (ncols/8)*128,
//     with the a_major / b_major bit of the instruction descriptor set.
// ======================================================================================================
namespace g4d {
namespace tc {

__host__ __device__ constexpr uint32_t img16_off(uint32_t row, uint32_t col, uint32_t ncols) {
    return (row >> 3) * ((ncols >> 3) * 128u) + (col >> 3) * 128u + (row & 7u) * 16u + (col & 7u) * 2u;
}

__device__ __forceinline__ uint16_t bf16_rn(float x) {
    uint16_t r;
    asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void bf16_split(float x, uint16_t& hi, uint16_t& lo) {
    hi = bf16_rn(x);
    lo = bf16_rn(x - __uint_as_float((uint32_t)hi << 16));
}

// instruction descriptor, kind::f16, A/B = BF16, D = F32; a_mn / b_mn select MN-major operands
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((N >> 3) << 17) |
           ((M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}

}  // namespace tc
}  // namespace g4d
