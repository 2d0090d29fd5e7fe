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
//   * epilogues (bias, ReLU, hi/lo split) are tcgen05.ld -> registers -> tcgen05.st, one thread per TMEM lane,
//     two warps per lane quadrant splitting the columns;
//   * the tile's 128 result rows stay in registers of their owner threads, which finish the Gaussian
//     (activations, EWA projection, SH colour) exactly like the SIMT path (geom_finish.cuh).
//
// TMEM map (512 columns): [0,128) A1 hi, [128,256) A1 lo, [256,384) D (layer-1 accumulators; D2 reuses [256,256+48)),
// [384,512) X: feature operand (hi [384,384+F), lo [448,448+F)) then the head's hidden half (hi [384,448), lo [448,512)).
// Compiled with -fmad=false (it contains the projection, see g4d_math.cuh).
#include "geom_finish.cuh"
#include "tc_umma.cuh"

namespace g4d {

constexpr uint32_t kColA1Hi = 0, kColA1Lo = 128, kColD = 256, kColX = 384, kColXLo = 448;

struct TcSmem { uint32_t w1, w2, w0, bias, coord, bars, total; };

inline TcSmem tc_smem_layout(int F, int max_kp16) {
    TcSmem s{};
    uint32_t off = 0;
    auto take = [&](uint32_t bytes) { uint32_t o = off; off += (bytes + 127u) & ~127u; return o; };
    s.w1 = take(2u * 128 * 128 * 4);
    s.w2 = take(2u * max_kp16 * 128 * 4);
    s.w0 = take(2u * 128 * F * 4);
    s.bias = take((128 + G4D_NUM_HEADS * 128 + 64) * 4);
    s.coord = take(128 * 16);
    s.bars = take(64);
    s.total = off;
    return s;
}

// ---- weight packing (once per parameter version) --------------------------------------------------------------
struct TcPackDesc {
    const float* src[1 + 2 * G4D_NUM_HEADS];
    float* dst[1 + 2 * G4D_NUM_HEADS];
    int rows_src[1 + 2 * G4D_NUM_HEADS], rows_dst[1 + 2 * G4D_NUM_HEADS], K[1 + 2 * G4D_NUM_HEADS];
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
    auto add = [&](const float* src, int rows_src, int rows_dst, int K) {
        p.src[m] = src; p.dst[m] = q; p.rows_src[m] = rows_src; p.rows_dst[m] = rows_dst; p.K[m] = K; p.start[m] = total;
        total += rows_dst * K; float* r = q; q += 2 * (size_t)rows_dst * K; ++m; return r;
    };
    out->w0 = add(prm.w0, 128, 128, F);
    for (int h = 0; h < G4D_NUM_HEADS; ++h) {
        out->kp16[h] = h == 4 ? 48 : 16;
        out->w1[h] = nullptr; out->w2[h] = nullptr;
        if (!(prm.head_mask & (1 << h))) continue;
        out->w1[h] = add(prm.w1[h], 128, 128, 128);
        out->w2[h] = add(prm.w2[h], head_out(h), out->kp16[h], 128);
    }
    p.start[m] = total; p.count = m;
    tc_pack_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(p);
    return cudaGetLastError();
}

// ---- the kernel -------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256, 1)
deform_tc_kernel(DeformDesc d, TcWeights tw, TcSmem L, const CameraDev* __restrict__ camp, float time_arg, int use_cam_time,
                 int64_t n, DeformIO io) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ CameraDev cam;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & 127, half = tid >> 7;
    const int64_t ntiles = (n + 127) / 128;
    if (MODE == 1 || use_cam_time) {
        for (int i = tid; i < (int)(sizeof(CameraDev) / 4); i += 256)
            reinterpret_cast<uint32_t*>(&cam)[i] = reinterpret_cast<const uint32_t*>(camp)[i];
    }
    float* sBias = reinterpret_cast<float*>(smem + L.bias);     // b0[128] | b1[5][128] | b2[64]
    float* coord = reinterpret_cast<float*>(smem + L.coord);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t *bar_w0 = bars, *bar_w1 = bars + 1, *bar_w2 = bars + 2, *bar_mma = bars + 3;
    int b2off[G4D_NUM_HEADS];
    {
        int o = 0;
        for (int h = 0; h < G4D_NUM_HEADS; ++h) { b2off[h] = o; if (d.head_mask & (1 << h)) o += head_out(h); }
    }
    for (int i = tid; i < 128; i += 256) sBias[i] = __ldg(d.b0 + i);
    for (int h = 0; h < G4D_NUM_HEADS; ++h) {
        if (!(d.head_mask & (1 << h))) continue;
        for (int i = tid; i < 128; i += 256) sBias[128 + h * 128 + i] = __ldg(d.b1[h] + i);
        for (int i = tid; i < head_out(h); i += 256) sBias[128 + G4D_NUM_HEADS * 128 + b2off[h] + i] = __ldg(d.b2[h] + i);
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, tc::kTmemCols);
    if (tid == 0) {
        mbar_init(bar_w0, 1); mbar_init(bar_w1, 1); mbar_init(bar_w2, 1); mbar_init(bar_mma, 1);
        fence_barrier_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t sW1 = tc::smem_addr(smem + L.w1), sW2 = tc::smem_addr(smem + L.w2), sW0 = tc::smem_addr(smem + L.w0);
    const int F = d.F;
    const bool have_tiles = (int64_t)blockIdx.x < ntiles;
    if (tid == 0 && have_tiles) {
        mbar_expect_tx(bar_w0, 2u * 128 * F * 4);
        tma_bulk_g2s(smem + L.w0, tw.w0, 2u * 128 * F * 4, bar_w0);
        if (d.head_mask) {
            const int h0 = __ffs(d.head_mask) - 1;
            mbar_expect_tx(bar_w1, 2u * 65536);
            tma_bulk_g2s(smem + L.w1, tw.w1[h0], 65536, bar_w1);
            tma_bulk_g2s(smem + L.w1 + 65536, tw.w1[h0] + 16384, 65536, bar_w1);
            const uint32_t w2b = (uint32_t)tw.kp16[h0] * 128 * 4;
            mbar_expect_tx(bar_w2, 2u * w2b);
            tma_bulk_g2s(smem + L.w2, tw.w2[h0], w2b, bar_w2);
            tma_bulk_g2s(smem + L.w2 + w2b, tw.w2[h0] + tw.kp16[h0] * 128, w2b, bar_w2);
        }
    }
    if (have_tiles) mbar_wait(bar_w0, 0);
    uint32_t ph_w1 = 0, ph_w2 = 0, ph_mma = 0;
    const float t = use_cam_time ? cam.time : time_arg;
    float amax[3], ascale[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { amax[a] = __ldg(d.aabb + a); ascale[a] = 2.0f / (__ldg(d.aabb + 3 + a) - amax[a]); }
    const int C4 = d.C >> 2, C4h = C4 >> 1;   // channel vectors per level, per half

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * 128, gi = base + row;
        const bool valid = gi < n;
        // ---- inputs of my Gaussian (owner threads: half == 0)
        Vec3 p{0.f, 0.f, 0.f};
        float sl[3] = {0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f}, ol = 0.f;
        if (half == 0) {
            if (valid) {
                p = Vec3{io.xyz[3 * gi], io.xyz[3 * gi + 1], io.xyz[3 * gi + 2]};
                if (io.scaling) { sl[0] = io.scaling[3 * gi]; sl[1] = io.scaling[3 * gi + 1]; sl[2] = io.scaling[3 * gi + 2]; }
                if (io.rotation) { const float4 r4 = *reinterpret_cast<const float4*>(io.rotation + 4 * gi); q[0] = r4.x; q[1] = r4.y; q[2] = r4.z; q[3] = r4.w; }
                if (io.opacity) ol = io.opacity[gi];
            }
            float4 c;
            c.x = (p.x - amax[0]) * ascale[0] - 1.0f; c.y = (p.y - amax[1]) * ascale[1] - 1.0f; c.z = (p.z - amax[2]) * ascale[2] - 1.0f;
            c.w = t;
            if (!valid) c = make_float4(0.f, 0.f, 0.f, t);
            *reinterpret_cast<float4*>(coord + 4 * row) = c;
        }
        __syncthreads();
        // ---- HexPlane features -> TMEM (X region), thread (row, half) owns channel vectors [half*C4h, (half+1)*C4h)
        {
            const float4 pc = *reinterpret_cast<const float4*>(coord + 4 * row);
            const float pcs[3] = {pc.x, pc.y, pc.z};
            for (int l = 0; l < d.levels; ++l) {
                Tap1D tx[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) tx[a] = make_tap(pcs[a], d.res[l][a]);
                for (int v0 = half * C4h; v0 < (half + 1) * C4h; v0 += 2) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int vv = 0; vv < 2; ++vv) {
                        const int v = v0 + vv;
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
                        tc::tf32_split(prod.x, hi[4 * vv + 0], lo[4 * vv + 0]); tc::tf32_split(prod.y, hi[4 * vv + 1], lo[4 * vv + 1]);
                        tc::tf32_split(prod.z, hi[4 * vv + 2], lo[4 * vv + 2]); tc::tf32_split(prod.w, hi[4 * vv + 3], lo[4 * vv + 3]);
                    }
                    const uint32_t col = (uint32_t)(l * d.C + 4 * v0);
                    tc::tmem_st8(lane_base + kColX + col, hi);
                    tc::tmem_st8(lane_base + kColXLo + col, lo);
                }
            }
            tc::wait_st();
        }
        tc::fence_before_sync();
        __syncthreads();
        // ---- layer 0: D = feat * W0^T
        if (tid == 0) {
            tc::fence_after_sync();
            tc::gemm_3xtf32(tbase + kColD, tbase + kColX, tbase + kColXLo, sW0, sW0 + 128u * F * 4, 128, F, F, 0, false);
            tc::umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, ph_mma); ph_mma ^= 1u;
        tc::fence_after_sync();
        // ---- epilogue 0: a1 = relu(D + b0) -> A1 (hi | lo); thread (row, half) handles columns [64*half, 64*half+64)
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            const uint32_t c0 = (uint32_t)(half * 64 + ch * 16);
            uint32_t v[16], hi[16], lo[16];
            tc::tmem_ld16(lane_base + kColD + c0, v);
            tc::wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) tc::tf32_split(fmaxf(__uint_as_float(v[j]) + sBias[c0 + j], 0.f), hi[j], lo[j]);
            tc::tmem_st16(lane_base + kColA1Hi + c0, hi);
            tc::tmem_st16(lane_base + kColA1Lo + c0, lo);
        }
        tc::wait_st();
        tc::fence_before_sync();
        __syncthreads();

        float dl[11];   // deltas of pos(3) scale(3) rot(4) opacity(1)
#pragma unroll
        for (int j = 0; j < 11; ++j) dl[j] = 0.f;
        float dsh[48];
#pragma unroll
        for (int j = 0; j < 48; ++j) dsh[j] = 0.f;

#pragma unroll
        for (int h = 0; h < G4D_NUM_HEADS; ++h) {
            if (!(d.head_mask & (1 << h))) continue;
            constexpr int kDummy = 0; (void)kDummy;
            const int kp16 = (h == 4) ? 48 : 16;
            const float* b1 = sBias + 128 + h * 128;
            // next head whose weights go into the buffers once this head has released them (-1: none)
            int nh = -1;
            {
                const int later = d.head_mask >> (h + 1);
                if (later) nh = h + 1 + (__ffs(later) - 1);
                else if (tile + gridDim.x < ntiles) nh = __ffs(d.head_mask) - 1;
            }
            // ---- layer 1: D = a1 * W1^T
            mbar_wait(bar_w1, ph_w1); ph_w1 ^= 1u;
            if (tid == 0) {
                tc::fence_after_sync();
                tc::gemm_3xtf32(tbase + kColD, tbase + kColA1Hi, tbase + kColA1Lo, sW1, sW1 + 65536u, 128, 128, 128, 0, false);
                tc::umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, ph_mma); ph_mma ^= 1u;
            tc::fence_after_sync();
            if (tid == 0) {   // W1 buffer is free: stream the next head's (or the next tile's first head's) W1
                if (nh >= 0) {
                    mbar_expect_tx(bar_w1, 2u * 65536);
                    tma_bulk_g2s(smem + L.w1, tw.w1[nh], 65536, bar_w1);
                    tma_bulk_g2s(smem + L.w1 + 65536, tw.w1[nh] + 16384, 65536, bar_w1);
                }
            }
            mbar_wait(bar_w2, ph_w2); ph_w2 ^= 1u;
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
                // hidden half hh: a2 = relu(D[:, 64hh : 64hh+64] + b1) -> X (hi | lo); thread handles 32 of the 64 columns
#pragma unroll 1
                for (int ch = 0; ch < 2; ++ch) {
                    const uint32_t cl = (uint32_t)(half * 32 + ch * 16);      // column inside the half
                    const uint32_t cg = (uint32_t)(hh * 64) + cl;             // column of the hidden layer
                    uint32_t v[16], hi[16], lo[16];
                    tc::tmem_ld16(lane_base + kColD + cg, v);
                    tc::wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) tc::tf32_split(fmaxf(__uint_as_float(v[j]) + b1[cg + j], 0.f), hi[j], lo[j]);
                    tc::tmem_st16(lane_base + kColX + cl, hi);
                    tc::tmem_st16(lane_base + kColXLo + cl, lo);
                }
                tc::wait_st();
                tc::fence_before_sync();
                __syncthreads();
                // ---- layer 2 partial: D2 (+)= a2_half * W2[:, 64hh : 64hh+64]^T     (D2 = D columns [0, kp16))
                if (tid == 0) {
                    tc::fence_after_sync();
                    tc::gemm_3xtf32(tbase + kColD, tbase + kColX, tbase + kColXLo, sW2, sW2 + (uint32_t)kp16 * 128 * 4, (uint32_t)kp16, 64,
                                    128, (uint32_t)(hh * 16), hh == 1);
                    tc::umma_commit(bar_mma);
                }
                mbar_wait(bar_mma, ph_mma); ph_mma ^= 1u;
                tc::fence_after_sync();
            }
            if (tid == 0) {   // W2 buffer is free
                if (nh >= 0) {
                    const uint32_t w2b = (uint32_t)tw.kp16[nh] * 128 * 4;
                    mbar_expect_tx(bar_w2, 2u * w2b);
                    tma_bulk_g2s(smem + L.w2, tw.w2[nh], w2b, bar_w2);
                    tma_bulk_g2s(smem + L.w2 + w2b, tw.w2[nh] + tw.kp16[nh] * 128, w2b, bar_w2);
                }
            }
            // ---- read the head's output row (owner threads)
            if (half == 0) {
                const float* b2 = sBias + 128 + G4D_NUM_HEADS * 128 + b2off[h];
                if (h < 4) {
                    uint32_t v[16];
                    tc::tmem_ld16(lane_base + kColD, v);
                    tc::wait_ld();
                    const int c = head_col(h), ko = head_out(h);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j < ko) dl[c + j] = __uint_as_float(v[j]) + b2[j];
                } else {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        uint32_t v[16];
                        tc::tmem_ld16(lane_base + kColD + ch * 16, v);
                        tc::wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; ++j) dsh[ch * 16 + j] = __uint_as_float(v[j]) + b2[ch * 16 + j];
                    }
                }
            }
            tc::fence_before_sync();
            __syncthreads();
        }
        // ---- finish my Gaussian
        if (half == 0 && valid) {
            p.x += dl[0]; p.y += dl[1]; p.z += dl[2];
            sl[0] += dl[3]; sl[1] += dl[4]; sl[2] += dl[5];
            q[0] += dl[6]; q[1] += dl[7]; q[2] += dl[8]; q[3] += dl[9];
            ol += dl[10];
            const bool hsh = d.head_mask & G4D_HEAD_SHS;
            if (MODE == 0) {
                io.out_xyz[3 * gi] = p.x; io.out_xyz[3 * gi + 1] = p.y; io.out_xyz[3 * gi + 2] = p.z;
                if (io.out_scaling) { io.out_scaling[3 * gi] = sl[0]; io.out_scaling[3 * gi + 1] = sl[1]; io.out_scaling[3 * gi + 2] = sl[2]; }
                if (io.out_rotation) *reinterpret_cast<float4*>(io.out_rotation + 4 * gi) = make_float4(q[0], q[1], q[2], q[3]);
                if (io.out_opacity) io.out_opacity[gi] = ol;
                if (io.out_shs && hsh) {
#pragma unroll
                    for (int j = 0; j < 48; ++j) io.out_shs[gi * 48 + j] = io.shs[gi * 48 + j] + dsh[j];
                }
            } else {
                if (hsh && io.fo.shs) {
#pragma unroll
                    for (int j = 0; j < 48; ++j) {
                        float b;
                        if (io.shs) b = io.shs[gi * 48 + j];
                        else b = j < 3 ? io.sh_dc[gi * 3 + j] : io.sh_rest[gi * 45 + (j - 3)];
                        io.fo.shs[gi * 48 + j] = b + dsh[j];
                    }
                }
                fused_finish(cam, io, gi, p, sl, q, ol, [&](int i) { return dsh[i]; });
            }
        }
        __syncthreads();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, tc::kTmemCols);
}

bool tc_deform_supported(const DeformDesc& d) {
    if (d.WD != 128) return false;
    if (d.C != 16 && d.C != 32) return false;
    if (d.F > 64 || (d.F % 8)) return false;
    const int max_kp = (d.head_mask & G4D_HEAD_SHS) ? 48 : 16;
    return tc_smem_layout(d.F, max_kp).total + 1024 <= 227 * 1024;
}

cudaError_t launch_deform_tc(const DeformDesc& d, const TcWeights& tw, int mode, const CameraDev* cam, float time,
                             bool use_cam_time, int64_t n, const DeformIO& io, int sm_count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const int max_kp = (d.head_mask & G4D_HEAD_SHS) ? 48 : 16;
    const TcSmem L = tc_smem_layout(d.F, max_kp);
    const size_t bytes = L.total + 1024;
    const int64_t ntiles = (n + 127) / 128;
    const int grid = (int)(ntiles < sm_count ? ntiles : sm_count);
    cudaError_t e;
    if (mode == 0) {
        e = cudaFuncSetAttribute(deform_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        deform_tc_kernel<0><<<grid, 256, bytes, st>>>(d, tw, L, cam, time, use_cam_time ? 1 : 0, n, io);
    } else {
        e = cudaFuncSetAttribute(deform_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        deform_tc_kernel<1><<<grid, 256, bytes, st>>>(d, tw, L, cam, time, use_cam_time ? 1 : 0, n, io);
    }
    return cudaGetLastError();
}

}  // namespace g4d
