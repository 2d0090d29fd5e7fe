// deform_bwd_common.cuh -- pieces shared by the FFMA and the tensor-core backward of the deformation network.
#pragma once
#include "g4d_internal.h"

namespace g4d {

struct DeformBwdBuffers {
    float* feat;                      // [N][F]
    float* a1;                        // [N][WD]
    float* da1[G4D_NUM_HEADS];        // [N][WD] per active head
    float* trow_grad[G4D_MAX_LEVELS][3];
};

struct DeformBwdDesc {
    DeformDesc d;
    const float* w0;                  // torch layout [WD][F]
    const float* w1[G4D_NUM_HEADS];   // torch layout [WD][WD]
    float* g_w0; float* g_b0;
    float* g_w1[G4D_NUM_HEADS]; float* g_b1[G4D_NUM_HEADS];
    float* g_w2[G4D_NUM_HEADS]; float* g_b2[G4D_NUM_HEADS];
    float* g_planes[G4D_MAX_LEVELS][6];
    const float* go[G4D_NUM_HEADS];   // dL/d(out) per head: xyz[N,3], scaling[N,3], rotation[N,4], opacity[N,1], shs[N,48] (NULL = 0)
    float* gi[G4D_NUM_HEADS];         // dL/d(in), same shapes (NULL = not wanted)
};

G4D_D void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// border-clamped tap with the coordinate-gradient multiplier ATen uses (0 where the coordinate was clipped)
struct TapG { int i0, i1; float w0, w1, gmul; };
G4D_D TapG make_tap_g(float u, int size) {
    float x = ((u + 1.f) / 2.f) * (float)(size - 1);
    TapG t;
    const float hi = (float)(size - 1);
    t.gmul = (x > 0.f && x < hi) ? 0.5f * hi : 0.f;
    x = fminf(fmaxf(x, 0.f), hi);
    const float x0 = floorf(x);
    t.i0 = (int)x0; t.i1 = min(t.i0 + 1, size - 1);
    t.w0 = (x0 + 1.f) - x; t.w1 = x - x0;
    return t;
}


// Backward of the HexPlane sampling for ONE channel vector (4 channels) of one level of one Gaussian:
// scatter d(feat) into the 4 / 2 taps of every plane (vector RED) and accumulate dL/d(normalised coordinate).
// pcs = normalised (x,y,z); df = dL/d(feat[l*C + 4v .. +3]).  (scene/hexplane.py:73-106 through autograd)
G4D_D void scatter_vector(const DeformDesc& d, float* const (*g_planes)[6], float* const (*trow_grad)[3], int l, int v, int C4,
                          const float pcs[3], float4 df, float gpix[3]) {
    TapG tx[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) tx[a] = make_tap_g(pcs[a], d.res[l][a]);
    float4 s[6], dsx[6], dsy[6];   // sample, d(sample)/d(x_pix of c0), d/d(y_pix of c1)
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int c0 = plane_axis0(k), c1 = plane_axis1(k);
        if (c1 == 3) {
            const float4* row = reinterpret_cast<const float4*>(d.trow[l][c0]);
            const float4 r0 = __ldg(row + tx[c0].i0 * C4 + v), r1 = __ldg(row + tx[c0].i1 * C4 + v);
            const float w0 = tx[c0].w0, w1 = tx[c0].w1;
            s[k] = make_float4(fmaf(r1.x, w1, r0.x * w0), fmaf(r1.y, w1, r0.y * w0), fmaf(r1.z, w1, r0.z * w0), fmaf(r1.w, w1, r0.w * w0));
            dsx[k] = make_float4(r1.x - r0.x, r1.y - r0.y, r1.z - r0.z, r1.w - r0.w);
            dsy[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            const int W = d.res[l][c0];
            const float4* pl = reinterpret_cast<const float4*>(d.planes[l][k]);
            const TapG &X = tx[c0], &Y = tx[c1];
            const float4 nw = __ldg(pl + (Y.i0 * W + X.i0) * C4 + v), ne = __ldg(pl + (Y.i0 * W + X.i1) * C4 + v);
            const float4 sw = __ldg(pl + (Y.i1 * W + X.i0) * C4 + v), se = __ldg(pl + (Y.i1 * W + X.i1) * C4 + v);
            const float wnw = X.w0 * Y.w0, wne = X.w1 * Y.w0, wsw = X.w0 * Y.w1, wse = X.w1 * Y.w1;
            s[k] = make_float4(fmaf(se.x, wse, fmaf(sw.x, wsw, fmaf(ne.x, wne, nw.x * wnw))),
                               fmaf(se.y, wse, fmaf(sw.y, wsw, fmaf(ne.y, wne, nw.y * wnw))),
                               fmaf(se.z, wse, fmaf(sw.z, wsw, fmaf(ne.z, wne, nw.z * wnw))),
                               fmaf(se.w, wse, fmaf(sw.w, wsw, fmaf(ne.w, wne, nw.w * wnw))));
            dsx[k] = make_float4((ne.x - nw.x) * Y.w0 + (se.x - sw.x) * Y.w1, (ne.y - nw.y) * Y.w0 + (se.y - sw.y) * Y.w1,
                                 (ne.z - nw.z) * Y.w0 + (se.z - sw.z) * Y.w1, (ne.w - nw.w) * Y.w0 + (se.w - sw.w) * Y.w1);
            dsy[k] = make_float4((sw.x - nw.x) * X.w0 + (se.x - ne.x) * X.w1, (sw.y - nw.y) * X.w0 + (se.y - ne.y) * X.w1,
                                 (sw.z - nw.z) * X.w0 + (se.z - ne.z) * X.w1, (sw.w - nw.w) * X.w0 + (se.w - ne.w) * X.w1);
        }
    }
    float4 pre[6], suf[6];   // prefix / suffix products so that a zero sample does not poison the others
    pre[0] = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
    for (int k = 1; k < 6; ++k) pre[k] = make_float4(pre[k - 1].x * s[k - 1].x, pre[k - 1].y * s[k - 1].y, pre[k - 1].z * s[k - 1].z, pre[k - 1].w * s[k - 1].w);
    suf[5] = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
    for (int k = 4; k >= 0; --k) suf[k] = make_float4(suf[k + 1].x * s[k + 1].x, suf[k + 1].y * s[k + 1].y, suf[k + 1].z * s[k + 1].z, suf[k + 1].w * s[k + 1].w);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int c0 = plane_axis0(k), c1 = plane_axis1(k);
        const float4 gs = make_float4(df.x * pre[k].x * suf[k].x, df.y * pre[k].y * suf[k].y, df.z * pre[k].z * suf[k].z, df.w * pre[k].w * suf[k].w);
        gpix[c0] += (gs.x * dsx[k].x + gs.y * dsx[k].y + gs.z * dsx[k].z + gs.w * dsx[k].w) * tx[c0].gmul;
        if (c1 == 3) {
            float* row = trow_grad[l][c0];
            const float w0 = tx[c0].w0, w1 = tx[c0].w1;
            red_add_v4(row + (tx[c0].i0 * C4 + v) * 4, make_float4(gs.x * w0, gs.y * w0, gs.z * w0, gs.w * w0));
            red_add_v4(row + (tx[c0].i1 * C4 + v) * 4, make_float4(gs.x * w1, gs.y * w1, gs.z * w1, gs.w * w1));
        } else {
            gpix[c1] += (gs.x * dsy[k].x + gs.y * dsy[k].y + gs.z * dsy[k].z + gs.w * dsy[k].w) * tx[c1].gmul;
            const int W = d.res[l][c0];
            float* pl = g_planes[l][k];
            const TapG &X = tx[c0], &Y = tx[c1];
            const float wnw = X.w0 * Y.w0, wne = X.w1 * Y.w0, wsw = X.w0 * Y.w1, wse = X.w1 * Y.w1;
            red_add_v4(pl + ((Y.i0 * W + X.i0) * C4 + v) * 4, make_float4(gs.x * wnw, gs.y * wnw, gs.z * wnw, gs.w * wnw));
            red_add_v4(pl + ((Y.i0 * W + X.i1) * C4 + v) * 4, make_float4(gs.x * wne, gs.y * wne, gs.z * wne, gs.w * wne));
            red_add_v4(pl + ((Y.i1 * W + X.i0) * C4 + v) * 4, make_float4(gs.x * wsw, gs.y * wsw, gs.z * wsw, gs.w * wsw));
            red_add_v4(pl + ((Y.i1 * W + X.i1) * C4 + v) * 4, make_float4(gs.x * wse, gs.y * wse, gs.z * wse, gs.w * wse));
        }
    }
}

}  // namespace g4d
