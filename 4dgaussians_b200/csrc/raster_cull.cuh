// raster_cull.cuh -- exact-image culling predicates and the warp-cooperative rect walk shared by the binning kernels
// (g4d_bin.cu) and the blend kernels (g4d_raster.cu).
//
// A (Gaussian, tile) pair can be dropped without changing a single pixel when even the best-placed point of the
// tile's pixel rectangle has alpha = opacity * exp(-q/2) < 1/255 (the blend stage skips such contributions, A.3).
// q is a convex quadratic, so its minimum over the rectangle is 0 (centre inside) or lies on one of the 4 edges.
// The 1e-4 margin covers the different rounding of the per-pixel evaluation in the blend kernel.
#pragma once
#include "g4d_common.cuh"

namespace g4d {

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
    return rect_contributes(r0, r1, (float)(tx * kTile), (float)(tx * kTile + kTile - 1), (float)(ty * kTile),
                            (float)(ty * kTile + kTile - 1));
}

// Warp-cooperative walk over the tile rects of a warp's 32 Gaussians: each visible one is broadcast by shuffle and the
// 32 lanes take 32 tiles of its rect at a time -- a thread-per-Gaussian loop would serialise on the largest rect.
// Rows are clipped to the tile-row band [band_y0, band_y1) the caller is working on.
struct TileJob { float4 r0, r1; int minx, miny, w, ntiles; };
G4D_D TileJob bcast_job(const float4& r0, const float4& r1, uint2 rc, int src, int band_y0, int band_y1) {
    TileJob j;
    j.r0.x = __shfl_sync(0xffffffffu, r0.x, src); j.r0.y = __shfl_sync(0xffffffffu, r0.y, src);
    j.r0.z = __shfl_sync(0xffffffffu, r0.z, src); j.r0.w = __shfl_sync(0xffffffffu, r0.w, src);
    j.r1.x = __shfl_sync(0xffffffffu, r1.x, src); j.r1.y = __shfl_sync(0xffffffffu, r1.y, src);
    j.r1.z = 0.f; j.r1.w = 0.f;
    const uint32_t rx = __shfl_sync(0xffffffffu, rc.x, src), ry = __shfl_sync(0xffffffffu, rc.y, src);
    j.minx = (int)(rx & 0xFFFFu);
    const int maxx = (int)(ry & 0xFFFFu);
    const int miny = max((int)(rx >> 16), band_y0), maxy = min((int)(ry >> 16), band_y1);
    j.miny = miny;
    j.w = maxx - j.minx;
    j.ntiles = maxy > miny ? j.w * (maxy - miny) : 0;
    return j;
}

}  // namespace g4d
