/*
 * CPU oracle for the rasterizer half of the hot path (tile-based 3D Gaussian splatting with the
 * extra depth output), forward AND backward.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or executed from the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * PARITY UNPINNED: the algorithm lives in a third-party dependency that is ABSENT from
 * /root/reference (empty submodule dir):
 *     ingra14m/depth-diff-gaussian-rasterization @ 9055fcfdde3c08e0ed602ded03d4379ee90b8986
 *     (/root/reference/.gitmodules:5-7; fork of graphdeco-inria/diff-gaussian-rasterization).
 * The reference ships no tests, golden vectors or fixtures for it.  This file restates the published
 * algorithm as specified in SURVEY.md Appendix A (A.1 preprocess, A.2 binning, A.3 blend forward,
 * A.4 backward) and is anchored on the reference's own call sites:
 *     /root/reference/gaussian_renderer/__init__.py:38-51   (settings fields)
 *     /root/reference/gaussian_renderer/__init__.py:120-128 (inputs / 3 outputs)
 *     /root/reference/scene/cameras.py:59-64                (row-vector view / proj matrices)
 * and cross-checked against in-tree sources of the same math:
 *     /root/reference/utils/sh_utils.py:57-112      (SH polynomial + constants)
 *     /root/reference/utils/general_utils.py:84-116 (quaternion -> rotation, Sigma = R S S^T R^T)
 * tests/test_oracle_raster.py additionally pins this file against an independent dense fp64
 * torch-autograd formulation (oracle/dense_ref.py) and closed-form cases.
 *
 * Numerics: fp32 arithmetic in the order written here; build with -ffp-contract=off so that the
 * index-producing stages (depth bits, radii, tile rects, sort keys) are reproducible bit-for-bit
 * by an implementation that evaluates the same expression trees without FMA contraction.
 * Gradient accumulations over pixels are done in double precision (the oracle is the yardstick).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16
#define SH_STRIDE 16 /* coefficients per channel slot: shs is [N,16,3] */

static const float kC0 = 0.28209479177387814f;
static const float kC1 = 0.4886025119029199f;
static const float kC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                             -1.0925484305920792f, 0.5462742152960396f};
static const float kC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                             -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

typedef struct {
    int32_t H, W, sh_degree, pad_;
    float tanfovx, tanfovy, scale_modifier, pad2_;
    float view[16]; /* world_view_transform, row-vector convention: x' = v0 x + v4 y + v8 z + v12 */
    float proj[16]; /* full_proj_transform, same convention */
    float campos[3];
    float bg[3];
} G4DRefCam;

static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
static inline int imin_(int a, int b) { return a < b ? a : b; }
static inline int imax_(int a, int b) { return a > b ? a : b; }
/* float -> int with round-toward-zero, saturating, NaN -> 0 (defined behaviour for absurd inputs) */
static inline int f2i_sat(float f) {
    if (!(f == f)) return 0;
    if (f >= 2147483520.f) return 2147483647;
    if (f <= -2147483648.f) return (-2147483647 - 1);
    return (int)f;
}

/* ---- A.1 helpers ------------------------------------------------------------------------- */
static void point4x3(const float *m, const float *p, float *o) {
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static void point4x4(const float *m, const float *p, float *o) {
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
    o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}

/* rotation matrix rows from quaternion (w,x,y,z) used as given (no normalisation) */
static void quat_to_rot(const float *q, float R[9]) {
    float r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1.f - 2.f * (y * y + z * z);
    R[1] = 2.f * (x * y - r * z);
    R[2] = 2.f * (x * z + r * y);
    R[3] = 2.f * (x * y + r * z);
    R[4] = 1.f - 2.f * (x * x + z * z);
    R[5] = 2.f * (y * z - r * x);
    R[6] = 2.f * (x * z - r * y);
    R[7] = 2.f * (y * z + r * x);
    R[8] = 1.f - 2.f * (x * x + y * y);
}

/* Sigma = R S S^T R^T; M[k][j] = s_k * R[j][k]; Sigma_ij = sum_k M[k][i] M[k][j].
 * Output order (xx, xy, xz, yy, yz, zz). */
static void cov3d_from_scale_rot(const float *scale, float mod, const float *q, float c[6]) {
    float R[9], M[9];
    quat_to_rot(q, R);
    float s[3] = {mod * scale[0], mod * scale[1], mod * scale[2]};
    for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 3; ++j) M[k * 3 + j] = s[k] * R[j * 3 + k];
    c[0] = M[0] * M[0] + M[3] * M[3] + M[6] * M[6];
    c[1] = M[0] * M[1] + M[3] * M[4] + M[6] * M[7];
    c[2] = M[0] * M[2] + M[3] * M[5] + M[6] * M[8];
    c[3] = M[1] * M[1] + M[4] * M[4] + M[7] * M[7];
    c[4] = M[1] * M[2] + M[4] * M[5] + M[7] * M[8];
    c[5] = M[2] * M[2] + M[5] * M[5] + M[8] * M[8];
}

/* EWA projection.  Returns the 2x3 matrix T = J*W (rows t0,t1) and cov2D (a,b,c) WITHOUT dilation;
 * clampx/clampy report whether the 1.3*tanfov guard band clamped t.x / t.y. */
static void cov2d_project(const G4DRefCam *cam, const float *pview, const float cov3[6], float fx, float fy,
                          float T0[3], float T1[3], float cov2[3], int *clampx, int *clampy, float tclamped[3]) {
    const float *v = cam->view;
    float limx = 1.3f * cam->tanfovx, limy = 1.3f * cam->tanfovy;
    float tz = pview[2];
    float txtz = pview[0] / tz, tytz = pview[1] / tz;
    *clampx = (txtz < -limx) || (txtz > limx);
    *clampy = (tytz < -limy) || (tytz > limy);
    float tx = fminf_(limx, fmaxf_(-limx, txtz)) * tz;
    float ty = fminf_(limy, fmaxf_(-limy, tytz)) * tz;
    tclamped[0] = tx; tclamped[1] = ty; tclamped[2] = tz;
    float j00 = fx / tz, j02 = -(fx * tx) / (tz * tz);
    float j11 = fy / tz, j12 = -(fy * ty) / (tz * tz);
    /* W rows: (v0,v4,v8), (v1,v5,v9), (v2,v6,v10) */
    T0[0] = j00 * v[0] + j02 * v[2];
    T0[1] = j00 * v[4] + j02 * v[6];
    T0[2] = j00 * v[8] + j02 * v[10];
    T1[0] = j11 * v[1] + j12 * v[2];
    T1[1] = j11 * v[5] + j12 * v[6];
    T1[2] = j11 * v[9] + j12 * v[10];
    /* u = Sigma * T0^T, w = Sigma * T1^T */
    float u0 = cov3[0] * T0[0] + cov3[1] * T0[1] + cov3[2] * T0[2];
    float u1 = cov3[1] * T0[0] + cov3[3] * T0[1] + cov3[4] * T0[2];
    float u2 = cov3[2] * T0[0] + cov3[4] * T0[1] + cov3[5] * T0[2];
    float w0 = cov3[0] * T1[0] + cov3[1] * T1[1] + cov3[2] * T1[2];
    float w1 = cov3[1] * T1[0] + cov3[3] * T1[1] + cov3[4] * T1[2];
    float w2 = cov3[2] * T1[0] + cov3[4] * T1[1] + cov3[5] * T1[2];
    cov2[0] = T0[0] * u0 + T0[1] * u1 + T0[2] * u2;
    cov2[1] = T0[0] * w0 + T0[1] * w1 + T0[2] * w2;
    cov2[2] = T1[0] * w0 + T1[1] * w1 + T1[2] * w2;
}

/* SH basis values for a unit direction; b[0..15]; only the first (deg+1)^2 are meaningful */
static void sh_basis(int deg, float x, float y, float z, float b[16]) {
    b[0] = kC0;
    if (deg > 0) {
        b[1] = -kC1 * y; b[2] = kC1 * z; b[3] = -kC1 * x;
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = kC2[0] * xy; b[5] = kC2[1] * yz; b[6] = kC2[2] * (2.f * zz - xx - yy);
            b[7] = kC2[3] * xz; b[8] = kC2[4] * (xx - yy);
            if (deg > 2) {
                b[9] = kC3[0] * y * (3.f * xx - yy);
                b[10] = kC3[1] * xy * z;
                b[11] = kC3[2] * y * (4.f * zz - xx - yy);
                b[12] = kC3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                b[13] = kC3[4] * x * (4.f * zz - xx - yy);
                b[14] = kC3[5] * z * (xx - yy);
                b[15] = kC3[6] * x * (xx - 3.f * yy);
            }
        }
    }
}

/* ---- A.1 preprocess ------------------------------------------------------------------------
 * Outputs (all [N,...]): depth, radii, xy[2], cov3d[6], conic_op[4], rgb[3], clamped[3] (u8),
 * rect[4] = (min_x, min_y, max_x, max_y) in tiles, tiles_touched. */
void g4dref_preprocess(const G4DRefCam *cam, int N, const float *means3D, const float *scales, const float *rots,
                       const float *opac, const float *shs, float *depth, int32_t *radii, float *xy, float *cov3d,
                       float *conic_op, float *rgb, uint8_t *clamped, int32_t *rect, uint32_t *tiles_touched) {
    const int gx = (cam->W + TILE - 1) / TILE, gy = (cam->H + TILE - 1) / TILE;
    const float fx = (float)cam->W / (2.f * cam->tanfovx), fy = (float)cam->H / (2.f * cam->tanfovy);
    const int ncoef = (cam->sh_degree + 1) * (cam->sh_degree + 1);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        const float *p = means3D + 3 * i;
        radii[i] = 0; tiles_touched[i] = 0; depth[i] = 0.f;
        xy[2 * i] = xy[2 * i + 1] = 0.f;
        for (int k = 0; k < 6; ++k) cov3d[6 * i + k] = 0.f;
        for (int k = 0; k < 4; ++k) { conic_op[4 * i + k] = 0.f; rect[4 * i + k] = 0; }
        for (int k = 0; k < 3; ++k) { rgb[3 * i + k] = 0.f; clamped[3 * i + k] = 0; }

        float pv[3], ph[4];
        point4x3(cam->view, p, pv);
        if (pv[2] <= 0.2f) continue; /* near cull (x/y frustum test is disabled upstream) */
        point4x4(cam->proj, p, ph);
        float pw = 1.0f / (ph[3] + 0.0000001f);
        float ndcx = ph[0] * pw, ndcy = ph[1] * pw;

        float c3[6];
        cov3d_from_scale_rot(scales + 3 * i, cam->scale_modifier, rots + 4 * i, c3);
        for (int k = 0; k < 6; ++k) cov3d[6 * i + k] = c3[k];

        float T0[3], T1[3], c2[3], tcl[3]; int cx, cy;
        cov2d_project(cam, pv, c3, fx, fy, T0, T1, c2, &cx, &cy, tcl);
        float a = c2[0] + 0.3f, b = c2[1], c = c2[2] + 0.3f;
        float det = a * c - b * b;
        if (det == 0.0f) continue;
        float det_inv = 1.f / det;
        float conx = c * det_inv, cony = -b * det_inv, conz = a * det_inv;
        float mid = 0.5f * (a + c);
        float root = sqrtf(fmaxf_(0.1f, mid * mid - det));
        float lam1 = mid + root, lam2 = mid - root;
        float rad = ceilf(3.f * sqrtf(fmaxf_(lam1, lam2)));
        float px = ((ndcx + 1.0f) * (float)cam->W - 1.0f) * 0.5f;
        float py = ((ndcy + 1.0f) * (float)cam->H - 1.0f) * 0.5f;
        int rminx = imin_(gx, imax_(0, f2i_sat((px - rad) / (float)TILE)));
        int rminy = imin_(gy, imax_(0, f2i_sat((py - rad) / (float)TILE)));
        int rmaxx = imin_(gx, imax_(0, f2i_sat((px + rad + (float)(TILE - 1)) / (float)TILE)));
        int rmaxy = imin_(gy, imax_(0, f2i_sat((py + rad + (float)(TILE - 1)) / (float)TILE)));
        int area = (rmaxx - rminx) * (rmaxy - rminy);
        if (area == 0) continue;

        /* SH -> RGB at the (deformed) position */
        float dx = p[0] - cam->campos[0], dy = p[1] - cam->campos[1], dz = p[2] - cam->campos[2];
        float len = sqrtf(dx * dx + dy * dy + dz * dz);
        dx = dx / len; dy = dy / len; dz = dz / len;
        float bas[16];
        sh_basis(cam->sh_degree, dx, dy, dz, bas);
        const float *sh = shs + (size_t)i * SH_STRIDE * 3;
        for (int ch = 0; ch < 3; ++ch) {
            float acc = bas[0] * sh[ch];
            for (int k = 1; k < ncoef; ++k) acc = acc + bas[k] * sh[3 * k + ch];
            acc = acc + 0.5f;
            clamped[3 * i + ch] = acc < 0.f;
            rgb[3 * i + ch] = fmaxf_(acc, 0.f);
        }
        depth[i] = pv[2];
        radii[i] = f2i_sat(rad);
        xy[2 * i] = px; xy[2 * i + 1] = py;
        conic_op[4 * i] = conx; conic_op[4 * i + 1] = cony; conic_op[4 * i + 2] = conz; conic_op[4 * i + 3] = opac[i];
        rect[4 * i] = rminx; rect[4 * i + 1] = rminy; rect[4 * i + 2] = rmaxx; rect[4 * i + 3] = rmaxy;
        tiles_touched[i] = (uint32_t)area;
    }
}

/* ---- A.2 binning ------------------------------------------------------------------------- */
int64_t g4dref_count_instances(int N, const uint32_t *tiles_touched) {
    int64_t r = 0;
    for (int i = 0; i < N; ++i) r += tiles_touched[i];
    return r;
}

static void radix_sort_pairs(uint64_t *keys, uint32_t *vals, int64_t n, int bits) {
    uint64_t *k2 = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
    uint32_t *v2 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    uint64_t *ka = keys, *kb = k2; uint32_t *va = vals, *vb = v2;
    for (int shift = 0; shift < bits; shift += 8) {
        int64_t cnt[257]; memset(cnt, 0, sizeof(cnt));
        for (int64_t i = 0; i < n; ++i) cnt[((ka[i] >> shift) & 0xFF) + 1]++;
        for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
        for (int64_t i = 0; i < n; ++i) {
            int64_t dst = cnt[(ka[i] >> shift) & 0xFF]++;
            kb[dst] = ka[i]; vb[dst] = va[i];
        }
        uint64_t *tk = ka; ka = kb; kb = tk;
        uint32_t *tv = va; va = vb; vb = tv;
    }
    if (ka != keys) { memcpy(keys, ka, sizeof(uint64_t) * (size_t)n); memcpy(vals, va, sizeof(uint32_t) * (size_t)n); }
    free(k2); free(v2);
}

/* keys_sorted / ids_sorted have R entries; ranges is [tiles,2] (start,end), (0,0) for empty tiles.
 * Order: tile-major, depth-ascending (positive float bits compare as uints), ties by Gaussian index. */
void g4dref_bin(const G4DRefCam *cam, int N, const float *depth, const int32_t *rect, const uint32_t *tiles_touched,
                int64_t R, uint64_t *keys_sorted, uint32_t *ids_sorted, uint32_t *ranges) {
    const int gx = (cam->W + TILE - 1) / TILE, gy = (cam->H + TILE - 1) / TILE;
    int64_t off = 0;
    for (int i = 0; i < N; ++i) {
        if (tiles_touched[i] == 0) continue;
        uint32_t dbits; memcpy(&dbits, depth + i, 4);
        for (int y = rect[4 * i + 1]; y < rect[4 * i + 3]; ++y)
            for (int x = rect[4 * i]; x < rect[4 * i + 2]; ++x) {
                uint64_t key = (uint64_t)(uint32_t)(y * gx + x);
                keys_sorted[off] = (key << 32) | dbits;
                ids_sorted[off] = (uint32_t)i;
                ++off;
            }
    }
    int tiles = gx * gy, tbits = 0;
    while ((1 << tbits) < tiles) ++tbits;
    radix_sort_pairs(keys_sorted, ids_sorted, R, 32 + ((tbits + 7) / 8) * 8);
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)tiles);
    for (int64_t i = 0; i < R; ++i) {
        uint32_t t = (uint32_t)(keys_sorted[i] >> 32);
        if (i == 0 || (uint32_t)(keys_sorted[i - 1] >> 32) != t) ranges[2 * t] = (uint32_t)i;
        if (i == R - 1 || (uint32_t)(keys_sorted[i + 1] >> 32) != t) ranges[2 * t + 1] = (uint32_t)(i + 1);
    }
}

/* ---- A.3 blend forward -------------------------------------------------------------------- */
void g4dref_blend_forward(const G4DRefCam *cam, const uint32_t *ids_sorted, const uint32_t *ranges, const float *xy,
                          const float *conic_op, const float *rgb, const float *depth, float *out_color,
                          float *out_depth, float *final_T, uint32_t *n_contrib) {
    const int H = cam->H, W = cam->W;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; ++tile) {
        int tx = tile % gx, ty = tile / gx;
        uint32_t beg = ranges[2 * tile], end = ranges[2 * tile + 1];
        for (int ly = 0; ly < TILE; ++ly)
            for (int lx = 0; lx < TILE; ++lx) {
                int pxi = tx * TILE + lx, pyi = ty * TILE + ly;
                if (pxi >= W || pyi >= H) continue;
                float pxf = (float)pxi, pyf = (float)pyi;
                float T = 1.f, C[3] = {0.f, 0.f, 0.f}, D = 0.f;
                uint32_t contributor = 0, last = 0;
                for (uint32_t k = beg; k < end; ++k) {
                    uint32_t g = ids_sorted[k];
                    contributor++;
                    float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
                    const float *co = conic_op + 4 * g;
                    float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0.f) continue;
                    float alpha = fminf_(0.99f, co[3] * expf(power));
                    if (alpha < 1.f / 255.f) continue;
                    float test_T = T * (1.f - alpha);
                    if (test_T < 0.0001f) break;
                    float w = alpha * T;
                    C[0] += rgb[3 * g] * w; C[1] += rgb[3 * g + 1] * w; C[2] += rgb[3 * g + 2] * w;
                    D += depth[g] * w;
                    T = test_T;
                    last = contributor;
                }
                size_t pix = (size_t)pyi * W + pxi;
                final_T[pix] = T; n_contrib[pix] = last;
                for (int ch = 0; ch < 3; ++ch) out_color[(size_t)ch * H * W + pix] = C[ch] + T * cam->bg[ch];
                out_depth[pix] = D;
            }
    }
}

/* ---- A.4 blend backward ------------------------------------------------------------------- *
 * dL_dpix is [3,H,W].  Outputs (zeroed here):
 *   g_mean2D [N,2]  d/d(pixel coordinate) * (0.5W, 0.5H)   (i.e. w.r.t. NDC; what means2D.grad carries)
 *   g_conic  [N,3]  d/d(conic.x), d/d(conic.y), d/d(conic.z) with power = -0.5(x dx^2 + z dy^2) - y dx dy
 *   g_opac   [N]    d/d(opacity)
 *   g_rgb    [N,3]  d/d(rgb)
 * The alpha clamp min(0.99, .) passes gradient through (as the published algorithm does); the depth
 * image has no gradient path. */
void g4dref_blend_backward(const G4DRefCam *cam, int N, const uint32_t *ids_sorted, const uint32_t *ranges,
                           const float *xy, const float *conic_op, const float *rgb, const float *final_T,
                           const uint32_t *n_contrib, const float *dL_dpix, float *g_mean2D, float *g_conic,
                           float *g_opac, float *g_rgb) {
    const int H = cam->H, W = cam->W;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    double *acc = (double *)calloc((size_t)N * 9, sizeof(double));
    const double ddx = 0.5 * W, ddy = 0.5 * H;
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; ++tile) {
        int tx = tile % gx, ty = tile / gx;
        uint32_t beg = ranges[2 * tile];
        for (int ly = 0; ly < TILE; ++ly)
            for (int lx = 0; lx < TILE; ++lx) {
                int pxi = tx * TILE + lx, pyi = ty * TILE + ly;
                if (pxi >= W || pyi >= H) continue;
                size_t pix = (size_t)pyi * W + pxi;
                float pxf = (float)pxi, pyf = (float)pyi;
                float Tfin = final_T[pix], T = Tfin;
                uint32_t last = n_contrib[pix];
                float dpix[3] = {dL_dpix[pix], dL_dpix[(size_t)H * W + pix], dL_dpix[(size_t)2 * H * W + pix]};
                float bgdot = cam->bg[0] * dpix[0] + cam->bg[1] * dpix[1] + cam->bg[2] * dpix[2];
                float accum[3] = {0.f, 0.f, 0.f}, lastc[3] = {0.f, 0.f, 0.f}, last_alpha = 0.f;
                for (uint32_t k = beg + last; k-- > beg;) { /* back to front over the first `last` entries */
                    uint32_t g = ids_sorted[k];
                    float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
                    const float *co = conic_op + 4 * g;
                    float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0.f) continue;
                    float G = expf(power);
                    float alpha = fminf_(0.99f, co[3] * G);
                    if (alpha < 1.f / 255.f) continue;
                    T = T / (1.f - alpha);
                    float w = alpha * T;
                    float dL_dalpha = 0.f;
                    double *a = acc + (size_t)g * 9;
                    for (int ch = 0; ch < 3; ++ch) {
                        float c = rgb[3 * g + ch];
                        accum[ch] = last_alpha * lastc[ch] + (1.f - last_alpha) * accum[ch];
                        lastc[ch] = c;
                        dL_dalpha += (c - accum[ch]) * dpix[ch];
#pragma omp atomic
                        a[6 + ch] += (double)(w * dpix[ch]);
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-Tfin / (1.f - alpha)) * bgdot;
                    float dL_dG = co[3] * dL_dalpha;
                    float gdx = G * dx, gdy = G * dy;
                    float dG_ddx = -gdx * co[0] - gdy * co[1];
                    float dG_ddy = -gdy * co[2] - gdx * co[1];
#pragma omp atomic
                    a[0] += (double)(dL_dG * dG_ddx) * ddx;
#pragma omp atomic
                    a[1] += (double)(dL_dG * dG_ddy) * ddy;
#pragma omp atomic
                    a[2] += (double)(-0.5f * gdx * dx * dL_dG);
#pragma omp atomic
                    a[3] += (double)(-gdx * dy * dL_dG);
#pragma omp atomic
                    a[4] += (double)(-0.5f * gdy * dy * dL_dG);
#pragma omp atomic
                    a[5] += (double)(G * dL_dalpha);
                }
            }
    }
    for (int i = 0; i < N; ++i) {
        const double *a = acc + (size_t)i * 9;
        g_mean2D[2 * i] = (float)a[0]; g_mean2D[2 * i + 1] = (float)a[1];
        g_conic[3 * i] = (float)a[2]; g_conic[3 * i + 1] = (float)a[3]; g_conic[3 * i + 2] = (float)a[4];
        g_opac[i] = (float)a[5];
        g_rgb[3 * i] = (float)a[6]; g_rgb[3 * i + 1] = (float)a[7]; g_rgb[3 * i + 2] = (float)a[8];
    }
    free(acc);
}

/* ---- A.4 per-Gaussian backward ------------------------------------------------------------- *
 * From (g_mean2D, g_conic, g_rgb) to gradients of means3D, scales, rots (quaternion as given) and shs.
 * Quirks kept from the published algorithm: the 1.3*tanfov clamp makes the clamped t.x/t.y a constant
 * (no gradient to t.x/t.y nor through it to t.z); colour gradient is zero where the forward clamped at 0;
 * 1/(det^2 + 1e-7) regularised inverse; the radius and depth carry no gradient.  d/dscale includes the
 * scale_modifier factor (mathematically exact; the published code omits it, identical at modifier 1). */
void g4dref_preprocess_backward(const G4DRefCam *cam, int N, const float *means3D, const float *scales,
                                const float *rots, const float *shs, const int32_t *radii, const uint8_t *clamped,
                                const float *g_mean2D, const float *g_conic, const float *g_rgb, float *g_means3D,
                                float *g_scales, float *g_rots, float *g_shs) {
    const float fx = (float)cam->W / (2.f * cam->tanfovx), fy = (float)cam->H / (2.f * cam->tanfovy);
    const int deg = cam->sh_degree, ncoef = (deg + 1) * (deg + 1);
    const float *v = cam->view, *pm = cam->proj;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        for (int k = 0; k < 3; ++k) { g_means3D[3 * i + k] = 0.f; g_scales[3 * i + k] = 0.f; }
        for (int k = 0; k < 4; ++k) g_rots[4 * i + k] = 0.f;
        for (int k = 0; k < SH_STRIDE * 3; ++k) g_shs[(size_t)i * SH_STRIDE * 3 + k] = 0.f;
        if (!(radii[i] > 0)) continue;
        const float *p = means3D + 3 * i;
        double gm[3] = {0, 0, 0};

        /* (1) conic -> cov2D -> (Sigma, T) */
        float pv[3]; point4x3(v, p, pv);
        float c3[6]; cov3d_from_scale_rot(scales + 3 * i, cam->scale_modifier, rots + 4 * i, c3);
        float T0[3], T1[3], c2[3], tcl[3]; int cx, cy;
        cov2d_project(cam, pv, c3, fx, fy, T0, T1, c2, &cx, &cy, tcl);
        double a = (double)c2[0] + 0.3, b = c2[1], c = (double)c2[2] + 0.3;
        double det = a * c - b * b;
        double d2inv = 1.0 / (det * det + 0.0000001);
        double gcx = g_conic[3 * i], gcy = g_conic[3 * i + 1], gcz = g_conic[3 * i + 2];
        /* conic = (c, -b, a)/det */
        double dL_da = d2inv * (-c * c * gcx + b * c * gcy + (det - a * c) * gcz);
        double dL_dc = d2inv * (-a * a * gcz + a * b * gcy + (det - a * c) * gcx);
        double dL_db = d2inv * (2.0 * b * c * gcx - (det + 2.0 * b * b) * gcy + 2.0 * a * b * gcz);
        /* cov2 = T Sigma T^T :  dL/dSigma_jk (full symmetric matrix S_) */
        double t0[3] = {T0[0], T0[1], T0[2]}, t1[3] = {T1[0], T1[1], T1[2]};
        double gS[3][3];
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k)
                gS[j][k] = dL_da * t0[j] * t0[k] + dL_dc * t1[j] * t1[k] + 0.5 * dL_db * (t0[j] * t1[k] + t1[j] * t0[k]);
        double S3[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
        double gT0[3], gT1[3];
        for (int j = 0; j < 3; ++j) {
            double s0 = S3[j][0] * t0[0] + S3[j][1] * t0[1] + S3[j][2] * t0[2];
            double s1 = S3[j][0] * t1[0] + S3[j][1] * t1[1] + S3[j][2] * t1[2];
            gT0[j] = 2.0 * dL_da * s0 + dL_db * s1;
            gT1[j] = 2.0 * dL_dc * s1 + dL_db * s0;
        }
        /* T = J W:  T0 = j00*Wrow0 + j02*Wrow2 ; T1 = j11*Wrow1 + j12*Wrow2 */
        double Wr[3][3] = {{v[0], v[4], v[8]}, {v[1], v[5], v[9]}, {v[2], v[6], v[10]}};
        double gJ00 = 0, gJ02 = 0, gJ11 = 0, gJ12 = 0;
        for (int j = 0; j < 3; ++j) {
            gJ00 += Wr[0][j] * gT0[j]; gJ02 += Wr[2][j] * gT0[j];
            gJ11 += Wr[1][j] * gT1[j]; gJ12 += Wr[2][j] * gT1[j];
        }
        double tx = tcl[0], ty = tcl[1], tz = tcl[2];
        double itz = 1.0 / tz, itz2 = itz * itz, itz3 = itz2 * itz;
        double gtx = cx ? 0.0 : -fx * itz2 * gJ02;
        double gty = cy ? 0.0 : -fy * itz2 * gJ12;
        double gtz = -fx * itz2 * gJ00 - fy * itz2 * gJ11 + (2.0 * fx * tx) * itz3 * gJ02 + (2.0 * fy * ty) * itz3 * gJ12;
        /* t = W p + trans  =>  dL/dp = W^T gt */
        gm[0] += Wr[0][0] * gtx + Wr[1][0] * gty + Wr[2][0] * gtz;
        gm[1] += Wr[0][1] * gtx + Wr[1][1] * gty + Wr[2][1] * gtz;
        gm[2] += Wr[0][2] * gtx + Wr[1][2] * gty + Wr[2][2] * gtz;

        /* (2) mean2D (NDC units) -> mean3D through the projective divide */
        {
            float ph[4]; point4x4(pm, p, ph);
            double mw = 1.0 / ((double)ph[3] + 0.0000001);
            double mul1 = (double)ph[0] * mw * mw, mul2 = (double)ph[1] * mw * mw;
            double g0 = g_mean2D[2 * i], g1 = g_mean2D[2 * i + 1];
            gm[0] += (pm[0] * mw - pm[3] * mul1) * g0 + (pm[1] * mw - pm[3] * mul2) * g1;
            gm[1] += (pm[4] * mw - pm[7] * mul1) * g0 + (pm[5] * mw - pm[7] * mul2) * g1;
            gm[2] += (pm[8] * mw - pm[11] * mul1) * g0 + (pm[9] * mw - pm[11] * mul2) * g1;
        }

        /* (3) colour -> SH coefficients and view direction -> mean3D */
        {
            double d0 = (double)p[0] - cam->campos[0], d1 = (double)p[1] - cam->campos[1], d2 = (double)p[2] - cam->campos[2];
            double len = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            double x = d0 / len, y = d1 / len, z = d2 / len;
            float bas[16]; sh_basis(deg, (float)x, (float)y, (float)z, bas);
            double gc[3];
            for (int ch = 0; ch < 3; ++ch) gc[ch] = clamped[3 * i + ch] ? 0.0 : (double)g_rgb[3 * i + ch];
            const float *sh = shs + (size_t)i * SH_STRIDE * 3;
            float *gsh = g_shs + (size_t)i * SH_STRIDE * 3;
            for (int k = 0; k < ncoef; ++k)
                for (int ch = 0; ch < 3; ++ch) gsh[3 * k + ch] = (float)((double)bas[k] * gc[ch]);
            /* d(basis_k)/d(x,y,z), analytic */
            double db[16][3]; memset(db, 0, sizeof(db));
            if (deg > 0) {
                db[1][1] = -kC1; db[2][2] = kC1; db[3][0] = -kC1;
                if (deg > 1) {
                    db[4][0] = kC2[0] * y; db[4][1] = kC2[0] * x;
                    db[5][1] = kC2[1] * z; db[5][2] = kC2[1] * y;
                    db[6][0] = kC2[2] * -2.0 * x; db[6][1] = kC2[2] * -2.0 * y; db[6][2] = kC2[2] * 4.0 * z;
                    db[7][0] = kC2[3] * z; db[7][2] = kC2[3] * x;
                    db[8][0] = kC2[4] * 2.0 * x; db[8][1] = kC2[4] * -2.0 * y;
                    if (deg > 2) {
                        double xx = x * x, yy = y * y, zz = z * z;
                        db[9][0] = kC3[0] * 6.0 * x * y; db[9][1] = kC3[0] * (3.0 * xx - 3.0 * yy);
                        db[10][0] = kC3[1] * y * z; db[10][1] = kC3[1] * x * z; db[10][2] = kC3[1] * x * y;
                        db[11][0] = kC3[2] * -2.0 * x * y; db[11][1] = kC3[2] * (4.0 * zz - xx - 3.0 * yy); db[11][2] = kC3[2] * 8.0 * y * z;
                        db[12][0] = kC3[3] * -6.0 * x * z; db[12][1] = kC3[3] * -6.0 * y * z; db[12][2] = kC3[3] * (6.0 * zz - 3.0 * xx - 3.0 * yy);
                        db[13][0] = kC3[4] * (4.0 * zz - 3.0 * xx - yy); db[13][1] = kC3[4] * -2.0 * x * y; db[13][2] = kC3[4] * 8.0 * x * z;
                        db[14][0] = kC3[5] * 2.0 * x * z; db[14][1] = kC3[5] * -2.0 * y * z; db[14][2] = kC3[5] * (xx - yy);
                        db[15][0] = kC3[6] * (3.0 * xx - 3.0 * yy); db[15][1] = kC3[6] * -6.0 * x * y;
                    }
                }
            }
            double gdir[3] = {0, 0, 0};
            for (int k = 1; k < ncoef; ++k) {
                double s = sh[3 * k] * gc[0] + sh[3 * k + 1] * gc[1] + sh[3 * k + 2] * gc[2];
                gdir[0] += db[k][0] * s; gdir[1] += db[k][1] * s; gdir[2] += db[k][2] * s;
            }
            /* d(normalize(d))/dd = (I - n n^T)/len */
            double dot = x * gdir[0] + y * gdir[1] + z * gdir[2];
            gm[0] += (gdir[0] - x * dot) / len; gm[1] += (gdir[1] - y * dot) / len; gm[2] += (gdir[2] - z * dot) / len;
        }
        g_means3D[3 * i] = (float)gm[0]; g_means3D[3 * i + 1] = (float)gm[1]; g_means3D[3 * i + 2] = (float)gm[2];

        /* (4) Sigma -> scale, quaternion.  Sigma = R diag(s^2) R^T, s = mod * scale. */
        {
            float Rf[9]; quat_to_rot(rots + 4 * i, Rf);
            double R[3][3] = {{Rf[0], Rf[1], Rf[2]}, {Rf[3], Rf[4], Rf[5]}, {Rf[6], Rf[7], Rf[8]}};
            double mod = cam->scale_modifier;
            double s[3] = {mod * scales[3 * i], mod * scales[3 * i + 1], mod * scales[3 * i + 2]};
            /* dL/ds_k = 2 s_k * r_k^T gS r_k  (r_k = column k of R) ; dL/dR = 2 gS R diag(s^2) */
            double gR[3][3];
            for (int k = 0; k < 3; ++k) {
                double q = 0;
                for (int a_ = 0; a_ < 3; ++a_)
                    for (int b_ = 0; b_ < 3; ++b_) q += R[a_][k] * gS[a_][b_] * R[b_][k];
                g_scales[3 * i + k] = (float)(2.0 * s[k] * q * mod);
            }
            for (int a_ = 0; a_ < 3; ++a_)
                for (int k = 0; k < 3; ++k) {
                    double q = 0;
                    for (int b_ = 0; b_ < 3; ++b_) q += gS[a_][b_] * R[b_][k];
                    gR[a_][k] = 2.0 * q * s[k] * s[k];
                }
            double r = rots[4 * i], x = rots[4 * i + 1], y = rots[4 * i + 2], z = rots[4 * i + 3];
            double gr = 2.0 * (-z * gR[0][1] + y * gR[0][2] + z * gR[1][0] - x * gR[1][2] - y * gR[2][0] + x * gR[2][1]);
            double gx_ = 2.0 * (y * gR[0][1] + z * gR[0][2] + y * gR[1][0] - 2.0 * x * gR[1][1] - r * gR[1][2] + z * gR[2][0] + r * gR[2][1] - 2.0 * x * gR[2][2]);
            double gy_ = 2.0 * (-2.0 * y * gR[0][0] + x * gR[0][1] + r * gR[0][2] + x * gR[1][0] + z * gR[1][2] - r * gR[2][0] + z * gR[2][1] - 2.0 * y * gR[2][2]);
            double gz_ = 2.0 * (-2.0 * z * gR[0][0] - r * gR[0][1] + x * gR[0][2] + r * gR[1][0] - 2.0 * z * gR[1][1] + y * gR[1][2] + x * gR[2][0] + y * gR[2][1]);
            g_rots[4 * i] = (float)gr; g_rots[4 * i + 1] = (float)gx_; g_rots[4 * i + 2] = (float)gy_; g_rots[4 * i + 3] = (float)gz_;
        }
    }
}

/* size of the camera struct, so the Python side can assert its ctypes mirror */
int g4dref_cam_sizeof(void) { return (int)sizeof(G4DRefCam); }
