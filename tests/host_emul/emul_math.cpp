// Host build of 4dgaussians_b200/csrc/g4d_math.cuh (the per-Gaussian device math) so that its logic can
// be checked against the oracle in the GPU-less build container.  TEST-ONLY: the product never runs this.
#include <cstring>
#include "../../4dgaussians_b200/csrc/g4d_math.cuh"

using namespace g4d;

struct RefCam {  // mirror of oracle/raster_ref.c::G4DRefCam
    int32_t H, W, sh_degree, pad_;
    float tanfovx, tanfovy, scale_modifier, pad2_;
    float view[16], proj[16], campos[3], bg[3];
};

static CameraDev to_dev(const RefCam* c) {
    CameraDev d;
    std::memset(&d, 0, sizeof(d));
    d.H = c->H; d.W = c->W; d.sh_degree = c->sh_degree;
    d.grid_x = (c->W + kTile - 1) / kTile; d.grid_y = (c->H + kTile - 1) / kTile; d.num_tiles = d.grid_x * d.grid_y;
    d.tanfovx = c->tanfovx; d.tanfovy = c->tanfovy; d.scale_modifier = c->scale_modifier;
    d.focal_x = (float)c->W / (2.f * c->tanfovx); d.focal_y = (float)c->H / (2.f * c->tanfovy);
    std::memcpy(d.view, c->view, 64); std::memcpy(d.proj, c->proj, 64);
    for (int i = 0; i < 3; ++i) { d.campos[i] = c->campos[i]; d.bg[i] = c->bg[i]; }
    return d;
}

extern "C" void emul_preprocess(const RefCam* rc, int n, const float* means, const float* scales, const float* rots,
                                const float* opac, const float* shs, float* depth, int32_t* radii, float* xy,
                                float* conic_op, float* rgb, uint8_t* clamped, int32_t* rect, uint32_t* tiles) {
    CameraDev cam = to_dev(rc);
    for (int i = 0; i < n; ++i) {
        Projected pr;
        Vec3 p{means[3 * i], means[3 * i + 1], means[3 * i + 2]};
        bool ok = project_gaussian(cam, p, Vec3{scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]},
                                   Quat{rots[4 * i], rots[4 * i + 1], rots[4 * i + 2], rots[4 * i + 3]}, pr);
        float c[3] = {0, 0, 0}; uint32_t bits = 0;
        if (ok) {
            const float* sh = shs + (size_t)i * 48;
            sh_to_rgb(cam, p, [&](int k, int ch) { return sh[3 * k + ch]; }, c, bits);
        }
        depth[i] = pr.depth; radii[i] = pr.radius; xy[2 * i] = pr.px; xy[2 * i + 1] = pr.py;
        conic_op[4 * i] = pr.conx; conic_op[4 * i + 1] = pr.cony; conic_op[4 * i + 2] = pr.conz;
        conic_op[4 * i + 3] = ok ? opac[i] : 0.f;
        for (int ch = 0; ch < 3; ++ch) { rgb[3 * i + ch] = c[ch]; clamped[3 * i + ch] = (bits >> ch) & 1; }
        rect[4 * i] = pr.rminx; rect[4 * i + 1] = pr.rminy; rect[4 * i + 2] = pr.rmaxx; rect[4 * i + 3] = pr.rmaxy;
        tiles[i] = pr.tiles;
    }
}

extern "C" void emul_backward(const RefCam* rc, int n, const float* means, const float* scales, const float* rots,
                              const float* shs, const int32_t* radii, const uint8_t* clamped, const float* g_mean2D,
                              const float* g_conic, const float* g_rgb, float* g_means, float* g_scales, float* g_rots,
                              float* g_shs) {
    CameraDev cam = to_dev(rc);
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) { g_means[3 * i + k] = 0; g_scales[3 * i + k] = 0; }
        for (int k = 0; k < 4; ++k) g_rots[4 * i + k] = 0;
        for (int k = 0; k < 48; ++k) g_shs[(size_t)i * 48 + k] = 0;
        if (radii[i] <= 0) continue;
        const float* sh = shs + (size_t)i * 48;
        float* gs = g_shs + (size_t)i * 48;
        uint32_t bits = clamped[3 * i] | (clamped[3 * i + 1] << 1) | (clamped[3 * i + 2] << 2);
        GaussGrad gg;
        gaussian_backward(cam, Vec3{means[3 * i], means[3 * i + 1], means[3 * i + 2]},
                          Vec3{scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]},
                          Quat{rots[4 * i], rots[4 * i + 1], rots[4 * i + 2], rots[4 * i + 3]}, bits, g_mean2D + 2 * i,
                          g_conic + 3 * i, g_rgb + 3 * i, [&](int k, int ch) { return sh[3 * k + ch]; },
                          [&](int k, int ch, float v) { gs[3 * k + ch] = v; }, gg);
        for (int k = 0; k < 3; ++k) { g_means[3 * i + k] = gg.mean[k]; g_scales[3 * i + k] = gg.scale[k]; }
        for (int k = 0; k < 4; ++k) g_rots[4 * i + k] = gg.rot[k];
    }
}
