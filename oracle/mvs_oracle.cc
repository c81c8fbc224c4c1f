/* TEST INFRASTRUCTURE - NOT PRODUCT CODE.  See mvs_oracle.h for role and pinning.
 *
 * Scalar CPU restatement of libs/dmrecon (reference citations are relative to
 * /root/reference).  Written as flat structs + free functions; arithmetic is
 * evaluated in the reference's source order (std::inner_product / accumulate
 * left to right from 0, libs/math/vector.h:441-444,542-545, matrix.h:473-492)
 * and compiled with -ffp-contract=off so the result does not depend on the
 * host CPU.  The reference itself is built with -funsafe-math-optimizations,
 * so agreement with oracle/_ref is statistical for floats (tolerances in
 * tests/test_oracle_vs_reference.py) and exact for integer results.
 */
#include "mvs_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <queue>
#include <unordered_set>
#include <vector>

namespace {

const int NS = 25;          /* filterWidth^2 samples               */
const int CENTER = 12;      /* nrSamples/2, patch_sampler.cc:73,96 */
const int MAXV = 512;       /* capacity of the per-patch view tables */
const float PI_F = 3.141592653589793f; /* defines.h:28 */

/* sRGB code value -> linear, same formula as the table comment at mvs_tools.cc:21-29.
 * The table entries are the float roundings of this double expression. */
float g_srgb2lin[256];
bool g_lut_ready = false;
void init_lut()
{
    if (g_lut_ready) return;
    for (int i = 0; i < 256; ++i) {
        double x = i / 255.0;
        double v = (i <= 0.04045 * 255.0) ? x / 12.92 : std::pow((x + 0.055) / 1.055, 2.4);
        g_srgb2lin[i] = (float)v;
    }
    g_lut_ready = true;
}

struct Level {
    int w = 0, h = 0;
    std::vector<uint8_t> img;      /* interleaved RGB, (y*w+x)*3+c (image.h:314-348) */
    float proj[9], invproj[9];     /* image_pyramid.h:51-59 */
};

struct View {
    bool valid = false;
    int w = 0, h = 0;
    float flen = 0, paspect = 1, pp[2] = {0.5f, 0.5f}, rot[9], trans[3];
    float campos[3];               /* camera.cc:34-39 */
    float w2c[12];                 /* rows 0..2 of camera.cc:61-67 */
    std::vector<Level> lv;
};

struct Feature {
    float pos[3];
    std::vector<int> refs;
};

} // namespace

struct mvs_oracle_scene {
    std::vector<View> views;
    std::vector<Feature> feats;
};

namespace {

/* ---------- libs/math restatements ---------- */
inline float dot3(const float* a, const float* b) { return ((0.0f + a[0] * b[0]) + a[1] * b[1]) + a[2] * b[2]; }
inline float sqn3(const float* a) { return ((0.0f + a[0] * a[0]) + a[1] * a[1]) + a[2] * a[2]; }
inline void normalize3(float* a) { float n = std::sqrt(sqn3(a)); a[0] /= n; a[1] /= n; a[2] /= n; }
inline void cross3(const float* a, const float* b, float* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
inline float round_mve(float x) { return x > 0.0f ? std::floor(x + 0.5f) : std::ceil(x - 0.5f); } /* functions.h:70-73 */
inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }     /* functions.h:204-207 */

/* Matrix4f::mult(Vec3f, 1) (matrix.h:485-492) */
inline void world_to_cam(const View& v, const float* p, float* o)
{
    for (int i = 0; i < 3; ++i)
        o[i] = dot3(v.w2c + 4 * i, p) + 1.0f * v.w2c[4 * i + 3];
}
inline void mat3_mul(const float* m, const float* x, float* o)
{
    for (int i = 0; i < 3; ++i) o[i] = dot3(m + 3 * i, x);
}

/* camera.cc:125-144,180-200 */
void fill_calibration(const View& v, float ppx, float ppy, float width, float height, float* K, float* Ki)
{
    float dim_aspect = width / height;
    float image_aspect = dim_aspect * v.paspect;
    float ax, ay;
    if (image_aspect < 1.0f) { ax = v.flen * height / v.paspect; ay = v.flen * height; }
    else                     { ax = v.flen * width;              ay = v.flen * width * v.paspect; }
    K[0] = ax;   K[1] = 0.f; K[2] = width * ppx;
    K[3] = 0.f;  K[4] = ay;  K[5] = height * ppy;
    K[6] = 0.f;  K[7] = 0.f; K[8] = 1.f;
    Ki[0] = 1.0f / ax; Ki[1] = 0.f;       Ki[2] = -width * ppx / ax;
    Ki[3] = 0.f;       Ki[4] = 1.0f / ay; Ki[5] = -height * ppy / ay;
    Ki[6] = 0.f;       Ki[7] = 0.f;       Ki[8] = 1.f;
}

/* image_tools.h:617-694 with Accum<uint8> (accum.h:117-170) */
void half_size_gaussian(const Level& in, Level& out)
{
    const int64_t iw = in.w, ih = in.h, ic = 3;
    const int64_t ow = (iw + 1) >> 1, oh = (ih + 1) >> 1;
    out.w = (int)ow; out.h = (int)oh;
    out.img.resize((size_t)(ow * oh * ic));
    const float w1 = std::exp(-0.5f / (2.0f * 1.0f));
    const float w2 = std::exp(-2.5f / (2.0f * 1.0f));
    const float w3 = std::exp(-4.5f / (2.0f * 1.0f));
    const float wt[4][4] = {{w3, w2, w2, w3}, {w2, w1, w1, w2}, {w2, w1, w1, w2}, {w3, w2, w2, w3}};
    int64_t outpos = 0;
    const int64_t rowstride = iw * ic;
    for (int64_t y = 0; y < oh; ++y) {
        const int64_t y2 = y * 2;
        const uint8_t* row[4];
        row[0] = &in.img[(size_t)(std::max<int64_t>(0, y2 - 1) * rowstride)];
        row[1] = &in.img[(size_t)(y2 * rowstride)];
        row[2] = &in.img[(size_t)(std::min(ih - 1, y2 + 1) * rowstride)];
        row[3] = &in.img[(size_t)(std::min(ih - 1, y2 + 2) * rowstride)];
        for (int64_t x = 0; x < ow; ++x) {
            const int64_t x2 = x * 2;
            int64_t xi[4];
            xi[0] = std::max<int64_t>(0, x2 - 1) * ic;
            xi[1] = x2 * ic;
            xi[2] = std::min(iw - 1, x2 + 1) * ic;
            xi[3] = std::min(iw - 1, x2 + 2) * ic;
            for (int64_t c = 0; c < ic; ++c) {
                float v = 0.0f, w = 0.0f;
                for (int r = 0; r < 4; ++r)
                    for (int k = 0; k < 4; ++k) {
                        v += (float)row[r][xi[k] + c] * wt[r][k];
                        w += wt[r][k];
                    }
                out.img[(size_t)outpos++] = (uint8_t)round_mve(v / w);
            }
        }
    }
}

/* image_pyramid.cc:19-95 */
void build_pyramid(View& v, const uint8_t* rgb)
{
    const int MIN_IMAGE_DIM = 30;
    v.lv.clear();
    float ppx = v.pp[0], ppy = v.pp[1];
    int cw = v.w, ch = v.h;
    v.lv.emplace_back();
    v.lv.back().w = cw; v.lv.back().h = ch;
    fill_calibration(v, ppx, ppy, (float)cw, (float)ch, v.lv.back().proj, v.lv.back().invproj);
    while (std::min(cw, ch) >= MIN_IMAGE_DIM) {
        if (cw % 2 == 1) ppx = ppx * float(cw) / float(cw + 1);
        if (ch % 2 == 1) ppy = ppy * float(ch) / float(ch + 1);
        cw = (cw + 1) / 2;
        ch = (ch + 1) / 2;
        v.lv.emplace_back();
        v.lv.back().w = cw; v.lv.back().h = ch;
        fill_calibration(v, ppx, ppy, (float)cw, (float)ch, v.lv.back().proj, v.lv.back().invproj);
    }
    v.lv[0].img.assign(rgb, rgb + (size_t)v.w * v.h * 3);
    for (size_t i = 1; i < v.lv.size(); ++i) {
        Level tmp;
        half_size_gaussian(v.lv[i - 1], tmp);
        v.lv[i].img.swap(tmp.img);   /* dims were already set from the level table */
    }
}

/* single_view.h:154-164 */
inline float foot_print(const View& v, const float* p) { float c[3]; world_to_cam(v, p, c); return c[2] * v.lv[0].invproj[0]; }
inline float foot_print_level(const View& v, int level, const float* p) { float c[3]; world_to_cam(v, p, c); return c[2] * v.lv[level].invproj[0]; }

/* single_view.h:176-195 */
inline void world_to_screen(const View& v, int level, const float* p, float* xy)
{
    float cp[3], sp[3];
    world_to_cam(v, p, cp);
    mat3_mul(v.lv[level].proj, cp, sp);
    xy[0] = sp[0] / sp[2] - 0.5f;
    xy[1] = sp[1] / sp[2] - 0.5f;
}

/* single_view.cc:99-106 + depthmap.cc:149-156 */
inline void view_ray(const View& v, int level, int x, int y, float* out)
{
    float px[3] = {(float)x + 0.5f, (float)y + 0.5f, 1.0f};
    float ray[3];
    mat3_mul(v.lv[level].invproj, px, ray);
    normalize3(ray);
    ray[0] *= 1.0f; ray[1] *= 1.0f; ray[2] *= 1.0f;
    for (int i = 0; i < 3; ++i)
        out[i] = ((0.0f + v.rot[0 + i] * ray[0]) + v.rot[3 + i] * ray[1]) + v.rot[6 + i] * ray[2];
}

/* single_view.cc:109-121 */
inline bool point_in_frustum(const View& v, const float* wp)
{
    float cp[3], sp[3];
    world_to_cam(v, wp, cp);
    if (cp[2] <= 0.0f) return false;
    mat3_mul(v.lv[0].proj, cp, sp);
    float x = sp[0] / sp[2] - 0.5f;
    float y = sp[1] / sp[2] - 0.5f;
    return x >= 0 && x <= v.lv[0].w - 1 && y >= 0 && y <= v.lv[0].h - 1;
}

/* mvs_tools.h:46-69 */
inline float parallax(const float* p, const View& a, const View& b)
{
    float d1[3] = {p[0] - a.campos[0], p[1] - a.campos[1], p[2] - a.campos[2]};
    float d2[3] = {p[0] - b.campos[0], p[1] - b.campos[1], p[2] - b.campos[2]};
    normalize3(d1); normalize3(d2);
    float dp = std::max(std::min(dot3(d1, d2), 1.f), -1.f);
    return std::acos(dp) * 180.f / PI_F;
}
inline float parallax_to_weight(float p)
{
    if (p < 0.f || p > 180.f) return 0.f;
    float sigma = (p <= 20.f) ? 5.f : 15.f;
    float mean = 20.f;
    float d = p - mean;
    return std::exp(-(d * d) / (2 * sigma * sigma));
}

/* ---------- PatchSampler (patch_sampler.cc) ---------- */
struct Sampler {
    const mvs_oracle_scene* sc;
    const mvs_oracle_settings* st;
    int ref;
    const View* rv;
    int nviews;
    int mx, my;
    float depth, dzI, dzJ;
    float rays[NS][3], pts[NS][3], mcol[NS][3];
    float masterMean, meanX[3], sqrDevX;
    uint8_t success[MAXV];
    uint8_t cache_filled[MAXV];       /* !neighColorSamples[v].empty() */
    float ncol[MAXV][NS][3];
    mvs_oracle_stats* stats;
};

/* patch_sampler.cc:274-295 */
void compute_patch_points(Sampler& s)
{
    int k = 0;
    for (int j = -2; j <= 2; ++j)
        for (int i = -2; i <= 2; ++i) {
            float tmp = s.depth + (float)i * s.dzI + (float)j * s.dzJ;
            if (tmp <= 0.f) { s.success[s.ref] = 0; return; }
            for (int c = 0; c < 3; ++c) s.pts[k][c] = s.rv->campos[c] + tmp * s.rays[k][c];
            ++k;
        }
}

/* patch_sampler.cc:298-345 */
void compute_master_samples(Sampler& s)
{
    const Level& L = s.rv->lv[s.st->scale];
    int k = 0;
    for (int j = s.my - 2; j <= s.my + 2; ++j)
        for (int i = s.mx - 2; i <= s.mx + 2; ++i) {
            size_t idx = ((size_t)j * L.w + i) * 3;
            for (int c = 0; c < 3; ++c) s.mcol[k][c] = g_srgb2lin[L.img[idx + c]];
            ++k;
        }
    s.masterMean = 0.f;
    for (int q = 0; q < NS; ++q)
        for (int c = 0; c < 3; ++c) s.masterMean += s.mcol[q][c];
    s.masterMean /= 3.f * NS;
    if (s.masterMean < 0.01f || s.masterMean > 0.99f) { s.success[s.ref] = 0; return; }
    s.meanX[0] = s.meanX[1] = s.meanX[2] = 0.f;
    for (int q = 0; q < NS; ++q)
        for (int c = 0; c < 3; ++c) { s.mcol[q][c] /= s.masterMean; s.meanX[c] += s.mcol[q][c]; }
    for (int c = 0; c < 3; ++c) s.meanX[c] /= NS;
    s.sqrDevX = 0.f;
    for (int q = 0; q < NS; ++q) {
        float d[3] = {s.mcol[q][0] - s.meanX[0], s.mcol[q][1] - s.meanX[1], s.mcol[q][2] - s.meanX[2]};
        s.sqrDevX += sqn3(d);
    }
}

/* patch_sampler.cc:19-62 */
void sampler_init(Sampler& s, const mvs_oracle_scene* sc, const mvs_oracle_settings* st, int ref,
                  int x, int y, float depth, float dzI, float dzJ, mvs_oracle_stats* stats)
{
    s.sc = sc; s.st = st; s.ref = ref; s.rv = &sc->views[ref]; s.nviews = (int)sc->views.size();
    s.mx = x; s.my = y; s.depth = depth; s.dzI = dzI; s.dzJ = dzJ; s.stats = stats;
    s.masterMean = 0.f;
    std::memset(s.success, 0, s.nviews);
    std::memset(s.cache_filled, 0, s.nviews);
    const Level& L = s.rv->lv[st->scale];
    if (x - 2 < 0 || y - 2 < 0 || x + 2 > L.w - 1 || y + 2 > L.h - 1) return;
    int k = 0;
    for (int j = y - 2; j <= y + 2; ++j)
        for (int i = x - 2; i <= x + 2; ++i)
            view_ray(*s.rv, st->scale, i, j, s.rays[k++]);
    s.success[ref] = 1;
    compute_master_samples(s);
    compute_patch_points(s);
}

/* patch_sampler.cc:259-271 */
void sampler_update(Sampler& s, float d, float dzI, float dzJ)
{
    if (s.stats) s.stats->n_update++;
    std::memset(s.success, 0, s.nviews);
    s.depth = d; s.dzI = dzI; s.dzJ = dzJ;
    s.success[s.ref] = 1;
    compute_patch_points(s);
    std::memset(s.cache_filled, 0, s.nviews);
}

/* mip level choice shared by patch_sampler.cc:76-91 and :359-375. Returns -1 when nfp <= 0. */
int pick_level(Sampler& s, int v)
{
    const float* p0 = s.pts[CENTER];
    float mfp = foot_print_level(*s.rv, s.st->scale, p0);
    float nfp = foot_print(s.sc->views[v], p0);
    if (mfp <= 0.f) return -2;   /* reference throws std::out_of_range (patch_sampler.cc:78-82) */
    if (nfp <= 0.f) return -1;
    float ratio = nfp / mfp;
    int mm = 0;
    while (ratio < 0.5f) { ++mm; ratio *= 2.f; }
    int maxl = (int)s.sc->views[v].lv.size() - 1;   /* clampLevel, minLevel = 0 for neighbours */
    if (mm < 0) mm = 0;
    if (mm > maxl) mm = maxl;
    return mm;
}

struct NegativeFootprint {};

/* patch_sampler.cc:65-133 + mvs_tools.cc:98-145 */
void fast_col_and_deriv(Sampler& s, int v, float col[NS][3], float der[NS][3])
{
    if (s.stats) s.stats->n_pse_deriv++;
    s.success[v] = 0;
    const View& nv = s.sc->views[v];
    int mm = pick_level(s, v);
    if (mm == -2) throw NegativeFootprint();
    if (mm < 0) return;
    float p1[3] = {s.pts[CENTER][0] + s.rays[CENTER][0], s.pts[CENTER][1] + s.rays[CENTER][1], s.pts[CENTER][2] + s.rays[CENTER][2]};
    float a[2], b[2];
    world_to_screen(nv, mm, p1, a);
    world_to_screen(nv, mm, s.pts[CENTER], b);
    float dd[2] = {a[0] - b[0], a[1] - b[1]};
    float d = std::sqrt((0.0f + dd[0] * dd[0]) + dd[1] * dd[1]);
    if (!(d > 0.f)) return;
    float step = 1.f / d;
    const Level& L = nv.lv[mm];
    const int w = L.w, h = L.h;
    float pos[NS][2], grad[NS][2];
    for (int i = 0; i < NS; ++i) {
        float q1[3] = {s.pts[i][0] + s.rays[i][0] * step, s.pts[i][1] + s.rays[i][1] * step, s.pts[i][2] + s.rays[i][2] * step};
        world_to_screen(nv, mm, s.pts[i], pos[i]);
        if (!(pos[i][0] > 0 && pos[i][0] < w - 1 && pos[i][1] > 0 && pos[i][1] < h - 1)) return;
        float t[2];
        world_to_screen(nv, mm, q1, t);
        grad[i][0] = t[0] - pos[i][0];
        grad[i][1] = t[1] - pos[i][1];
    }
    for (int i = 0; i < NS; ++i) {
        const int left = (int)std::floor(pos[i][0]);
        const int top = (int)std::floor(pos[i][1]);
        const float x = pos[i][0] - left;
        const float y = pos[i][1] - top;
        size_t p0 = ((size_t)top * w + left) * 3;
        size_t p1i = ((size_t)(top + 1) * w + left) * 3;
        const uint8_t* im = L.img.data();
        float x0 = (1.f - x) * g_srgb2lin[im[p0]]     + x * g_srgb2lin[im[p0 + 3]];
        float x1 = (1.f - x) * g_srgb2lin[im[p0 + 1]] + x * g_srgb2lin[im[p0 + 4]];
        float x2 = (1.f - x) * g_srgb2lin[im[p0 + 2]] + x * g_srgb2lin[im[p0 + 5]];
        float x3 = (1.f - x) * g_srgb2lin[im[p1i]]     + x * g_srgb2lin[im[p1i + 3]];
        float x4 = (1.f - x) * g_srgb2lin[im[p1i + 1]] + x * g_srgb2lin[im[p1i + 4]];
        float x5 = (1.f - x) * g_srgb2lin[im[p1i + 2]] + x * g_srgb2lin[im[p1i + 5]];
        col[i][0] = (1.f - y) * x0 + y * x3;
        col[i][1] = (1.f - y) * x1 + y * x4;
        col[i][2] = (1.f - y) * x2 + y * x5;
        float u = grad[i][0], vv = grad[i][1];
        for (int c = 0; c < 3; ++c) {
            float A = g_srgb2lin[im[p0 + c]], B = g_srgb2lin[im[p0 + 3 + c]];
            float C = g_srgb2lin[im[p1i + c]], D = g_srgb2lin[im[p1i + 3 + c]];
            der[i][c] = u * (B - A) + vv * (C - A) + (vv * x + u * y) * (A - B - C + D);
        }
    }
    for (int i = 0; i < NS; ++i)
        for (int c = 0; c < 3; ++c) der[i][c] /= step;
    s.success[v] = 1;
}

/* patch_sampler.cc:348-393 + mvs_tools.cc:169-199 */
void compute_neigh_color_samples(Sampler& s, int v)
{
    if (s.stats) s.stats->n_pse_color++;
    s.success[v] = 0;
    const View& nv = s.sc->views[v];
    int mm = pick_level(s, v);
    if (mm == -2) throw NegativeFootprint();
    if (mm < 0) return;                       /* cache stays empty */
    const Level& L = nv.lv[mm];
    const int w = L.w, h = L.h;
    s.cache_filled[v] = 1;                    /* color.resize(nrSamples) happened */
    for (int i = 0; i < NS; ++i) for (int c = 0; c < 3; ++c) s.ncol[v][i][c] = 0.f;
    float pos[NS][2];
    for (int i = 0; i < NS; ++i) {
        world_to_screen(nv, mm, s.pts[i], pos[i]);
        if (!(pos[i][0] > 0 && pos[i][0] < w - 1 && pos[i][1] > 0 && pos[i][1] < h - 1)) return;
    }
    const uint8_t* im = L.img.data();
    for (int i = 0; i < NS; ++i) {
        const int ii = (int)std::floor(pos[i][0]);
        const int jj = (int)std::floor(pos[i][1]);
        const float u = pos[i][0] - ii;
        const float vv = pos[i][1] - jj;
        size_t p0 = ((size_t)jj * w + ii) * 3;
        size_t p1 = ((size_t)(jj + 1) * w + ii) * 3;
        float x0 = (1.f - u) * g_srgb2lin[im[p0]]     + u * g_srgb2lin[im[p0 + 3]];
        float x1 = (1.f - u) * g_srgb2lin[im[p0 + 1]] + u * g_srgb2lin[im[p0 + 4]];
        float x2 = (1.f - u) * g_srgb2lin[im[p0 + 2]] + u * g_srgb2lin[im[p0 + 5]];
        float x3 = (1.f - u) * g_srgb2lin[im[p1]]     + u * g_srgb2lin[im[p1 + 3]];
        float x4 = (1.f - u) * g_srgb2lin[im[p1 + 1]] + u * g_srgb2lin[im[p1 + 4]];
        float x5 = (1.f - u) * g_srgb2lin[im[p1 + 2]] + u * g_srgb2lin[im[p1 + 5]];
        s.ncol[v][i][0] = (1.f - vv) * x0 + vv * x3;
        s.ncol[v][i][1] = (1.f - vv) * x1 + vv * x4;
        s.ncol[v][i][2] = (1.f - vv) * x2 + vv * x5;
    }
    s.success[v] = 1;
}

/* patch_sampler.cc:136-163 */
float get_fast_ncc(Sampler& s, int v)
{
    if (s.stats) s.stats->n_ncc++;
    if (!s.cache_filled[v]) compute_neigh_color_samples(s, v);
    if (!s.success[v]) return -1.f;
    float meanY[3] = {0.f, 0.f, 0.f};
    for (int i = 0; i < NS; ++i) for (int c = 0; c < 3; ++c) meanY[c] += s.ncol[v][i][c];
    for (int c = 0; c < 3; ++c) meanY[c] /= (float)NS;
    float sqrDevY = 0.f, devXY = 0.f;
    for (int i = 0; i < NS; ++i) {
        float dy[3] = {s.ncol[v][i][0] - meanY[0], s.ncol[v][i][1] - meanY[1], s.ncol[v][i][2] - meanY[2]};
        float dx[3] = {s.mcol[i][0] - s.meanX[0], s.mcol[i][1] - s.meanX[1], s.mcol[i][2] - s.meanX[2]};
        sqrDevY += sqn3(dy);
        devXY += dot3(dx, dy);
    }
    float tmp = std::sqrt(s.sqrDevX * sqrDevY);
    if (tmp > 0) return devXY / tmp;
    return -1.f;
}

/* patch_sampler.cc:243-256 */
void patch_normal(const Sampler& s, float* n)
{
    const int right = CENTER + 2, left = CENTER - 2, top = 2, bottom = NS - 1 - 2;
    float a[3], b[3];
    for (int c = 0; c < 3; ++c) { a[c] = s.pts[right][c] - s.pts[left][c]; b[c] = s.pts[top][c] - s.pts[bottom][c]; }
    cross3(a, b, n);
    normalize3(n);
}

/* ---------- LocalViewSelection + PatchOptimization ---------- */
struct Patch {
    Sampler smp;
    const mvs_oracle_scene* sc;
    const mvs_oracle_settings* st;
    int ref, nviews;
    float depth, dzI, dzJ;
    int iter;
    bool opti_success, converged;
    /* local view selection state */
    bool lvs_success;
    uint8_t available[MAXV];
    int sel[8]; int nsel;          /* ascending */
    float cscale[MAXV][3];
};

void sel_insert(Patch& p, int v)
{
    int k = p.nsel;
    while (k > 0 && p.sel[k - 1] > v) { p.sel[k] = p.sel[k - 1]; --k; }
    p.sel[k] = v; p.nsel++;
}
void sel_erase(Patch& p, int v)
{
    for (int k = 0; k < p.nsel; ++k)
        if (p.sel[k] == v) { for (int q = k; q + 1 < p.nsel; ++q) p.sel[q] = p.sel[q + 1]; p.nsel--; return; }
}

/* local_view_selection.cc:57-147 */
void lvs_perform(Patch& p)
{
    const mvs_oracle_settings& st = *p.st;
    if ((uint32_t)p.nsel == st.nr_recon_neighbors) { p.lvs_success = true; return; }
    const View& rv = p.sc->views[p.ref];
    const float* pt = p.smp.pts[CENTER];
    float mfp = foot_print_level(rv, st.scale, pt);
    float refDir[3] = {pt[0] - rv.campos[0], pt[1] - rv.campos[1], pt[2] - rv.campos[2]};
    normalize3(refDir);
    static thread_local float viewDir[MAXV][3], epi[MAXV][3], ncc[MAXV];
    for (int i = 0; i < p.nviews; ++i) {
        if (!p.available[i]) continue;
        float t = get_fast_ncc(p.smp, i);
        if (t < st.min_ncc) { p.available[i] = 0; continue; }
        ncc[i] = t;
        const View& v = p.sc->views[i];
        for (int c = 0; c < 3; ++c) viewDir[i][c] = pt[c] - v.campos[c];
        normalize3(viewDir[i]);
        cross3(viewDir[i], refDir, epi[i]);
        normalize3(epi[i]);
    }
    for (int k = 0; k < p.nsel; ++k) {
        int s = p.sel[k];
        const View& v = p.sc->views[s];
        for (int c = 0; c < 3; ++c) viewDir[s][c] = pt[c] - v.campos[c];
        normalize3(viewDir[s]);
        cross3(viewDir[s], refDir, epi[s]);
        normalize3(epi[s]);
    }
    bool foundOne = true;
    while ((uint32_t)p.nsel < st.nr_recon_neighbors && foundOne) {
        foundOne = false;
        int maxView = 0;
        float maxScore = 0.f;
        for (int i = 0; i < p.nviews; ++i) {
            if (!p.available[i]) continue;
            float score = ncc[i];
            float nfp = foot_print(p.sc->views[i], pt);
            if (mfp / nfp < 0.5f) score *= 0.01f;
            float dp = clampf(dot3(refDir, viewDir[i]), -1.f, 1.f);
            float plx = std::acos(dp) * 180.f / PI_F;
            score *= parallax_to_weight(plx);
            for (int k = 0; k < p.nsel; ++k) {
                int s = p.sel[k];
                dp = clampf(dot3(viewDir[s], viewDir[i]), -1.f, 1.f);
                plx = std::acos(dp) * 180.f / PI_F;
                score *= parallax_to_weight(plx);
                dp = dot3(epi[i], epi[s]);
                dp = clampf(dp, -1.f, 1.f);
                float angle = std::fabs(std::acos(dp) * 180.f / PI_F);
                if (angle > 90.f) angle = 180.f - angle;
                angle = std::max(angle, 1.f);
                if (angle < st.min_parallax) score *= angle / st.min_parallax;
            }
            if (score > maxScore) { foundOne = true; maxScore = score; maxView = i; }
        }
        if (foundOne) { sel_insert(p, maxView); p.available[maxView] = 0; }
    }
    if ((uint32_t)p.nsel == st.nr_recon_neighbors) p.lvs_success = true;
}

/* patch_optimization.cc:81-111 */
void compute_color_scale(Patch& p)
{
    if (!p.st->use_color_scale) return;
    for (int k = 0; k < p.nsel; ++k) {
        int id = p.sel[k];
        if (!p.smp.cache_filled[id]) compute_neigh_color_samples(p.smp, id);
        if (!p.smp.success[id]) return;
        for (int c = 0; c < 3; ++c) {
            float ab = 0.f, aa = 0.f;
            for (int i = 0; i < NS; ++i) {
                float n = p.smp.ncol[id][i][c];
                ab += (p.smp.mcol[i][c] - n * p.cscale[id][c]) * n;
                aa += n * n;
            }
            if ((double)std::fabs(aa) > 1e-6) {
                p.cscale[id][c] += ab / aa;
                if ((double)p.cscale[id][c] > 1e3) p.opti_success = false;
            } else
                p.opti_success = false;
        }
    }
}

/* patch_optimization.cc:21-78 + local_view_selection.cc:19-54 */
void patch_init(Patch& p, const mvs_oracle_scene* sc, const mvs_oracle_settings* st, int ref,
                const mvs_oracle_patch_in& in, const int32_t* gids, int ng, mvs_oracle_stats* stats)
{
    p.sc = sc; p.st = st; p.ref = ref; p.nviews = (int)sc->views.size();
    p.depth = in.depth; p.dzI = in.dz_i; p.dzJ = in.dz_j;
    sampler_init(p.smp, sc, st, ref, in.x, in.y, in.depth, in.dz_i, in.dz_j, stats);
    p.iter = 0; p.opti_success = true; p.converged = false;
    /* LocalViewSelection ctor */
    p.lvs_success = false;
    p.nsel = 0;
    for (int k = 0; k < in.n_local; ++k) sel_insert(p, in.local_ids[k]);
    std::memset(p.available, 0, p.nviews);
    if (p.smp.success[ref]) {
        if ((uint32_t)p.nsel == st->nr_recon_neighbors) p.lvs_success = true;
        else if ((uint32_t)p.nsel > st->nr_recon_neighbors) p.nsel = 0;
        for (int k = 0; k < ng; ++k) p.available[gids[k]] = 1;
        for (int k = 0; k < p.nsel; ++k) p.available[p.sel[k]] = 0;
    }
    if (!p.smp.success[ref]) { p.opti_success = false; return; }
    lvs_perform(p);
    if (!p.lvs_success) { p.opti_success = false; return; }
    float mm = p.smp.masterMean;
    for (int i = 0; i < p.nviews; ++i) p.cscale[i][0] = p.cscale[i][1] = p.cscale[i][2] = 1.f / mm;
    compute_color_scale(p);
}

/* patch_optimization.cc:265-299 */
void optimize_depth_only(Patch& p)
{
    float numerator = 0.f, denom = 0.f;
    float col[NS][3], der[NS][3];
    for (int k = 0; k < p.nsel; ++k) {
        int id = p.sel[k];
        fast_col_and_deriv(p.smp, id, col, der);
        if (!p.smp.success[id]) { p.opti_success = false; return; }
        const float* cs = p.cscale[id];
        for (int i = 0; i < NS; ++i) {
            float cd[3] = {cs[0] * der[i][0], cs[1] * der[i][1], cs[2] * der[i][2]};
            float rs[3] = {p.smp.mcol[i][0] - cs[0] * col[i][0], p.smp.mcol[i][1] - cs[1] * col[i][1], p.smp.mcol[i][2] - cs[2] * col[i][2]};
            numerator += 1.f * dot3(cd, rs);
            denom += 1.f * sqn3(cd);
        }
    }
    if (denom > 0) {
        p.depth += numerator / denom;
        sampler_update(p.smp, p.depth, p.dzI, p.dzJ);
        p.opti_success = p.smp.success[p.ref] ? true : false;
    }
}

/* patch_optimization.cc:302-364, matrix_tools.h:392-398,460-475 */
void optimize_depth_and_normal(Patch& p)
{
    if (!p.lvs_success) return;
    double ATA[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double ATb[3] = {0, 0, 0};
    float col[NS][3], der[NS][3];
    for (int k = 0; k < p.nsel; ++k) {
        int id = p.sel[k];
        fast_col_and_deriv(p.smp, id, col, der);
        if (!p.smp.success[id]) { p.opti_success = false; return; }
        const float* cs = p.cscale[id];
        for (int i = 0; i < NS; ++i) {
            const int ii = (i % 5) - 2, jj = (i / 5) - 2;
            for (int c = 0; c < 3; ++c) {
                float a0 = 1.f * cs[c] * der[i][c];
                float a1 = 1.f * ii * cs[c] * der[i][c];
                float a2 = 1.f * jj * cs[c] * der[i][c];
                float b = 1.f * (p.smp.mcol[i][c] - cs[c] * col[i][c]);
                ATA[0] += a0 * a0;
                ATA[1] += a0 * a1;
                ATA[2] += a0 * a2;
                ATA[4] += a1 * a1;
                ATA[5] += a1 * a2;
                ATA[8] += a2 * a2;
                ATb[0] += a0 * b;
                ATb[1] += a1 * b;
                ATb[2] += a2 * b;
            }
        }
    }
    ATA[3] = ATA[1]; ATA[6] = ATA[2]; ATA[7] = ATA[5];
    const double* m = ATA;
    double det = m[0] * m[4] * m[8] + m[1] * m[5] * m[6] + m[2] * m[3] * m[7]
               - m[2] * m[4] * m[6] - m[1] * m[3] * m[8] - m[0] * m[5] * m[7];
    if (det == 0.f) { p.opti_success = false; return; }
    double inv[9];
    inv[0] = m[4] * m[8] - m[5] * m[7];
    inv[1] = m[2] * m[7] - m[1] * m[8];
    inv[2] = m[1] * m[5] - m[2] * m[4];
    inv[3] = m[5] * m[6] - m[3] * m[8];
    inv[4] = m[0] * m[8] - m[2] * m[6];
    inv[5] = m[2] * m[3] - m[0] * m[5];
    inv[6] = m[3] * m[7] - m[4] * m[6];
    inv[7] = m[1] * m[6] - m[0] * m[7];
    inv[8] = m[0] * m[4] - m[1] * m[3];
    for (int i = 0; i < 9; ++i) inv[i] /= det;
    float X[3];
    for (int i = 0; i < 3; ++i)
        X[i] = (float)(((0.0 + inv[3 * i] * ATb[0]) + inv[3 * i + 1] * ATb[1]) + inv[3 * i + 2] * ATb[2]);
    p.dzI += X[1];
    p.dzJ += X[2];
    p.depth += X[0];
    sampler_update(p.smp, p.depth, p.dzI, p.dzJ);
    p.opti_success = p.smp.success[p.ref] ? true : false;
}

/* patch_optimization.cc:170-242 */
void do_auto_optimization(Patch& p)
{
    const mvs_oracle_settings& st = *p.st;
    if (!p.lvs_success || !p.opti_success) return;
    while (p.iter < 4 && p.opti_success) { optimize_depth_only(p); ++p.iter; }
    bool viewRemoved = false;
    while ((uint32_t)p.iter < st.max_iterations && p.lvs_success && p.opti_success) {
        float oldNCC[8];
        int nOld = p.nsel;
        for (int k = 0; k < p.nsel; ++k) oldNCC[k] = get_fast_ncc(p.smp, p.sel[k]);
        p.opti_success = false;
        if (p.iter % 5 == 4 || viewRemoved) {
            optimize_depth_and_normal(p);
            compute_color_scale(p);
            viewRemoved = false;
        } else
            optimize_depth_only(p);
        if (!p.opti_success) return;
        bool converged = true;
        int tbr[8]; int ntbr = 0;
        for (int k = 0; k < nOld; ++k) {
            float ncc = get_fast_ncc(p.smp, p.sel[k]);
            if (std::fabs(ncc - oldNCC[k]) > st.min_refine_diff) converged = false;
            if ((ncc < st.accept_ncc) || (p.iter == 14 && std::fabs(ncc - oldNCC[k]) > st.min_refine_diff)) {
                tbr[ntbr++] = p.sel[k];
                viewRemoved = true;
            }
        }
        if (viewRemoved) {
            /* LocalViewSelection::replaceViews (local_view_selection.cc:150-160) */
            for (int k = 0; k < ntbr; ++k) { p.available[tbr[k]] = 0; sel_erase(p, tbr[k]); }
            p.lvs_success = false;
            lvs_perform(p);
            if (!p.lvs_success) return;
            compute_color_scale(p);
        } else if (!p.opti_success) {
            return;
        } else if (converged) {
            p.converged = true;
            return;
        }
        ++p.iter;
    }
}

/* patch_optimization.cc:114-142 */
float compute_confidence(Patch& p)
{
    if (!p.converged) return 0.f;
    float meanNCC = 0.f;
    for (int k = 0; k < p.nsel; ++k) meanNCC += get_fast_ncc(p.smp, p.sel[k]);
    meanNCC /= p.nsel;
    float score = (meanNCC - p.st->accept_ncc) / (1.f - p.st->accept_ncc);
    float viewDir[3], n[3];
    view_ray(p.sc->views[p.ref], p.st->scale, p.smp.mx, p.smp.my, viewDir);
    patch_normal(p.smp, n);
    float dotP = -dot3(n, viewDir);
    if (dotP < 0.2f) return 0.f;
    return score;
}

void run_patch(const mvs_oracle_scene* sc, const mvs_oracle_settings* st, int ref,
               const int32_t* gids, int ng, const mvs_oracle_patch_in& in, mvs_oracle_patch_out& out,
               mvs_oracle_stats* stats)
{
    static thread_local Patch* pp = nullptr;
    if (!pp) pp = new Patch();
    Patch& p = *pp;
    if (stats) stats->n_opt++;
    std::memset(&out, 0, sizeof(out));
    for (int k = 0; k < 4; ++k) out.local_ids[k] = -1;
    try {
        patch_init(p, sc, st, ref, in, gids, ng, stats);
        do_auto_optimization(p);
        out.conf = compute_confidence(p);
    } catch (NegativeFootprint&) {
        /* The reference lets std::out_of_range escape DMRecon::start; callers see conf 0. */
        out.conf = 0.f; p.converged = false; p.opti_success = false;
        out.iterations = -1;
        return;
    }
    out.depth = p.depth; out.dz_i = p.dzI; out.dz_j = p.dzJ;
    if (p.smp.success[ref] || p.smp.masterMean != 0.f) {
        /* getNormal() is only read by callers when conf > 0; compute it whenever points exist */
    }
    if (out.conf > 0.f || p.converged) patch_normal(p.smp, out.normal);
    out.n_local = p.nsel;
    for (int k = 0; k < p.nsel && k < 4; ++k) out.local_ids[k] = p.sel[k];
    out.iterations = p.iter;
    out.converged = p.converged ? 1 : 0;
    out.opti_success = p.opti_success ? 1 : 0;
}

/* ---------- DMRecon (dmrecon.cc) ---------- */
bool feature_contains(const Feature& f, int id)
{
    for (size_t i = 0; i < f.refs.size(); ++i) if (f.refs[i] == id) return true;
    return false;
}
bool in_aabb(const float* p, const mvs_oracle_settings& st)
{
    for (int i = 0; i < 3; ++i) if (p[i] < st.aabb_min[i] || p[i] > st.aabb_max[i]) return false;
    return true;
}

/* dmrecon.cc:179-208 + global_view_selection.cc:17-101 */
std::vector<int> global_view_selection(const mvs_oracle_scene* sc, const mvs_oracle_settings& st, int ref)
{
    const int nv = (int)sc->views.size();
    const View& rv = sc->views[ref];
    std::vector<std::vector<int>> featInd(nv);
    for (size_t i = 0; i < sc->feats.size(); ++i) {
        const Feature& f = sc->feats[i];
        if (!feature_contains(f, ref)) continue;
        if (!point_in_frustum(rv, f.pos)) continue;
        if (!in_aabb(f.pos, st)) continue;
        for (size_t j = 0; j < f.refs.size(); ++j) {
            int vid = f.refs[j];
            if (vid < 0 || vid >= nv || !sc->views[vid].valid) continue;
            if (point_in_frustum(sc->views[vid], f.pos)) featInd[vid].push_back((int)i);
        }
    }
    std::vector<std::unordered_set<int>> sees(nv);
    for (int v = 0; v < nv; ++v) sees[v].insert(featInd[v].begin(), featInd[v].end());
    std::vector<uint8_t> available(nv, 1);
    available[ref] = 0;
    for (int v = 0; v < nv; ++v) if (!sc->views[v].valid) available[v] = 0;
    std::vector<int> selected;   /* kept ascending like std::set */
    bool foundOne = true;
    while (foundOne && selected.size() < st.global_vs_max) {
        float maxBenefit = 0.f;
        int maxView = 0;
        foundOne = false;
        for (int i = 0; i < nv; ++i) {
            if (!available[i]) continue;
            const View& tv = sc->views[i];
            float benefit = 0;
            for (size_t k = 0; k < featInd[i].size(); ++k) {
                float score = 1.f;
                const float* fp = sc->feats[featInd[i][k]].pos;
                float plx = parallax(fp, rv, tv);
                if (plx < st.min_parallax) { float q = plx / 10.f; score *= q * q; }
                float mfp = foot_print_level(rv, st.scale, fp);
                float nfp = foot_print(tv, fp);
                float ratio = mfp / nfp;
                if (ratio > 2.) ratio = (float)(2. / ratio);
                else if (ratio > 1.) ratio = 1.;
                score *= ratio;
                for (size_t s = 0; s < selected.size(); ++s) {
                    if (!sees[selected[s]].count(featInd[i][k])) continue;
                    plx = parallax(fp, sc->views[selected[s]], tv);
                    if (plx < st.min_parallax) { float q = plx / 10.f; score *= q * q; }
                }
                benefit += score;
            }
            if (benefit > maxBenefit) { maxBenefit = benefit; maxView = i; foundOne = true; }
        }
        if (foundOne) {
            selected.insert(std::upper_bound(selected.begin(), selected.end(), maxView), maxView);
            available[maxView] = 0;
        }
    }
    return selected;
}

struct QEntry {
    int x, y;
    float conf, depth, dzI, dzJ;
    int nloc; int loc[4];
    int64_t serial;          /* bookkeeping only (speculative-batch simulation); not part of the ordering */
    bool operator<(const QEntry& o) const { return conf < o.conf; }   /* dmrecon.h:72-76 */
};

} // namespace

extern "C" {

mvs_oracle_scene* mvs_oracle_create(int n_views)
{
    init_lut();
    if (n_views <= 0 || n_views > MAXV) return nullptr;
    mvs_oracle_scene* s = new mvs_oracle_scene();
    s->views.resize(n_views);
    return s;
}

void mvs_oracle_destroy(mvs_oracle_scene* s) { delete s; }

int mvs_oracle_set_view(mvs_oracle_scene* s, int id, const uint8_t* rgb, int w, int h,
                        float flen, float paspect, const float pp[2], const float rot[9], const float trans[3])
{
    if (!s || id < 0 || id >= (int)s->views.size() || w < 2 || h < 2) return -1;
    View& v = s->views[id];
    v.valid = true; v.w = w; v.h = h; v.flen = flen; v.paspect = paspect;
    v.pp[0] = pp[0]; v.pp[1] = pp[1];
    std::memcpy(v.rot, rot, sizeof(v.rot));
    std::memcpy(v.trans, trans, sizeof(v.trans));
    v.campos[0] = -rot[0] * trans[0] - rot[3] * trans[1] - rot[6] * trans[2];
    v.campos[1] = -rot[1] * trans[0] - rot[4] * trans[1] - rot[7] * trans[2];
    v.campos[2] = -rot[2] * trans[0] - rot[5] * trans[1] - rot[8] * trans[2];
    for (int r = 0; r < 3; ++r) {
        v.w2c[4 * r + 0] = rot[3 * r]; v.w2c[4 * r + 1] = rot[3 * r + 1]; v.w2c[4 * r + 2] = rot[3 * r + 2];
        v.w2c[4 * r + 3] = trans[r];
    }
    build_pyramid(v, rgb);
    return 0;
}

int mvs_oracle_set_features(mvs_oracle_scene* s, int n, const float* pos, const int32_t* off, const int32_t* ids)
{
    if (!s) return -1;
    s->feats.resize(n);
    for (int i = 0; i < n; ++i) {
        std::memcpy(s->feats[i].pos, pos + 3 * i, 3 * sizeof(float));
        s->feats[i].refs.assign(ids + off[i], ids + off[i + 1]);
    }
    return 0;
}

int mvs_oracle_num_levels(mvs_oracle_scene* s, int id) { return (int)s->views[id].lv.size(); }

int mvs_oracle_get_level(mvs_oracle_scene* s, int id, int level, int* w, int* h, uint8_t* rgb)
{
    if (!s || id < 0 || id >= (int)s->views.size()) return -1;
    const View& v = s->views[id];
    if (level < 0 || level >= (int)v.lv.size()) return -1;
    *w = v.lv[level].w; *h = v.lv[level].h;
    if (rgb) std::memcpy(rgb, v.lv[level].img.data(), v.lv[level].img.size());
    return 0;
}

int mvs_oracle_get_level_calib(mvs_oracle_scene* s, int id, int level, float proj[9], float invproj[9])
{
    const View& v = s->views[id];
    if (level < 0 || level >= (int)v.lv.size()) return -1;
    std::memcpy(proj, v.lv[level].proj, 36);
    std::memcpy(invproj, v.lv[level].invproj, 36);
    return 0;
}

int mvs_oracle_global_view_selection(mvs_oracle_scene* s, const mvs_oracle_settings* st, int ref, int32_t* out, int cap)
{
    std::vector<int> sel = global_view_selection(s, *st, ref);
    int n = (int)std::min<size_t>(sel.size(), (size_t)cap);
    for (int i = 0; i < n; ++i) out[i] = sel[i];
    return (int)sel.size();
}

int mvs_oracle_optimize_patches(mvs_oracle_scene* s, const mvs_oracle_settings* st, int ref,
                                const int32_t* gids, int ng, const mvs_oracle_patch_in* in, int n,
                                mvs_oracle_patch_out* out, mvs_oracle_stats* stats)
{
    if (!s || st->filter_width != 5) return -1;
    for (int i = 0; i < n; ++i) run_patch(s, st, ref, gids, ng, in[i], out[i], stats);
    return 0;
}

int mvs_oracle_reconstruct(mvs_oracle_scene* s, const mvs_oracle_settings* st, int ref,
                           float* depth, float* conf, float* dz, float* normal, int32_t* view_ids,
                           mvs_oracle_stats* stats,
                           mvs_oracle_patch_in* trace_in, mvs_oracle_patch_out* trace_out,
                           int64_t trace_cap, int64_t* trace_n, double max_seconds)
{
    if (!s || st->filter_width != 5) return -1;
    if (ref < 0 || ref >= (int)s->views.size() || !s->views[ref].valid) return -2;
    const View& rv = s->views[ref];
    if (st->scale < 0 || st->scale >= (int)rv.lv.size()) return -3;
    const int W = rv.lv[st->scale].w, H = rv.lv[st->scale].h;
    const size_t npix = (size_t)W * H;
    std::memset(depth, 0, npix * 4);
    std::memset(conf, 0, npix * 4);
    std::memset(dz, 0, npix * 8);
    std::memset(normal, 0, npix * 12);
    for (size_t i = 0; i < npix * 4; ++i) view_ids[i] = -1;
    mvs_oracle_stats local; std::memset(&local, 0, sizeof(local));
    mvs_oracle_stats& S = stats ? *stats : local;
    std::memset(&S, 0, sizeof(S));
    int64_t ntrace = 0;
    auto t0 = std::chrono::steady_clock::now();

    std::vector<int> gsel = global_view_selection(s, *st, ref);
    if (gsel.empty()) return -4;     /* "Global View Selection failed" (dmrecon.cc:222-223) */
    std::vector<int32_t> gids(gsel.begin(), gsel.end());

    std::priority_queue<QEntry> pq;
    int64_t serial = 0;
    int64_t spec_watermark = 0;   /* entries with serial below this were computed by a speculative batch */

    auto run = [&](const mvs_oracle_patch_in& in, mvs_oracle_patch_out& out) {
        run_patch(s, st, ref, gids.data(), (int)gids.size(), in, out, &S);
        if (ntrace < trace_cap && trace_in && trace_out) { trace_in[ntrace] = in; trace_out[ntrace] = out; }
        ++ntrace;
    };
    auto write_px = [&](int index, const mvs_oracle_patch_out& o) {
        depth[index] = o.depth;
        normal[3 * index] = o.normal[0]; normal[3 * index + 1] = o.normal[1]; normal[3 * index + 2] = o.normal[2];
        dz[2 * index] = o.dz_i; dz[2 * index + 1] = o.dz_j;
        conf[index] = o.conf;
        for (int k = 0; k < 4; ++k) view_ids[4 * index + k] = o.local_ids[k];
    };

    /* processFeatures (dmrecon.cc:244-331) */
    for (size_t i = 0; i < s->feats.size(); ++i) {
        const Feature& f = s->feats[i];
        bool use = feature_contains(f, ref);
        for (size_t k = 0; !use && k < gsel.size(); ++k) if (feature_contains(f, gsel[k])) use = true;
        if (!use) continue;
        if (!point_in_frustum(rv, f.pos)) continue;
        if (!in_aabb(f.pos, *st)) continue;
        S.n_seeds_processed++;
        float pix[2];
        world_to_screen(rv, st->scale, f.pos, pix);
        const int x = (int)round_mve(pix[0]);
        const int y = (int)round_mve(pix[1]);
        float dv[3] = {f.pos[0] - rv.campos[0], f.pos[1] - rv.campos[1], f.pos[2] - rv.campos[2]};
        mvs_oracle_patch_in in; std::memset(&in, 0, sizeof(in));
        in.x = x; in.y = y; in.depth = std::sqrt(sqn3(dv)); in.dz_i = 0.f; in.dz_j = 0.f; in.n_local = 0;
        for (int k = 0; k < 4; ++k) in.local_ids[k] = -1;
        mvs_oracle_patch_out o;
        run(in, o);
        if (o.conf <= 0.0f) continue;
        S.n_seeds_success++;
        const int index = y * W + x;
        if (conf[index] < o.conf) {
            if (conf[index] <= 0) S.n_filled++;
            write_px(index, o);
            QEntry e; e.x = x; e.y = y; e.conf = o.conf; e.depth = o.depth; e.dzI = o.dz_i; e.dzJ = o.dz_j;
            e.nloc = o.n_local; for (int k = 0; k < 4; ++k) e.loc[k] = o.local_ids[k];
            e.serial = serial++;
            pq.push(e);
        }
    }

    /* processQueue (dmrecon.cc:334-434) */
    while (!pq.empty()) {
        if (max_seconds > 0 && (S.n_pops & 255) == 0) {
            double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (el > max_seconds) break;
        }
        QEntry e = pq.top();
        pq.pop();
        S.n_pops++;
        float fx = (float)e.x, fy = (float)e.y;
        int index = (int)(fy * W + fx);       /* float-typed index (dmrecon.cc:368-370) */
        if (conf[index] > e.conf) {
            S.n_stale++;
            if (e.serial < spec_watermark) S.n_spec_wasted++;
            continue;
        }
        /* Simulation of the strict-order GPU schedule (bookkeeping only, the queue is untouched):
         * when the popped entry's result has not been batch-computed yet, one round computes every
         * entry currently queued, i.e. every serial below the current counter. */
        if (e.serial >= spec_watermark) { S.n_spec_rounds++; spec_watermark = serial; }
        mvs_oracle_patch_in in; std::memset(&in, 0, sizeof(in));
        in.x = e.x; in.y = e.y; in.depth = e.depth; in.dz_i = e.dzI; in.dz_j = e.dzJ; in.n_local = e.nloc;
        for (int k = 0; k < 4; ++k) in.local_ids[k] = e.loc[k];
        mvs_oracle_patch_out o;
        run(in, o);
        if (o.conf == 0) continue;
        if (conf[index] <= 0) S.n_filled++;
        if (conf[index] < o.conf) {
            write_px(index, o);
            QEntry c; c.conf = o.conf; c.depth = o.depth; c.dzI = o.dz_i; c.dzJ = o.dz_j;
            c.nloc = o.n_local; for (int k = 0; k < 4; ++k) c.loc[k] = o.local_ids[k];
            const int nx[4] = {e.x - 1, e.x + 1, e.x, e.x};
            const int ny[4] = {e.y, e.y, e.y - 1, e.y + 1};
            for (int d = 0; d < 4; ++d) {
                c.x = nx[d]; c.y = ny[d];
                int ni = c.y * W + c.x;
                if (conf[ni] < o.conf - 0.05f || conf[ni] == 0.f) { c.serial = serial++; pq.push(c); }
            }
        }
    }
    if (trace_n) *trace_n = ntrace;
    return 0;
}


/* Deterministic frontier ("wavefront") schedule of the same region growing, as executed by the GPU
 * implementation (DESIGN.md "Frontier schedule").  Every PatchOptimization is a pure function of its
 * queue entry (the maps are only read by the stale test and the commit, dmrecon.cc:371,388-431), so
 * the entries of one round may be evaluated in any order / in parallel:
 *   round: A drop entries with conf[pixel] > entry.conf (stale test, dmrecon.cc:371)
 *          B eligibility threshold of the round from the confidences of the non-stale entries, in bins of 1/8192
 *            (bin = min(8191, (int)(conf * 8192))): band > 0: bin >= bin(max) - max(1, (int)(band * 8192));
 *            topk > 0: bin >= the largest t with at least topk entries in bins >= t (0 if there are fewer);
 *            both: the larger threshold; neither: every entry is eligible.  Entries below are deferred unchanged.
 *          C per pixel the eligible entry with the largest (conf, then smallest direction code) runs,
 *            the others are carried over to the next round
 *          D results with conf != 0 are committed when conf[pixel] < result.conf (dmrecon.cc:377-398)
 *          E after ALL commits of the round, each committed result pushes its 4 neighbours under the
 *            reference rule conf[nb] < conf - 0.05 || conf[nb] == 0 (dmrecon.cc:400-431)
 * Seeds: all features are optimised, per pixel the most confident (first in feature order on ties)
 * is committed and becomes a round-0 entry (equivalent to dmrecon.cc:296-326 followed by the stale test). */
int mvs_oracle_reconstruct_wavefront(mvs_oracle_scene* s, const mvs_oracle_settings* st, int ref, float band, int topk,
                                     float* depth, float* conf, float* dz, float* normal, int32_t* view_ids,
                                     mvs_oracle_stats* stats)
{
    if (!s || st->filter_width != 5) return -1;
    if (ref < 0 || ref >= (int)s->views.size() || !s->views[ref].valid) return -2;
    const View& rv = s->views[ref];
    if (st->scale < 0 || st->scale >= (int)rv.lv.size()) return -3;
    const int W = rv.lv[st->scale].w, H = rv.lv[st->scale].h;
    const size_t npix = (size_t)W * H;
    std::memset(depth, 0, npix * 4);
    std::memset(conf, 0, npix * 4);
    std::memset(dz, 0, npix * 8);
    std::memset(normal, 0, npix * 12);
    for (size_t i = 0; i < npix * 4; ++i) view_ids[i] = -1;
    mvs_oracle_stats local; std::memset(&local, 0, sizeof(local));
    mvs_oracle_stats& S = stats ? *stats : local;
    std::memset(&S, 0, sizeof(S));

    std::vector<int> gsel = global_view_selection(s, *st, ref);
    if (gsel.empty()) return -4;
    std::vector<int32_t> gids(gsel.begin(), gsel.end());

    struct WEntry { int x, y, dir; float conf, depth, dzI, dzJ; int nloc; int loc[4]; };
    auto write_px = [&](int index, const mvs_oracle_patch_out& o) {
        depth[index] = o.depth;
        normal[3 * index] = o.normal[0]; normal[3 * index + 1] = o.normal[1]; normal[3 * index + 2] = o.normal[2];
        dz[2 * index] = o.dz_i; dz[2 * index + 1] = o.dz_j;
        conf[index] = o.conf;
        for (int k = 0; k < 4; ++k) view_ids[4 * index + k] = o.local_ids[k];
    };
    auto to_entry = [&](int x, int y, int dir, const mvs_oracle_patch_out& o) {
        WEntry e; e.x = x; e.y = y; e.dir = dir; e.conf = o.conf; e.depth = o.depth; e.dzI = o.dz_i; e.dzJ = o.dz_j;
        e.nloc = o.n_local; for (int k = 0; k < 4; ++k) e.loc[k] = o.local_ids[k];
        return e;
    };

    std::vector<WEntry> cur, next;
    /* seeds */
    for (size_t i = 0; i < s->feats.size(); ++i) {
        const Feature& f = s->feats[i];
        bool use = feature_contains(f, ref);
        for (size_t k = 0; !use && k < gsel.size(); ++k) if (feature_contains(f, gsel[k])) use = true;
        if (!use) continue;
        if (!point_in_frustum(rv, f.pos)) continue;
        if (!in_aabb(f.pos, *st)) continue;
        S.n_seeds_processed++;
        float pix[2];
        world_to_screen(rv, st->scale, f.pos, pix);
        const int x = (int)round_mve(pix[0]);
        const int y = (int)round_mve(pix[1]);
        float dv[3] = {f.pos[0] - rv.campos[0], f.pos[1] - rv.campos[1], f.pos[2] - rv.campos[2]};
        mvs_oracle_patch_in in; std::memset(&in, 0, sizeof(in));
        in.x = x; in.y = y; in.depth = std::sqrt(sqn3(dv));
        for (int k = 0; k < 4; ++k) in.local_ids[k] = -1;
        mvs_oracle_patch_out o;
        run_patch(s, st, ref, gids.data(), (int)gids.size(), in, o, &S);
        if (o.conf <= 0.0f) continue;
        S.n_seeds_success++;
        const int index = y * W + x;
        if (conf[index] < o.conf) write_px(index, o);
    }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const int index = y * W + x;
            if (conf[index] > 0.f) {
                mvs_oracle_patch_out o; std::memset(&o, 0, sizeof(o));
                o.conf = conf[index]; o.depth = depth[index]; o.dz_i = dz[2 * index]; o.dz_j = dz[2 * index + 1];
                o.n_local = 0;
                for (int k = 0; k < 4; ++k) { o.local_ids[k] = view_ids[4 * index + k]; if (o.local_ids[k] >= 0) o.n_local++; }
                cur.push_back(to_entry(x, y, 4, o));
            }
        }

    std::vector<int> winner(npix, -1);
    std::vector<mvs_oracle_patch_out> results;
    std::vector<int> run_list;
    auto conf_bin = [](float c) { int b = (int)(c * 8192.0f); return b < 0 ? 0 : (b > 8191 ? 8191 : b); };
    std::vector<unsigned> hist(8192);
    while (!cur.empty()) {
        S.n_spec_rounds++;
        next.clear();
        /* A: stale test; B: band threshold */
        float maxc = -1e30f;
        for (WEntry& e : cur) {
            const int index = e.y * W + e.x;
            S.n_pops++;
            if (conf[index] > e.conf) { e.dir = -1; S.n_stale++; continue; }
            maxc = std::max(maxc, e.conf);
        }
        int thr_bin = 0;
        if (band > 0.f || topk > 0) {
            std::fill(hist.begin(), hist.end(), 0u);
            int top = -1;
            for (const WEntry& e : cur) if (e.dir >= 0) { const int b = conf_bin(e.conf); hist[b]++; top = std::max(top, b); }
            if (band > 0.f && top >= 0) thr_bin = std::max(0, top - std::max(1, (int)(band * 8192.0f)));
            if (topk > 0) {
                unsigned long cum = 0; int t = 0;
                for (int b = 8191; b >= 0; --b) { cum += hist[b]; if (cum >= (unsigned long)topk) { t = b; break; } }
                thr_bin = std::max(thr_bin, t);
            }
        }
        /* C: per-pixel winner */
        run_list.clear();
        for (size_t i = 0; i < cur.size(); ++i) {
            WEntry& e = cur[i];
            if (e.dir < 0) continue;
            if (conf_bin(e.conf) < thr_bin) { next.push_back(e); e.dir = -1; S.n_pops--; continue; }
            const int index = e.y * W + e.x;
            int& w = winner[index];
            if (w < 0) { w = (int)i; continue; }
            const WEntry& o = cur[w];
            if (e.conf > o.conf || (e.conf == o.conf && e.dir < o.dir)) w = (int)i;
        }
        for (size_t i = 0; i < cur.size(); ++i) {
            WEntry& e = cur[i];
            if (e.dir < 0) continue;
            const int index = e.y * W + e.x;
            if (winner[index] == (int)i) run_list.push_back((int)i);
            else { next.push_back(e); S.n_pops--; }
        }
        /* optimise winners */
        results.resize(run_list.size());
        for (size_t k = 0; k < run_list.size(); ++k) {
            const WEntry& e = cur[run_list[k]];
            winner[e.y * W + e.x] = -1;
            mvs_oracle_patch_in in; std::memset(&in, 0, sizeof(in));
            in.x = e.x; in.y = e.y; in.depth = e.depth; in.dz_i = e.dzI; in.dz_j = e.dzJ; in.n_local = e.nloc;
            for (int q = 0; q < 4; ++q) in.local_ids[q] = e.loc[q];
            run_patch(s, st, ref, gids.data(), (int)gids.size(), in, results[k], &S);
        }
        /* D: commit */
        for (size_t k = 0; k < run_list.size(); ++k) {
            const WEntry& e = cur[run_list[k]];
            const int index = e.y * W + e.x;
            mvs_oracle_patch_out& o = results[k];
            if (o.conf == 0) { o.converged = -1; continue; }
            if (conf[index] < o.conf) write_px(index, o); else o.converged = -1;
        }
        /* E: expand */
        for (size_t k = 0; k < run_list.size(); ++k) {
            const WEntry& e = cur[run_list[k]];
            const mvs_oracle_patch_out& o = results[k];
            if (o.converged < 0) continue;
            const int nx[4] = {e.x - 1, e.x + 1, e.x, e.x};
            const int ny[4] = {e.y, e.y, e.y - 1, e.y + 1};
            for (int d = 0; d < 4; ++d) {
                const int ni = ny[d] * W + nx[d];
                if (conf[ni] < o.conf - 0.05f || conf[ni] == 0.f) next.push_back(to_entry(nx[d], ny[d], d, o));
            }
        }
        cur.swap(next);
    }
    for (size_t i = 0; i < npix; ++i) if (conf[i] > 0.f) S.n_filled++;
    return 0;
}

} // extern "C"
