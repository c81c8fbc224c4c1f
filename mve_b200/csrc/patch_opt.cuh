// Device-side patch optimisation: one warp per patch, lane k < 25 owns sample k of the 5x5 patch.
//
// Restates, for the GPU, what one mvs::PatchOptimization does in the reference
// (libs/dmrecon/patch_optimization.cc:21-364 with PatchSampler patch_sampler.cc:19-393,
// LocalViewSelection local_view_selection.cc:19-160 and mvs_tools.cc:98-199); SURVEY.md Appendix A
// is the line-by-line behavioural spec.  Differences in STRUCTURE (not in results):
//   * the colour-only sample set (computeNeighColorSamples) and the colour+derivative sample set
//     (fastColAndDeriv) of one view at one patch state use identical bilinear formulas
//     (mvs_tools.cc:119-128 vs :188-197), so a single fused sample set per (state, view) is drawn and
//     kept in registers until the state changes (PatchSampler::update, patch_sampler.cc:259-271);
//   * sums over the 25 samples are warp-shuffle reductions; the 3x3 normal equations are accumulated
//     per lane in fp64 from fp32 products exactly as patch_optimization.cc:326-343 and reduced once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstddef>

namespace b200mvs {

constexpr int MAX_LEVELS = 12;
constexpr int MAX_GLOBAL = 32;
constexpr int MAX_LOCAL = 4;
constexpr unsigned FULL = 0xffffffffu;
constexpr int NS = 25;
constexpr int CENTER = 12;   // patch_sampler.cc:73,96

struct alignas(16) LevelParams {   // ImagePyramidLevel (image_pyramid.h:28-59): K = [ax 0 cx; 0 ay cy; 0 0 1]
    float ax, ay, cx, cy;
    int w, h;
    int pitch;                // in texels (uchar4)
    int pad;
    const uchar4* img;        // RGBX8, row pitch 16-byte aligned
    unsigned long long pad2;
};
static_assert(sizeof(LevelParams) == 48, "LevelParams layout");

struct alignas(16) ViewParams {    // SingleView (single_view.h:28-133)
    float campos[3];
    float inv_ax0;            // source_level.invproj[0] (single_view.h:154-157)
    float w2c[12];            // rows 0..2 of worldToCam
    float rot[9];
    int nlevels;
    int valid;
    int pad;
    LevelParams lv[MAX_LEVELS];
};
static_assert(offsetof(ViewParams, lv) % 16 == 0, "ViewParams layout");

struct DevSettings {
    float min_ncc, min_parallax, accept_ncc, min_refine_diff;
    unsigned max_iterations, nr_recon_neighbors;
    int scale, use_color_scale;
};

struct JobParams {            // one reference view being reconstructed (DMRecon members, dmrecon.h:50-62)
    int ref_view, W, H, n_global;
    int gview[MAX_GLOBAL];    // global view ids by slot, ascending (neighViews)
    float ki0, ki2, ki4, ki5; // target_level.invproj entries [0],[2],[4],[5]
    const uchar4* ref_img;    // level `scale` of the reference view
    int ref_pitch;
    int pad;
    float* depth; float* conf; float* dz; float* normal;
    unsigned* slots;          // 4 x uint8 global slots per pixel (0xFF = none)
    unsigned long long* sel;  // per-pixel selection key of the current frontier round
};

struct PatchIn  { int x, y; float depth, dzI, dzJ; unsigned slots; };   // slots: 4 x uint8, 0xFF padded, ascending
struct PatchOut { float conf, depth, dzI, dzJ, nx, ny, nz; unsigned slots; int iterations; int flags; };
// flags: bit0 converged, bit1 opti_success

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

struct Patch {
    // ---- constants of the patch ----
    const DevSettings* st;
    const JobParams* job;
    const ViewParams* views;
    const ViewParams* rv;
    const float* lut;          // srgb2lin in shared memory (mvs_tools.cc:21-95)
    bool act;                  // lane < 25
    float fi, fj;              // sample offsets (patch_optimization.cc:56-64)
    // ---- per-lane sample state ----
    float rx, ry, rz;          // masterViewDirs[k]
    float px, py, pz;          // patchPoints[k]
    float m0, m1, m2;          // masterColorSamples[k] (normalised)
    float e0, e1, e2;          // masterColorSamples[k] - meanX
    // ---- warp-uniform state ----
    float crx, cry, crz;       // masterViewDirs[12]
    float cpx, cpy, cpz;       // patchPoints[12]
    float mfp;                 // footPrintScaled(patchPoints[12])
    float mm, sqrDevX;         // masterMeanCol, sqrDevX
    float depth, dzI, dzJ;
    bool ref_ok;               // sampler->success[refViewNr]
    int nsel;
    int sel[MAX_LOCAL];        // global slots, ascending
    float cs[MAX_LOCAL][3];    // colorScale of the selected views
    float cn[MAX_LOCAL][3];    // cached colour samples of this lane
    float cd[MAX_LOCAL][3];    // cached derivative samples of this lane
    unsigned valid, col_ok, der_ok;   // per selected position: cache filled / colour path ok / derivative path ok
    unsigned avail;            // LocalViewSelection::available over global slots
    int iter;
    bool opti, converged, lvs_ok;
    unsigned n_sets;

    // single_view.h:188-195 (K has the sparsity of camera.cc:125-144)
    __device__ __forceinline__ void project(const ViewParams* V, const LevelParams& L, float X, float Y, float Z,
                                            float& x, float& y) const
    {
        const float c0 = __ldg(&V->w2c[0]) * X + __ldg(&V->w2c[1]) * Y + __ldg(&V->w2c[2]) * Z + __ldg(&V->w2c[3]);
        const float c1 = __ldg(&V->w2c[4]) * X + __ldg(&V->w2c[5]) * Y + __ldg(&V->w2c[6]) * Z + __ldg(&V->w2c[7]);
        const float c2 = __ldg(&V->w2c[8]) * X + __ldg(&V->w2c[9]) * Y + __ldg(&V->w2c[10]) * Z + __ldg(&V->w2c[11]);
        x = (L.ax * c0 + L.cx * c2) / c2 - 0.5f;
        y = (L.ay * c1 + L.cy * c2) / c2 - 0.5f;
    }

    // patch_sampler.cc:274-295 (+ the centre point / master footprint used by every sample set)
    __device__ __forceinline__ void compute_points()
    {
        const float t = depth + fi * dzI + fj * dzJ;
        const bool bad = act && (t <= 0.f);
        if (__any_sync(FULL, bad)) ref_ok = false;
        px = __ldg(&rv->campos[0]) + t * rx;
        py = __ldg(&rv->campos[1]) + t * ry;
        pz = __ldg(&rv->campos[2]) + t * rz;
        cpx = __shfl_sync(FULL, px, CENTER);
        cpy = __shfl_sync(FULL, py, CENTER);
        cpz = __shfl_sync(FULL, pz, CENTER);
        const float z = __ldg(&rv->w2c[8]) * cpx + __ldg(&rv->w2c[9]) * cpy + __ldg(&rv->w2c[10]) * cpz + __ldg(&rv->w2c[11]);
        mfp = z * job->ki0;     // single_view.h:160-164
    }

    // PatchSampler ctor (patch_sampler.cc:19-62) + computeMasterSamples (:298-345)
    __device__ __forceinline__ void init_sampler(int lane, int x, int y)
    {
        act = lane < NS;
        const int di = act ? (lane % 5) - 2 : 0, dj = act ? (lane / 5) - 2 : 0;
        fi = (float)di; fj = (float)dj;
        ref_ok = false; mm = 0.f; sqrDevX = 0.f; valid = col_ok = der_ok = 0u; n_sets = 0u;
        rx = ry = rz = px = py = pz = 0.f; m0 = m1 = m2 = e0 = e1 = e2 = 0.f;
        crx = cry = crz = cpx = cpy = cpz = mfp = 0.f;
        if (x - 2 < 0 || y - 2 < 0 || x + 2 > job->W - 1 || y + 2 > job->H - 1) return;
        // viewRayScaled (single_view.cc:99-106, depthmap.cc:149-156)
        {
            const float fx = (float)(x + di) + 0.5f, fy = (float)(y + dj) + 0.5f;
            float vx = job->ki0 * fx + job->ki2;
            float vy = job->ki4 * fy + job->ki5;
            float vz = 1.0f;
            const float nrm = sqrtf(vx * vx + vy * vy + vz * vz);
            vx /= nrm; vy /= nrm; vz /= nrm;
            rx = __ldg(&rv->rot[0]) * vx + __ldg(&rv->rot[3]) * vy + __ldg(&rv->rot[6]) * vz;
            ry = __ldg(&rv->rot[1]) * vx + __ldg(&rv->rot[4]) * vy + __ldg(&rv->rot[7]) * vz;
            rz = __ldg(&rv->rot[2]) * vx + __ldg(&rv->rot[5]) * vy + __ldg(&rv->rot[8]) * vz;
        }
        crx = __shfl_sync(FULL, rx, CENTER);
        cry = __shfl_sync(FULL, ry, CENTER);
        crz = __shfl_sync(FULL, rz, CENTER);
        ref_ok = true;
        // master colours
        const uchar4 t = job->ref_img[(size_t)(y + dj) * job->ref_pitch + (x + di)];
        m0 = act ? lut[t.x] : 0.f; m1 = act ? lut[t.y] : 0.f; m2 = act ? lut[t.z] : 0.f;
        mm = warp_sum(m0 + m1 + m2) / (3.f * NS);
        if (mm < 0.01f || mm > 0.99f) { ref_ok = false; return; }
        m0 /= mm; m1 /= mm; m2 /= mm;
        const float mx0 = warp_sum(m0) / (float)NS, mx1 = warp_sum(m1) / (float)NS, mx2 = warp_sum(m2) / (float)NS;
        e0 = act ? m0 - mx0 : 0.f; e1 = act ? m1 - mx1 : 0.f; e2 = act ? m2 - mx2 : 0.f;
        sqrDevX = warp_sum(e0 * e0 + e1 * e1 + e2 * e2);
        compute_points();
    }

    // PatchSampler::update (patch_sampler.cc:259-271)
    __device__ __forceinline__ void update()
    {
        ref_ok = true;
        compute_points();
        valid = col_ok = der_ok = 0u;
    }

    // One fused sample set in view V at the current state: fastColAndDeriv (patch_sampler.cc:65-133 +
    // mvs_tools.cc:98-145) and computeNeighColorSamples (patch_sampler.cc:348-393 + mvs_tools.cc:169-199).
    // Returns bit0 = colour path succeeded, bit1 = derivative path succeeded.
    __device__ __forceinline__ unsigned sample(const ViewParams* V, float (&n)[3], float (&d)[3])
    {
        ++n_sets;
        n[0] = n[1] = n[2] = 0.f; d[0] = d[1] = d[2] = 0.f;
        const float nz = __ldg(&V->w2c[8]) * cpx + __ldg(&V->w2c[9]) * cpy + __ldg(&V->w2c[10]) * cpz + __ldg(&V->w2c[11]);
        const float nfp = nz * __ldg(&V->inv_ax0);
        // mfp <= 0 makes the reference throw std::out_of_range (patch_sampler.cc:78-82); it cannot happen for
        // depth > 0 because the centre ray has positive camera z.  Treated as a failed view here.
        if (!(mfp > 0.f) || nfp <= 0.f) return 0u;
        float ratio = nfp / mfp;
        int l = 0;
        while (ratio < 0.5f) { ++l; ratio *= 2.f; }
        const int nl = __ldg(&V->nlevels);
        if (l > nl - 1) l = nl - 1;                     // clampLevel, minLevel = 0 (single_view.h:113-123)
        LevelParams L;
        {
            const float4 k = __ldg(reinterpret_cast<const float4*>(&V->lv[l].ax));
            const int4 g = __ldg(reinterpret_cast<const int4*>(&V->lv[l].w));
            L.ax = k.x; L.ay = k.y; L.cx = k.z; L.cy = k.w; L.w = g.x; L.h = g.y; L.pitch = g.z;
            L.img = reinterpret_cast<const uchar4*>(__ldg(reinterpret_cast<const unsigned long long*>(&V->lv[l].img)));
        }
        // derivative step (patch_sampler.cc:94-100)
        float ax_, ay_, bx_, by_;
        project(V, L, cpx + crx, cpy + cry, cpz + crz, ax_, ay_);
        project(V, L, cpx, cpy, cpz, bx_, by_);
        const float ddx = ax_ - bx_, ddy = ay_ - by_;
        const float dd = sqrtf(ddx * ddx + ddy * ddy);
        const bool dok = dd > 0.f;
        const float step = 1.f / dd;
        float qx, qy;
        project(V, L, px, py, pz, qx, qy);
        const bool inb = qx > 0.f && qx < (float)(L.w - 1) && qy > 0.f && qy < (float)(L.h - 1);
        if (!__all_sync(FULL, inb || !act)) return 0u;
        if (!act) return dok ? 3u : 1u;
        float gx = 0.f, gy = 0.f;
        if (dok) {
            float tx, ty;
            project(V, L, px + rx * step, py + ry * step, pz + rz * step, tx, ty);
            gx = tx - qx; gy = ty - qy;
        }
        const int left = (int)floorf(qx), top = (int)floorf(qy);
        const float fx = qx - (float)left, fy = qy - (float)top;
        const uchar4* r0 = L.img + (size_t)top * L.pitch + left;
        const uchar4* r1 = r0 + L.pitch;
        const uchar4 A = __ldg(r0), B = __ldg(r0 + 1), C = __ldg(r1), D = __ldg(r1 + 1);
        const float a[3] = {lut[A.x], lut[A.y], lut[A.z]};
        const float b[3] = {lut[B.x], lut[B.y], lut[B.z]};
        const float c[3] = {lut[C.x], lut[C.y], lut[C.z]};
        const float e[3] = {lut[D.x], lut[D.y], lut[D.z]};
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float x0 = (1.f - fx) * a[ch] + fx * b[ch];
            const float x3 = (1.f - fx) * c[ch] + fx * e[ch];
            n[ch] = (1.f - fy) * x0 + fy * x3;
            const float der = gx * (b[ch] - a[ch]) + gy * (c[ch] - a[ch]) + (gy * fx + gx * fy) * (a[ch] - b[ch] - c[ch] + e[ch]);
            d[ch] = dok ? der / step : 0.f;
        }
        return dok ? 3u : 1u;
    }

    // getFastNCC on given colour samples (patch_sampler.cc:143-162)
    __device__ __forceinline__ float ncc_of(const float (&n)[3]) const
    {
        const float my0 = warp_sum(n[0]) / (float)NS, my1 = warp_sum(n[1]) / (float)NS, my2 = warp_sum(n[2]) / (float)NS;
        const float y0 = act ? n[0] - my0 : 0.f, y1 = act ? n[1] - my1 : 0.f, y2 = act ? n[2] - my2 : 0.f;
        const float sqrDevY = warp_sum(y0 * y0 + y1 * y1 + y2 * y2);
        const float devXY = warp_sum(e0 * y0 + e1 * y1 + e2 * y2);
        const float tmp = sqrtf(sqrDevX * sqrDevY);
        return tmp > 0.f ? devXY / tmp : -1.f;
    }

    template <int K> __device__ __forceinline__ void ensure()
    {
        if (valid & (1u << K)) return;
        const unsigned r = sample(&views[job->gview[sel[K]]], cn[K], cd[K]);
        valid |= 1u << K;
        if (r & 1u) col_ok |= 1u << K;
        if (r & 2u) der_ok |= 1u << K;
    }
    template <int K> __device__ __forceinline__ float ncc_sel()
    {
        ensure<K>();
        if (!(col_ok & (1u << K))) return -1.f;
        return ncc_of(cn[K]);
    }

    // patch_optimization.cc:81-111
    template <int K> __device__ __forceinline__ bool color_scale_one()
    {
        ensure<K>();
        if (!(col_ok & (1u << K))) return false;    // `return`, not `continue` (patch_optimization.cc:92-93)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float mc = ch == 0 ? m0 : (ch == 1 ? m1 : m2);
            const float nn = cn[K][ch];
            const float ab = warp_sum((mc - nn * cs[K][ch]) * nn);
            const float aa = warp_sum(nn * nn);
            if ((double)fabsf(aa) > 1e-6) {
                cs[K][ch] += ab / aa;
                if ((double)cs[K][ch] > 1e3) opti = false;
            } else
                opti = false;
        }
        return true;
    }
    __device__ __forceinline__ void color_scale()
    {
        if (!st->use_color_scale) return;
        if (nsel > 0 && !color_scale_one<0>()) return;
        if (nsel > 1 && !color_scale_one<1>()) return;
        if (nsel > 2 && !color_scale_one<2>()) return;
        if (nsel > 3 && !color_scale_one<3>()) return;
    }

    // patch_optimization.cc:265-299
    template <int K> __device__ __forceinline__ bool depth_acc(float& num, float& den)
    {
        ensure<K>();
        if (!(der_ok & (1u << K))) return false;
        const float c0 = cs[K][0] * cd[K][0], c1 = cs[K][1] * cd[K][1], c2 = cs[K][2] * cd[K][2];
        const float r0 = m0 - cs[K][0] * cn[K][0], r1 = m1 - cs[K][1] * cn[K][1], r2 = m2 - cs[K][2] * cn[K][2];
        num += c0 * r0 + c1 * r1 + c2 * r2;
        den += c0 * c0 + c1 * c1 + c2 * c2;
        return true;
    }
    __device__ __forceinline__ void depth_step()
    {
        float num = 0.f, den = 0.f;
        bool ok = true;
        if (ok && nsel > 0) ok = depth_acc<0>(num, den);
        if (ok && nsel > 1) ok = depth_acc<1>(num, den);
        if (ok && nsel > 2) ok = depth_acc<2>(num, den);
        if (ok && nsel > 3) ok = depth_acc<3>(num, den);
        if (!ok) { opti = false; return; }
        if (!act) { num = 0.f; den = 0.f; }
        num = warp_sum(num); den = warp_sum(den);
        if (den > 0.f) {
            depth += num / den;
            update();
            opti = ref_ok;
        }
    }

    // patch_optimization.cc:302-364 with matrix_tools.h:392-398,460-475
    template <int K> __device__ __forceinline__ bool normal_acc(double (&A)[6], double (&B)[3])
    {
        ensure<K>();
        if (!(der_ok & (1u << K))) return false;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float mc = ch == 0 ? m0 : (ch == 1 ? m1 : m2);
            const float a0 = cs[K][ch] * cd[K][ch];
            const float a1 = fi * cs[K][ch] * cd[K][ch];
            const float a2 = fj * cs[K][ch] * cd[K][ch];
            const float b = mc - cs[K][ch] * cn[K][ch];
            A[0] += (double)(a0 * a0); A[1] += (double)(a0 * a1); A[2] += (double)(a0 * a2);
            A[3] += (double)(a1 * a1); A[4] += (double)(a1 * a2); A[5] += (double)(a2 * a2);
            B[0] += (double)(a0 * b); B[1] += (double)(a1 * b); B[2] += (double)(a2 * b);
        }
        return true;
    }
    __device__ __forceinline__ void normal_step()
    {
        if (!lvs_ok) return;
        double A[6] = {0, 0, 0, 0, 0, 0}, B[3] = {0, 0, 0};
        bool ok = true;
        if (ok && nsel > 0) ok = normal_acc<0>(A, B);
        if (ok && nsel > 1) ok = normal_acc<1>(A, B);
        if (ok && nsel > 2) ok = normal_acc<2>(A, B);
        if (ok && nsel > 3) ok = normal_acc<3>(A, B);
        if (!ok) { opti = false; return; }
#pragma unroll
        for (int q = 0; q < 6; ++q) A[q] = warp_sum(act ? A[q] : 0.0);
#pragma unroll
        for (int q = 0; q < 3; ++q) B[q] = warp_sum(act ? B[q] : 0.0);
        const double m[9] = {A[0], A[1], A[2], A[1], A[3], A[4], A[2], A[4], A[5]};
        const double det = m[0] * m[4] * m[8] + m[1] * m[5] * m[6] + m[2] * m[3] * m[7]
                         - m[2] * m[4] * m[6] - m[1] * m[3] * m[8] - m[0] * m[5] * m[7];
        if (det == 0.0) { opti = false; return; }
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
#pragma unroll
        for (int q = 0; q < 9; ++q) inv[q] /= det;
        const float X0 = (float)(inv[0] * B[0] + inv[1] * B[1] + inv[2] * B[2]);
        const float X1 = (float)(inv[3] * B[0] + inv[4] * B[1] + inv[5] * B[2]);
        const float X2 = (float)(inv[6] * B[0] + inv[7] * B[1] + inv[8] * B[2]);
        dzI += X1; dzJ += X2; depth += X0;
        update();
        opti = ref_ok;
    }

    // ---- sorted insert / erase on the selected set (std::set semantics), compile-time indices only ----
    __device__ __forceinline__ void sel_erase_mask(unsigned mask)
    {
#pragma unroll
        for (int k = MAX_LOCAL - 1; k >= 0; --k) {
            if ((mask >> k) & 1u) {
#pragma unroll
                for (int q = k; q < MAX_LOCAL - 1; ++q) {
                    sel[q] = sel[q + 1];
                    cs[q][0] = cs[q + 1][0]; cs[q][1] = cs[q + 1][1]; cs[q][2] = cs[q + 1][2];
                }
                --nsel;
            }
        }
        valid = col_ok = der_ok = 0u;
    }
    __device__ __forceinline__ void sel_insert(int slot, float cs_init)
    {
        int pos = 0;
#pragma unroll
        for (int k = 0; k < MAX_LOCAL; ++k) if (k < nsel && sel[k] < slot) ++pos;
#pragma unroll
        for (int q = MAX_LOCAL - 1; q >= 1; --q) {
            if (q > pos) {
                sel[q] = sel[q - 1];
                cs[q][0] = cs[q - 1][0]; cs[q][1] = cs[q - 1][1]; cs[q][2] = cs[q - 1][2];
            }
        }
#pragma unroll
        for (int q = 0; q < MAX_LOCAL; ++q) {
            if (q == pos) { sel[q] = slot; cs[q][0] = cs[q][1] = cs[q][2] = cs_init; }
        }
        ++nsel;
        valid = col_ok = der_ok = 0u;
    }

    // mvs_tools.h:56-69
    static __device__ __forceinline__ float plx_weight(float p)
    {
        if (p < 0.f || p > 180.f) return 0.f;
        const float sigma = (p <= 20.f) ? 5.f : 15.f;
        const float dlt = p - 20.f;
        return expf(-(dlt * dlt) / (2.f * sigma * sigma));
    }
    static __device__ __forceinline__ float clamp1(float v) { return v < -1.f ? -1.f : (v > 1.f ? 1.f : v); }
    static __device__ __forceinline__ float deg_acos(float dp) { return acosf(dp) * 180.f / 3.141592653589793f; }

    // LocalViewSelection::performVS (local_view_selection.cc:57-147); lane i evaluates candidate slot i.
    __device__ __forceinline__ void lvs_perform(int lane, float cs_init)
    {
        const unsigned N = st->nr_recon_neighbors;
        if ((unsigned)nsel == N) { lvs_ok = true; return; }
        // refDir
        float rdx = cpx - __ldg(&rv->campos[0]), rdy = cpy - __ldg(&rv->campos[1]), rdz = cpz - __ldg(&rv->campos[2]);
        {
            const float nn = sqrtf(rdx * rdx + rdy * rdy + rdz * rdz);
            rdx /= nn; rdy /= nn; rdz /= nn;
        }
        // NCC of every available global view (one sample set each); lane i keeps candidate i's value
        float my_ncc = 0.f;
        const int G = job->n_global;
        for (int i = 0; i < G; ++i) {
            if (!((avail >> i) & 1u)) continue;
            float tn[3], td[3];
            const unsigned r = sample(&views[job->gview[i]], tn, td);
            const float v = (r & 1u) ? ncc_of(tn) : -1.f;
            if (v < st->min_ncc) { avail &= ~(1u << i); continue; }
            if (lane == i) my_ncc = v;
        }
        // per-lane candidate geometry (viewDir, epipolarPlane, footprint)
        float vdx = 0.f, vdy = 0.f, vdz = 1.f, epx = 0.f, epy = 0.f, epz = 1.f, nfp = 1.f;
        if (lane < G) {
            const ViewParams* V = &views[job->gview[lane]];
            vdx = cpx - __ldg(&V->campos[0]); vdy = cpy - __ldg(&V->campos[1]); vdz = cpz - __ldg(&V->campos[2]);
            const float nn = sqrtf(vdx * vdx + vdy * vdy + vdz * vdz);
            vdx /= nn; vdy /= nn; vdz /= nn;
            epx = vdy * rdz - vdz * rdy; epy = vdz * rdx - vdx * rdz; epz = vdx * rdy - vdy * rdx;
            const float en = sqrtf(epx * epx + epy * epy + epz * epz);
            epx /= en; epy /= en; epz /= en;
            const float z = __ldg(&V->w2c[8]) * cpx + __ldg(&V->w2c[9]) * cpy + __ldg(&V->w2c[10]) * cpz + __ldg(&V->w2c[11]);
            nfp = z * __ldg(&V->inv_ax0);
        }
        bool found = true;
        while ((unsigned)nsel < N && found) {
            found = false;
            float score = -1.f;
            if (lane < G && ((avail >> lane) & 1u)) {
                score = my_ncc;
                if (mfp / nfp < 0.5f) score *= 0.01f;
                float dp = clamp1(rdx * vdx + rdy * vdy + rdz * vdz);
                score *= plx_weight(deg_acos(dp));
            }
            // parallax / epipolar terms against every already selected view (geometry broadcast from its lane)
#pragma unroll
            for (int k = 0; k < MAX_LOCAL; ++k) {
                if (k < nsel) {
                    const int s = sel[k];
                    const float sx = __shfl_sync(FULL, vdx, s), sy = __shfl_sync(FULL, vdy, s), sz = __shfl_sync(FULL, vdz, s);
                    const float ex = __shfl_sync(FULL, epx, s), ey = __shfl_sync(FULL, epy, s), ez = __shfl_sync(FULL, epz, s);
                    if (lane < G && ((avail >> lane) & 1u)) {
                        float dp = clamp1(sx * vdx + sy * vdy + sz * vdz);
                        score *= plx_weight(deg_acos(dp));
                        dp = clamp1(epx * ex + epy * ey + epz * ez);
                        float angle = fabsf(deg_acos(dp));
                        if (angle > 90.f) angle = 180.f - angle;
                        angle = fmaxf(angle, 1.f);
                        if (angle < st->min_parallax) score *= angle / st->min_parallax;
                    }
                }
            }
            const bool cand = lane < G && ((avail >> lane) & 1u) && (score > 0.f);   // NaN compares false, like `score > maxScore`
            const float best = warp_max(cand ? score : -1.f);
            const unsigned winners = __ballot_sync(FULL, cand && score == best);
            if (best > 0.f && winners) {
                const int w = __ffs(winners) - 1;        // strict '>' in index order: lowest index wins ties
                found = true;
                sel_insert(w, cs_init);
                avail &= ~(1u << w);
            }
        }
        if ((unsigned)nsel == N) lvs_ok = true;
    }

    // PatchOptimization ctor (patch_optimization.cc:21-78) incl. LocalViewSelection ctor (local_view_selection.cc:19-54)
    __device__ __forceinline__ void init(int lane, const PatchIn& in)
    {
        rv = &views[job->ref_view];
        depth = in.depth; dzI = in.dzI; dzJ = in.dzJ;
        iter = 0; opti = true; converged = false; lvs_ok = false;
        nsel = 0; avail = 0u;
#pragma unroll
        for (int k = 0; k < MAX_LOCAL; ++k) {
            sel[k] = 0xFF; cs[k][0] = cs[k][1] = cs[k][2] = 0.f;
            cn[k][0] = cn[k][1] = cn[k][2] = 0.f; cd[k][0] = cd[k][1] = cd[k][2] = 0.f;
        }
        init_sampler(lane, in.x, in.y);
        // propagated ids arrive ascending
#pragma unroll
        for (int k = 0; k < MAX_LOCAL; ++k) {
            const int s = (in.slots >> (8 * k)) & 0xFF;
            if (s != 0xFF) { sel[k] = s; nsel = k + 1; }
        }
        if (!ref_ok) { opti = false; return; }
        const unsigned N = st->nr_recon_neighbors;
        if ((unsigned)nsel == N) lvs_ok = true;
        else if ((unsigned)nsel > N) nsel = 0;
        avail = job->n_global >= 32 ? FULL : ((1u << job->n_global) - 1u);
#pragma unroll
        for (int k = 0; k < MAX_LOCAL; ++k) if (k < nsel) avail &= ~(1u << sel[k]);
        const float cs_init = 1.f / mm;
#pragma unroll
        for (int k = 0; k < MAX_LOCAL; ++k) cs[k][0] = cs[k][1] = cs[k][2] = cs_init;
        lvs_perform(lane, cs_init);
        if (!lvs_ok) { opti = false; return; }
        color_scale();
    }

    __device__ __forceinline__ float ncc_at(int k)
    {
        switch (k) { case 0: return ncc_sel<0>(); case 1: return ncc_sel<1>(); case 2: return ncc_sel<2>(); default: return ncc_sel<3>(); }
    }

    // PatchOptimization::doAutoOptimization (patch_optimization.cc:170-242)
    __device__ __forceinline__ void auto_optimize(int lane)
    {
        if (!lvs_ok || !opti) return;
        while (iter < 4 && opti) { depth_step(); ++iter; }
        bool viewRemoved = false;
        while ((unsigned)iter < st->max_iterations && lvs_ok && opti) {
            float oldN[MAX_LOCAL];
            const int n_old = nsel;
#pragma unroll
            for (int k = 0; k < MAX_LOCAL; ++k) oldN[k] = (k < n_old) ? ncc_at(k) : 0.f;
            opti = false;
            if (iter % 5 == 4 || viewRemoved) {
                normal_step();
                color_scale();
                viewRemoved = false;
            } else
                depth_step();
            if (!opti) return;
            bool conv = true;
            unsigned tbr = 0u;
#pragma unroll
            for (int k = 0; k < MAX_LOCAL; ++k) {
                if (k < n_old) {
                    const float v = ncc_at(k);
                    const float df = fabsf(v - oldN[k]);
                    if (df > st->min_refine_diff) conv = false;
                    if (v < st->accept_ncc || (iter == 14 && df > st->min_refine_diff)) { tbr |= 1u << k; viewRemoved = true; }
                }
            }
            if (viewRemoved) {
                // LocalViewSelection::replaceViews (local_view_selection.cc:150-160)
                sel_erase_mask(tbr);
                lvs_ok = false;
                lvs_perform(lane, 1.f / mm);
                if (!lvs_ok) return;
                color_scale();
            } else if (!opti) {
                return;
            } else if (conv) {
                converged = true;
                return;
            }
            ++iter;
        }
    }

    // PatchOptimization::computeConfidence (patch_optimization.cc:114-142) + getPatchNormal (patch_sampler.cc:243-256)
    __device__ __forceinline__ void finish(PatchOut& out)
    {
        out.depth = depth; out.dzI = dzI; out.dzJ = dzJ;
        out.iterations = iter;
        out.flags = (converged ? 1 : 0) | (opti ? 2 : 0);
        unsigned s = 0u;
#pragma unroll
        for (int k = 0; k < MAX_LOCAL; ++k) s |= (unsigned)((k < nsel) ? (sel[k] & 0xFF) : 0xFF) << (8 * k);
        out.slots = s;
        out.conf = 0.f; out.nx = out.ny = out.nz = 0.f;
        if (!converged) return;
        float mean = 0.f;
#pragma unroll
        for (int k = 0; k < MAX_LOCAL; ++k) if (k < nsel) mean += ncc_at(k);
        mean /= (float)nsel;
        const float score = (mean - st->accept_ncc) / (1.f - st->accept_ncc);
        const float ax_ = __shfl_sync(FULL, px, CENTER + 2) - __shfl_sync(FULL, px, CENTER - 2);
        const float ay_ = __shfl_sync(FULL, py, CENTER + 2) - __shfl_sync(FULL, py, CENTER - 2);
        const float az_ = __shfl_sync(FULL, pz, CENTER + 2) - __shfl_sync(FULL, pz, CENTER - 2);
        const float bx_ = __shfl_sync(FULL, px, 2) - __shfl_sync(FULL, px, NS - 1 - 2);
        const float by_ = __shfl_sync(FULL, py, 2) - __shfl_sync(FULL, py, NS - 1 - 2);
        const float bz_ = __shfl_sync(FULL, pz, 2) - __shfl_sync(FULL, pz, NS - 1 - 2);
        float nx = ay_ * bz_ - az_ * by_, ny = az_ * bx_ - ax_ * bz_, nz = ax_ * by_ - ay_ * bx_;
        const float nn = sqrtf(nx * nx + ny * ny + nz * nz);
        nx /= nn; ny /= nn; nz /= nn;
        out.nx = nx; out.ny = ny; out.nz = nz;
        const float dotP = -(nx * crx + ny * cry + nz * crz);
        out.conf = (dotP < 0.2f) ? 0.f : score;
    }
};

// One PatchOptimization executed by the calling warp.  Returns the number of fused sample sets drawn.
__device__ __forceinline__ unsigned optimize_patch(const DevSettings* st, const JobParams* job, const ViewParams* views,
                                                   const float* lut, int lane, const PatchIn& in, PatchOut& out)
{
    Patch p;
    p.st = st; p.job = job; p.views = views; p.lut = lut;
    p.init(lane, in);
    p.auto_optimize(lane);
    p.finish(out);
    return p.n_sets;
}

} // namespace b200mvs
