// Device-side patch optimisation: EIGHT lanes per patch (four patches per warp), each lane owning up to four of the
// 25 samples of the 5x5 patch.
//
// Restates, for the GPU, what one mvs::PatchOptimization does in the reference
// (libs/dmrecon/patch_optimization.cc:21-364 with PatchSampler patch_sampler.cc:19-393,
// LocalViewSelection local_view_selection.cc:19-160 and mvs_tools.cc:98-199); SURVEY.md Appendix A
// is the line-by-line behavioural spec.  Differences in STRUCTURE (not in results):
//   * the colour-only sample set (computeNeighColorSamples) and the colour+derivative sample set
//     (fastColAndDeriv) of one view at one patch state use identical bilinear formulas
//     (mvs_tools.cc:119-128 vs :188-197), so ONE fused sample set per (state, view) is drawn;
//   * the optimisation is organised as one PASS per patch state (depth, dzI, dzJ): a loop over the
//     selected views that draws the fused sample set and immediately reduces it to what the reference
//     reads at that state - NCC of the view (getFastNCC), the colour-scale update when one is due
//     (computeColorScale), and the Gauss-Newton terms of the next step (optimizeDepthOnly /
//     optimizeDepthAndNormal);
//   * a group of 8 lanes is an independent unit: every collective names only the group's lanes, so the four
//     groups of a warp run different patches, in different stages, and re-converge at the one pass() call
//     site; a group that finishes fetches its next patch without waiting for the others (Patch::begin /
//     step / finish are driven by a flat loop in the kernels);
//   * a lane's four sample slots are unrolled: rays, master colours, drawn colours and derivatives of its samples stay in
//     registers between the sampling and the reductions that consume them (a rolled variant that kept them in shared
//     memory spent as many instructions re-reading them as on sampling, profiles/r2_notes.md); the four independent
//     sample computations give the scheduler instruction-level parallelism to hide the texel loads;
//   * the four bilinear taps of a sample come from ONE 16-byte load of a "quad" texel (the 2x2 neighbourhood
//     of every pixel is stored contiguously, DESIGN.md "Data layout"); sRGB code values are linearised through a
//     copy of the 256-entry table that is replicated per lane (no bank conflicts, mvs_tools.cc:21-95);
//   * small per-view arrays (selected slots, colour scales, NCCs) live one element per lane of the group and are
//     read with shuffles; the sums of the 3x3 normal equations (patch_optimization.cc:326-343) are formed
//     per lane in fp32 (<= 16 products) and across the lanes in fp64, then solved in fp64 inside the pass.
#pragma once
#if defined(B200MVS_HOST_EMU)
#include "simt_emu.h"      // tests/emu: runs this very file on the CPU, 32 host threads per warp (test infrastructure)
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <cstddef>

namespace b200mvs {

constexpr int MAX_LEVELS = 12;
constexpr int MAX_GLOBAL = 32;
constexpr int MAX_LOCAL = 4;
constexpr unsigned FULL = 0xffffffffu;
constexpr int NS = 25;
constexpr int CENTER = 12;   // patch_sampler.cc:73,96
constexpr int GROUP = 8;     // lanes per patch
constexpr int SLOTS = 4;     // sample slots per lane: slot 0 = samples 24.. (+ two helper points), slot s = samples 8(s-1)..8(s-1)+7
constexpr int LUT_REP = 32;  // replicas of the sRGB table in shared memory (one per lane)

struct alignas(16) LevelParams {   // ImagePyramidLevel (image_pyramid.h:28-59): K = [ax 0 cx; 0 ay cy; 0 0 1]
    float ax, ay, cx, cy;
    int w, h;
    int pitch;                // in texels (uchar4), also the pitch of the quad image (uint4)
    int pad;
    const uchar4* img;        // RGBX8, row pitch 16-byte aligned
    const uint4* quad;        // quad[y * pitch + x] = {img(x,y), img(x+1,y), img(x,y+1), img(x+1,y+1)} (clamped at the border)
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
    int tiles_x;              // 16x16-pixel tiles per row of the reference level (frontier sort)
    float* depth; float* conf; float* dz; float* normal;
    unsigned* slots;          // 4 x uint8 global slots per pixel (0xFF = none)
    unsigned long long* sel;  // per-pixel selection key of the current frontier round
    long long tile_base;      // first tile bin of this job
};

struct PatchIn  { int x, y; float depth, dzI, dzJ; unsigned slots; };   // slots: 4 x uint8, 0xFF padded, ascending
struct PatchOut { float conf, depth, dzI, dzJ, nx, ny, nz; unsigned slots; int iterations; int flags; };
// flags: bit0 converged, bit1 opti_success

// Hot-path approximations (each <= 2 ulp): MUFU reciprocal / reciprocal square root instead of the IEEE division and
// square root sequences.  The reference itself is built with -funsafe-math-optimizations (Makefile.inc:5), i.e. without
// IEEE guarantees for these operations; the effect on parity is measured by tests/test_gpu_parity.py.  Everything that
// decides integers on the host (global view selection, seeds, pyramid) stays IEEE.
#if defined(B200MVS_HOST_EMU)
__device__ __forceinline__ float rcp_fast(float x) { return 1.f / x; }
__device__ __forceinline__ float rsqrt_fast(float x) { return 1.f / sqrtf(x); }
#else
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rsqrt_fast(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif

struct Patch {
    // ---- constants of the thread ----
    const DevSettings* st;
    const ViewParams* views;
    const float* lutw;         // replicated srgb2lin table + lane: value v of this lane at lutw[v * LUT_REP]
    int gl;                    // lane within the group
    float fi[SLOTS], fj[SLOTS];   // patch offsets of the lane's samples (patch_optimization.cc:56-64), 0 for the helper points
    // ---- per-lane sample state ----
    float ray[SLOTS][3];       // masterViewDirs[k]
    float mc[SLOTS][3];        // masterColorSamples[k] (normalised)
    float sn[SLOTS][3];        // neighColorSamples[k] of the sample set drawn last
    float sd[SLOTS][3];        // its colour derivatives along the ray
    unsigned gmask;            // lanes of the group
    // ---- constants of the patch ----
    const JobParams* job;
    const ViewParams* rv;
    int x0, y0;
    // ---- group-uniform state ----
    float mx0, mx1, mx2;       // meanX per channel (patch_sampler.cc:333-339)
    float crx, cry, crz;       // masterViewDirs[12]
    float cpx, cpy, cpz;       // patchPoints[12]
    float mfp, inv_mfp;        // footPrintScaled(patchPoints[12]) and its reciprocal
    float mm, sqrDevX;         // masterMeanCol, sqrDevX
    float depth, dzI, dzJ;
    bool ref_ok;               // sampler->success[refViewNr]
    int nsel;
    unsigned avail;            // LocalViewSelection::available over global slots
    int iter;
    bool opti, converged, lvs_ok;
    unsigned n_sets;
    // state machine of doAutoOptimization
    int stage;
    bool viewRemoved, was_normal, normal;
    float old;                 // oldNCC of selected view `gl`
    // ---- lane-distributed small arrays: lane k (< nsel) of the group holds element k ----
    int sel_l;                 // selected global slot (ascending over lanes)
    float cs0_l, cs1_l, cs2_l; // colorScale of selected view k
    float ncc_l;               // NCC of selected view k at the state of the last pass
    float cand0, cand1, cand2, cand3;   // NCC of candidate global slot gl + 8 j (local view selection)
    // ---- results of the last pass (valid for the current state and selected set) ----
    unsigned p_col_ok, p_der_ok;   // bit k: colour / derivative path of selected view k succeeded
    float p_num, p_den;            // optimizeDepthOnly sums
    float nX0, nX1, nX2;           // solution of the 3x3 normal equations of optimizeDepthAndNormal
    bool n_singular;               // detATA == 0 (patch_optimization.cc:347-351)
    bool p_has_normal, p_has_ncc;

    enum Stage { LVS_CTOR, CTOR, FIRST, PRE, POST, LVS_REPL, REPL, DONE };

    // ---- group collectives ----
    template <typename T> __device__ __forceinline__ T gbcast(T v, int src) const { return __shfl_sync(gmask, v, src, GROUP); }
    __device__ __forceinline__ float gsum(float v) const
    {
#pragma unroll
        for (int o = GROUP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
        return v;
    }
    __device__ __forceinline__ double gsum(double v) const
    {
#pragma unroll
        for (int o = GROUP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
        return v;
    }
    __device__ __forceinline__ float gmax(float v) const
    {
#pragma unroll
        for (int o = GROUP / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(gmask, v, o));
        return v;
    }
    __device__ __forceinline__ bool gany(bool p) const { return __any_sync(gmask, p) != 0; }
    __device__ __forceinline__ unsigned gballot(bool p) const      // bit k = lane k of the group
    {
        const unsigned b = __ballot_sync(gmask, p) & gmask;
        return (b >> (__ffs(gmask) - 1)) & 0xFFu;
    }

    // sample handled by (slot s, this lane): index, kind and patch offsets (patch_optimization.cc:56-64)
    __device__ __forceinline__ int sample_index(int s) const { return s == 0 ? 24 + gl : (s - 1) * GROUP + gl; }
    __device__ __forceinline__ bool is_real(int s) const { return s != 0 || gl == 0; }

    // single_view.h:188-195 (K has the sparsity of camera.cc:125-144).  x = (K cp).x / cp.z - 0.5 is evaluated with one
    // reciprocal shared by x and y (<= 2 ulp from the reference's two divisions).
    __device__ __forceinline__ void project(const float (&w)[12], const LevelParams& L, float X, float Y, float Z,
                                            float& x, float& y) const
    {
        const float c0 = w[0] * X + w[1] * Y + w[2] * Z + w[3];
        const float c1 = w[4] * X + w[5] * Y + w[6] * Z + w[7];
        const float c2 = w[8] * X + w[9] * Y + w[10] * Z + w[11];
        const float ic2 = rcp_fast(c2);
        x = (L.ax * c0 + L.cx * c2) * ic2 - 0.5f;
        y = (L.ay * c1 + L.cy * c2) * ic2 - 0.5f;
    }

    // patch_sampler.cc:274-295 (+ the centre point / master footprint used by every sample set)
    __device__ __forceinline__ void compute_points()
    {
        bool bad = false;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const float t = depth + fi[s] * dzI + fj[s] * dzJ;
            bad |= is_real(s) && (t <= 0.f);
        }
        if (gany(bad)) ref_ok = false;
        cpx = __ldg(&rv->campos[0]) + depth * crx;       // the centre sample has offsets (0, 0): t = depth
        cpy = __ldg(&rv->campos[1]) + depth * cry;
        cpz = __ldg(&rv->campos[2]) + depth * crz;
        const float z = __ldg(&rv->w2c[8]) * cpx + __ldg(&rv->w2c[9]) * cpy + __ldg(&rv->w2c[10]) * cpz + __ldg(&rv->w2c[11]);
        mfp = z * job->ki0;     // single_view.h:160-164
        inv_mfp = rcp_fast(mfp);
    }

    // PatchSampler ctor (patch_sampler.cc:19-62) + computeMasterSamples (:298-345)
    __device__ __forceinline__ void init_sampler(int x, int y)
    {
        ref_ok = false; mm = 0.f; sqrDevX = 0.f;
        mx0 = mx1 = mx2 = 0.f;
        crx = cry = crz = cpx = cpy = cpz = mfp = inv_mfp = 0.f;
        if (x - 2 < 0 || y - 2 < 0 || x + 2 > job->W - 1 || y + 2 > job->H - 1) return;
        const float r0 = __ldg(&rv->rot[0]), r1 = __ldg(&rv->rot[1]), r2 = __ldg(&rv->rot[2]), r3 = __ldg(&rv->rot[3]), r4 = __ldg(&rv->rot[4]);
        const float r5 = __ldg(&rv->rot[5]), r6 = __ldg(&rv->rot[6]), r7 = __ldg(&rv->rot[7]), r8 = __ldg(&rv->rot[8]);
        float sum = 0.f;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const bool real = is_real(s);
            const int di = (int)fi[s], dj = (int)fj[s];
            // viewRayScaled (single_view.cc:99-106, depthmap.cc:149-156)
            const float fx = (float)(x + di) + 0.5f, fy = (float)(y + dj) + 0.5f;
            float vx = job->ki0 * fx + job->ki2;
            float vy = job->ki4 * fy + job->ki5;
            float vz = 1.0f;
            const float nrm = sqrtf(vx * vx + vy * vy + vz * vz);
            vx /= nrm; vy /= nrm; vz /= nrm;
            ray[s][0] = r0 * vx + r3 * vy + r6 * vz;
            ray[s][1] = r1 * vx + r4 * vy + r7 * vz;
            ray[s][2] = r2 * vx + r5 * vy + r8 * vz;
            // master colours
            mc[s][0] = mc[s][1] = mc[s][2] = 0.f;
            if (real) {
                const uchar4 t = job->ref_img[(size_t)(y + dj) * job->ref_pitch + (x + di)];
                mc[s][0] = lutw[t.x * LUT_REP]; mc[s][1] = lutw[t.y * LUT_REP]; mc[s][2] = lutw[t.z * LUT_REP];
            }
            sum += mc[s][0] + mc[s][1] + mc[s][2];
        }
        // sample 12 = slot 2, lane 4
        crx = gbcast(ray[2][0], CENTER - GROUP); cry = gbcast(ray[2][1], CENTER - GROUP); crz = gbcast(ray[2][2], CENTER - GROUP);
        ref_ok = true;
        mm = gsum(sum) / (3.f * NS);
        if (mm < 0.01f || mm > 0.99f) { ref_ok = false; return; }
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            mc[s][0] /= mm; mc[s][1] /= mm; mc[s][2] /= mm;
            s0 += mc[s][0]; s1 += mc[s][1]; s2 += mc[s][2];                 // helper points hold 0
        }
        mx0 = gsum(s0) / (float)NS; mx1 = gsum(s1) / (float)NS; mx2 = gsum(s2) / (float)NS;
        float dev = 0.f;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            if (!is_real(s)) continue;
            const float e0 = mc[s][0] - mx0, e1 = mc[s][1] - mx1, e2 = mc[s][2] - mx2;
            dev += e0 * e0 + e1 * e1 + e2 * e2;
        }
        sqrDevX = gsum(dev);
        compute_points();
    }

    // PatchSampler::update (patch_sampler.cc:259-271)
    __device__ __forceinline__ void update()
    {
        ref_ok = true;
        compute_points();
    }

    // One fused sample set in view V at the current state: fastColAndDeriv (patch_sampler.cc:65-133 +
    // mvs_tools.cc:98-145) and computeNeighColorSamples (patch_sampler.cc:348-393 + mvs_tools.cc:169-199).
    // Colours and derivatives of the lane's samples go to sn / sd.  Returns bit0 = colour path succeeded,
    // bit1 = derivative path succeeded.
    __device__ __forceinline__ unsigned sample(const ViewParams* V)
    {
        if (gl == 0) ++n_sets;
        float w[12];
        {
            const float4 a = __ldg(reinterpret_cast<const float4*>(&V->w2c[0]));
            const float4 b = __ldg(reinterpret_cast<const float4*>(&V->w2c[4]));
            const float4 c = __ldg(reinterpret_cast<const float4*>(&V->w2c[8]));
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
            w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
        }
        const float nz = w[8] * cpx + w[9] * cpy + w[10] * cpz + w[11];
        const float nfp = nz * __ldg(&V->inv_ax0);
        // mfp <= 0 makes the reference throw std::out_of_range (patch_sampler.cc:78-82); it cannot happen for
        // depth > 0 because the centre ray has positive camera z.  Treated as a failed view here.
        if (!(mfp > 0.f) || nfp <= 0.f) return 0u;
        float ratio = nfp * inv_mfp;
        int l = 0;
        while (ratio < 0.5f) { ++l; ratio *= 2.f; }
        const int nl = __ldg(&V->nlevels);
        if (l > nl - 1) l = nl - 1;                     // clampLevel, minLevel = 0 (single_view.h:113-123)
        LevelParams L;
        {
            const float4 k = __ldg(reinterpret_cast<const float4*>(&V->lv[l].ax));
            const int4 g = __ldg(reinterpret_cast<const int4*>(&V->lv[l].w));
            L.ax = k.x; L.ay = k.y; L.cx = k.z; L.cy = k.w; L.w = g.x; L.h = g.y; L.pitch = g.z;
            L.quad = reinterpret_cast<const uint4*>(__ldg(reinterpret_cast<const unsigned long long*>(&V->lv[l].quad)));
        }
        const float c0x = __ldg(&rv->campos[0]), c0y = __ldg(&rv->campos[1]), c0z = __ldg(&rv->campos[2]);
        // projections of all slots first (slot 0 carries the two helper points of the derivative step,
        // patch_sampler.cc:94-100: lane 1 projects patchPoints[12] + masterViewDirs[12], lane 2 patchPoints[12])
        float X[SLOTS], Y[SLOTS], Z[SLOTS], qx[SLOTS], qy[SLOTS];
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            if (s == 0) {
                const float a = gl == 1 ? 1.f : 0.f;
                const float t = depth + fi[0] * dzI + fj[0] * dzJ;
                X[0] = gl == 0 ? c0x + t * ray[0][0] : cpx + a * crx;
                Y[0] = gl == 0 ? c0y + t * ray[0][1] : cpy + a * cry;
                Z[0] = gl == 0 ? c0z + t * ray[0][2] : cpz + a * crz;
            } else {
                const float t = depth + fi[s] * dzI + fj[s] * dzJ;
                X[s] = c0x + t * ray[s][0]; Y[s] = c0y + t * ray[s][1]; Z[s] = c0z + t * ray[s][2];
            }
            project(w, L, X[s], Y[s], Z[s], qx[s], qy[s]);
        }
        const float ddx = gbcast(qx[0], 1) - gbcast(qx[0], 2);
        const float ddy = gbcast(qy[0], 1) - gbcast(qy[0], 2);
        const float dd2 = ddx * ddx + ddy * ddy;
        const float dd = dd2 * rsqrt_fast(dd2);        // |.|; NaN for dd2 == 0, which fails `d > 0` like the reference's 0
        const bool dok = dd > 0.f;
        const float step = rcp_fast(dd);
        const float wm1 = (float)(L.w - 1), hm1 = (float)(L.h - 1);
        bool oob = false;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s)
            if (is_real(s) && !(qx[s] > 0.f && qx[s] < wm1 && qy[s] > 0.f && qy[s] < hm1)) oob = true;
        if (gany(oob)) return 0u;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            sn[s][0] = sn[s][1] = sn[s][2] = 0.f; sd[s][0] = sd[s][1] = sd[s][2] = 0.f;
            if (!is_real(s)) continue;
            float gx = 0.f, gy = 0.f;
            if (dok) {
                float tx, ty;
                project(w, L, X[s] + ray[s][0] * step, Y[s] + ray[s][1] * step, Z[s] + ray[s][2] * step, tx, ty);
                gx = tx - qx[s]; gy = ty - qy[s];
            }
            const int left = (int)floorf(qx[s]), top = (int)floorf(qy[s]);
            const float fx = qx[s] - (float)left, fy = qy[s] - (float)top;
            const uint4 Q = __ldg(L.quad + (size_t)top * L.pitch + left);
            float a[3], b[3], c[3], e[3];
            a[0] = lutw[(Q.x & 0xFF) * LUT_REP]; a[1] = lutw[((Q.x >> 8) & 0xFF) * LUT_REP]; a[2] = lutw[((Q.x >> 16) & 0xFF) * LUT_REP];
            b[0] = lutw[(Q.y & 0xFF) * LUT_REP]; b[1] = lutw[((Q.y >> 8) & 0xFF) * LUT_REP]; b[2] = lutw[((Q.y >> 16) & 0xFF) * LUT_REP];
            c[0] = lutw[(Q.z & 0xFF) * LUT_REP]; c[1] = lutw[((Q.z >> 8) & 0xFF) * LUT_REP]; c[2] = lutw[((Q.z >> 16) & 0xFF) * LUT_REP];
            e[0] = lutw[(Q.w & 0xFF) * LUT_REP]; e[1] = lutw[((Q.w >> 8) & 0xFF) * LUT_REP]; e[2] = lutw[((Q.w >> 16) & 0xFF) * LUT_REP];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float x0 = (1.f - fx) * a[ch] + fx * b[ch];
                const float x3 = (1.f - fx) * c[ch] + fx * e[ch];
                sn[s][ch] = (1.f - fy) * x0 + fy * x3;
                const float der = gx * (b[ch] - a[ch]) + gy * (c[ch] - a[ch]) + (gy * fx + gx * fy) * (a[ch] - b[ch] - c[ch] + e[ch]);
                sd[s][ch] = dok ? der * dd : 0.f;        // deriv /= stepSize with stepSize = 1 / d (patch_sampler.cc:100,129-130)
            }
        }
        return dok ? 3u : 1u;
    }

    // One pass at the current state (see the header comment).
    //   candidates : false -> over the selected views; true -> over the AVAILABLE global views, only their NCC is
    //                computed (first half of LocalViewSelection::performVS, local_view_selection.cc:73-85)
    //   cs_pending : a computeColorScale() is due at this state (patch_optimization.cc:77,198,230)
    //   want_ncc   : the NCCs of this state are read by the reference (getFastNCC, patch_optimization.cc:192,213,126)
    //   want_normal: the next Gauss-Newton step is optimizeDepthAndNormal (else optimizeDepthOnly)
    // This is the only place a sample set is drawn, so its code exists once in the kernel.
    __device__ __forceinline__ void pass(bool candidates, bool cs_pending, bool want_ncc, bool want_normal)
    {
        float num = 0.f, den = 0.f;
        float A0 = 0.f, A1 = 0.f, A2 = 0.f, A3 = 0.f, A4 = 0.f, A5 = 0.f, B0 = 0.f, B1 = 0.f, B2 = 0.f;
        bool cs_active = cs_pending && st->use_color_scale;
        if (!candidates) { p_col_ok = p_der_ok = 0u; }
        const int count = candidates ? job->n_global : nsel;
#pragma unroll 1
        for (int k = 0; k < count; ++k) {
            int slot = k;
            if (candidates) { if (!((avail >> k) & 1u)) continue; }
            else slot = gbcast(sel_l, k);
            const unsigned r = sample(&views[job->gview[slot]]);
            const bool need_ncc = (candidates || want_ncc) && (r & 1u);
            float c0 = gbcast(cs0_l, k), c1 = gbcast(cs1_l, k), c2 = gbcast(cs2_l, k);
            const bool cs_now = !candidates && cs_active && (r & 1u);
            if (!candidates) {
                if (r & 1u) p_col_ok |= 1u << k;
                if (r & 2u) p_der_ok |= 1u << k;
                // computeColorScale: a failed view ends the whole update (`return`, not `continue`, patch_optimization.cc:92-93)
                if (cs_active && !(r & 1u)) cs_active = false;
            }
            float ncc = -1.f;
            if (need_ncc) {                               // getFastNCC (patch_sampler.cc:143-162)
                const float inv_n = 1.f / (float)NS;
                float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) { t0 += sn[s][0]; t1 += sn[s][1]; t2 += sn[s][2]; }
                const float my0 = gsum(t0) * inv_n, my1 = gsum(t1) * inv_n, my2 = gsum(t2) * inv_n;
                float sqrDevY = 0.f, devXY = 0.f;
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    if (!is_real(s)) continue;
                    const float y0 = sn[s][0] - my0, y1 = sn[s][1] - my1, y2 = sn[s][2] - my2;
                    sqrDevY += y0 * y0 + y1 * y1 + y2 * y2;
                    devXY += (mc[s][0] - mx0) * y0 + (mc[s][1] - mx1) * y1 + (mc[s][2] - mx2) * y2;
                }
                sqrDevY = gsum(sqrDevY); devXY = gsum(devXY);
                const float p = sqrDevX * sqrDevY;          // devXY / sqrt(p), -1 when sqrt(p) is not > 0
                ncc = p > 0.f ? devXY * rsqrt_fast(p) : -1.f;
            }
            if (cs_now) {                                 // computeColorScale for this view (patch_optimization.cc:88-110)
                float ab0 = 0.f, ab1 = 0.f, ab2 = 0.f, aa0 = 0.f, aa1 = 0.f, aa2 = 0.f;
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    if (!is_real(s)) continue;
                    ab0 += (mc[s][0] - sn[s][0] * c0) * sn[s][0]; ab1 += (mc[s][1] - sn[s][1] * c1) * sn[s][1]; ab2 += (mc[s][2] - sn[s][2] * c2) * sn[s][2];
                    aa0 += sn[s][0] * sn[s][0]; aa1 += sn[s][1] * sn[s][1]; aa2 += sn[s][2] * sn[s][2];
                }
                float cc[3] = {c0, c1, c2};
                const float ab[3] = {gsum(ab0), gsum(ab1), gsum(ab2)};
                const float aa[3] = {gsum(aa0), gsum(aa1), gsum(aa2)};
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    if ((double)fabsf(aa[ch]) > 1e-6) {
                        cc[ch] += ab[ch] * rcp_fast(aa[ch]);
                        if ((double)cc[ch] > 1e3) opti = false;
                    } else
                        opti = false;
                }
                c0 = cc[0]; c1 = cc[1]; c2 = cc[2];
                if (gl == k) { cs0_l = c0; cs1_l = c1; cs2_l = c2; }
            }
            // Gauss-Newton terms (patch_optimization.cc:283-288 / :324-343) with the colour scale of THIS state; only
            // meaningful when every view's derivative path succeeded, which the caller checks through p_der_ok
            if (!candidates && (r & 2u)) {
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    if (!is_real(s)) continue;
                    const float g0 = c0 * sd[s][0], g1 = c1 * sd[s][1], g2 = c2 * sd[s][2];
                    const float r0 = mc[s][0] - c0 * sn[s][0], r1 = mc[s][1] - c1 * sn[s][1], r2 = mc[s][2] - c2 * sn[s][2];
                    num += g0 * r0 + g1 * r1 + g2 * r2;
                    den += g0 * g0 + g1 * g1 + g2 * g2;
                    if (want_normal) {
                        const float gg[3] = {g0, g1, g2};
                        const float rr[3] = {r0, r1, r2};
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const float a0 = gg[ch];
                            const float a1 = fi[s] * a0;      // (ii * cs) * deriv == ii * (cs * deriv) exactly for ii in {-2..2}
                            const float a2 = fj[s] * a0;
                            A0 += a0 * a0; A1 += a0 * a1; A2 += a0 * a2;
                            A3 += a1 * a1; A4 += a1 * a2; A5 += a2 * a2;
                            B0 += a0 * rr[ch]; B1 += a1 * rr[ch]; B2 += a2 * rr[ch];
                        }
                    }
                }
            }
            if (candidates) {
                if (ncc < st->min_ncc) avail &= ~(1u << k);
                else if (gl == (k & (GROUP - 1))) {
                    const int j = k >> 3;
                    if (j == 0) cand0 = ncc; else if (j == 1) cand1 = ncc; else if (j == 2) cand2 = ncc; else cand3 = ncc;
                }
            } else if (want_ncc && gl == k)
                ncc_l = ncc;
        }
        if (candidates) return;
        p_num = gsum(num); p_den = gsum(den);
        p_has_normal = want_normal;
        p_has_ncc = want_ncc;
        if (want_normal) {
            // solve here so that the nine fp64 sums die with the pass (matrix_tools.h:392-398,460-475).  The lane's <= 48
            // products (3 channels x <= 4 samples x <= 4 views) are summed in fp32, the 8 lanes in fp64; the reference sums
            // all 300 fp32 products in fp64 (patch_optimization.cc:336-342).  The difference (~1e-7 relative on ATA) is far
            // below what the Gauss-Newton fixed point resolves; measured in tests/test_gpu_parity.py.
            const double D0 = gsum((double)A0), D1 = gsum((double)A1), D2 = gsum((double)A2), D3 = gsum((double)A3);
            const double D4 = gsum((double)A4), D5 = gsum((double)A5);
            const double E0 = gsum((double)B0), E1 = gsum((double)B1), E2 = gsum((double)B2);
            const double m[9] = {D0, D1, D2, D1, D3, D4, D2, D4, D5};
            const double det = m[0] * m[4] * m[8] + m[1] * m[5] * m[6] + m[2] * m[3] * m[7]
                             - m[2] * m[4] * m[6] - m[1] * m[3] * m[8] - m[0] * m[5] * m[7];
            n_singular = det == 0.0;
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
            nX0 = (float)(inv[0] * E0 + inv[1] * E1 + inv[2] * E2);
            nX1 = (float)(inv[3] * E0 + inv[4] * E1 + inv[5] * E2);
            nX2 = (float)(inv[6] * E0 + inv[7] * E1 + inv[8] * E2);
        }
    }

    __device__ __forceinline__ bool all_der_ok() const { return p_der_ok == ((1u << nsel) - 1u); }

    // optimizeDepthOnly (patch_optimization.cc:265-299) from the sums of the last pass. Returns true when the state moved.
    __device__ __forceinline__ bool depth_step()
    {
        if (!all_der_ok()) { opti = false; return false; }
        if (p_den > 0.f) {
            depth += p_num / p_den;
            update();
            opti = ref_ok;
            return true;
        }
        return false;
    }

    // optimizeDepthAndNormal (patch_optimization.cc:302-364) from the solution prepared by the last pass.
    __device__ __forceinline__ bool normal_step()
    {
        if (!all_der_ok()) { opti = false; return false; }
        if (n_singular) { opti = false; return false; }
        dzI += nX1; dzJ += nX2; depth += nX0;
        update();
        opti = ref_ok;
        return true;
    }

    // ---- sorted insert / erase on the lane-distributed selected set (std::set semantics) ----
    __device__ __forceinline__ void sel_erase_mask(unsigned mask)       // bit k: remove element k
    {
        // gather: new position p takes the p-th kept element
        int src = 0, cnt = 0;
        for (int k = 0; k < MAX_LOCAL; ++k) {
            const bool kk = k < nsel && !((mask >> k) & 1u);
            if (kk) { if (cnt == gl) src = k; ++cnt; }
        }
        const int s = gbcast(sel_l, src);
        const float a = gbcast(cs0_l, src), b = gbcast(cs1_l, src), c = gbcast(cs2_l, src);
        const float v = gbcast(ncc_l, src);
        nsel = cnt;
        if (gl < cnt) { sel_l = s; cs0_l = a; cs1_l = b; cs2_l = c; ncc_l = v; }
        else { sel_l = 0xFF; }
    }
    __device__ __forceinline__ void sel_insert(int slot, float cs_init)
    {
        // position = number of selected slots smaller than `slot`
        const unsigned smaller = gballot(gl < nsel && sel_l < slot);
        const int pos = __popc(smaller);
        const int s_up = __shfl_up_sync(gmask, sel_l, 1, GROUP);
        const float a_up = __shfl_up_sync(gmask, cs0_l, 1, GROUP), b_up = __shfl_up_sync(gmask, cs1_l, 1, GROUP), c_up = __shfl_up_sync(gmask, cs2_l, 1, GROUP);
        const float v_up = __shfl_up_sync(gmask, ncc_l, 1, GROUP);
        if (gl > pos && gl <= nsel) { sel_l = s_up; cs0_l = a_up; cs1_l = b_up; cs2_l = c_up; ncc_l = v_up; }
        if (gl == pos) { sel_l = slot; cs0_l = cs1_l = cs2_l = cs_init; ncc_l = 0.f; }
        ++nsel;
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

    // viewDir / epipolar plane / footprint of global slot `slot` at patchPoints[12] (local_view_selection.cc:93-131)
    __device__ __forceinline__ void cand_geometry(int slot, float rdx, float rdy, float rdz,
                                                  float& vdx, float& vdy, float& vdz, float& epx, float& epy, float& epz, float& nfp) const
    {
        const ViewParams* V = &views[job->gview[slot]];
        vdx = cpx - __ldg(&V->campos[0]); vdy = cpy - __ldg(&V->campos[1]); vdz = cpz - __ldg(&V->campos[2]);
        const float nn = sqrtf(vdx * vdx + vdy * vdy + vdz * vdz);
        vdx /= nn; vdy /= nn; vdz /= nn;
        epx = vdy * rdz - vdz * rdy; epy = vdz * rdx - vdx * rdz; epz = vdx * rdy - vdy * rdx;
        const float en = sqrtf(epx * epx + epy * epy + epz * epz);
        epx /= en; epy /= en; epz /= en;
        const float z = __ldg(&V->w2c[8]) * cpx + __ldg(&V->w2c[9]) * cpy + __ldg(&V->w2c[10]) * cpz + __ldg(&V->w2c[11]);
        nfp = z * __ldg(&V->inv_ax0);
    }

    // Second half of LocalViewSelection::performVS (local_view_selection.cc:86-147): greedy selection among the
    // candidates that survived the NCC test of pass(candidates = true); lane gl evaluates candidate slots gl + 8 j.
    __device__ __forceinline__ void lvs_greedy()
    {
        const unsigned N = st->nr_recon_neighbors;
        const float cs_init = 1.f / mm;
        float rdx = cpx - __ldg(&rv->campos[0]), rdy = cpy - __ldg(&rv->campos[1]), rdz = cpz - __ldg(&rv->campos[2]);
        {
            const float nn = sqrtf(rdx * rdx + rdy * rdy + rdz * rdz);
            rdx /= nn; rdy /= nn; rdz /= nn;
        }
        const int G = job->n_global;
        bool found = true;
        while ((unsigned)nsel < N && found) {
            found = false;
            float best_score = -1.f;
            int best_slot = 0x7FFFFFFF;
            // the already selected views, ascending (std::set iteration order); read before the lanes diverge
            const int ss0 = gbcast(sel_l, 0), ss1 = gbcast(sel_l, 1), ss2 = gbcast(sel_l, 2), ss3 = gbcast(sel_l, 3);
#pragma unroll 1
            for (int j = 0; j < MAX_GLOBAL / GROUP; ++j) {
                const int c = gl + GROUP * j;
                if (!(c < G && ((avail >> c) & 1u))) continue;
                float vdx, vdy, vdz, epx, epy, epz, nfp;
                cand_geometry(c, rdx, rdy, rdz, vdx, vdy, vdz, epx, epy, epz, nfp);
                float score = j == 0 ? cand0 : (j == 1 ? cand1 : (j == 2 ? cand2 : cand3));
                if (mfp / nfp < 0.5f) score *= 0.01f;
                float dp = clamp1(rdx * vdx + rdy * vdy + rdz * vdz);
                score *= plx_weight(deg_acos(dp));
                // parallax / epipolar terms against every already selected view
#pragma unroll 1
                for (int k = 0; k < nsel; ++k) {
                    const int s = k == 0 ? ss0 : (k == 1 ? ss1 : (k == 2 ? ss2 : ss3));
                    float sx, sy, sz, ex, ey, ez, sfp;
                    cand_geometry(s, rdx, rdy, rdz, sx, sy, sz, ex, ey, ez, sfp);
                    dp = clamp1(sx * vdx + sy * vdy + sz * vdz);
                    score *= plx_weight(deg_acos(dp));
                    dp = clamp1(epx * ex + epy * ey + epz * ez);
                    float angle = fabsf(deg_acos(dp));
                    if (angle > 90.f) angle = 180.f - angle;
                    angle = fmaxf(angle, 1.f);
                    if (angle < st->min_parallax) score *= angle / st->min_parallax;
                }
                // strict '>' in ascending slot order: the lowest slot wins ties (local_view_selection.cc:133-137); NaN never wins
                if (score > 0.f && score > best_score) { best_score = score; best_slot = c; }
            }
            const float best = gmax(best_score);
            const bool mine = best_score > 0.f && best_score == best;
            int w = mine ? best_slot : 0x7FFFFFFF;
#pragma unroll
            for (int o = GROUP / 2; o > 0; o >>= 1) w = min(w, __shfl_xor_sync(gmask, w, o));
            if (best > 0.f && w != 0x7FFFFFFF) {
                found = true;
                sel_insert(w, cs_init);
                avail &= ~(1u << w);
            }
        }
        if ((unsigned)nsel == N) lvs_ok = true;
    }

    // PatchOptimization ctor (patch_optimization.cc:21-78) incl. LocalViewSelection ctor (local_view_selection.cc:19-54),
    // up to the point where the first sample sets are needed; sets the first stage of the state machine.
    __device__ __forceinline__ void begin(const JobParams* j, const PatchIn& in)
    {
        job = j;
        rv = &views[job->ref_view];
        x0 = in.x; y0 = in.y;
        depth = in.depth; dzI = in.dzI; dzJ = in.dzJ;
        iter = 0; opti = true; converged = false; lvs_ok = false;
        nsel = 0; avail = 0u;
        sel_l = 0xFF; cs0_l = cs1_l = cs2_l = 0.f; ncc_l = 0.f; cand0 = cand1 = cand2 = cand3 = 0.f;
        p_col_ok = p_der_ok = 0u; p_num = p_den = 0.f; p_has_normal = p_has_ncc = false;
        nX0 = nX1 = nX2 = 0.f; n_singular = true;
        viewRemoved = was_normal = normal = false; old = 0.f;
        stage = DONE;
        init_sampler(in.x, in.y);
        // propagated ids arrive ascending, 0xFF padded
        if (gl < MAX_LOCAL) sel_l = (in.slots >> (8 * gl)) & 0xFF;
        nsel = __popc(gballot(gl < MAX_LOCAL && sel_l != 0xFF));
        if (!ref_ok) { opti = false; return; }
        const unsigned N = st->nr_recon_neighbors;
        if ((unsigned)nsel == N) lvs_ok = true;
        else if ((unsigned)nsel > N) { nsel = 0; sel_l = 0xFF; }
        avail = job->n_global >= 32 ? FULL : ((1u << job->n_global) - 1u);
        unsigned m = 0u;
        for (int k = 0; k < nsel; ++k) m |= 1u << gbcast(sel_l, k);
        avail &= ~m;
        cs0_l = cs1_l = cs2_l = 1.f / mm;
        stage = lvs_ok ? CTOR : LVS_CTOR;
    }

    // The rest of the ctor (performVS, computeColorScale) and PatchOptimization::doAutoOptimization
    // (patch_optimization.cc:66-77,170-242) as a state machine around the single pass() call site: one call = one pass
    // plus everything up to the next one.  Returns true when the optimisation is over.
    __device__ __forceinline__ bool step()
    {
        if (stage == DONE) return true;
        // arguments of the one pass() call, by stage
        const bool a_cand = (stage == LVS_CTOR) | (stage == LVS_REPL);
        const bool a_cs = (stage == CTOR) | (stage == REPL) | ((stage == POST) & was_normal);   // computeColorScale of :77, :230, :198
        const bool a_ncc = (stage == PRE) | (stage == POST) | (stage == REPL) | ((stage == FIRST) & (iter == 4));
        const bool a_normal = (stage == REPL) | ((stage == FIRST) & (iter == 4)) | ((stage == PRE) & normal) |
                              ((stage == POST) & ((iter + 1) % 5 == 4));
        pass(a_cand, a_cs, a_ncc, a_normal);
        if (stage == LVS_CTOR || stage == LVS_REPL) {
            lvs_greedy();
            if (!lvs_ok) { if (stage == LVS_CTOR) opti = false; stage = DONE; return true; }
            stage = (stage == LVS_CTOR) ? CTOR : REPL;
            return false;
        }
        if (!opti) { stage = DONE; return true; }        // a colour scale failed: every caller of computeColorScale gives up here
        if (stage == POST) {
            const float df = fabsf(ncc_l - old);
            const bool mine = gl < nsel;
            const bool conv = !gany(mine && df > st->min_refine_diff);
            const unsigned tbr = gballot(mine && (ncc_l < st->accept_ncc || (iter == 14 && df > st->min_refine_diff)));
            if (tbr) {
                viewRemoved = true;
                sel_erase_mask(tbr);          // LocalViewSelection::replaceViews (local_view_selection.cc:150-160)
                lvs_ok = false;
                stage = LVS_REPL;
                return false;
            }
            if (conv) { converged = true; stage = DONE; return true; }
            ++iter;
        } else if (stage == REPL) {
            ++iter;
        }
        // first four iterations only refine depth (:177-180)
        while (iter < 4 && opti) {
            const bool moved = depth_step();
            ++iter;
            if (moved && opti) { stage = FIRST; return false; }
        }
        if (!opti) { stage = DONE; return true; }
        // head of the main loop (:184-203)
        if (!((unsigned)iter < st->max_iterations && lvs_ok)) { stage = DONE; return true; }
        normal = (iter % 5 == 4) || viewRemoved;
        if (!p_has_ncc || (normal && !p_has_normal)) { stage = PRE; return false; }   // only after a depth step with denom <= 0
        old = ncc_l;                  // oldNCC (:190-193)
        opti = false;
        if (normal) { normal_step(); viewRemoved = false; was_normal = true; }
        else { depth_step(); was_normal = false; }
        if (!opti) { stage = DONE; return true; }
        stage = POST;
        return false;
    }

    // PatchOptimization::computeConfidence (patch_optimization.cc:114-142) + getPatchNormal (patch_sampler.cc:243-256)
    __device__ __forceinline__ void finish(PatchOut& out)
    {
        out.depth = depth; out.dzI = dzI; out.dzJ = dzJ;
        out.iterations = iter;
        out.flags = (converged ? 1 : 0) | (opti ? 2 : 0);
        unsigned s = 0u;
#pragma unroll
        for (int k = 0; k < MAX_LOCAL; ++k) {
            const int v = gbcast(sel_l, k);
            s |= (unsigned)((k < nsel) ? (v & 0xFF) : 0xFF) << (8 * k);
        }
        out.slots = s;
        out.conf = 0.f; out.nx = out.ny = out.nz = 0.f;
        if (!converged) return;
        // mean NCC of the final state (the NCCs of the last pass)
        float mean = 0.f;
        for (int k = 0; k < nsel; ++k) mean += gbcast(ncc_l, k);
        mean /= (float)nsel;
        const float score = (mean - st->accept_ncc) / (1.f - st->accept_ncc);
        // patchPoints[14] - patchPoints[10] and patchPoints[2] - patchPoints[22]: samples 10, 14 live in slot 2 (lanes 2, 6),
        // sample 2 in slot 1 (lane 2), sample 22 in slot 3 (lane 6)
        float q1x, q1y, q1z, q2x, q2y, q2z, q3x, q3y, q3z;
        {
            const float c0x = __ldg(&rv->campos[0]), c0y = __ldg(&rv->campos[1]), c0z = __ldg(&rv->campos[2]);
            float t = depth + fi[1] * dzI + fj[1] * dzJ;
            q1x = c0x + t * ray[1][0]; q1y = c0y + t * ray[1][1]; q1z = c0z + t * ray[1][2];
            t = depth + fi[2] * dzI + fj[2] * dzJ;
            q2x = c0x + t * ray[2][0]; q2y = c0y + t * ray[2][1]; q2z = c0z + t * ray[2][2];
            t = depth + fi[3] * dzI + fj[3] * dzJ;
            q3x = c0x + t * ray[3][0]; q3y = c0y + t * ray[3][1]; q3z = c0z + t * ray[3][2];
        }
        const float ax_ = gbcast(q2x, 6) - gbcast(q2x, 2);
        const float ay_ = gbcast(q2y, 6) - gbcast(q2y, 2);
        const float az_ = gbcast(q2z, 6) - gbcast(q2z, 2);
        const float bx_ = gbcast(q1x, 2) - gbcast(q3x, 6);
        const float by_ = gbcast(q1y, 2) - gbcast(q3y, 6);
        const float bz_ = gbcast(q1z, 2) - gbcast(q3z, 6);
        float nx = ay_ * bz_ - az_ * by_, ny = az_ * bx_ - ax_ * bz_, nz = ax_ * by_ - ay_ * bx_;
        const float nn = sqrtf(nx * nx + ny * ny + nz * nz);
        nx /= nn; ny /= nn; nz /= nn;
        out.nx = nx; out.ny = ny; out.nz = nz;
        const float dotP = -(nx * crx + ny * cry + nz * crz);
        out.conf = (dotP < 0.2f) ? 0.f : score;
    }
};

// Binds a thread to its group.  lut_rep: LUT_REP-fold replicated table in shared memory (lut_rep[v * LUT_REP + r] = srgb2lin[v]),
// tid: thread index in the block.
__device__ __forceinline__ void bind_thread(Patch& p, const DevSettings* st, const ViewParams* views, const float* lut_rep, int tid)
{
    const int lane = tid & 31;
    p.st = st; p.views = views;
    p.lutw = lut_rep + (lane & (LUT_REP - 1));
    p.gl = lane & (GROUP - 1);
    p.gmask = ((1u << GROUP) - 1u) << (lane & ~(GROUP - 1));
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int k = p.sample_index(s);
        const bool real = p.is_real(s);
        p.fi[s] = real ? (float)(k % 5 - 2) : 0.f;
        p.fj[s] = real ? (float)(k / 5 - 2) : 0.f;
    }
    p.stage = Patch::DONE;
    p.n_sets = 0u;
}

} // namespace b200mvs
