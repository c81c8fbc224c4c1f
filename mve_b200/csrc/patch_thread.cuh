// Device-side patch optimisation, THROUGHPUT variant: one THREAD per patch (32 patches per warp).
//
// Same function as PatchW in patch_warp.cuh - one mvs::PatchOptimization of the reference
// (libs/dmrecon/patch_optimization.cc:21-364 with PatchSampler patch_sampler.cc:19-393, LocalViewSelection
// local_view_selection.cc:19-160 and mvs_tools.cc:98-199) - organised for large frontier rounds, where there are far more
// queue entries than lanes on the chip:
//   * a lane walks the 25 samples of its own patch in a rolled loop, so no lane idles on a 25-of-32 mapping, nothing is
//     reduced across lanes (no shuffles at all) and everything that is uniform per patch - view constants, level choice,
//     colour scales, the state machine - is paid once per 32 patches instead of once per patch;
//   * nothing per-sample survives an iteration of the sample loop: the sums the reference forms in separate passes over
//     stored samples are accumulated in ONE sweep, in forms that do not need the means or the updated colour scale first:
//       NCC (patch_sampler.cc:143-162)      sum(n-p), sum((n-p)^2), sum(e (n-p)) around the pivot p = meanX * masterMeanCol
//                                           (the neighbour colours scatter around it, so the variance is formed without
//                                           cancellation)
//       colour scale (patch_optimization.cc:88-110)   sum(m n), sum(n n)  ->  ab = sum(m n) - cs * sum(n n)
//       Gauss-Newton, depth only, when the colour scale changes at this state (:283-288)
//                                           sum(d m), sum(d n), sum(d d) per channel -> numerator / denominator with the NEW scale
//       otherwise the products of :283-288 / :324-343 directly with the current scale;
//     only a view replacement (colour scale AND normal step at one state, rare) sweeps a view twice;
//   * geometry per sample is evaluated in the neighbour's camera frame directly: with u = R^T K^-1 (x, y, 1) the patch point
//     is C + t u / |u| and its image W (C + t u/|u|) + T = (W C + T) + (t / |u|) (W u); W C + T and W u0, W ua, W ub (u is
//     affine in the pixel offsets) are formed once per view and sweep - 12 FMAs per sample for BOTH projections of
//     patch_sampler.cc:94-133 instead of two full point transforms.  Same values up to rounding order (<= 2 ulp on pixel
//     coordinates; measured against the oracle in tests/test_gpu_parity.py and on the CPU in
//     tests/test_device_code_emulated.py).
// The state machine (stages, iteration counting, view replacement) is the one of PatchW::step() statement for statement.
#pragma once
#include "patch_opt.cuh"

namespace b200mvs {

struct PatchT {
    // ---- constants of the thread ----
    const DevSettings* st;
    const ViewParams* views;
#if defined(B200MVS_HOST_EMU)
    const float* lut_tab;      // lane-replicated srgb2lin table (lut_k); on the device: table() / lane_offset()
    unsigned lane4;            // 4 * (lane of this thread)
#endif
    // ---- constants of the patch ----
    const JobParams* job;
    unsigned xy;               // x | y << 16 of the patch centre
    float u0x, u0y, u0z;       // R^T K^-1 (x + .5, y + .5, 1): un-normalised ray of the centre pixel
    float uax, uay, uaz;       // R^T K^-1 (1, 0, 0): change of the ray per pixel in x
    float ubx, uby, ubz;       // R^T K^-1 (0, 1, 0): ... in y
    float c0x, c0y, c0z;       // camera centre of the reference view
    float mx0, mx1, mx2;       // meanX per channel (patch_sampler.cc:333-339)
    float crx, cry, crz;       // masterViewDirs[12]
    float cpx, cpy, cpz;       // patchPoints[12]
    float mfp, inv_mfp;        // footPrintScaled(patchPoints[12]) and its reciprocal
    float mm, inv_mm, sqrDevX; // masterMeanCol, its reciprocal, sqrDevX
    float depth, dzI, dzJ;
    unsigned avail;            // LocalViewSelection::available over global slots
    int iter;
    unsigned n_sets;
    // Everything small lives in ONE register (`pk`): the state of a patch stays live across the whole sample loop, and
    // every register it occupies there is a register the loop spills (profiles/r2_notes.md, "state packing").
    //   bits 0-9   flags F_*          bits 10-12 stage        bits 13-15 nsel (size of the selected set)
    //   bits 16-19 p_col_ok           bits 20-23 p_der_ok     (per selected view: colour / derivative samples valid in the last pass)
    //   bits 24-27 |NCC - oldNCC| > minRefineDiff per selected view, taken when the last pass overwrote the NCC (the
    //              reference keeps a copy, oldNCC, for this one comparison: patch_optimization.cc:196-199,217-225)
    unsigned pk;
    unsigned selp;             // the selected set: 4 x uint8 global slots, ascending, 0xFF = none
    // ---- per selected view (index = position in the ascending selected set) ----
    float cs[MAX_LOCAL][3];    // colorScale
    float ncc[MAX_LOCAL];      // NCC at the state of the last pass
    // (the sums a pass hands to the step that follows it - Gauss-Newton numerator / denominator, the solved normal
    //  equations - are consumed in the same step() call: PassOut, a local of step())


    struct PassOut { float num, den, nX0, nX1, nX2; };
    enum Stage { LVS_CTOR, CTOR, FIRST, PRE, POST, LVS_REPL, REPL, DONE };
    enum : unsigned {
        F_REF_OK = 1u << 0,        // sampler->success[refViewNr]
        F_OPTI = 1u << 1, F_CONVERGED = 1u << 2, F_LVS_OK = 1u << 3, F_VIEW_REMOVED = 1u << 4, F_WAS_NORMAL = 1u << 5,
        F_NORMAL = 1u << 6, F_SINGULAR = 1u << 7, F_HAS_NORMAL = 1u << 8, F_HAS_NCC = 1u << 9
    };
    static constexpr int SH_STAGE = 10, SH_NSEL = 13, SH_COL = 16, SH_DER = 20, SH_DF = 24;
    __device__ __forceinline__ bool is(unsigned f) const { return (pk & f) != 0u; }
    __device__ __forceinline__ void put(unsigned f, bool v) { pk = v ? (pk | f) : (pk & ~f); }
    __device__ __forceinline__ int stage() const { return (int)((pk >> SH_STAGE) & 7u); }
    __device__ __forceinline__ void set_stage(int v) { pk = (pk & ~(7u << SH_STAGE)) | ((unsigned)v << SH_STAGE); }
    __device__ __forceinline__ int nsel() const { return (int)((pk >> SH_NSEL) & 7u); }
    __device__ __forceinline__ void set_nsel(int v) { pk = (pk & ~(7u << SH_NSEL)) | ((unsigned)v << SH_NSEL); }
    __device__ __forceinline__ int sel(int k) const { return (int)((selp >> (8 * k)) & 0xFFu); }
    __device__ __forceinline__ int px() const { return (int)(xy & 0xFFFFu); }
    __device__ __forceinline__ int py() const { return (int)(xy >> 16); }

    // ---- register arrays with a dynamic index ----
    template <typename T> static __device__ __forceinline__ T get4(const T (&a)[MAX_LOCAL], int k)
    {
        return k == 0 ? a[0] : (k == 1 ? a[1] : (k == 2 ? a[2] : a[3]));
    }
    template <typename T> static __device__ __forceinline__ void set4(T (&a)[MAX_LOCAL], int k, T v)
    {
#pragma unroll
        for (int i = 0; i < MAX_LOCAL; ++i) if (i == k) a[i] = v;
    }

    // The table is the first thing in the kernels' dynamic shared memory (b200mvs.cu, OPT_SMEM_BYTES).  Read through the
    // member - the object itself lives in shared memory - the compiler would no longer know which address space it points
    // to and emit generic loads for the 15 look-ups of every sample.
    __device__ __forceinline__ unsigned lane_offset() const
    {
#if defined(B200MVS_HOST_EMU)
        return lane4;
#else
        return (threadIdx.x & 31u) << 2;
#endif
    }
    __device__ __forceinline__ const float* table() const
    {
#if defined(B200MVS_HOST_EMU)
        return lut_tab;
#else
        extern __shared__ float b200mvs_dyn_smem[];
        return b200mvs_dyn_smem;
#endif
    }

    // un-normalised ray of the sample with pixel offsets (di, dj), its reciprocal length and the sample's depth parameter
    __device__ __forceinline__ void sample_ray(float di, float dj, float& ux, float& uy, float& uz, float& inv) const
    {
        ux = u0x + di * uax + dj * ubx;
        uy = u0y + di * uay + dj * uby;
        uz = u0z + di * uaz + dj * ubz;
        inv = rsqrt_fast(ux * ux + uy * uy + uz * uz);
    }

    // patch_sampler.cc:274-295 (+ the centre point / master footprint used by every sample set)
    __device__ __forceinline__ void compute_points()
    {
        bool bad = false;
#pragma unroll 5
        for (int k = 0; k < NS; ++k) {
            const float t = depth + (float)(k % 5 - 2) * dzI + (float)(k / 5 - 2) * dzJ;
            bad |= t <= 0.f;
        }
        if (bad) pk &= ~F_REF_OK;
        cpx = c0x + depth * crx;
        cpy = c0y + depth * cry;
        cpz = c0z + depth * crz;
        const ViewParams* rv = &views[job->ref_view];
        const float z = __ldg(&rv->w2c[8]) * cpx + __ldg(&rv->w2c[9]) * cpy + __ldg(&rv->w2c[10]) * cpz + __ldg(&rv->w2c[11]);
        mfp = z * job->ki0;     // single_view.h:160-164
        inv_mfp = rcp_fast(mfp);
    }

    // PatchSampler ctor (patch_sampler.cc:19-62) + computeMasterSamples (:298-345)
    __device__ __forceinline__ void init_sampler(int x, int y)
    {
        const float* const lut_tab = table();
        const unsigned lane4 = lane_offset();
        pk &= ~F_REF_OK; mm = 0.f; inv_mm = 0.f; sqrDevX = 0.f;
        mx0 = mx1 = mx2 = 0.f;
        crx = cry = crz = cpx = cpy = cpz = mfp = inv_mfp = 0.f;
        if (x - 2 < 0 || y - 2 < 0 || x + 2 > job->W - 1 || y + 2 > job->H - 1) return;
        {
            // viewRayScaled (single_view.cc:99-106, depthmap.cc:149-156): K^-1 has the sparsity of camera.cc:180-200
            const ViewParams* rv = &views[job->ref_view];
            const float r0 = __ldg(&rv->rot[0]), r1 = __ldg(&rv->rot[1]), r2 = __ldg(&rv->rot[2]), r3 = __ldg(&rv->rot[3]), r4 = __ldg(&rv->rot[4]);
            const float r5 = __ldg(&rv->rot[5]), r6 = __ldg(&rv->rot[6]), r7 = __ldg(&rv->rot[7]), r8 = __ldg(&rv->rot[8]);
            const float vx = job->ki0 * ((float)x + 0.5f) + job->ki2, vy = job->ki4 * ((float)y + 0.5f) + job->ki5;
            u0x = r0 * vx + r3 * vy + r6; u0y = r1 * vx + r4 * vy + r7; u0z = r2 * vx + r5 * vy + r8;
            uax = r0 * job->ki0; uay = r1 * job->ki0; uaz = r2 * job->ki0;
            ubx = r3 * job->ki4; uby = r4 * job->ki4; ubz = r5 * job->ki4;
            c0x = __ldg(&rv->campos[0]); c0y = __ldg(&rv->campos[1]); c0z = __ldg(&rv->campos[2]);
            const float inv = rsqrt_fast(u0x * u0x + u0y * u0y + u0z * u0z);
            crx = u0x * inv; cry = u0y * inv; crz = u0z * inv;
        }
        pk |= F_REF_OK;
        // master colours: mean, then the per-channel means and deviations of the normalised colours
        const uchar4* row = job->ref_img + (size_t)(y - 2) * job->ref_pitch + (x - 2);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int j = 0; j < 5; ++j) {
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const unsigned t = reinterpret_cast<const unsigned*>(row)[i];
                s0 += lut_k<0>(lut_tab, lane4, t); s1 += lut_k<1>(lut_tab, lane4, t); s2 += lut_k<2>(lut_tab, lane4, t);
            }
            row += job->ref_pitch;
        }
        mm = (s0 + s1 + s2) / (3.f * NS);
        if (mm < 0.01f || mm > 0.99f) { pk &= ~F_REF_OK; return; }
        inv_mm = 1.f / mm;
        mx0 = (s0 * inv_mm) / (float)NS; mx1 = (s1 * inv_mm) / (float)NS; mx2 = (s2 * inv_mm) / (float)NS;
        row = job->ref_img + (size_t)(y - 2) * job->ref_pitch + (x - 2);
        float dev = 0.f;
#pragma unroll 1
        for (int j = 0; j < 5; ++j) {
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const unsigned t = reinterpret_cast<const unsigned*>(row)[i];
                const float e0 = lut_k<0>(lut_tab, lane4, t) * inv_mm - mx0, e1 = lut_k<1>(lut_tab, lane4, t) * inv_mm - mx1, e2 = lut_k<2>(lut_tab, lane4, t) * inv_mm - mx2;
                dev += e0 * e0 + e1 * e1 + e2 * e2;
            }
            row += job->ref_pitch;
        }
        sqrDevX = dev;
        compute_points();
    }

    // PatchSampler::update (patch_sampler.cc:259-271)
    __device__ __forceinline__ void update()
    {
        pk |= F_REF_OK;
        compute_points();
    }

    // One pass at the current state (same contract as PatchW::pass).
    // `cand`: NCC per candidate global slot, written by a candidates pass for the lvs_greedy() that follows it.
    __device__ __forceinline__ void pass(bool candidates, bool cs_pending, bool want_ncc, bool want_normal, float* cand, PassOut& po)
    {
        // The object lives in shared memory (k_frontier): what the sample loop reads 25 times per sweep is copied into
        // registers here, everything else is read where it is needed.
        const float* const lut_tab = table();
        const unsigned lane4 = lane_offset();
        const float u0x = this->u0x, u0y = this->u0y, u0z = this->u0z, uax = this->uax, uay = this->uay, uaz = this->uaz;
        const float ubx = this->ubx, uby = this->uby, ubz = this->ubz;
        const float depth = this->depth, dzI = this->dzI, dzJ = this->dzJ;
        const float mx0 = this->mx0, mx1 = this->mx1, mx2 = this->mx2, mm = this->mm, inv_mm = this->inv_mm;
        const JobParams* const job = this->job;
        float num = 0.f, den = 0.f;
        double D0 = 0.0, D1 = 0.0, D2 = 0.0, D3 = 0.0, D4 = 0.0, D5 = 0.0, E0 = 0.0, E1 = 0.0, E2 = 0.0;
        bool cs_active = cs_pending && st->use_color_scale;
        if (!candidates) pk &= want_ncc ? ~((0xFFu << SH_COL) | (0xFu << SH_DF)) : ~(0xFFu << SH_COL);     // p_col_ok = p_der_ok = 0
        const int count = candidates ? job->n_global : nsel();
        const int x0 = px(), y0 = py();
        const float pv0 = mx0 * mm, pv1 = mx1 * mm, pv2 = mx2 * mm;      // pivots of the NCC sums
#pragma unroll 1
        for (int k = 0; k < count; ++k) {
            int slot = k;
            if (candidates) { if (!((avail >> k) & 1u)) continue; }
            else slot = sel(k);
            const ViewParams* V = &views[job->gview[slot]];
            ++n_sets;
            float c0 = 1.f, c1 = 1.f, c2 = 1.f;
            if (!candidates) {
                c0 = k == 0 ? cs[0][0] : (k == 1 ? cs[1][0] : (k == 2 ? cs[2][0] : cs[3][0]));
                c1 = k == 0 ? cs[0][1] : (k == 1 ? cs[1][1] : (k == 2 ? cs[2][1] : cs[3][1]));
                c2 = k == 0 ? cs[0][2] : (k == 1 ? cs[1][2] : (k == 2 ? cs[2][2] : cs[3][2]));
            }
            // ---- view set-up: transform, level choice (patch_sampler.cc:76-91), derivative step (:94-100) ----
            unsigned r = 0u;
            float A0x = 0.f, A0y = 0.f, A0z = 0.f, W0x = 0.f, W0y = 0.f, W0z = 0.f, Wax = 0.f, Way = 0.f, Waz = 0.f, Wbx = 0.f, Wby = 0.f, Wbz = 0.f;
            float Lax = 0.f, Lay = 0.f, Lcx = 0.f, Lcy = 0.f, wm1 = 0.f, hm1 = 0.f, dd = 0.f, step = 0.f;
            int Lpitch = 0;
            const uint4* Lquad = nullptr;
            bool dok = false;
            {
                const float4 wa = __ldg(reinterpret_cast<const float4*>(&V->w2c[0]));
                const float4 wb = __ldg(reinterpret_cast<const float4*>(&V->w2c[4]));
                const float4 wc = __ldg(reinterpret_cast<const float4*>(&V->w2c[8]));
                A0x = wa.x * c0x + wa.y * c0y + wa.z * c0z + wa.w;
                A0y = wb.x * c0x + wb.y * c0y + wb.z * c0z + wb.w;
                A0z = wc.x * c0x + wc.y * c0y + wc.z * c0z + wc.w;
                W0x = wa.x * u0x + wa.y * u0y + wa.z * u0z; W0y = wb.x * u0x + wb.y * u0y + wb.z * u0z; W0z = wc.x * u0x + wc.y * u0y + wc.z * u0z;
                Wax = wa.x * uax + wa.y * uay + wa.z * uaz; Way = wb.x * uax + wb.y * uay + wb.z * uaz; Waz = wc.x * uax + wc.y * uay + wc.z * uaz;
                Wbx = wa.x * ubx + wa.y * uby + wa.z * ubz; Wby = wb.x * ubx + wb.y * uby + wb.z * ubz; Wbz = wc.x * ubx + wc.y * uby + wc.z * ubz;
                const float nz = wc.x * cpx + wc.y * cpy + wc.z * cpz + wc.w;
                const float nfp = nz * __ldg(&V->inv_ax0);
                // mfp <= 0 makes the reference throw std::out_of_range (patch_sampler.cc:78-82); it cannot happen for
                // depth > 0 because the centre ray has positive camera z.  Treated as a failed view here.
                if (mfp > 0.f && !(nfp <= 0.f)) {
                    float ratio = nfp * inv_mfp;
                    int l = 0;
                    while (ratio < 0.5f) { ++l; ratio *= 2.f; }
                    const int nl = __ldg(&V->nlevels);
                    if (l > nl - 1) l = nl - 1;                     // clampLevel, minLevel = 0 (single_view.h:113-123)
                    const float4 kk = __ldg(reinterpret_cast<const float4*>(&V->lv[l].ax));
                    const int4 g = __ldg(reinterpret_cast<const int4*>(&V->lv[l].w));
                    Lax = kk.x; Lay = kk.y; Lcx = kk.z; Lcy = kk.w; wm1 = (float)(g.x - 1); hm1 = (float)(g.y - 1); Lpitch = g.z;
                    Lquad = reinterpret_cast<const uint4*>(__ldg(reinterpret_cast<const unsigned long long*>(&V->lv[l].quad)));
                    // projections of patchPoints[12] and patchPoints[12] + masterViewDirs[12]
                    const float inv0 = rsqrt_fast(u0x * u0x + u0y * u0y + u0z * u0z);
                    const float sa = depth * inv0, sb = (depth + 1.f) * inv0;
                    const float hz1 = A0z + sa * W0z, hz2 = A0z + sb * W0z;
                    const float i1 = rcp_fast(hz1), i2 = rcp_fast(hz2);
                    const float ddx = (Lax * (A0x + sb * W0x) + Lcx * hz2) * i2 - (Lax * (A0x + sa * W0x) + Lcx * hz1) * i1;
                    const float ddy = (Lay * (A0y + sb * W0y) + Lcy * hz2) * i2 - (Lay * (A0y + sa * W0y) + Lcy * hz1) * i1;
                    const float dd2 = ddx * ddx + ddy * ddy;
                    dd = dd2 * rsqrt_fast(dd2);            // |.|; NaN for dd2 == 0, which fails `d > 0` like the reference's 0
                    dok = dd > 0.f;
                    step = rcp_fast(dd);
                    r = dok ? 3u : 1u;
                }
            }
            const bool need_ncc = candidates || want_ncc;
            const bool cs_view = !candidates && cs_active;         // a colour-scale update is due for this view (if it samples)
            float nccv = -1.f;
            // ---- sweeps over the 25 samples ----
#pragma unroll 1
            for (int rep = 0; rep < 2 && r != 0u; ++rep) {
                const bool second = rep == 1;
                if (second && !(cs_view && want_normal && (r & 2u))) break;
                const bool do_ncc = !second && need_ncc;
                const bool do_cs = !second && cs_view;
                // Gauss-Newton terms: directly when the colour scale of this state is known, through per-channel sums when it
                // is being updated and only the depth step follows, in the second sweep when the normal step follows
                const bool gn_direct = !candidates && (r & 2u) && (second || !cs_view);
                const bool gn_sums = !candidates && (r & 2u) && !second && cs_view && !want_normal;
                float S1a = 0.f, S1b = 0.f, S1c = 0.f, S2a = 0.f, S2b = 0.f, S2c = 0.f, Sen = 0.f;
                // A sweep either updates the colour scale (sums Mn, Nn and, when a depth step follows, Dm, Dn, Dd) or forms the
                // Gauss-Newton products directly (vnum, vden, A, B) - never both (do_cs excludes gn_direct), so the two groups
                // share their registers: 15 accumulators instead of 26 in the hottest loop of the kernel.
                float X0 = 0.f, X1 = 0.f, X2 = 0.f, X3 = 0.f, X4 = 0.f, X5 = 0.f, X6 = 0.f, X7 = 0.f, X8 = 0.f, X9 = 0.f, X10 = 0.f,
                      X11 = 0.f, X12 = 0.f, X13 = 0.f, X14 = 0.f;
                float &Mna = X0, &Mnb = X1, &Mnc = X2, &Nna = X3, &Nnb = X4, &Nnc = X5;
                float &Dma = X6, &Dmb = X7, &Dmc = X8, &Dna = X9, &Dnb = X10, &Dnc = X11, &Dda = X12, &Ddb = X13, &Ddc = X14;
                float &vnum = X0, &vden = X1;
                float &A0 = X2, &A1 = X3, &A2 = X4, &A3 = X5, &A4 = X6, &A5 = X7, &B0 = X8, &B1 = X9, &B2 = X10;
                bool oob = false;
                // The sample loop is software-pipelined: the geometry of sample k+1 is evaluated and its two loads (quad texel
                // of the neighbour, master texel) are issued BEFORE sample k is processed, so their latency is covered by the
                // ~150 instructions of table look-ups, interpolation and sums of sample k (an un-pipelined loop stalled on the
                // first use of each load, profiles/r2_notes.md).
                const uchar4* mptr = job->ref_img + (size_t)(y0 - 2) * job->ref_pitch + (x0 - 2);
                const int mskip = job->ref_pitch - 5;
                float ndi = -2.f, ndj = -2.f;              // offsets of the NEXT sample
                float nfx = 0.f, nfy = 0.f, ngx = 0.f, ngy = 0.f;
                uint4 nQ = make_uint4(0u, 0u, 0u, 0u);
                unsigned nmt = 0u;
                bool nvalid = false;
                // geometry + loads of the sample at (ndi, ndj)
#define B200MVS_STAGE_NEXT() do { \
                    const float ux_ = u0x + ndi * uax + ndj * ubx, uy_ = u0y + ndi * uay + ndj * uby, uz_ = u0z + ndi * uaz + ndj * ubz; \
                    const float inv_ = rsqrt_fast(ux_ * ux_ + uy_ * uy_ + uz_ * uz_); \
                    const float t_ = depth + ndi * dzI + ndj * dzJ; \
                    const float s1_ = t_ * inv_; \
                    const float wx_ = W0x + ndi * Wax + ndj * Wbx, wy_ = W0y + ndi * Way + ndj * Wby, wz_ = W0z + ndi * Waz + ndj * Wbz; \
                    const float hx_ = A0x + s1_ * wx_, hy_ = A0y + s1_ * wy_, hz_ = A0z + s1_ * wz_; \
                    const float ih_ = rcp_fast(hz_); \
                    const float qx_ = (Lax * hx_ + Lcx * hz_) * ih_ - 0.5f; \
                    const float qy_ = (Lay * hy_ + Lcy * hz_) * ih_ - 0.5f; \
                    nvalid = qx_ > 0.f && qx_ < wm1 && qy_ > 0.f && qy_ < hm1; \
                    if (nvalid) { \
                        ngx = 0.f; ngy = 0.f; \
                        if (dok) { \
                            const float s2_ = s1_ + step * inv_; \
                            const float kx_ = A0x + s2_ * wx_, ky_ = A0y + s2_ * wy_, kz_ = A0z + s2_ * wz_; \
                            const float ik_ = rcp_fast(kz_); \
                            ngx = (Lax * kx_ + Lcx * kz_) * ik_ - 0.5f - qx_; \
                            ngy = (Lay * ky_ + Lcy * kz_) * ik_ - 0.5f - qy_; \
                        } \
                        const int left_ = (int)floorf(qx_), top_ = (int)floorf(qy_); \
                        nfx = qx_ - (float)left_; nfy = qy_ - (float)top_; \
                        noff = (unsigned)top_ * (unsigned)Lpitch + (unsigned)left_; \
                        nQ = __ldg(Lquad + noff); \
                        nmt = *reinterpret_cast<const unsigned*>(mptr); \
                    } \
                } while (0)
                unsigned noff = 0u;
                B200MVS_STAGE_NEXT();
#ifdef B200MVS_T1_UNROLL2
#pragma unroll 2
#else
#pragma unroll 1
#endif
                for (int k = 0; k < NS; ++k) {
                    if (!nvalid) { oob = true; break; }
                    // the staged sample becomes the current one
                    const float di = ndi, dj = ndj, fx = nfx, fy = nfy, gx = ngx, gy = ngy;
                    const uint4 Q = nQ;
                    const unsigned mt = nmt;
                    if (k + 1 < NS) {
                        ndi += 1.f; ++mptr;
                        if (ndi > 2.f) { ndi = -2.f; ndj += 1.f; mptr += mskip; }
                        B200MVS_STAGE_NEXT();
                    }
                    {
                        const float m[3] = {lut_k<0>(lut_tab, lane4, mt) * inv_mm, lut_k<1>(lut_tab, lane4, mt) * inv_mm, lut_k<2>(lut_tab, lane4, mt) * inv_mm};
                        float a[3], b[3], c[3], e[3];
                        a[0] = lut_k<0>(lut_tab, lane4, Q.x); a[1] = lut_k<1>(lut_tab, lane4, Q.x); a[2] = lut_k<2>(lut_tab, lane4, Q.x);
                        b[0] = lut_k<0>(lut_tab, lane4, Q.y); b[1] = lut_k<1>(lut_tab, lane4, Q.y); b[2] = lut_k<2>(lut_tab, lane4, Q.y);
                        c[0] = lut_k<0>(lut_tab, lane4, Q.z); c[1] = lut_k<1>(lut_tab, lane4, Q.z); c[2] = lut_k<2>(lut_tab, lane4, Q.z);
                        e[0] = lut_k<0>(lut_tab, lane4, Q.w); e[1] = lut_k<1>(lut_tab, lane4, Q.w); e[2] = lut_k<2>(lut_tab, lane4, Q.w);
                        float n[3], d[3];
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const float x0_ = (1.f - fx) * a[ch] + fx * b[ch];
                            const float x3_ = (1.f - fx) * c[ch] + fx * e[ch];
                            n[ch] = (1.f - fy) * x0_ + fy * x3_;
                            const float der = gx * (b[ch] - a[ch]) + gy * (c[ch] - a[ch]) + (gy * fx + gx * fy) * (a[ch] - b[ch] - c[ch] + e[ch]);
                            d[ch] = dok ? der * dd : 0.f;        // deriv /= stepSize with stepSize = 1 / d (patch_sampler.cc:100,129-130)
                        }
                        if (do_ncc) {
                            const float y0_ = n[0] - pv0, y1_ = n[1] - pv1, y2_ = n[2] - pv2;
                            S1a += y0_; S1b += y1_; S1c += y2_;
                            S2a += y0_ * y0_; S2b += y1_ * y1_; S2c += y2_ * y2_;
                            Sen += (m[0] - mx0) * y0_ + (m[1] - mx1) * y1_ + (m[2] - mx2) * y2_;
                        }
                        if (do_cs) {
                            Mna += m[0] * n[0]; Mnb += m[1] * n[1]; Mnc += m[2] * n[2];
                            Nna += n[0] * n[0]; Nnb += n[1] * n[1]; Nnc += n[2] * n[2];
                        }
                        if (gn_sums) {
                            Dma += d[0] * m[0]; Dmb += d[1] * m[1]; Dmc += d[2] * m[2];
                            Dna += d[0] * n[0]; Dnb += d[1] * n[1]; Dnc += d[2] * n[2];
                            Dda += d[0] * d[0]; Ddb += d[1] * d[1]; Ddc += d[2] * d[2];
                        }
                        if (gn_direct) {
                            // patch_optimization.cc:283-288 / :324-343
                            const float g0 = c0 * d[0], g1 = c1 * d[1], g2 = c2 * d[2];
                            const float r0 = m[0] - c0 * n[0], r1 = m[1] - c1 * n[1], r2 = m[2] - c2 * n[2];
                            vnum += g0 * r0 + g1 * r1 + g2 * r2;
                            vden += g0 * g0 + g1 * g1 + g2 * g2;
                            if (want_normal) {
                                const float gg[3] = {g0, g1, g2};
                                const float rr[3] = {r0, r1, r2};
#pragma unroll
                                for (int ch = 0; ch < 3; ++ch) {
                                    const float a0 = gg[ch];
                                    const float a1 = di * a0;      // (ii * cs) * deriv == ii * (cs * deriv) exactly for ii in {-2..2}
                                    const float a2 = dj * a0;
                                    A0 += a0 * a0; A1 += a0 * a1; A2 += a0 * a2;
                                    A3 += a1 * a1; A4 += a1 * a2; A5 += a2 * a2;
                                    B0 += a0 * rr[ch]; B1 += a1 * rr[ch]; B2 += a2 * rr[ch];
                                }
                            }
                        }
                    }
                }
#undef B200MVS_STAGE_NEXT
                if (oob) { r = 0u; break; }
                if (do_ncc) {                             // getFastNCC (patch_sampler.cc:143-162)
                    const float inv_n = 1.f / (float)NS;
                    const float sqrDevY = (S2a - S1a * S1a * inv_n) + (S2b - S1b * S1b * inv_n) + (S2c - S1c * S1c * inv_n);
                    const float p = sqrDevX * sqrDevY;      // devXY / sqrt(p), -1 when sqrt(p) is not > 0
                    nccv = p > 0.f ? Sen * rsqrt_fast(p) : -1.f;
                }
                if (do_cs) {                              // computeColorScale for this view (patch_optimization.cc:88-110)
                    float cc[3] = {c0, c1, c2};
                    const float ab[3] = {Mna - c0 * Nna, Mnb - c1 * Nnb, Mnc - c2 * Nnc};
                    const float aa[3] = {Nna, Nnb, Nnc};
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        if ((double)fabsf(aa[ch]) > 1e-6) {
                            cc[ch] += ab[ch] * rcp_fast(aa[ch]);
                            if ((double)cc[ch] > 1e3) pk &= ~F_OPTI;
                        } else
                            pk &= ~F_OPTI;
                    }
                    c0 = cc[0]; c1 = cc[1]; c2 = cc[2];
#pragma unroll
                    for (int i = 0; i < MAX_LOCAL; ++i) if (i == k) { cs[i][0] = c0; cs[i][1] = c1; cs[i][2] = c2; }
                }
                if (gn_sums) {
                    const float vn = c0 * (Dma - c0 * Dna) + c1 * (Dmb - c1 * Dnb) + c2 * (Dmc - c2 * Dnc);
                    const float vd = c0 * c0 * Dda + c1 * c1 * Ddb + c2 * c2 * Ddc;
                    vnum = vn; vden = vd;                  // (Mn, Nn, whose registers these are, have been consumed above)
                }
                if (gn_direct || gn_sums) {
                    num += vnum; den += vden;
                    if (want_normal && gn_direct) {
                        // a view's <= 75 products are summed in fp32, the views in fp64; the reference sums all 300 fp32 products
                        // in fp64 (patch_optimization.cc:336-342) - far below what the Gauss-Newton fixed point resolves
                        D0 += (double)A0; D1 += (double)A1; D2 += (double)A2; D3 += (double)A3; D4 += (double)A4; D5 += (double)A5;
                        E0 += (double)B0; E1 += (double)B1; E2 += (double)B2;
                    }
                }
            }
            if (!candidates) {
                if (r & 1u) pk |= 1u << (SH_COL + k);
                if (r & 2u) pk |= 1u << (SH_DER + k);
                // computeColorScale: a failed view ends the whole update (`return`, not `continue`, patch_optimization.cc:92-93)
                if (cs_active && !(r & 1u)) cs_active = false;
                if (want_ncc) {
                    const float v = (r & 1u) ? nccv : -1.f;
                    if (fabsf(v - get4(ncc, k)) > st->min_refine_diff) pk |= 1u << (SH_DF + k);
                    set4(ncc, k, v);
                }
            } else {
                const float v = (r & 1u) ? nccv : -1.f;
                if (v < st->min_ncc) avail &= ~(1u << k);
                else cand[k] = v;
            }
        }
        if (candidates) return;
        po.num = num; po.den = den;
        put(F_HAS_NORMAL, want_normal);
        put(F_HAS_NCC, want_ncc);
        if (want_normal) {
            // matrix_tools.h:392-398,460-475
            const double m[9] = {D0, D1, D2, D1, D3, D4, D2, D4, D5};
            const double det = m[0] * m[4] * m[8] + m[1] * m[5] * m[6] + m[2] * m[3] * m[7]
                             - m[2] * m[4] * m[6] - m[1] * m[3] * m[8] - m[0] * m[5] * m[7];
            put(F_SINGULAR, det == 0.0);
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
            po.nX0 = (float)(inv[0] * E0 + inv[1] * E1 + inv[2] * E2);
            po.nX1 = (float)(inv[3] * E0 + inv[4] * E1 + inv[5] * E2);
            po.nX2 = (float)(inv[6] * E0 + inv[7] * E1 + inv[8] * E2);
        }
    }

    __device__ __forceinline__ bool all_der_ok() const { return ((pk >> SH_DER) & 0xFu) == ((1u << nsel()) - 1u); }

    // optimizeDepthOnly (patch_optimization.cc:265-299) from the sums of the last pass. Returns true when the state moved.
    __device__ __forceinline__ bool depth_step(const PassOut& po)
    {
        if (!all_der_ok()) { pk &= ~F_OPTI; return false; }
        if (po.den > 0.f) {
            depth += po.num / po.den;
            update();
            put(F_OPTI, is(F_REF_OK));
            return true;
        }
        return false;
    }

    // optimizeDepthAndNormal (patch_optimization.cc:302-364) from the solution prepared by the last pass.
    __device__ __forceinline__ bool normal_step(const PassOut& po)
    {
        if (!all_der_ok()) { pk &= ~F_OPTI; return false; }
        if (is(F_SINGULAR)) { pk &= ~F_OPTI; return false; }
        dzI += po.nX1; dzJ += po.nX2; depth += po.nX0;
        update();
        put(F_OPTI, is(F_REF_OK));
        return true;
    }

    // ---- sorted insert / erase on the selected set (std::set semantics) ----
    __device__ __forceinline__ void sel_erase_mask(unsigned mask)       // bit k: remove element k
    {
        const int n = nsel();
        int cnt = 0;
        unsigned packed = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < MAX_LOCAL; ++k) {
            if (k < n && !((mask >> k) & 1u)) {
                // element k moves to position cnt (cnt <= k)
                const unsigned s8 = (selp >> (8 * k)) & 0xFFu;
                const float a = cs[k][0], b = cs[k][1], c = cs[k][2], v = ncc[k];
                packed = (packed & ~(0xFFu << (8 * cnt))) | (s8 << (8 * cnt));
#pragma unroll
                for (int q = 0; q < MAX_LOCAL; ++q) if (q == cnt) { cs[q][0] = a; cs[q][1] = b; cs[q][2] = c; ncc[q] = v; }
                ++cnt;
            }
        }
        selp = packed;
        set_nsel(cnt);
    }
    __device__ __forceinline__ void sel_insert(int slot, float cs_init)
    {
        const int n = nsel();
        int pos = 0;
#pragma unroll
        for (int k = 0; k < MAX_LOCAL; ++k) if (k < n && sel(k) < slot) ++pos;
#pragma unroll
        for (int q = MAX_LOCAL - 1; q > 0; --q)
            if (q > pos && q <= n) { cs[q][0] = cs[q - 1][0]; cs[q][1] = cs[q - 1][1]; cs[q][2] = cs[q - 1][2]; ncc[q] = ncc[q - 1]; }
#pragma unroll
        for (int q = 0; q < MAX_LOCAL; ++q) if (q == pos) { cs[q][0] = cs[q][1] = cs[q][2] = cs_init; ncc[q] = 0.f; }
        // bytes below `pos` stay, the byte at `pos` becomes the new slot, the bytes above move up by one
        const unsigned low = pos == 0 ? 0u : (selp & (0xFFFFFFFFu >> (32 - 8 * pos)));
        const unsigned high = pos >= 3 ? 0u : ((selp >> (8 * pos)) << (8 * (pos + 1)));
        selp = low | ((unsigned)slot << (8 * pos)) | high;
        set_nsel(n + 1);
    }

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

    // Second half of LocalViewSelection::performVS (local_view_selection.cc:86-147)
    __device__ __forceinline__ void lvs_greedy(const float* cand)
    {
        const unsigned N = st->nr_recon_neighbors;
        const float cs_init = 1.f / mm;
        float rdx = cpx - c0x, rdy = cpy - c0y, rdz = cpz - c0z;
        {
            const float nn = sqrtf(rdx * rdx + rdy * rdy + rdz * rdz);
            rdx /= nn; rdy /= nn; rdz /= nn;
        }
        const int G = job->n_global;
        bool found = true;
        while ((unsigned)nsel() < N && found) {
            found = false;
            float maxScore = 0.f;
            int maxView = 0;
#pragma unroll 1
            for (int c = 0; c < G; ++c) {
                if (!((avail >> c) & 1u)) continue;
                float vdx, vdy, vdz, epx, epy, epz, nfp;
                cand_geometry(c, rdx, rdy, rdz, vdx, vdy, vdz, epx, epy, epz, nfp);
                float score = cand[c];
                if (mfp / nfp < 0.5f) score *= 0.01f;
                float dp = clamp1(rdx * vdx + rdy * vdy + rdz * vdz);
                score *= plx_weight(deg_acos(dp));
                const int ns = nsel();
#pragma unroll 1
                for (int k = 0; k < ns; ++k) {
                    float sx, sy, sz, ex, ey, ez, sfp;
                    cand_geometry(sel(k), rdx, rdy, rdz, sx, sy, sz, ex, ey, ez, sfp);
                    dp = clamp1(sx * vdx + sy * vdy + sz * vdz);
                    score *= plx_weight(deg_acos(dp));
                    dp = clamp1(epx * ex + epy * ey + epz * ez);
                    float angle = fabsf(deg_acos(dp));
                    if (angle > 90.f) angle = 180.f - angle;
                    angle = fmaxf(angle, 1.f);
                    if (angle < st->min_parallax) score *= angle / st->min_parallax;
                }
                if (score > maxScore) { maxScore = score; maxView = c; found = true; }     // local_view_selection.cc:133-137
            }
            if (found) {
                sel_insert(maxView, cs_init);
                avail &= ~(1u << maxView);
            }
        }
        if ((unsigned)nsel() == N) pk |= F_LVS_OK;
    }

    // PatchOptimization ctor (patch_optimization.cc:21-78) incl. LocalViewSelection ctor (local_view_selection.cc:19-54)
    __device__ __forceinline__ void begin(const JobParams* j, const PatchIn& in)
    {
        job = j;
        xy = ((unsigned)in.x & 0xFFFFu) | ((unsigned)in.y << 16);
        depth = in.depth; dzI = in.dzI; dzJ = in.dzJ;
        iter = 0;
        pk = F_OPTI | F_SINGULAR | ((unsigned)DONE << SH_STAGE);      // opti = true, n_singular = true, everything else clear
        avail = 0u;
        u0x = u0y = u0z = uax = uay = uaz = ubx = uby = ubz = c0x = c0y = c0z = 0.f;
        init_sampler(in.x, in.y);
        selp = in.slots;                                 // propagated ids arrive ascending, 0xFF padded
        int n = 0;
#pragma unroll
        for (int k = 0; k < MAX_LOCAL; ++k) {
            if (sel(k) != 0xFF) n = k + 1;
            cs[k][0] = cs[k][1] = cs[k][2] = 0.f; ncc[k] = 0.f;
        }
        set_nsel(n);
        if (!is(F_REF_OK)) { pk &= ~F_OPTI; return; }
        const unsigned N = st->nr_recon_neighbors;
        if ((unsigned)n == N) pk |= F_LVS_OK;
        else if ((unsigned)n > N) {
            n = 0;
            set_nsel(0);
            selp = 0xFFFFFFFFu;
        }
        avail = job->n_global >= 32 ? FULL : ((1u << job->n_global) - 1u);
#pragma unroll
        for (int k = 0; k < MAX_LOCAL; ++k) if (k < n) avail &= ~(1u << sel(k));
        const float ci = 1.f / mm;
#pragma unroll
        for (int k = 0; k < MAX_LOCAL; ++k) { cs[k][0] = cs[k][1] = cs[k][2] = ci; }
        set_stage(is(F_LVS_OK) ? CTOR : LVS_CTOR);
    }

    // doAutoOptimization as a state machine around the single pass() call site (see PatchW::step)
    __device__ __forceinline__ bool step()
    {
        int stage = this->stage();
        if (stage == DONE) return true;
        const bool a_cand = (stage == LVS_CTOR) | (stage == LVS_REPL);
        const bool a_cs = (stage == CTOR) | (stage == REPL) | ((stage == POST) & is(F_WAS_NORMAL));
        const bool a_ncc = (stage == PRE) | (stage == POST) | (stage == REPL) | ((stage == FIRST) & (iter == 4));
        const bool a_normal = (stage == REPL) | ((stage == FIRST) & (iter == 4)) | ((stage == PRE) & is(F_NORMAL)) |
                              ((stage == POST) & ((iter + 1) % 5 == 4));
        float cand[MAX_GLOBAL];      // candidates' NCCs: local memory, touched on the (rare) view-selection passes only
        PassOut po = {0.f, 0.f, 0.f, 0.f, 0.f};
        pass(a_cand, a_cs, a_ncc, a_normal, cand, po);
        if (stage == LVS_CTOR || stage == LVS_REPL) {
            lvs_greedy(cand);
            if (!is(F_LVS_OK)) { if (stage == LVS_CTOR) pk &= ~F_OPTI; set_stage(DONE); return true; }
            set_stage((stage == LVS_CTOR) ? CTOR : REPL);
            return false;
        }
        if (!is(F_OPTI)) { set_stage(DONE); return true; }
        if (stage == POST) {
            // `abs(ncc - oldNCC) > minRefineDiff` per view was taken when the pass overwrote the NCC (pass(), want_ncc)
            const int n = nsel();
            const unsigned dfb = (pk >> SH_DF) & ((1u << n) - 1u);
            unsigned tbr = (iter == 14) ? dfb : 0u;
#pragma unroll
            for (int k = 0; k < MAX_LOCAL; ++k)
                if (k < n && ncc[k] < st->accept_ncc) tbr |= 1u << k;
            if (tbr) {
                pk |= F_VIEW_REMOVED;
                sel_erase_mask(tbr);          // LocalViewSelection::replaceViews (local_view_selection.cc:150-160)
                pk &= ~F_LVS_OK;
                set_stage(LVS_REPL);
                return false;
            }
            if (dfb == 0u) { pk |= F_CONVERGED; set_stage(DONE); return true; }
            ++iter;
        } else if (stage == REPL) {
            ++iter;
        }
        while (iter < 4 && is(F_OPTI)) {
            const bool moved = depth_step(po);
            ++iter;
            if (moved && is(F_OPTI)) { set_stage(FIRST); return false; }
        }
        if (!is(F_OPTI)) { set_stage(DONE); return true; }
        if (!((unsigned)iter < st->max_iterations && is(F_LVS_OK))) { set_stage(DONE); return true; }
        const bool normal = (iter % 5 == 4) || is(F_VIEW_REMOVED);
        put(F_NORMAL, normal);
        if (!is(F_HAS_NCC) || (normal && !is(F_HAS_NORMAL))) { set_stage(PRE); return false; }
        // (oldNCC = ncc here in the reference: the comparison is taken when the next pass overwrites the NCC)
        pk &= ~F_OPTI;
        if (normal) { normal_step(po); pk &= ~F_VIEW_REMOVED; pk |= F_WAS_NORMAL; }
        else { depth_step(po); pk &= ~F_WAS_NORMAL; }
        if (!is(F_OPTI)) { set_stage(DONE); return true; }
        set_stage(POST);
        return false;
    }

    // PatchOptimization::computeConfidence (patch_optimization.cc:114-142) + getPatchNormal (patch_sampler.cc:243-256)
    __device__ __forceinline__ void finish(PatchOut& out)
    {
        out.depth = depth; out.dzI = dzI; out.dzJ = dzJ;
        out.iterations = iter;
        out.flags = (is(F_CONVERGED) ? 1 : 0) | (is(F_OPTI) ? 2 : 0);
        const int n = nsel();
        out.slots = n >= 4 ? selp : (selp | (0xFFFFFFFFu << (8 * n)));
        out.conf = 0.f; out.nx = out.ny = out.nz = 0.f;
        if (!is(F_CONVERGED)) return;
        float mean = 0.f;
#pragma unroll
        for (int k = 0; k < MAX_LOCAL; ++k) if (k < n) mean += ncc[k];
        mean /= (float)n;
        const float score = (mean - st->accept_ncc) / (1.f - st->accept_ncc);
        // patchPoints[14] - patchPoints[10] and patchPoints[2] - patchPoints[22]
        float px[4], py[4], pz[4];
        const float di_[4] = {2.f, -2.f, 0.f, 0.f}, dj_[4] = {0.f, 0.f, -2.f, 2.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float ux, uy, uz, inv;
            sample_ray(di_[q], dj_[q], ux, uy, uz, inv);
            const float s1 = (depth + di_[q] * dzI + dj_[q] * dzJ) * inv;
            px[q] = c0x + s1 * ux; py[q] = c0y + s1 * uy; pz[q] = c0z + s1 * uz;
        }
        const float ax_ = px[0] - px[1], ay_ = py[0] - py[1], az_ = pz[0] - pz[1];
        const float bx_ = px[2] - px[3], by_ = py[2] - py[3], bz_ = pz[2] - pz[3];
        float nx = ay_ * bz_ - az_ * by_, ny = az_ * bx_ - ax_ * bz_, nz = ax_ * by_ - ay_ * bx_;
        const float nn = sqrtf(nx * nx + ny * ny + nz * nz);
        nx /= nn; ny /= nn; nz /= nn;
        out.nx = nx; out.ny = ny; out.nz = nz;
        const float dotP = -(nx * crx + ny * cry + nz * crz);
        out.conf = (dotP < 0.2f) ? 0.f : score;
    }
};

#if !defined(B200MVS_HOST_EMU)
// PatchT objects sit side by side in shared memory (k_frontier).  With 8-byte members the stride cannot be an odd number of
// words; 2 mod 4 words keeps the conflicts of a warp's accesses to one member at 2-way (these are the once-per-sweep
// accesses, not the table look-ups of the sample loop).
static_assert(sizeof(PatchT) % 16 == 8 && sizeof(PatchT) <= 248, "PatchT stride in shared memory must be 8 mod 16 bytes, and 512 of them must fit beside the table");
#endif

__device__ __forceinline__ void bind_thread(PatchT& p, const DevSettings* st, const ViewParams* views, const float* lut_rep, int tid)
{
    p.st = st; p.views = views;
#if defined(B200MVS_HOST_EMU)
    p.lut_tab = lut_rep; p.lane4 = 4u * (unsigned)(tid & (LUT_REP - 1));
#else
    (void)lut_rep; (void)tid;      // the table is the start of the dynamic shared memory, the lane comes from threadIdx
#endif
    p.pk = (unsigned)PatchT::DONE << PatchT::SH_STAGE;
    p.n_sets = 0u;
}

} // namespace b200mvs
