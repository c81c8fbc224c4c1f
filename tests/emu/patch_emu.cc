/* TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 * Runs the product's device code for one mvs::PatchOptimization (mve_b200/csrc/patch_opt.cuh, compiled by g++ with the
 * SIMT emulation of simt_emu.h) on the CPU.  Inputs are plain arrays prepared by tests/test_device_code_emulated.py.
 * Built by tests/emu/build.py into tests/emu/libpatch_emu.so. */
#define B200MVS_HOST_EMU 1
#include "simt_emu.h"

namespace simt_emu {
thread_local Warp* t_warp = nullptr;
thread_local int t_lane = 0;
thread_local int t_sense_full = 0;
thread_local int t_sense_sub = 0;
}

#include "../../mve_b200/csrc/patch_opt.cuh"
#include "../../mve_b200/csrc/patch_warp.cuh"
#include "../../mve_b200/csrc/patch_thread.cuh"

#include <atomic>
#include <thread>
#include <vector>

using namespace b200mvs;

extern "C" {

/* flat description of a view: 32 floats + per level 8 values, see the Python side */
struct EmuView {
    float campos[3], inv_ax0, w2c[12], rot[9];
    int nlevels;
    float ax[MAX_LEVELS], ay[MAX_LEVELS], cx[MAX_LEVELS], cy[MAX_LEVELS];
    int w[MAX_LEVELS], h[MAX_LEVELS], pitch[MAX_LEVELS];
    const uchar4* img[MAX_LEVELS];
    const uint4* quad[MAX_LEVELS];
};

struct EmuPatchIn { int x, y; float depth, dzI, dzJ; unsigned slots; };
struct EmuPatchOut { float conf, depth, dzI, dzJ, nx, ny, nz; unsigned slots; int iterations, flags; unsigned sets; };

int emu_struct_sizes(int* view, int* pin, int* pout)
{
    *view = (int)sizeof(EmuView); *pin = (int)sizeof(EmuPatchIn); *pout = (int)sizeof(EmuPatchOut);
    return MAX_LEVELS;
}

/* settings: min_ncc, min_parallax, accept_ncc, min_refine_diff, max_iterations, nr_recon_neighbors, scale, use_color_scale */
int emu_optimize_patches(const EmuView* views, int n_views, int ref_view, int W, int H, const float ki[4],
                         const int* gview, int n_global, const float* fsettings, const int* isettings, const float* lut,
                         const EmuPatchIn* in, int n, EmuPatchOut* out, int mode)
{
    std::vector<ViewParams> vp(n_views);
    std::memset(vp.data(), 0, vp.size() * sizeof(ViewParams));
    for (int v = 0; v < n_views; ++v) {
        std::memcpy(vp[v].campos, views[v].campos, 12);
        vp[v].inv_ax0 = views[v].inv_ax0;
        std::memcpy(vp[v].w2c, views[v].w2c, 48);
        std::memcpy(vp[v].rot, views[v].rot, 36);
        vp[v].nlevels = views[v].nlevels;
        vp[v].valid = 1;
        for (int l = 0; l < views[v].nlevels; ++l) {
            LevelParams& L = vp[v].lv[l];
            L.ax = views[v].ax[l]; L.ay = views[v].ay[l]; L.cx = views[v].cx[l]; L.cy = views[v].cy[l];
            L.w = views[v].w[l]; L.h = views[v].h[l]; L.pitch = views[v].pitch[l]; L.img = views[v].img[l]; L.quad = views[v].quad[l];
        }
    }
    DevSettings st;
    st.min_ncc = fsettings[0]; st.min_parallax = fsettings[1]; st.accept_ncc = fsettings[2]; st.min_refine_diff = fsettings[3];
    st.max_iterations = (unsigned)isettings[0]; st.nr_recon_neighbors = (unsigned)isettings[1];
    st.scale = isettings[2]; st.use_color_scale = isettings[3];
    JobParams job;
    std::memset(&job, 0, sizeof(job));
    job.ref_view = ref_view; job.W = W; job.H = H; job.n_global = n_global;
    for (int k = 0; k < n_global; ++k) job.gview[k] = gview[k];
    job.ki0 = ki[0]; job.ki2 = ki[1]; job.ki4 = ki[2]; job.ki5 = ki[3];
    job.ref_img = views[ref_view].img[st.scale];
    job.ref_pitch = views[ref_view].pitch[st.scale];

    if (mode == 2) {
        /* one THREAD per patch (patch_thread.cuh): no collectives, so the device code simply runs patch after patch */
        std::vector<float> lut_rep(256 * LUT_STRIDE);
        for (int i = 0; i < 256 * LUT_STRIDE; ++i) lut_rep[i] = lut[i / LUT_STRIDE];
        PatchT p;
        bind_thread(p, &st, vp.data(), lut_rep.data(), 0);
        for (int i = 0; i < n; ++i) {
            PatchIn pi;
            pi.x = in[i].x; pi.y = in[i].y; pi.depth = in[i].depth; pi.dzI = in[i].dzI; pi.dzJ = in[i].dzJ; pi.slots = in[i].slots;
            const unsigned before = p.n_sets;
            p.begin(&job, pi);
            while (!p.step()) {}
            PatchOut po;
            p.finish(po);
            out[i].conf = po.conf; out[i].depth = po.depth; out[i].dzI = po.dzI; out[i].dzJ = po.dzJ;
            out[i].nx = po.nx; out[i].ny = po.ny; out[i].nz = po.nz; out[i].slots = po.slots;
            out[i].iterations = po.iterations; out[i].flags = po.flags; out[i].sets = p.n_sets - before;
        }
        return 0;
    }
    /* one warp per patch (patch_warp.cuh): 32 host threads run the device code lane by lane; "shared memory" = the
     * replicated table */
    simt_emu::Warp warp;
    std::vector<float> lut_rep(256 * LUT_STRIDE);
    for (int i = 0; i < 256 * LUT_STRIDE; ++i) lut_rep[i] = lut[i / LUT_STRIDE];
    std::vector<std::thread> lanes;
    for (int lane = 0; lane < 32; ++lane) {
        lanes.emplace_back([&, lane]() {
            simt_emu::t_warp = &warp;
            simt_emu::t_lane = lane;
            simt_emu::t_sense_full = 0;
            simt_emu::t_sense_sub = 0;
            PatchW p;
            bind_thread(p, &st, vp.data(), lut_rep.data(), lane);
            for (int i = 0; i < n; ++i) {
                PatchIn pi;
                pi.x = in[i].x; pi.y = in[i].y; pi.depth = in[i].depth; pi.dzI = in[i].dzI; pi.dzJ = in[i].dzJ; pi.slots = in[i].slots;
                const unsigned before = p.n_sets;
                p.begin(&job, pi);
                while (!p.step()) {}
                PatchOut po;
                p.finish(po);
                if (lane == 0) {
                    out[i].conf = po.conf; out[i].depth = po.depth; out[i].dzI = po.dzI; out[i].dzJ = po.dzJ;
                    out[i].nx = po.nx; out[i].ny = po.ny; out[i].nz = po.nz; out[i].slots = po.slots;
                    out[i].iterations = po.iterations; out[i].flags = po.flags; out[i].sets = p.n_sets - before;
                }
            }
        });
    }
    for (std::thread& t : lanes) t.join();
    return 0;
}

} // extern "C"
