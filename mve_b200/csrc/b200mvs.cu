// libb200mvs.so - C ABI of include/b200mvs.h: context, image pyramids, host-side view selection and seeds,
// frontier (region growing) orchestration and all CUDA kernels.  sm_100a only; no CPU fallback.
#include "../../include/b200mvs.h"
#include "patch_opt.cuh"
#include "patch_warp.cuh"
#include "patch_thread.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <chrono>
#include <ctime>
#include <mutex>
#include <string>
#include <thread>
#include <atomic>
#include <vector>

using namespace b200mvs;

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_create_error;      // last b200mvs_create error of the calling thread

struct HostLevel {
    int w = 0, h = 0, pitch = 0;
    float proj[9], invproj[9];
    uchar4* d_img = nullptr;
    uint4* d_quad = nullptr;       // 2x2 neighbourhood of every texel, 16 bytes (patch_opt.cuh LevelParams::quad)
};

struct HostView {
    bool valid = false;            // a SingleView exists (camera + image dimensions known)
    bool has_image = false;        // loadColorImage done: pyramid resident on the device
    int w = 0, h = 0;
    float flen = 0, paspect = 1, pp[2] = {0.5f, 0.5f}, rot[9], trans[3];
    float campos[3];
    float w2c[12];
    std::vector<HostLevel> lv;
    uchar4* d_base = nullptr;      // one allocation for every level: the RGBX8 images, then their quad images
    size_t bytes = 0;
};

struct SeedPoint { int x, y; float depth; };
struct HostPlan {                  // what DMRecon::start computes on the host before the queue runs (dmrecon.cc:179-292)
    b200mvs_settings settings;     // the settings it was made for
    std::vector<int> gsel;         // globalViewSelection result, ascending
    std::vector<SeedPoint> seeds;  // the feature loop of processFeatures
};

struct HostFeature {
    float pos[3];
    std::vector<int> refs;
};

struct Entry {                     // frontier queue entry = QueueData (dmrecon.h:28-38)
    int xy;                        // x | y << 16
    int jobdir;                    // job | dir << 24 ; -1 = dropped
    float conf, depth, dzI, dzJ;
    unsigned slots;
    int pad;
};
static_assert(sizeof(Entry) == 32, "Entry layout");

enum Counter { C_RUN = 0, C_NEXT = 1, C_OVERFLOW = 2, C_SETS = 3, C_OPTS = 4, C_SEED_OK = 5, C_FILLED = 6, C_TICKET = 7, C_NUM = 8 };

constexpr int HIST_FINE = 8192;                 // confidence bins of the eligibility threshold: bin = conf * 8192
constexpr int HIST_COARSE = 64;                 // one coarse bin per 128 fine bins
constexpr int HIST_PER_JOB = HIST_FINE + HIST_COARSE;
constexpr int DEFER_BIT = 1 << 27;              // Entry::jobdir flag: below this round's threshold, carried unchanged
enum Phase { PH_SEED = 0, PH_SELECT = 1, PH_THRESHOLD = 2, PH_PICK = 3, PH_OPT = 4, PH_COMMIT = 5, PH_EXPAND = 6, PH_SORT = 7, PH_OPT_THREAD = 8, PH_NUM = 10 };
enum Stop { ST_RUN = 0, ST_CANCELLED = 2, ST_OVERFLOW = 3 };

struct FrontierCtl {                            // device memory, zeroed before the launch
    unsigned long long bar;                     // grid barrier ticket counter
    unsigned long long nlist[2];                // entries in list[0] / list[1]
    unsigned long long nrun;                    // winners of the current round
    unsigned long long ticket;                  // next patch of the current optimise phase
    unsigned long long ticket2;                 // ... of the tail of a large round (warp-per-patch path)
    unsigned long long sort_cursor;             // next free position of run2 while a round is being grouped by tile
    unsigned long long small_cursor;            // ... behind the grouped entries, for the views that run one warp per patch
    unsigned long long rounds, peak, run_total, barriers;
    unsigned long long ns[PH_NUM];              // %globaltimer time per phase, measured by CTA 0
    unsigned long long thread_busy_ns;          // summed busy time of all warps in the one-thread-per-patch phases
    int stop;
    int pad;
};

struct HostMirror {                             // mapped pinned host memory; followed by filled[n_jobs] (device -> host)
    volatile int cancel;                        // host -> device: stop the whole batch at the next round
    volatile int pad;
    volatile unsigned long long round, queue;   // device -> host
};                                              // ... and by cancel_job[n_jobs] (host -> device: drop this view's queue)

template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

} // namespace

struct b200mvs_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::mutex mtx;
    std::string err;
    std::vector<HostView> views;
    std::vector<HostFeature> feats;
    std::vector<std::vector<int>> view_feats;   // inverted index: ascending ids of the features that reference a view
    ViewParams* d_views = nullptr;
    bool views_dirty = true;
    float* d_lut = nullptr;
    // workspace (grown on demand, reused across calls)
    DevBuf<Entry> ent_a, ent_b, run_in, run_sorted;
    DevBuf<unsigned> tile_cnt;
    DevBuf<unsigned long long> tile_off;
    DevBuf<PatchOut> run_out;
    DevBuf<unsigned char> written;
    DevBuf<unsigned long long> counters;
    DevBuf<JobParams> d_jobs;
    DevBuf<DevSettings> d_settings;
    DevBuf<unsigned char> maps;        // all per-job maps of the current batch
    DevBuf<FrontierCtl> ctl;
    DevBuf<unsigned> hist;
    DevBuf<int> thr_bin;
    DevBuf<int> job_cancel;
    DevBuf<unsigned long long> job_run;
    int frontier_grid = 0;             // CTAs of the cooperative launch (= what fits on the chip)
    int optimize_grid = 0;             // resident CTAs of k_optimize
    long long thread_min = -1;         // reconstruct: rounds with at least this many patches run one thread per patch (-1: default)
    int optimize_mode = 0;             // b200mvs_optimize_patches: 0 by batch size, 1 one warp per patch, 2 one thread per patch
    unsigned long long* h_counters = nullptr;   // pinned
    unsigned long long* h_mirror = nullptr;     // pinned + mapped: HostMirror
    std::vector<cudaEvent_t> ev_pool;
    // b200mvs_upload_view staging: two pinned host buffers + device buffers used alternately, so that the host copy of
    // view k+1 overlaps the H2D transfer and the pyramid kernels of view k
    struct Stage { uint8_t* host = nullptr; uint8_t* dev = nullptr; size_t cap = 0; cudaEvent_t done = nullptr; bool busy = false; };
    Stage stage[2];
    unsigned stage_next = 0;
    // host plans prepared ahead by b200mvs_plan_views (global view selection + seed list of a reference view)
    std::mutex plan_mtx;
    std::map<int, HostPlan> plans;
};

namespace {

int fail(b200mvs_ctx* c, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(ctx, B200MVS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// ---- reference float arithmetic on the host (libs/math conventions) ----
inline float dot3(const float* a, const float* b) { return ((0.0f + a[0] * b[0]) + a[1] * b[1]) + a[2] * b[2]; }
inline float sqn3(const float* a) { return dot3(a, a); }
inline void normalize3(float* a) { float n = std::sqrt(sqn3(a)); a[0] /= n; a[1] /= n; a[2] /= n; }
inline float round_mve(float x) { return x > 0.0f ? std::floor(x + 0.5f) : std::ceil(x - 0.5f); }
inline void world_to_cam(const HostView& v, const float* p, float* o)
{
    for (int i = 0; i < 3; ++i) o[i] = dot3(v.w2c + 4 * i, p) + 1.0f * v.w2c[4 * i + 3];
}
inline void mat3_mul(const float* m, const float* x, float* o) { for (int i = 0; i < 3; ++i) o[i] = dot3(m + 3 * i, x); }

// CameraInfo::fill_calibration / fill_inverse_calibration (camera.cc:125-144,180-200)
void fill_calibration(const HostView& v, float ppx, float ppy, float width, float height, float* K, float* Ki)
{
    const float dim_aspect = width / height;
    const float image_aspect = dim_aspect * v.paspect;
    float ax, ay;
    if (image_aspect < 1.0f) { ax = v.flen * height / v.paspect; ay = v.flen * height; }
    else                     { ax = v.flen * width;              ay = v.flen * width * v.paspect; }
    K[0] = ax;  K[1] = 0.f; K[2] = width * ppx;
    K[3] = 0.f; K[4] = ay;  K[5] = height * ppy;
    K[6] = 0.f; K[7] = 0.f; K[8] = 1.f;
    Ki[0] = 1.0f / ax; Ki[1] = 0.f;       Ki[2] = -width * ppx / ax;
    Ki[3] = 0.f;       Ki[4] = 1.0f / ay; Ki[5] = -height * ppy / ay;
    Ki[6] = 0.f;       Ki[7] = 0.f;       Ki[8] = 1.f;
}

// SingleView::pointInFrustum (single_view.cc:109-121)
bool point_in_frustum(const HostView& v, const float* wp)
{
    float cp[3], sp[3];
    world_to_cam(v, wp, cp);
    if (cp[2] <= 0.0f) return false;
    mat3_mul(v.lv[0].proj, cp, sp);
    const float x = sp[0] / sp[2] - 0.5f;
    const float y = sp[1] / sp[2] - 0.5f;
    return x >= 0 && x <= v.lv[0].w - 1 && y >= 0 && y <= v.lv[0].h - 1;
}
inline float foot_print(const HostView& v, int level, const float* p) { float c[3]; world_to_cam(v, p, c); return c[2] * v.lv[level].invproj[0]; }

bool in_aabb(const float* p, const b200mvs_settings& s)
{
    for (int i = 0; i < 3; ++i) if (p[i] < s.aabb_min[i] || p[i] > s.aabb_max[i]) return false;
    return true;
}

// DMRecon::analyzeFeatures (dmrecon.cc:179-208) + GlobalViewSelection (global_view_selection.cc:17-101).
// Kept on the host: its INTEGER result must match the reference bit for bit, so every float is produced by the
// reference's expression order.  What is restructured is only WHEN things are computed:
//   * (feature - camera centre).normalized() is computed once per (view, feature) instead of inside every parallax()
//     call (mvs_tools.h:46-53) - same operations, same values;
//   * `if (plx < minParallax) score *= sqr(plx / 10)` (global_view_selection.cc:80-81,93-97) needs acos only when the
//     directions are nearly parallel: for dot < cos(minParallax + 0.05 deg) the branch is certainly not taken;
//   * the factor of a selected view s on (candidate i, feature k) does not change between rounds; multiplying by the
//     exact 1.0f of a not-taken branch cannot change a rounding, so only the factors != 1 are stored (ascending s,
//     the std::set iteration order) and re-multiplied each round;
//   * the (candidate, feature) records are laid out feature-major, the order the greedy loop walks them in.
std::vector<int> global_view_selection(const b200mvs_ctx* c, const b200mvs_settings& st, int ref)
{
    const int nv = (int)c->views.size();
    const HostView& rv = c->views[ref];
    static const std::vector<int> no_feats;
    const std::vector<int>& of_ref = ref < (int)c->view_feats.size() ? c->view_feats[ref] : no_feats;
    const float dot_skip = (float)std::cos(((double)st.min_parallax + 0.05) * 3.14159265358979323846 / 180.0);
    auto unit_dir = [&](const HostView& v, const float* p, float* d) {
        d[0] = p[0] - v.campos[0]; d[1] = p[1] - v.campos[1]; d[2] = p[2] - v.campos[2];
        normalize3(d);
    };
    // parallax factor of global_view_selection.cc:80-81 / :93-97 from two unit directions
    auto plx_factor = [&](const float* d1, const float* d2) -> float {
        const float dt = dot3(d1, d2);
        if (dt < dot_skip) return 1.f;
        const float dp = std::max(std::min(dt, 1.f), -1.f);
        const float plx = std::acos(dp) * 180.f / 3.141592653589793f;
        if (plx < st.min_parallax) { const float q = plx / 10.f; return q * q; }
        return 1.f;
    };
    struct Extra { int k, view; float f; };       // a factor != 1 of selected view `view` on entry k of a candidate
    struct Cand {
        std::vector<float> base;         // per entry k (= SingleView::featInd order): parallax-with-ref and resolution terms (:78-87)
        std::vector<int> feat;           // per entry k: local index of its feature
        std::vector<int> ent;            // per entry k: its index in the feature-major arrays below
        std::vector<Extra> extra;        // ascending (k, view): the std::set iteration order of :90
        std::vector<Extra> fresh;        // the factors of the view selected last, before they are merged into `extra`
        float benefit = 0.f;             // benefitFromView of the last evaluation
        bool dirty = true;
    };
    std::vector<Cand> cand(nv);
    std::vector<char> avail(nv, 1);
    avail[ref] = 0;
    for (int v = 0; v < nv; ++v) if (!c->views[v].valid) avail[v] = 0;
    // Feature-major entries: everything the candidates hold about ONE feature is contiguous, because that is how the greedy
    // loop walks it - "which candidates see a feature the new view sees" is SingleView::seesFeature (single_view.h:166-174)
    // turned around.  A candidate's entries keep the order of the reference's featInd (ascending feature, then refs order).
    std::vector<int> foff(1, 0), ecand, ek;
    std::vector<float> edir;
    for (int fi : of_ref) {                      // the features with contains_view_id(refViewNr), ascending (dmrecon.cc:186-188)
        const HostFeature& f = c->feats[fi];
        if (!point_in_frustum(rv, f.pos)) continue;
        if (!in_aabb(f.pos, st)) continue;
        const int fl = (int)foff.size() - 1;
        bool have_ref = false;
        float dr[3] = {0.f, 0.f, 0.f}, mfp = 0.f;
        for (int vid : f.refs) {
            if (vid < 0 || vid >= nv || !avail[vid]) continue;            // the reference view itself is never a candidate
            const HostView& tv = c->views[vid];
            if (!point_in_frustum(tv, f.pos)) continue;
            if (!have_ref) { unit_dir(rv, f.pos, dr); mfp = foot_print(rv, st.scale, f.pos); have_ref = true; }
            Cand& C = cand[vid];
            const int e = (int)ecand.size();
            float d[3];
            unit_dir(tv, f.pos, d);
            float score = 1.f;
            score *= plx_factor(dr, d);
            const float nfp = foot_print(tv, 0, f.pos);
            float ratio = mfp / nfp;
            if (ratio > 2.) ratio = (float)(2. / ratio);
            else if (ratio > 1.) ratio = 1.;
            score *= ratio;
            ecand.push_back(vid); ek.push_back((int)C.base.size());
            edir.push_back(d[0]); edir.push_back(d[1]); edir.push_back(d[2]);
            C.base.push_back(score); C.feat.push_back(fl); C.ent.push_back(e);
        }
        if ((int)ecand.size() > foff.back()) foff.push_back((int)ecand.size());
    }
    std::vector<int> selected;
    std::vector<Extra> merged;
    bool found = true;
    while (found && selected.size() < st.global_vs_max) {
        float maxBenefit = 0.f;
        int maxView = 0;
        found = false;
        for (int i = 0; i < nv; ++i) {
            if (!avail[i]) continue;
            Cand& C = cand[i];
            if (C.dirty) {
                // recomputed only when a factor of this candidate changed since the last round: the same operations on
                // the same operands give the same float, so caching cannot change the result
                float benefit = 0;
                const size_t n = C.base.size();
                size_t x = 0;
                for (size_t k = 0; k < n; ++k) {
                    float score = C.base[k];
                    for (; x < C.extra.size() && C.extra[x].k == (int)k; ++x) score *= C.extra[x].f;
                    benefit += score;
                }
                C.benefit = benefit;
                C.dirty = false;
            }
            if (C.benefit > maxBenefit) { maxBenefit = C.benefit; maxView = i; found = true; }
        }
        if (!found) break;
        selected.insert(std::upper_bound(selected.begin(), selected.end(), maxView), maxView);
        avail[maxView] = 0;
        // fold the new view's factors into every remaining candidate.  Only (candidate, feature) entries whose feature
        // the new view sees can change (global_view_selection.cc:90-92).
        const Cand& S = cand[maxView];
        for (size_t ks = 0; ks < S.base.size(); ++ks) {
            const int fl = S.feat[ks];
            if (ks > 0 && S.feat[ks - 1] == fl) continue;              // seesFeature() is a predicate: a duplicate adds nothing
            const float* ds = &edir[3 * (size_t)S.ent[ks]];
            for (int e = foff[fl]; e < foff[fl + 1]; ++e) {
                const int i = ecand[e];
                if (!avail[i]) continue;
                const float f = plx_factor(ds, &edir[3 * (size_t)e]);
                if (f != 1.f) cand[i].fresh.push_back(Extra{ek[e], maxView, f});      // arrives with ascending k
            }
        }
        for (int i = 0; i < nv; ++i) {
            Cand& C = cand[i];
            if (C.fresh.empty()) continue;
            merged.resize(C.extra.size() + C.fresh.size());
            std::merge(C.extra.begin(), C.extra.end(), C.fresh.begin(), C.fresh.end(), merged.begin(),
                       [](const Extra& a, const Extra& b) { return a.k != b.k ? a.k < b.k : a.view < b.view; });
            C.extra.swap(merged);
            C.fresh.clear();
            C.dirty = true;
        }
    }
    return selected;
}

using Seed = SeedPoint;

// Host threads for the per-view host phase (global view selection + seed lists).  One process per GPU on a shared box
// should not start hardware_concurrency() threads each: B200MVS_HOST_THREADS caps it (bench.py sets cores / (2 * ranks)).
int host_threads(int n_jobs)
{
    int cap = (int)std::thread::hardware_concurrency();
    if (const char* e = std::getenv("B200MVS_HOST_THREADS")) { const int v = std::atoi(e); if (v > 0) cap = v; }
    return std::max(1, std::min(n_jobs, cap));
}

// The feature loop of DMRecon::processFeatures (dmrecon.cc:258-292): which features seed, at which pixel, with
// which initial depth.  The optimisation of the seeds runs on the device.
std::vector<Seed> collect_seeds(const b200mvs_ctx* c, const b200mvs_settings& st, int ref, const std::vector<int>& gsel)
{
    const HostView& rv = c->views[ref];
    std::vector<Seed> out;
    // "use feature if visible in reference view or at least one neighboring view" (dmrecon.cc:260-276), in feature order
    std::vector<char> wanted(c->feats.size(), 0);
    if (ref < (int)c->view_feats.size()) for (int fi : c->view_feats[ref]) wanted[fi] = 1;
    for (int g : gsel) if (g >= 0 && g < (int)c->view_feats.size()) for (int fi : c->view_feats[g]) wanted[fi] = 1;
    for (size_t fi = 0; fi < wanted.size(); ++fi) {
        if (!wanted[fi]) continue;
        const HostFeature& f = c->feats[fi];
        if (!point_in_frustum(rv, f.pos)) continue;
        if (!in_aabb(f.pos, st)) continue;
        float cp[3], sp[3];
        world_to_cam(rv, f.pos, cp);
        mat3_mul(rv.lv[st.scale].proj, cp, sp);
        const float px = sp[0] / sp[2] - 0.5f, py = sp[1] / sp[2] - 0.5f;
        const float dv[3] = {f.pos[0] - rv.campos[0], f.pos[1] - rv.campos[1], f.pos[2] - rv.campos[2]};
        Seed s;
        s.x = (int)round_mve(px);
        s.y = (int)round_mve(py);
        s.depth = std::sqrt(sqn3(dv));
        out.push_back(s);
    }
    return out;
}

// ------------------------------------------------------------------------------------------------
// kernels: image import + pyramid
// ------------------------------------------------------------------------------------------------
// `undistorted` bytes -> RGBX8 (alpha dropped, grey expanded: image_pyramid.cc:65-73)
__global__ void k_import_rgb(const uint8_t* __restrict__ src, int w, int h, int ch, uchar4* __restrict__ dst, int pitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t* p = src + ((size_t)y * w + x) * ch;
    uchar4 o;
    if (ch >= 3) { o.x = p[0]; o.y = p[1]; o.z = p[2]; }
    else         { o.x = o.y = o.z = p[0]; }
    o.w = 255;
    dst[(size_t)y * pitch + x] = o;
}

// mve::image::rescale_half_size_gaussian<uint8_t>(img, 1.f) (image_tools.h:617-694) with Accum<uint8>
// (accum.h:117-170): same 16 taps in the same order, fp32 accumulate, true division, math::round.
// Bit-exact with the reference (tests/test_gpu_parity.py::test_pyramid_bit_exact); products and sums are kept un-fused on purpose.
__global__ void k_half_gaussian(const uchar4* __restrict__ in, int iw, int ih, int ipitch,
                                uchar4* __restrict__ out, int ow, int oh, int opitch,
                                float w1, float w2, float w3)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= ow || y >= oh) return;
    const int y2 = y * 2, x2 = x * 2;
    const int ys[4] = {max(0, y2 - 1), y2, min(ih - 1, y2 + 1), min(ih - 1, y2 + 2)};
    const int xs[4] = {max(0, x2 - 1), x2, min(iw - 1, x2 + 1), min(iw - 1, x2 + 2)};
    const float wr[4][4] = {{w3, w2, w2, w3}, {w2, w1, w1, w2}, {w2, w1, w1, w2}, {w3, w2, w2, w3}};
    float v0 = 0.f, v1 = 0.f, v2 = 0.f, ws = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uchar4* row = in + (size_t)ys[r] * ipitch;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uchar4 t = row[xs[k]];
            v0 = __fadd_rn(v0, __fmul_rn((float)t.x, wr[r][k]));
            v1 = __fadd_rn(v1, __fmul_rn((float)t.y, wr[r][k]));
            v2 = __fadd_rn(v2, __fmul_rn((float)t.z, wr[r][k]));
            ws = __fadd_rn(ws, wr[r][k]);
        }
    }
    const float q0 = __fdiv_rn(v0, ws), q1 = __fdiv_rn(v1, ws), q2 = __fdiv_rn(v2, ws);
    uchar4 o;
    o.x = (unsigned char)(q0 > 0.f ? floorf(q0 + 0.5f) : ceilf(q0 - 0.5f));
    o.y = (unsigned char)(q1 > 0.f ? floorf(q1 + 0.5f) : ceilf(q1 - 0.5f));
    o.z = (unsigned char)(q2 > 0.f ? floorf(q2 + 0.5f) : ceilf(q2 - 0.5f));
    o.w = 255;
    out[(size_t)y * opitch + x] = o;
}

// The bilinear footprint of a sample at (x + fx, y + fy) is the 2x2 block {(x,y), (x+1,y), (x,y+1), (x+1,y+1)}
// (mvs_tools.cc:110-124): stored contiguously per (x, y) it is ONE aligned 16-byte load instead of four 4-byte ones.
__global__ void k_make_quads(const uchar4* __restrict__ img, int w, int h, int pitch, uint4* __restrict__ quad)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int x1 = min(x + 1, w - 1), y1 = min(y + 1, h - 1);
    const unsigned* p = reinterpret_cast<const unsigned*>(img);
    uint4 q;
    q.x = p[(size_t)y * pitch + x]; q.y = p[(size_t)y * pitch + x1];
    q.z = p[(size_t)y1 * pitch + x]; q.w = p[(size_t)y1 * pitch + x1];
    quad[(size_t)y * pitch + x] = q;
}

__global__ void k_export_rgb(const uchar4* __restrict__ src, int w, int h, int pitch, uint8_t* __restrict__ dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const uchar4 t = src[(size_t)y * pitch + x];
    uint8_t* p = dst + ((size_t)y * w + x) * 3;
    p[0] = t.x; p[1] = t.y; p[2] = t.z;
}

// ------------------------------------------------------------------------------------------------
// kernels: patch optimisation + frontier
// ------------------------------------------------------------------------------------------------
#ifndef OPT_TPB
#define OPT_TPB 512          // threads per CTA of the patch-optimisation kernels: one CTA per SM shares ONE 64 KB table
#endif
#ifndef OPT_MIN_BLOCKS
#define OPT_MIN_BLOCKS 1     // CTAs per SM: registers per thread <= 65536 / (OPT_MIN_BLOCKS * OPT_TPB) = 128; shared memory 180 KB per CTA
#endif
constexpr int OPT_WARPS = OPT_TPB / 32;
using PatchT = b200mvs::PatchW;    // one warp per patch (latency: small rounds)
using PatchT1 = b200mvs::PatchT;   // one thread per patch (throughput: large rounds)
// dynamic shared memory of the kernels that optimise patches: the lane-replicated sRGB table, then one PatchT1 per thread.
// The state of a thread's patch is live across the whole sample loop, which needs every register there is: left to the
// compiler it is spilled to local memory (~100 slots x 128 B per warp - more than L1 holds, so every reload in the per-view
// set-up and tear-down was an L2 round trip: 19 % of the kernel's time in ncu's stall samples).  In shared memory a reload
// costs a fixed ~30 cycles.
constexpr size_t OPT_LUT_BYTES = sizeof(float) * (256 * LUT_STRIDE);
constexpr size_t OPT_SMEM_BYTES = OPT_LUT_BYTES + (size_t)OPT_TPB * sizeof(PatchT1);
__device__ __forceinline__ PatchT1& thread_patch(float* smem)
{
    return reinterpret_cast<PatchT1*>(reinterpret_cast<unsigned char*>(smem) + OPT_LUT_BYTES)[threadIdx.x];
}

__device__ __forceinline__ Entry load_entry(const Entry* p)       // lists are rewritten by other SMs every round: bypass L1
{
    const int4 a = __ldcg(reinterpret_cast<const int4*>(p));
    const int4 b = __ldcg(reinterpret_cast<const int4*>(p) + 1);
    Entry e;
    e.xy = a.x; e.jobdir = a.y; e.conf = __int_as_float(a.z); e.depth = __int_as_float(a.w);
    e.dzI = __int_as_float(b.x); e.dzJ = __int_as_float(b.y); e.slots = (unsigned)b.z; e.pad = b.w;
    return e;
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// The PatchOptimizations of list[0..n): every warp takes entries through the ticket counter until none is left; a warp that
// finishes one fetches the next at once.  PatchOptimization ctor + doAutoOptimization + computeConfidence per entry.
__device__ __forceinline__ void optimise_entries(PatchT& p, const Entry* list, PatchOut* res, unsigned long long n,
                                                 unsigned long long* ticket, const JobParams* jobs, unsigned long long* counters)
{
    unsigned opts = 0u;
    for (;;) {
        unsigned long long w = 0ull;
        if (p.lane == 0) w = atomicAdd(ticket, 1ull);
        w = __shfl_sync(FULL, w, 0);
        if (w >= n) break;
        const Entry e = load_entry(&list[w]);
        PatchIn pi;
        pi.x = e.xy & 0xFFFF; pi.y = (e.xy >> 16) & 0xFFFF;
        pi.depth = e.depth; pi.dzI = e.dzI; pi.dzJ = e.dzJ; pi.slots = e.slots;
        p.begin(&jobs[e.jobdir & 0xFFFFFF], pi);
        ++opts;
        while (!p.step()) {}
        PatchOut po;
        p.finish(po);
        if (p.lane == 0) res[w] = po;
    }
    if (p.lane == 0 && opts) {
        atomicAdd(&counters[C_SETS], (unsigned long long)p.n_sets);
        atomicAdd(&counters[C_OPTS], (unsigned long long)opts);
        p.n_sets = 0u;
    }
}

// The same for LARGE lists: one thread per entry (patch_thread.cuh).  Lanes that need an entry take consecutive tickets with
// one atomic per converged subset of the warp; a lane that finishes fetches its next entry at once and meets the other
// lanes of its warp again at the pass() call site.
__device__ __forceinline__ void optimise_entries_t(PatchT1& p, const Entry* list, PatchOut* res, unsigned long long n,
                                                   unsigned long long* ticket, const JobParams* jobs, unsigned long long* counters,
                                                   unsigned long long* exit_sum = nullptr)
{
    const int lane = threadIdx.x & 31;
    const unsigned long long t_phase = exit_sum ? global_timer_ns() : 0ull;
    bool have = false;
    unsigned long long idx = 0ull;
    unsigned opts = 0u;
    for (;;) {
        if (!have) {
            const unsigned act = __activemask();
            const int leader = __ffs(act) - 1;
            unsigned long long base = 0ull;
            if (lane == leader) base = atomicAdd(ticket, (unsigned long long)__popc(act));
            base = __shfl_sync(act, base, leader);
            const unsigned long long w = base + (unsigned long long)__popc(act & ((1u << lane) - 1u));
            if (w >= n) break;
            idx = w;
            const Entry e = load_entry(&list[w]);
            PatchIn pi;
            pi.x = e.xy & 0xFFFF; pi.y = (e.xy >> 16) & 0xFFFF;
            pi.depth = e.depth; pi.dzI = e.dzI; pi.dzJ = e.dzJ; pi.slots = e.slots;
            p.begin(&jobs[e.jobdir & 0xFFFFFF], pi);
            have = true;
            ++opts;
        }
        if (p.step()) {
            PatchOut po;
            p.finish(po);
            res[idx] = po;
            have = false;
        }
    }
    if (opts) {
        atomicAdd(&counters[C_SETS], (unsigned long long)p.n_sets);
        atomicAdd(&counters[C_OPTS], (unsigned long long)opts);
        p.n_sets = 0u;
    }
    __syncwarp();
    if (lane == 0 && exit_sum) { atomicAdd(exit_sum, global_timer_ns() - t_phase); }     // busy time of this warp in the phase
}

#ifndef OPT_THREAD_MIN
#define OPT_THREAD_MIN 8192    // lists at least this long are optimised one thread per patch, shorter ones one warp per patch
#endif

// A batch of independent PatchOptimizations (b200mvs_optimize_patches).  mode: 0 = by list length, 1 = one warp per patch,
// 2 = one thread per patch.
__global__ void __launch_bounds__(OPT_TPB, OPT_MIN_BLOCKS)
k_optimize(const Entry* __restrict__ in, PatchOut* __restrict__ out, int n, int mode,
           const DevSettings* __restrict__ st, const JobParams* __restrict__ jobs, const ViewParams* __restrict__ views,
           const float* __restrict__ g_lut, unsigned long long* counters)
{
    extern __shared__ float smem[];
    for (int i = threadIdx.x; i < 256 * LUT_STRIDE; i += blockDim.x) smem[i] = g_lut[i / LUT_STRIDE];
    __syncthreads();
    if (mode == 2 || (mode == 0 && n >= OPT_THREAD_MIN)) {
        PatchT1& p = thread_patch(smem);
        bind_thread(p, st, views, smem, (int)threadIdx.x);
        optimise_entries_t(p, in, out, (unsigned long long)n, &counters[C_TICKET], jobs, counters);
    } else {
        PatchT p;
        bind_thread(p, st, views, smem, (int)threadIdx.x);
        optimise_entries(p, in, out, (unsigned long long)n, &counters[C_TICKET], jobs, counters);
    }
}

__device__ __forceinline__ unsigned long long entry_key(const Entry& e)
{
    // larger confidence first, then smaller direction code (DESIGN.md "Frontier schedule")
    return ((unsigned long long)__float_as_uint(e.conf) << 8) | (unsigned long long)(7 - ((e.jobdir >> 24) & 7));
}

// ---- the whole region growing of a batch in ONE persistent kernel -------------------------------------------------------
// processFeatures + processQueue (dmrecon.cc:244-434) as frontier rounds (DESIGN.md "Frontier schedule").  The grid is
// launched cooperatively with exactly as many CTAs as fit on the chip; the phases of a round are separated by a grid-wide
// barrier instead of kernel boundaries, the patch optimisations of a round are handed out warp by warp through a ticket
// counter, and the host is not involved until the queue is empty: progress goes out and the cancel flag comes in through
// mapped pinned memory once per round (Progress, progress.h:27-43; dmrecon.cc:353).
struct FrontierParams {
    Entry* list[2];
    Entry* run;
    Entry* run2;                                // the round's winners grouped by 16x16 tile (large rounds)
    unsigned* tile_cnt;                         // [n_tiles] entries per tile, zero between rounds
    unsigned long long* tile_off;               // [n_tiles] start of the tile's segment in run2
    long long n_tiles;
    PatchOut* res;
    unsigned char* written;
    unsigned long long cap;
    int n_seeds, n_jobs;
    const DevSettings* st;
    const JobParams* jobs;
    const ViewParams* views;
    const float* lut;
    unsigned long long* counters;               // Counter + filled per job
    FrontierCtl* ctl;
    unsigned* hist;                             // [n_jobs][HIST_PER_JOB], zero on entry (only with a threshold)
    int* thr_bin;                               // [n_jobs]
    HostMirror* host;
    volatile unsigned long long* host_filled;   // [n_jobs], mapped
    volatile int* host_cancel_job;              // [n_jobs], mapped
    int* job_cancel;                            // [n_jobs], device copy refreshed every round
    long long thread_min;                       // a view with at least this many patches in a round runs them one thread per patch
    unsigned long long* job_run;                // [n_jobs] winners of the round per view
    int band_bins;                              // frontier_band in fine bins (0 = off)
    int topk;                                   // frontier_topk (0 = off)
};

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// All CTAs are co-resident (cooperative launch), so a monotone ticket counter is a barrier: the k-th generation is
// complete when the counter reaches k * gridDim.x.  The wait polls with a RELAXED load: an acquire load in the loop
// (ld.acquire = LDG + CCTL.IVALL) would invalidate the SM's L1 on every poll and starve the CTAs of the same SM that are
// still sampling (measured: L1 hit rate 47 % -> profiles/r2_notes.md); one fence after the wait orders the phase.  Data that
// other SMs write during the kernel is read with ld.cg everywhere, so no L1 invalidation is needed for correctness.
__device__ __forceinline__ void grid_barrier(unsigned long long* bar)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long nb = gridDim.x;
        const unsigned long long t = atomicAdd(bar, 1ull);
        const unsigned long long target = (t / nb + 1ull) * nb;
        while (ld_relaxed_u64(bar) < target) __nanosleep(64);
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ int conf_bin(float c)
{
    const int b = (int)(c * (float)HIST_FINE);
    return b < 0 ? 0 : (b > HIST_FINE - 1 ? HIST_FINE - 1 : b);
}
__device__ __forceinline__ PatchOut load_result(const PatchOut* p)  // written by another SM in the optimise phase: bypass L1
{
    static_assert(sizeof(PatchOut) == 40, "PatchOut layout");
    const int2* q = reinterpret_cast<const int2*>(p);
    const int2 a = __ldcg(q), b = __ldcg(q + 1), c = __ldcg(q + 2), d = __ldcg(q + 3), e = __ldcg(q + 4);
    PatchOut r;
    r.conf = __int_as_float(a.x); r.depth = __int_as_float(a.y); r.dzI = __int_as_float(b.x); r.dzJ = __int_as_float(b.y);
    r.nx = __int_as_float(c.x); r.ny = __int_as_float(c.y); r.nz = __int_as_float(d.x); r.slots = (unsigned)d.y;
    r.iterations = e.x; r.flags = e.y;
    return r;
}
// Warp-aggregated slot allocation: the lanes of a warp that reach this point together take consecutive slots with ONE atomic
// (the queue bookkeeping appends ~20 M entries per step to three counters; one same-address atomic per entry cost ~20 ms).
__device__ __forceinline__ unsigned long long take_slot(unsigned long long* counter)
{
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned long long base = 0ull;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
}
// the same for a counter per job: lanes with the same job share one atomic
__device__ __forceinline__ void count_for_job(unsigned long long* per_job, int j)
{
    const unsigned m = __match_any_sync(__activemask(), j);
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&per_job[j], (unsigned long long)__popc(m));
}
__device__ __forceinline__ void push_entry(const FrontierParams& P, int q, const Entry& o)
{
    const unsigned long long pos = take_slot(&P.ctl->nlist[q]);
    if (pos < P.cap) P.list[q][pos] = o; else P.counters[C_OVERFLOW] = 1ull;
}
__device__ __forceinline__ void write_pixel(const JobParams& J, int idx, const PatchOut& r)
{
    J.depth[idx] = r.depth;
    J.conf[idx] = r.conf;
    J.dz[2 * idx] = r.dzI; J.dz[2 * idx + 1] = r.dzJ;
    J.normal[3 * idx] = r.nx; J.normal[3 * idx + 1] = r.ny; J.normal[3 * idx + 2] = r.nz;
    J.slots[idx] = r.slots;
}

// Eligibility threshold of one job from its confidence histogram (one warp).  topk: the largest bin t such that at least
// `topk` entries have bin >= t (0 when there are fewer); band: (highest non-empty bin) - band_bins.  Both: the larger.
__device__ __forceinline__ int threshold_of(const unsigned* hist, int lane, int band_bins, int topk)
{
    const unsigned* coarse = hist + HIST_FINE;
    int t_top = 0, t_band = 0;
    // coarse scan, descending: lane l looks at coarse bins 63 - l and 31 - l
    const unsigned c_hi = __ldcg(&coarse[63 - lane]), c_lo = __ldcg(&coarse[31 - lane]);
    if (band_bins > 0) {
        const unsigned m_hi = __ballot_sync(FULL, c_hi != 0u), m_lo = __ballot_sync(FULL, c_lo != 0u);
        int cb = -1;
        if (m_hi) cb = 63 - (__ffs(m_hi) - 1); else if (m_lo) cb = 31 - (__ffs(m_lo) - 1);
        if (cb >= 0) {
            int top = -1;
            for (int q = 3; q >= 0 && top < 0; --q) {
                const unsigned v = __ldcg(&hist[cb * 128 + q * 32 + (31 - lane)]);
                const unsigned m = __ballot_sync(FULL, v != 0u);
                if (m) top = cb * 128 + q * 32 + 31 - (__ffs(m) - 1);
            }
            t_band = top - band_bins;
            if (t_band < 0) t_band = 0;
        }
    }
    if (topk > 0) {
        unsigned need = (unsigned)topk, before = 0u;
        int cb = -1;
        for (int half = 0; half < 2 && cb < 0; ++half) {
            const unsigned c = half == 0 ? c_hi : c_lo;
            unsigned incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
            const unsigned m = __ballot_sync(FULL, before + incl >= need);
            if (m) {
                const int l = __ffs(m) - 1;
                cb = (half == 0 ? 63 : 31) - l;
                before += __shfl_sync(FULL, incl, l) - __shfl_sync(FULL, c, l);
            } else
                before += __shfl_sync(FULL, incl, 31);
        }
        if (cb >= 0) {
            for (int q = 3; q >= 0; --q) {
                const unsigned c = __ldcg(&hist[cb * 128 + q * 32 + (31 - lane)]);
                unsigned incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
                const unsigned m = __ballot_sync(FULL, before + incl >= need);
                if (m) { t_top = cb * 128 + q * 32 + 31 - (__ffs(m) - 1); break; }
                before += __shfl_sync(FULL, incl, 31);
            }
        }
    }
    return t_top > t_band ? t_top : t_band;
}

__global__ void __launch_bounds__(OPT_TPB, OPT_MIN_BLOCKS)
k_frontier(const FrontierParams P)
{
    extern __shared__ float smem[];
    for (int i = threadIdx.x; i < 256 * LUT_STRIDE; i += blockDim.x) smem[i] = P.lut[i / LUT_STRIDE];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (size_t)gridDim.x * blockDim.x;
    FrontierCtl* const ctl = P.ctl;
    unsigned long long* const cnt = P.counters;
    unsigned long long* const filled = cnt + C_NUM;
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    const bool thresholded = P.band_bins > 0 || P.topk > 0;
    unsigned long long t_prev = lead ? global_timer_ns() : 0ull;
    unsigned long long n_bar = 0ull;
#define PHASE_END(ph) do { grid_barrier(&ctl->bar); ++n_bar; if (lead) { const unsigned long long t_ = global_timer_ns(); ctl->ns[ph] += t_ - t_prev; t_prev = t_; } } while (0)

    // ---- processQueue as frontier rounds (dmrecon.cc:334-434) ----
    // Round 0 optimises the seeds (processFeatures, dmrecon.cc:293-326: per pixel the most confident seed is committed, first
    // in feature order on ties), every later round one frontier.  The optimise phase has ONE call site so that the patch
    // optimisation code exists once in the instruction stream.
    int p = 1;                                                   // the seed round pushes into list[0]
    bool seed_round = P.n_seeds > 0;
    for (;;) {
        unsigned long long n_cur = seed_round ? (unsigned long long)P.n_seeds : __ldcg(&ctl->nlist[p]);
        if (n_cur > P.cap) n_cur = P.cap;
        if (__ldcg(&ctl->stop) != ST_RUN || n_cur == 0ull) break;
        unsigned long long n_run = n_cur;
        if (!seed_round) {
        Entry* const cur = P.list[p];
        if (blockIdx.x == 0) {                                  // progress out (dmrecon.cc:355-364)
            if (threadIdx.x == 0) { P.host->round = ctl->rounds; P.host->queue = n_cur; }
            for (int j = threadIdx.x; j < P.n_jobs; j += blockDim.x) {
                P.host_filled[j] = __ldcg(&filled[j]);
                if (P.host_cancel_job[j]) P.job_cancel[j] = 1;   // this view's entries are dropped from now on (Progress::cancelled)
            }
        }
        // A: stale test (dmrecon.cc:371-373); without a threshold the per-pixel bid follows immediately
        for (size_t i = gtid; i < n_cur; i += gthreads) {
            const Entry e = load_entry(&cur[i]);
            const int j = e.jobdir & 0xFFFFFF;
            const JobParams& J = P.jobs[j];
            const int idx = ((e.xy >> 16) & 0xFFFF) * J.W + (e.xy & 0xFFFF);
            if (__ldcg(&J.conf[idx]) > e.conf || __ldcg(&P.job_cancel[j])) { cur[i].jobdir = -1; continue; }
            if (!thresholded) atomicMax(&J.sel[idx], entry_key(e));
            else {
                const int b = conf_bin(e.conf);
                atomicAdd(&P.hist[(size_t)j * HIST_PER_JOB + b], 1u);
                atomicAdd(&P.hist[(size_t)j * HIST_PER_JOB + HIST_FINE + (b >> 7)], 1u);
            }
        }
        PHASE_END(PH_SELECT);
        if (thresholded) {
            // B: this round's eligibility threshold of every job, then the bids of the eligible entries
            for (int j = blockIdx.x * OPT_WARPS + (threadIdx.x >> 5); j < P.n_jobs; j += gridDim.x * OPT_WARPS) {
                const int t = threshold_of(P.hist + (size_t)j * HIST_PER_JOB, lane, P.band_bins, P.topk);
                if (lane == 0) P.thr_bin[j] = t;
            }
            PHASE_END(PH_THRESHOLD);
            for (size_t i = gtid; i < n_cur; i += gthreads) {
                const Entry e = load_entry(&cur[i]);
                if (e.jobdir == -1) continue;
                const int j = e.jobdir & 0xFFFFFF;
                if (conf_bin(e.conf) < __ldcg(&P.thr_bin[j])) { cur[i].jobdir = e.jobdir | DEFER_BIT; continue; }
                const JobParams& J = P.jobs[j];
                const int idx = ((e.xy >> 16) & 0xFFFF) * J.W + (e.xy & 0xFFFF);
                atomicMax(&J.sel[idx], entry_key(e));
            }
            PHASE_END(PH_THRESHOLD);
        }
        // C: the winning bid of each pixel runs, every other live entry is carried to the next round (the reference would
        // pop it later and apply the stale test then)
        for (size_t i = gtid; i < n_cur; i += gthreads) {
            Entry e = load_entry(&cur[i]);
            if (e.jobdir == -1) continue;
            if (e.jobdir & DEFER_BIT) { e.jobdir &= ~DEFER_BIT; push_entry(P, 1 - p, e); continue; }
            const JobParams& J = P.jobs[e.jobdir & 0xFFFFFF];
            const int idx = ((e.xy >> 16) & 0xFFFF) * J.W + (e.xy & 0xFFFF);
            const unsigned long long key = entry_key(e);
            if (__ldcg(&J.sel[idx]) == key && atomicCAS(&J.sel[idx], key, 0ull) == key) {
                const unsigned long long pos = take_slot(&ctl->nrun);
                P.run[pos] = e;
                count_for_job(P.job_run, e.jobdir & 0xFFFFFF);                                   // |run| <= n_cur <= capacity
            } else
                push_entry(P, 1 - p, e);
        }
        if (thresholded)
            for (size_t i = gtid; i < (size_t)P.n_jobs * HIST_PER_JOB; i += gthreads) P.hist[i] = 0u;
        PHASE_END(PH_PICK);
        n_run = __ldcg(&ctl->nrun);
        if (lead) ctl->nlist[p] = 0ull;                           // consumed; the round after the next pushes into it
        }
        // Which implementation optimises an entry is decided PER VIEW: a view with at least thread_min winners in this round runs
        // one thread per patch, a view with fewer one warp per patch.  The rule depends on the view alone, so its maps do not
        // depend on which other views share the batch (the two implementations differ in rounding).  The entries of the
        // "thread" views are grouped by 16x16-pixel tile first, so that the lanes of a warp and the warps of an SM sample
        // overlapping windows of the neighbour images (L1 / coalescing): count per tile, one atomic cursor bump per non-empty
        // tile, scatter; the entries of the other views follow behind them.
        unsigned long long n_thread = 0ull;
        const Entry* run_cur = P.run;
        if (!seed_round && P.n_tiles > 0) {
            __shared__ unsigned long long s_big;
            if (threadIdx.x == 0) s_big = 0ull;
            __syncthreads();
            unsigned long long mine = 0ull;
            for (int j = threadIdx.x; j < P.n_jobs; j += blockDim.x) {
                const unsigned long long c = __ldcg(&P.job_run[j]);
                if (c >= (unsigned long long)P.thread_min) mine += c;
            }
            if (mine) atomicAdd(&s_big, mine);
            __syncthreads();
            n_thread = s_big;
        }
        if (n_thread > 0ull) {
            for (size_t i = gtid; i < n_run; i += gthreads) {
                const Entry e = load_entry(&P.run[i]);
                const int j = e.jobdir & 0xFFFFFF;
                if (__ldcg(&P.job_run[j]) < (unsigned long long)P.thread_min) continue;
                const JobParams& J = P.jobs[j];
                const long long bin = J.tile_base + (long long)(((e.xy >> 16) & 0xFFFF) >> 4) * J.tiles_x + ((e.xy & 0xFFFF) >> 4);
                atomicAdd(&P.tile_cnt[bin], 1u);
            }
            PHASE_END(PH_SORT);
            for (size_t b = gtid; b < (size_t)P.n_tiles; b += gthreads) {
                const unsigned c = __ldcg(&P.tile_cnt[b]);
                if (c) P.tile_off[b] = atomicAdd(&ctl->sort_cursor, (unsigned long long)c);
            }
            PHASE_END(PH_SORT);
            for (size_t i = gtid; i < n_run; i += gthreads) {
                const Entry e = load_entry(&P.run[i]);
                const int j = e.jobdir & 0xFFFFFF;
                if (__ldcg(&P.job_run[j]) < (unsigned long long)P.thread_min) {
                    P.run2[n_thread + take_slot(&ctl->small_cursor)] = e;
                    continue;
                }
                const JobParams& J = P.jobs[j];
                const long long bin = J.tile_base + (long long)(((e.xy >> 16) & 0xFFFF) >> 4) * J.tiles_x + ((e.xy & 0xFFFF) >> 4);
                // the tile's counter is consumed downwards: it is zero again when the tile's last entry has been placed
                const unsigned k = atomicSub(&P.tile_cnt[bin], 1u) - 1u;
                P.run2[__ldcg(&P.tile_off[bin]) + k] = e;
            }
            if (lead) { ctl->sort_cursor = 0ull; ctl->small_cursor = 0ull; }
            PHASE_END(PH_SORT);
            run_cur = P.run2;
        }
        const bool by_thread = n_thread > 0ull;
        // the PatchOptimizations of the round
        if (by_thread) {
            PatchT1& pt = thread_patch(smem);
            bind_thread(pt, P.st, P.views, smem, (int)threadIdx.x);
            optimise_entries_t(pt, run_cur, P.res, n_thread, &ctl->ticket, P.jobs, cnt, &ctl->thread_busy_ns);
            __syncwarp();
        }
        if (n_run > n_thread) {
            PatchT pg;
            bind_thread(pg, P.st, P.views, smem, (int)threadIdx.x);
            optimise_entries(pg, run_cur + n_thread, P.res + n_thread, n_run - n_thread, by_thread ? &ctl->ticket2 : &ctl->ticket, P.jobs, cnt);
        }
        if (seed_round) {
            PHASE_END(PH_SEED);
            for (size_t i = gtid; i < (size_t)P.n_seeds; i += gthreads) {
                const float c = __ldcg(&P.res[i].conf);
                if (!(c > 0.f)) continue;
                atomicAdd(&cnt[C_SEED_OK], 1ull);
                const Entry e = load_entry(&run_cur[i]);
                const JobParams& J = P.jobs[e.jobdir & 0xFFFFFF];
                const int idx = ((e.xy >> 16) & 0xFFFF) * J.W + (e.xy & 0xFFFF);
                atomicMax(&J.sel[idx], ((unsigned long long)__float_as_uint(c) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i));
            }
            PHASE_END(PH_SELECT);
            for (size_t i = gtid; i < (size_t)P.n_seeds; i += gthreads) {
                const float c = __ldcg(&P.res[i].conf);
                if (!(c > 0.f)) continue;
                const Entry e = load_entry(&run_cur[i]);
                const int j = e.jobdir & 0xFFFFFF;
                const JobParams& J = P.jobs[j];
                const int idx = ((e.xy >> 16) & 0xFFFF) * J.W + (e.xy & 0xFFFF);
                const unsigned long long key = ((unsigned long long)__float_as_uint(c) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
                if (__ldcg(&J.sel[idx]) != key) continue;
                J.sel[idx] = 0ull;
                count_for_job(filled, j);
                const PatchOut r = load_result(&P.res[i]);
                write_pixel(J, idx, r);
                Entry o;
                o.xy = e.xy; o.jobdir = j | (4 << 24);
                o.conf = r.conf; o.depth = r.depth; o.dzI = r.dzI; o.dzJ = r.dzJ; o.slots = r.slots; o.pad = 0;
                push_entry(P, 0, o);
            }
            if (lead) ctl->ticket = 0ull;
            PHASE_END(PH_COMMIT);
            seed_round = false;
            p = 0;
            continue;
        }
        if (by_thread) PHASE_END(PH_OPT_THREAD); else PHASE_END(PH_OPT);
        // D: commit (dmrecon.cc:377-398).  One winner per pixel, so plain stores.
        for (size_t i = gtid; i < n_run; i += gthreads) {
            const Entry e = load_entry(&run_cur[i]);
            const PatchOut r = load_result(&P.res[i]);
            const int j = e.jobdir & 0xFFFFFF;
            const JobParams& J = P.jobs[j];
            const int idx = ((e.xy >> 16) & 0xFFFF) * J.W + (e.xy & 0xFFFF);
            unsigned char w = 0;
            if (!(r.conf == 0.f)) {
                const float old = __ldcg(&J.conf[idx]);
                if (old <= 0.f) count_for_job(filled, j);
                if (old < r.conf) { write_pixel(J, idx, r); w = 1; }
            }
            P.written[i] = w;
        }
        if (lead) { ctl->nrun = 0ull; ctl->ticket = 0ull; ctl->ticket2 = 0ull; }
        if (blockIdx.x == 0) for (int j = threadIdx.x; j < P.n_jobs; j += blockDim.x) P.job_run[j] = 0ull;
        PHASE_END(PH_COMMIT);
        // E: after ALL commits, push the 4-neighbours of every committed pixel (dmrecon.cc:400-431)
        for (size_t i = gtid; i < n_run; i += gthreads) {
            if (!P.written[i]) continue;
            const Entry e = load_entry(&run_cur[i]);
            const PatchOut r = load_result(&P.res[i]);
            const int j = e.jobdir & 0xFFFFFF;
            const JobParams& J = P.jobs[j];
            const int x = e.xy & 0xFFFF, y = (e.xy >> 16) & 0xFFFF;
            const int nx[4] = {x - 1, x + 1, x, x};
            const int ny[4] = {y, y, y - 1, y + 1};
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const float c = __ldcg(&J.conf[ny[d] * J.W + nx[d]]);
                if (c < r.conf - 0.05f || c == 0.f) {
                    Entry o;
                    o.xy = nx[d] | (ny[d] << 16);
                    o.jobdir = j | (d << 24);
                    o.conf = r.conf; o.depth = r.depth; o.dzI = r.dzI; o.dzJ = r.dzJ; o.slots = r.slots; o.pad = 0;
                    push_entry(P, 1 - p, o);
                }
            }
        }
        if (lead) {
            ctl->rounds += 1ull;
            ctl->run_total += n_run;
            if (n_cur > ctl->peak) ctl->peak = n_cur;
            if (P.host->cancel) ctl->stop = ST_CANCELLED;
            else if (__ldcg(&cnt[C_OVERFLOW])) ctl->stop = ST_OVERFLOW;
        }
        PHASE_END(PH_EXPAND);
        p ^= 1;
    }
    if (lead) ctl->barriers = n_bar;
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) { P.host->round = ctl->rounds; P.host->queue = 0ull; }
        for (int j = threadIdx.x; j < P.n_jobs; j += blockDim.x) P.host_filled[j] = __ldcg(&filled[j]);
    }
#undef PHASE_END
}

// ------------------------------------------------------------------------------------------------
// host: view upload
// ------------------------------------------------------------------------------------------------
int check_settings(b200mvs_ctx* ctx, const b200mvs_settings* s)
{
    if (!s) return fail(ctx, B200MVS_ERR_INVALID_ARG, "settings is NULL");
    if (s->filter_width != 5)
        return fail(ctx, B200MVS_ERR_UNSUPPORTED, "filterWidth must be 5 (the reference hard-codes patchPoints[12], patch_sampler.cc:96)");
    if (s->scale < 0) return fail(ctx, B200MVS_ERR_INVALID_ARG, "Invalid scale factor");
    if (s->nr_recon_neighbors < 1 || s->nr_recon_neighbors > B200MVS_MAX_LOCAL_VIEWS)
        return fail(ctx, B200MVS_ERR_UNSUPPORTED, "nrReconNeighbors must be in 1..%d", B200MVS_MAX_LOCAL_VIEWS);
    if (s->global_vs_max < 1 || s->global_vs_max > B200MVS_MAX_GLOBAL_VIEWS)
        return fail(ctx, B200MVS_ERR_UNSUPPORTED, "globalVSMax must be in 1..%d", B200MVS_MAX_GLOBAL_VIEWS);
    if (!(s->frontier_band >= 0.f) || s->frontier_band > 1.f)
        return fail(ctx, B200MVS_ERR_INVALID_ARG, "frontier_band must be in [0, 1]");
    return 0;
}

DevSettings to_dev(const b200mvs_settings& s)
{
    DevSettings d;
    d.min_ncc = s.min_ncc; d.min_parallax = s.min_parallax; d.accept_ncc = s.accept_ncc; d.min_refine_diff = s.min_refine_diff;
    d.max_iterations = s.max_iterations; d.nr_recon_neighbors = s.nr_recon_neighbors;
    d.scale = s.scale; d.use_color_scale = s.use_color_scale;
    return d;
}

int upload_common(b200mvs_ctx* ctx, int id, const uint8_t* d_src, int w, int h, int ch, float flen, float paspect,
                  const float* pp, const float* rot, const float* trans, cudaStream_t stream)
{
    HostView& v = ctx->views[id];
    bool same_camera = false;
    {
        // prepared plans depend on cameras and image sizes only: a re-upload of the same view with the same camera keeps them
        same_camera = v.valid && !v.lv.empty() && v.w == w && v.h == h && v.flen == flen && v.paspect == paspect && v.pp[0] == pp[0] && v.pp[1] == pp[1] &&
                      std::memcmp(v.rot, rot, sizeof(v.rot)) == 0 && std::memcmp(v.trans, trans, sizeof(v.trans)) == 0;
        if (!same_camera) { std::lock_guard<std::mutex> pl(ctx->plan_mtx); ctx->plans.clear(); }
    }
    // An unchanged camera keeps its host-side record untouched: b200mvs_plan_views of other views may be reading it right
    // now (cameras, calibrations and level sizes; never the device pointers written below).
    if (!same_camera) {
        v.valid = false;
        v.has_image = false;
        v.w = w; v.h = h; v.flen = flen; v.paspect = paspect; v.pp[0] = pp[0]; v.pp[1] = pp[1];
        std::memcpy(v.rot, rot, sizeof(v.rot));
        std::memcpy(v.trans, trans, sizeof(v.trans));
        // CameraInfo::fill_camera_pos / fill_world_to_cam (camera.cc:34-39,61-67)
        v.campos[0] = -rot[0] * trans[0] - rot[3] * trans[1] - rot[6] * trans[2];
        v.campos[1] = -rot[1] * trans[0] - rot[4] * trans[1] - rot[7] * trans[2];
        v.campos[2] = -rot[2] * trans[0] - rot[5] * trans[1] - rot[8] * trans[2];
        for (int r = 0; r < 3; ++r) {
            v.w2c[4 * r] = rot[3 * r]; v.w2c[4 * r + 1] = rot[3 * r + 1]; v.w2c[4 * r + 2] = rot[3 * r + 2]; v.w2c[4 * r + 3] = trans[r];
        }
        // buildPyramid (image_pyramid.cc:22-53)
        v.lv.clear();
        float ppx = pp[0], ppy = pp[1];
        int cw = w, chh = h;
        auto push_level = [&]() {
            HostLevel L; L.w = cw; L.h = chh; L.pitch = (cw + 3) & ~3;
            fill_calibration(v, ppx, ppy, (float)cw, (float)chh, L.proj, L.invproj);
            v.lv.push_back(L);
        };
        push_level();
        while (std::min(cw, chh) >= 30) {
            if (cw % 2 == 1) ppx = ppx * float(cw) / float(cw + 1);
            if (chh % 2 == 1) ppy = ppy * float(chh) / float(chh + 1);
            cw = (cw + 1) / 2; chh = (chh + 1) / 2;
            push_level();
        }
    }
    if ((int)v.lv.size() > MAX_LEVELS) return fail(ctx, B200MVS_ERR_UNSUPPORTED, "image too large: %d pyramid levels", (int)v.lv.size());
    if (!d_src) {
        // camera only (SingleView::create, single_view.cc:24-53): the image follows with an upload if the view turns out
        // to be a reference view or a selected neighbour (loadColorImage, dmrecon.cc:78,238-240)
        if (v.d_base) { cudaFree(v.d_base); v.d_base = nullptr; v.bytes = 0; }
        v.has_image = false;
        v.valid = true;
        ctx->views_dirty = true;
        return 0;
    }
    size_t total = 0;
    for (HostLevel& L : v.lv) total += (size_t)L.pitch * L.h;
    const size_t need = total * (sizeof(uchar4) + sizeof(uint4));
    if (!v.d_base || v.bytes != need) {
        if (v.d_base) { cudaFree(v.d_base); v.d_base = nullptr; }
        CK(cudaMalloc(&v.d_base, need));
        v.bytes = need;
        ctx->views_dirty = true;
    }
    size_t off = 0;
    uint4* qbase = reinterpret_cast<uint4*>(v.d_base + total);      // total is a multiple of 4 texels: 16-byte aligned
    for (HostLevel& L : v.lv) { L.d_img = v.d_base + off; L.d_quad = qbase + off; off += (size_t)L.pitch * L.h; }
    const dim3 blk(32, 8);
    k_import_rgb<<<dim3((w + 31) / 32, (h + 7) / 8), blk, 0, stream>>>(d_src, w, h, ch, v.lv[0].d_img, v.lv[0].pitch);
    // ensureImages (image_pyramid.cc:56-95): rescale_half_size_gaussian(img, 1.f) level by level
    const float w1 = std::exp(-0.5f / (2.0f * 1.0f)), w2 = std::exp(-2.5f / (2.0f * 1.0f)), w3 = std::exp(-4.5f / (2.0f * 1.0f));
    for (size_t i = 1; i < v.lv.size(); ++i) {
        const HostLevel& a = v.lv[i - 1];
        const HostLevel& b = v.lv[i];
        k_half_gaussian<<<dim3((b.w + 31) / 32, (b.h + 7) / 8), blk, 0, stream>>>(a.d_img, a.w, a.h, a.pitch, b.d_img, b.w, b.h, b.pitch, w1, w2, w3);
    }
    for (const HostLevel& L : v.lv)
        k_make_quads<<<dim3((L.w + 31) / 32, (L.h + 7) / 8), blk, 0, stream>>>(L.d_img, L.w, L.h, L.pitch, L.d_quad);
    CK(cudaGetLastError());
    v.valid = true;
    v.has_image = true;
    ctx->views_dirty = true;
    return 0;
}

int sync_view_params(b200mvs_ctx* ctx)
{
    if (!ctx->views_dirty) return 0;
    std::vector<ViewParams> hp(ctx->views.size());
    std::memset(hp.data(), 0, hp.size() * sizeof(ViewParams));
    for (size_t i = 0; i < ctx->views.size(); ++i) {
        const HostView& v = ctx->views[i];
        ViewParams& p = hp[i];
        p.valid = v.valid ? 1 : 0;
        if (!v.valid) continue;
        std::memcpy(p.campos, v.campos, 12);
        std::memcpy(p.w2c, v.w2c, 48);
        std::memcpy(p.rot, v.rot, 36);
        p.inv_ax0 = v.lv[0].invproj[0];
        p.nlevels = (int)v.lv.size();
        for (size_t l = 0; l < v.lv.size(); ++l) {
            const HostLevel& L = v.lv[l];
            p.lv[l].ax = L.proj[0]; p.lv[l].ay = L.proj[4]; p.lv[l].cx = L.proj[2]; p.lv[l].cy = L.proj[5];
            p.lv[l].w = L.w; p.lv[l].h = L.h; p.lv[l].pitch = L.pitch; p.lv[l].img = v.has_image ? L.d_img : nullptr;
            p.lv[l].quad = v.has_image ? L.d_quad : nullptr;
        }
    }
    CK(cudaMemcpyAsync(ctx->d_views, hp.data(), hp.size() * sizeof(ViewParams), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->views_dirty = false;
    return 0;
}

cudaEvent_t get_event(b200mvs_ctx* ctx, size_t i)
{
    while (ctx->ev_pool.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); ctx->ev_pool.push_back(e); }
    return ctx->ev_pool[i];
}

// Opt-in to > 48 KB of dynamic shared memory and size the grids to what is resident on the chip (once per context).
int prepare_kernels(b200mvs_ctx* ctx)
{
    if (ctx->frontier_grid) return 0;
    CK(cudaFuncSetAttribute(k_frontier, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OPT_SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_optimize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OPT_SMEM_BYTES));
    int per_sm = 0, per_sm_opt = 0, sms = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_frontier, OPT_TPB, OPT_SMEM_BYTES));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_opt, k_optimize, OPT_TPB, OPT_SMEM_BYTES));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    if (per_sm < 1 || per_sm_opt < 1) return fail(ctx, B200MVS_ERR_CUDA, "patch-optimisation kernels do not fit on an SM");
    ctx->optimize_grid = per_sm_opt * sms;
    ctx->frontier_grid = per_sm * sms;
    return 0;
}

// JobParams of one reference view (everything except the map pointers)
int make_job(b200mvs_ctx* ctx, const b200mvs_settings& s, int ref, const std::vector<int>& gsel, JobParams& J)
{
    const HostView& rv = ctx->views[ref];
    if (s.scale >= (int)rv.lv.size()) return fail(ctx, B200MVS_ERR_INVALID_ARG, "Invalid scale factor");
    const HostLevel& L = rv.lv[s.scale];
    std::memset(&J, 0, sizeof(J));
    J.ref_view = ref; J.W = L.w; J.H = L.h; J.n_global = (int)gsel.size();
    for (size_t k = 0; k < gsel.size(); ++k) J.gview[k] = gsel[k];
    J.ki0 = L.invproj[0]; J.ki2 = L.invproj[2]; J.ki4 = L.invproj[4]; J.ki5 = L.invproj[5];
    J.ref_img = L.d_img; J.ref_pitch = L.pitch;
    return 0;
}

unsigned ids_to_slots(const std::vector<int>& gsel, const int32_t* ids, int n, bool* ok)
{
    unsigned s = 0xFFFFFFFFu;
    std::vector<int> slots;
    for (int k = 0; k < n; ++k) {
        auto it = std::lower_bound(gsel.begin(), gsel.end(), (int)ids[k]);
        if (it == gsel.end() || *it != ids[k]) { *ok = false; return s; }
        slots.push_back((int)(it - gsel.begin()));
    }
    std::sort(slots.begin(), slots.end());
    for (size_t k = 0; k < slots.size() && k < 4; ++k) s = (s & ~(0xFFu << (8 * k))) | ((unsigned)slots[k] << (8 * k));
    *ok = true;
    return s;
}

} // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* b200mvs_version(void) { return "b200mvs 0.1 (sm_100a)"; }

void b200mvs_default_settings(b200mvs_settings* s)
{
    if (!s) return;
    s->filter_width = 5; s->min_ncc = 0.3f; s->min_parallax = 10.0f; s->accept_ncc = 0.6f; s->min_refine_diff = 0.001f;
    s->max_iterations = 20; s->nr_recon_neighbors = 4; s->global_vs_max = 20; s->scale = 0; s->use_color_scale = 1;
    for (int i = 0; i < 3; ++i) { s->aabb_min[i] = -3.402823466e+38f; s->aabb_max[i] = 3.402823466e+38f; }
    s->frontier_band = 0.f; s->frontier_topk = 0;
}

const char* b200mvs_last_error(const b200mvs_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int b200mvs_create(int device, int n_views, b200mvs_ctx** out)
{
    b200mvs_ctx* ctx = nullptr;
    if (!out || n_views <= 0) return fail(nullptr, B200MVS_ERR_INVALID_ARG, "b200mvs_create: bad arguments");
    *out = nullptr;
    if (device == B200MVS_DEVICE_NONE) {
        // planning context: cameras, features, global view selection (pure host logic); every compute entry point fails
        ctx = new b200mvs_ctx();
        ctx->device = B200MVS_DEVICE_NONE;
        ctx->views.resize(n_views);
        *out = ctx;
        return 0;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, B200MVS_ERR_CUDA, "no CUDA device available (%s); b200mvs has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, B200MVS_ERR_INVALID_ARG, "device %d out of range (%d devices)", device, ndev);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, B200MVS_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    ctx = new b200mvs_ctx();
    ctx->device = device;
    ctx->views.resize(n_views);
    auto bail = [&](const char* what, cudaError_t ce) {
        fail(nullptr, B200MVS_ERR_CUDA, "%s: %s", what, cudaGetErrorString(ce));
        delete ctx;
        return B200MVS_ERR_CUDA;
    };
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaMalloc(&ctx->d_views, sizeof(ViewParams) * n_views)) != cudaSuccess) return bail("cudaMalloc(views)", e);
    if ((e = cudaMalloc(&ctx->d_lut, 256 * sizeof(float))) != cudaSuccess) return bail("cudaMalloc(lut)", e);
    if ((e = cudaMallocHost(&ctx->h_counters, sizeof(unsigned long long) * 4096)) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaHostAlloc(&ctx->h_mirror, sizeof(unsigned long long) * 8192, cudaHostAllocMapped)) != cudaSuccess) return bail("cudaHostAlloc", e);
    // sRGB code value -> linear: the formula documented at mvs_tools.cc:21-29; tests/test_oracle_vs_reference.py::test_srgb_table_matches_reference checks the
    // 256 floats against the reference table.
    float lut[256];
    for (int i = 0; i < 256; ++i) {
        const double x = i / 255.0;
        lut[i] = (float)((i <= 0.04045 * 255.0) ? x / 12.92 : std::pow((x + 0.055) / 1.055, 2.4));
    }
    if ((e = cudaMemcpy(ctx->d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice)) != cudaSuccess) return bail("cudaMemcpy(lut)", e);
    *out = ctx;
    return 0;
}

void b200mvs_destroy(b200mvs_ctx* ctx)
{
    if (!ctx) return;
    if (ctx->device == B200MVS_DEVICE_NONE) { delete ctx; return; }
    cudaSetDevice(ctx->device);
    for (HostView& v : ctx->views) if (v.d_base) cudaFree(v.d_base);
    if (ctx->d_views) cudaFree(ctx->d_views);
    if (ctx->d_lut) cudaFree(ctx->d_lut);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_mirror) cudaFreeHost(ctx->h_mirror);
    for (b200mvs_ctx::Stage& S : ctx->stage) {
        if (S.host) cudaFreeHost(S.host);
        if (S.dev) cudaFree(S.dev);
        if (S.done) cudaEventDestroy(S.done);
    }
    ctx->ent_a.release(); ctx->ent_b.release(); ctx->run_in.release(); ctx->run_sorted.release(); ctx->tile_cnt.release(); ctx->tile_off.release(); ctx->run_out.release(); ctx->written.release();
    ctx->counters.release(); ctx->d_jobs.release(); ctx->d_settings.release(); ctx->maps.release();
    ctx->ctl.release(); ctx->hist.release(); ctx->thr_bin.release(); ctx->job_cancel.release(); ctx->job_run.release();
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int b200mvs_upload_view(b200mvs_ctx* ctx, int id, const uint8_t* rgb, int w, int h, int channels,
                        float flen, float paspect, const float ppoint[2], const float rot[9], const float trans[3])
{
    if (!ctx) return B200MVS_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mtx);
    if (ctx->device == B200MVS_DEVICE_NONE) return fail(ctx, B200MVS_ERR_CUDA, "planning context (B200MVS_DEVICE_NONE): no CUDA device, b200mvs has no CPU fallback");
    if (id < 0 || id >= (int)ctx->views.size() || !rgb || w < 2 || h < 2 || !ppoint || !rot || !trans)
        return fail(ctx, B200MVS_ERR_INVALID_ARG, "b200mvs_upload_view: bad arguments");
    if (channels < 1 || channels > 4) return fail(ctx, B200MVS_ERR_INVALID_ARG, "Image with invalid number of channels");
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)w * h * channels;
    b200mvs_ctx::Stage& S = ctx->stage[ctx->stage_next++ & 1u];
    if (S.busy) { CK(cudaEventSynchronize(S.done)); S.busy = false; }      // the transfer that used this slot two uploads ago
    if (S.cap < bytes) {
        if (S.host) cudaFreeHost(S.host);
        if (S.dev) cudaFree(S.dev);
        S.host = nullptr; S.dev = nullptr; S.cap = 0;
        CK(cudaMallocHost(&S.host, bytes));
        CK(cudaMalloc(&S.dev, bytes));
        S.cap = bytes;
    }
    if (!S.done) CK(cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming));
    std::memcpy(S.host, rgb, bytes);                                         // pageable -> pinned; returns the caller's buffer
    CK(cudaMemcpyAsync(S.dev, S.host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    const int rc = upload_common(ctx, id, S.dev, w, h, channels, flen, paspect, ppoint, rot, trans, ctx->stream);
    CK(cudaEventRecord(S.done, ctx->stream));
    S.busy = true;
    // no synchronisation here: everything that reads the pyramid is ordered behind it on the context's stream
    return rc;
}

int b200mvs_upload_view_device(b200mvs_ctx* ctx, int id, const uint8_t* rgb_dev, int w, int h,
                               float flen, float paspect, const float ppoint[2], const float rot[9], const float trans[3],
                               void* cuda_stream)
{
    if (!ctx) return B200MVS_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mtx);
    if (ctx->device == B200MVS_DEVICE_NONE) return fail(ctx, B200MVS_ERR_CUDA, "planning context (B200MVS_DEVICE_NONE): no CUDA device, b200mvs has no CPU fallback");
    if (id < 0 || id >= (int)ctx->views.size() || !rgb_dev || w < 2 || h < 2 || !ppoint || !rot || !trans)
        return fail(ctx, B200MVS_ERR_INVALID_ARG, "b200mvs_upload_view_device: bad arguments");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    int rc = upload_common(ctx, id, rgb_dev, w, h, 3, flen, paspect, ppoint, rot, trans, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (!rc && e != cudaSuccess) rc = fail(ctx, B200MVS_ERR_CUDA, "upload sync: %s", cudaGetErrorString(e));
    return rc;
}

int b200mvs_set_view_camera(b200mvs_ctx* ctx, int id, int w, int h, float flen, float paspect, const float ppoint[2],
                            const float rot[9], const float trans[3])
{
    if (!ctx) return B200MVS_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mtx);
    if (id < 0 || id >= (int)ctx->views.size() || w < 2 || h < 2 || !ppoint || !rot || !trans)
        return fail(ctx, B200MVS_ERR_INVALID_ARG, "b200mvs_set_view_camera: bad arguments");
    if (ctx->device != B200MVS_DEVICE_NONE) CK(cudaSetDevice(ctx->device));
    return upload_common(ctx, id, nullptr, w, h, 3, flen, paspect, ppoint, rot, trans, ctx->stream);
}

int b200mvs_set_features(b200mvs_ctx* ctx, int n, const float* pos, const int32_t* off, const int32_t* ids)
{
    if (!ctx) return B200MVS_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mtx);
    if (n < 0 || (n > 0 && (!pos || !off || !ids))) return fail(ctx, B200MVS_ERR_INVALID_ARG, "b200mvs_set_features: bad arguments");
    { std::lock_guard<std::mutex> pl(ctx->plan_mtx); ctx->plans.clear(); }
    ctx->feats.resize(n);
    for (int i = 0; i < n; ++i) {
        std::memcpy(ctx->feats[i].pos, pos + 3 * i, 12);
        ctx->feats[i].refs.assign(ids + off[i], ids + off[i + 1]);
    }
    // Feature3D::contains_view_id (bundle.cc:15-21) for every view at once
    ctx->view_feats.assign(ctx->views.size(), std::vector<int>());
    for (int i = 0; i < n; ++i)
        for (int vid : ctx->feats[i].refs)
            if (vid >= 0 && vid < (int)ctx->views.size()) {
                std::vector<int>& vf = ctx->view_feats[vid];
                if (vf.empty() || vf.back() != i) vf.push_back(i);
            }
    return 0;
}

int b200mvs_num_levels(b200mvs_ctx* ctx, int id)
{
    if (!ctx) return B200MVS_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mtx);
    if (id < 0 || id >= (int)ctx->views.size() || !ctx->views[id].valid) return B200MVS_ERR_INVALID_ARG;
    return (int)ctx->views[id].lv.size();
}

int b200mvs_get_level(b200mvs_ctx* ctx, int id, int level, int* w, int* h, uint8_t* rgb)
{
    if (!ctx) return B200MVS_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mtx);
    if (id < 0 || id >= (int)ctx->views.size() || !ctx->views[id].valid) return fail(ctx, B200MVS_ERR_INVALID_ARG, "invalid view");
    const HostView& v = ctx->views[id];
    if (level < 0 || level >= (int)v.lv.size()) return fail(ctx, B200MVS_ERR_INVALID_ARG, "invalid level");
    const HostLevel& L = v.lv[level];
    if (w) *w = L.w;
    if (h) *h = L.h;
    if (!rgb) return 0;
    if (!v.has_image) return fail(ctx, B200MVS_ERR_INVALID_ARG, "color image of view %d is not loaded", id);
    if (ctx->device == B200MVS_DEVICE_NONE) return fail(ctx, B200MVS_ERR_CUDA, "planning context: no CUDA device");
    CK(cudaSetDevice(ctx->device));
    uint8_t* d = nullptr;
    CK(cudaMalloc(&d, (size_t)L.w * L.h * 3));
    k_export_rgb<<<dim3((L.w + 31) / 32, (L.h + 7) / 8), dim3(32, 8), 0, ctx->stream>>>(L.d_img, L.w, L.h, L.pitch, d);
    cudaError_t e = cudaMemcpyAsync(rgb, d, (size_t)L.w * L.h * 3, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, B200MVS_ERR_CUDA, "get_level copy: %s", cudaGetErrorString(e));
    return 0;
}

int b200mvs_global_view_selection(b200mvs_ctx* ctx, const b200mvs_settings* s, int ref, int32_t* ids_out, int cap)
{
    if (!ctx) return B200MVS_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mtx);
    int rc = check_settings(ctx, s);
    if (rc) return rc;
    if (ref < 0 || ref >= (int)ctx->views.size()) return fail(ctx, B200MVS_ERR_INVALID_ARG, "Master view index out of bounds");
    if (!ctx->views[ref].valid) return fail(ctx, B200MVS_ERR_INVALID_ARG, "Invalid master view");
    if (s->scale >= (int)ctx->views[ref].lv.size()) return fail(ctx, B200MVS_ERR_INVALID_ARG, "Invalid scale factor");
    if (cap > 0 && !ids_out) return fail(ctx, B200MVS_ERR_INVALID_ARG, "b200mvs_global_view_selection: ids_out is NULL");
    std::vector<int> sel;
    bool planned = false;
    {
        // a plan prepared by b200mvs_plan_views already holds the selection (it stays there for b200mvs_reconstruct)
        std::lock_guard<std::mutex> pl(ctx->plan_mtx);
        auto it = ctx->plans.find(ref);
        if (it != ctx->plans.end() && std::memcmp(&it->second.settings, s, sizeof(*s)) == 0) { sel = it->second.gsel; planned = true; }
    }
    if (!planned) sel = global_view_selection(ctx, *s, ref);
    for (int i = 0; i < (int)sel.size() && i < cap; ++i) ids_out[i] = sel[i];
    return (int)sel.size();
}

int b200mvs_set_patch_mode(b200mvs_ctx* ctx, int mode, int64_t thread_min)
{
    if (!ctx || mode < 0 || mode > 2) return B200MVS_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mtx);
    ctx->optimize_mode = mode;
    ctx->thread_min = thread_min;
    return 0;
}

int b200mvs_plan_views(b200mvs_ctx* ctx, const b200mvs_settings* s, int n_refs, const int32_t* refs)
{
    if (!ctx) return B200MVS_ERR_INVALID_ARG;
    // Deliberately NOT under ctx->mtx: planning reads only cameras and features and may overlap a running
    // b200mvs_reconstruct of another batch (the caller must not change cameras / features meanwhile).
    if (!s || n_refs < 0 || (n_refs > 0 && !refs)) return B200MVS_ERR_INVALID_ARG;
    if (s->filter_width != 5 || s->scale < 0 || s->global_vs_max < 1 || s->global_vs_max > B200MVS_MAX_GLOBAL_VIEWS) return B200MVS_ERR_INVALID_ARG;
    std::vector<HostPlan> made(n_refs);
    std::atomic<int> next_job(0);
    std::atomic<int> bad(0);
    auto worker = [&]() {
        for (;;) {
            const int j = next_job.fetch_add(1);
            if (j >= n_refs) break;
            const int r = refs[j];
            if (r < 0 || r >= (int)ctx->views.size() || !ctx->views[r].valid || s->scale >= (int)ctx->views[r].lv.size()) { bad = 1; continue; }
            made[j].settings = *s;
            made[j].gsel = global_view_selection(ctx, *s, r);
            if (!made[j].gsel.empty()) made[j].seeds = collect_seeds(ctx, *s, r, made[j].gsel);
        }
    };
    const int n_threads = host_threads(n_refs);
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker);
    worker();
    for (std::thread& t : pool) t.join();
    if (bad) return B200MVS_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> pl(ctx->plan_mtx);
    for (int j = 0; j < n_refs; ++j) ctx->plans[refs[j]] = std::move(made[j]);
    return 0;
}

int b200mvs_optimize_patches(b200mvs_ctx* ctx, const b200mvs_settings* s, int ref, const int32_t* gids, int ng,
                             const b200mvs_patch_in* in, int n, b200mvs_patch_out* out, b200mvs_stats* stats)
{
    if (!ctx) return B200MVS_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mtx);
    if (ctx->device == B200MVS_DEVICE_NONE) return fail(ctx, B200MVS_ERR_CUDA, "planning context (B200MVS_DEVICE_NONE): no CUDA device, b200mvs has no CPU fallback");
    int rc = check_settings(ctx, s);
    if (rc) return rc;
    if (ref < 0 || ref >= (int)ctx->views.size()) return fail(ctx, B200MVS_ERR_INVALID_ARG, "Master view index out of bounds");
    if (!ctx->views[ref].valid) return fail(ctx, B200MVS_ERR_INVALID_ARG, "Invalid master view");
    if (ng < 1 || ng > MAX_GLOBAL || !gids || n < 0 || (n > 0 && (!in || !out)))
        return fail(ctx, B200MVS_ERR_INVALID_ARG, "b200mvs_optimize_patches: bad arguments");
    std::vector<int> gsel(gids, gids + ng);
    if (!std::is_sorted(gsel.begin(), gsel.end())) return fail(ctx, B200MVS_ERR_INVALID_ARG, "global ids must be ascending");
    for (int g : gsel) if (g < 0 || g >= (int)ctx->views.size() || !ctx->views[g].valid) return fail(ctx, B200MVS_ERR_INVALID_ARG, "invalid global view id");
    if (!ctx->views[ref].has_image) return fail(ctx, B200MVS_ERR_INVALID_ARG, "color image of view %d is not loaded", ref);
    for (int g : gsel) if (!ctx->views[g].has_image) return fail(ctx, B200MVS_ERR_INVALID_ARG, "color image of view %d is not loaded", g);
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    if ((rc = sync_view_params(ctx))) return rc;
    JobParams J;
    if ((rc = make_job(ctx, *s, ref, gsel, J))) return rc;
    std::vector<Entry> he(n);
    for (int i = 0; i < n; ++i) {
        bool ok = true;
        if (in[i].n_local < 0 || in[i].n_local > 4) return fail(ctx, B200MVS_ERR_INVALID_ARG, "patch %d: n_local out of range", i);
        if (in[i].x < 0 || in[i].y < 0 || in[i].x > 0xFFFF || in[i].y > 0xFFFF) { he[i].xy = 0xFFFF | (0xFFFF << 16); }
        else he[i].xy = in[i].x | (in[i].y << 16);
        he[i].jobdir = 0;
        he[i].conf = 0.f; he[i].depth = in[i].depth; he[i].dzI = in[i].dz_i; he[i].dzJ = in[i].dz_j;
        he[i].slots = ids_to_slots(gsel, in[i].local_ids, in[i].n_local, &ok);
        he[i].pad = 0;
        if (!ok) return fail(ctx, B200MVS_ERR_INVALID_ARG, "patch %d: local view id not in the global set", i);
    }
    CK(ctx->run_in.reserve(n));
    CK(ctx->run_out.reserve(n));
    CK(ctx->counters.reserve(C_NUM + 8));
    CK(ctx->d_jobs.reserve(1));
    CK(ctx->d_settings.reserve(1));
    const DevSettings ds = to_dev(*s);
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->run_in.p, he.data(), sizeof(Entry) * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_jobs.p, &J, sizeof(J), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_settings.p, &ds, sizeof(ds), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->counters.p, 0, sizeof(unsigned long long) * C_NUM, st));
    cudaEvent_t e0 = get_event(ctx, 0), e1 = get_event(ctx, 1);
    CK(cudaEventRecord(e0, st));
    {
        int rc2 = prepare_kernels(ctx);
        if (rc2) return rc2;
        const int warps_per_block = OPT_TPB / 32;
        const int grid = std::max(1, std::min((n + warps_per_block - 1) / warps_per_block, ctx->optimize_grid));
        k_optimize<<<grid, OPT_TPB, OPT_SMEM_BYTES, st>>>(ctx->run_in.p, ctx->run_out.p, n, ctx->optimize_mode, ctx->d_settings.p,
                                                          ctx->d_jobs.p, ctx->d_views, ctx->d_lut, ctx->counters.p);
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1, st));
    std::vector<PatchOut> ho(n);
    CK(cudaMemcpyAsync(ho.data(), ctx->run_out.p, sizeof(PatchOut) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->h_counters, ctx->counters.p, sizeof(unsigned long long) * C_NUM, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int i = 0; i < n; ++i) {
        const PatchOut& r = ho[i];
        b200mvs_patch_out& o = out[i];
        o.conf = r.conf; o.depth = r.depth; o.dz_i = r.dzI; o.dz_j = r.dzJ;
        o.normal[0] = r.nx; o.normal[1] = r.ny; o.normal[2] = r.nz;
        o.n_local = 0;
        for (int k = 0; k < 4; ++k) {
            const int sl = (r.slots >> (8 * k)) & 0xFF;
            o.local_ids[k] = (sl != 0xFF && sl < ng) ? gsel[sl] : -1;
            if (o.local_ids[k] >= 0) o.n_local++;
        }
        o.iterations = r.iterations; o.converged = r.flags & 1; o.opti_success = (r.flags >> 1) & 1;
    }
    if (stats) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        stats->n_opt = ctx->h_counters[C_OPTS];
        stats->n_sample_sets = ctx->h_counters[C_SETS];
        stats->ms_patch_kernel = ms; stats->ms_total_device = ms;
        stats->n_patch_launches = 1; stats->n_kernel_launches = 1;
    }
    return 0;
}

int b200mvs_reconstruct(b200mvs_ctx* ctx, const b200mvs_settings* s, int n_refs, const int32_t* refs,
                        b200mvs_maps* maps, b200mvs_progress* progress, b200mvs_stats* stats, int32_t* failed_view)
{
    if (!ctx) return B200MVS_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mtx);
    if (ctx->device == B200MVS_DEVICE_NONE) return fail(ctx, B200MVS_ERR_CUDA, "planning context (B200MVS_DEVICE_NONE): no CUDA device, b200mvs has no CPU fallback");
    int rc = check_settings(ctx, s);
    if (rc) return rc;
    if (n_refs < 1 || !refs || n_refs > 4000) return fail(ctx, B200MVS_ERR_INVALID_ARG, "b200mvs_reconstruct: bad arguments");
    if (failed_view) *failed_view = -1;
    if (stats) std::memset(stats, 0, sizeof(*stats));
    for (int j = 0; j < n_refs; ++j) {
        const int r = refs[j];
        if (failed_view) *failed_view = r;
        if (r < 0 || r >= (int)ctx->views.size()) return fail(ctx, B200MVS_ERR_INVALID_ARG, "Master view index out of bounds");
        if (!ctx->views[r].valid) return fail(ctx, B200MVS_ERR_INVALID_ARG, "Invalid master view");
        if (s->scale >= (int)ctx->views[r].lv.size()) return fail(ctx, B200MVS_ERR_INVALID_ARG, "Invalid scale factor");
        if (ctx->views[r].lv[s->scale].w > 0xFFFF || ctx->views[r].lv[s->scale].h > 0xFFFF)
            return fail(ctx, B200MVS_ERR_UNSUPPORTED, "reference level larger than 65535 pixels per side");
    }
    if (failed_view) *failed_view = -1;
    CK(cudaSetDevice(ctx->device));
    if ((rc = sync_view_params(ctx))) return rc;
    cudaStream_t st = ctx->stream;
    const std::time_t t_start = std::time(nullptr);
    std::memset(ctx->h_counters, 0, sizeof(unsigned long long) * 4096);

    // ---- host phase per view: analyzeFeatures + globalViewSelection + seed list (dmrecon.cc:179-292) ----
    std::vector<JobParams> jobs(n_refs);
    std::vector<std::vector<int>> gsels(n_refs);
    std::vector<Entry> seeds;
    size_t total_px = 0, n_tiles = 0;
    std::vector<size_t> px_off(n_refs);
    // The per-view host work is independent (the reference runs whole DMRecons on OpenMP threads,
    // apps/dmrecon/dmrecon.cc:285): spread it over host threads.
    std::vector<std::vector<Seed>> seed_lists(n_refs);
    {
        // a plan prepared ahead (b200mvs_plan_views, possibly while the previous batch was running) is used once
        std::vector<int> todo;
        {
            std::lock_guard<std::mutex> pl(ctx->plan_mtx);
            for (int j = 0; j < n_refs; ++j) {
                if (progress) { progress[j].status = 1; progress[j].start_time = (uint64_t)t_start; progress[j].filled = 0; progress[j].queue_size = 0; }
                auto it = ctx->plans.find(refs[j]);
                if (it != ctx->plans.end() && std::memcmp(&it->second.settings, s, sizeof(*s)) == 0) {
                    gsels[j] = std::move(it->second.gsel);
                    seed_lists[j] = std::move(it->second.seeds);
                    ctx->plans.erase(it);
                    if (progress) progress[j].status = 2;
                } else {
                    todo.push_back(j);
                }
            }
        }
        std::atomic<size_t> next_job(0);
        auto worker = [&]() {
            for (;;) {
                const size_t k = next_job.fetch_add(1);
                if (k >= todo.size()) break;
                const int j = todo[k];
                gsels[j] = global_view_selection(ctx, *s, refs[j]);
                if (gsels[j].empty()) continue;
                if (progress) progress[j].status = 2;
                seed_lists[j] = collect_seeds(ctx, *s, refs[j], gsels[j]);
            }
        };
        const int n_threads = host_threads((int)todo.size());
        std::vector<std::thread> pool;
        for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker);
        worker();
        for (std::thread& t : pool) t.join();
    }
    for (int j = 0; j < n_refs; ++j) {
        if (gsels[j].empty()) {
            if (failed_view) *failed_view = refs[j];
            return fail(ctx, B200MVS_ERR_GLOBAL_VS, "Global View Selection failed");
        }
        if (!ctx->views[refs[j]].has_image) { if (failed_view) *failed_view = refs[j]; return fail(ctx, B200MVS_ERR_INVALID_ARG, "color image of view %d is not loaded", refs[j]); }
        for (int g : gsels[j])
            if (!ctx->views[g].has_image) { if (failed_view) *failed_view = refs[j]; return fail(ctx, B200MVS_ERR_INVALID_ARG, "color image of view %d (selected neighbour of view %d) is not loaded", g, refs[j]); }
        if ((rc = make_job(ctx, *s, refs[j], gsels[j], jobs[j]))) return rc;
        px_off[j] = total_px;
        total_px += (size_t)jobs[j].W * jobs[j].H;
        jobs[j].tiles_x = (jobs[j].W + 15) / 16;
        jobs[j].tile_base = (long long)n_tiles;
        n_tiles += (size_t)jobs[j].tiles_x * ((jobs[j].H + 15) / 16);
        for (const Seed& q : seed_lists[j]) {
            Entry e;
            // a seed outside the image fails in the PatchSampler ctor (patch_sampler.cc:47-50); keep it so that the
            // processed count matches, the kernel rejects it by the same bounds test
            const int x = std::min(std::max(q.x, -1), 0xFFFE), y = std::min(std::max(q.y, -1), 0xFFFE);
            e.xy = (x & 0xFFFF) | ((y & 0xFFFF) << 16);
            e.jobdir = j | (4 << 24);
            e.conf = 0.f; e.depth = q.depth; e.dzI = 0.f; e.dzJ = 0.f; e.slots = 0xFFFFFFFFu; e.pad = 0;
            seeds.push_back(e);
        }
        if (stats) stats->n_seeds_processed += seed_lists[j].size();
    }

    // ---- device buffers ----
    const size_t per_px = 4 + 4 + 8 + 12 + 4 + 8;   // depth, conf, dz, normal, slots, sel
    CK(ctx->maps.reserve(total_px * per_px + 256 * 8));
    unsigned char* base = ctx->maps.p;
    {
        // sel first (8-byte alignment), then the float maps
        unsigned long long* sel = reinterpret_cast<unsigned long long*>(base);
        float* depth = reinterpret_cast<float*>(base + total_px * 8);
        float* conf = depth + total_px;
        float* dz = conf + total_px;
        float* normal = dz + 2 * total_px;
        unsigned* slots = reinterpret_cast<unsigned*>(normal + 3 * total_px);
        for (int j = 0; j < n_refs; ++j) {
            jobs[j].sel = sel + px_off[j];
            jobs[j].depth = depth + px_off[j];
            jobs[j].conf = conf + px_off[j];
            jobs[j].dz = dz + 2 * px_off[j];
            jobs[j].normal = normal + 3 * px_off[j];
            jobs[j].slots = slots + px_off[j];
        }
        CK(cudaMemsetAsync(base, 0, total_px * (per_px - 4), st));               // sel, depth, conf, dz, normal = 0
        CK(cudaMemsetAsync(slots, 0xFF, total_px * 4, st));
    }
    // Frontier capacity: a round holds at most one running entry per pixel plus the carried losers; 2 entries per pixel
    // was never approached (peak on the BASELINE scenes: 0.2 per pixel).  Exceeding it fails the call (ERR_OVERFLOW).
    const size_t cap = std::max<size_t>(std::max<size_t>(2 * total_px, seeds.size()), 1u << 16);
    const bool thresholded = s->frontier_band > 0.f || s->frontier_topk > 0;
    CK(ctx->ent_a.reserve(cap));
    CK(ctx->ent_b.reserve(cap));
    CK(ctx->run_in.reserve(cap));
    CK(ctx->run_sorted.reserve(cap));
    CK(ctx->tile_cnt.reserve(n_tiles));
    CK(ctx->tile_off.reserve(n_tiles));
    CK(cudaMemsetAsync(ctx->tile_cnt.p, 0, sizeof(unsigned) * n_tiles, st));
    CK(ctx->run_out.reserve(cap));
    CK(ctx->written.reserve(cap));
    CK(ctx->counters.reserve(C_NUM + n_refs));
    CK(ctx->d_jobs.reserve(n_refs));
    CK(ctx->d_settings.reserve(1));
    CK(ctx->ctl.reserve(1));
    CK(ctx->thr_bin.reserve(n_refs));
    CK(ctx->job_cancel.reserve(n_refs));
    CK(ctx->job_run.reserve(n_refs));
    CK(cudaMemsetAsync(ctx->job_run.p, 0, sizeof(unsigned long long) * n_refs, st));
    if (thresholded) CK(ctx->hist.reserve((size_t)n_refs * HIST_PER_JOB));
    if ((size_t)(C_NUM + n_refs) > 4096) return fail(ctx, B200MVS_ERR_INVALID_ARG, "too many reference views in one batch");
    const DevSettings ds = to_dev(*s);
    CK(cudaMemcpyAsync(ctx->d_jobs.p, jobs.data(), sizeof(JobParams) * n_refs, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_settings.p, &ds, sizeof(ds), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->counters.p, 0, sizeof(unsigned long long) * (C_NUM + n_refs), st));
    CK(cudaMemsetAsync(ctx->ctl.p, 0, sizeof(FrontierCtl), st));
    if (thresholded) CK(cudaMemsetAsync(ctx->hist.p, 0, sizeof(unsigned) * (size_t)n_refs * HIST_PER_JOB, st));
    if (!seeds.empty()) CK(cudaMemcpyAsync(ctx->run_in.p, seeds.data(), sizeof(Entry) * seeds.size(), cudaMemcpyHostToDevice, st));
    HostMirror* mirror = reinterpret_cast<HostMirror*>(ctx->h_mirror);
    std::memset(ctx->h_mirror, 0, sizeof(HostMirror) + (sizeof(unsigned long long) + sizeof(int)) * n_refs);
    volatile unsigned long long* m_filled = reinterpret_cast<volatile unsigned long long*>(mirror + 1);
    volatile int* m_cancel_job = reinterpret_cast<volatile int*>(m_filled + n_refs);
    CK(cudaMemsetAsync(ctx->job_cancel.p, 0, sizeof(int) * n_refs, st));

    FrontierParams P;
    P.list[0] = ctx->ent_a.p; P.list[1] = ctx->ent_b.p;
    P.run = ctx->run_in.p; P.res = ctx->run_out.p; P.written = ctx->written.p;
    P.run2 = ctx->run_sorted.p; P.tile_cnt = ctx->tile_cnt.p; P.tile_off = ctx->tile_off.p;
    P.n_tiles = std::getenv("B200MVS_NO_TILE_SORT") ? 0 : (long long)n_tiles;
    P.cap = cap; P.n_seeds = (int)seeds.size(); P.n_jobs = n_refs;
    P.st = ctx->d_settings.p; P.jobs = ctx->d_jobs.p; P.views = ctx->d_views; P.lut = ctx->d_lut;
    P.counters = ctx->counters.p; P.ctl = ctx->ctl.p; P.hist = ctx->hist.p; P.thr_bin = ctx->thr_bin.p;
    P.host = mirror; P.host_filled = m_filled; P.host_cancel_job = m_cancel_job; P.job_cancel = ctx->job_cancel.p; P.job_run = ctx->job_run.p;
    P.thread_min = ctx->thread_min >= 0 ? ctx->thread_min : OPT_THREAD_MIN;
    if (const char* e = std::getenv("B200MVS_THREAD_MIN")) P.thread_min = std::atoll(e);      // tuning knob (tools/kbench.py)
    P.band_bins = s->frontier_band > 0.f ? std::max(1, (int)(s->frontier_band * (float)HIST_FINE)) : 0;
    P.topk = (int)std::min<uint32_t>(s->frontier_topk, 1u << 30);

    // `if (progress.cancelled) return` at the head of every stage (dmrecon.cc:100-104,336): views cancelled before the launch
    // never start; when every view is cancelled nothing runs at all
    std::vector<char> job_cancelled(n_refs, 0);
    if (progress) {
        int n_c = 0;
        for (int j = 0; j < n_refs; ++j) if (progress[j].cancelled) { job_cancelled[j] = 1; m_cancel_job[j] = 1; ++n_c; }
        if (n_c == n_refs) {
            for (int k = 0; k < n_refs; ++k) progress[k].status = 5;
            return fail(ctx, B200MVS_ERR_CANCELLED, "reconstruction cancelled");
        }
    }
    // ---- one cooperative launch: seeds + all frontier rounds (DESIGN.md "Frontier schedule") ----
    if ((rc = prepare_kernels(ctx))) return rc;
    cudaEvent_t ev_begin = get_event(ctx, 0), ev_end = get_event(ctx, 1);
    CK(cudaEventRecord(ev_begin, st));
    {
        void* args[] = {(void*)&P};
        CK(cudaLaunchCooperativeKernel((const void*)k_frontier, dim3(ctx->frontier_grid), dim3(OPT_TPB), args, OPT_SMEM_BYTES, st));
    }
    CK(cudaEventRecord(ev_end, st));
    CK(cudaMemcpyAsync(ctx->h_counters, ctx->counters.p, sizeof(unsigned long long) * (C_NUM + n_refs), cudaMemcpyDeviceToHost, st));
    FrontierCtl* h_ctl = reinterpret_cast<FrontierCtl*>(ctx->h_counters + 2048);
    CK(cudaMemcpyAsync(h_ctl, ctx->ctl.p, sizeof(FrontierCtl), cudaMemcpyDeviceToHost, st));
    cudaEvent_t ev_copied = get_event(ctx, 2);
    CK(cudaEventRecord(ev_copied, st));
    // While the kernel runs the host only relays: progress out (Progress::filled / queueSize, fancy_progress_printer.cc:84-91)
    // and cancel requests in (imageoperations.cc:177-184): a cancelled view's queue is dropped at the next round, the other
    // views of the batch go on; when every view is cancelled the kernel stops.
    if (progress) {
        for (;;) {
            const cudaError_t q = cudaEventQuery(ev_copied);
            if (q == cudaSuccess) break;
            if (q != cudaErrorNotReady) return fail(ctx, B200MVS_ERR_CUDA, "frontier kernel: %s", cudaGetErrorString(q));
            const unsigned long long qs = mirror->queue;
            int n_c = 0;
            for (int j = 0; j < n_refs; ++j) {
                if (progress[j].cancelled) { job_cancelled[j] = 1; m_cancel_job[j] = 1; }
                if (job_cancelled[j]) { ++n_c; continue; }
                progress[j].status = 3;
                progress[j].filled = m_filled[j];
                progress[j].queue_size = qs;
            }
            if (n_c == n_refs) mirror->cancel = 1;
            std::this_thread::sleep_for(std::chrono::microseconds(200));
        }
    }
    CK(cudaStreamSynchronize(st));
    if (progress) for (int j = 0; j < n_refs; ++j) if (progress[j].cancelled) job_cancelled[j] = 1;
    const bool cancelled = h_ctl->stop == ST_CANCELLED || std::all_of(job_cancelled.begin(), job_cancelled.end(), [](char c) { return c != 0; });
    if (ctx->h_counters[C_OVERFLOW] || h_ctl->stop == ST_OVERFLOW)
        return fail(ctx, B200MVS_ERR_OVERFLOW, "frontier buffer overflow (capacity %zu entries)", cap);
    if (stats) stats->n_seeds_success = ctx->h_counters[C_SEED_OK];

    // ---- results ----
    if (maps && !cancelled) {
        std::vector<unsigned> hslots;
        for (int j = 0; j < n_refs; ++j) {
            const size_t np = (size_t)jobs[j].W * jobs[j].H;
            maps[j].width = jobs[j].W; maps[j].height = jobs[j].H;
            if (job_cancelled[j]) continue;                  // RECON_CANCELLED: nothing is saved (dmrecon.cc:100-104)
            if (progress) progress[j].status = 4;
            if (maps[j].depth) CK(cudaMemcpyAsync(maps[j].depth, jobs[j].depth, np * 4, cudaMemcpyDeviceToHost, st));
            if (maps[j].conf) CK(cudaMemcpyAsync(maps[j].conf, jobs[j].conf, np * 4, cudaMemcpyDeviceToHost, st));
            if (maps[j].dz) CK(cudaMemcpyAsync(maps[j].dz, jobs[j].dz, np * 8, cudaMemcpyDeviceToHost, st));
            if (maps[j].normal) CK(cudaMemcpyAsync(maps[j].normal, jobs[j].normal, np * 12, cudaMemcpyDeviceToHost, st));
            if (maps[j].view_ids) {
                hslots.resize(np);
                CK(cudaMemcpyAsync(hslots.data(), jobs[j].slots, np * 4, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                const std::vector<int>& g = gsels[j];
                for (size_t p = 0; p < np; ++p)
                    for (int k = 0; k < 4; ++k) {
                        const unsigned sl = (hslots[p] >> (8 * k)) & 0xFF;
                        maps[j].view_ids[4 * p + k] = (sl < g.size()) ? g[sl] : -1;
                    }
            }
        }
    }
    CK(cudaStreamSynchronize(st));
    uint64_t filled = 0;
    for (int j = 0; j < n_refs; ++j) {
        filled += ctx->h_counters[C_NUM + j];
        if (progress) { progress[j].filled = ctx->h_counters[C_NUM + j]; progress[j].queue_size = 0; progress[j].status = (cancelled || job_cancelled[j]) ? 5 : 0; }
    }
    if (stats) {
        stats->n_opt = ctx->h_counters[C_OPTS];
        stats->n_sample_sets = ctx->h_counters[C_SETS];
        stats->n_rounds = h_ctl->rounds;
        stats->n_filled = filled;
        stats->n_entries_peak = h_ctl->peak;
        stats->n_patch_launches = 1;
        stats->n_kernel_launches = 1;
        float ms_all = 0.f;
        cudaEventElapsedTime(&ms_all, ev_begin, ev_end);
        // the optimise phases inside the persistent kernel, by %globaltimer of CTA 0 between the grid barriers
        double ns_all = 0.0;
        for (int k = 0; k < PH_NUM; ++k) ns_all += (double)h_ctl->ns[k];
        const double ns_opt = (double)h_ctl->ns[PH_OPT] + (double)h_ctl->ns[PH_SEED] + (double)h_ctl->ns[PH_OPT_THREAD];
        stats->ms_patch_kernel = ms_all;
        stats->ms_total_device = ms_all;
        stats->ms_optimise_phases = ns_all > 0.0 ? ms_all * ns_opt / ns_all : 0.0;
        stats->n_grid_barriers = h_ctl->barriers;
        if (std::getenv("B200MVS_PHASE_LOG")) {
            static const char* names[PH_NUM] = {"seed", "select", "threshold", "pick", "opt_warp", "commit", "expand", "sort", "opt_thread", "-"};
            std::fprintf(stderr, "[b200mvs] kernel %.2f ms, %llu rounds, %llu barriers:", ms_all, (unsigned long long)h_ctl->rounds, (unsigned long long)h_ctl->barriers);
            for (int k = 0; k < PH_NUM; ++k) std::fprintf(stderr, " %s=%.2f", names[k], ns_all > 0.0 ? ms_all * (double)h_ctl->ns[k] / ns_all : 0.0);
            std::fprintf(stderr, " | thread-phase warp utilisation %.1f %%\n", h_ctl->ns[PH_OPT_THREAD] ? 100.0 * (double)h_ctl->thread_busy_ns / ((double)h_ctl->ns[PH_OPT_THREAD] * (double)(ctx->frontier_grid * OPT_WARPS)) : 0.0);
        }
        stats->ms_optimise_thread_phases = ns_all > 0.0 ? ms_all * (double)h_ctl->ns[PH_OPT_THREAD] / ns_all : 0.0;
        stats->ms_sort_phases = ns_all > 0.0 ? ms_all * (double)h_ctl->ns[PH_SORT] / ns_all : 0.0;
    }
    if (cancelled) return fail(ctx, B200MVS_ERR_CANCELLED, "reconstruction cancelled");
    return 0;
}

} // extern "C"
