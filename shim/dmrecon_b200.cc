// Drop-in replacement for the reference's libmve_dmrecon.a: the class mvs::DMRecon with the reference's OWN header
// (libs/dmrecon/dmrecon.h:40-68, included from the reference tree at build time - nothing is copied into this repo),
// implemented as a thin host shim over the C ABI of libb200mvs.so (include/b200mvs.h).
//
// apps/dmrecon/dmrecon.cc and fancy_progress_printer.* compile and link UNCHANGED against this (shim/Makefile):
//   mvs::DMRecon recon(scene, settings); recon.start(); recon.getProgress(); recon.getRefViewNr();
// Everything below the boundary that is I/O stays the reference's (mve::Scene / mve::View / mve::Bundle, libmve.a).
//
// What the shim does, mirroring dmrecon.cc:
//   ctor  (dmrecon.cc:30-87)   same argument validation and exception types/messages; width/height of the scaled image
//   start (dmrecon.cc:90-172)  views + bundle -> b200mvs context (cached per scene like ImagePyramidCache,
//                              image_pyramid.cc:99-132), b200mvs_reconstruct, results attached with View::set_image
//                              under the reference's embedding names (depth-L<s>, dz-L<s>, conf-L<s>, undist-L<s>),
//                              writePlyFile / plyPath through the reference's own save_ply_view, same log lines, Progress
//                              updated live while the kernel runs, cancellation of a running view -> RECON_CANCELLED
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <ctime>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <thread>
#include <vector>

#include "dmrecon/dmrecon.h"
#include "dmrecon/settings.h"
#include "mve/image.h"
#include "mve/image_tools.h"
#include "mve/mesh_io_ply.h"
#include "util/file_system.h"
#include "util/string_utils.h"

#include "b200mvs.h"

namespace {

// One device context per (scene, embedding) AND per GPU, shared by all DMRecon objects of the process - the reference
// shares its image pyramids the same way through ImagePyramidCache's statics (image_pyramid.cc:157-160).
//
// The reference driver runs DMRecon::start() concurrently from OpenMP threads (apps/dmrecon/dmrecon.cc:285).  Here the
// concurrent calls are (i) spread round-robin over the GPUs named by B200MVS_DEVICES (default: device B200MVS_DEVICE or
// 0) and (ii) per GPU COMBINED: every caller only ENQUEUES its request; one of them becomes the leader, waits a short
// collection window for the other threads of the OpenMP team to arrive, then does for the whole batch what each DMRecon
// does for itself in the reference - global view selection, loading the colour images that are needed (once per view
// and context) - and submits ONE b200mvs_reconstruct in which all views advance together.  While the kernel runs a
// relay thread copies the live progress into every caller's mvs::Progress and forwards cancel requests.
struct Request {
    int32_t ref = 0;
    b200mvs_settings settings;
    b200mvs_maps maps;
    mvs::Progress* progress = nullptr;      // the caller's DMRecon::progress (read by progress printers / UMVE while we run)
    std::string embedding;
    bool quiet = true;
    b200mvs_stats stats;
    std::vector<int32_t> gvs;
    int rc = 0;
    std::string err;
    bool done = false;
};

struct DeviceCtx {
    std::mutex mtx;                 // protects everything below
    std::condition_variable cv;
    int device = 0;
    mve::Scene::Ptr scene;
    std::string embedding;
    b200mvs_ctx* ctx = nullptr;
    std::vector<char> uploaded;
    bool features_set = false;
    bool cameras_set = false;
    bool leader_active = false;
    int planners = 0;               // callers that run their global view selection right now and will enqueue next
    std::vector<Request*> pending;
    uint64_t arrivals = 0;          // bumped by every enqueue: the leader's collection window watches it
    size_t last_batch = 0;          // size of the previous batch = how many callers to expect
    ~DeviceCtx() { if (ctx) b200mvs_destroy(ctx); }
};

std::mutex g_mtx;
std::vector<std::unique_ptr<DeviceCtx>> g_devices;
std::atomic<unsigned> g_next(0);
std::mutex g_cout;

std::vector<int> device_list()
{
    std::vector<int> out;
    if (const char* e = std::getenv("B200MVS_DEVICES")) {
        std::string s(e);
        size_t pos = 0;
        while (pos < s.size()) {
            size_t q = s.find(',', pos);
            if (q == std::string::npos) q = s.size();
            if (q > pos) out.push_back(std::atoi(s.substr(pos, q - pos).c_str()));
            pos = q + 1;
        }
    }
    if (out.empty()) {
        const char* e = std::getenv("B200MVS_DEVICE");
        out.push_back(e ? std::atoi(e) : 0);
    }
    return out;
}

DeviceCtx& pick_device_ctx()
{
    std::lock_guard<std::mutex> lk(g_mtx);
    if (g_devices.empty())
        for (int d : device_list()) { g_devices.emplace_back(new DeviceCtx()); g_devices.back()->device = d; }
    return *g_devices[g_next.fetch_add(1) % g_devices.size()];
}

void throw_for(int rc, const std::string& msg)
{
    if (rc == B200MVS_ERR_INVALID_ARG || rc == B200MVS_ERR_UNSUPPORTED) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);     // B200MVS_ERR_GLOBAL_VS ("Global View Selection failed"), CUDA errors, overflow
}

bool same_settings(const b200mvs_settings& a, const b200mvs_settings& b) { return std::memcmp(&a, &b, sizeof(a)) == 0; }

// Global view selection + colour images of one request (dmrecon.cc:211-241), done by the batch leader.
// Called WITHOUT D.mtx (D.uploaded is only touched by the leader).
void prepare_request(DeviceCtx& D, Request& r)
{
    mve::Scene::ViewList const& mve_views(D.scene->get_views());
    r.progress->status = mvs::RECON_GLOBALVS;
    int32_t ids[B200MVS_MAX_GLOBAL_VIEWS];
    const int n = b200mvs_global_view_selection(D.ctx, &r.settings, r.ref, ids, B200MVS_MAX_GLOBAL_VIEWS);
    if (n < 0) { r.rc = n; r.err = b200mvs_last_error(D.ctx); return; }
    if (n == 0) { r.rc = B200MVS_ERR_GLOBAL_VS; r.err = "Global View Selection failed"; return; }
    r.gvs.assign(ids, ids + n);
    if (!r.quiet) {
        std::lock_guard<std::mutex> lk(g_cout);
        std::cout << "Global View Selection:";
        for (int i = 0; i < n; ++i) std::cout << " " << ids[i];
        std::cout << std::endl << "Loading color images..." << std::endl;
    }
    std::vector<int> need(ids, ids + n);
    need.push_back(r.ref);
    for (int id : need) {
        if (r.progress->cancelled) return;
        if (D.uploaded[id]) continue;
        mve::View::Ptr v = mve_views[id];
        mve::ByteImage::Ptr img = v->get_byte_image(r.embedding);
        mve::CameraInfo const& cam = v->get_camera();
        const int rc = b200mvs_upload_view(D.ctx, id, img->get_data_pointer(), img->width(), img->height(), img->channels(),
            cam.flen, cam.paspect, cam.ppoint, cam.rot, cam.trans);
        v->cache_cleanup();
        if (rc != 0) { r.rc = rc; r.err = b200mvs_last_error(D.ctx); return; }
        D.uploaded[id] = 1;
    }
    r.progress->status = mvs::RECON_FEATURES;
}

// Runs one batch on the device context (called by the leader WITHOUT holding D.mtx).
void run_batch(DeviceCtx& D, std::vector<Request*> batch)
{
    for (Request* r : batch) prepare_request(D, *r);
    // requests that failed in preparation or were cancelled meanwhile leave the batch with their own result
    std::vector<Request*> live;
    for (Request* r : batch) {
        if (r->rc != 0) continue;
        if (r->progress->cancelled) { r->rc = B200MVS_ERR_CANCELLED; continue; }
        live.push_back(r);
    }
    while (!live.empty()) {
        const size_t n = live.size();
        std::vector<int32_t> refs(n);
        std::vector<b200mvs_maps> maps(n);
        std::vector<b200mvs_progress> prog(n);
        std::memset(prog.data(), 0, sizeof(b200mvs_progress) * n);
        for (size_t i = 0; i < n; ++i) { refs[i] = live[i]->ref; maps[i] = live[i]->maps; live[i]->progress->status = mvs::RECON_QUEUE; }
        // relay: live progress out, cancel requests in (Progress is read/written without locks in the reference too,
        // fancy_progress_printer.cc:84-91, apps/umve/viewinspect/imageoperations.cc:177-184)
        std::atomic<bool> stop(false);
        std::thread relay([&]() {
            while (!stop.load()) {
                for (size_t i = 0; i < n; ++i) {
                    live[i]->progress->filled = prog[i].filled;
                    live[i]->progress->queueSize = prog[i].queue_size;
                    if (live[i]->progress->cancelled) prog[i].cancelled = 1;
                }
                std::this_thread::sleep_for(std::chrono::milliseconds(2));
            }
        });
        b200mvs_stats stats;
        int32_t failed = -1;
        const int rc = b200mvs_reconstruct(D.ctx, &live[0]->settings, (int)n, refs.data(), maps.data(), prog.data(), &stats, &failed);
        stop = true;
        relay.join();
        if (rc == 0 || rc == B200MVS_ERR_CANCELLED) {
            for (size_t i = 0; i < n; ++i) {
                // a cancelled view ends as RECON_CANCELLED; the other views of the batch keep their results
                const bool cancelled = rc == B200MVS_ERR_CANCELLED || prog[i].status == 5;
                live[i]->rc = cancelled ? B200MVS_ERR_CANCELLED : 0;
                live[i]->progress->filled = prog[i].filled;
                live[i]->maps = maps[i]; live[i]->stats = stats;
                live[i]->err = cancelled ? "reconstruction cancelled" : "";
            }
            return;
        }
        // one view made the call fail: give it its error, retry the others
        const std::string msg = b200mvs_last_error(D.ctx);
        bool removed = false;
        for (size_t i = 0; i < live.size(); ++i) {
            if (failed >= 0 && live[i]->ref != failed) continue;
            live[i]->rc = rc; live[i]->err = msg;
            if (failed >= 0) { live.erase(live.begin() + i); removed = true; break; }
        }
        if (failed < 0 || !removed) return;      // error not attributable to one view: every request got it
    }
}

} // namespace

MVS_NAMESPACE_BEGIN

DMRecon::DMRecon(mve::Scene::Ptr _scene, Settings const& _settings)
    : scene(_scene)
    , settings(_settings)
{
    mve::Scene::ViewList const& mve_views(scene->get_views());
    if (settings.refViewNr >= mve_views.size())
        throw std::invalid_argument("Master view index out of bounds");
    if (settings.scale < 0.f)
        throw std::invalid_argument("Invalid scale factor");
    if (settings.imageEmbedding.empty())
        throw std::invalid_argument("Invalid image embedding");
    try {
        this->bundle = this->scene->get_bundle();
    } catch (std::exception& e) {
        throw std::runtime_error(std::string("Error reading bundle file: ") + e.what());
    }
    mve::View::Ptr refV = mve_views[settings.refViewNr];
    if (refV == nullptr || !refV->is_camera_valid()
        || !refV->has_image(settings.imageEmbedding, mve::IMAGE_TYPE_UINT8))
        throw std::invalid_argument("Invalid master view");
    // size of pyramid level `scale` ((w+1)/2 per level, image_pyramid.cc:46-47)
    mve::View::ImageProxy const* proxy = refV->get_image_proxy(settings.imageEmbedding);
    int w = proxy->width, h = proxy->height;
    for (int l = 0; l < settings.scale; ++l) { w = (w + 1) / 2; h = (h + 1) / 2; }
    this->width = w;
    this->height = h;
    if (!settings.quiet)
        std::cout << "scaled image size: " << this->width << " x " << this->height << std::endl;
}

void
DMRecon::start()
{
    progress.start_time = std::time(nullptr);
    mve::Scene::ViewList const& mve_views(scene->get_views());
    DeviceCtx& D = pick_device_ctx();
    std::unique_lock<std::mutex> lock(D.mtx);

    /* (Re)create the device context for this scene; cameras and features are registered once per context. */
    if (D.ctx == nullptr || D.scene != scene || D.embedding != settings.imageEmbedding) {
        while (D.leader_active || D.planners > 0) D.cv.wait(lock);
        if (D.ctx) { b200mvs_destroy(D.ctx); D.ctx = nullptr; }
        int rc = b200mvs_create(D.device, (int)mve_views.size(), &D.ctx);
        if (rc != 0) throw std::runtime_error(b200mvs_last_error(nullptr));
        D.scene = scene;
        D.embedding = settings.imageEmbedding;
        D.uploaded.assign(mve_views.size(), 0);
        D.features_set = false;
        D.cameras_set = false;
    }
    b200mvs_ctx* ctx = D.ctx;

    /* Views: the same validity test as dmrecon.cc:62-71.  Every valid view gets its camera registered
       (SingleView::create); colour images are loaded by the batch leader, only for the master views and their selected
       neighbours (loadColorImage, dmrecon.cc:78,238-240), once per view and context. */
    progress.status = RECON_FEATURES;
    if (!D.cameras_set) {
        while (D.leader_active || D.planners > 0) D.cv.wait(lock);
        for (std::size_t i = 0; i < mve_views.size(); ++i) {
            mve::View::Ptr v = mve_views[i];
            if (v == nullptr || !v->is_camera_valid() || !v->has_image(settings.imageEmbedding, mve::IMAGE_TYPE_UINT8))
                continue;
            mve::View::ImageProxy const* proxy = v->get_image_proxy(settings.imageEmbedding);
            mve::CameraInfo const& cam = v->get_camera();
            int rc = b200mvs_set_view_camera(ctx, (int)i, proxy->width, proxy->height, cam.flen, cam.paspect, cam.ppoint,
                cam.rot, cam.trans);
            if (rc != 0) throw_for(rc, b200mvs_last_error(ctx));
        }
        D.cameras_set = true;
    }
    if (!D.features_set) {
        mve::Bundle::Features const& features = bundle->get_features();
        std::vector<float> pos(features.size() * 3);
        std::vector<int32_t> off(features.size() + 1, 0), ids;
        for (std::size_t i = 0; i < features.size(); ++i) {
            std::memcpy(&pos[3 * i], features[i].pos, 3 * sizeof(float));
            for (std::size_t j = 0; j < features[i].refs.size(); ++j) ids.push_back(features[i].refs[j].view_id);
            off[i + 1] = (int32_t)ids.size();
        }
        while (D.leader_active || D.planners > 0) D.cv.wait(lock);
        int rc = b200mvs_set_features(ctx, (int)features.size(), pos.data(), off.data(), ids.data());
        if (rc != 0) throw_for(rc, b200mvs_last_error(ctx));
        D.features_set = true;
    }
    if (progress.cancelled) { progress.status = RECON_CANCELLED; return; }

    /* Settings: POD part of mvs::Settings, field for field. */
    b200mvs_settings s;
    b200mvs_default_settings(&s);
    s.filter_width = settings.filterWidth;
    s.min_ncc = settings.minNCC;
    s.min_parallax = settings.minParallax;
    s.accept_ncc = settings.acceptNCC;
    s.min_refine_diff = settings.minRefineDiff;
    s.max_iterations = settings.maxIterations;
    s.nr_recon_neighbors = settings.nrReconNeighbors;
    s.global_vs_max = settings.globalVSMax;
    s.scale = settings.scale;
    s.use_color_scale = settings.useColorScale ? 1 : 0;
    for (int i = 0; i < 3; ++i) { s.aabb_min[i] = settings.aabbMin[i]; s.aabb_max[i] = settings.aabbMax[i]; }
    if (const char* e = std::getenv("B200MVS_FRONTIER_TOPK")) s.frontier_topk = (uint32_t)std::atoi(e);   // engine knobs, include/b200mvs.h
    if (const char* e = std::getenv("B200MVS_FRONTIER_BAND")) s.frontier_band = (float)std::atof(e);

    /* Result images, allocated like SingleView::prepareMasterView (single_view.cc:78-81). */
    mve::FloatImage::Ptr depthImg = mve::FloatImage::create(width, height, 1);
    mve::FloatImage::Ptr dzImg = mve::FloatImage::create(width, height, 2);
    mve::FloatImage::Ptr confImg = mve::FloatImage::create(width, height, 1);
    Request req;
    req.ref = (int32_t)settings.refViewNr;
    req.settings = s;
    req.progress = &progress;
    req.embedding = settings.imageEmbedding;
    req.quiet = settings.quiet;
    std::memset(&req.maps, 0, sizeof(req.maps));
    req.maps.depth = depthImg->get_data_pointer();
    req.maps.dz = dzImg->get_data_pointer();
    req.maps.conf = confImg->get_data_pointer();
    std::memset(&req.stats, 0, sizeof(req.stats));

    /* Global view selection + seed list (dmrecon.cc:96-98,211-292) on the CALLER's thread, like the reference, where every
       DMRecon of the OpenMP team does its own: in parallel over the team and while the GPU still runs the previous batch.
       The result is parked in the context (b200mvs_plan_views); the leader's b200mvs_global_view_selection and
       b200mvs_reconstruct pick it up instead of computing it again. */
    progress.status = RECON_GLOBALVS;
    D.planners++;
    lock.unlock();
    b200mvs_plan_views(ctx, &s, 1, &req.ref);        // a failure shows up again, with its message, in the leader's call
    lock.lock();
    D.planners--;
    if (progress.cancelled) { D.cv.notify_all(); progress.status = RECON_CANCELLED; return; }

    /* Submit.  Whoever finds no leader becomes one and runs batches until the queue is empty. */
    D.pending.push_back(&req);
    D.arrivals++;
    D.cv.notify_all();
    if (!D.leader_active) {
        D.leader_active = true;
        while (!D.pending.empty()) {
            /* Collection window: the other threads of the caller's OpenMP team reach this point within microseconds to
               milliseconds of each other (they all finished the previous batch together).  Callers that are computing
               their view selection right now (D.planners) are certain to arrive: wait for them (at most 250 ms).  Beyond
               that: until as many requests as the previous batch had are here, or nothing new has arrived for 3 ms, at
               most 30 ms. */
            const auto t_open = std::chrono::steady_clock::now();
            uint64_t seen = D.arrivals;
            for (;;) {
                if (D.planners == 0 && D.last_batch > 1 && D.pending.size() >= D.last_batch) break;
                const bool woke = D.cv.wait_for(lock, std::chrono::milliseconds(3), [&]() { return D.arrivals != seen; });
                if (!woke && D.planners == 0) break;
                seen = D.arrivals;
                if (std::chrono::steady_clock::now() - t_open > std::chrono::milliseconds(D.planners > 0 ? 250 : 30)) break;
            }
            std::vector<Request*> batch;
            std::vector<Request*> rest;
            for (Request* r : D.pending)
                (batch.empty() || (same_settings(r->settings, batch[0]->settings) && r->embedding == batch[0]->embedding) ? batch : rest).push_back(r);
            D.pending.swap(rest);
            D.last_batch = batch.size();
            lock.unlock();
            run_batch(D, batch);
            lock.lock();
            for (Request* r : batch) r->done = true;
            D.cv.notify_all();
        }
        D.leader_active = false;
        D.cv.notify_all();
    }
    while (!req.done) D.cv.wait(lock);
    lock.unlock();

    const b200mvs_stats& stats = req.stats;
    const int rc0 = req.rc;
    progress.queueSize = 0;
    if (rc0 == B200MVS_ERR_CANCELLED || progress.cancelled) { progress.status = RECON_CANCELLED; return; }
    if (rc0 != 0) throw_for(rc0, req.err);
    if (!settings.quiet) {
        std::ostringstream line;       // one write: the OpenMP threads of the driver print concurrently
        line << "Reconstructed view " << settings.refViewNr << " (batch of all views in flight: " << stats.n_seeds_processed
             << " features processed, " << stats.n_seeds_success << " succeeded optimization, " << stats.n_rounds
             << " frontier rounds)." << std::endl;
        std::lock_guard<std::mutex> lk(g_cout);
        std::cout << line.str() << std::flush;
    }

    progress.status = RECON_SAVING;
    mve::View::Ptr view = mve_views[settings.refViewNr];
    mve::ByteImage::Ptr scaled;
    if (settings.scale != 0 || settings.writePlyFile) {
        // level `scale` of the reference view as the device built it (bit-exact with the reference's pyramid)
        scaled = mve::ByteImage::create(width, height, 3);
        int w = 0, h = 0;
        int rc = b200mvs_get_level(ctx, (int)settings.refViewNr, settings.scale, &w, &h, scaled->get_data_pointer());
        if (rc != 0) throw_for(rc, b200mvs_last_error(ctx));
    }
    if (settings.writePlyFile) {
        // SingleView::saveReconAsPly (single_view.cc:123-138) through the same libmve writers
        if (settings.plyPath.empty()) throw std::invalid_argument("Empty path");
        std::string fname = "mvs-" + util::string::get_filled(settings.refViewNr, 4) + "-L" + util::string::get((float)settings.scale);
        if (!settings.quiet)
            std::cout << "Saving ply file as " << settings.plyPath << "/" << fname << ".ply" << std::endl;
        if (!util::fs::dir_exists(settings.plyPath.c_str())) util::fs::mkdir(settings.plyPath.c_str());
        mve::geom::save_ply_view(util::fs::join_path(settings.plyPath, fname + ".ply"), view->get_camera(), depthImg, confImg, scaled);
        mve::geom::save_xf_file(util::fs::join_path(settings.plyPath, fname + ".xf"), view->get_camera());
    }
    std::string name("depth-L");
    name += util::string::get(settings.scale);
    view->set_image(depthImg, name);
    if (settings.keepDzMap) {
        name = "dz-L";
        name += util::string::get(settings.scale);
        view->set_image(dzImg, name);
    }
    if (settings.keepConfidenceMap) {
        name = "conf-L";
        name += util::string::get(settings.scale);
        view->set_image(confImg, name);
    }
    if (settings.scale != 0) {
        name = "undist-L";
        name += util::string::get(settings.scale);
        view->set_image(scaled, name);
    }
    progress.status = RECON_IDLE;
    {
        int nrPix = this->width * this->height;
        float percent = (float) progress.filled / (float) nrPix;
        if (!settings.quiet)
            std::cout << "Filled " << progress.filled << " pixels, i.e. "
                      << util::string::get_fixed(percent * 100.f, 1) << " %." << std::endl;
    }
    size_t mvs_time = std::time(nullptr) - progress.start_time;
    if (!settings.quiet)
        std::cout << "MVS took " << mvs_time << " seconds." << std::endl;
}

MVS_NAMESPACE_END
