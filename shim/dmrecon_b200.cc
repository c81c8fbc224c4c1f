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
//                              same log lines, Progress kept up to date, cancellation -> RECON_CANCELLED
#include <cstring>
#include <ctime>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <vector>

#include "dmrecon/dmrecon.h"
#include "dmrecon/settings.h"
#include "mve/image.h"
#include "mve/image_tools.h"
#include "util/string_utils.h"

#include "b200mvs.h"

namespace {

// One device context per (scene, embedding), shared by all DMRecon objects of the process - the reference shares its
// image pyramids the same way through ImagePyramidCache's statics (image_pyramid.cc:157-160).
struct SharedCtx {
    std::mutex mtx;
    mve::Scene::Ptr scene;
    std::string embedding;
    b200mvs_ctx* ctx = nullptr;
    std::vector<char> uploaded;
    bool features_set = false;
    ~SharedCtx() { if (ctx) b200mvs_destroy(ctx); }
};
SharedCtx g_shared;

int pick_device()
{
    const char* e = std::getenv("B200MVS_DEVICE");
    return e ? std::atoi(e) : 0;
}

void throw_for(int rc, b200mvs_ctx* ctx)
{
    const std::string msg = b200mvs_last_error(ctx);
    if (rc == B200MVS_ERR_INVALID_ARG || rc == B200MVS_ERR_UNSUPPORTED) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);     // B200MVS_ERR_GLOBAL_VS ("Global View Selection failed"), CUDA errors, overflow
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
    std::unique_lock<std::mutex> lock(g_shared.mtx);

    /* (Re)create the device context for this scene. */
    if (g_shared.ctx == nullptr || g_shared.scene != scene || g_shared.embedding != settings.imageEmbedding) {
        if (g_shared.ctx) { b200mvs_destroy(g_shared.ctx); g_shared.ctx = nullptr; }
        int rc = b200mvs_create(pick_device(), (int)mve_views.size(), &g_shared.ctx);
        if (rc != 0) throw std::runtime_error(b200mvs_last_error(nullptr));
        g_shared.scene = scene;
        g_shared.embedding = settings.imageEmbedding;
        g_shared.uploaded.assign(mve_views.size(), 0);
        g_shared.features_set = false;
    }
    b200mvs_ctx* ctx = g_shared.ctx;

    /* Views: the same validity test as dmrecon.cc:62-71; images are uploaded once and their pyramids cached. */
    progress.status = RECON_FEATURES;
    for (std::size_t i = 0; i < mve_views.size() && !progress.cancelled; ++i) {
        if (g_shared.uploaded[i]) continue;
        mve::View::Ptr v = mve_views[i];
        if (v == nullptr || !v->is_camera_valid() || !v->has_image(settings.imageEmbedding, mve::IMAGE_TYPE_UINT8))
            continue;
        mve::ByteImage::Ptr img = v->get_byte_image(settings.imageEmbedding);
        mve::CameraInfo const& cam = v->get_camera();
        int rc = b200mvs_upload_view(ctx, (int)i, img->get_data_pointer(), img->width(), img->height(), img->channels(),
            cam.flen, cam.paspect, cam.ppoint, cam.rot, cam.trans);
        v->cache_cleanup();
        if (rc != 0) throw_for(rc, ctx);
        g_shared.uploaded[i] = 1;
    }
    if (!g_shared.features_set) {
        mve::Bundle::Features const& features = bundle->get_features();
        std::vector<float> pos(features.size() * 3);
        std::vector<int32_t> off(features.size() + 1, 0), ids;
        for (std::size_t i = 0; i < features.size(); ++i) {
            std::memcpy(&pos[3 * i], features[i].pos, 3 * sizeof(float));
            for (std::size_t j = 0; j < features[i].refs.size(); ++j) ids.push_back(features[i].refs[j].view_id);
            off[i + 1] = (int32_t)ids.size();
        }
        int rc = b200mvs_set_features(ctx, (int)features.size(), pos.data(), off.data(), ids.data());
        if (rc != 0) throw_for(rc, ctx);
        g_shared.features_set = true;
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

    if (!settings.quiet) {
        int32_t ids[B200MVS_MAX_GLOBAL_VIEWS];
        int n = b200mvs_global_view_selection(ctx, &s, (int)settings.refViewNr, ids, B200MVS_MAX_GLOBAL_VIEWS);
        if (n > 0) {
            std::cout << "Global View Selection:";
            for (int i = 0; i < n; ++i) std::cout << " " << ids[i];
            std::cout << std::endl;
        }
    }

    /* Result images, allocated like SingleView::prepareMasterView (single_view.cc:78-81). */
    mve::FloatImage::Ptr depthImg = mve::FloatImage::create(width, height, 1);
    mve::FloatImage::Ptr dzImg = mve::FloatImage::create(width, height, 2);
    mve::FloatImage::Ptr confImg = mve::FloatImage::create(width, height, 1);
    b200mvs_maps maps;
    std::memset(&maps, 0, sizeof(maps));
    maps.depth = depthImg->get_data_pointer();
    maps.dz = dzImg->get_data_pointer();
    maps.conf = confImg->get_data_pointer();
    b200mvs_progress prog;
    std::memset(&prog, 0, sizeof(prog));
    prog.cancelled = progress.cancelled ? 1 : 0;
    b200mvs_stats stats;
    int32_t ref = (int32_t)settings.refViewNr, failed = -1;
    progress.status = RECON_QUEUE;
    int rc = b200mvs_reconstruct(ctx, &s, 1, &ref, &maps, &prog, &stats, &failed);
    progress.filled = prog.filled;
    progress.queueSize = 0;
    if (rc == B200MVS_ERR_CANCELLED || progress.cancelled) { progress.status = RECON_CANCELLED; return; }
    if (rc != 0) throw_for(rc, ctx);
    if (!settings.quiet)
        std::cout << "Processed " << stats.n_seeds_processed << " features, from which "
                  << stats.n_seeds_success << " succeeded optimization." << std::endl;

    progress.status = RECON_SAVING;
    mve::View::Ptr view = mve_views[settings.refViewNr];
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
        mve::ByteImage::Ptr scaled = mve::ByteImage::create(width, height, 3);
        int w = 0, h = 0;
        rc = b200mvs_get_level(ctx, (int)settings.refViewNr, settings.scale, &w, &h, scaled->get_data_pointer());
        if (rc != 0) throw_for(rc, ctx);
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
