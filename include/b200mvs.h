/*
 * b200mvs - B200-native (sm_100a) dense multi-view-stereo depth-map engine behind the
 * interface of simonfuhrmann/mve's libs/dmrecon.
 *
 * This is the drop-in boundary: a plain C ABI (no C++/torch types) exported by
 * mve_b200/libb200mvs.so.  The reference has no FFI layer of its own for this path - its
 * boundary is the C++ class mvs::DMRecon in a static library (libs/dmrecon/dmrecon.h:40-68)
 * - so each entry point below names the reference interface it replaces; INTEGRATION.md
 * shows the header-identical mvs::DMRecon shim a maintainer links instead of
 * libmve_dmrecon.a.  All citations are relative to the reference tree.
 *
 * Conventions: every function returns 0 on success or a negative B200MVS_ERR_* code;
 * b200mvs_last_error() gives the message.  No exception crosses this boundary; the C++ shim
 * re-throws the exception types the reference throws (SURVEY.md §8b "Errors").
 * There is NO CPU fallback: b200mvs_create fails when no CUDA device is usable.
 */
#ifndef B200MVS_H
#define B200MVS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MVS_OK               0
#define B200MVS_ERR_INVALID_ARG (-1)  /* std::invalid_argument in the reference (dmrecon.cc:37-75)        */
#define B200MVS_ERR_CUDA        (-2)  /* CUDA runtime failure -> std::runtime_error in the shim           */
#define B200MVS_ERR_GLOBAL_VS   (-3)  /* "Global View Selection failed" (dmrecon.cc:222-223)               */
#define B200MVS_ERR_CANCELLED   (-4)  /* progress.cancelled was set (dmrecon.cc:100-104, RECON_CANCELLED) */
#define B200MVS_ERR_OVERFLOW    (-5)  /* frontier buffer capacity exceeded                                 */
#define B200MVS_ERR_UNSUPPORTED (-6)  /* setting outside the range the kernels implement (see below)       */

#define B200MVS_MAX_GLOBAL_VIEWS 32   /* settings.global_vs_max must be <= 32 (reference default 20)      */
#define B200MVS_MAX_LOCAL_VIEWS  4    /* settings.nr_recon_neighbors must be 1..4 (reference default 4)   */

typedef struct b200mvs_ctx b200mvs_ctx;

/* POD part of mvs::Settings (libs/dmrecon/settings.h:22-52), field for field.
 * filter_width must be 5: the reference hard-codes patchPoints[12] (patch_sampler.cc:96). */
typedef struct b200mvs_settings {
    uint32_t filter_width;        /* settings.h:31  = 5     */
    float    min_ncc;             /* settings.h:32  = 0.3   */
    float    min_parallax;        /* settings.h:33  = 10    */
    float    accept_ncc;          /* settings.h:34  = 0.6   */
    float    min_refine_diff;     /* settings.h:35  = 0.001 */
    uint32_t max_iterations;      /* settings.h:36  = 20    */
    uint32_t nr_recon_neighbors;  /* settings.h:37  = 4     */
    uint32_t global_vs_max;       /* settings.h:38  = 20    */
    int32_t  scale;               /* settings.h:39  = 0     */
    int32_t  use_color_scale;     /* settings.h:40  = 1     */
    float    aabb_min[3];         /* settings.h:44          */
    float    aabb_max[3];         /* settings.h:45          */
    /* Engine knobs (no reference counterpart; DESIGN.md "Frontier schedule").  The reference pops strictly by
     * descending confidence (dmrecon.h:72-76); the GPU advances a whole frontier per round.  Both knobs restrict a
     * round to the most confident queued entries of each view so that the order approaches the reference's:
     *   frontier_band  > 0: only entries within this confidence distance of the view's most confident entry run;
     *   frontier_topk  > 0: only about the K most confident entries of each view run (confidence resolution 1/8192).
     * 0 / 0 (default): every queued entry runs each round (fastest). */
    float    frontier_band;
    uint32_t frontier_topk;
} b200mvs_settings;

/* Fills the defaults of settings.h:22-52. */
void b200mvs_default_settings(b200mvs_settings* s);

/* mvs::Progress (libs/dmrecon/progress.h:27-43); read/written without locks like the reference
 * (fancy_progress_printer.cc:84-91, apps/umve/viewinspect/imageoperations.cc:177-184). */
typedef struct b200mvs_progress {
    volatile int32_t  status;     /* ReconStatus: 0 idle, 1 globalvs, 2 features, 3 queue, 4 saving, 5 cancelled */
    volatile int32_t  cancelled;  /* set from outside (any thread) to cancel THIS view; relayed to the running kernel within
                                     ~0.2 ms, its queue is dropped at the next frontier round, status ends as 5 and its maps
                                     are not written; the other views of the batch go on.  The call returns
                                     B200MVS_ERR_CANCELLED only when every view of the batch was cancelled */
    volatile uint64_t filled;
    volatile uint64_t queue_size;
    volatile uint64_t start_time;
} b200mvs_progress;

/* Result maps of one reference view, caller-owned HOST buffers of width*height pixels at pyramid
 * level settings.scale, row-major - the images DMRecon::start attaches to the view
 * (dmrecon.cc:119-145): depth (1 ch), conf (1 ch), dz (2 ch), plus normal (3 ch, computed but never
 * saved by the reference, single_view.cc:78-81) and the per-pixel local view ids (4 x int32, -1
 * padded, ascending) of the optimisation that wrote the pixel.  Any pointer except depth may be
 * NULL.  width/height are outputs. */
typedef struct b200mvs_maps {
    float*   depth;
    float*   conf;
    float*   dz;
    float*   normal;
    int32_t* view_ids;
    int32_t  width, height;
} b200mvs_maps;

/* One mvs::PatchOptimization (patch_optimization.cc:21-30): inputs and results. */
typedef struct b200mvs_patch_in {
    int32_t x, y;
    float   depth, dz_i, dz_j;
    int32_t n_local;              /* propagated local view ids (0 = run the full local view selection) */
    int32_t local_ids[4];
} b200mvs_patch_in;

typedef struct b200mvs_patch_out {
    float   conf;                 /* computeConfidence() (patch_optimization.cc:114-142) */
    float   depth, dz_i, dz_j;
    float   normal[3];
    int32_t n_local;
    int32_t local_ids[4];         /* ascending view ids, -1 padded */
    int32_t iterations;           /* status.iterationCount */
    int32_t converged;
    int32_t opti_success;
} b200mvs_patch_out;

/* Device-side work counters of one call (DESIGN.md "Measurement"). */
typedef struct b200mvs_stats {
    uint64_t n_opt;               /* patch optimisations executed                               */
    uint64_t n_sample_sets;       /* fused colour+derivative 5x5 sample sets drawn              */
    uint64_t n_rounds;            /* frontier rounds                                            */
    uint64_t n_filled;            /* pixels with conf > 0 (progress.filled)                     */
    uint64_t n_seeds_processed;   /* "Processed N features" (dmrecon.cc:286)                    */
    uint64_t n_seeds_success;     /* "... from which N succeeded optimization" (dmrecon.cc:301) */
    uint64_t n_entries_peak;      /* peak frontier size                                         */
    double   ms_patch_kernel;     /* CUDA-event time of the kernel that runs the patch optimisations (reconstruct: the
                                     persistent frontier kernel, seeds + all rounds in one launch) */
    double   ms_total_device;     /* CUDA-event time first launch -> last launch of the call    */
    uint64_t n_patch_launches;    /* launches of the kernel that runs the patch optimisations   */
    uint64_t n_kernel_launches;   /* all kernel launches of the call                            */
    double   ms_optimise_phases;  /* part of ms_patch_kernel spent in the optimise phases (rest: queue bookkeeping + barriers) */
    uint64_t n_grid_barriers;     /* grid-wide barriers executed by the persistent kernel       */
    double   ms_optimise_thread_phases; /* part of ms_optimise_phases in rounds run one thread per patch       */
    double   ms_sort_phases;      /* grouping the winners of large rounds by tile               */
} b200mvs_stats;

/* ---- lifecycle (mvs::DMRecon ctor/dtor, dmrecon.cc:30-87; ImagePyramidCache, image_pyramid.cc:99-160) ---- */
/* device >= 0: a CUDA device (fails loudly when there is none).  B200MVS_DEVICE_NONE creates a PLANNING context without a
 * GPU: b200mvs_set_view_camera, b200mvs_set_features and b200mvs_global_view_selection work (they are host logic in the
 * reference too); every entry point that computes on images fails with B200MVS_ERR_CUDA - there is no CPU fallback. */
#define B200MVS_DEVICE_NONE (-1)
int  b200mvs_create(int device, int n_views, b200mvs_ctx** out);
void b200mvs_destroy(b200mvs_ctx* ctx);
const char* b200mvs_last_error(const b200mvs_ctx* ctx);   /* ctx may be NULL: last create() error of the calling thread;
                                                              the message of a ctx belongs to the last failing call on it (not
                                                              synchronised: read it from the thread that got the error code) */
const char* b200mvs_version(void);

/* ---- inputs ---- */
/* Replaces SingleView::create + SingleView::loadColorImage for one mve::View (single_view.cc:24-66,
 * image_pyramid.cc:56-95): takes the `undistorted` uint8 image (1, 2, 3 or 4 channels; alpha dropped,
 * grey expanded as image_pyramid.cc:65-73) and the mve::CameraInfo fields (camera.h:23-170), builds the
 * Gaussian pyramid on the device.  rgb is a HOST pointer, h x w x channels, row-major. */
int b200mvs_upload_view(b200mvs_ctx* ctx, int view_id, const uint8_t* rgb, int w, int h, int channels,
                        float flen, float paspect, const float ppoint[2],
                        const float rot[9], const float trans[3]);
/* Same with a DEVICE pointer (3 channels) on the ctx's device; `cuda_stream` is a cudaStream_t or NULL.
 * Used for HBM-resident inputs and after an NCCL all-gather of the images (DESIGN.md "Multi-GPU"). */
int b200mvs_upload_view_device(b200mvs_ctx* ctx, int view_id, const uint8_t* rgb_dev, int w, int h,
                               float flen, float paspect, const float ppoint[2],
                               const float rot[9], const float trans[3], void* cuda_stream);
/* Camera and image size only - SingleView::create (single_view.cc:24-53).  The reference creates a SingleView for every
 * valid view but loads colour images only for the master view and its selected neighbours (dmrecon.cc:78,238-240); a
 * caller that wants the same economy registers all cameras, asks b200mvs_global_view_selection which views are needed
 * and uploads only those images.  b200mvs_reconstruct fails with B200MVS_ERR_INVALID_ARG ("color image of view N is not
 * loaded") when a needed image is missing. */
int b200mvs_set_view_camera(b200mvs_ctx* ctx, int view_id, int w, int h, float flen, float paspect,
                            const float ppoint[2], const float rot[9], const float trans[3]);
/* mve::Bundle::Features (bundle.h:51-60) as position + CSR list of referencing view ids. */
int b200mvs_set_features(b200mvs_ctx* ctx, int n_features, const float* pos,
                         const int32_t* ref_offsets, const int32_t* ref_view_ids);

/* ---- inspection (parity of the pyramid, image_tools.h:617-694) ---- */
int b200mvs_num_levels(b200mvs_ctx* ctx, int view_id);
int b200mvs_get_level(b200mvs_ctx* ctx, int view_id, int level, int* w, int* h, uint8_t* rgb_host_or_null);

/* ---- DMRecon::analyzeFeatures + globalViewSelection (dmrecon.cc:179-241, global_view_selection.cc) ---- */
/* Returns the number of selected views (ids ascending in ids_out) or a negative error. */
int b200mvs_global_view_selection(b200mvs_ctx* ctx, const b200mvs_settings* s, int ref_view,
                                  int32_t* ids_out, int cap);

/* Engine knob (no reference counterpart): which of the two device implementations of PatchOptimization runs.
 * mode (b200mvs_optimize_patches): 0 = by batch size, 1 = one warp per patch (low latency), 2 = one thread per patch
 * (throughput).  thread_min (b200mvs_reconstruct): in a frontier round, a VIEW with at least this many patches runs them one
 * thread per patch, a view with fewer one warp per patch (the rule looks at the view alone, so a view's maps do not depend on
 * which other views share the batch); -1 = built-in default, 0 = always, a huge value = never.  Both implementations are
 * checked against the oracle. */
int b200mvs_set_patch_mode(b200mvs_ctx* ctx, int mode, int64_t thread_min);

/* Prepares, on host threads, what DMRecon::start computes before its queue runs - analyzeFeatures, globalViewSelection and
 * the seed list of processFeatures (dmrecon.cc:179-292) - for the given reference views, so that a LATER
 * b200mvs_reconstruct of these views (same settings) starts its kernel at once.  May be called from another thread WHILE a
 * b200mvs_reconstruct of a previous batch is running (the reference overlaps them the same way: its OpenMP threads are in
 * different stages of different views, apps/dmrecon/dmrecon.cc:285); cameras and features must not change meanwhile
 * (re-uploading the image of a view with an UNCHANGED camera is allowed).  b200mvs_global_view_selection of a planned view
 * returns the plan's selection; b200mvs_reconstruct uses a plan once and drops it; changing a camera or the features drops
 * all plans.  Host threads: up to hardware_concurrency(), or the value of the environment variable B200MVS_HOST_THREADS
 * (several processes sharing one box, one per GPU). */
int b200mvs_plan_views(b200mvs_ctx* ctx, const b200mvs_settings* s, int n_refs, const int32_t* ref_views);

/* ---- batch of independent PatchOptimization runs: ctor + doAutoOptimization + computeConfidence
 *      (patch_optimization.cc:21-242); the patch-level parity entry ---- */
int b200mvs_optimize_patches(b200mvs_ctx* ctx, const b200mvs_settings* s, int ref_view,
                             const int32_t* global_ids, int n_global,
                             const b200mvs_patch_in* in, int n, b200mvs_patch_out* out,
                             b200mvs_stats* stats_or_null);

/* ---- DMRecon::start (dmrecon.cc:90-172) for a batch of reference views ----
 * Runs analyzeFeatures, globalViewSelection, processFeatures and processQueue for every view in
 * ref_views; all of them advance together, one frontier round per kernel sequence.
 * maps: array of n_refs entries, or NULL to leave the results on the device (HBM-resident timing);
 * progress: array of n_refs entries or NULL; stats: one aggregate or NULL.
 * A view whose global view selection is empty makes the call fail with B200MVS_ERR_GLOBAL_VS
 * (failed_view_or_null receives its id). */
int b200mvs_reconstruct(b200mvs_ctx* ctx, const b200mvs_settings* s, int n_refs, const int32_t* ref_views,
                        b200mvs_maps* maps, b200mvs_progress* progress, b200mvs_stats* stats,
                        int32_t* failed_view_or_null);

/* ---- consumers of the depth maps, on the device.  Stateless: host buffers in, host buffers
 *      out, `device` = CUDA device ordinal.  Errors: negative code, message from b200mvs_depthmap_last_error(). ---- */
const char* b200mvs_depthmap_last_error(void);
/* mve::image::depthmap_confidence_clean (libs/mve/depthmap.cc:118-131): depth = 0 where conf <= 0, in place. */
int b200mvs_depthmap_confidence_clean(int device, float* depth, const float* conf, int w, int h);
/* mve::image::depthmap_cleanup (depthmap.cc:25-113): 4-connected islands of depth != 0 smaller than thres pixels are erased. */
int b200mvs_depthmap_cleanup(int device, const float* depth, int w, int h, int64_t thres, float* out);
/* mve::geom::depthmap_triangulate (depthmap.cc:196-375, the per-view work of apps/scene2pset/scene2pset.cc:264-328):
 * vertex ids per pixel (0xFFFFFFFF = none), vertices (pixel_3dpos; transformed by the 4x4 row-major cam_to_world when given,
 * like mesh_transform), vertex colours (r, g, b, 1 as floats; NULL colour image = none) and faces, all in the reference's
 * order.  Outputs may be NULL except the counts; capacities in vertices / faces (w*h and 2*(w-1)*(h-1) always suffice). */
int b200mvs_depthmap_triangulate(int device, const float* depth, int w, int h, const float invproj[9], float dd_factor,
                                 const float* cam_to_world_or_null, const uint8_t* color_or_null, int color_channels,
                                 uint32_t* vertex_ids, float* vertices, float* colors, uint32_t* faces,
                                 uint64_t cap_vertices, uint64_t cap_faces, uint64_t* n_vertices, uint64_t* n_faces,
                                 double* device_ms_or_null);

/* The per-view work of apps/scene2pset (scene2pset.cc:264-358) in one call: depthmap_triangulate as above plus, per vertex,
 * the angle-weighted normals of TriangleMesh::recalc_normals (mesh.cc:45-151; normals NULL = skip), the boundary confidences
 * of depthmap_mesh_confidences(mesh, conf_iterations) (depthmap.cc:497-548; the app uses 4; confidences NULL or 0 = skip) and
 * the scale values (mean distance to the adjacent vertices of MeshInfo times scale_factor, scene2pset.cc:347-357; NULL = skip). */
int b200mvs_depthmap_pointset(int device, const float* depth, int w, int h, const float invproj[9], float dd_factor,
                              const float* cam_to_world_or_null, const uint8_t* color_or_null, int color_channels,
                              uint32_t* vertex_ids, float* vertices, float* colors, uint32_t* faces,
                              float* normals, float* confidences, int conf_iterations, float* scales, float scale_factor,
                              uint64_t cap_vertices, uint64_t cap_faces, uint64_t* n_vertices, uint64_t* n_faces,
                              double* device_ms_or_null);

#ifdef __cplusplus
}
#endif
#endif /* B200MVS_H */
