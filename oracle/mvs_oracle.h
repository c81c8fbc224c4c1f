/* TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's dense-MVS depth-map hot path
 * (libs/dmrecon of simonfuhrmann/mve) used only as the parity checker by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * Nothing under mve_b200/ may include, link or call this.
 *
 * PINNING: the reference holds no golden vectors or tests for this path
 * (SURVEY.md §4, §8c), so this restatement is pinned against OUTPUTS OF THE
 * REFERENCE ITSELF, compiled unmodified into oracle/_ref by oracle/Makefile:
 * whole depth/conf/dz maps and the printed global view selection of
 * oracle/_ref/dmrecon, and per-patch results of mvs::PatchOptimization through
 * oracle/_ref/ref_harness (tests/test_oracle_vs_reference.py, fixtures under
 * tests/golden/ minted by tests/golden/make_golden.py).
 *
 * Plain C ABI so tests can drive it through ctypes.
 */
#ifndef MVS_ORACLE_H
#define MVS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mvs_oracle_scene mvs_oracle_scene;

/* POD mirror of mvs::Settings (libs/dmrecon/settings.h:22-52). */
typedef struct {
    uint32_t filter_width;       /* must be 5 (patch_sampler.cc:96 hard-codes index 12) */
    float    min_ncc;
    float    min_parallax;
    float    accept_ncc;
    float    min_refine_diff;
    uint32_t max_iterations;
    uint32_t nr_recon_neighbors;
    uint32_t global_vs_max;
    int32_t  scale;
    int32_t  use_color_scale;
    float    aabb_min[3];
    float    aabb_max[3];
} mvs_oracle_settings;

/* One PatchOptimization: inputs (patch_optimization.cc:21-30) and outputs. */
typedef struct {
    int32_t x, y;
    float   depth, dz_i, dz_j;
    int32_t n_local;             /* number of propagated local view ids (0..4) */
    int32_t local_ids[4];
} mvs_oracle_patch_in;

typedef struct {
    float   conf;                /* computeConfidence(), 0 when not converged */
    float   depth, dz_i, dz_j;
    float   normal[3];
    int32_t n_local;
    int32_t local_ids[4];        /* ascending (std::set order) */
    int32_t iterations;          /* status.iterationCount */
    int32_t converged;
    int32_t opti_success;
} mvs_oracle_patch_out;

/* Work counters: scene constants for the roofline figure. */
typedef struct {
    uint64_t n_opt;              /* PatchOptimization objects constructed */
    uint64_t n_pse_deriv;        /* fastColAndDeriv executions      (patch_sampler.cc:65)  */
    uint64_t n_pse_color;        /* computeNeighColorSamples execs  (patch_sampler.cc:348) */
    uint64_t n_ncc;              /* getFastNCC calls                (patch_sampler.cc:136) */
    uint64_t n_update;           /* PatchSampler::update calls      (patch_sampler.cc:259) */
    uint64_t n_pops;             /* queue pops                      (dmrecon.cc:365)       */
    uint64_t n_stale;            /* pops dropped by the stale test  (dmrecon.cc:371)       */
    uint64_t n_filled;           /* progress.filled                                           */
    uint64_t n_seeds_processed;  /* "Processed N features"          (dmrecon.cc:286)       */
    uint64_t n_seeds_success;    /* "from which N succeeded"        (dmrecon.cc:301)       */
    uint64_t n_spec_rounds;      /* simulated speculative-batch rounds of the strict-order GPU schedule */
    uint64_t n_spec_wasted;      /* batch-computed entries later dropped as stale */
} mvs_oracle_stats;

mvs_oracle_scene* mvs_oracle_create(int n_views);
void mvs_oracle_destroy(mvs_oracle_scene*);

/* mve::View + SingleView::create + ImagePyramidCache (image_pyramid.cc:19-95). rgb is HxWx3 uint8. */
int mvs_oracle_set_view(mvs_oracle_scene*, int view_id, const uint8_t* rgb, int w, int h,
                        float flen, float paspect, const float ppoint[2],
                        const float rot[9], const float trans[3]);

/* mve::Bundle features: positions + CSR list of referencing view ids. */
int mvs_oracle_set_features(mvs_oracle_scene*, int n_feat, const float* pos,
                            const int32_t* ref_offsets, const int32_t* ref_view_ids);

int mvs_oracle_num_levels(mvs_oracle_scene*, int view_id);
/* Copies pyramid level `level` (rgb, w*h*3) and its dimensions. */
int mvs_oracle_get_level(mvs_oracle_scene*, int view_id, int level, int* w, int* h, uint8_t* rgb_or_null);
/* proj / invproj of a level (image_pyramid.h:51-59). */
int mvs_oracle_get_level_calib(mvs_oracle_scene*, int view_id, int level, float proj[9], float invproj[9]);

/* analyzeFeatures + GlobalViewSelection (dmrecon.cc:179-241). Returns count, ids ascending. */
int mvs_oracle_global_view_selection(mvs_oracle_scene*, const mvs_oracle_settings*, int ref_view,
                                     int32_t* ids_out, int cap);

/* Batch of independent PatchOptimization runs (ctor + doAutoOptimization + computeConfidence). */
int mvs_oracle_optimize_patches(mvs_oracle_scene*, const mvs_oracle_settings*, int ref_view,
                                const int32_t* global_ids, int n_global,
                                const mvs_oracle_patch_in* in, int n, mvs_oracle_patch_out* out,
                                mvs_oracle_stats* stats_or_null);

/* DMRecon::start (dmrecon.cc:90-172): analyzeFeatures, globalViewSelection, processFeatures, processQueue.
 * Maps are Ws*Hs (level `scale` of the ref view), row-major: depth[1], conf[1], dz[2], normal[3],
 * view_ids[4] (int32, -1 padded; the local view ids of the optimisation that wrote the pixel).
 * trace_in/out (optional, capacity trace_cap) receive the first PatchOptimizations in execution order;
 * *trace_n receives the total number executed (may exceed trace_cap).
 * order_mode 0 = reference strict priority order. */
int mvs_oracle_reconstruct(mvs_oracle_scene*, const mvs_oracle_settings*, int ref_view,
                           float* depth, float* conf, float* dz, float* normal, int32_t* view_ids,
                           mvs_oracle_stats* stats,
                           mvs_oracle_patch_in* trace_in, mvs_oracle_patch_out* trace_out,
                           int64_t trace_cap, int64_t* trace_n, double max_seconds);

/* Same reconstruction under the deterministic frontier schedule the GPU uses (see the comment at the
 * definition).  band <= 0 and topk <= 0: every queued entry runs each round; otherwise only the entries at or above the round's
 * confidence threshold (frontier_band / frontier_topk of include/b200mvs.h). stats.n_spec_rounds = number of rounds. */
int mvs_oracle_reconstruct_wavefront(mvs_oracle_scene*, const mvs_oracle_settings*, int ref_view, float band, int topk,
                                     float* depth, float* conf, float* dz, float* normal, int32_t* view_ids,
                                     mvs_oracle_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
