/*
 * sppm_b200.h — C ABI of the B200-native SPPM render hot path (libsppm_b200.so).
 *
 * This is the drop-in boundary for wondergo2017/raytrace3's render path.  The
 * reference has no FFI: its boundary is the header-only C++ class API
 * (raytracer/Raytracer.h:366 RayTracer::render, :423 SPPMRayTracer::render).  The
 * replacement keeps those classes (raytrace3_b200/include/) and implements their
 * bodies by flattening Scene/Camera state into the PODs below and calling these
 * entry points.  Each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - plain C, plain pointers and sizes; no C++/torch types cross the boundary;
 *   - every function returns 0 on success, a negative SppmError otherwise;
 *     sppm_last_error(ctx) returns a human-readable message;
 *   - a context owns all device memory; host arrays are borrowed for the duration
 *     of the call; outputs are copied into caller-owned host buffers;
 *   - a context is bound to ONE CUDA device and is not thread-safe;
 *   - all arithmetic is IEEE double, as in the reference (it uses no float anywhere);
 *   - there is no CPU fallback: without a CUDA device sppm_ctx_create fails.
 */
#ifndef SPPM_B200_H
#define SPPM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPPM_ABI_VERSION 3

typedef enum SppmError {
  SPPM_OK = 0,
  SPPM_ERR_INVALID = -1,   /* bad argument / call order */
  SPPM_ERR_CUDA = -2,      /* CUDA runtime failure (message has cudaGetErrorString) */
  SPPM_ERR_NO_DEVICE = -3, /* no usable sm_100 device */
  SPPM_ERR_CAPACITY = -4,  /* a device buffer overflowed (message says which) */
  SPPM_ERR_NCCL = -5       /* NCCL unavailable or failed */
} SppmError;

/* Object kinds: Obj.h:55 PlaneObj, Obj.h:102 SphereObj, Bezier.h:188 Bezier3Obj. */
enum { SPPM_OBJ_PLANE = 0, SPPM_OBJ_SPHERE = 1, SPPM_OBJ_BEZIER = 2 };
/* HitPoint::update() behaviour.  REFERENCE = Raytracer.h:69-79 verbatim (a no-op as
 * shipped because N starts at 0); PROGRESSIVE drops the N<=0 guard. */
enum { SPPM_UPDATE_REFERENCE = 0, SPPM_UPDATE_PROGRESSIVE = 1 };
/* Newton arithmetic.  EXACT = operation order of Bezier.h:69-134 without FMA
 * contraction (bit-identical to the oracle); FAST = the patch in the monomial basis,
 * Horner + FMA, 84 FMA per evaluation (within 1e-6 relative of EXACT, the north-star
 * tolerance); FAST_BERNSTEIN = Bernstein tensor form + FMA (~165 per evaluation). */
enum { SPPM_NEWTON_EXACT = 0, SPPM_NEWTON_FAST = 1, SPPM_NEWTON_FAST_BERNSTEIN = 2 };
/* Photon tracer organisation.  WAVEFRONT = one kernel step per bounce over a compacted
 * queue of live photons + a separate Newton kernel over the queued (ray, patch) items;
 * PERSISTENT = one persistent-warp megakernel.  Same results (the RNG is keyed by photon). */
enum { SPPM_TRACE_WAVEFRONT = 0, SPPM_TRACE_PERSISTENT = 1 };
/* Gather predicate evaluation (Raytracer.h:152-154).  EXACT = three FP64 comparisons on every
 * (hit point, deposit) pair of the 27-cell neighbourhood; FILTERED = the same decisions taken
 * from a packed-FP32 evaluation with a rigorous error bound, FP64 only for the pairs inside the
 * bound of a threshold: newN is bit-identical to EXACT, tao is summed in FP64 in both. */
enum { SPPM_GATHER_EXACT = 0, SPPM_GATHER_FILTERED = 1 };
/* Direction sampling arithmetic (Vec3.h:90-98 GetUnitRandRefl).  EXACT = theta = acos(sqrt(U1))
 * followed by sin/cos of theta as written; FAST = sin(theta) = sqrt(1-U1), cos(theta) = sqrt(U1)
 * taken directly (the same distribution, directions equal to a few ulps). */
enum { SPPM_MATH_EXACT = 0, SPPM_MATH_FAST = 1 };
/* Random-number domains of the counter RNG (who asks for Newton seeds). */
enum { SPPM_DOMAIN_PHOTON = 2, SPPM_DOMAIN_EYE = 3, SPPM_DOMAIN_TEST = 5 };

/* Element.h:7-19 Material (refln is never read on the path and is omitted). */
typedef struct SppmMaterial {
  double diff[3];
  double refl[3];
  double refr[3];
  double refrn;
} SppmMaterial;

/* One entry of Scene::objvec (Scene.h:96), flattened. */
typedef struct SppmObject {
  int32_t kind;         /* SPPM_OBJ_* */
  int32_t tex_id;       /* index into SppmSceneDesc.textures, -1 = untextured (Obj.h:9) */
  int32_t first_patch;  /* BEZIER: first patch in SppmSceneDesc.patches */
  int32_t n_patch;      /* BEZIER: number of patches (Bezier.h:193 Bvec.size()) */
  SppmMaterial mat;     /* Obj.h:17 */
  double color[3];      /* Obj.h:46 */
  double p0[3];         /* PLANE: P0 (Obj.h:58); SPHERE: org (Obj.h:105) */
  double norm[3];       /* PLANE: norm, already normalised as the ctor does (Obj.h:61) */
  double radius;        /* SPHERE (Obj.h:106) */
  double texU[3];       /* Obj.h:63 / Obj.h:107 */
  double texV[3];       /* Obj.h:64 / Obj.h:107 */
} SppmObject;

/* Light.h:5-17 (isotropic point emitter). */
typedef struct SppmLight {
  double pos[3];
  double color[3];
} SppmLight;

/* Element.h:43-75 Texture, pre-decoded: rows*cols*3 bytes, BGR order, row-major. */
typedef struct SppmTexture {
  int32_t rows;
  int32_t cols;
  const uint8_t* bgr8;
} SppmTexture;

/* Scene.h:93-97 flattened.  patches = world-space control points after the
 * Bezier3Obj ctor transform (Bezier.h:199-239), double[n_patches][16][3] with point
 * (j,k) of a patch at index 4*j+k.  Per-patch and per-object AABBs are derived by the
 * library exactly as AABBbox::addpoint does (Bezier.h:12-19). */
typedef struct SppmSceneDesc {
  int32_t n_objects;
  int32_t n_patches;
  int32_t n_lights;
  int32_t n_textures;
  const SppmObject* objects;
  const double* patches;
  const SppmLight* lights;
  const SppmTexture* textures;
} SppmSceneDesc;

/* Camera.h:4-22 after lookAt (Camera.h:32-54): dir carries the focal-length scale. */
typedef struct SppmCameraDesc {
  double pos[3];
  double dir[3];
  double du[3];
  double dv[3];
  int32_t w;
  int32_t h;
} SppmCameraDesc;

/* The reference's compile-time literals that a caller may change.
 * M_EPS=1e-4, PI=3.1415926, PI_INV=0.3183098862, MAX_NUM=1e20 and the 1e-3 normal
 * agreement threshold are compiled in (Vec3.h:6-11, Raytracer.h:154). */
typedef struct SppmParams {
  double init_r2;          /* Raytracer.h:13  INIT_R2 2.0 (also searchR2, :85) */
  double alpha;            /* Raytracer.h:45  0.7 */
  int32_t max_dep;         /* Raytracer.h:12  MAX_DEP 13 (<= 13) */
  int32_t newton_iters;    /* Bezier.h:6      MAX_ITER 10 */
  int32_t newton_restarts; /* Bezier.h:115    maxIter*5 = 50 (<= 4095) */
  int32_t update_mode;     /* SPPM_UPDATE_* */
  int32_t newton_mode;     /* SPPM_NEWTON_* */
  int32_t trace_mode;      /* SPPM_TRACE_* */
  uint64_t seed;           /* key of the Philox4x32-10 counter RNG */
  int32_t gather_mode;     /* SPPM_GATHER_* (ABI version 2) */
  int32_t count_pairs;     /* 1: the gather counts cand_pairs (diagnostic; a few % of its time), 0: cand_pairs stays 0 */
  int32_t math_mode;       /* SPPM_MATH_* */
  int32_t tex_lod;         /* mip level the tracers sample textures at: 0 = the uploaded image (Element.h:61-72, exact),
                            * l > 0 = the l-th 2x2 box-filtered level (an approximation; ABI version 3, was `reserved`) */
} SppmParams;

/* Result of one Scene::NearCollide (Scene.h:165-182 / Element.h:20-39 Collision). */
typedef struct SppmHit {
  double t;        /* Collision::dist; MAX_NUM (1e20) on a miss */
  double pos[3];
  double n[3];
  double color[3];
  double u, v;     /* Bezier parameters of the winning iterate (0 otherwise) */
  int32_t obj;     /* index into objects, -1 = miss (Collision::obj == nullptr) */
  int32_t patch;   /* global patch index for Bezier hits, -1 otherwise */
  int32_t inside;  /* Collision::inside */
  int32_t pad;
} SppmHit;

/* Work counters of the last eye pass / photon passes since sppm_reset_counters. */
typedef struct SppmCounters {
  uint64_t photons;          /* photons emitted */
  uint64_t isect_queries;    /* NearCollide calls */
  uint64_t patch_solves;     /* Bezier3::newton calls (ray,patch pairs past both AABBs) */
  uint64_t deposits;         /* diffuse photon hits (Raytracer.h:136) */
  uint64_t deposits_kept;    /* ... whose grid neighbourhood holds a hit point */
  uint64_t cand_pairs;       /* (deposit, hit point) pairs distance-tested in the gather */
  uint64_t acc_pairs;        /* pairs accepted (Raytracer.h:154) */
  uint64_t hit_points;       /* size of the current hit-point map */
  uint64_t grid_cells;       /* cells of the current hash grid */
  uint64_t launches;         /* kernels launched by this library */
  double ms_eye, ms_build, ms_trace, ms_sort, ms_gather, ms_update, ms_resolve; /* last call of each */
  uint64_t patch_culled;     /* ... of patch_solves, proven empty by the 13-DOP line test and not run */
  /* ABI version 3 */
  uint64_t spilled_photons;  /* photons the wavefront tracer handed to the persistent tracer (item list full) */
  /* per-kernel times of the photon passes run under sppm_set_profiling(ctx, 1) (serialised batches) */
  double ms_step, ms_newton, ms_spill, ms_bin_scan, ms_bin_scatter, ms_gather_prep, ms_gather_kernel;
  double ms_allreduce;       /* sppm_allreduce calls (event span on the communication stream) */
} SppmCounters;

typedef struct sppm_ctx sppm_ctx;

/* ---- lifetime / configuration ------------------------------------------------ */
int sppm_abi_version(void);
int sppm_ctx_create(int device, sppm_ctx** out);
void sppm_ctx_destroy(sppm_ctx* ctx);
const char* sppm_last_error(const sppm_ctx* ctx); /* ctx may be NULL: creation error */
int sppm_default_params(SppmParams* p);
int sppm_set_params(sppm_ctx* ctx, const SppmParams* p);
/* Replaces Scene::init's object graph (Scene.h:98-164) with a device copy. */
int sppm_scene_upload(sppm_ctx* ctx, const SppmSceneDesc* scene);

/* ---- the render path ---------------------------------------------------------- */
/* RayTracer::GetHitPoint + ViewTrace (Raytracer.h:102-116, 296-357). */
int sppm_eye_pass(sppm_ctx* ctx, const SppmCameraDesc* cam, uint64_t pass_id, int64_t* n_hitpoints);
/* Eye-pass sharding: the same pass restricted to film rows [row_begin, row_end);
 * append != 0 adds to the hit points already traced for this film.  Draws are addressed by the
 * pixel index, so any split of the rows yields the hit points of the whole pass. */
int sppm_eye_pass_rows(sppm_ctx* ctx, const SppmCameraDesc* cam, uint64_t pass_id, int32_t row_begin, int32_t row_end,
                       int32_t append, int64_t* n_hitpoints);
/* Exchange the ranks' shares over NCCL so that every rank holds the whole map (no-op without a
 * communicator); call before sppm_build_index. */
int sppm_allgather_hitpoints(sppm_ctx* ctx, int64_t* n_hitpoints);
/* Replaces the flann kd-tree build (Raytracer.h:370-380) by a cell-sorted hash grid. */
int sppm_build_index(sppm_ctx* ctx);
/* One round of RayTracer::PhotonMap (Raytracer.h:226-233) restricted to photons
 * [photon_begin, photon_begin+photon_count) of every light: Light::emit, PhotonTrace
 * and the radiusSearch gather (Raytracer.h:117-209).  Adds into the per-round
 * accumulators (dN, dtao); call sppm_update to fold them into the hit points.
 * Draws are addressed by (pass_id, photon id), never streamed: any split of an id range over
 * calls, batches or GPUs traces the same photons.  While HitPoint::update is the reference's
 * no-op (N = 0, Raytracer.h:70) several rounds may be one call: round i = ids
 * [i*onetime, (i+1)*onetime) -- what the host mirrors' PhotonMap does. */
int sppm_photon_pass(sppm_ctx* ctx, uint64_t photon_begin, uint64_t photon_count, uint64_t pass_id);
/* Sum the per-round accumulators over the ranks of the communicator (no-op without one).
 * Issued on the context's main stream, behind the round's gathers; the photon tracing of the next
 * rounds (lane streams) runs beside it.  SPPM_ASYNC_COMM=1 in the environment (REFERENCE update only)
 * moves this call and the sppm_update that follows it to a communication stream over a second set
 * of accumulators, so that the next round's GATHER overlaps them too; every call that reads
 * hit-point state is then ordered behind them. */
int sppm_allreduce(sppm_ctx* ctx);
/* newN += dN, tao += dtao, then HitPoint::update (Raytracer.h:69-79, 238-245). */
int sppm_update(sppm_ctx* ctx);
/* Film resolve (Raytracer.h:281-293): film[h][w][3] = sum tao/(PI*R2*cnt*onetime).
 * A pixel with several hit points (branching eye paths, Raytracer.h:312-335) sums them in canonical
 * order (eye-path node order, the order GetHitPoint appends them to Hvec) by one thread: the film is
 * the same bit for bit from run to run.  tao itself is not: a round's deposits are summed in the order
 * the binning placed them, which depends on the arrival order of the trace (<= 1e-15 relative; newN is
 * an integer and exact). */
int sppm_resolve(sppm_ctx* ctx, int32_t cnt, int32_t onetime, double* film_hw3);
/* film_acc += film (Raytracer.h:449-458) and tone map (Raytracer.h:24-26, 460-470):
 * bgr[h-1-i][j][k] = toInt(film[i][j][k] / div).  All host buffers. */
int sppm_tonemap(sppm_ctx* ctx, const double* film_hw3, int32_t w, int32_t h, double div, uint8_t* bgr_hw3);

/* ---- shared statistics across eye passes ------------------------------------
 * The reference averages independent PPM renders: every RayTracer::render restarts from
 * R2 = 2 and its HitPoint::update is a no-op (Raytracer.h:69-79, 429-458; README.md:9-17 names
 * SPPM).  These entry points keep ONE set of per-pixel statistics (R2, N, tau) while the camera
 * jitters between eye passes (Hachisuka & Jensen 2009):
 *   N' = N + alpha M,  R2' = R2 (N + alpha M)/(N + M),  tau' = (tau + Phi) R2'/R2,
 *   film = tau / (PI R2 photons_emitted).
 * Per pass: sppm_eye_pass, sppm_build_index, sppm_shared_load, photon rounds with
 * update_mode = REFERENCE (sppm_photon_pass / sppm_allreduce / sppm_update), sppm_shared_fold. */
int sppm_shared_begin(sppm_ctx* ctx, int32_t w, int32_t h);
/* Every hit point of the current map takes the radius of its pixel; its own counters restart. */
int sppm_shared_load(sppm_ctx* ctx);
/* Fold the pass into the pixels: M = sum newN, Phi = sum tao over the pixel's hit points. */
int sppm_shared_fold(sppm_ctx* ctx, uint64_t photons_emitted);
int sppm_shared_resolve(sppm_ctx* ctx, double* film_hw3);
/* Test hook: r2, n = double[h][w]; tau3 = double[3][h][w]; any may be NULL. */
int sppm_shared_export(sppm_ctx* ctx, double* r2_hw, double* n_hw, double* tau3, uint64_t* photons_emitted);
/* Restore statistics saved by sppm_shared_export (film checkpoint / resume); after sppm_shared_begin. */
int sppm_shared_import(sppm_ctx* ctx, const double* r2_hw, const double* n_hw, const double* tau3, uint64_t photons_emitted);

/* ---- host image I/O on the GPU ------------------------------------------------
 * Replaces cv::imread in the Texture constructor (Element.h:47-59) and cv::imwrite of the
 * tone-mapped film (Raytracer.h:474) by nvJPEG (dlopen).  Pixels are interleaved B,G,R bytes,
 * row-major, as cv::Mat / SppmTexture hold them.  All buffers are host memory. */
int sppm_jpeg_info(sppm_ctx* ctx, const uint8_t* jpeg, int64_t n_bytes, int32_t* rows, int32_t* cols);
int sppm_jpeg_decode(sppm_ctx* ctx, const uint8_t* jpeg, int64_t n_bytes, int32_t rows, int32_t cols, uint8_t* bgr8);
/* out may be NULL to query the size; n_bytes returns the length of the JPEG stream. */
int sppm_jpeg_encode(sppm_ctx* ctx, const uint8_t* bgr8, int32_t rows, int32_t cols, int32_t quality,
                     uint8_t* out, int64_t cap, int64_t* n_bytes);

/* ---- multi-GPU: one context per rank, NCCL found with dlopen ------------------- */
int sppm_comm_unique_id(void* id128);                         /* ncclGetUniqueId */
int sppm_comm_init_rank(sppm_ctx* ctx, int rank, int nranks, const void* id128);
/* Device pointers of the per-round accumulators, for callers that run their own
 * collective (torch.distributed): dN int32[n], dtao double[3][n] (SoA). */
int sppm_accum_device_ptrs(sppm_ctx* ctx, void** dN, void** dtao, int64_t* n);
int sppm_device_synchronize(sppm_ctx* ctx);
/* The context's cudaStream_t (all work is ordered on it), for callers that time or
 * chain work with their own CUDA events. */
void* sppm_stream_handle(sppm_ctx* ctx);
/* CUDA-event stopwatch on that stream: record slot 0 (start) / 1 (stop), then read. */
int sppm_timer_record(sppm_ctx* ctx, int slot);
int sppm_timer_elapsed_ms(sppm_ctx* ctx, double* ms);

/* ---- test / measurement hooks --------------------------------------------------- */
/* Scene::NearCollide (Scene.h:165) for n host rays (org,dir); Newton seeds come from
 * the counter RNG addressed by (pass_id, ids[i] or i, node, domain TEST). */
int sppm_intersect_batch(sppm_ctx* ctx, const double* rays6, int64_t n, const uint64_t* ids,
                         uint64_t pass_id, int32_t node, SppmHit* out);
/* Bezier3::newton (Bezier.h:112-159) for n (ray, patch) pairs, no AABB culling.
 * seeds_uv = double[n][restarts][2] injected (u0,v0), or NULL for the counter RNG.
 * tuv_out = double[n][3]; (1e20,1e20,1e20) when nothing was recorded. */
int sppm_bezier_patch_batch(sppm_ctx* ctx, const double* rays6, const int32_t* patch_idx, int64_t n,
                            const double* seeds_uv, double* tuv_out);
/* Same solve on n synthetic device-resident pairs, repeated reps times; returns the
 * average milliseconds of one repetition (CUDA events). */
int sppm_bezier_patch_bench(sppm_ctx* ctx, int64_t n, int32_t reps, uint64_t seed, double* ms_per_rep);
/* Measured FP64 FMA throughput of the device (TFLOP/s, FMA = 2), the roofline denominator
 * of the Newton and gather arithmetic. */
int sppm_fp64_peak(sppm_ctx* ctx, double* tflops);
/* Hit-point map import/export (AoS of the live HitPoint fields, Raytracer.h:47-58):
 * pos3,n3,wgt3 = double[n][3], r2 = double[n], pix = int32[n][2] (row, col). */
int sppm_import_hitpoints(sppm_ctx* ctx, int64_t n, const double* pos3, const double* n3,
                          const double* wgt3, const double* r2, const int32_t* pix, int32_t w, int32_t h);
/* Export in canonical order (sorted by pixel, then eye-path node). key may be NULL. */
int sppm_export_hitpoints(sppm_ctx* ctx, int64_t cap, double* pos3, double* n3, double* wgt3,
                          double* r2, int32_t* pix, uint32_t* node, int64_t* n);
/* The gather of Raytracer.h:139-159 for m host deposits (x, N, flux = double[m][3]). */
int sppm_gather_deposits(sppm_ctx* ctx, int64_t m, const double* x3, const double* n3, const double* flux3);
/* Per hit point, canonical order: newN, N (int32), R2, tao[3] (after sppm_update). */
int sppm_export_stats(sppm_ctx* ctx, int64_t cap, int32_t* newN, int32_t* N, double* r2, double* tao3);
/* Keep / export the deposit list of the next photon passes (test only). */
int sppm_keep_deposits(sppm_ctx* ctx, int32_t keep_all);
int sppm_export_deposits(sppm_ctx* ctx, int64_t cap, double* x3, double* n3, double* flux3,
                         uint64_t* photon_id, int32_t* bounce, int64_t* m);
int sppm_get_counters(sppm_ctx* ctx, SppmCounters* out);
/* Measurement aid: on = 1 makes every batch of sppm_photon_pass run alone (device sync after it) with a
 * CUDA event between its launches, so that SppmCounters.ms_step / ms_newton / ... hold the time of each
 * kernel class as an ncu launch list sees it; on = 0 (default) is the pipelined production path. */
int sppm_set_profiling(sppm_ctx* ctx, int32_t on);

/* Page-lock a host buffer the caller owns (and keeps alive until sppm_host_unpin), so that the film read-back of
 * sppm_resolve and the texture / film uploads copy by DMA at the full PCIe rate instead of through the driver's
 * staging buffer.  Optional: every entry point takes pageable memory as well.  What the reference's
 * camera->pic (Camera.h:13, filled at Raytracer.h:281-293) becomes in the drop-in Camera.  Needs no context. */
int sppm_host_pin(void* ptr, uint64_t bytes);
int sppm_host_unpin(void* ptr);
int sppm_reset_counters(sppm_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SPPM_B200_H */
