/*
 * sppm_oracle.h — C interface of the CPU oracle (TEST INFRASTRUCTURE, not product).
 *
 * The oracle is a dependency-free restatement of wondergo2017/raytrace3's render path
 * (raytracer/{Vec3,Element,Obj,Bezier,Light,Scene,Camera,Raytracer}.h).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load
 * it.  Parity status: the reference ships no tests or golden vectors, so the
 * oracle is pinned against THE REFERENCE'S OWN CODE run here: oracle/_ref/libref.so = the
 * reference's headers compiled from /root/reference/raytracer (three MSVC-only tokens patched
 * in a scratch copy) against a thin OpenCV/FLANN shim (oracle/ref_shim), with the random
 * numbers injected in the reference's consumption order; tests/test_ref_pin_cpu.py shows the
 * oracle bit-identical to it for patch evaluation, AABB, Newton, sampling, optics, camera,
 * textures, NearCollide, PhotonTrace + gather, ViewTrace, update, resolve and tone map.
 * The third-party arithmetic the shim restates is pinned separately: cv2.invert for
 * cv::Matx33d::inv, cv2.flann for the strict '<' radius search; Philox against Random123
 * known answers; plus committed golden vectors (tests/golden).
 *
 * The PODs are the ones of include/sppm_b200.h so the same flattened scene feeds both.
 */
#ifndef SPPM_ORACLE_H
#define SPPM_ORACLE_H
#include "../include/sppm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Philox4x32-10 (Salmon et al., SC'11), the shared counter RNG. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* Two uniforms in [0,1) (53-bit) for the draw addressed by the arguments. */
void orc_draw2(uint64_t seed, uint64_t pass, uint64_t id, uint32_t node, uint32_t purpose,
               uint32_t sub, uint32_t patch, double out[2]);

/* cv::Matx33d::inv (closed form, zeros when det == 0) — Bezier.h:126-131. */
void orc_inv3(const double m[9], double out[9]);
/* AABBbox::intersect — Bezier.h:20-51. box = {min[3], max[3]}. */
int orc_aabb_intersect(const double box[6], const double org[3], const double dir[3]);
/* evalPatchPoint / evalPatchDerivU / evalPatchDerivV — Bezier.h:85-111. */
void orc_bezier_eval(const double P[48], double u, double v, double S[3], double Su[3], double Sv[3]);
/* Bezier3::newton with injected (u0,v0) seeds — Bezier.h:112-159. */
void orc_newton(const double P[48], const double org[3], const double dir[3], const double* seeds_uv,
                int restarts, int iters, double tuv[3]);
/* Texture::colorUV — Element.h:61-72. */
void orc_color_uv(const SppmTexture* tex, double u, double v, double out[3]);

/* Vec3 helpers exposed for unit tests (Vec3.h:57-65, 90-115, 124-134). */
void orc_unit_rand_refl(const double norm[3], double u1, double u2, double out[3]);
void orc_refracted(const double I[3], const double N[3], double n, double out[3]);

/* Camera::lookAt (Camera.h:32-54) -> SppmCameraDesc. */
void orc_camera_lookat(const double pos[3], const double look[3], int w, int h, SppmCameraDesc* cam);
/* Camera::emit (Camera.h:18-22). */
void orc_camera_emit(const SppmCameraDesc* cam, int y, int x, double ray6[6]);

/* Scene::NearCollide on a batch (Scene.h:165-182); Newton seeds from the counter RNG
 * at (pass_id, ids[i] or i, node, domain). */
void orc_intersect_batch(const SppmSceneDesc* scene, const SppmParams* prm, const double* rays6, int64_t n,
                         const uint64_t* ids, uint64_t pass_id, int32_t node, int32_t domain, SppmHit* out);
void orc_bezier_patch_batch(const SppmSceneDesc* scene, const SppmParams* prm, const double* rays6,
                            const int32_t* patch_idx, int64_t n, const double* seeds_uv, double* tuv_out);

/* GetHitPoint + ViewTrace (Raytracer.h:102-116, 296-357).  Returns the hit-point count
 * (may exceed cap; only cap are written).  Order = the reference's Hvec order. */
int64_t orc_eye_pass(const SppmSceneDesc* scene, const SppmParams* prm, const SppmCameraDesc* cam,
                     uint64_t pass_id, int64_t cap, double* pos3, double* n3, double* wgt3, double* r2,
                     int32_t* pix, uint32_t* node);

int64_t orc_view_trace_pixel(const SppmSceneDesc* scene, const SppmParams* prm, const SppmCameraDesc* cam, uint64_t pass_id,
                             int i, int j, int64_t cap, double* pos3, double* n3, double* wgt3, double* r2, int32_t* pix);

/* The gather of Raytracer.h:139-159 by exhaustive search (the semantic definition). */
void orc_gather_brute(int64_t n_hp, const double* pos3, const double* n3, const double* wgt3,
                      const double* r2, int64_t m, const double* x3, const double* dn3, const double* flux3,
                      double search_r2, int32_t* newN, double* tao3, uint64_t* cand_pairs, uint64_t* acc_pairs);
/* Same through an exact kd-tree (leaf size 10), standing in for flann::KDTreeSingleIndex. */
void* orc_kdtree_build(int64_t n_hp, const double* pos3);
void orc_kdtree_free(void* tree);
void orc_gather_kdtree(void* tree, int64_t n_hp, const double* pos3, const double* n3, const double* wgt3,
                       const double* r2, int64_t m, const double* x3, const double* dn3, const double* flux3,
                       double search_r2, int32_t* newN, double* tao3);

/* One round of PhotonMap for photons [begin, begin+count) of every light
 * (Raytracer.h:226-233 + PhotonTrace :117-209).  tree = orc_kdtree_build handle or NULL
 * for brute force.  Optionally records every diffuse deposit (cap_dep entries). */
int64_t orc_photon_pass(const SppmSceneDesc* scene, const SppmParams* prm, uint64_t begin, uint64_t count,
                        uint64_t pass_id, void* tree, int64_t n_hp, const double* pos3, const double* n3,
                        const double* wgt3, const double* r2, int32_t* newN, double* tao3, int64_t cap_dep,
                        double* dep_x3, double* dep_n3, double* dep_flux3, uint64_t* dep_photon,
                        int32_t* dep_bounce, uint64_t* counters /* [4]: queries, solves, deposits, photons */);

/* Timing legs (bench.py cpu_baseline / --impl reference): see sppm_oracle.cpp. */
double orc_photon_bench(const SppmSceneDesc* scene, const SppmParams* prm, uint64_t count, uint64_t pass_id,
                        int32_t threads, void* tree, int64_t n_hp, const double* pos3, const double* n3,
                        const double* wgt3, const double* r2, uint64_t* counters);
double orc_patch_bench(const SppmSceneDesc* scene, const SppmParams* prm, const double* rays6,
                       const int32_t* patch_idx, int64_t n, int32_t threads, double* tuv_out);

/* HitPoint::update (Raytracer.h:69-79) over all hit points. */
void orc_update(int64_t n_hp, int32_t mode, double alpha, int32_t* newN, int32_t* N, double* r2, double* tao3);
/* Film resolve (Raytracer.h:281-293). film = double[h][w][3], zeroed first. */
void orc_resolve(int64_t n_hp, const int32_t* pix, const double* r2, const double* tao3, int32_t cnt,
                 int32_t onetime, int32_t w, int32_t h, double* film);
/* toInt + row flip (Raytracer.h:24-26, 460-470). */
void orc_tonemap(const double* film, int32_t w, int32_t h, double div, uint8_t* bgr);

/* The reference's whole CPU structure for timing (Raytracer.h:423-447): `threads`
 * independent jittered RayTracer replicas under OpenMP, each = eye pass + kd-tree
 * build + cnt rounds x onetime photons + resolve.  film (may be NULL) receives the sum.
 * times[4] = {eye_s, build_s, photon_s, total_s} (max over replicas).
 * counters[4] summed over replicas. */
void orc_render_replicas(const SppmSceneDesc* scene, const SppmParams* prm, const double cam_pos[3],
                         const double cam_look[3], int32_t w, int32_t h, int32_t cnt, int32_t onetime,
                         int32_t threads, double jitter, double* film, double* times, uint64_t* counters);
int orc_max_threads(void);
/* Draw log for feeding oracle/_ref (single-ray calls only; see sppm_oracle.cpp). */
void orc_log_begin(void);
int64_t orc_log_end(double* vals, signed char* kinds, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
