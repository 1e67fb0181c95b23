// trace.cu — ray/photon tracing kernels for sm_100a:
//   eye_pass_kernel      RayTracer::GetHitPoint + ViewTrace   (Raytracer.h:102-116, 296-357)
//   photon_trace_kernel  Light::emit + RayTracer::PhotonTrace (Light.h:8-13, Raytracer.h:117-209)
//   intersect_kernel     Scene::NearCollide on a ray batch     (Scene.h:165-182)
//   patch_batch_kernel   Bezier3::newton on (ray, patch) pairs (Bezier.h:112-159)
// All of them go through the warp-cooperative near_collide of scene.cuh.
#include "ctx.h"
#include "trace_kernels.cuh"

namespace sppm {

namespace {

TraceCommon make_common(sppm_ctx* c, uint64_t pass) {
  TraceCommon t;
  t.scene = c->scene;
  t.ncfg.restarts = c->prm.newton_restarts;
  t.ncfg.iters = c->prm.newton_iters;
  t.ncfg.mode = c->prm.newton_mode;
  t.ncfg.seed = c->prm.seed;
  t.ncfg.cull = c->newton_cull ? 1 : 0;
  t.patches_in_smem = c->patches_in_smem ? 1 : 0;
  t.max_dep = c->prm.max_dep;
  t.fast_math = c->prm.math_mode == SPPM_MATH_FAST ? 1 : 0;
  t.pass = pass;
  t.counters = (DevCounters*)c->counters.p;
  return t;
}

}  // namespace

cudaError_t trace_configure(sppm_ctx* c) {
  // objects always in shared memory; patches too when they fit beside the warp scratch
  const size_t scratch = TRACE_WARPS * sizeof(WarpScratch);
  size_t with = scene_smem_bytes(c->scene.n_obj, c->scene.n_patch, true) + scratch;
  c->patches_in_smem = with <= 100 * 1024;
  c->trace_smem = scene_smem_bytes(c->scene.n_obj, c->scene.n_patch, c->patches_in_smem) + scratch;
  c->wf_cfg.valid = false;
  c->trace_per_sm = 0;
  cudaError_t e;
  const int bytes = (int)c->trace_smem;
  if ((e = cudaFuncSetAttribute(eye_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
  if ((e = cudaFuncSetAttribute(photon_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
  if ((e = cudaFuncSetAttribute(intersect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
  if ((e = cudaFuncSetAttribute(patch_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
  return cudaSuccess;
}

cudaError_t launch_eye_pass(sppm_ctx* c, const SppmCameraDesc& cam, uint64_t pass_id, int row_begin, int row_end) {
  EyeArgs a;
  a.c = make_common(c, pass_id);
  a.cam = cam;
  a.out = (HpRaw*)c->hp_raw.p;
  a.cap = (unsigned long long)c->hp_cap;
  a.hp_count = (unsigned long long*)c->dev_scalars.p + 0;
  a.init_r2 = c->prm.init_r2;
  a.pix_begin = (long long)row_begin * cam.w;
  a.pix_end = (long long)row_end * cam.w;
  const long long npix = a.pix_end - a.pix_begin;
  if (npix <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((npix + TRACE_THREADS - 1) / TRACE_THREADS);
  eye_pass_kernel<<<blocks, TRACE_THREADS, c->trace_smem, c->stream>>>(a);
  c->host_counters.launches++;
  return cudaGetLastError();
}

cudaError_t launch_photon_trace(sppm_ctx* c, Lane& L, uint64_t begin, uint64_t count, uint64_t pass_id,
                                const unsigned* r_count, double* const* r_st, const unsigned long long* r_idd) {
  PhotonArgs a;
  a.c = make_common(c, pass_id);
  a.r_count = r_count;  // resume mode: finish the photons the wavefront tracer spilled (count stays on the device)
  for (int k = 0; k < 9; k++) a.r_st[k] = r_count ? r_st[k] : nullptr;
  a.r_idd = r_idd;
  a.grid = c->grid;
  a.cell_slot = c->index_ok ? (const int32_t*)c->cell_slot.p : nullptr;
  a.n_slots = c->n_slots;
  a.keep_all = c->keep_all_deposits ? 1 : 0;
  a.dep = (Deposit*)L.dep_raw.p;
  a.dep_key = (uint32_t*)L.dep_key.p;
  a.dep_rank = (uint32_t*)L.dep_rank.p;
  a.slot_count = (unsigned int*)L.slot_count.p;
  a.dep_cap = (unsigned long long)L.dep_cap;
  a.dep_count = (unsigned long long*)L.scalars.p + 1;
  a.cursor = (unsigned long long*)L.scalars.p + 2;
  a.begin = begin;
  a.count = count;
  if (c->trace_per_sm <= 0) {
    int q = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, photon_trace_kernel, TRACE_THREADS, c->trace_smem);
    if (e) return e;
    c->trace_per_sm = q < 1 ? 1 : q;
  }
  const int per_sm = c->trace_per_sm;
  const unsigned long long total = (unsigned long long)c->scene.n_light * count;
  unsigned long long want = (total + TRACE_THREADS - 1) / TRACE_THREADS;
  unsigned long long blocks = (unsigned long long)c->sm_count * per_sm;  // persistent: a whole number of waves
  if (want < blocks) blocks = want ? want : 1;
  photon_trace_kernel<<<(unsigned)blocks, TRACE_THREADS, c->trace_smem, L.stream>>>(a);
  c->host_counters.launches++;
  return cudaGetLastError();
}

cudaError_t launch_intersect(sppm_ctx* c, const double* d_rays6, int64_t n, const uint64_t* d_ids, uint64_t pass_id,
                             int node, SppmHit* d_out) {
  IsectArgs a;
  a.c = make_common(c, pass_id);
  a.rays6 = d_rays6;
  a.ids = d_ids;
  a.n = n;
  a.node = node;
  a.out = d_out;
  const unsigned blocks = (unsigned)((n + TRACE_THREADS - 1) / TRACE_THREADS);
  if (!blocks) return cudaSuccess;
  intersect_kernel<<<blocks, TRACE_THREADS, c->trace_smem, c->stream>>>(a);
  c->host_counters.launches++;
  return cudaGetLastError();
}

cudaError_t launch_patch_batch(sppm_ctx* c, const double* d_rays6, const int32_t* d_patch, int64_t n,
                               const double* d_seeds, double* d_tuv) {
  PatchArgs a;
  a.c = make_common(c, 0);
  a.rays6 = d_rays6;
  a.patch = d_patch;
  a.n = n;
  a.seeds = d_seeds;
  a.tuv = d_tuv;
  if (n <= 0) return cudaSuccess;
  int per_sm = 1;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, patch_batch_kernel, TRACE_THREADS, c->trace_smem);
  if (e) return e;
  if (per_sm < 1) per_sm = 1;
  long long want = (n + TRACE_THREADS - 1) / TRACE_THREADS;
  long long blocks = (long long)c->sm_count * per_sm;
  if (want < blocks) blocks = want;
  patch_batch_kernel<<<(unsigned)blocks, TRACE_THREADS, c->trace_smem, c->stream>>>(a);
  c->host_counters.launches++;
  return cudaGetLastError();
}

}  // namespace sppm
