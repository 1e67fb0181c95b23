// wavefront.cu — the photon pass as a bounce-synchronous wavefront (sm_100a).
//
// Light::emit + RayTracer::PhotonTrace (Light.h:8-13, Raytracer.h:117-209) are split into two
// kernels per bounce over a COMPACTED queue of live photons (SoA in HBM, ping-pong):
//
//   photon_step_kernel(b)   per live photon: [b = 0: emit | b > 0: finish the previous query —
//                           combine the analytic best with the Bezier solves, depth check,
//                           shade, deposit, Roulette, next direction]; then the next query's
//                           analytic tests (planes / spheres) and the cooperative Bezier AABB
//                           culling, which appends (photon slot, patch) items to a global list;
//                           survivors are re-compacted into four regions of the queue, by what
//                           their pending hit makes the next step do.
//   newton_items_kernel(b)  Bezier3::newton (Bezier.h:112-159) for all queued items, the
//                           item x restart tasks packed over full warps.
//
// Compared with the persistent megakernel (trace.cu) every kernel is coherent: no lane idles
// on a finished photon, the register-heavy Newton solver does not limit the occupancy of the
// light per-photon code, and the Newton tasks of the whole GPU are packed densely.
// Results are identical: the RNG is addressed by (pass, photon id, bounce, ...).
#include "ctx.h"
#include "wavefront_kernels.cuh"

namespace sppm {

cudaError_t wavefront_configure(sppm_ctx* c, int* step_blocks, int* newton_blocks, size_t* step_smem, size_t* newton_smem) {
  if (c->wf_cfg.valid) {  // once per scene: four driver queries less per batch
    *step_blocks = c->wf_cfg.step_blocks; *newton_blocks = c->wf_cfg.newton_blocks;
    *step_smem = c->wf_cfg.step_smem; *newton_smem = c->wf_cfg.newton_smem;
    return cudaSuccess;
  }
  const size_t scene = scene_smem_bytes(c->scene.n_obj, c->scene.n_patch, c->patches_in_smem, true);
  *step_smem = scene_smem_bytes(c->scene.n_obj, c->scene.n_patch, c->patches_in_smem, false);
  *newton_smem = scene + TRACE_WARPS * sizeof(WarpScratch);
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(photon_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*step_smem))) return e;
  if ((e = cudaFuncSetAttribute(newton_items_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*newton_smem))) return e;
  int a = 1, b = 1;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, photon_step_kernel, STEP_THREADS, *step_smem))) return e;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, newton_items_kernel, TRACE_THREADS, *newton_smem))) return e;
  if (const char* e = getenv("SPPM_STEP_CTAS")) a = std::max(1, std::min(a, atoi(e)));  // tuning knobs
  if (const char* e = getenv("SPPM_NEWTON_CTAS")) b = std::max(1, std::min(b, atoi(e)));
  *step_blocks = c->sm_count * (a < 1 ? 1 : a);
  *newton_blocks = c->sm_count * (b < 1 ? 1 : b);
  c->wf_cfg.step_blocks = *step_blocks; c->wf_cfg.newton_blocks = *newton_blocks;
  c->wf_cfg.step_smem = *step_smem; c->wf_cfg.newton_smem = *newton_smem;
  c->wf_cfg.valid = true;
  return cudaSuccess;
}

// One batch of photons [begin, begin+count) per light through all bounces.
cudaError_t launch_photon_trace_wavefront(sppm_ctx* c, Lane& Ln, uint64_t begin, uint64_t count, uint64_t pass_id) {
  const int64_t total = (int64_t)c->scene.n_light * (int64_t)count;
  if (total <= 0) return cudaSuccess;
  cudaError_t e;
  if (total > Ln.wf_cap) {
    if ((e = dev_reserve(Ln.wf_store, wf_bytes(total, wf_item_cap(total))))) return e;
    Ln.wf_cap = total;
  }
  WfLayout L = wf_layout(Ln.wf_store.p, Ln.wf_cap, wf_item_cap(Ln.wf_cap));
  // test knob: a short item list, so that the spill path (photons finished by the persistent tracer) runs
  if (const char* e = getenv("SPPM_WF_ITEM_CAP")) L.item_cap = (unsigned)std::max<long long>(1, std::min<long long>(L.item_cap, atoll(e)));
  int step_blocks, newton_blocks;
  size_t step_smem, newton_smem;
  if ((e = wavefront_configure(c, &step_blocks, &newton_blocks, &step_smem, &newton_smem))) return e;
  cudaStream_t st = Ln.stream;
  if ((e = cudaMemsetAsync(L.counters, 0, 128 * 4, st))) return e;
  WfBatch w;
  w.scene = c->scene;
  w.patches_in_smem = c->patches_in_smem ? 1 : 0;
  w.newton_cull = c->newton_cull ? 1 : 0;
  w.max_dep = c->prm.max_dep;
  w.newton_restarts = c->prm.newton_restarts;
  w.newton_iters = c->prm.newton_iters;
  w.newton_mode = c->prm.newton_mode;
  w.fast_math = c->prm.math_mode == SPPM_MATH_FAST ? 1 : 0;
  w.seed = c->prm.seed;
  w.pass = pass_id;
  w.counters = (DevCounters*)c->counters.p;
  w.begin = begin;
  w.count = count;
  // Measured alternatives (profiles/r2_D_ab.txt, r2_F_ab_configs.txt).  SPPM_WF_TAIL = b: from bounce b on the
  // Bezier-pending photons go to the persistent tracer in one launch instead of one latency-bound Newton launch
  // per bounce — the Newton launches shrink by 0.2 ms per round but the hand-over kernel costs 0.44 ms (a warp of
  // 32 such photons runs ~100 serial Newton rounds): off.  SPPM_NEWTON_DYNAMIC: chunks from a global cursor, which
  // newton_items_kernel uses only for bounces whose static chunk would hit its cap (four teapots): on.
  w.tail_bounce = -1;
  if (const char* e = getenv("SPPM_WF_TAIL")) w.tail_bounce = atoi(e);
  w.newton_dynamic = 1;
  if (const char* e = getenv("SPPM_NEWTON_DYNAMIC")) w.newton_dynamic = atoi(e);
  w.q_cap = (unsigned)Ln.wf_cap;
  w.grid = c->grid;
  w.cell_slot = c->index_ok ? (const int32_t*)c->cell_slot.p : nullptr;
  w.n_slots = c->n_slots;
  w.keep_all = c->keep_all_deposits ? 1 : 0;
  w.dep = (Deposit*)Ln.dep_raw.p;
  w.dep_key = (uint32_t*)Ln.dep_key.p;
  w.dep_rank = (uint32_t*)Ln.dep_rank.p;
  w.slot_count = (unsigned int*)Ln.slot_count.p;
  w.dep_cap = (unsigned long long)Ln.dep_cap;
  w.dep_count = (unsigned long long*)Ln.scalars.p + 1;
  const int n_steps = wf_n_steps(c->prm.max_dep);
  for (int b = 0; b < n_steps; b++) {
    const StepArgs a = wf_step_args(w, L, b);
    int blocks = step_blocks;
    if (b == 0) {
      const int64_t want = (total + STEP_THREADS - 1) / STEP_THREADS;
      if (want < blocks) blocks = (int)want;
    }
    if ((e = Ln.prof.mark(PK_STEP, st))) return e;
    photon_step_kernel<<<blocks, STEP_THREADS, step_smem, st>>>(a);
    c->host_counters.launches++;
    if (a.do_intersect && c->scene.n_patch > 0 && !a.spill_all) {
      const NewtonArgs n = wf_newton_args(w, L, a);
      if ((e = Ln.prof.mark(PK_NEWTON, st))) return e;
      newton_items_kernel<<<newton_blocks, TRACE_THREADS, newton_smem, st>>>(n);
      c->host_counters.launches++;
    }
  }
  if ((e = Ln.prof.mark(PK_SPILL, st))) return e;
  // Photons whose bounce found the item list full were taken out of the wavefront (spill list): the
  // persistent tracer finishes them (it returns at once when the list is empty, the usual case).
  if (c->scene.n_patch > 0) {
    if ((e = launch_photon_trace(c, Ln, begin, count, pass_id, L.spill.count, L.spill.st, L.spill.idd))) return e;
  }
  return cudaGetLastError();
}

}  // namespace sppm
