// device_host.cpp — TEST INFRASTRUCTURE, not product: the device functions of
// raytrace3_b200/csrc/{device_math.cuh, scene.cuh}, the host table builder scene_prep.h and the
// KERNELS of trace_kernels.cuh (eye pass, persistent photon tracer, NearCollide batch, Newton
// batch), compiled for the CPU by g++ through cuda_host_shim.h and exported as batch functions for
// tests/test_device_host_cpu.py.  The point is that the `-m "not gpu"` suite checks the SAME source
// lines the sm_100a kernels are built from (plane / sphere / AABB tests with their division-free
// rejects, patch evaluation in the three arithmetic modes, the Newton step and restart, the Newton
// culls, texture look-up, sampling and optics, the whole warp-cooperative near_collide + shade_hit)
// against the oracle, at fuzz sizes that would cost GPU minutes.  Built by tests/device_host/build.py:
//   g++ -O2 -ffp-contract=off (no FMA contraction, like the library's -fmad=false) -pthread
#include "cuda_host_shim.h"

#include <atomic>
#include <functional>
#include <numeric>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

// dynamic shared memory of the block a host thread is emulating (see __shared__ in the shim)
namespace sppm {
namespace {
alignas(16) thread_local unsigned char smem[160 << 10];   // trace / wavefront kernels
alignas(128) thread_local unsigned char gsm[64 << 10];    // gather_kernel
alignas(128) thread_local unsigned char wsm[64 << 10];    // gather_filtered_kernel
}
}  // namespace sppm
#include "../../raytrace3_b200/csrc/scene_prep.h"
#include "../../raytrace3_b200/csrc/trace_kernels.cuh"
#include "../../raytrace3_b200/csrc/wavefront_kernels.cuh"
#include "../../raytrace3_b200/csrc/gather_kernels.cuh"
#include "../../raytrace3_b200/csrc/image_kernels.cuh"
namespace sppm {
namespace {
unsigned long long dh_work[WK_N] = {};  // SPPM_WORK counters of gather_filtered_kernel
}
}  // namespace sppm

thread_local DhWarp* dh_warp = nullptr;

using namespace sppm;

namespace {

struct DhScene {
  PreppedScene ps;
  std::vector<std::vector<uint32_t>> bgrx;
  std::vector<DevTexture> tex;
  double lut255[256];
  unsigned long long culled = 0;
  SceneView sv;
  std::vector<SppmLight> lights;
  DevScene dev;  // what sppm_scene_upload hands the kernels (pointers into this object)
};

NewtonCfg make_cfg(int mode, int restarts, int iters, int cull, uint64_t seed) {
  NewtonCfg c;
  c.restarts = restarts;
  c.iters = iters;
  c.mode = mode;
  c.cull = cull;
  c.seed = seed;
  return c;
}

// run f(lane) for the 32 lanes of one emulated warp (fibers on the calling thread, see the shim)
thread_local std::function<void(int)>* dh_fn = nullptr;
void fiber_entry() {
  DhWarp* w = dh_warp;
  const int me = w->cur;
  (*dh_fn)(me);
  w->done[me] = true;
  // the lanes of a warp pass the same barriers, so lane 0 ends first and the others wait behind the
  // warp's last barrier: finish the warp, then go on with the next warp that has not finished (parked
  // at a block barrier, or not started)
  for (int k = 1; k < 32; k++) {
    const int next = (me & ~31) | ((me + k) & 31);
    if (!w->done[next]) {
      w->cur = next;
      setcontext(&w->lane_ctx[next]);
    }
  }
  for (int k = 1; k < w->n_threads; k++) {
    const int next = (me + k) % w->n_threads;
    if (!w->done[next]) {
      w->cur = next;
      setcontext(&w->lane_ctx[next]);
    }
  }
  setcontext(&w->main_ctx);
}
unsigned long long run_block(std::function<void(int)> f, int n_threads, unsigned block = 0, unsigned grid = 1) {
  constexpr size_t STACK = 160 << 10;
  struct Stacks {  // one set of thread stacks per host thread, reused by every block it emulates
    char* p = (char*)malloc((size_t)DH_MAX_THREADS * STACK);
    ~Stacks() { free(p); }
  };
  static thread_local Stacks lane_stacks;
  static thread_local DhWarp warp;  // ~300 KB of contexts: not on the stack
  char* stacks = lane_stacks.p;
  if (n_threads % 32 || n_threads > DH_MAX_THREADS) abort();
  warp.n_threads = n_threads;
  warp.barriers = 0;
  for (int l = 0; l < n_threads; l++) {
    warp.done[l] = false;
    getcontext(&warp.lane_ctx[l]);
    warp.lane_ctx[l].uc_stack.ss_sp = stacks + (size_t)l * STACK;
    warp.lane_ctx[l].uc_stack.ss_size = STACK;
    warp.lane_ctx[l].uc_link = nullptr;
    makecontext(&warp.lane_ctx[l], fiber_entry, 0);
  }
  dh_warp = &warp;
  dh_fn = &f;
  warp.cur = 0;
  warp.block = block;
  warp.grid = grid;
  swapcontext(&warp.main_ctx, &warp.lane_ctx[0]);
  dh_warp = nullptr;
  dh_fn = nullptr;
  for (int l = 0; l < n_threads; l++)
    if (!warp.done[l]) abort();  // a divergent collective
  return warp.barriers;
}

// kernel<<<blocks, threads, smem>>>(args): blocks handed out to host threads in order
template <class K, class A>
void call_kernel(K kernel, const A& args) { kernel(args); }
template <class K, class... As>
void call_kernel(K kernel, const std::tuple<As...>& args) { std::apply(kernel, args); }
template <class K, class A>
void launch(K kernel, const A& args, long long blocks, int threads = 32) {
  if (blocks <= 0) return;
  const int T = (int)std::max<long long>(1, std::min<long long>(blocks, std::min(16u, std::max(1u, std::thread::hardware_concurrency()))));
  std::atomic<long long> next{0};
  std::vector<std::thread> th;
  for (int w = 0; w < T; w++)
    th.emplace_back([&] {
      for (;;) {
        const long long b = next.fetch_add(1);
        if (b >= blocks) break;
        run_block([&](int) { call_kernel(kernel, args); }, threads, (unsigned)b, (unsigned)blocks);
      }
    });
  for (auto& t : th) t.join();
}

// Block sizes: the product's own (TRACE_THREADS, STEP_THREADS, W_THREADS, TPB, 320) by default, or
// one-warp blocks (dh_set_one_warp_blocks(1)), a second schedule every kernel but dep_scatter allows.
int g_one_warp = 0;
inline int bt(int product_threads) { return g_one_warp ? 32 : product_threads; }

// make_common (trace.cu) for an emulated launch; patches staged in shared memory like the library
// does whenever they fit (trace_configure)
TraceCommon common_of(DhScene* h, const SppmParams* prm, int cull, uint64_t pass, DevCounters* counters) {
  TraceCommon t;
  t.scene = h->dev;
  t.ncfg = make_cfg(prm->newton_mode, prm->newton_restarts, prm->newton_iters, cull, prm->seed);
  const size_t with = scene_smem_bytes(t.scene.n_obj, t.scene.n_patch, true) + TRACE_WARPS * sizeof(WarpScratch);
  t.patches_in_smem = with <= 100 * 1024 ? 1 : 0;
  t.max_dep = prm->max_dep;
  t.fast_math = prm->math_mode == SPPM_MATH_FAST ? 1 : 0;
  t.pass = pass;
  t.counters = counters;
  return t;
}

}  // namespace

extern "C" {

void dh_draw2(uint64_t seed, uint64_t pass, uint64_t id, uint32_t node, uint32_t purpose, uint32_t sub, uint32_t patch,
              double out[2]) {
  draw2(seed, pass, id, node, purpose, sub, patch, out[0], out[1]);
}
void dh_unit_rand_vec(uint64_t seed, uint64_t pass, int64_t n, const uint64_t* ids, uint32_t purpose, double* out3) {
  for (int64_t i = 0; i < n; i++) {
    const d3 v = unit_rand_vec(seed, pass, ids[i], purpose);
    out3[3 * i] = v.x; out3[3 * i + 1] = v.y; out3[3 * i + 2] = v.z;
  }
}
void dh_aabb_intersect(const double box[6], int64_t n, const double* rays6, int32_t* out) {
  for (int64_t i = 0; i < n; i++) out[i] = aabb_intersect(box, ld3(rays6 + 6 * i), ld3(rays6 + 6 * i + 3)) ? 1 : 0;
}
// mode 0: eval_exact(P), 1: eval_mono(A), 2: eval_fast(P); out[n][9] = S, Su, Sv
void dh_bezier_eval(int mode, const double* P48, int64_t n, const double* uv, double* out9) {
  for (int64_t i = 0; i < n; i++) {
    const double u = uv[2 * i], v = uv[2 * i + 1];
    const PatchEval e = mode == 0 ? eval_exact(P48, u, v) : (mode == 1 ? eval_mono(P48, u, v) : eval_fast(P48, u, v));
    double* o = out9 + 9 * i;
    o[0] = e.S.x; o[1] = e.S.y; o[2] = e.S.z;
    o[3] = e.Su.x; o[4] = e.Su.y; o[5] = e.Su.z;
    o[6] = e.Sv.x; o[7] = e.Sv.y; o[8] = e.Sv.z;
  }
}
// Bezier3::newton as the kernels run it: `restarts` x newton_restart<MODE>, the min over restarts
// kept with (t, restart index) order -- flush_items' segmented arg-min.  seeds[n][restarts][2].
void dh_newton(int mode, const double* P48, int64_t n, const double* rays6, const double* seeds, int restarts, int iters,
               double* tuv) {
  for (int64_t i = 0; i < n; i++) {
    const d3 org = ld3(rays6 + 6 * i), dir = ld3(rays6 + 6 * i + 3);
    double bt = SPPM_MAX_NUM, bu = SPPM_MAX_NUM, bv = SPPM_MAX_NUM;
    for (int r = 0; r < restarts; r++) {
      const double u0 = seeds[((size_t)i * restarts + r) * 2], v0 = seeds[((size_t)i * restarts + r) * 2 + 1];
      double t, u, v;
      const bool found = mode == 0   ? newton_restart<0>(P48, org, dir, u0, v0, iters, t, u, v)
                         : mode == 1 ? newton_restart<1>(P48, org, dir, u0, v0, iters, t, u, v)
                                     : newton_restart<2>(P48, org, dir, u0, v0, iters, t, u, v);
      if (found && t < bt) {
        bt = t; bu = u; bv = v;
      }
    }
    tuv[3 * i] = bt; tuv[3 * i + 1] = bu; tuv[3 * i + 2] = bv;
  }
}
// the tiled expansion + mip chain sppm_scene_upload builds on the device (abi.cu), by the product's kernels
static std::vector<uint32_t> expand_texels(const uint8_t* bgr, int rows, int cols, DevTexture& t) {
  std::vector<uint32_t> words(tex_layout_levels(rows, cols, nullptr, t));
  tex_layout_levels(rows, cols, words.data(), t);
  t.bgr = bgr;
  const long long n0 = (long long)rows * cols;
  launch(tex_expand_kernel, std::make_tuple(bgr, rows, cols, t.level[0].tiles_x, words.data()), (n0 + 255) / 256, 256);
  for (int l = 1; l < t.n_levels; l++) {
    const long long nl = (long long)t.level[l].rows * t.level[l].cols;
    launch(tex_mip_kernel, std::make_tuple(t.level[l - 1], t.level[l]), (nl + 255) / 256, 256);
  }
  return words;
}
void dh_color_uv(int rows, int cols, const uint8_t* bgr8, int64_t n, const double* uv, double* out3, int lod) {
  DevTexture t;
  const std::vector<uint32_t> words = expand_texels(bgr8, rows, cols, t);
  double lut[256];
  for (int i = 0; i < 256; i++) lut[i] = i / 255.0;
  for (int64_t i = 0; i < n; i++) {
    const d3 c = color_uv(t, lut, uv[2 * i], uv[2 * i + 1], lod);
    out3[3 * i] = c.x; out3[3 * i + 1] = c.y; out3[3 * i + 2] = c.z;
  }
}
void dh_unit_rand_refl(int64_t n, const double* norm3, const double* u12, int fast, double* out3) {
  for (int64_t i = 0; i < n; i++) {
    const d3 v = unit_rand_refl(ld3(norm3 + 3 * i), u12[2 * i], u12[2 * i + 1], fast != 0);
    out3[3 * i] = v.x; out3[3 * i + 1] = v.y; out3[3 * i + 2] = v.z;
  }
}
// normalize3(refracted(...)) as every call site has it (trace.cu, wavefront.cu), like orc_refracted
void dh_refracted(int64_t n, const double* I3, const double* N3, const double* nn, double* out3) {
  for (int64_t i = 0; i < n; i++) {
    const d3 v = normalize3(refracted(ld3(I3 + 3 * i), ld3(N3 + 3 * i), nn[i]));
    out3[3 * i] = v.x; out3[3 * i + 1] = v.y; out3[3 * i + 2] = v.z;
  }
}

// ---- a scene flattened by the product's own prep_scene ----------------------------------------
void* dh_scene_create(const SppmSceneDesc* s, char* err, int errlen) {
  DhScene* h = new DhScene();
  if (const char* bad = prep_scene(s, h->ps)) {
    if (err && errlen > 0) {
      strncpy(err, bad, (size_t)errlen - 1);
      err[errlen - 1] = 0;
    }
    delete h;
    return nullptr;
  }
  h->tex.resize((size_t)s->n_textures);
  for (int t = 0; t < s->n_textures; t++) {
    const SppmTexture& ht = s->textures[t];
    h->bgrx.push_back(expand_texels(ht.bgr8, ht.rows, ht.cols, h->tex[t]));  // level pointers go into the vector's heap block: stable
  }
  for (int i = 0; i < 256; i++) h->lut255[i] = i / 255.0;  // stage_scene
  SceneView& v = h->sv;
  v.obj = h->ps.objs.data();
  v.patchP = h->ps.patchP.data();
  v.patchA = h->ps.patchA.data();
  v.patchBox = h->ps.patchBox.data();
  v.patchDop = h->ps.patchDop.data();
  v.patchSub = h->ps.patchSub.data();
  v.culled = &h->culled;
  v.tex = h->tex.data();
  v.lut255 = h->lut255;
  v.n_obj = (int)h->ps.objs.size();
  h->lights.assign(s->lights, s->lights + s->n_lights);
  DevScene& d = h->dev;
  d.n_obj = v.n_obj;
  d.n_patch = s->n_patches;
  d.n_light = s->n_lights;
  d.n_tex = s->n_textures;
  d.obj = v.obj;
  d.patchP = v.patchP;
  d.patchA = v.patchA;
  d.patchBox = v.patchBox;
  d.patchDop = v.patchDop;
  d.patchSub = v.patchSub;
  d.culled = &h->culled;
  d.lights = h->lights.data();
  d.tex = v.tex;
  return h;
}
void dh_scene_free(void* p) { delete (DhScene*)p; }
int64_t dh_scene_n_patches(void* p) { return (int64_t)(((DhScene*)p)->ps.patchBox.size() / 6); }
uint64_t dh_scene_culled(void* p) { return ((DhScene*)p)->culled; }
void dh_scene_tables(void* p, double* box6, double* dop26, double* sub96, double* A48, double* objbox6) {
  const DhScene* h = (const DhScene*)p;
  memcpy(box6, h->ps.patchBox.data(), h->ps.patchBox.size() * 8);
  memcpy(dop26, h->ps.patchDop.data(), h->ps.patchDop.size() * 8);
  memcpy(sub96, h->ps.patchSub.data(), h->ps.patchSub.size() * 8);
  memcpy(A48, h->ps.patchA.data(), h->ps.patchA.size() * 8);
  for (size_t i = 0; i < h->ps.objs.size(); i++) memcpy(objbox6 + 6 * i, h->ps.objs[i].box, 48);
}
// the material flags and Roulette weights of object i (Element.h:16-18, Obj.h:12-15)
void dh_scene_material(void* p, int i, int32_t flags[3], double weights[4]) {
  const DevObject& o = ((const DhScene*)p)->ps.objs[(size_t)i];
  flags[0] = o.isDiff; flags[1] = o.isRefl; flags[2] = o.isRefr;
  weights[0] = o.diffr; weights[1] = o.reflr; weights[2] = o.refrr; weights[3] = o.allr;
}
// planes and spheres only, per ray, in object order: the analytic loop of photon_step_kernel
void dh_analytic_nearest(void* p, int64_t n, const double* rays6, int32_t* obj, double* t, int32_t* inside) {
  const SceneView& sv = ((const DhScene*)p)->sv;
  for (int64_t i = 0; i < n; i++) {
    const d3 org = ld3(rays6 + 6 * i), dir = ld3(rays6 + 6 * i + 3);
    Hit best;
    best.dist = SPPM_MAX_NUM;
    best.obj = -1;
    best.inside = false;
    for (int o = 0; o < sv.n_obj; o++) {
      const DevObject& ob = sv.obj[o];
      if (ob.kind == SPPM_OBJ_PLANE) test_plane(ob, o, org, dir, best);
      else if (ob.kind == SPPM_OBJ_SPHERE) test_sphere(ob, o, org, dir, best);
    }
    obj[i] = best.obj;
    t[i] = best.dist;
    inside[i] = best.inside ? 1 : 0;
  }
}
// the two Newton culls of flush_items on (ray, global patch) pairs: 1 = "this solve records nothing"
void dh_newton_cull(void* p, int64_t n, const double* rays6, const int32_t* patch, int32_t* dop_miss, int32_t* sub_miss) {
  const SceneView& sv = ((const DhScene*)p)->sv;
  for (int64_t i = 0; i < n; i++) {
    const d3 org = ld3(rays6 + 6 * i), dir = ld3(rays6 + 6 * i + 3);
    dop_miss[i] = dop_line_miss(sv.patchDop + (size_t)patch[i] * 26, org, dir) ? 1 : 0;
    sub_miss[i] = sub_boxes_miss(sv.patchSub + (size_t)patch[i] * 96, org, dir) ? 1 : 0;
  }
}
// ---- the kernels of trace_kernels.cuh, launched with one-warp blocks -----------------------------
// counters9 = DevCounters {photons, isect_queries, patch_solves, deposits, deposits_kept, cand_pairs,
// acc_pairs, overflow, patch_culled}
int64_t dh_eye_pass_kernel(void* p, const SppmParams* prm, const SppmCameraDesc* cam, uint64_t pass, int row_begin,
                           int row_end, int64_t cap, HpRaw* out, uint64_t* counters9) {
  DhScene* h = (DhScene*)p;
  DevCounters cnt;
  memset(&cnt, 0, sizeof(cnt));
  unsigned long long n_hp = 0;
  EyeArgs a;
  a.c = common_of(h, prm, 1, pass, &cnt);
  a.cam = *cam;
  a.out = out;
  a.cap = (unsigned long long)cap;
  a.hp_count = &n_hp;
  a.init_r2 = prm->init_r2;
  a.pix_begin = (long long)row_begin * cam->w;
  a.pix_end = (long long)row_end * cam->w;
  const long long npix = a.pix_end - a.pix_begin;
  if (npix > 0) launch(eye_pass_kernel, a, (npix + bt(TRACE_THREADS) - 1) / bt(TRACE_THREADS), bt(TRACE_THREADS));
  if (counters9) memcpy(counters9, &cnt, sizeof(cnt));
  return (int64_t)n_hp;
}
// photon_trace_kernel without a hit-point map (keep_all: every diffuse deposit is recorded)
int64_t dh_photon_trace_kernel(void* p, const SppmParams* prm, uint64_t begin, uint64_t count, uint64_t pass, int blocks,
                               int64_t cap, Deposit* out, uint64_t* counters9) {
  DhScene* h = (DhScene*)p;
  DevCounters cnt;
  memset(&cnt, 0, sizeof(cnt));
  std::vector<uint32_t> key((size_t)cap), rank((size_t)cap);
  unsigned int slot_count[2] = {0, 0};
  unsigned long long n_dep = 0, cursor = 0;
  PhotonArgs a;
  a.c = common_of(h, prm, 1, pass, &cnt);
  memset(&a.grid, 0, sizeof(a.grid));
  a.cell_slot = nullptr;
  a.n_slots = 0;
  a.keep_all = 1;
  a.dep = out;
  a.dep_key = key.data();
  a.dep_rank = rank.data();
  a.slot_count = slot_count;
  a.dep_cap = (unsigned long long)cap;
  a.dep_count = &n_dep;
  a.cursor = &cursor;
  a.begin = begin;
  a.count = count;
  a.r_count = nullptr;
  for (int k = 0; k < 9; k++) a.r_st[k] = nullptr;
  a.r_idd = nullptr;
  launch(photon_trace_kernel, a, blocks, bt(TRACE_THREADS));
  if (counters9) memcpy(counters9, &cnt, sizeof(cnt));
  return (int64_t)n_dep;
}
void dh_intersect_kernel(void* p, const SppmParams* prm, int cull, int64_t n, const double* rays6, const uint64_t* ids,
                         uint64_t pass, int node, SppmHit* out, uint64_t* counters9) {
  DhScene* h = (DhScene*)p;
  DevCounters cnt;
  memset(&cnt, 0, sizeof(cnt));
  IsectArgs a;
  a.c = common_of(h, prm, cull, pass, &cnt);
  a.rays6 = rays6;
  a.ids = ids;
  a.n = n;
  a.node = node;
  a.out = out;
  if (n > 0) launch(intersect_kernel, a, (n + bt(TRACE_THREADS) - 1) / bt(TRACE_THREADS), bt(TRACE_THREADS));
  if (counters9) memcpy(counters9, &cnt, sizeof(cnt));
}
void dh_patch_batch_kernel(void* p, const SppmParams* prm, int cull, int64_t n, const double* rays6, const int32_t* patch,
                           const double* seeds, int blocks, double* tuv) {
  DhScene* h = (DhScene*)p;
  DevCounters cnt;
  memset(&cnt, 0, sizeof(cnt));
  PatchArgs a;
  a.c = common_of(h, prm, cull, 0, &cnt);
  a.rays6 = rays6;
  a.patch = patch;
  a.n = n;
  a.seeds = seeds;
  a.tuv = tuv;
  if (n > 0) launch(patch_batch_kernel, a, blocks, bt(TRACE_THREADS));
}
// launch_photon_trace_wavefront (wavefront.cu) with the product's own queue layout and per-bounce
// arguments (wf_layout / wf_step_args / wf_newton_args): photon_step_kernel and newton_items_kernel
// per bounce, every diffuse deposit kept.  Returns the deposit count; -1 if the item list overflowed.
int64_t dh_photon_trace_wavefront(void* p, const SppmParams* prm, uint64_t begin, uint64_t count, uint64_t pass,
                                  int step_blocks, int newton_blocks, int64_t cap, Deposit* out, uint64_t* counters9,
                                  uint32_t* queue_counts /* [4 * (max_dep + 3)] or null */,
                                  int64_t item_cap_override /* > 0: a short item list, so that photons spill */,
                                  int tail_bounce /* WfBatch::tail_bounce, -1 = off */) {
  DhScene* h = (DhScene*)p;
  const int64_t total = (int64_t)h->dev.n_light * (int64_t)count;
  if (total <= 0) return 0;
  DevCounters cnt;
  memset(&cnt, 0, sizeof(cnt));
  const int64_t item_cap = wf_item_cap(total);
  std::vector<unsigned char> store(wf_bytes(total, item_cap) + 256);
  void* base = (void*)(((uintptr_t)store.data() + 255) & ~(uintptr_t)255);
  WfLayout L = wf_layout(base, total, item_cap);
  if (item_cap_override > 0) L.item_cap = (unsigned)std::min<int64_t>(item_cap, item_cap_override);  // SPPM_WF_ITEM_CAP of wavefront.cu
  memset(L.counters, 0, 128 * 4);
  std::vector<uint32_t> key((size_t)cap), rank((size_t)cap);
  unsigned int slot_count[2] = {0, 0};
  unsigned long long n_dep = 0;
  const TraceCommon tc = common_of(h, prm, 1, pass, &cnt);
  WfBatch w;
  w.scene = h->dev;
  w.patches_in_smem = tc.patches_in_smem;
  w.newton_cull = 1;
  w.max_dep = prm->max_dep;
  w.newton_restarts = prm->newton_restarts;
  w.newton_iters = prm->newton_iters;
  w.newton_mode = prm->newton_mode;
  w.fast_math = tc.fast_math;
  w.seed = prm->seed;
  w.pass = pass;
  w.counters = &cnt;
  w.begin = begin;
  w.count = count;
  w.q_cap = (unsigned)total;
  w.tail_bounce = tail_bounce;
  w.newton_dynamic = (tail_bounce & 1) ? 1 : 0;  // both chunking schemes get exercised
  memset(&w.grid, 0, sizeof(w.grid));
  w.cell_slot = nullptr;
  w.n_slots = 0;
  w.keep_all = 1;
  w.dep = out;
  w.dep_key = key.data();
  w.dep_rank = rank.data();
  w.slot_count = slot_count;
  w.dep_cap = (unsigned long long)cap;
  w.dep_count = &n_dep;
  const int n_steps = wf_n_steps(prm->max_dep);
  for (int b = 0; b < n_steps; b++) {
    const StepArgs a = wf_step_args(w, L, b);
    long long blocks = step_blocks;
    if (b == 0) blocks = std::min<long long>(blocks, (total + bt(STEP_THREADS) - 1) / bt(STEP_THREADS));
    launch(photon_step_kernel, a, blocks, bt(STEP_THREADS));
    if (a.do_intersect && h->dev.n_patch > 0 && !a.spill_all) {
      const NewtonArgs n = wf_newton_args(w, L, a);
      launch(newton_items_kernel, n, newton_blocks, bt(TRACE_THREADS));
    }
  }
  if (h->dev.n_patch > 0) {  // the spilled photons, finished by the persistent tracer (launch_photon_trace's resume mode)
    unsigned long long cursor = 0;
    PhotonArgs a;
    a.c = tc;
    a.grid = w.grid;
    a.cell_slot = nullptr;
    a.n_slots = 0;
    a.keep_all = 1;
    a.dep = out;
    a.dep_key = key.data();
    a.dep_rank = rank.data();
    a.slot_count = slot_count;
    a.dep_cap = (unsigned long long)cap;
    a.dep_count = &n_dep;
    a.cursor = &cursor;
    a.begin = begin;
    a.count = count;
    a.r_count = L.spill.count;
    for (int k = 0; k < 9; k++) a.r_st[k] = L.spill.st[k];
    a.r_idd = L.spill.idd;
    launch(photon_trace_kernel, a, newton_blocks, bt(TRACE_THREADS));
  }
  if (counters9) memcpy(counters9, &cnt, sizeof(cnt));
  if (queue_counts) memcpy(queue_counts, L.counters, sizeof(uint32_t) * 4 * (size_t)(n_steps + 1));
  return (int64_t)n_dep;
}
// ---- build_index + bin_deposits + launch_gather (gather.cu) with the product's kernels ----------
// The host orchestration is restated here with std:: algorithms in place of CUB (stable sort by cell
// key, run-length encode, exclusive scans, flagged select); every kernel between them is the
// product's: hp_key / hp_permute / mark_nbr / finalize_slot / run_flag / unit_count / unit_fill /
// occ_run / dep_key / dep_scatter / unit_parts and gather_filtered_kernel<COUNT> or, with `exact`, the
// all-FP64 gather_kernel (CTA per occupied cell).  raw = hit points in canonical order.
// Outputs in the CELL-SORTED order of the hit-point map: order[i] = raw index of sorted hit point i,
// dN[n], dtao[3][n]; info = {n_occ, n_slots, n_units, n_items, kept deposits}; pairs = {cand, acc}.
int dh_gather_filtered(int64_t n, const HpRaw* raw, const Grid* grid, int64_t m, const Deposit* dep_raw, double search_r2,
                       int count_pairs, int exact, int blocks, uint32_t* order, int32_t* dN, double* dtao, int64_t* info,
                       uint64_t* pairs) {
  const Grid g = *grid;
  if (n <= 0) return -1;
  const int T = bt(TPB);
  auto nb = [&](long long k) { return (k + T - 1) / T; };
  // hit points: cell keys, stable sort, SoA permute
  std::vector<uint32_t> iota((size_t)n), idx((size_t)n), skey((size_t)n);
  std::vector<uint64_t> key((size_t)n), skey64((size_t)n);  // (cell, Morton code in the cell)
  // canonical order first (sort by okey = pixel, eye-path node), then a stable sort by cell key
  std::vector<uint64_t> okey0((size_t)n);
  launch(hp_okey_kernel, std::make_tuple(raw, (long long)n, okey0.data(), iota.data()), nb(n), T);
  std::stable_sort(iota.begin(), iota.end(), [&](uint32_t a, uint32_t b) { return okey0[a] < okey0[b]; });
  launch(hp_key_kernel, std::make_tuple(raw, (const uint32_t*)iota.data(), (long long)n, g, key.data(), 1), nb(n), T);
  std::vector<uint32_t> pos((size_t)n);  // key[] is indexed by canonical position
  std::iota(pos.begin(), pos.end(), 0u);
  std::stable_sort(pos.begin(), pos.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
  for (int64_t i = 0; i < n; i++) {
    idx[(size_t)i] = iota[pos[(size_t)i]];
    skey64[(size_t)i] = key[pos[(size_t)i]];
  }
  launch(hp_cell_of_key_kernel, std::make_tuple((const uint64_t*)skey64.data(), (long long)n, skey.data()), nb(n), T);
  const size_t nd = (size_t)n;
  std::vector<double> dstore(nd * 16);
  std::vector<uint64_t> okey(nd);
  std::vector<int32_t> istore(nd * 5);
  HpSoA s;
  double* d = dstore.data();
  s.px = d; s.py = d + nd; s.pz = d + 2 * nd; s.nx = d + 3 * nd; s.ny = d + 4 * nd; s.nz = d + 5 * nd; s.r2 = d + 6 * nd;
  s.wx = d + 7 * nd; s.wy = d + 8 * nd; s.wz = d + 9 * nd; s.tao = d + 10 * nd; s.dtao = d + 13 * nd;
  s.okey = okey.data();
  int32_t* ip = istore.data();
  s.N = ip; s.newN = ip + nd; s.dN = ip + 2 * nd; s.row = ip + 3 * nd; s.col = ip + 4 * nd;
  launch(hp_permute_kernel, std::make_tuple(raw, (const uint32_t*)idx.data(), (long long)n, s), nb(n), T);
  // occupied cells and their hit-point ranges
  std::vector<uint32_t> occ_cell;
  std::vector<int32_t> occ_start;
  for (int64_t i = 0; i < n; i++)
    if (i == 0 || skey[(size_t)i] != skey[(size_t)i - 1]) {
      occ_cell.push_back(skey[(size_t)i]);
      occ_start.push_back((int32_t)i);
    }
  const int n_occ = (int)occ_cell.size();
  occ_start.push_back((int32_t)n);
  // compact slots of the cells whose 27-neighbourhood holds a hit point
  std::vector<int32_t> cell_slot((size_t)g.n_cells, 0), scan((size_t)g.n_cells);
  launch(mark_nbr_kernel, std::make_tuple((const uint32_t*)occ_cell.data(), n_occ, g, cell_slot.data()), nb((long long)n_occ * 27), T);
  std::exclusive_scan(cell_slot.begin(), cell_slot.end(), scan.begin(), 0);
  int32_t n_slots = 0;
  launch(finalize_slot_kernel, std::make_tuple(cell_slot.data(), (const int32_t*)scan.data(), (long long)g.n_cells, &n_slots),
         nb(g.n_cells), T);
  std::vector<int2> occ_run((size_t)n_occ * 9);
  launch(occ_run_kernel, std::make_tuple((const uint32_t*)occ_cell.data(), n_occ, g, (const int32_t*)cell_slot.data(), occ_run.data()),
         nb((long long)n_occ * 9), T);
  // work units
  std::vector<int32_t> flag((size_t)n_occ), run_first, ucnt((size_t)n_occ + 1, 0), uoff((size_t)n_occ + 1);
  launch(run_flag_kernel, std::make_tuple((const uint32_t*)occ_cell.data(), n_occ, g.dim[0], flag.data()), nb(n_occ), T);
  for (int o = 0; o < n_occ; o++)
    if (flag[(size_t)o]) run_first.push_back(o);
  int32_t n_runs = (int32_t)run_first.size();
  run_first.resize((size_t)n_occ + 1, 0);
  launch(unit_count_kernel, std::make_tuple((const int32_t*)run_first.data(), (const int32_t*)&n_runs, n_occ,
                                            (const int32_t*)occ_start.data(), ucnt.data()), nb(n_occ), T);
  std::exclusive_scan(ucnt.begin(), ucnt.end(), uoff.begin(), 0);
  const int n_units = uoff[(size_t)n_occ];
  std::vector<GUnit> units((size_t)std::max(n_units, 1));
  launch(unit_fill_kernel, std::make_tuple((const int32_t*)run_first.data(), (const int32_t*)&n_runs, n_occ,
                                           (const int32_t*)occ_start.data(), (const uint32_t*)occ_cell.data(),
                                           (const int32_t*)uoff.data(), n_units, g, (const int32_t*)cell_slot.data(), units.data()),
         nb(std::max(n_units, 1)), T);
  // deposits: slot key + arrival rank, CSR offsets, scatter (key n_slots = not kept)
  std::vector<uint32_t> dkey((size_t)std::max<int64_t>(m, 1)), drank((size_t)std::max<int64_t>(m, 1));
  std::vector<unsigned int> slot_count((size_t)n_slots + 2, 0);
  std::vector<int32_t> dep_off((size_t)n_slots + 2);
  if (m > 0)
    launch(dep_key_kernel, std::make_tuple(dep_raw, (long long)m, g, (const int32_t*)cell_slot.data(), n_slots, dkey.data(),
                                           drank.data(), slot_count.data()), nb(m), T);
  std::exclusive_scan(slot_count.begin(), slot_count.end(), dep_off.begin(), 0);
  std::vector<Deposit> sorted((size_t)std::max<int64_t>(m, 1));
  {
    // arrival ranks depend on the order the emulated blocks ran in; any order is a valid counting sort
    const unsigned long long m_dev = (unsigned long long)m;
    if (m > 0)
      launch(dep_scatter_kernel, std::make_tuple(dep_raw, (const uint32_t*)dkey.data(), (const uint32_t*)drank.data(), &m_dev,
                                                 (unsigned long long)m, (const int32_t*)dep_off.data(), sorted.data()), 12, 320);
  }
  // work items of the batch and the gather
  std::vector<int32_t> parts((size_t)n_units + 1), item_off((size_t)n_units + 1);
  launch(unit_parts_kernel, std::make_tuple((const GUnit*)units.data(), n_units, (const int32_t*)dep_off.data(), parts.data()),
         nb(n_units + 1), T);
  std::exclusive_scan(parts.begin(), parts.end(), item_off.begin(), 0);
  DevCounters cnt;
  memset(&cnt, 0, sizeof(cnt));
  GatherArgs a;
  memset(&a, 0, sizeof(a));
  a.s = s;
  a.n_hp = n;
  a.occ_start = occ_start.data();
  a.occ_run = occ_run.data();  // exact kernel only
  a.n_occ = n_occ;
  a.dep_off = dep_off.data();
  a.dep = sorted.data();
  a.search_r2 = search_r2;
  a.counters = &cnt;
  a.grid = g;
  a.occ_cell = occ_cell.data();
  a.units = units.data();
  a.n_units = n_units;
  a.full_accept = 1;
  a.item_off = item_off.data();
  unsigned int work_counter = 0;
  if (n_units > 0) {
    if (exact) launch(gather_kernel, std::make_tuple(a, &work_counter), std::min(blocks, n_occ), G_THREADS);
    else if (count_pairs) launch(gather_filtered_kernel<true>, std::make_tuple(a, &work_counter), blocks, bt(W_THREADS));
    else launch(gather_filtered_kernel<false>, std::make_tuple(a, &work_counter), blocks, bt(W_THREADS));
  }
  memcpy(order, idx.data(), nd * 4);
  memcpy(dN, s.dN, nd * 4);
  memcpy(dtao, s.dtao, nd * 24);
  info[0] = n_occ; info[1] = n_slots; info[2] = n_units; info[3] = item_off[(size_t)n_units]; info[4] = dep_off[(size_t)n_slots];
  pairs[0] = cnt.cand_pairs; pairs[1] = cnt.acc_pairs;
  return 0;
}
// update_kernel (HitPoint::update, Raytracer.h:69-79, on the round's accumulators) and resolve_kernel
// (Raytracer.h:281-293) over caller-owned structure-of-arrays; tao / dtao are [3][n], film is [h][w][3]
// (zeroed by the caller).  rounds = how many update launches (each folds dN / dtao, given once).
void dh_update_resolve(int64_t n, int mode, double alpha, int rounds, int32_t* N, int32_t* newN, int32_t* dN, double* r2,
                       double* tao, double* dtao, const int32_t* row, const int32_t* col, int cnt, int onetime, int w,
                       double* film, const uint32_t* raw_of_rank, const uint32_t* raw_of_pos) {
  HpSoA s;
  memset(&s, 0, sizeof(s));
  s.N = N; s.newN = newN; s.dN = dN; s.r2 = r2; s.tao = tao; s.dtao = dtao;
  s.row = const_cast<int32_t*>(row); s.col = const_cast<int32_t*>(col);
  const int T = bt(TPB);
  for (int k = 0; k < rounds; k++) launch(update_kernel, std::make_tuple(s, (long long)n, mode, alpha), (n + T - 1) / T, T);
  if (!film) return;
  // canonical rank -> position as build_index derives it from its two sorts
  std::vector<uint32_t> rank_of_raw((size_t)n), canon((size_t)n);
  launch(hp_rank_kernel, std::make_tuple(raw_of_rank, (long long)n, rank_of_raw.data()), (n + T - 1) / T, T);
  launch(hp_canon_kernel, std::make_tuple(raw_of_pos, (const uint32_t*)rank_of_raw.data(), (long long)n, canon.data()), (n + T - 1) / T, T);
  launch(resolve_kernel, std::make_tuple(s, (const uint32_t*)canon.data(), (long long)n, cnt, onetime, w, film), (n + T - 1) / T, T);
}
// Shared per-pixel statistics: shared_init (if init_r2 > 0) -> shared_load (the hit points take
// their pixel's radius; hp_r2 out) -> [the caller's newN / tao stand for the pass's gather] -> shared_accum
// -> shared_update -> shared_resolve.  Pixel arrays: r2[np], n[np], tau[3][np]; tao is [3][n_hp].
void dh_shared_pass(int64_t np, int w, double init_r2, double alpha, double photons, double* px_r2, double* px_n,
                    double* px_tau, int64_t n_hp, const int32_t* row, const int32_t* col, const int32_t* newN,
                    const double* tao, double* hp_r2_out, double* film) {
  std::vector<double> phi((size_t)np * 3, 0.0);
  std::vector<int32_t> m((size_t)np, 0);
  PxStats p;
  p.r2 = px_r2; p.n = px_n; p.tau = px_tau; p.phi = phi.data(); p.m = m.data(); p.np = np;
  const int T = bt(TPB);
  if (init_r2 > 0) launch(shared_init_kernel, std::make_tuple(p, init_r2), (np + T - 1) / T, T);
  std::vector<int32_t> iN((size_t)n_hp), inew((size_t)n_hp);
  std::vector<double> t((size_t)n_hp * 3);
  HpSoA s;
  memset(&s, 0, sizeof(s));
  s.row = const_cast<int32_t*>(row); s.col = const_cast<int32_t*>(col);
  s.r2 = hp_r2_out; s.N = iN.data(); s.newN = inew.data(); s.tao = t.data();
  launch(shared_load_kernel, std::make_tuple(s, (long long)n_hp, p, w), (n_hp + T - 1) / T, T);
  memcpy(inew.data(), newN, (size_t)n_hp * 4);   // the pass's photon rounds
  memcpy(t.data(), tao, (size_t)n_hp * 24);
  launch(shared_accum_kernel, std::make_tuple(s, (long long)n_hp, p, w), (n_hp + T - 1) / T, T);
  launch(shared_update_kernel, std::make_tuple(p, alpha), (np + T - 1) / T, T);
  if (film) launch(shared_resolve_kernel, std::make_tuple(p, photons, film), (np + T - 1) / T, T);
}
// tonemap_kernel (toInt + row flip, Raytracer.h:24-26, 460-470) and tex_expand_kernel (scene upload)
void dh_tonemap(const double* film, int w, int h, double div, uint8_t* bgr) {
  const long long n = (long long)w * h * 3;
  launch(tonemap_kernel, std::make_tuple(film, w, h, div, bgr), (n + 255) / 256, 256);
}
// all levels of the chain, untiled: out = level 0 [rows*cols] | level 1 | ...; returns the number of levels
int dh_tex_expand(const uint8_t* bgr, int rows, int cols, uint32_t* out, int32_t* dims /* [8][2] */) {
  DevTexture t;
  const std::vector<uint32_t> words = expand_texels(bgr, rows, cols, t);
  size_t o = 0;
  for (int l = 0; l < t.n_levels; l++) {
    const TexLevel& L = t.level[l];
    dims[2 * l] = L.rows; dims[2 * l + 1] = L.cols;
    for (int r = 0; r < L.rows; r++)
      for (int c = 0; c < L.cols; c++) out[o++] = L.texel[tex_tiled_index(r, c, L.tiles_x)];
  }
  return t.n_levels;
}
// bounds_kernel (build_index step 1): out9 = min[3], max[3], sum[3] of the hit-point positions
void dh_bounds(int64_t n, const HpRaw* raw, int blocks, double* out9) {
  unsigned long long mm[6];
  double sum[3] = {0, 0, 0};
  for (int k = 0; k < 3; k++) {
    mm[k] = ~0ull;
    mm[3 + k] = 0ull;
  }
  launch(bounds_kernel, std::make_tuple(raw, (long long)n, mm, sum), blocks, bt(TPB));
  for (int k = 0; k < 3; k++) {
    out9[k] = key_dbl(mm[k]);
    out9[3 + k] = key_dbl(mm[3 + k]);
    out9[6 + k] = sum[k];
  }
}
// items, slices, deposit visits, deposits kept by the cull, (hit point, deposit) pairs classified, pair slots
// issued (padding and idle lanes included), lanes that repeated a slice on the careful path
void dh_work_read(uint64_t* out7, int reset) {
  for (int k = 0; k < WK_N; k++) {
    out7[k] = dh_work[k];
    if (reset) dh_work[k] = 0;
  }
}
void dh_set_one_warp_blocks(int on) { g_one_warp = on ? 1 : 0; }
int dh_sizeof_hpraw() { return (int)sizeof(HpRaw); }
int dh_sizeof_deposit() { return (int)sizeof(Deposit); }

}  // extern "C"
