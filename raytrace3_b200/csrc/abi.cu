// abi.cu — the extern "C" boundary of libsppm_b200.so (include/sppm_b200.h) and the
// host-side orchestration of one render pass.  No CPU compute path exists here: every
// entry point that produces results launches the sm_100a kernels of trace.cu / gather.cu.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>

#include "ctx.h"
#include "scene_prep.h"
#include "image_kernels.cuh"

using namespace sppm;

namespace {

thread_local std::string g_create_error;

int fail(sppm_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  else g_create_error = msg;
  return code;
}
int cuda_fail(sppm_ctx* c, cudaError_t e, const char* where) {
  return fail(c, SPPM_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
#define CU(c, x)                                        \
  do {                                                  \
    cudaError_t e__ = (x);                              \
    if (e__ != cudaSuccess) return cuda_fail(c, e__, #x); \
  } while (0)

struct Bind {  // make the context's device current for the duration of a call
  int prev = -1;
  explicit Bind(const sppm_ctx* c) {
    cudaGetDevice(&prev);
    if (prev != c->device) cudaSetDevice(c->device);
    else prev = -1;
  }
  ~Bind() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Photons traced per batch (all lights).  A photon pass larger than this is cut into equal batches that rotate over
// the lanes.  8 Mi: on dense hit-point maps the gather's per-unit set-up and its partially filled 32-deposit
// slices are amortised over four times the deposits of the 2 Mi batches of round 1, the step / Newton launches
// over four times the photons (stress-4k 316 -> 348 M photons/s, caustics-1080p 800 -> 851;
// profiles/r2_b_ab_photon_batch.txt); a lane's buffers are sized for its largest batch (19 GB of deposit
// records + 4.4 GB of queues at 8 Mi photons, sized for max_dep + 1 deposits per photon).
constexpr uint64_t PHOTON_BATCH = 1ull << 23;

unsigned long long* scalars(sppm_ctx* c) { return (unsigned long long*)c->dev_scalars.p; }

// Per-batch timings (trace / bin / gather) are taken with a ring of CUDA events and folded
// into the host counters when the ring is full or at the next natural sync point, so the
// photon pass itself never waits for the device.
void fold_timings(sppm_ctx* c) {
  if (c->a_pending) {
    cudaEventSynchronize(c->aev[c->a_pending - 1][1]);
    for (int i = 0; i < c->a_pending; i++) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, c->aev[i][0], c->aev[i][1]) == cudaSuccess) c->host_counters.ms_allreduce += ms;
    }
    c->a_pending = 0;
  }
  if (c->t_pending == 0) return;
  cudaEventSynchronize(c->tev[c->t_pending - 1][c->t_has_gather[c->t_pending - 1] ? 3 : 1]);
  for (int i = 0; i < c->t_pending; i++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->tev[i][0], c->tev[i][1]) == cudaSuccess) c->host_counters.ms_trace += ms;
    if (c->t_has_gather[i]) {
      if (cudaEventElapsedTime(&ms, c->tev[i][1], c->tev[i][2]) == cudaSuccess) c->host_counters.ms_sort += ms;
      if (cudaEventElapsedTime(&ms, c->tev[i][2], c->tev[i][3]) == cudaSuccess) c->host_counters.ms_gather += ms;
    }
  }
  c->t_pending = 0;
}

// The deposit list's device-side count is checked at sync points (it is sized for max_dep + 1 deposits per
// photon, the most Raytracer.h:127-133 allows, so this cannot fire; the wavefront tracer's (ray, patch) item
// list cannot overflow either: photons that find it full are finished by the persistent tracer).
int check_overflow(sppm_ctx* c) {
  for (int j = 0; j < sppm_ctx::MAX_LANES; j++) {
    Lane& L = c->lanes[j];
    if (!L.scalars.p || L.dep_cap <= 0) continue;
    unsigned long long m = 0;
    if (cudaStreamSynchronize(L.stream) != cudaSuccess ||
        cudaMemcpyAsync(&m, (unsigned long long*)L.scalars.p + 1, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess)
      return cuda_fail(c, cudaGetLastError(), "check_overflow");
    if (j == c->last_lane) c->n_dep = (int64_t)std::min<unsigned long long>(m, (unsigned long long)L.dep_cap);
    if ((int64_t)m > L.dep_cap) return fail(c, SPPM_ERR_CAPACITY, "deposit buffer overflow");
  }
  return SPPM_OK;
}

// FP64 FMA throughput probe (the roofline denominator of the Newton / gather kernels; not in
// MEASURED_PEAKS.json): 8 independent DFMA chains per thread.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__global__ void synth_pairs_kernel(DevScene s, long long n, uint64_t seed, double* rays6, int32_t* patch) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int gp = (int)(i % s.n_patch);
  const double* b = s.patchBox + (size_t)gp * 6;
  double u0, u1, u2, u3;
  draw2(seed, 1, (uint64_t)i, 0, 6, 0, 0, u0, u1);
  draw2(seed, 1, (uint64_t)i, 0, 6, 1, 0, u2, u3);
  const d3 tgt = mk(b[0] + (b[3] - b[0]) * u0, b[1] + (b[4] - b[1]) * u1, b[2] + (b[5] - b[2]) * u2);
  const d3 dirv = unit_rand_vec(seed, 2, (uint64_t)i, 6);
  const d3 org = tgt + dirv * (40.0 + 40.0 * u3);
  const d3 dir = get_unit(tgt - org);
  rays6[6 * i + 0] = org.x; rays6[6 * i + 1] = org.y; rays6[6 * i + 2] = org.z;
  rays6[6 * i + 3] = dir.x; rays6[6 * i + 4] = dir.y; rays6[6 * i + 5] = dir.z;
  patch[i] = gp;
}

// ---- NCCL through dlopen (torch ships libnccl.so.2; nothing is linked) -------------------
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi& nccl() {
  static NcclApi api;
  if (api.h || api.ok) return api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.h) break;
  }
  if (!api.h) return api;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.h, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.h, "ncclCommInitRank");
  api.AllReduce = (decltype(api.AllReduce))dlsym(api.h, "ncclAllReduce");
  api.AllGather = (decltype(api.AllGather))dlsym(api.h, "ncclAllGather");
  api.Broadcast = (decltype(api.Broadcast))dlsym(api.h, "ncclBroadcast");
  api.GroupStart = (decltype(api.GroupStart))dlsym(api.h, "ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.h, "ncclGroupEnd");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.GroupStart && api.GroupEnd && api.CommDestroy;
  return api;
}

void free_scene(sppm_ctx* c) {
  for (auto& b : c->scene_bufs) dev_free(b);
  c->scene_bufs.clear();
  c->scene_ok = false;
}
// Scene arrays live in grow-only device buffers (slot i of scene_bufs) that are reused by the
// next upload: re-uploading a scene of the same shape costs only the copies.
template <class T>
cudaError_t upload(sppm_ctx* c, size_t slot, const T* host, size_t count, const T** out) {
  if (c->scene_bufs.size() <= slot) c->scene_bufs.resize(slot + 1);
  DevBuf& b = c->scene_bufs[slot];
  cudaError_t e = dev_reserve(b, std::max<size_t>(count * sizeof(T), 16));
  if (e) return e;
  if (count) e = cudaMemcpyAsync(b.p, host, count * sizeof(T), cudaMemcpyHostToDevice, c->stream);
  *out = (const T*)b.p;
  return e;
}

int ensure_deposit_capacity(sppm_ctx* c, Lane& L, int64_t cap) {
  const size_t slots = ((size_t)(c->index_ok ? c->n_slots : 0) + 4) * 4;
  CU(c, dev_reserve(L.slot_count, slots));
  CU(c, dev_reserve(L.dep_start, slots));
  if (cap <= L.dep_cap) return SPPM_OK;
  CU(c, dev_reserve(L.dep_raw, (size_t)cap * sizeof(Deposit)));
  CU(c, dev_reserve(L.dep_key, (size_t)cap * 4));
  CU(c, dev_reserve(L.dep_rank, (size_t)cap * 4));
  L.dep_cap = cap;
  return SPPM_OK;
}
// every later trace must see the scene / hit-point map that the main stream just produced
cudaError_t mark_setup(sppm_ctx* c) { return cudaEventRecord(c->ev_setup, c->stream); }

// Order the main stream after whatever the communication stream still holds (allreduce + update of the last
// round): every entry point that reads or rebuilds hit-point state on the main stream starts with this.
cudaError_t join_comm(sppm_ctx* c) {
  if (!c->comm_busy) return cudaSuccess;
  cudaError_t e = cudaStreamWaitEvent(c->stream, c->ev_comm, 0);
  if (e == cudaSuccess) c->comm_busy = false;
  return e;
}
bool use_async_comm(const sppm_ctx* c) {
  return c->async_comm && c->nccl_comm && c->nranks > 1 && c->comm_stream && c->prm.update_mode == SPPM_UPDATE_REFERENCE;
}

// canonical export order = ascending okey
std::vector<int64_t> canonical_order(const std::vector<uint64_t>& okey) {
  std::vector<int64_t> ord(okey.size());
  std::iota(ord.begin(), ord.end(), 0);
  std::sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b) { return okey[a] < okey[b]; });
  return ord;
}

}  // namespace

extern "C" {

int sppm_abi_version(void) { return SPPM_ABI_VERSION; }

const char* sppm_last_error(const sppm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int sppm_default_params(SppmParams* p) {
  if (!p) return SPPM_ERR_INVALID;
  p->init_r2 = 2.0;
  p->alpha = 0.7;
  p->max_dep = 13;
  p->newton_iters = 10;
  p->newton_restarts = 50;
  p->update_mode = SPPM_UPDATE_REFERENCE;
  p->newton_mode = SPPM_NEWTON_EXACT;
  p->trace_mode = SPPM_TRACE_WAVEFRONT;
  p->seed = 0x5EED5EEDull;
  p->gather_mode = SPPM_GATHER_FILTERED;
  p->count_pairs = 1;
  p->math_mode = SPPM_MATH_EXACT;
  p->tex_lod = 0;
  return SPPM_OK;
}

int sppm_ctx_create(int device, sppm_ctx** out) {
  if (!out) return fail(nullptr, SPPM_ERR_INVALID, "out is null");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return fail(nullptr, SPPM_ERR_NO_DEVICE,
                std::string("no CUDA device (this library has no CPU path): ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, SPPM_ERR_INVALID, "device index out of range");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device))) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(nullptr, SPPM_ERR_NO_DEVICE,
                std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
                    "; libsppm_b200 carries sm_100a code only");
  sppm_ctx* c = new sppm_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  sppm_default_params(&c->prm);
  if ((e = cudaSetDevice(device))) {
    delete c;
    return cuda_fail(nullptr, e, "cudaSetDevice");
  }
  bool ok;
  {  // SPPM_MAIN_PRIORITY=1 (A/B knob): the main stream (gathers, update) above the lanes' traces
    int lo_p = 0, hi_p = 0;
    cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
    const char* e = getenv("SPPM_MAIN_PRIORITY");
    ok = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, e && atoi(e) ? hi_p : 0) == cudaSuccess;
  }
  for (int i = 0; i < 8 && ok; i++) ok = cudaEventCreate(&c->ev[i]) == cudaSuccess;
  for (int i = 0; i < sppm_ctx::TIMING_RING && ok; i++)
    for (int k = 0; k < 4 && ok; k++) ok = cudaEventCreate(&c->tev[i][k]) == cudaSuccess;
  for (int i = 0; i < sppm_ctx::TIMING_RING && ok; i++)
    for (int k = 0; k < 2 && ok; k++) ok = cudaEventCreate(&c->aev[i][k]) == cudaSuccess;
  ok = ok && dev_reserve(c->dev_scalars, 64) == cudaSuccess && dev_reserve(c->counters, sizeof(DevCounters)) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&c->ev_setup, cudaEventDisableTiming) == cudaSuccess;
  {
    int lo_p = 0, hi_p = 0;  // the exchange gets the highest priority: its few CTAs should not queue behind a trace
    cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
    ok = ok && cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi_p) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_gathers_done, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->ev_acc_free[0], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->ev_acc_free[1], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->ev_comm, cudaEventDisableTiming) == cudaSuccess;
    if (const char* e = getenv("SPPM_ASYNC_COMM")) c->async_comm = atoi(e) != 0;
  }
  if (const char* e = getenv("SPPM_LANES")) c->n_lanes = c->n_lanes_small = std::max(1, std::min((int)sppm_ctx::MAX_LANES, atoi(e)));
  if (const char* e = getenv("SPPM_LANES_SMALL")) c->n_lanes_small = std::max(1, std::min((int)sppm_ctx::MAX_LANES, atoi(e)));
  for (int j = 0; j < sppm_ctx::MAX_LANES && ok; j++) {
    Lane& L = c->lanes[j];
    ok = cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&L.ev_trace_done, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&L.ev_gather_done, cudaEventDisableTiming) == cudaSuccess &&
         dev_reserve(L.scalars, 64) == cudaSuccess && cudaMemset(L.scalars.p, 0, 64) == cudaSuccess;
  }
  if (ok) {
    cudaMemset(c->dev_scalars.p, 0, 64);
    cudaMemset(c->counters.p, 0, sizeof(DevCounters));
    ok = cudaEventRecord(c->ev_setup, c->stream) == cudaSuccess;
  }
  if (!ok) {
    e = cudaGetLastError();
    sppm_ctx_destroy(c);
    return cuda_fail(nullptr, e, "context allocation");
  }
  *out = c;
  return SPPM_OK;
}

void sppm_ctx_destroy(sppm_ctx* c) {
  if (!c) return;
  Bind b(c);
  for (int j = 0; j < sppm_ctx::MAX_LANES; j++)
    if (c->lanes[j].stream) cudaStreamSynchronize(c->lanes[j].stream);
  if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->nccl_comm && nccl().ok) nccl().CommDestroy((ncclComm_t)c->nccl_comm);
  for (cudaEvent_t e : {c->ev_gathers_done, c->ev_acc_free[0], c->ev_acc_free[1], c->ev_comm})
    if (e) cudaEventDestroy(e);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  for (int j = 0; j < sppm_ctx::MAX_LANES; j++) {
    Lane& L = c->lanes[j];
    DevBuf* lb[] = {&L.dep_raw, &L.dep_sorted, &L.dep_key, &L.dep_rank, &L.slot_count, &L.dep_start, &L.wf_store, &L.scalars, &L.cub_tmp};
    for (DevBuf* d : lb) dev_free(*d);
    L.prof.destroy();
    if (L.ev_trace_done) cudaEventDestroy(L.ev_trace_done);
    if (L.ev_gather_done) cudaEventDestroy(L.ev_gather_done);
    if (L.stream) cudaStreamDestroy(L.stream);
  }
  if (c->ev_setup) cudaEventDestroy(c->ev_setup);
  c->prof.destroy();
  free_scene(c);
  DevBuf* bufs[] = {&c->hp_raw, &c->soa_store, &c->cell_slot, &c->occ_cell, &c->occ_start, &c->occ_nbr, &c->units, &c->item_parts, &c->item_off, &c->px_store, &c->hp_raw_alt, &c->dep_key, &c->dep_key2,
                    &c->dep_idx, &c->dep_idx2, &c->scan_scratch, &c->cub_tmp, &c->dev_scalars, &c->counters, &c->film, &c->scratch, &c->hp_canon};
  for (DevBuf* d : bufs) dev_free(*d);
  for (int i = 0; i < 8; i++)
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (int i = 0; i < sppm_ctx::TIMING_RING; i++)
    for (int k = 0; k < 4; k++)
      if (c->tev[i][k]) cudaEventDestroy(c->tev[i][k]);
  for (int i = 0; i < sppm_ctx::TIMING_RING; i++)
    for (int k = 0; k < 2; k++)
      if (c->aev[i][k]) cudaEventDestroy(c->aev[i][k]);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int sppm_set_params(sppm_ctx* c, const SppmParams* p) {
  if (!c || !p) return SPPM_ERR_INVALID;
  if (p->max_dep < 0 || p->max_dep > 13) return fail(c, SPPM_ERR_INVALID, "max_dep must be in [0,13]");
  if (p->newton_mode < 0 || p->newton_mode > 2) return fail(c, SPPM_ERR_INVALID, "bad newton_mode");
  if (p->trace_mode != SPPM_TRACE_WAVEFRONT && p->trace_mode != SPPM_TRACE_PERSISTENT) return fail(c, SPPM_ERR_INVALID, "bad trace_mode");
  if (p->gather_mode != SPPM_GATHER_EXACT && p->gather_mode != SPPM_GATHER_FILTERED) return fail(c, SPPM_ERR_INVALID, "bad gather_mode");
  if (p->math_mode != SPPM_MATH_EXACT && p->math_mode != SPPM_MATH_FAST) return fail(c, SPPM_ERR_INVALID, "bad math_mode");
  if (p->newton_restarts < 1 || p->newton_restarts > 4095) return fail(c, SPPM_ERR_INVALID, "newton_restarts out of range");
  if (p->newton_iters < 1) return fail(c, SPPM_ERR_INVALID, "newton_iters must be >= 1");
  if (!(p->init_r2 > 0)) return fail(c, SPPM_ERR_INVALID, "init_r2 must be > 0");
  if (p->tex_lod < 0 || p->tex_lod >= TEX_MAX_LEVELS) return fail(c, SPPM_ERR_INVALID, "tex_lod out of range");
  c->prm = *p;
  c->scene.tex_lod = p->tex_lod;
  return SPPM_OK;
}

int sppm_scene_upload(sppm_ctx* c, const SppmSceneDesc* s) {
  if (!c || !s) return SPPM_ERR_INVALID;
  if (s->n_objects < 0 || s->n_objects > MAX_OBJECTS) return fail(c, SPPM_ERR_INVALID, "n_objects out of range (max 64)");
  if (s->n_patches < 0 || s->n_patches >= (1 << 16)) return fail(c, SPPM_ERR_INVALID, "n_patches out of range");
  if (s->n_lights < 0 || s->n_textures < 0) return fail(c, SPPM_ERR_INVALID, "negative light / texture count");
  if ((s->n_objects > 0 && !s->objects) || (s->n_patches > 0 && !s->patches) || (s->n_lights > 0 && !s->lights) ||
      (s->n_textures > 0 && !s->textures))
    return fail(c, SPPM_ERR_INVALID, "a scene array is null while its count is positive");
  Bind b(c);
  CU(c, cudaStreamSynchronize(c->stream));
  c->scene_ok = false;
  PreppedScene ps;
  if (const char* bad = prep_scene(s, ps)) return fail(c, SPPM_ERR_INVALID, bad);
  const std::vector<double>&patchP = ps.patchP, &patchBox = ps.patchBox, &patchDop = ps.patchDop, &patchSub = ps.patchSub, &patchA = ps.patchA;
  const std::vector<DevObject>& objs = ps.objs;
  std::vector<SppmLight> lights(s->lights, s->lights + s->n_lights);
  std::vector<DevTexture> tex(s->n_textures);
  for (int t = 0; t < s->n_textures; t++) {
    const SppmTexture& ht = s->textures[t];
    if (ht.rows <= 0 || ht.cols <= 0 || !ht.bgr8) return fail(c, SPPM_ERR_INVALID, "empty texture (Q18: the reference would read out of range)");
    const uint8_t* dev = nullptr;
    CU(c, upload(c, 8 + (size_t)t, ht.bgr8, (size_t)ht.rows * ht.cols * 3, &dev));
    tex[t].rows = ht.rows;
    tex[t].cols = ht.cols;
    tex[t].bgr = dev;
    // one 32-bit word per texel in 8 x 4 tiles for the tracers (scene.cuh tex_fetch) + the mip chain below it,
    // expanded and filtered on the device
    const size_t slot = 8 + (size_t)s->n_textures + (size_t)t, ntexel = (size_t)ht.rows * ht.cols;
    if (c->scene_bufs.size() <= slot) c->scene_bufs.resize(slot + 1);
    const size_t words = tex_layout_levels(ht.rows, ht.cols, nullptr, tex[t]);
    CU(c, dev_reserve(c->scene_bufs[slot], words * 4));
    tex_layout_levels(ht.rows, ht.cols, (const uint32_t*)c->scene_bufs[slot].p, tex[t]);
    tex[t].bgr = dev;
    tex_expand_kernel<<<(unsigned)((ntexel + 255) / 256), 256, 0, c->stream>>>(dev, ht.rows, ht.cols, tex[t].level[0].tiles_x,
                                                                              (uint32_t*)c->scene_bufs[slot].p);
    c->host_counters.launches++;
    for (int l = 1; l < tex[t].n_levels; l++) {
      const size_t nl = (size_t)tex[t].level[l].rows * tex[t].level[l].cols;
      tex_mip_kernel<<<(unsigned)((nl + 255) / 256), 256, 0, c->stream>>>(tex[t].level[l - 1], tex[t].level[l]);
      c->host_counters.launches++;
    }
  }
  DevScene ds;
  ds.n_obj = s->n_objects;
  ds.n_patch = s->n_patches;
  ds.n_light = s->n_lights;
  ds.n_tex = s->n_textures;
  ds.tex_lod = c->prm.tex_lod;
  CU(c, upload(c, 0, objs.data(), objs.size(), &ds.obj));
  CU(c, upload(c, 1, patchP.data(), patchP.size(), &ds.patchP));
  CU(c, upload(c, 2, patchBox.data(), patchBox.size(), &ds.patchBox));
  CU(c, upload(c, 3, lights.data(), lights.size(), &ds.lights));
  CU(c, upload(c, 4, tex.data(), tex.size(), &ds.tex));
  CU(c, upload(c, 5, patchA.data(), patchA.size(), &ds.patchA));
  CU(c, upload(c, 6, patchDop.data(), patchDop.size(), &ds.patchDop));
  CU(c, upload(c, 7, patchSub.data(), patchSub.size(), &ds.patchSub));
  ds.culled = &((DevCounters*)c->counters.p)->patch_culled;
  CU(c, cudaStreamSynchronize(c->stream));  // host vectors go out of scope
  c->scene = ds;
  CU(c, trace_configure(c));
  c->scene_ok = true;
  CU(c, mark_setup(c));
  return SPPM_OK;
}

int sppm_eye_pass_rows(sppm_ctx* c, const SppmCameraDesc* cam, uint64_t pass_id, int32_t row_begin, int32_t row_end,
                       int32_t append, int64_t* n_hitpoints) {
  if (!c || !cam) return SPPM_ERR_INVALID;
  if (!c->scene_ok) return fail(c, SPPM_ERR_INVALID, "sppm_scene_upload must be called first");
  if (cam->w <= 0 || cam->h <= 0) return fail(c, SPPM_ERR_INVALID, "bad film size");
  if (row_begin < 0 || row_end > cam->h || row_begin > row_end) return fail(c, SPPM_ERR_INVALID, "bad row range");
  Bind b(c);
  const int64_t npix = (int64_t)cam->w * cam->h;
  if (append && (c->film_w != cam->w || c->film_h != cam->h || !c->hp_raw.p))
    return fail(c, SPPM_ERR_INVALID, "append: the film size differs from the hit points already traced");
  // 4 hit points per pixel to start with (the shipped scene makes exactly one); a scene whose
  // materials branch more (diffuse + mirror + glass on many objects: up to 2^14 per pixel,
  // Raytracer.h:312-335) reports its true count through the counter, and a whole-film pass is then
  // traced again into a buffer of that size (the pass is a function of the inputs alone).
  const int64_t want_cap = npix * 4 + 1024;
  if (!append || c->hp_cap < want_cap) c->hp_cap = want_cap;
  CU(c, dev_reserve(c->hp_raw, (size_t)c->hp_cap * sizeof(HpRaw)));
  c->film_w = cam->w;
  c->film_h = cam->h;
  c->index_ok = false;
  unsigned long long n = 0;
  for (int attempt = 0;; attempt++) {
    if (!append) CU(c, cudaMemsetAsync(scalars(c), 0, 8, c->stream));
    CU(c, cudaEventRecord(c->ev[4], c->stream));
    CU(c, launch_eye_pass(c, *cam, pass_id, row_begin, row_end));
    CU(c, cudaEventRecord(c->ev[5], c->stream));
    CU(c, cudaMemcpyAsync(&n, scalars(c), 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if ((int64_t)n <= c->hp_cap) break;
    if (attempt > 0) return fail(c, SPPM_ERR_CAPACITY, "hit-point buffer overflow (the eye pass grew between two identical traces)");
    // n is the true count (the kernel counts past the capacity): trace the same rows again into a buffer of
    // that size.  Appended row ranges share the buffer with the rows traced before, which cannot be re-made
    // here: they are carried over and the device-side count is put back to where this call found it.
    const int64_t new_cap = (int64_t)n + 1024;
    if (append) {
      const unsigned long long before = (unsigned long long)c->n_hp;
      DevBuf grown;
      CU(c, dev_reserve(grown, (size_t)new_cap * sizeof(HpRaw)));
      if (before) CU(c, cudaMemcpyAsync(grown.p, c->hp_raw.p, (size_t)before * sizeof(HpRaw), cudaMemcpyDeviceToDevice, c->stream));
      CU(c, cudaMemcpyAsync(scalars(c), &before, 8, cudaMemcpyHostToDevice, c->stream));
      CU(c, cudaStreamSynchronize(c->stream));  // `before` is a stack variable; the old buffer is freed next
      dev_free(c->hp_raw);
      c->hp_raw = grown;
    } else {
      CU(c, dev_reserve(c->hp_raw, (size_t)new_cap * sizeof(HpRaw)));
    }
    c->hp_cap = new_cap;
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]);
  c->host_counters.ms_eye = ms;
  c->n_hp = (int64_t)n;
  if (n_hitpoints) *n_hitpoints = c->n_hp;
  return SPPM_OK;
}

int sppm_eye_pass(sppm_ctx* c, const SppmCameraDesc* cam, uint64_t pass_id, int64_t* n_hitpoints) {
  if (!c || !cam) return SPPM_ERR_INVALID;
  return sppm_eye_pass_rows(c, cam, pass_id, 0, cam->h, 0, n_hitpoints);
}

// Eye-pass sharding: every rank traced its rows; exchange the raw hit points so that
// all ranks hold the whole map.  Counts differ per rank, so the exchange is one 8-byte all-gather
// of the counts followed by a group of broadcasts (rank r's records into slot r of the result).
// build_index sorts by (pixel, eye-path node), so the map is identical to a one-GPU eye pass.
int sppm_allgather_hitpoints(sppm_ctx* c, int64_t* n_hitpoints) {
  if (!c) return SPPM_ERR_INVALID;
  if (!c->hp_raw.p) return fail(c, SPPM_ERR_INVALID, "sppm_eye_pass_rows must be called first");
  if (!c->nccl_comm || c->nranks <= 1) {
    if (n_hitpoints) *n_hitpoints = c->n_hp;
    return SPPM_OK;
  }
  Bind b(c);
  CU(c, join_comm(c));
  NcclApi& n = nccl();
  if (!n.AllGather || !n.Broadcast) return fail(c, SPPM_ERR_NCCL, "ncclAllGather / ncclBroadcast not found");
  const int R = c->nranks;
  CU(c, dev_reserve(c->scratch, (size_t)(R + 1) * 8 + 256));
  unsigned long long* d_counts = (unsigned long long*)c->scratch.p;
  const unsigned long long mine = (unsigned long long)c->n_hp;
  CU(c, cudaMemcpyAsync(d_counts + R, &mine, 8, cudaMemcpyHostToDevice, c->stream));
  ncclResult_t r = n.AllGather(d_counts + R, d_counts, 1, ncclUint64, (ncclComm_t)c->nccl_comm, c->stream);
  if (r != ncclSuccess) return fail(c, SPPM_ERR_NCCL, "ncclAllGather (hit-point counts) failed");
  std::vector<unsigned long long> counts((size_t)R);
  CU(c, cudaMemcpyAsync(counts.data(), d_counts, (size_t)R * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  unsigned long long total = 0;
  for (auto v : counts) total += v;
  // The receive buffer is sized from `total`, which every rank computes from the same gathered counts: no
  // rank can leave before the collective below while the others enter it (the ranks' own hp_cap differ when
  // one of them re-traced a branching eye pass into a larger buffer).
  const int64_t cap_all = std::max<int64_t>(c->hp_cap, (int64_t)total + 1024);
  DevBuf& all = c->hp_raw_alt;  // second raw buffer, kept across calls (no allocation in steady state)
  CU(c, dev_reserve(all, (size_t)cap_all * sizeof(HpRaw)));
  r = n.GroupStart();
  unsigned long long off = 0;
  for (int k = 0; k < R && r == ncclSuccess; k++) {
    HpRaw* dst = (HpRaw*)all.p + off;
    if (counts[k])
      r = n.Broadcast(k == c->rank ? c->hp_raw.p : (const void*)dst, dst, (size_t)counts[k] * sizeof(HpRaw), ncclChar, k,
                      (ncclComm_t)c->nccl_comm, c->stream);
    off += counts[k];
  }
  const ncclResult_t r2 = n.GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess)
    return fail(c, SPPM_ERR_NCCL, std::string("hit-point exchange: ") + (n.GetErrorString ? n.GetErrorString(r) : "error"));
  std::swap(c->hp_raw, c->hp_raw_alt);  // stream-ordered: later work on c->stream sees the gathered map
  c->hp_cap = cap_all;
  c->n_hp = (int64_t)total;
  const unsigned long long t = total;
  CU(c, cudaMemcpyAsync(scalars(c), &t, 8, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));  // t is a stack variable
  c->index_ok = false;
  if (n_hitpoints) *n_hitpoints = c->n_hp;
  return SPPM_OK;
}

int sppm_build_index(sppm_ctx* c) {
  if (!c) return SPPM_ERR_INVALID;
  Bind b(c);
  CU(c, join_comm(c));
  CU(c, cudaEventRecord(c->ev[4], c->stream));
  CU(c, build_index(c));
  CU(c, cudaEventRecord(c->ev[5], c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]);
  c->host_counters.ms_build = ms;
  CU(c, mark_setup(c));
  return SPPM_OK;
}

int sppm_photon_pass(sppm_ctx* c, uint64_t photon_begin, uint64_t photon_count, uint64_t pass_id) {
  if (!c) return SPPM_ERR_INVALID;
  if (!c->scene_ok) return fail(c, SPPM_ERR_INVALID, "sppm_scene_upload must be called first");
  if (!c->index_ok && !c->keep_all_deposits) return fail(c, SPPM_ERR_INVALID, "sppm_build_index must be called first");
  if (c->scene.n_light <= 0 || photon_count == 0) return SPPM_OK;
  if (photon_begin + photon_count > (1ull << 40)) return fail(c, SPPM_ERR_INVALID, "photon index exceeds 2^40");
  Bind b(c);
  // Batches of up to PHOTON_BATCH photons rotate over the context's lanes (own stream + buffers):
  // lane stream = trace + deposit binning, main stream = gather.  The gather of one batch
  // overlaps the traces of the next ones — across rounds too, since nothing here waits for the host.
  // (A round is not split further: halving a batch doubles the per-cell and per-launch overheads.)
  const bool gather = c->index_ok && c->n_hp > 0;
  const uint64_t n_light = (uint64_t)c->scene.n_light;
  uint64_t batch = PHOTON_BATCH;
  if (const char* e = getenv("SPPM_PHOTON_BATCH")) batch = std::max<uint64_t>(1u << 16, std::min<uint64_t>(1ull << 24, strtoull(e, nullptr, 10)));  // A/B knob
  uint64_t per_light_batch = std::max<uint64_t>(1, batch / n_light);
  const bool split = !c->keep_all_deposits;
  if (photon_count > per_light_batch) {  // equal batches
    const uint64_t nb = (photon_count + per_light_batch - 1) / per_light_batch;
    per_light_batch = (photon_count + nb - 1) / nb;
  }
  for (uint64_t off = 0; off < photon_count; off += per_light_batch) {
    const uint64_t cnt = std::min(per_light_batch, photon_count - off);
    // small batches (the reference's own round sizes) are chains of latency-bound launches: more of them in flight
    const int lanes_now = cnt * n_light <= sppm_ctx::SMALL_BATCH ? c->n_lanes_small : c->n_lanes;
    const int j = split ? (c->lane_rr++ % lanes_now) : 0;
    Lane& L = c->lanes[j];
    c->last_lane = j;
    const int64_t cap = (int64_t)(cnt * n_light) * (c->prm.max_dep + 1);
    int rc = ensure_deposit_capacity(c, L, cap);
    if (rc) return rc;
    if (c->t_pending == sppm_ctx::TIMING_RING) fold_timings(c);
    cudaEvent_t* te = c->tev[c->t_pending];
    CU(c, cudaStreamWaitEvent(L.stream, c->ev_setup, 0));        // scene + hit-point map are ready
    // The gather of the lane's previous batch reads dep_sorted / dep_start only, and those are written by
    // bin_deposits: the trace starts at once, only the binning waits for that gather (SPPM_LATE_GATHER_WAIT=0, the
    // A/B knob, makes the whole batch wait as before: same rate at 2 M-photon rounds, 4 % lower at 100 k).
    static const bool late_wait = [] { const char* e = getenv("SPPM_LATE_GATHER_WAIT"); return !e || atoi(e) != 0; }();
    if (!late_wait) CU(c, cudaStreamWaitEvent(L.stream, L.ev_gather_done, 0));
    CU(c, cudaMemsetAsync((unsigned long long*)L.scalars.p + 1, 0, 16, L.stream));  // deposit count, photon cursor
    CU(c, cudaMemsetAsync(L.slot_count.p, 0, ((size_t)(c->index_ok ? c->n_slots : 0) + 2) * 4, L.stream));
    CU(c, cudaEventRecord(te[0], L.stream));
    if (c->prm.trace_mode == SPPM_TRACE_PERSISTENT) CU(c, launch_photon_trace(c, L, photon_begin + off, cnt, pass_id));
    else CU(c, launch_photon_trace_wavefront(c, L, photon_begin + off, cnt, pass_id));
    if (late_wait) CU(c, cudaStreamWaitEvent(L.stream, L.ev_gather_done, 0));   // the lane's previous deposits were consumed
    if (gather) CU(c, bin_deposits(c, L, L.stream));
    CU(c, L.prof.mark(PK_END, L.stream));
    CU(c, cudaEventRecord(te[1], L.stream));
    CU(c, cudaEventRecord(L.ev_trace_done, L.stream));
    CU(c, cudaStreamWaitEvent(c->stream, L.ev_trace_done, 0));
    if (gather) CU(c, launch_gather(c, L, te[2], te[3]));
    CU(c, cudaEventRecord(L.ev_gather_done, c->stream));
    c->t_has_gather[c->t_pending] = gather;
    c->t_pending++;
    if (c->profiling) {  // measurement mode: the batch runs alone, its launches are timed one by one
      CU(c, cudaStreamSynchronize(L.stream));
      CU(c, cudaStreamSynchronize(c->stream));
      double b[PK_N] = {};
      L.prof.fold(b);
      c->prof.fold(b);
      SppmCounters& h = c->host_counters;
      h.ms_step += b[PK_STEP]; h.ms_newton += b[PK_NEWTON]; h.ms_spill += b[PK_SPILL];
      h.ms_bin_scan += b[PK_BIN_SCAN]; h.ms_bin_scatter += b[PK_BIN_SCATTER];
      h.ms_gather_prep += b[PK_GATHER_PREP]; h.ms_gather_kernel += b[PK_GATHER_KERNEL];
    }
  }
  return SPPM_OK;  // stream-ordered: nothing above waited for the device
}

int sppm_allreduce(sppm_ctx* c) {
  if (!c) return SPPM_ERR_INVALID;
  if (!c->nccl_comm || c->nranks <= 1 || c->n_hp == 0) return SPPM_OK;
  Bind b(c);
  NcclApi& n = nccl();
  if (c->a_pending == sppm_ctx::TIMING_RING) fold_timings(c);
  cudaEvent_t* ae = c->aev[c->a_pending];
  // REFERENCE update: the exchange and the update that follows run on the communication stream, behind the
  // round's gathers, while the main stream goes on with the next round (see ctx.h); otherwise in line
  const bool async = use_async_comm(c) && !c->profiling;
  cudaStream_t st = async ? c->comm_stream : c->stream;
  if (async) {
    CU(c, cudaEventRecord(c->ev_gathers_done, c->stream));
    CU(c, cudaStreamWaitEvent(c->comm_stream, c->ev_gathers_done, 0));
  } else {
    CU(c, join_comm(c));
  }
  CU(c, cudaEventRecord(ae[0], st));
  ncclResult_t r = n.GroupStart();
  if (r == ncclSuccess) r = n.AllReduce(c->soa.dN, c->soa.dN, (size_t)c->n_hp, ncclInt32, ncclSum, (ncclComm_t)c->nccl_comm, st);
  if (r == ncclSuccess)
    r = n.AllReduce(c->soa.dtao, c->soa.dtao, (size_t)c->n_hp * 3, ncclFloat64, ncclSum, (ncclComm_t)c->nccl_comm, st);
  ncclResult_t r2 = n.GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) return fail(c, SPPM_ERR_NCCL, std::string("ncclAllReduce: ") + (n.GetErrorString ? n.GetErrorString(r) : "error"));
  CU(c, cudaEventRecord(ae[1], st));
  c->a_pending++;
  if (async) {
    CU(c, cudaEventRecord(c->ev_comm, c->comm_stream));
    c->comm_pending = true;
    c->comm_busy = true;
  }
  return SPPM_OK;
}

int sppm_update(sppm_ctx* c) {
  if (!c) return SPPM_ERR_INVALID;
  if (!c->index_ok) return fail(c, SPPM_ERR_INVALID, "sppm_build_index must be called first");
  Bind b(c);
  if (c->comm_pending) {
    // the update of the round whose exchange is in flight: same stream, then the accumulator sets swap
    CU(c, launch_update(c, c->comm_stream));
    CU(c, cudaEventRecord(c->ev_acc_free[c->acc_cur], c->comm_stream));
    CU(c, cudaEventRecord(c->ev_comm, c->comm_stream));
    c->comm_pending = false;
    c->comm_busy = true;
    c->acc_cur ^= 1;
    c->soa.dN = c->dN_set[c->acc_cur];
    c->soa.dtao = c->dtao_set[c->acc_cur];
    CU(c, cudaStreamWaitEvent(c->stream, c->ev_acc_free[c->acc_cur], 0));  // the update that last emptied this set
    if (c->prm.update_mode != SPPM_UPDATE_REFERENCE) CU(c, join_comm(c));  // the mode changed in between: radii may move
    return SPPM_OK;
  }
  CU(c, join_comm(c));
  CU(c, c->prof.mark(PK_UPDATE, c->stream));
  CU(c, launch_update(c, c->stream));
  if (c->profiling) {
    CU(c, c->prof.mark(PK_END, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    double bk[PK_N] = {};
    c->prof.fold(bk);
    c->host_counters.ms_update += bk[PK_UPDATE];
  }
  return SPPM_OK;
}

int sppm_resolve(sppm_ctx* c, int32_t cnt, int32_t onetime, double* film_hw3) {
  if (!c || !film_hw3) return SPPM_ERR_INVALID;
  if (!c->index_ok) return fail(c, SPPM_ERR_INVALID, "sppm_build_index must be called first");
  Bind b(c);
  CU(c, join_comm(c));
  CU(c, cudaEventRecord(c->ev[4], c->stream));
  CU(c, launch_resolve(c, cnt, onetime));
  CU(c, cudaEventRecord(c->ev[5], c->stream));
  CU(c, cudaMemcpyAsync(film_hw3, c->film.p, (size_t)c->film_w * c->film_h * 24, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]);
  c->host_counters.ms_resolve = ms;
  fold_timings(c);
  return check_overflow(c);
}

int sppm_shared_begin(sppm_ctx* c, int32_t w, int32_t h) {
  if (!c || w <= 0 || h <= 0) return SPPM_ERR_INVALID;
  Bind b(c);
  CU(c, shared_begin(c, w, h));
  return SPPM_OK;
}

int sppm_shared_load(sppm_ctx* c) {
  if (!c) return SPPM_ERR_INVALID;
  if (!c->index_ok) return fail(c, SPPM_ERR_INVALID, "sppm_build_index must be called first");
  if (!c->px_store.p || c->px_w != c->film_w || c->px_h != c->film_h)
    return fail(c, SPPM_ERR_INVALID, "sppm_shared_begin must be called with the film size of the eye pass");
  Bind b(c);
  CU(c, join_comm(c));
  CU(c, shared_load(c));
  CU(c, mark_setup(c));  // the lanes read the hit-point radii
  return SPPM_OK;
}

int sppm_shared_fold(sppm_ctx* c, uint64_t photons_emitted) {
  if (!c) return SPPM_ERR_INVALID;
  if (!c->index_ok || !c->px_store.p) return fail(c, SPPM_ERR_INVALID, "sppm_shared_load must be called first");
  Bind b(c);
  CU(c, join_comm(c));
  CU(c, shared_fold(c));
  c->px_photons += photons_emitted;
  return SPPM_OK;
}

int sppm_shared_resolve(sppm_ctx* c, double* film_hw3) {
  if (!c || !film_hw3) return SPPM_ERR_INVALID;
  if (!c->px_store.p) return fail(c, SPPM_ERR_INVALID, "sppm_shared_begin must be called first");
  Bind b(c);
  CU(c, join_comm(c));
  CU(c, shared_resolve(c));
  CU(c, cudaMemcpyAsync(film_hw3, c->film.p, (size_t)c->px_w * c->px_h * 24, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  fold_timings(c);
  return check_overflow(c);
}

int sppm_shared_export(sppm_ctx* c, double* r2_hw, double* n_hw, double* tau3, uint64_t* photons_emitted) {
  if (!c) return SPPM_ERR_INVALID;
  if (!c->px_store.p) return fail(c, SPPM_ERR_INVALID, "sppm_shared_begin must be called first");
  Bind b(c);
  CU(c, join_comm(c));
  CU(c, shared_export(c, r2_hw, n_hw, tau3));
  if (photons_emitted) *photons_emitted = c->px_photons;
  return SPPM_OK;
}

int sppm_shared_import(sppm_ctx* c, const double* r2_hw, const double* n_hw, const double* tau3, uint64_t photons_emitted) {
  if (!c || !r2_hw || !n_hw || !tau3) return SPPM_ERR_INVALID;
  if (!c->px_store.p) return fail(c, SPPM_ERR_INVALID, "sppm_shared_begin must be called first");
  Bind b(c);
  CU(c, join_comm(c));
  CU(c, shared_import(c, r2_hw, n_hw, tau3));
  c->px_photons = photons_emitted;
  return SPPM_OK;
}

int sppm_tonemap(sppm_ctx* c, const double* film_hw3, int32_t w, int32_t h, double div, uint8_t* bgr_hw3) {
  if (!c || !film_hw3 || !bgr_hw3 || w <= 0 || h <= 0) return SPPM_ERR_INVALID;
  Bind b(c);
  const size_t n = (size_t)w * h * 3;
  CU(c, dev_reserve(c->film, n * 8));
  CU(c, dev_reserve(c->scratch, n));
  CU(c, cudaMemcpyAsync(c->film.p, film_hw3, n * 8, cudaMemcpyHostToDevice, c->stream));
  tonemap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((const double*)c->film.p, w, h, div, (uint8_t*)c->scratch.p);
  c->host_counters.launches++;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(bgr_hw3, c->scratch.p, n, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return SPPM_OK;
}

int sppm_host_pin(void* ptr, uint64_t bytes) {
  if (!ptr || bytes == 0) return SPPM_ERR_INVALID;
  const cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();  // not sticky; the caller goes on with pageable memory
    g_create_error = std::string("cudaHostRegister: ") + cudaGetErrorString(e);
    return SPPM_ERR_CUDA;
  }
  return SPPM_OK;
}
int sppm_host_unpin(void* ptr) {
  if (!ptr) return SPPM_ERR_INVALID;
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return SPPM_ERR_CUDA;
  }
  return SPPM_OK;
}

int sppm_comm_unique_id(void* id128) {
  if (!id128) return SPPM_ERR_INVALID;
  NcclApi& n = nccl();
  if (!n.ok) return SPPM_ERR_NCCL;
  ncclUniqueId id;
  if (n.GetUniqueId(&id) != ncclSuccess) return SPPM_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return SPPM_OK;
}

int sppm_comm_init_rank(sppm_ctx* c, int rank, int nranks, const void* id128) {
  if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return SPPM_ERR_INVALID;
  NcclApi& n = nccl();
  if (!n.ok) return fail(c, SPPM_ERR_NCCL, "libnccl.so.2 not found (dlopen)");
  Bind b(c);
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm;
  ncclResult_t r = n.CommInitRank(&comm, nranks, id, rank);
  if (r != ncclSuccess) return fail(c, SPPM_ERR_NCCL, std::string("ncclCommInitRank: ") + (n.GetErrorString ? n.GetErrorString(r) : "error"));
  c->nccl_comm = comm;
  c->nranks = nranks;
  c->rank = rank;
  return SPPM_OK;
}

int sppm_accum_device_ptrs(sppm_ctx* c, void** dN, void** dtao, int64_t* n) {
  if (!c) return SPPM_ERR_INVALID;
  if (!c->index_ok) return fail(c, SPPM_ERR_INVALID, "sppm_build_index must be called first");
  if (dN) *dN = c->soa.dN;
  if (dtao) *dtao = c->soa.dtao;
  if (n) *n = c->n_hp;
  return SPPM_OK;
}

int sppm_device_synchronize(sppm_ctx* c) {
  if (!c) return SPPM_ERR_INVALID;
  Bind b(c);
  CU(c, join_comm(c));
  CU(c, cudaStreamSynchronize(c->stream));
  fold_timings(c);
  return check_overflow(c);
}

void* sppm_stream_handle(sppm_ctx* c) { return c ? (void*)c->stream : nullptr; }

/* CUDA-event stopwatch on the context's stream (slots 0..1 of ev[6..7]). */
int sppm_timer_record(sppm_ctx* c, int slot) {
  if (!c || slot < 0 || slot > 1) return SPPM_ERR_INVALID;
  Bind b(c);
  CU(c, join_comm(c));  // the stopwatch covers the exchange in flight
  CU(c, cudaEventRecord(c->ev[6 + slot], c->stream));
  return SPPM_OK;
}
int sppm_timer_elapsed_ms(sppm_ctx* c, double* ms) {
  if (!c || !ms) return SPPM_ERR_INVALID;
  Bind b(c);
  CU(c, cudaEventSynchronize(c->ev[7]));
  float f = 0;
  CU(c, cudaEventElapsedTime(&f, c->ev[6], c->ev[7]));
  *ms = f;
  return SPPM_OK;
}

// ---- test / measurement hooks ------------------------------------------------------------
int sppm_intersect_batch(sppm_ctx* c, const double* rays6, int64_t n, const uint64_t* ids, uint64_t pass_id,
                         int32_t node, SppmHit* out) {
  if (!c || !rays6 || !out || n < 0) return SPPM_ERR_INVALID;
  if (!c->scene_ok) return fail(c, SPPM_ERR_INVALID, "sppm_scene_upload must be called first");
  if (n == 0) return SPPM_OK;
  Bind b(c);
  DevBuf dr, di, dout;
  int rc = SPPM_OK;
  cudaError_t e;
  do {
    if ((e = dev_reserve(dr, (size_t)n * 48))) break;
    if ((e = dev_reserve(dout, (size_t)n * sizeof(SppmHit)))) break;
    if (ids && (e = dev_reserve(di, (size_t)n * 8))) break;
    if ((e = cudaMemcpyAsync(dr.p, rays6, (size_t)n * 48, cudaMemcpyHostToDevice, c->stream))) break;
    if (ids && (e = cudaMemcpyAsync(di.p, ids, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream))) break;
    if ((e = launch_intersect(c, (const double*)dr.p, n, ids ? (const uint64_t*)di.p : nullptr, pass_id, node, (SppmHit*)dout.p))) break;
    if ((e = cudaMemcpyAsync(out, dout.p, (size_t)n * sizeof(SppmHit), cudaMemcpyDeviceToHost, c->stream))) break;
    e = cudaStreamSynchronize(c->stream);
  } while (0);
  if (e) rc = cuda_fail(c, e, "sppm_intersect_batch");
  dev_free(dr);
  dev_free(di);
  dev_free(dout);
  return rc;
}

int sppm_bezier_patch_batch(sppm_ctx* c, const double* rays6, const int32_t* patch_idx, int64_t n, const double* seeds_uv,
                            double* tuv_out) {
  if (!c || !rays6 || !patch_idx || !tuv_out || n < 0) return SPPM_ERR_INVALID;
  if (!c->scene_ok) return fail(c, SPPM_ERR_INVALID, "sppm_scene_upload must be called first");
  for (int64_t i = 0; i < n; i++)
    if (patch_idx[i] < 0 || patch_idx[i] >= c->scene.n_patch) return fail(c, SPPM_ERR_INVALID, "patch index out of range");
  if (n == 0) return SPPM_OK;
  Bind b(c);
  DevBuf dr, dp, ds, dt;
  int rc = SPPM_OK;
  cudaError_t e;
  const size_t sbytes = (size_t)n * c->prm.newton_restarts * 16;
  do {
    if ((e = dev_reserve(dr, (size_t)n * 48))) break;
    if ((e = dev_reserve(dp, (size_t)n * 4))) break;
    if ((e = dev_reserve(dt, (size_t)n * 24))) break;
    if (seeds_uv && (e = dev_reserve(ds, sbytes))) break;
    if ((e = cudaMemcpyAsync(dr.p, rays6, (size_t)n * 48, cudaMemcpyHostToDevice, c->stream))) break;
    if ((e = cudaMemcpyAsync(dp.p, patch_idx, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream))) break;
    if (seeds_uv && (e = cudaMemcpyAsync(ds.p, seeds_uv, sbytes, cudaMemcpyHostToDevice, c->stream))) break;
    if ((e = launch_patch_batch(c, (const double*)dr.p, (const int32_t*)dp.p, n, seeds_uv ? (const double*)ds.p : nullptr, (double*)dt.p))) break;
    if ((e = cudaMemcpyAsync(tuv_out, dt.p, (size_t)n * 24, cudaMemcpyDeviceToHost, c->stream))) break;
    e = cudaStreamSynchronize(c->stream);
  } while (0);
  if (e) rc = cuda_fail(c, e, "sppm_bezier_patch_batch");
  dev_free(dr);
  dev_free(dp);
  dev_free(ds);
  dev_free(dt);
  return rc;
}

int sppm_bezier_patch_bench(sppm_ctx* c, int64_t n, int32_t reps, uint64_t seed, double* ms_per_rep) {
  if (!c || n <= 0 || reps <= 0 || !ms_per_rep) return SPPM_ERR_INVALID;
  if (!c->scene_ok || c->scene.n_patch <= 0) return fail(c, SPPM_ERR_INVALID, "a scene with Bezier patches must be uploaded first");
  Bind b(c);
  DevBuf dr, dp, dt;
  int rc = SPPM_OK;
  cudaError_t e;
  c->newton_cull = false;  // the microbenchmark times full 50 x 10 solves, not the 13-DOP shortcut
  do {
    if ((e = dev_reserve(dr, (size_t)n * 48))) break;
    if ((e = dev_reserve(dp, (size_t)n * 4))) break;
    if ((e = dev_reserve(dt, (size_t)n * 24))) break;
    synth_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->scene, n, seed, (double*)dr.p, (int32_t*)dp.p);
    if ((e = cudaGetLastError())) break;
    if ((e = launch_patch_batch(c, (const double*)dr.p, (const int32_t*)dp.p, n, nullptr, (double*)dt.p))) break;  // warm-up
    if ((e = cudaEventRecord(c->ev[4], c->stream))) break;
    for (int r = 0; r < reps && !e; r++) e = launch_patch_batch(c, (const double*)dr.p, (const int32_t*)dp.p, n, nullptr, (double*)dt.p);
    if (e) break;
    if ((e = cudaEventRecord(c->ev[5], c->stream))) break;
    if ((e = cudaStreamSynchronize(c->stream))) break;
    float ms = 0;
    if ((e = cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]))) break;
    *ms_per_rep = ms / reps;
  } while (0);
  c->newton_cull = true;
  if (e) rc = cuda_fail(c, e, "sppm_bezier_patch_bench");
  dev_free(dr);
  dev_free(dp);
  dev_free(dt);
  return rc;
}

int sppm_fp64_peak(sppm_ctx* c, double* tflops) {
  if (!c || !tflops) return SPPM_ERR_INVALID;
  Bind b(c);
  const int blocks = c->sm_count * 8, iters = 1 << 14;
  CU(c, dev_reserve(c->scratch, (size_t)blocks * 256 * 8));
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    CU(c, cudaEventRecord(c->ev[4], c->stream));
    fp64_peak_kernel<<<blocks, 256, 0, c->stream>>>((double*)c->scratch.p, iters, 0.999999, 1e-6);
    CU(c, cudaEventRecord(c->ev[5], c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    float ms = 0;
    CU(c, cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]));
    const double tf = 2.0 * 8 * iters * (double)blocks * 256 / (ms * 1e-3) * 1e-12;
    if (rep > 0 && tf > best) best = tf;
  }
  *tflops = best;
  return SPPM_OK;
}

int sppm_import_hitpoints(sppm_ctx* c, int64_t n, const double* pos3, const double* n3, const double* wgt3,
                          const double* r2, const int32_t* pix, int32_t w, int32_t h) {
  if (!c || n < 0 || w <= 0 || h <= 0) return SPPM_ERR_INVALID;
  if (n > 0 && (!pos3 || !n3 || !wgt3 || !r2 || !pix)) return SPPM_ERR_INVALID;
  Bind b(c);
  CU(c, join_comm(c));
  std::vector<HpRaw> raw((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    HpRaw& r = raw[i];
    for (int k = 0; k < 3; k++) {
      r.pos[k] = pos3[3 * i + k];
      r.n[k] = n3[3 * i + k];
      r.wgt[k] = wgt3[3 * i + k];
    }
    r.r2 = r2[i];
    r.row = pix[2 * i];
    r.col = pix[2 * i + 1];
    if (r.row < 0 || r.row >= h || r.col < 0 || r.col >= w) return fail(c, SPPM_ERR_INVALID, "hit-point pixel outside the film");
    r.okey = (uint64_t)i;
  }
  c->hp_cap = std::max<int64_t>(n, 1);
  CU(c, dev_reserve(c->hp_raw, (size_t)c->hp_cap * sizeof(HpRaw)));
  if (n) CU(c, cudaMemcpyAsync(c->hp_raw.p, raw.data(), (size_t)n * sizeof(HpRaw), cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->n_hp = n;
  c->film_w = w;
  c->film_h = h;
  c->index_ok = false;
  return SPPM_OK;
}

int sppm_export_hitpoints(sppm_ctx* c, int64_t cap, double* pos3, double* n3, double* wgt3, double* r2, int32_t* pix,
                          uint32_t* node, int64_t* n_out) {
  if (!c) return SPPM_ERR_INVALID;
  Bind b(c);
  CU(c, join_comm(c));
  const int64_t n = c->n_hp;
  if (n_out) *n_out = n;
  if (n == 0) return SPPM_OK;
  std::vector<HpRaw> raw((size_t)n);
  CU(c, cudaMemcpyAsync(raw.data(), c->hp_raw.p, (size_t)n * sizeof(HpRaw), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  std::vector<uint64_t> okey((size_t)n);
  for (int64_t i = 0; i < n; i++) okey[i] = raw[i].okey;
  const std::vector<int64_t> ord = canonical_order(okey);
  for (int64_t j = 0; j < std::min(n, cap); j++) {
    const HpRaw& r = raw[ord[j]];
    for (int k = 0; k < 3; k++) {
      if (pos3) pos3[3 * j + k] = r.pos[k];
      if (n3) n3[3 * j + k] = r.n[k];
      if (wgt3) wgt3[3 * j + k] = r.wgt[k];
    }
    if (r2) r2[j] = r.r2;
    if (pix) {
      pix[2 * j] = r.row;
      pix[2 * j + 1] = r.col;
    }
    if (node) node[j] = (uint32_t)(r.okey & 0xfffffu);
  }
  return SPPM_OK;
}

int sppm_gather_deposits(sppm_ctx* c, int64_t m, const double* x3, const double* n3, const double* flux3) {
  if (!c || m < 0) return SPPM_ERR_INVALID;
  if (!c->index_ok) return fail(c, SPPM_ERR_INVALID, "sppm_build_index must be called first");
  if (m == 0) return SPPM_OK;
  if (!x3 || !n3 || !flux3) return SPPM_ERR_INVALID;
  Bind b(c);
  CU(c, join_comm(c));
  Lane& L = c->lanes[0];
  c->last_lane = 0;
  CU(c, cudaStreamSynchronize(L.stream));
  int rc = ensure_deposit_capacity(c, L, m);
  if (rc) return rc;
  std::vector<Deposit> d((size_t)m);
  for (int64_t i = 0; i < m; i++) {
    for (int k = 0; k < 3; k++) {
      d[i].x[k] = x3[3 * i + k];
      d[i].n[k] = n3[3 * i + k];
      d[i].flux[k] = flux3[3 * i + k];
    }
    d[i].meta = (uint64_t)i;
  }
  CU(c, cudaMemcpyAsync(L.dep_raw.p, d.data(), (size_t)m * sizeof(Deposit), cudaMemcpyHostToDevice, c->stream));
  const unsigned long long m64 = (unsigned long long)m;
  CU(c, cudaMemcpyAsync((unsigned long long*)L.scalars.p + 1, &m64, 8, cudaMemcpyHostToDevice, c->stream));
  if (c->n_hp > 0) {
    if (c->t_pending == sppm_ctx::TIMING_RING) fold_timings(c);
    cudaEvent_t* te = c->tev[c->t_pending];
    CU(c, cudaMemsetAsync(L.slot_count.p, 0, ((size_t)c->n_slots + 2) * 4, c->stream));
    CU(c, cudaEventRecord(te[0], c->stream));
    CU(c, key_host_deposits(c, L, m, c->stream));
    CU(c, bin_deposits(c, L, c->stream));
    CU(c, cudaEventRecord(te[1], c->stream));
    CU(c, launch_gather(c, L, te[2], te[3]));
    CU(c, cudaEventRecord(L.ev_gather_done, c->stream));
    c->t_has_gather[c->t_pending] = true;
    c->t_pending++;
  }
  CU(c, cudaStreamSynchronize(c->stream));
  fold_timings(c);
  return SPPM_OK;
}

int sppm_export_stats(sppm_ctx* c, int64_t cap, int32_t* newN, int32_t* N, double* r2, double* tao3) {
  if (!c) return SPPM_ERR_INVALID;
  if (!c->index_ok) return fail(c, SPPM_ERR_INVALID, "sppm_build_index must be called first");
  Bind b(c);
  CU(c, join_comm(c));
  const int64_t n = c->n_hp;
  if (n == 0) return SPPM_OK;
  std::vector<uint64_t> okey((size_t)n);
  std::vector<int32_t> hN((size_t)n), hnewN((size_t)n);
  std::vector<double> hr2((size_t)n), htao((size_t)n * 3);
  CU(c, cudaMemcpyAsync(okey.data(), c->soa.okey, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(hN.data(), c->soa.N, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(hnewN.data(), c->soa.newN, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(hr2.data(), c->soa.r2, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(htao.data(), c->soa.tao, (size_t)n * 24, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  const std::vector<int64_t> ord = canonical_order(okey);
  for (int64_t j = 0; j < std::min(n, cap); j++) {
    const int64_t i = ord[j];
    if (newN) newN[j] = hnewN[i];
    if (N) N[j] = hN[i];
    if (r2) r2[j] = hr2[i];
    if (tao3)
      for (int k = 0; k < 3; k++) tao3[3 * j + k] = htao[(size_t)k * n + i];
  }
  return SPPM_OK;
}

int sppm_keep_deposits(sppm_ctx* c, int32_t keep_all) {
  if (!c) return SPPM_ERR_INVALID;
  c->keep_all_deposits = keep_all != 0;
  return SPPM_OK;
}

int sppm_export_deposits(sppm_ctx* c, int64_t cap, double* x3, double* n3, double* flux3, uint64_t* photon_id,
                         int32_t* bounce, int64_t* m_out) {
  if (!c) return SPPM_ERR_INVALID;
  Bind b(c);
  { int rc = check_overflow(c); if (rc) return rc; }  // also refreshes n_dep from the device counter
  const int64_t m = c->n_dep;
  if (m_out) *m_out = m;
  if (m == 0) return SPPM_OK;
  std::vector<Deposit> d((size_t)m);
  CU(c, cudaMemcpyAsync(d.data(), c->lanes[c->last_lane].dep_raw.p, (size_t)m * sizeof(Deposit), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  // canonical order: (photon id, bounce)
  std::vector<int64_t> ord((size_t)m);
  std::iota(ord.begin(), ord.end(), 0);
  auto keyf = [&](int64_t i) { return ((d[i].meta & 0x00ffffffffffffffull) << 8) | (d[i].meta >> 56); };
  std::sort(ord.begin(), ord.end(), [&](int64_t a, int64_t bb) { return keyf(a) < keyf(bb); });
  for (int64_t j = 0; j < std::min(m, cap); j++) {
    const Deposit& r = d[ord[j]];
    for (int k = 0; k < 3; k++) {
      if (x3) x3[3 * j + k] = r.x[k];
      if (n3) n3[3 * j + k] = r.n[k];
      if (flux3) flux3[3 * j + k] = r.flux[k];
    }
    if (photon_id) photon_id[j] = r.meta & 0x00ffffffffffffffull;
    if (bounce) bounce[j] = (int32_t)(r.meta >> 56);
  }
  return SPPM_OK;
}

int sppm_get_counters(sppm_ctx* c, SppmCounters* out) {
  if (!c || !out) return SPPM_ERR_INVALID;
  Bind b(c);
  CU(c, join_comm(c));
  CU(c, cudaStreamSynchronize(c->stream));
  fold_timings(c);
  DevCounters d;
  CU(c, cudaMemcpy(&d, c->counters.p, sizeof(d), cudaMemcpyDeviceToHost));
  SppmCounters& h = c->host_counters;
  h.photons = d.photons;
  h.isect_queries = d.isect_queries;
  h.patch_solves = d.patch_solves;
  h.patch_culled = d.patch_culled;
  h.spilled_photons = d.overflow;
  h.deposits = d.deposits;
  h.deposits_kept = d.deposits_kept;
  h.cand_pairs = d.cand_pairs;
  h.acc_pairs = d.acc_pairs;
  h.hit_points = (uint64_t)c->n_hp;
  h.grid_cells = (uint64_t)c->grid.n_cells;
  *out = h;
  return SPPM_OK;
}

int sppm_set_profiling(sppm_ctx* c, int32_t on) {
  if (!c) return SPPM_ERR_INVALID;
  c->profiling = on != 0;
  c->prof.on = c->profiling;
  c->prof.n = 0;
  for (int j = 0; j < sppm_ctx::MAX_LANES; j++) {
    c->lanes[j].prof.on = c->profiling;
    c->lanes[j].prof.n = 0;
  }
  return SPPM_OK;
}

int sppm_reset_counters(sppm_ctx* c) {
  if (!c) return SPPM_ERR_INVALID;
  Bind b(c);
  CU(c, join_comm(c));
  CU(c, cudaStreamSynchronize(c->stream));
  fold_timings(c);
  CU(c, cudaMemset(c->counters.p, 0, sizeof(DevCounters)));
  const uint64_t launches = c->host_counters.launches;
  c->host_counters = SppmCounters{};
  c->host_counters.launches = launches;
  return SPPM_OK;
}

}  // extern "C"
