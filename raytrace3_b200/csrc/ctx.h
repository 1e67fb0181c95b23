// ctx.h — host-side context of libsppm_b200.so and the launch interfaces between its
// translation units (abi.cu = C ABI + orchestration, trace.cu = eye/photon/intersection
// kernels, gather.cu = hash-grid build, deposit sort, gather, update, film).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/sppm_b200.h"
#include "scene.cuh"

namespace sppm {

// Raw (arrival-order) hit point written by the eye pass; 96 B.
struct HpRaw {
  double pos[3], n[3], wgt[3];
  double r2;
  int32_t row, col;
  uint64_t okey;  // canonical order key: pixel index << 20 | eye-path node
};
// Deposit record (Raytracer.h:139-156 inputs): x, N, flux; 80 B, 16-byte aligned for
// 128-bit broadcast loads in the gather.  meta = photon id | bounce << 56.
struct __align__(16) Deposit {
  double x[3], n[3], flux[3];
  uint64_t meta;
};

// Hash grid over the hit points (replaces the flann kd-tree, Raytracer.h:370-380).
struct Grid {
  double org[3];
  double inv_cell;
  int dim[3];
  int64_t n_cells;
  // A deposit outside [keep_lo, keep_hi] in some coordinate is farther than sqrt(searchR2) from
  // every hit point and is not kept: keep_lo/hi = the hit points' bounding box (true
  // coordinates) grown by one cell edge.  Without this, photons that escape to the far reaches
  // of the scene's infinite planes are clamped into the grid's boundary cells next to the
  // visible walls and streamed through the gather for nothing (43 % of the deposits of the
  // reference scene).
  double keep_lo[3], keep_hi[3];
};
__host__ __device__ inline bool grid_keeps(const Grid& g, double x, double y, double z) {
  return x >= g.keep_lo[0] && x <= g.keep_hi[0] && y >= g.keep_lo[1] && y <= g.keep_hi[1] && z >= g.keep_lo[2] &&
         z <= g.keep_hi[2];
}
__host__ __device__ inline int64_t grid_cell(const Grid& g, double x, double y, double z) {
#if defined(__CUDA_ARCH__) || defined(SPPM_HOST_EMULATION)  // the latter: tests/device_host only
  int ix = __double2int_rd((x - g.org[0]) * g.inv_cell);
  int iy = __double2int_rd((y - g.org[1]) * g.inv_cell);
  int iz = __double2int_rd((z - g.org[2]) * g.inv_cell);
#else
  int ix = 0, iy = 0, iz = 0;
  (void)x; (void)y; (void)z;
#endif
  ix = ix < 0 ? 0 : (ix >= g.dim[0] ? g.dim[0] - 1 : ix);
  iy = iy < 0 ? 0 : (iy >= g.dim[1] ? g.dim[1] - 1 : iy);
  iz = iz < 0 ? 0 : (iz >= g.dim[2] ? g.dim[2] - 1 : iz);
  return ((int64_t)iz * g.dim[1] + iy) * g.dim[0] + ix;
}

// Sorted structure-of-arrays hit-point map (device pointers, length n).
struct HpSoA {
  double *px, *py, *pz, *nx, *ny, *nz, *r2, *wx, *wy, *wz;
  double* tao;    // [3][n]
  double* dtao;   // [3][n] per-round accumulator
  int32_t *N, *newN, *dN;
  int32_t *row, *col;
  uint64_t* okey;
};

// Work unit of the filtered gather: up to 64 hit points that are consecutive in the cell-sorted
// map and lie in one run of x-adjacent occupied cells.  The deposits that can reach them are
// 9 contiguous ranges of the cell-sorted deposit list (rows dy,dz in -1..1, cells x_first-1 ..
// x_last+1), given as compact-slot runs.
struct GUnit {
  int32_t hp_begin, hp_count;
  int32_t n_cells_x, pad;
  double org[3];   // centre of the unit's cells: origin of the FP32 coordinates
  int2 run[9];     // [slot_lo, slot_hi_exclusive) per row, (0,0) outside the grid
};

struct DevCounters {
  unsigned long long photons, isect_queries, patch_solves, deposits, deposits_kept, cand_pairs, acc_pairs, overflow, patch_culled;
};

// sppm_set_profiling: one event before every launch of a batch; the span up to the next event goes to a bucket
enum ProfKind { PK_STEP, PK_NEWTON, PK_SPILL, PK_BIN_SCAN, PK_BIN_SCATTER, PK_GATHER_PREP, PK_GATHER_KERNEL, PK_UPDATE, PK_END, PK_N };
struct ProfMarks {
  std::vector<cudaEvent_t> ev;
  std::vector<int> kind;
  int n = 0;
  bool on = false;
  cudaError_t mark(int k, cudaStream_t st) {
    if (!on) return cudaSuccess;
    if (n == (int)ev.size()) {
      cudaEvent_t e;
      cudaError_t r = cudaEventCreate(&e);
      if (r) return r;
      ev.push_back(e);
      kind.push_back(0);
    }
    kind[n] = k;
    return cudaEventRecord(ev[n++], st);
  }
  void fold(double* bucket /*[PK_N]*/) {  // after the stream was synchronised
    for (int i = 0; i + 1 < n; i++) {
      float ms = 0;
      if (kind[i] != PK_END && cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) bucket[kind[i]] += ms;
    }
    n = 0;
  }
  void destroy() {
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    ev.clear();
    kind.clear();
    n = 0;
  }
};

struct DevBuf {  // growable device allocation
  void* p = nullptr;
  size_t bytes = 0;
};

// A photon "lane": its own stream and buffers for tracing one batch of photons and binning its
// deposits.  Batches rotate over the lanes (sppm_ctx::n_lanes) so that the traces of the next
// batches overlap the gather (main stream) of the current one.
struct Lane {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_trace_done = nullptr, ev_gather_done = nullptr;
  DevBuf dep_raw, dep_sorted, dep_key, dep_rank;  // deposits: arrival order / cell order, slot key, rank in slot
  DevBuf slot_count, dep_start;                   // u32[n_slots+2] counts and their exclusive scan (CSR offsets)
  DevBuf wf_store;                                // wavefront photon queues + item lists
  DevBuf scalars;                                 // u64[8]: [1] deposit count, [2] photon cursor
  DevBuf cub_tmp;
  int64_t dep_cap = 0, wf_cap = 0;
  ProfMarks prof;
};

}  // namespace sppm

struct sppm_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::string err;
  SppmParams prm;
  // scene
  sppm::DevScene scene{};
  std::vector<sppm::DevBuf> scene_bufs;
  bool scene_ok = false;
  bool patches_in_smem = true;
  size_t trace_smem = 0;
  // launch configurations, queried once per scene (trace_configure resets them) instead of once per batch
  struct WfCfg { bool valid = false; int step_blocks = 0, newton_blocks = 0; size_t step_smem = 0, newton_smem = 0; } wf_cfg;
  int trace_per_sm = 0;        // photon_trace_kernel CTAs per SM (0 = not queried yet)
  int gather_per_sm[3] = {};   // gather_kernel, gather_filtered_kernel<false>, <true>
  // hit points
  sppm::DevBuf hp_raw;     // HpRaw[cap]
  sppm::DevBuf hp_raw_alt; // receive buffer of sppm_allgather_hitpoints (swapped with hp_raw)
  int64_t hp_cap = 0, n_hp = 0;
  int film_w = 0, film_h = 0;
  sppm::DevBuf soa_store;  // backing store of HpSoA
  sppm::HpSoA soa{};
  sppm::DevBuf hp_canon;   // u32[n_hp]: position in the sorted map by canonical rank (pixel, eye-path node)
  bool index_ok = false;
  // grid
  sppm::Grid grid{};
  sppm::DevBuf cell_slot;     // int32[n_cells]: compact id of cells near hit points, -1 otherwise
  int32_t n_slots = 0;
  sppm::DevBuf occ_cell, occ_start;  // occupied cells: grid id [n_occ], hp start [n_occ+1]
  int32_t n_occ = 0;
  sppm::DevBuf occ_nbr;       // int2[n_occ][9]: the neighbourhood as 9 x-runs of compact slots
  sppm::DevBuf units;         // GUnit[n_units]: work units of the filtered gather
  int32_t n_units = 0;
  sppm::DevBuf item_parts, item_off;  // per-batch work-item counts of the units and their exclusive scan
  // deposits
  static constexpr int MAX_LANES = 8;
  sppm::Lane lanes[MAX_LANES];
  int n_lanes = 4;            // lanes used round-robin (SPPM_LANES, 1..MAX_LANES); 4 measured best (2: 2.80, 3: 2.72, 4: 2.68 ms; 5-8: no gain)
  int n_lanes_small = 6;      // for batches of at most SMALL_BATCH photons (SPPM_LANES_SMALL): their launches are latency floors (100 k-photon rounds: 4: 354, 6: 381, 8: 375 M photons/s)
  static constexpr uint64_t SMALL_BATCH = 1u << 19;
  int lane_rr = 0, last_lane = 0;
  cudaEvent_t ev_setup = nullptr;  // recorded on the main stream after scene / hit-point map changes
  sppm::DevBuf dep_key, dep_key2, dep_idx, dep_idx2, scan_scratch, cub_tmp;  // build_index scratch
  int64_t n_dep = 0;
  sppm::DevBuf dev_scalars;   // [0] hp count, [1] deposit count, [2] photon cursor ...
  sppm::DevBuf counters;      // DevCounters
  sppm::DevBuf film;          // double[h][w][3]
  sppm::DevBuf scratch;
  // shared statistics across eye passes: per pixel R2, N, tau[3], per-pass M, Phi[3]
  sppm::DevBuf px_store;
  int px_w = 0, px_h = 0;
  uint64_t px_photons = 0;   // photons emitted into the shared statistics so far
  bool newton_cull = true;    // 13-DOP line test before Newton solves (off only inside sppm_bezier_patch_bench)
  bool keep_all_deposits = false;
  SppmCounters host_counters{};
  bool profiling = false;     // sppm_set_profiling
  sppm::ProfMarks prof;       // marks on the main stream (gather)
  cudaEvent_t ev[8] = {};  // 4/5 generic, 6/7 user stopwatch
  static constexpr int TIMING_RING = 32;
  cudaEvent_t tev[TIMING_RING][4] = {};  // per batch: trace start, bin start, gather start, gather end
  int t_pending = 0;          // batches whose timings are not folded into host_counters yet
  cudaEvent_t aev[TIMING_RING][2] = {};  // sppm_allreduce: start / end on the main stream
  int a_pending = 0;
  bool t_has_gather[TIMING_RING] = {};
  // NCCL (dlopen)
  void* nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  // Optional (SPPM_ASYNC_COMM=1): exchange overlapped with the next round's gather (REFERENCE update only: HitPoint::update then never changes a radius,
  // Raytracer.h:70, so the next round's gather does not depend on it).  The per-round accumulators dN / dtao
  // exist twice: while the allreduce + update of round i run on comm_stream over one set, the gather of round
  // i+1 fills the other on the main stream.
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_gathers_done = nullptr, ev_acc_free[2] = {nullptr, nullptr}, ev_comm = nullptr;
  int32_t* dN_set[2] = {nullptr, nullptr};
  double* dtao_set[2] = {nullptr, nullptr};
  int acc_cur = 0;
  bool async_comm = false;    // SPPM_ASYNC_COMM=1 moves the exchange to comm_stream (measured: no gain)
  bool comm_pending = false;  // an allreduce was issued on comm_stream and its update has not been yet
  bool comm_busy = false;     // comm_stream holds work the main stream has not been ordered after
};

namespace sppm {
// trace.cu
cudaError_t launch_eye_pass(sppm_ctx* c, const SppmCameraDesc& cam, uint64_t pass_id, int row_begin, int row_end);
cudaError_t launch_photon_trace(sppm_ctx* c, Lane& L, uint64_t begin, uint64_t count, uint64_t pass_id,
                                const unsigned* r_count = nullptr, double* const* r_st = nullptr,
                                const unsigned long long* r_idd = nullptr);
cudaError_t launch_intersect(sppm_ctx* c, const double* d_rays6, int64_t n, const uint64_t* d_ids, uint64_t pass_id,
                             int node, SppmHit* d_out);
cudaError_t launch_patch_batch(sppm_ctx* c, const double* d_rays6, const int32_t* d_patch, int64_t n,
                               const double* d_seeds, double* d_tuv);
cudaError_t trace_configure(sppm_ctx* c);
// wavefront.cu
cudaError_t launch_photon_trace_wavefront(sppm_ctx* c, Lane& L, uint64_t begin, uint64_t count, uint64_t pass_id);
// gather.cu
cudaError_t build_index(sppm_ctx* c);
cudaError_t key_host_deposits(sppm_ctx* c, Lane& L, int64_t m, cudaStream_t st);
cudaError_t bin_deposits(sppm_ctx* c, Lane& L, cudaStream_t st);
cudaError_t launch_gather(sppm_ctx* c, Lane& L, cudaEvent_t ev_gather_start, cudaEvent_t ev_gather_end);
cudaError_t launch_update(sppm_ctx* c, cudaStream_t st);
cudaError_t launch_resolve(sppm_ctx* c, int cnt, int onetime);
cudaError_t shared_begin(sppm_ctx* c, int w, int h);
cudaError_t shared_load(sppm_ctx* c);
cudaError_t shared_fold(sppm_ctx* c);
cudaError_t shared_resolve(sppm_ctx* c);
cudaError_t shared_export(sppm_ctx* c, double* r2, double* n, double* tau3);
cudaError_t shared_import(sppm_ctx* c, const double* r2, const double* n, const double* tau3);
cudaError_t dev_reserve(DevBuf& b, size_t bytes);
void dev_free(DevBuf& b);
}  // namespace sppm
