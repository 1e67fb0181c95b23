// gather_kernels.cuh — the kernels of gather.cu: hash-grid build, deposit binning, the two gather
// kernels (exact and FP32-filtered), HitPoint::update, the film resolve and the shared per-pixel
// statistics; the host orchestration (CUB sorts / scans, buffers, launches) stays in gather.cu.  A
// header of their own so that tests/device_host can compile these kernel bodies for the CPU and hold
// the filtered gather to the exhaustive search without a GPU; gather.cu is the only product file that
// includes it.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ctx.h"

namespace sppm {

namespace {

constexpr int TPB = 256;
inline unsigned nblk(long long n, int tpb = TPB) { return (unsigned)((n + tpb - 1) / tpb); }
#define CK(x)                      \
  do {                             \
    cudaError_t e__ = (x);         \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

__global__ void bounds_kernel(const HpRaw* hp, long long n, unsigned long long* mm /*[6]*/, double* sum /*[3]*/) {
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, s[3] = {0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    for (int k = 0; k < 3; k++) {
      const double v = hp[i].pos[k];
      lo[k] = fmin(lo[k], v);
      hi[k] = fmax(hi[k], v);
      s[k] += v;
    }
  for (int k = 0; k < 3; k++) {
    for (int off = 16; off > 0; off >>= 1) {
      lo[k] = fmin(lo[k], __shfl_xor_sync(FULL, lo[k], off));
      hi[k] = fmax(hi[k], __shfl_xor_sync(FULL, hi[k], off));
      s[k] += __shfl_xor_sync(FULL, s[k], off);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(mm + k, dbl_key(lo[k]));
      atomicMax(mm + 3 + k, dbl_key(hi[k]));
      atomicAdd(sum + k, s[k]);
    }
  }
}

__global__ void hp_okey_kernel(const HpRaw* hp, long long n, uint64_t* okey, uint32_t* idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  okey[i] = hp[i].okey;
  idx[i] = (uint32_t)i;
}
// Sort key of the hit points taken in canonical (okey) order: grid cell in the high bits, below it a 15-bit
// Morton code of the position inside the cell (5 bits per axis).  Hit points that are consecutive in the sorted
// map are then neighbours in space, so a work unit of the gather (<= 64 consecutive hit points) is a compact
// cluster with a small bounding box instead of a strip of film rows — fewer deposits survive its cull.  Ties
// keep the canonical order (stable sort), so every rank of a multi-GPU run still holds the same map.
constexpr int HP_MORTON_BITS = 15;
__device__ __forceinline__ uint32_t spread5(uint32_t v) {  // abcde -> a00b00c00d00e
  v &= 31u;
  v = (v | (v << 8)) & 0x100fu;
  v = (v | (v << 4)) & 0x10c3u;
  v = (v | (v << 2)) & 0x1249u;
  return v;
}
__global__ void hp_key_kernel(const HpRaw* hp, const uint32_t* idx, long long n, Grid g, uint64_t* key, int morton) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const HpRaw& r = hp[idx[i]];
  const int64_t c = grid_cell(g, r.pos[0], r.pos[1], r.pos[2]);
  const int ix = (int)(c % g.dim[0]), iy = (int)((c / g.dim[0]) % g.dim[1]), iz = (int)(c / ((int64_t)g.dim[0] * g.dim[1]));
  uint32_t q[3];
  const int ic[3] = {ix, iy, iz};
#pragma unroll
  for (int k = 0; k < 3; k++) {  // a hit point clamped into a boundary cell lies outside it: clamp the code too
    const double f = ((r.pos[k] - g.org[k]) * g.inv_cell - ic[k]) * 32.0;
    q[k] = f > 0 ? (f < 31 ? (uint32_t)f : 31u) : 0u;
  }
  key[i] = ((uint64_t)c << HP_MORTON_BITS) | (morton ? spread5(q[0]) | (spread5(q[1]) << 1) | (spread5(q[2]) << 2) : 0u);
}
// the grid cell of every sorted key (run-length encoded into the occupied-cell list)
__global__ void hp_cell_of_key_kernel(const uint64_t* key, long long n, uint32_t* cell) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cell[i] = (uint32_t)(key[i] >> HP_MORTON_BITS);
}

__global__ void hp_permute_kernel(const HpRaw* hp, const uint32_t* idx, long long n, HpSoA s) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const HpRaw r = hp[idx[i]];
  s.px[i] = r.pos[0]; s.py[i] = r.pos[1]; s.pz[i] = r.pos[2];
  s.nx[i] = r.n[0]; s.ny[i] = r.n[1]; s.nz[i] = r.n[2];
  s.wx[i] = r.wgt[0]; s.wy[i] = r.wgt[1]; s.wz[i] = r.wgt[2];
  s.r2[i] = r.r2;
  s.row[i] = r.row; s.col[i] = r.col;
  s.okey[i] = r.okey;
  s.N[i] = 0; s.newN[i] = 0; s.dN[i] = 0;
  for (int k = 0; k < 3; k++) {
    s.tao[(size_t)k * n + i] = 0;
    s.dtao[(size_t)k * n + i] = 0;
  }
}

__global__ void mark_nbr_kernel(const uint32_t* occ_cell, int n_occ, Grid g, int32_t* flag) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_occ * 27) return;
  const int o = (int)(t / 27), k = (int)(t % 27);
  const uint32_t c = occ_cell[o];
  const int ix = c % g.dim[0], iy = (c / g.dim[0]) % g.dim[1], iz = c / (g.dim[0] * g.dim[1]);
  const int jx = ix + k % 3 - 1, jy = iy + (k / 3) % 3 - 1, jz = iz + k / 9 - 1;
  if (jx < 0 || jy < 0 || jz < 0 || jx >= g.dim[0] || jy >= g.dim[1] || jz >= g.dim[2]) return;
  flag[((long long)jz * g.dim[1] + jy) * g.dim[0] + jx] = 1;
}
__global__ void finalize_slot_kernel(int32_t* flag_to_slot, const int32_t* scan, long long n_cells, int32_t* n_slots) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cells) return;
  const int f = flag_to_slot[i];
  if (i == n_cells - 1) *n_slots = scan[i] + f;
  flag_to_slot[i] = f ? scan[i] : -1;
}
// The 27-neighbourhood of an occupied cell as 9 runs along x.  Compact slots are assigned in
// linear cell order and every neighbour of an occupied cell has a slot, so the (up to) three
// x-adjacent cells of a row have consecutive slots: run = [slot_lo, slot_hi_exclusive), or
// (0,0) for a row outside the grid.  With cell-sorted deposits a run is ONE contiguous range.
__global__ void occ_run_kernel(const uint32_t* occ_cell, int n_occ, Grid g, const int32_t* cell_slot, int2* occ_run) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_occ * 9) return;
  const int o = (int)(t / 9), k = (int)(t % 9);
  const uint32_t c = occ_cell[o];
  const int ix = c % g.dim[0], iy = (c / g.dim[0]) % g.dim[1], iz = c / (g.dim[0] * g.dim[1]);
  const int jy = iy + k % 3 - 1, jz = iz + k / 3 - 1;
  int2 r = make_int2(0, 0);
  if (!(jy < 0 || jz < 0 || jy >= g.dim[1] || jz >= g.dim[2])) {
    const int x0 = max(ix - 1, 0), x1 = min(ix + 1, g.dim[0] - 1);
    const long long row = ((long long)jz * g.dim[1] + jy) * g.dim[0];
    r.x = cell_slot[row + x0];
    r.y = cell_slot[row + x1] + 1;
  }
  occ_run[t] = r;
}
__global__ void set_last_kernel(int32_t* starts, int n_occ, int32_t n_hp) { starts[n_occ] = n_hp; }

// ---- work units of the filtered gather ----------------------------------------------------
#ifndef SPPM_UNIT_HP
#define SPPM_UNIT_HP 64
#endif
constexpr int UNIT_HP = SPPM_UNIT_HP;  // hit points per unit: 64 (two per lane) or 32
// flag[o] = 1 when occupied cell o starts a run of x-adjacent occupied cells
__global__ void run_flag_kernel(const uint32_t* occ_cell, int n_occ, int dim0, int32_t* flag) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_occ) return;
  const uint32_t c = occ_cell[o];
  flag[o] = (o == 0 || occ_cell[o - 1] + 1 != c || c % (uint32_t)dim0 == 0) ? 1 : 0;
}
// units per run (a run's hit points are contiguous: cut them every UNIT_HP)
__global__ void unit_count_kernel(const int32_t* run_first, const int32_t* n_runs_p, int n_occ, const int32_t* occ_start,
                                  int32_t* cnt) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_runs = *n_runs_p;
  if (r >= n_runs) return;
  const int o0 = run_first[r], o1 = r + 1 < n_runs ? run_first[r + 1] : n_occ;
  cnt[r] = (occ_start[o1] - occ_start[o0] + UNIT_HP - 1) / UNIT_HP;
}
// One thread per unit: unit u belongs to run r = the last run with unit_off[r] <= u and holds hit points
// [b, e) of it, b = (first hit point of the run) + (u - unit_off[r]) * UNIT_HP; its first / last cells are found
// by bisection in occ_start.  (One thread per RUN walking its units serially took 120 us at 786 k hit points:
// the long runs of the dense walls were the whole launch.)
__global__ void unit_fill_kernel(const int32_t* run_first, const int32_t* n_runs_p, int n_occ, const int32_t* occ_start,
                                 const uint32_t* occ_cell, const int32_t* unit_off, int n_units, Grid g,
                                 const int32_t* cell_slot, GUnit* units) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_units) return;
  const int n_runs = *n_runs_p;
  int lo = 0, hi = n_runs;  // last r in [0, n_runs) with unit_off[r] <= u (runs without units cannot own u: strict progress)
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (unit_off[mid] <= u) lo = mid; else hi = mid;
  }
  const int r = lo;
  const int o0 = run_first[r], o1 = r + 1 < n_runs ? run_first[r + 1] : n_occ;
  const int p1 = occ_start[o1];
  const int b = occ_start[o0] + (u - unit_off[r]) * UNIT_HP;
  const int e = min(b + UNIT_HP, p1);
  auto cell_of = [&](int p) {  // the cell o in [o0, o1) with occ_start[o] <= p < occ_start[o + 1]
    int a = o0, z = o1;
    while (z - a > 1) {
      const int mid = (a + z) >> 1;
      if (occ_start[mid] <= p) a = mid; else z = mid;
    }
    return a;
  };
  const int o = cell_of(b), ol = cell_of(e - 1);
  const double cell = 1.0 / g.inv_cell;
  const uint32_t c = occ_cell[o];
  const int ix = c % g.dim[0], iy = (c / g.dim[0]) % g.dim[1], iz = c / (g.dim[0] * g.dim[1]);
  const int nx = ol - o + 1;
  GUnit un;
  un.hp_begin = b;
  un.hp_count = e - b;
  un.n_cells_x = nx;
  un.pad = 0;
  un.org[0] = g.org[0] + (ix + 0.5 * nx) * cell;
  un.org[1] = g.org[1] + (iy + 0.5) * cell;
  un.org[2] = g.org[2] + (iz + 0.5) * cell;
  for (int k = 0; k < 9; k++) {
    const int jy = iy + k % 3 - 1, jz = iz + k / 3 - 1;
    int2 rr = make_int2(0, 0);
    if (!(jy < 0 || jz < 0 || jy >= g.dim[1] || jz >= g.dim[2])) {
      const int x0 = max(ix - 1, 0), x1 = min(ix + nx, g.dim[0] - 1);
      const long long row = ((long long)jz * g.dim[1] + jy) * g.dim[0];
      rr.x = cell_slot[row + x0];
      rr.y = cell_slot[row + x1] + 1;
    }
    un.run[k] = rr;
  }
  units[u] = un;
}

// ---- deposits --------------------------------------------------------------------------
__global__ void dep_key_kernel(const Deposit* dep, long long m, Grid g, const int32_t* cell_slot, int32_t n_slots,
                               uint32_t* key, uint32_t* rank, unsigned int* slot_count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int32_t s = grid_keeps(g, dep[i].x[0], dep[i].x[1], dep[i].x[2]) ? cell_slot[grid_cell(g, dep[i].x[0], dep[i].x[1], dep[i].x[2])] : -1;
  const uint32_t k = s >= 0 ? (uint32_t)s : (uint32_t)n_slots;
  key[i] = k;
  rank[i] = atomicAdd(slot_count + k, 1u);
}
// Counting sort, last step: off[] = exclusive scan of the per-slot counts (the CSR offsets the
// gather reads; off[n_slots] = number of kept deposits), every deposit knows its slot and its
// arrival rank within the slot, so sorted[off[key] + rank] = raw.  The deposit count stays on
// the device (no host round trip between tracing and gathering).
__global__ void __launch_bounds__(320) dep_scatter_kernel(const Deposit* raw, const uint32_t* key, const uint32_t* rank,
                                                          const unsigned long long* m_ptr, unsigned long long cap,
                                                          const int32_t* off, Deposit* sorted) {
  // five threads move one 80-byte record (16 B each), so every record is read and written as
  // one contiguous 80-byte piece of a warp-wide request; blockDim = 320 = 64 records per CTA pass
  const unsigned long long m = min(*m_ptr, cap);
  const unsigned part = threadIdx.x % 5;
  for (unsigned long long i = (unsigned long long)blockIdx.x * 64 + threadIdx.x / 5; i < m; i += (unsigned long long)gridDim.x * 64) {
    const double2* src = reinterpret_cast<const double2*>(raw + i);
    double2* dst = reinterpret_cast<double2*>(sorted + (off[key[i]] + (int32_t)rank[i]));
    dst[part] = src[part];
  }
}

// Per batch: a unit whose 9 deposit ranges hold D deposits becomes ceil(D / ITEM_DEP) work items
// (none when D = 0), so a boundary cell that collected 10^5 clamped deposits is spread over many
// warps.  parts[u] -> exclusive scan off[]; work item `it` is part it - off[u] of the unit u with
// off[u] <= it < off[u+1], which the warp finds with a 32-ary search (each step the 32 lanes probe
// 32 evenly spaced entries of off[] and a ballot narrows the range 32-fold): no item list has to
// be materialised, so no capacity can be exceeded however many units a deposit can reach.
constexpr int ITEM_DEP = 1024;               // deposits per work item
__global__ void unit_parts_kernel(const GUnit* units, int n_units, const int32_t* dep_off, int32_t* parts) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u > n_units) return;
  int d = 0;
  if (u < n_units)
    for (int k = 0; k < 9; k++) {
      const int2 r = units[u].run[k];
      if (r.y > r.x) d += dep_off[r.y] - dep_off[r.x];
    }
  parts[u] = (d + ITEM_DEP - 1) / ITEM_DEP;  // parts[n_units] = 0: the scan's last element is the item count
}
struct GatherArgs {
  HpSoA s;
  long long n_hp;
  const int32_t* occ_start;
  const int2* occ_run;  // [n_occ][9] slot runs
  int n_occ;
  const int32_t* dep_off;  // CSR offsets [n_slots + 1]
  const Deposit* dep;
  double search_r2;
  DevCounters* counters;
  Grid grid;                 // filtered kernel: cell centres
  const uint32_t* occ_cell;  // grid id of occupied cell c
  const GUnit* units;        // filtered kernel: work units
  int n_units;
  const int32_t* item_off;   // [n_units + 1] exclusive scan of the units' work-item counts (this batch)
  int full_accept;           // filtered kernel: whole-unit accept on (see gather_filtered_kernel)
};

// ---- TMA (cp.async.bulk) + mbarrier helpers -----------------------------------------------
#ifdef SPPM_HOST_EMULATION  // tests/device_host only (g++ cannot assemble PTX): the copy completes at issue
__device__ __forceinline__ void mbar_init(uint64_t*, unsigned) {}
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t*, unsigned) {}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void mbar_wait(uint64_t*, unsigned) {}
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spin = 0; !done; spin++) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();  // a lost transaction must abort, not hang the GPU
  }
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#endif

// One CTA per occupied hit-point cell (cells are handed out by a global work counter).  The
// deposits of the cell's 27 neighbour cells are 9 contiguous ranges of the cell-sorted list;
// one thread streams them into a double-buffered shared-memory tile with TMA bulk copies
// (cp.async.bulk + mbarrier) one tile ahead of the math.  The 256 threads are split hl x dl:
// hl threads hold hit points of the cell (pos, n, R2 in registers), dl = 256/hl threads stride
// over the tile, so a warp reads each 80-byte record as 128-bit shared-memory broadcasts.
// Partial sums are combined with shuffles + shared memory and every hit point's accumulators
// are written exactly once per batch (no atomics).
constexpr int G_THREADS = 256;
#ifndef SPPM_G_TILE
#define SPPM_G_TILE 256
#endif
constexpr int G_TILE = SPPM_G_TILE;  // deposits per tile: 20 KB at 256, two tiles in flight

struct RunCursor {  // walks the concatenation of the 9 deposit ranges tile by tile (uniform)
  int r, off;
};
// next piece of at most `room` deposits; returns its length and start (0 when exhausted)
__device__ __forceinline__ int next_piece(RunCursor& c, const int* rs, const int* rl, int room, int& start) {
  while (c.r < 9 && c.off >= rl[c.r]) {
    c.r++;
    c.off = 0;
  }
  if (c.r >= 9) return 0;
  const int take = min(rl[c.r] - c.off, room);
  start = rs[c.r] + c.off;
  c.off += take;
  return take;
}

__global__ void __launch_bounds__(G_THREADS) gather_kernel(const GatherArgs a, unsigned int* work_counter) {
  extern __shared__ __align__(128) unsigned char gsm[];
  double2(*tile)[G_TILE * 5] = reinterpret_cast<double2(*)[G_TILE * 5]>(gsm);  // tile[2][G_TILE * 5]
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ int rs[9], rl[9];
  __shared__ int s_cell, s_total;
  __shared__ int red_cnt[G_THREADS];
  __shared__ double red_s[3][G_THREADS];
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  unsigned phase[2] = {0, 0};  // parity of the next completion of each barrier (uniform)
  unsigned cand32 = 0, acc32 = 0;
  unsigned long long cand = 0, acc = 0;
  for (;;) {
    __syncthreads();  // previous cell fully done (rs/rl/s_cell/red reusable)
    if (tid == 0) s_cell = (int)atomicAdd(work_counter, 1u);
    __syncthreads();
    const int c = s_cell;
    if (c >= a.n_occ) break;
    if (tid < 32) {
      int ln = 0;
      if (tid < 9) {
        const int2 run = a.occ_run[(long long)c * 9 + tid];
        int st = 0;
        if (run.y > run.x) {
          st = a.dep_off[run.x];
          ln = a.dep_off[run.y] - st;
        }
        rs[tid] = st;
        rl[tid] = ln;
      }
      int tot = ln;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(FULL, tot, off);
      if (tid == 0) s_total = tot;
    }
    __syncthreads();
    const int D = s_total;  // deposits in the 27-cell neighbourhood = concatenation of the 9 runs
    if (D == 0) continue;   // no deposit near this cell in this batch
    const int n_tiles = (D + G_TILE - 1) / G_TILE;
    const int hs = a.occ_start[c], he = a.occ_start[c + 1];
    for (int h0 = hs; h0 < he; h0 += G_THREADS) {
      const int H = min(G_THREADS, he - h0);
      // thread -> (hit point, deposit slice).  Cells with few hit points use a power-of-two
      // split (partials combined with shuffles); otherwise hl = H exactly, dl = 256 / H slices
      // and at most H - 1 idle threads (90 % of the pair tests sit in cells with 17..128 points).
      const bool pow2 = H < 9;
      int hl = H;
      if (pow2) {
        hl = 1;
        while (hl < H) hl <<= 1;
      }
      const int dl = G_THREADS / hl;
      const int my_h = pow2 ? (tid & (hl - 1)) : (tid % hl), my_d = tid / hl;
      const int i = h0 + my_h;
      const bool valid = my_h < H && my_d < dl;
      double px = 0, py = 0, pz = 0, nx = 0, ny = 0, nz = 0, r2 = -1;
      if (valid) {
        px = a.s.px[i]; py = a.s.py[i]; pz = a.s.pz[i];
        nx = a.s.nx[i]; ny = a.s.ny[i]; nz = a.s.nz[i];
        r2 = a.s.r2[i];
      }
      int cnt = 0;
      double s0 = 0, s1 = 0, s2 = 0;
      // Tile k holds deposits [k*G_TILE, min(D, (k+1)*G_TILE)) of the concatenation.  Only the
      // producer (thread 0) walks the runs; it issues the bulk copies of tile k+1 while every
      // thread consumes tile k.
      RunCursor pc{0, 0};
      auto issue = [&](int k) {
        const int buf = k & 1;
        const int fill_k = min(G_TILE, D - k * G_TILE);
        mbar_expect_tx(&bar[buf], (unsigned)fill_k * (unsigned)sizeof(Deposit));
        int fill = 0, start = 0, take;
        while (fill < fill_k && (take = next_piece(pc, rs, rl, fill_k - fill, start)) > 0) {
          bulk_g2s(&tile[buf][fill * 5], a.dep + start, (unsigned)take * (unsigned)sizeof(Deposit), &bar[buf]);
          fill += take;
        }
      };
      if (tid == 0) issue(0);
      for (int k = 0; k < n_tiles; k++) {
        const int buf = k & 1;
        if (tid == 0 && k + 1 < n_tiles) issue(k + 1);  // prefetch: buffer (k+1)&1 was released by the barrier below
        const int fill_cur = min(G_TILE, D - k * G_TILE);
        mbar_wait(&bar[buf], phase[buf]);
        phase[buf] ^= 1;
        if (valid) {
          const double2* tb = tile[buf];
#pragma unroll 2
          for (int d = my_d; d < fill_cur; d += dl) {
            const double2* q = tb + d * 5;
            const double2 q0 = q[0], q1 = q[1], q2 = q[2];
            // x = (q0.x, q0.y, q1.x)  N = (q1.y, q2.x, q2.y)  flux = (q3.x, q3.y, q4.x)
            const double vx = px - q0.x, vy = py - q0.y, vz = pz - q1.x;  // Raytracer.h:152
            const double vv = (vx * vx + vy * vy) + vz * vz;
            if (vv < a.search_r2) {  // flann radiusSearch, strict '<' (Q3)
              cand32++;
              const double nn = (nx * q1.y + ny * q2.x) + nz * q2.y;
              if (nn > SPPM_NORMAL_AGREE && vv <= r2) {  // Raytracer.h:154
                const double2 q3 = q[3], q4 = q[4];
                cnt++;
                s0 += q3.x; s1 += q3.y; s2 += q4.x;
              }
            }
          }
        }
        __syncthreads();  // tile k fully consumed: its buffer may be refilled (tile k+2)
      }
      acc32 += (unsigned)cnt;
      // combine the dl partial sums of every hit point
      if (pow2) {
        for (int o = hl; o < 32; o <<= 1) {
          cnt += __shfl_xor_sync(FULL, cnt, o);
          s0 += __shfl_xor_sync(FULL, s0, o);
          s1 += __shfl_xor_sync(FULL, s1, o);
          s2 += __shfl_xor_sync(FULL, s2, o);
        }
      }
      red_cnt[tid] = valid ? cnt : 0;
      red_s[0][tid] = s0; red_s[1][tid] = s1; red_s[2][tid] = s2;
      __syncthreads();
      if (tid < H) {
        int tc = 0;
        double t0 = 0, t1 = 0, t2 = 0;
        if (pow2) {  // lanes 0..hl-1 of every warp hold that warp's sums
          for (int w = 0; w < G_THREADS / 32; w++) {
            tc += red_cnt[w * 32 + tid];
            t0 += red_s[0][w * 32 + tid]; t1 += red_s[1][w * 32 + tid]; t2 += red_s[2][w * 32 + tid];
          }
        } else {
          for (int d = 0; d < dl; d++) {
            tc += red_cnt[d * hl + tid];
            t0 += red_s[0][d * hl + tid]; t1 += red_s[1][d * hl + tid]; t2 += red_s[2][d * hl + tid];
          }
        }
        if (tc > 0) {  // hp->newN++, hp->tao += wgt.mul(flux)*PI_INV (Raytracer.h:155-156)
          a.s.dN[i] += tc;
          a.s.dtao[i] += (a.s.wx[i] * t0) * SPPM_PI_INV;
          a.s.dtao[a.n_hp + i] += (a.s.wy[i] * t1) * SPPM_PI_INV;
          a.s.dtao[2 * a.n_hp + i] += (a.s.wz[i] * t2) * SPPM_PI_INV;
        }
      }
      __syncthreads();
    }
    cand += cand32;
    acc += acc32;
    cand32 = acc32 = 0;
  }
  for (int off = 16; off > 0; off >>= 1) {
    cand += __shfl_xor_sync(FULL, cand, off);
    acc += __shfl_xor_sync(FULL, acc, off);
  }
  if (lane == 0 && cand) {
    atomicAdd(&a.counters->cand_pairs, cand);
    atomicAdd(&a.counters->acc_pairs, acc);
  }
}

// ---- filtered gather: FP32 interval classification, FP64 only where it matters -----------
//
// The acceptance predicate of Raytracer.h:152-154 is three FP64 comparisons per (hit point,
// deposit) pair, and ~80 % of the pairs of a 27-cell neighbourhood fail the first one.  The
// filtered kernel evaluates the predicate in packed FP32 (sm_100 FADD2/FMUL2/FFMA2, two
// deposits per instruction) on coordinates taken relative to an origin o close to both points
// (the centre of the work unit's cells), with a rigorous error bound, and falls back to the FP64
// evaluation of the reference for the (rare) pairs whose FP32 value lies within the bound of a
// threshold.  newN is therefore bit-identical to the all-FP64 kernel; tao is summed in FP64.
//
// Error bound (u = 2^-24; M_k bounds |p_k - o_k| and |x_k - o_k| in coordinate k for the points
// that take the FP32 path; v = p - x, vv = |v|^2 as the reference computes it in FP64):
//   p^_k = fl32(fl64(p_k - o_k)), x^_k likewise:  |p^_k - (p_k - o_k)| <= u M_k (1 + 2^-28)
//   v^_k = fl32(p^_k - x^_k):                     E_k = |v^_k - v_k| <= u (|v_k| + 2 M_k) + O(u^2)
//   S^   = fma(v^z,v^z, fma(v^y,v^y, v^x v^x)):   |S^ - sum v^_k^2| <= 3.01 u sum v^_k^2
//   |sum v^_k^2 - vv| <= sum E_k (2 |v_k| + E_k) <= 2 u vv + 4 u sqrt(vv) |M| + O(u^2)   (Cauchy-Schwarz)
//   => |S^ - vv64| <= u (5.01 vv + 4 sqrt(vv) |M|) + O(u^2) + 1e-15 vv,  |M| = sqrt(M_x^2 + M_y^2 + M_z^2).
//   vv -/+ this bound is increasing in vv, so its value at the threshold bounds the classification
//   error there; the kernel uses u (8 T + 6 sqrt(T) |M|) plus the FP32 rounding of T for
//   T = searchR2 and T = R2.  M_y = M_z = 1.5 cell edges, M_x = n_cells_x / 2 + 1 cell edges (all
//   x (1 + 1e-6)): every deposit of the unit's 9 ranges that was not clamped into the grid from
//   outside satisfies it, and one that does not is more than a cell edge > sqrt(searchR2) from
//   every hit point lying inside the unit's cells, so it is dropped.  A hit point that does not lie
//   inside the unit's cells (clamped), and a slice holding a deposit whose normal has a component
//   above 1.0001, take the FP64 path wholesale.
//   nn^ = fma(n^z,N^z, fma(n^y,N^y, n^x N^x)) with |n_k|, |N_k| <= 1.0001:
//   |nn^ - nn64| <= 5.1 u * 3.0006 < 9.2e-7; the kernel uses 1.5e-6.
// Pairs are classified by comparing S^, nn^ with the FP32 thresholds while FMNMX3 tracks the
// smallest |S^ - T|, |nn^ - 1e-3| a lane has seen in the slice; a lane whose minimum is inside the
// bound repeats its share pair by pair (careful_pairs: FP64 only for the pairs inside the bound).
#ifdef SPPM_HOST_EMULATION  // tests/device_host only: the two halves of a packed FP32 operation, same roundings
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#else
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rr;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(rr) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rr);
}
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rr;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(rr) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rr);
}
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rr;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rr) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rr);
}
#endif
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }

// Organisation: persistent, fully autonomous warps; no CTA-wide barrier after start-up.  A warp
// takes a work unit (GUnit: <= 64 consecutive hit points of one run of x-adjacent occupied cells,
// so sparse cells are merged and dense cells are split) from a global counter.  The deposits
// that can reach the unit are 9 contiguous ranges of the cell-sorted list; their concatenation
// is cut into slices of 32.  Lane 0 streams the slices into a warp-private, double-buffered raw
// buffer with TMA bulk copies (cp.async.bulk + one mbarrier per buffer); the 32 lanes convert
// one deposit each to FP32 relative to the unit's centre, drop those farther than
// sqrt(searchR2) from the bounding box of the unit's hit points (ballot compaction), and write
// the survivors as interleaved pairs.  Then lane = hit point (two hit points per lane for units
// of 33..64; 32/hl deposit sub-slices when the unit has hl <= 16 hit points) and every lane
// walks the warp's pairs with broadcast 128-bit shared-memory loads.  A deposit outside the
// 27-cell neighbourhood of a lane's hit point simply fails the distance test.  Every hit point
// belongs to exactly one unit.
// A unit whose ranges hold more than ITEM_DEP deposits is processed by several warps (work items,
// see unit_parts_kernel); only those units accumulate with atomics.
#ifndef SPPM_W_WARPS
#define SPPM_W_WARPS 8
#endif
constexpr int W_WARPS = SPPM_W_WARPS;
constexpr int W_THREADS = W_WARPS * 32;
#ifndef SPPM_W_SLICE
#define SPPM_W_SLICE 32
#endif
constexpr int W_SLICE = SPPM_W_SLICE;                          // deposits per warp slice (32 or 64)
constexpr int W_RAW_BYTES = W_SLICE * (int)sizeof(Deposit);   // 2560
constexpr int F_PAIR_F4 = 6;                                  // 16-byte words per converted pair (96 B)
constexpr int W_CONV_BYTES = (W_SLICE / 2) * F_PAIR_F4 * 16;  // 1536
constexpr int W_FULL_BYTES = 3 * 32 * 8;                       // per lane: flux sums of the deposits the whole unit accepts
constexpr int W_PER_WARP = 2 * W_RAW_BYTES + W_CONV_BYTES + 128 + W_FULL_BYTES;  // + kept-index list + full-accept sums
constexpr size_t F_SMEM = (size_t)W_WARPS * W_PER_WARP;
constexpr int W_CHUNK = 64;                                   // hit points per chunk (2 per lane)
#define SPPM_F_U 5.9604644775390625e-8 /* 2^-24 */
#define SPPM_F_NMAX 1.0001
#define SPPM_F_NBAND 1.5e-6f
#define SPPM_F_PAD 3.0e18f /* coordinate of a deposit that can never be close */

struct HpF {  // one hit point in FP32, relative to the cell centre
  float x, y, z, nx, ny, nz;
  float r2f;   // fl32(R2)
  float band;  // classification bound for S^ against searchR2 / R2
};

// FP64 evaluation (the reference's operation order) of one lane's share of a warp slice:
// list positions 2j, 2j+1 for j = j0, j0+jstep, ...; position q is raw record kidx[q] of the slice
// (use_list) or raw record q itself.
__device__ __noinline__ void exact_pairs(const HpSoA s, const int i, const double search_r2, const double2* raw,
                                         const unsigned char* kidx, const bool use_list, const int n_list, const int j0,
                                         const int jstep, int& cnt, unsigned& cand, double& t0, double& t1, double& t2) {
  const double px = s.px[i], py = s.py[i], pz = s.pz[i], nx = s.nx[i], ny = s.ny[i], nz = s.nz[i], r2 = s.r2[i];
  int c = 0;
  unsigned cd = 0;
  double a0 = 0, a1 = 0, a2 = 0;
  const int npair = (n_list + 1) >> 1;
  for (int j = j0; j < npair; j += jstep) {
    for (int h = 0; h < 2; h++) {
      const int qi = 2 * j + h;
      if (qi >= n_list) break;
      const int d = use_list ? (int)kidx[qi] : qi;
      const double2* q = raw + d * 5;
      const double2 q0 = q[0], q1 = q[1], q2 = q[2];
      const double vx = px - q0.x, vy = py - q0.y, vz = pz - q1.x;  // Raytracer.h:152
      const double vv = (vx * vx + vy * vy) + vz * vz;
      if (vv < search_r2) {  // flann radiusSearch, strict '<' (Q3)
        cd++;
        const double nn = (nx * q1.y + ny * q2.x) + nz * q2.y;
        if (nn > SPPM_NORMAL_AGREE && vv <= r2) {  // Raytracer.h:154
          const double2 q3 = q[3], q4 = q[4];
          c++;
          a0 += q3.x; a1 += q3.y; a2 += q4.x;
        }
      }
    }
  }
  cnt = c; cand = cd; t0 = a0; t1 = a1; t2 = a2;
}

// A lane whose fast pass came within the error bound of a threshold repeats its share of the
// slice pair by pair: FP32 classification where it is decisive, the FP64 evaluation of the
// reference for the (hit point, deposit) pairs inside the bound only.  Same order of summation.
__device__ __noinline__ void careful_pairs(const HpSoA s, const int i, const double search_r2, const double2* raw,
                                           const unsigned char* kidx, const int n_list, const float4* cv, const HpF h,
                                           const float sr2f, const int j0, const int jstep, int& cnt, unsigned& cand,
                                           double& t0, double& t1, double& t2) {
  const float thrf = (float)SPPM_NORMAL_AGREE;
  const double px = s.px[i], py = s.py[i], pz = s.pz[i], nx = s.nx[i], ny = s.ny[i], nz = s.nz[i], r2 = s.r2[i];
  int c = 0;
  unsigned cd = 0;
  double a0 = 0, a1 = 0, a2 = 0;
  const int npair = (n_list + 1) >> 1;
  for (int j = j0; j < npair; j += jstep) {
    const float* f = reinterpret_cast<const float*>(cv + j * F_PAIR_F4);
    const double* g = reinterpret_cast<const double*>(cv + j * F_PAIR_F4 + 3);
    for (int hf = 0; hf < 2; hf++) {
      const int qi = 2 * j + hf;
      if (qi >= n_list) break;
      const float vx = h.x - f[hf], vy = h.y - f[2 + hf], vz = h.z - f[4 + hf];
      const float vv = fmaf(vz, vz, fmaf(vy, vy, vx * vx));
      const float nn = fmaf(h.nz, f[10 + hf], fmaf(h.ny, f[8 + hf], h.nx * f[6 + hf]));
      const bool unc = !(fabsf(vv - sr2f) > h.band) || !(fabsf(vv - h.r2f) > h.band) || !(fabsf(nn - thrf) > SPPM_F_NBAND);
      bool is_cand, is_acc;
      if (unc) {
        const double2* q = raw + (int)kidx[qi] * 5;
        const double2 q0 = q[0], q1 = q[1], q2 = q[2];
        const double wx = px - q0.x, wy = py - q0.y, wz = pz - q1.x;  // Raytracer.h:152
        const double ww = (wx * wx + wy * wy) + wz * wz;
        const double mm = (nx * q1.y + ny * q2.x) + nz * q2.y;
        is_cand = ww < search_r2;                                      // flann radiusSearch, strict '<' (Q3)
        is_acc = is_cand && mm > SPPM_NORMAL_AGREE && ww <= r2;        // Raytracer.h:154
      } else {
        is_cand = vv < sr2f;
        is_acc = is_cand && vv < h.r2f && nn > thrf;
      }
      cd += is_cand ? 1u : 0u;
      if (is_acc) {
        c++;
        a0 += g[3 * hf]; a1 += g[3 * hf + 1]; a2 += g[3 * hf + 2];
      }
    }
  }
  cnt = c; cand = cd; t0 = a0; t1 = a1; t2 = a2;
}

// FP32 classification of one lane's pairs against its HB hit points; uv / un return the
// smallest distance of a classified value from its threshold.  UNI: R2 == searchR2 for all hit
// points of the chunk.  A converted pair is 96 bytes: {x0,x1,y0,y1} {z0,z1,nx0,nx1}
// {ny0,ny1,nz0,nz1} in FP32 and the two fluxes {fx0,fy0} {fz0,fx1} {fy1,fz1} in FP64; the
// accumulation is predicated, not branched.
constexpr int HBMAX = UNIT_HP > 32 ? 2 : 1;  // hit points per lane
#ifndef SPPM_PAIR_UNROLL
#define SPPM_PAIR_UNROLL 2
#endif
constexpr int PAIR_UNROLL = SPPM_PAIR_UNROLL;
// NN = false: every hit point of the unit has the same normal, the normal test of Raytracer.h:154 was made
// once per deposit (in FP64, exactly) when the slice was culled, and the loop tests distances only.
template <bool UNI, bool COUNT, int HB, bool NN = true>
__device__ __forceinline__ void fast_pairs(const float4* __restrict__ cv, const int j0, const int jstep, const int npair,
                                           const HpF (&h)[HBMAX], const float sr2f, int (&c)[HBMAX], unsigned& cand,
                                           double (&acc)[HBMAX][3], float (&uv)[HBMAX], float (&un)[HBMAX]) {
  const float thrf = (float)SPPM_NORMAL_AGREE;
  unsigned cd = 0;
  const float4* q = cv + j0 * F_PAIR_F4;
  const int stride = jstep * F_PAIR_F4;
#pragma unroll PAIR_UNROLL
  for (int j = j0; j < npair; j += jstep, q += stride) {
    const float4 A = q[0], B = q[1];
    float4 C = make_float4(0.f, 0.f, 0.f, 0.f);
    if (NN) C = q[2];
    const double2 F0 = *reinterpret_cast<const double2*>(q + 3), F1 = *reinterpret_cast<const double2*>(q + 4),
                  F2 = *reinterpret_cast<const double2*>(q + 5);
#pragma unroll
    for (int b = 0; b < HB; b++) {
      const float2 vx = f2sub(bc2(h[b].x), make_float2(A.x, A.y));
      const float2 vy = f2sub(bc2(h[b].y), make_float2(A.z, A.w));
      const float2 vz = f2sub(bc2(h[b].z), make_float2(B.x, B.y));
      float2 vv = f2mul(vx, vx);
      vv = f2fma(vy, vy, vv);
      vv = f2fma(vz, vz, vv);
      float2 nn = make_float2(1.f, 1.f);
      if (NN) {
        nn = f2mul(bc2(h[b].nx), make_float2(B.z, B.w));
        nn = f2fma(bc2(h[b].ny), make_float2(C.x, C.y), nn);
        nn = f2fma(bc2(h[b].nz), make_float2(C.z, C.w), nn);
        const float2 dn = f2sub(nn, bc2(thrf));
        un[b] = fminf(un[b], fminf(fabsf(dn.x), fabsf(dn.y)));
      }
      const float2 dv = f2sub(vv, bc2(sr2f));
      uv[b] = fminf(uv[b], fminf(fabsf(dv.x), fabsf(dv.y)));
      bool in0 = vv.x < sr2f, in1 = vv.y < sr2f;
      if (COUNT) cd += (__float_as_uint(dv.x) >> 31) + (__float_as_uint(dv.y) >> 31);  // dv < 0 (never -0)
      if (!UNI) {
        const float2 dr = f2sub(vv, bc2(h[b].r2f));
        uv[b] = fminf(uv[b], fminf(fabsf(dr.x), fabsf(dr.y)));
        in0 = in0 && vv.x < h[b].r2f;
        in1 = in1 && vv.y < h[b].r2f;
      }
      if (in0 && (!NN || nn.x > thrf)) {
        c[b]++;
        acc[b][0] += F0.x; acc[b][1] += F0.y; acc[b][2] += F1.x;
      }
      if (in1 && (!NN || nn.y > thrf)) {
        c[b]++;
        acc[b][0] += F1.y; acc[b][1] += F2.x; acc[b][2] += F2.y;
      }
    }
  }
  cand += cd;
}

#ifndef SPPM_GATHER_MINBLOCKS
#define SPPM_GATHER_MINBLOCKS 2
#endif
// Algorithmic work counters of the filtered gather (tools/gather_work_model.py): they exist in the CPU
// emulation of tests/device_host only and expand to nothing in the library.
#ifdef SPPM_HOST_EMULATION
enum { WK_ITEMS, WK_SLICES, WK_VISITS, WK_KEPT, WK_PAIRS, WK_PAIR_SLOTS, WK_REDO_LANES, WK_FULL, WK_N };
extern unsigned long long dh_work[WK_N];
#define SPPM_WORK(k, v)                                                                          \
  do {                                                                                           \
    if (lane == 0) __atomic_fetch_add(&dh_work[k], (unsigned long long)(v), __ATOMIC_RELAXED);   \
  } while (0)
#else
#define SPPM_WORK(k, v) ((void)0)
#endif
template <bool COUNT>
__global__ void __launch_bounds__(W_THREADS, SPPM_GATHER_MINBLOCKS) gather_filtered_kernel(const GatherArgs a, unsigned int* work_counter) {
  extern __shared__ __align__(128) unsigned char wsm[];
  __shared__ __align__(8) uint64_t bar[W_WARPS][2];
  __shared__ int run_pre[W_WARPS][12];  // per warp: run starts [0..8] are kept in registers; prefix of run lengths here
  __shared__ int run_st[W_WARPS][12];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  unsigned char* my = wsm + w * W_PER_WARP;
  float4* conv = reinterpret_cast<float4*>(my + 2 * W_RAW_BYTES);
  unsigned char* kidx = my + 2 * W_RAW_BYTES + W_CONV_BYTES;
  double* fsum = reinterpret_cast<double*>(my + 2 * W_RAW_BYTES + W_CONV_BYTES + 128);  // [3][32]
  if (lane == 0) {
    mbar_init(&bar[w][0], 1);
    mbar_init(&bar[w][1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const double cell = 1.0 / a.grid.inv_cell;
  const double M_yz = 1.5 * cell * (1.0 + 1e-6);
  const double half_in = 0.5 * cell * (1.0 + 1e-7);  // a hit point that really lies in its cell
  const float sr2f = __double2float_rn(a.search_r2);
  unsigned phase0 = 0, phase1 = 0;  // parity of the next completion of the warp's two barriers
  unsigned long long cand = 0, acc_total = 0;
  for (;;) {
    int it = 0;
    if (lane == 0) it = (int)atomicAdd(work_counter, 1u);
    it = __shfl_sync(FULL, it, 0);
    if (it >= a.item_off[a.n_units]) break;
    int u = 0;
    {  // 32-ary search: the last u with item_off[u] <= it
      int lo = 0, hi = a.n_units;
      while (hi - lo > 1) {
        const int step = (hi - lo + 31) >> 5;
        const int probe = lo + (lane + 1) * step;
        const bool le = probe < hi && a.item_off[probe] <= it;
        const int k = __popc(__ballot_sync(FULL, le));  // off[] is non-decreasing: the lanes that answer yes are a prefix
        lo += k * step;
        hi = min(hi, lo + step);
      }
      u = lo;
    }
    const int off_u = a.item_off[u];
    const int part = it - off_u;
    const bool shared_unit = a.item_off[u + 1] - off_u > 1;  // several warps accumulate into this unit's hit points
    const GUnit* un = a.units + u;
    const int h0 = un->hp_begin, H = un->hp_count, ncx = un->n_cells_x;
    const double ox = un->org[0], oy = un->org[1], oz = un->org[2];
    // the 9 deposit ranges and the prefix sums of their lengths (lanes 0..8, then shared)
    int D;
    {
      int st = 0, ln = 0;
      if (lane < 9) {
        const int2 run = un->run[lane];
        if (run.y > run.x) {
          st = a.dep_off[run.x];
          ln = a.dep_off[run.y] - st;
        }
      }
      int inc = ln;  // inclusive prefix over lanes 0..8
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, off);
        if (lane >= off) inc += t;
      }
      D = __shfl_sync(FULL, inc, 8);
      __syncwarp();
      if (lane < 9) {
        run_st[w][lane] = st;
        run_pre[w][lane] = inc - ln;  // exclusive prefix = position of the run's first deposit
      }
      if (lane == 9) run_pre[w][9] = D;
      __syncwarp();
    }
    if (D == 0) continue;
    SPPM_WORK(WK_ITEMS, 1);
    constexpr int ITEM_SLICES = ITEM_DEP / W_SLICE;
    const int sl0 = part * ITEM_SLICES;  // this item: slices [sl0, sl0 + n_sl) of the concatenation
    const int n_sl = min(ITEM_SLICES, (D + W_SLICE - 1) / W_SLICE - sl0);
    // lane 0: stream the item's next slice (32 deposits of the concatenation from position cur_p,
    // which lies in run cur_r) into raw buffer k & 1; the cursor advances slice by slice
    int cur_p = sl0 * W_SLICE, cur_r = 0;
    if (lane == 0) {
#pragma unroll
      for (int i = 1; i < 9; i++) cur_r += (cur_p >= run_pre[w][i]) ? 1 : 0;
    }
    auto issue = [&](int k) {
      const int buf = k & 1;
      int left = min(W_SLICE, D - cur_p);
      mbar_expect_tx(&bar[w][buf], (unsigned)left * (unsigned)sizeof(Deposit));
      unsigned char* dst = my + buf * W_RAW_BYTES;
      while (left > 0) {
        const int end = run_pre[w][cur_r + 1];
        const int take = min(end - cur_p, left);
        if (take > 0) {
          bulk_g2s(dst, a.dep + run_st[w][cur_r] + (cur_p - run_pre[w][cur_r]), (unsigned)take * (unsigned)sizeof(Deposit),
                   &bar[w][buf]);
          dst += take * sizeof(Deposit);
          left -= take;
          cur_p += take;
        }
        if (cur_p >= end) cur_r++;
      }
    };
    if (lane == 0) {
      issue(0);
      if (n_sl > 1) issue(1);
    }
    // hit points of the unit: lane = hit point (+32), sub-slices when there are <= 16
    const bool two = HBMAX > 1 && H > 32;
    int hl = 32, lg = 5;
    if (!two) {
      hl = 1; lg = 0;
      while (hl < H) { hl <<= 1; lg++; }
    }
    const int sub = 32 >> lg, my_h = lane & (hl - 1), sub_id = lane >> lg;
    int idx[HBMAX];
    bool valid[HBMAX];
    idx[0] = h0 + my_h;
    valid[0] = my_h < H;
    if (HBMAX > 1) {
      idx[HBMAX - 1] = h0 + 32 + lane;
      valid[HBMAX - 1] = two && 32 + lane < H;
    }
    const double half_x = 0.5 * ncx * cell * (1.0 + 1e-7), M_x = (0.5 * ncx + 1.0) * cell * (1.0 + 1e-6);
    const double Mn = sqrt(M_x * M_x + 2.0 * M_yz * M_yz);
    const double band_v = SPPM_F_U * (8.0 * a.search_r2 + 6.0 * sqrt(a.search_r2) * Mn) + 1e-13 * Mn * Mn;
    HpF hp[HBMAX];
    bool far_hp[HBMAX] = {};
    bool uni_me = true, nuni_me = true;
    const double rnx = a.s.nx[h0], rny = a.s.ny[h0], rnz = a.s.nz[h0];  // the unit's first normal (uniform loads)
    double rmax = 0, rmin = 1e300;  // largest / smallest R2 among the unit's hit points on the FP32 path
    float blo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, bhi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
#pragma unroll
    for (int b = 0; b < HBMAX; b++) {
      hp[b] = HpF{-SPPM_F_PAD, -SPPM_F_PAD, -SPPM_F_PAD, 0.f, 0.f, 0.f, 0.f, 0.f};  // never close, not even to a padding deposit
      if (valid[b]) {
        const int i = idx[b];
        const double ax = a.s.px[i] - ox, ay = a.s.py[i] - oy, az = a.s.pz[i] - oz;
        const double nx = a.s.nx[i], ny = a.s.ny[i], nz = a.s.nz[i], r2 = a.s.r2[i];
        far_hp[b] = !(fabs(ax) <= half_x && fabs(ay) <= half_in && fabs(az) <= half_in && fabs(nx) <= SPPM_F_NMAX &&
                      fabs(ny) <= SPPM_F_NMAX && fabs(nz) <= SPPM_F_NMAX && r2 >= 0 && r2 < 1e30);
        if (!far_hp[b]) {
          hp[b].x = __double2float_rn(ax); hp[b].y = __double2float_rn(ay); hp[b].z = __double2float_rn(az);
          hp[b].nx = __double2float_rn(nx); hp[b].ny = __double2float_rn(ny); hp[b].nz = __double2float_rn(nz);
          hp[b].r2f = __double2float_rn(r2);
          const double Tm = fmax(r2, a.search_r2);  // the larger threshold this hit point is classified against
          hp[b].band = __double2float_ru(1.01 * (SPPM_F_U * (8.0 * Tm + 6.0 * sqrt(Tm) * Mn + 2.0 * Tm) + 1e-13 * Mn * Mn));
          blo[0] = fminf(blo[0], hp[b].x); blo[1] = fminf(blo[1], hp[b].y); blo[2] = fminf(blo[2], hp[b].z);
          bhi[0] = fmaxf(bhi[0], hp[b].x); bhi[1] = fmaxf(bhi[1], hp[b].y); bhi[2] = fmaxf(bhi[2], hp[b].z);
          uni_me = uni_me && (r2 == a.search_r2);
          nuni_me = nuni_me && (nx == rnx && ny == rny && nz == rnz);
          rmax = fmax(rmax, r2);
          rmin = fmin(rmin, r2);
        }
      }
    }
    const bool uni = __all_sync(FULL, uni_me);
    // All hit points of the unit (on the FP32 path) carry the same normal, bit for bit — the units of a plane, i.e.
    // most of them: hp->n.dot(N) (Raytracer.h:154) then has ONE value per deposit, which the cull below computes in
    // FP64 with the reference's expression, and the pair loop leaves the normal out (fast_pairs<.., NN = false>).
    // Not in the diagnostic COUNT build: cand_pairs counts pairs by distance alone.
    const bool nuni = !COUNT && __all_sync(FULL, nuni_me);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      rmax = fmax(rmax, __shfl_xor_sync(FULL, rmax, off));
      rmin = fmin(rmin, __shfl_xor_sync(FULL, rmin, off));
    }
    // Cull radius: a pair is accepted only if vv < searchR2 AND vv <= R2 of its hit point (Raytracer.h:146,154), so a
    // deposit farther than min(searchR2, max R2 of the unit) from the unit's bounding box cannot be accepted by any
    // of its hit points.  Once HitPoint::update has shrunk the radii (PROGRESSIVE passes) this removes most of the
    // pairs between sqrt(R2) and sqrt(searchR2).  The diagnostic candidate count (vv < searchR2) needs them all.
    const float cull_thr = __double2float_ru((COUNT ? a.search_r2 : fmin(a.search_r2, rmax)) + 2.0 * band_v);
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        blo[k] = fminf(blo[k], __shfl_xor_sync(FULL, blo[k], off));
        bhi[k] = fmaxf(bhi[k], __shfl_xor_sync(FULL, bhi[k], off));
      }
    // Whole-unit accept.  A deposit whose distance to the FARTHEST corner of the unit's bounding box is below
    // min(searchR2, smallest R2 of the unit) by more than the classification bound is inside the ball of every hit
    // point of the unit; if it also passes the unit's (single, exact) normal test, all H hit points accept it
    // (Raytracer.h:154).  Such deposits never reach the pair loop: their flux goes to one sum per lane and the
    // unit's hit points receive count and sum once, at the end.  The bound: the far-corner offset per coordinate
    // is max(x^ - lo^, hi^ - x^) over FP32 values that bracket every hit point's FP32 coordinate, so the FP32 sum
    // of squares over-estimates each pair's S^ up to rounding, and S^ is within band_v of the reference's vv
    // (derivation above): S^far < T - 2 band_v  =>  vv < T for every hit point.  Dense maps (4K film) accept a
    // third to a half of the surviving deposits this way; at 1024 x 768 a unit is as wide as the radius and few do.
    // Only units that are small against the radius try: the deposits that qualify lie within sqrt(T) - half
    // the box diagonal of the box centre, and where that region is small the test costs more than it saves.
    const double full_T = fmin(a.search_r2, rmin) - 2.0 * band_v;
    const double hdx = 0.5 * ((double)bhi[0] - (double)blo[0]), hdy = 0.5 * ((double)bhi[1] - (double)blo[1]), hdz = 0.5 * ((double)bhi[2] - (double)blo[2]);
    const bool full_ok = !COUNT && a.full_accept && nuni && !__any_sync(FULL, far_hp[0] || far_hp[HBMAX - 1]) &&
                         (hdx * hdx + hdy * hdy + hdz * hdz) * 9.0 < full_T;  // half diagonal < a third of the radius
    const float full_thr = full_ok && full_T > 0 ? __double2float_rd(full_T) : -1.f;
    int fcnt = 0;
    if (full_ok) fsum[lane] = fsum[32 + lane] = fsum[64 + lane] = 0;
    int cnt[HBMAX] = {};
    double sum[HBMAX][3] = {};
    unsigned cand32 = 0;
    for (int k = 0; k < n_sl; k++) {
      const int buf = k & 1;
      const int fill = min(W_SLICE, D - (sl0 + k) * W_SLICE);
      const double2* raw = reinterpret_cast<const double2*>(my + buf * W_RAW_BYTES);
      if (buf == 0) { mbar_wait(&bar[w][0], phase0); phase0 ^= 1; } else { mbar_wait(&bar[w][1], phase1); phase1 ^= 1; }
      // convert one deposit per lane; keep it if it can be within sqrt(searchR2) of the unit's
      // bounding box.  A deposit farther than M from the centre in some coordinate (clamped into
      // a boundary cell from outside the grid) is more than a cell edge away from every hit
      // point lying inside the unit's cells and is dropped.
      bool badn = false;
      int n_k = 0;
#pragma unroll
      for (int rd = 0; rd < W_SLICE / 32; rd++) {  // W_SLICE / 32 deposits per lane
        const int di = rd * 32 + lane;
        bool keep = false;
        float fx = 0, fy = 0, fz = 0, gx = 0, gy = 0, gz = 0;
        double2 q3 = make_double2(0, 0), q4 = make_double2(0, 0);
        if (di < fill) {
          const double2* q = raw + di * 5;
          const double2 q0 = q[0], q1 = q[1], q2 = q[2];
          q3 = q[3]; q4 = q[4];
          const double ax = q0.x - ox, ay = q0.y - oy, az = q1.x - oz;
          const bool out = !(fabs(ax) <= M_x && fabs(ay) <= M_yz && fabs(az) <= M_yz);
          bool n_ok = true;
          if (nuni) n_ok = (rnx * q1.y + rny * q2.x) + rnz * q2.y > SPPM_NORMAL_AGREE;  // exact: Raytracer.h:154 for the whole unit
          else badn = badn || !(fabs(q1.y) <= SPPM_F_NMAX && fabs(q2.x) <= SPPM_F_NMAX && fabs(q2.y) <= SPPM_F_NMAX);
          fx = __double2float_rn(ax); fy = __double2float_rn(ay); fz = __double2float_rn(az);
          gx = __double2float_rn(q1.y); gy = __double2float_rn(q2.x); gz = __double2float_rn(q2.y);
          const float ex = fmaxf(fmaxf(blo[0] - fx, fx - bhi[0]), 0.f), ey = fmaxf(fmaxf(blo[1] - fy, fy - bhi[1]), 0.f),
                      ez = fmaxf(fmaxf(blo[2] - fz, fz - bhi[2]), 0.f);
          keep = !out && n_ok && !(fmaf(ez, ez, fmaf(ey, ey, ex * ex)) > cull_thr);
          if (keep && full_thr > 0.f) {
            const float mx = fmaxf(fx - blo[0], bhi[0] - fx), my2 = fmaxf(fy - blo[1], bhi[1] - fy), mz = fmaxf(fz - blo[2], bhi[2] - fz);
            if (fmaf(mz, mz, fmaf(my2, my2, mx * mx)) < full_thr) {  // every hit point of the unit accepts this deposit
              keep = false;
              fcnt++;
              fsum[lane] += q3.x; fsum[32 + lane] += q3.y; fsum[64 + lane] += q4.x;
            }
          }
        }
        const unsigned kmask = __ballot_sync(FULL, keep);
        if (keep) {
          const int pos = n_k + __popc(kmask & ((1u << lane) - 1u));
          float4* pr = conv + (pos >> 1) * F_PAIR_F4;
          float* f = reinterpret_cast<float*>(pr) + (pos & 1);
          f[0] = fx; f[2] = fy; f[4] = fz; f[6] = gx; f[8] = gy; f[10] = gz;
          double* g = reinterpret_cast<double*>(pr + 3) + (pos & 1) * 3;
          g[0] = q3.x; g[1] = q3.y; g[2] = q4.x;
          kidx[pos] = (unsigned char)di;
        }
        n_k += __popc(kmask);
      }
      const bool any_badn = __any_sync(FULL, badn);
      SPPM_WORK(WK_SLICES, 1);
      SPPM_WORK(WK_VISITS, fill);
      SPPM_WORK(WK_KEPT, n_k);
      SPPM_WORK(WK_PAIRS, (long long)n_k * H);
      SPPM_WORK(WK_PAIR_SLOTS, (long long)((n_k + 1) & ~1) * (two ? 64 : hl));
      if (lane == 0 && (n_k & 1)) {  // pad the last pair
        float4* pr = conv + (n_k >> 1) * F_PAIR_F4;
        float* f = reinterpret_cast<float*>(pr) + 1;
        f[0] = SPPM_F_PAD; f[2] = SPPM_F_PAD; f[4] = SPPM_F_PAD; f[6] = 0.f; f[8] = 0.f; f[10] = 0.f;
        double* g = reinterpret_cast<double*>(pr + 3) + 3;
        g[0] = 0; g[1] = 0; g[2] = 0;
      }
      __syncwarp();
      {
        int c_t[HBMAX] = {};
        double t[HBMAX][3] = {};
        float uv[HBMAX], un2[HBMAX];
#pragma unroll
        for (int b = 0; b < HBMAX; b++) uv[b] = un2[b] = 3.0e38f;
        unsigned cd_t = 0;
        const int npair = (n_k + 1) >> 1;
        if (!any_badn) {
          if (!COUNT && nuni) {
            if (HBMAX > 1 && two) {
              if (uni) fast_pairs<true, COUNT, HBMAX, false>(conv, sub_id, sub, npair, hp, sr2f, c_t, cd_t, t, uv, un2);
              else fast_pairs<false, COUNT, HBMAX, false>(conv, sub_id, sub, npair, hp, sr2f, c_t, cd_t, t, uv, un2);
            } else {
              if (uni) fast_pairs<true, COUNT, 1, false>(conv, sub_id, sub, npair, hp, sr2f, c_t, cd_t, t, uv, un2);
              else fast_pairs<false, COUNT, 1, false>(conv, sub_id, sub, npair, hp, sr2f, c_t, cd_t, t, uv, un2);
            }
          } else if (HBMAX > 1 && two) {
            if (uni) fast_pairs<true, COUNT, HBMAX>(conv, sub_id, sub, npair, hp, sr2f, c_t, cd_t, t, uv, un2);
            else fast_pairs<false, COUNT, HBMAX>(conv, sub_id, sub, npair, hp, sr2f, c_t, cd_t, t, uv, un2);
          } else {
            if (uni) fast_pairs<true, COUNT, 1>(conv, sub_id, sub, npair, hp, sr2f, c_t, cd_t, t, uv, un2);
            else fast_pairs<false, COUNT, 1>(conv, sub_id, sub, npair, hp, sr2f, c_t, cd_t, t, uv, un2);
          }
        }
        bool redo[HBMAX];
        bool any_redo = false;
#pragma unroll
        for (int b = 0; b < HBMAX; b++) {
          redo[b] = valid[b] && (any_badn || far_hp[b] || !(uv[b] > hp[b].band) || !(un2[b] > SPPM_F_NBAND));
          any_redo = any_redo || redo[b];
        }
#ifdef SPPM_HOST_EMULATION
        {
          const unsigned redo_mask = __ballot_sync(FULL, any_redo);
          SPPM_WORK(WK_REDO_LANES, __popc(redo_mask));
        }
#endif
        if (any_redo) {
          // the fast counters of this lane cover both hit points: recount them exactly
          unsigned cd_x = 0;
#pragma unroll
          for (int b = 0; b < HBMAX; b++) {
            if (!valid[b]) continue;
            const bool whole = any_badn || far_hp[b];  // not list based: every raw record of the slice in FP64
            int c_x;
            unsigned cd_b;
            double x0, x1, x2;
            if (whole) exact_pairs(a.s, idx[b], a.search_r2, raw, kidx, false, fill, sub_id, sub, c_x, cd_b, x0, x1, x2);
            else careful_pairs(a.s, idx[b], a.search_r2, raw, kidx, n_k, conv, hp[b], sr2f, sub_id, sub, c_x, cd_b, x0, x1, x2);
            c_t[b] = c_x; t[b][0] = x0; t[b][1] = x1; t[b][2] = x2;
            cd_x += cd_b;
          }
          cd_t = cd_x;
        }
#pragma unroll
        for (int b = 0; b < HBMAX; b++) {
          cnt[b] += c_t[b];
          sum[b][0] += t[b][0]; sum[b][1] += t[b][1]; sum[b][2] += t[b][2];
        }
        cand32 += cd_t;
      }
      __syncwarp();  // raw / conv / kidx consumed by every lane
      if (lane == 0 && k + 2 < n_sl) issue(k + 2);
    }
    acc_total += (unsigned)(cnt[0] + (HBMAX > 1 ? cnt[HBMAX - 1] : 0));
    cand += cand32;
    // whole-unit accepts of this item: one count and one flux sum for all H hit points
    int fc_all = 0;
    double fall[3] = {0, 0, 0};
    if (full_ok && __any_sync(FULL, fcnt > 0)) {
      fc_all = fcnt;
      fall[0] = fsum[lane]; fall[1] = fsum[32 + lane]; fall[2] = fsum[64 + lane];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        fc_all += __shfl_xor_sync(FULL, fc_all, off);
        fall[0] += __shfl_xor_sync(FULL, fall[0], off);
        fall[1] += __shfl_xor_sync(FULL, fall[1], off);
        fall[2] += __shfl_xor_sync(FULL, fall[2], off);
      }
      if (lane == 0) acc_total += (unsigned long long)fc_all * (unsigned)H;
      SPPM_WORK(WK_FULL, fc_all);
    }
    // combine the sub-slices of a hit point (lanes my_h, my_h + hl, ...)
    for (int o = hl; o < 32; o <<= 1) {
#pragma unroll
      for (int b = 0; b < HBMAX; b++) {
        cnt[b] += __shfl_xor_sync(FULL, cnt[b], o);
        sum[b][0] += __shfl_xor_sync(FULL, sum[b][0], o);
        sum[b][1] += __shfl_xor_sync(FULL, sum[b][1], o);
        sum[b][2] += __shfl_xor_sync(FULL, sum[b][2], o);
      }
    }
    if (sub_id == 0) {
#pragma unroll
      for (int b = 0; b < HBMAX; b++) {
        cnt[b] += fc_all;
        if (valid[b] && cnt[b] > 0) {  // hp->newN++, hp->tao += wgt.mul(flux)*PI_INV (Raytracer.h:155-156)
          const int i = idx[b];
          const double d0 = (a.s.wx[i] * (sum[b][0] + fall[0])) * SPPM_PI_INV, d1 = (a.s.wy[i] * (sum[b][1] + fall[1])) * SPPM_PI_INV,
                       d2 = (a.s.wz[i] * (sum[b][2] + fall[2])) * SPPM_PI_INV;
          if (shared_unit) {
            atomicAdd(&a.s.dN[i], cnt[b]);
            atomicAdd(&a.s.dtao[i], d0);
            atomicAdd(&a.s.dtao[a.n_hp + i], d1);
            atomicAdd(&a.s.dtao[2 * a.n_hp + i], d2);
          } else {
            a.s.dN[i] += cnt[b];
            a.s.dtao[i] += d0;
            a.s.dtao[a.n_hp + i] += d1;
            a.s.dtao[2 * a.n_hp + i] += d2;
          }
        }
      }
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    cand += __shfl_xor_sync(FULL, cand, off);
    acc_total += __shfl_xor_sync(FULL, acc_total, off);
  }
  if (lane == 0 && (cand || acc_total)) {
    atomicAdd(&a.counters->cand_pairs, cand);
    atomicAdd(&a.counters->acc_pairs, acc_total);
  }
}

// newN += dN, tao += dtao, then HitPoint::update (Raytracer.h:69-79)
__global__ void update_kernel(HpSoA s, long long n, int mode, double alpha) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int newN = s.newN[i] + s.dN[i];
  double t0 = s.tao[i] + s.dtao[i], t1 = s.tao[n + i] + s.dtao[n + i], t2 = s.tao[2 * n + i] + s.dtao[2 * n + i];
  s.dN[i] = 0;
  s.dtao[i] = 0; s.dtao[n + i] = 0; s.dtao[2 * n + i] = 0;
  int N = s.N[i];
  const bool skip = (mode == SPPM_UPDATE_REFERENCE) ? (N <= 0 || newN <= 0) : (newN <= 0);
  if (!skip) {
    const double k = (N + alpha * newN) / (N + newN);
    s.r2[i] *= k;
    t0 *= k; t1 *= k; t2 *= k;
    N = (int)(N + alpha * newN);  // int += double: computed in double, truncated
    newN = 0;
    s.N[i] = N;
  }
  s.newN[i] = newN;
  s.tao[i] = t0; s.tao[n + i] = t1; s.tao[2 * n + i] = t2;
}

// camera->pic[x][y] += tao * (1.0 / PI / (R2*cnt*onetime))  (Raytracer.h:288-293)
// Position of every hit point in the sorted map, by canonical rank (ascending okey = pixel, eye-path node):
// rank_of_raw scatters the first sort's order, canon gathers it through the second one.
__global__ void hp_rank_kernel(const uint32_t* raw_of_rank, long long n, uint32_t* rank_of_raw) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) rank_of_raw[raw_of_rank[r]] = (uint32_t)r;
}
__global__ void hp_canon_kernel(const uint32_t* raw_of_pos, const uint32_t* rank_of_raw, long long n, uint32_t* canon) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) canon[rank_of_raw[raw_of_pos[j]]] = (uint32_t)j;
}
// One thread per canonical rank; the first hit point of a pixel sums the pixel's hit points in canonical
// order (the order the reference's GetHitPoint appends a pixel's eye-path nodes to Hvec) and adds the sum
// to the zeroed film once: the film is the same bit for bit from run to run.  (A map imported in an order
// that separates a pixel's hit points gives several runs for the pixel, each added once.)
__global__ void resolve_kernel(HpSoA s, const uint32_t* canon, long long n, int cnt, int onetime, int w, double* film) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  uint32_t i = canon[r];
  const long long px = (long long)s.row[i] * w + s.col[i];
  if (r > 0) {
    const uint32_t q = canon[r - 1];
    if ((long long)s.row[q] * w + s.col[q] == px) return;
  }
  double a[3] = {0, 0, 0};
  for (long long k = r;;) {
    const double sc = 1.0 / SPPM_PI / (s.r2[i] * cnt * onetime);
    a[0] += s.tao[i] * sc;
    a[1] += s.tao[n + i] * sc;
    a[2] += s.tao[2 * n + i] * sc;
    if (++k >= n) break;
    i = canon[k];
    if ((long long)s.row[i] * w + s.col[i] != px) break;
  }
  double* o = film + (size_t)px * 3;
  atomicAdd(o + 0, a[0]);
  atomicAdd(o + 1, a[1]);
  atomicAdd(o + 2, a[2]);
}

// ---- shared statistics across eye passes -------------------------------------
// The reference averages independent PPM renders (every RayTracer::render starts from R2 = 2,
// Raytracer.h:429-458; its update() is a no-op, Q1).  Stochastic progressive photon mapping
// proper (Hachisuka & Jensen 2009, the method the reference's README names) keeps ONE set of
// statistics per pixel while the measurement point moves with the jittered camera:
//   N' = N + alpha M,   R2' = R2 (N + alpha M) / (N + M),   tau' = (tau + Phi) R2' / R2
// with M / Phi the photons / weighted flux gathered by the pixel's hit points in the pass, and
// L = tau / (PI R2 N_emitted).  Layout of px_store: r2[np] n[np] tau[3][np] phi[3][np] (f64) m[np] (i32).
struct PxStats {
  double *r2, *n, *tau, *phi;
  int32_t* m;
  long long np;
};
__global__ void shared_init_kernel(PxStats p, double r2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.np) return;
  p.r2[i] = r2;
  p.n[i] = 0;
  p.m[i] = 0;
  for (int k = 0; k < 3; k++) {
    p.tau[k * p.np + i] = 0;
    p.phi[k * p.np + i] = 0;
  }
}
// every hit point measures with the radius its pixel has reached; its own N / newN / tao restart
__global__ void shared_load_kernel(HpSoA s, long long n, PxStats p, int w) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long px = (long long)s.row[i] * w + s.col[i];
  s.r2[i] = p.r2[px];
  s.N[i] = 0;
  s.newN[i] = 0;
  s.tao[i] = 0; s.tao[n + i] = 0; s.tao[2 * n + i] = 0;
}
__global__ void shared_accum_kernel(HpSoA s, long long n, PxStats p, int w) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int m = s.newN[i];
  if (m <= 0) return;
  const long long px = (long long)s.row[i] * w + s.col[i];
  atomicAdd(p.m + px, m);
  atomicAdd(p.phi + px, s.tao[i]);
  atomicAdd(p.phi + p.np + px, s.tao[n + i]);
  atomicAdd(p.phi + 2 * p.np + px, s.tao[2 * n + i]);
}
__global__ void shared_update_kernel(PxStats p, double alpha) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.np) return;
  const int m = p.m[i];
  if (m <= 0) return;
  const double n0 = p.n[i], n1 = n0 + alpha * m;
  const double k = n1 / (n0 + m);
  p.n[i] = n1;
  p.r2[i] *= k;
  for (int c = 0; c < 3; c++) {
    p.tau[c * p.np + i] = (p.tau[c * p.np + i] + p.phi[c * p.np + i]) * k;
    p.phi[c * p.np + i] = 0;
  }
  p.m[i] = 0;
}
__global__ void shared_resolve_kernel(PxStats p, double photons, double* film) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.np) return;
  const double sc = photons > 0 ? 1.0 / SPPM_PI / (p.r2[i] * photons) : 0.0;
  for (int c = 0; c < 3; c++) film[i * 3 + c] = p.tau[c * p.np + i] * sc;
}
}  // namespace

}  // namespace sppm
