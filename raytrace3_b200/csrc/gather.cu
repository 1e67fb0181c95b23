// gather.cu — the hit-point hash grid and the photon gather for sm_100a.
//
// Replaces flann::KDTreeSingleIndex + radiusSearch + the accumulate loop of
// RayTracer::PhotonTrace (Raytracer.h:92-96, 139-159, 370-380) by
//   build_index:     cell key per hit point -> CUB radix sort -> SoA permute -> occupied-cell
//                    list, neighbourhood slot table (once per eye pass)
//   bin_and_gather:  deposits of one photon batch binned by cell with a counting sort (slot
//                    counts + ranks taken at append time, one scan, one scatter), then a
//                    hit-point-centric gather: each hit point scans the deposits of
//                    its 27 neighbour cells and accumulates newN / tao in registers — no
//                    atomics, each accumulator is written once per batch.
// and HitPoint::update / the film resolve (Raytracer.h:69-79, 281-293).
//
// Bit-exactness: the acceptance predicate is evaluated in FP64 with the reference's
// operation order and no FMA (the library is built with -fmad=false), so newN is exact.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>

#include "ctx.h"

namespace sppm {

cudaError_t dev_reserve(DevBuf& b, size_t bytes) {
  if (bytes <= b.bytes && b.p) return cudaSuccess;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
  size_t want = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e == cudaSuccess) b.bytes = want;
  return e;
}
void dev_free(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
}

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
// cell key of the hit points taken in canonical (okey) order
__global__ void hp_key_kernel(const HpRaw* hp, const uint32_t* idx, long long n, Grid g, uint32_t* key) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const HpRaw& r = hp[idx[i]];
  key[i] = (uint32_t)grid_cell(g, r.pos[0], r.pos[1], r.pos[2]);
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

// ---- deposits --------------------------------------------------------------------------
__global__ void dep_key_kernel(const Deposit* dep, long long m, Grid g, const int32_t* cell_slot, int32_t n_slots,
                               uint32_t* key, uint32_t* rank, unsigned int* slot_count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int32_t s = cell_slot[grid_cell(g, dep[i].x[0], dep[i].x[1], dep[i].x[2])];
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
};

// ---- TMA (cp.async.bulk) + mbarrier helpers -----------------------------------------------
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
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
__global__ void resolve_kernel(HpSoA s, long long n, int cnt, int onetime, int w, double* film) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double sc = 1.0 / SPPM_PI / (s.r2[i] * cnt * onetime);
  double* px = film + ((size_t)s.row[i] * w + s.col[i]) * 3;
  atomicAdd(px + 0, s.tao[i] * sc);
  atomicAdd(px + 1, s.tao[n + i] * sc);
  atomicAdd(px + 2, s.tao[2 * n + i] * sc);
}

inline int bits_for(uint64_t max_value) {
  int b = 1;
  while (b < 32 && (max_value >> b)) b++;
  return b;
}

cudaError_t layout_soa(sppm_ctx* c, int64_t n) {
  const size_t nd = (size_t)(n > 0 ? n : 1);
  const size_t bytes = nd * (10 * 8 + 6 * 8 + 3 * 4 + 2 * 4 + 8) + 4096;
  CK(dev_reserve(c->soa_store, bytes));
  unsigned char* p = (unsigned char*)c->soa_store.p;
  auto takeD = [&](size_t k) { double* r = (double*)p; p += k * 8; return r; };
  HpSoA& s = c->soa;
  s.px = takeD(nd); s.py = takeD(nd); s.pz = takeD(nd);
  s.nx = takeD(nd); s.ny = takeD(nd); s.nz = takeD(nd);
  s.r2 = takeD(nd);
  s.wx = takeD(nd); s.wy = takeD(nd); s.wz = takeD(nd);
  s.tao = takeD(3 * nd);
  s.dtao = takeD(3 * nd);
  s.okey = (uint64_t*)p; p += nd * 8;
  auto takeI = [&](size_t k) { int32_t* r = (int32_t*)p; p += k * 4; return r; };
  s.N = takeI(nd); s.newN = takeI(nd); s.dN = takeI(nd); s.row = takeI(nd); s.col = takeI(nd);
  return cudaSuccess;
}

}  // namespace

cudaError_t build_index(sppm_ctx* c) {
  cudaStream_t st = c->stream;
  const int64_t n = c->n_hp;
  c->index_ok = false;
  CK(layout_soa(c, n));
  if (n == 0) {
    c->n_occ = 0;
    c->n_slots = 0;
    c->grid = Grid{{0, 0, 0}, 1.0, {1, 1, 1}, 1};
    CK(dev_reserve(c->cell_slot, 4));
    CK(cudaMemsetAsync(c->cell_slot.p, 0xff, 4, st));
    CK(dev_reserve(c->occ_start, 8));
    CK(dev_reserve(c->occ_cell, 4));
    CK(dev_reserve(c->occ_nbr, 9 * 8));
    c->index_ok = true;
    return cudaSuccess;
  }
  const HpRaw* raw = (const HpRaw*)c->hp_raw.p;
  // 1. bounds + centroid
  CK(dev_reserve(c->scratch, 256));
  {
    unsigned long long init[9];
    for (int k = 0; k < 3; k++) {
      init[k] = ~0ull;
      init[3 + k] = 0ull;
      init[6 + k] = 0ull;  // double 0.0
    }
    CK(cudaMemcpyAsync(c->scratch.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    bounds_kernel<<<min(nblk(n), 1184u), TPB, 0, st>>>(raw, n, (unsigned long long*)c->scratch.p,
                                                        (double*)((unsigned long long*)c->scratch.p + 6));
    c->host_counters.launches++;
    CK(cudaMemcpyAsync(init, c->scratch.p, sizeof(init), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // 2. grid: cell edge = sqrt(searchR2) (1+1e-9) so a radius-sqrt(searchR2) ball spans <= 3 cells
    // per axis; coordinates outside the gridded region are clamped into its boundary cells, which
    // keeps neighbours neighbours.  The region is capped at 2^26 cells around the centroid.
    const double cell = sqrt(c->prm.init_r2) * (1.0 + 1e-9);
    double lo[3], hi[3], mean[3];
    int dim[3];
    for (int k = 0; k < 3; k++) {
      lo[k] = key_dbl(init[k]);
      hi[k] = key_dbl(init[3 + k]);
      double s;
      memcpy(&s, &init[6 + k], 8);
      mean[k] = s / (double)n;
      double cells = floor((hi[k] - lo[k]) / cell) + 1;
      dim[k] = cells > 1e9 ? 1000000000 : (int)cells;
    }
    auto shrink = [&](int k, int nd) {
      double half = 0.5 * nd * cell;
      double ctr = mean[k];
      if (ctr - half < lo[k]) ctr = lo[k] + half;
      if (ctr + half > hi[k]) ctr = hi[k] - half;
      lo[k] = ctr - half;
      dim[k] = nd;
    };
    for (int k = 0; k < 3; k++)
      if (dim[k] > 2048) shrink(k, 2048);
    while ((int64_t)dim[0] * dim[1] * dim[2] > ((int64_t)1 << 26)) {
      int k = 0;
      for (int j = 1; j < 3; j++)
        if (dim[j] > dim[k]) k = j;
      shrink(k, dim[k] / 2);
    }
    Grid g;
    for (int k = 0; k < 3; k++) {
      g.org[k] = lo[k];
      g.dim[k] = dim[k];
    }
    g.inv_cell = 1.0 / cell;
    g.n_cells = (int64_t)dim[0] * dim[1] * dim[2];
    c->grid = g;
  }
  const Grid g = c->grid;
  // 3. canonical order first (radix sort by okey = pixel, eye-path node: the eye pass appends in
  // arrival order), then a stable radix sort by cell key, then the SoA permute.  Every rank of a
  // multi-GPU run therefore holds the hit points in the same order.
  CK(dev_reserve(c->dep_key, (size_t)n * 8));  // reuse the deposit key/index buffers as sort scratch
  CK(dev_reserve(c->dep_key2, (size_t)n * 8));
  CK(dev_reserve(c->dep_idx, (size_t)n * 4));
  CK(dev_reserve(c->dep_idx2, (size_t)n * 4));
  uint32_t *k1 = (uint32_t*)c->dep_key.p, *k2 = (uint32_t*)c->dep_key2.p, *i1 = (uint32_t*)c->dep_idx.p,
           *i2 = (uint32_t*)c->dep_idx2.p;
  size_t tmp = 0;
  {
    uint64_t *o1 = (uint64_t*)c->dep_key.p, *o2 = (uint64_t*)c->dep_key2.p;
    hp_okey_kernel<<<nblk(n), TPB, 0, st>>>(raw, n, o1, i2);
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, o1, o2, i2, i1, (int)n, 0, 64, st));
    CK(dev_reserve(c->cub_tmp, tmp));
    CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, o1, o2, i2, i1, (int)n, 0, 64, st));
    c->host_counters.launches += 2;
  }
  hp_key_kernel<<<nblk(n), TPB, 0, st>>>(raw, i1, n, g, k1);
  c->host_counters.launches++;
  const int end_bit = bits_for((uint64_t)g.n_cells - 1);
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k1, k2, i1, i2, (int)n, 0, end_bit, st));
  CK(dev_reserve(c->cub_tmp, tmp));
  CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, k1, k2, i1, i2, (int)n, 0, end_bit, st));
  hp_permute_kernel<<<nblk(n), TPB, 0, st>>>(raw, i2, n, c->soa);
  c->host_counters.launches += 3;
  // 4. occupied cells (run-length encode of the sorted keys) and their hit-point ranges
  CK(dev_reserve(c->occ_cell, (size_t)n * 4));
  CK(dev_reserve(c->occ_start, ((size_t)n + 1) * 4));
  int32_t* counts = (int32_t*)c->dep_idx.p;  // i1 is free now
  int32_t* d_nruns = (int32_t*)c->scratch.p;
  CK(cub::DeviceRunLengthEncode::Encode(nullptr, tmp, k2, (uint32_t*)c->occ_cell.p, counts, d_nruns, (int)n, st));
  CK(dev_reserve(c->cub_tmp, tmp));
  CK(cub::DeviceRunLengthEncode::Encode(c->cub_tmp.p, tmp, k2, (uint32_t*)c->occ_cell.p, counts, d_nruns, (int)n, st));
  int32_t n_occ = 0;
  CK(cudaMemcpyAsync(&n_occ, d_nruns, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  c->n_occ = n_occ;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, counts, (int32_t*)c->occ_start.p, n_occ, st));
  CK(dev_reserve(c->cub_tmp, tmp));
  CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, counts, (int32_t*)c->occ_start.p, n_occ, st));
  set_last_kernel<<<1, 1, 0, st>>>((int32_t*)c->occ_start.p, n_occ, (int32_t)n);
  c->host_counters.launches += 4;
  // 5. cells whose 27-neighbourhood holds a hit point get a compact slot (others: -1)
  CK(dev_reserve(c->cell_slot, (size_t)g.n_cells * 4));
  CK(dev_reserve(c->scan_scratch, (size_t)g.n_cells * 4));
  CK(cudaMemsetAsync(c->cell_slot.p, 0, (size_t)g.n_cells * 4, st));
  mark_nbr_kernel<<<nblk((long long)n_occ * 27), TPB, 0, st>>>((uint32_t*)c->occ_cell.p, n_occ, g, (int32_t*)c->cell_slot.p);
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, (int32_t*)c->cell_slot.p, (int32_t*)c->scan_scratch.p, (int)g.n_cells, st));
  CK(dev_reserve(c->cub_tmp, tmp));
  CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, (int32_t*)c->cell_slot.p, (int32_t*)c->scan_scratch.p, (int)g.n_cells, st));
  finalize_slot_kernel<<<nblk(g.n_cells), TPB, 0, st>>>((int32_t*)c->cell_slot.p, (int32_t*)c->scan_scratch.p, g.n_cells, d_nruns);
  int32_t n_slots = 0;
  CK(cudaMemcpyAsync(&n_slots, d_nruns, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  c->n_slots = n_slots;
  CK(dev_reserve(c->occ_nbr, (size_t)n_occ * 9 * 8));
  occ_run_kernel<<<nblk((long long)n_occ * 9), TPB, 0, st>>>((uint32_t*)c->occ_cell.p, n_occ, g, (int32_t*)c->cell_slot.p,
                                                             (int2*)c->occ_nbr.p);
  c->host_counters.launches += 4;
  CK(cudaGetLastError());
  c->host_counters.hit_points = (uint64_t)n;
  c->host_counters.grid_cells = (uint64_t)g.n_cells;
  c->index_ok = true;
  return cudaSuccess;
}

cudaError_t key_host_deposits(sppm_ctx* c, Lane& L, int64_t m, cudaStream_t st) {
  if (m <= 0) return cudaSuccess;
  dep_key_kernel<<<nblk(m), TPB, 0, st>>>((const Deposit*)L.dep_raw.p, m, c->grid, (const int32_t*)c->cell_slot.p, c->n_slots,
                                          (uint32_t*)L.dep_key.p, (uint32_t*)L.dep_rank.p, (unsigned int*)L.slot_count.p);
  c->host_counters.launches++;
  return cudaGetLastError();
}

// Bin the lane's raw deposits by slot (counting sort: per-slot counts and ranks were taken when
// the deposits were appended): one scan + one scatter on stream st.  The deposit count is read
// on the device, nothing here waits for the host.
cudaError_t bin_deposits(sppm_ctx* c, Lane& L, cudaStream_t st) {
  if (c->n_hp == 0 || L.dep_cap <= 0) return cudaSuccess;
  CK(dev_reserve(L.dep_sorted, (size_t)L.dep_cap * sizeof(Deposit)));
  size_t tmp = 0;
  const int n_scan = c->n_slots + 2;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const int32_t*)L.slot_count.p, (int32_t*)L.dep_start.p, n_scan, st));
  CK(dev_reserve(L.cub_tmp, tmp));
  CK(cub::DeviceScan::ExclusiveSum(L.cub_tmp.p, tmp, (const int32_t*)L.slot_count.p, (int32_t*)L.dep_start.p, n_scan, st));
  dep_scatter_kernel<<<(unsigned)c->sm_count * 6, 320, 0, st>>>((const Deposit*)L.dep_raw.p, (const uint32_t*)L.dep_key.p,
                                                                (const uint32_t*)L.dep_rank.p,
                                                                (const unsigned long long*)L.scalars.p + 1,
                                                                (unsigned long long)L.dep_cap, (const int32_t*)L.dep_start.p,
                                                                (Deposit*)L.dep_sorted.p);
  c->host_counters.launches += 2;
  return cudaGetLastError();
}

// The gather of the lane's binned deposits into the per-round accumulators dN / dtao (main stream).
cudaError_t launch_gather(sppm_ctx* c, Lane& L, cudaEvent_t ev_gather_start, cudaEvent_t ev_gather_end) {
  cudaStream_t st = c->stream;
  if (ev_gather_start) CK(cudaEventRecord(ev_gather_start, st));
  if (c->n_hp > 0 && L.dep_cap > 0 && c->n_occ > 0) {
    GatherArgs a;
    a.s = c->soa;
    a.n_hp = c->n_hp;
    a.occ_start = (const int32_t*)c->occ_start.p;
    a.occ_run = (const int2*)c->occ_nbr.p;
    a.n_occ = c->n_occ;
    a.dep_off = (const int32_t*)L.dep_start.p;
    a.dep = (const Deposit*)L.dep_sorted.p;
    a.search_r2 = c->prm.init_r2;
    a.counters = (DevCounters*)c->counters.p;
    unsigned int* work_counter = (unsigned int*)((unsigned long long*)c->dev_scalars.p + 4);
    CK(cudaMemsetAsync(work_counter, 0, 4, st));
    const size_t tile_bytes = 2 * (size_t)G_TILE * sizeof(Deposit);
    CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes));  // per device
    int per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gather_kernel, G_THREADS, tile_bytes));
    if (const char* e = getenv("SPPM_GATHER_CTAS")) per_sm = std::max(1, std::min(per_sm, atoi(e)));  // tuning knob
    unsigned blocks = (unsigned)c->sm_count * (unsigned)(per_sm < 1 ? 1 : per_sm);  // persistent CTAs, cells handed out dynamically
    if (blocks > (unsigned)c->n_occ) blocks = (unsigned)c->n_occ;
    gather_kernel<<<blocks, G_THREADS, tile_bytes, st>>>(a, work_counter);
    c->host_counters.launches++;
  }
  if (ev_gather_end) CK(cudaEventRecord(ev_gather_end, st));
  return cudaGetLastError();
}

cudaError_t launch_update(sppm_ctx* c) {
  if (c->n_hp == 0) return cudaSuccess;
  update_kernel<<<nblk(c->n_hp), TPB, 0, c->stream>>>(c->soa, c->n_hp, c->prm.update_mode, c->prm.alpha);
  c->host_counters.launches++;
  return cudaGetLastError();
}

cudaError_t launch_resolve(sppm_ctx* c, int cnt, int onetime) {
  const size_t bytes = (size_t)c->film_w * c->film_h * 3 * 8;
  CK(dev_reserve(c->film, bytes));
  CK(cudaMemsetAsync(c->film.p, 0, bytes, c->stream));
  if (c->n_hp == 0) return cudaSuccess;
  resolve_kernel<<<nblk(c->n_hp), TPB, 0, c->stream>>>(c->soa, c->n_hp, cnt, onetime, c->film_w, (double*)c->film.p);
  c->host_counters.launches++;
  return cudaGetLastError();
}

}  // namespace sppm
