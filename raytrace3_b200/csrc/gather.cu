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
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <cstdlib>

#include "ctx.h"
#include "gather_kernels.cuh"

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

inline PxStats px_view(sppm_ctx* c) {
  PxStats p;
  p.np = (long long)c->px_w * c->px_h;
  double* d = (double*)c->px_store.p;
  p.r2 = d;
  p.n = d + p.np;
  p.tau = d + 2 * p.np;
  p.phi = d + 5 * p.np;
  p.m = (int32_t*)(d + 8 * p.np);
  return p;
}

inline int bits_for(uint64_t max_value) {
  int b = 1;
  while (b < 32 && (max_value >> b)) b++;
  return b;
}

cudaError_t layout_soa(sppm_ctx* c, int64_t n) {
  const size_t nd = (size_t)(n > 0 ? n : 1);
  const size_t bytes = nd * (10 * 8 + 6 * 8 + 3 * 4 + 2 * 4 + 8) + nd * (3 * 8 + 4) + 4096;  // + the second accumulator set
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
  c->dtao_set[0] = s.dtao;
  c->dN_set[0] = s.dN;
  c->dN_set[1] = takeI(nd);
  p = (unsigned char*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
  c->dtao_set[1] = (double*)p;
  c->acc_cur = 0;
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
    c->n_units = 0;
    c->grid = Grid{{0, 0, 0}, 1.0, {1, 1, 1}, 1, {-1e300, -1e300, -1e300}, {1e300, 1e300, 1e300}};
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
    double keep_lo[3], keep_hi[3];
    for (int k = 0; k < 3; k++) {  // the hit points' own bounds (true coordinates: capping the gridded region does not change them)
      keep_lo[k] = lo[k] - cell;
      keep_hi[k] = hi[k] + cell;
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
    for (int k = 0; k < 3; k++) {
      g.keep_lo[k] = keep_lo[k];
      g.keep_hi[k] = keep_hi[k];
    }
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
    // okey = pixel << 20 | eye-path node: only the bits a film of this size can set are sorted (5 radix passes
    // instead of 8 at 1024 x 768)
    const int okey_bits = std::min(64, 20 + bits_for((uint64_t)std::max<int64_t>((int64_t)c->film_w * c->film_h, 2) - 1));
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, o1, o2, i2, i1, (int)n, 0, okey_bits, st));
    CK(dev_reserve(c->cub_tmp, tmp));
    CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, o1, o2, i2, i1, (int)n, 0, okey_bits, st));
    c->host_counters.launches += 2;
  }
  uint64_t *m1 = (uint64_t*)c->dep_key.p, *m2 = (uint64_t*)c->dep_key2.p;  // (cell, Morton code in the cell) keys
  int morton = 1;  // SPPM_HP_MORTON=0: film order inside a cell, as round 1 (A/B knob)
  if (const char* e = getenv("SPPM_HP_MORTON")) morton = atoi(e);
  hp_key_kernel<<<nblk(n), TPB, 0, st>>>(raw, i1, n, g, m1, morton);
  c->host_counters.launches++;
  const int end_bit = bits_for((uint64_t)g.n_cells - 1) + HP_MORTON_BITS;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, m1, m2, i1, i2, (int)n, 0, end_bit, st));
  CK(dev_reserve(c->cub_tmp, tmp));
  CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, m1, m2, i1, i2, (int)n, 0, end_bit, st));
  hp_cell_of_key_kernel<<<nblk(n), TPB, 0, st>>>(m2, n, k1);  // k1 aliases m1, which the sort has consumed
  c->host_counters.launches++;
  k2 = k1;
  hp_permute_kernel<<<nblk(n), TPB, 0, st>>>(raw, i2, n, c->soa);
  {  // canonical rank -> position in the sorted map (resolve_kernel sums a pixel's hit points in that order);
     // i1 = raw index by rank is still intact (the sort reads it), the upper half of the key buffer is free
    CK(dev_reserve(c->hp_canon, (size_t)n * 4));
    uint32_t* rank_of_raw = (uint32_t*)c->dep_key.p + n;
    hp_rank_kernel<<<nblk(n), TPB, 0, st>>>(i1, n, rank_of_raw);
    hp_canon_kernel<<<nblk(n), TPB, 0, st>>>(i2, rank_of_raw, n, (uint32_t*)c->hp_canon.p);
    c->host_counters.launches += 2;
  }
  CK(cudaMemsetAsync(c->dN_set[1], 0, (size_t)n * 4, st));
  CK(cudaMemsetAsync(c->dtao_set[1], 0, (size_t)n * 24, st));
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
  // 6. work units of the filtered gather: runs of x-adjacent occupied cells, cut every UNIT_HP hit points
  {
    int32_t* flag = (int32_t*)c->dep_idx.p;        // [n_occ] (free again)
    int32_t* run_first = (int32_t*)c->dep_idx2.p;  // [n_runs]
    int32_t* ucnt = (int32_t*)c->dep_key.p;        // [n_runs]
    int32_t* uoff = (int32_t*)c->dep_key2.p;       // [n_runs + 1]
    int32_t* d_nr = d_nruns + 1;
    run_flag_kernel<<<nblk(n_occ), TPB, 0, st>>>((const uint32_t*)c->occ_cell.p, n_occ, g.dim[0], flag);
    thrust::counting_iterator<int32_t> iota(0);
    CK(cub::DeviceSelect::Flagged(nullptr, tmp, iota, flag, run_first, d_nr, n_occ, st));
    CK(dev_reserve(c->cub_tmp, tmp));
    CK(cub::DeviceSelect::Flagged(c->cub_tmp.p, tmp, iota, flag, run_first, d_nr, n_occ, st));
    CK(cudaMemsetAsync(ucnt, 0, ((size_t)n_occ + 1) * 4, st));
    unit_count_kernel<<<nblk(n_occ), TPB, 0, st>>>(run_first, d_nr, n_occ, (const int32_t*)c->occ_start.p, ucnt);
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, ucnt, uoff, n_occ + 1, st));
    CK(dev_reserve(c->cub_tmp, tmp));
    CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, ucnt, uoff, n_occ + 1, st));
    int32_t n_units = 0;
    CK(cudaMemcpyAsync(&n_units, uoff + n_occ, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    c->n_units = n_units;
    CK(dev_reserve(c->units, (size_t)(n_units > 0 ? n_units : 1) * sizeof(GUnit)));
    if (n_units > 0)
      unit_fill_kernel<<<nblk(n_units), TPB, 0, st>>>(run_first, d_nr, n_occ, (const int32_t*)c->occ_start.p,
                                                      (const uint32_t*)c->occ_cell.p, uoff, n_units, g,
                                                      (const int32_t*)c->cell_slot.p, (GUnit*)c->units.p);
    c->host_counters.launches += 5;
  }
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
  CK(L.prof.mark(PK_BIN_SCAN, st));
  CK(cub::DeviceScan::ExclusiveSum(L.cub_tmp.p, tmp, (const int32_t*)L.slot_count.p, (int32_t*)L.dep_start.p, n_scan, st));
  CK(L.prof.mark(PK_BIN_SCATTER, st));
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
    a.grid = c->grid;
    a.occ_cell = (const uint32_t*)c->occ_cell.p;
    a.units = (const GUnit*)c->units.p;
    a.n_units = c->n_units;
    a.full_accept = 1;
    if (const char* e = getenv("SPPM_GATHER_FULL")) a.full_accept = atoi(e);
    const bool filtered = c->prm.gather_mode == SPPM_GATHER_FILTERED;
    auto kern = !filtered ? gather_kernel : (c->prm.count_pairs ? gather_filtered_kernel<true> : gather_filtered_kernel<false>);
    const size_t tile_bytes = filtered ? F_SMEM : 2 * (size_t)G_TILE * sizeof(Deposit);
    const int threads = filtered ? W_THREADS : G_THREADS;
    int& cached = c->gather_per_sm[!filtered ? 0 : (c->prm.count_pairs ? 2 : 1)];
    if (cached <= 0) {  // once per context and kernel variant
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes));  // per device
      int q = 1;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, threads, tile_bytes));
      cached = q < 1 ? 1 : q;
    }
    int per_sm = cached;
    if (const char* e = getenv("SPPM_GATHER_CTAS")) per_sm = std::max(1, std::min(per_sm, atoi(e)));  // tuning knob
    unsigned blocks = (unsigned)c->sm_count * (unsigned)(per_sm < 1 ? 1 : per_sm);  // persistent CTAs, cells handed out dynamically
    const unsigned max_blocks = filtered ? (unsigned)((c->n_units + W_WARPS - 1) / W_WARPS) : (unsigned)c->n_occ;
    if (blocks > max_blocks) blocks = max_blocks;
    if (filtered) {
      // work items of this batch (see unit_parts_kernel); the item count stays on the device
      const int nu = c->n_units;
      CK(dev_reserve(c->item_parts, ((size_t)nu + 1) * 4));
      CK(dev_reserve(c->item_off, ((size_t)nu + 1) * 4));
      CK(c->prof.mark(PK_GATHER_PREP, st));
      unit_parts_kernel<<<nblk(nu + 1), TPB, 0, st>>>(a.units, nu, a.dep_off, (int32_t*)c->item_parts.p);
      size_t tmp = 0;
      CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, (int32_t*)c->item_parts.p, (int32_t*)c->item_off.p, nu + 1, st));
      CK(dev_reserve(c->cub_tmp, tmp));
      CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, (int32_t*)c->item_parts.p, (int32_t*)c->item_off.p, nu + 1, st));
      a.item_off = (const int32_t*)c->item_off.p;
      c->host_counters.launches += 2;
    }
    CK(c->prof.mark(PK_GATHER_KERNEL, st));
    kern<<<blocks, threads, tile_bytes, st>>>(a, work_counter);
    c->host_counters.launches++;
    CK(c->prof.mark(PK_END, st));
  }
  if (ev_gather_end) CK(cudaEventRecord(ev_gather_end, st));
  return cudaGetLastError();
}

cudaError_t launch_update(sppm_ctx* c, cudaStream_t st) {
  if (c->n_hp == 0) return cudaSuccess;
  update_kernel<<<nblk(c->n_hp), TPB, 0, st>>>(c->soa, c->n_hp, c->prm.update_mode, c->prm.alpha);
  c->host_counters.launches++;
  return cudaGetLastError();
}

cudaError_t launch_resolve(sppm_ctx* c, int cnt, int onetime) {
  const size_t bytes = (size_t)c->film_w * c->film_h * 3 * 8;
  CK(dev_reserve(c->film, bytes));
  CK(cudaMemsetAsync(c->film.p, 0, bytes, c->stream));
  if (c->n_hp == 0) return cudaSuccess;
  resolve_kernel<<<nblk(c->n_hp), TPB, 0, c->stream>>>(c->soa, (const uint32_t*)c->hp_canon.p, c->n_hp, cnt, onetime, c->film_w, (double*)c->film.p);
  c->host_counters.launches++;
  return cudaGetLastError();
}

cudaError_t shared_begin(sppm_ctx* c, int w, int h) {
  c->px_w = w;
  c->px_h = h;
  c->px_photons = 0;
  const size_t np = (size_t)w * h;
  CK(dev_reserve(c->px_store, np * (8 * 8 + 4) + 64));
  shared_init_kernel<<<nblk((long long)np), TPB, 0, c->stream>>>(px_view(c), c->prm.init_r2);
  c->host_counters.launches++;
  return cudaGetLastError();
}
cudaError_t shared_load(sppm_ctx* c) {
  if (c->n_hp == 0) return cudaSuccess;
  shared_load_kernel<<<nblk(c->n_hp), TPB, 0, c->stream>>>(c->soa, c->n_hp, px_view(c), c->px_w);
  c->host_counters.launches++;
  return cudaGetLastError();
}
cudaError_t shared_fold(sppm_ctx* c) {
  if (c->n_hp > 0) shared_accum_kernel<<<nblk(c->n_hp), TPB, 0, c->stream>>>(c->soa, c->n_hp, px_view(c), c->px_w);
  shared_update_kernel<<<nblk((long long)c->px_w * c->px_h), TPB, 0, c->stream>>>(px_view(c), c->prm.alpha);
  c->host_counters.launches += 2;
  return cudaGetLastError();
}
cudaError_t shared_resolve(sppm_ctx* c) {
  const size_t bytes = (size_t)c->px_w * c->px_h * 3 * 8;
  CK(dev_reserve(c->film, bytes));
  shared_resolve_kernel<<<nblk((long long)c->px_w * c->px_h), TPB, 0, c->stream>>>(px_view(c), (double)c->px_photons, (double*)c->film.p);
  c->host_counters.launches++;
  return cudaGetLastError();
}
cudaError_t shared_export(sppm_ctx* c, double* r2, double* n, double* tau3) {
  const PxStats p = px_view(c);
  const size_t np = (size_t)p.np;
  if (r2) CK(cudaMemcpyAsync(r2, p.r2, np * 8, cudaMemcpyDeviceToHost, c->stream));
  if (n) CK(cudaMemcpyAsync(n, p.n, np * 8, cudaMemcpyDeviceToHost, c->stream));
  if (tau3) CK(cudaMemcpyAsync(tau3, p.tau, 3 * np * 8, cudaMemcpyDeviceToHost, c->stream));  // [3][h][w] planes
  return cudaStreamSynchronize(c->stream);
}

cudaError_t shared_import(sppm_ctx* c, const double* r2, const double* n, const double* tau3) {
  const PxStats p = px_view(c);
  const size_t np = (size_t)p.np;
  CK(cudaMemcpyAsync(p.r2, r2, np * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(p.n, n, np * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(p.tau, tau3, 3 * np * 8, cudaMemcpyHostToDevice, c->stream));
  return cudaStreamSynchronize(c->stream);
}

}  // namespace sppm
