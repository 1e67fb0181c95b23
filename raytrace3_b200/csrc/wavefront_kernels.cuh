// wavefront_kernels.cuh — the kernels of wavefront.cu (photon_step_kernel, newton_items_kernel), the
// queue layout and the per-bounce launch arguments; the launch loop and context plumbing stay in
// wavefront.cu.  A header of their own so that tests/device_host can compile these very kernel bodies
// for the CPU (one-warp blocks on fibers) and hold the bounce-synchronous pass to the oracle without a
// GPU; wavefront.cu is the only product file that includes it.
#pragma once
#ifndef SPPM_HOST_EMULATION  // tests/device_host supplies a one-lane coalesced_group instead
#include <cooperative_groups.h>
#endif

#include <algorithm>
#include <cstdlib>

#include "ctx.h"

#ifndef SPPM_STEP_MINBLOCKS
#define SPPM_STEP_MINBLOCKS 6    // CTAs/SM the step kernel's register budget is sized for (80 regs; measured best)
#endif
#ifndef SPPM_NEWTON_MINBLOCKS
#define SPPM_NEWTON_MINBLOCKS 2  // same for the Newton item kernel
#endif

namespace sppm {

namespace {

constexpr int STEP_THREADS = 128;
constexpr unsigned long long ID_MASK = 0x00ffffffffffffffull;

struct Queue {  // SoA photon state, one entry per live photon
  double *ox, *oy, *oz, *dx, *dy, *dz, *fx, *fy, *fz, *hdist;
  unsigned long long* idd;  // photon id | dep << 56
  int* hobj;                // analytic best of the pending query: obj | inside << 16, -1 = none
  unsigned *item_begin, *item_cnt;
};
struct Items {
  unsigned* slot;   // photon slot in the queue the items were generated for
  unsigned* patch;  // global patch | object << 16
  double *t, *u, *v;
};

// Photons that leave the wavefront because the (ray, patch) item list of their bounce is full: the state
// before the pending query.  photon_trace_kernel (trace_kernels.cuh, persistent warps with the Newton
// solver inline) finishes them from here on — same RNG addresses, so the same path.
struct Spill {
  double* st[9];             // org, dir, flux
  unsigned long long* idd;   // photon id | dep << 56
  unsigned* count;
};

struct StepArgs {
  DevScene scene;
  int patches_in_smem, max_dep, bounce, do_intersect, fast_math;
  uint64_t seed, pass;
  DevCounters* counters;
  Queue qin, qout;
  Items it_in, it_out;
  const unsigned* n_in_ptr;  // [4]: photons queued in the four regions (see photon_step_kernel)
  unsigned n_in_host;
  unsigned* n_out;           // [4]
  unsigned q_cap;
  unsigned* n_items_out;
  unsigned item_cap;
  Spill spill;
  int spill_all;          // tail bounces: every photon with Bezier solves pending goes to the spill list
  uint64_t begin, count;  // photon range per light of this batch
  Grid grid;
  const int32_t* cell_slot;
  int32_t n_slots;
  int keep_all;
  Deposit* dep;
  uint32_t* dep_key;
  uint32_t* dep_rank;
  unsigned int* slot_count;
  unsigned long long dep_cap;
  unsigned long long* dep_count;
};

__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
  return v;
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }

__global__ void __launch_bounds__(STEP_THREADS, SPPM_STEP_MINBLOCKS) photon_step_kernel(const StepArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  // The queue arrays hold 2 x q_cap entries and are filled from four ends, by what the pending hit
  // will make the next step do, so that warps shade uniform work:
  //   region 0 [0 ->)             a plane with a diffuse part (most photons: one coherent path)
  //   region 1 [<- q_cap)         a purely specular object (mirror / glass: short reflect / refract paths)
  //   region 2 [q_cap ->)         a sphere with a diffuse part
  //   region 3 [<- 2 q_cap)       Bezier solves pending (the item merge + patch evaluation path)
  // Photons that hit nothing are dropped when they are found.  The virtual index space of a step is
  // [region 0 | region 1 | region 2 | region 3], the first three padded to whole warps.
  unsigned nreg[4];
  nreg[0] = a.bounce == 0 ? a.n_in_host : a.n_in_ptr[0];
#pragma unroll
  for (int r = 1; r < 4; r++) nreg[r] = a.bounce == 0 ? 0u : a.n_in_ptr[r];
  const unsigned c0 = ((nreg[0] + 31) >> 5) * 32, c1 = c0 + ((nreg[1] + 31) >> 5) * 32, c2 = c1 + ((nreg[2] + 31) >> 5) * 32;
  const unsigned n_in = c2 + nreg[3];
#ifndef SPPM_NO_EARLY_EXIT
  // A CTA whose first warp starts beyond the queue has no photon at all (the grid is sized for a full batch; in
  // the tail bounces a few CTAs hold every live photon): it leaves before staging the scene.
  if (blockIdx.x * blockDim.x >= n_in) return;
#endif
  const SceneView sv = stage_scene(a.scene, smem, a.patches_in_smem, false);
  const unsigned warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
  unsigned queries = 0, solves = 0, photons = 0, deposits = 0, kept = 0, spilled = 0;  // per thread, < 2^32

  for (unsigned base = warp_global * 32; base < n_in; base += n_warps * 32) {
    const unsigned vi = base + lane;
    const int reg = base < c0 ? 0 : (base < c1 ? 1 : (base < c2 ? 2 : 3));  // warp-uniform
    const unsigned k = vi - (reg == 0 ? 0u : (reg == 1 ? c0 : (reg == 2 ? c1 : c2)));  // index within the region
    const bool active = k < nreg[reg];
    const unsigned i = reg == 0 ? k : (reg == 1 ? a.q_cap - 1 - k : (reg == 2 ? a.q_cap + k : 2 * a.q_cap - 1 - k));
    bool alive = false;
    unsigned rank_val = 0;
    unsigned long long rank_at = ~0ull;  // deposit whose rank-in-slot is still in flight
    d3 org = mk(0, 0, 0), dir = mk(0, 0, 1), flux = mk(0, 0, 0);
    uint64_t id = 0;
    int dep = 0;
    if (active) {
      if (a.bounce == 0) {  // Light::emit (Light.h:8-13)
        const unsigned pl = (unsigned)(i / a.count);
        id = ((uint64_t)pl << 40) | (a.begin + i % a.count);
        const SppmLight& L = a.scene.lights[pl];
        flux = ld3(L.color) * SPPM_PI * 4;
        dir = unit_rand_vec(a.seed, a.pass, id, P_EMIT);
        org = ld3(L.pos);
        alive = true;
        photons++;
      } else {
        // ---- finish the pending query: nearest of the analytic best and the Bezier solves
        const unsigned long long idd = a.qin.idd[i];
        id = idd & ID_MASK;
        dep = (int)(idd >> 56);
        org = mk(a.qin.ox[i], a.qin.oy[i], a.qin.oz[i]);
        dir = mk(a.qin.dx[i], a.qin.dy[i], a.qin.dz[i]);
        Hit h;
        h.dist = a.qin.hdist[i];
        const int ho = a.qin.hobj[i];
        h.obj = ho < 0 ? -1 : (ho & 0xffff);
        h.inside = ho >= 0 && ((ho >> 16) & 1);
        h.patch = -1;
        h.u = h.v = 0;
        const unsigned ic = a.qin.item_cnt[i];
        if (ic) {  // Bezier3Obj::GetIntersect filter, per object in patch order (Bezier.h:246-261)
          const unsigned ib = a.qin.item_begin[i];
          int cur = -1, index = -1;
          double pt = SPPM_MAX_NUM, pu = SPPM_MAX_NUM, pv = SPPM_MAX_NUM;
          for (unsigned k = 0; k <= ic; k++) {
            const unsigned pk = k < ic ? a.it_in.patch[ib + k] : 0xffffffffu;
            const int ob = k < ic ? (int)(pk >> 16) : -2;
            if (ob != cur) {
              if (cur >= 0 && index != -1 && (pt < h.dist || (pt == h.dist && cur < h.obj))) {
                h.dist = pt; h.obj = cur; h.patch = index; h.u = pu; h.v = pv; h.inside = false;
              }
              cur = ob;
              index = -1;
              pt = pu = pv = SPPM_MAX_NUM;
            }
            if (k < ic) {
              const double qt = a.it_in.t[ib + k], qu = a.it_in.u[ib + k], qv = a.it_in.v[ib + k];
              if (!(qt < SPPM_M_EPS || qt > pt || qu < -SPPM_M_EPS || qu > 1 + SPPM_M_EPS || qv < -SPPM_M_EPS ||
                    qv > 1 + SPPM_M_EPS)) {
                if (qt < pt) {
                  pt = qt; pu = qu; pv = qv;
                  index = (int)(pk & 0xffffu);
                }
              }
            }
          }
        }
        if (!(h.obj < 0 || dep > a.max_dep)) {  // Raytracer.h:127-133
          flux = mk(a.qin.fx[i], a.qin.fy[i], a.qin.fz[i]);
          const int node = dep;
          dep++;
          const DevObject& ob = sv.obj[h.obj];
          const Surface s = shade_hit(sv, h, org, dir);
          // Raytracer.h:136-160: the gather input (x, N, incoming flux).  The slot lookup is issued
          // before the Roulette / direction arithmetic and consumed after it, and the rank atomic's
          // result is stored at the end of the iteration, so neither latency is waited for here.
          int32_t slot = -1;
          if (ob.isDiff && a.cell_slot && grid_keeps(a.grid, s.pos.x, s.pos.y, s.pos.z))
            slot = __ldg(a.cell_slot + grid_cell(a.grid, s.pos.x, s.pos.y, s.pos.z));
          // Obj::Roulette (Obj.h:30-45) and the single continuation (Raytracer.h:162-190)
          double ur, U1;
          draw2(a.seed, a.pass, id, (uint32_t)node, P_BOUNCE, 0, 0, ur, U1);
          const double r = randu(ur, 0., ob.allr);
          d3 ndir;
          if (ob.diffr > r) {
            double U2, unused;
            draw2(a.seed, a.pass, id, (uint32_t)node, P_BOUNCE, 1, 0, U2, unused);
            ndir = unit_rand_refl(s.n, U1, U2, a.fast_math != 0);
          } else if (ob.diffr + ob.reflr > r) {
            ndir = get_unit(get_refl(dir, s.n));
          } else {
            const double refrn = h.inside ? ob.refrn : 1 / ob.refrn;
            ndir = normalize3(refracted(dir, h.inside ? -s.n : s.n, refrn));
          }
          if (ob.isDiff) {  // Raytracer.h:136-160: the gather input (x, N, incoming flux)
            deposits++;
            if (slot >= 0 || a.keep_all) {
              // append with one atomic per group of lanes that are here together
              const cooperative_groups::coalesced_group g = cooperative_groups::coalesced_threads();
              unsigned long long dbase = 0;
              if (g.thread_rank() == 0) dbase = atomicAdd(a.dep_count, (unsigned long long)g.size());
              const unsigned long long ds = g.shfl(dbase, 0) + g.thread_rank();
              if (ds < a.dep_cap) {
                Deposit rec;
                rec.x[0] = s.pos.x; rec.x[1] = s.pos.y; rec.x[2] = s.pos.z;
                rec.n[0] = s.n.x; rec.n[1] = s.n.y; rec.n[2] = s.n.z;
                rec.flux[0] = flux.x; rec.flux[1] = flux.y; rec.flux[2] = flux.z;
                rec.meta = id | ((uint64_t)node << 56);
                a.dep[ds] = rec;
                const uint32_t key = slot >= 0 ? (uint32_t)slot : (uint32_t)a.n_slots;
                a.dep_key[ds] = key;
                rank_val = atomicAdd(a.slot_count + key, 1u);  // counting sort: rank within the slot
                rank_at = ds;
                kept++;
              }
            }
          }
          flux = mul3(s.color, flux);
          org = s.pos;
          dir = ndir;
          alive = true;
        }
      }
    }
    // ---- next query: analytic objects per lane, Bezier object boxes noted in a bit mask
    Hit best;
    best.dist = SPPM_MAX_NUM;
    best.obj = -1;
    best.inside = false;
    unsigned long long boxobj = 0;
    if (alive) {
      queries++;
      if (a.do_intersect) {
        for (int o = 0; o < sv.n_obj; o++) {
          const DevObject& ob = sv.obj[o];
          if (ob.kind == SPPM_OBJ_PLANE) test_plane(ob, o, org, dir, best);
          else if (ob.kind == SPPM_OBJ_SPHERE) test_sphere(ob, o, org, dir, best);
          else if (aabb_intersect(ob.box, org, dir)) boxobj |= 1ull << o;  // Bezier.h:244
        }
      } else {
        alive = false;  // the last query's result is never used (depth limit, Raytracer.h:129-132)
      }
    }
    if (rank_at != ~0ull) a.dep_rank[rank_at] = rank_val;  // the atomic issued before the Roulette has landed by now
    // ---- compaction of the survivors into the output queue (front: plane hits, back: the rest)
    if (alive && best.obj < 0 && boxobj == 0) alive = false;  // hit nothing (Raytracer.h:127): the photon is gone
    int oreg = 3;  // Bezier solves pending
    if (alive && boxobj == 0) {
      const DevObject& bo = sv.obj[best.obj];
      oreg = !bo.isDiff ? 1 : (bo.kind == SPPM_OBJ_PLANE ? 0 : 2);
    }
    const unsigned am = __ballot_sync(FULL, alive);
    if (!am) continue;
    unsigned rm[4];
#pragma unroll
    for (int r = 0; r < 4; r++) rm[r] = __ballot_sync(FULL, alive && oreg == r);
    unsigned rbase = 0;
    if (lane < 4 && rm[lane]) rbase = atomicAdd(a.n_out + lane, (unsigned)__popc(rm[lane]));  // one atomic per non-empty region
    rbase = __shfl_sync(FULL, rbase, oreg);
    const unsigned lt = (1u << lane) - 1;
    const unsigned kk = rbase + __popc(rm[oreg] & lt);
    const unsigned my_slot = oreg == 0 ? kk : (oreg == 1 ? a.q_cap - 1 - kk : (oreg == 2 ? a.q_cap + kk : 2 * a.q_cap - 1 - kk));
    // ---- cooperative Bezier culling: per in-box ray, 32 lanes test 32 patch boxes (Bezier.h:178)
    unsigned my_begin = 0, my_cnt = 0;
    bool spill_me = false;
    unsigned bm = __ballot_sync(FULL, alive && boxobj != 0);
    if (a.spill_all) {
      // Tail bounces (few photons alive, a handful of them inside a Bezier box): a Newton launch would cost its
      // serial chain (2 rounds x 10 iterations, ~35 us) for almost no work.  These photons leave the wavefront
      // here and the persistent tracer finishes them after the last bounce, all tails in one launch.
      spill_me = alive && boxobj != 0;
      bm = 0;
    }
    while (bm) {
      const int src = __ffs(bm) - 1;
      bm &= bm - 1;
      const d3 so = mk(shfl_d(org.x, src), shfl_d(org.y, src), shfl_d(org.z, src));
      const d3 sd = mk(shfl_d(dir.x, src), shfl_d(dir.y, src), shfl_d(dir.z, src));
      const unsigned long long omask = __shfl_sync(FULL, boxobj, src);
      const unsigned sslot = __shfl_sync(FULL, my_slot, src);
      unsigned total = 0, ibase = 0;
      for (int pass2 = 0; pass2 < 2; pass2++) {  // count, reserve one contiguous range, then write
        unsigned off = 0;
        unsigned long long om = omask;
        while (om) {
          const int o = __ffsll((long long)om) - 1;
          om &= om - 1;
          const DevObject& ob = sv.obj[o];
          for (int ch = 0; ch < ob.n_patch; ch += 32) {
            const int p = ch + lane;
            const bool hitp = p < ob.n_patch && aabb_intersect(sv.patchBox + (size_t)(ob.first_patch + p) * 6, so, sd);
            const unsigned b = __ballot_sync(FULL, hitp);
            if (pass2 && hitp) {
              const unsigned k = ibase + off + __popc(b & ((1u << lane) - 1));
              if (k < a.item_cap) {
                a.it_out.slot[k] = sslot;
                a.it_out.patch[k] = (unsigned)(ob.first_patch + p) | ((unsigned)o << 16);
              }
            }
            off += __popc(b);
          }
        }
        if (!pass2) {
          total = off;
          if (total == 0) break;
          if (lane == 0) ibase = atomicAdd(a.n_items_out, total);
          ibase = __shfl_sync(FULL, ibase, 0);
          if (ibase + total > a.item_cap) {
            // Item list full: this photon leaves the wavefront (spill list) and the persistent tracer
            // finishes it.  The first failed reservation may own the list's tail: give it well-formed
            // items (this ray, its first boxed object's first patch) — the Newton kernel runs every
            // entry below min(count, capacity); nobody reads their results.
            const unsigned fp = (unsigned)sv.obj[__ffsll((long long)omask) - 1].first_patch | ((unsigned)(__ffsll((long long)omask) - 1) << 16);
            for (unsigned k = ibase + lane; k < a.item_cap; k += 32) {
              a.it_out.slot[k] = sslot;
              a.it_out.patch[k] = fp;
            }
            total = 0;
            if (lane == src) spill_me = true;
            break;
          }
        }
      }
      if (lane == src) {
        my_begin = ibase;
        my_cnt = total;
      }
      if (lane == 0) solves += total;
    }
    if (spill_me) {  // rare: one atomic per photon
      const unsigned k = atomicAdd(a.spill.count, 1u);  // < q_cap: every photon spills at most once
      a.spill.st[0][k] = org.x; a.spill.st[1][k] = org.y; a.spill.st[2][k] = org.z;
      a.spill.st[3][k] = dir.x; a.spill.st[4][k] = dir.y; a.spill.st[5][k] = dir.z;
      a.spill.st[6][k] = flux.x; a.spill.st[7][k] = flux.y; a.spill.st[8][k] = flux.z;
      a.spill.idd[k] = id | ((unsigned long long)dep << 56);
      best.obj = -1;  // its queue entry reads "hit nothing" (no items): the next step drops it without a trace
      best.dist = SPPM_MAX_NUM;
      my_cnt = 0;
      queries--;      // the query is counted where it is answered
      spilled++;
    }
    if (alive) {
      const unsigned s = my_slot;
      a.qout.ox[s] = org.x; a.qout.oy[s] = org.y; a.qout.oz[s] = org.z;
      a.qout.dx[s] = dir.x; a.qout.dy[s] = dir.y; a.qout.dz[s] = dir.z;
      a.qout.fx[s] = flux.x; a.qout.fy[s] = flux.y; a.qout.fz[s] = flux.z;
      a.qout.idd[s] = id | ((unsigned long long)dep << 56);
      a.qout.hdist[s] = best.dist;
      a.qout.hobj[s] = best.obj < 0 ? -1 : (best.obj | (best.inside ? 1 << 16 : 0));
      a.qout.item_begin[s] = my_begin;
      a.qout.item_cnt[s] = my_cnt;
    }
  }
  const unsigned long long q64 = warp_sum64(queries), s64 = warp_sum64(solves), p64 = warp_sum64(photons),
                           d64 = warp_sum64(deposits), k64 = warp_sum64(kept), o64 = warp_sum64(spilled);
  if (lane == 0) {
    if (o64) atomicAdd(&a.counters->overflow, o64);
    if (q64) atomicAdd(&a.counters->isect_queries, q64);
    if (s64) atomicAdd(&a.counters->patch_solves, s64);
    if (p64) atomicAdd(&a.counters->photons, p64);
    if (d64) atomicAdd(&a.counters->deposits, d64);
    if (k64) atomicAdd(&a.counters->deposits_kept, k64);
  }
}

struct NewtonArgs {
  DevScene scene;
  NewtonCfg ncfg;
  int patches_in_smem;
  uint64_t pass;
  Queue q;
  Items it;
  const unsigned* n_items_ptr;
  unsigned item_cap;
  unsigned* work;  // chunk cursor of this launch (zeroed with the batch's counters)
};

// Bezier3::newton for the queued (photon slot, patch) items.  Each warp takes chunks of C items and runs the
// C*restarts tasks over its 32 lanes with flush_items (scene.cuh): lanes = restarts, arg-min on (t, restart)
// per item.  Chunks come from a static partition, or (WfBatch::newton_dynamic) from a global cursor: the work of
// an item is either a full solve or nothing (46 % are proven empty by the 13-DOP line test), but measured on
// the B200 the kernel is bound by the latency of the FP64 chain at 2 warps per scheduler, not by imbalance
// (0.83 ms per round either way), and warps that finish early make room for the other lanes' kernels.
__global__ void __launch_bounds__(TRACE_THREADS, SPPM_NEWTON_MINBLOCKS) newton_items_kernel(const NewtonArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const unsigned n = min(*a.n_items_ptr, a.item_cap);
  if (n == 0) return;
  const size_t scene_bytes = scene_smem_bytes(a.scene.n_obj, a.scene.n_patch, a.patches_in_smem);
  const int lane = threadIdx.x & 31;
  const unsigned warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
  // static: C items per warp so that all warps of the grid get work (few items: one per warp keeps the serial
  // Newton chain short); dynamic (a.work != null): about four chunks per warp from a global cursor
  // dynamic only where the static chunk would hit its cap of 32 items (4 teapots at 4K: 37 vs 46.5 ms per
  // 64 M photons); with fewer items the static partition is as fast and warps that run out of work early make
  // room for the other lanes' kernels (teapot room: 773-780 vs 755 M photons/s pipelined)
  const bool dynamic = a.work != nullptr && n > 32u * n_warps;
  const unsigned per = dynamic ? 4u * n_warps : n_warps;
  unsigned C = (n + per - 1) / per;
  C = C < 1 ? 1 : (C > 32 ? 32 : C);
#ifndef SPPM_NO_EARLY_EXIT
  // static partition: a CTA whose first warp's chunk starts beyond the list has no item (tail bounces: a handful of
  // items for a grid sized for thousands) and leaves before staging the scene
  if (!dynamic && (blockIdx.x * blockDim.x >> 5) * C >= n) return;
#endif
  const SceneView sv = stage_scene(a.scene, smem, a.patches_in_smem);
  WarpScratch* ws = reinterpret_cast<WarpScratch*>(smem + scene_bytes) + (threadIdx.x >> 5);
  for (unsigned c_static = warp_global * C;; c_static += n_warps * C) {
    unsigned c0 = c_static;
    if (dynamic) {
      if (lane == 0) c0 = atomicAdd(a.work, C);
      c0 = __shfl_sync(FULL, c0, 0);
    }
    if (c0 >= n) break;
    const int cnt = (int)min(C, n - c0);
    if (lane < cnt) {
      const unsigned s = a.it.slot[c0 + lane];
      const unsigned pk = a.it.patch[c0 + lane];
      ws->ray[0][lane] = a.q.ox[s]; ws->ray[1][lane] = a.q.oy[s]; ws->ray[2][lane] = a.q.oz[s];
      ws->ray[3][lane] = a.q.dx[s]; ws->ray[4][lane] = a.q.dy[s]; ws->ray[5][lane] = a.q.dz[s];
      const unsigned long long idd = a.q.idd[s];
      ws->id[lane] = idd & ID_MASK;
      ws->node[lane] = (unsigned)(idd >> 56);
      ws->item[lane] = (unsigned)lane | ((pk & 0xffffu) << 8);
    }
    __syncwarp();
    flush_items(sv, ws, cnt, a.ncfg, a.pass, SPPM_DOMAIN_PHOTON, nullptr, 0, lane);
    if (lane < cnt) {
      a.it.t[c0 + lane] = ws->it_t[lane];
      a.it.u[c0 + lane] = ws->it_u[lane];
      a.it.v[c0 + lane] = ws->it_v[lane];
    }
    __syncwarp();
  }
}

struct WfLayout {
  Queue q[2];
  Items it[2];
  unsigned* counters;  // [4*b + r] queue count of region r at bounce b, [64+b] item counts, [97] spilled photons, [100+b] Newton chunk cursors
  unsigned item_cap;
  Spill spill;         // [cap]
};

// (ray, patch) items a bounce of `cap` photons may queue before photons spill to the persistent tracer
inline int64_t wf_item_cap(int64_t cap) { return std::max<int64_t>(cap * 2, 1 << 16); }
size_t wf_bytes(int64_t cap, int64_t item_cap) {
  return 2 * (size_t)(2 * cap) * (10 * 8 + 8 + 4 + 4 + 4) + 2 * (size_t)item_cap * (4 + 4 + 24) + (size_t)cap * 80 + 128 * 4 + 65536;
}
WfLayout wf_layout(void* store, int64_t cap, int64_t item_cap) {
  WfLayout L;
  unsigned char* p = (unsigned char*)store;
  auto take = [&](size_t bytes) {
    void* r = p;
    p += (bytes + 255) & ~(size_t)255;
    return r;
  };
  for (int k = 0; k < 2; k++) {
    Queue& q = L.q[k];
    double** d[] = {&q.ox, &q.oy, &q.oz, &q.dx, &q.dy, &q.dz, &q.fx, &q.fy, &q.fz, &q.hdist};
    for (double** x : d) *x = (double*)take((size_t)cap * 16);  // 2 x cap entries: four regions, see photon_step_kernel
    q.idd = (unsigned long long*)take((size_t)cap * 16);
    q.hobj = (int*)take((size_t)cap * 8);
    q.item_begin = (unsigned*)take((size_t)cap * 8);
    q.item_cnt = (unsigned*)take((size_t)cap * 8);
    Items& it = L.it[k];
    it.slot = (unsigned*)take((size_t)item_cap * 4);
    it.patch = (unsigned*)take((size_t)item_cap * 4);
    it.t = (double*)take((size_t)item_cap * 8);
    it.u = (double*)take((size_t)item_cap * 8);
    it.v = (double*)take((size_t)item_cap * 8);
  }
  for (int k = 0; k < 9; k++) L.spill.st[k] = (double*)take((size_t)cap * 8);
  L.spill.idd = (unsigned long long*)take((size_t)cap * 8);
  L.counters = (unsigned*)take(128 * 4);
  L.spill.count = L.counters + 97;
  L.item_cap = (unsigned)item_cap;
  return L;
}

// ---- one batch's launch arguments (host) ------------------------------------------------------
// Everything a batch's launches need beyond the queue layout; filled from sppm_ctx + Lane by
// launch_photon_trace_wavefront (wavefront.cu) and, for the CPU emulation, by tests/device_host.
struct WfBatch {
  DevScene scene;
  int patches_in_smem, newton_cull;
  int max_dep, newton_restarts, newton_iters, newton_mode, fast_math;
  uint64_t seed, pass;
  DevCounters* counters;
  uint64_t begin, count;  // photon range per light
  int tail_bounce;        // >= 0: from this bounce on Bezier-pending photons spill to the persistent tracer
  int newton_dynamic;     // Newton chunks from a global cursor instead of a static partition
  unsigned q_cap;         // entries per queue half (>= n_light * count)
  Grid grid;
  const int32_t* cell_slot;
  int32_t n_slots;
  int keep_all;
  Deposit* dep;
  uint32_t* dep_key;
  uint32_t* dep_rank;
  unsigned int* slot_count;
  unsigned long long dep_cap;
  unsigned long long* dep_count;
};
inline int wf_n_steps(int max_dep) { return max_dep + 2; }  // queries at dep = 0 .. max_dep+1 (Q9)
// photon_step_kernel's arguments at bounce b
inline StepArgs wf_step_args(const WfBatch& w, const WfLayout& L, int b) {
  StepArgs a;
  a.scene = w.scene;
  a.patches_in_smem = w.patches_in_smem;
  a.max_dep = w.max_dep;
  a.bounce = b;
  a.do_intersect = b < wf_n_steps(w.max_dep) - 1;
  a.fast_math = w.fast_math;
  a.seed = w.seed;
  a.pass = w.pass;
  a.counters = w.counters;
  a.qin = L.q[b & 1];
  a.qout = L.q[(b + 1) & 1];
  a.it_in = L.it[b & 1];
  a.it_out = L.it[(b + 1) & 1];
  a.n_in_ptr = L.counters + 4 * b;
  a.n_in_host = (unsigned)((uint64_t)w.scene.n_light * w.count);
  a.n_out = L.counters + 4 * (b + 1);
  a.q_cap = w.q_cap;
  a.n_items_out = L.counters + 64 + b + 1;
  a.item_cap = L.item_cap;
  a.spill = L.spill;
  a.spill_all = w.tail_bounce >= 0 && b >= w.tail_bounce;
  a.begin = w.begin;
  a.count = w.count;
  a.grid = w.grid;
  a.cell_slot = w.cell_slot;
  a.n_slots = w.n_slots;
  a.keep_all = w.keep_all;
  a.dep = w.dep;
  a.dep_key = w.dep_key;
  a.dep_rank = w.dep_rank;
  a.slot_count = w.slot_count;
  a.dep_cap = w.dep_cap;
  a.dep_count = w.dep_count;
  return a;
}
// newton_items_kernel's arguments for the items step `a` queued
inline NewtonArgs wf_newton_args(const WfBatch& w, const WfLayout& L, const StepArgs& a) {
  NewtonArgs n;
  n.scene = w.scene;
  n.ncfg.restarts = w.newton_restarts;
  n.ncfg.iters = w.newton_iters;
  n.ncfg.mode = w.newton_mode;
  n.ncfg.seed = w.seed;
  n.ncfg.cull = w.newton_cull;
  n.patches_in_smem = a.patches_in_smem;
  n.pass = w.pass;
  n.q = a.qout;
  n.it = a.it_out;
  n.n_items_ptr = a.n_items_out;
  n.item_cap = L.item_cap;
  n.work = w.newton_dynamic ? L.counters + 100 + a.bounce : nullptr;
  return n;
}

}  // namespace

}  // namespace sppm
