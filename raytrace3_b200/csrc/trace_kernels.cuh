// trace_kernels.cuh — the kernels of trace.cu (launchers and context plumbing stay there):
//   eye_pass_kernel      RayTracer::GetHitPoint + ViewTrace   (Raytracer.h:102-116, 296-357)
//   photon_trace_kernel  Light::emit + RayTracer::PhotonTrace (Light.h:8-13, Raytracer.h:117-209)
//   intersect_kernel     Scene::NearCollide on a ray batch     (Scene.h:165-182)
//   patch_batch_kernel   Bezier3::newton on (ray, patch) pairs (Bezier.h:112-159)
// A header of their own so that tests/device_host can compile these very kernel bodies for the
// CPU (one-warp blocks on fibers) and hold them to the oracle without a GPU; trace.cu is the only
// product file that includes it.
#pragma once
#include "ctx.h"

#ifndef SPPM_TRACE_MINBLOCKS
#define SPPM_TRACE_MINBLOCKS 2  // CTAs per SM the photon kernel's register budget is sized for
#endif

namespace sppm {

namespace {

struct TraceCommon {
  DevScene scene;
  NewtonCfg ncfg;
  int patches_in_smem;
  int max_dep;
  int fast_math;
  uint64_t pass;
  DevCounters* counters;
};

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
  return v;
}

__device__ __forceinline__ WarpScratch* warp_scratch(unsigned char* smem, size_t scene_bytes) {
  return reinterpret_cast<WarpScratch*>(smem + scene_bytes) + (threadIdx.x >> 5);
}

// ---------------------------------------------------------------------------------------
struct EyeArgs {
  TraceCommon c;
  SppmCameraDesc cam;
  HpRaw* out;
  unsigned long long cap;
  unsigned long long* hp_count;
  double init_r2;
  long long pix_begin, pix_end;  // pixels [row_begin * w, row_end * w) of this launch (eye-pass sharding)
};
struct EyeStackEnt {
  d3 org, dir, wgt;
  int dep;
  unsigned node;
};

__global__ void __launch_bounds__(TRACE_THREADS) eye_pass_kernel(const EyeArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const size_t scene_bytes = scene_smem_bytes(a.c.scene.n_obj, a.c.scene.n_patch, a.c.patches_in_smem);
  const SceneView sv = stage_scene(a.c.scene, smem, a.c.patches_in_smem);
  WarpScratch* ws = warp_scratch(smem, scene_bytes);
  const int lane = threadIdx.x & 31;
  const long long npix = a.pix_end;
  const long long p = a.pix_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int row = (int)(p / a.cam.w), col = (int)(p % a.cam.w);

  EyeStackEnt stack[16];
  int sp = 0;
  bool work = p < npix;
  d3 org = ld3(a.cam.pos), dir = mk(0, 0, 1), wgt = mk(1, 1, 1);
  int dep = 0;
  unsigned node = 1;
  if (work) {  // Camera::emit (Camera.h:18-22)
    const d3 d = (ld3(a.cam.du) * ((col + 0.5) / a.cam.w - 0.5) + ld3(a.cam.dv) * ((row + 0.5) / a.cam.h - 0.5)) +
                 ld3(a.cam.dir);
    dir = get_unit(d);
  }
  unsigned long long queries = 0, solves = 0;
  while (__any_sync(FULL, work)) {
    const Hit h = near_collide(sv, ws, a.c.ncfg, work, org, dir, a.c.pass, (uint64_t)p, node, SPPM_DOMAIN_EYE, solves);
    bool emit_hp = false;
    HpRaw rec;
    if (work) {
      queries++;
      bool cont = false;
      if (h.obj >= 0 && !(dep > a.c.max_dep)) {  // Raytracer.h:304-310
        dep++;
        const DevObject& ob = sv.obj[h.obj];
        const Surface s = shade_hit(sv, h, org, dir);
        const d3 cw = mul3(s.color, wgt);
        if (ob.isDiff) {  // Raytracer.h:312-319
          const d3 w = mul3(cw, ld3(ob.diff));
          rec.pos[0] = s.pos.x; rec.pos[1] = s.pos.y; rec.pos[2] = s.pos.z;
          rec.n[0] = s.n.x; rec.n[1] = s.n.y; rec.n[2] = s.n.z;
          rec.wgt[0] = w.x; rec.wgt[1] = w.y; rec.wgt[2] = w.z;
          rec.r2 = a.init_r2;
          rec.row = row;
          rec.col = col;
          rec.okey = ((uint64_t)p << 20) | node;
          emit_hp = true;
        }
        d3 rdir = dir, tdir = dir;
        if (ob.isRefl) rdir = get_unit(get_refl(dir, s.n));  // Raytracer.h:320-326
        if (ob.isRefr) {                                      // Raytracer.h:327-335
          const double refrn = h.inside ? ob.refrn : 1 / ob.refrn;
          tdir = normalize3(refracted(dir, h.inside ? -s.n : s.n, refrn));
        }
        if (ob.isRefl && ob.isRefr) {  // defer the refraction subtree
          if (sp < 16) {
            stack[sp].org = s.pos;
            stack[sp].dir = tdir;
            stack[sp].wgt = mul3(cw, ld3(ob.refr));
            stack[sp].dep = dep;
            stack[sp].node = node * 2 + 1;
            sp++;
          }
        }
        if (ob.isRefl) {
          org = s.pos; dir = rdir; wgt = mul3(cw, ld3(ob.refl)); node = node * 2;
          cont = true;
        } else if (ob.isRefr) {
          org = s.pos; dir = tdir; wgt = mul3(cw, ld3(ob.refr)); node = node * 2 + 1;
          cont = true;
        }
      }
      if (!cont) {
        if (sp > 0) {
          sp--;
          org = stack[sp].org; dir = stack[sp].dir; wgt = stack[sp].wgt; dep = stack[sp].dep; node = stack[sp].node;
        } else {
          work = false;
        }
      }
    }
    // warp-aggregated append of the new hit points
    const unsigned m = __ballot_sync(FULL, emit_hp);
    if (m) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.hp_count, (unsigned long long)__popc(m));
      base = __shfl_sync(FULL, base, 0);
      if (emit_hp) {
        const unsigned long long slot = base + __popc(m & ((1u << lane) - 1));
        if (slot < a.cap) a.out[slot] = rec;
      }
    }
  }
  queries = warp_sum(queries);
  solves = warp_sum(solves);
  if (lane == 0) {
    atomicAdd(&a.c.counters->isect_queries, queries);
    atomicAdd(&a.c.counters->patch_solves, solves);
  }
}

// ---------------------------------------------------------------------------------------
struct PhotonArgs {
  TraceCommon c;
  Grid grid;
  const int32_t* cell_slot;  // null: no hit-point map (deposits kept only with keep_all)
  int32_t n_slots;
  int keep_all;
  Deposit* dep;
  uint32_t* dep_key;
  uint32_t* dep_rank;
  unsigned int* slot_count;
  unsigned long long dep_cap;
  unsigned long long* dep_count;
  unsigned long long* cursor;
  uint64_t begin, count;  // photon range per light
  // Resume mode (r_count != null): instead of emitting, take *r_count photons in flight from a list of
  // states (org, dir, flux, id | dep << 56) — the photons the wavefront tracer spilled when the (ray, patch)
  // item list of a bounce was full (wavefront_kernels.cuh).  Same RNG addresses, so the same paths.
  const unsigned* r_count;
  const double* r_st[9];
  const unsigned long long* r_idd;
};

// Persistent warps: a lane whose photon ended picks the next photon index from a global
// cursor, so lanes stay busy although path lengths differ (1..15 queries, Q9).  The RNG is
// addressed by photon id, so the result does not depend on which lane traces what.
__global__ void __launch_bounds__(TRACE_THREADS, SPPM_TRACE_MINBLOCKS) photon_trace_kernel(const PhotonArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  if (a.r_count && *a.r_count == 0) return;  // nothing was spilled (the usual case): no scene staging either
  const size_t scene_bytes = scene_smem_bytes(a.c.scene.n_obj, a.c.scene.n_patch, a.c.patches_in_smem);
  const SceneView sv = stage_scene(a.c.scene, smem, a.c.patches_in_smem);
  WarpScratch* ws = warp_scratch(smem, scene_bytes);
  const int lane = threadIdx.x & 31;
  const unsigned long long total = a.r_count ? (unsigned long long)*a.r_count : (unsigned long long)a.c.scene.n_light * a.count;

  bool alive = false, exhausted = false;
  d3 org = mk(0, 0, 0), dir = mk(0, 0, 1), flux = mk(0, 0, 0);
  int dep = 0;
  uint64_t id = 0;
  unsigned long long queries = 0, solves = 0, photons = 0, deposits = 0, kept = 0;

  for (;;) {
    const unsigned dead = __ballot_sync(FULL, !alive);
    if (dead && !exhausted) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.cursor, (unsigned long long)__popc(dead));
      base = __shfl_sync(FULL, base, 0);
      if (base + __popc(dead) >= total) exhausted = true;
      if (!alive) {
        const unsigned long long q = base + __popc(dead & ((1u << lane) - 1));
        if (q < total && a.r_count) {  // a photon in flight
          org = mk(a.r_st[0][q], a.r_st[1][q], a.r_st[2][q]);
          dir = mk(a.r_st[3][q], a.r_st[4][q], a.r_st[5][q]);
          flux = mk(a.r_st[6][q], a.r_st[7][q], a.r_st[8][q]);
          const unsigned long long idd = a.r_idd[q];
          id = idd & 0x00ffffffffffffffull;
          dep = (int)(idd >> 56);
          alive = true;
        } else if (q < total) {  // Light::emit (Light.h:8-13)
          const unsigned pl = (unsigned)(q / a.count);
          id = ((uint64_t)pl << 40) | (a.begin + q % a.count);
          const SppmLight& L = a.c.scene.lights[pl];
          flux = ld3(L.color) * SPPM_PI * 4;
          dir = unit_rand_vec(a.c.ncfg.seed, a.c.pass, id, P_EMIT);
          org = ld3(L.pos);
          dep = 0;
          alive = true;
          photons++;
        }
      }
    }
    if (!__any_sync(FULL, alive)) break;

    const Hit h = near_collide(sv, ws, a.c.ncfg, alive, org, dir, a.c.pass, id, (uint32_t)dep, SPPM_DOMAIN_PHOTON, solves);
    bool want = false;
    Deposit rec;
    uint32_t key = 0;
    if (alive) {
      queries++;
      if (h.obj < 0 || dep > a.c.max_dep) {  // Raytracer.h:127-132
        alive = false;
      } else {
        const int node = dep;
        dep++;
        const DevObject& ob = sv.obj[h.obj];
        const Surface s = shade_hit(sv, h, org, dir);
        if (ob.isDiff) {  // Raytracer.h:136-160: the gather input (x, N, incoming flux)
          deposits++;
          int32_t slot = -1;
          if (a.cell_slot && grid_keeps(a.grid, s.pos.x, s.pos.y, s.pos.z))
            slot = __ldg(a.cell_slot + grid_cell(a.grid, s.pos.x, s.pos.y, s.pos.z));
          if (slot >= 0 || a.keep_all) {
            want = true;
            key = slot >= 0 ? (uint32_t)slot : (uint32_t)a.n_slots;
            rec.x[0] = s.pos.x; rec.x[1] = s.pos.y; rec.x[2] = s.pos.z;
            rec.n[0] = s.n.x; rec.n[1] = s.n.y; rec.n[2] = s.n.z;
            rec.flux[0] = flux.x; rec.flux[1] = flux.y; rec.flux[2] = flux.z;
            rec.meta = id | ((uint64_t)node << 56);
          }
        }
        // Obj::Roulette (Obj.h:30-45) and the single continuation (Raytracer.h:162-190)
        double ur, U1;
        draw2(a.c.ncfg.seed, a.c.pass, id, (uint32_t)node, P_BOUNCE, 0, 0, ur, U1);
        const double r = randu(ur, 0., ob.allr);
        d3 ndir;
        if (ob.diffr > r) {
          double U2, unused;
          draw2(a.c.ncfg.seed, a.c.pass, id, (uint32_t)node, P_BOUNCE, 1, 0, U2, unused);
          ndir = unit_rand_refl(s.n, U1, U2, a.c.fast_math != 0);
        } else if (ob.diffr + ob.reflr > r) {
          ndir = get_unit(get_refl(dir, s.n));
        } else {
          const double refrn = h.inside ? ob.refrn : 1 / ob.refrn;
          ndir = normalize3(refracted(dir, h.inside ? -s.n : s.n, refrn));
        }
        flux = mul3(s.color, flux);
        org = s.pos;
        dir = ndir;
      }
    }
    const unsigned m = __ballot_sync(FULL, want);
    if (m) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.dep_count, (unsigned long long)__popc(m));
      base = __shfl_sync(FULL, base, 0);
      if (want) {
        const unsigned long long slot = base + __popc(m & ((1u << lane) - 1));
        if (slot < a.dep_cap) {
          a.dep[slot] = rec;
          a.dep_key[slot] = key;
          a.dep_rank[slot] = atomicAdd(a.slot_count + key, 1u);  // counting sort: rank within the slot
          kept++;
        }
      }
    }
  }
  queries = warp_sum(queries);
  solves = warp_sum(solves);
  photons = warp_sum(photons);
  deposits = warp_sum(deposits);
  kept = warp_sum(kept);
  if (lane == 0) {
    atomicAdd(&a.c.counters->isect_queries, queries);
    atomicAdd(&a.c.counters->patch_solves, solves);
    atomicAdd(&a.c.counters->photons, photons);
    atomicAdd(&a.c.counters->deposits, deposits);
    atomicAdd(&a.c.counters->deposits_kept, kept);
  }
}

// ---------------------------------------------------------------------------------------
struct IsectArgs {
  TraceCommon c;
  const double* rays6;
  const uint64_t* ids;
  long long n;
  int node;
  SppmHit* out;
};
__global__ void __launch_bounds__(TRACE_THREADS) intersect_kernel(const IsectArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const size_t scene_bytes = scene_smem_bytes(a.c.scene.n_obj, a.c.scene.n_patch, a.c.patches_in_smem);
  const SceneView sv = stage_scene(a.c.scene, smem, a.c.patches_in_smem);
  WarpScratch* ws = warp_scratch(smem, scene_bytes);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < a.n;
  d3 org = mk(0, 0, 0), dir = mk(0, 0, 1);
  uint64_t id = (uint64_t)i;
  if (active) {
    org = ld3(a.rays6 + 6 * i);
    dir = ld3(a.rays6 + 6 * i + 3);
    if (a.ids) id = a.ids[i];
  }
  unsigned long long solves = 0;
  const Hit h = near_collide(sv, ws, a.c.ncfg, active, org, dir, a.c.pass, id, (uint32_t)a.node, SPPM_DOMAIN_TEST, solves);
  if (active) {
    SppmHit o;
    o.t = h.obj >= 0 ? h.dist : SPPM_MAX_NUM;
    o.u = h.u; o.v = h.v;
    o.obj = h.obj; o.patch = h.patch; o.inside = h.inside ? 1 : 0; o.pad = 0;
    for (int k = 0; k < 3; k++) o.pos[k] = o.n[k] = o.color[k] = 0;
    if (h.obj >= 0) {
      const Surface s = shade_hit(sv, h, org, dir);
      o.pos[0] = s.pos.x; o.pos[1] = s.pos.y; o.pos[2] = s.pos.z;
      o.n[0] = s.n.x; o.n[1] = s.n.y; o.n[2] = s.n.z;
      o.color[0] = s.color.x; o.color[1] = s.color.y; o.color[2] = s.color.z;
    }
    a.out[i] = o;
  }
}

struct PatchArgs {
  TraceCommon c;
  const double* rays6;
  const int32_t* patch;
  long long n;
  const double* seeds;
  double* tuv;
};
// One warp solves 32 (ray, patch) pairs per round: 32 * restarts Newton tasks over 32 lanes.
__global__ void __launch_bounds__(TRACE_THREADS) patch_batch_kernel(const PatchArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const size_t scene_bytes = scene_smem_bytes(a.c.scene.n_obj, a.c.scene.n_patch, a.c.patches_in_smem);
  const SceneView sv = stage_scene(a.c.scene, smem, a.c.patches_in_smem);
  WarpScratch* ws = warp_scratch(smem, scene_bytes);
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long base = warp_global * 32; base < a.n; base += n_warps * 32) {
    const long long i = base + lane;
    const bool active = i < a.n;
    if (active) {
      for (int k = 0; k < 6; k++) ws->ray[k][lane] = a.rays6[6 * i + k];
      ws->id[lane] = (uint64_t)i;
      ws->node[lane] = 0;
      ws->item[lane] = (unsigned)lane | ((unsigned)a.patch[i] << 8);
    }
    const int n_items = (int)min((long long)32, a.n - base);
    __syncwarp();
    flush_items(sv, ws, n_items, a.c.ncfg, 0, SPPM_DOMAIN_TEST, a.seeds, base, lane);
    if (active) {
      a.tuv[3 * i + 0] = ws->it_t[lane];
      a.tuv[3 * i + 1] = ws->it_u[lane];
      a.tuv[3 * i + 2] = ws->it_v[lane];
    }
    __syncwarp();
  }
}

}  // namespace

}  // namespace sppm
