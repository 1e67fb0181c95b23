// scene.cuh — device scene view and Scene::NearCollide as a warp-cooperative routine.
//
// Analytic objects (planes, spheres) are tested per lane.  Bezier objects are handled by
// the whole warp: every lane culls its ray against the object / patch AABBs
// (Bezier.h:20-51, 176-186, 240-250), surviving (ray, patch) pairs go to a per-warp queue
// in shared memory, and the queue's (pair, restart) Newton tasks are spread over all 32
// lanes — lanes = restarts, control points read from shared memory, 3x3 inverse in
// registers — followed by a warp min-reduction on t (Bezier.h:146-158).
// Citations: /root/reference/raytracer/<file>:<line>.
#pragma once
#include "../../include/sppm_b200.h"
#include "device_math.cuh"

namespace sppm {

constexpr int MAX_OBJECTS = 64;
constexpr int QCAP = 96;  // (ray, patch) queue entries per warp
constexpr int TRACE_THREADS = 128;
constexpr int TRACE_WARPS = TRACE_THREADS / 32;

struct DevObject {
  int kind, tex_id, first_patch, n_patch;
  int isDiff, isRefl, isRefr;  // Element.h:16-18
  int udex, vdex;              // Obj.h:89-96
  int pad;
  double diffr, reflr, refrr, allr;  // Obj.h:12-15
  double diff[3], refl[3], refr[3], refrn;
  double color[3];
  double p0[3], norm[3], radius;
  double texU[3], texV[3], texW[3];  // texW = texU x texV (Obj.h:151)
  double texUmod, texVmod;           // Obj.h:97-98 Module()
  double box[6];                     // Bezier3Obj::aabb
};
// Texels are stored one 32-bit word each (B | G << 8 | R << 16) in 8 x 4 tiles (one 128-byte line per tile:
// block-linear), so the 2 x 2 footprint of a bilinear lookup (Element.h:61-72) falls into one line two times
// out of three instead of always two lines 4*cols bytes apart; below level 0 a chain of 2 x 2 box-filtered
// levels (SppmParams.tex_lod > 0: an opt-in approximation; level 0 is the reference's image, bit for bit).
constexpr int TEX_MAX_LEVELS = 8;
struct TexLevel {
  const uint32_t* texel;
  int rows, cols, tiles_x, pad;
};
struct DevTexture {
  int rows, cols;
  const uint8_t* bgr;   // the uploaded BGR8 image (row-major; JPEG encode / export paths)
  int n_levels, pad;
  TexLevel level[TEX_MAX_LEVELS];
};
__host__ __device__ inline size_t tex_tiled_index(int r, int c, int tiles_x) {
  return ((size_t)(r >> 2) * tiles_x + (c >> 3)) * 32 + (size_t)((r & 3) * 8 + (c & 7));
}
__host__ __device__ inline size_t tex_tiled_words(int rows, int cols) { return (size_t)((rows + 3) >> 2) * ((cols + 7) >> 3) * 32; }
struct DevScene {
  int n_obj, n_patch, n_light, n_tex;
  const DevObject* obj;     // [n_obj]
  const double* patchP;     // [n_patch][48] Bernstein control points
  const double* patchA;     // [n_patch][48] the same patches in the monomial basis: A[i][k][c], S = sum A u^k v^i
  const double* patchBox;   // [n_patch][6]
  const double* patchDop;   // [n_patch][26] 13-DOP of the control points, grown by the acceptance radius (dop_line_miss)
  const double* patchSub;   // [n_patch][16][6] AABBs of the 4x4 de Casteljau sub-patches' control points, grown likewise
  unsigned long long* culled;  // counter: Newton solves proven empty by the 13-DOP line test
  const SppmLight* lights;  // [n_light]
  const DevTexture* tex;    // [n_tex]
  int tex_lod;              // SppmParams.tex_lod: mip level the tracers sample (0 = the reference's image)
};
struct NewtonCfg {
  int restarts, iters, mode;
  int cull;  // 1: skip solves the 13-DOP line test proves empty (0 only for the M2 microbenchmark, which times full solves)
  uint64_t seed;
};

// per-warp shared-memory scratch of near_collide
struct WarpScratch {
  double ray[6][32];
  unsigned long long id[32];
  unsigned int node[32];
  unsigned int item[QCAP];  // lane | global patch << 8
  unsigned char live[QCAP]; // items that pass the 13-DOP line test (flush_items)
  double it_t[QCAP], it_u[QCAP], it_v[QCAP];
};
// block-level shared view of the scene (objects + patches staged by stage_scene)
struct SceneView {
  const DevObject* obj;
  const double* patchP;
  const double* patchA;
  const double* patchBox;
  const double* patchDop;
  const double* patchSub;
  unsigned long long* culled;
  const DevTexture* tex;
  const double* lut255;  // lut255[b] = b / 255.0 (Element.h:54-57), exact quotients
  int n_obj;
  int tex_lod;
};

// with_A: also stage the monomial coefficients (only kernels that run the FAST Newton need them)
__host__ __device__ inline size_t scene_smem_bytes(int n_obj, int n_patch, bool patches_in_smem, bool with_A = true) {
  size_t b = (size_t)n_obj * sizeof(DevObject) + 256 * sizeof(double);
  if (patches_in_smem) b += (size_t)n_patch * (with_A ? 102 : 54) * sizeof(double);
  return (b + 15) & ~(size_t)15;
}

// Global -> shared copy of n doubles by the whole block, eight loads in flight per thread (the plain loop
// `dst[i] = src[i]` compiles to LDG -> STS -> branch: one exposed L2 latency per element and thread, 31 of them
// for the scene with the monomial coefficients).  Measured on the B200 against the plain loop: within noise
// (782 vs 779-783 M photons/s, Newton launches 0.78 vs 0.78-0.79 ms) — the staging is not what a tail-bounce
// launch waits for; kept because it is never slower.
__device__ __forceinline__ void stage_copy(double* __restrict__ dst, const double* __restrict__ src, int n) {
  const int bd = (int)blockDim.x;
  for (int i = (int)threadIdx.x; i < n; i += 8 * bd) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = (i + k * bd < n) ? src[i + k * bd] : 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (i + k * bd < n) dst[i + k * bd] = v[k];
  }
}

// Copy objects (and patches if they fit) into dynamic shared memory; all threads call.
__device__ inline SceneView stage_scene(const DevScene& s, unsigned char* smem, bool patches_in_smem, bool with_A = true) {
  SceneView v;
  {
    const int n = s.n_obj * (int)(sizeof(DevObject) / 8);
    stage_copy(reinterpret_cast<double*>(smem), reinterpret_cast<const double*>(s.obj), n);
    v.obj = reinterpret_cast<const DevObject*>(smem);
  }
  {
    double* lut = reinterpret_cast<double*>(smem + (size_t)s.n_obj * sizeof(DevObject));
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = i / 255.0;
    v.lut255 = lut;
  }
  if (patches_in_smem) {
    double* dp = reinterpret_cast<double*>(smem + (size_t)s.n_obj * sizeof(DevObject)) + 256;
    double* db = dp + (size_t)s.n_patch * 48;
    stage_copy(dp, s.patchP, s.n_patch * 48);
    stage_copy(db, s.patchBox, s.n_patch * 6);
    v.patchP = dp;
    v.patchBox = db;
    v.patchA = s.patchA;
    if (with_A) {
      double* da = db + (size_t)s.n_patch * 6;
      stage_copy(da, s.patchA, s.n_patch * 48);
      v.patchA = da;
    }
  } else {
    v.patchP = s.patchP;
    v.patchA = s.patchA;
    v.patchBox = s.patchBox;
  }
  v.patchDop = s.patchDop;
  v.patchSub = s.patchSub;
  v.culled = s.culled;
  v.tex = s.tex;
  v.tex_lod = s.tex_lod;
  v.n_obj = s.n_obj;
  __syncthreads();
  return v;
}

// ---- Element.h:61-72 Texture::colorUV ------------------------------------------------
// tex_fetch issues the four texel loads (one 32-bit word per texel instead of three byte loads),
// tex_finish does the look-ups of b / 255.0 and the bilinear blend in the reference's expression
// order.  (Deferring tex_finish behind the direction arithmetic was measured: the extra live
// registers cost more than the hidden latency.)
struct TexFetch {
  uint32_t p11, p12, p21, p22;
  double detR, detC;
};
__device__ __forceinline__ TexFetch tex_fetch(const DevTexture& t, double u, double v, int lod = 0) {
  const TexLevel& L = t.level[lod < t.n_levels ? (lod > 0 ? lod : 0) : t.n_levels - 1];
  const int rows = L.rows, cols = L.cols;
  double row = (u - floor(u)) * rows, col = (v - floor(v)) * cols;
  int r1 = (int)floor(row + 1e-10), r2 = r1 + 1;
  int c1 = (int)floor(col + 1e-10), c2 = c1 + 1;
  TexFetch f;
  f.detR = r2 - row;
  f.detC = c2 - col;
  r1 = (r1 >= 0) ? (r1 >= rows ? 0 : r1) : (rows - 1);
  c1 = (c1 >= 0) ? (c1 >= cols ? 0 : c1) : (cols - 1);
  r2 = (r2 < rows) ? r2 : 0;
  c2 = (c2 < cols) ? c2 : 0;
  f.p11 = __ldg(L.texel + tex_tiled_index(r1, c1, L.tiles_x));
  f.p12 = __ldg(L.texel + tex_tiled_index(r1, c2, L.tiles_x));
  f.p21 = __ldg(L.texel + tex_tiled_index(r2, c1, L.tiles_x));
  f.p22 = __ldg(L.texel + tex_tiled_index(r2, c2, L.tiles_x));
  return f;
}
__device__ __forceinline__ d3 tex_finish(const TexFetch& f, const double* __restrict__ lut255) {
  double o[3];
  const double detR = f.detR, detC = f.detC;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double a = lut255[(f.p11 >> (8 * k)) & 255u], b = lut255[(f.p12 >> (8 * k)) & 255u], c = lut255[(f.p21 >> (8 * k)) & 255u],
                 d = lut255[(f.p22 >> (8 * k)) & 255u];
    o[k] = ((a * detR * detC + b * detR * (1 - detC)) + c * (1 - detR) * detC) + d * (1 - detR) * (1 - detC);
  }
  return d3{o[0], o[1], o[2]};
}
__device__ inline d3 color_uv(const DevTexture& t, const double* __restrict__ lut255, double u, double v, int lod = 0) {
  return tex_finish(tex_fetch(t, u, v, lod), lut255);
}

// ---- Bezier.h:20-51 AABBbox::intersect (entry faces only, Q5) ---------------------------
__device__ inline bool aabb_intersect(const double* box, d3 org, d3 dir) {
  const double o[3] = {org.x, org.y, org.z}, d[3] = {dir.x, dir.y, dir.z};
#pragma unroll
  for (int i = 0; i < 3; i++) {
    if (!is_zero(d[i])) {
      const double num = (d[i] > 0) ? (box[i] - o[i]) : (box[3 + i] - o[i]);
      if ((num > 0) != (d[i] > 0) || num == 0) continue;  // t <= 0: no division needed to know it
      const double t = num / d[i];
      if (t > 0) {
        bool flag = true;
#pragma unroll
        for (int j = 0; j < 3; j++) {
          if (j != i) {
            double c = o[j] + t * d[j];
            if (box[3 + j] < c || box[j] > c) flag = false;
          }
        }
        if (flag) return true;
      }
    }
  }
  return false;
}

// ---- 13-DOP line test: Newton solves that cannot record anything are skipped ---------------
// Bezier3::newton (Bezier.h:112-159) records an iterate only if u, v are in [0,1] and the point
// org + dir t lies within sqrt(1e-4) = 0.01 of S(u,v) (Bezier.h:135).  S([0,1]^2) lies in the
// convex hull of the 16 control points, hence inside their 13-DOP (slabs along the 13 directions
// with components in {0,+-1}); the slabs are stored grown by 0.0101 |a|.  If the INFINITE line
// (negative t counts, Q4) misses that polytope, no iterate of any restart can be recorded and the
// solve returns "nothing" without being run -- same result, 30 % fewer solves in the reference
// scene (the patch AABB test of Bezier.h:178 passes 2.5x more pairs than actually intersect).
static __device__ __noinline__ bool dop_line_miss(const double* __restrict__ dop, d3 o, d3 d) {
  const double ox[13] = {o.x, o.y, o.z, o.x + o.y, o.x - o.y, o.x + o.z, o.x - o.z, o.y + o.z, o.y - o.z,
                         (o.x + o.y) + o.z, (o.x + o.y) - o.z, (o.x - o.y) + o.z, (o.y - o.x) + o.z};
  const double dx[13] = {d.x, d.y, d.z, d.x + d.y, d.x - d.y, d.x + d.z, d.x - d.z, d.y + d.z, d.y - d.z,
                         (d.x + d.y) + d.z, (d.x + d.y) - d.z, (d.x - d.y) + d.z, (d.y - d.x) + d.z};
  double tmin = -1e300, tmax = 1e300;
#pragma unroll
  for (int k = 0; k < 13; k++) {
    const double lo = __ldg(dop + k) - ox[k], hi = __ldg(dop + 13 + k) - ox[k];  // lo <= dx t <= hi
    if (dx[k] == 0) {
      if (lo > 0 || hi < 0) return true;
    } else {
      const double inv = 1.0 / dx[k];
      const double t0 = lo * inv, t1 = hi * inv;
      tmin = fmax(tmin, fmin(t0, t1));
      tmax = fmin(tmax, fmax(t0, t1));
    }
  }
  // the slabs carry 1e-4 of slack beyond the acceptance radius; this guard covers the rounding of t
  return tmin > tmax + 1e-9 * (1.0 + fabs(tmin) + fabs(tmax));
}

// Second, tighter test of the same kind: the patch is the union of its 4x4 de Casteljau
// sub-patches, each inside the AABB of its own 16 control points (stored grown by 0.0101).  A line
// that misses all 16 boxes cannot come within the acceptance radius of the surface.  Together
// with the 13-DOP this proves 46 % of the reference scene's solves empty (39 % record a hit).
static __device__ __noinline__ bool sub_boxes_miss(const double* __restrict__ sub, d3 o, d3 d) {
  const double oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
  double inv[3];
#pragma unroll
  for (int k = 0; k < 3; k++) inv[k] = dd[k] != 0 ? 1.0 / dd[k] : 0.0;
  for (int b = 0; b < 16; b++) {
    const double* bx = sub + b * 6;
    double tmin = -1e300, tmax = 1e300;
    bool miss = false;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double lo = __ldg(bx + k) - oo[k], hi = __ldg(bx + 3 + k) - oo[k];
      if (dd[k] == 0) {
        miss = miss || lo > 0 || hi < 0;
      } else {
        const double t0 = lo * inv[k], t1 = hi * inv[k];
        tmin = fmax(tmin, fmin(t0, t1));
        tmax = fmin(tmax, fmax(t0, t1));
      }
    }
    if (!miss && !(tmin > tmax + 1e-9 * (1.0 + fabs(tmin) + fabs(tmax)))) return false;  // the line may touch this box
  }
  return true;
}

// ---- Bezier.h:69-111 patch evaluation --------------------------------------------------
struct PatchEval {
  d3 S, Su, Sv;
};
__device__ __forceinline__ void bern(double t, double c[4]) {  // Bezier.h:71-74
  c[0] = (1 - t) * (1 - t) * (1 - t);
  c[1] = 3 * t * (1 - t) * (1 - t);
  c[2] = 3 * t * t * (1 - t);
  c[3] = t * t * t;
}
__device__ __forceinline__ void dbern(double t, double c[4]) {  // Bezier.h:79-82
  c[0] = -3 * (1 - t) * (1 - t);
  c[1] = 3 * (1 - t) * (1 - t) - 6 * t * (1 - t);
  c[2] = 6 * t * (1 - t) - 3 * t * t;
  c[3] = 3 * t * t;
}
// P[0]*c0 + P[1]*c1 + P[2]*c2 + P[3]*c3, left-associated (Bezier.h:75)
#define SPPM_COMB(a, b, c, d, w) ((((a) * (w)[0] + (b) * (w)[1]) + (c) * (w)[2]) + (d) * (w)[3])

// EXACT: the operations of evalPatchPoint / evalPatchDerivU / evalPatchDerivV in the
// reference's order (rows at u are shared by Point and DerivV — same inputs, same bits).
__device__ __forceinline__ PatchEval eval_exact(const double* __restrict__ P, double u, double v) {
  double cu[4], cv[4], du[4], dv[4];
  bern(u, cu);
  bern(v, cv);
  dbern(u, du);
  dbern(v, dv);
  PatchEval e;
  double row[4][3];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < 3; k++)
      row[i][k] = SPPM_COMB(P[(4 * i) * 3 + k], P[(4 * i + 1) * 3 + k], P[(4 * i + 2) * 3 + k], P[(4 * i + 3) * 3 + k], cu);
  e.S.x = SPPM_COMB(row[0][0], row[1][0], row[2][0], row[3][0], cv);
  e.S.y = SPPM_COMB(row[0][1], row[1][1], row[2][1], row[3][1], cv);
  e.S.z = SPPM_COMB(row[0][2], row[1][2], row[2][2], row[3][2], cv);
  e.Sv.x = SPPM_COMB(row[0][0], row[1][0], row[2][0], row[3][0], dv);
  e.Sv.y = SPPM_COMB(row[0][1], row[1][1], row[2][1], row[3][1], dv);
  e.Sv.z = SPPM_COMB(row[0][2], row[1][2], row[2][2], row[3][2], dv);
  double col[4][3];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < 3; k++)
      col[i][k] = SPPM_COMB(P[i * 3 + k], P[(i + 4) * 3 + k], P[(i + 8) * 3 + k], P[(i + 12) * 3 + k], cv);
  e.Su.x = SPPM_COMB(col[0][0], col[1][0], col[2][0], col[3][0], du);
  e.Su.y = SPPM_COMB(col[0][1], col[1][1], col[2][1], col[3][1], du);
  e.Su.z = SPPM_COMB(col[0][2], col[1][2], col[2][2], col[3][2], du);
  return e;
}
// FAST: joint tensor-form evaluation with FMA: rows and row derivatives in u, then three
// combinations in v.  Same polynomial, different rounding (<= a few ulp).
__device__ __forceinline__ PatchEval eval_fast(const double* __restrict__ P, double u, double v) {
  double cu[4], cv[4], du[4], dv[4];
  const double mu = 1 - u, mv = 1 - v;
  cu[0] = mu * mu * mu; cu[1] = 3 * u * mu * mu; cu[2] = 3 * u * u * mu; cu[3] = u * u * u;
  cv[0] = mv * mv * mv; cv[1] = 3 * v * mv * mv; cv[2] = 3 * v * v * mv; cv[3] = v * v * v;
  du[0] = -3 * mu * mu; du[1] = 3 * mu * (mu - 2 * u); du[2] = 3 * u * (2 * mu - u); du[3] = 3 * u * u;
  dv[0] = -3 * mv * mv; dv[1] = 3 * mv * (mv - 2 * v); dv[2] = 3 * v * (2 * mv - v); dv[3] = 3 * v * v;
  double S[3] = {0, 0, 0}, Su[3] = {0, 0, 0}, Sv[3] = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; i++) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double p0 = P[(4 * i) * 3 + k], p1 = P[(4 * i + 1) * 3 + k], p2 = P[(4 * i + 2) * 3 + k],
                   p3 = P[(4 * i + 3) * 3 + k];
      const double r = fma(p3, cu[3], fma(p2, cu[2], fma(p1, cu[1], p0 * cu[0])));
      const double dr = fma(p3, du[3], fma(p2, du[2], fma(p1, du[1], p0 * du[0])));
      S[k] = fma(r, cv[i], S[k]);
      Sv[k] = fma(r, dv[i], Sv[k]);
      Su[k] = fma(dr, cv[i], Su[k]);
    }
  }
  PatchEval e;
  e.S = d3{S[0], S[1], S[2]};
  e.Su = d3{Su[0], Su[1], Su[2]};
  e.Sv = d3{Sv[0], Sv[1], Sv[2]};
  return e;
}

// FAST (default): the patch in the monomial basis (coefficients precomputed at scene upload,
// A = M P M^T), Horner in u for the four v-rows and their u-derivatives, Horner in v for S, Su
// and the v-derivative: 84 FMA per evaluation instead of ~165 in Bernstein tensor form.
__device__ __forceinline__ PatchEval eval_mono(const double* __restrict__ A, double u, double v) {
  const double u2 = 2 * u, u3 = 3 * u * u, v2 = 2 * v, v3 = 3 * v * v;
  double S[3], Su[3], Sv[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double r[4], d[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const double a0 = A[(4 * i) * 3 + k], a1 = A[(4 * i + 1) * 3 + k], a2 = A[(4 * i + 2) * 3 + k], a3 = A[(4 * i + 3) * 3 + k];
      r[i] = fma(fma(fma(a3, u, a2), u, a1), u, a0);
      d[i] = fma(a3, u3, fma(a2, u2, a1));
    }
    S[k] = fma(fma(fma(r[3], v, r[2]), v, r[1]), v, r[0]);
    Sv[k] = fma(r[3], v3, fma(r[2], v2, r[1]));
    Su[k] = fma(fma(fma(d[3], v, d[2]), v, d[1]), v, d[0]);
  }
  PatchEval e;
  e.S = d3{S[0], S[1], S[2]};
  e.Su = d3{Su[0], Su[1], Su[2]};
  e.Sv = d3{Sv[0], Sv[1], Sv[2]};
  return e;
}

// 1/x to ~1 ulp without the IEEE division routine: hardware seed (20 bits) + two Newton-Raphson
// steps in FMA.  FAST Newton mode only (its results are already "the same polynomial, different
// rounding", checked against EXACT to 1e-6 by the sweep tests).
__device__ __forceinline__ double rcp_newton(double x) {
  double r;
#ifdef SPPM_HOST_EMULATION  // tests/device_host only (g++ cannot assemble PTX); never defined in the library build
  r = dh_rcp_approx(x);
#else
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#endif
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
// One Newton step x = prevX - inv([dir, -Su, -Sv]) * (org + dir*t - S)  (Bezier.h:121-131).
// The inverse is cv::Matx33d::inv's closed form (adjugate / determinant, zeros when the
// determinant is exactly 0) held in registers.
template <bool FAST>
__device__ __forceinline__ void newton_step(d3 org, d3 dir, const PatchEval& e, double& t, double& u, double& v) {
  const d3 L = org + dir * t;
  const d3 r = L - e.S;
  const double a00 = dir.x, a01 = -e.Su.x, a02 = -e.Sv.x;
  const double a10 = dir.y, a11 = -e.Su.y, a12 = -e.Sv.y;
  const double a20 = dir.z, a21 = -e.Su.z, a22 = -e.Sv.z;
  if (FAST) {
    const double c00 = fma(a11, a22, -(a12 * a21)), c01 = fma(a02, a21, -(a01 * a22)), c02 = fma(a01, a12, -(a02 * a11));
    const double c10 = fma(a12, a20, -(a10 * a22)), c11 = fma(a00, a22, -(a02 * a20)), c12 = fma(a02, a10, -(a00 * a12));
    const double c20 = fma(a10, a21, -(a11 * a20)), c21 = fma(a01, a20, -(a00 * a21)), c22 = fma(a00, a11, -(a01 * a10));
    double d = fma(a02, c20, fma(a01, c10, a00 * c00));
    if (d != 0) {
      d = rcp_newton(d);
      t -= fma(c02, r.z, fma(c01, r.y, c00 * r.x)) * d;
      u -= fma(c12, r.z, fma(c11, r.y, c10 * r.x)) * d;
      v -= fma(c22, r.z, fma(c21, r.y, c20 * r.x)) * d;
    }
  } else {
    double d = a00 * (a11 * a22 - a21 * a12) - a01 * (a10 * a22 - a20 * a12) + a02 * (a10 * a21 - a20 * a11);
    double b00 = 0, b01 = 0, b02 = 0, b10 = 0, b11 = 0, b12 = 0, b20 = 0, b21 = 0, b22 = 0;
    if (d != 0) {
      d = 1 / d;
      b00 = (a11 * a22 - a12 * a21) * d;
      b01 = (a02 * a21 - a01 * a22) * d;
      b02 = (a01 * a12 - a02 * a11) * d;
      b10 = (a12 * a20 - a10 * a22) * d;
      b11 = (a00 * a22 - a02 * a20) * d;
      b12 = (a02 * a10 - a00 * a12) * d;
      b20 = (a10 * a21 - a11 * a20) * d;
      b21 = (a01 * a20 - a00 * a21) * d;
      b22 = (a00 * a11 - a01 * a10) * d;
    }
    // Matx * Vec: s = 0; s += m(i,k) * r(k)
    t = t - (((0.0 + b00 * r.x) + b01 * r.y) + b02 * r.z);
    u = u - (((0.0 + b10 * r.x) + b11 * r.y) + b12 * r.z);
    v = v - (((0.0 + b20 * r.x) + b21 * r.y) + b22 * r.z);
  }
}

// One restart of Bezier3::newton (Bezier.h:116-141): from (0,u0,v0) run `iters`
// iterations, keep the first strict minimum of t over the accepted iterates (Q4).
// MODE: 0 EXACT (P = control points), 1 FAST (P = monomial coefficients), 2 Bernstein tensor form + FMA
template <int MODE>
__device__ __forceinline__ PatchEval eval_mode(const double* __restrict__ P, double u, double v) {
  return MODE == 0 ? eval_exact(P, u, v) : (MODE == 1 ? eval_mono(P, u, v) : eval_fast(P, u, v));
}
template <int MODE>
__device__ __forceinline__ bool newton_restart(const double* __restrict__ P, d3 org, d3 dir, double u0, double v0,
                                               int iters, double& bt, double& bu, double& bv) {
  constexpr bool FAST = MODE != 0;
  double t = 0, u = u0, v = v0;
  bool found = false;
  bt = SPPM_MAX_NUM;
  PatchEval e = eval_mode<MODE>(P, u, v);
  for (int it = 0; it < iters; it++) {
    newton_step<FAST>(org, dir, e, t, u, v);
    e = eval_mode<MODE>(P, u, v);
    const d3 r = (org + dir * t) - e.S;
    const double res = ((0.0 + r.x * r.x) + r.y * r.y) + r.z * r.z;
    if (res < SPPM_M_EPS && u <= 1 && u >= 0 && v >= 0 && v <= 1) {
      if (t < bt) {
        bt = t;
        bu = u;
        bv = v;
        found = true;
      }
    }
  }
  return found;
}

// Process the warp's queued (ray, patch) items: n_items * restarts Newton tasks spread over
// the 32 lanes, per-item min-t written to ws->it_{t,u,v}.  seeds_uv (nullable) = injected
// seeds double[pair][restarts][2], pair = seed_pair_base + item.
__device__ inline void flush_items(const SceneView& sv, WarpScratch* ws, int n_items, const NewtonCfg& cfg,
                                   uint64_t pass, uint32_t domain, const double* seeds_uv, long long seed_pair_base,
                                   int lane) {
  for (int i = lane; i < n_items; i += 32) {
    ws->it_t[i] = SPPM_MAX_NUM;
    ws->it_u[i] = SPPM_MAX_NUM;
    ws->it_v[i] = SPPM_MAX_NUM;
  }
  __syncwarp();
  // 13-DOP line test (dop_line_miss): items whose solve cannot record anything keep MAX_NUM and
  // are not run; the live ones are listed in ws->live (in item order, which the arg-min relies on)
  int n_live = 0;
  for (int b0 = 0; b0 < n_items; b0 += 32) {
    const int i = b0 + lane;
    bool live = false;
    if (i < n_items) {
      const unsigned it = ws->item[i];
      const int owner = it & 31, gp = it >> 8;
      const d3 org = d3{ws->ray[0][owner], ws->ray[1][owner], ws->ray[2][owner]};
      const d3 dir = d3{ws->ray[3][owner], ws->ray[4][owner], ws->ray[5][owner]};
      live = !cfg.cull || !(dop_line_miss(sv.patchDop + (size_t)gp * 26, org, dir) || sub_boxes_miss(sv.patchSub + (size_t)gp * 96, org, dir));
    }
    const unsigned m = __ballot_sync(FULL, live);
    if (live) ws->live[n_live + __popc(m & ((1u << lane) - 1))] = (unsigned char)i;
    n_live += __popc(m);
  }
  if (lane == 0 && n_live < n_items) atomicAdd(sv.culled, (unsigned long long)(n_items - n_live));
  __syncwarp();
  const int R = cfg.restarts;
  const int total = n_live * R;
  for (int base = 0; base < total; base += 32) {
    const int k = base + lane;
    const bool valid = k < total;
    const int live_idx = valid ? k / R : -1;
    const int item = valid ? (int)ws->live[live_idx] : -1;
    double t = 0, u = 0, v = 0;
    bool found = false;
    if (valid) {
      const int r = k - live_idx * R;
      const unsigned it = ws->item[item];
      const int owner = it & 31, gp = it >> 8;
      const d3 org = d3{ws->ray[0][owner], ws->ray[1][owner], ws->ray[2][owner]};
      const d3 dir = d3{ws->ray[3][owner], ws->ray[4][owner], ws->ray[5][owner]};
      double u0, v0;
      if (seeds_uv) {
        const double* s = seeds_uv + ((size_t)(seed_pair_base + item) * R + r) * 2;
        u0 = s[0];
        v0 = s[1];
      } else {
        draw2(cfg.seed, pass, ws->id[owner], ws->node[owner], domain, (uint32_t)r, (uint32_t)gp, u0, v0);
      }
      const double* P = sv.patchP + (size_t)gp * 48;
      if (cfg.mode == SPPM_NEWTON_FAST) found = newton_restart<1>(sv.patchA + (size_t)gp * 48, org, dir, u0, v0, cfg.iters, t, u, v);
      else if (cfg.mode == SPPM_NEWTON_EXACT) found = newton_restart<0>(P, org, dir, u0, v0, cfg.iters, t, u, v);
      else found = newton_restart<2>(P, org, dir, u0, v0, cfg.iters, t, u, v);
    }
    // segmented arg-min over the (at most a few) items present in this round; restart order = lane order
    const int first = base / R;
    const int last = (min(base + 31, total - 1)) / R;
    for (int lx = first; lx <= last; lx++) {
      const int itx = (int)ws->live[lx];
      double tt = (found && item == itx) ? t : CUDART_INF;
      int ll = lane;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ot = __shfl_xor_sync(FULL, tt, off);
        const int ol = __shfl_xor_sync(FULL, ll, off);
        if (ot < tt || (ot == tt && ol < ll)) {
          tt = ot;
          ll = ol;
        }
      }
      if (tt < CUDART_INF) {  // warp-uniform
        const double wu = __shfl_sync(FULL, u, ll), wv = __shfl_sync(FULL, v, ll);
        if (lane == 0 && tt < ws->it_t[itx]) {
          ws->it_t[itx] = tt;
          ws->it_u[itx] = wu;
          ws->it_v[itx] = wv;
        }
      }
    }
    __syncwarp();
  }
}

struct Hit {  // Element.h:20-39 Collision, the fields the path reads
  double dist;
  int obj;    // -1 = NoneColl
  int patch;  // Bezier winner
  double u, v;
  bool inside;
};

// PlaneObj::GetIntersect (Obj.h:65-85): keep the hit if strictly nearer than best.
__device__ __forceinline__ void test_plane(const DevObject& ob, int a, d3 org, d3 dir, Hit& best) {
  const d3 norm = ld3(ob.norm);
  const double project = dot3(dir, norm);
  if (!is_zero(project)) {
    const double num = dot3(ld3(ob.p0) - org, norm);
    // dist = num / project is only needed when it can win.  It cannot when the signs differ
    // (dist <= 0), nor when an FP32 estimate of the quotient (relative error < 2^-20) already
    // exceeds the current best by more than that: both shortcuts skip exactly the cases the
    // reference's test rejects, so the result is unchanged and most FP64 divisions go away.
    if ((num > 0) != (project > 0)) return;
    const float est = __fdividef(__double2float_rn(num), __double2float_rn(project));
    if ((double)est > best.dist * (1.0 + 1e-5)) return;
    const double dist = num / project;
    if (dist > SPPM_M_EPS && dist < best.dist) {
      best.dist = dist;
      best.obj = a;
      best.inside = false;
    }
  }
}
// SphereObj::GetIntersect (Obj.h:111-139)
__device__ __forceinline__ void test_sphere(const DevObject& ob, int a, d3 org, d3 dir, Hit& best) {
  const d3 L = ld3(ob.p0) - org;
  const double project = dot3(L, dir);
  const double det2 = ob.radius * ob.radius - (module2(L) - project * project);
  if (!(det2 < SPPM_M_EPS)) {
    const double det = sqrt(det2);
    const double dist1 = project - det, dist2 = project + det;
    if (!(dist2 < SPPM_M_EPS)) {
      const double dist = (dist1 > SPPM_M_EPS) ? dist1 : dist2;
      if (dist < best.dist) {
        best.dist = dist;
        best.obj = a;
        best.inside = (dist1 < SPPM_M_EPS);
      }
    }
  }
}

// Scene::NearCollide (Scene.h:165-182) for the 32 rays of a warp.  All lanes must call
// (inactive lanes pass active=false).  Keeps the strictly nearest hit, first object wins ties.
__device__ inline Hit near_collide(const SceneView& sv, WarpScratch* ws, const NewtonCfg& cfg, bool active, d3 org,
                                   d3 dir, uint64_t pass, uint64_t id, uint32_t node, uint32_t domain,
                                   unsigned long long& solves) {
  const int lane = threadIdx.x & 31;
  Hit best;
  best.dist = SPPM_MAX_NUM;
  best.obj = -1;
  best.patch = -1;
  best.u = best.v = 0;
  best.inside = false;
  bool staged = false;
  for (int a = 0; a < sv.n_obj; a++) {
    const DevObject& ob = sv.obj[a];
    if (ob.kind == SPPM_OBJ_PLANE) {
      if (active) test_plane(ob, a, org, dir, best);
    } else if (ob.kind == SPPM_OBJ_SPHERE) {
      if (active) test_sphere(ob, a, org, dir, best);
    } else {  // Bezier3Obj::GetIntersect (Bezier.h:240-282)
      const bool in_box = active && aabb_intersect(ob.box, org, dir);
      if (__any_sync(FULL, in_box)) {
        if (!staged) {
          ws->ray[0][lane] = org.x; ws->ray[1][lane] = org.y; ws->ray[2][lane] = org.z;
          ws->ray[3][lane] = dir.x; ws->ray[4][lane] = dir.y; ws->ray[5][lane] = dir.z;
          ws->id[lane] = id;
          ws->node[lane] = node;
          staged = true;
          __syncwarp();
        }
        double pt = SPPM_MAX_NUM, pu = SPPM_MAX_NUM, pv = SPPM_MAX_NUM;  // para (Bezier.h:246)
        int index = -1;
        int n_items = 0;
        // Patch culling is cooperative: for every ray inside the object box (usually one or
        // two lanes of a photon warp) the 32 lanes test 32 patch boxes at once (Bezier.h:178);
        // items are queued in (ray lane, patch) order, which the per-lane filter below needs.
        // The queue is flushed (Newton tasks over all lanes) whenever it cannot take 32 more.
        unsigned boxmask = __ballot_sync(FULL, in_box);
        const int n_chunks = (ob.n_patch + 31) >> 5;
        int src = -1, ch = 0;  // ray (lane) being culled and its next 32-patch chunk; warp-uniform
        for (;;) {
          while (n_items <= QCAP - 32) {
            if (src < 0) {
              if (!boxmask) break;
              src = __ffs(boxmask) - 1;
              boxmask &= boxmask - 1;
              ch = 0;
            }
            const d3 so = d3{ws->ray[0][src], ws->ray[1][src], ws->ray[2][src]};
            const d3 sd = d3{ws->ray[3][src], ws->ray[4][src], ws->ray[5][src]};
            const int p = (ch << 5) + lane;
            const bool hitp = p < ob.n_patch && aabb_intersect(sv.patchBox + (size_t)(ob.first_patch + p) * 6, so, sd);
            const unsigned b = __ballot_sync(FULL, hitp);
            if (hitp) ws->item[n_items + __popc(b & ((1u << lane) - 1))] = (unsigned)src | ((unsigned)(ob.first_patch + p) << 8);
            n_items += __popc(b);
            if (++ch == n_chunks) src = -1;
          }
          if (n_items == 0) break;
          __syncwarp();
          flush_items(sv, ws, n_items, cfg, pass, domain, nullptr, 0, lane);
          if (lane == 0) solves += (unsigned long long)n_items;
          // per-lane filter in patch order (Bezier.h:249-261)
          for (int i = 0; i < n_items; i++) {
            const unsigned it = ws->item[i];
            if ((int)(it & 31) == lane) {
              const double qt = ws->it_t[i], qu = ws->it_u[i], qv = ws->it_v[i];
              if (qt < SPPM_M_EPS || qt > pt || qu < -SPPM_M_EPS || qu > 1 + SPPM_M_EPS || qv < -SPPM_M_EPS ||
                  qv > 1 + SPPM_M_EPS)
                continue;
              if (qt < pt) {
                pt = qt; pu = qu; pv = qv;
                index = (int)(it >> 8);
              }
            }
          }
          __syncwarp();
          n_items = 0;
        }
        if (index != -1 && pt < best.dist) {
          best.dist = pt;
          best.obj = a;
          best.patch = index;
          best.u = pu;
          best.v = pv;
          best.inside = false;
        }
      }
    }
  }
  return best;
}

struct Surface {  // what the callers need of the winning Collision
  d3 pos, n, color;
};
// Pos / N / color of the winning object (Obj.h:79-80,86-100; :131-137,140-153; Bezier.h:263-279).
__device__ inline Surface shade_hit(const SceneView& sv, const Hit& h, d3 org, d3 dir) {
  const DevObject& ob = sv.obj[h.obj];
  Surface s;
  s.pos = org + dir * h.dist;
  if (ob.kind == SPPM_OBJ_PLANE) {
    s.n = ld3(ob.norm);
    if (ob.tex_id < 0) {
      s.color = ld3(ob.color);
    } else {
      const double P[3] = {s.pos.x, s.pos.y, s.pos.z};
      const double v = 0.5 + (P[ob.vdex] - ob.p0[ob.vdex]) / ob.texUmod;
      const double u = 0.5 + (P[ob.udex] - ob.p0[ob.udex]) / ob.texVmod;
      s.color = color_uv(sv.tex[ob.tex_id], sv.lut255, u, v, sv.tex_lod);
    }
  } else if (ob.kind == SPPM_OBJ_SPHERE) {
    const d3 N = get_unit(s.pos - ld3(ob.p0));
    s.n = N;
    if (ob.tex_id < 0) {
      s.color = ld3(ob.color);
    } else {
      const double theta = acos(dot3(N, ld3(ob.texV)));
      double t = dot3(N, ld3(ob.texU)) / sin(theta);
      t = (t >= -1) ? t : -1;
      t = (t <= 1) ? t : 1;
      const double phi = acos(t);
      const double u = theta / SPPM_PI;
      double v = phi / (2 * SPPM_PI);
      v = (dot3(N, ld3(ob.texW)) < 0) ? (1 - v) : v;
      s.color = color_uv(sv.tex[ob.tex_id], sv.lut255, u, v, sv.tex_lod);
    }
  } else {
    const PatchEval e = eval_exact(sv.patchP + (size_t)h.patch * 48, h.u, h.v);
    d3 N = cross3(e.Su, e.Sv);
    if (dot3(N, dir) > 0) N = -N;
    s.n = normalize3(N);
    s.color = (ob.tex_id < 0) ? ld3(ob.color) : color_uv(sv.tex[ob.tex_id], sv.lut255, h.u, h.dist, sv.tex_lod);  // Q6
  }
  return s;
}

}  // namespace sppm
