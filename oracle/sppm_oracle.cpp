// sppm_oracle.cpp — CPU oracle: a line-by-line restatement of wondergo2017/raytrace3's
// render path.  TEST INFRASTRUCTURE ONLY (see sppm_oracle.h).  The reference has no tests of
// its own; this file is pinned bit-for-bit against the reference's own code compiled here
// (oracle/_ref, tests/test_ref_pin_cpu.py); all pins are listed in sppm_oracle.h / DESIGN.md.
//
// Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC (MSVC x64 /fp:precise does not
// contract to FMA either).  All arithmetic is double and keeps the reference's
// expression association; every function cites the file:line it follows (paths relative
// to /root/reference/raytracer/).
#include "sppm_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---- Vec3.h:6-14 constants -------------------------------------------------------
constexpr double M_EPS_ = 1e-4;
constexpr double PI_ = 3.1415926;
constexpr double PI_INV_ = 0.3183098862;
constexpr double MAX_NUM_ = 1e20;
constexpr double MIN_NUM_ = -1e20;

// ---- Vec3.h:28-155 (cv::Vec3d semantics: element-wise ops, sequential dot) --------
struct V3 {
  double v[3];
  V3() : v{0, 0, 0} {}
  V3(double a, double b, double c) : v{a, b, c} {}
  explicit V3(const double* p) : v{p[0], p[1], p[2]} {}
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};
inline V3 operator+(const V3& a, const V3& b) { return V3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline V3 operator-(const V3& a, const V3& b) { return V3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline V3 operator*(const V3& a, double s) { return V3(a[0] * s, a[1] * s, a[2] * s); }
inline V3 operator*(double s, const V3& a) { return V3(a[0] * s, a[1] * s, a[2] * s); }
inline V3 operator-(const V3& a) { return V3(-a[0], -a[1], -a[2]); }
inline V3 mul(const V3& a, const V3& b) { return V3(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }
// cv::Matx::dot: s = 0; s += a[i]*b[i] in order.
inline double dot(const V3& a, const V3& b) {
  double s = 0;
  for (int i = 0; i < 3; i++) s += a[i] * b[i];
  return s;
}
inline V3 cross(const V3& a, const V3& b) {
  return V3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
inline bool IsZero(double x) { return std::fabs(x) < M_EPS_; }  // Vec3.h:23
inline bool IsZeroV(const V3& a) {                               // Vec3.h:72-79: ANY channel
  for (int i = 0; i < 3; i++)
    if (IsZero(a[i])) return true;
  return false;
}
inline double Module2(const V3& a) {  // Vec3.h:44
  double res = 0;
  for (int i = 0; i < 3; i++) res += a[i] * a[i];
  return res;
}
inline double Module(const V3& a) { return std::sqrt(Module2(a)); }  // Vec3.h:45
inline V3 Normalize(const V3& a) {                                   // Vec3.h:48-55
  double mo = Module(a);
  if (!IsZero(mo)) {
    double mo_ = 1 / mo;
    return V3(a[0] * mo_, a[1] * mo_, a[2] * mo_);
  }
  return a;
}
inline V3 GetUnit(const V3& a) {  // Vec3.h:56
  double mo = Module(a);
  return IsZero(mo) ? a : a * (1.0 / mo);
}
inline V3 GetRefl(const V3& I, const V3& norm) {  // Vec3.h:80-83
  return I - (2 * dot(I, norm)) * norm;
}
inline V3 GetUnitRefl(const V3& I, const V3& norm) { return GetUnit(GetRefl(I, norm)); }  // Vec3.h:84
inline V3 GetAnormal(const V3& a) {                                                       // Vec3.h:85-89
  return (a[0] == 0 && a[1] == 0) ? V3(1, 0, 0) : GetUnit(V3(a[1], -a[0], 0));
}
V3 rotated(const V3& s, const V3& axis, double angle) {  // Vec3.h:99-115
  if (std::fabs(angle) < M_EPS_) return s;
  V3 ret;
  double cost = std::cos(angle);
  double sint = std::sin(angle);
  ret[0] += s[0] * (axis[0] * axis[0] + (1 - axis[0] * axis[0]) * cost);
  ret[0] += s[1] * (axis[0] * axis[1] * (1 - cost) - axis[2] * sint);
  ret[0] += s[2] * (axis[0] * axis[2] * (1 - cost) + axis[1] * sint);
  ret[1] += s[0] * (axis[1] * axis[0] * (1 - cost) + axis[2] * sint);
  ret[1] += s[1] * (axis[1] * axis[1] + (1 - axis[1] * axis[1]) * cost);
  ret[1] += s[2] * (axis[1] * axis[2] * (1 - cost) - axis[0] * sint);
  ret[2] += s[0] * (axis[2] * axis[0] * (1 - cost) - axis[1] * sint);
  ret[2] += s[1] * (axis[2] * axis[1] * (1 - cost) + axis[0] * sint);
  ret[2] += s[2] * (axis[2] * axis[2] + (1 - axis[2] * axis[2]) * cost);
  return ret;
}
V3 GetUnitRandRefl(const V3& norm, double U1, double U2) {  // Vec3.h:90-98
  double theta = std::acos(std::sqrt(U1));
  double phi = 2 * PI_ * U2;
  V3 ret = rotated(norm, GetAnormal(norm), theta);
  return GetUnit(rotated(ret, norm, phi));
}
V3 refracted(const V3& I, const V3& N, double n) {  // Vec3.h:124-134
  double cosI = -dot(N, I);
  double cosR2 = 1 - (1 - std::pow(cosI, 2)) * std::pow(n, 2);
  return (cosR2 > M_EPS_) ? I * n + N * (n * cosI - std::sqrt(cosR2)) : GetRefl(I, N);
}

// ---- counter RNG (replaces the irreproducible global cv::RNG, Vec3.h:5, Q15) -----
enum { P_EMIT = 0, P_BOUNCE = 1, P_JITTER = 4 };

inline void philox(const uint32_t c_in[4], const uint32_t k_in[2], uint32_t out[4]) {
  uint32_t c0 = c_in[0], c1 = c_in[1], c2 = c_in[2], c3 = c_in[3];
  uint32_t k0 = k_in[0], k1 = k_in[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
inline double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
inline void draw2(uint64_t seed, uint64_t pass, uint64_t id, uint32_t node, uint32_t purpose, uint32_t sub,
                  uint32_t patch, double out[2]) {
  uint32_t c[4], k[2], o[4];
  c[0] = (uint32_t)id;
  c[1] = (uint32_t)((id >> 32) & 0xffffu) | (node << 16);
  c[2] = (purpose & 0xfu) | ((sub & 0xfffu) << 4) | (patch << 16);
  c[3] = (uint32_t)pass;
  k[0] = (uint32_t)seed;
  k[1] = (uint32_t)(seed >> 32);
  philox(c, k, o);
  out[0] = u53(o[0], o[1]);
  out[1] = u53(o[2], o[3]);
}
// Optional log of the uniforms in the order the REFERENCE would consume them (used to feed the
// real reference code, oracle/_ref, with the same numbers).  kind: 0 single draw, 1 = u0 and
// 2 = v0 of a Newton restart (their consumption order in `Vec3 x(0, GetRandu(), GetRandu())`,
// Bezier.h:117, is compiler-dependent, Q16).  Not thread-safe: single-ray calls only.
std::vector<double>* g_draw_log = nullptr;
std::vector<signed char>* g_draw_kind = nullptr;
inline void log_draw(double u, int kind) {
  if (g_draw_log) {
    g_draw_log->push_back(u);
    g_draw_kind->push_back((signed char)kind);
  }
}
// cv::RNG::uniform(double a, double b) = u*(b-a)+a (Vec3.h:15-18).
inline double GetRandu(double u, double r1, double r2) { return u * (r2 - r1) + r1; }

// Vec3::GetUnitRandVec (Vec3.h:57-65) with addressed draws.
V3 GetUnitRandVec(uint64_t seed, uint64_t pass, uint64_t id, uint32_t purpose) {
  double x, y, z;
  uint32_t a = 0;
  do {
    double d0[2], d1[2];
    draw2(seed, pass, id, 0, purpose, 2 * a, 0, d0);
    draw2(seed, pass, id, 0, purpose, 2 * a + 1, 0, d1);
    x = GetRandu(d0[0], -1., 1.);
    y = GetRandu(d0[1], -1., 1.);
    z = GetRandu(d1[0], -1., 1.);
    log_draw(d0[0], 0);
    log_draw(d0[1], 0);
    log_draw(d1[0], 0);
    a++;
  } while ((x * x + y * y + z * z > 1 || x * x + y * y + z * z < M_EPS_) && a < 2000);
  return GetUnit(V3(x, y, z));
}

// ---- Element.h:61-72 Texture::colorUV ---------------------------------------------
V3 colorUV(const SppmTexture& t, double u, double v) {
  int rows = t.rows, cols = t.cols;
  double row = (u - std::floor(u)) * rows, col = (v - std::floor(v)) * cols;
  int r1 = (int)std::floor(row + 1e-10), r2 = r1 + 1;
  int c1 = (int)std::floor(col + 1e-10), c2 = c1 + 1;
  double detR = r2 - row, detC = c2 - col;
  r1 = (r1 >= 0) ? (r1 >= rows ? 0 : r1) : (rows - 1);
  c1 = (c1 >= 0) ? (c1 >= cols ? 0 : c1) : (cols - 1);
  r2 = (r2 < rows) ? r2 : 0;
  c2 = (c2 < cols) ? c2 : 0;
  auto px = [&](int r, int c) {  // Element.h:54-57: map(i,j)[co] / 255.0
    const uint8_t* p = t.bgr8 + ((size_t)r * cols + c) * 3;
    return V3(p[0] / 255.0, p[1] / 255.0, p[2] / 255.0);
  };
  return px(r1, c1) * detR * detC + px(r1, c2) * detR * (1 - detC) + px(r2, c1) * (1 - detR) * detC +
         px(r2, c2) * (1 - detR) * (1 - detC);
}

// ---- Bezier.h:7-57 AABBbox -----------------------------------------------------------
struct AABB {
  V3 pmin{MAX_NUM_, MAX_NUM_, MAX_NUM_}, pmax{MIN_NUM_, MIN_NUM_, MIN_NUM_};
  void addpoint(const V3& p) {
    for (int i = 0; i < 3; i++) {
      if (pmin[i] > p[i]) pmin[i] = p[i];
      if (pmax[i] < p[i]) pmax[i] = p[i];
    }
  }
  bool intersect(const V3& org, const V3& dir) const {  // Bezier.h:20-51 (entry faces only, Q5)
    double t = 0;
    V3 Colp;
    for (int i = 0; i < 3; i++) {
      if (!IsZero(dir[i])) {
        if (dir[i] > 0) t = (pmin[i] - org[i]) / dir[i];
        else t = (pmax[i] - org[i]) / dir[i];
        if (t > 0) {
          Colp = org + t * dir;
          bool flag = true;
          for (int j = 0; j < 3; j++)
            if (j != i)
              if (pmax[j] < Colp[j] || pmin[j] > Colp[j]) flag = false;
          if (flag) return true;
        }
      }
    }
    return false;
  }
};

// ---- Bezier.h:59-187 Bezier3 ---------------------------------------------------------
inline V3 evalCurvePoint(const V3 P[4], double t) {  // Bezier.h:69-76
  double c0 = (1 - t) * (1 - t) * (1 - t);
  double c1 = 3 * t * (1 - t) * (1 - t);
  double c2 = 3 * t * t * (1 - t);
  double c3 = t * t * t;
  return P[0] * c0 + P[1] * c1 + P[2] * c2 + P[3] * c3;
}
inline V3 evalCurveDeriv(const V3 P[4], double t) {  // Bezier.h:77-84
  double c0 = -3 * (1 - t) * (1 - t);
  double c1 = 3 * (1 - t) * (1 - t) - 6 * t * (1 - t);
  double c2 = 6 * t * (1 - t) - 3 * t * t;
  double c3 = 3 * t * t;
  return P[0] * c0 + P[1] * c1 + P[2] * c2 + P[3] * c3;
}
inline V3 evalPatchPoint(const V3 P[16], double u, double v) {  // Bezier.h:85-90
  V3 curve[4];
  for (int i = 0; i < 4; i++) curve[i] = evalCurvePoint(P + 4 * i, u);
  return evalCurvePoint(curve, v);
}
inline V3 evalPatchDerivU(const V3 P[16], double u, double v) {  // Bezier.h:91-105
  V3 Pv[4], curve[4];
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 4; j++) Pv[j] = P[i + 4 * j];
    curve[i] = evalCurvePoint(Pv, v);
  }
  return evalCurveDeriv(curve, u);
}
inline V3 evalPatchDerivV(const V3 P[16], double u, double v) {  // Bezier.h:106-111
  V3 curve[4];
  for (int i = 0; i < 4; i++) curve[i] = evalCurvePoint(P + 4 * i, u);
  return evalCurveDeriv(curve, v);
}
// cv::Matx33d::inv() default (DECOMP_LU) on a 3x3 Matx = closed-form adjugate; returns
// zeros when the determinant is exactly 0 (OpenCV operations.hpp Matx_FastInvOp<_Tp,3,3>,
// Matx_DetOp<_Tp,3>).  m is row-major m[3*i+j].
void inv3(const double a[9], double b[9]) {
#define A(i, j) a[3 * (i) + (j)]
#define B(i, j) b[3 * (i) + (j)]
  double d = A(0, 0) * (A(1, 1) * A(2, 2) - A(2, 1) * A(1, 2)) - A(0, 1) * (A(1, 0) * A(2, 2) - A(2, 0) * A(1, 2)) +
             A(0, 2) * (A(1, 0) * A(2, 1) - A(2, 0) * A(1, 1));
  if (d == 0) {
    for (int i = 0; i < 9; i++) b[i] = 0;
    return;
  }
  d = 1 / d;
  B(0, 0) = (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) * d;
  B(0, 1) = (A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2)) * d;
  B(0, 2) = (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1)) * d;
  B(1, 0) = (A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2)) * d;
  B(1, 1) = (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0)) * d;
  B(1, 2) = (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2)) * d;
  B(2, 0) = (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0)) * d;
  B(2, 1) = (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1)) * d;
  B(2, 2) = (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0)) * d;
#undef A
#undef B
}
const V3 MAX_VEC3_(MAX_NUM_, MAX_NUM_, MAX_NUM_);

// Bezier3::newton (Bezier.h:112-159).  seed(o, uv) supplies (GetRandu(0,1), GetRandu(0,1))
// of restart o as (u0, v0).  Every accepted iterate is recorded (no break, Q4); the
// answer is the first strict minimum of t in (restart, iteration) order.
template <class SeedFn>
V3 newton(const V3 P[16], const V3& org, const V3& dir, int restarts, int maxIter, SeedFn seed) {
  bool any = false;
  V3 best;
  double mindist = MAX_NUM_;
  for (int o = 0; o < restarts; o++) {
    double uv[2];
    seed(o, uv);
    V3 x(0, uv[0], uv[1]), prevX(-1, -1, -1);
    int iter = 0;
    while (iter < maxIter) {
      prevX = x;
      V3 L = org + dir * prevX[0];
      V3 P_ = evalPatchPoint(P, prevX[1], prevX[2]);
      V3 derivU = evalPatchDerivU(P, prevX[1], prevX[2]);
      V3 derivV = evalPatchDerivV(P, prevX[1], prevX[2]);
      double m[9], mi[9];
      for (int i = 0; i < 3; i++) m[3 * i + 0] = dir[i];
      for (int i = 0; i < 3; i++) m[3 * i + 1] = -derivU[i];
      for (int i = 0; i < 3; i++) m[3 * i + 2] = -derivV[i];
      inv3(m, mi);
      V3 r = L - P_;
      V3 step;  // Matx*Vec: s = 0; s += m(i,k)*r(k)
      for (int i = 0; i < 3; i++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += mi[3 * i + k] * r[k];
        step[i] = s;
      }
      x = prevX - step;
      iter++;
      L = org + dir * x[0];
      P_ = evalPatchPoint(P, x[1], x[2]);
      if (Module2(L - P_) < M_EPS_ && x[1] <= 1 && x[1] >= 0 && x[2] >= 0 && x[2] <= 1) {
        // xvec.push_back(x); later: first strict minimum over xvec (Bezier.h:146-158).
        // MAX_NUM start: an accepted iterate with t >= 1e20 would index -1 in the
        // reference (UB); it cannot be accepted with a finite residual in practice.
        if (x[0] < mindist) {
          mindist = x[0];
          best = x;
          any = true;
        }
      }
    }
  }
  return any ? best : MAX_VEC3_;
}

// ---- flattened scene ------------------------------------------------------------------
struct Obj {
  SppmObject o;
  bool isDiff, isRefl, isRefr;      // Element.h:16-18
  double diffr, reflr, refrr, allr;  // Obj.h:12-15
  AABB box;                          // Bezier3Obj::aabb (Bezier.h:194)
};
struct SceneC {
  std::vector<Obj> objs;
  std::vector<V3> P;          // 16 per patch
  std::vector<AABB> pbox;     // per patch (Bezier.h:165-175)
  std::vector<SppmLight> lights;
  std::vector<SppmTexture> tex;
  explicit SceneC(const SppmSceneDesc* s) {
    for (int i = 0; i < s->n_patches; i++) {
      AABB b;
      for (int k = 0; k < 16; k++) {
        V3 p(s->patches + ((size_t)i * 16 + k) * 3);
        P.push_back(p);
        b.addpoint(p);
      }
      pbox.push_back(b);
    }
    for (int i = 0; i < s->n_objects; i++) {
      Obj ob;
      ob.o = s->objects[i];
      V3 d(ob.o.mat.diff), r(ob.o.mat.refl), t(ob.o.mat.refr);
      ob.isDiff = !IsZeroV(d);
      ob.isRefl = !IsZeroV(r);
      ob.isRefr = !IsZeroV(t);
      ob.diffr = (d[0] + d[1] + d[2]) / 3;  // Vec3.h:116-119 GetPower
      ob.reflr = (r[0] + r[1] + r[2]) / 3;
      ob.refrr = (t[0] + t[1] + t[2]) / 3;
      ob.allr = ob.diffr + ob.reflr + ob.refrr;
      if (ob.o.kind == SPPM_OBJ_BEZIER)
        for (int p = ob.o.first_patch; p < ob.o.first_patch + ob.o.n_patch; p++)
          for (int k = 0; k < 16; k++) ob.box.addpoint(P[(size_t)p * 16 + k]);
      objs.push_back(ob);
    }
    for (int i = 0; i < s->n_lights; i++) lights.push_back(s->lights[i]);
    for (int i = 0; i < s->n_textures; i++) tex.push_back(s->textures[i]);
  }
};

struct Coll {  // Element.h:20-39
  V3 Pos, N, From, color;
  double dist = 0;
  int obj = -1;
  bool inside = false;
  int patch = -1;
  double u = 0, v = 0;
  bool IsValid() const { return obj >= 0; }
  bool IsFront() const { return !inside; }
};

struct RngAddr {
  uint64_t seed, pass, id;
  uint32_t node, domain;
};

// PlaneObj::GetColor (Obj.h:86-100)
V3 planeColor(const SceneC& sc, const Obj& ob, const V3& P) {
  if (ob.o.tex_id < 0) return V3(ob.o.color);
  int ndir = 0;
  for (int i = 0; i < 3; i++)
    if (ob.o.norm[i] != 0) ndir = i;
  int vdex = (ndir + 2) % 3;
  int udex = (ndir + 1) % 3;
  double v = 0.5 + (P[vdex] - ob.o.p0[vdex]) / Module(V3(ob.o.texU));
  double u = 0.5 + (P[udex] - ob.o.p0[udex]) / Module(V3(ob.o.texV));
  return colorUV(sc.tex[ob.o.tex_id], u, v);
}
// SphereObj::GetColor (Obj.h:140-153)
V3 sphereColor(const SceneC& sc, const Obj& ob, const V3& P) {
  if (ob.o.tex_id < 0) return V3(ob.o.color);
  V3 org(ob.o.p0), texU(ob.o.texU), texV(ob.o.texV);
  V3 N = GetUnit(P - org);
  double theta = std::acos(dot(N, texV));
  double t = dot(N, texU) / std::sin(theta);
  t = (t >= -1) ? t : -1;
  t = (t <= 1) ? t : 1;
  double phi = std::acos(t);
  double u = theta / PI_, v = phi / (2 * PI_);
  v = (dot(N, cross(texU, texV)) < 0) ? (1 - v) : v;
  return colorUV(sc.tex[ob.o.tex_id], u, v);
}

// PlaneObj::GetIntersect (Obj.h:65-85)
Coll planeIntersect(const SceneC& sc, int idx, const V3& org, const V3& dir) {
  const Obj& ob = sc.objs[idx];
  V3 norm(ob.o.norm), P0(ob.o.p0);
  double project = dot(dir, norm);
  if (IsZero(project)) return Coll();
  double dist = dot(P0 - org, norm) / project;
  if (dist > M_EPS_) {
    Coll c;
    c.Pos = org + dir * dist;
    c.From = dir;
    c.dist = dist;
    c.obj = idx;
    c.N = norm;
    c.color = planeColor(sc, ob, c.Pos);
    return c;
  }
  return Coll();
}
// SphereObj::GetIntersect (Obj.h:111-139)
Coll sphereIntersect(const SceneC& sc, int idx, const V3& o, const V3& d) {
  const Obj& ob = sc.objs[idx];
  V3 org(ob.o.p0);
  double radius = ob.o.radius;
  V3 L = org - o;
  double project = dot(L, d);
  double det2 = radius * radius - (Module2(L) - project * project);
  if (det2 < M_EPS_) return Coll();
  double det = std::sqrt(det2);
  double dist1 = project - det, dist2 = project + det;
  if (dist2 < M_EPS_) return Coll();
  double dist = (dist1 > M_EPS_) ? dist1 : dist2;
  Coll c;
  c.Pos = o + d * dist;
  c.From = d;
  c.N = GetUnit(c.Pos - org);
  c.obj = idx;
  c.dist = dist;
  c.inside = (dist1 < M_EPS_);
  c.color = sphereColor(sc, ob, c.Pos);
  return c;
}
// Bezier3Obj::GetIntersect (Bezier.h:240-282) incl. Bezier3::intersect (:176-186)
Coll bezierIntersect(const SceneC& sc, int idx, const V3& o, const V3& d, const SppmParams& prm, const RngAddr& ra,
                     uint64_t* solves) {
  const Obj& ob = sc.objs[idx];
  if (!ob.box.intersect(o, d)) return Coll();
  V3 para = MAX_VEC3_;
  int index = -1;
  for (int i = 0; i < ob.o.n_patch; i++) {
    int gp = ob.o.first_patch + i;
    V3 paratmp = MAX_VEC3_;
    if (sc.pbox[gp].intersect(o, d)) {
      if (solves) (*solves)++;
      paratmp = newton(&sc.P[(size_t)gp * 16], o, d, prm.newton_restarts, prm.newton_iters, [&](int r, double uv[2]) {
        draw2(ra.seed, ra.pass, ra.id, ra.node, ra.domain, (uint32_t)r, (uint32_t)gp, uv);
        log_draw(uv[0], 1);
        log_draw(uv[1], 2);
      });
    }
    if (paratmp[0] < M_EPS_ || paratmp[0] > para[0] || paratmp[1] < -M_EPS_ || paratmp[1] > 1 + M_EPS_ ||
        paratmp[2] < -M_EPS_ || paratmp[2] > 1 + M_EPS_)
      continue;
    if (paratmp[0] < para[0]) {
      para = paratmp;
      index = i;
    }
  }
  if (index == -1) return Coll();
  int gp = ob.o.first_patch + index;
  Coll c;
  c.Pos = o + d * para[0];
  c.From = d;
  V3 U = evalPatchDerivU(&sc.P[(size_t)gp * 16], para[1], para[2]);
  V3 V = evalPatchDerivV(&sc.P[(size_t)gp * 16], para[1], para[2]);
  V3 N = cross(U, V);
  if (dot(N, d) > 0) N = -N;
  N = Normalize(N);
  c.N = N;
  c.obj = idx;
  c.dist = para[0];
  c.inside = false;
  c.patch = gp;
  c.u = para[1];
  c.v = para[2];
  // Bezier3Obj::GetColor(u, t) (Bezier.h:278, 283-287; Q6: second coordinate is the ray distance)
  c.color = (ob.o.tex_id < 0) ? V3(ob.o.color) : colorUV(sc.tex[ob.o.tex_id], para[1], para[0]);
  return c;
}
// Scene::NearCollide (Scene.h:165-182)
Coll NearCollide(const SceneC& sc, const V3& org, const V3& dir, const SppmParams& prm, const RngAddr& ra,
                 uint64_t* solves) {
  double mindist = MAX_NUM_;
  Coll ctmp;
  for (int a = 0; a < (int)sc.objs.size(); a++) {
    Coll t;
    switch (sc.objs[a].o.kind) {
      case SPPM_OBJ_PLANE: t = planeIntersect(sc, a, org, dir); break;
      case SPPM_OBJ_SPHERE: t = sphereIntersect(sc, a, org, dir); break;
      default: t = bezierIntersect(sc, a, org, dir, prm, ra, solves); break;
    }
    if (t.IsValid() && t.dist < mindist) {
      mindist = t.dist;
      ctmp = t;
    }
  }
  return ctmp;
}

// ---- exact kd-tree (stands in for flann::KDTreeSingleIndex, Raytracer.h:92-96) ------
struct KdTree {
  struct Node {
    int left = -1, right = -1, dim = -1, lo = 0, hi = 0;
    double split = 0;
  };
  std::vector<Node> nodes;
  std::vector<int> idx;
  const double* pts = nullptr;
  int build(int lo, int hi) {
    int id = (int)nodes.size();
    nodes.emplace_back();
    nodes[id].lo = lo;
    nodes[id].hi = hi;
    if (hi - lo <= 10) return id;
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int i = lo; i < hi; i++)
      for (int k = 0; k < 3; k++) {
        double c = pts[3 * (size_t)idx[i] + k];
        mn[k] = std::min(mn[k], c);
        mx[k] = std::max(mx[k], c);
      }
    int dim = 0;
    for (int k = 1; k < 3; k++)
      if (mx[k] - mn[k] > mx[dim] - mn[dim]) dim = k;
    if (!(mx[dim] > mn[dim])) return id;  // all points identical
    int mid = (lo + hi) / 2;
    std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi,
                     [&](int a, int b) { return pts[3 * (size_t)a + dim] < pts[3 * (size_t)b + dim]; });
    nodes[id].dim = dim;
    nodes[id].split = pts[3 * (size_t)idx[mid] + dim];
    int l = build(lo, mid);
    int r = build(mid, hi);
    nodes[id].left = l;
    nodes[id].right = r;
    return id;
  }
  template <class F>
  void radius(int id, const double q[3], double r2, F&& f) const {
    const Node& nd = nodes[id];
    if (nd.dim < 0) {
      for (int i = nd.lo; i < nd.hi; i++) f(idx[i]);
      return;
    }
    double d = q[nd.dim] - nd.split;
    // left holds coords <= split, right coords >= split; prune only when the plane is
    // strictly farther than the radius (exact search, eps = 0)
    if (d <= 0) {
      radius(nd.left, q, r2, f);
      if (d * d <= r2) radius(nd.right, q, r2, f);
    } else {
      radius(nd.right, q, r2, f);
      if (d * d <= r2) radius(nd.left, q, r2, f);
    }
  }
};

struct HitPoints {  // the live fields of struct HitPoint (Raytracer.h:47-58)
  int64_t n;
  const double *pos, *nrm, *wgt, *r2;
};

// Raytracer.h:139-159 for one deposit; candidates from the kd-tree or all hit points.
inline void gatherOne(const HitPoints& hp, const KdTree* tree, const V3& x, const V3& cN, const V3& flux,
                      double searchR2, int32_t* newN, double* tao, uint64_t* cand, uint64_t* acc) {
  auto test = [&](int e2) {
    V3 p(hp.pos + 3 * (size_t)e2);
    V3 v = p - x;
    double vv = dot(v, v);
    if (!(vv < searchR2)) return;  // flann radiusSearch: squared radius, strict '<' (Q3)
    if (cand) (*cand)++;
    V3 n(hp.nrm + 3 * (size_t)e2);
    if ((dot(n, cN) > 1e-3) && (vv <= hp.r2[e2])) {
      newN[e2]++;
      V3 w(hp.wgt + 3 * (size_t)e2);
      V3 add = mul(w, flux) * PI_INV_;
      for (int k = 0; k < 3; k++) tao[3 * (size_t)e2 + k] += add[k];
      if (acc) (*acc)++;
    }
  };
  if (tree) {
    double q[3] = {x[0], x[1], x[2]};
    tree->radius(0, q, searchR2, test);
  } else {
    for (int64_t i = 0; i < hp.n; i++) test((int)i);
  }
}

struct DepositSink {
  int64_t cap = 0, n = 0;
  double *x = nullptr, *nn = nullptr, *flux = nullptr;
  uint64_t* photon = nullptr;
  int32_t* bounce = nullptr;
};

// RayTracer::PhotonTrace (Raytracer.h:117-209), tail recursion unrolled into a loop.
void PhotonTrace(const SceneC& sc, const SppmParams& prm, V3 org, V3 dir, V3 flux, uint64_t pass, uint64_t id,
                 const HitPoints& hp, const KdTree* tree, int32_t* newN, double* tao, DepositSink* sink,
                 uint64_t* counters) {
  int dep = 0;
  for (;;) {
    RngAddr ra{prm.seed, pass, id, (uint32_t)dep, SPPM_DOMAIN_PHOTON};
    if (counters) counters[0]++;
    Coll c1 = NearCollide(sc, org, dir, prm, ra, counters ? counters + 1 : nullptr);
    if (!c1.IsValid()) return;
    if (dep > prm.max_dep) return;
    int node = dep;
    dep++;
    const Obj& obj = sc.objs[c1.obj];
    if (obj.isDiff) {
      if (counters) counters[2]++;
      if (sink) {
        if (sink->n < sink->cap) {
          for (int k = 0; k < 3; k++) {
            sink->x[3 * sink->n + k] = c1.Pos[k];
            sink->nn[3 * sink->n + k] = c1.N[k];
            sink->flux[3 * sink->n + k] = flux[k];
          }
          if (sink->photon) sink->photon[sink->n] = id;
          if (sink->bounce) sink->bounce[sink->n] = node;
        }
        sink->n++;
      }
      if (hp.n > 0) gatherOne(hp, tree, c1.Pos, c1.N, flux, prm.init_r2, newN, tao, nullptr, nullptr);
    }
    // Obj::Roulette (Obj.h:30-45)
    double d0[2], d1[2];
    draw2(prm.seed, pass, id, (uint32_t)node, P_BOUNCE, 0, 0, d0);
    double r = GetRandu(d0[0], 0., obj.allr);
    log_draw(d0[0], 0);
    V3 nflux = mul(c1.color, flux);
    if (obj.diffr > r) {  // DIFF (Raytracer.h:163-166)
      draw2(prm.seed, pass, id, (uint32_t)node, P_BOUNCE, 1, 0, d1);
      log_draw(d0[1], 0);
      log_draw(d1[0], 0);
      dir = GetUnitRandRefl(c1.N, d0[1], d1[0]);
    } else if (obj.diffr + obj.reflr > r) {  // REFL (:177-181)
      dir = GetUnitRefl(c1.From, c1.N);
    } else {  // REFR (:182-190)
      double refrn = (c1.IsFront() ? 1 / obj.o.mat.refrn : obj.o.mat.refrn);
      V3 refr = refracted(c1.From, (c1.IsFront() ? c1.N : -c1.N), refrn);
      dir = Normalize(refr);
    }
    org = c1.Pos;
    flux = nflux;
  }
}

struct HpSink {
  int64_t cap, n;
  double *pos, *nrm, *wgt, *r2;
  int32_t* pix;
  uint32_t* node;
  double init_r2;
};

// RayTracer::ViewTrace (Raytracer.h:296-357); node: root 1, refl child 2n, refr child 2n+1.
void ViewTrace(const SceneC& sc, const SppmParams& prm, const V3& org, const V3& dir, int dep, const V3& wgt, int x,
               int y, uint64_t pass, uint64_t pixid, uint32_t node, HpSink& out, uint64_t* counters) {
  RngAddr ra{prm.seed, pass, pixid, node, SPPM_DOMAIN_EYE};
  if (counters) counters[0]++;
  Coll c1 = NearCollide(sc, org, dir, prm, ra, counters ? counters + 1 : nullptr);
  if (!c1.IsValid()) return;
  if (dep > prm.max_dep) return;
  dep++;
  const Obj& obj = sc.objs[c1.obj];
  if (obj.isDiff) {
    V3 w = mul(mul(c1.color, wgt), V3(obj.o.mat.diff));
    if (out.n < out.cap) {
      for (int k = 0; k < 3; k++) {
        out.pos[3 * out.n + k] = c1.Pos[k];
        out.nrm[3 * out.n + k] = c1.N[k];
        out.wgt[3 * out.n + k] = w[k];
      }
      out.r2[out.n] = out.init_r2;
      out.pix[2 * out.n] = x;
      out.pix[2 * out.n + 1] = y;
      if (out.node) out.node[out.n] = node;
    }
    out.n++;
  }
  if (obj.isRefl) {
    V3 w = mul(mul(c1.color, wgt), V3(obj.o.mat.refl));
    ViewTrace(sc, prm, c1.Pos, GetUnitRefl(c1.From, c1.N), dep, w, x, y, pass, pixid, node * 2, out, counters);
  }
  if (obj.isRefr) {
    double refrn = (c1.IsFront() ? 1 / obj.o.mat.refrn : obj.o.mat.refrn);
    V3 refr = Normalize(refracted(c1.From, (c1.IsFront() ? c1.N : -c1.N), refrn));
    V3 w = mul(mul(c1.color, wgt), V3(obj.o.mat.refr));
    ViewTrace(sc, prm, c1.Pos, refr, dep, w, x, y, pass, pixid, node * 2 + 1, out, counters);
  }
}

void cameraLookAt(const V3& pos, const V3& look, int w, int h, SppmCameraDesc* cam) {  // Camera.h:32-54
  const V3 up(0, 0, 1);
  V3 dir = GetUnit(look - pos);
  V3 du = Normalize(cross(dir, up));
  V3 dv = Normalize(-cross(dir, du));
  double fov = 50. * PI_ / 180.;
  dir = dir * (.5 / std::tan(fov / 2.));
  for (int k = 0; k < 3; k++) {
    cam->pos[k] = pos[k];
    cam->dir[k] = dir[k];
    cam->du[k] = du[k];
    cam->dv[k] = dv[k];
  }
  cam->w = w;
  cam->h = h;
}
void cameraEmit(const SppmCameraDesc* cam, int y, int x, V3& org, V3& dir) {  // Camera.h:18-22
  V3 du(cam->du), dv(cam->dv), cd(cam->dir);
  V3 d = du * ((x + 0.5) / cam->w - 0.5) + dv * ((y + 0.5) / cam->h - 0.5) + cd;
  org = V3(cam->pos);
  dir = GetUnit(d);
}

void updateAll(int64_t n, int mode, double alpha, int32_t* newN, int32_t* N, double* r2, double* tao) {
  for (int64_t i = 0; i < n; i++) {  // HitPoint::update (Raytracer.h:69-79)
    if (mode == SPPM_UPDATE_REFERENCE) {
      if (N[i] <= 0 || newN[i] <= 0) continue;
    } else {
      if (newN[i] <= 0) continue;
    }
    double k = (N[i] + alpha * newN[i]) / (N[i] + newN[i]);
    r2[i] *= k;
    for (int c = 0; c < 3; c++) tao[3 * i + c] *= k;
    N[i] += alpha * newN[i];  // int += double: evaluated in double, truncated
    newN[i] = 0;
  }
}
void resolveFilm(int64_t n, const int32_t* pix, const double* r2, const double* tao, int cnt, int onetime, int w, int h,
                 double* film) {  // Raytracer.h:281-293
  std::memset(film, 0, sizeof(double) * 3 * (size_t)w * h);
  for (int64_t i = 0; i < n; i++) {
    int x = pix[2 * i], y = pix[2 * i + 1];
    double s = 1.0 / PI_ / (r2[i] * cnt * onetime);
    for (int c = 0; c < 3; c++) film[((size_t)x * w + y) * 3 + c] = film[((size_t)x * w + y) * 3 + c] + tao[3 * i + c] * s;
  }
}
inline int toInt(double x) { return int(std::pow(1 - std::exp(-x), 1 / 2.2) * 255 + .5); }  // Raytracer.h:24-26

}  // namespace

// ======================================================================================
extern "C" {

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox(ctr, key, out); }
void orc_draw2(uint64_t seed, uint64_t pass, uint64_t id, uint32_t node, uint32_t purpose, uint32_t sub, uint32_t patch,
               double out[2]) {
  draw2(seed, pass, id, node, purpose, sub, patch, out);
}
void orc_inv3(const double m[9], double out[9]) { inv3(m, out); }
int orc_aabb_intersect(const double box[6], const double org[3], const double dir[3]) {
  AABB b;
  b.pmin = V3(box);
  b.pmax = V3(box + 3);
  return b.intersect(V3(org), V3(dir)) ? 1 : 0;
}
void orc_bezier_eval(const double P[48], double u, double v, double S[3], double Su[3], double Sv[3]) {
  V3 p[16];
  for (int i = 0; i < 16; i++) p[i] = V3(P + 3 * i);
  V3 s = evalPatchPoint(p, u, v), su = evalPatchDerivU(p, u, v), sv = evalPatchDerivV(p, u, v);
  for (int k = 0; k < 3; k++) {
    S[k] = s[k];
    Su[k] = su[k];
    Sv[k] = sv[k];
  }
}
void orc_newton(const double P[48], const double org[3], const double dir[3], const double* seeds_uv, int restarts,
                int iters, double tuv[3]) {
  V3 p[16];
  for (int i = 0; i < 16; i++) p[i] = V3(P + 3 * i);
  V3 r = newton(p, V3(org), V3(dir), restarts, iters, [&](int o, double uv[2]) {
    uv[0] = seeds_uv[2 * o];
    uv[1] = seeds_uv[2 * o + 1];
  });
  for (int k = 0; k < 3; k++) tuv[k] = r[k];
}
void orc_color_uv(const SppmTexture* tex, double u, double v, double out[3]) {
  V3 c = colorUV(*tex, u, v);
  for (int k = 0; k < 3; k++) out[k] = c[k];
}
void orc_unit_rand_refl(const double norm[3], double u1, double u2, double out[3]) {
  V3 r = GetUnitRandRefl(V3(norm), u1, u2);
  for (int k = 0; k < 3; k++) out[k] = r[k];
}
void orc_refracted(const double I[3], const double N[3], double n, double out[3]) {
  V3 r = Normalize(refracted(V3(I), V3(N), n));
  for (int k = 0; k < 3; k++) out[k] = r[k];
}
void orc_camera_lookat(const double pos[3], const double look[3], int w, int h, SppmCameraDesc* cam) {
  cameraLookAt(V3(pos), V3(look), w, h, cam);
}
void orc_camera_emit(const SppmCameraDesc* cam, int y, int x, double ray6[6]) {
  V3 o, d;
  cameraEmit(cam, y, x, o, d);
  for (int k = 0; k < 3; k++) {
    ray6[k] = o[k];
    ray6[3 + k] = d[k];
  }
}

static void fillHit(const Coll& c, SppmHit* h) {
  h->t = c.IsValid() ? c.dist : MAX_NUM_;
  for (int k = 0; k < 3; k++) {
    h->pos[k] = c.Pos[k];
    h->n[k] = c.N[k];
    h->color[k] = c.color[k];
  }
  h->u = c.u;
  h->v = c.v;
  h->obj = c.obj;
  h->patch = c.patch;
  h->inside = c.inside ? 1 : 0;
  h->pad = 0;
}

void orc_intersect_batch(const SppmSceneDesc* scene, const SppmParams* prm, const double* rays6, int64_t n,
                         const uint64_t* ids, uint64_t pass_id, int32_t node, int32_t domain, SppmHit* out) {
  SceneC sc(scene);
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < n; i++) {
    RngAddr ra{prm->seed, pass_id, ids ? ids[i] : (uint64_t)i, (uint32_t)node, (uint32_t)domain};
    Coll c = NearCollide(sc, V3(rays6 + 6 * i), V3(rays6 + 6 * i + 3), *prm, ra, nullptr);
    fillHit(c, out + i);
  }
}
void orc_bezier_patch_batch(const SppmSceneDesc* scene, const SppmParams* prm, const double* rays6,
                            const int32_t* patch_idx, int64_t n, const double* seeds_uv, double* tuv_out) {
  SceneC sc(scene);
  int R = prm->newton_restarts;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < n; i++) {
    int gp = patch_idx[i];
    V3 r;
    if (seeds_uv) {
      const double* s = seeds_uv + (size_t)i * R * 2;
      r = newton(&sc.P[(size_t)gp * 16], V3(rays6 + 6 * i), V3(rays6 + 6 * i + 3), R, prm->newton_iters,
                 [&](int o, double uv[2]) {
                   uv[0] = s[2 * o];
                   uv[1] = s[2 * o + 1];
                 });
    } else {
      r = newton(&sc.P[(size_t)gp * 16], V3(rays6 + 6 * i), V3(rays6 + 6 * i + 3), R, prm->newton_iters,
                 [&](int o, double uv[2]) {
                   draw2(prm->seed, 0, (uint64_t)i, 0, SPPM_DOMAIN_TEST, (uint32_t)o, (uint32_t)gp, uv);
                 });
    }
    for (int k = 0; k < 3; k++) tuv_out[3 * i + k] = r[k];
  }
}

int64_t orc_eye_pass(const SppmSceneDesc* scene, const SppmParams* prm, const SppmCameraDesc* cam, uint64_t pass_id,
                     int64_t cap, double* pos3, double* n3, double* wgt3, double* r2, int32_t* pix, uint32_t* node) {
  SceneC sc(scene);
  // GetHitPoint (Raytracer.h:105-113).  Rows are traced in parallel into per-row buffers and
  // concatenated in row order, so Hvec's order is the reference's (pixels are independent).
  struct Row {
    std::vector<double> pos, nrm, wgt, r2;
    std::vector<int32_t> pix;
    std::vector<uint32_t> node;
  };
  std::vector<Row> rows(cam->h);
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < cam->h; i++) {
    Row& R = rows[i];
    const int64_t rcap = (int64_t)cam->w * 64;
    R.pos.resize(3 * rcap); R.nrm.resize(3 * rcap); R.wgt.resize(3 * rcap); R.r2.resize(rcap);
    R.pix.resize(2 * rcap); R.node.resize(rcap);
    HpSink out{rcap, 0, R.pos.data(), R.nrm.data(), R.wgt.data(), R.r2.data(), R.pix.data(), R.node.data(), prm->init_r2};
    for (int j = 0; j < cam->w; j++) {
      V3 o, d;
      cameraEmit(cam, i, j, o, d);
      ViewTrace(sc, *prm, o, d, 0, V3(1, 1, 1), i, j, pass_id, (uint64_t)i * cam->w + j, 1, out, nullptr);
    }
    const int64_t n = std::min(out.n, rcap);
    R.pos.resize(3 * n); R.nrm.resize(3 * n); R.wgt.resize(3 * n); R.r2.resize(n); R.pix.resize(2 * n); R.node.resize(n);
  }
  int64_t total = 0;
  for (int i = 0; i < cam->h; i++) {
    const Row& R = rows[i];
    const int64_t n = (int64_t)R.r2.size();
    for (int64_t k = 0; k < n; k++, total++) {
      if (total >= cap) continue;
      for (int c = 0; c < 3; c++) {
        pos3[3 * total + c] = R.pos[3 * k + c];
        n3[3 * total + c] = R.nrm[3 * k + c];
        wgt3[3 * total + c] = R.wgt[3 * k + c];
      }
      r2[total] = R.r2[k];
      pix[2 * total] = R.pix[2 * k];
      pix[2 * total + 1] = R.pix[2 * k + 1];
      if (node) node[total] = R.node[k];
    }
  }
  return total;
}

// ViewTrace of one pixel (Raytracer.h:111), single-threaded (usable with the draw log).
int64_t orc_view_trace_pixel(const SppmSceneDesc* scene, const SppmParams* prm, const SppmCameraDesc* cam, uint64_t pass_id,
                             int i, int j, int64_t cap, double* pos3, double* n3, double* wgt3, double* r2, int32_t* pix) {
  SceneC sc(scene);
  HpSink out{cap, 0, pos3, n3, wgt3, r2, pix, nullptr, prm->init_r2};
  V3 o, d;
  cameraEmit(cam, i, j, o, d);
  ViewTrace(sc, *prm, o, d, 0, V3(1, 1, 1), i, j, pass_id, (uint64_t)i * cam->w + j, 1, out, nullptr);
  return out.n;
}

void orc_gather_brute(int64_t n_hp, const double* pos3, const double* n3, const double* wgt3, const double* r2,
                      int64_t m, const double* x3, const double* dn3, const double* flux3, double search_r2,
                      int32_t* newN, double* tao3, uint64_t* cand_pairs, uint64_t* acc_pairs) {
  HitPoints hp{n_hp, pos3, n3, wgt3, r2};
  uint64_t cand = 0, acc = 0;
  for (int64_t d = 0; d < m; d++)
    gatherOne(hp, nullptr, V3(x3 + 3 * d), V3(dn3 + 3 * d), V3(flux3 + 3 * d), search_r2, newN, tao3, &cand, &acc);
  if (cand_pairs) *cand_pairs = cand;
  if (acc_pairs) *acc_pairs = acc;
}
void* orc_kdtree_build(int64_t n_hp, const double* pos3) {
  KdTree* t = new KdTree();
  t->pts = pos3;
  t->idx.resize(n_hp);
  for (int64_t i = 0; i < n_hp; i++) t->idx[i] = (int)i;
  t->nodes.reserve(n_hp / 4 + 16);
  t->build(0, (int)n_hp);
  return t;
}
void orc_kdtree_free(void* tree) { delete (KdTree*)tree; }
void orc_gather_kdtree(void* tree, int64_t n_hp, const double* pos3, const double* n3, const double* wgt3,
                       const double* r2, int64_t m, const double* x3, const double* dn3, const double* flux3,
                       double search_r2, int32_t* newN, double* tao3) {
  HitPoints hp{n_hp, pos3, n3, wgt3, r2};
  for (int64_t d = 0; d < m; d++)
    gatherOne(hp, (KdTree*)tree, V3(x3 + 3 * d), V3(dn3 + 3 * d), V3(flux3 + 3 * d), search_r2, newN, tao3, nullptr,
              nullptr);
}

int64_t orc_photon_pass(const SppmSceneDesc* scene, const SppmParams* prm, uint64_t begin, uint64_t count,
                        uint64_t pass_id, void* tree, int64_t n_hp, const double* pos3, const double* n3,
                        const double* wgt3, const double* r2, int32_t* newN, double* tao3, int64_t cap_dep,
                        double* dep_x3, double* dep_n3, double* dep_flux3, uint64_t* dep_photon, int32_t* dep_bounce,
                        uint64_t* counters) {
  SceneC sc(scene);
  HitPoints hp{n_hp, pos3, n3, wgt3, r2};
  DepositSink sink;
  sink.cap = cap_dep;
  sink.x = dep_x3;
  sink.nn = dep_n3;
  sink.flux = dep_flux3;
  sink.photon = dep_photon;
  sink.bounce = dep_bounce;
  for (size_t pl = 0; pl < sc.lights.size(); pl++)  // Raytracer.h:226-233
    for (uint64_t j = begin; j < begin + count; j++) {
      uint64_t id = ((uint64_t)pl << 40) | j;
      // Light::emit (Light.h:8-13)
      V3 flux = V3(sc.lights[pl].color) * PI_ * 4;
      V3 dir = GetUnitRandVec(prm->seed, pass_id, id, P_EMIT);
      V3 org(sc.lights[pl].pos);
      if (counters) counters[3]++;
      PhotonTrace(sc, *prm, org, dir, flux, pass_id, id, hp, (KdTree*)tree, newN, tao3, dep_x3 ? &sink : nullptr,
                  counters);
    }
  return sink.n;
}

// Timing leg: `threads` OpenMP threads, each an independent photon tracer with private
// accumulators over the shared read-only hit-point map + kd-tree (the reference's replica
// structure, Raytracer.h:442-447, restricted to the photon pass).  Thread t traces photons
// [t*count, (t+1)*count) of every light.  Returns wall seconds of the parallel region.
double orc_photon_bench(const SppmSceneDesc* scene, const SppmParams* prm, uint64_t count, uint64_t pass_id,
                        int32_t threads, void* tree, int64_t n_hp, const double* pos3, const double* n3,
                        const double* wgt3, const double* r2, uint64_t* counters) {
  SceneC sc(scene);
  HitPoints hp{n_hp, pos3, n3, wgt3, r2};
  std::vector<std::vector<uint64_t>> cnts(threads, std::vector<uint64_t>(4, 0));
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  std::vector<std::vector<int32_t>> newN(threads);
  std::vector<std::vector<double>> tao(threads);
  for (int t = 0; t < threads; t++) {
    newN[t].assign(n_hp, 0);
    tao[t].assign(3 * n_hp, 0.0);
  }
  const double t0 = now();
#pragma omp parallel for num_threads(threads) schedule(static, 1)
  for (int t = 0; t < threads; t++) {
    uint64_t local[4] = {0, 0, 0, 0};  // on the thread's stack: the per-thread vectors share cache lines
    for (size_t pl = 0; pl < sc.lights.size(); pl++)
      for (uint64_t j = (uint64_t)t * count; j < (uint64_t)(t + 1) * count; j++) {
        uint64_t id = ((uint64_t)pl << 40) | j;
        V3 flux = V3(sc.lights[pl].color) * PI_ * 4;
        V3 dir = GetUnitRandVec(prm->seed, pass_id, id, P_EMIT);
        local[3]++;
        PhotonTrace(sc, *prm, V3(sc.lights[pl].pos), dir, flux, pass_id, id, hp, (KdTree*)tree, newN[t].data(),
                    tao[t].data(), nullptr, local);
      }
    for (int k = 0; k < 4; k++) cnts[t][k] = local[k];
  }
  const double t1 = now();
  if (counters)
    for (int k = 0; k < 4; k++) {
      counters[k] = 0;
      for (int t = 0; t < threads; t++) counters[k] += cnts[t][k];
    }
  return t1 - t0;
}

// Timing leg for M2: `threads` threads solve n (ray, patch) pairs with Bezier3::newton.
double orc_patch_bench(const SppmSceneDesc* scene, const SppmParams* prm, const double* rays6,
                       const int32_t* patch_idx, int64_t n, int32_t threads, double* tuv_out) {
  SceneC sc(scene);
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 8)
  for (int64_t i = 0; i < n; i++) {
    int gp = patch_idx[i];
    V3 r = newton(&sc.P[(size_t)gp * 16], V3(rays6 + 6 * i), V3(rays6 + 6 * i + 3), prm->newton_restarts,
                  prm->newton_iters, [&](int o, double uv[2]) {
                    draw2(prm->seed, 0, (uint64_t)i, 0, SPPM_DOMAIN_TEST, (uint32_t)o, (uint32_t)gp, uv);
                  });
    for (int k = 0; k < 3; k++) tuv_out[3 * i + k] = r[k];
  }
  return now() - t0;
}

void orc_update(int64_t n_hp, int32_t mode, double alpha, int32_t* newN, int32_t* N, double* r2, double* tao3) {
  updateAll(n_hp, mode, alpha, newN, N, r2, tao3);
}
void orc_resolve(int64_t n_hp, const int32_t* pix, const double* r2, const double* tao3, int32_t cnt, int32_t onetime,
                 int32_t w, int32_t h, double* film) {
  resolveFilm(n_hp, pix, r2, tao3, cnt, onetime, w, h, film);
}
void orc_tonemap(const double* film, int32_t w, int32_t h, double div, uint8_t* bgr) {
  for (int i1 = 0; i1 < h; i1++)  // Raytracer.h:460-470
    for (int j = 0; j < w; j++)
      for (int k = 0; k < 3; k++)
        bgr[((size_t)(h - 1 - i1) * w + j) * 3 + k] = (uint8_t)toInt(film[((size_t)i1 * w + j) * 3 + k] / div);
}

void orc_log_begin(void) {
  static std::vector<double> vals;
  static std::vector<signed char> kinds;
  vals.clear();
  kinds.clear();
  g_draw_log = &vals;
  g_draw_kind = &kinds;
}
int64_t orc_log_end(double* vals, signed char* kinds, int64_t cap) {
  int64_t n = g_draw_log ? (int64_t)g_draw_log->size() : 0;
  for (int64_t i = 0; i < n && i < cap; i++) {
    vals[i] = (*g_draw_log)[i];
    kinds[i] = (*g_draw_kind)[i];
  }
  g_draw_log = nullptr;
  g_draw_kind = nullptr;
  return n;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_render_replicas(const SppmSceneDesc* scene, const SppmParams* prm, const double cam_pos[3],
                         const double cam_look[3], int32_t w, int32_t h, int32_t cnt, int32_t onetime, int32_t threads,
                         double jitter, double* film, double* times, uint64_t* counters) {
  SceneC sc(scene);
  std::vector<std::vector<double>> films(threads);
  std::vector<double> t_eye(threads, 0), t_build(threads, 0), t_photon(threads, 0);
  std::vector<std::vector<uint64_t>> cnts(threads, std::vector<uint64_t>(4, 0));
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t0 = now();
#pragma omp parallel for num_threads(threads) schedule(static, 1)
  for (int j = 0; j < threads; j++) {  // Raytracer.h:442-447: one RayTracer per thread
    // Raytracer.h:430-441: c[i].MoveTo(camera.pos + 0.00015*GetUnitRandVec()); lookAt(look, 1.)
    V3 pos(cam_pos);
    if (jitter != 0) pos = pos + jitter * GetUnitRandVec(prm->seed, 0, (uint64_t)j, P_JITTER);
    SppmCameraDesc cam;
    cameraLookAt(pos, V3(cam_look), w, h, &cam);
    double a = now();
    // RayTracer::render (Raytracer.h:366-387)
    int64_t cap = (int64_t)w * h * 4 + 1024;
    std::vector<double> pos3(3 * cap), n3(3 * cap), wgt3(3 * cap), r2(cap);
    std::vector<int32_t> pix(2 * cap);
    HpSink out{cap, 0, pos3.data(), n3.data(), wgt3.data(), r2.data(), pix.data(), nullptr, prm->init_r2};
    for (int i = 0; i < h; i++)
      for (int jj = 0; jj < w; jj++) {
        V3 o, d;
        cameraEmit(&cam, i, jj, o, d);
        ViewTrace(sc, *prm, o, d, 0, V3(1, 1, 1), i, jj, (uint64_t)j, (uint64_t)i * w + jj, 1, out, nullptr);
      }
    int64_t n_hp = std::min(out.n, cap);
    double b = now();
    KdTree tree;
    tree.pts = pos3.data();
    tree.idx.resize(n_hp);
    for (int64_t i = 0; i < n_hp; i++) tree.idx[i] = (int)i;
    tree.build(0, (int)n_hp);
    double c = now();
    std::vector<int32_t> newN(n_hp, 0), N(n_hp, 0);
    std::vector<double> tao(3 * n_hp, 0.0);
    HitPoints hp{n_hp, pos3.data(), n3.data(), wgt3.data(), r2.data()};
    for (int i = 0; i < cnt; i++) {  // PhotonMap (Raytracer.h:220-246)
      for (size_t pl = 0; pl < sc.lights.size(); pl++)
        for (int k = 0; k < onetime; k++) {
          // draw addressing of the host mirrors: photon k of round i = photon i*onetime + k of pass j
          uint64_t id = ((uint64_t)pl << 40) | ((uint64_t)i * (uint64_t)onetime + (uint64_t)k);
          V3 flux = V3(sc.lights[pl].color) * PI_ * 4;
          V3 dir = GetUnitRandVec(prm->seed, (uint64_t)j, id, P_EMIT);
          cnts[j][3]++;
          PhotonTrace(sc, *prm, V3(sc.lights[pl].pos), dir, flux, (uint64_t)j, id, hp, &tree, newN.data(),
                      tao.data(), nullptr, cnts[j].data());
        }
      updateAll(n_hp, prm->update_mode, prm->alpha, newN.data(), N.data(), r2.data(), tao.data());
    }
    films[j].assign(3 * (size_t)w * h, 0.0);
    resolveFilm(n_hp, pix.data(), r2.data(), tao.data(), cnt, onetime, w, h, films[j].data());
    double d = now();
    t_eye[j] = b - a;
    t_build[j] = c - b;
    t_photon[j] = d - c;
  }
  double t1 = now();
  if (film) {
    std::memset(film, 0, sizeof(double) * 3 * (size_t)w * h);
    for (int j = 0; j < threads; j++)  // Raytracer.h:449-458
      for (size_t i = 0; i < 3 * (size_t)w * h; i++) film[i] += films[j][i];
  }
  if (times) {
    times[0] = *std::max_element(t_eye.begin(), t_eye.end());
    times[1] = *std::max_element(t_build.begin(), t_build.end());
    times[2] = *std::max_element(t_photon.begin(), t_photon.end());
    times[3] = t1 - t0;
  }
  if (counters)
    for (int k = 0; k < 4; k++) {
      counters[k] = 0;
      for (int j = 0; j < threads; j++) counters[k] += cnts[j][k];
    }
}

}  // extern "C"
