// device_math.cuh — FP64 vector helpers, constants and the counter RNG for the sm_100a
// kernels.  The whole library is compiled with -fmad=false, so every expression below is
// evaluated with the reference's association and without FMA contraction (the CPU
// reference is MSVC /fp:precise); FMA is used only where fma() is written out (FAST
// Newton).  Citations: /root/reference/raytracer/<file>:<line>.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace sppm {

// Vec3.h:6-14
#define SPPM_M_EPS 1e-4
#define SPPM_PI 3.1415926
#define SPPM_PI_INV 0.3183098862
#define SPPM_MAX_NUM 1e20
#define SPPM_MIN_NUM -1e20
#define SPPM_NORMAL_AGREE 1e-3  // Raytracer.h:154

constexpr unsigned FULL = 0xffffffffu;

struct d3 {
  double x, y, z;
};
__host__ __device__ inline d3 mk(double x, double y, double z) { return d3{x, y, z}; }
__host__ __device__ inline d3 ld3(const double* p) { return d3{p[0], p[1], p[2]}; }
__host__ __device__ inline d3 operator+(d3 a, d3 b) { return d3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ inline d3 operator-(d3 a, d3 b) { return d3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ inline d3 operator*(d3 a, double s) { return d3{a.x * s, a.y * s, a.z * s}; }
__host__ __device__ inline d3 operator-(d3 a) { return d3{-a.x, -a.y, -a.z}; }
__host__ __device__ inline d3 mul3(d3 a, d3 b) { return d3{a.x * b.x, a.y * b.y, a.z * b.z}; }
// cv::Matx::dot accumulates from 0 in order: ((0 + a0b0) + a1b1) + a2b2; 0 + p == p for
// every p except -0, whose sign cannot survive the following additions differently.
__host__ __device__ inline double dot3(d3 a, d3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__host__ __device__ inline d3 cross3(d3 a, d3 b) {
  return d3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__host__ __device__ inline bool is_zero(double x) { return fabs(x) < SPPM_M_EPS; }  // Vec3.h:23
__host__ __device__ inline bool is_zero3(d3 a) { return is_zero(a.x) || is_zero(a.y) || is_zero(a.z); }  // Vec3.h:72-79
__host__ __device__ inline double module2(d3 a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }  // Vec3.h:44
__host__ __device__ inline double module(d3 a) { return sqrt(module2(a)); }
__host__ __device__ inline d3 normalize3(d3 a) {  // Vec3.h:48-55
  double mo = module(a);
  if (!is_zero(mo)) {
    double mo_ = 1 / mo;
    return d3{a.x * mo_, a.y * mo_, a.z * mo_};
  }
  return a;
}
__host__ __device__ inline d3 get_unit(d3 a) {  // Vec3.h:56
  double mo = module(a);
  return is_zero(mo) ? a : a * (1.0 / mo);
}
__host__ __device__ inline d3 get_refl(d3 I, d3 n) { return I - n * (2 * dot3(I, n)); }  // Vec3.h:80-83
__host__ __device__ inline d3 get_anormal(d3 a) {                                         // Vec3.h:85-89
  return (a.x == 0 && a.y == 0) ? d3{1, 0, 0} : get_unit(d3{a.y, -a.x, 0});
}
// Rodrigues rotation of s about the unit axis ax, given sin and cos of the angle (Vec3.h:102-114)
__device__ inline d3 rotated_sc(d3 s, d3 ax, double sint, double cost) {
  d3 r;
  r.x = ((0.0 + s.x * (ax.x * ax.x + (1 - ax.x * ax.x) * cost)) + s.y * (ax.x * ax.y * (1 - cost) - ax.z * sint)) +
        s.z * (ax.x * ax.z * (1 - cost) + ax.y * sint);
  r.y = ((0.0 + s.x * (ax.y * ax.x * (1 - cost) + ax.z * sint)) + s.y * (ax.y * ax.y + (1 - ax.y * ax.y) * cost)) +
        s.z * (ax.y * ax.z * (1 - cost) - ax.x * sint);
  r.z = ((0.0 + s.x * (ax.z * ax.x * (1 - cost) - ax.y * sint)) + s.y * (ax.z * ax.y * (1 - cost) + ax.x * sint)) +
        s.z * (ax.z * ax.z + (1 - ax.z * ax.z) * cost);
  return r;
}
__device__ inline d3 rotated(d3 s, d3 ax, double angle) {  // Vec3.h:99-115
  if (fabs(angle) < SPPM_M_EPS) return s;
  double sint, cost;
  sincos(angle, &sint, &cost);  // one range reduction for both (same values as sin() and cos())
  return rotated_sc(s, ax, sint, cost);
}
// Cosine-weighted direction about norm (Vec3.h:90-98).  fast = SPPM_MATH_FAST: theta =
// acos(sqrt(U1)) is only ever used through its sine and cosine, which are sqrt(1-U1) and
// sqrt(U1) exactly; taking them directly saves acos() and one sincos(), and the two rotations
// are written in closed form; the direction differs from the reference's chain by a few ulps.  The reference's "no rotation below 1e-4 rad"
// rule (Vec3.h:100) is kept: theta < 1e-4  <=>  cos(theta) > cos(1e-4).
__device__ inline d3 unit_rand_refl(d3 norm, double U1, double U2, bool fast = false) {
  const double phi = 2 * SPPM_PI * U2;
  if (fast) {
    // With A = GetAnormal(N) (unit, A.N = 0) the two Rodrigues rotations collapse to
    //   dir = N cos(theta) + sin(theta) ((A x N) cos(phi) + A sin(phi)),
    // 21 flops instead of two 3x3 rotations and a normalisation; the reference's "no rotation
    // below 1e-4 rad" rule (Vec3.h:100) is kept for both angles.
    const double cost = sqrt(U1), sint = sqrt(1 - U1);
    const bool rot1 = !(cost > 0.999999995), rot2 = !(fabs(phi) < SPPM_M_EPS);
    const d3 A = get_anormal(norm);
    const d3 B = cross3(A, norm);
    double sp = 0, cp = 1;
    if (rot2) sincos(phi, &sp, &cp);
    if (!rot1) return norm;  // theta ~ 0: the second rotation is about the vector itself
    const d3 t = B * cp + A * sp;
    return norm * cost + t * sint;
  }
  const d3 ret = rotated(norm, get_anormal(norm), acos(sqrt(U1)));
  return get_unit(rotated(ret, norm, phi));
}
__device__ inline d3 refracted(d3 I, d3 N, double n) {  // Vec3.h:124-134
  double cosI = -dot3(N, I);
  double cosR2 = 1 - (1 - cosI * cosI) * (n * n);
  return (cosR2 > SPPM_M_EPS) ? I * n + N * (n * cosI - sqrt(cosR2)) : get_refl(I, N);
}

// ---- Philox4x32-10 counter RNG (replaces the global cv::RNG, Vec3.h:5) ---------------
// Draw address: (seed, pass, id, node, purpose, sub, patch) -> two 53-bit uniforms.
enum { P_EMIT = 0, P_BOUNCE = 1, P_JITTER = 4 };
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__host__ __device__ inline double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
__host__ __device__ inline void draw2(uint64_t seed, uint64_t pass, uint64_t id, uint32_t node, uint32_t purpose,
                                      uint32_t sub, uint32_t patch, double& a, double& b) {
  uint32_t o[4];
  philox4x32_10((uint32_t)id, (uint32_t)((id >> 32) & 0xffffu) | (node << 16),
                (purpose & 0xfu) | ((sub & 0xfffu) << 4) | (patch << 16), (uint32_t)pass, (uint32_t)seed,
                (uint32_t)(seed >> 32), o);
  a = u53(o[0], o[1]);
  b = u53(o[2], o[3]);
}
// cv::RNG::uniform(double a, double b) = u*(b-a)+a (Vec3.h:15-18)
__host__ __device__ inline double randu(double u, double r1, double r2) { return u * (r2 - r1) + r1; }

// Vec3::GetUnitRandVec (Vec3.h:57-65) with addressed draws.
__host__ __device__ inline d3 unit_rand_vec(uint64_t seed, uint64_t pass, uint64_t id, uint32_t purpose) {
  double x, y, z, r2;
  uint32_t a = 0;
  do {
    double p, q, r, s;
    draw2(seed, pass, id, 0, purpose, 2 * a, 0, p, q);
    draw2(seed, pass, id, 0, purpose, 2 * a + 1, 0, r, s);
    x = randu(p, -1., 1.);
    y = randu(q, -1., 1.);
    z = randu(r, -1., 1.);
    r2 = (x * x + y * y) + z * z;
    a++;
  } while ((r2 > 1 || r2 < SPPM_M_EPS) && a < 2000);
  return get_unit(d3{x, y, z});
}

// order-preserving map double -> uint64 (for atomic min/max on coordinates)
__device__ inline unsigned long long dbl_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline double key_dbl(unsigned long long k) {
  unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}

}  // namespace sppm
