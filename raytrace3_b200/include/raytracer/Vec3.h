// Vec3.h — drop-in for raytracer/Vec3.h (wondergo2017/raytrace3) without OpenCV.
// Same names and arithmetic (cv::Vec3d semantics: element-wise ops, sequential dot); the
// global cv::RNG (Vec3.h:5) is replaced by the library's counter RNG on the device, so the
// host only keeps a small deterministic generator for API compatibility (GetRandu).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <iostream>

#define M_EPS 1e-4
#define RAND_R 10000000
#define PI 3.1415926
#define PI_INV 0.3183098862
#define Color Vec3
#define MAX_NUM 1e20
#define MIN_NUM -1e20
#define MAX_VEC3 (Vec3(MAX_NUM, MAX_NUM, MAX_NUM))
#define MIN_VEC3 (Vec3(MIN_NUM, MIN_NUM, MIN_NUM))

namespace sppm_host {
// splitmix64: host-side stand-in for cv::RNG (only used by GetRandu / GetUnitRandVec callers
// such as the camera jitter of SPPMRayTracer::render, Raytracer.h:434).
inline uint64_t& rng_state() {
  static uint64_t s = 0x5EED5EEDull;
  return s;
}
inline double uniform01() {
  uint64_t z = (rng_state() += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
}  // namespace sppm_host

inline double GetRandu(double r1, double r2) { return sppm_host::uniform01() * (r2 - r1) + r1; }  // Vec3.h:15-18
inline double GetRandu() { return GetRandu(-RAND_R, RAND_R); }
inline bool IsZero(double i) { return std::fabs(i) < M_EPS; }  // Vec3.h:23
inline int GetRandu(int i, int j) { return i + (int)(sppm_host::uniform01() * (j - i)); }

class Vec3 {
 public:
  double val[3];
  Vec3() : val{0, 0, 0} {}
  Vec3(double a, double b, double c) : val{a, b, c} {}
  double& operator()(int i) { return val[i]; }
  double operator()(int i) const { return val[i]; }
  double& operator[](int i) { return val[i]; }
  double operator[](int i) const { return val[i]; }
  Vec3& operator+=(double num) { for (int i = 0; i < 3; i++) val[i] += num; return *this; }
  Vec3& operator-=(double num) { for (int i = 0; i < 3; i++) val[i] -= num; return *this; }
  Vec3& operator+=(const Vec3& v) { for (int i = 0; i < 3; i++) val[i] += v.val[i]; return *this; }
  Vec3& operator-=(const Vec3& v) { for (int i = 0; i < 3; i++) val[i] -= v.val[i]; return *this; }
  Vec3& operator*=(double s) { for (int i = 0; i < 3; i++) val[i] *= s; return *this; }
  Vec3 operator+(double num) const { return Vec3(val[0] + num, val[1] + num, val[2] + num); }
  Vec3 operator-(double num) const { return Vec3(val[0] - num, val[1] - num, val[2] - num); }
  Vec3 operator+(const Vec3& v) const { return Vec3(val[0] + v.val[0], val[1] + v.val[1], val[2] + v.val[2]); }
  Vec3 operator-(const Vec3& v) const { return Vec3(val[0] - v.val[0], val[1] - v.val[1], val[2] - v.val[2]); }
  Vec3 operator*(double s) const { return Vec3(val[0] * s, val[1] * s, val[2] * s); }
  Vec3 operator-() const { return Vec3(-val[0], -val[1], -val[2]); }
  bool operator<(const Vec3& v) const { return val[0] < v.val[0] && val[1] < v.val[1] && val[2] < v.val[2]; }
  bool operator<(const double& k) const { return val[0] < k && val[1] < k && val[2] < k; }
  bool operator>(const Vec3& v) const { return val[0] > v.val[0] && val[1] > v.val[1] && val[2] > v.val[2]; }
  double dot(const Vec3& b) const { double s = 0; for (int i = 0; i < 3; i++) s += val[i] * b.val[i]; return s; }
  Vec3 cross(const Vec3& b) const {
    return Vec3(val[1] * b.val[2] - val[2] * b.val[1], val[2] * b.val[0] - val[0] * b.val[2], val[0] * b.val[1] - val[1] * b.val[0]);
  }
  Vec3 mul(const Vec3& b) const { return Vec3(val[0] * b.val[0], val[1] * b.val[1], val[2] * b.val[2]); }
  double Module2() const { double res = 0; for (int i = 0; i < 3; i++) res += val[i] * val[i]; return res; }
  double Module() const { return std::sqrt(Module2()); }
  friend double dist2(const Vec3& v1, const Vec3& v2) { return (v1 - v2).Module2(); }
  friend double dist(const Vec3& v1, const Vec3& v2) { return std::sqrt(dist2(v1, v2)); }
  Vec3& Normalize() {  // Vec3.h:48-55
    double mo = Module();
    if (!IsZero(mo)) {
      double mo_ = 1 / mo;
      for (int i = 0; i < 3; i++) val[i] *= mo_;
    }
    return *this;
  }
  Vec3 GetUnit() const { double mo = Module(); return IsZero(mo) ? *this : *this * (1.0 / mo); }  // Vec3.h:56
  static Vec3 GetUnitRandVec() {  // Vec3.h:57-65
    double x, y, z;
    do {
      x = GetRandu(-1., 1.);
      y = GetRandu(-1., 1.);
      z = GetRandu(-1., 1.);
    } while (x * x + y * y + z * z > 1 || x * x + y * y + z * z < M_EPS);
    return Vec3(x, y, z).GetUnit();
  }
  friend void Confine(Vec3& v1) { for (int i = 0; i < 3; i++) if (v1.val[i] > 1) v1.val[i] = 1; }
  friend bool IsZero(const Vec3& v1) {  // Vec3.h:72-79: true if ANY channel is ~0
    for (int i = 0; i < 3; i++) if (IsZero(v1.val[i])) return true;
    return false;
  }
  Vec3 GetRefl(const Vec3& norm) const { return *this - norm * (2 * this->dot(norm)); }
  Vec3 GetUnitRefl(const Vec3& norm) const { return GetRefl(norm).GetUnit(); }
  Vec3 GetAnormal() const { return (val[0] == 0 && val[1] == 0) ? Vec3(1, 0, 0) : Vec3(val[1], -val[0], 0).GetUnit(); }
  double GetPower() const { return (val[0] + val[1] + val[2]) / 3; }
  bool isnan() const { for (int i = 0; i < 3; i++) if (!(val[i] < 2) && !(val[i] > 1)) return true; return false; }
};
inline Vec3 operator*(double s, const Vec3& v) { return v * s; }
inline std::ostream& operator<<(std::ostream& o, const Vec3& v) { return o << "[" << v.val[0] << ", " << v.val[1] << ", " << v.val[2] << "]"; }

// The one OpenCV type the scene code needs (Scene.h:142-153, Bezier.h:199): a 3x3 double matrix.
namespace cv {
struct Matx33d {
  double m[3][3];
  Matx33d() : m{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}} {}
  static Matx33d eye() { Matx33d r; r.m[0][0] = r.m[1][1] = r.m[2][2] = 1; return r; }
  double& operator()(int i, int j) { return m[i][j]; }
  double operator()(int i, int j) const { return m[i][j]; }
  Matx33d operator*(const Matx33d& b) const {
    Matx33d r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += m[i][k] * b.m[k][j]; r.m[i][j] = s; }
    return r;
  }
  Vec3 operator*(const Vec3& v) const {
    Vec3 r;
    for (int i = 0; i < 3; i++) { double s = 0; for (int k = 0; k < 3; k++) s += m[i][k] * v.val[k]; r.val[i] = s; }
    return r;
  }
};
}  // namespace cv
