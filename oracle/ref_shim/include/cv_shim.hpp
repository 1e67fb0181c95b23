// cv_shim.hpp — the slice of OpenCV the reference's headers use, restated from OpenCV's
// documented Matx/Vec semantics (core/matx.hpp, core/operations.hpp), so that the
// reference's OWN sources (/root/reference/raytracer/*.h) compile and run here without
// OpenCV.  TEST INFRASTRUCTURE (oracle/_ref build only).  What is restated:
//   cv::Vec<T,3>: element-wise + - * (scale), unary minus = scale by -1, dot = sequential
//     accumulation from 0, cross, mul; cv::Matx33d: zero default ctor, eye, (i,j), product
//     and Matx*Vec with sequential accumulation from 0, inv() = Matx_FastInvOp<_Tp,3,3>
//     (adjugate / determinant, zeros when det == 0) — pinned against cv2.invert in
//     tests/test_oracle_cpu.py;
//   cv::RNG::uniform(a,b) = u*(b-a)+a with u taken from an INJECTED sequence
//     (ref_shim::push_uniforms), because the real generator's stream is not a parity target
//     (timing runs switch to a free-running multiply-with-carry generator, ref_shim::free_state);
//   cv::imread: returns images registered by the harness (no codec); Mat/imwrite/imshow: stubs.
#pragma once
#include <math.h>
#include <stdlib.h>

#include <cstdio>
#include <deque>
#include <iostream>
#include <map>
#include <string>
#include <vector>

typedef unsigned char uchar;
#define CV_8UC3 16

namespace ref_shim {
inline std::deque<double>& uniforms() {
  static thread_local std::deque<double> q;
  return q;
}
inline long& underflows() {
  static thread_local long n = 0;
  return n;
}
// Timing runs (ref_time_photon_map) let the generator run freely instead: state != 0 selects
// cv::RNG's own multiply-with-carry recurrence (core.hpp: CV_RNG_COEFF 4164903690, a double is
// two 32-bit draws scaled by 2^-64), one state per thread.
inline unsigned long long& free_state() {
  static thread_local unsigned long long s = 0;
  return s;
}
struct Image {
  int rows = 0, cols = 0;
  const uchar* bgr = nullptr;
};
inline std::map<std::string, Image>& images() {
  static std::map<std::string, Image> m;
  return m;
}
}  // namespace ref_shim

namespace cv {

template <typename T, int cn>
class Vec {
 public:
  T val[cn];
  Vec() { for (int i = 0; i < cn; i++) val[i] = T(0); }
  Vec(T a, T b, T c) { val[0] = a; val[1] = b; val[2] = c; }
  T& operator()(int i) { return val[i]; }
  const T& operator()(int i) const { return val[i]; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
  T dot(const Vec& M) const {  // Matx::dot
    T s = 0;
    for (int i = 0; i < cn; i++) s += val[i] * M.val[i];
    return s;
  }
  Vec cross(const Vec& v) const {  // Vec<double,3>::cross
    return Vec(val[1] * v.val[2] - val[2] * v.val[1], val[2] * v.val[0] - val[0] * v.val[2],
               val[0] * v.val[1] - val[1] * v.val[0]);
  }
  Vec mul(const Vec& a) const {  // Matx_MulOp
    Vec r;
    for (int i = 0; i < cn; i++) r.val[i] = val[i] * a.val[i];
    return r;
  }
};
typedef Vec<double, 3> Vec3d;
typedef Vec<uchar, 3> Vec3b;

template <typename T, int cn> Vec<T, cn> operator+(const Vec<T, cn>& a, const Vec<T, cn>& b) { Vec<T, cn> r; for (int i = 0; i < cn; i++) r.val[i] = a.val[i] + b.val[i]; return r; }
template <typename T, int cn> Vec<T, cn> operator-(const Vec<T, cn>& a, const Vec<T, cn>& b) { Vec<T, cn> r; for (int i = 0; i < cn; i++) r.val[i] = a.val[i] - b.val[i]; return r; }
template <typename T, int cn> Vec<T, cn> operator*(const Vec<T, cn>& a, double alpha) { Vec<T, cn> r; for (int i = 0; i < cn; i++) r.val[i] = (T)(a.val[i] * alpha); return r; }
template <typename T, int cn> Vec<T, cn> operator*(double alpha, const Vec<T, cn>& a) { Vec<T, cn> r; for (int i = 0; i < cn; i++) r.val[i] = (T)(a.val[i] * alpha); return r; }
template <typename T, int cn> Vec<T, cn> operator*(const Vec<T, cn>& a, int alpha) { Vec<T, cn> r; for (int i = 0; i < cn; i++) r.val[i] = (T)(a.val[i] * alpha); return r; }
template <typename T, int cn> Vec<T, cn> operator*(int alpha, const Vec<T, cn>& a) { Vec<T, cn> r; for (int i = 0; i < cn; i++) r.val[i] = (T)(a.val[i] * alpha); return r; }
template <typename T, int cn> Vec<T, cn> operator-(const Vec<T, cn>& a) { Vec<T, cn> r; for (int i = 0; i < cn; i++) r.val[i] = (T)(a.val[i] * -1); return r; }  // Matx_ScaleOp(-1)
template <typename T, int cn> Vec<T, cn>& operator+=(Vec<T, cn>& a, const Vec<T, cn>& b) { for (int i = 0; i < cn; i++) a.val[i] = a.val[i] + b.val[i]; return a; }
template <typename T, int cn> Vec<T, cn>& operator-=(Vec<T, cn>& a, const Vec<T, cn>& b) { for (int i = 0; i < cn; i++) a.val[i] = a.val[i] - b.val[i]; return a; }
template <typename T, int cn> Vec<T, cn>& operator*=(Vec<T, cn>& a, double alpha) { for (int i = 0; i < cn; i++) a.val[i] = (T)(a.val[i] * alpha); return a; }
template <typename T, int cn> std::ostream& operator<<(std::ostream& o, const Vec<T, cn>& v) { o << "["; for (int i = 0; i < cn; i++) o << (i ? ", " : "") << v.val[i]; return o << "]"; }

class Matx33d {
 public:
  double val[9];
  Matx33d() { for (int i = 0; i < 9; i++) val[i] = 0; }
  static Matx33d eye() { Matx33d m; m.val[0] = m.val[4] = m.val[8] = 1; return m; }
  double& operator()(int i, int j) { return val[i * 3 + j]; }
  const double& operator()(int i, int j) const { return val[i * 3 + j]; }
  Matx33d inv() const {  // Matx_FastInvOp<_Tp,3,3> with Matx_DetOp<_Tp,3>; zeros if det == 0
    const Matx33d& a = *this;
    Matx33d b;
    double d = a(0, 0) * (a(1, 1) * a(2, 2) - a(2, 1) * a(1, 2)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(2, 0) * a(1, 2)) +
               a(0, 2) * (a(1, 0) * a(2, 1) - a(2, 0) * a(1, 1));
    if (d == 0) return Matx33d();
    d = 1 / d;
    b(0, 0) = (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) * d;
    b(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * d;
    b(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * d;
    b(1, 0) = (a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2)) * d;
    b(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * d;
    b(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * d;
    b(2, 0) = (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0)) * d;
    b(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * d;
    b(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * d;
    return b;
  }
};
inline Matx33d operator*(const Matx33d& a, const Matx33d& b) {  // Matx_MatMulOp
  Matx33d c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += a(i, k) * b(k, j);
      c(i, j) = s;
    }
  return c;
}
inline Vec3d operator*(const Matx33d& a, const Vec3d& b) {
  Vec3d c;
  for (int i = 0; i < 3; i++) {
    double s = 0;
    for (int k = 0; k < 3; k++) s += a(i, k) * b.val[k];
    c.val[i] = s;
  }
  return c;
}

class RNG {  // uniform(a,b) = ((double)*this)*(b-a)+a, the uniform coming from the injected queue
 public:
  double next01() {
    unsigned long long& st = ref_shim::free_state();
    if (st) {
      st = (unsigned long long)(unsigned)st * 4164903690U + (unsigned)(st >> 32);
      unsigned hi = (unsigned)st;
      st = (unsigned long long)(unsigned)st * 4164903690U + (unsigned)(st >> 32);
      return (((unsigned long long)hi << 32) | (unsigned)st) * 5.4210108624275221700372640043497e-20;
    }
    auto& q = ref_shim::uniforms();
    if (q.empty()) {
      ref_shim::underflows()++;
      return 0.5;
    }
    double u = q.front();
    q.pop_front();
    return u;
  }
  double uniform(double a, double b) { return next01() * (b - a) + a; }
  int uniform(int a, int b) { return a + (int)(next01() * (b - a)); }
};

class Mat {
 public:
  int rows = 0, cols = 0;
  std::vector<uchar> data;
  Mat() {}
  Mat(int r, int c, int) : rows(r), cols(c), data((size_t)r * c * 3, 0) {}
  template <class S> Mat(int r, int c, int, S) : rows(r), cols(c), data((size_t)r * c * 3, 0) {}
  template <class T> T* ptr(int y) { return (T*)(data.data() + (size_t)y * cols * 3); }
};
template <class T>
class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(const Mat& m) : Mat(m) {}
  T operator()(int i, int j) const {
    const uchar* p = data.data() + ((size_t)i * cols + j) * 3;
    return T(p[0], p[1], p[2]);
  }
};
inline Mat imread(const std::string& name) {
  Mat m;
  auto it = ref_shim::images().find(name);
  if (it != ref_shim::images().end()) {
    m.rows = it->second.rows;
    m.cols = it->second.cols;
    m.data.assign(it->second.bgr, it->second.bgr + (size_t)m.rows * m.cols * 3);
  }
  return m;
}
inline bool imwrite(const std::string&, const Mat&) { return true; }
inline void namedWindow(const char*) {}
inline void imshow(const char*, const Mat&) {}
}  // namespace cv
struct CvScalarShim {};
inline CvScalarShim cvScalar(double, double, double) { return CvScalarShim(); }
inline int cvWaitKey(int) { return 0; }
