// flann_shim.hpp — flann::Index<L2<double>> as the reference uses it (Raytracer.h:92-96, 146,
// 378-380): exact radius search on squared L2 distance, strict '<' (pinned against
// cv2.flann_Index in tests/test_oracle_cpu.py), results sorted by distance (SearchParams
// default sorted = true).  Exhaustive scan: this is a checker, not a fast index.
#pragma once
#include <algorithm>
#include <vector>
namespace flann {
template <class T> struct L2 { typedef T ElementType; };
template <class T>
struct Matrix {
  T* data = nullptr;
  size_t rows = 0, cols = 0;
  Matrix() {}
  Matrix(T* d, size_t r, size_t c) : data(d), rows(r), cols(c) {}
  T* operator[](size_t i) const { return data + i * cols; }
};
struct KDTreeSingleIndexParams {};
struct SearchParams {};
template <class Dist>
class Index {
  typedef typename Dist::ElementType T;
  Matrix<T> pts;
 public:
  template <class P> Index(const Matrix<T>& m, const P&) : pts(m) {}
  void buildIndex(const Matrix<T>& m) { pts = m; }
  int radiusSearch(const Matrix<T>& q, std::vector<std::vector<int> >& indices, std::vector<std::vector<T> >& dists,
                   float radius, const SearchParams&) {
    // NB: the real signature takes `float radius`; the reference passes searchR2 = 2.0 (exact in float)
    indices.assign(q.rows, std::vector<int>());
    dists.assign(q.rows, std::vector<T>());
    int total = 0;
    for (size_t r = 0; r < q.rows; r++) {
      std::vector<std::pair<T, int> > found;
      for (size_t i = 0; i < pts.rows; i++) {
        T d = 0;
        for (size_t k = 0; k < pts.cols; k++) {
          T diff = q[r][k] - pts[i][k];
          d += diff * diff;
        }
        if (d < (T)radius) found.push_back(std::make_pair(d, (int)i));
      }
      std::sort(found.begin(), found.end());
      for (auto& f : found) {
        indices[r].push_back(f.second);
        dists[r].push_back(f.first);
      }
      total += (int)found.size();
    }
    return total;
  }
};
}  // namespace flann
