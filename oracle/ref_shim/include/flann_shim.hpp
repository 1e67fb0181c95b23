// flann_shim.hpp — flann::Index<L2<double>> as the reference uses it (Raytracer.h:92-96, 146,
// 378-380): exact radius search on squared L2 distance, strict '<' (pinned against
// cv2.flann_Index in tests/test_oracle_cpu.py), results sorted by distance (SearchParams
// default sorted = true).  TEST INFRASTRUCTURE (oracle/_ref build only).
//
// buildIndex makes a kd-tree the way KDTreeSingleIndex does (median split on the widest axis,
// leaves of <= 10 points), so that the reference's own PhotonMap can also be TIMED with an index
// of the right complexity (bench.py, cpu_baseline.reference_code).  The search is exact: a
// subtree is skipped only when the squared distance to its splitting plane is already >= the
// radius, and the leaf test computes the distance exactly as the exhaustive scan does
// (sequential sum over the axes), so tree and scan return the same (index, distance) lists —
// ref_flann_selfcheck in the harness compares them.
#pragma once
#include <algorithm>
#include <numeric>
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
  struct Node {
    int lo, hi;        // range of perm (leaf) — [lo, hi)
    int dim;           // split axis, -1 for a leaf
    T split;           // left: coordinate <= split, right: coordinate >= split
    int left, right;
  };
  Matrix<T> pts;
  std::vector<int> perm;
  std::vector<Node> nodes;
  enum { LEAF = 10 };

  int build(int lo, int hi) {
    int id = (int)nodes.size();
    nodes.push_back(Node{lo, hi, -1, T(0), -1, -1});
    if (hi - lo <= LEAF) return id;
    int dim = 0;
    T best = T(-1);
    for (size_t k = 0; k < pts.cols; k++) {
      T mn = pts[perm[lo]][k], mx = mn;
      for (int i = lo + 1; i < hi; i++) {
        T v = pts[perm[i]][k];
        if (v < mn) mn = v;
        if (v > mx) mx = v;
      }
      if (mx - mn > best) { best = mx - mn; dim = (int)k; }
    }
    if (!(best > T(0))) return id;  // all points coincide: keep them in one leaf
    int mid = (lo + hi) / 2;
    std::nth_element(perm.begin() + lo, perm.begin() + mid, perm.begin() + hi,
                     [&](int a, int b) { return pts[a][dim] < pts[b][dim]; });
    T split = pts[perm[mid]][dim];
    int l = build(lo, mid);
    int r = build(mid, hi);
    nodes[id].dim = dim;
    nodes[id].split = split;
    nodes[id].left = l;
    nodes[id].right = r;
    return id;
  }
  T dist2(const T* q, size_t i) const {
    T d = 0;
    for (size_t k = 0; k < pts.cols; k++) {
      T diff = q[k] - pts[i][k];
      d += diff * diff;
    }
    return d;
  }
  void search(int id, const T* q, T radius, std::vector<std::pair<T, int> >& found) const {
    const Node& n = nodes[id];
    if (n.dim < 0) {
      for (int i = n.lo; i < n.hi; i++) {
        T d = dist2(q, (size_t)perm[i]);
        if (d < radius) found.push_back(std::make_pair(d, perm[i]));
      }
      return;
    }
    T diff = q[n.dim] - n.split;
    int near = diff < T(0) ? n.left : n.right, far = diff < T(0) ? n.right : n.left;
    search(near, q, radius, found);
    // every point of the far side is at least |diff| away along this axis (rounding is monotone,
    // so its computed squared distance is >= diff * diff)
    if (diff * diff < radius) search(far, q, radius, found);
  }

 public:
  template <class P> Index(const Matrix<T>& m, const P&) : pts(m) {}
  void buildIndex(const Matrix<T>& m) {
    pts = m;
    perm.resize(m.rows);
    std::iota(perm.begin(), perm.end(), 0);
    nodes.clear();
    if (m.rows) build(0, (int)m.rows);
  }
  // the exhaustive scan the tree is checked against
  void scan(const T* q, T radius, std::vector<std::pair<T, int> >& found) const {
    for (size_t i = 0; i < pts.rows; i++) {
      T d = dist2(q, i);
      if (d < radius) found.push_back(std::make_pair(d, (int)i));
    }
  }
  int radiusSearch(const Matrix<T>& q, std::vector<std::vector<int> >& indices, std::vector<std::vector<T> >& dists,
                   float radius, const SearchParams&, bool exhaustive = false) {
    // NB: the real signature takes `float radius`; the reference passes searchR2 = 2.0 (exact in float)
    indices.assign(q.rows, std::vector<int>());
    dists.assign(q.rows, std::vector<T>());
    int total = 0;
    std::vector<std::pair<T, int> > found;
    for (size_t r = 0; r < q.rows; r++) {
      found.clear();
      if (exhaustive || nodes.empty()) scan(q[r], (T)radius, found);
      else search(0, q[r], (T)radius, found);
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
