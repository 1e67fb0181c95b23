// Bezier.h — drop-in for raytracer/Bezier.h: AABBbox, Bezier3, Bezier3Obj.  The .bpt loader
// (Bezier.h:199-239) stays on the host (without rebinding stdin); the ray-patch Newton solver
// and the AABB culling run on the device (csrc/scene.cuh).
#pragma once
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "Element.h"
#include "Obj.h"
#include "Vec3.h"
#define MAX_ITER 10

class AABBbox {  // Bezier.h:7-57 (state only; the entry-face test runs on the device)
 public:
  Vec3 pmin = MAX_VEC3, pmax = MIN_VEC3;
  void addpoint(const Vec3& p) {
    for (int i = 0; i < 3; i++) {
      if (pmin.val[i] > p.val[i]) pmin.val[i] = p.val[i];
      if (pmax.val[i] < p.val[i]) pmax.val[i] = p.val[i];
    }
  }
  void clear() { pmin = MAX_VEC3; pmax = MIN_VEC3; }
};

class Bezier3 {  // Bezier.h:59-187
 public:
  Vec3 P[16];
  AABBbox aabb;
  void read(const Vec3 p[16]) {
    aabb.clear();
    for (int i = 0; i < 16; i++) {
      P[i] = p[i];
      aabb.addpoint(P[i]);
    }
  }
};

class Bezier3Obj : public Obj {  // Bezier.h:188-289
  std::vector<Bezier3> Bvec;
  AABBbox aabb;
 public:
  Bezier3Obj(const Vec3& C, double scale, const std::string& file, const Color& cc, const Material& mm,
             const cv::Matx33d& Trans = cv::Matx33d::eye()) : Obj(mm, cc) {
    aabb.clear();
    Bvec.clear();
    std::cerr << "read bezier3::" << file << std::endl;
    std::ifstream in(file.c_str());
    if (!in) throw std::runtime_error("Bezier3Obj: cannot open " + file);
    int size;
    in >> size;
    for (int i = 0; i < size; i++) {
      int m, n;
      in >> m >> n;
      if ((m + 1) * (n + 1) > 16) throw std::runtime_error("Bezier3Obj: only bicubic patches are supported");
      Vec3 vtmp[16];
      for (int j = 0; j < m + 1; j++)
        for (int k = 0; k < n + 1; k++) {
          int newv = j * (n + 1) + k;
          for (int l = 0; l < 3; l++) {
            in >> vtmp[newv].val[l];
            vtmp[newv].val[l] *= scale;
          }
          vtmp[newv] = Trans * vtmp[newv];
        }
      for (int q = 0; q < 16; q++) {
        vtmp[q] += C;
        aabb.addpoint(vtmp[q]);
      }
      Bezier3 bezier3;
      bezier3.read(vtmp);
      Bvec.push_back(bezier3);
    }
    std::cerr << "done!size::" << Bvec.size() << std::endl;
    std::cerr << "Bezier::aabb" << aabb.pmax << aabb.pmin << std::endl;
  }
  size_t patchCount() const { return Bvec.size(); }
  void flatten(SppmObject& o, std::vector<double>& patches) const override {
    flattenCommon(o);
    o.kind = SPPM_OBJ_BEZIER;
    o.first_patch = (int32_t)(patches.size() / 48);
    o.n_patch = (int32_t)Bvec.size();
    for (const Bezier3& b : Bvec)
      for (int q = 0; q < 16; q++)
        for (int l = 0; l < 3; l++) patches.push_back(b.P[q].val[l]);
  }
};
