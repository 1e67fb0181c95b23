// scene_prep.h — host-side flattening of a SppmSceneDesc into what the kernels read: per-patch
// and per-object AABBs, the cull tables of the Newton solves (13-DOP, sub-patch boxes), the
// monomial patch coefficients of the FAST Newton mode and the DevObject records.  Pure host C++
// (no CUDA calls), used by sppm_scene_upload (abi.cu); tests/device_host compiles it with g++ to
// check these tables and the device arithmetic of scene.cuh on the CPU.
// Citations: /root/reference/raytracer/<file>:<line>.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "scene.cuh"

namespace sppm {

struct PreppedScene {
  std::vector<double> patchP;    // [n_patch][48]
  std::vector<double> patchBox;  // [n_patch][6]
  std::vector<double> patchDop;  // [n_patch][26]
  std::vector<double> patchSub;  // [n_patch][16][6]
  std::vector<double> patchA;    // [n_patch][48]
  std::vector<DevObject> objs;   // [n_obj]
};

// Counts and array pointers of `s` are already validated by the caller.  Returns nullptr, or the
// message of the SPPM_ERR_INVALID the scene is rejected with.
inline const char* prep_scene(const SppmSceneDesc* s, PreppedScene& out) {
  // per-patch AABBs (Bezier.h:165-175) and per-object AABBs (Bezier.h:219-222), addpoint semantics
  std::vector<double>& patchP = out.patchP;
  std::vector<double>& patchBox = out.patchBox;
  patchP.assign((size_t)s->n_patches * 48, 0.0);
  patchBox.assign((size_t)s->n_patches * 6, 0.0);
  for (int p = 0; p < s->n_patches; p++) {
    double* bx = &patchBox[(size_t)p * 6];
    for (int k = 0; k < 3; k++) {
      bx[k] = SPPM_MAX_NUM;
      bx[3 + k] = SPPM_MIN_NUM;
    }
    for (int q = 0; q < 16; q++)
      for (int k = 0; k < 3; k++) {
        const double v = s->patches[((size_t)p * 16 + q) * 3 + k];
        patchP[((size_t)p * 16 + q) * 3 + k] = v;
        if (bx[k] > v) bx[k] = v;
        if (bx[3 + k] < v) bx[3 + k] = v;
      }
  }
  // 13-DOP of every patch's control points, grown by the Newton acceptance radius (scene.cuh dop_line_miss)
  std::vector<double>& patchDop = out.patchDop;
  patchDop.assign((size_t)s->n_patches * 26, 0.0);
  {
    static const double A13[13][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, -1, 0}, {1, 0, 1}, {1, 0, -1},
                                      {0, 1, 1}, {0, 1, -1}, {1, 1, 1}, {1, 1, -1}, {1, -1, 1}, {-1, 1, 1}};
    for (int p = 0; p < s->n_patches; p++)
      for (int k = 0; k < 13; k++) {
        double lo = 1e300, hi = -1e300, big = 0;
        for (int q = 0; q < 16; q++) {
          const double* v = &patchP[((size_t)p * 16 + q) * 3];
          const double e = A13[k][0] * v[0] + A13[k][1] * v[1] + A13[k][2] * v[2];
          lo = std::min(lo, e);
          hi = std::max(hi, e);
          big = std::max(big, std::fabs(v[0]) + std::fabs(v[1]) + std::fabs(v[2]));
        }
        const double len = std::sqrt(A13[k][0] * A13[k][0] + A13[k][1] * A13[k][1] + A13[k][2] * A13[k][2]);
        const double grow = 0.0101 * len + 1e-12 * big;  // sqrt(M_EPS) = 0.01 plus slack for rounding
        patchDop[(size_t)p * 26 + k] = lo - grow;
        patchDop[(size_t)p * 26 + 13 + k] = hi + grow;
      }
  }
  // AABBs of the 4x4 de Casteljau sub-patches (scene.cuh sub_boxes_miss), grown by the same radius
  std::vector<double>& patchSub = out.patchSub;
  patchSub.assign((size_t)s->n_patches * 96, 0.0);
  {
    // split a cubic Bezier curve c[4] at 1/4, 1/2, 3/4 into four cubic segments seg[4][4] (de Casteljau twice)
    auto split4 = [](const double c[4], double seg[4][4]) {
      auto halve = [](const double a[4], double l[4], double r[4]) {
        const double b0 = 0.5 * (a[0] + a[1]), b1 = 0.5 * (a[1] + a[2]), b2 = 0.5 * (a[2] + a[3]);
        const double c0 = 0.5 * (b0 + b1), c1 = 0.5 * (b1 + b2), d0 = 0.5 * (c0 + c1);
        l[0] = a[0]; l[1] = b0; l[2] = c0; l[3] = d0;
        r[0] = d0; r[1] = c1; r[2] = b2; r[3] = a[3];
      };
      double l[4], r[4];
      halve(c, l, r);
      halve(l, seg[0], seg[1]);
      halve(r, seg[2], seg[3]);
    };
    for (int p = 0; p < s->n_patches; p++) {
      double lo[16][3], hi[16][3], big = 0;
      for (int b = 0; b < 16; b++)
        for (int k = 0; k < 3; k++) {
          lo[b][k] = 1e300;
          hi[b][k] = -1e300;
        }
      for (int k = 0; k < 3; k++) {
        // rows (fixed v index i) split along u, then the columns of every u-segment split along v
        double rowseg[4][4][4];  // [i][su][j]
        for (int i = 0; i < 4; i++) {
          double c[4];
          for (int j = 0; j < 4; j++) {
            c[j] = patchP[((size_t)p * 16 + 4 * i + j) * 3 + k];
            big = std::max(big, std::fabs(c[j]));
          }
          split4(c, rowseg[i]);
        }
        for (int su = 0; su < 4; su++)
          for (int j = 0; j < 4; j++) {
            double c[4], colseg[4][4];
            for (int i = 0; i < 4; i++) c[i] = rowseg[i][su][j];
            split4(c, colseg);
            for (int sv = 0; sv < 4; sv++)
              for (int i = 0; i < 4; i++) {
                lo[su * 4 + sv][k] = std::min(lo[su * 4 + sv][k], colseg[sv][i]);
                hi[su * 4 + sv][k] = std::max(hi[su * 4 + sv][k], colseg[sv][i]);
              }
          }
      }
      const double grow = 0.0101 + 1e-12 * big;
      for (int b = 0; b < 16; b++)
        for (int k = 0; k < 3; k++) {
          patchSub[(size_t)p * 96 + b * 6 + k] = lo[b][k] - grow;
          patchSub[(size_t)p * 96 + b * 6 + 3 + k] = hi[b][k] + grow;
        }
    }
  }
  // monomial coefficients per patch and component: A = M G M^T with G[i][k] = P[4i+k] (i <-> v,
  // k <-> u) and M the cubic Bernstein -> power basis matrix
  std::vector<double>& patchA = out.patchA;
  patchA.assign((size_t)s->n_patches * 48, 0.0);
  {
    static const double M[4][4] = {{1, 0, 0, 0}, {-3, 3, 0, 0}, {3, -6, 3, 0}, {-1, 3, -3, 1}};
    for (int p = 0; p < s->n_patches; p++)
      for (int k = 0; k < 3; k++) {
        double G[4][4], T[4][4];
        for (int i = 0; i < 4; i++)
          for (int j = 0; j < 4; j++) G[i][j] = patchP[((size_t)p * 16 + 4 * i + j) * 3 + k];
        for (int i = 0; i < 4; i++)      // T = M G  (v direction)
          for (int j = 0; j < 4; j++) {
            double t = 0;
            for (int q = 0; q < 4; q++) t += M[i][q] * G[q][j];
            T[i][j] = t;
          }
        for (int i = 0; i < 4; i++)      // A = T M^T (u direction)
          for (int j = 0; j < 4; j++) {
            double t = 0;
            for (int q = 0; q < 4; q++) t += T[i][q] * M[j][q];
            patchA[((size_t)p * 16 + 4 * i + j) * 3 + k] = t;
          }
      }
  }
  std::vector<DevObject>& objs = out.objs;
  objs.assign((size_t)s->n_objects, DevObject{});
  for (int i = 0; i < s->n_objects; i++) {
    const SppmObject& o = s->objects[i];
    DevObject d;
    memset(&d, 0, sizeof(d));
    d.kind = o.kind;
    d.tex_id = o.tex_id;
    d.first_patch = o.first_patch;
    d.n_patch = o.n_patch;
    if (o.kind < 0 || o.kind > 2) return "bad object kind";
    if (o.tex_id >= s->n_textures) return "tex_id out of range";
    if (o.kind == SPPM_OBJ_BEZIER && (o.first_patch < 0 || o.n_patch < 0 || o.first_patch + o.n_patch > s->n_patches))
      return "bezier patch range out of bounds";
    const d3 df = ld3(o.mat.diff), rl = ld3(o.mat.refl), rr = ld3(o.mat.refr);
    d.isDiff = !is_zero3(df);
    d.isRefl = !is_zero3(rl);
    d.isRefr = !is_zero3(rr);
    d.diffr = (df.x + df.y + df.z) / 3;  // Vec3.h:116-119
    d.reflr = (rl.x + rl.y + rl.z) / 3;
    d.refrr = (rr.x + rr.y + rr.z) / 3;
    d.allr = d.diffr + d.reflr + d.refrr;
    for (int k = 0; k < 3; k++) {
      d.diff[k] = o.mat.diff[k];
      d.refl[k] = o.mat.refl[k];
      d.refr[k] = o.mat.refr[k];
      d.color[k] = o.color[k];
      d.p0[k] = o.p0[k];
      d.norm[k] = o.norm[k];
      d.texU[k] = o.texU[k];
      d.texV[k] = o.texV[k];
    }
    d.refrn = o.mat.refrn;
    d.radius = o.radius;
    const d3 w = cross3(ld3(o.texU), ld3(o.texV));
    d.texW[0] = w.x; d.texW[1] = w.y; d.texW[2] = w.z;
    d.texUmod = module(ld3(o.texU));
    d.texVmod = module(ld3(o.texV));
    int ndir = 0;  // Obj.h:89-96
    for (int k = 0; k < 3; k++)
      if (o.norm[k] != 0) ndir = k;
    d.vdex = (ndir + 2) % 3;
    d.udex = (ndir + 1) % 3;
    for (int k = 0; k < 3; k++) {
      d.box[k] = SPPM_MAX_NUM;
      d.box[3 + k] = SPPM_MIN_NUM;
    }
    if (o.kind == SPPM_OBJ_BEZIER)
      for (int p = o.first_patch; p < o.first_patch + o.n_patch; p++)
        for (int k = 0; k < 3; k++) {
          d.box[k] = std::min(d.box[k], patchBox[(size_t)p * 6 + k]);
          d.box[3 + k] = std::max(d.box[3 + k], patchBox[(size_t)p * 6 + 3 + k]);
        }
    objs[i] = d;
  }
  return nullptr;
}

}  // namespace sppm
