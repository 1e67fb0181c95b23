// Obj.h — drop-in for raytracer/Obj.h: Obj, PlaneObj, SphereObj.  The classes hold state and
// flatten themselves into SppmObject; intersection runs on the device.  GetIntersect stays
// callable for API compatibility and answers through the C ABI (sppm_intersect_batch on a
// one-object scene) — it is not a CPU implementation.
#pragma once
#include <cstring>

#include "Element.h"
#include "SppmBridge.h"
#include "Vec3.h"

class Obj {
 protected:
  Texture* texture = nullptr;
 public:
  Obj(const Material& mm, const Color& cc) : material(mm), color(cc) {
    diffr = mm.diff.GetPower();
    reflr = mm.refl.GetPower();
    refrr = mm.refr.GetPower();
    allr = diffr + reflr + refrr;
  }
  Material material;
  bool IsDiff() { return material.IsDiff(); }
  bool IsRefl() { return material.IsRefl(); }
  bool IsRefr() { return material.IsRefr(); }
  void GetDiff(Vec3& v) const { v = material.diff; }
  void GetRefl(Vec3& v) const { v = material.refl; }
  void GetRefr(Vec3& v) const { v = material.refr; }
  void GetRefl(Vec3& v, double& rln) const { v = material.refl; rln = material.refln; }
  void GetRefr(Vec3& v, double& rrn) const { v = material.refr; rrn = material.refrn; }
  enum React { DIFF, REFL, REFR };
  double diffr, reflr, refrr, allr;
  React Roulette() {  // Obj.h:30-45
    double r = GetRandu(0., allr);
    if (diffr > r) return DIFF;
    else if (diffr + reflr > r) return REFL;
    else return REFR;
  }
  Color color;
  virtual ~Obj() {}
  void setTexture(Texture* texture_) { texture = texture_; }
  Texture* getTexture() const { return texture; }

  // flatten into the ABI record; Bezier objects append their patches
  virtual void flatten(SppmObject& o, std::vector<double>& patches) const = 0;
  void flattenCommon(SppmObject& o) const {
    std::memset(&o, 0, sizeof(o));
    o.tex_id = -1;
    sppm_host::put3(o.mat.diff, material.diff.val);
    sppm_host::put3(o.mat.refl, material.refl.val);
    sppm_host::put3(o.mat.refr, material.refr.val);
    o.mat.refrn = material.refrn;
    sppm_host::put3(o.color, color.val);
  }
  // Obj.h:48 — answered by the device through the C ABI
  virtual Collision GetIntersect(const Ray& ray) {
    sppm_host::FlatScene fs;
    fs.objs.resize(1);
    flatten(fs.objs[0], fs.patches);
    if (texture) {
      fs.objs[0].tex_id = 0;
      fs.tex.push_back(SppmTexture{texture->rows, texture->cols, texture->bgr.data()});
    }
    sppm_ctx* c = sppm_host::compat_ctx();
    SppmSceneDesc d = fs.desc();
    sppm_host::check(sppm_scene_upload(c, &d), c, "sppm_scene_upload");
    double r6[6] = {ray.org.val[0], ray.org.val[1], ray.org.val[2], ray.dir.val[0], ray.dir.val[1], ray.dir.val[2]};
    SppmHit h;
    sppm_host::check(sppm_intersect_batch(c, r6, 1, nullptr, 0, 0, &h), c, "sppm_intersect_batch");
    if (h.obj < 0) return NoneColl;
    Collision col(Vec3(h.pos[0], h.pos[1], h.pos[2]), ray.dir, h.t, this, Vec3(h.n[0], h.n[1], h.n[2]),
                  Vec3(h.color[0], h.color[1], h.color[2]));
    col.inside = h.inside != 0;
    return col;
  }
};

class PlaneObj : public Obj {  // Obj.h:55-101
 public:
  Vec3 P0;
  Vec3 norm;
  PlaneObj(const Vec3& pp, const Vec3& nn, const Color& cc, const Material& mm) : Obj(mm, cc), P0(pp), norm(nn) { norm.Normalize(); }
  Vec3 texU = Vec3(400, 0, 0);
  Vec3 texV = Vec3(0, 0, 300);
  void flatten(SppmObject& o, std::vector<double>&) const override {
    flattenCommon(o);
    o.kind = SPPM_OBJ_PLANE;
    sppm_host::put3(o.p0, P0.val);
    sppm_host::put3(o.norm, norm.val);
    sppm_host::put3(o.texU, texU.val);
    sppm_host::put3(o.texV, texV.val);
  }
};

class SphereObj : public Obj {  // Obj.h:102-154
 public:
  Vec3 org;
  double radius;
  Vec3 texU = Vec3(0, 3, -3).GetUnit(), texV = Vec3(1, 0, 0).GetUnit();
  SphereObj(double rradius, const Vec3& orgg, const Color& cc, const Material& mm) : Obj(mm, cc), org(orgg), radius(rradius) {}
  void flatten(SppmObject& o, std::vector<double>&) const override {
    flattenCommon(o);
    o.kind = SPPM_OBJ_SPHERE;
    sppm_host::put3(o.p0, org.val);
    o.radius = radius;
    sppm_host::put3(o.texU, texU.val);
    sppm_host::put3(o.texV, texV.val);
  }
};
