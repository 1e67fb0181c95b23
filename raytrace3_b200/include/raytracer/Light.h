// Light.h — drop-in for raytracer/Light.h.  emit() stays callable on the host; the render
// path draws photons on the device (csrc/trace.cu photon_trace_kernel).
#pragma once
#include "Element.h"
#include "Vec3.h"
class Light {
 public:
  Light(Vec3 p, Color c) : pos(p), color(c) {}
  virtual ~Light() {}
  virtual void emit(Ray& ray, Vec3& flux) {  // Light.h:8-13
    flux = color * PI * 4;
    ray.dir = Vec3::GetUnitRandVec();
    ray.org = pos;
  }
  const Vec3& GetPos() const { return pos; }
  const Color& GetColor() const { return color; }
 private:
  Vec3 pos;
  Color color;
};
class SpotLight : public Light {
 public:
  SpotLight(Vec3 p, Color c) : Light(p, c) {}
};
