// Light.h — host mirror of the reference's point emitter (raytracer/Light.h).  A light is the POD the
// device consumes (SppmLight: position + power); the class only wraps it with the reference's
// constructor and emit() signature so that scene code written against the reference compiles.
// Photons of the render path are emitted on the device (csrc/wavefront.cu, bounce 0).
#pragma once
#include "Element.h"
#include "SppmBridge.h"
#include "Vec3.h"

class Light {
  SppmLight desc_{};

 public:
  Light(const Vec3& where, const Color& power) {
    sppm_host::put3(desc_.pos, where.val);
    sppm_host::put3(desc_.color, power.val);
  }
  virtual ~Light() = default;

  Vec3 GetPos() const { return Vec3(desc_.pos[0], desc_.pos[1], desc_.pos[2]); }
  Color GetColor() const { return Color(desc_.color[0], desc_.color[1], desc_.color[2]); }
  const SppmLight& descriptor() const { return desc_; }

  // One photon as the reference draws it (Light.h:8-13): isotropic direction, flux = 4 PI power.
  // Host-side compatibility only; uses the host's stand-in generator (Vec3.h).
  virtual void emit(Ray& photon, Vec3& flux) {
    photon.org = GetPos();
    photon.dir = Vec3::GetUnitRandVec();
    flux = GetColor() * PI * 4;
  }
};

// The reference's only concrete light; it adds nothing to Light (Light.h:20-27).
class SpotLight : public Light {
 public:
  using Light::Light;
};
