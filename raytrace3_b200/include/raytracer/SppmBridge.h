// SppmBridge.h — glue between the reference-shaped host classes and the C ABI
// (include/sppm_b200.h): error handling, the per-process device context used by the
// single-ray compatibility calls, and the flattened scene holder.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/sppm_b200.h"

namespace sppm_host {

inline void check(int rc, sppm_ctx* c, const char* what) {
  if (rc != SPPM_OK) throw std::runtime_error(std::string(what) + " failed [" + std::to_string(rc) + "]: " + sppm_last_error(c));
}
inline sppm_ctx* create_ctx(int device) {
  sppm_ctx* c = nullptr;
  int rc = sppm_ctx_create(device, &c);
  if (rc != SPPM_OK) throw std::runtime_error(std::string("sppm_ctx_create failed (no CPU fallback exists): ") + sppm_last_error(nullptr));
  return c;
}
// context for Scene::NearCollide / Obj::GetIntersect compatibility calls (device $SPPM_DEVICE or 0)
inline sppm_ctx* compat_ctx() {
  static sppm_ctx* c = create_ctx(std::getenv("SPPM_DEVICE") ? std::atoi(std::getenv("SPPM_DEVICE")) : 0);
  return c;
}
// Host-side phase times of the render loop (SPPM_HOST_TIMING=1): wall-clock seconds per named phase, summed over
// all threads and printed by SPPMRayTracer::render when it returns.  Off: one getenv per process.
inline bool host_timing_on() {
  static const bool on = std::getenv("SPPM_HOST_TIMING") && std::atoi(std::getenv("SPPM_HOST_TIMING")) != 0;
  return on;
}
inline std::map<std::string, std::pair<double, long>>& host_times() {
  static std::map<std::string, std::pair<double, long>> t;
  return t;
}
struct Phase {  // scope timer
  const char* name;
  std::chrono::steady_clock::time_point t0;
  explicit Phase(const char* n) : name(n) { if (host_timing_on()) t0 = std::chrono::steady_clock::now(); }
  ~Phase() {
    if (!host_timing_on()) return;
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
#pragma omp critical(sppm_host_times)
    {
      auto& e = host_times()[name];
      e.first += dt;
      e.second++;
    }
  }
};
inline void print_host_times() {
  if (!host_timing_on()) return;
  for (const auto& kv : host_times())
    std::printf("host phase %-22s %9.3f ms total  %6ld calls  %8.3f ms each\n", kv.first.c_str(), kv.second.first * 1e3, kv.second.second,
                kv.second.first * 1e3 / (double)kv.second.second);
}
// Rounds of `onetime` photons handed to one sppm_photon_pass call (RayTracer::PhotonMap): whole
// rounds, at most FUSE_PHOTONS photons per light and call (the per-call accumulators are int32 per
// hit point), one round when the rounds are not independent.  Python twin: api.rounds_per_call.
constexpr uint64_t FUSE_PHOTONS = 1ull << 28;
inline uint64_t rounds_per_call(uint64_t onetime, bool fuse) {
  if (!fuse || onetime == 0) return 1;
  const uint64_t r = FUSE_PHOTONS / onetime;
  return r < 1 ? 1 : r;
}
struct FlatScene {
  std::vector<SppmObject> objs;
  std::vector<double> patches;  // [n][16][3]
  std::vector<SppmLight> lights;
  std::vector<SppmTexture> tex;
  std::vector<const void*> tex_key;
  SppmSceneDesc desc() const {
    SppmSceneDesc d;
    d.n_objects = (int32_t)objs.size();
    d.n_patches = (int32_t)(patches.size() / 48);
    d.n_lights = (int32_t)lights.size();
    d.n_textures = (int32_t)tex.size();
    d.objects = objs.data();
    d.patches = patches.data();
    d.lights = lights.data();
    d.textures = tex.data();
    return d;
  }
};
inline void put3(double dst[3], const double src[3]) { dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; }

}  // namespace sppm_host
