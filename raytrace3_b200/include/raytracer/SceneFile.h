// SceneFile.h — scene description files and render jobs.
//
// The reference hard-codes its scene in Scene::init (Scene.h:98-164) and its camera, film size
// and loop counts in main() (main.cpp:19-41) and Raytracer.h:218,384,425.  This header reads the
// same information from a line-based text file, so BASELINE configs 2-5 run without recompiling:
//
//   film W H
//   camera px py pz  lx ly lz  [scale]
//   material NAME diff r g b refl r g b refr r g b [refrn x]
//   texture NAME FILE                     (.jpg/.jpeg decoded by nvJPEG, or .bgr8 raw pixels)
//   plane  p0 xyz  normal xyz  color rgb  MATERIAL [texture NAME]
//   sphere radius  centre xyz  color rgb  MATERIAL [texture NAME]
//   bezier FILE.bpt  centre xyz  scale  color rgb  MATERIAL [rot teapot | rot m00 .. m22] [texture NAME]
//   light  pos xyz  color rgb
//   render [passes N] [replicas N] [rounds N] [photons N] [seed S] [shared 0|1] [newton exact|fast]
//          [gather exact|filtered] [math exact|fast] [jitter x]
//   output [prefix P] [format ppm|jpg] [quality Q] [checkpoint FILE] [every N]
//
// '#' starts a comment.  File names are relative to the scene file.  Objects appear in objvec in
// file order (Scene::NearCollide keeps the first of equal distances, Scene.h:170-179).
#pragma once
#include <cmath>
#include <cstdint>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "Camera.h"
#include "Scene.h"

namespace sppm_host {

struct RenderJob {
  int w = 320, h = 240;
  Vec3 cam_pos = Vec3(50, 35, 230), cam_look = Vec3(50, 35.042612, 229);
  double cam_scale = 1.0;
  int passes = 1;          // SPPMRayTracer::sppmtime (Raytracer.h:425)
  int replicas = 4;        // Raytracer.h:426-429
  int rounds = 10;         // RayTracer cnt (Raytracer.h:384)
  int photons = 100000;    // photons per light and round (Raytracer.h:218 onetime)
  uint64_t seed = 0x5EED5EEDull;
  bool shared = false;     // shared per-pixel statistics
  int newton_mode = SPPM_NEWTON_EXACT, gather_mode = SPPM_GATHER_FILTERED, math_mode = SPPM_MATH_EXACT;
  double jitter = 0.00015;
  std::string out_prefix = "aaanew", format = "ppm", checkpoint;
  int quality = 95, checkpoint_every = 1;
};

inline cv::Matx33d teapot_rotation() {  // Scene.h:142-152: Trans2 * Trans, theta = (90/180) PI
  cv::Matx33d Trans;
  Trans(0, 0) = 1;
  Trans(1, 2) = 1;
  Trans(2, 1) = 1;
  cv::Matx33d Trans2;
  const double theta = (90. / 180.) * PI;
  Trans2(0, 0) = cos(theta);
  Trans2(0, 2) = sin(theta);
  Trans2(1, 1) = 1;
  Trans2(2, 0) = -sin(theta);
  Trans2(2, 2) = cos(theta);
  return Trans2 * Trans;
}

// Parse `path` into scene (objects, lights; textures decoded) and job.  Throws std::runtime_error
// with the line number on malformed input.
inline void load_scene_file(const std::string& path, Scene& scene, RenderJob& job) {
  std::ifstream in(path.c_str());
  if (!in) throw std::runtime_error("scene file: cannot open " + path);
  const size_t slash = path.find_last_of('/');
  const std::string dir = slash == std::string::npos ? std::string("") : path.substr(0, slash + 1);
  std::map<std::string, Material> materials;
  std::map<std::string, Texture*> textures;
  std::string line;
  int ln = 0;
  auto bad = [&](const std::string& why) { throw std::runtime_error(path + ":" + std::to_string(ln) + ": " + why); };
  while (std::getline(in, line)) {
    ln++;
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line.erase(hash);
    std::istringstream ss(line);
    std::string kw;
    if (!(ss >> kw)) continue;
    auto num = [&]() { double v; if (!(ss >> v)) bad("number expected in '" + kw + "'"); return v; };
    auto vec = [&]() { const double x = num(), y = num(), z = num(); return Vec3(x, y, z); };
    auto word = [&]() { std::string w; if (!(ss >> w)) bad("word expected in '" + kw + "'"); return w; };
    auto material = [&]() {
      const std::string n = word();
      auto it = materials.find(n);
      if (it == materials.end()) bad("unknown material " + n);
      return it->second;
    };
    auto options = [&](Obj* o, cv::Matx33d* rot) {  // trailing [rot ...] [texture NAME]
      std::string opt;
      while (ss >> opt) {
        if (opt == "texture") {
          const std::string n = word();
          auto it = textures.find(n);
          if (it == textures.end()) bad("unknown texture " + n);
          if (it->second) o->setTexture(it->second);
        } else if (opt == "rot" && rot) {
          std::string first = word();
          if (first == "teapot") *rot = teapot_rotation();
          else {
            (*rot)(0, 0) = std::stod(first);
            for (int k = 1; k < 9; k++) (*rot)(k / 3, k % 3) = num();
          }
        } else bad("unknown option " + opt);
      }
    };
    if (kw == "film") {
      job.w = (int)num();
      job.h = (int)num();
    } else if (kw == "camera") {
      job.cam_pos = vec();
      job.cam_look = vec();
      double s;
      if (ss >> s) job.cam_scale = s;
    } else if (kw == "material") {
      const std::string n = word();
      Vec3 d, s, r;
      double rn = 1.5;
      std::string k;
      while (ss >> k) {
        if (k == "diff") d = vec();
        else if (k == "refl") s = vec();
        else if (k == "refr") r = vec();
        else if (k == "refrn") rn = num();
        else bad("unknown material field " + k);
      }
      materials.erase(n);
      materials.insert(std::make_pair(n, Material(d, s, r, rn)));
    } else if (kw == "texture") {
      const std::string n = word(), f = word();
      Texture* t = nullptr;
      try {
        t = new Texture(dir + f);
      } catch (const std::exception& e) {
        std::cerr << "warning: " << e.what() << " — objects using '" << n << "' stay untextured" << std::endl;
      }
      textures[n] = t;
    } else if (kw == "plane") {
      const Vec3 p0 = vec(), nn = vec(), cc = vec();
      Obj* o = new PlaneObj(p0, nn, cc, material());
      options(o, nullptr);
      scene.objvec.push_back(o);
    } else if (kw == "sphere") {
      const double r = num();
      const Vec3 c = vec(), cc = vec();
      Obj* o = new SphereObj(r, c, cc, material());
      options(o, nullptr);
      scene.objvec.push_back(o);
    } else if (kw == "bezier") {
      const std::string f = word();
      const Vec3 c = vec();
      const double scale = num();
      const Vec3 cc = vec();
      const Material m = material();
      // the options follow the constructor arguments, but the rotation is a constructor argument:
      // parse them into a placeholder first
      cv::Matx33d rot = cv::Matx33d::eye();
      PlaneObj tmp(Vec3(), Vec3(0, 0, 1), cc, m);
      options(&tmp, &rot);
      Obj* o = new Bezier3Obj(c, scale, dir + f, cc, m, rot);
      if (tmp.getTexture()) o->setTexture(tmp.getTexture());
      scene.objvec.push_back(o);
    } else if (kw == "light") {
      const Vec3 p = vec(), c = vec();
      scene.plvec.push_back(new SpotLight(p, c));
    } else if (kw == "render") {
      std::string k;
      while (ss >> k) {
        if (k == "passes") job.passes = (int)num();
        else if (k == "replicas") job.replicas = (int)num();
        else if (k == "rounds") job.rounds = (int)num();
        else if (k == "photons") job.photons = (int)num();
        else if (k == "seed") job.seed = std::stoull(word(), nullptr, 0);
        else if (k == "shared") job.shared = num() != 0;
        else if (k == "jitter") job.jitter = num();
        else if (k == "newton") { const std::string v = word(); job.newton_mode = v == "fast" ? SPPM_NEWTON_FAST : SPPM_NEWTON_EXACT; }
        else if (k == "gather") { const std::string v = word(); job.gather_mode = v == "exact" ? SPPM_GATHER_EXACT : SPPM_GATHER_FILTERED; }
        else if (k == "math") { const std::string v = word(); job.math_mode = v == "fast" ? SPPM_MATH_FAST : SPPM_MATH_EXACT; }
        else bad("unknown render field " + k);
      }
    } else if (kw == "output") {
      std::string k;
      while (ss >> k) {
        if (k == "prefix") job.out_prefix = word();
        else if (k == "format") job.format = word();
        else if (k == "quality") job.quality = (int)num();
        else if (k == "checkpoint") job.checkpoint = word();
        else if (k == "every") job.checkpoint_every = (int)num();
        else bad("unknown output field " + k);
      }
    } else bad("unknown keyword " + kw);
  }
  if (scene.objvec.empty()) throw std::runtime_error(path + ": no objects");
  if (scene.plvec.empty()) throw std::runtime_error(path + ": no lights");
}

}  // namespace sppm_host
