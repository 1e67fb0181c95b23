// Scene.h — drop-in for raytracer/Scene.h: the objvec/plvec container, the hard-coded scene of
// Scene::init (Scene.h:98-164) and NearCollide.  flatten() produces the SppmSceneDesc the
// device consumes.
#pragma once
#include <cstdlib>
#include <string>
#include <vector>

#include "Bezier.h"
#include "Camera.h"
#include "Light.h"
#include "Obj.h"
#include "SppmBridge.h"

class Scene {
 public:
  std::vector<Obj*> objvec;
  std::vector<Light*> plvec;

  // Scene.h:98-164.  Assets (teapot.bpt and the textures pre-decoded to .bgr8 by
  // tools/decode_texture.py) are looked up in $SPPM_ASSET_DIR (default ".").  The reference
  // also loads blue.jpg for the teapot, which is absent upstream (Q18): the teapot stays
  // untextured unless blue.bgr8 exists.
  void init() {
    const char* env = std::getenv("SPPM_ASSET_DIR");
    const std::string dir = std::string(env ? env : ".") + "/";
    Material WHITE_DIFF(Vec3(.75, .75, .75), Vec3(), Vec3(), 0);
    Material MIRROR(Vec3(), Vec3(1, 1, 1) * .999, Vec3(), 0);
    Material REFR0(Vec3(), Vec3(), Vec3(.999, .999, .999), 1.5);
    Obj* obj[] = {
        new PlaneObj(Vec3(1, 40.8, 81.6), Vec3(1, 0, 0), Vec3(.75, .25, .25), WHITE_DIFF),   // Left
        new PlaneObj(Vec3(99, 40.8, 81.6), Vec3(1, 0, 0), Vec3(.25, .25, .75), WHITE_DIFF),  // Right
        new PlaneObj(Vec3(50, 40.8, 0), Vec3(0, 0, 1), Vec3(.75, 0.75, .75), MIRROR),        // Back
        new PlaneObj(Vec3(50, 0, 81.6), Vec3(0, -1, 0), Vec3(.75, .75, .75), WHITE_DIFF),    // Bottom
        new PlaneObj(Vec3(50, 81.6, 81.6), Vec3(0, 1, 0), Vec3(.75, .75, .75), WHITE_DIFF),  // Top
        new SphereObj(16.5, Vec3(27, 16.5, 47), Vec3(1, 1, 1) * .999, MIRROR),               // Mirror
        new SphereObj(16.5, Vec3(73, 16.5, 88), Vec3(1, 1, 1) * .999, REFR0),                // Glass
        new SphereObj(8.5, Vec3(50, 8.5, 60), Vec3(1, 1, 1) * .999, WHITE_DIFF)};
    for (int i = 0; i < 8; i++) objvec.push_back(obj[i]);
    Texture* timg = tryTexture(dir + "timg.bgr8");
    Texture* stone = tryTexture(dir + "wall.bgr8");
    if (stone) {
      objvec[0]->setTexture(stone);
      objvec[1]->setTexture(stone);
      objvec[4]->setTexture(stone);
    }
    Texture* planet = tryTexture(dir + "planet.bgr8");
    if (planet) objvec[7]->setTexture(planet);
    if (timg) objvec[3]->setTexture(timg);
    cv::Matx33d Trans;
    Trans(0, 0) = 1;
    Trans(1, 2) = 1;
    Trans(2, 1) = 1;
    cv::Matx33d Trans2;
    double theta = (90. / 180.) * PI;
    Trans2(0, 0) = cos(theta);
    Trans2(0, 2) = sin(theta);
    Trans2(1, 1) = 1;
    Trans2(2, 0) = -sin(theta);
    Trans2(2, 2) = cos(theta);
    Bezier3Obj* test = new Bezier3Obj(Vec3(20, 0, 120), 4, dir + "teapot.bpt", Vec3(0, 1, 1) * .999, WHITE_DIFF, Trans2 * Trans);
    objvec.push_back(test);
    Texture* blue = tryTexture(dir + "blue.bgr8", /*quiet=*/true);
    if (blue) objvec[8]->setTexture(blue);
    plvec.push_back(new SpotLight(Vec3(50, 60, 85), 5000 * Color(1, 1, 1)));
  }

  void flatten(sppm_host::FlatScene& fs) const {
    fs = sppm_host::FlatScene();
    fs.objs.resize(objvec.size());
    for (size_t i = 0; i < objvec.size(); i++) {
      objvec[i]->flatten(fs.objs[i], fs.patches);
      if (Texture* t = objvec[i]->getTexture()) {
        size_t k = 0;
        while (k < fs.tex_key.size() && fs.tex_key[k] != t) k++;
        if (k == fs.tex_key.size()) {
          fs.tex_key.push_back(t);
          fs.tex.push_back(SppmTexture{t->rows, t->cols, t->bgr.data()});
        }
        fs.objs[i].tex_id = (int32_t)k;
      }
    }
    for (Light* l : plvec) {
      SppmLight sl;
      sppm_host::put3(sl.pos, l->GetPos().val);
      sppm_host::put3(sl.color, l->GetColor().val);
      fs.lights.push_back(sl);
    }
  }

  // Scene.h:165-182 — answered by the device (compatibility call, one ray per launch)
  Collision NearCollide(const Ray& ray) {
    sppm_host::FlatScene fs;
    flatten(fs);
    sppm_ctx* c = sppm_host::compat_ctx();
    SppmSceneDesc d = fs.desc();
    sppm_host::check(sppm_scene_upload(c, &d), c, "sppm_scene_upload");
    double r6[6] = {ray.org.val[0], ray.org.val[1], ray.org.val[2], ray.dir.val[0], ray.dir.val[1], ray.dir.val[2]};
    SppmHit h;
    sppm_host::check(sppm_intersect_batch(c, r6, 1, nullptr, 0, 0, &h), c, "sppm_intersect_batch");
    if (h.obj < 0) return Collision();
    Collision col(Vec3(h.pos[0], h.pos[1], h.pos[2]), ray.dir, h.t, objvec[h.obj], Vec3(h.n[0], h.n[1], h.n[2]),
                  Vec3(h.color[0], h.color[1], h.color[2]));
    col.inside = h.inside != 0;
    return col;
  }

 private:
  static Texture* tryTexture(const std::string& path, bool quiet = false) {
    try {
      return new Texture(path);
    } catch (const std::exception& e) {
      if (!quiet) std::cerr << "warning: " << e.what() << " — object left untextured" << std::endl;
      return nullptr;
    }
  }
};
