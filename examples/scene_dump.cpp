// examples/scene_dump.cpp — test helper: load a scene description file with the C++ host mirror
// (SceneFile.h) and write the flattened scene the device would receive (SppmObject array, control
// points, lights, camera basis, job settings) as raw bytes, so that tests/test_scenefile_cpu.py can
// compare it with the Python twin byte for byte.  Runs without a GPU (textures must be .bgr8 or absent).
#include <cstdio>
#include <iostream>

#include "Vec3.h"
#include "Obj.h"
#include "Camera.h"
#include "Element.h"
#include "Light.h"
#include "Scene.h"
#include "SceneFile.h"

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cerr << "usage: scene_dump SCENE OUT.bin\n";
    return 2;
  }
  try {
    Scene scene;
    sppm_host::RenderJob job;
    sppm_host::load_scene_file(argv[1], scene, job);
    sppm_host::FlatScene fs;
    scene.flatten(fs);
    Camera cam;
    cam.init(job.w, job.h);
    cam.MoveTo(job.cam_pos);
    cam.lookAt(job.cam_look, job.cam_scale);
    const SppmCameraDesc cd = cam.desc();
    FILE* f = std::fopen(argv[2], "wb");
    if (!f) throw std::runtime_error("cannot write output");
    const int64_t hdr[4] = {(int64_t)fs.objs.size(), (int64_t)(fs.patches.size() / 48), (int64_t)fs.lights.size(), (int64_t)fs.tex.size()};
    std::fwrite(hdr, 8, 4, f);
    std::fwrite(fs.objs.data(), sizeof(SppmObject), fs.objs.size(), f);
    std::fwrite(fs.patches.data(), 8, fs.patches.size(), f);
    std::fwrite(fs.lights.data(), sizeof(SppmLight), fs.lights.size(), f);
    std::fwrite(&cd, sizeof(cd), 1, f);
    const int64_t jobv[10] = {job.w, job.h, job.passes, job.replicas, job.rounds, job.photons, (int64_t)job.seed, job.shared ? 1 : 0,
                              job.newton_mode * 100 + job.gather_mode * 10 + job.math_mode, job.quality};
    std::fwrite(jobv, 8, 10, f);
    std::fclose(f);
  } catch (const std::exception& e) {
    std::cerr << "scene_dump: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
