// Raytracer.h — drop-in for raytracer/Raytracer.h: HitPoint, RayTracer, SPPMRayTracer.
// render() keeps its signature; its body flattens the scene / camera and drives the sm_100a
// kernels through the C ABI (include/sppm_b200.h).  There is no CPU render path.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "Camera.h"
#include "Scene.h"
#include "SppmBridge.h"
using namespace std;

#define MIN_WGT 1e-3
#define MAX_DEP 13
#define INIT_R2 2.0

typedef unsigned char uchar;
inline int toInt(double x) { return int(pow(1 - exp(-x), 1 / 2.2) * 255 + .5); }  // Raytracer.h:24-26

const double alpha = 0.7;

struct HitPoint {  // Raytracer.h:47-80 (host mirror of one device hit point, see SyncHitPoints)
  Vec3 pos, n, v;
  Obj* obj = nullptr;
  int x = -1, y = -1;
  Color wgt;
  double R2 = INIT_R2;
  int N = 0;
  Color tao;
  int newN = 0;
};

class RayTracer {
 public:
  Camera* camera = nullptr;
  Scene* scene = nullptr;
  vector<HitPoint*> Hvec;  // filled on demand by SyncHitPoints(); the live map stays on the device
  int onetime = 100;       // Raytracer.h:218
  int cnt = 10000;         // Raytracer.h:384
  uint64_t pass_base = 0;  // distinguishes replicas in the counter RNG
  SppmParams params;
  // Photon k of round i is photon i*onetime + k of pass `pass_base` in the counter RNG.  While
  // HitPoint::update is the reference's -- a no-op for the hit points of a fresh eye pass (N = 0,
  // Raytracer.h:70, Q1) -- the rounds of PhotonMap are independent, so they are traced as one
  // photon pass of cnt*onetime photons (cut into batches of at most 8 Mi photons over the library's lanes)
  // instead of cnt launch sequences of onetime = 100 photons each: same photons, same newN, tao in
  // a different summation order.  PROGRESSIVE updates (or fuse_rounds = false) keep one pass per round.
  bool fuse_rounds = true;

  explicit RayTracer(sppm_ctx* shared = nullptr) : ctx(shared), own(shared == nullptr) { sppm_default_params(&params); }
  ~RayTracer() {
    clearHvec();
    if (own && ctx) sppm_ctx_destroy(ctx);
    camera = nullptr;
    scene = nullptr;
  }
  RayTracer(const RayTracer&) = delete;
  RayTracer& operator=(const RayTracer&) = delete;

  void GetHitPoint() {  // Raytracer.h:102-116
    ensure();
    {
      sppm_host::Phase ph("scene flatten+upload");
      sppm_host::FlatScene fs;
      scene->flatten(fs);
      SppmSceneDesc d = fs.desc();
      sppm_host::check(sppm_set_params(ctx, &params), ctx, "sppm_set_params");
      sppm_host::check(sppm_scene_upload(ctx, &d), ctx, "sppm_scene_upload");
    }
    sppm_host::Phase ph("eye pass");
    SppmCameraDesc cd = camera->desc();
    int64_t n = 0;
    sppm_host::check(sppm_eye_pass(ctx, &cd, pass_base, &n), ctx, "sppm_eye_pass");
    n_hitpoints = n;
    cout << "hitpoint::" << n << endl;
  }
  void PhotonMap(int cnt_) {  // Raytracer.h:210-295
    const bool fuse = fuse_rounds && params.update_mode == SPPM_UPDATE_REFERENCE;
    const uint64_t step = sppm_host::rounds_per_call((uint64_t)onetime, fuse);
    {
      sppm_host::Phase ph("photon passes (enqueue)");
      for (uint64_t i = 0; i < (uint64_t)cnt_; i += step) {
        const uint64_t nr = std::min<uint64_t>(step, (uint64_t)cnt_ - i);
        sppm_host::check(sppm_photon_pass(ctx, i * (uint64_t)onetime, nr * (uint64_t)onetime, pass_base), ctx, "sppm_photon_pass");
        sppm_host::check(sppm_allreduce(ctx), ctx, "sppm_allreduce");
        sppm_host::check(sppm_update(ctx), ctx, "sppm_update");
      }
    }
    // camera->pic is one page-locked double[h][w][3] block (Camera::film): the read-back lands in it directly
    sppm_host::Phase ph("resolve + read-back");
    sppm_host::check(sppm_resolve(ctx, cnt_, onetime, camera->film()), ctx, "sppm_resolve");
  }
  void render() {  // Raytracer.h:366-387
    GetHitPoint();
    {
      sppm_host::Phase ph("build index");
      sppm_host::check(sppm_build_index(ctx), ctx, "sppm_build_index");
    }
    PhotonMap(cnt);
  }
  // Copy the device hit-point map (positions, normals, weights, statistics) into Hvec.
  void SyncHitPoints() {
    clearHvec();
    int64_t n = n_hitpoints;
    vector<double> pos(3 * n), nn(3 * n), wgt(3 * n), r2(n), tao(3 * n);
    vector<int32_t> pix(2 * n), newN(n), N(n);
    sppm_host::check(sppm_export_hitpoints(ctx, n, pos.data(), nn.data(), wgt.data(), r2.data(), pix.data(), nullptr, &n), ctx, "sppm_export_hitpoints");
    sppm_host::check(sppm_export_stats(ctx, n, newN.data(), N.data(), r2.data(), tao.data()), ctx, "sppm_export_stats");
    for (int64_t i = 0; i < n; i++) {
      HitPoint* hp = new HitPoint();
      hp->pos = Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
      hp->n = Vec3(nn[3 * i], nn[3 * i + 1], nn[3 * i + 2]);
      hp->wgt = Vec3(wgt[3 * i], wgt[3 * i + 1], wgt[3 * i + 2]);
      hp->tao = Vec3(tao[3 * i], tao[3 * i + 1], tao[3 * i + 2]);
      hp->x = pix[2 * i];
      hp->y = pix[2 * i + 1];
      hp->R2 = r2[i];
      hp->N = N[i];
      hp->newN = newN[i];
      Hvec.push_back(hp);
    }
  }
  sppm_ctx* context() { ensure(); return ctx; }
  int64_t n_hitpoints = 0;

 private:
  sppm_ctx* ctx;
  bool own;
  void ensure() {
    if (!ctx) ctx = sppm_host::create_ctx(std::getenv("SPPM_DEVICE") ? std::atoi(std::getenv("SPPM_DEVICE")) : 0);
  }
  void clearHvec() {
    for (auto hp : Hvec) delete hp;
    vector<HitPoint*>().swap(Hvec);
  }
};

class SPPMRayTracer {  // Raytracer.h:421-477
 public:
  int sppmtime = 100000;      // Raytracer.h:425
  int replicas = 4;           // Raytracer.h:426-429
  int cnt = 10000, onetime = 100;
  double jitter = 0.00015;    // Raytracer.h:434
  std::vector<int> devices;   // GPUs the replicas are spread over (default: device 0)
  std::string out_prefix = "aaanew";
  std::vector<uchar> image;   // last tone-mapped image, BGR, row-flipped as the reference writes it
  std::vector<double> film;   // accumulated film [h][w][3] (camera is passed by value, as upstream)
  // false: the reference's estimator, the average of independent PPM renders (Raytracer.h:429-458).
  // true: one set of per-pixel statistics (R2, N, tau) shared by all eye passes while the camera
  // jitters -- SPPM proper (sppm_shared_*); runs on devices[0].
  bool shared_statistics = false;
  SppmParams params;                 // arithmetic modes / seed of every RayTracer of the loop
  bool fuse_rounds = true;           // see RayTracer::fuse_rounds
  std::string format = "ppm";        // "ppm" or "jpg" (nvJPEG, the reference's cv::imwrite target, Raytracer.h:474)
  int jpeg_quality = 95;
  // Film checkpoint: every `checkpoint_every` iterations the accumulated film (or
  // the shared statistics), the iteration count and the jitter generator are written to
  // `checkpoint_path`; with `resume` the loop continues from such a file.
  std::string checkpoint_path;
  int checkpoint_every = 0;
  bool resume = false;

  SPPMRayTracer() { sppm_default_params(&params); }

  void render(Scene scene, Camera camera) {
    if (devices.empty()) devices.push_back(std::getenv("SPPM_DEVICE") ? std::atoi(std::getenv("SPPM_DEVICE")) : 0);
    if (shared_statistics) {
      renderShared(scene, camera);
      return;
    }
    const int ndev = (int)devices.size();
    std::vector<sppm_ctx*> ctxs(ndev);
    for (int d = 0; d < ndev; d++) ctxs[d] = sppm_host::create_ctx(devices[d]);
    std::vector<Camera> c(replicas);
    int first = 0;
    if (resume) {
      std::vector<double> saved;
      first = loadCheckpoint(camera.w, camera.h, false, saved, nullptr);
      for (int i1 = 0; i1 < camera.h; i1++)
        for (int jj = 0; jj < camera.w; jj++)
          for (int k = 0; k < 3; k++) camera.pic[i1][jj].val[k] = saved[((size_t)i1 * camera.w + jj) * 3 + k];
    }
    for (int i = first; i < sppmtime; i++) {
      sppm_host::Phase whole("pass (everything)");
      {
        sppm_host::Phase ph("replica cameras");
        for (int j = 0; j < replicas; j++) {
          c[j] = camera;
          c[j].MoveTo(camera.pos + jitter * Vec3::GetUnitRandVec());
          c[j].lookAt(camera.look, 1.);
        }
      }
      // the reference runs its 4 RayTracers on 4 OpenMP threads; here one host thread per GPU
#pragma omp parallel for num_threads(ndev) schedule(static, 1)
      for (int d = 0; d < ndev; d++) {
        for (int j = d; j < replicas; j += ndev) {
          sppm_host::Phase whole_r("replica (everything)");
          RayTracer rt(ctxs[d]);
          rt.params = params;
          rt.fuse_rounds = fuse_rounds;
          rt.camera = &c[j];
          rt.scene = &scene;
          rt.cnt = cnt;
          rt.onetime = onetime;
          rt.pass_base = (uint64_t)i * replicas + j;
          rt.render();
        }
      }
      {
        sppm_host::Phase ph("film accumulate");
        film.resize((size_t)camera.w * camera.h * 3);
        const int hh = camera.h, ww = camera.w, nrep = replicas;
#pragma omp parallel for schedule(static)
        for (int i1 = 0; i1 < hh; i1++) {  // Raytracer.h:449-458, replicas summed in order j = 0, 1, ... per pixel
          for (int j = 0; j < nrep; j++)
            for (int jj = 0; jj < ww; jj++) camera.pic[i1][jj] += c[j].pic[i1][jj];
          std::memcpy(&film[(size_t)i1 * ww * 3], camera.pic[i1], (size_t)ww * 3 * sizeof(double));
        }
      }
      // tone map (Raytracer.h:460-470) on the device, then save (PPM instead of cv::imwrite)
      image.resize(film.size());
      {
        sppm_host::Phase ph("tone map");
        sppm_host::check(sppm_tonemap(ctxs[0], film.data(), camera.w, camera.h, (i + 1) * (double)replicas, image.data()), ctxs[0], "sppm_tonemap");
      }
      // checkpoint first, image second: a crash between the two repeats the image, not the iteration
      if (checkpoint_every > 0 && !checkpoint_path.empty() && (i + 1) % checkpoint_every == 0)
        saveCheckpoint(camera.w, camera.h, false, i + 1, film, 0);
      sppm_host::Phase ph("save image");
      saveImage(ctxs[0], i, camera.w, camera.h);
    }
    for (sppm_ctx* x : ctxs) sppm_ctx_destroy(x);
    sppm_host::print_host_times();
  }

 private:
  void renderShared(Scene& scene, Camera& camera) {
    sppm_ctx* ctx = sppm_host::create_ctx(devices[0]);
    SppmParams prm = params;
    prm.update_mode = SPPM_UPDATE_REFERENCE;  // the per-hit-point update stays the reference's no-op; the pixels are updated
    sppm_host::check(sppm_set_params(ctx, &prm), ctx, "sppm_set_params");
    sppm_host::FlatScene fs;
    scene.flatten(fs);
    SppmSceneDesc d = fs.desc();
    sppm_host::check(sppm_scene_upload(ctx, &d), ctx, "sppm_scene_upload");
    sppm_host::check(sppm_shared_begin(ctx, camera.w, camera.h), ctx, "sppm_shared_begin");
    film.assign((size_t)camera.w * camera.h * 3, 0.0);
    image.resize(film.size());
    const size_t np = (size_t)camera.w * camera.h;
    int first = 0;
    if (resume) {
      std::vector<double> st;
      uint64_t photons = 0;
      first = loadCheckpoint(camera.w, camera.h, true, st, &photons);
      sppm_host::check(sppm_shared_import(ctx, st.data(), st.data() + np, st.data() + 2 * np, photons), ctx, "sppm_shared_import");
    }
    for (int i = first; i < sppmtime; i++) {
      for (int j = 0; j < replicas; j++) {
        Camera c = camera;
        c.MoveTo(camera.pos + jitter * Vec3::GetUnitRandVec());
        c.lookAt(camera.look, 1.);
        SppmCameraDesc cd = c.desc();
        const uint64_t pass_base = (uint64_t)i * replicas + j;
        int64_t n = 0;
        sppm_host::check(sppm_eye_pass(ctx, &cd, pass_base, &n), ctx, "sppm_eye_pass");
        sppm_host::check(sppm_build_index(ctx), ctx, "sppm_build_index");
        sppm_host::check(sppm_shared_load(ctx), ctx, "sppm_shared_load");
        const uint64_t step = sppm_host::rounds_per_call((uint64_t)onetime, fuse_rounds);  // the update is the reference's no-op here
        for (uint64_t k = 0; k < (uint64_t)cnt; k += step) {
          const uint64_t nr = std::min<uint64_t>(step, (uint64_t)cnt - k);
          sppm_host::check(sppm_photon_pass(ctx, k * (uint64_t)onetime, nr * (uint64_t)onetime, pass_base), ctx, "sppm_photon_pass");
          sppm_host::check(sppm_update(ctx), ctx, "sppm_update");
        }
        sppm_host::check(sppm_shared_fold(ctx, (uint64_t)cnt * onetime * scene.plvec.size()), ctx, "sppm_shared_fold");
      }
      sppm_host::check(sppm_shared_resolve(ctx, film.data()), ctx, "sppm_shared_resolve");
      sppm_host::check(sppm_tonemap(ctx, film.data(), camera.w, camera.h, 1.0, image.data()), ctx, "sppm_tonemap");
      if (checkpoint_every > 0 && !checkpoint_path.empty() && (i + 1) % checkpoint_every == 0) {
        std::vector<double> st(5 * np);
        uint64_t photons = 0;
        sppm_host::check(sppm_shared_export(ctx, st.data(), st.data() + np, st.data() + 2 * np, &photons), ctx, "sppm_shared_export");
        saveCheckpoint(camera.w, camera.h, true, i + 1, st, photons);
      }
      saveImage(ctx, i, camera.w, camera.h);
    }
    sppm_ctx_destroy(ctx);
  }
  // checkpoint file: "SPPMCKP1", w, h, shared, iterations done, photons, jitter generator state, doubles
  void saveCheckpoint(int w, int h, bool shared, int done, const std::vector<double>& data, uint64_t photons) const {
    const std::string tmp = checkpoint_path + ".tmp";
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) throw std::runtime_error("checkpoint: cannot write " + tmp);
    const int64_t hdr[5] = {w, h, shared ? 1 : 0, done, (int64_t)data.size()};
    const uint64_t extra[2] = {photons, sppm_host::rng_state()};
    bool ok = std::fwrite("SPPMCKP1", 1, 8, f) == 8 && std::fwrite(hdr, 8, 5, f) == 5 && std::fwrite(extra, 8, 2, f) == 2 &&
              std::fwrite(data.data(), 8, data.size(), f) == data.size();
    ok = std::fclose(f) == 0 && ok;
    if (!ok || std::rename(tmp.c_str(), checkpoint_path.c_str()) != 0) throw std::runtime_error("checkpoint: write failed: " + checkpoint_path);
  }
  int loadCheckpoint(int w, int h, bool shared, std::vector<double>& data, uint64_t* photons) const {
    FILE* f = std::fopen(checkpoint_path.c_str(), "rb");
    if (!f) throw std::runtime_error("checkpoint: cannot open " + checkpoint_path);
    char magic[8];
    int64_t hdr[5];
    uint64_t extra[2];
    bool ok = std::fread(magic, 1, 8, f) == 8 && std::string(magic, 8) == "SPPMCKP1" && std::fread(hdr, 8, 5, f) == 5 &&
              std::fread(extra, 8, 2, f) == 2;
    if (ok && (hdr[0] != w || hdr[1] != h || hdr[2] != (shared ? 1 : 0))) {
      std::fclose(f);
      throw std::runtime_error("checkpoint: film size / estimator differ from this job: " + checkpoint_path);
    }
    // the payload length and the iteration count are the file's word: hold them to what this job can use
    // (film: w*h*3 doubles; shared statistics: R2, N, tau[3] per pixel) before anything is sized from them
    const int64_t want = (int64_t)(shared ? 5 : 3) * w * h;
    if (ok && (hdr[4] != want || hdr[3] < 0 || hdr[3] > sppmtime)) {
      std::fclose(f);
      throw std::runtime_error("checkpoint: payload length or iteration count does not fit this job: " + checkpoint_path);
    }
    if (ok) {
      data.resize((size_t)hdr[4]);
      ok = std::fread(data.data(), 8, data.size(), f) == data.size();
    }
    std::fclose(f);
    if (!ok) throw std::runtime_error("checkpoint: malformed file " + checkpoint_path);
    if (photons) *photons = extra[0];
    sppm_host::rng_state() = extra[1];
    return (int)hdr[3];
  }
  void saveImage(sppm_ctx* ctx, int i, int w, int h) const {
    if (format == "jpg") {
      // the tone-mapped image is stored bottom-up BGR (Raytracer.h:462-469); JPEG takes it as it is
      int64_t n = 0;
      sppm_host::check(sppm_jpeg_encode(ctx, image.data(), h, w, jpeg_quality, nullptr, 0, &n), ctx, "sppm_jpeg_encode");
      std::vector<uchar> out((size_t)n + 4096);
      sppm_host::check(sppm_jpeg_encode(ctx, image.data(), h, w, jpeg_quality, out.data(), (int64_t)out.size(), &n), ctx, "sppm_jpeg_encode");
      const std::string pic = out_prefix + std::to_string(i / 10) + ".jpg";
      std::cerr << "save:" << pic << std::endl;
      FILE* f = std::fopen(pic.c_str(), "wb");
      if (!f) return;
      std::fwrite(out.data(), 1, (size_t)n, f);
      std::fclose(f);
      return;
    }
    const std::string pic = out_prefix + std::to_string(i / 10) + ".ppm";
    std::cerr << "save:" << pic << std::endl;
    savePPM(pic, w, h);
  }
  void savePPM(const std::string& path, int w, int h) const {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return;
    std::fprintf(f, "P6\n%d %d\n255\n", w, h);
    std::vector<uchar> rgb((size_t)w * h * 3);
    for (size_t p = 0; p < (size_t)w * h; p++) {  // film channels are BGR (Q11)
      rgb[3 * p] = image[3 * p + 2];
      rgb[3 * p + 1] = image[3 * p + 1];
      rgb[3 * p + 2] = image[3 * p];
    }
    std::fwrite(rgb.data(), 1, rgb.size(), f);
    std::fclose(f);
  }
};
