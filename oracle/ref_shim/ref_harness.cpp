// ref_harness.cpp — C entry points into the reference's OWN code (compiled from
// /root/reference/raytracer/*.h against the shim).  TEST INFRASTRUCTURE: used only to pin the
// oracle (tests/test_ref_pin_cpu.py) and, in ref_time_photon_map, to time the reference's own
// RayTracer::PhotonMap on the host cores (bench.py, cpu_baseline.reference_code).  Random numbers
// are injected (ref_push_uniforms) in the order the reference consumes them; the timing entry
// point lets the generator run freely.
#include "Vec3.h"
#include "Obj.h"
#include "Camera.h"
#include "Element.h"
#include "Light.h"
#include "Bezier.h"
#include "Scene.h"
#include "Raytracer.h"

#include <fcntl.h>
#include <omp.h>
#include <unistd.h>

#include <chrono>
#include <fstream>

#include "sppm_b200.h"

static Vec3 V(const double* p) { return Vec3(p[0], p[1], p[2]); }
static void put(double* d, const cv::Vec3d& v) { d[0] = v.val[0]; d[1] = v.val[1]; d[2] = v.val[2]; }

struct RefScene {
  Scene scene;
  std::vector<std::string> names;
};

extern "C" {

void ref_push_uniforms(const double* u, int n) { for (int i = 0; i < n; i++) ref_shim::uniforms().push_back(u[i]); }
void ref_clear_uniforms() { ref_shim::uniforms().clear(); ref_shim::underflows() = 0; }
long ref_underflows() { return ref_shim::underflows(); }
long ref_uniforms_left() { return (long)ref_shim::uniforms().size(); }

// which of the two GetRandu calls in `Vec3 x(0, GetRandu(0.,1.), GetRandu(0.,1.))` (Bezier.h:117)
// this compiler evaluates first: returns 1 if u = x[1] takes the first draw (Q16)
int ref_arg_order_u_first() {
  ref_clear_uniforms();
  double s[2] = {0.25, 0.75};
  ref_push_uniforms(s, 2);
  Vec3 x(0, GetRandu(0., 1.), GetRandu(0., 1.));
  return x.val[1] == 0.25 ? 1 : 0;
}

int ref_aabb_intersect(const double box[6], const double org[3], const double dir[3]) {
  AABBbox b;
  b.pmin = V(box);
  b.pmax = V(box + 3);
  return b.intersect(V(org), V(dir)) ? 1 : 0;
}
void ref_bezier_eval(const double P[48], double u, double v, double S[3], double Su[3], double Sv[3]) {
  Vec3 p[16];
  for (int i = 0; i < 16; i++) p[i] = V(P + 3 * i);
  put(S, Bezier3::evalPatchPoint(p, u, v));
  put(Su, Bezier3::evalPatchDerivU(p, u, v));
  put(Sv, Bezier3::evalPatchDerivV(p, u, v));
}
// Bezier3::newton (Bezier.h:112-159); consumes 2 * MAX_ITER * 5 injected uniforms
void ref_newton(const double P[48], const double org[3], const double dir[3], double tuv[3]) {
  Vec3 p[16];
  for (int i = 0; i < 16; i++) p[i] = V(P + 3 * i);
  Bezier3 b;
  b.read(p);
  put(tuv, b.newton(V(org), V(dir), MAX_ITER));
}
void ref_unit_rand_refl(const double n[3], double out[3]) { put(out, Vec3::GetUnitRandRefl(V(n))); }
void ref_unit_rand_vec(double out[3]) { put(out, Vec3::GetUnitRandVec()); }
void ref_refracted(const double I[3], const double N[3], double n, double out[3]) {
  Vec3 r = V(I).refracted(V(N), n);
  r.Normalize();
  put(out, r);
}
void ref_camera(const double pos[3], const double look[3], int w, int h, SppmCameraDesc* out, int y, int x, double ray6[6]) {
  Camera c;
  c.init(w, h);
  c.MoveTo(V(pos));
  c.lookAt(V(look), 1.);
  put(out->pos, c.pos); put(out->dir, c.dir); put(out->du, c.du); put(out->dv, c.dv);
  out->w = c.w; out->h = c.h;
  Ray r = c.emit(y, x);
  put(ray6, r.org); put(ray6 + 3, r.dir);
}
int ref_to_int(double x) { return toInt(x); }

void ref_register_image(const char* name, int rows, int cols, const unsigned char* bgr) {
  ref_shim::Image im;
  im.rows = rows; im.cols = cols; im.bgr = bgr;
  ref_shim::images()[name] = im;
}
void ref_color_uv(const char* name, double u, double v, double out[3]) {
  Texture t(name);
  put(out, t.colorUV(u, v));
}

// Build a Scene of the reference's own PlaneObj / SphereObj / Bezier3Obj / SpotLight from the
// flattened description.  Bezier objects go through the real .bpt loader (Bezier.h:199-239):
// the world-space patches are written to tmpdir/<i>.bpt and loaded with C = 0, scale = 1,
// Trans = eye, which reproduces them exactly.
void* ref_scene_create(const SppmSceneDesc* d, const char* tmpdir) {
  RefScene* rs = new RefScene();
  std::vector<Texture*> tex;
  for (int t = 0; t < d->n_textures; t++) {
    std::string name = std::string(tmpdir) + "/tex" + std::to_string(t);
    ref_register_image(name.c_str(), d->textures[t].rows, d->textures[t].cols, d->textures[t].bgr8);
    tex.push_back(new Texture(name));
  }
  for (int i = 0; i < d->n_objects; i++) {
    const SppmObject& o = d->objects[i];
    Material m(V(o.mat.diff), V(o.mat.refl), V(o.mat.refr), o.mat.refrn);
    Obj* ob = nullptr;
    if (o.kind == SPPM_OBJ_PLANE) {
      PlaneObj* p = new PlaneObj(V(o.p0), V(o.norm), V(o.color), m);
      p->texU = V(o.texU); p->texV = V(o.texV);
      ob = p;
    } else if (o.kind == SPPM_OBJ_SPHERE) {
      SphereObj* s = new SphereObj(o.radius, V(o.p0), V(o.color), m);
      s->texU = V(o.texU); s->texV = V(o.texV);
      ob = s;
    } else {
      std::string file = std::string(tmpdir) + "/" + std::to_string(i) + ".bpt";
      {
        std::ofstream f(file);
        f.precision(17);
        f << o.n_patch << "\n";
        for (int p = o.first_patch; p < o.first_patch + o.n_patch; p++) {
          f << "3 3\n";
          for (int q = 0; q < 16; q++)
            f << d->patches[(p * 16 + q) * 3] << " " << d->patches[(p * 16 + q) * 3 + 1] << " " << d->patches[(p * 16 + q) * 3 + 2] << "\n";
        }
      }
      ob = new Bezier3Obj(Vec3(0, 0, 0), 1, file, V(o.color), m);
      std::cin.clear();
    }
    if (o.tex_id >= 0) ob->setTexture(tex[o.tex_id]);
    rs->scene.objvec.push_back(ob);
  }
  for (int l = 0; l < d->n_lights; l++) rs->scene.plvec.push_back(new SpotLight(V(d->lights[l].pos), V(d->lights[l].color)));
  return rs;
}
void ref_scene_destroy(void* h) { delete (RefScene*)h; }

static void fill(const Scene& sc, Collision& c, SppmHit* h) {
  h->t = c.IsValid() ? c.dist : MAX_NUM;
  put(h->pos, c.Pos); put(h->n, c.N); put(h->color, c.color);
  h->u = h->v = 0;
  h->obj = -1;
  for (size_t i = 0; i < sc.objvec.size(); i++) if (sc.objvec[i] == c.obj) h->obj = (int)i;
  h->patch = -1;
  h->inside = c.inside ? 1 : 0;
  h->pad = 0;
}
// Scene::NearCollide (Scene.h:165-182) for one ray
void ref_near_collide(void* h, const double ray6[6], SppmHit* out) {
  RefScene* rs = (RefScene*)h;
  Collision c = rs->scene.NearCollide(Ray(V(ray6), V(ray6 + 3)));
  fill(rs->scene, c, out);
}

static void load_hvec(RayTracer& rt, int64_t n, const double* pos3, const double* n3, const double* wgt3, const double* r2,
                      const int32_t* newN, const int32_t* N, const double* tao3, std::vector<double>& rdata) {
  for (int64_t i = 0; i < n; i++) {
    HitPoint* hp = new HitPoint(V(pos3 + 3 * i), V(n3 + 3 * i), Vec3(), nullptr, 0, 0, V(wgt3 + 3 * i));
    hp->R2 = r2[i];
    if (newN) hp->newN = newN[i];
    if (N) hp->N = N[i];
    if (tao3) hp->tao = V(tao3 + 3 * i);
    rt.Hvec.push_back(hp);
  }
  rdata.resize(3 * n + 3);
  for (int64_t i = 0; i < n; i++) for (int j = 0; j < 3; j++) rdata[i * 3 + j] = rt.Hvec[i]->pos[j];  // Raytracer.h:370-377
  flann::Matrix<double> dataset(rdata.data(), (size_t)n, 3);
  rt.index.buildIndex(dataset);                                                                       // Raytracer.h:378-380
}
// One photon of PhotonMap's inner loop (Raytracer.h:230-231): pl->emit(ray, flux); PhotonTrace(ray, 0, flux)
void ref_photon(void* h, int light, int64_t n_hp, const double* pos3, const double* n3, const double* wgt3, const double* r2,
                int32_t* newN, double* tao3) {
  RefScene* rs = (RefScene*)h;
  RayTracer rt;
  rt.scene = &rs->scene;
  rt.camera = nullptr;
  std::vector<double> rdata;
  load_hvec(rt, n_hp, pos3, n3, wgt3, r2, newN, nullptr, tao3, rdata);
  Ray ray;
  Color flux;
  rs->scene.plvec[light]->emit(ray, flux);
  rt.PhotonTrace(ray, 0, flux);
  for (int64_t i = 0; i < n_hp; i++) {
    newN[i] = rt.Hvec[i]->newN;
    put(tao3 + 3 * i, rt.Hvec[i]->tao);
  }
}
// ViewTrace for pixel (row i, col j) (Raytracer.h:111); returns the number of hit points written
int64_t ref_view_trace(void* h, const double cam_pos[3], const double cam_look[3], int w, int hgt, int i, int j, int64_t cap,
                       double* pos3, double* n3, double* wgt3, double* r2, int32_t* pix) {
  RefScene* rs = (RefScene*)h;
  Camera c;
  c.init(w, hgt);
  c.MoveTo(V(cam_pos));
  c.lookAt(V(cam_look), 1.);
  RayTracer rt;
  rt.scene = &rs->scene;
  rt.camera = &c;
  rt.ViewTrace(c.emit(i, j), 0, Color(1, 1, 1), i, j);
  int64_t n = (int64_t)rt.Hvec.size();
  for (int64_t k = 0; k < n && k < cap; k++) {
    put(pos3 + 3 * k, rt.Hvec[k]->pos); put(n3 + 3 * k, rt.Hvec[k]->n); put(wgt3 + 3 * k, rt.Hvec[k]->wgt);
    r2[k] = rt.Hvec[k]->R2;
    pix[2 * k] = rt.Hvec[k]->x; pix[2 * k + 1] = rt.Hvec[k]->y;
  }
  rt.camera = nullptr;
  return n;
}
// HitPoint::update (Raytracer.h:69-79) on arrays
void ref_update(int64_t n, int32_t* newN, int32_t* N, double* r2, double* tao3) {
  for (int64_t i = 0; i < n; i++) {
    HitPoint hp;
    hp.newN = newN[i]; hp.N = N[i]; hp.R2 = r2[i]; hp.tao = V(tao3 + 3 * i);
    hp.update();
    newN[i] = hp.newN; N[i] = hp.N; r2[i] = hp.R2; put(tao3 + 3 * i, hp.tao);
  }
}
// density estimate of one hit point (Raytracer.h:292)
void ref_resolve_one(const double tao[3], double R2, int cnt, int onetime, double out[3]) {
  HitPoint hp;
  hp.tao = V(tao); hp.R2 = R2;
  HitPoint* p = &hp;
  put(out, p->tao * (1.0 / PI / (p->R2 * cnt * onetime)));
}

// kd-tree of the FLANN shim against its exhaustive scan: n random points in [0,side)^3 (many of them
// on shared planes, like hit points on walls), nq queries; returns the number of queries whose
// (index, distance) lists differ — 0 expected.  total_found reports how many neighbours were compared.
long ref_flann_selfcheck(int n, int nq, double side, double radius2, unsigned seed, long* total_found) {
  std::vector<double> data((size_t)3 * n + 3), q(3);
  unsigned long long st = seed * 2654435761ULL + 12345;
  auto rnd = [&]() { st = st * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(st >> 11) * (1.0 / 9007199254740992.0); };
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < 3; k++) data[3 * i + k] = rnd() * side;
    int flat = (int)(rnd() * 6);                      // a third of the points lie exactly on a coordinate plane
    if (flat < 3 && (i & 1)) data[3 * i + flat] = 0.0;
    if (i % 97 == 0 && i) for (int k = 0; k < 3; k++) data[3 * i + k] = data[3 * (i - 1) + k];  // duplicates
  }
  flann::Index<flann::L2<double> > index(flann::Matrix<double>(data.data(), 1, 3), flann::KDTreeSingleIndexParams());
  index.buildIndex(flann::Matrix<double>(data.data(), (size_t)n, 3));
  std::vector<std::vector<int> > i1, i2;
  std::vector<std::vector<double> > d1, d2;
  long bad = 0, found = 0;
  for (int j = 0; j < nq; j++) {
    if (j & 1) for (int k = 0; k < 3; k++) q[k] = rnd() * side;
    else for (int k = 0; k < 3; k++) q[k] = data[3 * (size_t)(rnd() * n) + k] + (j % 4 == 0 ? 0.0 : (rnd() - 0.5) * 1e-3);
    flann::Matrix<double> qm(q.data(), 1, 3);
    found += index.radiusSearch(qm, i1, d1, (float)radius2, flann::SearchParams());
    index.radiusSearch(qm, i2, d2, (float)radius2, flann::SearchParams(), true);
    if (i1 != i2 || d1 != d2) bad++;
  }
  if (total_found) *total_found = found;
  return bad;
}

// The reference's own photon pass, timed.  `threads` independent replicas, one RayTracer + Camera
// per thread over the shared Scene, exactly the structure of SPPMRayTracer::render
// (Raytracer.h:429-447) minus the 0.00015 camera jitter: each runs GetHitPoint (Raytracer.h:102),
// flattens its hit points and builds the index as render() does (:369-380) — untimed set-up —
// and then, after a barrier, PhotonMap(cnt) (:210-295): cnt rounds of onetime = 100 photons per
// light (hard-coded, :219), each followed by the update sweep over ALL hit points (:238-245), and
// the final film resolve.  With target_s > 0 the round count is chosen from a 10-round probe so
// that the timed call lasts about target_s seconds (cnt is then only the upper limit).  out =
// {set-up seconds, PhotonMap seconds (wall, all threads), hit points per replica, photons traced
// by all threads in the timed call, rounds}.  The reference prints a progress line per
// row and per round; stdout/stderr are parked on /dev/null while it runs.  Known quirk kept: the
// query buffer `qdata` (Raytracer.h:43) is one global shared by all threads (a race in the
// reference, part of Q15), so a timing run is not a parity run.
int ref_time_photon_map(void* h, const double cam_pos[3], const double cam_look[3], int w, int hgt, int cnt, int threads,
                        double target_s, double out[5]) {
  RefScene* rs = (RefScene*)h;
  if (threads < 1) threads = 1;
  fflush(stdout); fflush(stderr); std::cout.flush(); std::cerr.flush();
  int null_fd = open("/dev/null", O_WRONLY), save1 = dup(1), save2 = dup(2);
  if (null_fd >= 0) { dup2(null_fd, 1); dup2(null_fd, 2); }
  double t_setup = 0, t_map = 0;
  long n_hp = 0;
  std::chrono::steady_clock::time_point t0, tp, t1, t2;
  t0 = std::chrono::steady_clock::now();
#pragma omp parallel num_threads(threads)
  {
    ref_shim::free_state() = 0x9E3779B97F4A7C15ULL ^ ((unsigned long long)(omp_get_thread_num() + 1) * 0xD1B54A32D192ED03ULL);
    Camera c;
    c.init(w, hgt);
    c.MoveTo(V(cam_pos));
    c.lookAt(V(cam_look), 1.);
    {
      RayTracer rt;
      rt.camera = &c;
      rt.scene = &rs->scene;
      rt.GetHitPoint();
      std::vector<double> rdata(3 * rt.Hvec.size() + 3);
      for (size_t i = 0; i < rt.Hvec.size(); i++) for (int j = 0; j < 3; j++) rdata[i * 3 + j] = rt.Hvec[i]->pos[j];
      flann::Matrix<double> dataset(rdata.data(), rt.Hvec.size(), 3);
      rt.index.buildIndex(dataset);
      if (target_s > 0) {
#pragma omp barrier
#pragma omp master
        tp = std::chrono::steady_clock::now();
        rt.PhotonMap(10);
#pragma omp barrier
#pragma omp master
        {
          double per_round = std::chrono::duration<double>(std::chrono::steady_clock::now() - tp).count() / 10.0;
          long want = (long)(target_s / (per_round > 1e-9 ? per_round : 1e-9));
          cnt = (int)std::max(10L, std::min((long)cnt, want));
        }
      }
#pragma omp barrier
#pragma omp master
      { t1 = std::chrono::steady_clock::now(); n_hp = (long)rt.Hvec.size(); }
      rt.PhotonMap(cnt);
#pragma omp barrier
#pragma omp master
      t2 = std::chrono::steady_clock::now();
    }
    ref_shim::free_state() = 0;
  }
  t_setup = std::chrono::duration<double>(t1 - t0).count();
  t_map = std::chrono::duration<double>(t2 - t1).count();
  fflush(stdout); fflush(stderr); std::cout.flush(); std::cerr.flush();
  if (null_fd >= 0) { dup2(save1, 1); dup2(save2, 2); close(null_fd); }
  if (save1 >= 0) close(save1);
  if (save2 >= 0) close(save2);
  out[0] = t_setup;
  out[1] = t_map;
  out[2] = (double)n_hp;
  out[3] = (double)threads * cnt * 100.0 * (double)rs->scene.plvec.size();
  out[4] = (double)cnt;
  return 0;
}

}  // extern "C"
