// ref_harness.cpp — C entry points into the reference's OWN code (compiled from
// /root/reference/raytracer/*.h against the shim).  TEST INFRASTRUCTURE: used only to pin the
// oracle (tests/test_ref_pin_cpu.py).  Random numbers are injected (ref_push_uniforms) in the
// order the reference consumes them.
#include "Vec3.h"
#include "Obj.h"
#include "Camera.h"
#include "Element.h"
#include "Light.h"
#include "Bezier.h"
#include "Scene.h"
#include "Raytracer.h"

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

}  // extern "C"
