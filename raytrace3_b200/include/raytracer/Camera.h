// Camera.h — drop-in for raytracer/Camera.h (pinhole camera + film).  Same members, same
// arithmetic; desc() hands the basis to the C ABI.
#pragma once
#include "Element.h"
#include "SppmBridge.h"
#include "Vec3.h"
class Camera {
 public:
  Vec3 pos;
  Vec3 look;
  double len_w = 0, len_h = 0;
  int w = 0, h = 0;
  Vec3 dir;
  Vec3 du, dv;
  Color** pic = nullptr;  // pic[i][j] as upstream (Camera.h:13); the rows are slices of ONE page-locked block, see film()
  Camera() {}
  static const int DEFAULT_LENS = 100;
  static const int DEFAULT_HEIGHT = 1024;
  static const int DEFAULT_WIDTH = 1024;
  Ray emit(int y, int x) {  // Camera.h:18-22
    Vec3 d = du * ((x + 0.5) / w - 0.5) + dv * ((y + 0.5) / h - 0.5) + dir;
    return Ray(pos, d.GetUnit());
  }
  void init(int w_ = DEFAULT_WIDTH, int h_ = DEFAULT_HEIGHT) { h = h_; w = w_; }
  void MoveTo(const Vec3& pos_) { pos = pos_; }
  void lookAt(const Vec3& P, double scale) {  // Camera.h:32-54
    look = P;
    const Vec3 up(0, 0, 1);
    w *= scale;
    h *= scale;
    dir = Vec3(look - pos).GetUnit();
    du = dir.cross(up);
    du.Normalize();
    dv = -dir.cross(du);
    dv.Normalize();
    double fov = 50. * PI / 180.;
    dir *= .5 / tan(fov / 2.);
    alloc();
  }
  Camera(const Camera& c) { copyPose(c); alloc(); }  // Camera.h:55-74: pose copied, fresh zeroed film
  Camera& operator=(const Camera& c) {
    if (pic && (c.w != w || c.h != h)) release();
    copyPose(c);
    alloc();
    return *this;
  }
  void clear() { for (int i = 0; i < h; i++) for (int j = 0; j < w; j++) pic[i][j] = Color(); }
  // The film as the C ABI sees it: double[h][w][3], row-major.  A Color is three doubles, so the block behind
  // pic IS that array: sppm_resolve writes the read-back straight into camera.pic (no staging copy), and the
  // block is page-locked (sppm_host_pin) so that the copy is one DMA transfer.
  double* film() { return pic ? pic[0][0].val : nullptr; }
  const double* film() const { return pic ? pic[0][0].val : nullptr; }
  ~Camera() { release(); }
  SppmCameraDesc desc() const {
    SppmCameraDesc d;
    sppm_host::put3(d.pos, pos.val);
    sppm_host::put3(d.dir, dir.val);
    sppm_host::put3(d.du, du.val);
    sppm_host::put3(d.dv, dv.val);
    d.w = w;
    d.h = h;
    return d;
  }
 private:
  int ah = 0;
  void copyPose(const Camera& c) {
    look = c.look; pos = c.pos; len_w = c.len_w; len_h = c.len_h; w = c.w; h = c.h; dir = c.dir; du = c.du; dv = c.dv;
  }
  Color* block = nullptr;
  bool pinned = false;
  void alloc() {
    static_assert(sizeof(Color) == 3 * sizeof(double), "a Color is three doubles: pic is double[h][w][3]");
    if (!pic && h > 0 && w > 0) {
      block = new Color[(size_t)h * w];
      pic = new Color*[h];
      for (int i = 0; i < h; i++) pic[i] = block + (size_t)i * w;
      ah = h;
      pinned = sppm_host_pin(block, (uint64_t)h * w * sizeof(Color)) == SPPM_OK;  // best effort: pageable works too
    }
  }
  void release() {
    if (!pic) return;
    if (pinned) sppm_host_unpin(block);
    pinned = false;
    delete[] block;
    delete[] pic;
    block = nullptr;
    pic = nullptr;
  }
};
