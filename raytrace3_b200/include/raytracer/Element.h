// Element.h — drop-in for raytracer/Element.h: Ray, Material, Collision, Texture.
#pragma once
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "Vec3.h"
#include "SppmBridge.h"

class Obj;
struct Ray { Vec3 org, dir; Ray() {} Ray(Vec3 o, Vec3 d) : org(o), dir(d) {} };
class Material {  // Element.h:7-19
 public:
  Material(Vec3 d, Vec3 s, Vec3 r, double rr = 1.5, double rl = 1) : diff(d), refl(s), refr(r), refrn(rr), refln(rl) {}
  Vec3 diff, refl, refr;
  double refrn, refln;
  bool IsDiff() { return !IsZero(diff); }
  bool IsRefl() { return !IsZero(refl); }
  bool IsRefr() { return !IsZero(refr); }
};
class Collision {  // Element.h:20-39
 public:
  bool IsValid() { return obj != nullptr; }
  bool IsFront() { return !inside; }
  Vec3 Pos, N, From;
  Obj* obj;
  Color color = Vec3();
  double dist;
  explicit Collision(const Vec3& pos = Vec3(), const Vec3& from = Vec3(), double d = 0, Obj* o = nullptr,
                     const Vec3& n = Vec3(), const Vec3& c = Vec3()) : Pos(pos), N(n), From(from) {
    dist = d;
    obj = o;
    color = c;
  }
  bool inside = 0;
};
const Collision NoneColl = Collision();

// Element.h:43-75.  The reference decodes a JPEG with cv::imread; here `Texture("wall.jpg")`
// decodes it on the GPU with nvJPEG (sppm_jpeg_decode; pixel values within the
// decoders' rounding of OpenCV's), and a texture may also be raw BGR8 pixels: handed over
// directly or read from a "P6"-style file written by tools/decode_texture.py
// ("BGR8 <rows> <cols>\n" + rows*cols*3 bytes) when bit-identical texels are wanted.  Bilinear
// lookups run on the device (Texture::colorUV -> csrc/scene.cuh color_uv).
struct Texture {
  int rows = 0, cols = 0;
  std::vector<uint8_t> bgr;
  Texture(int r, int c, const uint8_t* data) : rows(r), cols(c), bgr(data, data + (size_t)r * c * 3) {}
  explicit Texture(const std::string& fileName) {
    FILE* f = std::fopen(fileName.c_str(), "rb");
    if (!f) throw std::runtime_error("Texture: cannot open " + fileName + " (the reference would index out of range, Q18)");
    const int b0 = std::fgetc(f), b1 = std::fgetc(f);
    if (b0 == 0xFF && b1 == 0xD8) {  // JPEG: decode on the device
      std::fseek(f, 0, SEEK_END);
      const long len = std::ftell(f);
      std::fseek(f, 0, SEEK_SET);
      std::vector<uint8_t> data((size_t)len);
      const size_t got = std::fread(data.data(), 1, data.size(), f);
      std::fclose(f);
      if (got != data.size()) throw std::runtime_error("Texture: short read in " + fileName);
      sppm_ctx* c = sppm_host::compat_ctx();
      int32_t r = 0, cc = 0;
      sppm_host::check(sppm_jpeg_info(c, data.data(), (int64_t)data.size(), &r, &cc), c, "sppm_jpeg_info");
      rows = r;
      cols = cc;
      bgr.resize((size_t)rows * cols * 3);
      sppm_host::check(sppm_jpeg_decode(c, data.data(), (int64_t)data.size(), rows, cols, bgr.data()), c, "sppm_jpeg_decode");
      std::cout << "Texture:" << fileName << "::row:" << rows << " cols:" << cols << std::endl;
      return;
    }
    std::fseek(f, 0, SEEK_SET);
    if (std::fscanf(f, "BGR8 %d %d", &rows, &cols) != 2 || rows <= 0 || cols <= 0) { std::fclose(f); throw std::runtime_error("Texture: bad header in " + fileName); }
    std::fgetc(f);
    bgr.resize((size_t)rows * cols * 3);
    size_t got = std::fread(bgr.data(), 1, bgr.size(), f);
    std::fclose(f);
    if (got != bgr.size()) throw std::runtime_error("Texture: short read in " + fileName);
    std::cout << "Texture:" << fileName << "::row:" << rows << " cols:" << cols << std::endl;
  }
};
