// Element.h — drop-in for raytracer/Element.h: Ray, Material, Collision, Texture.
#pragma once
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "Vec3.h"

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

// Element.h:43-75.  The reference decodes a JPEG with cv::imread; here a texture is raw BGR8
// pixels: either handed over directly or read from a "P6"-style file written by
// tools/decode_texture.py ("BGR8 <rows> <cols>\n" + rows*cols*3 bytes).  Bilinear lookups run
// on the device (Texture::colorUV -> csrc/scene.cuh color_uv).
struct Texture {
  int rows = 0, cols = 0;
  std::vector<uint8_t> bgr;
  Texture(int r, int c, const uint8_t* data) : rows(r), cols(c), bgr(data, data + (size_t)r * c * 3) {}
  explicit Texture(const std::string& fileName) {
    FILE* f = std::fopen(fileName.c_str(), "rb");
    if (!f) throw std::runtime_error("Texture: cannot open " + fileName + " (the reference would index out of range, Q18)");
    if (std::fscanf(f, "BGR8 %d %d", &rows, &cols) != 2 || rows <= 0 || cols <= 0) { std::fclose(f); throw std::runtime_error("Texture: bad header in " + fileName); }
    std::fgetc(f);
    bgr.resize((size_t)rows * cols * 3);
    size_t got = std::fread(bgr.data(), 1, bgr.size(), f);
    std::fclose(f);
    if (got != bgr.size()) throw std::runtime_error("Texture: short read in " + fileName);
    std::cout << "Texture:" << fileName << "::row:" << rows << " cols:" << cols << std::endl;
  }
};
