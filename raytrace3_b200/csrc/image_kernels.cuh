// image_kernels.cuh — the two kernels of abi.cu that touch pixels: the tone map of the film and the
// texel expansion done at scene upload.  A header of their own so that tests/device_host can run them on
// the CPU against the oracle; abi.cu is the only product file that includes it.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "scene.cuh"

namespace {
using sppm::DevTexture;
using sppm::TexLevel;
using sppm::TEX_MAX_LEVELS;
using sppm::tex_tiled_index;
using sppm::tex_tiled_words;

__global__ void tonemap_kernel(const double* film, int w, int h, double div, uint8_t* bgr) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)w * h * 3) return;
  const int k = (int)(t % 3);
  const long long p = t / 3;
  const int j = (int)(p % w), i1 = (int)(p / w);
  const double x = film[t] / div;
  const int v = (int)(pow(1 - exp(-x), 1 / 2.2) * 255 + .5);  // toInt (Raytracer.h:24-26)
  bgr[((size_t)(h - 1 - i1) * w + j) * 3 + k] = (uint8_t)v;    // row flip (Raytracer.h:467)
}

// one 32-bit word per texel (B | G << 8 | R << 16) in 8 x 4 tiles: one load per texel in color_uv (scene.cuh)
__global__ void tex_expand_kernel(const uint8_t* bgr, int rows, int cols, int tiles_x, uint32_t* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  out[tex_tiled_index(r, c, tiles_x)] = (uint32_t)bgr[3 * i] | ((uint32_t)bgr[3 * i + 1] << 8) | ((uint32_t)bgr[3 * i + 2] << 16);
}
// next mip level: every texel the rounded mean of its 2 x 2 block of the level above (an odd last row / column
// repeats the edge texel), per channel
__global__ void tex_mip_kernel(TexLevel src, TexLevel dst) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)dst.rows * dst.cols) return;
  const int r = (int)(i / dst.cols), c = (int)(i % dst.cols);
  const int r0 = min(2 * r, src.rows - 1), r1 = min(2 * r + 1, src.rows - 1), c0 = min(2 * c, src.cols - 1), c1 = min(2 * c + 1, src.cols - 1);
  const uint32_t a = src.texel[tex_tiled_index(r0, c0, src.tiles_x)], b = src.texel[tex_tiled_index(r0, c1, src.tiles_x)],
                 d = src.texel[tex_tiled_index(r1, c0, src.tiles_x)], e = src.texel[tex_tiled_index(r1, c1, src.tiles_x)];
  uint32_t o = 0;
  for (int k = 0; k < 3; k++) {
    const uint32_t sum = ((a >> (8 * k)) & 255u) + ((b >> (8 * k)) & 255u) + ((d >> (8 * k)) & 255u) + ((e >> (8 * k)) & 255u);
    o |= ((sum + 2u) >> 2) << (8 * k);
  }
  const_cast<uint32_t*>(dst.texel)[tex_tiled_index(r, c, dst.tiles_x)] = o;
}
// Layout of a texture's level chain in one buffer of words: returns the total word count and fills the levels
// (texel pointers relative to `base`).
inline size_t tex_layout_levels(int rows, int cols, const uint32_t* base, DevTexture& t) {
  t.rows = rows;
  t.cols = cols;
  t.pad = 0;
  size_t off = 0;
  int n = 0;
  for (int r = rows, c = cols; n < TEX_MAX_LEVELS; n++) {
    TexLevel& L = t.level[n];
    L.rows = r; L.cols = c; L.tiles_x = (c + 7) >> 3; L.pad = 0;
    L.texel = base + off;
    off += tex_tiled_words(r, c);
    if (r == 1 && c == 1) { n++; break; }
    r = r > 1 ? r >> 1 : 1;
    c = c > 1 ? c >> 1 : 1;
  }
  t.n_levels = n;
  for (int k = n; k < TEX_MAX_LEVELS; k++) t.level[k] = t.level[n - 1];
  return off;
}

}  // namespace
