// image_kernels.cuh — the two kernels of abi.cu that touch pixels: the tone map of the film and the
// texel expansion done at scene upload.  A header of their own so that tests/device_host can run them on
// the CPU against the oracle; abi.cu is the only product file that includes it.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace {

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

// one 32-bit word per texel (B | G << 8 | R << 16): one load per texel in color_uv (scene.cuh)
__global__ void tex_expand_kernel(const uint8_t* bgr, size_t n, uint32_t* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)bgr[3 * i] | ((uint32_t)bgr[3 * i + 1] << 8) | ((uint32_t)bgr[3 * i + 2] << 16);
}

}  // namespace
