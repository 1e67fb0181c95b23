// imageio.cu — host image I/O of the render path on the GPU.
//
// The reference decodes its textures with cv::imread (Element.h:47-59) and saves the tone-mapped
// film with cv::imwrite every iteration (Raytracer.h:460-474).  Here both go through nvJPEG,
// which ships with the CUDA toolkit: baseline JPEG -> interleaved BGR8 (the layout Texture and
// the film use, Q11) decoded on the device, and BGR8 -> JPEG encoded on the device.  The library
// is found with dlopen so that libsppm_b200.so itself does not depend on it; without it the two
// entry points fail with SPPM_ERR_INVALID and a message (textures can still arrive pre-decoded).
#include <dlfcn.h>
#include <nvjpeg.h>

#include <mutex>

#include "ctx.h"

namespace {

struct NvJpegApi {
  bool ok = false;
  std::string why;
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
  nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*,
                           cudaStream_t) = nullptr;
  nvjpegStatus_t (*EncoderStateCreate)(nvjpegHandle_t, nvjpegEncoderState_t*, cudaStream_t) = nullptr;
  nvjpegStatus_t (*EncoderStateDestroy)(nvjpegEncoderState_t) = nullptr;
  nvjpegStatus_t (*EncoderParamsCreate)(nvjpegHandle_t, nvjpegEncoderParams_t*, cudaStream_t) = nullptr;
  nvjpegStatus_t (*EncoderParamsDestroy)(nvjpegEncoderParams_t) = nullptr;
  nvjpegStatus_t (*EncoderParamsSetQuality)(nvjpegEncoderParams_t, const int, cudaStream_t) = nullptr;
  nvjpegStatus_t (*EncoderParamsSetSamplingFactors)(nvjpegEncoderParams_t, const nvjpegChromaSubsampling_t, cudaStream_t) = nullptr;
  nvjpegStatus_t (*EncodeImage)(nvjpegHandle_t, nvjpegEncoderState_t, const nvjpegEncoderParams_t, const nvjpegImage_t*,
                                nvjpegInputFormat_t, int, int, cudaStream_t) = nullptr;
  nvjpegStatus_t (*EncodeRetrieveBitstream)(nvjpegHandle_t, nvjpegEncoderState_t, unsigned char*, size_t*, cudaStream_t) = nullptr;
};

NvJpegApi& nvjpeg_api() {
  static NvJpegApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) {
      api.why = "libnvjpeg.so.12 not found (dlopen)";
      return;
    }
    bool all = true;
    auto sym = [&](const char* n) {
      void* p = dlsym(h, n);
      if (!p) {
        all = false;
        api.why = std::string("nvJPEG symbol missing: ") + n;
      }
      return p;
    };
#define SPPM_NVJ(f) api.f = reinterpret_cast<decltype(api.f)>(sym("nvjpeg" #f))
    SPPM_NVJ(CreateSimple); SPPM_NVJ(Destroy); SPPM_NVJ(JpegStateCreate); SPPM_NVJ(JpegStateDestroy); SPPM_NVJ(GetImageInfo);
    SPPM_NVJ(Decode); SPPM_NVJ(EncoderStateCreate); SPPM_NVJ(EncoderStateDestroy); SPPM_NVJ(EncoderParamsCreate);
    SPPM_NVJ(EncoderParamsDestroy); SPPM_NVJ(EncoderParamsSetQuality); SPPM_NVJ(EncoderParamsSetSamplingFactors);
    SPPM_NVJ(EncodeImage); SPPM_NVJ(EncodeRetrieveBitstream);
#undef SPPM_NVJ
    api.ok = all;
  });
  return api;
}

int img_fail(sppm_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int d) {
    cudaGetDevice(&prev);
    cudaSetDevice(d);
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

}  // namespace

extern "C" {

int sppm_jpeg_info(sppm_ctx* c, const uint8_t* jpeg, int64_t n_bytes, int32_t* rows, int32_t* cols) {
  if (!c || !jpeg || n_bytes <= 0 || !rows || !cols) return SPPM_ERR_INVALID;
  NvJpegApi& nj = nvjpeg_api();
  if (!nj.ok) return img_fail(c, SPPM_ERR_INVALID, "nvJPEG unavailable: " + nj.why);
  DeviceGuard g(c->device);
  nvjpegHandle_t h = nullptr;
  if (nj.CreateSimple(&h) != NVJPEG_STATUS_SUCCESS) return img_fail(c, SPPM_ERR_CUDA, "nvjpegCreateSimple failed");
  int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
  nvjpegChromaSubsampling_t ss;
  const nvjpegStatus_t st = nj.GetImageInfo(h, jpeg, (size_t)n_bytes, &ncomp, &ss, ws, hs);
  nj.Destroy(h);
  if (st != NVJPEG_STATUS_SUCCESS) return img_fail(c, SPPM_ERR_INVALID, "not a JPEG stream nvJPEG can parse (status " + std::to_string((int)st) + ")");
  *rows = hs[0];
  *cols = ws[0];
  return SPPM_OK;
}

int sppm_jpeg_decode(sppm_ctx* c, const uint8_t* jpeg, int64_t n_bytes, int32_t rows, int32_t cols, uint8_t* bgr8) {
  if (!c || !jpeg || n_bytes <= 0 || rows <= 0 || cols <= 0 || !bgr8) return SPPM_ERR_INVALID;
  NvJpegApi& nj = nvjpeg_api();
  if (!nj.ok) return img_fail(c, SPPM_ERR_INVALID, "nvJPEG unavailable: " + nj.why);
  DeviceGuard g(c->device);
  nvjpegHandle_t h = nullptr;
  nvjpegJpegState_t st = nullptr;
  int rc = SPPM_OK;
  const size_t bytes = (size_t)rows * cols * 3;
  if (sppm::dev_reserve(c->scratch, bytes) != cudaSuccess) return img_fail(c, SPPM_ERR_CUDA, "cudaMalloc failed (JPEG decode buffer)");
  if (nj.CreateSimple(&h) != NVJPEG_STATUS_SUCCESS || nj.JpegStateCreate(h, &st) != NVJPEG_STATUS_SUCCESS) {
    rc = img_fail(c, SPPM_ERR_CUDA, "nvjpegCreateSimple / nvjpegJpegStateCreate failed");
  } else {
    int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t ss;
    nvjpegStatus_t s = nj.GetImageInfo(h, jpeg, (size_t)n_bytes, &ncomp, &ss, ws, hs);
    if (s != NVJPEG_STATUS_SUCCESS || hs[0] != rows || ws[0] != cols) {
      rc = img_fail(c, SPPM_ERR_INVALID, "JPEG size does not match rows x cols (call sppm_jpeg_info first)");
    } else {
      nvjpegImage_t img{};
      img.channel[0] = (unsigned char*)c->scratch.p;  // interleaved B,G,R: cv::imread's layout (Element.h:52-56)
      img.pitch[0] = (size_t)cols * 3;
      s = nj.Decode(h, st, jpeg, (size_t)n_bytes, NVJPEG_OUTPUT_BGRI, &img, c->stream);
      if (s != NVJPEG_STATUS_SUCCESS) rc = img_fail(c, SPPM_ERR_CUDA, "nvjpegDecode failed (status " + std::to_string((int)s) + ")");
      else if (cudaMemcpyAsync(bgr8, c->scratch.p, bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
               cudaStreamSynchronize(c->stream) != cudaSuccess)
        rc = img_fail(c, SPPM_ERR_CUDA, "copy of the decoded image failed");
    }
  }
  if (st) nj.JpegStateDestroy(st);
  if (h) nj.Destroy(h);
  return rc;
}

int sppm_jpeg_encode(sppm_ctx* c, const uint8_t* bgr8, int32_t rows, int32_t cols, int32_t quality, uint8_t* out, int64_t cap,
                     int64_t* n_bytes) {
  if (!c || !bgr8 || rows <= 0 || cols <= 0 || !n_bytes || quality < 1 || quality > 100) return SPPM_ERR_INVALID;
  NvJpegApi& nj = nvjpeg_api();
  if (!nj.ok) return img_fail(c, SPPM_ERR_INVALID, "nvJPEG unavailable: " + nj.why);
  DeviceGuard g(c->device);
  nvjpegHandle_t h = nullptr;
  nvjpegEncoderState_t es = nullptr;
  nvjpegEncoderParams_t ep = nullptr;
  int rc = SPPM_OK;
  const size_t bytes = (size_t)rows * cols * 3;
  if (sppm::dev_reserve(c->scratch, bytes) != cudaSuccess) return img_fail(c, SPPM_ERR_CUDA, "cudaMalloc failed (JPEG encode buffer)");
  if (nj.CreateSimple(&h) != NVJPEG_STATUS_SUCCESS || nj.EncoderStateCreate(h, &es, c->stream) != NVJPEG_STATUS_SUCCESS ||
      nj.EncoderParamsCreate(h, &ep, c->stream) != NVJPEG_STATUS_SUCCESS) {
    rc = img_fail(c, SPPM_ERR_CUDA, "nvJPEG encoder set-up failed");
  } else {
    nj.EncoderParamsSetQuality(ep, quality, c->stream);
    nj.EncoderParamsSetSamplingFactors(ep, NVJPEG_CSS_444, c->stream);
    nvjpegImage_t img{};
    img.channel[0] = (unsigned char*)c->scratch.p;
    img.pitch[0] = (size_t)cols * 3;
    size_t len = 0;
    if (cudaMemcpyAsync(c->scratch.p, bgr8, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
      rc = img_fail(c, SPPM_ERR_CUDA, "upload of the image failed");
    else if (nj.EncodeImage(h, es, ep, &img, NVJPEG_INPUT_BGRI, cols, rows, c->stream) != NVJPEG_STATUS_SUCCESS)
      rc = img_fail(c, SPPM_ERR_CUDA, "nvjpegEncodeImage failed");
    else if (nj.EncodeRetrieveBitstream(h, es, nullptr, &len, c->stream) != NVJPEG_STATUS_SUCCESS)
      rc = img_fail(c, SPPM_ERR_CUDA, "nvjpegEncodeRetrieveBitstream (size) failed");
    else {
      *n_bytes = (int64_t)len;
      if (out) {
        if ((int64_t)len > cap) rc = img_fail(c, SPPM_ERR_CAPACITY, "JPEG output buffer too small");
        else if (cudaStreamSynchronize(c->stream) != cudaSuccess ||
                 nj.EncodeRetrieveBitstream(h, es, out, &len, c->stream) != NVJPEG_STATUS_SUCCESS ||
                 cudaStreamSynchronize(c->stream) != cudaSuccess)
          rc = img_fail(c, SPPM_ERR_CUDA, "nvjpegEncodeRetrieveBitstream failed");
      }
    }
  }
  if (ep) nj.EncoderParamsDestroy(ep);
  if (es) nj.EncoderStateDestroy(es);
  if (h) nj.Destroy(h);
  return rc;
}

}  // extern "C"
