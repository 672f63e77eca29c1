// Decode front-end: JPEG bytes -> BGR uint8 HWC images in the device blob that pf_forward's pre-process reads,
// i.e. the GPU counterpart of `cv2.imread` in front of the path (demo/demo.py:151).  nvJPEG (CUDA toolkit library: library code,
// like cuBLAS) does the entropy decode + IDCT + colour conversion and writes interleaved BGR straight into the blob; this file
// only binds it at run time (dlopen: libpf_b200.so keeps working on hosts without nvJPEG, the entry points then fail loudly) and
// fans the images of a batch out over worker threads, each with its own decoder state and CUDA stream, joined into the caller's
// stream with events.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvjpeg.h>

#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace pf {

struct NvjpegApi {
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
  nvjpegStatus_t (*StateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*StateDestroy)(nvjpegJpegState_t) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*, cudaStream_t) = nullptr;
  const char* error = nullptr;
};

inline const NvjpegApi& nvjpeg_api() {
  static NvjpegApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    for (const char* name : {"libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so.12", "libnvjpeg.so"})
      if ((h = dlopen(name, RTLD_NOW | RTLD_LOCAL))) break;
    if (!h) { api.error = "libnvjpeg.so.12 could not be loaded"; return; }
    auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p && !api.error) api.error = "libnvjpeg lacks a required symbol"; return p; };
    api.CreateSimple = (decltype(api.CreateSimple))sym("nvjpegCreateSimple");
    api.Destroy = (decltype(api.Destroy))sym("nvjpegDestroy");
    api.StateCreate = (decltype(api.StateCreate))sym("nvjpegJpegStateCreate");
    api.StateDestroy = (decltype(api.StateDestroy))sym("nvjpegJpegStateDestroy");
    api.GetImageInfo = (decltype(api.GetImageInfo))sym("nvjpegGetImageInfo");
    api.Decode = (decltype(api.Decode))sym("nvjpegDecode");
  });
  return api;
}

}  // namespace pf

struct pf_jpeg {
  int device = 0;
  nvjpegHandle_t handle = nullptr;
  struct Worker { nvjpegJpegState_t state = nullptr; cudaStream_t stream = nullptr; cudaEvent_t done = nullptr; };
  std::vector<Worker> workers;
  cudaEvent_t start = nullptr;
};
