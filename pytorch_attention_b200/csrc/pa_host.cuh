// pa_host.cuh — host-side plumbing shared by the C-ABI entry points: error strings, device check,
// TMA tensor-map construction (driver entry point resolved at run time, no -lcuda link dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include "../../include/pa_b200.h"

namespace pa {

inline char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
#define PA_CUDA_OK(expr)                                                                       \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) return pa::fail(PA_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

inline std::atomic<unsigned long long>& launch_counter() {
  static std::atomic<unsigned long long> c{0};
  return c;
}

inline int current_device_check() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(PA_ERR_DEVICE, "no CUDA device (this library has no CPU fallback)");
  static int ok_dev[64];
  static std::once_flag once;
  std::call_once(once, [] { memset(ok_dev, 0, sizeof(ok_dev)); });
  if (dev >= 0 && dev < 64 && ok_dev[dev] == 1) return PA_OK;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return fail(PA_ERR_DEVICE, "cannot query device %d", dev);
  if (major != 10) return fail(PA_ERR_DEVICE, "device %d is compute capability %d.x; sm_100 (B200) required", dev, major);
  if (dev >= 0 && dev < 64) ok_dev[dev] = 1;
  return PA_OK;
}

inline int num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (sms[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms[dev] = n > 0 ? n : 148;
  }
  return sms[dev];
}

inline PFN_cuTensorMapEncodeTiled_v12000 tmap_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// rank-N tiled map over a 16-bit tensor.  dims[0] is the contiguous dimension; strides (bytes) for dims 1..rank-1.
enum TmapSwizzle { TM_SWZ_128 = 0, TM_SWZ_64 = 1, TM_SWZ_32 = 2 };
inline int make_tmap_16b(CUtensorMap* out, int dtype, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, TmapSwizzle swz = TM_SWZ_128) {
  auto fn = tmap_encode_fn();
  if (!fn) return fail(PA_ERR_CUDA, "cuTensorMapEncodeTiled driver entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(PA_ERR_MISALIGNED, "tensor base %p not 16-byte aligned", base);
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      if (strides_bytes[i - 1] % 16 != 0) return fail(PA_ERR_MISALIGNED, "tensor pitch %llu B not a multiple of 16", (unsigned long long)strides_bytes[i - 1]);
    }
    if (box[i] == 0 || box[i] > 256) return fail(PA_ERR_BAD_SHAPE, "TMA box dim %d = %u out of range", i, box[i]);
  }
  // Keyed cache (thread-local: no lock, one host thread per GPU is the supported concurrency): a tensor map is a pure function of
  // these arguments, a forward re-creates the same 8-12 maps every call (same workspace, same weights), and the driver's encoder
  // costs ~0.4 us each.  The key holds every argument, so a recycled address with another shape is simply another entry.
  struct Key {
    const void* base; int dtype, rank, swz; uint64_t dims[5]; uint64_t str[4]; uint32_t box[5];
    bool operator==(const Key& o) const { return memcmp(this, &o, sizeof(Key)) == 0; }
  };
  struct KeyHash {
    size_t operator()(const Key& k) const {
      const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
      uint64_t h = 1469598103934665603ull;
      for (size_t i = 0; i < sizeof(Key) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
      return (size_t)h;
    }
  };
  static_assert(sizeof(Key) % 8 == 0, "Key is hashed word-wise");
  static thread_local std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  Key key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.dtype = dtype; key.rank = rank; key.swz = (int)swz;
  for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; if (i > 0) key.str[i - 1] = strides_bytes[i - 1]; }
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return PA_OK; }
  CUresult r = fn(out, dtype == PA_DTYPE_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                       : dtype == PA_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                  (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz == TM_SWZ_128 ? CU_TENSOR_MAP_SWIZZLE_128B : swz == TM_SWZ_64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
                                     (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
  if (cache.size() >= 4096) cache.clear();
  cache.emplace(key, *out);
  return PA_OK;
}

}  // namespace pa
