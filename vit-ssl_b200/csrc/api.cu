// vitssl_b200 — runtime glue of the C-ABI: error slot, launch accounting, device check and
// host-side TMA descriptor encoding (driver entry point resolved at run time, so the library
// links against cudart only).
#include <cudaTypedefs.h>

#include <atomic>
#include <mutex>
#include <vector>
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include "common.cuh"
#include "vitssl_b200.h"

namespace vitssl {

static thread_local char g_error[512] = "";
// process-wide: backward kernels are launched from autograd's worker thread
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

long long launches_now() { return g_launches.load(); }
void add_launches(long long n) { g_launches += n; }

int check_launch(const char* what) {
  ++g_launches;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(err));
    return VITSSL_ERR_CUDA;
  }
  return VITSSL_OK;
}

// ---- per-launch profile (bench.py's roofline numbers for the C-sequenced encoder stack) ----------
struct ProfRec { char kind[40]; double work; cudaEvent_t e0, e1; };
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;  // forward runs on the main thread, backward on autograd's worker
static std::vector<ProfRec> g_prof;

bool prof_on() { return g_prof_on.load(std::memory_order_relaxed); }

void ProfScope::begin(const char* kind, double work) {
  ProfRec r{};
  strncpy(r.kind, kind, sizeof(r.kind) - 1);
  r.work = work;
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, stream_);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
  idx_ = static_cast<int>(g_prof.size()) - 1;
}

void ProfScope::end() {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (idx_ < static_cast<int>(g_prof.size())) cudaEventRecord(g_prof[idx_].e1, stream_);
}

bool pdl_enabled() {
  // measured neutral on the SimMIM step (15.68-15.87 ms with, 15.75-15.87 ms without: the hot kernels
  // fill every SM with one persistent CTA, so a dependent grid cannot become resident early) -> opt-in
  static const bool on = getenv("VITSSL_PDL") && atoi(getenv("VITSSL_PDL")) != 0;
  return on;
}

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t outer,
                 uint64_t ld_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  auto fn = encode_fn();
  VITSSL_REQUIRE(fn != nullptr, VITSSL_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  VITSSL_REQUIRE(elem_bytes == 2 || elem_bytes == 4, VITSSL_ERR_ARG, "tmap: element size %d", elem_bytes);
  const CUtensorMapSwizzle swz = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                       : CU_TENSOR_MAP_SWIZZLE_NONE;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VITSSL_REQUIRE(r == CUDA_SUCCESS, VITSSL_ERR_CUDA,
                 "cuTensorMapEncodeTiled(2d) failed: %d (inner %llu outer %llu pitch %llu box %u x %u)",
                 (int)r, (unsigned long long)inner, (unsigned long long)outer,
                 (unsigned long long)ld_bytes, box_inner, box_outer);
  return VITSSL_OK;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t ld_bytes, uint32_t box_inner, uint32_t box_outer) {
  return make_tmap_2d(out, base, 2, inner, outer, ld_bytes, box_inner, box_outer, 128);
}

int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t d1, uint64_t d2,
                      uint64_t ld1_bytes, uint64_t ld2_bytes, uint32_t box_inner, uint32_t box_d1,
                      uint32_t box_d2) {
  return make_tmap_bf16_3d_sw(out, base, inner, d1, d2, ld1_bytes, ld2_bytes, box_inner, box_d1, box_d2, 128);
}

int make_tmap_bf16_3d_sw(CUtensorMap* out, const void* base, uint64_t inner, uint64_t d1, uint64_t d2,
                         uint64_t ld1_bytes, uint64_t ld2_bytes, uint32_t box_inner, uint32_t box_d1,
                         uint32_t box_d2, int swizzle_bytes) {
  auto fn = encode_fn();
  const CUtensorMapSwizzle swz = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                       : CU_TENSOR_MAP_SWIZZLE_NONE;
  VITSSL_REQUIRE(fn != nullptr, VITSSL_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {inner, d1, d2};
  cuuint64_t strides[2] = {ld1_bytes, ld2_bytes};
  cuuint32_t box[3] = {box_inner, box_d1, box_d2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VITSSL_REQUIRE(r == CUDA_SUCCESS, VITSSL_ERR_CUDA,
                 "cuTensorMapEncodeTiled(3d) failed: %d (dims %llu %llu %llu pitches %llu %llu)",
                 (int)r, (unsigned long long)inner, (unsigned long long)d1, (unsigned long long)d2,
                 (unsigned long long)ld1_bytes, (unsigned long long)ld2_bytes);
  return VITSSL_OK;
}

}  // namespace vitssl

extern "C" int vitssl_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* vitssl_last_error(void) { return vitssl::g_error; }

extern "C" int vitssl_num_sms(void) { return vitssl::num_sms(); }

extern "C" int64_t vitssl_launch_count(int reset) {
  const long long v = vitssl::g_launches.load();
  if (reset) vitssl::g_launches.store(0);
  return v;
}

extern "C" int vitssl_profile_begin(void) {
  std::lock_guard<std::mutex> lk(vitssl::g_prof_mu);
  for (auto& r : vitssl::g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  vitssl::g_prof.clear();
  vitssl::g_prof_on.store(true);
  return 0;
}

extern "C" int64_t vitssl_profile_end(void) {
  vitssl::g_prof_on.store(false);
  std::lock_guard<std::mutex> lk(vitssl::g_prof_mu);
  return static_cast<int64_t>(vitssl::g_prof.size());
}

extern "C" int vitssl_profile_read(int64_t index, char* kind_out, int64_t kind_cap, double* work, float* ms) {
  std::lock_guard<std::mutex> lk(vitssl::g_prof_mu);
  VITSSL_REQUIRE(index >= 0 && index < static_cast<int64_t>(vitssl::g_prof.size()) && kind_out && kind_cap > 0 && work && ms,
                 VITSSL_ERR_ARG, "profile_read: bad index or null output");
  const auto& r = vitssl::g_prof[index];
  cudaError_t err = cudaEventSynchronize(r.e1);  // the one place this library waits for the device
  if (err == cudaSuccess) err = cudaEventElapsedTime(ms, r.e0, r.e1);
  VITSSL_REQUIRE(err == cudaSuccess, VITSSL_ERR_CUDA, "profile_read: %s", cudaGetErrorString(err));
  strncpy(kind_out, r.kind, kind_cap - 1);
  kind_out[kind_cap - 1] = 0;
  *work = r.work;
  return 0;
}

extern "C" int vitssl_device_check(void) {
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) {
    vitssl::set_error("no CUDA device: %s", cudaGetErrorString(err));
    return VITSSL_ERR_DEVICE;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    vitssl::set_error("device %d is sm_%d%d; this library contains sm_100a code only", dev, major,
                      minor);
    return VITSSL_ERR_DEVICE;
  }
  return 0;
}
