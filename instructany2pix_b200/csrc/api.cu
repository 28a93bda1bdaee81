// C-ABI plumbing: thread-local error string, device gate (sm_100 only; there is no CPU or other-arch fallback).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include <mutex>

#include "common.cuh"
#include "tensormap.cuh"

namespace ia2p {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<int> g_dev_ok[64];     // 0 unknown, 1 ok, -1 bad (written after g_sm_count: a reader that sees != 0 sees the count)
static std::atomic<int> g_sm_count[64];

static int query_device(int dev) {
  if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return IA2P_E_DEVICE; }
  if (g_dev_ok[dev] == 0) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) { set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return IA2P_E_DEVICE; }
    g_sm_count[dev].store(prop.multiProcessorCount);
    g_dev_ok[dev].store((prop.major == 10) ? 1 : -1);
    if (g_dev_ok[dev] < 0) set_error("device %d is sm_%d%d; this library only runs on sm_100 (B200)", dev, prop.major, prop.minor);
  }
  if (g_dev_ok[dev] < 0) {
    set_error("device %d is not sm_100 (B200); no fallback path exists", dev);
    return IA2P_E_DEVICE;
  }
  return 0;
}

int check_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("cudaGetDevice: %s (no CUDA device: this library has no CPU path)", cudaGetErrorString(e)); return IA2P_E_DEVICE; }
  return query_device(dev);
}

static thread_local int g_pdl_mode = -1;        // ia2p_set_pdl: -1 = the IA2P_PDL environment default, 0 = off, 1 = on
bool pdl_enabled() {
  if (g_pdl_mode >= 0) return g_pdl_mode != 0;
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("IA2P_PDL");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}

int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || g_sm_count[dev] <= 0) return 148;
  return g_sm_count[dev];
}

EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

int make_map_3d_bf16(CUtensorMap* m, const void* base, int64_t cols, int64_t ld, int64_t tokens, int64_t batch, int box_rows,
                     const char* what) {
  EncodeTiledFn enc = tensor_map_encoder();
  IA2P_REQUIRE(enc != nullptr, IA2P_E_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  IA2P_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, IA2P_E_ALIGN, "%s: q/k/v/out base not 16-byte aligned", what);
  const cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)tokens, (cuuint64_t)batch};
  const cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)tokens};
  const cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IA2P_REQUIRE(r == CUDA_SUCCESS, IA2P_E_DRIVER, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return 0;
}

}  // namespace ia2p

extern "C" int ia2p_version(void) { return 100; }
extern "C" int ia2p_set_pdl(int mode) {
  const int prev = ia2p::g_pdl_mode;
  ia2p::g_pdl_mode = mode < 0 ? -1 : (mode ? 1 : 0);
  return prev;
}
extern "C" const char* ia2p_last_error(void) { return ia2p::g_err; }
extern "C" int ia2p_device_check(int device) {
  if (device < 0) return ia2p::check_device();
  return ia2p::query_device(device);
}
