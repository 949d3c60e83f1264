#include "common.cuh"

#include <stdlib.h>

#include <atomic>

namespace vmm {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

char* error_buffer() { return g_err; }

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("VMM_PDL");
    return e != nullptr && e[0] != '0';
  }();
  return on;
}

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return VMM_ERR_CUDA;
}

int check_launch(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, where);
  return VMM_OK;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        sms <= 0)
      sms = 148;
  }
  return sms;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* gdim,
                      const uint64_t* gstride_bytes, const uint32_t* box, bool l2_256, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(VMM_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(VMM_ERR_ARG, "tensor map: base not 16-byte aligned");
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = gdim[i] ? gdim[i] : 1;
    bx[i] = box[i];
    es[i] = 1;
    if (bx[i] < 1 || bx[i] > 256) return set_error(VMM_ERR_ARG, "tensor map: box dim out of range");
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gs[i] = gstride_bytes[i];
    if (gs[i] == 0 || (gs[i] & 15) != 0) return set_error(VMM_ERR_ARG, "tensor map: stride must be a non-zero multiple of 16 bytes");
  }
  CUresult r = fn(map, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, l2_256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(error_buffer(), 512, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu, box %u %u)", (int)r, rank,
             (unsigned long long)gd[0], (unsigned long long)gd[1], bx[0], bx[1]);
    return VMM_ERR_CUDA;
  }
  return VMM_OK;
}

}  // namespace vmm

extern "C" const char* vmm_last_error(void) { return vmm::error_buffer(); }
extern "C" int vmm_abi_version(void) { return 8; }
extern "C" uint64_t vmm_launch_count(void) { return vmm::g_launches.load(); }
