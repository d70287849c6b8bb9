// Error plumbing, device queries and TMA tensor-map encoding shared by every entry point.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace srnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return static_cast<int>(e);
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return dev < 0 ? 0 : (dev >= kMaxDevices ? kMaxDevices - 1 : dev);
}

int sm_count() {
  static std::atomic<int> cached[kMaxDevices];
  const int dev = current_device();
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      return 148;
    }
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

static int make_tmap_impl(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, bool swizzle128) {
  return make_tmap_impl(out, base, rank, dims, strides_bytes, box,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
}
int make_tmap_bf16_sw64(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap_impl(out, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_64B);
}

static int make_tmap_impl(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return SRNN_ERR_DEVICE;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                  gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu] strides=[%llu,%llu] box=[%u,%u,%u] base=%p",
              static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
              (unsigned long long)(rank > 2 ? strides_bytes[1] : 0), box[0], rank > 1 ? box[1] : 0,
              rank > 2 ? box[2] : 0, base);
    return SRNN_ERR_ARG;
  }
  return SRNN_OK;
}

}  // namespace srnn

namespace srnn {
static bool g_pdl = false;
bool pdl_enabled() { return g_pdl; }
}  // namespace srnn

extern "C" int srnn_set_pdl(int32_t on) {
  srnn::g_pdl = on != 0;
  return SRNN_OK;
}
extern "C" const char* srnn_last_error(void) { return srnn::g_err; }
extern "C" int srnn_abi_version(void) { return SRNN_ABI_VERSION; }
extern "C" int srnn_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  SRNN_CUDA(cudaGetDevice(&dev));
  if (sms) SRNN_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  if (major) SRNN_CUDA(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
  if (minor) SRNN_CUDA(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
  return SRNN_OK;
}

// ---------------------------------------------------------------------------------------------
// Device probe: how many thread-block clusters of `cluster_size` CTAs (each with `smem_bytes` of dynamic shared memory
// and `threads` threads) can be co-resident?  Used by scripts/cluster_probe.py to choose decompositions (clusters of 16
// need the non-portable attribute and one whole GPC each).
// ---------------------------------------------------------------------------------------------
namespace srnn {
__global__ void probe_kernel(int* out) {
  extern __shared__ uint8_t probe_smem[];
  if (out && threadIdx.x == 0 && blockIdx.x == 0) out[0] = static_cast<int>(probe_smem[0]);
}
}  // namespace srnn

extern "C" int srnn_probe_clusters(int32_t cluster_size, int32_t smem_bytes, int32_t threads, int32_t* max_clusters) {
  SRNN_CHECK_ARG(cluster_size >= 1 && cluster_size <= 16 && smem_bytes >= 0 && threads > 0 && max_clusters,
                 "probe_clusters: bad arguments");
  auto kern = srnn::probe_kernel;
  SRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  if (cluster_size > 8) SRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cluster_size * 64);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  SRNN_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
  *max_clusters = n;
  return SRNN_OK;
}
