// Library runtime: error reporting, driver entry points, tensor-map construction.
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void gcc_set_error(const char* file, int line, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s:%d: %s", file, line, msg);
}

extern "C" const char* gcc_last_error(void) { return g_err; }

#include <stdlib.h>
int gcc_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("GCC_B200_PDL");
    on = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return on;
}

unsigned long long g_gcc_launches = 0;
extern "C" long long gcc_launch_count(void) { return (long long)g_gcc_launches; }

extern "C" int gcc_abi_version(void) { return 1; }

// The library links its own (static) CUDA runtime.  A host thread that has not yet issued a runtime
// call through THIS runtime instance (e.g. torch's autograd worker threads) has no context bound for
// the driver-API tensor-map encoder; bind the device's primary context once per thread.
extern "C" int gcc_bind_thread(int device) {
  if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess) {
    gcc_set_error(__FILE__, __LINE__, "gcc_bind_thread: cannot bind the device context");
    cudaGetLastError();
    return GCC_ERR_CUDA;
  }
  return GCC_OK;
}

// Fails (non-zero) unless the current device is an sm_100 part: there is no fallback path.
extern "C" int gcc_check_device(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    gcc_set_error(__FILE__, __LINE__, "no CUDA device");
    return GCC_ERR_CUDA;
  }
  if (prop.major != 10) {
    gcc_set_error(__FILE__, __LINE__, "gcc_b200 kernels are built for sm_100a only");
    return GCC_ERR_ARG;
  }
  return GCC_OK;
}

PFN_encodeTiled gcc_get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess) {
    gcc_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled entry point not available");
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

int gcc_make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box) {
  return gcc_make_tmap_bf16_sw(map, base, rank, dims, strides_bytes, box, 1);
}

// swizzle128 = 0: plain (un-swizzled) box, used for the 8-channel image operand whose 16-byte pixels land as
// canonical no-swizzle core matrices
int gcc_make_tmap_bf16_sw(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, int swizzle128) {
  PFN_encodeTiled enc = gcc_get_encode_tiled();
  if (!enc) return GCC_ERR_DRIVER;
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstrides[i] = strides_bytes[i];
  }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstrides,
                   gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf),
             "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu box %u %u %u %u base %p", (int)r,
             rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
             rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
    gcc_set_error(__FILE__, __LINE__, buf);
    return GCC_ERR_DRIVER;
  }
  return GCC_OK;
}
