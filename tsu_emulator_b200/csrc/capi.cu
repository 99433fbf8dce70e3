// Library-level entry points: version, error strings, device probe, Philox helpers.
#include <cstring>

#include "common.cuh"
#include "philox.cuh"

namespace {

__global__ void philox_fill_kernel(uint32_t* out, uint64_t n, uint32_t k0, uint32_t k1, uint32_t offset) {
  const uint64_t blk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one Philox block = 4 words
  const uint64_t base = blk * 4;
  if (base >= n) return;
  tsu_u32x4 o = tsu_philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), offset, TSU_STREAM_FILL, k0, k1);
  out[base] = o.x;
  if (base + 1 < n) out[base + 1] = o.y;
  if (base + 2 < n) out[base + 2] = o.z;
  if (base + 3 < n) out[base + 3] = o.w;
}

}  // namespace

extern "C" {

int tsu_version(void) { return TSU_B200_ABI_VERSION; }

const char* tsu_error_string(int code) {
  if (code == TSU_OK) return "ok";
  if (code == TSU_ERR_INVALID_ARG) return "invalid argument";
  if (code == TSU_ERR_UNSUPPORTED) return "unsupported configuration";
  if (code == TSU_ERR_NO_DEVICE) return "no CUDA device";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown error";
}

int tsu_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return TSU_ERR_NO_DEVICE;
  }
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return (int)e;
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return TSU_OK;
}

void tsu_philox4x32_10_host(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  tsu_u32x4 o = tsu_philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  out[0] = o.x;
  out[1] = o.y;
  out[2] = o.z;
  out[3] = o.w;
}

// ---- device buffers other ranks of the box can map (CUDA IPC): halos of the row-slab driver ----------------
int tsu_peer_alloc(void** d_ptr, size_t bytes) {
  TSU_CHECK_ARG(d_ptr && bytes > 0);
  cudaError_t e = cudaMalloc(d_ptr, bytes);  // plain cudaMalloc: pool / VMM allocations cannot be exported
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(*d_ptr, 0, bytes);
  return e == cudaSuccess ? TSU_OK : (int)e;
}

int tsu_peer_free(void* d_ptr) {
  cudaError_t e = cudaFree(d_ptr);
  return e == cudaSuccess ? TSU_OK : (int)e;
}

int tsu_peer_get_handle(void* d_ptr, unsigned char handle[64]) {
  TSU_CHECK_ARG(d_ptr && handle);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, d_ptr);
  if (e != cudaSuccess) return (int)e;
  memcpy(handle, &h, 64);
  return TSU_OK;
}

int tsu_peer_open_handle(const unsigned char handle[64], void** d_ptr) {
  TSU_CHECK_ARG(handle && d_ptr);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  cudaError_t e = cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  return e == cudaSuccess ? TSU_OK : (int)e;
}

int tsu_peer_read_u32(const void* d_ptr, uint32_t* h_out) {
  TSU_CHECK_ARG(d_ptr && h_out);
  cudaError_t e = cudaMemcpy(h_out, d_ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost);  // synchronises
  return e == cudaSuccess ? TSU_OK : (int)e;
}

int tsu_peer_close_handle(void* d_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(d_ptr);
  return e == cudaSuccess ? TSU_OK : (int)e;
}

int tsu_philox_fill_u32(uint32_t* d_out, uint64_t n, uint64_t seed, uint32_t offset, uintptr_t stream) {
  TSU_CHECK_ARG(d_out || n == 0);
  if (n == 0) return TSU_OK;
  const uint64_t blocks4 = (n + 3) / 4;
  const unsigned grid = (unsigned)((blocks4 + 255) / 256);
  philox_fill_kernel<<<grid, 256, 0, tsu_stream(stream)>>>(d_out, n, (uint32_t)seed, (uint32_t)(seed >> 32), offset);
  TSU_RETURN_LAUNCH_STATUS();
}

}  // extern "C"
