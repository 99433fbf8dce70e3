// NVRTC / driver-API plumbing shared by the run-time specialised kernels (see jit.cuh).
#include "jit.cuh"

#include <dlfcn.h>

#include <cstddef>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

namespace tsu_jit {
namespace {

struct Api {
  bool tried = false, ok = false;
  int (*nvrtcCreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*nvrtcCompileProgram)(void*, int, const char* const*) = nullptr;
  int (*nvrtcGetCUBINSize)(void*, size_t*) = nullptr;
  int (*nvrtcGetCUBIN)(void*, char*) = nullptr;
  int (*nvrtcGetProgramLogSize)(void*, size_t*) = nullptr;
  int (*nvrtcGetProgramLog)(void*, char*) = nullptr;
  int (*nvrtcDestroyProgram)(void**) = nullptr;
  int (*cuModuleLoadData)(void**, const void*) = nullptr;
  int (*cuModuleGetFunction)(void**, void*, const char*) = nullptr;
  int (*cuLaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**,
                        void**) = nullptr;
  int (*cuFuncSetAttribute)(void*, int, int) = nullptr;
};
Api g_api;
std::mutex g_mutex;

template <typename F>
bool sym(void* lib, const char* name, F& fn) {
  fn = reinterpret_cast<F>(dlsym(lib, name));
  return fn != nullptr;
}

bool load_locked() {
  if (g_api.tried) return g_api.ok;
  g_api.tried = true;
  void* nv = dlopen("libnvrtc.so", RTLD_NOW | RTLD_LOCAL);
  if (!nv) nv = dlopen("libnvrtc.so.12", RTLD_NOW | RTLD_LOCAL);
  void* cu = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
  if (!nv || !cu) return false;
  g_api.ok = sym(nv, "nvrtcCreateProgram", g_api.nvrtcCreateProgram) &&
             sym(nv, "nvrtcCompileProgram", g_api.nvrtcCompileProgram) &&
             sym(nv, "nvrtcGetCUBINSize", g_api.nvrtcGetCUBINSize) && sym(nv, "nvrtcGetCUBIN", g_api.nvrtcGetCUBIN) &&
             sym(nv, "nvrtcGetProgramLogSize", g_api.nvrtcGetProgramLogSize) &&
             sym(nv, "nvrtcGetProgramLog", g_api.nvrtcGetProgramLog) &&
             sym(nv, "nvrtcDestroyProgram", g_api.nvrtcDestroyProgram) &&
             sym(cu, "cuModuleLoadData", g_api.cuModuleLoadData) &&
             sym(cu, "cuModuleGetFunction", g_api.cuModuleGetFunction) && sym(cu, "cuLaunchKernel", g_api.cuLaunchKernel) &&
             sym(cu, "cuFuncSetAttribute", g_api.cuFuncSetAttribute);
  return g_api.ok;
}

}  // namespace

bool available() {
  std::lock_guard<std::mutex> lock(g_mutex);
  return load_locked();
}

void* compile(const std::string& source, const char* tu_name, const char* kernel_name, const char* include_dir,
              std::string& log) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (!load_locked()) {
    log = "libnvrtc.so / libcuda.so.1 not available";
    return nullptr;
  }
  if (cudaFree(0) != cudaSuccess) {  // makes the primary context of the current device current
    log = "no CUDA context";
    return nullptr;
  }
  void* prog = nullptr;
  if (g_api.nvrtcCreateProgram(&prog, source.c_str(), tu_name, 0, nullptr, nullptr) != 0) {
    log = "nvrtcCreateProgram failed";
    return nullptr;
  }
  const std::string inc = std::string("-I") + include_dir;
  const char* opts[] = {"--gpu-architecture=sm_100a", inc.c_str(), "-std=c++17", "-lineinfo"};
  const int rc = g_api.nvrtcCompileProgram(prog, 4, opts);
  if (rc != 0) {
    size_t n = 0;
    if (g_api.nvrtcGetProgramLogSize(prog, &n) == 0 && n > 1) {
      std::vector<char> buf(n);
      g_api.nvrtcGetProgramLog(prog, buf.data());
      log.assign(buf.data());
    } else {
      log = "nvrtcCompileProgram failed";
    }
    g_api.nvrtcDestroyProgram(&prog);
    return nullptr;
  }
  size_t n = 0;
  g_api.nvrtcGetCUBINSize(prog, &n);
  std::vector<char> cubin(n);
  g_api.nvrtcGetCUBIN(prog, cubin.data());
  g_api.nvrtcDestroyProgram(&prog);
  void *mod = nullptr, *fn = nullptr;
  if (g_api.cuModuleLoadData(&mod, cubin.data()) != 0 || g_api.cuModuleGetFunction(&fn, mod, kernel_name) != 0) {
    log = "cuModuleLoadData / cuModuleGetFunction failed";
    return nullptr;
  }
  return fn;
}

int launch(void* fn, unsigned grid, unsigned block, unsigned smem_bytes, void* stream, void** args) {
  const int rc = g_api.cuLaunchKernel(fn, grid, 1, 1, block, 1, 1, smem_bytes, stream, args, nullptr);
  return rc == 0 ? 0 : 999;
}

int set_max_dynamic_smem(void* fn, int bytes) {
  return g_api.cuFuncSetAttribute(fn, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, bytes) == 0 ? 0 : 999;
}

}  // namespace tsu_jit
