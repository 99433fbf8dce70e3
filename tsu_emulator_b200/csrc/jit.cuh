// Run-time compilation of kernel specialisations (NVRTC + driver API, both opened with dlopen so that the library has
// no link-time dependency on them).  Users: the table-specialised lattice half-sweep (ising2d.cu) and the Langevin
// kernel for traced Python energies (langevin.cu).
#pragma once
#include <string>

namespace tsu_jit {

// true when libnvrtc and libcuda could be opened and every entry point was found
bool available();

// compile `source` (translation unit name `tu_name`) for sm_100a with -I include_dir and return the CUfunction of the
// extern "C" kernel `kernel_name`, or nullptr with the reason in `log`
void* compile(const std::string& source, const char* tu_name, const char* kernel_name, const char* include_dir,
              std::string& log);

// cuLaunchKernel; returns 0 or 999 (a driver-API launch failure has no cudaError_t of its own)
int launch(void* fn, unsigned grid, unsigned block, unsigned smem_bytes, void* stream, void** args);

// raise the dynamic shared memory limit of a compiled kernel (cuFuncSetAttribute)
int set_max_dynamic_smem(void* fn, int bytes);

}  // namespace tsu_jit
