"""Compile a translation unit with NVRTC for sm_100a without a GPU (what the library does at run time):
    python tools/nvrtc_check.py file.cu        -> prints the log, exit code 0 on success"""
import ctypes, os, sys

def nvrtc_compile(src: str, name: str = "tu.cu", include_dir: str = None):
    nv = None
    for lib in ("libnvrtc.so", "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"):
        try:
            nv = ctypes.CDLL(lib); break
        except OSError:
            continue
    if nv is None:
        return None, "libnvrtc not found"
    prog = ctypes.c_void_p()
    nv.nvrtcCreateProgram(ctypes.byref(prog), src.encode(), name.encode(), 0, None, None)
    inc = include_dir or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tsu_emulator_b200", "csrc")
    opts = [b"--gpu-architecture=sm_100a", ("-I" + inc).encode(), b"-std=c++17", b"-lineinfo"]
    arr = (ctypes.c_char_p * len(opts))(*opts)
    rc = nv.nvrtcCompileProgram(prog, len(opts), arr)
    n = ctypes.c_size_t()
    nv.nvrtcGetProgramLogSize(prog, ctypes.byref(n))
    buf = ctypes.create_string_buffer(n.value or 1)
    nv.nvrtcGetProgramLog(prog, buf)
    size = ctypes.c_size_t()
    cubin = None
    if rc == 0:
        nv.nvrtcGetCUBINSize(prog, ctypes.byref(size))
        cubin = ctypes.create_string_buffer(size.value)
        nv.nvrtcGetCUBIN(prog, cubin)
    nv.nvrtcDestroyProgram(ctypes.byref(prog))
    return (cubin.raw if cubin else None), buf.value.decode(errors="replace")

if __name__ == "__main__":
    cubin, log = nvrtc_compile(open(sys.argv[1]).read(), os.path.basename(sys.argv[1]))
    print(log)
    print("ok, cubin bytes:", len(cubin)) if cubin else print("FAILED")
    sys.exit(0 if cubin else 1)
