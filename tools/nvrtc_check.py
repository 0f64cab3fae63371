"""Offline check (no GPU): the public device header + the test user kernels compile under NVRTC for sm_100a,
assembled exactly as hiprtBuildTraceKernels does."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiprt_b200 import api
from cuda.bindings import nvrtc


def compile_user_source(user, dispatch=""):
    L = api.lib()
    names = C.POINTER(C.c_char_p)()
    srcs = C.POINTER(C.c_char_p)()
    n = L.hiprtB200GetDeviceHeaders(C.byref(names), C.byref(srcs))
    hdr_names = [names[i] for i in range(n)]
    hdr_srcs = [srcs[i] for i in range(n)]
    full = "#include <hiprt/hiprt_device.h>\n" + dispatch + user
    err, prog = nvrtc.nvrtcCreateProgram(full.encode(), b"user.cu", n, hdr_srcs, hdr_names)
    opts = [b"--gpu-architecture=sm_100a", b"-std=c++17", b"-DHIPRT_B200=1"]
    (err,) = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    _, sz = nvrtc.nvrtcGetProgramLogSize(prog)
    log = b" " * sz
    nvrtc.nvrtcGetProgramLog(prog, log)
    return int(err), log.decode(errors="replace")


DISPATCH = '''
__device__ bool intersectSphere( const hiprtRay& ray, const void* data, void* payload, hiprtHit& hit );
__device__ bool cutoutFilter( const hiprtRay& ray, const void* data, void* payload, const hiprtHit& hit );
__device__ bool intersectFunc( uint32_t geomType, uint32_t rayType, const hiprtFuncTableHeader& tableHeader, const hiprtRay& ray, void* payload, hiprtHit& hit )
{ return intersectSphere( ray, tableHeader.funcDataSets[0].intersectFuncData, payload, hit ); }
__device__ bool filterFunc( uint32_t geomType, uint32_t rayType, const hiprtFuncTableHeader& tableHeader, const hiprtRay& ray, void* payload, const hiprtHit& hit )
{ return cutoutFilter( ray, tableHeader.funcDataSets[0].filterFuncData, payload, hit ); }
'''

if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    err, log = compile_user_source(open(os.path.join(root, "tests", "kernels", "user_kernels.cu")).read(), DISPATCH)
    print("nvrtc result", err)
    print(log[:4000])
    sys.exit(err)
