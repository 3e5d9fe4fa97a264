// Library-level entry points of libal3d.so: version, error text, device probe.
#include "common.cuh"
#include "../../include/al3d.h"

namespace al3d {

char *error_buffer()
{
    static thread_local char buf[512] = {0};
    return buf;
}

int set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return 1;
}

}  // namespace al3d

extern "C" int al3d_abi_version(void) { return AL3D_ABI_VERSION; }

extern "C" const char *al3d_last_error(void) { return al3d::error_buffer(); }

extern "C" int al3d_device_supports_tcgen05(void)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
    return (p.major == 10 && p.minor == 0) ? 1 : 0;
}
