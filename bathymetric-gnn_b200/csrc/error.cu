// Error channel, version and device query of libbgnn.so (see include/bgnn.h).
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void bgnn_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* bgnn_last_error(void) { return g_err; }
extern "C" int bgnn_version(void) { return BGNN_VERSION; }

extern "C" int bgnn_device_info(int device, int* out3) {
    BGNN_CHECK_ARG(out3, "out3 is null");
    cudaDeviceProp p;
    BGNN_CUDA(cudaGetDeviceProperties(&p, device));
    out3[0] = p.multiProcessorCount;
    out3[1] = p.major;
    out3[2] = p.minor;
    return BGNN_OK;
}
