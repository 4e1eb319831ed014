// Error channel, version and device query of libbgnn.so (see include/bgnn.h).
#include <stdarg.h>

#include <mutex>
#include <vector>

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

// 2-D strided copies between pinned host memory and the device (column blocks of survey planes: survey.py streams a band
// up block by block and its results down block by block; torch's copy_ stages strided host copies through temporaries)
extern "C" int bgnn_memcpy2d_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes, size_t height,
                                   int kind, void* stream) {
    BGNN_CHECK_ARG(dst && src && dpitch >= width_bytes && spitch >= width_bytes && (kind == 1 || kind == 2), "bad 2-D copy");
    if (width_bytes == 0 || height == 0) return BGNN_OK;
    BGNN_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, height,
                                kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return BGNN_OK;
}

// ---------------------------------------------------------------- per-kernel timing (bench aid)
struct ProfRec { int slot; cudaEvent_t a, b; cudaStream_t s; bool open; };
// The one piece of process-wide state in the library: a measurement aid, off by default, guarded by a mutex (launches from
// several host threads may record concurrently; a start/stop pair is matched per stream).
static bool g_prof_on = false;
static std::vector<ProfRec> g_recs;
static std::mutex g_prof_mu;

void bgnn_prof_start(int slot, cudaStream_t s) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r; r.slot = slot; r.s = s; r.open = true;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
    g_recs.push_back(r);
}
void bgnn_prof_stop(cudaStream_t s) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto it = g_recs.rbegin(); it != g_recs.rend(); ++it)
        if (it->open && it->s == s) { cudaEventRecord(it->b, s); it->open = false; return; }
}
extern "C" int bgnn_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_recs.clear();
    g_prof_on = on != 0;
    return BGNN_OK;
}
extern "C" int bgnn_profile_collect(int cap, int32_t* slots, float* ms) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    int n = 0;
    for (auto& r : g_recs) {
        BGNN_CUDA(cudaEventSynchronize(r.b));
        float t = 0.f;
        BGNN_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
        if (n < cap && slots && ms) { slots[n] = r.slot; ms[n] = t; }
        ++n;
    }
    return n;
}
