// Error channel, version and device query of libbgnn.so (see include/bgnn.h).
#include <stdarg.h>

#include <mutex>
#include <vector>

#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges are no-ops unless a profiler (nsys / ncu --nvtx) is attached

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

// NVTX range per library launch site: the slot names of include/bgnn.h, so a timeline groups the kernels of a
// step by stage.  Pushed / popped on the launching host thread around the launch call(s) of a ProfScope.
static const char* slot_name(int slot) {
    switch (slot) {
        case 0: return "bgnn:gb_count"; case 1: return "bgnn:gb_scan"; case 2: return "bgnn:gb_ids"; case 3: return "bgnn:gb_emit";
        case 4: return "bgnn:gb_coo"; case 10: return "bgnn:mlp0"; case 11: return "bgnn:mlp1";
        case 60: return "bgnn:heads"; case 61: return "bgnn:heads_finish"; case 70: return "bgnn:scatter_tile";
        case 71: return "bgnn:stitch_add_tile"; case 72: return "bgnn:stitch_finalize"; case 73: return "bgnn:apply_corrections";
        case 74: return "bgnn:graph_to_grid"; case 75: return "bgnn:vr_apply_results"; case 76: return "bgnn:sidecar_add";
        case 77: return "bgnn:survey_emit"; case 78: return "bgnn:stitch_add_tile_nodes"; case 80: return "bgnn:csr_aggregate";
        case 81: return "bgnn:linear";
        default: return slot >= 40 && slot < 60 ? "bgnn:aggregate_layer" : slot >= 20 && slot < 40 ? "bgnn:project_layer" : "bgnn:kernel";
    }
}

void bgnn_prof_start(int slot, cudaStream_t s) {
    nvtxRangePushA(slot_name(slot));
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r; r.slot = slot; r.s = s; r.open = true;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
    g_recs.push_back(r);
}
void bgnn_prof_stop(cudaStream_t s) {
    nvtxRangePop();
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
