// GPU timeline of a region of a process from CUPTI activity records (nsys is not installed in this
// image): start/end of every kernel and memcpy -> busy time (union of the intervals), span, idle
// fraction, launches.  Measurement tool only (tools/gpu_timeline.py); it is not linked into
// libbnpc_b200.so and never active in timed runs.
//
//   tl_start()                     enable CONCURRENT_KERNEL + MEMCPY activity records
//   tl_stop(double out[8])         flush; out = {span_ms, busy_ms, kernels, memcpys, kernel_ms_sum,
//                                  max_concurrency, first_ns (low 32 bits), dropped}
#include <cupti.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

namespace {
struct Iv { uint64_t a, b; };
std::vector<Iv> g_kernels, g_copies;
std::mutex g_mu;
size_t g_dropped = 0;

void CUPTIAPI buffer_requested(uint8_t** buffer, size_t* size, size_t* max_records) {
    *size = 8u << 20;
    *buffer = (uint8_t*)aligned_alloc(8, *size);
    *max_records = 0;
}

void CUPTIAPI buffer_completed(CUcontext, uint32_t, uint8_t* buffer, size_t, size_t valid) {
    CUpti_Activity* rec = nullptr;
    std::lock_guard<std::mutex> lock(g_mu);
    while (cuptiActivityGetNextRecord(buffer, valid, &rec) == CUPTI_SUCCESS) {
        if (rec->kind == CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL || rec->kind == CUPTI_ACTIVITY_KIND_KERNEL) {
            const CUpti_ActivityKernel9* k = (const CUpti_ActivityKernel9*)rec;
            g_kernels.push_back(Iv{k->start, k->end});
        } else if (rec->kind == CUPTI_ACTIVITY_KIND_MEMCPY) {
            const CUpti_ActivityMemcpy5* m = (const CUpti_ActivityMemcpy5*)rec;
            g_copies.push_back(Iv{m->start, m->end});
        }
    }
    free(buffer);
}

bool g_registered = false;
}  // namespace

extern "C" {

int tl_start(void) {
    {
        std::lock_guard<std::mutex> lock(g_mu);
        g_kernels.clear();
        g_copies.clear();
        g_dropped = 0;
    }
    if (!g_registered) {
        if (cuptiActivityRegisterCallbacks(buffer_requested, buffer_completed) != CUPTI_SUCCESS) return 1;
        g_registered = true;
    }
    if (cuptiActivityEnable(CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL) != CUPTI_SUCCESS) return 2;
    cuptiActivityEnable(CUPTI_ACTIVITY_KIND_MEMCPY);
    return 0;
}

int tl_stop(double* out) {
    cuptiActivityFlushAll(1);
    cuptiActivityDisable(CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL);
    cuptiActivityDisable(CUPTI_ACTIVITY_KIND_MEMCPY);
    std::lock_guard<std::mutex> lock(g_mu);
    for (int i = 0; i < 8; ++i) out[i] = 0.0;
    if (g_kernels.empty()) return 0;
    std::vector<Iv> v = g_kernels;
    std::sort(v.begin(), v.end(), [](const Iv& x, const Iv& y) { return x.a < y.a; });
    uint64_t first = v.front().a, last = 0, busy = 0, cur_a = v.front().a, cur_b = v.front().b, sum = 0;
    for (const Iv& iv : v) {
        sum += iv.b - iv.a;
        last = std::max(last, iv.b);
        if (iv.a > cur_b) { busy += cur_b - cur_a; cur_a = iv.a; cur_b = iv.b; }
        else cur_b = std::max(cur_b, iv.b);
    }
    busy += cur_b - cur_a;
    // largest number of kernels in flight at once
    std::vector<std::pair<uint64_t, int>> ev;
    ev.reserve(2 * v.size());
    for (const Iv& iv : v) { ev.push_back({iv.a, 1}); ev.push_back({iv.b, -1}); }
    std::sort(ev.begin(), ev.end());
    int depth = 0, max_depth = 0;
    for (auto& e : ev) { depth += e.second; max_depth = std::max(max_depth, depth); }
    out[0] = (last - first) * 1e-6;
    out[1] = busy * 1e-6;
    out[2] = (double)v.size();
    out[3] = (double)g_copies.size();
    out[4] = sum * 1e-6;
    out[5] = (double)max_depth;
    out[7] = (double)g_dropped;
    return 0;
}

}  // extern "C"
