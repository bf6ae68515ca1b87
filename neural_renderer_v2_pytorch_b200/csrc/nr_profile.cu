#include <cstdlib>
// nr_profile.cu -- optional per-kernel CUDA-event timing (nr_profile_enable / nr_profile_collect).
#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/nr_b200.h"
#include "nr_kernels.h"

namespace nr {

namespace {
std::atomic<int> g_on{0};
std::mutex g_mu;
struct Rec {
    int slot;
    cudaEvent_t a, b;
};
std::vector<Rec> g_recs;
}  // namespace

ProfScope::ProfScope(int slot, cudaStream_t stream) : slot_(slot), stream_(stream), start_(nullptr) {
    if (!g_on.load(std::memory_order_relaxed)) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, stream);
    start_ = e;
}

ProfScope::~ProfScope() {
    if (!start_) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) {
        cudaEventDestroy((cudaEvent_t)start_);
        return;
    }
    cudaEventRecord(e, stream_);
    std::lock_guard<std::mutex> lock(g_mu);
    g_recs.push_back({slot_, (cudaEvent_t)start_, e});
}

}  // namespace nr

extern "C" {

int nr_profile_enable(int on) {
    nr::g_on.store(on ? 1 : 0);
    return NR_OK;
}

int nr_profile_collect(float *ms, int32_t *launches) {
    std::vector<nr::Rec> recs;
    {
        std::lock_guard<std::mutex> lock(nr::g_mu);
        recs.swap(nr::g_recs);
    }
    for (const nr::Rec &r : recs) {
        float t = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess &&
            r.slot >= 0 && r.slot < NR_PROF_SLOTS) {
            if (ms) ms[r.slot] += t;
            if (launches) launches[r.slot] += 1;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    return NR_OK;
}

}  // extern "C"

namespace nr {
bool pdl_enabled(int which, long long pixels) {
    static int mask = -2;
    if (mask == -2) {
        const char *e = getenv("NR_PDL");
        mask = e ? atoi(e) : -1;
    }
    if (mask >= 0) return (mask & which) != 0;      // forced
    return pixels <= PDL_MAX_PIXELS;
}
}  // namespace nr
