// Optional measurement hooks used by bench.py to time the dominant kernel (the DMMA GEMM) live with
// CUDA events on the launching stream.  Disabled by default; the product path never enables it.
#include <atomic>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace gpb {

namespace {
struct Prof {
    std::atomic<int> enabled{0};
    std::atomic<long long> launches{0};
    std::mutex mu;
    std::vector<cudaEvent_t> begin, end;
    std::vector<int> tag;  // 0: DMMA GEMM, 1: Ozaki int8 kernel
    size_t used = 0;
    double oz_ops = 0.0;   // algorithmic int8 operations (2 x MACs) of the tag-1 launches
} g_prof;
}  // namespace

void profile_reset(int enable) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    g_prof.used = 0;
    g_prof.oz_ops = 0.0;
    g_prof.launches = 0;
    g_prof.enabled = enable;
}

bool profile_enabled() { return g_prof.enabled.load(std::memory_order_relaxed) != 0; }

void profile_count_launch() {
    if (g_prof.enabled.load(std::memory_order_relaxed)) g_prof.launches.fetch_add(1, std::memory_order_relaxed);
}

static void begin_tagged(stream_t s, int tag) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    if (g_prof.used == g_prof.begin.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        g_prof.begin.push_back(a);
        g_prof.end.push_back(b);
        g_prof.tag.push_back(0);
    }
    g_prof.tag[g_prof.used] = tag;
    cudaEventRecord(g_prof.begin[g_prof.used], to_stream(s));
}
void profile_gemm_begin(stream_t s) { begin_tagged(s, 0); }
void profile_ozaki_begin(stream_t s, double int8_ops) {
    begin_tagged(s, 1);
    std::lock_guard<std::mutex> lk(g_prof.mu);
    g_prof.oz_ops += int8_ops;
}

void profile_gemm_end(stream_t s) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    cudaEventRecord(g_prof.end[g_prof.used], to_stream(s));
    g_prof.used++;
}

static int read_tagged(int tag, double* ms_out, int64_t* n_out) {
    double total = 0.0;
    int64_t n = 0;
    for (size_t i = 0; i < g_prof.used; ++i) {
        if (g_prof.tag[i] != tag) continue;
        if (cudaEventSynchronize(g_prof.end[i]) != cudaSuccess) return GPB_ERR_LAUNCH;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_prof.begin[i], g_prof.end[i]) != cudaSuccess) return GPB_ERR_LAUNCH;
        total += ms;
        ++n;
    }
    if (ms_out) *ms_out = total;
    if (n_out) *n_out = n;
    return GPB_OK;
}
int profile_read(double* gemm_ms, int64_t* gemm_launches, int64_t* all_launches) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    int rc = read_tagged(0, gemm_ms, gemm_launches);
    if (all_launches) *all_launches = (int64_t)g_prof.launches.load();
    return rc;
}
int profile_read_ozaki(double* ms, int64_t* launches, double* int8_ops) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    if (int8_ops) *int8_ops = g_prof.oz_ops;
    return read_tagged(1, ms, launches);
}


// ---- helper streams for lookahead ---------------------------------------------------------------------
// Kept PER DEVICE (one process may drive several GPUs from several threads); created lazily, never destroyed.  Two host
// threads driving the same device share its helper streams: the fork / join events keep every call's own ordering intact.
namespace {
struct Side {
    std::mutex mu;
    cudaStream_t streams[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t events[64];
    int nev = 0, next = 0;
};
Side g_side[GPB_MAX_DEVICES];
}  // namespace

stream_t side_stream(stream_t main, int idx) {
    if (idx < 0 || idx >= 4) return main;
    Side& sd = g_side[current_device()];
    std::lock_guard<std::mutex> lk(sd.mu);
    if (!sd.streams[idx]) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
        if (cudaStreamCreateWithPriority(&sd.streams[idx], cudaStreamNonBlocking, hi) != cudaSuccess) return main;
    }
    return reinterpret_cast<stream_t>(sd.streams[idx]);
}

int stream_fork(stream_t from, stream_t to) {
    if (from == to) return GPB_OK;
    Side& sd = g_side[current_device()];
    std::lock_guard<std::mutex> lk(sd.mu);
    if (sd.nev < 64) {
        if (cudaEventCreateWithFlags(&sd.events[sd.nev], cudaEventDisableTiming) != cudaSuccess)
            return GPB_ERR_LAUNCH;
        sd.nev++;
    }
    cudaEvent_t ev = sd.events[sd.next % sd.nev];
    sd.next++;
    // re-recording an event does not disturb waits that were enqueued on its previous record
    if (cudaEventRecord(ev, to_stream(from)) != cudaSuccess) return GPB_ERR_LAUNCH;
    if (cudaStreamWaitEvent(to_stream(to), ev, 0) != cudaSuccess) return GPB_ERR_LAUNCH;
    return GPB_OK;
}

}  // namespace gpb
