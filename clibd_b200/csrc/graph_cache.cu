#include "graph_cache.h"

#include <cstdlib>
#include <list>
#include <mutex>
#include <unordered_map>
#include <unordered_set>

#include "common.cuh"

namespace clibd {

int64_t launch_count_now();            // loss_api.cu
void add_launches(int64_t n);          // loss_api.cu
bool profiling_enabled();              // loss_api.cu

namespace {

struct Entry {
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
    std::list<std::string>::iterator lru;
};

constexpr size_t kMaxEntries = 64;
constexpr int kMissStreak = 12;      // this many misses in a row: the caller's pointers do not repeat ...
constexpr int kCooldownCalls = 512;  // ... so run un-graphed for a while before trying again

std::mutex g_mu;
std::unordered_map<std::string, Entry> g_cache;
std::list<std::string> g_lru;  // front = most recently used
std::unordered_set<std::string> g_seen;  // argument tuples that have run once, un-captured
cudaStream_t g_capture_stream[64] = {};
int g_miss_streak = 0;
int g_cooldown = 0;
bool g_broken = false;  // a capture failed once: never try again in this process

bool env_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("CLIBD_GRAPHS");
        return e == nullptr || std::atoi(e) != 0;
    }();
    return on;
}

}  // namespace

int run_graphed(const GraphKey& key, bool eligible, cudaStream_t stream, const std::function<int(cudaStream_t)>& body) {
    if (!eligible || !env_enabled() || profiling_enabled()) return body(stream);
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
        cudaGetLastError();
        return body(stream);  // the caller is capturing its own graph: just contribute the launches
    }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return body(stream);
    std::unique_lock<std::mutex> lk(g_mu);
    if (g_broken) {
        lk.unlock();
        return body(stream);
    }
    if (g_cooldown > 0) {
        --g_cooldown;
        lk.unlock();
        return body(stream);
    }
    std::string full = key.bytes;
    full.append(reinterpret_cast<const char*>(&dev), sizeof(dev));
    auto it = g_cache.find(full);
    if (it != g_cache.end()) {
        g_miss_streak = 0;
        g_lru.splice(g_lru.begin(), g_lru, it->second.lru);
        add_launches(it->second.launches);
        CLIBD_CHECK_CUDA(cudaGraphLaunch(it->second.exec, stream));
        return 0;
    }
    // The first call with an argument tuple runs directly: it loads the kernels' modules (lazy loading), resolves driver
    // entry points and sets function attributes -- none of which may happen inside a capture -- and a tuple that never
    // comes back costs nothing.  The second call captures.
    if (g_seen.find(full) == g_seen.end()) {
        if (g_seen.size() > 4096) g_seen.clear();
        g_seen.insert(full);
        ++g_miss_streak;
        if (g_miss_streak >= kMissStreak) {
            g_miss_streak = 0;
            g_cooldown = kCooldownCalls;
        }
        lk.unlock();
        return body(stream);
    }
    if (++g_miss_streak >= kMissStreak) {
        g_miss_streak = 0;
        g_cooldown = kCooldownCalls;
        lk.unlock();
        return body(stream);
    }
    // ---- capture on an internal stream (the caller's may be the legacy default stream, which cannot be captured)
    if (g_capture_stream[dev] == nullptr &&
        cudaStreamCreateWithFlags(&g_capture_stream[dev], cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        g_broken = true;
        lk.unlock();
        return body(stream);
    }
    cudaStream_t cap = g_capture_stream[dev];
    if (cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        g_broken = true;
        lk.unlock();
        return body(stream);
    }
    const int64_t l0 = launch_count_now();
    const int rc = body(cap);
    const int64_t captured = launch_count_now() - l0;
    cudaGraph_t graph = nullptr;
    const cudaError_t e_end = cudaStreamEndCapture(cap, &graph);
    add_launches(-captured);  // nothing has run yet
    if (rc != 0) {            // an argument error inside body: report it, nothing was launched
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return rc;
    }
    cudaGraphExec_t exec = nullptr;
    if (e_end != cudaSuccess || graph == nullptr || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        g_broken = true;  // e.g. a library call inside body that cannot be captured on this driver
        lk.unlock();
        return body(stream);
    }
    cudaGraphDestroy(graph);
    if (g_cache.size() >= kMaxEntries) {
        const std::string& victim = g_lru.back();
        auto vit = g_cache.find(victim);
        if (vit != g_cache.end()) {
            cudaGraphExecDestroy(vit->second.exec);
            g_cache.erase(vit);
        }
        g_lru.pop_back();
    }
    g_lru.push_front(full);
    Entry en;
    en.exec = exec;
    en.launches = captured;
    en.lru = g_lru.begin();
    g_cache.emplace(full, en);
    add_launches(captured);
    CLIBD_CHECK_CUDA(cudaGraphLaunch(exec, stream));
    return 0;
}

}  // namespace clibd
