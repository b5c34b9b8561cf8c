// CUDA-graph replay of an entry point's launch sequence (launch-bound batches).
//
// A loss step is 40-60 kernel launches.  At the benchmark batch (N = 32768) they are long and the launch overhead is
// noise; at N = 4096 -- BASELINE config 2, or the per-rank share of a sharded step -- the step is a few hundred
// microseconds and the per-launch cost on the host and the gaps between dependent kernels on the device are most of
// it.  An entry point's launches depend only on its ARGUMENTS (shapes, pointers, scalar values passed by value), never
// on device data, so the sequence is captured once per distinct argument tuple and replayed with one cudaGraphLaunch.
// A training loop calls with the same pointers step after step (caching allocators hand the same blocks back); when
// it does not, the cache keeps missing and switches itself off for a while instead of paying an instantiate per call.
#pragma once
#include <cuda_runtime.h>

#include <cstring>
#include <functional>
#include <string>

namespace clibd {

struct GraphKey {
    std::string bytes;
    template <typename T>
    GraphKey& add(const T& v) {
        bytes.append(reinterpret_cast<const char*>(&v), sizeof(T));
        return *this;
    }
    template <typename T>
    GraphKey& add_array(const T* v, int count) {
        if (v != nullptr) bytes.append(reinterpret_cast<const char*>(v), sizeof(T) * count);
        else bytes.append(sizeof(T) * count, '\xff');
        return *this;
    }
};

// Runs body(s): directly on `stream` when `eligible` is false (or graphs are switched off: CLIBD_GRAPHS=0, event
// profiling on, a capture already in progress on `stream`), else through the graph cache.  Returns body's code.
int run_graphed(const GraphKey& key, bool eligible, cudaStream_t stream, const std::function<int(cudaStream_t)>& body);

// launch-bound regime: the O(N n d) kernels of a step take well under a millisecond
inline bool graph_worthwhile(int64_t N, int64_t n) { return N * n <= (int64_t(1) << 28); }

}  // namespace clibd
