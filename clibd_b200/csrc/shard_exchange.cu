// Exchanges of the row-sharded loss step over PEER-MAPPED memory (one process per GPU, NVLink / NVSwitch).
//
// The reference all-gathers labels and features with NCCL (bioscanclip/model/loss_func.py:143-157) and its autograd
// backward reduce-scatters the gathered gradients (loss_func.py:97).  Here every rank maps every other rank's exchange
// buffer (a symmetric allocation made by the caller) and the exchanges are plain stores into the owners' memory:
//
//   clibd_shard_push_rows    all-gather: raw rows + inverse norms + labels of the local block -> every rank
//   clibd_shard_push_stats   first half of the statistics all-reduce: disjoint segments in place, column-sum
//   clibd_shard_reduce_stats partials into per-rank slots, summed in rank order after the caller's barrier
//   clibd_shard_push_floats  small per-rank slot vectors (each rank's grad_output)
//   clibd_shard_barrier      barrier across the ranks on the caller's stream (flags in peer-mapped memory)
//
// (the reduce-scatter of the column-side gradients is not here: the gradient GEMM's epilogue stores its rows straight
//  into the owners' slot arrays, loss_grad_gemm.cu.)  Writers and readers are separated by a barrier across the ranks
// that the caller issues on the same stream; kernel completion makes the peer stores visible system-wide.
#include "../../include/clibd_b200.h"
#include "common.cuh"
#include "loss_plan.h"

namespace clibd {
namespace {

constexpr int kThreads = 256;

struct PeerRows {
    void* x[MAX_PEERS * 3];
    float* inv[MAX_PEERS * 3];
    int64_t* labels[MAX_PEERS];
};

// One warp per (modality, local row): the row is read once (16-byte chunks when the layout allows), its squared norm
// accumulated on the way, and every chunk stored to all `world` destinations; consecutive lanes write consecutive
// 16 bytes, so a warp store is one 512-byte burst per peer.
template <typename T, bool VEC>
__global__ void shard_push_rows_kernel(const T* __restrict__ x0, const T* __restrict__ x1, const T* __restrict__ x2,
                                       const int64_t* __restrict__ labels, int64_t n, int64_t d, int rank, int world,
                                       PeerRows peers) {
    const int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= 3 * n) return;
    const int m = static_cast<int>(w / n);
    const int64_t i = w - static_cast<int64_t>(m) * n;
    const T* xm = m == 0 ? x0 : (m == 1 ? x1 : x2);
    const int64_t grow = static_cast<int64_t>(rank) * n + i;
    if (m == 0 && labels != nullptr && lane < world) peers.labels[lane][grow] = labels[i];
    if (xm == nullptr) return;
    const T* src = xm + i * d;
    float ss = 0.f;
    if constexpr (VEC) {
        constexpr int E = 16 / sizeof(T);  // elements per 16-byte chunk
        const int64_t chunks = d / E;
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        for (int64_t c = lane; c < chunks; c += 32) {
            const uint4 v = s4[c];
            if constexpr (sizeof(T) == 4) {
                const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
                for (int k = 0; k < 4; ++k) ss = fmaf(f[k], f[k], ss);
            } else {
                const T* h = reinterpret_cast<const T*>(&v);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float f = load_as_float(h, k);
                    ss = fmaf(f, f, ss);
                }
            }
            for (int q = 0; q < world; ++q)
                reinterpret_cast<uint4*>(static_cast<T*>(peers.x[q * 3 + m]) + grow * d)[c] = v;
        }
    } else {
        for (int64_t c = lane; c < d; c += 32) {
            const T v = src[c];
            const float f = load_as_float(src, c);
            ss = fmaf(f, f, ss);
            for (int q = 0; q < world; ++q) (static_cast<T*>(peers.x[q * 3 + m]) + grow * d)[c] = v;
        }
    }
    ss = warp_sum(ss);
    const float iv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize's eps (loss_func.py:55-56)
    if (lane < world) peers.inv[lane * 3 + m][grow] = iv;
}

struct PeerStats {
    float* stats[MAX_PEERS];
    float* colslots[MAX_PEERS];
    double* posslots[MAX_PEERS];
};

// blockIdx.y = destination rank q.  Column-sum partials [3, N] -> slot `rank` of q's slot array; the local segments of
// rowsum / posrow (blocks 0..2 and 6..8 of the [9, N] statistics buffer) -> the same place of q's buffer (q != rank).
__global__ void shard_push_stats_kernel(const float* __restrict__ stats, const double* __restrict__ pos, int64_t N,
                                        int64_t row0, int64_t n, int rank, PeerStats peers) {
    const int q = blockIdx.y;
    const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k < 3 * N) peers.colslots[q][static_cast<int64_t>(rank) * 3 * N + k] = stats[3 * N + k];
    if (q != rank && k < 6 * n) {
        const int64_t blk = k / n, i = k - blk * n;
        const int64_t off = (blk < 3 ? blk : blk + 3) * N + row0 + i;
        peers.stats[q][off] = stats[off];
    }
    if (k < 3) peers.posslots[q][rank * 4 + k] = pos[k];
}

__global__ void shard_reduce_stats_kernel(const float* __restrict__ colslots, const double* __restrict__ posslots,
                                          int64_t N, int world, float* __restrict__ stats, double* __restrict__ pos) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k < 3 * N) {
        float acc = 0.f;
        for (int r = 0; r < world; ++r) acc += colslots[static_cast<int64_t>(r) * 3 * N + k];
        stats[3 * N + k] = acc;
    }
    if (k < 3) {
        double acc = 0.0;
        for (int r = 0; r < world; ++r) acc += posslots[r * 4 + k];
        pos[k] = acc;
    }
}

struct PeerFloats {
    float* p[MAX_PEERS];
};

__global__ void shard_push_floats_kernel(const float* __restrict__ src, int64_t count, int rank, PeerFloats peers) {
    const int q = blockIdx.y;
    const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k < count) peers.p[q][static_cast<int64_t>(rank) * count + k] = src[k];
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

constexpr int kBarrierChannels = 4;
static_assert(MAX_PEERS <= 32, "one warp signals and waits for all peers");

struct PeerFlags {
    unsigned long long* f[MAX_PEERS];  // rank q's block: [channel][writer rank] epochs, then [channel] own epoch counter
};

// Barrier across the ranks, one warp: lane q stores this rank's new epoch of the channel into rank q's block (release,
// system scope: everything earlier kernels of this stream wrote -- into peer memory too -- is visible to whoever
// acquires the flag) and then waits until rank q's epoch has arrived in its own block.  Epochs only grow and every
// (writer, reader) slot has one writer, so nothing is ever reset, and a rank that is a whole barrier ahead is harmless.
// The epoch counter lives in device memory: the kernel takes no per-call argument and can be replayed from a CUDA graph.
__global__ void shard_barrier_kernel(PeerFlags peers, int rank, int world, int channel) {
    unsigned long long* mine = peers.f[rank];
    unsigned long long* counter = mine + kBarrierChannels * MAX_PEERS + channel;
    const int q = threadIdx.x;
    unsigned long long epoch = 0;
    if (q == 0) {
        epoch = *counter + 1;
        *counter = epoch;
    }
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    if (q < world) {
        __threadfence_system();
        unsigned long long* theirs = peers.f[q] + channel * MAX_PEERS + rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
        const unsigned long long* slot = mine + channel * MAX_PEERS + q;
        unsigned long long seen = 0;
        unsigned long long t0 = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (true) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(slot) : "memory");
            if (seen >= epoch) break;
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 60ull * 1000000000ull) __trap();  // a peer never arrived: fail instead of hanging the GPU
            __nanosleep(64);
        }
    }
}

}  // namespace
}  // namespace clibd

using namespace clibd;

extern "C" {

int clibd_shard_push_rows(const void* const x_local[3], int dtype, const int64_t* labels_local, int64_t n, int64_t d,
                          int rank, int world, void* const peer_x[], float* const peer_inv[],
                          int64_t* const peer_labels[], clibd_stream_t stream) {
    NvtxRange nvtx_range("clibd_shard_push_rows");
    CLIBD_REQUIRE(x_local && peer_x && peer_inv && peer_labels, "null pointer");  // labels_local may be null: rows only
    CLIBD_REQUIRE(n > 0 && d > 0 && world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world, "bad shape or rank");
    CLIBD_REQUIRE(dtype == DT_F32 || dtype == DT_BF16 || dtype == DT_F16, "dtype must be 0, 1 or 2");
    const size_t esize = dtype == DT_F32 ? 4 : 2;
    PeerRows peers;
    bool vec = (d * esize) % 16 == 0;
    for (int q = 0; q < MAX_PEERS; ++q) {
        peers.labels[q] = q < world ? peer_labels[q] : nullptr;
        if (q < world && labels_local != nullptr) CLIBD_REQUIRE(peer_labels[q] != nullptr, "null peer label buffer");
        for (int m = 0; m < 3; ++m) {
            const bool on = q < world && x_local[m] != nullptr;
            peers.x[q * 3 + m] = on ? peer_x[q * 3 + m] : nullptr;
            peers.inv[q * 3 + m] = on ? peer_inv[q * 3 + m] : nullptr;
            if (on) {
                CLIBD_REQUIRE(peer_x[q * 3 + m] && peer_inv[q * 3 + m], "null peer buffer of a present modality");
                vec = vec && aligned16(peer_x[q * 3 + m]);
            }
        }
    }
    for (int m = 0; m < 3; ++m)
        if (x_local[m]) vec = vec && aligned16(x_local[m]);
    const int64_t blocks = ceil_div(3 * n * 32, kThreads);
#define CLIBD_PUSH(T)                                                                                                    \
    do {                                                                                                                 \
        if (vec)                                                                                                         \
            shard_push_rows_kernel<T, true><<<blocks, kThreads, 0, stream>>>(                                            \
                static_cast<const T*>(x_local[0]), static_cast<const T*>(x_local[1]), static_cast<const T*>(x_local[2]), \
                labels_local, n, d, rank, world, peers);                                                                 \
        else                                                                                                             \
            shard_push_rows_kernel<T, false><<<blocks, kThreads, 0, stream>>>(                                           \
                static_cast<const T*>(x_local[0]), static_cast<const T*>(x_local[1]), static_cast<const T*>(x_local[2]), \
                labels_local, n, d, rank, world, peers);                                                                 \
    } while (0)
    if (dtype == DT_F32) CLIBD_PUSH(float);
    else if (dtype == DT_BF16) CLIBD_PUSH(__nv_bfloat16);
    else CLIBD_PUSH(__half);
#undef CLIBD_PUSH
    CLIBD_KERNEL_CHECK();
    return 0;
}

int clibd_shard_push_stats(const float* stats, const double* pos, int64_t N, int64_t row0, int64_t n, int rank,
                           int world, float* const peer_stats[], float* const peer_colslots[],
                           double* const peer_posslots[], clibd_stream_t stream) {
    NvtxRange nvtx_range("clibd_shard_push_stats");
    CLIBD_REQUIRE(stats && pos && peer_stats && peer_colslots && peer_posslots, "null pointer");
    CLIBD_REQUIRE(N > 0 && n > 0 && row0 >= 0 && row0 + n <= N && world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world,
                  "bad shape or rank");
    PeerStats peers;
    for (int q = 0; q < MAX_PEERS; ++q) {
        peers.stats[q] = q < world ? peer_stats[q] : nullptr;
        peers.colslots[q] = q < world ? peer_colslots[q] : nullptr;
        peers.posslots[q] = q < world ? peer_posslots[q] : nullptr;
        if (q < world) CLIBD_REQUIRE(peer_stats[q] && peer_colslots[q] && peer_posslots[q], "null peer buffer");
    }
    const int64_t len = 3 * N > 6 * n ? 3 * N : 6 * n;
    dim3 grid(static_cast<unsigned>(ceil_div(len, kThreads)), static_cast<unsigned>(world));
    shard_push_stats_kernel<<<grid, kThreads, 0, stream>>>(stats, pos, N, row0, n, rank, peers);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int clibd_shard_reduce_stats(const float* colslots, const double* posslots, int64_t N, int world, float* stats,
                             double* pos, clibd_stream_t stream) {
    CLIBD_REQUIRE(colslots && posslots && stats && pos && N > 0 && world >= 1 && world <= MAX_PEERS, "bad arguments");
    shard_reduce_stats_kernel<<<static_cast<unsigned>(ceil_div(3 * N, kThreads)), kThreads, 0, stream>>>(colslots, posslots, N,
                                                                                                        world, stats, pos);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int clibd_shard_push_floats(const float* src, int64_t count, int rank, int world, float* const peer_slots[],
                            clibd_stream_t stream) {
    CLIBD_REQUIRE(src && peer_slots && count > 0 && world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world,
                  "bad arguments");
    PeerFloats peers;
    for (int q = 0; q < MAX_PEERS; ++q) {
        peers.p[q] = q < world ? peer_slots[q] : nullptr;
        if (q < world) CLIBD_REQUIRE(peer_slots[q] != nullptr, "null peer slot array");
    }
    dim3 grid(static_cast<unsigned>(ceil_div(count, kThreads)), static_cast<unsigned>(world));
    shard_push_floats_kernel<<<grid, kThreads, 0, stream>>>(src, count, rank, peers);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int clibd_shard_barrier(uint64_t* const peer_flags[], int rank, int world, int channel, clibd_stream_t stream) {
    CLIBD_REQUIRE(peer_flags && world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world && channel >= 0 &&
                      channel < kBarrierChannels,
                  "bad arguments");
    PeerFlags peers;
    for (int q = 0; q < MAX_PEERS; ++q) {
        peers.f[q] = q < world ? reinterpret_cast<unsigned long long*>(peer_flags[q]) : nullptr;
        if (q < world) CLIBD_REQUIRE(peer_flags[q] != nullptr, "null peer flag block");
    }
    shard_barrier_kernel<<<1, 32, 0, stream>>>(peers, rank, world, channel);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int64_t clibd_shard_barrier_bytes(void) {
    return sizeof(unsigned long long) * (kBarrierChannels * MAX_PEERS + kBarrierChannels);
}

}  // extern "C"
