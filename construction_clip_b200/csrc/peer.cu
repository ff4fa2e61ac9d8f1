// Peer-memory exchange over NVLink / NVSwitch: an all-gather written as ONE kernel per rank that publishes this
// rank's block, raises flags in every peer's memory, waits for the peers' flags and pulls their blocks with plain
// loads through the NVLink fabric -- straight into the layout the consumer wants (up to three independent segments
// per call, e.g. image features -> img_all[Bg, E] and text features -> txt_all[Bg, E]).
//
// It replaces the NCCL collectives that sit on the critical path between the two towers and the fused InfoNCE
// loss of the data-parallel step ("can instead be done ... with NVLink peer loads"): the feature
// gather (Bl x 2E fp32 per rank) and the gather of the row log-sum-exps + the three loss statistics.  These messages
// are tiny (512 KB / 1 KB per rank at global batch 1024 on 8 GPUs), so what matters is latency: a ring all-gather pays
// one hop per rank, the all-reduce form one reduction tree; here every rank reads every peer directly, all peers in
// parallel, after ONE flag round trip.
//
// Every rank owns a "symmetric" buffer (cudaMalloc + CUDA IPC, mapped into every peer process):
//     [0, 256)                        control words of the owner: epoch, finished-CTA counter, timeout flag
//     [256, 256 + world*kMaxCtas*4)   flags[src rank][cta] = last epoch that (rank, cta) has published   (written by peers)
//     two data slots of `slot_bytes`  (slot = epoch & 1)
// Protocol of call number e (the epoch lives in device memory and is advanced by the kernel itself, so the launch can
// be captured in a CUDA graph and replayed): CTA c copies its share of the local block into slot e & 1 (and into the
// local part of the destination), fences at system scope, stores e into flags[rank][c] of every peer
// (st.release.sys), spins until flags[p][c] >= e for every peer p in its own buffer (ld.acquire.sys) and then copies
// share c of every peer's slot.  Flags are per CTA because share c of a slot is written by the owner's CTA c: no grid
// barrier is needed anywhere.  Two slots make the WAR hazard impossible: a rank can reach call e + 2 (the next use of
// slot e & 1) only after every peer has published e + 1, i.e. after every peer finished pulling call e.
// A spin that lasts longer than kTimeoutNs sets the owner's timeout flag and gives up (garbage instead of a hang);
// the host checks the flag after its self-test and whenever it reads the loss.
#include <cstring>

#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr int kPeerMaxCtas = 32;
constexpr int kPeerThreads = 256;
constexpr int kPeerMaxWorld = 16;
constexpr int kPeerCtrlBytes = 256;
constexpr unsigned long long kTimeoutNs = 4000000000ull;  // 4 s

struct PeerCtrl {
    unsigned long long epoch;
    unsigned int done;
    unsigned int timeout;
};

struct PeerArgs {
    uint8_t* const* bufs;  // device array [world]: every rank's symmetric buffer as mapped in THIS process
    int world, rank, nseg, unit;  // unit: 16 (all segments 16-byte aligned / sized) or 4
    int64_t slot_bytes;
    const uint8_t* src[3];
    uint8_t* dst[3];
    int64_t bytes[3];  // per rank
};

__host__ __device__ inline int64_t peer_flags_bytes(int world) {
    return (static_cast<int64_t>(world) * kPeerMaxCtas * 4 + 255) / 256 * 256;
}

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// copies units [u0, u1) of the concatenation of the segments; T = uint4 or uint32_t
template <typename T, typename F>
__device__ __forceinline__ void for_units(const PeerArgs& a, int64_t u0, int64_t u1, F&& f) {
    int64_t base = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        if (s >= a.nseg) break;
        const int64_t n = a.bytes[s] / static_cast<int64_t>(sizeof(T));
        const int64_t lo = u0 > base ? u0 : base, hi = u1 < base + n ? u1 : base + n;
        for (int64_t u = lo + threadIdx.x; u < hi; u += blockDim.x) f(s, u - base, u);
        base += n;
    }
}

template <typename T>
__global__ void __launch_bounds__(kPeerThreads)
peer_allgather_kernel(const PeerArgs a) {
    uint8_t* mine = a.bufs[a.rank];
    PeerCtrl* ctrl = reinterpret_cast<PeerCtrl*>(mine);
    const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(&ctrl->epoch) + 1ull;
    const int64_t data0 = kPeerCtrlBytes + peer_flags_bytes(a.world) + static_cast<int64_t>(e & 1ull) * a.slot_bytes;
    int64_t total = 0;
    for (int s = 0; s < a.nseg; ++s) total += a.bytes[s] / static_cast<int64_t>(sizeof(T));
    const int64_t u0 = total * blockIdx.x / gridDim.x, u1 = total * (blockIdx.x + 1) / gridDim.x;

    // 1. publish this rank's share: slot (read by the peers) and the local block of the destination
    T* slot = reinterpret_cast<T*>(mine + data0);
    for_units<T>(a, u0, u1, [&](int s, int64_t i, int64_t u) {
        const T v = reinterpret_cast<const T*>(a.src[s])[i];
        slot[u] = v;
        reinterpret_cast<T*>(a.dst[s] + static_cast<int64_t>(a.rank) * a.bytes[s])[i] = v;
    });
    __threadfence_system();
    __syncthreads();
    // 2. + 3. one thread per peer: raise our flag in its buffer, then wait for its flag in ours
    if (threadIdx.x < a.world && static_cast<int>(threadIdx.x) != a.rank) {
        const int peer = threadIdx.x;
        unsigned int* theirs = reinterpret_cast<unsigned int*>(a.bufs[peer] + kPeerCtrlBytes) + a.rank * kPeerMaxCtas + blockIdx.x;
        st_release_sys(theirs, static_cast<unsigned int>(e));
        const unsigned int* ours = reinterpret_cast<const unsigned int*>(mine + kPeerCtrlBytes) + peer * kPeerMaxCtas + blockIdx.x;
        const unsigned long long t0 = globaltimer_ns();
        while (static_cast<int>(ld_acquire_sys(ours) - static_cast<unsigned int>(e)) < 0) {
            if (globaltimer_ns() - t0 > kTimeoutNs) {
                atomicExch(&ctrl->timeout, 1u);
                break;
            }
        }
    }
    __syncthreads();
    // 4. pull the peers' shares (nearest neighbour first, so that the ranks do not all hit the same peer at once)
    for (int k = 1; k < a.world; ++k) {
        const int peer = (a.rank + k) % a.world;
        const T* theirs = reinterpret_cast<const T*>(a.bufs[peer] + data0);
        for_units<T>(a, u0, u1, [&](int s, int64_t i, int64_t u) {
            reinterpret_cast<T*>(a.dst[s] + static_cast<int64_t>(peer) * a.bytes[s])[i] = __ldcv(theirs + u);
        });
    }
    // 5. the last CTA to finish advances the epoch (every CTA read it at the top, and the next call is stream-ordered)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&ctrl->done, 1u) == gridDim.x - 1) {
            ctrl->done = 0u;
            *reinterpret_cast<volatile unsigned long long*>(&ctrl->epoch) = e;
            __threadfence();
        }
    }
}

}  // namespace b200

using namespace b200;

extern "C" int64_t b200clip_peer_buffer_bytes(int world, int64_t slot_bytes) {
    if (world < 1 || world > kPeerMaxWorld || slot_bytes < 0) return -1;
    const int64_t slot = (slot_bytes + 255) / 256 * 256;
    return kPeerCtrlBytes + peer_flags_bytes(world) + 2 * slot;
}

extern "C" int b200clip_peer_alloc(b200clip_ctx* ctx, int64_t bytes, void** ptr, void* handle64) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(ptr && handle64 && bytes > 0, "peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    void* p = nullptr;
    B200_CHECK_CUDA(cudaMalloc(&p, static_cast<size_t>(bytes)));
    cudaError_t e = cudaMemset(p, 0, static_cast<size_t>(bytes));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), p);
    if (e != cudaSuccess) {
        set_error("peer_alloc: %s", cudaGetErrorString(e));
        (void)cudaFree(p);
        (void)cudaGetLastError();
        return B200CLIP_ERR_CUDA;
    }
    *ptr = p;
    return B200CLIP_OK;
}

extern "C" int b200clip_peer_open(b200clip_ctx* ctx, const void* handle64, void** ptr) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(ptr && handle64, "peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        set_error("peer_open: cudaIpcOpenMemHandle -> %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return B200CLIP_ERR_CUDA;
    }
    *ptr = p;
    return B200CLIP_OK;
}

extern "C" int b200clip_peer_close(b200clip_ctx* ctx, void* ptr) {
    B200_CHECK_CTX(ctx);
    if (ptr != nullptr) B200_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
    return B200CLIP_OK;
}

extern "C" int b200clip_peer_free(b200clip_ctx* ctx, void* ptr) {
    B200_CHECK_CTX(ctx);
    if (ptr != nullptr) B200_CHECK_CUDA(cudaFree(ptr));
    return B200CLIP_OK;
}

extern "C" int b200clip_peer_allgather(b200clip_ctx* ctx, const void* const* bufs_dev, int world, int rank,
                                       int64_t slot_bytes, int nseg, const void* const* src, void* const* dst,
                                       const int64_t* bytes, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(bufs_dev && src && dst && bytes, "peer_allgather: null argument");
    B200_CHECK_ARG(world >= 2 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "peer_allgather: bad world / rank");
    B200_CHECK_ARG(nseg >= 1 && nseg <= 3, "peer_allgather: 1..3 segments");
    PeerArgs a{};
    a.bufs = reinterpret_cast<uint8_t* const*>(const_cast<void* const*>(reinterpret_cast<const void* const*>(bufs_dev)));
    a.world = world;
    a.rank = rank;
    a.nseg = nseg;
    a.slot_bytes = (slot_bytes + 255) / 256 * 256;
    a.unit = 16;
    int64_t total = 0;
    for (int s = 0; s < nseg; ++s) {
        B200_CHECK_ARG(src[s] && dst[s] && bytes[s] > 0 && bytes[s] % 4 == 0, "peer_allgather: segment %d empty / not a multiple of 4 bytes", s);
        B200_CHECK_ARG((reinterpret_cast<uintptr_t>(src[s]) & 3) == 0 && (reinterpret_cast<uintptr_t>(dst[s]) & 3) == 0,
                       "peer_allgather: segment %d misaligned", s);
        if (bytes[s] % 16 || (reinterpret_cast<uintptr_t>(src[s]) & 15) || (reinterpret_cast<uintptr_t>(dst[s]) & 15)) a.unit = 4;
        a.src[s] = static_cast<const uint8_t*>(src[s]);
        a.dst[s] = static_cast<uint8_t*>(dst[s]);
        a.bytes[s] = bytes[s];
        total += bytes[s];
    }
    B200_CHECK_ARG(total <= slot_bytes, "peer_allgather: %lld bytes per rank exceed the slot (%lld)", (long long)total,
                   (long long)slot_bytes);
    // one CTA per 16 KB of this rank's block, at most kPeerMaxCtas (the same on every rank: the sizes are)
    int ctas = static_cast<int>((total + 16383) / 16384);
    ctas = ctas < 1 ? 1 : (ctas > kPeerMaxCtas ? kPeerMaxCtas : ctas);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (a.unit == 16)
        peer_allgather_kernel<uint4><<<ctas, kPeerThreads, 0, st>>>(a);
    else
        peer_allgather_kernel<uint32_t><<<ctas, kPeerThreads, 0, st>>>(a);
    B200_LAUNCH_CHECK();
    return B200CLIP_OK;
}

// control words of a rank's own buffer (host readable after a stream sync): out[0] = epoch, out[1] = timeout flag
extern "C" int b200clip_peer_status(b200clip_ctx* ctx, const void* own_buf, int64_t* out2) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(own_buf && out2, "peer_status: null argument");
    PeerCtrl c;
    B200_CHECK_CUDA(cudaMemcpy(&c, own_buf, sizeof(c), cudaMemcpyDeviceToHost));
    out2[0] = static_cast<int64_t>(c.epoch);
    out2[1] = static_cast<int64_t>(c.timeout);
    return B200CLIP_OK;
}
