// Host-side internals shared by the translation units of libb200clip.
#pragma once
#include <utility>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/b200clip.h"

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Tensor-map cache: a training / inference loop encodes the same few hundred (pointer, extents, pitch,
// box) descriptors every step -- weights never move, activations come back from the caching allocator at the same
// addresses -- so the driver's encoder (~1 us each, 2-4 per launch) is consulted only on a miss.  Direct mapped.
struct TmapCacheEntry {
    uint64_t key[4];   // ptr, dim0, dim1, pitch << 20 | box0 << 10 | box1   (all zero = empty)
    CUtensorMap map;
};
constexpr int kTmapCacheSize = 2048;

struct b200clip_ctx {
    int device;
    int num_sms;
    int max_quads;  // co-resident 4-CTA clusters of gemm_quad_bf16_kernel (0: kernel unavailable)
    PFN_encodeTiled encode_tiled;
    TmapCacheEntry* tmap_cache;   // kTmapCacheSize entries, owned
    uint64_t tmap_hits, tmap_misses;
};

namespace b200 {
void count_launch();
}

namespace b200 {

void set_error(const char* fmt, ...);

#define B200_CHECK_ARG(cond, ...)              \
    do {                                       \
        if (!(cond)) {                         \
            b200::set_error(__VA_ARGS__);      \
            return B200CLIP_ERR_ARG;           \
        }                                      \
    } while (0)

#define B200_CHECK_CUDA(expr)                                                                  \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            b200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return B200CLIP_ERR_CUDA;                                                          \
        }                                                                                      \
    } while (0)

#define B200_CHECK_CTX(ctx)                                       \
    do {                                                          \
        if ((ctx) == nullptr) {                                   \
            b200::set_error("null context");                      \
            return B200CLIP_ERR_ARG;                              \
        }                                                         \
    } while (0)

// every kernel launch of the library goes through this macro: it counts the launch (bench.py's
// `gpu_launches`) and surfaces launch-configuration errors
#define B200_LAUNCH_CHECK()                  \
    do {                                     \
        b200::count_launch();                \
        B200_CHECK_CUDA(cudaGetLastError()); \
    } while (0)

// Kernel launch with programmatic dependent launch enabled (see common.cuh: the kernel must call
// griddep_wait() before its first global-memory access).  B200CLIP_PDL=0 falls back to plain
// stream-ordered launches.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// 2-D bf16 tensor map, 128-byte swizzle.  dim0 = innermost extent (elements), dim1 = rows,
// pitch in elements; box = (box0 <= 64, box1 <= 256).
int make_tmap_bf16_2d(b200clip_ctx* ctx, CUtensorMap* out, const void* ptr, uint64_t dim0, uint64_t dim1,
                      uint64_t pitch_elems, uint32_t box0, uint32_t box1);

// the same with a 64-byte swizzle (box0 <= 32): the 32 x 32 bf16 staging tiles of the GEMM's TMA-store epilogue
int make_tmap_bf16_2d_sw(b200clip_ctx* ctx, CUtensorMap* out, const void* ptr, uint64_t dim0, uint64_t dim1,
                         uint64_t pitch_elems, uint32_t box0, uint32_t box1, int swizzle_bytes);

// 2-D uint8 tensor map, 32-byte swizzle, box0 <= 32 bytes (the 8-bit QuickGELU' codes of the c_fc forward epilogue)
int make_tmap_u8_2d_sw32(b200clip_ctx* ctx, CUtensorMap* out, const void* ptr, uint64_t dim0, uint64_t dim1,
                         uint64_t pitch_bytes, uint32_t box0, uint32_t box1);

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace b200
