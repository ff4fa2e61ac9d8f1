// Context, error reporting and tensor-map construction for libb200clip.
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "internal.h"

namespace b200 {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("B200CLIP_PDL");
        return !(e && atoi(e) == 0);
    }();
    return on;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int make_tmap_bf16_2d(b200clip_ctx* ctx, CUtensorMap* out, const void* ptr, uint64_t dim0, uint64_t dim1,
                      uint64_t pitch_elems, uint32_t box0, uint32_t box1) {
    return make_tmap_bf16_2d_sw(ctx, out, ptr, dim0, dim1, pitch_elems, box0, box1, 128);
}

int make_tmap_bf16_2d_sw(b200clip_ctx* ctx, CUtensorMap* out, const void* ptr, uint64_t dim0, uint64_t dim1,
                         uint64_t pitch_elems, uint32_t box0, uint32_t box1, int swizzle_bytes) {
    if (swizzle_bytes != 128 && swizzle_bytes != 64) {
        set_error("tensor map: swizzle %d not supported", swizzle_bytes);
        return B200CLIP_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (pitch_elems * 2) % 16 != 0) {
        set_error("tensor map: pointer %p / pitch %llu elements not 16-byte aligned", ptr,
                  (unsigned long long)pitch_elems);
        return B200CLIP_ERR_ARG;
    }
    if (box0 * 2 > static_cast<uint32_t>(swizzle_bytes) || box1 > 256 || box0 == 0 || box1 == 0) {
        set_error("tensor map: bad box %u x %u", box0, box1);
        return B200CLIP_ERR_ARG;
    }
    // cache lookup (one context is driven by one host thread at a time; entries are written whole before use)
    const uint64_t key[4] = {reinterpret_cast<uint64_t>(ptr), dim0, dim1,
                             (pitch_elems << 20) | (static_cast<uint64_t>(box0) << 10) | box1 |
                                 (static_cast<uint64_t>(swizzle_bytes == 64) << 63)};
    TmapCacheEntry* slot = nullptr;
    if (ctx->tmap_cache != nullptr) {
        uint64_t h = key[0] * 0x9E3779B97F4A7C15ull ^ key[1] * 0xC2B2AE3D27D4EB4Full ^ key[2] * 0x165667B19E3779F9ull ^ key[3];
        h ^= h >> 29;
        slot = &ctx->tmap_cache[h % kTmapCacheSize];
        if (slot->key[0] == key[0] && slot->key[1] == key[1] && slot->key[2] == key[2] && slot->key[3] == key[3]) {
            memcpy(out, &slot->map, sizeof(CUtensorMap));
            ++ctx->tmap_hits;
            return B200CLIP_OK;
        }
        ++ctx->tmap_misses;
    }
    cuuint64_t gdim[2] = {dim0, dim1};
    cuuint64_t gstride[1] = {pitch_elems * 2};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): dims %llu x %llu pitch %llu box %u x %u", (int)r,
                  (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)pitch_elems, box0, box1);
        return B200CLIP_ERR_CUDA;
    }
    if (slot != nullptr) {
        memcpy(&slot->map, out, sizeof(CUtensorMap));
        memcpy(slot->key, key, sizeof(key));
    }
    return B200CLIP_OK;
}

int make_tmap_u8_2d_sw32(b200clip_ctx* ctx, CUtensorMap* out, const void* ptr, uint64_t dim0, uint64_t dim1,
                         uint64_t pitch_bytes, uint32_t box0, uint32_t box1) {
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || pitch_bytes % 16 != 0 || box0 == 0 || box0 > 32 || box1 == 0 ||
        box1 > 256) {
        set_error("u8 tensor map: pointer %p / pitch %llu / box %u x %u", ptr, (unsigned long long)pitch_bytes, box0, box1);
        return B200CLIP_ERR_ARG;
    }
    const uint64_t key[4] = {reinterpret_cast<uint64_t>(ptr), dim0, dim1,
                             (pitch_bytes << 20) | (static_cast<uint64_t>(box0) << 10) | box1 | (1ull << 62)};
    TmapCacheEntry* slot = nullptr;
    if (ctx->tmap_cache != nullptr) {
        uint64_t h = key[0] * 0x9E3779B97F4A7C15ull ^ key[1] * 0xC2B2AE3D27D4EB4Full ^ key[2] * 0x165667B19E3779F9ull ^ key[3];
        h ^= h >> 29;
        slot = &ctx->tmap_cache[h % kTmapCacheSize];
        if (slot->key[0] == key[0] && slot->key[1] == key[1] && slot->key[2] == key[2] && slot->key[3] == key[3]) {
            memcpy(out, &slot->map, sizeof(CUtensorMap));
            ++ctx->tmap_hits;
            return B200CLIP_OK;
        }
        ++ctx->tmap_misses;
    }
    cuuint64_t gdim[2] = {dim0, dim1};
    cuuint64_t gstride[1] = {pitch_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(u8) failed (%d): dims %llu x %llu pitch %llu box %u x %u", (int)r,
                  (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)pitch_bytes, box0, box1);
        return B200CLIP_ERR_CUDA;
    }
    if (slot != nullptr) {
        memcpy(&slot->map, out, sizeof(CUtensorMap));
        memcpy(slot->key, key, sizeof(key));
    }
    return B200CLIP_OK;
}

int init_gemm(b200clip_ctx* ctx);
int init_attention(b200clip_ctx* ctx);

}  // namespace b200

extern "C" {

int b200clip_abi_version(void) { return B200CLIP_ABI_VERSION; }

uint64_t b200clip_launch_count(void) { return b200::g_launches.load(std::memory_order_relaxed); }

const char* b200clip_last_error(void) { return b200::g_err; }

int b200clip_ctx_create(b200clip_ctx** out, int device) {
    if (out == nullptr) {
        b200::set_error("ctx_create: null out");
        return B200CLIP_ERR_ARG;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        b200::set_error("no CUDA device available (%s); libb200clip has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        (void)cudaGetLastError();
        return B200CLIP_ERR_DEVICE;
    }
    if (device < 0 || device >= count) {
        b200::set_error("device %d out of range (count %d)", device, count);
        return B200CLIP_ERR_DEVICE;
    }
    cudaDeviceProp prop;
    B200_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        b200::set_error("device %d is sm_%d%d; libb200clip is built for sm_100a only", device, prop.major, prop.minor);
        return B200CLIP_ERR_DEVICE;
    }
    B200_CHECK_CUDA(cudaSetDevice(device));
    B200_CHECK_CUDA(cudaFree(0));  // make sure the primary context exists
    b200clip_ctx* ctx = new b200clip_ctx();
    ctx->tmap_cache = nullptr;
    ctx->tmap_hits = ctx->tmap_misses = 0;
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
        b200::set_error("cannot resolve cuTensorMapEncodeTiled from the driver");
        delete ctx;
        return B200CLIP_ERR_CUDA;
    }
    ctx->encode_tiled = reinterpret_cast<PFN_encodeTiled>(fn);
    ctx->tmap_cache = (getenv("B200CLIP_TMAP_CACHE") && atoi(getenv("B200CLIP_TMAP_CACHE")) == 0)
                          ? nullptr : new TmapCacheEntry[kTmapCacheSize]();
    int rc = b200::init_gemm(ctx);
    if (rc == 0) rc = b200::init_attention(ctx);
    if (rc != 0) {
        delete[] ctx->tmap_cache;
        delete ctx;
        return rc;
    }
    *out = ctx;
    return B200CLIP_OK;
}

int b200clip_ctx_destroy(b200clip_ctx* ctx) {
    if (ctx != nullptr) delete[] ctx->tmap_cache;
    delete ctx;
    return B200CLIP_OK;
}

}  // extern "C"
