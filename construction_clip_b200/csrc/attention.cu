// Fused multi-head attention (head_dim 64, sequence <= 128 = one tile) on tcgen05 tensor cores.
// Replaces the core of nn.MultiheadAttention(need_weights=False, attn_mask=causal|None) inside
// clip.model.ResidualAttentionBlock.attention: softmax(q k^T / sqrt(64) + mask) v
// (vision tower: S = 50, full; text tower: S = 77, causal with upstream's mask that does NOT
// mask padding).  Forward and backward; the backward recomputes the probabilities from q,k.
//
// One 128-thread CTA per (batch, head) work item, persistent over the work list, several CTAs
// per SM.  Q/K/V(/dO) tiles arrive by TMA straight out of the packed in_proj output
// [B*S, 3*H*64] into 128-byte-swizzled shared memory (one 128-byte row per token = exactly one
// swizzle row).  Thread r owns score row r (TMEM lane r):
//   fwd:  S = Q K^T (UMMA 128 x NPAD x 64)  -> softmax in registers -> P (bf16) to swizzled smem
//         O = P V   (A = P K-major, B = V MN-major) -> scale by 1/rowsum -> global
//   bwd:  S = Q K^T, dP = dO V^T -> P, dS = P o (dP - rowsum(P o dP)) / 8 to smem ->
//         dV = P^T dO, dK = dS^T Q (A read MN-major from the same P / dS tiles), dQ = dS K
#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr int kAttnThreads = 128;
constexpr int kTile = 128 * 128;  // bytes of one [128 rows x 64 bf16] swizzled tile (16 KB)
constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
    const __nv_bfloat16* dout;  // bwd only (unused; dO arrives by TMA)
    __nv_bfloat16* out;         // fwd: [B*S, H*64] ; bwd: dqkv [B*S, 3*H*64]
    int B, S, H, causal;
    int npad;  // S rounded up to a multiple of 16
};

// zero a shared-memory region cooperatively
__device__ __forceinline__ void zero_smem(uint8_t* p, int bytes) {
    uint4* q = reinterpret_cast<uint4*>(p);
    for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) q[i] = make_uint4(0u, 0u, 0u, 0u);
}

// store 8 consecutive bf16 (one 16-byte chunk) of row r, columns [c8*8, c8*8+8) into a K-major
// SW128 tile set (tiles of 64 columns, 16 KB each)
__device__ __forceinline__ void store_p_chunk(uint8_t* tile_base, int r, int c8, uint4 v) {
    const int atom = c8 >> 3, cc = c8 & 7;
    *reinterpret_cast<uint4*>(tile_base + atom * kTile + r * 128 + ((cc ^ (r & 7)) << 4)) = v;
}

__device__ __forceinline__ bool masked(int r, int c, int S, int causal) { return c >= S || (causal && c > r); }

// ------------------------------------------------------------------------------------------------
template <bool BIG>  // BIG: S > 64 (P needs two 64-column tiles)
__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int kPBytes = BIG ? 2 * kTile : kTile;
    uint8_t* sP = smem;  // aliases Q (Q is dead once S = Q K^T has completed)
    uint8_t* sQ = smem;
    uint8_t* sK = smem + kPBytes;
    uint8_t* sV = sK + kTile;
    constexpr int kTmemCols = 128;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = threadIdx.x;  // score row owned by this thread

    zero_smem(smem, kPBytes + 2 * kTile);
    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_qkv);
        mbar_init(&bar_load, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, kTmemCols);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);

    const int S = p.S, H = p.H, npad = p.npad;
    const int d = H * 64;
    const uint32_t load_bytes = 3u * S * 128u;
    const uint32_t idesc_s = make_idesc_bf16(128, npad, 0, 0);
    const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
    const float sc = 0.125f * kLog2e;
    const int num_work = p.B * H;
    uint32_t it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
        const int b = w / H, h = w % H;
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(&bar_load, load_bytes);
            tma_load_2d(sQ, &tm_qkv, &bar_load, h * 64, b * S);
            tma_load_2d(sK, &tm_qkv, &bar_load, d + h * 64, b * S);
            tma_load_2d(sV, &tm_qkv, &bar_load, 2 * d + h * 64, b * S);
            mbar_wait(&bar_load, it & 1u);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem, make_smem_desc_sw128(smem_u32(sQ) + k * 32, 16, 1024),
                          make_smem_desc_sw128(smem_u32(sK) + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
            umma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, 0u);
        __syncwarp();
        tc_fence_after();

        // ---- softmax over row r: pass 1 max, pass 2 exp / sum / write P ----
        float mx = -INFINITY;
        for (int c0 = 0; c0 < npad; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x16(trow + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (!masked(r, c0 + j, S, p.causal)) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
        float sum = 0.f;
        for (int c0 = 0; c0 < npad; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x16(trow + c0, v);
            tmem_ld_wait();
            float e[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                e[j] = masked(r, c0 + j, S, p.causal) ? 0.f : exp2f((__uint_as_float(v[j]) - mx) * sc);
                sum += e[j];
            }
            store_p_chunk(sP, r, (c0 >> 3), make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]),
                                                        pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7])));
            store_p_chunk(sP, r, (c0 >> 3) + 1, make_uint4(pack_bf16(e[8], e[9]), pack_bf16(e[10], e[11]),
                                                            pack_bf16(e[12], e[13]), pack_bf16(e[14], e[15])));
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc_fence_after();
            const int ksteps = npad >> 4;
            for (int k = 0; k < ksteps; ++k)
                umma_bf16(tmem,
                          make_smem_desc_sw128(smem_u32(sP) + (k >> 2) * kTile + (k & 3) * 32, 16, 1024),
                          make_smem_desc_sw128(smem_u32(sV) + k * 2048, 8192, 1024), idesc_o, k > 0 ? 1u : 0u);
            umma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, 1u);
        __syncwarp();
        tc_fence_after();
        {
            const float inv = 1.0f / sum;
            __nv_bfloat16* dst = p.out + (static_cast<int64_t>(b) * S + r) * d + h * 64;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(trow + c0, v);
                tmem_ld_wait();
                if (r < S) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 o;
                        o.x = pack_bf16(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
                        o.y = pack_bf16(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
                        o.z = pack_bf16(__uint_as_float(v[j + 4]) * inv, __uint_as_float(v[j + 5]) * inv);
                        o.w = pack_bf16(__uint_as_float(v[j + 6]) * inv, __uint_as_float(v[j + 7]) * inv);
                        *reinterpret_cast<uint4*>(dst + c0 + j) = o;
                    }
                }
            }
        }
        // P (generic-proxy writes) aliases Q (next TMA write): order them across proxies
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
template <bool BIG>
__global__ void __launch_bounds__(kAttnThreads)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int kPBytes = BIG ? 2 * kTile : kTile;
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kTile;
    uint8_t* sV = sK + kTile;
    uint8_t* sdO = sV + kTile;
    uint8_t* sP = sdO + kTile;
    uint8_t* sdS = sP + kPBytes;
    constexpr int kTmemCols = 256;
    // TMEM columns: [0,128) S, later dV [0,64) + dK [64,128);  [128,256) dP, later dQ [128,192)

    const int warp = threadIdx.x >> 5;
    const int r = threadIdx.x;

    zero_smem(smem, 4 * kTile + 2 * kPBytes);
    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_qkv);
        prefetch_tmap(&tm_do);
        mbar_init(&bar_load, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, kTmemCols);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);

    const int S = p.S, H = p.H, npad = p.npad;
    const int d = H * 64;
    const uint32_t load_bytes = 4u * S * 128u;
    const uint32_t idesc_s = make_idesc_bf16(128, npad, 0, 0);    // S = Q K^T, dP = dO V^T
    const uint32_t idesc_tn = make_idesc_bf16(128, 64, 1, 1);     // dV = P^T dO, dK = dS^T Q
    const uint32_t idesc_nn = make_idesc_bf16(128, 64, 0, 1);     // dQ = dS K
    const float sc = 0.125f * kLog2e;
    const int num_work = p.B * H;
    uint32_t it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
        const int b = w / H, h = w % H;
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(&bar_load, load_bytes);
            tma_load_2d(sQ, &tm_qkv, &bar_load, h * 64, b * S);
            tma_load_2d(sK, &tm_qkv, &bar_load, d + h * 64, b * S);
            tma_load_2d(sV, &tm_qkv, &bar_load, 2 * d + h * 64, b * S);
            tma_load_2d(sdO, &tm_do, &bar_load, h * 64, b * S);
            mbar_wait(&bar_load, it & 1u);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem, make_smem_desc_sw128(smem_u32(sQ) + k * 32, 16, 1024),
                          make_smem_desc_sw128(smem_u32(sK) + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem + 128, make_smem_desc_sw128(smem_u32(sdO) + k * 32, 16, 1024),
                          make_smem_desc_sw128(smem_u32(sV) + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
            umma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, 0u);
        __syncwarp();
        tc_fence_after();

        // pass A: row max of the masked scores
        float mx = -INFINITY;
        for (int c0 = 0; c0 < npad; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x16(trow + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (!masked(r, c0 + j, S, p.causal)) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
        // pass B: sum exp and sum exp * dP
        float sum = 0.f, dsum = 0.f;
        for (int c0 = 0; c0 < npad; c0 += 16) {
            uint32_t v[16], g[16];
            tmem_ld_32x16(trow + c0, v);
            tmem_ld_32x16(trow + 128 + c0, g);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (!masked(r, c0 + j, S, p.causal)) {
                    const float e = exp2f((__uint_as_float(v[j]) - mx) * sc);
                    sum += e;
                    dsum += e * __uint_as_float(g[j]);
                }
            }
        }
        const float inv = 1.0f / sum;
        const float D = dsum * inv;  // rowsum(P o dP)
        const bool live = r < S;     // rows >= S must contribute nothing to dV / dK
        // pass C: P and dS = P o (dP - D) / 8 -> swizzled smem
        for (int c0 = 0; c0 < npad; c0 += 16) {
            uint32_t v[16], g[16];
            tmem_ld_32x16(trow + c0, v);
            tmem_ld_32x16(trow + 128 + c0, g);
            tmem_ld_wait();
            float pe[16], ds[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const bool m = !live || masked(r, c0 + j, S, p.causal);
                const float e = m ? 0.f : exp2f((__uint_as_float(v[j]) - mx) * sc) * inv;
                pe[j] = e;
                ds[j] = m ? 0.f : e * (__uint_as_float(g[j]) - D) * 0.125f;
            }
            store_p_chunk(sP, r, (c0 >> 3), make_uint4(pack_bf16(pe[0], pe[1]), pack_bf16(pe[2], pe[3]),
                                                        pack_bf16(pe[4], pe[5]), pack_bf16(pe[6], pe[7])));
            store_p_chunk(sP, r, (c0 >> 3) + 1, make_uint4(pack_bf16(pe[8], pe[9]), pack_bf16(pe[10], pe[11]),
                                                            pack_bf16(pe[12], pe[13]), pack_bf16(pe[14], pe[15])));
            store_p_chunk(sdS, r, (c0 >> 3), make_uint4(pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]),
                                                         pack_bf16(ds[4], ds[5]), pack_bf16(ds[6], ds[7])));
            store_p_chunk(sdS, r, (c0 >> 3) + 1, make_uint4(pack_bf16(ds[8], ds[9]), pack_bf16(ds[10], ds[11]),
                                                             pack_bf16(ds[12], ds[13]), pack_bf16(ds[14], ds[15])));
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc_fence_after();
            const int ksteps = npad >> 4;  // reduction over q rows (dV, dK) / kv columns (dQ), both padded to npad
            for (int k = 0; k < ksteps; ++k) {
                // A = P^T : MN-major view of the [q][kv] tile; 16 q-rows per k-step, 64-kv groups 16 KB apart
                // (!BIG: one 64-kv group only; the upper 64 output rows alias it and are never stored)
                const uint64_t a_pt = make_smem_desc_sw128(smem_u32(sP) + k * 2048, BIG ? kTile : 0, 1024);
                const uint64_t a_dst = make_smem_desc_sw128(smem_u32(sdS) + k * 2048, BIG ? kTile : 0, 1024);
                const uint64_t b_do = make_smem_desc_sw128(smem_u32(sdO) + k * 2048, 8192, 1024);
                const uint64_t b_q = make_smem_desc_sw128(smem_u32(sQ) + k * 2048, 8192, 1024);
                umma_bf16(tmem, a_pt, b_do, idesc_tn, k > 0 ? 1u : 0u);        // dV
                umma_bf16(tmem + 64, a_dst, b_q, idesc_tn, k > 0 ? 1u : 0u);   // dK
                // dQ: A = dS K-major, B = K MN-major
                const uint64_t a_ds = make_smem_desc_sw128(smem_u32(sdS) + (k >> 2) * kTile + (k & 3) * 32, 16, 1024);
                const uint64_t b_k = make_smem_desc_sw128(smem_u32(sK) + k * 2048, 8192, 1024);
                umma_bf16(tmem + 128, a_ds, b_k, idesc_nn, k > 0 ? 1u : 0u);
            }
            umma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, 1u);
        __syncwarp();
        tc_fence_after();
        {
            __nv_bfloat16* dst = p.out + (static_cast<int64_t>(b) * S + r) * (3 * d) + h * 64;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                // t = 0: dQ (cols 128..191) -> q slot ; 1: dK (64..127) -> k slot ; 2: dV (0..63) -> v slot
                const uint32_t tcol = (t == 0) ? 128u : (t == 1 ? 64u : 0u);
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(trow + tcol + c0, v);
                    tmem_ld_wait();
                    if (r < S) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 o;
                            o.x = pack_bf16(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                            o.y = pack_bf16(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                            o.z = pack_bf16(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
                            o.w = pack_bf16(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
                            *reinterpret_cast<uint4*>(dst + t * d + c0 + j) = o;
                        }
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem, kTmemCols);
    }
}

static int fwd_smem(bool big) { return (big ? 2 : 1) * kTile + 2 * kTile + 1024; }
static int bwd_smem(bool big) { return 4 * kTile + 2 * (big ? 2 : 1) * kTile + 1024; }

int init_attention(b200clip_ctx*) {
    cudaError_t e;
    e = cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem(false));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem(true));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem(false));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem(true));
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(attention): %s", cudaGetErrorString(e));
        return B200CLIP_ERR_CUDA;
    }
    return 0;
}

static int check_attn_args(b200clip_ctx* ctx, const void* a, const void* b, int64_t B, int64_t S, int64_t H) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(a && b, "attention: null pointer");
    B200_CHECK_ARG(B > 0 && H > 0 && S > 0, "attention: bad shape");
    if (S > 128) {
        set_error("attention: S=%lld > 128 needs the multi-tile kernel (not in this version)", (long long)S);
        return B200CLIP_ERR_UNSUPPORTED;
    }
    B200_CHECK_ARG(B * S < (1ll << 31) && B * H < (1ll << 31), "attention: extent too large");
    return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200clip_attn_fwd(b200clip_ctx* ctx, const void* qkv, void* out, int64_t B, int64_t S, int64_t H,
                                 int causal, void* stream) {
    int rc = check_attn_args(ctx, qkv, out, B, S, H);
    if (rc) return rc;
    B200_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "attention: out not 16-byte aligned");
    CUtensorMap tm;
    if ((rc = make_tmap_bf16_2d(ctx, &tm, qkv, 3 * H * 64, B * S, 3 * H * 64, 64, static_cast<uint32_t>(S)))) return rc;
    AttnParams p{};
    p.out = static_cast<__nv_bfloat16*>(out);
    p.B = static_cast<int>(B);
    p.S = static_cast<int>(S);
    p.H = static_cast<int>(H);
    p.causal = causal ? 1 : 0;
    p.npad = static_cast<int>((S + 15) / 16 * 16);
    const bool big = S > 64;
    const int per_sm = big ? 3 : 4;
    const int64_t work = B * H;
    const int grid = static_cast<int>(work < ctx->num_sms * per_sm ? work : ctx->num_sms * per_sm);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (big)
        attn_fwd_kernel<true><<<grid, kAttnThreads, fwd_smem(true), st>>>(tm, p);
    else
        attn_fwd_kernel<false><<<grid, kAttnThreads, fwd_smem(false), st>>>(tm, p);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_attn_bwd(b200clip_ctx* ctx, const void* qkv, const void* dout, void* dqkv, int64_t B,
                                 int64_t S, int64_t H, int causal, void* stream) {
    int rc = check_attn_args(ctx, qkv, dout, B, S, H);
    if (rc) return rc;
    B200_CHECK_ARG(dqkv && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0, "attention: dqkv null / misaligned");
    CUtensorMap tm, tmdo;
    if ((rc = make_tmap_bf16_2d(ctx, &tm, qkv, 3 * H * 64, B * S, 3 * H * 64, 64, static_cast<uint32_t>(S)))) return rc;
    if ((rc = make_tmap_bf16_2d(ctx, &tmdo, dout, H * 64, B * S, H * 64, 64, static_cast<uint32_t>(S)))) return rc;
    AttnParams p{};
    p.out = static_cast<__nv_bfloat16*>(dqkv);
    p.B = static_cast<int>(B);
    p.S = static_cast<int>(S);
    p.H = static_cast<int>(H);
    p.causal = causal ? 1 : 0;
    p.npad = static_cast<int>((S + 15) / 16 * 16);
    const bool big = S > 64;
    const int per_sm = big ? 1 : 2;
    const int64_t work = B * H;
    const int grid = static_cast<int>(work < ctx->num_sms * per_sm ? work : ctx->num_sms * per_sm);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (big)
        attn_bwd_kernel<true><<<grid, kAttnThreads, bwd_smem(true), st>>>(tm, tmdo, p);
    else
        attn_bwd_kernel<false><<<grid, kAttnThreads, bwd_smem(false), st>>>(tm, tmdo, p);
    B200_LAUNCH_CHECK();
    return 0;
}
