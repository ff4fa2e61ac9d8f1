// Fused multi-head attention (head_dim 64) on tcgen05 tensor cores: single-tile kernels for
// sequences <= 128 (forward + backward; also in a packed "varlen" form where every sample owns its own
// number of rows), a KV-streaming forward and a two-pass blocked backward for longer sequences.
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
//         (+ the row log-sum-exp, saved for the backward)
//   bwd:  S = Q K^T, dP = dO V^T -> P = exp2(S - lse), dS = P o (dP - rowsum(dO o O)) / 8 to smem ->
//         dV = P^T dO, dK = dS^T Q (A read MN-major from the same P / dS tiles), dQ = dS K
#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr int kAttnThreads = 128;     // forward: one thread per score row
constexpr int kAttnBwdThreads = 256;  // backward: two threads per score row (column halves)
constexpr int kTile = 128 * 128;      // bytes an M=128 A-operand read spans (128 rows x 128 B)
constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
    const __nv_bfloat16* o;  // bwd: forward output [B*S, H*64] (for D = rowsum(dO o O))
    float* lse;              // [B*H*S] log2-domain row log-sum-exp: fwd writes, bwd reads
    __nv_bfloat16* out;      // fwd: [B*S, H*64] ; bwd: dqkv [B*S, 3*H*64]
    int B, S, H, causal;
    int npad;  // S rounded up to a multiple of 16
    // packed ("varlen") rows: sample b owns rows cu[b] .. cu[b+1]-1 (at most S of them) of qkv / out /
    // dqkv, and lse is indexed [row * H + h].  nullptr: every sample owns S rows, lse [(b*H + h)*S + r].
    const int32_t* cu;
    // packed rows with a STATIC row count (CUDA-graph buckets): rows cu[B] .. total_rows-1 belong to no sample.
    // The kernels zero-fill them in their output (out_ld elements per row) so that everything downstream
    // (row-wise GEMMs / LayerNorms, and the token-reductions of the weight gradients) sees finite zeros.
    int total_rows, out_ld;
    int kvb;  // attn_fwd_long_kernel: rows of a KV block (multiple of 16, <= kLongKvbMax)
    const __nv_bfloat16* qkv;  // attn_fwd_long_kernel: the packed in_proj output (SIMT tail rows read it directly)
    int tail_max;              // attn_fwd_long_kernel: at most this many rows past the last full query block go to the SIMT tail
};

// zero-fill of the surplus rows of a packed output (see AttnParams::total_rows); the whole grid takes part
template <bool VARLEN>
__device__ __forceinline__ void zero_surplus_rows(const AttnParams& p) {
    if constexpr (VARLEN) {
        const int live = min(__ldg(p.cu + p.B), p.total_rows);
        const int64_t n16 = static_cast<int64_t>(p.total_rows - live) * p.out_ld / 8;   // uint4 = 8 bf16
        uint4* base = reinterpret_cast<uint4*>(p.out + static_cast<int64_t>(live) * p.out_ld);
        for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n16;
             i += static_cast<int64_t>(gridDim.x) * blockDim.x)
            base[i] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// rows and length of work item b
template <bool VARLEN>
__device__ __forceinline__ void item_rows(const AttnParams& p, int b, int& row0, int& len) {
    if constexpr (VARLEN) {
        row0 = __ldg(p.cu + b);
        len = __ldg(p.cu + b + 1) - row0;
        len = len < 0 ? 0 : (len > p.S ? p.S : len);
    } else {
        row0 = b * p.S;
        len = p.S;
    }
}
template <bool VARLEN>
__device__ __forceinline__ int64_t lse_index(const AttnParams& p, int b, int h, int row0, int r) {
    return VARLEN ? static_cast<int64_t>(row0 + r) * p.H + h : (static_cast<int64_t>(b) * p.H + h) * p.S + r;
}

// Shared-memory tiles are COMPACT: npad rows x 128 B (one row per token, 128-byte swizzled).  The
// tensor core reads 128 rows for an M=128 A operand, i.e. past the tile into whatever follows;
// those rows only produce accumulator rows >= npad, which are never used.  The allocation is padded
// so that such reads stay inside it.

// zero a shared-memory region cooperatively
__device__ __forceinline__ void zero_smem(uint8_t* p, int bytes) {
    uint4* q = reinterpret_cast<uint4*>(p);
    for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) q[i] = make_uint4(0u, 0u, 0u, 0u);
}

// 16-byte chunk `c8` (8 bf16 columns) of row r inside a set of 64-column K-major SW128 tiles that are
// `atom_bytes` apart
__device__ __forceinline__ uint8_t* p_chunk(uint8_t* base, int atom_bytes, int r, int c8) {
    return base + (c8 >> 3) * atom_bytes + r * 128 + (((c8 & 7) ^ (r & 7)) << 4);
}

__device__ __forceinline__ bool masked(int r, int c, int S, int causal) { return c >= S || (causal && c > r); }

__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// sub-CTA barrier: `count` threads (a multiple of 32) meet on hardware barrier `id` (1..15; 0 = __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// Columns [0, lim) of score row r are attended: everything (lim = S) or the causal prefix; rows past the
// sequence have none.  Score positions outside that range are masked in EVERY work item of a launch,
// so the P / dS tiles are zeroed once and the masked positions are simply never written.
__device__ __forceinline__ int row_limit(int r, int S, int causal) { return r >= S ? 0 : (causal ? r + 1 : S); }
// the same bound for a whole warp (rows row0 .. row0 + 31): TMEM loads are warp-collective, so the
// chunk loop runs to the warp-uniform bound and each thread masks inside it
__device__ __forceinline__ int warp_limit(int row0, int S, int causal) {
    return row0 >= S ? 0 : (causal ? min(S, row0 + 32) : S);
}

// ------------------------------------------------------------------------------------------------
template <bool BIG, bool VARLEN = false>  // BIG: S > 64 (P needs two 64-column tiles); VARLEN: packed rows (p.cu)
__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int S = p.S, H = p.H, npad = p.npad;
    const int TB = npad * 128;  // compact tile bytes
    constexpr int kAtoms = BIG ? 2 : 1;
    // order [P atoms | Q | K | V]: the 128-row reads of the A operands (P, Q) spill into the NEXT tile,
    // never past the allocation
    uint8_t* sP = smem;  // kAtoms tiles of TB
    uint8_t* sQ = sP + kAtoms * TB;
    uint8_t* sK = sQ + TB;
    uint8_t* sV = sK + TB;
    constexpr int kTmemCols = BIG ? 128 : 64;

    const int warp = threadIdx.x >> 5;
    const int r = threadIdx.x;  // score row owned by this thread
    const bool warp_live = warp * 32 < npad;  // warps whose 32 rows are all padding only take part in the barriers

    zero_smem(smem, (3 + kAtoms) * TB);
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

    const int d = H * 64;
    const uint32_t load_bytes = 3u * S * 128u;
    const uint32_t idesc_s = make_idesc_bf16(128, npad, 0, 0);
    const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
    const float sc = 0.125f * kLog2e;
    const int num_work = p.B * H;
    constexpr bool varlen = VARLEN;
    int hw = 0;  // varlen: chunks of this warp's P rows that may still hold a previous item's values
    auto issue_loads = [&](int w) {
        const int b = w / H, h = w - b * H;
        int row0, len;
        item_rows<VARLEN>(p, b, row0, len);
        mbar_arrive_expect_tx(&bar_load, load_bytes);
        tma_load_2d(sQ, &tm_qkv, &bar_load, h * 64, row0);
        tma_load_2d(sK, &tm_qkv, &bar_load, d + h * 64, row0);
        tma_load_2d(sV, &tm_qkv, &bar_load, 2 * d + h * 64, row0);
    };
    griddep_launch_dependents();
    griddep_wait();  // the prologue above overlapped the previous kernel's tail; global memory from here on
    if (threadIdx.x == 0 && static_cast<int>(blockIdx.x) < num_work) issue_loads(blockIdx.x);
    zero_surplus_rows<VARLEN>(p);
    uint32_t it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
        const int b = w / H, h = w - b * H;
        int row0, Si;
        item_rows<VARLEN>(p, b, row0, Si);
        const int lim = row_limit(r, Si, p.causal);                            // attended columns of this row
        const int wchunks = (warp_limit(warp * 32, Si, p.causal) + 15) >> 4;   // 16-column chunks this warp visits
        if (threadIdx.x == 0) {
            mbar_wait(&bar_load, it & 1u);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem, make_smem_desc_sw128(smem_u32(sQ) + k * 32, 16, 1024),
                          make_smem_desc_sw128(smem_u32(sK) + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
            umma_commit(&bar_mma);
        }
        float sum = 1.f;
        if (warp_live) {
            mbar_wait(&bar_mma, 0u);
            __syncwarp();
            tc_fence_after();

            // ---- softmax over row r: pass 1 max, pass 2 exp / sum / write P ----
            float mx = -INFINITY;
            for (int ci = 0; ci < wchunks; ++ci) {
                const int c0 = ci << 4;
                uint32_t v[16];
                __syncwarp();  // tcgen05.ld is warp-collective: reconverge after the per-row branches
                tmem_ld_32x16(trow + c0, v);
                tmem_ld_wait();
                if (c0 + 16 <= lim) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < lim) mx = fmaxf(mx, __uint_as_float(v[j]));
                }
            }
            const float msc = mx * sc;
            sum = 0.f;
            // packed rows: the mask differs between work items, so positions a previous (longer) item
            // wrote and this one masks are cleared explicitly, up to the warp's high-water mark
            const int nproc = (varlen && hw > wchunks) ? hw : wchunks;
            for (int ci = 0; ci < nproc; ++ci) {
                const int c0 = ci << 4;
                uint32_t v[16];
                __syncwarp();
                if (ci < wchunks) {  // warp-uniform
                    tmem_ld_32x16(trow + c0, v);
                    tmem_ld_wait();
                }
                if (varlen && (ci >= wchunks || c0 >= lim)) {
                    if (r < npad) {
                        *reinterpret_cast<uint4*>(p_chunk(sP, TB, r, c0 >> 3)) = make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(p_chunk(sP, TB, r, (c0 >> 3) + 1)) = make_uint4(0u, 0u, 0u, 0u);
                    }
                } else if (c0 < lim) {  // else (fixed length): masked in every work item, the position stays zero
                    float e[16];
                    if (c0 + 16 <= lim) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) e[j] = ex2_fast(fmaf(__uint_as_float(v[j]), sc, -msc));
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            e[j] = (c0 + j < lim) ? ex2_fast(fmaf(__uint_as_float(v[j]), sc, -msc)) : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum += e[j];
                    *reinterpret_cast<uint4*>(p_chunk(sP, TB, r, c0 >> 3)) = make_uint4(
                        pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
                    *reinterpret_cast<uint4*>(p_chunk(sP, TB, r, (c0 >> 3) + 1)) = make_uint4(
                        pack_bf16(e[8], e[9]), pack_bf16(e[10], e[11]), pack_bf16(e[12], e[13]), pack_bf16(e[14], e[15]));
                }
            }
            hw = wchunks;
            if (p.lse != nullptr && r < Si)  // log2-domain: p = exp2(s * sc - lse)
                p.lse[lse_index<VARLEN>(p, b, h, row0, r)] = msc + log2f(sum);
        }  // warp_live
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc_fence_after();
            const int ksteps = npad >> 4;
            for (int k = 0; k < ksteps; ++k)
                umma_bf16(tmem, make_smem_desc_sw128(smem_u32(sP) + (k >> 2) * TB + (k & 3) * 32, 16, 1024),
                          make_smem_desc_sw128(smem_u32(sV) + k * 2048, 8192, 1024), idesc_o, k > 0 ? 1u : 0u);
            umma_commit(&bar_mma);
        }
        if (warp_live) {
            mbar_wait(&bar_mma, 1u);
            // every tile is free again: fetch the next work item under this one's epilogue
            if (threadIdx.x == 0 && w + static_cast<int>(gridDim.x) < num_work) issue_loads(w + gridDim.x);
            __syncwarp();
            tc_fence_after();
            const float inv = 1.0f / sum;
            __nv_bfloat16* dst = p.out + static_cast<int64_t>(row0 + r) * d + h * 64;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(trow + c0, v);
                tmem_ld_wait();
                if (r < Si) {
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
// Forward for sequences longer than one tile (ViT-B/16: 197, ViT-L/14: 257, ViT-L/14@336: 577 tokens):
// one CTA per (sample, head, 128-row query block), KV streamed in blocks with the online softmax
// recurrence; S_j = Q K_j^T and P_j V_j run on the tensor core, the running output lives in registers
// (64 fp32 per thread = one row) and is rescaled by exp2(m_old - m_new) per block.
// Round 2: (a) the KV block height is chosen per sequence length (p.kvb, a multiple of 16 up to 160) so that
// the blocks are evenly filled -- 257 tokens are 2 blocks of 144 (not 128 + 128 + 1), 577 are 4 of 160 (not 5);
// (b) the next block's K is requested as soon as S_j = Q K_j^T has retired and its V as soon as P_j V_j has,
// so the loads run under the softmax / the next score MMA instead of in front of them (no extra shared
// memory: each tile is simply re-filled the moment its last reader is done); (c) ex2.approx instead of exp2f.
constexpr int kLongKvbMax = 160;                       // 2 CTAs / SM: Q 16 KB + K, V 20 KB each + P 3 x 16 KB
constexpr int kLongTailMax = 8;                        // rows past the last full query block handled by the SIMT tail

// softmax(q k^T / 8) v for `n` query rows [q_first, q_first + n) of (sample b, head h), all S keys (no mask: the
// long kernels serve the vision tower), straight from global memory (L2: this CTA has just streamed the same K / V).
// 128 threads; thread t owns keys t, t + 128, ...  s_red: 4 x 66 floats.
__device__ __forceinline__ void long_tail_rows(const AttnParams& p, int b, int h, int q_first, int n, float* s_red) {
    const int S = p.S, d = p.H * 64;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float sc = 0.125f * kLog2e;
    const __nv_bfloat16* base = p.qkv + static_cast<int64_t>(b) * S * (3 * d) + h * 64;
    for (int i = 0; i < n; ++i) {
        const int row = q_first + i;
        const __nv_bfloat16* q = base + static_cast<int64_t>(row) * (3 * d);
        float qf[64];
#pragma unroll
        for (int c = 0; c < 64; c += 8) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(q + c));
            const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16(wv[j]);
                qf[c + 2 * j] = f.x;
                qf[c + 2 * j + 1] = f.y;
            }
        }
        // scores of this thread's keys (at most 5 for S <= 640), running max
        float sv[5];
        float mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            const int key = threadIdx.x + 128 * t;
            sv[t] = -INFINITY;
            if (key < S) {
                const __nv_bfloat16* k = base + static_cast<int64_t>(key) * (3 * d) + d;
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < 64; c += 8) {
                    const uint4 u = __ldg(reinterpret_cast<const uint4*>(k + c));
                    const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = unpack_bf16(wv[j]);
                        acc = fmaf(qf[c + 2 * j], f.x, acc);
                        acc = fmaf(qf[c + 2 * j + 1], f.y, acc);
                    }
                }
                sv[t] = acc;
                mx = fmaxf(mx, acc);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        __syncthreads();   // s_red is free (previous row / previous use)
        if (lane == 0) s_red[warp] = mx;
        __syncthreads();
        mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
        // p = exp2((s - max) * sc); partial sum and partial output of this thread's keys
        float sum = 0.f;
        float of[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) of[c] = 0.f;
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            const int key = threadIdx.x + 128 * t;
            if (key < S) {
                // the tensor path rounds P to bf16 before P V: do the same so both paths agree to rounding
                const float e = __bfloat162float(__float2bfloat16_rn(ex2_fast((sv[t] - mx) * sc)));
                sum += ex2_fast((sv[t] - mx) * sc);
                const __nv_bfloat16* v = base + static_cast<int64_t>(key) * (3 * d) + 2 * d;
#pragma unroll
                for (int c = 0; c < 64; c += 8) {
                    const uint4 u = __ldg(reinterpret_cast<const uint4*>(v + c));
                    const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = unpack_bf16(wv[j]);
                        of[c + 2 * j] = fmaf(e, f.x, of[c + 2 * j]);
                        of[c + 2 * j + 1] = fmaf(e, f.y, of[c + 2 * j + 1]);
                    }
                }
            }
        }
        // reduce the 65 values (64 outputs + the sum) over the warp, then over the 4 warps through shared memory
#pragma unroll
        for (int c = 0; c < 64; ++c)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) of[c] += __shfl_xor_sync(0xffffffffu, of[c], o);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 64; ++c) s_red[warp * 66 + c] = of[c];
            s_red[warp * 66 + 64] = sum;
        }
        __syncthreads();
        if (threadIdx.x < 64) {
            const int c = threadIdx.x;
            const float l = s_red[64] + s_red[66 + 64] + s_red[132 + 64] + s_red[198 + 64];
            const float v = s_red[c] + s_red[66 + c] + s_red[132 + c] + s_red[198 + c];
            p.out[(static_cast<int64_t>(b) * S + row) * d + h * 64 + c] = __float2bfloat16_rn(v / l);
            if (c == 0 && p.lse != nullptr) p.lse[(static_cast<int64_t>(b) * p.H + h) * S + row] = mx * sc + log2f(l);
        }
    }
    __syncthreads();
}
constexpr int kLongSmem = kTile + 2 * kLongKvbMax * 128 + 3 * kTile + 1024;
__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_long_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_q, bar_k, bar_v, bar_mma;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_red[4 * 66];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kTile;
    uint8_t* sV = sK + kLongKvbMax * 128;
    uint8_t* sP = sV + kLongKvbMax * 128;  // up to three 64-column tiles of 128 rows
    constexpr int kTmemCols = 256;          // [0,160) scores, [192,256) P V

    const int warp = threadIdx.x >> 5;
    const int r = threadIdx.x;
    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_q);
        prefetch_tmap(&tm_kv);
        mbar_init(&bar_q, 1);
        mbar_init(&bar_k, 1);
        mbar_init(&bar_v, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);

    const int S = p.S, H = p.H, KVB = p.kvb;
    const int d = H * 64;
    // A handful of rows past the last full 128-row query block (ViT-L/14: 257 = 2 x 128 + 1) do not get a tensor-core
    // block of their own -- it would cost as much as a full one -- but a SIMT pass by the CTA of the last full block.
    const int tail = (!p.causal && (S & 127) <= p.tail_max) ? (S & 127) : 0;
    const int nqb = tail ? S / 128 : (S + 127) / 128;
    const uint32_t kv_bytes = static_cast<uint32_t>(KVB) * 128u;
    const uint32_t idesc_s = make_idesc_bf16(128, KVB, 0, 0);
    const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
    const float sc = 0.125f * kLog2e;
    const int num_work = p.B * H * nqb;
    uint32_t it = 0, kv_it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
        const int qb = w % nqb;
        const int bh = w / nqb;
        const int b = bh / H, h = bh % H;
        const int q0 = qb * 128;
        const int row = q0 + r;  // query index inside the sample
        const int kv_end = p.causal ? min(S, q0 + 128) : S;
        if (threadIdx.x == 0) {   // Q and the first K / V block (every tile is free: the previous item ended with a block sync)
            mbar_arrive_expect_tx(&bar_q, kTile);
            tma_load_2d(sQ, &tm_q, &bar_q, h * 64, b * S + q0);
            mbar_arrive_expect_tx(&bar_k, kv_bytes);
            tma_load_2d(sK, &tm_kv, &bar_k, d + h * 64, b * S);
            mbar_arrive_expect_tx(&bar_v, kv_bytes);
            tma_load_2d(sV, &tm_kv, &bar_v, 2 * d + h * 64, b * S);
        }
        float m = -INFINITY, l = 0.f;
        float o[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) o[c] = 0.f;
        for (int k0 = 0; k0 < kv_end; k0 += KVB, ++kv_it) {
            const bool more = k0 + KVB < kv_end;
            if (threadIdx.x == 0) {
                if (k0 == 0) mbar_wait(&bar_q, it & 1u);
                mbar_wait(&bar_k, kv_it & 1u);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem, make_smem_desc_sw128(smem_u32(sQ) + k * 32, 16, 1024),
                              make_smem_desc_sw128(smem_u32(sK) + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
                umma_commit(&bar_mma);
            }
            mbar_wait(&bar_mma, 0u);
            if (threadIdx.x == 0 && more) {   // K_j has been consumed: fetch K_{j+1} under the softmax
                mbar_arrive_expect_tx(&bar_k, kv_bytes);
                tma_load_2d(sK, &tm_kv, &bar_k, d + h * 64, b * S + k0 + KVB);
            }
            __syncwarp();
            tc_fence_after();
            float mx = m;
            for (int c0 = 0; c0 < KVB; c0 += 16) {
                uint32_t v[16];
                tmem_ld_32x16(trow + c0, v);
                tmem_ld_wait();
                if (k0 + c0 + 16 <= S && !p.causal) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int col = k0 + c0 + j;
                        if (col < S && !(p.causal && col > row)) mx = fmaxf(mx, __uint_as_float(v[j]));
                    }
                }
            }
            // a fully masked row (padding rows past S when causal) keeps mx = -inf: use 0 to stay finite
            const float mref = (mx == -INFINITY) ? 0.f : mx;
            const float alpha = (m == -INFINITY) ? 0.f : ex2_fast((m - mref) * sc);
            const float msc = mref * sc;
            float sum = 0.f;
            for (int c0 = 0; c0 < KVB; c0 += 16) {
                uint32_t v[16];
                tmem_ld_32x16(trow + c0, v);
                tmem_ld_wait();
                float e[16];
                if (k0 + c0 + 16 <= S && !p.causal) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) e[j] = ex2_fast(fmaf(__uint_as_float(v[j]), sc, -msc));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int col = k0 + c0 + j;
                        const bool msk = col >= S || (p.causal && col > row);
                        e[j] = msk ? 0.f : ex2_fast(fmaf(__uint_as_float(v[j]), sc, -msc));
                    }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) sum += e[j];
                *reinterpret_cast<uint4*>(p_chunk(sP, kTile, r, c0 >> 3)) =
                    make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
                *reinterpret_cast<uint4*>(p_chunk(sP, kTile, r, (c0 >> 3) + 1)) =
                    make_uint4(pack_bf16(e[8], e[9]), pack_bf16(e[10], e[11]), pack_bf16(e[12], e[13]), pack_bf16(e[14], e[15]));
            }
            l = l * alpha + sum;
            m = mx;
            fence_proxy_async_smem();
            tc_fence_before();
            __syncthreads();
            if (threadIdx.x == 0) {
                mbar_wait(&bar_v, kv_it & 1u);
                tc_fence_after();
                const int ksteps = KVB >> 4;
                for (int k = 0; k < ksteps; ++k)
                    umma_bf16(tmem + 192, make_smem_desc_sw128(smem_u32(sP) + (k >> 2) * kTile + (k & 3) * 32, 16, 1024),
                              make_smem_desc_sw128(smem_u32(sV) + k * 2048, 8192, 1024), idesc_o, k > 0 ? 1u : 0u);
                umma_commit(&bar_mma);
            }
            mbar_wait(&bar_mma, 1u);
            if (threadIdx.x == 0 && more) {   // V_j has been consumed: fetch V_{j+1} under the next score MMA + softmax
                mbar_arrive_expect_tx(&bar_v, kv_bytes);
                tma_load_2d(sV, &tm_kv, &bar_v, 2 * d + h * 64, b * S + k0 + KVB);
            }
            __syncwarp();
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(trow + 192 + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) o[c0 + j] = fmaf(o[c0 + j], alpha, __uint_as_float(v[j]));
            }
            tc_fence_before();
            __syncthreads();  // P and both accumulators are reused by the next block
            tc_fence_after();
        }
        if (row < S) {
            const float inv = 1.0f / l;
            __nv_bfloat16* dst = p.out + (static_cast<int64_t>(b) * S + row) * d + h * 64;
#pragma unroll
            for (int j = 0; j < 64; j += 8) {
                uint4 ov;
                ov.x = pack_bf16(o[j] * inv, o[j + 1] * inv);
                ov.y = pack_bf16(o[j + 2] * inv, o[j + 3] * inv);
                ov.z = pack_bf16(o[j + 4] * inv, o[j + 5] * inv);
                ov.w = pack_bf16(o[j + 6] * inv, o[j + 7] * inv);
                *reinterpret_cast<uint4*>(dst + j) = ov;
            }
            if (p.lse != nullptr) p.lse[(static_cast<int64_t>(b) * H + h) * S + row] = m * sc + log2f(l);
        }
        if (tail && qb == nqb - 1) long_tail_rows(p, b, h, nqb * 128, tail, s_red);
    }
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// Backward.  With the forward's row log-sum-exp and D = rowsum(dO o O) known up front there is no
// row reduction left: ONE pass over (S, dP) produces P and dS.  Two threads share a row (they take
// different 16-column chunks), 8 warps per CTA, 2 CTAs per SM (256 TMEM columns each).
// All five input tiles of a work item (Q, dO, K, V and the forward's O) arrive by TMA.  With only two CTAs
// per SM there is little occupancy to hide memory latency behind, so sequences <= 64 (the vision tower:
// 8 KB tiles) keep TWO input buffers: the tiles of item i+1 are requested at the top of item i and have a
// whole item to land.  Longer sequences (S = 77: 10 KB tiles, P / dS in two atoms) would not fit two CTAs
// that way; they keep one buffer and refill it as soon as the second MMA batch has retired.
template <bool BIG, bool VARLEN = false>
__global__ void __launch_bounds__(kAttnBwdThreads, 2)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_o, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int NBUF = BIG ? 1 : 2;
    __shared__ uint64_t bar_load[NBUF], bar_mma;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_D[2][128];  // the two column-half partial sums of D = rowsum(dO o O) per row
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int S = p.S, H = p.H, npad = p.npad;
    const int TB = npad * 128;
    constexpr int kAtoms = BIG ? 2 : 1;
    // order [P | dS | NBUF x (Q | dO | K | V | O)]: 128-row A-operand reads (dS, Q, dO) spill into the following tiles only
    uint8_t* sP = smem;
    uint8_t* sdS = sP + kAtoms * TB;
    uint8_t* sIn = sdS + kAtoms * TB;
    constexpr int kTmemCols = 256;
    // TMEM columns: [0,128) S, later dV [0,64) + dK [64,128);  [128,256) dP, later dQ [128,192)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = (warp & 3) * 32 + lane;  // score row (TMEM lane) of this thread
    const int half = warp >> 2;            // which share of the columns / output chunks it takes
    const bool warp_live = (warp & 3) * 32 < npad;  // all-padding warps only take part in the barriers

    zero_smem(smem, (5 * NBUF + 2 * kAtoms) * TB);
    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_qkv);
        prefetch_tmap(&tm_do);
        prefetch_tmap(&tm_o);
#pragma unroll
        for (int i = 0; i < NBUF; ++i) mbar_init(&bar_load[i], 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, kTmemCols);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);

    const int d = H * 64;
    const uint32_t load_bytes = 5u * S * 128u;
    const uint32_t idesc_s = make_idesc_bf16(128, npad, 0, 0);    // S = Q K^T, dP = dO V^T
    const uint32_t idesc_tn = make_idesc_bf16(128, 64, 1, 1);     // dV = P^T dO, dK = dS^T Q
    const uint32_t idesc_nn = make_idesc_bf16(128, 64, 0, 1);     // dQ = dS K
    const float sc = 0.125f * kLog2e;
    const int num_work = p.B * H;
    constexpr bool varlen = VARLEN;
    int hw = 0;  // varlen: chunks of this warp pair's P / dS rows that may still hold a previous item's values
    auto issue_loads = [&](int w, int buf) {
        const int b = w / H, h = w - b * H;
        int row0, len;
        item_rows<VARLEN>(p, b, row0, len);
        uint8_t* base = sIn + buf * 5 * TB;
        mbar_arrive_expect_tx(&bar_load[buf], load_bytes);
        tma_load_2d(base, &tm_qkv, &bar_load[buf], h * 64, row0);
        tma_load_2d(base + TB, &tm_do, &bar_load[buf], h * 64, row0);
        tma_load_2d(base + 2 * TB, &tm_qkv, &bar_load[buf], d + h * 64, row0);
        tma_load_2d(base + 3 * TB, &tm_qkv, &bar_load[buf], 2 * d + h * 64, row0);
        tma_load_2d(base + 4 * TB, &tm_o, &bar_load[buf], h * 64, row0);
    };
    // the forward's row log-sum-exp of work item w for this thread's row (0 for rows past the sample)
    auto fetch_lse = [&](int w) -> float {
        if (w >= num_work) return 0.f;
        const int b = w / H, h = w - b * H;
        int row0, len;
        item_rows<VARLEN>(p, b, row0, len);
        return r < len ? __ldg(p.lse + lse_index<VARLEN>(p, b, h, row0, r)) : 0.f;
    };
    griddep_launch_dependents();
    griddep_wait();  // the prologue above overlapped the previous kernel's tail; global memory from here on
    if (threadIdx.x == 0 && static_cast<int>(blockIdx.x) < num_work) issue_loads(blockIdx.x, 0);
    float m2_next = fetch_lse(blockIdx.x);
    zero_surplus_rows<VARLEN>(p);
    uint32_t it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
        const int b = w / H, h = w - b * H;
        int row0, Si;
        item_rows<VARLEN>(p, b, row0, Si);
        const int buf = NBUF == 2 ? static_cast<int>(it & 1u) : 0;
        const uint32_t load_phase = NBUF == 2 ? ((it >> 1) & 1u) : (it & 1u);
        uint8_t* sQ = sIn + buf * 5 * TB;
        uint8_t* sdO = sQ + TB;
        uint8_t* sK = sdO + TB;
        uint8_t* sV = sK + TB;
        uint8_t* sO = sV + TB;
        // two buffers: the other one was last read by item it-1, which ended with a __syncthreads
        if (NBUF == 2 && threadIdx.x == 0 && w + static_cast<int>(gridDim.x) < num_work)
            issue_loads(w + gridDim.x, buf ^ 1);
        const bool live = r < Si;
        const int lim = row_limit(r, Si, p.causal);                                // attended columns of this row
        const int wchunks = (warp_limit((warp & 3) * 32, Si, p.causal) + 15) >> 4;  // 16-column chunks this warp pair visits
        // row statistic from the forward: fetched one item ahead, so its latency is never exposed
        const float m2 = m2_next;
        m2_next = fetch_lse(w + gridDim.x);
        if (threadIdx.x == 0) {
            mbar_wait(&bar_load[buf], load_phase);
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
        if (warp_live) {
            // D = rowsum(dO o O) while the tensor core works; the two threads of a row each take half of the 64 columns
            mbar_wait(&bar_load[buf], load_phase);
            float part = 0.f;
            if (live) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int cc = half * 4 + c;
                    const uint4 dv = *reinterpret_cast<const uint4*>(sdO + r * 128 + ((cc ^ (r & 7)) << 4));
                    const uint4 ov = *reinterpret_cast<const uint4*>(sO + r * 128 + ((cc ^ (r & 7)) << 4));
                    const uint32_t a[4] = {dv.x, dv.y, dv.z, dv.w}, o4[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 x = unpack_bf16(a[j]), y = unpack_bf16(o4[j]);
                        part = fmaf(x.x, y.x, part);
                        part = fmaf(x.y, y.y, part);
                    }
                }
            }
            s_D[half][r] = part;
            named_bar_sync(1 + (warp & 3), 64);  // the two warps that share these 32 rows
            const float D8 = (s_D[0][r] + s_D[1][r]) * 0.125f;
            mbar_wait(&bar_mma, 0u);
            __syncwarp();
            tc_fence_after();

            // single pass: P = exp2(S*sc - lse), dS = P o (dP - D) / 8 -> swizzled smem (bf16); the two
            // warps of a row group take alternate 16-column chunks
            // packed rows: the mask differs between work items, so positions a previous (longer) item
            // wrote and this one masks are cleared explicitly, up to the warp pair's high-water mark
            const int nproc = (varlen && hw > wchunks) ? hw : wchunks;
            for (int ci = half; ci < nproc; ci += 2) {
                const int c0 = ci << 4;
                uint32_t v[16], g[16];
                __syncwarp();  // tcgen05.ld is warp-collective: reconverge after the per-row branches
                if (ci < wchunks) {  // warp-uniform
                    tmem_ld_32x16(trow + c0, v);
                    tmem_ld_32x16(trow + 128 + c0, g);
                    tmem_ld_wait();
                }
                if (varlen && (ci >= wchunks || c0 >= lim)) {
                    if (r < npad) {
                        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(p_chunk(sP, TB, r, c0 >> 3)) = z;
                        *reinterpret_cast<uint4*>(p_chunk(sP, TB, r, (c0 >> 3) + 1)) = z;
                        *reinterpret_cast<uint4*>(p_chunk(sdS, TB, r, c0 >> 3)) = z;
                        *reinterpret_cast<uint4*>(p_chunk(sdS, TB, r, (c0 >> 3) + 1)) = z;
                    }
                } else if (c0 < lim) {  // else (fixed length): masked in every work item, the positions stay zero
                    float pe[16], ds[16];
                    if (c0 + 16 <= lim) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            pe[j] = ex2_fast(fmaf(__uint_as_float(v[j]), sc, -m2));
                            ds[j] = pe[j] * fmaf(__uint_as_float(g[j]), 0.125f, -D8);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const bool ok = c0 + j < lim;
                            pe[j] = ok ? ex2_fast(fmaf(__uint_as_float(v[j]), sc, -m2)) : 0.f;
                            ds[j] = ok ? pe[j] * fmaf(__uint_as_float(g[j]), 0.125f, -D8) : 0.f;
                        }
                    }
                    *reinterpret_cast<uint4*>(p_chunk(sP, TB, r, c0 >> 3)) = make_uint4(
                        pack_bf16(pe[0], pe[1]), pack_bf16(pe[2], pe[3]), pack_bf16(pe[4], pe[5]), pack_bf16(pe[6], pe[7]));
                    *reinterpret_cast<uint4*>(p_chunk(sP, TB, r, (c0 >> 3) + 1)) = make_uint4(
                        pack_bf16(pe[8], pe[9]), pack_bf16(pe[10], pe[11]), pack_bf16(pe[12], pe[13]), pack_bf16(pe[14], pe[15]));
                    *reinterpret_cast<uint4*>(p_chunk(sdS, TB, r, c0 >> 3)) = make_uint4(
                        pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]), pack_bf16(ds[4], ds[5]), pack_bf16(ds[6], ds[7]));
                    *reinterpret_cast<uint4*>(p_chunk(sdS, TB, r, (c0 >> 3) + 1)) = make_uint4(
                        pack_bf16(ds[8], ds[9]), pack_bf16(ds[10], ds[11]), pack_bf16(ds[12], ds[13]), pack_bf16(ds[14], ds[15]));
                }
            }
            hw = wchunks;
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc_fence_after();
            const int ksteps = npad >> 4;  // reduction over q rows (dV, dK) / kv columns (dQ), both padded to npad
            for (int k = 0; k < ksteps; ++k) {
                // A = P^T / dS^T : MN-major view of the [q][kv] tile; 16 q-rows per k-step; the second
                // 64-kv group lives one compact tile further (!BIG: a single group, the upper 64 output
                // rows alias it and are never stored)
                const uint64_t a_pt = make_smem_desc_sw128(smem_u32(sP) + k * 2048, BIG ? TB : 0, 1024);
                const uint64_t a_dst = make_smem_desc_sw128(smem_u32(sdS) + k * 2048, BIG ? TB : 0, 1024);
                const uint64_t b_do = make_smem_desc_sw128(smem_u32(sdO) + k * 2048, 8192, 1024);
                const uint64_t b_q = make_smem_desc_sw128(smem_u32(sQ) + k * 2048, 8192, 1024);
                umma_bf16(tmem, a_pt, b_do, idesc_tn, k > 0 ? 1u : 0u);        // dV
                umma_bf16(tmem + 64, a_dst, b_q, idesc_tn, k > 0 ? 1u : 0u);   // dK
                // dQ: A = dS K-major, B = K MN-major
                const uint64_t a_ds = make_smem_desc_sw128(smem_u32(sdS) + (k >> 2) * TB + (k & 3) * 32, 16, 1024);
                const uint64_t b_k = make_smem_desc_sw128(smem_u32(sK) + k * 2048, 8192, 1024);
                umma_bf16(tmem + 128, a_ds, b_k, idesc_nn, k > 0 ? 1u : 0u);
            }
            umma_commit(&bar_mma);
        }
        if (warp_live) {
            mbar_wait(&bar_mma, 1u);
            // one buffer: every tile is free again, fetch the next work item under this one's epilogue
            if (NBUF == 1 && threadIdx.x == 0 && w + static_cast<int>(gridDim.x) < num_work) issue_loads(w + gridDim.x, 0);
            __syncwarp();
            tc_fence_after();
            // six 32-column output chunks per row: dQ (TMEM 128..191), dK (64..127), dV (0..63);
            // the two threads of a row take three each
            __nv_bfloat16* dst = p.out + static_cast<int64_t>(row0 + r) * (3 * d) + h * 64;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int oc = half * 3 + q;         // 0..5
                const int t = oc >> 1, c0 = (oc & 1) * 32;
                const uint32_t tcol = (t == 0) ? 128u : (t == 1 ? 64u : 0u);
                uint32_t v[32];
                tmem_ld_32x32(trow + tcol + c0, v);
                tmem_ld_wait();
                if (live) {
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
// Backward for sequences longer than one tile (ViT-B/16 / ViT-L/14 fine-tuning): two passes over the
// 128 x 128 blocks of the score matrix, both recomputing S = Q K^T and dP = dO V^T on the tensor core.
//   MODE 0: one CTA per (sample, head, KV block j), loops over the query blocks i and accumulates
//           dV_j += P_ij^T dO_i, dK_j += dS_ij^T Q_i in TMEM;
//   MODE 1: one CTA per (sample, head, query block i), loops over the KV blocks j and accumulates
//           dQ_i += dS_ij K_j in TMEM.
// P = exp2(S * sc - lse) needs the forward's row log-sum-exp, dS = P o (dP - D) / 8 the row sums
// D = rowsum(dO o O), precomputed by attn_delta_kernel.  Two threads per score row, 256 threads.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, float* __restrict__ delta,
                  int B, int S, int H) {
    // delta[(b*H + h)*S + s] = sum_c dO[b*S+s, h*64+c] * O[b*S+s, h*64+c]; one thread per (token, head)
    const int64_t n = static_cast<int64_t>(B) * S * H;
    const int d = H * 64;
    for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < n;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int h = static_cast<int>(t % H);
        const int64_t tok = t / H;
        const uint4* po = reinterpret_cast<const uint4*>(o + tok * d + h * 64);
        const uint4* pd = reinterpret_cast<const uint4*>(dout + tok * d + h * 64);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 a = __ldg(po + c), g = __ldg(pd + c);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 x = unpack_bf16(aw[j]), y = unpack_bf16(gw[j]);
                acc = fmaf(x.x, y.x, acc);
                acc = fmaf(x.y, y.y, acc);
            }
        }
        const int64_t b = tok / S, srow = tok % S;
        delta[(b * H + h) * S + srow] = acc;
    }
}

template <int MODE>
__global__ void __launch_bounds__(kAttnBwdThreads)
attn_bwd_long_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                     const AttnParams p, const float* __restrict__ delta) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_outer, bar_inner, bar_mma;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sP = smem;               // 2 tiles
    uint8_t* sdS = sP + 2 * kTile;    // 2 tiles
    uint8_t* sQ = sdS + 2 * kTile;
    uint8_t* sdO = sQ + kTile;
    uint8_t* sK = sdO + kTile;
    uint8_t* sV = sK + kTile;
    constexpr int kTmemCols = 512;    // [0,128) S, [128,256) dP, [256,320) dV | dQ, [320,384) dK

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = (warp & 3) * 32 + lane;
    const int half = warp >> 2;
    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_qkv);
        prefetch_tmap(&tm_do);
        mbar_init(&bar_outer, 1);
        mbar_init(&bar_inner, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);

    const int S = p.S, H = p.H;
    const int d = H * 64;
    const int nblk = (S + 127) / 128;
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_tn = make_idesc_bf16(128, 64, 1, 1);
    const uint32_t idesc_nn = make_idesc_bf16(128, 64, 0, 1);
    const float sc = 0.125f * kLog2e;
    const int num_work = p.B * H * nblk;
    uint32_t it = 0, in_it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
        const int ob = w % nblk;  // outer block: KV block (MODE 0) / query block (MODE 1)
        const int bh = w / nblk;
        const int b = bh / H, h = bh % H;
        const int o0 = ob * 128;
        const int64_t tok0 = static_cast<int64_t>(b) * S;
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(&bar_outer, 2 * kTile);
            if (MODE == 0) {
                tma_load_2d(sK, &tm_qkv, &bar_outer, d + h * 64, static_cast<int>(tok0) + o0);
                tma_load_2d(sV, &tm_qkv, &bar_outer, 2 * d + h * 64, static_cast<int>(tok0) + o0);
            } else {
                tma_load_2d(sQ, &tm_qkv, &bar_outer, h * 64, static_cast<int>(tok0) + o0);
                tma_load_2d(sdO, &tm_do, &bar_outer, h * 64, static_cast<int>(tok0) + o0);
            }
        }
        // inner range (causal: only blocks that intersect the lower triangle)
        const int ib0 = (MODE == 0) ? (p.causal ? ob : 0) : 0;
        const int ib1 = (MODE == 0) ? nblk : (p.causal ? ob + 1 : nblk);
        bool first = true;
        for (int ib = ib0; ib < ib1; ++ib, ++in_it) {
            const int i0 = ib * 128;
            const int q0 = (MODE == 0) ? i0 : o0;   // query block start
            const int k0 = (MODE == 0) ? o0 : i0;   // kv block start
            if (threadIdx.x == 0) {
                mbar_arrive_expect_tx(&bar_inner, 2 * kTile);
                if (MODE == 0) {
                    tma_load_2d(sQ, &tm_qkv, &bar_inner, h * 64, static_cast<int>(tok0) + i0);
                    tma_load_2d(sdO, &tm_do, &bar_inner, h * 64, static_cast<int>(tok0) + i0);
                } else {
                    tma_load_2d(sK, &tm_qkv, &bar_inner, d + h * 64, static_cast<int>(tok0) + i0);
                    tma_load_2d(sV, &tm_qkv, &bar_inner, 2 * d + h * 64, static_cast<int>(tok0) + i0);
                }
            }
            const int qrow = q0 + r;
            const bool live = qrow < S;
            float m2 = 0.f, D = 0.f;
            if (live) {
                const int64_t si = (static_cast<int64_t>(b) * H + h) * S + qrow;
                m2 = p.lse[si];
                D = delta[si];
            }
            if (threadIdx.x == 0) {
                if (first) mbar_wait(&bar_outer, it & 1u);
                mbar_wait(&bar_inner, in_it & 1u);
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
            for (int ci = half * 4; ci < half * 4 + 4; ++ci) {
                const int c0 = ci << 4;
                uint32_t v[16], g[16];
                tmem_ld_32x16(trow + c0, v);
                tmem_ld_32x16(trow + 128 + c0, g);
                tmem_ld_wait();
                float pe[16], ds[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = k0 + c0 + j;
                    const bool msk = !live || col >= S || (p.causal && col > qrow);
                    const float e = msk ? 0.f : exp2f(__uint_as_float(v[j]) * sc - m2);
                    pe[j] = e;
                    ds[j] = msk ? 0.f : e * (__uint_as_float(g[j]) - D) * 0.125f;
                }
                *reinterpret_cast<uint4*>(p_chunk(sP, kTile, r, c0 >> 3)) =
                    make_uint4(pack_bf16(pe[0], pe[1]), pack_bf16(pe[2], pe[3]), pack_bf16(pe[4], pe[5]), pack_bf16(pe[6], pe[7]));
                *reinterpret_cast<uint4*>(p_chunk(sP, kTile, r, (c0 >> 3) + 1)) =
                    make_uint4(pack_bf16(pe[8], pe[9]), pack_bf16(pe[10], pe[11]), pack_bf16(pe[12], pe[13]), pack_bf16(pe[14], pe[15]));
                *reinterpret_cast<uint4*>(p_chunk(sdS, kTile, r, c0 >> 3)) =
                    make_uint4(pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]), pack_bf16(ds[4], ds[5]), pack_bf16(ds[6], ds[7]));
                *reinterpret_cast<uint4*>(p_chunk(sdS, kTile, r, (c0 >> 3) + 1)) =
                    make_uint4(pack_bf16(ds[8], ds[9]), pack_bf16(ds[10], ds[11]), pack_bf16(ds[12], ds[13]), pack_bf16(ds[14], ds[15]));
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncthreads();
            if (threadIdx.x == 0) {
                tc_fence_after();
                const uint32_t acc = first ? 0u : 1u;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (MODE == 0) {
                        const uint64_t a_pt = make_smem_desc_sw128(smem_u32(sP) + k * 2048, kTile, 1024);
                        const uint64_t a_dst = make_smem_desc_sw128(smem_u32(sdS) + k * 2048, kTile, 1024);
                        const uint64_t b_do = make_smem_desc_sw128(smem_u32(sdO) + k * 2048, 8192, 1024);
                        const uint64_t b_q = make_smem_desc_sw128(smem_u32(sQ) + k * 2048, 8192, 1024);
                        umma_bf16(tmem + 256, a_pt, b_do, idesc_tn, (k > 0) ? 1u : acc);   // dV_j += P^T dO_i
                        umma_bf16(tmem + 320, a_dst, b_q, idesc_tn, (k > 0) ? 1u : acc);   // dK_j += dS^T Q_i
                    } else {
                        const uint64_t a_ds = make_smem_desc_sw128(smem_u32(sdS) + (k >> 2) * kTile + (k & 3) * 32, 16, 1024);
                        const uint64_t b_k = make_smem_desc_sw128(smem_u32(sK) + k * 2048, 8192, 1024);
                        umma_bf16(tmem + 256, a_ds, b_k, idesc_nn, (k > 0) ? 1u : acc);    // dQ_i += dS K_j
                    }
                }
                umma_commit(&bar_mma);
            }
            mbar_wait(&bar_mma, 1u);   // the inner tiles, P and dS are rewritten by the next iteration
            __syncwarp();
            tc_fence_after();
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            first = false;
        }
        // ---- outputs of this work item
        {
            const int orow = o0 + r;
            const bool ok = orow < S && !first;
            __nv_bfloat16* dst = p.out + (tok0 + orow) * (3 * d) + h * 64;
            if (MODE == 0) {
                // half 0 stores dV (TMEM 256..319) into the v slot, half 1 stores dK (320..383) into the k slot
                const uint32_t tcol = half == 0 ? 256u : 320u;
                __nv_bfloat16* dd = dst + (half == 0 ? 2 * d : d);
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(trow + tcol + c0, v);
                    tmem_ld_wait();
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8)
                            *reinterpret_cast<uint4*>(dd + c0 + j) =
                                make_uint4(pack_bf16(__uint_as_float(v[j]), __uint_as_float(v[j + 1])),
                                           pack_bf16(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])),
                                           pack_bf16(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])),
                                           pack_bf16(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])));
                    }
                }
            } else {
                const int c0 = half * 32;
                uint32_t v[32];
                tmem_ld_32x32(trow + 256 + c0, v);
                tmem_ld_wait();
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8)
                        *reinterpret_cast<uint4*>(dst + c0 + j) =
                            make_uint4(pack_bf16(__uint_as_float(v[j]), __uint_as_float(v[j + 1])),
                                       pack_bf16(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])),
                                       pack_bf16(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])),
                                       pack_bf16(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])));
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

// dynamic shared memory: compact tiles (A operands first, so their 128-row reads stay inside) + alignment.
// The last A tile (Q fwd / dO bwd) is followed by >= 2 more tiles of npad >= 16 rows... not enough when
// npad < 43: keep the allocation at least kTile past the start of that tile.
static int fwd_smem(bool big, int npad) {
    const int tb = npad * 128, atoms = big ? 2 : 1;
    const int need = atoms * tb + kTile;  // Q starts after the P atoms and is read for 128 rows
    const int have = (3 + atoms) * tb;
    return (have > need ? have : need) + 1024;
}
static int bwd_smem(bool big, int npad) {
    const int tb = npad * 128, atoms = big ? 2 : 1, nbuf = big ? 1 : 2;  // (NBUF of attn_bwd_kernel)
    // the last buffer's dO starts after P, dS, the earlier buffers and its own Q, and is read for 128 rows
    const int need = (2 * atoms + 5 * (nbuf - 1) + 1) * tb + kTile;
    const int have = (5 * nbuf + 2 * atoms) * tb;
    return (have > need ? have : need) + 1024;
}

int init_attention(b200clip_ctx*) {
    cudaError_t e;
    e = cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem(false, 64));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem(true, 128));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attn_fwd_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLongSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attn_bwd_long_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * kTile + 1024);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attn_bwd_long_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * kTile + 1024);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem(false, 64));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem(true, 128));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attn_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem(false, 64));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attn_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem(true, 128));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attn_bwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem(false, 64));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attn_bwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem(true, 128));
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(attention): %s", cudaGetErrorString(e));
        return B200CLIP_ERR_CUDA;
    }
    return 0;
}

static int check_attn_args(b200clip_ctx* ctx, const void* a, const void* b, int64_t B, int64_t S, int64_t H,
                           bool allow_long) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(a && b, "attention: null pointer");
    B200_CHECK_ARG(B > 0 && H > 0 && S > 0, "attention: bad shape");
    if (S > 128 && !allow_long) {
        set_error("attention backward: S=%lld > 128 needs the multi-tile backward kernel (not in this version)",
                  (long long)S);
        return B200CLIP_ERR_UNSUPPORTED;
    }
    B200_CHECK_ARG(B * S < (1ll << 31) && B * H < (1ll << 31), "attention: extent too large");
    return 0;
}

}  // namespace b200

using namespace b200;

static int ctas_per_sm(int smem_bytes, int tmem_cols, int cap) {
    int n = (227 * 1024) / (smem_bytes + 1024);
    const int t = 512 / tmem_cols;
    if (n > t) n = t;
    if (n > cap) n = cap;
    return n < 1 ? 1 : n;
}

extern "C" int b200clip_attn_fwd(b200clip_ctx* ctx, const void* qkv, void* out, float* lse, int64_t B, int64_t S,
                                 int64_t H, int causal, void* stream) {
    int rc = check_attn_args(ctx, qkv, out, B, S, H, true);
    if (rc) return rc;
    B200_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "attention: out not 16-byte aligned");
    if (S > 128) {
        // KV block height: the fewest blocks of at most kLongKvbMax rows, evenly filled (257 -> 2 x 144, 577 -> 4 x 160)
        const int nkb = static_cast<int>((S + kLongKvbMax - 1) / kLongKvbMax);
        const int kvb = static_cast<int>(((S + nkb - 1) / nkb + 15) / 16 * 16);
        CUtensorMap tml, tmkv;
        if ((rc = make_tmap_bf16_2d(ctx, &tml, qkv, 3 * H * 64, B * S, 3 * H * 64, 64, 128))) return rc;
        if ((rc = make_tmap_bf16_2d(ctx, &tmkv, qkv, 3 * H * 64, B * S, 3 * H * 64, 64, static_cast<uint32_t>(kvb)))) return rc;
        AttnParams pl{};
        pl.kvb = kvb;
        pl.qkv = static_cast<const __nv_bfloat16*>(qkv);
        static const int tail_max = [] {
            const char* e = getenv("B200CLIP_LONG_TAIL");   // tuning: 0 disables the SIMT tail
            const int v = e ? atoi(e) : kLongTailMax;
            return v < 0 ? 0 : (v > kLongTailMax ? kLongTailMax : v);
        }();
        pl.tail_max = causal ? 0 : tail_max;
        pl.out = static_cast<__nv_bfloat16*>(out);
        pl.lse = lse;
        pl.B = static_cast<int>(B);
        pl.S = static_cast<int>(S);
        pl.H = static_cast<int>(H);
        pl.causal = causal ? 1 : 0;
        pl.npad = 128;
        const int64_t work_l = B * H * (((S & 127) <= pl.tail_max && (S & 127)) ? S / 128 : (S + 127) / 128);
        B200_CHECK_ARG(work_l < (1ll << 31), "attention: extent too large");
        const int grid_l = static_cast<int>(work_l < ctx->num_sms * 2 ? work_l : ctx->num_sms * 2);
        attn_fwd_long_kernel<<<grid_l, kAttnThreads, kLongSmem, static_cast<cudaStream_t>(stream)>>>(tml, tmkv, pl);
        B200_LAUNCH_CHECK();
        return 0;
    }
    CUtensorMap tm;
    if ((rc = make_tmap_bf16_2d(ctx, &tm, qkv, 3 * H * 64, B * S, 3 * H * 64, 64, static_cast<uint32_t>(S)))) return rc;
    AttnParams p{};
    p.out = static_cast<__nv_bfloat16*>(out);
    p.lse = lse;
    p.B = static_cast<int>(B);
    p.S = static_cast<int>(S);
    p.H = static_cast<int>(H);
    p.causal = causal ? 1 : 0;
    p.npad = static_cast<int>((S + 15) / 16 * 16);
    const bool big = S > 64;
    const int smem = fwd_smem(big, p.npad);
    const int per_sm = ctas_per_sm(smem, big ? 128 : 64, 8);
    const int64_t work = B * H;
    const int grid = static_cast<int>(work < ctx->num_sms * per_sm ? work : ctx->num_sms * per_sm);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (big)
        B200_CHECK_CUDA(launch_pdl(attn_fwd_kernel<true>, dim3(grid), dim3(kAttnThreads), smem, st, tm, p));
    else
        B200_CHECK_CUDA(launch_pdl(attn_fwd_kernel<false>, dim3(grid), dim3(kAttnThreads), smem, st, tm, p));
    B200_LAUNCH_CHECK();
    return 0;
}

// Packed rows: sample b owns rows cu[b] .. cu[b+1]-1 (1 .. S_max <= 128 of them) of qkv [total_rows, 3*H*64].
extern "C" int b200clip_attn_fwd_varlen(b200clip_ctx* ctx, const void* qkv, void* out, float* lse, const int32_t* cu,
                                        int64_t B, int64_t S_max, int64_t H, int64_t total_rows, int causal,
                                        void* stream) {
    int rc = check_attn_args(ctx, qkv, out, B, S_max, H, false);
    if (rc) return rc;
    B200_CHECK_ARG(cu != nullptr && total_rows > 0 && total_rows < (1ll << 31) && S_max <= 128,
                   "attn_fwd_varlen: needs cu, total_rows and S_max <= 128");
    B200_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "attention: out not 16-byte aligned");
    CUtensorMap tm;
    if ((rc = make_tmap_bf16_2d(ctx, &tm, qkv, 3 * H * 64, total_rows, 3 * H * 64, 64, static_cast<uint32_t>(S_max))))
        return rc;
    AttnParams p{};
    p.out = static_cast<__nv_bfloat16*>(out);
    p.lse = lse;
    p.B = static_cast<int>(B);
    p.S = static_cast<int>(S_max);
    p.H = static_cast<int>(H);
    p.causal = causal ? 1 : 0;
    p.npad = static_cast<int>((S_max + 15) / 16 * 16);
    p.cu = cu;
    p.total_rows = static_cast<int>(total_rows);
    p.out_ld = static_cast<int>(H * 64);
    const bool big = S_max > 64;
    const int smem = fwd_smem(big, p.npad);
    const int per_sm = ctas_per_sm(smem, big ? 128 : 64, 8);
    const int64_t work = B * H;
    const int grid = static_cast<int>(work < ctx->num_sms * per_sm ? work : ctx->num_sms * per_sm);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (big)
        B200_CHECK_CUDA(launch_pdl(attn_fwd_kernel<true, true>, dim3(grid), dim3(kAttnThreads), smem, st, tm, p));
    else
        B200_CHECK_CUDA(launch_pdl(attn_fwd_kernel<false, true>, dim3(grid), dim3(kAttnThreads), smem, st, tm, p));
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_attn_bwd_varlen(b200clip_ctx* ctx, const void* qkv, const void* out, const float* lse,
                                        const void* dout, void* dqkv, const int32_t* cu, int64_t B, int64_t S_max,
                                        int64_t H, int64_t total_rows, int causal, void* stream) {
    int rc = check_attn_args(ctx, qkv, dout, B, S_max, H, false);
    if (rc) return rc;
    B200_CHECK_ARG(out && lse && cu && total_rows > 0 && total_rows < (1ll << 31) && S_max <= 128,
                   "attn_bwd_varlen: needs out, lse, cu, total_rows and S_max <= 128");
    B200_CHECK_ARG(dqkv && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                   "attention: dqkv / out null or misaligned");
    CUtensorMap tm, tmdo, tmo;
    if ((rc = make_tmap_bf16_2d(ctx, &tm, qkv, 3 * H * 64, total_rows, 3 * H * 64, 64, static_cast<uint32_t>(S_max))))
        return rc;
    if ((rc = make_tmap_bf16_2d(ctx, &tmdo, dout, H * 64, total_rows, H * 64, 64, static_cast<uint32_t>(S_max))))
        return rc;
    if ((rc = make_tmap_bf16_2d(ctx, &tmo, out, H * 64, total_rows, H * 64, 64, static_cast<uint32_t>(S_max))))
        return rc;
    AttnParams p{};
    p.o = static_cast<const __nv_bfloat16*>(out);
    p.lse = const_cast<float*>(lse);
    p.out = static_cast<__nv_bfloat16*>(dqkv);
    p.B = static_cast<int>(B);
    p.S = static_cast<int>(S_max);
    p.H = static_cast<int>(H);
    p.causal = causal ? 1 : 0;
    p.npad = static_cast<int>((S_max + 15) / 16 * 16);
    p.cu = cu;
    p.total_rows = static_cast<int>(total_rows);
    p.out_ld = static_cast<int>(3 * H * 64);
    const bool big = S_max > 64;
    const int smem = bwd_smem(big, p.npad);
    const int per_sm = ctas_per_sm(smem, 256, 2);
    const int64_t work = B * H;
    const int grid = static_cast<int>(work < ctx->num_sms * per_sm ? work : ctx->num_sms * per_sm);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (big)
        B200_CHECK_CUDA(launch_pdl(attn_bwd_kernel<true, true>, dim3(grid), dim3(kAttnBwdThreads), smem, st, tm, tmdo, tmo, p));
    else
        B200_CHECK_CUDA(launch_pdl(attn_bwd_kernel<false, true>, dim3(grid), dim3(kAttnBwdThreads), smem, st, tm, tmdo, tmo, p));
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_attn_bwd(b200clip_ctx* ctx, const void* qkv, const void* out, const float* lse,
                                 const void* dout, void* dqkv, void* workspace, int64_t workspace_bytes, int64_t B,
                                 int64_t S, int64_t H, int causal, void* stream) {
    int rc = check_attn_args(ctx, qkv, dout, B, S, H, true);
    if (rc) return rc;
    B200_CHECK_ARG(out && lse, "attention bwd: needs the forward's out and lse");
    if (S > 128) {
        B200_CHECK_ARG(workspace != nullptr && workspace_bytes >= static_cast<int64_t>(B * H * S * sizeof(float)),
                       "attention bwd (S > 128): needs a workspace of B*H*S floats");
        B200_CHECK_ARG(dqkv && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0, "attention: dqkv null / misaligned");
        CUtensorMap tmq, tmd;
        if ((rc = make_tmap_bf16_2d(ctx, &tmq, qkv, 3 * H * 64, B * S, 3 * H * 64, 64, 128))) return rc;
        if ((rc = make_tmap_bf16_2d(ctx, &tmd, dout, H * 64, B * S, H * 64, 64, 128))) return rc;
        AttnParams pl{};
        pl.o = static_cast<const __nv_bfloat16*>(out);
        pl.lse = const_cast<float*>(lse);
        pl.out = static_cast<__nv_bfloat16*>(dqkv);
        pl.B = static_cast<int>(B);
        pl.S = static_cast<int>(S);
        pl.H = static_cast<int>(H);
        pl.causal = causal ? 1 : 0;
        pl.npad = 128;
        float* delta = static_cast<float*>(workspace);
        cudaStream_t stl = static_cast<cudaStream_t>(stream);
        const int64_t nd = B * S * H;
        const int gd = static_cast<int>(ceil_div(nd, 256) < ctx->num_sms * 8 ? ceil_div(nd, 256) : ctx->num_sms * 8);
        attn_delta_kernel<<<gd, 256, 0, stl>>>(pl.o, static_cast<const __nv_bfloat16*>(dout), delta, pl.B, pl.S, pl.H);
        B200_LAUNCH_CHECK();
        const int64_t work_l = B * H * ((S + 127) / 128);
        B200_CHECK_ARG(work_l < (1ll << 31), "attention: extent too large");
        const int grid_l = static_cast<int>(work_l < ctx->num_sms ? work_l : ctx->num_sms);
        attn_bwd_long_kernel<0><<<grid_l, kAttnBwdThreads, 8 * kTile + 1024, stl>>>(tmq, tmd, pl, delta);
        B200_LAUNCH_CHECK();
        attn_bwd_long_kernel<1><<<grid_l, kAttnBwdThreads, 8 * kTile + 1024, stl>>>(tmq, tmd, pl, delta);
        B200_LAUNCH_CHECK();
        return 0;
    }
    B200_CHECK_ARG(dqkv && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                   "attention: dqkv / out null or misaligned");
    CUtensorMap tm, tmdo, tmo;
    if ((rc = make_tmap_bf16_2d(ctx, &tm, qkv, 3 * H * 64, B * S, 3 * H * 64, 64, static_cast<uint32_t>(S)))) return rc;
    if ((rc = make_tmap_bf16_2d(ctx, &tmdo, dout, H * 64, B * S, H * 64, 64, static_cast<uint32_t>(S)))) return rc;
    if ((rc = make_tmap_bf16_2d(ctx, &tmo, out, H * 64, B * S, H * 64, 64, static_cast<uint32_t>(S)))) return rc;
    AttnParams p{};
    p.o = static_cast<const __nv_bfloat16*>(out);
    p.lse = const_cast<float*>(lse);
    p.out = static_cast<__nv_bfloat16*>(dqkv);
    p.B = static_cast<int>(B);
    p.S = static_cast<int>(S);
    p.H = static_cast<int>(H);
    p.causal = causal ? 1 : 0;
    p.npad = static_cast<int>((S + 15) / 16 * 16);
    const bool big = S > 64;
    const int smem = bwd_smem(big, p.npad);
    const int per_sm = ctas_per_sm(smem, 256, 2);
    const int64_t work = B * H;
    const int grid = static_cast<int>(work < ctx->num_sms * per_sm ? work : ctx->num_sms * per_sm);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (big)
        B200_CHECK_CUDA(launch_pdl(attn_bwd_kernel<true>, dim3(grid), dim3(kAttnBwdThreads), smem, st, tm, tmdo, tmo, p));
    else
        B200_CHECK_CUDA(launch_pdl(attn_bwd_kernel<false>, dim3(grid), dim3(kAttnBwdThreads), smem, st, tm, tmdo, tmo, p));
    B200_LAUNCH_CHECK();
    return 0;
}
