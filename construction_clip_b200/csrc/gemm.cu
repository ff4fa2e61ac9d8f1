// bf16 tensor-core GEMM for sm_100a: TMA -> 128B-swizzled shared-memory ring -> tcgen05.mma
// (accumulator in TMEM, double buffered) -> tcgen05.ld -> fused epilogue.
//
// Replaces nn.Linear / F.linear inside clip.model.ResidualAttentionBlock (attn.in_proj,
// attn.out_proj, mlp.c_fc + QuickGELU, mlp.c_proj + residual), visual.conv1 (after im2col),
// `x @ visual.proj`, `x @ text_projection` and their autograd dgrad / wgrad
// (reference call sites: CLIP/train.py:161 forward, CLIP/train.py:168 backward).
//
// Two kernels share the epilogue: gemm_pair_bf16_kernel (cta_group::2, one 256 x 256 tile per cluster of
// two CTAs, the default for large problems) and gemm_bf16_kernel<BN> (one CTA, 128 x {256,192,128}
// tiles); the launcher picks the shape with the smallest waves x tile-time makespan.
// One persistent CTA per SM, 10 warps:
//   warp 0     TMA producer (one elected lane)
//   warp 1     TMEM allocator + MMA issuer (one elected lane)
//   warps 2-9  epilogue: warp w drains TMEM lane quadrant w%4 (32 rows of the 128-row tile) for one
//              half of the tile's columns, transposing through shared memory so that every global
//              access (aux load, C store) is a full 128-byte row segment
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty pair (MMA <-> epilogue).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr int BM = 128;           // tile rows  (UMMA M, cta_group::1)
constexpr int BK = 64;            // k-block: 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;        // fixed for 16-bit inputs
// Epilogue scheduling knobs (compile-time; tools/build_variants.py builds A/B variants of the library).
// Same-box A/B over the 24 GEMM shapes of a ViT-B/32 layer (per-layer total, us; 3495 before any of them):
//   all off (only the per-tile bias fetch)            3449   <- default
//   LDTM look-ahead + early release                   3487
//   full-tile straight-line path                      3553
//   full-tile path + look-ahead + early release       3530
// The straight-line path halves the instruction count but bursts its 8 row stores; the predicated
// per-row form spreads them and is faster -- the epilogue is bound by the latency of its global
// accesses (8 warps issue all of an SM's output traffic), not by issue slots.
#ifndef B200_EPI_LDTM_AHEAD
#define B200_EPI_LDTM_AHEAD 0   // issue block j+1's TMEM load as soon as block j is staged
#endif
#ifndef B200_EPI_FULL_SPEC
#define B200_EPI_FULL_SPEC 0    // straight-line code path for tiles entirely inside the matrix
#endif
#ifndef B200_EPI_AUX_PREFETCH
#define B200_EPI_AUX_PREFETCH 1   // L2-prefetch the next tile's aux (residual / pre-activation) rows
#endif
#ifndef B200_EPI_EARLY_RELEASE
#define B200_EPI_EARLY_RELEASE 0  // hand the TMEM stage back after the last TMEM load, not the last store
#endif
// Epilogue warps: kEpiParts per TMEM lane quadrant, each takes a share of the tile's 32-column blocks.
// Measured with 12 warps (-DB200_EPI_WARPS=12: 3 column shares, one operand stage less to pay for their staging
// tiles, 128 registers per thread), same box, ViT-B/32 layer shapes at 1024 pairs: per-layer total 2730 us against
// 2682 us with 8 -- worse on 22 of 24 shapes (c_fc fwd 235 vs 227 us, out_proj fwd 93 vs 87), better only on the
// text c_proj dgrad (127 vs 137).  The fat epilogues are not short of warps.
#ifndef B200_EPI_WARPS
#define B200_EPI_WARPS 8
#endif
constexpr int kEpiWarps = B200_EPI_WARPS;
constexpr int kEpiParts = kEpiWarps / 4;
static_assert(kEpiWarps % 4 == 0 && kEpiParts >= 2 && kEpiParts <= 4, "epilogue warps come in groups of four (TMEM lane quadrants)");
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kStageBytesA = BM * BK * 2;  // 16 KB
constexpr int kEpiStageBytes = 32 * 128;   // per-warp staging: 32 rows x 128 B, XOR-swizzled

// Tile widths BN in {96, 128, 160, 192, 224, 256}: with M / 128 row tiles fixed by the batch, the width is what
// lets the launcher fit the tile count to the 148 SMs (e.g. 6400 x 768: 150 tiles of 256 = 2 waves, 250 tiles
// of 160 = 1.7 waves of 0.7-cost tiles).  Shared-memory stages are sized for BN rounded up to 64: an MN-major
// B operand arrives in 64-column swizzle atoms (the last atom of a 96 / 160 / 224 tile is loaded whole and
// read half).
template <int BN>
struct GemmCfg {
    static_assert(BN % 32 == 0 && BN >= 64 && BN <= 256, "tile width");
    static constexpr int kBNSmem = (BN + 63) / 64 * 64;
    static constexpr int kStageBytesB = kBNSmem * BK * 2;
    static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
    static constexpr int kFit = (kSmemBudget - 1024 - kEpiWarps * kEpiStageBytes) / kStageBytes;  // ring depth that fits
    static constexpr int kCap = (kBNSmem >= 192) ? 4 : 6;
    static constexpr int kStages = kFit < kCap ? kFit : kCap;
    static_assert(kStages >= 3, "operand ring too shallow");
    static constexpr int kTmemCols = (2 * BN > 256) ? 512 : 256;  // two accumulator stages (power of two)
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps * kEpiStageBytes + 1024;  // + align slack
};

struct GemmParams {
    void* C;
    const __nv_bfloat16* bias;
    const __nv_bfloat16* aux;
    __nv_bfloat16* preact;
    const float* scale;
    float* colsum;  // optional fp32 [N]: += column sums of C (bias gradient of the layer that produced A's grad)
    int64_t ldc, ldaux;
    int M, N, K;
    int num_m_tiles, num_n_tiles;
    int split_k, kb_total, kb_per_split;
    int out_f32, atomic_out, epilogue;
    int tma_store;  // EPI_QUICKGELU only: row-layout epilogue, outputs leave through TMA stores (tmC / tmP are valid)
    int d8;         // EPI_QUICKGELU: `preact` receives 8-bit codes of QuickGELU'(x) instead of x; EPI_QUICKGELU_BWD: `aux` holds them
};

// ------------------------------------------------------------------------------------------------
// Epilogue.  TMEM hands each thread one accumulator ROW (lane = row); writing rows straight to
// global memory would make every warp-level access touch 32 different lines.  Each epilogue warp
// therefore transposes 32 x 32 fp32 accumulator blocks through a private 4 KB shared-memory tile
// (16-byte chunks XOR-swizzled with the row: both access patterns are bank-conflict free) and does
// ALL the epilogue maths in the coalesced layout, where a lane owns 4 consecutive columns of 8
// different rows: bias is loaded once per block into registers, aux (fp32 residual stream or the
// saved bf16 pre-activation) is read with coalesced loads issued before the transpose, and the
// result leaves as 8/16-byte stores that cover whole 32-byte sectors.
// The tcgen05.ld of a block is in flight while its aux / bias loads are issued.
__device__ __forceinline__ uint32_t stage_off(int row, int chunk) {
    return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// x * sigmoid(1.702 x) = x * rcp(1 + 2^(-1.702 log2(e) x)) with ex2.approx / rcp.approx: 5 instructions,
// no range-check branches (the CUDA __fdividef / __expf pair costs ~3x as many), relative error
// ~1e-6.  The FORWARD uses this accurate form -- tanh.approx (2^-11) measurably eats into the 1e-2
// logit tolerance; the backward's QuickGELU' below can afford it.  x -> -inf gives x * 0, x -> +inf gives x.
__device__ __forceinline__ float qgelu_fast(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-2.4554669595930157f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}
// acc * d/dx[x sigmoid(1.702x)] = 0.5 acc (1 + t + u (1 - t^2)), u = 0.851 x, t = tanh(u)   (6 instr.)
__device__ __forceinline__ float qgelu_bwd_fast(float acc, float x) {
#ifdef B200_QGELU_BWD_EX2  // s = sigmoid(1.702 x) by ex2 + rcp:  acc * s * (1 + 1.702 x (1 - s))
    float e, sg;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-2.4554669595930157f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.0f + e));
    const float t1 = 1.702f * x;
    const float w = fmaf(-t1, sg, t1);  // 1.702 x (1 - s)
    const float as = acc * sg;
    return fmaf(as, w, as);
#endif
    const float u = 0.851f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float a = fmaf(-t, t, 1.0f);
    const float c = fmaf(u, a, t);
    const float h = 0.5f * acc;
    return fmaf(h, c, h);
}

// QuickGELU'(x) = s + 1.702 x s (1 - s), s = sigmoid(1.702 x), lives in [-0.1008, 1.1008].  The backward needs nothing
// else of the pre-activation, so the forward can save the derivative itself on an 8-bit grid instead of x in bf16:
// code = round(210 g' + 22), g' = (code - 22) / 210 -- step 1 / 210 (0 and 1, the two tails, are exact grid points),
// rms error 1.4e-3, evaluated from the fp32 pre-activation.  That is about what re-evaluating g' from a bf16-rounded
// x costs (|g''| <= 0.85 times a relative 2^-9 of x), at half the bytes in both directions and without a MUFU in
// the backward.  (B200CLIP_EPI_QUICKGELU_D8 / _BWD_D8.)
constexpr float kD8Scale = 210.0f, kD8Zero = 22.0f;
__device__ __forceinline__ uint32_t qgelu_fwd_d8(float x, float& g) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-2.4554669595930157f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    g = x * r;
    const float gp = fmaf(1.702f * g, 1.0f - r, r);
    // round to nearest through the magic-number add (1.5 * 2^23): the code lands in the low mantissa byte.  No F2I: the
    // conversion unit is the one the ex2 / rcp above already keep busy.  g' is inside [-0.1008, 1.1008] by construction
    // (the approximations are good to 1e-6), so 210 g' + 22 is inside [0.8, 253.2] and nothing has to saturate.
    return __float_as_uint(fmaf(gp, kD8Scale, kD8Zero + 12582912.0f)) & 0xffu;
}
__device__ __forceinline__ float d8_decode(uint32_t word, int k) {   // byte k of a little-endian word of four codes
    return fmaf(static_cast<float>((word >> (8 * k)) & 0xffu), 1.0f / kD8Scale, -kD8Zero / kD8Scale);
}

// per-lane aux operands of one 32x32 block (coalesced layout: 8 rows x 4 columns).  Fetched one block
// ahead so that global-memory latency is off the critical path (the first block of a tile is fetched
// before waiting for the MMA).  The bias of ALL the tile's blocks is fetched once per tile (drain_tile).
template <int EPI, bool OUT_F32>
struct EpiOperands {
    static constexpr bool AUX_F32 = (EPI == B200CLIP_EPI_RESIDUAL) && OUT_F32;
    static constexpr bool AUX_BF16 = (EPI == B200CLIP_EPI_QUICKGELU_BWD) || (EPI == B200CLIP_EPI_RESIDUAL && !OUT_F32);
    static constexpr bool AUX_U8 = (EPI == B200CLIP_EPI_QUICKGELU_BWD_D8);
    uint4 auxf[AUX_F32 ? 8 : 1];
    uint2 auxh[AUX_BF16 ? 8 : 1];
    uint32_t auxb[AUX_U8 ? 8 : 1];

    // Loads are UNCONDITIONAL from clamped (always valid) addresses so that all eight are in flight
    // at once -- a `cond ? load : 0` select puts a dependent MOV behind every load and serialises
    // them.  Values fetched for out-of-range rows / columns are never stored.
    // FULL: the whole 32 x 32 block is inside the matrix (warp-uniform), no clamping needed.
    template <bool FULL>
    __device__ __forceinline__ void load(const GemmParams& p, int m_base, int col0, int lane) {
        const int rrow = lane >> 3;
        int col = col0 + (lane & 7) * 4;
        if constexpr (!FULL) col = col < p.N ? col : 0;  // N % 8 == 0 and col % 4 == 0: the 4 columns are all in or all out
        if constexpr (AUX_F32) {
            const float* ap = reinterpret_cast<const float*>(p.aux) + col;
            if constexpr (FULL) {
                ap += static_cast<int64_t>(m_base + rrow) * p.ldaux;
                const int64_t rs = 4 * p.ldaux;
#pragma unroll
                for (int i = 0; i < 8; ++i) auxf[i] = *reinterpret_cast<const uint4*>(ap + i * rs);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int grow = min(m_base + 4 * i + rrow, p.M - 1);
                    auxf[i] = *reinterpret_cast<const uint4*>(ap + static_cast<int64_t>(grow) * p.ldaux);
                }
            }
        }
        if constexpr (AUX_U8) {
            const uint8_t* ap = reinterpret_cast<const uint8_t*>(p.aux) + col;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int grow = FULL ? m_base + 4 * i + rrow : min(m_base + 4 * i + rrow, p.M - 1);
                auxb[i] = *reinterpret_cast<const uint32_t*>(ap + static_cast<int64_t>(grow) * p.ldaux);
            }
        }
        if constexpr (AUX_BF16) {
            const __nv_bfloat16* ap = p.aux + col;
            if constexpr (FULL) {
                ap += static_cast<int64_t>(m_base + rrow) * p.ldaux;
                const int64_t rs = 4 * p.ldaux;
#pragma unroll
                for (int i = 0; i < 8; ++i) auxh[i] = *reinterpret_cast<const uint2*>(ap + i * rs);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int grow = min(m_base + 4 * i + rrow, p.M - 1);
                    auxh[i] = *reinterpret_cast<const uint2*>(ap + static_cast<int64_t>(grow) * p.ldaux);
                }
            }
        }
    }
};

// FULL (warp-uniform): all 32 rows and 32 columns of the block are inside the matrix, so the bounds
// predicates and the branches around the maths disappear; PRE: the pre-activation is stored too.
// Both are compile-time so that the hot loop is straight-line code (measured: the generic form spent
// ~15 instructions per element, 6 of them on predicates / pointer selects / re-materialised descriptors).
template <int EPI, bool OUT_F32, bool ATOMIC, bool FULL, bool PRE, bool PAIR>
__device__ __forceinline__ void epilogue_block(const GemmParams& p, const EpiOperands<EPI, OUT_F32>& op, uint2 bias_pk,
                                               uint32_t (&acc)[32], uint32_t taddr_cur, uint32_t taddr_next,
                                               uint64_t* release_bar, int m_base, int col0, float scale, uint8_t* stg,
                                               int lane) {
    using Op = EpiOperands<EPI, OUT_F32>;
    using OutT = typename std::conditional<OUT_F32, float, __nv_bfloat16>::type;
    // column sums of C ride along in the QuickGELU' epilogue (c_fc bias gradient) and in the plain bf16
    // one (out_proj dgrad: the V third of in_proj_bias' gradient is the column sum of d(attention out))
    constexpr bool kColsum = (EPI == B200CLIP_EPI_QUICKGELU_BWD) || (EPI == B200CLIP_EPI_QUICKGELU_BWD_D8) ||
                             (EPI == B200CLIP_EPI_NONE && !OUT_F32);
    const int rrow = lane >> 3, rch = lane & 7;
    const int col = col0 + rch * 4;
    const bool col_ok = FULL ? true : col < p.N;
    const float2 bf0 = unpack_bf16(bias_pk.x), bf1 = unpack_bf16(bias_pk.y);
#if !B200_EPI_LDTM_AHEAD
    tmem_ld_32x32(taddr_cur, acc);
#endif
    tmem_ld_wait();  // (look-ahead: this block's accumulator load was issued during the previous block)
    tc_fence_before();
#if defined(B200_EPI_DRY) && B200_EPI_DRY == 1  // measurement only: drain TMEM and drop the block
    __syncwarp();
    if (taddr_next == 0u && lane == 0) {
        if constexpr (PAIR)
            mbar_arrive_cluster(mapa_shared(smem_u32(release_bar), cluster_ctarank() & ~1u));
        else
            mbar_arrive(release_bar);
    }
    return;
#endif
    // ---- transpose: row layout -> staging
#pragma unroll
    for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(stg + stage_off(lane, c)) = make_uint4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
    __syncwarp();
    // the accumulator registers are free again: fetch the next block under this block's maths, or, after
    // the tile's last block, hand the TMEM stage back to the MMA warp right away
    if (taddr_next != 0u) {
#if B200_EPI_LDTM_AHEAD
        tmem_ld_32x32(taddr_next, acc);
#endif
    }
#if B200_EPI_EARLY_RELEASE
    else if (lane == 0) {  // every lane's loads completed before the __syncwarp above
        if constexpr (PAIR)
            mbar_arrive_cluster(mapa_shared(smem_u32(release_bar), cluster_ctarank() & ~1u));  // the leader's barrier
        else
            mbar_arrive(release_bar);
    }
#endif
    // ---- coalesced layout: maths + stores.  Row pointers advance by a constant stride (no per-row
    // 64-bit multiply); rows_left turns the row bound into a compare against the unrolled index.
    const int64_t first = static_cast<int64_t>(m_base + rrow) * p.ldc + col;
    OutT* cptr = reinterpret_cast<OutT*>(p.C) + first;
    [[maybe_unused]] __nv_bfloat16* pptr = PRE ? p.preact + first : nullptr;
    const int64_t rstride = 4 * p.ldc;
    // number of valid i (rows m_base+rrow+4i < M)
    [[maybe_unused]] const int rows_left = FULL ? 8 : (col_ok ? (p.M - m_base - rrow + 3) >> 2 : 0);
    [[maybe_unused]] float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f, cs3 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = 4 * i + rrow;
        const uint4 v = *reinterpret_cast<const uint4*>(stg + stage_off(row, rch));
#if defined(B200_EPI_DRY) && B200_EPI_DRY == 2  // measurement only: staging round trip, no maths / global stores
        if (p.M < 0) *reinterpret_cast<uint4*>(cptr) = v;
        continue;
#endif
        float x0 = fmaf(__uint_as_float(v.x), scale, bf0.x), x1 = fmaf(__uint_as_float(v.y), scale, bf0.y);
        float x2 = fmaf(__uint_as_float(v.z), scale, bf1.x), x3 = fmaf(__uint_as_float(v.w), scale, bf1.y);
#if defined(B200_EPI_DRY) && B200_EPI_DRY == 3  // measurement only: everything but the global stores
        const bool ok = (i < rows_left) && (p.M < 0);
#else
        const bool ok = FULL ? true : i < rows_left;
#endif
        if constexpr (EPI == B200CLIP_EPI_QUICKGELU) {
            if constexpr (PRE) {
                if (ok) *reinterpret_cast<uint2*>(pptr) = make_uint2(pack_bf16(x0, x1), pack_bf16(x2, x3));
            }
            x0 = qgelu_fast(x0); x1 = qgelu_fast(x1); x2 = qgelu_fast(x2); x3 = qgelu_fast(x3);
        } else if constexpr (Op::AUX_U8) {
            x0 *= d8_decode(op.auxb[i], 0); x1 *= d8_decode(op.auxb[i], 1);
            x2 *= d8_decode(op.auxb[i], 2); x3 *= d8_decode(op.auxb[i], 3);
        } else if constexpr (Op::AUX_F32) {
            x0 += __uint_as_float(op.auxf[i].x); x1 += __uint_as_float(op.auxf[i].y);
            x2 += __uint_as_float(op.auxf[i].z); x3 += __uint_as_float(op.auxf[i].w);
        } else if constexpr (Op::AUX_BF16) {
            const float2 f0 = unpack_bf16(op.auxh[i].x), f1 = unpack_bf16(op.auxh[i].y);
            if constexpr (EPI == B200CLIP_EPI_RESIDUAL) {
                x0 += f0.x; x1 += f0.y; x2 += f1.x; x3 += f1.y;
            } else {
                x0 = qgelu_bwd_fast(x0, f0.x); x1 = qgelu_bwd_fast(x1, f0.y);
                x2 = qgelu_bwd_fast(x2, f1.x); x3 = qgelu_bwd_fast(x3, f1.y);
            }
        }
        if (ok) {
            if constexpr (kColsum) {  // flavours that can carry a fused column sum (bias gradients)
                cs0 += x0; cs1 += x1; cs2 += x2; cs3 += x3;
            }
            if constexpr (OUT_F32) {
                if constexpr (ATOMIC)
                    red_add_v4(cptr, x0, x1, x2, x3);
                else
                    *reinterpret_cast<float4*>(cptr) = make_float4(x0, x1, x2, x3);
            } else {
                *reinterpret_cast<uint2*>(cptr) = make_uint2(pack_bf16(x0, x1), pack_bf16(x2, x3));
            }
        }
        cptr += rstride;
        if constexpr (PRE) pptr += rstride;
    }
    if constexpr (kColsum) {
        if (p.colsum != nullptr) {  // warp-uniform
            // lanes l, l^8, l^16, l^24 hold the same 4 columns for different rows
            cs0 += __shfl_xor_sync(0xffffffffu, cs0, 8);  cs1 += __shfl_xor_sync(0xffffffffu, cs1, 8);
            cs2 += __shfl_xor_sync(0xffffffffu, cs2, 8);  cs3 += __shfl_xor_sync(0xffffffffu, cs3, 8);
            cs0 += __shfl_xor_sync(0xffffffffu, cs0, 16); cs1 += __shfl_xor_sync(0xffffffffu, cs1, 16);
            cs2 += __shfl_xor_sync(0xffffffffu, cs2, 16); cs3 += __shfl_xor_sync(0xffffffffu, cs3, 16);
#ifdef B200_EPI_NO_COLSUM_RED  // measurement only
            if (lane < 8 && col_ok && p.M < 0) red_add_v4(p.colsum + col, cs0, cs1, cs2, cs3);
#else
            if (lane < 8 && col_ok) red_add_v4(p.colsum + col, cs0, cs1, cs2, cs3);
#endif
        }
    }
    __syncwarp();  // the staging tile is rewritten by the next block
#if !B200_EPI_EARLY_RELEASE
    if (taddr_next == 0u && lane == 0) {
        if constexpr (PAIR)
            mbar_arrive_cluster(mapa_shared(smem_u32(release_bar), cluster_ctarank() & ~1u));
        else
            mbar_arrive(release_bar);
    }
#endif
}

// One warp's share of one tile (nblk <= 4 blocks of 32 x 32), specialised on "entirely inside the
// matrix" (FULL) and "stores the pre-activation" (PRE).  Per-block latencies are taken off the
// critical path: the bias of every block is fetched before the accumulator is even complete, aux
// operands one block ahead, the TMEM load of block j+1 is issued as soon as block j is staged, and
// the TMEM stage is released right after the last TMEM load instead of after the last store.
template <int EPI, bool OUT_F32, bool ATOMIC, bool FULL, bool PRE, bool PAIR>
__device__ __forceinline__ void drain_tile(const GemmParams& p, uint64_t* full_bar, uint32_t phase,
                                           uint64_t* release_bar, uint32_t taddr, int m_base, int n_base, int nblk,
                                           float scale, uint8_t* stg, int lane) {
    uint2 bias_pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        bias_pk[j] = make_uint2(0u, 0u);
        if (p.bias != nullptr && j < nblk) {
            int col = n_base + j * 32 + (lane & 7) * 4;
            if constexpr (!FULL) col = col < p.N ? col : 0;
            bias_pk[j] = __ldg(reinterpret_cast<const uint2*>(p.bias + col));
        }
    }
    EpiOperands<EPI, OUT_F32> opA;
    if (nblk > 0) opA.template load<FULL>(p, m_base, n_base, lane);  // before waiting for the accumulator
    mbar_wait(full_bar, phase);
    __syncwarp();
    tc_fence_after();
    uint32_t acc[32];
    if (nblk > 0) {
#if B200_EPI_LDTM_AHEAD
        tmem_ld_32x32(taddr, acc);
#endif
    } else {  // nothing to drain (tile entirely past N): just hand the stage back
        tc_fence_before();
        if (lane == 0) {
            if constexpr (PAIR)
                mbar_arrive_cluster(mapa_shared(smem_u32(release_bar), cluster_ctarank() & ~1u));
            else
                mbar_arrive(release_bar);
        }
        return;
    }
#define B200_NEXT(j) taddr + (j) * 32, ((j) + 1 < nblk ? taddr + ((j) + 1) * 32 : 0u)
    if constexpr (!FULL && B200_EPI_FULL_SPEC) {  // edge tiles (rare): one operand set, no look-ahead
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < nblk) {
                if (j > 0) opA.template load<FULL>(p, m_base, n_base + j * 32, lane);
                epilogue_block<EPI, OUT_F32, ATOMIC, FULL, PRE, PAIR>(p, opA, bias_pk[j], acc, B200_NEXT(j), release_bar,
                                                                      m_base, n_base + j * 32, scale, stg, lane);
            }
        }
        return;
    }
    EpiOperands<EPI, OUT_F32> opB;  // ping-pong: the next block's operands load while this one runs
#pragma unroll
    for (int j = 0; j < 4; j += 2) {
        if (j < nblk) {
            if (j + 1 < nblk) opB.template load<FULL>(p, m_base, n_base + (j + 1) * 32, lane);
            epilogue_block<EPI, OUT_F32, ATOMIC, FULL, PRE, PAIR>(p, opA, bias_pk[j], acc, B200_NEXT(j), release_bar, m_base,
                                                                  n_base + j * 32, scale, stg, lane);
            if (j + 1 < nblk) {
                if (j + 2 < nblk) opA.template load<FULL>(p, m_base, n_base + (j + 2) * 32, lane);
                epilogue_block<EPI, OUT_F32, ATOMIC, FULL, PRE, PAIR>(p, opB, bias_pk[j + 1], acc, B200_NEXT(j + 1),
                                                                      release_bar, m_base, n_base + (j + 1) * 32, scale,
                                                                      stg, lane);
            }
        }
    }
#undef B200_NEXT
}

// ------------------------------------------------------------------------------------------------
// Row-layout epilogue with TMA stores, for the fattest epilogue of the model: c_fc forward = bias + QuickGELU with
// TWO bf16 outputs (activation and saved pre-activation).  The coalesced-layout path above spends 12.2 warp
// instructions per output element on this flavour (ncu source page, profiles/r02_gemm_fc_fwd_epilogue_sass.md): 6 of
// maths, the rest on the fp32 transpose through shared memory, per-row pointer / predicate arithmetic and 8-byte
// predicated stores -- and a tile takes 12.6 k cycles of epilogue against 6.1 k cycles of MMAs.  Here every thread
// keeps its accumulator ROW (the TMEM layout), does the maths in place, writes bf16 rows into a 64-byte-swizzled
// 32 x 32 staging tile (conflict free: chunk ^ ((row >> 1) & 3)) and one lane hands the tile to the TMA unit, which
// coalesces and clips it at the matrix edge: no transposition, no address arithmetic, no predicates, no LSU stores.
// The bias of a block is the same 32 values for every lane: four warp-uniform 16-byte loads.
// D8 (with PRE): the second output is not the pre-activation but the 8-bit code of QuickGELU'(x) (32 bytes per row of a
// block, 32-byte swizzle: chunk ^ ((row >> 2) & 1)).
template <bool PRE, bool D8, bool PAIR>
__device__ __forceinline__ void drain_tile_rows_qgelu(const GemmParams& p, const CUtensorMap* tmC, const CUtensorMap* tmP,
                                                      uint64_t* full_bar, uint32_t phase, uint64_t* release_bar,
                                                      uint32_t taddr, int m_base, int n_base, int nblk, float scale,
                                                      uint8_t* stg, int lane) {
    mbar_wait(full_bar, phase);
    __syncwarp();
    tc_fence_after();
    const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);
    uint8_t* row_g = stg + lane * 64;
    uint8_t* row_p = row_g + 2048;
#pragma unroll 1
    for (int j = 0; j < nblk; ++j) {
        const int cb = n_base + j * 32;
        uint32_t acc[32];
        tmem_ld_32x32(taddr + j * 32, acc);
        uint4 bq[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            bq[c] = make_uint4(0u, 0u, 0u, 0u);
            if (p.bias != nullptr && cb + 8 * c + 8 <= p.N)  // warp-uniform
                bq[c] = __ldg(reinterpret_cast<const uint4*>(p.bias + cb) + c);
        }
        tmem_ld_wait();
        if (j + 1 == nblk) {  // the accumulator stage is drained: hand it back before the maths
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR)
                    mbar_arrive_cluster(mapa_shared(smem_u32(release_bar), cluster_ctarank() & ~1u));
                else
                    mbar_arrive(release_bar);
            }
        }
        uint32_t pg[16];
        [[maybe_unused]] uint32_t pp[D8 ? 8 : 16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t bw[4] = {bq[c].x, bq[c].y, bq[c].z, bq[c].w};
            [[maybe_unused]] uint32_t codes[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 bf = unpack_bf16(bw[k]);
                const float x0 = fmaf(__uint_as_float(acc[8 * c + 2 * k]), scale, bf.x);
                const float x1 = fmaf(__uint_as_float(acc[8 * c + 2 * k + 1]), scale, bf.y);
                if constexpr (PRE && D8) {
                    float g0, g1;
                    codes[2 * k] = qgelu_fwd_d8(x0, g0);
                    codes[2 * k + 1] = qgelu_fwd_d8(x1, g1);
                    pg[4 * c + k] = pack_bf16(g0, g1);
                } else {
                    if constexpr (PRE) pp[4 * c + k] = pack_bf16(x0, x1);
                    pg[4 * c + k] = pack_bf16(qgelu_fast(x0), qgelu_fast(x1));
                }
            }
            if constexpr (PRE && D8) {  // eight codes -> two little-endian words
                pp[2 * c] = codes[0] | (codes[1] << 8) | (codes[2] << 16) | (codes[3] << 24);
                pp[2 * c + 1] = codes[4] | (codes[5] << 8) | (codes[6] << 16) | (codes[7] << 24);
            }
        }
        // the previous block's stores must have read the staging tile before it is overwritten
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t off = (static_cast<uint32_t>(c) ^ sw) << 4;
            *reinterpret_cast<uint4*>(row_g + off) = make_uint4(pg[4 * c], pg[4 * c + 1], pg[4 * c + 2], pg[4 * c + 3]);
            if constexpr (PRE && !D8)
                *reinterpret_cast<uint4*>(row_p + off) = make_uint4(pp[4 * c], pp[4 * c + 1], pp[4 * c + 2], pp[4 * c + 3]);
        }
        if constexpr (PRE && D8) {
            uint8_t* row_d = stg + 2048 + lane * 32;
            const uint32_t s32 = static_cast<uint32_t>((lane >> 2) & 1);
            *reinterpret_cast<uint4*>(row_d + ((0u ^ s32) << 4)) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
            *reinterpret_cast<uint4*>(row_d + ((1u ^ s32) << 4)) = make_uint4(pp[4], pp[5], pp[6], pp[7]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(tmC, stg, cb, m_base);
            if constexpr (PRE) tma_store_2d(tmP, stg + 2048, cb, m_base);
            tma_store_commit();
        }
    }
    if (nblk <= 0) {  // nothing to drain (tile entirely past N): just hand the stage back
        tc_fence_before();
        if (lane == 0) {
            if constexpr (PAIR)
                mbar_arrive_cluster(mapa_shared(smem_u32(release_bar), cluster_ctarank() & ~1u));
            else
                mbar_arrive(release_bar);
        }
    }
}

// The whole persistent loop of one epilogue warp, specialised on the epilogue flavour.
// CL > 1: the CTA is one of CL CTAs of a cluster (cta_group::2 pairs) working on a CL x 128-row tile; it owns rows
// [rank*128, rank*128+128) of the tile, iterates over the work list of its CLUSTER and releases the
// accumulator stage on the LEADER CTA's barrier.
template <int BN, int EPI, bool OUT_F32, bool ATOMIC, int CL = 1>
__device__ __forceinline__ void epilogue_loop(const GemmParams& p, uint32_t tmem_base, uint64_t* tmem_full_bar,
                                           uint64_t* tmem_empty_bar, uint8_t* stg, int warp, int lane, int num_work,
                                           [[maybe_unused]] const CUtensorMap* tmC = nullptr,
                                           [[maybe_unused]] const CUtensorMap* tmP = nullptr) {
    const int quad = warp & 3;         // TMEM lane quadrant this warp may access
    const int part = (warp - 2) >> 2;  // which share of the tile's columns this warp drains
    const float scale = (p.scale != nullptr) ? __ldg(p.scale) : 1.0f;
    int as = 0;
    uint32_t aphase = 0;
    constexpr bool PAIR = CL > 1;  // CL = CTAs per cluster: 1, 2 (one cta_group::2 pair) or 4 (two pairs stacked along M)
    const int rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
    const int w0 = static_cast<int>(blockIdx.x) / CL;
    const int wstep = static_cast<int>(gridDim.x) / CL;
    // The pre-activation operand of the QuickGELU' epilogue is streamed once with one block of look-ahead
    // (2 KB in flight per warp).  Each lane pulls its row of the NEXT tile's aux region into L2 one tile
    // ahead (plain prefetch.global.L2: -4 % on the two c_proj dgrad shapes).  Not for the fp32 residual
    // operand (measured +4 % there), and never with the bulk / TMA prefetch form (it queues behind the
    // mainloop's operand loads: up to 2.4x slower).
    constexpr bool kHasAux = (EPI == B200CLIP_EPI_QUICKGELU_BWD) || (EPI == B200CLIP_EPI_QUICKGELU_BWD_D8);
    constexpr int kAuxElem = (EPI == B200CLIP_EPI_RESIDUAL && OUT_F32) ? 4 : (EPI == B200CLIP_EPI_QUICKGELU_BWD_D8 ? 1 : 2);
    constexpr int kBlocks = BN / 32, kBase = kBlocks / kEpiParts, kRem = kBlocks % kEpiParts;
    constexpr int kBlocks0 = kBase + (kRem ? 1 : 0);                            // the largest share (<= 4)
    static_assert(kBlocks0 <= 4, "drain_tile handles at most four blocks per warp");
    const int blk0 = part * kBase + (part < kRem ? part : kRem);                 // first block of this warp's share
    const int nblk_half = kBase + (part < kRem ? 1 : 0);                         // blocks this warp drains per tile
    auto prefetch_aux = [&](int wq) {
        if constexpr (kHasAux && B200_EPI_AUX_PREFETCH) {
            if (wq < num_work) {
                const int tq = wq / p.split_k;
                const int row = (tq / p.num_n_tiles) * (CL * BM) + rank * BM + quad * 32 + lane;
                const int nb = (tq % p.num_n_tiles) * BN + blk0 * 32;
                const int ncol = min(nblk_half * 32, p.N - nb);  // N % 8 == 0: a multiple of 16 bytes either way
                if (row < p.M && ncol > 0) {
                    const uint8_t* a = reinterpret_cast<const uint8_t*>(p.aux) +
                                       (static_cast<int64_t>(row) * p.ldaux + nb) * kAuxElem;
#pragma unroll
                    for (int l = 0; l < kBlocks0 * 32 * kAuxElem / 128; ++l)
                        if (l * 128 < ncol * kAuxElem) prefetch_l2_line(a + l * 128);
                }
            }
        }
    };
    prefetch_aux(w0);
    for (int w = w0; w < num_work; w += wstep) {
        prefetch_aux(w + wstep);
        const int tile = w / p.split_k;
        const int m_base = (tile / p.num_n_tiles) * (CL * BM) + rank * BM + quad * 32;
        const int n_base = (tile % p.num_n_tiles) * BN + blk0 * 32;
        int nblk = nblk_half;
        const int valid = (p.N - n_base + 31) >> 5;  // blocks with at least one real column (warp-uniform)
        if (valid < nblk) nblk = valid < 0 ? 0 : valid;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                               static_cast<uint32_t>(as * BN + blk0 * 32);
        const bool full = B200_EPI_FULL_SPEC && (m_base + 32 <= p.M) && (n_base + nblk * 32 <= p.N);
#define B200_DRAIN(FULL, PRE)                                                                                      \
    drain_tile<EPI, OUT_F32, ATOMIC, FULL, PRE, PAIR>(p, &tmem_full_bar[as], aphase, &tmem_empty_bar[as], taddr, m_base, \
                                                      n_base, nblk, scale, stg, lane)
        if constexpr (EPI == B200CLIP_EPI_QUICKGELU) {
            if (p.tma_store) {  // kernel-uniform
                if (p.preact != nullptr && p.d8)
                    drain_tile_rows_qgelu<true, true, PAIR>(p, tmC, tmP, &tmem_full_bar[as], aphase, &tmem_empty_bar[as],
                                                            taddr, m_base, n_base, nblk, scale, stg, lane);
                else if (p.preact != nullptr)
                    drain_tile_rows_qgelu<true, false, PAIR>(p, tmC, tmP, &tmem_full_bar[as], aphase, &tmem_empty_bar[as],
                                                             taddr, m_base, n_base, nblk, scale, stg, lane);
                else
                    drain_tile_rows_qgelu<false, false, PAIR>(p, tmC, tmP, &tmem_full_bar[as], aphase, &tmem_empty_bar[as],
                                                              taddr, m_base, n_base, nblk, scale, stg, lane);
            } else if (p.preact != nullptr) {
                if (full) B200_DRAIN(true, true);
                else B200_DRAIN(false, true);
            } else {
                if (full) B200_DRAIN(true, false);
                else B200_DRAIN(false, false);
            }
        } else {
            if (full) B200_DRAIN(true, false);
            else B200_DRAIN(false, false);
        }
#undef B200_DRAIN
        if (++as == 2) {
            as = 0;
            aphase ^= 1u;
        }
    }
    if constexpr (EPI == B200CLIP_EPI_QUICKGELU) {
        // bulk stores in flight read this CTA's shared memory and must be complete before the grid is
        if (p.tma_store && lane == 0) tma_store_wait_all<0>();
    }
}

// ------------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP, const GemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int kStages = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[kStages];
    __shared__ uint64_t empty_bar[kStages];
    __shared__ uint64_t tmem_full_bar[2];
    __shared__ uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    // 128B swizzle needs 1024-byte aligned tiles
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        if (p.tma_store) {
            prefetch_tmap(&tmC);
            if (p.preact != nullptr) prefetch_tmap(&tmP);
        }
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], kEpiWarps);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_slot, Cfg::kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    griddep_launch_dependents();
    griddep_wait();  // everything above overlapped the previous kernel's tail; global memory from here on

    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int num_work = num_tiles * p.split_k;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
                const int split = w % p.split_k;
                const int tile = w / p.split_k;
                const int m0 = (tile / p.num_n_tiles) * BM;
                const int n0 = (tile % p.num_n_tiles) * BN;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    // bytes the loads below deliver: a K-major B box is exactly BN rows, MN-major B comes in 64-column atoms
                    mbar_arrive_expect_tx(&full_bar[stage], kStageBytesA + (B_MN ? Cfg::kBNSmem : BN) * BK * 2);
                    uint8_t* sA = smem + stage * Cfg::kStageBytes;
                    uint8_t* sB = sA + kStageBytesA;
                    if constexpr (!A_MN) {
                        tma_load_2d(sA, &tmA, &full_bar[stage], kb * BK, m0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j)
                            tma_load_2d(sA + j * (BK * 128), &tmA, &full_bar[stage], m0 + 64 * j, kb * BK);
                    }
                    if constexpr (!B_MN) {
                        tma_load_2d(sB, &tmB, &full_bar[stage], kb * BK, n0);
                    } else {
#pragma unroll
                        for (int j = 0; j < Cfg::kBNSmem / 64; ++j)
                            tma_load_2d(sB + j * (BK * 128), &tmB, &full_bar[stage], n0 + 64 * j, kb * BK);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
                const int split = w % p.split_k;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sA = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint32_t sB = sA + kStageBytesA;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // K-major : advance 16 elements (32 B) inside the 128-B swizzle row
                        // MN-major: advance 16 k-rows of 128 B
                        const uint64_t da = A_MN ? make_smem_desc_sw128(sA + k * (UMMA_K * 128), BK * 128, 1024)
                                                 : make_smem_desc_sw128(sA + k * (UMMA_K * 2), 16, 1024);
                        const uint64_t db = B_MN ? make_smem_desc_sw128(sB + k * (UMMA_K * 128), BK * 128, 1024)
                                                 : make_smem_desc_sw128(sB + k * (UMMA_K * 2), 16, 1024);
                        umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot once the MMAs have read it
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(&tmem_full_bar[as]);  // accumulator complete -> epilogue
                if (++as == 2) {
                    as = 0;
                    aphase ^= 1u;
                }
            }
        }
    } else {
        // ===================================== epilogue warps ===================================
        uint8_t* stg = smem + kStages * Cfg::kStageBytes + (warp - 2) * kEpiStageBytes;
#define EPI_LOOP(E, F32, AT) \
    epilogue_loop<BN, E, F32, AT>(p, tmem_base, tmem_full_bar, tmem_empty_bar, stg, warp, lane, num_work)
#define EPI_LOOP_TMA(E, F32, AT) \
    epilogue_loop<BN, E, F32, AT>(p, tmem_base, tmem_full_bar, tmem_empty_bar, stg, warp, lane, num_work, &tmC, &tmP)
        if (p.out_f32) {
            if (p.atomic_out)
                EPI_LOOP(B200CLIP_EPI_NONE, true, true);
            else if (p.epilogue == B200CLIP_EPI_RESIDUAL)
                EPI_LOOP(B200CLIP_EPI_RESIDUAL, true, false);
            else
                EPI_LOOP(B200CLIP_EPI_NONE, true, false);
        } else {
            switch (p.epilogue) {
                case B200CLIP_EPI_QUICKGELU: EPI_LOOP_TMA(B200CLIP_EPI_QUICKGELU, false, false); break;
                case B200CLIP_EPI_RESIDUAL: EPI_LOOP(B200CLIP_EPI_RESIDUAL, false, false); break;
                case B200CLIP_EPI_QUICKGELU_BWD: EPI_LOOP(B200CLIP_EPI_QUICKGELU_BWD, false, false); break;
                case B200CLIP_EPI_QUICKGELU_BWD_D8: EPI_LOOP(B200CLIP_EPI_QUICKGELU_BWD_D8, false, false); break;
                default: EPI_LOOP(B200CLIP_EPI_NONE, false, false); break;
            }
        }
#undef EPI_LOOP
#undef EPI_LOOP_TMA
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair GEMM (tcgen05 cta_group::2): two CTAs of a cluster compute one 256 x 256 tile.  Each CTA
// loads its own 128 rows of A and HALF of the B tile (128 of the 256 columns); the pair's tensor cores
// read both halves, so the operand bytes per flop fed through L2 / shared memory drop by a third
// against the 128 x 256 single-CTA tile -- the measured limiter of that kernel -- and a stage shrinks
// to 32 KB, which buys a 6-deep ring.  Only the leader CTA (cluster rank 0) issues MMAs; full barriers
// live in the leader (both CTAs' TMA transactions and producer arrivals are routed to it), empty /
// accumulator-full barriers are signalled in both CTAs with a multicast tcgen05.commit, and both
// epilogues release the accumulator on the leader's barrier.
constexpr int kStageBytesPair = 2 * (BM * BK * 2);  // A 16 KB + half of B 16 KB
constexpr int kStagesPairFit = (kSmemBudget - 1024 - kEpiWarps * kEpiStageBytes) / kStageBytesPair;
constexpr int kStagesPair = kStagesPairFit < 6 ? kStagesPairFit : 6;
constexpr int kSmemBytesPair = kStagesPair * kStageBytesPair + kEpiWarps * kEpiStageBytes + 1024;

// CL = 4 (gemm_quad_bf16_kernel): TWO pairs stacked along M share every B tile.  All these GEMMs run into the same
// wall -- the L2 slices deliver ~6300 B / clock to the whole chip (B300_MICROARCH.md "LTS throughput cap"; 11.0-11.4
// TB/s measured on every large shape of profiles/r01_gemm_shapes.txt once operand reads, aux reads and output
// writes are added up) -- so the bytes a tile pulls through L2 per flop are what sets the speed.  In a cluster of
// four, CTA (pair pp, half q) loads only a 64-column QUARTER of the B tile and the TMA unit multicasts it into
// both pairs' CTAs of the same half: 16 KB of A + 8 KB of B per CTA and k-block instead of 16 + 16 (-25 % of the L2
// reads).  Barrier protocol on top of the pair kernel's: a shared-memory slot of CTA X is also written by X's
// M-neighbour (rank X ^ 2), so its empty barrier counts the MMA commits of BOTH pair leaders (multicast to all four
// CTAs); the full barriers (in each pair leader) still count 2 producer arrivals and 64 KB of transactions, part of
// which now come from the neighbour pair's multicasts.
template <bool A_MN, bool B_MN, int CL>
__device__ __forceinline__ void gemm_cluster_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                                                  const CUtensorMap& tmP, const GemmParams& p) {
    static_assert(CL == 2 || CL == 4, "cluster of one or two cta_group::2 pairs");
    constexpr int BN = 256;
    constexpr int kStages = kStagesPair;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[kStages];
    __shared__ uint64_t empty_bar[kStages];
    __shared__ uint64_t tmem_full_bar[2];
    __shared__ uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());
    const int q = rank & 1;               // which half of the pair (rows and B columns)
    const int pp = rank >> 1;             // which pair of the cluster
    const bool leader = q == 0;           // issues the pair's MMAs, owns its full / accumulator-empty barriers
    const uint32_t leader_rank = static_cast<uint32_t>(rank & ~1);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        if (p.tma_store) {
            prefetch_tmap(&tmC);
            if (p.preact != nullptr) prefetch_tmap(&tmP);
        }
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 2);        // leader: its own arrive.expect_tx + the peer producer's arrive
            mbar_init(&empty_bar[s], CL / 2);  // multicast commit of every pair leader that reads data landing in this slot
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], 2 * kEpiWarps);  // leader: every epilogue warp of both CTAs of the pair
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(&tmem_base_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    griddep_launch_dependents();
    griddep_wait();  // everything above overlapped the previous kernel's tail; global memory from here on

    const int num_work = p.num_m_tiles * p.num_n_tiles * p.split_k;  // num_m_tiles counts (CL x 128)-row tiles
    const int w0 = static_cast<int>(blockIdx.x) / CL, wstep = static_cast<int>(gridDim.x) / CL;

    if (warp == 0) {
        // ================== TMA producer (every CTA: own A rows, own share of B) ==================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint16_t mc_mask = static_cast<uint16_t>(0x5u << q);  // CTAs (0, q) and (1, q)
            for (int w = w0; w < num_work; w += wstep) {
                const int split = w % p.split_k;
                const int tile = w / p.split_k;
                const int m0 = (tile / p.num_n_tiles) * (CL * BM) + rank * BM;
                const int n0 = (tile % p.num_n_tiles) * BN + q * (BN / 2);
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    if (leader)
                        mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytesPair);
                    else
                        mbar_arrive_cluster(mapa_shared(smem_u32(&full_bar[stage]), leader_rank));
                    uint8_t* sA = smem + stage * kStageBytesPair;
                    uint8_t* sB = sA + kStageBytesA;
                    if constexpr (!A_MN) {
                        tma_load_2d_2sm(sA, &tmA, &full_bar[stage], kb * BK, m0);
                    } else {
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            tma_load_2d_2sm(sA + j * (BK * 128), &tmA, &full_bar[stage], m0 + 64 * j, kb * BK);
                    }
                    if constexpr (CL == 2) {
                        if constexpr (!B_MN) {
                            tma_load_2d_2sm(sB, &tmB, &full_bar[stage], kb * BK, n0);
                        } else {
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                tma_load_2d_2sm(sB + j * (BK * 128), &tmB, &full_bar[stage], n0 + 64 * j, kb * BK);
                        }
                    } else {
                        // this CTA's 64-column quarter of the B tile, delivered to both pairs (8 KB at the same offset)
                        if constexpr (!B_MN)
                            tma_load_2d_2sm_mc(sB + pp * (BK * 128), &tmB, &full_bar[stage], kb * BK, n0 + 64 * pp, mc_mask);
                        else
                            tma_load_2d_2sm_mc(sB + pp * (BK * 128), &tmB, &full_bar[stage], n0 + 64 * pp, kb * BK, mc_mask);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================== MMA issuer (pair leaders only) ==================
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            constexpr uint16_t all_mask = static_cast<uint16_t>((1u << CL) - 1u);
            const uint16_t pair_mask = static_cast<uint16_t>(3u << (2 * pp));
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int w = w0; w < num_work; w += wstep) {
                const int split = w % p.split_k;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sA = smem_u32(smem + stage * kStageBytesPair);
                    const uint32_t sB = sA + kStageBytesA;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t da = A_MN ? make_smem_desc_sw128(sA + k * (UMMA_K * 128), BK * 128, 1024)
                                                 : make_smem_desc_sw128(sA + k * (UMMA_K * 2), 16, 1024);
                        const uint64_t db = B_MN ? make_smem_desc_sw128(sB + k * (UMMA_K * 128), BK * 128, 1024)
                                                 : make_smem_desc_sw128(sB + k * (UMMA_K * 2), 16, 1024);
                        umma_bf16_2sm(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit_2sm(&empty_bar[stage], all_mask);  // this pair is done with the slot, in EVERY CTA of the cluster
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit_2sm(&tmem_full_bar[as], pair_mask);  // accumulators complete -> both epilogues of the pair
                if (++as == 2) {
                    as = 0;
                    aphase ^= 1u;
                }
            }
        }
    } else {
        // ================== epilogue warps (every CTA drains its own 128 rows) ==================
        uint8_t* stg = smem + kStages * kStageBytesPair + (warp - 2) * kEpiStageBytes;
#define EPI_LOOP(E, F32, AT) \
    epilogue_loop<BN, E, F32, AT, CL>(p, tmem_base, tmem_full_bar, tmem_empty_bar, stg, warp, lane, num_work)
#define EPI_LOOP_TMA(E, F32, AT) \
    epilogue_loop<BN, E, F32, AT, CL>(p, tmem_base, tmem_full_bar, tmem_empty_bar, stg, warp, lane, num_work, &tmC, &tmP)
        if (p.out_f32) {
            if (p.atomic_out)
                EPI_LOOP(B200CLIP_EPI_NONE, true, true);
            else if (p.epilogue == B200CLIP_EPI_RESIDUAL)
                EPI_LOOP(B200CLIP_EPI_RESIDUAL, true, false);
            else
                EPI_LOOP(B200CLIP_EPI_NONE, true, false);
        } else {
            switch (p.epilogue) {
                case B200CLIP_EPI_QUICKGELU: EPI_LOOP_TMA(B200CLIP_EPI_QUICKGELU, false, false); break;
                case B200CLIP_EPI_RESIDUAL: EPI_LOOP(B200CLIP_EPI_RESIDUAL, false, false); break;
                case B200CLIP_EPI_QUICKGELU_BWD: EPI_LOOP(B200CLIP_EPI_QUICKGELU_BWD, false, false); break;
                case B200CLIP_EPI_QUICKGELU_BWD_D8: EPI_LOOP(B200CLIP_EPI_QUICKGELU_BWD_D8, false, false); break;
                default: EPI_LOOP(B200CLIP_EPI_NONE, false, false); break;
            }
        }
#undef EPI_LOOP
#undef EPI_LOOP_TMA
    }

    tc_fence_before();
    cluster_sync_all();  // peers' shared memory / TMEM must stay alive until every pair's MMAs and multicasts are done
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc_2sm(tmem_base, 512);
    }
}

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_pair_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP, const GemmParams p) {
    gemm_cluster_body<A_MN, B_MN, 2>(tmA, tmB, tmC, tmP, p);
}

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_quad_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP, const GemmParams p) {
    gemm_cluster_body<A_MN, B_MN, 4>(tmA, tmB, tmC, tmP, p);
}

template <bool A_MN, bool B_MN>
static int set_attr_pair() {
    cudaError_t e = cudaFuncSetAttribute(gemm_pair_bf16_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBytesPair);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(gemm pair): %s", cudaGetErrorString(e));
        return B200CLIP_ERR_CUDA;
    }
    return 0;
}

template <bool A_MN, bool B_MN>
static int set_attr_quad() {
    cudaError_t e = cudaFuncSetAttribute(gemm_quad_bf16_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBytesPair);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(gemm quad): %s", cudaGetErrorString(e));
        return B200CLIP_ERR_CUDA;
    }
    return 0;
}

// Output tensor maps of the TMA-store epilogue (p.tma_store): 32 x 32 bf16 boxes, 64-byte swizzle, clipped at M x N.
static int make_store_tmaps(b200clip_ctx* ctx, const GemmParams& p, CUtensorMap* tmC, CUtensorMap* tmP) {
    if (!p.tma_store) return 0;
    int rc = make_tmap_bf16_2d_sw(ctx, tmC, p.C, p.N, p.M, p.ldc, 32, 32, 64);
    if (rc == 0 && p.preact != nullptr)
        rc = p.d8 ? make_tmap_u8_2d_sw32(ctx, tmP, p.preact, p.N, p.M, p.ldc, 32, 32)   // 8-bit codes, same extents, pitch ldc BYTES
                  : make_tmap_bf16_2d_sw(ctx, tmP, p.preact, p.N, p.M, p.ldc, 32, 32, 64);
    return rc;
}

// Persistent grid size for `work` items on `units` SMs (or CTA pairs).  The makespan is ceil(work / units)
// waves whichever way the items are dealt, so the SMALLEST grid that still finishes in that many waves would
// leave SMs to the kernels of the other tower's stream (two-stream step, section 5 of DESIGN.md) instead of
// idling them through this kernel's partial last wave.  Measured at 128 pairs / GPU: 8.330 ms per step with it,
// 8.285 ms without -- no gain, so it is opt-in (B200CLIP_MIN_GRID=1) and the default grid is `units`.
static int persistent_grid(int64_t work, int units) {
    static const bool min_grid = [] {
        const char* e = getenv("B200CLIP_MIN_GRID");
        return e && atoi(e) != 0;
    }();
    if (work <= units) return static_cast<int>(work);
    if (!min_grid) return units;
    const int64_t waves = ceil_div(work, static_cast<int64_t>(units));
    return static_cast<int>(ceil_div(work, waves));
}

template <bool A_MN, bool B_MN>
static int launch_pair(b200clip_ctx* ctx, const void* A, int64_t lda, const void* B, int64_t ldb, GemmParams& p,
                       cudaStream_t stream) {
    CUtensorMap tmA, tmB;
    int rc;
    if (!A_MN)
        rc = make_tmap_bf16_2d(ctx, &tmA, A, p.K, p.M, lda, BK, BM);
    else
        rc = make_tmap_bf16_2d(ctx, &tmA, A, p.M, p.K, lda, 64, BK);
    if (rc) return rc;
    if (!B_MN)
        rc = make_tmap_bf16_2d(ctx, &tmB, B, p.K, p.N, ldb, BK, 128);  // this CTA's half of the 256 columns
    else
        rc = make_tmap_bf16_2d(ctx, &tmB, B, p.N, p.K, ldb, 64, BK);
    if (rc) return rc;
    CUtensorMap tmC{}, tmP{};
    if ((rc = make_store_tmaps(ctx, p, &tmC, &tmP))) return rc;
    p.num_m_tiles = static_cast<int>(ceil_div(p.M, 2 * BM));
    p.num_n_tiles = static_cast<int>(ceil_div(p.N, 256));
    const int64_t work = static_cast<int64_t>(p.num_m_tiles) * p.num_n_tiles * p.split_k;
    const int clusters = persistent_grid(work, ctx->num_sms / 2);
    B200_CHECK_CUDA(launch_pdl(gemm_pair_bf16_kernel<A_MN, B_MN>, dim3(2 * clusters), dim3(kGemmThreads),
                               kSmemBytesPair, stream, tmA, tmB, tmC, tmP, p));
    B200_LAUNCH_CHECK();
    return 0;
}

// two pairs per cluster, B multicast (gemm_quad_bf16_kernel): 512 x 256 tiles over ctx->max_quads clusters
template <bool A_MN, bool B_MN>
static int launch_quad(b200clip_ctx* ctx, const void* A, int64_t lda, const void* B, int64_t ldb, GemmParams& p,
                       cudaStream_t stream) {
    CUtensorMap tmA, tmB;
    int rc;
    if (!A_MN)
        rc = make_tmap_bf16_2d(ctx, &tmA, A, p.K, p.M, lda, BK, BM);
    else
        rc = make_tmap_bf16_2d(ctx, &tmA, A, p.M, p.K, lda, 64, BK);
    if (rc) return rc;
    if (!B_MN)
        rc = make_tmap_bf16_2d(ctx, &tmB, B, p.K, p.N, ldb, BK, 64);  // this CTA's quarter of the 256 columns
    else
        rc = make_tmap_bf16_2d(ctx, &tmB, B, p.N, p.K, ldb, 64, BK);
    if (rc) return rc;
    CUtensorMap tmC{}, tmP{};
    if ((rc = make_store_tmaps(ctx, p, &tmC, &tmP))) return rc;
    p.num_m_tiles = static_cast<int>(ceil_div(p.M, 4 * BM));
    p.num_n_tiles = static_cast<int>(ceil_div(p.N, 256));
    const int64_t work = static_cast<int64_t>(p.num_m_tiles) * p.num_n_tiles * p.split_k;
    const int clusters = persistent_grid(work, ctx->max_quads);
    B200_CHECK_CUDA(launch_pdl(gemm_quad_bf16_kernel<A_MN, B_MN>, dim3(4 * clusters), dim3(kGemmThreads),
                               kSmemBytesPair, stream, tmA, tmB, tmC, tmP, p));
    B200_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN>
static int set_attr() {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GemmCfg<BN>::kSmemBytes);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(gemm BN=%d): %s", BN, cudaGetErrorString(e));
        return B200CLIP_ERR_CUDA;
    }
    return 0;
}

int init_gemm(b200clip_ctx* ctx) {
    int rc = 0;
    if ((rc = set_attr<256, false, false>())) return rc;
    if ((rc = set_attr<256, false, true>())) return rc;
    if ((rc = set_attr<256, true, true>())) return rc;
    if ((rc = set_attr<192, false, false>())) return rc;
    if ((rc = set_attr<192, false, true>())) return rc;
    if ((rc = set_attr<192, true, true>())) return rc;
    if ((rc = set_attr<128, false, false>())) return rc;
    if ((rc = set_attr<128, false, true>())) return rc;
    if ((rc = set_attr<128, true, true>())) return rc;
    if ((rc = set_attr<224, false, false>())) return rc;
    if ((rc = set_attr<224, false, true>())) return rc;
    if ((rc = set_attr<160, false, false>())) return rc;
    if ((rc = set_attr<160, false, true>())) return rc;
    if ((rc = set_attr<96, false, false>())) return rc;
    if ((rc = set_attr<96, false, true>())) return rc;
    if ((rc = set_attr_pair<false, false>())) return rc;
    if ((rc = set_attr_pair<false, true>())) return rc;
    if ((rc = set_attr_pair<true, true>())) return rc;
    if ((rc = set_attr_quad<false, false>())) return rc;
    if ((rc = set_attr_quad<false, true>())) return rc;
    if ((rc = set_attr_quad<true, true>())) return rc;
    // how many 4-CTA clusters of the quad kernel the device can hold at once (GPCs whose SM count is not a
    // multiple of four strand one pair each); 0 disables the kernel
    ctx->max_quads = 0;
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(4 * (ctx->num_sms / 4));
        cfg.blockDim = dim3(kGemmThreads);
        cfg.dynamicSmemBytes = kSmemBytesPair;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, gemm_quad_bf16_kernel<false, false>, &cfg) == cudaSuccess && n > 0)
            ctx->max_quads = n < ctx->num_sms / 4 ? n : ctx->num_sms / 4;
        else
            (void)cudaGetLastError();
    }
    return 0;
}

template <int BN, bool A_MN, bool B_MN>
static int launch(b200clip_ctx* ctx, const void* A, int64_t lda, const void* B, int64_t ldb, GemmParams& p,
                  cudaStream_t stream) {
    CUtensorMap tmA, tmB;
    int rc;
    if (!A_MN)
        rc = make_tmap_bf16_2d(ctx, &tmA, A, p.K, p.M, lda, BK, BM);
    else
        rc = make_tmap_bf16_2d(ctx, &tmA, A, p.M, p.K, lda, 64, BK);
    if (rc) return rc;
    if (!B_MN)
        rc = make_tmap_bf16_2d(ctx, &tmB, B, p.K, p.N, ldb, BK, BN);
    else
        rc = make_tmap_bf16_2d(ctx, &tmB, B, p.N, p.K, ldb, 64, BK);
    if (rc) return rc;
    CUtensorMap tmC{}, tmP{};
    if ((rc = make_store_tmaps(ctx, p, &tmC, &tmP))) return rc;
    p.num_m_tiles = static_cast<int>(ceil_div(p.M, BM));
    p.num_n_tiles = static_cast<int>(ceil_div(p.N, BN));
    const int64_t work = static_cast<int64_t>(p.num_m_tiles) * p.num_n_tiles * p.split_k;
    const int grid = persistent_grid(work, ctx->num_sms);
    B200_CHECK_CUDA(launch_pdl(gemm_bf16_kernel<BN, A_MN, B_MN>, dim3(grid), dim3(kGemmThreads),
                               GemmCfg<BN>::kSmemBytes, stream, tmA, tmB, tmC, tmP, p));
    B200_LAUNCH_CHECK();
    return 0;
}

// pick split_k minimising the makespan (in k-blocks) over `sms` persistent CTAs
static int choose_split_k(int64_t tiles, int kb_total, int sms) {
    if (tiles >= sms) return 1;
    int best = 1;
    double best_cost = 1e30;
    const double epi = 6.0;  // epilogue (atomics) cost of one work item, in k-block units
    for (int s = 1; s <= 64 && s <= kb_total; ++s) {
        const int64_t waves = ceil_div(tiles * s, sms);
        const double cost = waves * (static_cast<double>(ceil_div(kb_total, s)) + epi);
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best = s;
        }
    }
    return best;
}

}  // namespace b200

using namespace b200;

extern "C" int b200clip_gemm_bf16(b200clip_ctx* ctx, const void* A, int64_t lda, int a_major, const void* B,
                                  int64_t ldb, int b_major, void* C, int64_t ldc, int out_dtype, const void* bias,
                                  const void* aux, int64_t ldaux, void* preact, const float* scale, float* colsum,
                                  int64_t M, int64_t N, int64_t K, int epilogue, int split_k, int accumulate,
                                  void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(A && B && C, "gemm: null operand");
    B200_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem %lld x %lld x %lld", (long long)M, (long long)N,
                   (long long)K);
    B200_CHECK_ARG(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm: extent too large");
    B200_CHECK_ARG(N % 8 == 0, "gemm: N (%lld) must be a multiple of 8", (long long)N);
    B200_CHECK_ARG(a_major == B200CLIP_MAJOR_K || a_major == B200CLIP_MAJOR_MN, "gemm: bad a_major");
    B200_CHECK_ARG(b_major == B200CLIP_MAJOR_K || b_major == B200CLIP_MAJOR_MN, "gemm: bad b_major");
    B200_CHECK_ARG(!(a_major == B200CLIP_MAJOR_MN && b_major == B200CLIP_MAJOR_K),
                   "gemm: (A MN-major, B K-major) is not instantiated");
    B200_CHECK_ARG(out_dtype == B200CLIP_DT_BF16 || out_dtype == B200CLIP_DT_F32, "gemm: bad out_dtype");
    B200_CHECK_ARG(epilogue >= 0 && epilogue <= 5, "gemm: bad epilogue %d", epilogue);
    // the two 8-bit-derivative flavours: the forward is EPI_QUICKGELU whose second output is the code of QuickGELU'(x)
    const bool fwd_d8 = epilogue == B200CLIP_EPI_QUICKGELU_D8;
    if (fwd_d8) {
        epilogue = B200CLIP_EPI_QUICKGELU;
        B200_CHECK_ARG(preact != nullptr && out_dtype == B200CLIP_DT_BF16 && ldc % 16 == 0,
                       "gemm: EPI_QUICKGELU_D8 needs the code output, a bf16 C and ldc %% 16 == 0");
    }
    const bool out_f32 = out_dtype == B200CLIP_DT_F32;
    if (out_f32) {
        B200_CHECK_ARG(epilogue == B200CLIP_EPI_NONE || epilogue == B200CLIP_EPI_RESIDUAL,
                       "gemm: fp32 output supports EPI_NONE / EPI_RESIDUAL only");
        B200_CHECK_ARG(epilogue == B200CLIP_EPI_NONE || (split_k == 1 && !accumulate),
                       "gemm: EPI_RESIDUAL with fp32 output needs split_k = 1 and no accumulate");
        B200_CHECK_ARG(ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, "gemm: fp32 C not 16B aligned");
    } else {
        B200_CHECK_ARG(split_k <= 1 && !accumulate, "gemm: split_k / accumulate need fp32 output");
        B200_CHECK_ARG(ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, "gemm: bf16 C not 16B aligned");
    }
    if (epilogue == B200CLIP_EPI_RESIDUAL || epilogue == B200CLIP_EPI_QUICKGELU_BWD) {
        B200_CHECK_ARG(aux != nullptr, "gemm: epilogue %d needs aux", epilogue);
        B200_CHECK_ARG(ldaux % (out_f32 ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0,
                       "gemm: aux not 16B aligned");
    }
    if (epilogue == B200CLIP_EPI_QUICKGELU_BWD_D8) {
        B200_CHECK_ARG(aux != nullptr && !out_f32, "gemm: EPI_QUICKGELU_BWD_D8 needs the 8-bit aux and a bf16 C");
        B200_CHECK_ARG(ldaux % 16 == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0, "gemm: 8-bit aux not 16B aligned");
    }
    B200_CHECK_ARG(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "gemm: bias not 16B aligned");
    B200_CHECK_ARG(preact == nullptr || (reinterpret_cast<uintptr_t>(preact) & 15) == 0, "gemm: preact misaligned");
    B200_CHECK_ARG(colsum == nullptr || ((reinterpret_cast<uintptr_t>(colsum) & 15) == 0 && !out_f32 &&
                                         (epilogue == B200CLIP_EPI_QUICKGELU_BWD || epilogue == B200CLIP_EPI_QUICKGELU_BWD_D8 ||
                                          epilogue == B200CLIP_EPI_NONE)),
                   "gemm: colsum is fused into the bf16 EPI_NONE / EPI_QUICKGELU_BWD epilogues only, 16-byte aligned");

    GemmParams p{};
    p.C = C;
    p.bias = static_cast<const __nv_bfloat16*>(bias);
    p.aux = static_cast<const __nv_bfloat16*>(aux);
    p.preact = static_cast<__nv_bfloat16*>(preact);
    p.scale = scale;
    p.colsum = colsum;
    p.ldc = ldc;
    p.ldaux = ldaux;
    p.M = static_cast<int>(M);
    p.N = static_cast<int>(N);
    p.K = static_cast<int>(K);
    p.kb_total = static_cast<int>(ceil_div(K, BK));
    p.out_f32 = out_f32 ? 1 : 0;
    p.epilogue = epilogue;
    // c_fc forward (bias + QuickGELU, activation + saved pre-activation): row-layout epilogue with TMA stores
    // (B200CLIP_EPI_TMA=0: the coalesced-layout path with LSU stores, for A/B measurements)
    static const bool epi_tma = [] {
        const char* e = getenv("B200CLIP_EPI_TMA");
        return !(e && atoi(e) == 0);
    }();
    p.tma_store = ((epi_tma || fwd_d8) && epilogue == B200CLIP_EPI_QUICKGELU && !out_f32) ? 1 : 0;
    p.d8 = fwd_d8 ? 1 : 0;

    // CTA-pair kernel (256 x 256 tiles over 74 clusters) when the problem has enough such tiles;
    // B200CLIP_GEMM_PAIR=0 disables it (tuning / bisecting)
    static const bool pair_enabled = [] {
        const char* e = getenv("B200CLIP_GEMM_PAIR");
        return !(e && atoi(e) == 0);
    }();
    // Two pairs per cluster with the B tile multicast (gemm_quad_bf16_kernel) is OFF by default: measured on the 24
    // GEMM shapes of a ViT-B/32 layer at 1024 pairs it LOSES 3 % (2769 vs 2685 us per layer; fc wgrad 203 vs 173 us)
    // although it pulls 25 % fewer operand bytes through L2 -- four CTAs that must all retire a slot before any of
    // them refills it, and 36 clusters instead of 74.  B200CLIP_GEMM_QUAD=1 enables it (read per call: tests flip it).
    const bool quad_enabled = [] {
        const char* e = getenv("B200CLIP_GEMM_QUAD");
        return e && atoi(e) != 0;
    }();
    const int clusters = ctx->num_sms / 2;
    const int sms = ctx->num_sms;
    const int64_t m_tiles = ceil_div(M, BM);
    const int64_t tiles_pair = ceil_div(M, 2 * BM) * ceil_div(N, 256);
    const int64_t tiles_quad = ceil_div(M, 4 * BM) * ceil_div(N, 256);
    const bool pair_ok = pair_enabled && N >= 256 && M >= 2 * BM;
    const bool quad_ok = pair_ok && quad_enabled && ctx->max_quads > 0 && M >= 4 * BM;
    bool use_quad = false;
    const bool amn = a_major == B200CLIP_MAJOR_MN, bmn = b_major == B200CLIP_MAJOR_MN;
    auto tiles_of = [&](int bn) { return m_tiles * ceil_div(N, bn); };
    // Tile shape.  Atomic (split-K / accumulating) fp32 outputs -- the weight gradients -- take the widest
    // tile that exists for them and balance the machine through split-K instead.  Everything else has ONE
    // work item per tile, so the shape with the smallest makespan  waves x tile-time  wins: small per-GPU
    // batches (8-GPU strong scaling) leave the 256 x 256 pair tiles with a nearly empty last wave (75 tiles
    // on 74 clusters = 2 waves), and with the row count fixed by the batch the tile WIDTH is the free
    // parameter.  Tile time relative to a 128 x 256 tile on one SM: 0.24 + 0.76 * BN / 256 (measured 1.00 /
    // 0.80 / 0.62 at 256 / 192 / 128 on the ViT-B/32 layer shapes); the pair kernel's 256 x 256 tile on two
    // SMs costs 0.95.  B200CLIP_FORCE_BN=<width> pins the single-CTA width (tuning / experiments only).
    bool use_pair = false;
    int bn = 256;
    const char* force_bn = getenv("B200CLIP_FORCE_BN");
    const bool atomic_like = out_f32 && (split_k != 1 || accumulate);
    if (force_bn) {
        bn = atoi(force_bn);
        if (bn != 96 && bn != 128 && bn != 160 && bn != 192 && bn != 224 && bn != 256) bn = 256;
        if (amn && (bn % 64)) bn = 256;
        while (bn > 96 && bn > N + 31) bn -= 32;
    } else if (atomic_like) {
        use_pair = pair_ok;
        // weight gradients balance the machine through split-K, so the quad kernel only has to avoid padding M much
        use_quad = quad_ok && (ceil_div(M, 4 * BM) * 4 * BM - M) * 10 <= M;
        bn = N >= 256 ? 256 : (N >= 192 ? 192 : 128);
    } else {
        double best = pair_ok ? 0.95 * static_cast<double>(ceil_div(tiles_pair, clusters)) : 1e30;
        use_pair = pair_ok;
        // a 512 x 256 quad tile takes kQuadCost of a pair tile's time (its two pairs run side by side on 3/4 of the L2 reads)
        static const double quad_cost = [] {
            const char* e = getenv("B200CLIP_QUAD_COST");
            return e ? atof(e) : 0.80;
        }();
        if (quad_ok) {
            const double c = quad_cost * static_cast<double>(ceil_div(tiles_quad, ctx->max_quads));
            if (c < best - 1e-9) best = c, use_quad = true;
        }
        // (224 and 96 exist and are tested, but lost on every shape they were predicted to win at 128 pairs / GPU:
        // fc fwd 33.9 -> 37.0 us, proj dgrad 38.8 -> 41.1 us with 224; 160 gains ~4 % on the N = 768 shapes)
        static const int widths[] = {256, 192, 160, 128};
        for (int w : widths) {
            if (w > N + 31 && w != 96) continue;          // wider than the problem: the next width covers it
            if (amn && (w % 64)) continue;                // (A MN-major, i.e. weight gradients: 64-multiples only)
            const double c = (0.24 + 0.76 * w / 256.0) * static_cast<double>(ceil_div(tiles_of(w), sms));
            if (c < best - 1e-9) best = c, use_pair = false, use_quad = false, bn = w;
        }
    }
    const int64_t tiles = use_quad ? tiles_quad : (use_pair ? tiles_pair : tiles_of(bn));
    if (split_k <= 0)
        split_k = out_f32 ? choose_split_k(tiles, p.kb_total, use_quad ? ctx->max_quads : (use_pair ? clusters : ctx->num_sms)) : 1;
    if (split_k > p.kb_total) split_k = p.kb_total;
    p.kb_per_split = static_cast<int>(ceil_div(p.kb_total, split_k));
    split_k = static_cast<int>(ceil_div(p.kb_total, p.kb_per_split));  // no empty splits
    p.split_k = split_k;
    p.atomic_out = (out_f32 && (split_k > 1 || accumulate)) ? 1 : 0;

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (use_quad) {
        if (!amn && !bmn) return launch_quad<false, false>(ctx, A, lda, B, ldb, p, st);
        if (!amn && bmn) return launch_quad<false, true>(ctx, A, lda, B, ldb, p, st);
        return launch_quad<true, true>(ctx, A, lda, B, ldb, p, st);
    }
    if (use_pair) {
        if (!amn && !bmn) return launch_pair<false, false>(ctx, A, lda, B, ldb, p, st);
        if (!amn && bmn) return launch_pair<false, true>(ctx, A, lda, B, ldb, p, st);
        return launch_pair<true, true>(ctx, A, lda, B, ldb, p, st);
    }
#define B200_LAUNCH_BN(W)                                                                      \
    case W:                                                                                    \
        if (!amn && !bmn) return launch<W, false, false>(ctx, A, lda, B, ldb, p, st);          \
        if (!amn && bmn) return launch<W, false, true>(ctx, A, lda, B, ldb, p, st);            \
        break;
    switch (bn) {
        B200_LAUNCH_BN(96)
        B200_LAUNCH_BN(160)
        B200_LAUNCH_BN(224)
        B200_LAUNCH_BN(128)
        B200_LAUNCH_BN(192)
        B200_LAUNCH_BN(256)
        default: break;
    }
#undef B200_LAUNCH_BN
    // A MN-major (weight gradients): 64-multiple widths only
    if (bn >= 256) return launch<256, true, true>(ctx, A, lda, B, ldb, p, st);
    if (bn >= 192) return launch<192, true, true>(ctx, A, lda, B, ldb, p, st);
    return launch<128, true, true>(ctx, A, lda, B, ldb, p, st);
}
