// bf16 tensor-core GEMM for sm_100a: TMA -> 128B-swizzled shared-memory ring -> tcgen05.mma
// (accumulator in TMEM, double buffered) -> tcgen05.ld -> fused epilogue.
//
// Replaces nn.Linear / F.linear inside clip.model.ResidualAttentionBlock (attn.in_proj,
// attn.out_proj, mlp.c_fc + QuickGELU, mlp.c_proj + residual), visual.conv1 (after im2col),
// `x @ visual.proj`, `x @ text_projection` and their autograd dgrad / wgrad
// (reference call sites: CLIP/train.py:161 forward, CLIP/train.py:168 backward).
//
// One persistent CTA per SM, 10 warps:
//   warp 0     TMA producer (one elected lane)
//   warp 1     TMEM allocator + MMA issuer (one elected lane)
//   warps 2-9  epilogue: warp w drains TMEM lane quadrant w%4 (32 rows of the 128-row tile) for one
//              half of the tile's columns, transposing through shared memory so that every global
//              access (aux load, C store) is a full 128-byte row segment
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty pair (MMA <-> epilogue).
#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr int BM = 128;           // tile rows  (UMMA M, cta_group::1)
constexpr int BK = 64;            // k-block: 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;        // fixed for 16-bit inputs
constexpr int kEpiWarps = 8;       // two warps per TMEM lane quadrant, each takes half of the tile's columns
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;
constexpr int kStageBytesA = BM * BK * 2;  // 16 KB
constexpr int kEpiStageBytes = 32 * 128;   // per-warp staging: 32 rows x 128 B, XOR-swizzled

template <int BN>
struct GemmCfg {
    static constexpr int kStageBytesB = BN * BK * 2;
    static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
    static constexpr int kStages = (BN == 256) ? 4 : 6;
    static constexpr int kTmemCols = 2 * BN;  // two accumulator stages
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps * kEpiStageBytes + 1024;  // + align slack
};

struct GemmParams {
    void* C;
    const __nv_bfloat16* bias;
    const __nv_bfloat16* aux;
    __nv_bfloat16* preact;
    const float* scale;
    int64_t ldc, ldaux;
    int M, N, K;
    int num_m_tiles, num_n_tiles;
    int split_k, kb_total, kb_per_split;
    int out_f32, atomic_out, epilogue;
};

// ------------------------------------------------------------------------------------------------
// Epilogue for one 32-column chunk held by one thread (one output row).
template <int EPI, bool OUT_F32, bool ATOMIC>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t (&acc)[32], int row, int col0,
                                               float scale) {
    if (row >= p.M) return;
    const int ncols = min(32, p.N - col0);  // N % 8 == 0 guaranteed by the host
#pragma unroll
    for (int v = 0; v < 4; ++v) {  // 4 vectors of 8 columns
        if (v * 8 >= ncols) break;
        const int col = col0 + v * 8;
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(acc[v * 8 + i]) * scale;
        if (p.bias != nullptr) {
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
            const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = unpack_bf16(bw[i]);
                x[2 * i] += f.x;
                x[2 * i + 1] += f.y;
            }
        }
        if constexpr (EPI == B200CLIP_EPI_QUICKGELU) {
            if (p.preact != nullptr) {
                uint4 o;
                o.x = pack_bf16(x[0], x[1]);
                o.y = pack_bf16(x[2], x[3]);
                o.z = pack_bf16(x[4], x[5]);
                o.w = pack_bf16(x[6], x[7]);
                *reinterpret_cast<uint4*>(p.preact + static_cast<int64_t>(row) * p.ldc + col) = o;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = quick_gelu(x[i]);
        } else if constexpr (EPI == B200CLIP_EPI_RESIDUAL && OUT_F32) {
            // fp32 residual stream: aux and C are fp32
            const float* ap = reinterpret_cast<const float*>(p.aux) + static_cast<int64_t>(row) * p.ldaux + col;
            const float4 a0 = *reinterpret_cast<const float4*>(ap);
            const float4 a1 = *reinterpret_cast<const float4*>(ap + 4);
            x[0] += a0.x; x[1] += a0.y; x[2] += a0.z; x[3] += a0.w;
            x[4] += a1.x; x[5] += a1.y; x[6] += a1.z; x[7] += a1.w;
        } else if constexpr (EPI == B200CLIP_EPI_RESIDUAL || EPI == B200CLIP_EPI_QUICKGELU_BWD) {
            const uint4 a = *reinterpret_cast<const uint4*>(p.aux + static_cast<int64_t>(row) * p.ldaux + col);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = unpack_bf16(aw[i]);
                if constexpr (EPI == B200CLIP_EPI_RESIDUAL) {
                    x[2 * i] += f.x;
                    x[2 * i + 1] += f.y;
                } else {
                    x[2 * i] *= quick_gelu_grad(f.x);
                    x[2 * i + 1] *= quick_gelu_grad(f.y);
                }
            }
        }
        if constexpr (OUT_F32) {
            float* dst = reinterpret_cast<float*>(p.C) + static_cast<int64_t>(row) * p.ldc + col;
            if constexpr (ATOMIC) {
#pragma unroll
                for (int i = 0; i < 8; ++i) atomicAdd(dst + i, x[i]);
            } else {
                *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(x[4], x[5], x[6], x[7]);
            }
        } else {
            uint4 o;
            o.x = pack_bf16(x[0], x[1]);
            o.y = pack_bf16(x[2], x[3]);
            o.z = pack_bf16(x[4], x[5]);
            o.w = pack_bf16(x[6], x[7]);
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<int64_t>(row) * p.ldc + col) = o;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Coalesced epilogue.  TMEM hands each thread one accumulator ROW; writing rows straight to global
// memory makes every warp-level store touch 32 different lines.  Instead each warp transposes
// through a private 4 KB shared-memory staging tile (32 rows x 128 B, 16-byte chunks XOR-swizzled
// with the row so both access patterns are bank-conflict free):
//   row layout       : thread = row, 8 chunks of 16 B                      (TMEM side, the maths)
//   coalesced layout : 8 lanes cover one 128-byte row segment, 4 rows/instr (global side)
// aux (residual / saved pre-activation) is fetched with coalesced loads through the same tile.
__device__ __forceinline__ uint32_t stage_off(int row, int chunk) {
    return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

template <int EPI, bool OUT_F32>
__device__ __forceinline__ void epilogue_staged(const GemmParams& p, uint32_t taddr, int m_base, int n_base,
                                                int ncols, float scale, uint8_t* stg, int lane) {
    constexpr int EPC = OUT_F32 ? 4 : 8;   // elements per 16-byte chunk of C (and of aux: same dtype as C,
    constexpr int CW = 8 * EPC;            //  except QUICKGELU_BWD whose aux and C are both bf16)
    constexpr bool HAS_AUX = (EPI == B200CLIP_EPI_RESIDUAL || EPI == B200CLIP_EPI_QUICKGELU_BWD);
    constexpr int ESZ = OUT_F32 ? 4 : 2;
    const int rrow = lane >> 3, rch = lane & 7;
    uint8_t* Cb = reinterpret_cast<uint8_t*>(p.C);
    const uint8_t* Ab = reinterpret_cast<const uint8_t*>(p.aux);
#pragma unroll 1
    for (int j = 0; j < ncols / CW; ++j) {
        const int col0 = n_base + j * CW;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t acc[CW];
        {
            uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&acc[0]);
            tmem_ld_32x32(taddr + j * CW, lo);
            if constexpr (CW == 64) {
                uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&acc[32]);
                tmem_ld_32x32(taddr + j * CW + 32, hi);
            }
        }
        uint4 auxv[8];
        if constexpr (HAS_AUX) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = 4 * i + rrow;
                const int grow = m_base + row, col = col0 + rch * EPC;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (grow < p.M && col < p.N)
                    v = *reinterpret_cast<const uint4*>(Ab + (static_cast<int64_t>(grow) * p.ldaux + col) * ESZ);
                *reinterpret_cast<uint4*>(stg + stage_off(row, rch)) = v;
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; ++c) auxv[c] = *reinterpret_cast<const uint4*>(stg + stage_off(lane, c));
            __syncwarp();
        }
        tmem_ld_wait();
        // ---- the maths, one 16-byte output chunk at a time (row layout)
        [[maybe_unused]] uint4 pre[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int col = col0 + c * EPC;
            float x[EPC];
#pragma unroll
            for (int e = 0; e < EPC; ++e) x[e] = __uint_as_float(acc[c * EPC + e]) * scale;
            if (p.bias != nullptr && col < p.N) {
                if constexpr (EPC == 8) {
                    const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
                    const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 f = unpack_bf16(bw[i]);
                        x[2 * i] += f.x;
                        x[2 * i + 1] += f.y;
                    }
                } else {
                    const uint2 b = __ldg(reinterpret_cast<const uint2*>(p.bias + col));
                    const float2 f0 = unpack_bf16(b.x), f1 = unpack_bf16(b.y);
                    x[0] += f0.x; x[1] += f0.y; x[2] += f1.x; x[3] += f1.y;
                }
            }
            if constexpr (EPI == B200CLIP_EPI_QUICKGELU) {
                if (p.preact != nullptr)
                    pre[c] = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]),
                                        pack_bf16(x[6], x[7]));
#pragma unroll
                for (int e = 0; e < EPC; ++e) x[e] = quick_gelu(x[e]);
            } else if constexpr (EPI == B200CLIP_EPI_RESIDUAL && OUT_F32) {
                x[0] += __uint_as_float(auxv[c].x);
                x[1] += __uint_as_float(auxv[c].y);
                x[2] += __uint_as_float(auxv[c].z);
                x[3] += __uint_as_float(auxv[c].w);
            } else if constexpr (HAS_AUX) {
                const uint32_t aw[4] = {auxv[c].x, auxv[c].y, auxv[c].z, auxv[c].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = unpack_bf16(aw[i]);
                    if constexpr (EPI == B200CLIP_EPI_RESIDUAL) {
                        x[2 * i] += f.x;
                        x[2 * i + 1] += f.y;
                    } else {
                        x[2 * i] *= quick_gelu_grad(f.x);
                        x[2 * i + 1] *= quick_gelu_grad(f.y);
                    }
                }
            }
            uint4 o;
            if constexpr (OUT_F32) {
                o = make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), __float_as_uint(x[3]));
            } else {
                o = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
            }
            *reinterpret_cast<uint4*>(stg + stage_off(lane, c)) = o;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = 4 * i + rrow;
            const int grow = m_base + row, col = col0 + rch * EPC;
            const uint4 v = *reinterpret_cast<const uint4*>(stg + stage_off(row, rch));
            if (grow < p.M && col < p.N)
                *reinterpret_cast<uint4*>(Cb + (static_cast<int64_t>(grow) * p.ldc + col) * ESZ) = v;
        }
        __syncwarp();
        if constexpr (EPI == B200CLIP_EPI_QUICKGELU) {
            if (p.preact != nullptr) {  // second output: the pre-activation (saved for the backward)
#pragma unroll
                for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(stg + stage_off(lane, c)) = pre[c];
                __syncwarp();
                uint8_t* Pb = reinterpret_cast<uint8_t*>(p.preact);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = 4 * i + rrow;
                    const int grow = m_base + row, col = col0 + rch * EPC;
                    const uint4 v = *reinterpret_cast<const uint4*>(stg + stage_off(row, rch));
                    if (grow < p.M && col < p.N)
                        *reinterpret_cast<uint4*>(Pb + (static_cast<int64_t>(grow) * p.ldc + col) * ESZ) = v;
                }
                __syncwarp();
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
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
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
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
                        for (int j = 0; j < BN / 64; ++j)
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
        const int quad = warp & 3;           // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;    // which half of the tile's columns this warp drains
        uint8_t* stg = smem + kStages * Cfg::kStageBytes + (warp - 2) * kEpiStageBytes;
        int as = 0;
        uint32_t aphase = 0;
        const float scale = (p.scale != nullptr) ? __ldg(p.scale) : 1.0f;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
            const int tile = w / p.split_k;
            const int m0 = (tile / p.num_n_tiles) * BM;
            const int n0 = (tile % p.num_n_tiles) * BN + half * (BN / 2);
            mbar_wait(&tmem_full_bar[as], aphase);
            __syncwarp();
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                   static_cast<uint32_t>(as * BN + half * (BN / 2));
            const int m_base = m0 + quad * 32;
            if (p.atomic_out) {
                // split-K / accumulate: fp32 atomics straight from the row layout
                const int row = m_base + lane;
#pragma unroll 1
                for (int c = 0; c < BN / 64; ++c) {
                    const int col0 = n0 + c * 32;
                    if (col0 >= p.N) break;  // warp-uniform
                    uint32_t acc[32];
                    tmem_ld_32x32(taddr + c * 32, acc);
                    tmem_ld_wait();
                    epilogue_chunk<B200CLIP_EPI_NONE, true, true>(p, acc, row, col0, scale);
                }
            } else if (p.out_f32) {
                if (p.epilogue == B200CLIP_EPI_RESIDUAL)
                    epilogue_staged<B200CLIP_EPI_RESIDUAL, true>(p, taddr, m_base, n0, BN / 2, scale, stg, lane);
                else
                    epilogue_staged<B200CLIP_EPI_NONE, true>(p, taddr, m_base, n0, BN / 2, scale, stg, lane);
            } else {
                switch (p.epilogue) {
                    case B200CLIP_EPI_QUICKGELU:
                        epilogue_staged<B200CLIP_EPI_QUICKGELU, false>(p, taddr, m_base, n0, BN / 2, scale, stg, lane);
                        break;
                    case B200CLIP_EPI_RESIDUAL:
                        epilogue_staged<B200CLIP_EPI_RESIDUAL, false>(p, taddr, m_base, n0, BN / 2, scale, stg, lane);
                        break;
                    case B200CLIP_EPI_QUICKGELU_BWD:
                        epilogue_staged<B200CLIP_EPI_QUICKGELU_BWD, false>(p, taddr, m_base, n0, BN / 2, scale, stg, lane);
                        break;
                    default:
                        epilogue_staged<B200CLIP_EPI_NONE, false>(p, taddr, m_base, n0, BN / 2, scale, stg, lane);
                        break;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
            if (++as == 2) {
                as = 0;
                aphase ^= 1u;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
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

int init_gemm(b200clip_ctx*) {
    int rc = 0;
    if ((rc = set_attr<256, false, false>())) return rc;
    if ((rc = set_attr<256, false, true>())) return rc;
    if ((rc = set_attr<256, true, true>())) return rc;
    if ((rc = set_attr<128, false, false>())) return rc;
    if ((rc = set_attr<128, false, true>())) return rc;
    if ((rc = set_attr<128, true, true>())) return rc;
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
    p.num_m_tiles = static_cast<int>(ceil_div(p.M, BM));
    p.num_n_tiles = static_cast<int>(ceil_div(p.N, BN));
    const int64_t work = static_cast<int64_t>(p.num_m_tiles) * p.num_n_tiles * p.split_k;
    const int grid = static_cast<int>(work < ctx->num_sms ? work : ctx->num_sms);
    gemm_bf16_kernel<BN, A_MN, B_MN><<<grid, kGemmThreads, GemmCfg<BN>::kSmemBytes, stream>>>(tmA, tmB, p);
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
                                  const void* aux, int64_t ldaux, void* preact, const float* scale, int64_t M,
                                  int64_t N, int64_t K, int epilogue, int split_k, int accumulate, void* stream) {
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
    B200_CHECK_ARG(epilogue >= 0 && epilogue <= 3, "gemm: bad epilogue %d", epilogue);
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
    B200_CHECK_ARG(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "gemm: bias not 16B aligned");
    B200_CHECK_ARG(preact == nullptr || (reinterpret_cast<uintptr_t>(preact) & 15) == 0, "gemm: preact misaligned");

    GemmParams p{};
    p.C = C;
    p.bias = static_cast<const __nv_bfloat16*>(bias);
    p.aux = static_cast<const __nv_bfloat16*>(aux);
    p.preact = static_cast<__nv_bfloat16*>(preact);
    p.scale = scale;
    p.ldc = ldc;
    p.ldaux = ldaux;
    p.M = static_cast<int>(M);
    p.N = static_cast<int>(N);
    p.K = static_cast<int>(K);
    p.kb_total = static_cast<int>(ceil_div(K, BK));
    p.out_f32 = out_f32 ? 1 : 0;
    p.epilogue = epilogue;

    // tile width: 256 when that still fills the machine, else 128 for more CTAs
    const int64_t tiles256 = ceil_div(M, BM) * ceil_div(N, 256);
    const bool use256 = (N >= 256) && (tiles256 >= ctx->num_sms || (out_f32 && epilogue == B200CLIP_EPI_NONE));
    const int64_t tiles = use256 ? tiles256 : ceil_div(M, BM) * ceil_div(N, 128);
    if (split_k <= 0) split_k = out_f32 ? choose_split_k(tiles, p.kb_total, ctx->num_sms) : 1;
    if (split_k > p.kb_total) split_k = p.kb_total;
    p.kb_per_split = static_cast<int>(ceil_div(p.kb_total, split_k));
    split_k = static_cast<int>(ceil_div(p.kb_total, p.kb_per_split));  // no empty splits
    p.split_k = split_k;
    p.atomic_out = (out_f32 && (split_k > 1 || accumulate)) ? 1 : 0;

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool amn = a_major == B200CLIP_MAJOR_MN, bmn = b_major == B200CLIP_MAJOR_MN;
    if (use256) {
        if (!amn && !bmn) return launch<256, false, false>(ctx, A, lda, B, ldb, p, st);
        if (!amn && bmn) return launch<256, false, true>(ctx, A, lda, B, ldb, p, st);
        return launch<256, true, true>(ctx, A, lda, B, ldb, p, st);
    } else {
        if (!amn && !bmn) return launch<128, false, false>(ctx, A, lda, B, ldb, p, st);
        if (!amn && bmn) return launch<128, false, true>(ctx, A, lda, B, ldb, p, st);
        return launch<128, true, true>(ctx, A, lda, B, ldb, p, st);
    }
}
