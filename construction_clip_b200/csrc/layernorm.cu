// LayerNorm forward / backward (HBM-bound): one warp per row, whole row held in registers,
// fp32 two-pass statistics, bf16 I/O.  Replaces clip.model.LayerNorm (ln_pre, ln_1, ln_2,
// ln_post, ln_final): upstream casts to fp32, runs nn.LayerNorm(eps=1e-5) and casts back.
// The forward optionally gathers source rows (CLS / EOT pooling: ln_post(x[:,0,:]),
// ln_final(x)[arange, argmax]) and fuses the vision token assembly
// (cat(class_embedding, conv1 output) + positional_embedding) in front of ln_pre.
#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr int kLnThreads = 256;  // 8 rows per block

// 8 consecutive elements <-> 8 floats, for bf16 or fp32 storage
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&f)[8]) {
    if constexpr (sizeof(T) == 2) {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 t = unpack_bf16(w[j]);
            f[2 * j] = t.x;
            f[2 * j + 1] = t.y;
        }
    } else {
        const float4 a = *reinterpret_cast<const float4*>(p);
        const float4 b = *reinterpret_cast<const float4*>(p + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
        f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&f)[8]) {
    if constexpr (sizeof(T) == 2) {
        *reinterpret_cast<uint4*>(p) =
            make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
    } else {
        *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
}

template <int NV, typename TX, typename TY>  // NV = 8-element vectors per lane, covers d <= NV*256
__global__ void __launch_bounds__(kLnThreads)
layernorm_fwd_kernel(const TX* __restrict__ x, int64_t ldx, const int32_t* __restrict__ row_index,
                     const __nv_bfloat16* __restrict__ neg_row, const __nv_bfloat16* __restrict__ add, int add_period,
                     const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                     TY* __restrict__ y, int64_t ldy,
                     __nv_bfloat16* __restrict__ pre_out, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     int rows, int d, float eps) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = kLnThreads / 32;
    const int nvec = d >> 3;
    griddep_launch_dependents();
    griddep_wait();
    for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
        const int64_t src = row_index ? row_index[r] : r;
        float v[NV][8];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int vec = lane + 32 * i;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
            if (vec < nvec) {
                if (src >= 0) {
                    load8(x + src * ldx + vec * 8, v[i]);
                } else if (neg_row != nullptr) {
                    load8(neg_row + vec * 8, v[i]);
                }
                if (add != nullptr) {
                    const uint4 u = __ldg(reinterpret_cast<const uint4*>(add + static_cast<int64_t>(r % add_period) * d + vec * 8));
                    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = unpack_bf16(w[j]);
                        v[i][2 * j] += f.x;
                        v[i][2 * j + 1] += f.y;
                    }
                }
                if (pre_out != nullptr) {
                    uint4 o;
                    o.x = pack_bf16(v[i][0], v[i][1]);
                    o.y = pack_bf16(v[i][2], v[i][3]);
                    o.z = pack_bf16(v[i][4], v[i][5]);
                    o.w = pack_bf16(v[i][6], v[i][7]);
                    *reinterpret_cast<uint4*>(pre_out + static_cast<int64_t>(r) * ldy + vec * 8) = o;
                    // statistics are taken on the stored (bf16-rounded) value so that the
                    // backward, which re-reads pre_out, sees exactly the same x
                    const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = unpack_bf16(w[j]);
                        v[i][2 * j] = f.x;
                        v[i][2 * j + 1] = f.y;
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) sum += v[i][j];
            }
        }
        const float mean = warp_sum(sum) / d;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (lane + 32 * i < nvec) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float c = v[i][j] - mean;
                    sq += c * c;
                }
            }
        }
        const float rstd = rsqrtf(warp_sum(sq) / d + eps);
        if (lane == 0) {
            if (mean_out) mean_out[r] = mean;
            if (rstd_out) rstd_out[r] = rstd;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int vec = lane + 32 * i;
            if (vec < nvec) {
                float gf[8], bf[8], o[8];
                load8(gamma + vec * 8, gf);
                load8(beta + vec * 8, bf);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * gf[j] + bf[j];
                store8(y + static_cast<int64_t>(r) * ldy + vec * 8, o);
            }
        }
    }
}

// 4 consecutive elements <-> 4 floats (shared-memory rows: one 8 / 16-byte access per lane, conflict-free)
template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&f)[4]) {
    if constexpr (sizeof(T) == 2) {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
        f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
    } else {
        const float4 a = *reinterpret_cast<const float4*>(p);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    }
}

// dx = dres + rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*gamma,  xhat = (x-mean)*rstd
// plus the column sums dgamma = sum dy*xhat, dbeta = sum dy and (optionally) sum dx.
// One warp per row.  Every warp runs its own two-stage pipeline: lane 0 fetches the NEXT row's x, dy and
// dres with 1-D bulk copies into the warp's private shared-memory stage (completion on a per-warp
// mbarrier) while the warp makes its two passes over the CURRENT row out of shared memory, so no DRAM
// latency is exposed and the only per-thread state is the column accumulators.
// Lane l owns the 4-column units u = l + 32*m (m < NU): d <= NU * 128.
template <int NU, typename TX, bool COLSUM>
__global__ void __launch_bounds__(kLnThreads, (NU <= 6) ? 2 : 1)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int64_t lddy, const TX* __restrict__ x,
                     int64_t ldx, const int32_t* __restrict__ row_index, const __nv_bfloat16* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const __nv_bfloat16* __restrict__ dres, int64_t lddres, __nv_bfloat16* __restrict__ dx,
                     int64_t lddx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ dx_colsum, int rows, int d) {
    extern __shared__ __align__(128) uint8_t s_ln[];  // bf16 gamma[d] | per warp: 2 stages x {x | dy | dres}
    __shared__ uint64_t s_bar[kLnThreads / 32][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int nunit = d >> 2;
    const uint32_t xb = static_cast<uint32_t>(d) * sizeof(TX), yb = static_cast<uint32_t>(d) * 2u;
    const uint32_t stage_bytes = xb + 2u * yb;
    __nv_bfloat16* s_gamma = reinterpret_cast<__nv_bfloat16*>(s_ln);
    uint8_t* wbase0 = s_ln + yb;
    uint8_t* wbase = wbase0 + static_cast<size_t>(warp) * 2u * stage_bytes;
    if (lane == 0) {   // per-warp barriers: only this warp ever touches them
        mbar_init(&s_bar[warp][0], 1);
        mbar_init(&s_bar[warp][1], 1);
        fence_barrier_init();
    }
    __syncwarp();
    griddep_launch_dependents();
    griddep_wait();  // global memory from here on

    auto issue = [&](int r, int stage) {  // lane 0
        const int64_t src = row_index ? row_index[r] : r;
        uint8_t* sb = wbase + stage * stage_bytes;
        uint64_t* bar = &s_bar[warp][stage];
        mbar_arrive_expect_tx(bar, xb + yb + (dres != nullptr ? yb : 0u));
        bulk_load_1d(sb, x + src * ldx, xb, bar);
        bulk_load_1d(sb + xb, dy + static_cast<int64_t>(r) * lddy, yb, bar);
        if (dres != nullptr) bulk_load_1d(sb + xb + yb, dres + static_cast<int64_t>(r) * lddres, yb, bar);
    };

    float ag[NU][4], ab[NU][4];
    float ac[COLSUM ? NU : 1][4];
#pragma unroll
    for (int m = 0; m < NU; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ag[m][j] = ab[m][j] = 0.f;
            if constexpr (COLSUM) ac[m][j] = 0.f;
        }
    const int stride = gridDim.x * wpb;
    int r = blockIdx.x * wpb + warp;
    float mu_n = 0.f, rs_n = 0.f;
    if (r < rows) {   // the first row is on its way before anything else touches global memory
        if (lane == 0) issue(r, 0);
        mu_n = mean[r];
        rs_n = rstd[r];
    }
    for (int i = threadIdx.x; i < d; i += blockDim.x) s_gamma[i] = gamma[i];
    __syncthreads();
    const float inv_d = 1.0f / d;
    for (uint32_t k = 0; r < rows; r += stride, ++k) {
        const uint32_t st = k & 1u;
        const float mu = mu_n, rs = rs_n;
        const int rn = r + stride;
        __syncwarp();  // every lane is done with the other stage (read during the previous row)
        if (rn < rows) {
            if (lane == 0) issue(rn, st ^ 1u);
            mu_n = mean[rn];
            rs_n = rstd[rn];
        }
        mbar_wait(&s_bar[warp][st], (k >> 1) & 1u);
        const uint8_t* sb = wbase + st * stage_bytes;
        const TX* xr = reinterpret_cast<const TX*>(sb);
        const __nv_bfloat16* dyr = reinterpret_cast<const __nv_bfloat16*>(sb + xb);
        const __nv_bfloat16* drr = reinterpret_cast<const __nv_bfloat16*>(sb + xb + yb);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int m = 0; m < NU; ++m) {
            const int u = lane + 32 * m;
            if (u < nunit) {
                float xf[4], df[4], gf[4];
                load4(xr + u * 4, xf);
                load4(dyr + u * 4, df);
                load4(s_gamma + u * 4, gf);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float xh = (xf[j] - mu) * rs;
                    const float g = df[j] * gf[j];
                    ab[m][j] += df[j];
                    ag[m][j] = fmaf(df[j], xh, ag[m][j]);
                    s1 += g;
                    s2 = fmaf(g, xh, s2);
                }
            }
        }
        s1 = warp_sum(s1) * inv_d;
        s2 = warp_sum(s2) * inv_d;
        const int64_t src = row_index ? row_index[r] : r;
        __nv_bfloat16* dxr = dx + src * lddx;
#pragma unroll
        for (int m = 0; m < NU; ++m) {
            const int u = lane + 32 * m;
            if (u < nunit) {
                float xf[4], df[4], gf[4], o[4];
                load4(xr + u * 4, xf);
                load4(dyr + u * 4, df);
                load4(s_gamma + u * 4, gf);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float xh = (xf[j] - mu) * rs;
                    o[j] = rs * (df[j] * gf[j] - s1 - xh * s2);
                }
                if (dres != nullptr) {
                    float rf[4];
                    load4(drr + u * 4, rf);
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] += rf[j];
                }
                *reinterpret_cast<uint2*>(dxr + u * 4) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
                if constexpr (COLSUM) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) ac[m][j] += o[j];
                }
            }
        }
    }
    // Column sums (dgamma, dbeta, dx_colsum).  Each warp parks its partial sums in its own pipeline buffers
    // (idle now: every row it requested has been consumed; 2 stages >= 3*d floats), then the block adds the
    // warps' partials column by column and issues ONE global atomic per column.  (Shared-memory fp32
    // atomicAdd compiles to a compare-and-swap spin loop: 72 of them per thread with 8-way contention were
    // the largest fixed cost of this kernel at small row counts.)
    if (dgamma != nullptr || COLSUM) {
        float* wp = reinterpret_cast<float*>(wbase);
        __syncwarp();
#pragma unroll
        for (int m = 0; m < NU; ++m) {
            const int u = lane + 32 * m;
            if (u < nunit) {
                *reinterpret_cast<float4*>(wp + u * 4) = make_float4(ag[m][0], ag[m][1], ag[m][2], ag[m][3]);
                *reinterpret_cast<float4*>(wp + d + u * 4) = make_float4(ab[m][0], ab[m][1], ab[m][2], ab[m][3]);
                if constexpr (COLSUM)
                    *reinterpret_cast<float4*>(wp + 2 * d + u * 4) = make_float4(ac[m][0], ac[m][1], ac[m][2], ac[m][3]);
            }
        }
        __syncthreads();
        const int ncol = (COLSUM ? 3 : 2) * d;
        for (int i = threadIdx.x; i < ncol; i += blockDim.x) {
            float t = 0.f;
            for (int w = 0; w < wpb; ++w)
                t += reinterpret_cast<const float*>(wbase0 + static_cast<size_t>(w) * 2u * stage_bytes)[i];
            if (i < d) {
                if (dgamma != nullptr) atomicAdd(&dgamma[i], t);
            } else if (i < 2 * d) {
                if (dgamma != nullptr) atomicAdd(&dbeta[i - d], t);
            } else {
                atomicAdd(&dx_colsum[i - 2 * d], t);
            }
        }
    }
}

template <int NU, typename TX, bool COLSUM>
static cudaError_t launch_ln_bwd(int grid, int threads, size_t smem, cudaStream_t st, const __nv_bfloat16* dy,
                                 int64_t lddy, const TX* x, int64_t ldx, const int32_t* row_index,
                                 const __nv_bfloat16* gamma, const float* mean, const float* rstd,
                                 const __nv_bfloat16* dres, int64_t lddres, __nv_bfloat16* dx, int64_t lddx,
                                 float* dgamma, float* dbeta, float* dx_colsum, int rows, int d) {
    static size_t configured = 0;  // per instantiation; the opt-in only ever grows
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(layernorm_bwd_kernel<NU, TX, COLSUM>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    return launch_pdl(layernorm_bwd_kernel<NU, TX, COLSUM>, dim3(grid), dim3(threads), smem, st, dy, lddy, x, ldx,
                      row_index, gamma, mean, rstd, dres, lddres, dx, lddx, dgamma, dbeta, dx_colsum, rows, d);
}

}  // namespace b200

using namespace b200;

#define LN_DISPATCH(d, CALL)                       \
    do {                                           \
        if ((d) <= 256) { CALL(1); }               \
        else if ((d) <= 512) { CALL(2); }          \
        else if ((d) <= 768) { CALL(3); }          \
        else if ((d) <= 1024) { CALL(4); }         \
        else if ((d) <= 1536) { CALL(6); }         \
        else { CALL(8); }                          \
    } while (0)

extern "C" int b200clip_layernorm_fwd(b200clip_ctx* ctx, const void* x, int64_t ldx, const int32_t* row_index,
                                      const void* neg_row, const void* add, int64_t add_period, const void* gamma, const void* beta,
                                      void* y, int64_t ldy, void* pre_out, float* mean, float* rstd, int64_t rows,
                                      int64_t d, float eps, int x_dtype, int y_dtype, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(x && gamma && beta && y, "layernorm_fwd: null pointer");
    B200_CHECK_ARG(rows > 0 && rows < (1ll << 31), "layernorm_fwd: bad rows %lld", (long long)rows);
    B200_CHECK_ARG(d >= 8 && d <= 2048 && d % 8 == 0, "layernorm_fwd: d=%lld must be a multiple of 8 in [8,2048]",
                   (long long)d);
    B200_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0, "layernorm_fwd: row pitches must be multiples of 8 elements");
    B200_CHECK_ARG(add == nullptr || add_period > 0, "layernorm_fwd: add_period must be > 0");
    const int wpb = kLnThreads / 32;
    const int64_t want = ceil_div(rows, wpb);
    const int grid = static_cast<int>(want < ctx->num_sms * 8 ? want : ctx->num_sms * 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t le = cudaSuccess;
    B200_CHECK_ARG((x_dtype == B200CLIP_DT_BF16 || x_dtype == B200CLIP_DT_F32) &&
                       (y_dtype == B200CLIP_DT_BF16 || y_dtype == B200CLIP_DT_F32), "layernorm_fwd: bad dtype");
#define CALL_T(NV, TX, TY)                                                                                         \
    le = launch_pdl(layernorm_fwd_kernel<NV, TX, TY>, dim3(grid), dim3(kLnThreads), 0, st,                         \
        static_cast<const TX*>(x), ldx, row_index, static_cast<const __nv_bfloat16*>(neg_row),                     \
        static_cast<const __nv_bfloat16*>(add), static_cast<int>(add ? add_period : 1),                            \
        static_cast<const __nv_bfloat16*>(gamma), static_cast<const __nv_bfloat16*>(beta), static_cast<TY*>(y),    \
        ldy, static_cast<__nv_bfloat16*>(pre_out), mean, rstd, static_cast<int>(rows), static_cast<int>(d), eps)
#define CALL(NV)                                                                       \
    do {                                                                               \
        if (x_dtype == B200CLIP_DT_BF16 && y_dtype == B200CLIP_DT_BF16) {              \
            CALL_T(NV, __nv_bfloat16, __nv_bfloat16);                                  \
        } else if (x_dtype == B200CLIP_DT_BF16) {                                      \
            CALL_T(NV, __nv_bfloat16, float);                                          \
        } else if (y_dtype == B200CLIP_DT_BF16) {                                      \
            CALL_T(NV, float, __nv_bfloat16);                                          \
        } else {                                                                       \
            CALL_T(NV, float, float);                                                  \
        }                                                                              \
    } while (0)
    LN_DISPATCH(d, CALL);
#undef CALL
#undef CALL_T
    B200_CHECK_CUDA(le);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_layernorm_bwd(b200clip_ctx* ctx, const void* dy, int64_t lddy, const void* x, int64_t ldx,
                                      const int32_t* row_index, const void* gamma, const float* mean,
                                      const float* rstd, const void* dres, int64_t lddres, void* dx, int64_t lddx,
                                      float* dgamma, float* dbeta, float* dx_colsum, int64_t rows, int64_t d,
                                      int x_dtype, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(dy && x && gamma && mean && rstd && dx, "layernorm_bwd: null pointer");
    B200_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), "layernorm_bwd: dgamma/dbeta must come together");
    B200_CHECK_ARG(rows > 0 && rows < (1ll << 31), "layernorm_bwd: bad rows");
    B200_CHECK_ARG(d >= 8 && d <= 2048 && d % 8 == 0, "layernorm_bwd: bad d=%lld", (long long)d);
    B200_CHECK_ARG(lddy % 8 == 0 && ldx % 8 == 0 && lddx % 8 == 0 && (dres == nullptr || lddres % 8 == 0),
                   "layernorm_bwd: row pitches must be multiples of 8");
    B200_CHECK_ARG(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) |
                     reinterpret_cast<uintptr_t>(dres) | reinterpret_cast<uintptr_t>(gamma)) & 15) == 0,
                   "layernorm_bwd: pointers must be 16-byte aligned");
    B200_CHECK_ARG(x_dtype == B200CLIP_DT_BF16 || x_dtype == B200CLIP_DT_F32, "layernorm_bwd: bad x dtype");
    // shared memory: gamma, and per warp two stages of {x row, dy row, dres row} (reused for the column sums)
    const size_t xsz = x_dtype == B200CLIP_DT_F32 ? 4 : 2;
    const size_t stage = static_cast<size_t>(d) * (xsz + 4);
    const size_t fixed = static_cast<size_t>(d) * 2;
    int wpb = kLnThreads / 32;
    while (wpb > 1 && fixed + wpb * 2 * stage > 200 * 1024) wpb >>= 1;
    const size_t smem = fixed + wpb * 2 * stage;
    const int per_sm = (d <= 768 && 2 * (smem + 1024) <= 227 * 1024) ? 2 : 1;
    const int64_t want = ceil_div(rows, wpb * 4);  // >= 4 rows per warp amortise the dgamma/dbeta atomics
    const int grid = static_cast<int>(want < ctx->num_sms * per_sm ? (want > 0 ? want : 1) : ctx->num_sms * per_sm);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t le = cudaSuccess;
#define CALL_T(NU, TX, CS)                                                                                         \
    le = launch_ln_bwd<NU, TX, CS>(grid, wpb * 32, smem, st, static_cast<const __nv_bfloat16*>(dy), lddy,          \
                                   static_cast<const TX*>(x), ldx, row_index,                                      \
                                   static_cast<const __nv_bfloat16*>(gamma), mean, rstd,                           \
                                   static_cast<const __nv_bfloat16*>(dres), lddres, static_cast<__nv_bfloat16*>(dx), \
                                   lddx, dgamma, dbeta, dx_colsum, static_cast<int>(rows), static_cast<int>(d))
#define CALL(NU)                                                       \
    do {                                                               \
        if (x_dtype == B200CLIP_DT_BF16) {                             \
            if (dx_colsum) { CALL_T(NU, __nv_bfloat16, true); }        \
            else { CALL_T(NU, __nv_bfloat16, false); }                 \
        } else {                                                       \
            if (dx_colsum) { CALL_T(NU, float, true); }                \
            else { CALL_T(NU, float, false); }                         \
        }                                                              \
    } while (0)
    if (d <= 256) { CALL(2); }
    else if (d <= 512) { CALL(4); }
    else if (d <= 768) { CALL(6); }
    else if (d <= 1024) { CALL(8); }
    else if (d <= 1536) { CALL(12); }
    else { CALL(16); }
#undef CALL
#undef CALL_T
    B200_CHECK_CUDA(le);
    B200_LAUNCH_CHECK();
    return 0;
}
