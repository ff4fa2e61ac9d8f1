// Similarity logits and the fused symmetric InfoNCE loss.
//
//   logits_per_image = logit_scale.exp() * I @ T.t()            (clip.model.CLIP.forward)
//   loss = (CE(logits_per_image, arange) + CE(logits_per_text, arange)) / 2   (CLIP/train.py:162-166)
//   accuracy = (argmax(logits_per_image, 1) == arange).sum()     (CLIP/train.py:173)
//
// The training path never writes the Bg x Bg logits: the forward streams 64x64 fp32 similarity
// tiles through an online log-sum-exp; the backward rebuilds each tile once, turns it into the
// combined softmax-gradient tile G (bf16) for the LOCAL rows only and feeds it to the tcgen05
// GEMM (d_img = c * G_img @ txt_all, d_txt = c * G_txt @ img_all).  Features stay fp32 in the
// similarity (SIMT fp32 FMA): the tolerance on logits (1e-2 abs at scale up to 100) is tighter
// than bf16 inputs allow.
#include <math_constants.h>

#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr int TS = 64;  // similarity tile (rows x cols), 256 threads x (4x4)
constexpr int TK = 16;

// acc[4][4] += X[i0 + ty*4 + a, :] . Y[j0 + tx*4 + b, :]
__device__ __forceinline__ void sim_tile(const float* __restrict__ X, int xrows, int i0, const float* __restrict__ Y,
                                         int yrows, int j0, int E, float (&acc)[4][4], float (*Xs)[TS + 4],
                                         float (*Ys)[TS + 4]) {
    const int t = threadIdx.x;
    const int lr = t >> 2, lk = (t & 3) * 4;
    const int ty = t >> 4, tx = t & 15;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int k0 = 0; k0 < E; k0 += TK) {
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), yv = xv;
        if (i0 + lr < xrows) xv = *reinterpret_cast<const float4*>(X + static_cast<int64_t>(i0 + lr) * E + k0 + lk);
        if (j0 + lr < yrows) yv = *reinterpret_cast<const float4*>(Y + static_cast<int64_t>(j0 + lr) * E + k0 + lk);
        __syncthreads();  // previous iteration's reads are done
        Xs[lk + 0][lr] = xv.x; Xs[lk + 1][lr] = xv.y; Xs[lk + 2][lr] = xv.z; Xs[lk + 3][lr] = xv.w;
        Ys[lk + 0][lr] = yv.x; Ys[lk + 1][lr] = yv.y; Ys[lk + 2][lr] = yv.z; Ys[lk + 3][lr] = yv.w;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&Xs[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Ys[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
logits_kernel(const float* __restrict__ img, const float* __restrict__ txt, const float* __restrict__ logit_scale,
              float* __restrict__ logits, int Bi, int Bt, int E) {
    __shared__ __align__(16) float Xs[TK][TS + 4];
    __shared__ __align__(16) float Ys[TK][TS + 4];
    const int i0 = blockIdx.y * TS, j0 = blockIdx.x * TS;
    float acc[4][4];
    sim_tile(img, Bi, i0, txt, Bt, j0, E, acc, Xs, Ys);
    const float s = expf(__ldg(logit_scale));
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + ty * 4 + a;
        if (i >= Bi) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = j0 + tx * 4 + b;
            if (j < Bt) logits[static_cast<int64_t>(i) * Bt + j] = s * acc[a][b];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward, pass 1: per (row tile, column split, direction) online max / sum-exp partials
struct LossFwdWs {
    float* pm;    // [2][nsplit][Bl] running max
    float* pl;    // [2][nsplit][Bl] sum exp(s - max)
    float* pv;    // [nsplit][Bl]    best logit (direction 0)
    int32_t* pi;  // [nsplit][Bl]    its column
};

__global__ void __launch_bounds__(256)
loss_fwd_partial_kernel(const float* __restrict__ img_all, const float* __restrict__ txt_all,
                        const float* __restrict__ logit_scale, int row0, int Bl, int Bg, int E, int nsplit,
                        int tiles_per_split, LossFwdWs ws) {
    __shared__ __align__(16) float Xs[TK][TS + 4];
    __shared__ __align__(16) float Ys[TK][TS + 4];
    const int dir = blockIdx.z, split = blockIdx.y;
    const float* X = (dir == 0 ? img_all : txt_all) + static_cast<int64_t>(row0) * E;
    const float* Y = (dir == 0 ? txt_all : img_all);
    const int i0 = blockIdx.x * TS;
    const float s = expf(__ldg(logit_scale));
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float m[4], l[4], bv[4];
    int bi[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        m[a] = -CUDART_INF_F;
        l[a] = 0.f;
        bv[a] = -CUDART_INF_F;
        bi[a] = 0x7fffffff;
    }
    const int ntiles = static_cast<int>(ceil_div(Bg, TS));
    const int t0 = split * tiles_per_split, t1 = min(ntiles, t0 + tiles_per_split);
    for (int jt = t0; jt < t1; ++jt) {
        const int j0 = jt * TS;
        float acc[4][4];
        sim_tile(X, Bl, i0, Y, Bg, j0, E, acc, Xs, Ys);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float tmax = -CUDART_INF_F;
            int targ = 0x7fffffff;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int j = j0 + tx * 4 + b;
                acc[a][b] = (j < Bg) ? s * acc[a][b] : -CUDART_INF_F;
                if (acc[a][b] > tmax) {
                    tmax = acc[a][b];
                    targ = j;
                }
            }
            // reduce over the 16 lanes that share this row (xor 8,4,2,1 stays inside the half-warp)
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, tmax, o);
                const int oi = __shfl_xor_sync(0xffffffffu, targ, o);
                if (ov > tmax || (ov == tmax && oi < targ)) {
                    tmax = ov;
                    targ = oi;
                }
            }
            if (tmax > bv[a] || (tmax == bv[a] && targ < bi[a])) {
                bv[a] = tmax;
                bi[a] = targ;
            }
            const float mnew = fmaxf(m[a], tmax);
            float sum = 0.f;
#pragma unroll
            for (int b = 0; b < 4; ++b) sum += __expf(acc[a][b] - mnew);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            l[a] = l[a] * __expf(m[a] - mnew) + sum;
            m[a] = mnew;
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = i0 + ty * 4 + a;
            if (i < Bl) {
                const int64_t k = (static_cast<int64_t>(dir) * nsplit + split) * Bl + i;
                ws.pm[k] = m[a];
                ws.pl[k] = l[a];
                if (dir == 0) {
                    ws.pv[static_cast<int64_t>(split) * Bl + i] = bv[a];
                    ws.pi[static_cast<int64_t>(split) * Bl + i] = bi[a];
                }
            }
        }
    }
}

// forward, pass 2: merge the splits per local row; add to loss sums / correct count
__global__ void __launch_bounds__(256)
loss_fwd_finalize_kernel(const float* __restrict__ img_all, const float* __restrict__ txt_all,
                         const float* __restrict__ logit_scale, int row0, int Bl, int E, int nsplit, LossFwdWs ws,
                         float* __restrict__ lse_i, float* __restrict__ lse_t, float* __restrict__ loss_sum,
                         int32_t* __restrict__ correct) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    float li = 0.f, lt = 0.f;
    int ok = 0;
    if (i < Bl) {
        const float s = expf(__ldg(logit_scale));
        const float* a = img_all + static_cast<int64_t>(row0 + i) * E;
        const float* b = txt_all + static_cast<int64_t>(row0 + i) * E;
        float dot = 0.f;
        for (int e = lane; e < E; e += 32) dot = fmaf(a[e], b[e], dot);
        const float diag = s * warp_sum(dot);
        float lse[2];
        for (int dir = 0; dir < 2; ++dir) {
            float m = -CUDART_INF_F;
            for (int sp = 0; sp < nsplit; ++sp) m = fmaxf(m, ws.pm[(static_cast<int64_t>(dir) * nsplit + sp) * Bl + i]);
            float l = 0.f;
            for (int sp = 0; sp < nsplit; ++sp) {
                const int64_t k = (static_cast<int64_t>(dir) * nsplit + sp) * Bl + i;
                l += ws.pl[k] * __expf(ws.pm[k] - m);
            }
            lse[dir] = m + logf(l);
        }
        float bv = -CUDART_INF_F;
        int bi = 0x7fffffff;
        for (int sp = 0; sp < nsplit; ++sp) {
            const float v = ws.pv[static_cast<int64_t>(sp) * Bl + i];
            const int c = ws.pi[static_cast<int64_t>(sp) * Bl + i];
            if (v > bv || (v == bv && c < bi)) {
                bv = v;
                bi = c;
            }
        }
        if (lane == 0) {
            lse_i[i] = lse[0];
            lse_t[i] = lse[1];
            li = lse[0] - diag;
            lt = lse[1] - diag;
            ok = (bi == row0 + i) ? 1 : 0;
        }
    }
    // block reduce (8 warps; only lane 0 of each holds a value)
    __shared__ float s_li[8], s_lt[8];
    __shared__ int s_ok[8];
    if (lane == 0) {
        s_li[threadIdx.x >> 5] = li;
        s_lt[threadIdx.x >> 5] = lt;
        s_ok[threadIdx.x >> 5] = ok;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        int c = 0;
        for (int w = 0; w < 8; ++w) {
            a += s_li[w];
            b += s_lt[w];
            c += s_ok[w];
        }
        atomicAdd(&loss_sum[0], a);
        atomicAdd(&loss_sum[1], b);
        if (correct != nullptr && c) atomicAdd(correct, c);
    }
}

// ------------------------------------------------------------------------------------------------
// backward: G tiles (bf16) for the local rows.
//   dir 0: rows = local img i, cols = all txt j : G = exp(s d - lse_i[i]) + exp(s d - lse_t[j]) - 2 delta
//   dir 1: rows = local txt i, cols = all img j : G = exp(s d - lse_t[i]) + exp(s d - lse_i[j]) - 2 delta
// plus (dir 0) the logit_scale gradient  sum G * d * coef,  coef = g * s / (2 Bg)
__global__ void __launch_bounds__(256)
loss_bwd_g_kernel(const float* __restrict__ img_all, const float* __restrict__ txt_all,
                  const float* __restrict__ logit_scale, const float* __restrict__ lse_i_all,
                  const float* __restrict__ lse_t_all, const float* __restrict__ grad_out, int row0, int Bl, int Bg,
                  int E, __nv_bfloat16* __restrict__ G, int ldg, float* __restrict__ coef_out,
                  float* __restrict__ d_logit_scale) {
    __shared__ __align__(16) float Xs[TK][TS + 4];
    __shared__ __align__(16) float Ys[TK][TS + 4];
    __shared__ float s_red[8];
    const int dir = blockIdx.z;
    const float* X = (dir == 0 ? img_all : txt_all) + static_cast<int64_t>(row0) * E;
    const float* Y = (dir == 0 ? txt_all : img_all);
    const float* lse_row = (dir == 0 ? lse_i_all : lse_t_all) + row0;
    const float* lse_col = (dir == 0 ? lse_t_all : lse_i_all);
    const int i0 = blockIdx.y * TS, j0 = blockIdx.x * TS;
    float acc[4][4];
    sim_tile(X, Bl, i0, Y, Bg, j0, E, acc, Xs, Ys);
    const float s = expf(__ldg(logit_scale));
    const float g = grad_out ? __ldg(grad_out) : 1.0f;
    const float coef = g * s / (2.0f * Bg);
    if (blockIdx.x == 0 && blockIdx.y == 0 && dir == 0 && threadIdx.x == 0) *coef_out = coef;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    __nv_bfloat16* Gd = G + static_cast<int64_t>(dir) * Bl * ldg;  // row pitch ldg = Bg rounded up to 8 (TMA pitch rule)
    float dsum = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + ty * 4 + a;
        if (i >= Bl) continue;
        const float lr = lse_row[i];
        float gv[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = j0 + tx * 4 + b;
            gv[b] = 0.f;
            if (j < Bg) {
                const float z = s * acc[a][b];
                gv[b] = __expf(z - lr) + __expf(z - lse_col[j]) - ((row0 + i == j) ? 2.0f : 0.0f);
                dsum += gv[b] * acc[a][b];
            }
        }
        const int j = j0 + tx * 4;
        if (j + 3 < ldg) {  // columns Bg .. ldg-1 are written as zeros (gv = 0 there)
            uint2 o;
            o.x = pack_bf16(gv[0], gv[1]);
            o.y = pack_bf16(gv[2], gv[3]);
            *reinterpret_cast<uint2*>(Gd + static_cast<int64_t>(i) * ldg + j) = o;
        }
    }
    if (dir == 0 && d_logit_scale != nullptr) {
        dsum = warp_sum(dsum);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = dsum;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int w = 0; w < 8; ++w) t += s_red[w];
            atomicAdd(d_logit_scale, t * coef);
        }
    }
}

static inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

static int loss_nsplit(int64_t Bl, int64_t Bg, int num_sms) {
    const int64_t row_tiles = ceil_div(Bl, TS);
    const int64_t col_tiles = ceil_div(Bg, TS);
    int64_t ns = ceil_div(static_cast<int64_t>(num_sms) * 2, row_tiles * 2);
    if (ns > col_tiles) ns = col_tiles;
    if (ns < 1) ns = 1;
    const int64_t tps = ceil_div(col_tiles, ns);
    return static_cast<int>(ceil_div(col_tiles, tps));
}

}  // namespace b200

using namespace b200;

extern "C" int64_t b200clip_clip_loss_workspace_bytes(b200clip_ctx* ctx, int64_t Bl, int64_t Bg, int64_t E) {
    if (ctx == nullptr || Bl <= 0 || Bg <= 0 || E <= 0) return -1;
    const int ns = loss_nsplit(Bl, Bg, ctx->num_sms);
    const size_t fwd = align256(sizeof(float) * 2 * ns * Bl) * 2 + align256(sizeof(float) * ns * Bl) * 2;
    const size_t ldg = static_cast<size_t>((Bg + 7) / 8 * 8);
    const size_t bwd = align256(2ull * Bl * ldg * 2) + 2 * align256(static_cast<size_t>(Bg) * E * 2) + 256;
    return static_cast<int64_t>(fwd > bwd ? fwd : bwd);
}

extern "C" int b200clip_logits(b200clip_ctx* ctx, const float* img, const float* txt, const float* logit_scale,
                               float* logits, int64_t Bi, int64_t Bt, int64_t E, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(img && txt && logit_scale && logits, "logits: null pointer");
    B200_CHECK_ARG(Bi > 0 && Bt > 0 && E > 0 && E % TK == 0, "logits: bad shape (E must be a multiple of 16)");
    dim3 grid(static_cast<unsigned>(ceil_div(Bt, TS)), static_cast<unsigned>(ceil_div(Bi, TS)));
    logits_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(img, txt, logit_scale, logits,
                                                                       static_cast<int>(Bi), static_cast<int>(Bt),
                                                                       static_cast<int>(E));
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_clip_loss_fwd(b200clip_ctx* ctx, const float* img_all, const float* txt_all,
                                      const float* logit_scale, int64_t row0, int64_t Bl, int64_t Bg, int64_t E,
                                      float* lse_i, float* lse_t, float* loss_sum, int32_t* correct, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(img_all && txt_all && logit_scale && lse_i && lse_t && loss_sum && workspace, "clip_loss_fwd: null pointer");
    B200_CHECK_ARG(Bl > 0 && Bg > 0 && row0 >= 0 && row0 + Bl <= Bg && E > 0 && E % TK == 0, "clip_loss_fwd: bad shape");
    B200_CHECK_ARG(workspace_bytes >= b200clip_clip_loss_workspace_bytes(ctx, Bl, Bg, E), "clip_loss_fwd: workspace too small");
    const int ns = loss_nsplit(Bl, Bg, ctx->num_sms);
    const int col_tiles = static_cast<int>(ceil_div(Bg, TS));
    const int tps = static_cast<int>(ceil_div(col_tiles, ns));
    uint8_t* w = static_cast<uint8_t*>(workspace);
    LossFwdWs ws;
    ws.pm = reinterpret_cast<float*>(w);
    w += align256(sizeof(float) * 2 * ns * Bl);
    ws.pl = reinterpret_cast<float*>(w);
    w += align256(sizeof(float) * 2 * ns * Bl);
    ws.pv = reinterpret_cast<float*>(w);
    w += align256(sizeof(float) * ns * Bl);
    ws.pi = reinterpret_cast<int32_t*>(w);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid(static_cast<unsigned>(ceil_div(Bl, TS)), ns, 2);
    loss_fwd_partial_kernel<<<grid, 256, 0, st>>>(img_all, txt_all, logit_scale, static_cast<int>(row0),
                                                  static_cast<int>(Bl), static_cast<int>(Bg), static_cast<int>(E), ns, tps, ws);
    B200_LAUNCH_CHECK();
    loss_fwd_finalize_kernel<<<static_cast<int>(ceil_div(Bl, 8)), 256, 0, st>>>(
        img_all, txt_all, logit_scale, static_cast<int>(row0), static_cast<int>(Bl), static_cast<int>(E), ns, ws, lse_i,
        lse_t, loss_sum, correct);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_clip_loss_bwd(b200clip_ctx* ctx, const float* img_all, const float* txt_all,
                                      const float* logit_scale, const float* lse_i_all, const float* lse_t_all,
                                      const float* grad_out, int64_t row0, int64_t Bl, int64_t Bg, int64_t E,
                                      float* d_img, float* d_txt, float* d_logit_scale, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(img_all && txt_all && logit_scale && lse_i_all && lse_t_all && d_img && d_txt && workspace,
                   "clip_loss_bwd: null pointer");
    B200_CHECK_ARG(Bl > 0 && Bg > 0 && row0 >= 0 && row0 + Bl <= Bg && E > 0 && E % TK == 0, "clip_loss_bwd: bad shape");
    B200_CHECK_ARG(E % 8 == 0, "clip_loss_bwd: E must be a multiple of 8 (tensor-core GEMM operand)");
    const int64_t ldg = (Bg + 7) / 8 * 8;  // any global batch (the reference trains with 9): G's pitch is padded
    B200_CHECK_ARG(workspace_bytes >= b200clip_clip_loss_workspace_bytes(ctx, Bl, Bg, E), "clip_loss_bwd: workspace too small");
    uint8_t* w = static_cast<uint8_t*>(workspace);
    __nv_bfloat16* G = reinterpret_cast<__nv_bfloat16*>(w);
    w += align256(2ull * Bl * ldg * 2);
    __nv_bfloat16* img_bf = reinterpret_cast<__nv_bfloat16*>(w);
    w += align256(static_cast<size_t>(Bg) * E * 2);
    __nv_bfloat16* txt_bf = reinterpret_cast<__nv_bfloat16*>(w);
    w += align256(static_cast<size_t>(Bg) * E * 2);
    float* coef = reinterpret_cast<float*>(w);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid(static_cast<unsigned>(ceil_div(ldg, TS)), static_cast<unsigned>(ceil_div(Bl, TS)), 2);
    loss_bwd_g_kernel<<<grid, 256, 0, st>>>(img_all, txt_all, logit_scale, lse_i_all, lse_t_all, grad_out,
                                            static_cast<int>(row0), static_cast<int>(Bl), static_cast<int>(Bg),
                                            static_cast<int>(E), G, static_cast<int>(ldg), coef, d_logit_scale);
    B200_LAUNCH_CHECK();
    int rc;
    if ((rc = b200clip_cast_f32_to_bf16(ctx, img_all, img_bf, Bg * E, stream))) return rc;
    if ((rc = b200clip_cast_f32_to_bf16(ctx, txt_all, txt_bf, Bg * E, stream))) return rc;
    // d_img[Bl,E] = coef * G0[Bl,Bg] @ txt_all[Bg,E] ; d_txt = coef * G1 @ img_all   (B operand MN-major)
    B200_CHECK_CUDA(cudaMemsetAsync(d_img, 0, sizeof(float) * Bl * E, st));
    B200_CHECK_CUDA(cudaMemsetAsync(d_txt, 0, sizeof(float) * Bl * E, st));
    if ((rc = b200clip_gemm_bf16(ctx, G, ldg, B200CLIP_MAJOR_K, txt_bf, E, B200CLIP_MAJOR_MN, d_img, E, B200CLIP_DT_F32,
                                 nullptr, nullptr, 0, nullptr, coef, nullptr, Bl, E, Bg, B200CLIP_EPI_NONE, 0, 1, stream)))
        return rc;
    if ((rc = b200clip_gemm_bf16(ctx, G + Bl * ldg, ldg, B200CLIP_MAJOR_K, img_bf, E, B200CLIP_MAJOR_MN, d_txt, E,
                                 B200CLIP_DT_F32, nullptr, nullptr, 0, nullptr, coef, nullptr, Bl, E, Bg, B200CLIP_EPI_NONE, 0,
                                 1, stream)))
        return rc;
    return 0;
}
