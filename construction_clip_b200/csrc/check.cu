// fp32 "check mode" of the forward path (BASELINE.json: logits within 1e-4 of the reference in an
// fp32 check mode).  Plain SIMT fp32 kernels: activations stay fp32 end to end, weights are the same
// bf16-representable values the tensor-core path reads (upcast exactly on load), accumulation is
// fp32 FMA, QuickGELU / softmax use expf.  Slow by design (no tensor cores) -- it exists to separate
// "the kernels implement the model" (checked here to 1e-4) from "bf16 storage costs ~5e-3"
// (checked on the fast path to 1e-2).  Forward only.
#include <math_constants.h>

#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr int CT = 64;  // tile
constexpr int CK = 16;

// C[m,n] = act(sum_k A[m,k] * Bw(n,k) + bias[n]) + res[m,n];  A fp32 [M,K]; B bf16 stored [N,K]
// (b_mn = 0) or [K,N] (b_mn = 1); bias bf16; res / C fp32.
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, int64_t lda, const __nv_bfloat16* __restrict__ B, int64_t ldb, int b_mn,
                const __nv_bfloat16* __restrict__ bias, const float* __restrict__ res, int64_t ldres,
                float* __restrict__ C, int64_t ldc, int M, int N, int K, int gelu) {
    __shared__ __align__(16) float As[CK][CT + 4];
    __shared__ __align__(16) float Bs[CK][CT + 4];
    const int t = threadIdx.x;
    const int ty = t >> 4, tx = t & 15;
    const int m0 = blockIdx.y * CT, n0 = blockIdx.x * CT;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int k0 = 0; k0 < K; k0 += CK) {
        // A tile: 64 rows x 16 k, one (row, 4 k) per thread
        {
            const int r = t >> 2, kk = (t & 3) * 4;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (m0 + r < M) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (k0 + kk + j < K) v[j] = A[static_cast<int64_t>(m0 + r) * lda + k0 + kk + j];
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j) As[kk + j][r] = v[j];
        }
        if (!b_mn) {  // B stored [N,K]
            const int r = t >> 2, kk = (t & 3) * 4;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (n0 + r < N) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (k0 + kk + j < K) v[j] = __bfloat162float(B[static_cast<int64_t>(n0 + r) * ldb + k0 + kk + j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) Bs[kk + j][r] = v[j];
        } else {      // B stored [K,N]
            const int kk = t >> 4, c = (t & 15) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = 0.f;
                if (k0 + kk < K && n0 + c + j < N) v = __bfloat162float(B[static_cast<int64_t>(k0 + kk) * ldb + n0 + c + j]);
                Bs[kk][c + j] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float x = acc[i][j];
            if (bias != nullptr) x += __bfloat162float(bias[n]);
            if (gelu) x = x / (1.0f + expf(-1.702f * x));
            if (res != nullptr) x += res[static_cast<int64_t>(m) * ldres + n];
            C[static_cast<int64_t>(m) * ldc + n] = x;
        }
    }
}

// softmax(q k^T / 8 + mask) v in fp32: one warp per query row, scores staged in shared memory
__global__ void __launch_bounds__(256)
attn_fwd_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int B, int S, int H, int causal) {
    extern __shared__ float s_scores[];  // [8 warps][S]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = H * 64;
    const int64_t rows = static_cast<int64_t>(B) * H * S;
    float* sc = s_scores + warp * S;
    for (int64_t w = blockIdx.x * 8 + warp; w < rows; w += static_cast<int64_t>(gridDim.x) * 8) {
        const int q = static_cast<int>(w % S);
        const int h = static_cast<int>((w / S) % H);
        const int b = static_cast<int>(w / (static_cast<int64_t>(S) * H));
        const float* qp = qkv + (static_cast<int64_t>(b) * S + q) * (3 * d) + h * 64;
        const float q0 = qp[lane], q1 = qp[lane + 32];
        const int kend = causal ? q + 1 : S;
        float mx = -CUDART_INF_F;
        for (int k = 0; k < kend; ++k) {
            const float* kp = qkv + (static_cast<int64_t>(b) * S + k) * (3 * d) + d + h * 64;
            float dot = q0 * kp[lane] + q1 * kp[lane + 32];
            dot = warp_sum(dot) * 0.125f;
            if (lane == 0) sc[k] = dot;
            mx = fmaxf(mx, dot);
        }
        __syncwarp();
        float sum = 0.f;
        for (int k = lane; k < kend; k += 32) {
            const float e = expf(sc[k] - mx);
            sc[k] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        __syncwarp();
        float o0 = 0.f, o1 = 0.f;
        for (int k = 0; k < kend; ++k) {
            const float* vp = qkv + (static_cast<int64_t>(b) * S + k) * (3 * d) + 2 * d + h * 64;
            const float pk = sc[k];
            o0 = fmaf(pk, vp[lane], o0);
            o1 = fmaf(pk, vp[lane + 32], o1);
        }
        float* op = out + (static_cast<int64_t>(b) * S + q) * d + h * 64;
        op[lane] = o0 / sum;
        op[lane + 32] = o1 / sum;
        __syncwarp();
    }
}

// im2col with fp32 output (image fp32 NCHW -> [B*g*g, ldcols] fp32, zero padded columns)
__global__ void im2col_f32_kernel(const float* __restrict__ img, float* __restrict__ cols, int64_t ldcols, int B, int R,
                                  int p, int g) {
    const int64_t n = static_cast<int64_t>(B) * g * g * ldcols;
    const int k = 3 * p * p;
    for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < n;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int col = static_cast<int>(t % ldcols);
        const int64_t prow = t / ldcols;
        float v = 0.f;
        if (col < k) {
            const int px = col % p, py = (col / p) % p, c = col / (p * p);
            const int gx = static_cast<int>(prow % g), gy = static_cast<int>((prow / g) % g);
            const int b = static_cast<int>(prow / (g * g));
            v = img[((static_cast<int64_t>(b) * 3 + c) * R + (gy * p + py)) * R + gx * p + px];
        }
        cols[t] = v;
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200clip_check_im2col_f32(b200clip_ctx* ctx, const float* image, float* cols, int64_t ldcols, int64_t B,
                                         int64_t R, int64_t patch, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(image && cols && B > 0 && R > 0 && patch > 0 && R % patch == 0 && ldcols >= 3 * patch * patch,
                   "check_im2col_f32: bad argument");
    const int g = static_cast<int>(R / patch);
    const int64_t n = B * g * g * ldcols;
    int64_t grid = ceil_div(n, 256);
    if (grid > static_cast<int64_t>(ctx->num_sms) * 16) grid = static_cast<int64_t>(ctx->num_sms) * 16;
    im2col_f32_kernel<<<static_cast<int>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        image, cols, ldcols, static_cast<int>(B), static_cast<int>(R), static_cast<int>(patch), g);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_check_gemm_f32(b200clip_ctx* ctx, const float* A, int64_t lda, const void* B_bf16, int64_t ldb,
                                       int b_major, const void* bias_bf16, const float* residual, int64_t ldres,
                                       float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int quickgelu,
                                       void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(A && B_bf16 && C && M > 0 && N > 0 && K > 0, "check_gemm_f32: bad argument");
    B200_CHECK_ARG(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "check_gemm_f32: extent too large");
    dim3 grid(static_cast<unsigned>(ceil_div(N, CT)), static_cast<unsigned>(ceil_div(M, CT)));
    gemm_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        A, lda, static_cast<const __nv_bfloat16*>(B_bf16), ldb, b_major == B200CLIP_MAJOR_MN ? 1 : 0,
        static_cast<const __nv_bfloat16*>(bias_bf16), residual, ldres, C, ldc, static_cast<int>(M), static_cast<int>(N),
        static_cast<int>(K), quickgelu ? 1 : 0);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_check_attn_fwd_f32(b200clip_ctx* ctx, const float* qkv, float* out, int64_t B, int64_t S,
                                           int64_t H, int causal, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(qkv && out && B > 0 && S > 0 && H > 0 && S <= 1024, "check_attn_fwd_f32: bad argument");
    const int64_t rows = B * H * S;
    int64_t g = ceil_div(rows, 8);
    if (g > static_cast<int64_t>(ctx->num_sms) * 8) g = static_cast<int64_t>(ctx->num_sms) * 8;
    attn_fwd_f32_kernel<<<static_cast<int>(g), 256, 8 * S * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        qkv, out, static_cast<int>(B), static_cast<int>(S), static_cast<int>(H), causal ? 1 : 0);
    B200_LAUNCH_CHECK();
    return 0;
}
