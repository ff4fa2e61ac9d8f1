// GPU side of the image preprocessing the reference runs in its DataLoader workers (CLIP/train.py:56,138:
// `self.preprocess(Image.open(...))` = upstream clip._transform): Resize(n_px, BICUBIC) + CenterCrop(n_px) on the decoded
// RGB pixels, bit-exact with Pillow's 8-bit resampler (src/libImaging/Resample.c: fixed-point coefficients with 22
// fractional bits, a horizontal pass and a vertical pass, each accumulating in int32 from 1 << 21, shifting right by
// 22 and clipping to 0..255).  ToTensor + Normalize stay fused into the patch-embedding im2col (elementwise.cu), so a
// decoded image goes host -> device once, as uint8, at its original size, and comes out as the uint8 [3, R, R] tensor
// `model(image_uint8, text)` / `ClipTrainer.step_from_host` take.
//
// The coefficient tables are built on the host (construction_clip_b200/data.py, cached per input size) and only for the
// R output columns / rows that survive the centre crop; the horizontal pass only touches the input rows the vertical
// taps of those R output rows read.  A pass whose output size equals its input size arrives as single-tap identity
// tables (coefficient 1 << 22), which reproduces Pillow's "skip this pass" exactly.
// Byte / integer work, a few MB per image: one thread per output pixel (3 channels), neighbouring threads read
// overlapping source windows (horizontal pass) or neighbouring columns of the same rows (vertical pass).
#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr int kPrecisionBits = 22;

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= kPrecisionBits;   // arithmetic shift, as Pillow's clip8 lookup is indexed
    return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[r][x][c] = clip8(2^21 + sum_k src[row0 + r][xmin(x) + k][c] * coef[x][k])
__global__ void __launch_bounds__(256)
resize_h_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, const int32_t* __restrict__ bounds,
                const int32_t* __restrict__ coef, int ksize, int row0, int rows, int R, uint8_t* __restrict__ tmp) {
    const int64_t n = static_cast<int64_t>(rows) * R;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % R), r = static_cast<int>(i / R);
        const int xmin = __ldg(bounds + 2 * x), cnt = __ldg(bounds + 2 * x + 1);
        const uint8_t* p = src + static_cast<int64_t>(row0 + r) * src_pitch + static_cast<int64_t>(xmin) * 3;
        const int32_t* k = coef + static_cast<int64_t>(x) * ksize;
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
        for (int j = 0; j < cnt; ++j) {
            const int c = __ldg(k + j);
            s0 += static_cast<int>(p[3 * j]) * c;
            s1 += static_cast<int>(p[3 * j + 1]) * c;
            s2 += static_cast<int>(p[3 * j + 2]) * c;
        }
        uint8_t* o = tmp + i * 3;
        o[0] = clip8(s0);
        o[1] = clip8(s1);
        o[2] = clip8(s2);
    }
}

// dst[c][y][x] = clip8(2^21 + sum_k tmp[ymin(y) + k - row0][x][c] * coef[y][k])
__global__ void __launch_bounds__(256)
resize_v_kernel(const uint8_t* __restrict__ tmp, const int32_t* __restrict__ bounds, const int32_t* __restrict__ coef,
                int ksize, int row0, int R, uint8_t* __restrict__ dst) {
    const int64_t n = static_cast<int64_t>(R) * R;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % R), y = static_cast<int>(i / R);
        const int ymin = __ldg(bounds + 2 * y), cnt = __ldg(bounds + 2 * y + 1);
        const uint8_t* p = tmp + (static_cast<int64_t>(ymin - row0) * R + x) * 3;
        const int32_t* k = coef + static_cast<int64_t>(y) * ksize;
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
        for (int j = 0; j < cnt; ++j) {
            const int c = __ldg(k + j);
            const uint8_t* q = p + static_cast<int64_t>(j) * R * 3;
            s0 += static_cast<int>(q[0]) * c;
            s1 += static_cast<int>(q[1]) * c;
            s2 += static_cast<int>(q[2]) * c;
        }
        dst[i] = clip8(s0);
        dst[n + i] = clip8(s1);
        dst[2 * n + i] = clip8(s2);
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200clip_resize_crop_u8(b200clip_ctx* ctx, const void* src, int64_t H, int64_t W, int64_t src_pitch,
                                       const int32_t* xbounds, const int32_t* xcoef, int64_t xk, const int32_t* ybounds,
                                       const int32_t* ycoef, int64_t yk, int64_t row0, int64_t rows, void* tmp, void* dst,
                                       int64_t R, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(src && xbounds && xcoef && ybounds && ycoef && tmp && dst, "resize_crop: null argument");
    B200_CHECK_ARG(H > 0 && W > 0 && R > 0 && R <= 4096 && H < (1 << 20) && W < (1 << 20), "resize_crop: bad extents");
    B200_CHECK_ARG(src_pitch >= 3 * W, "resize_crop: pitch %lld < 3 x width %lld", (long long)src_pitch, (long long)W);
    B200_CHECK_ARG(xk > 0 && yk > 0 && xk < 4096 && yk < 4096, "resize_crop: bad tap counts");
    B200_CHECK_ARG(row0 >= 0 && rows > 0 && row0 + rows <= H, "resize_crop: rows [%lld, %lld) outside the image",
                   (long long)row0, (long long)(row0 + rows));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t n1 = rows * R, n2 = R * R;
    const int g1 = static_cast<int>(ceil_div(n1, 256) < 8 * ctx->num_sms ? ceil_div(n1, 256) : 8 * ctx->num_sms);
    const int g2 = static_cast<int>(ceil_div(n2, 256) < 8 * ctx->num_sms ? ceil_div(n2, 256) : 8 * ctx->num_sms);
    resize_h_kernel<<<g1, 256, 0, st>>>(static_cast<const uint8_t*>(src), src_pitch, xbounds, xcoef, static_cast<int>(xk),
                                        static_cast<int>(row0), static_cast<int>(rows), static_cast<int>(R),
                                        static_cast<uint8_t*>(tmp));
    B200_LAUNCH_CHECK();
    resize_v_kernel<<<g2, 256, 0, st>>>(static_cast<const uint8_t*>(tmp), ybounds, ycoef, static_cast<int>(yk),
                                        static_cast<int>(row0), static_cast<int>(R), static_cast<uint8_t*>(dst));
    B200_LAUNCH_CHECK();
    return B200CLIP_OK;
}
