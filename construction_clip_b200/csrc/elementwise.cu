// HBM-bound helper kernels of the CLIP hot path: token embedding gather (+ positional add, EOT
// arg-max), its scatter-add backward, im2col for visual.conv1, bias-gradient column sums, vision
// token-assembly backward, L2 normalisation fwd/bwd, dtype casts and the AdamW update.
// All are plain coalesced 16-byte-vector kernels sized in multiples of the SM count.
#include "common.cuh"
#include "internal.h"

namespace b200 {

// ------------------------------------------------------------------------------------------------
// token_embedding(text) + positional_embedding       (clip.model.CLIP.encode_text, first two lines)
template <typename TOut>
__global__ void __launch_bounds__(256)
embed_tokens_fwd_kernel(const int32_t* __restrict__ ids, const __nv_bfloat16* __restrict__ table,
                        const __nv_bfloat16* __restrict__ pos, TOut* __restrict__ out, int rows, int S, int d,
                        int vocab) {
    const int lane = threadIdx.x & 31;
    const int nvec = d >> 3;
    for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8) {
        int id = ids[r];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        const int s = r % S;
        for (int vec = lane; vec < nvec; vec += 32) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(table + static_cast<int64_t>(id) * d + vec * 8));
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(pos + static_cast<int64_t>(s) * d + vec * 8));
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
            float o[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 fa = unpack_bf16(aw[j]), fb = unpack_bf16(bw[j]);
                o[2 * j] = fa.x + fb.x;
                o[2 * j + 1] = fa.y + fb.y;
            }
            TOut* dst = out + static_cast<int64_t>(r) * d + vec * 8;
            if constexpr (sizeof(TOut) == 2) {
                *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]),
                                                            pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
            } else {
                *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Packed text tower.  Under upstream's causal mask (clip.model.CLIP.build_attention_mask) nothing after a
// caption's EOT token can reach the pooled feature x[arange, text.argmax(-1)], and those positions
// receive exactly-zero gradients -- so they need not exist: caption b keeps len_b = argmax_s ids[b,s] + 1
// rows, the captions are stored back to back (cu = exclusive prefix sum of len), and every row-wise
// kernel simply sees fewer rows.  One block: a warp per caption finds its length, then a block-wide scan.
__global__ void __launch_bounds__(1024)
text_pack_plan_kernel(const int32_t* __restrict__ ids, int32_t* __restrict__ cu, int32_t* __restrict__ eot_row,
                      int B, int S, int rows_cap) {
    __shared__ int s_len[1024];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        s_carry = 0;
        cu[0] = 0;
    }
    for (int base = 0; base < B; base += 1024) {
        const int nb = min(1024, B - base);
        __syncthreads();
        for (int i = warp; i < nb; i += 32) {  // first arg-max of ids[base + i, :]
            int best = INT32_MIN, best_s = S;
            for (int t = lane; t < S; t += 32) {
                const int v = ids[static_cast<int64_t>(base + i) * S + t];
                if (v > best) {
                    best = v;
                    best_s = t;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int os = __shfl_xor_sync(0xffffffffu, best_s, o);
                if (ov > best || (ov == best && os < best_s)) {
                    best = ov;
                    best_s = os;
                }
            }
            if (lane == 0) s_len[i] = best_s + 1;
        }
        __syncthreads();
        // inclusive scan of s_len[0 .. nb) (Hillis-Steele; one element per thread)
        int v = threadIdx.x < nb ? s_len[threadIdx.x] : 0;
        for (int o = 1; o < 1024; o <<= 1) {
            __syncthreads();
            const int add = (static_cast<int>(threadIdx.x) >= o) ? s_len[threadIdx.x - o] : 0;
            __syncthreads();
            if (threadIdx.x < nb) {
                v += add;
                s_len[threadIdx.x] = v;
            }
        }
        __syncthreads();
        const int carry = s_carry;
        if (threadIdx.x < nb) {
            // never hand out rows past the caller's buffer: a too-small `rows_cap` truncates captions
            // (wrong features) instead of corrupting memory
            const int end = min(carry + v, rows_cap);
            cu[base + threadIdx.x + 1] = end;
            eot_row[base + threadIdx.x] = max(end - 1, 0);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + s_len[nb - 1];
    }
}

// token_embedding(text) + positional_embedding written straight into the packed layout; rows cu[B] .. rows_total-1
// (static-shape surplus) are zero-filled.
template <typename TOut>
__global__ void __launch_bounds__(256)
embed_tokens_packed_fwd_kernel(const int32_t* __restrict__ ids, const __nv_bfloat16* __restrict__ table,
                               const __nv_bfloat16* __restrict__ pos, const int32_t* __restrict__ cu,
                               TOut* __restrict__ out, int B, int S, int d, int vocab, int rows_total) {
    const int lane = threadIdx.x & 31;
    const int nvec = d >> 3;
    const int live = min(__ldg(cu + B), rows_total);
    const int total = B * S + (rows_total - live);   // (b, s) slots, then the surplus rows
    for (int t = blockIdx.x * 8 + (threadIdx.x >> 5); t < total; t += gridDim.x * 8) {
        if (t >= B * S) {   // zero a surplus row
            TOut* dst = out + static_cast<int64_t>(live + (t - B * S)) * d;
            for (int vec = lane; vec < nvec; vec += 32) {
                if constexpr (sizeof(TOut) == 2) {
                    *reinterpret_cast<uint4*>(dst + vec * 8) = make_uint4(0u, 0u, 0u, 0u);
                } else {
                    *reinterpret_cast<float4*>(dst + vec * 8) = make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(dst + vec * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            continue;
        }
        const int b = t / S, sp = t - b * S;
        const int row0 = __ldg(cu + b);
        if (sp >= __ldg(cu + b + 1) - row0) continue;   // after the EOT token: the position does not exist
        int id = ids[t];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        for (int vec = lane; vec < nvec; vec += 32) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(table + static_cast<int64_t>(id) * d + vec * 8));
            const uint4 bq = __ldg(reinterpret_cast<const uint4*>(pos + static_cast<int64_t>(sp) * d + vec * 8));
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bq.x, bq.y, bq.z, bq.w};
            float o[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 fa = unpack_bf16(aw[j]), fb = unpack_bf16(bw[j]);
                o[2 * j] = fa.x + fb.x;
                o[2 * j + 1] = fa.y + fb.y;
            }
            TOut* dst = out + static_cast<int64_t>(row0 + sp) * d + vec * 8;
            if constexpr (sizeof(TOut) == 2) {
                *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]),
                                                            pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
            } else {
                *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
            }
        }
    }
}

// eot_row[b] = b*S + argmax_s ids[b,s]  (first maximum, like torch.argmax)
__global__ void eot_argmax_kernel(const int32_t* __restrict__ ids, int32_t* __restrict__ eot_row, int B, int S) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    int best = INT32_MIN, best_s = S;
    for (int s = lane; s < S; s += 32) {
        const int v = ids[static_cast<int64_t>(b) * S + s];
        if (v > best) {
            best = v;
            best_s = s;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int os = __shfl_xor_sync(0xffffffffu, best_s, o);
        if (ov > best || (ov == best && os < best_s)) {
            best = ov;
            best_s = os;
        }
    }
    if (lane == 0) eot_row[b] = b * S + best_s;
}

// grid = (S, chunks): each warp walks a slice of the batch at a fixed position s, scatter-adds the
// row into dtable (skipping all-zero rows: positions after EOT receive exactly zero gradient under
// the causal mask) and keeps the positional gradient in registers until the end.
__global__ void __launch_bounds__(256)
embed_tokens_bwd_kernel(const int32_t* __restrict__ ids, const __nv_bfloat16* __restrict__ dout,
                        float* __restrict__ dtable, float* __restrict__ dpos, int B, int S, int d, int vocab,
                        const int32_t* __restrict__ cu) {  // cu != nullptr: dout is in the packed layout
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int s = blockIdx.x;
    const int nvec = d >> 3;
    const int warps_total = gridDim.y * 8;
    const int wid = blockIdx.y * 8 + warp;
    for (int vec0 = 0; vec0 < nvec; vec0 += 32) {
        const int vec = vec0 + lane;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int b = wid; b < B; b += warps_total) {
            int64_t r = static_cast<int64_t>(b) * S + s;
            if (cu != nullptr) {
                const int row0 = __ldg(cu + b);
                if (s >= __ldg(cu + b + 1) - row0) continue;   // position dropped by the packing: zero gradient
                r = row0 + s;
            }
            float f[8];
            bool nz = false;
            if (vec < nvec) {
                const uint4 u = *reinterpret_cast<const uint4*>(dout + r * d + vec * 8);
                nz = (u.x | u.y | u.z | u.w) & 0x7fff7fffu;
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 t = unpack_bf16(w[j]);
                    f[2 * j] = t.x;
                    f[2 * j + 1] = t.y;
                }
            }
            if (nz) {
                int id = ids[static_cast<int64_t>(b) * S + s];
                id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
                float* dst = dtable + static_cast<int64_t>(id) * d + vec * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    atomicAdd(dst + j, f[j]);
                    acc[j] += f[j];
                }
            }
        }
        if (vec < nvec) {
            float* dst = dpos + static_cast<int64_t>(s) * d + vec * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (acc[j] != 0.f) atomicAdd(dst + j, acc[j]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Row gather / scatter for the pooled last block (towers.py: only the CLS / EOT token of the last block's
// out_proj / ln_2 / MLP is live): dst[i,:] = src[idx[i],:]  /  dst[idx[i],:] = src[i,:], 16-byte vectors.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint4* __restrict__ src, int64_t src_ld16, const int32_t* __restrict__ idx,
                   uint4* __restrict__ dst, int n, int row16, bool scatter) {
    const int lane = threadIdx.x & 31;
    for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
        const int64_t r = idx[i];
        const uint4* sp = scatter ? src + static_cast<int64_t>(i) * row16 : src + r * src_ld16;
        uint4* dp = scatter ? dst + r * src_ld16 : dst + static_cast<int64_t>(i) * row16;
        for (int v = lane; v < row16; v += 32) dp[v] = sp[v];
    }
}

// ------------------------------------------------------------------------------------------------
// im2col for Conv2d(kernel = stride = patch): cols[(b*g + gy)*g + gx, (c*p + py)*p + px]
template <typename TIn>
__global__ void __launch_bounds__(256)
im2col_kernel(const TIn* __restrict__ img, __nv_bfloat16* __restrict__ cols, int64_t ldcols, int B, int R, int p, int g) {
    // one thread per (patch row, c, py) segment of p contiguous pixels
    const int64_t nseg = static_cast<int64_t>(B) * g * g * 3 * p;
    for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < nseg;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int py = static_cast<int>(t % p);
        const int c = static_cast<int>((t / p) % 3);
        const int64_t prow = t / (3 * p);
        const int gx = static_cast<int>(prow % g);
        const int gy = static_cast<int>((prow / g) % g);
        const int b = static_cast<int>(prow / (g * g));
        const TIn* src = img + ((static_cast<int64_t>(b) * 3 + c) * R + (gy * p + py)) * R + gx * p;
        __nv_bfloat16* dst = cols + prow * ldcols + (c * p + py) * p;
        if ((p & 7) == 0) {
            for (int px = 0; px < p; px += 8) {
                float f[8];
                if constexpr (sizeof(TIn) == 1) {   // raw pixels: ToTensor + Normalize fused (clip._transform)
                    const uint2 u = *reinterpret_cast<const uint2*>(src + px);
                    const float sc = c == 0 ? 1.0f / (255.0f * 0.26862954f) : (c == 1 ? 1.0f / (255.0f * 0.26130258f) : 1.0f / (255.0f * 0.27577711f));
                    const float sh = c == 0 ? -0.48145466f / 0.26862954f : (c == 1 ? -0.4578275f / 0.26130258f : -0.40821073f / 0.27577711f);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        f[j] = fmaf(static_cast<float>((u.x >> (8 * j)) & 0xffu), sc, sh);
                        f[4 + j] = fmaf(static_cast<float>((u.y >> (8 * j)) & 0xffu), sc, sh);
                    }
                    *reinterpret_cast<uint4*>(dst + px) = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]),
                                                                     pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                } else if constexpr (sizeof(TIn) == 4) {
                    const float4 a = *reinterpret_cast<const float4*>(src + px);
                    const float4 bq = *reinterpret_cast<const float4*>(src + px + 4);
                    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
                    f[4] = bq.x; f[5] = bq.y; f[6] = bq.z; f[7] = bq.w;
                    *reinterpret_cast<uint4*>(dst + px) = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]),
                                                                     pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                } else {
                    *reinterpret_cast<uint4*>(dst + px) = *reinterpret_cast<const uint4*>(src + px);
                }
            }
        } else {
            for (int px = 0; px < p; ++px) {
                if constexpr (sizeof(TIn) == 1) {
                    const float sc = c == 0 ? 1.0f / (255.0f * 0.26862954f) : (c == 1 ? 1.0f / (255.0f * 0.26130258f) : 1.0f / (255.0f * 0.27577711f));
                    const float sh = c == 0 ? -0.48145466f / 0.26862954f : (c == 1 ? -0.4578275f / 0.26130258f : -0.40821073f / 0.27577711f);
                    dst[px] = __float2bfloat16_rn(fmaf(static_cast<float>(src[px]), sc, sh));
                } else if constexpr (sizeof(TIn) == 4) {
                    dst[px] = __float2bfloat16_rn(static_cast<float>(src[px]));
                } else {
                    dst[px] = src[px];
                }
            }
        }
    }
}

__global__ void zero_pad_cols_kernel(__nv_bfloat16* __restrict__ cols, int64_t ldcols, int64_t rows, int k, int kpad) {
    const int w = kpad - k;
    const int64_t n = rows * w;
    for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < n;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x)
        cols[(t / w) * ldcols + k + (t % w)] = __float2bfloat16_rn(0.f);
}

// ------------------------------------------------------------------------------------------------
// out[n] += sum_m x[m,n]   (bias gradients).  block = 8 warps x 256 columns, rows chunked over grid.y
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, float* __restrict__ out, int M, int N, int rows_per_block) {
    __shared__ float s_part[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = blockIdx.x * 256 + lane * 8;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(M, r0 + rows_per_block);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    griddep_launch_dependents();
    griddep_wait();
    if (col < N) {
        for (int r = r0 + warp; r < r1; r += 8) {
            const uint4 u = *reinterpret_cast<const uint4*>(x + static_cast<int64_t>(r) * ldx + col);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16(w[j]);
                acc[2 * j] += f.x;
                acc[2 * j + 1] += f.y;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_part[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    const int c = threadIdx.x;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_part[w][c];
    if (blockIdx.x * 256 + c < N) atomicAdd(out + blockIdx.x * 256 + c, t);
}

// ------------------------------------------------------------------------------------------------
// vision assembly backward: dpre [B*n, d] -> dpatch rows (t>=1), dposcls[t,:] += sum_b dpre[b,t,:]
__global__ void __launch_bounds__(256)
vision_assemble_bwd_kernel(const __nv_bfloat16* __restrict__ dpre, __nv_bfloat16* __restrict__ dpatch,
                           float* __restrict__ dpos, float* __restrict__ dcls, int B, int n, int d) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x;
    const int nvec = d >> 3;
    const int warps_total = gridDim.y * 8;
    const int wid = blockIdx.y * 8 + warp;
    for (int vec0 = 0; vec0 < nvec; vec0 += 32) {
        const int vec = vec0 + lane;
        if (vec >= nvec) continue;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int b = wid; b < B; b += warps_total) {
            const uint4 u = *reinterpret_cast<const uint4*>(dpre + (static_cast<int64_t>(b) * n + t) * d + vec * 8);
            if (t > 0)
                *reinterpret_cast<uint4*>(dpatch + (static_cast<int64_t>(b) * (n - 1) + (t - 1)) * d + vec * 8) = u;
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16(w[j]);
                acc[2 * j] += f.x;
                acc[2 * j + 1] += f.y;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(dpos + static_cast<int64_t>(t) * d + vec * 8 + j, acc[j]);
            if (t == 0) atomicAdd(dcls + vec * 8 + j, acc[j]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// f / f.norm(dim=1, keepdim=True)       (CLIP.forward)
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ inv_norm, int B, int E) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= B) return;
    float s = 0.f;
    for (int i = lane; i < E; i += 32) {
        const float v = x[static_cast<int64_t>(r) * E + i];
        s += v * v;
    }
    const float inv = 1.0f / sqrtf(warp_sum(s));
    for (int i = lane; i < E; i += 32) y[static_cast<int64_t>(r) * E + i] = x[static_cast<int64_t>(r) * E + i] * inv;
    if (lane == 0) inv_norm[r] = inv;
}

__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ inv_norm,
                  __nv_bfloat16* __restrict__ dx, int B, int E) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= B) return;
    float dot = 0.f;
    for (int i = lane; i < E; i += 32) dot += dy[static_cast<int64_t>(r) * E + i] * y[static_cast<int64_t>(r) * E + i];
    dot = warp_sum(dot);
    const float inv = inv_norm[r];
    for (int i = lane; i < E; i += 32) {
        const int64_t k = static_cast<int64_t>(r) * E + i;
        dx[k] = __float2bfloat16_rn((dy[k] - y[k] * dot) * inv);
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void cast_f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t nv = n >> 3;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nv; i += stride) {
        const float4 a = reinterpret_cast<const float4*>(src)[2 * i];
        const float4 b = reinterpret_cast<const float4*>(src)[2 * i + 1];
        reinterpret_cast<uint4*>(dst)[i] =
            make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
    }
    for (int64_t i = (nv << 3) + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride)
        dst[i] = __float2bfloat16_rn(src[i]);
}
// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): two bf16 GEMM operands that together carry ~16 mantissa bits
__global__ void split_f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                         __nv_bfloat16* __restrict__ lo, int64_t n) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride) {
        const float x = src[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
    }
}
__global__ void cast_bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t n) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride)
        dst[i] = __bfloat162float(src[i]);
}

// ------------------------------------------------------------------------------------------------
// AdamW as the reference's optimiser computes it -- transformers.AdamW (CLIP/train.py:143, correct_bias=True):
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps) ; p -= lr wd p
// Unlike torch.optim.AdamW, eps is added to the UN-corrected sqrt(v) (an effective eps 1/sqrt(1-b2^t) times larger,
// ~32x at t = 1) and the decoupled decay is applied after the update with the plain lr.
// fp32 master weights + bf16 shadow copy.
template <typename TG>   // gradient storage: fp32 (flat accumulators) or bf16 (the wire format of the sharded reduce-scatter)
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ master, __nv_bfloat16* __restrict__ param, const TG* __restrict__ grad,
             float* __restrict__ m, float* __restrict__ v, int64_t n, float lr, float beta1, float beta2, float eps,
             float wd, float grad_scale, float bc1, float bc2, const float* __restrict__ hyper) {
    if (hyper != nullptr) {  // step-dependent scalars from device memory (CUDA-graph replay)
        lr = __ldg(hyper);
        bc1 = __ldg(hyper + 1);
        bc2 = __ldg(hyper + 2);
    }
    const float step_size = lr * sqrtf(bc2) / bc1;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride) {
        float g;
        if constexpr (sizeof(TG) == 2)
            g = __bfloat162float(grad[i]) * grad_scale;
        else
            g = grad[i] * grad_scale;
        float p = master[i];
        const float mi = beta1 * m[i] + (1.f - beta1) * g;
        const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
        m[i] = mi;
        v[i] = vi;
        p -= step_size * mi / (sqrtf(vi) + eps);
        p -= lr * wd * p;
        master[i] = p;
        if (param != nullptr) param[i] = __float2bfloat16_rn(p);
    }
}

static inline int grid_for(int64_t n, int threads, int num_sms, int per_sm = 8) {
    int64_t g = ceil_div(n, threads);
    const int64_t cap = static_cast<int64_t>(num_sms) * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

}  // namespace b200

using namespace b200;

extern "C" int b200clip_embed_tokens_fwd(b200clip_ctx* ctx, const int32_t* ids, const void* table, const void* pos,
                                         void* out, int out_dtype, int32_t* eot_row, int64_t B, int64_t S, int64_t d,
                                         int64_t vocab, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(ids && table && pos && out, "embed_tokens_fwd: null pointer");
    B200_CHECK_ARG(B > 0 && S > 0 && d > 0 && d % 8 == 0 && vocab > 0 && B * S < (1ll << 31), "embed_tokens_fwd: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rows = static_cast<int>(B * S);
    B200_CHECK_ARG(out_dtype == B200CLIP_DT_BF16 || out_dtype == B200CLIP_DT_F32, "embed_tokens_fwd: bad out_dtype");
    const int grid = grid_for(ceil_div(rows, 8) * 256, 256, ctx->num_sms);
    if (out_dtype == B200CLIP_DT_BF16)
        embed_tokens_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
            ids, static_cast<const __nv_bfloat16*>(table), static_cast<const __nv_bfloat16*>(pos),
            static_cast<__nv_bfloat16*>(out), rows, static_cast<int>(S), static_cast<int>(d), static_cast<int>(vocab));
    else
        embed_tokens_fwd_kernel<float><<<grid, 256, 0, st>>>(
            ids, static_cast<const __nv_bfloat16*>(table), static_cast<const __nv_bfloat16*>(pos),
            static_cast<float*>(out), rows, static_cast<int>(S), static_cast<int>(d), static_cast<int>(vocab));
    B200_LAUNCH_CHECK();
    if (eot_row != nullptr) {
        eot_argmax_kernel<<<static_cast<int>(ceil_div(B, 8)), 256, 0, st>>>(ids, eot_row, static_cast<int>(B),
                                                                           static_cast<int>(S));
        B200_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int b200clip_embed_tokens_bwd(b200clip_ctx* ctx, const int32_t* ids, const void* dout, float* dtable,
                                         float* dpos, int64_t B, int64_t S, int64_t d, int64_t vocab, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(ids && dout && dtable && dpos, "embed_tokens_bwd: null pointer");
    B200_CHECK_ARG(B > 0 && S > 0 && d > 0 && d % 8 == 0 && vocab > 0 && B * S < (1ll << 31), "embed_tokens_bwd: bad shape");
    int chunks = static_cast<int>(ceil_div(static_cast<int64_t>(ctx->num_sms) * 4, S));
    const int max_chunks = static_cast<int>(ceil_div(B, 8));
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    dim3 grid(static_cast<unsigned>(S), static_cast<unsigned>(chunks));
    embed_tokens_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        ids, static_cast<const __nv_bfloat16*>(dout), dtable, dpos, static_cast<int>(B), static_cast<int>(S),
        static_cast<int>(d), static_cast<int>(vocab), nullptr);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_gather_rows(b200clip_ctx* ctx, const void* src, int64_t src_ld_bytes, const int32_t* idx,
                                    void* dst, int64_t n, int64_t row_bytes, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(src && idx && dst && n > 0 && n < (1ll << 31), "gather_rows: bad argument");
    B200_CHECK_ARG(row_bytes > 0 && row_bytes % 16 == 0 && src_ld_bytes % 16 == 0 && src_ld_bytes >= row_bytes &&
                       (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                   "gather_rows: rows must be 16-byte multiples, 16-byte aligned");
    gather_rows_kernel<<<grid_for(ceil_div(n, 8) * 256, 256, ctx->num_sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(src), src_ld_bytes / 16, idx, static_cast<uint4*>(dst), static_cast<int>(n),
        static_cast<int>(row_bytes / 16), false);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_scatter_rows(b200clip_ctx* ctx, const void* src, const int32_t* idx, void* dst,
                                     int64_t dst_rows, int64_t n, int64_t row_bytes, int zero_first, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(src && idx && dst && n > 0 && n < (1ll << 31) && dst_rows > 0, "scatter_rows: bad argument");
    B200_CHECK_ARG(row_bytes > 0 && row_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                   "scatter_rows: rows must be 16-byte multiples, 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (zero_first) B200_CHECK_CUDA(cudaMemsetAsync(dst, 0, static_cast<size_t>(dst_rows * row_bytes), st));
    gather_rows_kernel<<<grid_for(ceil_div(n, 8) * 256, 256, ctx->num_sms), 256, 0, st>>>(
        static_cast<const uint4*>(src), row_bytes / 16, idx, static_cast<uint4*>(dst), static_cast<int>(n),
        static_cast<int>(row_bytes / 16), true);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_text_pack_plan(b200clip_ctx* ctx, const int32_t* ids, int32_t* cu, int32_t* eot_row, int64_t B,
                                       int64_t S, int64_t rows_cap, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(ids && cu && eot_row, "text_pack_plan: null pointer");
    B200_CHECK_ARG(B > 0 && S > 0 && B * S < (1ll << 31) && rows_cap > 0 && rows_cap < (1ll << 31),
                   "text_pack_plan: bad shape");
    text_pack_plan_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(ids, cu, eot_row, static_cast<int>(B),
                                                                            static_cast<int>(S),
                                                                            static_cast<int>(rows_cap));
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_embed_tokens_packed_fwd(b200clip_ctx* ctx, const int32_t* ids, const void* table,
                                                const void* pos, const int32_t* cu, void* out, int out_dtype, int64_t B,
                                                int64_t S, int64_t d, int64_t vocab, int64_t rows_total, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(ids && table && pos && cu && out, "embed_tokens_packed_fwd: null pointer");
    B200_CHECK_ARG(B > 0 && S > 0 && d > 0 && d % 8 == 0 && vocab > 0 && B * S < (1ll << 30) && rows_total > 0 &&
                       rows_total < (1ll << 30),
                   "embed_tokens_packed_fwd: bad shape");
    B200_CHECK_ARG(out_dtype == B200CLIP_DT_BF16 || out_dtype == B200CLIP_DT_F32, "embed_tokens_packed_fwd: bad out_dtype");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(ceil_div(B * S + rows_total, 8) * 256, 256, ctx->num_sms);
    if (out_dtype == B200CLIP_DT_BF16)
        embed_tokens_packed_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
            ids, static_cast<const __nv_bfloat16*>(table), static_cast<const __nv_bfloat16*>(pos), cu,
            static_cast<__nv_bfloat16*>(out), static_cast<int>(B), static_cast<int>(S), static_cast<int>(d),
            static_cast<int>(vocab), static_cast<int>(rows_total));
    else
        embed_tokens_packed_fwd_kernel<float><<<grid, 256, 0, st>>>(
            ids, static_cast<const __nv_bfloat16*>(table), static_cast<const __nv_bfloat16*>(pos), cu,
            static_cast<float*>(out), static_cast<int>(B), static_cast<int>(S), static_cast<int>(d),
            static_cast<int>(vocab), static_cast<int>(rows_total));
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_embed_tokens_packed_bwd(b200clip_ctx* ctx, const int32_t* ids, const void* dout,
                                                const int32_t* cu, float* dtable, float* dpos, int64_t B, int64_t S,
                                                int64_t d, int64_t vocab, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(ids && dout && cu && dtable && dpos, "embed_tokens_packed_bwd: null pointer");
    B200_CHECK_ARG(B > 0 && S > 0 && d > 0 && d % 8 == 0 && vocab > 0 && B * S < (1ll << 31),
                   "embed_tokens_packed_bwd: bad shape");
    int chunks = static_cast<int>(ceil_div(static_cast<int64_t>(ctx->num_sms) * 4, S));
    const int max_chunks = static_cast<int>(ceil_div(B, 8));
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    dim3 grid(static_cast<unsigned>(S), static_cast<unsigned>(chunks));
    embed_tokens_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        ids, static_cast<const __nv_bfloat16*>(dout), dtable, dpos, static_cast<int>(B), static_cast<int>(S),
        static_cast<int>(d), static_cast<int>(vocab), cu);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_im2col_patch(b200clip_ctx* ctx, const void* image, int in_dtype, void* cols, int64_t ldcols,
                                     int64_t B, int64_t R, int64_t patch, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(image && cols, "im2col: null pointer");
    B200_CHECK_ARG(B > 0 && R > 0 && patch > 0 && R % patch == 0, "im2col: bad shape B=%lld R=%lld p=%lld", (long long)B,
                   (long long)R, (long long)patch);
    const int64_t k = 3 * patch * patch;
    B200_CHECK_ARG(ldcols >= k && ldcols % 8 == 0, "im2col: ldcols=%lld must be >= %lld and a multiple of 8",
                   (long long)ldcols, (long long)k);
    B200_CHECK_ARG(in_dtype == B200CLIP_DT_BF16 || in_dtype == B200CLIP_DT_F32 || in_dtype == B200CLIP_DT_U8,
                   "im2col: bad in_dtype");
    const int g = static_cast<int>(R / patch);
    const int64_t nseg = B * g * g * 3 * patch;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(nseg, 256, ctx->num_sms, 16);
    if (in_dtype == B200CLIP_DT_U8)
        im2col_kernel<uint8_t><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(image), static_cast<__nv_bfloat16*>(cols),
                                                     ldcols, static_cast<int>(B), static_cast<int>(R),
                                                     static_cast<int>(patch), g);
    else if (in_dtype == B200CLIP_DT_F32)
        im2col_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(image), static_cast<__nv_bfloat16*>(cols),
                                                   ldcols, static_cast<int>(B), static_cast<int>(R),
                                                   static_cast<int>(patch), g);
    else
        im2col_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(image),
                                                           static_cast<__nv_bfloat16*>(cols), ldcols,
                                                           static_cast<int>(B), static_cast<int>(R),
                                                           static_cast<int>(patch), g);
    B200_LAUNCH_CHECK();
    if (ldcols > k) {
        const int64_t rows = B * g * g;
        zero_pad_cols_kernel<<<grid_for(rows * (ldcols - k), 256, ctx->num_sms), 256, 0, st>>>(
            static_cast<__nv_bfloat16*>(cols), ldcols, rows, static_cast<int>(k), static_cast<int>(ldcols));
        B200_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int b200clip_colsum(b200clip_ctx* ctx, const void* x, int64_t ldx, float* out, int64_t M, int64_t N,
                               void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(x && out, "colsum: null pointer");
    B200_CHECK_ARG(M > 0 && N > 0 && N % 8 == 0 && ldx % 8 == 0 && M < (1ll << 31), "colsum: bad shape");
    const int gx = static_cast<int>(ceil_div(N, 256));
    int gy = static_cast<int>(ceil_div(static_cast<int64_t>(ctx->num_sms) * 4, gx));
    const int max_gy = static_cast<int>(ceil_div(M, 64));
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    const int rows_per_block = static_cast<int>(ceil_div(M, gy));
    gy = static_cast<int>(ceil_div(M, rows_per_block));
    B200_CHECK_CUDA(launch_pdl(colsum_kernel, dim3(gx, gy), dim3(256), 0, static_cast<cudaStream_t>(stream),
                               static_cast<const __nv_bfloat16*>(x), ldx, out, static_cast<int>(M), static_cast<int>(N),
                               rows_per_block));
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_vision_assemble_bwd(b200clip_ctx* ctx, const void* dpre, void* dpatch, float* dpos,
                                            float* dcls, int64_t B, int64_t n, int64_t d, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(dpre && dpatch && dpos && dcls, "vision_assemble_bwd: null pointer");
    B200_CHECK_ARG(B > 0 && n > 1 && d > 0 && d % 8 == 0, "vision_assemble_bwd: bad shape");
    int chunks = static_cast<int>(ceil_div(static_cast<int64_t>(ctx->num_sms) * 4, n));
    const int max_chunks = static_cast<int>(ceil_div(B, 8));
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    vision_assemble_bwd_kernel<<<dim3(static_cast<unsigned>(n), chunks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(dpre), static_cast<__nv_bfloat16*>(dpatch), dpos, dcls, static_cast<int>(B),
        static_cast<int>(n), static_cast<int>(d));
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_l2norm_fwd(b200clip_ctx* ctx, const float* x, float* y, float* inv_norm, int64_t B, int64_t E,
                                   void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(x && y && inv_norm && B > 0 && E > 0, "l2norm_fwd: bad argument");
    l2norm_fwd_kernel<<<static_cast<int>(ceil_div(B, 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, y, inv_norm, static_cast<int>(B), static_cast<int>(E));
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_l2norm_bwd(b200clip_ctx* ctx, const float* dy, const float* y, const float* inv_norm,
                                   void* dx_bf16, int64_t B, int64_t E, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(dy && y && inv_norm && dx_bf16 && B > 0 && E > 0, "l2norm_bwd: bad argument");
    l2norm_bwd_kernel<<<static_cast<int>(ceil_div(B, 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dy, y, inv_norm, static_cast<__nv_bfloat16*>(dx_bf16), static_cast<int>(B), static_cast<int>(E));
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_cast_f32_to_bf16(b200clip_ctx* ctx, const float* src, void* dst, int64_t n, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(src && dst && n > 0, "cast_f32_to_bf16: bad argument");
    B200_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                   "cast_f32_to_bf16: pointers must be 16-byte aligned");
    cast_f32_to_bf16_kernel<<<grid_for(n / 8 + 1, 256, ctx->num_sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, static_cast<__nv_bfloat16*>(dst), n);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_split_f32_to_bf16(b200clip_ctx* ctx, const float* src, void* hi, void* lo, int64_t n,
                                          void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(src && hi && lo && n > 0, "split_f32_to_bf16: bad argument");
    split_f32_to_bf16_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), n);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_cast_bf16_to_f32(b200clip_ctx* ctx, const void* src, float* dst, int64_t n, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(src && dst && n > 0, "cast_bf16_to_f32: bad argument");
    cast_bf16_to_f32_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(src), dst, n);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_adamw(b200clip_ctx* ctx, float* master, void* param_bf16, const float* grad, float* m, float* v,
                              int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                              float grad_scale, int64_t step, const float* hyper_dev, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(master && grad && m && v && n > 0 && (step > 0 || hyper_dev), "adamw: bad argument");
    const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
    const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
    adamw_kernel<float><<<grid_for(n, 256, ctx->num_sms, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        master, static_cast<__nv_bfloat16*>(param_bf16), grad, m, v, n, lr, beta1, beta2, eps, weight_decay, grad_scale,
        bc1, bc2, hyper_dev);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200clip_adamw_g16(b200clip_ctx* ctx, float* master, void* param_bf16, const void* grad_bf16, float* m,
                                  float* v, int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                                  float grad_scale, int64_t step, const float* hyper_dev, void* stream) {
    B200_CHECK_CTX(ctx);
    B200_CHECK_ARG(master && grad_bf16 && m && v && n > 0 && (step > 0 || hyper_dev), "adamw_g16: bad argument");
    const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
    const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
    adamw_kernel<__nv_bfloat16><<<grid_for(n, 256, ctx->num_sms, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        master, static_cast<__nv_bfloat16*>(param_bf16), static_cast<const __nv_bfloat16*>(grad_bf16), m, v, n, lr, beta1,
        beta2, eps, weight_decay, grad_scale, bc1, bc2, hyper_dev);
    B200_LAUNCH_CHECK();
    return 0;
}
