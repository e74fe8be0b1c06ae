// CUDA-core kernels of the HRNet path that are HBM-bound (or tiny) rather than
// dense contractions: the median anchor (HRNet.py:200), the 2->64 first conv on
// (view, anchor) pairs with the repeat/cat of HRNet.py:201-204 fused away, and a
// layout-conversion helper used by the stage-dump test hook.
#include "internal.h"

namespace hrn {
namespace {

// ------------------------------------------------------------------ median anchor
// Lower median of the first k = min(L, 9) views per pixel: element (k-1)/2 of the
// sorted values (torch.median semantics; zero-padded views take part).
__global__ void median_anchor_kernel(const float* __restrict__ lrs, int L, int k, size_t hw, size_t total,
                                     float* __restrict__ anchor) {
    const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (idx >= total) return;
    const size_t b = idx / hw, p = idx % hw;
    const float* src = lrs + b * L * hw + p;
    float v[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) v[i] = i < k ? __ldg(src + i * hw) : __int_as_float(0x7f800000);
    // 9-element insertion sort, fully unrolled (registers only)
#pragma unroll
    for (int i = 1; i < 9; ++i) {
#pragma unroll
        for (int j = i; j > 0; --j) {
            const float lo = fminf(v[j - 1], v[j]), hi = fmaxf(v[j - 1], v[j]);
            v[j - 1] = lo;
            v[j] = hi;
        }
    }
    const int sel = (k - 1) >> 1;
    float out = v[0];
#pragma unroll
    for (int i = 1; i < 9; ++i) out = (i == sel) ? v[i] : out;
    anchor[idx] = out;
}

// ------------------------------------------------------------------ conv 2 -> 64 + PReLU
// Block = 64 x 8 output pixels of one (imageset, view), two horizontally adjacent pixels per thread so that every
// weight vector fetched from shared memory feeds two pixels.  Input channel 0 is the view, channel 1 the
// per-imageset anchor.  fp32 math, bf16 NHWC store (128 B per pixel, 256-bit stores).
constexpr int CI_TX = 32, CI_TY = 8, CI_PX = 2;
__global__ void __launch_bounds__(CI_TX* CI_TY)
conv_init_kernel(const float* __restrict__ lrs, const float* __restrict__ anchor, int L, int H, int W,
                 const float* __restrict__ w, const float* __restrict__ bias, float prelu,
                 __nv_bfloat16* __restrict__ out) {
    constexpr int TW = CI_TX * CI_PX;
    __shared__ float tile[2][CI_TY + 2][TW + 2];
    __shared__ __align__(16) float ws[18][64];   // [ci*9 + tap][co]
    __shared__ float bs[64];
    const int m = blockIdx.z;                    // image index b * L + view
    const int b = m / L;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * CI_TY;
    const int tid = threadIdx.y * CI_TX + threadIdx.x;
    const size_t hw = static_cast<size_t>(H) * W;
    const float* src0 = lrs + static_cast<size_t>(m) * hw;
    const float* src1 = anchor + static_cast<size_t>(b) * hw;
    for (int i = tid; i < 2 * (CI_TY + 2) * (TW + 2); i += CI_TX * CI_TY) {
        const int c = i / ((CI_TY + 2) * (TW + 2));
        const int r = (i / (TW + 2)) % (CI_TY + 2), q = i % (TW + 2);
        const int y = y0 + r - 1, x = x0 + q - 1;
        float v = 0.0f;
        if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg((c ? src1 : src0) + static_cast<size_t>(y) * W + x);
        tile[c][r][q] = v;
    }
    for (int i = tid; i < 18 * 64; i += CI_TX * CI_TY) {
        const int k = i / 64, co = i % 64;          // k = ci * 9 + tap ; w is (co, ci, ky, kx)
        ws[k][co] = __ldg(w + co * 18 + k);
    }
    if (tid < 64) bs[tid] = __ldg(bias + tid);
    __syncthreads();
    const int xa = x0 + CI_PX * threadIdx.x, y = y0 + threadIdx.y;
    if (xa >= W || y >= H) return;
    float in[2][18];                                // [pixel][ci*9 + ky*3 + kx]
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                const float v = tile[c][threadIdx.y + ky][CI_PX * threadIdx.x + kx];
                if (kx < 3) in[0][c * 9 + ky * 3 + kx] = v;
                if (kx > 0) in[1][c * 9 + ky * 3 + kx - 1] = v;
            }
    const float slope_m1 = prelu - 1.0f;
    __nv_bfloat16* op = out + (static_cast<size_t>(m) * hw + static_cast<size_t>(y) * W + xa) * 64;
    const bool second = xa + 1 < W;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        float acc[2][8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[0][e] = acc[1][e] = bs[g * 8 + e];
#pragma unroll
        for (int k = 0; k < 18; ++k) {
            const float4 wa = *reinterpret_cast<const float4*>(&ws[k][g * 8]);
            const float4 wb = *reinterpret_cast<const float4*>(&ws[k][g * 8 + 4]);
            const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                acc[0][e] = fmaf(in[0][k], wv[e], acc[0][e]);
                acc[1][e] = fmaf(in[1][k], wv[e], acc[1][e]);
            }
        }
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            uint4 o;
            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float p0 = acc[p][2 * e], p1 = acc[p][2 * e + 1];
                p0 = fmaf(slope_m1, fminf(p0, 0.0f), p0);      // PReLU(v) = v + (slope - 1) min(v, 0)
                p1 = fmaf(slope_m1, fminf(p1, 0.0f), p1);
                o2[e] = __floats2bfloat162_rn(p0, p1);
            }
            if (p == 0 || second) *reinterpret_cast<uint4*>(op + p * 64 + g * 8) = o;
        }
    }
}

// ------------------------------------------------------------------ bf16 NHWC -> fp32 NCHW (test hook)
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, size_t hw, int C, size_t total,
                                    float* __restrict__ out) {
    const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;   // over n * C * hw (NCHW order)
    if (idx >= total) return;
    const size_t p = idx % hw;
    const size_t c = (idx / hw) % C;
    const size_t n = idx / (hw * C);
    out[idx] = __bfloat162float(in[(n * hw + p) * C + c]);
}

}  // namespace

int median_anchor_launch(const float* lrs, int B, int L, int H, int W, float* anchor, cudaStream_t s) {
    const size_t hw = static_cast<size_t>(H) * W, total = hw * B;
    const int k = L < 9 ? L : 9;
    median_anchor_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(lrs, L, k, hw, total, anchor);
    note_launches(1);
    HRN_CUDA_OK(cudaGetLastError());
    return 0;
}

int conv_init_launch(const float* lrs, const float* anchor, int B, int L, int H, int W, const float* w,
                     const float* bias, float prelu, __nv_bfloat16* out, cudaStream_t s) {
    const long long imgs = static_cast<long long>(B) * L;
    if (imgs > 65535) {
        set_error("conv_init: B*L = %lld exceeds the grid z limit (65535); split the batch", imgs);
        return -1;
    }
    dim3 grid((W + CI_TX * CI_PX - 1) / (CI_TX * CI_PX), (H + CI_TY - 1) / CI_TY, static_cast<unsigned>(imgs));
    conv_init_kernel<<<grid, dim3(CI_TX, CI_TY), 0, s>>>(lrs, anchor, L, H, W, w, bias, prelu, out);
    note_launches(1);
    HRN_CUDA_OK(cudaGetLastError());
    return 0;
}

int nhwc_bf16_to_nchw_f32_launch(const __nv_bfloat16* in, int n, int H, int W, int C, float* out, cudaStream_t s) {
    const size_t hw = static_cast<size_t>(H) * W, total = hw * C * n;
    nhwc_to_nchw_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(in, hw, C, total, out);
    note_launches(1);
    HRN_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hrn
