// Small CUDA-core kernels of the HRNet path: the median anchor (HRNet.py:200) and a
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

int nhwc_bf16_to_nchw_f32_launch(const __nv_bfloat16* in, int n, int H, int W, int C, float* out, cudaStream_t s) {
    const size_t hw = static_cast<size_t>(H) * W, total = hw * C * n;
    nhwc_to_nchw_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(in, hw, C, total, out);
    note_launches(1);
    HRN_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hrn
