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
// Image n of the dump is image (n / group) * stride + n % group of the source (group == stride: contiguous).
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, size_t hw, int C, size_t total, int group,
                                    int stride, float* __restrict__ out) {
    const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;   // over n * C * hw (NCHW order)
    if (idx >= total) return;
    const size_t p = idx % hw;
    const size_t c = (idx / hw) % C;
    const size_t n = idx / (hw * C);
    const size_t src = (n / group) * stride + n % group;
    out[idx] = __bfloat162float(in[(src * hw + p) * C + c]);
}

// ------------------------------------------------------------------ uint16 pixels -> float32 in [0, 1]
// DataLoader.py:195-198: skimage.img_as_float(uint16).astype(float32) = RN32(x / 65535).  A true fp32 division gives the
// same bits for all 65536 inputs; multiplying by an fp32 reciprocal does not (512 values differ).
__global__ void u16_to_unit_float_kernel(const uint16_t* __restrict__ in, size_t n, float* __restrict__ out) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        out[i] = __fdiv_rn(static_cast<float>(in[i]), 65535.0f);
}

// predict.py:176 on the way out: skimage.img_as_uint(sr) for a float32 image (scikit-image 0.24, util/dtype.py _convert:
// float32 arithmetic because the output has 2 bytes) = clip(rint(x * 65535), 0, 65535) with round-half-to-even; values
// outside [-1, 1] make skimage raise, here they set *bad (the host wrapper raises).  NaN -> 0 without a flag.
__global__ void unit_float_to_u16_kernel(const float* __restrict__ in, size_t n, uint16_t* __restrict__ out, int* __restrict__ bad) {
    bool any_bad = false;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float x = in[i];
        any_bad |= (x < -1.0f) || (x > 1.0f);
        const float r = rintf(__fmul_rn(x, 65535.0f));
        out[i] = static_cast<uint16_t>(fminf(fmaxf(r, 0.0f), 65535.0f));      // fmaxf(NaN, 0) = 0
    }
    if (__any_sync(0xffffffffu, any_bad) && (threadIdx.x & 31) == 0 && bad != nullptr) atomicOr(bad, 1);
}

// ------------------------------------------------------------------ ragged collate (utils.py:63-113) on the device
// `packed` holds only the REAL views, imageset after imageset (views [offsets[b], offsets[b + 1]) belong to imageset b);
// the kernel scatters them into the padded (B, min_L, H, W) float32 batch the model takes: the first
// min(count_b, min_L) views are copied (uint16 views are scaled like DataLoader.py:195-198), the remaining planes are
// zero-filled and their alpha is 0 (utils.py:89-95).  No padded byte ever crosses PCIe.  One block walks a slice of
// one (imageset, view) plane; 128-bit stores when the plane size allows.
template <typename T>
__global__ void __launch_bounds__(256)
collate_kernel(const T* __restrict__ packed, const int* __restrict__ offsets, int min_L, size_t hw, float* __restrict__ lrs,
               float* __restrict__ alphas) {
    const int v = blockIdx.y, b = blockIdx.z;
    const int first = offsets[b], count = offsets[b + 1] - first;
    const bool real = v < count;                                     // v < min_L by the grid
    if (blockIdx.x == 0 && threadIdx.x == 0) alphas[static_cast<size_t>(b) * min_L + v] = real ? 1.0f : 0.0f;
    float* dst = lrs + (static_cast<size_t>(b) * min_L + v) * hw;
    const T* src = packed + static_cast<size_t>(first + (real ? v : 0)) * hw;
    const size_t step = static_cast<size_t>(gridDim.x) * blockDim.x;
    if ((hw & 3) == 0) {
        for (size_t q = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; q < hw / 4; q += step) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (real) {
                if (sizeof(T) == 4) {
                    o = __ldg(reinterpret_cast<const float4*>(src) + q);
                } else {
                    const ushort4 u = __ldg(reinterpret_cast<const ushort4*>(src) + q);
                    o = make_float4(__fdiv_rn(static_cast<float>(u.x), 65535.0f), __fdiv_rn(static_cast<float>(u.y), 65535.0f),
                                    __fdiv_rn(static_cast<float>(u.z), 65535.0f), __fdiv_rn(static_cast<float>(u.w), 65535.0f));
                }
            }
            reinterpret_cast<float4*>(dst)[q] = o;
        }
    } else {
        for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < hw; i += step) {
            float o = 0.0f;
            if (real) o = sizeof(T) == 4 ? static_cast<float>(src[i]) : __fdiv_rn(static_cast<float>(src[i]), 65535.0f);
            dst[i] = o;
        }
    }
}

// ------------------------------------------------------------------ live-work lists
// Which views and view pairs can reach the output?  HRNet.py:123-128 merges a pair as alice + alpha_bob * fuse(alice, bob),
// so a pair whose bob has alpha = 0 contributes nothing but alice, and everything that only feeds such pairs (the
// encoder work of zero-padded views, utils.py:89-95, and whole sub-trees of the fusion) never reaches the super-resolved
// image.  One thread per imageset walks the fusion tree backwards from view 0 of the last level:
//     need[l + 1][i]                        =>  need[l][i]                       (alice always survives)
//     need[l + 1][i] and alpha[top-1-i] != 0 =>  pair (l, i) is live, need[l][top-1-i]
// then the block compacts the flags into index lists (ascending, deterministic).  Layout of `lists`:
//     [0, LIVE_HDR)               counts: [0] = live encoder views, [1 + l] = live pairs of level l, [16 + l] = carried pairs
//     [LIVE_HDR, LIVE_HDR + B*L)  encoder list (image index b*L + v)
//     then per level l            pair list (pair index b*half_l + i), B*half_l entries
//     from LIVE_HDR + 2*B*L       per level l: carry list -- dead pairs whose alice the next level needs (the out-of-place
//                                 wavefront schedule copies those; the in-place three-launch schedule needs nothing)
// With skip == 0 every flag is set (dense lists; used by the stage-dump hook).
constexpr int LIVE_THREADS = 1024;

// `flag` may point to shared or global memory: a generic pointer, deliberately without __restrict__ / read-only hints.
__device__ void block_compact(const uint8_t* flag, int n, int* list, int* count) {
    __shared__ int warp_tot[LIVE_THREADS / 32];
    __shared__ int base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int start = 0; start < n; start += LIVE_THREADS) {
        const int i = start + threadIdx.x;
        const bool f = i < n && flag[i] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; ++w) off += warp_tot[w];
        if (f) list[off + __popc(bal & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < LIVE_THREADS / 32; ++w) tot += warp_tot[w];
            base += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

__global__ void __launch_bounds__(LIVE_THREADS)
live_lists_kernel(const float* __restrict__ alphas, int B, int L, int levels, int skip, int alpha_residual,
                  uint8_t* __restrict__ scratch_global, int use_smem, int* __restrict__ lists) {
    // scratch: need[l] at l * B * L (B * n_l entries), pair flags of level l at (levels + 1) * B * L + l * B * L;
    // kept in shared memory whenever it fits (always for Proba-V sized batches)
    extern __shared__ uint8_t scratch_smem[];
    uint8_t* scratch = use_smem ? scratch_smem : scratch_global;
    const size_t BL = static_cast<size_t>(B) * L;
    uint8_t* pair_flags = scratch + (levels + 1) * BL;
    uint8_t* carry_flags = pair_flags + levels * BL;     // dead pair whose alice is needed one level up (copied, not computed)
    for (int b = threadIdx.x; b < B; b += LIVE_THREADS) {
        int n_of[17];
        n_of[0] = L;
        for (int l = 0; l < levels; ++l) n_of[l + 1] = n_of[l] / 2;
        scratch[levels * BL + static_cast<size_t>(b) * n_of[levels]] = 1;        // n_levels == 1: the surviving view
        for (int l = levels - 1; l >= 0; --l) {
            const int n = n_of[l], half = n / 2, top = n - (n & 1);
            uint8_t* need = scratch + l * BL + static_cast<size_t>(b) * n;
            const uint8_t* need_up = scratch + (l + 1) * BL + static_cast<size_t>(b) * half;
            uint8_t* pf = pair_flags + l * BL + static_cast<size_t>(b) * half;
            for (int i = 0; i < n; ++i) need[i] = skip ? 0 : 1;
            for (int i = 0; i < half; ++i) {
                const bool wanted = !skip || need_up[i] != 0;
                const bool live = wanted && (!skip || !alpha_residual || alphas[static_cast<size_t>(b) * L + top - 1 - i] != 0.0f);
                pf[i] = live ? 1 : 0;
                carry_flags[l * BL + static_cast<size_t>(b) * half + i] = (wanted && !live) ? 1 : 0;
                if (wanted) need[i] = 1;
                if (live) need[top - 1 - i] = 1;
            }
        }
    }
    __syncthreads();
    block_compact(scratch, static_cast<int>(BL), lists + LIVE_HDR, lists);
    int off = LIVE_HDR + static_cast<int>(BL), n = L;
    for (int l = 0; l < levels; ++l) {
        const int half = n / 2;
        block_compact(pair_flags + l * BL, B * half, lists + off, lists + 1 + l);
        block_compact(carry_flags + l * BL, B * half, lists + off + static_cast<int>(BL), lists + 16 + l);
        off += B * half;
        n = half;
    }
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

int nhwc_bf16_to_nchw_f32_launch(const __nv_bfloat16* in, int n, int H, int W, int C, int group, int stride, float* out,
                                 cudaStream_t s) {
    const size_t hw = static_cast<size_t>(H) * W, total = hw * C * n;
    nhwc_to_nchw_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(in, hw, C, total, group, stride, out);
    note_launches(1);
    HRN_CUDA_OK(cudaGetLastError());
    return 0;
}

int u16_to_unit_float_launch(const uint16_t* in, size_t n, float* out, cudaStream_t s) {
    const size_t want = (n + 255) / 256;
    u16_to_unit_float_kernel<<<static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16), 256, 0, s>>>(in, n, out);
    note_launches(1);
    HRN_CUDA_OK(cudaGetLastError());
    return 0;
}

int unit_float_to_u16_launch(const float* in, size_t n, uint16_t* out, int* bad, cudaStream_t s) {
    const size_t want = (n + 255) / 256;
    if (bad != nullptr) HRN_CUDA_OK(cudaMemsetAsync(bad, 0, sizeof(int), s));
    unit_float_to_u16_kernel<<<static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16), 256, 0, s>>>(in, n, out, bad);
    note_launches(1);
    HRN_CUDA_OK(cudaGetLastError());
    return 0;
}

int collate_launch(const void* packed, int is_u16, const int* offsets, int B, int min_L, int H, int W, float* lrs,
                   float* alphas, cudaStream_t s) {
    const size_t hw = static_cast<size_t>(H) * W;
    if (B > 65535 || min_L > 65535) {
        set_error("collate: B and min_L must not exceed 65535");
        return -1;
    }
    if ((reinterpret_cast<uintptr_t>(packed) | reinterpret_cast<uintptr_t>(lrs)) & 15) {
        set_error("collate: buffers must be 16-byte aligned");
        return -1;
    }
    const size_t per_plane = (hw / 4 + 255) / 256;
    dim3 grid(static_cast<unsigned>(per_plane < 1 ? 1 : (per_plane > 64 ? 64 : per_plane)), min_L, B);
    if (is_u16)
        collate_kernel<uint16_t><<<grid, 256, 0, s>>>(static_cast<const uint16_t*>(packed), offsets, min_L, hw, lrs, alphas);
    else
        collate_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(packed), offsets, min_L, hw, lrs, alphas);
    note_launches(1);
    HRN_CUDA_OK(cudaGetLastError());
    return 0;
}

int live_levels(int L) {
    int levels = 0;
    for (int n = L; n / 2 > 0; n /= 2) ++levels;
    return levels;
}

size_t live_scratch_bytes(int B, int L) { return (3 * static_cast<size_t>(live_levels(L)) + 1) * B * L; }
size_t live_lists_ints(int B, int L) { return LIVE_HDR + 3 * static_cast<size_t>(B) * L; }

int live_lists_launch(const float* alphas, int B, int L, int skip, int alpha_residual, uint8_t* scratch, int* lists,
                      cudaStream_t s) {
    const int levels = live_levels(L);
    if (levels > 14) {
        set_error("hrn_forward: L=%d views need more than 14 fusion levels", L);
        return -1;
    }
    // A plain launch on purpose (no programmatic dependent launch): every later kernel of the forward pass reads the
    // lists without waiting, which is safe because the kernel after this one only starts once this one has finished.
    const size_t bytes = live_scratch_bytes(B, L);
    const int use_smem = bytes <= 40 * 1024;
    live_lists_kernel<<<1, LIVE_THREADS, use_smem ? bytes : 0, s>>>(alphas, B, L, levels, skip, alpha_residual, scratch,
                                                                    use_smem, lists);
    note_launches(1);
    HRN_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hrn
