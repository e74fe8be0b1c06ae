// First encoder layer (HRNet.py:51-53) on the tensor cores: Conv2d(2 -> 64, k3, p1) + PReLU over (view, anchor)
// pairs, with the repeat / cat / view of HRNet.py:201-204 fused away (channel 1 is the per-imageset anchor).
//
// K is only 18, far too thin for a TMA-fed implicit GEMM, so the A operand is BUILT in shared memory: four
// builder warps (thread = pixel) gather the 2 x 3 x 3 fp32 neighbourhood, split every value into hi + lo bf16
// parts (x = hi + lo keeps ~16 mantissa bits of the fp32 input image) and write one K-major SWIZZLE_128B row
//     [ hi(18) | lo(18) | 0 (12) ]            K = 48 = three UMMA k-steps
// per pixel; B = [ w(18) | w(18) | 0 ] in bf16 (8 KB, resident).  One tile = 128 pixels of one image row,
// D = 128 x 64 fp32 in one of eight TMEM slots; four epilogue warps apply bias + PReLU and store bf16 NHWC.
// Versus the CUDA-core version (1152 FMA + 288 LDS per pixel) this is ~5x fewer instructions per pixel and the
// kernel becomes bound by its 1.07 GB of output.
#include "umma_common.cuh"

#include <cstring>

namespace hrn {
namespace {

constexpr int TILE_M = 128;
constexpr int A_BYTES = TILE_M * 128;
constexpr int A_RING = 4;
constexpr int W_BYTES = 64 * 128;
constexpr int ACC_SLOTS = 8;
constexpr int NUM_THREADS = 128 + 256 + 256;    // warp 0: MMA, warp 1: TMEM alloc, warps 4-11: builders (two sets), warps 12-19: epilogue
constexpr int BAR_OFFSET = W_BYTES + A_RING * A_BYTES;
constexpr int BIAS_OFFSET = BAR_OFFSET + 512;   // barriers: (2 * A_RING + 2 * ACC_SLOTS) * 8 B + TMEM slot
constexpr int SMEM_BYTES = BIAS_OFFSET + 64 * 4 + 1024;
static_assert((2 * A_RING + 2 * ACC_SLOTS) * 8 + 8 <= 512, "barrier block overflows into the bias array");

struct InitArgs {
    const float* lrs;       // (B, L, H, W)
    const float* anchor;    // (B, H, W)
    int L, H, W, x_tiles;
    uint32_t tiles;         // B * L * H * x_tiles (checked to fit 31 bits by the launcher: tile arithmetic stays 32-bit,
                            // a 64-bit division per tile and warp used to be a third of this kernel's instructions)
    const uint8_t* w_img;   // conv_init_pack_weights() image
    const float* bias;
    float prelu;
    __nv_bfloat16* out;     // (B * L, H, W, 64)
    const int* live_list;   // live-work list (pointwise.cu): images live_list[0 .. *live_count), nullptr = all B * L
    const int* live_count;
};

// Walks the row tiles t0, t0 + step, t0 + 2 step, ... (flattened (image, row, column tile) order) without a division
// per tile.  The roles keep one cursor a tile AHEAD of the one they work on, so that the live-work list entry of the
// next image (a dependent L2 load) has a whole tile time to arrive.  (CTAs take interleaved tiles on purpose: giving
// every CTA a contiguous run of rows instead measured 15 % slower, 148 scattered 16 KB write streams.)
struct TileCursor {
    uint32_t idx, rem, per_img;      // position in the (listed) image sequence, tile within the image
    int m;                           // image index b * L + v
    __device__ TileCursor(uint32_t t, const InitArgs& a) : per_img(static_cast<uint32_t>(a.x_tiles) * a.H) {
        idx = t / per_img;
        rem = t - idx * per_img;
        m = a.live_list != nullptr ? a.live_list[idx] : static_cast<int>(idx);
    }
    __device__ void advance(uint32_t step, const InitArgs& a) {
        rem += step;
        if (rem >= per_img) {
            do {
                rem -= per_img;
                ++idx;
            } while (rem >= per_img);
            m = a.live_list != nullptr ? a.live_list[idx] : static_cast<int>(idx);
        }
    }
    __device__ void where(const InitArgs& a, int& y, int& xt) const {
        const uint32_t yy = a.x_tiles == 1 ? rem : rem / a.x_tiles;
        y = static_cast<int>(yy);
        xt = static_cast<int>(rem - yy * a.x_tiles);
    }
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_init_umma_kernel(const InitArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_s = base, ring_s = base + W_BYTES, bars = base + BAR_OFFSET;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * A_RING, bar_tfull = bars + 16 * A_RING;
    const uint32_t bar_tempty = bar_tfull + 8 * ACC_SLOTS, tmem_slot = bar_tempty + 8 * ACC_SLOTS;
    uint8_t* smem_gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));
    float* bias_s = reinterpret_cast<float*>(smem_gen + BIAS_OFFSET);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles = a.live_count != nullptr ? static_cast<uint32_t>(*a.live_count) * a.H * a.x_tiles : a.tiles;
    if (threadIdx.x == 0) {
        for (int i = 0; i < A_RING; ++i) {
            ptx::mbar_init(bar_full + 8 * i, 128);       // every builder thread arrives
            ptx::mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < ACC_SLOTS; ++i) {
            ptx::mbar_init(bar_tfull + 8 * i, 1);
            ptx::mbar_init(bar_tempty + 8 * i, 8);
        }
        ptx::fence_barrier_init();
    }
    ptx::pdl_launch_dependents();
    if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
    // weights (8 KB) and bias: plain loads, then made visible to the tensor-core (async) proxy
    for (int i = threadIdx.x; i < W_BYTES / 16; i += NUM_THREADS)
        reinterpret_cast<uint4*>(smem_gen)[i] = __ldg(reinterpret_cast<const uint4*>(a.w_img) + i);
    if (threadIdx.x < 64) bias_s[threadIdx.x] = a.bias[threadIdx.x];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        // ===================================================== MMA issuer
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(TILE_M, 64);
            const uint32_t a_lo0 = desc_lo(ring_s), b_lo0 = desc_lo(w_s);
            uint32_t it = 0;
            for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
                const uint32_t slot = it % A_RING, acc = it % ACC_SLOTS;
                ptx::mbar_wait(bar_tempty + 8 * acc, ((it / ACC_SLOTS) & 1) ^ 1, 4);
                ptx::mbar_wait(bar_full + 8 * slot, (it / A_RING) & 1, 3);
                ptx::tc_fence_after();
                uint64_t ad = make_desc(a_lo0 + slot * (A_BYTES / 16)), bd = make_desc(b_lo0);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    ptx::umma_bf16(tmem_base + acc * 64, ad, bd, idesc, j > 0 ? 1u : 0u);
                    ad += 2;
                    bd += 2;
                }
                ptx::umma_commit(bar_empty + 8 * slot);
                ptx::umma_commit(bar_tfull + 8 * acc);
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ===================================================== A builders: two sets of four warps take alternate tiles;
        // thread = pixel.  The 18 input values of the NEXT tile of the set are requested before the current row is
        // written, so the global-load latency overlaps the barrier wait and the other set's work.
        const int set = (warp - 4) >> 2;
        const int px = ((warp - 4) & 3) * 32 + lane;                    // row of the A tile
        const size_t hw = static_cast<size_t>(a.H) * a.W;
        auto gather = [&](const TileCursor& cur, float (&v)[18]) {
            int y, xt;
            cur.where(a, y, xt);
            const int m = cur.m;
            const int x = xt * TILE_M + px;
            const float* src[2] = {a.lrs + static_cast<size_t>(m) * hw, a.anchor + static_cast<size_t>(m / a.L) * hw};
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = y + ky - 1;
                const bool oky = yy >= 0 && yy < a.H;
                const int off = yy * a.W + x - 1;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const bool ok = oky && (x + kx - 1) >= 0 && (x + kx - 1) < a.W;
                    v[ky * 3 + kx] = ok ? __ldg(src[0] + off + kx) : 0.0f;
                    v[9 + ky * 3 + kx] = ok ? __ldg(src[1] + off + kx) : 0.0f;
                }
            }
        };
        float v[18];
        ptx::pdl_wait();                 // the anchor comes from the previous kernel
        const uint32_t first = blockIdx.x + static_cast<uint32_t>(set) * gridDim.x;
        const uint32_t stride = 2u * gridDim.x;
        TileCursor nxt(first < tiles ? first : 0u, a);          // the tile whose inputs are gathered next
        if (first < tiles) gather(nxt, v);
        if (first + stride < tiles) nxt.advance(stride, a);
        uint32_t it = set;
        for (uint32_t t = first; t < tiles; t += stride, it += 2) {
            // k = 0..17 hi parts, 18..35 lo parts, 36..47 zero
            __align__(16) __nv_bfloat16 row[48];
#pragma unroll
            for (int k = 0; k < 18; ++k) {
                const __nv_bfloat16 hi = __float2bfloat16_rn(v[k]);
                row[k] = hi;
                row[18 + k] = __float2bfloat16_rn(v[k] - __bfloat162float(hi));
            }
#pragma unroll
            for (int k = 36; k < 48; ++k) row[k] = __float2bfloat16_rn(0.0f);
            if (t + stride < tiles) {
                gather(nxt, v);
                if (t + 2 * stride < tiles) nxt.advance(stride, a);
            }
            const uint32_t slot = it % A_RING;
            ptx::mbar_wait(bar_empty + 8 * slot, ((it / A_RING) & 1) ^ 1, 1);
            uint8_t* dst = smem_gen + W_BYTES + slot * A_BYTES + px * 128;
#pragma unroll
            for (int j = 0; j < 6; ++j)       // 16-byte chunk j of the K-major row goes to chunk (j ^ (row % 8)): SWIZZLE_128B
                *reinterpret_cast<uint4*>(dst + ((j ^ (px & 7)) << 4)) = *reinterpret_cast<const uint4*>(&row[8 * j]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            ptx::mbar_arrive(bar_full + 8 * slot);
        }
    } else if (warp >= 12) {
        // ===================================================== epilogue: 8 warps = (lane quadrant) x (column half);
        // bias + PReLU in fp32 -> bf16 NHWC, 256-bit stores
        const int wq = warp & 3;
        const int hf = (warp - 12) >> 2;
        const float slope_m1 = a.prelu - 1.0f;
        const size_t hw = static_cast<size_t>(a.H) * a.W;
        float bias_r[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) bias_r[e] = bias_s[hf * 32 + e];
        ptx::pdl_wait();                 // the output buffer may still be read by the previous forward's kernels
        uint32_t it = 0;
        TileCursor cur(blockIdx.x < tiles ? blockIdx.x : 0u, a), nxt = cur;
        if (blockIdx.x + gridDim.x < tiles) nxt.advance(gridDim.x, a);
        for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
            int y, xt;
            cur.where(a, y, xt);
            const int m = cur.m;
            cur = nxt;
            if (t + 2 * gridDim.x < tiles) nxt.advance(gridDim.x, a);
            const int x = xt * TILE_M + wq * 32 + lane;
            const uint32_t acc = it % ACC_SLOTS;
            ptx::mbar_wait(bar_tfull + 8 * acc, (it / ACC_SLOTS) & 1, 5);
            ptx::tc_fence_after();
            uint32_t v[32];
            ptx::tmem_ld_x32(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * 64 + hf * 32, v);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * acc);
            uint32_t o[2][8];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                float x0 = __uint_as_float(v[2 * e]) + bias_r[2 * e];
                float x1 = __uint_as_float(v[2 * e + 1]) + bias_r[2 * e + 1];
                x0 = fmaf(slope_m1, fminf(x0, 0.0f), x0);
                x1 = fmaf(slope_m1, fminf(x1, 0.0f), x1);
                const __nv_bfloat162 p = __floats2bfloat162_rn(x0, x1);
                o[e >> 3][e & 7] = *reinterpret_cast<const uint32_t*>(&p);
            }
            if (x < a.W) {
                __nv_bfloat16* op = a.out + (static_cast<size_t>(m) * hw + static_cast<size_t>(y) * a.W + x) * 64 + hf * 32;
                ptx::stg_v8(op, o[0]);
                ptx::stg_v8(op + 16, o[1]);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 1) ptx::tmem_dealloc<512>(tmem_base);
}

}  // namespace

int conv_init_weight_image_bytes() { return W_BYTES; }

// Conv2d weight (64, 2, 3, 3) fp32 -> B image: row n = co, k = [w(ci*9+tap) for 18 | the same 18 again | 0 ...],
// bf16, K-major 128-byte rows, 16-byte chunks XOR-swizzled by (row % 8).
void conv_init_pack_weights(const float* w, uint8_t* dst) {
    std::memset(dst, 0, W_BYTES);
    for (int co = 0; co < 64; ++co)
        for (int k = 0; k < 36; ++k) {
            const __nv_bfloat16 h = __float2bfloat16_rn(w[co * 18 + (k % 18)]);
            const size_t off = static_cast<size_t>(co) * 128 + (((k >> 3) ^ (co & 7)) << 4) + (k & 7) * 2;
            std::memcpy(dst + off, &h, 2);
        }
}

int conv_init_umma_launch(const float* lrs, const float* anchor, int B, int L, int H, int W, const uint8_t* w_img,
                          const float* bias, float prelu, __nv_bfloat16* out, const int* live_list, const int* live_count,
                          int sm_count, cudaStream_t s) {
    InitArgs a;
    a.lrs = lrs;
    a.anchor = anchor;
    a.L = L;
    a.H = H;
    a.W = W;
    a.x_tiles = (W + TILE_M - 1) / TILE_M;
    const long long tiles64 = static_cast<long long>(B) * L * H * a.x_tiles;
    if (tiles64 >= (1LL << 31) - 2 * sm_count) {
        set_error("conv_init: %lld row tiles exceed the 31-bit tile index; split the batch", tiles64);
        return -1;
    }
    a.tiles = static_cast<uint32_t>(tiles64);
    a.live_list = live_list;
    a.live_count = live_count;
    a.w_img = w_img;
    a.bias = bias;
    a.prelu = prelu;
    a.out = out;
    static bool attr_set[64] = {};
    if (allow_dynamic_smem(conv_init_umma_kernel, SMEM_BYTES, attr_set)) return -1;
    const int ctas = static_cast<int>(a.tiles < static_cast<uint32_t>(sm_count) ? a.tiles : sm_count);
    HRN_CUDA_OK(launch_pdl(conv_init_umma_kernel, ctas, NUM_THREADS, SMEM_BYTES, s, 1, a));
    note_launches(1);
    return 0;
}

}  // namespace hrn
