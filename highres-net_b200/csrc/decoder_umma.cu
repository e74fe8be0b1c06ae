// Decoder (HRNet.py:147-156) on the tensor cores: ConvTranspose2d(64, 64, k=3, s=3) + PReLU + Conv2d(64, 1, k=1).
// stride == kernel, so nothing overlaps: every LR pixel yields its own 3x3 HR block,
//     sr[3y+ky, 3x+kx] = bf + sum_co wf[co] * PReLU(bd[co] + sum_ci x[ci] * Wd[ci, co, ky, kx])
// i.e. one GEMM  [pixels x 64] . [64 x 576]  followed by a per-pixel epilogue that never leaves the SM: the
// 64 x 3H x 3W deconv output of the reference (1.2 GB fp32 at B32 128x128, written and re-read there) is never
// materialised.  A = 128 consecutive pixels of one LR row (TMA, bf16 NHWC), B = the repacked deconv weight
// (576 rows (ky, kx, co) x 64 ci, bf16, 72 KB resident in smem), D = three N=192 accumulator groups (one per ky)
// ping-ponged between two TMEM buffers.  Epilogue warps 4..7 take even groups, warps 8..11 odd groups.
#include "umma_common.cuh"

#include <cstring>

namespace hrn {
namespace {

constexpr int TILE_M = 128;
constexpr int A_BYTES = TILE_M * 128;          // 128 px x 64 bf16
constexpr int A_RING = 4;
constexpr int W_BYTES = 576 * 128;             // 73,728
constexpr int GROUP_N = 192;                   // (kx, co) for one ky
constexpr int NUM_THREADS = 384;
constexpr int BAR_OFFSET = W_BYTES + A_RING * A_BYTES;
constexpr int PARAM_OFFSET = BAR_OFFSET + 256; // bd[64], wf[64]
constexpr int SMEM_BYTES = PARAM_OFFSET + 128 * 4 + 1024;
static_assert(16 * A_RING + 16 + 16 + 8 + 8 <= 256, "barrier block overflows into the parameter array");

struct DecArgs {
    int B, H, W, x_tiles;
    uint32_t tiles;                            // B * H * x_tiles (31 bits, checked by the launcher)
    const uint8_t* w_img;                      // pre-swizzled B image (device)
    const float* bd;                           // deconv bias (64)
    const float* wf;                           // final 1x1 weights (64)
    float prelu, bf;
    float* out;                                // (B, 3H, 3W) fp32
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
decoder_umma_kernel(const __grid_constant__ CUtensorMap in_map, const DecArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_s = base, ring_s = base + W_BYTES, bars = base + BAR_OFFSET;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * A_RING, bar_tfull = bars + 16 * A_RING;
    const uint32_t bar_tempty = bar_tfull + 16, bar_w = bar_tempty + 16, tmem_slot = bar_w + 8;
    uint8_t* smem_gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));
    float* bd_s = reinterpret_cast<float*>(smem_gen + PARAM_OFFSET);
    float* wf_s = bd_s + 64;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < A_RING; ++i) {
            ptx::mbar_init(bar_full + 8 * i, 1);
            ptx::mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(bar_tfull + 8 * i, 1);
            ptx::mbar_init(bar_tempty + 8 * i, 4);
        }
        ptx::mbar_init(bar_w, 1);
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&in_map);
    }
    ptx::pdl_launch_dependents();
    if (warp == 2) ptx::tmem_alloc<512>(tmem_slot);
    if (threadIdx.x >= 128 && threadIdx.x < 192) {
        bd_s[threadIdx.x - 128] = a.bd[threadIdx.x - 128];
        wf_s[threadIdx.x - 128] = a.wf[threadIdx.x - 128];
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(bar_w, W_BYTES);
            for (int off = 0; off < W_BYTES; off += 8192) ptx::bulk_copy_g2s(w_s + off, a.w_img + off, 8192, bar_w);
            ptx::pdl_wait();
            uint32_t it = 0;
            for (uint32_t t = blockIdx.x; t < a.tiles; t += gridDim.x, ++it) {
                const uint32_t per_img = static_cast<uint32_t>(a.x_tiles) * a.H;
                const uint32_t bi = t / per_img, rem = t - bi * per_img, yy = a.x_tiles == 1 ? rem : rem / a.x_tiles;
                const int xt = static_cast<int>(rem - yy * a.x_tiles), y = static_cast<int>(yy), b = static_cast<int>(bi);
                const uint32_t slot = it % A_RING;
                ptx::mbar_wait(bar_empty + 8 * slot, ((it / A_RING) & 1) ^ 1, 1);
                ptx::mbar_expect_tx(bar_full + 8 * slot, A_BYTES);
                ptx::tma_load_4d(ring_s + slot * A_BYTES, &in_map, 0, xt * TILE_M, y, b, bar_full + 8 * slot);
            }
        }
    } else if (warp == 1) {
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(TILE_M, GROUP_N);
            const uint32_t a_lo0 = desc_lo(ring_s), b_lo0 = desc_lo(w_s);
            ptx::mbar_wait(bar_w, 0, 2);
            uint32_t it = 0, grp = 0;
            for (uint32_t t = blockIdx.x; t < a.tiles; t += gridDim.x, ++it) {
                const uint32_t slot = it % A_RING;
                ptx::mbar_wait(bar_full + 8 * slot, (it / A_RING) & 1, 3);
                ptx::tc_fence_after();
#pragma unroll
                for (int g = 0; g < 3; ++g, ++grp) {
                    const uint32_t buf = grp & 1;
                    ptx::mbar_wait(bar_tempty + 8 * buf, ((grp >> 1) & 1) ^ 1, 4);
                    ptx::tc_fence_after();
                    uint64_t ad = make_desc(a_lo0 + slot * (A_BYTES / 16));
                    uint64_t bd = make_desc(b_lo0 + g * (GROUP_N * 128 / 16));
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        ptx::umma_bf16(tmem_base + buf * 256, ad, bd, idesc, j > 0 ? 1u : 0u);
                        ad += 2;
                        bd += 2;
                    }
                    ptx::umma_commit(bar_tfull + 8 * buf);
                }
                ptx::umma_commit(bar_empty + 8 * slot);
            }
        }
    } else if (warp >= 4) {
        const int wq = warp & 3;            // TMEM lane quadrant
        const int set = (warp - 4) >> 2;    // 0: even groups, 1: odd groups
        const float slope_m1 = a.prelu - 1.0f;
        ptx::pdl_wait();
        uint32_t grp = 0;
        for (uint32_t t = blockIdx.x; t < a.tiles; t += gridDim.x) {
            const uint32_t per_img = static_cast<uint32_t>(a.x_tiles) * a.H;
            const uint32_t bi = t / per_img, rem = t - bi * per_img, yy = a.x_tiles == 1 ? rem : rem / a.x_tiles;
            const int xt = static_cast<int>(rem - yy * a.x_tiles), y = static_cast<int>(yy), b = static_cast<int>(bi);
            const int x = xt * TILE_M + wq * 32 + lane;
#pragma unroll 1
            for (int g = 0; g < 3; ++g, ++grp) {
                if (static_cast<int>(grp & 1) != set) continue;
                const uint32_t buf = grp & 1;
                ptx::mbar_wait(bar_tfull + 8 * buf, (grp >> 1) & 1, 5);
                ptx::tc_fence_after();
                float sr[3];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    float acc = a.bf;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t v[32];
                        ptx::tmem_ld_x32(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + buf * 256 + kx * 64 + h * 32, v);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            float val = __uint_as_float(v[e]) + bd_s[h * 32 + e];
                            val = fmaf(slope_m1, fminf(val, 0.0f), val);          // PReLU(v) = v + (slope - 1) min(v, 0)
                            acc = fmaf(wf_s[h * 32 + e], val, acc);
                        }
                    }
                    sr[kx] = acc;
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * buf);
                if (x < a.W) {
                    float* op = a.out + (static_cast<size_t>(b) * 3 * a.H + 3 * y + g) * (3 * static_cast<size_t>(a.W)) + 3 * x;
                    op[0] = sr[0];
                    op[1] = sr[1];
                    op[2] = sr[2];
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

}  // namespace

int decoder_weight_image_bytes() { return W_BYTES; }

// ConvTranspose2d weight (ci, co, ky, kx) fp32 -> B image: row n = (ky*3 + kx)*64 + co, k = ci, bf16, K-major rows of
// 128 B with the 16-byte chunks XOR-swizzled by (row % 8) (SWIZZLE_128B).
void decoder_pack_weights(const float* w, uint8_t* dst) {
    for (int pos = 0; pos < 9; ++pos)
        for (int co = 0; co < 64; ++co)
            for (int ci = 0; ci < 64; ++ci) {
                const int row = pos * 64 + co;
                const __nv_bfloat16 h = __float2bfloat16_rn(w[(static_cast<size_t>(ci) * 64 + co) * 9 + pos]);
                const size_t off = static_cast<size_t>(row) * 128 + (((ci >> 3) ^ (row & 7)) << 4) + (ci & 7) * 2;
                std::memcpy(dst + off, &h, 2);
            }
}

int decoder_umma_launch(const __nv_bfloat16* in, int B, int image_stride, int H, int W, const uint8_t* w_img, const float* bd,
                        float prelu, const float* wf, float bf, float* out, int sm_count, cudaStream_t s) {
    DecArgs a;
    a.B = B;
    a.H = H;
    a.W = W;
    a.x_tiles = (W + TILE_M - 1) / TILE_M;
    const long long tiles64 = static_cast<long long>(B) * H * a.x_tiles;
    if (tiles64 >= (1LL << 31) - sm_count) {
        set_error("decoder: %lld row tiles exceed the 31-bit tile index; split the batch", tiles64);
        return -1;
    }
    a.tiles = static_cast<uint32_t>(tiles64);
    a.w_img = w_img;
    a.bd = bd;
    a.wf = wf;
    a.prelu = prelu;
    a.bf = bf;
    a.out = out;
    CUtensorMap map;
    if (encode_nhwc_map(&map, in, 64, W, H, B, TILE_M, image_stride)) return -1;
    static bool attr_set[64] = {};
    if (allow_dynamic_smem(decoder_umma_kernel, SMEM_BYTES, attr_set)) return -1;
    const int ctas = static_cast<int>(a.tiles < sm_count ? a.tiles : sm_count);
    HRN_CUDA_OK(launch_pdl(decoder_umma_kernel, ctas, NUM_THREADS, SMEM_BYTES, s, 1, map, a));
    note_launches(1);
    return 0;
}

}  // namespace hrn
