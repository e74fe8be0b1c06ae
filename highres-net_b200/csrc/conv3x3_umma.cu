// 3x3 / pad-1 convolution over bf16 NHWC activations as an implicit GEMM on the
// 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
// Replaces the cuDNN convs behind nn.Conv2d(64->64), (128->128), (128->64) of the
// reference (HRNet.py:18-21, 58-60, 94-97) and fuses what surrounds them there:
// bias, PReLU, the ResidualBlock skip (HRNet.py:32-33), the channel concat of a
// view pair (HRNet.py:114-119, done by addressing: the two 64-channel K chunks of
// the A operand come from two different views) and the alpha-masked merge
// `alice + alpha_bob * x` (HRNet.py:123-128).
//
// Mapping.  One UMMA tile = 128 consecutive output pixels of one image row (M)
// x N_TILE output channels (N), K = 9 taps x CIN.  Per tap the A operand is the
// SAME shared-memory row buffer read through a descriptor whose start address is
// shifted by (kx) pixels = kx * 128 B; ky selects one of three resident input
// rows.  Each CTA walks a vertical strip of rows, so an input row is fetched by
// TMA once (130 pixels: 128 + halo, out-of-bounds pixels zero-filled by TMA = the
// conv padding) and reused by the three output rows that touch it.  The weight
// slice of the CTA (9 x CIN x N_TILE bf16, 72 KB) stays resident in smem.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one
// thread), warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM -> registers ->
// bias/PReLU/residual -> bf16 NHWC global).  TMEM holds ACC_STAGES accumulators
// so the epilogue of tile t overlaps the MMAs of tiles t+1...
#include "internal.h"
#include "ptx.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace hrn {
namespace {

constexpr int TILE_M = 128;
constexpr int SLOT_PIX = TILE_M + 2;
constexpr int CHUNK_BYTES = 17408;   // 130 px * 128 B = 16640, rounded up to 1024 (keeps the SW128 phase)
constexpr int NUM_THREADS = 256;
constexpr int ACC_STAGES = 4;

template <int CIN>
struct Cfg {
    static constexpr int CHUNKS = CIN / 64;                 // 64-channel (128-byte) K chunks per pixel
    static constexpr int NT = (CIN == 64) ? 64 : 32;        // output channels per CTA (UMMA N)
    static constexpr int RING = (CIN == 64) ? 8 : 4;        // resident input rows
    static constexpr int SLOT_BYTES = CHUNKS * CHUNK_BYTES;
    static constexpr int WTILE_BYTES = NT * 128;            // one (tap, chunk) B tile: NT rows x 64 bf16
    static constexpr int W_BYTES = 9 * CHUNKS * WTILE_BYTES;
    static constexpr int TMEM_COLS = ACC_STAGES * NT;
    static constexpr int BAR_OFFSET = W_BYTES + RING * SLOT_BYTES;
    static constexpr int BIAS_OFFSET = BAR_OFFSET + 256;
    static constexpr int SMEM_BYTES = BIAS_OFFSET + NT * 4 + 1024;
    static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS >= 32 && TMEM_COLS <= 512, "TMEM columns");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct Geometry {
    int n_parts, items_per_part, strips, x_tiles;
};

struct Item {
    int m, xt, y0, rows;
};

__device__ __forceinline__ Item decode_item(int t, const ConvArgs& a, const Geometry& g) {
    Item it;
    it.xt = t % g.x_tiles;
    const int s = (t / g.x_tiles) % g.strips;
    it.m = t / (g.x_tiles * g.strips);
    it.y0 = s * a.strip_h;
    it.rows = min(a.strip_h, a.H - it.y0);
    return it;
}

template <int CIN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap in_map, const ConvArgs a, const Geometry g) {
    using C = Cfg<CIN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_s = base;
    const uint32_t ring_s = base + C::W_BYTES;
    const uint32_t bars = base + C::BAR_OFFSET;
    const uint32_t bar_full = bars;                                  // [RING]
    const uint32_t bar_empty = bars + 8 * C::RING;                   // [RING]
    const uint32_t bar_tfull = bars + 16 * C::RING;                  // [ACC_STAGES]
    const uint32_t bar_tempty = bar_tfull + 8 * ACC_STAGES;          // [ACC_STAGES]
    const uint32_t bar_w = bar_tempty + 8 * ACC_STAGES;
    const uint32_t tmem_slot = bar_w + 8;
    uint8_t* smem_gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));
    float* bias_s = reinterpret_cast<float*>(smem_gen + C::BIAS_OFFSET);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int part = blockIdx.x % g.n_parts;
    const int first = blockIdx.x / g.n_parts;
    const int step = gridDim.x / g.n_parts;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::RING; ++i) {
            ptx::mbar_init(bar_full + 8 * i, 1);
            ptx::mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < ACC_STAGES; ++i) {
            ptx::mbar_init(bar_tfull + 8 * i, 1);
            ptx::mbar_init(bar_tempty + 8 * i, 4);   // one arrive per epilogue warp
        }
        ptx::mbar_init(bar_w, 1);
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&in_map);
    }
    if (warp == 2) ptx::tmem_alloc<C::TMEM_COLS>(tmem_slot);
    if (threadIdx.x >= 128 && threadIdx.x < 128 + C::NT) bias_s[threadIdx.x - 128] = a.bias[part * C::NT + threadIdx.x - 128];
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            ptx::mbar_expect_tx(bar_w, C::W_BYTES);
            const uint8_t* wsrc = a.w_img + static_cast<size_t>(part) * C::W_BYTES;
            for (int off = 0; off < C::W_BYTES; off += C::WTILE_BYTES)
                ptx::bulk_copy_g2s(w_s + off, wsrc + off, C::WTILE_BYTES, bar_w);
            uint32_t it = 0;
            for (int t = first; t < g.items_per_part; t += step) {
                const Item w = decode_item(t, a, g);
                int img[2], ch[2];
                if (a.pair_mode) {
                    const int b = w.m / a.half, i = w.m % a.half;
                    img[0] = b * a.src_views + i;
                    img[1] = b * a.src_views + (a.top - 1 - i);
                    ch[0] = ch[1] = 0;
                } else {
                    img[0] = img[1] = w.m;
                    ch[0] = 0;
                    ch[1] = 64;
                }
                for (int r = 0; r < w.rows + 2; ++r, ++it) {
                    const uint32_t slot = it % C::RING, ph = (it / C::RING) & 1;
                    ptx::mbar_wait(bar_empty + 8 * slot, ph ^ 1, 1);
                    ptx::mbar_expect_tx(bar_full + 8 * slot, C::CHUNKS * SLOT_PIX * 128);
                    const int y = w.y0 - 1 + r;
#pragma unroll
                    for (int c = 0; c < C::CHUNKS; ++c)
                        ptx::tma_load_4d(ring_s + slot * C::SLOT_BYTES + c * CHUNK_BYTES, &in_map, ch[c],
                                         w.xt * TILE_M - 1, y, img[c], bar_full + 8 * slot);
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(TILE_M, C::NT);
            ptx::mbar_wait(bar_w, 0, 2);
            uint32_t it = 0, tile = 0;
            auto wait_full = [&](uint32_t k) {
                ptx::mbar_wait(bar_full + 8 * (k % C::RING), (k / C::RING) & 1, 3);
            };
            for (int t = first; t < g.items_per_part; t += step) {
                const Item w = decode_item(t, a, g);
                for (int i = 0; i < w.rows; ++i, ++tile) {
                    const uint32_t acc = tile % ACC_STAGES, aph = (tile / ACC_STAGES) & 1;
                    ptx::mbar_wait(bar_tempty + 8 * acc, aph ^ 1, 4);
                    if (i == 0) {
                        wait_full(it);
                        wait_full(it + 1);
                    }
                    wait_full(it + i + 2);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * C::NT;
                    uint32_t accumulate = 0;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const uint32_t a_row = ring_s + ((it + i + ky) % C::RING) * C::SLOT_BYTES;
#pragma unroll
                        for (int c = 0; c < C::CHUNKS; ++c) {
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const uint32_t a_addr = a_row + c * CHUNK_BYTES + kx * 128;
                                const uint32_t b_addr = w_s + ((ky * 3 + kx) * C::CHUNKS + c) * C::WTILE_BYTES;
                                const uint32_t bo = a.desc_base_offset_mode ? ((a_addr >> 7) & 7) : 0;
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    ptx::umma_bf16(d_tmem, ptx::smem_desc_sw128(a_addr + j * 32, bo),
                                                   ptx::smem_desc_sw128(b_addr + j * 32, 0), idesc, accumulate);
                                    accumulate = 1;
                                }
                            }
                        }
                    }
                    ptx::umma_commit(bar_empty + 8 * ((it + i) % C::RING));
                    if (i == w.rows - 1) {
                        ptx::umma_commit(bar_empty + 8 * ((it + i + 1) % C::RING));
                        ptx::umma_commit(bar_empty + 8 * ((it + i + 2) % C::RING));
                    }
                    ptx::umma_commit(bar_tfull + 8 * acc);
                }
                it += w.rows + 2;
            }
        }
    } else if (warp >= 4) {
        // ===================================================== epilogue
        const int wq = warp - 4;   // TMEM lane quadrant of this warp (warp % 4)
        const int co0 = part * C::NT;
        constexpr int VEC = C::NT / 8;   // 16-byte vectors of bf16 per pixel
        uint32_t tile = 0;
        for (int t = first; t < g.items_per_part; t += step) {
            const Item w = decode_item(t, a, g);
            const int x = w.xt * TILE_M + wq * 32 + lane;
            const bool valid = x < a.W;
            // residual source for this work item
            const __nv_bfloat16* res_img = nullptr;
            int res_c = 0;
            float scale = 1.0f;
            if (a.res_mode == RES_SAME) {
                res_img = a.res + (static_cast<size_t>(w.m) * a.H * a.W) * a.cout + co0;
                res_c = a.cout;
            } else if (a.res_mode == RES_PAIR) {
                const int b = w.m / a.half, i = w.m % a.half;
                const int side = co0 >= 64;
                const int img = b * a.src_views + (side ? (a.top - 1 - i) : i);
                res_img = a.res + (static_cast<size_t>(img) * a.H * a.W) * 64 + (co0 - 64 * side);
                res_c = 64;
            } else if (a.res_mode == RES_ALPHA) {
                const int b = w.m / a.half, i = w.m % a.half;
                res_img = a.res + (static_cast<size_t>(b * a.src_views + i) * a.H * a.W) * 64 + co0;
                res_c = 64;
                scale = a.alphas[b * a.alpha_stride + (a.top - 1 - i)];
            }
            for (int i = 0; i < w.rows; ++i, ++tile) {
                const uint32_t acc = tile % ACC_STAGES, aph = (tile / ACC_STAGES) & 1;
                const int y = w.y0 + i;
                const size_t pix = static_cast<size_t>(y) * a.W + x;
                uint4 rv[VEC];
                if (res_img != nullptr && valid) {
                    const uint4* rp = reinterpret_cast<const uint4*>(res_img + pix * res_c);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) rv[v] = __ldg(rp + v);
                }
                ptx::mbar_wait(bar_tfull + 8 * acc, aph, 5);
                ptx::tc_fence_after();
                uint32_t v[C::NT / 32][32];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * C::NT;
#pragma unroll
                for (int h = 0; h < C::NT / 32; ++h) ptx::tmem_ld_x32(taddr + h * 32, v[h]);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * acc);
                if (valid) {
                    uint4* op = reinterpret_cast<uint4*>(a.out + (static_cast<size_t>(w.m) * a.H * a.W + pix) * a.cout + co0);
#pragma unroll
                    for (int vv = 0; vv < VEC; ++vv) {
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int ch = vv * 8 + e;
                            float val = __uint_as_float(v[ch / 32][ch % 32]) + bias_s[ch];
                            if (a.has_prelu) val = val >= 0.0f ? val : a.prelu * val;
                            f[e] = val;
                        }
                        if (res_img != nullptr) {
                            const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv[vv]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 rf = __bfloat1622float2(r2[e]);
                                f[2 * e] = rf.x + scale * f[2 * e];
                                f[2 * e + 1] = rf.y + scale * f[2 * e + 1];
                            }
                        }
                        uint4 o;
                        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                        for (int e = 0; e < 4; ++e) o2[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                        op[vv] = o;
                    }
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 2) ptx::tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

template <int CIN>
int launch_impl(const ConvArgs& a, const CUtensorMap& map, const Geometry& g, int ctas, cudaStream_t stream) {
    using C = Cfg<CIN>;
    static bool attr_set = false;
    if (!attr_set) {
        HRN_CUDA_OK(cudaFuncSetAttribute(conv3x3_umma_kernel<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES));
        attr_set = true;
    }
    conv3x3_umma_kernel<CIN><<<ctas, NUM_THREADS, C::SMEM_BYTES, stream>>>(map, a, g);
    note_launches(1);
    HRN_CUDA_OK(cudaGetLastError());
    return 0;
}

inline int n_tile_for(int cin) { return cin == 64 ? 64 : 32; }

}  // namespace

int conv3x3_bytes_per_weight_image(int cin, int cout) { return 9 * cin * cout * 2; }

void conv3x3_pack_weights(const float* oihw, int cin, int cout, uint8_t* dst) {
    const int nt = n_tile_for(cin), chunks = cin / 64, parts = cout / nt;
    const size_t tile_bytes = static_cast<size_t>(nt) * 128;
    for (int p = 0; p < parts; ++p)
        for (int tap = 0; tap < 9; ++tap)
            for (int c = 0; c < chunks; ++c) {
                uint8_t* tile = dst + ((static_cast<size_t>(p) * 9 + tap) * chunks + c) * tile_bytes;
                for (int n = 0; n < nt; ++n)
                    for (int k = 0; k < 64; ++k) {
                        const int co = p * nt + n, ci = c * 64 + k;
                        const float w = oihw[(static_cast<size_t>(co) * cin + ci) * 9 + tap];
                        const __nv_bfloat16 h = __float2bfloat16_rn(w);
                        // K-major row of 64 bf16 = 128 B, 16-byte chunks XOR-swizzled by (row % 8): SWIZZLE_128B
                        const size_t off = static_cast<size_t>(n) * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
                        std::memcpy(tile + off, &h, 2);
                    }
            }
}

int conv3x3_launch(const ConvArgs& a_in, int sm_count, cudaStream_t stream) {
    ConvArgs a = a_in;
    if ((a.cin != 64 && a.cin != 128) || (a.cout != 64 && a.cout != 128)) {
        set_error("conv3x3: unsupported channels %d -> %d (need 64/128)", a.cin, a.cout);
        return -1;
    }
    if (a.n_img <= 0 || a.H <= 0 || a.W <= 0) {
        set_error("conv3x3: empty problem");
        return -1;
    }
    EncodeTiledFn encode = get_encode_fn();
    if (encode == nullptr) {
        set_error("conv3x3: cuTensorMapEncodeTiled not available from the driver");
        return -1;
    }
    const int nt = n_tile_for(a.cin);
    Geometry g;
    g.n_parts = a.cout / nt;
    g.x_tiles = (a.W + TILE_M - 1) / TILE_M;
    int ctas = std::max(g.n_parts, (sm_count / g.n_parts) * g.n_parts);
    const int groups = ctas / g.n_parts;
    if (a.strip_h <= 0) {
        // enough strips that every CTA group sees several work items, but strips no shorter than 8 rows
        const long long base_items = static_cast<long long>(a.n_img) * g.x_tiles;
        long long want = (6LL * groups + base_items - 1) / base_items;
        want = std::max(1LL, std::min<long long>(want, std::max(1, a.H / 8)));
        a.strip_h = static_cast<int>((a.H + want - 1) / want);
    }
    g.strips = (a.H + a.strip_h - 1) / a.strip_h;
    g.items_per_part = a.n_img * g.x_tiles * g.strips;
    ctas = std::min(ctas, g.items_per_part * g.n_parts);

    CUtensorMap map;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(a.in_c), static_cast<cuuint64_t>(a.W),
                                static_cast<cuuint64_t>(a.H), static_cast<cuuint64_t>(a.in_images)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(a.in_c) * 2, static_cast<cuuint64_t>(a.W) * a.in_c * 2,
                                   static_cast<cuuint64_t>(a.H) * a.W * a.in_c * 2};
    const cuuint32_t box[4] = {64, SLOT_PIX, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a.in), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("conv3x3: cuTensorMapEncodeTiled failed with CUresult %d (dims %d x %d x %d x %d)", (int)r, a.in_c,
                  a.W, a.H, a.in_images);
        return -1;
    }
    return a.cin == 64 ? launch_impl<64>(a, map, g, ctas, stream) : launch_impl<128>(a, map, g, ctas, stream);
}

}  // namespace hrn
