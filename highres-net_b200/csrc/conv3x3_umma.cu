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
// Mapping ("row-stationary, ky-stacked").  A = 128 consecutive pixels of ONE input
// row (M = 128, K = 64 channels per chunk), fetched once by TMA as 130 pixels
// (halo; out-of-bounds pixels are zero-filled by TMA = the conv padding).  The
// kx = 0,1,2 taps read the same smem buffer through descriptors whose start
// address is shifted by kx pixels (kx * 128 B; SWIZZLE_128B is address based, so
// the shifted view stays consistent with what TMA wrote).  Input row r feeds
// output rows r+1, r, r-1 through taps ky = 0, 1, 2, so B stacks
// [W(ky=2) | W(ky=1) | W(ky=0)] (3 x 64 output channels = N 192) and ONE MMA
// updates the three accumulators of output rows r-1, r, r+1, which live in
// adjacent 64-column TMEM slots.  Versus one MMA per (tap, output row) this
// issues a third of the MMAs and reads a third of the A bytes from shared memory.
// The CTA's weight slice (9 x CIN x 64 bf16) stays resident in smem; every input
// row is consumed exactly once, so the A ring is a plain FIFO.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one
// elected thread), warp 2 = TMEM allocator, warps 4..11 = epilogue (TMEM ->
// registers -> bias/PReLU/residual -> bf16 NHWC global).  TMEM holds 8
// accumulator slots (all 512 columns), so epilogues overlap the MMAs of later rows.
// Work is split evenly: each CTA (group) owns a contiguous range of the flattened
// (image, column tile, row) space and walks it as per-image strips.
#include "umma_common.cuh"
#include "strips.cuh"

#include <algorithm>
#include <cstring>

namespace hrn {
namespace {

constexpr int TILE_M = 128;
constexpr int SLOT_PIX = TILE_M + 2;
constexpr int CHUNK_BYTES = 17408;     // 130 px * 128 B = 16640, rounded up to 1024 (keeps the SW128 phase)
constexpr int CHUNK_TX = SLOT_PIX * 128;
constexpr int NT = 64;                 // output channels per CTA
constexpr int ACC_SLOTS = 8;           // 8 x 64 fp32 columns = the whole TMEM
constexpr int TMEM_COLS = ACC_SLOTS * NT;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + EPI_WARPS * 32;
constexpr int BTILE_BYTES = 3 * NT * 128;   // one (kx, chunk) B tile: 192 rows x 64 bf16

template <int CIN>
struct Cfg {
    static constexpr int CHUNKS = CIN / 64;
    static constexpr int RING = (CIN == 64) ? 8 : 4;          // resident (input row, chunk) buffers
    static constexpr int W_BYTES = 3 * CHUNKS * BTILE_BYTES;  // 73,728 or 147,456
    static constexpr int BAR_OFFSET = W_BYTES + RING * CHUNK_BYTES;
    static constexpr int BIAS_OFFSET = BAR_OFFSET + 512;
    static constexpr int SMEM_BYTES = BIAS_OFFSET + NT * 4 + 1024;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
    static_assert(16 * RING + 16 * ACC_SLOTS + 8 + 8 <= 512, "barrier block overflows into the bias array");
};

// MCAST (128 -> 128 layers): the two CTAs that compute the two 64-channel output halves of the same rows form a cluster;
// each loads one of the two K chunks of every input row and multicasts it to both, so every A row crosses L2 -> SM once
// instead of twice.  A slot may be overwritten only when BOTH CTAs have consumed it: the MMA threads multicast their
// "slot free" commits to both CTAs and the empty barriers count two arrivals.
template <int CIN, bool POOL, bool MCAST>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap in_map, const ConvArgs a, const Geometry geo) {
    using C = Cfg<CIN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_s = base;
    const uint32_t ring_s = base + C::W_BYTES;
    const uint32_t bars = base + C::BAR_OFFSET;
    const uint32_t bar_full = bars;                              // [RING]
    const uint32_t bar_empty = bars + 8 * C::RING;               // [RING]
    const uint32_t bar_tfull = bars + 16 * C::RING;              // [ACC_SLOTS]
    const uint32_t bar_tempty = bar_tfull + 8 * ACC_SLOTS;       // [ACC_SLOTS]
    const uint32_t bar_w = bar_tempty + 8 * ACC_SLOTS;
    const uint32_t tmem_slot = bar_w + 8;
    uint8_t* smem_gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));
    float* bias_s = reinterpret_cast<float*>(smem_gen + C::BIAS_OFFSET);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int part = blockIdx.x % geo.n_parts;
    const int group = blockIdx.x / geo.n_parts;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::RING; ++i) {
            ptx::mbar_init(bar_full + 8 * i, 1);
            ptx::mbar_init(bar_empty + 8 * i, MCAST ? 2 : 1);
        }
        for (int i = 0; i < ACC_SLOTS; ++i) {
            ptx::mbar_init(bar_tfull + 8 * i, 1);
            ptx::mbar_init(bar_tempty + 8 * i, EPI_WARPS);   // one arrive per epilogue warp
        }
        ptx::mbar_init(bar_w, 1);
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&in_map);
    }
    ptx::pdl_launch_dependents();            // the next kernel's prologue may overlap our tail
    if (warp == 2) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
    if (threadIdx.x >= 128 && threadIdx.x < 128 + NT) bias_s[threadIdx.x - 128] = a.bias[part * NT + threadIdx.x - 128];
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (MCAST) ptx::cluster_sync();          // the peer's barriers exist before anything is multicast into this CTA
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        // ===================================================== TMA producer: one elected thread
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(bar_w, C::W_BYTES);
            const uint8_t* wsrc = a.w_img + static_cast<size_t>(part) * C::W_BYTES;
            for (int off = 0; off < C::W_BYTES; off += 8192) ptx::bulk_copy_g2s(w_s + off, wsrc + off, 8192, bar_w);
            // The live-work list was finished long before the previous kernel started (live_lists_launch), so the
            // first strip is resolved while that kernel may still be running.
            StripWalker walk(geo, a, group);
            Strip s;
            bool have = walk.next(s);
            ptx::pdl_wait();                 // weights are constants; the activations come from the previous kernel
            uint32_t it = 0;
            // several narrow images per tile: one box = (W + 2 pixels from x = -1) x G images, segment after segment
            const int G = geo.img_group;
            const uint32_t chunk_tx = G > 1 ? static_cast<uint32_t>(G * (a.W + 2) * 128) : static_cast<uint32_t>(CHUNK_TX);
            for (; have; have = walk.next(s)) {
                int img[2], ch[2];
                if (a.pair_mode) {
                    const int b = s.m / a.half, i = s.m % a.half;
                    img[0] = b * a.src_views + i;
                    img[1] = b * a.src_views + (a.top - 1 - i);
                    ch[0] = ch[1] = 0;
                } else {
                    img[0] = img[1] = s.m * G;
                    ch[0] = 0;
                    ch[1] = 64;
                }
                for (int q = 0; q < s.rows + 2; ++q) {
#pragma unroll
                    for (int c = 0; c < C::CHUNKS; ++c, ++it) {
                        const uint32_t slot = it % C::RING, ph = (it / C::RING) & 1;
                        ptx::mbar_wait(bar_empty + 8 * slot, ph ^ 1, 1);
                        if (a.debug_flags & 4) {
                            ptx::mbar_arrive(bar_full + 8 * slot);
                        } else {
                            ptx::mbar_expect_tx(bar_full + 8 * slot, chunk_tx);
                            if (!MCAST)
                                ptx::tma_load_4d(ring_s + slot * CHUNK_BYTES, &in_map, ch[c], s.xt * TILE_M - 1,
                                                 s.y0 - 1 + q, img[c], bar_full + 8 * slot);
                            else if (c == part)          // chunk c is fetched by cluster rank c for both CTAs
                                ptx::tma_load_4d_mcast(ring_s + slot * CHUNK_BYTES, &in_map, ch[c], s.xt * TILE_M - 1,
                                                       s.y0 - 1 + q, img[c], bar_full + 8 * slot, 0x3);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer: ONE elected thread runs the whole role.
        // The tensor pipe queues about six N=192 MMAs (~600 cycles, measured with tools/umma_probe.cu), so the
        // per-item bookkeeping is hidden as long as it stays short: the barrier waits for the NEXT (row, chunk)
        // item are therefore issued in the middle of the current item's MMA burst, when the queue is full anyway.
        if (ptx::elect_one()) {
            constexpr uint32_t idesc_base = ptx::umma_idesc_bf16(TILE_M, 0);
            constexpr uint32_t idesc64 = ptx::umma_idesc_bf16(TILE_M, NT);
            constexpr uint32_t BLK = NT * 128 / 16;               // one 64-row ky block, in descriptor units (16 B)
            constexpr uint32_t B_KX = C::CHUNKS * BTILE_BYTES / 16 - 6;   // k-step 3 of kx -> k-step 0 of kx + 1
            const uint32_t a_lo0 = desc_lo(ring_s), b_lo0 = desc_lo(w_s);
            ptx::mbar_wait(bar_w, 0, 2);
            uint32_t it = 0, tile0 = 0;
            bool full_seen = false, tempty_seen = false;          // waits already done by the previous item
            StripWalker walk(geo, a, group, false);
            Strip s;
            bool have = walk.next(s);
            while (have) {
                Strip nxt;
                const bool have_next = walk.next(nxt);
                for (int q = 0; q < s.rows + 2; ++q) {
                    if (q >= 2 && q <= s.rows - 1) {
                        // ---- interior row (the bulk of the work): all three ky blocks are live and block 2 opens the
                        // accumulator of output row q.  Everything is derived from the slot of the oldest tile.
                        const uint32_t t_new = tile0 + q;
                        const uint32_t sl = (t_new - 2) % ACC_SLOTS;             // slot of output row q - 2
                        if (!tempty_seen)
                            ptx::mbar_wait(bar_tempty + 8 * (t_new % ACC_SLOTS), ((t_new / ACC_SLOTS) & 1) ^ 1, 4);
                        tempty_seen = false;
                        const bool nxt_opens = (q + 1 <= s.rows - 1);
#pragma unroll
                        for (int c = 0; c < C::CHUNKS; ++c, ++it) {
                            const uint32_t slot = it % C::RING;
                            if (!full_seen) ptx::mbar_wait(bar_full + 8 * slot, (it / C::RING) & 1, 3);
                            full_seen = false;
                            ptx::tc_fence_after();
                            uint64_t ad = make_desc(a_lo0 + slot * (CHUNK_BYTES / 16));
                            uint64_t bd = make_desc(b_lo0 + c * (BTILE_BYTES / 16));
                            const uint32_t dA = tmem_base + sl * NT;
                            // early waits for the next item, issued after k-step 7 while the MMA queue is full
                            auto early = [&]() {
                                const uint32_t itn = it + 1;
                                ptx::mbar_wait(bar_full + 8 * (itn % C::RING), (itn / C::RING) & 1, 6);
                                full_seen = true;
                                if (c + 1 == C::CHUNKS && nxt_opens) {
                                    const uint32_t tn = t_new + 1;
                                    ptx::mbar_wait(bar_tempty + 8 * (tn % ACC_SLOTS), ((tn / ACC_SLOTS) & 1) ^ 1, 7);
                                    tempty_seen = true;
                                }
                            };
                            if (sl <= ACC_SLOTS - 3) {
                                // slots sl, sl+1, sl+2 are contiguous: one N = 192 MMA per k-step
                                if (c == 0) {
                                    ptx::umma_bf16(dA, ad, bd, idesc_base | ((2 * NT >> 3) << 17), 1u);
                                    ptx::umma_bf16(dA + 2 * NT, ad, bd + 2 * BLK, idesc64, 0u);
                                } else {
                                    ptx::umma_bf16(dA, ad, bd, idesc_base | ((3 * NT >> 3) << 17), 1u);
                                }
#pragma unroll
                                for (int step = 1; step < 12; ++step) {
                                    ad += 2;
                                    bd += (step & 3) ? 2u : B_KX;
                                    ptx::umma_bf16(dA, ad, bd, idesc_base | ((3 * NT >> 3) << 17), 1u);
                                    if (step == 7) early();
                                }
                            } else {
                                // wrap: sl = 6 -> blocks {0,1} at slots 6,7 and block 2 at slot 0;
                                //       sl = 7 -> block 0 at slot 7 and blocks {1,2} at slots 0,1
                                const uint32_t n0 = (sl == ACC_SLOTS - 2) ? 2u : 1u, n1 = 3u - n0;
                                const uint32_t id0 = idesc_base | ((n0 * NT >> 3) << 17), id1 = idesc_base | ((n1 * NT >> 3) << 17);
                                uint64_t bd1 = bd + n0 * BLK;
                                if (c == 0) {
                                    ptx::umma_bf16(dA, ad, bd, id0, 1u);
                                    if (n1 == 2) ptx::umma_bf16(tmem_base, ad, bd1, idesc64, 1u);
                                    ptx::umma_bf16(tmem_base + (n1 - 1) * NT, ad, bd + 2 * BLK, idesc64, 0u);
                                } else {
                                    ptx::umma_bf16(dA, ad, bd, id0, 1u);
                                    ptx::umma_bf16(tmem_base, ad, bd1, id1, 1u);
                                }
#pragma unroll
                                for (int step = 1; step < 12; ++step) {
                                    ad += 2;
                                    bd += (step & 3) ? 2u : B_KX;
                                    bd1 += (step & 3) ? 2u : B_KX;
                                    ptx::umma_bf16(dA, ad, bd, id0, 1u);
                                    ptx::umma_bf16(tmem_base, ad, bd1, id1, 1u);
                                    if (step == 7) early();
                                }
                            }
                            if (MCAST) ptx::umma_commit_mcast(bar_empty + 8 * slot, 0x3); else ptx::umma_commit(bar_empty + 8 * slot);
                            if (c == C::CHUNKS - 1) ptx::umma_commit(bar_tfull + 8 * sl);
                        }
                        continue;
                    }
                    // ---- boundary rows of a strip (q = 0, 1, rows, rows + 1, or very short strips): generic path
                    // Input row q feeds output rows o = q - ky.  B block (2 - ky) <-> output row q - ky, so the blocks
                    // [blk_lo, blk_lo + nblk) map to the consecutive accumulators (tiles) t_lo, t_lo + 1, ...
                    const int ky_lo = max(0, q - (s.rows - 1)), ky_hi = min(2, q);
                    const int blk_lo = 2 - ky_hi, nblk = ky_hi - ky_lo + 1;
                    const uint32_t t_lo = tile0 + q - ky_hi;
                    const uint32_t s_lo = t_lo % ACC_SLOTS;
                    const bool opens = (ky_lo == 0);                    // the last block starts a new accumulator
                    if (opens && !tempty_seen && !(a.debug_flags & 16)) {
                        const uint32_t t_new = tile0 + q;
                        ptx::mbar_wait(bar_tempty + 8 * (t_new % ACC_SLOTS), ((t_new / ACC_SLOTS) & 1) ^ 1, 4);
                    }
                    tempty_seen = false;
                    // the accumulators are contiguous in TMEM except across the slot 7 -> 0 wrap: at most two segments
                    const int w0 = min(nblk, ACC_SLOTS - static_cast<int>(s_lo)), w1 = nblk - w0;
                    const uint32_t d0 = tmem_base + s_lo * NT, d1 = tmem_base;
                    const uint32_t id0 = idesc_base | (static_cast<uint32_t>(w0 * NT >> 3) << 17);
                    const uint32_t id1 = idesc_base | (static_cast<uint32_t>(w1 * NT >> 3) << 17);
                    // what follows this row (for the early waits)
                    const bool last_row = (q == s.rows + 1);
                    const bool more_rows = !last_row || have_next;
                    const bool next_opens = last_row ? true : (q + 1 <= s.rows - 1);
                    const uint32_t t_next = last_row ? tile0 + s.rows : tile0 + q + 1;
#pragma unroll
                    for (int c = 0; c < C::CHUNKS; ++c, ++it) {
                        const uint32_t slot = it % C::RING;
                        if (!full_seen && !(a.debug_flags & 16)) ptx::mbar_wait(bar_full + 8 * slot, (it / C::RING) & 1, 3);
                        full_seen = false;
                        ptx::tc_fence_after();
                        // Running 64-bit descriptors advanced in place: +2 (32 B) per k-step, +8 (one pixel) per kx
                        // for A; next (kx, chunk) tile for B.
                        uint64_t ad = make_desc(a_lo0 + slot * (CHUNK_BYTES / 16));
                        uint64_t bd0 = make_desc(b_lo0 + c * (BTILE_BYTES / 16) + blk_lo * BLK);
                        uint64_t bd1 = bd0 + w0 * BLK;
                        if (c == 0) {
                            // first k-step of the row: block by block, so that the opening accumulator is
                            // overwritten (accumulate = 0) while the older ones keep accumulating
                            for (int b = 0; b < nblk; ++b)
                                ptx::umma_bf16(tmem_base + ((t_lo + b) % ACC_SLOTS) * NT, ad, bd0 + b * BLK, idesc64,
                                               (opens && b == nblk - 1) ? 0u : 1u);
                        } else {
                            ptx::umma_bf16(d0, ad, bd0, id0, 1u);
                            if (w1 > 0) ptx::umma_bf16(d1, ad, bd1, id1, 1u);
                        }
#pragma unroll
                        for (int step = 1; step < 12; ++step) {
                            ad += 2;
                            bd0 += (step & 3) ? 2u : B_KX;
                            bd1 += (step & 3) ? 2u : B_KX;
                            ptx::umma_bf16(d0, ad, bd0, id0, 1u);
                            if (w1 > 0) ptx::umma_bf16(d1, ad, bd1, id1, 1u);
                            if (step == 7) {
                                // early waits for the next item: its A buffer, and its new accumulator if it opens one
                                const bool next_item = (c + 1 < C::CHUNKS) || more_rows;
                                if (next_item && !(a.debug_flags & 16)) {
                                    const uint32_t itn = it + 1;
                                    ptx::mbar_wait(bar_full + 8 * (itn % C::RING), (itn / C::RING) & 1, 6);
                                    full_seen = true;
                                    if (c + 1 == C::CHUNKS && next_opens) {
                                        ptx::mbar_wait(bar_tempty + 8 * (t_next % ACC_SLOTS),
                                                       ((t_next / ACC_SLOTS) & 1) ^ 1, 7);
                                        tempty_seen = true;
                                    }
                                }
                            }
                        }
                        if (MCAST) ptx::umma_commit_mcast(bar_empty + 8 * slot, 0x3); else ptx::umma_commit(bar_empty + 8 * slot);                 // this (row, chunk) buffer is consumed
                        if (c == C::CHUNKS - 1 && ky_hi == 2)                   // output row q-2 has all 9 taps
                            ptx::umma_commit(bar_tfull + 8 * ((tile0 + q - 2) % ACC_SLOTS));
                    }
                }
                tile0 += s.rows;
                s = nxt;
                have = have_next;
            }
        }
    } else if (warp >= 4) {
        // ===================================================== epilogue: 8 warps, (lane quadrant) x (column half)
        // Per thread and output row: one pixel x 32 channels.  Bias and PReLU in fp32 on the accumulators, one
        // rounding to bf16, then the residual / alpha merge as a packed bf16x2 FMA; 256-bit global loads/stores.
        const int wq = warp & 3;                 // TMEM lanes [32 wq, 32 wq + 32)
        const int hf = (warp - 4) >> 2;          // accumulator columns [32 hf, 32 hf + 32)
        const int co0 = part * NT + hf * 32;     // first output channel handled by this thread
        float bias_r[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) bias_r[e] = bias_s[hf * 32 + e];
        const bool has_prelu = a.has_prelu != 0;
        const float slope_m1 = a.prelu - 1.0f;   // PReLU(v) = v + (slope - 1) * min(v, 0)
        StripWalker walk(geo, a, group);
        Strip s;
        bool have = walk.next(s);
        ptx::pdl_wait();                         // residual reads and output writes touch the previous kernel's tensors
        uint32_t tile = 0;
        for (; have; have = walk.next(s)) {
            int x = s.xt * TILE_M + wq * 32 + lane, m_img = s.m;
            bool valid = x < a.W;
            if (geo.img_group > 1) {          // tile row = segment g (image s.m * G + g), pixel x of that image's row
                const int seg = a.W + 2, g = x / seg;
                x -= g * seg;
                m_img = s.m * geo.img_group + g;
                valid = g < geo.img_group && x < a.W && m_img < a.n_img;
            }
            const __nv_bfloat16* res_img = nullptr;
            int res_c = 0;
            float scale = 1.0f;
            if (a.res_mode == RES_SAME) {
                res_img = a.res + (static_cast<size_t>(s.m) * a.H * a.W) * a.cout + co0;
                res_c = a.cout;
            } else if (a.res_mode == RES_PAIR) {
                const int b = s.m / a.half, i = s.m % a.half;
                const int side = co0 >= 64;
                const int img = b * a.src_views + (side ? (a.top - 1 - i) : i);
                res_img = a.res + (static_cast<size_t>(img) * a.H * a.W) * 64 + (co0 - 64 * side);
                res_c = 64;
            } else if (a.res_mode == RES_ALPHA) {
                const int b = s.m / a.half, i = s.m % a.half;
                res_img = a.res + (static_cast<size_t>(b * a.src_views + i) * a.H * a.W) * 64 + co0;
                res_c = 64;
                scale = a.alphas[b * a.alpha_stride + (a.top - 1 - i)];
            }
            const bool use_res = res_img != nullptr && valid && !(a.debug_flags & 8);
            const __nv_bfloat162 scale2 = __floats2bfloat162_rn(scale, scale);
            const size_t pix0 = static_cast<size_t>(s.y0) * a.W + x;
            const __nv_bfloat16* rp = res_img + pix0 * res_c;
            const int out_img = a.out_in_stack ? (s.m / a.half) * a.src_views + s.m % a.half : m_img;
            __nv_bfloat16* op = a.out + (static_cast<size_t>(out_img) * a.H * a.W + pix0) * a.cout + co0;
            const size_t r_step = static_cast<size_t>(a.W) * res_c, o_step = static_cast<size_t>(a.W) * a.cout;
            // fused MaxPool2d(2): rows come in pairs (2k, 2k + 1) because every range starts on an even row; the pooled
            // pixel (y / 2, x / 2) is written by the even-x lane of the odd row
            __nv_bfloat16* pp = nullptr;
            uint32_t prev[16];
            if (POOL)
                pp = a.out + ((static_cast<size_t>(out_img) * (a.H / 2) + s.y0 / 2) * (a.W / 2) + x / 2) * a.cout + co0;
            for (int i = 0; i < s.rows; ++i, ++tile, rp += r_step, op += o_step) {
                const uint32_t acc = tile % ACC_SLOTS, aph = (tile / ACC_SLOTS) & 1;
                uint32_t rv[2][8];
                if (use_res) {
                    ptx::ldg_nc_v8(rp, rv[0]);
                    ptx::ldg_nc_v8(rp + 16, rv[1]);
                }
                ptx::mbar_wait(bar_tfull + 8 * acc, aph, 5);
                ptx::tc_fence_after();
                uint32_t v[32];
                if (a.debug_flags & 1) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = 0;
                } else {
                    ptx::tmem_ld_x32(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * NT + hf * 32, v);
                    ptx::tmem_ld_wait();
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * acc);
                uint32_t o[2][8];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float x0 = __uint_as_float(v[2 * e]) + bias_r[2 * e];
                    float x1 = __uint_as_float(v[2 * e + 1]) + bias_r[2 * e + 1];
                    if (has_prelu) {
                        x0 = fmaf(slope_m1, fminf(x0, 0.0f), x0);
                        x1 = fmaf(slope_m1, fminf(x1, 0.0f), x1);
                    }
                    __nv_bfloat162 y = __floats2bfloat162_rn(x0, x1);
                    if (use_res) y = __hfma2(scale2, y, *reinterpret_cast<const __nv_bfloat162*>(&rv[e >> 3][e & 7]));
                    o[e >> 3][e & 7] = *reinterpret_cast<const uint32_t*>(&y);
                }
                if (POOL) {
                    if (((s.y0 + i) & 1) == 0) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) prev[e] = o[e >> 3][e & 7];
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&prev[e]),
                                                       *reinterpret_cast<const __nv_bfloat162*>(&o[e >> 3][e & 7]));
                            uint32_t mu = *reinterpret_cast<const uint32_t*>(&m);
                            const uint32_t nu = __shfl_xor_sync(0xffffffffu, mu, 1);       // pixel x ^ 1 of the same rows
                            m = __hmax2(m, *reinterpret_cast<const __nv_bfloat162*>(&nu));
                            o[e >> 3][e & 7] = *reinterpret_cast<const uint32_t*>(&m);
                        }
                        if (valid && (x & 1) == 0) {
                            ptx::stg_v8(pp, o[0]);
                            ptx::stg_v8(pp + 16, o[1]);
                        }
                        pp += static_cast<size_t>(a.W / 2) * a.cout;
                    }
                } else if (valid && !(a.debug_flags & 2)) {
                    ptx::stg_v8(op, o[0]);
                    ptx::stg_v8(op + 16, o[1]);
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (MCAST) ptx::cluster_sync();          // no CTA leaves while its peer may still multicast into it
    if (warp == 2) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------- host side
template <int CIN, bool POOL, bool MCAST = false>
int launch_impl(const ConvArgs& a, const CUtensorMap& map, const Geometry& g, int ctas, cudaStream_t stream) {
    using C = Cfg<CIN>;
    static bool attr_set[64] = {};
    if (allow_dynamic_smem(conv3x3_umma_kernel<CIN, POOL, MCAST>, C::SMEM_BYTES, attr_set)) return -1;
    HRN_CUDA_OK(launch_pdl(conv3x3_umma_kernel<CIN, POOL, MCAST>, ctas, NUM_THREADS, C::SMEM_BYTES, stream, MCAST ? 2 : 1, map, a, g));
    note_launches(1);
    return 0;
}

}  // namespace

// How many two-CTA clusters of the multicast kernel can be resident on this device (queried once per device).
static bool cluster_pairs_fit(int pairs) {
    static int max_pairs[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    using C = Cfg<128>;
    static bool attr_set[64] = {};
    if (allow_dynamic_smem(conv3x3_umma_kernel<128, false, true>, C::SMEM_BYTES, attr_set)) return false;
    std::lock_guard<std::mutex> lock(lazy_init_mutex());
    if (max_pairs[dev] == 0) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * 148);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = C::SMEM_BYTES;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, conv3x3_umma_kernel<128, false, true>, &cfg) != cudaSuccess) {
            cudaGetLastError();
            n = -1;
        }
        max_pairs[dev] = n > 0 ? n : -1;
    }
    return max_pairs[dev] >= pairs;
}

int conv3x3_bytes_per_weight_image(int cin, int cout) { return 9 * cin * cout * 2; }

// OIHW fp32 -> per 64-channel output part: [kx][chunk][block = 2 - ky][n = 64 co][k = 64 ci] bf16, each
// (kx, chunk) tile being 192 K-major rows of 128 B whose 16-byte chunks are XOR-swizzled by (row % 8).
void conv3x3_pack_weights(const float* oihw, int cin, int cout, uint8_t* dst) {
    const int chunks = cin / 64, parts = cout / NT;
    for (int p = 0; p < parts; ++p)
        for (int kx = 0; kx < 3; ++kx)
            for (int c = 0; c < chunks; ++c) {
                uint8_t* tile = dst + ((static_cast<size_t>(p) * 3 + kx) * chunks + c) * BTILE_BYTES;
                for (int blk = 0; blk < 3; ++blk) {
                    const int ky = 2 - blk;
                    for (int n = 0; n < NT; ++n)
                        for (int k = 0; k < 64; ++k) {
                            const int co = p * NT + n, ci = c * 64 + k, row = blk * NT + n;
                            const float w = oihw[(static_cast<size_t>(co) * cin + ci) * 9 + ky * 3 + kx];
                            const __nv_bfloat16 h = __float2bfloat16_rn(w);
                            const size_t off = static_cast<size_t>(row) * 128 + (((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2;
                            std::memcpy(tile + off, &h, 2);
                        }
                }
            }
}

int conv3x3_launch(const ConvArgs& a, int sm_count, cudaStream_t stream) {
    if ((a.cin != 64 && a.cin != 128) || (a.cout != 64 && a.cout != 128)) {
        set_error("conv3x3: unsupported channels %d -> %d (need 64/128)", a.cin, a.cout);
        return -1;
    }
    if (a.n_img <= 0 || a.H <= 0 || a.W <= 0) {
        set_error("conv3x3: empty problem");
        return -1;
    }
    if (a.pool && ((a.H | a.W) & 1 || a.live_list != nullptr || a.pair_mode || a.res_mode != RES_NONE || a.out_in_stack)) {
        set_error("conv3x3: the fused max pool needs even H and W and a plain layer (no list, pair gather, residual, in-stack output)");
        return -1;
    }
    Geometry g;
    g.even_rows = a.pool ? 1 : 0;
    g.n_parts = a.cout / NT;
    g.x_tiles = (a.W + TILE_M - 1) / TILE_M;
    // Narrow images (W + 2 <= 64): an M tile of 128 pixels would be mostly empty, so G = 128 / (W + 2) images share one
    // tile, each with its own zero halo column on both sides (3 images at W = 32, 7 at W = 16).  Every output pixel
    // still sees exactly the same products in the same order, so results are bit-identical to G = 1.  Plain layers only
    // (no live-work list, pair gather, residual or in-stack output): that is what the ShiftNet path needs.
    g.img_group = 1;
    if (a.W + 2 <= TILE_M / 2 && a.n_img > 1 && !a.no_img_group && a.live_list == nullptr && !a.pair_mode &&
        a.res_mode == RES_NONE && !a.out_in_stack)
        g.img_group = TILE_M / (a.W + 2);
    const int n_groups = (a.n_img + g.img_group - 1) / g.img_group;
    g.total_rows = static_cast<long long>(n_groups) * g.x_tiles * a.H;
    int ctas = a.max_ctas > 0 ? std::min(a.max_ctas, sm_count) : sm_count;
    ctas = std::max(g.n_parts, (ctas / g.n_parts) * g.n_parts);
    g.groups = static_cast<int>(std::min<long long>(ctas / g.n_parts, g.total_rows));
    ctas = g.groups * g.n_parts;
    g.split = a.strip_split > 0 ? a.strip_split : 1;

    CUtensorMap map;
    if (encode_nhwc_map(&map, a.in, a.in_c, a.W, a.H, a.in_images, g.img_group > 1 ? a.W + 2 : SLOT_PIX, 1, g.img_group)) return -1;
    if (a.pool) return a.cin == 64 ? launch_impl<64, true>(a, map, g, ctas, stream) : launch_impl<128, true>(a, map, g, ctas, stream);
    // 128 -> 128: cluster pairs with multicast A rows (knob: ConvArgs::mcast, set by the handle) -- provided every pair
    // can be resident at once (one CTA per SM, a pair per TPC); otherwise a second wave would cost far more than the
    // multicast saves, and the plain launch is used.
    if (a.cin == 128 && a.cout == 128 && a.mcast && !(a.debug_flags & 4)) {
        if (cluster_pairs_fit(ctas / 2)) return launch_impl<128, false, true>(a, map, g, ctas, stream);
        if (a.mcast == 2) {              // test knob: the cluster path is REQUIRED
            set_error("conv3x3: %d cluster pairs cannot be resident at once on this device", ctas / 2);
            return -1;
        }
    }
    return a.cin == 64 ? launch_impl<64, false>(a, map, g, ctas, stream) : launch_impl<128, false>(a, map, g, ctas, stream);
}

}  // namespace hrn
