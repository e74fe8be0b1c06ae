// One encoder ResidualBlock (HRNet.py:17-33, 55-57) in ONE kernel:  out = x + PReLU2(conv2(PReLU1(conv1(x)))),  64 -> 64 -> 64.
//
// As two launches of conv3x3_umma the block moves 5.35 GB at C2 (x read, h written, h read, x re-read as the skip, out
// written) and its second launch is HBM-bound (3.2 GB in 0.64 ms, tensor pipe 50 % active).  Here the intermediate h
// never leaves the SM: two row-stationary conv pipelines run side by side in one CTA,
//     pipeline 1:  x rows (TMA ring, 2 slots)    -> accumulators in TMEM columns [0, 256)   -> epilogue 1 -> h rows in smem
//     pipeline 2:  h rows (smem ring, 2 slots)   -> accumulators in TMEM columns [256, 512) -> epilogue 2 -> out (global)
// with both weight sets resident (2 x 72 KB).  Epilogue 1 applies bias + PReLU, rounds to bf16 exactly like the
// stand-alone layer and writes the row in the K-major SWIZZLE_128B layout TMA would have produced (halo pixels and rows
// outside the image are zeros = conv2's padding), so results are bit-identical to the two-launch path.  The skip
// connection re-reads x rows that the TMA producer fetched a few rows earlier (L2 hits).  HBM traffic: 2.15 GB.
//
// Each pipeline has its own MMA-issuing thread (warps 1 and 3) running the same strip logic as conv3x3_umma; they only
// meet through mbarriers (h row full / empty).  A strip of R output rows needs h rows y0-1 .. y0+R, hence x rows
// y0-2 .. y0+R+1: the two extra h rows per strip are recomputed (about 2 % at C2).  Accumulator rings have 4 slots
// instead of 8, so every second row straddles the ring wrap and is issued as N = 128 + 64 instead of N = 192.
// One column tile only (W <= 128); wider images use the two-launch path.
#include "umma_common.cuh"
#include "strips.cuh"

#include <algorithm>

namespace hrn {
namespace {

constexpr int TILE_M = 128;
constexpr int SLOT_PIX = TILE_M + 2;
constexpr int CHUNK_BYTES = 17408;           // 130 px * 128 B rounded up to 1024 (keeps the SW128 phase)
constexpr int CHUNK_TX = SLOT_PIX * 128;
constexpr int NT = 64;
constexpr int ACCS = 4;                      // accumulator slots per pipeline
constexpr int RING = 2;                      // x rows / h rows resident per pipeline
constexpr int BTILE_BYTES = 3 * NT * 128;    // one kx tile: [W(ky=2) | W(ky=1) | W(ky=0)] x 64 ci
constexpr int W_BYTES = 3 * BTILE_BYTES;     // 73,728 per conv
constexpr int EPI2_WARPS = 8, EPI1_WARPS = 8;
constexpr int NUM_THREADS = 128 + EPI2_WARPS * 32 + EPI1_WARPS * 32;     // 640: at most 102 registers per thread

constexpr int X_OFFSET = 2 * W_BYTES;
constexpr int H_OFFSET = X_OFFSET + RING * CHUNK_BYTES;
constexpr int BAR_OFFSET = H_OFFSET + RING * CHUNK_BYTES;
constexpr int SMEM_BYTES = BAR_OFFSET + 512 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert((4 * RING + 4 * ACCS) * 8 + 16 <= 512, "barrier block");

struct RbArgs {
    ConvArgs c;                 // geometry, x (in), out, live list, conv 1 weights / PReLU
    const uint8_t* w2_img;      // conv 2: pre-swizzled weights, PReLU slope
    float prelu2;
    // Biases live in the kernel parameters (constant bank): the epilogues read them with LDC.  From shared memory the
    // loads queued behind the tensor cores' operand traffic (a fifth of the epilogue time), and 640 threads leave no
    // room to keep them in registers.
    float bias1[NT], bias2[NT];
};

// Barriers of one pipeline and where its operands / accumulators live.
struct Pipe {
    uint32_t ring_s, w_s, acc_base;                        // A ring, B image, first TMEM column
    uint32_t bar_full, bar_empty, bar_tfull, bar_tempty;   // [RING], [RING], [ACCS], [ACCS]
    int extra_rows;                                        // output rows beyond the strip (2 for pipeline 1)
};

// MMA issuer of one pipeline: ONE elected thread.  Same scheme as conv3x3_umma (CIN = 64, so one K chunk per row):
// interior rows take the lean path with early waits for the next row, strip-boundary rows the generic path.
__device__ __forceinline__ void mma_role(const Pipe& p, const Geometry& geo, const ConvArgs& a, int group, uint32_t bar_w) {
    constexpr uint32_t idesc_base = ptx::umma_idesc_bf16(TILE_M, 0);
    constexpr uint32_t idesc64 = ptx::umma_idesc_bf16(TILE_M, NT);
    constexpr uint32_t BLK = NT * 128 / 16;
    constexpr uint32_t B_KX = BTILE_BYTES / 16 - 6;        // k-step 3 of kx -> k-step 0 of kx + 1
    const uint32_t a_lo0 = desc_lo(p.ring_s), b_lo0 = desc_lo(p.w_s);
    ptx::mbar_wait(bar_w, 0, 2);
    uint32_t it = 0, tile0 = 0;
    bool full_seen = false, tempty_seen = false;
    StripWalker walk(geo, a, group, false);
    Strip s;
    bool have = walk.next(s);
    while (have) {
        Strip nxt;
        const bool have_next = walk.next(nxt);
        const int rows = s.rows + p.extra_rows;
        for (int q = 0; q < rows + 2; ++q, ++it) {
            const uint32_t slot = it % RING;
            if (q >= 2 && q <= rows - 1) {
                // ---- interior row: all three ky blocks live, block 2 opens the accumulator of output row q
                const uint32_t t_new = tile0 + q;
                const uint32_t sl = (t_new - 2) % ACCS;
                if (!tempty_seen) ptx::mbar_wait(p.bar_tempty + 8 * (t_new % ACCS), ((t_new / ACCS) & 1) ^ 1, 4);
                tempty_seen = false;
                const bool nxt_opens = (q + 1 <= rows - 1);
                if (!full_seen) ptx::mbar_wait(p.bar_full + 8 * slot, (it / RING) & 1, 3);
                full_seen = false;
                ptx::tc_fence_after();
                uint64_t ad = make_desc(a_lo0 + slot * (CHUNK_BYTES / 16));
                uint64_t bd = make_desc(b_lo0);
                const uint32_t dA = p.acc_base + sl * NT;
                // Probe (never block on) the barriers of the next row while the MMA queue is full: a blocking wait here
                // would hold back the last MMAs and the commits of THIS row until the other pipeline has produced the
                // next one, which chains the two pipelines' latencies together.
                auto early = [&]() {
                    const uint32_t itn = it + 1;
                    full_seen = ptx::mbar_test_wait(p.bar_full + 8 * (itn % RING), (itn / RING) & 1);
                    if (nxt_opens) {
                        const uint32_t tn = t_new + 1;
                        tempty_seen = ptx::mbar_test_wait(p.bar_tempty + 8 * (tn % ACCS), ((tn / ACCS) & 1) ^ 1);
                    }
                };
                if (sl <= ACCS - 3) {
                    ptx::umma_bf16(dA, ad, bd, idesc_base | ((2 * NT >> 3) << 17), 1u);
                    ptx::umma_bf16(dA + 2 * NT, ad, bd + 2 * BLK, idesc64, 0u);
#pragma unroll
                    for (int step = 1; step < 12; ++step) {
                        ad += 2;
                        bd += (step & 3) ? 2u : B_KX;
                        ptx::umma_bf16(dA, ad, bd, idesc_base | ((3 * NT >> 3) << 17), 1u);
                        if (step == 7) early();
                    }
                } else {
                    // wrap: sl = ACCS-2 -> blocks {0,1} at the last two slots, block 2 at slot 0;
                    //       sl = ACCS-1 -> block 0 at the last slot, blocks {1,2} at slots 0,1
                    const uint32_t n0 = (sl == ACCS - 2) ? 2u : 1u, n1 = 3u - n0;
                    const uint32_t id0 = idesc_base | ((n0 * NT >> 3) << 17), id1 = idesc_base | ((n1 * NT >> 3) << 17);
                    uint64_t bd1 = bd + n0 * BLK;
                    ptx::umma_bf16(dA, ad, bd, id0, 1u);
                    if (n1 == 2) ptx::umma_bf16(p.acc_base, ad, bd1, idesc64, 1u);
                    ptx::umma_bf16(p.acc_base + (n1 - 1) * NT, ad, bd + 2 * BLK, idesc64, 0u);
#pragma unroll
                    for (int step = 1; step < 12; ++step) {
                        ad += 2;
                        bd += (step & 3) ? 2u : B_KX;
                        bd1 += (step & 3) ? 2u : B_KX;
                        ptx::umma_bf16(dA, ad, bd, id0, 1u);
                        ptx::umma_bf16(p.acc_base, ad, bd1, id1, 1u);
                        if (step == 7) early();
                    }
                }
                ptx::umma_commit(p.bar_empty + 8 * slot);
                ptx::umma_commit(p.bar_tfull + 8 * sl);
                continue;
            }
            // ---- boundary rows of a strip: input row q feeds output rows q - ky; B block (2 - ky) <-> output row q - ky
            const int ky_lo = max(0, q - (rows - 1)), ky_hi = min(2, q);
            const int blk_lo = 2 - ky_hi, nblk = ky_hi - ky_lo + 1;
            const uint32_t t_lo = tile0 + q - ky_hi;
            const uint32_t s_lo = t_lo % ACCS;
            const bool opens = (ky_lo == 0);
            if (opens && !tempty_seen) {
                const uint32_t t_new = tile0 + q;
                ptx::mbar_wait(p.bar_tempty + 8 * (t_new % ACCS), ((t_new / ACCS) & 1) ^ 1, 4);
            }
            tempty_seen = false;
            const int w0 = min(nblk, ACCS - static_cast<int>(s_lo)), w1 = nblk - w0;
            const uint32_t d0 = p.acc_base + s_lo * NT, d1 = p.acc_base;
            const uint32_t id0 = idesc_base | (static_cast<uint32_t>(w0 * NT >> 3) << 17);
            const uint32_t id1 = idesc_base | (static_cast<uint32_t>(w1 * NT >> 3) << 17);
            const bool last_row = (q == rows + 1);
            const bool more_rows = !last_row || have_next;
            const bool next_opens = last_row ? true : (q + 1 <= rows - 1);
            const uint32_t t_next = last_row ? tile0 + rows : tile0 + q + 1;
            if (!full_seen) ptx::mbar_wait(p.bar_full + 8 * slot, (it / RING) & 1, 3);
            full_seen = false;
            ptx::tc_fence_after();
            uint64_t ad = make_desc(a_lo0 + slot * (CHUNK_BYTES / 16));
            uint64_t bd0 = make_desc(b_lo0 + blk_lo * BLK);
            uint64_t bd1 = bd0 + w0 * BLK;
            // first k-step block by block: the opening accumulator is overwritten while the older ones accumulate
            for (int b = 0; b < nblk; ++b)
                ptx::umma_bf16(p.acc_base + ((t_lo + b) % ACCS) * NT, ad, bd0 + b * BLK, idesc64,
                               (opens && b == nblk - 1) ? 0u : 1u);
#pragma unroll
            for (int step = 1; step < 12; ++step) {
                ad += 2;
                bd0 += (step & 3) ? 2u : B_KX;
                bd1 += (step & 3) ? 2u : B_KX;
                ptx::umma_bf16(d0, ad, bd0, id0, 1u);
                if (w1 > 0) ptx::umma_bf16(d1, ad, bd1, id1, 1u);
                if (step == 7 && more_rows) {
                    const uint32_t itn = it + 1;
                    full_seen = ptx::mbar_test_wait(p.bar_full + 8 * (itn % RING), (itn / RING) & 1);
                    if (next_opens) tempty_seen = ptx::mbar_test_wait(p.bar_tempty + 8 * (t_next % ACCS), ((t_next / ACCS) & 1) ^ 1);
                }
            }
            ptx::umma_commit(p.bar_empty + 8 * slot);
            if (ky_hi == 2) ptx::umma_commit(p.bar_tfull + 8 * ((tile0 + q - 2) % ACCS));
        }
        tile0 += rows;
        s = nxt;
        have = have_next;
    }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
resblock64_umma_kernel(const __grid_constant__ CUtensorMap in_map, const RbArgs r, const Geometry geo) {
    const ConvArgs& a = r.c;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    const uint32_t bars = base + BAR_OFFSET;
    Pipe p1, p2;
    p1.ring_s = base + X_OFFSET;
    p1.w_s = base;
    p1.bar_full = bars;
    p1.bar_empty = bars + 8 * RING;
    p2.ring_s = base + H_OFFSET;
    p2.w_s = base + W_BYTES;
    p2.bar_full = bars + 16 * RING;
    p2.bar_empty = bars + 24 * RING;
    p1.bar_tfull = bars + 32 * RING;
    p1.bar_tempty = p1.bar_tfull + 8 * ACCS;
    p2.bar_tfull = p1.bar_tempty + 8 * ACCS;
    p2.bar_tempty = p2.bar_tfull + 8 * ACCS;
    const uint32_t bar_w = p2.bar_tempty + 8 * ACCS;
    const uint32_t tmem_slot = bar_w + 8;
    p1.extra_rows = 2;
    p2.extra_rows = 0;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; ++i) {
            ptx::mbar_init(p1.bar_full + 8 * i, 1);              // TMA transaction
            ptx::mbar_init(p1.bar_empty + 8 * i, 1);             // tcgen05.commit
            ptx::mbar_init(p2.bar_full + 8 * i, EPI1_WARPS);     // one arrive per epilogue-1 warp
            ptx::mbar_init(p2.bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < ACCS; ++i) {
            ptx::mbar_init(p1.bar_tfull + 8 * i, 1);
            ptx::mbar_init(p1.bar_tempty + 8 * i, EPI1_WARPS);
            ptx::mbar_init(p2.bar_tfull + 8 * i, 1);
            ptx::mbar_init(p2.bar_tempty + 8 * i, EPI2_WARPS);
        }
        ptx::mbar_init(bar_w, 1);
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&in_map);
    }
    ptx::pdl_launch_dependents();
    if (warp == 2) ptx::tmem_alloc<512>(tmem_slot);
    // the h ring starts as zeros: halo pixels, pixels past W and nothing else are left untouched by epilogue 1
    for (int i = threadIdx.x; i < RING * CHUNK_BYTES / 16; i += NUM_THREADS)
        reinterpret_cast<uint4*>(smem_gen + H_OFFSET)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    p1.acc_base = tmem_base;
    p2.acc_base = tmem_base + ACCS * NT;

    if (warp == 0) {
        // ===================================================== TMA producer: weights of both convs, then x rows
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(bar_w, 2 * W_BYTES);
            for (int off = 0; off < W_BYTES; off += 8192) ptx::bulk_copy_g2s(p1.w_s + off, a.w_img + off, 8192, bar_w);
            for (int off = 0; off < W_BYTES; off += 8192) ptx::bulk_copy_g2s(p2.w_s + off, r.w2_img + off, 8192, bar_w);
            StripWalker walk(geo, a, group);
            Strip s;
            bool have = walk.next(s);
            ptx::pdl_wait();
            uint32_t it = 0;
            for (; have; have = walk.next(s)) {
                // With two slots the ring cannot hide an HBM round trip, so rows are pulled into L2 X_AHEAD rows ahead
                // of the load that brings them into shared memory.
                constexpr int X_AHEAD = 6;
                // x rows are read again as the skip connection a few rows later: keep them in L2 until then (evict_last);
                // debug flag 2048 = no eviction hints (A/B)
                const bool hint = !(a.debug_flags & 2048);
                const uint64_t keep = ptx::l2_policy_evict_last();
                for (int q = 0; q < X_AHEAD && q < s.rows + 4; ++q) {
                    if (hint) ptx::tma_prefetch_4d_hint(&in_map, 0, -1, s.y0 - 2 + q, s.m, keep);
                    else ptx::tma_prefetch_4d(&in_map, 0, -1, s.y0 - 2 + q, s.m);
                }
                for (int q = 0; q < s.rows + 4; ++q, ++it) {
                    const uint32_t slot = it % RING, ph = (it / RING) & 1;
                    if (q + X_AHEAD < s.rows + 4) {
                        if (hint) ptx::tma_prefetch_4d_hint(&in_map, 0, -1, s.y0 - 2 + q + X_AHEAD, s.m, keep);
                        else ptx::tma_prefetch_4d(&in_map, 0, -1, s.y0 - 2 + q + X_AHEAD, s.m);
                    }
                    ptx::mbar_wait(p1.bar_empty + 8 * slot, ph ^ 1, 1);
                    ptx::mbar_expect_tx(p1.bar_full + 8 * slot, CHUNK_TX);
                    if (hint) ptx::tma_load_4d_hint(p1.ring_s + slot * CHUNK_BYTES, &in_map, 0, -1, s.y0 - 2 + q, s.m, p1.bar_full + 8 * slot, keep);
                    else ptx::tma_load_4d(p1.ring_s + slot * CHUNK_BYTES, &in_map, 0, -1, s.y0 - 2 + q, s.m, p1.bar_full + 8 * slot);
                }
            }
        }
    } else if (warp == 1) {
        if (ptx::elect_one()) mma_role(p1, geo, a, group, bar_w);
    } else if (warp == 3) {
        if (ptx::elect_one()) mma_role(p2, geo, a, group, bar_w);
    } else if (warp >= 4 && warp < 4 + EPI2_WARPS) {
        // ===================================================== epilogue 2: out = x + PReLU2(acc + b2), bf16 NHWC global
        const int wq = warp & 3;
        const int hf = (warp - 4) >> 2;
        const int co0 = hf * 32;
        const float slope_m1 = r.prelu2 - 1.0f;
        const float* bias = r.bias2 + hf * 32;
        const bool skip_hint = !(a.debug_flags & 2048);
        StripWalker walk(geo, a, group);
        Strip s;
        bool have = walk.next(s);
        ptx::pdl_wait();
        uint32_t tile = 0;
        for (; have; have = walk.next(s)) {
            const int x = wq * 32 + lane;
            const bool valid = x < a.W;
            const size_t pix0 = static_cast<size_t>(s.y0) * a.W + x;
            const size_t img0 = static_cast<size_t>(s.m) * a.H * a.W;
            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(a.in) + (img0 + pix0) * NT + co0;
            __nv_bfloat16* op = a.out + (img0 + pix0) * NT + co0;
            const size_t step = static_cast<size_t>(a.W) * NT;
            for (int i = 0; i < s.rows; ++i, ++tile, rp += step, op += step) {
                const uint32_t acc = tile % ACCS, aph = (tile / ACCS) & 1;
                uint32_t rv[2][8];
                if (valid) {
                    if (skip_hint) {                                   // last use of this x row: free its L2 lines first
                        ptx::ldg_nc_v8_last_use(rp, rv[0]);
                        ptx::ldg_nc_v8_last_use(rp + 16, rv[1]);
                    } else {
                        ptx::ldg_nc_v8(rp, rv[0]);
                        ptx::ldg_nc_v8(rp + 16, rv[1]);
                    }
                }
                ptx::mbar_wait(p2.bar_tfull + 8 * acc, aph, 5);
                ptx::tc_fence_after();
                uint32_t v[32];
                ptx::tmem_ld_x32(p2.acc_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * NT + hf * 32, v);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(p2.bar_tempty + 8 * acc);
                const __nv_bfloat162 one2 = __floats2bfloat162_rn(1.0f, 1.0f);
#pragma unroll
                for (int g = 0; g < 2; ++g) {                          // 16 channels at a time
                    const float* bb = bias + 16 * g;
                    uint32_t o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float x0 = __uint_as_float(v[16 * g + 2 * e]) + bb[2 * e];
                        float x1 = __uint_as_float(v[16 * g + 2 * e + 1]) + bb[2 * e + 1];
                        x0 = fmaf(slope_m1, fminf(x0, 0.0f), x0);
                        x1 = fmaf(slope_m1, fminf(x1, 0.0f), x1);
                        __nv_bfloat162 y = __floats2bfloat162_rn(x0, x1);
                        if (valid) y = __hfma2(one2, y, *reinterpret_cast<const __nv_bfloat162*>(&rv[g][e]));
                        o[e] = *reinterpret_cast<const uint32_t*>(&y);
                    }
                    if (valid) {
                        if (skip_hint) ptx::stg_v8_stream(op + 16 * g, o);
                        else ptx::stg_v8(op + 16 * g, o);
                    }
                }
            }
        }
    } else if (warp >= 4 + EPI2_WARPS) {
        // ===================================================== epilogue 1: h = PReLU1(acc + b1) -> bf16 -> smem A row of conv 2
        const int wq = warp & 3;                       // TMEM lanes [32 wq, 32 wq + 32)
        const int hf = (warp - 4 - EPI2_WARPS) >> 2;   // channels [32 hf, 32 hf + 32) = 16-byte chunks 4 hf .. 4 hf + 3
        const int px = wq * 32 + lane;                 // pixel of the row; its smem row is px + 1 (slot 0 = left halo)
        const bool valid = px < a.W;
        const float slope_m1 = a.prelu - 1.0f;
        const float* bias = r.bias1 + hf * 32;
        const uint32_t row_off = static_cast<uint32_t>(px + 1) * 128u, sw = static_cast<uint32_t>((px + 1) & 7);
        StripWalker walk(geo, a, group, false);
        Strip s;
        uint32_t tile = 0;
        while (walk.next(s)) {
            for (int i = 0; i < s.rows + 2; ++i, ++tile) {
                const int ym = s.y0 - 1 + i;                           // image row of this h row
                const bool inside = ym >= 0 && ym < a.H;               // rows outside the image are conv 2's zero padding
                const uint32_t acc = tile % ACCS, aph = (tile / ACCS) & 1;
                const uint32_t hs = tile % RING, hph = (tile / RING) & 1;
                ptx::mbar_wait(p1.bar_tfull + 8 * acc, aph, 8);
                ptx::tc_fence_after();
                uint32_t v[32];
                ptx::tmem_ld_x32(p1.acc_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * NT + hf * 32, v);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(p1.bar_tempty + 8 * acc);
                ptx::mbar_wait(p2.bar_empty + 8 * hs, hph ^ 1, 9);     // conv 2 is done with the row that lived here
                if (valid) {
                    uint8_t* dst = smem_gen + H_OFFSET + hs * CHUNK_BYTES + row_off;
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const float* bb = bias + 16 * g;
#pragma unroll
                        for (int jj = 0; jj < 2; ++jj) {               // 16-byte chunk j = channels [8j, 8j + 8)
                            const uint32_t j = static_cast<uint32_t>(4 * hf + 2 * g + jj);
                            uint32_t o[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int c = 8 * jj + 2 * e;
                                float x0 = __uint_as_float(v[16 * g + c]) + bb[c];
                                float x1 = __uint_as_float(v[16 * g + c + 1]) + bb[c + 1];
                                x0 = fmaf(slope_m1, fminf(x0, 0.0f), x0);
                                x1 = fmaf(slope_m1, fminf(x1, 0.0f), x1);
                                const __nv_bfloat162 y = __floats2bfloat162_rn(x0, x1);
                                o[e] = inside ? *reinterpret_cast<const uint32_t*>(&y) : 0u;
                            }
                            *reinterpret_cast<uint4*>(dst + ((j ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(p2.bar_full + 8 * hs);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

}  // namespace

// One fused ResidualBlock(64): a1 describes conv 1 (x = a1.in, weights, PReLU) and the output tensor a1.out; w2 / prelu2
// are conv 2; the biases are HOST arrays of 64 floats.  Returns 1 if the shape is not supported (caller falls back to two launches).
int resblock64_launch(const ConvArgs& a1, const float* bias1_host, const uint8_t* w2_img, const float* bias2_host,
                      float prelu2, int sm_count, cudaStream_t stream) {
    if (a1.W > TILE_M || a1.cin != 64 || a1.cout != 64 || !a1.has_prelu) return 1;
    if (a1.n_img <= 0 || a1.H <= 0 || a1.W <= 0) {
        set_error("resblock64: empty problem");
        return -1;
    }
    RbArgs r;
    r.c = a1;
    r.w2_img = w2_img;
    r.prelu2 = prelu2;
    for (int i = 0; i < NT; ++i) {
        r.bias1[i] = bias1_host[i];
        r.bias2[i] = bias2_host[i];
    }
    Geometry g;
    g.n_parts = 1;
    g.x_tiles = 1;
    g.total_rows = static_cast<long long>(a1.n_img) * a1.H;
    const int ctas = a1.max_ctas > 0 ? std::min(a1.max_ctas, sm_count) : sm_count;
    g.groups = static_cast<int>(std::min<long long>(ctas, g.total_rows));
    g.split = a1.strip_split > 0 ? a1.strip_split : 1;
    g.img_group = 1;
    g.even_rows = 0;
    CUtensorMap map;
    if (encode_nhwc_map(&map, a1.in, 64, a1.W, a1.H, a1.in_images, SLOT_PIX)) return -1;
    static bool attr_set[64] = {};
    if (allow_dynamic_smem(resblock64_umma_kernel, SMEM_BYTES, attr_set)) return -1;
    HRN_CUDA_OK(launch_pdl(resblock64_umma_kernel, g.groups, NUM_THREADS, SMEM_BYTES, stream, 1, map, r, g));
    note_launches(1);
    return 0;
}

}  // namespace hrn
