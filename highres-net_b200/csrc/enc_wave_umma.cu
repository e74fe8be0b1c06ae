// The encoder's five 64 -> 64 convolutions (HRNet.py:55-60: two ResidualBlocks, HRNet.py:17-33, and the final conv) in ONE
// launch, as a ROW WAVEFRONT through the SMs -- the scheme of fuse_wave_umma.cu applied to the encoder:
//
//     stream = five CTAs:   R0a  conv 1 of ResidualBlock 0 + PReLU            (input: the first conv's output x0, from HBM)
//                           R0b  conv 2 of ResidualBlock 0 + PReLU + skip x0  -> x1
//                           R1a  conv 1 of ResidualBlock 1 + PReLU
//                           R1b  conv 2 of ResidualBlock 1 + PReLU + skip x1  -> x2
//                           FIN  the encoder's final conv (no activation)     -> the view stack (HBM)
//
// All five are the row-stationary, ky-stacked tcgen05 pipeline of conv3x3_umma<64> (K = 576, N = 64: 12 MMAs per row), so
// they advance in lockstep; each hands its rows to the next through a per-stream ring in global memory (16 rows x W x 64
// channels = 256 KB) that stays in L2, with the producer / consumer counters of fuse_wave_umma.cu (release by a publisher
// thread, acquire polls on the consumer's TMA thread, a shared-memory mirror of the ring-space credit).  One ring has TWO
// readers: x1 (R0b's output) is R1a's input and, eight to ten rows later, the skip connection of R1b, which reads it with
// L1-bypassing loads; R0b waits for both before it overwrites a row.  Of the 7.1 GB of DRAM traffic the three launches
// (2 x resblock64_umma + conv3x3_umma<64>) move at C2, only the read of x0 and the write of the view stack remain.
// A stream owns a contiguous range of the flattened (live view, row) space; stage k computes 4 - k halo rows beyond it on both
// sides (clipped to the image).  Results are bit-identical to the launches it replaces: every output element sees the same
// products in the same order and the same rounding points (the intermediate rows are rounded to bf16 in both).
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
constexpr int BTILE_BYTES = 3 * NT * 128;   // one kx B tile: 192 rows x 64 bf16
constexpr int CHUNKS = 1;                   // 64 input channels
constexpr int RING = 8;                     // resident input-row buffers
constexpr int W_BYTES = 3 * CHUNKS * BTILE_BYTES;   // 73,728
constexpr int ROWDONE = 16;                 // "row stored by all epilogue warps" barriers (skew between warps < 16 rows)
constexpr int BAR_OFFSET = W_BYTES + RING * CHUNK_BYTES;
constexpr int BIAS_OFFSET = BAR_OFFSET + 512;
constexpr int SMEM_BYTES = BIAS_OFFSET + NT * 4 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(16 * RING + 16 * ACC_SLOTS + 8 * ROWDONE + 8 + 8 + 4 <= 512, "barrier block overflows into the bias array");
constexpr int STAGES = 5;                   // CTAs per stream
constexpr int FLAG_STRIDE = 32;             // uint32 per counter: every hand-over counter sits in its own 128-byte line
// per stream: prod[k] = rows stage k has stored in ring k (k = 0..3), cons[k] = rows of ring k stage k + 1 has finished reading,
// skip = rows of ring 1 (x1) whose skip connection R1b has consumed
enum : int { F_PROD = 0, F_CONS = 4 * FLAG_STRIDE, F_SKIP = 8 * FLAG_STRIDE, FLAGS_PER_STREAM = 9 * FLAG_STRIDE };
constexpr int SKIP_RING = 1, SKIP_STAGE = 3;   // ring 1 (written by stage 1) is also read by the epilogue of stage 3

struct EncStage {
    const uint8_t* w_img;   // conv3x3_pack_weights image (64 -> 64)
    const float* bias;
    float prelu;
    int has_prelu;
};

struct EncWaveArgs {
    int H, W;
    const int* live_list;              // live views (image index) and their count, device side
    const int* live_count;
    const __nv_bfloat16* x0;           // (n_img, H, W, 64) bf16: output of the first conv
    __nv_bfloat16* out;                // (n_img, H, W, 64) bf16: the view stack
    __nv_bfloat16* ring[4];            // (streams, ring_rows, W, 64) bf16 each
    int ring_rows;
    int streams;
    uint32_t* flags;                   // (streams, FLAGS_PER_STREAM), zero before the launch
    EncStage stage[STAGES];
    int debug_flags;
    unsigned long long* stats;         // optional triage counters, 12 per CTA
};

// The stream's range of the flattened (live view, row) space, as per-image strips extended by `e` halo rows on both sides
// (clipped to the image): FIN computes the range itself, every earlier stage one row more on each side.
struct WaveWalker {
    long long g, g_end;
    int H, e;
    const int* list;
    __device__ WaveWalker(const EncWaveArgs& w, int stream, int e_, bool want_m = true)
        : H(w.H), e(e_), list(want_m ? w.live_list : nullptr) {
        const long long total = static_cast<long long>(*w.live_count) * w.H;
        g = total * stream / w.streams;
        g_end = total * (stream + 1) / w.streams;
    }
    __device__ bool next(Strip& s) {
        if (g >= g_end) return false;
        const long long col = g / H;
        const int y0 = static_cast<int>(g % H);
        const int rows = static_cast<int>(min(static_cast<long long>(H - y0), g_end - g));
        s.m = static_cast<int>(col);
        if (list != nullptr) s.m = list[s.m];
        s.xt = 0;
        s.y0 = max(0, y0 - e);
        s.rows = min(H, y0 + rows + e) - s.y0;
        g += rows;
        return true;
    }
};

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// 32 bytes from a ring row another SM wrote earlier in this launch: L2 only (the line may sit stale in this SM's L1 from the
// ring's previous lap)
__device__ __forceinline__ void ldg_cg_v8(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.cg.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}

// Bounded spin: a protocol bug must surface as a CUDA error, never as a hung GPU.
static __device__ __noinline__ uint32_t flag_acquire_slow(const uint32_t* p, uint32_t want, int tag) {
    const long long t0 = clock64();
    uint32_t v;
    while ((v = ld_acquire(p)) < want) {
        if (clock64() - t0 > HRN_WAIT_LIMIT_CYCLES) {
            printf("hrn_b200: encoder wavefront row wait timed out (block %d thread %d tag %d: have %u, want %u)\n",
                   (int)blockIdx.x, (int)threadIdx.x, tag, v, want);
            __trap();
        }
    }
    return v;
}

// Triage counters (per CTA, cycles), same layout as fuse_wave_umma.cu.
#define WAVE_STAT_BEGIN(var) const long long var = w.stats != nullptr ? clock64() : 0
#define WAVE_STAT_END(var, idx) \
    if (w.stats != nullptr) atomicAdd(w.stats + static_cast<size_t>(blockIdx.x) * 12 + (idx), static_cast<unsigned long long>(clock64() - var))

__global__ void __launch_bounds__(NUM_THREADS, 1)
enc_wave_kernel(const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_r0,
                const __grid_constant__ CUtensorMap map_r1, const __grid_constant__ CUtensorMap map_r2,
                const __grid_constant__ CUtensorMap map_r3, const EncWaveArgs w) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_s = base;
    const uint32_t ring_s = base + W_BYTES;
    const uint32_t bars = base + BAR_OFFSET;
    const uint32_t bar_full = bars;                              // [RING]
    const uint32_t bar_empty = bars + 8 * RING;                  // [RING]
    const uint32_t bar_tfull = bars + 16 * RING;                 // [ACC_SLOTS]
    const uint32_t bar_tempty = bar_tfull + 8 * ACC_SLOTS;       // [ACC_SLOTS]
    const uint32_t bar_rowdone = bar_tempty + 8 * ACC_SLOTS;     // [ROWDONE]
    const uint32_t bar_w = bar_rowdone + 8 * ROWDONE;
    const uint32_t tmem_slot = bar_w + 8;
    const uint32_t credit_slot = tmem_slot + 4;                  // ring rows the consumer(s) have finished reading (mirror)
    uint8_t* smem_gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));
    volatile uint32_t* credit_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (credit_slot - base));
    float* bias_s = reinterpret_cast<float*>(smem_gen + BIAS_OFFSET);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int stream = blockIdx.x / STAGES;
    const int conv = blockIdx.x % STAGES;                        // stage: 0 = R0a, 1 = R0b, 2 = R1a, 3 = R1b, 4 = FIN
    const int halo = STAGES - 1 - conv;                          // rows this stage computes beyond the stream's own range
    const EncStage& cv = w.stage[conv];
    const int R = w.ring_rows;
    uint32_t* flags = w.flags + static_cast<size_t>(stream) * FLAGS_PER_STREAM;
    const CUtensorMap* in_map = conv == 0 ? &map_x0 : (conv == 1 ? &map_r0 : (conv == 2 ? &map_r1 : (conv == 3 ? &map_r2 : &map_r3)));

    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; ++i) {
            ptx::mbar_init(bar_full + 8 * i, 1);
            ptx::mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < ACC_SLOTS; ++i) {
            ptx::mbar_init(bar_tfull + 8 * i, 1);
            ptx::mbar_init(bar_tempty + 8 * i, EPI_WARPS);   // one arrive per epilogue warp
        }
        for (int i = 0; i < ROWDONE; ++i) ptx::mbar_init(bar_rowdone + 8 * i, EPI_WARPS);
        ptx::mbar_init(bar_w, 1);
        *credit_gen = 0;
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(in_map);
    }
    ptx::pdl_launch_dependents();            // the next kernel's prologue may overlap our tail
    if (warp == 2) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
    if (threadIdx.x >= 128 && threadIdx.x < 128 + NT) bias_s[threadIdx.x - 128] = cv.bias[threadIdx.x - 128];
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        // ===================================================== TMA producer: one elected thread
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(bar_w, W_BYTES);
            for (int off = 0; off < W_BYTES; off += 8192) ptx::bulk_copy_g2s(w_s + off, cv.w_img + off, 8192, bar_w);
            WaveWalker walk(w, stream, halo);
            Strip s;
            bool have = walk.next(s);
            ptx::pdl_wait();                 // weights and lists are constants; x0 comes from the previous kernel
            uint32_t it = 0, cum = 0, seen = 0;   // cum: in-image input rows requested so far = rows the producer must have stored
            WAVE_STAT_BEGIN(t9);
            const uint32_t* prod = flags + F_PROD + (conv - 1) * FLAG_STRIDE;
            for (; have; have = walk.next(s)) {
                for (int q = 0; q < s.rows + 2; ++q, ++it) {
                    const int y = s.y0 - 1 + q;
                    int yc = y;                              // row coordinate of the load (rows outside the image: TMA zero fill)
                    if (conv != 0) {
                        if (y >= 0 && y < w.H) {
                            yc = static_cast<int>(cum % static_cast<uint32_t>(R));
                            ++cum;
                            if (seen < cum && !(w.debug_flags & 32)) {
                                WAVE_STAT_BEGIN(t0);
                                const uint32_t v = ld_acquire(prod);                 // pairs with the publisher's st.release
                                seen = v >= cum ? v : flag_acquire_slow(prod, cum, 20);
                                fence_proxy_async_global();  // generic-proxy writes -> async-proxy (TMA) reads
                                WAVE_STAT_END(t0, 1);
                            }
                        } else {
                            yc = -1;
                        }
                    }
                    const uint32_t slot = it % RING, ph = (it / RING) & 1;
                    {
                        WAVE_STAT_BEGIN(t1);
                        ptx::mbar_wait(bar_empty + 8 * slot, ph ^ 1, 1);
                        WAVE_STAT_END(t1, 2);
                    }
                    ptx::mbar_expect_tx(bar_full + 8 * slot, CHUNK_TX);
                    ptx::tma_load_4d(ring_s + slot * CHUNK_BYTES, in_map, 0, -1, yc, conv == 0 ? s.m : stream, bar_full + 8 * slot);
                }
            }
            WAVE_STAT_END(t9, 9);
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer: ONE elected thread (see conv3x3_umma.cu)
        if (ptx::elect_one()) {
            constexpr uint32_t idesc_base = ptx::umma_idesc_bf16(TILE_M, 0);
            constexpr uint32_t idesc64 = ptx::umma_idesc_bf16(TILE_M, NT);
            constexpr uint32_t BLK = NT * 128 / 16;               // one 64-row ky block, in descriptor units (16 B)
            constexpr uint32_t B_KX = CHUNKS * BTILE_BYTES / 16 - 6;   // k-step 3 of kx -> k-step 0 of kx + 1
            const uint32_t a_lo0 = desc_lo(ring_s), b_lo0 = desc_lo(w_s);
            ptx::mbar_wait(bar_w, 0, 2);
            uint32_t it = 0, tile0 = 0;
            bool full_seen = false, tempty_seen = false;          // waits already done by the previous item
            WaveWalker walk(w, stream, halo, false);
            Strip s;
            bool have = walk.next(s);
            while (have) {
                Strip nxt;
                const bool have_next = walk.next(nxt);
                for (int q = 0; q < s.rows + 2; ++q) {
                    if (q >= 2 && q <= s.rows - 1) {
                        // ---- interior row: all three ky blocks are live and block 2 opens the accumulator of output row q
                        const uint32_t t_new = tile0 + q;
                        const uint32_t sl = (t_new - 2) % ACC_SLOTS;             // slot of output row q - 2
                        if (!tempty_seen)
                            ptx::mbar_wait(bar_tempty + 8 * (t_new % ACC_SLOTS), ((t_new / ACC_SLOTS) & 1) ^ 1, 4);
                        tempty_seen = false;
                        const bool nxt_opens = (q + 1 <= s.rows - 1);
#pragma unroll
                        for (int c = 0; c < CHUNKS; ++c, ++it) {
                            const uint32_t slot = it % RING;
                            if (!full_seen) {
                                WAVE_STAT_BEGIN(t7);
                                ptx::mbar_wait(bar_full + 8 * slot, (it / RING) & 1, 3);
                                WAVE_STAT_END(t7, 7);
                            }
                            full_seen = false;
                            ptx::tc_fence_after();
                            uint64_t ad = make_desc(a_lo0 + slot * (CHUNK_BYTES / 16));
                            uint64_t bd = make_desc(b_lo0 + c * (BTILE_BYTES / 16));
                            const uint32_t dA = tmem_base + sl * NT;
                            // early waits for the next item, issued after k-step 7 while the MMA queue is full
                            auto early = [&]() {
                                const uint32_t itn = it + 1;
                                WAVE_STAT_BEGIN(t7);
                                ptx::mbar_wait(bar_full + 8 * (itn % RING), (itn / RING) & 1, 6);
                                WAVE_STAT_END(t7, 7);
                                full_seen = true;
                                if (c + 1 == CHUNKS && nxt_opens) {
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
                            ptx::umma_commit(bar_empty + 8 * slot);
                            if (c == CHUNKS - 1) ptx::umma_commit(bar_tfull + 8 * sl);
                        }
                        continue;
                    }
                    // ---- boundary rows of a strip (q = 0, 1, rows, rows + 1, or very short strips): generic path.
                    // Input row q feeds output rows o = q - ky; B block (2 - ky) <-> output row q - ky.
                    const int ky_lo = max(0, q - (s.rows - 1)), ky_hi = min(2, q);
                    const int blk_lo = 2 - ky_hi, nblk = ky_hi - ky_lo + 1;
                    const uint32_t t_lo = tile0 + q - ky_hi;
                    const uint32_t s_lo = t_lo % ACC_SLOTS;
                    const bool opens = (ky_lo == 0);                    // the last block starts a new accumulator
                    if (opens && !tempty_seen) {
                        const uint32_t t_new = tile0 + q;
                        ptx::mbar_wait(bar_tempty + 8 * (t_new % ACC_SLOTS), ((t_new / ACC_SLOTS) & 1) ^ 1, 4);
                    }
                    tempty_seen = false;
                    const int w0 = min(nblk, ACC_SLOTS - static_cast<int>(s_lo)), w1 = nblk - w0;
                    const uint32_t d0 = tmem_base + s_lo * NT, d1 = tmem_base;
                    const uint32_t id0 = idesc_base | (static_cast<uint32_t>(w0 * NT >> 3) << 17);
                    const uint32_t id1 = idesc_base | (static_cast<uint32_t>(w1 * NT >> 3) << 17);
                    const bool last_row = (q == s.rows + 1);
                    const bool more_rows = !last_row || have_next;
                    const bool next_opens = last_row ? true : (q + 1 <= s.rows - 1);
                    const uint32_t t_next = last_row ? tile0 + s.rows : tile0 + q + 1;
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c, ++it) {
                        const uint32_t slot = it % RING;
                        if (!full_seen) ptx::mbar_wait(bar_full + 8 * slot, (it / RING) & 1, 3);
                        full_seen = false;
                        ptx::tc_fence_after();
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
                                const bool next_item = (c + 1 < CHUNKS) || more_rows;
                                if (next_item) {
                                    const uint32_t itn = it + 1;
                                    ptx::mbar_wait(bar_full + 8 * (itn % RING), (itn / RING) & 1, 6);
                                    full_seen = true;
                                    if (c + 1 == CHUNKS && next_opens) {
                                        ptx::mbar_wait(bar_tempty + 8 * (t_next % ACC_SLOTS),
                                                       ((t_next / ACC_SLOTS) & 1) ^ 1, 7);
                                        tempty_seen = true;
                                    }
                                }
                            }
                        }
                        ptx::umma_commit(bar_empty + 8 * slot);                 // this (row, chunk) buffer is consumed
                        if (c == CHUNKS - 1 && ky_hi == 2)                      // output row q-2 has all 9 taps
                            ptx::umma_commit(bar_tfull + 8 * ((tile0 + q - 2) % ACC_SLOTS));
                    }
                }
                tile0 += s.rows;
                s = nxt;
                have = have_next;
            }
        }
    } else if (warp == 2) {
        // ===================================================== credit poller (stages with a ring to fill): ONE thread watches
        // the reader's counter(s) in global memory and mirrors the minimum into shared memory for the eight epilogue warps
        if (conv < STAGES - 1 && !(w.debug_flags & 32) && ptx::elect_one()) {
            uint32_t total_rows = 0;
            WaveWalker count(w, stream, halo, false);
            Strip s;
            for (bool have = count.next(s); have; have = count.next(s)) total_rows += s.rows;
            const uint32_t* cons = flags + F_CONS + conv * FLAG_STRIDE;
            const uint32_t last_needed = total_rows > static_cast<uint32_t>(R) ? total_rows - R : 0;   // the last store waits for this
            uint32_t have_credit = 0;
            const long long t0 = clock64();
            while (have_credit < last_needed) {
                uint32_t v = ld_relaxed(cons);
                if (conv == SKIP_RING) v = min(v, ld_relaxed(flags + F_SKIP));      // x1 has a second reader
                if (v > have_credit) {
                    have_credit = v;
                    *credit_gen = v;
                }
                if (clock64() - t0 > 40 * HRN_WAIT_LIMIT_CYCLES) {
                    printf("hrn_b200: encoder wavefront credit poller timed out (block %d: have %u, want %u)\n", (int)blockIdx.x, have_credit, last_needed);
                    __trap();
                }
            }
        }
    } else if (warp == 3) {
        // ===================================================== publisher: hand-over counters for the neighbours in the stream
        if (ptx::elect_one()) {
            WaveWalker walk(w, stream, halo, false);
            WaveWalker walk_skip(w, stream, STAGES - 1 - SKIP_RING, false);   // the strips of the stage that writes ring 1 (x1)
            Strip s, sk;
            uint32_t k = 0, in_base = 0, skip_base = 0, next_pub = 1, total_rows = 0;
            {
                WaveWalker count(w, stream, halo, false);
                for (bool have = count.next(s); have; have = count.next(s)) total_rows += s.rows;
            }
            uint32_t* prod = flags + F_PROD + conv * FLAG_STRIDE;
            uint32_t* my_cons = flags + F_CONS + (conv - 1) * FLAG_STRIDE;
            for (bool have = walk.next(s); have; have = walk.next(s)) {
                walk_skip.next(sk);
                const int in_lo = max(0, s.y0 - 1), in_hi = min(w.H, s.y0 + s.rows + 1);
                for (int i = 0; i < s.rows; ++i, ++k) {
                    {
                        WAVE_STAT_BEGIN(t5);
                        ptx::mbar_wait(bar_rowdone + 8 * (k % ROWDONE), (k / ROWDONE) & 1, 8);   // all eight epilogue warps are done with row k
                        WAVE_STAT_END(t5, 5);
                    }
                    WAVE_STAT_BEGIN(t6);
                    while (i + 1 < s.rows && ptx::mbar_test_wait(bar_rowdone + 8 * ((k + 1) % ROWDONE), ((k + 1) / ROWDONE) & 1)) {
                        ++i;
                        ++k;
                    }
                    // the tensor pipe has finished every input row up to y + 1 of this strip: those ring rows may be overwritten
                    if (conv > 0) st_relaxed(my_cons, in_base + static_cast<uint32_t>(min(s.y0 + i + 2, in_hi) - in_lo));
                    // R1b: the skip rows of x1 up to image row y have been read (at the end of a strip: all of the writer's rows)
                    if (conv == SKIP_STAGE)
                        st_relaxed(flags + F_SKIP, skip_base + static_cast<uint32_t>(i + 1 == s.rows ? sk.rows : s.y0 + i + 1 - sk.y0));
                    if (conv < STAGES - 1 && (k + 1 >= next_pub || k + 1 == total_rows)) {   // rows 0 .. k are in the ring
                        fence_proxy_async_global();      // the rows will be read through the async proxy (TMA)
                        st_release(prod, k + 1);
                        next_pub = k + 2;
                    }
                    WAVE_STAT_END(t6, 6);
                }
                in_base += static_cast<uint32_t>(in_hi - in_lo);
                skip_base += static_cast<uint32_t>(sk.rows);
            }
        }
    } else if (warp >= 4) {
        // ===================================================== epilogue: 8 warps, (lane quadrant) x (column half)
        const int wq = warp & 3;                 // TMEM lanes [32 wq, 32 wq + 32)
        const int hf = (warp - 4) >> 2;          // accumulator columns [32 hf, 32 hf + 32)
        const int co0 = hf * 32;                 // first output channel handled by this thread
        float bias_r[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) bias_r[e] = bias_s[hf * 32 + e];
        const bool has_prelu = cv.has_prelu != 0;
        const float slope_m1 = cv.prelu - 1.0f;  // PReLU(v) = v + (slope - 1) * min(v, 0)
        __nv_bfloat16* ring_out = conv < STAGES - 1 ? w.ring[conv] : nullptr;
        uint32_t seen_c = 0;
        WaveWalker walk(w, stream, halo);
        WaveWalker walk_skip(w, stream, STAGES - 1 - SKIP_RING, false);
        Strip s, sk;
        bool have = walk.next(s);
        ptx::pdl_wait();                         // residual reads and output writes touch the previous kernel's tensors
        uint32_t tile = 0;                       // rows produced so far = ring row counter of a producer
        uint32_t skip_base = 0;                  // rows the writer of ring 1 stored in the earlier strips
        const int x = wq * 32 + lane;
        const bool valid = x < w.W;
        for (; have; have = walk.next(s)) {
            walk_skip.next(sk);
            // skip connection: R0b adds x0 (a tensor of the previous kernel), R1b adds x1 (ring 1, written by R0b in this launch)
            const bool res_x0 = conv == 1, res_ring = conv == SKIP_STAGE;
            const bool use_res = (res_x0 || res_ring) && valid && !(w.debug_flags & 8);
            const __nv_bfloat16* rp = nullptr;
            if (res_x0) rp = w.x0 + ((static_cast<size_t>(s.m) * w.H + s.y0) * w.W + x) * 64 + co0;
            const uint32_t ring_row0 = skip_base + static_cast<uint32_t>(s.y0 - sk.y0);   // ring-1 row counter of image row s.y0
            auto res_ptr = [&](uint32_t rr) {
                return w.ring[SKIP_RING] + ((static_cast<size_t>(stream) * R + rr % R) * w.W + x) * 64 + co0;
            };
            const size_t r_step = static_cast<size_t>(w.W) * 64;
            __nv_bfloat16* op = nullptr;
            if (conv == STAGES - 1) op = w.out + ((static_cast<size_t>(s.m) * w.H + s.y0) * w.W + x) * 64 + co0;
            // The residual row is fetched TWO rows ahead, rotating three register buffers: a 64-channel row takes the tensor
            // pipe 1150 cycles, less than an L2 round trip under load, and a load issued one row ahead stalled the epilogue
            // (2100-3000 instead of 1500 cycles per row: the whole stream ran at the pace of the two skip stages).  x0 belongs
            // to the previous kernel and may be read at any time; rows y + 1 and y + 2 of x1 are known to be in ring 1 once the
            // accumulator of row y is complete (that needed R1a's rows up to y + 1, hence R0b's up to y + 2), so their loads
            // are issued behind that wait.
            uint32_t rv0[2][8], rv1[2][8], rv2[2][8];
            auto load_res = [&](int i, uint32_t (&dst)[2][8]) {      // residual row i of this strip
                if (res_x0) {
                    ptx::ldg_nc_v8(rp + static_cast<size_t>(i) * r_step, dst[0]);
                    ptx::ldg_nc_v8(rp + static_cast<size_t>(i) * r_step + 16, dst[1]);
                } else {
                    const __nv_bfloat16* q = res_ptr(ring_row0 + static_cast<uint32_t>(i));
                    ldg_cg_v8(q, dst[0]);
                    ldg_cg_v8(q + 16, dst[1]);
                }
            };
            if (use_res && res_x0) {
                load_res(0, rv0);
                if (s.rows > 1) load_res(1, rv1);
            }
            auto do_row = [&](int i, uint32_t (&rv)[2][8], uint32_t (&rv_n1)[2][8], uint32_t (&rv_n2)[2][8]) {
                const uint32_t acc = tile % ACC_SLOTS, aph = (tile / ACC_SLOTS) & 1;
                if (use_res && res_x0 && i + 2 < s.rows) load_res(i + 2, rv_n2);
                {
                    WAVE_STAT_BEGIN(t4);
                    ptx::mbar_wait(bar_tfull + 8 * acc, aph, 5);
                    if (warp == 4 && lane == 0) WAVE_STAT_END(t4, 4);
                }
                if (use_res && res_ring) {
                    if (i == 0) {
                        load_res(0, rv);
                        if (s.rows > 1) load_res(1, rv_n1);
                    }
                    if (i + 2 < s.rows) load_res(i + 2, rv_n2);
                }
                ptx::tc_fence_after();
                uint32_t v[32];
                ptx::tmem_ld_x32(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * NT + hf * 32, v);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * acc);
                uint32_t o[2][8];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float x0v = __uint_as_float(v[2 * e]) + bias_r[2 * e];
                    float x1v = __uint_as_float(v[2 * e + 1]) + bias_r[2 * e + 1];
                    if (has_prelu) {
                        x0v = fmaf(slope_m1, fminf(x0v, 0.0f), x0v);
                        x1v = fmaf(slope_m1, fminf(x1v, 0.0f), x1v);
                    }
                    __nv_bfloat162 yv = __floats2bfloat162_rn(x0v, x1v);
                    if (use_res) yv = __hadd2(yv, *reinterpret_cast<const __nv_bfloat162*>(&rv[e >> 3][e & 7]));
                    o[e >> 3][e & 7] = *reinterpret_cast<const uint32_t*>(&yv);
                }
                if (conv < STAGES - 1) {
                    // ring slot of produced row `tile`; it may be overwritten once the reader(s) are done with row tile - R
                    if (tile >= static_cast<uint32_t>(R) && !(w.debug_flags & 32)) {
                        const uint32_t want = tile - R + 1;
                        if (seen_c < want) {                 // shared-memory mirror kept by the credit poller (warp 2)
                            WAVE_STAT_BEGIN(t3);
                            const long long t0 = clock64();
                            while ((seen_c = *credit_gen) < want) {
                                if (clock64() - t0 > HRN_WAIT_LIMIT_CYCLES) {
                                    printf("hrn_b200: encoder wavefront ring-space wait timed out (block %d warp %d: have %u, want %u)\n",
                                           (int)blockIdx.x, warp, seen_c, want);
                                    __trap();
                                }
                            }
                            if (warp == 4 && lane == 0) WAVE_STAT_END(t3, 3);
                        }
                    }
                    if (valid && !(w.debug_flags & 2)) {
                        __nv_bfloat16* dst = ring_out + ((static_cast<size_t>(stream) * R + tile % R) * w.W + x) * 64 + co0;
                        ptx::stg_v8(dst, o[0]);
                        ptx::stg_v8(dst + 16, o[1]);
                    }
                } else {
                    if (valid) {
                        ptx::stg_v8(op, o[0]);
                        ptx::stg_v8(op + 16, o[1]);
                    }
                    op += r_step;
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar_rowdone + 8 * (tile % ROWDONE));      // -> publisher (warp 3)
                ++tile;
            };
            for (int i = 0; i < s.rows; i += 3) {
                do_row(i, rv0, rv1, rv2);
                if (i + 1 < s.rows) do_row(i + 1, rv1, rv2, rv0);
                if (i + 2 < s.rows) do_row(i + 2, rv2, rv0, rv1);
            }
            skip_base += static_cast<uint32_t>(sk.rows);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 2) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace

int enc_wave_streams(int sm_count) { return sm_count / STAGES; }
size_t enc_wave_ring_bytes(int sm_count, int ring_rows, int W) {
    return static_cast<size_t>(enc_wave_streams(sm_count)) * ring_rows * W * 64 * sizeof(__nv_bfloat16);
}
size_t enc_wave_flag_bytes(int sm_count) { return static_cast<size_t>(enc_wave_streams(sm_count)) * FLAGS_PER_STREAM * sizeof(uint32_t); }

// Every CTA of the grid has to be resident at once (see fuse_wave_fits).
bool enc_wave_fits(int sm_count) {
    static int max_ctas[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    static bool attr_set[64] = {};
    if (allow_dynamic_smem(enc_wave_kernel, SMEM_BYTES, attr_set)) return false;
    std::lock_guard<std::mutex> lock(lazy_init_mutex());
    if (max_ctas[dev] == 0) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, enc_wave_kernel, NUM_THREADS, SMEM_BYTES) != cudaSuccess) {
            cudaGetLastError();
            per_sm = 0;
        }
        max_ctas[dev] = per_sm > 0 ? per_sm * sm_count : -1;
    }
    return max_ctas[dev] >= enc_wave_streams(sm_count) * STAGES;
}

int enc_wave_launch(const EncWaveLaunch& a, int sm_count, cudaStream_t stream) {
    if (a.W > TILE_M || a.W <= 0 || a.H <= 0) {
        set_error("enc_wave: images wider than %d pixels take the per-layer launches", TILE_M);
        return -1;
    }
    int streams = enc_wave_streams(sm_count);
    if (a.streams > 0 && a.streams < streams) streams = a.streams;
    if (streams < 1 || a.ring_rows < 12) {
        set_error("enc_wave: needs at least %d SMs and 12 ring rows", STAGES);
        return -1;
    }
    EncWaveArgs w{};
    w.H = a.H;
    w.W = a.W;
    w.live_list = a.live_list;
    w.live_count = a.live_count;
    w.x0 = a.x0;
    w.out = a.out;
    for (int i = 0; i < 4; ++i) w.ring[i] = a.ring[i];
    w.ring_rows = a.ring_rows;
    w.streams = streams;
    w.flags = a.flags;
    w.debug_flags = a.debug_flags;
    w.stats = a.stats;
    for (int i = 0; i < STAGES; ++i) {
        w.stage[i].w_img = a.w_img[i];
        w.stage[i].bias = a.bias[i];
        w.stage[i].prelu = a.prelu[i];
        w.stage[i].has_prelu = a.has_prelu[i];
    }
    CUtensorMap map_x0, map_r[4];
    if (encode_nhwc_map(&map_x0, a.x0, 64, a.W, a.H, a.n_img, SLOT_PIX)) return -1;
    for (int i = 0; i < 4; ++i)
        if (encode_nhwc_map(&map_r[i], a.ring[i], 64, a.W, a.ring_rows, streams, SLOT_PIX)) return -1;
    static bool attr_set[64] = {};
    if (allow_dynamic_smem(enc_wave_kernel, SMEM_BYTES, attr_set)) return -1;
    HRN_CUDA_OK(launch_pdl(enc_wave_kernel, streams * STAGES, NUM_THREADS, SMEM_BYTES, stream, 1, map_x0, map_r[0], map_r[1],
                           map_r[2], map_r[3], w));
    note_launches(1);
    return 0;
}

}  // namespace hrn
