// One fusion level of RecuversiveNet (HRNet.py:99-134) in ONE launch: the three 3x3 convolutions of `fuse`
// (ResidualBlock(128) = conv A, conv B + skip; conv C 128 -> 64 + PReLU; HRNet.py:93-97) and the alpha merge
// `alice + alpha_bob * x` (HRNet.py:123-128) run as a ROW WAVEFRONT through the SMs instead of three
// launches that each round-trip a 128-channel tensor through HBM.
//
// The SMs are dealt out in "streams" of five CTAs (29 streams = 145 CTAs on a B200):
//
//     A0 A1   conv A on cat(alice, bob): the two 64-channel output halves       (K = 2 x 64 x 9, N = 64 each)
//     B0 B1   conv B on A's output + the ResidualBlock skip onto cat(alice, bob)
//     C       conv C on B's output + PReLU + alice + alpha_bob * x  -> next level's view stack
//
// Every CTA is the row-stationary, ky-stacked tcgen05 pipeline of conv3x3_umma.cu (one TMA thread, one MMA thread,
// eight epilogue warps, the CTA's 144 KB weight slice resident in shared memory), so all five advance at the same rows
// per second.  A stream owns a contiguous range of the flattened (pair, row) space; A writes its output rows into a
// small per-stream RING in global memory (16 rows x W x 128 channels = 512 KB) that B reads a few rows later, and B feeds
// C through a second ring.  The rings of all streams together are ~30 MB: they live in the 126 MB L2, are overwritten
// while still dirty there, and never reach HBM -- the 128-channel intermediates t1 / t2 of the three-launch schedule
// (4.2 of the 7.4 GB that fusion level 1 moves at B = 32, L = 16) disappear, and B's skip / C's alice re-reads hit rows
// A pulled into L2 microseconds earlier.
//
// Hand-over between the CTAs of a stream: device-scope counters in global memory.
//   producer:  epilogue warps store a row (generic proxy), __syncwarp, one lane arrives on a shared-memory mbarrier; the
//              publisher thread (warp 3) waits for all eight warps and does st.release.gpu(rows produced)
//   consumer:  its TMA thread spins on ld.acquire.gpu until the row is there, fence.proxy.async, then the TMA load
//   ring reuse: the consumer's epilogue publishes how many input rows its tensor pipe has finished reading (implied by
//              the accumulator-full commit); the producer's epilogue warps wait for row k - RING_ROWS before storing row k
// All CTAs of the grid are resident at once (one per SM, grid <= SM count), so the spin waits cannot deadlock; every
// wait is bounded and traps instead of hanging the GPU.
//
// Unlike the three-launch schedule this level is NOT in place: C writes into a second view stack (`stack_out`), because
// a neighbouring stream may still be reading alice's halo rows at a range boundary.  Pairs whose bob has alpha = 0 are
// not computed (HRNet.py:123-128 leaves alice unchanged); when the next level needs such an alice, the C CTAs copy it
// over ("carry" list of live_lists_kernel).  Output bits are identical to the three-launch schedule: every output
// element sees the same products in the same order and the same rounding points.
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
constexpr int CHUNKS = 2;                   // 128 input channels
constexpr int RING = 4;                     // resident (input row, chunk) buffers = 2 rows.  A fifth slot (16,640-byte slots) fits
                                            // and is bit-identical, but measured no faster: the loads are not what the streams wait for
constexpr int W_BYTES = 3 * CHUNKS * BTILE_BYTES;   // 147,456
constexpr int ROWDONE = 16;                 // "row stored by all epilogue warps" barriers (skew between warps < 16 rows)
constexpr int BAR_OFFSET = W_BYTES + RING * CHUNK_BYTES;
constexpr int BIAS_OFFSET = BAR_OFFSET + 512;
constexpr int SMEM_BYTES = BIAS_OFFSET + NT * 4 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(16 * RING + 16 * ACC_SLOTS + 8 * ROWDONE + 8 + 8 + 4 <= 512, "barrier block overflows into the bias array");
static_assert((W_BYTES + RING * CHUNK_BYTES) % 128 == 0, "barrier block alignment");
constexpr int CTAS_PER_STREAM = 5;
constexpr int FLAG_STRIDE = 32;             // uint32 per counter: every hand-over counter sits in its own 128-byte line
constexpr int FLAGS_PER_STREAM = 8 * FLAG_STRIDE;   // prod1[2], cons1[2], prod2[2], cons2, pad
enum : int { F_PROD1 = 0, F_CONS1 = 2 * FLAG_STRIDE, F_PROD2 = 4 * FLAG_STRIDE, F_CONS2 = 6 * FLAG_STRIDE };

struct WaveConv {
    const uint8_t* w_img;   // pre-swizzled weights of the whole layer (conv3x3_pack_weights)
    const float* bias;
    float prelu;
    int has_prelu;
};

struct WaveArgs {
    int H, W;
    int half, src_views, top;          // pair (b, i) = views i and top - 1 - i of imageset b; view stack stride = src_views
    const int* live_list;              // live pairs (index b * half + i) and their count, device side
    const int* live_count;
    const int* carry_list;             // dead pairs whose alice the next level needs
    const int* carry_count;
    const __nv_bfloat16* stack_in;     // (B * src_views, H, W, 64) bf16
    __nv_bfloat16* stack_out;
    __nv_bfloat16* ring1;              // (streams, ring_rows, W, 128) bf16
    __nv_bfloat16* ring2;
    int ring_rows;
    int streams;
    uint32_t* flags;                   // (streams, FLAGS_PER_STREAM), zero before the launch
    const float* alphas;
    int alpha_stride, alpha_residual;
    WaveConv conv[3];
    int debug_flags;
    unsigned long long* stats;         // optional (triage): per CTA 8 counters of clock64 cycles, see fuse_wave_kernel
    int publish_rows;                  // rows per "stored" publication (one device-scope release each), >= 1
    int lag_rows;                      // a consumer loads row k only when the producer has stored row k + lag_rows (triage)
};

// The stream's range of the flattened (live pair, row) space, as per-image strips extended by `e` halo rows on both
// sides (clipped to the image): conv C computes the range itself, B one row more on each side, A two.
struct WaveWalker {
    long long g, g_end;
    int H, e;
    const int* list;
    __device__ WaveWalker(const WaveArgs& w, int stream, int e_, bool want_m = true)
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
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Spin until *p >= want.  Bounded like the mbarrier waits: a protocol bug must surface as a CUDA error, never as a hung GPU.
static __device__ __noinline__ uint32_t flag_wait_slow(const uint32_t* p, uint32_t want, int tag) {
    const long long t0 = clock64();
    uint32_t v;
    while ((v = ld_relaxed(p)) < want) {
        if (clock64() - t0 > HRN_WAIT_LIMIT_CYCLES) {
            printf("hrn_b200: wavefront flag wait timed out (block %d thread %d tag %d: have %u, want %u)\n",
                   (int)blockIdx.x, (int)threadIdx.x, tag, v, want);
            __trap();
        }
    }
    return v;
}
// Relaxed polling; the caller adds the acquire fence where data written by the flag's producer is read afterwards (the
// ring-reuse wait needs none: it only orders this CTA's later stores after the consumer's completed reads).
__device__ __forceinline__ uint32_t flag_wait(const uint32_t* p, uint32_t want, int tag) {
    const uint32_t v = ld_relaxed(p);
    return v >= want ? v : flag_wait_slow(p, want, tag);
}
__device__ __forceinline__ void fence_acquire_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// Consumer side: the poll that succeeds must be an acquire (it pairs with the publisher's st.release).  ld.acquire.gpu is a
// strong load plus an L1 invalidate; a relaxed load followed by fence.acq_rel.gpu costs a full MEMBAR.GPU on the TMA thread
// for every batch of rows (4 % of the fusion stage, tools/wave_time.py).
static __device__ __noinline__ uint32_t flag_acquire_slow(const uint32_t* p, uint32_t want, int tag) {
    const long long t0 = clock64();
    uint32_t v;
    while ((v = ld_acquire(p)) < want) {
        if (clock64() - t0 > HRN_WAIT_LIMIT_CYCLES) {
            printf("hrn_b200: wavefront row wait timed out (block %d thread %d tag %d: have %u, want %u)\n",
                   (int)blockIdx.x, (int)threadIdx.x, tag, v, want);
            __trap();
        }
    }
    return v;
}

// Triage counters (per CTA, cycles): 0 = whole kernel, 1 = TMA thread waiting for the producer's rows, 2 = TMA thread
// waiting for a free shared-memory slot, 3 = epilogue warp 4 waiting for ring space, 4 = epilogue warp 4 waiting for a full
// accumulator, 5 = publisher waiting for stored rows, 6 = publisher inside its stores (release fence), 7 = MMA thread waiting
// for an input row in shared memory (interior rows).
#define WAVE_STAT_BEGIN(var) const long long var = w.stats != nullptr ? clock64() : 0
#define WAVE_STAT_END(var, idx) \
    if (w.stats != nullptr) atomicAdd(w.stats + static_cast<size_t>(blockIdx.x) * 12 + (idx), static_cast<unsigned long long>(clock64() - var))

__global__ void __launch_bounds__(NUM_THREADS, 1)
fuse_wave_kernel(const __grid_constant__ CUtensorMap map_stack, const __grid_constant__ CUtensorMap map_r1,
                 const __grid_constant__ CUtensorMap map_r2, const WaveArgs w) {
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
    const int stream = blockIdx.x / CTAS_PER_STREAM;
    const int role = blockIdx.x % CTAS_PER_STREAM;
    const int conv = role < 2 ? 0 : (role < 4 ? 1 : 2);         // 0 = A, 1 = B, 2 = C
    const int part = conv < 2 ? (role & 1) : 0;                  // 64-channel output half
    const int halo = 2 - conv;                                   // rows this conv computes beyond the stream's own range
    const WaveConv& cv = w.conv[conv];
    const int R = w.ring_rows;
    uint32_t* flags = w.flags + static_cast<size_t>(stream) * FLAGS_PER_STREAM;
    const CUtensorMap* in_map = conv == 0 ? &map_stack : (conv == 1 ? &map_r1 : &map_r2);

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
    if (threadIdx.x >= 128 && threadIdx.x < 128 + NT) bias_s[threadIdx.x - 128] = cv.bias[part * NT + threadIdx.x - 128];
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    WAVE_STAT_BEGIN(t_kernel);

    if (warp == 0) {
        // ===================================================== TMA producer: one elected thread
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(bar_w, W_BYTES);
            const uint8_t* wsrc = cv.w_img + static_cast<size_t>(part) * W_BYTES;
            for (int off = 0; off < W_BYTES; off += 8192) ptx::bulk_copy_g2s(w_s + off, wsrc + off, 8192, bar_w);
            WaveWalker walk(w, stream, halo);
            Strip s;
            bool have = walk.next(s);
            ptx::pdl_wait();                 // weights and lists are constants; the view stack comes from the previous kernel
            uint32_t total_in = 0;           // in-image input rows of this CTA = rows its producer will ever store
            if (conv != 0 && w.lag_rows > 0) {
                WaveWalker count(w, stream, halo, false);
                Strip c;
                for (bool more = count.next(c); more; more = count.next(c)) total_in += min(w.H, c.y0 + c.rows + 1) - max(0, c.y0 - 1);
            }
            uint32_t it = 0, cum = 0;        // cum: in-image input rows requested so far = rows the producer must have stored
            uint32_t seen0 = 0, seen1 = 0;   // last values read from the producer's two "rows stored" counters
            const uint32_t* prod = flags + (conv == 1 ? F_PROD1 : F_PROD2);
            WAVE_STAT_BEGIN(t9);
            const bool l2_hints = !(w.debug_flags & 2048);   // debug flag 2048 = no L2 eviction hints (A/B)
            const uint64_t keep = ptx::l2_policy_evict_last();
            for (; have; have = walk.next(s)) {
                int img[2] = {0, 0};
                if (conv == 0) {
                    const int b = s.m / w.half, i = s.m % w.half;
                    img[0] = b * w.src_views + i;
                    img[1] = b * w.src_views + (w.top - 1 - i);
                }
                for (int q = 0; q < s.rows + 2; ++q) {
                    const int y = s.y0 - 1 + q;
                    int yc = y;                              // row coordinate of the load (rows outside the image: TMA zero fill)
                    if (conv != 0) {
                        if (y >= 0 && y < w.H) {
                            yc = static_cast<int>(cum % static_cast<uint32_t>(R));
                            ++cum;
                            if (!(w.debug_flags & 32)) {
                                const uint32_t need = w.lag_rows > 0 ? min(cum + static_cast<uint32_t>(w.lag_rows), total_in) : cum;
                                if (seen0 < need || seen1 < need) {
                                    WAVE_STAT_BEGIN(t0);
                                    if (w.debug_flags & 128) {               // A/B: relaxed polls + one device-scope fence
                                        if (seen0 < need) seen0 = flag_wait(prod, need, 20);
                                        if (seen1 < need) seen1 = flag_wait(prod + FLAG_STRIDE, need, 21);
                                        fence_acquire_gpu();
                                    } else {                                 // acquire polls: pair with the publisher's st.release
                                        const uint32_t v0 = ld_acquire(prod), v1 = ld_acquire(prod + FLAG_STRIDE);   // both in flight
                                        seen0 = v0 >= need ? v0 : flag_acquire_slow(prod, need, 20);
                                        seen1 = v1 >= need ? v1 : flag_acquire_slow(prod + FLAG_STRIDE, need, 21);
                                    }
                                    WAVE_STAT_END(t0, 1);
                                    WAVE_STAT_BEGIN(t8);
                                    if (!(w.debug_flags & 256)) fence_proxy_async_global();  // generic-proxy writes -> async-proxy (TMA) reads
                                    WAVE_STAT_END(t8, 8);
                                }
                            }
                        } else {
                            yc = -1;
                        }
                    }
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c, ++it) {
                        const uint32_t slot = it % RING, ph = (it / RING) & 1;
                        WAVE_STAT_BEGIN(t1);
                        ptx::mbar_wait(bar_empty + 8 * slot, ph ^ 1, 1);
                        WAVE_STAT_END(t1, 2);
                        ptx::mbar_expect_tx(bar_full + 8 * slot, CHUNK_TX);
                        if (conv == 0 && l2_hints)      // view rows come back as the skip / alice rows of conv B and conv C: keep them in L2
                            ptx::tma_load_4d_hint(ring_s + slot * CHUNK_BYTES, in_map, 0, -1, yc, img[c], bar_full + 8 * slot, keep);
                        else if (conv == 0)
                            ptx::tma_load_4d(ring_s + slot * CHUNK_BYTES, in_map, 0, -1, yc, img[c], bar_full + 8 * slot);
                        else if (l2_hints && !(w.debug_flags & 4096))   // ring rows: written and read in L2, overwritten 16 rows later
                            ptx::tma_load_4d_hint(ring_s + slot * CHUNK_BYTES, in_map, 64 * c, -1, yc, stream, bar_full + 8 * slot, keep);
                        else
                            ptx::tma_load_4d(ring_s + slot * CHUNK_BYTES, in_map, 64 * c, -1, yc, stream, bar_full + 8 * slot);
                    }
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
        // ===================================================== credit poller (conv A, B): ONE thread per CTA watches the
        // consumer's "rows read" counter(s) in global memory and mirrors the minimum into shared memory, where the eight
        // epilogue warps look it up; without it every epilogue warp would spin on the same global line.
        if (conv < 2 && !(w.debug_flags & 1024) && ptx::elect_one()) {
            uint32_t total_rows = 0;
            WaveWalker count(w, stream, halo, false);
            Strip s;
            for (bool have = count.next(s); have; have = count.next(s)) total_rows += s.rows;
            const uint32_t* cons = flags + (conv == 0 ? F_CONS1 : F_CONS2);
            const uint32_t last_needed = total_rows > static_cast<uint32_t>(R) ? total_rows - R : 0;   // the last store waits for this
            uint32_t have_credit = 0;
            const long long t0 = clock64();
            while (have_credit < last_needed) {
                uint32_t v = ld_relaxed(cons);
                if (conv == 0) v = min(v, ld_relaxed(cons + FLAG_STRIDE));
                if (v > have_credit) {
                    have_credit = v;
                    *credit_gen = v;
                }
                if (clock64() - t0 > 40 * HRN_WAIT_LIMIT_CYCLES) {
                    printf("hrn_b200: wavefront credit poller timed out (block %d: have %u, want %u)\n", (int)blockIdx.x, have_credit, last_needed);
                    __trap();
                }
            }
        }
    } else if (warp == 3) {
        // ===================================================== publisher: hand-over counters for the neighbours in the stream.
        // Off the epilogue warps on purpose: a device-scope release (MEMBAR.GPU) costs about a microsecond under load, and
        // an epilogue warp that pays it every row holds back the accumulator ring and with it the whole CTA.
        if (ptx::elect_one()) {
            WaveWalker walk(w, stream, halo, false);
            Strip s;
            uint32_t k = 0, in_base = 0, next_pub = 1, total_rows = 0;
            {
                WaveWalker count(w, stream, halo, false);
                for (bool have = count.next(s); have; have = count.next(s)) total_rows += s.rows;
            }
            uint32_t* prod = flags + (conv == 0 ? F_PROD1 : F_PROD2) + part * FLAG_STRIDE;
            uint32_t* my_cons = flags + (conv == 1 ? F_CONS1 + part * FLAG_STRIDE : F_CONS2);
            for (bool have = walk.next(s); have; have = walk.next(s)) {
                const int in_lo = max(0, s.y0 - 1), in_hi = min(w.H, s.y0 + s.rows + 1);
                for (int i = 0; i < s.rows; ++i, ++k) {
                    WAVE_STAT_BEGIN(t5);
                    ptx::mbar_wait(bar_rowdone + 8 * (k % ROWDONE), (k / ROWDONE) & 1, 8);   // all eight epilogue warps are done with row k
                    WAVE_STAT_END(t5, 5);
                    WAVE_STAT_BEGIN(t6);
                    // rows of this strip that have been completed in the meantime are published with the same release: one
                    // device-scope fence per batch keeps the publisher from becoming the pace setter of the stream
                    while (i + 1 < s.rows && ptx::mbar_test_wait(bar_rowdone + 8 * ((k + 1) % ROWDONE), ((k + 1) / ROWDONE) & 1)) {
                        ++i;
                        ++k;
                    }
                    // conv B, C: the tensor pipe has finished every input row up to y + 1 of this strip (its accumulator was
                    // complete before the epilogue ran), so those ring rows may be overwritten.  Needs no fence, and goes out
                    // before the release below so that the credit does not wait for it.
                    if (conv > 0) st_relaxed(my_cons, in_base + static_cast<uint32_t>(min(s.y0 + i + 2, in_hi) - in_lo));
                    // conv A, B: rows 0 .. k are in the ring.  One device-scope release per `publish_rows` rows (and at the end
                    // of the stream): the fence drains the SM's store path, which every epilogue warp shares.
                    if (conv < 2 && (k + 1 >= next_pub || k + 1 == total_rows)) {
                        fence_proxy_async_global();      // the rows will be read through the async proxy (TMA)
                        if (w.debug_flags & 64) st_relaxed(prod, k + 1);
                        else st_release(prod, k + 1);
                        next_pub = k + 1 + w.publish_rows;
                    }
                    WAVE_STAT_END(t6, 6);
                }
                in_base += static_cast<uint32_t>(in_hi - in_lo);
            }
        }
    } else if (warp >= 4) {
        // ===================================================== epilogue: 8 warps, (lane quadrant) x (column half)
        const int wq = warp & 3;                 // TMEM lanes [32 wq, 32 wq + 32)
        const int hf = (warp - 4) >> 2;          // accumulator columns [32 hf, 32 hf + 32)
        const int co0 = part * NT + hf * 32;     // first output channel handled by this thread
        float bias_r[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) bias_r[e] = bias_s[hf * 32 + e];
        const bool has_prelu = cv.has_prelu != 0;
        const float slope_m1 = cv.prelu - 1.0f;  // PReLU(v) = v + (slope - 1) * min(v, 0)
        const int out_c = conv < 2 ? 128 : 64;
        __nv_bfloat16* ring_out = conv == 0 ? w.ring1 : w.ring2;
        const uint32_t* cons = flags + (conv == 0 ? F_CONS1 : F_CONS2);      // the consumer's "input rows read" counter(s)
        uint32_t seen_c0 = 0, seen_c1 = 0;
        WaveWalker walk(w, stream, halo);
        Strip s;
        bool have = walk.next(s);
        ptx::pdl_wait();                         // residual reads and output writes touch the previous kernel's tensors
        uint32_t tile = 0;                       // rows produced so far = ring row counter of a producer
        const int x = wq * 32 + lane;
        const bool valid = x < w.W;
        for (; have; have = walk.next(s)) {
            const int b = s.m / w.half, ip = s.m % w.half;
            const __nv_bfloat16* res_img = nullptr;
            float scale = 1.0f;
            if (conv == 1) {                     // ResidualBlock skip onto cat(alice, bob): this CTA's half is one of the two views
                const int img = b * w.src_views + (part ? (w.top - 1 - ip) : ip);
                res_img = w.stack_in + (static_cast<size_t>(img) * w.H * w.W) * 64 + hf * 32;
            } else if (conv == 2 && w.alpha_residual) {
                res_img = w.stack_in + (static_cast<size_t>(b * w.src_views + ip) * w.H * w.W) * 64 + hf * 32;
                scale = w.alphas[b * w.alpha_stride + (w.top - 1 - ip)];
            }
            const bool use_res = res_img != nullptr && valid && !(w.debug_flags & 8);
            const __nv_bfloat162 scale2 = __floats2bfloat162_rn(scale, scale);
            const size_t pix0 = static_cast<size_t>(s.y0) * w.W + x;
            const __nv_bfloat16* rp = res_img + pix0 * 64;
            const size_t r_step = static_cast<size_t>(w.W) * 64;
            __nv_bfloat16* op = nullptr;
            if (conv == 2) op = w.stack_out + (static_cast<size_t>(b * w.src_views + ip) * w.H * w.W + pix0) * 64 + co0;
            // The residual row (skip connection / alice) is fetched one whole row ahead into the other register buffer.
            // Measured alternative (not kept): the row by TMA into a shared-memory slot (a quarter of the load/store-unit
            // wavefronts: a warp's 256-bit global load touches 32 different lines) -- single-buffered in the 16 KB that are
            // left, it arrives too late to help (4.55 vs 4.53 ms), and it needs fence.proxy.async.shared::cta before every
            // refill: the epilogue's ld.shared (generic proxy) and the TMA write (async proxy) of the same slot are NOT ordered
            // by the mbarrier hand-shake alone (without the fence one forward in ten came back with rows off by 1e-3).
            // conv C's alice row and conv B's bob row are read for the last time here: first in line for eviction from L2
            // (conv B's alice row comes back once more in conv C)
            const bool last_use = !(w.debug_flags & 2048) && (conv == 2 || (conv == 1 && part == 1));
            auto res_load = [&](const __nv_bfloat16* p, uint32_t (&r)[2][8]) {
                if (last_use) {
                    ptx::ldg_nc_v8_last_use(p, r[0]);
                    ptx::ldg_nc_v8_last_use(p + 16, r[1]);
                } else {
                    ptx::ldg_nc_v8(p, r[0]);
                    ptx::ldg_nc_v8(p + 16, r[1]);
                }
            };
            uint32_t rv_a[2][8], rv_b[2][8];
            if (use_res) res_load(rp, rv_a);
            auto do_row = [&](int i, uint32_t (&rv)[2][8], uint32_t (&rv_next)[2][8]) {
                const uint32_t acc = tile % ACC_SLOTS, aph = (tile / ACC_SLOTS) & 1;
                if (use_res && i + 1 < s.rows) res_load(rp + r_step, rv_next);
                {
                    WAVE_STAT_BEGIN(t4);
                    ptx::mbar_wait(bar_tfull + 8 * acc, aph, 5);
                    if (warp == 4 && lane == 0) WAVE_STAT_END(t4, 4);
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
                    float x0 = __uint_as_float(v[2 * e]) + bias_r[2 * e];
                    float x1 = __uint_as_float(v[2 * e + 1]) + bias_r[2 * e + 1];
                    if (has_prelu) {
                        x0 = fmaf(slope_m1, fminf(x0, 0.0f), x0);
                        x1 = fmaf(slope_m1, fminf(x1, 0.0f), x1);
                    }
                    __nv_bfloat162 yv = __floats2bfloat162_rn(x0, x1);
                    if (use_res) yv = __hfma2(scale2, yv, *reinterpret_cast<const __nv_bfloat162*>(&rv[e >> 3][e & 7]));
                    o[e >> 3][e & 7] = *reinterpret_cast<const uint32_t*>(&yv);
                }
                if (conv < 2) {
                    // ring slot of produced row `tile`; it may be overwritten once the consumer(s) have read row tile - R
                    if (tile >= static_cast<uint32_t>(R) && !(w.debug_flags & 32)) {
                        const uint32_t want = tile - R + 1;
                        if (!(w.debug_flags & 1024)) {
                            if (seen_c0 < want) {            // shared-memory mirror kept by the credit poller (warp 2)
                                WAVE_STAT_BEGIN(t3);
                                const long long t0 = clock64();
                                while ((seen_c0 = *credit_gen) < want) {
                                    if (clock64() - t0 > HRN_WAIT_LIMIT_CYCLES) {
                                        printf("hrn_b200: wavefront ring-space wait timed out (block %d warp %d: have %u, want %u)\n",
                                               (int)blockIdx.x, warp, seen_c0, want);
                                        __trap();
                                    }
                                }
                                if (warp == 4 && lane == 0) WAVE_STAT_END(t3, 3);
                            }
                        } else if (seen_c0 < want || (conv == 0 && seen_c1 < want)) {
                            if (lane == 0) {
                                WAVE_STAT_BEGIN(t3);
                                if (seen_c0 < want) seen_c0 = flag_wait(cons, want, 30);
                                if (conv == 0 && seen_c1 < want) seen_c1 = flag_wait(cons + FLAG_STRIDE, want, 31);
                                if (warp == 4) WAVE_STAT_END(t3, 3);
                            }
                            seen_c0 = __shfl_sync(0xffffffffu, seen_c0, 0);
                            seen_c1 = __shfl_sync(0xffffffffu, seen_c1, 0);
                        }
                    }
                    if (valid) {
                        __nv_bfloat16* dst = ring_out + ((static_cast<size_t>(stream) * R + tile % R) * w.W + x) * out_c + co0;
                        if (w.debug_flags & 2048) {
                            ptx::stg_v8(dst, o[0]);
                            ptx::stg_v8(dst + 16, o[1]);
                        } else {                     // ring rows are consumed from L2 and overwritten 16 rows later: never worth a write-back
                            ptx::stg_v8_keep(dst, o[0]);
                            ptx::stg_v8_keep(dst + 16, o[1]);
                        }
                    }
                } else {
                    if (valid) {
                        if (w.debug_flags & 2048) {
                            ptx::stg_v8(op, o[0]);
                            ptx::stg_v8(op + 16, o[1]);
                        } else {                     // the level's output: read by the next launch only
                            ptx::stg_v8_stream(op, o[0]);
                            ptx::stg_v8_stream(op + 16, o[1]);
                        }
                    }
                    op += static_cast<size_t>(w.W) * 64;
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar_rowdone + 8 * (tile % ROWDONE));      // -> publisher (warp 3)
                ++tile;
                rp += r_step;
            };
            for (int i = 0; i < s.rows; i += 2) {
                do_row(i, rv_a, rv_b);
                if (i + 1 < s.rows) do_row(i + 1, rv_b, rv_a);
            }
        }
        // ---- carried views: alice of a pair whose bob has alpha = 0 goes to the next level unchanged (conv C CTAs only)
        if (conv == 2 && w.carry_count != nullptr) {
            const long long total = static_cast<long long>(*w.carry_count) * w.H;
            const long long r0 = total * stream / w.streams, r1 = total * (stream + 1) / w.streams;
            const int t = threadIdx.x - 128, row_vecs = w.W * 8;               // 16-byte vectors per 64-channel row
            for (long long r = r0; r < r1; ++r) {
                const int p = w.carry_list[r / w.H], y = static_cast<int>(r % w.H);
                const size_t off = ((static_cast<size_t>(p / w.half) * w.src_views + p % w.half) * w.H + y) * w.W * 64;
                const uint4* src = reinterpret_cast<const uint4*>(w.stack_in + off);
                uint4* dst = reinterpret_cast<uint4*>(w.stack_out + off);
                for (int v = t; v < row_vecs; v += EPI_WARPS * 32) dst[v] = __ldg(src + v);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (threadIdx.x == 0) WAVE_STAT_END(t_kernel, 0);
    if (warp == 2) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace

int fuse_wave_streams(int sm_count) { return sm_count / CTAS_PER_STREAM; }

// The CTAs of a wavefront grid wait for one another, so the whole grid has to be resident at once.  Asked once per device:
// how many CTAs of this kernel fit (one per SM with 214 KB of shared memory, fewer when the context is limited to a share of
// the SMs); the caller falls back to the three-launch schedule when the answer is smaller than the grid.
bool fuse_wave_fits(int sm_count) {
    static int max_ctas[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    static bool attr_set[64] = {};
    if (allow_dynamic_smem(fuse_wave_kernel, SMEM_BYTES, attr_set)) return false;
    std::lock_guard<std::mutex> lock(lazy_init_mutex());
    if (max_ctas[dev] == 0) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fuse_wave_kernel, NUM_THREADS, SMEM_BYTES) != cudaSuccess) {
            cudaGetLastError();
            per_sm = 0;
        }
        max_ctas[dev] = per_sm > 0 ? per_sm * sm_count : -1;
    }
    return max_ctas[dev] >= fuse_wave_streams(sm_count) * CTAS_PER_STREAM;
}
size_t fuse_wave_ring_bytes(int sm_count, int ring_rows, int W) {
    return static_cast<size_t>(fuse_wave_streams(sm_count)) * ring_rows * W * 128 * sizeof(__nv_bfloat16);
}
size_t fuse_wave_flag_bytes(int sm_count) { return static_cast<size_t>(fuse_wave_streams(sm_count)) * FLAGS_PER_STREAM * sizeof(uint32_t); }

int fuse_wave_launch(const FuseWaveArgs& a, int sm_count, cudaStream_t stream) {
    if (a.W > TILE_M || a.W <= 0 || a.H <= 0) {
        set_error("fuse_wave: images wider than %d pixels take the three-launch schedule", TILE_M);
        return -1;
    }
    int streams = fuse_wave_streams(sm_count);
    if (a.streams > 0 && a.streams < streams) streams = a.streams;
    if (streams < 1 || a.ring_rows < 8) {
        set_error("fuse_wave: needs at least %d SMs and 8 ring rows", CTAS_PER_STREAM);
        return -1;
    }
    WaveArgs w{};
    w.H = a.H;
    w.W = a.W;
    w.half = a.half;
    w.src_views = a.src_views;
    w.top = a.top;
    w.live_list = a.live_list;
    w.live_count = a.live_count;
    w.carry_list = a.carry_list;
    w.carry_count = a.carry_count;
    w.stack_in = a.stack_in;
    w.stack_out = a.stack_out;
    w.ring1 = a.ring1;
    w.ring2 = a.ring2;
    w.ring_rows = a.ring_rows;
    w.streams = streams;
    w.flags = a.flags;
    w.alphas = a.alphas;
    w.alpha_stride = a.alpha_stride;
    w.alpha_residual = a.alpha_residual;
    w.debug_flags = a.debug_flags;
    w.stats = a.stats;
    w.publish_rows = a.publish_rows > 0 ? a.publish_rows : 1;
    w.lag_rows = a.lag_rows > 0 ? a.lag_rows : 0;
    if (a.ring_rows < w.publish_rows + w.lag_rows + 4) {
        set_error("fuse_wave: ring_rows must be at least publish_rows + lag_rows + 4");
        return -1;
    }
    if (a.ring_rows < w.publish_rows + 4) {
        set_error("fuse_wave: ring_rows must be at least publish_rows + 4");
        return -1;
    }
    for (int i = 0; i < 3; ++i) {
        w.conv[i].w_img = a.w_img[i];
        w.conv[i].bias = a.bias[i];
        w.conv[i].prelu = a.prelu[i];
        w.conv[i].has_prelu = a.has_prelu[i];
    }
    CUtensorMap map_stack, map_r1, map_r2;
    if (encode_nhwc_map(&map_stack, a.stack_in, 64, a.W, a.H, a.stack_images, SLOT_PIX)) return -1;
    if (encode_nhwc_map(&map_r1, a.ring1, 128, a.W, a.ring_rows, streams, SLOT_PIX)) return -1;
    if (encode_nhwc_map(&map_r2, a.ring2, 128, a.W, a.ring_rows, streams, SLOT_PIX)) return -1;
    static bool attr_set[64] = {};
    if (allow_dynamic_smem(fuse_wave_kernel, SMEM_BYTES, attr_set)) return -1;
    HRN_CUDA_OK(launch_pdl(fuse_wave_kernel, streams * CTAS_PER_STREAM, NUM_THREADS, SMEM_BYTES, stream, 1, map_stack, map_r1,
                           map_r2, w));
    note_launches(1);
    return 0;
}

}  // namespace hrn
