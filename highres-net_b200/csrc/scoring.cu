// Scoring half of the hot path:
//   * lanczos_shift_kernel : separable 7-tap Lanczos sub-pixel shift with reflect padding
//     (lanczos.py:47-107 + lanczos_kernel lanczos.py:5-43); one launch for all images
//     instead of ~25 tiny launches per image.
//   * cpsnr_*              : brightness-bias-corrected clear-PSNR for all (2*border+1)^2
//     integer shifts of HR against the border-cropped SR (Evaluator.py:11-43, 52-73),
//     with the max AND the argmax (first maximum in row-major (x, y) order, np.argmax).
//
// cPSNR arithmetic mirrors the reference element-wise ops in fp32 (diff, diff*map,
// (diff-bias)*map, square); only the big sums are carried in fp64 and reduced in a
// fixed order, so results are deterministic and within ~1e-6 dB of the fp32 numpy loop.
#include "internal.h"

#include <mutex>

namespace hrn {
namespace {

constexpr int MAX_TAPS = 15;

// ------------------------------------------------------------------ Lanczos
constexpr int LZ_TW = 64, LZ_TH = 32, LZ_THREADS = 256;

// lanczos.py:26-41 in fp32: w(t) = sinc(pi t) sinc(pi t / a), pi t == 0 -> 1e-6, no |t| < a support clamp,
// normalised to sum 1.
__device__ __forceinline__ void lanczos_taps_device(float d, int a, int ntaps, float* taps) {
    const int half = ntaps / 2;
    float k[MAX_TAPS], sum = 0.0f;
    for (int i = 0; i < ntaps; ++i) {
        const float x = static_cast<float>(i - half) - d;
        float pix = 3.14159265358979323846f * x;
        pix = pix == 0.0f ? 1e-6f : pix;
        const float pa = pix / static_cast<float>(a);
        k[i] = (sinf(pix) / pix) * (sinf(pa) / pa);
        sum += k[i];
    }
    for (int i = 0; i < ntaps; ++i) taps[i] = k[i] / sum;
}

__global__ void lanczos_taps_kernel(const float* __restrict__ d, int n, int a, int ntaps, float* __restrict__ out) {
    // lets a kernel launched with programmatic stream serialisation (lanczos7_tma.cu) run its prologue underneath this one;
    // it still waits for this grid to finish before it reads the taps
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float t[MAX_TAPS];
    lanczos_taps_device(d[i], a, ntaps, t);
    for (int k = 0; k < ntaps; ++k) out[static_cast<size_t>(i) * ntaps + k] = t[k];
}

__device__ __forceinline__ int padded_index(int q, int size, int p, bool* ok) {
    // reflect (no edge repeat) inside the p-wide ReflectionPad2d ring, zero outside (conv2d zero padding)
    *ok = (q >= -p) && (q < size + p);
    int r = q < 0 ? -q : q;
    r = r >= size ? 2 * (size - 1) - r : r;
    return min(max(r, 0), size - 1);
}

__global__ void __launch_bounds__(LZ_THREADS)
lanczos_shift_kernel(const float* __restrict__ img, const float* __restrict__ shift, int C, int H, int W, int p,
                     int a, int ntaps, float* __restrict__ out) {
    extern __shared__ float lz_smem[];
    const int half = ntaps / 2;
    const int in_w = LZ_TW + 2 * half, in_h = LZ_TH + 2 * half;
    float* tile = lz_smem;                        // [in_h][in_w]
    float* tmp = tile + in_h * in_w;              // [LZ_TH][in_w]
    __shared__ float taps[2][MAX_TAPS];
    const int plane = blockIdx.z;                 // n * C + c
    const int c = plane % C;
    const int x0 = blockIdx.x * LZ_TW, y0 = blockIdx.y * LZ_TH;
    const float* src = img + static_cast<size_t>(plane) * H * W;

    if (threadIdx.x < 2) lanczos_taps_device(shift[c * 2 + threadIdx.x], a, ntaps, taps[threadIdx.x]);
    for (int i = threadIdx.x; i < in_h * in_w; i += LZ_THREADS) {
        const int r = i / in_w, q = i % in_w;
        bool oky, okx;
        const int yy = padded_index(y0 + r - half, H, p, &oky);
        const int xx = padded_index(x0 + q - half, W, p, &okx);
        tile[i] = (oky && okx) ? __ldg(src + static_cast<size_t>(yy) * W + xx) : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LZ_TH * in_w; i += LZ_THREADS) {      // y pass (dim 0 taps first, lanczos.py:90)
        const int r = i / in_w, q = i % in_w;
        float acc = 0.0f;
        for (int t = 0; t < ntaps; ++t) acc = fmaf(taps[0][t], tile[(r + t) * in_w + q], acc);
        tmp[i] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LZ_TH * LZ_TW; i += LZ_THREADS) {     // x pass (lanczos.py:94)
        const int r = i / LZ_TW, q = i % LZ_TW;
        const int y = y0 + r, x = x0 + q;
        if (y < H && x < W) {
            float acc = 0.0f;
            for (int t = 0; t < ntaps; ++t) acc = fmaf(taps[1][t], tmp[r * in_w + q + t], acc);
            out[static_cast<size_t>(plane) * H * W + static_cast<size_t>(y) * W + x] = acc;
        }
    }
}

// N = 7 specialisation (the only width the reference uses).  128 x 32 output tile per block.  The y pass runs
// straight from global memory: each thread owns one column of the (128 + 6)-wide strip and slides a 7-row register
// window down 16 output rows (22 coalesced loads for 16 outputs, nothing staged in shared memory); the x pass then
// computes four outputs per thread from three 128-bit shared-memory reads and stores one float4.
constexpr int L7_TW = 128, L7_TH = 32, L7_THREADS = 288, L7_HALF = 3;   // 288 >= 134 columns x 2 segments: one y-pass round
constexpr int L7_INW = L7_TW + 2 * L7_HALF, L7_PITCH = L7_INW + 2;   // 134, 136
constexpr int L7_SEG = 16;                                            // output rows per y-pass work item

__global__ void __launch_bounds__(L7_THREADS, 4)
lanczos_shift7_kernel(const float* __restrict__ img, const float* __restrict__ taps, int C, int H, int W, int p,
                      float* __restrict__ out) {
    __shared__ __align__(16) float tmp[L7_TH][L7_PITCH];
    const int plane = blockIdx.z, c = plane % C;
    const int x0 = blockIdx.x * L7_TW, y0 = blockIdx.y * L7_TH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* src = img + static_cast<size_t>(plane) * H * W;
    // taps[(c * 2 + axis) * 7 + t] come from lanczos_taps_kernel (one launch for all channels): computing them here cost
    // every block two serial sinf chains with every other thread waiting at a barrier (ncu: barrier stalls first)
    float ky[7], kx[7];
#pragma unroll
    for (int t = 0; t < 7; ++t) {
        ky[t] = __ldg(taps + (c * 2) * 7 + t);
        kx[t] = __ldg(taps + (c * 2 + 1) * 7 + t);
    }
    const bool y_interior = (y0 - L7_HALF >= 0) && (y0 + L7_TH + L7_HALF <= H);
    // y pass (dim 0 taps first, lanczos.py:90): work item = (column q of the strip, 16-row segment)
    for (int item = threadIdx.x; item < L7_INW * (L7_TH / L7_SEG); item += L7_THREADS) {
        const int q = item % L7_INW, seg = item / L7_INW;
        bool okx;
        const int xx = padded_index(x0 + q - L7_HALF, W, p, &okx);
        const int r0 = seg * L7_SEG;
        // all 22 loads are issued before the first use (memory-level parallelism), then 16 outputs are formed
        float col[L7_SEG + 6];
        if (y_interior) {                       // no reflection needed in y for this block (10 of 12 row tiles at 384)
            // xx is clamped into the image, so the loads are unconditional (a column outside the zero ring is cleared
            // afterwards) and the row offsets are 32-bit: a predicated load with a size_t product rebuilt the whole
            // 64-bit address per element, and address arithmetic was a third of all issued instructions.
            const float* base_ptr = src + static_cast<size_t>(y0 + r0 - L7_HALF) * W + xx;
#pragma unroll
            for (int t = 0; t < L7_SEG + 6; ++t) col[t] = __ldg(base_ptr + t * W);
            if (!okx) {
#pragma unroll
                for (int t = 0; t < L7_SEG + 6; ++t) col[t] = 0.0f;
            }
        } else {
#pragma unroll
            for (int t = 0; t < L7_SEG + 6; ++t) {
                bool oky;
                const int yy = padded_index(y0 + r0 + t - L7_HALF, H, p, &oky);
                col[t] = (oky && okx) ? __ldg(src + static_cast<size_t>(yy) * W + xx) : 0.0f;
            }
        }
#pragma unroll
        for (int r = 0; r < L7_SEG; ++r) {
            float acc = 0.0f;
#pragma unroll
            for (int t = 0; t < 7; ++t) acc = fmaf(ky[t], col[r + t], acc);
            tmp[r0 + r][q] = acc;
        }
    }
    __syncthreads();
    const bool vec_ok = (W & 3) == 0;
    for (int r = warp; r < L7_TH; r += L7_THREADS / 32) {            // x pass (lanczos.py:94), 4 outputs per thread
        const int y = y0 + r, x = x0 + 4 * lane;
        if (y >= H || x >= W) continue;
        float v[12];
#pragma unroll
        for (int k = 0; k < 3; ++k) *reinterpret_cast<float4*>(&v[4 * k]) = *reinterpret_cast<const float4*>(&tmp[r][4 * lane + 4 * k]);
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float acc = 0.0f;
#pragma unroll
            for (int t = 0; t < 7; ++t) acc = fmaf(kx[t], v[e + t], acc);
            o[e] = acc;
        }
        float* dst = out + static_cast<size_t>(plane) * H * W + static_cast<size_t>(y) * W + x;
        if (vec_ok && x + 3 < W) {
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (x + e < W) dst[e] = o[e];
        }
    }
}

// ------------------------------------------------------------------ cPSNR shift search
constexpr int CP_LANES = 32;                       // threadIdx.x
constexpr int CP_TCOLS = 4;                        // crop columns per thread
constexpr int CP_COLS = CP_LANES * CP_TCOLS;       // crop columns per block
constexpr int CP_MAXS = 7;                         // shifts per axis supported (border_w <= 3)
constexpr int CP_WIN = 12;                         // hr / map columns per thread and row: 4 + 7 - 1 = 10, as three float4
constexpr int CP_TARGET_BLOCKS = 148 * 8;

struct CpGeom {
    int H, W, border, S, size, col_blocks, band_rows, blocks_per_set, vec_ok;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// PASS 1: per site sum(m) and sum((hr - sr) * m).  PASS 2: per site sum(((hr - sr - bias) * m)^2).
// Block = (32 lanes x 4 crop columns) x (S row-shifts x).  A thread loads 10 consecutive hr / map values of one row (three
// aligned float4 each) and 4 sr values, and feeds 4 columns x S column-shifts y from them: 11 load instructions per
// 28 (pixel, site) terms instead of 15 per 7, which moves the kernel from the load pipe to the fp32 pipe.
// SFIX = 7 is the border_w = 3 search with every per-site loop bound known at compile time; SFIX = 0 takes S from g.
template <int PASS, int SFIX>
__global__ void __launch_bounds__(CP_LANES* CP_MAXS)
cpsnr_pass_kernel(const float* __restrict__ sr, const float* __restrict__ hr, const float* __restrict__ hm,
                  CpGeom g, int clip_sr, const float* __restrict__ bias, double* __restrict__ partial) {
    const int S = SFIX ? SFIX : g.S;
    const int set = blockIdx.y;
    const int band = blockIdx.x / g.col_blocks, cb = blockIdx.x % g.col_blocks;
    const int j0 = cb * CP_COLS + threadIdx.x * CP_TCOLS;      // first crop column of this thread
    const int ncol = min(CP_TCOLS, g.size - j0);               // <= 0: nothing to do
    const int x = threadIdx.y;                                 // row shift handled by this thread
    const size_t plane = static_cast<size_t>(g.H) * g.W;
    const float* srp = sr + set * plane;
    const float* hrp = hr + set * plane;
    const float* hmp = hm + set * plane;
    const int i0 = band * g.band_rows, i1 = min(g.size, i0 + g.band_rows);
    const bool vec = g.vec_ok && j0 + CP_WIN <= g.W;
    const bool full = ncol == CP_TCOLS;
    double s0[CP_MAXS], s1[CP_MAXS];
    float b[CP_MAXS];
#pragma unroll
    for (int y = 0; y < CP_MAXS; ++y) {
        s0[y] = 0.0;
        s1[y] = 0.0;
        b[y] = (PASS == 2 && y < S) ? bias[(set * S + x) * S + y] : 0.0f;
    }
    if (ncol > 0) {
        // Rows are taken four at a time: the per-element arithmetic is the reference's fp32 arithmetic, the 16 terms of a
        // site (4 rows x 4 columns) are added in fp32 (the reference sums everything in fp32) and only that partial
        // sum is folded into the fp64 accumulator.
        for (int i = i0; i < i1; i += 4) {
            float p0[CP_MAXS], p1[CP_MAXS];
#pragma unroll
            for (int y = 0; y < CP_MAXS; ++y) p0[y] = p1[y] = 0.0f;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (i + r < i1) {
                    const float* srow = srp + static_cast<size_t>(i + r + g.border) * g.W + j0 + g.border;
                    float sv[CP_TCOLS];
#pragma unroll
                    for (int c = 0; c < CP_TCOLS; ++c) {
                        sv[c] = c < ncol ? __ldg(srow + c) : 0.0f;
                        if (clip_sr) sv[c] = fminf(fmaxf(sv[c], 0.0f), 1.0f);
                    }
                    const size_t off = static_cast<size_t>(i + r + x) * g.W + j0;
                    float hw[CP_WIN], mw[CP_WIN];
                    if (vec) {
#pragma unroll
                        for (int q = 0; q < CP_WIN / 4; ++q) {
                            const float4 h4 = __ldg(reinterpret_cast<const float4*>(hrp + off) + q);
                            const float4 m4 = __ldg(reinterpret_cast<const float4*>(hmp + off) + q);
                            hw[4 * q] = h4.x, hw[4 * q + 1] = h4.y, hw[4 * q + 2] = h4.z, hw[4 * q + 3] = h4.w;
                            mw[4 * q] = m4.x, mw[4 * q + 1] = m4.y, mw[4 * q + 2] = m4.z, mw[4 * q + 3] = m4.w;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < CP_TCOLS + CP_MAXS - 1; ++k) {
                            const bool ok = j0 + k < g.W;
                            hw[k] = ok ? __ldg(hrp + off + k) : 0.0f;
                            mw[k] = ok ? __ldg(hmp + off + k) : 0.0f;
                        }
                    }
#pragma unroll
                    for (int c = 0; c < CP_TCOLS; ++c) {
                        if (full || c < ncol) {
#pragma unroll
                            for (int y = 0; y < CP_MAXS; ++y) {
                                if (y < S) {
                                    const float m = mw[c + y];
                                    const float d = hw[c + y] - sv[c];     // diff = hr - sr            (Evaluator.py:35)
                                    if (PASS == 1) {
                                        p0[y] += m;                        // n_clear                   (Evaluator.py:34)
                                        p1[y] += d * m;                    // sum(diff * hr_map)        (Evaluator.py:36)
                                    } else {
                                        const float t = (d - b[y]) * m;    // (diff - bias) * hr_map    (Evaluator.py:37)
                                        p0[y] += t * t;
                                    }
                                }
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int y = 0; y < CP_MAXS; ++y) {
                s0[y] += static_cast<double>(p0[y]);
                if (PASS == 1) s1[y] += static_cast<double>(p1[y]);
            }
        }
    }
    // fixed-order block reduction over the 32 lanes (one warp per x), then one partial per block
    double* dst = partial + ((static_cast<size_t>(set) * g.blocks_per_set + blockIdx.x) * S + x) * S * 2;
#pragma unroll
    for (int y = 0; y < CP_MAXS; ++y) {
        if (y < S) {
            const double r0 = warp_sum(s0[y]);
            const double r1 = PASS == 1 ? warp_sum(s1[y]) : 0.0;
            if (threadIdx.x == 0) {
                dst[y * 2] = r0;
                dst[y * 2 + 1] = r1;
            }
        }
    }
}

// border_w = 3 search, second generation: ONE warp per block owns 128 crop columns x a band of rows and ALL 49 sites.
// A lane sweeps the hr / map rows of its band once (three aligned float4 each = 10 useful values for its 4 columns x 7
// column shifts) and keeps the last 7 sr rows of its 4 columns in registers: hr row h meets sr rows h - x, x = 0..6, so
// one row of loads (8 float4, issued one row ahead) feeds 196 (pixel, site) terms instead of 28.
// The 49 (or 2 x 49) per-lane accumulators are fp32 and live for at most CW_FLUSH rows (<= 64 terms); they are then
// staged through shared memory and folded into fp64 totals in a fixed order (lane l owns quantities l, l + 32, ...).
// Element-wise arithmetic is the reference's fp32 arithmetic, exactly as in cpsnr_pass_kernel.
constexpr int CW_COLS = 128, CW_FLUSH = 16, CW_S = 7, CW_SITES = CW_S * CW_S;
constexpr int CW_TARGET_WARPS = 148 * 11;      // one-warp blocks resident per wave (168 registers, 19 KB of shared memory)

struct CwRowS {                                // the same with the four sr values loaded as float, float2, float (cpsnr_window_kernel)
    float4 h[3], m[3];
    float s0;
    float2 s12;
    float s3;
};
struct CwRow {                                 // one hr / map row segment (12 columns) and the sr row that enters the window
    float4 h[3], m[3], s[2];
};

// All terms of one hr row: sites (x, y) with x in [xlo, xhi], columns c < ncol (CPRED) or all four.
template <int PASS, bool CPRED>
__device__ __forceinline__ void cw_terms(float (&acc)[CW_S][CW_S], const float (&aux)[CW_S][CW_S], float (&rs)[CW_S],
                                         const float (&hw)[12], const float (&mw)[12], const float (&svw)[CW_S][4],
                                         int xlo, int xhi, int ncol) {
    if (PASS == 1) {                           // rs = sum of the map over this lane's columns, per column shift
#pragma unroll
        for (int y = 0; y < CW_S; ++y) {
            float r = 0.0f;
#pragma unroll
            for (int c = 0; c < 4; ++c) r += (!CPRED || c < ncol) ? mw[c + y] : 0.0f;      // n_clear (Evaluator.py:34)
            rs[y] = r;
        }
    }
#pragma unroll
    for (int x = 0; x < CW_S; ++x) {
        if (x < xlo || x > xhi) continue;      // warp-uniform: sr crop row h - x lies outside this band
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (CPRED && c >= ncol) continue;
#pragma unroll
            for (int y = 0; y < CW_S; ++y) {
                const float m = mw[c + y];
                const float d = hw[c + y] - svw[x][c];               // diff = hr - sr            (Evaluator.py:35)
                if (PASS == 1) {
                    acc[x][y] += d * m;                              // sum(diff * hr_map)        (Evaluator.py:36)
                } else {
                    const float t = (d - aux[x][y]) * m;             // (diff - bias) * hr_map    (Evaluator.py:37)
                    acc[x][y] += t * t;
                }
            }
        }
    }
}

// Out of line on purpose: inlined, its fp64 temporaries pushed the row loop of the kernel below over its register budget.
static __device__ __noinline__ void cw_fallback_bias(const double* __restrict__ pass1, int set, int blocks_per_set, float* bias_sh) {
    for (int site = threadIdx.x; site < CW_SITES; site += 32) {
        double a0 = 0.0, a1 = 0.0;
        const double* src = pass1 + (static_cast<size_t>(set) * blocks_per_set * CW_SITES + site) * 2;
        for (int blk = 0; blk < blocks_per_set; ++blk) {
            a0 += src[static_cast<size_t>(blk) * CW_SITES * 2];
            a1 += src[static_cast<size_t>(blk) * CW_SITES * 2 + 1];
        }
        bias_sh[site] = static_cast<float>(a1 / a0);                  // 0/0 -> NaN like numpy
    }
    __syncwarp();
}

template <int PASS, bool FALLBACK = false>       // FALLBACK: the whole-imageset fallback of the one-pass path (set_sel, pass1)
__global__ void __launch_bounds__(32, 9)
cpsnr_window_kernel(const float* __restrict__ sr, const float* __restrict__ hr, const float* __restrict__ hm, CpGeom g,
                    int clip_sr, const float* __restrict__ bias, double* __restrict__ partial,
                    const uint8_t* __restrict__ set_sel = nullptr, const double* __restrict__ pass1 = nullptr) {
    if (FALLBACK && !set_sel[blockIdx.y]) return;                // selected imagesets only
    // PASS 2 of the fallback gets the pass-1 partial sums instead of a bias array and forms the bias itself, exactly like
    // cpsnr_finalize_kernel<1> (same order of additions), which saves a launch that nearly always has nothing to do
    __shared__ float bias_sh[PASS == 2 && FALLBACK ? CW_SITES : 1];
    if (PASS == 2 && FALLBACK) cw_fallback_bias(pass1, blockIdx.y, g.blocks_per_set, bias_sh);
    constexpr int NQ = PASS == 1 ? 2 * CW_SITES : CW_SITES;      // quantities per block: [site][n, sum d*m] or [site]
    constexpr int NK = (NQ + 31) / 32;
    __shared__ float stage[NQ][33];
    // PASS 1, n_clear: a row whose 7 sr partners all lie in the band adds rs[y] to every site (x, y), so those rows go
    // into 7 registers; the few rows at the top and bottom of a band go to the per-site array in shared memory.
    __shared__ float edge_n[PASS == 1 ? CW_SITES : 1][32];
    const int lane = threadIdx.x;
    const int set = blockIdx.y;
    const int band = blockIdx.x / g.col_blocks, cb = blockIdx.x % g.col_blocks;
    const int j0 = cb * CW_COLS + lane * 4;                      // first crop column of this lane
    const int ncol = min(4, g.size - j0);                        // <= 0: idle lane
    const bool all_full = __all_sync(0xffffffffu, ncol == 4);
    const size_t plane = static_cast<size_t>(g.H) * g.W;
    const float* srp = sr + set * plane + static_cast<size_t>(g.border) * g.W + j0;
    const float* hrp = hr + set * plane + j0;
    const float* hmp = hm + set * plane + j0;
    const int i0 = band * g.band_rows, i1 = min(g.size, i0 + g.band_rows);
    const int h_end = i1 + CW_S - 1;                             // hr rows [i0, h_end) meet sr crop rows [i0, i1)
    const bool q_ok[3] = {j0 + 4 <= g.W, j0 + 8 <= g.W, j0 + 12 <= g.W};

    float acc[CW_S][CW_S], aux[CW_S][CW_S];     // PASS 1: acc = sum d*m (aux unused).  PASS 2: acc = sum t*t, aux = bias
    float n_all[CW_S];                          // PASS 1: sum of rs over the rows that count for every x
#pragma unroll
    for (int x = 0; x < CW_S; ++x) {
        n_all[x] = 0.0f;
#pragma unroll
        for (int y = 0; y < CW_S; ++y) {
            acc[x][y] = 0.0f;
            aux[x][y] = PASS == 2 ? (FALLBACK ? bias_sh[x * CW_S + y] : bias[(set * CW_S + x) * CW_S + y]) : 0.0f;
            if (PASS == 1) edge_n[x * CW_S + y][lane] = 0.0f;
        }
    }
    double tot[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) tot[k] = 0.0;

    auto flush = [&]() {
#pragma unroll
        for (int x = 0; x < CW_S; ++x)
#pragma unroll
            for (int y = 0; y < CW_S; ++y) {
                if (PASS == 1) {
                    stage[(x * CW_S + y) * 2][lane] = n_all[y] + edge_n[x * CW_S + y][lane];
                    stage[(x * CW_S + y) * 2 + 1][lane] = acc[x][y];
                    edge_n[x * CW_S + y][lane] = 0.0f;
                } else {
                    stage[x * CW_S + y][lane] = acc[x][y];
                }
                acc[x][y] = 0.0f;
            }
        if (PASS == 1) {
#pragma unroll
            for (int y = 0; y < CW_S; ++y) n_all[y] = 0.0f;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const int q = lane + 32 * k;
            if (q < NQ) {
                double a = 0.0;
#pragma unroll 8
                for (int l = 0; l < 32; ++l) a += static_cast<double>(stage[q][l]);
                tot[k] += a;
            }
        }
        __syncwarp();
    };

    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load_row = [&](int h, CwRowS& r) {
        const size_t off = static_cast<size_t>(h) * g.W;
        const bool in = h < h_end;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            r.h[q] = (in && q_ok[q]) ? __ldg(reinterpret_cast<const float4*>(hrp + off) + q) : zero4;
            r.m[q] = (in && q_ok[q]) ? __ldg(reinterpret_cast<const float4*>(hmp + off) + q) : zero4;
        }
        const bool sin = h < i1;                                     // sr crop row h exists in this band
        // crop column j0 + c is image column j0 + c + 3: one float, one float2, one float.  Two aligned 128-bit loads would
        // be fewer instructions, but the unused fourth float of the second quad got its register reused as scratch a few
        // instructions after the load was issued -- a write-after-write hazard on a load in flight, the whole memory latency
        // once per row (ncu: 22 % of all stall samples sat on that one MOV).
        r.s0 = (sin && q_ok[0]) ? __ldg(srp + off + 3) : 0.0f;
        r.s12 = (sin && q_ok[1]) ? __ldg(reinterpret_cast<const float2*>(srp + off + 4)) : make_float2(0.f, 0.f);
        r.s3 = (sin && q_ok[1]) ? __ldg(srp + off + 6) : 0.0f;
    };

    float svw[CW_S][4];                                              // svw[x] = sr crop row h - x (columns j0 .. j0 + 3)
#pragma unroll
    for (int x = 0; x < CW_S; ++x)
#pragma unroll
        for (int c = 0; c < 4; ++c) svw[x][c] = 0.0f;
    CwRowS nxt;
    load_row(i0, nxt);
    int since_flush = 0;
    for (int h = i0; h < h_end; ++h) {
        const CwRowS cur = nxt;
        load_row(h + 1, nxt);                                        // in flight while this row is being consumed
        float hw[12], mw[12];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            hw[4 * q] = cur.h[q].x, hw[4 * q + 1] = cur.h[q].y, hw[4 * q + 2] = cur.h[q].z, hw[4 * q + 3] = cur.h[q].w;
            mw[4 * q] = cur.m[q].x, mw[4 * q + 1] = cur.m[q].y, mw[4 * q + 2] = cur.m[q].z, mw[4 * q + 3] = cur.m[q].w;
        }
#pragma unroll
        for (int x = CW_S - 1; x > 0; --x)
#pragma unroll
            for (int c = 0; c < 4; ++c) svw[x][c] = svw[x - 1][c];
        svw[0][0] = cur.s0, svw[0][1] = cur.s12.x, svw[0][2] = cur.s12.y, svw[0][3] = cur.s3;   // crop column j0 + c = image column j0 + c + 3
        if (clip_sr) {
#pragma unroll
            for (int c = 0; c < 4; ++c) svw[0][c] = fminf(fmaxf(svw[0][c], 0.0f), 1.0f);
        }
        // sr crop rows i = h - x that lie in this band: x in [xlo, xhi]
        const int xlo = max(0, h - i1 + 1), xhi = min(CW_S - 1, h - i0);
        float rs[CW_S];
        if (all_full)
            cw_terms<PASS, false>(acc, aux, rs, hw, mw, svw, xlo, xhi, 4);
        else
            cw_terms<PASS, true>(acc, aux, rs, hw, mw, svw, xlo, xhi, ncol);
        if (PASS == 1) {
            if (xlo == 0 && xhi == CW_S - 1) {
#pragma unroll
                for (int y = 0; y < CW_S; ++y) n_all[y] += rs[y];
            } else {
                for (int x = xlo; x <= xhi; ++x)
#pragma unroll
                    for (int y = 0; y < CW_S; ++y) edge_n[x * CW_S + y][lane] += rs[y];
            }
        }
        if (++since_flush == CW_FLUSH) {
            flush();
            since_flush = 0;
        }
    }
    if (since_flush > 0) flush();
    double* dst = partial + (static_cast<size_t>(set) * g.blocks_per_set + blockIdx.x) * CW_SITES * 2;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int q = lane + 32 * k;
        if (q < NQ) {
            if (PASS == 1) {
                dst[q] = tot[k];
            } else {
                dst[2 * q] = tot[k];
                dst[2 * q + 1] = 0.0;
            }
        }
    }
}

// ------------------------------------------------------------------ 49-site window kernel, third generation
// Same sweep and the same fp32 element arithmetic in the same order as cpsnr_window_kernel (so the site tables are
// bit-identical for 0/1 status maps), with two changes that attack what ncu showed -- an fp32-issue-bound kernel at 12
// warps per SM:
//   * the 7 row shifts are split over TWO one-warp blocks (x in [0, 4) and x in [4, 7), blockIdx.z): 16 / 12 instead of 49
//     accumulator pairs per lane, 14-16 warps per SM instead of 12, and twice as many blocks for small batches;
//   * Blackwell's packed fp32 pipe: the column shifts are processed in pairs (y, y + 1) with add/mul/fma.f32x2 (FADD2 /
//     FMUL2 / FFMA2 -- per-lane IEEE results identical to the scalar instructions), the sr value enters as a broadcast
//     operand.  y = 7 is a dummy lane of the last pair (its sums are dropped).  3 (pass 1) / 4 (pass 2) issue slots per
//     TWO (pixel, site) terms instead of 2 / 4 per term.
using f2 = unsigned long long;
__device__ __forceinline__ f2 pk(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// All terms of one hr row for the shifts xl in [xlo, xhi] (local to the block) and the columns c < ncol (CPRED) or all four.
// A pair of adjacent hr / map columns (k, k + 1) is a 64-bit register pair only for even k (the halves of the loaded
// float4s).  Even columns c therefore pair the shifts (y, y + 1) = (0,1) (2,3) (4,5) (6,7*) and odd columns pair
// (-1*,0) (1,2) (3,4) (5,6): k = c + y is even in both cases (* = dummy lane, dropped at the flush), and the two kinds of
// columns keep their own accumulators (acc[x][c & 1][..]), added when the block flushes.
template <int PASS, int XN, bool CPRED>
__device__ __forceinline__ void cw2_terms(f2 (&acc)[XN][2][4], f2 (*aux)[2][4], const float (&hw)[12], const float (&mw)[12],
                                          const float (&svw)[XN][4], int xlo, int xhi, int ncol) {
#pragma unroll
    for (int x = 0; x < XN; ++x) {
        if (x < xlo || x > xhi) continue;                            // warp-uniform: sr crop row h - x lies outside this band
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (CPRED && c >= ncol) continue;
            const f2 s2 = pk(svw[x][c], svw[x][c]);
#pragma unroll
            for (int yp = 0; yp < 4; ++yp) {
                const int k = (c & ~1) + 2 * yp;                     // even: columns (k, k + 1) <-> y = k - c, k + 1 - c
                const f2 d2 = sub2(pk(hw[k], hw[k + 1]), s2);        // diff = hr - sr            (Evaluator.py:35)
                const f2 m2 = pk(mw[k], mw[k + 1]);
                if (PASS == 1) {
                    acc[x][c & 1][yp] = fma2(d2, m2, acc[x][c & 1][yp]);      // sum(diff * hr_map)   (Evaluator.py:36)
                } else {
                    const f2 t2 = mul2(sub2(d2, aux[x][c & 1][yp]), m2);      // (diff - bias) * hr_map (Evaluator.py:37)
                    acc[x][c & 1][yp] = fma2(t2, t2, acc[x][c & 1][yp]);
                }
            }
        }
    }
}

// The same terms with scalar fp32 instructions (the arithmetic and sum order of cpsnr_window_kernel, for the shifts of one
// half-block): with the packed variant this separates the two changes of the third generation -- more warps per SM from the
// split, fewer instructions from the packing.
template <int PASS, int XN, bool CPRED>
__device__ __forceinline__ void cw2_terms_scalar(float (&acc)[XN][CW_S], const float* bias_s, const float (&hw)[12],
                                                 const float (&mw)[12], const float (&svw)[XN][4], int xlo, int xhi, int ncol) {
#pragma unroll
    for (int x = 0; x < XN; ++x) {
        if (x < xlo || x > xhi) continue;
#pragma unroll
        for (int y = 0; y < CW_S; ++y) {
            const float b = PASS == 2 ? bias_s[x * 8 + y] : 0.0f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (CPRED && c >= ncol) continue;
                const float m = mw[c + y];
                const float d = hw[c + y] - svw[x][c];               // diff = hr - sr            (Evaluator.py:35)
                if (PASS == 1) {
                    acc[x][y] += d * m;                              // sum(diff * hr_map)        (Evaluator.py:36)
                } else {
                    const float t = (d - b) * m;                     // (diff - bias) * hr_map    (Evaluator.py:37)
                    acc[x][y] += t * t;
                }
            }
        }
    }
}

template <int PASS, int X0, int XN, bool PACKED>
__device__ __forceinline__ void cpsnr_window2_body(const float* __restrict__ sr, const float* __restrict__ hr,
                                                   const float* __restrict__ hm, const CpGeom& g, int clip_sr,
                                                   const float* __restrict__ bias, double* __restrict__ partial,
                                                   float (*stage)[33], float (*edge_n)[32], f2 (*aux)[2][4]) {
    constexpr int NS = XN * CW_S;                                 // sites of this block: x in [X0, X0 + XN), all y
    constexpr int NQ = PASS == 1 ? 2 * NS : NS;
    constexpr int NK = (NQ + 31) / 32;
    const int lane = threadIdx.x;
    const int set = blockIdx.y;
    const int band = blockIdx.x / g.col_blocks, cb = blockIdx.x % g.col_blocks;
    const int j0 = cb * CW_COLS + lane * 4;                      // first crop column of this lane
    const int ncol = min(4, g.size - j0);                        // <= 0: idle lane
    const bool all_full = __all_sync(0xffffffffu, ncol == 4);
    const size_t plane = static_cast<size_t>(g.H) * g.W;
    const float* srp = sr + set * plane + static_cast<size_t>(g.border) * g.W + j0;
    const float* hrp = hr + set * plane + j0;
    const float* hmp = hm + set * plane + j0;
    const int i0 = band * g.band_rows, i1 = min(g.size, i0 + g.band_rows);
    const int h_begin = i0 + X0, h_end = i1 + X0 + XN - 1;       // hr rows that meet sr crop rows [i0, i1) at these shifts
    const bool q_ok[3] = {j0 + 4 <= g.W, j0 + 8 <= g.W, j0 + 12 <= g.W};

    // acc[x - X0][0][yp] = shifts (2 yp, 2 yp + 1) summed over the even columns, acc[x - X0][1][yp] = shifts (2 yp - 1, 2 yp)
    // over the odd columns (see cw2_terms).  PASS 2: the sites' biases sit in shared memory as the same pairs (broadcast
    // LDS.64) instead of 32 more registers per lane.
    f2 acc[XN][2][4];
    float accs[XN][CW_S];                        // !PACKED: plain fp32 accumulators, one per (x, y)
    float* bias_s = reinterpret_cast<float*>(aux);   // !PACKED, PASS 2: bias of site (x, y) at [x * 8 + y]
    float n_all[CW_S];                           // PASS 1: sum of rs over the rows that count for every x of this block
#pragma unroll
    for (int y = 0; y < CW_S; ++y) n_all[y] = 0.0f;
#pragma unroll
    for (int x = 0; x < XN; ++x) {
#pragma unroll
        for (int yp = 0; yp < 4; ++yp) acc[x][0][yp] = acc[x][1][yp] = pk(0.0f, 0.0f);
#pragma unroll
        for (int y = 0; y < CW_S; ++y) accs[x][y] = 0.0f;
    }
    if (PASS == 2 && !PACKED) {
        if (lane < XN * 8) bias_s[lane] = (lane & 7) < CW_S ? bias[(set * CW_S + X0 + (lane >> 3)) * CW_S + (lane & 7)] : 0.0f;
        __syncwarp();
    }
    if (PASS == 2 && PACKED) {
        if (lane < XN * 8) {
            const int x = lane >> 3, odd = (lane >> 2) & 1, yp = lane & 3;
            const int y0 = 2 * yp - odd, y1 = y0 + 1;                // the two shifts of this pair; -1 and 7 are dummies
            const float* b = bias + (set * CW_S + X0 + x) * CW_S;
            aux[x][odd][yp] = pk(y0 >= 0 ? b[y0] : 0.0f, y1 < CW_S ? b[y1] : 0.0f);
        }
        __syncwarp();
    }
    if (PASS == 1) {
#pragma unroll
        for (int q = 0; q < NS; ++q) edge_n[q][lane] = 0.0f;
    }
    double tot[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) tot[k] = 0.0;

    auto flush = [&]() {
#pragma unroll
        for (int x = 0; x < XN; ++x) {
            float ev[8], od[8];                                      // ev[y] for y = 0..7, od[y + 1] for y = -1..6
#pragma unroll
            for (int yp = 0; yp < 4; ++yp) {
                upk(acc[x][0][yp], ev[2 * yp], ev[2 * yp + 1]);
                upk(acc[x][1][yp], od[2 * yp], od[2 * yp + 1]);
                acc[x][0][yp] = acc[x][1][yp] = pk(0.0f, 0.0f);
            }
#pragma unroll
            for (int y = 0; y < CW_S; ++y) {
                const int q = x * CW_S + y;
                const float a = PACKED ? ev[y] + od[y + 1] : accs[x][y];     // even columns + odd columns
                accs[x][y] = 0.0f;
                if (PASS == 1) {
                    stage[q * 2][lane] = n_all[y] + edge_n[q][lane];
                    stage[q * 2 + 1][lane] = a;
                    edge_n[q][lane] = 0.0f;
                } else {
                    stage[q][lane] = a;
                }
            }
        }
        if (PASS == 1) {
#pragma unroll
            for (int y = 0; y < CW_S; ++y) n_all[y] = 0.0f;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const int q = lane + 32 * k;
            if (q < NQ) {
                double a = 0.0;
#pragma unroll 8
                for (int l = 0; l < 32; ++l) a += static_cast<double>(stage[q][l]);
                tot[k] += a;
            }
        }
        __syncwarp();
    };

    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load_row = [&](int h, CwRow& r) {
        const size_t off = static_cast<size_t>(h) * g.W;
        const bool in = h < h_end;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            r.h[q] = (in && q_ok[q]) ? __ldg(reinterpret_cast<const float4*>(hrp + off) + q) : zero4;
            r.m[q] = (in && q_ok[q]) ? __ldg(reinterpret_cast<const float4*>(hmp + off) + q) : zero4;
        }
        const int i = h - X0;                                        // the sr crop row that enters the window with hr row h
        const bool sin = i < i1;                                     // (i >= i0 by the loop bounds)
        const size_t soff = static_cast<size_t>(i) * g.W;
#pragma unroll
        for (int q = 0; q < 2; ++q) r.s[q] = (sin && q_ok[q]) ? __ldg(reinterpret_cast<const float4*>(srp + soff) + q) : zero4;
    };

    float svw[XN][4];                                                // svw[x - X0] = sr crop row h - x (columns j0 .. j0 + 3)
#pragma unroll
    for (int x = 0; x < XN; ++x)
#pragma unroll
        for (int c = 0; c < 4; ++c) svw[x][c] = 0.0f;
    int since_flush = 0;
    for (int h = h_begin; h < h_end; ++h) {
        // No software prefetch here: at 16 one-warp blocks per SM the other warps cover the L1 / L2 latency, and the 32
        // registers a prefetched row costs would spill.
        CwRow cur;
        load_row(h, cur);
        float hw[12], mw[12];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            hw[4 * q] = cur.h[q].x, hw[4 * q + 1] = cur.h[q].y, hw[4 * q + 2] = cur.h[q].z, hw[4 * q + 3] = cur.h[q].w;
            mw[4 * q] = cur.m[q].x, mw[4 * q + 1] = cur.m[q].y, mw[4 * q + 2] = cur.m[q].z, mw[4 * q + 3] = cur.m[q].w;
        }
#pragma unroll
        for (int x = XN - 1; x > 0; --x)
#pragma unroll
            for (int c = 0; c < 4; ++c) svw[x][c] = svw[x - 1][c];
        svw[0][0] = cur.s[0].w, svw[0][1] = cur.s[1].x, svw[0][2] = cur.s[1].y, svw[0][3] = cur.s[1].z;   // crop column j0 + c = image column j0 + c + 3
        if (clip_sr) {
#pragma unroll
            for (int c = 0; c < 4; ++c) svw[0][c] = fminf(fmaxf(svw[0][c], 0.0f), 1.0f);
        }
        // sr crop rows i = h - x of this band: x in [xlo, xhi] (clamped to this block's shifts)
        const int xlo = max(X0, h - i1 + 1), xhi = min(X0 + XN - 1, h - i0);
        if (PASS == 1) {
            float rs[CW_S];                                          // sum of the map over this lane's columns, per column shift
#pragma unroll
            for (int y = 0; y < CW_S; ++y) {
                float r = 0.0f;
#pragma unroll
                for (int c = 0; c < 4; ++c) r += (all_full || c < ncol) ? mw[c + y] : 0.0f;      // n_clear (Evaluator.py:34)
                rs[y] = r;
            }
            if (xlo == X0 && xhi == X0 + XN - 1) {
#pragma unroll
                for (int y = 0; y < CW_S; ++y) n_all[y] += rs[y];
            } else {
                for (int x = xlo; x <= xhi; ++x)
#pragma unroll
                    for (int y = 0; y < CW_S; ++y) edge_n[(x - X0) * CW_S + y][lane] += rs[y];
            }
        }
        if (PACKED) {
            if (all_full)
                cw2_terms<PASS, XN, false>(acc, aux, hw, mw, svw, xlo - X0, xhi - X0, 4);
            else
                cw2_terms<PASS, XN, true>(acc, aux, hw, mw, svw, xlo - X0, xhi - X0, ncol);
        } else {
            if (all_full)
                cw2_terms_scalar<PASS, XN, false>(accs, bias_s, hw, mw, svw, xlo - X0, xhi - X0, 4);
            else
                cw2_terms_scalar<PASS, XN, true>(accs, bias_s, hw, mw, svw, xlo - X0, xhi - X0, ncol);
        }
        if (++since_flush == CW_FLUSH) {
            flush();
            since_flush = 0;
        }
    }
    if (since_flush > 0) flush();
    double* dst = partial + ((static_cast<size_t>(set) * g.blocks_per_set + blockIdx.x) * CW_SITES + X0 * CW_S) * 2;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int q = lane + 32 * k;
        if (q < NQ) {
            if (PASS == 1) {
                dst[q] = tot[k];
            } else {
                dst[2 * q] = tot[k];
                dst[2 * q + 1] = 0.0;
            }
        }
    }
}

constexpr int CW2_X_SPLIT = 4;                 // blockIdx.z = 0: x in [0, 4); 1: x in [4, 7)
template <int PASS, bool PACKED>
__global__ void __launch_bounds__(32, 16)
cpsnr_window2_kernel(const float* __restrict__ sr, const float* __restrict__ hr, const float* __restrict__ hm, CpGeom g,
                     int clip_sr, const float* __restrict__ bias, double* __restrict__ partial) {
    constexpr int NQ_MAX = (PASS == 1 ? 2 : 1) * CW2_X_SPLIT * CW_S;
    __shared__ float stage[NQ_MAX][33];
    __shared__ float edge_n[PASS == 1 ? CW2_X_SPLIT * CW_S : 1][32];
    __shared__ f2 aux[CW2_X_SPLIT][2][4];
    if (blockIdx.z == 0)
        cpsnr_window2_body<PASS, 0, CW2_X_SPLIT, PACKED>(sr, hr, hm, g, clip_sr, bias, partial, stage, edge_n, aux);
    else
        cpsnr_window2_body<PASS, CW2_X_SPLIT, CW_S - CW2_X_SPLIT, PACKED>(sr, hr, hm, g, clip_sr, bias, partial, stage, edge_n, aux);
}

// One block per imageset, one thread per site.  MODE 1: bias = sum(d*m) / n.  MODE 2: scores + argmax.
template <int MODE>
__global__ void cpsnr_finalize_kernel(const double* __restrict__ partial, CpGeom g, float* __restrict__ bias,
                                      double* __restrict__ nclear, float* __restrict__ best_db,
                                      int32_t* __restrict__ best_site, float* __restrict__ site_db) {
    __shared__ float score[CP_MAXS * CP_MAXS];
    const int set = blockIdx.x, site = threadIdx.x, sites = g.S * g.S;
    if (site < sites) {
        double a0 = 0.0, a1 = 0.0;
        const double* src = partial + (static_cast<size_t>(set) * g.blocks_per_set * sites + site) * 2;
        for (int blk = 0; blk < g.blocks_per_set; ++blk) {
            a0 += src[static_cast<size_t>(blk) * sites * 2];
            a1 += src[static_cast<size_t>(blk) * sites * 2 + 1];
        }
        if (MODE == 1) {
            nclear[set * sites + site] = a0;
            bias[set * sites + site] = static_cast<float>(a1 / a0);          // 0/0 -> NaN like numpy
        } else {
            const double cmse = a0 / nclear[set * sites + site];
            const float s = static_cast<float>(-10.0 * log10(cmse));          // cMSE = 0 -> +inf
            score[site] = s;
            if (site_db != nullptr) site_db[set * sites + site] = s;
        }
    }
    if (MODE == 2) {
        // np.max / np.argmax over the shift window as a warp-shuffle reduction: NaN beats everything and the first NaN
        // (else the first maximum) is the argmax, so the comparator orders by (is NaN, value, lower site index).
        float v = site < sites ? score[site] : -INFINITY;
        int arg = site < sites ? site : 0x7fffffff;
        auto better = [](float av, int ai, float bv, int bi) {
            const bool an = av != av, bn = bv != bv;
            if (an || bn) return (an && bn) ? ai < bi : an;
            if (av != bv) return av > bv;
            return ai < bi;
        };
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
            if (better(ov, oi, v, arg)) {
                v = ov;
                arg = oi;
            }
        }
        __shared__ float wv[2];
        __shared__ int wi[2];
        if ((threadIdx.x & 31) == 0) {
            wv[threadIdx.x >> 5] = v;
            wi[threadIdx.x >> 5] = arg;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const bool second = better(wv[1], wi[1], wv[0], wi[0]);
            best_db[set] = second ? wv[1] : wv[0];
            best_site[set] = second ? wi[1] : wi[0];
        }
    }
}

// ------------------------------------------------------------------ any border_w (Evaluator.py:52 takes any): one block per
// (site, imageset) walks the whole crop; same fp32 element arithmetic, fp64 sums in a fixed order.  Slow path for
// border_w > 3 only (the reference's own callers always use 3).
constexpr int CA_THREADS = 256;
template <int PASS>
__global__ void __launch_bounds__(CA_THREADS)
cpsnr_any_pass_kernel(const float* __restrict__ sr, const float* __restrict__ hr, const float* __restrict__ hm, CpGeom g,
                      int clip_sr, const float* __restrict__ bias, double* __restrict__ partial) {
    __shared__ double red[2][CA_THREADS / 32];
    const int site = blockIdx.x, set = blockIdx.y, sites = g.S * g.S;
    const int x = site / g.S, y = site % g.S;
    const size_t plane = static_cast<size_t>(g.H) * g.W;
    const float* srp = sr + set * plane + static_cast<size_t>(g.border) * g.W + g.border;
    const float* hrp = hr + set * plane + static_cast<size_t>(x) * g.W + y;
    const float* hmp = hm + set * plane + static_cast<size_t>(x) * g.W + y;
    const float b = PASS == 2 ? bias[set * sites + site] : 0.0f;
    double a0 = 0.0, a1 = 0.0;
    for (int i = 0; i < g.size; ++i) {
        float p0 = 0.0f, p1 = 0.0f;
        for (int j = threadIdx.x; j < g.size; j += CA_THREADS) {
            const size_t off = static_cast<size_t>(i) * g.W + j;
            float sv = __ldg(srp + off);
            if (clip_sr) sv = fminf(fmaxf(sv, 0.0f), 1.0f);
            const float m = __ldg(hmp + off), d = __ldg(hrp + off) - sv;
            if (PASS == 1) {
                p0 += m;
                p1 += d * m;
            } else {
                const float t = (d - b) * m;
                p0 += t * t;
            }
        }
        a0 += static_cast<double>(p0);
        a1 += static_cast<double>(p1);
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = a0;
        red[1][threadIdx.x >> 5] = a1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double r0 = 0.0, r1 = 0.0;
        for (int w = 0; w < CA_THREADS / 32; ++w) {
            r0 += red[0][w];
            r1 += red[1][w];
        }
        partial[(static_cast<size_t>(set) * sites + site) * 2] = r0;
        partial[(static_cast<size_t>(set) * sites + site) * 2 + 1] = PASS == 1 ? r1 : 0.0;
    }
}

// One block per imageset; MODE 1: bias and n_clear per site; MODE 2: scores + first-maximum argmax (NaN wins, like numpy).
template <int MODE>
__global__ void __launch_bounds__(CA_THREADS)
cpsnr_any_finalize_kernel(const double* __restrict__ partial, int sites, float* __restrict__ bias, double* __restrict__ nclear,
                          float* __restrict__ best_db, int32_t* __restrict__ best_site, float* __restrict__ site_db) {
    __shared__ float bv[CA_THREADS];
    __shared__ int bi[CA_THREADS];
    const int set = blockIdx.x;
    auto better = [](float av, int ai, float cv, int ci) {
        const bool an = av != av, cn = cv != cv;
        if (an || cn) return (an && cn) ? ai < ci : an;
        if (av != cv) return av > cv;
        return ai < ci;
    };
    float v = -INFINITY;
    int arg = 0x7fffffff;
    for (int site = threadIdx.x; site < sites; site += CA_THREADS) {
        const double a0 = partial[(static_cast<size_t>(set) * sites + site) * 2];
        const double a1 = partial[(static_cast<size_t>(set) * sites + site) * 2 + 1];
        if (MODE == 1) {
            nclear[set * sites + site] = a0;
            bias[set * sites + site] = static_cast<float>(a1 / a0);
        } else {
            const float sc = static_cast<float>(-10.0 * log10(a0 / nclear[set * sites + site]));
            if (site_db != nullptr) site_db[set * sites + site] = sc;
            if (arg == 0x7fffffff || better(sc, site, v, arg)) {
                v = sc;
                arg = site;
            }
        }
    }
    if (MODE == 2) {
        bv[threadIdx.x] = v;
        bi[threadIdx.x] = arg;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int t = 1; t < CA_THREADS; ++t)
                if (bi[t] != 0x7fffffff && better(bv[t], bi[t], v, arg)) {
                    v = bv[t];
                    arg = bi[t];
                }
            best_db[set] = v;
            best_site[set] = arg;
        }
    }
}

// ------------------------------------------------------------------ 49-site search in ONE pass over the data (border_w = 3)
// The two-pass kernels above read sr, hr and the map twice and spend 2 + 4 fp32 instructions per (site, pixel): pass 1 for the
// bias b = sum(m d) / n, pass 2 for sum(((d - b) m)^2).  For a 0/1 map the second sum is sum(m d^2) - n b^2, so one pass that
// accumulates n, sum(m d) and sum(m d^2) per site (4 instructions per (site, pixel)) gives the same number -- provided the
// cancellation in the last step is harmless.  It is made harmless by CENTRING: every work item subtracts a constant c of its
// own from d (the masked mean difference of its first row, i.e. a good guess of the bias), accumulates the centred sums in
// fp32 over at most 16 rows and folds them into fp64; the finalize kernel un-centres every item's sums in fp64
// (sum(m d) = S1 + c n, sum(m d^2) = S2 + 2 c S1 + c^2 n), where a cancellation of even 1e6 costs nothing.  What is left is the
// fp32 rounding of the centred partial sums, ~1e-7 of sum(m (d - c)^2): sites where that is more than 16 x the cMSE (or whose
// result is not a positive finite number: empty masks, NaN / inf pixels, exact matches) are flagged and recomputed by the
// two-pass fallback kernel below with the reference's element arithmetic; imagesets whose map is not 0/1 are flagged whole.
// Tile: one warp per block, 6 crop columns per lane (192 per block: 378 = 192 + 186, two blocks, 63 of 64 lanes busy); a
// block walks work items (imageset, column block, row band) in a fixed round-robin, the grid is one resident wave.
constexpr int OP_TC = 6, OP_COLS = 32 * OP_TC, OP_S = 7, OP_SITES = OP_S * OP_S, OP_FLUSH = 16;   // OP_TC: even
constexpr int OP_WIN = OP_TC + OP_S - 1;         // 12 hr / map columns per lane and row, as six float2
constexpr int OP_NQ = 3 * OP_SITES;              // per site: n, S1, S2
constexpr int OP_STRIDE = OP_NQ + 2;             // doubles per work item: the sums, c, "map is not 0/1"
constexpr float OP_TRUST = 16.0f;                // flag a site when sum(m (d - c)^2) > OP_TRUST * n * cMSE
constexpr int OP_MAX_REDO = 6;                   // more flagged sites than this in one imageset: redo the whole imageset instead

struct OpGeom {
    int H, W, size, col_blocks, band_rows, bands, items, items_per_set;
};
struct OpRow {
    float2 h[OP_WIN / 2], m[OP_WIN / 2];
    float s0;                                    // sr crop columns j0 .. j0 + OP_TC - 1 are image columns j0 + 3 ..: float,
    float2 smid[(OP_TC - 2) / 2];                // (OP_TC - 2) / 2 aligned float2,
    float slast;                                 // float
};

// All terms of one hr row.  The seven shifted differences of one (sr row, column) pair are formed side by side -- seven FADD,
// seven FMUL, seven FADD, seven FFMA -- so that no instruction waits for its predecessor: with two scratch registers (what the
// compiler chose for the straightforward loop once the 98 accumulators had filled the register file) every instruction
// waited out the 4-cycle latency of the one before it.  The sr window lives in shared memory for the same reason (42 registers).
template <bool CPRED, bool XALL>
__device__ __forceinline__ void op_terms(float (&s2)[OP_S][OP_S], float (&s1)[OP_S][OP_S], const float (&hw)[OP_WIN],
                                         const float (&mw)[OP_WIN], const float (*svs)[OP_TC][32], int lane, int slot0,
                                         int xlo, int xhi, int ncol) {
    int slot = slot0;                          // slot of sr crop row h (x = 0); row h - x sits x slots back in the ring of 7
#pragma unroll
    for (int x = 0; x < OP_S; ++x) {
        if (XALL || (x >= xlo && x <= xhi)) {  // warp-uniform: sr crop row h - x lies inside this band
#pragma unroll
            for (int c = 0; c < OP_TC; ++c) {
                if (CPRED && c >= ncol) continue;
                const float sv = svs[slot][c][lane];
                float d[OP_S], md[OP_S];
#pragma unroll
                for (int y = 0; y < OP_S; ++y) d[y] = hw[c + y] - sv;            // (hr - sr) - centre          (Evaluator.py:35)
#pragma unroll
                for (int y = 0; y < OP_S; ++y) md[y] = d[y] * mw[c + y];        // exact for a 0/1 map
#pragma unroll
                for (int y = 0; y < OP_S; ++y) s1[x][y] += md[y];
#pragma unroll
                for (int y = 0; y < OP_S; ++y) s2[x][y] = fmaf(md[y], d[y], s2[x][y]);
            }
        }
        slot = slot == 0 ? OP_S - 1 : slot - 1;
    }
}

__global__ void __launch_bounds__(32, 7)
cpsnr_onepass_kernel(const float* __restrict__ sr, const float* __restrict__ hr, const float* __restrict__ hm, OpGeom g,
                     int clip_sr, double* __restrict__ partial, int* __restrict__ redo_list) {
    __shared__ float stage[OP_SITES][33];
    if (blockIdx.x == 0 && threadIdx.x == 0) redo_list[0] = 0;       // number of flagged (site, imageset) pairs, filled in by the scores kernel
    __shared__ float edge_n[OP_SITES][32];       // n of the rows at the top / bottom of a band (not all 7 sr partners inside)
    __shared__ float svs[OP_S][OP_TC][32];       // ring of the last seven sr crop rows (+ centre), [slot][column][lane]
    __shared__ double tot[6][32];                // fp64 running sums: [quantity * 2 + k][lane] for site lane + 32 k
    const int lane = threadIdx.x;
    const size_t plane = static_cast<size_t>(g.H) * g.W;
    const float2 zero2 = make_float2(0.f, 0.f);
    for (int item = blockIdx.x; item < g.items; item += gridDim.x) {
        const int set = item / g.items_per_set, rest = item % g.items_per_set;
        const int cb = rest / g.bands, band = rest % g.bands;
        const int j0 = cb * OP_COLS + lane * OP_TC;                  // first crop column of this lane
        const int ncol = max(0, min(OP_TC, g.size - j0));
        // a lane beyond the crop loads nothing: zeros for hr, sr and the map add nothing to any sum, so it can run the same
        // straight-line code as the others (378 = 63 x 6: with six columns per lane no lane is ever partly filled at 384^2)
        const bool active = ncol > 0;
        const bool all_full = __all_sync(0xffffffffu, ncol == OP_TC || ncol == 0);
        const float* srp = sr + set * plane + static_cast<size_t>(3) * g.W + j0;     // crop (i, j) = image (i + 3, j + 3)
        const float* hrp = hr + set * plane + j0;
        const float* hmp = hm + set * plane + j0;
        const int i0 = band * g.band_rows, i1 = min(g.size, i0 + g.band_rows);
        const int h_end = i1 + OP_S - 1;                             // hr rows [i0, h_end) meet sr crop rows [i0, i1)
        bool q_ok[OP_WIN / 2];
#pragma unroll
        for (int q = 0; q < OP_WIN / 2; ++q) q_ok[q] = j0 + 2 * q + 2 <= g.W;

        float s1[OP_S][OP_S], s2[OP_S][OP_S], n_all[OP_S];
#pragma unroll
        for (int x = 0; x < OP_S; ++x) {
            n_all[x] = 0.0f;
#pragma unroll
            for (int y = 0; y < OP_S; ++y) {
                s1[x][y] = 0.0f;
                s2[x][y] = 0.0f;
                edge_n[x * OP_S + y][lane] = 0.0f;
            }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) tot[k][lane] = 0.0;
        uint32_t not_binary = 0;

        // fold the fp32 partial sums into fp64: quantity by quantity (n, S1, S2) through one 49 x 32 staging tile -- every lane
        // writes its 49 values, then lanes 0 .. 48 (two rounds) each add up one site's 32 values in a fixed order
        auto fold = [&](int which) {
            __syncwarp();
#pragma unroll 1
            for (int k = 0; k < 2; ++k) {
                const int site = lane + 32 * k;
                if (site < OP_SITES) {
                    double a = 0.0;
#pragma unroll 8
                    for (int l = 0; l < 32; ++l) a += static_cast<double>(stage[site][l]);
                    tot[which * 2 + k][lane] += a;
                }
            }
            __syncwarp();
        };
        auto flush = [&]() {
#pragma unroll
            for (int x = 0; x < OP_S; ++x)
#pragma unroll
                for (int y = 0; y < OP_S; ++y) {
                    stage[x * OP_S + y][lane] = n_all[y] + edge_n[x * OP_S + y][lane];
                    edge_n[x * OP_S + y][lane] = 0.0f;
                }
#pragma unroll
            for (int y = 0; y < OP_S; ++y) n_all[y] = 0.0f;
            fold(0);
#pragma unroll
            for (int x = 0; x < OP_S; ++x)
#pragma unroll
                for (int y = 0; y < OP_S; ++y) {
                    stage[x * OP_S + y][lane] = s1[x][y];
                    s1[x][y] = 0.0f;
                }
            fold(1);
#pragma unroll
            for (int x = 0; x < OP_S; ++x)
#pragma unroll
                for (int y = 0; y < OP_S; ++y) {
                    stage[x * OP_S + y][lane] = s2[x][y];
                    s2[x][y] = 0.0f;
                }
            fold(2);
        };
        auto load_row = [&](int h, OpRow& r) {
            const size_t off = static_cast<size_t>(h) * g.W;
            const bool in = active && h < h_end;
#pragma unroll
            for (int q = 0; q < OP_WIN / 2; ++q) {
                r.h[q] = (in && q_ok[q]) ? __ldg(reinterpret_cast<const float2*>(hrp + off) + q) : zero2;
                r.m[q] = (in && q_ok[q]) ? __ldg(reinterpret_cast<const float2*>(hmp + off) + q) : zero2;
            }
            const bool sin = active && h < i1;                       // sr crop row h exists in this band
            // every loaded value is used: a dead lane of a wider load gets its register recycled while the load is in flight
            // (see cpsnr_window_kernel)
            r.s0 = (sin && q_ok[1]) ? __ldg(srp + off + 3) : 0.0f;
#pragma unroll
            for (int q = 0; q < (OP_TC - 2) / 2; ++q)
                r.smid[q] = (sin && q_ok[q + 2]) ? __ldg(reinterpret_cast<const float2*>(srp + off + 4) + q) : zero2;
            r.slast = (sin && q_ok[OP_TC / 2 + 1]) ? __ldg(srp + off + OP_TC + 2) : 0.0f;
        };

        float centre = 0.0f;
        OpRow nxt;
        load_row(i0, nxt);
        int since_flush = 0, slot = 0;
        for (int h = i0; h < h_end; ++h) {
            const OpRow cur = nxt;
            load_row(h + 1, nxt);                                    // in flight while this row is being consumed
            float hw[OP_WIN], mw[OP_WIN];
#pragma unroll
            for (int q = 0; q < OP_WIN / 2; ++q) {
                hw[2 * q] = cur.h[q].x, hw[2 * q + 1] = cur.h[q].y;
                mw[2 * q] = cur.m[q].x, mw[2 * q + 1] = cur.m[q].y;
            }
#pragma unroll
            for (int k = 0; k < OP_WIN; ++k) {                       // anything but +0.0f or 1.0f in the map?
                const uint32_t bits = __float_as_uint(mw[k]);
                not_binary |= min(bits, bits ^ 0x3f800000u);
            }
            float sv[OP_TC];
            sv[0] = cur.s0;
#pragma unroll
            for (int q = 0; q < (OP_TC - 2) / 2; ++q) sv[1 + 2 * q] = cur.smid[q].x, sv[2 + 2 * q] = cur.smid[q].y;
            sv[OP_TC - 1] = cur.slast;
            if (clip_sr) {
#pragma unroll
                for (int c = 0; c < OP_TC; ++c) sv[c] = fminf(fmaxf(sv[c], 0.0f), 1.0f);
            }
            if (h == i0) {
                // centring constant of this item: masked mean of hr - sr over its first row at shift (0, 3); any finite
                // value is correct, a good one keeps the fp32 partial sums small
                float num = 0.0f, den = 0.0f;
#pragma unroll
                for (int c = 0; c < OP_TC; ++c)
                    if (c < ncol) {
                        num = fmaf(mw[c + 3], hw[c + 3] - sv[c], num);
                        den += mw[c + 3];
                    }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    num += __shfl_xor_sync(0xffffffffu, num, o);
                    den += __shfl_xor_sync(0xffffffffu, den, o);
                }
                const float c0 = num / den;
                centre = (den > 0.0f && fabsf(c0) <= 4.0f) ? c0 : 0.0f;   // also false for NaN
            }
            slot = slot == OP_S - 1 ? 0 : slot + 1;                  // sr crop row h goes into the next slot of the ring
#pragma unroll
            for (int c = 0; c < OP_TC; ++c) svs[slot][c][lane] = sv[c] + centre;
            // sr crop rows i = h - x that lie in this band: x in [xlo, xhi]
            const int xlo = max(0, h - i1 + 1), xhi = min(OP_S - 1, h - i0);
            if (all_full && xlo == 0 && xhi == OP_S - 1) {
                // the hot path, straight-line code: every lane has all its columns (or none) and all seven sr partners of this
                // hr row lie inside the band
#pragma unroll
                for (int y = 0; y < OP_S; ++y) {
                    float r = 0.0f;
#pragma unroll
                    for (int c = 0; c < OP_TC; ++c) r += mw[c + y];                               // n_clear (Evaluator.py:34)
                    n_all[y] += r;
                }
                op_terms<false, true>(s2, s1, hw, mw, svs, lane, slot, 0, OP_S - 1, OP_TC);
            } else {
                float rs[OP_S];                                      // sum of the map over this lane's columns, per column shift
#pragma unroll
                for (int y = 0; y < OP_S; ++y) {
                    float r = 0.0f;
#pragma unroll
                    for (int c = 0; c < OP_TC; ++c) r += (all_full || c < ncol) ? mw[c + y] : 0.0f;
                    rs[y] = r;
                }
                if (all_full)                                        // the rows at the top / bottom of a band
                    op_terms<false, false>(s2, s1, hw, mw, svs, lane, slot, xlo, xhi, OP_TC);
                else                                                 // a partly filled lane (crop widths that 6 does not divide)
                    op_terms<true, false>(s2, s1, hw, mw, svs, lane, slot, xlo, xhi, ncol);
                for (int x = xlo; x <= xhi; ++x)
#pragma unroll
                    for (int y = 0; y < OP_S; ++y) edge_n[x * OP_S + y][lane] += rs[y];
            }
            if (++since_flush == OP_FLUSH) {
                flush();
                since_flush = 0;
            }
        }
        if (since_flush > 0) flush();
        double* dst = partial + static_cast<size_t>(item) * OP_STRIDE;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int site = lane + 32 * k;
            if (site < OP_SITES) {
#pragma unroll
                for (int which = 0; which < 3; ++which) dst[site * 3 + which] = tot[which * 2 + k][lane];
            }
        }
        const bool any_bad = __any_sync(0xffffffffu, not_binary != 0);
        if (lane == 0) {
            dst[OP_NQ] = static_cast<double>(centre);
            dst[OP_NQ + 1] = any_bad ? 1.0 : 0.0;
        }
        __syncwarp();
    }
}

// One block per imageset, one thread per site: un-centre and add the items' sums in fp64, score, decide which sites the
// fallback has to redo.
__global__ void cpsnr_onepass_scores_kernel(const double* __restrict__ partial, OpGeom g, float* __restrict__ score,
                                            uint8_t* __restrict__ redo, int* __restrict__ redo_list, uint8_t* __restrict__ set_sel) {
    const int set = blockIdx.x, site = threadIdx.x;
    bool flagged = false, bad_map = false;
    if (site < OP_SITES) {
        double n = 0.0, a = 0.0, q = 0.0, e = 0.0;
        const double* src = partial + static_cast<size_t>(set) * g.items_per_set * OP_STRIDE;
        for (int it = 0; it < g.items_per_set; ++it, src += OP_STRIDE) {
            const double nk = src[site * 3], s1 = src[site * 3 + 1], s2 = src[site * 3 + 2], c = src[OP_NQ];
            n += nk;
            a += s1 + c * nk;
            q += s2 + 2.0 * c * s1 + c * c * nk;
            e += s2;
            bad_map = bad_map || src[OP_NQ + 1] != 0.0;
        }
        const double b = a / n, cmse = q / n - b * b;
        const bool trusted = cmse > 0.0 && cmse < 1e300 && e <= static_cast<double>(OP_TRUST) * n * cmse;   // false for NaN
        score[set * OP_SITES + site] = static_cast<float>(-10.0 * log10(cmse));
        flagged = !trusted;
    }
    // A few flagged sites are redone one by one (each re-reads the imageset twice); an imageset with many of them, or with a
    // map that is not 0/1, goes through the two-pass window kernels as a whole instead (two more reads for all 49 sites).
    const int count = __syncthreads_count(flagged);
    const bool whole = __syncthreads_or(bad_map) || count > OP_MAX_REDO;
    if (threadIdx.x == 0) set_sel[set] = whole ? 1 : 0;
    if (site < OP_SITES) {
        const bool single = flagged && !whole;
        redo[set * OP_SITES + site] = single ? 1 : 0;
        if (single) redo_list[1 + atomicAdd(redo_list, 1)] = set * OP_SITES + site;   // [0] = count (zeroed by the one-pass kernel)
    }
}

// Fallback for flagged (site, imageset) pairs: the reference's two passes with the reference's element arithmetic (diff,
// diff * map, (diff - bias) * map, square in fp32; sums in fp64).  RD_SLICES blocks share the rows of a pair, so 32 flagged
// pairs already fill the GPU; pass 2 adds up the slices of pass 1 itself, the argmax kernel those of pass 2.  Blocks of
// the flagged pairs come as a list (count first) written by the scores kernel: usually empty, two launches of a few microseconds.
constexpr int RD_THREADS = 256, RD_SLICES = 16;
template <int PASS>
__global__ void __launch_bounds__(RD_THREADS)
cpsnr_redo_kernel(const float* __restrict__ sr, const float* __restrict__ hr, const float* __restrict__ hm, int H, int W,
                  int clip_sr, const int* __restrict__ redo_list, double* __restrict__ slices) {
    __shared__ double red[2][RD_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units = redo_list[0] * RD_SLICES;                      // usually 0: the whole grid leaves at once
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int pair = redo_list[1 + unit / RD_SLICES], slice = unit % RD_SLICES;
    const int set = pair / OP_SITES, site = pair % OP_SITES;
    const int x = site / OP_S, y = site % OP_S, size = W - 6;
    const size_t plane = static_cast<size_t>(H) * W;
    const float* srp = sr + set * plane + static_cast<size_t>(3) * W + 3;
    const float* hrp = hr + set * plane + static_cast<size_t>(x) * W + y;
    const float* hmp = hm + set * plane + static_cast<size_t>(x) * W + y;
    // slices[pair][slice] = {n, sum d m, sum t t}
    float b = 0.0f;
    if (PASS == 2) {
        double n = 0.0, a = 0.0;
        for (int k = 0; k < RD_SLICES; ++k) {
            n += slices[(static_cast<size_t>(pair) * RD_SLICES + k) * 3];
            a += slices[(static_cast<size_t>(pair) * RD_SLICES + k) * 3 + 1];
        }
        b = static_cast<float>(a / n);                               // 0/0 -> NaN like numpy
    }
    const int rows = (size + RD_SLICES - 1) / RD_SLICES, i0 = slice * rows, i1 = min(size, i0 + rows);
    double a0 = 0.0, a1 = 0.0;
    for (int i = i0 + warp; i < i1; i += RD_THREADS / 32) {
        float p0 = 0.0f, p1 = 0.0f;
#pragma unroll 12
        for (int j = lane; j < size; j += 32) {
            const size_t off = static_cast<size_t>(i) * W + j;
            float sv = __ldg(srp + off);
            if (clip_sr) sv = fminf(fmaxf(sv, 0.0f), 1.0f);
            const float m = __ldg(hmp + off), d = __ldg(hrp + off) - sv;
            if (PASS == 1) {
                p0 += m;
                p1 += d * m;
            } else {
                const float t = (d - b) * m;
                p0 += t * t;
            }
        }
        a0 += static_cast<double>(p0);
        a1 += static_cast<double>(p1);
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) {
        red[0][warp] = a0;
        red[1][warp] = a1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double r0 = 0.0, r1 = 0.0;
        for (int w = 0; w < RD_THREADS / 32; ++w) {
            r0 += red[0][w];
            r1 += red[1][w];
        }
        double* dst = slices + (static_cast<size_t>(pair) * RD_SLICES + slice) * 3;
        if (PASS == 1) {
            dst[0] = r0;
            dst[1] = r1;
        } else {
            dst[2] = r0;
        }
    }
    __syncthreads();                                                 // red[] is reused by the next unit
    }
}

// np.max / np.argmax over the 49 scores of an imageset: NaN beats everything, the first NaN (else the first maximum) wins.
__global__ void cpsnr_argmax_kernel(const float* __restrict__ score, const uint8_t* __restrict__ redo,
                                    const double* __restrict__ slices, const uint8_t* __restrict__ set_sel,
                                    const double* __restrict__ pass1, const double* __restrict__ pass2, int blocks_per_set,
                                    float* __restrict__ best_db, int32_t* __restrict__ best_site, float* __restrict__ site_db) {
    const int set = blockIdx.x, site = threadIdx.x;
    float v = site < OP_SITES ? score[set * OP_SITES + site] : -INFINITY;
    if (site < OP_SITES && set_sel[set]) {                           // imageset redone by the two-pass window kernels: cpsnr_finalize_kernel<2>
        double n = 0.0, q = 0.0;
        const size_t first = (static_cast<size_t>(set) * blocks_per_set * OP_SITES + site) * 2;
        for (int blk = 0; blk < blocks_per_set; ++blk) {
            n += pass1[first + static_cast<size_t>(blk) * OP_SITES * 2];
            q += pass2[first + static_cast<size_t>(blk) * OP_SITES * 2];
        }
        v = static_cast<float>(-10.0 * log10(q / n));                // cMSE = 0 -> +inf
    }
    if (site < OP_SITES && redo[set * OP_SITES + site]) {            // the fallback's answer replaces the one-pass score
        double n = 0.0, q = 0.0;
        for (int k = 0; k < RD_SLICES; ++k) {
            n += slices[(static_cast<size_t>(set * OP_SITES + site) * RD_SLICES + k) * 3];
            q += slices[(static_cast<size_t>(set * OP_SITES + site) * RD_SLICES + k) * 3 + 2];
        }
        v = static_cast<float>(-10.0 * log10(q / n));                // cMSE = 0 -> +inf
    }
    int arg = site < OP_SITES ? site : 0x7fffffff;
    if (site < OP_SITES && site_db != nullptr) site_db[set * OP_SITES + site] = v;
    auto better = [](float av, int ai, float bv, int bi) {
        const bool an = av != av, bn = bv != bv;
        if (an || bn) return (an && bn) ? ai < bi : an;
        if (av != bv) return av > bv;
        return ai < bi;
    };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
        if (better(ov, oi, v, arg)) {
            v = ov;
            arg = oi;
        }
    }
    __shared__ float wv[2];
    __shared__ int wi[2];
    if ((threadIdx.x & 31) == 0) {
        wv[threadIdx.x >> 5] = v;
        wi[threadIdx.x >> 5] = arg;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const bool second = better(wv[1], wi[1], wv[0], wi[0]);
        best_db[set] = second ? wv[1] : wv[0];
        best_site[set] = second ? wi[1] : wi[0];
    }
}

// ------------------------------------------------------------------ clear loss (train.py:66-87 without autograd)
// metric 0: masked_MSE = mean over ALL pixels of (m*sr - m*hr)^2
// metric 1: cMSE       = sum(m * (sr + b - hr)^2) / sum(m),  b = sum(m * (hr - sr)) / sum(m)   (weight m, not m^2)
// metric 2: cPSNR      = -10 log10(cMSE)
constexpr int CL_BLOCKS = 16, CL_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* scratch) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < CL_THREADS / 32; ++w) r += scratch[w];       // fixed order: deterministic
    return r;
}

// PASS 0: sum((m*sr - m*hr)^2).  PASS 1: sum(m), sum(m * (hr - sr)).  PASS 2: sum(m * ((sr + b) - hr)^2).
template <int PASS>
__global__ void __launch_bounds__(CL_THREADS)
clear_loss_pass_kernel(const float* __restrict__ sr, const float* __restrict__ hr, const float* __restrict__ hm,
                       size_t hw, const float* __restrict__ bias, double* __restrict__ partial) {
    __shared__ double scratch[CL_THREADS / 32];
    const int img = blockIdx.y;
    const float *s = sr + img * hw, *h = hr + img * hw, *m = hm + img * hw;
    const float b = PASS == 2 ? bias[img] : 0.0f;
    double a0 = 0.0, a1 = 0.0;
    for (size_t i = blockIdx.x * static_cast<size_t>(CL_THREADS) + threadIdx.x; i < hw; i += static_cast<size_t>(CL_BLOCKS) * CL_THREADS) {
        const float sv = __ldg(s + i), hv = __ldg(h + i), mv = __ldg(m + i);
        if (PASS == 0) {
            const float e = mv * sv - mv * hv;
            a0 += static_cast<double>(e * e);
        } else if (PASS == 1) {
            a0 += static_cast<double>(mv);
            a1 += static_cast<double>(mv * (hv - sv));
        } else {
            const float e = (sv + b) - hv;
            a0 += static_cast<double>(mv * (e * e));
        }
    }
    const double r0 = block_sum(a0, scratch);
    const double r1 = PASS == 1 ? block_sum(a1, scratch) : 0.0;
    if (threadIdx.x == 0) {
        partial[(static_cast<size_t>(img) * CL_BLOCKS + blockIdx.x) * 2] = r0;
        partial[(static_cast<size_t>(img) * CL_BLOCKS + blockIdx.x) * 2 + 1] = r1;
    }
}

// MODE 0: masked_MSE out.  MODE 1: bias + n_clear.  MODE 2: cMSE / cPSNR out.
template <int MODE>
__global__ void clear_loss_finalize_kernel(const double* __restrict__ partial, int B, double n_pix, int metric,
                                           float* __restrict__ bias, double* __restrict__ nclear, float* __restrict__ out) {
    const int img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= B) return;
    double a0 = 0.0, a1 = 0.0;
    for (int k = 0; k < CL_BLOCKS; ++k) {
        a0 += partial[(static_cast<size_t>(img) * CL_BLOCKS + k) * 2];
        a1 += partial[(static_cast<size_t>(img) * CL_BLOCKS + k) * 2 + 1];
    }
    if (MODE == 0) {
        out[img] = static_cast<float>(a0 / n_pix);
    } else if (MODE == 1) {
        nclear[img] = a0;
        bias[img] = static_cast<float>(a1 / a0);
    } else {
        const double cmse = a0 / nclear[img];
        out[img] = metric == 1 ? static_cast<float>(cmse) : static_cast<float>(-10.0 * log10(cmse));
    }
}

}  // namespace

// Scratch of the scoring entry points (they take no handle): a private stream-ordered pool per device that KEEPS its
// memory across synchronisations.  With the default pool (release threshold 0) every call that follows a
// cudaStreamSynchronize -- the normal pattern of a validation loop -- paid a fresh driver allocation: 3.8 ms for a 0.15 ms
// search on 32 imagesets (tools/cpsnr_trace.py).
static int scratch_alloc(void** p, size_t bytes, cudaStream_t s) {
    constexpr int MAX_DEV = 64;
    static cudaMemPool_t pools[MAX_DEV] = {};
    int dev = 0;
    HRN_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEV) {
        set_error("scratch_alloc: device %d out of range", dev);
        return -1;
    }
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (pools[dev] == nullptr) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        HRN_CUDA_OK(cudaMemPoolCreate(&pool, &props));
        unsigned long long keep = ~0ull;
        HRN_CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        pools[dev] = pool;
    }
    HRN_CUDA_OK(cudaMallocFromPoolAsync(p, bytes, pools[dev], s));
    return 0;
}

int lanczos_shift_launch(const float* img, const float* shift, int nb, int c, int H, int W, int p, int a, int ntaps,
                         float* out, cudaStream_t s) {
    if (ntaps < 1 || ntaps > MAX_TAPS || (ntaps & 1) == 0) {
        set_error("lanczos_shift: kernel width N=%d unsupported (odd, <= %d)", ntaps, MAX_TAPS);
        return -1;
    }
    if (p < 0 || p >= H || p >= W) {
        set_error("lanczos_shift: reflect padding p=%d must be smaller than the image (%d x %d)", p, H, W);
        return -1;
    }
    if (a <= 0) {
        set_error("lanczos_shift: a must be positive");
        return -1;
    }
    const long long planes = static_cast<long long>(nb) * c;
    if (planes <= 0 || planes > 65535) {
        set_error("lanczos_shift: %lld planes outside [1, 65535]", planes);
        return -1;
    }
    if (ntaps == 7) {
        // the width the reference uses everywhere (ShiftNet.py:87-89): compile-time tile geometry
        float* taps = nullptr;
        if (scratch_alloc(reinterpret_cast<void**>(&taps), static_cast<size_t>(c) * 2 * 7 * sizeof(float), s)) return -1;
        lanczos_taps_kernel<<<(2 * c + 127) / 128, 128, 0, s>>>(shift, 2 * c, a, 7, taps);
        if (!g_lanczos_scalar && lanczos7_tma_usable(img, out, H, W)) {
            // TMA-fed tiles (lanczos7_tma.cu): 16-byte aligned rows
            if (lanczos7_tma_launch(img, taps, static_cast<int>(planes), c, H, W, p, out, s)) return -1;
        } else {
            dim3 grid((W + L7_TW - 1) / L7_TW, (H + L7_TH - 1) / L7_TH, static_cast<unsigned>(planes));
            lanczos_shift7_kernel<<<grid, L7_THREADS, 0, s>>>(img, taps, c, H, W, p, out);
        }
        note_launches(1);
        HRN_CUDA_OK(cudaFreeAsync(taps, s));
    } else {
        const int half = ntaps / 2;
        const size_t smem = (static_cast<size_t>(LZ_TH + 2 * half) + LZ_TH) * (LZ_TW + 2 * half) * sizeof(float);
        dim3 grid((W + LZ_TW - 1) / LZ_TW, (H + LZ_TH - 1) / LZ_TH, static_cast<unsigned>(planes));
        lanczos_shift_kernel<<<grid, LZ_THREADS, smem, s>>>(img, shift, c, H, W, p, a, ntaps, out);
    }
    note_launches(1);
    HRN_CUDA_OK(cudaGetLastError());
    return 0;
}

int lanczos_taps_launch(const float* d, int n, int a, int ntaps, float* out, cudaStream_t s) {
    if (ntaps < 1 || ntaps > MAX_TAPS || (ntaps & 1) == 0 || a <= 0 || n <= 0) {
        set_error("lanczos_taps: need n > 0, a > 0 and odd N <= %d (got n=%d a=%d N=%d)", MAX_TAPS, n, a, ntaps);
        return -1;
    }
    lanczos_taps_kernel<<<(n + 127) / 128, 128, 0, s>>>(d, n, a, ntaps, out);
    note_launches(1);
    HRN_CUDA_OK(cudaGetLastError());
    return 0;
}

int g_cpsnr_generic = 0;
int g_lanczos_scalar = 0;       // test knob: 1 = the register-window Lanczos kernel also for rows that TMA can address
// Measured on a B200 (profiles/r02_cpsnr_ab.log, 512 x 384^2): scalar window kernel 1.86 ms, packed / split kernel 2.25 ms;
// with the batch cut into L2-sized chunks 2.13-3.85 ms (chunk 64 ... 12).  ncu on the scalar kernel: issue slots 44-51 % active,
// fp32 pipe 27-40 %, DRAM 12-14 %: latency-bound at three warps per scheduler, and an FFMA2 is no cheaper for the pipe than
// two FFMAs, so neither packing nor L2 residency of pass 2 pays; both stay behind knobs as measured alternatives.
int g_cpsnr_window_v1 = -1;    // -1 (default) = automatic: the split scalar kernel for small batches (twice the blocks: 0.131 vs 0.145 ms
                               // on 32 imagesets), the 49-sites-per-warp kernel for large ones (1.86 vs 2.01 ms on 512); 1 = scalar
                               // 49-sites-per-warp kernel; 0 = x split over two warps + packed fp32x2; 2 = x split, scalar fp32
int g_cpsnr_chunk = 0;         // imagesets per pass-1 / pass-2 round trip: 0 (default) = whole batch; -1 = by L2 budget; n > 0 = n
constexpr size_t CP_L2_BUDGET = 56ull << 20;   // chunk = -1: bytes of sr + hr + map per chunk that pass 2 should still find in L2
constexpr int CW2_TARGET_WARPS = 148 * 14;

int g_cpsnr_onepass = 1;       // 1 (default) = border_w = 3 on aligned rows takes the one-pass kernel; 0 = the two-pass window kernels

// border_w = 3, rows 16-byte aligned: one pass over the data, scores + flags, the two fallback passes for single flagged sites, the
// two-pass window kernels for imagesets flagged as a whole (pass 2 forms the bias, the argmax kernel the scores), argmax.  The
// four fallback launches leave at once when nothing is flagged.
static int shift_cpsnr_onepass(const float* sr, const float* hr, const float* hm, int B, int H, int W, int clip_sr,
                               float* best_db, int32_t* best_site, float* site_db, cudaStream_t s) {
    int dev = 0, sm_count = 0, per_sm = 0;
    HRN_CUDA_OK(cudaGetDevice(&dev));
    HRN_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    HRN_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cpsnr_onepass_kernel, 32, 0));
    const int resident = sm_count * (per_sm < 1 ? 1 : per_sm);
    OpGeom g;
    g.H = H;
    g.W = W;
    g.size = W - 6;
    g.col_blocks = (g.size + OP_COLS - 1) / OP_COLS;
    // Row bands: every block of the one resident wave walks items round-robin, so the cost is the number of rounds; pick the
    // band count (bands of at least 24 rows, each pays 6 extra row loads) that wastes the least of the last round.
    const long long columns = static_cast<long long>(B) * g.col_blocks;
    const int max_bands = g.size / 24 < 1 ? 1 : g.size / 24;
    int best_bands = 1;
    double best_cost = 1e300;
    for (int bands = 1; bands <= max_bands; ++bands) {
        const int rows = (g.size + bands - 1) / bands;
        const long long items = columns * ((g.size + rows - 1) / rows);
        const long long rounds = (items + resident - 1) / resident;
        const double cost = static_cast<double>(rounds) * (rows + 8);           // row iterations of the busiest block (+ fill / flush)
        if (cost < best_cost * 0.995) {
            best_cost = cost;
            best_bands = bands;
        }
    }
    g.band_rows = (g.size + best_bands - 1) / best_bands;
    g.bands = (g.size + g.band_rows - 1) / g.band_rows;
    g.items_per_set = g.col_blocks * g.bands;
    const long long items = static_cast<long long>(B) * g.items_per_set;
    g.items = static_cast<int>(items);
    const size_t partial_bytes = static_cast<size_t>(items) * OP_STRIDE * sizeof(double);
    const size_t slice_bytes = static_cast<size_t>(B) * OP_SITES * RD_SLICES * 3 * sizeof(double);
    const size_t score_bytes = static_cast<size_t>(B) * OP_SITES * sizeof(float);
    const size_t list_bytes = (static_cast<size_t>(B) * OP_SITES + 1) * sizeof(int);
    // whole-imageset fallback: the two-pass 49-sites-per-warp window kernels on the selected imagesets
    CpGeom g2;
    g2.H = H;
    g2.W = W;
    g2.border = 3;
    g2.S = CW_S;
    g2.size = g.size;
    g2.vec_ok = 1;
    g2.col_blocks = (g.size + CW_COLS - 1) / CW_COLS;
    int want = CW_TARGET_WARPS / (B * g2.col_blocks);
    want = want < 1 ? 1 : (want > (g.size + 7) / 8 ? (g.size + 7) / 8 : want);
    g2.band_rows = (g.size + want - 1) / want;
    g2.blocks_per_set = ((g.size + g2.band_rows - 1) / g2.band_rows) * g2.col_blocks;
    const size_t partial2_bytes = static_cast<size_t>(B) * g2.blocks_per_set * OP_SITES * 2 * sizeof(double);
    uint8_t* ws = nullptr;
    if (scratch_alloc(reinterpret_cast<void**>(&ws), partial_bytes + slice_bytes + 2 * partial2_bytes + score_bytes + list_bytes +
                                                     static_cast<size_t>(B) * (OP_SITES + 1), s))
        return -1;
    uint8_t* at = ws;
    auto take = [&](size_t bytes) {
        uint8_t* p = at;
        at += bytes;
        return p;
    };
    double* partial = reinterpret_cast<double*>(take(partial_bytes));
    double* slices = reinterpret_cast<double*>(take(slice_bytes));
    double* partial2 = reinterpret_cast<double*>(take(partial2_bytes));      // pass 1 of the whole-imageset fallback: n, sum d m
    double* partial3 = reinterpret_cast<double*>(take(partial2_bytes));      // pass 2: sum t t
    float* score = reinterpret_cast<float*>(take(score_bytes));
    int* redo_list = reinterpret_cast<int*>(take(list_bytes));
    uint8_t* redo = take(static_cast<size_t>(B) * OP_SITES);
    uint8_t* set_sel = take(B);
    const int grid = items < resident ? static_cast<int>(items) : resident;
    cpsnr_onepass_kernel<<<grid, 32, 0, s>>>(sr, hr, hm, g, clip_sr, partial, redo_list);
    cpsnr_onepass_scores_kernel<<<B, 64, 0, s>>>(partial, g, score, redo, redo_list, set_sel);
    const long long max_units = static_cast<long long>(B) * OP_SITES * RD_SLICES;
    const int rgrid = static_cast<int>(max_units < 4LL * sm_count ? max_units : 4LL * sm_count);
    cpsnr_redo_kernel<1><<<rgrid, RD_THREADS, 0, s>>>(sr, hr, hm, H, W, clip_sr, redo_list, slices);
    cpsnr_redo_kernel<2><<<rgrid, RD_THREADS, 0, s>>>(sr, hr, hm, H, W, clip_sr, redo_list, slices);
    const dim3 wgrid(g2.blocks_per_set, B);
    cpsnr_window_kernel<1, true><<<wgrid, 32, 0, s>>>(sr, hr, hm, g2, clip_sr, nullptr, partial2, set_sel, nullptr);
    cpsnr_window_kernel<2, true><<<wgrid, 32, 0, s>>>(sr, hr, hm, g2, clip_sr, nullptr, partial3, set_sel, partial2);
    cpsnr_argmax_kernel<<<B, 64, 0, s>>>(score, redo, slices, set_sel, partial2, partial3, g2.blocks_per_set, best_db, best_site, site_db);
    note_launches(7);
    HRN_CUDA_OK(cudaGetLastError());
    HRN_CUDA_OK(cudaFreeAsync(ws, s));
    return 0;
}

int shift_cpsnr_launch(const float* sr, const float* hr, const float* hm, int B, int H, int W, int border,
                       int clip_sr, float* best_db, int32_t* best_site, float* site_db, cudaStream_t s) {
    if (H != W) {
        set_error("shift_cpsnr: square images only (got %d x %d), like Evaluator.py:64", H, W);
        return -1;
    }
    if (border < 0 || W - 2 * border <= 0) {
        set_error("shift_cpsnr: border_w=%d leaves no crop of a %d x %d image", border, H, W);
        return -1;
    }
    if (2 * border + 1 > 255) {
        set_error("shift_cpsnr: border_w=%d: more than 255 x 255 shifts are not supported", border);
        return -1;
    }
    if (B <= 0 || B > 65535) {
        set_error("shift_cpsnr: batch %d outside [1, 65535]", B);
        return -1;
    }
    CpGeom g;
    g.H = H;
    g.W = W;
    g.border = border;
    g.S = 2 * border + 1;
    g.size = W - 2 * border;
    if (g.S > CP_MAXS) {
        // border_w > 3: no caller of the reference uses it, but Evaluator.py:52 accepts it -> one block per (site, imageset)
        g.col_blocks = g.band_rows = g.blocks_per_set = 1;
        g.vec_ok = 0;
        const int sites = g.S * g.S;
        const size_t partial_bytes = static_cast<size_t>(B) * sites * 2 * sizeof(double);
        const size_t nclear_bytes = static_cast<size_t>(B) * sites * sizeof(double);
        uint8_t* ws = nullptr;
        if (scratch_alloc(reinterpret_cast<void**>(&ws), partial_bytes + nclear_bytes + static_cast<size_t>(B) * sites * sizeof(float), s)) return -1;
        double* partial = reinterpret_cast<double*>(ws);
        double* nclear = reinterpret_cast<double*>(ws + partial_bytes);
        float* bias = reinterpret_cast<float*>(ws + partial_bytes + nclear_bytes);
        dim3 grid(sites, B);
        cpsnr_any_pass_kernel<1><<<grid, CA_THREADS, 0, s>>>(sr, hr, hm, g, clip_sr, nullptr, partial);
        cpsnr_any_finalize_kernel<1><<<B, CA_THREADS, 0, s>>>(partial, sites, bias, nclear, nullptr, nullptr, nullptr);
        cpsnr_any_pass_kernel<2><<<grid, CA_THREADS, 0, s>>>(sr, hr, hm, g, clip_sr, bias, partial);
        cpsnr_any_finalize_kernel<2><<<B, CA_THREADS, 0, s>>>(partial, sites, bias, nclear, best_db, best_site, site_db);
        note_launches(4);
        HRN_CUDA_OK(cudaGetLastError());
        HRN_CUDA_OK(cudaFreeAsync(ws, s));
        return 0;
    }
    g.vec_ok = (W % 4 == 0) && (((reinterpret_cast<uintptr_t>(sr) | reinterpret_cast<uintptr_t>(hr) | reinterpret_cast<uintptr_t>(hm)) & 15) == 0);
    if (g.S == OP_S && g.vec_ok && g_cpsnr_generic == 0 && g_cpsnr_onepass != 0 && g_cpsnr_window_v1 < 0 && g_cpsnr_chunk == 0)
        return shift_cpsnr_onepass(sr, hr, hm, B, H, W, clip_sr, best_db, best_site, site_db, s);
    // border_w = 3 on 16-byte aligned rows (every case the reference produces) takes the 49-site window kernel
    const bool window = g.S == CW_S && g.vec_ok && g_cpsnr_generic == 0;
    const int variant = g_cpsnr_window_v1 >= 0 ? g_cpsnr_window_v1 : (B <= 128 ? 2 : 1);
    const bool window2 = window && variant != 1;
    // Both passes read sr, hr and the map.  Optionally (knob cpsnr_chunk) a batch that does not fit in L2 is processed in
    // chunks of imagesets -- pass 1 -> bias -> pass 2 of one chunk back to back, so that pass 2 finds the chunk in L2 and
    // every byte comes from HBM once.  Off by default: the kernels are fp32-bound and the extra launches cost more.
    const size_t set_bytes = 3 * static_cast<size_t>(H) * W * sizeof(float);
    int chunk = B;
    if (g_cpsnr_chunk > 0) chunk = g_cpsnr_chunk;
    else if (g_cpsnr_chunk < 0 && window && set_bytes * B > CP_L2_BUDGET) chunk = static_cast<int>(CP_L2_BUDGET / set_bytes);
    chunk = chunk < 1 ? 1 : (chunk > B ? B : chunk);
    const int per_pass = chunk < B ? chunk : B;            // imagesets per launch: sizes the row bands
    if (window) {
        g.col_blocks = (g.size + CW_COLS - 1) / CW_COLS;
        // one-warp blocks, at most one full wave of them when the batch is small; a band is at least 8 rows
        const int target = window2 ? CW2_TARGET_WARPS / 2 : CW_TARGET_WARPS;
        int want = target / (per_pass * g.col_blocks);
        want = want < 1 ? 1 : (want > (g.size + 7) / 8 ? (g.size + 7) / 8 : want);
        g.band_rows = (g.size + want - 1) / want;
    } else {
        g.col_blocks = (g.size + CP_COLS - 1) / CP_COLS;
        // enough row bands to fill the GPU (224-thread blocks, ~8 per SM), at least 8 rows each, a multiple of 4 rows
        int want = (CP_TARGET_BLOCKS + B * g.col_blocks - 1) / (B * g.col_blocks);
        want = want < 1 ? 1 : (want > (g.size + 7) / 8 ? (g.size + 7) / 8 : want);
        g.band_rows = (((g.size + want - 1) / want) + 3) & ~3;
    }
    const int bands = (g.size + g.band_rows - 1) / g.band_rows;
    g.blocks_per_set = bands * g.col_blocks;
    const int sites = g.S * g.S;
    const size_t partial_bytes = static_cast<size_t>(B) * g.blocks_per_set * sites * 2 * sizeof(double);
    const size_t nclear_bytes = static_cast<size_t>(B) * sites * sizeof(double);
    const size_t bias_bytes = static_cast<size_t>(B) * sites * sizeof(float);
    uint8_t* ws = nullptr;
    if (scratch_alloc(reinterpret_cast<void**>(&ws), partial_bytes + nclear_bytes + bias_bytes, s)) return -1;
    double* partial_all = reinterpret_cast<double*>(ws);
    double* nclear_all = reinterpret_cast<double*>(ws + partial_bytes);
    float* bias_all = reinterpret_cast<float*>(ws + partial_bytes + nclear_bytes);
    const size_t plane = static_cast<size_t>(H) * W;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = b0 + chunk <= B ? chunk : B - b0;
        const float *sr_c = sr + b0 * plane, *hr_c = hr + b0 * plane, *hm_c = hm + b0 * plane;
        double* partial = partial_all + static_cast<size_t>(b0) * g.blocks_per_set * sites * 2;
        double* nclear = nclear_all + static_cast<size_t>(b0) * sites;
        float* bias = bias_all + static_cast<size_t>(b0) * sites;
        float* site_c = site_db != nullptr ? site_db + static_cast<size_t>(b0) * sites : nullptr;
        dim3 grid(g.blocks_per_set, nb), grid2(g.blocks_per_set, nb, 2), block(CP_LANES, g.S);
        if (window2 && variant == 2)
            cpsnr_window2_kernel<1, false><<<grid2, 32, 0, s>>>(sr_c, hr_c, hm_c, g, clip_sr, nullptr, partial);
        else if (window2)
            cpsnr_window2_kernel<1, true><<<grid2, 32, 0, s>>>(sr_c, hr_c, hm_c, g, clip_sr, nullptr, partial);
        else if (window)
            cpsnr_window_kernel<1><<<grid, 32, 0, s>>>(sr_c, hr_c, hm_c, g, clip_sr, nullptr, partial);
        else if (g.S == 7)
            cpsnr_pass_kernel<1, 7><<<grid, block, 0, s>>>(sr_c, hr_c, hm_c, g, clip_sr, nullptr, partial);
        else
            cpsnr_pass_kernel<1, 0><<<grid, block, 0, s>>>(sr_c, hr_c, hm_c, g, clip_sr, nullptr, partial);
        cpsnr_finalize_kernel<1><<<nb, 64, 0, s>>>(partial, g, bias, nclear, nullptr, nullptr, nullptr);
        if (window2 && variant == 2)
            cpsnr_window2_kernel<2, false><<<grid2, 32, 0, s>>>(sr_c, hr_c, hm_c, g, clip_sr, bias, partial);
        else if (window2)
            cpsnr_window2_kernel<2, true><<<grid2, 32, 0, s>>>(sr_c, hr_c, hm_c, g, clip_sr, bias, partial);
        else if (window)
            cpsnr_window_kernel<2><<<grid, 32, 0, s>>>(sr_c, hr_c, hm_c, g, clip_sr, bias, partial);
        else if (g.S == 7)
            cpsnr_pass_kernel<2, 7><<<grid, block, 0, s>>>(sr_c, hr_c, hm_c, g, clip_sr, bias, partial);
        else
            cpsnr_pass_kernel<2, 0><<<grid, block, 0, s>>>(sr_c, hr_c, hm_c, g, clip_sr, bias, partial);
        cpsnr_finalize_kernel<2><<<nb, 64, 0, s>>>(partial, g, bias, nclear, best_db + b0, best_site + b0, site_c);
        note_launches(4);
    }
    HRN_CUDA_OK(cudaGetLastError());
    HRN_CUDA_OK(cudaFreeAsync(ws, s));
    return 0;
}

int clear_loss_launch(const float* sr, const float* hr, const float* hm, int B, int H, int W, int metric, float* out,
                      cudaStream_t s) {
    if (metric < 0 || metric > 2) {
        set_error("clear_loss: metric %d unknown (0 = masked_MSE, 1 = cMSE, 2 = cPSNR)", metric);
        return -1;
    }
    if (B <= 0 || B > 65535 || H <= 0 || W <= 0) {
        set_error("clear_loss: bad sizes (B=%d H=%d W=%d)", B, H, W);
        return -1;
    }
    const size_t hw = static_cast<size_t>(H) * W;
    const size_t partial_bytes = static_cast<size_t>(B) * CL_BLOCKS * 2 * sizeof(double);
    uint8_t* ws = nullptr;
    if (scratch_alloc(reinterpret_cast<void**>(&ws), partial_bytes + B * sizeof(double) + B * sizeof(float), s)) return -1;
    double* partial = reinterpret_cast<double*>(ws);
    double* nclear = reinterpret_cast<double*>(ws + partial_bytes);
    float* bias = reinterpret_cast<float*>(ws + partial_bytes + B * sizeof(double));
    dim3 grid(CL_BLOCKS, B);
    const int fb = (B + 127) / 128;
    if (metric == 0) {
        clear_loss_pass_kernel<0><<<grid, CL_THREADS, 0, s>>>(sr, hr, hm, hw, nullptr, partial);
        clear_loss_finalize_kernel<0><<<fb, 128, 0, s>>>(partial, B, static_cast<double>(hw), metric, nullptr, nullptr, out);
        note_launches(2);
    } else {
        clear_loss_pass_kernel<1><<<grid, CL_THREADS, 0, s>>>(sr, hr, hm, hw, nullptr, partial);
        clear_loss_finalize_kernel<1><<<fb, 128, 0, s>>>(partial, B, 0.0, metric, bias, nclear, nullptr);
        clear_loss_pass_kernel<2><<<grid, CL_THREADS, 0, s>>>(sr, hr, hm, hw, bias, partial);
        clear_loss_finalize_kernel<2><<<fb, 128, 0, s>>>(partial, B, 0.0, metric, bias, nclear, out);
        note_launches(4);
    }
    HRN_CUDA_OK(cudaGetLastError());
    HRN_CUDA_OK(cudaFreeAsync(ws, s));
    return 0;
}

}  // namespace hrn
