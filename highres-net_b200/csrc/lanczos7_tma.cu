// N = 7 Lanczos shift for rows that TMA can address (W % 4 == 0, 16-byte aligned base): lanczos.py:47-107 with the
// kernel width the reference uses everywhere (ShiftNet.py:87-89).
//
// ncu on the register-window kernel in scoring.cu (4.6 TB/s at 512 x 384^2, DRAM 54 %) showed issue slots 77 % active at
// ~35 instructions per output pixel, of which 14 are the FFMAs of the two passes: bound by instruction issue, two thirds of
// the rest being scalar loads and their addresses.  A first TMA version with CTA-wide tiles (one 136 x 38 window per 288
// threads, two __syncthreads per tile) halved the instructions but spent its time at the barriers (4.8 TB/s); a second one
// with warp-private tiles and the y-filtered rows handed to the x pass through shared memory ran the LSU data pipe at 68 %
// (28 bytes of shared-memory traffic per output pixel).  This one has NO block-level synchronisation and no intermediate in
// shared memory: every WARP owns a pipeline of its own --
//   * a sub-tile is 56 x 16 outputs; its (56 + 8) x (16 + 6) input window is ONE instruction (a 3-D TMA box, rows and
//     columns outside the image arrive as zeros) into one of the warp's two window buffers, the next window always in flight;
//   * lane = (group of four columns, 8-row segment).  y pass (dim 0 taps first, lanczos.py:90): 14 LDS.128 for 8 x 4 values,
//     which stay in registers; x pass (lanczos.py:94): the three columns either side come from the neighbouring lanes
//     (six shuffles per row), then 8 x 4 outputs and eight 128-bit stores.  The two groups at the window's edge are halo only;
//   * the reflect ring (ReflectionPad2d, lanczos.py:83) is a row remap in the y pass of the sub-tiles that touch the top /
//     bottom of the image, and a register copy in the x pass of the lanes that hold column 0 / W - 1 (filtering a column
//     commutes with copying it), so nothing is patched in shared memory.
// Arithmetic order per output is the one of lanczos_shift7_kernel (taps ascending, fmaf): both kernels agree bit for bit.
#include "umma_common.cuh"

namespace hrn {
namespace {

constexpr int T_W = 56, T_H = 16, HALF = 3;
constexpr int IN_W = T_W + 8, IN_H = T_H + 2 * HALF;            // window: columns x0 - 4 .. x0 + 59 (16-byte aligned), rows y0 - 3 .. y0 + 18
constexpr int GROUPS = IN_W / 4, SEG = 8;                       // y pass: 16 column groups x 2 row segments = 32 lanes
constexpr int WARPS = 8, THREADS = WARPS * 32, CTAS_PER_SM = 2; // 16 independent warp pipelines per SM (10 x 2 and 6 x 3 at <= 96 registers measured 15-30 % slower)
constexpr uint32_t IN_BYTES = IN_W * IN_H * sizeof(float);      // 5632
static_assert(GROUPS * (T_H / SEG) == 32 && GROUPS == 16, "lane mapping");

struct __align__(128) WarpSmem {                                // a TMA destination must start on a 128-byte boundary
    float in[2][IN_H][IN_W];
};
static_assert(sizeof(WarpSmem) % 128 == 0 && IN_BYTES % 128 == 0, "window alignment");
struct __align__(128) Smem {
    WarpSmem w[WARPS];
    uint64_t full[WARPS][2];
};

struct L7Args {
    const float* taps;      // [(c * 2 + axis) * 7 + t]
    float* out;
    int C, H, W, p, tiles_x, tiles_y, tiles;
};

// Tile cursor: tile index -> (tx, ty, plane) once, then advanced by a constant step without divisions.
struct Cursor {
    int tx, ty, plane;
    __device__ __forceinline__ void set(int tile, const L7Args& a) {
        tx = tile % a.tiles_x;
        const int rest = tile / a.tiles_x;
        ty = rest % a.tiles_y;
        plane = rest / a.tiles_y;
    }
    __device__ __forceinline__ void advance(const Cursor& step, const L7Args& a) {
        tx += step.tx;
        ty += step.ty;
        plane += step.plane;
        if (tx >= a.tiles_x) { tx -= a.tiles_x; ++ty; }
        if (ty >= a.tiles_y) { ty -= a.tiles_y; ++plane; }
    }
};

// Row of the window that holds image row gy of the reflect-padded image (lanczos.py:83), or a row of zeros: rows outside
// the image arrive zero-filled, so a row beyond the p-wide ring simply maps to itself.
__device__ __forceinline__ int window_row(int wy, int y0, int H, int p) {
    const int gy = y0 - HALF + wy;
    int r = gy;
    if (gy < 0 && -gy <= p) r = -gy;
    if (gy >= H && gy - H < p) r = 2 * (H - 1) - gy;
    return min(max(r - y0 + HALF, 0), IN_H - 1);
}

template <bool EDGE_Y>
__device__ __forceinline__ void y_pass(const float (*in)[IN_W], float4 (&acc)[SEG], const float (&ky)[7], int g, int r0, int y0,
                                       int H, int p) {
#pragma unroll
    for (int r = 0; r < SEG; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < SEG + 6; ++t) {
        const int wy = EDGE_Y ? window_row(r0 + t, y0, H, p) : r0 + t;
        const float4 v = *reinterpret_cast<const float4*>(&in[wy][4 * g]);
#pragma unroll
        for (int r = 0; r < SEG; ++r) {
            const int k = t - r;                                 // tap of window row r0 + t for output row r0 + r: ascending in t
            if (k >= 0 && k < 7) {
                acc[r].x = fmaf(ky[k], v.x, acc[r].x);
                acc[r].y = fmaf(ky[k], v.y, acc[r].y);
                acc[r].z = fmaf(ky[k], v.z, acc[r].z);
                acc[r].w = fmaf(ky[k], v.w, acc[r].w);
            }
        }
    }
}

__global__ void __launch_bounds__(THREADS, CTAS_PER_SM)
lanczos7_tma_kernel(const __grid_constant__ CUtensorMap map, const L7Args a) {
    extern __shared__ __align__(128) uint8_t l7_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(l7_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSmem& ws = sm.w[warp];
    const int n_pipes = static_cast<int>(gridDim.x) * WARPS, pipe = static_cast<int>(blockIdx.x) * WARPS + warp;
    const int n_my = (a.tiles - pipe + n_pipes - 1) / n_pipes;  // this warp's sub-tiles: pipe, pipe + n_pipes, ...
    if (n_my <= 0) return;
    const uint32_t bar0 = ptx::smem_u32(&sm.full[warp][0]), bar1 = ptx::smem_u32(&sm.full[warp][1]);
    Cursor step, cur, ahead;                                    // ahead: the sub-tile whose window is requested next (lane 0)
    step.set(n_pipes, a);
    cur.set(pipe, a);
    ahead = cur;
    auto issue = [&](int i) {                                   // one lane: window of my i-th sub-tile into buffer i & 1
        const uint32_t bar = (i & 1) ? bar1 : bar0;
        ptx::mbar_expect_tx(bar, IN_BYTES);
        ptx::tma_load_3d(ptx::smem_u32(&ws.in[i & 1][0][0]), &map, ahead.tx * T_W - 4, ahead.ty * T_H - HALF, ahead.plane, bar);
        ahead.advance(step, a);
    };
    if (lane == 0) {
        ptx::prefetch_tensormap(&map);
        ptx::mbar_init(bar0, 1);
        ptx::mbar_init(bar1, 1);
        ptx::fence_barrier_init();
        issue(0);
        if (n_my > 1) issue(1);
    }
    __syncwarp();
    const int g = lane % GROUPS, r0 = (lane / GROUPS) * SEG;    // (group of four columns, segment of eight output rows)
    // taps of a sub-tile's channel: 14 floats, ky[7] then kx[7] (lanczos_taps_kernel).  They are fetched one sub-tile ahead,
    // underneath the x pass, so the y pass never waits for them.
    auto load_taps = [&](int plane, float (&k14)[14]) {
        const float2* tp = reinterpret_cast<const float2*>(a.taps) + (plane % a.C) * 7;
#pragma unroll
        for (int t = 0; t < 7; ++t) {
            const float2 v = __ldg(tp + t);
            k14[2 * t] = v.x;
            k14[2 * t + 1] = v.y;
        }
    };
    ptx::pdl_wait();                                            // the taps come from the kernel launched just before this one
    float nk[14];
    load_taps(cur.plane, nk);
    for (int i = 0; i < n_my; ++i) {
        const int b = i & 1;
        const int x0 = cur.tx * T_W, y0 = cur.ty * T_H, plane = cur.plane;
        cur.advance(step, a);
        float ky[7], kx[7];
#pragma unroll
        for (int t = 0; t < 7; ++t) {
            ky[t] = nk[t];
            kx[t] = nk[7 + t];
        }
        ptx::mbar_wait(b ? bar1 : bar0, (i >> 1) & 1, 1);
        float4 yv[SEG];                                          // y-filtered columns x .. x + 3 of rows y0 + r0 .. + 7
        if ((y0 - HALF < 0) || (y0 + T_H + HALF > a.H)) y_pass<true>(ws.in[b], yv, ky, g, r0, y0, a.H, a.p);
        else y_pass<false>(ws.in[b], yv, ky, g, r0, y0, a.H, a.p);
        __syncwarp();                                            // in[b] free
        if (lane == 0 && i + 2 < n_my) {
            ptx::fence_proxy_async_shared();                     // the warp's ld.shared of in[b] before the async-proxy refill
            issue(i + 2);
        }
        if (i + 1 < n_my) load_taps(cur.plane, nk);
        const int x = x0 - 4 + 4 * g;
        // W % 4 == 0: the image ends at a lane boundary.  Columns -1, -2, -3 mirror 1, 2, 3 and W, W + 1, W + 2 mirror
        // W - 2, W - 3, W - 4 inside the p-wide ring; beyond it they stay zero.
        const bool left = x == 0, right = x + 4 == a.W;
        const bool l1 = left && a.p >= 1, l2 = left && a.p >= 2, l3 = left && a.p >= 3;
        const bool r1 = right && a.p >= 1, r2 = right && a.p >= 2, r3 = right && a.p >= 3;
        // rows of this lane's segment it has to store: none for the two halo groups and for columns beyond the image
        const int rows = (g >= 1 && g <= GROUPS - 2 && x < a.W) ? a.H - y0 - r0 : 0;
        float* dst = a.out + (static_cast<size_t>(plane) * a.H + y0 + r0) * a.W + x;
#pragma unroll
        for (int r = 0; r < SEG; ++r) {
            float v[10];                                         // v[j] = column x - 3 + j of the y-filtered row
            v[0] = __shfl_up_sync(0xffffffffu, yv[r].y, 1);
            v[1] = __shfl_up_sync(0xffffffffu, yv[r].z, 1);
            v[2] = __shfl_up_sync(0xffffffffu, yv[r].w, 1);
            v[3] = yv[r].x;
            v[4] = yv[r].y;
            v[5] = yv[r].z;
            v[6] = yv[r].w;
            v[7] = __shfl_down_sync(0xffffffffu, yv[r].x, 1);
            v[8] = __shfl_down_sync(0xffffffffu, yv[r].y, 1);
            v[9] = __shfl_down_sync(0xffffffffu, yv[r].z, 1);
            v[2] = l1 ? v[4] : v[2];
            v[1] = l2 ? v[5] : v[1];
            v[0] = l3 ? v[6] : v[0];
            v[7] = r1 ? v[5] : v[7];
            v[8] = r2 ? v[4] : v[8];
            v[9] = r3 ? v[3] : v[9];
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float acc = 0.0f;
#pragma unroll
                for (int t = 0; t < 7; ++t) acc = fmaf(kx[t], v[e + t], acc);
                o[e] = acc;
            }
            // predicated store, no branch: the rows of a lane stay one basic block that the scheduler can interleave
            asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.s32 q, %5, %6;\n\t@q st.global.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
                         ::"l"(dst + static_cast<size_t>(r) * a.W), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]), "r"(r), "r"(rows)
                         : "memory");
        }
    }
}

}  // namespace

bool lanczos7_tma_usable(const float* img, const float* out, int H, int W) {
    return (W % 4 == 0) && W >= 8 && H >= 1 &&
           (((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
}

// taps: device array [(c * 2 + axis) * 7 + t] from lanczos_taps_kernel; planes = nb * c images of H x W.
int lanczos7_tma_launch(const float* img, const float* taps, int planes, int c, int H, int W, int p, float* out, cudaStream_t s) {
    int dev = 0, sm_count = 0;
    HRN_CUDA_OK(cudaGetDevice(&dev));
    HRN_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    EncodeTiledFn encode = get_encode_fn();
    if (encode == nullptr) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return -1;
    }
    CUtensorMap map;
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(planes)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(W) * 4, static_cast<cuuint64_t>(H) * W * 4};
    const cuuint32_t box[3] = {IN_W, IN_H, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(img), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("lanczos_shift: cuTensorMapEncodeTiled failed with CUresult %d (%d planes of %d x %d)", (int)r, planes, H, W);
        return -1;
    }
    L7Args a;
    a.taps = taps;
    a.out = out;
    a.C = c;
    a.H = H;
    a.W = W;
    a.p = p;
    a.tiles_x = (W + T_W - 1) / T_W;
    a.tiles_y = (H + T_H - 1) / T_H;
    const long long tiles = static_cast<long long>(a.tiles_x) * a.tiles_y * planes;
    if (tiles > 0x7fffffffLL) {
        set_error("lanczos_shift: %lld tiles exceed the 31-bit tile index; split the batch", tiles);
        return -1;
    }
    a.tiles = static_cast<int>(tiles);
    static bool attr_set[64] = {};
    if (allow_dynamic_smem(lanczos7_tma_kernel, static_cast<int>(sizeof(Smem)), attr_set)) return -1;
    const int resident = sm_count * CTAS_PER_SM, wanted = (a.tiles + WARPS - 1) / WARPS;
    const int ctas = wanted < resident ? wanted : resident;
    HRN_CUDA_OK(launch_pdl(lanczos7_tma_kernel, ctas, THREADS, sizeof(Smem), s, 1, map, a));
    HRN_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hrn
