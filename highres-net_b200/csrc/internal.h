// Internal declarations shared by the translation units of libhrn_b200.so.
// Nothing here crosses the C ABI (see include/hrn_b200.h for that).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace hrn {

// ------------------------------------------------------------------ errors
void set_error(const char* fmt, ...);
void note_launches(int n);   // feeds hrn_kernel_launch_count()
#define HRN_CUDA_OK(expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::hrn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                             __LINE__);                                                        \
            return -1;                                                                         \
        }                                                                                      \
    } while (0)

// Entry points that take a handle run on the handle's device and leave the caller's current device as they found it.
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(device) == cudaSuccess;
        if (!ok) set_error("cudaSetDevice(%d) failed", device);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// ------------------------------------------------------------------ conv3x3 (tcgen05)
// Residual / merge modes of the conv epilogue.
enum ResMode : int {
    RES_NONE = 0,
    RES_SAME = 1,    // out = y + res[m]                         (ResidualBlock skip, HRNet.py:32-33)
    RES_PAIR = 2,    // out = y + cat(alice, bob)[m]             (ResidualBlock(128) skip on the fused pair)
    RES_ALPHA = 3,   // out = alice[m] + alpha_bob[m] * y        (HRNet.py:123-128)
};

struct ConvArgs {
    // geometry of the OUTPUT images (same H, W as the input: 3x3, pad 1)
    int n_img, H, W;
    int cin, cout;            // cin in {64, 128}; cout in {64, 128}
    int max_ctas;             // 0 = one CTA per SM; tests shrink it to force odd strip boundaries
    int strip_split;          // 0/1 = one contiguous row range per CTA; k > 1 = k shorter ranges per CTA, dealt round-robin
    int debug_flags;          // perf triage only: 1 = no TMEM load, 2 = no stores, 4 = no TMA loads, 8 = no residual loads; 2048 = no L2 eviction hints
    // A operand gather.  pair_mode: output image m = (b, i) reads chunk 0 from view i and chunk 1 from
    // view top-1-i of the 64-channel view stack (HRNet.py:114-119); otherwise chunk c = channels [64c, 64c+64).
    // src_views is the STRIDE of the view stack in images per imageset (the original L at every level: the fusion
    // writes each level in place over the alice slots), top the number of views the level pairs up.
    int pair_mode, half, src_views, top;
    const void* in;           // bf16 NHWC source tensor (for the TMA map)
    int in_images, in_c;      // its image count and channels per pixel
    // B operand: pre-swizzled smem image, [cout / 64][kx][cin / 64][3 ky-blocks x 64 rows x 128 B]
    const uint8_t* w_img;
    const float* bias;        // [cout]
    float prelu;
    int has_prelu;
    // epilogue
    __nv_bfloat16* out;       // bf16 NHWC, cout channels per pixel
    int res_mode;
    const __nv_bfloat16* res; // RES_SAME: cout-channel tensor; RES_PAIR / RES_ALPHA: the 64-channel view stack
    const float* alphas;      // (B, alpha_stride) original alphas, RES_ALPHA only
    int alpha_stride;
    int out_in_stack;         // 1: output image m = (b, i) is written to slot b * src_views + i of the view stack
    // live-work list (pointwise.cu: live_lists_kernel): the kernel processes images live_list[0 .. *live_count) instead
    // of 0 .. n_img; both device pointers, nullptr = dense
    const int* live_list;
    const int* live_count;
    int mcast;                // 1: 128 -> 128 layers run as cluster pairs that multicast the A rows (see conv3x3_umma.cu)
    int pool;                 // 1: MaxPool2d(2) fused into the epilogue; `out` is (n_img, H / 2, W / 2, cout).  Plain layers only.
    int no_img_group;         // test knob: 1 = never put several narrow images into one M tile (see conv3x3_launch)
};
int conv3x3_bytes_per_weight_image(int cin, int cout);
// Repack OIHW fp32 (cout, cin, 3, 3) host weights into the pre-swizzled bf16 smem image (host memory).
void conv3x3_pack_weights(const float* oihw, int cin, int cout, uint8_t* dst);
int conv3x3_launch(const ConvArgs& a, int sm_count, cudaStream_t stream);
// resblock64_umma.cu: x + PReLU2(conv2(PReLU1(conv1(x)))) for 64 channels in one launch (the intermediate stays in shared
// memory).  a1 = conv 1 (x = a1.in, out = a1.out).  Returns 1 when the shape is not supported (W > 128): use two launches.
// bias1_host / bias2_host: HOST pointers to the 64 biases of each conv (they travel as kernel parameters).
int resblock64_launch(const ConvArgs& a1, const float* bias1_host, const uint8_t* w2_img, const float* bias2_host,
                      float prelu2, int sm_count, cudaStream_t stream);

// fuse_wave_umma.cu: one fusion level (conv A, conv B + skip, conv C + PReLU + alpha merge) as a single wavefront launch.
struct FuseWaveArgs {
    int H, W;
    int half, src_views, top;          // as in ConvArgs: pair (b, i) = views i and top - 1 - i; src_views = view stack stride
    const int* live_list;              // live pairs of this level and their count (device)
    const int* live_count;
    const int* carry_list;             // dead pairs whose alice the next level needs (device)
    const int* carry_count;
    const __nv_bfloat16* stack_in;     // (stack_images, H, W, 64) bf16
    int stack_images;
    __nv_bfloat16* stack_out;          // same shape; merged pairs and carried views are written to their alice slot
    __nv_bfloat16* ring1;              // fuse_wave_ring_bytes() each
    __nv_bfloat16* ring2;
    int ring_rows;
    uint32_t* flags;                   // fuse_wave_flag_bytes(), all zero when the launch starts
    const float* alphas;
    int alpha_stride, alpha_residual;
    const uint8_t* w_img[3];           // conv A, B, C: conv3x3_pack_weights images
    const float* bias[3];
    float prelu[3];
    int has_prelu[3];
    int debug_flags;
    int streams;                       // 0 = as many as the SMs allow (sm_count / 5); tests force odd partitions with fewer
    unsigned long long* stats;         // optional triage counters: 8 per CTA (see fuse_wave_umma.cu), accumulated over launches
    int publish_rows;                  // rows per hand-over publication (one device-scope release each); 0 = 1
    int lag_rows;                      // triage: consumers stay this many rows behind their producer
};
int fuse_wave_streams(int sm_count);
bool fuse_wave_fits(int sm_count);        // every CTA of a wavefront grid can be resident at once on the current device
size_t fuse_wave_ring_bytes(int sm_count, int ring_rows, int W);
size_t fuse_wave_flag_bytes(int sm_count);
int fuse_wave_launch(const FuseWaveArgs& a, int sm_count, cudaStream_t stream);

// enc_wave_umma.cu: the encoder's two ResidualBlocks and its final conv (five 64 -> 64 convolutions) as one wavefront launch.
struct EncWaveLaunch {
    int H, W;
    int n_img;                         // images of x0 / out (B * L)
    const int* live_list;              // live views and their count (device)
    const int* live_count;
    const __nv_bfloat16* x0;           // (n_img, H, W, 64) bf16: output of the first conv
    __nv_bfloat16* out;                // (n_img, H, W, 64) bf16: the view stack
    __nv_bfloat16* ring[4];            // enc_wave_ring_bytes() each
    int ring_rows;
    uint32_t* flags;                   // enc_wave_flag_bytes(), all zero when the launch starts
    const uint8_t* w_img[5];           // R0a, R0b, R1a, R1b, FIN: conv3x3_pack_weights images
    const float* bias[5];
    float prelu[5];
    int has_prelu[5];
    int debug_flags;
    int streams;                       // 0 = sm_count / 5
    unsigned long long* stats;         // optional triage counters, 12 per CTA (knob "enc_stats")
};
int enc_wave_streams(int sm_count);
bool enc_wave_fits(int sm_count);
size_t enc_wave_ring_bytes(int sm_count, int ring_rows, int W);
size_t enc_wave_flag_bytes(int sm_count);
int enc_wave_launch(const EncWaveLaunch& a, int sm_count, cudaStream_t stream);

// ------------------------------------------------------------------ pointwise / CUDA-core kernels
int median_anchor_launch(const float* lrs, int B, int L, int H, int W, float* anchor, cudaStream_t s);
// conv_init_umma.cu: conv 2->64 + PReLU on (view, anchor) pairs on the tensor cores (A operand built in smem with a
// hi/lo bf16 split of the fp32 inputs); writes bf16 NHWC (B*L, H, W, 64).  w_img = conv_init_pack_weights() (device).
int conv_init_weight_image_bytes();
void conv_init_pack_weights(const float* w_co_ci_ky_kx, uint8_t* dst);
int conv_init_umma_launch(const float* lrs, const float* anchor, int B, int L, int H, int W, const uint8_t* w_img,
                          const float* bias, float prelu, __nv_bfloat16* out, const int* live_list, const int* live_count,
                          int sm_count, cudaStream_t s);
// decoder_umma.cu: stride-3 deconv + PReLU + 1x1 conv fused on the tensor cores; in: bf16 NHWC (B, H, W, 64) ->
// out fp32 (B, 3H, 3W).  w_img = decoder_pack_weights() image (device), bd (64), wf (64) fp32 device.
int decoder_weight_image_bytes();
void decoder_pack_weights(const float* w_ci_co_ky_kx, uint8_t* dst);
// image_stride: imageset b is image b * image_stride of `in` (view 0 of the in-place view stack).
int decoder_umma_launch(const __nv_bfloat16* in, int B, int image_stride, int H, int W, const uint8_t* w_img, const float* bd,
                        float prelu, const float* wf, float bf, float* out, int sm_count, cudaStream_t s);
int nhwc_bf16_to_nchw_f32_launch(const __nv_bfloat16* in, int n, int H, int W, int C, int group, int stride, float* out,
                                 cudaStream_t s);
int u16_to_unit_float_launch(const uint16_t* in, size_t n, float* out, cudaStream_t s);
int unit_float_to_u16_launch(const float* in, size_t n, uint16_t* out, int* bad, cudaStream_t s);
int collate_launch(const void* packed, int is_u16, const int* offsets, int B, int min_L, int H, int W, float* lrs,
                   float* alphas, cudaStream_t s);
// live-work lists (see pointwise.cu).  Layout of `lists`: [0] = live encoder views, [1 + l] = live pairs of level l,
// [16 + l] = carried pairs of level l (dead pairs whose alice the next level needs); [LIVE_HDR, LIVE_HDR + B*L) encoder
// list, then the pair lists level by level; the carry lists start at LIVE_HDR + 2*B*L, level by level.
constexpr int LIVE_HDR = 32;
int live_levels(int L);
size_t live_scratch_bytes(int B, int L);
size_t live_lists_ints(int B, int L);
int live_lists_launch(const float* alphas, int B, int L, int skip, int alpha_residual, uint8_t* scratch, int* lists,
                      cudaStream_t s);

// ------------------------------------------------------------------ scoring
int lanczos_shift_launch(const float* img, const float* shift, int nb, int c, int H, int W, int p, int a,
                         int ntaps, float* out, cudaStream_t s);
int lanczos_taps_launch(const float* d, int n, int a, int ntaps, float* out, cudaStream_t s);
// lanczos7_tma.cu: the N = 7 kernel for rows that TMA can address (W % 4 == 0, 16-byte aligned pointers)
bool lanczos7_tma_usable(const float* img, const float* out, int H, int W);
int lanczos7_tma_launch(const float* img, const float* taps_dev, int planes, int c, int H, int W, int p, float* out, cudaStream_t s);
int clear_loss_launch(const float* sr, const float* hr, const float* hm, int B, int H, int W, int metric, float* out,
                      cudaStream_t s);
extern int g_cpsnr_generic, g_cpsnr_window_v1, g_cpsnr_chunk, g_cpsnr_onepass, g_lanczos_scalar;   // test knobs (hrn_scoring_debug_set)
int shift_cpsnr_launch(const float* sr, const float* hr, const float* hm, int B, int H, int W, int border,
                       int clip_sr, float* best_db, int32_t* best_site, float* site_db, cudaStream_t s);

}  // namespace hrn
