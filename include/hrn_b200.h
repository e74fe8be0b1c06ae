/* hrn_b200 -- C ABI of the B200-native HighRes-net inference + scoring path.
 *
 * The reference (gwall-ceres/HighRes-net) has no FFI: its boundary is the Python
 * module API.  Each entry point below names the reference interface it replaces;
 * INTEGRATION.md shows the ctypes binding a maintainer would add on the reference
 * side.  Conventions: plain pointers and sizes only, status-code returns (0 = ok,
 * non-zero = error, message via hrn_last_error()), no exceptions across the ABI,
 * caller-owned buffers, handle-owned weights and workspace, all work enqueued on
 * the CUDA stream passed in (a cudaStream_t cast to void*; NULL = default stream).
 * One handle per device; a handle is not thread-safe.  Functions that take a handle run on
 * the handle's device and restore the caller's current device before they return; the
 * handle-less scoring functions run on the current device.  There is NO CPU fallback:
 * every function fails with an error if no sm_100 device is usable.
 */
#ifndef HRN_B200_H
#define HRN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HRN_ABI_VERSION 1

typedef struct hrn_handle hrn_handle;

/* Mirrors config/config.json:9-35 ("network") of the reference.  The kernels are
 * specialised for the shipped values; anything else is rejected by hrn_create. */
typedef struct hrn_config {
    int32_t enc_in_channels;   /* network.encoder.in_channels     (2)  */
    int32_t enc_num_layers;    /* network.encoder.num_layers      (2)  residual blocks */
    int32_t enc_kernel_size;   /* network.encoder.kernel_size     (3)  */
    int32_t enc_channels;      /* network.encoder.channel_size    (64) */
    int32_t rec_alpha_residual;/* network.recursive.alpha_residual (1) */
    int32_t rec_in_channels;   /* network.recursive.in_channels   (64) */
    int32_t rec_kernel_size;   /* network.recursive.kernel_size   (3)  */
    int32_t dec_in_channels;   /* network.decoder.deconv.in_channels  (64) */
    int32_t dec_kernel_size;   /* network.decoder.deconv.kernel_size  (3)  */
    int32_t dec_stride;        /* network.decoder.deconv.stride       (3)  */
    int32_t dec_out_channels;  /* network.decoder.deconv.out_channels (64) */
    int32_t fin_in_channels;   /* network.decoder.final.in_channels   (64) */
    int32_t fin_kernel_size;   /* network.decoder.final.kernel_size   (1)  */
    int32_t fin_out_channels;  /* network.decoder.final.out_channels  (1)  */
} hrn_config;

/* ABI version of the loaded library (== HRN_ABI_VERSION of the header it was built from). */
int32_t hrn_abi_version(void);

/* Last error message of the calling thread ("" if none). */
const char* hrn_last_error(void);

/* Replaces HRNet.__init__ (src/DeepNetworks/HRNet.py:175-184) + .to(device).
 * Creates a handle on CUDA device `device`; fails unless it is compute capability 10.x. */
int32_t hrn_create(const hrn_config* cfg, int32_t device, hrn_handle** out);
void hrn_destroy(hrn_handle* h);

/* Replaces load_state_dict (predict.py:98-99): upload one tensor by its reference
 * state_dict key (e.g. "fuse.fuse.0.block.2.weight").  `data` is
 * HOST fp32 in the reference layout (Conv2d OIHW, ConvTranspose2d (in, out, kH, kW),
 * PReLU (1,)); the handle repacks it (bf16, pre-swizzled UMMA B-operand images). */
int32_t hrn_set_weight(hrn_handle* h, const char* key, const float* data, const int64_t* shape, int32_t ndim);
/* Number of the 31 expected tensors still missing (0 = ready to run). */
int32_t hrn_missing_weights(const hrn_handle* h);

/* Replaces HRNet.forward (HRNet.py:186-211).  DEVICE pointers:
 *   lrs (B, L, H, W) fp32 contiguous, alphas (B, L) fp32, sr (B, 1, 3H, 3W) fp32 out.
 * Square inputs only (H == W), as the reference's view() at HRNet.py:204 requires.  The handle keeps 5 activation
 * buffers of B*L*H*W*128 bytes (bf16 NHWC), capped at 64 GiB by slicing the batch.
 * Views and view pairs that cannot reach the output because of alpha = 0 (HRNet.py:123-128: alice + alpha_bob * x) are
 * not computed; the result is the one the reference produces for finite inputs.  Everything is enqueued on `stream`,
 * nothing synchronises with the host. */
int32_t hrn_forward(hrn_handle* h, const float* lrs, const float* alphas, int32_t B, int32_t L, int32_t H,
                    int32_t W, float* sr, void* stream);
/* Sizes the handle's workspace for forwards of up to (B, L, H, W) now, so that no later hrn_forward* call has to free and
 * reallocate it (which synchronises the device in the middle of the caller's stream).  Optional: without it the first
 * call of a larger shape grows the workspace itself.  Replaces nothing in the reference (PyTorch's caching allocator
 * plays this role behind HRNet.forward, HRNet.py:186-211). */
int32_t hrn_reserve(hrn_handle* h, int32_t B, int32_t L, int32_t H, int32_t W);

/* Same with HOST buffers (the train.py:200-208 / predict.py:36-40 pattern: H2D of
 * lrs/alphas, forward, D2H of sr), synchronous.  The batch is cut into chunks whose copies overlap the
 * kernels of the neighbouring chunks (pinned host memory is needed for that overlap; pageable works too). */
int32_t hrn_forward_host(hrn_handle* h, const float* lrs_host, const float* alphas_host, int32_t B, int32_t L,
                         int32_t H, int32_t W, float* sr_host, void* stream);

/* The same H2D -> forward -> D2H pattern for a LOOP over batches (the validation loop train.py:199-208, where the
 * DataLoader hands over pinned batches, train.py:276-287 pin_memory=True): submit returns once copies and kernels are
 * enqueued and hands back a ticket; wait blocks until that call's sr_host is complete.  Two calls may be in flight, so
 * the H2D copy of batch n + 1 and the D2H copy of batch n - 1 overlap the kernels of batch n; a third submit first
 * retires the oldest call.  lrs_host / alphas_host / sr_host must stay valid (and unmodified / unread) until the wait;
 * pinned memory is needed for the overlap.  All submits of a handle should use the same `stream`. */
int32_t hrn_forward_host_submit(hrn_handle* h, const float* lrs_host, const float* alphas_host, int32_t B, int32_t L,
                                int32_t H, int32_t W, float* sr_host, void* stream, int64_t* ticket);
int32_t hrn_forward_host_wait(hrn_handle* h, int64_t ticket);

/* hrn_forward_host for views still in the on-disk 16-bit format (DataLoader.py:134, 195-198: io.imread -> uint16,
 * skimage.img_as_float(...).astype(float32) = x / 65535): lrs_host is (B, L, H, W) uint16, copied as is (half the H2D
 * bytes) and converted on the device with the same rounding.  hrn_u16_to_unit_float is that conversion alone on
 * DEVICE pointers (n elements). */
int32_t hrn_forward_host_u16(hrn_handle* h, const uint16_t* lrs_host, const float* alphas_host, int32_t B, int32_t L,
                             int32_t H, int32_t W, float* sr_host, void* stream);
int32_t hrn_u16_to_unit_float(const uint16_t* src, int64_t n, float* dst, void* stream);
/* Replaces utils.collateFunction (src/utils.py:63-113) on the device.  `packed` (DEVICE) holds only the real views of the
 * batch, imageset after imageset: views [offsets[b], offsets[b + 1]) belong to imageset b (`offsets`: B + 1 int32 on the
 * DEVICE; float32 planes of H * W, or raw uint16 planes when packed_is_u16 != 0, which are scaled like
 * DataLoader.py:195-198).  Writes lrs (B, min_L, H, W) float32: the first min(count_b, min_L) views of every imageset
 * (truncation, utils.py:89-91), zero planes after them (utils.py:92-95), and alphas (B, min_L) = 1 for real views, 0 for
 * padding -- the two arguments of hrn_forward.  Only real views ever cross PCIe. */
int32_t hrn_collate(const void* packed, int32_t packed_is_u16, const int32_t* offsets, int32_t B, int32_t min_L, int32_t H,
                    int32_t W, float* lrs, float* alphas, void* stream);
/* The way out (predict.py:176, generate_submission_file: sr = skimage.img_as_uint(sr) before io.imsave): float32 in
 * [-1, 1] -> uint16 = clip(rint(x * 65535 in fp32), 0, 65535), round half to even, on DEVICE pointers (n elements).
 * *out_of_range (device int32, may be NULL) is set to 1 if any value lies outside [-1, 1] (skimage raises there). */
int32_t hrn_unit_float_to_u16(const float* src, int64_t n, uint16_t* dst, int32_t* out_of_range, void* stream);

/* Replaces lanczos.lanczos_shift (src/lanczos.py:47-107) incl. lanczos_kernel (5-43).
 * DEVICE pointers: img (Nb, C, H, W) fp32, shift (C, 2) = (dy, dx) per channel, out like img.
 * p = reflect padding width, a = lobes, ntaps = N (odd). */
int32_t hrn_lanczos_shift(const float* img, const float* shift, int32_t Nb, int32_t C, int32_t H, int32_t W,
                          int32_t p, int32_t a, int32_t ntaps, float* out, void* stream);

/* Replaces lanczos.lanczos_kernel (src/lanczos.py:5-43).  DEVICE pointers: d (n) fp32 shifts ->
 * taps (n, ntaps) fp32, normalised; a = lobes, ntaps = N (odd). */
int32_t hrn_lanczos_taps(const float* d, int32_t n, int32_t a, int32_t ntaps, float* taps, void* stream);

/* Replaces Evaluator.shift_cPSNR (src/Evaluator.py:52-73, cPSNR 11-43) for a batch.
 * DEVICE pointers: sr, hr, hr_map (B, H, W) fp32 (H == W); best_db (B) fp32 = max cPSNR,
 * best_site (B) int32 = argmax over itertools.product(range(S), range(S)) with S = 2*border_w+1
 * (site = x*S + y, x = row offset), site_db (B, S*S) fp32 or NULL.  clip_sr != 0 fuses the
 * np.clip(sr, 0, 1) of the call sites (train.py:212, predict.py:43). */
int32_t hrn_shift_cpsnr(const float* sr, const float* hr, const float* hr_map, int32_t B, int32_t H, int32_t W,
                        int32_t border_w, int32_t clip_sr, float* best_db, int32_t* best_site, float* site_db,
                        void* stream);

/* Replaces train.get_loss (src/train.py:66-87) without autograd: one value per image.  DEVICE pointers: sr, hr, hr_map
 * (B, H, W) fp32, loss (B) fp32.  metric 0 = masked_MSE, 1 = cMSE (brightness-bias corrected, weighted by hr_map),
 * 2 = cPSNR = -10 log10(cMSE). */
#define HRN_LOSS_MASKED_MSE 0
#define HRN_LOSS_CMSE 1
#define HRN_LOSS_CPSNR 2
int32_t hrn_clear_loss(const float* sr, const float* hr, const float* hr_map, int32_t B, int32_t H, int32_t W,
                       int32_t metric, float* loss, void* stream);

/* ---- test / profiling hooks (not part of the reference surface) ---- */

/* Stage identifiers for hrn_forward_dump. */
#define HRN_STAGE_ANCHOR 1            /* (B, 1, H, W)  median anchor                       */
#define HRN_STAGE_ENC(i) (0x100 + (i))/* (B*L, 64, H, W) after encoder conv i (0 = init, .. last = encoder out) */
#define HRN_STAGE_FUSE(level, j) (0x200 + 4 * (level) + (j))
                                      /* j = 0: (B*half, 128, H, W) after block conv 1; 1: after the residual block;
                                         2: (B*half, 64, H, W) merged views of the next level */
/* Runs the forward like hrn_forward and additionally copies the named intermediate,
 * converted to fp32 NCHW, into `dump` (device, caller-sized).  Returns -1 if the stage does not exist.
 * This hook always computes every view and pair (no dead-view skipping), so all intermediates are defined. */
int32_t hrn_forward_dump(hrn_handle* h, const float* lrs, const float* alphas, int32_t B, int32_t L, int32_t H,
                         int32_t W, float* sr, int32_t stage, float* dump, void* stream);

/* Test knobs.  "max_ctas" = N > 0 limits the tcgen05 conv kernels to N CTAs (0 = one per SM), which moves
 * the strip boundaries of the row partition; "strip_split" = k > 1 gives every CTA k shorter row ranges dealt round-robin
 * instead of one contiguous range (measured slower: tools/strip_split.py); results must not change.  "host_chunks" = pipeline depth of
 * hrn_forward_host (0 = automatic, 1 = no overlap).  "workspace_mb" caps the activation workspace (default 65536 MB);
 * batches that need more are run as consecutive slices with identical results.  "skip_dead_views" = 0 makes the forward
 * compute every view and pair even when alpha = 0 padding keeps it from reaching the output (default 1: skipped; the
 * super-resolved image is bit-identical either way).  "mcast" = 0 launches the 128 -> 128 convolutions of the fusion stage
 * as independent CTAs instead of cluster pairs that multicast their input rows, "fuse_resblock" = 0 runs an encoder
 * ResidualBlock as two launches instead of one (defaults 1; bit-identical either way).  "fuse_wave" = 0 runs every fusion
 * level as three launches of the conv kernel instead of one wavefront launch (default 1 for images up to 128 pixels wide),
 * "enc_wave" = 1 runs the encoder's two ResidualBlocks and final conv as one wavefront launch (default 0: measured slower);
 * "wave_streams" (0 = SM count / 5), "wave_ring_rows" (16), "enc_ring_rows" (24), "wave_publish_rows" (1) and
 * "wave_lag_rows" (0) move the stream partition, ring depth and hand-over granularity of the wavefront launches -- all
 * bit-identical.  "wave_stats" / "enc_stats" = 1 start per-CTA wait counters of the wavefront kernels, = 0 print their per-role
 * means to stderr.  "debug_flags" disables parts of the conv / wavefront kernels for performance triage (results are then
 * garbage): 1 no TMEM load, 2 no stores, 4 no TMA loads, 8 no residual loads, 32 no hand-over waits, 64 relaxed publication,
 * 128 relaxed polls + fence instead of acquire polls, 256 no proxy fence, 1024 every epilogue warp polls global memory; and two
 * that keep results intact: 2048 no L2 eviction hints in the ResidualBlock / wavefront kernels, 4096 none on the wavefront's ring reads. */
int32_t hrn_debug_set(hrn_handle* h, const char* knob, int32_t value);
/* Process-wide test knobs of the scoring entry points (they take no handle).  hrn_shift_cpsnr with border_w = 3 on 16-byte
 * aligned rows runs the one-pass kernel (centred sums, per-site trust test, two-pass fallback for flagged sites);
 * "cpsnr_onepass" = 0 selects the two-pass window kernels instead, and so does any explicit "cpsnr_window_v1" / "cpsnr_chunk".
 * "cpsnr_generic" = 1 uses the general shift-window kernel also for border_w = 3; "cpsnr_window_v1" picks the two-pass window
 * kernel variant: 1 = scalar, all 49 sites per warp; 2 = the 7 row shifts split over two warps, scalar; 0 = split + packed
 * fp32x2 (measured slowest); default -1 = automatic (2 for batches up to 128 imagesets, 1 above); "cpsnr_chunk" = n > 0 runs
 * pass 1 -> bias -> pass 2 per chunk of n imagesets, -1 sizes the chunk by the L2 (default 0: the whole batch).  All agree.
 * "lanczos_scalar" = 1 makes hrn_lanczos_shift use the register-window kernel also for rows that TMA can address. */
int32_t hrn_scoring_debug_set(const char* knob, int32_t value);

/* Per-launch device timing of HRNet.forward, by kernel class, with CUDA events recorded on the stream the
 * kernels run on.  hrn_profile_begin arms it; every later hrn_forward* call records one event pair per launch;
 * hrn_profile_end synchronises the device and returns, per class, the summed milliseconds, the summed
 * algorithmic FLOPs (MAC = 2) and the launch count (arrays of HRN_PROF_CLASSES entries), then disarms. */
#define HRN_PROF_CONV64 0     /* tcgen05 conv 64 -> 64            */
#define HRN_PROF_CONV128 1    /* tcgen05 conv 128 -> 128 / 64     */
#define HRN_PROF_CONV_INIT 2  /* conv 2 -> 64 (+ PReLU)           */
#define HRN_PROF_DECODER 3    /* deconv + PReLU + 1x1             */
#define HRN_PROF_MEDIAN 4     /* median anchor                    */
#define HRN_PROF_RESBLOCK64 5 /* fused encoder ResidualBlock (two 64 -> 64 convs in one launch) */
#define HRN_PROF_LIVE_LISTS 6 /* live-work lists (first launch of every forward) */
#define HRN_PROF_FORWARD 7    /* one span around the whole forward: FORWARD - sum(classes 0..6) = gaps between launches */
#define HRN_PROF_FUSE_WAVE 8  /* fused fusion level (three convs of one level in one wavefront launch) */
#define HRN_PROF_ENC_WAVE 9   /* encoder ResidualBlocks + final conv (five 64 -> 64 convs in one wavefront launch) */
#define HRN_PROF_CLASSES 10
int32_t hrn_profile_begin(hrn_handle* h);
int32_t hrn_profile_end(hrn_handle* h, double* ms, double* flops, int64_t* launches);

/* ---- The file formats on both ends of the path; host code, no GPU involved, `threads` = 0 means all
 * host cores.  PNG: non-interlaced greyscale, bit depth 1/2/4/8/16 (what the Proba-V files are).
 * hrn_png_info: header of one file.  hrn_png_read_gray_u16: decodes n files of height x width into dst
 * (n, height, width) uint16 -- caller-owned, e.g. pinned memory that hrn_forward_host_u16 / hrn_collate then take -- with
 * the sample values as stored (0..65535 for 16-bit views, 0..255 / 0..1 for 8- / 1-bit status maps); replaces the
 * io.imread loops of DataLoader.py:134-140.  hrn_clearance_scores: sum of every QM status map, replaces
 * save_clearance.py:13-27.  hrn_clearance_order: the view order of DataLoader.py:128-131, np.argsort(clearances)[::-1]
 * (equal scores by descending index).  hrn_png_write_gray_u16: n images (n, height, width) uint16 -> 16-bit PNG files,
 * replaces io.imsave of predict.py:181; hrn_zip_store: the stored archive of predict.py:186-195 (ZipFile(mode='w')). */
int32_t hrn_png_info(const char* path, int32_t* width, int32_t* height, int32_t* bit_depth, int32_t* color_type);
int32_t hrn_png_read_gray_u16(const char* const* paths, int32_t n, int32_t height, int32_t width, uint16_t* dst,
                              int32_t threads);
int32_t hrn_png_write_gray_u16(const char* const* paths, int32_t n, int32_t height, int32_t width, const uint16_t* src,
                               int32_t threads);
int32_t hrn_clearance_scores(const char* const* qm_paths, int32_t n, int32_t height, int32_t width, int32_t threads,
                             double* scores);
int32_t hrn_clearance_order(const double* clearances, int32_t n, int32_t* order);
int32_t hrn_zip_store(const char* zip_path, const char* const* files, const char* const* arcnames, int32_t n);

/* Number of kernels launched by this library (all handles) since load; bench.py reports the delta. */
int64_t hrn_kernel_launch_count(void);

/* ---- ShiftNet, the registration network on the other side of lanczos_shift in the training graph.
 * Replaces DeepNetworks.ShiftNet.ShiftNet.forward (src/DeepNetworks/ShiftNet.py:49-75) in EVAL mode: BatchNorm uses its
 * running statistics (folded into the convs), Dropout is the identity; the caller is train.register_batch
 * (src/train.py:26-44), which feeds pairs cat([reference, view], 1) of 128 x 128 crops.  Autograd and train-mode batch
 * statistics are out of scope.  One handle per device; keys and shapes of hrn_shiftnet_set_weight are those of
 * ShiftNet(in_channel=1).state_dict() (layerK.0.weight/bias, layerK.1.weight/bias/running_mean/running_var for
 * K = 1..8, fc1.weight, fc1.bias, fc2.weight; num_batches_tracked is not needed), HOST fp32 pointers.
 * hrn_shiftnet_forward: x (N, 2, 128, 128) fp32 DEVICE -> theta (N, 2) fp32 DEVICE, enqueued on `stream`. */
typedef struct hrn_shiftnet hrn_shiftnet;
int32_t hrn_shiftnet_create(int32_t device, hrn_shiftnet** out);
void hrn_shiftnet_destroy(hrn_shiftnet* h);
int32_t hrn_shiftnet_set_weight(hrn_shiftnet* h, const char* state_dict_key, const float* host_data, const int64_t* shape,
                                int32_t ndim);
int32_t hrn_shiftnet_missing_weights(hrn_shiftnet* h);
int32_t hrn_shiftnet_forward(hrn_shiftnet* h, const float* x, int32_t N, int32_t H, int32_t W, float* theta, void* stream);
/* Test knob: "img_group" = 0 makes the conv layers on narrow feature maps (32 x 32, 16 x 16) use one image row per MMA
 * tile instead of several images side by side (default 1); "fused_pool" = 0 runs MaxPool2d(2) as its own launch instead of
 * inside the conv epilogue (default 1); thetas must be bit-identical either way. */
int32_t hrn_shiftnet_debug_set(hrn_shiftnet* h, const char* knob, int32_t value);

#ifdef __cplusplus
}
#endif
#endif /* HRN_B200_H */
