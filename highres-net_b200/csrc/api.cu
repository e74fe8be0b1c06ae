// C ABI of libhrn_b200.so (declared in include/hrn_b200.h): handle, weight
// repacking, workspace and the layer schedule of HRNet.forward (HRNet.py:186-211).
#include "../../include/hrn_b200.h"
#include "internal.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <set>
#include <string>
#include <vector>

namespace hrn {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct ConvLayer {
    int cin = 0, cout = 0;
    uint8_t* w_img = nullptr;   // device, pre-swizzled bf16
    float* bias = nullptr;      // device
    float bias_host[128] = {};  // host copy (the fused ResidualBlock kernel takes its biases as kernel parameters)
    float prelu = 0.0f;
    bool has_prelu = false;
};

}  // namespace hrn

struct hrn_handle {
    hrn_config cfg;
    int device = 0, sm_count = 0;
    std::set<std::string> have;
    int expected = 0;
    // first conv (2 -> 64) and decoder parameters (fp32, device)
    uint8_t* w_init_img = nullptr;     // conv_init_pack_weights() image
    float* b_init = nullptr;
    float prelu_init = 0.0f;
    std::vector<hrn::ConvLayer> enc;   // 2 * num_layers residual convs + the final conv
    hrn::ConvLayer fuse[3];
    uint8_t* wd_img = nullptr;         // decoder_pack_weights() image
    float *bd = nullptr, *wf = nullptr;
    float prelu_dec = 0.0f, bf = 0.0f;
    // workspace
    size_t act_cap = 0;                // bytes of each activation buffer
    __nv_bfloat16* act[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t anchor_cap = 0;
    float* anchor = nullptr;
    size_t lists_cap = 0, live_scratch_cap = 0;
    int* lists = nullptr;              // live-work lists of the current forward call (pointwise.cu: live_lists_kernel)
    uint8_t* live_scratch = nullptr;
    int skip_dead = 1;                 // 0: process every view and pair even when it cannot reach the output (test knob)
    size_t io_cap[3] = {0, 0, 0};
    float* io[3] = {nullptr, nullptr, nullptr};   // device staging for hrn_forward_host: lrs, alphas, sr
    uint16_t* io_u16 = nullptr;                   // raw uint16 views of hrn_forward_host_u16
    size_t io_u16_cap = 0;
    int max_ctas = 0;                  // 0 = one CTA per SM (test knob)
    int strip_split = 0;               // ranges of the row space per CTA (0/1 = one contiguous range)
    int mcast = 1;                     // 128 -> 128 convs as cluster pairs with multicast A rows (0: plain launch, test knob)
    int fuse_resblock = 1;             // encoder ResidualBlocks as one launch each (resblock64_umma.cu) when W <= 128
    int fuse_wave = 1;                 // fusion levels as one wavefront launch each (fuse_wave_umma.cu) when W <= 128
    int enc_wave = 0;                  // 1: encoder ResidualBlocks + final conv as one wavefront launch (enc_wave_umma.cu, W <= 128).
                                       // Off by default: bit-identical and 4.5 GB less DRAM traffic, but 3.4 ms against 2.3-2.6 ms
                                       // (the two skip stages are bound by the load/store unit; profiles/r02_enc_wave_triage.log)
    int enc_ring_rows = 24;            // rows per stream ring of the encoder wavefront (x1 has two readers, the second ~10 rows later)
    __nv_bfloat16* enc_ring[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t enc_ring_cap[4] = {0, 0, 0, 0};
    int wave_ring_rows = 16;           // rows per stream ring of the wavefront schedule (test knob, >= 8)
    int wave_streams = 0;              // streams of the wavefront schedule (0 = sm_count / 5; test knob)
    int wave_publish_rows = 1;         // rows per hand-over publication of the wavefront schedule
    int wave_lag_rows = 0;             // triage: consumers stay this many rows behind their producer
    __nv_bfloat16* wave_ring[2] = {nullptr, nullptr};
    size_t wave_ring_cap[2] = {0, 0};
    unsigned long long* wave_stats = nullptr;   // triage counters of the fusion wavefront kernel (knob "wave_stats"), 12 per CTA
    unsigned long long* enc_stats = nullptr;    // the same for the encoder wavefront (knob "enc_stats")
    uint32_t* wave_flags = nullptr;    // 16 levels x fuse_wave_flag_bytes()
    size_t wave_flags_cap = 0;
    int host_chunks = 0;               // hrn_forward_host pipeline depth (0 = automatic)
    long long workspace_mb = 65536;    // cap on the activation workspace; larger batches are run in slices
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // H2D / D2H streams of hrn_forward_host
    cudaEvent_t ev_in[8] = {}, ev_done[8] = {};
    // hrn_forward_host_submit / _wait: two calls in flight, each with its own device staging and events
    struct HostSlot {
        float *lrs = nullptr, *alphas = nullptr, *sr = nullptr;
        size_t cap[3] = {0, 0, 0};
        cudaEvent_t ev_in = nullptr, ev_kernels = nullptr, ev_out = nullptr;   // H2D done, kernels done, D2H done
        int64_t ticket = 0;                                                      // 0 = free
    } slots[2];
    int64_t next_ticket = 1;
    int debug_flags = 0;
    bool work_enqueued = false;        // a forward may still be running: hrn_set_weight drains the device before it overwrites weights
    // optional per-launch timing (hrn_profile_begin / hrn_profile_end)
    bool profiling = false, prof_error = false;
    cudaEvent_t ev_fwd_done = nullptr;  // recorded after the last kernel of every forward (the workspace is shared)
    cudaStream_t last_stream = nullptr;
    bool have_last = false;
    struct Span { cudaEvent_t e0, e1; int cls; double flops; };
    std::vector<Span> spans;
};

namespace {

using hrn::set_error;

int grow(void** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p != nullptr) HRN_CUDA_OK(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    HRN_CUDA_OK(cudaMalloc(p, need));
    *cap = need;
    return 0;
}

int upload(float** dst, const float* src, size_t n) {
    if (*dst == nullptr) HRN_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(dst), n * sizeof(float)));
    HRN_CUDA_OK(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

bool shape_is(const int64_t* shape, int ndim, std::initializer_list<int64_t> want) {
    if (ndim != static_cast<int>(want.size())) return false;
    int i = 0;
    for (int64_t w : want)
        if (shape[i++] != w) return false;
    return true;
}

int set_conv_weight(hrn::ConvLayer& l, const float* data, const int64_t* shape, int ndim) {
    if (!shape_is(shape, ndim, {l.cout, l.cin, 3, 3})) {
        set_error("conv weight: expected shape (%d, %d, 3, 3)", l.cout, l.cin);
        return -1;
    }
    const size_t bytes = hrn::conv3x3_bytes_per_weight_image(l.cin, l.cout);
    std::vector<uint8_t> img(bytes);
    hrn::conv3x3_pack_weights(data, l.cin, l.cout, img.data());
    if (l.w_img == nullptr) HRN_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&l.w_img), bytes));
    HRN_CUDA_OK(cudaMemcpy(l.w_img, img.data(), bytes, cudaMemcpyHostToDevice));
    return 0;
}

int set_conv_bias(hrn::ConvLayer& l, const float* data, const int64_t* shape, int ndim) {
    if (!shape_is(shape, ndim, {l.cout})) {
        set_error("conv bias: expected shape (%d,)", l.cout);
        return -1;
    }
    std::memcpy(l.bias_host, data, static_cast<size_t>(l.cout) * sizeof(float));
    return upload(&l.bias, data, l.cout);
}

int set_prelu(float* slot, bool* flag, const int64_t* shape, int ndim, const float* data) {
    if (!shape_is(shape, ndim, {1})) {
        set_error("PReLU weight: expected shape (1,) (single shared slope, nn.PReLU() default)");
        return -1;
    }
    *slot = data[0];
    if (flag != nullptr) *flag = true;
    return 0;
}

struct Dump {
    int stage;
    float* dst;
    bool hit;
};

int maybe_dump(Dump* d, int stage, const __nv_bfloat16* t, int n, int H, int W, int C, cudaStream_t s, int group = 1,
               int stride = 1) {
    if (d == nullptr || d->stage != stage) return 0;
    d->hit = true;
    return hrn::nhwc_bf16_to_nchw_f32_launch(t, n, H, W, C, group, stride, d->dst, s);
}

// Sizes the handle-owned workspace (five bf16 NHWC activation buffers, the anchor plane, the live-work lists) for a
// (B, L, H, W) forward.  Growing frees and reallocates, which synchronises the device: callers that must never stall
// mid-stream size the workspace up front with hrn_reserve().
int ensure_workspace(hrn_handle* h, int B, int L, int H, int W) {
    const size_t hw = static_cast<size_t>(H) * W;
    const size_t n_img = static_cast<size_t>(B) * L;
    const size_t act_bytes = n_img * hw * 64 * sizeof(__nv_bfloat16);
    if (act_bytes > h->act_cap) {
        for (int i = 0; i < 5; ++i) {
            if (h->act[i] != nullptr) HRN_CUDA_OK(cudaFree(h->act[i]));
            h->act[i] = nullptr;
        }
        h->act_cap = 0;
        for (int i = 0; i < 5; ++i) HRN_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->act[i]), act_bytes));
        h->act_cap = act_bytes;
    }
    if (grow(reinterpret_cast<void**>(&h->anchor), &h->anchor_cap, static_cast<size_t>(B) * hw * sizeof(float))) return -1;
    if (grow(reinterpret_cast<void**>(&h->lists), &h->lists_cap, hrn::live_lists_ints(B, L) * sizeof(int))) return -1;
    if (grow(reinterpret_cast<void**>(&h->live_scratch), &h->live_scratch_cap, hrn::live_scratch_bytes(B, L))) return -1;
    if ((h->fuse_wave || h->enc_wave) && W <= 128 && hrn::fuse_wave_streams(h->sm_count) >= 1 && hrn::fuse_wave_fits(h->sm_count)) {
        const size_t ring = hrn::fuse_wave_ring_bytes(h->sm_count, h->wave_ring_rows, W);
        for (int i = 0; i < 2; ++i)
            if (grow(reinterpret_cast<void**>(&h->wave_ring[i]), &h->wave_ring_cap[i], ring)) return -1;
        if (grow(reinterpret_cast<void**>(&h->wave_flags), &h->wave_flags_cap,
                 16 * hrn::fuse_wave_flag_bytes(h->sm_count) + hrn::enc_wave_flag_bytes(h->sm_count))) return -1;
        if (h->enc_wave && hrn::enc_wave_fits(h->sm_count)) {
            const size_t ering = hrn::enc_wave_ring_bytes(h->sm_count, h->enc_ring_rows, W);
            for (int i = 0; i < 4; ++i)
                if (grow(reinterpret_cast<void**>(&h->enc_ring[i]), &h->enc_ring_cap[i], ering)) return -1;
        }
    }
    return 0;
}

// Wavefront launches spin on flags written by sibling CTAs, so every CTA of such a grid must become resident.  Two of them
// enqueued on different streams could each grab half of the SMs and wait for the other half for ever; they are therefore
// serialised per device with an event (each one fills the GPU anyway).  The mutex covers "wait, launch, record".
std::mutex g_wave_mutex;
cudaEvent_t g_wave_event[64] = {};

// Brackets one kernel launch with CUDA events on its own stream when profiling is on.
struct SpanGuard {
    hrn_handle* h;
    cudaStream_t s;
    hrn_handle::Span sp{};
    bool on;
    SpanGuard(hrn_handle* h_, cudaStream_t s_, int cls, double flops) : h(h_), s(s_), on(h_->profiling) {
        if (!on) return;
        sp.cls = cls;
        sp.flops = flops;
        sp.e0 = sp.e1 = nullptr;
        // a failed event call must not pass silently: hrn_profile_end reports it instead of returning bogus times
        if (cudaEventCreate(&sp.e0) != cudaSuccess || cudaEventCreate(&sp.e1) != cudaSuccess ||
            cudaEventRecord(sp.e0, s) != cudaSuccess) {
            h->prof_error = true;
            if (sp.e0 != nullptr) cudaEventDestroy(sp.e0);
            if (sp.e1 != nullptr) cudaEventDestroy(sp.e1);
            on = false;
        }
    }
    ~SpanGuard() {
        if (!on) return;
        if (cudaEventRecord(sp.e1, s) != cudaSuccess) {
            h->prof_error = true;
            cudaEventDestroy(sp.e0);
            cudaEventDestroy(sp.e1);
            return;
        }
        h->spans.push_back(sp);
    }
};

int run_conv(hrn_handle* h, const hrn::ConvLayer& l, hrn::ConvArgs a, cudaStream_t s) {
    SpanGuard guard(h, s, l.cin == 64 ? HRN_PROF_CONV64 : HRN_PROF_CONV128,
                    2.0 * 9.0 * l.cin * l.cout * static_cast<double>(a.n_img) * a.H * a.W);
    a.cin = l.cin;
    a.cout = l.cout;
    a.w_img = l.w_img;
    a.bias = l.bias;
    a.prelu = l.prelu;
    a.has_prelu = l.has_prelu ? 1 : 0;
    a.max_ctas = h->max_ctas;
    a.strip_split = h->strip_split;
    a.debug_flags = h->debug_flags;
    a.mcast = h->mcast;
    return hrn::conv3x3_launch(a, h->sm_count, s);
}

int forward_impl(hrn_handle* h, const float* lrs, const float* alphas, int B, int L, int H, int W, float* sr,
                 cudaStream_t s, Dump* dump) {
    if (h == nullptr) {
        set_error("null handle");
        return -1;
    }
    if (hrn_missing_weights(h) != 0) {
        set_error("hrn_forward: %d of %d state_dict tensors have not been set", hrn_missing_weights(h), h->expected);
        return -1;
    }
    if (B <= 0 || L <= 0 || H <= 0 || W <= 0) {
        set_error("hrn_forward: empty input (B=%d L=%d H=%d W=%d)", B, L, H, W);
        return -1;
    }
    if (H != W) {
        set_error("hrn_forward: square inputs only (got %d x %d); the reference view() at HRNet.py:204 swaps H and W",
                  H, W);
        return -1;
    }
    hrn::DeviceGuard on_device(h->device);
    if (!on_device.ok) return -1;
    h->work_enqueued = true;
    const size_t hw = static_cast<size_t>(H) * W;
    const size_t n_img = static_cast<size_t>(B) * L;
    if (ensure_workspace(h, B, L, H, W)) return -1;
    // The activation workspace and the live-work lists belong to the handle: a forward on another stream than the
    // previous one must not start before that one has finished (same-stream calls are ordered anyway).
    if (h->ev_fwd_done == nullptr) HRN_CUDA_OK(cudaEventCreateWithFlags(&h->ev_fwd_done, cudaEventDisableTiming));
    if (h->have_last && h->last_stream != s) HRN_CUDA_OK(cudaStreamWaitEvent(s, h->ev_fwd_done, 0));
    SpanGuard whole(h, s, HRN_PROF_FORWARD, 0.0);
    if (h->wave_flags != nullptr && dump == nullptr && (h->fuse_wave || h->enc_wave) && W <= 128)     // hand-over counters of the wavefront launches
        HRN_CUDA_OK(cudaMemsetAsync(h->wave_flags, 0, h->wave_flags_cap, s));

    // ---- live-work lists: views / pairs that cannot reach the output (alpha = 0 padding) are not computed at all.
    // The stage-dump hook asks for dense lists so that every intermediate tensor is defined.  Must stay the first
    // launch of the pass (see live_lists_launch).
    {
        SpanGuard guard(h, s, HRN_PROF_LIVE_LISTS, 0.0);
        if (hrn::live_lists_launch(alphas, B, L, (dump == nullptr && h->skip_dead) ? 1 : 0, h->cfg.rec_alpha_residual ? 1 : 0,
                                   h->live_scratch, h->lists, s))
            return -1;
    }
    const int* enc_list = h->lists + hrn::LIVE_HDR;
    const int* enc_count = h->lists;

    // ---- anchor + first conv (HRNet.py:200-204, 51-53)
    {
        SpanGuard guard(h, s, HRN_PROF_MEDIAN, 0.0);
        if (hrn::median_anchor_launch(lrs, B, L, H, W, h->anchor, s)) return -1;
    }
    if (dump != nullptr && dump->stage == HRN_STAGE_ANCHOR) {
        HRN_CUDA_OK(cudaMemcpyAsync(dump->dst, h->anchor, static_cast<size_t>(B) * hw * sizeof(float),
                                    cudaMemcpyDeviceToDevice, s));
        dump->hit = true;
    }
    {
        SpanGuard guard(h, s, HRN_PROF_CONV_INIT, 2.0 * 18.0 * 64.0 * static_cast<double>(n_img) * hw);
        if (hrn::conv_init_umma_launch(lrs, h->anchor, B, L, H, W, h->w_init_img, h->b_init, h->prelu_init, h->act[0], enc_list, enc_count, h->sm_count, s)) return -1;
    }
    int stage = 0;
    if (maybe_dump(dump, HRN_STAGE_ENC(stage), h->act[0], static_cast<int>(n_img), H, W, 64, s)) return -1;

    // ---- encoder residual blocks + final conv (HRNet.py:55-60, 17-33)
    hrn::ConvArgs base{};
    base.n_img = static_cast<int>(n_img);
    base.H = H;
    base.W = W;
    base.in_images = static_cast<int>(n_img);
    base.in_c = 64;
    base.live_list = enc_list;
    base.live_count = enc_count;
    int cur = 0;
    // Wavefront schedules spin on flags written by sibling CTAs: they are serialised per device (see g_wave_mutex); the lock
    // covers the enqueue of every wavefront launch of this forward.
    const bool wave_ok = dump == nullptr && W <= 128 && h->wave_flags != nullptr;
    bool enc_as_wave = wave_ok && h->enc_wave && h->cfg.enc_num_layers == 2 && h->enc_ring[0] != nullptr;
    for (int i = 0; i < 4 && enc_as_wave; ++i) enc_as_wave = h->enc[i].has_prelu;
    const bool fuse_as_wave = wave_ok && h->fuse_wave && L >= 2 && h->fuse[0].has_prelu && h->fuse[1].has_prelu && h->fuse[2].has_prelu;
    std::unique_lock<std::mutex> wave_lock(g_wave_mutex, std::defer_lock);
    if (enc_as_wave || fuse_as_wave) {
        wave_lock.lock();
        cudaEvent_t& ev = g_wave_event[h->device];
        if (ev == nullptr) HRN_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        else HRN_CUDA_OK(cudaStreamWaitEvent(s, ev, 0));
    }
    if (enc_as_wave) {
        // the two ResidualBlocks and the final conv in one launch (enc_wave_umma.cu): x0 = act[0] -> view stack = act[1]
        hrn::EncWaveLaunch ew{};
        ew.H = H;
        ew.W = W;
        ew.n_img = static_cast<int>(n_img);
        ew.live_list = enc_list;
        ew.live_count = enc_count;
        ew.x0 = h->act[0];
        ew.out = h->act[1];
        for (int i = 0; i < 4; ++i) ew.ring[i] = h->enc_ring[i];
        ew.ring_rows = h->enc_ring_rows;
        ew.flags = h->wave_flags + 16 * (hrn::fuse_wave_flag_bytes(h->sm_count) / sizeof(uint32_t));
        ew.debug_flags = h->debug_flags;
        ew.streams = h->wave_streams;
        ew.stats = h->enc_stats;
        for (int i = 0; i < 5; ++i) {
            ew.w_img[i] = h->enc[i].w_img;
            ew.bias[i] = h->enc[i].bias;
            ew.prelu[i] = h->enc[i].prelu;
            ew.has_prelu[i] = h->enc[i].has_prelu ? 1 : 0;
        }
        {
            SpanGuard guard(h, s, HRN_PROF_ENC_WAVE, 5.0 * 2.0 * 9.0 * 64 * 64 * static_cast<double>(n_img) * hw);
            if (hrn::enc_wave_launch(ew, h->sm_count, s)) return -1;
        }
        cur = 1;
        stage += 5;
    } else {
    for (int r = 0; r < h->cfg.enc_num_layers; ++r) {
        const int t1 = (cur + 1) % 3, t2 = (cur + 2) % 3;
        hrn::ConvArgs a = base;
        a.in = h->act[cur];
        // One launch for the whole ResidualBlock when the image is a single column tile wide (the intermediate never
        // leaves the SM); the stage-dump hook and wider images take the two-launch path below.
        if (h->fuse_resblock && dump == nullptr && h->enc[2 * r].has_prelu && h->enc[2 * r + 1].has_prelu) {
            const hrn::ConvLayer &l1 = h->enc[2 * r], &l2 = h->enc[2 * r + 1];
            a.out = h->act[t2];
            a.cin = a.cout = 64;
            a.w_img = l1.w_img;
            a.bias = l1.bias;
            a.prelu = l1.prelu;
            a.has_prelu = 1;
            a.max_ctas = h->max_ctas;
            a.strip_split = h->strip_split;
            int rc;
            {
                SpanGuard guard(h, s, HRN_PROF_RESBLOCK64, 2.0 * 2.0 * 9.0 * 64 * 64 * static_cast<double>(a.n_img) * H * W);
                rc = hrn::resblock64_launch(a, l1.bias_host, l2.w_img, l2.bias_host, l2.prelu, h->sm_count, s);
            }
            if (rc < 0) return -1;
            if (rc == 0) {
                stage += 2;
                cur = t2;
                continue;
            }
            a = base;
            a.in = h->act[cur];
        }
        a.out = h->act[t1];
        a.res_mode = hrn::RES_NONE;
        if (run_conv(h, h->enc[2 * r], a, s)) return -1;
        if (maybe_dump(dump, HRN_STAGE_ENC(++stage), h->act[t1], base.n_img, H, W, 64, s)) return -1;
        a.in = h->act[t1];
        a.out = h->act[t2];
        a.res_mode = hrn::RES_SAME;
        a.res = h->act[cur];
        if (run_conv(h, h->enc[2 * r + 1], a, s)) return -1;
        if (maybe_dump(dump, HRN_STAGE_ENC(++stage), h->act[t2], base.n_img, H, W, 64, s)) return -1;
        cur = t2;
    }
    {
        hrn::ConvArgs a = base;
        a.in = h->act[cur];
        a.out = h->act[(cur + 1) % 3];
        a.res_mode = hrn::RES_NONE;
        if (run_conv(h, h->enc.back(), a, s)) return -1;
        cur = (cur + 1) % 3;
        if (maybe_dump(dump, HRN_STAGE_ENC(++stage), h->act[cur], base.n_img, H, W, 64, s)) return -1;
    }

    }

    // ---- recursive fusion (HRNet.py:99-134), in place: the 64-channel view stack keeps its stride of L images per
    // imageset at every level and the merged pair (b, i) overwrites alice's slot b * L + i.  A pair that is not live
    // (alpha_bob = 0) therefore needs no work at all: alice is already where the next level expects it.
    int n = L, level = 0;
    __nv_bfloat16 *stack = h->act[cur], *t1 = h->act[3], *t2 = h->act[4];
    const int* pair_list = h->lists + hrn::LIVE_HDR + n_img;
    // Wavefront schedule (fuse_wave_umma.cu): one launch per level, the 128-channel intermediates stay in L2-resident rings,
    // the level writes into a second view stack (act[3]); images wider than one column tile and the stage-dump hook take
    // the three-launch schedule below.
    if (fuse_as_wave) {
        const size_t flag_ints = hrn::fuse_wave_flag_bytes(h->sm_count) / sizeof(uint32_t);
        __nv_bfloat16* other = h->act[3];
        while (n / 2 > 0) {
            const int half = n / 2, top = n - (n % 2);
            hrn::FuseWaveArgs fw{};
            fw.H = H;
            fw.W = W;
            fw.half = half;
            fw.src_views = L;
            fw.top = top;
            fw.live_list = pair_list;
            fw.live_count = h->lists + 1 + level;
            fw.carry_list = pair_list + n_img;               // the carry lists mirror the pair lists, B*L ints further on
            fw.carry_count = h->lists + 16 + level;
            fw.stack_in = stack;
            fw.stack_images = B * L;
            fw.stack_out = other;
            fw.ring1 = h->wave_ring[0];
            fw.ring2 = h->wave_ring[1];
            fw.ring_rows = h->wave_ring_rows;
            fw.flags = h->wave_flags + static_cast<size_t>(level) * flag_ints;
            fw.alphas = alphas;
            fw.alpha_stride = L;
            fw.alpha_residual = h->cfg.rec_alpha_residual ? 1 : 0;
            fw.debug_flags = h->debug_flags;
            fw.streams = h->wave_streams;
            fw.stats = h->wave_stats;
            fw.publish_rows = h->wave_publish_rows;
            fw.lag_rows = h->wave_lag_rows;
            for (int i = 0; i < 3; ++i) {
                fw.w_img[i] = h->fuse[i].w_img;
                fw.bias[i] = h->fuse[i].bias;
                fw.prelu[i] = h->fuse[i].prelu;
                fw.has_prelu[i] = 1;
            }
            {
                SpanGuard guard(h, s, HRN_PROF_FUSE_WAVE, 2.0 * 9.0 * (128.0 * 128 * 2 + 128.0 * 64) * B * half * static_cast<double>(hw));
                if (hrn::fuse_wave_launch(fw, h->sm_count, s)) return -1;
            }
            std::swap(stack, other);
            pair_list += B * half;
            n = half;
            ++level;
        }
    }
    if (wave_lock.owns_lock()) {
        HRN_CUDA_OK(cudaEventRecord(g_wave_event[h->device], s));
        wave_lock.unlock();
    }
    while (n / 2 > 0) {
        const int half = n / 2, top = n - (n % 2);
        hrn::ConvArgs a{};
        a.n_img = B * half;
        a.H = H;
        a.W = W;
        a.half = half;
        a.src_views = L;
        a.top = top;
        a.live_list = pair_list;
        a.live_count = h->lists + 1 + level;
        // conv 1 of the residual block on cat(alice, bob): the concat is two K chunks from two views
        a.pair_mode = 1;
        a.in = stack;
        a.in_images = B * L;
        a.in_c = 64;
        a.out = t1;
        a.res_mode = hrn::RES_NONE;
        if (run_conv(h, h->fuse[0], a, s)) return -1;
        if (maybe_dump(dump, HRN_STAGE_FUSE(level, 0), t1, B * half, H, W, 128, s)) return -1;
        // conv 2 + skip connection onto cat(alice, bob)
        a.pair_mode = 0;
        a.in = t1;
        a.in_images = B * half;
        a.in_c = 128;
        a.out = t2;
        a.res_mode = hrn::RES_PAIR;
        a.res = stack;
        if (run_conv(h, h->fuse[1], a, s)) return -1;
        if (maybe_dump(dump, HRN_STAGE_FUSE(level, 1), t2, B * half, H, W, 128, s)) return -1;
        // conv 128 -> 64 + PReLU, then alice + alpha_bob * x, written over alice
        a.in = t2;
        a.out = stack;
        a.out_in_stack = 1;
        a.res_mode = h->cfg.rec_alpha_residual ? hrn::RES_ALPHA : hrn::RES_NONE;
        a.res = stack;
        a.alphas = alphas;
        a.alpha_stride = L;
        if (run_conv(h, h->fuse[2], a, s)) return -1;
        if (maybe_dump(dump, HRN_STAGE_FUSE(level, 2), stack, B * half, H, W, 64, s, half, L)) return -1;
        pair_list += B * half;
        n = half;
        ++level;
    }
    // torch.mean over the single remaining view (HRNet.py:134) is the identity: the loop always ends at n == 1.

    // ---- decoder (HRNet.py:147-156) on view 0 of every imageset
    {
        SpanGuard guard(h, s, HRN_PROF_DECODER, 74880.0 * static_cast<double>(B) * hw);
        if (hrn::decoder_umma_launch(stack, B, L, H, W, h->wd_img, h->bd, h->prelu_dec, h->wf, h->bf, sr, h->sm_count, s)) return -1;
    }
    if (dump != nullptr && !dump->hit) {
        set_error("hrn_forward_dump: stage 0x%x does not exist for L=%d", dump->stage, L);
        return -1;
    }
    HRN_CUDA_OK(cudaEventRecord(h->ev_fwd_done, s));
    h->last_stream = s;
    h->have_last = true;
    return 0;
}

}  // namespace

// ============================================================================ C ABI
extern "C" {

int32_t hrn_abi_version(void) { return HRN_ABI_VERSION; }
const char* hrn_last_error(void) { return hrn::g_error; }
int64_t hrn_kernel_launch_count(void) { return hrn::g_launches.load(); }

int32_t hrn_create(const hrn_config* cfg, int32_t device, hrn_handle** out) {
    if (cfg == nullptr || out == nullptr) {
        set_error("hrn_create: null argument");
        return -1;
    }
    *out = nullptr;
    const hrn_config& c = *cfg;
    if (c.enc_in_channels != 2 || c.enc_kernel_size != 3 || c.enc_channels != 64 || c.rec_in_channels != 64 ||
        c.rec_kernel_size != 3 || c.dec_in_channels != 64 || c.dec_out_channels != 64 || c.dec_kernel_size != 3 ||
        c.dec_stride != 3 || c.fin_in_channels != 64 || c.fin_kernel_size != 1 || c.fin_out_channels != 1 ||
        c.enc_num_layers < 0 || c.enc_num_layers > 16) {
        set_error("hrn_create: unsupported network config; kernels are specialised for encoder 2->64 k3, "
                  "recursive 64 k3, deconv 64->64 k3 s3, final 64->1 k1 (config/config.json), num_layers 0..16");
        return -1;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        set_error("hrn_create: no CUDA device visible; this library has no CPU fallback");
        return -1;
    }
    if (device < 0 || device >= ndev) {
        set_error("hrn_create: device %d out of range (0..%d)", device, ndev - 1);
        return -1;
    }
    cudaDeviceProp prop;
    HRN_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("hrn_create: device %d is sm_%d%d; kernels are built for sm_100a (B200) only", device, prop.major,
                  prop.minor);
        return -1;
    }
    hrn::DeviceGuard on_device(device);
    if (!on_device.ok) return -1;
    hrn_handle* h = new hrn_handle();
    h->cfg = c;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->enc.resize(2 * c.enc_num_layers + 1);
    for (size_t i = 0; i < h->enc.size(); ++i) {
        h->enc[i].cin = h->enc[i].cout = 64;
        h->enc[i].has_prelu = false;   // set when its PReLU slope arrives; the final conv has none
    }
    h->fuse[0].cin = h->fuse[0].cout = 128;
    h->fuse[1].cin = h->fuse[1].cout = 128;
    h->fuse[2].cin = 128;
    h->fuse[2].cout = 64;
    h->expected = 3 + 6 * c.enc_num_layers + 2 + 9 + 5;
    *out = h;
    return 0;
}

void hrn_destroy(hrn_handle* h) {
    if (h == nullptr) return;
    hrn::DeviceGuard on_device(h->device);
    auto rel = [](void* p) {
        if (p != nullptr) cudaFree(p);
    };
    rel(h->w_init_img);
    rel(h->b_init);
    for (auto& l : h->enc) {
        rel(l.w_img);
        rel(l.bias);
    }
    for (auto& l : h->fuse) {
        rel(l.w_img);
        rel(l.bias);
    }
    rel(h->wd_img);
    rel(h->bd);
    rel(h->wf);
    for (auto* p : h->act) rel(p);
    rel(h->anchor);
    rel(h->lists);
    rel(h->live_scratch);
    rel(h->wave_ring[0]);
    rel(h->wave_ring[1]);
    rel(h->wave_flags);
    rel(h->wave_stats);
    rel(h->enc_stats);
    for (auto* p : h->enc_ring) rel(p);
    for (auto* p : h->io) rel(p);
    rel(h->io_u16);
    if (h->ev_fwd_done != nullptr) cudaEventDestroy(h->ev_fwd_done);
    for (auto& sp : h->spans) {
        cudaEventDestroy(sp.e0);
        cudaEventDestroy(sp.e1);
    }
    for (auto& sl : h->slots) {
        rel(sl.lrs);
        rel(sl.alphas);
        rel(sl.sr);
        if (sl.ev_in != nullptr) {
            cudaEventDestroy(sl.ev_in);
            cudaEventDestroy(sl.ev_kernels);
            cudaEventDestroy(sl.ev_out);
        }
    }
    if (h->copy_in != nullptr) {
        cudaStreamDestroy(h->copy_in);
        cudaStreamDestroy(h->copy_out);
        for (int i = 0; i < 8; ++i) {
            cudaEventDestroy(h->ev_in[i]);
            cudaEventDestroy(h->ev_done[i]);
        }
    }
    delete h;
}

int32_t hrn_missing_weights(const hrn_handle* h) {
    return h == nullptr ? -1 : h->expected - static_cast<int>(h->have.size());
}

int32_t hrn_set_weight(hrn_handle* h, const char* key, const float* data, const int64_t* shape, int32_t ndim) {
    if (h == nullptr || key == nullptr || data == nullptr || shape == nullptr) {
        set_error("hrn_set_weight: null argument");
        return -1;
    }
    hrn::DeviceGuard on_device(h->device);
    if (!on_device.ok) return -1;
    if (h->work_enqueued) {            // forwards are asynchronous and may run on non-blocking streams
        HRN_CUDA_OK(cudaDeviceSynchronize());
        h->work_enqueued = false;
    }
    const std::string k(key);
    int rc = -2;
    int r = 0, j = 0;
    char tail[16] = "";
    if (k == "encode.init_layer.0.weight") {
        if (shape_is(shape, ndim, {64, 2, 3, 3})) {
            std::vector<uint8_t> img(hrn::conv_init_weight_image_bytes());
            hrn::conv_init_pack_weights(data, img.data());
            rc = 0;
            if (h->w_init_img == nullptr && cudaMalloc(reinterpret_cast<void**>(&h->w_init_img), img.size()) != cudaSuccess) rc = -4;
            if (rc == 0 && cudaMemcpy(h->w_init_img, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) rc = -4;
        } else rc = -3;
    } else if (k == "encode.init_layer.0.bias") {
        rc = shape_is(shape, ndim, {64}) ? upload(&h->b_init, data, 64) : -3;
    } else if (k == "encode.init_layer.1.weight") {
        rc = set_prelu(&h->prelu_init, nullptr, shape, ndim, data);
    } else if (sscanf(key, "encode.res_layers.%d.block.%d.%15s", &r, &j, tail) == 3 && r >= 0 &&
               r < h->cfg.enc_num_layers && j >= 0 && j <= 3) {
        hrn::ConvLayer& l = h->enc[2 * r + j / 2];
        if ((j & 1) == 0 && strcmp(tail, "weight") == 0) rc = set_conv_weight(l, data, shape, ndim);
        else if ((j & 1) == 0 && strcmp(tail, "bias") == 0) rc = set_conv_bias(l, data, shape, ndim);
        else if ((j & 1) == 1 && strcmp(tail, "weight") == 0) rc = set_prelu(&l.prelu, &l.has_prelu, shape, ndim, data);
    } else if (k == "encode.final.0.weight") {
        rc = set_conv_weight(h->enc.back(), data, shape, ndim);
    } else if (k == "encode.final.0.bias") {
        rc = set_conv_bias(h->enc.back(), data, shape, ndim);
    } else if (sscanf(key, "fuse.fuse.0.block.%d.%15s", &j, tail) == 2 && j >= 0 && j <= 3) {
        hrn::ConvLayer& l = h->fuse[j / 2];
        if ((j & 1) == 0 && strcmp(tail, "weight") == 0) rc = set_conv_weight(l, data, shape, ndim);
        else if ((j & 1) == 0 && strcmp(tail, "bias") == 0) rc = set_conv_bias(l, data, shape, ndim);
        else if ((j & 1) == 1 && strcmp(tail, "weight") == 0) rc = set_prelu(&l.prelu, &l.has_prelu, shape, ndim, data);
    } else if (k == "fuse.fuse.1.weight") {
        rc = set_conv_weight(h->fuse[2], data, shape, ndim);
    } else if (k == "fuse.fuse.1.bias") {
        rc = set_conv_bias(h->fuse[2], data, shape, ndim);
    } else if (k == "fuse.fuse.2.weight") {
        rc = set_prelu(&h->fuse[2].prelu, &h->fuse[2].has_prelu, shape, ndim, data);
    } else if (k == "decode.deconv.0.weight") {
        if (shape_is(shape, ndim, {64, 64, 3, 3})) {
            std::vector<uint8_t> img(hrn::decoder_weight_image_bytes());
            hrn::decoder_pack_weights(data, img.data());
            rc = 0;
            if (h->wd_img == nullptr && cudaMalloc(reinterpret_cast<void**>(&h->wd_img), img.size()) != cudaSuccess) rc = -4;
            if (rc == 0 && cudaMemcpy(h->wd_img, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) rc = -4;
        } else rc = -3;
    } else if (k == "decode.deconv.0.bias") {
        rc = shape_is(shape, ndim, {64}) ? upload(&h->bd, data, 64) : -3;
    } else if (k == "decode.deconv.1.weight") {
        rc = set_prelu(&h->prelu_dec, nullptr, shape, ndim, data);
    } else if (k == "decode.final.weight") {
        rc = shape_is(shape, ndim, {1, 64, 1, 1}) ? upload(&h->wf, data, 64) : -3;
    } else if (k == "decode.final.bias") {
        if (shape_is(shape, ndim, {1})) {
            h->bf = data[0];
            rc = 0;
        } else rc = -3;
    }
    if (rc == -2) {
        set_error("hrn_set_weight: unknown state_dict key '%s'", key);
        return -1;
    }
    if (rc == -3) {
        set_error("hrn_set_weight: wrong shape for '%s'", key);
        return -1;
    }
    if (rc == -4) {
        set_error("hrn_set_weight: CUDA allocation/copy failed for '%s'", key);
        return -1;
    }
    if (rc != 0) return -1;
    h->have.insert(k);
    return 0;
}

// Imagesets are independent, so a batch whose activation workspace (5 buffers of B*L*H*W*128 bytes) would exceed the
// cap is run as consecutive slices on the same stream; results are identical to one big call.
static int forward_sliced(hrn_handle* h, const float* lrs, const float* alphas, int B, int L, int H, int W, float* sr,
                          cudaStream_t s) {
    if (h == nullptr) {
        set_error("null handle");
        return -1;
    }
    if (B <= 0 || L <= 0 || H <= 0 || W <= 0) return forward_impl(h, lrs, alphas, B, L, H, W, sr, s, nullptr);
    const double per_set = 5.0 * L * static_cast<double>(H) * W * 128.0;
    const double budget = static_cast<double>(h->workspace_mb) * 1048576.0;
    long long slice = static_cast<long long>(budget / per_set);
    slice = slice < 1 ? 1 : (slice > B ? B : slice);
    const size_t set_in = static_cast<size_t>(L) * H * W, set_out = static_cast<size_t>(9) * H * W;
    for (long long b0 = 0; b0 < B; b0 += slice) {
        const int nb = static_cast<int>(b0 + slice <= B ? slice : B - b0);
        if (forward_impl(h, lrs + b0 * set_in, alphas + b0 * L, nb, L, H, W, sr + b0 * set_out, s, nullptr)) return -1;
    }
    return 0;
}

int32_t hrn_reserve(hrn_handle* h, int32_t B, int32_t L, int32_t H, int32_t W) {
    if (h == nullptr) {
        set_error("null handle");
        return -1;
    }
    if (B <= 0 || L <= 0 || H <= 0 || W <= 0) {
        set_error("hrn_reserve: empty shape (B=%d L=%d H=%d W=%d)", B, L, H, W);
        return -1;
    }
    hrn::DeviceGuard on_device(h->device);
    if (!on_device.ok) return -1;
    // same slicing rule as hrn_forward: a batch above the workspace cap runs in slices of this many imagesets
    const double per_set = 5.0 * L * static_cast<double>(H) * W * 128.0;
    long long slice = static_cast<long long>(static_cast<double>(h->workspace_mb) * 1048576.0 / per_set);
    slice = slice < 1 ? 1 : (slice > B ? B : slice);
    return ensure_workspace(h, static_cast<int>(slice), L, H, W);
}

int32_t hrn_forward(hrn_handle* h, const float* lrs, const float* alphas, int32_t B, int32_t L, int32_t H, int32_t W,
                    float* sr, void* stream) {
    return forward_sliced(h, lrs, alphas, B, L, H, W, sr, static_cast<cudaStream_t>(stream));
}

int32_t hrn_forward_dump(hrn_handle* h, const float* lrs, const float* alphas, int32_t B, int32_t L, int32_t H,
                         int32_t W, float* sr, int32_t stage, float* dump, void* stream) {
    Dump d{stage, dump, false};
    return forward_impl(h, lrs, alphas, B, L, H, W, sr, static_cast<cudaStream_t>(stream), &d);
}

static int ensure_copy_streams(hrn_handle* h) {
    if (h->copy_in != nullptr) return 0;
    HRN_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking));
    HRN_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
    for (int i = 0; i < 8; ++i) {
        HRN_CUDA_OK(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
        HRN_CUDA_OK(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
    }
    return 0;
}

// Shared body of hrn_forward_host / hrn_forward_host_u16.  lrs_u16 != nullptr: the views arrive as raw uint16 (half the
// H2D bytes) and become float32 on the device.
static int forward_host_impl(hrn_handle* h, const float* lrs_host, const uint16_t* lrs_u16, const float* alphas_host,
                             int32_t B, int32_t L, int32_t H, int32_t W, float* sr_host, void* stream) {
    if (h == nullptr) {
        set_error("null handle");
        return -1;
    }
    if (B <= 0 || L <= 0 || H <= 0 || W <= 0) {
        set_error("hrn_forward_host: empty input");
        return -1;
    }
    hrn::DeviceGuard on_device(h->device);
    if (!on_device.ok) return -1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t set_in = static_cast<size_t>(L) * H * W, set_out = static_cast<size_t>(9) * H * W;
    if (grow(reinterpret_cast<void**>(&h->io[0]), &h->io_cap[0], B * set_in * 4)) return -1;
    if (grow(reinterpret_cast<void**>(&h->io[1]), &h->io_cap[1], static_cast<size_t>(B) * L * 4)) return -1;
    if (grow(reinterpret_cast<void**>(&h->io[2]), &h->io_cap[2], B * set_out * 4)) return -1;
    if (lrs_u16 != nullptr && grow(reinterpret_cast<void**>(&h->io_u16), &h->io_u16_cap, B * set_in * 2)) return -1;
    if (ensure_copy_streams(h)) return -1;
    // Imagesets are independent, so the batch is cut into chunks and pipelined: the H2D copy of chunk k+1 and the
    // D2H copy of chunk k-1 overlap the kernels of chunk k (three streams, events in between).
    // Measured on B200 at B = 32, L = 16, 128x128 (tools/host_chunks.py): 1 / 2 / 4 chunks = 8.99 / 9.00 / 9.14 ms, i.e. the
    // smaller per-chunk kernels cost what the overlap saves, so small batches are not split by default.
    int chunks = h->host_chunks > 0 ? h->host_chunks : (B >= 64 ? 2 : 1);
    chunks = chunks > 8 ? 8 : (chunks > B ? B : chunks);
    const int per = (B + chunks - 1) / chunks;
    // the staging buffers may still be read by work the caller queued on `s` before this call
    HRN_CUDA_OK(cudaEventRecord(h->ev_done[0], s));
    HRN_CUDA_OK(cudaStreamWaitEvent(h->copy_in, h->ev_done[0], 0));
    for (int k = 0; k < chunks; ++k) {
        const int b0 = k * per, nb = (b0 + per <= B ? per : B - b0);
        if (nb <= 0) break;
        if (lrs_u16 != nullptr)
            HRN_CUDA_OK(cudaMemcpyAsync(h->io_u16 + b0 * set_in, lrs_u16 + b0 * set_in, nb * set_in * 2,
                                        cudaMemcpyHostToDevice, h->copy_in));
        else
            HRN_CUDA_OK(cudaMemcpyAsync(h->io[0] + b0 * set_in, lrs_host + b0 * set_in, nb * set_in * 4,
                                        cudaMemcpyHostToDevice, h->copy_in));
        HRN_CUDA_OK(cudaMemcpyAsync(h->io[1] + static_cast<size_t>(b0) * L, alphas_host + static_cast<size_t>(b0) * L,
                                    static_cast<size_t>(nb) * L * 4, cudaMemcpyHostToDevice, h->copy_in));
        HRN_CUDA_OK(cudaEventRecord(h->ev_in[k], h->copy_in));
    }
    for (int k = 0; k < chunks; ++k) {
        const int b0 = k * per, nb = (b0 + per <= B ? per : B - b0);
        if (nb <= 0) break;
        HRN_CUDA_OK(cudaStreamWaitEvent(s, h->ev_in[k], 0));
        if (lrs_u16 != nullptr && hrn::u16_to_unit_float_launch(h->io_u16 + b0 * set_in, nb * set_in, h->io[0] + b0 * set_in, s))
            return -1;
        if (forward_sliced(h, h->io[0] + b0 * set_in, h->io[1] + static_cast<size_t>(b0) * L, nb, L, H, W,
                           h->io[2] + b0 * set_out, s))
            return -1;
        HRN_CUDA_OK(cudaEventRecord(h->ev_done[k], s));
        HRN_CUDA_OK(cudaStreamWaitEvent(h->copy_out, h->ev_done[k], 0));
        HRN_CUDA_OK(cudaMemcpyAsync(sr_host + b0 * set_out, h->io[2] + b0 * set_out, nb * set_out * 4,
                                    cudaMemcpyDeviceToHost, h->copy_out));
    }
    HRN_CUDA_OK(cudaStreamSynchronize(h->copy_out));
    HRN_CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

int32_t hrn_forward_host(hrn_handle* h, const float* lrs_host, const float* alphas_host, int32_t B, int32_t L,
                         int32_t H, int32_t W, float* sr_host, void* stream) {
    if (lrs_host == nullptr || alphas_host == nullptr || sr_host == nullptr) {
        set_error("hrn_forward_host: null pointer");
        return -1;
    }
    return forward_host_impl(h, lrs_host, nullptr, alphas_host, B, L, H, W, sr_host, stream);
}

int32_t hrn_forward_host_u16(hrn_handle* h, const uint16_t* lrs_host, const float* alphas_host, int32_t B, int32_t L,
                             int32_t H, int32_t W, float* sr_host, void* stream) {
    if (lrs_host == nullptr || alphas_host == nullptr || sr_host == nullptr) {
        set_error("hrn_forward_host_u16: null pointer");
        return -1;
    }
    return forward_host_impl(h, nullptr, lrs_host, alphas_host, B, L, H, W, sr_host, stream);
}

// Pipelined host-buffer forward: submit returns as soon as the copies and kernels are enqueued, so the H2D copy of call
// n + 1 and the D2H copy of call n - 1 overlap the kernels of call n.  Two calls may be in flight (two staging slots);
// a third submit first retires the oldest one.
int32_t hrn_forward_host_submit(hrn_handle* h, const float* lrs_host, const float* alphas_host, int32_t B, int32_t L,
                                int32_t H, int32_t W, float* sr_host, void* stream, int64_t* ticket) {
    if (h == nullptr || lrs_host == nullptr || alphas_host == nullptr || sr_host == nullptr || ticket == nullptr) {
        set_error("hrn_forward_host_submit: null argument");
        return -1;
    }
    *ticket = 0;
    if (B <= 0 || L <= 0 || H <= 0 || W <= 0) {
        set_error("hrn_forward_host_submit: empty input");
        return -1;
    }
    hrn::DeviceGuard on_device(h->device);
    if (!on_device.ok) return -1;
    if (ensure_copy_streams(h)) return -1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    hrn_handle::HostSlot& sl = h->slots[h->next_ticket & 1];
    hrn_handle::HostSlot& other = h->slots[(h->next_ticket & 1) ^ 1];
    if (sl.ev_in == nullptr) {
        HRN_CUDA_OK(cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
        HRN_CUDA_OK(cudaEventCreateWithFlags(&sl.ev_kernels, cudaEventDisableTiming));
        HRN_CUDA_OK(cudaEventCreateWithFlags(&sl.ev_out, cudaEventDisableTiming));
    }
    if (sl.ticket != 0) {      // the call that used this slot two submits ago: its results are complete after this
        HRN_CUDA_OK(cudaEventSynchronize(sl.ev_out));
        sl.ticket = 0;
    }
    const size_t n_in = static_cast<size_t>(B) * L * H * W, n_out = static_cast<size_t>(B) * 9 * H * W;
    if (grow(reinterpret_cast<void**>(&sl.lrs), &sl.cap[0], n_in * 4)) return -1;
    if (grow(reinterpret_cast<void**>(&sl.alphas), &sl.cap[1], static_cast<size_t>(B) * L * 4)) return -1;
    if (grow(reinterpret_cast<void**>(&sl.sr), &sl.cap[2], n_out * 4)) return -1;
    HRN_CUDA_OK(cudaMemcpyAsync(sl.lrs, lrs_host, n_in * 4, cudaMemcpyHostToDevice, h->copy_in));
    HRN_CUDA_OK(cudaMemcpyAsync(sl.alphas, alphas_host, static_cast<size_t>(B) * L * 4, cudaMemcpyHostToDevice, h->copy_in));
    HRN_CUDA_OK(cudaEventRecord(sl.ev_in, h->copy_in));
    // the activation workspace is shared: kernels of this call run after the kernels of the previous one even if the
    // caller switched streams in between
    if (other.ticket != 0) HRN_CUDA_OK(cudaStreamWaitEvent(s, other.ev_kernels, 0));
    HRN_CUDA_OK(cudaStreamWaitEvent(s, sl.ev_in, 0));
    if (forward_sliced(h, sl.lrs, sl.alphas, B, L, H, W, sl.sr, s)) return -1;
    HRN_CUDA_OK(cudaEventRecord(sl.ev_kernels, s));
    HRN_CUDA_OK(cudaStreamWaitEvent(h->copy_out, sl.ev_kernels, 0));
    HRN_CUDA_OK(cudaMemcpyAsync(sr_host, sl.sr, n_out * 4, cudaMemcpyDeviceToHost, h->copy_out));
    HRN_CUDA_OK(cudaEventRecord(sl.ev_out, h->copy_out));
    sl.ticket = h->next_ticket++;
    *ticket = sl.ticket;
    return 0;
}

int32_t hrn_forward_host_wait(hrn_handle* h, int64_t ticket) {
    if (h == nullptr) {
        set_error("null handle");
        return -1;
    }
    if (ticket <= 0 || ticket >= h->next_ticket) {
        set_error("hrn_forward_host_wait: unknown ticket %lld", static_cast<long long>(ticket));
        return -1;
    }
    hrn_handle::HostSlot& sl = h->slots[ticket & 1];
    if (sl.ticket != ticket) return 0;      // already retired by a later submit (or waited before): sr_host is complete
    hrn::DeviceGuard on_device(h->device);
    if (!on_device.ok) return -1;
    HRN_CUDA_OK(cudaEventSynchronize(sl.ev_out));
    sl.ticket = 0;
    return 0;
}

int32_t hrn_u16_to_unit_float(const uint16_t* src, int64_t n, float* dst, void* stream) {
    if (src == nullptr || dst == nullptr || n < 0) {
        set_error("hrn_u16_to_unit_float: bad argument");
        return -1;
    }
    return n == 0 ? 0 : hrn::u16_to_unit_float_launch(src, static_cast<size_t>(n), dst, static_cast<cudaStream_t>(stream));
}

int32_t hrn_unit_float_to_u16(const float* src, int64_t n, uint16_t* dst, int32_t* out_of_range, void* stream) {
    if (src == nullptr || dst == nullptr || n < 0) {
        set_error("hrn_unit_float_to_u16: bad argument");
        return -1;
    }
    return n == 0 ? 0 : hrn::unit_float_to_u16_launch(src, static_cast<size_t>(n), dst, out_of_range, static_cast<cudaStream_t>(stream));
}

int32_t hrn_collate(const void* packed, int32_t packed_is_u16, const int32_t* offsets, int32_t B, int32_t min_L, int32_t H,
                    int32_t W, float* lrs, float* alphas, void* stream) {
    if (packed == nullptr || offsets == nullptr || lrs == nullptr || alphas == nullptr) {
        set_error("hrn_collate: null pointer");
        return -1;
    }
    if (B <= 0 || min_L <= 0 || H <= 0 || W <= 0) {
        set_error("hrn_collate: empty shape (B=%d min_L=%d H=%d W=%d)", B, min_L, H, W);
        return -1;
    }
    return hrn::collate_launch(packed, packed_is_u16, offsets, B, min_L, H, W, lrs, alphas, static_cast<cudaStream_t>(stream));
}

int32_t hrn_lanczos_shift(const float* img, const float* shift, int32_t Nb, int32_t C, int32_t H, int32_t W, int32_t p,
                          int32_t a, int32_t ntaps, float* out, void* stream) {
    if (img == nullptr || shift == nullptr || out == nullptr) {
        set_error("hrn_lanczos_shift: null pointer");
        return -1;
    }
    return hrn::lanczos_shift_launch(img, shift, Nb, C, H, W, p, a, ntaps, out, static_cast<cudaStream_t>(stream));
}

int32_t hrn_lanczos_taps(const float* d, int32_t n, int32_t a, int32_t ntaps, float* taps, void* stream) {
    if (d == nullptr || taps == nullptr) {
        set_error("hrn_lanczos_taps: null pointer");
        return -1;
    }
    return hrn::lanczos_taps_launch(d, n, a, ntaps, taps, static_cast<cudaStream_t>(stream));
}

int32_t hrn_shift_cpsnr(const float* sr, const float* hr, const float* hr_map, int32_t B, int32_t H, int32_t W,
                        int32_t border_w, int32_t clip_sr, float* best_db, int32_t* best_site, float* site_db,
                        void* stream) {
    if (sr == nullptr || hr == nullptr || hr_map == nullptr || best_db == nullptr || best_site == nullptr) {
        set_error("hrn_shift_cpsnr: null pointer");
        return -1;
    }
    return hrn::shift_cpsnr_launch(sr, hr, hr_map, B, H, W, border_w, clip_sr, best_db, best_site, site_db,
                                   static_cast<cudaStream_t>(stream));
}

int32_t hrn_scoring_debug_set(const char* knob, int32_t value) {
    if (knob != nullptr && strcmp(knob, "cpsnr_generic") == 0) {
        hrn::g_cpsnr_generic = value != 0;
        return 0;
    }
    if (knob != nullptr && strcmp(knob, "cpsnr_window_v1") == 0) {
        hrn::g_cpsnr_window_v1 = value;           // 1 = scalar window kernel (default), 0 = split + packed fp32x2, 2 = split, scalar
        return 0;
    }
    if (knob != nullptr && strcmp(knob, "lanczos_scalar") == 0) {
        hrn::g_lanczos_scalar = value != 0;
        return 0;
    }
    if (knob != nullptr && strcmp(knob, "cpsnr_onepass") == 0) {
        hrn::g_cpsnr_onepass = value != 0;
        return 0;
    }
    if (knob != nullptr && strcmp(knob, "cpsnr_chunk") == 0) {
        hrn::g_cpsnr_chunk = value;
        return 0;
    }
    set_error("hrn_scoring_debug_set: unknown knob");
    return -1;
}

int32_t hrn_clear_loss(const float* sr, const float* hr, const float* hr_map, int32_t B, int32_t H, int32_t W,
                       int32_t metric, float* loss, void* stream) {
    if (sr == nullptr || hr == nullptr || hr_map == nullptr || loss == nullptr) {
        set_error("hrn_clear_loss: null pointer");
        return -1;
    }
    return hrn::clear_loss_launch(sr, hr, hr_map, B, H, W, metric, loss, static_cast<cudaStream_t>(stream));
}

int32_t hrn_profile_begin(hrn_handle* h) {
    if (h == nullptr) {
        set_error("null handle");
        return -1;
    }
    for (auto& sp : h->spans) {
        cudaEventDestroy(sp.e0);
        cudaEventDestroy(sp.e1);
    }
    h->spans.clear();
    h->profiling = true;
    h->prof_error = false;
    return 0;
}

int32_t hrn_profile_end(hrn_handle* h, double* ms, double* flops, int64_t* launches) {
    if (h == nullptr || ms == nullptr || flops == nullptr || launches == nullptr) {
        set_error("hrn_profile_end: null argument");
        return -1;
    }
    h->profiling = false;
    hrn::DeviceGuard on_device(h->device);
    if (!on_device.ok) return -1;
    HRN_CUDA_OK(cudaDeviceSynchronize());
    for (int c = 0; c < HRN_PROF_CLASSES; ++c) {
        ms[c] = 0.0;
        flops[c] = 0.0;
        launches[c] = 0;
    }
    for (auto& sp : h->spans) {
        float t = 0.0f;
        HRN_CUDA_OK(cudaEventElapsedTime(&t, sp.e0, sp.e1));
        ms[sp.cls] += t;
        flops[sp.cls] += sp.flops;
        launches[sp.cls] += 1;
        cudaEventDestroy(sp.e0);
        cudaEventDestroy(sp.e1);
    }
    h->spans.clear();
    if (h->prof_error) {
        set_error("hrn_profile_end: a CUDA event call failed while profiling; the times are incomplete");
        return -1;
    }
    return 0;
}

int32_t hrn_debug_set(hrn_handle* h, const char* knob, int32_t value) {
    if (h == nullptr || knob == nullptr) {
        set_error("hrn_debug_set: null argument");
        return -1;
    }
    if (strcmp(knob, "max_ctas") == 0) h->max_ctas = value;
    else if (strcmp(knob, "debug_flags") == 0) h->debug_flags = value;
    else if (strcmp(knob, "host_chunks") == 0) h->host_chunks = value;
    else if (strcmp(knob, "workspace_mb") == 0) h->workspace_mb = value > 0 ? value : 65536;
    else if (strcmp(knob, "skip_dead_views") == 0) h->skip_dead = value != 0;
    else if (strcmp(knob, "strip_split") == 0) h->strip_split = value;
    else if (strcmp(knob, "fuse_resblock") == 0) h->fuse_resblock = value != 0;
    else if (strcmp(knob, "fuse_wave") == 0) h->fuse_wave = value != 0;
    else if (strcmp(knob, "enc_wave") == 0) h->enc_wave = value != 0;
    else if (strcmp(knob, "enc_ring_rows") == 0) {
        if (value < 12 || value > 4096) {
            set_error("hrn_debug_set: enc_ring_rows must be in [12, 4096]");
            return -1;
        }
        if (value > h->enc_ring_rows) h->enc_ring_cap[0] = h->enc_ring_cap[1] = h->enc_ring_cap[2] = h->enc_ring_cap[3] = 0;
        h->enc_ring_rows = value;
    }
    else if (strcmp(knob, "wave_streams") == 0) h->wave_streams = value;
    else if (strcmp(knob, "wave_publish_rows") == 0) h->wave_publish_rows = value < 1 ? 1 : value;
    else if (strcmp(knob, "wave_lag_rows") == 0) h->wave_lag_rows = value < 0 ? 0 : value;
    else if (strcmp(knob, "wave_stats") == 0) {
        // 1: start collecting per-CTA wait counters of the wavefront kernel; 0: print their per-role means to stderr and stop
        hrn::DeviceGuard on_device(h->device);
        const size_t n = static_cast<size_t>(h->sm_count) * 12;
        if (value) {
            if (h->wave_stats == nullptr) HRN_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->wave_stats), n * sizeof(unsigned long long)));
            HRN_CUDA_OK(cudaDeviceSynchronize());
            HRN_CUDA_OK(cudaMemset(h->wave_stats, 0, n * sizeof(unsigned long long)));
        } else if (h->wave_stats != nullptr) {
            HRN_CUDA_OK(cudaDeviceSynchronize());
            std::vector<unsigned long long> host(n);
            HRN_CUDA_OK(cudaMemcpy(host.data(), h->wave_stats, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            const char* names[5] = {"A0", "A1", "B0", "B1", "C "};
            const int streams = h->wave_streams > 0 ? h->wave_streams : h->sm_count / 5;
            fprintf(stderr, "wavefront triage (mean cycles per CTA over %d streams): kernel | TMA wait rows | TMA wait smem slot | epi wait ring space | epi wait accumulator | publisher wait rows | publisher stores | MMA wait input row\n", streams);
            for (int r = 0; r < 5; ++r) {
                double m[12] = {};
                for (int st = 0; st < streams; ++st)
                    for (int k = 0; k < 12; ++k) m[k] += static_cast<double>(host[(static_cast<size_t>(st) * 5 + r) * 12 + k]) / streams;
                fprintf(stderr, "  %s %12.0f %12.0f %12.0f %12.0f %12.0f %12.0f %12.0f %10.0f | TMA fences %10.0f TMA loop %10.0f\n", names[r], m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9]);
            }
            HRN_CUDA_OK(cudaFree(h->wave_stats));
            h->wave_stats = nullptr;
        }
    }
    else if (strcmp(knob, "wave_ring_rows") == 0) {
        if (value < 8 || value > 4096) {
            set_error("hrn_debug_set: wave_ring_rows must be in [8, 4096]");
            return -1;
        }
        if (value > h->wave_ring_rows) h->wave_ring_cap[0] = h->wave_ring_cap[1] = 0;   // regrown by the next forward
        h->wave_ring_rows = value;
    }
    else if (strcmp(knob, "mcast") == 0) h->mcast = value;        // 0 = plain, 1 = cluster pairs if they fit, 2 = required
    else if (strcmp(knob, "enc_stats") == 0) {
        // 1: start collecting per-CTA wait counters of the wavefront kernel; 0: print their per-role means to stderr and stop
        hrn::DeviceGuard on_device(h->device);
        const size_t n = static_cast<size_t>(h->sm_count) * 12;
        if (value) {
            if (h->enc_stats == nullptr) HRN_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->enc_stats), n * sizeof(unsigned long long)));
            HRN_CUDA_OK(cudaDeviceSynchronize());
            HRN_CUDA_OK(cudaMemset(h->enc_stats, 0, n * sizeof(unsigned long long)));
        } else if (h->enc_stats != nullptr) {
            HRN_CUDA_OK(cudaDeviceSynchronize());
            std::vector<unsigned long long> host(n);
            HRN_CUDA_OK(cudaMemcpy(host.data(), h->enc_stats, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            const char* names[5] = {"R0a", "R0b", "R1a", "R1b", "FIN"};
            const int streams = h->wave_streams > 0 ? h->wave_streams : h->sm_count / 5;
            fprintf(stderr, "encoder wavefront triage (mean cycles per CTA over %d streams): kernel | TMA wait rows | TMA wait smem slot | epi wait ring space | epi wait accumulator | publisher wait rows | publisher stores | MMA wait input row\n", streams);
            for (int r = 0; r < 5; ++r) {
                double m[12] = {};
                for (int st = 0; st < streams; ++st)
                    for (int k = 0; k < 12; ++k) m[k] += static_cast<double>(host[(static_cast<size_t>(st) * 5 + r) * 12 + k]) / streams;
                fprintf(stderr, "  %s %12.0f %12.0f %12.0f %12.0f %12.0f %12.0f %12.0f %10.0f | TMA fences %10.0f TMA loop %10.0f\n", names[r], m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9]);
            }
            HRN_CUDA_OK(cudaFree(h->enc_stats));
            h->enc_stats = nullptr;
        }
    }
    else if (strcmp(knob, "wave_ring_rows") == 0) {
        if (value < 8 || value > 4096) {
            set_error("hrn_debug_set: wave_ring_rows must be in [8, 4096]");
            return -1;
        }
        if (value > h->wave_ring_rows) h->wave_ring_cap[0] = h->wave_ring_cap[1] = 0;   // regrown by the next forward
        h->wave_ring_rows = value;
    }
    else if (strcmp(knob, "mcast") == 0) h->mcast = value;        // 0 = plain, 1 = cluster pairs if they fit, 2 = required
    else {
        set_error("hrn_debug_set: unknown knob '%s'", knob);
        return -1;
    }
    return 0;
}

}  // extern "C"
