// model.cu -- the C ABI of librcn_cuda.so (include/rcn_cuda.h): model handle, device buffers, host<->device
// staging, and the layer dispatch that replaces rcn's per-sample CPU loops (rcn/src/rcn.rs) with batched
// kernels (features.cu, dense.cu).
#include <cuda.h>
#include <algorithm>
#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <vector>

#include "dense.cuh"
#include "features.cuh"
#include "smallnet.cuh"
#include "opctx.cuh"
#include "dp.cuh"
#include "ozaki.cuh"

namespace rcn {

std::string& last_error_ref() {
    thread_local std::string s;
    return s;
}

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}

// ---- launch accounting / per-kernel event timing --------------------------------------------------------
namespace {
struct ProfRecord { const char* name; cudaEvent_t start, stop; };
std::atomic<unsigned long long> g_launches{0};
std::atomic<bool> g_profiling{false};
std::mutex g_prof_mu;
std::vector<ProfRecord> g_prof;
}  // namespace

LaunchScope::LaunchScope(const char* name, cudaStream_t s) : stream(s), slot(-1) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_profiling.load(std::memory_order_relaxed)) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return; }
    ProfRecord r{name, nullptr, nullptr};
    if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) { cudaGetLastError(); return; }
    cudaEventRecord(r.start, s);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
    slot = (int)g_prof.size() - 1;
}

LaunchScope::~LaunchScope() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].stop, stream);
}

// ---- streaming from a pinned HOST dataset (rcn_cuda_train_epoch_host fast path) ---------------------------------------
// State block in device memory: [0] cursor (samples, advanced by the update), [1] n_steps, [2] host image base,
// [3] prefetch launches so far, [4] ticket of the running prefetch launch, [5] pad; the step's labels follow.
// Launch k of the prefetch kernel (its own counter: it runs on a parallel graph branch and must not race with the
// cursor update at the end of step k) reads chunk k+1 straight from pinned host memory (zero-copy loads over PCIe, 16 B
// per thread, coalesced) into ring slot (k+1) % 2 while the training kernels of chunk k run on the other branch.
constexpr int kHsStateSlots = 6;
// the streamed epoch's cursor never wraps (every call re-initialises the state block), so the captured graphs are independent
// of the epoch length and serve every later call with the same batch shape
constexpr long long kHsNoWrap = 1ll << 62;
// consecutive steps captured into one CUDA graph by rcn_cuda_train_epoch_host (RCN_CUDA_HOST_STEPS_PER_GRAPH overrides)
// Default: 2 in pull mode (every step joins its prefetch branch anyway), 40 in dma mode -- the steps inside a graph follow each
// other without the host or the PCIe link in between, while every graph boundary cost 4-5 us (and now and then 27 us) with
// copies in flight on the link (profiles/r2_e2e_timeline.txt): 24.8 / 23.4 / 22.3 / 21.8 us per step at 5 / 10 / 20 / 40.
static int hs_steps_per_graph(bool dma) {
    static const int v = []() { const char* e = getenv("RCN_CUDA_HOST_STEPS_PER_GRAPH"); int n = e ? atoi(e) : 0; return n < 1 ? 0 : (n > 64 ? 64 : n); }();
    return v ? v : (dma ? 40 : 2);
}
// How the images of the streamed epoch cross PCIe (RCN_CUDA_HOST_COPY = dma | pull):
//  dma  (default) the copy engine.  The copy stream walks the dataset with back-to-back cudaMemcpyAsync calls into a ring of
//       `kHsRingSteps` chunks, an event behind each; a second stream waits for each event and publishes, with an 8-byte copy
//       from a pinned table, how many images have landed (state[3]).  Kernel A's image-loading lanes wait for their own image on
//       that counter (wait_arrived), so the compute stream carries nothing but back-to-back graph launches -- no events, no
//       per-step join.  Once per quarter of the ring the copy stream waits (cuStreamWaitValue64) until the training cursor
//       (state[0]) has passed the quarter's previous occupants.  Copies ramp 1 ... 2 ... 4 ... 8 chunks: the first step starts
//       after ONE chunk has crossed the link, later copies run at the link's large-transfer rate.  Measured on a warm link
//       (profiles/r2_pcie_dma.json): 43.6 GB/s for one 0.8 MB chunk, 48.7 for two, 53.6 for eight, against 38 GB/s for
//       SM-issued zero-copy loads; and no CTA of the step's kernels shares its SM with a copy kernel.  (Publishing the counter
//       on the copy stream itself -- cuStreamWriteValue64 or the 8-byte copy -- put ~22 us between consecutive data copies,
//       which then no longer pipeline: a one-chunk copy took 40 us instead of 18 and the copy stream became the bottleneck.)
//  pull the prefetch kernel below on a parallel branch of the step's graph (round 1's design; also the fallback when the
//       driver has no 64-bit stream memory operations or the staged front end is not the one selected).
static bool hs_dma_env() {
    static const bool v = []() { const char* e = getenv("RCN_CUDA_HOST_COPY"); return !(e && (e[0] == 'p' || e[0] == 'P')); }();
    return v;
}
constexpr size_t kHsRingSteps = 64;   // dma: chunks the ring holds
constexpr size_t kHsQuarter = kHsRingSteps / 4;   // dma: granularity of the copy stream's wait for free ring slots
constexpr size_t kHsMaxCopy = 8;      // dma: chunks per cudaMemcpyAsync once the ramp is over (divides kHsQuarter)
constexpr int kHsEvents = 64;         // dma: one event behind each copy, reused round robin
struct StreamMemOps {
    CUresult (*wait64)(CUstream, CUdeviceptr, cuuint64_t, unsigned int) = nullptr;

    bool ok = false;
};
static const StreamMemOps& stream_mem_ops(int device) {
    static const StreamMemOps ops = [device]() {
        StreamMemOps o;
        auto entry = [](const char* name) -> void* {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
            return p;
        };
        auto get_attr = reinterpret_cast<CUresult (*)(int*, CUdevice_attribute, CUdevice)>(entry("cuDeviceGetAttribute"));
        o.wait64 = reinterpret_cast<decltype(o.wait64)>(entry("cuStreamWaitValue64"));

        int can = 0;
        o.ok = get_attr && o.wait64 && get_attr(&can, CU_DEVICE_ATTRIBUTE_CAN_USE_64_BIT_STREAM_MEM_OPS, device) == CUDA_SUCCESS && can;
        return o;
    }();
    return ops;
}
// PCIe pulls are latency-bound per request: every thread first issues ALL the 16-byte loads of its round (U independent
// zero-copy reads in flight, the whole 0.8 MB chunk of the canonical config in one round trip) and only then stores them.
// (Written as `d4[i] = s4[i]` the compiler must keep load/store pairs in order -- source and ring may alias as far as it
// knows -- which leaves ONE load in flight per thread and costs a PCIe round trip per 128 KB.)
template <int U>
__global__ void __launch_bounds__(256) host_prefetch_kernel(long long* __restrict__ state, unsigned char* __restrict__ ring,
                                                            long long B, long long img_bytes) {
    const long long k1 = *reinterpret_cast<volatile long long*>(state + 3) + 1;
    if (k1 < state[1]) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(state[2]) + k1 * B * img_bytes;
        unsigned char* dst = ring + ((k1 * B) % (2 * B)) * img_bytes;
        const long long n16 = B * img_bytes / 16;
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        const long long stride = (long long)gridDim.x * blockDim.x;
        for (long long base = blockIdx.x * (long long)blockDim.x + threadIdx.x; base < n16; base += stride * U) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (base + u * stride < n16) v[u] = s4[base + u * stride];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (base + u * stride < n16) d4[base + u * stride] = v[u];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {   // the last CTA to finish publishes the counter: every CTA has read it by then
        __threadfence();
        if (atomicAdd(reinterpret_cast<unsigned long long*>(state + 4), 1ull) == gridDim.x - 1) {
            state[4] = 0;
            *reinterpret_cast<volatile long long*>(state + 3) = k1;
        }
    }
}
// RCN_CUDA_PREFETCH_UNROLL = loads in flight per thread (1 = the load/store-paired loop, default 8);
// RCN_CUDA_PREFETCH_CTAS = grid size (default 32: the SMs the 128-CTA step kernels leave free, and then some)
static int prefetch_unroll() {
    static const int v = []() { const char* e = getenv("RCN_CUDA_PREFETCH_UNROLL"); int n = e ? atoi(e) : 8; return n <= 1 ? 1 : (n <= 4 ? 4 : 8); }();
    return v;
}
static int prefetch_ctas() {
    static const int v = []() { const char* e = getenv("RCN_CUDA_PREFETCH_CTAS"); int n = e ? atoi(e) : 32; return n < 1 ? 1 : (n > 1024 ? 1024 : n); }();
    return v;
}
static void launch_host_prefetch(long long* state, unsigned char* ring, long long B, long long img_bytes, cudaStream_t s) {
    const int g = prefetch_ctas();
    switch (prefetch_unroll()) {
        case 1: host_prefetch_kernel<1><<<g, 256, 0, s>>>(state, ring, B, img_bytes); break;
        case 4: host_prefetch_kernel<4><<<g, 256, 0, s>>>(state, ring, B, img_bytes); break;
        default: host_prefetch_kernel<8><<<g, 256, 0, s>>>(state, ring, B, img_bytes); break;
    }
}

bool is_pinned_host_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost && a.devicePointer != nullptr;
}


}  // namespace rcn

using namespace rcn;

struct rcn_cuda_model {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    size_t classes = 0;
    std::vector<int32_t> cfg;
    std::vector<size_t> ff;

    // parameters: flat [W0|b0|W1|b1|...], W_l column-major rows[l] x cols[l] (serialization.rs:19-22 order)
    bool params_ready = false;
    std::vector<size_t> rows, cols, w_off, b_off;
    size_t n_params = 0, sum_rows = 0;
    DevBuf params, grads_own;
    double* grads = nullptr;
    bool grads_bound = false;
    double mean = 1.0, sd = 1.0;  // scale_set starts at (1, 1)  (rcn.rs:71)
    SmallNetDesc small_desc{};      // fused path for narrow networks (smallnet.cu)
    bool use_small = false;

    // feature plan cache
    bool plan_valid = false;
    size_t plan_H = 0, plan_W = 0;
    FeaturePlan plan;
    FeatureScratch fscratch;

    // scratch (grow-only)
    DevBuf in_stage, tgt_stage, feats, acts, deltas, gemm_ws, out_stage, small, red_ws;
    ReduceScratch rs;           // deterministic two-stage reductions (bias gradient, batch statistics)
    // epoch mode (rcn.rs:144-149 on a resident dataset)
    const void* ep_images = nullptr;
    const int64_t* ep_labels = nullptr;
    const int64_t* ep_perm = nullptr;
    int ep_fmt = 0;
    size_t ep_n = 0, ep_H = 0, ep_W = 0, ep_B = 0;
    DevBuf ep_state;            // kEpSlots int64 state words (features.cuh: the cursor) | labels_batch [B]
    size_t last_B = 0;          // batch of the last accumulate call (taps)
    bool stats_valid = false;
    // data-parallel group (dp.cu) and the pipelined host-dataset loop (rcn_cuda_train_epoch_host)
    DpState dp;
    SnUpdate pending_upd{};         // set by the step entry points: the next small-network accumulate applies the update itself
    bool upd_fused = false;         // ... and did (the standalone update kernel is then skipped)
    bool cursor_fused = false;      // ... or at least advanced the cursor / wrote the result ring (data-parallel groups)
    bool x_owns_cursor = false;     // an exchange kernel that advances a cursor may be in flight: kernel A must not read the
                                    // cursor ahead of its griddepcontrol.wait (cleared wherever the stream is drained)
    bool dp_pushed = false;         // the last accumulate pushed its gradients to the peers itself (kernel B epilogue)
    bool dp_push_suppress = false;  // warm-up launches must not push (a push is consumed by exactly one receive)
    DevBuf timeline;                // device-side launch timeline (timeline.cuh), allocated by rcn_cuda_timeline_enable
    bool tl_on = false;
    OzakiWorkspace oz;              // tcgen05 integer-slice GEMM scratch (wide dense layers)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    DevBuf host_slot[2];
    double* stats_host = nullptr;   // pinned, 2 doubles per step
    size_t stats_host_cap = 0;
    // streaming variant: the GPU pulls chunk k+1 from pinned host memory while it trains on chunk k; one graph per step
    DevBuf hs_ring, hs_state;
    cudaStream_t flag_stream = nullptr;     // dma mode: publishes the arrival counter behind each copy
    cudaEvent_t hs_ev[kHsEvents] = {};
    long long* hs_counts = nullptr;         // pinned: "images arrived" value behind each copy of the streamed epoch (dma mode)
    size_t hs_counts_cap = 0;
    std::map<int, cudaGraphExec_t> hs_graphs;   // streamed epoch: g consecutive steps per graph, captured on first use
    cudaEvent_t hs_fork = nullptr, hs_join = nullptr;
    struct HsKey {
        size_t B = 0, H = 0, W = 0, n_steps = 0;
        unsigned long long alloc_gen = 0;   // no device buffer has moved since the capture (common.cuh)
        uint64_t mean_bits = 0, sd_bits = 0; // scale_set travels BY VALUE in the captured kernel arguments
        double scale = 0.0;
        const void *labels = nullptr, *stats = nullptr, *ring = nullptr, *state = nullptr, *grads = nullptr;
        cudaStream_t stream = nullptr;
        bool dp = false;
        long long window = 0;               // ring size in images (captured in kernel A's arguments)
        bool operator==(const HsKey& o) const {
            return B == o.B && H == o.H && W == o.W && n_steps == o.n_steps && alloc_gen == o.alloc_gen && mean_bits == o.mean_bits &&
                   sd_bits == o.sd_bits && scale == o.scale && labels == o.labels &&
                   stats == o.stats && ring == o.ring && state == o.state && grads == o.grads && stream == o.stream && dp == o.dp && window == o.window;
        }
    } hs_key;

    double* act(size_t l, size_t B) const {
        size_t off = 0;
        for (size_t j = 0; j < l; ++j) off += rows[j];
        return acts.as<double>() + off * B;
    }
    double* delta(size_t l, size_t B) const {
        size_t off = 0;
        for (size_t j = 0; j < l; ++j) off += rows[j];
        return deltas.as<double>() + off * B;
    }
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() { /* leave the device selected: one process drives one GPU */ }
};

#define RCN_ENTER(h)                                                                     \
    if (!(h)) return fail(RCN_ERR_INVALID, "null model handle");                         \
    DeviceGuard _guard((h)->device);                                                     \
    if (!_guard.ok) return fail(RCN_ERR_CUDA, "cudaSetDevice(%d) failed", (h)->device);

bool dp_fused_push_enabled() {
    static const bool on = []() { const char* e = getenv("RCN_CUDA_DP_FUSED_PUSH"); return !(e && e[0] == '0'); }();
    return on;
}

// Consumes the "already pushed" flag of the last accumulate: every push is matched by exactly one receive.
bool take_dp_pushed(rcn_cuda_model* h) {
    const bool p = h->dp_pushed;
    h->dp_pushed = false;
    return p;
}

int ensure_plan(rcn_cuda_model* h, size_t H, size_t W) {
    if (h->plan_valid && h->plan_H == H && h->plan_W == W) return RCN_OK;
    if (H == 0 || W == 0 || H > 32768 || W > 32768) return fail(RCN_ERR_INVALID, "bad image size %zu x %zu", H, W);
    RCN_TRY(plan_features(h->cfg.data(), h->cfg.size(), H, W, &h->plan));
    h->plan_valid = true; h->plan_H = H; h->plan_W = W;
    return RCN_OK;
}

size_t pixel_bytes(int fmt) { return fmt == RCN_PIXELS_U8_ROWMAJOR ? 1 : 8; }

// Copies a device result to a possibly-host destination; `synced` is set when the stream was drained.
int deliver(rcn_cuda_model* h, void* dst, const void* dev_src, size_t bytes) {
    if (bytes == 0) return RCN_OK;
    if (is_device_ptr(dst)) {
        if (dst != dev_src) RCN_CUDA_TRY(cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToDevice, h->stream));
        return RCN_OK;
    }
    RCN_CUDA_TRY(cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToHost, h->stream));
    RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return RCN_OK;
}

int require_params(rcn_cuda_model* h) {
    if (!h->params_ready) return fail(RCN_ERR_STATE, "parameters not initialised: call rcn_cuda_init_params first");
    return RCN_OK;
}

// features (+ optional standardise) into h->feats for device or host images
int features_into(rcn_cuda_model* h, const void* images, int fmt, size_t B, size_t H, size_t W, bool standardise,
                  double* out_dev) {
    RCN_TRY(ensure_plan(h, H, W));
    if (fmt != RCN_PIXELS_U8_ROWMAJOR && fmt != RCN_PIXELS_F64_COLMAJOR) return fail(RCN_ERR_INVALID, "unknown pixel format %d", fmt);
    if (B == 0 || h->plan.L == 0) return RCN_OK;
    if (!images) return fail(RCN_ERR_INVALID, "null images");
    StagedIn in;
    RCN_TRY(in.stage(images, B * H * W * pixel_bytes(fmt), h->in_stage, h->stream, nullptr));
    return launch_features(h->plan, in.dev, fmt, B, H, W, standardise, h->mean, h->sd, out_dev, h->fscratch, h->stream);
}

// forward pass over device feats (cols[0] x B); fills acts; last layer also fills deltas when targets given
int forward_dev(rcn_cuda_model* h, const double* feats, size_t B, const double* onehot, const int64_t* labels,
                bool want_delta) {
    const size_t n = h->rows.size();
    RCN_TRY(h->acts.reserve(h->sum_rows * B * sizeof(double)));
    if (want_delta) RCN_TRY(h->deltas.reserve(h->sum_rows * B * sizeof(double)));
    if (h->use_small && !want_delta)
        return launch_smallnet_forward(h->small_desc, h->params.as<double>(), const_cast<double*>(feats), B,
                                       h->acts.as<double>(), nullptr, h->stream);
    for (size_t l = 0; l < n; ++l) {
        const double* a_in = l == 0 ? feats : h->act(l - 1, B);
        const bool last = (l + 1 == n);
        RCN_TRY(launch_dense_forward(h->params.as<double>() + h->w_off[l], h->params.as<double>() + h->b_off[l], a_in,
                                     h->rows[l], h->cols[l], B, h->act(l, B),
                                     (last && want_delta) ? h->delta(l, B) : nullptr, onehot, labels, h->stream, &h->oz));
    }
    return RCN_OK;
}

int accumulate_dev(rcn_cuda_model* h, const double* feats, const double* onehot, const int64_t* labels, size_t B,
                   const SmallNetFront* front = nullptr) {
    const size_t n = h->rows.size();
    const SnUpdate upd = h->pending_upd;   // consumed by this call whichever path it takes
    h->pending_upd = SnUpdate{};
    h->upd_fused = false;
    if (h->dp_pushed)   // a pushed gradient must be received by the group's update before the next one is produced
        return fail(RCN_ERR_STATE, "data-parallel group: the previous gradients were pushed to the peers but never applied "
                                   "(every accumulate needs exactly one apply on every rank)");
    if (B == 0) {
        RCN_CUDA_TRY(cudaMemsetAsync(h->grads, 0, h->n_params * sizeof(double), h->stream));
        return RCN_OK;
    }
    if (h->use_small && B <= smallnet_max_batch()) {
        RCN_TRY(h->acts.reserve(h->sum_rows * B * sizeof(double)));
        RCN_TRY(h->deltas.reserve(h->sum_rows * B * sizeof(double)));
        RCN_TRY(h->small.reserve(64));
        DpPush push{};
        const bool fuse_push = h->dp.connected && dp_fused_push_enabled() && !h->dp_push_suppress;
        if (fuse_push) push = dp_push_desc(h->dp);
        RCN_TRY(launch_smallnet_backprop(h->small_desc, h->params.as<double>(), const_cast<double*>(feats), B, onehot,
                                         labels, h->acts.as<double>(), h->deltas.as<double>(), h->grads,
                                         h->small.as<double>(), h->gemm_ws, front, h->stream, fuse_push ? &push : nullptr,
                                         ((upd.params || upd.cursor) && (!h->dp.connected || fuse_push)) ? &upd : nullptr));
        h->upd_fused = upd.params && (!h->dp.connected || fuse_push);
        h->cursor_fused = upd.cursor && (!h->dp.connected || fuse_push);
        h->dp_pushed = fuse_push && !h->upd_fused;   // pushed but not yet received: the exchange kernel must follow
        h->stats_valid = true;
        h->last_B = B;
        return RCN_OK;
    }
    if (front) return fail(RCN_ERR_STATE, "internal: fused front end requested on the generic path");
    RCN_TRY(forward_dev(h, feats, B, onehot, labels, true));
    for (size_t l = n - 1; l-- > 0;)  // rcn.rs:305-311
        RCN_TRY(launch_dense_backward_data(h->params.as<double>() + h->w_off[l + 1], h->delta(l + 1, B), h->act(l, B),
                                           h->rows[l], h->rows[l + 1], B, h->delta(l, B), h->stream, &h->oz));
    for (size_t l = 0; l < n; ++l) {
        const double* a_prev = l == 0 ? feats : h->act(l - 1, B);
        RCN_TRY(launch_dense_backward_weight(h->delta(l, B), a_prev, h->rows[l], h->cols[l], B, h->grads + h->w_off[l],
                                             h->grads + h->b_off[l], h->gemm_ws, h->rs, h->stream, &h->oz));
    }
    RCN_TRY(h->small.reserve(64));
    RCN_TRY(launch_batch_stats(h->act(n - 1, B), h->rows[n - 1], B, onehot, labels, h->small.as<double>(), h->rs, h->stream));
    h->stats_valid = true;
    h->last_B = B;
    return RCN_OK;
}

// images (DEVICE) -> gradient sums.  The canonical narrow network with u8 pixels runs the convpool stack inside
// kernel A (smallnet.cu); everything else runs the feature kernel(s) first.  `bi` (optional) is the epoch-mode
// device-side batch selection; without it `labels` are this batch's labels.
// How far the single-GPU prewait goes (kernel A's front end ahead of griddepcontrol.wait, under the previous kernel B's tail;
// RCN_CUDA_PREWAIT): 0 = nowhere, 1 = in the streamed host epoch only (default), 2 = in every training step.  Measured: the host
// epoch gains 3-6 % (the arrival-counter wait and the ring loads move under kernel B); the device-resident step does not (kernel
// A needs a whole SM's registers, so all but ~26 of its CTAs start only when kernel B's CTAs leave, and its parameter loads are
// no longer hidden behind the front end).
static int prewait_mode() {
    static const int v = []() { const char* e = getenv("RCN_CUDA_PREWAIT"); return e ? atoi(e) : 1; }();
    return v;
}

static bool fused_front_enabled() {
    static const bool v = []() { const char* e = getenv("RCN_CUDA_FUSED_FRONT"); return !(e && e[0] == '0'); }();
    return v;
}

int accumulate_images_dev(rcn_cuda_model* h, const void* images, int fmt, const int64_t* labels, size_t B, size_t H,
                          size_t W, const BatchIndex* bi) {
    RCN_TRY(h->feats.reserve(h->plan.L * B * sizeof(double)));
    const int64_t* step_labels = bi ? (const int64_t*)bi->labels_batch : labels;
    if (fused_front_enabled() && h->use_small && B <= smallnet_max_batch() && fmt == RCN_PIXELS_U8_ROWMAJOR && h->plan.n_conv <= 10 &&
        h->plan.L > 0) {
        SmallNetFront fr{};
        fr.images = (const uint8_t*)images;
        fr.H = (int)H; fr.W = (int)W;
        fr.max_elems = (int)h->plan.max_elems;
        fr.stages = h->plan.stages;
        fr.sc = make_standardise(h->plan, fmt, true, h->mean, h->sd);
        if (bi) fr.bi = *bi;
        smallnet_front_select(h->plan, &fr);
        // Data-parallel step behind the exchange kernel: kernel A may run its front end ahead of griddepcontrol.wait
        // (smallnet.cu) when this step's kernel B pushes to the peers (so the exchange kernel X precedes the next kernel
        // A), the update is not folded into kernel B, and any device-side cursor it reads is advanced by kernel B -- this
        // step's and the previous one's (x_owns_cursor) -- not by X, which may still be running then.
        static const bool prewait_env = []() { const char* e = getenv("RCN_CUDA_DP_PREWAIT"); return !(e && e[0] == '0'); }();
        const bool fuse_push = h->dp.connected && dp_fused_push_enabled() && !h->dp_push_suppress;
        fr.prewait = (prewait_env && fr.use_cp && fuse_push && !h->pending_upd.params && !h->x_owns_cursor &&
                      (!fr.bi.cursor || h->pending_upd.cursor == fr.bi.cursor)) ? 1 : 0;
        // One GPU: the same front end under the tail of the previous step's kernel B (which triggers its dependents early).
        // Safe whatever precedes this launch: a kernel that does not trigger early has completed before kernel A starts; kernel B
        // advances a device-side cursor before it triggers, and the only buffer both touch -- the features -- is written by
        // this kernel A after its wait.
        if (!h->dp.connected && fr.use_cp && (prewait_mode() == 2 || (prewait_mode() == 1 && fr.bi.window != 0))) fr.prewait = 2;
        if (smallnet_front_fits(h->small_desc, fr))
            return accumulate_dev(h, h->feats.as<double>(), nullptr, step_labels, B, &fr);
    }
    RCN_TRY(launch_features(h->plan, images, fmt, B, H, W, true, h->mean, h->sd, h->feats.as<double>(), h->fscratch,
                            h->stream, bi));
    return accumulate_dev(h, h->feats.as<double>(), nullptr, step_labels, B);
}

// stage targets (exactly one of onehot / labels)
int stage_targets(rcn_cuda_model* h, const double* onehot, const int64_t* labels, size_t B, const double** oh_dev,
                  const int64_t** lb_dev) {
    *oh_dev = nullptr; *lb_dev = nullptr;
    if ((onehot == nullptr) == (labels == nullptr))
        return fail(RCN_ERR_INVALID, "exactly one of onehot / labels must be given");
    StagedIn in;
    if (onehot) {
        RCN_TRY(in.stage(onehot, h->classes * B * sizeof(double), h->tgt_stage, h->stream, nullptr));
        *oh_dev = (const double*)in.dev;
    } else {
        RCN_TRY(in.stage(labels, B * sizeof(int64_t), h->tgt_stage, h->stream, nullptr));
        *lb_dev = (const int64_t*)in.dev;
    }
    return RCN_OK;
}

int check_feature_width(rcn_cuda_model* h, size_t L) {
    if (h->cols[0] != L)  // nalgebra `w * a` dimension mismatch (reference panics; see rcn.rs:443 quirk)
        return fail(RCN_ERR_SHAPE, "Matrix multiplication dimensions mismatch: first layer expects %zu inputs, feature vector has %zu",
                    h->cols[0], L);
    return RCN_OK;
}

}  // namespace

extern "C" {

const char* rcn_cuda_last_error(void) { return last_error_ref().c_str(); }
int rcn_cuda_version(void) { return 100; }

int rcn_cuda_device_count(int* count) {
    if (!count) return fail(RCN_ERR_INVALID, "null count");
    RCN_CUDA_TRY(cudaGetDeviceCount(count));
    return RCN_OK;
}

int rcn_cuda_create(size_t classes, const int32_t* convpool_cfg, size_t n_convpool, const size_t* feedforward_cfg,
                    size_t n_feedforward, int device, rcn_cuda_handle* out) {
    if (!out) return fail(RCN_ERR_INVALID, "null out handle");
    *out = nullptr;
    if (n_convpool && !convpool_cfg) return fail(RCN_ERR_INVALID, "null convpool_cfg");
    if (n_feedforward && !feedforward_cfg) return fail(RCN_ERR_INVALID, "null feedforward_cfg");
    for (size_t i = 0; i < n_convpool; ++i)
        if (convpool_cfg[i] < 0 || convpool_cfg[i] > 3) return fail(RCN_ERR_INVALID, "unknown RCNLayer code %d", convpool_cfg[i]);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(RCN_ERR_CUDA, "no CUDA device available (%s); librcn_cuda has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= ndev) return fail(RCN_ERR_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    RCN_CUDA_TRY(cudaSetDevice(device));
    std::unique_ptr<rcn_cuda_model> m(new (std::nothrow) rcn_cuda_model());
    if (!m) return fail(RCN_ERR_INVALID, "out of host memory");
    m->device = device;
    m->classes = classes;
    m->cfg.assign(convpool_cfg, convpool_cfg + n_convpool);
    m->ff.assign(feedforward_cfg, feedforward_cfg + n_feedforward);
    RCN_CUDA_TRY(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
    m->stream = m->own_stream;
    *out = m.release();
    return RCN_OK;
}

int rcn_cuda_destroy(rcn_cuda_handle h) {
    if (!h) return RCN_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->own_stream) { cudaStreamSynchronize(h->own_stream); cudaStreamDestroy(h->own_stream); }
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    for (int i = 0; i < 2; ++i) {
        if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
        if (h->ev_consumed[i]) cudaEventDestroy(h->ev_consumed[i]);
        h->host_slot[i].release();
    }
    if (h->stats_host) cudaFreeHost(h->stats_host);
    for (auto& kv : h->hs_graphs) cudaGraphExecDestroy(kv.second);
    if (h->hs_fork) cudaEventDestroy(h->hs_fork);
    if (h->hs_join) cudaEventDestroy(h->hs_join);

    h->hs_ring.release(); h->hs_state.release();
    if (h->hs_counts) cudaFreeHost(h->hs_counts);
    if (h->flag_stream) cudaStreamDestroy(h->flag_stream);
    for (int i = 0; i < kHsEvents; ++i)
        if (h->hs_ev[i]) cudaEventDestroy(h->hs_ev[i]);

    dp_release(h->dp);
    h->oz.release();
    DevBuf* bufs[] = {&h->params, &h->grads_own, &h->in_stage, &h->tgt_stage, &h->feats, &h->acts, &h->deltas,
                      &h->gemm_ws, &h->out_stage, &h->small, &h->red_ws, &h->ep_state, &h->fscratch.a, &h->fscratch.b, &h->rs.buf};
    for (DevBuf* b : bufs) b->release();
    delete h;
    return RCN_OK;
}

int rcn_cuda_set_stream(rcn_cuda_handle h, void* cuda_stream) {
    RCN_ENTER(h);
    h->stream = (cuda_stream == RCN_STREAM_OWN) ? h->own_stream : (cudaStream_t)cuda_stream;
    return RCN_OK;
}

static void stream_drained(rcn_cuda_model* h);

int rcn_cuda_synchronize(rcn_cuda_handle h) {
    RCN_ENTER(h);
    RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));
    stream_drained(h);
    return RCN_OK;
}

int rcn_cuda_feature_shape(rcn_cuda_handle h, size_t H, size_t W, size_t* n_maps, size_t* map_h, size_t* map_w) {
    RCN_ENTER(h);
    RCN_TRY(ensure_plan(h, H, W));
    if (n_maps) *n_maps = h->plan.n_maps;
    if (map_h) *map_h = h->plan.map_h;
    if (map_w) *map_w = h->plan.map_w;
    return RCN_OK;
}

namespace {
// Installs explicit layer shapes (rows[i] x cols[i]), allocates zeroed parameters / gradients, picks the fused path.
int install_shapes(rcn_cuda_model* h, const std::vector<size_t>& rows, const std::vector<size_t>& cols) {
    const size_t n = rows.size();
    // a captured step holds the layer shapes by value (SmallNetDesc): new shapes invalidate it like a moved buffer does
    alloc_generation().fetch_add(1, std::memory_order_relaxed);
    h->rows.clear(); h->cols.clear(); h->w_off.clear(); h->b_off.clear();
    size_t off = 0, sum_rows = 0;
    for (size_t i = 0; i < n; ++i) {
        if (rows[i] == 0) return fail(RCN_ERR_INVALID, "layer %zu has zero neurons", i);
        h->rows.push_back(rows[i]); h->cols.push_back(cols[i]);  // dims = (output_size, input_size)  (rcn.rs:502)
        h->w_off.push_back(off); off += rows[i] * cols[i];
        h->b_off.push_back(off); off += rows[i];
        sum_rows += rows[i];
    }
    h->n_params = off; h->sum_rows = sum_rows;
    {   // fused small-network path (smallnet.cu) unless RCN_CUDA_SMALLNET=0
        SmallNetDesc d{};
        h->use_small = false;
        const char* env = getenv("RCN_CUDA_SMALLNET");
        if (n <= (size_t)kSmallNetMaxLayers && h->cols[0] >= 1 && h->cols[0] < (1u << 24) && off < (1u << 30) &&
            !(env && env[0] == '0')) {
            d.n_layers = (int)n; d.n_in = (int)h->cols[0]; d.n_params = (int)off;
            for (size_t i = 0; i < n; ++i) { d.rows[i] = (int)h->rows[i]; d.w_off[i] = (int)h->w_off[i]; d.b_off[i] = (int)h->b_off[i]; }
            bool ok = true;
            for (size_t i = 0; i < n; ++i) ok = ok && h->rows[i] <= 32;
            if (ok && smallnet_eligible(d)) { h->small_desc = d; h->small_desc.tl = h->tl_on ? h->timeline.as<Timeline>() : nullptr; h->use_small = true; }
        }
    }
    RCN_TRY(h->params.reserve(off * sizeof(double)));
    RCN_CUDA_TRY(cudaMemsetAsync(h->params.p, 0, off * sizeof(double), h->stream));
    if (!h->grads_bound) {
        RCN_TRY(h->grads_own.reserve(off * sizeof(double)));
        h->grads = h->grads_own.as<double>();
        RCN_CUDA_TRY(cudaMemsetAsync(h->grads, 0, off * sizeof(double), h->stream));
    }
    h->params_ready = true;
    h->stats_valid = false;
    return RCN_OK;
}
}  // namespace

int rcn_cuda_init_params(rcn_cuda_handle h, size_t l) {
    RCN_ENTER(h);
    // load_weights_and_bias (rcn.rs:425-457)
    if (h->ff.empty()) return fail(RCN_ERR_OUT_OF_BOUNDS, "index out of bounds: feedforward_cfg is empty (rcn.rs:444)");
    unsigned c = 0, p = 0;
    for (int32_t layer : h->cfg) {
        if (layer == RCN_LAYER_CONV_NONE || layer == RCN_LAYER_CONV_SAME) c += 1; else p += 2;
    }
    if (c > 20 || p > 60) return fail(RCN_ERR_INVALID, "convpool stack too deep");
    size_t pc = 1, pp = 1;
    for (unsigned i = 0; i < c; ++i) pc *= 4;
    for (unsigned i = 0; i < p; ++i) pp *= 2;
    size_t a = pc / pp * l;  // 4^c / 2^p * l, integer division first (rcn.rs:443)
    size_t b = h->ff[0];
    const size_t n = h->ff.size() + 1;
    std::vector<size_t> rows, cols;
    for (size_t i = 0; i < n; ++i) {
        rows.push_back(b); cols.push_back(a);
        a = b;
        b = (i + 1 < h->ff.size()) ? h->ff[i + 1] : h->classes;
    }
    return install_shapes(h, rows, cols);
}

// A deserialized model carries whatever matrices its checkpoint holds (main.rs:50): install explicit shapes.
int rcn_cuda_init_params_shapes(rcn_cuda_handle h, const size_t* rows, const size_t* cols, size_t n_layers) {
    RCN_ENTER(h);
    if (!rows || !cols || n_layers == 0) return fail(RCN_ERR_INVALID, "need at least one layer shape");
    for (size_t i = 1; i < n_layers; ++i)
        if (cols[i] != rows[i - 1])
            return fail(RCN_ERR_SHAPE, "Matrix multiplication dimensions mismatch: layer %zu has %zu inputs, layer %zu has %zu outputs", i,
                        cols[i], i - 1, rows[i - 1]);
    return install_shapes(h, std::vector<size_t>(rows, rows + n_layers), std::vector<size_t>(cols, cols + n_layers));
}

int rcn_cuda_num_layers(rcn_cuda_handle h, size_t* n_layers) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (n_layers) *n_layers = h->rows.size();
    return RCN_OK;
}

int rcn_cuda_layer_shape(rcn_cuda_handle h, size_t layer, size_t* rows, size_t* cols) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (layer >= h->rows.size()) return fail(RCN_ERR_INVALID, "layer %zu out of range", layer);
    if (rows) *rows = h->rows[layer];
    if (cols) *cols = h->cols[layer];
    return RCN_OK;
}

int rcn_cuda_param_count(rcn_cuda_handle h, size_t* n) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (n) *n = h->n_params;
    return RCN_OK;
}

static int copy_in(rcn_cuda_handle h, double* dst_dev, const double* src, size_t n) {
    if (!src) return fail(RCN_ERR_INVALID, "null source");
    if (n == 0) return RCN_OK;
    const bool dev = is_device_ptr(src);
    RCN_CUDA_TRY(cudaMemcpyAsync(dst_dev, src, n * sizeof(double), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    if (!dev) RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));  // pageable source may be reused by the caller
    return RCN_OK;
}

int rcn_cuda_set_weights(rcn_cuda_handle h, size_t layer, size_t rows, size_t cols, const double* w) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (layer >= h->rows.size()) return fail(RCN_ERR_INVALID, "layer %zu out of range", layer);
    if (rows != h->rows[layer] || cols != h->cols[layer])
        return fail(RCN_ERR_SHAPE, "weights for layer %zu must be %zu x %zu, got %zu x %zu", layer, h->rows[layer], h->cols[layer], rows, cols);
    return copy_in(h, h->params.as<double>() + h->w_off[layer], w, rows * cols);
}

int rcn_cuda_get_weights(rcn_cuda_handle h, size_t layer, double* w) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (layer >= h->rows.size() || !w) return fail(RCN_ERR_INVALID, "bad layer / null destination");
    return deliver(h, w, h->params.as<double>() + h->w_off[layer], h->rows[layer] * h->cols[layer] * sizeof(double));
}

int rcn_cuda_set_bias(rcn_cuda_handle h, size_t layer, size_t n, const double* b) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (layer >= h->rows.size()) return fail(RCN_ERR_INVALID, "layer %zu out of range", layer);
    if (n != h->rows[layer]) return fail(RCN_ERR_SHAPE, "bias for layer %zu must have %zu entries, got %zu", layer, h->rows[layer], n);
    return copy_in(h, h->params.as<double>() + h->b_off[layer], b, n);
}

int rcn_cuda_get_bias(rcn_cuda_handle h, size_t layer, double* b) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (layer >= h->rows.size() || !b) return fail(RCN_ERR_INVALID, "bad layer / null destination");
    return deliver(h, b, h->params.as<double>() + h->b_off[layer], h->rows[layer] * sizeof(double));
}

int rcn_cuda_set_params(rcn_cuda_handle h, const double* flat, size_t n) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (n != h->n_params) return fail(RCN_ERR_SHAPE, "model has %zu parameters, got %zu", h->n_params, n);
    return copy_in(h, h->params.as<double>(), flat, n);
}

int rcn_cuda_get_params(rcn_cuda_handle h, double* flat, size_t n) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (n != h->n_params || !flat) return fail(RCN_ERR_SHAPE, "model has %zu parameters, got %zu", h->n_params, n);
    return deliver(h, flat, h->params.p, n * sizeof(double));
}

int rcn_cuda_set_scale(rcn_cuda_handle h, double mean, double sd) {
    RCN_ENTER(h);
    h->mean = mean; h->sd = sd;
    return RCN_OK;
}

int rcn_cuda_get_scale(rcn_cuda_handle h, double* mean, double* sd) {
    RCN_ENTER(h);
    if (mean) *mean = h->mean;
    if (sd) *sd = h->sd;
    return RCN_OK;
}

int rcn_cuda_features(rcn_cuda_handle h, const void* images, int pixel_format, size_t B, size_t H, size_t W,
                      int standardise, double* out) {
    RCN_ENTER(h);
    RCN_TRY(ensure_plan(h, H, W));
    const size_t bytes = h->plan.L * B * sizeof(double);
    if (bytes == 0) return RCN_OK;
    if (!out) return fail(RCN_ERR_INVALID, "null output");
    double* out_dev = out;
    const bool out_is_dev = is_device_ptr(out);
    if (!out_is_dev) { RCN_TRY(h->feats.reserve(bytes)); out_dev = h->feats.as<double>(); }
    RCN_TRY(features_into(h, images, pixel_format, B, H, W, standardise != 0, out_dev));
    if (!out_is_dev) RCN_TRY(deliver(h, out, out_dev, bytes));
    return RCN_OK;
}

int rcn_cuda_gen_scales(rcn_cuda_handle h, const double* feats, size_t L, size_t B, double* mean, double* sd) {
    RCN_ENTER(h);
    if (!feats || L * B == 0) return fail(RCN_ERR_INVALID, "gen_scales needs a non-empty set");
    StagedIn in;
    RCN_TRY(in.stage(feats, L * B * sizeof(double), h->in_stage, h->stream, nullptr));
    RCN_TRY(h->small.reserve(64));
    double* res = h->small.as<double>() + 4;
    RCN_TRY(launch_gen_scales((const double*)in.dev, L * B, res, h->red_ws, h->stream));
    double host[2];
    RCN_CUDA_TRY(cudaMemcpyAsync(host, res, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
    RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->mean = host[0]; h->sd = host[1];  // self.scale_set = (mean, sd)  (rcn.rs:249-250)
    if (mean) *mean = host[0];
    if (sd) *sd = host[1];
    return RCN_OK;
}

int rcn_cuda_standardise(rcn_cuda_handle h, double* feats, size_t n) {
    RCN_ENTER(h);
    if (n == 0) return RCN_OK;
    if (!feats) return fail(RCN_ERR_INVALID, "null feats");
    if (is_device_ptr(feats)) return launch_standardise(feats, n, h->mean, h->sd, h->stream);
    RCN_TRY(h->in_stage.reserve(n * sizeof(double)));
    RCN_CUDA_TRY(cudaMemcpyAsync(h->in_stage.p, feats, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    RCN_TRY(launch_standardise(h->in_stage.as<double>(), n, h->mean, h->sd, h->stream));
    return deliver(h, feats, h->in_stage.p, n * sizeof(double));
}

int rcn_cuda_forward(rcn_cuda_handle h, const double* feats, size_t B, double* out_acts) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (B == 0) return RCN_OK;
    if (!feats || !out_acts) return fail(RCN_ERR_INVALID, "null feats / output");
    StagedIn in;
    RCN_TRY(in.stage(feats, h->cols[0] * B * sizeof(double), h->in_stage, h->stream, nullptr));
    RCN_TRY(forward_dev(h, (const double*)in.dev, B, nullptr, nullptr, false));
    const size_t n = h->rows.size();
    return deliver(h, out_acts, h->act(n - 1, B), h->rows[n - 1] * B * sizeof(double));
}

static int classify_dev_feats(rcn_cuda_handle h, const double* feats_dev, size_t B, int64_t* labels_out) {
    RCN_TRY(forward_dev(h, feats_dev, B, nullptr, nullptr, false));
    const size_t n = h->rows.size();
    int64_t* lab_dev = labels_out;
    const bool out_dev = is_device_ptr(labels_out);
    if (!out_dev) { RCN_TRY(h->out_stage.reserve(B * sizeof(int64_t))); lab_dev = h->out_stage.as<int64_t>(); }
    RCN_TRY(launch_argmax_last(h->act(n - 1, B), h->rows[n - 1], B, lab_dev, h->stream));
    if (!out_dev) RCN_TRY(deliver(h, labels_out, lab_dev, B * sizeof(int64_t)));
    return RCN_OK;
}

int rcn_cuda_classify_features(rcn_cuda_handle h, const double* feats, size_t B, int64_t* labels_out) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (B == 0) return RCN_OK;
    if (!feats || !labels_out) return fail(RCN_ERR_INVALID, "null feats / output");
    StagedIn in;
    RCN_TRY(in.stage(feats, h->cols[0] * B * sizeof(double), h->in_stage, h->stream, nullptr));
    return classify_dev_feats(h, (const double*)in.dev, B, labels_out);
}

int rcn_cuda_classify(rcn_cuda_handle h, const void* images, int pixel_format, size_t B, size_t H, size_t W,
                      int64_t* labels_out) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (B == 0) return RCN_OK;
    if (!labels_out) return fail(RCN_ERR_INVALID, "null output");
    RCN_TRY(ensure_plan(h, H, W));
    RCN_TRY(check_feature_width(h, h->plan.L));
    RCN_TRY(h->feats.reserve(h->plan.L * B * sizeof(double)));
    RCN_TRY(features_into(h, images, pixel_format, B, H, W, true, h->feats.as<double>()));  // rcn.rs:84-89
    return classify_dev_feats(h, h->feats.as<double>(), B, labels_out);
}

int rcn_cuda_evaluate(rcn_cuda_handle h, const double* feats, const int64_t* labels, size_t B, uint64_t* accept) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (!accept) return fail(RCN_ERR_INVALID, "null accept");
    *accept = 0;
    if (B == 0) return RCN_OK;
    if (!feats || !labels) return fail(RCN_ERR_INVALID, "null feats / labels");
    StagedIn in, lb;
    RCN_TRY(in.stage(feats, h->cols[0] * B * sizeof(double), h->in_stage, h->stream, nullptr));
    RCN_TRY(lb.stage(labels, B * sizeof(int64_t), h->tgt_stage, h->stream, nullptr));
    RCN_TRY(forward_dev(h, (const double*)in.dev, B, nullptr, nullptr, false));
    const size_t n = h->rows.size();
    RCN_TRY(h->small.reserve(64));
    double* st = h->small.as<double>() + 2;
    RCN_TRY(launch_batch_stats(h->act(n - 1, B), h->rows[n - 1], B, nullptr, (const int64_t*)lb.dev, st, h->rs, h->stream));
    uint64_t host[2];
    RCN_CUDA_TRY(cudaMemcpyAsync(host, st, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
    RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));
    *accept = host[1];
    return RCN_OK;
}

int rcn_cuda_accumulate_gradients(rcn_cuda_handle h, const double* feats, const double* onehot, const int64_t* labels,
                                  size_t B) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    const double* oh = nullptr; const int64_t* lb = nullptr;
    const double* f_dev = nullptr;
    bool host_in = false;
    if (B) {
        if (!feats) return fail(RCN_ERR_INVALID, "null feats");
        RCN_TRY(stage_targets(h, onehot, labels, B, &oh, &lb));
        StagedIn in;
        RCN_TRY(in.stage(feats, h->cols[0] * B * sizeof(double), h->in_stage, h->stream, &host_in));
        f_dev = (const double*)in.dev;
    }
    RCN_TRY(accumulate_dev(h, f_dev, oh, lb, B));
    if (host_in || (onehot && !is_device_ptr(onehot)) || (labels && !is_device_ptr(labels)))
        RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));  // caller may reuse host buffers
    return RCN_OK;
}

int rcn_cuda_accumulate_gradients_images(rcn_cuda_handle h, const void* images, int pixel_format, const int64_t* labels,
                                         size_t B, size_t H, size_t W) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    RCN_TRY(ensure_plan(h, H, W));
    RCN_TRY(check_feature_width(h, h->plan.L));
    const double* oh = nullptr; const int64_t* lb = nullptr;
    if (B) {
        if (!labels) return fail(RCN_ERR_INVALID, "null labels");
        RCN_TRY(stage_targets(h, nullptr, labels, B, &oh, &lb));
        if (pixel_format != RCN_PIXELS_U8_ROWMAJOR && pixel_format != RCN_PIXELS_F64_COLMAJOR)
            return fail(RCN_ERR_INVALID, "unknown pixel format %d", pixel_format);
        if (!images) return fail(RCN_ERR_INVALID, "null images");
        StagedIn in;
        RCN_TRY(in.stage(images, B * H * W * pixel_bytes(pixel_format), h->in_stage, h->stream, nullptr));
        RCN_TRY(accumulate_images_dev(h, in.dev, pixel_format, lb, B, H, W, nullptr));
    } else {
        RCN_TRY(accumulate_dev(h, nullptr, nullptr, nullptr, 0));
    }
    if (B && (!is_device_ptr(images) || !is_device_ptr(labels))) RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return RCN_OK;
}

int rcn_cuda_apply_gradients(rcn_cuda_handle h, double eta, size_t batch) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (batch == 0) return RCN_OK;  // chunks_exact never yields an empty batch (rcn.rs:147)
    const double scale = eta / (double)batch;  // (eta / batch.len() as f64)  (rcn.rs:214)
    if (h->dp.connected)   // exchange over NVLink peer memory fused with the update (dp.cu)
        return launch_dp_allreduce_sgd(h->dp, h->params.as<double>(), h->grads, scale, h->stream, nullptr, 0, 0, nullptr, nullptr,
                                       take_dp_pushed(h), h->small_desc.tl);
    return launch_sgd_update(h->params.as<double>(), h->grads, h->n_params, scale, h->stream);
}

// Every entry point that drains the model's stream ends the hazard x_owns_cursor guards against.
static void stream_drained(rcn_cuda_model* h) { h->x_owns_cursor = false; }

// Single-GPU steps of the fused small-network path let the weight-gradient kernel apply the update (SnUpdate): arm it
// before the accumulate; if that accumulate took another path the standalone update kernel runs as before.
static void arm_fused_update(rcn_cuda_model* h, double scale, long long* cursor, long long batch, long long n_samples,
                             double* stats_ring, bool streamed = false) {
    static const bool on = []() { const char* e = getenv("RCN_CUDA_FUSED_UPDATE"); return !(e && e[0] == '0'); }();
    h->pending_upd = SnUpdate{};
    if (!on || !h || !h->params_ready || !h->use_small) return;
    // Data-parallel groups: receiving inside the weight-gradient kernel (MODE 3) was measured SLOWER than the separate
    // exchange kernel (2 GPUs, c2: 30.8 vs 28.4 us/step) -- the NVLink round trip is then exposed inside the kernel
    // instead of overlapping the next launch -- so it is opt-in (RCN_CUDA_DP_FUSED_UPDATE=1).  By default kernel B only
    // takes over the cursor / result ring there (params stays null) and the exchange kernel applies the update.
    static const bool dp_on = []() { const char* e = getenv("RCN_CUDA_DP_FUSED_UPDATE"); return e && e[0] == '1'; }();
    if (h->dp.connected && (!dp_fused_push_enabled() || h->dp_push_suppress)) return;
    if (!h->dp.connected || dp_on) h->pending_upd.params = h->params.as<double>();
    h->pending_upd.scale = scale;
    h->pending_upd.cursor = cursor;
    h->pending_upd.batch = batch;
    h->pending_upd.n_samples = n_samples;
    h->pending_upd.stats_ring = stats_ring;
    h->pending_upd.early_cursor = (!h->dp.connected && (prewait_mode() == 2 || (prewait_mode() == 1 && streamed))) ? 1 : 0;
}
// What the accumulate did with the armed update: bit 0 = parameters updated, bit 1 = cursor advanced / result ring written.
static int take_upd_fused(rcn_cuda_model* h) {
    const int f = (h->upd_fused ? 1 : 0) | (h->cursor_fused ? 2 : 0);
    h->upd_fused = false;
    h->cursor_fused = false;
    h->pending_upd = SnUpdate{};
    return f;
}

int rcn_cuda_train_batch(rcn_cuda_handle h, const double* feats, const double* onehot, const int64_t* labels, size_t B,
                         double eta) {
    // on a connected data-parallel group B is this rank's shard and the update uses the global batch (rcn.rs:214)
    const size_t global = h ? B * (size_t)(h->dp.connected ? h->dp.world : 1) : B;
    if (h && B) arm_fused_update(h, eta / (double)global, nullptr, 0, 0, nullptr);
    const int rc = rcn_cuda_accumulate_gradients(h, feats, onehot, labels, B);
    if (rc != RCN_OK) { if (h) take_upd_fused(h); return rc; }
    if (take_upd_fused(h) & 1) return RCN_OK;
    return rcn_cuda_apply_gradients(h, eta, global);
}

int rcn_cuda_train_batch_images(rcn_cuda_handle h, const void* images, int pixel_format, const int64_t* labels, size_t B,
                                size_t H, size_t W, double eta) {
    const size_t global = h ? B * (size_t)(h->dp.connected ? h->dp.world : 1) : B;
    if (h && B) arm_fused_update(h, eta / (double)global, nullptr, 0, 0, nullptr);
    const int rc = rcn_cuda_accumulate_gradients_images(h, images, pixel_format, labels, B, H, W);
    if (rc != RCN_OK) { if (h) take_upd_fused(h); return rc; }
    if (take_upd_fused(h) & 1) return RCN_OK;
    return rcn_cuda_apply_gradients(h, eta, global);
}

int rcn_cuda_last_batch_stats(rcn_cuda_handle h, double* cost, uint64_t* hits) {
    RCN_ENTER(h);
    if (!h->stats_valid) return fail(RCN_ERR_STATE, "no batch has been accumulated yet");
    double host[2];
    RCN_CUDA_TRY(cudaMemcpyAsync(host, h->small.p, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
    RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (cost) *cost = host[0];
    if (hits) memcpy(hits, &host[1], sizeof(uint64_t));
    return RCN_OK;
}

// ---- epoch mode: rcn.rs:144-149 over a dataset resident in device memory ---------------------------------------------
int rcn_cuda_epoch_bind(rcn_cuda_handle h, const void* images, int pixel_format, const int64_t* labels,
                        const int64_t* perm, size_t n_samples, size_t H, size_t W, size_t B) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (!images || !labels || B == 0 || n_samples < B) return fail(RCN_ERR_INVALID, "epoch_bind needs images, labels and n_samples >= B > 0");
    if (!is_device_ptr(images) || !is_device_ptr(labels) || (perm && !is_device_ptr(perm)))
        return fail(RCN_ERR_INVALID, "epoch mode needs a dataset resident in device memory");
    if (pixel_format != RCN_PIXELS_U8_ROWMAJOR && pixel_format != RCN_PIXELS_F64_COLMAJOR)
        return fail(RCN_ERR_INVALID, "unknown pixel format %d", pixel_format);
    RCN_TRY(ensure_plan(h, H, W));
    RCN_TRY(check_feature_width(h, h->plan.L));
    RCN_TRY(h->ep_state.reserve((kEpSlots + B) * sizeof(int64_t)));
    RCN_CUDA_TRY(cudaMemsetAsync(h->ep_state.p, 0, (kEpSlots + B) * sizeof(int64_t), h->stream));
    RCN_TRY(h->feats.reserve(h->plan.L * B * sizeof(double)));
    h->ep_images = images; h->ep_labels = labels; h->ep_perm = perm; h->ep_fmt = pixel_format;
    h->ep_n = n_samples; h->ep_H = H; h->ep_W = W; h->ep_B = B;
    RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));   // the cursor is zero before any step kernel can look at it
    stream_drained(h);
    return RCN_OK;
}

int rcn_cuda_epoch_seek(rcn_cuda_handle h, size_t position) {
    RCN_ENTER(h);
    if (!h->ep_images) return fail(RCN_ERR_STATE, "no dataset bound: call rcn_cuda_epoch_bind first");
    if (position + h->ep_B > h->ep_n) return fail(RCN_ERR_INVALID, "position %zu leaves fewer than B samples", position);
    const long long pos = (long long)position;
    RCN_CUDA_TRY(cudaMemcpyAsync(h->ep_state.p, &pos, sizeof(pos), cudaMemcpyHostToDevice, h->stream));
    RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));
    stream_drained(h);
    return RCN_OK;
}

int rcn_cuda_epoch_position(rcn_cuda_handle h, size_t* position) {
    RCN_ENTER(h);
    if (!h->ep_images || !position) return fail(RCN_ERR_STATE, "no dataset bound / null output");
    long long pos = 0;
    RCN_CUDA_TRY(cudaMemcpyAsync(&pos, h->ep_state.p, sizeof(pos), cudaMemcpyDeviceToHost, h->stream));
    RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));
    stream_drained(h);
    *position = (size_t)pos;
    return RCN_OK;
}

int rcn_cuda_epoch_accumulate(rcn_cuda_handle h) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (!h->ep_images) return fail(RCN_ERR_STATE, "no dataset bound: call rcn_cuda_epoch_bind first");
    BatchIndex bi;
    bi.cursor = h->ep_state.as<long long>();
    bi.perm = (const long long*)h->ep_perm;
    bi.labels_all = (const long long*)h->ep_labels;
    bi.labels_batch = h->ep_state.as<long long>() + kEpSlots;
    static const bool l2_prefetch = []() { const char* e = getenv("RCN_CUDA_L2_PREFETCH"); return !(e && e[0] == '0'); }();
    if (l2_prefetch) {   // lets kernel A ask L2 for the next step's images (smallnet.cu)
        bi.batch = (long long)h->ep_B;
        bi.n_samples = (long long)h->ep_n;
    }
    return accumulate_images_dev(h, h->ep_images, h->ep_fmt, nullptr, h->ep_B, h->ep_H, h->ep_W, &bi);
}

static int epoch_apply_impl(rcn_cuda_model* h, double eta, size_t global_batch, bool cursor_done) {
    if (global_batch == 0) return fail(RCN_ERR_INVALID, "global batch is zero");
    const double scale = eta / (double)global_batch;
    long long* cursor = cursor_done ? nullptr : h->ep_state.as<long long>();   // kernel B already advanced it
    if (h->dp.connected) {
        if (cursor) h->x_owns_cursor = true;
        return launch_dp_allreduce_sgd(h->dp, h->params.as<double>(), h->grads, scale, h->stream, cursor, (long long)h->ep_B,
                                       (long long)h->ep_n, nullptr, nullptr, take_dp_pushed(h), h->small_desc.tl);
    }
    return launch_sgd_update(h->params.as<double>(), h->grads, h->n_params, scale, h->stream, cursor, (long long)h->ep_B,
                             (long long)h->ep_n, nullptr, nullptr);
}

int rcn_cuda_epoch_apply(rcn_cuda_handle h, double eta, size_t global_batch) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (!h->ep_images) return fail(RCN_ERR_STATE, "no dataset bound: call rcn_cuda_epoch_bind first");
    return epoch_apply_impl(h, eta, global_batch, false);
}

int rcn_cuda_epoch_step(rcn_cuda_handle h, double eta) {
    const size_t global = h ? h->ep_B * (size_t)(h->dp.connected ? h->dp.world : 1) : 0;
    if (h && h->ep_images && h->ep_B)
        arm_fused_update(h, eta / (double)global, h->ep_state.as<long long>(), (long long)h->ep_B, (long long)h->ep_n, nullptr);
    const int rc = rcn_cuda_epoch_accumulate(h);
    if (rc != RCN_OK) { if (h) take_upd_fused(h); return rc; }
    const int fused = take_upd_fused(h);
    if (fused & 1) return RCN_OK;
    RCN_ENTER(h);
    return epoch_apply_impl(h, eta, global, (fused & 2) != 0);
}

int rcn_cuda_epoch_run(rcn_cuda_handle h, double eta, size_t n_steps) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (!h->ep_images) return fail(RCN_ERR_STATE, "no dataset bound: call rcn_cuda_epoch_bind first");
    if (n_steps > (size_t)1 << 30) return fail(RCN_ERR_INVALID, "too many steps in one call");
    for (size_t k = 0; k < n_steps; ++k) RCN_TRY(rcn_cuda_epoch_step(h, eta));
    return RCN_OK;
}

// ---- pipelined loop over a HOST-resident dataset: rcn.rs:147-149 ------------------------------------------------------
int rcn_cuda_train_epoch_host(rcn_cuda_handle h, const void* images, int pixel_format, const int64_t* labels,
                              size_t n_samples, size_t H, size_t W, size_t B, double eta, size_t global_batch,
                              double* cost_out, uint64_t* hits_out, size_t* n_steps_out) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (n_steps_out) *n_steps_out = 0;
    if (B == 0) return fail(RCN_ERR_INVALID, "batch size is zero");
    if (pixel_format != RCN_PIXELS_U8_ROWMAJOR && pixel_format != RCN_PIXELS_F64_COLMAJOR)
        return fail(RCN_ERR_INVALID, "unknown pixel format %d", pixel_format);
    const size_t n_steps = n_samples / B;       // chunks_exact: the remainder is dropped (rcn.rs:147)
    if (n_steps == 0) return RCN_OK;
    if (!images || !labels) return fail(RCN_ERR_INVALID, "null images / labels");
    if (is_device_ptr(images) || is_device_ptr(labels))
        return fail(RCN_ERR_INVALID, "train_epoch_host takes HOST buffers; use rcn_cuda_epoch_bind for a device-resident dataset");
    RCN_TRY(ensure_plan(h, H, W));
    RCN_TRY(check_feature_width(h, h->plan.L));
    if (global_batch == 0) global_batch = B * (size_t)(h->dp.connected ? h->dp.world : 1);
    const size_t img_bytes = B * H * W * pixel_bytes(pixel_format);
    if (!h->copy_stream) RCN_CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->hs_fork) {
        for (int i = 0; i < 2; ++i) {
            RCN_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
            RCN_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_consumed[i], cudaEventDisableTiming));
        }
        RCN_CUDA_TRY(cudaEventCreateWithFlags(&h->hs_fork, cudaEventDisableTiming));
        RCN_CUDA_TRY(cudaEventCreateWithFlags(&h->hs_join, cudaEventDisableTiming));
    }
    if (h->stats_host_cap < n_steps) {
        if (h->stats_host) cudaFreeHost(h->stats_host);
        h->stats_host = nullptr; h->stats_host_cap = 0;
        RCN_CUDA_TRY(cudaHostAlloc((void**)&h->stats_host, n_steps * 2 * sizeof(double), cudaHostAllocDefault));
        h->stats_host_cap = n_steps;
    }
    // ---- streaming fast path: pinned u8 dataset + fused small-network kernels: the GPU pulls chunk k+1 over PCIe itself
    // (zero-copy loads) while chunk k trains; per-step results are written straight into pinned host memory; the host
    // launches ONE CUDA graph per step and nothing else.
    static const bool stream_env = []() { const char* e = getenv("RCN_CUDA_HOST_STREAMING"); return !(e && e[0] == '0'); }();
    if (stream_env && pixel_format == RCN_PIXELS_U8_ROWMAJOR && h->use_small && B <= smallnet_max_batch() && img_bytes % 16 == 0 &&
        (reinterpret_cast<uintptr_t>(images) & 15) == 0 && is_pinned_host_ptr(images) && h->plan.L > 0 && h->plan.n_conv <= 10) {
        SmallNetFront probe{};
        probe.H = (int)H; probe.W = (int)W; probe.max_elems = (int)h->plan.max_elems; probe.stages = h->plan.stages;
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        RCN_CUDA_TRY(cudaStreamIsCapturing(h->stream, &cs));
        if (smallnet_front_fits(h->small_desc, probe) && cs == cudaStreamCaptureStatusNone) {
            // The legacy default stream cannot be captured: run the loop on the model's own stream, ordered after whatever
            // the caller has enqueued on the legacy stream; the call drains that stream before it returns.
            struct StreamSwap {
                rcn_cuda_model* m; cudaStream_t saved;
                StreamSwap(rcn_cuda_model* m_) : m(m_), saved(m_->stream) {}
                ~StreamSwap() { m->stream = saved; }
            } swap_guard(h);
            if (h->stream == nullptr) {
                RCN_CUDA_TRY(cudaEventRecord(h->hs_fork, nullptr));
                RCN_CUDA_TRY(cudaStreamWaitEvent(h->own_stream, h->hs_fork, 0));
                h->stream = h->own_stream;
            }
            RCN_TRY(h->hs_ring.reserve(kHsRingSteps * img_bytes));
            // copy-engine mode needs the staged front end (its image-loading lanes do the waiting) and stream memory operations
            bool dma = hs_dma_env() && fused_front_enabled() && kHsRingSteps * B < (1ull << 31) && stream_mem_ops(h->device).ok;
            if (dma) {
                probe.images = h->hs_ring.as<uint8_t>();
                smallnet_front_select(h->plan, &probe);
                dma = probe.use_cp != 0;
            }
            const size_t spg = (size_t)hs_steps_per_graph(dma);
            const size_t ring_steps = dma ? kHsRingSteps : 2;
            // The epoch as a sequence of graph launches: as many `spg`-step graphs as fit, then (dma) the largest of 20 / 10 / 5 / 2 /
            // 1 steps that still fits -- as few launches as possible: with copies on the link a graph boundary costs 4-5 us and
            // now and then 27 us (a short FIRST graph to get going sooner measured worse for that reason: 507-519 us against 479
            // for a 20-step epoch; its launch latency passes while the first chunk crosses the link anyway).
            std::vector<int> plan_g;
            for (size_t left = n_steps; left;) {
                size_t g = 1;
                if (left >= spg) g = spg;
                else if (dma)
                    for (size_t c : {(size_t)20, (size_t)10, (size_t)5, (size_t)2})
                        if (c < spg && left >= c) { g = c; break; }
                plan_g.push_back((int)g);
                left -= g;
            }
            RCN_TRY(h->hs_state.reserve((kHsStateSlots + B) * sizeof(long long)));
            RCN_TRY(h->tgt_stage.reserve(n_steps * B * sizeof(int64_t)));
            RCN_TRY(h->feats.reserve(h->plan.L * B * sizeof(double)));
            const double scale_s = eta / (double)global_batch;
            long long* st = h->hs_state.as<long long>();
            BatchIndex bi;
            bi.cursor = st;
            bi.labels_all = h->tgt_stage.as<long long>();
            bi.labels_batch = st + kHsStateSlots;
            bi.window = (long long)(ring_steps * B);
            bi.arrived = dma ? st + 3 : nullptr;
            // state, labels of the whole epoch (and, pull mode, chunk 0) go up with plain copies; after that graph replays
            const long long init[kHsStateSlots] = {0, (long long)n_steps, (long long)reinterpret_cast<uintptr_t>(images), 0, 0, 0};
            RCN_CUDA_TRY(cudaMemcpyAsync(st, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
            RCN_CUDA_TRY(cudaMemcpyAsync(h->tgt_stage.p, labels, n_steps * B * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
            if (!dma) RCN_CUDA_TRY(cudaMemcpyAsync(h->hs_ring.p, images, img_bytes, cudaMemcpyHostToDevice, h->stream));
            rcn_cuda_model::HsKey key;
            key.B = B; key.H = H; key.W = W; key.n_steps = 0; key.scale = scale_s;   // the graphs do not depend on the epoch length
            key.labels = h->tgt_stage.p;   // (grows with the epoch length: a longer epoch than any before re-captures)
            key.alloc_gen = alloc_generation().load(std::memory_order_relaxed);
            memcpy(&key.mean_bits, &h->mean, sizeof(double));   // a set_scale / gen_scales since the capture re-captures
            memcpy(&key.sd_bits, &h->sd, sizeof(double));
            key.stats = h->stats_host; key.ring = h->hs_ring.p; key.state = st; key.grads = h->grads; key.stream = h->stream;
            key.dp = h->dp.connected;
            key.window = dma ? -bi.window : bi.window;   // (sign: the pull graphs carry the prefetch branch, the dma graphs do not)
            if (h->hs_graphs.empty() || !(key == h->hs_key)) {
                for (auto& kv : h->hs_graphs) cudaGraphExecDestroy(kv.second);
                h->hs_graphs.clear();
                // warm-up outside capture: reserves every scratch buffer and sets kernel attributes (no parameter update)
                h->dp_push_suppress = true;
                BatchIndex bi_warm = bi;
                bi_warm.arrived = nullptr;   // nothing has been copied yet: the warm-up trains on whatever the ring holds
                const int wrc = accumulate_images_dev(h, h->hs_ring.p, pixel_format, nullptr, B, H, W, &bi_warm);
                h->dp_push_suppress = false;
                RCN_TRY(wrc);
                RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));
                key.alloc_gen = alloc_generation().load(std::memory_order_relaxed);   // the warm-up may have grown scratch buffers
                h->hs_key = key;
            }
            {
                // `g` consecutive steps in ONE graph: a step is only tens of microseconds long, so the host's per-launch cost
                // is spread over several of them; the remainder of the epoch replays the one-step graph.
                auto capture_steps = [&](int g, cudaGraphExec_t* out) -> int {
                    cudaGraph_t graph = nullptr;
                    RCN_CUDA_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
                    int rc = RCN_OK;
                    for (int q = 0; q < g && rc == RCN_OK; ++q) {
                        if (!dma) {
                        if (cudaEventRecord(h->hs_fork, h->stream) != cudaSuccess || cudaStreamWaitEvent(h->copy_stream, h->hs_fork, 0) != cudaSuccess) { rc = fail(RCN_ERR_CUDA, "graph fork failed"); break; }
                        {
                            LaunchScope ls("host_prefetch_kernel", h->copy_stream);
                            launch_host_prefetch(st, h->hs_ring.as<unsigned char>(), (long long)B, (long long)(H * W), h->copy_stream);
                        }
                        if (cudaPeekAtLastError() != cudaSuccess) { rc = fail(RCN_ERR_CUDA, "prefetch kernel launch failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
                        if (cudaEventRecord(h->hs_join, h->copy_stream) != cudaSuccess) { rc = fail(RCN_ERR_CUDA, "graph join failed"); break; }
                        }
                        arm_fused_update(h, scale_s, st, (long long)B, kHsNoWrap, h->stats_host, true);
                        rc = accumulate_images_dev(h, h->hs_ring.p, pixel_format, nullptr, B, H, W, &bi);
                        const int fused = take_upd_fused(h);
                        if (rc != RCN_OK) break;
                        if (fused & 1) {   // the weight-gradient kernel applied the update, advanced the cursor, wrote the result
                            // the prefetch branch joins at the end of its step: the next step starts after both
                            if (!dma && cudaStreamWaitEvent(h->stream, h->hs_join, 0) != cudaSuccess) { rc = fail(RCN_ERR_CUDA, "graph join failed"); break; }
                            continue;
                        }
                        // (fused & 2: kernel B advanced the cursor and wrote the result; the update kernel gets no cursor)
                        long long* cur = (fused & 2) ? nullptr : st;
                        if (h->dp.connected) {
                            if (cur) h->x_owns_cursor = true;
                            rc = launch_dp_allreduce_sgd(h->dp, h->params.as<double>(), h->grads, scale_s, h->stream, cur, (long long)B,
                                                         kHsNoWrap, h->small.as<double>(), h->stats_host, take_dp_pushed(h), h->small_desc.tl);
                        } else {
                            rc = launch_sgd_update(h->params.as<double>(), h->grads, h->n_params, scale_s, h->stream, cur, (long long)B,
                                                   kHsNoWrap, h->small.as<double>(), h->stats_host);
                        }
                        if (rc != RCN_OK) break;
                        // the prefetch branch joins AFTER the exchange / update kernel was enqueued: the next step's kernel A
                        // then directly follows that kernel in the stream (its programmatic dependent) and also waits for
                        // the prefetched chunk
                        if (!dma && cudaStreamWaitEvent(h->stream, h->hs_join, 0) != cudaSuccess) { rc = fail(RCN_ERR_CUDA, "graph join failed"); break; }
                    }
                    cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
                    if (rc != RCN_OK) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
                    if (ce != cudaSuccess || !graph) return fail(RCN_ERR_CUDA, "stream capture of the training step failed: %s", cudaGetErrorString(ce));
                    ce = cudaGraphInstantiate(out, graph, 0);
                    cudaGraphDestroy(graph);
                    if (ce != cudaSuccess) { *out = nullptr; return fail(RCN_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce)); }
                    return RCN_OK;
                };
                for (int g : plan_g) {
                    if (h->hs_graphs.count(g)) continue;
                    cudaGraphExec_t ge = nullptr;
                    RCN_TRY(capture_steps(g, &ge));
                    h->hs_graphs[g] = ge;
                }
            }
            if (!dma) {
                for (int g : plan_g) RCN_CUDA_TRY(cudaGraphLaunch(h->hs_graphs[g], h->stream));
            } else {
                // Copies first (the driver queues them; each waits on the device for its ring slots), then the graph launches.
                const StreamMemOps& mo = stream_mem_ops(h->device);
                const CUdeviceptr cursor_dev = (CUdeviceptr)reinterpret_cast<uintptr_t>(st);
                const char* src = static_cast<const char*>(images);
                if (h->hs_counts_cap < n_steps) {   // pinned table of the counter values, one per copy
                    if (h->hs_counts) cudaFreeHost(h->hs_counts);
                    h->hs_counts = nullptr; h->hs_counts_cap = 0;
                    RCN_CUDA_TRY(cudaHostAlloc((void**)&h->hs_counts, n_steps * sizeof(long long), cudaHostAllocDefault));
                    h->hs_counts_cap = n_steps;
                }
                if (!h->flag_stream) {
                    RCN_CUDA_TRY(cudaStreamCreateWithFlags(&h->flag_stream, cudaStreamNonBlocking));
                    for (int i = 0; i < kHsEvents; ++i) RCN_CUDA_TRY(cudaEventCreateWithFlags(&h->hs_ev[i], cudaEventDisableTiming));
                }
                // the state block (cursor 0, arrived 0) must be in place before the first counter copy can land
                RCN_CUDA_TRY(cudaEventRecord(h->hs_fork, h->stream));
                RCN_CUDA_TRY(cudaStreamWaitEvent(h->flag_stream, h->hs_fork, 0));
                size_t a = 0, k = 0, n_copy = 0;
                auto enqueue_copy = [&]() -> int {
                    // ramp: while the copies are only a few chunks ahead of the training steps a long copy would hold back the
                    // steps waiting for its first chunk (the counter moves when a copy has landed completely)
                    size_t g = a < 2 ? 1 : (a < 16 ? 2 : (a < 32 ? 4 : kHsMaxCopy));
                    g = std::min(g, n_steps - a);
                    g = std::min(g, kHsQuarter - a % kHsQuarter);                      // never across a quarter of the ring
                    if (a >= ring_steps && a % kHsQuarter == 0) {
                        // entering a quarter whose slots held steps [a - ring, a + quarter - ring): all of them must have trained
                        if (mo.wait64(h->copy_stream, cursor_dev, (cuuint64_t)((a + kHsQuarter - ring_steps) * B), CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                            return fail(RCN_ERR_CUDA, "cuStreamWaitValue64 failed");
                    }
                    RCN_CUDA_TRY(cudaMemcpyAsync(h->hs_ring.as<char>() + (a % ring_steps) * img_bytes, src + a * img_bytes, g * img_bytes,
                                                 cudaMemcpyHostToDevice, h->copy_stream));
                    a += g;
                    h->hs_counts[n_copy] = (long long)(a * B);
                    cudaEvent_t ev = h->hs_ev[n_copy % kHsEvents];
                    RCN_CUDA_TRY(cudaEventRecord(ev, h->copy_stream));
                    RCN_CUDA_TRY(cudaStreamWaitEvent(h->flag_stream, ev, 0));
                    // (cuStreamWriteValue64 here instead measured the same: 476 vs 478 us for a 20-step epoch)
                    RCN_CUDA_TRY(cudaMemcpyAsync(st + 3, h->hs_counts + n_copy, sizeof(long long), cudaMemcpyHostToDevice, h->flag_stream));
                    ++n_copy;
                    return RCN_OK;
                };
                // One chunk's copy, then the first graph launch (its launch latency passes while the chunk crosses the link; kernel
                // A waits on the device for whatever has not arrived yet), then copies and launches interleaved, the copies
                // enqueued half a ring ahead of the launches.
                RCN_TRY(enqueue_copy());
                for (int g : plan_g) {
                    RCN_CUDA_TRY(cudaGraphLaunch(h->hs_graphs[g], h->stream));
                    k += (size_t)g;
                    while (a < n_steps && a < k + ring_steps / 2) RCN_TRY(enqueue_copy());
                }
            }
            // (pull: prefetch +) A + B (+ exchange/update) per replayed step
            g_launches.fetch_add((unsigned long long)n_steps * ((dma ? 2 : 3) + (h->dp.connected ? 1 : 0)), std::memory_order_relaxed);
            RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));
            stream_drained(h);
            for (size_t k = 0; k < n_steps; ++k) {
                if (cost_out) cost_out[k] = h->stats_host[2 * k];
                if (hits_out) memcpy(&hits_out[k], &h->stats_host[2 * k + 1], sizeof(uint64_t));
            }
            h->stats_valid = true;
            h->last_B = B;
            if (n_steps_out) *n_steps_out = n_steps;
            return RCN_OK;
        }
    }
    // ---- general path: double-buffered cudaMemcpyAsync on the copy stream -----------------------------------------------
    for (int i = 0; i < 2; ++i) RCN_TRY(h->host_slot[i].reserve(img_bytes));
    // the labels of the whole epoch go up in ONE copy (one driver call per step less); images stream chunk by chunk
    RCN_TRY(h->tgt_stage.reserve(n_steps * B * sizeof(int64_t)));
    RCN_CUDA_TRY(cudaMemcpyAsync(h->tgt_stage.p, labels, n_steps * B * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    const int64_t* labels_dev = h->tgt_stage.as<int64_t>();
    // everything the previous user of the staging slots enqueued on the compute stream must be finished first
    RCN_CUDA_TRY(cudaEventRecord(h->ev_consumed[0], h->stream));
    RCN_CUDA_TRY(cudaEventRecord(h->ev_consumed[1], h->stream));
    const double scale = eta / (double)global_batch;   // (eta / batch.len() as f64)  (rcn.rs:214)
    const char* img = (const char*)images;
    for (size_t k = 0; k < n_steps; ++k) {
        const int slot = (int)(k & 1);
        char* dst = h->host_slot[slot].as<char>();
        // H2D of chunk k on the copy stream: overlaps the kernels of chunk k-1
        RCN_CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_consumed[slot], 0));
        RCN_CUDA_TRY(cudaMemcpyAsync(dst, img + k * img_bytes, img_bytes, cudaMemcpyHostToDevice, h->copy_stream));
        RCN_CUDA_TRY(cudaEventRecord(h->ev_copied[slot], h->copy_stream));
        RCN_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_copied[slot], 0));
        RCN_TRY(accumulate_images_dev(h, dst, pixel_format, labels_dev + k * B, B, H, W, nullptr));
        RCN_CUDA_TRY(cudaEventRecord(h->ev_consumed[slot], h->stream));
        if (h->dp.connected)
            RCN_TRY(launch_dp_allreduce_sgd(h->dp, h->params.as<double>(), h->grads, scale, h->stream, nullptr, 0, 0, nullptr, nullptr,
                                            take_dp_pushed(h), h->small_desc.tl));
        else
            RCN_TRY(launch_sgd_update(h->params.as<double>(), h->grads, h->n_params, scale, h->stream));
        // D2H of this step's result (cost, hits evaluated with the pre-update parameters)
        RCN_CUDA_TRY(cudaMemcpyAsync(h->stats_host + 2 * k, h->small.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));
    RCN_CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
    for (size_t k = 0; k < n_steps; ++k) {
        if (cost_out) cost_out[k] = h->stats_host[2 * k];
        if (hits_out) memcpy(&hits_out[k], &h->stats_host[2 * k + 1], sizeof(uint64_t));
    }
    if (n_steps_out) *n_steps_out = n_steps;
    return RCN_OK;
}

// ---- data-parallel group (dp.cu) -------------------------------------------------------------------------------------
int rcn_cuda_dp_init(rcn_cuda_handle h, int world, int rank, void* ipc_handle_out) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    RCN_TRY(dp_alloc(h->dp, world, rank, h->n_params, h->stream));
    if (ipc_handle_out) {
        cudaIpcMemHandle_t hd;
        RCN_CUDA_TRY(cudaIpcGetMemHandle(&hd, h->dp.block));
        static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(ipc_handle_out, &hd, sizeof(hd));
    }
    return RCN_OK;
}

int rcn_cuda_dp_connect_ipc(rcn_cuda_handle h, const void* all_handles) {
    RCN_ENTER(h);
    if (!h->dp.block) return fail(RCN_ERR_STATE, "call rcn_cuda_dp_init first");
    if (!all_handles) return fail(RCN_ERR_INVALID, "null handles");
    for (int q = 0; q < h->dp.world; ++q) {
        if (q == h->dp.rank) continue;
        cudaIpcMemHandle_t hd;
        memcpy(&hd, (const char*)all_handles + (size_t)q * sizeof(hd), sizeof(hd));
        void* p = nullptr;
        RCN_CUDA_TRY(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
        h->dp.peers[q] = p; h->dp.imported[q] = true;
    }
    h->dp.connected = h->dp.world > 1;
    return RCN_OK;
}

int rcn_cuda_dp_connect_local(rcn_cuda_handle h, const rcn_cuda_handle* group) {
    RCN_ENTER(h);
    if (!h->dp.block) return fail(RCN_ERR_STATE, "call rcn_cuda_dp_init first");
    if (!group) return fail(RCN_ERR_INVALID, "null group");
    for (int q = 0; q < h->dp.world; ++q) {
        if (q == h->dp.rank) continue;
        rcn_cuda_model* o = group[q];
        if (!o || !o->dp.block || o->dp.world != h->dp.world || o->dp.rank != q || o->dp.n != h->dp.n)
            return fail(RCN_ERR_INVALID, "group member %d is not an initialised rank %d of a %d-rank group with %zu parameters", q, q,
                        h->dp.world, h->dp.n);
        if (o->device != h->device) {
            int can = 0;
            RCN_CUDA_TRY(cudaDeviceCanAccessPeer(&can, h->device, o->device));
            if (!can) return fail(RCN_ERR_CUDA, "device %d cannot access device %d as a peer", h->device, o->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(RCN_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", o->device, cudaGetErrorString(e));
            cudaGetLastError();
        }
        h->dp.peers[q] = o->dp.block; h->dp.imported[q] = false;
    }
    h->dp.connected = h->dp.world > 1;
    return RCN_OK;
}

int rcn_cuda_dp_shutdown(rcn_cuda_handle h) {
    RCN_ENTER(h);
    RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));
    stream_drained(h);
    dp_release(h->dp);
    h->dp_pushed = false;
    return RCN_OK;
}

static int timeline_reset(rcn_cuda_model* h) {
    static Timeline init;   // 4.3 KB: too large for the stack of a small thread
    memset(&init, 0, sizeof(init));
    memset(init.t0, 0xff, sizeof(init.t0));
    RCN_CUDA_TRY(cudaMemcpyAsync(h->timeline.p, &init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return RCN_OK;
}

int rcn_cuda_timeline_enable(rcn_cuda_handle h, int on) {
    RCN_ENTER(h);
    RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));
    stream_drained(h);
    if (on) {
        RCN_TRY(h->timeline.reserve(sizeof(Timeline)));
        RCN_TRY(timeline_reset(h));
    }
    if ((on != 0) != h->tl_on) alloc_generation().fetch_add(1, std::memory_order_relaxed);   // captured steps hold the pointer by value
    h->tl_on = on != 0;
    h->small_desc.tl = h->tl_on ? h->timeline.as<Timeline>() : nullptr;
    return RCN_OK;
}

int rcn_cuda_timeline_read(rcn_cuda_handle h, uint64_t* stamps, uint32_t* launches) {
    RCN_ENTER(h);
    if (!h->tl_on) return fail(RCN_ERR_STATE, "the timeline is not enabled");
    if (!stamps || !launches) return fail(RCN_ERR_INVALID, "null output");
    static Timeline host;
    RCN_CUDA_TRY(cudaMemcpyAsync(&host, h->timeline.p, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
    RCN_CUDA_TRY(cudaStreamSynchronize(h->stream));
    stream_drained(h);
    memcpy(stamps, host.t0, sizeof(host.t0));
    memcpy(stamps + kTlKernels * kTlRing, host.t1, sizeof(host.t1));
    for (int k = 0; k < kTlKernels; ++k) launches[k] = host.seq[k];
    return RCN_OK;
}

int rcn_cuda_dp_error(rcn_cuda_handle h, int* error) {
    RCN_ENTER(h);
    if (!error) return fail(RCN_ERR_INVALID, "null output");
    unsigned e = 0;
    RCN_TRY(dp_read_error(h->dp, h->stream, &e));
    stream_drained(h);
    *error = (int)e;
    if (e) return fail(RCN_ERR_STATE, "data-parallel group: a gradient exchange timed out waiting for a peer (a rank died, raised before "
                                      "its launch, or ran a different number of steps); the parameters of this replica are NaN");
    return RCN_OK;
}

int rcn_cuda_bind_gradient_buffer(rcn_cuda_handle h, double* device_ptr, size_t n) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (!device_ptr) {
        RCN_TRY(h->grads_own.reserve(h->n_params * sizeof(double)));
        h->grads = h->grads_own.as<double>();
        h->grads_bound = false;
        return RCN_OK;
    }
    if (n != h->n_params) return fail(RCN_ERR_SHAPE, "gradient buffer must hold %zu doubles, got %zu", h->n_params, n);
    if (!is_device_ptr(device_ptr)) return fail(RCN_ERR_INVALID, "gradient buffer must be device memory");
    h->grads = device_ptr;
    h->grads_bound = true;
    return RCN_OK;
}

int rcn_cuda_gradient_buffer(rcn_cuda_handle h, double** device_ptr, size_t* n) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (device_ptr) *device_ptr = h->grads;
    if (n) *n = h->n_params;
    return RCN_OK;
}

int rcn_cuda_get_gradients(rcn_cuda_handle h, double* flat, size_t n) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (n != h->n_params || !flat) return fail(RCN_ERR_SHAPE, "model has %zu parameters, got %zu", h->n_params, n);
    return deliver(h, flat, h->grads, n * sizeof(double));
}

int rcn_cuda_get_activations(rcn_cuda_handle h, size_t layer, double* out) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (layer >= h->rows.size() || !out) return fail(RCN_ERR_INVALID, "bad layer / null destination");
    if (!h->last_B) return fail(RCN_ERR_STATE, "no batch has been accumulated yet");
    return deliver(h, out, h->act(layer, h->last_B), h->rows[layer] * h->last_B * sizeof(double));
}

int rcn_cuda_get_deltas(rcn_cuda_handle h, size_t layer, double* out) {
    RCN_ENTER(h);
    RCN_TRY(require_params(h));
    if (layer >= h->rows.size() || !out) return fail(RCN_ERR_INVALID, "bad layer / null destination");
    if (!h->last_B) return fail(RCN_ERR_STATE, "no batch has been accumulated yet");
    return deliver(h, out, h->delta(layer, h->last_B), h->rows[layer] * h->last_B * sizeof(double));
}

// ---- launch accounting ------------------------------------------------------------------------------
int rcn_cuda_kernel_launches(uint64_t* count) {
    if (!count) return fail(RCN_ERR_INVALID, "null count");
    *count = g_launches.load();
    return RCN_OK;
}

int rcn_cuda_allocation_generation(uint64_t* generation) {
    if (!generation) return fail(RCN_ERR_INVALID, "null generation");
    *generation = alloc_generation().load(std::memory_order_relaxed);
    return RCN_OK;
}

int rcn_cuda_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (on) {
        for (auto& r : g_prof) { cudaEventDestroy(r.start); cudaEventDestroy(r.stop); }
        g_prof.clear();
    }
    g_profiling.store(on != 0);
    return RCN_OK;
}

int rcn_cuda_profile_report(char* json_out, size_t capacity) {
    if (!json_out || capacity < 3) return fail(RCN_ERR_INVALID, "report buffer too small");
    RCN_CUDA_TRY(cudaDeviceSynchronize());
    std::map<std::string, std::pair<unsigned long long, double>> agg;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        for (auto& r : g_prof) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, r.start, r.stop) != cudaSuccess) { cudaGetLastError(); continue; }
            auto& a = agg[r.name];
            a.first += 1; a.second += ms;
        }
    }
    std::string js = "{";
    bool first = true;
    for (auto& kv : agg) {
        char buf[256];
        snprintf(buf, sizeof(buf), "%s\"%s\": {\"launches\": %llu, \"total_ms\": %.6f}", first ? "" : ", ", kv.first.c_str(),
                 kv.second.first, kv.second.second);
        js += buf;
        first = false;
    }
    js += "}";
    if (js.size() + 1 > capacity) return fail(RCN_ERR_INVALID, "report needs %zu bytes", js.size() + 1);
    memcpy(json_out, js.c_str(), js.size() + 1);
    return RCN_OK;
}

// ---- op-level API ---------------------------------------------------------------------------------

int rcn_cuda_convolve_2d(int device, void* cuda_stream, const double* m, size_t H, size_t W, const double* kernel,
                         size_t kh, size_t kw, int padding, double* out) {
    if (!m || !kernel || !out) return fail(RCN_ERR_INVALID, "null pointer");
    if (padding != RCN_PADDING_NONE && padding != RCN_PADDING_SAME) return fail(RCN_ERR_INVALID, "unknown padding %d", padding);
    // kernel.rs:123-128
    if ((kh == 0 && kw == 0) || kh > H || kw > W || kh == 0 || kw == 0)
        return fail(RCN_ERR_SHAPE, "convolve_2d expects 'self.shape() >= kernel_shape() > 0', received (%zu, %zu) and (%zu, %zu) respectively.", H, W, kh, kw);
    // kernel.rs:131-135
    if ((kh % 2 == 0 || kw % 2 == 0) && padding == RCN_PADDING_SAME)
        return fail(RCN_ERR_SHAPE, "convolve_2d expects kernel dimensions to be odd when padding mode set to 'SAME', got (%zu, %zu)", kh, kw);
    // kernel.rs:154-158: the padded copy indexes self[(cy-1, cx-1)] up to cy = H + kh/2 - 1 => out of bounds for kh/2 >= 2
    if (padding == RCN_PADDING_SAME && (kh / 2 >= 2 || kw / 2 >= 2))
        return fail(RCN_ERR_OUT_OF_BOUNDS, "Matrix index out of bounds. (SAME padding with a (%zu, %zu) kernel, kernel.rs:156)", kh, kw);
    if (H > 32768 || W > 32768) return fail(RCN_ERR_INVALID, "matrix too large");
    OpCtx c;
    RCN_TRY(c.enter(device, cuda_stream));
    const bool same = padding == RCN_PADDING_SAME;
    const size_t on = same ? H * W : (H - kh + 1) * (W - kw + 1);
    const void *m_dev = nullptr, *k_dev = nullptr; void* o_dev = nullptr; bool host = false;
    RCN_TRY(c.in(m, H * W * 8, tl_op_in, &m_dev));
    RCN_TRY(c.in(kernel, kh * kw * 8, tl_op_k, &k_dev));
    RCN_TRY(c.out(out, on * 8, tl_op_out, &o_dev, &host));
    RCN_TRY(launch_convolve_2d((const double*)m_dev, H, W, (const double*)k_dev, kh, kw, padding, (double*)o_dev, c.stream));
    return c.finish(out, o_dev, on * 8, host);
}

int rcn_cuda_convolve_2d_separated(int device, void* cuda_stream, const double* m, size_t H, size_t W, int op, int padding,
                                   double* out) {
    if (!m || !out) return fail(RCN_ERR_INVALID, "null pointer");
    if (padding != RCN_PADDING_NONE && padding != RCN_PADDING_SAME) return fail(RCN_ERR_INVALID, "unknown padding %d", padding);
    if (op < 0 || op > 3) return fail(RCN_ERR_INVALID, "unknown SeparableOperator %d", op);
    if (H < 3 || W < 3)  // kernel.rs:199-201
        return fail(RCN_ERR_SHAPE, "convolve_2d_separated expects 'self.shape() >= kernel_shape() > 0', received (%zu, %zu) and (3, 3) respectively.", H, W);
    if (H > 32768 || W > 32768) return fail(RCN_ERR_INVALID, "matrix too large");
    OpCtx c;
    RCN_TRY(c.enter(device, cuda_stream));
    const bool same = padding == RCN_PADDING_SAME;
    const size_t on = same ? H * W : (H - 2) * (W - 2);
    const void* m_dev = nullptr; void* o_dev = nullptr; bool host = false;
    RCN_TRY(c.in(m, H * W * 8, tl_op_in, &m_dev));
    RCN_TRY(c.out(out, on * 8, tl_op_out, &o_dev, &host));
    RCN_TRY(launch_convolve_2d_separated((const double*)m_dev, H, W, op, padding, (double*)o_dev, c.stream));
    return c.finish(out, o_dev, on * 8, host);
}

int rcn_cuda_relu(int device, void* cuda_stream, const double* m, size_t n, double* out) {
    if (n == 0) return RCN_OK;
    if (!m || !out) return fail(RCN_ERR_INVALID, "null pointer");
    OpCtx c;
    RCN_TRY(c.enter(device, cuda_stream));
    const void* m_dev = nullptr; void* o_dev = nullptr; bool host = false;
    RCN_TRY(c.in(m, n * 8, tl_op_in, &m_dev));
    RCN_TRY(c.out(out, n * 8, tl_op_out, &o_dev, &host));
    RCN_TRY(launch_relu((const double*)m_dev, n, (double*)o_dev, c.stream));
    return c.finish(out, o_dev, n * 8, host);
}

int rcn_cuda_pool_2d(int device, void* cuda_stream, const double* m, size_t H, size_t W, int padding, int pooling,
                     double* out, uint8_t* argmax_out) {
    if (!m || !out) return fail(RCN_ERR_INVALID, "null pointer");
    if (padding != RCN_PADDING_NONE && padding != RCN_PADDING_SAME) return fail(RCN_ERR_INVALID, "unknown padding %d", padding);
    if (pooling != RCN_POOLING_AVERAGE && pooling != RCN_POOLING_MAX) return fail(RCN_ERR_INVALID, "unknown pooling %d", pooling);
    if (H < 2 || W < 2)  // kernel.rs:246-251
        return fail(RCN_ERR_SHAPE, "stride_2d expected a matrix with dimensions greater than (2, 2), got (%zu, %zu)", H, W);
    if (pooling != RCN_POOLING_MAX) return fail(RCN_ERR_NOT_IMPLEMENTED, "Not implemented");  // kernel.rs:283-285
    if (H > 32768 || W > 32768) return fail(RCN_ERR_INVALID, "matrix too large");
    OpCtx c;
    RCN_TRY(c.enter(device, cuda_stream));
    const bool same = padding == RCN_PADDING_SAME;
    const size_t oh = same ? (H + 1) / 2 : H / 2, ow = same ? (W + 1) / 2 : W / 2;
    const size_t on = oh * ow;
    const void* m_dev = nullptr; void* o_dev = nullptr; bool host = false;
    RCN_TRY(c.in(m, H * W * 8, tl_op_in, &m_dev));
    RCN_TRY(c.out(out, on * 8, tl_op_out, &o_dev, &host));
    // aux: [int nan_flag | pad to 16 | argmax bytes]
    RCN_TRY(tl_op_aux.reserve(16 + on));
    int* flag = tl_op_aux.as<int>();
    RCN_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), c.stream));
    uint8_t* am_dev = nullptr;
    bool am_host = false;
    if (argmax_out) {
        if (is_device_ptr(argmax_out)) am_dev = argmax_out;
        else { am_dev = tl_op_aux.as<uint8_t>() + 16; am_host = true; }
    }
    RCN_TRY(launch_pool_2d((const double*)m_dev, H, W, padding, (double*)o_dev, am_dev, flag, c.stream));
    int nan = 0;
    RCN_CUDA_TRY(cudaMemcpyAsync(&nan, flag, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    if (am_host) RCN_CUDA_TRY(cudaMemcpyAsync(argmax_out, am_dev, on, cudaMemcpyDeviceToHost, c.stream));
    if (host) RCN_CUDA_TRY(cudaMemcpyAsync(out, o_dev, on * 8, cudaMemcpyDeviceToHost, c.stream));
    RCN_CUDA_TRY(cudaStreamSynchronize(c.stream));
    if (nan) return fail(RCN_ERR_NAN, "called `Option::unwrap()` on a `None` value (partial_cmp on NaN, kernel.rs:280)");
    return RCN_OK;
}

}  // extern "C"
