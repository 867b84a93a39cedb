// features.cuh -- feature-stage declarations (Sobel-separated conv -> ReLU -> 2x2 max-pool stacks).
#pragma once
#include "common.cuh"

namespace rcn {

// One executable stage of a convpool stack (rcn.rs:317-348). A Convolve2D layer directly followed by a
// Pool2D layer is fused into one stage so the 4x-larger conv output never leaves registers.
struct Stage {
    int kind;         // 0 = conv (+relu), 1 = conv (+relu) + max-pool, 2 = max-pool only
    int same;         // conv padding == Padding::Same
    int first;        // conv input is the raw image: one map, output slots [T,L,R,B] (rcn.rs:339)
    int n_in;         // input maps
    int h_in, w_in;   // input map size
    int h_c, w_c;     // conv output size (pool-only: == input size)
    int h_out, w_out; // stage output map size
    int n_out;        // output maps
};
constexpr int kMaxStages = 24;
struct StageList {
    Stage s[kMaxStages];
    int n;
};

struct FeaturePlan {
    StageList stages;
    size_t n_maps = 0, map_h = 0, map_w = 0, L = 0;
    size_t max_elems = 0;  // largest per-image map set any stage reads (elements)
    int n_conv = 0;
};

// Geometry of the staged path (u8 images through [Convolve2D(Same), Pool2D(Max)]^n, W % 4 == 0): every stage input is a
// set of int32 column-major tiles in shared memory with a zero frame (2 rows above / 2 columns left, >= 1 below / right,
// and the never-read last column of SURVEY.md A.2 kept at zero).  Shared by features_cp_kernel and kernel A's front end.
constexpr int kCpMaxStages = 8;
struct CpStage {
    int n_in, h, w;          // input maps of this stage
    int hp, map_elems;       // padded column pitch (even), padded elements per map
    int h_out, w_out;        // pooled output size
    int off;                 // offset (ints) of this stage's input tiles within one image's tile block
    unsigned magic_hw, magic_h, magic_items;   // ceil(2^32 / d) for d = h_out*w_out, h_out, n_in*h_out*w_out (all >= 2)
};
struct CpPlan {
    CpStage s[kCpMaxStages];
    int n;
    int stage_bytes;         // one u8 staging buffer (H*W rounded up to 128)
    int tile_ints;           // all padded tiles of one image
    int tasks_per_image;     // transpose warp-tasks per image
    unsigned magic_w4;       // ceil(2^32 / (W/4))            (0 when W/4 < 2)
    unsigned magic_chunks;   // ceil(2^32 / chunks per row group)  (0 when < 2)
    unsigned magic_tasks;    // ceil(2^32 / tasks_per_image)  (0 when < 2)
};
// false when the plan does not qualify (other layer kinds, W % 4 != 0, 1-row maps, > 10 conv layers).
bool make_cp_plan(const FeaturePlan& plan, size_t H, size_t W, CpPlan* out);

// Walks convpool_cfg for an H x W input; returns RCN_ERR_SHAPE / RCN_ERR_NOT_IMPLEMENTED where the
// reference would panic.
int plan_features(const int32_t* cfg, size_t n_cfg, size_t H, size_t W, FeaturePlan* plan);

// Scratch owned by the caller (model): ping-pong buffers for the layer-by-layer path.
struct FeatureScratch {
    DevBuf a, b;
};

// Standardise + clamp epilogue (rcn.rs:407-412). mode 0: off; 1: IEEE division; 2: host-verified exact fast division.
struct Standardise {
    int mode;
    double mean, sd, rcp;
};

// Optional device-side indirection for epoch training (rcn.rs:144-149: the shuffled training set is walked in
// chunks_exact(batch) steps): image b of this step is dataset image perm[*cursor + b] (perm == NULL: identity) and
// its label is copied to labels_batch[b].  Everything is read on the device, so one captured CUDA graph replays
// for every step of every epoch.
struct BatchIndex {
    const long long* cursor = nullptr;
    const long long* perm = nullptr;
    const long long* labels_all = nullptr;
    long long* labels_batch = nullptr;
    // Streaming from a HOST dataset (rcn_cuda_train_epoch_host): `images` is a ring of `window` image slots and image
    // i of the walk lives in slot i % window (labels are still indexed by i).  0 = images hold the whole dataset.
    long long window = 0;
    // ... filled by the copy engine: `*arrived` = how many images of the walk have landed in the ring so far (written by a
    // stream memory operation behind each copy).  A kernel reading image i first waits for *arrived > i.  nullptr = no wait.
    const long long* arrived = nullptr;
    // Epoch mode: how the cursor advances after this step (chunks_exact wrap-around, rcn.rs:147) -- lets a kernel ask L2 for
    // the NEXT step's images ahead of time.  0 = unknown (no prefetch).
    long long batch = 0, n_samples = 0;
};
// epoch state block (int64 slots): the training cursor, padding; this step's labels follow
constexpr int kEpCursor = 0, kEpSlots = 8;

// Epilogue descriptor for a batch: picks the host-verified exact fast division when the input is u8.
Standardise make_standardise(const FeaturePlan& plan, int pixel_format, bool standardise, double mean, double sd);

// images: DEVICE pointer, B images in `pixel_format`; out: DEVICE L x B.
int launch_features(const FeaturePlan& plan, const void* images, int pixel_format, size_t B, size_t H, size_t W,
                    bool standardise, double mean, double sd, double* out, FeatureScratch& scratch,
                    cudaStream_t stream, const BatchIndex* index = nullptr);

int launch_standardise(double* feats, size_t n, double mean, double sd, cudaStream_t stream);
// Deterministic two-pass mean / population-sd; result[0]=mean, result[1]=sd (device), needs scratch.
int launch_gen_scales(const double* feats, size_t n, double* result_dev, DevBuf& scratch, cudaStream_t stream);

// op-level kernels (single matrix, device pointers)
int launch_convolve_2d(const double* m, size_t H, size_t W, const double* k, size_t kh, size_t kw, int padding,
                       double* out, cudaStream_t stream);
int launch_convolve_2d_separated(const double* m, size_t H, size_t W, int op, int padding, double* out,
                                 cudaStream_t stream);
int launch_relu(const double* m, size_t n, double* out, cudaStream_t stream);
int launch_pool_2d(const double* m, size_t H, size_t W, int padding, double* out, uint8_t* argmax, int* nan_flag,
                   cudaStream_t stream);

}  // namespace rcn
