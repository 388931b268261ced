// r360_api.cu -- host side of the C ABI declared in include/r360.h.
//
// Owns device memory, streams and launch order; all arithmetic of the path runs in the kernels
// of r360_kernels.cu.  There is no CPU fallback: every entry point needs a CUDA device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <dlfcn.h>
#include <cctype>
#include <map>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <nvtx3/nvToolsExt.h>          // header-only: ranges show up in nsys / ncu timelines, cost nothing without a tool attached
#include "r360_kernels.h"
#include "synth.h"

namespace {

struct NvtxRange {                      // one named range per C-ABI call that enqueues device work
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

thread_local std::string g_create_error;     // r360_create failures (no ctx yet); ranks may be host threads

#ifndef R360_CHUNK_FRAMES
#define R360_CHUNK_FRAMES 128
#endif
constexpr int kChunkFrames = R360_CHUNK_FRAMES;     // max frames per pyramid-build launch / H2D staging buffer
constexpr int kStages = 4;           // staging buffers: the copy stream runs up to kStages - 1 chunks ahead
constexpr int kStreamPairs = 64;     // pairs per registration batch of r360_register_host_pairs
constexpr int kLatencyPairs = 4;     // calls with at most this many pairs watch the active lists from the host and stop enqueueing
                                     // passes of a level once every pair has left it (the single-pair call shape of the class mirror)
constexpr int kOccPairs = 64;        // pairs per batch of the occlusion variants (bounds their scratch: 12 B per pixel and pair)

struct Ctx {
    int device = 0, sm_count = 0, pass_grid = 0;
    int rows = 0, cols = 0, L = 0, max_frames = 0, max_pairs = 0;
    r360_params P{};
    R360Level lv[R360_MAX_LEVELS]{};
    long long px_total = 0;
    cudaStream_t st = nullptr, cs = nullptr, st2 = nullptr;      // compute, copy, and the side stream of the error-only passes
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool overlap_err = false;                    // R360_OVERLAP_ERR=1: error-only launches on a side stream (measured: no gain -- the
                                                 // persistent grids fill the chip either way; profiles/r02_overlap_err_ab.txt)
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_copy[kStages]{}, ev_done[kStages]{};
    int chunk = kChunkFrames;                    // frames per chunk = min(kChunkFrames, max_frames)
    bool pin_packed = true;                      // R360_PIN_PACKED=0: the scalar one-pixel-per-thread pinhole evaluation (A/B measurements)
    bool use_pyr_mid = true;                     // R360_PYR_MID=0: k_down + k_texel for the levels >= 1 (A/B measurements)
    bool use_pyr_head = true;                    // R360_PYR_HEAD=0 in the environment: the separate level-0 kernels (A/B measurements)
    long long n_chunks_done = 0;                 // staging-buffer rotation across calls
    std::vector<cudaEvent_t> ev_pass;            // pairs of events around each pixel pass
    float* d_tables = nullptr;
    // frame slots
    std::vector<float2*> src;                    // {depth, gray} pyramids
    std::vector<float*> trg;                     // texel pyramids
    // which pyramid of a slot holds the frame LAST set there: an allocation outlives a re-set with a narrower
    // role, its content must not be used afterwards (check_pair / dump_* test these, not the pointers)
    std::vector<uint8_t> have_src, have_trg;
    float2* scratch = nullptr;                   // `chunk` source-format pyramids (target-only frames)
    uint8_t* stage_rgb[kStages]{};
    uint16_t* stage_depth[kStages]{};            // also holds float depth (sized for it)
    // pointer tables for the pyramid kernels (pinned host + device)
    float2** h_pyr = nullptr; float2** d_pyr = nullptr;
    float2** h_pyr_t = nullptr; float2** d_pyr_t = nullptr;
    float** h_trg_t = nullptr; float** d_trg_t = nullptr;
    // fused pyramid head: per frame of a chunk, where level 0 goes (null: target-only frame), where level 1
    // goes, where the level-0 texels go (null: source-only frame)
    float2** h_l0 = nullptr; float2** d_l0 = nullptr;
    float2** h_l1 = nullptr; float2** d_l1 = nullptr;
    float** h_tex = nullptr; float** d_tex = nullptr;
    // pair batch
    R360Pair* d_pairs = nullptr;
    R360Fx* d_acc = nullptr; int* d_cnt = nullptr;     // per pair: R360_ACC_STRIDE fixed-point sums, R360_ACC_INTS counters
    int* d_active = nullptr; int* d_nactive = nullptr;
    int* h_nactive = nullptr;                    // pinned copy of d_nactive (latency mode)
    bool early_exit = true;                      // R360_EARLY_EXIT=0: always enqueue the full schedule (A/B)
    int* d_active_err = nullptr;                 // pairs whose next pass is error-only
    int pass_grid_err = 0;
    bool speculate = true;                       // R360_SPECULATE=0 in the environment: every pass is the fused one (A/B)
    float spec_margin = 1.0f;                    // R360_SPEC_MARGIN: threshold of the prediction in units of tol_residual (A/B)
    int dyn_permille = 150;                      // R360_DYN_PERMILLE: share of a k_pass launch's items handed out dynamically (A/B; 0 = static)
    const float2** h_srcb = nullptr; const float2** d_srcb = nullptr;
    const float** h_trgb = nullptr; const float** d_trgb = nullptr;
    int32_t* h_idx = nullptr; int32_t* d_idx = nullptr;          // src idx | trg idx
    float* h_pose = nullptr; float* d_pose = nullptr;
    r360_result* d_res = nullptr; r360_result* h_res = nullptr;
    r360_iter_record* d_trace = nullptr; size_t trace_cap = 0;
    float* d_cams = nullptr; float* h_cams = nullptr;
    int* occ_head = nullptr; int* occ_next = nullptr; float* occ_dinv = nullptr;   // occlusion 1/2: per-texel candidate lists
    int occ_cap = 0;                                                                // pairs the scratch holds
    float cam[4] = {0.f, 0.f, 0.f, 0.f}; bool have_cam = false;                     // setCameraMatrix (pinhole contexts)
    const float2** h_rig_src = nullptr; const float2** d_rig_src = nullptr;          // the 8-sensor rig: per pair 8 source /
    const float** h_rig_trg = nullptr; const float** d_rig_trg = nullptr;            // target pyramids (allocated on first use)
    std::map<void*, std::pair<size_t, bool>> host_allocs;                           // r360_host_alloc: ptr -> (bytes, mmap'ed + registered?)
    uint8_t* d_gather = nullptr; size_t gather_cap = 0;                             // r360_allgather_results: send | receive records
    uint8_t* d_sens_rgb = nullptr; uint16_t* d_sens_depth = nullptr; size_t sens_cap = 0;   // ingest: sensor images of one chunk
    // stats
    float last_ms = 0.f, pass_ms = 0.f;
    int pass_launches = 0;
    double pass_bytes = 0.0;
    int64_t launches = 0;
    std::string err;
};

// NCCL entry points, resolved at run time: the library the host process already uses (a C++ host links it, a
// Python host has it loaded by torch), else the system libnccl.  No link-time dependency.
typedef int (*NcclAllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*NcclErrorStringFn)(int);
struct NcclApi { NcclAllGatherFn all_gather = nullptr; NcclErrorStringFn error_string = nullptr; };
NcclApi load_nccl() {
    NcclApi api;
    void* f = dlsym(RTLD_DEFAULT, "ncclAllGather");
    void* e = f ? dlsym(RTLD_DEFAULT, "ncclGetErrorString") : nullptr;
    if (!f) {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) { f = dlsym(h, "ncclAllGather"); e = dlsym(h, "ncclGetErrorString"); }
    }
    api.all_gather = reinterpret_cast<NcclAllGatherFn>(f);
    api.error_string = reinterpret_cast<NcclErrorStringFn>(e);
    return api;
}
const NcclApi& nccl_api() {
    static const NcclApi api = load_nccl();          // thread-safe initialisation: ranks may be host threads
    return api;
}

int fail(Ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(c, call)                                                                           \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail((c), e_ == cudaErrorMemoryAllocation ? R360_E_NOMEM : R360_E_CUDA,    \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int ensure_src(Ctx* c, int slot) {
    // + 2: the pass kernel reads pixel pairs as float4, an odd-sized last level reads one pixel past its end
    if (!c->src[slot]) CK(c, cudaMalloc(&c->src[slot], sizeof(float2) * (c->px_total + 2)));
    return R360_OK;
}
int ensure_trg(Ctx* c, int slot) {
    if (!c->trg[slot]) CK(c, cudaMalloc(&c->trg[slot], sizeof(float) * R360_TEXEL_FLOATS * c->px_total));
    return R360_OK;
}

R360PassArgs pass_args(Ctx* c, int level, int n_pairs_hint, int first = 0, bool err_only = false) {
    R360PassArgs a{};
    a.lv = c->lv[level];
    a.params = c->P;
    a.inv_std_photo = (float)(1. / c->P.std_photo);
    a.one = 1.0f;
    // items: the granularity of the contiguous per-CTA runs (>= 8 items per CTA when there is enough work)
    const long long n = c->lv[level].n;
    const long long blk = 2LL * R360_PASS_THREADS;                  // pixels per CTA iteration
    long long want = n * (long long)std::max(n_pairs_hint, 1) / (8LL * (err_only ? c->pass_grid_err : c->pass_grid));
    long long ppi = ((want + blk - 1) / blk) * blk;
    ppi = std::min<long long>(std::max<long long>(ppi, blk), 16 * blk);
    a.px_per_item = (int)ppi;
    a.items_per_pair = (int)((n + ppi - 1) / ppi);
    // pair-indexed arrays are addressed relative to `first` (pair ids inside the kernels are batch-local)
    a.n_active = err_only ? c->d_nactive + 3 : c->d_nactive;
    a.work_counter = c->d_nactive + (err_only ? 5 : 4);
    a.dyn_permille = c->dyn_permille;
    a.active_list = (err_only ? c->d_active_err : c->d_active) + first;
    a.pairs = c->d_pairs + first;
    a.src_base = c->d_srcb + first;
    a.trg_base = c->d_trgb + first;
    a.acc = c->d_acc + (size_t)first * R360_ACC_STRIDE;
    a.cnt = c->d_cnt + (size_t)first * R360_ACC_INTS;
    return a;
}

int trace_per_level(const Ctx* c) { return c->P.projection == R360_PINHOLE ? 2 * c->P.max_iters + 2 : c->P.max_iters + 2; }

R360GnArgs gn_args(Ctx* c, int n_pairs, r360_iter_record* trace, int first = 0) {
    R360GnArgs g{};
    g.params = c->P;
    g.n_pairs = n_pairs;
    g.pairs = c->d_pairs + first;
    g.acc = c->d_acc + (size_t)first * R360_ACC_STRIDE;
    g.cnt = c->d_cnt + (size_t)first * R360_ACC_INTS;
    g.active_list = c->d_active + first;
    g.n_active = c->d_nactive;
    g.active_list_err = c->d_active_err + first;
    g.n_active_err = c->d_nactive + 3;
    g.speculate = c->speculate ? 1 : 0;
    g.spec_margin = c->spec_margin;
    g.ticket = c->d_nactive + 2;
    g.work_counters = c->d_nactive + 4;
    g.trace = trace ? trace + (size_t)first * c->L * trace_per_level(c) : nullptr;
    return g;
}

// Host side of the per-frame pointer tables of ONE chunk (frames [first, first+n), table entries [table_off, table_off+n)):
// where each frame's pyramid / texels go, by role.  Returns the number of target frames through n_t.
int fill_tables(Ctx* c, int first, int n, const uint8_t* roles, int table_off, int* n_t_out) {
    int n_t = 0;
    for (int k = 0; k < n; ++k) {
        const int slot = first + k;
        const int role = roles ? roles[k] : R360_ROLE_BOTH;
        c->have_src[slot] = (role & R360_ROLE_SOURCE) ? 1 : 0;
        c->have_trg[slot] = (role & R360_ROLE_TARGET) ? 1 : 0;
        if (role & R360_ROLE_SOURCE) {
            int rc = ensure_src(c, slot);
            if (rc) return rc;
            c->h_pyr[table_off + k] = c->src[slot];
        } else {
            c->h_pyr[table_off + k] = c->scratch + (size_t)k * c->px_total;
        }
        if (role & R360_ROLE_TARGET) {
            int rc = ensure_trg(c, slot);
            if (rc) return rc;
            c->h_pyr_t[table_off + n_t] = c->h_pyr[table_off + k];
            c->h_trg_t[table_off + n_t] = c->trg[slot];
            ++n_t;
        }
        c->h_l0[table_off + k] = (role & R360_ROLE_SOURCE) ? c->h_pyr[table_off + k] : nullptr;
        c->h_l1[table_off + k] = c->L >= 2 ? c->h_pyr[table_off + k] + c->lv[1].px_off : nullptr;
        c->h_tex[table_off + k] = (role & R360_ROLE_TARGET) ? c->trg[slot] : nullptr;
    }
    *n_t_out = n_t;
    return R360_OK;
}
// Device copies of table entries [table_off, table_off+n): once per call for all its chunks (set_frames), not per chunk --
// six small copies in front of every chunk's kernels cost the build 0.2 ms per 1024 frames.
int upload_tables(Ctx* c, int table_off, int n) {
    CK(c, cudaMemcpyAsync(c->d_pyr + table_off, c->h_pyr + table_off, sizeof(float2*) * n, cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemcpyAsync(c->d_pyr_t + table_off, c->h_pyr_t + table_off, sizeof(float2*) * n, cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemcpyAsync(c->d_trg_t + table_off, c->h_trg_t + table_off, sizeof(float*) * n, cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemcpyAsync(c->d_l0 + table_off, c->h_l0 + table_off, sizeof(float2*) * n, cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemcpyAsync(c->d_l1 + table_off, c->h_l1 + table_off, sizeof(float2*) * n, cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemcpyAsync(c->d_tex + table_off, c->h_tex + table_off, sizeof(float*) * n, cudaMemcpyHostToDevice, c->st));
    return R360_OK;
}

// Pyramid build of frames [first, first+n) from inputs resident on the device (or staged there).  n_t_ready >= 0: the
// tables of this chunk are already on the device (the caller filled and uploaded them for the whole call).
int build_chunk(Ctx* c, int first, int n, const uint8_t* rgb_dev, const uint16_t* depth_mm_dev,
                const float* depth_m_dev, const uint8_t* roles, int table_off, int n_t_ready = -1) {
    int n_t = n_t_ready;
    if (n_t_ready < 0) {
        int rc = fill_tables(c, first, n, roles, table_off, &n_t);
        if (rc) return rc;
        rc = upload_tables(c, table_off, n);
        if (rc) return rc;
    }
    // Levels 0 and 1 and the level-0 texels come from ONE pass over the raw input (k_pyr_head) when the
    // pyramid has a level 1; the remaining levels from k_down / k_texel.
    const bool fused_head = c->L >= 2 && c->use_pyr_head;
    if (fused_head) {
        r360_launch_pyr_head(c->st, rgb_dev, depth_mm_dev, depth_m_dev, c->d_l0 + table_off, c->d_l1 + table_off,
                             c->d_tex + table_off, c->rows, c->cols, c->P.min_depth, c->P.max_depth, c->P.n_sensors_mask, n);
        ++c->launches;
    } else {
        r360_launch_level0(c->st, rgb_dev, depth_mm_dev, depth_m_dev, c->d_pyr + table_off, n, c->rows * c->cols, c->sm_count);
        ++c->launches;
    }
    if (fused_head && c->use_pyr_mid) {
        // levels >= 1: the level's target texels and the next level from ONE read of its plane (k_texel + k_down in one kernel)
        for (int l = 1; l < c->L; ++l) {
            const bool next = l + 1 < c->L;
            if (!next && !n_t) break;
            r360_launch_pyr_mid(c->st, c->d_pyr + table_off, n_t ? c->d_tex + table_off : nullptr, c->lv[l].px_off,
                                next ? c->lv[l + 1].px_off : -1, c->lv[l].rows, c->lv[l].cols, c->P.min_depth, c->P.max_depth,
                                c->P.n_sensors_mask, n);
            ++c->launches;
        }
        CK(c, cudaGetLastError());
        return R360_OK;
    }
    for (int l = fused_head ? 2 : 1; l < c->L; ++l) {
        r360_launch_down(c->st, c->d_pyr + table_off, c->lv[l - 1].px_off, c->lv[l].px_off, c->lv[l - 1].rows,
                         c->lv[l - 1].cols, c->P.min_depth, c->P.max_depth, n, c->sm_count);
        ++c->launches;
    }
    if (n_t)
        for (int l = fused_head ? 1 : 0; l < c->L; ++l) {
            r360_launch_texel(c->st, c->d_pyr_t + table_off, c->d_trg_t + table_off, c->lv[l].px_off, c->lv[l].rows,
                              c->lv[l].cols, c->P.n_sensors_mask, n_t, c->sm_count);
            ++c->launches;
        }
    CK(c, cudaGetLastError());
    return R360_OK;
}

// One chunk of host frames: H2D on the copy stream into the next staging buffer (the copy stream
// runs up to kStages - 1 chunks ahead of the compute stream), pyramid build on the compute stream.
int stage_and_build(Ctx* c, int first_slot, int m, const uint8_t* rgb, const uint8_t* depth, bool depth_is_f32,
                    const uint8_t* roles, int table_off, int n_t_ready = -1) {
    const size_t npx = (size_t)c->rows * c->cols;
    const size_t dsz = depth_is_f32 ? sizeof(float) : sizeof(uint16_t);
    const int b = (int)(c->n_chunks_done % kStages);
    if (c->n_chunks_done >= kStages) CK(c, cudaStreamWaitEvent(c->cs, c->ev_done[b], 0));
    CK(c, cudaMemcpyAsync(c->stage_rgb[b], rgb, (size_t)m * npx * 3, cudaMemcpyHostToDevice, c->cs));
    CK(c, cudaMemcpyAsync(c->stage_depth[b], depth, (size_t)m * npx * dsz, cudaMemcpyHostToDevice, c->cs));
    CK(c, cudaEventRecord(c->ev_copy[b], c->cs));
    CK(c, cudaStreamWaitEvent(c->st, c->ev_copy[b], 0));
    int rc = build_chunk(c, first_slot, m, c->stage_rgb[b], depth_is_f32 ? nullptr : (const uint16_t*)c->stage_depth[b],
                         depth_is_f32 ? (const float*)c->stage_depth[b] : nullptr, roles, table_off, n_t_ready);
    if (rc) return rc;
    CK(c, cudaEventRecord(c->ev_done[b], c->st));
    ++c->n_chunks_done;
    return R360_OK;
}

int set_frames_impl(Ctx* c, int first, int n, const uint8_t* rgb, const void* depth, bool depth_is_f32,
                    bool on_device, const uint8_t* roles) {
    if (!c) return R360_E_ARG;
    NvtxRange nvtx("r360_set_frames");
    if (first < 0 || n < 0 || first + n > c->max_frames || !rgb || !depth)
        return fail(c, R360_E_ARG, "set_frames: bad range [%d,%d) of %d slots or null input", first, first + n, c->max_frames);
    if (roles)
        for (int k = 0; k < n; ++k)
            if (roles[k] < 1 || roles[k] > 3) return fail(c, R360_E_ARG, "set_frames: role %d of frame %d invalid", roles[k], k);
    CK(c, cudaSetDevice(c->device));
    const size_t npx = (size_t)c->rows * c->cols;
    const size_t dsz = depth_is_f32 ? sizeof(float) : sizeof(uint16_t);
    CK(c, cudaEventRecord(c->ev_t0, c->st));
    std::vector<int> n_t((size_t)(n + c->chunk - 1) / c->chunk + 1, 0);
    for (int off = 0; off < n; off += c->chunk) {                    // the pointer tables of every chunk, uploaded once
        int rc = fill_tables(c, first + off, std::min(c->chunk, n - off), roles ? roles + off : nullptr, off, &n_t[off / c->chunk]);
        if (rc) return rc;
    }
    if (n > 0) {
        int rc = upload_tables(c, 0, n);
        if (rc) return rc;
    }
    for (int off = 0; off < n; off += c->chunk) {
        const int m = std::min(c->chunk, n - off);
        int rc = on_device
                     ? build_chunk(c, first + off, m, rgb + (size_t)off * npx * 3,
                                   depth_is_f32 ? nullptr : (const uint16_t*)((const uint8_t*)depth + (size_t)off * npx * dsz),
                                   depth_is_f32 ? (const float*)((const uint8_t*)depth + (size_t)off * npx * dsz) : nullptr,
                                   roles ? roles + off : nullptr, off, n_t[off / c->chunk])
                     : stage_and_build(c, first + off, m, rgb + (size_t)off * npx * 3, (const uint8_t*)depth + (size_t)off * npx * dsz,
                                       depth_is_f32, roles ? roles + off : nullptr, off, n_t[off / c->chunk]);
        if (rc) return rc;
    }
    CK(c, cudaEventRecord(c->ev_t1, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    CK(c, cudaEventElapsedTime(&c->last_ms, c->ev_t0, c->ev_t1));
    return R360_OK;
}

int check_pair(Ctx* c, int src, int trg) {
    if (src < 0 || src >= c->max_frames || trg < 0 || trg >= c->max_frames)
        return fail(c, R360_E_ARG, "frame index out of range (src %d, trg %d, %d slots)", src, trg, c->max_frames);
    if (!c->have_src[src]) return fail(c, R360_E_STATE, "frame %d has no source pyramid (r360_set_frames with a SOURCE role first)", src);
    if (!c->have_trg[trg]) return fail(c, R360_E_STATE, "frame %d has no target pyramid (r360_set_frames with a TARGET role first)", trg);
    return R360_OK;
}

// One fused pixel pass over a single pair at `pose` (the eval / dump hooks).
int eval_setup(Ctx* c, int src, int trg, int level, const float pose[16], R360PassArgs* a) {
    if (level < 0 || level >= c->L || !pose) return fail(c, R360_E_ARG, "eval: bad level %d or null pose", level);
    int rc = check_pair(c, src, trg);
    if (rc) return rc;
    CK(c, cudaSetDevice(c->device));
    const int slot = c->max_pairs;          // spare pair slot
    R360Pair hp;
    memset(&hp, 0, sizeof(hp));
    memcpy(hp.pose_eval, pose, 64);
    memcpy(hp.pose_estim, pose, 64);
    hp.active = 1;
    CK(c, cudaMemcpyAsync(c->d_pairs + slot, &hp, sizeof(hp), cudaMemcpyHostToDevice, c->st));
    const float2* sb = c->src[src];
    const float* tb = c->trg[trg];
    CK(c, cudaMemcpyAsync(c->d_srcb + slot, &sb, sizeof(sb), cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemcpyAsync(c->d_trgb + slot, &tb, sizeof(tb), cudaMemcpyHostToDevice, c->st));
    const int one = 1;
    CK(c, cudaMemcpyAsync(c->d_active + slot, &slot, sizeof(int), cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemcpyAsync(c->d_nactive + 1, &one, sizeof(int), cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemsetAsync(c->d_nactive + 4, 0, sizeof(int) * 2, c->st));     // dynamic-item counters
    CK(c, cudaMemsetAsync(c->d_acc + (size_t)slot * R360_ACC_STRIDE, 0, sizeof(R360Fx) * R360_ACC_STRIDE, c->st));
    CK(c, cudaMemsetAsync(c->d_cnt + (size_t)slot * R360_ACC_INTS, 0, sizeof(int) * R360_ACC_INTS, c->st));
    CK(c, cudaStreamSynchronize(c->st));    // hp / sb / tb are stack variables
    *a = pass_args(c, level, 1);
    a->n_active = c->d_nactive + 1;
    a->active_list = c->d_active + slot;
    return R360_OK;
}

// One evaluation of the active pairs of `a` at their pose_eval: the fused pass (occlusion 0) or the
// scatter + evaluate pair of the occlusion variants.
R360PinLevel pin_level(const Ctx* c, int level) {
    R360PinLevel pl;
    const float scaleFactor = 1.0 / pow(2, level);              // RPI.h:569
    pl.fx = c->cam[0] * scaleFactor; pl.fy = c->cam[1] * scaleFactor;
    pl.ox = c->cam[2] * scaleFactor; pl.oy = c->cam[3] * scaleFactor;
    pl.inv_fx = 1. / pl.fx; pl.inv_fy = 1. / pl.fy;              // RPI.h:574-575
    return pl;
}

void launch_evaluation(Ctx* c, const R360PassArgs& a, int n_pairs, int level) {
    if (c->P.projection == R360_PINHOLE) {
        r360_launch_pin_eval(c->st, a, pin_level(c, level), n_pairs, c->sm_count, c->pin_packed);
        ++c->launches;
    } else if (c->P.occlusion == 0) {
        r360_launch_pass(c->st, a, c->pass_grid);
        ++c->launches;
    } else {
        r360_launch_occ_pass(c->st, a, n_pairs, c->occ_head, c->occ_next, c->occ_dinv, c->sm_count);
        c->launches += 3;
    }
}

int eval_pass(Ctx* c, int src, int trg, int level, const float pose[16], double acc[R360_ACC_STRIDE], int cnt[R360_ACC_INTS]) {
    R360PassArgs a;
    int rc = eval_setup(c, src, trg, level, pose, &a);
    if (rc) return rc;
    if (c->P.projection == R360_PINHOLE && !c->have_cam) return fail(c, R360_E_STATE, "pinhole context: call r360_set_camera first (setCameraMatrix, RPI.h:254)");
    launch_evaluation(c, a, 1, level);
    CK(c, cudaGetLastError());
    const int slot = c->max_pairs;
    R360Fx fx[R360_ACC_STRIDE];
    CK(c, cudaMemcpyAsync(fx, c->d_acc + (size_t)slot * R360_ACC_STRIDE, sizeof(R360Fx) * R360_ACC_STRIDE, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaMemcpyAsync(cnt, c->d_cnt + (size_t)slot * R360_ACC_INTS, sizeof(int) * R360_ACC_INTS, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    for (int k = 0; k < R360_ACC_STRIDE; ++k) acc[k] = k < R360_ACC_DOUBLES + 1 ? r360_fx_get(fx, k) : 0.0;
    return R360_OK;
}

}  // namespace

// ============================================================================ C ABI
extern "C" {

struct r360_ctx : Ctx {};

int r360_version(void) { return 100; }

void r360_default_params(r360_params* p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->n_levels = 4;                       // RPI.h:204
    p->min_depth = 0.3f;                   // RPI.h:202   (float member = 0.3 double literal)
    p->max_depth = 6.0f;                   // RPI.h:203
    p->std_photo = (float)(6. / 255);      // RPI.h:208
    p->std_depth = (float)0.2;             // RPI.h:211
    p->thres_sal_int = 0.01f;              // RPI.h:218
    p->thres_sal_depth = 0.01f;            // RPI.h:219
    p->max_iters = 10;                     // RPI.h:4593
    p->tol_residual = 1e-3;                // RPI.h:4594
    p->tol_update = 1e-4;                  // RPI.h:4595
    p->method = R360_PHOTO_DEPTH;          // what every caller passes
    p->occlusion = 0;
    p->n_sensors_mask = 8;                 // RPI.h:4537
}

void r360_default_params_pinhole(r360_params* p) {
    if (!p) return;
    r360_default_params(p);
    p->projection = R360_PINHOLE;
    p->tol_residual = 1e-4;                // RPI.h:4308
    p->n_sensors_mask = 0;                 // the sensor-joint mask belongs to alignFrames360 (RPI.h:4537)
}

const char* r360_last_error(const r360_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void r360_destroy(r360_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->st) cudaStreamSynchronize(c->st);
    if (c->cs) cudaStreamSynchronize(c->cs);
    for (auto p : c->src) if (p) cudaFree(p);
    for (auto p : c->trg) if (p) cudaFree(p);
    cudaFree(c->d_tables); cudaFree(c->scratch);
    for (int b = 0; b < kStages; ++b) { cudaFree(c->stage_rgb[b]); cudaFree(c->stage_depth[b]); }
    cudaFree(c->d_pyr); cudaFree(c->d_pyr_t); cudaFree(c->d_trg_t);
    cudaFreeHost(c->h_pyr); cudaFreeHost(c->h_pyr_t); cudaFreeHost(c->h_trg_t);
    cudaFree(c->d_l0); cudaFree(c->d_l1); cudaFree(c->d_tex);
    cudaFreeHost(c->h_l0); cudaFreeHost(c->h_l1); cudaFreeHost(c->h_tex);
    cudaFree(c->d_pairs); cudaFree(c->d_acc); cudaFree(c->d_cnt); cudaFree(c->d_active); cudaFree(c->d_nactive);
    cudaFree(c->d_active_err); cudaFreeHost(c->h_nactive);
    cudaFree(c->d_srcb); cudaFree(c->d_trgb); cudaFreeHost(c->h_srcb); cudaFreeHost(c->h_trgb);
    cudaFree(c->d_idx); cudaFreeHost(c->h_idx); cudaFree(c->d_pose); cudaFreeHost(c->h_pose);
    cudaFree(c->d_res); cudaFreeHost(c->h_res); cudaFree(c->d_trace);
    cudaFree(c->d_cams); cudaFreeHost(c->h_cams);
    cudaFree(c->d_sens_rgb); cudaFree(c->d_sens_depth); cudaFree(c->d_gather);
    cudaFree(c->occ_head); cudaFree(c->occ_next); cudaFree(c->occ_dinv);
    cudaFreeHost(c->h_rig_src); cudaFree(c->d_rig_src); cudaFreeHost(c->h_rig_trg); cudaFree(c->d_rig_trg);
    for (auto& kv : c->host_allocs) {
        if (kv.second.second) { cudaHostUnregister(kv.first); munmap(kv.first, kv.second.first); }
        else cudaFreeHost(kv.first);
    }
    for (auto e : c->ev_pass) cudaEventDestroy(e);
    for (int b = 0; b < kStages; ++b) { if (c->ev_copy[b]) cudaEventDestroy(c->ev_copy[b]); if (c->ev_done[b]) cudaEventDestroy(c->ev_done[b]); }
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->st2) cudaStreamDestroy(c->st2);
    if (c->st) cudaStreamDestroy(c->st);
    if (c->cs) cudaStreamDestroy(c->cs);
    delete c;
}

static int create_impl(r360_ctx* c, int device, int rows, int cols, int max_frames, int max_pairs, const r360_params* params) {
    c->device = device; c->rows = rows; c->cols = cols; c->max_frames = max_frames; c->max_pairs = max_pairs;
    c->P = *params; c->L = params->n_levels;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(c, R360_E_CUDA, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    CK(c, cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(c, cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->pass_grid = c->sm_count * R360_PASS_CTAS;
    c->pass_grid_err = c->sm_count * R360_ERR_CTAS;
    if (const char* e = getenv("R360_SPECULATE")) c->speculate = atoi(e) != 0;
    if (const char* e = getenv("R360_SPEC_MARGIN")) c->spec_margin = (float)atof(e);
    if (const char* e = getenv("R360_DYN_PERMILLE")) c->dyn_permille = std::min(1000, std::max(0, atoi(e)));
    CK(c, r360_pass_init());
    if (params->occlusion != 0) CK(c, r360_occ_init());
    CK(c, cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    CK(c, cudaStreamCreateWithFlags(&c->cs, cudaStreamNonBlocking));
    CK(c, cudaStreamCreateWithFlags(&c->st2, cudaStreamNonBlocking));
    CK(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CK(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    if (const char* e = getenv("R360_OVERLAP_ERR")) c->overlap_err = atoi(e) != 0;
    CK(c, cudaEventCreate(&c->ev_t0));
    CK(c, cudaEventCreate(&c->ev_t1));
    c->chunk = std::min(kChunkFrames, max_frames);
    for (int b = 0; b < kStages; ++b) {
        CK(c, cudaEventCreateWithFlags(&c->ev_copy[b], cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&c->ev_done[b], cudaEventDisableTiming));
    }
    c->ev_pass.resize((size_t)2 * c->L * (params->max_iters + 1 + R360_SPEC_EXTRA));
    for (auto& ev : c->ev_pass) CK(c, cudaEventCreate(&ev));

    // ---- per-level geometry and trig tables (RPI.h:2553-2556, 4555-4569), host-computed with
    //      the pinned sin/cos so that they equal the oracle's tables bit for bit.
    size_t n_tab = 0;
    long long off = 0;
    for (int l = 0; l < c->L; ++l) {
        R360Level& v = c->lv[l];
        v.rows = rows >> l; v.cols = cols >> l; v.n = v.rows * v.cols;
        v.div_magic = ((1ULL << 40) + v.cols - 1) / v.cols;
        v.stride_r = (2 * R360_PASS_THREADS) / v.cols;
        v.stride_c = (2 * R360_PASS_THREADS) % v.cols;
        v.px_off = off; off += v.n;
        v.res = (float)(2 * R360_PI_D / v.cols);
        v.res_inv = 1 / v.res;
        v.half_rows = (float)(0.5 * v.rows - 0.5);
        n_tab += 2 * (size_t)(v.rows + v.cols);
    }
    c->px_total = off;
    std::vector<float> tab(n_tab);
    CK(c, cudaMalloc(&c->d_tables, sizeof(float) * n_tab));
    size_t t = 0;
    for (int l = 0; l < c->L; ++l) {
        R360Level& v = c->lv[l];
        // {sin c, sin c+1, cos c, cos c+1}(theta) per column pair and {sin, -cos}(phi_r) per row:
        // one 16-byte + one 8-byte load per pixel pair
        v.tab_t = reinterpret_cast<const float4*>(c->d_tables + t);
        v.tab_p = reinterpret_cast<const float2*>(c->d_tables + t + 2 * (size_t)v.cols);
        for (int k = 0; k < v.cols; ++k) {
            float th = k * v.res;
            r360_sincosf(th, &tab[t + 4 * (k >> 1) + (k & 1)], &tab[t + 4 * (k >> 1) + 2 + (k & 1)]);
        }
        for (int r = 0; r < v.rows; ++r) {
            float ph = (v.half_rows - r) * v.res, sp, cp;
            r360_sincosf(ph, &sp, &cp);
            tab[t + 2 * (size_t)v.cols + 2 * r] = sp;
            tab[t + 2 * (size_t)v.cols + 2 * r + 1] = -cp;
        }
        t += 2 * (size_t)(v.rows + v.cols);
    }
    CK(c, cudaMemcpy(c->d_tables, tab.data(), sizeof(float) * n_tab, cudaMemcpyHostToDevice));

    c->src.assign(max_frames, nullptr);
    c->trg.assign(max_frames, nullptr);
    c->have_src.assign(max_frames, 0);
    c->have_trg.assign(max_frames, 0);
    const size_t npx = (size_t)rows * cols;
    CK(c, cudaMalloc(&c->scratch, sizeof(float2) * c->px_total * c->chunk));
    for (int b = 0; b < kStages; ++b) {
        CK(c, cudaMalloc(&c->stage_rgb[b], npx * 3 * c->chunk));
        CK(c, cudaMalloc(&c->stage_depth[b], npx * sizeof(float) * c->chunk));
    }
    CK(c, cudaMallocHost(&c->h_pyr, sizeof(void*) * max_frames)); CK(c, cudaMalloc(&c->d_pyr, sizeof(void*) * max_frames));
    CK(c, cudaMallocHost(&c->h_pyr_t, sizeof(void*) * max_frames)); CK(c, cudaMalloc(&c->d_pyr_t, sizeof(void*) * max_frames));
    CK(c, cudaMallocHost(&c->h_trg_t, sizeof(void*) * max_frames)); CK(c, cudaMalloc(&c->d_trg_t, sizeof(void*) * max_frames));
    CK(c, cudaMallocHost(&c->h_l0, sizeof(void*) * max_frames)); CK(c, cudaMalloc(&c->d_l0, sizeof(void*) * max_frames));
    CK(c, cudaMallocHost(&c->h_l1, sizeof(void*) * max_frames)); CK(c, cudaMalloc(&c->d_l1, sizeof(void*) * max_frames));
    CK(c, cudaMallocHost(&c->h_tex, sizeof(void*) * max_frames)); CK(c, cudaMalloc(&c->d_tex, sizeof(void*) * max_frames));
    if (const char* e = getenv("R360_PYR_HEAD")) c->use_pyr_head = atoi(e) != 0;
    if (const char* e = getenv("R360_PYR_MID")) c->use_pyr_mid = atoi(e) != 0;
    if (const char* e = getenv("R360_PIN_PACKED")) c->pin_packed = atoi(e) != 0;

    const int np = max_pairs + 1;     // + 1 spare slot for the eval hooks
    CK(c, cudaMalloc(&c->d_pairs, sizeof(R360Pair) * np));
    CK(c, cudaMalloc(&c->d_acc, sizeof(R360Fx) * R360_ACC_STRIDE * np));
    CK(c, cudaMalloc(&c->d_cnt, sizeof(int) * R360_ACC_INTS * np));
    CK(c, cudaMalloc(&c->d_active, sizeof(int) * np));
    CK(c, cudaMalloc(&c->d_active_err, sizeof(int) * np));
    // [0] batch (fused passes), [1] eval hooks, [2] completion ticket, [3] batch (error-only passes),
    // [4] / [5] dynamic-item counters of the fused / error-only k_pass launch
    CK(c, cudaMalloc(&c->d_nactive, sizeof(int) * 8));
    CK(c, cudaMemset(c->d_nactive, 0, sizeof(int) * 8));
    CK(c, cudaMallocHost(&c->h_nactive, sizeof(int) * 8));
    if (const char* e = getenv("R360_EARLY_EXIT")) c->early_exit = atoi(e) != 0;
    CK(c, cudaMallocHost(&c->h_srcb, sizeof(void*) * np)); CK(c, cudaMalloc(&c->d_srcb, sizeof(void*) * np));
    CK(c, cudaMallocHost(&c->h_trgb, sizeof(void*) * np)); CK(c, cudaMalloc(&c->d_trgb, sizeof(void*) * np));
    CK(c, cudaMallocHost(&c->h_idx, sizeof(int32_t) * 2 * np)); CK(c, cudaMalloc(&c->d_idx, sizeof(int32_t) * 2 * np));
    CK(c, cudaMallocHost(&c->h_pose, sizeof(float) * 16 * np)); CK(c, cudaMalloc(&c->d_pose, sizeof(float) * 16 * np));
    CK(c, cudaMalloc(&c->d_res, sizeof(r360_result) * np)); CK(c, cudaMallocHost(&c->h_res, sizeof(r360_result) * np));
    if (params->occlusion != 0) {
        c->occ_cap = std::min(max_pairs, kOccPairs);
        const size_t n = (size_t)c->occ_cap * npx;
        CK(c, cudaMalloc(&c->occ_head, sizeof(int) * n));
        CK(c, cudaMalloc(&c->occ_next, sizeof(int) * n));
        CK(c, cudaMemset(c->occ_next, 0xFF, sizeof(int) * n));      // k_occ_eval loads a pixel's own link before it knows whether the pixel is listed
        CK(c, cudaMalloc(&c->occ_dinv, sizeof(float) * n));
    }
    CK(c, cudaMalloc(&c->d_cams, sizeof(float) * 12 * kChunkFrames)); CK(c, cudaMallocHost(&c->h_cams, sizeof(float) * 12 * kChunkFrames));
    return R360_OK;
}

int r360_create(r360_ctx** out, int device, int rows, int cols, int max_frames, int max_pairs, const r360_params* params) {
    if (!out) return fail(nullptr, R360_E_ARG, "r360_create: null ctx pointer");
    *out = nullptr;
    r360_params def;
    if (!params) { r360_default_params(&def); params = &def; }
    if (rows <= 0 || cols <= 0 || max_frames <= 0 || max_pairs <= 0)
        return fail(nullptr, R360_E_ARG, "r360_create: rows/cols/max_frames/max_pairs must be positive");
    if (params->n_levels < 1 || params->n_levels > R360_MAX_LEVELS)
        return fail(nullptr, R360_E_ARG, "r360_create: n_levels %d not in [1,%d]", params->n_levels, R360_MAX_LEVELS);
    if ((rows % (1 << (params->n_levels - 1))) || (cols % (1 << params->n_levels)))
        return fail(nullptr, R360_E_ARG, "r360_create: %dx%d: rows must be divisible by 2^(n_levels-1) and cols by 2^n_levels "
                                         "(every level keeps an even width)", cols, rows);
    if (params->occlusion < 0 || params->occlusion > 2)
        return fail(nullptr, R360_E_ARG, "r360_create: occlusion %d not in {0, 1, 2} (RPI.h:4517)", params->occlusion);
    if (params->projection != R360_SPHERE && params->projection != R360_PINHOLE)
        return fail(nullptr, R360_E_ARG, "r360_create: projection %d is neither R360_SPHERE nor R360_PINHOLE", params->projection);
    if (params->projection == R360_PINHOLE && params->occlusion != 0)
        return fail(nullptr, R360_E_ARG, "r360_create: the pinhole occlusion variants (errorPhotoICP_Occ1/2, RPI.h:1107-2023) are not built");
    if (params->method < 0 || params->method > 2) return fail(nullptr, R360_E_ARG, "r360_create: bad method %d", params->method);
    if (params->max_iters < 1 || params->max_iters > 64) return fail(nullptr, R360_E_ARG, "r360_create: max_iters %d not in [1,64]", params->max_iters);
    if ((long long)rows * cols >= (1LL << 27)) return fail(nullptr, R360_E_ARG, "r360_create: image too large");
    // the kernels split a pixel index into (row, col) by multiplication with ceil(2^40 / cols): exact while
    // index * cols < 2^40, or for any index when cols is a power of two
    if (cols > 8192 && (cols & (cols - 1)) != 0)
        return fail(nullptr, R360_E_ARG, "r360_create: cols %d > 8192 must be a power of two", cols);
    r360_ctx* c = new r360_ctx;
    int rc = create_impl(c, device, rows, cols, max_frames, max_pairs, params);
    if (rc) { g_create_error = c->err; r360_destroy(c); return rc; }
    *out = c;
    return R360_OK;
}

int r360_set_frames(r360_ctx* c, int first, int n, const uint8_t* rgb, const uint16_t* depth_mm, const uint8_t* roles) {
    return set_frames_impl(c, first, n, rgb, depth_mm, false, false, roles);
}
int r360_set_frames_dev(r360_ctx* c, int first, int n, const uint8_t* rgb_dev, const uint16_t* depth_mm_dev, const uint8_t* roles) {
    return set_frames_impl(c, first, n, rgb_dev, depth_mm_dev, false, true, roles);
}
int r360_set_frames_f32(r360_ctx* c, int first, int n, const uint8_t* rgb, const float* depth_m, const uint8_t* roles) {
    return set_frames_impl(c, first, n, rgb, depth_m, true, false, roles);
}

// Enqueues alignFrames360 for pairs [first, first + n) of the current call on the compute stream
// (no host synchronisation): per-pair tables H2D, state init, then per level max_iters + 1 fused
// passes with the on-device Gauss-Newton step in between, results into d_res[first ..].
// The host tables (h_srcb, h_trgb, h_idx, h_pose) must already hold the whole call.
// latency_mode (small calls only; needs a host synchronisation per pass, so never inside the streaming pipeline): after
// every state-machine step the two list lengths come back to the host, and the rest of a level's schedule -- up to
// max_iters + 1 + R360_SPEC_EXTRA passes of which a converged pair needs two or three -- is not enqueued once both are 0.
static int enqueue_register(r360_ctx* c, int first, int n, int n_total, bool has_pose, r360_iter_record* d_trace,
                            bool time_passes, int* n_ev, bool latency_mode = false) {
    CK(c, cudaMemcpyAsync(c->d_srcb + first, c->h_srcb + first, sizeof(void*) * n, cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemcpyAsync(c->d_trgb + first, c->h_trgb + first, sizeof(void*) * n, cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemcpyAsync(c->d_idx + first, c->h_idx + first, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemcpyAsync(c->d_idx + c->max_pairs + 1 + first, c->h_idx + n_total + first, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->st));
    if (has_pose)
        CK(c, cudaMemcpyAsync(c->d_pose + 16 * (size_t)first, c->h_pose + 16 * (size_t)first, sizeof(float) * 16 * n, cudaMemcpyHostToDevice, c->st));
    R360GnArgs g = gn_args(c, n, d_trace, first);
    r360_launch_pairs_init(c->st, g, c->d_idx + first, c->d_idx + c->max_pairs + 1 + first,
                           has_pose ? c->d_pose + 16 * (size_t)first : nullptr);
    ++c->launches;
    for (int level = c->L - 1; level >= 0; --level) {                 // RPI.h:4531
        r360_launch_level_begin(c->st, g, level);
        ++c->launches;
        R360PassArgs a = pass_args(c, level, n, first);
        const bool pin = c->P.projection == R360_PINHOLE;
        const bool two_lists = !pin && c->P.occlusion == 0 && c->speculate;    // fused + error-only launches
        R360PassArgs a_err = pass_args(c, level, n, first, true);
        // sphere: 1 initial + <= max_iters loop bodies; pinhole: every loop body may add one damped retry
        const int n_eval = pin ? 2 * c->P.max_iters + 1 : c->P.max_iters + 1 + (two_lists ? R360_SPEC_EXTRA : 0);
        for (int k = 0; k < n_eval; ++k) {
            if (time_passes) CK(c, cudaEventRecord(c->ev_pass[(*n_ev)++], c->st));
            // The fused and the error-only launch of a pass work on disjoint pair lists: the error-only one goes to a side
            // stream, so whichever is small (a few stragglers) runs in the shadow of the other instead of after it.
            const bool side = two_lists && k > 0 && c->overlap_err && !latency_mode;      // the first pass of a level is always fused
            if (side) {
                CK(c, cudaEventRecord(c->ev_fork, c->st));
                CK(c, cudaStreamWaitEvent(c->st2, c->ev_fork, 0));
                r360_launch_pass(c->st2, a_err, c->pass_grid_err, false);
                CK(c, cudaEventRecord(c->ev_join, c->st2));
                ++c->launches;
            }
            launch_evaluation(c, a, n, level);
            if (side) {
                CK(c, cudaStreamWaitEvent(c->st, c->ev_join, 0));
            } else if (two_lists && k > 0) {
                r360_launch_pass(c->st, a_err, c->pass_grid_err, false);
                ++c->launches;
            }
            if (time_passes) CK(c, cudaEventRecord(c->ev_pass[(*n_ev)++], c->st));
            if (pin) r360_launch_gn_step_pin(c->st, g, level);
            else r360_launch_gn_step(c->st, g, level);
            ++c->launches;
            if (latency_mode) {
                CK(c, cudaMemcpyAsync(c->h_nactive, c->d_nactive, sizeof(int) * 4, cudaMemcpyDeviceToHost, c->st));
                CK(c, cudaStreamSynchronize(c->st));
                if (c->h_nactive[0] == 0 && c->h_nactive[3] == 0) break;      // every pair has left this level
            }
        }
    }
    r360_launch_finalize(c->st, g, c->d_res + first, c->rows, c->cols, first);
    ++c->launches;
    CK(c, cudaGetLastError());
    return R360_OK;
}

int r360_register_pairs(r360_ctx* c, int n_pairs, const int32_t* src_idx, const int32_t* trg_idx, const float* init_pose,
                        r360_result* out, r360_iter_record* trace) {
    if (!c) return R360_E_ARG;
    NvtxRange nvtx("r360_register_pairs");
    if (n_pairs < 0 || n_pairs > c->max_pairs || !src_idx || !trg_idx || !out)
        return fail(c, R360_E_ARG, "register_pairs: n_pairs %d not in [0,%d] or null argument", n_pairs, c->max_pairs);
    if (n_pairs == 0) return R360_OK;
    CK(c, cudaSetDevice(c->device));
    for (int p = 0; p < n_pairs; ++p) {
        int rc = check_pair(c, src_idx[p], trg_idx[p]);
        if (rc) return rc;
        c->h_srcb[p] = c->src[src_idx[p]];
        c->h_trgb[p] = c->trg[trg_idx[p]];
        c->h_idx[p] = src_idx[p];
        c->h_idx[n_pairs + p] = trg_idx[p];
    }
    if (c->P.projection == R360_PINHOLE && !c->have_cam)
        return fail(c, R360_E_STATE, "pinhole context: call r360_set_camera first (setCameraMatrix, RPI.h:254)");
    const int per_level = trace_per_level(c);
    const size_t n_rec = (size_t)n_pairs * c->L * per_level;
    if (trace && c->trace_cap < n_rec) {
        if (c->d_trace) cudaFree(c->d_trace);
        c->d_trace = nullptr; c->trace_cap = 0;
        CK(c, cudaMalloc(&c->d_trace, sizeof(r360_iter_record) * n_rec));
        c->trace_cap = n_rec;
    }
    CK(c, cudaEventRecord(c->ev_t0, c->st));
    if (init_pose) memcpy(c->h_pose, init_pose, sizeof(float) * 16 * n_pairs);
    if (trace) CK(c, cudaMemsetAsync(c->d_trace, 0, sizeof(r360_iter_record) * n_rec, c->st));
    int n_ev = 0;
    if (c->P.occlusion == 0) {
        const bool latency_mode = c->early_exit && n_pairs <= kLatencyPairs;
        int rc = enqueue_register(c, 0, n_pairs, n_pairs, init_pose != nullptr, trace ? c->d_trace : nullptr,
                                  c->P.projection == R360_SPHERE && !latency_mode, &n_ev, latency_mode);
        if (rc) return rc;
    } else {
        // the occlusion variants keep per-texel candidate lists for every pair of a batch: bounded batches
        for (int first = 0; first < n_pairs; first += c->occ_cap) {
            int rc = enqueue_register(c, first, std::min(c->occ_cap, n_pairs - first), n_pairs, init_pose != nullptr,
                                      trace ? c->d_trace : nullptr, false, &n_ev);
            if (rc) return rc;
        }
    }
    CK(c, cudaMemcpyAsync(c->h_res, c->d_res, sizeof(r360_result) * n_pairs, cudaMemcpyDeviceToHost, c->st));
    if (trace) CK(c, cudaMemcpyAsync(trace, c->d_trace, sizeof(r360_iter_record) * n_rec, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaEventRecord(c->ev_t1, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    memcpy(out, c->h_res, sizeof(r360_result) * n_pairs);
    CK(c, cudaEventElapsedTime(&c->last_ms, c->ev_t0, c->ev_t1));
    // pass statistics: device time of the fused kernel and the algorithmic bytes it moved
    // (32 B per source pixel per executed pass)
    c->pass_ms = 0.f; c->pass_launches = n_ev / 2; c->pass_bytes = 0.0;
    for (int k = 0; k < n_ev; k += 2) {
        float ms = 0.f;
        CK(c, cudaEventElapsedTime(&ms, c->ev_pass[k], c->ev_pass[k + 1]));
        c->pass_ms += ms;
    }
    for (int p = 0; p < n_pairs; ++p)
        for (int l = 0; l < c->L; ++l) c->pass_bytes += 32.0 * (double)c->lv[l].n * out[p].passes[l];
    return R360_OK;
}

// setTargetFrame + setSourceFrame + alignFrames360 for n_pairs pairs whose frames are in HOST
// memory, as one pipelined call: frame 2p is the target and frame 2p+1 the source of pair p.
// Uploads (copy stream, kStages staging buffers), pyramid builds and batched registrations of
// kStreamPairs pairs (compute stream) overlap; the host blocks once, at the end.
int r360_register_host_pairs(r360_ctx* c, int n_pairs, const uint8_t* rgb, const uint16_t* depth_mm,
                             const float* init_pose, r360_result* out) {
    if (!c) return R360_E_ARG;
    NvtxRange nvtx("r360_register_host_pairs");
    if (n_pairs < 0 || n_pairs > c->max_pairs || 2 * n_pairs > c->max_frames || !rgb || !depth_mm || !out)
        return fail(c, R360_E_ARG, "register_host_pairs: n_pairs %d needs max_pairs >= n_pairs and max_frames >= 2 n_pairs, non-null buffers", n_pairs);
    if (n_pairs == 0) return R360_OK;
    if (c->P.projection == R360_PINHOLE && !c->have_cam)
        return fail(c, R360_E_STATE, "pinhole context: call r360_set_camera first (setCameraMatrix, RPI.h:254)");
    CK(c, cudaSetDevice(c->device));
    const size_t npx = (size_t)c->rows * c->cols;
    for (int p = 0; p < n_pairs; ++p) {                               // allocate before the pipeline starts
        int rc = ensure_trg(c, 2 * p);
        if (!rc) rc = ensure_src(c, 2 * p + 1);
        if (rc) return rc;
        c->h_srcb[p] = c->src[2 * p + 1];
        c->h_trgb[p] = c->trg[2 * p];
        c->h_idx[p] = 2 * p + 1;
        c->h_idx[n_pairs + p] = 2 * p;
    }
    if (init_pose) memcpy(c->h_pose, init_pose, sizeof(float) * 16 * n_pairs);
    std::vector<uint8_t> roles(2 * (size_t)std::min(n_pairs, kStreamPairs));
    for (size_t k = 0; k < roles.size(); ++k) roles[k] = (k & 1) ? R360_ROLE_SOURCE : R360_ROLE_TARGET;
    CK(c, cudaEventRecord(c->ev_t0, c->st));
    int n_ev = 0;
    // Batches of kStreamPairs; in a call of several batches the last kStreamPairs pairs go in halves (32, 16, 8, 8): the
    // upload is the bottleneck of the call, and what the device still has to do after the last byte arrived -- the
    // pipeline's drain -- is the last batch.  A call of at most one batch stays one batch (the decomposition, and with it
    // every bit of the records, is that of r360_register_pairs).
    const bool halve_tail = n_pairs > kStreamPairs;
    for (int first = 0, nb = 0; first < n_pairs; first += nb) {
        const int left = n_pairs - first;
        nb = left > kStreamPairs ? kStreamPairs : (halve_tail && left > 8 ? (left + 1) / 2 : left);
        for (int off = 0; off < 2 * nb; off += c->chunk) {
            const int m = std::min(c->chunk, 2 * nb - off);
            const size_t f0 = 2 * (size_t)first + off;
            int rc = stage_and_build(c, (int)f0, m, rgb + f0 * npx * 3, (const uint8_t*)(depth_mm + f0 * npx), false,
                                     roles.data() + off, (int)f0);
            if (rc) return rc;
        }
        int rc = enqueue_register(c, first, nb, n_pairs, init_pose != nullptr, nullptr, false, &n_ev);
        if (rc) return rc;
    }
    CK(c, cudaMemcpyAsync(c->h_res, c->d_res, sizeof(r360_result) * n_pairs, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaEventRecord(c->ev_t1, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    memcpy(out, c->h_res, sizeof(r360_result) * n_pairs);
    CK(c, cudaEventElapsedTime(&c->last_ms, c->ev_t0, c->ev_t1));
    return R360_OK;
}

int r360_eval_error(r360_ctx* c, int src, int trg, int level, const float pose[16], double* err2, int32_t* n_valid) {
    if (!c) return R360_E_ARG;
    double acc[R360_ACC_STRIDE]; int cnt[R360_ACC_INTS];
    int rc = eval_pass(c, src, trg, level, pose, acc, cnt);
    if (rc) return rc;
    if (c->P.projection == R360_PINHOLE) {
        if (err2) *err2 = acc[27] + acc[28];
        if (n_valid) *n_valid = cnt[2];
    } else if (c->P.occlusion == 0) {
        if (err2) *err2 = acc[27];
        if (n_valid) *n_valid = cnt[1] + cnt[2];
    } else {
        if (err2) *err2 = acc[27] + acc[28];
        if (n_valid) *n_valid = c->P.occlusion == 1 ? cnt[1] + cnt[2] : cnt[2];
    }
    return R360_OK;
}

int r360_set_camera(r360_ctx* c, float fx, float fy, float ox, float oy) {
    if (!c) return R360_E_ARG;
    if (c->P.projection != R360_PINHOLE) return fail(c, R360_E_STATE, "set_camera: the context was created with projection = R360_SPHERE");
    if (!(fx > 0.f) || !(fy > 0.f)) return fail(c, R360_E_ARG, "set_camera: focal lengths must be positive");
    c->cam[0] = fx; c->cam[1] = fy; c->cam[2] = ox; c->cam[3] = oy;
    c->have_cam = true;
    return R360_OK;
}

int r360_eval_error_pinhole(r360_ctx* c, int src, int trg, int level, const float pose[16], double* photo_residual,
                            double* depth_residual, int32_t* n_valid_photo, int32_t* n_valid_depth, double* error) {
    if (!c) return R360_E_ARG;
    if (c->P.projection != R360_PINHOLE) return fail(c, R360_E_STATE, "eval_error_pinhole: the context was created with projection = R360_SPHERE");
    double acc[R360_ACC_STRIDE]; int cnt[R360_ACC_INTS];
    int rc = eval_pass(c, src, trg, level, pose, acc, cnt);
    if (rc) return rc;
    if (photo_residual) *photo_residual = acc[27];
    if (depth_residual) *depth_residual = acc[28];
    if (n_valid_photo) *n_valid_photo = cnt[1];
    if (n_valid_depth) *n_valid_depth = cnt[2];
    if (error) *error = (double)(float)(std::sqrt(acc[27] / (double)cnt[2]) + std::sqrt(acc[28] / (double)cnt[2]));   // RPI.h:768-771
    return R360_OK;
}

// ---------------------------------------------------------------- the 8-sensor rig (RegisterRGBD360::RegisterDensePhotoICP)
static int rig_check(r360_ctx* c, const float* Rt) {
    if (c->P.projection != R360_PINHOLE) return fail(c, R360_E_STATE, "rig: the context was created with projection = R360_SPHERE");
    if (!c->have_cam) return fail(c, R360_E_STATE, "rig: call r360_set_camera first (RegisterRGBD360.h:361-369)");
    if (c->P.method != R360_PHOTO_CONSISTENCY)
        return fail(c, R360_E_STATE, "rig: only PHOTO_CONSISTENCY is defined (calcHessianGradient_robot's depth row reads a matrix that "
                                     "is never assigned upstream, RPI.h:5372-5374)");
    if (c->P.occlusion != 0) return fail(c, R360_E_STATE, "rig: occlusion must be 0");
    if (!Rt) return fail(c, R360_E_ARG, "rig: null extrinsics");
    if (!c->d_rig_src) {
        const size_t n = 8 * (size_t)(c->max_pairs + 1);
        CK(c, cudaMallocHost(&c->h_rig_src, sizeof(void*) * n)); CK(c, cudaMalloc(&c->d_rig_src, sizeof(void*) * n));
        CK(c, cudaMallocHost(&c->h_rig_trg, sizeof(void*) * n)); CK(c, cudaMalloc(&c->d_rig_trg, sizeof(void*) * n));
    }
    return R360_OK;
}
static int rig_frames(r360_ctx* c, int slot, int src_first, int trg_first) {
    if (src_first < 0 || src_first + 8 > c->max_frames || trg_first < 0 || trg_first + 8 > c->max_frames)
        return fail(c, R360_E_ARG, "rig: a rig frame occupies 8 consecutive slots (src %d, trg %d, %d slots)", src_first, trg_first, c->max_frames);
    for (int s = 0; s < 8; ++s) {
        int rc = check_pair(c, src_first + s, trg_first + s);
        if (rc) return rc;
        c->h_rig_src[8 * (size_t)slot + s] = c->src[src_first + s];
        c->h_rig_trg[8 * (size_t)slot + s] = c->trg[trg_first + s];
    }
    return R360_OK;
}
static R360RigArgs rig_args(const r360_ctx* c, const float* Rt, int level) {
    R360RigArgs r;
    memcpy(r.Rt, Rt, sizeof(r.Rt));
    for (int s = 0; s < 8; ++s) r360_inverse4(r.Rt[s], r.Rt_inv[s]);          // poseCamRobot.inverse(), RPI.h:4923 / 5125
    r.src8 = c->d_rig_src; r.trg8 = c->d_rig_trg;
    const R360PinLevel pl = pin_level(c, level);                              // RPI.h:4915-4921 (float)
    r.fx = pl.fx; r.fy = pl.fy; r.ox = pl.ox; r.oy = pl.oy; r.inv_fx = pl.inv_fx; r.inv_fy = pl.inv_fy;
    const double scaleFactor = 1.0 / pow(2, level);                           // RPI.h:5108-5114 (double)
    r.dfx = c->cam[0] * scaleFactor; r.dfy = c->cam[1] * scaleFactor; r.dox = c->cam[2] * scaleFactor; r.doy = c->cam[3] * scaleFactor;
    r.dinv_fx = 1. / r.dfx; r.dinv_fy = 1. / r.dfy;
    return r;
}

int r360_eval_rig(r360_ctx* c, int src_first, int trg_first, int level, const float pose[16], const float* Rt,
                  double* error2, float H[36], float g[6], int32_t* n_visible, int32_t* n_error_terms) {
    if (!c) return R360_E_ARG;
    int rc = rig_check(c, Rt);
    if (rc) return rc;
    const int slot = c->max_pairs;
    rc = rig_frames(c, slot, src_first, trg_first);
    if (rc) return rc;
    R360PassArgs a;
    rc = eval_setup(c, src_first, trg_first, level, pose, &a);
    if (rc) return rc;
    CK(c, cudaMemcpyAsync(c->d_rig_src + 8 * (size_t)slot, c->h_rig_src + 8 * (size_t)slot, sizeof(void*) * 8, cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemcpyAsync(c->d_rig_trg + 8 * (size_t)slot, c->h_rig_trg + 8 * (size_t)slot, sizeof(void*) * 8, cudaMemcpyHostToDevice, c->st));
    const R360RigArgs rig = rig_args(c, Rt, level);
    // (the frame tables are indexed by the pair id, here the spare slot)
    r360_launch_rig_eval(c->st, a, rig, 1, c->sm_count);
    ++c->launches;
    CK(c, cudaGetLastError());
    R360Fx fx[R360_ACC_STRIDE];
    int cnt[R360_ACC_INTS];
    CK(c, cudaMemcpyAsync(fx, c->d_acc + (size_t)slot * R360_ACC_STRIDE, sizeof(fx), cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaMemcpyAsync(cnt, c->d_cnt + (size_t)slot * R360_ACC_INTS, sizeof(cnt), cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    if (error2) *error2 = r360_fx_get(fx, 27);
    if (H) {
        int q = 0;
        for (int x = 0; x < 6; ++x)
            for (int y = x; y < 6; ++y, ++q) H[x + 6 * y] = H[y + 6 * x] = (float)r360_fx_get(fx, q);
    }
    if (g) for (int x = 0; x < 6; ++x) g[x] = (float)r360_fx_get(fx, 21 + x);
    if (n_visible) *n_visible = cnt[0];
    if (n_error_terms) *n_error_terms = cnt[1];
    return R360_OK;
}

int r360_register_rig_pairs(r360_ctx* c, int n_pairs, const int32_t* src_first, const int32_t* trg_first, const float* Rt,
                            const float* init_pose, int faithful_new_error, r360_result* out) {
    if (!c) return R360_E_ARG;
    NvtxRange nvtx("r360_register_rig_pairs");
    if (n_pairs < 0 || n_pairs > c->max_pairs || !src_first || !trg_first || !out)
        return fail(c, R360_E_ARG, "register_rig_pairs: n_pairs %d not in [0,%d] or null argument", n_pairs, c->max_pairs);
    if (n_pairs == 0) return R360_OK;
    int rc = rig_check(c, Rt);
    if (rc) return rc;
    CK(c, cudaSetDevice(c->device));
    for (int p = 0; p < n_pairs; ++p) {
        rc = rig_frames(c, p, src_first[p], trg_first[p]);
        if (rc) return rc;
        c->h_srcb[p] = c->src[src_first[p]]; c->h_trgb[p] = c->trg[trg_first[p]];
        c->h_idx[p] = src_first[p]; c->h_idx[n_pairs + p] = trg_first[p];
    }
    if (init_pose) memcpy(c->h_pose, init_pose, sizeof(float) * 16 * n_pairs);
    CK(c, cudaEventRecord(c->ev_t0, c->st));
    CK(c, cudaMemcpyAsync(c->d_rig_src, c->h_rig_src, sizeof(void*) * 8 * n_pairs, cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemcpyAsync(c->d_rig_trg, c->h_rig_trg, sizeof(void*) * 8 * n_pairs, cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemcpyAsync(c->d_idx, c->h_idx, sizeof(int32_t) * n_pairs, cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemcpyAsync(c->d_idx + c->max_pairs + 1, c->h_idx + n_pairs, sizeof(int32_t) * n_pairs, cudaMemcpyHostToDevice, c->st));
    if (init_pose) CK(c, cudaMemcpyAsync(c->d_pose, c->h_pose, sizeof(float) * 16 * n_pairs, cudaMemcpyHostToDevice, c->st));
    R360GnArgs g = gn_args(c, n_pairs, nullptr, 0);
    g.params.max_iters = 10;                                   // RegisterRGBD360.h:395-397
    g.params.tol_residual = pow(10, -1);
    g.params.tol_update = pow(10, -6);
    g.lambda0 = 0.001;                                         // :391
    g.rig_faithful = faithful_new_error ? 1 : 0;
    r360_launch_pairs_init(c->st, g, c->d_idx, c->d_idx + c->max_pairs + 1, init_pose ? c->d_pose : nullptr);
    ++c->launches;
    for (int level = c->L - 1; level >= 0; --level) {
        r360_launch_level_begin(c->st, g, level);
        ++c->launches;
        R360PassArgs a = pass_args(c, level, n_pairs, 0);
        const R360RigArgs rig = rig_args(c, Rt, level);
        const int n_eval = 2 * g.params.max_iters + 1;         // one initial evaluation; every loop body may add one damped retry
        const bool latency_mode = c->early_exit && n_pairs <= kLatencyPairs;     // as r360_register_pairs: stop a level once every pair has left it
        for (int k = 0; k < n_eval; ++k) {
            r360_launch_rig_eval(c->st, a, rig, n_pairs, c->sm_count);
            r360_launch_gn_step_rig(c->st, g, level);
            c->launches += 2;
            if (latency_mode) {
                CK(c, cudaMemcpyAsync(c->h_nactive, c->d_nactive, sizeof(int) * 4, cudaMemcpyDeviceToHost, c->st));
                CK(c, cudaStreamSynchronize(c->st));
                if (c->h_nactive[0] == 0 && c->h_nactive[3] == 0) break;
            }
        }
    }
    r360_launch_finalize(c->st, g, c->d_res, c->rows, c->cols, 0);
    ++c->launches;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(c->h_res, c->d_res, sizeof(r360_result) * n_pairs, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaEventRecord(c->ev_t1, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    memcpy(out, c->h_res, sizeof(r360_result) * n_pairs);
    CK(c, cudaEventElapsedTime(&c->last_ms, c->ev_t0, c->ev_t1));
    return R360_OK;
}

int r360_eval_error_occ(r360_ctx* c, int src, int trg, int level, const float pose[16], double* photo_residual,
                        double* depth_residual, int32_t* n_valid_photo, int32_t* n_valid_depth, double* error) {
    if (!c) return R360_E_ARG;
    if (c->P.occlusion == 0) return fail(c, R360_E_STATE, "eval_error_occ: the context was created with occlusion = 0");
    double acc[R360_ACC_STRIDE]; int cnt[R360_ACC_INTS];
    int rc = eval_pass(c, src, trg, level, pose, acc, cnt);
    if (rc) return rc;
    if (photo_residual) *photo_residual = acc[27];
    if (depth_residual) *depth_residual = acc[28];
    if (n_valid_photo) *n_valid_photo = cnt[1];
    if (n_valid_depth) *n_valid_depth = cnt[2];
    if (error) {
        // RPI.h:3360-3367 (Occ1: each RMS over its own counter) / 3849-3856 (Occ2: both over nValidDepthPts)
        const double n_p = (double)(c->P.occlusion == 1 ? cnt[1] : cnt[2]), n_d = (double)cnt[2];
        *error = std::sqrt(acc[27] / n_p) + std::sqrt(acc[28] / n_d);
    }
    return R360_OK;
}

int r360_eval_hessgrad(r360_ctx* c, int src, int trg, int level, const float pose[16], float H[36], float g[6], int32_t* n_visible) {
    if (!c) return R360_E_ARG;
    double acc[R360_ACC_STRIDE]; int cnt[R360_ACC_INTS];
    int rc = eval_pass(c, src, trg, level, pose, acc, cnt);
    if (rc) return rc;
    if (H) {
        int q = 0;
        for (int a = 0; a < 6; ++a)
            for (int b = a; b < 6; ++b, ++q) H[a + 6 * b] = H[b + 6 * a] = (float)acc[q];
    }
    if (g) for (int a = 0; a < 6; ++a) g[a] = (float)acc[21 + a];
    if (n_visible) *n_visible = cnt[0];
    return R360_OK;
}

int r360_dump_level(r360_ctx* c, int frame, int level, float* gray, float* depth, float* ggx, float* ggy, float* dgx, float* dgy) {
    if (!c) return R360_E_ARG;
    if (frame < 0 || frame >= c->max_frames || level < 0 || level >= c->L) return fail(c, R360_E_ARG, "dump_level: bad frame %d / level %d", frame, level);
    CK(c, cudaSetDevice(c->device));
    const R360Level& v = c->lv[level];
    const bool want_grad = ggx || ggy || dgx || dgy;
    if (want_grad && !c->have_trg[frame]) return fail(c, R360_E_STATE, "dump_level: frame %d has no target pyramid", frame);
    if (c->have_trg[frame]) {
        std::vector<float> buf((size_t)v.n * R360_TEXEL_FLOATS);
        CK(c, cudaMemcpy(buf.data(), c->trg[frame] + v.px_off * R360_TEXEL_FLOATS, buf.size() * sizeof(float), cudaMemcpyDeviceToHost));
        for (int i = 0; i < v.n; ++i) {
            const float* t = &buf[(size_t)i * R360_TEXEL_FLOATS];
            if (gray) gray[i] = t[0];
            if (depth) depth[i] = t[1];
            if (ggx) ggx[i] = t[2];
            if (ggy) ggy[i] = t[3];
            if (dgx) dgx[i] = t[4];
            if (dgy) dgy[i] = t[5];
        }
    } else if (c->have_src[frame]) {
        std::vector<float2> buf(v.n);
        CK(c, cudaMemcpy(buf.data(), c->src[frame] + v.px_off, buf.size() * sizeof(float2), cudaMemcpyDeviceToHost));
        for (int i = 0; i < v.n; ++i) { if (depth) depth[i] = buf[i].x; if (gray) gray[i] = buf[i].y; }
    } else {
        return fail(c, R360_E_STATE, "dump_level: frame %d was never set", frame);
    }
    return R360_OK;
}

// Source-role planes of a frame (the {depth, gray} pyramid the warp reads).
int r360_dump_source_level(r360_ctx* c, int frame, int level, float* gray, float* depth) {
    if (!c) return R360_E_ARG;
    if (frame < 0 || frame >= c->max_frames || level < 0 || level >= c->L) return fail(c, R360_E_ARG, "dump_source_level: bad frame/level");
    if (!c->have_src[frame]) return fail(c, R360_E_STATE, "dump_source_level: frame %d has no source pyramid", frame);
    CK(c, cudaSetDevice(c->device));
    const R360Level& v = c->lv[level];
    std::vector<float2> buf(v.n);
    CK(c, cudaMemcpy(buf.data(), c->src[frame] + v.px_off, buf.size() * sizeof(float2), cudaMemcpyDeviceToHost));
    for (int i = 0; i < v.n; ++i) { if (depth) depth[i] = buf[i].x; if (gray) gray[i] = buf[i].y; }
    return R360_OK;
}

int r360_dump_warp(r360_ctx* c, int src, int trg, int level, const float pose[16], int32_t* r_idx, int32_t* c_idx,
                   uint8_t* valid_photo, uint8_t* valid_depth) {
    if (!c) return R360_E_ARG;
    R360PassArgs a;
    int rc = eval_setup(c, src, trg, level, pose, &a);
    if (rc) return rc;
    const int n = c->lv[level].n;
    int32_t* d_r = nullptr; int32_t* d_c = nullptr; uint8_t* d_vp = nullptr; uint8_t* d_vd = nullptr;
    CK(c, cudaMalloc(&d_r, sizeof(int32_t) * n)); CK(c, cudaMalloc(&d_c, sizeof(int32_t) * n));
    CK(c, cudaMalloc(&d_vp, n)); CK(c, cudaMalloc(&d_vd, n));
    r360_launch_warp_dump(c->st, a, c->max_pairs, d_r, d_c, d_vp, d_vd, c->sm_count);
    ++c->launches;
    cudaError_t e = cudaStreamSynchronize(c->st);
    if (e == cudaSuccess && r_idx) e = cudaMemcpy(r_idx, d_r, sizeof(int32_t) * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && c_idx) e = cudaMemcpy(c_idx, d_c, sizeof(int32_t) * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && valid_photo) e = cudaMemcpy(valid_photo, d_vp, n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && valid_depth) e = cudaMemcpy(valid_depth, d_vd, n, cudaMemcpyDeviceToHost);
    cudaFree(d_r); cudaFree(d_c); cudaFree(d_vp); cudaFree(d_vd);
    if (e != cudaSuccess) return fail(c, R360_E_CUDA, "dump_warp: %s", cudaGetErrorString(e));
    return R360_OK;
}

int r360_index_stats(r360_ctx* c, int src, int trg, int level, const float pose[16], uint64_t out[3]) {
    if (!c || !out) return R360_E_ARG;
    R360PassArgs a;
    int rc = eval_setup(c, src, trg, level, pose, &a);
    if (rc) return rc;
    unsigned long long* d_out = nullptr;
    CK(c, cudaMalloc(&d_out, 4 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d_out, 0, 4 * sizeof(unsigned long long), c->st);
    r360_launch_index_stats(c->st, a, c->max_pairs, d_out, c->sm_count);
    ++c->launches;
    unsigned long long h[4] = { 0, 0, 0, 0 };
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, c->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(c, R360_E_CUDA, "index_stats: %s", cudaGetErrorString(e));
    for (int k = 0; k < 3; ++k) out[k] = h[k];
    return R360_OK;
}

int r360_synth_frames_dev(r360_ctx* c, int kind, int first_id, int n, uint8_t* rgb_dev, uint16_t* depth_mm_dev) {
    if (!c) return R360_E_ARG;
    if (n < 0 || !rgb_dev || !depth_mm_dev || kind < 0 || kind > 1) return fail(c, R360_E_ARG, "synth_frames: bad arguments");
    CK(c, cudaSetDevice(c->device));
    const size_t npx = (size_t)c->rows * c->cols;
    for (int off = 0; off < n; off += kChunkFrames) {
        const int m = std::min(kChunkFrames, n - off);
        CK(c, cudaStreamSynchronize(c->st));       // h_cams reuse
        for (int k = 0; k < m; ++k) {
            double R[9], t[3];
            r360_synth_pose(kind, first_id + off + k, R, t);
            for (int q = 0; q < 9; ++q) c->h_cams[12 * k + q] = (float)R[q];
            for (int q = 0; q < 3; ++q) c->h_cams[12 * k + 9 + q] = (float)t[q];
        }
        CK(c, cudaMemcpyAsync(c->d_cams, c->h_cams, sizeof(float) * 12 * m, cudaMemcpyHostToDevice, c->st));
        r360_launch_synth(c->st, kind, first_id + off, c->rows, c->cols, c->d_cams, m, rgb_dev + (size_t)off * npx * 3,
                          depth_mm_dev + (size_t)off * npx, c->sm_count);
        ++c->launches;
    }
    CK(c, cudaGetLastError());
    CK(c, cudaStreamSynchronize(c->st));
    return R360_OK;
}

int r360_synth_frames(r360_ctx* c, int kind, int first_id, int n, uint8_t* rgb, uint16_t* depth_mm) {
    if (!c) return R360_E_ARG;
    if (n < 0 || !rgb || !depth_mm) return fail(c, R360_E_ARG, "synth_frames: bad arguments");
    const size_t npx = (size_t)c->rows * c->cols;
    for (int off = 0; off < n; off += c->chunk) {
        const int m = std::min(c->chunk, n - off);
        int rc = r360_synth_frames_dev(c, kind, first_id + off, m, c->stage_rgb[0], c->stage_depth[0]);
        if (rc) return rc;
        CK(c, cudaMemcpy(rgb + (size_t)off * npx * 3, c->stage_rgb[0], (size_t)m * npx * 3, cudaMemcpyDeviceToHost));
        CK(c, cudaMemcpy(depth_mm + (size_t)off * npx, c->stage_depth[0], (size_t)m * npx * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    }
    return R360_OK;
}

// ---------------------------------------------------------------- Frame360 ingest
void r360_default_rig(r360_rig* rig) {
    if (!rig) return;
    memset(rig, 0, sizeof(*rig));
    rig->sensor_rows = 240; rig->sensor_cols = 320;
    rig->fx = 262.5f; rig->fy = 262.5f; rig->cx = 159.5f; rig->cy = 119.5f;          // Calib360.h:75-77
    for (int s = 0; s < 8; ++s)
        for (int k = 0; k < 4; ++k) rig->Rt_inv[s][5 * k] = 1.0f;
}

// boost::archive::binary_oarchive layout as written by the reference's grabber (64-bit little endian
// build): 4-byte length-prefixed signature "serialization::archive" with its version and flags (45
// bytes in all up to the first object), then per cv::Mat {int32 cols, int32 rows, uint64 elemSize,
// uint64 type, raw pixels} (cvmat_serialization.h:22-37).
int r360_frame360_parse(const uint8_t* b, size_t n_bytes, int32_t* sensor_rows, int32_t* sensor_cols,
                        uint8_t* rgb, size_t rgb_capacity, uint16_t* depth_mm, size_t depth_capacity) {
    static const char sig[] = "serialization::archive";
    if (!b || n_bytes < 45 + 24) return R360_E_ARG;
    if (memcmp(b + 8, sig, sizeof(sig) - 1) != 0) return R360_E_ARG;
    size_t off = 45;
    int rows0 = 0, cols0 = 0;
    for (int k = 0; k < 16; ++k) {
        if (off + 24 > n_bytes) return R360_E_ARG;
        int32_t cols, rows;
        uint64_t es, ty;
        memcpy(&cols, b + off, 4); memcpy(&rows, b + off + 4, 4);
        memcpy(&es, b + off + 8, 8); memcpy(&ty, b + off + 16, 8);
        off += 24;
        const bool is_rgb = (k % 2) == 0;
        if (cols <= 0 || rows <= 0 || cols > 4096 || rows > 4096) return R360_E_ARG;
        if (is_rgb ? (es != 3 || ty != 16) : (es != 2 || ty != 2)) return R360_E_ARG;     // CV_8UC3 / CV_16UC1
        if (k == 0) { rows0 = rows; cols0 = cols; }
        else if (rows != rows0 || cols != cols0) return R360_E_ARG;
        const size_t nb = (size_t)rows * cols * es;
        if (off + nb > n_bytes) return R360_E_ARG;
        const size_t s = (size_t)(k / 2);
        if (is_rgb && rgb) {
            if ((s + 1) * nb > rgb_capacity) return R360_E_ARG;
            memcpy(rgb + s * nb, b + off, nb);
        }
        if (!is_rgb && depth_mm) {
            if ((s + 1) * nb > depth_capacity * sizeof(uint16_t)) return R360_E_ARG;
            memcpy((uint8_t*)depth_mm + s * nb, b + off, nb);
        }
        off += nb;
    }
    if (sensor_rows) *sensor_rows = rows0;
    if (sensor_cols) *sensor_cols = cols0;
    return R360_OK;
}

int r360_stitch_frames(r360_ctx* c, const r360_rig* rig, int first, int n, const uint8_t* sensor_rgb,
                       const uint16_t* sensor_depth_mm, const uint8_t* roles, uint8_t* sphere_rgb,
                       uint16_t* sphere_depth_mm) {
    if (!c) return R360_E_ARG;
    if (!rig || !sensor_rgb || !sensor_depth_mm || n < 0 || first < 0 || first + n > c->max_frames)
        return fail(c, R360_E_ARG, "stitch_frames: bad range [%d,%d) of %d slots or null input", first, first + n, c->max_frames);
    R360StitchArgs a;
    a.g = r360_stitch_geom(rig->sensor_rows, rig->sensor_cols, rig->fx, rig->fy, rig->cx, rig->cy);
    if (a.g.rows != c->rows || a.g.cols != c->cols)
        return fail(c, R360_E_ARG, "stitch_frames: rig gives a %dx%d sphere, the context holds %dx%d", a.g.cols, a.g.rows, c->cols, c->rows);
    if (roles)
        for (int k = 0; k < n; ++k)
            if (roles[k] < 1 || roles[k] > 3) return fail(c, R360_E_ARG, "stitch_frames: role %d of frame %d invalid", roles[k], k);
    memcpy(a.Rt_inv, rig->Rt_inv, sizeof(a.Rt_inv));
    CK(c, cudaSetDevice(c->device));
    const size_t spx = (size_t)rig->sensor_rows * rig->sensor_cols * 8, npx = (size_t)c->rows * c->cols;
    if (c->sens_cap < spx * c->chunk) {
        cudaFree(c->d_sens_rgb); cudaFree(c->d_sens_depth);
        c->d_sens_rgb = nullptr; c->d_sens_depth = nullptr; c->sens_cap = 0;
        CK(c, cudaMalloc(&c->d_sens_rgb, spx * 3 * c->chunk));
        CK(c, cudaMalloc(&c->d_sens_depth, spx * sizeof(uint16_t) * c->chunk));
        c->sens_cap = spx * c->chunk;
    }
    CK(c, cudaEventRecord(c->ev_t0, c->st));
    for (int off = 0; off < n; off += c->chunk) {
        const int m = std::min(c->chunk, n - off);
        const int b = (int)(c->n_chunks_done % kStages);
        CK(c, cudaMemcpyAsync(c->d_sens_rgb, sensor_rgb + (size_t)off * spx * 3, (size_t)m * spx * 3, cudaMemcpyHostToDevice, c->st));
        CK(c, cudaMemcpyAsync(c->d_sens_depth, sensor_depth_mm + (size_t)off * spx, (size_t)m * spx * sizeof(uint16_t), cudaMemcpyHostToDevice, c->st));
        r360_launch_stitch(c->st, a, c->d_sens_rgb, c->d_sens_depth, c->stage_rgb[b], c->stage_depth[b], m, c->sm_count);
        ++c->launches;
        int rc = build_chunk(c, first + off, m, c->stage_rgb[b], c->stage_depth[b], nullptr, roles ? roles + off : nullptr, off);
        if (rc) return rc;
        if (sphere_rgb) CK(c, cudaMemcpyAsync(sphere_rgb + (size_t)off * npx * 3, c->stage_rgb[b], (size_t)m * npx * 3, cudaMemcpyDeviceToHost, c->st));
        if (sphere_depth_mm) CK(c, cudaMemcpyAsync(sphere_depth_mm + (size_t)off * npx, c->stage_depth[b], (size_t)m * npx * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->st));
        CK(c, cudaEventRecord(c->ev_done[b], c->st));
        ++c->n_chunks_done;
        CK(c, cudaStreamSynchronize(c->st));             // the sensor buffer is reused by the next chunk
    }
    CK(c, cudaEventRecord(c->ev_t1, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    CK(c, cudaEventElapsedTime(&c->last_ms, c->ev_t0, c->ev_t1));
    return R360_OK;
}

void r360_synth_gt_pose(int kind, int src_id, int trg_id, double T[16]) { r360_synth_relpose(kind, src_id, trg_id, T); }

int r360_allgather_results(r360_ctx* c, void* nccl_comm, const r360_result* local, int n_local, int n_ranks, r360_result* all) {
    if (!c) return R360_E_ARG;
    NvtxRange nvtx("r360_allgather_results");
    if (!nccl_comm || n_local < 0 || n_ranks < 1 || (n_local > 0 && (!local || !all)))
        return fail(c, R360_E_ARG, "allgather_results: null communicator / buffer or bad counts (n_local %d, n_ranks %d)", n_local, n_ranks);
    if (n_local == 0) return R360_OK;
    const NcclApi& nccl = nccl_api();
    if (!nccl.all_gather) return fail(c, R360_E_STATE, "allgather_results: no NCCL in this process and libnccl.so.2 not found");
    CK(c, cudaSetDevice(c->device));
    const size_t bytes = sizeof(r360_result) * (size_t)n_local;
    const size_t need = bytes * ((size_t)n_ranks + 1);
    if (c->gather_cap < need) {
        if (c->d_gather) cudaFree(c->d_gather);
        c->d_gather = nullptr; c->gather_cap = 0;
        CK(c, cudaMalloc(&c->d_gather, need));
        c->gather_cap = need;
    }
    uint8_t* d_in = c->d_gather;
    uint8_t* d_out = c->d_gather + bytes;
    CK(c, cudaMemcpyAsync(d_in, local, bytes, cudaMemcpyHostToDevice, c->st));
    const int rc = nccl.all_gather(d_in, d_out, bytes, /*ncclUint8*/ 1, nccl_comm, c->st);
    if (rc != 0)
        return fail(c, R360_E_CUDA, "ncclAllGather failed: %s", nccl.error_string ? nccl.error_string(rc) : "unknown NCCL error");
    CK(c, cudaMemcpyAsync(all, d_out, bytes * (size_t)n_ranks, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    return R360_OK;
}

// NUMA node of the ctx's GPU from sysfs (-1: not exposed), and the CPUs of that node.
static int gpu_numa_node(r360_ctx* c, std::string* why) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), c->device) != cudaSuccess) { *why = "cudaDeviceGetPCIBusId failed"; return -1; }
    for (char* p = bus; *p; ++p) *p = (char)tolower(*p);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) { *why = path + " not readable"; return -1; }
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    if (node < 0) *why = path + " = -1 (the platform exposes no NUMA node for this device)";
    return node;
}
static bool node_cpus(int node, cpu_set_t* set) {
    char path[96];
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    CPU_ZERO(set);
    int a, b, n = 0;
    while (fscanf(f, "%d", &a) == 1) {
        b = a;
        int ch = fgetc(f);
        if (ch == '-') { if (fscanf(f, "%d", &b) != 1) break; ch = fgetc(f); }
        for (int k = a; k <= b && k < CPU_SETSIZE; ++k) { CPU_SET(k, set); ++n; }
        if (ch != ',') break;
    }
    fclose(f);
    return n > 0;
}

int r360_host_alloc(r360_ctx* c, size_t bytes, void** ptr, int32_t* numa_node) {
    if (!c || !ptr || bytes == 0) return R360_E_ARG;
    *ptr = nullptr;
    if (numa_node) *numa_node = -1;
    CK(c, cudaSetDevice(c->device));
    std::string why;
    const int node = gpu_numa_node(c, &why);
    if (node < 0) {
        CK(c, cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
        c->host_allocs[*ptr] = std::make_pair(bytes, false);
        c->err = "host_alloc: no NUMA placement: " + why;
        return R360_OK;
    }
    const size_t page = 2u << 20;
    const size_t len = (bytes + page - 1) / page * page;
    void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return fail(c, R360_E_NOMEM, "host_alloc: mmap of %zu bytes failed", len);
    // bind the range to the node (needs no privilege for the caller's own memory; seccomp profiles may refuse it) ...
    unsigned long mask[16] = {0};
    bool bound = false;
    if (node < (int)(sizeof(mask) * 8)) {
        mask[node / (8 * sizeof(unsigned long))] |= 1UL << (node % (8 * sizeof(unsigned long)));
        bound = syscall(SYS_mbind, p, len, /*MPOL_BIND*/ 2, mask, sizeof(mask) * 8, 0) == 0;
    }
    // ... and first-touch the pages from a CPU of that node (the default policy then places them there)
    cpu_set_t old_set, node_set;
    const bool have_old = sched_getaffinity(0, sizeof(old_set), &old_set) == 0;
    bool pinned = false;
    if (have_old && node_cpus(node, &node_set)) {
        cpu_set_t both;
        CPU_AND(&both, &old_set, &node_set);
        if (CPU_COUNT(&both) > 0) pinned = sched_setaffinity(0, sizeof(both), &both) == 0;
    }
    for (size_t off = 0; off < len; off += 4096) ((volatile char*)p)[off] = 0;
    if (pinned) sched_setaffinity(0, sizeof(old_set), &old_set);
    cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterPortable);
    if (e != cudaSuccess) { munmap(p, len); return fail(c, R360_E_CUDA, "host_alloc: cudaHostRegister: %s", cudaGetErrorString(e)); }
    int where = -1;
    if (syscall(SYS_get_mempolicy, &where, nullptr, 0UL, p, /*MPOL_F_NODE | MPOL_F_ADDR*/ 3UL) != 0) where = -1;
    if (numa_node) *numa_node = where;
    char note[256];
    snprintf(note, sizeof(note), "host_alloc: GPU %d on NUMA node %d; mbind %s, first touch %s; first page on node %d", c->device, node,
             bound ? "ok" : "refused", pinned ? "from a CPU of the node" : "from the caller's CPU (affinity excludes the node)", where);
    c->err = note;
    c->host_allocs[p] = std::make_pair(len, true);
    *ptr = p;
    return R360_OK;
}

int r360_host_free(r360_ctx* c, void* p) {
    if (!c) return R360_E_ARG;
    if (!p) return R360_OK;
    auto it = c->host_allocs.find(p);
    if (it == c->host_allocs.end()) return fail(c, R360_E_ARG, "host_free: pointer was not returned by r360_host_alloc of this ctx");
    CK(c, cudaSetDevice(c->device));
    if (it->second.second) { cudaHostUnregister(p); munmap(p, it->second.first); }
    else cudaFreeHost(p);
    c->host_allocs.erase(it);
    return R360_OK;
}

int r360_device_alloc(r360_ctx* c, size_t bytes, void** ptr_dev) {
    if (!c || !ptr_dev) return R360_E_ARG;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaMalloc(ptr_dev, bytes));
    return R360_OK;
}
int r360_device_free(r360_ctx* c, void* ptr_dev) {
    if (!c) return R360_E_ARG;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaFree(ptr_dev));
    return R360_OK;
}
int r360_synchronize(r360_ctx* c) {
    if (!c) return R360_E_ARG;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->cs));
    CK(c, cudaStreamSynchronize(c->st));
    return R360_OK;
}
float r360_last_device_ms(const r360_ctx* c) { return c ? c->last_ms : 0.f; }
int64_t r360_kernel_launches(const r360_ctx* c) { return c ? c->launches : 0; }
int r360_last_pass_stats(const r360_ctx* c, float* total_ms, int32_t* launches, double* alg_bytes) {
    if (!c) return R360_E_ARG;
    if (total_ms) *total_ms = c->pass_ms;
    if (launches) *launches = c->pass_launches;
    if (alg_bytes) *alg_bytes = c->pass_bytes;
    return R360_OK;
}

}  // extern "C"
