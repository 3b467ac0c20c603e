// context.cuh -- host-side state behind the C ABI: one mlvb_context per GPU holding the EV tables,
// a ring of frame slots (stream + device buffers + pinned staging), a scratch arena for batches and
// the per-clip state the reference keeps in file-static lists.
#pragma once
#include <atomic>
#include <condition_variable>
#include <chrono>
#include <deque>
#include <set>
#include <thread>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mlvfs_b200.h"
#include "kernels.cuh"

// glibc-compatible rand() (TYPE_3 additive feedback generator), the dither source of stripes.c:129-130.
struct GlibcRand {
    uint32_t r[34];
    int f = 3, b = 0;
    GlibcRand() { seed(1); }
    void seed(unsigned s);
    int next();
};

struct FrameGeom;
// A pixel list with its level schedule, resident on the device (bad-pixel map, focus-pixel map).
struct PixelList {
    std::vector<PixelXY> host;             // original (reference) order
    PixelXY *d_by_level = nullptr;         // sorted by level, stable
    unsigned *d_level_start = nullptr;
    // dual-ISO (horizontal-only) application: entries grouped by row in list order and cut into segments
    // that cannot see each other (x gap > 3); one thread walks one segment sequentially
    PixelXY *d_by_row = nullptr;
    unsigned *d_seg_start = nullptr;       // segments of the sparse rows: [2 * s], [2 * s + 1] = first, end entry
    unsigned nseg = 0;
    unsigned *d_long_rows = nullptr;       // rows with many entries (one warp walks the whole row in shared memory)
    unsigned nlong = 0;
    std::vector<unsigned> level_start;     // nlevels + 1 entries
    unsigned nlevels = 0;
    // focus maps only: some entry lies outside [0, w) in x but still acts, on the wrapped linear index
    // i = x + y * w (cs.c:467, 479-500).  Such a write leaves its row: the row-wise dual-ISO walk is not used then.
    bool has_wrapped = false;
    ~PixelList();
    // Builds the level schedule from `host`.  Without a geometry (bad-pixel maps: interior entries only) the
    // +-3 cross stencil in map coordinates; with one (focus maps, whose border rules depend on where an entry
    // falls in the frame) entries that do nothing in this frame are dropped and levels come from linear indices.
    int upload(const struct FrameGeom *focus_geom = nullptr);
};

struct BadPixelMap {                       // reference cs.c:186-193 + the 8-slot ring cs.c:215-217
    uint64_t file_guid = 0;
    int aggressive = 0;
    bool valid = false;
    std::shared_ptr<PixelList> list;
    std::shared_ptr<void> fused_plan;      // fused.cu: per-warp patch buckets for this map (built lazily)
};

struct FocusPixelMap {                     // reference cs.c:176-184
    uint32_t camera = 0;
    int rawi_width = 0, rawi_height = 0;
    std::vector<PixelXY> entries;          // file order; empty: no map file for this camera/size
    // one schedule per frame geometry the map has been applied to (the border rules of cs.c:479-500 and the
    // wrapped linear indices depend on width, height and crop offsets)
    struct Schedule { int w, h, crop_x, crop_y; std::shared_ptr<PixelList> list; };
    std::vector<Schedule> schedules;
};

struct StripesState {                      // reference stripes.h:30-36, keyed by MLV path (stripes.c:29-38)
    bool computed = false;
    StripeCoef coef{};
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    uint8_t *d_packed = nullptr;  size_t packed_cap = 0;
    uint16_t *d_a = nullptr, *d_b = nullptr;  size_t frame_cap = 0;     // bytes each
    uint8_t *h_in = nullptr;   size_t h_in_cap = 0;                    // pinned staging
    uint16_t *h_out = nullptr; size_t h_out_cap = 0;
    void *d_aux = nullptr; size_t aux_cap = 0;                          // per-slot scratch (pattern noise, dual ISO)
    int *d_status = nullptr, *h_status = nullptr;                       // codec status (LJ92), device + pinned
    // in-flight bookkeeping
    bool busy = false;
    int64_t ticket = -1;
    uint16_t *user_dst = nullptr;          // non-null: copy h_out -> user_dst in wait()
    size_t out_bytes = 0;
    mlvb_frame_result result{};
    // asynchronous submission (frames whose pipeline waits on the host, see abi.cu): set under mlvb_context::mu
    bool async = false, job_done = false;
    int job_rc = 0;
};

// A host batch in flight (mlvb_process_frames): device staging for n payloads, n work frames and n finished frames
// of one clip, its own stream and scratch, so that batches issued by different host threads overlap.
struct BatchSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    uint8_t *d_in = nullptr;     size_t in_cap = 0;
    uint16_t *d_work = nullptr;  size_t work_cap = 0;
    uint16_t *d_out = nullptr;   size_t out_cap = 0;
    void *d_aux = nullptr;       size_t aux_cap = 0;
    int *d_status = nullptr, *h_status = nullptr; int status_cap = 0;
    bool busy = false;
};

struct AsyncJob {                          // one submitted frame waiting for a submit worker
    Slot *slot;
    struct frame_headers hdr;
    mlvb_options opts;
    std::string clip;
    const void *src;
    size_t payload_bytes;
    uint16_t *dst;
};

struct mlvb_context {
    int device = 0;
    EvLuts luts{};
    int *d_raw2ev_base = nullptr;
    uint16_t *d_ev2raw_pos = nullptr;
    int *d_ev2raw_full = nullptr;
    bool ev2raw_octaves_ok = false;        // ev2raw[e] == ev2raw[13 EV + (e mod EV)] >> (13 - e / EV) for all e (checked at creation)
    int sm_count = 0;

    std::mutex mu;                         // slots + tickets
    std::condition_variable cv;
    std::vector<Slot> slots;
    int64_t next_ticket = 0;

    // submit workers: frames of the full dual-ISO pipeline block on statistics read-backs, so mlvb_submit hands them
    // to a few host threads (one frame each, the slot's own stream) once the clip's per-clip state exists
    std::vector<std::thread> submit_workers;
    std::deque<AsyncJob> jobs;
    std::mutex job_mu;
    std::condition_variable job_cv;
    bool stopping = false;
    std::set<std::string> async_clips;     // clip + option keys whose first frame has been processed (under job_mu)

    std::mutex clip_mu;                    // per-clip state (creation is once-only, under this lock)
    std::map<std::string, StripesState> stripes;
    BadPixelMap bad_maps[8];
    int bad_map_cursor = 0;
    std::vector<FocusPixelMap> focus_maps;
    GlibcRand dither_rng;

    // scratch for the device-batch entry point and the per-clip statistics passes
    cudaStream_t batch_stream = nullptr;
    void *d_scratch = nullptr;  size_t scratch_cap = 0;
    void *d_stat = nullptr;     size_t stat_cap = 0;
    int *d_batch_status = nullptr; size_t batch_status_cap = 0;
    void *d_batch_aux = nullptr; size_t batch_aux_cap = 0;        // per-frame codec status of a batch

    // batch lanes: extra streams + scratch so that independent dual-ISO frames of one device batch overlap their
    // statistics read-backs and host epilogues with each other's kernels (abi.cu: run_pipeline)
    struct BatchLane { cudaStream_t stream = nullptr; cudaEvent_t done = nullptr; void *d_aux = nullptr; size_t aux_cap = 0; };
    std::vector<BatchLane> batch_lanes;
    cudaEvent_t batch_fork = nullptr;
    std::mutex lanes_mu;                   // one device batch at a time owns the lanes (host batches run concurrently)

    // host batches (mlvb_process_frames)
    std::mutex hb_mu;
    std::condition_variable hb_cv;
    std::vector<BatchSlot> host_batches;
    bool blocking_sync = true;             // waits sleep on the event instead of spinning ($MLVB_BLOCKING_SYNC=0: spin)
    bool sync_submit = false, no_wide = false, wide_segments = false;   // $MLVB_SYNC_SUBMIT, $MLVB_NO_WIDE, $MLVB_WIDE_SEGMENTS (read once)
    int batch_lane_count = 0;              // dual-ISO frames of a device batch in flight at once ($MLVB_BATCH_LANES; 0: by scratch size)
    int spin_us = 150;                     // stream_wait polls this long before it sleeps ($MLVB_SPIN_US)

    std::atomic<uint64_t> launches{0};
    std::atomic<uint64_t> path_count[3] = {{0}, {0}, {0}};   // fused strip kernel, fused wide kernel, host batches (mlvb_path_count)

    // optional per-stage device timing (mlvb_profile_begin / mlvb_profile_end), bench.py's roofline leg
    bool profiling = false;
    struct StageSpan { int stage; cudaEvent_t a, b; };
    std::vector<StageSpan> spans;

    int ensure_scratch(size_t bytes);
    int ensure_stat(size_t bytes);
};

// Wait for a stream from a host thread.  cudaStreamSynchronize spins by default, which costs a core per waiting
// thread for as long as the wait lasts; a blocking event sleeps but wakes up tens of microseconds late, and the
// dual-ISO path waits four times per frame for kernels that take 20 .. 60 us.  So: poll for a short while (most waits
// end there), then sleep on a per-thread blocking event.  $MLVB_BLOCKING_SYNC=0: plain cudaStreamSynchronize.
inline cudaError_t stream_wait(const mlvb_context *ctx, cudaStream_t st)
{
    if (!ctx->blocking_sync) return cudaStreamSynchronize(st);
    constexpr int MAX_DEV = 64;
    thread_local cudaEvent_t ev[MAX_DEV] = {};
    const int dev = ctx->device;
    if (dev < 0 || dev >= MAX_DEV) return cudaStreamSynchronize(st);
    if (!ev[dev]) {
        cudaError_t e = cudaEventCreateWithFlags(&ev[dev], cudaEventDisableTiming | cudaEventBlockingSync);
        if (e != cudaSuccess) { ev[dev] = nullptr; return e; }
    }
    cudaError_t e = cudaEventRecord(ev[dev], st);
    if (e != cudaSuccess) return e;
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        e = cudaEventQuery(ev[dev]);
        if (e != cudaErrorNotReady) return e;
        if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(ctx->spin_us)) break;
    }
    return cudaEventSynchronize(ev[dev]);
}

// stage ids reported by mlvb_profile_end
enum { ST_UNPACK = 0, ST_PIXFIX = 1, ST_CHROMA = 2, ST_STRIPES = 3, ST_PATTERN = 4, ST_DUALISO = 5, ST_LJ92 = 6, ST_COUNT = 8 };

// RAII span: records a cudaEvent pair around a stage when profiling is on
struct StageTimer {
    mlvb_context *ctx; cudaStream_t st; cudaEvent_t b = nullptr;
    StageTimer(mlvb_context *c, int stage, cudaStream_t s) : ctx(c), st(s)
    {
        if (!c->profiling) return;
        cudaEvent_t a;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st);
        c->spans.push_back({stage, a, b});
    }
    ~StageTimer() { if (b) cudaEventRecord(b, st); }
};

// host EV tables (built once per process with libm, uploaded per context)
const int *host_raw2ev_base();
const double *host_raw2evf_base();
const int *host_ev2raw_base();             // 24*EV entries, index 0 <-> e = -10 EV

struct FrameGeom {
    int w, h, bpp, black, white, crop_x, crop_y, frame_size;
    size_t npix;
};
FrameGeom geom_from_headers(const struct frame_headers *hdr);

// the single-ISO correction chain on device buffers (process_frame order, main.c:966-997).
// d_a holds the unpacked frames on entry; the finished frames are written to d_out (may alias d_a
// only when no chroma smoothing is requested).  Creates per-clip state from frame 0 when missing.
int run_single_iso_chain(mlvb_context *ctx, const struct frame_headers *hdr, const FrameGeom &g,
                         const mlvb_options &opts, const char *mlv_filename, uint16_t *d_a, uint16_t *d_out,
                         size_t frame_stride, int nframes, int skip_chroma, int skip_pixfix, cudaStream_t st);

// dualiso.cu
size_t dual_iso_scratch_bytes(int w, int h, int interp_method);
int run_cr2hdr20(mlvb_context *ctx, const struct frame_headers *hdr, const FrameGeom &g, uint16_t *d_img, int interp_method,
                 int use_fullres, int use_alias_map, int cs_method, int fix_bad_pixels_mode, void *d_aux, cudaStream_t st);
void dual_iso_reset_tables(mlvb_context *ctx);
void dual_iso_free_tables(mlvb_context *ctx);

// fused.cu: MLVB_OK = enqueued, 1 = not eligible (use the general path), < 0 = error
int try_fused_single_iso(mlvb_context *ctx, const struct frame_headers *hdr, const FrameGeom &g, const mlvb_options &opts,
                         const char *mlv_filename, const void *d_payload, size_t payload_stride, size_t payload_bytes,
                         uint16_t *d_out, size_t out_stride_px, int nframes, cudaStream_t st);

// hdrpreview.cu
size_t hdr_preview_scratch_bytes(int white);
size_t deflicker_scratch_bytes(int bpp);
int run_hdr_preview(mlvb_context *ctx, const struct frame_headers *hdr, const FrameGeom &g, uint16_t *d_img, void *d_aux,
                    cudaStream_t st);
int run_deflicker(mlvb_context *ctx, const FrameGeom &g, const uint16_t *d_img, int target, void *d_aux, cudaStream_t st,
                  int32_t bias[2]);

// per-clip state accessors (call with ctx->clip_mu held)
int get_bad_pixel_map(mlvb_context *ctx, const struct frame_headers *hdr, const FrameGeom &g, int aggressive,
                      const uint16_t *d_img, cudaStream_t st, std::shared_ptr<PixelList> *out);
int get_focus_pixel_map(mlvb_context *ctx, const struct frame_headers *hdr, const FrameGeom &g, std::shared_ptr<PixelList> *out);
int compute_stripes(mlvb_context *ctx, const FrameGeom &g, const uint16_t *d_img, cudaStream_t st, StripeCoef *out);

// apply a pixel list to device frames: level schedule for the 2-D interpolator, independent row segments for
// the horizontal one (dual ISO); counts the launches
int apply_pixel_list(mlvb_context *ctx, const PixelList &L, uint16_t *d_img, const FrameGeom &g, size_t frame_stride, int nframes,
                     int dual_iso, int edge_rules, cudaStream_t st);

// slot lease for the synchronous drop-in entry points (dropin.cu)
Slot *acquire_slot(mlvb_context *ctx, bool may_block = true);   // nullptr only when !may_block and every slot is busy
void release_slot(mlvb_context *ctx, Slot *s);
int slot_reserve(Slot &s, size_t packed_bytes, size_t frame_bytes);
int reserve_device(void **p, size_t *cap, size_t bytes);

// stage scratch requirement for a frame of this geometry under these options (0 if none)
size_t aux_bytes_for(const FrameGeom &g, const mlvb_options &opts);
