// abi.cu -- the mlvb_* entry points of include/mlvfs_b200.h: context lifecycle, the slot ring that
// pipelines host frames (pinned H2D -> kernels -> pinned D2H on one stream per slot), and the
// device-resident batch form.  The reference-named drop-in symbols live in dropin.cu.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <thread>
#include <vector>

#include "context.cuh"

namespace {

// $MLVB_TRACE: timestamps of the host-batch phases on stderr (debugging aid: who waits for whom)
static void trace_point(const char *what)
{
    static const bool on = getenv("MLVB_TRACE") != nullptr;
    if (!on) return;
    static const auto t0 = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[mlvb %9.3f ms] thread %zx: %s\n", ms, std::hash<std::thread::id>()(std::this_thread::get_id()) & 0xFFFF, what);
}

// clip + the options that shape the per-clip state: frames of a key whose first frame has been through are independent
// of each other (ctx->async_clips, under ctx->job_mu)
static std::string clip_option_key(const char *mlv_filename, const mlvb_options *opts)
{
    char tag[64];
    snprintf(tag, sizeof(tag), "|%d.%d.%d.%d.%d.%d", opts->fix_bad_pixels, opts->fix_stripes, opts->chroma_smooth,
             opts->hdr_interpolation_method, opts->hdr_no_fullres, opts->hdr_no_alias_map);
    return std::string(mlv_filename ? mlv_filename : "") + tag;
}

int round_up(size_t v, size_t a, size_t *out) { *out = (v + a - 1) / a * a; return 0; }

// Pinned host memory handed out by mlvb_host_alloc: a process-wide pool.  Freed blocks are kept (up to
// $MLVB_PIN_POOL_MB, default 4096) and handed out again, because cudaHostAlloc / cudaFreeHost are millisecond-scale,
// driver-serialised calls and the frame builder needs a frame-sized buffer per frame (the reference's two
// malloc()s at main.c:931-933).  The table also answers "is this caller buffer pinned?" without a driver call.
struct PinPool {
    std::mutex mu;
    std::map<uintptr_t, size_t> blocks;                  // every block we own (in use or free): base -> bytes
    std::multimap<size_t, void *> free_blocks;           // bytes -> base
    size_t free_bytes = 0, cap_bytes = 0;
    bool cap_read = false;
};
PinPool &pin_pool() { static PinPool *p = new PinPool(); return *p; }     // leaked on purpose: outlives static destructors

constexpr size_t PIN_GRANULE = 64 * 1024;

bool in_pin_pool(const void *p, size_t bytes)
{
    PinPool &P = pin_pool();
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.blocks.upper_bound((uintptr_t)p);
    if (it == P.blocks.begin()) return false;
    --it;
    return (uintptr_t)p + bytes <= it->first + it->second;
}

// caller buffers that did not come from mlvb_host_alloc (torch pinned tensors, cudaHostRegister) cost one driver query
bool host_is_pinned(const void *p, size_t bytes)
{
    if (in_pin_pool(p, bytes)) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

}  // namespace

int slot_reserve(Slot &s, size_t packed_bytes, size_t frame_bytes)
{
    size_t pb, fb;
    round_up(packed_bytes + 1024, 256, &pb);
    round_up(frame_bytes, 256, &fb);
    if (pb > s.packed_cap) {
        if (s.d_packed) cudaFree(s.d_packed);
        if (s.h_in) cudaFreeHost(s.h_in);
        s.d_packed = nullptr; s.h_in = nullptr; s.packed_cap = s.h_in_cap = 0;
        MLVB_CUDA_OK(cudaMalloc(&s.d_packed, pb));
        MLVB_CUDA_OK(cudaHostAlloc(&s.h_in, pb, cudaHostAllocDefault));
        s.packed_cap = s.h_in_cap = pb;
    }
    if (fb > s.frame_cap) {
        if (s.d_a) cudaFree(s.d_a);
        if (s.d_b) cudaFree(s.d_b);
        if (s.h_out) cudaFreeHost(s.h_out);
        s.d_a = s.d_b = nullptr; s.h_out = nullptr; s.frame_cap = s.h_out_cap = 0;
        MLVB_CUDA_OK(cudaMalloc(&s.d_a, fb));
        MLVB_CUDA_OK(cudaMalloc(&s.d_b, fb));
        MLVB_CUDA_OK(cudaHostAlloc(&s.h_out, fb, cudaHostAllocDefault));
        s.frame_cap = s.h_out_cap = fb;
    }
    return MLVB_OK;
}

int reserve_device(void **p, size_t *cap, size_t bytes)
{
    if (bytes <= *cap) return MLVB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    MLVB_CUDA_OK(cudaMalloc(p, bytes));
    *cap = bytes;
    return MLVB_OK;
}

size_t aux_bytes_for(const FrameGeom &g, const mlvb_options &opts)
{
    size_t need = 0;
    if (opts.fix_pattern_noise) need = std::max(need, pattern_noise_scratch_bytes(g.w, g.h));
    if (opts.dual_iso == 2) need = std::max(need, dual_iso_scratch_bytes(g.w, g.h, opts.hdr_interpolation_method));
    if (opts.dual_iso == 1) need = std::max(need, hdr_preview_scratch_bytes((uint16_t)g.white));
    if (opts.deflicker) need = std::max(need, deflicker_scratch_bytes(g.bpp));
    return need;
}

Slot *acquire_slot(mlvb_context *ctx, bool may_block)
{
    std::unique_lock<std::mutex> lk(ctx->mu);
    for (;;) {
        for (auto &c : ctx->slots)
            if (!c.busy) { c.busy = true; c.ticket = ctx->next_ticket++; return &c; }
        if (!may_block) return nullptr;
        ctx->cv.wait(lk);
    }
}

void release_slot(mlvb_context *ctx, Slot *s)
{
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        s->busy = false;
        s->ticket = -1;
    }
    ctx->cv.notify_all();            // the condition variable is shared with mlvb_wait's job_done waiters: wake them all
}

namespace {

// Decode the VIDF payload into unpacked 16-bit frames at d_frames (main.c:569-706 dispatch).
int decode_payload(mlvb_context *ctx, const struct frame_headers *hdr, const FrameGeom &g, const void *d_payload,
                   size_t payload_stride, size_t payload_bytes, uint16_t *d_frames, size_t frame_stride, int nframes,
                   int *d_status, void *d_aux, size_t aux_cap, cudaStream_t st)
{
    const int vc = hdr->file_hdr.videoClass;
    if (vc & MLVB_VIDEO_CLASS_FLAG_LZMA) return MLVB_ERR_UNSUPPORTED;      // legacy codec stays on the CPU side
    if (vc & MLVB_VIDEO_CLASS_FLAG_LJ92) {                                 // main.c:617-681
        StageTimer t(ctx, ST_LJ92, st);
        int rc = launch_lj92_decode(d_payload, payload_stride, payload_bytes, d_frames, frame_stride, g.w, g.h, nframes,
                                    d_status, d_aux, aux_cap, st);
        if (rc > 0) { ctx->launches += rc; rc = MLVB_OK; }
        return rc;
    }
    if (payload_bytes < mlvb_packed_bytes((uint32_t)g.npix, g.bpp)) return MLVB_ERR_ARG;
    StageTimer t(ctx, ST_UNPACK, st);
    int rc = launch_unpack(d_payload, payload_stride, payload_bytes, d_frames, frame_stride, (uint32_t)g.npix, g.bpp,
                           nframes, st);
    if (rc == MLVB_OK) ctx->launches += 1;
    return rc;
}

// Everything process_frame does between get_image_data and the final header (main.c:942-997),
// on device buffers.  Finished frames end up in d_out.
int run_pipeline(mlvb_context *ctx, const struct frame_headers *hdr, const mlvb_options &opts, const char *mlv_filename,
                 const void *d_payload, size_t payload_stride, size_t payload_bytes, uint16_t *d_work, uint16_t *d_out,
                 size_t frame_stride, int nframes, int *d_status, void *d_aux, size_t aux_cap, cudaStream_t st,
                 mlvb_frame_result *res, mlvb_frame_result *per_frame = nullptr)
{
    // res: what frame 0 reports; per_frame (optional, nframes entries, pre-filled by the caller): the fields that can
    // differ between the frames of a batch -- whether the frame was converted and the levels that follow from it
    const FrameGeom g = geom_from_headers(hdr);
    if (g.w <= 0 || g.h <= 0) return MLVB_ERR_ARG;
    res->is_dual_iso = 0;
    res->black_level = g.black;
    res->white_level = g.white;
    res->exposure_bias[0] = hdr->rawi_hdr.raw_info.exposure_bias[0];
    res->exposure_bias[1] = hdr->rawi_hdr.raw_info.exposure_bias[1];

    // steady state of the full single-ISO chain: one fused kernel, one pass over HBM (fused.cu)
    int rc = try_fused_single_iso(ctx, hdr, g, opts, mlv_filename, d_payload, payload_stride, payload_bytes, d_out, frame_stride,
                                  nframes, st);
    if (rc <= 0) return rc;

    const bool cs = (opts.chroma_smooth == 2 || opts.chroma_smooth == 3 || opts.chroma_smooth == 5) && opts.dual_iso != 2 &&
                    g.black <= MLVB_MAX_BLACK;
    // without an out-of-place stage the chain can run directly in d_out
    uint16_t *d_a = cs ? d_work : d_out;
    rc = decode_payload(ctx, hdr, g, d_payload, payload_stride, payload_bytes, d_a, frame_stride, nframes, d_status, d_aux,
                        aux_cap, st);
    if (rc) return rc;
    if (opts.deflicker) {                                                           // main.c:943, 895-906
        int32_t bias[2];
        for (int f = nframes - 1; f >= 0; f--) {                                    // frame 0 last: its bias is reported
            rc = run_deflicker(ctx, g, d_a + (size_t)f * frame_stride, opts.deflicker, d_aux, st, bias);
            if (rc) return rc;
        }
        res->exposure_bias[0] = bias[0];
        res->exposure_bias[1] = bias[1];
    }
    if (opts.fix_pattern_noise) {                                                   // main.c:946-949
        StageTimer t(ctx, ST_PATTERN, st);
        for (int f = 0; f < nframes; f++) {
            rc = launch_pattern_noise((int16_t *)(d_a + (size_t)f * frame_stride), g.w, g.h, g.white, d_aux, st);
            if (rc) return rc;
            ctx->launches += 10;
        }
    }
    if (opts.dual_iso == 1) {                                                       // main.c:952-955
        for (int f = 0; f < nframes; f++) {
            uint16_t *fa = d_a + (size_t)f * frame_stride, *fo = d_out + (size_t)f * frame_stride;
            {
                StageTimer t(ctx, ST_DUALISO, st);
                rc = run_hdr_preview(ctx, hdr, g, fa, d_aux, st);
            }
            if (rc < 0) return rc;
            FrameGeom gf = g;
            if (rc == 1) { gf.black *= 4; gf.white *= 4; }                         // hdr.c:222-223
            if (f == 0) { res->is_dual_iso = rc; res->black_level = gf.black; res->white_level = gf.white; }
            // Converted frames skip the focus / bad-pixel stage (main.c:961-973).  The reference would
            // still chroma-smooth them with the x4 black level, which indexes its EV table out of bounds
            // (values reach 65535 against 24576 entries); we do not smooth converted preview frames.
            rc = run_single_iso_chain(ctx, hdr, gf, opts, mlv_filename, fa, fo, frame_stride, 1, rc == 1, rc == 1, st);
            if (rc) return rc;
        }
        return MLVB_OK;
    }
    if (opts.dual_iso == 2) {                                                       // main.c:956-973
        auto one_frame = [&](int f, void *aux, cudaStream_t s) -> int {
            uint16_t *fa = d_a + (size_t)f * frame_stride, *fo = d_out + (size_t)f * frame_stride;
            int r = run_cr2hdr20(ctx, hdr, g, fa, opts.hdr_interpolation_method, !opts.hdr_no_fullres, !opts.hdr_no_alias_map,
                                 opts.chroma_smooth, opts.fix_bad_pixels, aux, s);
            if (r < 0) return r;
            FrameGeom gf = g;
            if (r == 1) { gf.black *= 4; gf.white *= 4; }                          // hdr.c:1951-1952
            if (f == 0) { res->is_dual_iso = r; res->black_level = gf.black; res->white_level = gf.white; }
            if (per_frame) { per_frame[f].is_dual_iso = r; per_frame[f].black_level = gf.black; per_frame[f].white_level = gf.white; }
            // converted: only stripes remain; not converted: the usual chain minus chroma smoothing (main.c:975)
            return run_single_iso_chain(ctx, hdr, gf, opts, mlv_filename, fa, fo, frame_stride, 1, 1, r == 1, s);
        };
        StageTimer t(ctx, ST_DUALISO, st);
        // The batch's first frame goes first, on the caller's stream, when the clip's per-clip state (bad-pixel map, EV
        // tables, stripes) may not exist yet: it creates it.  Once a frame of this clip + option set has been through
        // (the same key the submit path keeps), every frame of the batch is independent.
        const std::string clip_key = clip_option_key(mlv_filename, &opts);
        bool primed;
        { std::lock_guard<std::mutex> lk(ctx->job_mu); primed = ctx->async_clips.count(clip_key) != 0; }
        const int first = primed ? 0 : 1;
        if (!primed) {
            rc = one_frame(0, d_aux, st);
            if (rc) return rc;
            { std::lock_guard<std::mutex> lk(ctx->job_mu); ctx->async_clips.insert(clip_key); }
            if (nframes == 1) return rc;
        } else if (nframes == 1) return one_frame(0, d_aux, st);
        // the frames are independent: a few host threads, each with its own stream and scratch, so that one
        // frame's statistics read-backs and scalar epilogues overlap the kernels of the others
        // the lanes (streams, scratch, fork event) belong to the context: concurrent host batches (mlvb_process_frames
        // has three slots) take turns here, while their copies and single-stream stages still overlap
        trace_point("dual ISO batch: waiting for the lanes");
        std::lock_guard<std::mutex> lanes_lock(ctx->lanes_mu);
        trace_point("dual ISO batch: lanes acquired");
        const size_t need = aux_bytes_for(g, opts);
        // measured on B200: frames with small scratch (C3: 5.9 MP mean23) gain up to 16 in flight, the AMaZE frames
        // (C4: 2.7 GB of tile workspaces each) are best at 8
        // (each lane is a host thread that polls: with fewer than 8 host cores per GPU -- eight ranks on a 32-core box -- 8)
        const int lane_cap = ctx->batch_lane_count > 0 ? ctx->batch_lane_count
                                                       : ((need > ((size_t)1 << 30) || ctx->spin_us < 1000) ? 8 : 16);
        const int nlanes = std::min(nframes - first, lane_cap);
        if (!ctx->batch_fork) MLVB_CUDA_OK(cudaEventCreateWithFlags(&ctx->batch_fork, cudaEventDisableTiming));
        while ((int)ctx->batch_lanes.size() < nlanes) {
            mlvb_context::BatchLane l;
            MLVB_CUDA_OK(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
            MLVB_CUDA_OK(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
            ctx->batch_lanes.push_back(l);
        }
        for (int l = 0; l < nlanes; l++) {
            rc = reserve_device(&ctx->batch_lanes[l].d_aux, &ctx->batch_lanes[l].aux_cap, need);
            if (rc) return rc;
        }
        MLVB_CUDA_OK(cudaEventRecord(ctx->batch_fork, st));
        std::vector<int> lane_rc(nlanes, MLVB_OK);
        std::vector<std::thread> workers;
        const bool was_profiling = ctx->profiling;
        ctx->profiling = false;                                                     // the span list is single-threaded; this stage is timed as a whole
        for (int l = 0; l < nlanes; l++)
            workers.emplace_back([&, l] {
                mlvb_context::BatchLane &L = ctx->batch_lanes[l];
                if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamWaitEvent(L.stream, ctx->batch_fork, 0) != cudaSuccess) {
                    lane_rc[l] = MLVB_ERR_CUDA;
                    return;
                }
                for (int f = first + l; f < nframes && lane_rc[l] == MLVB_OK; f += nlanes) lane_rc[l] = one_frame(f, L.d_aux, L.stream);
                if (cudaEventRecord(L.done, L.stream) != cudaSuccess) lane_rc[l] = MLVB_ERR_CUDA;
            });
        for (auto &th : workers) th.join();
        trace_point("dual ISO batch: lane threads joined");
        ctx->profiling = was_profiling;
        for (int l = 0; l < nlanes; l++) {
            MLVB_CUDA_OK(cudaStreamWaitEvent(st, ctx->batch_lanes[l].done, 0));
            if (lane_rc[l]) rc = lane_rc[l];
        }
        return rc;
    }
    return run_single_iso_chain(ctx, hdr, g, opts, mlv_filename, d_a, d_out, frame_stride, nframes, 0, 0, st);
}

std::mutex g_default_mu;
mlvb_context *g_default_ctx = nullptr;

}  // namespace

extern "C" {

int mlvb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int mlvb_context_create(int device, int nslots, mlvb_context **out)
{
    if (!out) return MLVB_ERR_ARG;
    *out = nullptr;
    if (mlvb_device_count() <= device) {
        fprintf(stderr, "libmlvfs_b200: CUDA device %d not available -- this library has no CPU path\n", device);
        return MLVB_ERR_CUDA;
    }
    MLVB_CUDA_OK(cudaSetDevice(device));
    mlvb_context *ctx = new mlvb_context();
    ctx->device = device;
    // a failure below returns from this lambda; the partly built context is torn down instead of leaked
    const int rc_create = [&]() -> int {
    const size_t n1 = (16384 + MLVB_MAX_BLACK) * sizeof(int), n3 = 24 * MLVB_EV_RES * sizeof(int);
    const size_t n2 = 14 * MLVB_EV_RES * sizeof(uint16_t);
    std::vector<uint16_t> pos(14 * MLVB_EV_RES);
    const int *ev2raw = host_ev2raw_base();
    for (int e = 0; e < 14 * MLVB_EV_RES; e++) pos[e] = (uint16_t)ev2raw[e + 10 * MLVB_EV_RES];
    // 2^(e/EV) = 2^k * 2^(f/EV): every octave of the table is the top octave shifted right (exact, not assumed:
    // checked here against the libm-built table; fused.cu keeps only the 64 KiB top octave in shared memory)
    ctx->ev2raw_octaves_ok = true;
    for (int e = 0; e < 14 * MLVB_EV_RES && ctx->ev2raw_octaves_ok; e++)
        ctx->ev2raw_octaves_ok = pos[e] == (pos[13 * MLVB_EV_RES + (e & (MLVB_EV_RES - 1))] >> (13 - (e >> 15)));
    if (cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) ctx->sm_count = 148;
    MLVB_CUDA_OK(cudaMalloc(&ctx->d_raw2ev_base, n1));
    MLVB_CUDA_OK(cudaMalloc(&ctx->d_ev2raw_pos, n2));
    MLVB_CUDA_OK(cudaMalloc(&ctx->d_ev2raw_full, n3));
    MLVB_CUDA_OK(cudaMemcpy(ctx->d_raw2ev_base, host_raw2ev_base(), n1, cudaMemcpyHostToDevice));
    MLVB_CUDA_OK(cudaMemcpy(ctx->d_ev2raw_pos, pos.data(), n2, cudaMemcpyHostToDevice));
    MLVB_CUDA_OK(cudaMemcpy(ctx->d_ev2raw_full, ev2raw, n3, cudaMemcpyHostToDevice));
    ctx->luts.raw2ev_base = ctx->d_raw2ev_base;
    ctx->luts.ev2raw_pos = ctx->d_ev2raw_pos;
    ctx->luts.ev2raw_full = ctx->d_ev2raw_full + 10 * MLVB_EV_RES;
    if (nslots <= 0) nslots = 4;
    ctx->slots.resize(nslots);
    {
        // A wait that spins burns a host core per frame in flight; with several GPUs (or several processes) on one
        // box those cores are what the copies and the other ranks need.  Events are created blocking by default.
        const char *bs = getenv("MLVB_BLOCKING_SYNC");
        ctx->blocking_sync = !(bs && *bs == '0');
        ctx->sync_submit = getenv("MLVB_SYNC_SUBMIT") != nullptr;
        ctx->no_wide = getenv("MLVB_NO_WIDE") != nullptr;
        ctx->wide_segments = getenv("MLVB_WIDE_SEGMENTS") != nullptr;
        const char *bl = getenv("MLVB_BATCH_LANES");
        if (bl) ctx->batch_lane_count = std::min(std::max(atoi(bl), 1), 16);
        // plenty of host cores per GPU: a waiting thread may poll for as long as a statistics kernel runs (lowest
        // latency); few cores per GPU (8 GPUs on a 32-core box, one process each): poll briefly, then sleep
        int ngpu = 1;
        if (cudaGetDeviceCount(&ngpu) != cudaSuccess || ngpu < 1) ngpu = 1;
        const unsigned cores = std::thread::hardware_concurrency();
        ctx->spin_us = (cores >= 8u * (unsigned)ngpu) ? 5000 : 150;
        const char *su = getenv("MLVB_SPIN_US");
        if (su) ctx->spin_us = std::max(atoi(su), 0);
    }
    const unsigned ev_flags = cudaEventDisableTiming | (ctx->blocking_sync ? cudaEventBlockingSync : 0);
    for (auto &s : ctx->slots) {
        MLVB_CUDA_OK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        MLVB_CUDA_OK(cudaEventCreateWithFlags(&s.done, ev_flags));
        MLVB_CUDA_OK(cudaMalloc(&s.d_status, sizeof(int)));
        MLVB_CUDA_OK(cudaMemset(s.d_status, 0, sizeof(int)));
        MLVB_CUDA_OK(cudaHostAlloc(&s.h_status, sizeof(int), cudaHostAllocDefault));
        *s.h_status = 0;
    }
    MLVB_CUDA_OK(cudaStreamCreateWithFlags(&ctx->batch_stream, cudaStreamNonBlocking));
    ctx->host_batches.resize(3);
    for (auto &b : ctx->host_batches) {
        MLVB_CUDA_OK(cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking));
        MLVB_CUDA_OK(cudaEventCreateWithFlags(&b.done, ev_flags));
    }
    return MLVB_OK;
    }();
    if (rc_create != MLVB_OK) {
        mlvb_context_destroy(ctx);
        return rc_create;
    }
    *out = ctx;
    return MLVB_OK;
}

void mlvb_context_destroy(mlvb_context *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    {
        std::lock_guard<std::mutex> lk(ctx->job_mu);
        ctx->stopping = true;
    }
    ctx->job_cv.notify_all();
    for (auto &t : ctx->submit_workers) t.join();
    cudaDeviceSynchronize();
    for (auto &s : ctx->slots) {
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.done) cudaEventDestroy(s.done);
        if (s.d_packed) cudaFree(s.d_packed);
        if (s.d_a) cudaFree(s.d_a);
        if (s.d_b) cudaFree(s.d_b);
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.d_status) cudaFree(s.d_status);
        if (s.d_aux) cudaFree(s.d_aux);
        if (s.h_status) cudaFreeHost(s.h_status);
    }
    if (ctx->d_batch_status) cudaFree(ctx->d_batch_status);
    if (ctx->d_batch_aux) cudaFree(ctx->d_batch_aux);
    for (auto &b : ctx->host_batches) {
        if (b.stream) cudaStreamDestroy(b.stream);
        if (b.done) cudaEventDestroy(b.done);
        if (b.d_in) cudaFree(b.d_in);
        if (b.d_work) cudaFree(b.d_work);
        if (b.d_out) cudaFree(b.d_out);
        if (b.d_aux) cudaFree(b.d_aux);
        if (b.d_status) cudaFree(b.d_status);
        if (b.h_status) cudaFreeHost(b.h_status);
    }
    for (auto &l : ctx->batch_lanes) {
        if (l.d_aux) cudaFree(l.d_aux);
        if (l.done) cudaEventDestroy(l.done);
        if (l.stream) cudaStreamDestroy(l.stream);
    }
    if (ctx->batch_fork) cudaEventDestroy(ctx->batch_fork);
    if (ctx->batch_stream) cudaStreamDestroy(ctx->batch_stream);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->d_stat) cudaFree(ctx->d_stat);
    mlvb_reset_clip_state(ctx);
    dual_iso_free_tables(ctx);
    cudaFree(ctx->d_raw2ev_base);
    cudaFree(ctx->d_ev2raw_pos);
    cudaFree(ctx->d_ev2raw_full);
    delete ctx;
}

mlvb_context *mlvb_default_context(void)
{
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (!g_default_ctx) {
        const char *dev = getenv("MLVB_DEVICE");
        if (mlvb_context_create(dev ? atoi(dev) : 0, 0, &g_default_ctx) != MLVB_OK) g_default_ctx = nullptr;
    }
    return g_default_ctx;
}

void *mlvb_host_alloc(size_t bytes)
{
    const size_t want = (std::max<size_t>(bytes, 1) + PIN_GRANULE - 1) / PIN_GRANULE * PIN_GRANULE;
    PinPool &P = pin_pool();
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.free_blocks.lower_bound(want);
        if (it != P.free_blocks.end() && it->first <= want + want / 4) {          // close enough in size: reuse
            void *p = it->second;
            P.free_bytes -= it->first;
            P.free_blocks.erase(it);
            return p;
        }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        mlvb_host_pool_trim();                                                     // give the cached blocks back and retry once
        if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    std::lock_guard<std::mutex> lk(P.mu);
    P.blocks[(uintptr_t)p] = want;
    return p;
}

void mlvb_host_free(void *p)
{
    if (!p) return;
    PinPool &P = pin_pool();
    {
        std::lock_guard<std::mutex> lk(P.mu);
        if (!P.cap_read) {
            const char *mb = getenv("MLVB_PIN_POOL_MB");
            P.cap_bytes = (size_t)(mb ? std::max(atol(mb), 0L) : 4096) << 20;
            P.cap_read = true;
        }
        auto it = P.blocks.find((uintptr_t)p);
        if (it == P.blocks.end()) { fprintf(stderr, "libmlvfs_b200: mlvb_host_free(%p): not from mlvb_host_alloc\n", p); return; }
        if (P.free_bytes + it->second <= P.cap_bytes) {
            P.free_blocks.emplace(it->second, p);
            P.free_bytes += it->second;
            return;
        }
        P.blocks.erase(it);
    }
    cudaFreeHost(p);
}

void mlvb_host_pool_trim(void)
{
    PinPool &P = pin_pool();
    std::vector<void *> victims;
    {
        std::lock_guard<std::mutex> lk(P.mu);
        for (auto &kv : P.free_blocks) { victims.push_back(kv.second); P.blocks.erase((uintptr_t)kv.second); }
        P.free_blocks.clear();
        P.free_bytes = 0;
    }
    for (void *v : victims) cudaFreeHost(v);
}

void mlvb_reset_clip_state(mlvb_context *ctx)
{
    if (!ctx) return;
    std::lock_guard<std::mutex> lk(ctx->clip_mu);
    ctx->stripes.clear();
    for (auto &m : ctx->bad_maps) m = BadPixelMap();
    ctx->bad_map_cursor = 0;
    ctx->focus_maps.clear();
    dual_iso_reset_tables(ctx);          // the 20-bit EV tables remember the first frame's white level (hdr.c:1089-1093)
    std::lock_guard<std::mutex> jl(ctx->job_mu);
    ctx->async_clips.clear();            // the next frame of every clip recreates its state synchronously
}

void mlvb_seed_dither(mlvb_context *ctx, unsigned seed)
{
    if (!ctx) return;
    std::lock_guard<std::mutex> lk(ctx->clip_mu);
    ctx->dither_rng.seed(seed);
}

int mlvb_get_stripes(mlvb_context *ctx, const char *mlv_filename, int *needed, int coef[8])
{
    if (!ctx) return MLVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->clip_mu);
    auto it = ctx->stripes.find(mlv_filename ? mlv_filename : "");
    if (it == ctx->stripes.end() || !it->second.computed) return MLVB_ERR_ARG;
    if (needed) *needed = it->second.coef.needed;
    if (coef) memcpy(coef, it->second.coef.coef, sizeof(int) * 8);
    return 8;
}

int mlvb_get_bad_pixels(mlvb_context *ctx, uint64_t file_guid, int aggressive, int *xy, int cap)
{
    if (!ctx) return MLVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->clip_mu);
    for (auto &m : ctx->bad_maps)
        if (m.valid && m.file_guid == file_guid && m.aggressive == aggressive && m.list) {
            const int n = (int)m.list->host.size();
            for (int i = 0; i < n && i < cap; i++) { xy[2 * i] = m.list->host[i].x; xy[2 * i + 1] = m.list->host[i].y; }
            return n;
        }
    return MLVB_ERR_ARG;
}

void mlvb_profile_begin(mlvb_context *ctx)
{
    if (!ctx) return;
    for (auto &sp : ctx->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    ctx->spans.clear();
    ctx->profiling = true;
}

int mlvb_profile_end(mlvb_context *ctx, float *ms_per_stage, int *spans_per_stage, int nstages)
{
    if (!ctx) return MLVB_ERR_ARG;
    ctx->profiling = false;
    for (int i = 0; i < nstages; i++) { ms_per_stage[i] = 0.f; spans_per_stage[i] = 0; }
    int rc = MLVB_OK;
    for (auto &sp : ctx->spans) {
        float ms = 0.f;
        if (cudaEventSynchronize(sp.b) != cudaSuccess || cudaEventElapsedTime(&ms, sp.a, sp.b) != cudaSuccess) rc = MLVB_ERR_CUDA;
        else if (sp.stage < nstages) { ms_per_stage[sp.stage] += ms; spans_per_stage[sp.stage]++; }
        cudaEventDestroy(sp.a); cudaEventDestroy(sp.b);
    }
    ctx->spans.clear();
    return rc;
}

uint64_t mlvb_launch_count(mlvb_context *ctx) { return ctx ? (uint64_t)ctx->launches : 0; }
uint64_t mlvb_path_count(mlvb_context *ctx, int which) { return (ctx && which >= 0 && which < 3) ? (uint64_t)ctx->path_count[which] : 0; }

// ------------------------------------------------------------------ host-buffer pipeline

// H2D -> pipeline -> D2H -> event on the slot's stream (everything mlvb_submit enqueues for one frame)
static int run_slot_job(mlvb_context *ctx, Slot *s, const struct frame_headers *hdr, const mlvb_options *opts, const char *mlv_filename,
                        const void *src, size_t payload_bytes, uint16_t *dst)
{
    const FrameGeom g = geom_from_headers(hdr);
    const size_t frame_bytes = g.npix * 2;
    if (cudaMemcpyAsync(s->d_packed, src, payload_bytes, cudaMemcpyHostToDevice, s->stream) != cudaSuccess) return MLVB_ERR_CUDA;
    s->result = mlvb_frame_result();
    *s->h_status = 0;
    const bool coded = (hdr->file_hdr.videoClass & MLVB_VIDEO_CLASS_FLAG_LJ92) != 0;
    int rc = reserve_device(&s->d_aux, &s->aux_cap,
                            std::max(aux_bytes_for(g, *opts), coded ? lj92_scratch_bytes(payload_bytes, g.npix, 1) : (size_t)0));
    if (rc) return rc;
    rc = run_pipeline(ctx, hdr, *opts, mlv_filename, s->d_packed, 0, payload_bytes, s->d_a, s->d_b, g.npix, 1, s->d_status,
                      s->d_aux, s->aux_cap, s->stream, &s->result);
    if (rc == MLVB_OK && coded &&
        cudaMemcpyAsync(s->h_status, s->d_status, sizeof(int), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess)
        rc = MLVB_ERR_CUDA;
    if (rc) { cudaStreamSynchronize(s->stream); return rc; }
    s->out_bytes = frame_bytes;
    const bool dst_pinned = host_is_pinned(dst, frame_bytes);
    s->user_dst = dst_pinned ? nullptr : dst;
    if (cudaMemcpyAsync(dst_pinned ? (void *)dst : (void *)s->h_out, s->d_b, frame_bytes, cudaMemcpyDeviceToHost,
                        s->stream) != cudaSuccess ||
        cudaEventRecord(s->done, s->stream) != cudaSuccess) {
        cudaStreamSynchronize(s->stream);
        return MLVB_ERR_CUDA;
    }
    return MLVB_OK;
}

static void submit_worker(mlvb_context *ctx)
{
    cudaSetDevice(ctx->device);
    for (;;) {
        AsyncJob job;
        {
            std::unique_lock<std::mutex> lk(ctx->job_mu);
            ctx->job_cv.wait(lk, [&] { return ctx->stopping || !ctx->jobs.empty(); });
            if (ctx->jobs.empty()) return;                                  // stopping
            job = ctx->jobs.front();
            ctx->jobs.pop_front();
        }
        const int rc = run_slot_job(ctx, job.slot, &job.hdr, &job.opts, job.clip.c_str(), job.src, job.payload_bytes, job.dst);
        {
            std::lock_guard<std::mutex> lk(ctx->mu);
            job.slot->job_rc = rc;
            job.slot->job_done = true;
        }
        ctx->cv.notify_all();
    }
}

constexpr int SUBMIT_WORKERS = 8;

constexpr mlvb_ticket SUBMIT_WOULD_BLOCK = -100;          // internal: no free slot and the caller asked not to wait

static mlvb_ticket submit_frame(mlvb_context *ctx, const struct frame_headers *hdr, const void *payload, size_t payload_bytes,
                                const mlvb_options *opts, const char *mlv_filename, uint16_t *dst, bool may_block);

mlvb_ticket mlvb_submit(mlvb_context *ctx, const struct frame_headers *hdr, const void *payload, size_t payload_bytes,
                        const mlvb_options *opts, const char *mlv_filename, uint16_t *dst)
{
    return submit_frame(ctx, hdr, payload, payload_bytes, opts, mlv_filename, dst, true);
}

static mlvb_ticket submit_frame(mlvb_context *ctx, const struct frame_headers *hdr, const void *payload, size_t payload_bytes,
                                const mlvb_options *opts, const char *mlv_filename, uint16_t *dst, bool may_block)
{
    if (!ctx || !hdr || !payload || !opts || !dst) return MLVB_ERR_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return MLVB_ERR_CUDA;
    const FrameGeom g = geom_from_headers(hdr);
    const size_t frame_bytes = g.npix * 2;

    // Legacy LZMA clips (main.c:598-616): the range coder is serial, so the payload is expanded on the host -- straight
    // into the slot's pinned stage -- and the frame continues as an uncompressed one (unpack and everything after it on
    // the GPU), exactly where the reference calls dng_get_image_data on LzmaUncompress's output.
    struct frame_headers lzma_hdr;
    size_t lzma_out = 0;
    if (hdr->file_hdr.videoClass & MLVB_VIDEO_CLASS_FLAG_LZMA) {
        if (payload_bytes < 4 + 5 + 5) return MLVB_ERR_ARG;
        uint32_t stored;
        memcpy(&stored, payload, 4);
        lzma_out = stored;
        if (lzma_out < mlvb_packed_bytes((uint32_t)g.npix, g.bpp)) {
            fprintf(stderr, "libmlvfs_b200: LZMA frame expands to %zu bytes, the frame needs %zu\n", lzma_out,
                    mlvb_packed_bytes((uint32_t)g.npix, g.bpp));
            return MLVB_ERR_ARG;
        }
        lzma_hdr = *hdr;
        lzma_hdr.file_hdr.videoClass &= ~MLVB_VIDEO_CLASS_FLAG_LZMA;
    }

    Slot *s = acquire_slot(ctx, may_block);
    if (!s) return SUBMIT_WOULD_BLOCK;
    auto fail = [&](int rc) -> mlvb_ticket { release_slot(ctx, s); return rc; };
    int rc = slot_reserve(*s, std::max(payload_bytes, lzma_out), frame_bytes);
    if (rc) return fail(rc);
    s->async = false; s->job_done = false; s->job_rc = MLVB_OK;

    // H2D: straight from the caller's buffer when it is pinned, else through the slot's pinned stage
    const void *src = payload;
    if (lzma_out) {
        const uint8_t *in = (const uint8_t *)payload;
        size_t produced = lzma_out;
        if (mlvb_lzma_decode(s->h_in, &produced, in + 9, payload_bytes - 9, in + 4) != MLVB_OK) {
            fprintf(stderr, "libmlvfs_b200: LZMA Failed!\n");                             // main.c:612
            return fail(MLVB_ERR_ARG);
        }
        if (produced < lzma_out) memset(s->h_in + produced, 0, lzma_out - produced);      // stream ended early (end marker)
        src = s->h_in;
        payload_bytes = lzma_out;
        hdr = &lzma_hdr;
    } else if (!host_is_pinned(payload, payload_bytes)) { memcpy(s->h_in, payload, payload_bytes); src = s->h_in; }

    // The full dual-ISO pipeline waits on the host several times per frame (statistics read-backs): once the clip's
    // per-clip state exists (its first frame went through synchronously, below), such frames go to the submit workers
    // so that the frames in flight overlap.  payload (when pinned) and dst must stay valid until mlvb_wait.
    const bool blocking = opts->dual_iso == 2 && !ctx->profiling && !ctx->sync_submit;
    std::string key;
    if (blocking) {
        key = clip_option_key(mlv_filename, opts);
        std::unique_lock<std::mutex> lk(ctx->job_mu);
        if (ctx->async_clips.count(key)) {
            if (ctx->submit_workers.empty())
                for (int i = 0; i < SUBMIT_WORKERS; i++) ctx->submit_workers.emplace_back(submit_worker, ctx);
            s->async = true;
            AsyncJob job;
            job.slot = s; job.hdr = *hdr; job.opts = *opts; job.clip = mlv_filename ? mlv_filename : "";
            job.src = src; job.payload_bytes = payload_bytes; job.dst = dst;
            ctx->jobs.push_back(std::move(job));
            lk.unlock();
            ctx->job_cv.notify_one();
            return s->ticket;
        }
    }
    rc = run_slot_job(ctx, s, hdr, opts, mlv_filename, src, payload_bytes, dst);
    if (rc) return fail(rc);
    if (blocking) {
        std::lock_guard<std::mutex> lk(ctx->job_mu);
        ctx->async_clips.insert(key);
    }
    return s->ticket;
}

int mlvb_wait(mlvb_context *ctx, mlvb_ticket ticket, mlvb_frame_result *res)
{
    if (!ctx || ticket < 0) return MLVB_ERR_ARG;
    Slot *s = nullptr;
    int rc = MLVB_OK;
    {
        std::unique_lock<std::mutex> lk(ctx->mu);
        for (auto &c : ctx->slots) if (c.busy && c.ticket == ticket) { s = &c; break; }
        if (!s) return MLVB_ERR_ARG;
        if (s->async) {
            ctx->cv.wait(lk, [&] { return s->job_done; });
            rc = s->job_rc;
        }
    }
    if (rc != MLVB_OK) {
        // the worker already synchronised the stream
    } else if (cudaEventSynchronize(s->done) != cudaSuccess) {
        fprintf(stderr, "libmlvfs_b200: frame failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        rc = MLVB_ERR_CUDA;
    } else if (*s->h_status != 0) {
        fprintf(stderr, "libmlvfs_b200: LJ92: Failed (%d)\n", *s->h_status);      // main.c:671-679
        rc = MLVB_ERR_ARG;
    } else if (s->user_dst) {
        memcpy(s->user_dst, s->h_out, s->out_bytes);
    }
    s->result.status = rc;
    if (res) *res = s->result;
    release_slot(ctx, s);
    return rc;
}

int mlvb_process_frame(mlvb_context *ctx, const struct frame_headers *hdr, const void *payload, size_t payload_bytes,
                       const mlvb_options *opts, const char *mlv_filename, uint16_t *dst, mlvb_frame_result *res)
{
    const mlvb_ticket t = mlvb_submit(ctx, hdr, payload, payload_bytes, opts, mlv_filename, dst);
    if (t < 0) {
        if (res) { *res = mlvb_frame_result(); res->status = (int)t; }
        return (int)t;
    }
    return mlvb_wait(ctx, t, res);
}

// ------------------------------------------------------------------ host batches

namespace {

BatchSlot *acquire_batch_slot(mlvb_context *ctx)
{
    std::unique_lock<std::mutex> lk(ctx->hb_mu);
    for (;;) {
        for (auto &b : ctx->host_batches)
            if (!b.busy) { b.busy = true; return &b; }
        ctx->hb_cv.wait(lk);
    }
}

void release_batch_slot(mlvb_context *ctx, BatchSlot *b)
{
    { std::lock_guard<std::mutex> lk(ctx->hb_mu); b->busy = false; }
    ctx->hb_cv.notify_one();
}

// frames one device batch can hold: same geometry, levels, codec and crop offsets (one clip, no panning)
bool same_batch_shape(const struct frame_headers &a, const struct frame_headers &b)
{
    const FrameGeom x = geom_from_headers(&a), y = geom_from_headers(&b);
    return x.w == y.w && x.h == y.h && x.bpp == y.bpp && x.black == y.black && x.white == y.white && x.crop_x == y.crop_x &&
           x.crop_y == y.crop_y && a.file_hdr.videoClass == b.file_hdr.videoClass && a.file_hdr.fileGuid == b.file_hdr.fileGuid &&
           a.idnt_hdr.cameraModel == b.idnt_hdr.cameraModel && a.rawi_hdr.raw_info.width == b.rawi_hdr.raw_info.width &&
           a.rawi_hdr.raw_info.height == b.rawi_hdr.raw_info.height;
}

}  // namespace

int mlvb_process_frames(mlvb_context *ctx, int nframes, const struct frame_headers *hdrs, const void *const *payloads,
                        const size_t *payload_bytes, const mlvb_options *opts, const char *mlv_filename, uint16_t *const *dsts,
                        mlvb_frame_result *results)
{
    if (!ctx || nframes < 0 || (nframes && (!hdrs || !payloads || !payload_bytes || !opts || !dsts))) return MLVB_ERR_ARG;
    if (nframes == 0) return MLVB_OK;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return MLVB_ERR_CUDA;

    // One device batch needs frame-independent results and one shape; anything else is pipelined frame by frame
    // over the context's slots (still one call for the host).
    // (full dual-ISO frames included: a device batch runs them on the context's lanes, all at once when the clip is
    // primed; the preview conversion dual_iso == 1 stays per frame)
    static const bool diso_batches = getenv("MLVB_DISO_HOST_BATCH") != nullptr;     // measured slower end to end (below)
    bool batch = nframes >= 2 && (opts->dual_iso == 0 || (opts->dual_iso == 2 && diso_batches)) && opts->deflicker == 0 &&
                 !(hdrs[0].file_hdr.videoClass & MLVB_VIDEO_CLASS_FLAG_LZMA);
    for (int f = 1; f < nframes && batch; f++) batch = same_batch_shape(hdrs[0], hdrs[f]);
    if (!batch) {
        // Several threads may be in here at once, sharing the context's slots: a thread that already has frames in
        // flight never WAITS for a slot (two such threads could wait for each other for ever) -- it retires its own
        // oldest frame instead and tries again.
        std::deque<std::pair<int, mlvb_ticket>> inflight;
        int rc_all = MLVB_OK;
        auto retire_oldest = [&] {
            const int f = inflight.front().first;
            const int rc = mlvb_wait(ctx, inflight.front().second, results ? &results[f] : nullptr);
            inflight.pop_front();
            if (rc) rc_all = rc;
        };
        for (int f = 0; f < nframes; f++) {
            mlvb_ticket t;
            for (;;) {
                t = submit_frame(ctx, &hdrs[f], payloads[f], payload_bytes[f], opts, mlv_filename, dsts[f], inflight.empty());
                if (t != SUBMIT_WOULD_BLOCK) break;
                retire_oldest();
            }
            if (t < 0) {
                rc_all = (int)t;
                if (results) { results[f] = mlvb_frame_result(); results[f].status = (int)t; }
            } else inflight.emplace_back(f, t);
        }
        while (!inflight.empty()) retire_oldest();
        return rc_all;
    }

    const FrameGeom g = geom_from_headers(&hdrs[0]);
    if (g.w <= 0 || g.h <= 0) return MLVB_ERR_ARG;
    const bool coded = (hdrs[0].file_hdr.videoClass & MLVB_VIDEO_CLASS_FLAG_LJ92) != 0;
    size_t max_bytes = 0;
    for (int f = 0; f < nframes; f++) max_bytes = std::max(max_bytes, payload_bytes[f]);
    const size_t stride = (max_bytes + (coded ? 1024 : 0) + 255) / 256 * 256;
    const size_t frame_px = (g.npix + 127) / 128 * 128;                       // 256-byte aligned frames

    trace_point("process_frames: enter");
    BatchSlot *b = acquire_batch_slot(ctx);
    cudaStream_t st = b->stream;
    bool per_frame_status = false;                       // results[] already tell which frames failed
    int rc = [&]() -> int {
        int r = reserve_device((void **)&b->d_in, &b->in_cap, stride * nframes + 1024);
        if (!r) r = reserve_device((void **)&b->d_work, &b->work_cap, frame_px * 2 * nframes);
        if (!r) r = reserve_device((void **)&b->d_out, &b->out_cap, frame_px * 2 * nframes);
        if (!r) r = reserve_device(&b->d_aux, &b->aux_cap,
                                   std::max(aux_bytes_for(g, *opts), coded ? lj92_scratch_bytes(stride, g.npix, nframes) : (size_t)0));
        if (r) return r;
        if (b->status_cap < nframes) {
            if (b->d_status) cudaFree(b->d_status);
            if (b->h_status) cudaFreeHost(b->h_status);
            b->d_status = nullptr; b->h_status = nullptr; b->status_cap = 0;
            MLVB_CUDA_OK(cudaMalloc(&b->d_status, sizeof(int) * nframes));
            MLVB_CUDA_OK(cudaHostAlloc(&b->h_status, sizeof(int) * nframes, cudaHostAllocDefault));
            b->status_cap = nframes;
        }
        MLVB_CUDA_OK(cudaMemsetAsync(b->d_status, 0, sizeof(int) * nframes, st));
        for (int f = 0; f < nframes; f++) {
            if (!coded && payload_bytes[f] < mlvb_packed_bytes((uint32_t)g.npix, g.bpp)) return MLVB_ERR_ARG;
            MLVB_CUDA_OK(cudaMemcpyAsync(b->d_in + f * stride, payloads[f], payload_bytes[f], cudaMemcpyHostToDevice, st));
            if (coded && payload_bytes[f] < stride)                           // a stale tail must not look like stream data
                MLVB_CUDA_OK(cudaMemsetAsync(b->d_in + f * stride + payload_bytes[f], 0, stride - payload_bytes[f], st));
        }
        trace_point("process_frames: H2D enqueued");
        mlvb_frame_result res0;
        std::vector<mlvb_frame_result> per_frame;
        if (opts->dual_iso == 2) {                                            // frames of one clip may differ in "looks like dual ISO"
            mlvb_frame_result init = mlvb_frame_result();
            const FrameGeom g0 = geom_from_headers(&hdrs[0]);
            init.black_level = g0.black; init.white_level = g0.white;
            init.exposure_bias[0] = hdrs[0].rawi_hdr.raw_info.exposure_bias[0];
            init.exposure_bias[1] = hdrs[0].rawi_hdr.raw_info.exposure_bias[1];
            per_frame.assign(nframes, init);
        }
        r = run_pipeline(ctx, &hdrs[0], *opts, mlv_filename, b->d_in, stride, coded ? stride : max_bytes, b->d_work, b->d_out, frame_px,
                         nframes, b->d_status, b->d_aux, b->aux_cap, st, &res0, per_frame.empty() ? nullptr : per_frame.data());
        if (r) return r;
        if (coded) MLVB_CUDA_OK(cudaMemcpyAsync(b->h_status, b->d_status, sizeof(int) * nframes, cudaMemcpyDeviceToHost, st));
        for (int f = 0; f < nframes; f++)
            MLVB_CUDA_OK(cudaMemcpyAsync(dsts[f], b->d_out + f * frame_px, g.npix * 2, cudaMemcpyDeviceToHost, st));
        MLVB_CUDA_OK(cudaEventRecord(b->done, st));
        trace_point("process_frames: D2H enqueued");
        MLVB_CUDA_OK(cudaEventSynchronize(b->done));
        trace_point("process_frames: done");
        int rr = MLVB_OK;
        for (int f = 0; f < nframes; f++) {
            int s = MLVB_OK;
            if (coded && b->h_status[f] != 0) {
                fprintf(stderr, "libmlvfs_b200: LJ92: frame %d of the batch failed (%d)\n", f, b->h_status[f]);    // main.c:671-679
                s = rr = MLVB_ERR_ARG;
            }
            if (results) { results[f] = per_frame.empty() ? res0 : per_frame[f]; results[f].status = s; }
        }
        per_frame_status = true;
        return rr;
    }();
    if (rc && !per_frame_status) {
        cudaStreamSynchronize(st);                       // nothing of this batch may still be running when the slot is reused
        if (results)
            for (int f = 0; f < nframes; f++) { results[f] = mlvb_frame_result(); results[f].status = rc; }
    }
    ctx->path_count[2] += 1;
    release_batch_slot(ctx, b);
    return rc;
}

// ------------------------------------------------------------------ device-resident batch

int mlvb_process_batch_device(mlvb_context *ctx, const struct frame_headers *hdr, const mlvb_options *opts,
                              const char *mlv_filename, const void *d_payload, size_t payload_stride,
                              size_t payload_bytes, uint16_t *d_out, size_t out_stride_px, int nframes,
                              void *cuda_stream)
{
    if (!ctx || !hdr || !opts || !d_payload || !d_out || nframes <= 0) return MLVB_ERR_ARG;
    MLVB_CUDA_OK(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->batch_stream;
    const FrameGeom g = geom_from_headers(hdr);
    if (out_stride_px < g.npix) return MLVB_ERR_ARG;
    // one work frame per output frame, same stride as the output
    int rc = ctx->ensure_scratch((size_t)nframes * out_stride_px * sizeof(uint16_t));
    if (rc) return rc;
    mlvb_frame_result res;
    const bool coded = (hdr->file_hdr.videoClass & MLVB_VIDEO_CLASS_FLAG_LJ92) != 0;
    if (coded && ctx->batch_status_cap < (size_t)nframes) {
        if (ctx->d_batch_status) cudaFree(ctx->d_batch_status);
        ctx->d_batch_status = nullptr; ctx->batch_status_cap = 0;
        MLVB_CUDA_OK(cudaMalloc(&ctx->d_batch_status, sizeof(int) * nframes));
        ctx->batch_status_cap = nframes;
    }
    rc = reserve_device(&ctx->d_batch_aux, &ctx->batch_aux_cap,
                        std::max(aux_bytes_for(g, *opts), coded ? lj92_scratch_bytes(payload_bytes, g.npix, nframes) : (size_t)0));
    if (rc) return rc;
    rc = run_pipeline(ctx, hdr, *opts, mlv_filename, d_payload, payload_stride, payload_bytes, (uint16_t *)ctx->d_scratch,
                      d_out, out_stride_px, nframes, ctx->d_batch_status, ctx->d_batch_aux, ctx->batch_aux_cap, st, &res);
    if (rc == MLVB_OK && coded) {
        // a compressed batch reports corrupt streams synchronously (the decode dwarfs the sync)
        std::vector<int> status(nframes);
        MLVB_CUDA_OK(cudaMemcpyAsync(status.data(), ctx->d_batch_status, sizeof(int) * nframes, cudaMemcpyDeviceToHost, st));
        MLVB_CUDA_OK(stream_wait(ctx, st));
        for (int f = 0; f < nframes; f++)
            if (status[f] != 0) { fprintf(stderr, "libmlvfs_b200: LJ92: frame %d failed (%d)\n", f, status[f]); rc = MLVB_ERR_ARG; }
    }
    return rc;
}

}  // extern "C"
