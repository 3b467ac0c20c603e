// fused.cu -- the steady-state single-ISO chain of BASELINE config 2 in ONE pass over HBM:
//   14-bit packed payload -> unpack -> bad-pixel repair -> 3x3 median chroma smoothing -> stripe gains
//   -> 16-bit frame.
//
// Replaces, fused: dng.c:813-872 (unpack), cs.c:135-168 + 314-330 (bad-pixel interpolation),
// cs.c:49-84 + chroma_smooth.c:22-71 (3x3), stripes.c:250-266 (apply).  Traffic is the compulsory
// 1.75 B/px in + 2 B/px out; no intermediate frame is written.
//
// Structure ("warp strip", see chroma.cu): a warp owns 32 adjacent RGGB quad columns (lanes 1..30 write,
// lanes 0/31 are halo) and walks down a segment of quad rows.  Each lane pulls its two pixels of a
// row straight out of the packed bit stream (two aligned 32-bit loads + funnel shift; the bit phase of
// a lane is constant down the column because 14*W is a multiple of 32), keeps the (ge, dr, db) EV
// triplets of rows y-1, y, y+1 in registers, and gets the neighbours' sorted 3-element columns by
// shuffle: median9 = med3(max(lows), med3(mids), min(highs)).
//
// Bad pixels: when the clip's list has no dependent entries (level schedule depth 1 -- the normal
// case for sparse sensor defects) the repaired values are computed first by a tiny per-entry kernel
// that reads its 12 neighbours directly from the packed stream, and the strip kernel patches them in as
// the rows go by (per-warp buckets built once per clip).  Lists with dependencies, focus-pixel maps,
// other bit depths or smoothing sizes use the general multi-kernel path.
#include <algorithm>
#include <map>
#include <mutex>
#include <type_traits>
#include <vector>

#include "context.cuh"

namespace {

constexpr int FS_ROWS = 34;          // output quad rows per warp
constexpr int FS_WARPS = 4;

struct PatchItem { int qrow; short lane; short sub; unsigned entry; };   // sub: 0 r, 1 g1, 2 g2, 3 b

struct FusedParams {
    const uint8_t *packed; size_t payload_stride;
    uint16_t *out; size_t out_stride;
    int w, h, black;
    const int *raw2ev;                // indexed by raw value
    const uint16_t *ev2raw;
    int stripes, black16, white16;
    int coef[8];
    const PatchItem *items; const unsigned *bucket_start; const uint16_t *vals; unsigned n_entries;
    int nstrips;
};

__device__ __forceinline__ int packed_px(const uint8_t *frame, int W, int x, int y)
{
    const size_t bit = ((size_t)y * W + x) * 14;
    const uint16_t *wds = reinterpret_cast<const uint16_t *>(frame);
    const size_t k = bit >> 4;
    const int s = (int)(bit & 15);
    uint32_t pair = (uint32_t)wds[k] << 16;
    if (s + 14 > 16) pair |= wds[k + 1];
    return (int)((pair >> (18 - s)) & 0x3FFF);
}

// cs.c:135-168 (interpolate_pixel) evaluated on the packed frame; interior entries only
__global__ void patch_values_kernel(const uint8_t *__restrict__ packed, size_t payload_stride, int w, int h, int black,
                                    const int *__restrict__ raw2ev, const uint16_t *__restrict__ ev2raw,
                                    const PixelXY *__restrict__ list, unsigned n, int crop_x, int crop_y,
                                    uint16_t *__restrict__ vals)
{
    const unsigned m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const uint8_t *f = packed + (size_t)blockIdx.y * payload_stride;
    const int x = list[m].x - crop_x, y = list[m].y - crop_y;
    if (!(x > 2 && x < w - 3 && y > 2 && y < h - 3)) return;
    auto ev = [&](int dx, int dy) { return __ldg(raw2ev + packed_px(f, w, x + dx, y + dy)); };
    const int dv1 = wabs(wsub(ev(0, 3), ev(0, 1))), dv2 = wabs(wsub(ev(0, -1), ev(0, -3)));
    const int dh1 = wabs(wsub(ev(3, 0), ev(1, 0))), dh2 = wabs(wsub(ev(-1, 0), ev(-3, 0)));
    const int sum = wadd(wadd(dh1, dh2), wadd(dv1, dv2));
    int v;
    if (sum == 0) v = packed_px(f, w, x + 2, y);
    else {
        const int den = wmul(3, sum);
        const int cv1 = ((sum - dv1) << 8) / den, cv2 = ((sum - dv2) << 8) / den;
        const int ch1 = ((sum - dh1) << 8) / den, ch2 = ((sum - dh2) << 8) / den;
        const int e = (wmul(ev(0, 2), cv1) >> 8) + (wmul(ev(0, -2), cv2) >> 8) + (wmul(ev(2, 0), ch1) >> 8) + (wmul(ev(-2, 0), ch2) >> 8);
        v = __ldg(ev2raw + clamp_ev(e)) + black;
    }
    vals[(size_t)blockIdx.y * n + m] = (uint16_t)v;
}

struct RowQ { uint32_t top, bot; int ge, dr, db; };

__device__ __forceinline__ void sort3(int &a, int &b, int &c)
{
    int t = min(a, b); b = max(a, b); a = t;
    t = min(b, c); c = max(b, c); b = t;
    t = min(a, b); b = max(a, b); a = t;
}
__device__ __forceinline__ int med3(int a, int b, int c) { return max(min(a, b), min(max(a, b), c)); }

__device__ __forceinline__ int median9_columns(int a, int b, int c)
{
    sort3(a, b, c);
    const int al = __shfl_up_sync(0xFFFFFFFFu, a, 1), ar = __shfl_down_sync(0xFFFFFFFFu, a, 1);
    const int bl = __shfl_up_sync(0xFFFFFFFFu, b, 1), br = __shfl_down_sync(0xFFFFFFFFu, b, 1);
    const int cl = __shfl_up_sync(0xFFFFFFFFu, c, 1), cr = __shfl_down_sync(0xFFFFFFFFu, c, 1);
    return med3(max(max(al, a), ar), med3(bl, b, br), min(min(cl, c), cr));
}

__device__ __forceinline__ uint32_t stripe_gain_fast(uint32_t v, uint32_t coef, int black16, int white16)
{
    if ((int)v > black16 + 64) {
        const uint32_t t = __umulhi((v - (uint32_t)black16) << 16, coef) + (uint32_t)black16;
        return min(t, (uint32_t)white16);
    }
    return v;
}

// two 14-bit pixels starting at bit phase s of the 32-bit word pair at p (stream is 16-bit LE words, MSB first)
__device__ __forceinline__ uint32_t two_px(const uint32_t *p, int s)
{
    const uint32_t w0 = __byte_perm(__ldg(p), 0, 0x1032);
    const uint32_t w1 = (s > 4) ? __byte_perm(__ldg(p + 1), 0, 0x1032) : 0u;
    const uint32_t bits = __funnelshift_l(w1, w0, s) >> 4;                  // 28 bits: p0 p1
    return (bits >> 14) | ((bits & 0x3FFFu) << 16);
}

template <bool STRIPES>
__global__ void __launch_bounds__(FS_WARPS * 32)
fused3_strip_kernel(const FusedParams P)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = P.w, pw = w >> 1, ph = P.h >> 1;
    const int strip = blockIdx.x * FS_WARPS + warp;
    if (strip * 30 >= pw) return;
    const int qc = strip * 30 + lane - 1;
    const bool col_ok = qc >= 0 && qc < pw;
    const int x = 2 * qc;
    const uint8_t *frame = P.packed + (size_t)blockIdx.z * P.payload_stride;
    uint16_t *out = P.out + (size_t)blockIdx.z * P.out_stride;
    const int *raw2ev = P.raw2ev;
    const int qr0 = blockIdx.y * FS_ROWS, qr1 = min(qr0 + FS_ROWS, ph);
    const bool x_inside = x >= 4 && x < w - 4;
    const bool writer = col_ok && lane >= 1 && lane <= 30;
    const int black = P.black, h = P.h, black16 = P.black16, white16 = P.white16;
    const uint16_t *ev2raw = P.ev2raw;
    uint32_t c0 = 0, c1 = 0;
    if (STRIPES) { c0 = (uint32_t)P.coef[x & 7]; c1 = (uint32_t)P.coef[(x + 1) & 7]; }

    // bit phase of this lane's pixel pair; constant down the column, rows are 14*w bits apart
    const long long bit_top0 = (long long)x * 14;
    const int words_per_row = (14 * w) >> 5;                                // 14*w is a multiple of 32 (w % 16 == 0)
    const int s = (int)(bit_top0 & 31);
    const uint32_t *colp = reinterpret_cast<const uint32_t *>(frame) + (bit_top0 >> 5);

    // patches of this warp (sorted by the quad row they belong to)
    const unsigned bucket = blockIdx.y * P.nstrips + strip;
    unsigned pc = P.items ? P.bucket_start[bucket] : 0u;
    const unsigned pe = P.items ? P.bucket_start[bucket + 1] : 0u;
    const uint16_t *vals = P.vals + (size_t)blockIdx.z * P.n_entries;

    auto load_row = [&](int qr) {
        RowQ q = {0u, 0u, 0, 0, 0};
        const bool ok = col_ok && qr >= 0 && qr < ph;
        if (ok) {
            const uint32_t *pt = colp + (size_t)(2 * qr) * words_per_row;
            q.top = two_px(pt, s);
            q.bot = two_px(pt + words_per_row, s);
        }
        while (pc < pe && P.items[pc].qrow == qr) {                         // warp-uniform
            const PatchItem it = P.items[pc];
            if (lane == it.lane) {
                const uint32_t v = vals[it.entry];
                if (it.sub == 0) q.top = (q.top & 0xFFFF0000u) | v;
                else if (it.sub == 1) q.top = (q.top & 0x0000FFFFu) | (v << 16);
                else if (it.sub == 2) q.bot = (q.bot & 0xFFFF0000u) | v;
                else q.bot = (q.bot & 0x0000FFFFu) | (v << 16);
            }
            pc++;
        }
        if (ok) {
            q.ge = wadd(__ldg(raw2ev + (q.top >> 16)), __ldg(raw2ev + (q.bot & 0xFFFF))) / 2;
            q.dr = wsub(__ldg(raw2ev + (q.top & 0xFFFF)), q.ge);
            q.db = wsub(__ldg(raw2ev + (q.bot >> 16)), q.ge);
        }
        return q;
    };

    RowQ A = load_row(qr0 - 1);
    RowQ B = load_row(qr0);
    uint16_t *orow = out + (size_t)(2 * qr0) * w + x;
    for (int qr = qr0; qr < qr1; qr++) {
        const RowQ Cq = load_row(qr + 1);
        const int mr = median9_columns(A.dr, B.dr, Cq.dr);
        const int mb = median9_columns(A.db, B.db, Cq.db);
        uint32_t top = B.top, bot = B.bot;
        const int y = 2 * qr;
        if (x_inside && y >= 4 && y < h - 5 && B.ge >= 2 * MLVB_EV_RES) {
            const int er = wadd(B.ge, mr), eb = wadd(B.ge, mb);
            if (er > MLVB_EV_RES && eb > MLVB_EV_RES) {
                const uint32_t r = (uint32_t)(__ldg(ev2raw + clamp_ev(er)) + black) & 0xFFFFu;
                const uint32_t b = (uint32_t)(__ldg(ev2raw + clamp_ev(eb)) + black) & 0xFFFFu;
                top = (top & 0xFFFF0000u) | r;
                bot = (bot & 0x0000FFFFu) | (b << 16);
            }
        }
        if (STRIPES) {
            top = stripe_gain_fast(top & 0xFFFF, c0, black16, white16) | (stripe_gain_fast(top >> 16, c1, black16, white16) << 16);
            bot = stripe_gain_fast(bot & 0xFFFF, c0, black16, white16) | (stripe_gain_fast(bot >> 16, c1, black16, white16) << 16);
        }
        if (writer) {
            *reinterpret_cast<uint32_t *>(orow) = top;
            *reinterpret_cast<uint32_t *>(orow + w) = bot;
        }
        orow += 2 * w;
        A = B;
        B = Cq;
    }
}

}  // namespace

#include "fused_wide.cuh"

// Per-clip patch buckets for the fused kernels (built once, with the bad-pixel map).
struct FusedPatchPlan {
    int w = 0, h = 0, crop_x = 0, crop_y = 0;
    unsigned n_entries = 0;
    PatchItem *d_items = nullptr;
    unsigned *d_bucket_start = nullptr;
    WideItem *d_wide_items = nullptr;           // wide kernel: entries by (strip, quad row)
    unsigned *d_wide_row_start = nullptr;
    struct Vals { uint16_t *d = nullptr; size_t cap = 0; };
    std::map<cudaStream_t, Vals> vals;          // repaired values of the frames in flight, one buffer per stream
    ~FusedPatchPlan()
    {
        if (d_items) cudaFree(d_items);
        if (d_bucket_start) cudaFree(d_bucket_start);
        if (d_wide_items) cudaFree(d_wide_items);
        if (d_wide_row_start) cudaFree(d_wide_row_start);
        for (auto &kv : vals) if (kv.second.d) cudaFree(kv.second.d);
    }
};

static int build_patch_plan(const PixelList &list, int w, int h, int crop_x, int crop_y, FusedPatchPlan *plan)
{
    const int pw = w / 2, ph = h / 2;
    const int nstrips = ceil_div(pw, 30), nseg = ceil_div(ph, FS_ROWS);
    std::vector<std::vector<PatchItem>> buckets((size_t)nstrips * nseg);
    for (size_t m = 0; m < list.host.size(); m++) {
        const int x = list.host[m].x - crop_x, y = list.host[m].y - crop_y;
        if (!(x > 2 && x < w - 3 && y > 2 && y < h - 3)) continue;            // cs.c:317: only interior entries are repaired
        const int qc = x >> 1, qr = y >> 1, sub = (y & 1) * 2 + (x & 1);
        for (int s = std::max(qc / 30 - 1, 0); s <= std::min(qc / 30 + 1, nstrips - 1); s++) {
            const int lane = qc - 30 * s + 1;
            if (lane < 0 || lane > 31) continue;
            for (int g = std::max(qr / FS_ROWS - 1, 0); g <= std::min(qr / FS_ROWS + 1, nseg - 1); g++) {
                const int lo = g * FS_ROWS - 1, hi = std::min(g * FS_ROWS + FS_ROWS, ph);   // rows this segment loads
                if (qr < lo || qr > hi) continue;
                buckets[(size_t)g * nstrips + s].push_back(PatchItem{qr, (short)lane, (short)sub, (unsigned)m});
            }
        }
    }
    std::vector<unsigned> start(buckets.size() + 1, 0);
    std::vector<PatchItem> items;
    for (size_t b = 0; b < buckets.size(); b++) {
        std::stable_sort(buckets[b].begin(), buckets[b].end(), [](const PatchItem &a, const PatchItem &c) { return a.qrow < c.qrow; });
        start[b] = (unsigned)items.size();
        items.insert(items.end(), buckets[b].begin(), buckets[b].end());
    }
    start[buckets.size()] = (unsigned)items.size();
    plan->w = w; plan->h = h; plan->crop_x = crop_x; plan->crop_y = crop_y;
    plan->n_entries = (unsigned)list.host.size();
    MLVB_CUDA_OK(cudaMalloc(&plan->d_items, std::max<size_t>(items.size(), 1) * sizeof(PatchItem)));
    MLVB_CUDA_OK(cudaMalloc(&plan->d_bucket_start, start.size() * sizeof(unsigned)));
    if (!items.empty()) MLVB_CUDA_OK(cudaMemcpy(plan->d_items, items.data(), items.size() * sizeof(PatchItem), cudaMemcpyHostToDevice));
    MLVB_CUDA_OK(cudaMemcpy(plan->d_bucket_start, start.data(), start.size() * sizeof(unsigned), cudaMemcpyHostToDevice));

    // wide kernel: per strip and quad row (fused_wide.cuh::wide_build_items)
    std::vector<unsigned> wstart;
    std::vector<WideItem> witems;
    wide_build_items(list.host.size(), [&](size_t m, int &x, int &y) { x = list.host[m].x - crop_x; y = list.host[m].y - crop_y; },
                     w, h, witems, wstart);
    MLVB_CUDA_OK(cudaMalloc(&plan->d_wide_items, std::max<size_t>(witems.size(), 1) * sizeof(WideItem)));
    MLVB_CUDA_OK(cudaMalloc(&plan->d_wide_row_start, wstart.size() * sizeof(unsigned)));
    if (!witems.empty()) MLVB_CUDA_OK(cudaMemcpy(plan->d_wide_items, witems.data(), witems.size() * sizeof(WideItem), cudaMemcpyHostToDevice));
    MLVB_CUDA_OK(cudaMemcpy(plan->d_wide_row_start, wstart.data(), wstart.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
    return MLVB_OK;
}

// The wide kernel needs enough warp rows (frames x strips x quad rows) to occupy one persistent CTA per SM;
// MLVB_WIDE_MIN_ROWS overrides the threshold (tests run it on small batches).
static long long wide_min_rows(const mlvb_context *ctx)
{
    const char *e = getenv("MLVB_WIDE_MIN_ROWS");
    if (e && *e) return atoll(e);
    return 8LL * ctx->sm_count * FW_WARPS;
}

// Segment length of the wide kernel: items = frames x strips x segments are dealt round-robin to
// sm_count x 16 persistent warps; pick the split that loses least to the last partial round and to the two
// halo rows every segment re-reads.
static int wide_pick_segments(int nframes, int nstrips, int ph, int nwarps)
{
    int best = 1;
    double best_eff = 0.0;
    for (int nseg = 1; nseg <= std::max(1, ph / 8); nseg++) {
        const int rows = ceil_div(ph, nseg);
        if (ceil_div(ph, rows) != nseg) continue;
        const long long items = (long long)nframes * nstrips * nseg;
        const long long rounds = (items + nwarps - 1) / nwarps;
        const double eff = (double)items / (double)(rounds * nwarps) * rows / (rows + 2.0);
        if (eff > best_eff) { best_eff = eff; best = nseg; }
    }
    return best;
}

// Returns MLVB_OK when the fused kernel was enqueued, 1 when this call is not eligible (use the general
// path), < 0 on error.  Needs the clip's per-clip state to exist already.
int try_fused_single_iso(mlvb_context *ctx, const struct frame_headers *hdr, const FrameGeom &g, const mlvb_options &opts,
                         const char *mlv_filename, const void *d_payload, size_t payload_stride, size_t payload_bytes,
                         uint16_t *d_out, size_t out_stride_px, int nframes, cudaStream_t st)
{
    if (opts.chroma_smooth != 3 || opts.dual_iso || opts.fix_pattern_noise || opts.deflicker) return 1;
    if (g.bpp != 14 || (g.w % 16) || (g.h % 2) || g.w < 32 || g.h < 8 || g.black > MLVB_MAX_BLACK) return 1;
    if (hdr->file_hdr.videoClass & (MLVB_VIDEO_CLASS_FLAG_LJ92 | MLVB_VIDEO_CLASS_FLAG_LZMA)) return 1;
    if (((uintptr_t)d_payload % 16) || (payload_stride % 4) || ((uintptr_t)d_out % 4) || (out_stride_px % 2)) return 1;
    if (payload_bytes < mlvb_packed_bytes((uint32_t)g.npix, 14)) return MLVB_ERR_ARG;

    FusedParams P;
    memset(&P, 0, sizeof(P));
    std::shared_ptr<PixelList> bad;
    std::shared_ptr<FusedPatchPlan> plan;
    {
        std::lock_guard<std::mutex> lk(ctx->clip_mu);
        // focus-pixel maps are chains of neighbouring entries: general path
        for (auto &m : ctx->focus_maps)
            if (m.camera == hdr->idnt_hdr.cameraModel && m.rawi_width == hdr->rawi_hdr.raw_info.width &&
                m.rawi_height == hdr->rawi_hdr.raw_info.height && !m.entries.empty()) return 1;
        bool focus_known = false;
        for (auto &m : ctx->focus_maps)
            focus_known |= (m.camera == hdr->idnt_hdr.cameraModel && m.rawi_width == hdr->rawi_hdr.raw_info.width &&
                            m.rawi_height == hdr->rawi_hdr.raw_info.height);
        if (!focus_known) return 1;                                           // first frame of this camera: let the general path look for a map
        if (opts.fix_bad_pixels) {
            const uint64_t guid = hdr->file_hdr.fileGuid;
            BadPixelMap *bm = nullptr;
            for (auto &m : ctx->bad_maps)
                if (m.valid && guid && m.file_guid == guid && m.aggressive == (opts.fix_bad_pixels == 2)) bm = &m;
            if (!bm) return 1;                                                // map not detected yet
            bad = bm->list;
            if (bad && bad->nlevels > 1) return 1;                            // dependent entries: level-scheduled general path
            if (bad && bad->nlevels == 1) {
                auto &slot = bm->fused_plan;
                auto cur = std::static_pointer_cast<FusedPatchPlan>(slot);
                if (!cur || cur->w != g.w || cur->h != g.h || cur->crop_x != g.crop_x || cur->crop_y != g.crop_y) {
                    cur = std::make_shared<FusedPatchPlan>();
                    int rc = build_patch_plan(*bad, g.w, g.h, g.crop_x, g.crop_y, cur.get());
                    if (rc) return rc;
                    slot = cur;
                }
                plan = cur;
            }
        }
        if (opts.fix_stripes) {
            auto it = ctx->stripes.find(mlv_filename ? mlv_filename : "");
            if (it == ctx->stripes.end() || !it->second.computed) return 1;  // coefficients come from the first frame (general path)
            const StripeCoef &sc = it->second.coef;
            if (sc.needed) {
                for (int i = 0; i < 8; i++) { if (sc.coef[i] <= 0) return 1; P.coef[i] = sc.coef[i]; }
                P.stripes = 1;
            }
        }
    }
    P.packed = (const uint8_t *)d_payload; P.payload_stride = payload_stride;
    P.out = d_out; P.out_stride = out_stride_px;
    P.w = g.w; P.h = g.h; P.black = g.black;
    P.raw2ev = ctx->luts.raw2ev_base + (MLVB_MAX_BLACK - g.black);
    P.ev2raw = ctx->luts.ev2raw_pos;
    P.black16 = (uint16_t)g.black; P.white16 = (uint16_t)g.white;
    P.nstrips = ceil_div(g.w / 2, 30);
    if (plan) {
        const size_t need = (size_t)nframes * plan->n_entries * sizeof(uint16_t);
        uint16_t *d_vals = nullptr;
        {
            std::lock_guard<std::mutex> lk(ctx->clip_mu);
            FusedPatchPlan::Vals &v = plan->vals[st];
            if (need > v.cap) {
                if (v.d) cudaFree(v.d);
                v.d = nullptr; v.cap = 0;
                MLVB_CUDA_OK(cudaMalloc(&v.d, need));
                v.cap = need;
            }
            d_vals = v.d;
        }
        StageTimer t(ctx, ST_PIXFIX, st);
        patch_values_kernel<<<dim3(ceil_div(plan->n_entries, 128), nframes), 128, 0, st>>>(
            P.packed, payload_stride, g.w, g.h, g.black, P.raw2ev, P.ev2raw, bad->d_by_level, plan->n_entries, g.crop_x, g.crop_y,
            d_vals);
        ctx->launches += 1;
        P.items = plan->d_items; P.bucket_start = plan->d_bucket_start; P.vals = d_vals; P.n_entries = plan->n_entries;
    }
    // batches large enough to keep one persistent CTA per SM busy take the wide kernel (fused_wide.cuh)
    bool wide = ctx->ev2raw_octaves_ok && ctx->sm_count > 0 && (g.w % 64) == 0 && ((uintptr_t)d_out % 16) == 0 &&
                (out_stride_px % 8) == 0 && (payload_stride % 16) == 0 && !ctx->no_wide &&
                (long long)nframes * ceil_div(g.w, FW_STRIP_PX) * (g.h / 2) >= wide_min_rows(ctx);
    // (v - black) * gain fits 32 bits, and so does (v - black) * gain + (black << 16) (the high-half form of FW_GAIN_X)
    for (int i = 0; i < 8 && P.stripes; i++)
        wide = wide && P.coef[i] > 0 && P.coef[i] < (1 << 18) &&
               (16383LL - P.black16) * P.coef[i] + ((long long)P.black16 << 16) < (1LL << 32);
    if (wide) {
        WideParams Q;
        memset(&Q, 0, sizeof(Q));
        Q.packed = P.packed; Q.payload_stride = payload_stride; Q.out = d_out; Q.out_stride = out_stride_px;
        Q.w = g.w; Q.h = g.h; Q.black = g.black; Q.raw2ev = P.raw2ev; Q.ev2raw13 = ctx->luts.ev2raw_pos + 13 * MLVB_EV_RES;
        Q.black16 = P.black16; Q.white16 = P.white16;
        for (int i = 0; i < 8; i++) {
            Q.gain[i].coef = (unsigned)P.coef[i]; Q.gain[i].k1 = 0u - (unsigned)P.black16 * (unsigned)P.coef[i];
            Q.coefh[i] = (unsigned)P.coef[i] << 14;       // coef < 2^18 (checked above)
        }
        Q.k4 = 0u - 4u * (unsigned)P.black16;
        for (int i = 0; i < 8; i++) {
            Q.gainx[i].coef = (unsigned)P.coef[i];
            Q.gainx[i].kx = ((unsigned)P.black16 << 16) - (unsigned)P.black16 * (unsigned)P.coef[i];
        }
        Q.whitex = ((unsigned)P.white16 << 16) | 0xFFFFu;
        if (plan) { Q.items = plan->d_wide_items; Q.row_start = plan->d_wide_row_start; Q.vals = P.vals; Q.n_entries = plan->n_entries; }
        Q.nstrips = ceil_div(g.w, FW_STRIP_PX); Q.nframes = nframes;
        Q.one = 1; Q.mone = -1;
        Q.shr[0] = 1u << 14; Q.shr[1] = 1u << 17;
        Q.shl[0] = 1u << 14; Q.shl[1] = 1u << 10; Q.shl[2] = 1u << 6; Q.shl[3] = 1u << 2;
        // one contiguous run of a strip's quad rows per warp (nseg = 0); MLVB_WIDE_SEGMENTS=1 keeps the equal-segment split
        const long long all_rows = (long long)nframes * Q.nstrips * (g.h / 2);
        const int nwarps = ctx->sm_count * FW_WARPS;
        if (all_rows < (1LL << 30) && nwarps >= Q.nstrips && !ctx->wide_segments) {
            Q.nseg = 0;             // the kernel cuts each strip's column of nframes x h/2 rows among that strip's warps
            Q.seg_rows = 0;
        } else {
            Q.nseg = wide_pick_segments(nframes, Q.nstrips, g.h / 2, nwarps);
            Q.seg_rows = ceil_div(g.h / 2, Q.nseg);
        }
        // the opt-in to > 48 KB of dynamic shared memory is per device: set it with every launch (a context may live on any GPU)
        const bool unit01 = P.coef[0] == 65536 && P.coef[1] == 65536 && P.white16 > P.black16 + 64;
        auto launch = [&](auto kernel) {
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FW_SMEM_BYTES);
            kernel<<<ctx->sm_count, FW_THREADS, FW_SMEM_BYTES, st>>>(Q);
        };
        StageTimer t(ctx, ST_CHROMA, st);
        if (P.stripes && unit01) launch(fused3_wide_kernel<2>);
        else if (P.stripes) launch(fused3_wide_kernel<1>);
        else launch(fused3_wide_kernel<0>);
        ctx->launches += 1;
        ctx->path_count[1] += 1;
    } else {
        StageTimer t(ctx, ST_CHROMA, st);
        dim3 grid(ceil_div(P.nstrips, FS_WARPS), ceil_div(g.h / 2, FS_ROWS), nframes);
        if (P.stripes) fused3_strip_kernel<true><<<grid, FS_WARPS * 32, 0, st>>>(P);
        else fused3_strip_kernel<false><<<grid, FS_WARPS * 32, 0, st>>>(P);
        ctx->launches += 1;
        ctx->path_count[0] += 1;
    }
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}
