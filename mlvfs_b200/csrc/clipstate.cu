// clipstate.cu -- EV tables, per-clip state (bad-pixel map, focus-pixel map, stripe coefficients)
// and the single-ISO stage sequencer.
//
// Per-clip state follows the reference's "first frame processed creates it" rule:
//   stripes coefficients  keyed by MLV path          main.c:980-997, stripes.c:29-83
//   bad-pixel map         keyed by fileGuid+aggr.    cs.c:233-254 (8-slot ring)
//   focus-pixel map       keyed by camera/raw size   cs.c:421-438, loaded from "<id>_<w>x<h>.fpm" in CWD
// Creation is once-only under ctx->clip_mu; the pixel statistics run on the GPU, only list
// bookkeeping (level schedule, the eight final pow() calls) happens on the host.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <unordered_map>

#include "context.cuh"

// ------------------------------------------------------------------ EV tables (main.c:128-196)

namespace {
int g_raw2ev[16384 + MLVB_MAX_BLACK];
double g_raw2evf[16384 + MLVB_MAX_BLACK];
int g_ev2raw[24 * MLVB_EV_RES];
std::once_flag g_lut_once;

void build_luts()
{
    memset(g_raw2ev, 0, sizeof(g_raw2ev));
    memset(g_raw2evf, 0, sizeof(g_raw2evf));
    for (int d = 0; d < 16384; d++) {
        const double e = log2((double)d) * MLVB_EV_RES;
        g_raw2evf[d + MLVB_MAX_BLACK] = e;
        g_raw2ev[d + MLVB_MAX_BLACK] = (d == 0) ? INT32_MIN : (int)e;    // (int)-inf on x86 == INT_MIN
    }
    for (int e = -10 * MLVB_EV_RES; e < 14 * MLVB_EV_RES; e++)
        g_ev2raw[e + 10 * MLVB_EV_RES] = (int)pow(2, (float)e / MLVB_EV_RES);
}
}  // namespace

const int *host_raw2ev_base() { std::call_once(g_lut_once, build_luts); return g_raw2ev; }
const double *host_raw2evf_base() { std::call_once(g_lut_once, build_luts); return g_raw2evf; }
const int *host_ev2raw_base() { std::call_once(g_lut_once, build_luts); return g_ev2raw; }

// ------------------------------------------------------------------ glibc rand()

void GlibcRand::seed(unsigned s)
{
    int32_t word = s ? (int32_t)s : 1;
    r[0] = (uint32_t)word;
    for (int i = 1; i < 31; i++) {
        const long hi = word / 127773, lo = word % 127773;
        word = (int32_t)(16807 * lo - 2836 * hi);
        if (word < 0) word += 2147483647;
        r[i] = (uint32_t)word;
    }
    f = 3; b = 0;
    for (int i = 0; i < 310; i++) next();
}

int GlibcRand::next()
{
    r[f] += r[b];
    const int out = (int)(r[f] >> 1);
    f = (f + 1) % 31; b = (b + 1) % 31;
    return out;
}

// ------------------------------------------------------------------ level schedule

PixelList::~PixelList()
{
    if (d_by_level) cudaFree(d_by_level);
    if (d_level_start) cudaFree(d_level_start);
    if (d_by_row) cudaFree(d_by_row);
    if (d_seg_start) cudaFree(d_seg_start);
    if (d_long_rows) cudaFree(d_long_rows);
}

int PixelList::upload(const FrameGeom *fg)
{
    level_start.assign(2, 0);
    nlevels = 0;
    has_wrapped = false;
    if (fg) {
        // cs.c:462-501: keep the entries that act in this frame, in file order
        const long long w = fg->w, h = fg->h;
        std::vector<PixelXY> kept;
        kept.reserve(host.size());
        for (const PixelXY &p : host) {
            const long long x = (long long)p.x - fg->crop_x, y = (long long)p.y - fg->crop_y, i = x + y * w;
            const bool interior = x > 2 && x < w - 3 && y > 2 && y < h - 3;
            if (!interior && !(i > 0 && i < w * h)) continue;
            if (!interior) {
                const bool hedge = (x >= w - 3 && x < w) || (x >= 0 && x <= 3);
                const bool vedge = (y >= h - 3 && y < h) || (y >= 0 && y <= 3);
                if (!hedge && !vedge) continue;                    // x outside the frame on an interior row: no rule applies
                if (x < 0 || x >= w) has_wrapped = true;
            }
            kept.push_back(p);
        }
        host.swap(kept);
    }
    const size_t n = host.size();
    if (n == 0) return MLVB_OK;
    // level(m) = 1 + max level over earlier entries in the +-3 cross stencil (and at the same site).  The stencil is
    // symmetric, so an entry also waits for earlier entries that READ its site.
    std::vector<unsigned> level(n, 0);
    std::unordered_map<uint64_t, unsigned> last_at;       // site -> latest entry index so far
    last_at.reserve(n * 2);
    auto key = [](int x, int y) { return ((uint64_t)(uint32_t)x << 32) | (uint32_t)y; };
    unsigned maxlevel = 0;
    for (size_t m = 0; m < n; m++) {
        unsigned lv = 0;
        if (fg) {
            // sites are linear indices: every rule reads i +- 1..3 and / or i +- (1..3) * w (interpolate_* and the
            // +-2 copies), also when the index wrapped into a neighbouring row
            const long long w = fg->w;
            const long long i = ((long long)host[m].x - fg->crop_x) + ((long long)host[m].y - fg->crop_y) * w;
            for (int d = -3; d <= 3; d++) {
                auto it = last_at.find((uint64_t)(i + d));
                if (it != last_at.end()) lv = std::max(lv, level[it->second] + 1);
                if (d != 0) {
                    it = last_at.find((uint64_t)(i + d * w));
                    if (it != last_at.end()) lv = std::max(lv, level[it->second] + 1);
                }
            }
            last_at[(uint64_t)i] = (unsigned)m;
        } else {
            const int x = host[m].x, y = host[m].y;
            for (int d = -3; d <= 3; d++) {
                auto it = last_at.find(key(x + d, y));
                if (it != last_at.end()) lv = std::max(lv, level[it->second] + 1);
                if (d != 0) {
                    it = last_at.find(key(x, y + d));
                    if (it != last_at.end()) lv = std::max(lv, level[it->second] + 1);
                }
            }
            last_at[key(x, y)] = (unsigned)m;
        }
        level[m] = lv;
        maxlevel = std::max(maxlevel, lv);
    }
    nlevels = maxlevel + 1;
    level_start.assign(nlevels + 1, 0);
    for (size_t m = 0; m < n; m++) level_start[level[m] + 1]++;
    for (unsigned l = 0; l < nlevels; l++) level_start[l + 1] += level_start[l];
    std::vector<PixelXY> sorted(n);
    std::vector<unsigned> cursor(level_start.begin(), level_start.end() - 1);
    for (size_t m = 0; m < n; m++) sorted[cursor[level[m]]++] = host[m];
    MLVB_CUDA_OK(cudaMalloc(&d_by_level, n * sizeof(PixelXY)));
    MLVB_CUDA_OK(cudaMalloc(&d_level_start, level_start.size() * sizeof(unsigned)));
    MLVB_CUDA_OK(cudaMemcpy(d_by_level, sorted.data(), n * sizeof(PixelXY), cudaMemcpyHostToDevice));
    MLVB_CUDA_OK(cudaMemcpy(d_level_start, level_start.data(), level_start.size() * sizeof(unsigned),
                            cudaMemcpyHostToDevice));

    // Row-wise form for the horizontal interpolator (dual ISO, cs.c:321-328 / 470-478 with dual_iso set):
    // an entry then reads and writes its own row only, so rows are independent and, inside a row whose
    // entries come in increasing x, so are runs further than 3 columns apart.  Stable counting sort by y.
    int ylo = host[0].y, yhi = host[0].y;
    for (size_t m = 0; m < n; m++) { ylo = std::min(ylo, host[m].y); yhi = std::max(yhi, host[m].y); }
    const size_t nrows = (size_t)(yhi - ylo) + 1;
    std::vector<unsigned> row_start(nrows + 1, 0);
    for (size_t m = 0; m < n; m++) row_start[(size_t)(host[m].y - ylo) + 1]++;
    for (size_t r = 0; r < nrows; r++) row_start[r + 1] += row_start[r];
    std::vector<PixelXY> by_row(n);
    {
        std::vector<unsigned> cur(row_start.begin(), row_start.end() - 1);
        for (size_t m = 0; m < n; m++) by_row[cur[(size_t)(host[m].y - ylo)]++] = host[m];
    }
    // rows with >= 128 entries (dual ISO with --really-bad-pix lists every bright-row pixel: one serial chain per
    // row) are walked by a warp in shared memory; the others are cut into independent segments, one thread each
    std::vector<unsigned> seg, longr;
    for (size_t r = 0; r < nrows; r++) {
        const unsigned lo = row_start[r], hi = row_start[r + 1];
        if (lo == hi) continue;
        if (hi - lo >= 128) { longr.push_back(lo); longr.push_back(hi); continue; }
        bool monotone = true;
        for (unsigned m = lo + 1; m < hi; m++) monotone &= by_row[m].x >= by_row[m - 1].x;
        unsigned first = lo;
        if (monotone)
            for (unsigned m = lo + 1; m < hi; m++)
                if (by_row[m].x - by_row[m - 1].x > 3) { seg.push_back(first); seg.push_back(m); first = m; }
        seg.push_back(first); seg.push_back(hi);
    }
    nseg = (unsigned)(seg.size() / 2);
    nlong = (unsigned)(longr.size() / 2);
    MLVB_CUDA_OK(cudaMalloc(&d_by_row, n * sizeof(PixelXY)));
    MLVB_CUDA_OK(cudaMalloc(&d_seg_start, std::max<size_t>(seg.size(), 2) * sizeof(unsigned)));
    MLVB_CUDA_OK(cudaMalloc(&d_long_rows, std::max<size_t>(longr.size(), 2) * sizeof(unsigned)));
    MLVB_CUDA_OK(cudaMemcpy(d_by_row, by_row.data(), n * sizeof(PixelXY), cudaMemcpyHostToDevice));
    if (!seg.empty()) MLVB_CUDA_OK(cudaMemcpy(d_seg_start, seg.data(), seg.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
    if (!longr.empty()) MLVB_CUDA_OK(cudaMemcpy(d_long_rows, longr.data(), longr.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
    return MLVB_OK;
}

int apply_pixel_list(mlvb_context *ctx, const PixelList &L, uint16_t *d_img, const FrameGeom &g, size_t frame_stride, int nframes,
                     int dual_iso, int edge_rules, cudaStream_t st)
{
    if (!L.nlevels) return MLVB_OK;
    if (dual_iso && !L.has_wrapped) {
        ctx->launches += (L.nseg > 0) + (L.nlong > 0);
        return launch_pixel_fix_rows(d_img, g.w, g.h, frame_stride, nframes, g.black, g.crop_x, g.crop_y, edge_rules, L.d_by_row,
                                     L.d_seg_start, L.nseg, L.d_long_rows, L.nlong, ctx->luts, ctx->ev2raw_octaves_ok, ctx->sm_count, st);
    }
    ctx->launches += 1 + (L.nlevels > 1);
    return launch_pixel_fix(d_img, g.w, g.h, frame_stride, nframes, g.black, g.crop_x, g.crop_y, dual_iso, edge_rules, L.d_by_level,
                            L.d_level_start, L.level_start.data(), L.nlevels, ctx->luts, st);
}

// ------------------------------------------------------------------ geometry

FrameGeom geom_from_headers(const struct frame_headers *hdr)
{
    FrameGeom g;
    g.w = hdr->rawi_hdr.xRes;
    g.h = hdr->rawi_hdr.yRes;
    g.bpp = hdr->rawi_hdr.raw_info.bits_per_pixel;
    g.black = hdr->rawi_hdr.raw_info.black_level;
    g.white = hdr->rawi_hdr.raw_info.white_level;
    g.crop_x = (hdr->vidf_hdr.panPosX + 7) & ~7;           // cs.c:225-226
    g.crop_y = hdr->vidf_hdr.panPosY & ~1;
    g.frame_size = hdr->rawi_hdr.raw_info.frame_size;
    g.npix = (size_t)g.w * g.h;
    return g;
}

int mlvb_context::ensure_scratch(size_t bytes)
{
    if (bytes <= scratch_cap) return MLVB_OK;
    if (d_scratch) cudaFree(d_scratch);
    d_scratch = nullptr; scratch_cap = 0;
    MLVB_CUDA_OK(cudaMalloc(&d_scratch, bytes));
    scratch_cap = bytes;
    return MLVB_OK;
}

int mlvb_context::ensure_stat(size_t bytes)
{
    if (bytes <= stat_cap) return MLVB_OK;
    if (d_stat) cudaFree(d_stat);
    d_stat = nullptr; stat_cap = 0;
    MLVB_CUDA_OK(cudaMalloc(&d_stat, bytes));
    stat_cap = bytes;
    return MLVB_OK;
}

// ------------------------------------------------------------------ per-clip state creation

// cs.c:233-312: look the map up by (fileGuid, aggressive); on a miss detect on this frame and take
// the next ring slot.  fileGuid == 0 never matches, so such clips re-detect every frame (A.4).
int get_bad_pixel_map(mlvb_context *ctx, const struct frame_headers *hdr, const FrameGeom &g, int aggressive,
                             const uint16_t *d_img, cudaStream_t st, std::shared_ptr<PixelList> *out)
{
    const uint64_t guid = hdr->file_hdr.fileGuid;
    for (auto &m : ctx->bad_maps)
        if (m.valid && guid && m.file_guid == guid && m.aggressive == aggressive) { *out = m.list; return MLVB_OK; }

    size_t fb, cb;
    const int nctas = badpix_detect_scratch_bytes(g.w, g.h, &fb, &cb);
    int rc = ctx->ensure_stat(fb + cb);
    if (rc) return rc;
    uint8_t *d_flags = (uint8_t *)ctx->d_stat;
    unsigned long long *d_counts = (unsigned long long *)((uint8_t *)ctx->d_stat + fb);
    rc = launch_badpix_detect_count(d_img, g.w, g.h, g.black, aggressive, ctx->luts, d_flags, d_counts, st);
    if (rc) return rc;
    ctx->launches += 2;
    unsigned long long total = 0;
    MLVB_CUDA_OK(cudaMemcpyAsync(&total, d_counts + nctas, sizeof(total), cudaMemcpyDeviceToHost, st));
    MLVB_CUDA_OK(cudaStreamSynchronize(st));

    auto list = std::make_shared<PixelList>();
    if (total) {
        PixelXY *d_list = nullptr;
        MLVB_CUDA_OK(cudaMalloc(&d_list, total * sizeof(PixelXY)));
        rc = launch_badpix_detect_scatter(d_flags, d_counts, g.w, g.h, g.crop_x, g.crop_y, d_list, st);
        ctx->launches += 1;
        list->host.resize(total);
        if (rc == MLVB_OK) {
            cudaError_t e = cudaMemcpyAsync(list->host.data(), d_list, total * sizeof(PixelXY), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) rc = MLVB_ERR_CUDA;
        }
        cudaFree(d_list);
        if (rc) return rc;
    }
    rc = list->upload();
    if (rc) return rc;

    BadPixelMap &slot = ctx->bad_maps[ctx->bad_map_cursor];
    ctx->bad_map_cursor = (ctx->bad_map_cursor + 1) % 8;
    slot.file_guid = guid; slot.aggressive = aggressive; slot.valid = true; slot.list = list;
    *out = list;
    return MLVB_OK;
}

// cs.c:355-438: "<cameraModel hex>_<raw width>x<raw height>.fpm" in the current directory, one "x y" per line
int get_focus_pixel_map(mlvb_context *ctx, const struct frame_headers *hdr, const FrameGeom &g, std::shared_ptr<PixelList> *out)
{
    const uint32_t cam = hdr->idnt_hdr.cameraModel;
    const int rw = hdr->rawi_hdr.raw_info.width, rh = hdr->rawi_hdr.raw_info.height;
    FocusPixelMap *fm = nullptr;
    for (auto &m : ctx->focus_maps)
        if (m.camera == cam && m.rawi_width == rw && m.rawi_height == rh) { fm = &m; break; }
    if (!fm) {
        FocusPixelMap nm;
        nm.camera = cam; nm.rawi_width = rw; nm.rawi_height = rh;
        char name[1024];
        snprintf(name, sizeof(name), "%x_%ix%i.fpm", cam, rw, rh);
        FILE *f = fopen(name, "r");
        if (f) {
            int x = 0, y = 0, ret;
            while ((ret = fscanf(f, "%i %i", &x, &y)) != EOF) {
                if (ret == 2) nm.entries.push_back(PixelXY{x, y});
                else break;
            }
            fclose(f);
        }
        ctx->focus_maps.push_back(std::move(nm));
        fm = &ctx->focus_maps.back();
    }
    out->reset();
    if (fm->entries.empty()) return MLVB_OK;
    for (auto &s : fm->schedules)
        if (s.w == g.w && s.h == g.h && s.crop_x == g.crop_x && s.crop_y == g.crop_y) { *out = s.list; return MLVB_OK; }
    auto list = std::make_shared<PixelList>();
    list->host = fm->entries;
    int rc = list->upload(&g);
    if (rc) return rc;
    if (fm->schedules.size() >= 8) fm->schedules.erase(fm->schedules.begin());   // panning clips: keep the latest offsets
    fm->schedules.push_back({g.w, g.h, g.crop_x, g.crop_y, list});
    *out = list;
    return MLVB_OK;
}

// stripes.c:143-248 on the GPU (statistics) + the eight pow() calls on the host
int compute_stripes(mlvb_context *ctx, const FrameGeom &g, const uint16_t *d_img, cudaStream_t st, StripeCoef *out)
{
    const int nctas = stripes_count_ctas(g.w, g.h);
    const size_t counts_bytes = ((size_t)nctas + 1) * sizeof(unsigned long long);
    const size_t hist_bytes = 8 * 65536 * sizeof(unsigned), tail_bytes = 256;
    unsigned long long total = 0;
    int rc = ctx->ensure_stat(counts_bytes + hist_bytes + tail_bytes);
    if (rc) return rc;
    unsigned long long *d_counts = (unsigned long long *)ctx->d_stat;
    unsigned *d_hist = (unsigned *)((uint8_t *)ctx->d_stat + counts_bytes);
    unsigned *d_num = (unsigned *)((uint8_t *)d_hist + hist_bytes);
    int *d_med = (int *)(d_num + 8);
    if (nctas) {
        rc = launch_stripes_count(d_img, g.w, g.h, g.black, g.white, d_counts, st);
        if (rc) return rc;
        ctx->launches += 2;
        MLVB_CUDA_OK(cudaMemcpyAsync(&total, d_counts + nctas, sizeof(total), cudaMemcpyDeviceToHost, st));
        MLVB_CUDA_OK(cudaStreamSynchronize(st));
    }
    // dither table: the next 2*total values of the process-wide rand() stream, reduced mod 1024
    std::vector<uint16_t> dither(2 * total + 2);
    for (size_t i = 0; i < 2 * total; i++) dither[i] = (uint16_t)(ctx->dither_rng.next() % 1024);
    uint16_t *d_dither = nullptr;
    MLVB_CUDA_OK(cudaMalloc(&d_dither, dither.size() * sizeof(uint16_t)));
    cudaError_t e = cudaMemcpyAsync(d_dither, dither.data(), dither.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, st);
    unsigned num[8] = {0};
    int med[8] = {0};
    if (e == cudaSuccess) {
        rc = launch_stripes_hist(d_img, g.w, g.h, g.black, g.white, d_counts, d_dither, d_hist, d_num, d_med, st);
        ctx->launches += 2;
        if (rc == MLVB_OK) {
            e = cudaMemcpyAsync(num, d_num, sizeof(num), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(med, d_med, sizeof(med), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
    }
    cudaFree(d_dither);
    if (e != cudaSuccess) { fprintf(stderr, "libmlvfs_b200: stripes statistics failed: %s\n", cudaGetErrorString(e)); return MLVB_ERR_CUDA; }
    if (rc) return rc;

    // stripes.c:219-246.  Groups with too few samples keep identity gain (the reference leaves
    // them uninitialised).
    for (int j = 0; j < 8; j++) out->coef[j] = 65536;
    for (int j = 2; j < 8; j++) {
        if ((int)num[j] < g.frame_size / 128) continue;
        const double ev = (double)(med[j] - 32768) / 32768;             // H2F
        out->coef[j] = (int)(pow(2, ev) * 65536);
    }
    out->needed = 0;
    for (int j = 0; j < 8; j++) {
        const double c = (double)out->coef[j] / 65536;
        if (c < 0.998 || c > 1.002) out->needed = 1;
    }
    return MLVB_OK;
}

// ------------------------------------------------------------------ single-ISO stage sequencer

int run_single_iso_chain(mlvb_context *ctx, const struct frame_headers *hdr, const FrameGeom &g,
                         const mlvb_options &opts, const char *mlv_filename, uint16_t *d_a, uint16_t *d_out,
                         size_t frame_stride, int nframes, int skip_chroma, int skip_pixfix, cudaStream_t st)
{
    int rc = MLVB_OK;
    // --- focus pixels, then bad pixels (main.c:966-973), in place on d_a
    std::shared_ptr<PixelList> focus, bad;
    if (!skip_pixfix) {
        std::lock_guard<std::mutex> lk(ctx->clip_mu);
        rc = get_focus_pixel_map(ctx, hdr, g, &focus);
        if (rc) return rc;
    }
    if (focus && focus->nlevels && g.black <= MLVB_MAX_BLACK) {
        StageTimer t(ctx, ST_PIXFIX, st);
        rc = launch_pixel_fix(d_a, g.w, g.h, frame_stride, nframes, g.black, g.crop_x, g.crop_y, 0, 1, focus->d_by_level,
                              focus->d_level_start, focus->level_start.data(), focus->nlevels, ctx->luts, st);
        if (rc) return rc;
        ctx->launches += 1 + (focus->nlevels > 1);
    }
    if (!skip_pixfix && opts.fix_bad_pixels && g.black <= MLVB_MAX_BLACK) {
        {
            std::lock_guard<std::mutex> lk(ctx->clip_mu);
            rc = get_bad_pixel_map(ctx, hdr, g, opts.fix_bad_pixels == 2, d_a, st, &bad);
            if (rc) return rc;
        }
        if (bad && bad->nlevels) {
            StageTimer t(ctx, ST_PIXFIX, st);
            rc = launch_pixel_fix(d_a, g.w, g.h, frame_stride, nframes, g.black, g.crop_x, g.crop_y, 0, 0, bad->d_by_level,
                                  bad->d_level_start, bad->level_start.data(), bad->nlevels, ctx->luts, st);
            if (rc) return rc;
            ctx->launches += 1 + (bad->nlevels > 1);
        }
    }

    // --- stripes state (main.c:980-989): known -> fuse into the chroma-smooth store
    StripeCoef sc{};
    bool have_sc = false;
    const std::string clip = mlv_filename ? mlv_filename : "";
    if (opts.fix_stripes) {
        std::lock_guard<std::mutex> lk(ctx->clip_mu);
        auto it = ctx->stripes.find(clip);
        if (it != ctx->stripes.end() && it->second.computed) { sc = it->second.coef; have_sc = true; }
    }

    // --- chroma smoothing (main.c:975-978); black > MAX_BLACK leaves the frame alone (cs.c:58)
    const bool do_cs = !skip_chroma && (opts.chroma_smooth == 2 || opts.chroma_smooth == 3 || opts.chroma_smooth == 5) &&
                       g.black <= MLVB_MAX_BLACK;
    uint16_t *cur = d_a;
    if (do_cs) {
        if (d_out == d_a) return MLVB_ERR_ARG;
        StageTimer t(ctx, ST_CHROMA, st);
        rc = launch_chroma_smooth_u16(d_a, d_out, g.w, g.h, frame_stride, nframes, g.black, opts.chroma_smooth, ctx->luts,
                                      (opts.fix_stripes && have_sc) ? &sc : nullptr, g.white, st);
        if (rc) return rc;
        ctx->launches += 1;
        cur = d_out;
    }

    // --- stripes (main.c:980-997)
    if (opts.fix_stripes) {
        if (!have_sc) {
            // first frame of the clip: statistics come from frame 0 AFTER the corrections above
            std::lock_guard<std::mutex> lk(ctx->clip_mu);
            StripesState &ss = ctx->stripes[clip];
            if (!ss.computed) {
                rc = compute_stripes(ctx, g, cur, st, &ss.coef);
                if (rc) return rc;
                ss.computed = true;
            }
            sc = ss.coef;
            rc = launch_stripes_apply(cur, g.w, g.npix, frame_stride, nframes, g.black, g.white, &sc, st);
            if (rc) return rc;
            ctx->launches += (sc.needed && g.w % 8 == 0);
        } else if (!do_cs) {
            StageTimer t(ctx, ST_STRIPES, st);
            rc = launch_stripes_apply(cur, g.w, g.npix, frame_stride, nframes, g.black, g.white, &sc, st);
            if (rc) return rc;
            ctx->launches += (sc.needed && g.w % 8 == 0);
        }
    }
    if (cur != d_out) {
        for (int f = 0; f < nframes; f++)
            MLVB_CUDA_OK(cudaMemcpyAsync(d_out + f * frame_stride, cur + f * frame_stride, g.npix * 2,
                                         cudaMemcpyDeviceToDevice, st));
    }
    return MLVB_OK;
}
