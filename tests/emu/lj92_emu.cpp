// tests/emu/lj92_emu.cpp -- TEST INFRASTRUCTURE.  Runs the cooperative programs of
// mlvfs_b200/csrc/lj92_core.cuh (parallel LJ92 decode) on the host: one std::thread per CUDA thread,
// std::barrier for __syncthreads / __syncwarp, an exchange array for warp shuffles, so that the
// block-cooperative logic (subsequence synchronisation, boundary resolution, prefix sums, the skewed
// prediction wavefront and its strip pipeline) can be checked against the oracle without a GPU.
// Built by tests/test_lj92_emu.py with  g++ -O1 -std=c++20 -shared -fPIC -pthread.
//
//   flags bit 0: skip the parallel re-synchronisation pass (dec_body<1> over all blocks), so that
//                resolve_body alone has to repair every block boundary;
//   flags bit 1: run the blocks of each pass in reverse order;
//   flags bit 2: predict with the wavefront even where predictor 6 separates.
#include <atomic>
#include <barrier>
#include <cstdlib>
#include <memory>
#include <thread>
#include <vector>

#include "../../mlvfs_b200/csrc/lj92_core.cuh"

namespace {

using namespace lj92;

struct WarpX {
    std::barrier<> bar{32};
    uint32_t x[32];
};

struct HostCtx {
    int tid, nthr;
    std::barrier<> *bar;
    std::atomic<int> *orflag;        // two counters, alternating
    int phase = 0;
    WarpX *wx = nullptr;
    void sync() { bar->arrive_and_wait(); }
    int sync_or(int p)
    {
        std::atomic<int> &f = orflag[phase & 1];
        if (p) f.store(1);
        bar->arrive_and_wait();
        const int r = f.load();
        bar->arrive_and_wait();
        if (tid == 0) f.store(0);
        bar->arrive_and_wait();
        phase++;
        return r;
    }
    void syncwarp() { wx->bar.arrive_and_wait(); }
    int all(int p)
    {
        wx->x[tid & 31] = (uint32_t)p;
        wx->bar.arrive_and_wait();
        int a = 1;
        for (int i = 0; i < 32; i++) a &= wx->x[i] != 0;
        wx->bar.arrive_and_wait();
        return a;
    }
    uint32_t atomic_add(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
    void store16(uint16_t *p, const uint32_t w[4]) { memcpy(p, w, 16); }
    uint32_t load_volatile32(const uint32_t *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
    void store_volatile32(uint32_t *p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
    void pause() { std::this_thread::yield(); }
    uint32_t shfl_up(uint32_t v, int d)
    {
        const int lane = tid & 31;
        wx->x[lane] = v;
        wx->bar.arrive_and_wait();
        const uint32_t r = lane >= d ? wx->x[lane - d] : v;
        wx->bar.arrive_and_wait();
        return r;
    }
    uint32_t shfl(uint32_t v, int src)
    {
        wx->x[tid & 31] = v;
        wx->bar.arrive_and_wait();
        const uint32_t r = wx->x[src];
        wx->bar.arrive_and_wait();
        return r;
    }
    void set_status(int *p, int v) { int z = 0; __atomic_compare_exchange_n(p, &z, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); }
    void load64(const uint16_t *p, uint32_t nx[16]) { memcpy(nx, p, 64); }
    void store64(uint16_t *p, const uint32_t o[16]) { memcpy(p, o, 64); }
};

template <class Fn>
void run_block(int nthr, Fn fn)
{
    std::barrier<> bar(nthr);
    std::atomic<int> orflag[2];
    orflag[0] = 0; orflag[1] = 0;
    std::vector<std::unique_ptr<WarpX>> wx;
    for (int w = 0; w < (nthr + 31) / 32; w++) wx.emplace_back(new WarpX);
    std::vector<std::thread> th;
    for (int t = 0; t < nthr; t++)
        th.emplace_back([&, t]() {
            HostCtx C{t, nthr, &bar, orflag, 0, wx[t >> 5].get()};
            fn(C);
        });
    for (auto &x : th) x.join();
}

}  // namespace

// payload = uint32 size + JPEG stream.  out: W*H samples, untiled.  Returns the frame status.
// stats[0] = decode blocks, stats[1] = blocks repaired by resolve_body, stats[2] = clean bytes.
extern "C" int lj92_emu_decode(const uint8_t *payload, size_t payload_bytes, int W, int H, uint16_t *out, int flags,
                               int predict_warps, int predict_parts, unsigned *stats)
{
    const size_t npix = (size_t)W * H;
    const Layout L = make_layout(payload_bytes, npix);
    std::vector<char> scratch(L.frame_stride + 64, (char)0xA5);              // poisoned like fresh device memory
    char *base = scratch.data();
    base += (64 - ((uintptr_t)base & 63)) & 63;
    FrameWork F = frame_work(base, L, 0);

    // 0. headers (lj92_parse_kernel)
    int rc = parse_headers(payload + 4, (int)(payload_bytes - 4), *F.T);
    if (rc == ST_OK && (long long)F.T->lw * F.T->lh != (long long)W * H) rc = ST_HEADER;
    F.T->status = rc;
    if (rc != ST_OK) return rc;
    for (int i = 0; i < (1 << LUT_BITS); i++) F.T->lut[i] = lut_entry(*F.T, i);
    for (int i = 0; i < (1 << LUT1_BITS); i++) F.T->lut1[i] = lut_entry(*F.T, i, LUT1_BITS);

    // 1. unstuff (count / scan / scatter kernels, serialised per 16-byte thread)
    size_t clean = 0;
    for (size_t q = 0; q < payload_bytes; q += 16) {
        uint8_t b[16];
        for (int j = 0; j < 16; j++) b[j] = q + j < payload_bytes ? payload[q + j] : 0;
        const unsigned m = keep_mask16(payload, b, (long long)q, 4ll + F.T->scan_off, (long long)payload_bytes);
        for (int j = 0; j < 16; j++) if (m >> j & 1) F.clean[clean++] = b[j];
    }
    F.T->clean_bytes = (unsigned)clean;

    const unsigned ncta = (unsigned)(((unsigned long long)clean * 8 + CHUNK_BITS - 1) / CHUNK_BITS);
    auto order = [&](unsigned i) { return (flags & 2) ? ncta - 1 - i : i; };
    std::unique_ptr<DecShared> S(new DecShared);

    // 2. synchronise
    for (unsigned i = 0; i < ncta; i++)
        run_block(DEC_THREADS, [&](HostCtx &C) { dec_body<0>(C, *S, F, order(i), (uint32_t)npix); });
    if (!(flags & 1))
        for (unsigned i = 0; i < ncta; i++)
            run_block(DEC_THREADS, [&](HostCtx &C) { dec_body<1>(C, *S, F, order(i), (uint32_t)npix); });
    unsigned stale = 0;
    for (unsigned b = 1; b < ncta; b++) stale += F.sub_end[b * DEC_THREADS - 1] != F.cta_in[b];
    // 3. resolve + index
    run_block(DEC_THREADS, [&](HostCtx &C) { resolve_body(C, *S, F, (uint32_t)npix); });
    // 4. write (+ boundary words cleared for the prediction pass)
    for (unsigned i = 0; i < ncta; i++)
        run_block(DEC_THREADS, [&](HostCtx &C) {
            clear_boundary(C, F, L, order(i), ncta);
            dec_body<2>(C, *S, F, order(i), (uint32_t)npix);
        });
    // 5b. predictor 6 separated (lj92_row_kernel + lj92_col_kernel<0..2>, restated serially); flags bit 2
    //     forces the wavefront instead
    const bool sep = !(flags & 4) && separable(*F.T, W) && W % 4 == 0 && H % 2 == 0;
    if (sep) {
        for (int r = 0; r < H; r++) {
            int U = r == 0 ? 1 << (F.T->bits - 1) : 0;
            for (int c = 0; c < W; c++) {
                U = row_step(r == 0, U, (int)(int16_t)F.tiled[(size_t)r * W + c]);
                F.tiled[(size_t)r * W + c] = (uint16_t)U;
            }
        }
        const int nch = (H + CH_ROWS - 1) / CH_ROWS;
        std::vector<uint16_t> sums((size_t)nch * W, 0);
        for (int ch = 0; ch < nch; ch++)
            for (int y = ch * CH_ROWS; y < H && y < (ch + 1) * CH_ROWS; y++)
                for (int x = 0; x < W; x++) sums[(size_t)ch * W + x] += F.tiled[(size_t)y * W + x];
        std::vector<uint16_t> run(W, 0);
        for (int ch = 0; ch < nch; ch++)
            for (int x = 0; x < W; x++) { const uint16_t v = sums[(size_t)ch * W + x]; sums[(size_t)ch * W + x] = run[x]; run[x] += v; }
        for (int ch = 0; ch < nch; ch++)
            for (int x = 0; x < W; x++) {
                uint16_t acc = sums[(size_t)ch * W + x];
                for (int y = ch * CH_ROWS; y < H && y < (ch + 1) * CH_ROWS; y++) {
                    acc += F.tiled[(size_t)y * W + x];
                    const int dy = 2 * y < H ? 2 * y : 2 * y - H + 1, dx = 2 * x < W ? 2 * x : 2 * x - W + 1;
                    out[(size_t)dy * W + dx] = acc;
                }
            }
        if (stats) { stats[0] = ncta; stats[1] = stale; stats[2] = (unsigned)clean; }
        return F.T->status;
    }
    // 5. predict: `predict_parts` blocks run concurrently and take strip groups by ticket
    {
        const int parts = predict_parts > 0 ? predict_parts : 1;
        std::vector<std::vector<uint16_t>> rings(parts, std::vector<uint16_t>((size_t)predict_warps * RING_ELEMS));
        std::vector<uint32_t> tickets(parts);
        std::vector<std::thread> blocks;
        for (int p = 0; p < parts; p++)
            blocks.emplace_back([&, p]() {
                run_block(predict_warps * 32, [&](HostCtx &C) { predict_body(C, rings[p].data(), &tickets[p], F, false, W); });
            });
        for (auto &b : blocks) b.join();
    }
    if (stats) { stats[0] = ncta; stats[1] = stale; stats[2] = (unsigned)clean; }
    if (F.T->status != ST_OK) return F.T->status;
    // 6. untile (lj92_untile_kernel's gather form for even sizes, the scatter form otherwise)
    if (W % 2 == 0 && H % 2 == 0) {
        for (int dy = 0; dy < H; dy++)
            for (int dx = 0; dx < W; dx++) {
                const int y = (dy >> 1) + (dy & 1) * (H >> 1), x = (dx >> 1) + (dx & 1) * (W >> 1);
                out[(size_t)dy * W + dx] = F.tiled[(size_t)y * W + x];
            }
    } else {
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                out[(size_t)((2 * y) % H + (2 * y) / H) * W + (2 * x) % W + (2 * x) / W] = F.tiled[(size_t)y * W + x];
    }
    return ST_OK;
}
