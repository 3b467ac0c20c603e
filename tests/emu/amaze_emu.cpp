// tests/emu/amaze_emu.cpp -- TEST INFRASTRUCTURE.  Runs the CUDA tile program of
// mlvfs_b200/csrc/amaze_tile.cuh on the host: one std::thread per CUDA thread, std::barrier for
// __syncthreads, so that the block-cooperative logic (pass order, sequential passes, over-run lanes) can
// be checked against the oracle without a GPU.  Built by tests/test_amaze_emu.py with
//   g++ -O1 -std=c++20 -ffp-contract=off -shared -fPIC
// `order` permutes which host thread plays which CUDA thread id, to shake out hidden intra-phase
// dependencies; `poison` fills the workspace with NaN first (a reused workspace holds another tile's data).
#include <barrier>
#include <cstdlib>
#include <memory>
#include <thread>
#include <vector>

#include "../../mlvfs_b200/csrc/amaze_tile.cuh"

namespace {
struct HostCtx {
    int tid, nthr;
    std::barrier<> *bar, *wbar;                    // block barrier; this thread's warp barrier (32 threads)
    void sync() { bar->arrive_and_wait(); }
    void syncwarp() { wbar->arrive_and_wait(); }
    void mark(int) {}
    void prefetch(const void *) {}
    float ld_stream(const float *p) { return *p; }
    void st_stream(float *p, float v) { *p = v; }
    void atomic_add(int *p, int v) { __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
};
}  // namespace

extern "C" void amaze_emu(const float *raw, float *red, float *green, float *blue, int stride, int width, int height,
                          int nthr, int order, int poison)
{
    std::vector<char> block(amaze::WS_BYTES + 64 * 1024);          // slack: the reference's bottom-border overrun stays inside
    for (int top = -16; top < height; top += amaze::TS - 32)
        for (int left = -16; left < width; left += amaze::TS - 32) {
            if (poison) memset(block.data(), 0xFF, block.size());
            amaze::Ws W = amaze::carve(block.data());
            amaze::Geom G = amaze::tile_geom(width, height, top, left);
            amaze::Shared S;
            std::barrier<> bar(nthr);
            std::vector<std::unique_ptr<std::barrier<>>> wbars;     // nthr is a multiple of 32 (like the CUDA launch)
            for (int wdx = 0; wdx < nthr / 32; wdx++) wbars.emplace_back(new std::barrier<>(32));
            std::vector<std::thread> th;
            for (int t = 0; t < nthr; t++) {
                const int tid = order == 0 ? t : order == 1 ? nthr - 1 - t : (t * 37 + 11) % nthr;
                th.emplace_back([&, tid]() {
                    HostCtx C{tid, nthr, &bar, wbars[tid / 32].get()};
                    amaze::tile_body(C, W, G, S, raw, red, green, blue, stride);
                });
            }
            for (auto &x : th) x.join();
        }
}
