// TEST INFRASTRUCTURE ONLY.  A lockstep thread block on the CPU: every lane of every warp is a
// ucontext fiber; warp collectives (shuffle, ballot) are an exchange through a double-buffered
// per-warp mailbox plus a cooperative barrier, named block barriers (bar.sync / bar.arrive) are
// arrival counters with a generation number.  Lets tests/emul run the marching-warp CUDA source
// (monodepth2.jl_b200/csrc/md2_march.cuh, compiled with MD2_WARP_EMU) without a GPU.
#pragma once
#include <stdlib.h>
#include <ucontext.h>

#include <functional>

namespace md2 {

struct WarpEmu {
    static constexpr int MAXW = 2, NL = 32 * MAXW, NBAR = 16;
    static constexpr size_t STACK = 512 * 1024;
    ucontext_t main_ctx, ctx[NL];
    char* stacks = nullptr;
    int nthreads = 32, cur = 0, ndone = 0;
    bool done[NL];
    unsigned int buf[MAXW][2][32];
    int phase[NL];
    long arrived[MAXW], passed[NL];
    int bar_count[NBAR];
    long bar_gen[NBAR];
    std::function<void(int)> body;

    WarpEmu() { stacks = (char*)malloc(STACK * NL); }
    ~WarpEmu() { free(stacks); }

    static WarpEmu*& current() { static WarpEmu* w = nullptr; return w; }

    static void trampoline() {
        WarpEmu* w = current();
        const int me = w->cur;
        w->body(me);
        w->done[me] = true;
        w->ndone++;
        if (w->ndone == w->nthreads) { setcontext(&w->main_ctx); }
        w->yield_from(me);   // never returns
        abort();
    }
    void yield_from(int me) {
        int nxt = me;
        do { nxt = (nxt + 1) % nthreads; } while (done[nxt] && nxt != me);
        if (nxt == me) return;
        cur = nxt;
        swapcontext(&ctx[me], &ctx[nxt]);
    }
    // all 32 lanes of the calling warp deposit a value; returns the warp's mailbox
    const unsigned int* exchange(unsigned int v) {
        const int me = cur, w = me >> 5, lane = me & 31;
        const int ph = phase[me];
        buf[w][ph][lane] = v;
        phase[me] ^= 1;
        arrived[w]++;
        const long target = 32L * (++passed[me]);
        while (arrived[w] < target) yield_from(me);
        return buf[w][ph];
    }
    // named barrier: completes when `count` threads have arrived (bar.sync waits, bar.arrive does not)
    void barrier(int id, int count, int wait) {
        const int me = cur;
        const long gen = bar_gen[id];
        if (++bar_count[id] == count) { bar_count[id] = 0; bar_gen[id]++; return; }
        if (wait)
            while (bar_gen[id] == gen) yield_from(me);
    }
    void run(int threads, std::function<void(int)> f) {
        body = f;
        nthreads = threads;
        current() = this;
        ndone = 0;
        for (int w = 0; w < MAXW; ++w) arrived[w] = 0;
        for (int b = 0; b < NBAR; ++b) { bar_count[b] = 0; bar_gen[b] = 0; }
        for (int l = 0; l < threads; ++l) {
            done[l] = false; phase[l] = 0; passed[l] = 0;
            getcontext(&ctx[l]);
            ctx[l].uc_stack.ss_sp = stacks + STACK * l;
            ctx[l].uc_stack.ss_size = STACK;
            ctx[l].uc_link = nullptr;
            makecontext(&ctx[l], (void (*)())trampoline, 0);
        }
        cur = 0;
        swapcontext(&main_ctx, &ctx[0]);
        current() = nullptr;
    }
};

inline unsigned int emu_xchg(unsigned int v, int src_lane) {
    const unsigned int* b = WarpEmu::current()->exchange(v);
    return b[src_lane & 31];
}
inline unsigned int emu_ballot(int pred) {
    const unsigned int* b = WarpEmu::current()->exchange(pred ? 1u : 0u);
    unsigned int m = 0;
    for (int l = 0; l < 32; ++l) m |= (b[l] ? 1u : 0u) << l;
    return m;
}
inline void emu_bar(int id, int count, int wait) { WarpEmu::current()->barrier(id, count, wait); }
// mbarrier objects: bars[idx] holds (completed phases << 32) | pending arrivals; 32 arrivals per phase
inline void emu_mb_init(unsigned long long* bars, int idx) { bars[idx] = 0ull; }
inline void emu_mb_arrive(unsigned long long* bars, int idx) {
    unsigned long long v = bars[idx] + 1ull;
    if ((v & 0xffffffffull) == 32ull) v = ((v >> 32) + 1ull) << 32;
    bars[idx] = v;
}
inline void emu_mb_wait(unsigned long long* bars, int idx, int parity) {
    WarpEmu* w = WarpEmu::current();
    while ((int)((bars[idx] >> 32) & 1ull) == parity) w->yield_from(w->cur);
}

}  // namespace md2
