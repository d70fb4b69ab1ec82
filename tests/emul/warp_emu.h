// TEST INFRASTRUCTURE ONLY.  A 32-lane lockstep "warp" on the CPU: every lane is a ucontext fiber;
// warp collectives (shuffle, ballot, syncwarp) are an exchange through a double-buffered mailbox
// plus a cooperative barrier.  Lets tests/emul run the marching-warp CUDA source
// (monodepth2.jl_b200/csrc/md2_march.cuh, compiled with MD2_WARP_EMU) without a GPU.
#pragma once
#include <stdlib.h>
#include <ucontext.h>

#include <functional>

namespace md2 {

struct WarpEmu {
    static constexpr int NL = 32;
    static constexpr size_t STACK = 512 * 1024;
    ucontext_t main_ctx, ctx[NL];
    char* stacks = nullptr;
    int cur = 0, ndone = 0;
    bool done[NL];
    unsigned int buf[2][NL];
    int phase[NL];
    long arrived = 0, passed[NL];
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
        if (w->ndone == NL) { setcontext(&w->main_ctx); }
        w->yield_from(me);   // never returns
        abort();
    }
    void yield_from(int me) {
        int nxt = me;
        do { nxt = (nxt + 1) % NL; } while (done[nxt] && nxt != me);
        if (nxt == me) return;
        cur = nxt;
        swapcontext(&ctx[me], &ctx[nxt]);
    }
    const unsigned int* exchange(unsigned int v) {
        const int me = cur;
        const int ph = phase[me];
        buf[ph][me] = v;
        phase[me] ^= 1;
        arrived++;
        const long target = (long)NL * (++passed[me]);
        while (arrived < target) yield_from(me);
        return buf[ph];
    }
    void run(std::function<void(int)> f) {
        body = f;
        current() = this;
        arrived = 0; ndone = 0;
        for (int l = 0; l < NL; ++l) {
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
    WarpEmu* w = WarpEmu::current();
    const unsigned int* b = w->exchange(v);
    return b[src_lane & 31];
}
inline unsigned int emu_ballot(int pred) {
    WarpEmu* w = WarpEmu::current();
    const unsigned int* b = w->exchange(pred ? 1u : 0u);
    unsigned int m = 0;
    for (int l = 0; l < 32; ++l) m |= (b[l] ? 1u : 0u) << l;
    return m;
}

}  // namespace md2
