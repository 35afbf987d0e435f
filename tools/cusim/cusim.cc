// cusim scheduler: runs the blocks of a launch one after the other; the threads of a block are fibers
// (hand-rolled x86-64 context switch) that give up the CPU at barriers, warp collectives and
// __nanosleep.  See cuda_runtime.h in this directory.  DEVELOPMENT / TEST TOOL ONLY.
#include "cuda_runtime.h"

#include <sys/mman.h>

#include <cstdio>
#include <mutex>
#include <vector>

#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim

#if !defined(__x86_64__)
#error "cusim's context switch is written for x86-64"
#endif

extern "C" void cusim_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl cusim_switch
.hidden cusim_switch
.type cusim_switch,@function
cusim_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size cusim_switch,.-cusim_switch
)");

namespace cusim {

Fiber *cur = nullptr;
uint3 g_blockIdx;
dim3 g_blockDim, g_gridDim;

namespace {
enum { RUNNABLE = 0, DONE = 1, BLOCKED_BAR = 2, BLOCKED_WARP = 3 };
constexpr size_t kStack = 256 * 1024;

// A rendezvous in progress: the lanes of `mask` doing operation `kind`.  A warp can have several at once
// (e.g. the lanes of a ballot-derived sub-mask inside match_any while the others already wait in __syncwarp).
struct Slot {
    unsigned long long vals[32];
    int args[32];
    unsigned arrived = 0, mask = 0;
    int kind = -1;
};
constexpr int kSlots = 6;
struct Warp {
    Slot slots[kSlots];
    unsigned long long res[32];
    unsigned live = 0;
};

struct Block {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    int live = 0, bar_arrived = 0;
    int acc_or = 0, acc_and = 1, acc_cnt = 0, res_or = 0, res_and = 1, res_cnt = 0;
    unsigned long long progress = 0;
};

Block *blk = nullptr;
void *sched_sp = nullptr;
void (*g_thunk)(void *) = nullptr;
void *g_closure = nullptr;
char *g_stacks = nullptr;
size_t g_nstacks = 0;
std::mutex g_mu;

[[noreturn]] void die(const char *msg) {
    fprintf(stderr, "cusim: %s (block %u,%u,%u thread %d)\n", msg, g_blockIdx.x, g_blockIdx.y, g_blockIdx.z, cur ? cur->linear : -1);
    abort();
}

void to_scheduler() { cusim_switch(&cur->sp, sched_sp); }

void release_barrier() {
    blk->res_or = blk->acc_or; blk->res_and = blk->acc_and; blk->res_cnt = blk->acc_cnt;
    blk->acc_or = 0; blk->acc_and = 1; blk->acc_cnt = 0; blk->bar_arrived = 0;
    for (auto &f : blk->fibers) if (f.state == BLOCKED_BAR) f.state = RUNNABLE;
    blk->progress++;
}

void complete_warp(Warp &W, Slot &w, int warp_index) {
    const unsigned m = w.arrived;
    unsigned ballot = 0;
    if (w.kind == 4) for (int l = 0; l < 32; ++l) if ((m >> l) & 1) ballot |= (unsigned)(w.vals[l] & 1) << l;
    for (int l = 0; l < 32; ++l) {
        if (!((m >> l) & 1)) continue;
        unsigned long long r = 0;
        int src = l;
        switch (w.kind) {
        case 0: src = w.args[l] & 31; break;
        case 1: src = l - w.args[l]; if (src < 0) src = l; break;
        case 2: src = l + w.args[l]; if (src > 31) src = l; break;
        case 3: src = (l ^ w.args[l]) & 31; break;
        default: break;
        }
        if (w.kind <= 3) r = ((m >> src) & 1) ? w.vals[src] : w.vals[l];
        else if (w.kind == 4) r = ballot;
        else if (w.kind == 5) { unsigned mm = 0; for (int k = 0; k < 32; ++k) if (((m >> k) & 1) && w.vals[k] == w.vals[l]) mm |= 1u << k; r = mm; }
        W.res[l] = r;
    }
    w.arrived = 0; w.kind = -1;
    Fiber *base = blk->fibers.data() + (size_t)warp_index * 32;
    const int n = (int)blk->fibers.size() - warp_index * 32;
    for (int l = 0; l < 32 && l < n; ++l) if (((m >> l) & 1) && base[l].state == BLOCKED_WARP) base[l].state = RUNNABLE;
    blk->progress++;
}

void fiber_main() {
    g_thunk(g_closure);
    Fiber *f = cur;
    f->state = DONE;
    blk->live--;
    blk->progress++;
    Warp &w = blk->warps[f->warp];
    w.live &= ~(1u << f->lane);
    for (Slot &s : w.slots) if (s.arrived && (s.mask & w.live) == s.arrived) complete_warp(w, s, f->warp);
    if (blk->live > 0 && blk->bar_arrived == blk->live) release_barrier();
    to_scheduler();
    die("resumed a finished fiber");
}

void init_fiber(Fiber &f, char *stack_end) {
    uintptr_t top = reinterpret_cast<uintptr_t>(stack_end) & ~uintptr_t(15);
    void **a = reinterpret_cast<void **>(top - 16);        // return-address slot (16-byte aligned)
    *a = reinterpret_cast<void *>(&fiber_main);
    void **sp = a - 6;                                     // r15 r14 r13 r12 rbx rbp
    for (int i = 0; i < 6; ++i) sp[i] = nullptr;
    f.sp = sp;
}
}  // namespace

void yield() { to_scheduler(); }

void barrier(int pred, int *out_or, int *out_and, int *out_count) {
    Block *b = blk;
    b->acc_or |= pred ? 1 : 0; b->acc_and &= pred ? 1 : 0; b->acc_cnt += pred ? 1 : 0;
    b->bar_arrived++;
    if (b->bar_arrived == b->live) release_barrier();
    else { cur->state = BLOCKED_BAR; to_scheduler(); }
    if (out_or) *out_or = b->res_or;
    if (out_and) *out_and = b->res_and;
    if (out_count) *out_count = b->res_cnt;
}

unsigned long long warp_exchange(unsigned mask, unsigned long long v, int kind, int arg) {
    Fiber *f = cur;
    Warp &w = blk->warps[f->warp];
    mask &= w.live;
    if (!((mask >> f->lane) & 1)) die("warp collective: calling lane is not in the mask");
    Slot *s = nullptr;
    for (Slot &c : w.slots) if (c.arrived && c.mask == mask && c.kind == kind) { s = &c; break; }
    if (!s) {
        for (Slot &c : w.slots) if (!c.arrived) { s = &c; break; }
        if (!s) die("warp collective: too many different rendezvous in flight in one warp");
        s->mask = mask; s->kind = kind;
    }
    if ((s->arrived >> f->lane) & 1) die("warp collective: lane arrived twice");
    s->vals[f->lane] = v; s->args[f->lane] = arg; s->arrived |= 1u << f->lane;
    if (s->arrived == (s->mask & w.live)) complete_warp(w, *s, f->warp);
    else { f->state = BLOCKED_WARP; to_scheduler(); }
    return w.res[f->lane];
}

void run_grid(dim3 grid, dim3 block, size_t, void (*thunk)(void *), void *closure) {
    std::lock_guard<std::mutex> lock(g_mu);
    const size_t nt = (size_t)block.x * block.y * block.z;
    if (nt == 0 || nt > 1024) die("bad block size");
    if (g_nstacks < nt) {
        if (g_stacks) munmap(g_stacks, g_nstacks * kStack);
        g_stacks = static_cast<char *>(mmap(nullptr, nt * kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
        if (g_stacks == MAP_FAILED) die("mmap of fiber stacks failed");
        g_nstacks = nt;
    }
    g_thunk = thunk; g_closure = closure;
    g_blockDim = block; g_gridDim = grid;
    Block b;
    blk = &b;
    b.fibers.resize(nt);
    b.warps.resize((nt + 31) / 32);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_blockIdx = uint3{bx, by, bz};
                for (auto &w : b.warps) { for (Slot &s : w.slots) { s.arrived = 0; s.kind = -1; } w.live = 0; }
                for (size_t i = 0; i < nt; ++i) {
                    Fiber &f = b.fibers[i];
                    f.linear = (int)i; f.lane = (int)(i & 31); f.warp = (int)(i >> 5); f.state = RUNNABLE;
                    f.tid = uint3{(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y), (unsigned)(i / ((size_t)block.x * block.y))};
                    b.warps[f.warp].live |= 1u << f.lane;
                    init_fiber(f, g_stacks + (i + 1) * kStack);
                }
                b.live = (int)nt; b.bar_arrived = 0; b.acc_or = 0; b.acc_and = 1; b.acc_cnt = 0;
                unsigned long long last_progress = b.progress, idle_sweeps = 0;
                while (b.live > 0) {
                    bool ran = false;
                    for (size_t i = 0; i < nt; ++i) {
                        Fiber &f = b.fibers[i];
                        if (f.state != RUNNABLE) continue;
                        cur = &f; ran = true;
                        cusim_switch(&sched_sp, f.sp);
                    }
                    cur = nullptr;
                    if (!ran) die("deadlock: every live thread is blocked in a barrier or warp collective");
                    if (b.progress == last_progress) { if (++idle_sweeps > 2000000ull) die("deadlock: threads spin without progress"); }
                    else { last_progress = b.progress; idle_sweeps = 0; }
                }
            }
    blk = nullptr;
}

}  // namespace cusim
