// Runtime of the host emulation (see host_emul.h).  Test tool only; never linked into libvanerf_b200.so.
//
// A thread block is a set of ucontext fibers multiplexed on one OS thread (a barrier = yield until every member
// has arrived); blocks are distributed over a small pool of OS threads.
#ifdef VANERF_HOST_EMUL
#include "host_emul.h"

#include <ucontext.h>

#include <atomic>
#include <memory>

namespace emul {
thread_local uint3_ t_threadIdx, t_blockIdx;
thread_local Block* t_block = nullptr;
dim3 g_blockDim, g_gridDim;

namespace {
constexpr size_t kStack = 256 * 1024;

struct Fiber {
    ucontext_t ctx;
    std::unique_ptr<char[]> stack;
    bool done = false;
    uint3_ tid;
};

struct Runner {                       // one per OS thread
    ucontext_t sched;
    std::vector<Fiber> fibers;
    int cur = -1;
    const std::function<void()>* body = nullptr;
};
thread_local Runner* t_runner = nullptr;

void fiber_entry() {
    Runner* r = t_runner;
    (*r->body)();
    r->fibers[r->cur].done = true;
    swapcontext(&r->fibers[r->cur].ctx, &r->sched);
}
}  // namespace

void Barrier::wait() {
    Runner* r = t_runner;
    const int g = gen;
    if (++count == n) { count = 0; gen++; return; }
    while (g == gen) {
        swapcontext(&r->fibers[r->cur].ctx, &r->sched);
        t_threadIdx = r->fibers[r->cur].tid;
    }
}

static void run_block(Runner& r, Block& blk, dim3 block, int nt) {
    for (int t = 0; t < nt; ++t) {
        Fiber& f = r.fibers[t];
        f.done = false;
        f.tid.x = t % block.x;
        f.tid.y = (t / block.x) % block.y;
        f.tid.z = t / (block.x * block.y);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.get();
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &r.sched;
        makecontext(&f.ctx, fiber_entry, 0);
    }
    int remaining = nt;
    while (remaining > 0) {
        for (int t = 0; t < nt; ++t) {
            Fiber& f = r.fibers[t];
            if (f.done) continue;
            r.cur = t;
            t_threadIdx = f.tid;
            swapcontext(&r.sched, &f.ctx);
            if (f.done) remaining--;
        }
    }
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    g_blockDim = block;
    g_gridDim = grid;
    const int nt = (int)(block.x * block.y * block.z);
    const int nwarps = (nt + 31) / 32;
    const long nblocks = (long)grid.x * grid.y * grid.z;
    const int n_os = (int)std::max(1L, std::min<long>(nblocks, std::thread::hardware_concurrency()));
    std::atomic<long> next{0};
    auto worker = [&]() {
        Runner r;
        r.body = &body;
        r.fibers.resize(nt);
        for (auto& f : r.fibers) f.stack.reset(new char[kStack]);
        t_runner = &r;
        Block blk;
        blk.warp_bar = std::vector<Barrier>(nwarps);
        blk.xchg.assign((size_t)nwarps * 32, 0);
        std::vector<char> dyn(smem + 128);
        blk.dyn_smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(dyn.data()) + 127) & ~uintptr_t(127));
        t_block = &blk;
        for (long b = next.fetch_add(1); b < nblocks; b = next.fetch_add(1)) {
            blk.bar.init(nt);
            for (int w = 0; w < nwarps; ++w) blk.warp_bar[w].init((w == nwarps - 1) ? nt - 32 * w : 32);
            t_blockIdx.x = (unsigned)(b % grid.x);
            t_blockIdx.y = (unsigned)((b / grid.x) % grid.y);
            t_blockIdx.z = (unsigned)(b / ((long)grid.x * grid.y));
            run_block(r, blk, block, nt);
        }
        t_runner = nullptr;
        t_block = nullptr;
    };
    std::vector<std::thread> th;
    for (int i = 0; i < n_os; ++i) th.emplace_back(worker);
    for (auto& x : th) x.join();
}
}  // namespace emul
#endif
