// Runtime of the host emulation (see host_emul.h).  Test tool only; never linked into libvanerf_b200.so.
#ifdef VANERF_HOST_EMUL
#include "host_emul.h"

namespace emul {
thread_local uint3_ t_threadIdx, t_blockIdx;
thread_local Block* t_block = nullptr;
dim3 g_blockDim, g_gridDim;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    g_blockDim = block;
    g_gridDim = grid;
    const int nt = (int)(block.x * block.y * block.z);
    const int nwarps = (nt + 31) / 32;
    Block blk;
    blk.bar.init(nt);
    blk.warp_bar = std::vector<Barrier>(nwarps);
    for (int w = 0; w < nwarps; ++w) {
        int lanes = (w == nwarps - 1) ? nt - 32 * w : 32;
        blk.warp_bar[w].init(lanes);
    }
    blk.xchg.assign((size_t)nwarps * 32, 0);
    std::vector<char> dyn(smem + 64);
    blk.dyn_smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(dyn.data()) + 63) & ~uintptr_t(63));
    const long nblocks = (long)grid.x * grid.y * grid.z;
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; ++t) {
        th.emplace_back([&, t]() {
            t_block = &blk;
            t_threadIdx.x = t % block.x;
            t_threadIdx.y = (t / block.x) % block.y;
            t_threadIdx.z = t / (block.x * block.y);
            for (long b = 0; b < nblocks; ++b) {
                t_blockIdx.x = (unsigned)(b % grid.x);
                t_blockIdx.y = (unsigned)((b / grid.x) % grid.y);
                t_blockIdx.z = (unsigned)(b / ((long)grid.x * grid.y));
                body();
                blk.bar.wait();     // all threads leave block b before anyone enters b+1 (static __shared__ reuse)
            }
        });
    }
    for (auto& x : th) x.join();
}
}  // namespace emul
#endif
