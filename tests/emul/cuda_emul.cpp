// cuda_emul.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emul.h).
// Fiber scheduler: one ucontext fiber per CUDA thread, blocks distributed over
// host threads.
#include "cuda_emul.h"

namespace qi_emul {

thread_local Fiber* cur = nullptr;

static const size_t kStackBytes = 96 * 1024;

struct Worker {
    std::vector<unsigned char> stacks;
    std::vector<unsigned char> smem;
};

static void fiber_entry() {
    Fiber* f = cur;
    (*f->blk->body)();
    f->done = true;
    BlockState* b = f->blk;
    b->live--;
    b->warps[f->linear >> 5].live--;
    // A thread that exits may complete a barrier the others are waiting on.
    if (b->live > 0 && b->bar_arrived >= b->live) { b->bar_arrived = 0; b->bar_gen++; }
    WarpState& w = b->warps[f->linear >> 5];
    if (w.live > 0 && w.arrived >= w.live) { w.arrived = 0; w.gen++; }
    if (w.live > 0 && w.arrived2 >= w.live) { w.arrived2 = 0; w.gen2++; }
    swapcontext(&f->ctx, &b->sched);
}

void yield() {
    Fiber* f = cur;
    swapcontext(&f->ctx, &f->blk->sched);
}

static void run_block(Worker& wk, dim3 grid, dim3 block, size_t smem_bytes, uint3 bid,
                      const std::function<void()>& body) {
    BlockState b;
    b.bid = bid; b.bdim = block; b.gdim = grid;
    b.nthreads = (int)(block.x * block.y * block.z);
    b.live = b.nthreads;
    b.body = &body;
    if (wk.smem.size() < smem_bytes + 256) wk.smem.resize(smem_bytes + 256);
    // 128-byte aligned dynamic shared memory
    uintptr_t sp = reinterpret_cast<uintptr_t>(wk.smem.data());
    sp = (sp + 127) & ~uintptr_t(127);
    b.smem = reinterpret_cast<unsigned char*>(sp);
    if (wk.stacks.size() < kStackBytes * (size_t)b.nthreads) wk.stacks.resize(kStackBytes * (size_t)b.nthreads);
    b.fibers.resize(b.nthreads);
    b.warps.resize((b.nthreads + 31) / 32);
    for (int i = 0; i < b.nthreads; ++i) {
        Fiber& f = b.fibers[i];
        f.linear = i;
        f.tid.x = i % block.x;
        f.tid.y = (i / block.x) % block.y;
        f.tid.z = i / (block.x * block.y);
        f.blk = &b;
        f.done = false;
        b.warps[i >> 5].live++;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = wk.stacks.data() + kStackBytes * (size_t)i;
        f.ctx.uc_stack.ss_size = kStackBytes;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, fiber_entry, 0);
    }
    while (b.live > 0) {
        for (int i = 0; i < b.nthreads; ++i) {
            Fiber& f = b.fibers[i];
            if (f.done) continue;
            cur = &f;
            swapcontext(&b.sched, &f.ctx);
        }
    }
    cur = nullptr;
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    if (nblocks == 0) return;
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    const char* env = std::getenv("QI_EMUL_THREADS");
    if (env) hw = (unsigned)std::max(1, std::atoi(env));
    size_t nworkers = std::min<size_t>(hw, nblocks);
    std::atomic<size_t> next{0};
    auto work = [&]() {
        Worker wk;
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= nblocks) break;
            uint3 bid;
            bid.x = (unsigned)(i % grid.x);
            bid.y = (unsigned)((i / grid.x) % grid.y);
            bid.z = (unsigned)(i / ((size_t)grid.x * grid.y));
            run_block(wk, grid, block, smem_bytes, bid, body);
        }
    };
    if (nworkers <= 1) { work(); return; }
    std::vector<std::thread> th;
    for (size_t w = 0; w < nworkers; ++w) th.emplace_back(work);
    for (auto& t : th) t.join();
}

}  // namespace qi_emul
