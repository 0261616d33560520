// cuda_emul.cpp -- fiber scheduler of the TEST-ONLY SIMT emulator (see cuda_emul.h).
#include "cuda_emul.h"

#include <ucontext.h>

#include <vector>

namespace ts_emul {
dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
unsigned char *g_smem = nullptr;
unsigned long long g_launches = 0;

namespace {
struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = false;
};
ucontext_t g_sched;
Fiber *g_cur = nullptr;
const std::function<void()> *g_body = nullptr;
constexpr size_t kStack = 256 * 1024;
std::vector<char *> g_stack_pool;

void trampoline() {
    (*g_body)();
    g_cur->done = true;
    swapcontext(&g_cur->ctx, &g_sched);
}
}  // namespace

void syncthreads() { swapcontext(&g_cur->ctx, &g_sched); }

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
    g_launches++;
    const unsigned nt = block.x * block.y * block.z;
    while (g_stack_pool.size() < nt) g_stack_pool.push_back((char *)malloc(kStack));
    std::vector<Fiber> fibers(nt);
    std::vector<unsigned char> smem(smem_bytes + 64);
    g_blockDim = block;
    g_gridDim = grid;
    g_body = &body;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                g_blockIdx = dim3(bx, by, bz);
                std::fill(smem.begin(), smem.end(), (unsigned char)0xCD);
                g_smem = smem.data();
                for (unsigned t = 0; t < nt; t++) {
                    Fiber &f = fibers[t];
                    f.done = false;
                    f.stack = g_stack_pool[t];
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack;
                    f.ctx.uc_stack.ss_size = kStack;
                    f.ctx.uc_link = &g_sched;
                    makecontext(&f.ctx, trampoline, 0);
                }
                unsigned alive = nt;
                // TS_EMUL_SHUFFLE=<seed>: resume the threads of a block in a different pseudo-random order in
                // every barrier interval, so code that silently depends on thread order (a missing
                // __syncthreads) fails here instead of racing on the GPU.
                static const char *shuf = getenv("TS_EMUL_SHUFFLE");
                static uint64_t rng = shuf ? strtoull(shuf, nullptr, 10) * 2654435761u + 1 : 0;
                std::vector<unsigned> order(nt);
                for (unsigned t = 0; t < nt; t++) order[t] = t;
                while (alive) {
                    if (shuf)
                        for (unsigned t = nt - 1; t > 0; t--) {
                            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                            std::swap(order[t], order[(rng >> 33) % (t + 1)]);
                        }
                    for (unsigned oi = 0; oi < nt; oi++) {
                        const unsigned t = order[oi];
                        Fiber &f = fibers[t];
                        if (f.done) continue;
                        g_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        g_cur = &f;
                        swapcontext(&g_sched, &f.ctx);
                        if (f.done) alive--;
                    }
                }
            }
    g_smem = nullptr;
}
}  // namespace ts_emul
