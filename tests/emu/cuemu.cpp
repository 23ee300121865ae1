// tests/emu/cuemu.cpp — runtime of the kernel-logic emulator (see cuemu.h).
// TEST INFRASTRUCTURE ONLY; never part of the product library.
#define BSB_EMU 1
#include "cuemu.h"

#include <ucontext.h>
#include <xmmintrin.h>

#include <map>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim

namespace cuemu {

thread_local uint3 t_threadIdx, t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;

namespace {

struct NeedFibers {};

enum State { READY, WAIT_BLOCK, WAIT_WARP, DONE };

struct Fiber {
  ucontext_t ctx;
  State state;
  uint3 tid;
  int lin;
  int and_gen;
};

constexpr size_t kStack = 128 * 1024;
constexpr size_t kMaxDynSmem = 256 * 1024;

struct Worker {
  bool fiber_mode = false;
  ucontext_t sched;
  std::vector<Fiber> fibers;
  std::vector<char> stacks;
  Fiber* cur = nullptr;
  const std::function<void()>* body = nullptr;
  std::vector<unsigned> slots;      // per-thread shuffle slots
  std::vector<unsigned char> smem;  // dynamic shared memory
  int nthreads = 0;
  int and_acc[4] = {1, 1, 1, 1};
};
thread_local Worker W;

std::mutex g_mu;
std::map<const void*, bool> g_needs_fibers;
long g_launches = 0;

void fiber_entry() {
  (*W.body)();
  W.cur->state = DONE;
  swapcontext(&W.cur->ctx, &W.sched);
}

void yield_to_sched() {
  Fiber* f = W.cur;
  swapcontext(&f->ctx, &W.sched);
}

void run_block_fibers(dim3 block, const std::function<void()>& body) {
  const int n = (int)(block.x * block.y * block.z);
  W.nthreads = n;
  for (int q = 0; q < 4; ++q) W.and_acc[q] = 1;
  W.body = &body;
  W.fiber_mode = true;
  if ((int)W.fibers.size() < n) W.fibers.resize(n);
  if (W.stacks.size() < (size_t)n * kStack) W.stacks.resize((size_t)n * kStack);
  if ((int)W.slots.size() < n) W.slots.resize(n);
  int i = 0;
  for (unsigned z = 0; z < block.z; ++z)
    for (unsigned y = 0; y < block.y; ++y)
      for (unsigned x = 0; x < block.x; ++x, ++i) {
        Fiber& f = W.fibers[i];
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = W.stacks.data() + (size_t)i * kStack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, fiber_entry, 0);
        f.state = READY;
        f.tid = uint3{x, y, z};
        f.lin = i;
        f.and_gen = 0;
      }
  for (;;) {
    bool ran = false;
    int live = 0, wait_block = 0;
    for (int k = 0; k < n; ++k) {
      Fiber& f = W.fibers[k];
      if (f.state == READY) {
        W.cur = &f;
        t_threadIdx = f.tid;
        swapcontext(&W.sched, &f.ctx);
        ran = true;
      }
    }
    for (int k = 0; k < n; ++k) {
      State s = W.fibers[k].state;
      if (s != DONE) ++live;
      if (s == WAIT_BLOCK) ++wait_block;
    }
    if (live == 0) break;
    bool released = false;
    if (wait_block == live) {
      for (int k = 0; k < n; ++k) if (W.fibers[k].state == WAIT_BLOCK) W.fibers[k].state = READY;
      released = true;
    } else {
      for (int w0 = 0; w0 < n; w0 += 32) {
        int wl = 0, ww = 0;
        for (int k = w0; k < n && k < w0 + 32; ++k) {
          if (W.fibers[k].state != DONE) ++wl;
          if (W.fibers[k].state == WAIT_WARP) ++ww;
        }
        if (wl && ww == wl) {
          for (int k = w0; k < n && k < w0 + 32; ++k) if (W.fibers[k].state == WAIT_WARP) W.fibers[k].state = READY;
          released = true;
        }
      }
    }
    if (!ran && !released) {
      fprintf(stderr, "cuemu: barrier deadlock (divergent __syncthreads / shuffle?)\n");
      abort();
    }
  }
  W.fiber_mode = false;
}

}  // namespace

unsigned char* dyn_smem() {
  if (W.smem.size() < kMaxDynSmem) W.smem.resize(kMaxDynSmem);
  return W.smem.data();
}

void syncthreads() {
  if (!W.fiber_mode) throw NeedFibers{};
  W.cur->state = WAIT_BLOCK;
  yield_to_sched();
}

int syncthreads_and(int pred) {
  if (!W.fiber_mode) throw NeedFibers{};
  // four rotating accumulators: a fiber can run at most one call ahead of the slowest one, so the
  // slot two generations ahead can be reset while this generation is being read.
  const int gen = W.cur->and_gen++;
  if (!pred) W.and_acc[gen & 3] = 0;
  syncthreads();
  const int r = W.and_acc[gen & 3];
  W.and_acc[(gen + 2) & 3] = 1;
  return r;
}

void syncwarp() {
  if (!W.fiber_mode) throw NeedFibers{};
  W.cur->state = WAIT_WARP;
  yield_to_sched();
}

unsigned shfl_exchange(unsigned value, int arg, int mode, int width) {
  if (!W.fiber_mode) throw NeedFibers{};
  const int lin = W.cur->lin, lane = lin & 31, base = lin - lane;
  W.slots[lin] = value;
  syncwarp();
  const int seg = lane & ~(width - 1);
  int src = lane;
  switch (mode) {
    case 0: src = seg | (arg & (width - 1)); break;
    case 1: src = lane + arg; if (src >= seg + width) src = lane; break;
    case 2: src = lane - arg; if (src < seg) src = lane; break;
    case 3: src = lane ^ arg; if (src >= seg + width || src < seg) src = lane; break;
  }
  unsigned r = value;
  if (base + src < W.nthreads && W.fibers[base + src].state != DONE) r = W.slots[base + src];
  syncwarp();
  return r;
}

unsigned ballot(int pred) {
  if (!W.fiber_mode) throw NeedFibers{};
  const int lin = W.cur->lin, lane = lin & 31, base = lin - lane;
  W.slots[lin] = pred ? 1u : 0u;
  syncwarp();
  unsigned m = 0;
  for (int k = 0; k < 32 && base + k < W.nthreads; ++k)
    if (W.fibers[base + k].state != DONE && W.slots[base + k]) m |= 1u << k;
  syncwarp();
  return m;
}

long launches() { return g_launches; }

static void run_blocks(dim3 grid, dim3 block, size_t /*smem*/, const std::function<void()>& body,
                       bool* needs_fibers, unsigned long first, unsigned long last) {
  const unsigned csr = _mm_getcsr();
  _mm_setcsr(csr | 0x8040u);  // FTZ | DAZ, like nvcc -ftz=true
  t_blockDim = block;
  t_gridDim = grid;
  for (unsigned long b = first; b < last; ++b) {
    t_blockIdx = uint3{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((unsigned long)grid.x * grid.y))};
    bool done = false;
    if (!*needs_fibers) {
      try {
        for (unsigned z = 0; z < block.z; ++z)
          for (unsigned y = 0; y < block.y; ++y)
            for (unsigned x = 0; x < block.x; ++x) {
              t_threadIdx = uint3{x, y, z};
              body();
            }
        done = true;
      } catch (NeedFibers&) {
        *needs_fibers = true;
      }
    }
    if (!done) run_block_fibers(block, body);
  }
  _mm_setcsr(csr);
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body, const void* key) {
  if (smem_bytes > kMaxDynSmem) { fprintf(stderr, "cuemu: dynamic smem %zu too large\n", smem_bytes); abort(); }
  if (block.x * block.y * block.z > 1024) { fprintf(stderr, "cuemu: block too large\n"); abort(); }
  bool needs;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    static const bool trace = getenv("CUEMU_TRACE") != nullptr;     // name every kernel once, the first time it is launched
    if (trace && g_needs_fibers.find(key) == g_needs_fibers.end()) fprintf(stderr, "cuemu: launch %s\n", (const char*)key);
    needs = g_needs_fibers[key];
    ++g_launches;
  }
  const unsigned long total = (unsigned long)grid.x * grid.y * grid.z;
  static const int n_workers = [] {
    const char* e = getenv("CUEMU_THREADS");
    int n = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    return n < 1 ? 1 : (n > 16 ? 16 : n);
  }();
  if (n_workers == 1 || total < 8) {
    run_blocks(grid, block, smem_bytes, body, &needs, 0, total);
  } else {
    // blocks are independent by the CUDA programming model: spread them over host threads
    std::vector<std::thread> th;
    std::vector<char> flags(n_workers, needs ? 1 : 0);
    for (int w = 0; w < n_workers; ++w) {
      unsigned long a = total * w / n_workers, b = total * (w + 1) / n_workers;
      th.emplace_back([&, a, b, w] { bool nf = flags[w] != 0; run_blocks(grid, block, smem_bytes, body, &nf, a, b); flags[w] = nf; });
    }
    for (auto& t : th) t.join();
    for (char f : flags) needs = needs || f;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  g_needs_fibers[key] = needs;
}

}  // namespace cuemu
