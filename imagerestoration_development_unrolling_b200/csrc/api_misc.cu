// api_misc.cu - version / device / error plumbing of the C ABI.
#include "common.cuh"

#ifdef GLRGTV_EMU
thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
alignas(64) static thread_local float emu_smem_storage[64 * 1024];
thread_local float* emu_smem = emu_smem_storage;

// ---- cooperative fibers: one per CUDA thread of the block being emulated (ucontext, round-robin) ----
#include <ucontext.h>
#include <vector>
thread_local bool emu_fiber_mode = false;
namespace {
constexpr size_t kStack = 256 * 1024;
struct FiberRt {
    std::vector<ucontext_t> ctx;
    std::vector<char*> stacks;
    std::vector<char> done;
    ucontext_t sched;
    unsigned n = 0, cur = 0, alive = 0;
    unsigned bar_count = 0, bar_gen = 0;
    std::vector<float> buf[2];
    std::vector<unsigned> cnt, gen;
    void (*fn)(void*) = nullptr;
    void* arg = nullptr;
};
thread_local FiberRt rt;
void fiber_yield() { swapcontext(&rt.ctx[rt.cur], &rt.sched); }
void fiber_main() {
    rt.fn(rt.arg);
    rt.done[rt.cur] = 1;
    swapcontext(&rt.ctx[rt.cur], &rt.sched);
}
}  // namespace
void emu_barrier() {
    const unsigned g = rt.bar_gen;
    if (++rt.bar_count >= rt.alive) { rt.bar_count = 0; ++rt.bar_gen; return; }
    while (rt.bar_gen == g) fiber_yield();
}
float emu_shfl(float v, int src_lane) {
    if (!emu_fiber_mode) return v;
    const unsigned me = rt.cur, w = me >> 5, base = w << 5;
    const unsigned wsize = rt.n - base < 32u ? rt.n - base : 32u;
    const unsigned g = rt.gen[w];
    rt.buf[g & 1][me] = v;
    if (++rt.cnt[w] == wsize) { rt.cnt[w] = 0; ++rt.gen[w]; }
    else while (rt.gen[w] == g) fiber_yield();
    return rt.buf[g & 1][base + (unsigned)src_lane];
}
void emu_run_block(unsigned nthreads, void (*fn)(void*), void* arg) {
    rt.n = nthreads; rt.alive = nthreads; rt.fn = fn; rt.arg = arg;
    rt.bar_count = 0; rt.bar_gen = 0;
    if (rt.ctx.size() < nthreads) {
        const size_t old = rt.ctx.size();
        rt.ctx.resize(nthreads);
        rt.stacks.resize(nthreads, nullptr);
        for (size_t i = old; i < nthreads; ++i) rt.stacks[i] = (char*)malloc(kStack);
    }
    rt.done.assign(nthreads, 0);
    rt.buf[0].assign(nthreads, 0.f); rt.buf[1].assign(nthreads, 0.f);
    rt.cnt.assign((nthreads + 31) / 32, 0); rt.gen.assign((nthreads + 31) / 32, 0);
    for (unsigned i = 0; i < nthreads; ++i) {
        getcontext(&rt.ctx[i]);
        rt.ctx[i].uc_stack.ss_sp = rt.stacks[i];
        rt.ctx[i].uc_stack.ss_size = kStack;
        rt.ctx[i].uc_link = &rt.sched;
        makecontext(&rt.ctx[i], fiber_main, 0);
    }
    emu_fiber_mode = true;
    unsigned remaining = nthreads;
    while (remaining) {
        for (unsigned i = 0; i < nthreads; ++i) {
            if (rt.done[i]) continue;
            rt.cur = i;
            threadIdx = emu_dim3(i, 0, 0);
            swapcontext(&rt.sched, &rt.ctx[i]);
            if (rt.done[i]) {
                --remaining; --rt.alive;
                if (rt.bar_count && rt.bar_count >= rt.alive) { rt.bar_count = 0; ++rt.bar_gen; }
            }
        }
    }
    emu_fiber_mode = false;
}
extern "C" {
int glrgtv_abi_version(void) { return GLRGTV_ABI_VERSION; }
const char* glrgtv_last_cuda_error(void) { return "emulation build"; }
int glrgtv_check_device(void) { return GLRGTV_OK; }
unsigned long long glrgtv_launch_count(void) { return 0; }
int glrgtv_profile_enable(int) { return 0; }
int glrgtv_profile_read(float*, int*, int) { return 0; }
}
#else
#include <string.h>
static thread_local char g_last_error[256] = "";

int glr_record_launch_error(void) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return GLRGTV_OK;
    strncpy(g_last_error, cudaGetErrorString(e), sizeof(g_last_error) - 1);
    return GLRGTV_ERR_CUDA;
}

unsigned long long g_glr_launches = 0;
int g_glr_prof_on = 0;
#define GLR_PROF_SLOTS 32
#define GLR_PROF_POOL 8192
static cudaEvent_t g_ev[GLR_PROF_POOL][2];
static int g_ev_slot[GLR_PROF_POOL];
static int g_ev_made = 0, g_ev_used = 0, g_ev_open[GLR_PROF_SLOTS];

void glr_prof_mark(int slot, int end, void* stream) {
    if (slot < 0 || slot >= GLR_PROF_SLOTS) return;
    if (!end) {
        if (g_ev_used >= GLR_PROF_POOL) { g_ev_open[slot] = -1; return; }
        if (g_ev_used >= g_ev_made) {
            if (cudaEventCreate(&g_ev[g_ev_made][0]) != cudaSuccess || cudaEventCreate(&g_ev[g_ev_made][1]) != cudaSuccess) {
                g_ev_open[slot] = -1;
                return;
            }
            ++g_ev_made;
        }
        g_ev_slot[g_ev_used] = slot;
        g_ev_open[slot] = g_ev_used;
        cudaEventRecord(g_ev[g_ev_used][0], (cudaStream_t)stream);
        ++g_ev_used;
    } else if (g_ev_open[slot] >= 0) {
        cudaEventRecord(g_ev[g_ev_open[slot]][1], (cudaStream_t)stream);
        g_ev_open[slot] = -1;
    }
}

extern "C" {
unsigned long long glrgtv_launch_count(void) { return g_glr_launches; }
// start (on=1: also clears) / stop recording one CUDA-event pair around every kernel of the block entry points
int glrgtv_profile_enable(int on) {
    if (on) {
        g_ev_used = 0;
        for (int i = 0; i < GLR_PROF_SLOTS; ++i) g_ev_open[i] = -1;
    }
    g_glr_prof_on = on;
    return GLRGTV_OK;
}
// synchronises on the recorded events; ms[slot] = summed duration, count[slot] = launches.  returns #pairs
int glrgtv_profile_read(float* ms, int* count, int n_slots) {
    for (int i = 0; i < n_slots; ++i) { ms[i] = 0.f; count[i] = 0; }
    int n = 0;
    for (int i = 0; i < g_ev_used; ++i) {
        float t = 0.f;
        if (cudaEventSynchronize(g_ev[i][1]) != cudaSuccess) continue;
        if (cudaEventElapsedTime(&t, g_ev[i][0], g_ev[i][1]) != cudaSuccess) continue;
        if (g_ev_slot[i] < n_slots) { ms[g_ev_slot[i]] += t; count[g_ev_slot[i]] += 1; ++n; }
    }
    return n;
}
int glrgtv_abi_version(void) { return GLRGTV_ABI_VERSION; }
const char* glrgtv_last_cuda_error(void) { return g_last_error; }
int glrgtv_check_device(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return GLRGTV_ERR_DEVICE;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return GLRGTV_ERR_DEVICE;
    return major == 10 ? GLRGTV_OK : GLRGTV_ERR_DEVICE;
}
}
#endif
