// rtb_warp.cuh — the warp-level operations the traversal scheduler (rtb_trace.cuh) is written
// against.  On the device they are the sm_100a intrinsics.  In a plain g++ translation unit
// (tests/hostcheck) they are an EMULATION: the 32 lanes of a warp run as 32 cooperative fibers
// on one thread and every collective (ballot, shuffle, match, sync) is a rendezvous of all
// lanes, so the scheduler's control flow — votes, window sort, refill, queue pushes — is
// rehearsed by the CPU suite in the GPU-less build container.  Test infrastructure on the host
// side; the product only ever runs the device branch.
#ifndef RTB_WARP_CUH
#define RTB_WARP_CUH

#include "rtb_math.cuh"

#if defined(__CUDACC__)
#define RTB_WD __device__ __forceinline__
#else
#define RTB_WD inline
#include <cassert>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <ucontext.h>
#include <vector>
#endif

namespace rtb {

#if defined(__CUDACC__)
typedef float4 Vec4f;
typedef uint4 Vec4u;
typedef float2 Vec2f;
typedef uint2 Vec2u;
#else
struct alignas(16) Vec4f {
    float x, y, z, w;
};
struct alignas(16) Vec4u {
    uint32_t x, y, z, w;
};
struct alignas(8) Vec2f {
    float x, y;
};
struct alignas(8) Vec2u {
    uint32_t x, y;
};
#endif

RTB_WD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
RTB_WD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

#if defined(__CUDACC__)

struct WarpOps {
    static __device__ __forceinline__ uint32_t lane() { return threadIdx.x & 31u; }
    static __device__ __forceinline__ uint32_t ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
    template <class T> static __device__ __forceinline__ T shfl(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
    static __device__ __forceinline__ uint32_t match_any(uint32_t k) { return __match_any_sync(0xffffffffu, k); }
    static __device__ __forceinline__ void sync() { __syncwarp(); }
    static __device__ __forceinline__ uint32_t popc(uint32_t m) { return uint32_t(__popc(m)); }
    static __device__ __forceinline__ int ffs(uint32_t m) { return __ffs(m); }
    static __device__ __forceinline__ uint32_t atomic_add(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
    static __device__ __forceinline__ unsigned long long atomic_add(unsigned long long *p, unsigned long long v) {
        return atomicAdd(p, v);
    }
    static __device__ __forceinline__ void atomic_add(float *p, float v) { atomicAdd(p, v); }
};

#else // ---- host emulation ------------------------------------------------------------------

// One warp = 32 fibers.  run() resumes the lanes round-robin; a lane that reaches a collective
// publishes its operand and yields until all 32 have arrived.  Operands are double-buffered by
// generation parity: a fast lane may already be publishing for the next collective while a slow
// one still reads the previous one.
class HostWarp {
  public:
    static HostWarp *&current() {
        static thread_local HostWarp *w = nullptr;
        return w;
    }
    explicit HostWarp(size_t stack_bytes = 1u << 20) : stack_bytes_(stack_bytes) {}
    void run(const std::function<void()> &body) {
        body_ = &body;
        HostWarp *prev = current();
        current() = this;
        std::vector<std::vector<char>> stacks(32, std::vector<char>(stack_bytes_));
        for (int l = 0; l < 32; ++l) {
            getcontext(&ctx_[l]);
            ctx_[l].uc_stack.ss_sp = stacks[l].data();
            ctx_[l].uc_stack.ss_size = stack_bytes_;
            ctx_[l].uc_link = &main_;
            makecontext(&ctx_[l], reinterpret_cast<void (*)()>(&HostWarp::entry), 0);
            finished_[l] = false;
            waiting_[l] = false;
        }
        arrived_ = 0;
        gen_ = 0;
        int live = 32;
        while (live > 0) {
            bool progressed = false;
            for (int l = 0; l < 32; ++l) {
                if (finished_[l] || waiting_[l])
                    continue;
                lane_ = l;
                swapcontext(&main_, &ctx_[l]);
                progressed = true;
                if (finished_[l])
                    --live;
            }
            if (arrived_ == 32) { // release the collective
                arrived_ = 0;
                ++gen_;
                for (int l = 0; l < 32; ++l)
                    waiting_[l] = false;
                progressed = true;
            }
            if (!progressed) {
                current() = prev;
                throw std::runtime_error("HostWarp: deadlock (a lane left a collective's warp)");
            }
        }
        current() = prev;
    }
    uint32_t lane() const { return uint32_t(lane_); }
    // publish `v`, wait for all lanes, return a pointer to the 32 published values
    const uint64_t *exchange(uint64_t v) {
        const int b = int(gen_ & 1u);
        slot_[b][lane_] = v;
        waiting_[lane_] = true;
        ++arrived_;
        const int me = lane_;
        swapcontext(&ctx_[me], &main_);
        lane_ = me;
        return slot_[b];
    }

  private:
    static void entry() {
        HostWarp *w = current();
        const int me = w->lane_;
        (*w->body_)();
        w->finished_[me] = true;
    }
    size_t stack_bytes_;
    const std::function<void()> *body_ = nullptr;
    ucontext_t main_, ctx_[32];
    bool finished_[32], waiting_[32];
    int lane_ = 0, arrived_ = 0;
    uint32_t gen_ = 0;
    uint64_t slot_[2][32];
};

struct WarpOps {
    static uint32_t lane() { return HostWarp::current()->lane(); }
    static uint32_t ballot(bool p) {
        const uint64_t *s = HostWarp::current()->exchange(p ? 1u : 0u);
        uint32_t m = 0;
        for (int i = 0; i < 32; ++i)
            m |= uint32_t(s[i] & 1u) << i;
        return m;
    }
    template <class T> static T shfl(T v, int src) {
        static_assert(sizeof(T) <= 8, "shfl: up to 64-bit values");
        uint64_t bits = 0;
        memcpy(&bits, &v, sizeof(T));
        const uint64_t *s = HostWarp::current()->exchange(bits);
        T out;
        memcpy(&out, &s[src & 31], sizeof(T));
        return out;
    }
    static uint32_t match_any(uint32_t k) {
        const uint64_t *s = HostWarp::current()->exchange(k);
        uint32_t m = 0;
        for (int i = 0; i < 32; ++i)
            m |= uint32_t(uint32_t(s[i]) == k) << i;
        return m;
    }
    static void sync() { HostWarp::current()->exchange(0); }
    static uint32_t popc(uint32_t m) { return uint32_t(__builtin_popcount(m)); }
    static int ffs(uint32_t m) { return __builtin_ffs(int(m)); }
    // one host thread runs every lane: plain read-modify-write
    static uint32_t atomic_add(uint32_t *p, uint32_t v) {
        const uint32_t o = *p;
        *p = o + v;
        return o;
    }
    static unsigned long long atomic_add(unsigned long long *p, unsigned long long v) {
        const unsigned long long o = *p;
        *p = o + v;
        return o;
    }
    static void atomic_add(float *p, float v) { *p += v; }
};

#endif

} // namespace rtb

#endif // RTB_WARP_CUH
