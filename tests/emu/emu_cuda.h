// TEST INFRASTRUCTURE: a tiny "CUDA on pthreads" shim so the SIMT kernels under
// hyprec_b200/csrc/*.cuh can be compiled with g++ and executed on the CPU-only build box.
// One block runs at a time; every CUDA thread of the block is a real pthread, __syncthreads()
// is a pthread barrier and warp shuffles go through a per-warp exchange slot.  It exists to
// catch indexing / synchronisation mistakes before GPU time is spent; it is never part of the
// product path (the product is the nvcc-built libhyprec_b200.so).
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(x) __attribute__((aligned(x)))

struct emu_dim3 {
    unsigned x, y, z;
    emu_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef emu_dim3 dim3;

namespace emu {
struct Block {
    pthread_barrier_t bar;
    std::vector<pthread_barrier_t> warp_bar;
    std::vector<uint64_t> xchg;      // 32 slots per warp
    std::vector<unsigned char> smem;
    unsigned nthreads;
};
extern thread_local emu_dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
extern thread_local Block* t_block;

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    unsigned nthreads = block.x * block.y * block.z;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                Block blk;
                blk.nthreads = nthreads;
                pthread_barrier_init(&blk.bar, nullptr, nthreads);
                unsigned nwarps = (nthreads + 31) / 32;
                blk.warp_bar.resize(nwarps);
                for (unsigned w = 0; w < nwarps; ++w)
                    pthread_barrier_init(&blk.warp_bar[w], nullptr, std::min(32u, nthreads - 32 * w));
                blk.xchg.assign(32 * nwarps, 0);
                blk.smem.assign(smem_bytes + 64, 0xCD);   // poison: uninitialised reads show up
                std::vector<std::thread> ts;
                ts.reserve(nthreads);
                for (unsigned t = 0; t < nthreads; ++t)
                    ts.emplace_back([&, t]() {
                        t_block = &blk;
                        t_blockDim = block;
                        t_gridDim = grid;
                        t_blockIdx = emu_dim3(bx, by, bz);
                        t_threadIdx = emu_dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        body();
                    });
                for (auto& th : ts) th.join();
                pthread_barrier_destroy(&blk.bar);
                for (auto& wb : blk.warp_bar) pthread_barrier_destroy(&wb);
            }
}
inline unsigned linear_tid() {
    return t_threadIdx.x + t_blockDim.x * (t_threadIdx.y + t_blockDim.y * t_threadIdx.z);
}
inline unsigned char* dyn_smem() {
    uintptr_t p = (uintptr_t)t_block->smem.data();
    return (unsigned char*)((p + 15) & ~(uintptr_t)15);
}
template <typename T>
inline T exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    unsigned tid = linear_tid(), w = tid / 32, lane = tid % 32;
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    t_block->xchg[32 * w + lane] = raw;
    pthread_barrier_wait(&t_block->warp_bar[w]);
    unsigned width = std::min(32u, t_block->nthreads - 32 * w);
    uint64_t got = (src_lane >= 0 && (unsigned)src_lane < width) ? t_block->xchg[32 * w + src_lane] : raw;
    pthread_barrier_wait(&t_block->warp_bar[w]);
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
}  // namespace emu

#define threadIdx emu::t_threadIdx
#define blockIdx emu::t_blockIdx
#define blockDim emu::t_blockDim
#define gridDim emu::t_gridDim
#define HYPREC_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()

inline void __syncthreads() { pthread_barrier_wait(&emu::t_block->bar); }
inline void __syncwarp(unsigned = 0xffffffffu) {
    pthread_barrier_wait(&emu::t_block->warp_bar[emu::linear_tid() / 32]);
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu::exchange(v, src); }
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    return emu::exchange(v, (int)(emu::linear_tid() % 32) ^ m);
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    return emu::exchange(v, (int)(emu::linear_tid() % 32 + d));
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    int lane = (int)(emu::linear_tid() % 32);
    return emu::exchange(v, lane - (int)d >= 0 ? lane - (int)d : lane);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned out = 0;
    for (int l = 0; l < 32; ++l) out |= (emu::exchange(pred ? 1 : 0, l) ? 1u : 0u) << l;
    return out;
}

template <typename T>
inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
inline double atomicAdd(double* p, double v) {
    uint64_t* ip = (uint64_t*)p;
    uint64_t old = __atomic_load_n(ip, __ATOMIC_SEQ_CST), want;
    double cur;
    do {
        memcpy(&cur, &old, 8);
        double nv = cur + v;
        memcpy(&want, &nv, 8);
    } while (!__atomic_compare_exchange_n(ip, &old, want, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    return cur;
}
inline unsigned atomicCAS(unsigned* p, unsigned compare, unsigned val) {
    __atomic_compare_exchange_n(p, &compare, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return compare;
}
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline float __uint_as_float(unsigned v) { float r; memcpy(&r, &v, 4); return r; }
inline unsigned __float_as_uint(float v) { unsigned r; memcpy(&r, &v, 4); return r; }
struct uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
template <typename T>
inline T __ldg(const T* p) { return *p; }
inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
