// Multi-GPU exchange over NVLink peer memory (one process per GPU).
//
// The reference has no distributed code; what shards is the independence of rows inside als_step
// (/root/reference/lib/collaborative_filtering.py:82-86 reads only `fixed_vecs`).  Every rank keeps
// a full replica of U, V and of the whitened factors; the replicas and a small mailbox live in ONE
// cudaMalloc'd arena per rank that the other ranks map with cudaIpcOpenMemHandle.  A rank that has
// solved its rows stores them straight into every replica (from the GEMM epilogues of score.cuh /
// tc_score.cuh, or with push_rows_kernel for the few rows solved elsewhere), so the "all-gather of
// the updated factor matrix after each half-iteration" is the epilogue of the kernel that produced
// the rows.  The k x k Gram partials and the RMSE scalars travel through the mailbox: every rank
// writes its slot into every peer and each rank sums the slots in rank order (bit-identical on all
// ranks, no floating-point atomics).
//
// Ordering: every rank runs the same sequence of signal / wait pairs.  signal(seq) is a one-CTA
// kernel behind the stores it publishes (stream order + __threadfence_system) that writes `seq`
// into flags[my_rank] of every peer; wait(seq) spins until all of this rank's flags reached `seq`.
// The spin is bounded (globaltimer): a dead peer raises status 2 instead of hanging the GPU.
#pragma once
#include <stdint.h>
#ifdef HYPREC_EMU
#include <sched.h>
#include <time.h>
#endif

#include "common.cuh"

namespace hyprec {

constexpr int kMaxRanks = 8;

struct PeerList {
    void* base[kMaxRanks];   // arena of every rank as mapped into this process (base[rank] = own)
    int world;
    int rank;
};

#ifdef HYPREC_EMU
// CPU emulator (tests/emu): ranks are host threads, arenas host memory
inline void st_release_sys_u64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline unsigned long long ld_acquire_sys_u64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline unsigned long long global_timer_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __nanosleep(unsigned) { sched_yield(); }
struct int4 { int x, y, z, w; };
#else
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

// flags of rank r: world x u64 at base[r] + off_flags; entry [src] is written by rank src only.
__global__ void comm_signal_kernel(PeerList peers, size_t off_flags, unsigned long long seq) {
    const int dst = threadIdx.x;
    if (dst >= peers.world) return;
    __threadfence_system();
    unsigned long long* flags = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(peers.base[dst]) + off_flags);
    st_release_sys_u64(flags + peers.rank, seq);
}

__global__ void comm_wait_kernel(PeerList peers, size_t off_flags, unsigned long long seq, unsigned long long timeout_ns,
                                 int* status) {
    const int src = threadIdx.x;
    if (src >= peers.world) return;
    const unsigned long long* flags =
        reinterpret_cast<const unsigned long long*>(static_cast<const unsigned char*>(peers.base[peers.rank]) + off_flags);
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys_u64(flags + src) < seq) {
        // (a barrier that already timed out makes the later ones fall through: the run is lost, end it quickly)
        if (*reinterpret_cast<volatile int*>(status) != 0 || global_timer_ns() - t0 > timeout_ns) {
            *status = 2;       // a peer never arrived
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

// Rows of a row-major [.][k] matrix that sits at the same arena offset on every rank: copy the rows
// (a contiguous range, or a list) from the local replica into every other replica.  16-byte stores,
// one warp per row and destination.
template <typename T>
__global__ void push_rows_kernel(PeerList peers, size_t off_matrix, int k, int64_t row_lo, int64_t n_rows,
                                 const int32_t* __restrict__ row_list) {
    const int lane = threadIdx.x % 32;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x / 32);
    const T* local = reinterpret_cast<const T*>(static_cast<const unsigned char*>(peers.base[peers.rank]) + off_matrix);
    const int vec = 16 / (int)sizeof(T);
    const bool aligned = (k % vec) == 0;
    for (int64_t w = warp; w < n_rows; w += n_warps) {
        const int64_t row = row_list ? (int64_t)row_list[w] : row_lo + w;
        const T* src = local + row * k;
        for (int d = 0; d < peers.world; ++d) {
            if (d == peers.rank) continue;
            T* dst = reinterpret_cast<T*>(static_cast<unsigned char*>(peers.base[d]) + off_matrix) + row * k;
            if (aligned) {
                const int4* s4 = reinterpret_cast<const int4*>(src);
                int4* d4 = reinterpret_cast<int4*>(dst);
                for (int c = lane; c < k / vec; c += 32) d4[c] = s4[c];
            } else {
                for (int c = lane; c < k; c += 32) dst[c] = src[c];
            }
        }
    }
    __threadfence_system();
}

// Mailbox: `count` doubles from a local buffer into slot [rank] of every rank (own included).
__global__ void push_slot_kernel(PeerList peers, size_t off_slots, size_t slot_doubles, const double* __restrict__ src, int count) {
    for (int d = 0; d < peers.world; ++d) {
        double* dst = reinterpret_cast<double*>(static_cast<unsigned char*>(peers.base[d]) + off_slots) + (size_t)peers.rank * slot_doubles;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
}

// out[i] = sum over ranks (in rank order) of slot[r][i], i < count; optionally a cast copy.
template <typename T>
__global__ void reduce_slots_kernel(const double* __restrict__ slots, size_t slot_doubles, int world, int count,
                                    double* __restrict__ out, T* __restrict__ out_cast) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += slots[(size_t)r * slot_doubles + i];
    out[i] = s;
    if (out_cast) out_cast[i] = (T)s;
}

__global__ void set_double_kernel(double* p, double v) { *p = v; }

// Reproducible initial factors without a host copy (the scaled workload: 3 GB of float32 factors):
// element e of the matrix = the float32 in [0, 1) made of the top 24 bits of mix64(seed, e).
// Restated in numpy by oracle.als.uniform_factors.
HYPREC_HD unsigned long long mix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
template <typename T>
__global__ void uniform_fill_kernel(T* __restrict__ out, int64_t count, unsigned long long seed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const unsigned long long r = mix64(seed * 0xD1342543DE82EF95ull + (unsigned long long)i);
    out[i] = (T)((float)(r >> 40) * (1.0f / 16777216.0f));
}

// out[w][0..k) = (double) matrix[rows[w]][0..k): a few rows of a device-resident factor matrix for the host
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ matrix, int k, const int64_t* __restrict__ rows, int64_t n_rows,
                                   double* __restrict__ out) {
    const int64_t w = blockIdx.x;
    if (w >= n_rows) return;
    const T* src = matrix + rows[w] * k;
    for (int c = threadIdx.x; c < k; c += blockDim.x) out[w * k + c] = (double)src[c];
}

}  // namespace hyprec
