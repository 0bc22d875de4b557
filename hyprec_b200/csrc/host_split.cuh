// Host-side k-fold candidate lists: what Evaluator.get_kfold_indices does per user
// (/root/reference/lib/evaluator.py:136-195) -- shuffle the user's rated items, shuffle its non-rated items, deal both
// into k_folds lists -- consuming numpy's legacy global random stream draw for draw, so the lists are bit-identical to
// the reference's.  No CUDA here: plain C++ that runs before anything reaches the GPU.
//
// The stream is numpy.random's MT19937 (state = 624 key words + a position, numpy.random.get_state()); a legacy
// `shuffle` of a 1-d array is a Fisher-Yates walk from the last element down whose draws are
// `random_interval(max = i)`: 32-bit outputs masked to the smallest 2^b - 1 >= i, rejected while > i
// (numpy/random/mtrand.pyx `_shuffle_raw`, numpy/random/src/distributions/distributions.c `random_interval`;
// numpy itself is a third-party dependency of the reference, pinned by tests/test_host_mirror.py against the
// installed numpy).
#pragma once
#include <cmath>
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <thread>
#include <vector>

namespace hyprec_host {

struct Mt19937 {
    uint32_t* key;   // 624 words, caller-owned (in / out)
    int pos;

    void refill() {
        constexpr int N = 624, M = 397;
        constexpr uint32_t A = 0x9908b0dfu, UP = 0x80000000u, LO = 0x7fffffffu;
        int kk = 0;
        uint32_t y;
        for (; kk < N - M; ++kk) {
            y = (key[kk] & UP) | (key[kk + 1] & LO);
            key[kk] = key[kk + M] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        }
        for (; kk < N - 1; ++kk) {
            y = (key[kk] & UP) | (key[kk + 1] & LO);
            key[kk] = key[kk + (M - N)] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        }
        y = (key[N - 1] & UP) | (key[0] & LO);
        key[N - 1] = key[M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        pos = 0;
    }
    inline uint32_t next32() {
        if (pos >= 624) refill();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    // random_interval(max), max < 2^32
    inline uint32_t interval(uint32_t max) {
        if (max == 0) return 0;
        uint32_t mask = max;
        mask |= mask >> 1;
        mask |= mask >> 2;
        mask |= mask >> 4;
        mask |= mask >> 8;
        mask |= mask >> 16;
        uint32_t v;
        while ((v = (next32() & mask)) > max) {
        }
        return v;
    }
    // Fisher-Yates from the last element down, one random_interval(i) per step.  Written without a data-dependent
    // branch: a rejected draw (about one in four) swaps x[i] with itself and leaves i where it is, so the only
    // serial chain is i -> mask -> compare -> i (the rejection loop as a branch mispredicts ~0.4 times per step
    // and costs 11 ns per step, this form ~3 ns; same draws, same result).
    // SWAP = false: only walk the stream (which draws a shuffle of this length consumes does not depend on the data)
    template <typename T, bool SWAP = true>
    void shuffle(T* x, int64_t n) {
        int64_t i = n - 1;
        while (i >= 1) {
            // all i of one binade share the mask: the inner loop's serial chain is compare -> decrement only
            const uint32_t mask = 0xffffffffu >> __builtin_clz((uint32_t)i);
            const int64_t floor_i = (int64_t)(mask >> 1) + 1;      // smallest i with this mask
            while (i >= floor_i) {
                if (pos >= 624) refill();
                const int64_t budget = 624 - pos;                   // draws left before the next refill
                const uint32_t* src = key + pos;
                int64_t c = 0;
                while (c < budget && i >= floor_i) {
                    uint32_t y = src[c++];
                    y ^= (y >> 11);
                    y ^= (y << 7) & 0x9d2c5680u;
                    y ^= (y << 15) & 0xefc60000u;
                    y ^= (y >> 18);
                    const uint32_t v = y & mask;
                    const bool ok = v <= (uint32_t)i;
                    if (SWAP) {
                        const int64_t j = ok ? (int64_t)v : i;
                        const T t = x[i];
                        x[i] = x[j];
                        x[j] = t;
                    }
                    i -= ok ? 1 : 0;
                }
                pos += (int)c;
            }
        }
    }
};

// Python slice x[a:b] of a sequence of length n, for a, b >= 0: clipped bounds
static inline void clip_slice(int64_t a, int64_t b, int64_t n, int64_t* lo, int64_t* hi) {
    if (a > n) a = n;
    if (b > n) b = n;
    if (b < a) b = a;
    *lo = a;
    *hi = b;
}

// Lengths of one user's k_folds lists -- they depend on the counts only, not on the shuffles: positives per fold,
// slice of the shuffled non-rated items per fold (evaluator.py:176-192).  False when a list would need the reference's
// negative slice bounds (a fold's share of positives longer than n_items / k_folds).
struct FoldSlices {
    int64_t p_lo, p_hi, x_lo, x_hi;
};
static inline bool fold_slices(int64_t n_rated, int64_t n_non, int64_t n_items, int k_folds, FoldSlices* out) {
    const int64_t share = (int64_t)std::nearbyint((1.0 / (double)k_folds) * (double)n_rated);   // :165, half to even
    int64_t start = 0, prev_fill = 0;
    for (int f = 0; f < k_folds; ++f) {
        FoldSlices& s = out[f];
        if (f == k_folds - 1) clip_slice(start, n_rated, n_rated, &s.p_lo, &s.p_hi);
        else clip_slice(start, start + share, n_rated, &s.p_lo, &s.p_hi);
        start += share;
        const double fill_d = std::trunc((double)n_items / (double)k_folds - (double)(s.p_hi - s.p_lo));   // :184
        if (fill_d < 0.0) return false;
        const int64_t fill = (int64_t)fill_d;
        if (f > 0 && fill != prev_fill) clip_slice(f * prev_fill, prev_fill * f + fill, n_non, &s.x_lo, &s.x_hi);   // :185-187
        else clip_slice(f * fill, fill * (f + 1), n_non, &s.x_lo, &s.x_hi);                                        // :189
        prev_fill = fill;
    }
    return true;
}

// Returns 0, or 1 when a user's lists cannot be written with non-negative slice bounds (the caller falls back to its
// Python loop; the random state is left untouched), or 2 when `capacity` is too small.
//
// The random stream is sequential, but how many draws a shuffle consumes depends only on its length and on the draws
// themselves, never on the data being shuffled.  So: (1) the list lengths follow from the counts; (2) ONE sequential
// pass walks the stream through every user's two shuffles without touching any data (masked rejection only, ~2 ns per
// step) and records the generator state at the first user of every chunk; (3) the chunks are shuffled and dealt on
// `n_threads` threads, each from its recorded state.  Same draws, same lists, the memory-bound part in parallel.
static int kfold_lists(uint32_t* key, int* pos_io, int64_t n_users, int64_t n_items, int k_folds,
                       const int64_t* indptr, const int32_t* indices, int64_t capacity, int64_t* slot_ptr,
                       int64_t* slot_positives, int64_t* ids64, int32_t* ids32, int n_threads) {
    // (1) lengths
    std::vector<FoldSlices> fs((size_t)k_folds);
    slot_ptr[0] = 0;
    for (int64_t user = 0; user < n_users; ++user) {
        const int64_t n_rated = indptr[user + 1] - indptr[user];
        if (!fold_slices(n_rated, n_items - n_rated, n_items, k_folds, fs.data())) return 1;
        for (int f = 0; f < k_folds; ++f) {
            const int64_t slot = user * k_folds + f;
            slot_positives[slot] = fs[(size_t)f].p_hi - fs[(size_t)f].p_lo;
            slot_ptr[slot + 1] = slot_ptr[slot] + slot_positives[slot] + (fs[(size_t)f].x_hi - fs[(size_t)f].x_lo);
        }
    }
    if (slot_ptr[n_users * k_folds] > capacity) return 2;
    // (2) the stream, sequentially: generator state at the first user of every chunk
    if (n_threads < 1) n_threads = 1;
    const int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>(n_users, (int64_t)n_threads * 8));
    const int64_t per_chunk = (n_users + n_chunks - 1) / std::max<int64_t>(n_chunks, 1);
    std::vector<uint32_t> states((size_t)n_chunks * 624);
    std::vector<int> positions((size_t)n_chunks);
    {
        Mt19937 rng{key, *pos_io};
        for (int64_t user = 0; user < n_users; ++user) {
            if (user % per_chunk == 0) {
                const int64_t c = user / per_chunk;
                for (int i = 0; i < 624; ++i) states[(size_t)c * 624 + i] = key[i];
                positions[(size_t)c] = rng.pos;
            }
            const int64_t n_rated = indptr[user + 1] - indptr[user];
            rng.template shuffle<int32_t, false>(nullptr, n_rated);              // evaluator.py:162
            rng.template shuffle<int32_t, false>(nullptr, n_items - n_rated);    // :171
        }
        *pos_io = rng.pos;      // key[] now holds the state behind the last shuffle
    }
    // (3) shuffle and deal, chunk by chunk
    auto work = [&](int64_t c) {
        const int64_t u0 = c * per_chunk, u1 = std::min<int64_t>(n_users, u0 + per_chunk);
        std::vector<uint32_t> local(states.begin() + c * 624, states.begin() + (c + 1) * 624);
        Mt19937 rng{local.data(), positions[(size_t)c]};
        std::vector<int32_t> rated, non_rated((size_t)n_items);
        std::vector<FoldSlices> mine((size_t)k_folds);
        for (int64_t user = u0; user < u1; ++user) {
            const int64_t e0 = indptr[user], e1 = indptr[user + 1];
            rated.assign(indices + e0, indices + e1);
            // the non-rated items in increasing order: the gaps between consecutive (sorted) rated ids
            int64_t n_non = 0;
            int32_t next = 0;
            for (int64_t e = e0; e <= e1; ++e) {
                const int32_t stop = e < e1 ? indices[e] : (int32_t)n_items;
                for (int32_t i = next; i < stop; ++i) non_rated[(size_t)n_non++] = i;
                next = stop + 1;
            }
            const int64_t n_rated = e1 - e0;
            rng.template shuffle<int32_t, true>(rated.data(), n_rated);
            rng.template shuffle<int32_t, true>(non_rated.data(), n_non);
            fold_slices(n_rated, n_non, n_items, k_folds, mine.data());
            for (int f = 0; f < k_folds; ++f) {
                int64_t at = slot_ptr[user * k_folds + f];
                const FoldSlices& s = mine[(size_t)f];
                for (int64_t i = s.p_lo; i < s.p_hi; ++i, ++at) {
                    if (ids64) ids64[at] = rated[(size_t)i];
                    if (ids32) ids32[at] = rated[(size_t)i];
                }
                for (int64_t i = s.x_lo; i < s.x_hi; ++i, ++at) {
                    if (ids64) ids64[at] = non_rated[(size_t)i];
                    if (ids32) ids32[at] = non_rated[(size_t)i];
                }
            }
        }
    };
    const int64_t used_chunks = (n_users + per_chunk - 1) / per_chunk;
    if (n_threads == 1 || used_chunks <= 1) {
        for (int64_t c = 0; c < used_chunks; ++c) work(c);
    } else {
        std::atomic<int64_t> cursor(0);
        std::vector<std::thread> pool;
        const int n_pool = (int)std::min<int64_t>(n_threads, used_chunks);
        for (int t = 0; t < n_pool; ++t)
            pool.emplace_back([&]() {
                for (int64_t c = cursor.fetch_add(1); c < used_chunks; c = cursor.fetch_add(1)) work(c);
            });
        for (auto& th : pool) th.join();
    }
    return 0;
}

// The draws of Evaluator.naive_split_users (evaluator.py:92-93): per user, in user order,
// numpy.random.choice(non_zeros, size) = `size` legacy randint(0, len) values, i.e. masked rejection draws below
// len - 1 (no draw at all when the user has a single rating).  picks[] receives the CSR entry numbers drawn
// (with repeats: the reference samples with replacement), user after user; sizes[u] entries each.
static void naive_picks(uint32_t* key, int* pos_io, int64_t n_users, const int64_t* indptr, const int64_t* sizes,
                        int64_t* picks) {
    Mt19937 rng{key, *pos_io};
    int64_t at = 0;
    for (int64_t user = 0; user < n_users; ++user) {
        const int64_t len = indptr[user + 1] - indptr[user];
        for (int64_t i = 0; i < sizes[user]; ++i)
            picks[at++] = indptr[user] + (len > 1 ? (int64_t)rng.interval((uint32_t)(len - 1)) : 0);
    }
    *pos_io = rng.pos;
}

}  // namespace hyprec_host
