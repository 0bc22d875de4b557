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
#include <cstdint>
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
    template <typename T>
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
                    const int64_t j = ok ? (int64_t)v : i;
                    const T t = x[i];
                    x[i] = x[j];
                    x[j] = t;
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

// Returns 0, or 1 when a user's lists cannot be written with non-negative slice bounds (a fold's share of positives
// longer than n_items / k_folds: the reference then slices with negative numbers; the caller falls back to its Python
// loop and the random state is left untouched), or 2 when `capacity` is too small.
static int kfold_lists(uint32_t* key, int* pos_io, int64_t n_users, int64_t n_items, int k_folds,
                       const int64_t* indptr, const int32_t* indices, int64_t capacity, int64_t* slot_ptr,
                       int64_t* slot_positives, int64_t* ids64, int32_t* ids32) {
    std::vector<uint32_t> saved(key, key + 624);
    const int saved_pos = *pos_io;
    Mt19937 rng{key, saved_pos};
    std::vector<int32_t> rated, non_rated;
    std::vector<uint8_t> mask((size_t)n_items);
    std::vector<int64_t> fills((size_t)k_folds);
    non_rated.reserve((size_t)n_items);
    int64_t at = 0;
    slot_ptr[0] = 0;
    auto bail = [&](int code) {
        for (int i = 0; i < 624; ++i) key[i] = saved[i];
        *pos_io = saved_pos;
        return code;
    };
    for (int64_t user = 0; user < n_users; ++user) {
        const int64_t e0 = indptr[user], e1 = indptr[user + 1];
        rated.assign(indices + e0, indices + e1);
        std::fill(mask.begin(), mask.end(), (uint8_t)1);
        for (int64_t e = e0; e < e1; ++e) mask[(size_t)indices[e]] = 0;
        non_rated.clear();
        for (int64_t i = 0; i < n_items; ++i)
            if (mask[(size_t)i]) non_rated.push_back((int32_t)i);
        const int64_t n_rated = (int64_t)rated.size(), n_non = (int64_t)non_rated.size();
        rng.shuffle(rated.data(), n_rated);                                        // evaluator.py:162
        const int64_t share = (int64_t)std::nearbyint((1.0 / (double)k_folds) * (double)n_rated);   // :165, half to even
        rng.shuffle(non_rated.data(), n_non);                                      // :171
        int64_t start = 0;
        for (int f = 0; f < k_folds; ++f) {
            int64_t p_lo, p_hi;
            if (f == k_folds - 1) clip_slice(start, n_rated, n_rated, &p_lo, &p_hi);
            else clip_slice(start, start + share, n_rated, &p_lo, &p_hi);
            start += share;
            const int64_t n_pos = p_hi - p_lo;
            const double fill = std::trunc((double)n_items / (double)k_folds - (double)n_pos);   // :184
            if (fill < 0.0) return bail(1);
            fills[(size_t)f] = (int64_t)fill;
            int64_t x_lo, x_hi;
            if (f > 0 && fills[(size_t)f] != fills[(size_t)f - 1])                  // :185-187
                clip_slice(f * fills[(size_t)f - 1], fills[(size_t)f - 1] * f + fills[(size_t)f], n_non, &x_lo, &x_hi);
            else                                                                   // :189
                clip_slice(f * fills[(size_t)f], fills[(size_t)f] * (f + 1), n_non, &x_lo, &x_hi);
            const int64_t len = n_pos + (x_hi - x_lo);
            if (at + len > capacity) return bail(2);
            for (int64_t i = p_lo; i < p_hi; ++i, ++at) {
                if (ids64) ids64[at] = rated[(size_t)i];
                if (ids32) ids32[at] = rated[(size_t)i];
            }
            for (int64_t i = x_lo; i < x_hi; ++i, ++at) {
                if (ids64) ids64[at] = non_rated[(size_t)i];
                if (ids32) ids32[at] = non_rated[(size_t)i];
            }
            slot_positives[user * k_folds + f] = n_pos;
            slot_ptr[user * k_folds + f + 1] = at;
        }
    }
    *pos_io = rng.pos;
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
