// Host-side random number plumbing for the ensemble input producers.
//
// The reference draws its randomness from third-party crates that are NOT under
// /root/reference: `rand ^0.8` (Cargo.toml:34) and `rand_pcg 0.3.1` (Cargo.toml:46).
// What is restated here is their published algorithm, as used at the reference's call
// sites (er_c.rs:133,158,350-351; er_m.rs:203-204,290-291; sw.rs:289,302,328;
// barabasi_albert.rs:203,215,218).  It is pinned by the reference's own seeded fixtures
// TestData/unchaning_*.json (adjacency order AND final generator state), see
// tests/test_generators.py.
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

namespace nec_host {

using u128 = unsigned __int128;

inline uint64_t rotr64(uint64_t x, unsigned r) { r &= 63; return r ? (x >> r) | (x << (64 - r)) : x; }
inline uint32_t rotr32(uint32_t x, unsigned r) { r &= 31; return r ? (x >> r) | (x << (32 - r)) : x; }

// rand_pcg::Pcg64 = Lcg128Xsl64: 128-bit LCG state, XSL-RR output.
class Pcg64 {
public:
    u128 state = 0, increment = 1;

    static constexpr u128 multiplier() {
        return ((u128)0x2360ED051FC65DA4ULL << 64) | (u128)0x4385DF649FCCF645ULL;
    }

    // rand_core::SeedableRng::seed_from_u64 (a PCG32 stream fills the 32 seed bytes),
    // followed by Lcg128Xsl64::from_seed / from_state_incr.
    static Pcg64 seed_from_u64(uint64_t s) {
        uint32_t words[8];
        for (int i = 0; i < 8; ++i) {
            s = s * 6364136223846793005ULL + 11634580027462260723ULL;
            uint32_t xorshifted = (uint32_t)(((s >> 18) ^ s) >> 27);
            words[i] = rotr32(xorshifted, (unsigned)(s >> 59));
        }
        auto u64_at = [&](int k) { return (uint64_t)words[2 * k] | ((uint64_t)words[2 * k + 1] << 32); };
        Pcg64 g;
        u128 st = (u128)u64_at(0) | ((u128)u64_at(1) << 64);
        g.increment = ((u128)u64_at(2) | ((u128)u64_at(3) << 64)) | 1;
        g.state = st + g.increment;
        g.step();
        return g;
    }

    static Pcg64 from_state(u128 state, u128 increment) {
        Pcg64 g;
        g.state = state;
        g.increment = increment;
        return g;
    }

    void step() { state = state * multiplier() + increment; }

    uint64_t next_u64() {
        step();
        unsigned rot = (unsigned)(state >> 122);
        uint64_t folded = (uint64_t)(state >> 64) ^ (uint64_t)state;
        return rotr64(folded, rot);
    }
    uint32_t next_u32() { return (uint32_t)next_u64(); }

    // rng.gen::<f64>() — 53 random bits scaled into [0,1).
    double gen_f64() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }

    // rng.gen_range(0..n) for usize (64-bit): widening multiply with the
    // "conservative zone" rejection of rand 0.8's sample_single.
    uint64_t gen_below_u64(uint64_t n) {
        uint64_t zone = (n << __builtin_clzll(n)) - 1;
        for (;;) {
            u128 wide = (u128)next_u64() * (u128)n;
            if ((uint64_t)wide <= zone) return (uint64_t)(wide >> 64);
        }
    }
    // rng.gen_range(0..n) for u32.
    uint32_t gen_below_u32(uint32_t n) {
        uint32_t zone = (n << __builtin_clz(n)) - 1;
        for (;;) {
            uint64_t wide = (uint64_t)next_u32() * (uint64_t)n;
            if ((uint32_t)wide <= zone) return (uint32_t)(wide >> 32);
        }
    }
};

// rand 0.8 SliceRandom::shuffle: Fisher-Yates from the back; the index is drawn with
// the u32 sampler whenever the bound fits in u32.
template <class T>
void shuffle(std::vector<T> &v, Pcg64 &rng) {
    for (size_t i = v.size(); i-- > 1;) {
        size_t bound = i + 1;
        size_t j = bound <= 0xFFFFFFFFULL ? (size_t)rng.gen_below_u32((uint32_t)bound)
                                          : (size_t)rng.gen_below_u64(bound);
        std::swap(v[i], v[j]);
    }
}

// rand 0.8 Uniform<usize>::new(0, total).sample(): unlike gen_range this uses the exact
// rejection zone (ints_to_reject = (2^64 - range) % range).
struct UniformBelow {
    uint64_t range, zone;
    explicit UniformBelow(uint64_t total) : range(total) {
        uint64_t reject = (UINT64_MAX - range + 1) % range;
        zone = UINT64_MAX - reject;
    }
    uint64_t sample(Pcg64 &rng) const {
        for (;;) {
            u128 wide = (u128)rng.next_u64() * (u128)range;
            if ((uint64_t)wide <= zone) return (uint64_t)(wide >> 64);
        }
    }
};

// rand 0.8 WeightedIndex<usize>: cumulative weights (last one dropped), a uniform draw in
// [0,total) and the first index whose cumulative weight exceeds the draw.
// No reference fixture pins this sampler (barabasi_albert.rs has only a smoke test):
// BA construction is "parity unpinned" at the generator level.
struct WeightedIndex {
    std::vector<uint64_t> cumulative;
    UniformBelow uniform;
    static uint64_t total_of(const std::vector<uint64_t> &w) {
        uint64_t t = 0;
        for (uint64_t x : w) t += x;
        return t;
    }
    explicit WeightedIndex(const std::vector<uint64_t> &w) : uniform(total_of(w)) {
        cumulative.reserve(w.size() ? w.size() - 1 : 0);
        uint64_t acc = 0;
        for (size_t i = 0; i + 1 < w.size(); ++i) {
            acc += w[i];
            cumulative.push_back(acc);
        }
    }
    size_t sample(Pcg64 &rng) const {
        uint64_t chosen = uniform.sample(rng);
        size_t lo = 0, hi = cumulative.size();
        while (lo < hi) {  // partition point of (w <= chosen)
            size_t mid = (lo + hi) / 2;
            if (cumulative[mid] <= chosen) lo = mid + 1; else hi = mid;
        }
        return lo;
    }
};

}  // namespace nec_host
