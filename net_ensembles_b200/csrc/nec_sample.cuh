// On-device ErEnsembleC::randomize, bit-identical to the crate:
//   for i in 0..n { for j in i+1..n { if rng.gen::<f64>() <= prob { adj[i].push(j); adj[j].push(i) } } }   (src/er_c.rs:128-143)
// with rng = rand_pcg::Pcg64 (Lcg128Xsl64; rand_pcg 0.3.1, not under /root/reference, restated from its published
// algorithm exactly as csrc/host/nec_rand.hpp does and pinned by TestData/unchaning_erc_{1,2}.json).
// Every unordered pair consumes exactly ONE 64-bit output, in lexicographic order, so pair (i, j) uses output
// number P(i) + (j - i), P(i) = i(n-1) - i(i-1)/2, and an LCG can jump: state_k = A^k s + C (A^k - 1)/(A - 1).
// One warp takes a row i of the pair triangle: the row's start state comes from a log-time jump (one thread per
// row), lane l then sits l+1 outputs further (precomputed A^(l+1), C_(l+1)) and advances 32 outputs per step
// (A^32, C_32): one 128-bit multiply-add per draw, like the sequential generator.
// Adjacency ORDER: vertex v's list receives its smaller neighbours first (pairs (k, v), k ascending - they are
// drawn in rows k < v) and then its larger ones (pairs (v, j), j ascending): each list is simply ascending.  The
// "upper" part is written in order by the row's warp (ballot ranks); the "lower" part is collected through atomic
// cursors and then sorted, which makes the result deterministic and equal to the crate's lists.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nec {

typedef unsigned __int128 u128;

struct LcgStep { unsigned long long mul_lo, mul_hi, add_lo, add_hi; };     // state -> mul * state + add  (mod 2^128)

__host__ __device__ inline u128 u128_of(unsigned long long lo, unsigned long long hi) { return ((u128)hi << 64) | (u128)lo; }
__host__ __device__ inline u128 pcg64_multiplier() { return u128_of(0x4385DF649FCCF645ULL, 0x2360ED051FC65DA4ULL); }

// k steps of state -> A state + C as one affine map (binary method, O(log k) 128-bit multiplies)
__host__ __device__ inline void lcg_jump(unsigned long long k, u128 inc, u128 &mul_out, u128 &add_out) {
    u128 acc_mul = 1, acc_add = 0, cur_mul = pcg64_multiplier(), cur_add = inc;
    while (k) {
        if (k & 1) { acc_mul *= cur_mul; acc_add = acc_add * cur_mul + cur_add; }
        cur_add = (cur_mul + 1) * cur_add;
        cur_mul *= cur_mul;
        k >>= 1;
    }
    mul_out = acc_mul; add_out = acc_add;
}

// XSL-RR output of a state that has already been stepped (Lcg128Xsl64::next_u64 after its step)
__host__ __device__ inline unsigned long long pcg64_output(u128 state) {
    const unsigned rot = (unsigned)(state >> 122);
    const unsigned long long folded = (unsigned long long)(state >> 64) ^ (unsigned long long)state;
    return (folded >> rot) | (folded << ((64 - rot) & 63));
}
// rng.gen::<f64>() <= prob  (53 random bits scaled into [0, 1): exact in f64, so the comparison is the crate's)
__host__ __device__ inline bool erc_accept(u128 state, double prob) {
    return (double)(pcg64_output(state) >> 11) * (1.0 / 9007199254740992.0) <= prob;
}

struct ErcParams {
    uint32_t n;
    double prob;
    unsigned long long state_lo, state_hi, inc_lo, inc_hi;   // generator before the first draw
    const LcgStep *__restrict__ lane_step;                   // [33]: l = 0..31 -> l+1 outputs further; [32] -> 32 outputs further
    unsigned long long *row_state;                           // [n][2]: state after P(i) outputs
    uint32_t *lower_deg, *upper_deg, *cursor;                // [n] each
    uint32_t *row_ptr;                                       // [n+1]
    uint32_t *col_idx;
    unsigned long long *totals;                              // [0] nnz, [1] rows with degree > 7, [2] max degree
};

__device__ __forceinline__ unsigned long long erc_row_offset(unsigned long long i, unsigned long long n) {
    return i * (n - 1) - (i * (i - 1)) / 2;                  // pairs drawn before row i
}

__global__ void erc_row_states_kernel(const ErcParams p) {
    const u128 s0 = u128_of(p.state_lo, p.state_hi), inc = u128_of(p.inc_lo, p.inc_hi);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
        u128 m, a;
        lcg_jump(erc_row_offset(i, p.n), inc, m, a);
        const u128 s = m * s0 + a;
        p.row_state[2 * (size_t)i] = (unsigned long long)s;
        p.row_state[2 * (size_t)i + 1] = (unsigned long long)(s >> 64);
        p.lower_deg[i] = 0; p.upper_deg[i] = 0; p.cursor[i] = 0;
    }
}

// kFill == false: count the two halves of every degree.  kFill == true: write the neighbour ids.
template <bool kFill>
__global__ void erc_rows_kernel(const ErcParams p) {
    const int lane = threadIdx.x & 31;
    const uint32_t lane_lt = (1u << lane) - 1u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = (gridDim.x * blockDim.x) >> 5;
    const LcgStep my = p.lane_step[lane], stride = p.lane_step[32];
    const u128 my_mul = u128_of(my.mul_lo, my.mul_hi), my_add = u128_of(my.add_lo, my.add_hi);
    const u128 st_mul = u128_of(stride.mul_lo, stride.mul_hi), st_add = u128_of(stride.add_lo, stride.add_hi);
    for (uint32_t i = warp; i < p.n; i += warps) {
        u128 s = my_mul * u128_of(p.row_state[2 * (size_t)i], p.row_state[2 * (size_t)i + 1]) + my_add;   // output of pair (i, i+1+lane)
        uint32_t found = 0;
        const uint32_t base = kFill ? p.row_ptr[i] + p.lower_deg[i] : 0u;
        for (uint32_t j0 = i + 1; j0 < p.n; j0 += 32) {
            const uint32_t j = j0 + lane;
            const bool acc = j < p.n && erc_accept(s, p.prob);
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, acc);
            if (acc) {
                if (kFill) {
                    p.col_idx[base + found + __popc(bal & lane_lt)] = j;                   // upper part of row i: ascending j
                    p.col_idx[p.row_ptr[j] + atomicAdd(&p.cursor[j], 1u)] = i;            // lower part of row j: sorted afterwards
                } else {
                    atomicAdd(&p.lower_deg[j], 1u);
                }
            }
            found += __popc(bal);
            s = st_mul * s + st_add;
        }
        if (!kFill && lane == 0) p.upper_deg[i] = found;
    }
}

// exclusive scan of the degrees into row_ptr: one CTA (n is at most a few million; each thread takes a block)
__global__ void erc_scan_kernel(const ErcParams p) {
    __shared__ unsigned long long s_part[1024];
    const uint32_t n = p.n, per = (n + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
    unsigned long long sum = 0;
    uint32_t over8 = 0, maxd = 0;
    for (uint32_t v = lo; v < hi; ++v) {
        const uint32_t d = p.lower_deg[v] + p.upper_deg[v];
        sum += d; over8 += d > 7; maxd = max(maxd, d);
    }
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (uint32_t t = 0; t < blockDim.x; ++t) { const unsigned long long x = s_part[t]; s_part[t] = run; run += x; }
        p.totals[0] = run;
        p.row_ptr[n] = (uint32_t)run;
    }
    __syncthreads();
    unsigned long long at = s_part[threadIdx.x];
    for (uint32_t v = lo; v < hi; ++v) { p.row_ptr[v] = (uint32_t)at; at += p.lower_deg[v] + p.upper_deg[v]; }
    if (over8) atomicAdd(&p.totals[1], (unsigned long long)over8);
    atomicMax(&p.totals[2], (unsigned long long)maxd);
}

// lower parts (neighbours k < v, collected in arrival order) -> ascending, one thread per row (lists are short:
// half the mean degree; long ones just take longer)
__global__ void erc_sort_lower_kernel(const ErcParams p) {
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < p.n; v += gridDim.x * blockDim.x) {
        uint32_t *row = p.col_idx + p.row_ptr[v];
        const uint32_t d = p.lower_deg[v];
        for (uint32_t a = 1; a < d; ++a) {
            const uint32_t x = row[a];
            uint32_t b = a;
            while (b > 0 && row[b - 1] > x) { row[b] = row[b - 1]; --b; }
            row[b] = x;
        }
    }
}

}  // namespace nec
