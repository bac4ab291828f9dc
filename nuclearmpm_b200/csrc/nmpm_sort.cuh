// K0 — stable LSD radix sort of (cell key, slot) pairs, 8 bits per pass, hand-written for sm_100a.
//
// The reference never sorts (it walks particles in input order, src/nclr.h:112); binning is what
// makes the GPU scatter/gather coalesced.  The order produced here is exactly
// std::stable_sort by key = ascending (key, previous slot), reproduced on the CPU by
// nclr_oracle_stable_sort (oracle/nclr_oracle.c) and compared bit for bit in tests/.
//
// Per pass (3 kernels):
//   sort_hist    : per-tile digit histogram (shared-memory integer atomics)  -> hist[digit][tile]
//   sort_scan    : one CTA per digit scans its row over tiles; the last CTA to finish scans the
//                  digit totals (threadfence + counter) -> digit_base
// Digit width (NMPM_RADIX_BITS): 10-bit digits sort cfg4's 28-bit keys (+1 bit for the special keys) in 3 passes instead
// of 4, but every pass is slower (1 024 scan rows, 32-byte instead of 64-byte runs in the scatter): measured 0.76 against
// 0.71 ms per 16.8 M pairs on cfg4 and 0.110 against 0.122 ms on cfg3 (gpurun r3q) — 8 stays.
//   sort_scatter : warp-synchronous stable ranking (__match_any_sync) + scatter
// Integer atomics only: the result is deterministic.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nmpm {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys per CTA
#ifndef NMPM_RADIX_BITS
#define NMPM_RADIX_BITS 8
#endif
constexpr int kRadixBits = NMPM_RADIX_BITS;
constexpr int kRadix = 1 << kRadixBits;
constexpr uint32_t kRadixMask = kRadix - 1;

__global__ void __launch_bounds__(kSortThreads) sort_hist(const uint32_t* __restrict__ keys, uint32_t n, int shift,
                                                          uint32_t* __restrict__ hist, uint32_t ntiles) {
    __shared__ uint32_t h[kRadix];
    for (int d = threadIdx.x; d < kRadix; d += kSortThreads) h[d] = 0;
    __syncthreads();
    const uint32_t start = blockIdx.x * kSortTile;
    // (plain shared-memory atomics: aggregating the lanes of equal digit with __match_any_sync first — cell-sorted keys
    // repeat their digit ~8 times in a row — was measured slower, 0.184 against 0.178 ms of sort per step, gpurun r3r)
    uint32_t k[kSortItems];  // all loads in flight before the first atomic (the atomics order the loads behind them)
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t idx = start + i * kSortThreads + threadIdx.x;
        k[i] = (idx < n) ? __ldg(keys + idx) : 0u;
    }
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t idx = start + i * kSortThreads + threadIdx.x;
        if (idx < n) atomicAdd(&h[(k[i] >> shift) & kRadixMask], 1u);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < kRadix; d += kSortThreads) hist[(size_t) d * ntiles + blockIdx.x] = h[d];
}

// grid = kRadix CTAs (one per digit).  Exclusive scan of hist[d][0..ntiles) in place; row totals to
// digit_total; the last CTA done turns digit_total into the exclusive digit_base.
__global__ void __launch_bounds__(256) sort_scan(uint32_t* __restrict__ hist, uint32_t ntiles,
                                                 uint32_t* __restrict__ digit_total, uint32_t* __restrict__ digit_base,
                                                 unsigned int* __restrict__ done_counter) {
    __shared__ uint32_t warp_sums[8];
    __shared__ uint32_t carry;
    __shared__ bool is_last;
    uint32_t* row = hist + (size_t) blockIdx.x * ntiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < ntiles; base += 256) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = (i < ntiles) ? row[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w)
            if (w < warp) woff += warp_sums[w];
        const uint32_t c = carry;
        if (i < ntiles) row[i] = c + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = c + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        digit_total[blockIdx.x] = carry;
        __threadfence();
        const unsigned int t = atomicAdd(done_counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // exclusive scan of the kRadix digit totals: kRadix / 256 consecutive digits per thread
        constexpr int PER = kRadix / 256;
        uint32_t v[PER], sum = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            v[q] = ((volatile uint32_t*) digit_total)[threadIdx.x * PER + q];
            sum += v[q];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w)
            if (w < warp) woff += warp_sums[w];
        uint32_t run = woff + incl - sum;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            digit_base[threadIdx.x * PER + q] = run;
            run += v[q];
        }
        if (threadIdx.x == 0) *done_counter = 0;  // re-arm for the next pass
    }
}

// vals_in == nullptr means the identity (first pass).
template <int MINB>
__global__ void __launch_bounds__(kSortThreads, MINB) sort_scatter(const uint32_t* __restrict__ keys_in,
                                                             const uint32_t* __restrict__ vals_in,
                                                             uint32_t* __restrict__ keys_out,
                                                             uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                             const uint32_t* __restrict__ hist, uint32_t ntiles,
                                                             const uint32_t* __restrict__ digit_base) {
    constexpr int kWarps = kSortThreads / 32;
    __shared__ uint32_t warp_cnt[kWarps][kRadix];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < kWarps * kRadix; k += kSortThreads) (&warp_cnt[0][0])[k] = 0;
    __syncthreads();

    // stable order inside the tile: (warp, item, lane); each warp owns a contiguous 512-key run
    const uint32_t wstart = blockIdx.x * kSortTile + warp * (32 * kSortItems);
    uint32_t key[kSortItems], val[kSortItems];
    uint16_t rank[kSortItems];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {  // all loads in flight before the ranking (its __syncwarp()s order them)
        const uint32_t idx = wstart + i * 32 + lane;
        const bool valid = idx < n;
        key[i] = valid ? __ldg(keys_in + idx) : 0xFFFFFFFFu;
        val[i] = valid ? (vals_in ? __ldg(vals_in + idx) : idx) : 0u;
    }
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t idx = wstart + i * 32 + lane;
        const bool valid = idx < n;
        const uint32_t digit = (key[i] >> shift) & kRadixMask;
        // invalid lanes get a private pseudo-digit so that they match nobody
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : ((uint32_t) kRadix | (uint32_t) lane));
        const uint32_t before = warp_cnt[warp][digit];
        rank[i] = (uint16_t) (before + __popc(peers & lt_mask));
        __syncwarp();
        if (valid && (peers & lt_mask) == 0u) warp_cnt[warp][digit] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    for (int d = threadIdx.x; d < kRadix; d += kSortThreads) {
        // exclusive prefix over warps per digit, seeded with the global base of (digit, tile)
        uint32_t run = digit_base[d] + hist[(size_t) d * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t t = warp_cnt[w][d];
            warp_cnt[w][d] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t idx = wstart + i * 32 + lane;
        if (idx < n) {
            const uint32_t pos = warp_cnt[warp][(key[i] >> shift) & kRadixMask] + rank[i];
            keys_out[pos] = key[i];
            vals_out[pos] = val[i];
        }
    }
}

struct SortWorkspace {
    uint32_t *keys_a = nullptr, *keys_b = nullptr, *vals_a = nullptr, *vals_b = nullptr;
    uint32_t *hist = nullptr, *digit_total = nullptr, *digit_base = nullptr;
    unsigned int* done_counter = nullptr;
    uint32_t ntiles = 0;
    // CTAs per SM sort_scatter is compiled for (env NMPM_SORT_MINB: experiments).  3 (80 registers, 40 B of spills):
    // 0.163 ms of sort per step on cfg4 against 0.170 without a bound (97 registers) and 0.214 at 2 (gpurun r3t, r3w)
    int scatter_minb = 3;
};

// Sorts ws.keys_a (n keys) carrying slot indices; returns pointers to the sorted keys / permutation
// (they alias ws buffers).  `key_bits` = number of significant key bits.  Returns launches issued.
inline int radix_sort_pairs(SortWorkspace& ws, uint32_t n, int key_bits, cudaStream_t st, uint32_t** keys_sorted,
                            uint32_t** perm) {
    // one spare bit above the cell keys so that the special keys (out of grid, migrated away) stay on top
    const int passes = (key_bits + kRadixBits) / kRadixBits;
    uint32_t *kin = ws.keys_a, *kout = ws.keys_b, *vin = nullptr, *vout = ws.vals_a;
    uint32_t* vother = ws.vals_b;
    int launches = 0;
    for (int p = 0; p < passes; ++p) {
        const int shift = kRadixBits * p;
        sort_hist<<<ws.ntiles, kSortThreads, 0, st>>>(kin, n, shift, ws.hist, ws.ntiles);
        sort_scan<<<kRadix, 256, 0, st>>>(ws.hist, ws.ntiles, ws.digit_total, ws.digit_base, ws.done_counter);
        if (ws.scatter_minb >= 3)
            sort_scatter<3><<<ws.ntiles, kSortThreads, 0, st>>>(kin, vin, kout, vout, n, shift, ws.hist, ws.ntiles, ws.digit_base);
        else
            sort_scatter<2><<<ws.ntiles, kSortThreads, 0, st>>>(kin, vin, kout, vout, n, shift, ws.hist, ws.ntiles, ws.digit_base);
        launches += 3;
        uint32_t* t = kin;
        kin = kout;
        kout = t;
        vin = vout;
        vout = vother;
        vother = vin;
    }
    *keys_sorted = kin;
    *perm = vin;
    return launches;
}

}  // namespace nmpm
