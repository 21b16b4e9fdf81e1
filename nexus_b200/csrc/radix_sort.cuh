// Stable LSD radix sort of (Morton key, primitive index) pairs for sm_100a: replaces the cub::DeviceRadixSort::SortPairs call the
// reference makes (vendor/NexusBVH/NexusBVH/src/Cuda/Setup.cu:63-112) over the same bit windows, [2, 32) for 32-bit keys and
// [1, 64) for 64-bit keys.  A stable sort has exactly one result, so the sorted order - and with it every tree - is unchanged.
//
// Single-pass-per-digit ("onesweep", Adinets & Merrill 2022) organisation, 8-bit digits:
//   * the digit histograms of ALL passes are accumulated by the kernel that generates the keys (morton_kernel), so there is no
//     histogram pass over the keys at all;
//   * one kernel per digit reads every pair once and writes it once: a CTA ranks its tile (warp-level match-any multisplit,
//     stable), obtains the number of equal digits in all earlier tiles by decoupled look-back over a per-tile status word
//     (flag + count packed in 32 bits, so no fence is needed), and scatters through shared memory so that global writes are
//     runs of consecutive addresses per digit;
//   * tiles are handed out by an atomic counter in launch order, which is what makes the look-back deadlock free.
// Algorithmic bytes: (sizeof(key) + 4) x 2 per pair and pass; 4 passes for 32-bit keys, 8 for 64-bit keys.
#pragma once
#include "nx_common.cuh"

#ifndef NX_SORT_THREADS
#define NX_SORT_THREADS 256
#endif
#ifndef NX_SORT_MINB
#define NX_SORT_MINB 5
#endif
#ifndef NX_SORT_ITEMS
#define NX_SORT_ITEMS 16
#endif
constexpr int kSortThreads = NX_SORT_THREADS;      // a tile is kSortThreads x kItems pairs: the longer the tile, the longer the runs written per digit
constexpr int kSortWarps = kSortThreads / 32;
constexpr uint32_t kSortFlagAgg = 1u << 30, kSortFlagPrefix = 2u << 30, kSortValueMask = (1u << 30) - 1u;

template <typename KeyT> struct SortShape;
template <> struct SortShape<uint32_t> { static constexpr int kItems = NX_SORT_ITEMS, kPasses = 4, kFirstBit = 2; };
template <> struct SortShape<uint64_t> { static constexpr int kItems = NX_SORT_ITEMS / 2, kPasses = 8, kFirstBit = 1; };

// digit of pass p: bits [first + 8p, first + 8p + 8) - the top pass simply sees zeros above the key's width
template <typename KeyT> __device__ __forceinline__ uint32_t sort_digit(KeyT key, int pass)
{
    return (uint32_t)(key >> (SortShape<KeyT>::kFirstBit + 8 * pass)) & 0xffu;
}

// exclusive prefix sums of the per-pass digit histograms, in place: hist[p][d] -> number of keys with a smaller digit in pass p
__global__ void __launch_bounds__(256) sort_prefix_kernel(uint32_t* hist, int passes)
{
    __shared__ uint32_t warpSum[8];
    for (int p = 0; p < passes; p++)
    {
        const uint32_t v = hist[p * 256 + threadIdx.x];
        uint32_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(NX_FULL, s, o); if (lane_id() >= (uint32_t)o) s += t; }
        if (lane_id() == 31) warpSum[threadIdx.x >> 5] = s;
        __syncthreads();
        uint32_t before = 0;
        for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) before += warpSum[w];
        hist[p * 256 + threadIdx.x] = before + s - v;
        __syncthreads();
    }
}

template <typename KeyT> struct SortSmem {
    static constexpr int TILE = kSortThreads * SortShape<KeyT>::kItems;
    uint32_t warpHist[kSortWarps][256];     // per warp: running digit counts while ranking, then exclusive offsets over the warps
    uint32_t digitOffset[256];              // first position of the digit inside the sorted tile
    uint32_t digitGlobal[256];              // global position of the digit's first element of this tile, minus digitOffset
    uint32_t tileHist[256];                 // digit counts of the tile (published before the ranking)
    uint32_t warpTotals[kSortWarps];
    uint32_t tile;
    KeyT keys[TILE];
    uint32_t vals[TILE];
};

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads, NX_SORT_MINB) onesweep_kernel(const KeyT* __restrict__ keysIn, const uint32_t* __restrict__ valsIn, KeyT* __restrict__ keysOut,
                                                                uint32_t* __restrict__ valsOut, uint32_t n, int pass, const uint32_t* __restrict__ digitBase,
                                                                uint32_t* status, uint32_t* tileCounter)
{
    constexpr int ITEMS = SortShape<KeyT>::kItems, TILE = kSortThreads * ITEMS;
    extern __shared__ __align__(16) unsigned char sort_raw[];
    SortSmem<KeyT>& S = *reinterpret_cast<SortSmem<KeyT>*>(sort_raw);
    auto& warpHist = S.warpHist; auto& digitOffset = S.digitOffset; auto& digitGlobal = S.digitGlobal; auto& warpTotals = S.warpTotals;
    KeyT* const sKeys = S.keys; uint32_t* const sVals = S.vals; uint32_t& sTile = S.tile;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t lane_lt; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lane_lt));
    if (tid == 0) sTile = atomicAdd(tileCounter, 1u);
    for (uint32_t i = tid; i < kSortWarps * 256; i += kSortThreads) (&warpHist[0][0])[i] = 0u;
    for (uint32_t i = tid; i < 256; i += kSortThreads) S.tileHist[i] = 0u;
    __syncthreads();
    const uint32_t tile = sTile;
    const uint32_t base = tile * (uint32_t)TILE;
    const uint32_t valid = min((uint32_t)TILE, n - base);

    // ---- load, warp striped: item k of lane l in warp w is element w * 32 * ITEMS + k * 32 + l (coalesced, order preserving per round)
    KeyT key[ITEMS]; uint32_t rank[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const uint32_t i = warp * 32u * ITEMS + (uint32_t)k * 32u + lane;
        key[k] = i < valid ? __ldg(keysIn + base + i) : (KeyT)~(KeyT)0;      // padding sorts behind every real key of the tile
    }
    // ---- the tile's digit counts first, published at once: the tiles behind this one look back at them, and by the time this tile has
    // ranked its keys and looks back itself, the tiles before it have published theirs (ncu on the first version, where the counts
    // came out of the ranking: 17 % of the stall samples in the look-back spin)
#pragma unroll
    for (int k = 0; k < ITEMS; k++) atomicAdd(&S.tileHist[sort_digit<KeyT>(key[k], pass)], 1u);
    __syncthreads();
    const bool owner = tid < 256u;
    uint32_t* const mine = status + (size_t)tile * 256u + (tid & 255u);
    const uint32_t count = owner ? S.tileHist[tid] : 0u;
    if (owner) __stcg(mine, (tile == 0u ? kSortFlagPrefix : kSortFlagAgg) | count);

    // ---- rank inside the warp, round by round: lanes with the same digit find each other with match.any; the first of them reads
    // and advances the warp's counter of that digit.  All matches are issued before the first counter is touched: they do not depend
    // on one another, and back to back behind the shared-memory updates they were a third of the kernel's stall samples.
    uint32_t peers[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) peers[k] = __match_any_sync(NX_FULL, sort_digit<KeyT>(key[k], pass));
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const uint32_t d = sort_digit<KeyT>(key[k], pass);
        const uint32_t before = __popc(peers[k] & lane_lt);
        uint32_t old = 0;
        if (before == 0u) { old = warpHist[warp][d]; warpHist[warp][d] = old + __popc(peers[k]); }
        rank[k] = __shfl_sync(NX_FULL, old, __ffs(peers[k]) - 1) + before;
        __syncwarp();
    }
    __syncthreads();

    // ---- thread d (< 256) owns digit d: offsets of the warps, the tile's count, and the look-back over earlier tiles
    if (owner) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) { const uint32_t t = warpHist[w][tid]; warpHist[w][tid] = run; run += t; }
    }
    // exclusive scan of the 256 counts -> position of each digit inside the sorted tile
    uint32_t scan = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(NX_FULL, scan, o); if (lane >= (uint32_t)o) scan += t; }
    if (lane == 31u) warpTotals[warp] = scan;
    __syncthreads();
    if (owner) {
        uint32_t before = 0;
        for (uint32_t w = 0; w < warp; w++) before += warpTotals[w];
        const uint32_t offset = before + scan - count;
        digitOffset[tid] = offset;
        uint32_t earlier = 0;
        if (tile != 0u) {
            for (int t = (int)tile - 1; t >= 0; t--) {
                const volatile uint32_t* p = status + (size_t)t * 256u + tid;
                uint32_t v;
                do { v = *p; } while ((v & ~kSortValueMask) == 0u);
                earlier += v & kSortValueMask;
                if (v & kSortFlagPrefix) break;
            }
            __stcg(mine, kSortFlagPrefix | (earlier + count));
        }
        digitGlobal[tid] = __ldg(digitBase + pass * 256 + tid) + earlier - offset;
    }
    __syncthreads();

    // ---- scatter into the sorted tile in shared memory, then stream it out: consecutive threads write consecutive addresses
    // the values are only read now (they were not needed for ranking, and sixteen fewer live registers is one more resident block)
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const uint32_t i = warp * 32u * ITEMS + (uint32_t)k * 32u + lane;
        const uint32_t d = sort_digit<KeyT>(key[k], pass);
        const uint32_t pos = digitOffset[d] + warpHist[warp][d] + rank[k];
        sKeys[pos] = key[k]; sVals[pos] = i < valid ? __ldg(valsIn + base + i) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const uint32_t i = tid + (uint32_t)k * kSortThreads;
        if (i < valid) {
            const KeyT kk = sKeys[i];
            const uint32_t g = digitGlobal[sort_digit<KeyT>(kk, pass)] + i;
            keysOut[g] = kk; valsOut[g] = sVals[i];
        }
    }
}

// Sorts n pairs; the result ends in (keys, vals) because the number of passes is even.  `hist` holds the digit histograms of all
// passes (filled by the key generator), `status` at least tiles * 256 words, `counters` one word per pass (zero).
template <typename KeyT>
cudaError_t radix_sort_pairs(KeyT* keys, uint32_t* vals, KeyT* keysAlt, uint32_t* valsAlt, uint32_t n, uint32_t* hist, uint32_t* status, uint32_t* counters, cudaStream_t s)
{
    constexpr int TILE = kSortThreads * SortShape<KeyT>::kItems, PASSES = SortShape<KeyT>::kPasses;
    static_assert(PASSES % 2 == 0, "the result must land in the primary buffers");
    const uint32_t tiles = (n + TILE - 1) / TILE;
    static bool attr = false;   // a property of the loaded kernel, not of a context
    if (!attr) { cudaFuncSetAttribute((const void*)onesweep_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem<KeyT>)); attr = true; }
    sort_prefix_kernel<<<1, 256, 0, s>>>(hist, PASSES);
    for (int p = 0; p < PASSES; p++) {
        cudaError_t e = cudaMemsetAsync(status, 0, (size_t)tiles * 256u * sizeof(uint32_t), s);
        if (e != cudaSuccess) return e;
        if (p & 1) onesweep_kernel<KeyT><<<tiles, kSortThreads, sizeof(SortSmem<KeyT>), s>>>(keysAlt, valsAlt, keys, vals, n, p, hist, status, counters + p);
        else onesweep_kernel<KeyT><<<tiles, kSortThreads, sizeof(SortSmem<KeyT>), s>>>(keys, vals, keysAlt, valsAlt, n, p, hist, status, counters + p);
    }
    return cudaGetLastError();
}
template <typename KeyT> inline size_t radix_sort_status_words(uint32_t n)
{
    constexpr int TILE = kSortThreads * SortShape<KeyT>::kItems;
    return (size_t)((n + TILE - 1) / TILE) * 256u;
}
