// kernels_sort.cuh — the device primitives of the set build, hand-written for sm_100a (no library sort / scan on the hot
// path): a stable least-significant-digit radix sort of (64-bit key, 32-bit value) pairs and single-pass chained scans.
//
// Radix sort ("one sweep" per digit): the digit histograms of ALL passes are taken in one read of the keys (fused with the
// Morton-code generation of the set build, k_sort_prepare); every pass is then ONE kernel: a block takes the next tile
// (atomic ticket, so that a tile only ever waits for tiles that already run), ranks its keys stably (warp-striped items,
// __match_any_sync multi-split, per-warp digit counters in shared memory), publishes its per-digit counts, obtains the
// counts of all earlier tiles by decoupled look-back (status word = count | PARTIAL / INCLUSIVE flag) and scatters through
// shared memory so that the global writes of one digit are contiguous.  Up to RS_MAXSEG independent segments (the two
// resolution levels of createGaussianSets, DmsaOptimizer.h:81-86) are sorted by the same launches.
//
// Chained scans: k_segment (run heads of the sorted codes: leaf numbering, leaf starts and the ring-id test of
// DmsaOptimizer.h:303-307 in one pass), k_emit (acceptance + set numbering + emission), k_scan_excl (plain exclusive sum).
#pragma once
#include <cuda_runtime.h>

#include "pdl.cuh"

#include "kernels_sets.cuh"

namespace dmsa {

typedef unsigned long long u64_t;
typedef unsigned int u32_t;

#define RS_BINS 256
#define RS_T 256
#define RS_ITEMS 16
#define RS_TILE (RS_T * RS_ITEMS)
#define RS_MAXPASS 8
#define RS_MAXSEG 2
#define RS_PART (1u << 30)
#define RS_INCL (2u << 30)
#define RS_VMASK ((1u << 30) - 1)
#define RS_SMEM (RS_TILE * 12 + (RS_T / 32) * RS_BINS * 4 + 2 * RS_BINS * 4)

struct SortSeg {
    u64_t* keyA;
    u64_t* keyB;
    u32_t* valA;  // may be null for pass 0: the value of item i is i
    u32_t* valB;
};
struct SortArgs {
    SortSeg seg[RS_MAXSEG];
    int nseg, n, tiles, npass;
    int iota;     // pass 0 takes the value of item i to be i (valA is then only written, by pass 1)
    u32_t* hist;  // [seg][RS_MAXPASS][RS_BINS] digit histograms of every pass (zeroed before k_sort_prepare / k_sort_hist)
    u32_t* look;  // [pass][seg][tiles][RS_BINS] look-back status words (zeroed per sort)
    int* ticket;  // [RS_MAXPASS] tile tickets (zeroed per sort)
    int pass;
};

__device__ __forceinline__ u32_t lanemask_lt() {
    u32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ u32_t ld_volatile_u32(const u32_t* p) {
    u32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// status words are published with an atomic exchange: performed at the L2 right away (a plain store may sit in the SM's
// write path for microseconds while thousands of readers poll for it)
__device__ __forceinline__ void st_volatile_u32(u32_t* p, u32_t v) { atomicExch(p, v); }
__device__ __forceinline__ u64_t ld_volatile_u64(const u64_t* p) {
    u64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(u64_t* p, u64_t v) { atomicExch(p, v); }

// exclusive prefix sum over the NT threads of a block (s_w: NT / 32 words of shared scratch); also returns the block total
template <int NT>
__device__ __forceinline__ u32_t block_excl_scan(u32_t v, u32_t* s_w, u32_t* total = nullptr) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();  // s_w may still be read by a previous scan
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    u32_t base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < NT / 32; ++k) {
        const u32_t x = s_w[k];
        if (k < w) base += x;
        tot += x;
    }
    if (total) *total = tot;
    return base + inc - v;
}
__device__ __forceinline__ u32_t block_excl_scan_256(u32_t v, u32_t* s_w, u32_t* total = nullptr) { return block_excl_scan<256>(v, s_w, total); }

// warp-aggregated shared-memory histogram update of one digit per lane (equal digits of a warp cost one atomic)
__device__ __forceinline__ void hist_add_warp(u32_t* sh, u32_t d, bool valid) {
    const u32_t key = valid ? d : 0xffffffffu;
    const u32_t m = __match_any_sync(0xffffffffu, key);
    if (valid && (threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(sh + d, __popc(m));
}

// ---- pass 0 input of the set build: Morton codes (kernels_sets.cuh k_morton) of both levels + the digit histograms of every pass
struct PrepArgs {
    const int* keys_all;      // [level][3 N] voxel keys
    const LevelInfo* infos;
    int level[RS_MAXSEG];
    const int* ring;          // ring id per point
    const int* flags;         // flags[1] != 0: some ring id lies outside [0, 65535]
};
// The ring id of a point rides in bits 48..63 of its sort key when the Morton code leaves them free (octree depth <= 15) and
// every ring id fits 16 bits: the sort only looks at the code bits, and k_segment's ring test (DmsaOptimizer.h:303-307)
// then needs no gathers.  Decided on the device, identically by k_sort_prepare and k_segment.
#define RING_SHIFT 48
__device__ __forceinline__ bool ring_packed(const LevelInfo* info, const int* flags) { return info->depth <= 15 && flags[1] == 0; }
__global__ void __launch_bounds__(RS_T) k_sort_prepare(SortArgs a, PrepArgs pa) {
    DMSA_PDL_ENTER();
    __shared__ u32_t sh[RS_MAXPASS * RS_BINS];
    const int seg = blockIdx.y, lvl = pa.level[seg];
    const LevelInfo* __restrict__ info = pa.infos + lvl;
    const int* __restrict__ keys = pa.keys_all + (size_t)3 * a.n * lvl;
    u64_t* __restrict__ code = a.seg[seg].keyA;
    for (int q = threadIdx.x; q < a.npass * RS_BINS; q += RS_T) sh[q] = 0;
    __syncthreads();
    const long long lo0 = info->lo[0], lo1 = info->lo[1], lo2 = info->lo[2];
    const u64_t inval = 1ull << (3 * info->depth);  // non-finite points sort behind every leaf (PCL skips them)
    const bool packed = ring_packed(info, pa.flags);
    const int base = blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int i = base + k * RS_T + threadIdx.x;
        const bool valid = i < a.n;
        u64_t m = 0;
        if (valid) {
            const int kx = keys[3 * (size_t)i], ky = keys[3 * (size_t)i + 1], kz = keys[3 * (size_t)i + 2];
            if (kx == (-2147483647 - 1)) {
                m = inval;
            } else {
                const u64_t x = (u64_t)((long long)kx - lo0), y = (u64_t)((long long)ky - lo1), z = (u64_t)((long long)kz - lo2);
                m = (spread3(x) << 2) | (spread3(y) << 1) | spread3(z);
            }
            code[i] = (packed && pa.ring) ? (m | ((u64_t)(u32_t)pa.ring[i] << RING_SHIFT)) : m;
        }
        // the low digits of a Morton code are well spread (plain shared-memory atomics); the top digits take few values
        // (equal digits of a warp are counted once)
        for (int p = 0; p < a.npass; ++p) {
            const u32_t d = (u32_t)(m >> (8 * p)) & 255u;
            if (p + 2 < a.npass) {
                if (valid) atomicAdd(sh + p * RS_BINS + d, 1u);
            } else {
                hist_add_warp(sh + p * RS_BINS, d, valid);
            }
        }
    }
    __syncthreads();
    u32_t* __restrict__ gh = a.hist + (size_t)seg * RS_MAXPASS * RS_BINS;
    for (int q = threadIdx.x; q < a.npass * RS_BINS; q += RS_T)
        if (sh[q]) atomicAdd(gh + (q / RS_BINS) * RS_BINS + (q % RS_BINS), sh[q]);
}
// the same histograms for keys that already exist (keyA of every segment)
__global__ void __launch_bounds__(RS_T) k_sort_hist(SortArgs a) {
    DMSA_PDL_ENTER();
    __shared__ u32_t sh[RS_MAXPASS * RS_BINS];
    const int seg = blockIdx.y;
    const u64_t* __restrict__ code = a.seg[seg].keyA;
    for (int q = threadIdx.x; q < a.npass * RS_BINS; q += RS_T) sh[q] = 0;
    __syncthreads();
    const int base = blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int i = base + k * RS_T + threadIdx.x;
        const bool valid = i < a.n;
        const u64_t m = valid ? code[i] : 0;
        for (int p = 0; p < a.npass; ++p) hist_add_warp(sh + p * RS_BINS, (u32_t)(m >> (8 * p)) & 255u, valid);
    }
    __syncthreads();
    u32_t* __restrict__ gh = a.hist + (size_t)seg * RS_MAXPASS * RS_BINS;
    for (int q = threadIdx.x; q < a.npass * RS_BINS; q += RS_T)
        if (sh[q]) atomicAdd(gh + (q / RS_BINS) * RS_BINS + (q % RS_BINS), sh[q]);
}

#ifdef DMSA_TIMELINE
__device__ unsigned long long g_dbg_t[16384];
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define DMSA_TLK(kern, slot) do { if (DMSA_TIMELINE == (kern) && threadIdx.x == 0 && blockIdx.x < 2048) g_dbg_t[8 * blockIdx.x + (slot)] = gtimer(); } while (0)
#else
#define DMSA_TLK(kern, slot) do { } while (0)
#endif
#define DMSA_TL(slot) DMSA_TLK(1, slot)
// ---- one digit pass over every segment --------------------------------------------------------------------------------
// 3 blocks per SM (<= 85 registers, 58 KB of shared memory each): the 346 tiles of the two 705 k-point levels of BASELINE
// config 2 are resident together, one wave.
__global__ void __launch_bounds__(RS_T, 3) k_sort_pass(SortArgs a) {
    DMSA_PDL_ENTER();
    extern __shared__ __align__(16) unsigned char rs_smem[];
    u64_t* s_keys = reinterpret_cast<u64_t*>(rs_smem);           // [RS_TILE]
    u32_t* s_vals = reinterpret_cast<u32_t*>(s_keys + RS_TILE);  // [RS_TILE]
    u32_t* s_hist = s_vals + RS_TILE;                            // [warps][RS_BINS]
    u32_t* s_tileoff = s_hist + (RS_T / 32) * RS_BINS;           // [RS_BINS]
    u32_t* s_gbase = s_tileoff + RS_BINS;                        // [RS_BINS]
    __shared__ u32_t s_w[RS_T / 32];
    __shared__ int s_t;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DMSA_TL(0);
    if (tid == 0) s_t = atomicAdd(a.ticket + a.pass, 1);
    for (int q = tid; q < (RS_T / 32) * RS_BINS; q += RS_T) s_hist[q] = 0;
    __syncthreads();
    DMSA_TL(1);
    const int t = s_t, seg = t / a.tiles, tile = t - seg * a.tiles;
    const bool even = (a.pass & 1) == 0;
    const u64_t* __restrict__ kin = even ? a.seg[seg].keyA : a.seg[seg].keyB;
    u64_t* __restrict__ kout = even ? a.seg[seg].keyB : a.seg[seg].keyA;
    const u32_t* __restrict__ vin = even ? a.seg[seg].valA : a.seg[seg].valB;
    u32_t* __restrict__ vout = even ? a.seg[seg].valB : a.seg[seg].valA;
    const bool iota = a.pass == 0 && a.iota;
    const int shift = 8 * a.pass;
    const int base = tile * RS_TILE, cnt = min(RS_TILE, a.n - base);
    const int wbase = warp * 32 * RS_ITEMS;
    u64_t key[RS_ITEMS];
    u32_t rank2[RS_ITEMS / 2];  // two 16-bit ranks per word
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int p = wbase + i * 32 + lane;
        key[i] = p < cnt ? kin[base + p] : ~0ull;
    }
    // stable ranking inside the warp: items in (i, lane) order; equal digits of one step are ranked by lane
    u32_t* __restrict__ wh = s_hist + warp * RS_BINS;
    const u32_t lt = lanemask_lt();
    if (key[RS_ITEMS - 1] == 12345ull) DMSA_TL(7);  // (keeps the loads ahead of the time stamp)
    DMSA_TL(2);
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const bool valid = wbase + i * 32 + lane < cnt;
        const u32_t d = valid ? ((u32_t)(key[i] >> shift) & 255u) : 256u;
        const u32_t m = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(m) - 1;
        u32_t old = 0;
        if (lane == leader && valid) {
            old = wh[d];
            wh[d] = old + __popc(m);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        const u32_t r = old + __popc(m & lt);
        if (i & 1)
            rank2[i >> 1] |= r << 16;
        else
            rank2[i >> 1] = r;
        __syncwarp();
    }
    // the values: independent loads, in flight during the look-back
    u32_t val[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int p = wbase + i * 32 + lane;
        val[i] = (p < cnt && !iota) ? vin[base + p] : (u32_t)(base + p);
    }
    __syncthreads();
    DMSA_TL(3);
    // thread d: exclusive prefix of digit d over the warps, tile count
    u32_t count = 0;
#pragma unroll
    for (int w = 0; w < RS_T / 32; ++w) {
        const u32_t c = s_hist[w * RS_BINS + tid];
        s_hist[w * RS_BINS + tid] = count;
        count += c;
    }
    // decoupled look-back over the earlier tiles of this segment (thread d walks digit d); the partial count is published
    // first so that later tiles never wait for this tile's own walk
    u32_t* __restrict__ look = a.look + (((size_t)a.pass * a.nseg + seg) * a.tiles) * RS_BINS;
    st_volatile_u32(look + (size_t)tile * RS_BINS + tid, count | (tile == 0 ? RS_INCL : RS_PART));
    const u32_t tileoff = block_excl_scan_256(count, s_w);
    const u32_t binstart = block_excl_scan_256(a.hist[((size_t)seg * RS_MAXPASS + a.pass) * RS_BINS + tid], s_w);
    u32_t excl = 0;
    if (tile > 0) {
        // all tiles of a pass are resident together and publish their partial counts at about the same time: the walk back
        // to the nearest inclusive word is a chain of L2 round trips, so eight status words are fetched per step
        bool done = false;
        for (int j = tile - 1; j >= 0 && !done; j -= 8) {
            u32_t v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = (j - q >= 0) ? ld_volatile_u32(look + (size_t)(j - q) * RS_BINS + tid) : RS_INCL;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (done) continue;
                while (v[q] == 0) {
                    __nanosleep(200);
                    v[q] = ld_volatile_u32(look + (size_t)(j - q) * RS_BINS + tid);
                }
                excl += v[q] & RS_VMASK;
                if (v[q] & RS_INCL) done = true;
            }
        }
        st_volatile_u32(look + (size_t)tile * RS_BINS + tid, (excl + count) | RS_INCL);
    }
    s_tileoff[tid] = tileoff;
    s_gbase[tid] = binstart + excl - tileoff;
    __syncthreads();
    DMSA_TL(4);
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int p = wbase + i * 32 + lane;
        if (p < cnt) {
            const u32_t d = (u32_t)(key[i] >> shift) & 255u;
            const u32_t r = (i & 1) ? (rank2[i >> 1] >> 16) : (rank2[i >> 1] & 0xffffu);
            const u32_t pos = s_tileoff[d] + wh[d] + r;
            s_keys[pos] = key[i];
            s_vals[pos] = val[i];
        }
    }
    __syncthreads();
    DMSA_TL(5);
#pragma unroll 4
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int p = k * RS_T + tid;
        if (p < cnt) {
            const u64_t kk = s_keys[p];
            const u32_t o = s_gbase[(u32_t)(kk >> shift) & 255u] + (u32_t)p;
            kout[o] = kk;
            vout[o] = s_vals[p];
        }
    }
    DMSA_TL(6);
}

// ---- chained scans ------------------------------------------------------------------------------------------------------
// Status word of a tile: bits 62..63 flag (1 partial, 2 inclusive), payload below.  `combine` semantics are supplied by the
// callers (sum, or sum | max packed in the payload).
#define CS_T 512
#define CS_ITEMS 8
#define CS_TILE (CS_T * CS_ITEMS)
#define EM_T 256  // k_emit: few leaves per thread (every accepted leaf costs a chain of dependent gathers)
#define EM_ITEMS 4
#define EM_TILE (EM_T * EM_ITEMS)
#define CS_PART (1ull << 62)
#define CS_INCL (2ull << 62)
#define CS_PMASK ((1ull << 62) - 1)

// payload of k_segment: low 31 bits = number of run heads, next 31 bits = 1 + position of the last head (0: none)
__device__ __forceinline__ u64_t seg_combine(u64_t earlier, u64_t later) {
    const u64_t sum = (earlier & 0x7fffffffull) + (later & 0x7fffffffull);
    const u64_t me = earlier >> 31, ml = later >> 31;
    return sum | ((ml > me ? ml : me) << 31);
}
__device__ __forceinline__ u64_t sum_combine(u64_t a, u64_t b) { return a + b; }

// Decoupled look-back with the WHOLE block reading status words: thread q takes the q-th nearest earlier tile, so up to NT
// predecessors cost one L2 round trip plus a block reduction (all tiles of a scan are usually resident together and publish
// their partial aggregates at about the same time; a warp-wide window would walk them 32 at a time).  COMBINE must be
// commutative with 0 as identity.  Returns the exclusive prefix to every thread.  s_red: NT / 32 + 2 words of shared scratch.
template <u64_t (*COMBINE)(u64_t, u64_t), int NT>
__device__ __forceinline__ u64_t block_lookback(u64_t* status, int tile, u64_t aggregate, u64_t* s_red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) st_volatile_u64(status + tile, aggregate | (tile == 0 ? CS_INCL : CS_PART));
    u64_t excl = 0;
    for (int hi = tile - 1; hi >= 0; hi -= NT) {
        const int jj = hi - tid;
        u64_t v = CS_INCL;  // tiles before the first count as an inclusive 0
        if (jj >= 0) {
            v = ld_volatile_u64(status + jj);
            while ((v >> 62) == 0) {  // (back off: thousands of threads poll while the tiles of a wave publish)
                __nanosleep(200);
                v = ld_volatile_u64(status + jj);
            }
        }
        // nearest inclusive predecessor of this chunk (smallest tid): everything behind it is already folded into it
        const u32_t incl = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        __syncthreads();  // s_red is free
        DMSA_TLK(2, 5);
        if (lane == 0) s_red[warp] = incl ? (u64_t)(warp * 32 + __ffs(incl) - 1) : (u64_t)NT;
        __syncthreads();
        int stop = NT;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) stop = min(stop, (int)s_red[w]);
        u64_t c = (tid <= stop) ? (v & CS_PMASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c = COMBINE(c, __shfl_xor_sync(0xffffffffu, c, o));
        __syncthreads();
        if (lane == 0) s_red[warp] = c;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) excl = COMBINE(excl, s_red[w]);
        if (stop < NT) break;
    }
    if (tid == 0 && tile > 0) st_volatile_u64(status + tile, COMBINE(excl, aggregate) | CS_INCL);
    return excl;
}

// Run heads of the sorted Morton codes of every segment, in one pass:
//   scan[i]       = number of heads at positions <= i (leaf of member i = scan[i] - 1)            [was k_heads + scan]
//   raw_start[c]  = position of leaf c's first member, raw_start[R] = n_valid, info->R, n_valid    [was k_raw_starts]
//   raw_diff[c]   = 1 if some member's ring id differs from the first member's                     [was k_ring_diff]
#define SG_T 256
#define SG_ITEMS 16
#define SG_TILE (SG_T * SG_ITEMS)
struct SegmentArgs {
    const u64_t* code[RS_MAXSEG];
    const u32_t* idx[RS_MAXSEG];
    int* scan[RS_MAXSEG];
    int* raw_start[RS_MAXSEG];
    int* raw_diff[RS_MAXSEG];  // zeroed before the launch
    LevelInfo* infos;
    int level[RS_MAXSEG];
    const int* ring;
    const int* flags;
    int n, tiles;
    u64_t* status;  // [seg][tiles], zeroed
    int* ticket;    // zeroed
};
__global__ void __launch_bounds__(SG_T, 3) k_segment(SegmentArgs a) {
    DMSA_PDL_ENTER();
    __shared__ u64_t s_agg[SG_T / 32 + 2];
    __shared__ u64_t s_code[SG_TILE + SG_TILE / SG_ITEMS + 2 + 18];  // tile + halo, one pad word per SG_ITEMS (conflict-free thread-blocked reads)
    __shared__ int s_t;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DMSA_TLK(2, 0);
    if (tid == 0) s_t = atomicAdd(a.ticket, 1);
    __syncthreads();
    DMSA_TLK(2, 1);
    const int t = s_t, seg = t / a.tiles, tile = t - seg * a.tiles;
    LevelInfo* __restrict__ info = a.infos + a.level[seg];
    const u64_t* __restrict__ code = a.code[seg];
    const u32_t* __restrict__ idx = a.idx[seg];
    const u64_t inval = 1ull << (3 * info->depth);
    const bool packed = ring_packed(info, a.flags);
    const u64_t cmask = packed ? ((1ull << RING_SHIFT) - 1) : ~0ull;
    const int n = a.n;
    const int i0 = tile * SG_TILE + tid * SG_ITEMS;
    // run heads and validity of this thread's items (bit k), ring ids of the members
    u32_t heads = 0, valid = 0;
    int rg[SG_ITEMS];
    bool next_invalid;  // the item behind this thread's last one is not a member of any leaf
    {
        // coalesced loads of the tile's codes [tbase - 1, tbase + SG_TILE] into shared memory; slot of position q (relative to
        // tbase - 1) is q + q / SG_ITEMS
        const int tbase = tile * SG_TILE;
#pragma unroll
        for (int k = 0; k < SG_ITEMS + 1; ++k) {
            const int q = k * SG_T + tid;
            if (q < SG_TILE + 2) {
                const int gi = tbase - 1 + q;
                s_code[q + q / SG_ITEMS] = (gi >= 0 && gi < n) ? code[gi] : ~0ull;
            }
        }
        __syncthreads();
        u64_t c[SG_ITEMS + 2];
#pragma unroll
        for (int k = 0; k < SG_ITEMS + 2; ++k) {
            const int q = tid * SG_ITEMS + k;
            c[k] = s_code[q + q / SG_ITEMS];
        }
#pragma unroll
        for (int k = 0; k < SG_ITEMS; ++k) {
            const u64_t cur = c[k + 1] & cmask, prv = c[k] & cmask;
            const bool v = (i0 + k < n) && cur < inval;
            if (v) valid |= 1u << k;
            if (v && (i0 + k == 0 || prv != cur)) heads |= 1u << k;
            rg[k] = (int)(c[k + 1] >> RING_SHIFT);
        }
        next_invalid = (i0 + SG_ITEMS >= n) || (c[SG_ITEMS + 1] & cmask) >= inval;
    }
    if (!packed && a.ring) {  // ring ids by gather (independent loads, in flight while the scan and the look-back run)
#pragma unroll
        for (int k = 0; k < SG_ITEMS; ++k) rg[k] = (i0 + k < n) ? (int)idx[i0 + k] : 0;
#pragma unroll
        for (int k = 0; k < SG_ITEMS; ++k) rg[k] = (i0 + k < n) ? a.ring[rg[k]] : 0;
    }
    u64_t mine = 0;  // packed (count, 1 + last head position)
    if (heads) mine = (u64_t)__popc(heads) | ((u64_t)(i0 + (31 - __clz(heads)) + 1) << 31);
    // inclusive warp scan of the packed pairs, then across warps
    u64_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64_t up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = seg_combine(up, inc);
    }
    if (lane == 31) s_agg[warp] = inc;
    __syncthreads();
    DMSA_TLK(2, 2);
    u64_t wex = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SG_T / 32; ++w) {
        const u64_t x = s_agg[w];
        if (w < warp) wex = seg_combine(wex, x);
        total = seg_combine(total, x);
    }
    const u64_t tiles_before = block_lookback<seg_combine, SG_T>(a.status + (size_t)seg * a.tiles, tile, total, s_agg);
    DMSA_TLK(2, 3);
    // exclusive prefix of this thread = tiles before | warps before | lanes before   (the zero word is the identity)
    u64_t ex = seg_combine(tiles_before, wex);
    const u64_t lanes_before = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane > 0) ex = seg_combine(ex, lanes_before);
    int cnt = (int)(ex & 0x7fffffffull);
    const int last = (int)(ex >> 31) - 1;
    // ring id of the run head that precedes this thread's items
    int hring = 0;
    if ((valid & 1u) && !(heads & 1u) && last >= 0) hring = packed ? (int)(code[last] >> RING_SHIFT) : (a.ring ? a.ring[idx[last]] : 0);
    DMSA_TLK(2, 4);
    int* __restrict__ scan = a.scan[seg];
    int* __restrict__ raw_start = a.raw_start[seg];
    int* __restrict__ raw_diff = a.raw_diff[seg];
    int sc[SG_ITEMS];
#pragma unroll
    for (int k = 0; k < SG_ITEMS; ++k) {
        const int i = i0 + k;
        if (heads & (1u << k)) {
            raw_start[cnt] = i;
            ++cnt;
            hring = rg[k];
        }
        sc[k] = cnt;
        if ((valid & (1u << k)) && rg[k] != hring) raw_diff[cnt - 1] = 1;  // DmsaOptimizer.h:303-307
        const bool nxt_invalid = (k + 1 < SG_ITEMS) ? !(valid & (1u << ((k + 1) & 31))) : next_invalid;
        if ((valid & (1u << k)) && nxt_invalid) {  // the last member of the last leaf
            info->n_valid = i + 1;
            info->R = cnt;
            raw_start[cnt] = i + 1;
        }
    }
    if (i0 + SG_ITEMS <= n) {
        int4* __restrict__ dst = reinterpret_cast<int4*>(scan + i0);
#pragma unroll
        for (int k = 0; k < SG_ITEMS / 4; ++k) dst[k] = make_int4(sc[4 * k], sc[4 * k + 1], sc[4 * k + 2], sc[4 * k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < SG_ITEMS; ++k)
            if (i0 + k < n) scan[i0 + k] = sc[k];
    }
    DMSA_TLK(2, 6);
#ifdef DMSA_TIMELINE
    if (DMSA_TIMELINE == 2 && threadIdx.x == 0 && blockIdx.x < 2048) g_dbg_t[8 * blockIdx.x + 7] = (unsigned long long)t;
#endif
}

// Acceptance (DmsaOptimizer.h:307), set numbering over BOTH levels (level 1's sets follow level 0's) and emission, one pass.
// FROM_PLAN = false: every leaf emits at most one unsplit set, decided here.  FROM_PLAN = true: the emission plan
// (out_cnt / sub_*) was written by k_accept + the splitSet kernels.
struct EmitArgs {
    const int* raw_start[RS_MAXSEG];
    const int* raw_diff[RS_MAXSEG];
    const int* out_cnt[RS_MAXSEG];
    const int* sub_start[RS_MAXSEG];
    const int* sub_n[RS_MAXSEG];
    const int* sub_code[RS_MAXSEG];
    const u32_t* idx[RS_MAXSEG];
    const int* keys[RS_MAXSEG];
    int mbase[RS_MAXSEG];
    LevelInfo* infos;
    int level[RS_MAXSEG];
    int nseg, n, tiles, minPts, cap;
    CellStore cs;
    u64_t* status;  // [nseg * tiles] one chain over all segments, zeroed
    int* ticket;    // zeroed
    int* gbase_ready;  // [nseg] zeroed: 1 + gbase of the segment once its first tile knows it
};
template <bool FROM_PLAN>
__global__ void __launch_bounds__(EM_T) k_emit(EmitArgs a) {
    DMSA_PDL_ENTER();
    __shared__ u32_t s_w[EM_T / 32];
    __shared__ u64_t s_red[EM_T / 32 + 2];
    __shared__ int s_t;
    const int tid = threadIdx.x;
    // the number of tiles follows the leaf counts, which live on the device: a small persistent grid draws tickets
    int tl[RS_MAXSEG], total_tiles = 0;
#pragma unroll
    for (int s_ = 0; s_ < RS_MAXSEG; ++s_) {
        tl[s_] = s_ < a.nseg ? max(1, (a.infos[a.level[s_]].R + EM_TILE - 1) / EM_TILE) : 0;
        total_tiles += tl[s_];
    }
    while (true) {
        __syncthreads();
        if (tid == 0) s_t = atomicAdd(a.ticket, 1);
        __syncthreads();
        const int t = s_t;
        if (t >= total_tiles) return;
        int seg = 0, tile = t;
#pragma unroll
        for (int s_ = 0; s_ < RS_MAXSEG - 1; ++s_)
            if (seg == s_ && tile >= tl[s_]) {
                tile -= tl[s_];
                seg = s_ + 1;
            }
        LevelInfo* __restrict__ info = a.infos + a.level[seg];
        const int R = info->R;
        const int c0 = tile * EM_TILE + tid * EM_ITEMS;
        const int* __restrict__ raw_start = a.raw_start[seg];
        const int* __restrict__ raw_diff = a.raw_diff[seg];
        const int* __restrict__ out_cnt = a.out_cnt[seg];
        // phase 1: independent loads
        int rs[EM_ITEMS + 1], aux[EM_ITEMS];
#pragma unroll
        for (int k = 0; k <= EM_ITEMS; ++k) rs[k] = (c0 + k <= R) ? raw_start[c0 + k] : 0;
#pragma unroll
        for (int k = 0; k < EM_ITEMS; ++k) aux[k] = (c0 + k < R) ? (FROM_PLAN ? out_cnt[c0 + k] : raw_diff[c0 + k]) : 0;
        int oc[EM_ITEMS];
        u32_t mine = 0;
#pragma unroll
        for (int k = 0; k < EM_ITEMS; ++k) {
            int v = 0;
            if (c0 + k < R) v = FROM_PLAN ? aux[k] : ((rs[k + 1] - rs[k] >= a.minPts && aux[k]) ? 1 : 0);
            oc[k] = v;
            mine += (u32_t)v;
        }
        // phase 2 (issued before the look-back, consumed after it): first member and voxel key of every emitting leaf
        int kx[EM_ITEMS], ky[EM_ITEMS], kz[EM_ITEMS];
        {
            int pm[EM_ITEMS];
#pragma unroll
            for (int k = 0; k < EM_ITEMS; ++k) pm[k] = oc[k] ? (int)a.idx[seg][rs[k]] : 0;
#pragma unroll
            for (int k = 0; k < EM_ITEMS; ++k) {
                const int* __restrict__ kk = a.keys[seg] + 3 * (size_t)pm[k];
                kx[k] = oc[k] ? kk[0] : 0;
                ky[k] = oc[k] ? kk[1] : 0;
                kz[k] = oc[k] ? kk[2] : 0;
            }
        }
        u32_t total = 0;
        const u32_t bex = block_excl_scan<EM_T>(mine, s_w, &total);
        const int before = (int)block_lookback<sum_combine, EM_T>(a.status, t, (u64_t)total, s_red);  // sets emitted by all earlier tiles (earlier segments included)
        // per-level bookkeeping: the first tile of a segment knows gbase, the last one the level's count (the first tile holds an
        // earlier ticket, so it runs or has finished: waiting for its word cannot deadlock)
        if (tid == 0) {
            if (tile == 0) {
                info->gbase = before;
                st_volatile_u32(reinterpret_cast<u32_t*>(a.gbase_ready) + seg, (u32_t)before + 1u);
            }
            if (tile == tl[seg] - 1) {
                u32_t gb;
                do {
                    gb = ld_volatile_u32(reinterpret_cast<const u32_t*>(a.gbase_ready) + seg);
                } while (gb == 0);
                info->G = before + (int)total - (int)(gb - 1u);
            }
        }
        int g = before + (int)bex;
#pragma unroll
        for (int k = 0; k < EM_ITEMS; ++k) {
            const int c = c0 + k;
            for (int e = 0; e < oc[k]; ++e, ++g) {
                if (g >= a.cap) continue;
                int st, nn, code;
                if (FROM_PLAN) {
                    st = a.sub_start[seg][2 * c + e];
                    nn = a.sub_n[seg][2 * c + e];
                    code = a.sub_code[seg][2 * c + e];
                } else {
                    st = rs[k];
                    nn = rs[k + 1] - rs[k];
                    code = 0;
                }
                a.cs.start[g] = a.mbase[seg] + st;
                a.cs.n[g] = nn;
                a.cs.level[g] = a.level[seg];
                a.cs.sub[g] = code;
                a.cs.key[3 * g] = kx[k];
                a.cs.key[3 * g + 1] = ky[k];
                a.cs.key[3 * g + 2] = kz[k];
            }
        }
    }
}

// member records of every segment in sorted order (kernels_sets.cuh k_gather for both levels in one launch)
struct GatherArgs {
    const u32_t* idx[RS_MAXSEG];
    float4* rec[RS_MAXSEG];
    float4* wrec[RS_MAXSEG];
    const LevelInfo* infos;
    int level[RS_MAXSEG];
};
__global__ void k_gather2(GatherArgs a, const float4* __restrict__ local, const int* __restrict__ tid, int identity_row, const float4* __restrict__ world) {
    DMSA_PDL_ENTER();
    const int seg = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.infos[a.level[seg]].n_valid) return;
    const int p = (int)a.idx[seg][i];
    float4 l = local[p];
    const int t = tid[p];
    l.w = __int_as_float(t < 0 ? identity_row : t);  // static points: the table's identity row reproduces them exactly
    a.rec[seg][i] = l;
    a.wrec[seg][i] = world[p];
}

// plain exclusive sum of n ints (out[i] = in[0] + .. + in[i-1]), single pass
struct ScanArgs {
    const int* in;
    int* out;
    int n, tiles;
    u64_t* status;  // [tiles], zeroed
    int* ticket;    // zeroed
};
__global__ void __launch_bounds__(CS_T) k_scan_excl(ScanArgs a) {
    DMSA_PDL_ENTER();
    __shared__ u32_t s_w[CS_T / 32];
    __shared__ u64_t s_red[CS_T / 32 + 2];
    __shared__ int s_t;
    const int tid = threadIdx.x;
    if (tid == 0) s_t = atomicAdd(a.ticket, 1);
    __syncthreads();
    const int tile = s_t;
    const int i0 = tile * CS_TILE + tid * CS_ITEMS;
    int v[CS_ITEMS];
    u32_t mine = 0;
#pragma unroll
    for (int k = 0; k < CS_ITEMS; ++k) {
        v[k] = (i0 + k < a.n) ? a.in[i0 + k] : 0;
        mine += (u32_t)v[k];
    }
    u32_t total = 0;
    const u32_t bex = block_excl_scan<CS_T>(mine, s_w, &total);
    int run = (int)block_lookback<sum_combine, CS_T>(a.status, tile, (u64_t)total, s_red) + (int)bex;
#pragma unroll
    for (int k = 0; k < CS_ITEMS; ++k) {
        if (i0 + k >= a.n) break;
        a.out[i0 + k] = run;
        run += v[k];
    }
}

}  // namespace dmsa
