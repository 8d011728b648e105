// ccl.cu -- 4-connected labelling + per-droplet statistics on the GPU, run-based on a bit-packed mask.
//
// Replaces the device-side work of quantify() (reference quantify_droplets_batch.py:81-95):
//   label(mask, connectivity=1)            -> ccl_tile_kernel + ccl_border_kernel (union-find over RUN STARTS,
//                                             root = smallest raster index of the component)
//   per-label `< min_area` filter (:83-85) -> ccl_area_kernel + the keep predicate of ccl_mark_kernel
//   label(lbl) again = compaction (:86)    -> ccl_mark / ccl_scan / ccl_ids (prefix sum over kept roots in raster
//                                             order = skimage's numbering)
//   regionprops_table + micron columns     -> per-run moments (closed form per run, no per-pixel work) summed per
//                                             tile-local component, ccl_accumulate_kernel (one 64-bit atomic triple
//                                             per tile-local component) + ccl_finalize_kernel (IEEE f64 divide/sqrt)
//
// Data layout.  The u8 mask is read ONCE (1 B/px) and packed to 1 bit/px (`bits`: 32-px words on the droplet path,
// 64-px words for the overlay's background; row pitch WW = ceil(W / word) words).  Everything after that works on
// runs (maximal horizontal stretches of set bits inside one word): a 1024^2 frame with ~3 k droplets has ~30 k runs,
// so the int32 label plane of a per-pixel algorithm
// (4 B/px written, then re-read by every later pass) never exists.  Union-find state lives in planes indexed by the
// pixel index of a run's FIRST pixel and is touched only there (sparse: a few sectors per droplet):
//   P   int32  parent (pixel index of another run start of the same component, smaller or equal)
//   ACC u64    per tile-local root: area | sum(col - tile_x0) << 16 | sum(row - tile_y0) << 36 of its tile-local
//              component (field maxima for the widest, 64 x 128, tile: 8192, 258 048, 520 192: no carries between fields)
//   AUX u32    per final root: total area (min_area > 1), then the final label id (0 = filtered out)
// `rootbits` marks the run starts that are tile-local roots, `keptbits` the final roots that survive min_area.
// The int32 label image is written only when the caller asks for it (labels_out), straight from the runs.
//
// Everything is integer until the last kernel, so labels / counts / areas are bit-exact and the centroids are the
// exact integer sums divided once in f64, as numpy does.
#include "common.cuh"

namespace dc {

namespace {

typedef unsigned long long u64;

constexpr int TILE_H = 128;       // rows per tile = threads per block (one thread per word-row); tile width = one word
constexpr int WORDS_PER_BLOCK = 256;   // compaction granularity of the kept-root scan

__device__ __forceinline__ int find_root(const volatile int* L, int x) {
    int p = L[x];
    while (p != x) { x = p; p = L[x]; }
    return x;
}

// Union by smaller index (so the surviving root is the component's first pixel in raster order).
__device__ __forceinline__ void union_min(int* L, int a, int b) {
    bool done;
    do {
        a = find_root(L, a);
        b = find_root(L, b);
        if (a < b) {
            int old = atomicMin(&L[b], a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            int old = atomicMin(&L[a], b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

// Word type of the bit planes.  The droplet path packs 32 pixels per word (one 32 x 128 tile per 128-thread block:
// 16 KB of shared parent array, 13 blocks per SM, single-instruction bit scans); the overlay's background labelling
// packs 64 (few, long runs: fewer clipped runs and half as many words for the bit stencil).
template <class T> struct WT;
template <> struct WT<u64> {
    static constexpr int BITS = 64;
    static __device__ __forceinline__ int ffs(u64 w) { return __ffsll((long long)w) - 1; }              // lowest set bit
    static __device__ __forceinline__ int top(u64 z) { return 64 - __clzll((long long)z); }            // 1 + highest set bit
};
template <> struct WT<unsigned> {
    static constexpr int BITS = 32;
    static __device__ __forceinline__ int ffs(unsigned w) { return __ffs((int)w) - 1; }
    static __device__ __forceinline__ int top(unsigned z) { return 32 - __clz((int)z); }
};

template <class T> __device__ __forceinline__ T bits_below(int b) { return b >= WT<T>::BITS ? (T)~(T)0 : (T)(((T)1 << b) - (T)1); }       // bits [0, b)
template <class T> __device__ __forceinline__ T bits_upto(int b) { return b >= WT<T>::BITS - 1 ? (T)~(T)0 : (T)(((T)2 << b) - (T)1); }     // bits [0, b]

// first / last bit of the run of ones of `w` that contains bit b (bit b is set)
template <class T> __device__ __forceinline__ int run_start(T w, int b) {
    const T z = (T)~w & bits_below<T>(b);
    return z ? WT<T>::top(z) : 0;
}
template <class T> __device__ __forceinline__ int run_end(T w, int b) {
    const T z = (T)~w & (T)~bits_upto<T>(b);
    return z ? WT<T>::ffs(z) - 1 : WT<T>::BITS - 1;
}

// For every pair (run of `dn`, run of `up`) that overlaps in x: unite(start of the dn run, start of the up run).
// A maximal stretch of dn & up lies inside exactly one run of each word, and two stretches never share both runs,
// so every overlapping pair is visited exactly once.
template <class T, class U>
__device__ __forceinline__ void merge_rows(T dn, T up, U&& unite) {
    T ov = dn & up;
    while (ov) {
        const int b = WT<T>::ffs(ov);
        unite(run_start<T>(dn, b), run_start<T>(up, b));
        ov &= (T)~bits_upto<T>(min(run_end<T>(dn, b), run_end<T>(up, b)));
    }
}

// Shared-memory union-find of the tile kernel.  Ids are raster positions row * BITS + bit (so that the minimum is the
// first pixel), but a plain array would put the run starts of a column -- e.g. bit 0 of every row of a background
// tile -- into one bank: slot(id) rotates each row by its row number.
template <int BITS> __device__ __forceinline__ int slot(int id) {
    return (id & ~(BITS - 1)) | ((id + (id / BITS)) & (BITS - 1));
}
template <int BITS> __device__ __forceinline__ int find_root_s(const volatile int* L, int x) {
    int p = L[slot<BITS>(x)];
    while (p != x) { x = p; p = L[slot<BITS>(x)]; }
    return x;
}
template <int BITS> __device__ __forceinline__ void union_min_s(int* L, int a, int b) {
    bool done;
    do {
        a = find_root_s<BITS>(L, a);
        b = find_root_s<BITS>(L, b);
        if (a < b) {
            int old = atomicMin(&L[slot<BITS>(b)], a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            int old = atomicMin(&L[slot<BITS>(a)], b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

// bit k = (byte k of v != 0)
__device__ __forceinline__ unsigned nz4(unsigned v) {
    const unsigned h = (v | ((v | 0x80808080u) - 0x01010101u)) & 0x80808080u;   // bit 7 of every non-zero byte
    return (((h >> 7) * 0x01020408u) >> 24) & 0xFu;
}

__device__ __forceinline__ u64 pack_acc(unsigned len, unsigned srow, unsigned scol) {
    return (u64)len | ((u64)scol << 16) | ((u64)srow << 36);
}

// ---- K1: one BITS x 128 tile per block: pack the mask to bits, union-find over the tile's runs in shared memory,
//          per-run moments summed into the tile-local roots.  invert != 0 labels the ZERO pixels instead (the
//          background components the overlay stencil needs).
// SINGLE_PASS: every run start zeroes its record, then every run adds its moments to its root atomically (one pass
// over the runs; wins when a tile has few, long runs -- the overlay's background).  Otherwise the roots store their
// own run and only the other runs add (two passes, but half the global writes: wins on droplet masks, measured).
template <class T, bool SINGLE_PASS>
__global__ void __launch_bounds__(TILE_H) ccl_tile_kernel(const uint8_t* __restrict__ mask, T* __restrict__ bits,
                                                          T* __restrict__ rootbits, int* __restrict__ P,
                                                          u64* __restrict__ ACC, unsigned* __restrict__ AUX, int H, int W,
                                                          int WW, int invert, int zero_aux) {
    constexpr int BITS = WT<T>::BITS;
    constexpr int LPR = BITS / 16;                         // lanes (16-pixel loads) per tile row
    __shared__ int par[BITS * TILE_H];
    __shared__ T rowbits[TILE_H];
    const int t = threadIdx.x;
    const int wx = blockIdx.x, ty0 = blockIdx.y * TILE_H, img = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const uint8_t* m = mask + (size_t)img * HW;
    const int x0 = wx * BITS;

    // ---- pack: LPR lanes per row, 16 pixels (one 16-byte load) per lane
    if (((W & 15) == 0) && ((reinterpret_cast<uintptr_t>(mask) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < LPR; ++j) {
            const int c = t + j * TILE_H;
            const int row = c / LPR, q = c % LPR;
            const int y = ty0 + row, x = x0 + q * 16;
            unsigned b16 = 0;
            if (y < H && x < W) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(m + (size_t)y * W + x));
                b16 = nz4(v.x) | (nz4(v.y) << 4) | (nz4(v.z) << 8) | (nz4(v.w) << 12);
                if (invert) b16 ^= 0xFFFFu;
            }
            T part = (T)b16 << (16 * q);
#pragma unroll
            for (int d = 1; d < LPR; d <<= 1) part |= __shfl_xor_sync(0xffffffffu, part, d);
            if (q == 0) rowbits[row] = part;
        }
    } else {
        const int y = ty0 + t;
        T w = 0;
        if (y < H) {
            const int n = min(BITS, W - x0);
            const uint8_t* r = m + (size_t)y * W + x0;
            for (int k = 0; k < n; ++k) w |= (T)((r[k] != 0) != (invert != 0)) << k;
        }
        rowbits[t] = w;
    }
    __syncthreads();

    const int y = ty0 + t;
    const bool inb = y < H;
    const T w = rowbits[t];                                // 0 for rows below the image
    const T up = t > 0 ? rowbits[t - 1] : (T)0;
    const size_t word_idx = ((size_t)img * H + (inb ? y : 0)) * WW + wx;
    if (__syncthreads_or(w != (T)0) == 0) {                // empty tile: nothing to label
        if (inb) { bits[word_idx] = (T)0; rootbits[word_idx] = (T)0; }
        return;
    }
    if constexpr (SINGLE_PASS) {
        if (inb) bits[word_idx] = w;
        int* Pi = P + (size_t)img * HW;
        u64* Ai = ACC + (size_t)img * HW;
        unsigned* Xi = AUX + (size_t)img * HW;
        const int gbase = y * W + x0;                          // pixel index of bit 0 of this word (H*W < 2^31)
        const T starts = w & (T)~(w << 1);
        for (T s = starts; s; s &= s - 1) {
            const int b = WT<T>::ffs(s);
            par[slot<BITS>(t * BITS + b)] = t * BITS + b;
            Ai[gbase + b] = 0ull;                              // any run start may turn out to be its component's root
        }
        __syncthreads();
        merge_rows<T>(w, up, [&](int sd, int su) { union_min_s<BITS>(par, t * BITS + sd, (t - 1) * BITS + su); });
        __syncthreads();                                       // (also: the zeroed records are visible to the whole block)

        // ---- every run start gets its parent in the global plane and adds its moments to its tile-local root
        T rootw = 0;
        for (T s = starts; s; s &= s - 1) {
            const int b = WT<T>::ffs(s);
            const int self = t * BITS + b;
            const int r = find_root_s<BITS>(par, self);
            const int e = run_end<T>(w, b);
            const unsigned len = (unsigned)(e - b + 1);
            const int groot = (ty0 + r / BITS) * W + x0 + (r & (BITS - 1));
            Pi[gbase + b] = groot;
            atomicAdd(&Ai[groot], pack_acc(len, (unsigned)t * len, (unsigned)(b + e) * len / 2u));
            if (r == self) {
                if (zero_aux) Xi[groot] = 0u;
                rootw |= (T)1 << b;
            }
        }
        if (inb) rootbits[word_idx] = rootw;
    } else {
        if (inb) bits[word_idx] = w;
        const T starts = w & (T)~(w << 1);
        for (T s = starts; s; s &= s - 1) {
            const int b = WT<T>::ffs(s);
            par[slot<BITS>(t * BITS + b)] = t * BITS + b;
        }
        __syncthreads();
        merge_rows<T>(w, up, [&](int sd, int su) { union_min_s<BITS>(par, t * BITS + sd, (t - 1) * BITS + su); });
        __syncthreads();

        // ---- tile-local roots publish their own run; every run start gets its parent in the global plane
        int* Pi = P + (size_t)img * HW;
        u64* Ai = ACC + (size_t)img * HW;
        unsigned* Xi = AUX + (size_t)img * HW;
        const int gbase = y * W + x0;                          // pixel index of bit 0 of this word (H*W < 2^31)
        T rootw = 0;
        for (T s = starts; s; s &= s - 1) {
            const int b = WT<T>::ffs(s);
            const int self = t * BITS + b;
            const int r = find_root_s<BITS>(par, self);
            if (r == self) {
                const int e = run_end<T>(w, b);
                const unsigned len = (unsigned)(e - b + 1);
                Ai[gbase + b] = pack_acc(len, (unsigned)t * len, (unsigned)(b + e) * len / 2u);
                Pi[gbase + b] = gbase + b;
                if (zero_aux) Xi[gbase + b] = 0u;
                rootw |= (T)1 << b;
            } else {
                Pi[gbase + b] = (ty0 + r / BITS) * W + x0 + (r & (BITS - 1));
            }
        }
        if (inb) rootbits[word_idx] = rootw;
        __syncthreads();                                       // the roots' records are visible to the whole block
        for (T s = starts & (T)~rootw; s; s &= s - 1) {
            const int b = WT<T>::ffs(s);
            const int r = find_root_s<BITS>(par, t * BITS + b);
            const int e = run_end<T>(w, b);
            const unsigned len = (unsigned)(e - b + 1);
            atomicAdd(&Ai[(ty0 + r / BITS) * W + x0 + (r & (BITS - 1))], pack_acc(len, (unsigned)t * len, (unsigned)(b + e) * len / 2u));
        }
    }
}

// ---- K2: merge across tile borders (runs are clipped at tile borders, so every start has a P entry)
template <class T>
__global__ void ccl_border_kernel(const T* __restrict__ bits, int* P, int H, int W, int WW) {
    constexpr int BITS = WT<T>::BITS;
    const int nbr = (H - 1) / TILE_H;                      // horizontal borders: rows y = TILE_H * k, k = 1..nbr
    const long long nh = (long long)nbr * WW, nv = (long long)H * (WW - 1);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const T* Bi = bits + (size_t)blockIdx.y * H * WW;
    int* Pi = P + (size_t)blockIdx.y * H * W;
    if (i < nh) {
        const int k = (int)(i / WW) + 1, wx = (int)(i % WW);
        const int y = k * TILE_H;
        const T dn = Bi[(size_t)y * WW + wx], up = Bi[(size_t)(y - 1) * WW + wx];
        const int g = y * W + wx * BITS;
        merge_rows<T>(dn, up, [&](int sd, int su) { union_min(Pi, g + sd, g - W + su); });
    } else if (i < nh + nv) {
        i -= nh;
        const int y = (int)(i / (WW - 1)), wx = (int)(i % (WW - 1)) + 1;
        const T a = Bi[(size_t)y * WW + wx - 1], b = Bi[(size_t)y * WW + wx];
        if ((a >> (BITS - 1)) & b & (T)1) {
            const int g = y * W + wx * BITS;
            union_min(Pi, g, g - BITS + run_start<T>(a, BITS - 1));
        }
    }
}

// ---- K2b (only when min_area > 1): total pixel count per final root
template <class T>
__global__ void ccl_area_kernel(const T* __restrict__ rootbits, const int* __restrict__ P, const u64* __restrict__ ACC,
                                unsigned* __restrict__ AUX, int H, int W, int WW) {
    const long long NW = (long long)H * WW;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NW) return;
    const T rw = rootbits[(size_t)blockIdx.y * NW + i];
    if (!rw) return;
    const size_t HW = (size_t)H * W;
    const int* Pi = P + blockIdx.y * HW;
    const int gbase = (int)(i / WW) * W + (int)(i % WW) * WT<T>::BITS;
    for (T s = rw; s; s &= s - 1) {
        const int gi = gbase + WT<T>::ffs(s);
        atomicAdd(&AUX[blockIdx.y * HW + find_root(Pi, gi)], (unsigned)(ACC[blockIdx.y * HW + gi] & 0xFFFFull));
    }
}

// ---- K3: final roots that survive the min_area filter -> keptbits, and their count per WORDS_PER_BLOCK words
template <class T>
__global__ void __launch_bounds__(WORDS_PER_BLOCK) ccl_mark_kernel(const T* __restrict__ rootbits, const int* __restrict__ P,
                                                                   unsigned* __restrict__ AUX, T* __restrict__ keptbits,
                                                                   int* __restrict__ blockcnt, int H, int W, int WW, int nblk,
                                                                   long long min_area) {
    const long long NW = (long long)H * WW;
    const long long i = (long long)blockIdx.x * WORDS_PER_BLOCK + threadIdx.x;
    const size_t HW = (size_t)H * W;
    T kept = 0;
    if (i < NW) {
        const T rw = rootbits[(size_t)blockIdx.y * NW + i];
        if (rw) {
            const int* Pi = P + blockIdx.y * HW;
            unsigned* Xi = AUX + blockIdx.y * HW;
            const int gbase = (int)(i / WW) * W + (int)(i % WW) * WT<T>::BITS;
            for (T s = rw; s; s &= s - 1) {
                const int b = WT<T>::ffs(s);
                const int gi = gbase + b;
                if (Pi[gi] != gi) continue;                           // merged into a component with an earlier first pixel
                if (min_area <= 1 || (long long)Xi[gi] >= min_area) kept |= (T)1 << b;
                else Xi[gi] = 0u;                                     // label 0: filtered out (qdb:83-85)
            }
        }
        keptbits[(size_t)blockIdx.y * NW + i] = kept;
    }
    // block sum of the popcounts
    __shared__ int warp_sum[WORDS_PER_BLOCK / 32];
    int v = __popcll((u64)kept);
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int k = 0; k < WORDS_PER_BLOCK / 32; ++k) tot += warp_sum[k];
        blockcnt[(size_t)blockIdx.y * nblk + blockIdx.x] = tot;
    }
}

// ---- K4: exclusive scan of the block counts of one image (one block per image); zeroes the table rows in use
__global__ void __launch_bounds__(1024) ccl_scan_kernel(const int* __restrict__ blockcnt, int* __restrict__ blockoff,
                                                        int* __restrict__ counts, int nblk, int capacity, long long* area,
                                                        long long* s0, long long* s1) {
    __shared__ int warp_sum[32];
    __shared__ int carry_s;
    const int* c = blockcnt + (size_t)blockIdx.x * nblk;
    int* o = blockoff + (size_t)blockIdx.x * nblk;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < nblk; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nblk ? c[i] : 0;
        int incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += n;
        }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = warp_sum[lane], wi = w;
            for (int d = 1; d < 32; d <<= 1) {
                int n = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += n;
            }
            warp_sum[lane] = wi - w;   // exclusive
        }
        __syncthreads();
        int carry = carry_s;
        if (i < nblk) o[i] = carry + warp_sum[wid] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sum[wid] + incl;
        __syncthreads();
    }
    const int total = carry_s;
    if (threadIdx.x == 0) counts[blockIdx.x] = total;
    const int n = min(total, capacity);
    for (int r = threadIdx.x; r < n; r += 1024) {
        const size_t q = (size_t)blockIdx.x * capacity + r;
        area[q] = 0; s0[q] = 0; s1[q] = 0;
    }
}

// ---- K5: consecutive ids (1..n, raster order of first pixel) written at the kept roots
template <class T>
__global__ void __launch_bounds__(WORDS_PER_BLOCK) ccl_ids_kernel(const T* __restrict__ keptbits, const int* __restrict__ blockoff,
                                                                  unsigned* __restrict__ AUX, int H, int W, int WW, int nblk) {
    __shared__ int warp_cnt[WORDS_PER_BLOCK / 32];
    const long long NW = (long long)H * WW;
    const long long i = (long long)blockIdx.x * WORDS_PER_BLOCK + threadIdx.x;
    const T kw = i < NW ? keptbits[(size_t)blockIdx.y * NW + i] : (T)0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c = __popcll((u64)kw);
    int incl = c;
    for (int d = 1; d < 32; d <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += n;
    }
    if (lane == 31) warp_cnt[wid] = incl;
    __syncthreads();
    int before = 0;
    for (int k = 0; k < wid; ++k) before += warp_cnt[k];
    if (!kw) return;
    unsigned id = (unsigned)(blockoff[(size_t)blockIdx.y * nblk + blockIdx.x] + before + incl - c);
    unsigned* Xi = AUX + (size_t)blockIdx.y * H * W;
    const int gbase = (int)(i / WW) * W + (int)(i % WW) * WT<T>::BITS;
    for (T s = kw; s; s &= s - 1) Xi[gbase + WT<T>::ffs(s)] = ++id;
}

// ---- K6: every tile-local component adds its moments to the table row of its final root
template <class T>
__global__ void ccl_accumulate_kernel(const T* __restrict__ rootbits, const int* __restrict__ P, const u64* __restrict__ ACC,
                                      const unsigned* __restrict__ AUX, int H, int W, int WW, int capacity,
                                      u64* __restrict__ area, u64* __restrict__ s0, u64* __restrict__ s1) {
    const long long NW = (long long)H * WW;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NW) return;
    const T rw = rootbits[(size_t)blockIdx.y * NW + i];
    if (!rw) return;
    const size_t HW = (size_t)H * W;
    const int* Pi = P + blockIdx.y * HW;
    const int y = (int)(i / WW), x0 = (int)(i % WW) * WT<T>::BITS;
    const u64 ty0 = (u64)((y / TILE_H) * TILE_H);
    const int gbase = y * W + x0;
    for (T s = rw; s; s &= s - 1) {
        const int gi = gbase + WT<T>::ffs(s);
        const unsigned id = AUX[blockIdx.y * HW + find_root(Pi, gi)];
        if (id == 0u || id > (unsigned)capacity) continue;
        const u64 pk = ACC[blockIdx.y * HW + gi];
        const u64 a = pk & 0xFFFFull, sc = (pk >> 16) & 0xFFFFFull, sr = pk >> 36;
        const size_t q = (size_t)blockIdx.y * capacity + (id - 1);
        atomicAdd(&area[q], a);
        atomicAdd(&s0[q], sr + a * ty0);
        atomicAdd(&s1[q], sc + a * (u64)x0);
    }
}

// ---- K7: integer sums -> f64 columns (IEEE divide / sqrt, as numpy evaluates them) ----
__global__ void ccl_finalize_kernel(const int* __restrict__ counts, int capacity, const long long* __restrict__ area,
                                    double* c0, double* c1, double* diam, double* area_um2, double* diam_um,
                                    double px_per_um) {
    int b = blockIdx.y;
    int n = min(counts[b], capacity);
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    size_t o = (size_t)b * capacity + r;
    double a = (double)area[o];
    long long sr = reinterpret_cast<const long long*>(c0)[o];
    long long sc = reinterpret_cast<const long long*>(c1)[o];
    const double pi = 3.14159265358979323846;
    double d = sqrt(__ddiv_rn(__dmul_rn(4.0, a), pi));      // regionprops: sqrt(4 * area / pi)
    c0[o] = __ddiv_rn((double)sr, a);                         // mean of integer row coordinates
    c1[o] = __ddiv_rn((double)sc, a);
    diam[o] = d;
    if (px_per_um > 0.0 && area_um2 && diam_um) {
        area_um2[o] = __ddiv_rn(a, __dmul_rn(px_per_um, px_per_um));   // qdb:93
        diam_um[o] = __ddiv_rn(d, px_per_um);                           // qdb:94
    }
}

// ---- K8 (only when the caller wants the label image): 4 pixels per thread, straight from the runs
template <class T>
__global__ void ccl_labels_kernel(const T* __restrict__ bits, const int* __restrict__ P, const unsigned* __restrict__ AUX,
                                  int* __restrict__ labels_out, int H, int W, int WW) {
    constexpr int BITS = WT<T>::BITS;
    const int W4 = (W + 3) >> 2;
    const long long n = (long long)H * W4;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int y = (int)(i / W4), x = (int)(i % W4) * 4;
    const size_t HW = (size_t)H * W;
    const T w = bits[((size_t)blockIdx.y * H + y) * WW + x / BITS];
    const int sh = x & (BITS - 1);
    int out[4] = {0, 0, 0, 0};
    if ((w >> sh) & (T)0xF) {
        const int* Pi = P + blockIdx.y * HW;
        const int gbase = y * W + (x & ~(BITS - 1));
        int prev_start = -1, prev_id = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!((w >> (sh + k)) & (T)1)) continue;
            const int st = run_start<T>(w, sh + k);
            if (st != prev_start) {
                prev_start = st;
                prev_id = (int)AUX[blockIdx.y * HW + find_root(Pi, gbase + st)];
            }
            out[k] = prev_id;
        }
    }
    int* dst = labels_out + blockIdx.y * HW + (size_t)y * W + x;
    if ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(labels_out) & 15) == 0) {
        *reinterpret_cast<int4*>(dst) = make_int4(out[0], out[1], out[2], out[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (x + k < W) dst[k] = out[k];
    }
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct Workspace {
    void *bits, *rootbits, *keptbits;      // bit planes of u32 (droplet path) or u64 (overlay) words
    int* P;
    u64* ACC;
    unsigned* AUX;
    int *blockcnt, *blockoff;
    size_t total;
};

Workspace carve_ws(char* base, int B, int H, int W, int word_bits) {
    Workspace ws;
    const size_t hw = (size_t)H * W * B;
    const size_t nw = (size_t)H * ceil_div(W, word_bits) * B;
    const size_t nblk = (size_t)ceil_div(H * ceil_div(W, word_bits), WORDS_PER_BLOCK) * B;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += align256(bytes);
        return p;
    };
    ws.bits = take(nw * (word_bits / 8));
    ws.rootbits = take(nw * (word_bits / 8));
    ws.keptbits = take(nw * (word_bits / 8));
    ws.P = (int*)take(hw * 4);
    ws.ACC = (u64*)take(hw * 8);
    ws.AUX = (unsigned*)take(hw * 4);
    ws.blockcnt = (int*)take(nblk * 4);
    ws.blockoff = (int*)take(nblk * 4);
    ws.total = off;
    return ws;
}

// ------------------------------------------------------------------------------------------------ overlay stencil
// The pixels that cv2.drawContours(img, findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE), -1, color, 2)
// paints (reference quantify_droplets_batch.py:74-79), derived without tracing (the rule is pinned against cv2
// itself on adversarial and random masks by the CPU and GPU overlay tests under tests/):
//   outer background = zero pixels 4-connected to the image frame (holes are not);
//   contour          = non-zero pixels with a 4-neighbour in the outer background or outside the image
//                      (exactly the point set of the external contours: only top-level borders are retrieved);
//   thickness 2      = every contour pixel plus its 4-neighbours (cv2's radius-1 caps and 3-wide bands), plus, for
//                      every diagonal step of a border -- two diagonal non-zero pixels whose common 4-neighbour on
//                      one side is outer background -- the pixels one step to either side of both ends, which the
//                      outline of cv2's rotated band quad rounds onto.
// Everything runs on the bit planes: the BACKGROUND is labelled with the run-based kernels above (a handful of long
// runs per row, so the one huge outer component costs little), frame-touching roots are flagged, and the rule is
// evaluated 64 pixels at a time with shifts and ANDs.

// frame-touching background components: flag their roots (AUX = 1; the tile kernel zeroed it at every tile-local root)
__global__ void ovl_mark_kernel(const u64* __restrict__ bits, const int* __restrict__ P, unsigned* __restrict__ AUX, int H,
                                int W, int WW) {
    const int per = 2 * W + 2 * H;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= per) return;
    int y, x;
    if (t < W) { y = 0; x = t; }
    else if (t < 2 * W) { y = H - 1; x = t - W; }
    else if (t < 2 * W + H) { y = t - 2 * W; x = 0; }
    else { y = t - 2 * W - H; x = W - 1; }
    const size_t HW = (size_t)H * W;
    const u64 w = bits[((size_t)blockIdx.y * H + y) * WW + (x >> 6)];
    if (!((w >> (x & 63)) & 1ull)) return;
    const int g = y * W + (x & ~63) + run_start<u64>(w, x & 63);
    AUX[blockIdx.y * HW + find_root(P + blockIdx.y * HW, g)] = 1u;
}

// outer word = the background runs whose component reaches the frame
__global__ void ovl_outer_kernel(const u64* __restrict__ bits, const int* __restrict__ P, const unsigned* __restrict__ AUX,
                                 u64* __restrict__ outer, int H, int W, int WW) {
    const long long NW = (long long)H * WW;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NW) return;
    const size_t HW = (size_t)H * W;
    const u64 w = bits[(size_t)blockIdx.y * NW + i];
    u64 o = 0;
    if (w) {
        const int* Pi = P + blockIdx.y * HW;
        const int gbase = (int)(i / WW) * W + (int)(i % WW) * 64;
        for (u64 s = w & ~(w << 1); s; s &= s - 1) {
            const int b = __ffsll((long long)s) - 1;
            if (AUX[blockIdx.y * HW + find_root(Pi, gbase + b)]) o |= bits_upto<u64>(run_end<u64>(w, b)) & ~bits_below<u64>(b);
        }
    }
    outer[(size_t)blockIdx.y * NW + i] = o;
}

// Row accessor for the stencil: word wx of row y of a bit plane, `fill` outside the image; `pad` = value of the
// bits of the last word beyond column W.
struct PlaneRow {
    const u64* row;       // nullptr: outside the image
    int WW;
    u64 fill, tailmask;   // tailmask: valid bits of the last word
    __device__ __forceinline__ u64 word(int wx) const {
        if (!row || wx < 0 || wx >= WW) return fill;
        u64 v = row[wx];
        if (wx == WW - 1) v = (v & tailmask) | (fill & ~tailmask);
        return v;
    }
    // bit x of the result = plane[x + dx] for the 64 columns of word wx (|dx| <= 3)
    __device__ __forceinline__ u64 shifted(int wx, int dx) const {
        const u64 c = word(wx);
        if (dx == 0) return c;
        if (dx > 0) return (c >> dx) | (word(wx + 1) << (64 - dx));
        return (c << -dx) | (word(wx - 1) >> (64 + dx));
    }
};

__device__ __forceinline__ u64 spread8(unsigned b) {       // 8 bits -> 8 bytes of 0 / 1
    u64 v = ((u64)b * 0x0101010101010101ull) & 0x8040201008040201ull;
    return ((v + 0x7F7F7F7F7F7F7F7Full) >> 7) & 0x0101010101010101ull;
}

__global__ void ovl_stencil_kernel(const u64* __restrict__ bg, const u64* __restrict__ outer, uint8_t* __restrict__ stencil,
                                   int H, int W, int WW) {
    const long long NW = (long long)H * WW;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NW) return;
    const int y = (int)(i / WW), wx = (int)(i % WW);
    const u64* bgi = bg + (size_t)blockIdx.y * NW;
    const u64* oui = outer + (size_t)blockIdx.y * NW;
    const u64 tailmask = (W & 63) ? bits_below<u64>(W & 63) : ~0ull;
    // FG(r): foreground = not background, 0 outside the image.  OUT(r): outer background, 1 outside the image.
    auto BG = [&](int r) { return PlaneRow{(r >= 0 && r < H) ? bgi + (size_t)r * WW : nullptr, WW, ~0ull, tailmask}; };
    auto OUT = [&](int r) { return PlaneRow{(r >= 0 && r < H) ? oui + (size_t)r * WW : nullptr, WW, ~0ull, tailmask}; };
    auto fg = [&](int r, int dx) { return ~BG(r).shifted(wx, dx); };          // background is 1 outside the image
    auto out = [&](int r, int dx) { return OUT(r).shifted(wx, dx); };
    // contour pixels of row r, shifted by dx
    auto ct = [&](int r, int dx) {
        return fg(r, dx) & (out(r - 1, dx) | out(r + 1, dx) | out(r, dx - 1) | out(r, dx + 1));
    };
    u64 v = ct(y, 0) | ct(y - 1, 0) | ct(y + 1, 0) | ct(y, -1) | ct(y, 1);
    // diagonal step p -> p + (1, 1) (rows, cols): both non-zero, (p.y, p.x + 1) or (p.y + 1, p.x) outer background;
    // paints p + (-1, +1), p + (+1, -1), p + (0, +2), p + (+2, 0)
    auto step_dr = [&](int r, int dx) { return fg(r, dx) & fg(r + 1, dx + 1) & (out(r, dx + 1) | out(r + 1, dx)); };
    // anti-diagonal step p -> p + (1, -1): (p.y, p.x - 1) or (p.y + 1, p.x) outer background;
    // paints p + (+1, +1), p + (-1, -1), p + (+2, 0), p + (0, -2)
    auto step_dl = [&](int r, int dx) { return fg(r, dx) & fg(r + 1, dx - 1) & (out(r, dx - 1) | out(r + 1, dx)); };
    v |= step_dr(y + 1, -1) | step_dr(y - 1, 1) | step_dr(y, -2) | step_dr(y - 2, 0);
    v |= step_dl(y - 1, -1) | step_dl(y + 1, 1) | step_dl(y - 2, 0) | step_dl(y, 2);
    v &= (wx == WW - 1) ? tailmask : ~0ull;
    uint8_t* dst = stencil + (size_t)blockIdx.y * H * W + (size_t)y * W + (size_t)wx * 64;
    const int n = min(64, W - wx * 64);
    if (n == 64 && (W & 15) == 0 && (reinterpret_cast<uintptr_t>(stencil) & 15) == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned h = (unsigned)(v >> (16 * k)) & 0xFFFFu;
            const u64 lo = spread8(h & 0xFFu), hi = spread8(h >> 8);
            *reinterpret_cast<uint4*>(dst + 16 * k) =
                make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)hi, (unsigned)(hi >> 32));
        }
    } else {
        for (int k = 0; k < n; ++k) dst[k] = (uint8_t)((v >> k) & 1ull);
    }
}

}  // namespace

constexpr int LABEL_WORD_BITS = 32;       // droplet labelling: 32-pixel words (one 32 x 128 tile per block)
typedef unsigned LabelWord;
constexpr int OVERLAY_WORD_BITS = 64;     // background labelling of the overlay stencil: 64-pixel words

size_t label_workspace_bytes(int B, int H, int W) { return carve_ws(nullptr, B, H, W, LABEL_WORD_BITS).total; }

int launch_label_stats(const dc_label_args_t* a, cudaStream_t stream) {
    DC_REQUIRE(a && a->mask && a->counts && a->area && a->centroid0 && a->centroid1 && a->eq_diam, DC_EINVAL,
               "dc_label_stats: null pointer argument");
    DC_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->capacity > 0, DC_EINVAL, "dc_label_stats: bad shape %d %d %d cap %d",
               a->B, a->H, a->W, a->capacity);
    DC_REQUIRE((long long)a->H * a->W < (1ll << 31), DC_EINVAL, "dc_label_stats: image too large for int32 indices");
    DC_REQUIRE(a->px_per_um <= 0.0 || (a->area_um2 && a->diam_um), DC_EINVAL,
               "dc_label_stats: px_per_um given but micron columns are NULL");
    const int B = a->B, H = a->H, W = a->W;
    DC_REQUIRE(a->workspace && a->workspace_bytes >= label_workspace_bytes(B, H, W), DC_EWORKSPACE,
               "dc_label_stats: workspace too small (%zu < %zu)", a->workspace_bytes, label_workspace_bytes(B, H, W));
    DC_REQUIRE(((uintptr_t)a->workspace & 7) == 0, DC_EINVAL, "dc_label_stats: workspace must be 8-byte aligned");
    DC_REQUIRE(B <= 65535, DC_EINVAL, "dc_label_stats: batch > 65535");
    typedef LabelWord T;
    const int WW = ceil_div(W, LABEL_WORD_BITS);
    const long long NW = (long long)H * WW;
    const int nblk = (int)((NW + WORDS_PER_BLOCK - 1) / WORDS_PER_BLOCK);
    DC_REQUIRE(ceil_div(H, TILE_H) <= 65535, DC_EINVAL, "dc_label_stats: image too tall");
    const Workspace ws = carve_ws((char*)a->workspace, B, H, W, LABEL_WORD_BITS);
    T* bits = (T*)ws.bits;
    T* rootbits = (T*)ws.rootbits;
    T* keptbits = (T*)ws.keptbits;
    const int filter = a->min_area > 1;

    ccl_tile_kernel<T, false><<<dim3(WW, ceil_div(H, TILE_H), B), TILE_H, 0, stream>>>(a->mask, bits, rootbits, ws.P, ws.ACC, ws.AUX, H,
                                                                               W, WW, 0, filter);
    const long long nborder = (long long)((H - 1) / TILE_H) * WW + (long long)H * (WW - 1);
    if (nborder > 0)
        ccl_border_kernel<T><<<dim3((unsigned)((nborder + 255) / 256), B), 256, 0, stream>>>(bits, ws.P, H, W, WW);
    const dim3 wg((unsigned)((NW + 255) / 256), B);
    if (filter) ccl_area_kernel<T><<<wg, 256, 0, stream>>>(rootbits, ws.P, ws.ACC, ws.AUX, H, W, WW);
    ccl_mark_kernel<T><<<dim3(nblk, B), WORDS_PER_BLOCK, 0, stream>>>(rootbits, ws.P, ws.AUX, keptbits, ws.blockcnt, H, W, WW, nblk,
                                                                     a->min_area);
    // accumulators live in the caller's table: area (i64) and, until finalize, centroid0/1 reused as i64 sums
    long long* s0 = reinterpret_cast<long long*>(a->centroid0);
    long long* s1 = reinterpret_cast<long long*>(a->centroid1);
    ccl_scan_kernel<<<B, 1024, 0, stream>>>(ws.blockcnt, ws.blockoff, a->counts, nblk, a->capacity, (long long*)a->area, s0, s1);
    ccl_ids_kernel<T><<<dim3(nblk, B), WORDS_PER_BLOCK, 0, stream>>>(keptbits, ws.blockoff, ws.AUX, H, W, WW, nblk);
    ccl_accumulate_kernel<T><<<wg, 256, 0, stream>>>(rootbits, ws.P, ws.ACC, ws.AUX, H, W, WW, a->capacity, (u64*)a->area,
                                                     (u64*)s0, (u64*)s1);
    const long long hw = (long long)H * W;
    const int maxrows = (int)(a->capacity < (hw + 1) / 2 ? a->capacity : (hw + 1) / 2);
    ccl_finalize_kernel<<<dim3(ceil_div(maxrows, 256), B), 256, 0, stream>>>(a->counts, a->capacity, (const long long*)a->area,
                                                                            a->centroid0, a->centroid1, a->eq_diam,
                                                                            a->area_um2, a->diam_um, a->px_per_um);
    if (a->labels_out) {
        const long long n4 = (long long)H * ((W + 3) >> 2);
        ccl_labels_kernel<T><<<dim3((unsigned)((n4 + 255) / 256), B), 256, 0, stream>>>(bits, ws.P, ws.AUX, a->labels_out, H, W, WW);
    }
    DC_CUDA(cudaGetLastError());
    return DC_OK;
}

size_t overlay_workspace_bytes(int B, int H, int W) { return carve_ws(nullptr, B, H, W, OVERLAY_WORD_BITS).total; }

int launch_overlay_stencil(const dc_overlay_args_t* a, cudaStream_t stream) {
    DC_REQUIRE(a && a->mask && a->stencil, DC_EINVAL, "dc_overlay_stencil: null pointer argument");
    DC_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0, DC_EINVAL, "dc_overlay_stencil: bad shape %d %d %d", a->B, a->H, a->W);
    DC_REQUIRE((long long)a->H * a->W < (1ll << 31), DC_EINVAL, "dc_overlay_stencil: image too large for int32 indices");
    DC_REQUIRE(a->B <= 65535, DC_EINVAL, "dc_overlay_stencil: batch > 65535");
    const int B = a->B, H = a->H, W = a->W;
    DC_REQUIRE(a->workspace && a->workspace_bytes >= overlay_workspace_bytes(B, H, W), DC_EWORKSPACE,
               "dc_overlay_stencil: workspace too small (%zu < %zu)", a->workspace_bytes, overlay_workspace_bytes(B, H, W));
    DC_REQUIRE(((uintptr_t)a->workspace & 7) == 0, DC_EINVAL, "dc_overlay_stencil: workspace must be 8-byte aligned");
    DC_REQUIRE(ceil_div(H, TILE_H) <= 65535, DC_EINVAL, "dc_overlay_stencil: image too tall");
    const int WW = ceil_div(W, OVERLAY_WORD_BITS);
    const long long NW = (long long)H * WW;
    const Workspace ws = carve_ws((char*)a->workspace, B, H, W, OVERLAY_WORD_BITS);
    u64* bits = (u64*)ws.bits;
    u64* rootbits = (u64*)ws.rootbits;
    u64* outer = (u64*)ws.keptbits;

    // label the background (runs of zero pixels); AUX is zeroed at every tile-local root
    ccl_tile_kernel<u64, true><<<dim3(WW, ceil_div(H, TILE_H), B), TILE_H, 0, stream>>>(a->mask, bits, rootbits, ws.P, ws.ACC, ws.AUX, H,
                                                                                 W, WW, 1, 1);
    const long long nborder = (long long)((H - 1) / TILE_H) * WW + (long long)H * (WW - 1);
    if (nborder > 0)
        ccl_border_kernel<u64><<<dim3((unsigned)((nborder + 255) / 256), B), 256, 0, stream>>>(bits, ws.P, H, W, WW);
    ovl_mark_kernel<<<dim3(ceil_div(2 * W + 2 * H, 256), B), 256, 0, stream>>>(bits, ws.P, ws.AUX, H, W, WW);
    const dim3 wg((unsigned)((NW + 255) / 256), B);
    ovl_outer_kernel<<<wg, 256, 0, stream>>>(bits, ws.P, ws.AUX, outer, H, W, WW);
    ovl_stencil_kernel<<<wg, 256, 0, stream>>>(bits, outer, a->stencil, H, W, WW);
    DC_CUDA(cudaGetLastError());
    return DC_OK;
}

}  // namespace dc
