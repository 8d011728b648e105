// ccl.cu -- 4-connected labelling + per-droplet statistics on the GPU.
//
// Replaces the device-side work of quantify() (reference quantify_droplets_batch.py:81-95):
//   label(mask, connectivity=1)            -> ccl_local_kernel + ccl_border_kernel (union-find,
//                                             root = smallest raster index of the component)
//   per-label `< min_area` filter (:83-85) -> ccl_flatten_area_kernel + the keep predicate
//   label(lbl) again = compaction (:86)    -> ccl_count / ccl_scan / ccl_assign (prefix sum over
//                                             kept roots in raster order = skimage's numbering)
//   regionprops_table + micron columns     -> ccl_moments_kernel (warp-aggregated 64-bit integer
//                                             atomics) + ccl_finalize_kernel (IEEE f64 divide/sqrt)
//
// Everything is integer until the last kernel, so labels / counts / areas are bit-exact and the
// centroids are the exact integer sums divided once in f64, as numpy does.
#include "common.cuh"

namespace dc {

namespace {

constexpr int TILE = 32;          // ccl_local tile edge (one warp per tile row)
constexpr int SCAN_BLOCK = 1024;  // pixels per compaction block

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

// ---- K1: per 32x32 tile union-find in shared memory; rows are merged with a warp ballot ----
// invert != 0 labels the ZERO pixels instead (the background components the overlay stencil needs).
__global__ void __launch_bounds__(TILE* TILE) ccl_local_kernel(const uint8_t* __restrict__ mask, int* __restrict__ L,
                                                              int H, int W, int invert) {
    __shared__ int s[TILE * TILE];
    __shared__ unsigned rowbits[TILE];
    const int lx = threadIdx.x, ly = threadIdx.y;
    const int x = blockIdx.x * TILE + lx, y = blockIdx.y * TILE + ly;
    const size_t img = (size_t)blockIdx.z * H * W;
    const bool inb = (x < W) && (y < H);
    const bool fg = inb && ((mask[img + (size_t)y * W + x] != 0) != (invert != 0));
    const unsigned bits = __ballot_sync(0xffffffffu, fg);
    const int tid = ly * TILE + lx;
    int lab = -1;
    if (fg) {
        unsigned below = (1u << lx) - 1u;
        unsigned zeros = ~bits & below;                 // background pixels left of me in this row
        int start = zeros ? (32 - __clz(zeros)) : 0;    // first pixel of my horizontal run
        lab = ly * TILE + start;
    }
    s[tid] = lab;
    if (lx == 0) rowbits[ly] = bits;
    __syncthreads();
    if (fg && ly > 0) {
        unsigned up = rowbits[ly - 1];
        if ((up >> lx) & 1u) {
            // one union per overlapping run pair: skip when the pixel to the left already links them
            bool left_links = lx > 0 && ((bits >> (lx - 1)) & 1u) && ((up >> (lx - 1)) & 1u);
            if (!left_links) union_min(s, tid, tid - TILE);
        }
    }
    __syncthreads();
    if (inb) {
        int out = -1;
        if (fg) {
            int r = find_root(s, tid);
            int ry = r / TILE, rx = r % TILE;
            out = (blockIdx.y * TILE + ry) * W + (blockIdx.x * TILE + rx);
        }
        L[img + (size_t)y * W + x] = out;
    }
}

// ---- K2: merge across tile borders in global memory ----
__global__ void ccl_border_kernel(int* L, int H, int W) {
    const int nbr = (H - 1) / TILE;   // horizontal borders: rows y = TILE*k, k = 1..nbr
    const int nbc = (W - 1) / TILE;   // vertical borders:   cols x = TILE*k
    const long long nh = (long long)nbr * W, nv = (long long)nbc * H;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int* Li = L + (size_t)blockIdx.y * H * W;
    if (t < nh) {
        int y = (int)(t / W + 1) * TILE, x = (int)(t % W);
        int i = y * W + x;
        if (Li[i] >= 0 && Li[i - W] >= 0) {
            bool left_links = (x % TILE != 0) && Li[i - 1] >= 0 && Li[i - W - 1] >= 0;
            if (!left_links) union_min(Li, i, i - W);
        }
    } else if (t < nh + nv) {
        t -= nh;
        int x = (int)(t / H + 1) * TILE, y = (int)(t % H);
        int i = y * W + x;
        if (Li[i] >= 0 && Li[i - 1] >= 0) {
            bool up_links = (y % TILE != 0) && Li[i - W] >= 0 && Li[i - W - 1] >= 0;
            if (!up_links) union_min(Li, i, i - 1);
        }
    }
}

// ---- K3 (only when min_area > 1): flatten + per-root pixel count ----
__global__ void ccl_flatten_area_kernel(int* L, int* area, int HW) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int* Li = L + (size_t)blockIdx.y * HW;
    int* Ai = area + (size_t)blockIdx.y * HW;
    int r = -1;
    if (i < HW && Li[i] >= 0) {
        r = find_root(Li, i);
        Li[i] = r;
    }
    unsigned act = __ballot_sync(0xffffffffu, r >= 0);
    if (r >= 0) {
        unsigned peers = __match_any_sync(act, r);
        if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&Ai[r], __popc(peers));
    }
}

__device__ __forceinline__ bool keep_root(const int* Li, const int* Ai, int i, int HW, long long min_area) {
    if (i >= HW || Li[i] != i) return false;
    return min_area <= 1 || (long long)Ai[i] >= min_area;
}

// ---- K4a: kept roots per SCAN_BLOCK pixels ----
__global__ void __launch_bounds__(SCAN_BLOCK) ccl_count_kernel(const int* __restrict__ L, const int* __restrict__ aux,
                                                               int* __restrict__ blockcnt, int HW, int nblk,
                                                               long long min_area) {
    const int* Li = L + (size_t)blockIdx.y * HW;
    const int* Ai = aux + (size_t)blockIdx.y * HW;
    int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    int c = __syncthreads_count(keep_root(Li, Ai, i, HW, min_area));
    if (threadIdx.x == 0) blockcnt[(size_t)blockIdx.y * nblk + blockIdx.x] = c;
}

// ---- K4b: exclusive scan of the block counts of one image (one block per image) ----
__global__ void __launch_bounds__(1024) ccl_scan_kernel(const int* __restrict__ blockcnt, int* __restrict__ blockoff,
                                                        int* __restrict__ counts, int nblk) {
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
    if (threadIdx.x == 0) counts[blockIdx.x] = carry_s;
}

// ---- K4c: consecutive ids (1..n, raster order of first pixel) written at the root positions ----
__global__ void __launch_bounds__(SCAN_BLOCK) ccl_assign_kernel(const int* __restrict__ L, int* __restrict__ aux,
                                                                const int* __restrict__ blockoff, int HW, int nblk,
                                                                long long min_area) {
    __shared__ int warp_cnt[SCAN_BLOCK / 32];
    const int* Li = L + (size_t)blockIdx.y * HW;
    int* Ai = aux + (size_t)blockIdx.y * HW;
    int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    bool is_root = (i < HW) && Li[i] == i;
    bool keep = keep_root(Li, Ai, i, HW, min_area);
    unsigned b = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[wid] = __popc(b);
    __syncthreads();
    if (wid == 0) {
        int w = warp_cnt[lane], wi = w;
        for (int d = 1; d < 32; d <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += n;
        }
        warp_cnt[lane] = wi - w;
    }
    __syncthreads();
    if (is_root) {
        int id = 0;
        if (keep) id = blockoff[(size_t)blockIdx.y * nblk + blockIdx.x] + warp_cnt[wid] + __popc(b & ((1u << lane) - 1u)) + 1;
        Ai[i] = id;   // aux now holds the final label of every root (0 = filtered out)
    }
}

// ---- zero the accumulator rows that will be used ----
__global__ void ccl_zero_rows_kernel(const int* __restrict__ counts, int capacity, long long* area, long long* s0,
                                     long long* s1) {
    int b = blockIdx.y;
    int n = min(counts[b], capacity);
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) {
        size_t o = (size_t)b * capacity + r;
        area[o] = 0; s0[o] = 0; s1[o] = 0;
    }
}

// ---- K5: final label image + area / sum(row) / sum(col) per droplet ----
__global__ void ccl_moments_kernel(const int* __restrict__ L, const int* __restrict__ aux, int* __restrict__ labels_out,
                                   int H, int W, int capacity, unsigned long long* __restrict__ area,
                                   unsigned long long* __restrict__ s0, unsigned long long* __restrict__ s1) {
    const int HW = H * W;
    const int b = blockIdx.y;
    const int* Li = L + (size_t)b * HW;
    const int* Ai = aux + (size_t)b * HW;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int id = 0;
    if (i < HW && Li[i] >= 0) id = Ai[find_root(Li, i)];
    if (labels_out && i < HW) labels_out[(size_t)b * HW + i] = id;
    const bool acc = id > 0 && id <= capacity;
    unsigned act = __ballot_sync(0xffffffffu, acc);
    if (acc) {
        const int lane = threadIdx.x & 31;
        const int py = i / W, px = i - py * W;
        unsigned peers = __match_any_sync(act, id);
        unsigned sr = __reduce_add_sync(peers, (unsigned)py);   // REDUX.SUM over the peer set
        unsigned sc = __reduce_add_sync(peers, (unsigned)px);
        if ((__ffs(peers) - 1) == lane) {
            size_t o = (size_t)b * capacity + (id - 1);
            atomicAdd(&area[o], (unsigned long long)__popc(peers));
            atomicAdd(&s0[o], (unsigned long long)sr);
            atomicAdd(&s1[o], (unsigned long long)sc);
        }
    }
}

// ---- K6: integer sums -> f64 columns (IEEE divide / sqrt, as numpy evaluates them) ----
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

struct Workspace {
    int* L;
    int* aux;
    int* blockcnt;
    int* blockoff;
};

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

// frame-touching background components: flag their roots
__global__ void ovl_mark_kernel(const int* __restrict__ L, uint8_t* __restrict__ flag, int H, int W) {
    const int per = 2 * W + 2 * H;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= per) return;
    int y, x;
    if (t < W) { y = 0; x = t; }
    else if (t < 2 * W) { y = H - 1; x = t - W; }
    else if (t < 2 * W + H) { y = t - 2 * W; x = 0; }
    else { y = t - 2 * W - H; x = W - 1; }
    const size_t img = (size_t)blockIdx.y * H * W;
    const int i = y * W + x;
    if (L[img + i] >= 0) flag[img + find_root(L + img, i)] = 1;
}

// outer[i] = 1 for zero pixels whose component reaches the frame
__global__ void ovl_outer_kernel(const int* __restrict__ L, const uint8_t* __restrict__ flag, uint8_t* __restrict__ outer,
                                 int HW) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const size_t img = (size_t)blockIdx.y * HW;
    uint8_t o = 0;
    if (L[img + i] >= 0) o = flag[img + find_root(L + img, i)];
    outer[img + i] = o;
}

constexpr int OVL_T = 32;         // output tile edge
constexpr int OVL_R = 3;          // halo: contour test at distance <= 2 needs `outer` at distance <= 3
__global__ void __launch_bounds__(OVL_T* OVL_T) ovl_stencil_kernel(const uint8_t* __restrict__ mask,
                                                                  const uint8_t* __restrict__ outer,
                                                                  uint8_t* __restrict__ stencil, int H, int W) {
    constexpr int S = OVL_T + 2 * OVL_R;
    __shared__ uint8_t fg_s[S][S + 2], out_s[S][S + 2], ct_s[S][S + 2];
    const size_t img = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * OVL_T - OVL_R, y0 = blockIdx.y * OVL_T - OVL_R;
    const int tid = threadIdx.y * OVL_T + threadIdx.x;
    for (int i = tid; i < S * S; i += OVL_T * OVL_T) {
        const int ly = i / S, lx = i % S;
        const int y = y0 + ly, x = x0 + lx;
        const bool inb = y >= 0 && y < H && x >= 0 && x < W;
        fg_s[ly][lx] = inb && mask[img + (size_t)y * W + x] != 0;
        out_s[ly][lx] = inb ? outer[img + (size_t)y * W + x] : 1;     // outside the image = the frame
    }
    __syncthreads();
    for (int i = tid; i < S * S; i += OVL_T * OVL_T) {
        const int ly = i / S, lx = i % S;
        uint8_t c = 0;
        if (fg_s[ly][lx] && ly > 0 && ly < S - 1 && lx > 0 && lx < S - 1)
            c = out_s[ly - 1][lx] | out_s[ly + 1][lx] | out_s[ly][lx - 1] | out_s[ly][lx + 1];
        ct_s[ly][lx] = c;
    }
    __syncthreads();
    const int x = blockIdx.x * OVL_T + threadIdx.x, y = blockIdx.y * OVL_T + threadIdx.y;
    if (x >= W || y >= H) return;
    const int ly = threadIdx.y + OVL_R, lx = threadIdx.x + OVL_R;
    uint8_t v = ct_s[ly][lx] | ct_s[ly - 1][lx] | ct_s[ly + 1][lx] | ct_s[ly][lx - 1] | ct_s[ly][lx + 1];
    // diagonal step p -> p + (1, 1) (rows, cols): both non-zero, (p.y, p.x + 1) or (p.y + 1, p.x) outer background;
    // paints p + (-1, +1), p + (+1, -1), p + (0, +2), p + (+2, 0)
    auto step_dr = [&](int py, int px) -> uint8_t {
        return fg_s[py][px] & fg_s[py + 1][px + 1] & (out_s[py][px + 1] | out_s[py + 1][px]);
    };
    // anti-diagonal step p -> p + (1, -1): (p.y, p.x - 1) or (p.y + 1, p.x) outer background;
    // paints p + (+1, +1), p + (-1, -1), p + (+2, 0), p + (0, -2)
    auto step_dl = [&](int py, int px) -> uint8_t {
        return fg_s[py][px] & fg_s[py + 1][px - 1] & (out_s[py][px - 1] | out_s[py + 1][px]);
    };
    v |= step_dr(ly + 1, lx - 1) | step_dr(ly - 1, lx + 1) | step_dr(ly, lx - 2) | step_dr(ly - 2, lx);
    v |= step_dl(ly - 1, lx - 1) | step_dl(ly + 1, lx + 1) | step_dl(ly - 2, lx) | step_dl(ly, lx + 2);
    stencil[img + (size_t)y * W + x] = v ? 1 : 0;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

size_t label_workspace_bytes(int B, int H, int W) {
    size_t hw = (size_t)H * W;
    int nblk = (int)((hw + SCAN_BLOCK - 1) / SCAN_BLOCK);
    return 2 * align256(sizeof(int) * hw * B) + 2 * align256(sizeof(int) * (size_t)nblk * B);
}

int launch_label_stats(const dc_label_args_t* a, cudaStream_t stream) {
    DC_REQUIRE(a && a->mask && a->counts && a->area && a->centroid0 && a->centroid1 && a->eq_diam, DC_EINVAL,
               "dc_label_stats: null pointer argument");
    DC_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->capacity > 0, DC_EINVAL, "dc_label_stats: bad shape %d %d %d cap %d",
               a->B, a->H, a->W, a->capacity);
    DC_REQUIRE((long long)a->H * a->W < (1ll << 31), DC_EINVAL, "dc_label_stats: image too large for int32 indices");
    DC_REQUIRE(a->px_per_um <= 0.0 || (a->area_um2 && a->diam_um), DC_EINVAL,
               "dc_label_stats: px_per_um given but micron columns are NULL");
    const int B = a->B, H = a->H, W = a->W, HW = H * W;
    const int nblk = ceil_div(HW, SCAN_BLOCK);
    DC_REQUIRE(a->workspace && a->workspace_bytes >= label_workspace_bytes(B, H, W), DC_EWORKSPACE,
               "dc_label_stats: workspace too small (%zu < %zu)", a->workspace_bytes, label_workspace_bytes(B, H, W));
    DC_REQUIRE(B <= 65535, DC_EINVAL, "dc_label_stats: batch > 65535");

    char* p = (char*)a->workspace;
    Workspace ws;
    ws.L = (int*)p;        p += align256(sizeof(int) * (size_t)HW * B);
    ws.aux = (int*)p;      p += align256(sizeof(int) * (size_t)HW * B);
    ws.blockcnt = (int*)p; p += align256(sizeof(int) * (size_t)nblk * B);
    ws.blockoff = (int*)p;

    dim3 tb(TILE, TILE);
    dim3 tg(ceil_div(W, TILE), ceil_div(H, TILE), B);
    ccl_local_kernel<<<tg, tb, 0, stream>>>(a->mask, ws.L, H, W, 0);

    long long nborder = (long long)((H - 1) / TILE) * W + (long long)((W - 1) / TILE) * H;
    if (nborder > 0) {
        dim3 bg((unsigned)((nborder + 255) / 256), B);
        ccl_border_kernel<<<bg, 256, 0, stream>>>(ws.L, H, W);
    }
    if (a->min_area > 1) {
        DC_CUDA(cudaMemsetAsync(ws.aux, 0, sizeof(int) * (size_t)HW * B, stream));
        dim3 fg(ceil_div(HW, 256), B);
        ccl_flatten_area_kernel<<<fg, 256, 0, stream>>>(ws.L, ws.aux, HW);
    }
    dim3 sg(nblk, B);
    ccl_count_kernel<<<sg, SCAN_BLOCK, 0, stream>>>(ws.L, ws.aux, ws.blockcnt, HW, nblk, a->min_area);
    ccl_scan_kernel<<<B, 1024, 0, stream>>>(ws.blockcnt, ws.blockoff, a->counts, nblk);
    ccl_assign_kernel<<<sg, SCAN_BLOCK, 0, stream>>>(ws.L, ws.aux, ws.blockoff, HW, nblk, a->min_area);

    // accumulators live in the caller's table: area (i64) and, until finalize, centroid0/1 reused as i64 sums
    long long* s0 = reinterpret_cast<long long*>(a->centroid0);
    long long* s1 = reinterpret_cast<long long*>(a->centroid1);
    int maxrows = a->capacity < (HW + 1) / 2 ? a->capacity : (HW + 1) / 2;
    dim3 zg(ceil_div(maxrows, 256), B);
    ccl_zero_rows_kernel<<<zg, 256, 0, stream>>>(a->counts, a->capacity, (long long*)a->area, s0, s1);
    dim3 mg(ceil_div(HW, 256), B);
    ccl_moments_kernel<<<mg, 256, 0, stream>>>(ws.L, ws.aux, a->labels_out, H, W, a->capacity,
                                               (unsigned long long*)a->area, (unsigned long long*)s0,
                                               (unsigned long long*)s1);
    ccl_finalize_kernel<<<zg, 256, 0, stream>>>(a->counts, a->capacity, (const long long*)a->area, a->centroid0,
                                                a->centroid1, a->eq_diam, a->area_um2, a->diam_um, a->px_per_um);
    DC_CUDA(cudaGetLastError());
    return DC_OK;
}

size_t overlay_workspace_bytes(int B, int H, int W) {
    size_t hw = (size_t)H * W;
    return align256(sizeof(int) * hw * B) + 2 * align256(hw * B);
}

int launch_overlay_stencil(const dc_overlay_args_t* a, cudaStream_t stream) {
    DC_REQUIRE(a && a->mask && a->stencil, DC_EINVAL, "dc_overlay_stencil: null pointer argument");
    DC_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0, DC_EINVAL, "dc_overlay_stencil: bad shape %d %d %d", a->B, a->H, a->W);
    DC_REQUIRE((long long)a->H * a->W < (1ll << 31), DC_EINVAL, "dc_overlay_stencil: image too large for int32 indices");
    DC_REQUIRE(a->B <= 65535, DC_EINVAL, "dc_overlay_stencil: batch > 65535");
    const int B = a->B, H = a->H, W = a->W, HW = H * W;
    DC_REQUIRE(a->workspace && a->workspace_bytes >= overlay_workspace_bytes(B, H, W), DC_EWORKSPACE,
               "dc_overlay_stencil: workspace too small (%zu < %zu)", a->workspace_bytes, overlay_workspace_bytes(B, H, W));
    char* p = (char*)a->workspace;
    int* L = (int*)p;             p += align256(sizeof(int) * (size_t)HW * B);
    uint8_t* flag = (uint8_t*)p;  p += align256((size_t)HW * B);
    uint8_t* outer = (uint8_t*)p;

    dim3 tb(TILE, TILE);
    dim3 tg(ceil_div(W, TILE), ceil_div(H, TILE), B);
    ccl_local_kernel<<<tg, tb, 0, stream>>>(a->mask, L, H, W, 1);          // label the background
    long long nborder = (long long)((H - 1) / TILE) * W + (long long)((W - 1) / TILE) * H;
    if (nborder > 0) {
        dim3 bg((unsigned)((nborder + 255) / 256), B);
        ccl_border_kernel<<<bg, 256, 0, stream>>>(L, H, W);
    }
    DC_CUDA(cudaMemsetAsync(flag, 0, (size_t)HW * B, stream));
    dim3 mg(ceil_div(2 * W + 2 * H, 256), B);
    ovl_mark_kernel<<<mg, 256, 0, stream>>>(L, flag, H, W);
    dim3 og(ceil_div(HW, 256), B);
    ovl_outer_kernel<<<og, 256, 0, stream>>>(L, flag, outer, HW);
    dim3 sb(OVL_T, OVL_T);
    dim3 sg(ceil_div(W, OVL_T), ceil_div(H, OVL_T), B);
    ovl_stencil_kernel<<<sg, sb, 0, stream>>>(a->mask, outer, a->stencil, H, W);
    DC_CUDA(cudaGetLastError());
    return DC_OK;
}

}  // namespace dc
