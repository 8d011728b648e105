// density.cu -- the per-image maps of the reference's alternate front end quantify_pipline.py:
//
//   dc_roi_mask          generate_roi_mask            quantify_pipline.py:44-51   + the cv2.moments centroid :133-136
//   dc_radial_density    get_targets                  quantify_pipline.py:61-91   (droplet centroids: dc_label_stats)
//   dc_spatial_density   density_maps                 quantify_pipline.py:93-97
//
// Everything here is integer or correctly-rounded IEEE arithmetic issued in the order the reference's libraries use,
// so the results are bit-exact against OpenCV / numpy / scipy (tests/test_gpu_density.py):
//   cvtColor RGB2GRAY      (9798 R + 19235 G + 3735 B + 2^14) >> 15
//   GaussianBlur 15x15, 0  OpenCV's fixed-point path for u8: the error-diffused 8-bit kernel of sigma 2.6
//                          {1,3,6,12,20,30,36,40,...} (sum 256), rows then columns, BORDER_REFLECT_101, one rounding
//   THRESH_OTSU            getThreshVal_Otsu_8u: 256-bin histogram, the f64 recurrence, one thread per image
//   MORPH_CLOSE / OPEN     15x15 rectangle, out-of-image taps ignored; on bit-packed rows (32 pixels per word): a dilate is
//                          an OR of funnel-shifted words, an erode the complement of a dilate of the complement
//   moments                m00, m10, m01 as exact integer sums, cx = int(m10 / m00) in f64
//   get_targets            d = sqrt((double)((x-cx)^2 + (y-cy)^2)); np.linspace bounds i * (max_d / n), last = max_d;
//                          ring i: bounds[i] < d <= bounds[i+1]
//   gaussian_filter        scipy NI_Correlate1D, symmetric kernel: f64 accumulation, centre first, then tap pairs from
//                          the outside in; axis 0 then axis 1, each rounded to f32; mode 'reflect' (dcba|abcd|dcba)
//   density                g(mask) / (g(roi) + 1e-5f) * 100f in f32
// These are HBM-/latency-bound stencils over a handful of byte planes; none of them is on the BASELINE metric's path.
#include "common.cuh"

namespace dc {

namespace {

__constant__ int kGauss15[15] = {1, 3, 6, 12, 20, 30, 36, 40, 36, 30, 20, 12, 6, 3, 1};

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    const int p = 2 * (n - 1);
    i %= p;
    if (i < 0) i += p;
    return i >= n ? p - i : i;
}

// (d c b a | a b c d | d c b a)
__device__ __forceinline__ int reflect_half(int i, int n) {
    const int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i >= n ? p - 1 - i : i;
}

// ---- ROI: gray + 15x15 fixed-point Gaussian + histogram, one 64 x 32 tile (+7 halo) per block in shared memory
constexpr int BT_W = 64, BT_H = 32, BT_R = 7;
__global__ void __launch_bounds__(256) roi_blur_hist_kernel(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ blurred,
                                                            unsigned int* __restrict__ hist, int H, int W) {
    __shared__ uint8_t gray_s[BT_H + 2 * BT_R][BT_W + 2 * BT_R + 2];
    __shared__ uint16_t hb_s[BT_H + 2 * BT_R][BT_W];
    __shared__ unsigned int sh[256];
    const int tid = threadIdx.x;
    sh[tid] = 0;
    const size_t img = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * BT_W, y0 = blockIdx.y * BT_H;
    for (int i = tid; i < (BT_H + 2 * BT_R) * (BT_W + 2 * BT_R); i += 256) {
        const int ly = i / (BT_W + 2 * BT_R), lx = i % (BT_W + 2 * BT_R);
        const int y = reflect101(y0 + ly - BT_R, H), x = reflect101(x0 + lx - BT_R, W);    // BORDER_REFLECT_101
        const uint8_t* px = rgb + (img + (size_t)y * W + x) * 3;
        gray_s[ly][lx] = (uint8_t)((px[0] * 9798 + px[1] * 19235 + px[2] * 3735 + (1 << 14)) >> 15);
    }
    __syncthreads();
    for (int i = tid; i < (BT_H + 2 * BT_R) * BT_W; i += 256) {
        const int ly = i / BT_W, lx = i % BT_W;
        int acc = 0;
#pragma unroll
        for (int j = 0; j < 15; ++j) acc += gray_s[ly][lx + j] * kGauss15[j];
        hb_s[ly][lx] = (uint16_t)acc;                     // <= 255 * 256
    }
    __syncthreads();
    for (int i = tid; i < BT_H * BT_W; i += 256) {
        const int ly = i / BT_W, lx = i % BT_W;
        const int y = y0 + ly, x = x0 + lx;
        if (y >= H || x >= W) continue;
        int acc = 0;
#pragma unroll
        for (int j = 0; j < 15; ++j) acc += hb_s[ly + j][lx] * kGauss15[j];
        const int v = (acc + (1 << 15)) >> 16;
        blurred[img + (size_t)y * W + x] = (uint8_t)v;
        atomicAdd(&sh[v], 1u);
    }
    __syncthreads();
    if (sh[tid]) atomicAdd(&hist[(size_t)blockIdx.z * 256 + tid], sh[tid]);
}

// ---- ROI: Otsu threshold, OpenCV's recurrence, one thread per image (strict f64, no contraction)
__global__ void roi_otsu_kernel(const unsigned int* __restrict__ hist, int* __restrict__ thresh, int B, int HW) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const unsigned int* h = hist + (size_t)b * 256;
    const double scale = __ddiv_rn(1.0, (double)HW);
    double mu = 0.0;
    for (int i = 0; i < 256; ++i) mu = __dadd_rn(mu, __dmul_rn((double)i, (double)h[i]));
    mu = __dmul_rn(mu, scale);
    double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
    int max_val = 0;
    const double eps = 1.1920928955078125e-07;     // FLT_EPSILON
    for (int i = 0; i < 256; ++i) {
        const double p_i = __dmul_rn((double)h[i], scale);
        mu1 = __dmul_rn(mu1, q1);
        q1 = __dadd_rn(q1, p_i);
        const double q2 = __dsub_rn(1.0, q1);
        if (fmin(q1, q2) < eps || fmax(q1, q2) > __dsub_rn(1.0, eps)) continue;
        mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p_i)), q1);
        const double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
        const double dm = __dsub_rn(mu1, mu2);
        const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), dm), dm);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    thresh[b] = max_val;
}

// ---- ROI: binarise (> Otsu threshold) and pack 32 pixels per word: bit i of word wx = pixel 32 wx + i.
// The morphology runs on these words: a rectangular dilate is an OR of shifted words, and an erode that ignores
// out-of-image taps is the complement of a dilate of the complement (both with zeros outside the image).
__global__ void roi_pack_kernel(const uint8_t* __restrict__ blurred, const int* __restrict__ thresh,
                                uint32_t* __restrict__ bits, int H, int W, int Wp) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;     // blockDim.x multiple of 32
    const size_t img = (size_t)blockIdx.z * H * W;
    const bool v = x < W && blurred[img + (size_t)y * W + x] > thresh[blockIdx.z];
    const unsigned w = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && (x >> 5) < Wp) bits[((size_t)blockIdx.z * H + y) * Wp + (x >> 5)] = w;
}

// Horizontal pass of a (2R+1)-wide rectangular dilate (ERODE: erode) on packed rows.
template <bool ERODE>
__global__ void roi_morph_h_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int H, int W, int Wp, int R) {
    const int wx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (wx >= Wp) return;
    const uint32_t* row = in + ((size_t)blockIdx.z * H + y) * Wp;
    const uint32_t valid = (wx == Wp - 1 && (W & 31)) ? ((1u << (W & 31)) - 1u) : 0xffffffffu;
    auto ld = [&](int k) -> uint32_t {                      // complemented for the erode; zero outside the image
        if (k < 0 || k >= Wp) return 0u;
        uint32_t v = row[k];
        if (ERODE) v = ~v;
        return (k == Wp - 1 && (W & 31)) ? (v & ((1u << (W & 31)) - 1u)) : v;
    };
    const uint32_t l = ld(wx - 1), c = ld(wx), r = ld(wx + 1);
    uint32_t acc = c;
    for (int s = 1; s <= R; ++s) acc |= __funnelshift_l(l, c, s) | __funnelshift_r(c, r, s);      // R < 32
    if (ERODE) acc = ~acc;
    out[((size_t)blockIdx.z * H + y) * Wp + wx] = acc & valid;
}

// Vertical pass: OR (ERODE: AND, rows outside the image ignored) of rows y-R .. y+R.
template <bool ERODE>
__global__ void roi_morph_v_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int H, int Wp, int R) {
    const int wx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (wx >= Wp) return;
    const uint32_t* col = in + (size_t)blockIdx.z * H * Wp + wx;
    uint32_t acc = ERODE ? 0xffffffffu : 0u;
    const int ya = max(0, y - R), yb = min(H - 1, y + R);
    for (int yy = ya; yy <= yb; ++yy) {
        const uint32_t v = col[(size_t)yy * Wp];
        acc = ERODE ? (acc & v) : (acc | v);
    }
    out[((size_t)blockIdx.z * H + y) * Wp + wx] = acc;
}

// ---- ROI: unpack to {0,1} bytes + moment sums (m00, m10 = sum x, m01 = sum y) as exact integers
__global__ void roi_finish_kernel(const uint32_t* __restrict__ bits, uint8_t* __restrict__ roi,
                                  unsigned long long* __restrict__ sums, int H, int W, int Wp) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;      // blockDim.x multiple of 32
    const size_t img = (size_t)blockIdx.z * H * W;
    uint32_t w = 0;
    if ((x >> 5) < Wp) w = bits[((size_t)blockIdx.z * H + y) * Wp + (x >> 5)];
    if (x < W) roi[img + (size_t)y * W + x] = (uint8_t)((w >> (x & 31)) & 1u);
    if ((threadIdx.x & 31) == 0 && w) {
        const unsigned n = __popc(w);
        // sum of the set bit positions: sum_k 2^k * popc(w & mask_k)
        const unsigned si = __popc(w & 0xaaaaaaaau) + 2 * __popc(w & 0xccccccccu) + 4 * __popc(w & 0xf0f0f0f0u) +
                            8 * __popc(w & 0xff00ff00u) + 16 * __popc(w & 0xffff0000u);
        atomicAdd(&sums[(size_t)blockIdx.z * 3 + 0], (unsigned long long)n);
        atomicAdd(&sums[(size_t)blockIdx.z * 3 + 1], (unsigned long long)n * (unsigned)(x & ~31) + si);
        atomicAdd(&sums[(size_t)blockIdx.z * 3 + 2], (unsigned long long)n * (unsigned)y);
    }
}

__global__ void roi_centroid_kernel(const unsigned long long* __restrict__ sums, int* __restrict__ centroid, int B, int H, int W) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const unsigned long long m00 = sums[(size_t)b * 3], m10 = sums[(size_t)b * 3 + 1], m01 = sums[(size_t)b * 3 + 2];
    int cy = H / 2, cx = W / 2;                                    // quantify_pipline.py:135-136
    if (m00) {
        cx = (int)__ddiv_rn((double)m10, (double)m00);
        cy = (int)__ddiv_rn((double)m01, (double)m00);
    }
    centroid[2 * b] = cy;
    centroid[2 * b + 1] = cx;
}

// ---- radial: max squared distance of the ROI pixels to the centroid
__global__ void radial_maxd2_kernel(const uint8_t* __restrict__ roi, const int* __restrict__ centroid,
                                    unsigned long long* __restrict__ maxd2, int H, int W) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int HW = H * W;
    const size_t img = (size_t)blockIdx.y * HW;
    const long long cy = centroid[2 * blockIdx.y], cx = centroid[2 * blockIdx.y + 1];
    unsigned long long d2 = 0;
    bool any = false;
    if (i < HW && roi[img + i]) {
        const long long dx = (i % W) - cx, dy = (i / W) - cy;
        d2 = (unsigned long long)(dx * dx + dy * dy);
        any = true;
    }
    for (int o = 16; o > 0; o >>= 1) {
        d2 = max(d2, __shfl_down_sync(0xffffffffu, d2, o));
    }
    const unsigned anyw = __ballot_sync(0xffffffffu, any);
    // slot 0: max d^2; slot 1: number of ROI pixels seen (only "non-zero" matters)
    if ((threadIdx.x & 31) == 0 && anyw) {
        atomicMax(&maxd2[(size_t)blockIdx.y * 2], d2);
        atomicAdd(&maxd2[(size_t)blockIdx.y * 2 + 1], (unsigned long long)__popc(anyw));
    }
}

constexpr int RADIAL_MAX_LAYERS = 64;

__device__ __forceinline__ int ring_of(double d, const double* bounds, int n) {
    for (int i = 0; i < n; ++i)
        if (bounds[i] < d && d <= bounds[i + 1]) return i;
    return -1;
}

// ---- radial: ring bounds (np.linspace) + droplets per ring; one block per image
__global__ void radial_rings_kernel(const unsigned long long* __restrict__ maxd2, const int* __restrict__ centroid,
                                    const int* __restrict__ counts, const double* __restrict__ c0,
                                    const double* __restrict__ c1, int capacity, int nb, double* __restrict__ bounds_out,
                                    int* __restrict__ ring_counts) {
    __shared__ double bounds[RADIAL_MAX_LAYERS + 1];
    __shared__ int cnt[RADIAL_MAX_LAYERS];
    const int b = blockIdx.x;
    if (threadIdx.x <= nb) {
        const double max_d = sqrt((double)maxd2[(size_t)b * 2]);
        const double step = __ddiv_rn(max_d, (double)nb);
        bounds[threadIdx.x] = threadIdx.x == nb ? max_d : __dmul_rn((double)threadIdx.x, step);
    }
    if (threadIdx.x < nb) cnt[threadIdx.x] = 0;
    __syncthreads();
    const double cy = (double)centroid[2 * b], cx = (double)centroid[2 * b + 1];
    int n = counts[b];
    if (n > capacity) n = capacity;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const double dx = __dsub_rn(c1[(size_t)b * capacity + r], cx), dy = __dsub_rn(c0[(size_t)b * capacity + r], cy);
        const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        const int k = ring_of(d, bounds, nb);
        if (k >= 0) atomicAdd(&cnt[k], 1);
    }
    __syncthreads();
    if (threadIdx.x <= nb) bounds_out[(size_t)b * (RADIAL_MAX_LAYERS + 1) + threadIdx.x] = bounds[threadIdx.x];
    if (threadIdx.x < nb) ring_counts[(size_t)b * RADIAL_MAX_LAYERS + threadIdx.x] = cnt[threadIdx.x];
}

// ---- radial: paint each ROI pixel with its ring's droplet count
__global__ void radial_paint_kernel(const uint8_t* __restrict__ roi, const int* __restrict__ centroid,
                                    const unsigned long long* __restrict__ maxd2, const int* __restrict__ counts,
                                    const double* __restrict__ bounds_g, const int* __restrict__ ring_counts, int nb,
                                    float* __restrict__ out, int H, int W) {
    __shared__ double bounds[RADIAL_MAX_LAYERS + 1];
    __shared__ int cnt[RADIAL_MAX_LAYERS];
    const int b = blockIdx.y;
    if (threadIdx.x <= nb) bounds[threadIdx.x] = bounds_g[(size_t)b * (RADIAL_MAX_LAYERS + 1) + threadIdx.x];
    if (threadIdx.x < nb) cnt[threadIdx.x] = ring_counts[(size_t)b * RADIAL_MAX_LAYERS + threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int HW = H * W;
    if (i >= HW) return;
    const size_t img = (size_t)b * HW;
    float v = 0.f;
    // no ROI pixel or no droplet: all zeros (quantify_pipline.py:71-72)
    if (roi[img + i] && maxd2[(size_t)b * 2 + 1] != 0 && counts[b] > 0) {
        const long long dx = (i % W) - (long long)centroid[2 * b + 1], dy = (i / W) - (long long)centroid[2 * b];
        const double d = sqrt((double)(dx * dx + dy * dy));
        const int k = ring_of(d, bounds, nb);
        if (k >= 0) v = (float)cnt[k];
    }
    out[img + i] = v;
}

// ---- spatial: one pass of scipy's symmetric correlate1d, f64 accumulation, f32 result
struct GaussWeights {
    int radius;
    double w[DC_GAUSS_MAX_RADIUS + 1];       // w[k] = weight at distance k from the centre
};

// Both planes (droplet mask and ROI) go through the same pass together; a block stages its tile plus `radius` halo
// rows / columns in shared memory as f32 (u8 and f32 inputs are both exact in it).
constexpr int GT = 64;              // tile edge along x (threads per row of the block)
constexpr int GT_ROWS0 = 64;        // axis-0 pass: rows per tile (16 per thread)
constexpr int GT_ROWS1 = 16;        // axis-1 pass: rows per tile (4 per thread)

__device__ __forceinline__ double gauss_dot(const float* c, int stride, const GaussWeights& gw) {
    double acc = __dmul_rn((double)c[0], gw.w[0]);
    for (int k = gw.radius; k >= 1; --k)
        acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)c[-k * stride], (double)c[k * stride]), gw.w[k]));
    return acc;
}

// axis 0 (along y) of both u8 planes -> two f32 planes
__global__ void __launch_bounds__(256) gauss_axis0_kernel(const uint8_t* __restrict__ in_m, const uint8_t* __restrict__ in_r,
                                                          float* __restrict__ out_m, float* __restrict__ out_r, int H, int W,
                                                          const __grid_constant__ GaussWeights gw) {
    extern __shared__ float gs[];
    const int r = gw.radius, rows = GT_ROWS0 + 2 * r;
    float* sm = gs;
    float* sr = gs + (size_t)rows * GT;
    const size_t img = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * GT, y0 = blockIdx.y * GT_ROWS0;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int x = x0 + tx;
    for (int ly = ty; ly < rows; ly += 4) {
        const int y = reflect_half(y0 + ly - r, H);
        float vm = 0.f, vr = 0.f;
        if (x < W) { vm = (float)in_m[img + (size_t)y * W + x]; vr = (float)in_r[img + (size_t)y * W + x]; }
        sm[ly * GT + tx] = vm;
        sr[ly * GT + tx] = vr;
    }
    __syncthreads();
    if (x >= W) return;
    for (int ly = ty; ly < GT_ROWS0; ly += 4) {
        const int y = y0 + ly;
        if (y >= H) break;
        out_m[img + (size_t)y * W + x] = (float)gauss_dot(sm + (ly + r) * GT + tx, GT, gw);
        out_r[img + (size_t)y * W + x] = (float)gauss_dot(sr + (ly + r) * GT + tx, GT, gw);
    }
}

// axis 1 (along x) of both f32 planes, then 100 * g(mask) / (g(roi) + 1e-5) in f32
__global__ void __launch_bounds__(256) gauss_axis1_ratio_kernel(const float* __restrict__ in_m, const float* __restrict__ in_r,
                                                                float* __restrict__ out, int H, int W,
                                                                const __grid_constant__ GaussWeights gw) {
    extern __shared__ float gs[];
    const int r = gw.radius, cols = GT + 2 * r;
    float* sm = gs;
    float* sr = gs + (size_t)GT_ROWS1 * cols;
    const size_t img = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * GT, y0 = blockIdx.y * GT_ROWS1;
    const int tid = threadIdx.y * GT + threadIdx.x;
    for (int i = tid; i < GT_ROWS1 * cols; i += 256) {
        const int ly = i / cols, lx = i - ly * cols;
        const int y = y0 + ly, xs = reflect_half(x0 + lx - r, W);
        float vm = 0.f, vr = 0.f;
        if (y < H) { vm = in_m[img + (size_t)y * W + xs]; vr = in_r[img + (size_t)y * W + xs]; }
        sm[i] = vm;
        sr[i] = vr;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= W) return;
    for (int ly = threadIdx.y; ly < GT_ROWS1; ly += 4) {
        const int y = y0 + ly;
        if (y >= H) break;
        const float gm = (float)gauss_dot(sm + ly * cols + threadIdx.x + r, 1, gw);
        const float gr = (float)gauss_dot(sr + ly * cols + threadIdx.x + r, 1, gw);
        out[img + (size_t)y * W + x] = __fmul_rn(__fdiv_rn(gm, __fadd_rn(gr, 1e-5f)), 100.f);
    }
}

size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

size_t roi_workspace_bytes(int B, int H, int W) {
    const size_t hw = (size_t)H * W * B;
    const size_t words = (size_t)B * H * ((W + 31) / 32);
    return al256(hw) + 2 * al256(words * 4) + al256((size_t)B * 256 * 4) + al256((size_t)B * 4) + al256((size_t)B * 3 * 8);
}

int launch_roi_mask(const dc_roi_args_t* a, cudaStream_t stream) {
    DC_REQUIRE(a && a->rgb && a->roi && a->centroid, DC_EINVAL, "dc_roi_mask: null pointer argument");
    DC_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->B <= 65535 && a->H <= 65535, DC_EINVAL, "dc_roi_mask: bad shape %d %d %d",
               a->B, a->H, a->W);
    DC_REQUIRE((long long)a->H * a->W < (1ll << 31), DC_EINVAL, "dc_roi_mask: image too large for int32 indices");
    const int B = a->B, H = a->H, W = a->W, Wp = (W + 31) / 32;
    DC_REQUIRE(a->workspace && a->workspace_bytes >= roi_workspace_bytes(B, H, W), DC_EWORKSPACE,
               "dc_roi_mask: workspace too small (%zu < %zu)", a->workspace_bytes, roi_workspace_bytes(B, H, W));
    const size_t hw = (size_t)H * W * B, words = (size_t)B * H * Wp;
    char* p = (char*)a->workspace;
    uint8_t* blurred = (uint8_t*)p;              p += al256(hw);
    uint32_t* b0 = (uint32_t*)p;                 p += al256(words * 4);
    uint32_t* b1 = (uint32_t*)p;                 p += al256(words * 4);
    unsigned int* hist = (unsigned int*)p;       p += al256((size_t)B * 256 * 4);
    int* thresh = (int*)p;                       p += al256((size_t)B * 4);
    unsigned long long* sums = (unsigned long long*)p;

    DC_CUDA(cudaMemsetAsync(hist, 0, (size_t)B * 256 * 4, stream));
    DC_CUDA(cudaMemsetAsync(sums, 0, (size_t)B * 3 * 8, stream));
    roi_blur_hist_kernel<<<dim3(ceil_div(W, BT_W), ceil_div(H, BT_H), B), 256, 0, stream>>>(a->rgb, blurred, hist, H, W);
    roi_otsu_kernel<<<ceil_div(B, 64), 64, 0, stream>>>(hist, thresh, B, H * W);
    dim3 pg(ceil_div(W, 256), H, B);
    roi_pack_kernel<<<pg, 256, 0, stream>>>(blurred, thresh, b0, H, W, Wp);
    // MORPH_CLOSE = erode(dilate(m)), MORPH_OPEN = dilate(erode(m)), 15 x 15 rectangle; the two erodes in the middle
    // compose into one 29 x 29 erode (out-of-image taps ignored in both forms)
    dim3 mg(ceil_div(Wp, 64), H, B);
    roi_morph_h_kernel<false><<<mg, 64, 0, stream>>>(b0, b1, H, W, Wp, 7);
    roi_morph_v_kernel<false><<<mg, 64, 0, stream>>>(b1, b0, H, Wp, 7);
    roi_morph_h_kernel<true><<<mg, 64, 0, stream>>>(b0, b1, H, W, Wp, 14);
    roi_morph_v_kernel<true><<<mg, 64, 0, stream>>>(b1, b0, H, Wp, 14);
    roi_morph_h_kernel<false><<<mg, 64, 0, stream>>>(b0, b1, H, W, Wp, 7);
    roi_morph_v_kernel<false><<<mg, 64, 0, stream>>>(b1, b0, H, Wp, 7);
    roi_finish_kernel<<<pg, 256, 0, stream>>>(b0, a->roi, sums, H, W, Wp);
    roi_centroid_kernel<<<ceil_div(B, 64), 64, 0, stream>>>(sums, a->centroid, B, H, W);
    DC_CUDA(cudaGetLastError());
    return DC_OK;
}

size_t radial_workspace_bytes(int B) {
    return al256((size_t)B * 2 * 8) + al256((size_t)B * (RADIAL_MAX_LAYERS + 1) * 8) + al256((size_t)B * RADIAL_MAX_LAYERS * 4);
}

int launch_radial_density(const dc_radial_args_t* a, cudaStream_t stream) {
    DC_REQUIRE(a && a->roi && a->centroid && a->counts && a->centroid0 && a->centroid1 && a->out, DC_EINVAL,
               "dc_radial_density: null pointer argument");
    DC_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->B <= 65535 && a->capacity > 0, DC_EINVAL,
               "dc_radial_density: bad shape %d %d %d cap %d", a->B, a->H, a->W, a->capacity);
    DC_REQUIRE(a->nb_layers >= 1 && a->nb_layers <= RADIAL_MAX_LAYERS, DC_EINVAL, "dc_radial_density: nb_layers %d (1..%d)",
               a->nb_layers, RADIAL_MAX_LAYERS);
    DC_REQUIRE((long long)a->H * a->W < (1ll << 31), DC_EINVAL, "dc_radial_density: image too large for int32 indices");
    const int B = a->B, H = a->H, W = a->W;
    DC_REQUIRE(a->workspace && a->workspace_bytes >= radial_workspace_bytes(B), DC_EWORKSPACE,
               "dc_radial_density: workspace too small (%zu < %zu)", a->workspace_bytes, radial_workspace_bytes(B));
    char* p = (char*)a->workspace;
    unsigned long long* maxd2 = (unsigned long long*)p;  p += al256((size_t)B * 2 * 8);
    double* bounds = (double*)p;                         p += al256((size_t)B * (RADIAL_MAX_LAYERS + 1) * 8);
    int* ring_counts = (int*)p;
    DC_CUDA(cudaMemsetAsync(maxd2, 0, (size_t)B * 2 * 8, stream));
    dim3 g(ceil_div(H * W, 256), B);
    radial_maxd2_kernel<<<g, 256, 0, stream>>>(a->roi, a->centroid, maxd2, H, W);
    radial_rings_kernel<<<B, 256, 0, stream>>>(maxd2, a->centroid, a->counts, a->centroid0, a->centroid1, a->capacity,
                                               a->nb_layers, bounds, ring_counts);
    radial_paint_kernel<<<g, 256, 0, stream>>>(a->roi, a->centroid, maxd2, a->counts, bounds, ring_counts, a->nb_layers,
                                               a->out, H, W);
    DC_CUDA(cudaGetLastError());
    return DC_OK;
}

size_t spatial_workspace_bytes(int B, int H, int W) { return 2 * al256((size_t)B * H * W * 4); }

int launch_spatial_density(const dc_spatial_args_t* a, cudaStream_t stream) {
    DC_REQUIRE(a && a->mask && a->roi && a->out && a->weights, DC_EINVAL, "dc_spatial_density: null pointer argument");
    DC_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->B <= 65535, DC_EINVAL, "dc_spatial_density: bad shape %d %d %d", a->B, a->H,
               a->W);
    DC_REQUIRE(a->radius >= 0 && a->radius <= DC_GAUSS_MAX_RADIUS, DC_EINVAL, "dc_spatial_density: radius %d (0..%d)", a->radius,
               DC_GAUSS_MAX_RADIUS);
    const int B = a->B, H = a->H, W = a->W;
    DC_REQUIRE(ceil_div(H, GT_ROWS1) <= 65535, DC_EINVAL, "dc_spatial_density: image too tall (%d rows)", H);
    DC_REQUIRE(a->workspace && a->workspace_bytes >= spatial_workspace_bytes(B, H, W), DC_EWORKSPACE,
               "dc_spatial_density: workspace too small (%zu < %zu)", a->workspace_bytes, spatial_workspace_bytes(B, H, W));
    GaussWeights gw;
    gw.radius = a->radius;
    for (int k = 0; k <= a->radius; ++k) gw.w[k] = a->weights[a->radius + k];      // symmetric: keep centre..edge
    const size_t n = (size_t)B * H * W;
    char* p = (char*)a->workspace;
    float* tm = (float*)p;   p += al256(n * 4);
    float* tr = (float*)p;
    const size_t smem0 = (size_t)2 * (GT_ROWS0 + 2 * a->radius) * GT * sizeof(float);
    const size_t smem1 = (size_t)2 * GT_ROWS1 * (GT + 2 * a->radius) * sizeof(float);
    // function attributes are per device: set on every launch
    DC_CUDA(cudaFuncSetAttribute(gauss_axis0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
    dim3 blk(GT, 4);
    gauss_axis0_kernel<<<dim3(ceil_div(W, GT), ceil_div(H, GT_ROWS0), B), blk, smem0, stream>>>(a->mask, a->roi, tm, tr, H, W, gw);
    gauss_axis1_ratio_kernel<<<dim3(ceil_div(W, GT), ceil_div(H, GT_ROWS1), B), blk, smem1, stream>>>(tm, tr, a->out, H, W, gw);
    DC_CUDA(cudaGetLastError());
    return DC_OK;
}

}  // namespace dc
