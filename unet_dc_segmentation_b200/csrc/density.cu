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
//   MORPH_CLOSE / OPEN     15x15 rectangle, out-of-image taps ignored; separable min / max passes
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

// ---- ROI: gray + horizontal blur pass -> u16 (<= 255 * 256)
__global__ void roi_gray_hblur_kernel(const uint8_t* __restrict__ rgb, uint16_t* __restrict__ hb, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t img = (size_t)blockIdx.z * H * W;
    const uint8_t* row = rgb + (img + (size_t)y * W) * 3;
    int acc = 0;
#pragma unroll
    for (int j = 0; j < 15; ++j) {
        const int xs = reflect101(x + j - 7, W);
        const int r = row[3 * xs], g = row[3 * xs + 1], b = row[3 * xs + 2];
        const int gray = (r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15;
        acc += gray * kGauss15[j];
    }
    hb[img + (size_t)y * W + x] = (uint16_t)acc;
}

// ---- ROI: vertical blur pass -> u8 + per-image histogram
__global__ void roi_vblur_hist_kernel(const uint16_t* __restrict__ hb, uint8_t* __restrict__ blurred,
                                      unsigned int* __restrict__ hist, int H, int W) {
    __shared__ unsigned int sh[256];
    const int tid = threadIdx.x;
    sh[tid & 255] = 0;
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const size_t img = (size_t)blockIdx.z * H * W;
    if (x < W) {
        long long acc = 0;
#pragma unroll
        for (int j = 0; j < 15; ++j) acc += (long long)hb[img + (size_t)reflect101(y + j - 7, H) * W + x] * kGauss15[j];
        const int v = (int)((acc + (1 << 15)) >> 16);
        blurred[img + (size_t)y * W + x] = (uint8_t)v;
        atomicAdd(&sh[v], 1u);
    }
    __syncthreads();
    if (sh[tid & 255] && tid < 256) atomicAdd(&hist[(size_t)blockIdx.z * 256 + tid], sh[tid]);
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

// ---- ROI: one separable pass of a 15-wide rectangular erode / dilate on a 0/255 plane; out-of-image taps ignored.
// thresh != NULL: the input is the blurred plane and is binarised first (in > thresh[b]).
template <bool MAX, bool HORIZ>
__global__ void roi_rect15_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int* __restrict__ thresh,
                                  int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t img = (size_t)blockIdx.z * H * W;
    const int t = thresh ? thresh[blockIdx.z] : 0;
    int acc = MAX ? 0 : 255;
#pragma unroll
    for (int j = -7; j <= 7; ++j) {
        const int xs = HORIZ ? x + j : x, ys = HORIZ ? y : y + j;
        if (xs < 0 || xs >= W || ys < 0 || ys >= H) continue;
        int v = in[img + (size_t)ys * W + xs];
        if (thresh) v = v > t ? 255 : 0;
        acc = MAX ? max(acc, v) : min(acc, v);
    }
    out[img + (size_t)y * W + x] = (uint8_t)acc;
}

// ---- ROI: last dilate pass output -> {0,1} + moment sums
__global__ void roi_finish_kernel(const uint8_t* __restrict__ m, uint8_t* __restrict__ roi,
                                  unsigned long long* __restrict__ sums, int H, int W) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int HW = H * W;
    const size_t img = (size_t)blockIdx.y * HW;
    unsigned long long s0 = 0, sx = 0, sy = 0;
    if (i < HW) {
        const int v = m[img + i] > 0;
        roi[img + i] = (uint8_t)v;
        if (v) { s0 = 1; sx = (unsigned long long)(i % W); sy = (unsigned long long)(i / W); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_down_sync(0xffffffffu, s0, o);
        sx += __shfl_down_sync(0xffffffffu, sx, o);
        sy += __shfl_down_sync(0xffffffffu, sy, o);
    }
    if ((threadIdx.x & 31) == 0 && s0) {
        atomicAdd(&sums[(size_t)blockIdx.y * 3 + 0], s0);
        atomicAdd(&sums[(size_t)blockIdx.y * 3 + 1], sx);
        atomicAdd(&sums[(size_t)blockIdx.y * 3 + 2], sy);
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

template <typename TIn, bool AXIS0>
__global__ void gauss_pass_kernel(const TIn* __restrict__ in, float* __restrict__ out, int H, int W,
                                  const __grid_constant__ GaussWeights gw) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t img = (size_t)blockIdx.z * H * W;
    const int n = AXIS0 ? H : W, c = AXIS0 ? y : x;
    auto at = [&](int k) -> double {
        const int r = reflect_half(k, n);
        return (double)in[img + (AXIS0 ? (size_t)r * W + x : (size_t)y * W + r)];
    };
    double acc = __dmul_rn(at(c), gw.w[0]);
    for (int k = gw.radius; k >= 1; --k) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(at(c - k), at(c + k)), gw.w[k]));
    out[img + (size_t)y * W + x] = (float)acc;
}

__global__ void density_ratio_kernel(const float* __restrict__ gm, const float* __restrict__ gr, float* __restrict__ out,
                                     size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = __fmul_rn(__fdiv_rn(gm[i], __fadd_rn(gr[i], 1e-5f)), 100.f);
}

size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

size_t roi_workspace_bytes(int B, int H, int W) {
    const size_t hw = (size_t)H * W * B;
    return al256(hw * 2) + 2 * al256(hw) + al256((size_t)B * 256 * 4) + al256((size_t)B * 4) + al256((size_t)B * 3 * 8);
}

int launch_roi_mask(const dc_roi_args_t* a, cudaStream_t stream) {
    DC_REQUIRE(a && a->rgb && a->roi && a->centroid, DC_EINVAL, "dc_roi_mask: null pointer argument");
    DC_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->B <= 65535 && a->H <= 65535, DC_EINVAL, "dc_roi_mask: bad shape %d %d %d",
               a->B, a->H, a->W);
    DC_REQUIRE((long long)a->H * a->W < (1ll << 31), DC_EINVAL, "dc_roi_mask: image too large for int32 indices");
    const int B = a->B, H = a->H, W = a->W;
    DC_REQUIRE(a->workspace && a->workspace_bytes >= roi_workspace_bytes(B, H, W), DC_EWORKSPACE,
               "dc_roi_mask: workspace too small (%zu < %zu)", a->workspace_bytes, roi_workspace_bytes(B, H, W));
    const size_t hw = (size_t)H * W * B;
    char* p = (char*)a->workspace;
    uint16_t* hb = (uint16_t*)p;                 p += al256(hw * 2);
    uint8_t* t0 = (uint8_t*)p;                   p += al256(hw);
    uint8_t* t1 = (uint8_t*)p;                   p += al256(hw);
    unsigned int* hist = (unsigned int*)p;       p += al256((size_t)B * 256 * 4);
    int* thresh = (int*)p;                       p += al256((size_t)B * 4);
    unsigned long long* sums = (unsigned long long*)p;

    DC_CUDA(cudaMemsetAsync(hist, 0, (size_t)B * 256 * 4, stream));
    DC_CUDA(cudaMemsetAsync(sums, 0, (size_t)B * 3 * 8, stream));
    dim3 blk(256), grd(ceil_div(W, 256), H, B);
    roi_gray_hblur_kernel<<<grd, blk, 0, stream>>>(a->rgb, hb, H, W);
    roi_vblur_hist_kernel<<<grd, blk, 0, stream>>>(hb, t0, hist, H, W);                       // t0 = blurred
    roi_otsu_kernel<<<ceil_div(B, 64), 64, 0, stream>>>(hist, thresh, B, H * W);
    // close = erode(dilate(m)), open = dilate(erode(m)); each a horizontal + a vertical pass
    roi_rect15_kernel<true, true><<<grd, blk, 0, stream>>>(t0, t1, thresh, H, W);             // threshold + dilate x
    roi_rect15_kernel<true, false><<<grd, blk, 0, stream>>>(t1, t0, nullptr, H, W);           // dilate y
    roi_rect15_kernel<false, true><<<grd, blk, 0, stream>>>(t0, t1, nullptr, H, W);           // erode x
    roi_rect15_kernel<false, false><<<grd, blk, 0, stream>>>(t1, t0, nullptr, H, W);          // erode y   (closed)
    roi_rect15_kernel<false, true><<<grd, blk, 0, stream>>>(t0, t1, nullptr, H, W);           // erode x
    roi_rect15_kernel<false, false><<<grd, blk, 0, stream>>>(t1, t0, nullptr, H, W);          // erode y
    roi_rect15_kernel<true, true><<<grd, blk, 0, stream>>>(t0, t1, nullptr, H, W);            // dilate x
    roi_rect15_kernel<true, false><<<grd, blk, 0, stream>>>(t1, t0, nullptr, H, W);           // dilate y  (opened)
    dim3 fg(ceil_div(H * W, 256), B);
    roi_finish_kernel<<<fg, 256, 0, stream>>>(t0, a->roi, sums, H, W);
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

size_t spatial_workspace_bytes(int B, int H, int W) { return 3 * al256((size_t)B * H * W * 4); }

int launch_spatial_density(const dc_spatial_args_t* a, cudaStream_t stream) {
    DC_REQUIRE(a && a->mask && a->roi && a->out && a->weights, DC_EINVAL, "dc_spatial_density: null pointer argument");
    DC_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->B <= 65535 && a->H <= 65535, DC_EINVAL,
               "dc_spatial_density: bad shape %d %d %d", a->B, a->H, a->W);
    DC_REQUIRE(a->radius >= 0 && a->radius <= DC_GAUSS_MAX_RADIUS, DC_EINVAL, "dc_spatial_density: radius %d (0..%d)", a->radius,
               DC_GAUSS_MAX_RADIUS);
    const int B = a->B, H = a->H, W = a->W;
    DC_REQUIRE(a->workspace && a->workspace_bytes >= spatial_workspace_bytes(B, H, W), DC_EWORKSPACE,
               "dc_spatial_density: workspace too small (%zu < %zu)", a->workspace_bytes, spatial_workspace_bytes(B, H, W));
    GaussWeights gw;
    gw.radius = a->radius;
    for (int k = 0; k <= a->radius; ++k) gw.w[k] = a->weights[a->radius + k];      // symmetric: keep centre..edge
    const size_t n = (size_t)B * H * W;
    char* p = (char*)a->workspace;
    float* t = (float*)p;    p += al256(n * 4);
    float* gm = (float*)p;   p += al256(n * 4);
    float* gr = (float*)p;
    dim3 blk(256), grd(ceil_div(W, 256), H, B);
    gauss_pass_kernel<uint8_t, true><<<grd, blk, 0, stream>>>(a->mask, t, H, W, gw);
    gauss_pass_kernel<float, false><<<grd, blk, 0, stream>>>(t, gm, H, W, gw);
    gauss_pass_kernel<uint8_t, true><<<grd, blk, 0, stream>>>(a->roi, t, H, W, gw);
    gauss_pass_kernel<float, false><<<grd, blk, 0, stream>>>(t, gr, H, W, gw);
    density_ratio_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(gm, gr, a->out, n);
    DC_CUDA(cudaGetLastError());
    return DC_OK;
}

}  // namespace dc
