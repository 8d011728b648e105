// morph.cu -- "rolling ball" background correction on the GPU.
//
// Replaces rolling_ball_correction_rgb (reference utils/data_loader.py:11-24), per plane:
//   kernel  = cv2.getStructuringElement(MORPH_ELLIPSE, (radius, radius))        :17
//   bg      = cv2.morphologyEx(channel, MORPH_OPEN, kernel)   erode, then dilate  :19
//   corr    = cv2.subtract(channel, bg)                       saturating u8       :20
//   out     = cv2.normalize(corr, None, 0, 255, NORM_MINMAX)                      :21
//
// The element is a radius x radius flat ellipse (NOT separable, not symmetric for even sizes):
// row i covers columns [j1[i], j2[i]).  erode(y,x) = min_i min_{j in row i} src[y+i-an, x+j-an],
// dilate uses max over the SAME offsets, taps outside the image are ignored (OpenCV semantics).
//
// Kernel plan: a block stages its tile + halo in shared memory, builds a power-of-two
// range-min (or max) table along x (levels 2^0..2^L), then every thread owns 4 adjacent pixels
// and combines, for each of the `radius` element rows, two overlapping 2^l windows with byte-SIMD
// __vminu4 / __vmaxu4.  All arithmetic is u8, so the result is bit-exact.
#include "common.cuh"
#include <math.h>

namespace dc {

namespace {

constexpr int TW = 64;           // tile width  (16 threads x 4 pixels)
constexpr int TH = 64;           // tile height (64: the 49-row halo costs 1.8x the tile, not 2.5x; 2 blocks of 1024 threads per SM)
constexpr int MAX_RADIUS = 100;
constexpr int MAX_LEVELS = 7;    // windows up to 64 wide

struct SERows {
    int k;                        // element size (= radius), anchor = k/2
    int nlevels;                  // levels 0..nlevels-1 are built
    unsigned char j1[MAX_RADIUS]; // first column of row i
    unsigned char j2m[MAX_RADIUS];// j2 - 2^level: start of the second window
    unsigned char lvl[MAX_RADIUS];// level used by row i; 255 = empty row
};

template <bool IS_MAX>
__device__ __forceinline__ unsigned vop(unsigned a, unsigned b) {
    return IS_MAX ? __vmaxu4(a, b) : __vminu4(a, b);
}

// One morphology pass over planes addressed as in[((p/C)*H*W + y*W + x)*C + p%C] (in_c = C) and
// written planar.  SUBTRACT: out = saturate(orig - result), plus a per-plane min/max reduction.
template <bool IS_MAX, bool SUBTRACT>
__global__ void __launch_bounds__(TW / 4 * TH) morph_pass_kernel(const uint8_t* __restrict__ in, int in_c,
                                                         uint8_t* __restrict__ out, const uint8_t* __restrict__ orig,
                                                         int orig_c, int H, int W, int* __restrict__ minmax,
                                                         const __grid_constant__ SERows se) {
    extern __shared__ unsigned smem[];
    const int k = se.k, an = k / 2;
    const int pad = (4 - (an & 3)) & 3;          // extra columns on the left: the region starts on a 4-pixel boundary
    const int RW = TW + k + pad;                 // region width in bytes (tile + halo), halo = k-1 (+1 pad)
    const int pitch = ((RW + 3) >> 2) + 1;       // words per region row (+1 so unaligned fetches stay in-row)
    const int RH = TH + k - 1;
    const int level_words = pitch * RH;
    const unsigned ident = IS_MAX ? 0u : 0xffffffffu;

    const int plane = blockIdx.z;
    const int b = plane / in_c, c = plane % in_c;
    const int x0 = blockIdx.x * TW - an - pad, y0 = blockIdx.y * TH - an;   // region origin in the image (x0 % 4 == 0)
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthreads = blockDim.x * blockDim.y;

    __shared__ int2 rowtab[MAX_RADIUS];
    __shared__ int nrows_s;
    if (tid == 0) {
        int n = 0;
        for (int i = 0; i < k; ++i) {
            const int l = se.lvl[i];
            if (l == 255) continue;
            const int base = (l * level_words + i * pitch) * 4;
            rowtab[n++] = make_int2(base + se.j1[i] + pad, base + se.j2m[i] + pad);
        }
        nrows_s = n;
    }
    // ---- stage level 0 (tile + halo), identity outside the image ----
    uint8_t* s8 = reinterpret_cast<uint8_t*>(smem);
    const uint8_t* src = in + (size_t)b * H * W * in_c + c;
    if (in_c == 1 && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0) {
        // planar input with 4-pixel-aligned rows: one aligned word per 4 region bytes (x0 % 4 == 0 and W % 4 == 0,
        // so a word lies wholly inside or wholly outside the image)
        for (int i = tid; i < level_words; i += nthreads) {
            const int ry = i / pitch, rxw = i - ry * pitch;
            const int gy = y0 + ry, gx = x0 + 4 * rxw;
            unsigned v = ident;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = *reinterpret_cast<const unsigned*>(src + (size_t)gy * W + gx);
            smem[i] = v;
        }
    } else {
        for (int i = tid; i < RH * pitch * 4; i += nthreads) {
            int ry = i / (pitch * 4), rx = i - ry * (pitch * 4);
            int gy = y0 + ry, gx = x0 + rx;
            uint8_t v = IS_MAX ? 0 : 255;
            if (rx < RW && gy >= 0 && gy < H && gx >= 0 && gx < W) v = src[((size_t)gy * W + gx) * in_c];
            s8[i] = v;
        }
    }
    __syncthreads();
    // ---- levels 1..: T_l[x] = op(T_{l-1}[x], T_{l-1}[x + 2^(l-1)]) ----
    for (int l = 1; l < se.nlevels; ++l) {
        const unsigned* prev = smem + (l - 1) * level_words;
        unsigned* cur = smem + l * level_words;
        const int step = 1 << (l - 1);           // bytes
        for (int i = tid; i < level_words; i += nthreads) {
            int rx = i % pitch;
            unsigned a = prev[i], bb;
            if (step < 4) {
                unsigned nxt = (rx + 1 < pitch) ? prev[i + 1] : ident;
                bb = __funnelshift_r(a, nxt, step * 8);
            } else {
                int ws = step >> 2;
                bb = (rx + ws < pitch) ? prev[i + ws] : ident;
            }
            cur[i] = vop<IS_MAX>(a, bb);
        }
        __syncthreads();
    }
    // ---- combine the element rows ----
    // sm_100a has no byte-lane SIMD min/max (__vminu4 expands to 7 instructions) but it has the DPX 16-bit-lane
    // three-operand form (VIMNMX3.U16x2).  The accumulator is kept as two words of 16-bit lanes (even / odd
    // pixels); each fetched u8x4 window is split with two PRMTs, so a row costs 4 PRMT + 2 VIMNMX3 instead of 14.
    // rowtab: per non-empty element row, the byte offsets of its two windows inside the level tables (one
    // broadcast LDS.64 instead of three dependent constant-bank loads plus address arithmetic).
    const int lx4 = threadIdx.x * 4, ly = threadIdx.y;
    const uint8_t* tbase = reinterpret_cast<const uint8_t*>(smem) + (ly * pitch) * 4 + lx4;
    const unsigned id16 = IS_MAX ? 0u : 0x00ff00ffu;
    unsigned acc_e = id16, acc_o = id16;
    const int nrows = nrows_s;
#pragma unroll 4
    for (int n = 0; n < nrows; ++n) {
        const int2 t = rowtab[n];
        const unsigned* ra = reinterpret_cast<const unsigned*>(tbase + (t.x & ~3));
        const unsigned* rb = reinterpret_cast<const unsigned*>(tbase + (t.y & ~3));
        const unsigned a = __funnelshift_r(ra[0], ra[1], (t.x & 3) * 8);
        const unsigned bb = __funnelshift_r(rb[0], rb[1], (t.y & 3) * 8);
        const unsigned a_e = __byte_perm(a, 0, 0x4240), a_o = __byte_perm(a, 0, 0x4341);
        const unsigned b_e = __byte_perm(bb, 0, 0x4240), b_o = __byte_perm(bb, 0, 0x4341);
        acc_e = IS_MAX ? __vimax3_u16x2(acc_e, a_e, b_e) : __vimin3_u16x2(acc_e, a_e, b_e);
        acc_o = IS_MAX ? __vimax3_u16x2(acc_o, a_o, b_o) : __vimin3_u16x2(acc_o, a_o, b_o);
    }
    const unsigned acc = __byte_perm(acc_e, acc_o, 0x6240);     // bytes: e.lo, o.lo, e.hi, o.hi
    // ---- write (and, for the second pass, subtract + reduce) ----
    const int gx = blockIdx.x * TW + lx4, gy = blockIdx.y * TH + ly;
    int mn = 255, mx = 0;
    if (gy < H && gx < W) {
        uint8_t* dst = out + ((size_t)plane * H + gy) * W + gx;
        unsigned res = acc;
        if (SUBTRACT) {
            const int ob = plane / orig_c, oc = plane % orig_c;
            const uint8_t* o = orig + ((size_t)ob * H * W + (size_t)gy * W + gx) * orig_c + oc;
            unsigned ov = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (gx + j < W) ov |= (unsigned)o[(size_t)j * orig_c] << (8 * j);
            res = __vsubus4(ov, acc);            // cv2.subtract saturates at 0
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (gx + j < W) {
                int v = (res >> (8 * j)) & 0xff;
                dst[j] = (uint8_t)v;
                mn = min(mn, v); mx = max(mx, v);
            }
        }
    }
    if (SUBTRACT) {
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
        if ((tid & 31) == 0) {
            atomicMin(&minmax[2 * plane], mn);
            atomicMax(&minmax[2 * plane + 1], mx);
        }
    }
}

__global__ void init_minmax_kernel(int* minmax, int planes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < planes) { minmax[2 * i] = 255; minmax[2 * i + 1] = 0; }
}

// cv2.normalize(NORM_MINMAX, 0..255) on u8: scale = 255 * (1/(max-min)) and shift = -min*scale in
// double, then fp32 fma(v, scale, shift) rounded half-to-even (matches cv2 4.13 on every input).
__global__ void __launch_bounds__(256) stretch_kernel(const uint8_t* __restrict__ corr, uint8_t* __restrict__ out,
                                                      int out_c, int H, int W, const int* __restrict__ minmax) {
    __shared__ uint8_t lut[256];
    const int plane = blockIdx.y;
    const int b = plane / out_c, c = plane % out_c;
    {
        int mn = minmax[2 * plane], mx = minmax[2 * plane + 1];
        double scale = (mx - mn) > 0 ? __dmul_rn(255.0, __ddiv_rn(1.0, (double)(mx - mn))) : 0.0;
        double shift = __dsub_rn(0.0, __dmul_rn((double)mn, scale));
        float a = (float)scale, sh = (float)shift;
        int q = __float2int_rn(__fmaf_rn((float)threadIdx.x, a, sh));
        lut[threadIdx.x] = (uint8_t)min(255, max(0, q));
    }
    __syncthreads();
    const size_t HW = (size_t)H * W;
    const uint8_t* src = corr + (size_t)plane * HW;
    uint8_t* dst = out + (size_t)b * HW * out_c + c;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (size_t)gridDim.x * blockDim.x)
        dst[i * out_c] = lut[src[i]];
}

int build_rows(int radius, SERows* se) {
    const int k = radius;
    se->k = k;
    int r = k / 2, c = k / 2;
    double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    int maxw = 1;
    for (int i = 0; i < k; ++i) {
        int dy = i - r, j1 = 0, j2 = 0;
        if (abs(dy) <= r) {
            int dx = (int)nearbyint(c * sqrt(((double)r * r - (double)dy * dy) * inv_r2));   // cvRound
            j1 = c - dx < 0 ? 0 : c - dx;
            j2 = c + dx + 1 > k ? k : c + dx + 1;
        }
        int w = j2 - j1;
        if (w <= 0) { se->lvl[i] = 255; se->j1[i] = 0; se->j2m[i] = 0; continue; }
        int l = 0;
        while ((2 << l) <= w) ++l;               // 2^l <= w < 2^(l+1)
        se->lvl[i] = (unsigned char)l;
        se->j1[i] = (unsigned char)j1;
        se->j2m[i] = (unsigned char)(j2 - (1 << l));
        if (w > maxw) maxw = w;
    }
    int nl = 1;
    while ((1 << nl) <= maxw) ++nl;
    se->nlevels = nl;
    return nl <= MAX_LEVELS ? 0 : -1;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

size_t rolling_ball_workspace_bytes(int planes, int H, int W) {
    return 2 * align256((size_t)planes * H * W) + align256(sizeof(int) * 2 * (size_t)planes);
}

int launch_rolling_ball(const dc_rolling_ball_args_t* a, cudaStream_t stream) {
    DC_REQUIRE(a && a->in && a->out, DC_EINVAL, "dc_rolling_ball: null pointer argument");
    DC_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->C > 0, DC_EINVAL, "dc_rolling_ball: bad shape");
    DC_REQUIRE(a->radius >= 1 && a->radius <= MAX_RADIUS, DC_EINVAL, "dc_rolling_ball: radius %d outside [1,%d]",
               a->radius, MAX_RADIUS);
    const int planes = a->B * a->C, H = a->H, W = a->W;
    DC_REQUIRE(planes <= 65535, DC_EINVAL, "dc_rolling_ball: more than 65535 planes in one call");
    DC_REQUIRE(a->workspace && a->workspace_bytes >= rolling_ball_workspace_bytes(planes, H, W), DC_EWORKSPACE,
               "dc_rolling_ball: workspace too small");
    SERows se;
    DC_REQUIRE(build_rows(a->radius, &se) == 0, DC_EINVAL, "dc_rolling_ball: element too wide");

    char* p = (char*)a->workspace;
    uint8_t* er = (uint8_t*)p;   p += align256((size_t)planes * H * W);
    uint8_t* corr = (uint8_t*)p; p += align256((size_t)planes * H * W);
    int* minmax = (int*)p;

    const int an = se.k / 2, pad = (4 - (an & 3)) & 3;
    const int RW = TW + se.k + pad, pitch = ((RW + 3) >> 2) + 1, RH = TH + se.k - 1;
    const size_t smem = (size_t)se.nlevels * pitch * RH * 4;
    DC_REQUIRE(smem <= 226 * 1024, DC_EINVAL, "dc_rolling_ball: radius %d needs %zu B of shared memory", a->radius, smem);
    // function attributes are per device: set on every launch (a host-side call of well under a microsecond)
    DC_CUDA(cudaFuncSetAttribute(morph_pass_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));   // + 816 B static (rowtab)
    DC_CUDA(cudaFuncSetAttribute(morph_pass_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));   // + 816 B static (rowtab)
    dim3 block(TW / 4, TH);
    dim3 grid(ceil_div(W, TW), ceil_div(H, TH), planes);
    init_minmax_kernel<<<ceil_div(planes, 256), 256, 0, stream>>>(minmax, planes);
    morph_pass_kernel<false, false><<<grid, block, smem, stream>>>(a->in, a->C, er, nullptr, 0, H, W, nullptr, se);
    morph_pass_kernel<true, true><<<grid, block, smem, stream>>>(er, 1, corr, a->in, a->C, H, W, minmax, se);
    int sblocks = ceil_div(H * W, 256 * 8);
    if (sblocks > 4096) sblocks = 4096;
    stretch_kernel<<<dim3(sblocks, planes), 256, 0, stream>>>(corr, a->out, a->C, H, W, minmax);
    DC_CUDA(cudaGetLastError());
    return DC_OK;
}

}  // namespace dc
