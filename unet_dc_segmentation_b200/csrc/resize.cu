// resize.cu -- the two cv2.resize calls of the hot path on the GPU.
//
// Reference quantify_droplets_batch.py:44 (`cv2.resize(im, (IMG_SIZE, IMG_SIZE), cv2.INTER_AREA)`) and :57
// (`cv2.resize(mask512, (ow, oh), cv2.INTER_NEAREST)`) both pass the interpolation flag in the `dst` slot, so both
// run OpenCV's default INTER_LINEAR on 8-bit data: 11-bit fixed-point coefficients, horizontal pass in int,
// vertical pass (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2.  x fractions are clamped at the borders, y
// weights are not (the row indices are clipped instead).  Integer work: bit-exact against cv2.
//
// One thread per output pixel (all channels); coefficients are recomputed per thread with the same IEEE
// double / float sequence OpenCV uses (no fast-math), which is cheaper than a table round trip through HBM.
// HBM-bound: algorithmic bytes = src read once + dst written once.
#include "common.cuh"

namespace dc {

namespace {

struct LinCoef { int i0, i1, a0, a1; };

__device__ __forceinline__ LinCoef lin_coef(int d, double scale, int sn, bool clamp_frac) {
    float f = (float)__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    LinCoef c;
    if (clamp_frac) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= sn - 1) { s = sn - 1; f = 0.f; }
        c.i0 = s;
        c.i1 = min(s + 1, sn - 1);
    } else {
        c.i0 = min(max(s, 0), sn - 1);
        c.i1 = min(max(s + 1, 0), sn - 1);
    }
    c.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));    // saturate_cast<short>: round half to even
    c.a1 = __float2int_rn(__fmul_rn(f, 2048.f));
    return c;
}

template <int CN>
__global__ void __launch_bounds__(256) resize_linear_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                               int sh, int sw, int dh, int dw, double scale_x,
                                                               double scale_y) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= dw) return;
    const LinCoef cx = lin_coef(x, scale_x, sw, true);
    const LinCoef cy = lin_coef(y, scale_y, sh, false);
    const uint8_t* s = src + (size_t)blockIdx.z * sh * sw * CN;
    const uint8_t* r0 = s + (size_t)cy.i0 * sw * CN;
    const uint8_t* r1 = s + (size_t)cy.i1 * sw * CN;
    uint8_t* o = dst + (((size_t)blockIdx.z * dh + y) * dw + x) * CN;
#pragma unroll
    for (int c = 0; c < CN; ++c) {
        const int S0 = r0[cx.i0 * CN + c] * cx.a0 + r0[cx.i1 * CN + c] * cx.a1;
        const int S1 = r1[cx.i0 * CN + c] * cx.a0 + r1[cx.i1 * CN + c] * cx.a1;
        const int v = (((cy.a0 * (S0 >> 4)) >> 16) + ((cy.a1 * (S1 >> 4)) >> 16) + 2) >> 2;
        o[c] = (uint8_t)min(255, max(0, v));
    }
}

}  // namespace

int launch_resize_linear_u8(const dc_resize_args_t* a, cudaStream_t stream) {
    DC_REQUIRE(a && a->in && a->out, DC_EINVAL, "dc_resize_linear_u8: null pointer argument");
    DC_REQUIRE(a->B > 0 && a->src_h > 0 && a->src_w > 0 && a->dst_h > 0 && a->dst_w > 0, DC_EINVAL,
               "dc_resize_linear_u8: bad shape");
    DC_REQUIRE(a->C == 1 || a->C == 3, DC_EINVAL, "dc_resize_linear_u8: C must be 1 or 3 (got %d)", a->C);
    DC_REQUIRE(a->dst_h <= 65535 && a->B <= 65535, DC_EINVAL, "dc_resize_linear_u8: dst_h / B above 65535");
    // OpenCV: inv_scale = dsize / ssize; scale = 1. / inv_scale (double)
    const double scale_x = 1.0 / ((double)a->dst_w / (double)a->src_w);
    const double scale_y = 1.0 / ((double)a->dst_h / (double)a->src_h);
    dim3 grid(ceil_div(a->dst_w, 256), a->dst_h, a->B);
    if (a->C == 1)
        resize_linear_u8_kernel<1><<<grid, 256, 0, stream>>>(a->in, a->out, a->src_h, a->src_w, a->dst_h, a->dst_w, scale_x, scale_y);
    else
        resize_linear_u8_kernel<3><<<grid, 256, 0, stream>>>(a->in, a->out, a->src_h, a->src_w, a->dst_h, a->dst_w, scale_x, scale_y);
    DC_CUDA(cudaGetLastError());
    return DC_OK;
}

}  // namespace dc
