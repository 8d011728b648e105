// stem.cu -- first layer of UNetDC: Conv2d(3, 64, 3, padding=d, dilation=d) + BatchNorm + ReLU
// (reference models/model_2.py:10 -> :41-46, first conv of enc1).  K = 27 is too thin for the
// tensor cores to matter (3.6 of 1541 GFLOP per 1024^2 image), and the layer is bound by writing
// its 64-channel bf16 output, so it is a direct fp32 convolution that also does the layout change
// NCHW fp32 (or u8 image / 255, reference quantify_droplets_batch.py:45) -> NHWC bf16.
#include "common.cuh"

namespace dc {

namespace {

constexpr int COUT = 64;

template <int IN_KIND>
__device__ __forceinline__ float load_px(const void* in, int b, int c, int y, int x, int H, int W) {
    if (y < 0 || y >= H || x < 0 || x >= W) return 0.f;   // zero padding
    if (IN_KIND == 0) {
        return reinterpret_cast<const float*>(in)[(((size_t)b * 3 + c) * H + y) * W + x];
    } else if (IN_KIND == 1) {
        return __fdiv_rn((float)reinterpret_cast<const uint8_t*>(in)[((size_t)b * H + y) * W + x], 255.0f);
    } else {
        return __fdiv_rn((float)reinterpret_cast<const uint8_t*>(in)[(((size_t)b * H + y) * W + x) * 3 + c], 255.0f);
    }
}

// weight: fp32 [64][3][3][3] (co, ci, ky, kx) with BN folded; staged as ws[(ci*9 + ky*3 + kx)][co].
template <int IN_KIND>
__global__ void __launch_bounds__(128) stem_kernel(const void* __restrict__ in, const float* __restrict__ weight,
                                                   const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                   int B, int H, int W, int d, int out_stride, int out_offset) {
    __shared__ __align__(16) float ws[27 * COUT];
    __shared__ __align__(16) float bs[COUT];
    for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) {
        int co = i % COUT, t = i / COUT;
        ws[i] = weight[co * 27 + t];
    }
    if (threadIdx.x < COUT) bs[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();

    const size_t npix = (size_t)B * H * W;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((size_t)W * H));
        float v[27];
        if (IN_KIND == 1) {
            // grayscale: the three input channels are the same plane
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                float g = load_px<1>(in, b, 0, y + (t / 3 - 1) * d, x + (t % 3 - 1) * d, H, W);
                v[t] = g; v[9 + t] = g; v[18 + t] = g;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int t = 0; t < 9; ++t)
                    v[c * 9 + t] = load_px<IN_KIND>(in, b, c, y + (t / 3 - 1) * d, x + (t % 3 - 1) * d, H, W);
        }
        __nv_bfloat16* o = out + p * out_stride + out_offset;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float acc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = bs[half * 32 + j];
#pragma unroll
            for (int t = 0; t < 27; ++t) {
                const float4* w4 = reinterpret_cast<const float4*>(&ws[t * COUT + half * 32]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 w = w4[j];
                    acc[4 * j + 0] = fmaf(v[t], w.x, acc[4 * j + 0]);
                    acc[4 * j + 1] = fmaf(v[t], w.y, acc[4 * j + 1]);
                    acc[4 * j + 2] = fmaf(v[t], w.z, acc[4 * j + 2]);
                    acc[4 * j + 3] = fmaf(v[t], w.w, acc[4 * j + 3]);
                }
            }
            uint4* o4 = reinterpret_cast<uint4*>(o + half * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(fmaxf(acc[8 * j + 0], 0.f), fmaxf(acc[8 * j + 1], 0.f));
                __nv_bfloat162 h1 = __floats2bfloat162_rn(fmaxf(acc[8 * j + 2], 0.f), fmaxf(acc[8 * j + 3], 0.f));
                __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaxf(acc[8 * j + 4], 0.f), fmaxf(acc[8 * j + 5], 0.f));
                __nv_bfloat162 h3 = __floats2bfloat162_rn(fmaxf(acc[8 * j + 6], 0.f), fmaxf(acc[8 * j + 7], 0.f));
                uint4 u;
                u.x = *reinterpret_cast<unsigned*>(&h0); u.y = *reinterpret_cast<unsigned*>(&h1);
                u.z = *reinterpret_cast<unsigned*>(&h2); u.w = *reinterpret_cast<unsigned*>(&h3);
                o4[j] = u;
            }
        }
    }
}

}  // namespace

int launch_stem(const dc_stem_args_t* a, cudaStream_t stream) {
    DC_REQUIRE(a && a->in && a->weight && a->bias && a->out, DC_EINVAL, "dc_stem: null pointer argument");
    DC_REQUIRE(a->Cout == COUT, DC_EINVAL, "dc_stem: Cout must be 64 (got %d)", a->Cout);
    DC_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->dilation >= 1, DC_EINVAL, "dc_stem: bad shape");
    DC_REQUIRE(a->out_stride % 8 == 0 && a->out_offset % 8 == 0 && a->out_stride >= a->out_offset + COUT, DC_EINVAL,
               "dc_stem: output stride/offset must be multiples of 8 channels");
    DC_REQUIRE(a->in_kind >= 0 && a->in_kind <= 2, DC_EINVAL, "dc_stem: in_kind %d", a->in_kind);
    const size_t npix = (size_t)a->B * a->H * a->W;
    int blocks = (int)((npix + 127) / 128);
    const int cap = num_sms() * 16;
    if (blocks > cap) blocks = cap;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a->out);
    switch (a->in_kind) {
        case 0: stem_kernel<0><<<blocks, 128, 0, stream>>>(a->in, a->weight, a->bias, out, a->B, a->H, a->W, a->dilation, a->out_stride, a->out_offset); break;
        case 1: stem_kernel<1><<<blocks, 128, 0, stream>>>(a->in, a->weight, a->bias, out, a->B, a->H, a->W, a->dilation, a->out_stride, a->out_offset); break;
        default: stem_kernel<2><<<blocks, 128, 0, stream>>>(a->in, a->weight, a->bias, out, a->B, a->H, a->W, a->dilation, a->out_stride, a->out_offset); break;
    }
    DC_CUDA(cudaGetLastError());
    return DC_OK;
}

}  // namespace dc
