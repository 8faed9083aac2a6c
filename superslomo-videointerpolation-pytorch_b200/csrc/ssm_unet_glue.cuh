// ssm_unet_glue.cuh -- element-wise steps BETWEEN the cuDNN convolutions of the two flow U-Nets, for
// channels-last (N x H x W x C) activations in bf16 or fp32 (sm_100a), forward and backward.
//
// The convolutions themselves stay on PyTorch / cuDNN (the one dense contraction of the system).  A whole
// 1080p inference step, however, spends only 26 % of its time in them (profiles/r01s_pipeline_profile.txt):
// ATen's channels-last bilinear upsampling kernel takes 38 % (1.7 ms per call, ~10x off its bandwidth
// bound), the separate bias add after every convolution 12 %, LeakyReLU 3 %, average pooling 6 %.  These
// three kernels do the same arithmetic as the ATen ops they replace, in the same order, at HBM speed:
//
//   upsample2x_nhwc_kernel   F.interpolate(x, size=(2H, 2W), mode="bilinear", align_corners=False)
//                            reference: the lambda upsampleN of scripts/models/flow_computation.py:92-94, 103-105, ... (applied at :236-272)
//   bias_leaky_nhwc_kernel   conv bias add + LeakyReLU(0.1)      reference: layers.conv, scripts/models/layers.py:21-33
//   avgpool2_nhwc_kernel     AvgPool2d(2)                        reference: layers.avg_pool, scripts/models/layers.py:60-63
//
// One thread owns 8 consecutive channels (16 B in bf16, 32 B in fp32) of one pixel, so a warp moves 512 /
// 1024 contiguous bytes per access; all arithmetic is fp32 (ATen's accscalar_t), rounded once on store.
#pragma once
#include "ssm_device.cuh"

namespace ssm {

struct Vec8 { float v[8]; };

template <typename T> __device__ __forceinline__ Vec8 ld8(const T* p);
template <> __device__ __forceinline__ Vec8 ld8<float>(const float* p) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    Vec8 r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
template <> __device__ __forceinline__ Vec8 ld8<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const unsigned w[4] = {q.x, q.y, q.z, q.w};
    Vec8 r;
#pragma unroll
    for (int k = 0; k < 4; ++k) { r.v[2 * k] = __uint_as_float(w[k] << 16); r.v[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u); }
    return r;
}
template <typename T> __device__ __forceinline__ void st8(T* p, const Vec8& r);
template <> __device__ __forceinline__ void st8<float>(float* p, const Vec8& r) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
    __stcs(reinterpret_cast<float4*>(p) + 1, make_float4(r.v[4], r.v[5], r.v[6], r.v[7]));
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const Vec8& r) {
    unsigned w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(r.v[2 * k], r.v[2 * k + 1]);
        w[k] = *reinterpret_cast<const unsigned*>(&h);
    }
    __stcs(reinterpret_cast<uint4*>(p), make_uint4(w[0], w[1], w[2], w[3]));
}

// in: M x H x W x C   out: M x 2H x 2W x CO (channels 0..C-1 of each pixel).  Per INPUT pixel (y, x) and 8 channels: the 3x3
// neighbourhood rows {max(y-1,0), y, min(y+1,H-1)} x columns {max(x-1,0), x, min(x+1,W-1)} and writes the 2x2
// output block (2y..2y+1, 2x..2x+1).  ATen (UpSampleBilinear2d.cu) takes, per output index o, the source index
// s = max(0.5*(o + 0.5) - 0.5, 0), the taps i1 = int(s) and i1 + (i1 < size-1), and the weights (1 - l, l) with
// l = s - i1:   o = 2i, i >= 1: taps (i-1, i), weights (0.25, 0.75);   o = 0: taps (0, min(1, size-1)), weights (1, 0);
//               o = 2i+1:      taps (i, min(i+1, size-1)), weights (0.75, 0.25);
// and sums  h0*(w0*a + w1*b) + h1*(w0*c + w1*d)  in fp32 -- the same expression, in the same order, here.
// A thread walks UPS_ROWS consecutive input rows of its column and carries the horizontal sums of the previous and
// current row along, so a row is loaded and summed horizontally once per thread instead of three times
// ((UPS_ROWS + 2) x 3 loads per UPS_ROWS x 4 stores instead of 9 per 4): the kernel is instruction-issue-bound
// (bf16 unpack + fp32 arithmetic), not bandwidth-bound, in its one-row form.
#ifndef SSM_UPS_ROWS
#define SSM_UPS_ROWS 4
#endif
#ifndef SSM_UPS_BLOCK
#define SSM_UPS_BLOCK 128
#endif
constexpr int UPS_ROWS = SSM_UPS_ROWS;
constexpr int UPS_BLOCK = SSM_UPS_BLOCK;

template <typename T>
__global__ void __launch_bounds__(UPS_BLOCK)
upsample2x_nhwc_kernel(const T* __restrict__ in, T* __restrict__ out, int H, int W, int C8, long long CO, long long total) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int HB = (H + UPS_ROWS - 1) / UPS_ROWS;                      // row blocks per image
    const int c8 = (int)(i % C8);
    long long r = i / C8;
    const int x = (int)(r % W); r /= W;
    const int yb = (int)(r % HB);
    const long long m = r / HB;
    const int C = C8 * 8;
    const int xs[3] = {max(x - 1, 0), x, min(x + 1, W - 1)};
    const bool x0 = x == 0;
    const float we0 = x0 ? 1.0f : 0.25f, we1 = x0 ? 0.0f : 0.75f;     // even output column
    const T* base = in + (m * H * (long long)W) * C + c8 * 8;
    // horizontal sums of one input row: e = even output column (taps x-1, x; at x = 0: taps 0, 1 with weights 1, 0),
    // o = odd output column (taps x, x+1)
    auto hsum = [&](int row, float (&e)[8], float (&o)[8]) {
        const T* rp = base + (long long)row * W * C;
        const Vec8 a = ld8<T>(rp + (long long)xs[0] * C), b = ld8<T>(rp + (long long)xs[1] * C), c = ld8<T>(rp + (long long)xs[2] * C);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            e[k] = we0 * a.v[k] + we1 * (x0 ? c.v[k] : b.v[k]);
            o[k] = 0.75f * b.v[k] + 0.25f * c.v[k];
        }
    };
    const int y_begin = yb * UPS_ROWS, y_end = min(y_begin + UPS_ROWS, H);
    float pe[8], po[8], ce[8], co[8], ne[8], no[8];                    // previous / current / next row
    hsum(max(y_begin - 1, 0), pe, po);
    hsum(y_begin, ce, co);
    const int W2 = 2 * W;
    for (int y = y_begin; y < y_end; ++y) {
        hsum(min(y + 1, H - 1), ne, no);
        const bool y0 = y == 0;
        const float he0 = y0 ? 1.0f : 0.25f, he1 = y0 ? 0.0f : 0.75f;  // even output row: rows (y-1, y), at y = 0: (0, 1)
        Vec8 o00, o01, o10, o11;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            o00.v[k] = he0 * pe[k] + he1 * (y0 ? ne[k] : ce[k]);
            o01.v[k] = he0 * po[k] + he1 * (y0 ? no[k] : co[k]);
            o10.v[k] = 0.75f * ce[k] + 0.25f * ne[k];                  // odd output row: rows (y, y+1)
            o11.v[k] = 0.75f * co[k] + 0.25f * no[k];
        }
        // CO = pixel stride of `out` in elements (>= C): the result may be a channel slice of a wider tensor, so that the
        // torch.cat in front of the upsampling (flow_computation.py:244-245, 253-254, ...) needs no pass of its own
        T* ob = out + ((m * 2 * H + 2 * y) * (long long)W2 + 2 * x) * CO + c8 * 8;
        st8<T>(ob, o00);
        st8<T>(ob + CO, o01);
        st8<T>(ob + (long long)W2 * CO, o10);
        st8<T>(ob + (long long)W2 * CO + CO, o11);
#pragma unroll
        for (int k = 0; k < 8; ++k) { pe[k] = ce[k]; po[k] = co[k]; ce[k] = ne[k]; co[k] = no[k]; }
    }
}

// out1 <- leaky_relu(y + bias[c], slope); out1 == y with pixel stride C is the in-place form.  bias: fp32, C values.
// out1 (pixel stride CO1 elements) and the optional second copy out2 (pixel stride CO2) may be channel slices of
// wider tensors: the activation is then written where the next torch.cat would have copied it
// (flow_computation.py:277 cat([conv11b_out, conv1b_out]) in front of fuse_conv), and that cat needs no pass.
template <typename T>
__global__ void __launch_bounds__(256)
bias_leaky_nhwc_kernel(const T* __restrict__ y, const float* __restrict__ bias, int C8, float slope, long long total,
                       T* __restrict__ out1, long long CO1, T* __restrict__ out2, long long CO2) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c8 = (int)(i % C8);
    const long long px = i / C8;
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8) + 1);
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    Vec8 r = ld8<T>(y + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        // the two-step form rounds the sum to the storage type before the activation (aten::add_ then leaky_relu_)
        const float s = storage_round<T>(r.v[k] + b[k]);
        r.v[k] = s > 0.0f ? s : s * slope;
    }
    st8<T>(out1 + px * CO1 + c8 * 8, r);
    if (out2) st8<T>(out2 + px * CO2 + c8 * 8, r);
}

// in: M x H x W x C (H, W even)   out: M x H/2 x W/2 x C, mean of the 2x2 window, summed in ATen's order
template <typename T>
__global__ void __launch_bounds__(256)
avgpool2_nhwc_kernel(const T* __restrict__ in, T* __restrict__ out, int Ho, int Wo, int C8, long long total) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c8 = (int)(i % C8);
    long long r = i / C8;
    const int x = (int)(r % Wo); r /= Wo;
    const int y = (int)(r % Ho);
    const long long m = r / Ho;
    const int C = C8 * 8, W = 2 * Wo;
    const T* p = in + ((m * 2 * Ho + 2 * y) * (long long)W + 2 * x) * C + c8 * 8;
    const Vec8 a = ld8<T>(p), b = ld8<T>(p + C), c = ld8<T>(p + (long long)W * C), d = ld8<T>(p + (long long)W * C + C);
    Vec8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = (((a.v[k] + b.v[k]) + c.v[k]) + d.v[k]) / 4.0f;
    st8<T>(out + i * 8, o);
}

// ---- backward of the three steps (training).  All are gathers: deterministic, no atomics. ------------------------
// d/d(in) of upsample2x: input pixel (y, x) is a tap of output rows 2y-1 .. 2y+2 with weights 0.25, 0.75, 0.75, 0.25
// (see the forward), except at the borders: row 0 takes output row 0 with weight 1 and has no row -1; row H-1 takes
// output row 2H-1 with weight 1 (both of its taps clamp to H-1) and has no row 2H.  Columns alike.
// g: M x 2H x 2W x CO (channel slice starting at `g`), gin: M x H x W x C dense.
template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_bwd_nhwc_kernel(const T* __restrict__ g, T* __restrict__ gin, int H, int W, int C8, long long CO, long long total) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c8 = (int)(i % C8);
    long long r = i / C8;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const long long m = r / H;
    float wr[4] = {0.25f, 0.75f, 0.75f, 0.25f}, wc[4] = {0.25f, 0.75f, 0.75f, 0.25f};
    if (y == 0) { wr[0] = 0.0f; wr[1] = 1.0f; }
    if (y == H - 1) { wr[3] = 0.0f; wr[2] = 1.0f; }
    if (x == 0) { wc[0] = 0.0f; wc[1] = 1.0f; }
    if (x == W - 1) { wc[3] = 0.0f; wc[2] = 1.0f; }
    const int W2 = 2 * W;
    const T* gb = g + (m * 2 * H * (long long)W2) * CO + c8 * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int oy = 2 * y - 1 + a;
        if (wr[a] == 0.0f) continue;
        float row[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (wc[b] == 0.0f) continue;
            const Vec8 v = ld8<T>(gb + ((long long)oy * W2 + (2 * x - 1 + b)) * CO);
#pragma unroll
            for (int k = 0; k < 8; ++k) row[k] = fmaf(wc[b], v.v[k], row[k]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(wr[a], row[k], acc[k]);
    }
    Vec8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = acc[k];
    st8<T>(gin + i * 8, o);
}

// gx = gy * (y > 0 ? 1 : slope), y = the activation OUTPUT (same sign as the pre-activation for slope > 0)
template <typename T>
__global__ void __launch_bounds__(256)
leaky_bwd_nhwc_kernel(const T* __restrict__ gy, const T* __restrict__ y, T* __restrict__ gx, float slope, long long total) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const Vec8 g = ld8<T>(gy + i * 8), v = ld8<T>(y + i * 8);
    Vec8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = v.v[k] > 0.0f ? g.v[k] : g.v[k] * slope;
    st8<T>(gx + i * 8, o);
}

// gin[2y+a, 2x+b] = gout[y, x] / 4
template <typename T>
__global__ void __launch_bounds__(256)
avgpool2_bwd_nhwc_kernel(const T* __restrict__ gout, T* __restrict__ gin, int Ho, int Wo, int C8, long long total) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c8 = (int)(i % C8);
    long long r = i / C8;
    const int x = (int)(r % Wo); r /= Wo;
    const int y = (int)(r % Ho);
    const long long m = r / Ho;
    const int C = C8 * 8, W = 2 * Wo;
    Vec8 v = ld8<T>(gout + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] = v.v[k] / 4.0f;
    T* p = gin + ((m * 2 * Ho + 2 * y) * (long long)W + 2 * x) * C + c8 * 8;
    st8<T>(p, v); st8<T>(p + C, v); st8<T>(p + (long long)W * C, v); st8<T>(p + (long long)W * C + C, v);
}

}  // namespace ssm
