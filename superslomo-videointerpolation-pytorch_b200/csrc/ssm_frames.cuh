// ssm_frames.cuh -- the steps either side of the synthesis path:
//   frames_from_u8   uint8 H x W x 3 images (cv2 BGR or RGB)  ->  normalised, zero- or mean-padded
//                    planar NCHW frames AND (optionally) the RGBx copy the gather kernels read,
//                    in one pass (reference: scripts/visualize_interpolation.py:61-88, 257-262;
//                    scripts/utils/dataloaders/augmentations.py:181-190, default_reader.py:266-271)
//   frames_to_u8     planar NCHW frames -> crop, de-normalise, uint8 H x W x 3
//                    (reference: scripts/evaluate_interpolation_results.py:143-163, 192-202;
//                    scripts/visualize_interpolation.py:221-232, 264-268)
// A uint8 input has 256 possible values per channel, so the normalisation is a 3 x 256 table the
// caller fills with the reference's own expression (on the device it wants to match bit for bit);
// the kernel is then a pure HBM-bound re-layout: 3 B/px read, 12 (+16) B/px written.
#pragma once
#include "ssm_kernels.cuh"

namespace ssm {

// one thread = 4 consecutive output pixels of one row: float4 / 4 x float4 stores
template <typename T>
__global__ void __launch_bounds__(256)
frames_from_u8_kernel(const unsigned char* __restrict__ src, long long src_frame_stride, int src_row_stride, int bgr,
                      int H_in, int W_in, int H, int W, int top, int left, const float* __restrict__ lut,
                      float pad0, float pad1, float pad2, View<T> planar, T* __restrict__ rgbx, long long total_quads) {
    __shared__ float s_lut[3 * 256];
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) s_lut[i] = lut[i];
    __syncthreads();
    const int qpr = W / 4;                                   // quads per row (W % 4 == 0, checked on the host)
    const float pad[3] = {pad0, pad1, pad2};
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total_quads;
         q += (long long)gridDim.x * blockDim.x) {
        const int xq = (int)(q % qpr);
        const long long r = q / qpr;
        const int y = (int)(r % H);
        const long long f = r / H;
        const int x0 = xq * 4;
        float v[3][4];
        const int sy = y - top;
        const bool row_in = (unsigned)sy < (unsigned)H_in;
        const unsigned char* row = src + f * src_frame_stride + (long long)(row_in ? sy : 0) * src_row_stride;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int sx = x0 + k - left;
            const bool in = row_in && (unsigned)sx < (unsigned)W_in;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                // output channel c is R,G,B; a BGR source stores it at byte 2-c
                v[c][k] = in ? s_lut[c * 256 + __ldg(row + (long long)sx * 3 + (bgr ? 2 - c : c))] : pad[c];
            }
        }
        const long long p = (long long)y * W + x0;
        if (planar.p) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                T* o = planar.p + f * planar.sb + c * planar.sc + p;
                if (sizeof(T) == 4) {
                    __stcs(reinterpret_cast<float4*>(o), make_float4(v[c][0], v[c][1], v[c][2], v[c][3]));
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) sts_(o + k, v[c][k]);
                }
            }
        }
        if (rgbx) {
            T* o = rgbx + (f * (long long)H * W + p) * 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (sizeof(T) == 4) {
                    *reinterpret_cast<float4*>(o + 4 * k) = make_float4(v[0][k], v[1][k], v[2][k], 0.0f);
                } else {
                    const unsigned lo = (unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(v[0][k])) |
                                        ((unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(v[1][k])) << 16);
                    const unsigned hi = (unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(v[2][k]));
                    *reinterpret_cast<uint2*>(o + 4 * k) = make_uint2(lo, hi);
                }
            }
        }
    }
}

// (x * std + mean) * scale with every product / sum rounded separately, as the three torch kernels
// of evaluate_interpolation_results.py:199-201 do, then the float -> uint8 conversion of
// numpy.astype(np.uint8) on x86-64 (truncate toward zero to int32, keep the low 8 bits: out-of-range
// values WRAP, which is what the reference writes) or, with saturate, a clamp to [0, 255] first.
__device__ __forceinline__ unsigned char to_u8(float x, float sd, float mu, float scale, int saturate) {
    float v = __fmul_rn(__fadd_rn(__fmul_rn(x, sd), mu), scale);
    if (saturate) v = fminf(fmaxf(v, 0.0f), 255.0f);
    if (!(fabsf(v) < 2147483648.0f)) return 0;             // cvttss2si "indefinite" value 0x80000000 -> low byte 0
    return (unsigned char)((int)v & 0xff);
}

// one thread = one output pixel (3 bytes); consecutive threads write consecutive bytes of a row
template <typename T>
__global__ void __launch_bounds__(256)
frames_to_u8_kernel(View<const T> planar, int H, int W, int top, int left, int H_out, int W_out,
                    float m0, float m1, float m2, float s0, float s1, float s2, float scale, int bgr, int saturate,
                    unsigned char* __restrict__ dst, long long dst_frame_stride, int dst_row_stride, long long total) {
    const float mu[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W_out);
        const long long r = i / W_out;
        const int y = (int)(r % H_out);
        const long long f = r / H_out;
        const T* s = planar.p + f * planar.sb + (long long)(y + top) * W + (x + left);
        unsigned char* o = dst + f * dst_frame_stride + (long long)y * dst_row_stride + (long long)x * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            o[bgr ? 2 - c : c] = to_u8(lds_(s + c * planar.sc), sd[c], mu[c], scale, saturate);
    }
}

}  // namespace ssm
