// ssm_q8.cuh -- the synthesis kernels for frames that arrive as 8-bit images (sm_100a).
//
// Real frames are uint8: the reference reads them with cv2 and normalises them on the fly
// (scripts/visualize_interpolation.py:61-88, 257-262).  The fp32 RGBx gathers of ssm_kernels.cuh are
// bound by the L1 data stage (10 wavefronts per 16-byte-per-lane request, 8 requests per pixel and
// timestep, profiles/r01p).  Here the gather source is a table of 2x2 ENTRIES of raw bytes:
//
//   entry(x0, y0) = { RGB(x0,y0), RGB(x0+1,y0), RGB(x0,y0+1), RGB(x0+1,y0+1), 4 spare bytes } = 16 B
//   for x0 in [-1, W-1], y0 in [-1, H-1]  ((H+1) x (W+1) entries per frame), zero bytes outside the frame,
//
// so ONE 16-byte request per bilinear sample brings all four taps of all three channels (4x fewer L1
// wavefronts, profiles/r02a_exp_gather2*), at the same 16 B/px footprint as the fp32 RGBx copy (a pair's
// two tables, 67 MB, stay in the 126 MB L2).  The normalisation (b/255 - mean)/std is affine in the byte,
// so it commutes with the interpolation:  sum_k w_k (a b_k + c) = a sum_k w_k b_k + c sum_{k in frame} w_k
// -- within 1e-6 of interpolating the normalised fp32 values as the reference does (zeros padding: taps
// outside the frame drop out of both sums).  The pass-through channels of compute_inputs (I0, I1 at the pixel,
// flow_interpolation.py:364-367) are read from the planar normalised frames (bit-exact with the reference's
// normalisation, ssm_frames_from_u8).
//
// One thread owns TWO horizontally adjacent pixels: every streaming access is 8 bytes per lane (a warp moves
// 256 contiguous bytes per plane row: 6.96 TB/s instead of 5.70 TB/s for the 16-plane store pattern of
// compute_inputs, profiles/r02b_exp_store.jsonl).  W must be even (padded sizes are multiples of 32).
#pragma once
#include "ssm_frames.cuh"

namespace ssm {

struct Norm3 { float a[3], c[3]; };       // normalised value of byte b in channel k: a[k] * b + c[k]

constexpr int Q8_TILE_W = 64;              // 32 lanes x 2 pixels
constexpr int Q8_TILE_H = 8;
constexpr int Q8_THREADS = 256;

__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// 1 / (1 + 2^(-x log2 e)): two MUFU + two FP32 ops; |error| < 1e-7 (not coordinate arithmetic)
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }

struct QTap {
    float wnw, wne, wsw, wse;   // corner weights (ATen: area of the opposite sub-rectangle)
    float wsum;                 // sum of the weights of the corners inside the frame
    int idx;                    // entry index (y0+1)*(W+1) + (x0+1); only dereferenced when ok
    bool ok;                    // at least one corner may be inside
};

template <int MODE>
__device__ __forceinline__ QTap make_qtap(int x, int y, float u, float v, const Geom& g) {
    float ix = sample_coord<MODE>((float)x, u, g.xnorm, g.xinv, g.xm1);
    float iy = sample_coord<MODE>((float)y, v, g.ynorm, g.yinv, g.ym1);
    ix = fminf(fmaxf(ix, -2.0f), (float)g.W + 1.0f);      // far outside (or NaN): no corner inside
    iy = fminf(fmaxf(iy, -2.0f), (float)g.H + 1.0f);
    const float fx = floorf(ix), fy = floorf(iy);
    // ix - fx is exact; 1 - (ix - fx) equals ATen's (fx + 1) - ix except for |ix| < 1, where it may differ by one
    // rounding (6e-8) -- one instruction less per axis
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
    const int x0 = (int)fx, y0 = (int)fy;
    QTap t;
    t.wnw = wx0 * wy0; t.wne = wx1 * wy0; t.wsw = wx0 * wy1; t.wse = wx1 * wy1;
    const float sx = ((unsigned)x0 < (unsigned)g.W ? wx0 : 0.0f) + ((unsigned)(x0 + 1) < (unsigned)g.W ? wx1 : 0.0f);
    const float sy = ((unsigned)y0 < (unsigned)g.H ? wy0 : 0.0f) + ((unsigned)(y0 + 1) < (unsigned)g.H ? wy1 : 0.0f);
    t.wsum = sx * sy;
    t.ok = (unsigned)(x0 + 1) <= (unsigned)g.W && (unsigned)(y0 + 1) <= (unsigned)g.H;
    t.idx = (y0 + 1) * (g.W + 1) + (x0 + 1);
    return t;
}

__device__ __forceinline__ uint4 load_entry(const uint4* __restrict__ table, const QTap& t) {
    // unsigned 32-bit entry index: the address is one IMAD.WIDE.U32 (base + idx * 16)
    return t.ok ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(table) + (size_t)(unsigned)t.idx * 16u))
                : make_uint4(0u, 0u, 0u, 0u);
}

// byte k of w as a float: PRMT builds the bit pattern of 2^23 + b, one FADD removes the 2^23 (exact)
__device__ __forceinline__ float byte_f(unsigned w, int k) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 + k)) - 8388608.0f;
}

// normalised bilinear sample of the three channels from one entry
__device__ __forceinline__ void q8_sample(const uint4& q, const QTap& t, const Norm3& nm, float (&out)[3]) {
    const unsigned w[3] = {q.x, q.y, q.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // bytes 0-2 nw, 3-5 ne, 6-8 sw, 9-11 se
        float s = byte_f(w[c >> 2], c & 3) * t.wnw;
        s = fmaf(byte_f(w[(3 + c) >> 2], (3 + c) & 3), t.wne, s);
        s = fmaf(byte_f(w[(6 + c) >> 2], (6 + c) & 3), t.wsw, s);
        s = fmaf(byte_f(w[(9 + c) >> 2], (9 + c) & 3), t.wse, s);
        out[c] = fmaf(nm.a[c], s, nm.c[c] * t.wsum);
    }
}

struct Q8Idx { int b, x, y; bool valid; };
__device__ __forceinline__ Q8Idx q8_index(int H, int W) {
    const int tiles_x = (W + Q8_TILE_W - 1) / Q8_TILE_W, tiles_y = (H + Q8_TILE_H - 1) / Q8_TILE_H;
    const int tpp = tiles_x * tiles_y;
    Q8Idx t;
    t.b = blockIdx.x / tpp;
    const int r = blockIdx.x - t.b * tpp, ty = r / tiles_x, tx = r - ty * tiles_x;
    t.x = tx * Q8_TILE_W + 2 * (threadIdx.x & 31);
    t.y = ty * Q8_TILE_H + (threadIdx.x >> 5);
    t.valid = t.x < W && t.y < H;        // W even: x + 1 < W as well
    return t;
}

__device__ __forceinline__ float2 lds2(const float* p) { return __ldcs(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 lds2(const __nv_bfloat16* p) {
    const unsigned w = __ldcs(reinterpret_cast<const unsigned*>(p));
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ void sts2(float* p, float a, float b) { __stcs(reinterpret_cast<float2*>(p), make_float2(a, b)); }
__device__ __forceinline__ void sts2(__nv_bfloat16* p, float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);          // .x = low half = the first pixel
    __stcs(reinterpret_cast<unsigned*>(p), *reinterpret_cast<const unsigned*>(&h));
}

// =============================================================================================
// entry tables from uint8 images: F images (H_in x W_in x 3, RGB or BGR bytes) placed at (top, left) of an
// H x W frame whose other pixels are byte 0 (visualize_interpolation.py:76-87 pads the raw image with 0 and
// normalises afterwards, :137)  ->  F x (H+1) x (W+1) entries.  One thread builds 4 consecutive entries of a row.
// =============================================================================================
__global__ void __launch_bounds__(256)
quads_from_u8_kernel(const unsigned char* __restrict__ src, long long src_frame_stride, int src_row_stride, int bgr,
                     int H_in, int W_in, int H, int W, int top, int left, uint4* __restrict__ quads, long long total_groups) {
    const int gpr = (W + 1 + 3) / 4;                       // groups of 4 entries per entry row
    for (long long gidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; gidx < total_groups;
         gidx += (long long)gridDim.x * blockDim.x) {
        const int gx = (int)(gidx % gpr);
        const long long r = gidx / gpr;
        const int ey = (int)(r % (H + 1));
        const long long f = r / (H + 1);
        const int ex0 = gx * 4;
        // texel columns ex0-1 .. ex0+3, rows ey-1, ey (frame coordinates); 0 outside the frame or the source
        unsigned tex[2][5];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const int sy = ey - 1 + dy - top;
            const bool row_in = (unsigned)(ey - 1 + dy) < (unsigned)H && (unsigned)sy < (unsigned)H_in;
            const unsigned char* row = src + f * src_frame_stride + (long long)(row_in ? sy : 0) * src_row_stride;
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) {
                const int fxp = ex0 - 1 + dx, sx = fxp - left;
                unsigned v = 0u;
                if (row_in && (unsigned)fxp < (unsigned)W && (unsigned)sx < (unsigned)W_in) {
                    const unsigned char* px = row + (long long)sx * 3;
                    const unsigned b0 = __ldg(px), b1 = __ldg(px + 1), b2 = __ldg(px + 2);
                    v = bgr ? (b2 | (b1 << 8) | (b0 << 16)) : (b0 | (b1 << 8) | (b2 << 16));    // R | G<<8 | B<<16
                }
                tex[dy][dx] = v;
            }
        }
        uint4* o = quads + (f * (H + 1) + ey) * (long long)(W + 1) + ex0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (ex0 + k > W) break;
            const unsigned t00 = tex[0][k], t10 = tex[0][k + 1], t01 = tex[1][k], t11 = tex[1][k + 1];
            o[k] = make_uint4(t00 | (t10 << 24), (t10 >> 8) | (t01 << 16), (t01 >> 16) | (t11 << 8), 0u);
        }
    }
}

// =============================================================================================
// a2: compute_inputs from entry tables     reference scripts/models/flow_interpolation.py:338-372
//     (batched over N timesteps, as flow_pack_fwd_kernel).  TO / NHWC: see flow_pack_fwd_kernel.
// =============================================================================================
#ifndef SSM_Q8_PACK_MIN_BLOCKS
#define SSM_Q8_PACK_MIN_BLOCKS 3
#endif
// measured (profiles/r02f_q8_timing_*.json): compute_output_image gains from 4 CTAs per SM (64 registers, 2.40 -> 2.25 ms)
// except with the fused uint8 output, which then spills (2.53 -> 2.94 ms); compute_inputs is best at 3 (80 registers)
#ifndef SSM_Q8_FUSE_MIN_BLOCKS
#define SSM_Q8_FUSE_MIN_BLOCKS 4
#endif
#ifndef SSM_Q8_FUSE_U8_MIN_BLOCKS
#define SSM_Q8_FUSE_U8_MIN_BLOCKS 3
#endif
// T: storage type of img6 / flow4 (and of out16 in the planar layout): fp32, or bf16 with fp32 arithmetic
template <typename T, int MODE, typename TO, bool NHWC>
__global__ void __launch_bounds__(Q8_THREADS, SSM_Q8_PACK_MIN_BLOCKS)
flow_pack_fwd_q8_kernel(View<const T> img6, const uint4* __restrict__ quads, View<const T> flow4,
                        const float* __restrict__ tv, View<TO> out16, int N, Geom g, Norm3 nm) {
    const Q8Idx ti = q8_index(g.H, g.W);
    if (!ti.valid) return;
    const int p = ti.y * g.W + ti.x;
    const long long epf = (long long)(g.H + 1) * (g.W + 1);            // entries per frame
    const uint4* __restrict__ tab0 = quads + (long long)ti.b * 2 * epf;
    const uint4* __restrict__ tab1 = tab0 + epf;
    const T* F = flow4.p + ti.b * flow4.sb + p;
    const int fsc = (int)flow4.sc, isc = (int)img6.sc;
    const float2 f01x = lds2(F), f01y = lds2(F + fsc), f10x = lds2(F + 2 * fsc), f10y = lds2(F + 3 * fsc);
    const T* I = img6.p + ti.b * img6.sb + p;
    float2 c0[3], c1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { c0[c] = lds2(I + c * isc); c1[c] = lds2(I + (3 + c) * isc); }
    const float* tp = tv + ti.b * N;
    TO* __restrict__ O = out16.p + ti.b * out16.sb + (NHWC ? (long long)p * 16 : (long long)p);
    const int osc = (int)out16.sc;
    const float fa[2][4] = {{f01x.x, f01y.x, f10x.x, f10y.x}, {f01x.y, f01y.y, f10x.y, f10y.y}};
    for (int n = 0; n < N; ++n, O += out16.sn) {
        const Coef k = make_coef(__ldg(tp + n));
        float e1x[2], e1y[2], e0x[2], e0y[2];
        QTap t1[2], t0[2];
        uint4 q1[2], q0[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            e0x[i] = storage_round<T>(est_t0(k, fa[i][0], fa[i][2])); e0y[i] = storage_round<T>(est_t0(k, fa[i][1], fa[i][3]));   // F_t0  :353
            e1x[i] = storage_round<T>(est_t1(k, fa[i][0], fa[i][2])); e1y[i] = storage_round<T>(est_t1(k, fa[i][1], fa[i][3]));   // F_t1  :356
            t1[i] = make_qtap<MODE>(ti.x + i, ti.y, e1x[i], e1y[i], g);                       // warp(img_1, F_t1) :361
            t0[i] = make_qtap<MODE>(ti.x + i, ti.y, e0x[i], e0y[i], g);                       // warp(img_0, F_t0) :362
            q1[i] = load_entry(tab1, t1[i]);
            q0[i] = load_entry(tab0, t0[i]);
        }
        float w1[2][3], w0[2][3];
#pragma unroll
        for (int i = 0; i < 2; ++i) { q8_sample(q1[i], t1[i], nm, w1[i]); q8_sample(q0[i], t0[i], nm, w0[i]); }
        if (NHWC) {                                                                           // :364-367, channels-last
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float o[16];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    o[c] = i ? c1[c].y : c1[c].x; o[3 + c] = w1[i][c]; o[10 + c] = w0[i][c]; o[13 + c] = i ? c0[c].y : c0[c].x;
                }
                o[6] = e1x[i]; o[7] = e1y[i]; o[8] = e0x[i]; o[9] = e0y[i];
                store16_nhwc<TO>(O + 16 * i, o);
            }
            continue;
        }
        if constexpr (!NHWC) {
            TO* Of = O;
#pragma unroll
            for (int c = 0; c < 3; ++c) {                                                     // :364-367
                sts2(Of + (0 + c) * osc, c1[c].x, c1[c].y);
                sts2(Of + (3 + c) * osc, w1[0][c], w1[1][c]);
                sts2(Of + (10 + c) * osc, w0[0][c], w0[1][c]);
                sts2(Of + (13 + c) * osc, c0[c].x, c0[c].y);
            }
            sts2(Of + 6 * osc, e1x[0], e1x[1]); sts2(Of + 7 * osc, e1y[0], e1y[1]);
            sts2(Of + 8 * osc, e0x[0], e0x[1]); sts2(Of + 9 * osc, e0y[0], e0y[1]);
        }
    }
}

// =============================================================================================
// a3 + a4: extract_outputs + compute_output_image from entry tables   flow_interpolation.py:374-429
//     estimated flows recomputed from the stage-1 flows (as fuse_fwd_kernel<RECOMP>); TY = storage type of the
//     U-Net output.  OUT_U8: the fused frame is de-normalised and written as uint8 H_out x W_out x 3 images
//     (crop at (top, left)) with the arithmetic of frames_to_u8_kernel, instead of as fp32 planes.
// =============================================================================================
struct U8Out {
    unsigned char* dst; long long frame_stride; int row_stride;    // frame index = b * N + n
    int top, left, H_out, W_out, bgr, saturate;
    float mean[3], std[3], scale;
};

template <typename T, int MODE, typename TY, bool OUT_U8>
__global__ void __launch_bounds__(Q8_THREADS, OUT_U8 ? SSM_Q8_FUSE_U8_MIN_BLOCKS : SSM_Q8_FUSE_MIN_BLOCKS)
fuse_fwd_q8_kernel(const uint4* __restrict__ quads, View<const T> flow4, View<const TY> out5,
                   const float* __restrict__ tv, View<T> out3, U8Out u8, int N, Geom g, Norm3 nm) {
    const Q8Idx ti = q8_index(g.H, g.W);
    if (!ti.valid) return;
    const int p = ti.y * g.W + ti.x;
    const long long epf = (long long)(g.H + 1) * (g.W + 1);
    const uint4* __restrict__ tab0 = quads + (long long)ti.b * 2 * epf;
    const uint4* __restrict__ tab1 = tab0 + epf;
    const float* tp = tv + ti.b * N;
    const T* F = flow4.p + ti.b * flow4.sb + p;
    const int fsc = (int)flow4.sc, ysc = (int)out5.sc, osc = (int)out3.sc;
    const float2 f01x = lds2(F), f01y = lds2(F + fsc), f10x = lds2(F + 2 * fsc), f10y = lds2(F + 3 * fsc);
    const float fa[2][4] = {{f01x.x, f01y.x, f10x.x, f10y.x}, {f01x.y, f01y.y, f10x.y, f10y.y}};
    const TY* __restrict__ Y = out5.p + ti.b * out5.sb + p;
    T* __restrict__ O = OUT_U8 ? nullptr : out3.p + ti.b * out3.sb + p;
    float2 ys[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) ys[c] = lds2(Y + c * ysc);
    for (int n = 0; n < N; ++n) {
        const float tt = __ldg(tp + n);
        const Coef k = make_coef(tt);
        const float omt = k.omt;
        const float ya[2][5] = {{ys[0].x, ys[1].x, ys[2].x, ys[3].x, ys[4].x}, {ys[0].y, ys[1].y, ys[2].y, ys[3].y, ys[4].y}};
        if (n + 1 < N) {             // streaming loads of the next timestep, in flight during the gathers
            Y += out5.sn;
#pragma unroll
            for (int c = 0; c < 5; ++c) ys[c] = lds2(Y + c * ysc);
        }
        QTap t0[2], t1[2];
        uint4 q0[2], q1[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float f1x = __fadd_rn(storage_round<T>(est_t1(k, fa[i][0], fa[i][2])), ya[i][1]);          // :412
            const float f1y = __fadd_rn(storage_round<T>(est_t1(k, fa[i][1], fa[i][3])), ya[i][2]);
            const float f0x = __fadd_rn(storage_round<T>(est_t0(k, fa[i][0], fa[i][2])), ya[i][3]);          // :413
            const float f0y = __fadd_rn(storage_round<T>(est_t0(k, fa[i][1], fa[i][3])), ya[i][4]);
            t0[i] = make_qtap<MODE>(ti.x + i, ti.y, f0x, f0y, g);                          // :416
            t1[i] = make_qtap<MODE>(ti.x + i, ti.y, f1x, f1y, g);                          // :418
            q0[i] = load_entry(tab0, t0[i]);
            q1[i] = load_entry(tab1, t1[i]);
        }
        float res[2][3];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float v1 = sigmoid_fast(ya[i][0]);                                        // :386-388
            const float v0 = 1.0f - v1;                                                     // :390
            const float a0 = omt * v0, a1 = tt * v1;
            const float rz = rcp_approx(a0 + a1);                                           // 1/Z  :425
            float s0[3], s1[3];
            q8_sample(q0[i], t0[i], nm, s0);
            q8_sample(q1[i], t1[i], nm, s1);
#pragma unroll
            for (int c = 0; c < 3; ++c) res[i][c] = (a0 * s0[c] + a1 * s1[c]) * rz;        // :420-427
        }
        if (OUT_U8) {
            const int oy = ti.y - u8.top;
            if ((unsigned)oy < (unsigned)u8.H_out) {
                unsigned char* row = u8.dst + (long long)(ti.b * N + n) * u8.frame_stride + (long long)oy * u8.row_stride;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int ox = ti.x + i - u8.left;
                    if ((unsigned)ox < (unsigned)u8.W_out) {
                        unsigned char* o = row + (long long)ox * 3;
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            o[u8.bgr ? 2 - c : c] = to_u8(res[i][c], u8.std[c], u8.mean[c], u8.scale, u8.saturate);
                    }
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) sts2(O + c * osc, res[0][c], res[1][c]);
            O += out3.sn;
        }
    }
}

}  // namespace ssm
