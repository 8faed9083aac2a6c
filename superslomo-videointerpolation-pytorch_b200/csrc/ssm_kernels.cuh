// ssm_kernels.cuh -- the synthesis kernels (sm_100a).  One thread owns one pixel of one frame
// pair and loops over the N timesteps of that pair, so everything that does not depend on t
// (flows F01/F10, the centre pixels of I0/I1, the accumulators of the flow gradient) lives in
// registers across the timestep loop, and the gather footprints of consecutive timesteps overlap
// in L1.  A CTA is a 32 x 8 pixel tile: a warp is 32 consecutive pixels of one row, so every
// streaming load/store is one full 128-byte line per warp and the 32 gather addresses of a tap fall
// into one or two lines when the flow is smooth; the 8 rows share the north/south tap lines in L1.
// CTAs are numbered pair-major so that the two frames of a pair (50 MB planar / 67 MB packed at
// 1088x1920 fp32) stay resident in the 126 MB L2 while all tiles and timesteps of that pair are
// processed.
//
// Every kernel that gathers from the frames is templated on PACKED: gather from the RGBx copy made
// by pack_frames_kernel (one 16-byte request per tap) or from the planar frames (three 4-byte
// requests per tap).  See ssm_device.cuh for why.
//
// All of it is HBM-bound element-wise + gather work: no tensor cores.
#pragma once
#include "ssm_device.cuh"

namespace ssm {

constexpr int TILE_W = 32;
constexpr int TILE_H = 8;
constexpr int TILE_THREADS = TILE_W * TILE_H;

struct TileIdx { int b, x, y; bool valid; };

__device__ __forceinline__ TileIdx tile_index(int H, int W) {
    // blockIdx.x enumerates (pair, tile_y, tile_x) with tile_x fastest
    int tiles_x = (W + TILE_W - 1) / TILE_W;
    int tiles_y = (H + TILE_H - 1) / TILE_H;
    int tpp = tiles_x * tiles_y;
    int b = blockIdx.x / tpp;
    int r = blockIdx.x - b * tpp;
    int ty = r / tiles_x, tx = r - ty * tiles_x;
    TileIdx t;
    t.b = b;
    t.x = tx * TILE_W + (threadIdx.x & (TILE_W - 1));
    t.y = ty * TILE_H + (threadIdx.x / TILE_W);
    t.valid = t.x < W && t.y < H;
    return t;
}

// the two frames of pair b, as the gather source of the PACKED / planar variants
template <typename T, bool PACKED> struct Frames {
    const T* f0;
    const T* f1;
    long long sc;
    __device__ __forceinline__ Frames(const View<const T>& img6, const T* __restrict__ packed, int b, long long npx) {
        if (PACKED) {
            f0 = packed + (long long)b * 8 * npx;
            f1 = f0 + 4 * npx;
            sc = 0;
        } else {
            f0 = img6.p + b * img6.sb;
            f1 = f0 + 3 * img6.sc;
            sc = img6.sc;
        }
    }
};

// =============================================================================================
// frame re-layout: B x 6 x H x W planar  ->  B x 2 x H x W x 4 (RGBx), once per batch of pairs
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(TILE_THREADS)
pack_frames_kernel(View<const T> img6, T* __restrict__ packed, Geom g) {
    TileIdx ti = tile_index(g.H, g.W);
    if (!ti.valid) return;
    const int p = ti.y * g.W + ti.x;
    const long long npx = (long long)g.H * g.W;
    const T* I = img6.p + ti.b * img6.sb + p;
    T* o = packed + ((long long)ti.b * 8 * npx) + (long long)p * 4;
#pragma unroll
    for (int f = 0; f < 2; ++f) {
        const float r = lds_(I + (3 * f + 0) * img6.sc), gg = lds_(I + (3 * f + 1) * img6.sc), bb = lds_(I + (3 * f + 2) * img6.sc);
        if (sizeof(T) == 4) {
            *reinterpret_cast<float4*>(o + f * 4 * npx) = make_float4(r, gg, bb, 0.0f);
        } else {
            const unsigned lo = (unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(r)) |
                                ((unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(gg)) << 16);
            const unsigned hi = (unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(bb));
            *reinterpret_cast<uint2*>(o + f * 4 * npx) = make_uint2(lo, hi);
        }
    }
}

// =============================================================================================
// a1: warp forward          reference scripts/models/layers.py:73-120   (any channel count: planar)
// =============================================================================================
template <typename T, int MODE>
__global__ void __launch_bounds__(TILE_THREADS)
warp_fwd_kernel(View<const T> img, View<const T> flow, View<T> out, int C, Geom g) {
    TileIdx ti = tile_index(g.H, g.W);
    if (!ti.valid) return;
    const int p = ti.y * g.W + ti.x;
    const T* fl = flow.p + ti.b * flow.sb + p;
    float u = lds_(fl), v = lds_(fl + flow.sc);
    Taps t = make_taps<MODE>(ti.x, ti.y, u, v, g);
    const T* ip = img.p + ti.b * img.sb;
    T* op = out.p + ti.b * out.sb + p;
    for (int c = 0; c < C; ++c) {
        Quad q = gather_quad(ip + c * img.sc, t, g.W);
        sts_(op + c * out.sc, bilerp(q, t));
    }
}

// warp backward, gather part: gradient w.r.t. the flow.  When hdr is given it also records
// max |grad_out| for the deterministic image-gradient pass (ssm_scatter.cuh), which re-reads
// grad_out and flow directly.
template <typename T, int MODE>
__global__ void __launch_bounds__(TILE_THREADS)
warp_bwd_flow_kernel(View<const T> gout, View<const T> img, View<const T> flow, View<T> gflow, int C, Geom g,
                     ScatterHdr* hdr) {
    TileIdx ti = tile_index(g.H, g.W);
    float amax = 0.0f;
    if (ti.valid) {
        const int p = ti.y * g.W + ti.x;
        const T* fl = flow.p + ti.b * flow.sb + p;
        float u = lds_(fl), v = lds_(fl + flow.sc);
        Taps t = make_taps<MODE>(ti.x, ti.y, u, v, g);
        const T* ip = img.p + ti.b * img.sb;
        const T* gp = gout.p + ti.b * gout.sb + p;
        float gix = 0.0f, giy = 0.0f;
        for (int c = 0; c < C; ++c) {
            Quad q = gather_quad(ip + c * img.sc, t, g.W);
            float gc = lds_(gp + c * gout.sc);
            amax = fmaxf(amax, fabsf(gc));
            bilerp_grad(q, t, gc, gix, giy);
        }
        T* o = gflow.p + ti.b * gflow.sb + p;
        sts_(o, coord_grad_to_flow<MODE>(gix, g.xgrad, g.xnorm, g.xinv));
        sts_(o + gflow.sc, coord_grad_to_flow<MODE>(giy, g.ygrad, g.ynorm, g.yinv));
    }
    if (hdr) record_absmax(&hdr->absmax_bits, amax);
}

// =============================================================================================
// a2: compute_inputs forward    reference scripts/models/flow_interpolation.py:338-372
//     batched over N timesteps (absorbs the loop + torch.stack of superslomo_r.py:167-179 and the
//     three torch.cat of :364-367): reads 10 channels once, writes 16 channels per timestep.
// =============================================================================================
template <typename T, int MODE, bool PACKED>
__global__ void __launch_bounds__(TILE_THREADS)
flow_pack_fwd_kernel(View<const T> img6, const T* __restrict__ packed, View<const T> flow4,
                     const float* __restrict__ tv, View<T> out16, int N, Geom g) {
    TileIdx ti = tile_index(g.H, g.W);
    if (!ti.valid) return;
    const int p = ti.y * g.W + ti.x;
    const long long npx = (long long)g.H * g.W;
    const Frames<T, PACKED> fr(img6, packed, ti.b, npx);
    const T* F = flow4.p + ti.b * flow4.sb + p;
    const float f01x = lds_(F), f01y = lds_(F + flow4.sc);
    const float f10x = lds_(F + 2 * flow4.sc), f10y = lds_(F + 3 * flow4.sc);
    float c0[3], c1[3];
    if (PACKED) {
        load_px(fr.f0 + (long long)p * 4, c0);
        load_px(fr.f1 + (long long)p * 4, c1);
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            c0[c] = ldg_(fr.f0 + c * fr.sc + p);
            c1[c] = ldg_(fr.f1 + c * fr.sc + p);
        }
    }
    const float* tp = tv + ti.b * N;
    T* __restrict__ O = out16.p + ti.b * out16.sb + p;
    const int osc = (int)out16.sc;   // fits 32 bits (checked on the host): one IMAD.WIDE per address
    for (int n = 0; n < N; ++n, O += out16.sn) {
        const Coef k = make_coef(__ldg(tp + n));
        const float e0x = est_t0(k, f01x, f10x), e0y = est_t0(k, f01y, f10y);   // F_t0  :353
        const float e1x = est_t1(k, f01x, f10x), e1y = est_t1(k, f01y, f10y);   // F_t1  :356
        const Taps t1 = make_taps<MODE>(ti.x, ti.y, e1x, e1y, g);               // warp(img_1, F_t1) :361
        const Taps t0 = make_taps<MODE>(ti.x, ti.y, e0x, e0y, g);               // warp(img_0, F_t0) :362
        Quad q1[3], q0[3];
        gather3<T, PACKED>(fr.f1, fr.sc, t1, g.W, q1);
        gather3<T, PACKED>(fr.f0, fr.sc, t0, g.W, q0);
#pragma unroll
        for (int c = 0; c < 3; ++c) {                                           // :364-367
            sts_(O + (0 + c) * osc, c1[c]);
            sts_(O + (3 + c) * osc, bilerp(q1[c], t1));
            sts_(O + (10 + c) * osc, bilerp(q0[c], t0));
            sts_(O + (13 + c) * osc, c0[c]);
        }
        sts_(O + 6 * osc, e1x); sts_(O + 7 * osc, e1y);
        sts_(O + 8 * osc, e0x); sts_(O + 9 * osc, e0y);
    }
}

// a2 backward, gather part: gradient w.r.t. flow_pred_tensor, summed over the N timesteps in
// registers (deterministic).  With IMG_GRAD the direct image gradients (channels 0:3 and 13:16 of
// grad16, summed over timesteps) go to an fp32 staging buffer and max |grad16[:, 3:6|10:13]| is
// recorded; the warped-image part is added by the scatter pass (ssm_scatter.cuh).
template <typename T, int MODE, bool PACKED, bool IMG_GRAD>
__global__ void __launch_bounds__(TILE_THREADS)
flow_pack_bwd_kernel(View<const T> g16, View<const T> img6, const T* __restrict__ packed, View<const T> flow4,
                     const float* __restrict__ tv, View<T> gflow4, float* __restrict__ gimg_direct,
                     ScatterHdr* hdr, int N, Geom g) {
    TileIdx ti = tile_index(g.H, g.W);
    float amax = 0.0f;
    if (ti.valid) {
        const bool want_flow = gflow4.p != nullptr;
        const int p = ti.y * g.W + ti.x;
        const long long npx = (long long)g.H * g.W;
        const Frames<T, PACKED> fr(img6, packed, ti.b, npx);
        const T* F = flow4.p + ti.b * flow4.sb + p;
        const float f01x = lds_(F), f01y = lds_(F + flow4.sc);
        const float f10x = lds_(F + 2 * flow4.sc), f10y = lds_(F + 3 * flow4.sc);
        float d01x = 0, d01y = 0, d10x = 0, d10y = 0;
        float di0[3] = {0, 0, 0}, di1[3] = {0, 0, 0};
        const float* tp = tv + ti.b * N;
        const T* G = g16.p + ti.b * g16.sb + p;
        for (int n = 0; n < N; ++n, G += g16.sn) {
            const Coef k = make_coef(__ldg(tp + n));
            float gw1[3], gw0[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                gw1[c] = lds_(G + (3 + c) * g16.sc);
                gw0[c] = lds_(G + (10 + c) * g16.sc);
                if (IMG_GRAD) {
                    amax = fmaxf(amax, fmaxf(fabsf(gw1[c]), fabsf(gw0[c])));
                    di1[c] += lds_(G + (0 + c) * g16.sc);
                    di0[c] += lds_(G + (13 + c) * g16.sc);
                }
            }
            if (!want_flow) continue;
            const float e0x = est_t0(k, f01x, f10x), e0y = est_t0(k, f01y, f10y);
            const float e1x = est_t1(k, f01x, f10x), e1y = est_t1(k, f01y, f10y);
            const Taps t1 = make_taps<MODE>(ti.x, ti.y, e1x, e1y, g);
            const Taps t0 = make_taps<MODE>(ti.x, ti.y, e0x, e0y, g);
            Quad q1[3], q0[3];
            gather3<T, PACKED>(fr.f1, fr.sc, t1, g.W, q1);
            gather3<T, PACKED>(fr.f0, fr.sc, t0, g.W, q0);
            float g1x = 0, g1y = 0, g0x = 0, g0y = 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                bilerp_grad(q1[c], t1, gw1[c], g1x, g1y);
                bilerp_grad(q0[c], t0, gw0[c], g0x, g0y);
            }
            const float de1x = lds_(G + 6 * g16.sc) + coord_grad_to_flow<MODE>(g1x, g.xgrad, g.xnorm, g.xinv);
            const float de1y = lds_(G + 7 * g16.sc) + coord_grad_to_flow<MODE>(g1y, g.ygrad, g.ynorm, g.yinv);
            const float de0x = lds_(G + 8 * g16.sc) + coord_grad_to_flow<MODE>(g0x, g.xgrad, g.xnorm, g.xinv);
            const float de0y = lds_(G + 9 * g16.sc) + coord_grad_to_flow<MODE>(g0y, g.ygrad, g.ynorm, g.yinv);
            d01x += k.c00 * de0x + k.c10 * de1x;
            d01y += k.c00 * de0y + k.c10 * de1y;
            d10x += k.c01 * de0x - k.c11 * de1x;
            d10y += k.c01 * de0y - k.c11 * de1y;
        }
        if (want_flow) {
            T* o = gflow4.p + ti.b * gflow4.sb + p;
            sts_(o, d01x); sts_(o + gflow4.sc, d01y);
            sts_(o + 2 * gflow4.sc, d10x); sts_(o + 3 * gflow4.sc, d10y);
        }
        if (IMG_GRAD) {
            // dense fp32 B x 6 x H x W; the finalise pass adds the scattered part
            float* d = gimg_direct + (long long)ti.b * 6 * npx + p;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                d[c * npx] = di0[c];
                d[(3 + c) * npx] = di1[c];
            }
        }
    }
    if (IMG_GRAD) record_absmax(&hdr->absmax_bits, amax);
}

// =============================================================================================
// a3 + a4: extract_outputs + compute_output_image   flow_interpolation.py:374-429
//     batched over N timesteps (the loop of superslomo_r.py:215-238)
// =============================================================================================
template <typename T, int MODE, bool PACKED>
#ifndef SSM_FUSE_MIN_BLOCKS
#define SSM_FUSE_MIN_BLOCKS 4
#endif
#ifndef SSM_FUSE_PREFETCH
#define SSM_FUSE_PREFETCH 1
#endif
__global__ void __launch_bounds__(TILE_THREADS, SSM_FUSE_MIN_BLOCKS)
fuse_fwd_kernel(View<const T> img6, const T* __restrict__ packed, View<const T> flows4, View<const T> out5,
                const float* __restrict__ tv, View<T> out3, int N, Geom g) {
    TileIdx ti = tile_index(g.H, g.W);
    if (!ti.valid) return;
    const int p = ti.y * g.W + ti.x;
    const long long npx = (long long)g.H * g.W;
    const Frames<T, PACKED> fr(img6, packed, ti.b, npx);
    const float* tp = tv + ti.b * N;
    const T* __restrict__ X = flows4.p + ti.b * flows4.sb + p;
    const T* __restrict__ Y = out5.p + ti.b * out5.sb + p;
    T* __restrict__ O = out3.p + ti.b * out3.sb + p;
    // channel strides fit 32 bits (checked on the host): one IMAD.WIDE per address
    const int xsc = (int)flows4.sc, ysc = (int)out5.sc, osc = (int)out3.sc;
    // The 9 streaming loads of timestep n+1 are issued before the gathers of timestep n, so that two
    // dependent long-latency phases (HBM stream, then L2/L1 gather) overlap across iterations.
    float xs[4], ys[5];
    if (SSM_FUSE_PREFETCH) {
#pragma unroll
        for (int k = 0; k < 4; ++k) xs[k] = lds_(X + k * xsc);
#pragma unroll
        for (int k = 0; k < 5; ++k) ys[k] = lds_(Y + k * ysc);
    }
    for (int n = 0; n < N; ++n, O += out3.sn) {
        const float tt = __ldg(tp + n);
        const float omt = __fsub_rn(1.0f, tt);
        if (!SSM_FUSE_PREFETCH) {
#pragma unroll
            for (int k = 0; k < 4; ++k) xs[k] = lds_(X + k * xsc);
#pragma unroll
            for (int k = 0; k < 5; ++k) ys[k] = lds_(Y + k * ysc);
            X += flows4.sn; Y += out5.sn;
        }
        const float logit = ys[0];
        const float f1x = __fadd_rn(xs[0], ys[1]);                                   // :412
        const float f1y = __fadd_rn(xs[1], ys[2]);
        const float f0x = __fadd_rn(xs[2], ys[3]);                                   // :413
        const float f0y = __fadd_rn(xs[3], ys[4]);
        if (SSM_FUSE_PREFETCH && n + 1 < N) {
            X += flows4.sn; Y += out5.sn;
#pragma unroll
            for (int k = 0; k < 4; ++k) xs[k] = lds_(X + k * xsc);
#pragma unroll
            for (int k = 0; k < 5; ++k) ys[k] = lds_(Y + k * ysc);
        }
        const float v1 = sigmoid_(logit);                                            // :386-388
        const float v0 = 1.0f - v1;                                                  // :390
        const Taps t0 = make_taps<MODE>(ti.x, ti.y, f0x, f0y, g);                    // :416
        const Taps t1 = make_taps<MODE>(ti.x, ti.y, f1x, f1y, g);                    // :418
        Quad q0[3], q1[3];
        gather3<T, PACKED>(fr.f0, fr.sc, t0, g.W, q0);
        gather3<T, PACKED>(fr.f1, fr.sc, t1, g.W, q1);
        const float rz = __frcp_rn(omt * v0 + tt * v1);                              // 1/Z  :425
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float w0 = v0 * bilerp(q0[c], t0);                                 // :420
            const float w1 = v1 * bilerp(q1[c], t1);                                 // :421
            const float s = omt * w0 + tt * w1;                                      // :423
            sts_(O + c * osc, s * rz);                                               // :427
        }
    }
}

// a3 + a4 backward, gather part: gradients w.r.t. the U-Net output (5 ch) and the estimated flows
// (input_tensor[:, 6:10]).  With STAGE, d/d(warped I0), d/d(warped I1) are written to an fp32
// staging buffer (B x N x 6 x H x W) and their max magnitude is recorded for the deterministic
// image-gradient pass (ssm_scatter.cuh).
template <typename T, int MODE, bool PACKED, bool STAGE>
__global__ void __launch_bounds__(TILE_THREADS)
fuse_bwd_kernel(View<const T> g3, View<const T> img6, const T* __restrict__ packed, View<const T> flows4,
                View<const T> out5, const float* __restrict__ tv, View<T> gout5, View<T> gflows4,
                float* __restrict__ stage, ScatterHdr* hdr, int N, Geom g) {
    TileIdx ti = tile_index(g.H, g.W);
    float amax = 0.0f;
    if (ti.valid) {
        const int p = ti.y * g.W + ti.x;
        const long long npx = (long long)g.H * g.W;
        const Frames<T, PACKED> fr(img6, packed, ti.b, npx);
        const float* tp = tv + ti.b * N;
        for (int n = 0; n < N; ++n) {
            const float tt = __ldg(tp + n);
            const float omt = __fsub_rn(1.0f, tt);
            const T* X = flows4.p + ti.b * flows4.sb + n * flows4.sn + p;
            const T* Y = out5.p + ti.b * out5.sb + n * out5.sn + p;
            const T* G = g3.p + ti.b * g3.sb + n * g3.sn + p;
            const float v1 = sigmoid_(lds_(Y));
            const float v0 = 1.0f - v1;
            const float f1x = __fadd_rn(lds_(X), lds_(Y + out5.sc));
            const float f1y = __fadd_rn(lds_(X + flows4.sc), lds_(Y + 2 * out5.sc));
            const float f0x = __fadd_rn(lds_(X + 2 * flows4.sc), lds_(Y + 3 * out5.sc));
            const float f0y = __fadd_rn(lds_(X + 3 * flows4.sc), lds_(Y + 4 * out5.sc));
            const Taps t0 = make_taps<MODE>(ti.x, ti.y, f0x, f0y, g);
            const Taps t1 = make_taps<MODE>(ti.x, ti.y, f1x, f1y, g);
            Quad q0[3], q1[3];
            gather3<T, PACKED>(fr.f0, fr.sc, t0, g.W, q0);
            gather3<T, PACKED>(fr.f1, fr.sc, t1, g.W, q1);
            const float z = omt * v0 + tt * v1;
            const float rz = __frcp_rn(z);
            float dz = 0, dv0 = 0, dv1 = 0, g0x = 0, g0y = 0, g1x = 0, g1y = 0;
            float* st = STAGE ? stage + ((long long)(ti.b * N + n) * 6) * npx + p : nullptr;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float s0 = bilerp(q0[c], t0), s1 = bilerp(q1[c], t1);
                const float o = (omt * (v0 * s0) + tt * (v1 * s1)) * rz;
                const float gc = lds_(G + c * g3.sc);
                const float ds = gc * rz;          // d/d(weighted_sum)
                dz -= gc * o * rz;                 // d/d(normalization_factor)
                const float dw0 = omt * ds, dw1 = tt * ds;
                dv0 += dw0 * s0; dv1 += dw1 * s1;
                const float ds0 = dw0 * v0, ds1 = dw1 * v1;   // d/d(warped frames)
                bilerp_grad(q0[c], t0, ds0, g0x, g0y);
                bilerp_grad(q1[c], t1, ds1, g1x, g1y);
                if (STAGE) {
                    st[c * npx] = ds0; st[(3 + c) * npx] = ds1;
                    amax = fmaxf(amax, fmaxf(fabsf(ds0), fabsf(ds1)));
                }
            }
            dv0 += omt * dz; dv1 += tt * dz;
            const float df1x = coord_grad_to_flow<MODE>(g1x, g.xgrad, g.xnorm, g.xinv);
            const float df1y = coord_grad_to_flow<MODE>(g1y, g.ygrad, g.ynorm, g.yinv);
            const float df0x = coord_grad_to_flow<MODE>(g0x, g.xgrad, g.xnorm, g.xinv);
            const float df0y = coord_grad_to_flow<MODE>(g0y, g.ygrad, g.ynorm, g.yinv);
            if (gout5.p) {
                T* o = gout5.p + ti.b * gout5.sb + n * gout5.sn + p;
                sts_(o, (dv1 - dv0) * (v1 * (1.0f - v1)));
                sts_(o + gout5.sc, df1x); sts_(o + 2 * gout5.sc, df1y);
                sts_(o + 3 * gout5.sc, df0x); sts_(o + 4 * gout5.sc, df0y);
            }
            if (gflows4.p) {
                T* o = gflows4.p + ti.b * gflows4.sb + n * gflows4.sn + p;
                sts_(o, df1x); sts_(o + gflows4.sc, df1y);
                sts_(o + 2 * gflows4.sc, df0x); sts_(o + 3 * gflows4.sc, df0y);
            }
        }
    }
    if (STAGE) record_absmax(&hdr->absmax_bits, amax);
}

}  // namespace ssm
