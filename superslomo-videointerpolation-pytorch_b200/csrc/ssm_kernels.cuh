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

#ifndef SSM_TILE_W
#define SSM_TILE_W 32
#endif
#ifndef SSM_TILE_H
#define SSM_TILE_H 8
#endif
constexpr int TILE_W = SSM_TILE_W;   // a multiple of 32: a warp is always 32 consecutive pixels of a row
constexpr int TILE_H = SSM_TILE_H;
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

// one image (B x 3 x H x W planar) -> B x H x W x 4 (RGBx): the stand-alone warp's staging copy, for callers that
// warp the same image more than once (losses.py:152-162 warps each frame by two flows)
template <typename T>
__global__ void __launch_bounds__(TILE_THREADS)
pack_image_kernel(View<const T> img3, T* __restrict__ packed, Geom g) {
    TileIdx ti = tile_index(g.H, g.W);
    if (!ti.valid) return;
    const int p = ti.y * g.W + ti.x;
    const long long npx = (long long)g.H * g.W;
    const T* I = img3.p + ti.b * img3.sb + p;
    T* o = packed + ((long long)ti.b * npx + p) * 4;
    const float r = lds_(I), gg = lds_(I + img3.sc), bb = lds_(I + 2 * img3.sc);
    if (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(o) = make_float4(r, gg, bb, 0.0f);
    } else {
        const unsigned lo = (unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(r)) |
                            ((unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(gg)) << 16);
        const unsigned hi = (unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(bb));
        *reinterpret_cast<uint2*>(o) = make_uint2(lo, hi);
    }
}

// =============================================================================================
// a1: warp forward          reference scripts/models/layers.py:73-120
//     planar: any channel count, one 4-byte gather per tap and channel.  PACKED (C = 3): img.p points at the RGBx
//     copy of the image (pack_image_kernel, batch stride H*W*4): one 16-byte gather per tap.
// =============================================================================================
template <typename T, int MODE, bool PACKED>
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
    if constexpr (PACKED) {
        Quad q[3];
        gather3<T, true>(ip, 0, t, g.W, q);
#pragma unroll
        for (int c = 0; c < 3; ++c) sts_(op + c * out.sc, bilerp(q[c], t));
    } else {
        for (int c = 0; c < C; ++c) {
            Quad q = gather_quad(ip + c * img.sc, t, g.W);
            sts_(op + c * out.sc, bilerp(q, t));
        }
    }
}

// warp backward, gather part: gradient w.r.t. the flow.  When hdr is given it also records
// max |grad_out| for the deterministic image-gradient pass (ssm_scatter.cuh), which re-reads
// grad_out and flow directly.
template <typename T, int MODE, bool PACKED>
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
        if constexpr (PACKED) {
            Quad q[3];
            gather3<T, true>(ip, 0, t, g.W, q);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float gc = lds_(gp + c * gout.sc);
                amax = fmaxf(amax, fabsf(gc));
                bilerp_grad(q[c], t, gc, gix, giy);
            }
        } else {
            for (int c = 0; c < C; ++c) {
                Quad q = gather_quad(ip + c * img.sc, t, g.W);
                float gc = lds_(gp + c * gout.sc);
                amax = fmaxf(amax, fabsf(gc));
                bilerp_grad(q, t, gc, gix, giy);
            }
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
#ifndef SSM_PACK_MIN_BLOCKS
#define SSM_PACK_MIN_BLOCKS 4
#endif
// NHWC: the 16 channels of a pixel are written next to each other
// (B x N x H x W x 16, i.e. torch.channels_last per (pair, timestep)) in the storage type TO, which may
// differ from the input type -- with TO = bf16 this is the tensor conv1a of the stage-2 U-Net consumes
// under channels-last bf16 autocast (flow_interpolation.py:36-38), so no layout or dtype conversion pass
// runs between compute_inputs and the U-Net; a thread then writes 32 (bf16) or 64 (fp32) contiguous
// bytes per timestep as one or two 32-byte stores instead of 16 4-byte ones.
// One 256-bit streaming store (sm_100: STG.256): a thread writes whole 32-byte sectors, so the two
// halves of a sector never travel as separate byte-masked writes.  p must be 32-byte aligned.
__device__ __forceinline__ void stcs256(void* p, const unsigned (&w)[8]) {
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
template <typename TO> __device__ __forceinline__ void store16_nhwc(TO* o, const float (&v)[16]);
template <> __device__ __forceinline__ void store16_nhwc<float>(float* o, const float (&v)[16]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        unsigned w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = __float_as_uint(v[8 * h + k]);
        stcs256(o + 8 * h, w);
    }
}
template <> __device__ __forceinline__ void store16_nhwc<__nv_bfloat16>(__nv_bfloat16* o, const float (&v)[16]) {
    unsigned w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);   // .x = low half = even channel
        w[k] = *reinterpret_cast<const unsigned*>(&h);
    }
    stcs256(o, w);
}

template <typename T, int MODE, bool PACKED, typename TO = T, bool NHWC = false>
__global__ void __launch_bounds__(TILE_THREADS, SSM_PACK_MIN_BLOCKS)
flow_pack_fwd_kernel(View<const T> img6, const T* __restrict__ packed, View<const T> flow4,
                     const float* __restrict__ tv, View<TO> out16, int N, Geom g) {
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
    TO* __restrict__ O = out16.p + ti.b * out16.sb + (NHWC ? (long long)p * 16 : (long long)p);
    const int osc = (int)out16.sc;   // fits 32 bits (checked on the host): one IMAD.WIDE per address
    for (int n = 0; n < N; ++n, O += out16.sn) {
        const Coef k = make_coef(__ldg(tp + n));
        // rounded to the storage type before they are used: the warped channels then correspond to the flows a
        // consumer reads back from channels 6:10 (and to what fuse_fwd / est_flows recompute); a no-op in fp32
        const float e0x = storage_round<T>(est_t0(k, f01x, f10x)), e0y = storage_round<T>(est_t0(k, f01y, f10y));   // F_t0  :353
        const float e1x = storage_round<T>(est_t1(k, f01x, f10x)), e1y = storage_round<T>(est_t1(k, f01y, f10y));   // F_t1  :356
        const Taps t1 = make_taps<MODE>(ti.x, ti.y, e1x, e1y, g);               // warp(img_1, F_t1) :361
        const Taps t0 = make_taps<MODE>(ti.x, ti.y, e0x, e0y, g);               // warp(img_0, F_t0) :362
        Quad q1[3], q0[3];
        gather3<T, PACKED>(fr.f1, fr.sc, t1, g.W, q1);
        gather3<T, PACKED>(fr.f0, fr.sc, t0, g.W, q0);
        if constexpr (NHWC) {                                                   // :364-367, channels-last
            float o[16];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                o[c] = c1[c]; o[3 + c] = bilerp(q1[c], t1); o[10 + c] = bilerp(q0[c], t0); o[13 + c] = c0[c];
            }
            o[6] = e1x; o[7] = e1y; o[8] = e0x; o[9] = e0y;
            store16_nhwc<TO>(O, o);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {                                       // :364-367
                sts_(O + (0 + c) * osc, c1[c]);
                sts_(O + (3 + c) * osc, bilerp(q1[c], t1));
                sts_(O + (10 + c) * osc, bilerp(q0[c], t0));
                sts_(O + (13 + c) * osc, c0[c]);
            }
            sts_(O + 6 * osc, e1x); sts_(O + 7 * osc, e1y);
            sts_(O + 8 * osc, e0x); sts_(O + 9 * osc, e0y);
        }
    }
}

// a2 backward, gather part: gradient w.r.t. flow_pred_tensor, summed over the N timesteps in
// registers (deterministic).  With IMG_GRAD the direct image gradients (channels 0:3 and 13:16 of
// grad16, summed over timesteps) go to an fp32 staging buffer and max |grad16[:, 3:6|10:13]| is
// recorded; the warped-image part is added by the scatter pass (ssm_scatter.cuh).
// The streaming loads of timestep n+1 are issued before the gathers of timestep n (two dependent
// long-latency phases overlap across iterations, as in fuse_fwd_kernel).
#ifndef SSM_PACK_BWD_MIN_BLOCKS
#define SSM_PACK_BWD_MIN_BLOCKS 4
#endif
template <typename T, int MODE, bool PACKED, bool IMG_GRAD>
__global__ void __launch_bounds__(TILE_THREADS, IMG_GRAD ? 3 : SSM_PACK_BWD_MIN_BLOCKS)
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
        const T* __restrict__ G = g16.p + ti.b * g16.sb + p;
        const int gsc = (int)g16.sc;   // fits 32 bits (checked on the host)
        // gs[0:3] = d/d(warped I1) (ch 3:6), gs[3:7] = direct flow terms (ch 6:10), gs[7:10] = d/d(warped I0)
        float gs[10], gd[6];
#pragma unroll
        for (int k = 0; k < 10; ++k) gs[k] = lds_(G + (3 + k) * gsc);
        if (IMG_GRAD) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { gd[c] = lds_(G + c * gsc); gd[3 + c] = lds_(G + (13 + c) * gsc); }
        }
        for (int n = 0; n < N; ++n) {
            const Coef k = make_coef(__ldg(tp + n));
            float gw1[3], gw0[3], gdir[4];
#pragma unroll
            for (int c = 0; c < 3; ++c) { gw1[c] = gs[c]; gw0[c] = gs[7 + c]; }
#pragma unroll
            for (int c = 0; c < 4; ++c) gdir[c] = gs[3 + c];
            if (IMG_GRAD) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    amax = fmaxf(amax, fmaxf(fabsf(gw1[c]), fabsf(gw0[c])));
                    di1[c] += gd[c];
                    di0[c] += gd[3 + c];
                }
            }
            if (n + 1 < N) {
                G += g16.sn;
#pragma unroll
                for (int c = 0; c < 10; ++c) gs[c] = lds_(G + (3 + c) * gsc);
                if (IMG_GRAD) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) { gd[c] = lds_(G + c * gsc); gd[3 + c] = lds_(G + (13 + c) * gsc); }
                }
            }
            if (!want_flow) continue;
            const float e0x = storage_round<T>(est_t0(k, f01x, f10x)), e0y = storage_round<T>(est_t0(k, f01y, f10y));
            const float e1x = storage_round<T>(est_t1(k, f01x, f10x)), e1y = storage_round<T>(est_t1(k, f01y, f10y));
            float g1x = 0, g1y = 0, g0x = 0, g0y = 0;
            {
                const Taps t1 = make_taps<MODE>(ti.x, ti.y, e1x, e1y, g);
                Quad q1[3];
                gather3<T, PACKED>(fr.f1, fr.sc, t1, g.W, q1);
#pragma unroll
                for (int c = 0; c < 3; ++c) bilerp_grad(q1[c], t1, gw1[c], g1x, g1y);
            }
            {
                const Taps t0 = make_taps<MODE>(ti.x, ti.y, e0x, e0y, g);
                Quad q0[3];
                gather3<T, PACKED>(fr.f0, fr.sc, t0, g.W, q0);
#pragma unroll
                for (int c = 0; c < 3; ++c) bilerp_grad(q0[c], t0, gw0[c], g0x, g0y);
            }
            const float de1x = gdir[0] + coord_grad_to_flow<MODE>(g1x, g.xgrad, g.xnorm, g.xinv);
            const float de1y = gdir[1] + coord_grad_to_flow<MODE>(g1y, g.ygrad, g.ynorm, g.yinv);
            const float de0x = gdir[2] + coord_grad_to_flow<MODE>(g0x, g.xgrad, g.xnorm, g.xinv);
            const float de0y = gdir[3] + coord_grad_to_flow<MODE>(g0y, g.ygrad, g.ynorm, g.yinv);
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
//     RECOMP: the estimated flows F_t1, F_t0 (input_tensor[:, 6:10]) are not read back but recomputed
//     from the stage-1 flows and t with the arithmetic of flow_pack_fwd_kernel (bit-identical to the
//     stored values): `flows4` is then flow_pred_tensor (B x 4 x H x W) and 4 of the 12 streamed
//     channels per timestep disappear.
// =============================================================================================
#ifndef SSM_FUSE_MIN_BLOCKS
#define SSM_FUSE_MIN_BLOCKS 4
#endif
// the estimated flows of one timestep in input_tensor order (F_t1.x, F_t1.y, F_t0.x, F_t0.y), rounded
// to the storage type exactly as flow_pack_fwd_kernel stores them
template <typename T>
__device__ __forceinline__ void est_flows(float tt, const float (&f)[4], float (&xs)[4]) {
    const Coef k = make_coef(tt);
    xs[0] = storage_round<T>(est_t1(k, f[0], f[2])); xs[1] = storage_round<T>(est_t1(k, f[1], f[3]));
    xs[2] = storage_round<T>(est_t0(k, f[0], f[2])); xs[3] = storage_round<T>(est_t0(k, f[1], f[3]));
}

// TY: storage type of the U-Net output (out5).  It may differ from T: under bf16 autocast final_conv
// (flow_interpolation.py:149-157) produces bf16 while frames and flows stay fp32; reading it as it is
// gives exactly the values of out5.float() without that conversion pass.
template <typename T, int MODE, bool PACKED, bool RECOMP, typename TY = T>
__global__ void __launch_bounds__(TILE_THREADS, SSM_FUSE_MIN_BLOCKS)
fuse_fwd_kernel(View<const T> img6, const T* __restrict__ packed, View<const T> flows4, View<const TY> out5,
                const float* __restrict__ tv, View<T> out3, int N, Geom g) {
    TileIdx ti = tile_index(g.H, g.W);
    if (!ti.valid) return;
    const int p = ti.y * g.W + ti.x;
    const long long npx = (long long)g.H * g.W;
    const Frames<T, PACKED> fr(img6, packed, ti.b, npx);
    const float* tp = tv + ti.b * N;
    const T* __restrict__ X = flows4.p + ti.b * flows4.sb + p;
    const TY* __restrict__ Y = out5.p + ti.b * out5.sb + p;
    T* __restrict__ O = out3.p + ti.b * out3.sb + p;
    // channel strides fit 32 bits (checked on the host): one IMAD.WIDE per address
    const int xsc = (int)flows4.sc, ysc = (int)out5.sc, osc = (int)out3.sc;
    // The streaming loads of timestep n+1 are issued before the gathers of timestep n, so that two
    // dependent long-latency phases (HBM stream, then L2/L1 gather) overlap across iterations.
    float xs[4], ys[5], f[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) (RECOMP ? f[k] : xs[k]) = lds_(X + k * xsc);
#pragma unroll
    for (int k = 0; k < 5; ++k) ys[k] = lds_(Y + k * ysc);
    for (int n = 0; n < N; ++n, O += out3.sn) {
        const float tt = __ldg(tp + n);
        const float omt = __fsub_rn(1.0f, tt);
        if (RECOMP) est_flows<T>(tt, f, xs);
        const float logit = ys[0];
        const float f1x = __fadd_rn(xs[0], ys[1]);                                   // :412
        const float f1y = __fadd_rn(xs[1], ys[2]);
        const float f0x = __fadd_rn(xs[2], ys[3]);                                   // :413
        const float f0y = __fadd_rn(xs[3], ys[4]);
        if (n + 1 < N) {
            Y += out5.sn;
#pragma unroll
            for (int k = 0; k < 5; ++k) ys[k] = lds_(Y + k * ysc);
            if (!RECOMP) {
                X += flows4.sn;
#pragma unroll
                for (int k = 0; k < 4; ++k) xs[k] = lds_(X + k * xsc);
            }
        }
        const float v1 = sigmoid_(logit);                                            // :386-388
        const float v0 = 1.0f - v1;                                                  // :390
        const Taps t0 = make_taps<MODE>(ti.x, ti.y, f0x, f0y, g);                    // :416
        const Taps t1 = make_taps<MODE>(ti.x, ti.y, f1x, f1y, g);                    // :418
        Quad q0[3], q1[3];
        gather3<T, PACKED>(fr.f0, fr.sc, t0, g.W, q0);
        gather3<T, PACKED>(fr.f1, fr.sc, t1, g.W, q1);
        const float rz = rcp_approx(omt * v0 + tt * v1);                              // 1/Z  :425
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
// (input_tensor[:, 6:10]).  With STAGE (image gradients wanted) max |G| of the launch is recorded for the
// deterministic image-gradient pass (ssm_scatter.cuh): it bounds every d/d(warped I_f) = k_f V_f G_c because
// k_f V_f <= 1.  That pass recomputes d/d(warped I_f) from G and the logit itself; round 1 staged them in a
// B x N x 6 x H x W fp32 buffer (5.6 GB written and read back at 16 x 1088 x 1920 x 7, and 128 registers here).
//
// With G = d/d(out3), S = (1-t) V0 w0 + t V1 w1, Z = (1-t) V0 + t V1 (SURVEY.md section 8 note):
//   d/d(w_f,c) = k_f V_f G_c,  k_0 = (1-t)/Z, k_1 = t/Z
//   A_f = sum_c G_c w_f,c ;  dZ = -(k_0 V0 A_0 + k_1 V1 A_1) / Z ;  dV_0 = k_0 A_0 + (1-t) dZ ; dV_1 alike
// so each frame contributes three scalars (A_f and the two coordinate gradients) and the frames are
// processed one after the other: 12 gathered values live at a time instead of 24.
#ifndef SSM_FUSE_BWD_MIN_BLOCKS
#define SSM_FUSE_BWD_MIN_BLOCKS 4
#endif
template <typename T, bool PACKED, bool STAGE>
__device__ __forceinline__ void fuse_bwd_frame(const T* __restrict__ frame, long long sc, const Taps& t, int W,
                                               const float (&gc)[3], float kv, float& A, float& gx, float& gy,
                                               float* __restrict__ st, long long npx) {
    Quad q[3];
    gather3<T, PACKED>(frame, sc, t, W, q);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        A = fmaf(gc[c], bilerp(q[c], t), A);
        const float ds = kv * gc[c];                   // d/d(warped frame)
        bilerp_grad(q[c], t, ds, gx, gy);
    }
}

// RECOMP (see fuse_fwd_kernel): `flows4` is flow_pred_tensor (B x 4 x H x W) and `gflows4` its
// gradient (B x 4 x H x W), accumulated over the N timesteps in registers through the coefficients
// of flow_interpolation.py:353,356 -- no B x N x 4 gradient (nor the 12 zero channels around it in
// the gradient of input_tensor) is ever materialised.
template <typename T, int MODE, bool PACKED, bool STAGE, bool RECOMP>
__global__ void __launch_bounds__(TILE_THREADS, SSM_FUSE_BWD_MIN_BLOCKS)
fuse_bwd_kernel(View<const T> g3, View<const T> img6, const T* __restrict__ packed, View<const T> flows4,
                View<const T> out5, const float* __restrict__ tv, View<T> gout5, View<T> gflows4,
                float* __restrict__ stage, ScatterHdr* hdr, int N, Geom g) {
    TileIdx ti = tile_index(g.H, g.W);
    if (ti.valid) {
        const int p = ti.y * g.W + ti.x;
        const long long npx = (long long)g.H * g.W;
        const Frames<T, PACKED> fr(img6, packed, ti.b, npx);
        const float* tp = tv + ti.b * N;
        const T* __restrict__ X = flows4.p + ti.b * flows4.sb + p;
        const T* __restrict__ Y = out5.p + ti.b * out5.sb + p;
        const T* __restrict__ G = g3.p + ti.b * g3.sb + p;
        T* __restrict__ O5 = gout5.p ? gout5.p + ti.b * gout5.sb + p : nullptr;
        T* __restrict__ O4 = gflows4.p ? gflows4.p + ti.b * gflows4.sb + p : nullptr;
        // channel strides fit 32 bits (checked on the host)
        const int xsc = (int)flows4.sc, ysc = (int)out5.sc, gsc = (int)g3.sc;
        const int o5sc = (int)gout5.sc, o4sc = (int)gflows4.sc;
        float xs[4], ys[5], gs[3], f[4];
        float d01x = 0, d01y = 0, d10x = 0, d10y = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) (RECOMP ? f[k] : xs[k]) = lds_(X + k * xsc);
#pragma unroll
        for (int k = 0; k < 5; ++k) ys[k] = lds_(Y + k * ysc);
#pragma unroll
        for (int k = 0; k < 3; ++k) gs[k] = lds_(G + k * gsc);
        for (int n = 0; n < N; ++n) {
            const float tt = __ldg(tp + n);
            const float omt = __fsub_rn(1.0f, tt);
            if (RECOMP) est_flows<T>(tt, f, xs);
            const float logit = ys[0];
            const float f1x = __fadd_rn(xs[0], ys[1]);
            const float f1y = __fadd_rn(xs[1], ys[2]);
            const float f0x = __fadd_rn(xs[2], ys[3]);
            const float f0y = __fadd_rn(xs[3], ys[4]);
            const float gc[3] = {gs[0], gs[1], gs[2]};
            if (STAGE) {
                // max |G| of the launch, recorded timestep by timestep instead of carried in a register to the end of the
                // kernel (that register, and the code behind the pixel's block, cost ~80 bytes of spills at the 64
                // registers of 4 CTAs/SM: 5.3 ms instead of 3.8, profiles/r04g, r04h).  Integer max of the |.| bit patterns:
                // NaN compares above every finite value and poisons the scale.  The pre-check reads through L1: a stale
                // value is only ever too small (the slot grows monotonically) -- a redundant atomic, never a wrong result.
                const unsigned int mb = max(max(__float_as_uint(fabsf(gc[0])), __float_as_uint(fabsf(gc[1]))), __float_as_uint(fabsf(gc[2])));
                if (mb > __ldca(&hdr->absmax_bits)) atomicMax(&hdr->absmax_bits, mb);
            }
            if (n + 1 < N) {     // streaming loads of the next timestep, in flight during the gathers
                Y += out5.sn; G += g3.sn;
                if (!RECOMP) {
                    X += flows4.sn;
#pragma unroll
                    for (int k = 0; k < 4; ++k) xs[k] = lds_(X + k * xsc);
                }
#pragma unroll
                for (int k = 0; k < 5; ++k) ys[k] = lds_(Y + k * ysc);
#pragma unroll
                for (int k = 0; k < 3; ++k) gs[k] = lds_(G + k * gsc);
            }
            const float v1 = sigmoid_(logit);
            const float v0 = 1.0f - v1;
            const float rz = rcp_approx(omt * v0 + tt * v1);
            const float k0 = omt * rz, k1 = tt * rz;
            float A0 = 0, A1 = 0, g0x = 0, g0y = 0, g1x = 0, g1y = 0;
            float* st = nullptr;          // (round 1: the staging buffer of d/d(warped I_f))
            {
                const Taps t0 = make_taps<MODE>(ti.x, ti.y, f0x, f0y, g);
                fuse_bwd_frame<T, PACKED, STAGE>(fr.f0, fr.sc, t0, g.W, gc, k0 * v0, A0, g0x, g0y, st, npx);
            }
            {
                const Taps t1 = make_taps<MODE>(ti.x, ti.y, f1x, f1y, g);
                fuse_bwd_frame<T, PACKED, STAGE>(fr.f1, fr.sc, t1, g.W, gc, k1 * v1, A1, g1x, g1y, st, npx);
            }
            const float dz = -rz * (k0 * v0 * A0 + k1 * v1 * A1);      // d/d(normalization_factor)
            const float dv0 = k0 * A0 + omt * dz, dv1 = k1 * A1 + tt * dz;
            const float df1x = coord_grad_to_flow<MODE>(g1x, g.xgrad, g.xnorm, g.xinv);
            const float df1y = coord_grad_to_flow<MODE>(g1y, g.ygrad, g.ynorm, g.yinv);
            const float df0x = coord_grad_to_flow<MODE>(g0x, g.xgrad, g.xnorm, g.xinv);
            const float df0y = coord_grad_to_flow<MODE>(g0y, g.ygrad, g.ynorm, g.yinv);
            if (O5) {
                sts_(O5, (dv1 - dv0) * (v1 * (1.0f - v1)));
                sts_(O5 + o5sc, df1x); sts_(O5 + 2 * o5sc, df1y);
                sts_(O5 + 3 * o5sc, df0x); sts_(O5 + 4 * o5sc, df0y);
                O5 += gout5.sn;
            }
            if (RECOMP) {
                const Coef k = make_coef(tt);
                d01x += k.c00 * df0x + k.c10 * df1x;
                d01y += k.c00 * df0y + k.c10 * df1y;
                d10x += k.c01 * df0x - k.c11 * df1x;
                d10y += k.c01 * df0y - k.c11 * df1y;
            } else if (O4) {
                sts_(O4, df1x); sts_(O4 + o4sc, df1y);
                sts_(O4 + 2 * o4sc, df0x); sts_(O4 + 3 * o4sc, df0y);
                O4 += gflows4.sn;
            }
        }
        if (RECOMP && O4) {
            sts_(O4, d01x); sts_(O4 + o4sc, d01y);
            sts_(O4 + 2 * o4sc, d10x); sts_(O4 + 3 * o4sc, d10y);
        }
    }
}

// =============================================================================================
// a9 + a4 fused: compute_output_image together with the loss
// front-end of scripts/models/losses.py -- the L1 reconstruction term |I_t - target| (:111, :217),
// the stage-2 warp loss |g(I0, F^_t0) - target| + |g(I1, F^_t1) - target| (:152-154, :166-167, whose
// two warps are exactly the ones compute_output_image performs, flow_interpolation.py:416-418) and
// the stage-1 warp loss |g(I1, F01) - I0| + |g(I0, F10) - I1| (:160-163).  The kernel writes the fused
// frame and per-CTA partial sums of the three L1 terms; loss_reduce_kernel adds the partials of
// each sample in a fixed order (deterministic, fp64).  The per-sample means and the lambda weights
// (losses.py:213-233) are applied by the caller.
// partials: [grid][2 N + 1] floats = (rec, warp2) per timestep, then warp1.
// =============================================================================================
__device__ __forceinline__ float block_sum(float v, float* red) {   // result valid in thread 0
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.0f;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < TILE_THREADS / 32; ++w) r += red[w];
    }
    __syncthreads();
    return r;
}

__device__ __forceinline__ float sign_(float v) { return (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f); }

template <typename T, int MODE, bool PACKED>
__global__ void __launch_bounds__(TILE_THREADS, 3)
fuse_loss_fwd_kernel(View<const T> img6, const T* __restrict__ packed, View<const T> flow4, View<const T> out5,
                     View<const T> target, const float* __restrict__ tv, View<T> out3, float* __restrict__ partials,
                     int N, Geom g, int stage1_on, int stage2_on) {
    __shared__ float red[TILE_THREADS / 32];
    const TileIdx ti = tile_index(g.H, g.W);
    const int p = ti.y * g.W + ti.x;
    const long long npx = (long long)g.H * g.W;
    const Frames<T, PACKED> fr(img6, packed, ti.b, npx);
    const float* tp = tv + ti.b * N;
    float* part = partials + (long long)blockIdx.x * (2 * N + 1);
    float f[4] = {0.f, 0.f, 0.f, 0.f};
    float w1sum = 0.0f;
    if (ti.valid) {
        const T* F = flow4.p + ti.b * flow4.sb + p;
#pragma unroll
        for (int k = 0; k < 4; ++k) f[k] = lds_(F + k * (int)flow4.sc);
        if (stage1_on) {                                              // losses.py:160-163
            float c0[3], c1[3];
            if (PACKED) { load_px(fr.f0 + (long long)p * 4, c0); load_px(fr.f1 + (long long)p * 4, c1); }
            else {
#pragma unroll
                for (int c = 0; c < 3; ++c) { c0[c] = ldg_(fr.f0 + c * fr.sc + p); c1[c] = ldg_(fr.f1 + c * fr.sc + p); }
            }
            const Taps ta = make_taps<MODE>(ti.x, ti.y, f[0], f[1], g);   // warp(img_1, flow_01) vs img_0
            const Taps tb = make_taps<MODE>(ti.x, ti.y, f[2], f[3], g);   // warp(img_0, flow_10) vs img_1
            Quad qa[3], qb[3];
            gather3<T, PACKED>(fr.f1, fr.sc, ta, g.W, qa);
            gather3<T, PACKED>(fr.f0, fr.sc, tb, g.W, qb);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                w1sum += fabsf(storage_round<T>(bilerp(qa[c], ta)) - c0[c]) + fabsf(storage_round<T>(bilerp(qb[c], tb)) - c1[c]);
        }
    }
    for (int n = 0; n < N; ++n) {
        float rec = 0.0f, w2 = 0.0f;
        if (ti.valid) {
            const float tt = __ldg(tp + n);
            const float omt = __fsub_rn(1.0f, tt);
            float xs[4];
            est_flows<T>(tt, f, xs);
            const T* Y = out5.p + ti.b * out5.sb + n * out5.sn + p;
            const T* TG = target.p + ti.b * target.sb + n * target.sn + p;
            T* O = out3.p + ti.b * out3.sb + n * out3.sn + p;
            const int ysc = (int)out5.sc, tsc = (int)target.sc, osc = (int)out3.sc;
            const float logit = lds_(Y);
            const float f1x = __fadd_rn(xs[0], lds_(Y + ysc)), f1y = __fadd_rn(xs[1], lds_(Y + 2 * ysc));
            const float f0x = __fadd_rn(xs[2], lds_(Y + 3 * ysc)), f0y = __fadd_rn(xs[3], lds_(Y + 4 * ysc));
            float tg[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) tg[c] = lds_(TG + c * tsc);
            const float v1 = sigmoid_(logit);
            const float v0 = 1.0f - v1;
            const Taps t0 = make_taps<MODE>(ti.x, ti.y, f0x, f0y, g);
            const Taps t1 = make_taps<MODE>(ti.x, ti.y, f1x, f1y, g);
            Quad q0[3], q1[3];
            gather3<T, PACKED>(fr.f0, fr.sc, t0, g.W, q0);
            gather3<T, PACKED>(fr.f1, fr.sc, t1, g.W, q1);
            const float rz = rcp_approx(omt * v0 + tt * v1);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float s0 = bilerp(q0[c], t0), s1 = bilerp(q1[c], t1);
                const float o = storage_round<T>((omt * (v0 * s0) + tt * (v1 * s1)) * rz);
                sts_(O + c * osc, o);
                rec += fabsf(o - tg[c]);                                              // losses.py:111
                if (stage2_on) w2 += fabsf(storage_round<T>(s0) - tg[c]) + fabsf(storage_round<T>(s1) - tg[c]);   // :166-167
            }
        }
        const float r = block_sum(rec, red);
        const float w = block_sum(w2, red);
        if (threadIdx.x == 0) { part[2 * n] = r; part[2 * n + 1] = w; }
    }
    const float w1 = block_sum(w1sum, red);
    if (threadIdx.x == 0) part[2 * N] = w1;
}

// sums[b][k] = sum over the tiles of pair b of partials[b * tiles + tile][k], k < K = 2 N + 1.
// One CTA per pair; fixed assignment of tiles to threads and a fixed-order tree: deterministic.
__global__ void __launch_bounds__(256)
loss_reduce_kernel(const float* __restrict__ partials, int tiles, int K, float* __restrict__ sums) {
    __shared__ double red[256];
    const int b = blockIdx.x;
    for (int k = 0; k < K; ++k) {
        double acc = 0.0;
        for (int i = threadIdx.x; i < tiles; i += 256) acc += (double)partials[((long long)b * tiles + i) * K + k];
        red[threadIdx.x] = acc;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) sums[(long long)b * K + k] = (float)red[0];
        __syncthreads();
    }
}

// Backward of fuse_loss_fwd_kernel: gradients w.r.t. the U-Net output (B x N x 5) and the stage-1
// flows (B x 4, summed over timesteps and over both loss paths).  g3 (may be null) is the dense
// upstream gradient of the fused frames (e.g. from the perceptual loss); gsum[b][2N+1] holds
// dL/d(rec_n), dL/d(warp2_n), dL/d(warp1).  The L1 terms differentiate to sign() (torch: sign(0) = 0);
// the fused frame is re-read (out3) so that the effective frame gradient is known before the gathers
// and the per-frame reduced form of fuse_bwd_kernel applies.  Frames and targets are data: no image
// gradients here (callers that need them use the unfused path).
template <typename T, int MODE, bool PACKED>
__global__ void __launch_bounds__(TILE_THREADS, 3)
fuse_loss_bwd_kernel(View<const T> g3, const float* __restrict__ gsum, View<const T> img6, const T* __restrict__ packed,
                     View<const T> flow4, View<const T> out5, View<const T> target, View<const T> out3,
                     const float* __restrict__ tv, View<T> gout5, View<T> gflow4, int N, Geom g,
                     int stage1_on, int stage2_on) {
    const TileIdx ti = tile_index(g.H, g.W);
    if (!ti.valid) return;
    const int p = ti.y * g.W + ti.x;
    const long long npx = (long long)g.H * g.W;
    const Frames<T, PACKED> fr(img6, packed, ti.b, npx);
    const float* tp = tv + ti.b * N;
    const float* gs = gsum + (long long)ti.b * (2 * N + 1);
    float f[4];
    {
        const T* F = flow4.p + ti.b * flow4.sb + p;
#pragma unroll
        for (int k = 0; k < 4; ++k) f[k] = lds_(F + k * (int)flow4.sc);
    }
    float d01x = 0, d01y = 0, d10x = 0, d10y = 0;
    if (stage1_on) {
        const float gw1 = __ldg(gs + 2 * N);
        float c0[3], c1[3];
        if (PACKED) { load_px(fr.f0 + (long long)p * 4, c0); load_px(fr.f1 + (long long)p * 4, c1); }
        else {
#pragma unroll
            for (int c = 0; c < 3; ++c) { c0[c] = ldg_(fr.f0 + c * fr.sc + p); c1[c] = ldg_(fr.f1 + c * fr.sc + p); }
        }
        float gax = 0, gay = 0, gbx = 0, gby = 0;
        {
            const Taps ta = make_taps<MODE>(ti.x, ti.y, f[0], f[1], g);
            Quad qa[3];
            gather3<T, PACKED>(fr.f1, fr.sc, ta, g.W, qa);
#pragma unroll
            for (int c = 0; c < 3; ++c) bilerp_grad(qa[c], ta, gw1 * sign_(storage_round<T>(bilerp(qa[c], ta)) - c0[c]), gax, gay);
        }
        {
            const Taps tb = make_taps<MODE>(ti.x, ti.y, f[2], f[3], g);
            Quad qb[3];
            gather3<T, PACKED>(fr.f0, fr.sc, tb, g.W, qb);
#pragma unroll
            for (int c = 0; c < 3; ++c) bilerp_grad(qb[c], tb, gw1 * sign_(storage_round<T>(bilerp(qb[c], tb)) - c1[c]), gbx, gby);
        }
        d01x = coord_grad_to_flow<MODE>(gax, g.xgrad, g.xnorm, g.xinv);
        d01y = coord_grad_to_flow<MODE>(gay, g.ygrad, g.ynorm, g.yinv);
        d10x = coord_grad_to_flow<MODE>(gbx, g.xgrad, g.xnorm, g.xinv);
        d10y = coord_grad_to_flow<MODE>(gby, g.ygrad, g.ynorm, g.yinv);
    }
    for (int n = 0; n < N; ++n) {
        const float tt = __ldg(tp + n);
        const float omt = __fsub_rn(1.0f, tt);
        const float grec = __ldg(gs + 2 * n), gw2 = stage2_on ? __ldg(gs + 2 * n + 1) : 0.0f;
        float xs[4];
        est_flows<T>(tt, f, xs);
        const T* Y = out5.p + ti.b * out5.sb + n * out5.sn + p;
        const T* TG = target.p + ti.b * target.sb + n * target.sn + p;
        const T* O = out3.p + ti.b * out3.sb + n * out3.sn + p;
        const int ysc = (int)out5.sc, tsc = (int)target.sc, osc = (int)out3.sc;
        const float logit = lds_(Y);
        const float f1x = __fadd_rn(xs[0], lds_(Y + ysc)), f1y = __fadd_rn(xs[1], lds_(Y + 2 * ysc));
        const float f0x = __fadd_rn(xs[2], lds_(Y + 3 * ysc)), f0y = __fadd_rn(xs[3], lds_(Y + 4 * ysc));
        float tg[3], gc[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            tg[c] = lds_(TG + c * tsc);
            gc[c] = grec * sign_(lds_(O + c * osc) - tg[c]);
            if (g3.p) gc[c] += lds_(g3.p + ti.b * g3.sb + n * g3.sn + p + c * (int)g3.sc);
        }
        const float v1 = sigmoid_(logit);
        const float v0 = 1.0f - v1;
        const float rz = rcp_approx(omt * v0 + tt * v1);
        const float k0 = omt * rz, k1 = tt * rz;
        float A0 = 0, A1 = 0, g0x = 0, g0y = 0, g1x = 0, g1y = 0;
        {
            const Taps t0 = make_taps<MODE>(ti.x, ti.y, f0x, f0y, g);
            Quad q[3];
            gather3<T, PACKED>(fr.f0, fr.sc, t0, g.W, q);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float sv = bilerp(q[c], t0);
                A0 = fmaf(gc[c], sv, A0);
                bilerp_grad(q[c], t0, k0 * v0 * gc[c] + gw2 * sign_(storage_round<T>(sv) - tg[c]), g0x, g0y);
            }
        }
        {
            const Taps t1 = make_taps<MODE>(ti.x, ti.y, f1x, f1y, g);
            Quad q[3];
            gather3<T, PACKED>(fr.f1, fr.sc, t1, g.W, q);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float sv = bilerp(q[c], t1);
                A1 = fmaf(gc[c], sv, A1);
                bilerp_grad(q[c], t1, k1 * v1 * gc[c] + gw2 * sign_(storage_round<T>(sv) - tg[c]), g1x, g1y);
            }
        }
        const float dz = -rz * (k0 * v0 * A0 + k1 * v1 * A1);
        const float dv0 = k0 * A0 + omt * dz, dv1 = k1 * A1 + tt * dz;
        const float df1x = coord_grad_to_flow<MODE>(g1x, g.xgrad, g.xnorm, g.xinv);
        const float df1y = coord_grad_to_flow<MODE>(g1y, g.ygrad, g.ynorm, g.yinv);
        const float df0x = coord_grad_to_flow<MODE>(g0x, g.xgrad, g.xnorm, g.xinv);
        const float df0y = coord_grad_to_flow<MODE>(g0y, g.ygrad, g.ynorm, g.yinv);
        if (gout5.p) {
            T* o = gout5.p + ti.b * gout5.sb + n * gout5.sn + p;
            const int o5sc = (int)gout5.sc;
            sts_(o, (dv1 - dv0) * (v1 * (1.0f - v1)));
            sts_(o + o5sc, df1x); sts_(o + 2 * o5sc, df1y);
            sts_(o + 3 * o5sc, df0x); sts_(o + 4 * o5sc, df0y);
        }
        const Coef k = make_coef(tt);
        d01x += k.c00 * df0x + k.c10 * df1x;
        d01y += k.c00 * df0y + k.c10 * df1y;
        d10x += k.c01 * df0x - k.c11 * df1x;
        d10y += k.c01 * df0y - k.c11 * df1y;
    }
    if (gflow4.p) {
        T* o = gflow4.p + ti.b * gflow4.sb + p;
        const int osc = (int)gflow4.sc;
        sts_(o, d01x); sts_(o + osc, d01y);
        sts_(o + 2 * osc, d10x); sts_(o + 3 * osc, d10y);
    }
}

}  // namespace ssm
