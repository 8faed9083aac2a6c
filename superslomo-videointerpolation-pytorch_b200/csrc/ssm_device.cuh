// ssm_device.cuh -- device-side building blocks shared by the synthesis kernels (sm_100a).
//
// The arithmetic that fixes the SAMPLING COORDINATE is written with explicitly rounded
// intrinsics (__fadd_rn, __fmul_rn, __fdiv_rn: never contracted into FMAs), because the reference
// evaluates it as separate torch kernels (scripts/models/layers.py:100-116, then ATen's
// grid_sampler un-normalisation) and a 1-ulp difference in the coordinate flips floor() at cell
// borders, which changes the flow gradient by O(1) (SURVEY.md finding 2).  Everything downstream
// of the coordinate (tap weights, interpolation, fusion) only has to agree to rounding error and
// is free to use FMAs.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ssm_b200.h"

namespace ssm {

// ---- storage access -------------------------------------------------------------------------
// ldg_: read-only, L1-cached (gather taps and frames: reused by neighbouring threads/timesteps)
// lds_: streaming read (touched once per launch)      sts_: streaming write
template <typename T> __device__ __forceinline__ float ldg_(const T* p);
template <> __device__ __forceinline__ float ldg_<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldg_<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(__ldg(p));
}
template <typename T> __device__ __forceinline__ float lds_(const T* p);
template <> __device__ __forceinline__ float lds_<float>(const float* p) { return __ldcs(p); }
template <> __device__ __forceinline__ float lds_<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(__ldcs(p));
}
template <typename T> __device__ __forceinline__ void sts_(T* p, float v);
template <> __device__ __forceinline__ void sts_<float>(float* p, float v) { __stcs(p, v); }
template <> __device__ __forceinline__ void sts_<__nv_bfloat16>(__nv_bfloat16* p, float v) {
    __stcs(p, __float2bfloat16_rn(v));
}
// round a computed flow to the storage type and back, so that a value written by one kernel and
// re-read by the next gives the same coordinate as the value used in-kernel (bf16 storage only)
template <typename T> __device__ __forceinline__ float storage_round(float v);
template <> __device__ __forceinline__ float storage_round<float>(float v) { return v; }
template <> __device__ __forceinline__ float storage_round<__nv_bfloat16>(float v) {
    return __bfloat162float(__float2bfloat16_rn(v));
}

template <typename T> struct View {
    T* p;
    long long sb, sn, sc;
};

// ---- geometry -------------------------------------------------------------------------------
struct Geom {
    int H, W;
    float xnorm, ynorm;  // float(max(W-1,1)), float(max(H-1,1))     layers.py:112-113
    float xinv, yinv;    // fp32 1/xnorm, 1/ynorm                    (SSM_COORD_RCP)
    float xm1, ym1;      // float(W-1), float(H-1)                   ATen un-normalise
    float xgrad, ygrad;  // float(W-1)/2, float(H-1)/2               ATen backward multiplier
};

// Correctly rounded s / d for a launch-constant divisor d with y = RN(1/d) precomputed on the host:
// two Markstein correction steps (q <- q + (s - q*d)*y with the residual exact in an FMA).  After
// the first step q is a faithful rounding of s/d; with y correctly rounded and the significand of
// d not all ones (d = W-1 is a small integer) the second step then returns RN(s/d).  5 instructions
// instead of the ~27 of the generic IEEE division sequence; checked exhaustively against
// __fdiv_rn over every fp32 s by ssm_selftest_division (tests/test_parity_gpu.py).
__device__ __forceinline__ float div_rn_const(float s, float d, float y) {
    float q = __fmul_rn(s, y);
    float r = __fmaf_rn(-q, d, s);
    q = __fmaf_rn(r, y, q);
    r = __fmaf_rn(-q, d, s);
    return __fmaf_rn(r, y, q);
}

// layers.py:100 (grid + flo), :112 (2.0*u/max(W-1,1) - 1.0), ATen ((c + 1)/2)*(size - 1)
template <int MODE>
__device__ __forceinline__ float sample_coord(float pos, float flow, float norm, float inv, float m1) {
    float g = __fadd_rn(pos, flow);
    float s = __fmul_rn(2.0f, g);
    float n = (MODE == SSM_COORD_DIV) ? div_rn_const(s, norm, inv) : __fmul_rn(s, inv);
    n = __fsub_rn(n, 1.0f);
    float a = __fadd_rn(n, 1.0f);
    a = __fmul_rn(a, 0.5f);
    return __fmul_rn(a, m1);
}

// chain rule from d/d(ix) to d/d(flow): ATen multiplies by (size-1)/2; autograd of layers.py:112
// divides by max(size-1,1) in the same rounding mode as the forward and multiplies by 2.
template <int MODE>
__device__ __forceinline__ float coord_grad_to_flow(float gi, float gmul, float norm, float inv) {
    float g = __fmul_rn(gi, gmul);
    g = (MODE == SSM_COORD_DIV) ? div_rn_const(g, norm, inv) : __fmul_rn(g, inv);
    return __fmul_rn(g, 2.0f);
}

struct Taps {
    float ix, iy, fx, fy;        // position and its floor
    float wnw, wne, wsw, wse;    // corner weights = area of the opposite sub-rectangle
    int off;                     // y0*W + x0 (only dereferenced under the masks)
    bool nw, ne, sw, se;         // corner inside the image (zeros padding otherwise)
};

template <int MODE>
__device__ __forceinline__ Taps make_taps(int x, int y, float u, float v, const Geom& g) {
    Taps t;
    float ix = sample_coord<MODE>((float)x, u, g.xnorm, g.xinv, g.xm1);
    float iy = sample_coord<MODE>((float)y, v, g.ynorm, g.yinv, g.ym1);
    // far outside (or NaN): every corner is out of bounds; keep the int conversion defined
    ix = fminf(fmaxf(ix, -2.0f), (float)g.W + 1.0f);
    iy = fminf(fmaxf(iy, -2.0f), (float)g.H + 1.0f);
    float fx = floorf(ix), fy = floorf(iy);
    float xe = fx + 1.0f, ys = fy + 1.0f;
    t.ix = ix; t.iy = iy; t.fx = fx; t.fy = fy;
    t.wnw = (xe - ix) * (ys - iy);
    t.wne = (ix - fx) * (ys - iy);
    t.wsw = (xe - ix) * (iy - fy);
    t.wse = (ix - fx) * (iy - fy);
    int x0 = (int)fx, y0 = (int)fy;
    bool xin0 = (unsigned)x0 < (unsigned)g.W, xin1 = (unsigned)(x0 + 1) < (unsigned)g.W;
    bool yin0 = (unsigned)y0 < (unsigned)g.H, yin1 = (unsigned)(y0 + 1) < (unsigned)g.H;
    t.nw = xin0 && yin0; t.ne = xin1 && yin0; t.sw = xin0 && yin1; t.se = xin1 && yin1;
    t.off = y0 * g.W + x0;
    return t;
}

// the four corner values of one plane (0 where masked)
struct Quad { float nw, ne, sw, se; };

template <typename T>
__device__ __forceinline__ Quad gather_quad(const T* __restrict__ plane, const Taps& t, int W) {
    const T* q = plane + t.off;
    Quad v;
    v.nw = t.nw ? ldg_(q) : 0.0f;
    v.ne = t.ne ? ldg_(q + 1) : 0.0f;
    v.sw = t.sw ? ldg_(q + W) : 0.0f;
    v.se = t.se ? ldg_(q + W + 1) : 0.0f;
    return v;
}

// bilinear value; masked corners have value 0 and therefore add exactly 0 (ATen skips them)
__device__ __forceinline__ float bilerp(const Quad& v, const Taps& t) {
    float acc = v.nw * t.wnw;
    acc = fmaf(v.ne, t.wne, acc);
    acc = fmaf(v.sw, t.wsw, acc);
    acc = fmaf(v.se, t.wse, acc);
    return acc;
}

// ---- packed frames -----------------------------------------------------------------------------
// The gathers are the L1-bound part of every kernel (ncu, profiles/r01a: 79 % L1tex throughput at
// 48 % DRAM): with the flows of the bench workload the 32 lanes of a warp hit ~17 different
// sectors per tap, and a planar NCHW frame needs one such request per tap AND channel.  A
// pre-pass (pack_frames_kernel) therefore re-lays each frame as pixel-interleaved RGBx (4 elements
// per pixel: 16 B in fp32, 8 B in bf16), so one request per tap brings all three channels.
template <typename T> __device__ __forceinline__ void load_px(const T* p, float (&v)[3]);
template <> __device__ __forceinline__ void load_px<float>(const float* p, float (&v)[3]) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = q.x; v[1] = q.y; v[2] = q.z;
}
template <> __device__ __forceinline__ void load_px<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[3]) {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = __uint_as_float(q.x << 16);
    v[1] = __uint_as_float(q.x & 0xffff0000u);
    v[2] = __uint_as_float(q.y << 16);
}

// corner values of the three colour planes of one frame.  PACKED: `frame` points at the RGBx copy
// (H*W*4 elements); otherwise at plane 0 of the planar frame with channel stride sc.
template <typename T, bool PACKED>
__device__ __forceinline__ void gather3(const T* __restrict__ frame, long long sc, const Taps& t, int W, Quad (&q)[3]) {
    if (PACKED) {
        const T* p = frame + (long long)t.off * 4;
        float a[3] = {0.f, 0.f, 0.f}, b[3] = {0.f, 0.f, 0.f}, c[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
        if (t.nw) load_px(p, a);
        if (t.ne) load_px(p + 4, b);
        if (t.sw) load_px(p + 4 * W, c);
        if (t.se) load_px(p + 4 * W + 4, d);
#pragma unroll
        for (int k = 0; k < 3; ++k) { q[k].nw = a[k]; q[k].ne = b[k]; q[k].sw = c[k]; q[k].se = d[k]; }
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) q[k] = gather_quad(frame + k * sc, t, W);
    }
}

// accumulate d(bilerp)/d(ix), d(bilerp)/d(iy) times upstream gradient g.  ATen sums the eight
// tap terms -nw*(ys-iy)*g + ne*(ys-iy)*g - ... one by one; grouping them by row / column is the
// same sum up to rounding (masked taps have value 0 either way) and takes a third of the
// instructions: d/dix = (ne-nw)*(ys-iy) + (se-sw)*(iy-fy), d/diy = (sw-nw)*(xe-ix) + (se-ne)*(ix-fx).
__device__ __forceinline__ void bilerp_grad(const Quad& v, const Taps& t, float g, float& gix, float& giy) {
    const float dys = t.iy - t.fy, dyn = (t.fy + 1.0f) - t.iy;   // weights of the south / north rows
    const float dxe = t.ix - t.fx, dxw = (t.fx + 1.0f) - t.ix;   // weights of the east / west columns
    const float dx = fmaf(v.ne - v.nw, dyn, (v.se - v.sw) * dys);
    const float dy = fmaf(v.sw - v.nw, dxw, (v.se - v.ne) * dxe);
    gix = fmaf(dx, g, gix);
    giy = fmaf(dy, g, giy);
}

// flow_interpolation.py:353,356 coefficients, evaluated as torch does on the B x 1 x 1 x 1 tensor
struct Coef { float c00, c01, c10, c11, omt, t; };
__device__ __forceinline__ Coef make_coef(float t) {
    Coef c;
    float omt = __fsub_rn(1.0f, t);
    c.c00 = __fmul_rn(-omt, t);
    c.c01 = __fmul_rn(t, t);
    c.c10 = __fmul_rn(omt, omt);
    c.c11 = __fmul_rn(t, omt);
    c.omt = omt; c.t = t;
    return c;
}

// F_t0 = -(1-t)t F01 + t^2 F10 ; F_t1 = (1-t)^2 F01 - t(1-t) F10, products rounded before the sum
__device__ __forceinline__ float est_t0(const Coef& c, float f01, float f10) {
    return __fadd_rn(__fmul_rn(c.c00, f01), __fmul_rn(c.c01, f10));
}
__device__ __forceinline__ float est_t1(const Coef& c, float f01, float f10) {
    return __fsub_rn(__fmul_rn(c.c10, f01), __fmul_rn(c.c11, f10));
}

// sigmoid and the final normalisation 1/Z are not coordinate arithmetic: the hardware approximations (MUFU.EX2,
// MUFU.RCP: relative error ~1e-7) are far inside the 1e-5 bar.  Round 1 used expf + __frcp_rn, which cost two
// subroutine calls and ~40 instructions per pixel and timestep in kernels that are bound by instruction issue.
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// 1 / (1 + 2^(-x log2 e)): two MUFU + two FP32 operations
__device__ __forceinline__ float sigmoid_(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }

// ---- bookkeeping of the deterministic image-gradient accumulation (see ssm_scatter.cuh) --------
struct ScatterHdr {          // lives at the start of the workspace, zeroed before every launch
    unsigned int absmax_bits;   // max |upstream gradient| as fp32 bits
    unsigned int pad[3];
};

// Must be reached by all 32 lanes of the warp.  m >= 0, or NaN whose bits compare above every
// finite value and poison the result.
__device__ __forceinline__ void record_absmax(unsigned int* slot, float m) {
#ifdef SSM_EXP_NO_ABSMAX       // timing experiment only (wrong scale): what the recording itself costs
    if (m < -1.0f) *slot = 0u;
    return;
#endif
    unsigned int bits = __reduce_max_sync(0xffffffffu, __float_as_uint(fabsf(m)));
    // The pre-check reads through L1 (ld.ca): a stale value is only ever too SMALL (the slot grows monotonically), which
    // costs a redundant atomicMax, never a wrong result.  Read from L2 (ld.cg) the one address took a request from every
    // warp of the launch -- 1 M requests on one L2 sector at 16 x 1088 x 1920, 1.5 ms of a 5.3 ms kernel
    // (profiles/r04g_bwd_timing_*.json).
    if ((threadIdx.x & 31) == 0 && bits > __ldca(slot)) atomicMax(slot, bits);
}

}  // namespace ssm
