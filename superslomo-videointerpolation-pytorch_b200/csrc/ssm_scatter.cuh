// ssm_scatter.cuh -- deterministic, segmented image-gradient accumulation for the warp backward (sm_100a).
//
// The image gradient of a backward warp is a scatter: every source pixel adds weight*grad to four
// data-dependent destination pixels.  The reference (ATen / cuDNN grid_sampler backward) does this
// with fp32 atomicAdd, whose result depends on the order in which the adds land and therefore
// changes from run to run.  Here every contribution is an INTEGER on one launch-wide grid, so any
// order of additions gives the same bits:
//
//   1. the gather kernels, which read the upstream gradient anyway, record max |grad| of the launch
//      (integer atomicMax on the bit pattern: order-independent);
//   2. scale = 2^k puts that maximum in [2^28, 2^29): a contribution rn(weight*grad) * 2^k (exact) is rounded
//      ONCE to a 32-bit integer, 29 bits below the largest possible one;
//   3. SEGMENTED accumulation (round 2): a CTA owns a 64 x 16 tile of source pixels and, per frame, a shared-memory
//      window of int32 accumulators covering the tile shifted by the displacement of its centre pixel plus a 12-pixel
//      halo.  Contributions that land in the window -- nearly all of them for a piecewise-smooth flow -- are added
//      with shared-memory integer atomics; the few that fall outside go to the 64-bit global accumulators directly.
//      The window SURVIVES the timesteps of a frame: it is flushed (one 64-bit global atomic per non-zero cell) and
//      re-anchored only when the centre displacement has drifted more than SW_DRIFT pixels from its origin, and at the
//      end of the frame -- for the flows of one pair at N = 7 that is once per frame instead of seven times.  int32
//      wrap-around is exact modulo 2^32 and every add checks its own overflow (the atomic returns the previous value),
//      correcting the global cell by +-2^32, so the sums are exact whatever the data;
//   4. a finalise pass converts the 64-bit sums back (and adds the direct, non-warped gradient terms).
//
// Measured (tools/exp_scatter.cu, profiles/r02k_exp_scatter.jsonl; 16 pairs x 2 frames x 7 timesteps at
// 1088x1920): global 64-bit atomics alone 19.5 ms (rough flow) / 10.1 ms (smooth); windowed 8.5 / 7.6 ms.
// The shipped kernels (tools/exp_bwd_timing.py, profiles/r04h-r04j): one flush per (frame, timestep) 10.9 / 9.7 ms,
// windows kept across the timesteps of a frame 7.8 / 7.3 ms (halo 12; halo 8: 8.2 / 7.6; a drift limit of 12, 16, 24 px
// or none: the same on these fields -- the limit is there for flows that sweep far over the N timesteps).
#pragma once
#include "ssm_kernels.cuh"

namespace ssm {

// power-of-two scale: the largest contribution (<= absmax, up to a rounding) lands in [2^28, 2^29].  64-bit global
// sums cannot overflow: the weights of one source pixel sum to 1, so a cell receives at most N*H*W * absmax < 2^25 * 2^29.
// count_bits is unused (kept in the signatures of the kernels that predate the windowed scheme).
__device__ __forceinline__ float scatter_scale(const ScatterHdr* h, int /*count_bits*/) {
    unsigned int bits = h->absmax_bits;
    if (bits == 0u) return 1.0f;
    if (bits >= 0x7f800000u) return __int_as_float(0x7fc00000);   // inf/NaN upstream: poison
    int e = (int)(bits >> 23) - 127;           // floor(log2(absmax)) (denormals: -127, fine)
    int k = 28 - e;
    k = min(k, 126); k = max(k, -126);
    return __int_as_float((k + 127) << 23);
}

__device__ __forceinline__ void fx_add(long long* dst, float contrib, float scale) {
    const int q = __float2int_rn(contrib * scale);       // |contrib * scale| <= 2^29; NaN -> 0 (the scale poisons the result)
    if (q != 0) atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)(long long)q);
}

// ---- the shared-memory window -----------------------------------------------------------------------------------
#ifndef SSM_SW_TILE_H
#define SSM_SW_TILE_H 16
#endif
constexpr int SW_TILE_W = 64, SW_TILE_H = SSM_SW_TILE_H;  // source pixels per CTA: 256 threads x (SW_TILE_H / 4) rows
#ifndef SSM_SW_HALO
#define SSM_SW_HALO 12
#endif
constexpr int SW_HALO = SSM_SW_HALO;
#ifndef SSM_SW_DRIFT
#define SSM_SW_DRIFT SSM_SW_HALO       // re-anchor a frame's window when its centre displacement has moved further than this
#endif
constexpr int SW_DRIFT = SSM_SW_DRIFT;
constexpr int SW_W = SW_TILE_W + 2 * SW_HALO, SW_H = SW_TILE_H + 2 * SW_HALO, SW_PLANE = SW_W * SW_H;   // 88 x 40 cells x 3 planes: 42 KB
constexpr int SW_THREADS = 256;
#ifndef SSM_SW_MIN_BLOCKS
#define SSM_SW_MIN_BLOCKS 4        // 64 registers: 7.8 ms; 3 CTAs/SM (80 registers) 9.2 ms, 5 (48, spilling) 8.8 ms (profiles/r04o); 2: 14.5 -> 9.9 ms in r02p
#endif
constexpr int SW_MIN_BLOCKS = SSM_SW_MIN_BLOCKS;
constexpr int SW_CENTRE_LX = SW_TILE_W / 2, SW_CENTRE_ROW = SW_TILE_H / 2;

struct SwTile { int b, x0, y0, lx, ly; };
__device__ __forceinline__ SwTile sw_tile(int H, int W) {
    const int tiles_x = (W + SW_TILE_W - 1) / SW_TILE_W, tiles_y = (H + SW_TILE_H - 1) / SW_TILE_H, tpp = tiles_x * tiles_y;
    SwTile t;
    t.b = blockIdx.x / tpp;
    const int r = blockIdx.x - t.b * tpp, ty = r / tiles_x, tx = r - ty * tiles_x;
    t.x0 = tx * SW_TILE_W; t.y0 = ty * SW_TILE_H;
    t.lx = threadIdx.x & (SW_TILE_W - 1); t.ly = threadIdx.x / SW_TILE_W;        // pixel j of a thread: row ly + 4 j
    return t;
}

// one int32 add into the window; an overflow of the cell (possible only after more than two maximal contributions)
// is repaired exactly by moving 2^32 into the 64-bit global cell
__device__ __forceinline__ void sw_add(int* cell, int q, long long* gcell) {
    const int old = atomicAdd(cell, q);
    const int nw = (int)((unsigned)old + (unsigned)q);
    if (((old ^ nw) & (q ^ nw)) < 0)
        atomicAdd(reinterpret_cast<unsigned long long*>(gcell), (unsigned long long)(q > 0 ? (1ll << 32) : -(1ll << 32)));
}

// the contributions of one source pixel (taps t, three channel values gv) to planes plane0 + c * npx.
// ONE code path for every pixel: a tap inside the image goes to the window when its cell is inside it and to the global
// accumulator otherwise, by predication.  (The first version had a fast path for footprints entirely inside image and
// window and a per-tap slow path; with a rough flow field most warps hold both kinds of lanes and executed both paths,
// 550 instructions per pixel and phase.)  All twelve shared-memory atomics are issued before any of their results is
// looked at, and one combined test decides whether some cell may have wrapped.
__device__ __forceinline__ void sw_splat3(int* win, int ox, int oy, const Taps& t, const float (&gv)[3], float scale,
                                          long long* __restrict__ plane0, long long npx, int W) {
    const int cx = (int)t.fx - ox, cy = (int)t.fy - oy;
    const float w[4] = {t.wnw, t.wne, t.wsw, t.wse};
    const bool in[4] = {t.nw, t.ne, t.sw, t.se};
    const bool xin[2] = {(unsigned)cx < (unsigned)SW_W, (unsigned)(cx + 1) < (unsigned)SW_W};
    const bool yin[2] = {(unsigned)cy < (unsigned)SW_H, (unsigned)(cy + 1) < (unsigned)SW_H};
    bool smem[4], glob[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool inside = xin[k & 1] && yin[k >> 1];
        smem[k] = in[k] && inside;
        glob[k] = in[k] && !inside;
    }
    int* c0 = win + cy * SW_W + cx;                 // dereferenced under smem[k] only
    long long* g0 = plane0 + t.off;                 // dereferenced under in[k] only
    const float gs[3] = {gv[0] * scale, gv[1] * scale, gv[2] * scale};      // scale is a power of two: exact
    int q[12], old[12];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) q[4 * c + k] = __float2int_rn(w[k] * gs[c]);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = 4 * c + k;
            old[i] = 0;
            if (smem[k]) old[i] = atomicAdd(c0 + c * SW_PLANE + (k & 1) + (k >> 1) * SW_W, q[i]);
        }
    if (glob[0] || glob[1] || glob[2] || glob[3]) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!glob[k]) continue;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (q[4 * c + k])
                    atomicAdd(reinterpret_cast<unsigned long long*>(g0 + c * npx + (k & 1) + (k >> 1) * W), (unsigned long long)(long long)q[4 * c + k]);
        }
    }
    // |q| <= 2^29 (+ a rounding): a cell can only wrap if it held more than 2^30 in magnitude -- one add + one or per
    // atomic, with a wide margin
    unsigned flag = 0u;
#pragma unroll
    for (int i = 0; i < 12; ++i) flag |= (unsigned)old[i] + 0x40000000u;
    if (flag & 0x80000000u) {            // rare: repair the cells that wrapped by moving 2^32 into the global cell
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = 4 * c + k, nw = (int)((unsigned)old[i] + (unsigned)q[i]);
                if (smem[k] && ((old[i] ^ nw) & (q[i] ^ nw)) < 0)
                    atomicAdd(reinterpret_cast<unsigned long long*>(g0 + c * npx + (k & 1) + (k >> 1) * W),
                              (unsigned long long)(q[i] > 0 ? (1ll << 32) : -(1ll << 32)));
            }
    }
}

// non-zero cells -> one 64-bit global atomic each; the window is left zeroed for the next phase.  Four cells per
// 16-byte shared-memory load; an all-zero quadruple (half the window on average) costs three instructions.
__device__ __forceinline__ void sw_flush(int* win, int ox, int oy, long long* __restrict__ plane0, long long npx, int H, int W) {
    static_assert(SW_W % 4 == 0, "window rows are read four cells at a time");
    int4* w4 = reinterpret_cast<int4*>(win);
    for (int i = threadIdx.x; i < 3 * SW_PLANE / 4; i += SW_THREADS) {
        const int4 v = w4[i];
        if ((v.x | v.y | v.z | v.w) == 0) continue;
        w4[i] = make_int4(0, 0, 0, 0);
        const int cell = 4 * i, c = cell / SW_PLANE, r = cell - c * SW_PLANE, cy = r / SW_W, cx = r - cy * SW_W;
        // only taps inside the image were added: a non-zero cell (ox + cx + j, oy + cy) is a valid pixel
        long long* g = plane0 + c * npx + (long long)(oy + cy) * W + (ox + cx);
        if (v.x) atomicAdd(reinterpret_cast<unsigned long long*>(g), (unsigned long long)(long long)v.x);
        if (v.y) atomicAdd(reinterpret_cast<unsigned long long*>(g + 1), (unsigned long long)(long long)v.y);
        if (v.z) atomicAdd(reinterpret_cast<unsigned long long*>(g + 2), (unsigned long long)(long long)v.z);
        if (v.w) atomicAdd(reinterpret_cast<unsigned long long*>(g + 3), (unsigned long long)(long long)v.w);
    }
}

// One phase = the contributions of the CTA's 64 x 16 source pixels to one frame's three planes.
// (win and s_org are deliberately NOT __restrict__: with it the compiler hoists the s_org load above the barrier.)
// pixel(x, y, t, gv) fills the taps and the three channel values of source pixel (x, y) of this phase.
template <typename PixelFn>
__device__ __forceinline__ void sw_phase(int* win, volatile int* s_org, const SwTile& ti, const Geom& g, float scale,
                                         long long* __restrict__ plane0, long long npx, PixelFn&& pixel) {
    // window origin: the tile shifted by the displacement of its centre pixel (any origin is correct; this one
    // captures the most)
    if (ti.lx == SW_CENTRE_LX && ti.ly == (SW_CENTRE_ROW & 3)) {
        const int x = min(ti.x0 + SW_CENTRE_LX, g.W - 1), y = min(ti.y0 + SW_CENTRE_ROW, g.H - 1);
        Taps t; float gv[3];
        pixel(x, y, t, gv);
        s_org[0] = (int)t.fx - (x - ti.x0) - SW_HALO;
        s_org[1] = (int)t.fy - (y - ti.y0) - SW_HALO;
    }
    __syncthreads();                         // origin visible; the previous phase's flush is complete
    const int ox = s_org[0], oy = s_org[1];
#pragma unroll 1
    for (int j = 0; j < SW_TILE_H / 4; ++j) {
        const int x = ti.x0 + ti.lx, y = ti.y0 + ti.ly + 4 * j;
        if (x < g.W && y < g.H) {
            Taps t; float gv[3];
            pixel(x, y, t, gv);
            sw_splat3(win, ox, oy, t, gv, scale, plane0, npx, g.W);
        }
    }
    __syncthreads();
    sw_flush(win, ox, oy, plane0, npx, g.H, g.W);
    // no barrier here: the next phase's first barrier orders this flush before its adds, and every thread has read
    // s_org before the barrier above
}

// One timestep of a frame whose window is kept across timesteps (compute_inputs / compute_output_image backward): the
// window is flushed and re-anchored only when the displacement of the tile's centre pixel has drifted more than SW_DRIFT
// pixels from the window's origin (first timestep: anchored).  s_org (4 ints) is double-buffered by the timestep parity;
// `have` / (ox, oy) are uniform over the CTA.  The caller flushes after the frame's last timestep.
template <typename PixelFn>
__device__ __forceinline__ void sw_timestep(int* win, volatile int* s_org, int parity, const SwTile& ti, const Geom& g, float scale,
                                              long long* __restrict__ plane0, long long npx, bool& have, int& ox, int& oy,
                                              PixelFn&& pixel) {
    volatile int* so = s_org + 2 * parity;
    if (ti.lx == SW_CENTRE_LX && ti.ly == (SW_CENTRE_ROW & 3)) {
        const int x = min(ti.x0 + SW_CENTRE_LX, g.W - 1), y = min(ti.y0 + SW_CENTRE_ROW, g.H - 1);
        Taps t; float gv[3];
        pixel(x, y, t, gv);
        so[0] = (int)t.fx - (x - ti.x0) - SW_HALO;
        so[1] = (int)t.fy - (y - ti.y0) - SW_HALO;
    }
    __syncthreads();                         // candidate origin visible; the previous phase's adds are complete
    const int cx = so[0], cy = so[1];
    if (!have || abs(cx - ox) > SW_DRIFT || abs(cy - oy) > SW_DRIFT) {
        if (have) {
            sw_flush(win, ox, oy, plane0, npx, g.H, g.W);
            __syncthreads();                 // the window is zero again before anything is added at the new origin
        }
        ox = cx; oy = cy; have = true;
    }
#pragma unroll 1
    for (int j = 0; j < SW_TILE_H / 4; ++j) {
        const int x = ti.x0 + ti.lx, y = ti.y0 + ti.ly + 4 * j;
        if (x < g.W && y < g.H) {
            Taps t; float gv[3];
            pixel(x, y, t, gv);
            sw_splat3(win, ox, oy, t, gv, scale, plane0, npx, g.W);
        }
    }
}

__device__ __forceinline__ void sw_clear(int* win) {
    int4* w4 = reinterpret_cast<int4*>(win);
    for (int i = threadIdx.x; i < 3 * SW_PLANE / 4; i += SW_THREADS) w4[i] = make_int4(0, 0, 0, 0);
}

__device__ __forceinline__ void scatter_quad(long long* plane, const Taps& t, int W, float gv, float scale) {
    long long* q = plane + t.off;
    if (t.nw) fx_add(q, t.wnw * gv, scale);
    if (t.ne) fx_add(q + 1, t.wne * gv, scale);
    if (t.sw) fx_add(q + W, t.wsw * gv, scale);
    if (t.se) fx_add(q + W + 1, t.wse * gv, scale);
}

// ---- max |grad| pre-pass for the stand-alone warp when only the image gradient is wanted ------
template <typename T>
__global__ void __launch_bounds__(TILE_THREADS)
absmax_kernel(View<const T> v, int C, Geom g, ScatterHdr* hdr) {
    TileIdx ti = tile_index(g.H, g.W);
    float m = 0.0f;
    if (ti.valid) {
        const T* p = v.p + ti.b * v.sb + ti.y * g.W + ti.x;
        for (int c = 0; c < C; ++c) m = fmaxf(m, fabsf(lds_(p + c * v.sc)));
    }
    record_absmax(&hdr->absmax_bits, m);
}

// ---- a1: scatter of grad_out through the flow ---------------------------------------------------
// any channel count: global atomics only (the window scheme below is built for the three colour planes)
template <typename T, int MODE>
__global__ void __launch_bounds__(TILE_THREADS)
warp_scatter_kernel(View<const T> gout, View<const T> flow, long long* __restrict__ acc, int C, Geom g,
                    const ScatterHdr* __restrict__ hdr, int count_bits) {
    TileIdx ti = tile_index(g.H, g.W);
    if (!ti.valid) return;
    const float scale = scatter_scale(hdr, count_bits);
    const int p = ti.y * g.W + ti.x;
    const long long npx = (long long)g.H * g.W;
    const T* fl = flow.p + ti.b * flow.sb + p;
    Taps t = make_taps<MODE>(ti.x, ti.y, lds_(fl), lds_(fl + flow.sc), g);
    const T* gp = gout.p + ti.b * gout.sb + p;
    for (int c = 0; c < C; ++c)
        scatter_quad(acc + ((long long)ti.b * C + c) * npx, t, g.W, lds_(gp + c * gout.sc), scale);
}

// C = 3: windowed
template <typename T, int MODE>
__global__ void __launch_bounds__(SW_THREADS, SW_MIN_BLOCKS)
warp_scatter_win_kernel(View<const T> gout, View<const T> flow, long long* __restrict__ acc, Geom g,
                        const ScatterHdr* __restrict__ hdr) {
    __shared__ __align__(16) int win[3 * SW_PLANE];
    __shared__ int s_org[2];
    const SwTile ti = sw_tile(g.H, g.W);
    const float scale = scatter_scale(hdr, 0);
    const long long npx = (long long)g.H * g.W;
    sw_clear(win);
    const T* fl = flow.p + ti.b * flow.sb;
    const T* gp = gout.p + ti.b * gout.sb;
    const int fsc = (int)flow.sc, gsc = (int)gout.sc;
    sw_phase(win, s_org, ti, g, scale, acc + (long long)ti.b * 3 * npx, npx, [&](int x, int y, Taps& t, float (&gv)[3]) {
        const int p = y * g.W + x;
        t = make_taps<MODE>(x, y, lds_(fl + p), lds_(fl + fsc + p), g);
#pragma unroll
        for (int c = 0; c < 3; ++c) gv[c] = lds_(gp + c * gsc + p);
    });
}

// ---- a2: scatter of grad16[:, 3:6] through F_t1 into I1 and grad16[:, 10:13] through F_t0 into I0
template <typename T, int MODE>
__global__ void __launch_bounds__(SW_THREADS, SW_MIN_BLOCKS)
flow_pack_scatter_kernel(View<const T> g16, View<const T> flow4, const float* __restrict__ tv,
                         long long* __restrict__ acc, int N, Geom g,
                         const ScatterHdr* __restrict__ hdr, int count_bits) {
    __shared__ __align__(16) int win[3 * SW_PLANE];
    __shared__ int s_org[4];
    const SwTile ti = sw_tile(g.H, g.W);
    const float scale = scatter_scale(hdr, count_bits);
    const long long npx = (long long)g.H * g.W;
    sw_clear(win);
    const T* F = flow4.p + ti.b * flow4.sb;
    const int fsc = (int)flow4.sc, gsc = (int)g16.sc;
    long long* a0 = acc + (long long)ti.b * 6 * npx;      // I0 planes 0-2, I1 planes 3-5
    for (int frame = 0; frame < 2; ++frame) {
        bool have = false; int ox = 0, oy = 0;
        long long* plane0 = a0 + frame * 3 * npx;
        for (int n = 0; n < N; ++n) {
            const Coef k = make_coef(__ldg(tv + ti.b * N + n));
            const T* G = g16.p + ti.b * g16.sb + n * g16.sn;
            sw_timestep(win, s_org, n & 1, ti, g, scale, plane0, npx, have, ox, oy, [&](int x, int y, Taps& t, float (&gv)[3]) {
                const int p = y * g.W + x;
                const float f01x = ldg_(F + p), f01y = ldg_(F + fsc + p), f10x = ldg_(F + 2 * fsc + p), f10y = ldg_(F + 3 * fsc + p);
                if (frame) t = make_taps<MODE>(x, y, storage_round<T>(est_t1(k, f01x, f10x)), storage_round<T>(est_t1(k, f01y, f10y)), g);
                else t = make_taps<MODE>(x, y, storage_round<T>(est_t0(k, f01x, f10x)), storage_round<T>(est_t0(k, f01y, f10y)), g);
                const T* gp = G + (frame ? 3 : 10) * gsc + p;
#pragma unroll
                for (int c = 0; c < 3; ++c) gv[c] = lds_(gp + c * gsc);
            });
        }
        __syncthreads();                     // every add of the frame's last timestep is in the window
        if (have) sw_flush(win, ox, oy, plane0, npx, g.H, g.W);
    }
}

// ---- a4: scatter of d/d(warped I0), d/d(warped I1) through the refined flows --------------------
// d/d(warped I_f)_c = k_f V_f G_c (see fuse_bwd_kernel) is recomputed here from G = grad of the fused frame and the
// visibility logit with the arithmetic of fuse_bwd_kernel, instead of being staged by that kernel.
template <typename T, int MODE, bool RECOMP>
__global__ void __launch_bounds__(SW_THREADS, SW_MIN_BLOCKS)
fuse_scatter_kernel(View<const T> g3, View<const T> flows4, View<const T> out5,
                    const float* __restrict__ tv, long long* __restrict__ acc, int N, Geom g,
                    const ScatterHdr* __restrict__ hdr, int count_bits) {
    __shared__ __align__(16) int win[3 * SW_PLANE];
    __shared__ int s_org[4];
    const SwTile ti = sw_tile(g.H, g.W);
    const float scale = scatter_scale(hdr, count_bits);
    const long long npx = (long long)g.H * g.W;
    sw_clear(win);
    long long* a0 = acc + (long long)ti.b * 6 * npx;
    const int fsc = (int)flows4.sc, ysc = (int)out5.sc, gsc = (int)g3.sc;
    for (int frame = 0; frame < 2; ++frame) {
        bool have = false; int ox = 0, oy = 0;
        long long* plane0 = a0 + frame * 3 * npx;
        for (int n = 0; n < N; ++n) {
            const float tt = __ldg(tv + ti.b * N + n);
            const float omt = __fsub_rn(1.0f, tt);
            const T* X = flows4.p + ti.b * flows4.sb + (RECOMP ? 0 : n * flows4.sn);
            const T* Y = out5.p + ti.b * out5.sb + n * out5.sn;
            const T* G = g3.p + ti.b * g3.sb + n * g3.sn;
            sw_timestep(win, s_org, n & 1, ti, g, scale, plane0, npx, have, ox, oy, [&](int x, int y, Taps& t, float (&gv)[3]) {
                const int p = y * g.W + x;
                float xs[4];
                if (RECOMP) {
                    const float f[4] = {ldg_(X + p), ldg_(X + fsc + p), ldg_(X + 2 * fsc + p), ldg_(X + 3 * fsc + p)};
                    est_flows<T>(tt, f, xs);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) xs[k] = ldg_(X + k * fsc + p);
                }
                const int o = frame ? 0 : 2;
                const float fx = __fadd_rn(xs[o], ldg_(Y + (1 + o) * ysc + p));
                const float fy = __fadd_rn(xs[o + 1], ldg_(Y + (2 + o) * ysc + p));
                t = make_taps<MODE>(x, y, fx, fy, g);
                const float v1 = sigmoid_(ldg_(Y + p));
                const float v0 = 1.0f - v1;
                const float rz = rcp_approx(omt * v0 + tt * v1);
                const float kv = frame ? (tt * rz) * v1 : (omt * rz) * v0;            // k_f V_f
#pragma unroll
                for (int c = 0; c < 3; ++c) gv[c] = kv * ldg_(G + c * gsc + p);
            });
        }
        __syncthreads();                     // every add of the frame's last timestep is in the window
        if (have) sw_flush(win, ox, oy, plane0, npx, g.H, g.W);
        // the next frame's first barrier orders this flush before its adds
    }
}

// ---- finalise: fixed point -> storage type, plus the direct (non-warped) gradient if any --------
// One CTA converts 1024 consecutive pixels of one plane, four per thread (32 bytes of accumulator in, 16 bytes of fp32
// out); plane and chunk come from ONE 32-bit division per CTA.  (The first version was a grid-stride loop with two 64-bit
// divisions per element: 0.83 ms for the 1.6 GB of accumulators of 16 pairs, 2.9 TB/s.)
constexpr int FIN_THREADS = 256, FIN_PER_THREAD = 4;
__device__ __forceinline__ void store4(float* o, const float (&r)[4]) { __stcs(reinterpret_cast<float4*>(o), make_float4(r[0], r[1], r[2], r[3])); }
__device__ __forceinline__ void store4(__nv_bfloat16* o, const float (&r)[4]) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(r[0], r[1]), hi = __floats2bfloat162_rn(r[2], r[3]);
    __stcs(reinterpret_cast<uint2*>(o), make_uint2(*reinterpret_cast<const unsigned*>(&lo), *reinterpret_cast<const unsigned*>(&hi)));
}
template <typename T>
__global__ void __launch_bounds__(FIN_THREADS)
scatter_finalize_kernel(const long long* __restrict__ acc, const float* __restrict__ direct,
                        View<T> gimg, int C, long long npx, unsigned chunks_per_plane,
                        const ScatterHdr* __restrict__ hdr, int count_bits) {
    const float scale = scatter_scale(hdr, count_bits);
    const double inv = 1.0 / (double)scale;   // NaN scale poisons the output, as intended
    const unsigned plane = blockIdx.x / chunks_per_plane, chunk = blockIdx.x - plane * chunks_per_plane;
    const unsigned b = plane / (unsigned)C, c = plane - b * (unsigned)C;
    const long long p0 = ((long long)chunk * FIN_THREADS + threadIdx.x) * FIN_PER_THREAD;
    if (p0 >= npx) return;
    const long long* a = acc + (long long)plane * npx + p0;
    const float* d = direct ? direct + (long long)plane * npx + p0 : nullptr;
    T* o = gimg.p + b * gimg.sb + c * gimg.sc + p0;
    const bool whole = p0 + FIN_PER_THREAD <= npx && (reinterpret_cast<uintptr_t>(a) & 15) == 0 &&
                       (!d || (reinterpret_cast<uintptr_t>(d) & 15) == 0) &&
                       (reinterpret_cast<uintptr_t>(o) & (FIN_PER_THREAD * sizeof(T) - 1)) == 0;
    if (whole) {
        const longlong2 v0 = __ldcs(reinterpret_cast<const longlong2*>(a)), v1 = __ldcs(reinterpret_cast<const longlong2*>(a) + 1);
        float r[4] = {(float)((double)v0.x * inv), (float)((double)v0.y * inv), (float)((double)v1.x * inv), (float)((double)v1.y * inv)};
        if (d) {
            const float4 dd = __ldcs(reinterpret_cast<const float4*>(d));
            r[0] += dd.x; r[1] += dd.y; r[2] += dd.z; r[3] += dd.w;
        }
        store4(o, r);
    } else {
        for (int k = 0; k < FIN_PER_THREAD && p0 + k < npx; ++k) {
            float v = (float)((double)__ldcs(a + k) * inv);
            if (d) v += __ldcs(d + k);
            sts_(o + k, v);
        }
    }
}

}  // namespace ssm
