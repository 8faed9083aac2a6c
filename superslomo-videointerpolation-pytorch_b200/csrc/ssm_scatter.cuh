// ssm_scatter.cuh -- deterministic image-gradient accumulation for the warp backward (sm_100a).
//
// The image gradient of a backward warp is a scatter: every source pixel adds weight*grad to four
// data-dependent destination pixels.  The reference (ATen / cuDNN grid_sampler backward) does this
// with fp32 atomicAdd, whose result depends on the order in which the adds land and therefore
// changes from run to run.  Here the destination accumulators are 64-bit fixed point:
//
//   1. the gather kernels, which read the upstream gradient anyway, record max |grad| of the launch
//      (integer atomicMax on the bit pattern: order-independent);
//   2. scale = 2^k is chosen from that maximum and the largest possible number of addends per
//      destination (N*H*W) so that no sum can overflow 63 bits; k leaves >= 34 fraction bits
//      below the largest contribution at 1080p x 7 timesteps;
//   3. every contribution rn(weight*grad) is multiplied by 2^k (exact), rounded to an integer once
//      and added with an integer atomic -- integer addition is associative, so the sum is
//      bit-identical run to run whatever the order;
//   4. a finalise pass converts the sum back (and adds the direct, non-warped gradient terms).
//
// Each destination plane segment is therefore accumulated exactly (to 2^-k) and deterministically;
// the only rounding is the one fp32 product per contribution, which the reference has as well.
#pragma once
#include "ssm_kernels.cuh"

namespace ssm {

// power-of-two scale: the largest contribution (<= absmax) lands below 2^(62 - count_bits)
__device__ __forceinline__ float scatter_scale(const ScatterHdr* h, int count_bits) {
    unsigned int bits = h->absmax_bits;
    if (bits == 0u) return 1.0f;
    if (bits >= 0x7f800000u) return __int_as_float(0x7fc00000);   // inf/NaN upstream: poison
    int e = (int)(bits >> 23) - 127;           // floor(log2(absmax)) (denormals: -127, fine)
    int k = (62 - count_bits) - (e + 1);
    k = min(k, 126); k = max(k, -126);
    return __int_as_float((k + 127) << 23);
}

__device__ __forceinline__ void fx_add(long long* dst, float contrib, float scale) {
    long long q = __float2ll_rn(contrib * scale);
    if (q != 0) atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)q);
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

// ---- a2: scatter of grad16[:, 3:6] through F_t1 into I1 and grad16[:, 10:13] through F_t0 into I0
template <typename T, int MODE>
__global__ void __launch_bounds__(TILE_THREADS)
flow_pack_scatter_kernel(View<const T> g16, View<const T> flow4, const float* __restrict__ tv,
                         long long* __restrict__ acc, int N, Geom g,
                         const ScatterHdr* __restrict__ hdr, int count_bits) {
    TileIdx ti = tile_index(g.H, g.W);
    if (!ti.valid) return;
    const float scale = scatter_scale(hdr, count_bits);
    const int p = ti.y * g.W + ti.x;
    const long long npx = (long long)g.H * g.W;
    const T* F = flow4.p + ti.b * flow4.sb + p;
    const float f01x = lds_(F), f01y = lds_(F + flow4.sc);
    const float f10x = lds_(F + 2 * flow4.sc), f10y = lds_(F + 3 * flow4.sc);
    long long* a0 = acc + (long long)ti.b * 6 * npx;      // I0 planes 0-2, I1 planes 3-5
    for (int n = 0; n < N; ++n) {
        const Coef k = make_coef(__ldg(tv + ti.b * N + n));
        const Taps t1 = make_taps<MODE>(ti.x, ti.y, storage_round<T>(est_t1(k, f01x, f10x)), storage_round<T>(est_t1(k, f01y, f10y)), g);
        const Taps t0 = make_taps<MODE>(ti.x, ti.y, storage_round<T>(est_t0(k, f01x, f10x)), storage_round<T>(est_t0(k, f01y, f10y)), g);
        const T* G = g16.p + ti.b * g16.sb + n * g16.sn + p;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            scatter_quad(a0 + (3 + c) * npx, t1, g.W, lds_(G + (3 + c) * g16.sc), scale);
            scatter_quad(a0 + c * npx, t0, g.W, lds_(G + (10 + c) * g16.sc), scale);
        }
    }
}

// ---- a4: scatter of the staged d/d(warped I0), d/d(warped I1) through the refined flows ---------
template <typename T, int MODE, bool RECOMP>
__global__ void __launch_bounds__(TILE_THREADS)
fuse_scatter_kernel(const float* __restrict__ stage, View<const T> flows4, View<const T> out5,
                    const float* __restrict__ tv, long long* __restrict__ acc, int N, Geom g,
                    const ScatterHdr* __restrict__ hdr, int count_bits) {
    TileIdx ti = tile_index(g.H, g.W);
    if (!ti.valid) return;
    const float scale = scatter_scale(hdr, count_bits);
    const int p = ti.y * g.W + ti.x;
    const long long npx = (long long)g.H * g.W;
    long long* a0 = acc + (long long)ti.b * 6 * npx;
    float f[4] = {0.f, 0.f, 0.f, 0.f};
    if (RECOMP) {
        const T* F = flows4.p + ti.b * flows4.sb + p;
#pragma unroll
        for (int k = 0; k < 4; ++k) f[k] = lds_(F + k * flows4.sc);
    }
    for (int n = 0; n < N; ++n) {
        float xs[4];
        if (RECOMP) {
            est_flows<T>(__ldg(tv + ti.b * N + n), f, xs);
        } else {
            const T* X = flows4.p + ti.b * flows4.sb + n * flows4.sn + p;
#pragma unroll
            for (int k = 0; k < 4; ++k) xs[k] = lds_(X + k * flows4.sc);
        }
        const T* Y = out5.p + ti.b * out5.sb + n * out5.sn + p;
        const float f1x = __fadd_rn(xs[0], lds_(Y + out5.sc));
        const float f1y = __fadd_rn(xs[1], lds_(Y + 2 * out5.sc));
        const float f0x = __fadd_rn(xs[2], lds_(Y + 3 * out5.sc));
        const float f0y = __fadd_rn(xs[3], lds_(Y + 4 * out5.sc));
        const Taps t0 = make_taps<MODE>(ti.x, ti.y, f0x, f0y, g);
        const Taps t1 = make_taps<MODE>(ti.x, ti.y, f1x, f1y, g);
        const float* st = stage + ((long long)(ti.b * N + n) * 6) * npx + p;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            scatter_quad(a0 + c * npx, t0, g.W, __ldcs(st + c * npx), scale);
            scatter_quad(a0 + (3 + c) * npx, t1, g.W, __ldcs(st + (3 + c) * npx), scale);
        }
    }
}

// ---- finalise: fixed point -> storage type, plus the direct (non-warped) gradient if any --------
template <typename T>
__global__ void __launch_bounds__(256)
scatter_finalize_kernel(const long long* __restrict__ acc, const float* __restrict__ direct,
                        View<T> gimg, int C, long long npx, long long total,
                        const ScatterHdr* __restrict__ hdr, int count_bits) {
    const float scale = scatter_scale(hdr, count_bits);
    const double inv = 1.0 / (double)scale;   // NaN scale poisons the output, as intended
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long bc = i / npx, p = i - bc * npx;
        long long b = bc / C, c = bc - b * C;
        float v = (float)((double)__ldcs(acc + i) * inv);
        if (direct) v += __ldcs(direct + i);
        sts_(gimg.p + b * gimg.sb + c * gimg.sc + p, v);
    }
}

}  // namespace ssm
