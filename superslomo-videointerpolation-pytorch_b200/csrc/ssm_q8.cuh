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
// flow_interpolation.py:364-367) are bit-exact with the reference's normalisation either way they are obtained: read
// from the planar normalised frames (ssm_frames_from_u8), or looked up from the tables' own bytes through the 3 x 256
// table that kernel applies (lut != NULL: the planar frames are then not read at all).
//
// One thread owns TWO horizontally adjacent pixels: every streaming access is 8 bytes per lane (a warp moves
// 256 contiguous bytes per plane row: 6.96 TB/s instead of 5.70 TB/s for the 16-plane store pattern of
// compute_inputs, profiles/r02b_exp_store.jsonl).  W must be even (padded sizes are multiples of 32).
#pragma once
#include "ssm_frames.cuh"

// Interpolation in difference form (12 packed operations fewer per timestep, errors 1.2e-6 instead of 2.6e-6): measured
// SLOWER for the headline kernel (compute_output_image, fp32 U-Net output: 1.85 -> 1.94 ms, longer dependency chains at
// 24 warps per SM), faster only with a bf16 U-Net output (2.14 -> 1.94 ms); off (profiles/r03s_q8_timing_*.json).
#ifndef SSM_Q8_DIFF_FORM
#define SSM_Q8_DIFF_FORM 0
#endif
// Experiment (off): with bf16 storage, sample at x + flow instead of reproducing the reference's normalise / de-normalise
// round trip bit for bit (40 of ~140 packed operations per timestep).  bf16 path 4.16 -> 3.92 ms (0.46 -> 0.49 of the HBM
// peak), 3.87 ms with the difference form on top: the bf16 kernels are not bound by that arithmetic either, and the
// coordinate modes would stop meaning anything for one dtype -- not shipped (profiles/r04f_bf16_*.json, tools/gpu_exp16.sh).
#ifndef SSM_Q8_BF16_DIRECT_COORD
#define SSM_Q8_BF16_DIRECT_COORD 0
#endif

namespace ssm {

struct Norm3 { float a[3], c[3]; };       // normalised value of byte b in channel k: a[k] * b + c[k]

constexpr int Q8_TILE_W = 64;              // 32 lanes x 2 pixels
// rows per CTA (threads = 32 x rows).  64 x 4 tiles with 128-thread CTAs, 6 per SM: compute_output_image 1.85 -> 1.81 ms,
// compute_inputs 2.86 -> 2.91 ms -- a wash (profiles/r03u_q8_timing_128_thread_ctas_*.json)
#ifndef SSM_Q8_TILE_H
#define SSM_Q8_TILE_H 8
#endif
constexpr int Q8_TILE_H = SSM_Q8_TILE_H;
constexpr int Q8_THREADS = 32 * Q8_TILE_H;


// ---- two pixels at a time: packed fp32 arithmetic ------------------------------------------------
// sm_100 issues one instruction for two independent fp32 operations on a register pair (FADD2 / FMUL2 /
// FFMA2, each with its own IEEE rounding -- the same bits as two scalar instructions, never contracted).
// A thread owns two horizontally adjacent pixels, so every quantity of the sampling arithmetic is a natural
// pair (.x = pixel x, .y = pixel x + 1) and the kernels, which were bound by instruction issue (75-82 % of
// the issue slots, profiles/r02e), need about half the instructions for it.
typedef float2 f2;
__device__ __forceinline__ f2 bc2(float s) { return make_float2(s, s); }
__device__ __forceinline__ f2 neg2(f2 a) { return make_float2(-a.x, -a.y); }
// add2 / mul2 / fma2: free to be contracted by the compiler (weights, interpolation, fusion)
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __fadd2_rn(a, neg2(b)); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
// add2x / mul2x / add2s: the coordinate arithmetic, one IEEE rounding per operation.  Two traps (both seen in the SASS):
// the __f*2_rn intrinsics, unlike their scalar namesakes, lower to plain vector fmul / fadd, which nvcc contracts
// into FFMA2; and ptxas (12.9) fuses even an explicit mul.rn.f32x2 feeding an add.rn.f32x2 into FFMA2, with or
// without -fmad=false (it also rewrites fma(x, 1, y) to that add first).  So: PTX with explicit .rn for the packed
// operations, and wherever a packed PRODUCT is the operand of a SUM that the reference rounds separately, the sum is
// two scalar add.rn.f32 (add2s), which ptxas never fuses.
__device__ __forceinline__ unsigned long long f2_bits(f2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ f2 bits_f2(unsigned long long r) {
    f2 a;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
    return a;
}
__device__ __forceinline__ f2 add2x(f2 a, f2 b) {          // operands must not be packed products (see above)
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(d);
}
__device__ __forceinline__ f2 mul2x(f2 a, f2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(d);
}
__device__ __forceinline__ f2 add2s(f2 a, f2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ f2 sub2s(f2 a, f2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }

// div_rn_const / sample_coord of ssm_device.cuh on a pair (same operations, same roundings)
template <int MODE>
__device__ __forceinline__ f2 sample_coord2(f2 pos, f2 flow, float norm, float inv, float m1) {
    const f2 g = add2x(pos, flow);
    const f2 s = mul2x(g, bc2(2.0f));
    f2 n;
    if (MODE == SSM_COORD_DIV) {
        const f2 d = bc2(norm), y = bc2(inv);
        f2 q = mul2x(s, y);
        f2 r = fma2(neg2(q), d, s);
        q = fma2(r, y, q);
        r = fma2(neg2(q), d, s);
        n = fma2(r, y, q);
    } else {
        n = mul2x(s, bc2(inv));
    }
    n = (MODE == SSM_COORD_DIV) ? add2x(n, bc2(-1.0f)) : add2s(n, bc2(-1.0f));
    f2 a = add2x(n, bc2(1.0f));
    a = mul2x(a, bc2(0.5f));
    return mul2x(a, bc2(m1));
}

// F_t0 / F_t1 of flow_interpolation.py:353,356 on a pair (products rounded before the sum, as est_t0 / est_t1)
__device__ __forceinline__ f2 est2_t0(const Coef& c, f2 f01, f2 f10) { return add2s(mul2x(bc2(c.c00), f01), mul2x(bc2(c.c01), f10)); }
__device__ __forceinline__ f2 est2_t1(const Coef& c, f2 f01, f2 f10) { return sub2s(mul2x(bc2(c.c10), f01), mul2x(bc2(c.c11), f10)); }
template <typename T> __device__ __forceinline__ f2 storage_round2(f2 v) { return v; }
template <> __device__ __forceinline__ f2 storage_round2<__nv_bfloat16>(f2 v) {
    return __bfloat1622float2(__floats2bfloat162_rn(v.x, v.y));
}

// the bilinear sample of one frame for the two pixels of a thread
struct QTap2 {
#if SSM_Q8_DIFF_FORM
    f2 wx1, wy1;                // fractional position inside the 2 x 2 cell (east / south weights)
#else
    f2 wnw, wne, wsw, wse;      // corner weights (ATen: area of the opposite sub-rectangle)
#endif
    f2 wsum;                    // sum of the weights of the corners inside the frame
    unsigned idx[2];            // entry index in the launch-wide table, frame base + (y0+1)*(W+1) + (x0+1); only dereferenced when ok
    bool ok[2];                 // the sample has weight inside the frame (then the entry exists)
};

constexpr float Q8_FLOOR_MAGIC = 12582912.0f;       // 1.5 * 2^23: ulp 1 on [2^23, 2^24)
constexpr int Q8_FLOOR_BITS = 0x4B400000;

// xbias = bit pattern of the floor constant - 1 - index of the frame's first entry in the launch-wide table (q8_xbias)
template <int MODE, bool DIRECT = false>
__device__ __forceinline__ QTap2 make_qtap2(f2 posx, f2 posy, f2 u, f2 v, const Geom& g, unsigned xbias) {
    // DIRECT (bf16 storage): the sampling coordinate is x + flow.  The reference's normalise / de-normalise round trip
    // (layers.py:100-112) is the identity up to ~4 roundings (< 3e-4 px at 2048); a flow stored in bf16 is itself only
    // known to 2^-9 relative (0.06-0.25 px at 32-128 px), so reproducing those roundings bit for bit buys nothing there.
    const f2 ix = DIRECT ? add2(posx, u) : sample_coord2<MODE>(posx, u, g.xnorm, g.xinv, g.xm1);
    const f2 iy = DIRECT ? add2(posy, v) : sample_coord2<MODE>(posy, v, g.ynorm, g.yinv, g.ym1);
    // floor without the conversion unit: RD(ix + 1.5 * 2^23) is floor(ix) + 1.5 * 2^23 exactly for |ix| < 2^22, its low
    // mantissa bits are the integer, and subtracting the constant gives floor(ix) as a float.
    const f2 rx = __fadd2_rd(ix, bc2(Q8_FLOOR_MAGIC)), ry = __fadd2_rd(iy, bc2(Q8_FLOOR_MAGIC));
    const f2 fx = add2(rx, bc2(-Q8_FLOOR_MAGIC)), fy = add2(ry, bc2(-Q8_FLOOR_MAGIC));
    // ix - fx is exact; saturation only matters for non-finite coordinates (NaN, inf - inf -> weight 0: such samples
    // read nothing and give 0).  1 - (ix - fx) equals ATen's (fx + 1) - ix except for |ix| < 1, where it may differ
    // by one rounding (6e-8).
    const f2 wx1 = make_float2(__saturatef(ix.x - fx.x), __saturatef(ix.y - fx.y));
    const f2 wy1 = make_float2(__saturatef(iy.x - fy.x), __saturatef(iy.y - fy.y));
    QTap2 t;
#if SSM_Q8_DIFF_FORM
    t.wx1 = wx1; t.wy1 = wy1;
#else
    const f2 wx0 = sub2(bc2(1.0f), wx1), wy0 = sub2(bc2(1.0f), wy1);
    t.wnw = mul2(wx0, wy0); t.wne = mul2(wx1, wy0); t.wsw = mul2(wx0, wy1); t.wse = mul2(wx1, wy1);
#endif
    // total weight of the columns inside the frame: 1 in the interior, ix + 1 for ix in [-1, 0), W - ix for
    // ix in [W-1, W), 0 outside or NaN = min(sat(ix + 1), sat(W - ix)), the same roundings as wx1 / wx0 there
    const float Wf = g.xm1 + 1.0f, Hf = g.ym1 + 1.0f;
    const f2 sx = make_float2(fminf(__saturatef(ix.x + 1.0f), __saturatef(Wf - ix.x)), fminf(__saturatef(ix.y + 1.0f), __saturatef(Wf - ix.y)));
    const f2 sy = make_float2(fminf(__saturatef(iy.x + 1.0f), __saturatef(Hf - iy.x)), fminf(__saturatef(iy.y + 1.0f), __saturatef(Hf - iy.y)));
    t.wsum = mul2(sx, sy);
    // wsum > 0  =>  -1 < ix < W and -1 < iy < H  =>  the entry (x0, y0) in [-1, W-1] x [-1, H-1] exists.  wsum == 0: every
    // corner inside the frame has weight 0 (or the coordinate is not finite): the sample is 0 without reading anything.
    t.ok[0] = t.wsum.x > 0.0f; t.ok[1] = t.wsum.y > 0.0f;
    const unsigned w1 = (unsigned)(g.W + 1);
    t.idx[0] = ((unsigned)__float_as_int(ry.x) - (unsigned)(Q8_FLOOR_BITS - 1)) * w1 + ((unsigned)__float_as_int(rx.x) - xbias);
    t.idx[1] = ((unsigned)__float_as_int(ry.y) - (unsigned)(Q8_FLOOR_BITS - 1)) * w1 + ((unsigned)__float_as_int(rx.y) - xbias);
    return t;
}

// (experiment, see SSM_Q8_BF16_DIRECT_COORD; a shipped version would have to keep the reference's expression for 1-pixel
// axes, which maps every sample of such an axis to coordinate 0)
template <typename T> constexpr bool Q8_DIRECT = SSM_Q8_BF16_DIRECT_COORD && sizeof(T) == 2;

// kept in a register for the whole kernel (the compiler otherwise recomputes the table offset, ~8 instructions, in front
// of every gather)
__device__ __forceinline__ unsigned q8_xbias(unsigned first_entry) {
    unsigned r = (unsigned)(Q8_FLOOR_BITS - 1) - first_entry;
    asm volatile("" : "+r"(r));
    return r;
}

__device__ __forceinline__ uint4 load_entry(const uint4* __restrict__ quads, unsigned idx, bool ok) {
    // unsigned 32-bit entry index: the address is one IMAD.WIDE.U32 (base + idx * 16)
    return ok ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(quads) + (size_t)idx * 16u))
              : make_uint4(0u, 0u, 0u, 0u);
}

// byte k of w as the bit pattern of 2^23 + b (one PRMT); subtracting 2^23 (exact) gives the byte as a float
__device__ __forceinline__ float byte_bits(unsigned w, int k) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 + k)); }

// normalised bilinear samples of the three channels for the two pixels, from their two entries
__device__ __forceinline__ void q8_sample2(const uint4& qa, const uint4& qb, const QTap2& t, const Norm3& nm, f2 (&out)[3]) {
    const unsigned wa[3] = {qa.x, qa.y, qa.z}, wb[3] = {qb.x, qb.y, qb.z};
    const f2 m23 = bc2(-8388608.0f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // bytes 0-2 nw, 3-5 ne, 6-8 sw, 9-11 se
#if SSM_Q8_DIFF_FORM
        // interpolation in difference form: b_nw + wx (b_ne - b_nw) + wy ((b_sw - b_nw) + wx (b_se - b_sw - b_ne + b_nw)).
        // The differences of the 2^23-biased bit patterns are exact and drop the bias; the four corner weights and their
        // six packed operations per frame are never formed.
        const f2 Bnw = make_float2(byte_bits(wa[c >> 2], c & 3), byte_bits(wb[c >> 2], c & 3));
        const f2 Bne = make_float2(byte_bits(wa[(3 + c) >> 2], (3 + c) & 3), byte_bits(wb[(3 + c) >> 2], (3 + c) & 3));
        const f2 Bsw = make_float2(byte_bits(wa[(6 + c) >> 2], (6 + c) & 3), byte_bits(wb[(6 + c) >> 2], (6 + c) & 3));
        const f2 Bse = make_float2(byte_bits(wa[(9 + c) >> 2], (9 + c) & 3), byte_bits(wb[(9 + c) >> 2], (9 + c) & 3));
        const f2 dx = sub2(Bne, Bnw), dy = sub2(Bsw, Bnw), dxy = sub2(sub2(Bse, Bsw), dx);
        f2 s = fma2(t.wx1, dx, add2(Bnw, m23));
        s = fma2(t.wy1, fma2(t.wx1, dxy, dy), s);
#else
        const f2 bnw = add2(make_float2(byte_bits(wa[c >> 2], c & 3), byte_bits(wb[c >> 2], c & 3)), m23);
        const f2 bne = add2(make_float2(byte_bits(wa[(3 + c) >> 2], (3 + c) & 3), byte_bits(wb[(3 + c) >> 2], (3 + c) & 3)), m23);
        const f2 bsw = add2(make_float2(byte_bits(wa[(6 + c) >> 2], (6 + c) & 3), byte_bits(wb[(6 + c) >> 2], (6 + c) & 3)), m23);
        const f2 bse = add2(make_float2(byte_bits(wa[(9 + c) >> 2], (9 + c) & 3), byte_bits(wb[(9 + c) >> 2], (9 + c) & 3)), m23);
        f2 s = mul2(bnw, t.wnw);
        s = fma2(bne, t.wne, s);
        s = fma2(bsw, t.wsw, s);
        s = fma2(bse, t.wse, s);
#endif
        out[c] = fma2(bc2(nm.a[c]), s, mul2(bc2(nm.c[c]), t.wsum));
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
__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 ldg2(const __nv_bfloat16* p) {
    const unsigned w = __ldg(reinterpret_cast<const unsigned*>(p));
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
// 0.35 ms per 32 frames of 1088x1920 (1.2 GB moved).  Tried and measured slower or equal (profiles/r02s-r02v): the rows
// read as aligned 32-bit words cut up with funnel shifts (10 loads per thread instead of 30: 0.41 ms), one entry per
// thread with fully coalesced 16-byte stores (0.43 ms; 0.41 with a 2-D grid and no 64-bit divisions), streaming stores
// (no change), both texel rows staged in shared memory with coalesced word loads and one entry per lane (0.43 ms with one
// tile per CTA, 0.53 ms with persistent CTAs; profiles/r04b, r04c); 32-bit instead of 64-bit index divisions (no change, r04w).
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
// CTAs per SM, measured with the packed-arithmetic kernels (profiles/r02h_q8_timing_*.json; 3 CTAs = 80 registers,
// 4 CTAs = 64 registers with a few spilled bytes): compute_inputs 2.86 ms at 3, 2.94 at 4; compute_output_image with an
// fp32 U-Net output 1.86 ms at 3, 2.03 at 4; with a bf16 U-Net output 2.25 at 3, 2.12 at 4; with the fused uint8
// output 2.13 at 3, 2.36 at 4.
#ifndef SSM_Q8_PACK_MIN_BLOCKS
#define SSM_Q8_PACK_MIN_BLOCKS 3
#endif
#ifndef SSM_Q8_FUSE_MIN_BLOCKS
#define SSM_Q8_FUSE_MIN_BLOCKS 3
#endif
#ifndef SSM_Q8_FUSE_BF16_MIN_BLOCKS
#define SSM_Q8_FUSE_BF16_MIN_BLOCKS 4
#endif
// channels-last output only: the six pass-through channels are re-read every timestep (L1 / L2 hits) instead of living
// in 12 registers, which lets 4 CTAs fit per SM: 2.05 -> 1.98 ms (rough flow) / 1.98 -> 1.84 ms (smooth) for the bf16
// channels-last tensor; the planar fp32 kernel is DRAM-bound and gets slower that way (2.90 -> 2.99 ms)
// (profiles/r03h_q8_timing_pack_*.json)
#ifndef SSM_Q8_RELOAD_BF16
#define SSM_Q8_RELOAD_BF16 1       // the same for bf16 storage (planar): 1.92 -> 1.83 ms (profiles/r03n_bf16_*.json)
#endif
#ifndef SSM_Q8_PACK_NHWC_MIN_BLOCKS
#define SSM_Q8_PACK_NHWC_MIN_BLOCKS 4
#endif
#ifndef SSM_Q8_FUSE_PREFETCH
#define SSM_Q8_FUSE_PREFETCH 1     // U-Net output of timestep n+1 loaded during the gathers of timestep n: 1.85 ms; without it
                                   // 2.10 ms at 3 CTAs/SM and 1.99 ms at 4 (profiles/r03c_q8_timing_fuse_prefetch_*.json)
#endif
// Experiment (off): the gathers of timestep n+1 are issued before the interpolation of timestep n (a whole loop trip in
// flight); only the fractional weights, the in-frame weight and the logit travel with the entries.  102-127 registers
// (2 CTAs/SM) or 80 with spills (3 CTAs/SM).  Measured SLOWER: 1.91 ms shipped -> 2.53 ms (2 CTAs/SM) / 2.27 ms (3, spilling);
// the shipped loop at 2 CTAs/SM takes 2.63 ms -- the kernel needs its 24 warps per SM more than a longer load-to-use
// distance (profiles/r04e_q8_timing_fuse_*.json, tools/gpu_exp15.sh).
#ifndef SSM_Q8_FUSE_PIPE
#define SSM_Q8_FUSE_PIPE 0
#endif
#ifndef SSM_Q8_FUSE_U8_MIN_BLOCKS
#define SSM_Q8_FUSE_U8_MIN_BLOCKS 3
#endif
// T: storage type of img6 / flow4 (and of out16 in the planar layout): fp32, or bf16 with fp32 arithmetic
template <typename T, int MODE, typename TO, bool NHWC>
__global__ void __launch_bounds__(Q8_THREADS, (NHWC || (SSM_Q8_RELOAD_BF16 && sizeof(T) == 2)) ? SSM_Q8_PACK_NHWC_MIN_BLOCKS : SSM_Q8_PACK_MIN_BLOCKS)
flow_pack_fwd_q8_kernel(View<const T> img6, const uint4* __restrict__ quads, View<const T> flow4,
                        const float* __restrict__ tv, View<TO> out16, int N, Geom g, Norm3 nm,
                        const float* __restrict__ lut) {
    const Q8Idx ti = q8_index(g.H, g.W);
    if (!ti.valid) return;
    const int p = ti.y * g.W + ti.x;
    // entry index of the launch-wide table as an unsigned 32-bit number (B * 2 * entries per frame < 2^32, checked on
    // the host): the address of a gather is one IMAD.WIDE.U32 on the kernel parameter, no pointer kept in registers
    const unsigned epf = (unsigned)(g.H + 1) * (unsigned)(g.W + 1);
    const unsigned xb0 = q8_xbias((unsigned)ti.b * 2u * epf), xb1 = q8_xbias((unsigned)ti.b * 2u * epf + epf);
    const T* F = flow4.p + ti.b * flow4.sb + p;
    const int fsc = (int)flow4.sc, isc = (int)img6.sc;
    const float2 f01x = lds2(F), f01y = lds2(F + fsc), f10x = lds2(F + 2 * fsc), f10y = lds2(F + 3 * fsc);
    const T* I = img6.p + ti.b * img6.sb + p;
    constexpr bool RELOAD = NHWC || (SSM_Q8_RELOAD_BF16 && sizeof(T) == 2);
    float2 c0[3], c1[3];
    // lut != NULL (3 x 256 normalised values, the table ssm_frames_from_u8 applies): the pass-through channels I0, I1 at the
    // thread's two pixels come from the entry tables the kernel gathers from anyway -- entry (x, y) holds RGB(x, y) in
    // bytes 0-2 and RGB(x+1, y) in bytes 3-5 -- and the planar frames are not read at all: 0.8 GB of 17.3 GB less DRAM
    // traffic at 16 x 1088 x 1920, bit-identical values (the same table entry either way).
    const unsigned own = (unsigned)ti.b * 2u * epf + (unsigned)(ti.y + 1) * (unsigned)(g.W + 1) + (unsigned)(ti.x + 1);
    auto from_tables = [&]() {
        const uint4 e0 = load_entry(quads, own, true), e1 = load_entry(quads, own + epf, true);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* l = lut + 256 * c;
            c0[c] = make_float2(__ldg(l + ((e0.x >> (8 * c)) & 0xffu)), __ldg(l + (c == 0 ? (e0.x >> 24) : ((e0.y >> (8 * (c - 1))) & 0xffu))));
            c1[c] = make_float2(__ldg(l + ((e1.x >> (8 * c)) & 0xffu)), __ldg(l + (c == 0 ? (e1.x >> 24) : ((e1.y >> (8 * (c - 1))) & 0xffu))));
        }
    };
    if constexpr (!RELOAD) {
        if (lut) from_tables();
        else {
#pragma unroll
            for (int c = 0; c < 3; ++c) { c0[c] = lds2(I + c * isc); c1[c] = lds2(I + (3 + c) * isc); }
        }
    }
    const float* tp = tv + ti.b * N;
    TO* __restrict__ O = out16.p + ti.b * out16.sb + (NHWC ? (long long)p * 16 : (long long)p);
    const int osc = (int)out16.sc;
    const f2 posx = make_float2((float)ti.x, (float)(ti.x + 1)), posy = bc2((float)ti.y);
    for (int n = 0; n < N; ++n, O += out16.sn) {
        const Coef k = make_coef(__ldg(tp + n));
        if constexpr (RELOAD) {
            if (lut) from_tables();
            else {
#pragma unroll
                for (int c = 0; c < 3; ++c) { c0[c] = ldg2(I + c * isc); c1[c] = ldg2(I + (3 + c) * isc); }
            }
        }
        const f2 e0x = storage_round2<T>(est2_t0(k, f01x, f10x)), e0y = storage_round2<T>(est2_t0(k, f01y, f10y));   // F_t0  :353
        const f2 e1x = storage_round2<T>(est2_t1(k, f01x, f10x)), e1y = storage_round2<T>(est2_t1(k, f01y, f10y));   // F_t1  :356
        const QTap2 t1 = make_qtap2<MODE, Q8_DIRECT<T>>(posx, posy, e1x, e1y, g, xb1);                             // warp(img_1, F_t1) :361
        const uint4 q1a = load_entry(quads, t1.idx[0], t1.ok[0]), q1b = load_entry(quads, t1.idx[1], t1.ok[1]);
        const QTap2 t0 = make_qtap2<MODE, Q8_DIRECT<T>>(posx, posy, e0x, e0y, g, xb0);                             // warp(img_0, F_t0) :362
        const uint4 q0a = load_entry(quads, t0.idx[0], t0.ok[0]), q0b = load_entry(quads, t0.idx[1], t0.ok[1]);
        f2 w1[3], w0[3];
        q8_sample2(q1a, q1b, t1, nm, w1);
        q8_sample2(q0a, q0b, t0, nm, w0);
        if constexpr (NHWC) {                                                                 // :364-367, channels-last
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float o[16];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    o[c] = i ? c1[c].y : c1[c].x; o[3 + c] = i ? w1[c].y : w1[c].x;
                    o[10 + c] = i ? w0[c].y : w0[c].x; o[13 + c] = i ? c0[c].y : c0[c].x;
                }
                o[6] = i ? e1x.y : e1x.x; o[7] = i ? e1y.y : e1y.x; o[8] = i ? e0x.y : e0x.x; o[9] = i ? e0y.y : e0y.x;
                store16_nhwc<TO>(O + 16 * i, o);
            }
        } else {
            TO* Of = O;
#pragma unroll
            for (int c = 0; c < 3; ++c) {                                                     // :364-367
                sts2(Of + (0 + c) * osc, c1[c].x, c1[c].y);
                sts2(Of + (3 + c) * osc, w1[c].x, w1[c].y);
                sts2(Of + (10 + c) * osc, w0[c].x, w0[c].y);
                sts2(Of + (13 + c) * osc, c0[c].x, c0[c].y);
            }
            sts2(Of + 6 * osc, e1x.x, e1x.y); sts2(Of + 7 * osc, e1y.x, e1y.y);
            sts2(Of + 8 * osc, e0x.x, e0x.y); sts2(Of + 9 * osc, e0y.x, e0y.y);
        }
    }
}

// =============================================================================================
// a3 + a4: extract_outputs + compute_output_image from entry tables   flow_interpolation.py:374-429
//     estimated flows recomputed from the stage-1 flows (as fuse_fwd_kernel<RECOMP>); TY = storage type of the
//     U-Net output.  OUT_U8: the fused frame is de-normalised and written as uint8 H_out x W_out x 3 images
//     (crop at (top, left)) with the arithmetic of frames_to_u8_kernel, instead of as fp32 planes.
// =============================================================================================
#if SSM_Q8_FUSE_PIPE
struct QFrac2 { f2 wx1, wy1, wsum; unsigned idx[2]; bool ok[2]; };
template <int MODE>
__device__ __forceinline__ QFrac2 make_qfrac2(f2 posx, f2 posy, f2 u, f2 v, const Geom& g, unsigned xbias) {   // make_qtap2 without the corner products
    const f2 ix = sample_coord2<MODE>(posx, u, g.xnorm, g.xinv, g.xm1);
    const f2 iy = sample_coord2<MODE>(posy, v, g.ynorm, g.yinv, g.ym1);
    const f2 rx = __fadd2_rd(ix, bc2(Q8_FLOOR_MAGIC)), ry = __fadd2_rd(iy, bc2(Q8_FLOOR_MAGIC));
    const f2 fx = add2(rx, bc2(-Q8_FLOOR_MAGIC)), fy = add2(ry, bc2(-Q8_FLOOR_MAGIC));
    QFrac2 t;
    t.wx1 = make_float2(__saturatef(ix.x - fx.x), __saturatef(ix.y - fx.y));
    t.wy1 = make_float2(__saturatef(iy.x - fy.x), __saturatef(iy.y - fy.y));
    const float Wf = g.xm1 + 1.0f, Hf = g.ym1 + 1.0f;
    const f2 sx = make_float2(fminf(__saturatef(ix.x + 1.0f), __saturatef(Wf - ix.x)), fminf(__saturatef(ix.y + 1.0f), __saturatef(Wf - ix.y)));
    const f2 sy = make_float2(fminf(__saturatef(iy.x + 1.0f), __saturatef(Hf - iy.x)), fminf(__saturatef(iy.y + 1.0f), __saturatef(Hf - iy.y)));
    t.wsum = mul2(sx, sy);
    t.ok[0] = t.wsum.x > 0.0f; t.ok[1] = t.wsum.y > 0.0f;
    const unsigned w1 = (unsigned)(g.W + 1);
    t.idx[0] = ((unsigned)__float_as_int(ry.x) - (unsigned)(Q8_FLOOR_BITS - 1)) * w1 + ((unsigned)__float_as_int(rx.x) - xbias);
    t.idx[1] = ((unsigned)__float_as_int(ry.y) - (unsigned)(Q8_FLOOR_BITS - 1)) * w1 + ((unsigned)__float_as_int(rx.y) - xbias);
    return t;
}
__device__ __forceinline__ QTap2 corner_weights(f2 wx1, f2 wy1, f2 wsum) {
    QTap2 t;
    const f2 wx0 = sub2(bc2(1.0f), wx1), wy0 = sub2(bc2(1.0f), wy1);
    t.wnw = mul2(wx0, wy0); t.wne = mul2(wx1, wy0); t.wsw = mul2(wx0, wy1); t.wse = mul2(wx1, wy1);
    t.wsum = wsum;
    return t;
}
// a timestep whose gathers are in flight
struct FuseStage { uint4 q0a, q0b, q1a, q1b; f2 wx0, wy0, ws0, wx1, wy1, ws1, y0; float tt; };
#endif

struct U8Out {
    unsigned char* dst; long long frame_stride; int row_stride;    // frame index = b * N + n
    int top, left, H_out, W_out, bgr, saturate;
    float mean[3], std[3], scale;
};

template <typename T, int MODE, typename TY, bool OUT_U8>
__global__ void __launch_bounds__(Q8_THREADS, OUT_U8 ? SSM_Q8_FUSE_U8_MIN_BLOCKS : (sizeof(TY) == 2 ? SSM_Q8_FUSE_BF16_MIN_BLOCKS : SSM_Q8_FUSE_MIN_BLOCKS))
fuse_fwd_q8_kernel(const uint4* __restrict__ quads, View<const T> flow4, View<const TY> out5,
                   const float* __restrict__ tv, View<T> out3, U8Out u8, int N, Geom g, Norm3 nm) {
    const Q8Idx ti = q8_index(g.H, g.W);
    if (!ti.valid) return;
    const int p = ti.y * g.W + ti.x;
    const unsigned epf = (unsigned)(g.H + 1) * (unsigned)(g.W + 1);
    const unsigned xb0 = q8_xbias((unsigned)ti.b * 2u * epf), xb1 = q8_xbias((unsigned)ti.b * 2u * epf + epf);
    const float* tp = tv + ti.b * N;
    const T* F = flow4.p + ti.b * flow4.sb + p;
    const int fsc = (int)flow4.sc, ysc = (int)out5.sc, osc = (int)out3.sc;
    const float2 f01x = lds2(F), f01y = lds2(F + fsc), f10x = lds2(F + 2 * fsc), f10y = lds2(F + 3 * fsc);
    const TY* __restrict__ Y = out5.p + ti.b * out5.sb + p;
    T* __restrict__ O = OUT_U8 ? nullptr : out3.p + ti.b * out3.sb + p;
    const f2 posx = make_float2((float)ti.x, (float)(ti.x + 1)), posy = bc2((float)ti.y);
#if SSM_Q8_FUSE_PIPE
    f2 ys[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) ys[c] = lds2(Y + c * ysc);
    // gather half of timestep m: coordinates and the four entry loads; starts the streaming loads of timestep m + 1
    auto gather = [&](int m) {
        FuseStage st;
        st.tt = __ldg(tp + m);
        const Coef k = make_coef(st.tt);
        const f2 y1 = ys[1], y2 = ys[2], y3 = ys[3], y4 = ys[4];
        st.y0 = ys[0];
        if (m + 1 < N) {
            Y += out5.sn;
#pragma unroll
            for (int c = 0; c < 5; ++c) ys[c] = lds2(Y + c * ysc);
        }
        const f2 f1x = add2x(storage_round2<T>(est2_t1(k, f01x, f10x)), y1);                  // :412
        const f2 f1y = add2x(storage_round2<T>(est2_t1(k, f01y, f10y)), y2);
        const f2 f0x = add2x(storage_round2<T>(est2_t0(k, f01x, f10x)), y3);                  // :413
        const f2 f0y = add2x(storage_round2<T>(est2_t0(k, f01y, f10y)), y4);
        const QFrac2 t0 = make_qfrac2<MODE>(posx, posy, f0x, f0y, g, xb0);                          // :416
        st.q0a = load_entry(quads, t0.idx[0], t0.ok[0]); st.q0b = load_entry(quads, t0.idx[1], t0.ok[1]);
        const QFrac2 t1 = make_qfrac2<MODE>(posx, posy, f1x, f1y, g, xb1);                          // :418
        st.q1a = load_entry(quads, t1.idx[0], t1.ok[0]); st.q1b = load_entry(quads, t1.idx[1], t1.ok[1]);
        st.wx0 = t0.wx1; st.wy0 = t0.wy1; st.ws0 = t0.wsum; st.wx1 = t1.wx1; st.wy1 = t1.wy1; st.ws1 = t1.wsum;
        return st;
    };
    FuseStage cur = gather(0);
    for (int n = 0; n < N; ++n) {
        FuseStage nxt = cur;
        if (n + 1 < N) nxt = gather(n + 1);
        const float tt = cur.tt;
        const float omt_ = __fsub_rn(1.0f, tt);
        struct { float omt; } k = {omt_};
        const f2 y0 = cur.y0;
        const QTap2 t0 = corner_weights(cur.wx0, cur.wy0, cur.ws0), t1 = corner_weights(cur.wx1, cur.wy1, cur.ws1);
        const uint4 q0a = cur.q0a, q0b = cur.q0b, q1a = cur.q1a, q1b = cur.q1b;
        cur = nxt;
        const f2 e = mul2(y0, bc2(-1.4426950408889634f));
        const f2 d = add2(make_float2(ex2_approx(e.x), ex2_approx(e.y)), bc2(1.0f));
        const f2 v1 = make_float2(rcp_approx(d.x), rcp_approx(d.y));
        const f2 v0 = sub2(bc2(1.0f), v1);
        const f2 a0 = mul2(bc2(k.omt), v0), a1 = mul2(bc2(tt), v1);
        const f2 z = add2(a0, a1);
        const f2 rz = make_float2(rcp_approx(z.x), rcp_approx(z.y));                           // 1/Z  :425
        f2 s0[3], s1[3], res[3];
        q8_sample2(q0a, q0b, t0, nm, s0);
        q8_sample2(q1a, q1b, t1, nm, s1);
#else
    f2 ys[5];
#if SSM_Q8_FUSE_PREFETCH
#pragma unroll
    for (int c = 0; c < 5; ++c) ys[c] = lds2(Y + c * ysc);
#endif
    for (int n = 0; n < N; ++n) {
        const float tt = __ldg(tp + n);
        const Coef k = make_coef(tt);
#if SSM_Q8_FUSE_PREFETCH
        const f2 y0 = ys[0], y1 = ys[1], y2 = ys[2], y3 = ys[3], y4 = ys[4];
        if (n + 1 < N) {             // streaming loads of the next timestep, in flight during the gathers
            Y += out5.sn;
#pragma unroll
            for (int c = 0; c < 5; ++c) ys[c] = lds2(Y + c * ysc);
        }
#else
#pragma unroll
        for (int c = 0; c < 5; ++c) ys[c] = lds2(Y + c * ysc);
        Y += out5.sn;
        const f2 y0 = ys[0], y1 = ys[1], y2 = ys[2], y3 = ys[3], y4 = ys[4];
#endif
        const f2 f1x = add2x(storage_round2<T>(est2_t1(k, f01x, f10x)), y1);                  // :412
        const f2 f1y = add2x(storage_round2<T>(est2_t1(k, f01y, f10y)), y2);
        const f2 f0x = add2x(storage_round2<T>(est2_t0(k, f01x, f10x)), y3);                  // :413
        const f2 f0y = add2x(storage_round2<T>(est2_t0(k, f01y, f10y)), y4);
        const QTap2 t0 = make_qtap2<MODE, Q8_DIRECT<T>>(posx, posy, f0x, f0y, g, xb0);                            // :416
        const uint4 q0a = load_entry(quads, t0.idx[0], t0.ok[0]), q0b = load_entry(quads, t0.idx[1], t0.ok[1]);
        const QTap2 t1 = make_qtap2<MODE, Q8_DIRECT<T>>(posx, posy, f1x, f1y, g, xb1);                            // :418
        const uint4 q1a = load_entry(quads, t1.idx[0], t1.ok[0]), q1b = load_entry(quads, t1.idx[1], t1.ok[1]);
        // V_t<-1 = sigmoid(out[:, 0]) (:386-388) as 1 / (1 + 2^(-x log2 e)); V_t<-0 = 1 - V_t<-1 (:390)
        const f2 e = mul2(y0, bc2(-1.4426950408889634f));
        const f2 d = add2(make_float2(ex2_approx(e.x), ex2_approx(e.y)), bc2(1.0f));
        const f2 v1 = make_float2(rcp_approx(d.x), rcp_approx(d.y));
        const f2 v0 = sub2(bc2(1.0f), v1);
        const f2 a0 = mul2(bc2(k.omt), v0), a1 = mul2(bc2(tt), v1);
        const f2 z = add2(a0, a1);
        const f2 rz = make_float2(rcp_approx(z.x), rcp_approx(z.y));                           // 1/Z  :425
        f2 s0[3], s1[3], res[3];
        q8_sample2(q0a, q0b, t0, nm, s0);
        q8_sample2(q1a, q1b, t1, nm, s1);
#endif
#pragma unroll
        for (int c = 0; c < 3; ++c) res[c] = mul2(fma2(a1, s1[c], mul2(a0, s0[c])), rz);       // :420-427
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
                            o[u8.bgr ? 2 - c : c] = to_u8(i ? res[c].y : res[c].x, u8.std[c], u8.mean[c], u8.scale, u8.saturate);
                    }
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) sts2(O + c * osc, res[c].x, res[c].y);
            O += out3.sn;
        }
    }
}

}  // namespace ssm
