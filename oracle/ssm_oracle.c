/*
 * ssm_oracle.c -- CPU ORACLE for the Super SloMo intermediate-frame synthesis path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
 * __graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py may build,
 * load or call it.  The product path (libssm_b200.so, CUDA only) never links or calls it.
 *
 * What it restates (plain C, scalar fp32, every operation individually rounded -- build with
 * -ffp-contract=off, see oracle/Makefile):
 *
 *   warp                  reference scripts/models/layers.py:73-120
 *                         (mesh grid :92-96, vgrid = grid + flo :100, normalisation
 *                          2.0*u/max(W-1,1)-1.0 :112-113, grid_sample(align_corners=True) :119)
 *   bilinear sampler      third-party: PyTorch ATen grid_sampler_2d, mode=bilinear,
 *                         padding_mode=zeros, align_corners=True.  Not vendored in
 *                         /root/reference; the reference pins pytorch=1.6.0
 *                         (configs/conda_env.yml:137), this image runs torch 2.11.0.
 *                         Published algorithm: un-normalise ((c+1)/2)*(size-1); corners
 *                         nw=(floor ix, floor iy), ne, sw, se; weight of a corner = area of the
 *                         opposite sub-rectangle; taps outside the image contribute zero.
 *   flow approximation    scripts/models/flow_interpolation.py:353 (F_t0) and :356 (F_t1)
 *   16-channel packing    scripts/models/flow_interpolation.py:364-367
 *   visibility + fusion   scripts/models/flow_interpolation.py:382-392, 402-427
 *   backward              the reference has no hand-written backward; this restates what
 *                         autograd derives for the functions above (SURVEY.md section 8 note).
 *
 * Parity pin: the reference publishes no golden vectors or tests.  This oracle is pinned against
 * outputs of the reference itself, imported from /root/reference/scripts and run on CPU in the
 * build container; the inputs/outputs are committed under tests/golden/ together with the script
 * that generated them (tests/golden/make_golden.py); tests/test_oracle_golden.py checks them.
 *
 * coord_mode selects how the division by max(W-1,1) in layers.py:112-113 is rounded:
 *   0 (SSM_COORD_DIV)  true IEEE division        -- what torch's CPU kernel does
 *   1 (SSM_COORD_RCP)  multiply by fp32 (1/(W-1)) -- what torch's CUDA kernel does for a
 *                      Python-scalar divisor
 * Sampling coordinates must be bit-identical to the reference's or floor() flips at cell borders
 * change the flow gradient by O(1); everything downstream of the coordinate only needs to agree
 * to rounding error.
 *
 * All tensors are dense NCHW fp32.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* coordinate arithmetic                                                                       */

/* layers.py:100,112-113 followed by ATen's unnormalise for align_corners=True. */
static inline float sample_coord(float pos, float flow, int size, int coord_mode)
{
    const float norm = (float)(size - 1 > 1 ? size - 1 : 1);  /* max(size-1, 1), layers.py:112 */
    float g = pos + flow;                                      /* grid + flo, layers.py:100 */
    float s = 2.0f * g;
    float n;
    if (coord_mode == 0) {
        n = s / norm;
    } else {
        float inv = 1.0f / norm;
        n = s * inv;
    }
    n = n - 1.0f;
    /* grid_sampler_unnormalize, align_corners: ((coord + 1) / 2) * (size - 1) */
    float a = n + 1.0f;
    a = a / 2.0f;
    return a * (float)(size - 1);
}

typedef struct {
    float ix, iy;          /* un-normalised sampling position */
    int x0, y0;            /* north-west corner */
    float wnw, wne, wsw, wse;
    int in_nw, in_ne, in_sw, in_se;
} tap_t;

static inline void make_taps(tap_t* t, int x, int y, float u, float v, int H, int W, int coord_mode)
{
    float ix = sample_coord((float)x, u, W, coord_mode);
    float iy = sample_coord((float)y, v, H, coord_mode);
    /* far outside: every tap is out of bounds; keep the int conversion defined */
    if (!(ix > -2.0f)) ix = -2.0f;
    if (!(ix < (float)W + 1.0f)) ix = (float)W + 1.0f;
    if (!(iy > -2.0f)) iy = -2.0f;
    if (!(iy < (float)H + 1.0f)) iy = (float)H + 1.0f;
    float fx = floorf(ix), fy = floorf(iy);
    t->ix = ix; t->iy = iy;
    t->x0 = (int)fx; t->y0 = (int)fy;
    float xe = fx + 1.0f, ys = fy + 1.0f;   /* ix_se, iy_se */
    t->wnw = (xe - ix) * (ys - iy);
    t->wne = (ix - fx) * (ys - iy);
    t->wsw = (xe - ix) * (iy - fy);
    t->wse = (ix - fx) * (iy - fy);
    int xin0 = t->x0 >= 0 && t->x0 < W, xin1 = t->x0 + 1 >= 0 && t->x0 + 1 < W;
    int yin0 = t->y0 >= 0 && t->y0 < H, yin1 = t->y0 + 1 >= 0 && t->y0 + 1 < H;
    t->in_nw = xin0 && yin0; t->in_ne = xin1 && yin0;
    t->in_sw = xin0 && yin1; t->in_se = xin1 && yin1;
}

static inline float sample_plane(const float* p, const tap_t* t, int W)
{
    float acc = 0.0f;
    const float* q = p + (ptrdiff_t)t->y0 * W + t->x0;
    if (t->in_nw) acc += q[0] * t->wnw;
    if (t->in_ne) acc += q[1] * t->wne;
    if (t->in_sw) acc += q[W] * t->wsw;
    if (t->in_se) acc += q[W + 1] * t->wse;
    return acc;
}

/* d(sample)/d(ix), d(sample)/d(iy) contributions of one plane with upstream gradient g. */
static inline void sample_plane_grad(const float* p, const tap_t* t, int W, float g, float* gix, float* giy)
{
    const float* q = p + (ptrdiff_t)t->y0 * W + t->x0;
    float fx = (float)t->x0, fy = (float)t->y0;
    float xe = fx + 1.0f, ys = fy + 1.0f;
    float ix = t->ix, iy = t->iy;
    if (t->in_nw) { float v = q[0];     *gix -= v * (ys - iy) * g; *giy -= v * (xe - ix) * g; }
    if (t->in_ne) { float v = q[1];     *gix += v * (ys - iy) * g; *giy -= v * (ix - fx) * g; }
    if (t->in_sw) { float v = q[W];     *gix -= v * (iy - fy) * g; *giy += v * (xe - ix) * g; }
    if (t->in_se) { float v = q[W + 1]; *gix += v * (iy - fy) * g; *giy += v * (ix - fx) * g; }
}

/* Chain rule from d/d(ix) back to d/d(flow): ATen multiplies by (size-1)/2, autograd of
 * layers.py:112 then divides by max(size-1,1) (same rounding mode as the forward) and
 * multiplies by 2.0. */
static inline float coord_grad_to_flow(float gi, int size, int coord_mode)
{
    const float norm = (float)(size - 1 > 1 ? size - 1 : 1);
    float g = gi * ((float)(size - 1) / 2.0f);
    if (coord_mode == 0) g = g / norm; else g = g * (1.0f / norm);
    return g * 2.0f;
}

static inline void scatter_plane(float* p, const tap_t* t, int W, float g)
{
    float* q = p + (ptrdiff_t)t->y0 * W + t->x0;
    if (t->in_nw) q[0]     += t->wnw * g;
    if (t->in_ne) q[1]     += t->wne * g;
    if (t->in_sw) q[W]     += t->wsw * g;
    if (t->in_se) q[W + 1] += t->wse * g;
}

/* ------------------------------------------------------------------------------------------ */
/* a1: warp  (layers.py:73-120)                                                                */

void ssm_oracle_warp_fwd(const float* img, const float* flo, float* out,
                         int B, int C, int H, int W, int coord_mode)
{
    const size_t npx = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                size_t p = (size_t)y * W + x;
                tap_t t;
                make_taps(&t, x, y, flo[((size_t)b * 2 + 0) * npx + p], flo[((size_t)b * 2 + 1) * npx + p], H, W, coord_mode);
                for (int c = 0; c < C; ++c)
                    out[((size_t)b * C + c) * npx + p] = sample_plane(img + ((size_t)b * C + c) * npx, &t, W);
            }
    }
}

/* gimg / gflo may be NULL.  gimg is overwritten (zero-filled first). */
void ssm_oracle_warp_bwd(const float* gout, const float* img, const float* flo,
                         float* gimg, float* gflo, int B, int C, int H, int W, int coord_mode)
{
    const size_t npx = (size_t)H * W;
    if (gimg) memset(gimg, 0, sizeof(float) * B * C * npx);
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                size_t p = (size_t)y * W + x;
                tap_t t;
                make_taps(&t, x, y, flo[((size_t)b * 2 + 0) * npx + p], flo[((size_t)b * 2 + 1) * npx + p], H, W, coord_mode);
                float gix = 0.0f, giy = 0.0f;
                for (int c = 0; c < C; ++c) {
                    float g = gout[((size_t)b * C + c) * npx + p];
                    if (gflo) sample_plane_grad(img + ((size_t)b * C + c) * npx, &t, W, g, &gix, &giy);
                    if (gimg) scatter_plane(gimg + ((size_t)b * C + c) * npx, &t, W, g);
                }
                if (gflo) {
                    gflo[((size_t)b * 2 + 0) * npx + p] = coord_grad_to_flow(gix, W, coord_mode);
                    gflo[((size_t)b * 2 + 1) * npx + p] = coord_grad_to_flow(giy, H, coord_mode);
                }
            }
}

/* ------------------------------------------------------------------------------------------ */
/* a2: compute_inputs  (flow_interpolation.py:338-372).  t has one value per sample.           */

typedef struct { float c00, c01, c10, c11, omt, t; } coef_t;

static inline coef_t make_coef(float t)
{
    coef_t c;
    float omt = 1.0f - t;          /* (1 - t) */
    c.c00 = (-omt) * t;            /* -(1 - t) * t      :353 */
    c.c01 = t * t;                 /* t ** 2            :353 */
    c.c10 = omt * omt;             /* (1 - t) ** 2      :356 */
    c.c11 = t * omt;               /* t * (1 - t)       :356 */
    c.omt = omt; c.t = t;
    return c;
}

void ssm_oracle_flow_pack_fwd(const float* img6, const float* flow4, const float* t,
                              float* out16, int B, int H, int W, int coord_mode)
{
    const size_t npx = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
        const coef_t k = make_coef(t[b]);
        const float* I0 = img6 + (size_t)b * 6 * npx;
        const float* I1 = I0 + 3 * npx;
        const float* F = flow4 + (size_t)b * 4 * npx;
        float* O = out16 + (size_t)b * 16 * npx;
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                size_t p = (size_t)y * W + x;
                float f01x = F[p], f01y = F[npx + p], f10x = F[2 * npx + p], f10y = F[3 * npx + p];
                float a, bb;
                a = k.c00 * f01x; bb = k.c01 * f10x; float e0x = a + bb;   /* F_t0 :353 */
                a = k.c00 * f01y; bb = k.c01 * f10y; float e0y = a + bb;
                a = k.c10 * f01x; bb = k.c11 * f10x; float e1x = a - bb;   /* F_t1 :356 */
                a = k.c10 * f01y; bb = k.c11 * f10y; float e1y = a - bb;
                tap_t t0, t1;
                make_taps(&t1, x, y, e1x, e1y, H, W, coord_mode);           /* warp(img_1, F_t1) :361 */
                make_taps(&t0, x, y, e0x, e0y, H, W, coord_mode);           /* warp(img_0, F_t0) :362 */
                for (int c = 0; c < 3; ++c) {
                    O[(0 + c) * npx + p] = I1[c * npx + p];                  /* :364-367 */
                    O[(3 + c) * npx + p] = sample_plane(I1 + c * npx, &t1, W);
                    O[(10 + c) * npx + p] = sample_plane(I0 + c * npx, &t0, W);
                    O[(13 + c) * npx + p] = I0[c * npx + p];
                }
                O[6 * npx + p] = e1x; O[7 * npx + p] = e1y;
                O[8 * npx + p] = e0x; O[9 * npx + p] = e0y;
            }
    }
}

/* Backward of compute_inputs.  g16 is dL/d(out16).  gflow4 and gimg6 may be NULL. */
void ssm_oracle_flow_pack_bwd(const float* g16, const float* img6, const float* flow4, const float* t,
                              float* gflow4, float* gimg6, int B, int H, int W, int coord_mode)
{
    const size_t npx = (size_t)H * W;
    if (gimg6) memset(gimg6, 0, sizeof(float) * B * 6 * npx);
    for (int b = 0; b < B; ++b) {
        const coef_t k = make_coef(t[b]);
        const float* I0 = img6 + (size_t)b * 6 * npx;
        const float* I1 = I0 + 3 * npx;
        const float* F = flow4 + (size_t)b * 4 * npx;
        const float* G = g16 + (size_t)b * 16 * npx;
        float* gI0 = gimg6 ? gimg6 + (size_t)b * 6 * npx : NULL;
        float* gI1 = gimg6 ? gI0 + 3 * npx : NULL;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                size_t p = (size_t)y * W + x;
                float f01x = F[p], f01y = F[npx + p], f10x = F[2 * npx + p], f10y = F[3 * npx + p];
                float a, bb;
                a = k.c00 * f01x; bb = k.c01 * f10x; float e0x = a + bb;
                a = k.c00 * f01y; bb = k.c01 * f10y; float e0y = a + bb;
                a = k.c10 * f01x; bb = k.c11 * f10x; float e1x = a - bb;
                a = k.c10 * f01y; bb = k.c11 * f10y; float e1y = a - bb;
                tap_t t0, t1;
                make_taps(&t1, x, y, e1x, e1y, H, W, coord_mode);
                make_taps(&t0, x, y, e0x, e0y, H, W, coord_mode);
                float g1x = 0, g1y = 0, g0x = 0, g0y = 0;
                for (int c = 0; c < 3; ++c) {
                    float gw1 = G[(3 + c) * npx + p], gw0 = G[(10 + c) * npx + p];
                    sample_plane_grad(I1 + c * npx, &t1, W, gw1, &g1x, &g1y);
                    sample_plane_grad(I0 + c * npx, &t0, W, gw0, &g0x, &g0y);
                    if (gimg6) {
                        scatter_plane(gI1 + c * npx, &t1, W, gw1);
                        scatter_plane(gI0 + c * npx, &t0, W, gw0);
                        gI1[c * npx + p] += G[(0 + c) * npx + p];
                        gI0[c * npx + p] += G[(13 + c) * npx + p];
                    }
                }
                if (gflow4) {
                    float de1x = G[6 * npx + p] + coord_grad_to_flow(g1x, W, coord_mode);
                    float de1y = G[7 * npx + p] + coord_grad_to_flow(g1y, H, coord_mode);
                    float de0x = G[8 * npx + p] + coord_grad_to_flow(g0x, W, coord_mode);
                    float de0y = G[9 * npx + p] + coord_grad_to_flow(g0y, H, coord_mode);
                    float* gF = gflow4 + (size_t)b * 4 * npx;
                    gF[p]           = k.c00 * de0x + k.c10 * de1x;    /* dF01 */
                    gF[npx + p]     = k.c00 * de0y + k.c10 * de1y;
                    gF[2 * npx + p] = k.c01 * de0x - k.c11 * de1x;    /* dF10 */
                    gF[3 * npx + p] = k.c01 * de0y - k.c11 * de1y;
                }
            }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a3 + a4: extract_outputs + compute_output_image (flow_interpolation.py:374-429)             */

static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

void ssm_oracle_fuse_fwd(const float* img6, const float* in16, const float* out5, const float* t,
                         float* out3, int B, int H, int W, int coord_mode)
{
    const size_t npx = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
        const coef_t k = make_coef(t[b]);
        const float* I0 = img6 + (size_t)b * 6 * npx;
        const float* I1 = I0 + 3 * npx;
        const float* X = in16 + (size_t)b * 16 * npx;
        const float* Y = out5 + (size_t)b * 5 * npx;
        float* O = out3 + (size_t)b * 3 * npx;
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                size_t p = (size_t)y * W + x;
                float v1 = sigmoidf_(Y[p]);                  /* :386-388 */
                float v0 = 1.0f - v1;                        /* :390 */
                float f1x = X[6 * npx + p] + Y[1 * npx + p]; /* :412 */
                float f1y = X[7 * npx + p] + Y[2 * npx + p];
                float f0x = X[8 * npx + p] + Y[3 * npx + p]; /* :413 */
                float f0y = X[9 * npx + p] + Y[4 * npx + p];
                tap_t t0, t1;
                make_taps(&t0, x, y, f0x, f0y, H, W, coord_mode);  /* :416 */
                make_taps(&t1, x, y, f1x, f1y, H, W, coord_mode);  /* :418 */
                float z0 = k.omt * v0, z1 = k.t * v1;
                float z = z0 + z1;                           /* :425 */
                for (int c = 0; c < 3; ++c) {
                    float w0 = v0 * sample_plane(I0 + c * npx, &t0, W);   /* :420 */
                    float w1 = v1 * sample_plane(I1 + c * npx, &t1, W);   /* :421 */
                    float a = k.omt * w0, bb = k.t * w1;
                    float s = a + bb;                        /* :423 */
                    O[c * npx + p] = s / z;                  /* :427 */
                }
            }
    }
}

/* Backward of compute_output_image.  g3 = dL/d(out3).  Outputs (any may be NULL):
 * gout5 [B,5,H,W], gin16 [B,16,H,W] (zero outside channels 6:10), gimg6 [B,6,H,W]. */
void ssm_oracle_fuse_bwd(const float* g3, const float* img6, const float* in16, const float* out5, const float* t,
                         float* gout5, float* gin16, float* gimg6, int B, int H, int W, int coord_mode)
{
    const size_t npx = (size_t)H * W;
    if (gimg6) memset(gimg6, 0, sizeof(float) * B * 6 * npx);
    if (gin16) memset(gin16, 0, sizeof(float) * B * 16 * npx);
    for (int b = 0; b < B; ++b) {
        const coef_t k = make_coef(t[b]);
        const float* I0 = img6 + (size_t)b * 6 * npx;
        const float* I1 = I0 + 3 * npx;
        const float* X = in16 + (size_t)b * 16 * npx;
        const float* Y = out5 + (size_t)b * 5 * npx;
        const float* G = g3 + (size_t)b * 3 * npx;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                size_t p = (size_t)y * W + x;
                float v1 = sigmoidf_(Y[p]);
                float v0 = 1.0f - v1;
                float f1x = X[6 * npx + p] + Y[1 * npx + p];
                float f1y = X[7 * npx + p] + Y[2 * npx + p];
                float f0x = X[8 * npx + p] + Y[3 * npx + p];
                float f0y = X[9 * npx + p] + Y[4 * npx + p];
                tap_t t0, t1;
                make_taps(&t0, x, y, f0x, f0y, H, W, coord_mode);
                make_taps(&t1, x, y, f1x, f1y, H, W, coord_mode);
                float z = k.omt * v0 + k.t * v1;
                float dz = 0.0f, dv0 = 0.0f, dv1 = 0.0f;
                float g0x = 0, g0y = 0, g1x = 0, g1y = 0;
                for (int c = 0; c < 3; ++c) {
                    float s0 = sample_plane(I0 + c * npx, &t0, W);
                    float s1 = sample_plane(I1 + c * npx, &t1, W);
                    float num = k.omt * (v0 * s0) + k.t * (v1 * s1);
                    float o = num / z;
                    float g = G[c * npx + p];
                    float ds = g / z;                 /* d/d(weighted_sum) */
                    dz -= g * o / z;                  /* d/d(normalization_factor) */
                    float dw0 = k.omt * ds;           /* d/d(pred_v_0t * warped0) */
                    float dw1 = k.t * ds;
                    dv0 += dw0 * s0; dv1 += dw1 * s1;
                    float ds0 = dw0 * v0, ds1 = dw1 * v1;   /* d/d(warped) */
                    sample_plane_grad(I0 + c * npx, &t0, W, ds0, &g0x, &g0y);
                    sample_plane_grad(I1 + c * npx, &t1, W, ds1, &g1x, &g1y);
                    if (gimg6) {
                        scatter_plane(gimg6 + ((size_t)b * 6 + c) * npx, &t0, W, ds0);
                        scatter_plane(gimg6 + ((size_t)b * 6 + 3 + c) * npx, &t1, W, ds1);
                    }
                }
                dv0 += k.omt * dz; dv1 += k.t * dz;
                float df1x = coord_grad_to_flow(g1x, W, coord_mode), df1y = coord_grad_to_flow(g1y, H, coord_mode);
                float df0x = coord_grad_to_flow(g0x, W, coord_mode), df0y = coord_grad_to_flow(g0y, H, coord_mode);
                if (gout5) {
                    float* gY = gout5 + (size_t)b * 5 * npx;
                    gY[p] = (dv1 - dv0) * (v1 * (1.0f - v1));       /* v0 = 1 - v1; sigmoid' */
                    gY[1 * npx + p] = df1x; gY[2 * npx + p] = df1y;
                    gY[3 * npx + p] = df0x; gY[4 * npx + p] = df0y;
                }
                if (gin16) {
                    float* gX = gin16 + (size_t)b * 16 * npx;
                    gX[6 * npx + p] = df1x; gX[7 * npx + p] = df1y;
                    gX[8 * npx + p] = df0x; gX[9 * npx + p] = df0y;
                }
            }
    }
}

int ssm_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
