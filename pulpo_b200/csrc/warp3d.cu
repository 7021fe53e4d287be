// warp3d.cu -- SpatialTransformer forward/backward (reference: src/network_blocks.py:101-121,
// ATen grid_sampler_3d / grid_sampler_3d_backward semantics: bilinear, border, align_corners=False),
// optionally fused with L2_reg of the displacement field (src/losses.py:208-222), which reads the
// same full-resolution field in the same step (src/models.py:160-162).
//
// HBM-bound streaming gather: at 20 B/voxel (C=1) the roofline leaves ~110 issue slots and ~1 L1-pipe
// cycle per voxel per SM, so the kernel is written for both: lane = x (every field load / store one
// cache line, corner gathers 1-2 lines; see the work-mapping note below), no 64-bit div/mod
// (FastDiv decode), 32-bit offsets, exact constant division in 5 FMAs, floor via an RZ add whose
// float bits index memory directly, a footprint that is always 2x2x2 in-bounds (fixed +1
// neighbours, no predicated gathers).
// Algorithmic bytes: fwd 12 + 8C per voxel; bwd 4C (gout) + 12 (df) + 4C (img) + 12 (gdf) [+ 4C gimg].
#include "common.cuh"

namespace pulpo {

// Work mapping.  A warp owns 32 consecutive voxels along D2 times WR consecutive rows along D1 at
// one plane; lane = x.  What bounds these kernels after HBM is the SM's L1 load/store pipe: a warp
// load that touches n cache lines occupies it for ~2n cycles.  With lane = x every field load and
// every store is one line, and each of the 8 corner gathers of a smooth field touches 1-2 lines
// (the previous mapping, 4 consecutive voxels per thread and 128-bit field loads, spread every
// gather over 4 lines: 2.3 pipe cycles per voxel against 1.0 now).  WR rows per thread keep
// several independent gathers in flight and let the fused regulariser take its y neighbours from
// registers and its x neighbours from warp shuffles.
#ifndef PULPO_WARP_WR
#define PULPO_WARP_WR 4
#endif
#ifndef PULPO_WARP_FWD_CTAS
#define PULPO_WARP_FWD_CTAS 3
#endif
#ifndef PULPO_WARP_BWD_CTAS
#define PULPO_WARP_BWD_CTAS 2
#endif
#ifndef PULPO_WARP_FWDG_CTAS
#define PULPO_WARP_FWDG_CTAS 2   // forward that also stores the interpolant's spatial gradient (dpos)
#endif
#ifndef PULPO_WARP_PF
#define PULPO_WARP_PF 1   // L2 prefetch of the next item's field rows
#endif
constexpr int WR = PULPO_WARP_WR;

struct WarpGeom {
    int B, C, D0, D1, D2, S;   // output / displacement-field grid
    int I1, I2, Si;            // sampled image: rows, row length, voxels per channel (== D1, D2, S unless the sizes differ)
    int nxb, nyb;            // 32-voxel blocks per row, WR-row blocks per plane
    unsigned int items;      // B * D0 * nyb * nxb  (one per warp)
    int unbias;
    FastDiv dnxb, dnyb, dD0;
    AxisConst2 a0, a1, a2;   // packed: the kernels process the WR rows of a thread as WR/2 pairs (FADD2 / FMUL2 / FFMA2)
};

static int make_geom(WarpGeom &g, int B, int C, int D0, int D1, int D2, int I0 = 0, int I1 = 0, int I2 = 0)
{
    if (I0 <= 0) { I0 = D0; I1 = D1; I2 = D2; }
    i64 S = (i64)D0 * D1 * D2, Si = (i64)I0 * I1 * I2;
    if (S >= (1ll << 31) || D0 > (1 << 22) || D1 > (1 << 22) || D2 > (1 << 22)) return PULPO_ERR_INVALID_SHAPE;
    if (Si >= (1ll << 31) || I0 > (1 << 22) || I1 > (1 << 22) || I2 > (1 << 22) || I0 < 2 || I1 < 2 || I2 < 2) return PULPO_ERR_INVALID_SHAPE;
    g.B = B; g.C = C; g.D0 = D0; g.D1 = D1; g.D2 = D2; g.S = (int)S;
    g.I1 = I1; g.I2 = I2; g.Si = (int)Si;
    g.nxb = (D2 + 31) / 32;
    g.nyb = (D1 + WR - 1) / WR;
    i64 items = (i64)B * D0 * g.nyb * g.nxb;
    if (items * 32 >= (1ll << 32)) return PULPO_ERR_INVALID_SHAPE;
    g.items = (unsigned int)items;
    g.unbias = tap_unbias(I1, I2);
    g.dnxb = make_fastdiv(g.nxb); g.dnyb = make_fastdiv(g.nyb); g.dD0 = make_fastdiv(D0);
    g.a0 = make_axis2(D0, I0); g.a1 = make_axis2(D1, I1); g.a2 = make_axis2(D2, I2);
    return PULPO_OK;
}

// 8 corners of one footprint (fixed neighbour offsets)
struct C8 {
    float c000, c001, c010, c011, c100, c101, c110, c111;
};

// strides as 64-bit BYTE offsets computed once per thread: each corner address is then one 64-bit add
// (the int-stride version spent ~35 integer instructions per voxel on sign extensions and carries)
struct GStride {
    i64 y, z, zy;
};
__device__ __forceinline__ GStride make_gstride(int sy, int sz)
{
    GStride s;
    s.y = (i64)sy * 4; s.z = (i64)sz * 4; s.zy = s.y + s.z;
    return s;
}

__device__ __forceinline__ C8 gather8(const float *im, int base, const GStride &st)
{
    const char *p = reinterpret_cast<const char *>(im) + (i64)base * 4;
    const float *p0 = reinterpret_cast<const float *>(p), *py = reinterpret_cast<const float *>(p + st.y);
    const float *pz = reinterpret_cast<const float *>(p + st.z), *pzy = reinterpret_cast<const float *>(p + st.zy);
    C8 k;
    k.c000 = __ldg(p0); k.c001 = __ldg(p0 + 1); k.c010 = __ldg(py); k.c011 = __ldg(py + 1);
    k.c100 = __ldg(pz); k.c101 = __ldg(pz + 1); k.c110 = __ldg(pzy); k.c111 = __ldg(pzy + 1);
    return k;
}

// ---- two voxels (rows y, y+1 of the same column and plane) per instruction
struct Foot2 {
    int base0, base1;
    float2 wx0, wx1, wy0, wy1, wz0, wz1;
    float2 uz, uy, ux;   // unclamped sample positions (backward)
};

template <int MODE>
__device__ __forceinline__ Foot2 make_foot2(float zf, float yf, float xf, float2 dz, float2 dy, float2 dx, const WarpGeom &g,
                                            int *fl = nullptr)
{
    const Tap2 tz = make_tap2<MODE>(splat2(zf), dz, g.a0);
    const Tap2 ty = make_tap2<MODE>(make_float2(yf, yf + 1.0f), dy, g.a1);
    const Tap2 tx = make_tap2<MODE>(splat2(xf), dx, g.a2);
    Foot2 f;
    f.base0 = (tz.bits0 * g.I1 + ty.bits0) * g.I2 + tx.bits0 - g.unbias;
    f.base1 = (tz.bits1 * g.I1 + ty.bits1) * g.I2 + tx.bits1 - g.unbias;
    f.wx0 = tx.w0; f.wx1 = tx.w1; f.wy0 = ty.w0; f.wy1 = ty.w1; f.wz0 = tz.w0; f.wz1 = tz.w1;
    f.uz = tz.u; f.uy = ty.u; f.ux = tx.u;
    if (fl) { fl[0] = tz.fl0; fl[1] = ty.fl0; fl[2] = tx.fl0; fl[3] = tz.fl1; fl[4] = ty.fl1; fl[5] = tx.fl1; }
    return f;
}

__device__ __forceinline__ float2 pair(float a, float b) { return make_float2(a, b); }

// trilinear interpolation of two voxels at once: same corner order and op order as the CPU grid sampler (products and
// sums rounded separately, no FMA), so each half is bit-identical to torch-CPU
__device__ __forceinline__ float2 interp8x2(const C8 &a, const C8 &b, const Foot2 &f)
{
    const float2 w00 = __fmul2_rn(f.wx0, f.wy0), w01 = __fmul2_rn(f.wx1, f.wy0);
    const float2 w10 = __fmul2_rn(f.wx0, f.wy1), w11 = __fmul2_rn(f.wx1, f.wy1);
    float2 acc = __fmul2_rn(pair(a.c000, b.c000), __fmul2_rn(w00, f.wz0));
    acc = add2_nofuse(acc, __fmul2_rn(pair(a.c001, b.c001), __fmul2_rn(w01, f.wz0)));
    acc = add2_nofuse(acc, __fmul2_rn(pair(a.c010, b.c010), __fmul2_rn(w10, f.wz0)));
    acc = add2_nofuse(acc, __fmul2_rn(pair(a.c011, b.c011), __fmul2_rn(w11, f.wz0)));
    acc = add2_nofuse(acc, __fmul2_rn(pair(a.c100, b.c100), __fmul2_rn(w00, f.wz1)));
    acc = add2_nofuse(acc, __fmul2_rn(pair(a.c101, b.c101), __fmul2_rn(w01, f.wz1)));
    acc = add2_nofuse(acc, __fmul2_rn(pair(a.c110, b.c110), __fmul2_rn(w10, f.wz1)));
    acc = add2_nofuse(acc, __fmul2_rn(pair(a.c111, b.c111), __fmul2_rn(w11, f.wz1)));
    return acc;
}

// spatial gradient of the interpolant at the two sample points (backward, gather half): difference along one axis,
// interpolated along the other two; FMA contraction is fine here (gradients are compared with a tolerance)
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, splat2(-1.0f), a); }
__device__ __forceinline__ float2 lerp_diff2(float2 d00, float2 d01, float2 d10, float2 d11, float2 u0, float2 u1, float2 v0,
                                             float2 v1)
{
    const float2 lo = __ffma2_rn(d01, u1, __fmul2_rn(d00, u0));
    const float2 hi = __ffma2_rn(d11, u1, __fmul2_rn(d10, u0));
    return __ffma2_rn(hi, v1, __fmul2_rn(lo, v0));
}

struct WItem {
    int b, z, y0, x;
    bool xok;
};

__device__ __forceinline__ WItem decode_witem(unsigned int w, int lane, const WarpGeom &g)
{
    unsigned int r, xb, r2, yb, b, z;
    fast_divmod(w, g.dnxb, r, xb);
    fast_divmod(r, g.dnyb, r2, yb);
    fast_divmod(r2, g.dD0, b, z);
    WItem t;
    t.b = (int)b; t.z = (int)z; t.y0 = (int)yb * WR; t.x = (int)xb * 32 + lane;
    t.xok = t.x < g.D2;
    return t;
}

// the WR values of one field channel this thread owns (0 outside the volume).  STREAM: read-once data
// (field without the fused regulariser, upstream gradient) bypasses L1 so that it does not evict the
// image lines the corner gathers live on.
template <bool STREAM>
__device__ __forceinline__ void load_rows(const float *f, int sy, const bool (&ok)[WR], float (&v)[WR])
{
#pragma unroll
    for (int j = 0; j < WR; ++j) v[j] = ok[j] ? (STREAM ? ld_stream(f + j * sy) : __ldg(f + j * sy)) : 0.0f;
}

// L2_reg of one field channel at the WR voxels this thread owns (forward differences on the
// [1:,1:,1:] crop, src/losses.py:217-221): sum of squared differences.  y neighbours come from the
// thread's own registers, x neighbours from the lane to the left.  `inner` (warp-uniform): the block
// lies inside the crop in z and y and inside the volume, so only the x edge needs a per-lane mask.
__device__ __forceinline__ float l2_fwd_terms(const float *f, const float (&c)[WR], const bool (&ok)[WR], const WItem &t,
                                              int lane, int sy, int sz, bool inner)
{
    float acc = 0.0f;
    if (inner) {
        // interior block: every neighbour address is inside the allocation (z > 0, y0 > 0), so the x
        // neighbour is simply the cached load one element to the left (the value read by lane 0 at x == 0
        // is masked) -- no shuffles, no predicated edge loads
        const float xin = t.x > 0 ? 1.0f : 0.0f;
        const char *row = reinterpret_cast<const char *>(f);
        const i64 sy8 = (i64)sy * 4, sz8 = (i64)sz * 4;
        float py = __ldg(reinterpret_cast<const float *>(row - sy8));
#pragma unroll
        for (int j = 0; j < WR; ++j, row += sy8) {
            const float left = __ldg(reinterpret_cast<const float *>(row) - 1);
            const float dzz = c[j] - __ldg(reinterpret_cast<const float *>(row - sz8)), dyy = c[j] - py, dxx = c[j] - left;
            acc += xin * (dzz * dzz + dyy * dyy + dxx * dxx);
            py = c[j];
        }
        return acc;
    }
    const bool zin = t.z > 0;
    // row above the block (needed by row 0)
    const float up = (ok[0] && t.y0 > 0) ? __ldg(f - sy) : 0.0f;
#pragma unroll
    for (int j = 0; j < WR; ++j) {
        float left = __shfl_up_sync(0xffffffffu, c[j], 1);
        if (lane == 0 && ok[j] && t.x > 0) left = __ldg(f + j * sy - 1);
        const bool in = ok[j] && zin && (t.y0 + j > 0) && (t.x > 0);
        if (in) {
            const float pz = __ldg(f + j * sy - sz);
            const float py = j > 0 ? c[j > 0 ? j - 1 : 0] : up;
            float d = c[j] - pz; acc += d * d;
            d = c[j] - py; acc += d * d;
            d = c[j] - left; acc += d * d;
        }
    }
    return acc;
}

// Spatial gradient of the interpolant at the two sample points of a pair, masked and scaled exactly like the gather
// half of the backward (zero wherever the border clamp is active; S_img/2 of the unnormalise times 2/(S-1) of the
// normalise): d out / d df along z, y, x.  With one image channel the whole gather half of the backward is
// gdf = gout * dpos, so a forward that stores dpos (GRAD) turns the backward into a streaming product (see
// warp3d_bwd_dpos_kernel and resize.cu's fused adjoint) -- no second position chain, no second round of gathers.
__device__ __forceinline__ void foot_grad2(const C8 &a, const C8 &b, const Foot2 &k, const WarpGeom &g, float2 &gz,
                                           float2 &gy, float2 &gx)
{
    const float2 c000 = pair(a.c000, b.c000), c001 = pair(a.c001, b.c001), c010 = pair(a.c010, b.c010),
                 c011 = pair(a.c011, b.c011), c100 = pair(a.c100, b.c100), c101 = pair(a.c101, b.c101),
                 c110 = pair(a.c110, b.c110), c111 = pair(a.c111, b.c111);
    const float2 sx = lerp_diff2(sub2(c001, c000), sub2(c011, c010), sub2(c101, c100), sub2(c111, c110), k.wy0, k.wy1,
                                 k.wz0, k.wz1);
    const float2 sy_ = lerp_diff2(sub2(c010, c000), sub2(c011, c001), sub2(c110, c100), sub2(c111, c101), k.wx0, k.wx1,
                                  k.wz0, k.wz1);
    const float2 sz_ = lerp_diff2(sub2(c100, c000), sub2(c101, c001), sub2(c110, c010), sub2(c111, c011), k.wx0, k.wx1,
                                  k.wy0, k.wy1);
    const float kz = 2.0f * g.a0.rcp.x, ky = 2.0f * g.a1.rcp.x, kx = 2.0f * g.a2.rcp.x;
    const float2 mz = pair((k.uz.x > 0.0f && k.uz.x < g.a0.Sm1.x) ? g.a0.gmul.x : 0.0f, (k.uz.y > 0.0f && k.uz.y < g.a0.Sm1.x) ? g.a0.gmul.x : 0.0f);
    const float2 my = pair((k.uy.x > 0.0f && k.uy.x < g.a1.Sm1.x) ? g.a1.gmul.x : 0.0f, (k.uy.y > 0.0f && k.uy.y < g.a1.Sm1.x) ? g.a1.gmul.x : 0.0f);
    const float2 mx = pair((k.ux.x > 0.0f && k.ux.x < g.a2.Sm1.x) ? g.a2.gmul.x : 0.0f, (k.ux.y > 0.0f && k.ux.y < g.a2.Sm1.x) ? g.a2.gmul.x : 0.0f);
    gz = __fmul2_rn(__fmul2_rn(mz, sz_), splat2(kz));
    gy = __fmul2_rn(__fmul2_rn(my, sy_), splat2(ky));
    gx = __fmul2_rn(__fmul2_rn(mx, sx), splat2(kx));
}

template <int MODE, bool IDX, bool REG, bool GRAD>
__global__ void __launch_bounds__(256, GRAD ? PULPO_WARP_FWDG_CTAS : PULPO_WARP_FWD_CTAS)
warp3d_fwd_kernel(const float *__restrict__ img, const float *__restrict__ df, float *__restrict__ out,
                  int32_t *__restrict__ idx, float *reg_out, ReduceWs *ws, double reg_scale, float *__restrict__ dpos,
                  const WarpGeom g)
{
    __shared__ double red[32];
    const unsigned int nwarps = gridDim.x * 8u;
    const int lane = threadIdx.x & 31;
    const int S = g.S, sy = g.D2, sz = g.D1 * g.D2;
    const GStride gst = make_gstride(g.I2, g.I1 * g.I2);
    float reg_acc = 0.0f;
    // persistent: whole waves of resident CTAs stride over the work items (one warp-item at a time)
    for (unsigned int w = (blockIdx.x * 256u + threadIdx.x) >> 5; w < g.items; w += nwarps) {
        const WItem t = decode_witem(w, lane, g);
        const int v0 = (t.z * g.D1 + t.y0) * g.D2 + t.x;
        bool ok[WR];
#pragma unroll
        for (int j = 0; j < WR; ++j) ok[j] = t.xok && (t.y0 + j < g.D1);
        const float *f = df + (i64)t.b * 3 * S + v0;
#if PULPO_WARP_PF
        if (w + nwarps < g.items && lane < 3 * WR) {
            // the field rows of this warp's NEXT item -> L2 (one 128-byte line per lane: channel lane / WR, row lane % WR),
            // so the first of the item's two dependent memory round trips is an L2 hit
            const WItem n = decode_witem(w + nwarps, 0, g);
            const i64 nv = (i64)n.b * 3 * S + ((i64)n.z * g.D1 + n.y0) * g.D2 + n.x;
            if (n.y0 + lane % WR < g.D1) prefetch_l2(df + nv + (i64)(lane / WR) * S + (lane % WR) * sy);
        }
#endif
        float dz[WR], dy[WR], dx[WR];
        load_rows<!REG>(f, sy, ok, dz);
        load_rows<!REG>(f + S, sy, ok, dy);
        load_rows<!REG>(f + 2 * S, sy, ok, dx);
        if (REG) {
            const bool inner = t.z > 0 && t.y0 > 0 && t.y0 + WR <= g.D1 && __all_sync(0xffffffffu, t.xok);
            reg_acc += l2_fwd_terms(f, dz, ok, t, lane, sy, sz, inner) + l2_fwd_terms(f + S, dy, ok, t, lane, sy, sz, inner) +
                       l2_fwd_terms(f + 2 * S, dx, ok, t, lane, sy, sz, inner);
        }
        const float zf = (float)t.z, xf = (float)t.x;
        Foot2 ft[WR / 2];
#pragma unroll
        for (int jp = 0; jp < WR / 2; ++jp) {
            const int j = 2 * jp;
            int fl[6];
            ft[jp] = make_foot2<MODE>(zf, (float)(t.y0 + j), xf, pair(dz[j], dz[j + 1]), pair(dy[j], dy[j + 1]),
                                      pair(dx[j], dx[j + 1]), g, IDX ? fl : nullptr);
            if (IDX) {
                int32_t *o = idx + (i64)t.b * 3 * S + v0 + j * sy;
                if (ok[j]) { o[0] = fl[0]; o[S] = fl[1]; o[2 * S] = fl[2]; }
                if (ok[j + 1]) { o[sy] = fl[3]; o[S + sy] = fl[4]; o[2 * S + sy] = fl[5]; }
            }
        }
        for (int c = 0; c < g.C; ++c) {
            const float *im = img + ((i64)t.b * g.C + c) * g.Si;
            float *o = out + ((i64)t.b * g.C + c) * S + v0;
            // all 8*WR corner gathers are issued before the first interpolation (no branch in between: the
            // footprint of a lane outside the volume is clamped in-bounds, so its loads are legal and just
            // unused) -- the kernel is bound by load latency and instruction issue, not by the L1 pipe
            C8 q[WR];
#pragma unroll
            for (int jp = 0; jp < WR / 2; ++jp) {
                q[2 * jp] = gather8(im, ft[jp].base0, gst);
                q[2 * jp + 1] = gather8(im, ft[jp].base1, gst);
            }
#pragma unroll
            for (int jp = 0; jp < WR / 2; ++jp) {
                const float2 res = interp8x2(q[2 * jp], q[2 * jp + 1], ft[jp]);
                if (ok[2 * jp]) o[(2 * jp) * sy] = res.x;
                if (ok[2 * jp + 1]) o[(2 * jp + 1) * sy] = res.y;
                if (GRAD) {   // C == 1 (checked by the entry point)
                    float2 gz, gy, gx;
                    foot_grad2(q[2 * jp], q[2 * jp + 1], ft[jp], g, gz, gy, gx);
                    float *d = dpos + (i64)t.b * 3 * S + v0 + (2 * jp) * sy;
                    if (ok[2 * jp]) { d[0] = gz.x; d[S] = gy.x; d[2 * S] = gx.x; }
                    if (ok[2 * jp + 1]) { d[sy] = gz.y; d[S + sy] = gy.y; d[2 * S + sy] = gx.y; }
                }
            }
        }
    }
    if (REG) {
        double bt = block_sum((double)reg_acc, red);
        grid_reduce_finish(bt, ws, reg_out, reg_scale, red);
    }
}

// gradient of L2_reg w.r.t. one field channel at the WR voxels this thread owns, gather form (see
// l2reg_bwd_v4_kernel in losses.cu): voxel v collects its own three differences if it is inside
// the crop, minus the difference of each forward neighbour that is inside the crop.  `inner`
// (warp-uniform): every row of the block has all four z / y neighbours inside the volume and the
// crop, which leaves the discrete Laplacian with per-lane masks for the two x edges.
__device__ __forceinline__ void l2_bwd_terms(const float *f, const float (&c)[WR], const bool (&ok)[WR], float (&r)[WR],
                                             const WItem &t, int lane, const WarpGeom &g, int sy, int sz, bool inner)
{
    const bool xin = t.x > 0, xn = t.x + 1 < g.D2;
    if (inner) {
        // interior block (all z / y neighbours and the rows around exist): x neighbours are cached loads at
        // +-1 element (masked at the two x faces), no shuffles, no predicated edge loads
        const float mi = xin ? 1.0f : 0.0f, mn = xn ? 1.0f : 0.0f;
        const char *row = reinterpret_cast<const char *>(f);
        const i64 sy8 = (i64)sy * 4, sz8 = (i64)sz * 4;
        float py = __ldg(reinterpret_cast<const float *>(row - sy8));
        const float down = __ldg(reinterpret_cast<const float *>(row + WR * sy8));
#pragma unroll
        for (int j = 0; j < WR; ++j, row += sy8) {
            const float *rp = reinterpret_cast<const float *>(row);
            const float left = __ldg(rp - 1), right = __ldg(rp + 1);
            const float cc = c[j];
            const float ny = j + 1 < WR ? c[j + 1 < WR ? j + 1 : 0] : down;
            const float nb = (__ldg(reinterpret_cast<const float *>(row - sz8)) + __ldg(reinterpret_cast<const float *>(row + sz8))) +
                             (py + ny);
            r[j] = mi * ((5.0f * cc - left) - nb) - mn * (right - cc);
            py = cc;
        }
        return;
    }
    const bool zin = t.z > 0, zn = t.z + 1 < g.D0;
    const float up = (ok[0] && t.y0 > 0) ? __ldg(f - sy) : 0.0f;
    const float down = (t.xok && t.y0 + WR < g.D1) ? __ldg(f + WR * sy) : 0.0f;
#pragma unroll
    for (int j = 0; j < WR; ++j) {
        float left = __shfl_up_sync(0xffffffffu, c[j], 1), right = __shfl_down_sync(0xffffffffu, c[j], 1);
        if (lane == 0 && ok[j] && xin) left = __ldg(f + j * sy - 1);
        if (lane == 31 && ok[j] && xn) right = __ldg(f + j * sy + 1);
        float a = 0.0f;
        if (ok[j]) {
            const bool yin = t.y0 + j > 0, yn = t.y0 + j + 1 < g.D1;
            const float cc = c[j];
            const float pz = zin ? __ldg(f + j * sy - sz) : 0.0f, nz = zn ? __ldg(f + j * sy + sz) : 0.0f;
            const float py = j > 0 ? c[j > 0 ? j - 1 : 0] : up;
            const float ny = j + 1 < WR ? c[j + 1 < WR ? j + 1 : 0] : down;
            if (xin && yin && zin) a += (cc - pz) + (cc - py) + (cc - left);
            if (zn && yin && xin) a -= nz - cc;
            if (yn && zin && xin) a -= ny - cc;
            if (xn && zin && yin) a -= right - cc;
        }
        r[j] = a;
    }
}

// Backward: gather half (gdf) always, scatter half (gimg) only when requested.  REG adds the
// gradient of the fused L2_reg term to gdf.
template <int MODE, bool SCATTER, bool REG>
__global__ void __launch_bounds__(256, PULPO_WARP_BWD_CTAS)
warp3d_bwd_kernel(const float *__restrict__ gout, const float *__restrict__ img, const float *__restrict__ df,
                  float *__restrict__ gimg, float *__restrict__ gdf, const float *__restrict__ reg_gloss,
                  float reg_k, const WarpGeom g)
{
    const unsigned int nwarps = gridDim.x * 8u;
    const int lane = threadIdx.x & 31;
    const int S = g.S, sy = g.D2, sz = g.D1 * g.D2;
    const int isy = g.I2, isz = g.I1 * g.I2;
    const GStride gst = make_gstride(isy, isz);
    const float reg_scale_k = REG ? (reg_gloss ? __ldg(reg_gloss) : 1.0f) * reg_k : 0.0f;
    for (unsigned int w = (blockIdx.x * 256u + threadIdx.x) >> 5; w < g.items; w += nwarps) {   // persistent, warp-uniform
    const WItem t = decode_witem(w, lane, g);
    const int v0 = (t.z * g.D1 + t.y0) * g.D2 + t.x;
    bool ok[WR];
#pragma unroll
    for (int j = 0; j < WR; ++j) ok[j] = t.xok && (t.y0 + j < g.D1);
    const float *f = df + (i64)t.b * 3 * S + v0;
    float dz[WR], dy[WR], dx[WR];
    load_rows<!REG>(f, sy, ok, dz);
    load_rows<!REG>(f + S, sy, ok, dy);
    load_rows<!REG>(f + 2 * S, sy, ok, dx);
    const float zf = (float)t.z, xf = (float)t.x;
    // autograd chain of 2*(loc/(S-1)-0.5) after the sampler's S/2:  (m*g*2)/(S-1)
    const float kz = 2.0f * g.a0.rcp.x, ky = 2.0f * g.a1.rcp.x, kx = 2.0f * g.a2.rcp.x;
    float2 rz2[WR / 2], ry2[WR / 2], rx2[WR / 2];
    Foot2 ft[WR / 2];
#pragma unroll
    for (int jp = 0; jp < WR / 2; ++jp) {
        const int j = 2 * jp;
        ft[jp] = make_foot2<MODE>(zf, (float)(t.y0 + j), xf, pair(dz[j], dz[j + 1]), pair(dy[j], dy[j + 1]),
                                  pair(dx[j], dx[j + 1]), g);
        rz2[jp] = ry2[jp] = rx2[jp] = splat2(0.0f);
    }
    for (int c = 0; c < g.C; ++c) {
        const i64 off = ((i64)t.b * g.C + c) * S, ioff = ((i64)t.b * g.C + c) * g.Si;
        float go[WR];
        load_rows<true>(gout + off + v0, sy, ok, go);
        C8 qq[WR];   // all gathers first (see the forward)
#pragma unroll
        for (int jp = 0; jp < WR / 2; ++jp) {
            qq[2 * jp] = gather8(img + ioff, ft[jp].base0, gst);
            qq[2 * jp + 1] = gather8(img + ioff, ft[jp].base1, gst);
        }
#pragma unroll
        for (int jp = 0; jp < WR / 2; ++jp) {
            const Foot2 &k = ft[jp];
            const C8 &a = qq[2 * jp], &b = qq[2 * jp + 1];
            const float2 c000 = pair(a.c000, b.c000), c001 = pair(a.c001, b.c001), c010 = pair(a.c010, b.c010),
                         c011 = pair(a.c011, b.c011), c100 = pair(a.c100, b.c100), c101 = pair(a.c101, b.c101),
                         c110 = pair(a.c110, b.c110), c111 = pair(a.c111, b.c111);
            // d/dx: difference along x, interpolated along y and z; likewise for y and z
            const float2 sx = lerp_diff2(sub2(c001, c000), sub2(c011, c010), sub2(c101, c100), sub2(c111, c110), k.wy0, k.wy1,
                                         k.wz0, k.wz1);
            const float2 sy_ = lerp_diff2(sub2(c010, c000), sub2(c011, c001), sub2(c110, c100), sub2(c111, c101), k.wx0, k.wx1,
                                          k.wz0, k.wz1);
            const float2 sz_ = lerp_diff2(sub2(c100, c000), sub2(c101, c001), sub2(c110, c010), sub2(c111, c011), k.wx0, k.wx1,
                                          k.wy0, k.wy1);
            const float2 go2 = pair(go[2 * jp], go[2 * jp + 1]);
            rx2[jp] = __ffma2_rn(sx, go2, rx2[jp]);
            ry2[jp] = __ffma2_rn(sy_, go2, ry2[jp]);
            rz2[jp] = __ffma2_rn(sz_, go2, rz2[jp]);
            if (SCATTER) {
                // Scatter half (grad w.r.t. the image): 8 weighted copies of the upstream gradient go to the footprint's
                // corners with red.global.add.f32.  Contention-aware: for a smooth field the footprint of lane l + 1
                // starts one voxel to the right of lane l's, so lane l + 1's (dx = 0) column IS lane l's (dx = 1)
                // column -- the same four addresses would be hit by both lanes.  Neighbouring lanes are combined
                // first (one shuffle per corner pair): lane l adds its right neighbour's dx = 0 contributions to its
                // own dx = 1 ones and the neighbour does not issue them: ~4.1 instead of 8 reductions per voxel, and no
                // two lanes of a warp instruction hit the same address.  Warp-uniform: every lane takes the shuffles.
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const bool on = ok[2 * jp + h];
                    const int base = h ? k.base1 : k.base0;
                    const int mybase = on ? base : -0x40000000;
                    const int right = __shfl_down_sync(0xffffffffu, mybase, 1);
                    const bool take = on && lane < 31 && right == base + 1;            // I carry my right neighbour's dx = 0 column
                    const bool given = __shfl_up_sync(0xffffffffu, (int)take, 1) != 0 && lane > 0;   // my dx = 0 column is carried by lane - 1
                    float *o = gimg + ioff + base;
                    const float wx0 = h ? k.wx0.y : k.wx0.x, wx1 = h ? k.wx1.y : k.wx1.x, wy0 = h ? k.wy0.y : k.wy0.x,
                                wy1 = h ? k.wy1.y : k.wy1.x, wz0 = h ? k.wz0.y : k.wz0.x, wz1 = h ? k.wz1.y : k.wz1.x;
                    const float gz[2] = {go[2 * jp + h] * wz0, go[2 * jp + h] * wz1};
                    const float wy[2] = {wy0, wy1};
#pragma unroll
                    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
                        for (int dy = 0; dy < 2; ++dy) {
                            const float v0 = (wx0 * wy[dy]) * gz[dz];
                            float v1 = (wx1 * wy[dy]) * gz[dz];
                            const float nv0 = __shfl_down_sync(0xffffffffu, on ? v0 : 0.0f, 1);
                            if (take) v1 += nv0;
                            if (on) {
                                float *q = o + dz * isz + dy * isy;
                                if (!given) atomicAdd(q, v0);
                                atomicAdd(q + 1, v1);
                            }
                        }
                }
            }
        }
    }
    float rz[WR], ry[WR], rx[WR];
#pragma unroll
    for (int jp = 0; jp < WR / 2; ++jp) {
        const Foot2 &k = ft[jp];
        // zero gradient wherever the border clamp is active (p <= 0 or p >= S-1)
        const float2 mz = pair((k.uz.x > 0.0f && k.uz.x < g.a0.Sm1.x) ? g.a0.gmul.x : 0.0f, (k.uz.y > 0.0f && k.uz.y < g.a0.Sm1.x) ? g.a0.gmul.x : 0.0f);
        const float2 my = pair((k.uy.x > 0.0f && k.uy.x < g.a1.Sm1.x) ? g.a1.gmul.x : 0.0f, (k.uy.y > 0.0f && k.uy.y < g.a1.Sm1.x) ? g.a1.gmul.x : 0.0f);
        const float2 mx = pair((k.ux.x > 0.0f && k.ux.x < g.a2.Sm1.x) ? g.a2.gmul.x : 0.0f, (k.ux.y > 0.0f && k.ux.y < g.a2.Sm1.x) ? g.a2.gmul.x : 0.0f);
        const float2 z2 = __fmul2_rn(__fmul2_rn(mz, rz2[jp]), splat2(kz));
        const float2 y2 = __fmul2_rn(__fmul2_rn(my, ry2[jp]), splat2(ky));
        const float2 x2 = __fmul2_rn(__fmul2_rn(mx, rx2[jp]), splat2(kx));
        rz[2 * jp] = z2.x; rz[2 * jp + 1] = z2.y;
        ry[2 * jp] = y2.x; ry[2 * jp + 1] = y2.y;
        rx[2 * jp] = x2.x; rx[2 * jp + 1] = x2.y;
    }
    if (REG) {
        const float k = reg_scale_k;
        const bool inner = t.z > 0 && t.z + 1 < g.D0 && t.y0 > 0 && t.y0 + WR < g.D1 && __all_sync(0xffffffffu, t.xok);
        float tt[WR];
        l2_bwd_terms(f, dz, ok, tt, t, lane, g, sy, sz, inner);
#pragma unroll
        for (int j = 0; j < WR; ++j) rz[j] += k * tt[j];
        l2_bwd_terms(f + S, dy, ok, tt, t, lane, g, sy, sz, inner);
#pragma unroll
        for (int j = 0; j < WR; ++j) ry[j] += k * tt[j];
        l2_bwd_terms(f + 2 * S, dx, ok, tt, t, lane, g, sy, sz, inner);
#pragma unroll
        for (int j = 0; j < WR; ++j) rx[j] += k * tt[j];
    }
    if (gdf) {
        float *o = gdf + (i64)t.b * 3 * S + v0;
#pragma unroll
        for (int j = 0; j < WR; ++j)
            if (ok[j]) {
                o[j * sy] = rz[j];
                o[S + j * sy] = ry[j];
                o[2 * S + j * sy] = rx[j];
            }
    }
    }   // persistent loop
}

// 8 warps per CTA; at most `per_sm` resident CTAs per SM, each striding over the items
static unsigned int persistent_grid(unsigned int items, int per_sm)
{
    int dev = 0, sms = kSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned int want = (items + 7) / 8, cap = (unsigned int)(sms * per_sm);
    return want < cap ? (want ? want : 1) : cap;
}

template <int MODE>
static int launch_fwd(const float *img, const float *df, float *out, int32_t *idx, float *reg_out, ReduceWs *ws,
                      double reg_scale, int B, int C, int D0, int D1, int D2, int I0, int I1, int I2, cudaStream_t st,
                      float *dpos = nullptr)
{
    WarpGeom g;
    int rc = make_geom(g, B, C, D0, D1, D2, I0, I1, I2);
    if (rc != PULPO_OK) return rc;
    const unsigned int grid = persistent_grid(g.items, dpos ? PULPO_WARP_FWDG_CTAS : PULPO_WARP_FWD_CTAS);
    if (dpos)
        warp3d_fwd_kernel<MODE, false, false, true><<<grid, 256, 0, st>>>(img, df, out, nullptr, nullptr, nullptr, 0.0, dpos, g);
    else if (idx)
        warp3d_fwd_kernel<MODE, true, false, false><<<grid, 256, 0, st>>>(img, df, out, idx, nullptr, nullptr, 0.0, nullptr, g);
    else if (reg_out)
        warp3d_fwd_kernel<MODE, false, true, false><<<grid, 256, 0, st>>>(img, df, out, nullptr, reg_out, ws, reg_scale, nullptr, g);
    else
        warp3d_fwd_kernel<MODE, false, false, false><<<grid, 256, 0, st>>>(img, df, out, nullptr, nullptr, nullptr, 0.0, nullptr, g);
    return launch_status();
}

template <int MODE>
static int launch_bwd(const float *gout, const float *img, const float *df, float *gimg, float *gdf,
                      const float *reg_gloss, float reg_k, bool reg, int B, int C, int D0, int D1, int D2, int I0, int I1,
                      int I2, cudaStream_t st)
{
    WarpGeom g;
    int rc = make_geom(g, B, C, D0, D1, D2, I0, I1, I2);
    if (rc != PULPO_OK) return rc;
    const unsigned int grid = persistent_grid(g.items, PULPO_WARP_BWD_CTAS);
    if (gimg && reg)
        warp3d_bwd_kernel<MODE, true, true><<<grid, 256, 0, st>>>(gout, img, df, gimg, gdf, reg_gloss, reg_k, g);
    else if (gimg)
        warp3d_bwd_kernel<MODE, true, false><<<grid, 256, 0, st>>>(gout, img, df, gimg, gdf, nullptr, 0.0f, g);
    else if (reg)
        warp3d_bwd_kernel<MODE, false, true><<<grid, 256, 0, st>>>(gout, img, df, gimg, gdf, reg_gloss, reg_k, g);
    else
        warp3d_bwd_kernel<MODE, false, false><<<grid, 256, 0, st>>>(gout, img, df, gimg, gdf, nullptr, 0.0f, g);
    return launch_status();
}

static int warp_fwd_dispatch(const float *img, const float *df, float *out, int32_t *idx, float *reg_out, void *ws,
                             double reg_scale, int B, int C, int D0, int D1, int D2, int coord_mode, cudaStream_t st,
                             int I0 = 0, int I1 = 0, int I2 = 0, float *dpos = nullptr)
{
    ReduceWs *w = (ReduceWs *)ws;
    if (coord_mode == PULPO_COORD_CPU_EXACT)
        return launch_fwd<0>(img, df, out, idx, reg_out, w, reg_scale, B, C, D0, D1, D2, I0, I1, I2, st, dpos);
    return launch_fwd<1>(img, df, out, idx, reg_out, w, reg_scale, B, C, D0, D1, D2, I0, I1, I2, st, dpos);
}

// gdf (+)= gout * dpos: the gather half of the backward once the forward has stored dpos (one image channel).
// Pure streaming, 128-bit when the volume allows.
template <bool ACC, typename V>
__global__ void __launch_bounds__(256)
warp3d_bwd_dpos_kernel(const V *__restrict__ gout, const V *__restrict__ dpos, V *__restrict__ gdf, unsigned int Sv,
                       unsigned int total)   // Sv: vectors per channel; total = B * Sv
{
    for (unsigned int i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
        const unsigned int b = i / Sv, v = i - b * Sv;
        const V go = gout[i];
        const size_t o = (size_t)b * 3 * Sv + v;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const V d = dpos[o + (size_t)c * Sv];
            V r;
            if constexpr (sizeof(V) == 16) {
                r.x = go.x * d.x; r.y = go.y * d.y; r.z = go.z * d.z; r.w = go.w * d.w;
                if (ACC) { const V old = gdf[o + (size_t)c * Sv]; r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w; }
            } else {
                r = go * d;
                if (ACC) r += gdf[o + (size_t)c * Sv];
            }
            gdf[o + (size_t)c * Sv] = r;
        }
    }
}

static int warp_bwd_dispatch(const float *gout, const float *img, const float *df, float *gimg, float *gdf,
                             const float *reg_gloss, float reg_k, bool reg, int B, int C, int D0, int D1, int D2,
                             int coord_mode, cudaStream_t st, int I0 = 0, int I1 = 0, int I2 = 0)
{
    if (coord_mode == PULPO_COORD_CPU_EXACT)
        return launch_bwd<0>(gout, img, df, gimg, gdf, reg_gloss, reg_k, reg, B, C, D0, D1, D2, I0, I1, I2, st);
    return launch_bwd<1>(gout, img, df, gimg, gdf, reg_gloss, reg_k, reg, B, C, D0, D1, D2, I0, I1, I2, st);
}

static double l2reg_scale(float lamb, int B, int D0, int D1, int D2)
{
    // mean over the [1:,1:,1:] crop of 3 channels, times lamb * D0*D1*D2 (src/losses.py:221-222)
    const double cnt = (double)B * 3 * (D0 - 1) * (double)(D1 - 1) * (D2 - 1);
    return (double)lamb * D0 * D1 * D2 / cnt;
}

}  // namespace pulpo

using namespace pulpo;

extern "C" int pulpo_warp3d_fwd(const float *img, const float *df, float *out, int32_t *idx_dbg, int B, int C,
                                int D0, int D1, int D2, int coord_mode, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_warp3d_fwd");
    PULPO_REQUIRE(img && df && out, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(coord_mode == 0 || coord_mode == 1, PULPO_ERR_UNSUPPORTED);
    return warp_fwd_dispatch(img, df, out, idx_dbg, nullptr, nullptr, 0.0, B, C, D0, D1, D2, coord_mode,
                             (cudaStream_t)stream);
}

extern "C" int pulpo_warp3d_bwd(const float *gout, const float *img, const float *df, float *gimg, float *gdf,
                                int B, int C, int D0, int D1, int D2, int coord_mode, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_warp3d_bwd");
    PULPO_REQUIRE(gout && img && df, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(gimg || gdf, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(coord_mode == 0 || coord_mode == 1, PULPO_ERR_UNSUPPORTED);
    return warp_bwd_dispatch(gout, img, df, gimg, gdf, nullptr, 0.0f, false, B, C, D0, D1, D2, coord_mode,
                             (cudaStream_t)stream);
}

extern "C" int pulpo_warp3d_fwd_dpos(const float *img, const float *df, float *out, float *dpos, int B, int D0, int D1,
                                     int D2, int coord_mode, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_warp3d_fwd_dpos");
    PULPO_REQUIRE(img && df && out && dpos, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(coord_mode == 0 || coord_mode == 1, PULPO_ERR_UNSUPPORTED);
    return warp_fwd_dispatch(img, df, out, nullptr, nullptr, nullptr, 0.0, B, 1, D0, D1, D2, coord_mode,
                             (cudaStream_t)stream, 0, 0, 0, dpos);
}

extern "C" int pulpo_warp3d_bwd_dpos(const float *gout, const float *dpos, float *gdf, int accumulate, int B, int D0,
                                     int D1, int D2, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_warp3d_bwd_dpos");
    PULPO_REQUIRE(gout && dpos && gdf, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && D0 >= 1 && D1 >= 1 && D2 >= 1, PULPO_ERR_INVALID_SHAPE);
    const i64 S = (i64)D0 * D1 * D2;
    PULPO_REQUIRE(S * B < (1ll << 31), PULPO_ERR_INVALID_SHAPE);
    cudaStream_t st = (cudaStream_t)stream;
    if (S % 4 == 0 && aligned16(gout) && aligned16(dpos) && aligned16(gdf)) {
        const unsigned int Sv = (unsigned int)(S / 4), total = Sv * (unsigned int)B;
        const int grid = grid_for(total, 256, 8);
        if (accumulate)
            warp3d_bwd_dpos_kernel<true, float4><<<grid, 256, 0, st>>>((const float4 *)gout, (const float4 *)dpos, (float4 *)gdf, Sv, total);
        else
            warp3d_bwd_dpos_kernel<false, float4><<<grid, 256, 0, st>>>((const float4 *)gout, (const float4 *)dpos, (float4 *)gdf, Sv, total);
    } else {
        const unsigned int Sv = (unsigned int)S, total = Sv * (unsigned int)B;
        const int grid = grid_for(total, 256, 8);
        if (accumulate)
            warp3d_bwd_dpos_kernel<true, float><<<grid, 256, 0, st>>>(gout, dpos, gdf, Sv, total);
        else
            warp3d_bwd_dpos_kernel<false, float><<<grid, 256, 0, st>>>(gout, dpos, gdf, Sv, total);
    }
    return launch_status();
}

extern "C" int pulpo_warp3d_fwd_img(const float *img, const float *df, float *out, int32_t *idx_dbg, int B, int C,
                                    int D0, int D1, int D2, int I0, int I1, int I2, int coord_mode, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_warp3d_fwd_img");
    PULPO_REQUIRE(img && df && out, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2 && I0 >= 2 && I1 >= 2 && I2 >= 2, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(coord_mode == 0 || coord_mode == 1, PULPO_ERR_UNSUPPORTED);
    return warp_fwd_dispatch(img, df, out, idx_dbg, nullptr, nullptr, 0.0, B, C, D0, D1, D2, coord_mode,
                             (cudaStream_t)stream, I0, I1, I2);
}

extern "C" int pulpo_warp3d_bwd_img(const float *gout, const float *img, const float *df, float *gimg, float *gdf,
                                    int B, int C, int D0, int D1, int D2, int I0, int I1, int I2, int coord_mode,
                                    pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_warp3d_bwd_img");
    PULPO_REQUIRE(gout && img && df, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(gimg || gdf, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2 && I0 >= 2 && I1 >= 2 && I2 >= 2, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(coord_mode == 0 || coord_mode == 1, PULPO_ERR_UNSUPPORTED);
    return warp_bwd_dispatch(gout, img, df, gimg, gdf, nullptr, 0.0f, false, B, C, D0, D1, D2, coord_mode,
                             (cudaStream_t)stream, I0, I1, I2);
}

extern "C" int pulpo_warp3d_l2reg_fwd(const float *img, const float *df, float *out, float lamb, float *reg_out,
                                      void *ws, size_t ws_bytes, int B, int C, int D0, int D1, int D2,
                                      int coord_mode, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_warp3d_l2reg_fwd");
    PULPO_REQUIRE(img && df && out && reg_out && ws, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(coord_mode == 0 || coord_mode == 1, PULPO_ERR_UNSUPPORTED);
    PULPO_REQUIRE(ws_bytes >= kReduceWsBytes, PULPO_ERR_WORKSPACE);
    return warp_fwd_dispatch(img, df, out, nullptr, reg_out, ws, l2reg_scale(lamb, B, D0, D1, D2), B, C, D0, D1, D2,
                             coord_mode, (cudaStream_t)stream);
}

extern "C" int pulpo_warp3d_l2reg_bwd(const float *gout, const float *img, const float *df, float *gdf, float lamb,
                                      const float *reg_gloss, int B, int C, int D0, int D1, int D2, int coord_mode,
                                      pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_warp3d_l2reg_bwd");
    PULPO_REQUIRE(gout && img && df && gdf, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(coord_mode == 0 || coord_mode == 1, PULPO_ERR_UNSUPPORTED);
    const float kk = (float)(2.0 * l2reg_scale(lamb, B, D0, D1, D2));
    return warp_bwd_dispatch(gout, img, df, nullptr, gdf, reg_gloss, kk, true, B, C, D0, D1, D2, coord_mode,
                             (cudaStream_t)stream);
}
