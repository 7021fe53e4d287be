// warp3d.cu -- SpatialTransformer forward/backward (reference: src/network_blocks.py:101-121,
// ATen grid_sampler_3d / grid_sampler_3d_backward semantics: bilinear, border, align_corners=False),
// optionally fused with L2_reg of the displacement field (src/losses.py:208-222), which reads the
// same full-resolution field in the same step (src/models.py:160-162).
//
// HBM-bound streaming gather.  One thread owns VEC consecutive voxels along D2: the three
// displacement channels are read with 128-bit loads, the 8 corner gathers go through L1/L2
// (neighbouring voxels share cache lines), the output is written with 128-bit stores.
// At 20 B/voxel (C=1) the HBM roofline leaves ~110 issue slots per voxel per SM, so the kernel is
// written for instruction count: no 64-bit div/mod (FastDiv decode), 32-bit offsets, exact
// constant division in 5 FMAs, floor via an RZ add whose float bits index memory directly, a
// footprint that is always 2x2x2 in-bounds (fixed +1 neighbours, no predicated loads).
// Algorithmic bytes: fwd 12 + 8C per voxel; bwd 4C (gout) + 12 (df) + 4C (img) + 12 (gdf) [+ 4C gimg].
#include "common.cuh"

namespace pulpo {

struct WarpGeom {
    int B, C, D0, D1, D2, XG;
    unsigned int groups;
    int unbias;
    FastDiv dXG, dD1, dD0;
    AxisConst a0, a1, a2;
};

static int make_geom(WarpGeom &g, int B, int C, int D0, int D1, int D2, int vec)
{
    i64 S = (i64)D0 * D1 * D2;
    if (S >= (1ll << 31) || (i64)B * S / vec >= (1ll << 31) || D0 > (1 << 22) || D1 > (1 << 22) || D2 > (1 << 22))
        return PULPO_ERR_INVALID_SHAPE;
    g.B = B; g.C = C; g.D0 = D0; g.D1 = D1; g.D2 = D2; g.XG = D2 / vec;
    g.groups = (unsigned int)((i64)B * D0 * D1 * g.XG);
    g.unbias = tap_unbias(D1, D2);
    g.dXG = make_fastdiv(g.XG); g.dD1 = make_fastdiv(D1); g.dD0 = make_fastdiv(D0);
    g.a0 = make_axis(D0); g.a1 = make_axis(D1); g.a2 = make_axis(D2);
    return PULPO_OK;
}

struct Foot {        // trilinear footprint of one voxel: 2x2x2 corners, always inside the volume
    int base;        // offset of the low corner inside one [D0,D1,D2] volume
    float wx0, wx1, wy0, wy1, wz0, wz1;
};

template <int MODE>
__device__ __forceinline__ Foot make_foot(float zf, float yf, float xf, float dz, float dy, float dx,
                                          const WarpGeom &g, float *uz = nullptr, float *uy = nullptr,
                                          float *ux = nullptr, int *fl = nullptr)
{
    Tap tz = make_tap<MODE>(zf, dz, g.a0, uz);
    Tap ty = make_tap<MODE>(yf, dy, g.a1, uy);
    Tap tx = make_tap<MODE>(xf, dx, g.a2, ux);
    Foot f;
    f.base = tap_base(tz, ty, tx, g.D1, g.D2, g.unbias);
    f.wx0 = tx.w0; f.wx1 = tx.w1; f.wy0 = ty.w0; f.wy1 = ty.w1; f.wz0 = tz.w0; f.wz1 = tz.w1;
    if (fl) { fl[0] = tz.floor_p; fl[1] = ty.floor_p; fl[2] = tx.floor_p; }
    return f;
}

template <int VEC>
__device__ __forceinline__ void load_vec(const float *p, float (&v)[VEC])
{
    if (VEC == 4) {
        float4 t = ld_stream4(p);
        v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
    } else {
        v[0] = __ldg(p);
    }
}

template <int VEC>
__device__ __forceinline__ void load_vec_cached(const float *p, float (&v)[VEC])
{
    if (VEC == 4) {
        float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
    } else {
        v[0] = __ldg(p);
    }
}

template <int VEC>
__device__ __forceinline__ void store_vec(float *p, const float (&v)[VEC])
{
    if (VEC == 4)
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
    else
        p[0] = v[0];
}

// 8 corners of one footprint (fixed neighbour offsets)
struct C8 {
    float c000, c001, c010, c011, c100, c101, c110, c111;
};

__device__ __forceinline__ C8 gather8(const float *p, int sy, int sz)
{
    const float *py = p + sy, *pz = p + sz, *pzy = pz + sy;
    C8 k;
    k.c000 = __ldg(p); k.c001 = __ldg(p + 1); k.c010 = __ldg(py); k.c011 = __ldg(py + 1);
    k.c100 = __ldg(pz); k.c101 = __ldg(pz + 1); k.c110 = __ldg(pzy); k.c111 = __ldg(pzy + 1);
    return k;
}

// same corner order and op order as the CPU grid sampler: bit-identical to torch-CPU
__device__ __forceinline__ float interp8(const C8 &k, const Foot &f)
{
    const float w00 = __fmul_rn(f.wx0, f.wy0), w01 = __fmul_rn(f.wx1, f.wy0);
    const float w10 = __fmul_rn(f.wx0, f.wy1), w11 = __fmul_rn(f.wx1, f.wy1);
    float acc = __fmul_rn(k.c000, __fmul_rn(w00, f.wz0));
    acc = __fadd_rn(acc, __fmul_rn(k.c001, __fmul_rn(w01, f.wz0)));
    acc = __fadd_rn(acc, __fmul_rn(k.c010, __fmul_rn(w10, f.wz0)));
    acc = __fadd_rn(acc, __fmul_rn(k.c011, __fmul_rn(w11, f.wz0)));
    acc = __fadd_rn(acc, __fmul_rn(k.c100, __fmul_rn(w00, f.wz1)));
    acc = __fadd_rn(acc, __fmul_rn(k.c101, __fmul_rn(w01, f.wz1)));
    acc = __fadd_rn(acc, __fmul_rn(k.c110, __fmul_rn(w10, f.wz1)));
    acc = __fadd_rn(acc, __fmul_rn(k.c111, __fmul_rn(w11, f.wz1)));
    return acc;
}

// L2_reg of the field at the VEC voxels this thread owns (forward differences on the
// [1:,1:,1:] crop, src/losses.py:217-221): returns sum of squared differences
template <int VEC>
__device__ __forceinline__ float l2_fwd_terms(const float *f, const float (&c)[VEC], int x0, bool crop_zy, int sy, int sz)
{
    if (!crop_zy) return 0.0f;
    float pz[VEC], py[VEC];
    load_vec_cached<VEC>(f - sz, pz);
    load_vec_cached<VEC>(f - sy, py);
    float acc = 0.0f;
    if (x0 > 0) {
        const float px = __ldg(f - 1);
        float t = c[0] - pz[0]; acc += t * t;
        t = c[0] - py[0]; acc += t * t;
        t = c[0] - px; acc += t * t;
    }
#pragma unroll
    for (int j = 1; j < VEC; ++j) {
        float t = c[j] - pz[j]; acc += t * t;
        t = c[j] - py[j]; acc += t * t;
        t = c[j] - c[j - 1]; acc += t * t;
    }
    return acc;
}

template <int MODE, int VEC, bool IDX, bool REG>
__global__ void __launch_bounds__(256, REG ? 2 : 3)
warp3d_fwd_kernel(const float *__restrict__ img, const float *__restrict__ df, float *__restrict__ out,
                  int32_t *__restrict__ idx, float *reg_out, ReduceWs *ws, double reg_scale, const WarpGeom g)
{
    __shared__ double red[32];
    const unsigned int gid = blockIdx.x * 256u + threadIdx.x;
    float reg_acc = 0.0f;
    if (gid < g.groups) {
        unsigned int row, xg, zb, y, b, z;
        fast_divmod(gid, g.dXG, row, xg);
        fast_divmod(row, g.dD1, zb, y);
        fast_divmod(zb, g.dD0, b, z);
        const int S = g.D0 * g.D1 * g.D2;
        const int x0 = xg * VEC;
        const int v0 = (z * g.D1 + y) * g.D2 + x0;
        const int sy = g.D2, sz = g.D1 * g.D2;
        const float *f = df + (i64)b * 3 * S + v0;

        float dz[VEC], dy[VEC], dx[VEC];
        if (REG) {   // the field is re-read by neighbouring threads for the differences: keep it cached
            load_vec_cached<VEC>(f, dz);
            load_vec_cached<VEC>(f + S, dy);
            load_vec_cached<VEC>(f + 2 * S, dx);
        } else {
            load_vec<VEC>(f, dz);
            load_vec<VEC>(f + S, dy);
            load_vec<VEC>(f + 2 * S, dx);
        }
        if (REG) {   // first, while only the field values are live
            const bool crop = (z > 0 && y > 0);
            reg_acc = l2_fwd_terms<VEC>(f, dz, x0, crop, sy, sz) + l2_fwd_terms<VEC>(f + S, dy, x0, crop, sy, sz) +
                      l2_fwd_terms<VEC>(f + 2 * S, dx, x0, crop, sy, sz);
        }
        const float zf = (float)(int)z, yf = (float)(int)y, xf0 = (float)x0;

        Foot ft[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            int fl[3];
            ft[j] = make_foot<MODE>(zf, yf, xf0 + (float)j, dz[j], dy[j], dx[j], g, nullptr, nullptr, nullptr,
                                    IDX ? fl : nullptr);
            if (IDX) {
                int32_t *o = idx + (i64)b * 3 * S + v0 + j;
                o[0] = fl[0]; o[S] = fl[1]; o[2 * S] = fl[2];
            }
        }
        for (int c = 0; c < g.C; ++c) {
            const float *im = img + ((i64)b * g.C + c) * S;
            float res[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) res[j] = interp8(gather8(im + ft[j].base, sy, sz), ft[j]);
            store_vec<VEC>(out + ((i64)b * g.C + c) * S + v0, res);
        }
    }
    if (REG) {
        double bt = block_sum((double)reg_acc, red);
        grid_reduce_finish_atomic(bt, ws, reg_out, reg_scale);
    }
}

// gradient of L2_reg w.r.t. the field at the VEC voxels this thread owns, gather form (see
// l2reg_bwd_v4_kernel in losses.cu): voxel v collects its own three differences if it is inside
// the crop, minus the difference of each forward neighbour that is inside the crop
template <int VEC>
__device__ __forceinline__ void l2_bwd_terms(const float *f, const float (&c)[VEC], float (&r)[VEC], int x0, int D2,
                                             bool zin, bool yin, bool zn, bool yn, int sy, int sz)
{
    float pz[VEC], py[VEC], nz[VEC], ny[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) pz[j] = py[j] = nz[j] = ny[j] = 0.0f;
    if (zin) load_vec_cached<VEC>(f - sz, pz);
    if (yin) load_vec_cached<VEC>(f - sy, py);
    if (zn) load_vec_cached<VEC>(f + sz, nz);
    if (yn) load_vec_cached<VEC>(f + sy, ny);
    const float left = x0 > 0 ? __ldg(f - 1) : 0.0f;
    const float right = x0 + VEC < D2 ? __ldg(f + VEC) : 0.0f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int x = x0 + j;
        const bool xin = x > 0, xn = x + 1 < D2;
        const float cc = c[j];
        const float cl = j > 0 ? c[j > 0 ? j - 1 : 0] : left;
        const float cr = j + 1 < VEC ? c[j + 1 < VEC ? j + 1 : 0] : right;
        float a = 0.0f;
        if (xin && yin && zin) a += (cc - pz[j]) + (cc - py[j]) + (cc - cl);
        if (zn && yin && xin) a -= nz[j] - cc;
        if (yn && zin && xin) a -= ny[j] - cc;
        if (xn && zin && yin) a -= cr - cc;
        r[j] = a;
    }
}

// Backward: gather half (gdf) always, scatter half (gimg) only when requested.  REG adds the
// gradient of the fused L2_reg term to gdf.
template <int MODE, int VEC, bool SCATTER, bool REG>
__global__ void __launch_bounds__(256, 2)
warp3d_bwd_kernel(const float *__restrict__ gout, const float *__restrict__ img, const float *__restrict__ df,
                  float *__restrict__ gimg, float *__restrict__ gdf, const float *__restrict__ reg_gloss,
                  float reg_k, const WarpGeom g)
{
    const unsigned int gid = blockIdx.x * 256u + threadIdx.x;
    if (gid >= g.groups) return;
    unsigned int row, xg, zb, y, b, z;
    fast_divmod(gid, g.dXG, row, xg);
    fast_divmod(row, g.dD1, zb, y);
    fast_divmod(zb, g.dD0, b, z);
    const int S = g.D0 * g.D1 * g.D2;
    const int x0 = xg * VEC;
    const int v0 = (z * g.D1 + y) * g.D2 + x0;
    const float *f = df + (i64)b * 3 * S + v0;

    float dz[VEC], dy[VEC], dx[VEC];
    if (REG) {
        load_vec_cached<VEC>(f, dz);
        load_vec_cached<VEC>(f + S, dy);
        load_vec_cached<VEC>(f + 2 * S, dx);
    } else {
        load_vec<VEC>(f, dz);
        load_vec<VEC>(f + S, dy);
        load_vec<VEC>(f + 2 * S, dx);
    }
    const float zf = (float)(int)z, yf = (float)(int)y, xf0 = (float)x0;

    const int sy = g.D2, sz = g.D1 * g.D2;
    float rz[VEC], ry[VEC], rx[VEC];
    // autograd chain of 2*(loc/(S-1)-0.5) after the sampler's S/2:  (m*g*2)/(S-1)
    const float kz = 2.0f * g.a0.rcp, ky = 2.0f * g.a1.rcp, kx = 2.0f * g.a2.rcp;
    float go1[VEC];
    if (g.C == 1) load_vec<VEC>(gout + (i64)b * S + v0, go1);   // the common case: one 128-bit load
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        float uz, uy, ux;
        const Foot k = make_foot<MODE>(zf, yf, xf0 + (float)j, dz[j], dy[j], dx[j], g, &uz, &uy, &ux);
        // zero gradient wherever the border clamp is active (p <= 0 or p >= S-1)
        const float mz = (uz > 0.0f && uz < g.a0.Sm1) ? g.a0.gmul : 0.0f;
        const float my = (uy > 0.0f && uy < g.a1.Sm1) ? g.a1.gmul : 0.0f;
        const float mx = (ux > 0.0f && ux < g.a2.Sm1) ? g.a2.gmul : 0.0f;
        float gz = 0.0f, gy = 0.0f, gx = 0.0f;
        for (int c = 0; c < g.C; ++c) {
            const i64 off = ((i64)b * g.C + c) * S;
            const float go = (g.C == 1) ? go1[j] : __ldg(gout + off + v0 + j);
            const C8 q = gather8(img + off + k.base, sy, sz);
            // d/dx: difference along x, interpolated along y and z; likewise for y and z
            const float sx = ((q.c001 - q.c000) * k.wy0 + (q.c011 - q.c010) * k.wy1) * k.wz0 +
                             ((q.c101 - q.c100) * k.wy0 + (q.c111 - q.c110) * k.wy1) * k.wz1;
            const float sy_ = ((q.c010 - q.c000) * k.wx0 + (q.c011 - q.c001) * k.wx1) * k.wz0 +
                              ((q.c110 - q.c100) * k.wx0 + (q.c111 - q.c101) * k.wx1) * k.wz1;
            const float sz_ = ((q.c100 - q.c000) * k.wx0 + (q.c101 - q.c001) * k.wx1) * k.wy0 +
                              ((q.c110 - q.c010) * k.wx0 + (q.c111 - q.c011) * k.wx1) * k.wy1;
            gx += sx * go;
            gy += sy_ * go;
            gz += sz_ * go;
            if (SCATTER) {
                float *o = gimg + off + k.base;
                const float w00 = k.wx0 * k.wy0, w01 = k.wx1 * k.wy0, w10 = k.wx0 * k.wy1, w11 = k.wx1 * k.wy1;
                const float g0 = go * k.wz0, g1 = go * k.wz1;
                atomicAdd(o, w00 * g0);
                atomicAdd(o + 1, w01 * g0);
                atomicAdd(o + sy, w10 * g0);
                atomicAdd(o + sy + 1, w11 * g0);
                atomicAdd(o + sz, w00 * g1);
                atomicAdd(o + sz + 1, w01 * g1);
                atomicAdd(o + sz + sy, w10 * g1);
                atomicAdd(o + sz + sy + 1, w11 * g1);
            }
        }
        rz[j] = (mz * gz) * kz;
        ry[j] = (my * gy) * ky;
        rx[j] = (mx * gx) * kx;
    }
    if (REG) {
        const float k = (reg_gloss ? __ldg(reg_gloss) : 1.0f) * reg_k;
        const bool zin = z > 0, yin = y > 0, zn = (int)z + 1 < g.D0, yn = (int)y + 1 < g.D1;
        float t[VEC];
        l2_bwd_terms<VEC>(f, dz, t, x0, g.D2, zin, yin, zn, yn, sy, sz);
#pragma unroll
        for (int j = 0; j < VEC; ++j) rz[j] += k * t[j];
        l2_bwd_terms<VEC>(f + S, dy, t, x0, g.D2, zin, yin, zn, yn, sy, sz);
#pragma unroll
        for (int j = 0; j < VEC; ++j) ry[j] += k * t[j];
        l2_bwd_terms<VEC>(f + 2 * S, dx, t, x0, g.D2, zin, yin, zn, yn, sy, sz);
#pragma unroll
        for (int j = 0; j < VEC; ++j) rx[j] += k * t[j];
    }
    if (gdf) {
        float *o = gdf + (i64)b * 3 * S + v0;
        store_vec<VEC>(o, rz);
        store_vec<VEC>(o + S, ry);
        store_vec<VEC>(o + 2 * S, rx);
    }
}

template <int MODE, int VEC>
static int launch_fwd(const float *img, const float *df, float *out, int32_t *idx, float *reg_out, ReduceWs *ws,
                      double reg_scale, int B, int C, int D0, int D1, int D2, cudaStream_t st)
{
    WarpGeom g;
    int rc = make_geom(g, B, C, D0, D1, D2, VEC);
    if (rc != PULPO_OK) return rc;
    const unsigned int grid = (g.groups + 255) / 256;
    if (idx)
        warp3d_fwd_kernel<MODE, VEC, true, false><<<grid, 256, 0, st>>>(img, df, out, idx, nullptr, nullptr, 0.0, g);
    else if (reg_out)
        warp3d_fwd_kernel<MODE, VEC, false, true><<<grid, 256, 0, st>>>(img, df, out, nullptr, reg_out, ws, reg_scale, g);
    else
        warp3d_fwd_kernel<MODE, VEC, false, false><<<grid, 256, 0, st>>>(img, df, out, nullptr, nullptr, nullptr, 0.0, g);
    return launch_status();
}

template <int MODE, int VEC>
static int launch_bwd(const float *gout, const float *img, const float *df, float *gimg, float *gdf,
                      const float *reg_gloss, float reg_k, bool reg, int B, int C, int D0, int D1, int D2,
                      cudaStream_t st)
{
    WarpGeom g;
    int rc = make_geom(g, B, C, D0, D1, D2, VEC);
    if (rc != PULPO_OK) return rc;
    const unsigned int grid = (g.groups + 255) / 256;
    if (gimg && reg)
        warp3d_bwd_kernel<MODE, VEC, true, true><<<grid, 256, 0, st>>>(gout, img, df, gimg, gdf, reg_gloss, reg_k, g);
    else if (gimg)
        warp3d_bwd_kernel<MODE, VEC, true, false><<<grid, 256, 0, st>>>(gout, img, df, gimg, gdf, nullptr, 0.0f, g);
    else if (reg)
        warp3d_bwd_kernel<MODE, VEC, false, true><<<grid, 256, 0, st>>>(gout, img, df, gimg, gdf, reg_gloss, reg_k, g);
    else
        warp3d_bwd_kernel<MODE, VEC, false, false><<<grid, 256, 0, st>>>(gout, img, df, gimg, gdf, nullptr, 0.0f, g);
    return launch_status();
}

static int warp_fwd_dispatch(const float *img, const float *df, float *out, int32_t *idx, float *reg_out, void *ws,
                             double reg_scale, int B, int C, int D0, int D1, int D2, int coord_mode, cudaStream_t st)
{
    bool v4 = (D2 % 4 == 0) && aligned16(df) && aligned16(out);
    ReduceWs *w = (ReduceWs *)ws;
    if (coord_mode == PULPO_COORD_CPU_EXACT)
        return v4 ? launch_fwd<0, 4>(img, df, out, idx, reg_out, w, reg_scale, B, C, D0, D1, D2, st)
                  : launch_fwd<0, 1>(img, df, out, idx, reg_out, w, reg_scale, B, C, D0, D1, D2, st);
    return v4 ? launch_fwd<1, 4>(img, df, out, idx, reg_out, w, reg_scale, B, C, D0, D1, D2, st)
              : launch_fwd<1, 1>(img, df, out, idx, reg_out, w, reg_scale, B, C, D0, D1, D2, st);
}

static int warp_bwd_dispatch(const float *gout, const float *img, const float *df, float *gimg, float *gdf,
                             const float *reg_gloss, float reg_k, bool reg, int B, int C, int D0, int D1, int D2,
                             int coord_mode, cudaStream_t st)
{
    bool v4 = (D2 % 4 == 0) && aligned16(df) && aligned16(gout) && (!gdf || aligned16(gdf));
    if (coord_mode == PULPO_COORD_CPU_EXACT)
        return v4 ? launch_bwd<0, 4>(gout, img, df, gimg, gdf, reg_gloss, reg_k, reg, B, C, D0, D1, D2, st)
                  : launch_bwd<0, 1>(gout, img, df, gimg, gdf, reg_gloss, reg_k, reg, B, C, D0, D1, D2, st);
    return v4 ? launch_bwd<1, 4>(gout, img, df, gimg, gdf, reg_gloss, reg_k, reg, B, C, D0, D1, D2, st)
              : launch_bwd<1, 1>(gout, img, df, gimg, gdf, reg_gloss, reg_k, reg, B, C, D0, D1, D2, st);
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
    PULPO_REQUIRE(img && df && out, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(coord_mode == 0 || coord_mode == 1, PULPO_ERR_UNSUPPORTED);
    return warp_fwd_dispatch(img, df, out, idx_dbg, nullptr, nullptr, 0.0, B, C, D0, D1, D2, coord_mode,
                             (cudaStream_t)stream);
}

extern "C" int pulpo_warp3d_bwd(const float *gout, const float *img, const float *df, float *gimg, float *gdf,
                                int B, int C, int D0, int D1, int D2, int coord_mode, pulpo_stream_t stream)
{
    PULPO_REQUIRE(gout && img && df, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(gimg || gdf, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(coord_mode == 0 || coord_mode == 1, PULPO_ERR_UNSUPPORTED);
    return warp_bwd_dispatch(gout, img, df, gimg, gdf, nullptr, 0.0f, false, B, C, D0, D1, D2, coord_mode,
                             (cudaStream_t)stream);
}

extern "C" int pulpo_warp3d_l2reg_fwd(const float *img, const float *df, float *out, float lamb, float *reg_out,
                                      void *ws, size_t ws_bytes, int B, int C, int D0, int D1, int D2,
                                      int coord_mode, pulpo_stream_t stream)
{
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
    PULPO_REQUIRE(gout && img && df && gdf, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(coord_mode == 0 || coord_mode == 1, PULPO_ERR_UNSUPPORTED);
    const float kk = (float)(2.0 * l2reg_scale(lamb, B, D0, D1, D2));
    return warp_bwd_dispatch(gout, img, df, nullptr, gdf, reg_gloss, kk, true, B, C, D0, D1, D2, coord_mode,
                             (cudaStream_t)stream);
}
