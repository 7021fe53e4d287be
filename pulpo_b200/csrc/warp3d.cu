// warp3d.cu -- SpatialTransformer forward/backward (reference: src/network_blocks.py:101-121,
// ATen grid_sampler_3d / grid_sampler_3d_backward semantics: bilinear, border, align_corners=False).
//
// HBM-bound streaming gather.  One thread owns VEC consecutive voxels along D2: the three
// displacement channels are read with 128-bit loads, the 8 corner gathers go through L1/L2
// (neighbouring voxels share cache lines), the output is written with 128-bit stores.
// At 20 B/voxel (C=1) the HBM roofline leaves ~110 issue slots per voxel per SM, so the kernel is
// written for instruction count: no 64-bit div/mod (FastDiv decode), 32-bit offsets, exact
// constant division in 5 FMAs, floor via an RZ add (no conversion unit), clamped corner offsets
// instead of predicated loads.
// Algorithmic bytes: fwd 12 + 8C per voxel; bwd 4C (gout) + 12 (df) + 4C (img) + 12 (gdf) [+ 4C gimg].
#include "common.cuh"

namespace pulpo {

struct WarpGeom {
    int B, C, D0, D1, D2, XG;
    unsigned int groups;
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
    g.dXG = make_fastdiv(g.XG); g.dD1 = make_fastdiv(D1); g.dD0 = make_fastdiv(D0);
    g.a0 = make_axis(D0); g.a1 = make_axis(D1); g.a2 = make_axis(D2);
    return PULPO_OK;
}

struct Foot {        // trilinear footprint of one voxel: 2x2x2 corners, always inside the volume
    int base;        // offset of the low corner inside one [D0,D1,D2] volume
    int hz, hy, hx;  // upper-border shifts (only needed to report floor(p))
    float wx0, wx1, wy0, wy1, wz0, wz1;
};

template <int MODE>
__device__ __forceinline__ Foot make_foot(float zf, float yf, float xf, float dz, float dy, float dx,
                                          const WarpGeom &g, float *uz = nullptr, float *uy = nullptr,
                                          float *ux = nullptr)
{
    Tap tz = make_tap<MODE>(zf, dz, g.a0, g.D0, uz);
    Tap ty = make_tap<MODE>(yf, dy, g.a1, g.D1, uy);
    Tap tx = make_tap<MODE>(xf, dx, g.a2, g.D2, ux);
    Foot f;
    f.base = (tz.i * g.D1 + ty.i) * g.D2 + tx.i;
    f.hz = tz.hi; f.hy = ty.hi; f.hx = tx.hi;
    f.wx0 = tx.w0; f.wx1 = tx.w1; f.wy0 = ty.w0; f.wy1 = ty.w1; f.wz0 = tz.w0; f.wz1 = tz.w1;
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
__device__ __forceinline__ void store_vec(float *p, const float (&v)[VEC])
{
    if (VEC == 4)
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
    else
        p[0] = v[0];
}

template <int MODE, int VEC>
__global__ void __launch_bounds__(256)
warp3d_fwd_kernel(const float *__restrict__ img, const float *__restrict__ df, float *__restrict__ out,
                  int32_t *__restrict__ idx, const WarpGeom g)
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
    load_vec<VEC>(f, dz);
    load_vec<VEC>(f + S, dy);
    load_vec<VEC>(f + 2 * S, dx);
    const float zf = (float)(int)z, yf = (float)(int)y, xf0 = (float)x0;

    Foot ft[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) ft[j] = make_foot<MODE>(zf, yf, xf0 + (float)j, dz[j], dy[j], dx[j], g);
    if (idx) {
        int32_t *o = idx + (i64)b * 3 * S + v0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            // recover floor(p) = (iz, iy, ix) from the (possibly border-shifted) base offset
            int rem = ft[j].base;
            int iz = rem / (g.D1 * g.D2);
            rem -= iz * g.D1 * g.D2;
            int iy = rem / g.D2;
            o[j] = iz + ft[j].hz; o[S + j] = iy + ft[j].hy; o[2 * S + j] = rem - iy * g.D2 + ft[j].hx;
        }
    }
    const int sy = g.D2, sz = g.D1 * g.D2;
    for (int c = 0; c < g.C; ++c) {
        const float *im = img + ((i64)b * g.C + c) * S;
        float res[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const Foot &k = ft[j];
            const float *p = im + k.base;
            const float *py = p + sy, *pz = p + sz, *pzy = pz + sy;
            const float c000 = __ldg(p), c001 = __ldg(p + 1), c010 = __ldg(py), c011 = __ldg(py + 1);
            const float c100 = __ldg(pz), c101 = __ldg(pz + 1), c110 = __ldg(pzy), c111 = __ldg(pzy + 1);
            // same corner order and op order as the CPU grid sampler: bit-identical to torch-CPU
            const float w00 = __fmul_rn(k.wx0, k.wy0), w01 = __fmul_rn(k.wx1, k.wy0);
            const float w10 = __fmul_rn(k.wx0, k.wy1), w11 = __fmul_rn(k.wx1, k.wy1);
            float acc = __fmul_rn(c000, __fmul_rn(w00, k.wz0));
            acc = __fadd_rn(acc, __fmul_rn(c001, __fmul_rn(w01, k.wz0)));
            acc = __fadd_rn(acc, __fmul_rn(c010, __fmul_rn(w10, k.wz0)));
            acc = __fadd_rn(acc, __fmul_rn(c011, __fmul_rn(w11, k.wz0)));
            acc = __fadd_rn(acc, __fmul_rn(c100, __fmul_rn(w00, k.wz1)));
            acc = __fadd_rn(acc, __fmul_rn(c101, __fmul_rn(w01, k.wz1)));
            acc = __fadd_rn(acc, __fmul_rn(c110, __fmul_rn(w10, k.wz1)));
            acc = __fadd_rn(acc, __fmul_rn(c111, __fmul_rn(w11, k.wz1)));
            res[j] = acc;
        }
        store_vec<VEC>(out + ((i64)b * g.C + c) * S + v0, res);
    }
}

// Backward: gather half (gdf) always, scatter half (gimg) only when requested.
template <int MODE, int VEC, bool SCATTER>
__global__ void __launch_bounds__(256)
warp3d_bwd_kernel(const float *__restrict__ gout, const float *__restrict__ img, const float *__restrict__ df,
                  float *__restrict__ gimg, float *__restrict__ gdf, const WarpGeom g)
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
    load_vec<VEC>(f, dz);
    load_vec<VEC>(f + S, dy);
    load_vec<VEC>(f + 2 * S, dx);
    const float zf = (float)(int)z, yf = (float)(int)y, xf0 = (float)x0;

    const int sy = g.D2, sz = g.D1 * g.D2;
    float rz[VEC], ry[VEC], rx[VEC];
    // autograd chain of 2*(loc/(S-1)-0.5) after the sampler's S/2:  (m*g*2)/(S-1)
    const float kz = 2.0f * g.a0.rcp, ky = 2.0f * g.a1.rcp, kx = 2.0f * g.a2.rcp;
    float go1[VEC];
    if (g.C == 1) load_vec<VEC>(gout + (i64)b * S + v0, go1);   // the common case: one 128-bit load
    // voxel-outer loop keeps only one footprint live (register pressure -> occupancy)
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        float uz, uy, ux;
        const Foot k = make_foot<MODE>(zf, yf, xf0 + (float)j, dz[j], dy[j], dx[j], g, &uz, &uy, &ux);
        // zero gradient wherever the border clamp is active (p <= 0 or p >= S-1)
        const float mz = (uz <= 0.0f || uz >= g.a0.Sm1) ? 0.0f : g.a0.gmul;
        const float my = (uy <= 0.0f || uy >= g.a1.Sm1) ? 0.0f : g.a1.gmul;
        const float mx = (ux <= 0.0f || ux >= g.a2.Sm1) ? 0.0f : g.a2.gmul;
        float gz = 0.0f, gy = 0.0f, gx = 0.0f;
        for (int c = 0; c < g.C; ++c) {
            const i64 off = ((i64)b * g.C + c) * S;
            const float go = (g.C == 1) ? go1[j] : __ldg(gout + off + v0 + j);
            const float *p = img + off + k.base;
            const float *py = p + sy, *pz = p + sz, *pzy = pz + sy;
            const float c000 = __ldg(p), c001 = __ldg(p + 1), c010 = __ldg(py), c011 = __ldg(py + 1);
            const float c100 = __ldg(pz), c101 = __ldg(pz + 1), c110 = __ldg(pzy), c111 = __ldg(pzy + 1);
            // d/dx: difference along x, interpolated along y and z; likewise for y and z
            const float sx = ((c001 - c000) * k.wy0 + (c011 - c010) * k.wy1) * k.wz0 +
                             ((c101 - c100) * k.wy0 + (c111 - c110) * k.wy1) * k.wz1;
            const float sy_ = ((c010 - c000) * k.wx0 + (c011 - c001) * k.wx1) * k.wz0 +
                              ((c110 - c100) * k.wx0 + (c111 - c101) * k.wx1) * k.wz1;
            const float sz_ = ((c100 - c000) * k.wx0 + (c101 - c001) * k.wx1) * k.wy0 +
                              ((c110 - c010) * k.wx0 + (c111 - c011) * k.wx1) * k.wy1;
            gx += sx * go;
            gy += sy_ * go;
            gz += sz_ * go;
            if (SCATTER) {
                float *q = gimg + off + k.base;
                const float w00 = k.wx0 * k.wy0, w01 = k.wx1 * k.wy0, w10 = k.wx0 * k.wy1, w11 = k.wx1 * k.wy1;
                const float g0 = go * k.wz0, g1 = go * k.wz1;
                atomicAdd(q, w00 * g0);
                atomicAdd(q + 1, w01 * g0);
                atomicAdd(q + sy, w10 * g0);
                atomicAdd(q + sy + 1, w11 * g0);
                atomicAdd(q + sz, w00 * g1);
                atomicAdd(q + sz + 1, w01 * g1);
                atomicAdd(q + sz + sy, w10 * g1);
                atomicAdd(q + sz + sy + 1, w11 * g1);
            }
        }
        rz[j] = (mz * gz) * kz;
        ry[j] = (my * gy) * ky;
        rx[j] = (mx * gx) * kx;
    }
    if (gdf) {
        float *o = gdf + (i64)b * 3 * S + v0;
        store_vec<VEC>(o, rz);
        store_vec<VEC>(o + S, ry);
        store_vec<VEC>(o + 2 * S, rx);
    }
}

template <int MODE, int VEC>
static int launch_fwd(const float *img, const float *df, float *out, int32_t *idx, int B, int C, int D0, int D1,
                      int D2, cudaStream_t st)
{
    WarpGeom g;
    int rc = make_geom(g, B, C, D0, D1, D2, VEC);
    if (rc != PULPO_OK) return rc;
    warp3d_fwd_kernel<MODE, VEC><<<(g.groups + 255) / 256, 256, 0, st>>>(img, df, out, idx, g);
    return launch_status();
}

template <int MODE, int VEC>
static int launch_bwd(const float *gout, const float *img, const float *df, float *gimg, float *gdf, int B, int C,
                      int D0, int D1, int D2, cudaStream_t st)
{
    WarpGeom g;
    int rc = make_geom(g, B, C, D0, D1, D2, VEC);
    if (rc != PULPO_OK) return rc;
    const unsigned int grid = (g.groups + 255) / 256;
    if (gimg)
        warp3d_bwd_kernel<MODE, VEC, true><<<grid, 256, 0, st>>>(gout, img, df, gimg, gdf, g);
    else
        warp3d_bwd_kernel<MODE, VEC, false><<<grid, 256, 0, st>>>(gout, img, df, gimg, gdf, g);
    return launch_status();
}

}  // namespace pulpo

using namespace pulpo;

extern "C" int pulpo_warp3d_fwd(const float *img, const float *df, float *out, int32_t *idx_dbg, int B, int C,
                                int D0, int D1, int D2, int coord_mode, pulpo_stream_t stream)
{
    PULPO_REQUIRE(img && df && out, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(coord_mode == 0 || coord_mode == 1, PULPO_ERR_UNSUPPORTED);
    cudaStream_t st = (cudaStream_t)stream;
    bool v4 = (D2 % 4 == 0) && aligned16(df) && aligned16(out) && !idx_dbg;
    if (coord_mode == PULPO_COORD_CPU_EXACT)
        return v4 ? launch_fwd<0, 4>(img, df, out, idx_dbg, B, C, D0, D1, D2, st)
                  : launch_fwd<0, 1>(img, df, out, idx_dbg, B, C, D0, D1, D2, st);
    return v4 ? launch_fwd<1, 4>(img, df, out, idx_dbg, B, C, D0, D1, D2, st)
              : launch_fwd<1, 1>(img, df, out, idx_dbg, B, C, D0, D1, D2, st);
}

extern "C" int pulpo_warp3d_bwd(const float *gout, const float *img, const float *df, float *gimg, float *gdf,
                                int B, int C, int D0, int D1, int D2, int coord_mode, pulpo_stream_t stream)
{
    PULPO_REQUIRE(gout && img && df, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(gimg || gdf, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(coord_mode == 0 || coord_mode == 1, PULPO_ERR_UNSUPPORTED);
    cudaStream_t st = (cudaStream_t)stream;
    bool v4 = (D2 % 4 == 0) && aligned16(df) && aligned16(gout) && (!gdf || aligned16(gdf));
    if (coord_mode == PULPO_COORD_CPU_EXACT)
        return v4 ? launch_bwd<0, 4>(gout, img, df, gimg, gdf, B, C, D0, D1, D2, st)
                  : launch_bwd<0, 1>(gout, img, df, gimg, gdf, B, C, D0, D1, D2, st);
    return v4 ? launch_bwd<1, 4>(gout, img, df, gimg, gdf, B, C, D0, D1, D2, st)
              : launch_bwd<1, 1>(gout, img, df, gimg, gdf, B, C, D0, D1, D2, st);
}
