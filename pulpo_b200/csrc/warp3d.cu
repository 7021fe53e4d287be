// warp3d.cu -- SpatialTransformer forward/backward (reference: src/network_blocks.py:101-121,
// ATen grid_sampler_3d / grid_sampler_3d_backward semantics: bilinear, border, align_corners=False).
//
// HBM-bound streaming gather.  One thread owns VEC consecutive voxels along D2: the three
// displacement channels are read with 128-bit loads, the 8 corner gathers go through L1/L2
// (neighbouring voxels share cache lines), the output is written with 128-bit stores.
// Algorithmic bytes: fwd 12 + 8C per voxel; bwd 4C (gout) + 12 (df) + 4C (img) + 12 (gdf) [+ 4C gimg].
#include "common.cuh"

namespace pulpo {

struct VoxTaps {
    Tap z, y, x;
    i64 base;
};

template <int MODE, int VEC>
__global__ void __launch_bounds__(256)
warp3d_fwd_kernel(const float *__restrict__ img, const float *__restrict__ df, float *__restrict__ out,
                  int32_t *__restrict__ idx, int B, int C, int D0, int D1, int D2, AxisConst a0,
                  AxisConst a1, AxisConst a2)
{
    const int XG = D2 / VEC;
    const i64 S = (i64)D0 * D1 * D2;
    const i64 sy = D2, sz = (i64)D1 * D2;
    const i64 groups = (i64)B * D0 * D1 * XG;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < groups; g += (i64)gridDim.x * blockDim.x) {
        int xg = (int)(g % XG);
        i64 r = g / XG;
        int y = (int)(r % D1);
        r /= D1;
        int z = (int)(r % D0);
        int b = (int)(r / D0);
        const int x0 = xg * VEC;
        const i64 v0 = ((i64)z * D1 + y) * D2 + x0;
        const float *f = df + (i64)b * 3 * S + v0;

        float dz[VEC], dy[VEC], dx[VEC];
        if (VEC == 4) {
            float4 t0 = ld_stream4(f), t1 = ld_stream4(f + S), t2 = ld_stream4(f + 2 * S);
            dz[0] = t0.x; dz[1 % VEC] = t0.y; dz[2 % VEC] = t0.z; dz[3 % VEC] = t0.w;
            dy[0] = t1.x; dy[1 % VEC] = t1.y; dy[2 % VEC] = t1.z; dy[3 % VEC] = t1.w;
            dx[0] = t2.x; dx[1 % VEC] = t2.y; dx[2 % VEC] = t2.z; dx[3 % VEC] = t2.w;
        } else {
            dz[0] = __ldg(f); dy[0] = __ldg(f + S); dx[0] = __ldg(f + 2 * S);
        }

        VoxTaps tp[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            tp[j].z = make_tap<MODE>(z, dz[j], a0, D0);
            tp[j].y = make_tap<MODE>(y, dy[j], a1, D1);
            tp[j].x = make_tap<MODE>(x0 + j, dx[j], a2, D2);
            tp[j].base = ((i64)tp[j].z.i * D1 + tp[j].y.i) * D2 + tp[j].x.i;
        }
        if (idx) {
            int32_t *o = idx + (i64)b * 3 * S + v0;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                o[j] = tp[j].z.i; o[S + j] = tp[j].y.i; o[2 * S + j] = tp[j].x.i;
            }
        }
        for (int c = 0; c < C; ++c) {
            const float *im = img + ((i64)b * C + c) * S;
            float res[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const Tap &tz = tp[j].z, &ty = tp[j].y, &tx = tp[j].x;
                const float *p = im + tp[j].base;
                // same corner order and op order as the CPU grid sampler: bit-exact vs torch-CPU
                float c000 = __ldg(p);
                float c001 = tx.in1 ? __ldg(p + 1) : 0.0f;
                float c010 = ty.in1 ? __ldg(p + sy) : 0.0f;
                float c011 = (ty.in1 && tx.in1) ? __ldg(p + sy + 1) : 0.0f;
                float c100 = tz.in1 ? __ldg(p + sz) : 0.0f;
                float c101 = (tz.in1 && tx.in1) ? __ldg(p + sz + 1) : 0.0f;
                float c110 = (tz.in1 && ty.in1) ? __ldg(p + sz + sy) : 0.0f;
                float c111 = (tz.in1 && ty.in1 && tx.in1) ? __ldg(p + sz + sy + 1) : 0.0f;
                float w00 = __fmul_rn(tx.w0, ty.w0), w01 = __fmul_rn(tx.w1, ty.w0);
                float w10 = __fmul_rn(tx.w0, ty.w1), w11 = __fmul_rn(tx.w1, ty.w1);
                float acc = __fmul_rn(c000, __fmul_rn(w00, tz.w0));
                acc = __fadd_rn(acc, __fmul_rn(c001, __fmul_rn(w01, tz.w0)));
                acc = __fadd_rn(acc, __fmul_rn(c010, __fmul_rn(w10, tz.w0)));
                acc = __fadd_rn(acc, __fmul_rn(c011, __fmul_rn(w11, tz.w0)));
                acc = __fadd_rn(acc, __fmul_rn(c100, __fmul_rn(w00, tz.w1)));
                acc = __fadd_rn(acc, __fmul_rn(c101, __fmul_rn(w01, tz.w1)));
                acc = __fadd_rn(acc, __fmul_rn(c110, __fmul_rn(w10, tz.w1)));
                acc = __fadd_rn(acc, __fmul_rn(c111, __fmul_rn(w11, tz.w1)));
                res[j] = acc;
            }
            float *o = out + ((i64)b * C + c) * S + v0;
            if (VEC == 4)
                *reinterpret_cast<float4 *>(o) = make_float4(res[0], res[1 % VEC], res[2 % VEC], res[3 % VEC]);
            else
                o[0] = res[0];
        }
    }
}

// Backward: gather half (gdf) always, scatter half (gimg) only when requested.
template <int MODE, int VEC, bool SCATTER>
__global__ void __launch_bounds__(256)
warp3d_bwd_kernel(const float *__restrict__ gout, const float *__restrict__ img, const float *__restrict__ df,
                  float *__restrict__ gimg, float *__restrict__ gdf, int B, int C, int D0, int D1, int D2,
                  AxisConst a0, AxisConst a1, AxisConst a2)
{
    const int XG = D2 / VEC;
    const i64 S = (i64)D0 * D1 * D2;
    const i64 sy = D2, sz = (i64)D1 * D2;
    const i64 groups = (i64)B * D0 * D1 * XG;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < groups; g += (i64)gridDim.x * blockDim.x) {
        int xg = (int)(g % XG);
        i64 r = g / XG;
        int y = (int)(r % D1);
        r /= D1;
        int z = (int)(r % D0);
        int b = (int)(r / D0);
        const int x0 = xg * VEC;
        const i64 v0 = ((i64)z * D1 + y) * D2 + x0;
        const float *f = df + (i64)b * 3 * S + v0;

        float dz[VEC], dy[VEC], dx[VEC];
        if (VEC == 4) {
            float4 t0 = ld_stream4(f), t1 = ld_stream4(f + S), t2 = ld_stream4(f + 2 * S);
            dz[0] = t0.x; dz[1 % VEC] = t0.y; dz[2 % VEC] = t0.z; dz[3 % VEC] = t0.w;
            dy[0] = t1.x; dy[1 % VEC] = t1.y; dy[2 % VEC] = t1.z; dy[3 % VEC] = t1.w;
            dx[0] = t2.x; dx[1 % VEC] = t2.y; dx[2 % VEC] = t2.z; dx[3 % VEC] = t2.w;
        } else {
            dz[0] = __ldg(f); dy[0] = __ldg(f + S); dx[0] = __ldg(f + 2 * S);
        }

        VoxTaps tp[VEC];
        float mz[VEC], my[VEC], mx[VEC], gz[VEC], gy[VEC], gx[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float uz, uy, ux;
            tp[j].z = make_tap<MODE>(z, dz[j], a0, D0, &uz);
            tp[j].y = make_tap<MODE>(y, dy[j], a1, D1, &uy);
            tp[j].x = make_tap<MODE>(x0 + j, dx[j], a2, D2, &ux);
            tp[j].base = ((i64)tp[j].z.i * D1 + tp[j].y.i) * D2 + tp[j].x.i;
            // zero gradient wherever the border clamp is active (p <= 0 or p >= S-1)
            mz[j] = (uz <= 0.0f || uz >= a0.Sm1) ? 0.0f : a0.gmul;
            my[j] = (uy <= 0.0f || uy >= a1.Sm1) ? 0.0f : a1.gmul;
            mx[j] = (ux <= 0.0f || ux >= a2.Sm1) ? 0.0f : a2.gmul;
            gz[j] = gy[j] = gx[j] = 0.0f;
        }
        for (int c = 0; c < C; ++c) {
            const i64 off = ((i64)b * C + c) * S;
            const float *im = img + off;
            float go[VEC];
            if (VEC == 4) {
                float4 t = ld_stream4(gout + off + v0);
                go[0] = t.x; go[1 % VEC] = t.y; go[2 % VEC] = t.z; go[3 % VEC] = t.w;
            } else {
                go[0] = __ldg(gout + off + v0);
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const Tap &tz = tp[j].z, &ty = tp[j].y, &tx = tp[j].x;
                const float *p = im + tp[j].base;
                float c000 = __ldg(p);
                float c001 = tx.in1 ? __ldg(p + 1) : 0.0f;
                float c010 = ty.in1 ? __ldg(p + sy) : 0.0f;
                float c011 = (ty.in1 && tx.in1) ? __ldg(p + sy + 1) : 0.0f;
                float c100 = tz.in1 ? __ldg(p + sz) : 0.0f;
                float c101 = (tz.in1 && tx.in1) ? __ldg(p + sz + 1) : 0.0f;
                float c110 = (tz.in1 && ty.in1) ? __ldg(p + sz + sy) : 0.0f;
                float c111 = (tz.in1 && ty.in1 && tx.in1) ? __ldg(p + sz + sy + 1) : 0.0f;
                // d/dx: difference along x, interpolated along y and z; likewise for y, z
                float ex0 = c001 - c000, ex1 = c011 - c010, ex2 = c101 - c100, ex3 = c111 - c110;
                float sx = (ex0 * ty.w0 + ex1 * ty.w1) * tz.w0 + (ex2 * ty.w0 + ex3 * ty.w1) * tz.w1;
                float ey0 = c010 - c000, ey1 = c011 - c001, ey2 = c110 - c100, ey3 = c111 - c101;
                float sy_ = (ey0 * tx.w0 + ey1 * tx.w1) * tz.w0 + (ey2 * tx.w0 + ey3 * tx.w1) * tz.w1;
                float ez0 = c100 - c000, ez1 = c101 - c001, ez2 = c110 - c010, ez3 = c111 - c011;
                float sz_ = (ez0 * tx.w0 + ez1 * tx.w1) * ty.w0 + (ez2 * tx.w0 + ez3 * tx.w1) * ty.w1;
                gx[j] += sx * go[j];
                gy[j] += sy_ * go[j];
                gz[j] += sz_ * go[j];
                if (SCATTER) {
                    float *q = gimg + off + tp[j].base;
                    float w00 = tx.w0 * ty.w0, w01 = tx.w1 * ty.w0, w10 = tx.w0 * ty.w1, w11 = tx.w1 * ty.w1;
                    float g0 = go[j] * tz.w0, g1 = go[j] * tz.w1;
                    atomicAdd(q, w00 * g0);
                    if (tx.in1) atomicAdd(q + 1, w01 * g0);
                    if (ty.in1) atomicAdd(q + sy, w10 * g0);
                    if (ty.in1 && tx.in1) atomicAdd(q + sy + 1, w11 * g0);
                    if (tz.in1) {
                        atomicAdd(q + sz, w00 * g1);
                        if (tx.in1) atomicAdd(q + sz + 1, w01 * g1);
                        if (ty.in1) atomicAdd(q + sz + sy, w10 * g1);
                        if (ty.in1 && tx.in1) atomicAdd(q + sz + sy + 1, w11 * g1);
                    }
                }
            }
        }
        if (gdf) {
            float *o = gdf + (i64)b * 3 * S + v0;
            float rz[VEC], ry[VEC], rx[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                // autograd chain of 2*(loc/(S-1)-0.5): (g*2)/(S-1), after the sampler's S/2
                rz[j] = __fdiv_rn((mz[j] * gz[j]) * 2.0f, a0.Sm1);
                ry[j] = __fdiv_rn((my[j] * gy[j]) * 2.0f, a1.Sm1);
                rx[j] = __fdiv_rn((mx[j] * gx[j]) * 2.0f, a2.Sm1);
            }
            if (VEC == 4) {
                *reinterpret_cast<float4 *>(o) = make_float4(rz[0], rz[1 % VEC], rz[2 % VEC], rz[3 % VEC]);
                *reinterpret_cast<float4 *>(o + S) = make_float4(ry[0], ry[1 % VEC], ry[2 % VEC], ry[3 % VEC]);
                *reinterpret_cast<float4 *>(o + 2 * S) = make_float4(rx[0], rx[1 % VEC], rx[2 % VEC], rx[3 % VEC]);
            } else {
                o[0] = rz[0]; o[S] = ry[0]; o[2 * S] = rx[0];
            }
        }
    }
}

template <int MODE, int VEC>
static int launch_fwd(const float *img, const float *df, float *out, int32_t *idx, int B, int C, int D0, int D1,
                      int D2, cudaStream_t st)
{
    i64 groups = (i64)B * D0 * D1 * (D2 / VEC);
    warp3d_fwd_kernel<MODE, VEC><<<grid_for(groups, 256), 256, 0, st>>>(img, df, out, idx, B, C, D0, D1, D2,
                                                                      make_axis(D0), make_axis(D1), make_axis(D2));
    return launch_status();
}

template <int MODE, int VEC>
static int launch_bwd(const float *gout, const float *img, const float *df, float *gimg, float *gdf, int B, int C,
                      int D0, int D1, int D2, cudaStream_t st)
{
    i64 groups = (i64)B * D0 * D1 * (D2 / VEC);
    int grid = grid_for(groups, 256);
    if (gimg)
        warp3d_bwd_kernel<MODE, VEC, true><<<grid, 256, 0, st>>>(gout, img, df, gimg, gdf, B, C, D0, D1, D2,
                                                                 make_axis(D0), make_axis(D1), make_axis(D2));
    else
        warp3d_bwd_kernel<MODE, VEC, false><<<grid, 256, 0, st>>>(gout, img, df, gimg, gdf, B, C, D0, D1, D2,
                                                                  make_axis(D0), make_axis(D1), make_axis(D2));
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
