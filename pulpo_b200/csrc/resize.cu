// resize.cu -- Laplacian-pyramid plumbing:
//   * ResizeTransform(factor>1) fused with DFAdder      src/network_blocks.py:138-158
//   * its exact adjoint in gather form (no atomics)     (autograd of the above)
//   * F.interpolate(size=...) target pyramid            src/losses.py:313
//   * avg_pool3d(2,2,ceil_mode=True) moving pyramid     src/components/pulpo.py:171-179
// All follow ATen's upsample_trilinear3d (align_corners=False) index/weight rules
// (SURVEY.md 9.4).  HBM-bound: one thread per output voxel, D2 innermost -> coalesced
// stores; the 8 input taps of neighbouring outputs share cache lines.
#include "common.cuh"

namespace pulpo {

struct LinTap {
    int i0, i1;
    float l0, l1;
};

// ATen: src = max(0, scale*(o+0.5)-0.5); i0 = min(floor(src), n-1); l1 = clamp(src-i0, 0, 1)
__device__ __forceinline__ LinTap lin_tap(int o, int n_in, int n_out, float scale)
{
    LinTap t;
    if (n_in == n_out) {
        t.i0 = o; t.i1 = o; t.l0 = 1.0f; t.l1 = 0.0f;
        return t;
    }
    float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)o, 0.5f)), 0.5f);
    src = src < 0.0f ? 0.0f : src;
    int i = (int)floorf(src);
    i = i > n_in - 1 ? n_in - 1 : i;
    float lam = __fsub_rn(src, (float)i);
    lam = fminf(fmaxf(lam, 0.0f), 1.0f);
    t.i0 = i;
    t.i1 = i + (i < n_in - 1 ? 1 : 0);
    t.l1 = lam;
    t.l0 = __fsub_rn(1.0f, lam);
    return t;
}

// out[bc, z, y, x] = trilinear(premul * in) (+ addend)
__global__ void __launch_bounds__(256)
trilinear_fwd_kernel(const float *__restrict__ in, const float *__restrict__ addend, float *__restrict__ out,
                     float premul, int BC, int i0n, int i1n, int i2n, int o0n, int o1n, int o2n, float s0, float s1,
                     float s2)
{
    const i64 Si = (i64)i0n * i1n * i2n, So = (i64)o0n * o1n * o2n;
    const i64 total = (i64)BC * So;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < total; g += (i64)gridDim.x * blockDim.x) {
        int x = (int)(g % o2n);
        i64 r = g / o2n;
        int y = (int)(r % o1n);
        r /= o1n;
        int z = (int)(r % o0n);
        int bc = (int)(r / o0n);
        LinTap tz = lin_tap(z, i0n, o0n, s0), ty = lin_tap(y, i1n, o1n, s1), tx = lin_tap(x, i2n, o2n, s2);
        const float *p = in + (i64)bc * Si;
        const float *r00 = p + ((i64)tz.i0 * i1n + ty.i0) * i2n, *r01 = p + ((i64)tz.i0 * i1n + ty.i1) * i2n;
        const float *r10 = p + ((i64)tz.i1 * i1n + ty.i0) * i2n, *r11 = p + ((i64)tz.i1 * i1n + ty.i1) * i2n;
        float a00 = premul * __ldg(r00 + tx.i0) * tx.l0 + premul * __ldg(r00 + tx.i1) * tx.l1;
        float a01 = premul * __ldg(r01 + tx.i0) * tx.l0 + premul * __ldg(r01 + tx.i1) * tx.l1;
        float a10 = premul * __ldg(r10 + tx.i0) * tx.l0 + premul * __ldg(r10 + tx.i1) * tx.l1;
        float a11 = premul * __ldg(r11 + tx.i0) * tx.l0 + premul * __ldg(r11 + tx.i1) * tx.l1;
        float v = (a00 * ty.l0 + a01 * ty.l1) * tz.l0 + (a10 * ty.l0 + a11 * ty.l1) * tz.l1;
        if (addend) v += __ldg(addend + g);
        out[g] = v;
    }
}

// weight with which output o (along one axis) reads input i
__device__ __forceinline__ float adj_w(int o, int i, int n_in, int n_out, float scale)
{
    LinTap t = lin_tap(o, n_in, n_out, scale);
    return (t.i0 == i ? t.l0 : 0.0f) + (t.i1 == i ? t.l1 : 0.0f);
}

// gx[bc, i] = scale * sum_o w(o -> i) * gout[bc, o]; each input voxel gathers from the <= 2f+2
// outputs per axis whose footprint touches it.
__global__ void __launch_bounds__(128)
upsample_bwd_kernel(const float *__restrict__ gout, float *__restrict__ gx, int f, float scale, int BC, int d0,
                    int d1, int d2, int accumulate)
{
    const int o0n = f * d0, o1n = f * d1, o2n = f * d2;
    const float s = 1.0f / (float)f;
    const i64 Si = (i64)d0 * d1 * d2, So = (i64)o0n * o1n * o2n;
    const i64 total = (i64)BC * Si;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < total; g += (i64)gridDim.x * blockDim.x) {
        int x = (int)(g % d2);
        i64 r = g / d2;
        int y = (int)(r % d1);
        r /= d1;
        int z = (int)(r % d0);
        int bc = (int)(r / d0);
        const float *go = gout + (i64)bc * So;
        const int zl = max(0, f * z - f / 2 - 1), zh = min(o0n - 1, f * z + (3 * f) / 2);
        const int yl = max(0, f * y - f / 2 - 1), yh = min(o1n - 1, f * y + (3 * f) / 2);
        const int xl = max(0, f * x - f / 2 - 1), xh = min(o2n - 1, f * x + (3 * f) / 2);
        float acc = 0.0f;
        for (int oz = zl; oz <= zh; ++oz) {
            float wz = adj_w(oz, z, d0, o0n, s);
            if (wz == 0.0f) continue;
            float accy = 0.0f;
            for (int oy = yl; oy <= yh; ++oy) {
                float wy = adj_w(oy, y, d1, o1n, s);
                if (wy == 0.0f) continue;
                const float *row = go + ((i64)oz * o1n + oy) * o2n;
                float accx = 0.0f;
                for (int ox = xl; ox <= xh; ++ox) accx += adj_w(ox, x, d2, o2n, s) * __ldg(row + ox);
                accy += wy * accx;
            }
            acc += wz * accy;
        }
        gx[g] = accumulate ? gx[g] + scale * acc : scale * acc;
    }
}

// ---------------------------------------------------------------------------------------------
// Specialised x2 path (every field resize on the level_res hot path).  For scale factor 2 the
// ATen taps are  out[2j] = .25*in[j-1] + .75*in[j]  (out[0] = in[0]),  out[2j+1] = .75*in[j] +
// .25*in[min(j+1,n-1)].  A thread produces a 2x2x4 block of outputs from a 3x3x4 block of inputs,
// interpolating x, then y, then z in registers (the same nesting order as ATen), so it issues
// ~2 loads and ~10 FP ops per output element instead of 8 loads + 3 tap computations.
struct Up2Geom {
    int BC, d0, d1, d2, XQ;   // XQ = d2 / 2 output quads per row
    unsigned int threads;
    FastDiv dXQ, dd1, dd0;
};

template <bool ADD>
__global__ void __launch_bounds__(128)
up2_fwd_kernel(const float *__restrict__ in, const float *__restrict__ addend, float *__restrict__ out, float premul,
               const Up2Geom g)
{
    const unsigned int gid = blockIdx.x * 128u + threadIdx.x;
    if (gid >= g.threads) return;
    unsigned int r, m, r2, k, bc, j;
    fast_divmod(gid, g.dXQ, r, m);
    fast_divmod(r, g.dd1, r2, k);
    fast_divmod(r2, g.dd0, bc, j);
    const int d0 = g.d0, d1 = g.d1, d2 = g.d2;
    const float *p = in + (i64)bc * d0 * d1 * d2;
    const int zr[3] = {max((int)j - 1, 0), (int)j, min((int)j + 1, d0 - 1)};
    const int yr[3] = {max((int)k - 1, 0), (int)k, min((int)k + 1, d1 - 1)};
    const int xc[4] = {max(2 * (int)m - 1, 0), 2 * (int)m, 2 * (int)m + 1, min(2 * (int)m + 2, d2 - 1)};
    // even-output tap weights on (prev, cur): (.25,.75), or (1,0) on the first output of an axis
    const float xe0 = m ? 0.25f : 1.0f, xe1 = m ? 0.75f : 0.0f;
    const float ye0 = k ? 0.25f : 1.0f, ye1 = k ? 0.75f : 0.0f;
    const float ze0 = j ? 0.25f : 1.0f, ze1 = j ? 0.75f : 0.0f;
    const int xa = m ? xc[0] : xc[1];   // first output of the row reads (in[0], in[0]) with weights (1,0)

    float Y[3][2][4];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float X[3][4];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const float *row = p + ((i64)zr[a] * d1 + yr[b]) * d2;
            const float v0 = premul * __ldg(row + xa), v1 = premul * __ldg(row + xc[1]);
            const float v2 = premul * __ldg(row + xc[2]), v3 = premul * __ldg(row + xc[3]);
            X[b][0] = v0 * xe0 + v1 * xe1;
            X[b][1] = v1 * 0.75f + v2 * 0.25f;
            X[b][2] = v1 * 0.25f + v2 * 0.75f;
            X[b][3] = v2 * 0.75f + v3 * 0.25f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            Y[a][0][q] = (k ? X[0][q] : X[1][q]) * ye0 + X[1][q] * ye1;
            Y[a][1][q] = X[1][q] * 0.75f + X[2][q] * 0.25f;
        }
    }
    const int o1 = 2 * d1, o2 = 2 * d2;
    const i64 obase = (((i64)bc * 2 * d0 + 2 * j) * o1 + 2 * k) * o2 + 4 * m;
#pragma unroll
    for (int yy = 0; yy < 2; ++yy) {
        float e[4], o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            e[q] = (j ? Y[0][yy][q] : Y[1][yy][q]) * ze0 + Y[1][yy][q] * ze1;
            o[q] = Y[1][yy][q] * 0.75f + Y[2][yy][q] * 0.25f;
        }
        const i64 oe = obase + (i64)yy * o2, oo = oe + (i64)o1 * o2;
        if (ADD) {
            const float4 ae = ld_stream4(addend + oe), ao = ld_stream4(addend + oo);
            e[0] += ae.x; e[1] += ae.y; e[2] += ae.z; e[3] += ae.w;
            o[0] += ao.x; o[1] += ao.y; o[2] += ao.z; o[3] += ao.w;
        }
        *reinterpret_cast<float4 *>(out + oe) = make_float4(e[0], e[1], e[2], e[3]);
        *reinterpret_cast<float4 *>(out + oo) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ---- x2 up-sampling, z-marching version (the default for factor 2).
// The kernel above recomputes the x/y interpolation of three input planes for every pair of output planes
// (36 loads, ~190 flops per 16 outputs; 70 % issue-bound in ncu).  Here a thread keeps the x/y-interpolated
// 2x4 blocks of the previous, current and next input plane in registers while it walks a run of planes:
// 12 loads and ~70 flops per 16 outputs.  Same operations in the same order per output value, so the
// results are bit-identical to the kernel above.
struct Up2FGeom {
    int BC, d0, d1, d2, XQ, zrun, nzrun;
    unsigned int threads;      // BC * nzrun * d1 * XQ
    FastDiv dXQ, dd1, dnz;
};

struct XY8 {
    float v[2][4];   // [output row 2k + yy][output column 4m + q]
};

__device__ __forceinline__ XY8 up2_plane_xy(const float *__restrict__ p, int z, int d1, int d2, int k, int m, float premul)
{
    const int yr[3] = {max(k - 1, 0), k, min(k + 1, d1 - 1)};
    const int xc1 = 2 * m, xc2 = 2 * m + 1, xc3 = min(2 * m + 2, d2 - 1);
    const int xa = m ? 2 * m - 1 : xc1;   // first output of the row reads (in[0], in[0]) with weights (1,0)
    const float xe0 = m ? 0.25f : 1.0f, xe1 = m ? 0.75f : 0.0f;
    const float ye0 = k ? 0.25f : 1.0f, ye1 = k ? 0.75f : 0.0f;
    float r[3][4];
#pragma unroll
    for (int b = 0; b < 3; ++b) {   // all 12 loads first
        const float *row = p + ((i64)z * d1 + yr[b]) * d2;
        r[b][0] = __ldg(row + xa); r[b][1] = __ldg(row + xc1); r[b][2] = __ldg(row + xc2); r[b][3] = __ldg(row + xc3);
    }
    float X[3][4];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const float v0 = premul * r[b][0], v1 = premul * r[b][1], v2 = premul * r[b][2], v3 = premul * r[b][3];
        X[b][0] = v0 * xe0 + v1 * xe1;
        X[b][1] = v1 * 0.75f + v2 * 0.25f;
        X[b][2] = v1 * 0.25f + v2 * 0.75f;
        X[b][3] = v2 * 0.75f + v3 * 0.25f;
    }
    XY8 o;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        o.v[0][q] = (k ? X[0][q] : X[1][q]) * ye0 + X[1][q] * ye1;
        o.v[1][q] = X[1][q] * 0.75f + X[2][q] * 0.25f;
    }
    return o;
}

template <bool ADD>
__global__ void __launch_bounds__(128)
up2_fwd_march_kernel(const float *__restrict__ in, const float *__restrict__ addend, float *__restrict__ out, float premul,
                     const Up2FGeom g)
{
    const unsigned int gid = blockIdx.x * 128u + threadIdx.x;
    if (gid >= g.threads) return;
    unsigned int r, m, r2, k, bc, zr;
    fast_divmod(gid, g.dXQ, r, m);
    fast_divmod(r, g.dd1, r2, k);
    fast_divmod(r2, g.dnz, bc, zr);
    const int d0 = g.d0, d1 = g.d1, d2 = g.d2;
    const int z0 = (int)zr * g.zrun, z1 = min(d0, z0 + g.zrun);
    const float *p = in + (i64)bc * d0 * d1 * d2;
    const int o1 = 2 * d1, o2 = 2 * d2;
    const i64 oplane = (i64)o1 * o2;
    i64 obase = (((i64)bc * 2 * d0 + 2 * z0) * o1 + 2 * (int)k) * o2 + 4 * (int)m;
    XY8 yp = up2_plane_xy(p, max(z0 - 1, 0), d1, d2, (int)k, (int)m, premul);
    XY8 yc = up2_plane_xy(p, z0, d1, d2, (int)k, (int)m, premul);
    for (int z = z0; z < z1; ++z, obase += 2 * oplane) {
        const XY8 yn = up2_plane_xy(p, min(z + 1, d0 - 1), d1, d2, (int)k, (int)m, premul);
        const float ze0 = z ? 0.25f : 1.0f, ze1 = z ? 0.75f : 0.0f;
#pragma unroll
        for (int yy = 0; yy < 2; ++yy) {
            float e[4], o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                e[q] = (z ? yp.v[yy][q] : yc.v[yy][q]) * ze0 + yc.v[yy][q] * ze1;
                o[q] = yc.v[yy][q] * 0.75f + yn.v[yy][q] * 0.25f;
            }
            const i64 oe = obase + (i64)yy * o2, oo = oe + oplane;
            if (ADD) {
                const float4 ae = ld_stream4(addend + oe), ao = ld_stream4(addend + oo);
                e[0] += ae.x; e[1] += ae.y; e[2] += ae.z; e[3] += ae.w;
                o[0] += ao.x; o[1] += ao.y; o[2] += ao.z; o[3] += ao.w;
            }
            *reinterpret_cast<float4 *>(out + oe) = make_float4(e[0], e[1], e[2], e[3]);
            *reinterpret_cast<float4 *>(out + oo) = make_float4(o[0], o[1], o[2], o[3]);
        }
        yp = yc;
        yc = yn;
    }
}

// Adjoint of the x2 up-sampling in gather form: input j collects
//   wa*go[2j-1] + wb*go[2j] + wc*go[2j+1] + wd*go[2j+2],  (wa..wd) = (.25,.75,.75,.25),
// (0,1,.75,.25) at j = 0 and (.25,.75,1,0) at j = n-1.  A thread produces 4 consecutive inputs
// along x from a 4x4x10 block of gout (two 128-bit loads + two scalars per row).
struct Up2BGeom {
    int BC, d0, d1, d2, XQ;   // XQ = d2 / 4
    unsigned int threads;
    FastDiv dXQ, dd1, dd0;
};

__device__ __forceinline__ void adj4(int j, int n, float (&w)[4])
{
    w[0] = j > 0 ? 0.25f : 0.0f;
    w[1] = j > 0 ? 0.75f : 1.0f;
    w[2] = j < n - 1 ? 0.75f : 1.0f;
    w[3] = j < n - 1 ? 0.25f : 0.0f;
}

template <bool ACC>
__global__ void __launch_bounds__(128)
up2_bwd_kernel(const float *__restrict__ gout, float *__restrict__ gx, float scale, const Up2BGeom g)
{
    const unsigned int gid = blockIdx.x * 128u + threadIdx.x;
    if (gid >= g.threads) return;
    unsigned int r, m, r2, y, bc, z;
    fast_divmod(gid, g.dXQ, r, m);
    fast_divmod(r, g.dd1, r2, y);
    fast_divmod(r2, g.dd0, bc, z);
    const int d0 = g.d0, d1 = g.d1, d2 = g.d2, o0 = 2 * d0, o1 = 2 * d1, o2 = 2 * d2;
    const float *go = gout + (i64)bc * o0 * o1 * o2;
    float wz[4], wy[4], wx[4][4];
    adj4((int)z, d0, wz);
    adj4((int)y, d1, wy);
#pragma unroll
    for (int t = 0; t < 4; ++t) adj4(4 * (int)m + t, d2, wx[t]);
    const int xl = 8 * (int)m;                        // aligned start of the 8 central outputs
    const int xm1 = max(xl - 1, 0), xp8 = min(xl + 8, o2 - 1);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int oz = min(max(2 * (int)z - 1 + a, 0), o0 - 1);
        float accy[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int oy = min(max(2 * (int)y - 1 + b, 0), o1 - 1);
            const float *row = go + ((i64)oz * o1 + oy) * o2;
            const float4 c0 = ld_stream4(row + xl), c1 = ld_stream4(row + xl + 4);
            const float v[10] = {__ldg(row + xm1), c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, __ldg(row + xp8)};
#pragma unroll
            for (int t = 0; t < 4; ++t)
                accy[t] += wy[b] * (wx[t][0] * v[2 * t] + wx[t][1] * v[2 * t + 1] + wx[t][2] * v[2 * t + 2] +
                                    wx[t][3] * v[2 * t + 3]);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[t] += wz[a] * accy[t];
    }
    float *o = gx + (((i64)bc * d0 + z) * d1 + y) * d2 + 4 * m;
    float4 res = make_float4(scale * acc[0], scale * acc[1], scale * acc[2], scale * acc[3]);
    if (ACC) {
        const float4 old = *reinterpret_cast<float4 *>(o);
        res.x += old.x; res.y += old.y; res.z += old.z; res.w += old.w;
    }
    *reinterpret_cast<float4 *>(o) = res;
}

// ---- x2 adjoint, z-marching version (the default for factor 2).
// The 4x4x10 gather above reads every gradient row through L1 four times (two y and two z neighbours
// need it) and is bound by the L1 pipe (81 % in ncu).  Here a warp owns 64 consecutive inputs along x
// (two per lane) times two rows, and walks a run of input planes: every gradient plane is turned ONCE
// into its x/y-adjoint T(oz) (six rows, one 128-bit load + two shuffles per row), and an input plane
// is the z-adjoint of four consecutive T's, two of which are carried in registers from the previous
// plane.  Gradient rows are read 1.5x instead of 4x, all loads are full cache lines.
constexpr int UB_RY = 2;       // input rows per thread
#ifndef PULPO_UP2B_PF
#define PULPO_UP2B_PF 0        // request the next iteration's gradient planes into L2 while this one is computed
#endif
struct Up2MGeom {
    int BC, d0, d1, d2;
    int nxb, nyb, zrun, nzrun;
    unsigned int items;
    FastDiv dnxb, dnyb, dnz;
};

struct T4 {
    float v[UB_RY][2];   // [row][x]
};

// the six gradient rows one plane contributes to this thread's 2 rows x 2 columns
struct PlaneRaw {
    float4 c[UB_RY + 4];              // gradients 2*x0 .. 2*x0+3 of rows 2*y0-1 .. 2*y0+4
    float e[UB_RY + 4];               // the warp's edge lanes only: gradient 2*x0-1 (lane 0) / 2*x0+4 (lane 31) -- one register
                                      // for both (the kernel is register-bound: separate arrays spilled 250 bytes)
};

// All loads of a plane are issued back to back (six independent 128-bit loads + the edge lanes'
// scalars); the shuffles and the arithmetic come later (a load placed after a shuffle waits for it:
// that version ran at 0.8 TB/s with one load in flight per warp).
// PROD: the gradient plane is formed on the fly as gm * plane -- the warp's gather-half backward gdf = gmoved * dpos
// (warp3d.cu, pulpo_warp3d_fwd_dpos), so the full-resolution field gradient never exists in HBM; `mul` is the same
// plane of the one-channel upstream gradient.
template <bool PROD>
__device__ __forceinline__ void up2_plane_load(PlaneRaw &p, const float *__restrict__ vol, const float *__restrict__ mvol,
                                               i64 plane_off, int o1, int o2, int y0, int x0, bool xok, int lane)
{
    const float *plane = vol + plane_off, *mul = PROD ? mvol + plane_off : nullptr;
    const bool need_l = (lane == 0) && xok && x0 > 0, need_r = (lane == 31) && xok && (2 * x0 + 4 < o2);
#pragma unroll
    for (int b = 0; b < UB_RY + 4; ++b) {
        int oy = 2 * y0 - 1 + b;
        oy = oy < 0 ? 0 : (oy > o1 - 1 ? o1 - 1 : oy);   // clamped rows replicate the edge rows (see up2_plane_adjoint)
        const i64 ro = (i64)oy * o2 + 2 * x0;
        p.c[b] = xok ? ld_stream4(plane + ro) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int eo = need_l ? -1 : 4;
        p.e[b] = (need_l || need_r) ? __ldg(plane + ro + eo) : 0.0f;
        if (PROD) {
            const float4 m = xok ? ld_stream4(mul + ro) : make_float4(0.f, 0.f, 0.f, 0.f);
            p.c[b].x *= m.x; p.c[b].y *= m.y; p.c[b].z *= m.z; p.c[b].w *= m.w;
            if (need_l || need_r) p.e[b] *= __ldg(mul + ro + eo);
        }
    }
}

// x/y adjoint of one gradient plane for this thread's 2 rows x 2 columns.  Constant weights (.25, .75, .75, .25) on
// every tap: at a face the out-of-range tap is the replicated edge gradient (.25 g + .75 g = g, the adjoint of ATen's
// edge clamp), so there are no per-thread weight tables (16 registers and a select per tap in a kernel that spilled).
// xfirst / xlast: this thread's first / second column is the first / last input of the row.
__device__ __forceinline__ T4 up2_plane_adjoint(const PlaneRaw &p, int lane, bool xfirst, bool xlast)
{
    float tx[UB_RY + 4][2];
#pragma unroll
    for (int b = 0; b < UB_RY + 4; ++b) {
        float left = __shfl_up_sync(0xffffffffu, p.c[b].w, 1), right = __shfl_down_sync(0xffffffffu, p.c[b].x, 1);
        if (lane == 0) left = p.e[b];
        if (lane == 31) right = p.e[b];
        if (xfirst) left = p.c[b].x;
        if (xlast) right = p.c[b].w;
        tx[b][0] = 0.25f * (left + p.c[b].z) + 0.75f * (p.c[b].x + p.c[b].y);
        tx[b][1] = 0.25f * (p.c[b].y + right) + 0.75f * (p.c[b].z + p.c[b].w);
    }
    T4 t;
#pragma unroll
    for (int r = 0; r < UB_RY; ++r)
#pragma unroll
        for (int q = 0; q < 2; ++q)   // rows 2r .. 2r+3 of the six: clamped row indices replicate the edge rows
            t.v[r][q] = 0.25f * (tx[2 * r][q] + tx[2 * r + 3][q]) + 0.75f * (tx[2 * r + 1][q] + tx[2 * r + 2][q]);
    return t;
}

template <bool ACC, bool PROD>
__global__ void __launch_bounds__(256, 2)
up2_bwd_march_kernel(const float *__restrict__ gout, const float *__restrict__ gmul, float *__restrict__ gx, float scale,
                     const Up2MGeom g)
{
    const unsigned int nwarps = gridDim.x * 8u;
    const int lane = threadIdx.x & 31;
    const int d0 = g.d0, d1 = g.d1, d2 = g.d2, o1 = 2 * d1, o2 = 2 * d2;
    for (unsigned int w = (blockIdx.x * 256u + threadIdx.x) >> 5; w < g.items; w += nwarps) {
        unsigned int r, xb, r2, yb, bc, zr;
        fast_divmod(w, g.dnxb, r, xb);
        fast_divmod(r, g.dnyb, r2, yb);
        fast_divmod(r2, g.dnz, bc, zr);
        const int x0 = ((int)xb * 32 + lane) * 2, y0 = (int)yb * UB_RY;
        const bool xok = x0 < d2;                       // d2 is even: x0 + 1 < d2 as well
        const bool xfirst = x0 == 0, xlast = x0 + 2 == d2;
        const int z0 = (int)zr * g.zrun, z1 = min(d0, z0 + g.zrun);
        const float *gb = gout + (i64)bc * 2 * d0 * o1 * o2;
        const float *gm = PROD ? gmul + (i64)(bc / 3u) * 2 * d0 * o1 * o2 : nullptr;   // fields have 3 channels
        const i64 plane = (i64)o1 * o2;
        // T(2z-1), T(2z) carried; T(2z+1), T(2z+2) computed per plane (both planes' loads in flight together).
        // Beyond the two z faces T replicates the face plane (see up2_plane_adjoint).
        PlaneRaw ra, rb;
        T4 ta, tb;
        if (PROD) {
            up2_plane_load<PROD>(rb, gb, gm, (i64)(2 * z0) * plane, o1, o2, y0, x0, xok, lane);
            tb = up2_plane_adjoint(rb, lane, xfirst, xlast);
            asm volatile("" ::: "memory");
            if (z0 > 0) up2_plane_load<PROD>(ra, gb, gm, (i64)(2 * z0 - 1) * plane, o1, o2, y0, x0, xok, lane);
            ta = (z0 > 0) ? up2_plane_adjoint(ra, lane, xfirst, xlast) : tb;
            asm volatile("" ::: "memory");
        } else {
            if (z0 > 0) up2_plane_load<PROD>(ra, gb, gm, (i64)(2 * z0 - 1) * plane, o1, o2, y0, x0, xok, lane);
            up2_plane_load<PROD>(rb, gb, gm, (i64)(2 * z0) * plane, o1, o2, y0, x0, xok, lane);
            tb = up2_plane_adjoint(rb, lane, xfirst, xlast);
            ta = (z0 > 0) ? up2_plane_adjoint(ra, lane, xfirst, xlast) : tb;
        }
        for (int z = z0; z < z1; ++z) {
#if PULPO_UP2B_PF
            if (xok && z + 1 < z1) {   // the next iteration's two gradient planes -> L2 (one request per row and lane quad)
#pragma unroll
                for (int b = 0; b < UB_RY + 4; ++b) {
                    int oy = 2 * y0 - 1 + b;
                    oy = oy < 0 ? 0 : (oy > o1 - 1 ? o1 - 1 : oy);
                    const float *row = gb + (i64)(2 * z + 3) * plane + (i64)oy * o2 + 2 * x0;
                    prefetch_l2(row);
                    if (z + 2 < d0) prefetch_l2(row + plane);
                }
            }
#endif
            up2_plane_load<PROD>(ra, gb, gm, (i64)(2 * z + 1) * plane, o1, o2, y0, x0, xok, lane);
            T4 tc, td;
            if (PROD) {
                // two sources per row: one plane at a time, or the 24 128-bit loads of both planes spill
                tc = up2_plane_adjoint(ra, lane, xfirst, xlast);
                asm volatile("" ::: "memory");
                if (z + 1 < d0) up2_plane_load<PROD>(rb, gb, gm, (i64)(2 * z + 2) * plane, o1, o2, y0, x0, xok, lane);
                td = (z + 1 < d0) ? up2_plane_adjoint(rb, lane, xfirst, xlast) : tc;
            } else {
                if (z + 1 < d0) up2_plane_load<PROD>(rb, gb, gm, (i64)(2 * z + 2) * plane, o1, o2, y0, x0, xok, lane);
                tc = up2_plane_adjoint(ra, lane, xfirst, xlast);
                td = (z + 1 < d0) ? up2_plane_adjoint(rb, lane, xfirst, xlast) : tc;
            }
#pragma unroll
            for (int rr = 0; rr < UB_RY; ++rr) {
                if (!xok || y0 + rr >= d1) continue;
                float2 res;
                res.x = scale * (0.25f * (ta.v[rr][0] + td.v[rr][0]) + 0.75f * (tb.v[rr][0] + tc.v[rr][0]));
                res.y = scale * (0.25f * (ta.v[rr][1] + td.v[rr][1]) + 0.75f * (tb.v[rr][1] + tc.v[rr][1]));
                float2 *o = reinterpret_cast<float2 *>(gx + (((i64)bc * d0 + z) * d1 + y0 + rr) * d2 + x0);
                if (ACC) {
                    const float2 old = *o;
                    res.x += old.x; res.y += old.y;
                }
                *o = res;
            }
            ta = tc;
            tb = td;
        }
    }
}

// avg_pool3d(kernel 2, stride 2, pad 0, ceil_mode=True): clipped windows, divide by clipped count
__global__ void __launch_bounds__(256)
avgpool2_kernel(const float *__restrict__ in, float *__restrict__ out, int BC, int D0, int D1, int D2)
{
    const int o0 = (D0 + 1) / 2, o1 = (D1 + 1) / 2, o2 = (D2 + 1) / 2;
    const i64 Si = (i64)D0 * D1 * D2, So = (i64)o0 * o1 * o2, total = (i64)BC * So;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < total; g += (i64)gridDim.x * blockDim.x) {
        int x = (int)(g % o2);
        i64 r = g / o2;
        int y = (int)(r % o1);
        r /= o1;
        int z = (int)(r % o0);
        int bc = (int)(r / o0);
        const int z1 = min(2 * z + 2, D0), y1 = min(2 * y + 2, D1), x1 = min(2 * x + 2, D2);
        float s = 0.0f;
        for (int a = 2 * z; a < z1; ++a)
            for (int b = 2 * y; b < y1; ++b)
                for (int c = 2 * x; c < x1; ++c) s += __ldg(in + (i64)bc * Si + ((i64)a * D1 + b) * D2 + c);
        int cnt = (z1 - 2 * z) * (y1 - 2 * y) * (x1 - 2 * x);
        out[g] = __fdiv_rn(s, (float)cnt);
    }
}

// The whole 2x pooling pyramid in one launch (Autoencoder.forward, src/components/pulpo.py:168-179: NL
// successive avg_pool3d(2,2)), for volumes whose sizes are multiples of 2^NL (all windows full, count 8).
// A CTA stages one E^3 block (E = 2^NL <= 16) of the input in shared memory with 128-bit loads and reduces it
// level by level; every level is written to its own output.  The input is read once (the chain of NL launches
// re-reads each level and is launch-latency bound below the first).  Same summation order per window as
// avgpool2_kernel (z, then y, then x) and the same division, so the results are bit-identical.
constexpr int POOL_MAXL = 4;
struct PoolPyr {
    float *out[POOL_MAXL];
    int nl, E;
    int BC, D0, D1, D2;
    int nb0, nb1, nb2;   // blocks per axis
};

__global__ void __launch_bounds__(256)
avgpool2_pyramid_kernel(const float *__restrict__ in, const PoolPyr p)
{
    __shared__ __align__(16) float sa[16 * 16 * 16];
    __shared__ __align__(16) float sb[8 * 8 * 8];
    const int E = p.E;
    unsigned int blk = blockIdx.x;
    const int bx = blk % p.nb2; blk /= p.nb2;
    const int by = blk % p.nb1; blk /= p.nb1;
    const int bz = blk % p.nb0;
    const int bc = blk / p.nb0;
    // stage the block: E*E rows of E floats (E / 4 quads per row)
    const int qpr = E >> 2, nquads = E * E * qpr;
    const float *src = in + (((i64)bc * p.D0 + (i64)bz * E) * p.D1 + (i64)by * E) * p.D2 + (i64)bx * E;
    for (int q = threadIdx.x; q < nquads; q += 256) {
        const int xq = q % qpr, row = q / qpr, y = row % E, z = row / E;
        const float4 v = ld_stream4(src + ((i64)z * p.D1 + y) * p.D2 + 4 * xq);
        *reinterpret_cast<float4 *>(sa + (z * E + y) * E + 4 * xq) = v;
    }
    __syncthreads();
    float *cur = sa, *nxt = sb;
    int e = E;   // edge of the block held in `cur`
#pragma unroll
    for (int l = 0; l < POOL_MAXL; ++l) {   // unrolled: constant indices into the parameter's pointer array
        if (l >= p.nl) break;
        const int h = e >> 1;
        const int d0 = p.D0 >> (l + 1), d1 = p.D1 >> (l + 1), d2 = p.D2 >> (l + 1);
        float *o = p.out[l] + (((i64)bc * d0 + (i64)bz * h) * d1 + (i64)by * h) * d2 + (i64)bx * h;
        for (int t = threadIdx.x; t < h * h * h; t += 256) {
            const int x = t % h, y = (t / h) % h, z = t / (h * h);
            float s = 0.0f;
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const float2 v = *reinterpret_cast<const float2 *>(cur + ((2 * z + a) * e + 2 * y + b) * e + 2 * x);
                    s += v.x;
                    s += v.y;
                }
            const float r = __fdiv_rn(s, 8.0f);
            o[((i64)z * d1 + y) * d2 + x] = r;
            nxt[(z * h + y) * h + x] = r;
        }
        __syncthreads();
        float *tmp = cur; cur = nxt; nxt = tmp;
        e = h;
    }
}

}  // namespace pulpo

using namespace pulpo;

extern "C" int pulpo_resize_up_fwd(const float *x, const float *addend, float *out, int factor, float scale, int B,
                                   int C, int d0, int d1, int d2, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_resize_up_fwd");
    PULPO_REQUIRE(x && out, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && d0 > 0 && d1 > 0 && d2 > 0, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(factor >= 2 && factor <= 64, PULPO_ERR_UNSUPPORTED);
    if (factor == 2 && (d2 % 2 == 0) && aligned16(out) && (!addend || aligned16(addend)) &&
        (i64)B * C * d0 * d1 * d2 < (1ll << 28)) {
        Up2FGeom g;
        g.BC = B * C; g.d0 = d0; g.d1 = d1; g.d2 = d2; g.XQ = d2 / 2;
        int dev = 0, sms = kSMs;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const i64 columns = (i64)g.BC * d1 * g.XQ;
        i64 per_col = ((i64)sms * 2048 + columns - 1) / columns;     // ~2048 threads per SM
        if (per_col < 1) per_col = 1;
        if (per_col > d0) per_col = d0;
        g.zrun = (int)((d0 + per_col - 1) / per_col);
        g.nzrun = (d0 + g.zrun - 1) / g.zrun;
        g.threads = (unsigned int)(columns * g.nzrun);
        g.dXQ = make_fastdiv(g.XQ); g.dd1 = make_fastdiv(d1); g.dnz = make_fastdiv(g.nzrun);
        const unsigned int grid = (g.threads + 127) / 128;
        if (addend)
            up2_fwd_march_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(x, addend, out, scale, g);
        else
            up2_fwd_march_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(x, addend, out, scale, g);
        return launch_status();
    }
    if (factor == 2 && (d2 % 2 == 0) && aligned16(out) && (!addend || aligned16(addend)) &&
        (i64)B * C * d0 * d1 * d2 < (1ll << 31)) {
        Up2Geom g;
        g.BC = B * C; g.d0 = d0; g.d1 = d1; g.d2 = d2; g.XQ = d2 / 2;
        g.threads = (unsigned int)((i64)B * C * d0 * d1 * g.XQ);
        g.dXQ = make_fastdiv(g.XQ); g.dd1 = make_fastdiv(d1); g.dd0 = make_fastdiv(d0);
        const unsigned int grid = (g.threads + 127) / 128;
        if (addend)
            up2_fwd_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(x, addend, out, scale, g);
        else
            up2_fwd_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(x, addend, out, scale, g);
        return launch_status();
    }
    float s = (float)(1.0 / (double)factor);
    i64 total = (i64)B * C * d0 * d1 * d2 * factor * factor * factor;
    trilinear_fwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        x, addend, out, scale, B * C, d0, d1, d2, factor * d0, factor * d1, factor * d2, s, s, s);
    return launch_status();
}

// preconditions of the z-marching x2 adjoint
static bool up2_march_ok(const void *gout, const void *gx, int factor, int BC, int d0, int d1, int d2)
{
    return factor == 2 && (d2 % 2 == 0) && aligned16(gout) && aligned16(gx) && d0 >= 2 &&
           (i64)BC * d0 * d1 * d2 * 8 < (1ll << 31);
}

// gmul != nullptr: the gradient is gout * gmul, formed on the fly (see up2_plane_load)
static int launch_up2_bwd_march(const float *gout, const float *gmul, float *gx, float scale, int accumulate, int BC, int d0,
                                int d1, int d2, cudaStream_t st)
{
    Up2MGeom g;
    g.BC = BC; g.d0 = d0; g.d1 = d1; g.d2 = d2;
    g.nxb = (d2 / 2 + 31) / 32;
    g.nyb = (d1 + UB_RY - 1) / UB_RY;
    int dev = 0, sms = kSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const i64 columns = (i64)g.BC * g.nyb * g.nxb;
    i64 grid = (columns * d0 + 7) / 8;                 // CTAs if every warp took one plane
    if (grid > (i64)sms * 2) grid = (i64)sms * 2;
    if (grid < 1) grid = 1;
    i64 per_col = (2 * grid * 8 + columns - 1) / columns;   // ~2 runs per resident warp
    if (per_col < 1) per_col = 1;
    if (per_col > d0) per_col = d0;
    g.zrun = (int)((d0 + per_col - 1) / per_col);
    g.nzrun = (d0 + g.zrun - 1) / g.zrun;
    g.items = (unsigned int)(columns * g.nzrun);
    g.dnxb = make_fastdiv(g.nxb); g.dnyb = make_fastdiv(g.nyb); g.dnz = make_fastdiv(g.nzrun);
    const unsigned int gr = (unsigned int)grid;
    if (gmul) {
        if (accumulate) up2_bwd_march_kernel<true, true><<<gr, 256, 0, st>>>(gout, gmul, gx, scale, g);
        else up2_bwd_march_kernel<false, true><<<gr, 256, 0, st>>>(gout, gmul, gx, scale, g);
    } else {
        if (accumulate) up2_bwd_march_kernel<true, false><<<gr, 256, 0, st>>>(gout, nullptr, gx, scale, g);
        else up2_bwd_march_kernel<false, false><<<gr, 256, 0, st>>>(gout, nullptr, gx, scale, g);
    }
    return launch_status();
}

extern "C" int pulpo_resize_up_bwd(const float *gout, float *gx, int factor, float scale, int accumulate, int B,
                                   int C, int d0, int d1, int d2, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_resize_up_bwd");
    PULPO_REQUIRE(gout && gx, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && d0 > 0 && d1 > 0 && d2 > 0, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(factor >= 2 && factor <= 64, PULPO_ERR_UNSUPPORTED);
    if (up2_march_ok(gout, gx, factor, B * C, d0, d1, d2))
        return launch_up2_bwd_march(gout, nullptr, gx, scale, accumulate, B * C, d0, d1, d2, (cudaStream_t)stream);
    if (factor == 2 && (d2 % 4 == 0) && aligned16(gout) && aligned16(gx) && (i64)B * C * d0 * d1 * d2 < (1ll << 31)) {
        Up2BGeom g;
        g.BC = B * C; g.d0 = d0; g.d1 = d1; g.d2 = d2; g.XQ = d2 / 4;
        g.threads = (unsigned int)((i64)B * C * d0 * d1 * g.XQ);
        g.dXQ = make_fastdiv(g.XQ); g.dd1 = make_fastdiv(d1); g.dd0 = make_fastdiv(d0);
        if (accumulate)
            up2_bwd_kernel<true><<<(g.threads + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gout, gx, scale, g);
        else
            up2_bwd_kernel<false><<<(g.threads + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gout, gx, scale, g);
        return launch_status();
    }
    i64 total = (i64)B * C * d0 * d1 * d2;
    upsample_bwd_kernel<<<grid_for(total, 128, 16), 128, 0, (cudaStream_t)stream>>>(gout, gx, factor, scale, B * C,
                                                                                  d0, d1, d2, accumulate);
    return launch_status();
}

extern "C" int pulpo_resize_up2_bwd_dpos(const float *gout, const float *dpos, float *gx, float scale, int accumulate,
                                         int B, int d0, int d1, int d2, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_resize_up2_bwd_dpos");
    PULPO_REQUIRE(gout && dpos && gx, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && d0 > 0 && d1 > 0 && d2 > 0, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(up2_march_ok(dpos, gx, 2, B * 3, d0, d1, d2) && aligned16(gout), PULPO_ERR_UNSUPPORTED);
    return launch_up2_bwd_march(dpos, gout, gx, scale, accumulate, B * 3, d0, d1, d2, (cudaStream_t)stream);
}

extern "C" int pulpo_interp_size_fwd(const float *x, float *out, int B, int C, int i0, int i1, int i2, int o0,
                                     int o1, int o2, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_interp_size_fwd");
    PULPO_REQUIRE(x && out, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && i0 > 0 && i1 > 0 && i2 > 0 && o0 > 0 && o1 > 0 && o2 > 0,
                  PULPO_ERR_INVALID_SHAPE);
    i64 total = (i64)B * C * o0 * o1 * o2;
    trilinear_fwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        x, nullptr, out, 1.0f, B * C, i0, i1, i2, o0, o1, o2, (float)i0 / (float)o0, (float)i1 / (float)o1,
        (float)i2 / (float)o2);
    return launch_status();
}

extern "C" int pulpo_avgpool2_fwd(const float *x, float *out, int B, int C, int D0, int D1, int D2,
                                  pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_avgpool2_fwd");
    PULPO_REQUIRE(x && out, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 > 0 && D1 > 0 && D2 > 0, PULPO_ERR_INVALID_SHAPE);
    i64 total = (i64)B * C * ((D0 + 1) / 2) * ((D1 + 1) / 2) * ((D2 + 1) / 2);
    avgpool2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, out, B * C, D0, D1, D2);
    return launch_status();
}

extern "C" int pulpo_avgpool2_pyramid_fwd(const float *x, float *const *outs, int nlevels, int B, int C, int D0, int D1,
                                          int D2, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_avgpool2_pyramid_fwd");
    PULPO_REQUIRE(x && outs, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 > 0 && D1 > 0 && D2 > 0, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(nlevels >= 1 && nlevels <= POOL_MAXL, PULPO_ERR_INVALID_SHAPE);
    const int nl = nlevels, E = nl < 2 ? 4 : (1 << nl);   // a single level still stages 4^3 blocks (quads)
    PULPO_REQUIRE(D0 % E == 0 && D1 % E == 0 && D2 % E == 0 && aligned16(x), PULPO_ERR_UNSUPPORTED);
    PoolPyr p;
    p.nl = nl; p.E = E; p.BC = B * C; p.D0 = D0; p.D1 = D1; p.D2 = D2;
    p.nb0 = D0 / E; p.nb1 = D1 / E; p.nb2 = D2 / E;
    for (int l = 0; l < POOL_MAXL; ++l) p.out[l] = nullptr;
    for (int l = 0; l < nl; ++l) {
        PULPO_REQUIRE(outs[l], PULPO_ERR_NULL_POINTER);
        p.out[l] = outs[l];
    }
    const i64 blocks = (i64)p.BC * p.nb0 * p.nb1 * p.nb2;
    PULPO_REQUIRE(blocks < (1ll << 31), PULPO_ERR_INVALID_SHAPE);
    avgpool2_pyramid_kernel<<<(unsigned int)blocks, 256, 0, (cudaStream_t)stream>>>(x, p);
    return launch_status();
}
