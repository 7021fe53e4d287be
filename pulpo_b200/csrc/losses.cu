// losses.cu -- streaming scalar losses and MC moments:
//   * KL_two_gauss_with_diag_cov       src/losses.py:47-76   (call order :271-273)
//   * L2_reg (3-D branch)              src/losses.py:208-222
//   * per-voxel Welford/Chan moments   evaluate.py:243-251 (std over MC samples)
// Pure HBM streaming: 128-bit loads where alignment allows, warp-shuffle -> CTA -> deterministic
// two-stage grid reduction in double for the scalars.
#include "common.cuh"

namespace pulpo {

__device__ __forceinline__ float kl_term(float m0, float s0, float m1, float s1, float eps)
{
    const float v0 = __fmul_rn(s0, s0), v1 = __fmul_rn(s1, s1);
    const float dm = __fsub_rn(m1, m0);
    const float num = __fadd_rn(v0, __fmul_rn(dm, dm));
    const float den = __fadd_rn(v1, eps);
    float t = __fdiv_rn(num, den);
    t = __fadd_rn(t, logf(den));
    t = __fsub_rn(t, logf(__fadd_rn(v0, eps)));
    return __fsub_rn(t, 1.0f);
}

__global__ void __launch_bounds__(256)
kl_fwd_kernel(const float *__restrict__ mu0, const float *__restrict__ sg0, const float *__restrict__ mu1,
              const float *__restrict__ sg1, float eps, float *out, ReduceWs *ws, double scale, i64 total, int vec)
{
    __shared__ double red[32];
    float acc = 0.0f;
    const i64 tid = blockIdx.x * (i64)blockDim.x + threadIdx.x, nthr = (i64)gridDim.x * blockDim.x;
    if (vec) {
        for (i64 i = tid * 4; i < total; i += nthr * 4) {
            float4 m = ld_stream4(mu0 + i), s = ld_stream4(sg0 + i);
            float4 m1 = mu1 ? ld_stream4(mu1 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 s1 = sg1 ? ld_stream4(sg1 + i) : make_float4(1.f, 1.f, 1.f, 1.f);
            float t = kl_term(m.x, s.x, m1.x, s1.x, eps);
            t += kl_term(m.y, s.y, m1.y, s1.y, eps);
            t += kl_term(m.z, s.z, m1.z, s1.z, eps);
            t += kl_term(m.w, s.w, m1.w, s1.w, eps);
            acc += t;
        }
    } else {
        for (i64 i = tid; i < total; i += nthr)
            acc += kl_term(mu0[i], sg0[i], mu1 ? mu1[i] : 0.0f, sg1 ? sg1[i] : 1.0f, eps);
    }
    double bt = block_sum((double)acc, red);
    grid_reduce_finish(bt, ws, out, scale, red);
}

__global__ void __launch_bounds__(256)
kl_bwd_kernel(const float *__restrict__ gloss, const float *__restrict__ mu0, const float *__restrict__ sg0,
              const float *__restrict__ mu1, const float *__restrict__ sg1, float eps, float *__restrict__ gmu,
              float *__restrict__ gsg, float invB, i64 total)
{
    const float k = (gloss ? __ldg(gloss) : 1.0f) * invB;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const float s = sg0[i];
        const float s1 = sg1 ? sg1[i] : 1.0f;
        const float m1 = mu1 ? mu1[i] : 0.0f;
        const float den = s1 * s1 + eps;
        gmu[i] = k * (mu0[i] - m1) / den;
        gsg[i] = k * (s / den - s / (s * s + eps));
    }
}

// ---- f-4: gauss_sampler (src/network_blocks.py:7-8) fused with the level's KL term (src/losses.py:47-76,
// call order :271-273): z = mu + sigma * (var * noise) and KL[N(mu, sigma) || p1] from ONE read of mu, sigma.
// The noise is an input (torch's Philox stream stays the source of randomness, so the samples are the
// reference's for the same generator state).
__global__ void __launch_bounds__(256)
gauss_kl_fwd_kernel(const float *__restrict__ mu, const float *__restrict__ sg, const float *__restrict__ noise,
                    const float *__restrict__ mu1, const float *__restrict__ sg1, float var, float eps,
                    float *__restrict__ z, float *out, ReduceWs *ws, double scale, i64 total, int vec)
{
    __shared__ double red[32];
    float acc = 0.0f;
    const i64 tid = blockIdx.x * (i64)blockDim.x + threadIdx.x, nthr = (i64)gridDim.x * blockDim.x;
    if (vec) {
        for (i64 i = tid * 4; i < total; i += nthr * 4) {
            const float4 m = ld_stream4(mu + i), s = ld_stream4(sg + i), e = ld_stream4(noise + i);
            const float4 m1 = mu1 ? ld_stream4(mu1 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 s1 = sg1 ? ld_stream4(sg1 + i) : make_float4(1.f, 1.f, 1.f, 1.f);
            float4 o;
            o.x = __fadd_rn(m.x, __fmul_rn(s.x, __fmul_rn(var, e.x)));
            o.y = __fadd_rn(m.y, __fmul_rn(s.y, __fmul_rn(var, e.y)));
            o.z = __fadd_rn(m.z, __fmul_rn(s.z, __fmul_rn(var, e.z)));
            o.w = __fadd_rn(m.w, __fmul_rn(s.w, __fmul_rn(var, e.w)));
            *reinterpret_cast<float4 *>(z + i) = o;
            float t = kl_term(m.x, s.x, m1.x, s1.x, eps);
            t += kl_term(m.y, s.y, m1.y, s1.y, eps);
            t += kl_term(m.z, s.z, m1.z, s1.z, eps);
            t += kl_term(m.w, s.w, m1.w, s1.w, eps);
            acc += t;
        }
    } else {
        for (i64 i = tid; i < total; i += nthr) {
            const float m = mu[i], s = sg[i];
            z[i] = __fadd_rn(m, __fmul_rn(s, __fmul_rn(var, noise[i])));
            acc += kl_term(m, s, mu1 ? mu1[i] : 0.0f, sg1 ? sg1[i] : 1.0f, eps);
        }
    }
    double bt = block_sum((double)acc, red);
    grid_reduce_finish(bt, ws, out, scale, red);
}

// gmu = gz + gloss * dKL/dmu,  gsigma = gz * var * noise + gloss * dKL/dsigma   (gz, gloss nullable)
__global__ void __launch_bounds__(256)
gauss_kl_bwd_kernel(const float *__restrict__ gz, const float *__restrict__ gloss, const float *__restrict__ mu,
                    const float *__restrict__ sg, const float *__restrict__ noise, const float *__restrict__ mu1,
                    const float *__restrict__ sg1, float var, float eps, float *__restrict__ gmu,
                    float *__restrict__ gsg, float k0, int have_kl, i64 total)
{
    const float k = have_kl ? (gloss ? __ldg(gloss) : 1.0f) * k0 : 0.0f;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const float s = sg[i];
        const float s1 = sg1 ? sg1[i] : 1.0f;
        const float m1 = mu1 ? mu1[i] : 0.0f;
        const float den = s1 * s1 + eps;
        const float g = gz ? gz[i] : 0.0f;
        gmu[i] = g + k * (mu[i] - m1) / den;
        gsg[i] = g * (var * noise[i]) + k * (s / den - s / (s * s + eps));
    }
}

// ---- all pyramid levels, value and gradients, in ONE launch (HotPathPlan: the KL terms have no data
// dependence on anything else in the step, so eight tiny launches become one streaming pass that reads
// mu / sigma once).  N(0,1) prior only (src/components/pulpo.py:337-339).
constexpr int KL_MAXL = 6;
struct KlLevel {
    const float *mu, *sg;
    float *gmu, *gsg, *out;
    i64 total;      // B * n
    int vec;        // total % 4 == 0 and all four pointers 16-byte aligned
    float k;        // weight / B : scale of the gradients
    double scale;   // 0.5 * weight / B : scale of the value
};
struct KlMulti {
    int n;
    KlLevel l[KL_MAXL];
};
struct KlMultiWs {
    unsigned int ticket, pad;
    double partial[1];   // [KL_MAXL][ctas]
};

__global__ void __launch_bounds__(256)
kl_multi_kernel(const KlMulti m, float eps, KlMultiWs *ws)
{
    __shared__ double red[32];
    __shared__ bool is_last;
    const i64 tid = blockIdx.x * (i64)blockDim.x + threadIdx.x, nthr = (i64)gridDim.x * blockDim.x;
    const float den = 1.0f + eps;
    for (int lv = 0; lv < m.n; ++lv) {
        const KlLevel &L = m.l[lv];
        float acc = 0.0f;
        if (L.vec) {   // 128-bit accesses, two quads in flight per thread (the scalar loop is latency-bound)
            const i64 nq = L.total >> 2;
            for (i64 q = tid; q < nq; q += 2 * nthr) {
                const i64 q1 = q + nthr;
                const bool two = q1 < nq;
                const float4 m0 = ld_stream4(L.mu + 4 * q), s0 = ld_stream4(L.sg + 4 * q);
                const float4 m1 = two ? ld_stream4(L.mu + 4 * q1) : m0, s1 = two ? ld_stream4(L.sg + 4 * q1) : s0;
                float4 gm, gs;
                float t = kl_term(m0.x, s0.x, 0.0f, 1.0f, eps);
                t += kl_term(m0.y, s0.y, 0.0f, 1.0f, eps);
                t += kl_term(m0.z, s0.z, 0.0f, 1.0f, eps);
                t += kl_term(m0.w, s0.w, 0.0f, 1.0f, eps);
                acc += t;
                gm.x = L.k * m0.x / den; gm.y = L.k * m0.y / den; gm.z = L.k * m0.z / den; gm.w = L.k * m0.w / den;
                gs.x = L.k * (s0.x / den - s0.x / (s0.x * s0.x + eps)); gs.y = L.k * (s0.y / den - s0.y / (s0.y * s0.y + eps));
                gs.z = L.k * (s0.z / den - s0.z / (s0.z * s0.z + eps)); gs.w = L.k * (s0.w / den - s0.w / (s0.w * s0.w + eps));
                *reinterpret_cast<float4 *>(L.gmu + 4 * q) = gm;
                *reinterpret_cast<float4 *>(L.gsg + 4 * q) = gs;
                if (two) {
                    t = kl_term(m1.x, s1.x, 0.0f, 1.0f, eps);
                    t += kl_term(m1.y, s1.y, 0.0f, 1.0f, eps);
                    t += kl_term(m1.z, s1.z, 0.0f, 1.0f, eps);
                    t += kl_term(m1.w, s1.w, 0.0f, 1.0f, eps);
                    acc += t;
                    gm.x = L.k * m1.x / den; gm.y = L.k * m1.y / den; gm.z = L.k * m1.z / den; gm.w = L.k * m1.w / den;
                    gs.x = L.k * (s1.x / den - s1.x / (s1.x * s1.x + eps)); gs.y = L.k * (s1.y / den - s1.y / (s1.y * s1.y + eps));
                    gs.z = L.k * (s1.z / den - s1.z / (s1.z * s1.z + eps)); gs.w = L.k * (s1.w / den - s1.w / (s1.w * s1.w + eps));
                    *reinterpret_cast<float4 *>(L.gmu + 4 * q1) = gm;
                    *reinterpret_cast<float4 *>(L.gsg + 4 * q1) = gs;
                }
            }
        } else {
            for (i64 i = tid; i < L.total; i += nthr) {
                const float mu = __ldg(L.mu + i), s = __ldg(L.sg + i);
                acc += kl_term(mu, s, 0.0f, 1.0f, eps);
                L.gmu[i] = L.k * mu / den;
                L.gsg[i] = L.k * (s / den - s / (s * s + eps));
            }
        }
        const double bt = block_sum((double)acc, red);
        if (threadIdx.x == 0) ws->partial[(i64)lv * gridDim.x + blockIdx.x] = bt;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int lv = 0; lv < m.n; ++lv) {
            double sum = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x)
                sum += ((volatile double *)ws->partial)[(i64)lv * gridDim.x + i];
            sum = block_sum(sum, red);
            if (threadIdx.x == 0) *m.l[lv].out = (float)(sum * m.l[lv].scale);
        }
        if (threadIdx.x == 0) ws->ticket = 0;
    }
}

// L2_reg: forward differences on the [1:,1:,1:] crop
__global__ void __launch_bounds__(256)
l2reg_fwd_kernel(const float *__restrict__ f, float *out, ReduceWs *ws, double scale, int BC, int D0, int D1, int D2)
{
    __shared__ double red[32];
    const i64 sy = D2, sz = (i64)D1 * D2, S = (i64)D0 * sz, total = (i64)BC * S;
    float acc = 0.0f;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < total; g += (i64)gridDim.x * blockDim.x) {
        int x = (int)(g % D2);
        i64 r = g / D2;
        int y = (int)(r % D1);
        int z = (int)((r / D1) % D0);
        if (x == 0 || y == 0 || z == 0) continue;
        const float c = __ldg(f + g);
        const float a = c - __ldg(f + g - sz), b = c - __ldg(f + g - sy), d = c - __ldg(f + g - 1);
        acc += a * a + b * b + d * d;
    }
    double bt = block_sum((double)acc, red);
    grid_reduce_finish(bt, ws, out, scale, red);
}

// gradient in gather form: voxel v collects its own three differences (if v is in the crop)
// minus the difference of each forward neighbour that is in the crop
__global__ void __launch_bounds__(256)
l2reg_bwd_kernel(const float *__restrict__ gloss, const float *__restrict__ f, float *__restrict__ gf, float kk,
                 int BC, int D0, int D1, int D2, int accumulate)
{
    const i64 sy = D2, sz = (i64)D1 * D2, S = (i64)D0 * sz, total = (i64)BC * S;
    const float k = (gloss ? __ldg(gloss) : 1.0f) * kk;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < total; g += (i64)gridDim.x * blockDim.x) {
        int x = (int)(g % D2);
        i64 r = g / D2;
        int y = (int)(r % D1);
        int z = (int)((r / D1) % D0);
        const float c = __ldg(f + g);
        float acc = 0.0f;
        if (x > 0 && y > 0 && z > 0)
            acc += (c - __ldg(f + g - sz)) + (c - __ldg(f + g - sy)) + (c - __ldg(f + g - 1));
        if (z + 1 < D0 && y > 0 && x > 0) acc -= __ldg(f + g + sz) - c;
        if (y + 1 < D1 && z > 0 && x > 0) acc -= __ldg(f + g + sy) - c;
        if (x + 1 < D2 && z > 0 && y > 0) acc -= __ldg(f + g + 1) - c;
        gf[g] = accumulate ? gf[g] + k * acc : k * acc;
    }
}

// 128-bit variants (D2 % 4 == 0): one thread owns 4 consecutive voxels of a row, FastDiv decode
struct RowGeom {
    int BC, D0, D1, D2, XG;
    unsigned int groups;
    FastDiv dXG, dD1, dD0;
};

static bool make_rowgeom(RowGeom &g, int BC, int D0, int D1, int D2)
{
    if (D2 % 4 || (i64)BC * D0 * D1 * D2 >= (1ll << 31)) return false;
    g.BC = BC; g.D0 = D0; g.D1 = D1; g.D2 = D2; g.XG = D2 / 4;
    g.groups = (unsigned int)((i64)BC * D0 * D1 * g.XG);
    g.dXG = make_fastdiv(g.XG); g.dD1 = make_fastdiv(D1); g.dD0 = make_fastdiv(D0);
    return true;
}

__global__ void __launch_bounds__(256)
l2reg_fwd_v4_kernel(const float *__restrict__ f, float *out, ReduceWs *ws, double scale, const RowGeom g)
{
    __shared__ double red[32];
    const int sy = g.D2, sz = g.D1 * g.D2;
    float acc = 0.0f;
    for (unsigned int gid = blockIdx.x * 256u + threadIdx.x; gid < g.groups; gid += gridDim.x * 256u) {
        unsigned int row, xg, zb, y, bc, z;
        fast_divmod(gid, g.dXG, row, xg);
        fast_divmod(row, g.dD1, zb, y);
        fast_divmod(zb, g.dD0, bc, z);
        if (y == 0 || z == 0) continue;
        const float *p = f + (i64)gid * 4;
        const float4 c = ld_stream4(p), pz = ld_stream4(p - sz), py = ld_stream4(p - sy);
        const float px = xg ? __ldg(p - 1) : c.x;   // x == 0 is outside the crop
        float t;
        if (xg) { t = c.x - pz.x; acc += t * t; t = c.x - py.x; acc += t * t; t = c.x - px; acc += t * t; }
        t = c.y - pz.y; acc += t * t; t = c.y - py.y; acc += t * t; t = c.y - c.x; acc += t * t;
        t = c.z - pz.z; acc += t * t; t = c.z - py.z; acc += t * t; t = c.z - c.y; acc += t * t;
        t = c.w - pz.w; acc += t * t; t = c.w - py.w; acc += t * t; t = c.w - c.z; acc += t * t;
    }
    double bt = block_sum((double)acc, red);
    grid_reduce_finish(bt, ws, out, scale, red);
}

// ACC: gf += ...  (lets the caller fold this gradient into an existing one without an extra pass)
template <bool ACC>
__global__ void __launch_bounds__(256)
l2reg_bwd_v4_kernel(const float *__restrict__ gloss, const float *__restrict__ f, float *__restrict__ gf, float kk,
                    const RowGeom g)
{
    const unsigned int gid = blockIdx.x * 256u + threadIdx.x;
    if (gid >= g.groups) return;
    const int sy = g.D2, sz = g.D1 * g.D2;
    const float k = (gloss ? __ldg(gloss) : 1.0f) * kk;
    unsigned int row, xg, zb, y, bc, z;
    fast_divmod(gid, g.dXG, row, xg);
    fast_divmod(row, g.dD1, zb, y);
    fast_divmod(zb, g.dD0, bc, z);
    const float *p = f + (i64)gid * 4;
    const float4 c4 = ld_stream4(p);
    const float c[6] = {xg ? __ldg(p - 1) : 0.0f, c4.x, c4.y, c4.z, c4.w, (int)xg + 1 < g.XG ? __ldg(p + 4) : 0.0f};
    const bool zin = z > 0, yin = y > 0, zn = (int)z + 1 < g.D0, yn = (int)y + 1 < g.D1;
    float4 pz = make_float4(0.f, 0.f, 0.f, 0.f), py = pz, nz = pz, ny = pz;
    if (zin) pz = ld_stream4(p - sz);
    if (yin) py = ld_stream4(p - sy);
    if (zn) nz = ld_stream4(p + sz);
    if (yn) ny = ld_stream4(p + sy);
    const float pzv[4] = {pz.x, pz.y, pz.z, pz.w}, pyv[4] = {py.x, py.y, py.z, py.w};
    const float nzv[4] = {nz.x, nz.y, nz.z, nz.w}, nyv[4] = {ny.x, ny.y, ny.z, ny.w};
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = 4 * (int)xg + j;
        const bool xin = x > 0, xn = x + 1 < g.D2;
        const float cc = c[j + 1];
        float a = 0.0f;
        if (xin && yin && zin) a += (cc - pzv[j]) + (cc - pyv[j]) + (cc - c[j]);
        if (zn && yin && xin) a -= nzv[j] - cc;
        if (yn && zin && xin) a -= nyv[j] - cc;
        if (xn && zin && yin) a -= c[j + 2] - cc;
        r[j] = k * a;
    }
    float4 *o = reinterpret_cast<float4 *>(gf + (i64)gid * 4);
    if (ACC) {
        const float4 old = *o;
        r[0] += old.x; r[1] += old.y; r[2] += old.z; r[3] += old.w;
    }
    *o = make_float4(r[0], r[1], r[2], r[3]);
}

// Value AND gradient of L2_reg in one pass over the field (the hot-path plan's form: the loss weights are folded into
// the kernels, so the gradient does not wait for an upstream scalar).  Same per-voxel arithmetic as the two kernels
// above; every thread owns 4 consecutive voxels of a row, the five neighbour quads come through L2 (read-once
// streaming loads), persistent grid so that the value can use the deterministic two-stage reduction.
// PROD: gf additionally receives gout * dpos -- the gather half of the warp's backward when its forward stored dpos
// (warp3d.cu, pulpo_warp3d_fwd_dpos); gout has one channel per batch item, f / dpos / gf have C.
template <bool ACC, bool PROD>
__global__ void __launch_bounds__(256)
l2reg_fwd_bwd_v4_kernel(const float *__restrict__ f, float *__restrict__ gf, float kk, float *out, ReduceWs *ws,
                        double scale, const float *__restrict__ gout, const float *__restrict__ dpos, int C,
                        const RowGeom g)
{
    __shared__ double red[32];
    const int sy = g.D2, sz = g.D1 * g.D2;
    float vacc = 0.0f;
    for (unsigned int gid = blockIdx.x * 256u + threadIdx.x; gid < g.groups; gid += gridDim.x * 256u) {
        unsigned int row, xg, zb, y, bc, z;
        fast_divmod(gid, g.dXG, row, xg);
        fast_divmod(row, g.dD1, zb, y);
        fast_divmod(zb, g.dD0, bc, z);
        const float *p = f + (i64)gid * 4;
        const float4 c4 = ld_stream4(p);
        const bool zin = z > 0, yin = y > 0, zn = (int)z + 1 < g.D0, yn = (int)y + 1 < g.D1;
        float4 pz = make_float4(0.f, 0.f, 0.f, 0.f), py = pz, nz = pz, ny = pz, go = pz, dp = pz;
        if (PROD) {
            const unsigned int b = bc / (unsigned int)C;
            const i64 S4 = (i64)g.D0 * g.D1 * g.XG;                 // quads per channel
            go = ld_stream4(gout + ((i64)gid - (i64)(bc - b) * S4) * 4);
            dp = ld_stream4(dpos + (i64)gid * 4);
        }
        if (zin) pz = ld_stream4(p - sz);
        if (yin) py = ld_stream4(p - sy);
        if (zn) nz = ld_stream4(p + sz);
        if (yn) ny = ld_stream4(p + sy);
        const float c[6] = {xg ? __ldg(p - 1) : 0.0f, c4.x, c4.y, c4.z, c4.w, (int)xg + 1 < g.XG ? __ldg(p + 4) : 0.0f};
        const float pzv[4] = {pz.x, pz.y, pz.z, pz.w}, pyv[4] = {py.x, py.y, py.z, py.w};
        const float nzv[4] = {nz.x, nz.y, nz.z, nz.w}, nyv[4] = {ny.x, ny.y, ny.z, ny.w};
        float r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = 4 * (int)xg + j;
            const bool xin = x > 0, xn = x + 1 < g.D2;
            const float cc = c[j + 1];
            float a = 0.0f;
            if (xin && yin && zin) {
                const float dz = cc - pzv[j], dy = cc - pyv[j], dx = cc - c[j];
                a += dz + dy + dx;
                vacc += dz * dz; vacc += dy * dy; vacc += dx * dx;
            }
            if (zn && yin && xin) a -= nzv[j] - cc;
            if (yn && zin && xin) a -= nyv[j] - cc;
            if (xn && zin && yin) a -= c[j + 2] - cc;
            r[j] = kk * a;
        }
        if (PROD) {
            r[0] += go.x * dp.x; r[1] += go.y * dp.y; r[2] += go.z * dp.z; r[3] += go.w * dp.w;
        }
        float4 *o = reinterpret_cast<float4 *>(gf + (i64)gid * 4);
        if (ACC) {
            const float4 old = *o;
            r[0] += old.x; r[1] += old.y; r[2] += old.z; r[3] += old.w;
        }
        *o = make_float4(r[0], r[1], r[2], r[3]);
    }
    double bt = block_sum((double)vacc, red);
    grid_reduce_finish(bt, ws, out, scale, red);
}

// z-marching form of the kernel above (the default when the volume has >= 8 planes).  The one-plane-per-thread version
// pulls five neighbour quads of the field through L2 for every quad it produces (~660 MB of L2 traffic for 275 MB of
// HBM bytes at 160x192x224: L2-bound).  Here a thread owns one quad column and walks a run of planes: the quads of the
// previous and the next plane are carried in registers, the y neighbours of a plane come from L1 (the rows above and
// below belong to threads of the same CTA at the same plane), and everything plane z + 1 needs is requested before
// plane z is computed.
struct L2MarchGeom {
    int BC, D0, D1, D2, XG, zrun, nzrun, C;
    unsigned int items;     // BC * nzrun * D1 * XG
    FastDiv dXG, dD1, dnz, dC;
};

struct L2Plane {
    float4 c, py, ny, go, dp, old;
    float xl, xr;
};

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

template <bool ACC, bool PROD>
__device__ __forceinline__ void l2_plane_load(L2Plane &p, const float *fq, const float *goq, const float *dpq,
                                              const float *gfq, int sy, bool full, bool yin, bool yn, bool xl_ok, bool xr_ok)
{
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    p.c = ldg4(fq);
    p.py = (full && yin) ? ldg4(fq - sy) : z4;
    p.ny = (full && yn) ? ldg4(fq + sy) : z4;
    p.xl = (full && xl_ok) ? __ldg(fq - 1) : 0.0f;
    p.xr = (full && xr_ok) ? __ldg(fq + 4) : 0.0f;
    if (PROD) {
        p.go = full ? ld_stream4(goq) : z4;
        p.dp = full ? ld_stream4(dpq) : z4;
    }
    if (ACC) p.old = full ? ld_stream4(gfq) : z4;
}

#ifndef PULPO_L2M_CTAS
#define PULPO_L2M_CTAS 3
#endif
#ifndef PULPO_L2M_PF
#define PULPO_L2M_PF 1      // planes ahead of the register prefetch that are requested into L2 (0: off)
#endif
template <bool ACC, bool PROD>
__global__ void __launch_bounds__(256, PULPO_L2M_CTAS)
l2reg_fwd_bwd_march_kernel(const float *__restrict__ f, float *__restrict__ gf, float kk, float *out, ReduceWs *ws,
                           double scale, const float *__restrict__ gout, const float *__restrict__ dpos,
                           const L2MarchGeom g)
{
    __shared__ double red[32];
    const int sy = g.D2;
    const i64 sz = (i64)g.D1 * g.D2, S = (i64)g.D0 * sz;
    float vacc = 0.0f;
    for (unsigned int it = blockIdx.x * 256u + threadIdx.x; it < g.items; it += gridDim.x * 256u) {
        // item order: x quad, row, CHANNEL, z run, batch item -- the C channels of a (z run, batch item) are in flight
        // together, so the one-channel gout they share is read from HBM once and from L2 afterwards
        unsigned int row, xg, r2, y, r3, ch, b, zr;
        fast_divmod(it, g.dXG, row, xg);
        fast_divmod(row, g.dD1, r2, y);
        fast_divmod(r2, g.dC, r3, ch);
        fast_divmod(r3, g.dnz, b, zr);
        const unsigned int bc = b * (unsigned int)g.C + ch;
        const int z0 = (int)zr * g.zrun, z1 = min(g.D0, z0 + g.zrun);
        const bool yin = y > 0, yn = (int)y + 1 < g.D1, xl_ok = xg > 0, xr_ok = (int)xg + 1 < g.XG;
        const i64 q0 = (i64)bc * S + (i64)z0 * sz + (i64)y * g.D2 + 4 * (i64)xg;     // this thread's quad at plane z0
        const i64 qg = q0 - (i64)(bc - b) * S;                                       // the same quad of gout [B,1,...]
        const float *fq = f + q0, *goq = PROD ? gout + qg : nullptr, *dpq = PROD ? dpos + q0 : nullptr;
        float *gfq = gf + q0;
        float4 prev = z0 > 0 ? ldg4(fq - sz) : make_float4(0.f, 0.f, 0.f, 0.f);
        L2Plane cur;
        l2_plane_load<ACC, PROD>(cur, fq, goq, dpq, gfq, sy, true, yin, yn, xl_ok, xr_ok);
        for (int z = z0; z < z1; ++z) {
            L2Plane nxt;
            const bool zn = z + 1 < g.D0, zin = z > 0;
            if (PULPO_L2M_PF > 0 && z + 1 + PULPO_L2M_PF < z1) {
                prefetch_l2(fq + (1 + PULPO_L2M_PF) * sz);
                if (PROD) { prefetch_l2(goq + (1 + PULPO_L2M_PF) * sz); prefetch_l2(dpq + (1 + PULPO_L2M_PF) * sz); }
            }
            if (zn)
                l2_plane_load<ACC, PROD>(nxt, fq + sz, PROD ? goq + sz : nullptr, PROD ? dpq + sz : nullptr, gfq + sz, sy,
                                         z + 1 < z1, yin, yn, xl_ok, xr_ok);
            else
                nxt.c = make_float4(0.f, 0.f, 0.f, 0.f);
            const float c[6] = {cur.xl, cur.c.x, cur.c.y, cur.c.z, cur.c.w, cur.xr};
            const float pzv[4] = {prev.x, prev.y, prev.z, prev.w}, pyv[4] = {cur.py.x, cur.py.y, cur.py.z, cur.py.w};
            const float nzv[4] = {nxt.c.x, nxt.c.y, nxt.c.z, nxt.c.w}, nyv[4] = {cur.ny.x, cur.ny.y, cur.ny.z, cur.ny.w};
            float r[4];
            if (zin && zn && yin && yn) {
                // interior plane and row (almost every iteration): only the two x faces need masks -- no predicates
                const float mi0 = xl_ok ? 1.0f : 0.0f, mn3 = xr_ok ? 1.0f : 0.0f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float cc = c[j + 1];
                    const float dz = cc - pzv[j], dy = cc - pyv[j], dx = cc - c[j];
                    const float own = (dz + dy + dx) - (nzv[j] - cc) - (nyv[j] - cc);
                    const float sq = dz * dz + dy * dy + dx * dx;
                    const float fwd = c[j + 2] - cc;
                    if (j == 0) {          // x == 0 (first quad of the row) lies outside the crop
                        vacc += mi0 * sq;
                        r[j] = kk * (mi0 * own - fwd);
                    } else if (j == 3) {   // x == D2 - 1 (last quad) has no forward x neighbour
                        vacc += sq;
                        r[j] = kk * (own - mn3 * fwd);
                    } else {
                        vacc += sq;
                        r[j] = kk * (own - fwd);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int x = 4 * (int)xg + j;
                    const bool xin = x > 0, xn = x + 1 < g.D2;
                    const float cc = c[j + 1];
                    float a = 0.0f;
                    if (xin && yin && zin) {
                        const float dz = cc - pzv[j], dy = cc - pyv[j], dx = cc - c[j];
                        a += dz + dy + dx;
                        vacc += dz * dz; vacc += dy * dy; vacc += dx * dx;
                    }
                    if (zn && yin && xin) a -= nzv[j] - cc;
                    if (yn && zin && xin) a -= nyv[j] - cc;
                    if (xn && zin && yin) a -= c[j + 2] - cc;
                    r[j] = kk * a;
                }
            }
            if (PROD) {
                r[0] += cur.go.x * cur.dp.x; r[1] += cur.go.y * cur.dp.y; r[2] += cur.go.z * cur.dp.z; r[3] += cur.go.w * cur.dp.w;
            }
            if (ACC) { r[0] += cur.old.x; r[1] += cur.old.y; r[2] += cur.old.z; r[3] += cur.old.w; }
            *reinterpret_cast<float4 *>(gfq) = make_float4(r[0], r[1], r[2], r[3]);
            prev = cur.c;
            cur = nxt;
            fq += sz; gfq += sz;
            if (PROD) { goq += sz; dpq += sz; }
        }
    }
    double bt = block_sum((double)vacc, red);
    grid_reduce_finish(bt, ws, out, scale, red);
}

// ---------------------------------------------------------------------------------------------
// L2_reg of a x2 up-sampled field, evaluated on the COARSE grid (value and gradient w.r.t. the coarse field).
// The hot path regularises final = ResizeTransform(1/2)(integrated) = up2(2 * v) (src/components/pulpo.py:314,
// src/models.py:162): a trilinear interpolant of v.  Every forward difference of the fine field along axis a is a fixed
// combination of the coarse differences e_a = v[. + 1_a] - v[.]:   d[2j+1] = (e[j-1] + e[j]) / 2,  d[2j+2] = e[j]
// (e = 0 outside the volume; ATen's edge clamp gives exactly this at the faces), and along the other two axes the fine
// values are the 0.25 / 0.75 blends.  Summing the squares over the fine [1:,1:,1:] crop gives
//     sum (d_a final)^2 = < e_a , (K_a x B_b x B_c) e_a >,   K = tridiag(.25, 1.5, .25) on e,
//                                                            B = tridiag(.375, 1.25, .375), B[0][0] = .625, B[n-1][n-1] = 1.625
// (B = U^T C U with U the 1-D interpolation and C the crop mask), so
//     L2_reg(final) = scale * sum_a < e_a, q_a >,  q_a = (K x B x B) e_a,     d L2_reg / d v = 2 scale sum_a E_a^T q_a.
// All coefficients are positive and act on differences, so the conditioning is that of the reference's own
// difference-of-interpolated-values (checked against it in fp64: identical to 1e-16).  The full-resolution pass over
// the final field (82.6 MB read + its gradient written and read again) becomes a 10 MB stencil on L2-resident data.
// Separable evaluation (three 1-D passes over L2-resident coarse arrays, x then y then z):
//     x pass:  X1 = B_x v,            Qx1 = K_x E_x v
//     y pass:  Y1 = B_y X1,           Qy1 = K_y E_y X1,        Qx2 = B_y Qx1
//     z pass:  q_z = K_z E_z Y1,      q_y = B_z Qy1,           q_x = B_z Qx2;
//              value += e_z q_z + e_y q_y + e_x q_x,   gv (+)= 2 scale sum_a (q_a[. - 1_a] - q_a[.])
// (E = forward difference, zero beyond the last voxel).  ~35 cached loads per coarse voxel in total; the brute-force
// 27-tap form needed 135 and ran 10x slower.
struct RegUpGeom {
    int BC, d0, d1, d2;
    unsigned int total;      // BC * d0 * d1 * d2
    FastDiv dd2, dd1, dd0;
};

// (B Z)[j] along an axis of size n and element stride `st`; p points at Z[j]
__device__ __forceinline__ float regup_B(const float *__restrict__ p, int j, int n, int st)
{
    const float c = j == 0 ? (n == 1 ? 1.0f : 0.625f) : (j == n - 1 ? 1.625f : 1.25f);
    float r = c * __ldg(p);
    if (j > 0) r += 0.375f * __ldg(p - st);
    if (j + 1 < n) r += 0.375f * __ldg(p + st);
    return r;
}
// (K E Z)[j]: .25 t[j-1] + 1.5 t[j] + .25 t[j+1],  t[i] = Z[i+1] - Z[i] for 0 <= i <= n-2, zero otherwise
__device__ __forceinline__ float regup_KE(const float *__restrict__ p, int j, int n, int st)
{
    if (j < 0 || j + 1 >= n) return 0.0f;
    const float z0 = __ldg(p), z1 = __ldg(p + st);
    float r = 1.5f * (z1 - z0);
    if (j > 0) r += 0.25f * (z0 - __ldg(p - st));
    if (j + 2 < n) r += 0.25f * (__ldg(p + 2 * st) - z1);
    return r;
}

__global__ void __launch_bounds__(256)
regup_x_kernel(const float *__restrict__ v, float *__restrict__ X1, float *__restrict__ Qx1, const RegUpGeom g)
{
    for (unsigned int i = blockIdx.x * 256u + threadIdx.x; i < g.total; i += gridDim.x * 256u) {
        unsigned int r, x;
        fast_divmod(i, g.dd2, r, x);
        X1[i] = regup_B(v + i, (int)x, g.d2, 1);
        Qx1[i] = regup_KE(v + i, (int)x, g.d2, 1);
    }
}

__global__ void __launch_bounds__(256)
regup_y_kernel(const float *__restrict__ X1, const float *__restrict__ Qx1, float *__restrict__ Y1, float *__restrict__ Qy1,
               float *__restrict__ Qx2, const RegUpGeom g)
{
    for (unsigned int i = blockIdx.x * 256u + threadIdx.x; i < g.total; i += gridDim.x * 256u) {
        unsigned int r, x, r2, y;
        fast_divmod(i, g.dd2, r, x);
        fast_divmod(r, g.dd1, r2, y);
        Y1[i] = regup_B(X1 + i, (int)y, g.d1, g.d2);
        Qy1[i] = regup_KE(X1 + i, (int)y, g.d1, g.d2);
        Qx2[i] = regup_B(Qx1 + i, (int)y, g.d1, g.d2);
    }
}

template <bool ACC>
__global__ void __launch_bounds__(256)
regup_z_kernel(const float *__restrict__ v, const float *__restrict__ Y1, const float *__restrict__ Qy1,
               const float *__restrict__ Qx2, float *__restrict__ gv, float k2, float *out, ReduceWs *ws, double scale,
               const RegUpGeom g)
{
    __shared__ double red[32];
    const int sy = g.d2, sz = g.d1 * g.d2;
    float vacc = 0.0f;
    for (unsigned int i = blockIdx.x * 256u + threadIdx.x; i < g.total; i += gridDim.x * 256u) {
        unsigned int r, ux, r2, uy, bc, uz;
        fast_divmod(i, g.dd2, r, ux);
        fast_divmod(r, g.dd1, r2, uy);
        fast_divmod(r2, g.dd0, bc, uz);
        const int x = (int)ux, y = (int)uy, z = (int)uz;
        const float c = __ldg(v + i);
        // term z
        const float qz = regup_KE(Y1 + i, z, g.d0, sz), qzm = regup_KE(Y1 + i - sz, z - 1, g.d0, sz);
        // term y: q_y at rows y and y - 1 (Qy1 is zero in the last row)
        const float qy = regup_B(Qy1 + i, z, g.d0, sz), qym = y > 0 ? regup_B(Qy1 + i - sy, z, g.d0, sz) : 0.0f;
        // term x
        const float qx = regup_B(Qx2 + i, z, g.d0, sz), qxm = x > 0 ? regup_B(Qx2 + i - 1, z, g.d0, sz) : 0.0f;
        const float ez = z + 1 < g.d0 ? __ldg(v + i + sz) - c : 0.0f;
        const float ey = y + 1 < g.d1 ? __ldg(v + i + sy) - c : 0.0f;
        const float ex = x + 1 < g.d2 ? __ldg(v + i + 1) - c : 0.0f;
        vacc += ez * qz; vacc += ey * qy; vacc += ex * qx;
        const float grad = (qzm - qz) + (qym - qy) + (qxm - qx);
        gv[i] = ACC ? gv[i] + k2 * grad : k2 * grad;
    }
    double bt = block_sum((double)vacc, red);
    grid_reduce_finish(bt, ws, out, scale, red);
}

// Welford: count = number of samples including x
__global__ void __launch_bounds__(256)
moments_update_kernel(const float *__restrict__ x, float *__restrict__ mean, float *__restrict__ m2, float inv_count,
                      int first, i64 n)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const float v = x[i];
        if (first) {
            mean[i] = v;
            m2[i] = 0.0f;
        } else {
            const float mu = mean[i];
            const float d = v - mu;
            const float mu2 = mu + d * inv_count;
            mean[i] = mu2;
            m2[i] += d * (v - mu2);
        }
    }
}

// Chan pairwise merge of (mean_a, m2_a, na) with (mean_b, m2_b, nb) into a
__global__ void __launch_bounds__(256)
moments_merge_kernel(float *__restrict__ mean_a, float *__restrict__ m2_a, const float *__restrict__ mean_b,
                     const float *__restrict__ m2_b, float wb, float wab, i64 n)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const float d = mean_b[i] - mean_a[i];
        mean_a[i] += d * wb;              // nb / (na + nb)
        m2_a[i] += m2_b[i] + d * d * wab;  // na * nb / (na + nb)
    }
}

__global__ void __launch_bounds__(256)
moments_std_kernel(const float *__restrict__ m2, float *__restrict__ out, float inv_nm1, i64 n)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        out[i] = sqrtf(m2[i] * inv_nm1);
}


// streaming per-voxel squared error of MC samples against the fixed image: acc += (x - y)^2  (the MSE map of
// evaluate.py:1538, mean_n((all_moved - y)^2), without the [N, ...] sample stack)
__global__ void __launch_bounds__(256)
sqerr_update_kernel(const float *__restrict__ x, const float *__restrict__ y, float *__restrict__ acc, int first, i64 n)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const float d = x[i] - y[i];
        acc[i] = first ? d * d : acc[i] + d * d;
    }
}

// Evaluate.ncc(var, mse) (evaluate.py:334-353, zero_norm=True): global zero-normalised cross-correlation of two
// maps, sum((a - mean a) / (std a * n + eps) * (v - mean v) / (std v + eps)) with population stds and eps = 1e-15.
// One pass: per-CTA double partials of (sum a, sum v, sum aa, sum vv, sum av) scaled per element by `sa` / `sv`
// (a = sa * a_in: the variance map is std^2 and the MSE map sum / N, so neither needs a pass of its own); the last
// CTA to arrive combines them in a fixed order.
constexpr int GNCC_MAX_CTAS = 1024;
struct GnccWs {
    unsigned int ticket, pad;
    double part[5][GNCC_MAX_CTAS];
};
__global__ void __launch_bounds__(256)
global_ncc_kernel(const float *__restrict__ a, const float *__restrict__ v, float sa, float sv, int square_a, i64 n,
                  GnccWs *ws, float *out)
{
    __shared__ double red[32];
    __shared__ bool is_last;
    double s[5] = {0, 0, 0, 0, 0};
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        float af = a[i];
        if (square_a) af = af * af;          // var = moved_std ** 2 (evaluate.py:1539), fp32 like the reference
        const double x = (double)(af * sa), y = (double)(v[i] * sv);
        s[0] += x; s[1] += y; s[2] += x * x; s[3] += y * y; s[4] += x * y;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double t = block_sum(s[k], red);
        if (threadIdx.x == 0) ws->part[k][blockIdx.x] = t;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double tot[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        double t = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += ((volatile double *)ws->part[k])[i];
        tot[k] = block_sum(t, red);
    }
    if (threadIdx.x == 0) {
        const double N = (double)n, ma = tot[0] / N, mv = tot[1] / N;
        const double va = fmax(tot[2] / N - ma * ma, 0.0), vv = fmax(tot[3] / N - mv * mv, 0.0);
        const double cov = tot[4] - N * ma * mv;
        const double eps = 1e-15;
        out[0] = (float)(cov / ((sqrt(va) * N + eps) * (sqrt(vv) + eps)));
        out[1] = (float)ma;                 // var.mean() of evaluate.py:1541 comes for free
        ws->ticket = 0;
    }
}

// total = sum of the step's per-term / per-level loss scalars (fixed order), optionally also accumulated into a
// running per-term sum -- replaces the ATen reduction that used to sit in the captured step, and lets multi-GPU
// runs all-reduce the loss scalars once per K steps instead of once per step
__global__ void loss_total_kernel(const float *__restrict__ losses, int rows, int cols, float *total, float *running,
                                  int accumulate)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float t = 0.0f;
    for (int r = 0; r < rows; ++r) {
        float rs = 0.0f;
        for (int c = 0; c < cols; ++c) rs += losses[r * cols + c];
        if (running) running[r] = accumulate ? running[r] + rs : rs;
        t += rs;
    }
    if (total) *total = t;
}

// All tracked maps of one MC sample in ONE launch, graph-capturable: the sample count lives in device memory (the
// kernel uses *count_dev + 1; pulpo_counter_add bumps it afterwards), so the same captured launch serves every
// sample.  Per map: Welford update of (mean, M2) and, where a target is given, the squared-error sum.
constexpr int MM_MAXMAPS = 32;
struct MomentsMaps {
    int n;
    pulpo_moments_map m[MM_MAXMAPS];
};
__global__ void __launch_bounds__(256)
moments_update_multi_kernel(const MomentsMaps maps, const int *__restrict__ count_dev)
{
    const int count = *count_dev + 1;
    const float inv_count = 1.0f / (float)count;
    const bool first = count == 1;
    for (int k = 0; k < maps.n; ++k) {
        const pulpo_moments_map &mp = maps.m[k];
        const float *__restrict__ x = mp.x;
        float *__restrict__ mean = mp.mean, *__restrict__ m2 = mp.m2, *__restrict__ acc = mp.sqerr_acc;
        const float *__restrict__ y = mp.target;
        const i64 n = mp.n;
        for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
            const float v = x[i];
            if (first) {
                mean[i] = v;
                m2[i] = 0.0f;
            } else {
                const float mu = mean[i];
                const float d = v - mu;
                const float mu2 = mu + d * inv_count;
                mean[i] = mu2;
                m2[i] += d * (v - mu2);
            }
            if (acc) {
                const float e = v - y[i];
                acc[i] = first ? e * e : acc[i] + e * e;
            }
        }
    }
}
__global__ void counter_add_kernel(int *ctr, int v, int reset)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) *ctr = reset ? v : *ctr + v;
}

// ---- MC sampling: z = mu + sigma * eps (gauss_sampler, src/network_blocks.py:7-8) for every level of one MC
// deformation sample in ONE graph-capturable launch.  eps comes from a counter-based generator (Philox4x32-10 +
// Box-Muller) keyed by (seed, sample id) and indexed by (level, element), so sample i is the same bits on whatever
// rank draws it; the sample id is first_id + id_stride * (*count_dev): the MC loop's running count lives on the
// device (pulpo_counter_add), so the captured launch serves every sample of the loop at evaluate.py:227-235.
constexpr int GS_MAXL = 8;
struct GaussLevels {
    int n;
    pulpo_gauss_level l[GS_MAXL];
};
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ float2 box_muller(unsigned int a, unsigned int b)
{
    const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);    // (0, 1)
    const float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    return make_float2(r * c, r * s);
}
__global__ void __launch_bounds__(256)
gauss_sample_multi_kernel(const GaussLevels lv, unsigned long long seed, const int *__restrict__ count_dev, int first_id,
                          int id_stride, float var)
{
    const unsigned int sample = (unsigned int)(first_id + id_stride * (count_dev ? *count_dev : 0));
    const uint2 key = make_uint2((unsigned int)seed, (unsigned int)(seed >> 32));
    for (int k = 0; k < lv.n; ++k) {
        const pulpo_gauss_level &L = lv.l[k];
        const i64 quads = (L.n + 3) / 4;
        for (i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x; q < quads; q += (i64)gridDim.x * blockDim.x) {
            const uint4 r = philox4x32_10(make_uint4((unsigned int)q, (unsigned int)(q >> 32), sample, (unsigned int)k), key);
            const float2 e0 = box_muller(r.x, r.y), e1 = box_muller(r.z, r.w);
            const float e[4] = {e0.x, e0.y, e1.x, e1.y};
            const i64 i0 = 4 * q;
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (i0 + t < L.n) {
                    const float eps = __fmul_rn(var, e[t]);
                    if (L.eps_out) L.eps_out[i0 + t] = eps;
                    L.z[i0 + t] = __fadd_rn(L.mu[i0 + t], __fmul_rn(L.sigma[i0 + t], eps));
                }
        }
    }
}

// Chan merge of W partial (mean, M2) slices in rank order and the unbiased std of the result in one pass (the reduction
// of MC statistics after the all_to_all: part r of a statistic sits at r * chunk).  Sequential merge in registers: same
// arithmetic and order as W - 1 calls of moments_merge_kernel followed by moments_std_kernel.
constexpr int MS_MAXW = 16;
struct MergeCounts {
    int w;
    int n[MS_MAXW];
};
__global__ void __launch_bounds__(256)
moments_merge_std_kernel(const float *__restrict__ mean_parts, const float *__restrict__ m2_parts, const MergeCounts c, i64 chunk,
                         float inv_nm1, float *__restrict__ std_out)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < chunk; i += (i64)gridDim.x * blockDim.x) {
        float mean = 0.0f, m2 = 0.0f;
        int seen = 0;
        for (int r = 0; r < c.w; ++r) {
            const int nb = c.n[r];
            if (nb == 0) continue;
            const float mb = mean_parts[r * chunk + i], qb = m2_parts[r * chunk + i];
            if (seen == 0) {
                mean = mb; m2 = qb;
            } else {
                const float tot = (float)(seen + nb);
                const float d = mb - mean;
                mean += d * ((float)nb / tot);
                m2 += qb + d * d * ((float)seen * (float)nb / tot);
            }
            seen += nb;
        }
        std_out[i] = sqrtf(m2 * inv_nm1);
    }
}

}  // namespace pulpo

using namespace pulpo;

extern "C" size_t pulpo_reduce_ws_bytes(void) { return kReduceWsBytes; }

extern "C" int pulpo_kl_diag_fwd(const float *mu0, const float *sigma0, const float *mu1, const float *sigma1,
                                 float eps, float weight, float *out, void *ws, size_t ws_bytes, int B, long long n,
                                 pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_kl_diag_fwd");
    PULPO_REQUIRE(mu0 && sigma0 && out && ws, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && n > 0, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(ws_bytes >= kReduceWsBytes, PULPO_ERR_WORKSPACE);
    const i64 total = (i64)B * n;
    bool al = aligned16(mu0) && aligned16(sigma0) && (!mu1 || aligned16(mu1)) && (!sigma1 || aligned16(sigma1));
    int grid = grid_for((total + 3) / 4, 256, 4);
    kl_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mu0, sigma0, mu1, sigma1, eps, out, (ReduceWs *)ws,
                                                        0.5 * (double)weight / (double)B, total, (al && (total & 3) == 0) ? 1 : 0);
    return launch_status();
}

extern "C" int pulpo_gauss_sample_kl_fwd(const float *mu, const float *sigma, const float *noise, const float *mu1,
                                        const float *sigma1, float var, float eps, float weight, float *z, float *out,
                                        void *ws, size_t ws_bytes, int B, long long n, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_gauss_sample_kl_fwd");
    PULPO_REQUIRE(mu && sigma && noise && z && out && ws, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && n > 0, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(ws_bytes >= kReduceWsBytes, PULPO_ERR_WORKSPACE);
    const i64 total = (i64)B * n;
    bool al = aligned16(mu) && aligned16(sigma) && aligned16(noise) && aligned16(z) && (!mu1 || aligned16(mu1)) &&
              (!sigma1 || aligned16(sigma1));
    int grid = grid_for((total + 3) / 4, 256, 4);
    gauss_kl_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mu, sigma, noise, mu1, sigma1, var, eps, z, out,
                                                              (ReduceWs *)ws, 0.5 * (double)weight / (double)B, total,
                                                              (al && (total & 3) == 0) ? 1 : 0);
    return launch_status();
}

extern "C" int pulpo_gauss_sample_kl_bwd(const float *gz, const float *gloss, int have_kl, const float *mu,
                                        const float *sigma, const float *noise, const float *mu1, const float *sigma1,
                                        float var, float eps, float weight, float *gmu, float *gsigma, int B,
                                        long long n, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_gauss_sample_kl_bwd");
    PULPO_REQUIRE(mu && sigma && noise && gmu && gsigma, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && n > 0, PULPO_ERR_INVALID_SHAPE);
    const i64 total = (i64)B * n;
    gauss_kl_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        gz, gloss, mu, sigma, noise, mu1, sigma1, var, eps, gmu, gsigma, weight / (float)B, have_kl, total);
    return launch_status();
}

extern "C" size_t pulpo_kl_multi_ws_bytes(void) { return 16 + sizeof(double) * KL_MAXL * 592; }

extern "C" int pulpo_kl_n01_multi(const pulpo_kl_level *levels, int nlevels, float eps, int B, void *ws, size_t ws_bytes,
                                  pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_kl_n01_multi");
    PULPO_REQUIRE(levels && ws, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(nlevels >= 1 && nlevels <= KL_MAXL && B > 0, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(ws_bytes >= pulpo_kl_multi_ws_bytes(), PULPO_ERR_WORKSPACE);
    KlMulti m;
    m.n = nlevels;
    i64 most = 0;
    for (int l = 0; l < nlevels; ++l) {
        const pulpo_kl_level &v = levels[l];
        PULPO_REQUIRE(v.mu && v.sigma && v.gmu && v.gsigma && v.out, PULPO_ERR_NULL_POINTER);
        PULPO_REQUIRE(v.n > 0, PULPO_ERR_INVALID_SHAPE);
        m.l[l].mu = v.mu; m.l[l].sg = v.sigma; m.l[l].gmu = v.gmu; m.l[l].gsg = v.gsigma; m.l[l].out = v.out;
        m.l[l].total = (i64)B * v.n;
        m.l[l].vec = (m.l[l].total % 4 == 0 && aligned16(v.mu) && aligned16(v.sigma) && aligned16(v.gmu) && aligned16(v.gsigma)) ? 1 : 0;
        m.l[l].k = v.weight / (float)B;
        m.l[l].scale = 0.5 * (double)v.weight / (double)B;
        if (m.l[l].total > most) most = m.l[l].total;
    }
    int grid = grid_for((most + 7) / 8, 256, 4);   // two quads per thread and pass
    if (grid > 592) grid = 592;
    kl_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(m, eps, (KlMultiWs *)ws);
    return launch_status();
}

extern "C" int pulpo_kl_diag_bwd(const float *gloss, const float *mu0, const float *sigma0, const float *mu1,
                                 const float *sigma1, float eps, float weight, float *gmu0, float *gsigma0, int B,
                                 long long n, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_kl_diag_bwd");
    PULPO_REQUIRE(mu0 && sigma0 && gmu0 && gsigma0, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && n > 0, PULPO_ERR_INVALID_SHAPE);
    const i64 total = (i64)B * n;
    kl_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(gloss, mu0, sigma0, mu1, sigma1, eps, gmu0,
                                                                        gsigma0, weight / (float)B, total);
    return launch_status();
}

extern "C" int pulpo_l2reg_fwd(const float *f, float lamb, float *out, void *ws, size_t ws_bytes, int B, int C,
                               int D0, int D1, int D2, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_l2reg_fwd");
    PULPO_REQUIRE(f && out && ws, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(ws_bytes >= kReduceWsBytes, PULPO_ERR_WORKSPACE);
    const i64 total = (i64)B * C * D0 * D1 * D2;
    const double cnt = (double)B * C * (D0 - 1) * (double)(D1 - 1) * (D2 - 1);
    const double scale = (double)lamb * D0 * D1 * D2 / cnt;
    RowGeom g;
    if (make_rowgeom(g, B * C, D0, D1, D2) && aligned16(f))
        l2reg_fwd_v4_kernel<<<grid_for(g.groups, 256, 4), 256, 0, (cudaStream_t)stream>>>(f, out, (ReduceWs *)ws, scale, g);
    else
        l2reg_fwd_kernel<<<grid_for(total, 256, 4), 256, 0, (cudaStream_t)stream>>>(f, out, (ReduceWs *)ws, scale,
                                                                                  B * C, D0, D1, D2);
    return launch_status();
}

extern "C" int pulpo_l2reg_bwd(const float *gloss, const float *f, float lamb, float *gf, int accumulate, int B,
                               int C, int D0, int D1, int D2, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_l2reg_bwd");
    PULPO_REQUIRE(f && gf, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2, PULPO_ERR_INVALID_SHAPE);
    const i64 total = (i64)B * C * D0 * D1 * D2;
    const double cnt = (double)B * C * (D0 - 1) * (double)(D1 - 1) * (D2 - 1);
    const float kk = (float)(2.0 * (double)lamb * D0 * D1 * D2 / cnt);
    RowGeom g;
    if (make_rowgeom(g, B * C, D0, D1, D2) && aligned16(f) && aligned16(gf))
        if (accumulate)
            l2reg_bwd_v4_kernel<true><<<(g.groups + 255) / 256, 256, 0, (cudaStream_t)stream>>>(gloss, f, gf, kk, g);
        else
            l2reg_bwd_v4_kernel<false><<<(g.groups + 255) / 256, 256, 0, (cudaStream_t)stream>>>(gloss, f, gf, kk, g);
    else
        l2reg_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(gloss, f, gf, kk, B * C, D0, D1, D2, accumulate);
    return launch_status();
}

// generic product for the fallback path of pulpo_l2reg_fwd_bwd
__global__ void __launch_bounds__(256)
prod_acc_kernel(const float *__restrict__ gout, const float *__restrict__ dpos, float *__restrict__ gf, i64 S, int C, i64 total)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const i64 bc = i / S, b = bc / C;
        gf[i] += gout[i - (bc - b) * S] * dpos[i];
    }
}

extern "C" int pulpo_l2reg_fwd_bwd(const float *f, float lamb, float *out, const float *gout, const float *dpos,
                                   float *gf, int accumulate, void *ws, size_t ws_bytes, int B, int C, int D0, int D1,
                                   int D2, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_l2reg_fwd_bwd");
    PULPO_REQUIRE(f && out && gf && ws, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE((gout == nullptr) == (dpos == nullptr), PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(ws_bytes >= kReduceWsBytes, PULPO_ERR_WORKSPACE);
    const i64 total = (i64)B * C * D0 * D1 * D2;
    const double cnt = (double)B * C * (D0 - 1) * (double)(D1 - 1) * (D2 - 1);
    const double scale = (double)lamb * D0 * D1 * D2 / cnt;
    const float kk = (float)(2.0 * scale);
    cudaStream_t st = (cudaStream_t)stream;
    RowGeom g;
    const bool vec_ok = make_rowgeom(g, B * C, D0, D1, D2) && aligned16(f) && aligned16(gf) &&
                        (!gout || (aligned16(gout) && aligned16(dpos)));
    if (vec_ok && D0 >= 8) {
        L2MarchGeom m;
        m.BC = B * C; m.D0 = D0; m.D1 = D1; m.D2 = D2; m.XG = D2 / 4; m.C = C;
        const i64 columns = (i64)m.BC * D1 * m.XG;
        i64 nz = (2ll * kSMs * 768 + columns - 1) / columns;      // ~2 runs per resident thread
        if (nz > D0 / 4) nz = D0 / 4;
        if (nz < 1) nz = 1;
        m.zrun = (int)((D0 + nz - 1) / nz);
        m.nzrun = (D0 + m.zrun - 1) / m.zrun;
        m.items = (unsigned int)(columns * m.nzrun);
        m.dXG = make_fastdiv(m.XG); m.dD1 = make_fastdiv(D1); m.dnz = make_fastdiv(m.nzrun); m.dC = make_fastdiv(C);
        const int grid = grid_for(m.items, 256, PULPO_L2M_CTAS);
        ReduceWs *w = (ReduceWs *)ws;
        if (gout) {
            if (accumulate) l2reg_fwd_bwd_march_kernel<true, true><<<grid, 256, 0, st>>>(f, gf, kk, out, w, scale, gout, dpos, m);
            else l2reg_fwd_bwd_march_kernel<false, true><<<grid, 256, 0, st>>>(f, gf, kk, out, w, scale, gout, dpos, m);
        } else {
            if (accumulate) l2reg_fwd_bwd_march_kernel<true, false><<<grid, 256, 0, st>>>(f, gf, kk, out, w, scale, nullptr, nullptr, m);
            else l2reg_fwd_bwd_march_kernel<false, false><<<grid, 256, 0, st>>>(f, gf, kk, out, w, scale, nullptr, nullptr, m);
        }
    } else if (vec_ok) {
        const int grid = grid_for(g.groups, 256, 8);
        ReduceWs *w = (ReduceWs *)ws;
        if (gout) {
            if (accumulate) l2reg_fwd_bwd_v4_kernel<true, true><<<grid, 256, 0, st>>>(f, gf, kk, out, w, scale, gout, dpos, C, g);
            else l2reg_fwd_bwd_v4_kernel<false, true><<<grid, 256, 0, st>>>(f, gf, kk, out, w, scale, gout, dpos, C, g);
        } else {
            if (accumulate) l2reg_fwd_bwd_v4_kernel<true, false><<<grid, 256, 0, st>>>(f, gf, kk, out, w, scale, nullptr, nullptr, C, g);
            else l2reg_fwd_bwd_v4_kernel<false, false><<<grid, 256, 0, st>>>(f, gf, kk, out, w, scale, nullptr, nullptr, C, g);
        }
    } else {   // rows that are not a multiple of 4 voxels: the generic kernels
        l2reg_fwd_kernel<<<grid_for(total, 256, 4), 256, 0, st>>>(f, out, (ReduceWs *)ws, scale, B * C, D0, D1, D2);
        l2reg_bwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(nullptr, f, gf, kk, B * C, D0, D1, D2, accumulate);
        if (gout) prod_acc_kernel<<<grid_for(total, 256), 256, 0, st>>>(gout, dpos, gf, (i64)D0 * D1 * D2, C, total);
    }
    return launch_status();
}

extern "C" size_t pulpo_l2reg_up2_scratch_bytes(int B, int C, int d0, int d1, int d2)
{
    if (B <= 0 || C <= 0 || d0 <= 0 || d1 <= 0 || d2 <= 0) return 0;
    return (size_t)5 * B * C * d0 * d1 * d2 * sizeof(float);
}

extern "C" int pulpo_l2reg_up2_fwd_bwd(const float *v, float lamb, float *out, float *gv, int accumulate, void *scratch,
                                       size_t scratch_bytes, void *ws, size_t ws_bytes, int B, int C, int d0, int d1,
                                       int d2, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_l2reg_up2_fwd_bwd");
    PULPO_REQUIRE(v && out && gv && ws && scratch, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && d0 >= 1 && d1 >= 1 && d2 >= 1, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(ws_bytes >= kReduceWsBytes, PULPO_ERR_WORKSPACE);
    PULPO_REQUIRE(scratch_bytes >= pulpo_l2reg_up2_scratch_bytes(B, C, d0, d1, d2), PULPO_ERR_WORKSPACE);
    const i64 total = (i64)B * C * d0 * d1 * d2;
    PULPO_REQUIRE(total < (1ll << 31), PULPO_ERR_INVALID_SHAPE);
    // the regulariser's own normalisation, on the FINE grid (src/losses.py:221-222)
    const double D0 = 2.0 * d0, D1 = 2.0 * d1, D2 = 2.0 * d2;
    const double cnt = (double)B * C * (D0 - 1) * (D1 - 1) * (D2 - 1);
    const double scale = (double)lamb * D0 * D1 * D2 / cnt;
    RegUpGeom g;
    g.BC = B * C; g.d0 = d0; g.d1 = d1; g.d2 = d2; g.total = (unsigned int)total;
    g.dd2 = make_fastdiv(d2); g.dd1 = make_fastdiv(d1); g.dd0 = make_fastdiv(d0);
    float *X1 = (float *)scratch, *Qx1 = X1 + total, *Y1 = Qx1 + total, *Qy1 = Y1 + total, *Qx2 = Qy1 + total;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(total, 256, 8);
    regup_x_kernel<<<grid, 256, 0, st>>>(v, X1, Qx1, g);
    regup_y_kernel<<<grid, 256, 0, st>>>(X1, Qx1, Y1, Qy1, Qx2, g);
    if (accumulate)
        regup_z_kernel<true><<<grid, 256, 0, st>>>(v, Y1, Qy1, Qx2, gv, (float)(2.0 * scale), out, (ReduceWs *)ws, scale, g);
    else
        regup_z_kernel<false><<<grid, 256, 0, st>>>(v, Y1, Qy1, Qx2, gv, (float)(2.0 * scale), out, (ReduceWs *)ws, scale, g);
    return launch_status();
}

extern "C" int pulpo_moments_update(const float *x, float *mean, float *m2, int count, long long n,
                                    pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_moments_update");
    PULPO_REQUIRE(x && mean && m2, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(count >= 1 && n > 0, PULPO_ERR_INVALID_SHAPE);
    moments_update_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, mean, m2, 1.0f / (float)count,
                                                                            count == 1, n);
    return launch_status();
}

extern "C" int pulpo_moments_merge(float *mean_a, float *m2_a, int count_a, const float *mean_b, const float *m2_b,
                                   int count_b, long long n, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_moments_merge");
    PULPO_REQUIRE(mean_a && m2_a && mean_b && m2_b, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(count_a >= 0 && count_b >= 0 && count_a + count_b > 0 && n > 0, PULPO_ERR_INVALID_SHAPE);
    const float tot = (float)(count_a + count_b);
    moments_merge_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        mean_a, m2_a, mean_b, m2_b, (float)count_b / tot, (float)count_a * (float)count_b / tot, n);
    return launch_status();
}

extern "C" int pulpo_moments_std(const float *m2, float *std_out, int count, long long n, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_moments_std");
    PULPO_REQUIRE(m2 && std_out, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(count >= 2 && n > 0, PULPO_ERR_INVALID_SHAPE);
    moments_std_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(m2, std_out, 1.0f / (float)(count - 1), n);
    return launch_status();
}

extern "C" int pulpo_sqerr_update(const float *x, const float *y, float *acc, int first, long long n, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_sqerr_update");
    PULPO_REQUIRE(x && y && acc, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(n > 0, PULPO_ERR_INVALID_SHAPE);
    sqerr_update_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, acc, first, n);
    return launch_status();
}

extern "C" size_t pulpo_global_ncc_ws_bytes(void) { return sizeof(GnccWs); }

extern "C" int pulpo_global_ncc(const float *a, const float *v, float scale_a, float scale_v, int square_a, long long n,
                                float *out2, void *ws, size_t ws_bytes, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_global_ncc");
    PULPO_REQUIRE(a && v && out2 && ws, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(n > 0, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(ws_bytes >= sizeof(GnccWs), PULPO_ERR_WORKSPACE);
    int grid = grid_for(n, 256, 4);
    if (grid > GNCC_MAX_CTAS) grid = GNCC_MAX_CTAS;
    global_ncc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, v, scale_a, scale_v, square_a, n, (GnccWs *)ws, out2);
    return launch_status();
}

extern "C" int pulpo_loss_total(const float *losses, int rows, int cols, float *total, float *running, int accumulate,
                                pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_loss_total");
    PULPO_REQUIRE(losses && (total || running), PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(rows > 0 && cols > 0 && rows * cols <= 4096, PULPO_ERR_INVALID_SHAPE);
    loss_total_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(losses, rows, cols, total, running, accumulate);
    return launch_status();
}

extern "C" int pulpo_moments_update_multi(const pulpo_moments_map *maps, int nmaps, const int *count_dev, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_moments_update_multi");
    PULPO_REQUIRE(maps && count_dev, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(nmaps >= 1 && nmaps <= MM_MAXMAPS, PULPO_ERR_INVALID_SHAPE);
    MomentsMaps mm;
    mm.n = nmaps;
    long long biggest = 0;
    for (int k = 0; k < nmaps; ++k) {
        PULPO_REQUIRE(maps[k].x && maps[k].mean && maps[k].m2 && (!maps[k].sqerr_acc || maps[k].target), PULPO_ERR_NULL_POINTER);
        PULPO_REQUIRE(maps[k].n > 0, PULPO_ERR_INVALID_SHAPE);
        mm.m[k] = maps[k];
        if (maps[k].n > biggest) biggest = maps[k].n;
    }
    moments_update_multi_kernel<<<grid_for(biggest, 256, 8), 256, 0, (cudaStream_t)stream>>>(mm, count_dev);
    return launch_status();
}

extern "C" int pulpo_counter_add(int *counter_dev, int value, int reset, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_counter_add");
    PULPO_REQUIRE(counter_dev, PULPO_ERR_NULL_POINTER);
    counter_add_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counter_dev, value, reset);
    return launch_status();
}

extern "C" int pulpo_gauss_sample_multi(const pulpo_gauss_level *levels, int nlevels, unsigned long long seed,
                                        const int *count_dev, int first_id, int id_stride, float var, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_gauss_sample_multi");
    PULPO_REQUIRE(levels, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(nlevels >= 1 && nlevels <= GS_MAXL, PULPO_ERR_INVALID_SHAPE);
    GaussLevels g;
    g.n = nlevels;
    long long biggest = 0;
    for (int k = 0; k < nlevels; ++k) {
        PULPO_REQUIRE(levels[k].mu && levels[k].sigma && levels[k].z, PULPO_ERR_NULL_POINTER);
        PULPO_REQUIRE(levels[k].n > 0, PULPO_ERR_INVALID_SHAPE);
        g.l[k] = levels[k];
        if (levels[k].n > biggest) biggest = levels[k].n;
    }
    gauss_sample_multi_kernel<<<grid_for((biggest + 3) / 4, 256, 8), 256, 0, (cudaStream_t)stream>>>(g, seed, count_dev, first_id,
                                                                                                 id_stride, var);
    return launch_status();
}

extern "C" int pulpo_moments_merge_std(const float *mean_parts, const float *m2_parts, const int *counts, int nparts,
                                       long long chunk, float *std_out, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_moments_merge_std");
    PULPO_REQUIRE(mean_parts && m2_parts && counts && std_out, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(nparts >= 1 && nparts <= MS_MAXW && chunk > 0, PULPO_ERR_INVALID_SHAPE);
    MergeCounts c;
    c.w = nparts;
    long long total = 0;
    for (int r = 0; r < nparts; ++r) {
        PULPO_REQUIRE(counts[r] >= 0, PULPO_ERR_INVALID_SHAPE);
        c.n[r] = counts[r];
        total += counts[r];
    }
    PULPO_REQUIRE(total >= 2, PULPO_ERR_INVALID_SHAPE);
    moments_merge_std_kernel<<<grid_for(chunk, 256, 8), 256, 0, (cudaStream_t)stream>>>(mean_parts, m2_parts, c, chunk,
                                                                                       1.0f / (float)(total - 1), std_out);
    return launch_status();
}

extern "C" int pulpo_version(void) { return PULPO_B200_VERSION; }

extern "C" const char *pulpo_strerror(int status)
{
    switch (status) {
        case PULPO_OK: return "ok";
        case PULPO_ERR_NULL_POINTER: return "null pointer for a required argument";
        case PULPO_ERR_INVALID_SHAPE: return "invalid shape (sizes must be positive; warp axes need S >= 2)";
        case PULPO_ERR_UNSUPPORTED: return "unsupported argument (window/factor/coord_mode/alignment)";
        case PULPO_ERR_WORKSPACE: return "workspace too small or misaligned";
        case PULPO_ERR_CUDA: return "CUDA launch error";
        default: return "unknown status";
    }
}
