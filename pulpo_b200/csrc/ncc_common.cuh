// ncc_common.cuh -- arithmetic shared by the two NCC box-sum kernels (ncc.cu: generic loads,
// ncc_tma.cu: TMA-staged tiles).
#pragma once
#include "common.cuh"

namespace pulpo {

__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float add2(float a, float b) { return __fadd_rn(a, b); }

// sums of W consecutive values for 4 consecutive outputs: o[j] = sum a[j .. j+W-1]
template <int W, typename T>
__device__ __forceinline__ void xsum4(const T (&a)[W + 3], T (&o)[4])
{
    if (W >= 5) {
        T core = a[3];
#pragma unroll
        for (int t = 4; t < W; ++t) core = add2(core, a[t]);
        const T p12 = add2(a[1], a[2]), pw = add2(a[W], a[W + 1]);
        o[0] = add2(core, add2(a[0], p12));
        o[1] = add2(core, add2(p12, a[W]));
        o[2] = add2(core, add2(a[2], pw));
        o[3] = add2(core, add2(pw, a[W + 2]));
    } else {  // W == 3
        const T p12 = add2(a[1], a[2]), p34 = add2(a[3], a[4]);
        o[0] = add2(a[0], p12);
        o[1] = add2(p12, a[3]);
        o[2] = add2(a[2], p34);
        o[3] = add2(p34, a[5]);
    }
}

// correctly rounded x / W with the precomputed reciprocal (see div_by_axis)
__device__ __forceinline__ float div_by_W(float x, float Wf, float rcpW)
{
    float q0 = __fmul_rn(x, rcpW);
    float r0 = __fmaf_rn(-Wf, q0, x);
    float q1 = __fmaf_rn(r0, rcpW, q0);
    float r1 = __fmaf_rn(-Wf, q1, x);
    return __fmaf_rn(r1, rcpW, q1);
}


// fwd epilogue: cc and the backward coefficients from the five window sums; same expanded
// (cancellation-prone) formulas as the reference (src/losses.py:124-133), op by op.  u = sum / W
// uses the rounded reciprocal: cross, I_var and J_var are first-order insensitive to the rounding
// of u (d cross / d u_I = -J_sum + u_J * W = 0), and exact zeros stay exact zeros.
struct NccPoint {
    float cc, a, b, c;
};

template <bool COEF>
__device__ __forceinline__ NccPoint ncc_point(float sI, float sJ, float sII, float sJJ, float sIJ, float Wf, float rcpW)
{
    const float uI = __fmul_rn(sI, rcpW), uJ = __fmul_rn(sJ, rcpW);
    float cross = __fsub_rn(sIJ, __fmul_rn(uJ, sI));
    cross = __fsub_rn(cross, __fmul_rn(uI, sJ));
    cross = __fadd_rn(cross, __fmul_rn(__fmul_rn(uI, uJ), Wf));
    float Iv = __fsub_rn(sII, __fmul_rn(__fmul_rn(2.0f, uI), sI));
    Iv = __fadd_rn(Iv, __fmul_rn(__fmul_rn(uI, uI), Wf));
    float Jv = __fsub_rn(sJJ, __fmul_rn(__fmul_rn(2.0f, uJ), sJ));
    Jv = __fadd_rn(Jv, __fmul_rn(__fmul_rn(uJ, uJ), Wf));
    const float Dn = __fadd_rn(__fmul_rn(Iv, Jv), 1e-8f);
    const float c2 = __fmul_rn(cross, cross);
    const float rD = __fdividef(1.0f, Dn);
    NccPoint r;
    r.cc = c2 * rD;
    if (COEF) {
        r.a = 2.0f * cross * rD;
        r.c = -(c2 * Iv) * (rD * rD);
        r.b = -(r.a * sI + 2.0f * r.c * sJ) * rcpW;
    } else {
        r.a = r.b = r.c = 0.0f;
    }
    return r;
}

constexpr int NCC_MAX_CTAS = 320;   // >= 2 CTAs per SM
struct NccTmaParams {
    const float *I, *J;            // bwd epilogue (target, pred)
    float *o0, *o1, *o2;           // fwd: a, b, c (nullable) ; bwd: o0 = gpred
    const float *gloss;            // bwd: upstream scalar (nullable)
    float *loss;
    ReduceWs *ws;
    double loss_scale;             // -gamma / B
    float k;                       // bwd: -gamma / B
    float Wf, rcpW;
    int BC, D0, D1, D2, xt, yt;
    long long total_planes;        // BC * yt * xt * D0
    int zchunks;                   // > 0: aligned (column, z chunk) grid with this many chunks per column
    int nbounds;                   // > 0: cost-balanced split, CTA i owns linear planes [bounds[i], bounds[i+1])
    int bounds[NCC_MAX_CTAS + 1];
};


// ncc_tma.cu
bool ncc_tma_eligible(const float *in0, const float *in1, const float *in2, int D0, int D1, int D2, int win);
int ncc_tma_grid(int BC, int D0, int D1, int D2, int win);
int ncc_tma_launch(const float *in0, const float *in1, const float *in2, NccTmaParams p, int win, cudaStream_t st);

}  // namespace pulpo
