// ncc.cu -- windowed local NCC loss (reference: NCC_loss, src/losses.py:85-135) and its
// closed-form backward.
//
// The reference computes five win^3 box sums (I, J, I^2, J^2, IJ) as dense conv3d calls
// (729 taps each at win=9).  Here they are separable direct sums inside one streaming kernel:
// a CTA owns a (TY x TX) tile in (D1, D2) and marches along D0.
//   phase 1  x-sums: a thread reads 12 consecutive values per input with three 128-bit loads
//            (zero padded at the volume border exactly like conv3d's padding), forms the
//            products in registers and writes 4 x-summed values per quantity to shared memory
//            (15 adds per quantity for 4 outputs by sharing the common core of the windows);
//   phase 2  y-sums from shared memory (two output rows per thread, shared core), then the z
//            window as a register ring of the last W plane sums -- no subtraction, so no
//            running-sum drift and exact zeros stay exact zeros (the conditioning issue of
//            SURVEY.md 9.5).
// The kernel is bound by instruction issue, not HBM (8 B/voxel leaves ~45 issue slots per voxel
// per SM), so the five quantities travel as two packed float2 (I,J), (I^2,J^2) plus one scalar IJ
// and are summed with Blackwell's packed-fp32 FADD2/FMUL2 (__fadd2_rn/__fmul2_rn): 3 adds per
// tap instead of 5.  The loss is reduced warp-shuffle -> CTA -> deterministic two-stage grid sum
// in double.  Forward optionally emits the coefficient volumes (a, b, c) whose box filter is the
// gradient; backward runs the same box kernel on them (packed (a,b) + scalar c).
// Algorithmic bytes: fwd 8 B/voxel (+12 when saving a,b,c), bwd 12 B/voxel (+12 reading a,b,c).
#include "common.cuh"
#include "ncc_common.cuh"

namespace pulpo {

constexpr int NCC_TX = 32;  // tile width along D2  (one warp of columns)
constexpr int NCC_TY = 16;  // tile height along D1
constexpr int NCC_XS = 4;   // x outputs per phase-1 work item
constexpr int NCC_THREADS = NCC_TX * NCC_TY / 2;

struct NccParams {
    const float *in0, *in1, *in2;  // fwd: target I, pred J ; bwd: a, b, c
    const float *I, *J;            // bwd epilogue
    float *o0, *o1, *o2;           // fwd: a, b, c (nullable) ; bwd: o0 = gpred
    const float *gloss;            // bwd: upstream scalar (nullable)
    float *loss;
    ReduceWs *ws;
    double loss_scale;             // -gamma / B
    float k;                       // bwd: -gamma / B
    float Wf, rcpW;                // win^3 and RN(1/win^3)
    int BC, D0, D1, D2, zchunk, nzchunks;
};

template <int W, bool FWD, bool VECLOAD>
__global__ void __launch_bounds__(NCC_THREADS, 2)
ncc_box_kernel(const NccParams p)
{
    constexpr int R = W / 2;
    constexpr int ROWS = NCC_TY + 2 * R;
    constexpr int SEGS = NCC_TX / NCC_XS;
    constexpr int NV = NCC_XS + 2 * R;  // inputs per item
    constexpr int PAD = ((R + 3) / 4) * 4;  // aligned halo: loads cover [xq-PAD, xq+4+PAD)
    constexpr int NL = 2 * PAD + 4;
    // packed quantities: A = (I, J) | (a, b);  B = (I^2, J^2) [fwd only];  C = IJ | c
    __shared__ __align__(16) float2 XA[ROWS][NCC_TX];
    __shared__ __align__(16) float2 XB[FWD ? ROWS : 1][NCC_TX];
    __shared__ __align__(16) float XC[ROWS][NCC_TX];
    __shared__ double red[32];

    const int D0 = p.D0, D1 = p.D1, D2 = p.D2;
    const int sy = D2, sz = D1 * D2;
    const i64 S = (i64)D0 * sz;
    const int x0 = blockIdx.x * NCC_TX, y0 = blockIdx.y * NCC_TY;
    const int bc = blockIdx.z / p.nzchunks, zc = blockIdx.z % p.nzchunks;
    const int z_start = zc * p.zchunk, z_end = min(D0, z_start + p.zchunk);
    const float *in0 = p.in0 + (i64)bc * S, *in1 = p.in1 + (i64)bc * S;
    const float *in2 = FWD ? nullptr : p.in2 + (i64)bc * S;

    // phase-2 ownership: column tx, output rows 2*yp and 2*yp+1
    const int tx = threadIdx.x % NCC_TX, yp = threadIdx.x / NCC_TX;
    const int gx = x0 + tx, gy = y0 + 2 * yp;
    float2 rA[W][2], rB[FWD ? W : 1][2];
    float rC[W][2];
#pragma unroll
    for (int s = 0; s < W; ++s) {
        rA[s][0] = rA[s][1] = make_float2(0.f, 0.f);
        if (FWD) rB[s % (FWD ? W : 1)][0] = rB[s % (FWD ? W : 1)][1] = make_float2(0.f, 0.f);
        rC[s][0] = rC[s][1] = 0.0f;
    }

    float cc_acc = 0.0f;
    const float gk = FWD ? 0.0f : ((p.gloss ? __ldg(p.gloss) : 1.0f) * p.k);
    const int nplanes = (z_end - z_start) + 2 * R;

    for (int zp0 = 0; zp0 < nplanes; zp0 += W) {
#pragma unroll
        for (int s = 0; s < W; ++s) {
            const int zp = zp0 + s;
            if (zp >= nplanes) break;
            const int zin = z_start - R + zp;
            const bool plane_ok = (zin >= 0 && zin < D0);
            // ---------------- phase 1: x-sums of this input plane into shared memory
            for (int item = threadIdx.x; item < ROWS * SEGS; item += NCC_THREADS) {
                const int r = item / SEGS, sg = item % SEGS;
                const int yy = y0 - R + r;
                float2 oA[4], oB[4];
                float oC[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    oA[j] = oB[j] = make_float2(0.f, 0.f);
                    oC[j] = 0.0f;
                }
                if (plane_ok && yy >= 0 && yy < D1) {
                    const int rowoff = zin * sz + yy * sy;
                    float v0[NL], v1[NL], v2[NL];
                    if (VECLOAD) {
                        // aligned 128-bit loads cover [xq-PAD, xq+4+PAD); a float4 is wholly in or out
                        const int xq = x0 + sg * NCC_XS;
#pragma unroll
                        for (int q4 = 0; q4 < NL / 4; ++q4) {
                            const int xx = xq - PAD + 4 * q4;
                            const bool ok = (xx >= 0 && xx < D2);
                            float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0, t2 = t0;
                            if (ok) {
                                t0 = __ldg(reinterpret_cast<const float4 *>(in0 + rowoff + xx));
                                t1 = __ldg(reinterpret_cast<const float4 *>(in1 + rowoff + xx));
                                if (!FWD) t2 = __ldg(reinterpret_cast<const float4 *>(in2 + rowoff + xx));
                            }
                            v0[4 * q4] = t0.x; v0[4 * q4 + 1] = t0.y; v0[4 * q4 + 2] = t0.z; v0[4 * q4 + 3] = t0.w;
                            v1[4 * q4] = t1.x; v1[4 * q4 + 1] = t1.y; v1[4 * q4 + 2] = t1.z; v1[4 * q4 + 3] = t1.w;
                            v2[4 * q4] = t2.x; v2[4 * q4 + 1] = t2.y; v2[4 * q4 + 2] = t2.z; v2[4 * q4 + 3] = t2.w;
                        }
                    } else {
                        const int xb = x0 + sg * NCC_XS - PAD;
#pragma unroll
                        for (int j = 0; j < NL; ++j) {
                            const int xx = xb + j;
                            const bool ok = (xx >= 0 && xx < D2) && (j >= PAD - R) && (j < PAD - R + NV);
                            v0[j] = ok ? __ldg(in0 + rowoff + xx) : 0.0f;
                            v1[j] = ok ? __ldg(in1 + rowoff + xx) : 0.0f;
                            v2[j] = (ok && !FWD) ? __ldg(in2 + rowoff + xx) : 0.0f;
                        }
                    }
                    float2 a[NV], b[NV];
                    float c[NV];
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        const int jj = j + PAD - R;   // window of output 0 starts at x - R
                        a[j] = make_float2(v0[jj], v1[jj]);
                        if (FWD) {
                            b[j] = __fmul2_rn(a[j], a[j]);
                            c[j] = __fmul_rn(v0[jj], v1[jj]);
                        } else {
                            c[j] = v2[jj];
                        }
                    }
                    xsum4<W>(a, oA);
                    if (FWD) xsum4<W>(b, oB);
                    xsum4<W>(c, oC);
                }
                float4 *da = reinterpret_cast<float4 *>(&XA[r][sg * NCC_XS]);
                da[0] = make_float4(oA[0].x, oA[0].y, oA[1].x, oA[1].y);
                da[1] = make_float4(oA[2].x, oA[2].y, oA[3].x, oA[3].y);
                if (FWD) {
                    float4 *db = reinterpret_cast<float4 *>(&XB[r][sg * NCC_XS]);
                    db[0] = make_float4(oB[0].x, oB[0].y, oB[1].x, oB[1].y);
                    db[1] = make_float4(oB[2].x, oB[2].y, oB[3].x, oB[3].y);
                }
                *reinterpret_cast<float4 *>(&XC[r][sg * NCC_XS]) = make_float4(oC[0], oC[1], oC[2], oC[3]);
            }
            __syncthreads();
            // ---------------- phase 2: y-sums (two rows sharing the window core) + z ring
            {
                float2 coreA = XA[2 * yp + 1][tx];
                float coreC = XC[2 * yp + 1][tx];
#pragma unroll
                for (int t = 2; t < W; ++t) {
                    coreA = add2(coreA, XA[2 * yp + t][tx]);
                    coreC = add2(coreC, XC[2 * yp + t][tx]);
                }
                rA[s][0] = add2(XA[2 * yp][tx], coreA);
                rA[s][1] = add2(coreA, XA[2 * yp + W][tx]);
                rC[s][0] = add2(XC[2 * yp][tx], coreC);
                rC[s][1] = add2(coreC, XC[2 * yp + W][tx]);
                if (FWD) {
                    float2 coreB = XB[2 * yp + 1][tx];
#pragma unroll
                    for (int t = 2; t < W; ++t) coreB = add2(coreB, XB[2 * yp + t][tx]);
                    rB[FWD ? s : 0][0] = add2(XB[2 * yp][tx], coreB);
                    rB[FWD ? s : 0][1] = add2(coreB, XB[2 * yp + W][tx]);
                }
            }
            __syncthreads();
            const int zout = zin - R;
            if (zout >= z_start && zout < z_end && gx < D2) {
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    if (gy + o >= D1) continue;
                    float2 sA = rA[0][o], sB = FWD ? rB[0][o] : make_float2(0.f, 0.f);
                    float sC = rC[0][o];
#pragma unroll
                    for (int u = 1; u < W; ++u) {
                        sA = add2(sA, rA[u][o]);
                        if (FWD) sB = add2(sB, rB[u % (FWD ? W : 1)][o]);
                        sC = add2(sC, rC[u][o]);
                    }
                    const i64 off = (i64)bc * S + (i64)zout * sz + (gy + o) * sy + gx;
                    if (FWD) {
                        const float sI = sA.x, sJ = sA.y, sII = sB.x, sJJ = sB.y, sIJ = sC;
                        const float Wf = p.Wf;
                        const float uI = div_by_W(sI, Wf, p.rcpW), uJ = div_by_W(sJ, Wf, p.rcpW);
                        // same expanded (cancellation-prone) formulas as the reference, op by op
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
                        cc_acc += c2 * rD;
                        if (p.o0) {
                            const float a = 2.0f * cross * rD;
                            const float c = -(c2 * Iv) * (rD * rD);
                            p.o0[off] = a;
                            p.o1[off] = -(a * sI + 2.0f * c * sJ) * p.rcpW;
                            p.o2[off] = c;
                        }
                    } else {
                        const float Iv = __ldg(p.I + off), Jv = __ldg(p.J + off);
                        p.o0[off] = gk * (Iv * sA.x + sA.y + 2.0f * Jv * sC);
                    }
                }
            }
        }
    }
    if (FWD) {
        double bt = block_sum((double)cc_acc, red);
        grid_reduce_finish(bt, p.ws, p.loss, p.loss_scale, red);
    }
}

struct NccGrid {
    dim3 grid;
    int zchunk, nzchunks;
};

static NccGrid ncc_grid(int BC, int D0, int D1, int D2, int win)
{
    NccGrid g;
    const int xt = (D2 + NCC_TX - 1) / NCC_TX, yt = (D1 + NCC_TY - 1) / NCC_TY;
    const i64 tiles = (i64)xt * yt * BC;
    // enough z-chunks for ~2 full waves (148 SMs x 2 resident CTAs), but keep the 2R halo
    // planes a small fraction of each chunk
    i64 want = (2 * kSMs * 2 + tiles - 1) / tiles;
    int maxchunks = D0 / (2 * win) > 0 ? D0 / (2 * win) : 1;
    int n = (int)(want < 1 ? 1 : (want > maxchunks ? maxchunks : want));
    g.zchunk = (D0 + n - 1) / n;
    g.nzchunks = (D0 + g.zchunk - 1) / g.zchunk;
    g.grid = dim3(xt, yt, BC * g.nzchunks);
    return g;
}

template <bool FWD, bool VEC>
static int ncc_launch_w(const NccParams &p, const NccGrid &g, int win, cudaStream_t st)
{
    switch (win) {
        case 3: ncc_box_kernel<3, FWD, VEC><<<g.grid, NCC_THREADS, 0, st>>>(p); break;
        case 5: ncc_box_kernel<5, FWD, VEC><<<g.grid, NCC_THREADS, 0, st>>>(p); break;
        case 7: ncc_box_kernel<7, FWD, VEC><<<g.grid, NCC_THREADS, 0, st>>>(p); break;
        case 9: ncc_box_kernel<9, FWD, VEC><<<g.grid, NCC_THREADS, 0, st>>>(p); break;
        case 11: ncc_box_kernel<11, FWD, VEC><<<g.grid, NCC_THREADS, 0, st>>>(p); break;
        default: return PULPO_ERR_UNSUPPORTED;
    }
    return launch_status();
}

template <bool FWD>
static int ncc_launch(const NccParams &p, const NccGrid &g, int win, cudaStream_t st)
{
    bool vec = (p.D2 % 4 == 0) && aligned16(p.in0) && aligned16(p.in1) && (FWD || aligned16(p.in2)) &&
               (i64)p.D0 * p.D1 * p.D2 < (1ll << 31);
    return vec ? ncc_launch_w<FWD, true>(p, g, win, st) : ncc_launch_w<FWD, false>(p, g, win, st);
}

}  // namespace pulpo

using namespace pulpo;

extern "C" size_t pulpo_ncc_ws_bytes(int B, int C, int D0, int D1, int D2)
{
    // worst case over the supported windows: the grid is largest for the smallest window
    NccGrid g = ncc_grid(B * C, D0, D1, D2, 3);
    size_t ctas = (size_t)g.grid.x * g.grid.y * g.grid.z;
    if (ctas < 1024) ctas = 1024;   // the persistent TMA kernel deposits one partial per SM
    return 16 + sizeof(double) * ctas;
}

extern "C" int pulpo_ncc_fwd(const float *pred, const float *target, float *loss, float *abc, void *ws,
                             size_t ws_bytes, int win, float gamma, int B, int C, int D0, int D1, int D2,
                             pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_ncc_fwd");
    PULPO_REQUIRE(pred && target && loss && ws, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 > 0 && D1 > 0 && D2 > 0, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE((i64)D0 * D1 * D2 < (1ll << 31), PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(win >= 3 && win <= 11 && (win & 1), PULPO_ERR_UNSUPPORTED);
    const i64 n = (i64)B * C * D0 * D1 * D2;
    if (ncc_tma_eligible(target, pred, nullptr, D0, D1, D2, win) &&
        ws_bytes >= 16 + sizeof(double) * (size_t)ncc_tma_grid(B * C, D0, D1, D2, win)) {
        NccTmaParams t{};
        t.o0 = abc; t.o1 = abc ? abc + n : nullptr; t.o2 = abc ? abc + 2 * n : nullptr;
        t.loss = loss; t.ws = (ReduceWs *)ws;
        t.loss_scale = -(double)gamma / (double)B;
        t.Wf = (float)(win * win * win);
        t.rcpW = 1.0f / t.Wf;
        t.BC = B * C; t.D0 = D0; t.D1 = D1; t.D2 = D2;
        return ncc_tma_launch(target, pred, nullptr, t, win, (cudaStream_t)stream);
    }
    NccGrid g = ncc_grid(B * C, D0, D1, D2, win);
    PULPO_REQUIRE(g.grid.y <= 65535 && g.grid.z <= 65535, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(ws_bytes >= 16 + sizeof(double) * (size_t)g.grid.x * g.grid.y * g.grid.z, PULPO_ERR_WORKSPACE);
    NccParams p{};
    p.in0 = target; p.in1 = pred; p.in2 = nullptr;
    p.o0 = abc; p.o1 = abc ? abc + n : nullptr; p.o2 = abc ? abc + 2 * n : nullptr;
    p.loss = loss; p.ws = (ReduceWs *)ws;
    p.loss_scale = -(double)gamma / (double)B;
    p.Wf = (float)(win * win * win);
    p.rcpW = 1.0f / p.Wf;
    p.BC = B * C; p.D0 = D0; p.D1 = D1; p.D2 = D2; p.zchunk = g.zchunk; p.nzchunks = g.nzchunks;
    return ncc_launch<true>(p, g, win, (cudaStream_t)stream);
}

extern "C" int pulpo_ncc_bwd(const float *abc, const float *pred, const float *target, const float *gloss,
                             float *gpred, int win, float gamma, int B, int C, int D0, int D1, int D2,
                             pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_ncc_bwd");
    PULPO_REQUIRE(abc && pred && target && gpred, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 > 0 && D1 > 0 && D2 > 0, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE((i64)D0 * D1 * D2 < (1ll << 31), PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(win >= 3 && win <= 11 && (win & 1), PULPO_ERR_UNSUPPORTED);
    const i64 n = (i64)B * C * D0 * D1 * D2;
    if (ncc_tma_eligible(abc, abc + n, abc + 2 * n, D0, D1, D2, win) && aligned16(gpred)) {
        NccTmaParams t{};
        t.I = target; t.J = pred; t.o0 = gpred; t.gloss = gloss;
        t.k = -gamma / (float)B;
        t.Wf = (float)(win * win * win);
        t.rcpW = 1.0f / t.Wf;
        t.BC = B * C; t.D0 = D0; t.D1 = D1; t.D2 = D2;
        return ncc_tma_launch(abc, abc + n, abc + 2 * n, t, win, (cudaStream_t)stream);
    }
    NccGrid g = ncc_grid(B * C, D0, D1, D2, win);
    PULPO_REQUIRE(g.grid.y <= 65535 && g.grid.z <= 65535, PULPO_ERR_INVALID_SHAPE);
    NccParams p{};
    p.in0 = abc; p.in1 = abc + n; p.in2 = abc + 2 * n;
    p.I = target; p.J = pred; p.o0 = gpred; p.gloss = gloss;
    p.k = -gamma / (float)B;
    p.Wf = (float)(win * win * win);
    p.rcpW = 1.0f / p.Wf;
    p.BC = B * C; p.D0 = D0; p.D1 = D1; p.D2 = D2; p.zchunk = g.zchunk; p.nzchunks = g.nzchunks;
    return ncc_launch<false>(p, g, win, (cudaStream_t)stream);
}
