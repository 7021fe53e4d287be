// ncc.cu -- windowed local NCC loss (reference: NCC_loss, src/losses.py:85-135) and its
// closed-form backward.
//
// The reference computes five win^3 box sums (I, J, I^2, J^2, IJ) as dense conv3d calls
// (729 taps each at win=9).  Here they are separable direct sums inside one streaming kernel:
// a CTA owns a (TY x TX) tile in (D1, D2) and marches along D0.
//   phase 1  x-sums: a thread reads XS+2R consecutive values per input from global (zero padded
//            at the volume border exactly like conv3d's padding), forms the products in
//            registers and writes XS x-summed values per quantity to shared memory;
//   phase 2  y-sums from shared memory (two output rows per thread), then the z window as a
//            register ring of the last W plane sums -- no subtraction, so no running-sum drift
//            and exact zeros stay exact zeros (the conditioning issue of SURVEY.md 9.5).
// The loss is reduced warp-shuffle -> CTA -> deterministic two-stage grid sum in double.
// Forward optionally emits the three coefficient volumes (a, b, c) whose box filter is the
// gradient; backward runs the same box kernel on them.
// Algorithmic bytes: fwd 8 B/voxel (+12 when saving a,b,c), bwd 12 B/voxel (+12 reading a,b,c).
#include "common.cuh"

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
    float Wf;                      // win^3
    int BC, D0, D1, D2, zchunk, nzchunks;
};

template <int W, bool FWD>
__global__ void __launch_bounds__(NCC_THREADS)
ncc_box_kernel(const NccParams p)
{
    constexpr int R = W / 2;
    constexpr int NQ = FWD ? 5 : 3;
    constexpr int NIN = FWD ? 2 : 3;
    constexpr int ROWS = NCC_TY + 2 * R;
    constexpr int SEGS = NCC_TX / NCC_XS;
    __shared__ __align__(16) float X[NQ][ROWS][NCC_TX];
    __shared__ double red[32];

    const int D0 = p.D0, D1 = p.D1, D2 = p.D2;
    const i64 sy = D2, sz = (i64)D1 * D2, S = (i64)D0 * sz;
    const int x0 = blockIdx.x * NCC_TX, y0 = blockIdx.y * NCC_TY;
    const int bc = blockIdx.z / p.nzchunks, zc = blockIdx.z % p.nzchunks;
    const int z_start = zc * p.zchunk, z_end = min(D0, z_start + p.zchunk);
    const float *in[3] = {p.in0 + (i64)bc * S, p.in1 + (i64)bc * S, FWD ? nullptr : p.in2 + (i64)bc * S};

    // phase-2 ownership: column x, output rows 2*yp and 2*yp+1
    const int tx = threadIdx.x % NCC_TX, yp = threadIdx.x / NCC_TX;
    const int gx = x0 + tx, gy = y0 + 2 * yp;
    float ring[W][NQ][2];
#pragma unroll
    for (int s = 0; s < W; ++s)
#pragma unroll
        for (int q = 0; q < NQ; ++q) ring[s][q][0] = ring[s][q][1] = 0.0f;

    float cc_acc = 0.0f;
    const float gl = (!FWD && p.gloss) ? __ldg(p.gloss) : 1.0f;
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
                float acc[NQ][NCC_XS];
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int j = 0; j < NCC_XS; ++j) acc[q][j] = 0.0f;
                if (plane_ok && yy >= 0 && yy < D1) {
                    const i64 rowoff = (i64)zin * sz + (i64)yy * sy;
                    const int xb = x0 + sg * NCC_XS - R;
                    float v[NIN][NCC_XS + 2 * R];
#pragma unroll
                    for (int n = 0; n < NIN; ++n)
#pragma unroll
                        for (int j = 0; j < NCC_XS + 2 * R; ++j) {
                            const int xx = xb + j;
                            v[n][j] = (xx >= 0 && xx < D2) ? __ldg(in[n] + rowoff + xx) : 0.0f;
                        }
#pragma unroll
                    for (int j = 0; j < NCC_XS; ++j)
#pragma unroll
                        for (int t = 0; t < W; ++t) {
                            if (FWD) {
                                const float a = v[0][j + t], b = v[1][j + t];
                                acc[0][j] += a;
                                acc[1][j] += b;
                                acc[2][j] = __fadd_rn(acc[2][j], __fmul_rn(a, a));
                                acc[3][j] = __fadd_rn(acc[3][j], __fmul_rn(b, b));
                                acc[4 % NQ][j] = __fadd_rn(acc[4 % NQ][j], __fmul_rn(a, b));
                            } else {
                                acc[0][j] += v[0][j + t];
                                acc[1][j] += v[1][j + t];
                                acc[2][j] += v[2 % NIN][j + t];
                            }
                        }
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    *reinterpret_cast<float4 *>(&X[q][r][sg * NCC_XS]) =
                        make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
            }
            __syncthreads();
            // ---------------- phase 2: y-sums (two rows) + z ring
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                float core = 0.0f;
#pragma unroll
                for (int t = 1; t < W; ++t) core += X[q][2 * yp + t][tx];
                ring[s][q][0] = X[q][2 * yp][tx] + core;
                ring[s][q][1] = core + X[q][2 * yp + W][tx];
            }
            __syncthreads();
            const int zout = zin - R;
            if (zout >= z_start && zout < z_end && gx < D2) {
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    if (gy + o >= D1) continue;
                    float sum[NQ];
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        float t = 0.0f;
#pragma unroll
                        for (int u = 0; u < W; ++u) t += ring[u][q][o];
                        sum[q] = t;
                    }
                    const i64 off = (i64)bc * S + (i64)zout * sz + (i64)(gy + o) * sy + gx;
                    if (FWD) {
                        const float sI = sum[0], sJ = sum[1], sII = sum[2], sJJ = sum[3], sIJ = sum[4 % NQ];
                        const float Wf = p.Wf;
                        const float uI = __fdiv_rn(sI, Wf), uJ = __fdiv_rn(sJ, Wf);
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
                        cc_acc += __fdiv_rn(c2, Dn);
                        if (p.o0) {
                            const float a = __fdiv_rn(2.0f * cross, Dn);
                            const float c = -(c2 * Iv) / (Dn * Dn);
                            p.o0[off] = a;
                            p.o1[off] = -(a * sI) / Wf - (2.0f * c * sJ) / Wf;
                            p.o2[off] = c;
                        }
                    } else {
                        const float Iv = __ldg(p.I + off), Jv = __ldg(p.J + off);
                        p.o0[off] = (gl * p.k) * (Iv * sum[0] + sum[1] + 2.0f * Jv * sum[2 % NQ]);
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

template <bool FWD>
static int ncc_launch(const NccParams &p, const NccGrid &g, int win, cudaStream_t st)
{
    switch (win) {
        case 3: ncc_box_kernel<3, FWD><<<g.grid, NCC_THREADS, 0, st>>>(p); break;
        case 5: ncc_box_kernel<5, FWD><<<g.grid, NCC_THREADS, 0, st>>>(p); break;
        case 7: ncc_box_kernel<7, FWD><<<g.grid, NCC_THREADS, 0, st>>>(p); break;
        case 9: ncc_box_kernel<9, FWD><<<g.grid, NCC_THREADS, 0, st>>>(p); break;
        case 11: ncc_box_kernel<11, FWD><<<g.grid, NCC_THREADS, 0, st>>>(p); break;
        default: return PULPO_ERR_UNSUPPORTED;
    }
    return launch_status();
}

}  // namespace pulpo

using namespace pulpo;

extern "C" size_t pulpo_ncc_ws_bytes(int B, int C, int D0, int D1, int D2)
{
    // worst case over the supported windows: the grid is largest for the smallest window
    NccGrid g = ncc_grid(B * C, D0, D1, D2, 3);
    size_t ctas = (size_t)g.grid.x * g.grid.y * g.grid.z;
    return 16 + sizeof(double) * ctas;
}

extern "C" int pulpo_ncc_fwd(const float *pred, const float *target, float *loss, float *abc, void *ws,
                             size_t ws_bytes, int win, float gamma, int B, int C, int D0, int D1, int D2,
                             pulpo_stream_t stream)
{
    PULPO_REQUIRE(pred && target && loss && ws, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 > 0 && D1 > 0 && D2 > 0, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(win >= 3 && win <= 11 && (win & 1), PULPO_ERR_UNSUPPORTED);
    NccGrid g = ncc_grid(B * C, D0, D1, D2, win);
    PULPO_REQUIRE(g.grid.y <= 65535 && g.grid.z <= 65535, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(ws_bytes >= 16 + sizeof(double) * (size_t)g.grid.x * g.grid.y * g.grid.z, PULPO_ERR_WORKSPACE);
    const i64 n = (i64)B * C * D0 * D1 * D2;
    NccParams p{};
    p.in0 = target; p.in1 = pred; p.in2 = nullptr;
    p.o0 = abc; p.o1 = abc ? abc + n : nullptr; p.o2 = abc ? abc + 2 * n : nullptr;
    p.loss = loss; p.ws = (ReduceWs *)ws;
    p.loss_scale = -(double)gamma / (double)B;
    p.Wf = (float)(win * win * win);
    p.BC = B * C; p.D0 = D0; p.D1 = D1; p.D2 = D2; p.zchunk = g.zchunk; p.nzchunks = g.nzchunks;
    return ncc_launch<true>(p, g, win, (cudaStream_t)stream);
}

extern "C" int pulpo_ncc_bwd(const float *abc, const float *pred, const float *target, const float *gloss,
                             float *gpred, int win, float gamma, int B, int C, int D0, int D1, int D2,
                             pulpo_stream_t stream)
{
    PULPO_REQUIRE(abc && pred && target && gpred, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && C > 0 && D0 > 0 && D1 > 0 && D2 > 0, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(win >= 3 && win <= 11 && (win & 1), PULPO_ERR_UNSUPPORTED);
    NccGrid g = ncc_grid(B * C, D0, D1, D2, win);
    PULPO_REQUIRE(g.grid.y <= 65535 && g.grid.z <= 65535, PULPO_ERR_INVALID_SHAPE);
    const i64 n = (i64)B * C * D0 * D1 * D2;
    NccParams p{};
    p.in0 = abc; p.in1 = abc + n; p.in2 = abc + 2 * n;
    p.I = target; p.J = pred; p.o0 = gpred; p.gloss = gloss;
    p.k = -gamma / (float)B;
    p.Wf = (float)(win * win * win);
    p.BC = B * C; p.D0 = D0; p.D1 = D1; p.D2 = D2; p.zchunk = g.zchunk; p.nzchunks = g.nzchunks;
    return ncc_launch<false>(p, g, win, (cudaStream_t)stream);
}
