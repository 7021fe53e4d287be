// ncc_tma.cu -- windowed local NCC (reference: NCC_loss, src/losses.py:85-135) for volumes large
// enough to be worth a persistent, TMA-fed kernel (D2 % 4 == 0, D2 >= 44, win <= 9); ncc.cu keeps the
// generic path.  Same arithmetic as ncc.cu: five win^3 box sums as separable direct sums (no
// running-sum subtraction), closed-form backward through the saved coefficient volumes.
//
// B200 design.  The box sums are bound by the SM's shared-memory pipe (128 B/cycle) and by
// instruction issue, not by HBM (8 B/voxel), so the kernel is organised around shared-memory bytes
// per voxel:
//   * persistent CTAs walk a contiguous range of the linearised (tile column, z) space
//     (all CTAs get the same number of planes whatever the volume size; a range that crosses a
//     column boundary is processed as two segments).  Tile = 16 x 32 in (D1, D2), two CTAs per SM
//     (the 2 x 45-register z rings per thread cap the SM at 16 warps; independent CTAs keep the
//     pipes busy across each other's per-plane barrier); a CTA marches
//     along D0 with the W-plane z window as a register ring, so the 2R halo planes are paid once
//     per ~45 planes instead of once per 20;
//   * input planes (tile + R halo) are staged by TMA (cp.async.bulk.tensor.4d, 3-stage mbarrier
//     pipeline): the hardware zero-fills outside the volume, which IS conv3d's zero padding, and
//     global-load latency never reaches the math warps;
//   * pass 1 (y sums, from the staged plane): an item = one column x 4 output rows; its 4+2R staged
//     values per input are scalar LDS (conflict-free: lanes = consecutive columns) that land in
//     (I,J) register pairs, so products and every later sum are packed FMUL2 / FADD2 with no
//     register shuffling; results go to a double-buffered shared array (one __syncthreads per plane);
//   * pass 2 (x sums) + z ring + epilogue: a thread owns two adjacent columns of one row (shared
//     window core, 128-bit LDS of two (I,J) pairs), keeps the W-plane z window of its two outputs as
//     a register ring, and writes 64-bit results.
// Algorithmic bytes: fwd 8 B/voxel (+12 when saving a,b,c), bwd 12 B/voxel (+12 reading a,b,c).
#include <type_traits>

#include <cuda.h>   // CUtensorMap and its enums; cuTensorMapEncodeTiled is looked up at run time

#include "ncc_common.cuh"

namespace pulpo {

constexpr int NT_TX = 32;         // tile width  (D2)
constexpr int NT_TY = 16;         // tile height (D1)
constexpr int NT_THREADS = 256;   // pass 2: thread (xp, ty) owns columns 2*xp, 2*xp+1 of row ty
constexpr int NT_CTAS_PER_SM = 2; // two independent CTAs per SM: one computes while the other sits at its plane barrier
constexpr int NT_BOXW = 44;       // staged row: PADL + TX + R (R <= 4), padded
constexpr int NT_PADL = 4;        // the box starts at x0 - 4: TMA needs a 16-byte aligned start along the inner dimension
constexpr int NT_STAGES = 3;
constexpr int NT_RY = 4;          // output rows per pass-1 work item
constexpr int NT_YC = NT_TX + 2 * NT_PADL;   // columns of the y-summed arrays (pitch: 40 float2 = 320 B)

// compile-time loop with early exit: f(integral_constant<int, I>) for I = 0 .. N-1 until it returns false
template <int I, int N, typename F>
__device__ __forceinline__ void static_for_impl(F &f)
{
    if constexpr (I < N) {
        if (f(std::integral_constant<int, I>{})) static_for_impl<I + 1, N>(f);
    }
}
template <int N, typename F>
__device__ __forceinline__ void static_for(F &&f)
{
    static_for_impl<0, N>(f);
}

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned int bar, unsigned int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned int bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned int bar, unsigned int parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
// one [1, 1, rows, NT_BOXW] box of a [BC, D0, D1, D2] fp32 tensor -> shared memory; coordinates may lie
// outside the tensor (negative or too large): those elements arrive as zeros
__device__ __forceinline__ void tma_load_box(unsigned int dst, const CUtensorMap *tm, unsigned int bar, int x, int y,
                                             int z, int bc)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"((unsigned long long)tm), "r"(bar), "r"(x), "r"(y), "r"(z), "r"(bc)
        : "memory");
}

template <int W, bool FWD>
struct NccTmaSmem {
    static constexpr int R = W / 2;
    static constexpr int ROWS = NT_TY + 2 * R;                              // staged rows
    static constexpr int NIN = FWD ? 2 : 3;
    static constexpr int RAW_FLOATS = ((ROWS * NT_BOXW + 31) / 32) * 32;   // one input plane, 128-byte aligned slots
    static constexpr int RAW_BYTES = NT_STAGES * NIN * RAW_FLOATS * 4;
    static constexpr int YA_FLOATS = NT_TY * NT_YC * 2, YC_FLOATS = NT_TY * NT_YC;
    static constexpr int YS_FLOATS = YA_FLOATS * (FWD ? 2 : 1) + YC_FLOATS;   // one buffer: A, [B], C
    static constexpr int BAR_OFF = RAW_BYTES + 2 * YS_FLOATS * 4;
    static constexpr int RED_OFF = BAR_OFF + 64;
    static constexpr int BYTES = RED_OFF + 32 * 8;
};

template <int W, bool FWD>
__global__ void __launch_bounds__(NT_THREADS, NT_CTAS_PER_SM)
ncc_tma_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
               const __grid_constant__ CUtensorMap tm2, const NccTmaParams p)
{
    using L = NccTmaSmem<W, FWD>;
    constexpr int R = L::R, ROWS = L::ROWS, NIN = L::NIN;
    constexpr int XOFF = NT_PADL - R;            // first staged / y-summed column any window needs
    constexpr int NC = NT_TX + 2 * R;            // y-summed columns that are needed: XOFF .. XOFF + NC - 1
    constexpr int NVY = NT_RY + 2 * R;           // staged rows per pass-1 item
    constexpr int ITEMS = NC * (NT_TY / NT_RY);
    constexpr int NP = 2 * R + 2;                // y-summed columns per pass-2 thread (two adjacent outputs)
    constexpr unsigned int STAGE_BYTES = NIN * ROWS * NT_BOXW * 4;
    extern __shared__ __align__(128) unsigned char smem[];
    float *raw = reinterpret_cast<float *>(smem);                        // [stage][input][ROWS][BOXW]
    float *ys = reinterpret_cast<float *>(smem + L::RAW_BYTES);          // [2][YS_FLOATS]
    double *red = reinterpret_cast<double *>(smem + L::RED_OFF);
    const unsigned int raw_s = smem_u32(smem), bar_s = raw_s + L::BAR_OFF;

    const int tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NT_STAGES; ++s) mbar_init(bar_s + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int D0 = p.D0, D1 = p.D1, D2 = p.D2;
    const int sy = D2, sz = D1 * D2;
    const i64 S = (i64)D0 * sz;
    // pass-2 ownership: columns 2*xp, 2*xp+1 of row ty
    const int xp = tid % (NT_TX / 2), ty = tid / (NT_TX / 2);
    // pass-1 ownership: column fastest across lanes (conflict-free 32-bit loads / 64-bit stores)
    const bool has_item = tid < ITEMS;
    const int yc = tid % NC, yrg = tid / NC;

    // this CTA's share of the linearised (column, z) space
    const long long T = p.total_planes;
    long long q = (T * blockIdx.x) / gridDim.x;
    long long q_end = (T * (blockIdx.x + 1)) / gridDim.x;
    if (p.nbounds > 0) {
        // cost-balanced mode: every segment a CTA starts costs its 2R halo planes; the host split the linear space so
        // that planes + halos are even across CTAs (a plain even split leaves the CTAs that cross a column boundary
        // with two halos, the aligned grid leaves CTA slots empty)
        q = p.bounds[blockIdx.x];
        q_end = p.bounds[blockIdx.x + 1];
    } else if (p.zchunks > 0) {
        // aligned mode: CTA = (column, z chunk) with the same chunk boundaries in every column and neighbouring
        // columns on neighbouring CTAs, so the CTAs that share halo rows read them at about the same time (L2 hits)
        const int ncols = (int)(T / D0);
        const int colx = blockIdx.x % ncols, ch = blockIdx.x / ncols;
        const int zc = (D0 + p.zchunks - 1) / p.zchunks;
        const int za = ch * zc, zb = min(D0, za + zc);
        q = (long long)colx * D0 + za;
        q_end = (long long)colx * D0 + (zb > za ? zb : za);
    }

    float cc_acc = 0.0f;
    const float gk = FWD ? 0.0f : ((p.gloss ? __ldg(p.gloss) : 1.0f) * p.k);
    unsigned int gp = 0;   // planes staged so far by this CTA (stage = gp % NT_STAGES, parity = (gp / NT_STAGES) & 1)

    while (q < q_end) {
        const int col = (int)(q / D0);
        const int z_start = (int)(q - (long long)col * D0);
        const long long left = q_end - q;
        const int z_end = (left < (long long)(D0 - z_start)) ? z_start + (int)left : D0;
        q += z_end - z_start;
        const int xtile = col % p.xt, ytile = (col / p.xt) % p.yt, bc = col / (p.xt * p.yt);
        const int x0 = xtile * NT_TX, y0 = ytile * NT_TY;
        const int gx = x0 + 2 * xp, gy = y0 + ty;
        const bool live = (gx < D2) && (gy < D1);     // D2 is even: gx < D2 implies gx + 1 < D2
        const i64 obase = (i64)bc * S + (i64)gy * sy + gx;
        const int nplanes = (z_end - z_start) + 2 * R;

        __syncthreads();   // every thread is done with the previous segment's shared memory
        if (tid == 0) {
            const int n0 = nplanes < NT_STAGES ? nplanes : NT_STAGES;
            for (int i = 0; i < n0; ++i) {
                const unsigned int st = (gp + i) % NT_STAGES;
                const unsigned int bar = bar_s + 8 * st, dst = raw_s + st * NIN * L::RAW_FLOATS * 4;
                mbar_expect_tx(bar, STAGE_BYTES);
                tma_load_box(dst, &tm0, bar, x0 - NT_PADL, y0 - R, z_start - R + i, bc);
                tma_load_box(dst + L::RAW_FLOATS * 4, &tm1, bar, x0 - NT_PADL, y0 - R, z_start - R + i, bc);
                if (!FWD) tma_load_box(dst + 2 * L::RAW_FLOATS * 4, &tm2, bar, x0 - NT_PADL, y0 - R, z_start - R + i, bc);
            }
        }

        float2 rA[W][2], rB[FWD ? W : 1][2];
        float rC[W][2];
#pragma unroll
        for (int s = 0; s < W; ++s) {
            rA[s][0] = rA[s][1] = make_float2(0.f, 0.f);
            if (FWD) rB[s % (FWD ? W : 1)][0] = rB[s % (FWD ? W : 1)][1] = make_float2(0.f, 0.f);
            rC[s][0] = rC[s][1] = 0.0f;
        }

        if constexpr (FWD) {
        // forward: pass 1 and pass 2 of a plane back to back across the plane barrier (its two 45-register z rings leave
        // no room to overlap the passes of consecutive planes: the pipelined order below spills, 81 -> 94 us; keeping the
        // z windows as sums of three planes -- 4 instead of 8 additions, 8 instead of 9 registers per output -- was
        // measured too: no gain in the backward, more spills in the forward)
        for (int p0 = 0; p0 < nplanes; p0 += W) {
#pragma unroll
            for (int s = 0; s < W; ++s) {
                const int pl = p0 + s;
                if (pl >= nplanes) break;
                const unsigned int st = (gp + pl) % NT_STAGES, par = ((gp + pl) / NT_STAGES) & 1u;
                float *ybuf = ys + (pl & 1) * L::YS_FLOATS;
                float2 *YA = reinterpret_cast<float2 *>(ybuf);
                float2 *YB = reinterpret_cast<float2 *>(ybuf + L::YA_FLOATS);
                float *YC = ybuf + L::YA_FLOATS * (FWD ? 2 : 1);
                const int zout = z_start + pl - 2 * R;
                // backward epilogue operands: requested now, used after both passes
                float2 Iv = make_float2(0.f, 0.f), Jv = Iv;
                if (!FWD && pl >= 2 * R && live) {
                    Iv = __ldg(reinterpret_cast<const float2 *>(p.I + obase + (i64)zout * sz));
                    Jv = __ldg(reinterpret_cast<const float2 *>(p.J + obase + (i64)zout * sz));
                }
                mbar_wait(bar_s + 8 * st, par);
                // ---------------- pass 1: y sums of the staged plane (products formed in registers).  Scalar loads
                // let (I, J) land in one register pair, so everything downstream is packed FADD2 / FMUL2.
                if (has_item) {
                    // the forward lives at the register limit (two 45-register z rings per thread): the three y sums
                    // are formed one after the other from the staged (I, J) pairs -- sum, store, then the next product
                    // in the same registers -- and pass 2 below handles one quantity at a time as well.  Forming all
                    // products / loading all staged columns at once spilled ~70 registers: 98 -> 81 us without.
                    const float *r0 = raw + st * NIN * L::RAW_FLOATS + (NT_RY * yrg) * NT_BOXW + XOFF + yc;
                    const int idx0 = (NT_RY * yrg) * NT_YC + XOFF + yc;
                    float2 a[NVY];
#pragma unroll
                    for (int j = 0; j < NVY; ++j) a[j] = make_float2(r0[j * NT_BOXW], r0[L::RAW_FLOATS + j * NT_BOXW]);
                    {
                        float2 oA[4];
                        xsum4<W>(a, oA);
#pragma unroll
                        for (int o = 0; o < NT_RY; ++o) YA[idx0 + o * NT_YC] = oA[o];
                    }
                    asm volatile("" ::: "memory");
                    {
                        float c[NVY], oC[4];
#pragma unroll
                        for (int j = 0; j < NVY; ++j) c[j] = __fmul_rn(a[j].x, a[j].y);
                        xsum4<W>(c, oC);
#pragma unroll
                        for (int o = 0; o < NT_RY; ++o) YC[idx0 + o * NT_YC] = oC[o];
                    }
                    asm volatile("" ::: "memory");
                    {
                        float2 oB[4];
#pragma unroll
                        for (int j = 0; j < NVY; ++j) a[j] = __fmul2_rn(a[j], a[j]);
                        xsum4<W>(a, oB);
#pragma unroll
                        for (int o = 0; o < NT_RY; ++o) YB[idx0 + o * NT_YC] = oB[o];
                    }
                }
                __syncthreads();
                // the staged plane has been consumed by every thread: refill its slot
                if (tid == 0 && pl + NT_STAGES < nplanes) {
                    const unsigned int bar = bar_s + 8 * st, dst = raw_s + st * NIN * L::RAW_FLOATS * 4;
                    const int zin = z_start - R + pl + NT_STAGES;
                    mbar_expect_tx(bar, STAGE_BYTES);
                    tma_load_box(dst, &tm0, bar, x0 - NT_PADL, y0 - R, zin, bc);
                    tma_load_box(dst + L::RAW_FLOATS * 4, &tm1, bar, x0 - NT_PADL, y0 - R, zin, bc);
                    if (!FWD) tma_load_box(dst + 2 * L::RAW_FLOATS * 4, &tm2, bar, x0 - NT_PADL, y0 - R, zin, bc);
                }
                // ---------------- pass 2: x sums for two adjacent outputs (shared window core) into ring slot s, one
                // quantity after the other (A, then C, then B: at most 10 staged columns in registers next to the rings),
                // each followed at once by its z window sum (computed on the halo planes too: guarding it spills again)
                const int cb = ty * NT_YC + 2 * xp + XOFF;   // first y-summed column of this thread's windows
                float2 sA[2], sB[2];
                float sC[2];
                {
                    float2 PA[NP];
#pragma unroll
                    for (int t = 0; t < NP / 2; ++t) {
                        if ((XOFF & 1) == 0) {
                            const float4 qa = *reinterpret_cast<const float4 *>(YA + cb + 2 * t);
                            PA[2 * t] = make_float2(qa.x, qa.y); PA[2 * t + 1] = make_float2(qa.z, qa.w);
                        } else {
                            PA[2 * t] = YA[cb + 2 * t]; PA[2 * t + 1] = YA[cb + 2 * t + 1];
                        }
                    }
                    float2 coreA = PA[1];
#pragma unroll
                    for (int t = 2; t < W; ++t) coreA = add2(coreA, PA[t]);
                    rA[s][0] = add2(PA[0], coreA);
                    rA[s][1] = add2(coreA, PA[W]);
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        sA[o] = rA[0][o];
#pragma unroll
                        for (int u = 1; u < W; ++u) sA[o] = add2(sA[o], rA[u][o]);
                    }
                }
                asm volatile("" ::: "memory");
                {
                    float PC[NP];
#pragma unroll
                    for (int t = 0; t < NP / 2; ++t) {
                        if ((XOFF & 1) == 0) {
                            const float2 qc = *reinterpret_cast<const float2 *>(YC + cb + 2 * t);
                            PC[2 * t] = qc.x; PC[2 * t + 1] = qc.y;
                        } else {
                            PC[2 * t] = YC[cb + 2 * t]; PC[2 * t + 1] = YC[cb + 2 * t + 1];
                        }
                    }
                    float coreC = PC[1];
#pragma unroll
                    for (int t = 2; t < W; ++t) coreC = add2(coreC, PC[t]);
                    rC[s][0] = add2(PC[0], coreC);
                    rC[s][1] = add2(coreC, PC[W]);
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        sC[o] = rC[0][o];
#pragma unroll
                        for (int u = 1; u < W; ++u) sC[o] = add2(sC[o], rC[u][o]);
                    }
                }
                asm volatile("" ::: "memory");
                {
                    float2 PB[NP];
#pragma unroll
                    for (int t = 0; t < NP / 2; ++t) {
                        if ((XOFF & 1) == 0) {
                            const float4 qb = *reinterpret_cast<const float4 *>(YB + cb + 2 * t);
                            PB[2 * t] = make_float2(qb.x, qb.y); PB[2 * t + 1] = make_float2(qb.z, qb.w);
                        } else {
                            PB[2 * t] = YB[cb + 2 * t]; PB[2 * t + 1] = YB[cb + 2 * t + 1];
                        }
                    }
                    float2 coreB = PB[1];
#pragma unroll
                    for (int t = 2; t < W; ++t) coreB = add2(coreB, PB[t]);
                    rB[s][0] = add2(PB[0], coreB);
                    rB[s][1] = add2(coreB, PB[W]);
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        sB[o] = rB[0][o];
#pragma unroll
                        for (int u = 1; u < W; ++u) sB[o] = add2(sB[o], rB[u][o]);
                    }
                }
                // ---------------- epilogue
                if (pl >= 2 * R && live) {
                    const i64 off = obase + (i64)zout * sz;
                    float ra[2], rb[2], rc[2];
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        if (p.o0) {
                            const NccPoint r = ncc_point<true>(sA[o].x, sA[o].y, sB[o].x, sB[o].y, sC[o], p.Wf, p.rcpW);
                            cc_acc += r.cc;
                            ra[o] = r.a; rb[o] = r.b; rc[o] = r.c;
                        } else {
                            cc_acc += ncc_point<false>(sA[o].x, sA[o].y, sB[o].x, sB[o].y, sC[o], p.Wf, p.rcpW).cc;
                        }
                    }
                    if (p.o0) {
                        *reinterpret_cast<float2 *>(p.o0 + off) = make_float2(ra[0], ra[1]);
                        *reinterpret_cast<float2 *>(p.o1 + off) = make_float2(rb[0], rb[1]);
                        *reinterpret_cast<float2 *>(p.o2 + off) = make_float2(rc[0], rc[1]);
                    }
                }
            }
        }
        } else {
        // backward: software-pipelined across the plane barrier (79 -> 70 us at 160x192x224)
        // Pass 2 (x sums into ring slot SLOT) + z window + epilogue of plane `pl`, from the y sums in buffer pl & 1.
        // Called one barrier interval AFTER the plane's pass 1: in every interval a thread runs pass 1 of plane pl and
        // pass 2 of plane pl - 1, two independent instruction streams, instead of the two passes of one plane back to
        // back across the barrier -- the kernel is bound by the length of the per-plane dependency chain (measured:
        // skipping the arithmetic of all-zero planes changed nothing), not by throughput.
        constexpr bool PIPE = true;
        auto pass2 = [&](auto slot_c, const int pl, const float2 Iv, const float2 Jv) {
            constexpr int SLOT = decltype(slot_c)::value;
            float *ybuf = ys + (pl & 1) * L::YS_FLOATS;
            const float2 *YA = reinterpret_cast<const float2 *>(ybuf);
            const float2 *YB = reinterpret_cast<const float2 *>(ybuf + L::YA_FLOATS);
            const float *YC = ybuf + L::YA_FLOATS * (FWD ? 2 : 1);
            const int zout = z_start + pl - 2 * R;
            {
                float2 PA[NP], PB[NP];
                float PC[NP];
                const int cb = ty * NT_YC + 2 * xp + XOFF;   // first y-summed column of this thread's windows
                if ((XOFF & 1) == 0) {
#pragma unroll
                    for (int t = 0; t < NP / 2; ++t) {
                        const float4 qa = *reinterpret_cast<const float4 *>(YA + cb + 2 * t);
                        PA[2 * t] = make_float2(qa.x, qa.y); PA[2 * t + 1] = make_float2(qa.z, qa.w);
                        if (FWD) {
                            const float4 qb = *reinterpret_cast<const float4 *>(YB + cb + 2 * t);
                            PB[2 * t] = make_float2(qb.x, qb.y); PB[2 * t + 1] = make_float2(qb.z, qb.w);
                        }
                        const float2 qc = *reinterpret_cast<const float2 *>(YC + cb + 2 * t);
                        PC[2 * t] = qc.x; PC[2 * t + 1] = qc.y;
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < NP; ++t) {
                        PA[t] = YA[cb + t];
                        if (FWD) PB[t] = YB[cb + t];
                        PC[t] = YC[cb + t];
                    }
                }
                float2 coreA = PA[1];
                float coreC = PC[1];
#pragma unroll
                for (int t = 2; t < W; ++t) {
                    coreA = add2(coreA, PA[t]);
                    coreC = add2(coreC, PC[t]);
                }
                rA[SLOT][0] = add2(PA[0], coreA);
                rA[SLOT][1] = add2(coreA, PA[W]);
                rC[SLOT][0] = add2(PC[0], coreC);
                rC[SLOT][1] = add2(coreC, PC[W]);
                if (FWD) {
                    float2 coreB = PB[1];
#pragma unroll
                    for (int t = 2; t < W; ++t) coreB = add2(coreB, PB[t]);
                    rB[FWD ? SLOT : 0][0] = add2(PB[0], coreB);
                    rB[FWD ? SLOT : 0][1] = add2(coreB, PB[W]);
                }
            }
            // ---------------- z window + epilogue
            if (pl >= 2 * R && live) {
                const i64 off = obase + (i64)zout * sz;
                float ra[2], rb[2], rc[2];
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    float2 sA = rA[0][o], sB = FWD ? rB[0][o] : make_float2(0.f, 0.f);
                    float sC = rC[0][o];
#pragma unroll
                    for (int u = 1; u < W; ++u) {
                        sA = add2(sA, rA[u][o]);
                        if (FWD) sB = add2(sB, rB[u % (FWD ? W : 1)][o]);
                        sC = add2(sC, rC[u][o]);
                    }
                    if (FWD) {
                        if (p.o0) {
                            const NccPoint r = ncc_point<true>(sA.x, sA.y, sB.x, sB.y, sC, p.Wf, p.rcpW);
                            cc_acc += r.cc;
                            ra[o] = r.a; rb[o] = r.b; rc[o] = r.c;
                        } else {
                            cc_acc += ncc_point<false>(sA.x, sA.y, sB.x, sB.y, sC, p.Wf, p.rcpW).cc;
                        }
                    } else {
                        const float iv = o ? Iv.y : Iv.x, jv = o ? Jv.y : Jv.x;
                        ra[o] = gk * (iv * sA.x + sA.y + 2.0f * jv * sC);
                    }
                }
                if (FWD) {
                    if (p.o0) {
                        *reinterpret_cast<float2 *>(p.o0 + off) = make_float2(ra[0], ra[1]);
                        *reinterpret_cast<float2 *>(p.o1 + off) = make_float2(rb[0], rb[1]);
                        *reinterpret_cast<float2 *>(p.o2 + off) = make_float2(rc[0], rc[1]);
                    }
                } else {
                    *reinterpret_cast<float2 *>(p.o0 + off) = make_float2(ra[0], ra[1]);
                }
            }
        };

        float2 Iv_prev = make_float2(0.f, 0.f), Jv_prev = Iv_prev;   // backward epilogue operands of the plane whose pass 2 is pending
        for (int p0 = 0; p0 <= nplanes; p0 += W) {
            static_for<W>([&](auto s_c) -> bool {      // ring slot s is a compile-time constant: the rings stay in registers
                constexpr int s = decltype(s_c)::value;
                const int pl = p0 + s;
                if (pl > nplanes) return false;
                const bool have = pl < nplanes;       // the last turn only drains pass 2 of the final plane
                const unsigned int st = (gp + pl) % NT_STAGES, par = ((gp + pl) / NT_STAGES) & 1u;
                float *ybuf = ys + (pl & 1) * L::YS_FLOATS;
                float2 *YA = reinterpret_cast<float2 *>(ybuf);
                float2 *YB = reinterpret_cast<float2 *>(ybuf + L::YA_FLOATS);
                float *YC = ybuf + L::YA_FLOATS * (FWD ? 2 : 1);
                const int zout = z_start + pl - 2 * R;
                // backward epilogue operands: requested now, used one interval later
                float2 Iv = make_float2(0.f, 0.f), Jv = Iv;
                if (!FWD && have && pl >= 2 * R && live) {
                    Iv = __ldg(reinterpret_cast<const float2 *>(p.I + obase + (i64)zout * sz));
                    Jv = __ldg(reinterpret_cast<const float2 *>(p.J + obase + (i64)zout * sz));
                }
                if (have) mbar_wait(bar_s + 8 * st, par);
                // ---------------- pass 1: y sums of the staged plane (products formed in registers).  Scalar loads
                // let (I, J) land in one register pair, so everything downstream is packed FADD2 / FMUL2.
                if (has_item && have) {
                    const float *r0 = raw + st * NIN * L::RAW_FLOATS + (NT_RY * yrg) * NT_BOXW + XOFF + yc;
                    float2 a[NVY], b[NVY];
                    float c[NVY];
#pragma unroll
                    for (int j = 0; j < NVY; ++j) {
                        a[j] = make_float2(r0[j * NT_BOXW], r0[L::RAW_FLOATS + j * NT_BOXW]);
                        if (FWD) {
                            b[j] = __fmul2_rn(a[j], a[j]);
                            c[j] = __fmul_rn(a[j].x, a[j].y);
                        } else {
                            b[j] = make_float2(0.f, 0.f);
                            c[j] = r0[2 * L::RAW_FLOATS + j * NT_BOXW];
                        }
                    }
                    float2 oA[4], oB[4];
                    float oC[4];
                    xsum4<W>(a, oA);
                    if (FWD) xsum4<W>(b, oB);
                    xsum4<W>(c, oC);
#pragma unroll
                    for (int o = 0; o < NT_RY; ++o) {
                        const int idx = (NT_RY * yrg + o) * NT_YC + XOFF + yc;
                        YA[idx] = oA[o];
                        if (FWD) YB[idx] = oB[o];
                        YC[idx] = oC[o];
                    }
                }
                // ---------------- pass 2 + epilogue of the PREVIOUS plane (its y sums were published by the last barrier)
                if (PIPE) {
                    if (pl > 0) pass2(std::integral_constant<int, (s + W - 1) % W>{}, pl - 1, Iv_prev, Jv_prev);
                    Iv_prev = Iv; Jv_prev = Jv;
                }
                __syncthreads();
                // the staged plane has been consumed by every thread: refill its slot
                if (tid == 0 && have && pl + NT_STAGES < nplanes) {
                    const unsigned int bar = bar_s + 8 * st, dst = raw_s + st * NIN * L::RAW_FLOATS * 4;
                    const int zin = z_start - R + pl + NT_STAGES;
                    mbar_expect_tx(bar, STAGE_BYTES);
                    tma_load_box(dst, &tm0, bar, x0 - NT_PADL, y0 - R, zin, bc);
                    tma_load_box(dst + L::RAW_FLOATS * 4, &tm1, bar, x0 - NT_PADL, y0 - R, zin, bc);
                    if (!FWD) tma_load_box(dst + 2 * L::RAW_FLOATS * 4, &tm2, bar, x0 - NT_PADL, y0 - R, zin, bc);
                }
                // not pipelined: the two passes of the same plane, back to back across the barrier
                if (!PIPE && have) pass2(std::integral_constant<int, s>{}, pl, Iv, Jv);
                return true;
            });
        }
        }
        gp += (unsigned int)nplanes;
    }
    if (FWD) {
        double bt = block_sum((double)cc_acc, red);
        grid_reduce_finish(bt, p.ws, p.loss, p.loss_scale, red);
    }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    // resolved once; a plain function pointer, not mutable state that affects results
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess ||
            qr != cudaDriverEntryPointSuccess)
            return nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}

static bool make_map(CUtensorMap *tm, const float *ptr, int BC, int D0, int D1, int D2, int rows)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[4] = {(cuuint64_t)D2, (cuuint64_t)D1, (cuuint64_t)D0, (cuuint64_t)BC};
    cuuint64_t strides[3] = {(cuuint64_t)D2 * 4, (cuuint64_t)D2 * D1 * 4, (cuuint64_t)D2 * D1 * D0 * 4};
    cuuint32_t box[4] = {(cuuint32_t)NT_BOXW, (cuuint32_t)rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)ptr, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {
        // a thread that has only used the runtime API lazily (e.g. torch's autograd worker) may have no
        // context bound yet for a direct driver call: bind the primary context and retry
        cudaFree(0);
        r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS && getenv("PULPO_B200_DEBUG"))
        fprintf(stderr, "libpulpo_b200: cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d] box rows %d\n", (int)r, BC, D0, D1,
                D2, rows);
    return r == CUDA_SUCCESS;
}

bool ncc_tma_eligible(const float *in0, const float *in1, const float *in2, int D0, int D1, int D2, int win)
{
    if (win < 3 || win > 9 || !(win & 1)) return false;   // R <= NT_PADL
    if (D2 % 4 != 0 || D2 < NT_BOXW || D1 < NT_TY / 2 || D0 < 2 * win) return false;   // small volumes: generic kernel
    if (!aligned16(in0) || !aligned16(in1) || (in2 && !aligned16(in2))) return false;
    return encode_fn() != nullptr;
}

template <int W, bool FWD>
static int ncc_tma_launch_w(const CUtensorMap &t0, const CUtensorMap &t1, const CUtensorMap &t2, const NccTmaParams &p,
                            int grid, cudaStream_t st)
{
    using L = NccTmaSmem<W, FWD>;
    static bool configured[64] = {};   // per device; setting the attribute is idempotent
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        if (cudaFuncSetAttribute(ncc_tma_kernel<W, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES) != cudaSuccess)
            return launch_status();
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    ncc_tma_kernel<W, FWD><<<grid, NT_THREADS, L::BYTES, st>>>(t0, t1, t2, p);
    return launch_status();
}

int ncc_tma_grid(int BC, int D0, int D1, int D2, int win)
{
    const long long planes = (long long)BC * ((D2 + NT_TX - 1) / NT_TX) * ((D1 + NT_TY - 1) / NT_TY) * D0;
    int dev = 0, sms = kSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // at least ~2*win planes per CTA so the z halo stays a fraction of the work
    const long long want = planes / (2 * win);
    const long long cap = (long long)sms * NT_CTAS_PER_SM;
    const long long ncols = planes / D0;
    if (ncols <= cap && want >= cap) {   // aligned (column, z chunk) grid
        const long long k = cap / ncols;
        if (k * ncols * 10 >= cap * 8) return (int)(k * ncols);          // keeps >= 80 % of the CTA slots busy
    }
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

// z chunks per column of the aligned grid (0: linearised split)
static int ncc_tma_zchunks(int BC, int D0, int D1, int D2, int grid)
{
    const long long ncols = (long long)BC * ((D2 + NT_TX - 1) / NT_TX) * ((D1 + NT_TY - 1) / NT_TY);
    if (grid % ncols != 0) return 0;
    int dev = 0, sms = kSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long cap = (long long)sms * NT_CTAS_PER_SM;
    return (grid / ncols == cap / ncols && ncols <= cap) ? (int)(grid / ncols) : 0;
}

// Cost-balanced split of the linearised (column, z) space over `grid` CTAs: a CTA pays H = 2R halo planes for every
// segment it starts.  Smallest per-CTA budget T for which a greedy walk needs at most `grid` CTAs (binary search);
// a CTA does not start a segment it could not extend by at least `minseg` planes.
static bool ncc_balanced_bounds(NccTmaParams &p, int grid, int win)
{
    const long long ncols = p.total_planes / p.D0;
    if (grid > NCC_MAX_CTAS || p.total_planes >= (1ll << 31) || grid < 2) return false;
    const int H = 2 * (win / 2), D0 = p.D0, minseg = H;
    auto walk = [&](int T, int *bounds) -> int {
        int cta = 0, budget = T;
        long long q = 0;
        if (bounds) bounds[0] = 0;
        for (long long c = 0; c < ncols; ++c) {
            int remaining = D0;
            while (remaining > 0) {
                if (budget < H + (remaining < minseg ? remaining : minseg)) {
                    ++cta;
                    if (bounds && cta <= NCC_MAX_CTAS) bounds[cta] = (int)q;
                    budget = T;
                }
                int take = budget - H;
                if (take > remaining) take = remaining;
                remaining -= take; q += take; budget -= H + take;
            }
        }
        return cta + 1;
    };
    int lo = H + 1, hi = D0 + H;
    if (walk(hi, nullptr) > grid) return false;
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (walk(mid, nullptr) <= grid) hi = mid; else lo = mid + 1;
    }
    const int used = walk(lo, p.bounds);
    for (int i = used; i <= grid; ++i) p.bounds[i] = (int)p.total_planes;
    p.nbounds = grid;
    return true;
}

static int ncc_tma_dispatch(const CUtensorMap &t0, const CUtensorMap &t1, const CUtensorMap &t2, const NccTmaParams &p, int grid,
                            int win, bool fwd, cudaStream_t st);

// in2 == nullptr selects the forward (inputs: target, pred), else the backward (a, b, c)
int ncc_tma_launch(const float *in0, const float *in1, const float *in2, NccTmaParams p, int win, cudaStream_t st)
{
    const bool fwd = (in2 == nullptr);
    const int rows = NT_TY + 2 * (win / 2);
    CUtensorMap t0, t1, t2;
    if (!make_map(&t0, in0, p.BC, p.D0, p.D1, p.D2, rows) || !make_map(&t1, in1, p.BC, p.D0, p.D1, p.D2, rows))
        return PULPO_ERR_CUDA;
    if (fwd)
        t2 = t1;
    else if (!make_map(&t2, in2, p.BC, p.D0, p.D1, p.D2, rows))
        return PULPO_ERR_CUDA;
    p.xt = (p.D2 + NT_TX - 1) / NT_TX;
    p.yt = (p.D1 + NT_TY - 1) / NT_TY;
    p.total_planes = (long long)p.BC * p.xt * p.yt * p.D0;
    const int grid = ncc_tma_grid(p.BC, p.D0, p.D1, p.D2, win);
    p.zchunks = ncc_tma_zchunks(p.BC, p.D0, p.D1, p.D2, grid);
    p.nbounds = 0;
#ifdef PULPO_NCC_BALANCED
    {
        int dev = 0, sms = kSMs;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int cap = sms * NT_CTAS_PER_SM;
        if (p.total_planes / (2 * win) >= cap && ncc_balanced_bounds(p, cap, win)) {
            p.zchunks = 0;
            return ncc_tma_dispatch(t0, t1, t2, p, cap, win, fwd, st);
        }
    }
#endif
    return ncc_tma_dispatch(t0, t1, t2, p, grid, win, fwd, st);
}

static int ncc_tma_dispatch(const CUtensorMap &t0, const CUtensorMap &t1, const CUtensorMap &t2, const NccTmaParams &p, int grid,
                            int win, bool fwd, cudaStream_t st)
{
#define PULPO_NCC_TMA_CASE(WW)                                                   \
    case WW:                                                                     \
        return fwd ? ncc_tma_launch_w<WW, true>(t0, t1, t2, p, grid, st)         \
                   : ncc_tma_launch_w<WW, false>(t0, t1, t2, p, grid, st);
    switch (win) {
        PULPO_NCC_TMA_CASE(3)
        PULPO_NCC_TMA_CASE(5)
        PULPO_NCC_TMA_CASE(7)
        PULPO_NCC_TMA_CASE(9)
    }
#undef PULPO_NCC_TMA_CASE
    return PULPO_ERR_UNSUPPORTED;
}

}  // namespace pulpo
