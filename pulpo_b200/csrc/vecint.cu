// vecint.cu -- scaling-and-squaring integration (reference: VecInt, src/network_blocks.py:165-177):
//     v <- vec * 2^-nsteps ;  nsteps times:  v <- v + warp(v, v)
//
// B200 design.  The field is both the image and the displacement, so each step is a full
// grid-wide dependency.  All steps run in ONE cooperative launch with grid.sync() (1.25 us measured,
// scripts/micro/sync_bench.cu) between steps instead of 7 launches; the integration states live in
// a channel-interleaved float4 layout ([B,S] x (c0,c1,c2,pad)) so that every trilinear corner is one
// 128-bit gather and every scatter one red.global.add.v4.f32.  The states of one level
// (<= 13.8 MB at 80x96x112) stay resident in the 126 MB L2 between steps.
//
// What bounds it (ncu, profiles/): not HBM and not instruction issue but the SM's L1 load/store
// pipe -- a 128-bit warp gather that touches n cache lines occupies it for ~2n cycles, a vector
// reduction for ~0.6 cycles per lane -- plus the grid barrier.  Hence:
//   * work mapping: a warp owns an 8 (x) by 4 (y) patch (4 lines per aligned gather) and walks a
//     run of z planes, one work item per resident warp (one CTA per SM);
//   * z reuse: the four upper corners of one plane are the four lower corners of the next when the
//     footprint moved by exactly one plane (the usual case for a smooth field) -> 4 gathers per
//     plane instead of 8; the backward mirrors it for the scatter (the four upper-corner
//     contributions are carried in registers and merged into the next plane's) -> 4 reductions;
//   * ordered loads: the next plane's own values are requested before this plane's gathers (volatile
//     asm keeps the order), otherwise every plane pays two serialised L2 round trips;
//   * PULPO_COORD_FAST: sample position in one FMA, FMA interpolation (no index contract here).
// In the exact modes the forward arithmetic is op-for-op the CPU grid sampler's: results are
// bit-identical to torch-CPU (PULPO_COORD_CPU_EXACT).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pulpo {

#ifndef PULPO_VI_PX
#define PULPO_VI_PX 8
#endif
constexpr int VI_PX = PULPO_VI_PX, VI_PY = 32 / VI_PX;   // lanes of a warp: VI_PX along x, VI_PY along y
constexpr int VI_PX_LOG2 = VI_PX == 8 ? 3 : VI_PX == 16 ? 4 : 5;
static_assert(VI_PX == 8 || VI_PX == 16 || VI_PX == 32, "patch width");
// One CTA per SM: a CTA that reaches grid.sync() early polls the barrier with acquire loads,
// and every poll invalidates that SM's L1 (CCTL.IVALL) -- with several CTAs per SM this slowed
// the CTAs still working next to it.  With one CTA per SM nobody is left to disturb.
// Thread counts from a sweep on B200 (config-2 levels, bench kernel table): fewer, longer z runs win
// (more corner reuse / scatter carry per run, no spills): forward 1024 -> 768: 126 -> 118 us,
// backward 768 / 640 / 512 / 448 / 384: 343 / 347 / 322 / 356 / 414 us.
#ifndef PULPO_VI_BWD_PY
#define PULPO_VI_BWD_XYZ 1   // three rotating gradient states (default); -DPULPO_VI_BWD_PY: two (own, scatter) pairs
#endif
#ifndef PULPO_VI_FWD_THREADS
#define PULPO_VI_FWD_THREADS 768
#endif
#ifndef PULPO_VI_BWD_THREADS
#define PULPO_VI_BWD_THREADS 512
#endif
constexpr int VI_FWD_THREADS = PULPO_VI_FWD_THREADS;
constexpr int VI_BWD_THREADS = PULPO_VI_BWD_THREADS;

struct VGeom {
    int B, D0, D1, D2;
    int S;            // voxels per volume (< 2^31)
    unsigned int N;   // B * S
    int unbias;
    int npx, npy;     // patches per row / per plane column
    int zrun, nzrun;  // planes per work item, items per column
    unsigned int items;   // B * npy * npx * nzrun
    FastDiv dnpx, dnpy, dnz;
    AxisConst a0, a1, a2;
};

// One launch integrates several pyramid levels at once (they are independent fields with the same
// step count): a cooperative kernel owns every SM, so per-level launches would serialise and each
// would pay its own grid barriers.  Work items of all levels form one list.
constexpr int VI_MAXL = 6;
struct VLevel {
    const float *in;    // fwd: vec ; bwd: gout            [B,3,D0,D1,D2]
    float *out;         // fwd: integrated field ; bwd: gvec
    float4 *ws;         // fwd: states ; bwd: saved states
    float4 *scr;        // bwd: two (own, scatter) pairs
    unsigned int item0; // first work item of this level in the joint list
    // dataflow synchronisation (see "Step synchronisation" below): tail of the caller's workspace
    unsigned int *rowdone;   // [B * nzrun * npy] steps-completed counters, one per row of x-adjacent work items
    float *ctamax;           // [gridDim.x] per-CTA max |v_0| of this level
    int l1safe;              // every 128-byte line of a state belongs to exactly one work item (D2 % 8 == 0, aligned base)
    VGeom g;
};
struct VMulti {
    int n;
    unsigned int items;   // total
    int combine;          // fold the Laplacian-pyramid combination into the launch (see below)
    int dataflow;         // steps synchronised item to item through rowdone counters instead of grid barriers
    unsigned int *err;    // dataflow: sticky "a wait timed out" flag
    const float *indiv[VI_MAXL];   // combine: the levels' individual fields
    VLevel l[VI_MAXL];
};

// ---- coarse-to-fine combination folded into the integration launch (SVFDecoder.forward, src/components/pulpo.py:308;
// PULPo.combine_dfs, src/models.py:356-367):  combined_l = 2 * up2(combined_{l+1}) + individual_l, one grid barrier
// per level instead of one launch (and the adjoint behind the backward).  Optional: measured at config 2 the
// separate launches win (+27 us forward, +31 us backward here vs 20 + 25 us of small kernels that overlap with
// other streams): these phases run on the cooperative grid's ~100 k threads and are latency-bound.  Same taps,
// weights and nesting order (x, y, z) as the x2 kernels of resize.cu: bit-identical combined fields.
struct Up2Ax {
    int ia, ib;
    float wa, wb;
};
__device__ __forceinline__ Up2Ax up2_axis(int o, int n)   // output index o, input size n
{
    const int j = o >> 1;
    Up2Ax t;
    if (o & 1) {
        t.ia = j; t.ib = min(j + 1, n - 1); t.wa = 0.75f; t.wb = 0.25f;
    } else {
        t.ia = j ? j - 1 : j; t.ib = j; t.wa = j ? 0.25f : 1.0f; t.wb = j ? 0.75f : 0.0f;
    }
    return t;
}
__device__ __forceinline__ float up2_point(const float *__restrict__ p, int d1, int d2, const Up2Ax &tz, const Up2Ax &ty,
                                           const Up2Ax &tx, float premul)
{
    const float *r00 = p + ((i64)tz.ia * d1 + ty.ia) * d2, *r01 = p + ((i64)tz.ia * d1 + ty.ib) * d2;
    const float *r10 = p + ((i64)tz.ib * d1 + ty.ia) * d2, *r11 = p + ((i64)tz.ib * d1 + ty.ib) * d2;
    const float a0 = __ldcg(r00 + tx.ia), a1 = __ldcg(r00 + tx.ib), b0 = __ldcg(r01 + tx.ia), b1 = __ldcg(r01 + tx.ib);
    const float c0 = __ldcg(r10 + tx.ia), c1 = __ldcg(r10 + tx.ib), e0 = __ldcg(r11 + tx.ia), e1 = __ldcg(r11 + tx.ib);
    const float x00 = (premul * a0) * tx.wa + (premul * a1) * tx.wb, x01 = (premul * b0) * tx.wa + (premul * b1) * tx.wb;
    const float x10 = (premul * c0) * tx.wa + (premul * c1) * tx.wb, x11 = (premul * e0) * tx.wa + (premul * e1) * tx.wb;
    const float y0 = x00 * ty.wa + x01 * ty.wb, y1 = x10 * ty.wa + x11 * ty.wb;
    return y0 * tz.wa + y1 * tz.wb;
}
// adjoint weights of input j for outputs 2j-1 .. 2j+2 (resize.cu: adj4)
__device__ __forceinline__ void up2_adj_w(int j, int n, float (&w)[4])
{
    w[0] = j > 0 ? 0.25f : 0.0f;
    w[1] = j > 0 ? 0.75f : 1.0f;
    w[2] = j < n - 1 ? 0.75f : 1.0f;
    w[3] = j < n - 1 ? 0.25f : 0.0f;
}
// sum over the 4x4x4 outputs that read input (z, y, x); go: one channel volume at twice the size
// (not inlined: keeps the rarely-run combination code out of the integration loops' register allocation)
__device__ __noinline__ float up2_adjoint_point(const float *__restrict__ go, int d0, int d1, int d2, int z, int y, int x)
{
    float wz[4], wy[4], wx[4];
    up2_adj_w(z, d0, wz); up2_adj_w(y, d1, wy); up2_adj_w(x, d2, wx);
    const int o1 = 2 * d1, o2 = 2 * d2;
    float acc = 0.0f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int oz = min(max(2 * z - 1 + a, 0), 2 * d0 - 1);   // clamped taps carry weight 0
        float accy = 0.0f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int oy = min(max(2 * y - 1 + b, 0), o1 - 1);
            const float *row = go + ((i64)oz * o1 + oy) * o2;
            const float g0 = __ldcg(row + max(2 * x - 1, 0)), g1 = __ldcg(row + 2 * x), g2 = __ldcg(row + 2 * x + 1),
                        g3 = __ldcg(row + min(2 * x + 2, o2 - 1));
            accy += wy[b] * (wx[0] * g0 + wx[1] * g1 + wx[2] * g2 + wx[3] * g3);
        }
        acc += wz[a] * accy;
    }
    return acc;
}

static int make_vgeom(VGeom &g, int B, int D0, int D1, int D2)
{
    i64 S = (i64)D0 * D1 * D2;
    if ((i64)B * S >= (1ll << 31) || D0 > (1 << 22) || D1 > (1 << 22) || D2 > (1 << 22)) return PULPO_ERR_INVALID_SHAPE;
    g.B = B; g.D0 = D0; g.D1 = D1; g.D2 = D2; g.S = (int)S; g.N = (unsigned int)(B * S);
    g.unbias = tap_unbias(D1, D2);
    g.npx = (D2 + VI_PX - 1) / VI_PX;
    g.npy = (D1 + VI_PY - 1) / VI_PY;
    g.a0 = make_axis(D0); g.a1 = make_axis(D1); g.a2 = make_axis(D2);
    return PULPO_OK;
}

// grid.sync() behind a condition the compiler cannot fold (`live` is a kernel parameter, always > 0).  Measured on
// B200 (CUDA 12.9): with the bare call in the step loop a step cost ~5 us on top of its work, with the call inside a
// conditional ~1.4 us (forward 100 -> 73 us, backward 211 -> 176 us at 80x96x112; no-barrier bound 64 / 160 us).  The
// SASS differs only by a BSSY / BSYNC convergence-barrier pair that the conditional wraps around the barrier sequence
// (thread 0 of warp 0 arrives and polls alone): the warps leave the barrier reconverged.
__device__ __forceinline__ void grid_barrier(cg::grid_group &grid, int live)
{
    if (live > 0) grid.sync();
    __syncwarp();
}

struct Item {
    int b, z0, z1, y, x;
    bool valid;   // this lane's voxel column is inside the volume
};

__device__ __forceinline__ Item decode_item(unsigned int it, const VGeom &g, int lane)
{
    // x fastest: the warps of a CTA own x-adjacent patches of the same rows and planes, so they
    // share the cache lines on their common borders and spread over the L1 sets
    unsigned int r, zr, r2, px, py, b;
    fast_divmod(it, g.dnpx, r, px);
    fast_divmod(r, g.dnpy, r2, py);
    fast_divmod(r2, g.dnz, b, zr);
    Item t;
    t.b = (int)b;
    t.z0 = (int)zr * g.zrun;
    t.z1 = min(g.D0, t.z0 + g.zrun);
    t.x = (int)px * VI_PX + (lane & (VI_PX - 1));
    t.y = (int)py * VI_PY + (lane >> VI_PX_LOG2);
    t.valid = (t.x < g.D2) && (t.y < g.D1);
    return t;
}

struct Corners {
    float4 c[8];  // index = dz*4 + dy*2 + dx
};
constexpr int NOBASE = -0x40000000;   // "no previous footprint" (never equals base - plane stride)

// 128-bit load whose position in the instruction stream is kept (volatile asms are not reordered
// against each other): the next plane's own value must be requested BEFORE this plane's corner
// gathers, otherwise its latency is serialised behind theirs (two L2 round trips per plane).
__device__ __forceinline__ float4 ld4v(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// Asynchronous 16-byte global -> shared copies (LDGSTS): the voxel's own values of the next planes travel to a
// per-lane shared-memory slot without holding registers, so several planes can be in flight per warp.  The
// integration kernels are bound by memory-level parallelism (a warp walks its z run serially and 16-24 warps
// per SM with one plane each in flight cover ~1/4 of the bytes HBM latency needs): timing experiments without
// gathers and with one reduction still took 216 of 311 us.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
// same, allocating in L1: the saved state is gathered again by this and neighbouring warps
__device__ __forceinline__ void cp_async16_ca(void *smem_dst, const void *gsrc)
{
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
#ifndef PULPO_VI_RING
#define PULPO_VI_RING 4
#endif
constexpr int VI_RING = PULPO_VI_RING;   // planes of own values staged per warp (VI_RING - 1 in flight)
// Backward only.  Measured: backward 309 -> 269 us (ring depths 2, 3, 4 and 6 alike); the forward, whose gathers hit
// L1 70 % of the time and whose register prefetch one plane ahead suffices, got slower with the same ring
// (115 -> 128 us) and keeps its register prefetch.  Also measured and not kept: staging the NEXT plane's four upper
// corners the same way (footprint base from the next plane's own value in the ring, double-buffered slots): correct,
// but 279 vs 269 us -- the gathers cost L1 throughput, not exposed latency, and the extra sample-position
// computation plus 64 KB more shared memory (less L1) outweigh the overlap.  Timing experiments on the kernel
// before the ring (wrong results, timing only): no gathers 269 us, one scatter reduction instead of four 277 us,
// both 216 us of 311 us -- most of the time was the serial own-load chain, which the ring removes.

// The footprint is always 2x2x2 in-bounds (see make_tap): fixed +1 / +D2 / +D1*D2 neighbours.
// Walking a z run, the upper four corners of one plane are the lower four of the next whenever the
// footprint moved by exactly one plane (the usual case for a smooth field): those are kept in
// registers, so a plane costs four 128-bit gathers instead of eight.
// Strides are 64-bit BYTE offsets computed once per work item, so each corner address is one 64-bit add.
struct VStride {
    i64 y, z, zy;
};
__device__ __forceinline__ VStride make_vstride(int sy, int sz)
{
    VStride s;
    s.y = (i64)sy * 16; s.z = (i64)sz * 16; s.zy = s.y + s.z;
    return s;
}

__device__ __forceinline__ void gather8(const float4 *vol, int base, const VStride &st, bool reuse_lower, Corners &k)
{
    const char *p = reinterpret_cast<const char *>(vol) + (i64)base * 16;
    if (reuse_lower) {
        k.c[0] = k.c[4]; k.c[1] = k.c[5]; k.c[2] = k.c[6]; k.c[3] = k.c[7];
    } else {
        k.c[0] = ld4v(reinterpret_cast<const float4 *>(p));
        k.c[1] = ld4v(reinterpret_cast<const float4 *>(p) + 1);
        k.c[2] = ld4v(reinterpret_cast<const float4 *>(p + st.y));
        k.c[3] = ld4v(reinterpret_cast<const float4 *>(p + st.y) + 1);
    }
    k.c[4] = ld4v(reinterpret_cast<const float4 *>(p + st.z));
    k.c[5] = ld4v(reinterpret_cast<const float4 *>(p + st.z) + 1);
    k.c[6] = ld4v(reinterpret_cast<const float4 *>(p + st.zy));
    k.c[7] = ld4v(reinterpret_cast<const float4 *>(p + st.zy) + 1);
}

// same corner order / op order as the CPU grid sampler (tnw, tne, tsw, tse, bnw, ...)
template <int MODE>
__device__ __forceinline__ float interp8(float c0, float c1, float c2, float c3, float c4, float c5, float c6,
                                         float c7, const float w[8], float own)
{
    if (MODE == PULPO_COORD_FAST) {   // one FMA chain seeded with the voxel's own value: v + sum w_d c_d
        float acc = __fmaf_rn(c0, w[0], own);
        acc = __fmaf_rn(c1, w[1], acc);
        acc = __fmaf_rn(c2, w[2], acc);
        acc = __fmaf_rn(c3, w[3], acc);
        acc = __fmaf_rn(c4, w[4], acc);
        acc = __fmaf_rn(c5, w[5], acc);
        acc = __fmaf_rn(c6, w[6], acc);
        return __fmaf_rn(c7, w[7], acc);
    }
    float acc = __fmul_rn(c0, w[0]);
    acc = __fadd_rn(acc, __fmul_rn(c1, w[1]));
    acc = __fadd_rn(acc, __fmul_rn(c2, w[2]));
    acc = __fadd_rn(acc, __fmul_rn(c3, w[3]));
    acc = __fadd_rn(acc, __fmul_rn(c4, w[4]));
    acc = __fadd_rn(acc, __fmul_rn(c5, w[5]));
    acc = __fadd_rn(acc, __fmul_rn(c6, w[6]));
    acc = __fadd_rn(acc, __fmul_rn(c7, w[7]));
    return __fadd_rn(own, acc);
}

struct VFoot {
    int base;   // offset of the low corner inside one volume
    float w[8];
    float wx0, wx1, wy0, wy1, wz0, wz1;
};

template <int MODE>
__device__ __forceinline__ VFoot make_vfoot(float zf, float yf, float xf, const float4 &v, const VGeom &g,
                                            float *uz = nullptr, float *uy = nullptr, float *ux = nullptr)
{
    Tap tz = make_tap<MODE>(zf, v.x, g.a0, uz);
    Tap ty = make_tap<MODE>(yf, v.y, g.a1, uy);
    Tap tx = make_tap<MODE>(xf, v.z, g.a2, ux);
    VFoot f;
    f.base = tap_base(tz, ty, tx, g.D1, g.D2, g.unbias);
    f.wx0 = tx.w0; f.wx1 = tx.w1; f.wy0 = ty.w0; f.wy1 = ty.w1; f.wz0 = tz.w0; f.wz1 = tz.w1;
    const float w00 = __fmul_rn(tx.w0, ty.w0), w01 = __fmul_rn(tx.w1, ty.w0);
    const float w10 = __fmul_rn(tx.w0, ty.w1), w11 = __fmul_rn(tx.w1, ty.w1);
    f.w[0] = __fmul_rn(w00, tz.w0); f.w[1] = __fmul_rn(w01, tz.w0);
    f.w[2] = __fmul_rn(w10, tz.w0); f.w[3] = __fmul_rn(w11, tz.w0);
    f.w[4] = __fmul_rn(w00, tz.w1); f.w[5] = __fmul_rn(w01, tz.w1);
    f.w[6] = __fmul_rn(w10, tz.w1); f.w[7] = __fmul_rn(w11, tz.w1);
    return f;
}

// ---- Step synchronisation.  Step k+1 of a work item reads, besides its own voxels, only voxels within the
// reach of the step-k field: |p - x| <= |v_k| + 0.5 per axis (sample_pos: p = (x + v) S / (S - 1) - 0.5, clamped
// into the volume), and |v_k| <= 2^k max|v_0| because each step adds an interpolated -- convex -- value of the
// field to itself.  So instead of a grid-wide barrier per step (7 + 7 of them; the barriers with their cold-L1
// restart and tail were ~35 % of the forward and ~22 % of the backward, measured with the barriers removed),
// an item waits only for the rows of items within that reach: every item bumps its row's counter when it
// finishes a step (red.release.gpu: MEMBAR + REDG, no L1 invalidation) and polls its neighbours' rows with relaxed
// loads.  No acquire fence (CCTL.IVALL would empty the SM's L1 on every wait): a state is written once (save_steps)
// and, where `l1safe`, every 128-byte line of it belongs to one work item, so nobody can have cached a line of it
// before its owner signalled; levels with ragged rows take the fence.  All CTAs are co-resident (cooperative
// launch) and step k+1 depends on step k only, so the waits cannot deadlock; they give up after ~0.5 s anyway
// (sticky error flag; the results are then wrong, the GPU is not hung).
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned int *p, unsigned int v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

struct ItemBox {   // patch-row coordinates of a work item
    int b, zr, py, z0, z1;
};
__device__ __forceinline__ ItemBox decode_box(unsigned int it, const VGeom &g)
{
    unsigned int r, zr, r2, px, py, b;
    fast_divmod(it, g.dnpx, r, px);
    fast_divmod(r, g.dnpy, r2, py);
    fast_divmod(r2, g.dnz, b, zr);
    ItemBox x;
    x.b = (int)b; x.zr = (int)zr; x.py = (int)py;
    x.z0 = (int)zr * g.zrun; x.z1 = min(g.D0, x.z0 + g.zrun);
    return x;
}

// max |v_0| of a level from the per-CTA partials (NaN / Inf propagate: the reach then covers the volume)
__device__ __forceinline__ float level_max(const float *ctamax, int nctas, int lane)
{
    float mx = 0.0f;
    for (int i = lane; i < nctas; i += 32) {
        const float v = __ldcg(ctamax + i);
        mx = (v > mx || v != v) ? v : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float v = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = (v > mx || v != v) ? v : mx;
    }
    return mx;
}

// voxels a step can reach when max |v_0| = m0 and the field has doubled `k` times (margins: fp32 rounding of the
// interpolation weights and of the sample position)
__device__ __forceinline__ int step_reach(float m0, int k, int cap)
{
    const float bound = m0 * (float)(1u << k) * 1.0005f + 0.501f;
    if (!(bound < (float)cap)) return cap;     // also NaN
    return (int)bound + 1;
}

// wait until every row of work items within `reach` voxels of the item has completed `need_steps` steps
__device__ __forceinline__ void wait_rows(const VLevel &L, const ItemBox &x, int reach, unsigned int need_steps, int lane,
                                          unsigned int *err)
{
    const VGeom &g = L.g;
    const int zlo = max(x.z0 - reach, 0) / g.zrun, zhi = min(x.z1 - 1 + reach, g.D0 - 1) / g.zrun;
    const int ylo = max(x.py * VI_PY - reach, 0) / VI_PY, yhi = min(x.py * VI_PY + VI_PY - 1 + reach, g.D1 - 1) / VI_PY;
    const int ny = yhi - ylo + 1, n = (zhi - zlo + 1) * ny;
    const unsigned int need = need_steps * (unsigned int)g.npx;
    const unsigned int *base = L.rowdone + (i64)x.b * g.nzrun * g.npy;
    for (int i = lane; i < n; i += 32) {
        const int zr = zlo + i / ny, py = ylo + i % ny;
        const unsigned int *p = base + zr * g.npy + py;
        int spins = 0;
        while (ld_relaxed_u32(p) < need) {
            __nanosleep(100);
            if (++spins > (1 << 16) && (ld_relaxed_u32(err) != 0u || spins > (1 << 22))) {
                atomicExch(err, 1u);
                break;
            }
        }
    }
    __syncwarp();
    if (!L.l1safe) fence_acq_rel_gpu();
}

__device__ __forceinline__ void signal_row(const VLevel &L, const ItemBox &x, int lane)
{
    __syncwarp();
    if (lane == 0) red_release_add(L.rowdone + ((i64)x.b * L.g.nzrun + x.zr) * L.g.npy + x.py, 1u);
}

// One forward work item.  L0 = true: the item belongs to level 0 (the bulk of the work), whose geometry is then
// addressed at constant offsets of the kernel parameter (constant-bank operands) instead of through a run-time
// level index (one LDC per use).
template <int MODE, bool L0>
__device__ __forceinline__ void vi_fwd_item(const VMulti &m, int lv, unsigned int it, int lane, int k, int save, bool last)
{
            const VLevel &L = L0 ? m.l[0] : m.l[lv];
            const VGeom &g = L.g;
            const unsigned int N = g.N, S = g.S;
            const int sy = g.D2, sz = g.D1 * g.D2;
            const VStride vst = make_vstride(sy, sz);
            const float4 *src = save ? L.ws + (i64)k * N : L.ws + (i64)(k & 1) * N;
            float4 *dst = save ? L.ws + (i64)(k + 1) * N : L.ws + (i64)((k + 1) & 1) * N;
            const Item t = decode_item(it - L.item0, g, lane);
            if (!t.valid) return;
            const float yf = (float)t.y, xf = (float)t.x;
            const float4 *vol = src + (i64)t.b * S;
            int off = (t.z0 * g.D1 + t.y) * g.D2 + t.x;
            float4 v = ld4v(vol + off);
            Corners kc;
            int prev_base = NOBASE;
            for (int z = t.z0; z < t.z1; ++z, off += sz) {
                float4 vn = v;
                if (z + 1 < t.z1) vn = ld4v(vol + off + sz);   // the next plane's own value, ahead of this plane's gathers
                const VFoot f = make_vfoot<MODE>((float)z, yf, xf, v, g);
                gather8(vol, f.base, vst, f.base == prev_base + sz, kc);
                prev_base = f.base;
                const float r0 = interp8<MODE>(kc.c[0].x, kc.c[1].x, kc.c[2].x, kc.c[3].x, kc.c[4].x, kc.c[5].x, kc.c[6].x, kc.c[7].x, f.w, v.x);
                const float r1 = interp8<MODE>(kc.c[0].y, kc.c[1].y, kc.c[2].y, kc.c[3].y, kc.c[4].y, kc.c[5].y, kc.c[6].y, kc.c[7].y, f.w, v.y);
                const float r2 = interp8<MODE>(kc.c[0].z, kc.c[1].z, kc.c[2].z, kc.c[3].z, kc.c[4].z, kc.c[5].z, kc.c[6].z, kc.c[7].z, f.w, v.z);
                if (last) {
                    float *o = L.out + (i64)t.b * 3 * S + off;
                    o[0] = r0; o[S] = r1; o[2 * S] = r2;
                } else {
                    dst[(i64)t.b * S + off] = make_float4(r0, r1, r2, 0.0f);
                }
                v = vn;
            }
}

template <int MODE>
__global__ void __launch_bounds__(VI_FWD_THREADS, 1)
vecint_fwd_kernel(const VMulti m, int nsteps, int save, float scale)
{
    cg::grid_group grid = cg::this_grid();
    const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const unsigned int warp = tid >> 5, nwarps = nthr >> 5;

    // v_0 = vec * 2^-nsteps, planar -> interleaved
    if (m.combine) {
        // coarse to fine: combined_l = 2 * up2(combined_{l+1}) + individual_l, written to L.in (an output here) and,
        // scaled, to the first integration state; one grid barrier per level
        for (int lv = m.n - 1; lv >= 0; --lv) {
            const VLevel &L = m.l[lv];
            const unsigned int N = L.g.N, S = L.g.S;
            const float *ind = m.indiv[lv];
            if (lv == m.n - 1) {
                for (unsigned int i = tid; i < N; i += nthr) {
                    unsigned int b = i / S, v = i - b * S;
                    const float *f = ind + (i64)b * 3 * S + v;
                    L.ws[i] = make_float4(__fmul_rn(__ldg(f), scale), __fmul_rn(__ldg(f + S), scale),
                                          __fmul_rn(__ldg(f + 2 * S), scale), 0.0f);
                }
                continue;
            }
            grid_barrier(grid, m.n);
            const VLevel &C = m.l[lv + 1];
            const float *lower = (lv + 1 == m.n - 1) ? m.indiv[lv + 1] : C.in;
            const int c1 = C.g.D1, c2 = C.g.D2;
            const unsigned int Sc = C.g.S;
            float *comb = const_cast<float *>(L.in);
            for (unsigned int i = tid; i < N; i += nthr) {
                unsigned int b = i / S, v = i - b * S;
                unsigned int zy = v / (unsigned int)L.g.D2, x = v - zy * (unsigned int)L.g.D2;
                unsigned int z = zy / (unsigned int)L.g.D1, y = zy - z * (unsigned int)L.g.D1;
                const Up2Ax tz = up2_axis((int)z, C.g.D0), ty = up2_axis((int)y, c1), tx = up2_axis((int)x, c2);
                const float *f = ind + (i64)b * 3 * S + v;
                const float *lo = lower + (i64)b * 3 * Sc;
                float r0 = up2_point(lo, c1, c2, tz, ty, tx, 2.0f);
                float r1 = up2_point(lo + Sc, c1, c2, tz, ty, tx, 2.0f);
                float r2 = up2_point(lo + 2 * (i64)Sc, c1, c2, tz, ty, tx, 2.0f);
                r0 += __ldg(f); r1 += __ldg(f + S); r2 += __ldg(f + 2 * S);
                float *o = comb + (i64)b * 3 * S + v;
                o[0] = r0; o[S] = r1; o[2 * S] = r2;
                L.ws[i] = make_float4(__fmul_rn(r0, scale), __fmul_rn(r1, scale), __fmul_rn(r2, scale), 0.0f);
            }
        }
    } else {
        __shared__ float vi_red[32];
        for (int lv = 0; lv < m.n; ++lv) {
            const VLevel &L = m.l[lv];
            const unsigned int N = L.g.N, S = L.g.S;
            float mx = 0.0f;   // max |v_0| of this thread's voxels (NaN sticks)
            for (unsigned int i = tid; i < N; i += nthr) {
                unsigned int b = i / S, v = i - b * S;
                const float *f = L.in + (i64)b * 3 * S + v;
                const float4 q = make_float4(__fmul_rn(__ldg(f), scale), __fmul_rn(__ldg(f + S), scale),
                                             __fmul_rn(__ldg(f + 2 * S), scale), 0.0f);
                L.ws[i] = q;
                const float a = fmaxf(fmaxf(fabsf(q.x), fabsf(q.y)), fabsf(q.z));
                mx = (q.x != q.x || q.y != q.y || q.z != q.z) ? __int_as_float(0x7fc00000) : ((a > mx) ? a : mx);
            }
            if (m.dataflow) {
                // per-CTA partial of the level's max |v_0| and zeroed row counters, published by the grid barrier below
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float v = __shfl_xor_sync(0xffffffffu, mx, o);
                    mx = (v > mx || v != v) ? v : mx;
                }
                __syncthreads();
                if (lane == 0) vi_red[threadIdx.x >> 5] = mx;
                __syncthreads();
                if (threadIdx.x == 0) {
                    float t = 0.0f;
                    for (unsigned int w = 0; w < (blockDim.x >> 5); ++w) {
                        const float v = vi_red[w];
                        t = (v > t || v != v) ? v : t;
                    }
                    L.ctamax[blockIdx.x] = t;
                }
                const unsigned int rows = (unsigned int)(L.g.B * L.g.nzrun * L.g.npy);
                for (unsigned int i = tid; i < rows; i += nthr) L.rowdone[i] = 0u;
            }
        }
        if (m.dataflow && tid == 0) *m.err = 0u;
    }
    const bool flow = m.dataflow != 0;
    int cached_lv = -1;
    float cached_max = 0.0f;
    for (int k = 0; k < nsteps; ++k) {
#ifdef PULPO_VI_NOSYNC   // timing experiment only (wrong results): upper bound of what removing the step barriers buys
        if (k == 0)
#endif
        if (k == 0 || !flow) grid_barrier(grid, m.n);
        const bool last = (k == nsteps - 1);
        for (unsigned int it = warp; it < m.items; it += nwarps) {
            int lv = 0;
#pragma unroll
            for (int j = 1; j < VI_MAXL; ++j) lv += (j < m.n && it >= m.l[j].item0) ? 1 : 0;
            ItemBox box;
            if (flow) {
                const VLevel &L = m.l[lv];
                box = decode_box(it - L.item0, L.g);
                if (k > 0) {
                    if (lv != cached_lv) {
                        cached_max = level_max(L.ctamax, (int)gridDim.x, lane);
                        cached_lv = lv;
                    }
                    const int cap = max(L.g.D0, max(L.g.D1, L.g.D2));
                    wait_rows(L, box, step_reach(cached_max, k, cap), (unsigned int)k, lane, m.err);
                }
            }
            if (lv == 0)
                vi_fwd_item<MODE, true>(m, 0, it, lane, k, save, last);
            else
                vi_fwd_item<MODE, false>(m, lv, it, lane, k, save, last);
            if (flow && !last) signal_row(m.l[lv], box, lane);
        }
    }
    if (nsteps == 0) {
        grid_barrier(grid, m.n);
        for (int lv = 0; lv < m.n; ++lv) {
            const VLevel &L = m.l[lv];
            const unsigned int N = L.g.N, S = L.g.S;
            for (unsigned int i = tid; i < N; i += nthr) {
                unsigned int b = i / S, v = i - b * S;
                float4 t = L.ws[i];
                float *o = L.out + (i64)b * 3 * S + v;
                o[0] = t.x; o[S] = t.y; o[2 * S] = t.z;
            }
        }
    }
}

#ifdef PULPO_VI_TRACE
// tuning builds only: per-CTA (step start, work done) timestamps of the backward, read back by
// pulpo_debug_vi_trace() to look at load balance across SMs
__device__ unsigned long long g_vi_trace[148 * 64];
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

struct F3 {
    float x, y, z;
};

__device__ __forceinline__ F3 shfl_up3(const F3 &a, int d)
{
    F3 r;
    r.x = __shfl_up_sync(0xffffffffu, a.x, d);
    r.y = __shfl_up_sync(0xffffffffu, a.y, d);
    r.z = __shfl_up_sync(0xffffffffu, a.z, d);
    return r;
}

__device__ __forceinline__ void red3(float4 *addr, const F3 &a)
{
    red_add_v4(reinterpret_cast<float *>(addr), a.x, a.y, a.z, 0.0f);
}

// Backward of one step  v' = v + W(v) v :   g = g' + W(v)^T g' + (dW/dv : v)^T g'
//   own    : g' + gather-form gradient through the sample position
//   scatter: w_d * g' onto the 8 corners -> red.global.add.v4.f32, the four upper-corner contributions
//            carried in registers to the next plane of the z run (COMBINE == 2, default)
// Gradient states (float4 per voxel).  Default: three states rotate -- X is read, Y (zero at step start)
// receives both parts through vector reductions, Z is cleared for the next step by the thread that owns the
// voxel; 48 B/voxel of scratch keeps more of the working set in L2 than the alternative
// (-DPULPO_VI_BWD_PY: two (own P, scatter Y) pairs, the own part a plain store; 64 B/voxel; measured
// 324 vs 312 us at config 2, 979 vs 904 us at two pairs per GPU).
template <int MODE, int COMBINE>
__global__ void __launch_bounds__(VI_BWD_THREADS, 1)
vecint_bwd_kernel(const VMulti m, int nsteps, float scale)
{
    cg::grid_group grid = cg::this_grid();
    const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, lx = lane & (VI_PX - 1), ly = lane >> VI_PX_LOG2;
    const unsigned int warp = tid >> 5, nwarps = nthr >> 5;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int NOADDR = NOBASE;
    (void)lx; (void)ly;

    for (int lv = 0; lv < m.n; ++lv) {
        const VLevel &L = m.l[lv];
        const unsigned int N = L.g.N, S = L.g.S;
#ifdef PULPO_VI_BWD_XYZ
        float4 *Pa = L.scr, *Ya = L.scr + N, *Yb = L.scr + N;      // X = gout, Y = 0
#else
        float4 *Pa = L.scr, *Ya = L.scr + N, *Yb = L.scr + 3 * (i64)N;
#endif
        float mx = 0.0f;   // dataflow: max |v_0| over the saved first state (bounds every step's reach)
        for (unsigned int i = tid; i < N; i += nthr) {
            unsigned int b = i / S, v = i - b * S;
            const float *f = L.in + (i64)b * 3 * S + v;
            Pa[i] = make_float4(__ldg(f), __ldg(f + S), __ldg(f + 2 * S), 0.0f);
            Ya[i] = zero4;
#ifndef PULPO_VI_BWD_XYZ
            Yb[i] = zero4;
#endif
            if (m.dataflow && nsteps > 0) {
                const float4 q = __ldg(L.ws + i);
                const float a = fmaxf(fmaxf(fabsf(q.x), fabsf(q.y)), fabsf(q.z));
                mx = (q.x != q.x || q.y != q.y || q.z != q.z) ? __int_as_float(0x7fc00000) : ((a > mx) ? a : mx);
            }
        }
        if (m.dataflow) {
            __shared__ float vi_red[32];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float v = __shfl_xor_sync(0xffffffffu, mx, o);
                mx = (v > mx || v != v) ? v : mx;
            }
            __syncthreads();
            if (lane == 0) vi_red[threadIdx.x >> 5] = mx;
            __syncthreads();
            if (threadIdx.x == 0) {
                float t = 0.0f;
                for (unsigned int w = 0; w < (blockDim.x >> 5); ++w) {
                    const float v = vi_red[w];
                    t = (v > t || v != v) ? v : t;
                }
                L.ctamax[blockIdx.x] = t;
            }
            const unsigned int rows = (unsigned int)(L.g.B * L.g.nzrun * L.g.npy);
            for (unsigned int i = tid; i < rows; i += nthr) L.rowdone[i] = 0u;
        }
    }
    if (m.dataflow && tid == 0) *m.err = 0u;
    const bool flow = m.dataflow != 0;
    int cached_lv = -1;
    float cached_max = 0.0f;
    int flip = 0;   // which (P, Y) pair holds the incoming gradient of the current step
    for (int k = nsteps - 1; k >= 0; --k, flip ^= 1) {
#ifdef PULPO_VI_NOSYNC
        if (k == nsteps - 1)
#endif
        if (k == nsteps - 1 || !flow) grid_barrier(grid, m.n);
#ifdef PULPO_VI_TRACE
        if (threadIdx.x == 0 && blockIdx.x < 148 && k < 32) g_vi_trace[blockIdx.x * 64 + 2 * k] = gtime();
#endif
        for (unsigned int it = warp; it < m.items; it += nwarps) {   // warp-uniform loop: all lanes shuffle
            int lv = 0;
#pragma unroll
            for (int j = 1; j < VI_MAXL; ++j) lv += (j < m.n && it >= m.l[j].item0) ? 1 : 0;
            const VLevel &L = m.l[lv];
            const VGeom &g = L.g;
            const unsigned int N = g.N, S = g.S;
            const int sy = g.D2, sz = g.D1 * g.D2;
            const VStride vst = make_vstride(sy, sz);
            ItemBox box;
            if (flow) {
                // this step reads the gradient state the previous backward step (forward step k+1) scattered into,
                // within that step's reach, and scatters into / clears states whose owners must be past it too
                box = decode_box(it - L.item0, g);
                if (k < nsteps - 1) {
                    if (lv != cached_lv) {
                        cached_max = level_max(L.ctamax, (int)gridDim.x, lane);
                        cached_lv = lv;
                    }
                    wait_rows(L, box, step_reach(cached_max, k + 1, max(g.D0, max(g.D1, g.D2))), (unsigned int)(nsteps - 1 - k), lane,
                              m.err);
                }
            }
#ifdef PULPO_VI_BWD_XYZ
            // three rotating states: X (incoming gradient), Y (accumulates own + scatter parts, zero at step start),
            // Z (being zeroed for the next step); Pa = X, Yb = Y, Ya = Z in the names below
            const int rot = (nsteps - 1 - k) % 3;
            float4 *Pa = L.scr + (i64)rot * N, *Yb = L.scr + (i64)((rot + 1) % 3) * N, *Ya = L.scr + (i64)((rot + 2) % 3) * N;
            float4 *Pb = Yb;
#else
            float4 *Pa = L.scr + (flip ? 2 : 0) * (i64)N, *Ya = Pa + N;
            float4 *Pb = L.scr + (flip ? 0 : 2) * (i64)N, *Yb = Pb + N;
#endif
            // autograd chain of the sample position: (S/2) * 2/(S-1) per axis where the clamp is inactive
            const float kz = (MODE == PULPO_COORD_FAST) ? g.a0.kf : g.a0.gmul * (2.0f * g.a0.rcp);
            const float ky = (MODE == PULPO_COORD_FAST) ? g.a1.kf : g.a1.gmul * (2.0f * g.a1.rcp);
            const float kx = (MODE == PULPO_COORD_FAST) ? g.a2.kf : g.a2.gmul * (2.0f * g.a2.rcp);
            const float4 *vk = L.ws + (i64)k * N;
            const Item t = decode_item(it - L.item0, g, lane);
            const float yf = (float)t.y, xf = (float)t.x;
            const i64 vb = (i64)t.b * S;
            const float4 *vol = vk + vb;
            float4 *acc = Yb + vb;
            int off = t.valid ? (t.z0 * g.D1 + t.y) * g.D2 + t.x : 0;
#ifdef PULPO_VI_BWD_XYZ
            // own values (incoming gradient X, saved state v_k) of the next planes: asynchronous copies into this
            // lane's slots of the warp's ring, VI_RING - 1 planes ahead
            extern __shared__ __align__(16) float4 vi_ring[];
            float4 *ring = vi_ring + (threadIdx.x >> 5) * (VI_RING * 64) + lane;   // slot d: X at ring[d*64], v at ring[d*64+32]
            const int nz = t.z1 - t.z0;
#pragma unroll
            for (int d = 0; d < VI_RING - 1; ++d) {
                if (t.valid && d < nz) {
                    cp_async16(ring + d * 64, Pa + vb + off + (i64)d * sz);
                    cp_async16_ca(ring + d * 64 + 32, vol + off + (i64)d * sz);
                }
                cp_async_commit();
            }
#endif
            F3 carry = {0.f, 0.f, 0.f};   // corner (dz=1, dy=0, dx=0) of the previous plane, already combined in x and y
            F3 up[4];                     // COMBINE == 2: all four upper corners of the previous plane
#pragma unroll
            for (int d = 0; d < 4; ++d) up[d] = carry;
            int carry_addr = NOADDR;
#ifndef PULPO_VI_BWD_XYZ
            float4 Gn = zero4, vn = zero4;
            if (t.valid) {
                const float4 p = ld4v(Pa + vb + off), y = ld4v(Ya + vb + off);
                Gn = make_float4(p.x + y.x, p.y + y.y, p.z + y.z, 0.0f);
                vn = ld4v(vol + off);
            }
#endif
            Corners kc;
            int prev_base = NOBASE;
            int slot = 0;   // ring slot of the current plane
            for (int z = t.z0; z < t.z1; ++z, off += sz) {
#ifdef PULPO_VI_BWD_XYZ
                {   // request plane z + VI_RING - 1 into the slot consumed one iteration ago, then wait for plane z
                    const int ahead = z - t.z0 + VI_RING - 1;
                    const int wslot = slot == 0 ? VI_RING - 1 : slot - 1;
                    if (t.valid && ahead < nz) {
                        cp_async16(ring + wslot * 64, Pa + vb + off + (i64)(VI_RING - 1) * sz);
                        cp_async16_ca(ring + wslot * 64 + 32, vol + off + (i64)(VI_RING - 1) * sz);
                    }
                    cp_async_commit();
                    cp_async_wait<VI_RING - 1>();
                }
                float4 G = zero4, v = zero4;
                if (t.valid) {
                    G = ring[slot * 64];
                    v = ring[slot * 64 + 32];
                    Ya[vb + off] = zero4;   // clear it for the step after next
                }
                slot = slot == VI_RING - 1 ? 0 : slot + 1;
#else
                const float4 G = Gn, v = vn;
                if (t.valid) {
                    Ya[vb + off] = zero4;   // read above (or one iteration ago): clear it for the step after next
                    if (z + 1 < t.z1) {     // next plane's own values, requested ahead of this plane's gathers
                        const float4 p = ld4v(Pa + vb + off + sz), y = ld4v(Ya + vb + off + sz);
                        Gn = make_float4(p.x + y.x, p.y + y.y, p.z + y.z, 0.0f);
                        vn = ld4v(vol + off + sz);
                    }
                }
#endif
                float uz, uy, ux;
                const VFoot f = make_vfoot<MODE>((float)z, yf, xf, v, g, &uz, &uy, &ux);
                const int A = t.valid ? f.base : NOADDR;
                if (t.valid) {
                    gather8(vol, f.base, vst, f.base == prev_base + sz, kc);
                    prev_base = f.base;
                    // q[d] = <corner_d, G> over the 3 channels
                    float q[8];
#pragma unroll
                    for (int d = 0; d < 8; ++d) q[d] = kc.c[d].x * G.x + kc.c[d].y * G.y + kc.c[d].z * G.z;
                    const float sx = ((q[1] - q[0]) * f.wy0 + (q[3] - q[2]) * f.wy1) * f.wz0 +
                                     ((q[5] - q[4]) * f.wy0 + (q[7] - q[6]) * f.wy1) * f.wz1;
                    const float sy_ = ((q[2] - q[0]) * f.wx0 + (q[3] - q[1]) * f.wx1) * f.wz0 +
                                      ((q[6] - q[4]) * f.wx0 + (q[7] - q[5]) * f.wx1) * f.wz1;
                    const float sz_ = ((q[4] - q[0]) * f.wx0 + (q[5] - q[1]) * f.wx1) * f.wy0 +
                                      ((q[6] - q[2]) * f.wx0 + (q[7] - q[3]) * f.wx1) * f.wy1;
                    const float mz = (uz > 0.0f && uz < g.a0.Sm1) ? kz : 0.0f;
                    const float my = (uy > 0.0f && uy < g.a1.Sm1) ? ky : 0.0f;
                    const float mx = (ux > 0.0f && ux < g.a2.Sm1) ? kx : 0.0f;
#ifdef PULPO_VI_BWD_XYZ
                    red_add_v4(reinterpret_cast<float *>(Pb + vb + off), G.x + mz * sz_, G.y + my * sy_, G.z + mx * sx, 0.0f);
#else
                    Pb[vb + off] = make_float4(G.x + mz * sz_, G.y + my * sy_, G.z + mx * sx, 0.0f);
#endif
                }
                // ---- scatter half.  c[d] = w_d * G for the 8 corners (all in-bounds; border corners carry weight 0)
                F3 c[8];
#pragma unroll
                for (int d = 0; d < 8; ++d) {
                    c[d].x = f.w[d] * G.x; c[d].y = f.w[d] * G.y; c[d].z = f.w[d] * G.z;
                }
                if (COMBINE == 0) {        // one reduction per corner
                    if (t.valid) {
                        float4 *q = acc + A;
                        red3(q, c[0]); red3(q + 1, c[1]); red3(q + sy, c[2]); red3(q + sy + 1, c[3]);
                        red3(q + sz, c[4]); red3(q + sz + 1, c[5]); red3(q + sz + sy, c[6]); red3(q + sz + sy + 1, c[7]);
                    }
                    continue;
                }
                if (COMBINE == 2) {
                    // z-carry only: the four upper corners of the previous plane are this plane's lower corners
                    // whenever the footprint moved by exactly one plane -> 4 reductions per voxel, no shuffles
                    if (t.valid) {
                        if (carry_addr == A) {
#pragma unroll
                            for (int d = 0; d < 4; ++d) { c[d].x += up[d].x; c[d].y += up[d].y; c[d].z += up[d].z; }
                        } else if (carry_addr != NOADDR) {
                            float4 *q = acc + carry_addr;
                            red3(q, up[0]); red3(q + 1, up[1]); red3(q + sy, up[2]); red3(q + sy + 1, up[3]);
                        }
                        float4 *q = acc + A;
                        red3(q, c[0]); red3(q + 1, c[1]); red3(q + sy, c[2]); red3(q + sy + 1, c[3]);
#pragma unroll
                        for (int d = 0; d < 4; ++d) up[d] = c[4 + d];
                        carry_addr = A + sz;
                    }
                    continue;
                }
                // combine along x: my dx=1 corners are lane+1's dx=0 corners when its footprint starts one voxel right
                const int A_left = __shfl_up_sync(0xffffffffu, A, 1), A_right = __shfl_down_sync(0xffffffffu, A, 1);
                const bool recv_x = (lx > 0) && (A_left + 1 == A);
                const bool sent_x = (lx < VI_PX - 1) && (A + 1 == A_right);
#pragma unroll
                for (int m = 0; m < 4; ++m) {   // m = dz*2 + dy
                    const F3 r = shfl_up3(c[2 * m + 1], 1);
                    if (recv_x) { c[2 * m].x += r.x; c[2 * m].y += r.y; c[2 * m].z += r.z; }
                }
                // combine along y: my (dy=1, dx=0) corners are lane+8's (dy=0, dx=0) corners
                const int A_up = __shfl_up_sync(0xffffffffu, A, VI_PX), A_down = __shfl_down_sync(0xffffffffu, A, VI_PX);
                const bool recv_y = (ly > 0) && (A_up + sy == A);
                const bool sent_y = (ly < VI_PY - 1) && (A + sy == A_down);
#pragma unroll
                for (int dz = 0; dz < 2; ++dz) {
                    const F3 r = shfl_up3(c[4 * dz + 2], VI_PX);
                    if (recv_y) { c[4 * dz].x += r.x; c[4 * dz].y += r.y; c[4 * dz].z += r.z; }
                }
                if (t.valid) {
                    // combine along z: the previous plane's (dz=1,dy=0,dx=0) corner is usually this plane's (0,0,0)
                    if (carry_addr == A) {
                        c[0].x += carry.x; c[0].y += carry.y; c[0].z += carry.z;
                    } else if (carry_addr != NOADDR) {
                        red3(acc + carry_addr, carry);
                    }
                    carry = c[4];
                    carry_addr = A + sz;
                    float4 *q = acc + A;
                    red3(q, c[0]);
                    if (!sent_y) {
                        red3(q + sy, c[2]);
                        red3(q + sz + sy, c[6]);
                    }
                    if (!sent_x) {
                        red3(q + 1, c[1]);
                        red3(q + sy + 1, c[3]);
                        red3(q + sz + 1, c[5]);
                        red3(q + sz + sy + 1, c[7]);
                    }
                }
            }
            if (COMBINE == 1 && carry_addr != NOADDR) red3(acc + carry_addr, carry);
            if (COMBINE == 2 && carry_addr != NOADDR) {
                float4 *q = acc + carry_addr;
                red3(q, up[0]); red3(q + 1, up[1]); red3(q + sy, up[2]); red3(q + sy + 1, up[3]);
            }
            if (flow && k > 0) signal_row(L, box, lane);
        }
#ifdef PULPO_VI_TRACE
        __syncthreads();
        if (threadIdx.x == 0 && blockIdx.x < 148 && k < 32) g_vi_trace[blockIdx.x * 64 + 2 * k + 1] = gtime();
#endif
    }
    grid_barrier(grid, m.n);
    for (int lv = 0; lv < m.n; ++lv) {
        const VLevel &L = m.l[lv];
        const unsigned int N = L.g.N, S = L.g.S;
#ifdef PULPO_VI_BWD_XYZ
        const float4 *Pa = L.scr + (i64)(nsteps % 3) * N, *Ya = L.scr + (i64)((nsteps + 1) % 3) * N;   // X; the last step's Z is all zero
#else
        const float4 *Pa = L.scr + (flip ? 2 : 0) * (i64)N, *Ya = Pa + N;
#endif
        if (m.combine && lv > 0) {
            // adjoint of the combination: the finer level's (complete) gradient flows into this level through
            // 2 * up2^T; fine to coarse, one grid barrier per level
            grid_barrier(grid, m.n);
            const VLevel &F = m.l[lv - 1];
            const unsigned int Sf = F.g.S;
            const int d0 = L.g.D0, d1 = L.g.D1, d2 = L.g.D2;
            for (unsigned int i = tid; i < N; i += nthr) {
                unsigned int b = i / S, v = i - b * S;
                unsigned int zy = v / (unsigned int)d2, x = v - zy * (unsigned int)d2;
                unsigned int z = zy / (unsigned int)d1, yy = zy - z * (unsigned int)d1;
#ifdef PULPO_VI_BWD_XYZ
                const float4 p = Pa[i], y = zero4;   // the last step's Z state is all zero: not read
#else
                const float4 p = Pa[i], y = Ya[i];
#endif
                const float *go = F.out + (i64)b * 3 * Sf;
                float *o = L.out + (i64)b * 3 * S + v;
                o[0] = (p.x + y.x) * scale + 2.0f * up2_adjoint_point(go, d0, d1, d2, (int)z, (int)yy, (int)x);
                o[S] = (p.y + y.y) * scale + 2.0f * up2_adjoint_point(go + Sf, d0, d1, d2, (int)z, (int)yy, (int)x);
                o[2 * S] = (p.z + y.z) * scale + 2.0f * up2_adjoint_point(go + 2 * (i64)Sf, d0, d1, d2, (int)z, (int)yy, (int)x);
            }
            continue;
        }
        for (unsigned int i = tid; i < N; i += nthr) {
            unsigned int b = i / S, v = i - b * S;
#ifdef PULPO_VI_BWD_XYZ
            const float4 p = Pa[i], y = zero4;   // the last step's Z state is all zero: not read
#else
            const float4 p = Pa[i], y = Ya[i];
#endif
            float *o = L.out + (i64)b * 3 * S + v;
            o[0] = (p.x + y.x) * scale; o[S] = (p.y + y.y) * scale; o[2 * S] = (p.z + y.z) * scale;
        }
    }
}

constexpr size_t VI_TAIL_CTAS = 2048;   // per-CTA slots in a workspace tail (dataflow synchronisation)

template <typename K>
static int coop_ctas(K kernel, int threads, size_t smem = 0)
{
    int dev = 0, sms = kSMs, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

// Split every patch column of every level into z runs of (at most) t planes, t the smallest run length
// for which the joint item list fits the resident warps: one work item per warp, no second round.
static void plan_items(VMulti &m, int total_warps)
{
    i64 colplanes = 0;
    for (int l = 0; l < m.n; ++l) colplanes += (i64)m.l[l].g.B * m.l[l].g.npy * m.l[l].g.npx * m.l[l].g.D0;
    int t = (int)((colplanes + total_warps - 1) / total_warps);
    if (t < 1) t = 1;
    for (;; ++t) {
        i64 items = 0;
        int maxd = 1;
        for (int l = 0; l < m.n; ++l) {
            VGeom &g = m.l[l].g;
            g.nzrun = (g.D0 + t - 1) / t;
            g.zrun = (g.D0 + g.nzrun - 1) / g.nzrun;   // balanced runs: no warp gets more than ceil(D0 / nzrun) planes
            items += (i64)g.B * g.npy * g.npx * g.nzrun;
            if (g.D0 > maxd) maxd = g.D0;
        }
        if (items <= total_warps || t >= maxd) break;
    }
    unsigned int first = 0;
    for (int l = 0; l < m.n; ++l) {
        VGeom &g = m.l[l].g;
        g.items = (unsigned int)((i64)g.B * g.npy * g.npx * g.nzrun);
        g.dnpx = make_fastdiv(g.npx); g.dnpy = make_fastdiv(g.npy); g.dnz = make_fastdiv(g.nzrun);
        m.l[l].item0 = first;
        first += g.items;
    }
    m.items = first;
}

template <typename K>
static int launch_coop(K kernel, int threads, VMulti &m, void **args, cudaStream_t st, size_t smem = 0)
{
    const int ctas = coop_ctas(kernel, threads, smem);
    // lanes cover whole 8x4 patches, so size the grid by patch-padded voxels
    i64 padded = 0;
    for (int l = 0; l < m.n; ++l) padded += (i64)m.l[l].g.B * m.l[l].g.D0 * m.l[l].g.npy * VI_PY * m.l[l].g.npx * VI_PX;
    i64 grid = (padded + threads - 1) / threads;
    if (grid > ctas) grid = ctas;
    if (grid < 1) grid = 1;
    if ((size_t)grid > VI_TAIL_CTAS) m.dataflow = 0;
    plan_items(m, (int)grid * (threads / 32));
    cudaError_t e = cudaLaunchCooperativeKernel((void *)kernel, dim3((unsigned int)grid), dim3(threads), args, smem, st);
    return e == cudaSuccess ? launch_status() : PULPO_ERR_CUDA;
}

// Tail of a level's state workspace (forward) / gradient scratch (backward) used by the dataflow synchronisation:
// row counters (upper bound: one z run per plane), per-CTA maxima, the error flag.
static size_t vi_tail_bytes(int B, int D0, int D1)
{
    const size_t rows = (size_t)B * D0 * ((D1 + VI_PY - 1) / VI_PY);
    return 128 + ((rows * 4 + 127) / 128) * 128 + VI_TAIL_CTAS * 4 + 128;
}
static void vi_tail_ptrs(VLevel &L, char *tail, int B, int D0, int D1)
{
    tail = (char *)(((uintptr_t)tail + 127) & ~(uintptr_t)127);
    const size_t rows = (size_t)B * D0 * ((D1 + VI_PY - 1) / VI_PY);
    L.rowdone = (unsigned int *)tail;
    L.ctamax = (float *)(tail + ((rows * 4 + 127) / 128) * 128);
}
static bool vi_dataflow_enabled()
{
    // Opt-in ("1"; read per call).  Measured on B200 at config 2 (scripts/vi_ab.py, profiles/r2_vecint_sync.md): the
    // item-to-item synchronisation is correct (bit-identical forward) but SLOWER than grid barriers -- forward 107 vs
    // 74 us, backward 188 vs 176 us at 80x96x112 -- because every item pays a MEMBAR.GPU + an L2 round trip of polling
    // per step (~2-3 us), more than the 1.4 us a grid barrier costs once the steps are balanced.
    const char *e = getenv("PULPO_VI_DATAFLOW");
    return e && e[0] == '1';
}

static int fill_levels(VMulti &m, const pulpo_vecint_level *levels, int nlevels, int B, bool bwd, int nsteps, int save)
{
    PULPO_REQUIRE(levels && nlevels >= 1 && nlevels <= VI_MAXL, PULPO_ERR_INVALID_SHAPE);
    m.n = nlevels;
    m.combine = 0;
    for (int l = 0; l < VI_MAXL; ++l) m.indiv[l] = nullptr;
    i64 total = 0;
    for (int l = 0; l < nlevels; ++l) {
        const pulpo_vecint_level &v = levels[l];
        PULPO_REQUIRE(v.in && v.out && v.ws, PULPO_ERR_NULL_POINTER);
        PULPO_REQUIRE(v.D0 >= 2 && v.D1 >= 2 && v.D2 >= 2, PULPO_ERR_INVALID_SHAPE);
        int rc = make_vgeom(m.l[l].g, B, v.D0, v.D1, v.D2);
        if (rc != PULPO_OK) return rc;
        if (bwd) {
            PULPO_REQUIRE(v.scratch, PULPO_ERR_NULL_POINTER);
            PULPO_REQUIRE(v.scratch_bytes >= pulpo_vecint_bwd_scratch_bytes(B, v.D0, v.D1, v.D2) && aligned16(v.scratch) &&
                              aligned16(v.ws),
                          PULPO_ERR_WORKSPACE);
        } else {
            PULPO_REQUIRE(v.ws_bytes >= pulpo_vecint_ws_bytes(nsteps, save, B, v.D0, v.D1, v.D2) && aligned16(v.ws),
                          PULPO_ERR_WORKSPACE);
        }
        m.l[l].in = v.in; m.l[l].out = v.out; m.l[l].ws = (float4 *)v.ws; m.l[l].scr = (float4 *)v.scratch;
        m.l[l].rowdone = nullptr; m.l[l].ctamax = nullptr; m.l[l].l1safe = 1;
        const size_t state = (size_t)B * v.D0 * v.D1 * v.D2 * sizeof(float4);
        if (bwd) {
#ifdef PULPO_VI_BWD_XYZ
            vi_tail_ptrs(m.l[l], (char *)v.scratch + 3 * state, B, v.D0, v.D1);   // backward: L1-cached loads touch read-only states only
#endif
        } else if (save && nsteps >= 1) {
            vi_tail_ptrs(m.l[l], (char *)v.ws + (size_t)nsteps * state, B, v.D0, v.D1);
            m.l[l].l1safe = (v.D2 % VI_PX == 0) && (((uintptr_t)v.ws) % 128 == 0) && (VI_PX * sizeof(float4) == 128);
        }
        total += (i64)B * v.D0 * v.D1 * v.D2;
    }
    PULPO_REQUIRE(total < (1ll << 31), PULPO_ERR_INVALID_SHAPE);
    // item-to-item step synchronisation needs every state to be written exactly once (saved steps) and room for the
    // counters in the workspace tail; otherwise the steps are separated by grid barriers
    m.dataflow = 0;
    m.err = nullptr;
    if (vi_dataflow_enabled() && nsteps >= 2 && m.l[0].rowdone && (bwd || save)) {
        m.dataflow = 1;
        m.err = (unsigned int *)((char *)m.l[0].ctamax + VI_TAIL_CTAS * 4);
    }
    return PULPO_OK;
}

}  // namespace pulpo

using namespace pulpo;

#ifdef PULPO_VI_TRACE
extern "C" int pulpo_debug_vi_trace(unsigned long long *host_out)
{
    return cudaMemcpyFromSymbol(host_out, g_vi_trace, sizeof(unsigned long long) * 148 * 64) == cudaSuccess ? 0 : -5;
}
#endif

extern "C" size_t pulpo_vecint_ws_bytes(int nsteps, int save_steps, int B, int D0, int D1, int D2)
{
    size_t state = (size_t)B * D0 * D1 * D2 * sizeof(float4);
    int n = save_steps ? (nsteps < 1 ? 1 : nsteps) : 2;
    // one synchronisation tail per batch item: batches of large volumes run item by item, each on its own slice
    return state * (size_t)n + (save_steps ? (size_t)B * vi_tail_bytes(1, D0, D1) : 0);
}

extern "C" size_t pulpo_vecint_bwd_scratch_bytes(int B, int D0, int D1, int D2)
{
#ifdef PULPO_VI_BWD_XYZ
    return (size_t)B * D0 * D1 * D2 * sizeof(float4) * 3 + (size_t)B * vi_tail_bytes(1, D0, D1);   // three rotating gradient states + sync tails
#else
    return (size_t)B * D0 * D1 * D2 * sizeof(float4) * 4;   // two (P, Y) pairs
#endif
}

// the levels of a combined launch are a x2 pyramid, fine to coarse
static int check_pyramid(const pulpo_vecint_level *levels, int nlevels)
{
    for (int l = 0; l + 1 < nlevels; ++l)
        PULPO_REQUIRE(levels[l].D0 == 2 * levels[l + 1].D0 && levels[l].D1 == 2 * levels[l + 1].D1 &&
                          levels[l].D2 == 2 * levels[l + 1].D2,
                      PULPO_ERR_INVALID_SHAPE);
    return PULPO_OK;
}

// Batches of large volumes are integrated one item after the other (B launches): the states of ONE item
// (13.8 MB per state at 80x96x112, 7 saved + 3 gradient states) about fill the 126 MB L2, and a joint launch
// walks all items in every step, pushing the scatter/gather traffic out to HBM (backward at config 2: 0.31 ms
// for one pair, 0.90 / 1.85 ms for 2 / 4 pairs jointly).  Small volumes stay in one launch (launch/barrier bound).
// The saved states are therefore laid out per item ([B][steps][S]) whenever this returns true; forward and
// backward take the same decision from the same arguments.
static bool split_batch(const pulpo_vecint_level *levels, int nlevels, int B)
{
    if (B <= 1 || !levels) return false;
    i64 vox = 0;
    for (int l = 0; l < nlevels; ++l) vox += (i64)levels[l].D0 * levels[l].D1 * levels[l].D2;
    return vox >= (1 << 19);
}

static void item_levels(pulpo_vecint_level *dst, const pulpo_vecint_level *src, int nlevels, int b, int nsteps, int save,
                        bool bwd)
{
    for (int l = 0; l < nlevels; ++l) {
        const pulpo_vecint_level &v = src[l];
        const size_t S = (size_t)v.D0 * v.D1 * v.D2;
        const size_t wsb = pulpo_vecint_ws_bytes(nsteps, save, 1, v.D0, v.D1, v.D2);
        const size_t scb = pulpo_vecint_bwd_scratch_bytes(1, v.D0, v.D1, v.D2);
        dst[l] = v;
        dst[l].in = v.in ? v.in + (size_t)b * 3 * S : nullptr;
        dst[l].out = v.out ? v.out + (size_t)b * 3 * S : nullptr;
        dst[l].ws = v.ws ? (char *)v.ws + (size_t)b * wsb : nullptr;
        dst[l].ws_bytes = v.ws_bytes >= (size_t)(b + 1) * wsb ? wsb : 0;
        dst[l].scratch = (bwd && v.scratch) ? (char *)v.scratch + (size_t)b * scb : v.scratch;
        dst[l].scratch_bytes = bwd ? (v.scratch_bytes >= (size_t)(b + 1) * scb ? scb : 0) : v.scratch_bytes;
    }
}

static int vecint_multi_fwd_impl(const pulpo_vecint_level *levels, const float *const *indiv, int nlevels, int nsteps,
                                 int save_steps, int B, int coord_mode, pulpo_stream_t stream)
{
    PULPO_REQUIRE(B > 0 && nsteps >= 0 && nsteps <= 30, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(levels && nlevels >= 1 && nlevels <= VI_MAXL, PULPO_ERR_INVALID_SHAPE);
    if (split_batch(levels, nlevels, B)) {
        pulpo_vecint_level one[VI_MAXL];
        const float *ind[VI_MAXL];
        for (int b = 0; b < B; ++b) {
            item_levels(one, levels, nlevels, b, nsteps, save_steps, false);
            if (indiv)
                for (int l = 0; l < nlevels; ++l)
                    ind[l] = indiv[l] ? indiv[l] + (size_t)b * 3 * levels[l].D0 * levels[l].D1 * levels[l].D2 : nullptr;
            int rc = vecint_multi_fwd_impl(one, indiv ? ind : nullptr, nlevels, nsteps, save_steps, 1, coord_mode, stream);
            if (rc != PULPO_OK) return rc;
        }
        return PULPO_OK;
    }
    coord_mode &= 0xff;   // high bits: backward tuning switches
    PULPO_REQUIRE(coord_mode >= 0 && coord_mode <= 2, PULPO_ERR_UNSUPPORTED);
    VMulti m;
    int rc = fill_levels(m, levels, nlevels, B, false, nsteps, save_steps);
    if (rc != PULPO_OK) return rc;
    if (indiv) {
        rc = check_pyramid(levels, nlevels);
        if (rc != PULPO_OK) return rc;
        for (int l = 0; l < nlevels; ++l) {
            PULPO_REQUIRE(indiv[l], PULPO_ERR_NULL_POINTER);
            m.indiv[l] = indiv[l];
        }
        m.combine = 1;
        m.dataflow = 0;   // the in-launch combination has its own grid-wide phases
    }
    float scale = 1.0f / (float)(1u << nsteps);
    void *args[] = {&m, &nsteps, &save_steps, &scale};
    cudaStream_t st = (cudaStream_t)stream;
    if (coord_mode == 0) return launch_coop(vecint_fwd_kernel<0>, VI_FWD_THREADS, m, args, st);
    if (coord_mode == 1) return launch_coop(vecint_fwd_kernel<1>, VI_FWD_THREADS, m, args, st);
    return launch_coop(vecint_fwd_kernel<2>, VI_FWD_THREADS, m, args, st);
}

extern "C" int pulpo_vecint_multi_fwd(const pulpo_vecint_level *levels, int nlevels, int nsteps, int save_steps, int B,
                                      int coord_mode, pulpo_stream_t stream)
{
    return vecint_multi_fwd_impl(levels, nullptr, nlevels, nsteps, save_steps, B, coord_mode, stream);
}

extern "C" int pulpo_combine_vecint_multi_fwd(const pulpo_vecint_level *levels, const float *const *indiv, int nlevels,
                                              int nsteps, int save_steps, int B, int coord_mode, pulpo_stream_t stream)
{
    PULPO_REQUIRE(indiv, PULPO_ERR_NULL_POINTER);
    return vecint_multi_fwd_impl(levels, indiv, nlevels, nsteps, save_steps, B, coord_mode, stream);
}

static int vecint_multi_bwd_impl(const pulpo_vecint_level *levels, int combine, int nlevels, int nsteps, int B,
                                 int coord_mode, pulpo_stream_t stream)
{
    PULPO_REQUIRE(B > 0 && nsteps >= 0 && nsteps <= 30, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(levels && nlevels >= 1 && nlevels <= VI_MAXL, PULPO_ERR_INVALID_SHAPE);
    if (split_batch(levels, nlevels, B)) {
        pulpo_vecint_level one[VI_MAXL];
        for (int b = 0; b < B; ++b) {
            item_levels(one, levels, nlevels, b, nsteps, 1, true);
            int rc = vecint_multi_bwd_impl(one, combine, nlevels, nsteps, 1, coord_mode, stream);
            if (rc != PULPO_OK) return rc;
        }
        return PULPO_OK;
    }
    const int variant = (coord_mode >> 8) & 0xf;   // tuning switch, see below; 0 = default
    coord_mode &= 0xff;
    PULPO_REQUIRE(coord_mode >= 0 && coord_mode <= 2, PULPO_ERR_UNSUPPORTED);
    VMulti m;
    int rc = fill_levels(m, levels, nlevels, B, true, nsteps, 1);
    if (rc != PULPO_OK) return rc;
    if (combine) {
        rc = check_pyramid(levels, nlevels);
        if (rc != PULPO_OK) return rc;
        m.combine = 1;
        m.dataflow = 0;
    }
    float scale = 1.0f / (float)(1u << nsteps);
    void *args[] = {&m, &nsteps, &scale};
    cudaStream_t st = (cudaStream_t)stream;
    // scatter strategy (same result up to summation order): default z-carry (2); 0x100 -> one reduction per corner
    // (0), 0x200 -> lane + plane combining (1)
    PULPO_REQUIRE(variant <= 2, PULPO_ERR_UNSUPPORTED);
    const int comb = variant == 0 ? 2 : variant == 1 ? 0 : 1;
#ifdef PULPO_VI_BWD_XYZ
    const size_t bwd_smem = (size_t)(VI_BWD_THREADS / 32) * VI_RING * 64 * sizeof(float4);   // per-warp rings of own values
#else
    const size_t bwd_smem = 0;
#endif
#define PULPO_VI_BWD_CASE(M, C) \
    if (coord_mode == M && comb == C) return launch_coop(vecint_bwd_kernel<M, C>, VI_BWD_THREADS, m, args, st, bwd_smem);
    PULPO_VI_BWD_CASE(0, 2) PULPO_VI_BWD_CASE(1, 2) PULPO_VI_BWD_CASE(2, 2)
    PULPO_VI_BWD_CASE(0, 0) PULPO_VI_BWD_CASE(2, 0)
    PULPO_VI_BWD_CASE(0, 1) PULPO_VI_BWD_CASE(2, 1)
#undef PULPO_VI_BWD_CASE
    return PULPO_ERR_UNSUPPORTED;
}

extern "C" int pulpo_vecint_multi_bwd(const pulpo_vecint_level *levels, int nlevels, int nsteps, int B, int coord_mode,
                                      pulpo_stream_t stream)
{
    return vecint_multi_bwd_impl(levels, 0, nlevels, nsteps, B, coord_mode, stream);
}

extern "C" int pulpo_combine_vecint_multi_bwd(const pulpo_vecint_level *levels, int nlevels, int nsteps, int B,
                                              int coord_mode, pulpo_stream_t stream)
{
    return vecint_multi_bwd_impl(levels, 1, nlevels, nsteps, B, coord_mode, stream);
}

extern "C" int pulpo_vecint_fwd(const float *vec, float *out, void *ws, size_t ws_bytes, int nsteps, int save_steps,
                                int B, int D0, int D1, int D2, int coord_mode, pulpo_stream_t stream)
{
    pulpo_vecint_level lv = {vec, out, ws, ws_bytes, nullptr, 0, D0, D1, D2};
    return pulpo_vecint_multi_fwd(&lv, 1, nsteps, save_steps, B, coord_mode, stream);
}

extern "C" int pulpo_vecint_bwd(const float *gout, const void *saved, float *gvec, void *scratch, size_t scratch_bytes,
                                int nsteps, int B, int D0, int D1, int D2, int coord_mode, pulpo_stream_t stream)
{
    PULPO_REQUIRE(saved || nsteps == 0 || !gout, PULPO_ERR_NULL_POINTER);
    pulpo_vecint_level lv = {gout, gvec, const_cast<void *>(saved), 0, scratch, scratch_bytes, D0, D1, D2};
    if (!saved && nsteps == 0) lv.ws = scratch;   // unused by the kernel when there are no steps
    return pulpo_vecint_multi_bwd(&lv, 1, nsteps, B, coord_mode, stream);
}
