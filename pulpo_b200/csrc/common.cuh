// common.cuh -- shared device helpers for libpulpo_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <nvtx3/nvToolsExt.h>

#include "../../include/pulpo_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpulpo_b200 is written for sm_100a (B200) only"
#endif

namespace pulpo {

typedef long long i64;

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

// NVTX range around every C-ABI entry point (the reference has no tracing hooks; Nsight shows one range
// per call).  Header-only NVTX v3: a no-op costing a few nanoseconds when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define PULPO_NVTX(name) ::pulpo::NvtxRange pulpo_nvtx_range_(name)

#define PULPO_REQUIRE(cond, code) \
    do {                          \
        if (!(cond)) return (code); \
    } while (0)

static inline int launch_status()
{
    cudaError_t e = cudaPeekAtLastError();
    if (e == cudaSuccess) return PULPO_OK;
    if (getenv("PULPO_B200_DEBUG")) fprintf(stderr, "libpulpo_b200: CUDA error: %s\n", cudaGetErrorString(e));
    return PULPO_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// Sample position of SpatialTransformer (src/network_blocks.py:103-107 + ATen unnormalise for
// align_corners=False).  Every op is rounded separately with _rn intrinsics so the compiler can
// neither contract to FMA nor replace the division by a reciprocal multiply: floor(p) must match
// torch bit for bit.  MODE 1 reproduces what torch-CUDA does instead.
struct AxisConst {
    float S;     // float(S)
    float Sm1;   // float(S-1)
    float rcp;   // RN(1 / float(S-1))
    float gmul;  // S/2: d p / d n of the unnormalise (backward)
    float tmax;  // 2^23 + (S-2): largest low-corner index, biased (see make_tap)
    float kf;    // RN(S / (S-1)): d p / d loc; the whole normalise/unnormalise chain in one factor (MODE 2)
};

__host__ __device__ inline AxisConst make_axis(int S)
{
    AxisConst a;
    a.S = (float)S;
    a.Sm1 = (float)(S - 1);
    a.rcp = 1.0f / (float)(S - 1);
    a.gmul = (float)S / 2.0f;
    a.tmax = 8388608.0f + (float)(S - 2);
    a.kf = (float)S / (float)(S - 1);
    return a;
}

// Correctly rounded x / c for the per-axis constant c = S-1 without the generic division
// sequence: y = RN(1/c) is precomputed, two FMA residual corrections make q1 faithful and the
// last FMA rounds correctly (Markstein's theorem; it only fails for a divisor whose significand
// is all ones, i.e. S-1 = 2^24-1).  Non-finite x is passed through.  Tiny |x| may lose
// denormal bits, which the following "- 0.5" absorbs.
__device__ __forceinline__ float div_by_axis(float x, const AxisConst &a)
{
    float q0 = __fmul_rn(x, a.rcp);
    float r0 = __fmaf_rn(-a.Sm1, q0, x);
    float q1 = __fmaf_rn(r0, a.rcp, q0);
    float r1 = __fmaf_rn(-a.Sm1, q1, x);
    float q = __fmaf_rn(r1, a.rcp, q1);
    // non-finite or absurdly large x: x itself has the right sign/NaN-ness and clamps identically
    return (fabsf(x) <= 1e30f) ? q : x;
}

// vf = float(voxel index along this axis)
template <int MODE>
__device__ __forceinline__ float sample_pos(float vf, float d, const AxisConst &a)
{
    float loc = __fadd_rn(vf, d);
    // MODE 2 (PULPO_COORD_FAST): p = loc * S/(S-1) - 0.5 in one FMA.  Same value as the chain below
    // up to a few ulp of p (~1e-5 voxel at p ~ 200); not index-exact, so only VecInt offers it.
    if (MODE == PULPO_COORD_FAST) return __fmaf_rn(loc, a.kf, -0.5f);
    float q = (MODE == PULPO_COORD_CPU_EXACT) ? div_by_axis(loc, a) : __fmul_rn(loc, a.rcp);
    float n = __fmul_rn(2.0f, __fsub_rn(q, 0.5f));
    float t = (MODE == PULPO_COORD_CPU_EXACT) ? __fsub_rn(__fmul_rn(__fadd_rn(n, 1.0f), a.S), 1.0f)
                                              : __fmaf_rn(__fadd_rn(n, 1.0f), a.S, -1.0f);
    return __fmul_rn(t, 0.5f);  // "/ 2" is exact either way
}

// border padding: std::min(S-1, std::max(p, 0)) including its NaN behaviour (NaN -> S-1) in two
// instructions: NaN-propagating max, then IEEE minNum
__device__ __forceinline__ float clip_pos(float p, float Sm1)
{
    float lo;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(lo) : "f"(p), "f"(0.0f));
    return fminf(lo, Sm1);
}

constexpr int kTapBias = 0x4B000000;   // float bits of 2^23

// one axis of the trilinear footprint
struct Tap {
    int bits;   // float bits of 2^23 + i;  i = index of the low corner actually read
    float w0;   // weight of corner i
    float w1;   // weight of corner i+1
    int floor_p;  // floor(p) (== i except at the upper border); only the index-dump path uses it
};

template <int MODE>
__device__ __forceinline__ Tap make_tap(float vf, float d, const AxisConst &a, float *unclamped = nullptr)
{
    float u = sample_pos<MODE>(vf, d, a);
    if (unclamped) *unclamped = u;
    float p = clip_pos(u, a.Sm1);
    // floor of 0 <= p < 2^22 without the conversion unit: adding 2^23 in round-toward-zero
    // truncates the fraction; the integer sits in the low mantissa bits of t.
    float t = __fadd_rz(p, 8388608.0f);
    // p == S-1 exactly: corner i+1 would be outside the volume (weight 0).  Clamp the low corner to
    // S-2 so that every footprint is 2x2x2 in-bounds with fixed +1 neighbours; the weights become
    // (0, 1) and 0*x + 1*y == y keeps the result bit-identical.
    float tc = fminf(t, a.tmax);
    float fl = __fsub_rn(tc, 8388608.0f);
    Tap tp;
    tp.bits = __float_as_int(tc);
    tp.floor_p = __float_as_int(t) - kTapBias;
    tp.w1 = __fsub_rn(p, fl);          // p - i   (exact)
    tp.w0 = __fsub_rn(1.0f, tp.w1);    // == (i+1) - p bit for bit (both differences are exact or the same op)
    return tp;
}

// ---------------------------------------------------------------------------------------------
// Packed (f32x2) variant: two voxels per instruction on Blackwell's FADD2 / FMUL2 / FFMA2.  The warp and
// integration kernels are bound by instruction issue, and the sample-position chain is most of their
// arithmetic; two voxels of the same axis share every constant, so the chain packs perfectly.  Same values,
// bit for bit, as sample_pos / make_tap above -- two steps of the chain are fused where the fusion is exact:
//   * n + 1 with n = 2 * (q - 0.5):  n is exact (a power-of-two scaling), so RN(n + 1) == fma(q - 0.5, 2, 1);
//   * (t - 1) * 0.5 with t = RN((n + 1) * S):  halving commutes with rounding (t - 1 is never subnormal: it is 0
//     or at least one ulp of a value near 1), so RN(t - 1) / 2 == fma(t, 0.5, -0.5).
struct AxisConst2 {   // AxisConst as register pairs: the same axis twice (two voxels) or two axes of one voxel
    float2 S, nSm1, rcp, Sm1, tmax, gmul, kf;
};

// S: size of the axis the displacement field lives on (normalisation, src/network_blocks.py:106-107); Simg: size of
// the sampled image along that axis (grid_sample's unnormalise, clamp and gather).  They differ when a level-sized
// field resamples a full-resolution image (evaluate.py:198,240,246).
__host__ __device__ inline AxisConst2 make_axis_pair(int Sx, int Simgx, int Sy, int Simgy)
{
    const AxisConst ax = make_axis(Sx), ix = make_axis(Simgx), ay = make_axis(Sy), iy = make_axis(Simgy);
    AxisConst2 r;
    r.S.x = ix.S; r.S.y = iy.S;
    r.nSm1.x = -ax.Sm1; r.nSm1.y = -ay.Sm1;
    r.rcp.x = ax.rcp; r.rcp.y = ay.rcp;
    r.Sm1.x = ix.Sm1; r.Sm1.y = iy.Sm1;
    r.tmax.x = ix.tmax; r.tmax.y = iy.tmax;
    r.gmul.x = ix.gmul; r.gmul.y = iy.gmul;
    r.kf.x = ix.S / ax.Sm1; r.kf.y = iy.S / ay.Sm1;
    return r;
}
__host__ __device__ inline AxisConst2 make_axis2(int S, int Simg) { return make_axis_pair(S, Simg, S, Simg); }
__host__ __device__ inline AxisConst2 make_axis2(int S) { return make_axis2(S, S); }

__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

// a + b, both halves rounded to nearest, that ptxas does NOT contract with a preceding packed multiply: ptxas 12.9
// fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (even with -fmad=false), which would break bit-identity with the CPU
// sampler's separately rounded products and sums.  The .ftz flag on the add alone blocks the fusion; it only matters
// for subnormal operands / results (|x| < 1.2e-38), which the interpolation of image intensities never produces
// unless the exact value is already that small.
__device__ __forceinline__ float2 add2_nofuse(float2 a, float2 b)
{
    unsigned long long ua, ub, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
    asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(ua), "l"(ub));
    float2 o;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
    return o;
}

// non-finite / absurd loc: NaN and +huge must end at S-1, -huge at 0, like the division-based chain does; clamping
// loc itself (IEEE minNum drops the NaN in favour of the bound) keeps the FMA chain below finite
__device__ __forceinline__ float tame(float x) { return fmaxf(fminf(x, 1e30f), -1e30f); }

template <int MODE>
__device__ __forceinline__ float2 sample_pos2(float2 vf, float2 d, const AxisConst2 &a)
{
    float2 loc = __fadd2_rn(vf, d);
    if (MODE == PULPO_COORD_FAST) return __ffma2_rn(loc, a.kf, splat2(-0.5f));
    float2 q;
    if (MODE == PULPO_COORD_CPU_EXACT) {
        loc.x = tame(loc.x); loc.y = tame(loc.y);
        const float2 q0 = __fmul2_rn(loc, a.rcp);
        const float2 r0 = __ffma2_rn(a.nSm1, q0, loc);
        const float2 q1 = __ffma2_rn(r0, a.rcp, q0);
        const float2 r1 = __ffma2_rn(a.nSm1, q1, loc);
        q = __ffma2_rn(r1, a.rcp, q1);
    } else {
        q = __fmul2_rn(loc, a.rcp);
    }
    const float2 h = __fadd2_rn(q, splat2(-0.5f));
    const float2 n1 = __ffma2_rn(h, splat2(2.0f), splat2(1.0f));          // == RN(2 * (q - 0.5) + 1)
    if (MODE == PULPO_COORD_CPU_EXACT) {
        const float2 t = __fmul2_rn(n1, a.S);
        return __ffma2_rn(t, splat2(0.5f), splat2(-0.5f));               // == RN(t - 1) * 0.5
    }
    const float2 t = __ffma2_rn(n1, a.S, splat2(-1.0f));                   // torch-CUDA contracts (n + 1) * S - 1
    return __fmul2_rn(t, splat2(0.5f));
}

struct Tap2 {
    int bits0, bits1;      // float bits of 2^23 + i for the two voxels
    float2 w0, w1;         // weights of corners i and i+1
    float2 u;              // unclamped sample positions (backward: gradient mask)
    int fl0, fl1;          // floor(p) (index dump)
};

template <int MODE>
__device__ __forceinline__ Tap2 make_tap2(float2 vf, float2 d, const AxisConst2 &a)
{
    Tap2 tp;
    tp.u = sample_pos2<MODE>(vf, d, a);
    float2 p;
    p.x = clip_pos(tp.u.x, a.Sm1.x); p.y = clip_pos(tp.u.y, a.Sm1.y);
    const float2 t = __fadd2_rz(p, splat2(8388608.0f));
    float2 tc;
    tc.x = fminf(t.x, a.tmax.x); tc.y = fminf(t.y, a.tmax.y);
    const float2 fl = __fadd2_rn(tc, splat2(-8388608.0f));
    tp.bits0 = __float_as_int(tc.x); tp.bits1 = __float_as_int(tc.y);
    tp.fl0 = __float_as_int(t.x) - kTapBias; tp.fl1 = __float_as_int(t.y) - kTapBias;
    tp.w1 = __ffma2_rn(fl, splat2(-1.0f), p);              // p - i   (exact)
    tp.w0 = __ffma2_rn(tp.w1, splat2(-1.0f), splat2(1.0f));   // == (i+1) - p bit for bit
    return tp;
}

// offset of the low corner inside one [D0,D1,D2] volume from the three taps' float bits; the
// 2^23 biases are removed by one precomputed constant (32-bit wrap-around arithmetic)
__device__ __forceinline__ int tap_base(const Tap &tz, const Tap &ty, const Tap &tx, int D1, int D2, int unbias)
{
    return (tz.bits * D1 + ty.bits) * D2 + tx.bits - unbias;
}

static inline int tap_unbias(int D1, int D2)
{
    return (int)((unsigned int)kTapBias * ((unsigned int)D1 * (unsigned int)D2 + (unsigned int)D2 + 1u));
}

// 32-bit division by a launch constant (q = umulhi(n, mul) >> shr, valid for n < 2^31)
struct FastDiv {
    unsigned int d, mul, shr;
};

static inline FastDiv make_fastdiv(unsigned int d)
{
    FastDiv f;
    f.d = d;
    if (d == 1) {
        f.mul = 0; f.shr = 0;
        return f;
    }
    unsigned int lg = 0;
    while ((1ull << lg) < d) ++lg;
    unsigned int p = 31 + lg;
    f.mul = (unsigned int)(((1ull << p) + d - 1) / d);
    f.shr = p - 32;
    return f;
}

__device__ __forceinline__ void fast_divmod(unsigned int n, const FastDiv &f, unsigned int &q, unsigned int &r)
{
    q = (f.d == 1) ? n : (__umulhi(n, f.mul) >> f.shr);
    r = n - q * f.d;
}

// ---------------------------------------------------------------------------------------------
// reductions
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum in double (result valid in thread 0).  smem: >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double *smem)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    v = warp_sum(v);
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = (lane < (nthreads + 31) / 32) ? smem[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// Deterministic two-stage scalar reduction: every CTA deposits one double partial in the
// caller's workspace; the last CTA to arrive (ticket counter, self-resetting) sums the
// partials in a fixed order and writes `scale * sum` to *out as fp32.
struct ReduceWs {
    unsigned int ticket;
    unsigned int pad;
    double acc;         // accumulator of the atomic variant (grid_reduce_finish_atomic)
    double partial[1];  // [max_ctas]
};
constexpr int kMaxReduceCtas = 4096;
constexpr size_t kReduceWsBytes = 16 + sizeof(double) * kMaxReduceCtas;

__device__ __forceinline__ void grid_reduce_finish(double block_total, ReduceWs *ws, float *out, double scale,
                                                   double *smem)
{
    __shared__ bool is_last;
    const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    const int nctas = gridDim.x * gridDim.y * gridDim.z;
    if (threadIdx.x == 0) {
        ws->partial[cta] = block_total;
        __threadfence();
        unsigned int t = atomicAdd(&ws->ticket, 1u);
        is_last = (t == (unsigned int)(nctas - 1));
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double s = 0.0;
        for (int i = threadIdx.x; i < nctas; i += blockDim.x) s += ((volatile double *)ws->partial)[i];
        s = block_sum(s, smem);
        if (threadIdx.x == 0) {
            *out = (float)(s * scale);
            ws->ticket = 0;  // ready for the next launch on this workspace
        }
    }
}

// Variant for grids of any size: CTAs add their double partial into ws->acc with one fp64
// atomic; the last CTA to arrive publishes and resets.  The summation order is not fixed, but
// the partials are doubles of fp32 data, so the fp32 result is reproducible in practice.
__device__ __forceinline__ void grid_reduce_finish_atomic(double block_total, ReduceWs *ws, float *out, double scale)
{
    if (threadIdx.x == 0) {
        const unsigned int nctas = gridDim.x * gridDim.y * gridDim.z;
        atomicAdd(&ws->acc, block_total);
        __threadfence();
        unsigned int t = atomicAdd(&ws->ticket, 1u);
        if (t == nctas - 1) {
            __threadfence();
            double s = __longlong_as_double(atomicExch((unsigned long long *)&ws->acc, 0ull));
            *out = (float)(s * scale);
            ws->ticket = 0;
        }
    }
}

// streaming (read-once) 128-bit load that does not pollute L1
__device__ __forceinline__ float4 ld_stream4(const float *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float ld_stream(const float *p)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// request a line into L2 ahead of its use (no register, no shared memory: for z-marching kernels whose next planes'
// addresses are known long before the loads that need them)
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// sm_90+ vector reduction: one 16-byte red instead of four scalar atomics
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

static inline bool aligned16(const void *p) { return (((uintptr_t)p) & 15u) == 0; }

static inline int grid_for(i64 work_items, int threads, int max_ctas_per_sm = 8)
{
    i64 g = (work_items + threads - 1) / threads;
    i64 cap = (i64)kSMs * max_ctas_per_sm;
    if (g > cap) g = cap;  // grid-stride loops: whole waves of 148 SMs
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace pulpo
