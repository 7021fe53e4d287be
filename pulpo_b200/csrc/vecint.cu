// vecint.cu -- scaling-and-squaring integration (reference: VecInt, src/network_blocks.py:165-177):
//     v <- vec * 2^-nsteps ;  nsteps times:  v <- v + warp(v, v)
//
// B200 design.  The field is both the image and the displacement, so each step is a full
// grid-wide dependency.  All steps run in ONE cooperative launch with grid.sync() between
// steps instead of 7 launches; the integration states live in a channel-interleaved float4
// layout ([B,S] x (c0,c1,c2,pad)) so that every trilinear corner is one 128-bit gather
// instead of three 32-bit ones, and the scatter half of the backward is one
// red.global.add.v4.f32 per corner instead of three scalar atomics.  The states of one level
// (<= 13.8 MB at 80x96x112) stay resident in the 126 MB L2 between steps.
// Forward arithmetic is op-for-op the CPU grid sampler's, so results are bit-identical to
// torch-CPU in PULPO_COORD_CPU_EXACT mode.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pulpo {

struct VGeom {
    int B, D0, D1, D2;
    int S;            // voxels per volume (< 2^31)
    unsigned int N;   // B * S
    FastDiv dD2, dD1, dD0;
    AxisConst a0, a1, a2;
};

static int make_vgeom(VGeom &g, int B, int D0, int D1, int D2)
{
    i64 S = (i64)D0 * D1 * D2;
    if ((i64)B * S >= (1ll << 31) || D0 > (1 << 22) || D1 > (1 << 22) || D2 > (1 << 22)) return PULPO_ERR_INVALID_SHAPE;
    g.B = B; g.D0 = D0; g.D1 = D1; g.D2 = D2; g.S = (int)S; g.N = (unsigned int)(B * S);
    g.dD2 = make_fastdiv(D2); g.dD1 = make_fastdiv(D1); g.dD0 = make_fastdiv(D0);
    g.a0 = make_axis(D0); g.a1 = make_axis(D1); g.a2 = make_axis(D2);
    return PULPO_OK;
}

__device__ __forceinline__ void decode(unsigned int i, const VGeom &g, unsigned int &b, unsigned int &z,
                                       unsigned int &y, unsigned int &x)
{
    unsigned int r, zb;
    fast_divmod(i, g.dD2, r, x);
    fast_divmod(r, g.dD1, zb, y);
    fast_divmod(zb, g.dD0, b, z);
}

struct Corners {
    float4 c[8];  // index = dz*4 + dy*2 + dx
};

// the footprint is always 2x2x2 in-bounds (see make_tap): fixed +1 / +D2 / +D1*D2 neighbours
__device__ __forceinline__ void gather8(const float4 *p, int sy, int sz, Corners &k)
{
    const float4 *py = p + sy, *pz = p + sz, *pzy = pz + sy;
    k.c[0] = p[0];
    k.c[1] = p[1];
    k.c[2] = py[0];
    k.c[3] = py[1];
    k.c[4] = pz[0];
    k.c[5] = pz[1];
    k.c[6] = pzy[0];
    k.c[7] = pzy[1];
}

// same corner order / op order as the CPU grid sampler (tnw, tne, tsw, tse, bnw, ...)
__device__ __forceinline__ float interp_exact(float c0, float c1, float c2, float c3, float c4, float c5, float c6,
                                              float c7, const float w[8])
{
    float acc = __fmul_rn(c0, w[0]);
    acc = __fadd_rn(acc, __fmul_rn(c1, w[1]));
    acc = __fadd_rn(acc, __fmul_rn(c2, w[2]));
    acc = __fadd_rn(acc, __fmul_rn(c3, w[3]));
    acc = __fadd_rn(acc, __fmul_rn(c4, w[4]));
    acc = __fadd_rn(acc, __fmul_rn(c5, w[5]));
    acc = __fadd_rn(acc, __fmul_rn(c6, w[6]));
    acc = __fadd_rn(acc, __fmul_rn(c7, w[7]));
    return acc;
}

struct VFoot {
    int base;
    float w[8];
    float wx0, wx1, wy0, wy1, wz0, wz1;
};

template <int MODE>
__device__ __forceinline__ VFoot make_vfoot(unsigned int b, unsigned int z, unsigned int y, unsigned int x,
                                            const float4 &v, const VGeom &g, float *uz = nullptr,
                                            float *uy = nullptr, float *ux = nullptr)
{
    Tap tz = make_tap<MODE>((float)(int)z, v.x, g.a0, g.D0, uz);
    Tap ty = make_tap<MODE>((float)(int)y, v.y, g.a1, g.D1, uy);
    Tap tx = make_tap<MODE>((float)(int)x, v.z, g.a2, g.D2, ux);
    VFoot f;
    f.base = (int)b * g.S + (tz.i * g.D1 + ty.i) * g.D2 + tx.i;
    f.wx0 = tx.w0; f.wx1 = tx.w1; f.wy0 = ty.w0; f.wy1 = ty.w1; f.wz0 = tz.w0; f.wz1 = tz.w1;
    const float w00 = __fmul_rn(tx.w0, ty.w0), w01 = __fmul_rn(tx.w1, ty.w0);
    const float w10 = __fmul_rn(tx.w0, ty.w1), w11 = __fmul_rn(tx.w1, ty.w1);
    f.w[0] = __fmul_rn(w00, tz.w0); f.w[1] = __fmul_rn(w01, tz.w0);
    f.w[2] = __fmul_rn(w10, tz.w0); f.w[3] = __fmul_rn(w11, tz.w0);
    f.w[4] = __fmul_rn(w00, tz.w1); f.w[5] = __fmul_rn(w01, tz.w1);
    f.w[6] = __fmul_rn(w10, tz.w1); f.w[7] = __fmul_rn(w11, tz.w1);
    return f;
}

template <int MODE>
__global__ void __launch_bounds__(256)
vecint_fwd_kernel(const float *__restrict__ vec, float *__restrict__ out, float4 *ws, int nsteps, int save,
                  float scale, const VGeom g)
{
    cg::grid_group grid = cg::this_grid();
    const unsigned int N = g.N, S = g.S;
    const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;

    // v_0 = vec * 2^-nsteps, planar -> interleaved
    for (unsigned int i = tid; i < N; i += nthr) {
        unsigned int b = i / S, v = i - b * S;
        const float *f = vec + (i64)b * 3 * S + v;
        ws[i] = make_float4(__fmul_rn(__ldg(f), scale), __fmul_rn(__ldg(f + S), scale),
                            __fmul_rn(__ldg(f + 2 * S), scale), 0.0f);
    }
    for (int k = 0; k < nsteps; ++k) {
        grid.sync();
        const float4 *src = save ? ws + (i64)k * N : ws + (i64)(k & 1) * N;
        float4 *dst = save ? ws + (i64)(k + 1) * N : ws + (i64)((k + 1) & 1) * N;
        const bool last = (k == nsteps - 1);
        for (unsigned int i = tid; i < N; i += nthr) {
            unsigned int b, z, y, x;
            decode(i, g, b, z, y, x);
            const float4 v = src[i];
            const VFoot f = make_vfoot<MODE>(b, z, y, x, v, g);
            Corners kc;
            gather8(src + f.base, g.D2, g.D1 * g.D2, kc);
            float r0 = interp_exact(kc.c[0].x, kc.c[1].x, kc.c[2].x, kc.c[3].x, kc.c[4].x, kc.c[5].x, kc.c[6].x, kc.c[7].x, f.w);
            float r1 = interp_exact(kc.c[0].y, kc.c[1].y, kc.c[2].y, kc.c[3].y, kc.c[4].y, kc.c[5].y, kc.c[6].y, kc.c[7].y, f.w);
            float r2 = interp_exact(kc.c[0].z, kc.c[1].z, kc.c[2].z, kc.c[3].z, kc.c[4].z, kc.c[5].z, kc.c[6].z, kc.c[7].z, f.w);
            r0 = __fadd_rn(v.x, r0);
            r1 = __fadd_rn(v.y, r1);
            r2 = __fadd_rn(v.z, r2);
            if (last) {
                float *o = out + (i64)b * 3 * S + (i - b * S);
                o[0] = r0; o[S] = r1; o[2 * S] = r2;
            } else {
                dst[i] = make_float4(r0, r1, r2, 0.0f);
            }
        }
    }
    if (nsteps == 0) {
        grid.sync();
        for (unsigned int i = tid; i < N; i += nthr) {
            unsigned int b = i / S, v = i - b * S;
            float4 t = ws[i];
            float *o = out + (i64)b * 3 * S + v;
            o[0] = t.x; o[S] = t.y; o[2 * S] = t.z;
        }
    }
}

// Backward of one step  v' = v + W(v) v :   g = g' + W(v)^T g' + (dW/dv : v)^T g'
//   own   : g' + gather-form gradient through the sample position  -> one red.v4 on the voxel
//   scatter: w_d * g' onto the 8 corners                            -> one red.v4 per corner
// All contributions go through red.global.add.v4.f32 into a pre-zeroed state, so there is no
// ordering hazard between the plain part and the scatter part; three states rotate
// (read / accumulate / being zeroed for the next step).
template <int MODE>
__global__ void __launch_bounds__(256)
vecint_bwd_kernel(const float *__restrict__ gout, const float4 *saved, float *__restrict__ gvec, float4 *scr,
                  int nsteps, float scale, const VGeom g)
{
    cg::grid_group grid = cg::this_grid();
    const unsigned int N = g.N, S = g.S;
    const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    float4 *X = scr, *Y = scr + N, *Z = scr + 2 * (i64)N;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float kz = 2.0f * g.a0.rcp, ky = 2.0f * g.a1.rcp, kx = 2.0f * g.a2.rcp;

    for (unsigned int i = tid; i < N; i += nthr) {
        unsigned int b = i / S, v = i - b * S;
        const float *f = gout + (i64)b * 3 * S + v;
        X[i] = make_float4(__ldg(f), __ldg(f + S), __ldg(f + 2 * S), 0.0f);
        Y[i] = zero4;
    }
    for (int k = nsteps - 1; k >= 0; --k) {
        grid.sync();
        const float4 *vk = saved + (i64)k * N;
        for (unsigned int i = tid; i < N; i += nthr) {
            unsigned int b, z, y, x;
            decode(i, g, b, z, y, x);
            const float4 G = X[i];
            const float4 v = vk[i];
            float uz, uy, ux;
            const VFoot f = make_vfoot<MODE>(b, z, y, x, v, g, &uz, &uy, &ux);
            Corners kc;
            gather8(vk + f.base, g.D2, g.D1 * g.D2, kc);
            // t[d] = <corner_d, G> over the 3 channels
            float t[8];
#pragma unroll
            for (int d = 0; d < 8; ++d) t[d] = kc.c[d].x * G.x + kc.c[d].y * G.y + kc.c[d].z * G.z;
            const float sx = ((t[1] - t[0]) * f.wy0 + (t[3] - t[2]) * f.wy1) * f.wz0 +
                             ((t[5] - t[4]) * f.wy0 + (t[7] - t[6]) * f.wy1) * f.wz1;
            const float sy_ = ((t[2] - t[0]) * f.wx0 + (t[3] - t[1]) * f.wx1) * f.wz0 +
                             ((t[6] - t[4]) * f.wx0 + (t[7] - t[5]) * f.wx1) * f.wz1;
            const float sz = ((t[4] - t[0]) * f.wx0 + (t[5] - t[1]) * f.wx1) * f.wy0 +
                             ((t[6] - t[2]) * f.wx0 + (t[7] - t[3]) * f.wx1) * f.wy1;
            const float mz = (uz <= 0.0f || uz >= g.a0.Sm1) ? 0.0f : g.a0.gmul;
            const float my = (uy <= 0.0f || uy >= g.a1.Sm1) ? 0.0f : g.a1.gmul;
            const float mx = (ux <= 0.0f || ux >= g.a2.Sm1) ? 0.0f : g.a2.gmul;
            red_add_v4(reinterpret_cast<float *>(Y + i), G.x + (mz * sz) * kz, G.y + (my * sy_) * ky,
                       G.z + (mx * sx) * kx, 0.0f);
            // scatter half (all 8 corners are in-bounds; border corners carry weight 0)
            float *q = reinterpret_cast<float *>(Y + f.base);
            const int sy4 = 4 * g.D2, sz4 = 4 * g.D1 * g.D2;
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const int o = ((d & 1) ? 4 : 0) + ((d & 2) ? sy4 : 0) + ((d & 4) ? sz4 : 0);
                const float w = f.w[d];
                red_add_v4(q + o, w * G.x, w * G.y, w * G.z, 0.0f);
            }
            Z[i] = zero4;  // accumulation target of the next step
        }
        float4 *t = X; X = Y; Y = Z; Z = t;
    }
    grid.sync();
    for (unsigned int i = tid; i < N; i += nthr) {
        unsigned int b = i / S, v = i - b * S;
        float4 gq = X[i];
        float *o = gvec + (i64)b * 3 * S + v;
        o[0] = gq.x * scale; o[S] = gq.y * scale; o[2 * S] = gq.z * scale;
    }
}

template <typename K>
static int coop_grid(K kernel, i64 work, int threads)
{
    int dev = 0, sms = kSMs, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
    if (per_sm < 1) per_sm = 1;
    i64 g = (work + threads - 1) / threads;
    i64 cap = (i64)sms * per_sm;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace pulpo

using namespace pulpo;

extern "C" size_t pulpo_vecint_ws_bytes(int nsteps, int save_steps, int B, int D0, int D1, int D2)
{
    size_t state = (size_t)B * D0 * D1 * D2 * sizeof(float4);
    int n = save_steps ? (nsteps < 1 ? 1 : nsteps) : 2;
    return state * (size_t)n;
}

extern "C" size_t pulpo_vecint_bwd_scratch_bytes(int B, int D0, int D1, int D2)
{
    return (size_t)B * D0 * D1 * D2 * sizeof(float4) * 3;
}

extern "C" int pulpo_vecint_fwd(const float *vec, float *out, void *ws, size_t ws_bytes, int nsteps, int save_steps,
                                int B, int D0, int D1, int D2, int coord_mode, pulpo_stream_t stream)
{
    PULPO_REQUIRE(vec && out && ws, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2 && nsteps >= 0 && nsteps <= 30, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(coord_mode == 0 || coord_mode == 1, PULPO_ERR_UNSUPPORTED);
    PULPO_REQUIRE(ws_bytes >= pulpo_vecint_ws_bytes(nsteps, save_steps, B, D0, D1, D2) && aligned16(ws),
                  PULPO_ERR_WORKSPACE);
    VGeom g;
    int rc = make_vgeom(g, B, D0, D1, D2);
    if (rc != PULPO_OK) return rc;
    float scale = 1.0f / (float)(1u << nsteps);
    float4 *w4 = (float4 *)ws;
    void *args[] = {&vec, &out, &w4, &nsteps, &save_steps, &scale, &g};
    cudaError_t e;
    if (coord_mode == 0) {
        int grid = coop_grid(vecint_fwd_kernel<0>, g.N, 256);
        e = cudaLaunchCooperativeKernel((void *)vecint_fwd_kernel<0>, dim3(grid), dim3(256), args, 0, (cudaStream_t)stream);
    } else {
        int grid = coop_grid(vecint_fwd_kernel<1>, g.N, 256);
        e = cudaLaunchCooperativeKernel((void *)vecint_fwd_kernel<1>, dim3(grid), dim3(256), args, 0, (cudaStream_t)stream);
    }
    return e == cudaSuccess ? launch_status() : PULPO_ERR_CUDA;
}

extern "C" int pulpo_vecint_bwd(const float *gout, const void *saved, float *gvec, void *scratch, size_t scratch_bytes,
                                int nsteps, int B, int D0, int D1, int D2, int coord_mode, pulpo_stream_t stream)
{
    PULPO_REQUIRE(gout && gvec && scratch && (saved || nsteps == 0), PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && D0 >= 2 && D1 >= 2 && D2 >= 2 && nsteps >= 0 && nsteps <= 30, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(coord_mode == 0 || coord_mode == 1, PULPO_ERR_UNSUPPORTED);
    PULPO_REQUIRE(scratch_bytes >= pulpo_vecint_bwd_scratch_bytes(B, D0, D1, D2) && aligned16(scratch) &&
                      aligned16(saved),
                  PULPO_ERR_WORKSPACE);
    VGeom g;
    int rc = make_vgeom(g, B, D0, D1, D2);
    if (rc != PULPO_OK) return rc;
    float scale = 1.0f / (float)(1u << nsteps);
    const float4 *sv = (const float4 *)saved;
    float4 *scr = (float4 *)scratch;
    void *args[] = {&gout, &sv, &gvec, &scr, &nsteps, &scale, &g};
    cudaError_t e;
    if (coord_mode == 0) {
        int grid = coop_grid(vecint_bwd_kernel<0>, g.N, 256);
        e = cudaLaunchCooperativeKernel((void *)vecint_bwd_kernel<0>, dim3(grid), dim3(256), args, 0, (cudaStream_t)stream);
    } else {
        int grid = coop_grid(vecint_bwd_kernel<1>, g.N, 256);
        e = cudaLaunchCooperativeKernel((void *)vecint_bwd_kernel<1>, dim3(grid), dim3(256), args, 0, (cudaStream_t)stream);
    }
    return e == cudaSuccess ? launch_status() : PULPO_ERR_CUDA;
}
