// jacdet.cu -- Jacobian determinant of a displacement field and its standard deviation
// (reference: jacobian_det / JDetStd, src/losses.py:147-204, 3-D branch): the alternative regulariser
// (`--regularizer jdet`, src/models.py:96-97) and the folding metric of evaluate.py:583-591.
//
// Reference arithmetic, kept op for op (including its quirky scaling: the field is normalised by
// 2/shape per channel, channel-flipped, and the flipped channel j is then scaled by (shape[j]-2)/2):
//   phi_j = ((df[2-j] * 2) / shape[2-j]) * (shape[j] - 2) / 2
//   J[a][j] = 0.5 * (phi_j(clamp(v + e_a)) - phi_j(clamp(v - e_a))) + delta_aj     (replication padding)
//   det = J00 (J11 J22 - J21 J12) - J01 (J10 J22 - J20 J12) + J02 (J10 J21 - J20 J11)
// HBM streaming: 12 B/voxel read (neighbours from L1/L2), 4 B/voxel written.  The backward stores the
// nine upstream-weighted cofactors (36 B/voxel) and applies the adjoint of the central differences in
// gather form (no atomics).
#include "common.cuh"

namespace pulpo {

struct JGeom {
    int B, D0, D1, D2, S;
    float shape[3];   // float(D0), float(D1), float(D2)
    float sm2[3];     // float(D0 - 2), ...
    int normalize;
};

static int make_jgeom(JGeom &g, int B, int D0, int D1, int D2, int normalize)
{
    const i64 S = (i64)D0 * D1 * D2;
    if (S >= (1ll << 31) || (i64)B * S >= (1ll << 31)) return PULPO_ERR_INVALID_SHAPE;
    g.B = B; g.D0 = D0; g.D1 = D1; g.D2 = D2; g.S = (int)S;
    g.shape[0] = (float)D0; g.shape[1] = (float)D1; g.shape[2] = (float)D2;
    g.sm2[0] = (float)(D0 - 2); g.sm2[1] = (float)(D1 - 2); g.sm2[2] = (float)(D2 - 2);
    g.normalize = normalize;
    return PULPO_OK;
}

// flipped channel j of the scaled field at one voxel (value of df channel 2-j passed in)
__device__ __forceinline__ float phi(float v, int j, const JGeom &g)
{
    if (g.normalize) v = __fdiv_rn(__fmul_rn(v, 2.0f), g.shape[2 - j]);
    return __fmul_rn(__fmul_rn(v, g.sm2[j]), 0.5f);
}

__device__ __forceinline__ void jacobian_at(const float *f, int z, int y, int x, const JGeom &g, float (&J)[3][3])
{
    const int sy = g.D2, sz = g.D1 * g.D2;
    const int off = (z * g.D1 + y) * g.D2 + x;
    const int dzm = z > 0 ? -sz : 0, dzp = z + 1 < g.D0 ? sz : 0;
    const int dym = y > 0 ? -sy : 0, dyp = y + 1 < g.D1 ? sy : 0;
    const int dxm = x > 0 ? -1 : 0, dxp = x + 1 < g.D2 ? 1 : 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float *c = f + (i64)(2 - j) * g.S + off;
        J[0][j] = __fmul_rn(0.5f, __fsub_rn(phi(__ldg(c + dzp), j, g), phi(__ldg(c + dzm), j, g)));
        J[1][j] = __fmul_rn(0.5f, __fsub_rn(phi(__ldg(c + dyp), j, g), phi(__ldg(c + dym), j, g)));
        J[2][j] = __fmul_rn(0.5f, __fsub_rn(phi(__ldg(c + dxp), j, g), phi(__ldg(c + dxm), j, g)));
        J[j][j] = __fadd_rn(J[j][j], 1.0f);
    }
}

__device__ __forceinline__ float det3(const float (&J)[3][3])
{
    const float m0 = __fsub_rn(__fmul_rn(J[1][1], J[2][2]), __fmul_rn(J[2][1], J[1][2]));
    const float m1 = __fsub_rn(__fmul_rn(J[1][0], J[2][2]), __fmul_rn(J[2][0], J[1][2]));
    const float m2 = __fsub_rn(__fmul_rn(J[1][0], J[2][1]), __fmul_rn(J[2][0], J[1][1]));
    return __fadd_rn(__fsub_rn(__fmul_rn(J[0][0], m0), __fmul_rn(J[0][1], m1)), __fmul_rn(J[0][2], m2));
}

__global__ void __launch_bounds__(256)
jacdet_fwd_kernel(const float *__restrict__ df, float *__restrict__ det, const JGeom g)
{
    const i64 total = (i64)g.B * g.S;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const int b = (int)(i / g.S), v = (int)(i - (i64)b * g.S);
        const int x = v % g.D2, r = v / g.D2, y = r % g.D1, z = r / g.D1;
        float J[3][3];
        jacobian_at(df + (i64)b * 3 * g.S, z, y, x, g, J);
        det[i] = det3(J);
    }
}

// upstream-weighted cofactors gC[a][j] = gdet * d det / d J[a][j], nine planes [9][B*S]
__global__ void __launch_bounds__(256)
jacdet_cof_kernel(const float *__restrict__ gdet, const float *__restrict__ df, float *__restrict__ gc, const JGeom g)
{
    const i64 total = (i64)g.B * g.S;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const int b = (int)(i / g.S), v = (int)(i - (i64)b * g.S);
        const int x = v % g.D2, r = v / g.D2, y = r % g.D1, z = r / g.D1;
        float J[3][3];
        jacobian_at(df + (i64)b * 3 * g.S, z, y, x, g, J);
        const float u = gdet[i];
        float C[3][3];
        C[0][0] = J[1][1] * J[2][2] - J[2][1] * J[1][2];
        C[0][1] = -(J[1][0] * J[2][2] - J[2][0] * J[1][2]);
        C[0][2] = J[1][0] * J[2][1] - J[2][0] * J[1][1];
        C[1][0] = -(J[0][1] * J[2][2] - J[0][2] * J[2][1]);
        C[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
        C[1][2] = -(J[0][0] * J[2][1] - J[0][1] * J[2][0]);
        C[2][0] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
        C[2][1] = -(J[0][0] * J[1][2] - J[0][2] * J[1][0]);
        C[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int j = 0; j < 3; ++j) gc[(i64)(a * 3 + j) * total + i] = u * C[a][j];
    }
}

// adjoint of the replication-padded central differences, gather form:
//   g_phi_j(w) = sum_a 0.5 * [ gC[a][j](w - e_a) - gC[a][j](w + e_a) ]   (neighbours inside the volume)
//                      + 0.5 * gC[a][j](w) at the upper face, - 0.5 * gC[a][j](w) at the lower face
__global__ void __launch_bounds__(256)
jacdet_adj_kernel(const float *__restrict__ gc, float *__restrict__ gdf, const JGeom g)
{
    const i64 total = (i64)g.B * g.S;
    const int sy = g.D2, sz = g.D1 * g.D2;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const int b = (int)(i / g.S), v = (int)(i - (i64)b * g.S);
        const int x = v % g.D2, r = v / g.D2, y = r % g.D1, z = r / g.D1;
        const int pos[3] = {z, y, x}, ext[3] = {g.D0, g.D1, g.D2}, stride[3] = {sz, sy, 1};
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float acc = 0.0f;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float *p = gc + (i64)(a * 3 + j) * total + i;
                const float lo = pos[a] > 0 ? __ldg(p - stride[a]) : -__ldg(p);
                const float hi = pos[a] + 1 < ext[a] ? __ldg(p + stride[a]) : -__ldg(p);
                acc += 0.5f * (lo - hi);
            }
            // d phi_j / d df[2-j]
            float k = g.sm2[j] * 0.5f;
            if (g.normalize) k *= 2.0f / g.shape[2 - j];
            gdf[((i64)b * 3 + (2 - j)) * g.S + v] = acc * k;
        }
    }
}

// ---- unbiased std over all elements (torch.Tensor.std()), deterministic two-stage reduction in double
struct StdWs {
    unsigned int ticket, pad;
    double mean, std;
    double partial[2];   // [2 * ctas]: sum, sum of squares
};
constexpr int kStdCtas = 592;

__global__ void __launch_bounds__(256)
std_fwd_kernel(const float *__restrict__ x, float lamb, float *out, StdWs *ws, i64 n)
{
    __shared__ double red[32];
    __shared__ bool is_last;
    double s = 0.0, q = 0.0;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double v = (double)x[i];
        s += v;
        q += v * v;
    }
    s = block_sum(s, red);
    q = block_sum(q, red);
    if (threadIdx.x == 0) {
        ws->partial[2 * blockIdx.x] = s;
        ws->partial[2 * blockIdx.x + 1] = q;
        __threadfence();
        is_last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double ts = 0.0, tq = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
            ts += ((volatile double *)ws->partial)[2 * i];
            tq += ((volatile double *)ws->partial)[2 * i + 1];
        }
        ts = block_sum(ts, red);
        tq = block_sum(tq, red);
        if (threadIdx.x == 0) {
            const double mean = ts / (double)n;
            double var = (tq - ts * mean) / (double)(n - 1);
            if (var < 0.0) var = 0.0;
            ws->mean = mean;
            ws->std = sqrt(var);
            *out = (float)((double)lamb * ws->std);
            ws->ticket = 0;
        }
    }
}

// d (lamb * std) / d x_i = lamb * (x_i - mean) / ((n - 1) * std)
__global__ void __launch_bounds__(256)
std_bwd_kernel(const float *__restrict__ gloss, const float *__restrict__ x, const StdWs *ws, float lamb,
               float *__restrict__ gx, i64 n)
{
    const double k = (double)((gloss ? __ldg(gloss) : 1.0f) * lamb) / ((double)(n - 1) * fmax(ws->std, 1e-300));
    const double mean = ws->mean;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        gx[i] = (float)(k * ((double)x[i] - mean));
}

}  // namespace pulpo

using namespace pulpo;

extern "C" int pulpo_jacdet_fwd(const float *df, float *det, int normalize, int B, int D0, int D1, int D2,
                                pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_jacdet_fwd");
    PULPO_REQUIRE(df && det, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && D0 > 0 && D1 > 0 && D2 > 0, PULPO_ERR_INVALID_SHAPE);
    JGeom g;
    int rc = make_jgeom(g, B, D0, D1, D2, normalize);
    if (rc != PULPO_OK) return rc;
    jacdet_fwd_kernel<<<grid_for((i64)B * g.S, 256), 256, 0, (cudaStream_t)stream>>>(df, det, g);
    return launch_status();
}

extern "C" size_t pulpo_jacdet_bwd_ws_bytes(int B, int D0, int D1, int D2)
{
    return (size_t)9 * B * D0 * D1 * D2 * sizeof(float);
}

extern "C" int pulpo_jacdet_bwd(const float *gdet, const float *df, float *gdf, void *ws, size_t ws_bytes, int normalize,
                                int B, int D0, int D1, int D2, pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_jacdet_bwd");
    PULPO_REQUIRE(gdet && df && gdf && ws, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(B > 0 && D0 > 0 && D1 > 0 && D2 > 0, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(ws_bytes >= pulpo_jacdet_bwd_ws_bytes(B, D0, D1, D2), PULPO_ERR_WORKSPACE);
    JGeom g;
    int rc = make_jgeom(g, B, D0, D1, D2, normalize);
    if (rc != PULPO_OK) return rc;
    const int grid = grid_for((i64)B * g.S, 256);
    jacdet_cof_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gdet, df, (float *)ws, g);
    jacdet_adj_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float *)ws, gdf, g);
    return launch_status();
}

extern "C" size_t pulpo_std_ws_bytes(void) { return 32 + sizeof(double) * 2 * kStdCtas; }

extern "C" int pulpo_std_fwd(const float *x, float lamb, float *out, void *ws, size_t ws_bytes, long long n,
                             pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_std_fwd");
    PULPO_REQUIRE(x && out && ws, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(n >= 2, PULPO_ERR_INVALID_SHAPE);
    PULPO_REQUIRE(ws_bytes >= pulpo_std_ws_bytes(), PULPO_ERR_WORKSPACE);
    int grid = grid_for(n, 256, 4);
    if (grid > kStdCtas) grid = kStdCtas;
    std_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, lamb, out, (StdWs *)ws, n);
    return launch_status();
}

extern "C" int pulpo_std_bwd(const float *gloss, const float *x, const void *ws, float lamb, float *gx, long long n,
                             pulpo_stream_t stream)
{
    PULPO_NVTX("pulpo_std_bwd");
    PULPO_REQUIRE(x && ws && gx, PULPO_ERR_NULL_POINTER);
    PULPO_REQUIRE(n >= 2, PULPO_ERR_INVALID_SHAPE);
    std_bwd_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(gloss, x, (const StdWs *)ws, lamb, gx, n);
    return launch_status();
}
