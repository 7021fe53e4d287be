/*
 * pulpo_oracle.c -- TEST INFRASTRUCTURE ONLY (parity oracle, never shipped, never measured
 * as the product).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.
 *
 * Plain-C, single-precision CPU restatement of PULPo's dense-3D registration hot path,
 * written from the arithmetic specification in SURVEY.md section 9 (not from the
 * reference sources).  Every op is rounded separately (build with -ffp-contract=off) so
 * the integer sampling indices are reproducible bit for bit.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * file is pinned against outputs of the reference itself, generated in the build
 * container by oracle/gen_golden.py (imports /root/reference) and committed under
 * tests/golden/.  tests/test_oracle_golden.py checks every function here against them.
 *
 * Reference call sites restated (path:line under the upstream repo):
 *   orc_warp3d_fwd/bwd     src/network_blocks.py:101-121  (+ ATen grid_sampler_3d, bilinear,
 *                          border, align_corners=False)
 *   orc_vecint_fwd/bwd     src/network_blocks.py:173-177
 *   orc_resize_up_fwd/bwd  src/network_blocks.py:138-150, DFAdder :152-158
 *                          (+ ATen upsample_trilinear3d, align_corners=False)
 *   orc_interp_size_fwd    src/losses.py:313, src/components/pulpo.py:202
 *   orc_avgpool2_fwd       src/components/pulpo.py:171-179, src/models.py:376-384
 *   orc_ncc_fwd/bwd        src/losses.py:85-135
 *   orc_kl_diag_fwd/bwd    src/losses.py:47-76
 *   orc_l2reg_fwd/bwd      src/losses.py:208-222
 *
 * Layout everywhere: contiguous fp32 [B, C, D0, D1, D2], D2 innermost.  Field channel a
 * displaces along spatial axis a (network_blocks.py:94-103).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_COORD_CPU_EXACT 0 /* x/(S-1) true division, unnormalise rounded op by op */
#define ORC_COORD_CUDA_RCP 1  /* x*(1/(S-1)) and fma((n+1),S,-1): what torch-CUDA does */

typedef long long i64;

/* ------------------------------------------------------------------ sample position */
/* network_blocks.py:103-107 then ATen unnormalise for align_corners=False.           */
static inline float orc_sample_pos(int v, float d, int S, int mode)
{
    float loc = (float)v + d;
    float q;
    if (mode == ORC_COORD_CUDA_RCP) {
        float r = 1.0f / (float)(S - 1);
        q = loc * r;
    } else {
        q = loc / (float)(S - 1);
    }
    float n = 2.0f * (q - 0.5f);
    float p;
    if (mode == ORC_COORD_CUDA_RCP)
        p = fmaf(n + 1.0f, (float)S, -1.0f) / 2.0f;
    else {
        float t = (n + 1.0f) * (float)S;
        p = (t - 1.0f) / 2.0f;
    }
    return p; /* unclamped */
}

/* border padding: clamp to [0, S-1] the way std::min(S-1, std::max(p, 0)) does */
static inline float orc_clip(float p, int S)
{
    float lo = (p < 0.0f) ? 0.0f : p;      /* std::max(p, 0): NaN stays NaN */
    float hi = (float)(S - 1);
    return (lo < hi) ? lo : hi;            /* std::min(hi, lo) */
}

/* ------------------------------------------------------------------ warp forward    */
void orc_warp3d_fwd(const float *img, const float *df, float *out, int32_t *idx,
                    int B, int C, int D0, int D1, int D2, int mode)
{
    const i64 S = (i64)D0 * D1 * D2;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int z = 0; z < D0; ++z)
            for (int y = 0; y < D1; ++y)
                for (int x = 0; x < D2; ++x) {
                    i64 v = ((i64)z * D1 + y) * D2 + x;
                    const float *f = df + (i64)b * 3 * S;
                    float pz = orc_clip(orc_sample_pos(z, f[v], D0, mode), D0);
                    float py = orc_clip(orc_sample_pos(y, f[S + v], D1, mode), D1);
                    float px = orc_clip(orc_sample_pos(x, f[2 * S + v], D2, mode), D2);
                    int iz = (int)floorf(pz), iy = (int)floorf(py), ix = (int)floorf(px);
                    if (idx) {
                        int32_t *o = idx + (i64)b * 3 * S;
                        o[v] = iz; o[S + v] = iy; o[2 * S + v] = ix;
                    }
                    /* corner weights: (i+1)-p on the low side, p-i on the high side */
                    float wz0 = (float)(iz + 1) - pz, wz1 = pz - (float)iz;
                    float wy0 = (float)(iy + 1) - py, wy1 = py - (float)iy;
                    float wx0 = (float)(ix + 1) - px, wx1 = px - (float)ix;
                    int zin1 = iz + 1 < D0, yin1 = iy + 1 < D1, xin1 = ix + 1 < D2;
                    float w000 = wx0 * wy0 * wz0, w001 = wx1 * wy0 * wz0;
                    float w010 = wx0 * wy1 * wz0, w011 = wx1 * wy1 * wz0;
                    float w100 = wx0 * wy0 * wz1, w101 = wx1 * wy0 * wz1;
                    float w110 = wx0 * wy1 * wz1, w111 = wx1 * wy1 * wz1;
                    i64 base = ((i64)iz * D1 + iy) * D2 + ix;
                    i64 sy = D2, sz = (i64)D1 * D2;
                    for (int c = 0; c < C; ++c) {
                        const float *im = img + ((i64)b * C + c) * S;
                        float acc = 0.0f;
                        acc += im[base] * w000;
                        if (xin1) acc += im[base + 1] * w001;
                        if (yin1) acc += im[base + sy] * w010;
                        if (yin1 && xin1) acc += im[base + sy + 1] * w011;
                        if (zin1) acc += im[base + sz] * w100;
                        if (zin1 && xin1) acc += im[base + sz + 1] * w101;
                        if (zin1 && yin1) acc += im[base + sz + sy] * w110;
                        if (zin1 && yin1 && xin1) acc += im[base + sz + sy + 1] * w111;
                        out[((i64)b * C + c) * S + v] = acc;
                    }
                }
}

/* ------------------------------------------------------------------ warp backward   */
/* SURVEY.md 9.2.  gimg (nullable) is ACCUMULATED into (caller zeroes); gdf (nullable)   */
/* is overwritten.  Serial over voxels inside one (b) so the scatter is deterministic. */
void orc_warp3d_bwd(const float *gout, const float *img, const float *df, float *gimg,
                    float *gdf, int B, int C, int D0, int D1, int D2, int mode)
{
    const i64 S = (i64)D0 * D1 * D2;
    const i64 sy = D2, sz = (i64)D1 * D2;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        for (int z = 0; z < D0; ++z)
            for (int y = 0; y < D1; ++y)
                for (int x = 0; x < D2; ++x) {
                    i64 v = ((i64)z * D1 + y) * D2 + x;
                    const float *f = df + (i64)b * 3 * S;
                    float uz = orc_sample_pos(z, f[v], D0, mode);
                    float uy = orc_sample_pos(y, f[S + v], D1, mode);
                    float ux = orc_sample_pos(x, f[2 * S + v], D2, mode);
                    /* d p / d df: (S/2) from the unnormalise, zero where the clamp is active,
                       then autograd of 2*(loc/(S-1)-0.5): (g*2)/(S-1) */
                    float mz = (uz <= 0.0f || uz >= (float)(D0 - 1)) ? 0.0f : (float)D0 / 2.0f;
                    float my = (uy <= 0.0f || uy >= (float)(D1 - 1)) ? 0.0f : (float)D1 / 2.0f;
                    float mx = (ux <= 0.0f || ux >= (float)(D2 - 1)) ? 0.0f : (float)D2 / 2.0f;
                    float pz = orc_clip(uz, D0), py = orc_clip(uy, D1), px = orc_clip(ux, D2);
                    int iz = (int)floorf(pz), iy = (int)floorf(py), ix = (int)floorf(px);
                    float wz0 = (float)(iz + 1) - pz, wz1 = pz - (float)iz;
                    float wy0 = (float)(iy + 1) - py, wy1 = py - (float)iy;
                    float wx0 = (float)(ix + 1) - px, wx1 = px - (float)ix;
                    int zin1 = iz + 1 < D0, yin1 = iy + 1 < D1, xin1 = ix + 1 < D2;
                    i64 base = ((i64)iz * D1 + iy) * D2 + ix;
                    float gx = 0.f, gy = 0.f, gz = 0.f;
                    for (int c = 0; c < C; ++c) {
                        i64 off = ((i64)b * C + c) * S;
                        const float *im = img + off;
                        float g = gout[off + v];
                        for (int k = 0; k < 8; ++k) {
                            int dz = k >> 2, dy = (k >> 1) & 1, dx = k & 1;
                            if ((dz && !zin1) || (dy && !yin1) || (dx && !xin1)) continue;
                            float wz = dz ? wz1 : wz0, wy = dy ? wy1 : wy0, wx = dx ? wx1 : wx0;
                            i64 a = base + dz * sz + dy * sy + dx;
                            if (gimg) gimg[off + a] += (wx * wy * wz) * g;
                            float val = im[a];
                            float tx = val * wy * wz * g, ty = val * wx * wz * g, tz = val * wx * wy * g;
                            gx = dx ? gx + tx : gx - tx;
                            gy = dy ? gy + ty : gy - ty;
                            gz = dz ? gz + tz : gz - tz;
                        }
                    }
                    if (gdf) {
                        float *o = gdf + (i64)b * 3 * S;
                        o[v] = ((mz * gz) * 2.0f) / (float)(D0 - 1);
                        o[S + v] = ((my * gy) * 2.0f) / (float)(D1 - 1);
                        o[2 * S + v] = ((mx * gx) * 2.0f) / (float)(D2 - 1);
                    }
                }
}

/* ------------------------------------------------------------------ VecInt           */
/* network_blocks.py:173-177.  steps = [nsteps+1][B,3,S]; steps[0]=vec*2^-nsteps,       */
/* steps[k+1] = steps[k] + warp(steps[k], steps[k]); result = steps[nsteps].           */
void orc_vecint_fwd(const float *vec, float *steps, int nsteps, int B, int D0, int D1, int D2,
                    int mode)
{
    const i64 n = (i64)B * 3 * D0 * D1 * D2;
    const float scale = 1.0f / (float)(1 << nsteps);
    for (i64 i = 0; i < n; ++i) steps[i] = vec[i] * scale;
    for (int k = 0; k < nsteps; ++k) {
        float *cur = steps + (i64)k * n, *nxt = steps + (i64)(k + 1) * n;
        orc_warp3d_fwd(cur, cur, nxt, NULL, B, 3, D0, D1, D2, mode);
        for (i64 i = 0; i < n; ++i) nxt[i] = cur[i] + nxt[i];
    }
}

/* gvec = d loss / d vec given gout = d loss / d steps[nsteps] (SURVEY.md 9.2 last para) */
void orc_vecint_bwd(const float *gout, const float *steps, float *gvec, int nsteps, int B,
                    int D0, int D1, int D2, int mode)
{
    const i64 n = (i64)B * 3 * D0 * D1 * D2;
    const float scale = 1.0f / (float)(1 << nsteps);
    float *g = (float *)malloc(sizeof(float) * n);
    float *gi = (float *)malloc(sizeof(float) * n);
    float *gd = (float *)malloc(sizeof(float) * n);
    memcpy(g, gout, sizeof(float) * n);
    for (int k = nsteps - 1; k >= 0; --k) {
        const float *cur = steps + (i64)k * n;
        memset(gi, 0, sizeof(float) * n);
        orc_warp3d_bwd(g, cur, cur, gi, gd, B, 3, D0, D1, D2, mode);
        for (i64 i = 0; i < n; ++i) g[i] = g[i] + gi[i] + gd[i];
    }
    for (i64 i = 0; i < n; ++i) gvec[i] = g[i] * scale;
    free(g); free(gi); free(gd);
}

/* ------------------------------------------------------------------ trilinear resize */
/* ATen upsample (align_corners=False): src = max(0, scale*(o+.5)-.5); i0=min(floor,n-1); */
/* l1 = clamp(src-i0,0,1); i1 = i0 + (i0<n-1); l0 = 1-l1.  Equal sizes -> plain copy.     */
static inline void orc_lin_tap(int o, int n_in, int n_out, float scale, int *i0, int *i1,
                               float *l0, float *l1)
{
    if (n_in == n_out) { *i0 = o; *i1 = o; *l0 = 1.0f; *l1 = 0.0f; return; }
    float src = scale * ((float)o + 0.5f) - 0.5f;
    if (src < 0.0f) src = 0.0f;
    int i = (int)floorf(src);
    if (i > n_in - 1) i = n_in - 1;
    float lam = src - (float)i;
    if (lam < 0.0f) lam = 0.0f;
    if (lam > 1.0f) lam = 1.0f;
    *i0 = i; *i1 = i + ((i < n_in - 1) ? 1 : 0); *l1 = lam; *l0 = 1.0f - lam;
}

static void orc_trilinear(const float *x, const float *addend, float *out, float premul,
                          int BC, int i0n, int i1n, int i2n, int o0n, int o1n, int o2n,
                          float s0, float s1, float s2)
{
    const i64 Si = (i64)i0n * i1n * i2n, So = (i64)o0n * o1n * o2n;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bc = 0; bc < BC; ++bc)
        for (int z = 0; z < o0n; ++z) {
            int za, zb; float lz0, lz1;
            orc_lin_tap(z, i0n, o0n, s0, &za, &zb, &lz0, &lz1);
            for (int y = 0; y < o1n; ++y) {
                int ya, yb; float ly0, ly1;
                orc_lin_tap(y, i1n, o1n, s1, &ya, &yb, &ly0, &ly1);
                for (int xx = 0; xx < o2n; ++xx) {
                    int xa, xb; float lx0, lx1;
                    orc_lin_tap(xx, i2n, o2n, s2, &xa, &xb, &lx0, &lx1);
                    const float *p = x + (i64)bc * Si;
#define AT(zz, yy, xq) (premul * p[((i64)(zz) * i1n + (yy)) * i2n + (xq)])
                    float r00 = AT(za, ya, xa) * lx0 + AT(za, ya, xb) * lx1;
                    float r01 = AT(za, yb, xa) * lx0 + AT(za, yb, xb) * lx1;
                    float r10 = AT(zb, ya, xa) * lx0 + AT(zb, ya, xb) * lx1;
                    float r11 = AT(zb, yb, xa) * lx0 + AT(zb, yb, xb) * lx1;
#undef AT
                    float r0 = r00 * ly0 + r01 * ly1;
                    float r1 = r10 * ly0 + r11 * ly1;
                    float r = r0 * lz0 + r1 * lz1;
                    i64 o = (i64)bc * So + ((i64)z * o1n + y) * o2n + xx;
                    out[o] = addend ? r + addend[o] : r;
                }
            }
        }
}

/* ResizeTransform(factor>1) + optional DFAdder: out = interp_f(scale * x) (+ addend).  */
/* x: [B,C,d0,d1,d2] -> out: [B,C,f*d0,f*d1,f*d2]; the op receives scale_factor=f so    */
/* the source scale is exactly 1/f.                                                    */
void orc_resize_up_fwd(const float *x, const float *addend, float *out, int factor, float scale,
                       int B, int C, int d0, int d1, int d2)
{
    float s = (float)(1.0 / (double)factor);
    orc_trilinear(x, addend, out, scale, B * C, d0, d1, d2, factor * d0, factor * d1,
                  factor * d2, s, s, s);
}

/* exact adjoint of the above w.r.t. x (scatter form, serial per (b,c)) */
void orc_resize_up_bwd(const float *gout, float *gx, int factor, float scale, int B, int C,
                       int d0, int d1, int d2)
{
    const int o0n = factor * d0, o1n = factor * d1, o2n = factor * d2;
    const i64 Si = (i64)d0 * d1 * d2, So = (i64)o0n * o1n * o2n;
    float s = (float)(1.0 / (double)factor);
    memset(gx, 0, sizeof(float) * (size_t)(B * C) * Si);
#pragma omp parallel for schedule(static)
    for (int bc = 0; bc < B * C; ++bc) {
        float *g = gx + (i64)bc * Si;
        for (int z = 0; z < o0n; ++z) {
            int za, zb; float lz0, lz1;
            orc_lin_tap(z, d0, o0n, s, &za, &zb, &lz0, &lz1);
            for (int y = 0; y < o1n; ++y) {
                int ya, yb; float ly0, ly1;
                orc_lin_tap(y, d1, o1n, s, &ya, &yb, &ly0, &ly1);
                for (int xx = 0; xx < o2n; ++xx) {
                    int xa, xb; float lx0, lx1;
                    orc_lin_tap(xx, d2, o2n, s, &xa, &xb, &lx0, &lx1);
                    float go = gout[(i64)bc * So + ((i64)z * o1n + y) * o2n + xx];
#define G(zz, yy, xq) g[((i64)(zz) * d1 + (yy)) * d2 + (xq)]
                    G(za, ya, xa) += lz0 * ly0 * lx0 * go; G(za, ya, xb) += lz0 * ly0 * lx1 * go;
                    G(za, yb, xa) += lz0 * ly1 * lx0 * go; G(za, yb, xb) += lz0 * ly1 * lx1 * go;
                    G(zb, ya, xa) += lz1 * ly0 * lx0 * go; G(zb, ya, xb) += lz1 * ly0 * lx1 * go;
                    G(zb, yb, xa) += lz1 * ly1 * lx0 * go; G(zb, yb, xb) += lz1 * ly1 * lx1 * go;
#undef G
                }
            }
        }
        for (i64 i = 0; i < Si; ++i) g[i] *= scale;
    }
}

/* F.interpolate(x, size=(o0,o1,o2), trilinear, align_corners=False): scale = in/out   */
void orc_interp_size_fwd(const float *x, float *out, int B, int C, int i0n, int i1n, int i2n,
                         int o0n, int o1n, int o2n)
{
    orc_trilinear(x, NULL, out, 1.0f, B * C, i0n, i1n, i2n, o0n, o1n, o2n,
                  (float)i0n / (float)o0n, (float)i1n / (float)o1n, (float)i2n / (float)o2n);
}

/* avg_pool3d(kernel 2, stride 2, pad 0, ceil_mode=True): border windows are clipped   */
/* to the input and divided by the clipped element count.                              */
void orc_avgpool2_fwd(const float *x, float *out, int B, int C, int D0, int D1, int D2)
{
    const int o0 = (D0 + 1) / 2, o1 = (D1 + 1) / 2, o2 = (D2 + 1) / 2;
    const i64 Si = (i64)D0 * D1 * D2, So = (i64)o0 * o1 * o2;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bc = 0; bc < B * C; ++bc)
        for (int z = 0; z < o0; ++z)
            for (int y = 0; y < o1; ++y)
                for (int xx = 0; xx < o2; ++xx) {
                    int z1 = 2 * z + 2 < D0 ? 2 * z + 2 : D0;
                    int y1 = 2 * y + 2 < D1 ? 2 * y + 2 : D1;
                    int x1 = 2 * xx + 2 < D2 ? 2 * xx + 2 : D2;
                    float s = 0.0f;
                    for (int a = 2 * z; a < z1; ++a)
                        for (int b2 = 2 * y; b2 < y1; ++b2)
                            for (int c = 2 * xx; c < x1; ++c)
                                s += x[(i64)bc * Si + ((i64)a * D1 + b2) * D2 + c];
                    int cnt = (z1 - 2 * z) * (y1 - 2 * y) * (x1 - 2 * xx);
                    out[(i64)bc * So + ((i64)z * o1 + y) * o2 + xx] = s / (float)cnt;
                }
}

/* ------------------------------------------------------------------ local NCC         */
/* zero-padded win^3 box sum of src -> dst (one [D0,D1,D2] volume), separable, fp32     */
static void orc_box3(const float *src, float *dst, float *tmp, int D0, int D1, int D2, int win)
{
    const int r = win / 2;
    const i64 sy = D2, sz = (i64)D1 * D2;
    /* along x: src -> dst */
    for (i64 zy = 0; zy < (i64)D0 * D1; ++zy)
        for (int x = 0; x < D2; ++x) {
            float s = 0.0f;
            int a = x - r < 0 ? 0 : x - r, b = x + r >= D2 ? D2 - 1 : x + r;
            for (int k = a; k <= b; ++k) s += src[zy * D2 + k];
            dst[zy * D2 + x] = s;
        }
    /* along y: dst -> tmp */
    for (int z = 0; z < D0; ++z)
        for (int y = 0; y < D1; ++y) {
            int a = y - r < 0 ? 0 : y - r, b = y + r >= D1 ? D1 - 1 : y + r;
            for (int x = 0; x < D2; ++x) {
                float s = 0.0f;
                for (int k = a; k <= b; ++k) s += dst[z * sz + k * sy + x];
                tmp[z * sz + y * sy + x] = s;
            }
        }
    /* along z: tmp -> dst */
    for (int z = 0; z < D0; ++z) {
        int a = z - r < 0 ? 0 : z - r, b = z + r >= D0 ? D0 - 1 : z + r;
        for (i64 yx = 0; yx < sz; ++yx) {
            float s = 0.0f;
            for (int k = a; k <= b; ++k) s += tmp[k * sz + yx];
            dst[z * sz + yx] = s;
        }
    }
}

/* losses.py:85-135.  pred = J (y_pred), target = I (y_true).  Returns the loss; if     */
/* gpred != NULL also writes d loss / d pred (closed form).              */
double orc_ncc(const float *pred, const float *target, float *gpred, int win, float gamma,
               int B, int C, int D0, int D1, int D2)
{
    const i64 S = (i64)D0 * D1 * D2;
    const float W = (float)((i64)win * win * win);
    double total = 0.0;
#pragma omp parallel for schedule(dynamic) reduction(+ : total)
    for (int bc = 0; bc < B * C; ++bc) {
        const float *I = target + (i64)bc * S, *J = pred + (i64)bc * S;
        float *buf = (float *)malloc(sizeof(float) * S * 10);
        float *prod = buf, *tmp = buf + S, *sI = buf + 2 * S, *sJ = buf + 3 * S, *sII = buf + 4 * S,
              *sJJ = buf + 5 * S, *sIJ = buf + 6 * S, *fa = buf + 7 * S, *fb = buf + 8 * S,
              *fc = buf + 9 * S;
        orc_box3(I, sI, tmp, D0, D1, D2, win);
        orc_box3(J, sJ, tmp, D0, D1, D2, win);
        for (i64 i = 0; i < S; ++i) prod[i] = I[i] * I[i];
        orc_box3(prod, sII, tmp, D0, D1, D2, win);
        for (i64 i = 0; i < S; ++i) prod[i] = J[i] * J[i];
        orc_box3(prod, sJJ, tmp, D0, D1, D2, win);
        for (i64 i = 0; i < S; ++i) prod[i] = I[i] * J[i];
        orc_box3(prod, sIJ, tmp, D0, D1, D2, win);
        double acc = 0.0;
        for (i64 i = 0; i < S; ++i) {
            float uI = sI[i] / W, uJ = sJ[i] / W;
            float cross = sIJ[i] - uJ * sI[i] - uI * sJ[i] + uI * uJ * W;
            float Iv = sII[i] - 2.0f * uI * sI[i] + uI * uI * W;
            float Jv = sJJ[i] - 2.0f * uJ * sJ[i] + uJ * uJ * W;
            float D = Iv * Jv + 1e-8f;
            float cc = cross * cross / D;
            acc += (double)cc;
            if (gpred) {
                float a = 2.0f * cross / D;
                float c = -(cross * cross) * Iv / (D * D);
                fa[i] = a; fc[i] = c;
                fb[i] = -a * sI[i] / W - 2.0f * c * sJ[i] / W;
            }
        }
        total += acc;
        if (gpred) {
            float *Ba = sI, *Bb = sJ, *Bc = sII;
            orc_box3(fa, Ba, tmp, D0, D1, D2, win);
            orc_box3(fb, Bb, tmp, D0, D1, D2, win);
            orc_box3(fc, Bc, tmp, D0, D1, D2, win);
            float k = -gamma / (float)B;
            float *g = gpred + (i64)bc * S;
            for (i64 i = 0; i < S; ++i) g[i] = k * (I[i] * Ba[i] + Bb[i] + 2.0f * J[i] * Bc[i]);
        }
        free(buf);
    }
    return -(double)gamma * total / (double)B;
}

/* ------------------------------------------------------------------ diagonal KL      */
/* losses.py:47-76: KL[p0||p1], mean over batch of 0.5*sum(...).  mu1/sigma1 may be     */
/* NULL meaning the N(0,1) prior of pulpo.py:337-339.                                   */
double orc_kl_diag_fwd(const float *mu0, const float *sg0, const float *mu1, const float *sg1,
                       float eps, int B, i64 n)
{
    double total = 0.0;
    for (i64 i = 0; i < (i64)B * n; ++i) {
        float s0 = sg0[i] * sg0[i];
        float s1 = sg1 ? sg1[i] * sg1[i] : 1.0f;
        float m1 = mu1 ? mu1[i] : 0.0f;
        float dm = m1 - mu0[i];
        float t = (s0 + dm * dm) / (s1 + eps) + logf(s1 + eps) - logf(s0 + eps) - 1.0f;
        total += (double)t;
    }
    return 0.5 * total / (double)B;
}

void orc_kl_diag_bwd(const float *mu0, const float *sg0, const float *mu1, const float *sg1,
                     float eps, float gscale, float *gmu0, float *gsg0, int B, i64 n)
{
    float k = gscale / (float)B;
    for (i64 i = 0; i < (i64)B * n; ++i) {
        float s1 = sg1 ? sg1[i] * sg1[i] : 1.0f;
        float m1 = mu1 ? mu1[i] : 0.0f;
        float den = s1 + eps;
        float s = sg0[i];
        gmu0[i] = k * (mu0[i] - m1) / den;
        gsg0[i] = k * (s / den - s / (s * s + eps));
    }
}

/* ------------------------------------------------------------------ L2 regulariser   */
/* losses.py:208-222 (3-D branch): forward differences on the [1:,1:,1:] crop,          */
/* mean * lamb * D0*D1*D2.                                                              */
double orc_l2reg_fwd(const float *f, float lamb, int B, int C, int D0, int D1, int D2)
{
    const i64 S = (i64)D0 * D1 * D2, sy = D2, sz = (i64)D1 * D2;
    double total = 0.0;
    for (int bc = 0; bc < B * C; ++bc)
        for (int z = 1; z < D0; ++z)
            for (int y = 1; y < D1; ++y)
                for (int x = 1; x < D2; ++x) {
                    const float *p = f + (i64)bc * S + z * sz + y * sy + x;
                    float a = p[0] - p[-sz], b = p[0] - p[-sy], c = p[0] - p[-1];
                    total += (double)(a * a + b * b + c * c);
                }
    double cnt = (double)B * C * (D0 - 1) * (D1 - 1) * (D2 - 1);
    return total / cnt * (double)lamb * D0 * D1 * D2;
}

void orc_l2reg_bwd(const float *f, float lamb, float gscale, float *gf, int B, int C, int D0,
                   int D1, int D2)
{
    const i64 S = (i64)D0 * D1 * D2, sy = D2, sz = (i64)D1 * D2;
    double cnt = (double)B * C * (D0 - 1) * (D1 - 1) * (D2 - 1);
    float k = (float)(2.0 * (double)gscale * (double)lamb * D0 * D1 * D2 / cnt);
    memset(gf, 0, sizeof(float) * (size_t)(B * C) * S);
    for (int bc = 0; bc < B * C; ++bc)
        for (int z = 1; z < D0; ++z)
            for (int y = 1; y < D1; ++y)
                for (int x = 1; x < D2; ++x) {
                    i64 o = (i64)bc * S + z * sz + y * sy + x;
                    float a = f[o] - f[o - sz], b = f[o] - f[o - sy], c = f[o] - f[o - 1];
                    gf[o] += k * (a + b + c);
                    gf[o - sz] -= k * a; gf[o - sy] -= k * b; gf[o - 1] -= k * c;
                }
}
