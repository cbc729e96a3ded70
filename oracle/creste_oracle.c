/*
 * creste_oracle.c -- CPU restatement of the reference's HBM/latency-bound hot-path functions.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the oracle (checker) for the CUDA kernels in
 * creste_public_b200/csrc/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product path never calls into it.
 *
 * Parity status: PINNED -- every function here is checked bit-for-bit / to tolerance against
 * the reference's own Python code executed in the build container (tests/test_oracle_cpu.py,
 * fixtures under tests/golden/ made by oracle/gen_golden.py).  The reference itself ships no
 * tests or golden vectors.
 *
 * fp32 arithmetic is written with explicit fmaf() in the operation order torch's CPU kernels
 * were measured to use (see DESIGN.md "Arithmetic order"), and the file is compiled with
 * -ffp-contract=off so the compiler neither fuses nor splits anything on its own.
 *
 * Functions (reference file:line each one follows):
 *   oracle_vi_solve          creste/models/blocks/vin.py:36-80      (stencil :36-46, loop :48-80)
 *   oracle_svf               creste/models/lfd.py:156-277           (+ train_utils.py:765-803)
 *   oracle_frustum_to_bev    creste/models/blocks/splat_projection.py:19-51, :169, :175-189
 *   oracle_splat_soft        creste/models/blocks/splat_projection.py:262-354
 *   oracle_lidar_raster      creste/utils/projection.py:64-134 + scripts/preprocessing/
 *                            build_dense_depth.py:461-463, creste/utils/depth_utils.py:14-39
 *   oracle_depth_expectation creste/utils/depth_utils.py:300-313, creste/models/depth.py:60-100
 *   oracle_expert_visitation creste/utils/loss_utils.py:1055-1116 (second definition)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- VI stencil (vin.py:36-46): per action, the three non-zero taps in (ky,kx) raster order.
 * torch CPU conv2d == in-order FMA chain over the 3x3 taps starting from acc = 0; zero-weight
 * taps leave acc unchanged, so only the non-zero taps are listed (measured: 0 mismatches). */
typedef struct { int dy, dx; float w; } tap_t;
static const tap_t VI_TAPS[8][3] = {
    /* a0 dest(-1,-1): c(0,0) l(1,0) r(0,1) */ {{-1,-1,0.8f},{-1, 0,0.1f},{ 0,-1,0.1f}},
    /* a1 dest(-1, 0): c(0,1) l(0,0) r(0,2) */ {{-1,-1,0.1f},{-1, 0,0.8f},{-1, 1,0.1f}},
    /* a2 dest(-1, 1): c(0,2) l(0,1) r(1,2) */ {{-1, 0,0.1f},{-1, 1,0.8f},{ 0, 1,0.1f}},
    /* a3 dest( 0,-1): c(1,0) l(2,0) r(0,0) */ {{-1,-1,0.1f},{ 0,-1,0.8f},{ 1,-1,0.1f}},
    /* a4 dest( 0, 1): c(1,2) l(0,2) r(2,2) */ {{-1, 1,0.1f},{ 0, 1,0.8f},{ 1, 1,0.1f}},
    /* a5 dest( 1,-1): c(2,0) l(2,1) r(1,0) */ {{ 0,-1,0.1f},{ 1,-1,0.8f},{ 1, 0,0.1f}},
    /* a6 dest( 1, 0): c(2,1) l(2,2) r(2,0) */ {{ 1,-1,0.1f},{ 1, 0,0.8f},{ 1, 1,0.1f}},
    /* a7 dest( 1, 1): c(2,2) l(1,2) r(2,1) */ {{ 0, 1,0.1f},{ 1, 0,0.1f},{ 1, 1,0.8f}},
};

static inline float xval(const float* X, int H, int W, int y, int x) {
    return (y < 0 || y >= H || x < 0 || x >= W) ? 0.0f : X[(size_t)y * W + x];
}

static void vi_q(const float* X, int H, int W, int y, int x, float q[8]) {
    for (int a = 0; a < 8; ++a) {
        float acc = 0.0f;
        for (int t = 0; t < 3; ++t)
            acc = fmaf(VI_TAPS[a][t].w, xval(X, H, W, y + VI_TAPS[a][t].dy, x + VI_TAPS[a][t].dx), acc);
        q[a] = acc;
    }
}

/* r,v: [B,H,W]; q,pi: [B,8,H,W] (may be NULL).  Returns 0; *sweeps = number of Bellman sweeps. */
int oracle_vi_solve(const float* r, float* v, float* q, float* pi, int B, int H, int W,
                    float gamma, float thr, int max_sweeps, int* sweeps) {
    const size_t HW = (size_t)H * W, N = (size_t)B * HW;
    float* X = (float*)malloc(N * sizeof(float));
    float* vn = (float*)malloc(N * sizeof(float));
    if (!X || !vn) return 1;
    memset(v, 0, N * sizeof(float));
    int K = 0;
    float delta = INFINITY;
    while (delta > thr && K < max_sweeps) {
        float dmax = 0.0f;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)N; ++i) X[i] = r[i] + v[i] * gamma; /* two roundings (no contraction) */
#pragma omp parallel for reduction(max : dmax) schedule(static)
        for (long by = 0; by < (long)B * H; ++by) {
            const int b = (int)(by / H), y = (int)(by % H);
            const float* Xb = X + (size_t)b * HW;
            for (int x = 0; x < W; ++x) {
                float qq[8];
                vi_q(Xb, H, W, y, x, qq);
                float m = qq[0];
                for (int a = 1; a < 8; ++a) m = qq[a] > m ? qq[a] : m;
                const size_t i = (size_t)b * HW + (size_t)y * W + x;
                vn[i] = m;
                const float d = fabsf(m - v[i]);
                dmax = d > dmax ? d : dmax;
            }
        }
        memcpy(v, vn, N * sizeof(float));
        delta = dmax;
        ++K;
    }
    if (sweeps) *sweeps = K;
    if (q || pi) {
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)N; ++i) X[i] = r[i] + v[i] * gamma;
#pragma omp parallel for schedule(static)
        for (long by = 0; by < (long)B * H; ++by) {
            const int b = (int)(by / H), y = (int)(by % H);
            const float* Xb = X + (size_t)b * HW;
            for (int x = 0; x < W; ++x) {
                float qq[8], e[8];
                vi_q(Xb, H, W, y, x, qq);
                float m = qq[0];
                for (int a = 1; a < 8; ++a) m = qq[a] > m ? qq[a] : m;
                float s = 0.0f;
                for (int a = 0; a < 8; ++a) { e[a] = expf(qq[a] - m); s += e[a]; }
                for (int a = 0; a < 8; ++a) {
                    const size_t o = ((size_t)b * 8 + a) * HW + (size_t)y * W + x;
                    if (q) q[o] = qq[a];
                    if (pi) pi[o] = e[a] / s;
                }
            }
        }
    }
    free(X);
    free(vn);
    return 0;
}

/* ---- SVF (lfd.py:156-277).  dynamics rows (lfd.py:37-46). */
static const int DYN[8][2] = {{-1,-1},{-1,0},{-1,1},{0,-1},{0,1},{1,-1},{1,0},{1,1}};

/* policy [B,8,H,W]; expert_rc [B,T,2] (row, col) in un-pooled BEV cells (float: expert[:,:,:2,2]);
 * fov [H,W] uint8; outputs exp_svf [B,H,W], states [B,T,2] int64, states_grid [B,H,W].
 * sharpen != 0 => policy <- softmax((pi - max pi)/temperature) (lfd.py:190-194). */
int oracle_svf(const float* policy, const float* expert_rc, const uint8_t* fov, int B, int H, int W,
               int T, int ds, int sharpen, float temperature, int zero_terminal, float* exp_svf,
               int64_t* states, float* states_grid) {
    const size_t HW = (size_t)H * W;
    float* pol = (float*)malloc(8 * HW * sizeof(float));
    float* mu = (float*)malloc(HW * sizeof(float));
    float* mun = (float*)malloc(HW * sizeof(float));
    if (!pol || !mu || !mun) return 1;
    for (int b = 0; b < B; ++b) {
        const float* P = policy + (size_t)b * 8 * HW;
        /* S = (expert // ds).long() clamped (lfd.py:171-173); S0 = earliest pose in FOV
         * (train_utils.py:765-803), default (H-1, W//2); S1 = last pose. */
        long s0r = H - 1, s0c = W / 2, s1r = 0, s1c = 0;
        int found = 0;
        for (int t = 0; t < T; ++t) {
            long rr = (long)floorf(expert_rc[((size_t)b * T + t) * 2 + 0] / (float)ds);
            long cc = (long)floorf(expert_rc[((size_t)b * T + t) * 2 + 1] / (float)ds);
            rr = rr < 0 ? 0 : (rr > H - 1 ? H - 1 : rr);
            cc = cc < 0 ? 0 : (cc > W - 1 ? W - 1 : cc);
            if (!found && fov[rr * W + cc]) { s0r = rr; s0c = cc; found = 1; }
            if (t == T - 1) { s1r = rr; s1c = cc; }
        }
        const size_t S0 = (size_t)s0r * W + s0c, S1 = (size_t)s1r * W + s1c;
        for (size_t i = 0; i < HW; ++i) {
            if (sharpen) {
                float m = P[i];
                for (int a = 1; a < 8; ++a) m = P[a * HW + i] > m ? P[a * HW + i] : m;
                float l[8], lm, e[8], s = 0.0f;
                for (int a = 0; a < 8; ++a) l[a] = (P[a * HW + i] - m) / temperature;
                lm = l[0];
                for (int a = 1; a < 8; ++a) lm = l[a] > lm ? l[a] : lm;
                for (int a = 0; a < 8; ++a) { e[a] = expf(l[a] - lm); s += e[a]; }
                for (int a = 0; a < 8; ++a) pol[a * HW + i] = e[a] / s;
            } else {
                for (int a = 0; a < 8; ++a) pol[a * HW + i] = P[a * HW + i];
            }
        }
        float* out = exp_svf + (size_t)b * HW;
        memset(mu, 0, HW * sizeof(float));
        mu[S0] = 1.0f;
        /* The reference zeroes mu[t-1][S1] in place *before* propagating (lfd.py:202-203), so the
         * zeroed value is also what ends up in the time-sum; mu[T-1] is never zeroed. */
        memset(out, 0, HW * sizeof(float));
        for (int t = 1; t < T; ++t) {
            if (zero_terminal) mu[S1] = 0.0f;
            for (size_t i = 0; i < HW; ++i) out[i] += mu[i];
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    float acc = 0.0f;
                    for (int a = 0; a < 8; ++a) {
                        const int sy = y - DYN[a][0], sx = x - DYN[a][1];
                        float c = 0.0f;
                        if (sy >= 0 && sy < H && sx >= 0 && sx < W)
                            c = pol[a * HW + (size_t)sy * W + sx] * mu[(size_t)sy * W + sx];
                        acc += c;
                    }
                    mun[(size_t)y * W + x] = acc;
                }
            float* tmp = mu; mu = mun; mun = tmp;
        }
        for (size_t i = 0; i < HW; ++i) out[i] += mu[i];
        /* greedy rollout (lfd.py:230-248) on the ORIGINAL (un-sharpened) policy */
        if (states && states_grid) {
            float* g = states_grid + (size_t)b * HW;
            memset(g, 0, HW * sizeof(float));
            long cr = s0r, cc = s0c;
            states[((size_t)b * T) * 2 + 0] = cr;
            states[((size_t)b * T) * 2 + 1] = cc;
            g[cr * W + cc] += 1.0f;
            for (int t = 1; t < T; ++t) {
                const size_t st = (size_t)cr * W + cc;
                int best = 0;
                float bv = P[st];
                for (int a = 1; a < 8; ++a)
                    if (P[a * HW + st] > bv) { bv = P[a * HW + st]; best = a; }
                cr += DYN[best][0]; cc += DYN[best][1];
                cr = cr < 0 ? 0 : (cr > H - 1 ? H - 1 : cr);
                cc = cc < 0 ? 0 : (cc > W - 1 ? W - 1 : cc);
                states[((size_t)b * T + t) * 2 + 0] = cr;
                states[((size_t)b * T + t) * 2 + 1] = cc;
                g[cr * W + cc] += 1.0f;
            }
        }
    }
    free(pol); free(mu); free(mun);
    return 0;
}

/* ---- frustum -> LiDAR xyz -> BEV voxel coords (splat_projection.py:19-51, :169, :175-189).
 * depth [N,Hs,Ws] (m), p2p [N,4,4] row-major, range = [xmin,ymin,zmin,xmax,ymax,zmax],
 * voxel = [vx,vy].  Outputs xyz [N,3,Hs*Ws] (channel-major, as the reference's view),
 * xy [N,Hs*Ws,2] float voxel coords, mask [N,Hs*Ws] uint8.
 * bmm K=4 == acc=a0*b0; acc=fma(a1,b1,acc); ... (measured on torch CPU). */
int oracle_frustum_to_bev(const float* depth, const float* p2p, const float* range,
                          const float* voxel, int N, int Hs, int Ws, float* xyz, float* xy,
                          uint8_t* mask) {
    const size_t P = (size_t)Hs * Ws;
    for (int n = 0; n < N; ++n) {
        const float* M = p2p + (size_t)n * 16;
#pragma omp parallel for schedule(static)
        for (long p = 0; p < (long)P; ++p) {
            const int vv = (int)(p / Ws), uu = (int)(p % Ws);
            const float d = depth[(size_t)n * P + p];
            const float c[4] = {(float)uu * d, (float)vv * d, 1.0f * d, 1.0f};
            float o[3];
            for (int i = 0; i < 3; ++i) {
                float acc = M[i * 4 + 0] * c[0];
                acc = fmaf(M[i * 4 + 1], c[1], acc);
                acc = fmaf(M[i * 4 + 2], c[2], acc);
                acc = fmaf(M[i * 4 + 3], c[3], acc);
                o[i] = acc;
                if (xyz) xyz[((size_t)n * 3 + i) * P + p] = acc;
            }
            int ok = 1;
            for (int i = 0; i < 3; ++i) ok = ok && (o[i] < range[3 + i]) && (o[i] >= range[i]);
            if (mask) mask[(size_t)n * P + p] = (uint8_t)ok;
            /* lidar2map (splat_projection.py:81-88): x_map = -y - ymin... literally
             * [0,-1,0,-xmin; -1,0,0,-ymin]: one inexact add each, then true division. */
            const float xm = -range[0] + (-o[1]);
            const float ym = -range[1] + (-o[0]);
            if (xy) {
                xy[((size_t)n * P + p) * 2 + 0] = xm / voxel[0];
                xy[((size_t)n * P + p) * 2 + 1] = ym / voxel[1];
            }
        }
    }
    return 0;
}

/* ---- bilinear splat (splat_projection.py:262-354), scatter_mode='mean'.
 * xy [N,P,2]; feats [N,F,P] (already multiplied by the mask); outputs vol [N,F,H*W],
 * dens [N,H*W], idx [N,P,4] int64 linear cell index per tap in (xdiff,ydiff) loop order
 * ((0,0),(0,1),(1,0),(1,1)), -1 where the tap is out of bounds; wts [N,P,4]. */
int oracle_splat_soft(const float* xy, const float* feats, int N, int P, int F, int H, int W,
                      float min_weight, float* vol, float* dens, int64_t* idx, float* wts) {
    const size_t G = (size_t)H * W;
    if (vol) memset(vol, 0, (size_t)N * F * G * sizeof(float));
    if (dens) memset(dens, 0, (size_t)N * G * sizeof(float));
    for (int n = 0; n < N; ++n) {
        for (int p = 0; p < P; ++p) {
            const float X = xy[((size_t)n * P + p) * 2 + 0], Y = xy[((size_t)n * P + p) * 2 + 1];
            const float fX = floorf(X), fY = floorf(Y);
            const long X0 = (long)fX, Y0 = (long)fY;
            const float rX = X - (float)X0, rY = Y - (float)Y0;
            int t = 0;
            for (int dx = 0; dx < 2; ++dx) {
                const float wX = (float)(1 - dx) + (float)(2 * dx - 1) * rX;
                for (int dy = 0; dy < 2; ++dy, ++t) {
                    const float wY = (float)(1 - dy) + (float)(2 * dy - 1) * rY;
                    const float w = wX * wY;
                    const long X_ = X0 + dx, Y_ = Y0 + dy;
                    const int valid = (0 <= X_) && (X_ < W) && (0 <= Y_) && (Y_ < H);
                    const long id = Y_ * W + X_;
                    if (idx) idx[((size_t)n * P + p) * 4 + t] = valid ? id : -1;
                    if (wts) wts[((size_t)n * P + p) * 4 + t] = valid ? w : 0.0f;
                    if (!valid) continue;
                    if (dens) dens[(size_t)n * G + id] += w;
                    if (vol && feats)
                        for (int f = 0; f < F; ++f)
                            vol[((size_t)n * F + f) * G + id] += w * feats[((size_t)n * F + f) * P + p];
                }
            }
        }
        if (vol && dens)
            for (int f = 0; f < F; ++f)
                for (size_t g = 0; g < G; ++g) {
                    const float d = dens[(size_t)n * G + g];
                    vol[((size_t)n * F + f) * G + g] /= (d < min_weight ? min_weight : d);
                }
    }
    return 0;
}

/* ---- LiDAR -> sparse depth raster (projection.py:88-134; build_dense_depth.py:461-463).
 * pc [npts, stride] float32 (xyz first), P [3,4] float64 row-major (lidar2camrect).
 * depth_m [H,W] float32 = per-pixel max z (0 where empty); depth_mm [H,W] float32 =
 * float(uint16(clip(depth_m*1000, 0, 65535))) as channel 3 of the network input is built. */
int oracle_lidar_raster(const float* pc, int npts, int stride, const double* P, int H, int W,
                        float* depth_m, float* depth_mm) {
    const size_t G = (size_t)H * W;
    double* zmax = (double*)calloc(G, sizeof(double));
    if (!zmax) return 1;
    for (int i = 0; i < npts; ++i) {
        const double x = pc[(size_t)i * stride], y = pc[(size_t)i * stride + 1],
                     z = pc[(size_t)i * stride + 2];
        double c[3];
        for (int r = 0; r < 3; ++r) {
            double acc = P[r * 4 + 0] * x;
            acc = fma(P[r * 4 + 1], y, acc);
            acc = fma(P[r * 4 + 2], z, acc);
            acc = fma(P[r * 4 + 3], 1.0, acc);
            c[r] = acc;
        }
        double u = c[0] / c[2], v = c[1] / c[2];
        /* np.clip to int32 range then astype(int32): truncation toward zero; NaN -> INT_MIN */
        if (!(c[2] > 0.0)) continue;
        if (u != u || v != v) continue;
        if (u > 2147483647.0) u = 2147483647.0;
        if (u < -2147483648.0) u = -2147483648.0;
        if (v > 2147483647.0) v = 2147483647.0;
        if (v < -2147483648.0) v = -2147483648.0;
        const long ui = (long)u, vi = (long)v; /* C cast truncates toward zero */
        if (ui < 0 || ui >= W || vi < 0 || vi >= H) continue;
        const size_t g = (size_t)vi * W + ui;
        if (c[2] > zmax[g]) zmax[g] = c[2];
    }
    for (size_t g = 0; g < G; ++g) {
        const float d = (float)zmax[g];
        if (depth_m) depth_m[g] = d;
        if (depth_mm) {
            float mm = d * 1000.0f;
            mm = mm < 0.0f ? 0.0f : (mm > 65535.0f ? 65535.0f : mm);
            depth_mm[g] = (float)(uint16_t)mm;
        }
    }
    free(zmax);
    return 0;
}

/* ---- depth expectation (depth_utils.py:300-313, depth.py:70,100).  logits [N,D,P] -> metric [N,P]
 * (metres), bins [N,P] int64 (first arg-max). */
int oracle_depth_expectation(const float* logits, int N, int D, int P, float dmin, float dmax,
                             float* metric, int64_t* bins) {
    for (int n = 0; n < N; ++n)
#pragma omp parallel for schedule(static)
        for (long p = 0; p < (long)P; ++p) {
            const float* L = logits + (size_t)n * D * P + p;
            float m = L[0];
            int arg = 0;
            for (int k = 1; k < D; ++k)
                if (L[(size_t)k * P] > m) { m = L[(size_t)k * P]; arg = k; }
            double s = 0.0, e = 0.0;
            const float step = (dmax - dmin) / (float)(D - 1);
            for (int k = 0; k < D; ++k) {
                /* torch.linspace: symmetric evaluation around the midpoint */
                const float val = (k < D / 2) ? dmin + step * (float)k : dmax - step * (float)(D - 1 - k);
                const float ex = expf(L[(size_t)k * P] - m);
                s += ex;
                e += (double)ex * (double)val;
            }
            if (metric) metric[(size_t)n * P + p] = (float)(e / s) / 1000.0f;
            if (bins) bins[(size_t)n * P + p] = arg;
        }
    return 0;
}

/* ---- expert / counterfactual visitation raster (loss_utils.py:1055-1116, 2nd definition).
 * rc [B,T,2] float64 (row, col) un-pooled cells; counts [B,H,W] float32 in {0,1}. fp32 path:
 * the expert poses are float32 tensors; the counterfactual trajectories are float64 -- the
 * caller picks by `is_f64`. */
int oracle_expert_visitation(const double* rc, int B, int T, double map_ds, int H, int W,
                             int is_f64, float* counts) {
    const size_t G = (size_t)H * W;
    memset(counts, 0, (size_t)B * G * sizeof(float));
    /* max_steps is a global (whole batch) quantity: ceil(max segment length) */
    long max_steps = 0;
    for (int b = 0; b < B; ++b)
        for (int t = 0; t + 1 < T; ++t) {
            double d;
            if (is_f64) {
                const double r0 = rc[((size_t)b * T + t) * 2] / map_ds, c0 = rc[((size_t)b * T + t) * 2 + 1] / map_ds;
                const double r1 = rc[((size_t)b * T + t + 1) * 2] / map_ds, c1 = rc[((size_t)b * T + t + 1) * 2 + 1] / map_ds;
                d = sqrt((r1 - r0) * (r1 - r0) + (c1 - c0) * (c1 - c0));
            } else {
                const float r0 = (float)rc[((size_t)b * T + t) * 2] / (float)map_ds, c0 = (float)rc[((size_t)b * T + t) * 2 + 1] / (float)map_ds;
                const float r1 = (float)rc[((size_t)b * T + t + 1) * 2] / (float)map_ds, c1 = (float)rc[((size_t)b * T + t + 1) * 2 + 1] / (float)map_ds;
                d = sqrtf((r1 - r0) * (r1 - r0) + (c1 - c0) * (c1 - c0));
            }
            const long s = (long)ceil(d);
            if (s > max_steps) max_steps = s;
        }
    for (int b = 0; b < B; ++b) {
        float* C = counts + (size_t)b * G;
        for (int t = 0; t < T; ++t) {
            const int last = (t == T - 1);
            const long ns = last ? 1 : max_steps;
            for (long k = 0; k < ns; ++k) {
                double pr, pc;
                if (is_f64) {
                    const double r0 = rc[((size_t)b * T + t) * 2] / map_ds, c0 = rc[((size_t)b * T + t) * 2 + 1] / map_ds;
                    if (last) { pr = r0; pc = c0; }
                    else {
                        const double r1 = rc[((size_t)b * T + t + 1) * 2] / map_ds, c1 = rc[((size_t)b * T + t + 1) * 2 + 1] / map_ds;
                        /* torch.linspace(0,1,n) is float32 even for float64 trajectories:
                         * step=(1-0)/(n-1); first half start+step*i, second half end-step*(n-1-i) */
                        float ff;
                        if (max_steps == 1) ff = 0.0f;
                        else {
                            const float step = 1.0f / (float)(max_steps - 1);
                            ff = (k < max_steps / 2) ? step * (float)k : 1.0f - step * (float)(max_steps - 1 - k);
                        }
                        const double f = (double)ff;
                        pr = r0 + f * (r1 - r0); pc = c0 + f * (c1 - c0);
                    }
                } else {
                    const float r0 = (float)rc[((size_t)b * T + t) * 2] / (float)map_ds, c0 = (float)rc[((size_t)b * T + t) * 2 + 1] / (float)map_ds;
                    if (last) { pr = r0; pc = c0; }
                    else {
                        const float r1 = (float)rc[((size_t)b * T + t + 1) * 2] / (float)map_ds, c1 = (float)rc[((size_t)b * T + t + 1) * 2 + 1] / (float)map_ds;
                        float f;
                        if (max_steps == 1) f = 0.0f;
                        else {
                            const float step = 1.0f / (float)(max_steps - 1);
                            f = (k < max_steps / 2) ? step * (float)k : 1.0f - step * (float)(max_steps - 1 - k);
                        }
                        const float fr = r0 + f * (r1 - r0), fc = c0 + f * (c1 - c0);
                        pr = fr; pc = fc;
                    }
                }
                double cr = pr < 0 ? 0 : (pr > H - 1 ? H - 1 : pr);
                double cc = pc < 0 ? 0 : (pc > W - 1 ? W - 1 : pc);
                const long ir = (long)cr, ic = (long)cc;
                C[ir * W + ic] = 1.0f; /* scatter_add of ones then clip to 1 */
            }
        }
    }
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
