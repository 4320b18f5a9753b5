// N3: backward kernels of the frame path (a14..a17) so that a torch frame-predicting network trains through the
// point-cloud / n-body canonicalizers as it does in the reference (torch autograd through gram_schmidt,
// common/utils.py:22-51; bmm in pointcloud/canonicalization/continuous_group.py:77-79; the row products of
// nbody/canonicalization/euclidean_group.py:114-122, :133-136, :150-157).  All per-sample 3x3 algebra in registers.
#include "common.cuh"

namespace eqb {

// fp64 registers: an ill-conditioned triple (nearly parallel rows) divides by a tiny norm twice
struct V3 { double x, y, z; };
__device__ __forceinline__ V3 v3(double x, double y, double z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// u = w / |w|:  dw = (du - u (u . du)) / |w|
__device__ __forceinline__ V3 normalize_bwd(V3 u, double norm, V3 du) { return (1.0 / norm) * (du - dot(u, du) * u); }

// reverse mode through gs3 (small_ops.cu): rows a, b, c of R from rows v1, v2, v3
__global__ void gram_schmidt3_backward_kernel(const float *__restrict__ v, const float *__restrict__ dR, float *__restrict__ dv,
                                              int B, int modified) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B) return;
    const float *p = v + (size_t)s * 9, *g = dR + (size_t)s * 9;
    const V3 v1 = v3(p[0], p[1], p[2]), v2 = v3(p[3], p[4], p[5]), v3_ = v3(p[6], p[7], p[8]);
    // forward, keeping the intermediates
    const double n1 = sqrt(dot(v1, v1));
    const V3 a = (1.0 / n1) * v1;
    const double d = dot(v2, a);
    const V3 bp = v2 - d * a;
    const double n2 = sqrt(dot(bp, bp));
    const V3 b = (1.0 / n2) * bp;
    const double d1 = dot(v3_, a);
    V3 c1 = v3_ - d1 * a;                       // modified: the once-projected third vector
    const double d2 = modified ? dot(c1, b) : dot(v3_, b);
    const V3 cp = c1 - d2 * b;
    const double n3 = sqrt(dot(cp, cp));
    const V3 c = (1.0 / n3) * cp;
    // reverse
    V3 da = v3(g[0], g[1], g[2]), db = v3(g[3], g[4], g[5]);
    const V3 dc = v3(g[6], g[7], g[8]);
    const V3 dcp = normalize_bwd(c, n3, dc);
    V3 dc1 = dcp;                               // cp = c1 - d2 b
    const double dd2 = -dot(b, dcp);
    db = db - d2 * dcp;
    V3 dv3;
    if (modified) {                             // d2 = c1 . b
        dc1 = dc1 + dd2 * b;
        db = db + dd2 * c1;
        dv3 = dc1;
    } else {                                    // d2 = v3 . b
        db = db + dd2 * v3_;
        dv3 = dc1 + dd2 * b;
    }
    const double dd1 = -dot(a, dc1);             // c1 = v3 - d1 a
    da = da - d1 * dc1;
    dv3 = dv3 + dd1 * a;                        // d1 = v3 . a
    da = da + dd1 * v3_;
    const V3 dbp = normalize_bwd(b, n2, db);
    V3 dv2 = dbp;                               // bp = v2 - d a
    const double dd = -dot(a, dbp);
    da = da - d * dbp;
    dv2 = dv2 + dd * a;                         // d = v2 . a
    da = da + dd * v2;
    const V3 dv1 = normalize_bwd(a, n1, da);
    float *o = dv + (size_t)s * 9;
    o[0] = (float)dv1.x; o[1] = (float)dv1.y; o[2] = (float)dv1.z;
    o[3] = (float)dv2.x; o[4] = (float)dv2.y; o[5] = (float)dv2.z;
    o[6] = (float)dv3.x; o[7] = (float)dv3.y; o[8] = (float)dv3.z;
}

// y_j[n] = sum_k R[j][k] x_k[n]  ->  dx_k[n] = sum_j R[j][k] dy_j[n],  dR[j][k] = sum_n dy_j[n] x_k[n]
__global__ void __launch_bounds__(256) so3_apply_backward_kernel(const float *__restrict__ x, const float *__restrict__ R,
                                                                 const float *__restrict__ dy, float *__restrict__ dx,
                                                                 float *__restrict__ dR, int N, int chunks) {
    const int b = blockIdx.x / chunks, ch = blockIdx.x % chunks;
    float r[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) r[i] = __ldg(R + (size_t)b * 9 + i);
    const float *xb = x + (size_t)b * 3 * N, *gb = dy + (size_t)b * 3 * N;
    float acc[9] = {};
    for (int n = ch * blockDim.x + threadIdx.x; n < N; n += chunks * blockDim.x) {
        const float g0 = gb[n], g1 = gb[N + n], g2 = gb[2 * N + n];
        if (dx) {
            float *db = dx + (size_t)b * 3 * N;
            db[n] = g0 * r[0] + g1 * r[3] + g2 * r[6];
            db[N + n] = g0 * r[1] + g1 * r[4] + g2 * r[7];
            db[2 * N + n] = g0 * r[2] + g1 * r[5] + g2 * r[8];
        }
        if (dR) {
            const float p0 = xb[n], p1 = xb[N + n], p2 = xb[2 * N + n];
            acc[0] += g0 * p0; acc[1] += g0 * p1; acc[2] += g0 * p2;
            acc[3] += g1 * p0; acc[4] += g1 * p1; acc[5] += g1 * p2;
            acc[6] += g2 * p0; acc[7] += g2 * p1; acc[8] += g2 * p2;
        }
    }
    if (!dR) return;
    __shared__ float red[8][9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        float a = acc[i];
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
        if (lane == 0) red[warp][i] = a;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        atomicAdd(dR + (size_t)b * 9 + threadIdx.x, t);
    }
}

// e3_apply: lc_j = sum_k (loc_k - t_k) R[j][k], vc_j = sum_k vel_k R[j][k]   (one row per thread, own R and t)
__global__ void __launch_bounds__(256) e3_apply_backward_kernel(const float *__restrict__ loc, const float *__restrict__ vel,
                                                                const float *__restrict__ R, const float *__restrict__ t,
                                                                const float *__restrict__ dlc, const float *__restrict__ dvc,
                                                                float *__restrict__ dloc, float *__restrict__ dvel,
                                                                float *__restrict__ dR, float *__restrict__ dt, int M) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float r[9], gl[3], gv[3], l[3], v[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) r[i] = R[(size_t)m * 9 + i];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        gl[k] = dlc ? dlc[3 * m + k] : 0.f;
        gv[k] = dvc ? dvc[3 * m + k] : 0.f;
        l[k] = loc[3 * m + k] - t[3 * m + k];
        v[k] = vel[3 * m + k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float gk = gl[0] * r[k] + gl[1] * r[3 + k] + gl[2] * r[6 + k];
        if (dloc) dloc[3 * m + k] = gk;
        if (dt) dt[3 * m + k] = -gk;
        if (dvel) dvel[3 * m + k] = gv[0] * r[k] + gv[1] * r[3 + k] + gv[2] * r[6 + k];
    }
    if (dR) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) dR[(size_t)m * 9 + 3 * j + k] = gl[j] * l[k] + gv[j] * v[k];
    }
}

// e3_invert: y_j = sum_k x_k R[k][j] + t_j
__global__ void __launch_bounds__(256) e3_invert_backward_kernel(const float *__restrict__ x, const float *__restrict__ R,
                                                                 const float *__restrict__ dy, float *__restrict__ dx,
                                                                 float *__restrict__ dR, float *__restrict__ dt, int M) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float g[3], xv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { g[k] = dy[3 * m + k]; xv[k] = x[3 * m + k]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (dx) dx[3 * m + k] = g[0] * R[(size_t)m * 9 + 3 * k] + g[1] * R[(size_t)m * 9 + 3 * k + 1] + g[2] * R[(size_t)m * 9 + 3 * k + 2];
        if (dt) dt[3 * m + k] = g[k];
        if (dR) {
#pragma unroll
            for (int j = 0; j < 3; ++j) dR[(size_t)m * 9 + 3 * k + j] = xv[k] * g[j];
        }
    }
}

}  // namespace eqb

using namespace eqb;

extern "C" int eqb_gram_schmidt3_backward(const float *v, const float *dR, float *dv, int B, int modified, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0, "eqb_gram_schmidt3_backward: bad batch");
    if (B == 0) return 0;
    EQB_REQUIRE(v && dR && dv, "eqb_gram_schmidt3_backward: null pointer");
    gram_schmidt3_backward_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(v, dR, dv, B, modified);
    return finish_launch("eqb_gram_schmidt3_backward");
}

extern "C" int eqb_so3_apply_backward(const float *x, const float *R, const float *dy, float *dx, float *dR, int B, int N,
                                      void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && N >= 0, "eqb_so3_apply_backward: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (dR && B > 0) EQB_CUDA(cudaMemsetAsync(dR, 0, (size_t)B * 9 * sizeof(float), st));
    if (B == 0 || N == 0) return 0;
    EQB_REQUIRE(x && R && dy && (dx || dR), "eqb_so3_apply_backward: null pointer");
    int chunks = (N + 1023) / 1024;
    if (chunks < 1) chunks = 1;
    EQB_REQUIRE((long long)B * chunks < (1LL << 31), "eqb_so3_apply_backward: grid too large");
    so3_apply_backward_kernel<<<(unsigned)(B * chunks), 256, 0, st>>>(x, R, dy, dx, dR, N, chunks);
    return finish_launch("eqb_so3_apply_backward");
}

extern "C" int eqb_e3_apply_backward(const float *loc, const float *vel, const float *R, const float *t, const float *dloc_c,
                                     const float *dvel_c, float *dloc, float *dvel, float *dR, float *dt, int M, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(M >= 0, "eqb_e3_apply_backward: bad row count");
    if (M == 0) return 0;
    EQB_REQUIRE(loc && vel && R && t && (dloc_c || dvel_c), "eqb_e3_apply_backward: null pointer");
    e3_apply_backward_kernel<<<(M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(loc, vel, R, t, dloc_c, dvel_c, dloc, dvel, dR, dt, M);
    return finish_launch("eqb_e3_apply_backward");
}

extern "C" int eqb_e3_invert_backward(const float *x, const float *R, const float *dy, float *dx, float *dR, float *dt, int M,
                                      void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(M >= 0, "eqb_e3_invert_backward: bad row count");
    if (M == 0) return 0;
    EQB_REQUIRE(x && R && dy, "eqb_e3_invert_backward: null pointer");
    e3_invert_backward_kernel<<<(M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, R, dy, dx, dR, dt, M);
    return finish_launch("eqb_e3_invert_backward");
}
