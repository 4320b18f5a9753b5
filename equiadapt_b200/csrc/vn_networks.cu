// N1: the frame-predicting vector-neuron networks as fused kernels (eval mode).
//
// VNSmall (pointcloud/canonicalization_networks/equivariant_networks.py:79-150, vector_neuron_layers.py:210-324):
//   kNN graph (k nearest by -|xi|^2 + 2 xi.xj - |xj|^2, :15-33) -> per edge the three vector channels
//   [xj - xi, xi, xj x xi] (:36-76) -> VNLinearLeakyReLU(3 -> 21) -> mean over the k neighbours ->
//   VNLinearLeakyReLU(21 -> 21) -> VNBatchNorm(21) -> VNLinearLeakyReLU(21 -> 4) -> (dropout = identity in eval) ->
//   mean over the points -> first three channels = three 3-vectors per cloud.
// The reference materialises (B, 9, N, k), (B, 21, 3, N, k) ... tensors (5.2 MB per cloud and tensor at N = 1024,
// k = 20) and spends > 99 % of the point-cloud configuration there.  Here one CTA owns one cloud:
// the cloud lives in shared memory, every thread walks its points, keeps the k best neighbours in registers, pushes
// each edge through the first layer, averages, runs the two point-wise layers and adds its result to the CTA's sum.
// Nothing but the cloud is read and nothing but 9 floats per cloud is written.
//
// VNDeepSets (nbody/canonicalization_networks/custom_equivariant_networks.py:13-252): see the second half of the file.
#include <stdlib.h>

#include "common.cuh"

namespace eqb {

constexpr int VN_C1 = 21;   // 64 // 3 hidden vector channels (fixed by the reference, equivariant_networks.py:115-118)
constexpr int VN_C2 = 4;    // 12 // 3 output vector channels, the first 3 are used (:150)
constexpr float VN_EPS = 1e-6f;
constexpr int VN_MAX_THREADS = 512;
constexpr int VN_CAP = 16;    // pending neighbour candidates per lane between two insertion rounds

// flat parameter block (floats), raw tensors of the reference module in this order:
//   conv_pos: map_to_feat (21x3), map_to_dir (21x3), batchnorm.bn2d {weight, bias, running_mean, running_var} (4x21)
//   conv1:    map_to_feat (21x21), map_to_dir (21x21), batchnorm.bn1d {w, b, rm, rv} (4x21)
//   bn1:      bn1d {w, b, rm, rv} (4x21)
//   conv2:    map_to_feat (4x21), map_to_dir (4x21), batchnorm.bn1d {w, b, rm, rv} (4x4)
constexpr int VP_F0 = 0, VP_D0 = VP_F0 + VN_C1 * 3, VP_BN0 = VP_D0 + VN_C1 * 3;
constexpr int VP_F1 = VP_BN0 + 4 * VN_C1, VP_D1 = VP_F1 + VN_C1 * VN_C1, VP_BN1C = VP_D1 + VN_C1 * VN_C1;
constexpr int VP_BN1 = VP_BN1C + 4 * VN_C1;
constexpr int VP_F2 = VP_BN1 + 4 * VN_C1, VP_D2 = VP_F2 + VN_C2 * VN_C1, VP_BN2 = VP_D2 + VN_C2 * VN_C1;
constexpr int VP_TOTAL = VP_BN2 + 4 * VN_C2;

// VNBatchNorm in eval mode (vector_neuron_layers.py:309-322): v <- v / (|v| + EPS) * ((|v| + EPS - rm) * s + b),
// s = weight / sqrt(running_var + eps); (s, t = b - rm * s) are folded when the parameters are staged.
__device__ __forceinline__ void vn_bn(float v[3], float s, float t) {
    const float norm = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) + VN_EPS;
    const float f = (norm * s + t) / norm;
    v[0] *= f; v[1] *= f; v[2] *= f;
}
// VN leaky ReLU with negative_slope 0 (vector_neuron_layers.py:267-273): p <- p - (p.d / (|d|^2 + EPS)) d where p.d < 0
__device__ __forceinline__ void vn_relu(float p[3], const float d[3]) {
    const float dot = p[0] * d[0] + p[1] * d[1] + p[2] * d[2];
    if (dot < 0.f) {
        const float f = dot / (d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + VN_EPS);
        p[0] -= f * d[0]; p[1] -= f * d[1]; p[2] -= f * d[2];
    }
}

template <int K, int VN_THREADS>
__global__ void __launch_bounds__(VN_THREADS, 1) vnsmall_kernel(const float *__restrict__ x, const float *__restrict__ prm,
                                                                double *__restrict__ part, int N, int k, float bn_eps) {
    extern __shared__ __align__(16) float vsm[];
    float *xs = vsm;               // [3][N]
    float *xx = xs + 3 * N;        // [N]  |x_i|^2
    float *P = xx + N;             // staged parameters, batch norms folded to (scale, shift) pairs
    int *nbr = reinterpret_cast<int *>(P + VP_TOTAL);   // [K][VN_THREADS] neighbour lists (dynamic indexing lives here)
    float *bufv = reinterpret_cast<float *>(nbr + K * VN_THREADS);   // [VN_CAP][VN_THREADS] pending candidates: value,
    int *bufi = reinterpret_cast<int *>(bufv + VN_CAP * VN_THREADS); //                       index
    float4 *x4 = reinterpret_cast<float4 *>(bufi + VN_CAP * VN_THREADS);   // [N] (x, y, z, |x|^2): one 16-byte broadcast per candidate
    __shared__ double red[VN_THREADS / 32][9];
    // grid (clouds, splits): CTA (b, sp) owns points [sp * per, (sp + 1) * per) of cloud b - small batches (the 16 clouds
    // per GPU of BASELINE configs[3]) still fill the SMs; every CTA stages the whole cloud (all points are candidates)
    const int b = blockIdx.x, tid = threadIdx.x;
    const int per = (N + (int)gridDim.y - 1) / (int)gridDim.y;
    const int i_lo = (int)blockIdx.y * per, i_hi = min(N, i_lo + per);
    const float *xb = x + (size_t)b * 3 * N;
    for (int i = tid; i < 3 * N; i += VN_THREADS) xs[i] = xb[i];
    for (int i = tid; i < VP_TOTAL; i += VN_THREADS) P[i] = prm[i];
    __syncthreads();
    // fold the four batch norms in place: {w, b, rm, rv} -> {s, t, -, -}
    auto fold = [&](int off, int C) {
        for (int c = tid; c < C; c += VN_THREADS) {
            const float w = P[off + c], bb = P[off + C + c], rm = P[off + 2 * C + c], rv = P[off + 3 * C + c];
            const float s = w / sqrtf(rv + bn_eps);
            P[off + c] = s;
            P[off + C + c] = bb - rm * s;
        }
    };
    fold(VP_BN0, VN_C1); fold(VP_BN1C, VN_C1); fold(VP_BN1, VN_C1); fold(VP_BN2, VN_C2);
    for (int i = tid; i < N; i += VN_THREADS) {
        const float a0 = xs[i], a1 = xs[N + i], a2 = xs[2 * N + i];
        xx[i] = a0 * a0 + a1 * a1 + a2 * a2;   // torch.sum(x**2, dim=1): ((a0^2 + a1^2) + a2^2)
        x4[i] = make_float4(a0, a1, a2, xx[i]);
    }
    __syncthreads();

    float total[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) total[q] = 0.f;

    // (warp-uniform trip count: lanes past the end of the split redo its last point and are left out of the sum)
    for (int i0 = i_lo; i0 < i_hi; i0 += VN_THREADS) {
        const bool live = i0 + tid < i_hi;
        const int i = live ? i0 + tid : i_hi - 1;
        const float xi0 = xs[i], xi1 = xs[N + i], xi2 = xs[2 * N + i], xxi = xx[i];
        // ---- k nearest neighbours: the k largest of (-xx_j - inner_ij) - xx_i, inner = -2 x_i.x_j (:28-32) ------------
        // A sorted insertion costs ~6 K instructions and a warp pays for it whenever ANY lane inserts -- with one point per
        // lane that is nearly every candidate (r3o profile: 60 % of the kernel's instructions).  So candidates that beat the
        // lane's k-th best are only APPENDED to a small per-lane buffer in shared memory; the warp runs the insertions when
        // some lane's buffer is full (and once at the end).  Same neighbours in the same order (ties: lower index first).
        float val[K];
        int id[K];
#pragma unroll
        for (int s = 0; s < K; ++s) { val[s] = -INFINITY; id[s] = i; }
        float thr = -INFINITY;
        int cnt = 0;
        float *bv = bufv + tid;               // this lane's next free slot (slots are VN_THREADS apart)
        int *bi = bufi + tid;
        auto flush = [&]() {
#pragma unroll 1
            for (int e = 0; e < VN_CAP; ++e) {
                if (e < cnt) {
                    const float cv = bufv[e * VN_THREADS + tid];
                    const int ci = bufi[e * VN_THREADS + tid];
                    if (cv > thr) {
                        // shift-insert into the descending list, bottom up: slot s takes its upper neighbour when the
                        // candidate beats that one, the candidate when it beats only slot s, else it keeps its value
                        bool above[K];
#pragma unroll
                        for (int s = 0; s < K; ++s) above[s] = cv > val[s];
#pragma unroll
                        for (int s = K - 1; s >= 1; --s) {
                            val[s] = above[s - 1] ? val[s - 1] : (above[s] ? cv : val[s]);
                            id[s] = above[s - 1] ? id[s - 1] : (above[s] ? ci : id[s]);
                        }
                        if (above[0]) { val[0] = cv; id[0] = ci; }
                        // the admission threshold is the k-th best (k <= K: static indexing keeps the list in registers)
#pragma unroll
                        for (int s = 0; s < K; ++s)
                            if (s == k - 1) thr = val[s];
                    }
                }
            }
            cnt = 0;
            bv = bufv + tid;
            bi = bufi + tid;
        };
        auto candidate = [&](const float4 c, int j) {
            const float m = fmaf(xi2, c.z, fmaf(xi1, c.y, xi0 * c.x));
            const float pd = (-c.w - (-2.f * m)) - xxi;
            if (pd > thr) {
                *bv = pd;
                *bi = j;
                bv += VN_THREADS;
                bi += VN_THREADS;
                ++cnt;
            }
        };
        // four candidates per round; a round starts with room for four in every lane's buffer.  ONE call site for the
        // insertion rounds: the unrolled K-slot insertion is ~1 000 instructions and the kernel is already bound by
        // instruction fetch when it is inlined three times (r3q profile)
#pragma unroll 1
        for (int j = 0;; j += 4) {
            const bool done = j >= N;
            if (done || __any_sync(0xffffffffu, cnt > VN_CAP - 4)) {
                flush();
                if (done) break;
            }
            if (j + 4 <= N) {
                const float4 c0 = x4[j], c1 = x4[j + 1], c2 = x4[j + 2], c3 = x4[j + 3];   // (same address in every lane: broadcast)
                candidate(c0, j);
                candidate(c1, j + 1);
                candidate(c2, j + 2);
                candidate(c3, j + 3);
            } else {
                for (int q = j; q < N; ++q) candidate(x4[q], q);
            }
        }
        // ---- edges -> VNLinearLeakyReLU(3 -> 21) -> mean over the neighbours ----------------------------------------
        float h[VN_C1][3];
#pragma unroll
        for (int c = 0; c < VN_C1; ++c) h[c][0] = h[c][1] = h[c][2] = 0.f;
#pragma unroll
        for (int s = 0; s < K; ++s) nbr[s * VN_THREADS + tid] = id[s];
#pragma unroll 1
        for (int s = 0; s < k; ++s) {
            const int j = nbr[s * VN_THREADS + tid];
            const float xj0 = xs[j], xj1 = xs[N + j], xj2 = xs[2 * N + j];
            // channels: feature - x, x, cross(feature, x)  (:69-73)
            const float f0[3] = {xj0 - xi0, xj1 - xi1, xj2 - xi2};
            const float f2[3] = {xj1 * xi2 - xj2 * xi1, xj2 * xi0 - xj0 * xi2, xj0 * xi1 - xj1 * xi0};
#pragma unroll
            for (int c = 0; c < VN_C1; ++c) {
                const float wf0 = P[VP_F0 + 3 * c], wf1 = P[VP_F0 + 3 * c + 1], wf2 = P[VP_F0 + 3 * c + 2];
                const float wd0 = P[VP_D0 + 3 * c], wd1 = P[VP_D0 + 3 * c + 1], wd2 = P[VP_D0 + 3 * c + 2];
                float p[3] = {fmaf(wf2, f2[0], fmaf(wf1, xi0, wf0 * f0[0])), fmaf(wf2, f2[1], fmaf(wf1, xi1, wf0 * f0[1])),
                              fmaf(wf2, f2[2], fmaf(wf1, xi2, wf0 * f0[2]))};
                const float d[3] = {fmaf(wd2, f2[0], fmaf(wd1, xi0, wd0 * f0[0])), fmaf(wd2, f2[1], fmaf(wd1, xi1, wd0 * f0[1])),
                                    fmaf(wd2, f2[2], fmaf(wd1, xi2, wd0 * f0[2]))};
                vn_bn(p, P[VP_BN0 + c], P[VP_BN0 + VN_C1 + c]);
                vn_relu(p, d);
                h[c][0] += p[0]; h[c][1] += p[1]; h[c][2] += p[2];
            }
        }
        const float invk = 1.f / (float)k;
#pragma unroll
        for (int c = 0; c < VN_C1; ++c) { h[c][0] *= invk; h[c][1] *= invk; h[c][2] *= invk; }
        // ---- VNLinearLeakyReLU(21 -> 21) -> VNBatchNorm -> VNLinearLeakyReLU(21 -> 4), channel by channel ------------
        float p2[3][3], d2[3][3];
#pragma unroll
        for (int c2 = 0; c2 < 3; ++c2)
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) p2[c2][dd] = d2[c2][dd] = 0.f;
#pragma unroll 1
        for (int c = 0; c < VN_C1; ++c) {
            float p[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int m = 0; m < VN_C1; ++m) {
                const float wf = P[VP_F1 + c * VN_C1 + m], wd = P[VP_D1 + c * VN_C1 + m];
                p[0] = fmaf(wf, h[m][0], p[0]); p[1] = fmaf(wf, h[m][1], p[1]); p[2] = fmaf(wf, h[m][2], p[2]);
                d[0] = fmaf(wd, h[m][0], d[0]); d[1] = fmaf(wd, h[m][1], d[1]); d[2] = fmaf(wd, h[m][2], d[2]);
            }
            vn_bn(p, P[VP_BN1C + c], P[VP_BN1C + VN_C1 + c]);
            vn_relu(p, d);
            vn_bn(p, P[VP_BN1 + c], P[VP_BN1 + VN_C1 + c]);
#pragma unroll
            for (int c2 = 0; c2 < 3; ++c2) {
                const float wf = P[VP_F2 + c2 * VN_C1 + c], wd = P[VP_D2 + c2 * VN_C1 + c];
#pragma unroll
                for (int dd = 0; dd < 3; ++dd) {
                    p2[c2][dd] = fmaf(wf, p[dd], p2[c2][dd]);
                    d2[c2][dd] = fmaf(wd, p[dd], d2[c2][dd]);
                }
            }
        }
#pragma unroll
        for (int c2 = 0; c2 < 3; ++c2) {
            vn_bn(p2[c2], P[VP_BN2 + c2], P[VP_BN2 + VN_C2 + c2]);
            vn_relu(p2[c2], d2[c2]);
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) total[3 * c2 + dd] += live ? p2[c2][dd] : 0.f;
        }
    }
    // ---- mean over the points (:150) -----------------------------------------------------------------------------
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        double v = (double)total[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[tid >> 5][q] = v;
    }
    __syncthreads();
    if (tid < 9) {
        double v = 0.0;
        for (int w = 0; w < VN_THREADS / 32; ++w) v += red[w][tid];
        part[((size_t)b * gridDim.y + blockIdx.y) * 9 + tid] = v;
    }
}

// out[b][q] = (sum over the splits of cloud b) / N, in split order (deterministic)
__global__ void vnsmall_finish_kernel(const double *__restrict__ part, int B, int S, int N, float *__restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * 9) return;
    const int b = t / 9, q = t - b * 9;
    double v = 0.0;
    for (int s = 0; s < S; ++s) v += part[((size_t)b * S + s) * 9 + q];
    out[t] = (float)(v / (double)N);
}


// =====================================================================================================================
// VNDeepSets (nbody/canonicalization_networks/custom_equivariant_networks.py:13-252; VNLeakyReLU / VNSoftplus of
// nbody/canonicalization_networks/custom_group_equivariant_layers.py:7-99), eval mode:
//   mean_loc = scatter(loc, system, reduce=layer_pooling); features = stack(canon_feature channels) (M, 3, Cin)   (:129-158)
//   L x VNDeepSetLayer: identity_linear(x) + pooling_linear(scatter(x[edges[0]], edges[1], reduce=pooling))
//                       -> VN nonlinearity with its own map_to_dir -> (dropout = identity) -> + x (layers after the first)
//   scatter over the 5 particles of a system (final_pooling) -> output_layer (H -> 4) -> rotation vectors (3 x 3) and
//   translation (+ mean_loc), repeated for the 5 particle rows of the system                                   (:159-172)
// Like the reference (batch_indices = arange(batch).repeat(1, 5), :129-131) systems are 5 consecutive rows.  Edges
// must stay inside a system (true for the K5 graphs of examples/nbody/model_utils.py:60-89); a tiny pre-kernel turns
// the edge list into 5 x 5 multiplicity tables and counts edges that do not.  One thread per particle, 32 systems per
// CTA, features ping-pong through shared memory between layers; the reference's ~40 launches and (M, 3, H)
// intermediates per layer become one launch that reads loc / vel / charges and writes (R vectors, t).
// =====================================================================================================================
constexpr int DS_P = 5;            // particles per system
constexpr int DS_SYS = 32;         // systems per CTA
constexpr int DS_THREADS = DS_P * DS_SYS;
enum { DS_RELU = 0, DS_LEAKY = 1, DS_SOFTPLUS = 2 };

struct DeepSetsArgs {
    const float *loc, *vel, *charges;
    const unsigned char *adj;   // [S][5][5] multiplicity of edge (src j -> dst i)
    const float *prm;
    float *rot, *trans;         // (M, 3, 3), (M, 3)
    int S, cin, L, nonlin, pool_mean, final_mean, canon_translation;
    int feat_v, feat_a, feat_c; // which channels follow the canonical location: velocity, angular, charge-weighted location
};

__global__ void ds_adjacency_kernel(const long long *__restrict__ edges, long long E, int S, unsigned char *__restrict__ adj,
                                    int *__restrict__ bad) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
        const long long src = edges[e], dst = edges[E + e];
        const long long s1 = src / DS_P, s2 = dst / DS_P;
        if (src < 0 || dst < 0 || s1 != s2 || s1 >= S) {
            if (bad) atomicAdd(bad, 1);
            continue;
        }
        const int i = (int)(dst - s2 * DS_P), j = (int)(src - s1 * DS_P);
        // byte-wise atomic increment through the containing 32-bit word
        const size_t idx = (size_t)s2 * 25 + i * 5 + j;
        unsigned int *word = reinterpret_cast<unsigned int *>(adj) + (idx >> 2);
        atomicAdd(word, 1u << (8 * (idx & 3)));
    }
}

template <int H>
__global__ void __launch_bounds__(DS_THREADS) vndeepsets_kernel(const DeepSetsArgs a) {
    extern __shared__ __align__(16) float dsm[];
    float *xa = dsm;                               // [DS_THREADS][3][H]
    float *xb = xa + DS_THREADS * 3 * H;           // [DS_THREADS][3][H]
    float *W = xb + DS_THREADS * 3 * H;            // parameters
    const int tid = threadIdx.x, ls = tid / DS_P, i = tid - ls * DS_P;
    const int sys = blockIdx.x * DS_SYS + ls;
    const bool live = sys < a.S;
    const int m = sys * DS_P + i;
    // parameter block: layer 0 {Wid (H x cin), bid (H), Wpool (H x cin), bpool (H), Wdir (H x H)}, layers 1.. with cin = H,
    // then Wout (4 x H), bout (4)
    const int l0 = 2 * H * a.cin + 2 * H + H * H, ll = 3 * H * H + 2 * H;
    const int nprm = l0 + (a.L - 1) * ll + 4 * H + 4;
    for (int q = tid; q < nprm; q += DS_THREADS) W[q] = a.prm[q];

    // ---- features (:129-158) ------------------------------------------------------------------------------------
    float lc[3] = {0.f, 0.f, 0.f}, vl[3] = {0.f, 0.f, 0.f}, ch = 0.f;
    if (live) {
#pragma unroll
        for (int d = 0; d < 3; ++d) { lc[d] = a.loc[(size_t)m * 3 + d]; vl[d] = a.vel ? a.vel[(size_t)m * 3 + d] : 0.f; }
        ch = a.charges ? a.charges[m] : 0.f;
    }
    float *tmp = xb;   // [DS_THREADS][3] staging of loc for the per-system reduction
#pragma unroll
    for (int d = 0; d < 3; ++d) tmp[tid * 3 + d] = lc[d];
    __syncthreads();
    float mean_loc[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float s = 0.f;
        for (int j = 0; j < DS_P; ++j) s += tmp[(ls * DS_P + j) * 3 + d];
        mean_loc[d] = a.pool_mean ? s / (float)DS_P : s;
    }
    __syncthreads();
    {
        const float cl[3] = {lc[0] - mean_loc[0], lc[1] - mean_loc[1], lc[2] - mean_loc[2]};
        const float ang[3] = {cl[1] * vl[2] - cl[2] * vl[1], cl[2] * vl[0] - cl[0] * vl[2], cl[0] * vl[1] - cl[1] * vl[0]};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float *f = xa + (tid * 3 + d) * H;
            int c = 0;
            f[c++] = cl[d];
            if (a.feat_v) f[c++] = vl[d];
            if (a.feat_a) f[c++] = ang[d];
            if (a.feat_c) f[c++] = cl[d] * ch;
        }
    }
    // adjacency row of this particle: multiplicities of edges j -> i, and the in-degree
    float adj[DS_P], deg = 0.f;
#pragma unroll
    for (int j = 0; j < DS_P; ++j) {
        adj[j] = live ? (float)a.adj[(size_t)sys * 25 + i * 5 + j] : 0.f;
        deg += adj[j];
    }
    const float inv_deg = a.pool_mean ? (deg > 0.f ? 1.f / deg : 0.f) : 1.f;
    __syncthreads();

    // ---- layers ---------------------------------------------------------------------------------------------------
    float *xin = xa, *xout = xb;
    int cin = a.cin;
    const float *Wl = W;
    for (int l = 0; l < a.L; ++l) {
        const float *Wid = Wl, *bid = Wid + H * cin, *Wpool = bid + H, *bpool = Wpool + H * cin, *Wdir = bpool + H;
        float y[3][H];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float *xi = xin + (tid * 3 + d) * H;
            for (int o = 0; o < H; ++o) y[d][o] = bid[o] + bpool[o];
            for (int c = 0; c < cin; ++c) {
                float pooled = 0.f;
#pragma unroll
                for (int j = 0; j < DS_P; ++j) pooled = fmaf(adj[j], xin[((ls * DS_P + j) * 3 + d) * H + c], pooled);
                pooled *= inv_deg;
                const float xv = xi[c];
#pragma unroll
                for (int o = 0; o < H; ++o) y[d][o] = fmaf(Wpool[o * cin + c], pooled, fmaf(Wid[o * cin + c], xv, y[d][o]));
            }
        }
        // VN nonlinearity: direction d = map_to_dir(y), per channel o
        const float ns = a.nonlin == DS_LEAKY ? 0.2f : 0.f;
        for (int o = 0; o < H; ++o) {
            float dv[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < H; ++c) {
                const float w = Wdir[o * H + c];
                dv[0] = fmaf(w, y[0][c], dv[0]); dv[1] = fmaf(w, y[1][c], dv[1]); dv[2] = fmaf(w, y[2][c], dv[2]);
            }
            // (y is read at channel c above and written at channel o below: results go to shared memory, y stays intact)
            const float p0 = y[0][o], p1 = y[1][o], p2 = y[2][o];
            const float dot = p0 * dv[0] + p1 * dv[1] + p2 * dv[2];
            const float dn2 = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2];
            float mask;
            if (a.nonlin == DS_SOFTPLUS) {
                const float pn = sqrtf(p0 * p0 + p1 * p1 + p2 * p2), dn = sqrtf(dn2);
                const float ang = acosf(dot / (pn * dn + VN_EPS));
                const float cs = cosf(0.5f * ang);
                mask = cs * cs;
            } else {
                mask = dot >= 0.f ? 1.f : 0.f;
            }
            const float f = dot / (dn2 + VN_EPS);
            float r[3] = {p0 - f * dv[0], p1 - f * dv[1], p2 - f * dv[2]};
            const float pv[3] = {p0, p1, p2};
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float v = ns * pv[d] + (1.f - ns) * (mask * pv[d] + (1.f - mask) * r[d]);
                if (l > 0) v += xin[(tid * 3 + d) * H + o];    // residual (layers after the first, :62-75)
                xout[(tid * 3 + d) * H + o] = v;
            }
        }
        __syncthreads();
        float *t = xin; xin = xout; xout = t;
        Wl += l == 0 ? l0 : ll;
        cin = H;
    }
    // ---- final pooling over the particles of the system, output layer (:159-172) ----------------------------------
    const float *Wout = Wl, *bout = Wout + 4 * H;
    if (live) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float o4[4] = {bout[0], bout[1], bout[2], bout[3]};
            for (int c = 0; c < H; ++c) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < DS_P; ++j) s += xin[((ls * DS_P + j) * 3 + d) * H + c];
                if (a.final_mean) s /= (float)DS_P;
#pragma unroll
                for (int o = 0; o < 4; ++o) o4[o] = fmaf(Wout[o * H + c], s, o4[o]);
            }
            a.rot[(size_t)m * 9 + d * 3 + 0] = o4[0];
            a.rot[(size_t)m * 9 + d * 3 + 1] = o4[1];
            a.rot[(size_t)m * 9 + d * 3 + 2] = o4[2];
            a.trans[(size_t)m * 3 + d] = (a.canon_translation ? o4[3] : 0.f) + mean_loc[d];
        }
    }
}

}  // namespace eqb

using namespace eqb;

extern "C" int eqb_vndeepsets_param_count(int in_dim, int hidden, int num_layers) {
    return 2 * hidden * in_dim + 2 * hidden + hidden * hidden + (num_layers - 1) * (3 * hidden * hidden + 2 * hidden) +
           4 * hidden + 4;
}

extern "C" int64_t eqb_vndeepsets_workspace_bytes(int S) { return (int64_t)S * 25 + 64; }

extern "C" int eqb_vndeepsets_forward(const float *loc, const float *vel, const float *charges, const int64_t *edges,
                                      int64_t E, int S, const float *params, int in_dim, int hidden, int num_layers,
                                      int feat_v, int feat_a, int feat_c, int nonlinearity, int layer_pool_mean,
                                      int final_pool_mean, int canon_translation, float *rot_vectors, float *translation,
                                      void *workspace, int64_t workspace_bytes, int32_t *bad_edges, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(S >= 0 && E >= 0 && in_dim >= 1 && in_dim <= 4 && num_layers >= 1, "eqb_vndeepsets_forward: bad argument");
    EQB_REQUIRE(in_dim == 1 + (feat_v != 0) + (feat_a != 0) + (feat_c != 0), "eqb_vndeepsets_forward: in_dim does not match the feature flags");
    EQB_REQUIRE(nonlinearity >= DS_RELU && nonlinearity <= DS_SOFTPLUS, "eqb_vndeepsets_forward: unknown nonlinearity");
    EQB_UNSUPPORTED(hidden != 8 && hidden != 16 && hidden != 32, "eqb_vndeepsets_forward: hidden_dim %d not in {8, 16, 32}", hidden);
    if (S == 0) return 0;
    EQB_REQUIRE(loc && params && rot_vectors && translation && workspace && (E == 0 || edges), "eqb_vndeepsets_forward: null pointer");
    EQB_REQUIRE((feat_v == 0 && feat_a == 0) || vel, "eqb_vndeepsets_forward: velocity features need vel");
    EQB_REQUIRE(feat_c == 0 || charges, "eqb_vndeepsets_forward: charge features need charges");
    EQB_REQUIRE(workspace_bytes >= eqb_vndeepsets_workspace_bytes(S), "eqb_vndeepsets_forward: workspace too small");
    EQB_REQUIRE(((uintptr_t)workspace & 3) == 0, "eqb_vndeepsets_forward: workspace must be 4-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *adj = (unsigned char *)workspace;
    EQB_CUDA(cudaMemsetAsync(adj, 0, ((size_t)S * 25 + 3) & ~(size_t)3, st));
    if (bad_edges) EQB_CUDA(cudaMemsetAsync(bad_edges, 0, sizeof(int32_t), st));
    if (E > 0) {
        long long blocks = (E + 255) / 256;
        if (blocks > 1024) blocks = 1024;
        ds_adjacency_kernel<<<(unsigned)blocks, 256, 0, st>>>((const long long *)edges, (long long)E, S, adj, bad_edges);
    }
    DeepSetsArgs a{};
    a.loc = loc; a.vel = vel; a.charges = charges; a.adj = adj; a.prm = params; a.rot = rot_vectors; a.trans = translation;
    a.S = S; a.cin = in_dim; a.L = num_layers; a.nonlin = nonlinearity; a.pool_mean = layer_pool_mean != 0;
    a.final_mean = final_pool_mean != 0; a.canon_translation = canon_translation != 0;
    a.feat_v = feat_v != 0; a.feat_a = feat_a != 0; a.feat_c = feat_c != 0;
    const int nprm = eqb_vndeepsets_param_count(in_dim, hidden, num_layers);
    const size_t smem = ((size_t)2 * DS_THREADS * 3 * hidden + nprm) * sizeof(float);
    EQB_UNSUPPORTED(smem > 200 * 1024, "eqb_vndeepsets_forward: %d layers of width %d do not fit in shared memory", num_layers, hidden);
    const unsigned grid = (unsigned)((S + DS_SYS - 1) / DS_SYS);
    switch (hidden) {
        case 8:
            EQB_CUDA(cudaFuncSetAttribute(vndeepsets_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            vndeepsets_kernel<8><<<grid, DS_THREADS, smem, st>>>(a);
            break;
        case 16:
            EQB_CUDA(cudaFuncSetAttribute(vndeepsets_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            vndeepsets_kernel<16><<<grid, DS_THREADS, smem, st>>>(a);
            break;
        default:
            EQB_CUDA(cudaFuncSetAttribute(vndeepsets_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            vndeepsets_kernel<32><<<grid, DS_THREADS, smem, st>>>(a);
            break;
    }
    return finish_launch("vndeepsets_kernel");
}

extern "C" int eqb_vnsmall_param_count(void) { return VP_TOTAL; }

static int vnsmall_splits(int B, int N) {
    // One thread walks one point at a time, so a split pays only while a thread still owns more than one point and the
    // batch leaves SMs idle: at most ceil(N / 512) splits, and only as many as there are spare SMs per cloud
    // (measured: B = 128 is fastest un-split at 1.54 ms; B = 16 goes from 1.5 ms to 0.85 ms with a split).
    if (B <= 0) return 1;
    const int spare = num_sms() / B, by_points = (N + VN_MAX_THREADS - 1) / VN_MAX_THREADS;
    const int s = spare < by_points ? spare : by_points;
    return s < 1 ? 1 : s;
}

extern "C" int64_t eqb_vnsmall_workspace_bytes(int B, int N) {
    return (int64_t)(B > 0 ? B : 1) * vnsmall_splits(B, N) * 9 * (int64_t)sizeof(double);
}

extern "C" int eqb_vnsmall_forward(const float *x, int B, int N, const float *params, int n_knn, float bn_eps,
                                   float *out, void *workspace, int64_t workspace_bytes, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && N > 0 && n_knn > 0, "eqb_vnsmall_forward: bad shape");
    EQB_REQUIRE(n_knn <= N, "eqb_vnsmall_forward: n_knn = %d exceeds the %d points of a cloud", n_knn, N);
    EQB_UNSUPPORTED(n_knn > 32, "eqb_vnsmall_forward: n_knn = %d > 32 not supported by this build", n_knn);
    if (B == 0) return 0;
    EQB_REQUIRE(x && params && out && workspace, "eqb_vnsmall_forward: null pointer");
    EQB_REQUIRE(workspace_bytes >= eqb_vnsmall_workspace_bytes(B, N) && ((uintptr_t)workspace & 7) == 0,
                "eqb_vnsmall_forward: workspace too small or misaligned");
    const int S = vnsmall_splits(B, N);
    double *part = (double *)workspace;
    // 512 threads (128 registers, a few hundred bytes of spills) hide latency better than 256 (255 registers) when
    // there is at most one cloud per SM; EQB_VN_THREADS=256 selects the other build
    const char *tv = getenv("EQB_VN_THREADS");
    const int threads = tv && atoi(tv) == 256 ? 256 : 512;
    const size_t smem = ((size_t)8 * N + VP_TOTAL + (size_t)((n_knn == 20 ? 20 : 32) + 2 * VN_CAP) * threads) * sizeof(float);
    EQB_UNSUPPORTED(smem > 200 * 1024, "eqb_vnsmall_forward: clouds of %d points do not fit in shared memory", N);
    cudaStream_t st = (cudaStream_t)stream;
#define EQB_VN_LAUNCH(KK, TT)                                                                                          \
    do {                                                                                                               \
        EQB_CUDA(cudaFuncSetAttribute(vnsmall_kernel<KK, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
        vnsmall_kernel<KK, TT><<<dim3(B, S), TT, smem, st>>>(x, params, part, N, n_knn, bn_eps);                         \
    } while (0)
    if (n_knn == 20) {
        if (threads == 256) EQB_VN_LAUNCH(20, 256); else EQB_VN_LAUNCH(20, 512);
    } else {
        if (threads == 256) EQB_VN_LAUNCH(32, 256); else EQB_VN_LAUNCH(32, 512);
    }
#undef EQB_VN_LAUNCH
    vnsmall_finish_kernel<<<(B * 9 + 127) / 128, 128, 0, st>>>(part, B, S, N, out);
    return finish_launch("vnsmall_kernel");
}
