// N3: training kernels of the group-conv canonicalization network (CustomEquivariantNetwork).
//
// The reference trains this network through torch autograd over, per layer, the filter-orbit construction
// (custom_group_equivariant_layers.py:62-90, :169-199, :298-334, :461-507), F.conv2d (:104-112, :352-361), ReLU and the
// final mean (custom_equivariant_networks.py:80-93).  The fused inference stack (gconv_stack*.cu) keeps no
// activations, so training uses a layer-wise path that does: the forward conv saves every post-ReLU feature map, the
// backward needs, per layer,
//     dW_x[n, (ci,ky,kx)] = sum_{b,p} dY[b,n,p] X[b,ci,p+(ky,kx)]              (conv2d_weight_grad_kernel)
//     dX[b,c,p]           = [X[b,c,p] > 0] sum_n W_x[n,c] dY[b,n,p]            (conv2d_forward_kernel on W_x^T, 1x1, masked)
//     db_x[n]             = sum_{b,p} dY[b,n,p]                                (plane_sums_kernel, then a (B,N) sum)
// and the adjoints of the two (linear) filter-orbit maps (orbit_adjoint kernels).  These are fp32 SIMT kernels with
// 128 x 128 tiles and 8 x 8 register blocks: correctness and completeness of the training step first; the tensor-core treatment
// the inference stack got is the obvious successor (weight gradients are NT GEMMs with K = B*H*W).
#include <stdlib.h>

#include "common.cuh"
#include "gconv_stack_tc.cuh"

namespace eqb {

constexpr int GT_T = 128;   // tile side (outputs): 8 x 8 register block per thread
constexpr int GT_K = 16;    // reduction chunk
constexpr int GT_THREADS = 256;
constexpr int GT_PITCH = GT_T + 4;

// 8 x 8 outer-product update from one reduction row of the two shared-memory tiles: the thread owns rows
// {4 ty .. +3, 64 + 4 ty .. +3} and columns {4 tx .. +3, 64 + 4 tx .. +3} (two float4 reads per operand, conflict-free)
__device__ __forceinline__ void gt_fma_row(const float *__restrict__ arow, const float *__restrict__ brow, int ty, int tx,
                                           float (&acc)[8][8]) {
    const float4 a0 = *reinterpret_cast<const float4 *>(arow + 4 * ty), a1 = *reinterpret_cast<const float4 *>(arow + 64 + 4 * ty);
    const float4 b0 = *reinterpret_cast<const float4 *>(brow + 4 * tx), b1 = *reinterpret_cast<const float4 *>(brow + 64 + 4 * tx);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
}
// 4-byte cp.async with zero fill (src_bytes = 0 leaves the destination zeroed): the next chunk streams into the other
// shared-memory stage while the FMAs of the current one run
__device__ __forceinline__ void gt_cp4(float *smem_dst, const float *gmem_src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "r"(bytes));
}
__device__ __forceinline__ void gt_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void gt_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }
__device__ __forceinline__ int gt_index(int t, int i) { return (i < 4 ? 0 : 60) + 4 * t + i; }   // i-th owned row / column

// ---------------------------------------------------------------------------------------------------------------------
// Y[b,n,p] = act( bias[n] + sum_kk W[n,kk] * X[b, ci(kk), y(p)+ky(kk), x(p)+kx(kk)] ) [* (mask[b,n,p] > 0)]
// valid k x k convolution, NCHW fp32.  grid (ceil(P/128), ceil(N/128), B).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GT_THREADS) conv2d_forward_kernel(const float *__restrict__ X, const float *__restrict__ Wn,
                                                                    const float *__restrict__ bias, const float *__restrict__ mask,
                                                                    float *__restrict__ Y, int Cin, int H, int W, int k, int N,
                                                                    int relu) {
    __shared__ __align__(16) float ws[2][GT_K][GT_PITCH];   // [stage][kk][n]
    __shared__ __align__(16) float xs[2][GT_K][GT_PITCH];   // [stage][kk][p]
    const int Ho = H - k + 1, Wo = W - k + 1, P = Ho * Wo, K = Cin * k * k, kk2 = k * k;
    const int b = blockIdx.z, n0 = blockIdx.y * GT_T, p0 = blockIdx.x * GT_T;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const float *xb = X + (size_t)b * Cin * H * W;
    float acc[8][8] = {};
    // loaders.  Weights: lanes run along the 16 reduction entries of the chunk (contiguous in W[n][.]), thread owns tile
    // rows wn + 16 j; a lane-per-row layout reads 32 sectors per warp load and competes with the FMAs for the LSU.
    // Image: column lc = tid & 127 of the tile (contiguous pixels), reduction rows lr + 2 j.
    const int wk = tid & 15, wn = tid >> 4;
    const int lc = tid & 127, lr = tid >> 7;
    const int lp = min(p0 + lc, P - 1);
    const int pix = (lp / Wo) * W + (lp % Wo);
    auto issue = [&](int stage, int k0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + wn + 16 * j;
            const bool wok = k0 + wk < K && n < N;
            gt_cp4(&ws[stage][wk][wn + 16 * j], wok ? Wn + (size_t)n * K + k0 + wk : Wn, wok);
            const int r = lr + 2 * j, kk = k0 + r;
            const bool xok = kk < K;
            const int kc = xok ? kk : 0;
            const int ci = kc / kk2, rem = kc - ci * kk2, ky = rem / k, kx = rem - ky * k;
            gt_cp4(&xs[stage][r][lc], xb + (size_t)ci * H * W + ky * W + kx + pix, xok);
        }
        gt_commit();
    };
    issue(0, 0);
    int stage = 0;
    for (int k0 = 0; k0 < K; k0 += GT_K, stage ^= 1) {
        if (k0 + GT_K < K) { issue(stage ^ 1, k0 + GT_K); gt_wait<1>(); } else { gt_wait<0>(); }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < GT_K; ++q) gt_fma_row(ws[stage][q], xs[stage][q], ty, tx, acc);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = n0 + gt_index(ty, i);
        if (n >= N) continue;
        const float bv = bias ? bias[n] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = p0 + gt_index(tx, j);
            if (p >= P) continue;
            const size_t o = ((size_t)b * N + n) * P + p;
            float v = acc[i][j] + bv;
            if (relu) v = fmaxf(v, 0.f);
            if (mask) v = mask[o] > 0.f ? v : 0.f;
            Y[o] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// dW[n, kk] += sum over this CTA's (b, pixel) range of dY[b,n,p] * X[b, ci(kk), y(p)+ky, x(p)+kx].
// grid (ceil(K/128), ceil(N/128), splits); the reduction range B*P is cut into `splits` contiguous pieces of whole
// 16-pixel chunks; partial tiles are added atomically (dW is zeroed by the caller).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GT_THREADS) conv2d_weight_grad_kernel(const float *__restrict__ dY, const float *__restrict__ X,
                                                                        float *__restrict__ dW, int B, int Cin, int H, int W,
                                                                        int k, int N, int chunks_per_split) {
    __shared__ __align__(16) float gs[2][GT_K][GT_PITCH];   // [stage][p][n]
    __shared__ __align__(16) float xs[2][GT_K][GT_PITCH];   // [stage][p][kk]
    const int Ho = H - k + 1, Wo = W - k + 1, P = Ho * Wo, K = Cin * k * k, kk2 = k * k;
    const int kk0 = blockIdx.x * GT_T, n0 = blockIdx.y * GT_T;
    const int chunks_per_image = (P + GT_K - 1) / GT_K;
    const long long total_chunks = (long long)B * chunks_per_image;
    const long long c_begin = (long long)blockIdx.z * chunks_per_split;
    const long long c_end = min(total_chunks, c_begin + chunks_per_split);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    // loaders: lanes run along the 16 pixels of the chunk (contiguous in memory), thread owns columns lc + 16 j (j < 8)
    const int lp = tid & 15, lc = tid >> 4;
    int xoff[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int kk = kk0 + lc + 16 * j;
        xoff[j] = -1;
        if (kk < K) {
            const int ci = kk / kk2, r = kk - ci * kk2, ky = r / k, kx = r - ky * k;
            xoff[j] = ci * H * W + ky * W + kx;
        }
    }
    float acc[8][8] = {};
    auto issue = [&](int stage, long long c) {
        const int b = (int)(c / chunks_per_image), pc = (int)(c - (long long)b * chunks_per_image) * GT_K;
        const float *gyb = dY + ((size_t)b * N) * P;
        const float *xb = X + (size_t)b * Cin * H * W;
        const int p = pc + lp;
        const bool pok = p < P;
        const int pix = pok ? (p / Wo) * W + (p % Wo) : 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = lc + 16 * j, n = n0 + col;
            const bool gok = pok && n < N, xok = pok && xoff[j] >= 0;
            gt_cp4(&gs[stage][lp][col], gok ? gyb + (size_t)n * P + p : dY, gok);
            gt_cp4(&xs[stage][lp][col], xok ? xb + xoff[j] + pix : X, xok);
        }
        gt_commit();
    };
    if (c_begin < c_end) issue(0, c_begin);
    int stage = 0;
    for (long long c = c_begin; c < c_end; ++c, stage ^= 1) {
        if (c + 1 < c_end) { issue(stage ^ 1, c + 1); gt_wait<1>(); } else { gt_wait<0>(); }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < GT_K; ++q) gt_fma_row(gs[stage][q], xs[stage][q], ty, tx, acc);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = n0 + gt_index(ty, i);
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int kk = kk0 + gt_index(tx, j);
            if (kk < K) atomicAdd(dW + (size_t)n * K + kk, acc[i][j]);
        }
    }
}

// out[row] = sum_p x[row, p]   (fp64 accumulation, one CTA per row)
__global__ void __launch_bounds__(256) plane_sums_kernel(const float *__restrict__ x, long long P, float *__restrict__ out) {
    const float *xr = x + (size_t)blockIdx.x * P;
    double acc = 0.0;
    for (long long p = threadIdx.x; p < P; p += blockDim.x) acc += (double)__ldg(xr + p);
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        out[blockIdx.x] = (float)t;
    }
}

// dY[b, o*G + g, p] = scale * dact[b, g]     (backward of the mean over (o, p), custom_equivariant_networks.py:91)
__global__ void __launch_bounds__(256) group_mean_backward_kernel(const float *__restrict__ dact, float *__restrict__ dY,
                                                                  int N, int G, long long P, float scale) {
    const int row = blockIdx.x;            // b * N + n
    const int b = row / N, n = row - b * N;
    const float v = scale * dact[(size_t)b * G + n % G];
    float *d = dY + (size_t)row * P;
    for (long long p = threadIdx.x; p < P; p += blockDim.x) d[p] = v;
}

// ---- adjoints of the filter orbits (small_ops.cu: lift_orbit_kernel / regular_orbit_kernel) -----------------------
// orbit[n, kk] = rotated_tap(w_slice, ...) is a 4-tap bilinear gather; its adjoint scatters the same four weights.
__device__ __forceinline__ void rotated_tap_scatter(float *__restrict__ dw, float gv, int k, int y, int x, int r, int N, bool mirror) {
    if (mirror) x = k - 1 - x;
    double c, s;
    rot_cs(r, N, 1.0, c, s);
    const double ctr = 0.5 * (k - 1);
    const double u = x - ctr, v = y - ctr;
    const double xs = ctr + c * u - s * v, ys = ctr + s * u + c * v;
    const double xf = floor(xs), yf = floor(ys);
    const float fx = (float)(xs - xf), fy = (float)(ys - yf);
    const int x0 = (int)xf, y0 = (int)yf;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int xi = x0 + dx, yi = y0 + dy;
            if (xi >= 0 && xi < k && yi >= 0 && yi < k) {
                const float wt = (dy ? fy : 1.f - fy) * (dx ? fx : 1.f - fx);
                if (wt != 0.f) atomicAdd(dw + yi * k + xi, wt * gv);
            }
        }
}

__device__ __forceinline__ int regular_src_slice_t(int g, int h, int N) {   // same table as small_ops.cu
    if (g < N) return h < N ? (h - g + N) % N : N + (h - N + g) % N;
    const int gp = g - N;
    return h < N ? N + (h + gp) % N : (h - N - gp + N) % N;
}

__global__ void lift_orbit_adjoint_kernel(const float *__restrict__ dorbit, float *__restrict__ dw, int cout, int cin, int k,
                                          int N, int G) {
    const int kk2 = k * k, K = cin * kk2;
    const long long total = (long long)cout * G * K;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(t % K), n = (int)(t / K);
        const int o = n / G, g = n % G, i = kk / kk2, yx = kk % kk2;
        rotated_tap_scatter(dw + ((size_t)o * cin + i) * kk2, dorbit[t], k, yx / k, yx % k, g % N, N, g >= N);
    }
}

__global__ void regular_orbit_adjoint_kernel(const float *__restrict__ dorbit, float *__restrict__ dw, int cout, int cin, int k,
                                             int N, int G) {
    const int kk2 = k * k, K = cin * G * kk2;
    const long long total = (long long)cout * G * K;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(t % K), n = (int)(t / K);
        const int o = n / G, g = n % G, ih = kk / kk2, yx = kk % kk2;
        const int i = ih / G, h = ih % G;
        const int src = regular_src_slice_t(g, h, N);
        rotated_tap_scatter(dw + (((size_t)o * cin + i) * G + src) * kk2, dorbit[t], k, yx / k, yx % k, g % N, N, g >= N);
    }
}

}  // namespace eqb

using namespace eqb;

// x_absmax (in) / y_absmax (out): per-image max |.| (B floats each, may be NULL).  The tensor-core kernels scale their fp16
// operand split per image; a caller that chains layers hands the maxima on instead of paying a pass over the feature map.
extern "C" int eqb_conv2d_forward_scaled(const float *x, const float *w, const float *bias, const float *mask, float *y, int B,
                                         int cin, int H, int W, int N, int k, int relu, const float *x_absmax, float *y_absmax,
                                         void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && cin > 0 && N > 0 && k > 0 && H >= k && W >= k, "eqb_conv2d_forward: bad shape");
    EQB_REQUIRE(B == 0 || (x && w && y), "eqb_conv2d_forward: null pointer");
    EQB_REQUIRE(B <= 65535 && (N + GT_T - 1) / GT_T <= 65535, "eqb_conv2d_forward: grid too large");
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int P = (H - k + 1) * (W - k + 1);
    {
        // 256 output channels, cin * k * k <= 256 (the lift and the 1x1 layers of the flagship network): CTA-pair tcgen05
        // kernel with the fp16 hi/lo operand split (gconv_stack_tc.cu, namespace pw)
        int handled = 0;
        if (int err = tc_pw_conv(x, w, bias, mask, y, B, cin, H, W, N, k, relu, x_absmax, y_absmax, st, &handled)) return err;
        if (handled) return 0;
    }
    dim3 grid((P + GT_T - 1) / GT_T, (N + GT_T - 1) / GT_T, B);
    conv2d_forward_kernel<<<grid, GT_THREADS, 0, st>>>(x, w, bias, mask, y, cin, H, W, k, N, relu);
    if (int err = finish_launch("eqb_conv2d_forward")) return err;
    if (y_absmax) return tc_absmax(y, B, (size_t)N * (size_t)P, y_absmax, st);
    return 0;
}

extern "C" int eqb_conv2d_forward(const float *x, const float *w, const float *bias, const float *mask, float *y, int B,
                                  int cin, int H, int W, int N, int k, int relu, void *stream) {
    return eqb_conv2d_forward_scaled(x, w, bias, mask, y, B, cin, H, W, N, k, relu, nullptr, nullptr, stream);
}

// out[n] = sum_{b, p} dy[b, n, p] (the SIMT path's version of the row sums the tensor-core kernel folds into its converters)
__global__ void __launch_bounds__(256) channel_sums_kernel(const float *__restrict__ dy, int B, int N, long long P,
                                                           float *__restrict__ out) {
    __shared__ float red[8];
    const int n = blockIdx.x;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
        const float *row = dy + ((size_t)b * N + n) * P;
        for (long long p = threadIdx.x; p < P; p += 256) acc += row[p];
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        out[n] = t;
    }
}

extern "C" int eqb_conv2d_weight_grad_scaled(const float *dy, const float *x, float *dw, int B, int cin, int H, int W, int N,
                                             int k, const float *dy_absmax, const float *x_absmax, float *dy_rowsum,
                                             void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && cin > 0 && N > 0 && k > 0 && H >= k && W >= k, "eqb_conv2d_weight_grad: bad shape");
    EQB_REQUIRE(dw && (B == 0 || (dy && x)), "eqb_conv2d_weight_grad: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int K = cin * k * k, P = (H - k + 1) * (W - k + 1);
    EQB_CUDA(cudaMemsetAsync(dw, 0, (size_t)N * K * sizeof(float), st));
    if (dy_rowsum) EQB_CUDA(cudaMemsetAsync(dy_rowsum, 0, (size_t)N * sizeof(float), st));
    if (B == 0) return 0;
    {
        int handled = 0;
        if (int err = tc_pw_wgrad(dy, x, dw, dy_rowsum, B, cin, H, W, N, k, dy_absmax, x_absmax, st, &handled)) return err;
        if (handled) return 0;
    }
    if (dy_rowsum) channel_sums_kernel<<<N, 256, 0, st>>>(dy, B, N, (long long)P, dy_rowsum);
    const int tiles = ((K + GT_T - 1) / GT_T) * ((N + GT_T - 1) / GT_T);
    const long long total_chunks = (long long)B * ((P + GT_K - 1) / GT_K);
    long long splits = (4LL * num_sms() + tiles - 1) / tiles;       // a few waves of CTAs
    splits = std::max(1LL, std::min(std::min(splits, total_chunks), 65535LL));
    const long long per = (total_chunks + splits - 1) / splits;
    splits = (total_chunks + per - 1) / per;
    EQB_REQUIRE(per < (1LL << 31), "eqb_conv2d_weight_grad: reduction too long");
    dim3 grid((K + GT_T - 1) / GT_T, (N + GT_T - 1) / GT_T, (unsigned)splits);
    conv2d_weight_grad_kernel<<<grid, GT_THREADS, 0, st>>>(dy, x, dw, B, cin, H, W, k, N, (int)per);
    return finish_launch("eqb_conv2d_weight_grad");
}

extern "C" int eqb_conv2d_weight_grad(const float *dy, const float *x, float *dw, int B, int cin, int H, int W, int N, int k,
                                      void *stream) {
    return eqb_conv2d_weight_grad_scaled(dy, x, dw, B, cin, H, W, N, k, nullptr, nullptr, nullptr, stream);
}

extern "C" int eqb_plane_sums(const float *x, int64_t rows, int64_t P, float *out, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(rows >= 0 && P > 0 && rows < (1LL << 31), "eqb_plane_sums: bad shape");
    EQB_REQUIRE(rows == 0 || (x && out), "eqb_plane_sums: null pointer");
    if (rows == 0) return 0;
    plane_sums_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, P, out);
    return finish_launch("eqb_plane_sums");
}

extern "C" int eqb_group_mean_backward(const float *dact, float *dy, int B, int cout, int num_group, int64_t P, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && cout > 0 && num_group > 0 && P > 0, "eqb_group_mean_backward: bad shape");
    EQB_REQUIRE(B == 0 || (dact && dy), "eqb_group_mean_backward: null pointer");
    const long long rows = (long long)B * cout * num_group;
    EQB_REQUIRE(rows < (1LL << 31), "eqb_group_mean_backward: too many planes");
    if (rows == 0) return 0;
    group_mean_backward_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(dact, dy, cout * num_group, num_group, P,
                                                                                 (float)(1.0 / ((double)cout * (double)P)));
    return finish_launch("eqb_group_mean_backward");
}

extern "C" int eqb_lift_filter_orbit_adjoint(const float *dorbit, float *dw, int cout, int cin, int k, int num_rotations,
                                             int reflect, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(cout > 0 && cin > 0 && k > 0 && num_rotations > 0 && dorbit && dw, "eqb_lift_filter_orbit_adjoint: bad argument");
    const int G = num_rotations * (reflect ? 2 : 1);
    cudaStream_t st = (cudaStream_t)stream;
    EQB_CUDA(cudaMemsetAsync(dw, 0, (size_t)cout * cin * k * k * sizeof(float), st));
    const long long total = (long long)cout * G * cin * k * k;
    lift_orbit_adjoint_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 4096), 256, 0, st>>>(dorbit, dw, cout, cin, k,
                                                                                                      num_rotations, G);
    return finish_launch("eqb_lift_filter_orbit_adjoint");
}

extern "C" int eqb_regular_filter_orbit_adjoint(const float *dorbit, float *dw, int cout, int cin, int k, int num_rotations,
                                                int reflect, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(cout > 0 && cin > 0 && k > 0 && num_rotations > 0 && dorbit && dw, "eqb_regular_filter_orbit_adjoint: bad argument");
    const int G = num_rotations * (reflect ? 2 : 1);
    cudaStream_t st = (cudaStream_t)stream;
    EQB_CUDA(cudaMemsetAsync(dw, 0, (size_t)cout * cin * G * k * k * sizeof(float), st));
    const long long total = (long long)cout * G * cin * G * k * k;
    regular_orbit_adjoint_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 4096), 256, 0, st>>>(dorbit, dw, cout, cin, k,
                                                                                                         num_rotations, G);
    return finish_launch("eqb_regular_filter_orbit_adjoint");
}
