// a7: e2cnn-style group-equivariant conv stack with EXPANDED filters -> group activations.
//
// Reference: ESCNNEquivariantNetwork (escnn_networks.py:59-117).  In eval() every e2cnn module of that network is a
// plain dense op on the expanded tensors e2cnn caches (R2Conv.filter / .expanded_bias; InnerBatchNorm = one affine
// per field shared by its |G| channels; ReLU; PointwiseDropout = identity):
//   x -> [ conv2d(k x k, valid) + bias -> scale * . + shift -> ReLU ] x (L-1) -> conv2d(k x k) + bias
//     -> reshape (B, Cout, |G|, H', W') -> mean over (Cout, H', W')                      => (B, |G|)
// Unlike CustomEquivariantNetwork (gconv_stack*.cu) every layer is k x k, so layers cannot be chained per pixel
// tile without halos: each layer is one implicit-GEMM pass (M = output pixels, N = Cout*|G| <= 256, K = Cin*k*k)
// whose epilogue applies bias / affine / ReLU and writes the next layer's NCHW input; the LAST layer's epilogue
// does not write its output at all - it accumulates the masked spatial sums that the group pool needs.
//
// This is the first correct CUDA path for a7: fp32 SIMT, 64-pixel x Npad tiles, 8 x NT register blocks, K streamed
// in chunks of 16 (weights by cp.async, the im2col chunk gathered one chunk ahead into registers so the global
// latency hides behind the FMAs).  A tcgen05 variant along the lines of gconv_stack_tc.cu is the planned successor.
#include <stdlib.h>

#include "common.cuh"
#include "conv_stack_tc.cuh"
#include "gconv_stack_tc.cuh"

namespace eqb {

constexpr int CV_TM = 64;      // pixels per tile
constexpr int CV_PITCH = 68;   // floats per im2col row
constexpr int CV_KC = 16;      // K rows per chunk
constexpr int CV_THREADS = 256;

struct ConvArgs {
    const float *x;        // (B, cin, H, W)
    float *y;              // (B, N, Ho, Wo) or nullptr for the pooled last layer
    const float *Wt;       // [Kpad][Npad] K-major
    const float *bias;     // [Npad]
    const float *scale;    // [Npad] (1 when no affine)
    const float *shift;    // [Npad]
    double *S_part;        // [B][chunks][Npad] (pooled layer)
    __half *y_hi, *y_lo;   // (B, Ho, Wo, Npad) fp16 hi / lo pair of s_out * value (input of a tcgen05 layer), or null
    const float *lay;      // device layer record (LAY_SOUT) when y_hi is set
    int B, cin, H, W, ksz, Ho, Wo, P, K, Kpad, N, Npad, relu;
    int tiles, chunks, tiles_per_chunk;
};

__device__ __forceinline__ void cv_cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src));
}

template <int NT>
__global__ void __launch_bounds__(CV_THREADS, 2) conv_layer_kernel(const ConvArgs a) {
    extern __shared__ __align__(16) float sm[];
    float *abuf = sm;                                   // [2][CV_KC][CV_PITCH]  im2col chunk
    float *wbuf = abuf + 2 * CV_KC * CV_PITCH;          // [2][CV_KC][Npad]
    int *koff = reinterpret_cast<int *>(wbuf + 2 * CV_KC * a.Npad);   // [Kpad] offset of tap k relative to the patch origin
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int b = blockIdx.x / a.chunks, ch = blockIdx.x - b * a.chunks;
    const int tile_begin = ch * a.tiles_per_chunk, tile_end = min(a.tiles, tile_begin + a.tiles_per_chunk);
    const int chunk_f4 = CV_KC * a.Npad / 4;
    const int kk2 = a.ksz * a.ksz;
    const float *xb = a.x + (size_t)b * a.cin * a.H * a.W;

    for (int k = tid; k < a.Kpad; k += CV_THREADS) {
        int off = -1;
        if (k < a.K) {
            const int c = k / kk2, rem = k - c * kk2, ky = rem / a.ksz, kx = rem - ky * a.ksz;
            off = (c * a.H + ky) * a.W + kx;
        }
        koff[k] = off;
    }
    __syncthreads();

    // im2col gather assignment: thread -> pixel p = tid & 63, K rows kq, kq + 4, kq + 8, kq + 12 of the chunk
    const int gp = tid & (CV_TM - 1), gq = tid >> 6;
    const int nchunks = a.Kpad / CV_KC;

    double colsum[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) colsum[j] = 0.0;

    for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int p0 = tile * CV_TM;
        const int pix = p0 + gp;
        const bool ok = pix < a.P;
        const int oy = ok ? pix / a.Wo : 0, ox = ok ? pix - oy * a.Wo : 0;
        const float *xp = xb + (size_t)oy * a.W + ox;

        float acc[8][NT];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;

        float g[4];
        auto gather = [&](int kc) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int off = koff[kc * CV_KC + gq + 4 * i];
                g[i] = (ok && off >= 0) ? __ldg(xp + off) : 0.f;
            }
        };
        auto stash = [&](int buf) {
#pragma unroll
            for (int i = 0; i < 4; ++i) abuf[(buf * CV_KC + gq + 4 * i) * CV_PITCH + gp] = g[i];
        };
        // prologue: chunk 0
        for (int i = tid; i < chunk_f4; i += CV_THREADS) cv_cp_async16(wbuf + 4 * i, a.Wt + 4 * i);
        asm volatile("cp.async.commit_group;");
        gather(0);
        stash(0);
        for (int kc = 0; kc < nchunks; ++kc) {
            const int cur = kc & 1, nxt = cur ^ 1;
            if (kc + 1 < nchunks) {
                const float *src = a.Wt + (size_t)(kc + 1) * CV_KC * a.Npad;
                float *wn = wbuf + nxt * (CV_KC * a.Npad);
                for (int i = tid; i < chunk_f4; i += CV_THREADS) cv_cp_async16(wn + 4 * i, src + 4 * i);
                asm volatile("cp.async.commit_group;");
                gather(kc + 1);                       // global loads in flight during the FMAs below
                asm volatile("cp.async.wait_group 1;");
            } else {
                asm volatile("cp.async.wait_group 0;");
            }
            __syncthreads();                          // chunk kc: weights landed, im2col rows stashed
            const float *ap = abuf + cur * (CV_KC * CV_PITCH) + ty * 8;
            const float *wb = wbuf + cur * (CV_KC * a.Npad);
#pragma unroll
            for (int kk = 0; kk < CV_KC; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4 *>(ap + kk * CV_PITCH);
                const float4 a1 = *reinterpret_cast<const float4 *>(ap + kk * CV_PITCH + 4);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                float wv[NT];
#pragma unroll
                for (int j = 0; j < NT; ++j) wv[j] = wb[kk * a.Npad + j * 32 + tx];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
            }
            if (kc + 1 < nchunks) stash(nxt);         // buffer nxt was last read in iteration kc-1 (barrier above)
            __syncthreads();                          // reads of buffers cur done before they are refilled
        }

        // ---- epilogue -----------------------------------------------------------------------------------
        const float s_out = a.y_hi ? a.lay[LAY_SOUT] : 1.f;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int n = j * 32 + tx;
            const float bv = a.bias[n], sc = a.scale[n], sh = a.shift[n];
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v = fmaf(acc[i][j] + bv, sc, sh);
                if (a.relu) v = fmaxf(v, 0.f);
                const int p = p0 + ty * 8 + i;
                if (p < a.P) {
                    if (a.y_hi) {   // NHWC: the 32 lanes of a warp write 32 consecutive channels of one pixel
                        const float vs = v * s_out;
                        const __half hi = __float2half_rn(vs);
                        const size_t o = ((size_t)b * a.P + p) * a.Npad + n;
                        a.y_hi[o] = hi;
                        a.y_lo[o] = __float2half_rn(vs - __half2float(hi));
                    } else if (a.y) {
                        if (n < a.N) a.y[((size_t)b * a.N + n) * a.P + p] = v;
                    } else {
                        s += v;
                    }
                }
            }
            colsum[j] += (double)s;
        }
    }

    if (!a.y && !a.y_hi) {
        __syncthreads();
        double *red = reinterpret_cast<double *>(sm);  // [8][Npad] doubles <= the two weight buffers
#pragma unroll
        for (int j = 0; j < NT; ++j) red[ty * a.Npad + j * 32 + tx] = colsum[j];
        __syncthreads();
        for (int n = tid; n < a.Npad; n += CV_THREADS) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w * a.Npad + n];
            a.S_part[((size_t)b * a.chunks + ch) * a.Npad + n] = s;
        }
    }
}

// filter (N, Cin, k, k) -> K-major zero-padded Wt [Kpad][Npad]; per-channel vectors padded to Npad
__global__ void conv_pack_kernel(const float *__restrict__ w, float *__restrict__ Wt, int N, int K, int Kpad, int Npad) {
    const int total = Kpad * Npad;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int k = t / Npad, n = t - k * Npad;
        Wt[t] = (k < K && n < N) ? w[(size_t)n * K + k] : 0.f;
    }
}
__global__ void conv_vec_kernel(const float *__restrict__ bias, const float *__restrict__ scale,
                                const float *__restrict__ shift, float *__restrict__ out, int N, int Npad) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < Npad; n += gridDim.x * blockDim.x) {
        out[n] = (bias && n < N) ? bias[n] : 0.f;
        out[Npad + n] = (scale && n < N) ? scale[n] : 1.f;
        out[2 * Npad + n] = (shift && n < N) ? shift[n] : 0.f;
    }
}

// act[b][g] = sum_{o} S[b][o*G+g] / (cout * P)
__global__ void conv_pool_finish_kernel(const double *__restrict__ S_part, int chunks, int Npad, int cout, int G,
                                        double inv_count, float *__restrict__ act) {
    const int b = blockIdx.x, g = threadIdx.x;
    if (g >= G) return;
    double s = 0.0;
    for (int o = 0; o < cout; ++o)
        for (int c = 0; c < chunks; ++c) s += S_part[((size_t)b * chunks + c) * Npad + o * G + g];
    act[(size_t)b * G + g] = (float)(s * inv_count);
}

// ---- the last layer is linear and followed only by the mean: fold it through the pool ------------------------------
//   mean_{o,y,x} (W * in + b)[(o,g)] = sum_{c,ky,kx} Wfold[g][c][ky][kx] * S[c][ky][kx] / (Cout * P) + mean_o b[(o,g)]
//   S[c][ky][kx]   = sum of in[c] over the window [ky, ky + Ho) x [kx, kx + Wo)      (k x k shifted box sums)
//   Wfold[g][...]  = sum_o W[(o,g)][c][ky][kx]
// so the last conv (half of the FLOPs of the reference's default 3-layer network) is never executed.
// One block per (image, input channel), one thread per COLUMN: thread x streams in[0..H)[x] (row-contiguous across the
// block -> coalesced, all loads independent), keeps the column total in fp64 and turns it into the k vertical window
// sums by removing the leading / trailing k-1 elements (re-read, L1 hits); k*k threads then add the columns of their
// horizontal window.  HBM-bound: one read of the plane.
__global__ void __launch_bounds__(256) conv_window_sums_kernel(const float *__restrict__ in, int C, int H, int W, int k,
                                                               double *__restrict__ S) {
    extern __shared__ __align__(16) unsigned char wsm[];
    double *cols = reinterpret_cast<double *>(wsm);     // [k][W]  vertical window sums per column
    const int bc = blockIdx.x, Ho = H - k + 1, Wo = W - k + 1;
    const float *src = in + (size_t)bc * H * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int y = 0;
        for (; y + 3 < H; y += 4) {
            s0 += (double)__ldg(src + (size_t)y * W + x);
            s1 += (double)__ldg(src + (size_t)(y + 1) * W + x);
            s2 += (double)__ldg(src + (size_t)(y + 2) * W + x);
            s3 += (double)__ldg(src + (size_t)(y + 3) * W + x);
        }
        for (; y < H; ++y) s0 += (double)__ldg(src + (size_t)y * W + x);
        const double total = (s0 + s1) + (s2 + s3);
        double tail = 0.0, head = 0.0;
        for (int yy = Ho; yy < H; ++yy) tail += (double)__ldg(src + (size_t)yy * W + x);
        for (int ky = 0; ky < k; ++ky) {
            cols[ky * W + x] = total - head - tail;      // rows [ky, ky + Ho)
            if (ky + 1 < k) {
                head += (double)__ldg(src + (size_t)ky * W + x);
                tail -= (double)__ldg(src + (size_t)(ky + Ho) * W + x);
            }
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < k * k; t += blockDim.x) {
        const int ky = t / k, kx = t - ky * k;
        double s = 0.0;
        for (int x = kx; x < kx + Wo; ++x) s += cols[ky * W + x];
        S[(size_t)bc * k * k + t] = s;
    }
}

// Wfold[g][kk] = sum_o w[(o*G+g)][kk], kk over (c, ky, kx);  bmean[g] = mean_o bias[o*G+g]
__global__ void conv_fold_weights_kernel(const float *__restrict__ w, const float *__restrict__ bias, int cout, int G, int K,
                                         double *__restrict__ Wfold, double *__restrict__ bmean) {
    const int total = G * K;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int g = t / K, kk = t - g * K;
        double s = 0.0;
        for (int o = 0; o < cout; ++o) s += (double)w[(size_t)(o * G + g) * K + kk];
        Wfold[t] = s;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < G) {
        double s = 0.0;
        if (bias)
            for (int o = 0; o < cout; ++o) s += (double)bias[o * G + threadIdx.x];
        bmean[threadIdx.x] = s / (double)cout;
    }
}

// act[b][g] = Wfold[g] . S[b] * inv_count + bmean[g]; one block per image
__global__ void __launch_bounds__(256) conv_fold_apply_kernel(const double *__restrict__ S, const double *__restrict__ Wfold,
                                                              const double *__restrict__ bmean, int K, int G, double inv_count,
                                                              float *__restrict__ act) {
    __shared__ double red[8];
    const int b = blockIdx.x;
    const double *Sb = S + (size_t)b * K;
    for (int g = 0; g < G; ++g) {
        double v = 0.0;
        for (int kk = threadIdx.x; kk < K; kk += blockDim.x) v += Sb[kk] * Wfold[(size_t)g * K + kk];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
            act[(size_t)b * G + g] = (float)(t * inv_count + bmean[g]);
        }
        __syncthreads();
    }
}

struct ConvPlan {
    int N, Npad, P[16], Ho[16], Wo[16], cin[16], K[16], Kpad[16];
    int tiles, chunks, tiles_per_chunk;   // of the pooled (last) layer
    size_t off_wt[16], off_vec[16], off_buf[2], off_S, off_fold, off_bmean, off_win, total;
    size_t smem[16];
    bool fold;             // last layer folded through the pool (its input plane fits the window-sum kernel's staging)
    size_t win_smem;
    // tcgen05 inner layers (conv_stack_tc.cu): fp16 hi/lo NHWC ping-pong buffers, packed weights, layer records
    bool tc;
    size_t off_h[2][2], off_wp[16], off_lay, off_absmax, off_rowstat, off_xin[2];
    int cpad0;             // input channels padded to a whole K atom: 16, or a multiple of 32 (layer 0 on the tensor path)
};

static int conv_make_plan(int B, int cin, int H, int W, int cout, int k, int G, int L, ConvPlan &p) {
    EQB_REQUIRE(B >= 0 && cin > 0 && H > 0 && W > 0 && cout > 0 && k > 0 && G > 0 && L >= 1, "eqb_conv_stack: bad argument");
    EQB_UNSUPPORTED(L > 16, "eqb_conv_stack: more than 16 layers not supported");
    p.N = cout * G;
    EQB_UNSUPPORTED(p.N > 256, "eqb_conv_stack: out_channels*|G| = %d > 256 not supported by this build", p.N);
    p.Npad = p.N <= 32 ? 32 : p.N <= 64 ? 64 : p.N <= 128 ? 128 : 256;
    int h = H, w = W, c = cin;
    size_t off = 0;
    for (int l = 0; l < L; ++l) {
        EQB_REQUIRE(h >= k && w >= k, "eqb_conv_stack: feature map %dx%d of layer %d smaller than the %dx%d kernel", h, w, l, k, k);
        p.cin[l] = c;
        p.K[l] = c * k * k;
        p.Kpad[l] = (p.K[l] + CV_KC - 1) / CV_KC * CV_KC;
        h -= k - 1; w -= k - 1;
        p.Ho[l] = h; p.Wo[l] = w; p.P[l] = h * w;
        p.off_wt[l] = off; off += (size_t)p.Kpad[l] * p.Npad * sizeof(float);
        p.off_vec[l] = off; off += (size_t)3 * p.Npad * sizeof(float);
        p.smem[l] = ((size_t)2 * CV_KC * CV_PITCH + (size_t)2 * CV_KC * p.Npad) * sizeof(float) + (size_t)p.Kpad[l] * sizeof(int);
        EQB_UNSUPPORTED(p.smem[l] > 200 * 1024, "eqb_conv_stack: Cin*k*k = %d too large for the tap table", p.K[l]);
        c = p.N;
    }
    off = (off + 255) & ~(size_t)255;
    // ping-pong activation buffers (layer outputs shrink, the first is the largest)
    const size_t act_bytes = L > 1 ? (size_t)(B > 0 ? B : 1) * p.N * p.P[0] * sizeof(float) : 0;
    p.off_buf[0] = off; off += (act_bytes + 255) & ~(size_t)255;
    p.off_buf[1] = off; off += L > 2 ? ((act_bytes + 255) & ~(size_t)255) : 0;
    p.tiles = (p.P[L - 1] + CV_TM - 1) / CV_TM;
    const int target = num_sms() * 2 * 4;
    int chunks = B > 0 ? (target + B - 1) / B : 1;
    chunks = chunks < 1 ? 1 : chunks > p.tiles ? p.tiles : chunks;
    p.tiles_per_chunk = (p.tiles + chunks - 1) / chunks;
    p.chunks = (p.tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
    p.off_S = off; off += (size_t)(B > 0 ? B : 1) * p.chunks * p.Npad * sizeof(double);
    // fold of the last layer: input plane (H_in x W_in of layer L-1) staged in shared memory
    const int Hin = p.Ho[L - 1] + k - 1, Win = p.Wo[L - 1] + k - 1;
    p.win_smem = (size_t)Win * k * sizeof(double);
    (void)Hin;
    p.fold = p.win_smem <= 200 * 1024 && !getenv("EQB_CONV_NO_FOLD");
    off = (off + 255) & ~(size_t)255;
    p.off_fold = off; off += (size_t)G * p.K[L - 1] * sizeof(double);
    p.off_bmean = off; off += 64 * sizeof(double);
    p.off_win = off; off += (size_t)(B > 0 ? B : 1) * p.K[L - 1] * sizeof(double);
    p.tc = p.fold && ctc_eligible(p.Npad, L);
    for (int l = 1; l <= L - 2; ++l)   // every tcgen05 layer's input must hold at least one halo box (8+k-1) x (16+k-1)
        if (p.Wo[l - 1] < 8 + k - 1 || p.Ho[l - 1] < 16 + k - 1) p.tc = false;
    if (W < 8 + k - 1 || H < 16 + k - 1) p.tc = false;
    p.cpad0 = cin <= 16 ? 16 : (cin + 31) / 32 * 32;
    if (p.tc) {
        off = (off + 1023) & ~(size_t)1023;
        const size_t hbytes = (((size_t)(B > 0 ? B : 1) * p.P[0] * p.Npad * sizeof(__half)) + 1023) & ~(size_t)1023;
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) {
                p.off_h[i][j] = off;
                off += (i == 0 || L > 3) ? hbytes : 0;
            }
        for (int l = 1; l <= L - 2; ++l) {
            p.off_wp[l] = off;
            off += (ctc_pack_bytes(p.Npad, p.Npad, k) + 1023) & ~(size_t)1023;
        }
        p.off_wp[0] = off; off += (ctc_pack_bytes(p.Npad, p.cpad0, k) + 1023) & ~(size_t)1023;
        const size_t xbytes = (((size_t)(B > 0 ? B : 1) * H * W * p.cpad0 * sizeof(__half)) + 1023) & ~(size_t)1023;
        p.off_xin[0] = off; off += xbytes;
        p.off_xin[1] = off; off += xbytes;
        p.off_lay = off; off += (size_t)(L + 1) * (B > 0 ? B : 1) * LAY_FLOATS * sizeof(float);   // one record per layer AND image
        p.off_absmax = off; off += (((size_t)(B > 0 ? B : 1) * sizeof(float)) + 63) & ~(size_t)63;
        p.off_rowstat = off; off += (size_t)2 * p.Npad * sizeof(float);
    }
    p.total = off;
    return 0;
}

template <int NT>
static int conv_launch(const ConvArgs &a, size_t smem, cudaStream_t st) {
    EQB_CUDA(cudaFuncSetAttribute(conv_layer_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    EQB_REQUIRE((long long)a.B * a.chunks < (1LL << 31), "eqb_conv_stack: grid too large");
    conv_layer_kernel<NT><<<(unsigned)(a.B * a.chunks), CV_THREADS, smem, st>>>(a);
    return finish_launch("conv_layer_kernel");
}

}  // namespace eqb

using namespace eqb;

extern "C" int64_t eqb_conv_stack_workspace_bytes(int B, int cin, int H, int W, int cout, int k, int num_group,
                                                  int num_layers) {
    ConvPlan p;
    const int rc = conv_make_plan(B, cin, H, W, cout, k, num_group, num_layers, p);
    if (rc) return rc;
    return (int64_t)p.total;
}

extern "C" int eqb_conv_stack_forward(const float *x, int B, int cin, int H, int W, const float *const *filters,
                                      const float *const *biases, const float *const *scales,
                                      const float *const *shifts, int cout, int k, int num_group, int num_layers,
                                      float *act, void *workspace, int64_t workspace_bytes, void *stream) {
    EQB_NVTX_RANGE();
    ConvPlan p;
    const int rc = conv_make_plan(B, cin, H, W, cout, k, num_group, num_layers, p);
    if (rc) return rc;
    if (B == 0) return 0;
    EQB_REQUIRE(x && filters && act && workspace, "eqb_conv_stack_forward: null pointer");
    EQB_REQUIRE(workspace_bytes >= (int64_t)p.total, "eqb_conv_stack_forward: workspace %lld < %lld bytes",
                (long long)workspace_bytes, (long long)p.total);
    EQB_REQUIRE(((uintptr_t)workspace & 15) == 0, "eqb_conv_stack_forward: workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    const int L = num_layers;
    const float *in = x;
    int h = H, w = W;
    float *lay = p.tc ? (float *)(ws + p.off_lay) : nullptr;
    const __half *tin_hi = nullptr, *tin_lo = nullptr;   // fp16 pair feeding the next tcgen05 layer
    if (p.tc) {
        // operand scales are chained on the device from max |x| (conv_stack_tc.cu)
        int e = tc_absmax(x, B, (size_t)cin * H * W, (float *)(ws + p.off_absmax), st);   // max |x| of every image
        if (e) return e;
    }
    for (int l = 0; l < L; ++l) {
        EQB_REQUIRE(filters[l], "eqb_conv_stack_forward: null filter for layer %d", l);
        if (p.tc && l <= L - 2) {
            // ---- layer on the tensor cores: fp16 hi/lo NHWC in, fp16 pair or (last such layer) fp32 NCHW out ----------
            float *vec = (float *)(ws + p.off_vec[l]);
            conv_vec_kernel<<<1, 256, 0, st>>>(biases ? biases[l] : nullptr, scales ? scales[l] : nullptr,
                                               shifts ? shifts[l] : nullptr, vec, p.N, p.Npad);
            float *lay_l = lay + (size_t)l * B * LAY_FLOATS;          // B records of this layer
            const float *bound_in = l == 0 ? (const float *)(ws + p.off_absmax) : lay_l + LAY_INBOUND;
            int e = ctc_layer_stats(filters[l], vec, p.N, p.Npad, p.K[l], bound_in, l == 0 ? 1 : LAY_FLOATS, B, 1, lay_l,
                                    lay_l + (size_t)B * LAY_FLOATS, (float *)(ws + p.off_rowstat), st);
            if (e) return e;
            unsigned char *wp = (unsigned char *)(ws + p.off_wp[l]);
            const int cpad_in = l == 0 ? p.cpad0 : p.Npad;
            e = ctc_pack(filters[l], lay_l, p.N, p.cin[l], cpad_in, k, p.Npad, wp, st);
            if (e) return e;
            if (l == 0) {   // the network input becomes the first fp16 pair (channels padded to a whole K atom)
                __half *xh = (__half *)(ws + p.off_xin[0]), *xl = (__half *)(ws + p.off_xin[1]);
                e = ctc_input_split(x, (const float *)(ws + p.off_absmax), B, cin, H, W, cpad_in, xh, xl, st);
                if (e) return e;
                tin_hi = xh; tin_lo = xl;
            }
            const bool last_inner = l == L - 2;
            float *out32 = last_inner ? (float *)(ws + p.off_buf[0]) : nullptr;
            __half *oh = last_inner ? nullptr : (__half *)(ws + p.off_h[l & 1][0]);
            __half *ol = last_inner ? nullptr : (__half *)(ws + p.off_h[l & 1][1]);
            e = ctc_conv_layer(tin_hi, tin_lo, B, cpad_in, h, w, k, wp, vec, lay_l, p.N, p.Npad, 1, out32, oh, ol, p.Npad, st);
            if (e) return e;
            in = out32; tin_hi = oh; tin_lo = ol;
            h = p.Ho[l]; w = p.Wo[l];
            continue;
        }
        if (l == L - 1 && p.fold) {
            // `in` = input of the last layer, (B, cin_l, h, w) fp32
            const int Kl = p.K[l];
            double *Wfold = (double *)(ws + p.off_fold), *bmean = (double *)(ws + p.off_bmean), *Swin = (double *)(ws + p.off_win);
            conv_fold_weights_kernel<<<64, 256, 0, st>>>(filters[l], biases ? biases[l] : nullptr, cout, num_group, Kl, Wfold, bmean);
            EQB_CUDA(cudaFuncSetAttribute(conv_window_sums_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            conv_window_sums_kernel<<<(unsigned)(B * p.cin[l]), w <= 96 ? 96 : w <= 128 ? 128 : 256, p.win_smem, st>>>(in, p.cin[l], h, w, k, Swin);
            const double inv = 1.0 / ((double)cout * (double)p.P[l]);
            conv_fold_apply_kernel<<<B, 256, 0, st>>>(Swin, Wfold, bmean, Kl, num_group, inv, act);
            return finish_launch("conv_fold_apply_kernel");
        }
        float *Wt = (float *)(ws + p.off_wt[l]), *vec = (float *)(ws + p.off_vec[l]);
        conv_pack_kernel<<<64, 256, 0, st>>>(filters[l], Wt, p.N, p.K[l], p.Kpad[l], p.Npad);
        const bool last = l == L - 1;
        conv_vec_kernel<<<1, 256, 0, st>>>(biases ? biases[l] : nullptr, (!last && scales) ? scales[l] : nullptr,
                                           (!last && shifts) ? shifts[l] : nullptr, vec, p.N, p.Npad);
        ConvArgs a{};
        a.x = in; a.B = B; a.cin = p.cin[l]; a.H = h; a.W = w; a.ksz = k; a.Ho = p.Ho[l]; a.Wo = p.Wo[l]; a.P = p.P[l];
        a.K = p.K[l]; a.Kpad = p.Kpad[l]; a.N = p.N; a.Npad = p.Npad; a.relu = !last;
        a.Wt = Wt; a.bias = vec; a.scale = vec + p.Npad; a.shift = vec + 2 * p.Npad;
        if (last) {
            a.y = nullptr; a.S_part = (double *)(ws + p.off_S);
            a.tiles = p.tiles; a.chunks = p.chunks; a.tiles_per_chunk = p.tiles_per_chunk;
        } else {
            a.y = a.y_hi ? nullptr : (float *)(ws + p.off_buf[l & 1]); a.S_part = nullptr;
            a.tiles = (p.P[l] + CV_TM - 1) / CV_TM;
            const int target = num_sms() * 2 * 4;
            int chunks = (target + B - 1) / B;
            chunks = chunks < 1 ? 1 : chunks > a.tiles ? a.tiles : chunks;
            a.tiles_per_chunk = (a.tiles + chunks - 1) / chunks;
            a.chunks = (a.tiles + a.tiles_per_chunk - 1) / a.tiles_per_chunk;
        }
        int e;
        switch (p.Npad / 32) {
            case 1: e = conv_launch<1>(a, p.smem[l], st); break;
            case 2: e = conv_launch<2>(a, p.smem[l], st); break;
            case 4: e = conv_launch<4>(a, p.smem[l], st); break;
            default: e = conv_launch<8>(a, p.smem[l], st); break;
        }
        if (e) return e;
        in = a.y; tin_hi = a.y_hi; tin_lo = a.y_lo;
        h = p.Ho[l]; w = p.Wo[l];
    }
    const double inv_count = 1.0 / ((double)cout * (double)p.P[L - 1]);
    conv_pool_finish_kernel<<<B, 32 * ((num_group + 31) / 32), 0, st>>>((const double *)(ws + p.off_S), p.chunks, p.Npad,
                                                                         cout, num_group, inv_count, act);
    return finish_launch("conv_pool_finish_kernel");
}
