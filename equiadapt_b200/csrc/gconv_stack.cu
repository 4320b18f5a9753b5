// Fused group-equivariant conv stack -> group activations (a4..a6).
//
// Reference: CustomEquivariantNetwork.forward (custom_equivariant_networks.py:80-93) =
//   lift conv k x k -> [ReLU -> 1x1 regular group conv] x (L-1) -> mean over (Cout, H', W').
// The reference materialises every (B, Cout*|G|, H', W') intermediate in HBM (4.4 GB each at
// BASELINE cfg2) plus separate bias / ReLU passes.  Here the whole stack runs per tile of 64 output
// pixels without leaving the SM:
//   im2col tile (K0 x 64) in shared memory -> GEMM with the lifted filter orbit -> +bias, ReLU in
//   registers -> written back to shared memory as the next layer's operand -> GEMM with the
//   group-circulant 1x1 orbit -> ... -> masked column sums accumulated in fp64 registers.
// The LAST layer is linear and followed only by the mean, so it is folded through the pool:
//   act[g] = sum_k S[k] * (sum_o W'[(o,g),k]) / (Cout*P) + mean(b),   S = spatial sum of its input,
// an |G| x K mat-vec per image in the finish kernel (same result up to fp32 summation order).
//
// This file is the fp32 SIMT implementation (8x8 register tiles, weights streamed K-major through a
// cp.async double buffer).  It handles any Cout*|G| <= 256.
#include "common.cuh"
#include "gconv_stack_tc.cuh"

namespace eqb {

int launch_lift_orbit(const float *w, float *out, int cout, int cin, int k, int N, int reflect, long long sn,
                      long long sk, cudaStream_t st);
int launch_regular_orbit(const float *w, float *out, int cout, int cin, int k, int N, int reflect, long long sn,
                         long long sk, cudaStream_t st);

constexpr int GT_TM = 64;      // pixels per tile
constexpr int GT_PITCH = 68;   // floats per operand row (64 pixels + 4 pad: conflict-free float4 epilogue stores)
constexpr int GT_KC = 16;      // K rows per streamed weight chunk
constexpr int GT_THREADS = 256;
constexpr int GT_MAX_GEMM = 8;
// per-call scalars ahead of the partial sums: max |x| of every image (tcgen05 path)
// and the per-image cross entropies / identity flags of the fused select (2 B doubles)
static inline size_t scratch_amax(int B) { return (((size_t)(B > 0 ? B : 1) * sizeof(float)) + 255) & ~(size_t)255; }
static inline size_t scratch_head(int B) { return scratch_amax(B) + (((size_t)(B > 0 ? B : 1) * 2 * sizeof(double) + 255) & ~(size_t)255); }

struct StackArgs {
    const float *x;
    int B, cin, H, W, ksz, Ho, Wo, P;
    int K0, K0pad, N, Npad, rows;
    int n_gemm;
    int relu_last;  // ReLU on the last executed GEMM layer (num_layers > 1)
    const float *Wt[GT_MAX_GEMM];    // [Kpad][Npad], K-major
    const float *bias[GT_MAX_GEMM];  // [Npad]
    int Kpad[GT_MAX_GEMM];
    double *S_part;  // [B][chunks][Npad]
    int tiles, chunks, tiles_per_chunk;
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

template <int NT>
__global__ void __launch_bounds__(GT_THREADS, 2) gconv_stack_kernel(const StackArgs a) {
    extern __shared__ __align__(16) float sm[];
    float *act = sm;                                   // [rows][GT_PITCH]
    float *wbuf = sm + (size_t)a.rows * GT_PITCH;      // [2][GT_KC][Npad]
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int b = blockIdx.x / a.chunks, ch = blockIdx.x - b * a.chunks;
    const int tile_begin = ch * a.tiles_per_chunk, tile_end = min(a.tiles, tile_begin + a.tiles_per_chunk);
    const int chunk_f4 = GT_KC * a.Npad / 4;  // float4 per weight chunk
    const int kk2 = a.ksz * a.ksz;
    const float *xb = a.x + (size_t)b * a.cin * a.H * a.W;

    double colsum[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) colsum[j] = 0.0;

    for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int p0 = tile * GT_TM;
        // ---- im2col: act[k][p] = x[c][oy+ky][ox+kx], zero rows up to K0pad, zero columns past P ------
        {
            const int p = tid & (GT_TM - 1), kq = tid >> 6;
            const int pix = p0 + p;
            const bool ok = pix < a.P;
            const int oy = ok ? pix / a.Wo : 0, ox = ok ? pix - oy * a.Wo : 0;
            const float *xp = xb + (size_t)oy * a.W + ox;
            for (int k = kq; k < a.K0pad; k += GT_THREADS / GT_TM) {
                float v = 0.f;
                if (ok && k < a.K0) {
                    const int c = k / kk2, rem = k - c * kk2, ky = rem / a.ksz, kx = rem - ky * a.ksz;
                    v = __ldg(xp + ((size_t)c * a.H + ky) * a.W + kx);
                }
                act[k * GT_PITCH + p] = v;
            }
        }
        __syncthreads();

        for (int l = 0; l < a.n_gemm; ++l) {
            float acc[8][NT];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;

            const float *Wt = a.Wt[l];
            const int nchunks = a.Kpad[l] / GT_KC;
            for (int i = tid; i < chunk_f4; i += GT_THREADS) cp_async16(wbuf + 4 * i, Wt + 4 * i);
            cp_async_commit();
            for (int kc = 0; kc < nchunks; ++kc) {
                float *wb = wbuf + (kc & 1) * (GT_KC * a.Npad);
                if (kc + 1 < nchunks) {
                    float *wn = wbuf + ((kc + 1) & 1) * (GT_KC * a.Npad);
                    const float *src = Wt + (size_t)(kc + 1) * GT_KC * a.Npad;
                    for (int i = tid; i < chunk_f4; i += GT_THREADS) cp_async16(wn + 4 * i, src + 4 * i);
                    cp_async_commit();
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
                const float *ap = act + (size_t)kc * GT_KC * GT_PITCH + ty * 8;
#pragma unroll
                for (int kk = 0; kk < GT_KC; ++kk) {
                    const float4 a0 = *reinterpret_cast<const float4 *>(ap + kk * GT_PITCH);
                    const float4 a1 = *reinterpret_cast<const float4 *>(ap + kk * GT_PITCH + 4);
                    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    float wv[NT];
#pragma unroll
                    for (int j = 0; j < NT; ++j) wv[j] = wb[kk * a.Npad + j * 32 + tx];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < NT; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
                }
                __syncthreads();  // all reads of wb (and, on the last chunk, of act) are done
            }

            const float *bias = a.bias[l];
            if (l + 1 < a.n_gemm) {
                // +bias, ReLU, becomes the next layer's operand (row = channel, column = pixel)
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const int n = j * 32 + tx;
                    const float bv = bias[n];
                    float4 o0, o1;
                    o0.x = fmaxf(acc[0][j] + bv, 0.f); o0.y = fmaxf(acc[1][j] + bv, 0.f);
                    o0.z = fmaxf(acc[2][j] + bv, 0.f); o0.w = fmaxf(acc[3][j] + bv, 0.f);
                    o1.x = fmaxf(acc[4][j] + bv, 0.f); o1.y = fmaxf(acc[5][j] + bv, 0.f);
                    o1.z = fmaxf(acc[6][j] + bv, 0.f); o1.w = fmaxf(acc[7][j] + bv, 0.f);
                    *reinterpret_cast<float4 *>(act + n * GT_PITCH + ty * 8) = o0;
                    *reinterpret_cast<float4 *>(act + n * GT_PITCH + ty * 8 + 4) = o1;
                }
                __syncthreads();
            } else {
                // last executed layer: masked spatial sum of this thread's 8 pixels
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const float bv = bias[j * 32 + tx];
                    float s = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float v = acc[i][j] + bv;
                        if (a.relu_last) v = fmaxf(v, 0.f);
                        if (p0 + ty * 8 + i < a.P) s += v;
                    }
                    colsum[j] += (double)s;
                }
            }
        }
    }

    // ---- reduce the 8 pixel groups (warps) and emit this chunk's partial sums ------------------------
    __syncthreads();
    double *red = reinterpret_cast<double *>(sm);  // [8][Npad] doubles, fits: rows*68*4 >= 8*Npad*8
#pragma unroll
    for (int j = 0; j < NT; ++j) red[ty * a.Npad + j * 32 + tx] = colsum[j];
    __syncthreads();
    for (int n = tid; n < a.Npad; n += GT_THREADS) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w * a.Npad + n];
        a.S_part[((size_t)b * a.chunks + ch) * a.Npad + n] = s;
    }
}

// M[k][g]: what multiplies the spatial sums.  fold: sum_o W_last'[(o,g),k] for the 1x1 regular layer
// (k = i*|G|+h, W'[(o,g),(i,h)] = W[o,i,slice(g,h)]);  no fold (single layer): indicator k%|G|==g.
__device__ __forceinline__ int fold_src_slice(int g, int h, int N) {
    if (g < N) return h < N ? (h - g + N) % N : N + (h - N + g) % N;
    const int gp = g - N;
    return h < N ? N + (h + gp) % N : (h - N - gp + N) % N;
}

__global__ void build_fold_matrix_kernel(const float *__restrict__ w_last, double *__restrict__ M, int cout, int N,
                                         int G, int Npad, int fold) {
    const int total = Npad * G;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int k = t / G, g = t - k * G;
        double v = 0.0;
        if (k < cout * G) {
            if (fold) {
                const int i = k / G, h = k - i * G;
                const int src = fold_src_slice(g, h, N);
                for (int o = 0; o < cout; ++o) v += (double)w_last[((size_t)o * cout + i) * G + src];
            } else {
                v = (k % G == g) ? 1.0 : 0.0;
            }
        }
        M[t] = v;
    }
}

__global__ void expand_bias_kernel(const float *__restrict__ bias, float *__restrict__ out, int N, int G, int Npad) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < Npad; n += gridDim.x * blockDim.x)
        out[n] = (bias && n < N) ? bias[n / G] : 0.f;
}

// Fused group pool / select (a9 + a13) riding on the finish kernel: what eqb_group_pool_select computes from the activations
// (small_ops.cu: same per-sample arithmetic, same order of the batch sums, hence the same bits), so a canonicalizer step needs
// no separate select launch.  All pointers null -> plain finish.
struct SelectOut {
    int32_t *idx;
    float *rotation, *reflection, *onehot, *stats;   // reflection / onehot may be null
    double *per_image;                               // scratch: [B] cross entropies, then [B] identity flags
    unsigned int *ticket;                            // zero at entry, reset by the last block
    int N;                                           // rotations (|G| = N or 2 N)
};

// one block per image: S = sum of chunk partials; act[g] = S . M[:,g] / (Cout*P) + mean(last bias); warp w owns the group
// elements g = w, w + 8, ...; [select:] thread 0 picks the element, the last block to finish reduces the batch statistic
__global__ void __launch_bounds__(256) gconv_finish_kernel(const double *__restrict__ S_part,
                                                           const double *__restrict__ M,
                                                           const float *__restrict__ last_bias, int cout, int chunks,
                                                           int Npad, int G, double inv_count, float *__restrict__ act,
                                                           const SelectOut sel) {
    extern __shared__ double S[];  // [Npad]
    __shared__ float act_s[64];
    __shared__ double s_ce[8], s_id[8];
    __shared__ unsigned int s_last;
    const int b = blockIdx.x, B = gridDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = threadIdx.x; n < Npad; n += blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < chunks; ++c) s += S_part[((size_t)b * chunks + c) * Npad + n];
        S[n] = s;
    }
    __syncthreads();
    double bmean = 0.0;
    if (last_bias) {
        for (int o = 0; o < cout; ++o) bmean += (double)last_bias[o];
        bmean /= cout;
    }
    for (int g = warp; g < G; g += 8) {
        double v = 0.0;
        for (int n = lane; n < Npad; n += 32) v += S[n] * M[(size_t)n * G + g];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
            const float a = (float)(v * inv_count + bmean);
            act[(size_t)b * G + g] = a;
            if (g < 64) act_s[g] = a;
        }
    }
    if (!sel.idx) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        // per-sample part of group_pool_select_kernel, verbatim
        const float *a = act_s;
        float best = a[0];
        int bi = 0;
        for (int g = 1; g < G; ++g) {
            const float v = a[g];
            if (v > best || (v != v && best == best)) {
                best = v;
                bi = g;
            }
        }
        float se = 0.f;
        for (int g = 0; g < G; ++g) se += expf(a[g] - best);
        sel.per_image[b] = (double)((best + logf(se)) - a[0]);
        sel.per_image[B + b] = bi == 0 ? 1.0 : 0.0;
        sel.idx[b] = bi;
        const int N = sel.N, r = bi % N;
        const float step = 360.0f / (float)N;
        sel.rotation[b] = (r < (N + 1) / 2) ? __fmul_rn(step, (float)r) : 360.0f - __fmul_rn(step, (float)(N - r));
        if (sel.reflection) sel.reflection[b] = bi >= N ? 1.f : 0.f;
        if (sel.onehot)
            for (int g = 0; g < G; ++g) sel.onehot[(size_t)b * G + g] = g == bi ? 1.f : 0.f;
        __threadfence();
        s_last = atomicAdd(sel.ticket, 1u) == (unsigned int)(B - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    // last block: the batch statistic in the summation order of group_pool_select_kernel's single-CTA path
    __threadfence();
    double ce = 0.0, ident = 0.0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        ce += __ldcg(sel.per_image + i);
        ident += __ldcg(sel.per_image + B + i);
    }
    for (int o = 16; o > 0; o >>= 1) {
        ce += __shfl_xor_sync(0xffffffffu, ce, o);
        ident += __shfl_xor_sync(0xffffffffu, ident, o);
    }
    if (lane == 0) {
        s_ce[warp] = ce;
        s_id[warp] = ident;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0, d = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            c += s_ce[w];
            d += s_id[w];
        }
        sel.stats[0] = (float)c;
        sel.stats[1] = (float)d;
        sel.stats[2] = (float)B;
        sel.stats[3] = (float)(c / (double)B);
        sel.stats[4] = (float)(d / (double)B);
        *sel.ticket = 0u;                 // ready for the next launch
    }
}

struct StackPlan {
    int G, N, Npad, K0, K0pad, Ho, Wo, P, rows, n_gemm, tiles, chunks, tiles_per_chunk;
    size_t off_wt[GT_MAX_GEMM], off_bias[GT_MAX_GEMM], off_M, off_tc, off_ticket, off_S, total;
    size_t smem;
    // tcgen05 path (gconv_stack_tc.cu): 128-pixel tiles, its own chunking
    bool tc;
    int tc_tiles, tc_chunks, tc_tiles_per_chunk;
    // CTA-pair kernel: 256-pixel pair-tiles
    bool tc_pair;
    int tc2_tiles, tc2_chunks, tc2_tiles_per_chunk;
};

static int make_plan(int B, int cin, int H, int W, int cout, int k, int N, int reflect, int L, StackPlan &p) {
    EQB_REQUIRE(B >= 0 && cin > 0 && H > 0 && W > 0 && cout > 0 && k > 0 && N > 0 && L >= 1,
                "eqb_gconv_stack: bad argument");
    EQB_REQUIRE(H >= k && W >= k, "eqb_gconv_stack: image %dx%d smaller than the %dx%d kernel", H, W, k, k);
    p.G = N * (reflect ? 2 : 1);
    p.N = cout * p.G;
    EQB_UNSUPPORTED(p.N > 256, "eqb_gconv_stack: out_channels*|G| = %d > 256 not supported by this build", p.N);
    EQB_UNSUPPORTED(L - 1 > GT_MAX_GEMM, "eqb_gconv_stack: more than %d layers not supported", GT_MAX_GEMM + 1);
    const int nt = p.N <= 32 ? 1 : p.N <= 64 ? 2 : p.N <= 128 ? 4 : 8;
    p.Npad = 32 * nt;
    p.K0 = cin * k * k;
    p.K0pad = (p.K0 + GT_KC - 1) / GT_KC * GT_KC;
    p.Ho = H - k + 1;
    p.Wo = W - k + 1;
    p.P = p.Ho * p.Wo;
    p.rows = p.K0pad > p.Npad ? p.K0pad : p.Npad;
    p.n_gemm = L == 1 ? 1 : L - 1;
    p.tiles = (p.P + GT_TM - 1) / GT_TM;
    // enough CTAs for ~8 waves at 2 CTAs/SM, but never fewer than ~4 tiles per CTA when avoidable
    const int target = num_sms() * 2 * 8;
    int chunks = B > 0 ? (target + B - 1) / B : 1;
    if (chunks > p.tiles) chunks = p.tiles;
    if (chunks < 1) chunks = 1;
    p.tiles_per_chunk = (p.tiles + chunks - 1) / chunks;
    p.chunks = (p.tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
    p.smem = ((size_t)p.rows * GT_PITCH + 2 * GT_KC * p.Npad) * sizeof(float);
    EQB_UNSUPPORTED(p.smem > 200 * 1024, "eqb_gconv_stack: Cin*k*k = %d too large for the shared-memory tile", p.K0);
    size_t off = 0;
    for (int l = 0; l < p.n_gemm; ++l) {
        const size_t kp = l == 0 ? p.K0pad : p.Npad;
        p.off_wt[l] = off;
        off += kp * p.Npad * sizeof(float);
    }
    for (int l = 0; l < p.n_gemm; ++l) {
        p.off_bias[l] = off;
        off += (size_t)p.Npad * sizeof(float);
    }
    p.off_M = off;
    off += (size_t)p.Npad * p.G * sizeof(double);
    p.tc = tc_eligible(p.N, p.K0, p.n_gemm);
    p.tc_pair = false;
    p.off_tc = off;
    int max_chunks = p.chunks;
    if (p.tc) {
        off = (off + 1023) & ~(size_t)1023;
        p.off_tc = off;
        off += tc_pack_bytes(p.N, p.K0);
        p.tc_tiles = (p.P + 127) / 128;
        // work items of <= 8 tiles; finer when the batch alone cannot occupy every SM four times over
        long long tpc = ((long long)(B > 0 ? B : 1) * p.tc_tiles + 4LL * num_sms() - 1) / (4LL * num_sms());
        p.tc_tiles_per_chunk = (int)(tpc < 1 ? 1 : tpc > 8 ? 8 : tpc);
        p.tc_chunks = (p.tc_tiles + p.tc_tiles_per_chunk - 1) / p.tc_tiles_per_chunk;
        if (p.tc_chunks > max_chunks) max_chunks = p.tc_chunks;
        p.tc_pair = tc_use_pair(p.N, p.K0);
        p.tc2_tiles = (p.P + 255) / 256;
        p.tc2_tiles_per_chunk = p.tc_tiles_per_chunk / 2 > 0 ? p.tc_tiles_per_chunk / 2 : 1;
        p.tc2_chunks = (p.tc2_tiles + p.tc2_tiles_per_chunk - 1) / p.tc2_tiles_per_chunk;
        if (2 * p.tc2_chunks > max_chunks) max_chunks = 2 * p.tc2_chunks;   // two partial-sum rows per chunk
    }
    off = (off + 15) & ~(size_t)15;
    p.off_ticket = off;      // last-block ticket of the fused select (zeroed by the pack call, reset by the kernel)
    off += 64;
    p.off_S = off;
    off += scratch_head(B);
    off += (size_t)(B > 0 ? B : 1) * max_chunks * p.Npad * sizeof(double);
    p.total = off;
    return 0;
}

template <int NT>
static int launch_stack(const StackArgs &a, size_t smem, cudaStream_t st) {
    static PerDeviceOnce configured;
    if (configured.first()) {
        EQB_CUDA(cudaFuncSetAttribute(gconv_stack_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    gconv_stack_kernel<NT><<<(unsigned)(a.B * a.chunks), GT_THREADS, smem, st>>>(a);
    return finish_launch("gconv_stack_kernel");
}

}  // namespace eqb

using namespace eqb;

// Workspace layout: [ packed weights | expanded biases | fold matrix ] = "packed" part (depends on the
// parameters only) followed by the per-call partial sums.
extern "C" int64_t eqb_gconv_stack_workspace_bytes(int B, int cin, int H, int W, int cout, int k, int num_rotations,
                                                   int reflect, int num_layers) {
    StackPlan p;
    const int rc = make_plan(B, cin, H, W, cout, k, num_rotations, reflect, num_layers, p);
    if (rc) return rc;
    return (int64_t)p.total;
}

extern "C" int64_t eqb_gconv_stack_packed_bytes(int cin, int cout, int k, int num_rotations, int reflect,
                                                int num_layers) {
    StackPlan p;
    const int rc = make_plan(1, cin, k, k, cout, k, num_rotations, reflect, num_layers, p);
    if (rc) return rc;
    return (int64_t)p.off_S;
}

extern "C" int eqb_gconv_stack_pack(const float *lift_w, const float *lift_b, const float *const *reg_w,
                                    const float *const *reg_b, int cin, int cout, int k, int num_rotations, int reflect,
                                    int num_layers, void *packed, int64_t packed_bytes, void *stream) {
    EQB_NVTX_RANGE();
    StackPlan p;
    const int rc = make_plan(1, cin, k, k, cout, k, num_rotations, reflect, num_layers, p);
    if (rc) return rc;
    EQB_REQUIRE(lift_w && packed, "eqb_gconv_stack_pack: null pointer");
    EQB_REQUIRE(num_layers == 1 || (reg_w && reg_b), "eqb_gconv_stack_pack: null layer table");
    EQB_REQUIRE(packed_bytes >= (int64_t)p.off_S, "eqb_gconv_stack_pack: buffer %lld < %lld bytes",
                (long long)packed_bytes, (long long)p.off_S);
    EQB_REQUIRE(((uintptr_t)packed & 15) == 0, "eqb_gconv_stack_pack: buffer must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)packed;
    const int L = num_layers, reflect01 = reflect != 0;
    // filter orbits as zero-padded K-major GEMM operands, expanded biases, fold matrix
    EQB_CUDA(cudaMemsetAsync(ws, 0, p.off_bias[0], st));
    EQB_CUDA(cudaMemsetAsync(ws + p.off_ticket, 0, 64, st));
    int e = launch_lift_orbit(lift_w, (float *)(ws + p.off_wt[0]), cout, cin, k, num_rotations, reflect01, 1, p.Npad, st);
    if (e) return e;
    for (int l = 1; l < p.n_gemm; ++l) {
        EQB_REQUIRE(reg_w[l - 1], "eqb_gconv_stack_pack: null weight for layer %d", l);
        e = launch_regular_orbit(reg_w[l - 1], (float *)(ws + p.off_wt[l]), cout, cout, 1, num_rotations, reflect01, 1,
                                 p.Npad, st);
        if (e) return e;
    }
    for (int l = 0; l < p.n_gemm; ++l) {
        const float *bsrc = l == 0 ? lift_b : reg_b[l - 1];
        expand_bias_kernel<<<1, 256, 0, st>>>(bsrc, (float *)(ws + p.off_bias[l]), p.N, p.G, p.Npad);
    }
    const float *w_last = L > 1 ? reg_w[L - 2] : nullptr;
    EQB_REQUIRE(L == 1 || w_last, "eqb_gconv_stack_pack: null weight for the last layer");
    build_fold_matrix_kernel<<<(p.Npad * p.G + 255) / 256, 256, 0, st>>>(w_last, (double *)(ws + p.off_M), cout,
                                                                         num_rotations, p.G, p.Npad, L > 1);
    e = finish_launch("eqb_gconv_stack_pack");
    if (e) return e;
    if (p.tc)  // the same operands as UMMA hi / lo images for the tcgen05 kernel
        return tc_pack((const float *)(ws + p.off_wt[0]), p.K0, (const float *)(ws + p.off_wt[1]),
                       (const float *)(ws + p.off_bias[0]), p.Npad, p.N, (unsigned char *)(ws + p.off_tc), st);
    return 0;
}

static int stack_run(const float *x, const float *x_absmax, int B, int cin, int H, int W, const void *packed,
                     const float *last_bias, int cout, int k, int num_rotations, int reflect, int num_layers, float *act,
                     SelectOut sel, void *scratch, int64_t scratch_bytes, void *stream) {
    StackPlan p;
    const int rc = make_plan(B, cin, H, W, cout, k, num_rotations, reflect, num_layers, p);
    if (rc) return rc;
    if (B == 0) return 0;
    EQB_REQUIRE(x && packed && act && scratch, "eqb_gconv_stack_run: null pointer");
    EQB_REQUIRE(scratch_bytes >= (int64_t)(p.total - p.off_S), "eqb_gconv_stack_run: scratch %lld < %lld bytes",
                (long long)scratch_bytes, (long long)(p.total - p.off_S));
    EQB_REQUIRE(((uintptr_t)packed & 15) == 0 && ((uintptr_t)scratch & 7) == 0, "eqb_gconv_stack_run: misaligned buffer");
    EQB_REQUIRE((long long)B * p.chunks < (1LL << 31), "eqb_gconv_stack_run: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    const char *ws = (const char *)packed;
    const int L = num_layers;

    StackArgs a{};
    a.x = x; a.B = B; a.cin = cin; a.H = H; a.W = W; a.ksz = k; a.Ho = p.Ho; a.Wo = p.Wo; a.P = p.P;
    a.K0 = p.K0; a.K0pad = p.K0pad; a.N = p.N; a.Npad = p.Npad; a.rows = p.rows;
    a.n_gemm = p.n_gemm; a.relu_last = L > 1;
    for (int l = 0; l < p.n_gemm; ++l) {
        a.Wt[l] = (const float *)(ws + p.off_wt[l]);
        a.bias[l] = (const float *)(ws + p.off_bias[l]);
        a.Kpad[l] = l == 0 ? p.K0pad : p.Npad;
    }
    a.S_part = (double *)((char *)scratch + scratch_head(B));
    a.tiles = p.tiles; a.chunks = p.chunks; a.tiles_per_chunk = p.tiles_per_chunk;
    int e, chunks = p.chunks;
    if (p.tc) {
        TcArgs t{};
        t.x = x; t.B = B; t.cin = cin; t.H = H; t.W = W; t.ksz = k; t.Wo = p.Wo; t.P = p.P;
        t.K0 = p.K0; t.N = p.N; t.Npad = p.Npad;
        t.bias1 = a.bias[0]; t.bias2 = a.bias[1];
        t.wpack = (const unsigned char *)(ws + p.off_tc);
        t.S_part = a.S_part;
        if (x_absmax) {
            t.absmax = x_absmax;          // the producer of x (eqb_crop_resize_aa_absmax) already reduced max |x| per image
        } else {
            t.absmax = (const float *)scratch;
            e = tc_absmax(x, B, (size_t)cin * H * W, (float *)scratch, st);
            if (e) return e;
        }
        t.tiles = p.tc_tiles; t.chunks = p.tc_chunks; t.tiles_per_chunk = p.tc_tiles_per_chunk;
        t.tiles2 = p.tc2_tiles; t.chunks2 = p.tc2_chunks; t.tiles_per_chunk2 = p.tc2_tiles_per_chunk;
        chunks = p.tc_pair ? 2 * p.tc2_chunks : p.tc_chunks;
        e = tc_launch(t, st);
    } else {
        switch (p.Npad / 32) {
            case 1: e = launch_stack<1>(a, p.smem, st); break;
            case 2: e = launch_stack<2>(a, p.smem, st); break;
            case 4: e = launch_stack<4>(a, p.smem, st); break;
            default: e = launch_stack<8>(a, p.smem, st); break;
        }
    }
    if (e) return e;
    const double inv_count = 1.0 / ((double)cout * (double)p.P);
    if (sel.idx) {
        EQB_UNSUPPORTED(p.G > 64, "eqb_gconv_stack_run_select: more than 64 group elements");
        sel.per_image = (double *)((char *)scratch + scratch_amax(B));
        sel.ticket = (unsigned int *)(const_cast<char *>(ws) + p.off_ticket);
        sel.N = num_rotations;
    }
    gconv_finish_kernel<<<B, 256, p.Npad * sizeof(double), st>>>(a.S_part, (const double *)(ws + p.off_M),
                                                                L > 1 ? last_bias : nullptr, cout, chunks, p.Npad, p.G,
                                                                inv_count, act, sel);
    return finish_launch("gconv_finish_kernel");
}

extern "C" int eqb_gconv_stack_run(const float *x, int B, int cin, int H, int W, const void *packed,
                                   const float *last_bias, int cout, int k, int num_rotations, int reflect,
                                   int num_layers, float *act, void *scratch, int64_t scratch_bytes, void *stream) {
    EQB_NVTX_RANGE();
    return stack_run(x, nullptr, B, cin, H, W, packed, last_bias, cout, k, num_rotations, reflect, num_layers, act, SelectOut{},
                     scratch, scratch_bytes, stream);
}

extern "C" int eqb_gconv_stack_run_scaled(const float *x, const float *x_absmax, int B, int cin, int H, int W,
                                          const void *packed, const float *last_bias, int cout, int k, int num_rotations,
                                          int reflect, int num_layers, float *act, void *scratch, int64_t scratch_bytes,
                                          void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(x_absmax, "eqb_gconv_stack_run_scaled: null absmax");
    return stack_run(x, x_absmax, B, cin, H, W, packed, last_bias, cout, k, num_rotations, reflect, num_layers, act, SelectOut{},
                     scratch, scratch_bytes, stream);
}

extern "C" int eqb_gconv_stack_run_select(const float *x, const float *x_absmax, int B, int cin, int H, int W,
                                          void *packed, const float *last_bias, int cout, int k, int num_rotations,
                                          int reflect, int num_layers, float *act, int32_t *idx, float *rotation,
                                          float *reflection, float *onehot, float *stats, void *scratch,
                                          int64_t scratch_bytes, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(stats && (B == 0 || (idx && rotation)), "eqb_gconv_stack_run_select: null pointer");
    if (B == 0) {   // the statistic of an empty batch, as eqb_group_pool_select leaves it
        EQB_CUDA(cudaMemsetAsync(stats, 0, 5 * sizeof(float), (cudaStream_t)stream));
        return 0;
    }
    SelectOut sel{};
    sel.idx = idx; sel.rotation = rotation; sel.reflection = reflect ? reflection : nullptr; sel.onehot = onehot; sel.stats = stats;
    return stack_run(x, x_absmax, B, cin, H, W, packed, last_bias, cout, k, num_rotations, reflect, num_layers, act, sel,
                     scratch, scratch_bytes, stream);
}

extern "C" int eqb_debug_last_stall(int *out5) { return tc_last_stall(out5); }
extern "C" int eqb_debug_stack_trace(void *device_buffer, int tiles) { return tc_set_trace((long long *)device_buffer, tiles); }

extern "C" int eqb_gconv_stack_forward(const float *x, int B, int cin, int H, int W, const float *lift_w,
                                       const float *lift_b, const float *const *reg_w, const float *const *reg_b,
                                       int cout, int k, int num_rotations, int reflect, int num_layers, float *act,
                                       void *workspace, int64_t workspace_bytes, void *stream) {
    EQB_NVTX_RANGE();
    StackPlan p;
    const int rc = make_plan(B, cin, H, W, cout, k, num_rotations, reflect, num_layers, p);
    if (rc) return rc;
    if (B == 0) return 0;
    EQB_REQUIRE(workspace, "eqb_gconv_stack_forward: null workspace");
    EQB_REQUIRE(workspace_bytes >= (int64_t)p.total, "eqb_gconv_stack_forward: workspace %lld < %lld bytes",
                (long long)workspace_bytes, (long long)p.total);
    int e = eqb_gconv_stack_pack(lift_w, lift_b, reg_w, reg_b, cin, cout, k, num_rotations, reflect, num_layers,
                                 workspace, (int64_t)p.off_S, stream);
    if (e) return e;
    const float *b_last = num_layers > 1 ? reg_b[num_layers - 2] : nullptr;
    return eqb_gconv_stack_run(x, B, cin, H, W, workspace, b_last, cout, k, num_rotations, reflect, num_layers, act,
                               (char *)workspace + p.off_S, workspace_bytes - (int64_t)p.off_S, stream);
}
