// Latency-bound pieces of the hot path: pre-network crop+resize (a3), filter orbits (a4/a5),
// group pool / select + prior statistic (a9/a13), cosine activations (a12), frames (a14..a17).
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace eqb {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// -------------------------------------------------------------------------------------------------
// a3: CenterCrop + antialiased bilinear resize.  Separable triangle filter of ATen's
// _upsample_bilinear2d_aa (align_corners=False): per output index i, centre = scale*(i+0.5),
// support = max(scale,1), taps j in [max(0,int(centre-support+.5)), min(in,int(centre+support+.5))),
// w_j = max(0, 1-|(j-centre+.5)/max(scale,1)|) normalised to sum 1.
// Horizontal pass first, then vertical, as ATen does (same fma order per output as one thread per pixel).
//
// One CTA = one plane x a band of RS_ROWS output rows x up to RS_COLS output columns.  The tap tables of the
// band are built once into shared memory; thread (x, y) owns output column x: it keeps that column's
// horizontal weights in registers, filters the band's input rows straight from global memory (neighbouring
// lanes read overlapping 128-byte lines -> L1 hits) into a shared-memory strip, then the strip is filtered
// vertically.  Reads: the crop window once (+ band overlap), writes: the resized plane once.
// -------------------------------------------------------------------------------------------------
constexpr int AA_MAX_TAPS = 16;
constexpr int RS_ROWS = 16, RS_COLS = 128;

struct AaAxis {   // 17 words: odd stride -> conflict-free per-lane reads
    int lo_n;     // lo | n << 20
    float w[AA_MAX_TAPS];
};

__device__ __forceinline__ void aa_axis(int i, int in_size, float scale, AaAxis &ax) {
    const float support = scale >= 1.f ? scale : 1.f;
    const float invscale = scale >= 1.f ? 1.f / scale : 1.f;
    const float center = scale * (i + 0.5f);
    const int lo = max((int)(center - support + 0.5f), 0);
    const int hi = min((int)(center + support + 0.5f), in_size);
    const int n = min(hi - lo, AA_MAX_TAPS);
    float tot = 0.f;
    for (int j = 0; j < n; ++j) {
        const float v = fmaxf(1.f - fabsf((j + lo - center + 0.5f) * invscale), 0.f);
        ax.w[j] = v;
        tot += v;
    }
    for (int j = 0; j < n; ++j) ax.w[j] = tot != 0.f ? ax.w[j] / tot : 0.f;
    for (int j = n > 0 ? n : 0; j < AA_MAX_TAPS; ++j) ax.w[j] = 0.f;
    ax.lo_n = lo | (max(n, 0) << 20);
}

template <int MAXT>
__global__ void __launch_bounds__(256) crop_resize_aa_kernel(const float *__restrict__ x, float *__restrict__ y, int H,
                                                             int W, int top, int left, int ch, int cw, int oh, int ow,
                                                             float sy, float sx, int bands, int chunks, int strip_rows) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    AaAxis *tx = reinterpret_cast<AaAxis *>(rs_smem);         // [RS_COLS]
    AaAxis *ty = tx + RS_COLS;                                // [RS_ROWS]
    float *strip = reinterpret_cast<float *>(ty + RS_ROWS);   // [strip_rows][RS_COLS]
    int b = blockIdx.x;
    const int chunk = b % chunks; b /= chunks;
    const int band = b % bands;
    const size_t plane = b / bands;
    const int ox0 = chunk * RS_COLS, oy0 = band * RS_ROWS;
    const int ncols = min(RS_COLS, ow - ox0), nrows = min(RS_ROWS, oh - oy0);
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthreads = blockDim.x * blockDim.y;

    for (int t = tid; t < ncols + nrows; t += nthreads) {
        if (t < ncols) aa_axis(ox0 + t, cw, sx, tx[t]);
        else aa_axis(oy0 + t - ncols, ch, sy, ty[t - ncols]);
    }
    __syncthreads();
    const int row_lo = ty[0].lo_n & 0xfffff;
    const int row_hi = (ty[nrows - 1].lo_n & 0xfffff) + (ty[nrows - 1].lo_n >> 20);
    const int nr = min(row_hi - row_lo, strip_rows);
    const int ox = threadIdx.x;
    if (ox < ncols) {
        // horizontal pass: this thread's column weights live in registers
        const int lo = tx[ox].lo_n & 0xfffff, n = tx[ox].lo_n >> 20;
        float w[MAXT];
#pragma unroll
        for (int i = 0; i < MAXT; ++i) w[i] = tx[ox].w[i];
        const float *src = x + plane * (size_t)H * W + (size_t)(top + row_lo) * W + left + lo;
        // four rows per iteration: 4 x MAXT independent loads in flight per thread (r1q profile: this pass waits on
        // its global loads for 60 % of the kernel's stall cycles)
        const int rstep = blockDim.y;
        int r = threadIdx.y;
        for (; r + 3 * rstep < nr; r += 4 * rstep) {
            float v[4][MAXT];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float *rp = src + (size_t)(r + q * rstep) * W;
#pragma unroll
                for (int i = 0; i < MAXT; ++i) v[q][i] = i < n ? __ldg(rp + i) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float h = 0.f;
#pragma unroll
                for (int i = 0; i < MAXT; ++i)
                    if (i < n) h = fmaf(v[q][i], w[i], h);
                strip[(r + q * rstep) * RS_COLS + ox] = h;
            }
        }
        for (; r < nr; r += rstep) {
            const float *rp = src + (size_t)r * W;
            float h = 0.f;
#pragma unroll
            for (int i = 0; i < MAXT; ++i)
                if (i < n) h = fmaf(__ldg(rp + i), w[i], h);
            strip[r * RS_COLS + ox] = h;
        }
    }
    __syncthreads();
    if (ox < ncols) {
        for (int oy = threadIdx.y; oy < nrows; oy += blockDim.y) {
            const int lo = (ty[oy].lo_n & 0xfffff) - row_lo, n = ty[oy].lo_n >> 20;
            float acc = 0.f;
            for (int j = 0; j < n; ++j) acc = fmaf(strip[(lo + j) * RS_COLS + ox], ty[oy].w[j], acc);
            y[(plane * oh + (oy0 + oy)) * (size_t)ow + ox0 + ox] = acc;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// a4 / a5: filter orbits.  T_g(w)(y,x) = bilinear zero-fill sample of w rotated by angle_g about the
// kernel centre (kornia.rotate, positive = counter-clockwise), then hflip for reflected elements.
// Output element (n, kk) is written at out[n*sn + kk*sk] so the same kernel emits the conv2d layout
// (sn = K, sk = 1) and the K-major, zero-padded GEMM operand of the fused stack (sn = 1, sk = Npad).
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rotated_tap(const float *__restrict__ w, int k, int y, int x, int r, int N,
                                             bool mirror) {
    if (mirror) x = k - 1 - x;  // hflip AFTER the rotation: out(x) = rot(k-1-x)
    double c, s;
    rot_cs(r, N, 1.0, c, s);
    const double ctr = 0.5 * (k - 1);
    const double u = x - ctr, v = y - ctr;
    const double xs = ctr + c * u - s * v, ys = ctr + s * u + c * v;
    const double xf = floor(xs), yf = floor(ys);
    const float fx = (float)(xs - xf), fy = (float)(ys - yf);
    const int x0 = (int)xf, y0 = (int)yf;
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int xi = x0 + dx, yi = y0 + dy;
            if (xi >= 0 && xi < k && yi >= 0 && yi < k) {
                const float wt = (dy ? fy : 1.f - fy) * (dx ? fx : 1.f - fx);
                acc = fmaf(w[yi * k + xi], wt, acc);
            }
        }
    return acc;
}

__global__ void lift_orbit_kernel(const float *__restrict__ w, float *__restrict__ out, int cout, int cin, int k,
                                  int N, int G, long long sn, long long sk) {
    const int kk2 = k * k, K = cin * kk2;
    const long long total = (long long)cout * G * K;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(t % K);
        const int n = (int)(t / K);
        const int o = n / G, g = n % G;
        const int i = kk / kk2, yx = kk % kk2;
        out[n * sn + kk * sk] = rotated_tap(w + ((size_t)o * cin + i) * kk2, k, yx / k, yx % k, g % N, N, g >= N);
    }
}

// input-group slice of W feeding (g,h): C_N custom_group_equivariant_layers.py:283-293, D_N :420-449
__device__ __forceinline__ int regular_src_slice(int g, int h, int N) {
    if (g < N) return h < N ? (h - g + N) % N : N + (h - N + g) % N;
    const int gp = g - N;
    return h < N ? N + (h + gp) % N : (h - N - gp + N) % N;
}

__global__ void regular_orbit_kernel(const float *__restrict__ w, float *__restrict__ out, int cout, int cin, int k,
                                     int N, int G, long long sn, long long sk) {
    const int kk2 = k * k, K = cin * G * kk2;
    const long long total = (long long)cout * G * K;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(t % K);
        const int n = (int)(t / K);
        const int o = n / G, g = n % G;
        const int ih = kk / kk2, yx = kk % kk2;
        const int i = ih / G, h = ih % G;
        const int src = regular_src_slice(g, h, N);
        out[n * sn + kk * sk] =
            rotated_tap(w + (((size_t)o * cin + i) * G + src) * kk2, k, yx / k, yx % k, g % N, N, g >= N);
    }
}

// -------------------------------------------------------------------------------------------------
// a9 + a13: group pool / select.  One thread per sample (|G| <= 64 values), grid-stride; CE and
// identity counts are reduced per block with warp shuffles and finished, in block order (so the
// result is run-to-run deterministic), by the last block to arrive.
// -------------------------------------------------------------------------------------------------
// per-launch slots of the multi-block path (per-device globals; the host hands the slots out round-robin)
constexpr int SEL_BLOCKS = 1024, SEL_SLOTS = 8;
__device__ unsigned int g_sel_ticket[SEL_SLOTS];
__device__ double g_sel_partial[SEL_SLOTS][2 * SEL_BLOCKS];

__global__ void __launch_bounds__(256) group_pool_select_kernel(const float *__restrict__ act, int B, int N, int G,
                                                                int32_t *__restrict__ idx, float *__restrict__ rotation,
                                                                float *__restrict__ reflection,
                                                                float *__restrict__ onehot, float *__restrict__ stats,
                                                                int slot) {
    double ce = 0.0, ident = 0.0;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const float *a = act + (size_t)b * G;
        float best = a[0];
        int bi = 0;
        for (int g = 1; g < G; ++g) {
            const float v = a[g];
            if (v > best || (v != v && best == best)) {  // first max; NaN wins like torch.argmax
                best = v;
                bi = g;
            }
        }
        float se = 0.f;
        for (int g = 0; g < G; ++g) se += expf(a[g] - best);
        ce += (double)((best + logf(se)) - a[0]);  // logsumexp(a) - a[0] == CE(a, class 0)
        ident += bi == 0 ? 1.0 : 0.0;
        idx[b] = bi;
        // torch.linspace(0, 360, N+1)[r] in fp32 (two symmetric halves), discrete_group.py:110-112
        const int r = bi % N;
        const float step = 360.0f / (float)N;
        rotation[b] = (r < (N + 1) / 2) ? __fmul_rn(step, (float)r) : 360.0f - __fmul_rn(step, (float)(N - r));
        if (reflection) reflection[b] = bi >= N ? 1.f : 0.f;
        if (onehot)
            for (int g = 0; g < G; ++g) onehot[(size_t)b * G + g] = g == bi ? 1.f : 0.f;
    }
    __shared__ double s_ce[8], s_id[8];
    for (int o = 16; o > 0; o >>= 1) {
        ce += __shfl_xor_sync(0xffffffffu, ce, o);
        ident += __shfl_xor_sync(0xffffffffu, ident, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_ce[threadIdx.x >> 5] = ce;
        s_id[threadIdx.x >> 5] = ident;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0, d = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            c += s_ce[w];
            d += s_id[w];
        }
        if (gridDim.x > 1) {
            // several blocks (B > 8192): partials to a per-launch slot, the last block to finish adds them in block order
            g_sel_partial[slot][2 * blockIdx.x] = c;
            g_sel_partial[slot][2 * blockIdx.x + 1] = d;
            __threadfence();
            if (atomicAdd(&g_sel_ticket[slot], 1u) != gridDim.x - 1) return;
            __threadfence();
            c = d = 0.0;
            for (unsigned int k = 0; k < gridDim.x; ++k) {
                c += __ldcg(&g_sel_partial[slot][2 * k]);
                d += __ldcg(&g_sel_partial[slot][2 * k + 1]);
            }
            g_sel_ticket[slot] = 0u;
        }
        stats[0] = (float)c;
        stats[1] = (float)d;
        stats[2] = (float)B;
        stats[3] = (float)(c / (double)B);   // this batch's own means: read directly when there is one rank
        stats[4] = (float)(d / (double)B);
    }
}

// a12: cosine similarity to the reference vector, (|G|*B, V) -> (B,|G|).  One warp per row.
// ATen cosine_similarity: x/max(||x||,eps) . y/max(||y||,eps), eps = 1e-8.
__global__ void __launch_bounds__(256) cosine_act_kernel(const float *__restrict__ vec, const float *__restrict__ ref,
                                                         float *__restrict__ act, int B, int G, int V) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B * G) return;
    const float *v = vec + (size_t)row * V;
    float vv = 0.f, rr = 0.f;
    for (int i = lane; i < V; i += 32) {
        vv = fmaf(v[i], v[i], vv);
        rr = fmaf(ref[i], ref[i], rr);
    }
    for (int o = 16; o > 0; o >>= 1) {
        vv += __shfl_xor_sync(0xffffffffu, vv, o);
        rr += __shfl_xor_sync(0xffffffffu, rr, o);
    }
    const float nv = fmaxf(sqrtf(vv), 1e-8f), nr = fmaxf(sqrtf(rr), 1e-8f);
    float dot = 0.f;
    for (int i = lane; i < V; i += 32) dot = fmaf(ref[i] / nr, v[i] / nv, dot);
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) {
        const int g = row / B, b = row - g * B;  // rows are group-major
        act[(size_t)b * G + g] = dot;
    }
}

// N3: backward of cosine_act_kernel.  With v^ = v / nv, r^ = ref / nr, c = v^ . r^ (nv, nr clamped at eps as above):
//   d c / d v = r^ / nv - c v / nv^2   (second term absent when |v| sits below the clamp),   d c / d ref symmetric.
// One warp per row; d ref accumulates over rows with atomics (zeroed by the caller).
__global__ void __launch_bounds__(256) cosine_act_backward_kernel(const float *__restrict__ vec, const float *__restrict__ ref,
                                                                  const float *__restrict__ dact, float *__restrict__ dvec,
                                                                  float *__restrict__ dref, int B, int G, int V) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B * G) return;
    const float *v = vec + (size_t)row * V;
    float vv = 0.f, rr = 0.f;
    for (int i = lane; i < V; i += 32) {
        vv = fmaf(v[i], v[i], vv);
        rr = fmaf(ref[i], ref[i], rr);
    }
    for (int o = 16; o > 0; o >>= 1) {
        vv += __shfl_xor_sync(0xffffffffu, vv, o);
        rr += __shfl_xor_sync(0xffffffffu, rr, o);
    }
    const float sv = sqrtf(vv), sr = sqrtf(rr);
    const float nv = fmaxf(sv, 1e-8f), nr = fmaxf(sr, 1e-8f);
    float dot = 0.f;
    for (int i = lane; i < V; i += 32) dot = fmaf(ref[i] / nr, v[i] / nv, dot);
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const int g = row / B, b = row - g * B;
    const float go = dact[(size_t)b * G + g];
    const float kv = sv > 1e-8f ? dot / (nv * nv) : 0.f, kr = sr > 1e-8f ? dot / (nr * nr) : 0.f;
    for (int i = lane; i < V; i += 32) {
        if (dvec) dvec[(size_t)row * V + i] = go * (ref[i] / (nr * nv) - kv * v[i]);
        if (dref) atomicAdd(dref + i, go * (v[i] / (nv * nr) - kr * ref[i]));
    }
}

// -------------------------------------------------------------------------------------------------
// a14..a17: frames
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ float norm3(float x, float y, float z) {
    // torch.norm over 3 elements: sqrt of the plain fp32 sum of squares
    return sqrtf(x * x + y * y + z * z);
}

__device__ __forceinline__ void gs3(const float *v, float *R, bool modified) {
    float a0 = v[0], a1 = v[1], a2 = v[2];
    float n = norm3(a0, a1, a2);
    a0 /= n; a1 /= n; a2 /= n;
    float d = v[3] * a0 + v[4] * a1 + v[5] * a2;
    float b0 = v[3] - d * a0, b1 = v[4] - d * a1, b2 = v[5] - d * a2;
    n = norm3(b0, b1, b2);
    b0 /= n; b1 /= n; b2 /= n;
    float c0, c1, c2;
    const float d1 = v[6] * a0 + v[7] * a1 + v[8] * a2;
    if (!modified) {
        const float d2 = v[6] * b0 + v[7] * b1 + v[8] * b2;
        c0 = v[6] - d1 * a0 - d2 * b0;
        c1 = v[7] - d1 * a1 - d2 * b1;
        c2 = v[8] - d1 * a2 - d2 * b2;
    } else {
        c0 = v[6] - d1 * a0; c1 = v[7] - d1 * a1; c2 = v[8] - d1 * a2;
        const float d2 = c0 * b0 + c1 * b1 + c2 * b2;
        c0 -= d2 * b0; c1 -= d2 * b1; c2 -= d2 * b2;
    }
    n = norm3(c0, c1, c2);
    R[0] = a0; R[1] = a1; R[2] = a2;
    R[3] = b0; R[4] = b1; R[5] = b2;
    R[6] = c0 / n; R[7] = c1 / n; R[8] = c2 / n;
}

__global__ void gram_schmidt3_kernel(const float *__restrict__ v, float *__restrict__ R, int B, int modified) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float in[9], out[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) in[i] = v[(size_t)b * 9 + i];
    gs3(in, out, modified != 0);
#pragma unroll
    for (int i = 0; i < 9; ++i) R[(size_t)b * 9 + i] = out[i];
}

// y[b,:,n] = R[b] x[b,:,n]; x is (B,3,N): three coalesced streams per cloud, R in registers.
__global__ void __launch_bounds__(256) so3_apply_kernel(const float *__restrict__ x, const float *__restrict__ R,
                                                        float *__restrict__ y, int B, int N, int chunks) {
    const int b = blockIdx.x / chunks, ch = blockIdx.x % chunks;
    float r[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) r[i] = __ldg(R + (size_t)b * 9 + i);
    const float *xb = x + (size_t)b * 3 * N;
    float *yb = y + (size_t)b * 3 * N;
    for (int n = ch * blockDim.x + threadIdx.x; n < N; n += chunks * blockDim.x) {
        const float p0 = xb[n], p1 = xb[N + n], p2 = xb[2 * N + n];
        // bmm(x^T, R^T): out_j = sum_k x_k R[j][k], k ascending
        yb[n] = fmaf(p2, r[2], fmaf(p1, r[1], p0 * r[0]));
        yb[N + n] = fmaf(p2, r[5], fmaf(p1, r[4], p0 * r[3]));
        yb[2 * N + n] = fmaf(p2, r[8], fmaf(p1, r[7], p0 * r[6]));
    }
}

__global__ void __launch_bounds__(256) e3_apply_kernel(const float *__restrict__ loc, const float *__restrict__ vel,
                                                       const float *__restrict__ R, const float *__restrict__ t,
                                                       float *__restrict__ lc, float *__restrict__ vc, int M) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float r[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) r[i] = R[(size_t)m * 9 + i];
    const float l0 = loc[3 * m], l1 = loc[3 * m + 1], l2 = loc[3 * m + 2];
    const float v0 = vel[3 * m], v1 = vel[3 * m + 1], v2 = vel[3 * m + 2];
    const float t0 = t[3 * m], t1 = t[3 * m + 1], t2 = t[3 * m + 2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        // row-vector times R^T: out_j = sum_k in_k R[j][k]; loc R^T - t R^T as two separate products
        const float a = fmaf(l2, r[3 * j + 2], fmaf(l1, r[3 * j + 1], l0 * r[3 * j]));
        const float b = fmaf(t2, r[3 * j + 2], fmaf(t1, r[3 * j + 1], t0 * r[3 * j]));
        lc[3 * m + j] = a - b;
        vc[3 * m + j] = fmaf(v2, r[3 * j + 2], fmaf(v1, r[3 * j + 1], v0 * r[3 * j]));
    }
}

__global__ void __launch_bounds__(256) e3_invert_kernel(const float *__restrict__ x, const float *__restrict__ R,
                                                        const float *__restrict__ t, float *__restrict__ y, int M) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float x0 = x[3 * m], x1 = x[3 * m + 1], x2 = x[3 * m + 2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        // x R + t: out_j = sum_k x_k R[k][j]
        const float a = fmaf(x2, R[(size_t)m * 9 + 6 + j], fmaf(x1, R[(size_t)m * 9 + 3 + j], x0 * R[(size_t)m * 9 + j]));
        y[3 * m + j] = a + t[3 * m + j];
    }
}

// ONE launch for any size: up to PS_BLOCKS blocks leave fp64 partial sums, the last block to finish (ticket) adds them in block
// order (run-to-run deterministic) and writes the five floats.  Tickets / partials live in per-device globals, one slot per
// launch in flight (the host hands out slots round-robin), so launches on different streams do not share state and nothing is
// allocated per call (round 1: a single CTA with a 64-bit modulo per element up to 256 Ki entries -- 265 us for 5 000 n-body
// systems -- and cudaMallocAsync + a second launch above that).
constexpr int PS_BLOCKS = 64, PS_SLOTS = 32;
__device__ unsigned int g_ps_ticket[PS_SLOTS];
__device__ double g_ps_partial[PS_SLOTS][PS_BLOCKS];

__global__ void __launch_bounds__(256) prior_stats_continuous_kernel(const float *__restrict__ R, long long total, int d,
                                                                     float count, float *__restrict__ stats, int slot) {
    double acc = 0.0;
    const int dd = d * d;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    // position inside the d x d matrix, advanced without a division per element
    int e = (int)(i % dd);
    const int step = (int)(stride % dd);
    for (; i < total; i += stride) {
        const float diff = R[i] - ((e / d == e % d) ? 1.f : 0.f);
        acc += (double)(diff * diff);
        e += step;
        if (e >= dd) e -= dd;
    }
    __shared__ double s[8];
    __shared__ unsigned int last;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) c += s[w];
        g_ps_partial[slot][blockIdx.x] = c;
        __threadfence();
        last = atomicAdd(&g_ps_ticket[slot], 1u) == gridDim.x - 1 ? 1u : 0u;
        if (last) {
            __threadfence();
            double t = 0.0;
            for (unsigned int k = 0; k < gridDim.x; ++k) t += __ldcg(&g_ps_partial[slot][k]);
            stats[0] = (float)t;
            stats[1] = count;
            stats[2] = 0.f;
            stats[3] = (float)(t / (double)count);
            stats[4] = 1.f - stats[3];
            g_ps_ticket[slot] = 0u;
        }
    }
}

static inline unsigned grid_for(long long total, int threads, int max_blocks) {
    long long b = (total + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (unsigned)b;
}

}  // namespace eqb

using namespace eqb;

extern "C" int eqb_abi_version(void) { return EQB_ABI_VERSION; }
extern "C" const char *eqb_last_error(void) { return g_err; }

namespace eqb {
int launch_crop_resize_tma(const float *x, float *y, float *amax, int B, int C, int H, int W, int top, int left, int ch,
                           int cw, int oh, int ow, cudaStream_t st, int *handled);   // crop_resize_tma.cu
int tc_absmax(const float *x, int B, size_t n_per_image, float *absmax, cudaStream_t st);   // gconv_stack_tc.cu
}

static int crop_resize(const float *x, float *y, float *amax, int B, int C, int H, int W, int top, int left, int crop_h,
                       int crop_w, int out_h, int out_w, void *stream) {
    EQB_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0, "eqb_crop_resize_aa: bad shape");
    EQB_REQUIRE(top >= 0 && left >= 0 && crop_h > 0 && crop_w > 0 && top + crop_h <= H && left + crop_w <= W,
                "eqb_crop_resize_aa: crop window [%d+%d, %d+%d] outside %dx%d", top, crop_h, left, crop_w, H, W);
    const float sy = (float)crop_h / (float)out_h, sx = (float)crop_w / (float)out_w;
    const float sup_y = sy >= 1.f ? sy : 1.f, sup_x = sx >= 1.f ? sx : 1.f;
    EQB_UNSUPPORTED((int)(2 * sup_y + 2) > AA_MAX_TAPS || (int)(2 * sup_x + 2) > AA_MAX_TAPS,
                    "eqb_crop_resize_aa: down-scale factor above %d not supported", (AA_MAX_TAPS - 2) / 2);
    if (B == 0) return 0;
    EQB_REQUIRE(x && y, "eqb_crop_resize_aa: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (amax) EQB_CUDA(cudaMemsetAsync(amax, 0, (size_t)B * sizeof(float), st));
    // TMA-staged kernel (crop_resize_tma.cu) when the tensor meets the TMA layout rules; same bits as the scalar kernel
    if (!getenv("EQB_RESIZE_SCALAR")) {
        int handled = 0;
        const int e = launch_crop_resize_tma(x, y, amax, B, C, H, W, top, left, crop_h, crop_w, out_h, out_w, st, &handled);
        if (e) return e;
        if (handled) return 0;
    }
    const int bands = (out_h + RS_ROWS - 1) / RS_ROWS, chunks = (out_w + RS_COLS - 1) / RS_COLS;
    const long long blocks = (long long)B * C * bands * chunks;
    EQB_REQUIRE(blocks < (1LL << 31), "eqb_crop_resize_aa: grid too large");
    // input rows one band can touch: RS_ROWS output rows apart by sy, plus the filter support on both sides
    const int strip_rows = (int)((RS_ROWS - 1) * sy + 2 * sup_y + 3);
    const size_t smem = (size_t)(RS_COLS + RS_ROWS) * sizeof(AaAxis) + (size_t)strip_rows * RS_COLS * sizeof(float);
    EQB_UNSUPPORTED(smem > 200 * 1024, "eqb_crop_resize_aa: down-scale factor too large for the shared-memory strip");
    const int bx = 32 * ((std::min(out_w, RS_COLS) + 31) / 32);
    const dim3 block(bx, 256 / bx);
    const bool narrow = (int)(2 * sup_x + 2) <= 6;
    auto kern = narrow ? crop_resize_aa_kernel<6> : crop_resize_aa_kernel<AA_MAX_TAPS>;
    if (smem > 48 * 1024) EQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, block, smem, st>>>(x, y, H, W, top, left, crop_h, crop_w, out_h, out_w, sy, sx, bands, chunks,
                                                strip_rows);
    int e = finish_launch("eqb_crop_resize_aa");
    if (e || !amax) return e;
    return tc_absmax(y, B, (size_t)C * out_h * out_w, amax, st);
}

extern "C" int eqb_crop_resize_aa(const float *x, float *y, int B, int C, int H, int W, int top, int left, int crop_h,
                                  int crop_w, int out_h, int out_w, void *stream) {
    EQB_NVTX_RANGE();
    return crop_resize(x, y, nullptr, B, C, H, W, top, left, crop_h, crop_w, out_h, out_w, stream);
}

extern "C" int eqb_crop_resize_aa_absmax(const float *x, float *y, float *y_absmax, int B, int C, int H, int W, int top,
                                         int left, int crop_h, int crop_w, int out_h, int out_w, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(y_absmax || B == 0, "eqb_crop_resize_aa_absmax: null absmax");
    return crop_resize(x, y, y_absmax, B, C, H, W, top, left, crop_h, crop_w, out_h, out_w, stream);
}

namespace eqb {
int launch_lift_orbit(const float *w, float *out, int cout, int cin, int k, int N, int reflect, long long sn,
                      long long sk, cudaStream_t st) {
    const int G = N * (reflect ? 2 : 1);
    const long long total = (long long)cout * G * cin * k * k;
    lift_orbit_kernel<<<grid_for(total, 256, 4096), 256, 0, st>>>(w, out, cout, cin, k, N, G, sn, sk);
    return finish_launch("lift_orbit");
}
int launch_regular_orbit(const float *w, float *out, int cout, int cin, int k, int N, int reflect, long long sn,
                         long long sk, cudaStream_t st) {
    const int G = N * (reflect ? 2 : 1);
    const long long total = (long long)cout * G * cin * G * k * k;
    regular_orbit_kernel<<<grid_for(total, 256, 4096), 256, 0, st>>>(w, out, cout, cin, k, N, G, sn, sk);
    return finish_launch("regular_orbit");
}
}  // namespace eqb

extern "C" int eqb_lift_filter_orbit(const float *w, float *orbit, int cout, int cin, int k, int num_rotations,
                                     int reflect, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(cout > 0 && cin > 0 && k > 0 && num_rotations > 0 && w && orbit, "eqb_lift_filter_orbit: bad argument");
    return launch_lift_orbit(w, orbit, cout, cin, k, num_rotations, reflect, (long long)cin * k * k, 1,
                             (cudaStream_t)stream);
}

extern "C" int eqb_regular_filter_orbit(const float *w, float *orbit, int cout, int cin, int k, int num_rotations,
                                        int reflect, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(cout > 0 && cin > 0 && k > 0 && num_rotations > 0 && w && orbit,
                "eqb_regular_filter_orbit: bad argument");
    const int G = num_rotations * (reflect ? 2 : 1);
    return launch_regular_orbit(w, orbit, cout, cin, k, num_rotations, reflect, (long long)cin * G * k * k, 1,
                                (cudaStream_t)stream);
}

extern "C" int eqb_group_pool_select(const float *act, int B, int num_rotations, int reflect, int32_t *idx,
                                     float *rotation, float *reflection, float *onehot, float *stats, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && num_rotations > 0, "eqb_group_pool_select: bad shape");
    EQB_REQUIRE(stats && (B == 0 || (act && idx && rotation)), "eqb_group_pool_select: null pointer");
    const int G = num_rotations * (reflect ? 2 : 1);
    // one CTA up to 8192 samples (32 per thread): the kernel is latency-bound either way, and a single CTA finishes the
    // statistic itself - no stream-ordered scratch allocation and no second launch on the path of every step (the
    // cudaMallocAsync / cudaFreeAsync pair cost 8 us on some boxes of the pool and 60 us on others)
    const unsigned blocks = B <= 8192 ? 1u : grid_for(B, 256, SEL_BLOCKS);
    cudaStream_t st = (cudaStream_t)stream;
    static unsigned int next_slot = 0;
    const int slot = (int)(next_slot++ % SEL_SLOTS);
    group_pool_select_kernel<<<blocks, 256, 0, st>>>(act, B, num_rotations, G, idx, rotation,
                                                     reflect ? reflection : nullptr, onehot, stats, slot);
    return finish_launch("eqb_group_pool_select");
}

extern "C" int eqb_cosine_group_activations(const float *vec, const float *ref, float *act, int B, int num_group, int V,
                                            void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && num_group > 0 && V > 0, "eqb_cosine_group_activations: bad shape");
    if (B == 0) return 0;
    EQB_REQUIRE(vec && ref && act, "eqb_cosine_group_activations: null pointer");
    const int rows = B * num_group;
    cosine_act_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(vec, ref, act, B, num_group, V);
    return finish_launch("eqb_cosine_group_activations");
}

extern "C" int eqb_cosine_group_activations_backward(const float *vec, const float *ref, const float *dact, float *dvec,
                                                     float *dref, int B, int num_group, int V, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && num_group > 0 && V > 0, "eqb_cosine_group_activations_backward: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (dref) EQB_CUDA(cudaMemsetAsync(dref, 0, (size_t)V * sizeof(float), st));
    if (B == 0) return 0;
    EQB_REQUIRE(vec && ref && dact && (dvec || dref), "eqb_cosine_group_activations_backward: null pointer");
    const int rows = B * num_group;
    cosine_act_backward_kernel<<<(rows + 7) / 8, 256, 0, st>>>(vec, ref, dact, dvec, dref, B, num_group, V);
    return finish_launch("eqb_cosine_group_activations_backward");
}

extern "C" int eqb_gram_schmidt3(const float *v, float *R, int B, int modified, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0, "eqb_gram_schmidt3: bad batch");
    if (B == 0) return 0;
    EQB_REQUIRE(v && R, "eqb_gram_schmidt3: null pointer");
    gram_schmidt3_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(v, R, B, modified);
    return finish_launch("eqb_gram_schmidt3");
}

extern "C" int eqb_so3_apply(const float *x, const float *R, float *y, int B, int N, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && N >= 0, "eqb_so3_apply: bad shape");
    if (B == 0 || N == 0) return 0;
    EQB_REQUIRE(x && R && y, "eqb_so3_apply: null pointer");
    int chunks = (N + 1023) / 1024;  // 4 points per thread
    if (chunks < 1) chunks = 1;
    EQB_REQUIRE((long long)B * chunks < (1LL << 31), "eqb_so3_apply: grid too large");
    so3_apply_kernel<<<(unsigned)(B * chunks), 256, 0, (cudaStream_t)stream>>>(x, R, y, B, N, chunks);
    return finish_launch("eqb_so3_apply");
}

extern "C" int eqb_e3_apply(const float *loc, const float *vel, const float *R, const float *t, float *loc_c,
                            float *vel_c, int M, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(M >= 0, "eqb_e3_apply: bad row count");
    if (M == 0) return 0;
    EQB_REQUIRE(loc && vel && R && t && loc_c && vel_c, "eqb_e3_apply: null pointer");
    e3_apply_kernel<<<(M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(loc, vel, R, t, loc_c, vel_c, M);
    return finish_launch("eqb_e3_apply");
}

extern "C" int eqb_e3_invert(const float *x, const float *R, const float *t, float *y, int M, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(M >= 0, "eqb_e3_invert: bad row count");
    if (M == 0) return 0;
    EQB_REQUIRE(x && R && t && y, "eqb_e3_invert: null pointer");
    e3_invert_kernel<<<(M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, R, t, y, M);
    return finish_launch("eqb_e3_invert");
}

extern "C" int eqb_prior_stats_continuous(const float *R, int B, int d, float *stats, void *stream) {
    EQB_NVTX_RANGE();
    EQB_REQUIRE(B >= 0 && d > 0 && stats, "eqb_prior_stats_continuous: bad argument");
    EQB_REQUIRE(B == 0 || R, "eqb_prior_stats_continuous: null pointer");
    const long long total = (long long)B * d * d;
    cudaStream_t st = (cudaStream_t)stream;
    static unsigned int next_slot = 0;
    const int slot = (int)(next_slot++ % PS_SLOTS);
    long long blocks = (total + 2047) / 2048;        // >= 8 entries per thread before another block pays off
    if (blocks < 1) blocks = 1;
    if (blocks > PS_BLOCKS) blocks = PS_BLOCKS;
    prior_stats_continuous_kernel<<<(unsigned)blocks, 256, 0, st>>>(R, total, d, (float)total, stats, slot);
    return finish_launch("eqb_prior_stats_continuous");
}
