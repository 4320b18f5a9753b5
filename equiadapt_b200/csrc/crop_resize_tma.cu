// a3 on the TMA unit: CenterCrop + antialiased bilinear resize (discrete_group.py:174-188), persistent CTAs over (plane,
// band) items, plus the per-IMAGE max |x_pre| the conv stack scales its fp16 operand split with.
//
// The scalar kernel in small_ops.cu is latency-bound on its per-tap global loads (0.33 of the HBM copy peak, r1 profile:
// every thread issues its taps as 4-byte LDGs and waits for them).  Here
//   * ONE cp.async.bulk.tensor.3d per item lands the band's source rows (crop_w x rows box) in shared memory; a two-stage
//     ring per persistent CTA (two CTAs per SM) keeps four boxes in flight per SM behind the one being filtered;
//   * the separable triangle filter (ATen _upsample_bilinear2d_aa: horizontal pass, then vertical, same per-output FMA
//     order as small_ops.cu, so both kernels return the same bits) runs from shared memory;
//   * each CTA folds max |y| of its outputs into amax[image] with one atomicMax (non-negative floats order like their bit
//     patterns), so no separate pass over x_pre is needed and a sample's scale never depends on its batch-mates.
// Algorithmic traffic: the crop window once (+ the 4-row overlap of neighbouring bands, served by L2) and the resized
// plane once: (388 800 + 110 592) B per 3x224x224 image at crop 180 / resize 96.
#include <cuda.h>

#include <algorithm>

#include "resample.cuh"

namespace eqb {

namespace rs2 {

constexpr int MAX_TAPS = 16;
constexpr int BAND = 24;          // output rows per CTA
constexpr int THREADS2 = 384;

struct Axis {       // 17 words: odd stride -> conflict-free per-lane reads of the weights
    int lo_n;       // lo | n << 20
    float w[MAX_TAPS];
};

// identical to aa_axis in small_ops.cu (ATen's tap rule)
__device__ __forceinline__ void axis(int i, int in_size, float scale, Axis &ax) {
    const float support = scale >= 1.f ? scale : 1.f;
    const float invscale = scale >= 1.f ? 1.f / scale : 1.f;
    const float center = scale * (i + 0.5f);
    const int lo = max((int)(center - support + 0.5f), 0);
    const int hi = min((int)(center + support + 0.5f), in_size);
    const int n = min(hi - lo, MAX_TAPS);
    float tot = 0.f;
    for (int j = 0; j < n; ++j) {
        const float v = fmaxf(1.f - fabsf((j + lo - center + 0.5f) * invscale), 0.f);
        ax.w[j] = v;
        tot += v;
    }
    for (int j = 0; j < n; ++j) ax.w[j] = tot != 0.f ? ax.w[j] / tot : 0.f;
    for (int j = n > 0 ? n : 0; j < MAX_TAPS; ++j) ax.w[j] = 0.f;
    ax.lo_n = lo | (max(n, 0) << 20);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

struct Args {
    float *y;
    float *amax;            // (B) per-image max |y|, zeroed by the caller (may be null)
    int C, top, box_x, dx, ch, cw, oh, ow, bands, box_w, box_rows;   // box_x = left rounded down to 16 bytes, dx = left - box_x
    int items;              // planes * bands
    float sy, sx;
};

constexpr int STAGES = 2;

// Persistent CTA: work items (plane, band of BAND output rows) round-robin; a two-stage ring of source boxes keeps TWO
// cp.async.bulk.tensor loads in flight per CTA while the threads filter the current one.  blockDim = (ow rounded up to 32,
// 256 / that): thread x owns output column x and keeps its horizontal weights in registers.
template <int MAXT>
__global__ void __launch_bounds__(THREADS2) crop_resize_tma_kernel(const __grid_constant__ CUtensorMap map, const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    // [box 0][box 1] (each box_rows x box_w floats, 128-byte aligned) [strip: box_rows x ow][tx: ow][ty: oh][mbarriers][red]
    const size_t box_floats = (((size_t)a.box_rows * a.box_w * 4 + 127) & ~(size_t)127) / 4;
    float *box0 = reinterpret_cast<float *>(smem);
    float *strip = box0 + STAGES * box_floats;
    Axis *tx = reinterpret_cast<Axis *>(strip + (size_t)a.box_rows * a.ow);
    Axis *ty = tx + a.ow;
    // (the tables are 68-byte records: round up to the 16-byte boundary the mbarriers need)
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(
        smem + ((reinterpret_cast<unsigned char *>(ty + a.oh) - smem + 15) & ~(size_t)15));
    float *red = reinterpret_cast<float *>(bars + STAGES);
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthreads = blockDim.x * blockDim.y;
    const uint32_t box_bytes = (uint32_t)(a.box_rows * a.box_w * 4);

    // tap tables of the whole plane, once per CTA (they are the same for every item)
    for (int t = tid; t < a.ow + a.oh; t += nthreads) {
        if (t < a.ow) axis(t, a.cw, a.sx, tx[t]);
        else axis(t - a.ow, a.ch, a.sy, ty[t - a.ow]);
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + s)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int item, int stage) {      // (thread 0) source rows of `item` -> box[stage]
        const int band = item % a.bands, plane = item / a.bands;
        const int row0 = ty[band * BAND].lo_n & 0xfffff;
        const uint32_t bar_s = smem_u32(bars + stage);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(box_bytes) : "memory");
        // columns / rows past the image are zero-filled by the TMA unit and never used as taps
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(smem_u32(box0 + stage * box_floats)), "l"((uint64_t)&map), "r"(bar_s), "r"(a.box_x), "r"(a.top + row0), "r"(plane)
            : "memory");
    };
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            const int item = (int)blockIdx.x + s * (int)gridDim.x;
            if (item < a.items) issue(item, s);
        }
    }
    const int ox = threadIdx.x;
    const bool col_ok = ox < a.ow;
    // this thread's column: horizontal taps in registers
    int xlo = 0, xn = 0;
    float wx[MAXT];
#pragma unroll
    for (int i = 0; i < MAXT; ++i) wx[i] = 0.f;
    if (col_ok) {
        xlo = tx[ox].lo_n & 0xfffff;
        xn = tx[ox].lo_n >> 20;
#pragma unroll
        for (int i = 0; i < MAXT; ++i) wx[i] = tx[ox].w[i];
    }
    int k = 0;
    for (int item = blockIdx.x; item < a.items; item += gridDim.x, ++k) {
        const int stage = k % STAGES;
        const uint32_t parity = (uint32_t)(k / STAGES) & 1u;
        const int band = item % a.bands, plane = item / a.bands;
        const int oy0 = band * BAND, nrows = min(BAND, a.oh - oy0);
        const int row_lo = ty[oy0].lo_n & 0xfffff;
        const int row_hi = (ty[oy0 + nrows - 1].lo_n & 0xfffff) + (ty[oy0 + nrows - 1].lo_n >> 20);
        const int nr = min(row_hi - row_lo, a.box_rows);
        const float *box = box0 + stage * box_floats;
        {   // wait for the box
            const uint32_t bar_s = smem_u32(bars + stage);
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n"
                    ".reg .pred p;\n"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                    "selp.u32 %0, 1, 0, p;\n"
                    "}\n"
                    : "=r"(done)
                    : "r"(bar_s), "r"(parity)
                    : "memory");
            }
        }
        // horizontal pass (same FMA order per output as the scalar kernel): four rows per iteration, 32-bit shared-memory
        // addresses, taps predicated on this thread's tap count (no branches)
        const uint32_t strip_s = smem_u32(strip) + 4u * (uint32_t)ox;
        const uint32_t ow4 = 4u * (uint32_t)a.ow;
        if (col_ok) {
            const uint32_t cp = smem_u32(box) + 4u * (uint32_t)(a.dx + xlo), pitch = 4u * (uint32_t)a.box_w;
            const int by = blockDim.y;
            int r = threadIdx.y;
            for (; r + 3 * by < nr; r += 4 * by) {
                float h[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < MAXT; ++i) {
                    if (i < xn) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) h[q] = fmaf(lds(cp + (uint32_t)(r + q * by) * pitch + 4u * i), wx[i], h[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) sts(strip_s + (uint32_t)(r + q * by) * ow4, h[q]);
            }
            for (; r < nr; r += by) {
                float h = 0.f;
#pragma unroll
                for (int i = 0; i < MAXT; ++i)
                    if (i < xn) h = fmaf(lds(cp + (uint32_t)r * pitch + 4u * i), wx[i], h);
                sts(strip_s + (uint32_t)r * ow4, h);
            }
        }
        __syncthreads();               // strip complete; nobody reads this box any more
        if (tid == 0) {
            const int next = item + STAGES * (int)gridDim.x;
            if (next < a.items) issue(next, stage);
        }
        // vertical pass: taps predicated on the row's tap count, weights of the row in registers
        float m = 0.f;
        if (col_ok) {
            // (running addresses: the pass is bound by its instruction count -- 52 per output with the multiplies in the loop)
            float *yp = a.y + ((size_t)plane * a.oh + oy0 + threadIdx.y) * a.ow + ox;
            const size_t ystep = (size_t)blockDim.y * a.ow;
            const uint32_t ty_s = smem_u32(ty + oy0);
            for (int oy = threadIdx.y; oy < nrows; oy += blockDim.y, yp += ystep) {
                const uint32_t ay_s = ty_s + (uint32_t)oy * (uint32_t)sizeof(Axis);
                const int lo_n = __float_as_int(lds(ay_s));
                const int lo = (lo_n & 0xfffff) - row_lo, n = lo_n >> 20;
                uint32_t sp = strip_s + (uint32_t)lo * ow4;
                float acc = 0.f;
                if (n <= MAXT) {
#pragma unroll
                    for (int j = 0; j < MAXT; ++j) {
                        if (j < n) acc = fmaf(lds(sp), lds(ay_s + 4u + 4u * j), acc);
                        sp += ow4;
                    }
                } else {
                    for (int j = 0; j < n; ++j, sp += ow4) acc = fmaf(lds(sp), lds(ay_s + 4u + 4u * j), acc);
                }
                *yp = acc;
                m = fmaxf(m, fabsf(acc));
                if (!(acc == acc)) m = __int_as_float(0x7f800000);   // NaN poisons only its own image's scale: +inf
            }
        }
        if (a.amax) {
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if ((tid & 31) == 0) red[tid >> 5] = m;
        }
        __syncthreads();               // strip free for the next item; red complete
        if (a.amax && tid == 0) {
            for (int w = 1; w < (nthreads + 31) / 32; ++w) m = fmaxf(m, red[w]);
            atomicMax(reinterpret_cast<unsigned int *>(a.amax + plane / a.C), __float_as_uint(m));
        }
    }
}

}  // namespace rs2

// -> 0 launched, *handled = 1; *handled = 0 when the tensor does not meet the TMA rules (caller takes the scalar kernel)
int launch_crop_resize_tma(const float *x, float *y, float *amax, int B, int C, int H, int W, int top, int left, int ch,
                           int cw, int oh, int ow, cudaStream_t st, int *handled) {
    *handled = 0;
    const float sy = (float)ch / (float)oh, sx = (float)cw / (float)ow;
    const float sup_y = sy >= 1.f ? sy : 1.f, sup_x = sx >= 1.f ? sx : 1.f;
    const int box_x = left & ~3, dx = left - box_x;                        // a box must START on a 16-byte boundary (tools/tma_probe.cu)
    const int box_w = (cw + dx + 3) & ~3;                                  // inner box extent: a multiple of 16 bytes
    const int box_rows = (int)((rs2::BAND - 1) * sy + 2 * sup_y + 3);      // input rows one band can touch
    const long long planes = (long long)B * C;
    if (((uintptr_t)x & 15) != 0 || (W & 3) != 0 || box_w > 256 || box_rows > 256 || planes <= 0 ||
        planes >= (1LL << 31) || (int)(2 * sup_x + 2) > rs2::MAX_TAPS || (int)(2 * sup_y + 2) > rs2::MAX_TAPS)
        return 0;
    if (ow > 128) return 0;
    const int bands = (oh + rs2::BAND - 1) / rs2::BAND;
    const long long items = planes * bands;
    if (items >= (1LL << 31)) return 0;
    const size_t box_bytes = (((size_t)box_rows * box_w * 4) + 127) & ~(size_t)127;
    size_t smem = rs2::STAGES * box_bytes + (size_t)box_rows * ow * sizeof(float) + (size_t)(ow + oh) * sizeof(rs2::Axis);
    smem = (smem + 15) & ~(size_t)15;
    smem += 8 * rs2::STAGES + 64;
    if (smem > 110 * 1024) return 0;                                       // two persistent CTAs per SM
    CUtensorMap map;
    int e = make_plane_map(&map, x, W, H, planes, box_w, box_rows, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (e) return e;
    rs2::Args a{};
    a.y = y; a.amax = amax; a.C = C; a.top = top; a.box_x = box_x; a.dx = dx; a.ch = ch; a.cw = cw; a.oh = oh; a.ow = ow;
    a.bands = bands; a.box_w = box_w; a.box_rows = box_rows; a.items = (int)items; a.sy = sy; a.sx = sx;
    const bool narrow = (int)(2 * sup_x + 2) <= 6;
    auto kern = narrow ? rs2::crop_resize_tma_kernel<6> : rs2::crop_resize_tma_kernel<rs2::MAX_TAPS>;
    EQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int bx = 32 * ((ow + 31) / 32);
    const dim3 block(bx, rs2::THREADS2 / bx);
    const long long grid = std::min<long long>(items, 2LL * num_sms());
    kern<<<(unsigned)grid, block, smem, st>>>(map, a);
    *handled = 1;
    return finish_launch("crop_resize_tma_kernel");
}

}  // namespace eqb
