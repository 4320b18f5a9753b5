// Probe: one 256 x N x K fp16 UMMA tile issued by a CTA PAIR (tcgen05.mma.cta_group::2, cluster of 2): checks
//   - tcgen05.alloc/dealloc.cta_group::2 issued by one warp of EACH CTA,
//   - operand placement: CTA r supplies A rows [128 r, 128 r + 128) and B rows [N/2 r, N/2 r + N/2) from its own
//     shared memory at the SAME offsets, K-major, SWIZZLE_64B atoms of 32 fp16,
//   - accumulator placement: CTA r's TMEM lanes hold D rows [128 r, +128), all N columns,
//   - tcgen05.commit.cta_group::2 ... multicast::cluster arriving on the barrier at the same offset in both CTAs.
//   umma_pair_probe <N> <K>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc64(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)4 << 61);
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    probe(const float *A, const float *B, float *D, int N, int K, int *status) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    unsigned char *sm = raw + (base - smem_u32(raw));
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int atoms = K / 32, NH = N / 2;
    const uint32_t a_off = 0, a_atom = 128 * 64, b_off = atoms * a_atom, b_atom = NH * 64;
    const uint32_t bar = base + b_off + atoms * b_atom, slot = bar + 8;
    const int tid = threadIdx.x;
    auto phys = [&](int row, int k) {
        const int chunk = (k % 32) / 8, within = (k % 8) * 2;
        return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4) + within);
    };
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        *(__half *)(sm + a_off + (k / 32) * a_atom + phys(r, k)) = __float2half(A[(size_t)(128 * rank + r) * K + k]);
    }
    for (int i = tid; i < NH * K; i += 128) {
        const int r = i / K, k = i % K;
        *(__half *)(sm + b_off + (k / 32) * b_atom + phys(r, k)) = __float2half(B[(size_t)(NH * rank + r) * K + k]);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *(volatile uint32_t *)(sm + (slot - base));
    if (tid == 0) status[rank] = (int)tmem;
    if (rank == 0 && tid == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        for (int ks = 0; ks < K / 16; ++ks) {
            const int atom = ks / 2, j = ks % 2;
            const uint64_t da = umma_desc64(base + a_off + atom * a_atom + 32 * j), db = umma_desc64(base + b_off + atom * b_atom + 32 * j);
            const uint32_t acc = ks != 0;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
    }
    // bounded wait on the local barrier
    {
        uint32_t ok = 0;
        const long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
            if (!ok && clock64() - t0 > 2000000000LL) {
                if (tid == 0) status[2 + rank] = -1;
                break;
            }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = tid >> 5, lane = tid & 31;
    for (int c = 0; c < N / 32; ++c) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c * 32;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) D[(size_t)(128 * rank + warp * 32 + lane) * N + c * 32 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}
int main(int argc, char **argv) {
    const int N = atoi(argv[1]), K = atoi(argv[2]);
    float *hA = (float *)malloc(256 * K * 4), *hB = (float *)malloc(N * K * 4), *hD = (float *)malloc(256 * N * 4);
    srand(1);
    for (int i = 0; i < 256 * K; ++i) hA[i] = (float)(rand() % 17 - 8) / 8.f;
    for (int i = 0; i < N * K; ++i) hB[i] = (float)(rand() % 17 - 8) / 16.f;
    float *A, *B, *D; int *st;
    cudaMalloc(&A, 256 * K * 4); cudaMalloc(&B, N * K * 4); cudaMalloc(&D, 256 * N * 4); cudaMalloc(&st, 16);
    cudaMemcpy(A, hA, 256 * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(B, hB, N * K * 4, cudaMemcpyHostToDevice);
    cudaMemset(D, 0, 256 * N * 4); cudaMemset(st, 0, 16);
    const size_t smem = (size_t)(128 + N / 2) * K * 2 + 2048;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<2, 128, smem>>>(A, B, D, N, K, st);
    cudaError_t e = cudaGetLastError();
    if (!e) e = cudaDeviceSynchronize();
    printf("pair N=%d K=%d smem=%zu: %s", N, K, smem, cudaGetErrorString(e));
    if (e) { printf("\n"); return 1; }
    int hs[4]; cudaMemcpy(hs, st, 16, cudaMemcpyDeviceToHost);
    cudaMemcpy(hD, D, 256 * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0, bad_lo = 0;
    for (int m = 0; m < 256; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)hA[m * K + k] * hB[n * K + k];
        const double err = fabs(ref - hD[m * N + n]); if (err > maxerr) maxerr = err; if (err > 1e-4) { ++bad; if (m < 128) ++bad_lo; }
    }
    printf("  tmem %d %d  timeouts %d %d  max err %.3g  bad %d (rows<128: %d) / %d   D[0][0..3] = %g %g %g %g  D[128][0..1] = %g %g\n",
           hs[0], hs[1], hs[2], hs[3], maxerr, bad, bad_lo, 256 * N, hD[0], hD[1], hD[2], hD[3], hD[128 * N], hD[128 * N + 1]);
    return 0;
}
