// Probe: does a K-major SWIZZLE_64B UMMA A-descriptor accept a start address shifted by whole 64-byte rows inside a
// larger swizzled buffer (as a TMA-written halo patch would need), and with which base_offset?  The A buffer holds
// (16 groups x PITCH rows) x 64 B, written with the swizzle of the ABSOLUTE row index (what TMA does); the MMA reads the
// 128 rows { g * PITCH + shift + r : g < 16, r < 8 } via SBO = PITCH * 64.
//   umma_shift_probe <pitch_rows> <shift_rows> <base_offset_mode 0|1>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) probe(const float *A, const float *B, float *D, int pitch, int shift, int bo_mode) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const uint32_t base = smem_u32(sm);
    const int N = 64, K = 32, rowsA = 16 * pitch + 16;
    const uint32_t a_off = 0, b_off = ((rowsA * 64 + 1023) / 1024) * 1024, bar = base + b_off + N * 64, slot = bar + 8;
    const int tid = threadIdx.x;
    auto phys = [&](int row, int k) { return (uint32_t)(row * 64 + ((((k % 32) / 8) ^ ((row >> 1) & 3)) << 4) + (k % 8) * 2); };
    // logical A row m = g*8 + r lives at buffer row g*pitch + shift + r
    for (int i = tid; i < rowsA * K; i += 128) *(__half *)(sm + a_off + phys(i / K, i % K)) = __float2half(0.f);
    __syncthreads();
    for (int i = tid; i < 128 * K; i += 128) {
        const int m = i / K, k = i % K, row = (m / 8) * pitch + shift + (m % 8);
        *(__half *)(sm + a_off + phys(row, k)) = __float2half(A[i]);
    }
    for (int i = tid; i < N * K; i += 128) *(__half *)(sm + b_off + phys(i / K, i % K)) = __float2half(B[i]);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *(volatile uint32_t *)(sm + (slot - base));
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int j = 0; j < 2; ++j) {
            const uint32_t a_addr = base + a_off + shift * 64 + 32 * j, b_addr = base + b_off + 32 * j;
            const uint64_t bo = bo_mode ? (uint64_t)((a_addr >> 7) & 7) : 0;
            const uint64_t da = (uint64_t)((a_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((pitch * 64) >> 4) << 32) |
                                ((uint64_t)1 << 46) | (bo << 49) | ((uint64_t)4 << 61);
            const uint64_t db = (uint64_t)((b_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) |
                                ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
            const uint32_t acc = j != 0;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra Dn;\nbra W;\nDn:\n}\n" ::"r"(bar), "r"(0) : "memory");
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
        for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * N + c * 32 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}
int main(int argc, char **argv) {
    const int pitch = atoi(argv[1]), shift = atoi(argv[2]), bo = atoi(argv[3]);
    const int N = 64, K = 32;
    float *hA = (float *)malloc(128 * K * 4), *hB = (float *)malloc(N * K * 4), *hD = (float *)malloc(128 * N * 4);
    srand(1);
    for (int i = 0; i < 128 * K; ++i) hA[i] = (float)(rand() % 17 - 8) / 8.f;
    for (int i = 0; i < N * K; ++i) hB[i] = (float)(rand() % 17 - 8) / 16.f;
    float *A, *B, *D;
    cudaMalloc(&A, 128 * K * 4); cudaMalloc(&B, N * K * 4); cudaMalloc(&D, 128 * N * 4);
    cudaMemcpy(A, hA, 128 * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(B, hB, N * K * 4, cudaMemcpyHostToDevice);
    cudaMemset(D, 0, 128 * N * 4);
    const size_t smem = (size_t)(16 * pitch + 16) * 64 + 1024 + N * 64 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(A, B, D, pitch, shift, bo);
    cudaError_t e = cudaGetLastError();
    if (!e) e = cudaDeviceSynchronize();
    printf("pitch=%d shift=%d base_offset_mode=%d: %s", pitch, shift, bo, cudaGetErrorString(e));
    if (e) { printf("\n"); return 1; }
    cudaMemcpy(hD, D, 128 * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)hA[m * K + k] * hB[n * K + k];
        const double err = fabs(ref - hD[m * N + n]); if (err > maxerr) maxerr = err; if (err > 1e-4) ++bad;
    }
    printf("  max err %.3g  bad %d / %d\n", maxerr, bad, 128 * N);
    return 0;
}
