// Probe: one 128 x N x K fp16 (kind::f16, fp32 accumulate) UMMA tile with manually written K-major swizzled operands;
// checks the SWIZZLE_32B / 64B / 128B descriptors for 2-byte elements.
//   umma_probe16 <swizzle_bytes 128|64|32> <N> <K>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <cuda_fp16.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, int swz) {
    const uint64_t sbo = 8 * swz, lt = swz == 128 ? 2 : swz == 64 ? 4 : 6;
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | (lt << 61);
}
__global__ void __launch_bounds__(128, 1) probe(const float *A, const float *B, float *D, int N, int K, int swz) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    unsigned char *sm = raw + (base - smem_u32(raw));
    const int row_bytes = swz, kper = swz / 2, atoms = K / kper;   // K values per swizzle row
    const uint32_t a_off = 0, a_atom = 128 * row_bytes, b_off = atoms * a_atom, b_atom = N * row_bytes;
    const uint32_t bar = base + b_off + atoms * b_atom, slot = bar + 8;
    const int tid = threadIdx.x;
    auto phys = [&](int row, int k) {   // byte offset inside an atom
        const int chunk = (k % kper) / 8, within = (k % 8) * 2;
        const int x = swz == 128 ? (row & 7) : swz == 64 ? ((row >> 1) & 3) : ((row >> 2) & 1);
        return (uint32_t)(row * row_bytes + ((chunk ^ x) << 4) + within);
    };
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        *(__half *)(sm + a_off + (k / kper) * a_atom + phys(r, k)) = __float2half(A[i]);
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        *(__half *)(sm + b_off + (k / kper) * b_atom + phys(r, k)) = __float2half(B[i]);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *(volatile uint32_t *)(sm + (slot - base));
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int ks = 0; ks < K / 16; ++ks) {
            const int atom = ks / (kper / 16), j = ks % (kper / 16);
            const uint64_t da = umma_desc(base + a_off + atom * a_atom + 32 * j, swz), db = umma_desc(base + b_off + atom * b_atom + 32 * j, swz);
            const uint32_t acc = ks != 0;
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
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}
int main(int argc, char **argv) {
    const int swz = atoi(argv[1]), N = atoi(argv[2]), K = atoi(argv[3]);
    float *hA = (float *)malloc(128 * K * 4), *hB = (float *)malloc(N * K * 4), *hD = (float *)malloc(128 * N * 4);
    srand(1);
    for (int i = 0; i < 128 * K; ++i) hA[i] = (float)(rand() % 17 - 8) / 8.f;   // exactly representable in fp16
    for (int i = 0; i < N * K; ++i) hB[i] = (float)(rand() % 17 - 8) / 16.f;
    float *A, *B, *D;
    cudaMalloc(&A, 128 * K * 4); cudaMalloc(&B, N * K * 4); cudaMalloc(&D, 128 * N * 4);
    cudaMemcpy(A, hA, 128 * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(B, hB, N * K * 4, cudaMemcpyHostToDevice);
    cudaMemset(D, 0, 128 * N * 4);
    const size_t smem = (size_t)(128 + N) * K * 4 + 2048;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(A, B, D, N, K, swz);
    cudaError_t e = cudaDeviceSynchronize();
    printf("swz=%d N=%d K=%d: %s", swz, N, K, cudaGetErrorString(e));
    if (e) { printf("\n"); return 1; }
    cudaMemcpy(hD, D, 128 * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)hA[m * K + k] * hB[n * K + k];
        const double err = fabs(ref - hD[m * N + n]); if (err > maxerr) maxerr = err; if (err > 1e-4) ++bad;
    }
    printf("  max err %.3g  bad %d / %d   D[0][0..3] = %g %g %g %g\n", maxerr, bad, 128 * N, hD[0], hD[1], hD[2], hD[3]);
    return 0;
}
