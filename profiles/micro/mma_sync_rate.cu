// Throughput and dependent-issue latency of the warp-level tensor-core MMA (mma.sync m16n8k16 bf16, fp32
// accumulate) and of ldmatrix on sm_100a, per SM sub-partition; and the LDGSTS (cp.async 16 B) streaming rate
// of one CTA per SM from L2.  These size the persistent training kernel (train_fused.cu).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_rate mma_sync_rate.cu && ./mma_sync_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int ILP>
__global__ void k_mma(float* out, long long* cyc, int iters) {
    float c[ILP][4];
    uint32_t a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f803f80u};
#pragma unroll
    for (int i = 0; i < ILP; ++i) for (int q = 0; q < 4; ++q) c[i][q] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) mma(c[i], a, b);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) for (int q = 0; q < 4; ++q) s += c[i][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_ldsm(float* out, long long* cyc, int iters, int trans) {
    extern __shared__ __align__(16) unsigned char sm[];
    uint32_t acc = 0;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 432 + (threadIdx.x >> 5) * 16;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t r0, r1, r2, r3;
            if (trans) asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(base + i * 32));
            else asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(base + i * 32));
            acc ^= r0 ^ r1 ^ r2 ^ r3;
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// one CTA per SM streams `bytes` per round from a per-SM-distinct (or shared) L2-resident region with cp.async 16 B
__global__ void k_ldgsts(const unsigned char* src, size_t region, int bytes, int rounds, int shared_src, long long* cyc, float* out) {
    extern __shared__ __align__(16) unsigned char sm[];
    const unsigned char* s = src + (shared_src ? 0 : (size_t)blockIdx.x * region);
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        const unsigned char* p = s + (size_t)(r % 8) * bytes;
        for (int o = threadIdx.x * 16; o < bytes; o += blockDim.x * 16)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(sm + o)), "l"(p + o));
        asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = sm[5];
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float* out; long long* cyc; long long h = 0;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2048;
#define RUN_MMA(ILP, WARPS)                                                                                   \
    k_mma<ILP><<<148, 32 * WARPS>>>(out, cyc, iters); k_mma<ILP><<<148, 32 * WARPS>>>(out, cyc, iters);         \
    cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                  \
    printf("mma.sync m16n8k16 bf16  ILP %2d, %2d warps/SM: %.2f cycles per MMA per warp, %.2f per sub-partition, %.0f flop/clk/SM\n", \
           ILP, WARPS, (double)h / (iters * ILP), (double)h / (iters * ILP * (WARPS < 4 ? 1.0 : WARPS / 4.0)),  \
           4096.0 * iters * ILP * WARPS / (double)h);
    RUN_MMA(1, 4) RUN_MMA(2, 4) RUN_MMA(4, 4) RUN_MMA(8, 4) RUN_MMA(16, 4)
    RUN_MMA(1, 8) RUN_MMA(4, 8) RUN_MMA(8, 8) RUN_MMA(16, 8) RUN_MMA(4, 16) RUN_MMA(8, 16)
    cudaFuncSetAttribute(k_ldsm, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int trans = 0; trans < 2; ++trans)
        for (int warps = 4; warps <= 8; warps += 4) {
            k_ldsm<<<148, 32 * warps, 64 * 1024>>>(out, cyc, iters, trans);
            cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("ldmatrix.x4%s %d warps/SM: %.2f cycles per instruction per sub-partition\n", trans ? ".trans" : "      ", warps,
                   (double)h / (iters * 8 * (warps / 4.0)));
        }
    unsigned char* src; const size_t region = 8 * 64 * 1024;
    cudaMalloc(&src, 148 * region); cudaMemset(src, 1, 148 * region);
    cudaFuncSetAttribute(k_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int ctas = 16; ctas <= 148; ctas += 132 - 36)   // 16, 112
        for (int shared_src = 0; shared_src < 2; ++shared_src)
            for (int bytes = 16 * 1024; bytes <= 64 * 1024; bytes *= 2) {
                k_ldgsts<<<ctas, 256, 64 * 1024>>>(src, region, bytes, 64, shared_src, cyc, out);
                k_ldgsts<<<ctas, 256, 64 * 1024>>>(src, region, bytes, 64, shared_src, cyc, out);
                cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
                printf("cp.async 16B stream, %3d CTAs, %s source, %2d KB per round: %.0f cycles per round = %.1f B/clk/SM\n", ctas,
                       shared_src ? "shared  " : "distinct", bytes / 1024, (double)h / 64, bytes * 64.0 / (double)h);
            }
    return 0;
}
