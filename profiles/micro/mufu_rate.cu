// Throughput of the MUFU ops the rollout epilogue uses, in cycles per warp instruction per SM
// sub-partition: one warp per sub-partition (128-thread blocks), 16 independent chains.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ float op(float x) {
    float y;
    if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    else if (OP == 1) {     // + 1 so that ptxas cannot fold rcp(rcp(x)); costs one extra issue slot
        asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        asm volatile("add.ftz.f32 %0, %0, 0f3F800000;" : "+f"(y));
    }
    else if (OP == 2) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    else if (OP == 3) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    else if (OP == 5) {     // F2FP.BF16.F32.PACK_AB: which pipe does the bf16 pack run on?
        unsigned u;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(u) : "f"(x));
        y = __uint_as_float(u);
    } else if (OP == 6) {   // PRMT (upper halves of two words): the ALU-pipe alternative (truncation)
        unsigned u;
        asm volatile("prmt.b32 %0, %1, %1, 0x7632;" : "=r"(u) : "r"(__float_as_uint(x)));
        y = __uint_as_float(u);
    } else asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int OP, int WARPS>
__global__ void k(float* out, long long* cyc, int iters) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.5f + 0.01f * (threadIdx.x + i);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = op<OP>(v[i]);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP, int WARPS>
void run(const char* name) {
    float* out; long long* cyc; long long h = 0;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<OP, WARPS><<<148, 32 * WARPS>>>(out, cyc, iters);
    k<OP, WARPS><<<148, 32 * WARPS>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // WARPS/4 warps per sub-partition, 16 instructions per iteration each
    printf("%-6s %2d warps/SM: %.2f cycles per warp instruction per sub-partition\n", name, WARPS,
           (double)h / ((double)iters * 16 * (WARPS / 4)));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0, 4>("ex2"); run<1, 4>("rcp"); run<2, 4>("lg2"); run<3, 4>("tanh"); run<4, 4>("ffma");
    run<5, 4>("f2fp"); run<6, 4>("prmt");
    run<0, 8>("ex2"); run<1, 8>("rcp"); run<4, 8>("ffma"); run<5, 8>("f2fp");
    return 0;
}
