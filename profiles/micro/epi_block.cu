// The arithmetic of ONE epilogue block of the rollout kernel (16 pre-activations per lane -> swish ->
// bf16 hi/lo split; pe_tc.cu do_block) in isolation: registers in, registers out, no tensor
// memory and no barriers.  Cycles per block per SM sub-partition with 1, 2 and 3 warps per
// sub-partition show whether the 400 cycles/block seen in the kernel are the instruction mix.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../mirl_b200/csrc/tc_ptx.cuh"
using namespace mb;

template <int VARIANT>
__device__ __forceinline__ void block(const uint32_t (&cur)[16], uint32_t (&h)[8], uint32_t (&lo)[8]) {
    const uint64_t kNegLog2e = pack2(-1.4426950408889634f, -1.4426950408889634f);
    const uint64_t kOne = pack2(1.0f, 1.0f);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const uint64_t z = pack2(__uint_as_float(cur[2 * t]), __uint_as_float(cur[2 * t + 1]));
        float e0, e1;
        unpack2(mul2(z, kNegLog2e), e0, e1);
        const uint64_t y = add2(pack2(ex2_ftz(e0), ex2_ftz(e1)), kOne);
        float y0, y1;
        unpack2(y, y0, y1);
        uint64_t a = mul2(z, pack2(rcp_ftz(y0), rcp_ftz(y1)));
        float a0, a1;
        unpack2(a, a0, a1);
        if (VARIANT == 0) {                     // the kernel's split: round-to-nearest packs
            h[t] = cvt_bf16x2(a0, a1);
            const uint64_t hf = pack2(__uint_as_float(h[t] << 16), __uint_as_float(h[t] & 0xffff0000u));
            float l0, l1;
            unpack2(sub2(a, hf), l0, l1);
            lo[t] = cvt_bf16x2(l0, l1);
        } else if (VARIANT == 1) {              // swish only (no split)
            h[t] = __float_as_uint(a0);
            lo[t] = __float_as_uint(a1);
        } else {                                // truncating split: PRMT packs, no F2FP
            const uint32_t u0 = __float_as_uint(a0), u1 = __float_as_uint(a1);
            asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(h[t]) : "r"(u0), "r"(u1));
            const uint64_t hf = pack2(__uint_as_float(u0 & 0xffff0000u), __uint_as_float(u1 & 0xffff0000u));
            float l0, l1;
            unpack2(sub2(a, hf), l0, l1);
            asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(lo[t]) : "r"(__float_as_uint(l0)), "r"(__float_as_uint(l1)));
        }
    }
}

template <int VARIANT>
__global__ void k(uint32_t* out, long long* cyc, int iters) {
    uint32_t cur[16], h[8], lo[8], acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) cur[i] = __float_as_uint(0.01f * (threadIdx.x + 3 * i) - 1.0f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        block<VARIANT>(cur, h, lo);
#pragma unroll
        for (int i = 0; i < 8; ++i) {           // next block's inputs depend on this one's outputs (2 ALU ops per word)
            acc ^= h[i] + lo[i];
            cur[2 * i] = (cur[2 * i] & 0xbfffffffu) ^ (h[i] & 0x00010000u);
            cur[2 * i + 1] = (cur[2 * i + 1] & 0xbfffffffu) ^ (lo[i] & 0x00010000u);
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int VARIANT>
void run(const char* name, int warps) {
    uint32_t* out; long long* cyc; long long hcyc = 0;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2048;
    k<VARIANT><<<148, 32 * warps>>>(out, cyc, iters);
    k<VARIANT><<<148, 32 * warps>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(&hcyc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-22s %d warp(s)/sub-partition: %6.1f cycles per block per warp, %6.1f per block per sub-partition\n", name,
           warps / 4, (double)hcyc / iters, (double)hcyc / iters / (warps / 4));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w = 4; w <= 12; w += 4) run<0>("swish + RN split", w);
    for (int w = 4; w <= 12; w += 4) run<1>("swish only", w);
    for (int w = 4; w <= 12; w += 4) run<2>("swish + trunc split", w);
    return 0;
}
