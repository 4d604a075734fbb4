// Inner loop of train_fused.cu's forward pass in isolation: one 64-row weight chunk resident in shared memory,
// 16-row activation tile, 8 warps, tiles warp + 8 i (26 tiles -> 4,4,3,3,3,3,3,3), bf16x3 MMAs.  Reports cycles per
// chunk for: fragment loads + MMAs, fragment loads only, MMAs only -- to see which of them bounds the pass.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fwd_loop fwd_loop.cu && ./fwd_loop
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int Ns = 216, As = 216;
// MODE 0: loads + MMAs; 1: loads only; 2: MMAs only; 3: loads + MMAs with B fragments fetched as x4.trans pairs
template <int MODE, int NTW>
__device__ __forceinline__ void chunk(float (&acc)[4][4], const __nv_bfloat16* Ah, const __nv_bfloat16* Al, const __nv_bfloat16* Wh,
                                      const __nv_bfloat16* Wl, int a_off, int b_off, int warp, uint32_t& sink) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        uint32_t ah[4], al[4], bh[NTW][2], bl[NTW][2];
        if (MODE != 2) {
            ldsm_x4(ah, smem_u32(Ah + a_off + kk * 16));
            ldsm_x4(al, smem_u32(Al + a_off + kk * 16));
#pragma unroll
            for (int i = 0; i < NTW; ++i) {
                const int off = kk * 16 * Ns + b_off + (warp + 8 * i) * 8;
                ldsm_x2_t(bh[i], smem_u32(Wh + off));
                ldsm_x2_t(bl[i], smem_u32(Wl + off));
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) { ah[q] = 0x3f803f80u + kk; al[q] = 0x3c003c00u + q; }
#pragma unroll
            for (int i = 0; i < NTW; ++i) { bh[i][0] = 0x3f803f80u + i; bh[i][1] = 0x3f803f81u; bl[i][0] = 0x3c003c00u; bl[i][1] = 0x3c003c01u + kk; }
        }
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < NTW; ++i) sink ^= bh[i][0] ^ bh[i][1] ^ bl[i][0] ^ bl[i][1];
            sink ^= ah[0] ^ ah[1] ^ ah[2] ^ ah[3] ^ al[0] ^ al[1] ^ al[2] ^ al[3];
        } else {
#pragma unroll
            for (int i = 0; i < NTW; ++i) { mma16816(acc[i], al, bh[i]); mma16816(acc[i], ah, bl[i]); mma16816(acc[i], ah, bh[i]); }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k_fwd(float* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) unsigned char sm[];
    __nv_bfloat16* Wh = reinterpret_cast<__nv_bfloat16*>(sm);
    __nv_bfloat16* Wl = Wh + 64 * Ns;
    __nv_bfloat16* Ah = Wl + 64 * Ns;
    __nv_bfloat16* Al = Ah + 16 * As;
    for (int i = threadIdx.x; i < (2 * 64 * Ns + 2 * 16 * As); i += 256) Wh[i] = __float2bfloat16(0.001f * (i % 97));
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a_off = ((lane & 7) + 8 * ((lane >> 3) & 1)) * As + 8 * (lane >> 4), b_off = (lane & 15) * Ns;
    float acc[4][4] = {};
    uint32_t sink = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (warp < 2) chunk<MODE, 4>(acc, Ah, Al, Wh, Wl, a_off + (it & 1) * 64, b_off, warp, sink);
        else chunk<MODE, 3>(acc, Ah, Al, Wh, Wl, a_off + (it & 1) * 64, b_off, warp, sink);
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = __uint_as_float(sink);
    for (int i = 0; i < 4; ++i) for (int q = 0; q < 4; ++q) s += acc[i][q];
    out[blockIdx.x * 256 + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}


// dX chunk of train_fused.cu: 8 output tiles per 64-row chunk; warps as 4 pairs, each warp two tiles x half of the reduction
// (KH k-steps), g fragments (hi + lo) resident in registers, one ldmatrix.x4 per tile and k-step.
// MODE 0: loads + MMAs; 1: loads only; 2: MMAs only
template <int MODE>
__global__ void __launch_bounds__(256, 1) k_dx(float* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) unsigned char sm[];
    __nv_bfloat16* Wh = reinterpret_cast<__nv_bfloat16*>(sm);
    __nv_bfloat16* Ah = Wh + 2 * 64 * Ns;
    __nv_bfloat16* Al = Ah + 16 * As;
    for (int i = threadIdx.x; i < (2 * 64 * Ns + 2 * 16 * As); i += 256) Wh[i] = __float2bfloat16(0.001f * (i % 97));
    __syncthreads();
    constexpr int KH = 7;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, upper = warp >> 2, pair = warp & 3;
    const int k_begin = upper ? KH : 0, k_count = upper ? 6 : 7;
    const int a_off = ((lane & 7) + 8 * ((lane >> 3) & 1)) * As + 8 * (lane >> 4);
    uint32_t gfh[KH][4], gfl[KH][4];
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) {
        ldsm_x4(gfh[kk], smem_u32(Ah + a_off + (k_begin + (kk < k_count ? kk : 0)) * 16));
        ldsm_x4(gfl[kk], smem_u32(Al + a_off + (k_begin + (kk < k_count ? kk : 0)) * 16));
    }
    const int w_off = (pair * 16 + (lane & 7)) * Ns + 8 * ((lane >> 3) & 1) + k_begin * 16 + (lane >> 4) * (64 * Ns);
    float tot[2][4] = {};
    uint32_t sink = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float c3[2][3][4] = {};
#pragma unroll
        for (int kk = 0; kk < KH; ++kk) {
            if (kk < k_count) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint32_t b4[4];
                    if (MODE != 2) ldsm_x4(b4, smem_u32(Wh + w_off + j * 8 * Ns + kk * 16 + (it & 1) * 8));
                    else { b4[0] = 0x3f803f80u + kk; b4[1] = 0x3f803f81u + j; b4[2] = 0x3c003c00u; b4[3] = 0x3c003c01u + it; }
                    if (MODE == 1) { sink ^= b4[0] ^ b4[1] ^ b4[2] ^ b4[3]; continue; }
                    const uint32_t bh[2] = {b4[0], b4[1]}, bl[2] = {b4[2], b4[3]};
                    mma16816(c3[j][0], gfl[kk], bh);
                    mma16816(c3[j][1], gfh[kk], bl);
                    mma16816(c3[j][2], gfh[kk], bh);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) tot[j][q] += (c3[j][0][q] + c3[j][1][q]) + c3[j][2][q];
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = __uint_as_float(sink);
    for (int j = 0; j < 2; ++j) for (int q = 0; q < 4; ++q) s += tot[j][q];
    out[blockIdx.x * 256 + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float* out; long long* cyc; long long h = 0;
    cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
    const int iters = 1000, smem = (2 * 64 * Ns + 2 * 16 * As) * 2;
#define RUN(MODE, NAME)                                                                              \
    cudaFuncSetAttribute(k_fwd<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);            \
    k_fwd<MODE><<<148, 256, smem>>>(out, cyc, iters); k_fwd<MODE><<<148, 256, smem>>>(out, cyc, iters); \
    cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                         \
    printf("%-28s %.0f cycles per 64-row chunk (26 tiles x 4 k-steps x 3 MMAs = 312 MMAs, 78 per sub-partition)  %s\n", NAME, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
    RUN(0, "fragment loads + MMAs") RUN(1, "fragment loads only") RUN(2, "MMAs only")
#define RUNX(MODE, NAME)                                                                             \
    cudaFuncSetAttribute(k_dx<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);             \
    k_dx<MODE><<<148, 256, smem>>>(out, cyc, iters); k_dx<MODE><<<148, 256, smem>>>(out, cyc, iters); \
    cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                         \
    printf("dX pairs: %-24s %.0f cycles per 64-row chunk (8 tiles x 13 k-steps x 3 MMAs = 312 MMAs)  %s\n", NAME, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
    RUNX(0, "fragment loads + MMAs") RUNX(1, "fragment loads only") RUNX(2, "MMAs only")
    return 0;
}
