// Microbenchmark (development tool, not product): per-SM issue throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100+)
// and of a mixed FFMA2 + LDS.128 stream shaped like the stencil gathers.  Prints FMA lanes / clk / SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_pipes fma_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
template <int ILP>
__global__ void k_fmuladd2(float* out, int iters, float x, float y) {
    unsigned long long acc[ILP];
    float2 xx = make_float2(x, x), yy = make_float2(y, y);
    unsigned long long X = *reinterpret_cast<unsigned long long*>(&xx), Y = *reinterpret_cast<unsigned long long*>(&yy);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float2 a = make_float2(threadIdx.x + i, i); acc[i] = *reinterpret_cast<unsigned long long*>(&a); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = (i & 1) ? fmul2(acc[i], X) : fadd2(acc[i], Y);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float2 a = *reinterpret_cast<float2*>(&acc[i]); s += a.x + a.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_ffma(float* out, int iters, float x, float y) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = ffma1(acc[i], x, y);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_ffma2(float* out, int iters, float x, float y) {
    unsigned long long acc[ILP];
    float2 xx = make_float2(x, x), yy = make_float2(y, y);
    unsigned long long X = *reinterpret_cast<unsigned long long*>(&xx), Y = *reinterpret_cast<unsigned long long*>(&yy);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float2 a = make_float2(threadIdx.x + i, i); acc[i] = *reinterpret_cast<unsigned long long*>(&a); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = ffma2(acc[i], X, Y);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float2 a = *reinterpret_cast<float2*>(&acc[i]); s += a.x + a.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// alternate FFMA (fma pipe) with IADD3/LOP3-class integer ops (alu pipe): can both pipes be kept busy?
template <int ILP>
__global__ void k_mix_alu(float* out, int iters, float x, float y, int z) {
    float acc[ILP]; int ia[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { acc[i] = threadIdx.x + i; ia[i] = threadIdx.x * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) { acc[i] = ffma1(acc[i], x, y); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ia[i]) : "r"(z), "r"(it)); }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i] + ia[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// gather-shaped: one LDS.128 feeds 6 FFMA2 (packed) or 9 FFMA (scalar)
template <int PACKED>
__global__ void k_gather(float* out, int iters, float x) {
    __shared__ float4 tile[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tile[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float s = 0;
    if (PACKED) {
        unsigned long long a[6] = {0, 0, 0, 0, 0, 0};
        float2 w0 = make_float2(x, x), w1 = make_float2(x + 1, x + 1), w2 = make_float2(x + 2, x + 2);
        unsigned long long W0 = *reinterpret_cast<unsigned long long*>(&w0), W1 = *reinterpret_cast<unsigned long long*>(&w1), W2 = *reinterpret_cast<unsigned long long*>(&w2);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int n = 0; n < 16; ++n) {
                const float4 v = tile[(threadIdx.x + n * 12 + it) & 1023];
                const float2 lo = make_float2(v.x, v.y), hi = make_float2(v.z, v.w);
                const unsigned long long L = *reinterpret_cast<const unsigned long long*>(&lo), H = *reinterpret_cast<const unsigned long long*>(&hi);
                a[0] = ffma2(L, W0, a[0]); a[1] = ffma2(H, W0, a[1]); a[2] = ffma2(L, W1, a[2]); a[3] = ffma2(H, W1, a[3]);
                a[4] = ffma2(L, W2, a[4]); a[5] = ffma2(H, W2, a[5]);
            }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) { float2 t = *reinterpret_cast<float2*>(&a[i]); s += t.x + t.y; }
    } else {
        float a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        const float w0 = x, w1 = x + 1, w2 = x + 2;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int n = 0; n < 16; ++n) {
                const float4 v = tile[(threadIdx.x + n * 12 + it) & 1023];
                a[0] = ffma1(v.x, w0, a[0]); a[1] = ffma1(v.y, w0, a[1]); a[2] = ffma1(v.z, w0, a[2]);
                a[3] = ffma1(v.x, w1, a[3]); a[4] = ffma1(v.y, w1, a[4]); a[5] = ffma1(v.z, w1, a[5]);
                a[6] = ffma1(v.x, w2, a[6]); a[7] = ffma1(v.y, w2, a[7]); a[8] = ffma1(v.z, w2, a[8]);
            }
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) s += a[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = pr.multiProcessorCount, nt = 512, nb = sms * 4, iters = 20000;
    float* out; cudaMalloc(&out, sizeof(float) * nb * nt);
    printf("%s, %d SMs, attr clock %d MHz (lanes/clk/SM below assume that clock; compare rows, not absolutes)\n", pr.name, sms, clk_khz / 1000);
    auto report = [&](const char* name, float ms, double fma_per_thread_iter) {
        const double fmas = (double)nb * nt * iters * fma_per_thread_iter;
        printf("%-34s %8.3f ms  %8.2f TFMA/s  %7.1f FMA lanes/clk/SM\n", name, ms, fmas / ms * 1e-9, fmas / (ms * 1e-3) / ((double)clk_khz * 1e3) / sms);
    };
    report("FFMA  ILP8",  time_ms([&] { k_ffma<8><<<nb, nt>>>(out, iters, 1.0001f, 0.5f); }), 8);
    report("FFMA  ILP16", time_ms([&] { k_ffma<16><<<nb, nt>>>(out, iters, 1.0001f, 0.5f); }), 16);
    report("FFMA2 ILP8",  time_ms([&] { k_ffma2<8><<<nb, nt>>>(out, iters, 1.0001f, 0.5f); }), 16);
    report("FFMA2 ILP16", time_ms([&] { k_ffma2<16><<<nb, nt>>>(out, iters, 1.0001f, 0.5f); }), 32);
    report("FMUL2/FADD2 ILP8 (op count)", time_ms([&] { k_fmuladd2<8><<<nb, nt>>>(out, iters, 1.0001f, 0.5f); }), 16);
    report("FFMA + LOP3 interleaved (FMA count)", time_ms([&] { k_mix_alu<8><<<nb, nt>>>(out, iters, 1.0001f, 0.5f, 77); }), 8);
    report("gather 16x(LDS.128 + 9 FFMA)",  time_ms([&] { k_gather<0><<<nb, nt>>>(out, iters / 16, 1.5f); }), 9);
    report("gather 16x(LDS.128 + 6 FFMA2)", time_ms([&] { k_gather<1><<<nb, nt>>>(out, iters / 16, 1.5f); }), 9);
    cudaFree(out);
    return 0;
}
