// Latency and issue rate of Blackwell's 2-wide fp32 instructions next to their scalar forms (sm_100a).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o f32x2_rates f32x2_rates.cu && ./f32x2_rates
// One block per SM of 128 / 512 threads (1 / 4 warps per scheduler); every thread runs CHAINS independent dependent
// chains of LEN operations.  cycles / (LEN * CHAINS) per warp = issue cost when CHAINS is large, latency when CHAINS = 1.
#include <cstdio>
#include <cuda_runtime.h>

#define OPS(X) X(ADD1) X(MUL1) X(FMA1) X(FMA1I) X(ADD2) X(MUL2) X(FMA2) X(FMA2I) X(ADD2MUL2) X(ADD2FMA2I)
enum Op {
#define E(n) n,
    OPS(E)
#undef E
    NOPS
};
const char* NAMES[] = {
#define S(n) #n,
    OPS(S)
#undef S
};

template <int OP>
__device__ __forceinline__ void step(float2& v, float2 c) {
    unsigned long long a, b, r;
    a = *reinterpret_cast<unsigned long long*>(&v);
    b = *reinterpret_cast<unsigned long long*>(&c);
    if (OP == ADD1) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v.x) : "f"(c.x)); return; }
    if (OP == MUL1) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(v.x) : "f"(c.x)); return; }
    if (OP == FMA1) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v.x) : "f"(c.x), "f"(c.y)); return; }
    if (OP == FMA1I) { asm volatile("fma.rn.f32 %0, %0, 0f3E800000, %1;" : "+f"(v.x) : "f"(c.y)); return; }
    if (OP == ADD2) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    if (OP == MUL2) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    if (OP == FMA2) asm volatile("fma.rn.f32x2 %0, %1, %2, %2;" : "=l"(r) : "l"(a), "l"(b));
    if (OP == FMA2I) asm volatile("{ .reg .b64 q; mov.b64 q, 0x3E8000003E800000; fma.rn.f32x2 %0, %1, q, %2; }" : "=l"(r) : "l"(a), "l"(b));
    if (OP == ADD2MUL2) asm volatile("{ .reg .b64 q, t; mov.b64 q, 0x3E8000003E800000; add.rn.f32x2 t, %1, %2; mul.rn.f32x2 %0, t, q; }" : "=l"(r) : "l"(a), "l"(b));
    if (OP == ADD2FMA2I) asm volatile("{ .reg .b64 q, t; mov.b64 q, 0x3E8000003E800000; add.rn.f32x2 t, %1, %2; fma.rn.f32x2 %0, t, q, %2; }" : "=l"(r) : "l"(a), "l"(b));
    v = *reinterpret_cast<float2*>(&r);
}

template <int OP, int CHAINS>
__global__ void k(float2* out, long long* cyc, int len, float2 c) {
    float2 v[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) v[i] = make_float2(threadIdx.x * 1e-3f + i, 1.0f + i);
    __syncthreads();
    const long long t0 = clock64();
    for (int n = 0; n < len; n += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) step<OP>(v[i], c);
    }
    const long long t1 = clock64();
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { s.x += v[i].x; s.y += v[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP, int CHAINS>
double run(int threads) {
    float2* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float2));
    cudaMallocManaged(&cyc, 8);
    const int len = 4096;
    k<OP, CHAINS><<<148, threads>>>(out, cyc, len, make_float2(1.0000001f, 1e-7f));
    cudaDeviceSynchronize();
    k<OP, CHAINS><<<148, threads>>>(out, cyc, len, make_float2(1.0000001f, 1e-7f));
    cudaDeviceSynchronize();
    const double per = (double)*cyc / ((double)len * CHAINS);
    cudaFree(out); cudaFree(cyc);
    return per;
}

template <int OP>
void row() {
    printf("%-10s  latency (1 warp/sched, 1 chain) %6.2f | cycles per warp-instruction slot: 1 warp x 8 chains %5.2f, 4 warps x 8 chains %5.2f (x4 warps = %5.2f per scheduler)\n",
           NAMES[OP], run<OP, 1>(128), run<OP, 8>(128), run<OP, 8>(512), 4 * run<OP, 8>(512));
}

int main() {
    printf("ADD2MUL2 / ADD2FMA2I are two dependent instructions per step\n");
    row<ADD1>(); row<MUL1>(); row<FMA1>(); row<FMA1I>(); row<ADD2>(); row<MUL2>(); row<FMA2>(); row<FMA2I>(); row<ADD2MUL2>(); row<ADD2FMA2I>();
    return 0;
}
