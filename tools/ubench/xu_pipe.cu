// micro-benchmark of the softmax instruction mix on one SM sub-partition: MUFU.EX2, F2FP, FFMA2, FADD2, FMNMX3 alone and mixed.
// Prints clocks per loop iteration (8 elements per thread) for 1, 2, 4 warps per scheduler.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o xu_pipe xu_pipe.cu ; run: ./xu_pipe
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
// bit 0 MUFU x8, bit 1 F2FP x4, bit 2 FFMA2 x4, bit 3 FADD2 x4, bit 4 FMNMX3 x4
template <int M>
__global__ void k(float* out, long long* clk, float seed) {
    float a[8], mx = -1e30f;
    uint64_t acc[4] = {0, 0, 0, 0};
    uint32_t h[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i * 0.01f;
    const uint64_t c1 = pack2(0.999f, 0.999f), c2 = pack2(-0.001f, -0.001f);
    const long long t0 = clock64();
    for (int it = 0; it < 512; ++it) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            uint64_t v = pack2(a[i], a[i + 1]);
            if (M & 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(c1), "l"(c2));
            unpack2(v, x[i], x[i + 1]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (M & 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            if (M & 8) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc[i >> 1]) : "l"(pack2(x[i], x[i + 1])));
            if (M & 2) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i >> 1]) : "f"(x[i]), "f"(x[i + 1]));
            if (M & 16) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(mx) : "f"(x[i]), "f"(x[i + 1]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = x[i];
    }
    const long long t1 = clock64();
    float s = mx;
    for (int i = 0; i < 8; ++i) s += a[i];
    for (int i = 0; i < 4; ++i) { float p, q; unpack2(acc[i], p, q); s += p + q + h[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int M>
void run(const char* name) {
    float* out; long long* clk; cudaMalloc(&out, 4 * 1024 * 148); cudaMalloc(&clk, 8);
    printf("%-44s", name);
    for (int warps : {4, 8, 16}) {
        k<M><<<148, warps * 32>>>(out, clk, 0.5f); k<M><<<148, warps * 32>>>(out, clk, 0.5f);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
        printf("  %dw/sched: %6.1f clk/iter/warp-slot", warps / 4, (double)c / 512.0 / (warps / 4.0));
    }
    printf("\n");
    cudaFree(out); cudaFree(clk);
}
int main() {
    printf("per iteration = 8 elements per thread; MUFU floor = 64 clk per iteration per warp on a scheduler\n");
    run<1>("MUFU x8");
    run<2>("F2FP x4");
    run<4>("FFMA2 x4");
    run<8>("FADD2 x4");
    run<16>("FMNMX3 x4");
    run<1 | 2>("MUFU x8 + F2FP x4");
    run<1 | 4>("MUFU x8 + FFMA2 x4");
    run<1 | 8>("MUFU x8 + FADD2 x4");
    run<1 | 4 | 8>("MUFU x8 + FFMA2 x4 + FADD2 x4");
    run<1 | 2 | 4 | 8>("MUFU x8 + FFMA2 x4 + FADD2 x4 + F2FP x4");
    run<31>("all: + FMNMX3 x4 (the softmax mix)");
    return 0;
}
