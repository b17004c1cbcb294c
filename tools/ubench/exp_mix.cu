// micro-benchmark: the attention softmax inner loop (per 16 elements: scale FFMA2, ex2, row-sum FADD2, F16 pack, row-max FMNMX3) with
// K of every 16 exponentials evaluated on the FMA pipe instead of the MUFU:
//     2^x = 2^n * p(f),  n = floor(x) (add.rm with the 1.5 * 2^23 magic), f = x - n in [0, 1), p cubic (max rel. error 7.5e-5 < F16 ulp / 2),
//     exponent inserted with one integer shift-add.
// Prints clocks per 16 elements per warp slot for 1 / 2 / 4 warps per scheduler; the pure-MUFU floor is 128 clk.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_mix exp_mix.cu ; run: ./exp_mix
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t add2_rm(uint64_t a, uint64_t b) { uint64_t d; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// two exponentials on the FMA pipe (packed): 3 FADD2-class + 3 FFMA2 + 2 integer ops
__device__ __forceinline__ void exp2_poly2(uint64_t x, float& p0, float& p1) {
    const float MAGIC = 12582912.0f;                       // 1.5 * 2^23: x + MAGIC rounded down leaves floor(x) in the low mantissa bits
    const uint64_t magic2 = pack2(MAGIC, MAGIC), nmagic2 = pack2(-MAGIC, -MAGIC);
    const uint64_t t = add2_rm(x, magic2);
    const uint64_t n = add2(t, nmagic2);                   // floor(x) as float
    const uint64_t f = fma2(n, pack2(-1.f, -1.f), x);      // x - floor(x) in [0, 1)
    uint64_t p = fma2(pack2(0.07802403f, 0.07802403f), f, pack2(0.22606707f, 0.22606707f));
    p = fma2(p, f, pack2(0.69583398f, 0.69583398f));
    p = fma2(p, f, pack2(0.99992514f, 0.99992514f));
    float t0, t1, q0, q1;
    unpack2(t, t0, t1);
    unpack2(p, q0, q1);
    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

template <int K>   // K of 16 elements (K even) take the polynomial path
__global__ void k(float* out, long long* clk, float seed, float* err) {
    float s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = -seed * (1 + ((threadIdx.x * 16 + i) % 97)) * 0.13f;
    uint64_t rs2 = 0;
    float mx = -1e30f;
    uint32_t h[8];
    const uint64_t l2e2 = pack2(1.4426950408889634f, 1.4426950408889634f), nm2 = pack2(-0.25f, -0.25f);
    const long long t0 = clock64();
    for (int it = 0; it < 512; ++it) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            mx = fmaxf(mx, fmaxf(s[i], s[i + 1]));
            const uint64_t x = fma2(pack2(s[i], s[i + 1]), l2e2, nm2);
            float p0, p1;
            // polynomial pairs are spread evenly through the 8 pairs so that MUFU and FMA work interleave in the instruction stream
            if ((i / 2) * K / 16 != (i / 2 + 1) * K / 16) {
                exp2_poly2(x, p0, p1);
            } else {
                float x0, x1;
                unpack2(x, x0, x1);
                p0 = ex2(x0);
                p1 = ex2(x1);
            }
            rs2 = add2(rs2, pack2(p0, p1));
            asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i >> 1]) : "f"(p1), "f"(p0));
            s[i] = s[i] * 0.999f + p0 * 1e-6f;          // keep the loop-carried dependency honest without changing the mix much
        }
    }
    const long long t1 = clock64();
    float a, b;
    unpack2(rs2, a, b);
    float acc = a + b + mx;
    for (int i = 0; i < 8; ++i) acc += h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        *clk = t1 - t0;
        float worst = 0.f;                                 // accuracy of the polynomial path over [-40, 8]
        for (int j = 0; j < 4000; ++j) {
            const float x = -40.f + 48.f * j / 4000.f;
            float p0, p1;
            exp2_poly2(pack2(x, x + 0.003f), p0, p1);
            worst = fmaxf(worst, fabsf(p0 / exp2f(x) - 1.f));
        }
        *err = worst;
    }
}

template <int K>
void run() {
    float *out, *err; long long* clk;
    cudaMalloc(&out, 4 * 1024 * 148); cudaMalloc(&clk, 8); cudaMalloc(&err, 4);
    printf("%2d of 16 exponentials on the FMA pipe:", K);
    for (int warps : {4, 8, 16}) {
        k<K><<<148, warps * 32>>>(out, clk, 0.5f, err); k<K><<<148, warps * 32>>>(out, clk, 0.5f, err);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
        printf("  %dw/sched %6.1f clk per 16 elements per warp slot", warps / 4, (double)c / 512.0 / (warps / 4.0));
    }
    float e; cudaMemcpy(&e, err, 4, cudaMemcpyDeviceToHost);
    printf("   (poly max rel err %.2e)\n", e);
    cudaFree(out); cudaFree(clk); cudaFree(err);
}
int main() {
    printf("pure-MUFU floor: 16 x 8 = 128 clk per 16 elements per warp on one scheduler\n");
    run<0>(); run<2>(); run<4>(); run<6>(); run<8>(); run<10>(); run<16>();
    return 0;
}
