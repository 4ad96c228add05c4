// Issue rates of the fp32 forms the HCA transform kernel uses, per SM sub-partition (SMSP), on sm_100a:
// scalar FADD / FMUL-immediate / FMUL-register, and the two-wide FADD2 / FFMA2. Each warp runs N_CHAIN independent
// dependency chains so that latency does not bound the rate; W warps per SMSP.  nvcc -arch=sm_100a -o fp32_rates fp32_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long ull;
constexpr int N_CHAIN = 16, ITERS = 4096;

template <int OP>
__global__ void k(float* out, ull one, float seed, long long* cycles) {
    float v[N_CHAIN];
    ull p[N_CHAIN / 2];
    for (int i = 0; i < N_CHAIN; i++) v[i] = seed + i + threadIdx.x;
    for (int i = 0; i < N_CHAIN / 2; i++) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(v[2 * i]), "f"(v[2 * i + 1]));
    ull q; asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(seed), "f"(seed));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < N_CHAIN; i++) {
            if (OP == 0) v[i] = __fadd_rn(v[i], seed);
            if (OP == 1) v[i] = __fmul_rn(v[i], 1.0000001f);
            if (OP == 2) v[i] = __fmul_rn(v[i], seed);
            if (OP == 3 && i < N_CHAIN / 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q));
            if (OP == 4 && i < N_CHAIN / 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(one), "l"(q));
            if (OP == 5) { if (i & 1) v[i] = __fmul_rn(v[i], 1.0000001f); else if (i < N_CHAIN / 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(one), "l"(q)); }
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < N_CHAIN; i++) s += v[i];
    for (int i = 0; i < N_CHAIN / 2; i++) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p[i])); s += a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, int warps_per_smsp, int per_iter) {
    float* out; long long* cyc; long long h = 0;
    cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 8);
    const int threads = warps_per_smsp * 4 * 32;
    k<OP><<<148, threads>>>(out, 0x3F8000003F800000ull, 1.0f, cyc);
    k<OP><<<148, threads>>>(out, 0x3F8000003F800000ull, 1.0f, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double instr = (double)ITERS * per_iter * warps_per_smsp;       // warp instructions per SMSP
    printf("%-28s warps/SMSP %d: %.3f warp-instr/clk/SMSP (%.2f clk each)\n", name, warps_per_smsp, instr / h, h / instr);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {1, 2, 4}) {
        run<0>("FADD reg,reg", w, N_CHAIN);
        run<1>("FMUL reg,imm", w, N_CHAIN);
        run<2>("FMUL reg,reg", w, N_CHAIN);
        run<3>("FADD2", w, N_CHAIN / 2);
        run<4>("FFMA2 (uniform multiplier)", w, N_CHAIN / 2);
        run<5>("FMUL imm + FFMA2 mixed 2:1", w, N_CHAIN / 2 + N_CHAIN / 4);
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
