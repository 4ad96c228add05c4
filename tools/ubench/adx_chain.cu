// Why does one lone warp need ~32 cycles per ADX sample? The decode worker's recurrence in isolation, one warp per
// SM sub-partition (as in adx_decode_fast_kernel), variants that add the worker's other per-sample work one piece at a
// time. Cycles per sample = clock64 difference / samples.   nvcc -arch=sm_100a -o adx_chain adx_chain.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int BLOCKS = 512;      // blocks of 32 samples per run
constexpr int SPB = 32;

// V = 0: s = (c0*s >> 12) + x, nothing else (latency of multiply -> shift-add)
// V = 1: + the c1 term one sample ahead (products from registers)
// V = 2: + products from nibbles (shift, shift, multiply-add)
// V = 3: + int16 store to shared memory per sample
// V = 4: + running min / max
// V = 5: V = 4 with the 18 code bytes loaded from shared memory per block
// `active`: bit w set = warp w of the CTA runs the chain, the others leave at once (which warps share a sub-partition?)
template <int V>
__global__ void k(const int* __restrict__ args, int* out, long long* cycles, unsigned active) {
    if (!((active >> (threadIdx.x >> 5)) & 1u)) return;
    __shared__ short pcm[32 * (SPB * 2 + 2) + 64];
    __shared__ unsigned char code[32 * 20 * 8];
    const int lane = threadIdx.x & 31;
    const int c0 = args[0], c1 = args[1], scale = args[2], zero = args[3];
    int s1 = args[4] + lane, s2 = args[5];
    for (int i = threadIdx.x; i < (int)sizeof(code); i += blockDim.x) code[i] = (unsigned char)(i * 7 + args[6]);
    int lo = 0, hi = 0;
    short* dst = pcm + lane * (SPB * 2 + 2) / 2 * 0 + (lane >> 1) * (SPB * 2 + 2) + (lane & 1);
    const long long t0 = clock64();
    for (int b = 0; b < BLOCKS; b++) {
        int blk[18];
        if (V >= 5) {
            const unsigned char* src = code + (lane * 20 * 8 + (b & 7) * 18) % (32 * 20 * 8 - 18);
#pragma unroll
            for (int i = 0; i < 18; i++) blk[i] = src[i];
        } else {
#pragma unroll
            for (int i = 0; i < 18; i++) blk[i] = (args[6] + i * 37 + b) & 0xFF;
        }
        int prod[SPB];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (V >= 2) {
                prod[2 * i] = (((int)(blk[2 + i] << 24)) >> 28) * scale + zero;
                prod[2 * i + 1] = (((int)(blk[2 + i] << 28)) >> 28) * scale + zero;
            } else {
                prod[2 * i] = blk[2 + i] + zero;
                prod[2 * i + 1] = blk[2 + i] - zero;
            }
        }
        int x = prod[0] + ((c1 * s2) >> 12);
#pragma unroll
        for (int n = 0; n < SPB; n++) {
            const int sn = ((c0 * s1) >> 12) + x;
            if (V >= 1) { if (n + 1 < SPB) x = prod[n + 1] + ((c1 * s1) >> 12); }
            else x = prod[(n + 1) & 31];
            if (V >= 3) dst[n * 2] = (short)sn;
            if (V >= 4) { lo = min(lo, sn); hi = max(hi, sn); }
            s2 = s1; s1 = sn;
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s1 + s2 + lo + hi + pcm[lane];
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) atomicMax((unsigned long long*)cycles, (unsigned long long)(t1 - t0));
}

// The same per-sample work (V = 4) with U samples per unrolled loop body: does the pace depend on the body's size?
// WAIT = 1: the other warps of the CTA do not leave, they wait at a barrier the chain warps reach when they are done
// WAIT = 2: they wait in cp.async.wait_group / spin on a shared flag instead
template <int U, int WAIT = 0>
__global__ void kbody(const int* __restrict__ args, int* out, long long* cycles, unsigned active) {
    __shared__ volatile int done_flag;
    if (threadIdx.x == 0) done_flag = 0;
    if (WAIT) __syncthreads();
    if (!((active >> (threadIdx.x >> 5)) & 1u)) {
        if (WAIT == 1) __syncthreads();
        if (WAIT == 2) { while (done_flag < 4) __nanosleep(200); }
        return;
    }
    __shared__ short pcm[16 * (2 * 256 + 2) + 64];
    const int lane = threadIdx.x & 31;
    const int c0 = args[0], c1 = args[1], scale = args[2], zero = args[3];
    int s1 = args[4] + lane, s2 = args[5];
    int lo = 0, hi = 0;
    short* dst = pcm + (lane >> 1) * (2 * 256 + 2) + (lane & 1);
    constexpr int TOTAL = BLOCKS * SPB;
    const long long t0 = clock64();
    for (int b = 0; b < TOTAL / U; b++) {
        int prod[U];
#pragma unroll
        for (int i = 0; i < U / 2; i++) {
            const int byte = (args[6] + i * 37 + b) & 0xFF;
            prod[2 * i] = (((int)(byte << 24)) >> 28) * scale + zero;
            prod[2 * i + 1] = (((int)(byte << 28)) >> 28) * scale + zero;
        }
        int x = prod[0] + ((c1 * s2) >> 12);
#pragma unroll
        for (int n = 0; n < U; n++) {
            const int sn = ((c0 * s1) >> 12) + x;
            if (n + 1 < U) x = prod[n + 1] + ((c1 * s1) >> 12);
            dst[(n & 255) * 2] = (short)sn;
            lo = min(lo, sn); hi = max(hi, sn);
            s2 = s1; s1 = sn;
        }
    }
    const long long t1 = clock64();
    if (WAIT == 1) __syncthreads();
    if (WAIT == 2 && lane == 0) atomicAdd((int*)&done_flag, 1);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s1 + s2 + lo + hi + pcm[lane];
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) atomicMax((unsigned long long*)cycles, (unsigned long long)(t1 - t0));
}

template <int WAIT>
void run_wait() {
    int h_args[8] = {0x73A, -0x4E2, 0x123, 0, 100, -50, 3, 0};
    int *d_args, *out; long long* cyc; long long h = 0;
    cudaMalloc(&d_args, sizeof h_args); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    cudaMemcpy(d_args, h_args, sizeof h_args, cudaMemcpyHostToDevice);
    kbody<32, WAIT><<<148, 512>>>(d_args, out, cyc, 0x000Fu);
    cudaMemset(cyc, 0, 8);
    kbody<32, WAIT><<<148, 512>>>(d_args, out, cyc, 0x000Fu);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps 0-3 run the chain, the other twelve %s: %6.2f cycles per sample\n", WAIT == 1 ? "wait at the CTA barrier" : WAIT == 2 ? "sleep-poll a flag" : "exit", (double)h / (BLOCKS * SPB));
    cudaFree(d_args); cudaFree(out); cudaFree(cyc);
}

template <int U>
void run_body() {
    int h_args[8] = {0x73A, -0x4E2, 0x123, 0, 100, -50, 3, 0};
    int *d_args, *out; long long* cyc; long long h = 0;
    cudaMalloc(&d_args, sizeof h_args); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    cudaMemcpy(d_args, h_args, sizeof h_args, cudaMemcpyHostToDevice);
    for (unsigned active : {0x000Fu, 0xFFFFu}) {
        kbody<U><<<148, 512>>>(d_args, out, cyc, active);
        cudaMemset(cyc, 0, 8);
        kbody<U><<<148, 512>>>(d_args, out, cyc, active);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("body of %3d samples, warps active %04x: %6.2f cycles per sample\n", U, active, (double)h / (BLOCKS * SPB));
    }
    cudaFree(d_args); cudaFree(out); cudaFree(cyc);
}

template <int V>
void run(const char* what, int warps, unsigned active = 0xFFFFFFFFu) {
    int h_args[8] = {0x73A, -0x4E2, 0x123, 0, 100, -50, 3, 0};
    int *d_args, *out; long long* cyc; long long h = 0;
    cudaMalloc(&d_args, sizeof h_args); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    cudaMemcpy(d_args, h_args, sizeof h_args, cudaMemcpyHostToDevice);
    k<V><<<148, warps * 32>>>(d_args, out, cyc, active);
    cudaMemset(cyc, 0, 8);
    k<V><<<148, warps * 32>>>(d_args, out, cyc, active);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("V=%d %-58s warps/CTA %2d active %08x: %6.2f cycles per sample\n", V, what, warps, active, (double)h / (BLOCKS * SPB));
    cudaFree(d_args); cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int warps : {4, 8, 16}) {     // 1, 2, 4 warps per sub-partition
        run<0>("multiply -> shift-add chain only", warps);
        run<1>("+ c1 term one sample ahead", warps);
        run<2>("+ products from nibbles", warps);
        run<3>("+ int16 store to shared memory", warps);
        run<4>("+ running min / max", warps);
        run<5>("+ code bytes from shared memory", warps);
    }
    // four chain warps in a CTA of sixteen: which four?
    run<4>("warps 0,1,2,3 of 16", 16, 0x000Fu);
    run<4>("warps 0,4,8,12 of 16", 16, 0x1111u);
    run<4>("warps 0,5,10,15 of 16", 16, 0x8421u);
    run<4>("warps 0,1 of 16", 16, 0x0003u);
    run<4>("warps 0,4 of 16", 16, 0x0011u);
    run<4>("warps 0,2 of 16", 16, 0x0005u);
    run_wait<0>(); run_wait<1>(); run_wait<2>();
    run_body<8>(); run_body<16>(); run_body<32>(); run_body<48>(); run_body<64>(); run_body<96>(); run_body<256>();
    return 0;
}
