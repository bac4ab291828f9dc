// Pipe-throughput microbenchmark for sm_100a: how many warp instructions per cycle per SM the scalar
// FFMA, the packed FFMA2 (fp32x2), MUFU and mixed FMA+ALU streams sustain.  Decides whether packing the
// 3x3 Jacobi rotations into fp32x2 pairs buys issue slots.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_pipes tools/ubench_pipes.cu && gpurun_out/ubench_pipes
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;
constexpr int ACC = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
    float x[ACC];
    float2 y[ACC];
    unsigned u[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
        x[i] = threadIdx.x * 1e-3f + i;
        y[i] = make_float2(x[i], x[i] + 0.5f);
        u[i] = threadIdx.x + i;
    }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            if (MODE == 0) x[i] = fmaf(x[i], a, b);                         // FFMA
            if (MODE == 1) y[i] = __ffma2_rn(y[i], a2, b2);                 // FFMA2
            if (MODE == 2) {                                                // FFMA + LOP3/IADD (alu pipe)
                x[i] = fmaf(x[i], a, b);
                u[i] = (u[i] ^ 0x9e3779b9u) + u[(i + 1) % ACC];
            }
            if (MODE == 3) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i]));  // MUFU
            if (MODE == 4) {                                                // FFMA2 + FFMA
                y[i] = __ffma2_rn(y[i], a2, b2);
                x[i] = fmaf(x[i], a, b);
            }
            if (MODE == 5) y[i] = __fmul2_rn(y[i], a2);                     // FMUL2
            if (MODE == 6) y[i] = __fadd2_rn(y[i], a2);                     // FADD2
            if (MODE == 7) {                                                // FFMA2 + alu
                y[i] = __ffma2_rn(y[i], a2, b2);
                u[i] = (u[i] ^ 0x9e3779b9u) + u[(i + 1) % ACC];
            }
            if (MODE == 8) x[i] = x[i] * a;                                  // FMUL
            if (MODE == 9) x[i] = x[i] + a;                                  // FADD
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += x[i] + y[i].x + y[i].y + (float) u[i];
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char* name, int inst_per_slot, float* d) {
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 1.0001f, 1e-7f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(d, 1.0001f, 1e-7f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double warp_inst = (double) blocks * 8 * ITER * ACC * inst_per_slot;
    const double cyc = best * 1e-3 * khz * 1e3;
    printf("%-22s %8.3f ms  %6.2f warp-inst/clk/SM (at max clock %d MHz)  %6.2f per SMSP\n", name, best,
           warp_inst / cyc / sms, khz / 1000, warp_inst / cyc / sms / 4);
}

int main() {
    float* d;
    cudaMalloc(&d, 4);
    run<0>("FFMA", 1, d);
    run<8>("FMUL", 1, d);
    run<9>("FADD", 1, d);
    run<1>("FFMA2", 1, d);
    run<5>("FMUL2", 1, d);
    run<6>("FADD2", 1, d);
    run<2>("FFMA+LOP+IADD", 3, d);
    run<7>("FFMA2+LOP+IADD", 3, d);
    run<4>("FFMA2+FFMA", 2, d);
    run<3>("MUFU.RSQ", 1, d);
    return 0;
}
