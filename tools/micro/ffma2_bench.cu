// Micro-benchmark (development tool): issue rate and dependent latency of FFMA vs FFMA2 (packed f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, bool PACKED>
__global__ void kern (float* out, int iters, float s)
{
    float2 a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i)
        a[i] = make_float2 (threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
    const float2 b = make_float2 (s, s * 0.5f), c = make_float2 (1e-3f, 2e-3f);
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
        {
            if (PACKED)
                a[i] = __ffma2_rn (a[i], b, c);
            else
            {
                a[i].x = __fmaf_rn (a[i].x, b.x, c.x);
                a[i].y = __fmaf_rn (a[i].y, b.y, c.y);
            }
        }
    }
    float acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i)
        acc += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int ILP, bool PACKED>
void run (const char* name, int warps_per_sm)
{
    float* out;
    cudaMalloc (&out, 148 * 1024 * 4 * 4);
    const int iters = 20000;
    const int threads = 32 * warps_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate (&e0);
    cudaEventCreate (&e1);
    kern<ILP, PACKED><<<148, threads>>> (out, 100, 0.999f);
    cudaEventRecord (e0);
    kern<ILP, PACKED><<<148, threads>>> (out, iters, 0.999f);
    cudaEventRecord (e1);
    cudaEventSynchronize (e1);
    float ms;
    cudaEventElapsedTime (&ms, e0, e1);
    // per SMSP: warps_per_sm/4 warps; fma-lanes = ILP*2 per iteration per thread
    const double flop_per_thread = (double) iters * ILP * 2;
    const double cycles = ms * 1e-3 * 1.965e9;
    printf ("%-28s warps/SM %2d ILP %d: %.3f ms, %.2f cycles per iteration per warp-slot, %.2f fp32-fma lanes/clk/SM\n", name, warps_per_sm, ILP, ms, cycles / iters,
            flop_per_thread * threads / cycles);
    cudaFree (out);
}

int main ()
{
    run<1, false> ("FFMA  dependent pair", 4);
    run<1, true> ("FFMA2 dependent", 4);
    run<8, false> ("FFMA  8 pairs", 4);
    run<8, true> ("FFMA2 8 packed", 4);
    run<8, false> ("FFMA  8 pairs", 16);
    run<8, true> ("FFMA2 8 packed", 16);
    run<4, false> ("FFMA  4 pairs", 32);
    run<4, true> ("FFMA2 4 packed", 32);
    return 0;
}
