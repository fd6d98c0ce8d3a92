// dpe_microbench.cu -- roofline denominators measured in the same run as the
// kernels they bound: FP32 FMA-pipe throughput (scalar FFMA and packed FFMA2)
// and a device-memory copy.  bench.py reports fractions against these and the
// driver-written MEASURED_PEAKS.json.
#include "dpe_internal.cuh"

namespace dpe {

// 16 independent accumulator chains per thread, 8 warps per SM sub-partition's worth
// of CTAs: nothing but FMAs in the loop body.  The packed variant uses the operand shape that
// reaches the FFMA2 issue limit on B200 (one warp instruction per 2 cycles per sub-partition =
// 73 TFLOP/s at 1965 MHz): a scalar .F32 multiplier per chain and one 64-bit pair shared by all
// chains; `acc = ffma2(acc, a2, b2)` with three 64-bit register sources tops out 7 % lower.
template <int PACKED>
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float a, float b) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 a2 = make_float2(a, a * 0.999f), b2 = make_float2(b, b * 1.001f);
    float cm[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) cm[i] = a + 1e-3f * (float)((threadIdx.x + i) & 7);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (PACKED) {
                    acc[i] = __ffma2_rn(make_float2(cm[i], cm[i]), b2, acc[i]);
                } else {
                    acc[i].x = fmaf(acc[i].x, a2.x, b2.x);
                    acc[i].y = fmaf(acc[i].y, a2.y, b2.y);
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    if (s == 12345.678f) out[0] = s;   // keep the chains alive
}

// FP64 pipe: 16 independent DFMA chains per thread (the lookup-path scoring kernels are bound by this pipe)
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) b[i] = a[i];
}

}  // namespace dpe

using namespace dpe;

extern "C" int dpe_microbench_fp32(int device, int use_ffma2, double* tflops) {
    if (!tflops) { set_error("null argument"); return DPE_EINVAL; }
    DPE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    DPE_CUDA(cudaGetDeviceProperties(&prop, device));
    float* out;
    DPE_CUDA(cudaMalloc(&out, 64));
    const int iters = 4096, blocks = prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    DPE_CUDA(cudaEventCreate(&e0));
    DPE_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        DPE_CUDA(cudaEventRecord(e0));
        if (use_ffma2) k_fma_peak<1><<<blocks, 256>>>(out, iters, 0.999f, 1e-3f);
        else k_fma_peak<0><<<blocks, 256>>>(out, iters, 0.999f, 1e-3f);
        DPE_CUDA(cudaEventRecord(e1));
        DPE_CUDA(cudaEventSynchronize(e1));
        float ms;
        DPE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = (double)blocks * 256 * iters * 4 * 16 * 2 * 2;   // 2 lanes x (mul+add)
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *tflops = best;
    return DPE_OK;
}

extern "C" int dpe_microbench_fp64(int device, double* tflops) {
    if (!tflops) { set_error("null argument"); return DPE_EINVAL; }
    DPE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    DPE_CUDA(cudaGetDeviceProperties(&prop, device));
    double* out;
    DPE_CUDA(cudaMalloc(&out, 64));
    const int iters = 512, blocks = prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    DPE_CUDA(cudaEventCreate(&e0));
    DPE_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        DPE_CUDA(cudaEventRecord(e0));
        k_dfma_peak<<<blocks, 256>>>(out, iters, 0.999, 1e-3);
        DPE_CUDA(cudaEventRecord(e1));
        DPE_CUDA(cudaEventSynchronize(e1));
        float ms;
        DPE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = (double)blocks * 256 * iters * 4 * 16 * 2;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *tflops = best;
    return DPE_OK;
}

extern "C" int dpe_microbench_hbm(int device, size_t bytes, double* gbs) {
    if (!gbs) { set_error("null argument"); return DPE_EINVAL; }
    DPE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    DPE_CUDA(cudaGetDeviceProperties(&prop, device));
    bytes &= ~(size_t)15;
    float4 *a, *b;
    DPE_CUDA(cudaMalloc(&a, bytes));
    DPE_CUDA(cudaMalloc(&b, bytes));
    DPE_CUDA(cudaMemset(a, 1, bytes));
    cudaEvent_t e0, e1;
    DPE_CUDA(cudaEventCreate(&e0));
    DPE_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        DPE_CUDA(cudaEventRecord(e0));
        k_copy<<<prop.multiProcessorCount * 16, 256>>>(a, b, bytes / 16);
        DPE_CUDA(cudaEventRecord(e1));
        DPE_CUDA(cudaEventSynchronize(e1));
        float ms;
        DPE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double g = 2.0 * bytes / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(a); cudaFree(b);
    *gbs = best;
    return DPE_OK;
}
