// ffma2_pattern.cu -- does the k_brute instruction pattern (32 blends, then 64 accumulates per pair of
// positions, 64 accumulator pairs + 32 alphas live) reach the FFMA2 issue limit when fed from registers?
// cycles per FFMA2 per SM sub-partition, 2 warps per sub-partition.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int NC, int ORDER>
__global__ void __launch_bounds__(256, 1) k_pat(const float* __restrict__ in, float* out, int iters) {
    float al[NC]; float2 acc[NC];
    float4 x[4], rd[4];
    for (int j = 0; j < NC; ++j) { al[j] = in[j + threadIdx.x]; acc[j] = make_float2(0.f, 0.f); }
    for (int q = 0; q < 4; ++q) {
        x[q] = make_float4(in[100 + q + threadIdx.x], in[110 + q], in[120 + q + threadIdx.x], in[130 + q]);
        rd[q] = make_float4(in[200 + q + threadIdx.x], in[210 + q], in[220 + q + threadIdx.x], in[230 + q]);
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const float eps = __ldcg(in + 4000 + (it & 7));     // 0 at run time, unknown at compile time
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            // make every shared operand loop-variant so nothing is hoisted (8 cheap ops per q, all modes alike)
            rd[q].x += eps; rd[q].y += eps; rd[q].z += eps; rd[q].w += eps;
            x[q].x += eps; x[q].y += eps; x[q].z += eps; x[q].w += eps;
            const float2 dp = make_float2(rd[q].x, rd[q].y), r0 = make_float2(rd[q].z, rd[q].w);
            const float2 xa = make_float2(x[q].x, x[q].y), xb = make_float2(x[q].z, x[q].w);
            if (ORDER == 0) {                 // as compiled today: per candidate blend + 2 accumulates (compiler reorders)
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    const float2 bp = __ffma2_rn(make_float2(al[j], al[j]), dp, r0);
                    acc[j] = __ffma2_rn(make_float2(bp.x, bp.x), xa, acc[j]);
                    acc[j] = __ffma2_rn(make_float2(bp.y, bp.y), xb, acc[j]);
                }
            } else if (ORDER == 2) {          // 3 accumulates, scalar = half of a long-lived register pair
#pragma unroll
                for (int j = 0; j < NC; j += 2) {
                    const float2 a2 = make_float2(al[j], al[j + 1]);
                    acc[j] = __ffma2_rn(make_float2(a2.x, a2.x), xa, acc[j]);
                    acc[j] = __ffma2_rn(make_float2(a2.y, a2.y), xb, acc[j]);
                    acc[j] = __ffma2_rn(make_float2(a2.x, a2.x), dp, acc[j]);
                    acc[j + 1] = __ffma2_rn(make_float2(a2.y, a2.y), xa, acc[j + 1]);
                    acc[j + 1] = __ffma2_rn(make_float2(a2.x, a2.x), xb, acc[j + 1]);
                    acc[j + 1] = __ffma2_rn(make_float2(a2.y, a2.y), dp, acc[j + 1]);
                }
            } else if (ORDER == 3) {          // halves of 16: blend 16, accumulate 32, twice
#pragma unroll
                for (int h = 0; h < NC; h += 16) {
                    float2 bp[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) bp[j] = __ffma2_rn(make_float2(al[h + j], al[h + j]), dp, r0);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        acc[h + j] = __ffma2_rn(make_float2(bp[j].x, bp[j].x), xa, acc[h + j]);
                        acc[h + j] = __ffma2_rn(make_float2(bp[j].y, bp[j].y), xb, acc[h + j]);
                    }
                }
            } else if (ORDER == 4) {          // blend by candidate pair: (b_j, b_j+1) for one position, alpha pair long-lived
#pragma unroll
                for (int j = 0; j < NC; j += 2) {
                    const float2 a2 = make_float2(al[j], al[j + 1]);
                    const float2 b0 = __ffma2_rn(a2, make_float2(dp.x, dp.x), make_float2(r0.x, r0.x));   // position p, candidates j, j+1
                    const float2 b1 = __ffma2_rn(a2, make_float2(dp.y, dp.y), make_float2(r0.y, r0.y));   // position p+1
                    acc[j] = __ffma2_rn(make_float2(b0.x, b0.x), xa, acc[j]);
                    acc[j] = __ffma2_rn(make_float2(b1.x, b1.x), xb, acc[j]);
                    acc[j + 1] = __ffma2_rn(make_float2(b0.y, b0.y), xa, acc[j + 1]);
                    acc[j + 1] = __ffma2_rn(make_float2(b1.y, b1.y), xb, acc[j + 1]);
                }
            } else {                          // accumulate only (the 2.03-cycle pattern), 3 per candidate
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    acc[j] = __ffma2_rn(make_float2(al[j], al[j]), xa, acc[j]);
                    acc[j] = __ffma2_rn(make_float2(al[j], al[j]), xb, acc[j]);
                    acc[j] = __ffma2_rn(make_float2(al[j], al[j]), dp, acc[j]);
                }
            }
        }
    }
    float s = 0.f;
    for (int j = 0; j < NC; ++j) s += acc[j].x + acc[j].y;
    if (s == 1234.5678f) out[0] = s;
}

template <int NC, int ORDER>
void run(const char* name, const float* in, float* out) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_pat<NC, ORDER><<<p.multiProcessorCount, 256>>>(in, out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ffma2 = (double)iters * 4 * NC * 3 * 2;          // per sub-partition: 2 warps
    printf("%-44s NC %2d: %8.3f ms, %.3f cycles per FFMA2 per SMSP\n", name, NC, best, best * 1e-3 * clk * 1e3 / ffma2);
}

int main() {
    float *in, *out; cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 64);
    float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = (i >= 4000) ? 0.f : 0.5f + 1e-3f * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<32, 0>("blend + 2 accumulates (k_brute pattern)", in, out);
    run<32, 1>("3 accumulates", in, out);
    run<32, 2>("3 accumulates, scalar from a pair half", in, out);
    run<32, 3>("blend 16 / accumulate 32, twice", in, out);
    run<32, 4>("blend by candidate pair", in, out);
    run<16, 0>("blend + 2 accumulates", in, out);
    run<16, 1>("3 accumulates", in, out);
    run<24, 0>("blend + 2 accumulates", in, out);
    run<8, 0>("blend + 2 accumulates", in, out);
    return 0;
}
