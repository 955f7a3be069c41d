// Issue-rate microbenchmarks for the roofline denominators of this path (FP32 FMA and SFU/MUFU issue, not HBM,
// bound the AGBNP1 pair passes).  MEASURED_PEAKS.json only carries HBM and bf16 figures, so bench.py measures these two
// on the GPU it runs on, at the clock the chip sustains under this kind of load, and reports fractions "of measured".
#ifndef AGBNP_PEAKS_CUH_
#define AGBNP_PEAKS_CUH_

#include <cuda_runtime.h>

namespace agbnp_b200_impl {

constexpr int PEAK_ITERS = 2048;
constexpr int PEAK_CHAINS = 8;

// mode 0: scalar FFMA, 1: packed fma.rn.f32x2 (FFMA2), 2: MUFU.EX2, 3: MUFU.RSQ, 4: GB-like mix (14 FFMA-class : 1 MUFU)
template <int MODE>
__global__ void __launch_bounds__(256) k_peak(float* out, float seed) {
    float a[PEAK_CHAINS], b = seed, c = 0.5f*seed;
    float2 v[PEAK_CHAINS];
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; i++) { a[i] = seed + (float) (threadIdx.x+i); v[i] = make_float2(a[i], a[i]+1.f); }
    const float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
#pragma unroll 1
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < PEAK_CHAINS; i++) {
            if (MODE == 0) a[i] = fmaf(a[i], b, c);
            else if (MODE == 1) v[i] = __ffma2_rn(v[i], b2, c2);
            else if (MODE == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            else if (MODE == 3) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            else {
                float t = a[i];
#pragma unroll
                for (int k = 0; k < 14; k++) t = fmaf(t, b, c);
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(t));
                a[i] = t;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; i++) s += a[i] + v[i].x + v[i].y;
    if (s == 12345.678f) out[0] = s;          // keeps the chains alive without a store on the timed path
}

// returns operations per second of the mode's unit: FMA lanes/s (modes 0,1), MUFU ops/s (2,3), instr-lanes/s (4)
template <int MODE>
inline double run_peak(int num_sm, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1, float* d_out) {
    const int grid = num_sm*8;
    k_peak<MODE><<<grid, 256, 0, s>>>(d_out, 1.0f);                  // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, s);
        k_peak<MODE><<<grid, 256, 0, s>>>(d_out, 1.0f);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double per_thread = (double) PEAK_ITERS*PEAK_CHAINS;
        if (MODE == 1) per_thread *= 2.0;
        if (MODE == 4) per_thread *= 15.0;
        const double rate = per_thread*256.0*grid/(ms*1e-3);
        if (rate > best) best = rate;
    }
    return best;
}

} // namespace agbnp_b200_impl
#endif
