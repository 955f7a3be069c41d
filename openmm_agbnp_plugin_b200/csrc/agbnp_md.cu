// libagbnp_md.so -- the integrator side of the MD driver used where OpenMM is absent: one fused kernel per
// Langevin step, so that "MD steps per second" measures the AGBNP evaluation and not a dozen element-wise launches.
// Not part of the plugin's force path: OpenMM's own integrators play this role there (example/hivrt_benchmark.py:20 uses
// LangevinIntegrator(300 K, 1/ps, 1 fs)).  The update is that integrator's: v <- a v + (1-a)/gamma F/m + sqrt(kT (1-a^2)/m) N(0,1),
// a = exp(-gamma dt); x <- x + v dt.
#include <cuda_runtime.h>
#include <math.h>

namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {          // splitmix64 finaliser
    z = (z ^ (z >> 30))*0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27))*0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// two independent standard normals from a counter (seed, step, index): Box-Muller on two 24-bit uniforms
__device__ __forceinline__ float2 normal2(unsigned long long seed, unsigned long long step, unsigned long long idx) {
    const unsigned long long h = mix64(seed ^ mix64(step*0x9e3779b97f4a7c15ull + idx));
    const float u1 = ((float) ((h >> 40) & 0xffffff) + 0.5f)*(1.0f/16777216.0f);
    const float u2 = ((float) ((h >> 8) & 0xffffff) + 0.5f)*(1.0f/16777216.0f);
    const float r = sqrtf(-2.0f*__logf(u1));
    float s, c;
    __sincosf(6.283185307179586f*u2, &s, &c);
    return make_float2(r*c, r*s);
}

struct StepArgs {
    float4* posq;            // [n] x, y, z, (charge slot untouched)
    float* vel;              // [3n]
    float* frc;              // [3n] force of the step's evaluation; zeroed here for the next one
    const float* inv_mass;   // [n]
    const float4* x0;        // [n] tether positions, or null
    float k_res;             // kJ/mol/nm^2
    int n;
    float dt, vscale, fscale, noise;    // ps; a; (1-a)/gamma; sqrt(kT (1-a^2))
    unsigned long long seed, step;
    double* ke2;             // optional: sum of m v^2 after the update (kJ/mol), accumulated
};

__global__ void __launch_bounds__(256) k_langevin(StepArgs A) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    double ke = 0.0;
    if (i < A.n) {
        float4 p = A.posq[i];
        float fx = A.frc[3*i], fy = A.frc[3*i+1], fz = A.frc[3*i+2];
        A.frc[3*i] = 0.f; A.frc[3*i+1] = 0.f; A.frc[3*i+2] = 0.f;
        if (A.x0) { const float4 q = A.x0[i]; fx -= A.k_res*(p.x-q.x); fy -= A.k_res*(p.y-q.y); fz -= A.k_res*(p.z-q.z); }
        const float im = A.inv_mass[i];
        const float sd = A.noise*sqrtf(im);
        const float2 n01 = normal2(A.seed, A.step, 2ull*i), n2 = normal2(A.seed, A.step, 2ull*i+1);
        float vx = A.vel[3*i], vy = A.vel[3*i+1], vz = A.vel[3*i+2];
        vx = A.vscale*vx + A.fscale*im*fx + sd*n01.x;
        vy = A.vscale*vy + A.fscale*im*fy + sd*n01.y;
        vz = A.vscale*vz + A.fscale*im*fz + sd*n2.x;
        A.vel[3*i] = vx; A.vel[3*i+1] = vy; A.vel[3*i+2] = vz;
        p.x += A.dt*vx; p.y += A.dt*vy; p.z += A.dt*vz;
        A.posq[i] = p;
        ke = (double) ((vx*vx + vy*vy + vz*vz)/im);
    }
    if (A.ke2) {
        for (int o = 16; o > 0; o >>= 1) ke += __shfl_xor_sync(0xffffffffu, ke, o);
        if ((threadIdx.x & 31) == 0 && ke != 0.0) atomicAdd(A.ke2, ke);
    }
}

} // namespace

extern "C" int agbnp_md_langevin_step(void* posq, float* vel, float* frc, const float* inv_mass, const void* x0, float k_res, int n,
                                      float dt_ps, float friction_per_ps, float kT, unsigned long long seed, unsigned long long step,
                                      double* d_ke2, void* stream) {
    if (!posq || !vel || !frc || !inv_mass || n <= 0) return -1;
    StepArgs a;
    a.posq = (float4*) posq; a.vel = vel; a.frc = frc; a.inv_mass = inv_mass; a.x0 = (const float4*) x0; a.k_res = k_res; a.n = n;
    a.dt = dt_ps;
    const double vs = exp(-(double) friction_per_ps*dt_ps);
    a.vscale = (float) vs;
    a.fscale = (float) (friction_per_ps > 0 ? (1.0-vs)/friction_per_ps : dt_ps);
    a.noise = (float) sqrt((double) kT*(1.0-vs*vs));
    a.seed = seed; a.step = step; a.ke2 = d_ke2;
    k_langevin<<<(n+255)/256, 256, 0, (cudaStream_t) stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
