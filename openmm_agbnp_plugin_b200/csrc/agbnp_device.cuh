// Shared device-side definitions for the sm_100a AGBNP1/GaussVol kernels.
//
// HBM layout (all arrays in the library's internal "sorted" atom order: heavy atoms first in Morton order, padded to a
// multiple of 32, then hydrogens in Morton order, padded to a multiple of 32):
//   posq      float4[np]   x,y,z (nm), charge                     (SoA float4, one 16-byte load per atom)
//   orig      int[np]      caller's atom index, -1 for padding
//   aL,vL,aS,vS double[np] Gaussian exponent / volume for enlarged and vdW radii (v = 0 for hydrogens / padding)
//   accumulators (zeroed per evaluation by one memset): see Accum
#ifndef AGBNP_DEVICE_CUH_
#define AGBNP_DEVICE_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

namespace agbnp_b200_impl {

constexpr int TILE = 32;                      // atoms per block (one warp lane each)
constexpr double FORCE_SCALE = 4294967296.0;  // 2^32, OpenMM's fixed-point force scale
constexpr unsigned FULL = 0xffffffffu;

// status bits written by kernels when a capacity is exceeded (host grows the buffer and re-runs the evaluation)
enum StatusBits { ST_NBR_OVERFLOW = 1, ST_NODE_OVERFLOW = 2, ST_LEVEL_OVERFLOW = 4, ST_TREE_OVERFLOW = 8, ST_PAIRLIST_OVERFLOW = 16,
                  ST_TREE_STALE = 32 /* a rescan found no valid stored tree (its build evaluation overflowed): rebuild */,
                  ST_PEER_OVERFLOW = 64 /* sharded evaluation: another shard overflowed, nothing was delivered here either */,
                  ST_PEER_TIMEOUT = 128 /* peer-memory exchange: a peer's flag did not arrive within the spin limit */ };

// energy / diagnostic scalar slots (double)
enum ScalarSlot { SC_EVOL_L = 0, SC_EVOL_S = 1, SC_EGB = 2, SC_EVDW = 3, SC_VOL_L = 4, SC_VOL_S = 5, SC_TOTAL = 6, SC_FAULT = 7, SC_COUNT = 8 };
// work counters (unsigned long long)
enum CounterSlot { CT_PGB = 0, CT_PQ = 1, CT_C2 = 2, CT_C3 = 3, CT_M = 4, CT_TILES_GB = 5, CT_TILES_Q = 6, CT_SPARE = 7, CT_COUNT = 8 };

// Programmatic dependent launch (sm_90+): every kernel of an evaluation lets its successor be scheduled at once
// (pdl_release, first instruction) and itself waits for its predecessor -- completion AND memory visibility of the whole
// grid -- before it touches anything an earlier kernel reads or writes (pdl_acquire).  Launch latency and prologues that
// read only per-context constants (spline tables) hide behind the predecessor's tail.  Without the launch attribute both
// are no-ops.  Every CTA executes pdl_acquire, so completion stays transitive along the kernel chain.
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_acquire() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Work stealing of the persistent kernels, two flavours.
// (a) claim_unit: every unit comes from the atomic counter.
// (b) first_unit / next_unit: a warp's first unit is its own global index, later ones come from the counter offset by the
//     number of warps -- no burst of thousands of same-address atomics at the start of the kernel (20 % of k_born's stall
//     samples).  Measured: (b) helps the range-limited pair passes and the tree build of small systems (RNase H: k_born
//     26 -> 21 us, k_deriv 30 -> 24 us; Trp-cage k_tree 57 -> 50 us) and is neutral for them on 18 k atoms; k_gb and the gamma
//     sweep lose 5-10 % with it (their warps then run in lockstep through equal-sized units, and the math-bound and the
//     latency-bound phases of the warps of a scheduler no longer overlap), so those two keep (a).
// Tried (r2r): claims issued one unit ahead of the work (the atomic's round trip shows as 10-20 % of the stall samples of
// k_born / k_deriv / k_gb) -- every kernel got SLOWER (k_gb 154 -> 159 us, 2clr k_born 23 -> 30 us): other warps already cover
// the wait, and a warp that holds its next unit early unbalances the tail.
// The counter starts at 0 for every evaluation; grids using (b) must be fully resident.
__device__ __forceinline__ int claim_unit(int* counter, int lane) {
    int u = 0;
    if (lane == 0) u = atomicAdd(counter, 1);
    return __shfl_sync(0xffffffffu, u, 0);
}
__device__ __forceinline__ int first_unit() { return (int) (blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5)); }
__device__ __forceinline__ int next_unit(int* counter, int lane) { return claim_unit(counter, lane) + (int) (gridDim.x*(blockDim.x >> 5)); }

#ifdef TAIL_DEBUG
// Diagnostic build (tools/build_variant.sh taildbg -DTAIL_DEBUG; tools/tail_probe.py): every warp of a persistent kernel stamps
// %globaltimer when it enters and leaves its work loop -- the mean leaving time against the kernel's span says how much of the
// kernel is tail (r2ag: k_gb 98 %, k_tree 89 %, k_deriv 87 %, k_born 81 % on HIV-RT; 58-67 % for the tree and Born kernels of 2clr).
// per kernel slot k: [0] min start, [1] max end, [2] sum of warp end times, [3] number of warps, [4] sum of squares (unused)
__device__ unsigned long long g_tail[16][8];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t & ((1ull << 40)-1); }
__device__ __forceinline__ void tail_begin(int k) { if ((threadIdx.x & 31) == 0) atomicMin(&g_tail[k][0], gtimer()); }
__device__ __forceinline__ void tail_end(int k) {
    if ((threadIdx.x & 31) == 0) { const unsigned long long t = gtimer(); atomicMax(&g_tail[k][1], t); atomicAdd(&g_tail[k][2], t); atomicAdd(&g_tail[k][3], 1ull); }
}
#else
__device__ __forceinline__ void tail_begin(int) {}
__device__ __forceinline__ void tail_end(int) {}
#endif
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(FULL, v, o);
        if (lane_id() >= o) v += t;
    }
    return v;
}

// fixed-point force accumulation (bitwise reproducible; same representation as OpenMM's CUDA force buffer)
__device__ __forceinline__ void add_force_fixed(unsigned long long* f, float v) {
    atomicAdd(f, (unsigned long long) (long long) (v*(float) 4294967296.0f));
}
__device__ __forceinline__ void add_force_fixed(unsigned long long* f, double v) {
    atomicAdd(f, (unsigned long long) (long long) (v*FORCE_SCALE));
}

// squared distance between two axis-aligned boxes given as centre/half-extent
__device__ __forceinline__ float box_box_dist2(float4 ca, float4 ha, float4 cb, float4 hb) {
    float dx = fmaxf(0.f, fabsf(ca.x-cb.x) - ha.x - hb.x);
    float dy = fmaxf(0.f, fabsf(ca.y-cb.y) - ha.y - hb.y);
    float dz = fmaxf(0.f, fabsf(ca.z-cb.z) - ha.z - hb.z);
    return dx*dx + dy*dy + dz*dz;
}
__device__ __forceinline__ float point_box_dist2(float x, float y, float z, float4 cb, float4 hb) {
    float dx = fmaxf(0.f, fabsf(x-cb.x) - hb.x);
    float dy = fmaxf(0.f, fabsf(y-cb.y) - hb.y);
    float dz = fmaxf(0.f, fabsf(z-cb.z) - hb.z);
    return dx*dx + dy*dy + dz*dz;
}

// r2 exactly as the oracle's membership rule evaluates it: float, left to right, no FMA contraction
__device__ __forceinline__ float dist2_exact(float dx, float dy, float dz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

} // namespace agbnp_b200_impl
#endif
