// Pair-pass kernels of the AGBNP1 path for sm_100a: atom gather + block bounding boxes, inverse Born radii (S4-S5),
// GB pair energy/force + Y (S6), per-atom vdW / self terms (S7-S8), Born-radius derivative pass (S9), force scatter.
// Reference semantics: platforms/reference/src/ReferenceAGBNPKernels.cpp:421-586 (double loops over all pairs).
// Decomposition here: atoms in blocks of 32 (sorted order, heavy first); passes bounded by the 2.0 nm table range or
// by the cutoff cull block pairs on the fly with bounding boxes (no stored neighbor list); the Born and derivative
// passes are written as pure row sums (every per-atom output has one owner, no atomics on the hot side); the GB pass,
// which has no range limit without a cutoff, uses symmetric 32x32 register tiles (4 i-atoms x 8 j-atoms per lane),
// warp-shuffle reduce-scatter of the partial sums and one fixed-point atomic per atom and component.
#ifndef AGBNP_PAIR_CUH_
#define AGBNP_PAIR_CUH_

#include "agbnp_device.cuh"

namespace agbnp_b200_impl {

constexpr int PAIR_THREADS = 256;       // Born / derivative kernels: 8 warps share one row block
constexpr int PAIR_WARPS = PAIR_THREADS/32;
constexpr int GB_THREADS = 128;
constexpr int GB_CHUNK = 8;             // column tiles per GB work unit
constexpr int I4_INTERVALS = 15;        // AGBNP_I4LOOKUP_NA - 1
constexpr float PIFAC = 0.07957747154594767f;   // 1/(4 pi)

// ---------------------------------------------------------------------------------------------------------------
// k_prep: gather caller-order positions into the sorted SoA float4 array and compute block bounding boxes
// ---------------------------------------------------------------------------------------------------------------
struct PrepArgs {
    int np;
    const float4* posq_in;      // caller order (w ignored)
    const int* orig;            // sorted -> caller index, -1 padding
    const float* charge;        // sorted
    float4* posq;               // sorted out
    float4 *bbc, *bbh;
};

__global__ void __launch_bounds__(256) k_prep(PrepArgs A) {
    const int lane = threadIdx.x & 31;
    const int blk = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
    if (blk*TILE >= A.np) return;
    const int k = blk*TILE+lane;
    const int o = A.orig[k];
    float4 p;
    float lo[3], hi[3];
    if (o >= 0) {
        p = A.posq_in[o];
        p.w = A.charge[k];
        lo[0] = hi[0] = p.x; lo[1] = hi[1] = p.y; lo[2] = hi[2] = p.z;
    } else {
        // padding: far away, pairwise distinct, zero charge -- contributes exactly nothing anywhere
        p = make_float4(1.0e4f + 50.f*lane, 1.0e4f + 50.f*(blk % 1000), 1.0e4f + 50.f*(blk/1000), 0.f);
        lo[0] = lo[1] = lo[2] = 3.0e38f; hi[0] = hi[1] = hi[2] = -3.0e38f;
    }
    A.posq[k] = p;
#pragma unroll
    for (int c = 0; c < 3; c++) { lo[c] = warp_min(lo[c]); hi[c] = warp_max(hi[c]); }
    if (lane == 0) {
        A.bbc[blk] = make_float4(0.5f*(lo[0]+hi[0]), 0.5f*(lo[1]+hi[1]), 0.5f*(lo[2]+hi[2]), 0.f);
        A.bbh[blk] = make_float4(0.5f*(hi[0]-lo[0]), 0.5f*(hi[1]-lo[1]), 0.5f*(hi[2]-lo[2]), 0.f);
    }
}

// cubic-spline value / derivative from a packed interval (y_k, y_{k+1}, y2_k h^2/6, y2_{k+1} h^2/6); b = fraction in [0,1)
__device__ __forceinline__ float spline_value(float4 c, float b) {
    const float a = 1.f-b;
    return a*c.x + b*c.y + (a*a*a-a)*c.z + (b*b*b-b)*c.w;
}
__device__ __forceinline__ float spline_deriv(float4 c, float b, float inv_h) {
    const float a = 1.f-b;
    return ((c.y-c.x) + ((1.f-3.f*a*a)*c.z + (3.f*b*b-1.f)*c.w))*inv_h;
}

struct PairCommon {
    int np, nhb, nb;            // padded atoms, heavy blocks, all blocks
    const float4* posq;
    const int* orig;
    const float4 *bbc, *bbh;
    const unsigned char* ts;    // screened radius type
    const signed char* tj;      // screener radius type, -1 hydrogens / padding
    const float4* i4;           // packed tables
    int ntj, ntables;
    float inv_h;
    float range2;               // (2.0 nm)^2 table range
    float cut2;                 // cutoff^2 (float product), used when CUTOFF
    int row_begin, row_end;     // row blocks this shard owns
};

// ---------------------------------------------------------------------------------------------------------------
// k_born: beta_i = 1/r_i - (1/4pi) sum_{j heavy, j != i, d < 2.0} s_j Q(d; type_i, type_j); B_i = 1/swf(beta_i)
// (ReferenceAGBNPKernels.cpp:41-55,421-454) + the per-atom GB self energy (:477), vdW energy (:513-517), brw (:524-528)
// ---------------------------------------------------------------------------------------------------------------
struct BornArgs {
    PairCommon c;
    const double* svS;          // self volumes (vdW radii)
    const double* vS;           // atomic volumes (vdW radii)
    const float* radius;        // sorted vdW radii
    const float* alpha;         // sorted vdW alpha
    float* vsf;                 // out: volume scaling factors s_i
    float* born;                // out: B_i
    float* bfp;                 // out: d swf / d beta
    float* brw;                 // out
    float4* gbj;                // out [3*np]: GB atom records in broadcast form (see k_gb)
    float qscale;               // sqrt(-2k)
    double* scalars;
    unsigned long long* counters;
    float kdiel;                // dielectric_factor
    float hb_radius;
    int own_row_begin, own_row_end;   // rows whose per-atom energies / counters this shard reports (the pass itself is replicated)
};

template <bool CUTOFF>
__global__ void __launch_bounds__(PAIR_THREADS) k_born(BornArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* tab = (float4*) smem_raw;                                   // [ntables*15]
    float4* s_pos = tab + A.c.ntables*I4_INTERVALS;                     // [PAIR_WARPS][32]  x,y,z, s_j/(4pi)
    int* s_tj = (int*) (s_pos + PAIR_WARPS*TILE);                       // [PAIR_WARPS][32]
    float* s_red = (float*) (s_tj + PAIR_WARPS*TILE);                   // [PAIR_WARPS][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < A.c.ntables*I4_INTERVALS; i += blockDim.x) tab[i] = A.c.i4[i];
    __syncthreads();

    const int rowb = A.c.row_begin + blockIdx.x;
    const int a = rowb*TILE+lane;
    const float4 pa = A.c.posq[a];
    const int tbase = (int) A.c.ts[a]*A.c.ntj*I4_INTERVALS;
    const float4 ca = A.c.bbc[rowb], ha = A.c.bbh[rowb];
    const float lim2 = CUTOFF ? fminf(A.c.range2, A.c.cut2) : A.c.range2;
    float sum = 0.f;
    unsigned long long npair = 0;
    float4* my_pos = s_pos + warp*TILE;
    int* my_tj = s_tj + warp*TILE;
    for (int cb = warp; cb < A.c.nhb; cb += PAIR_WARPS) {
        if (box_box_dist2(ca, ha, A.c.bbc[cb], A.c.bbh[cb]) >= lim2) continue;      // warp-uniform
        const int j = cb*TILE+lane;
        float4 pj = A.c.posq[j];
        const double vj = A.vS[j];
        pj.w = vj > 0 ? PIFAC*((float) A.svS[j]/(float) vj) : 0.f;
        __syncwarp();
        my_pos[lane] = pj;
        my_tj[lane] = A.c.tj[j];
        __syncwarp();
#pragma unroll 4
        for (int jj = 0; jj < TILE; jj++) {
            const float4 q = my_pos[jj];
            const int tj = my_tj[jj];
            const float dx = q.x-pa.x, dy = q.y-pa.y, dz = q.z-pa.z;
            float d2;
            bool ok;
            if (CUTOFF) { d2 = dist2_exact(dx, dy, dz); ok = d2 < A.c.cut2 && d2 < A.c.range2; }
            else { d2 = dx*dx + dy*dy + dz*dz; ok = d2 < A.c.range2; }
            ok = ok && (cb*TILE+jj != a) && tj >= 0;
            if (ok) {
                const float d = d2*rsqrtf(fmaxf(d2, 1e-20f));
                const float t = d*A.c.inv_h;
                const int k = min((int) t, I4_INTERVALS-1);
                const float4 c = tab[tbase + tj*I4_INTERVALS + k];
                sum += q.w*spline_value(c, t-(float) k);
                npair++;
            }
        }
    }
    s_red[warp*TILE+lane] = sum;
    __syncthreads();
    if (warp == 0) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < PAIR_WARPS; w++) tot += s_red[w*TILE+lane];
        float evdw = 0.f, eself = 0.f;
        const bool real = A.c.orig[a] >= 0;
        if (real) {
            const float beta = 1.f/A.radius[a] - tot;
            // agbnp_swf_invbr (ReferenceAGBNPKernels.cpp:41-55)
            const float ia = 0.5f, ia2 = 0.25f;          // 1/2.0, 1/2.0^2
            float t, fp;
            if (beta < 0.f) { t = ia; fp = 0.f; }
            else { t = sqrtf(ia2 + beta*beta); fp = beta/t; }
            const float br = 1.f/t;
            A.born[a] = br;
            A.bfp[a] = fp;
            const double va = A.vS[a];
            A.vsf[a] = va > 0 ? (float) A.svS[a]/(float) va : 0.f;
            const float q = pa.w, al = A.alpha[a];
            eself = A.kdiel*q*q/br;
            const float bh = br + A.hb_radius;
            const float bh3 = bh*bh*bh;
            evdw = al/bh3;
            A.brw[a] = -PIFAC*3.f*al*br*br*fp/(bh3*bh);
        } else {
            A.born[a] = 1.f; A.bfp[a] = 0.f; A.vsf[a] = 0.f; A.brw[a] = 0.f;
        }
        {
            const float br = A.born[a];
            const float qs = pa.w*A.qscale, ib = 0.60056120439322491f/br;   // sqrt(log2(e)/4)/B
            float4* rec = A.gbj + 3*(size_t) a;
            rec[0] = make_float4(pa.x, pa.x, pa.y, pa.y);
            rec[1] = make_float4(pa.z, pa.z, qs, qs);
            rec[2] = make_float4(br, br, ib, ib);
        }
        const double es = warp_sum((double) eself), ev = warp_sum((double) evdw);
        if (lane == 0 && rowb >= A.own_row_begin && rowb < A.own_row_end) {
            atomicAdd(&A.scalars[SC_EGB], es); atomicAdd(&A.scalars[SC_EVDW], ev);
        }
    }
    npair = (unsigned long long) warp_sum((double) npair);
    if (lane == 0 && npair && rowb >= A.own_row_begin && rowb < A.own_row_end) atomicAdd(&A.counters[CT_PQ], npair);
}

// ---------------------------------------------------------------------------------------------------------------
// k_gb: GB pair energy, direct force and Y accumulators over symmetric 32x32 tiles
// (ReferenceAGBNPKernels.cpp:476-498), in packed FP32x2 arithmetic (fma.rn.f32x2 -> FFMA2 on sm_100a).
//
// The pass is bound by FP32 issue: one pair costs 27 FP32 operations + 2 MUFU (ex2, rsqrt).  Scalar code spends an
// issue slot per operation; here every arithmetic instruction handles the pairs (i0,j) and (i1,j) of two row atoms at
// once, which halves the issue slots of the FP32 part and leaves the FMA pipe itself as the limit.
//   lane grid 4 (li) x 8 (lj): a lane owns 8 row atoms (4 packed pairs) for a whole work unit and 4 column atoms per
//   tile; column atoms come from `gbj`, written by the Born kernel in broadcast form {x,x,y,y | z,z,q,q | B,B,ib,ib}
//   so that three 16-byte loads land directly in aligned register pairs;
//   column-side sums are reduce-scattered over the 4 lanes that share them (12 shuffles per tile) and leave as ONE
//   vector atomic (red.global.add.v4.f32) per lane and tile; row-side sums are flushed once per work unit.
// Charges are pre-scaled by sqrt(-2k), so q_i q_j carries the GB prefactor; ib = sqrt(log2(e)/4)/B, so that
// exp(-d2/(4 B_i B_j)) = ex2(-d2 ib_i ib_j).
// ---------------------------------------------------------------------------------------------------------------
struct GBArgs {
    PairCommon c;
    const float4* gbj;          // [3*np] broadcast-form GB atom records
    const int2* units;          // (row block, first column block): triangular cover in chunks of GB_CHUNK column tiles
    int nunits;
    int shard_rank, shard_count;
    float4* gbacc;              // out [np]: fx, fy, fz (GB pair force), Y_i*(-2k)
    double* scalars;
    unsigned long long* counters;
    int* work_counter;
};

__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// one 32x32 tile: 4 column atoms x 4 packed row pairs per lane
template <bool CUTOFF, bool DIAG>
__device__ __forceinline__ void gb_tile(const GBArgs& A, int cb, int li, int lj,
                                        const float2 (&nx)[4], const float2 (&ny)[4], const float2 (&nz)[4],
                                        const float2 (&qi)[4], const float2 (&bi)[4], const float2 (&nib)[4],
                                        float2 (&fi)[4][4], float2& e2, unsigned& npair) {
    const float2 m025 = make_float2(-0.25f, -0.25f), p025 = make_float2(0.25f, 0.25f);
    const float2 half = make_float2(-0.5f, -0.5f), three_half = make_float2(1.5f, 1.5f);
    const float2 cut2 = make_float2(A.c.cut2, A.c.cut2);
    float sj[4][4];
#pragma unroll
    for (int n = 0; n < 4; n++) {
        const int jl = lj*4 + n;                                  // column atom within the tile
        const float4* rec = A.gbj + 3*(size_t) (cb*TILE + jl);
        const float4 r0 = __ldg(rec), r1 = __ldg(rec+1), r2 = __ldg(rec+2);
        const float2 xj = make_float2(r0.x, r0.y), yj = make_float2(r0.z, r0.w), zj = make_float2(r1.x, r1.y);
        const float2 qj = make_float2(r1.z, r1.w), bj = make_float2(r2.x, r2.y), ibj = make_float2(r2.z, r2.w);
        float2 ax = make_float2(0.f, 0.f), ay = ax, az = ax, aY = ax;
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const float2 dx = __fadd2_rn(xj, nx[m]), dy = __fadd2_rn(yj, ny[m]), dz = __fadd2_rn(zj, nz[m]);
            float2 d2;
            if (CUTOFF) d2 = __fadd2_rn(__fadd2_rn(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy)), __fmul2_rn(dz, dz));   // membership rule: no contraction
            else d2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
            const float2 bb = __fmul2_rn(bi[m], bj);
            const float2 arg = __fmul2_rn(d2, __fmul2_rn(nib[m], ibj));
            const float2 et = make_float2(fast_ex2(arg.x), fast_ex2(arg.y));
            const float2 t = __ffma2_rn(bb, et, d2);
            const float2 f = make_float2(fast_rsqrt(t.x), fast_rsqrt(t.y));
            float2 qq = __fmul2_rn(qi[m], qj);
            if (DIAG) {                                           // pairs i < j only (also removes i == j)
                const int i0 = li*8 + 2*m;
                qq.x = i0 < jl ? qq.x : 0.f;
                qq.y = i0+1 < jl ? qq.y : 0.f;
            }
            if (CUTOFF) {
                const bool k0 = d2.x < cut2.x, k1 = d2.y < cut2.y;
                qq.x = k0 ? qq.x : 0.f;
                qq.y = k1 ? qq.y : 0.f;
                if (DIAG) npair += (k0 && li*8 + 2*m < jl) + (k1 && li*8 + 2*m+1 < jl);
                else npair += (unsigned) k0 + (unsigned) k1;
            }
            const float2 qf = __fmul2_rn(qq, f);
            const float2 ff = __fmul2_rn(f, f);
            // energy with one Newton step on the reciprocal square root: f (3/2 - t f^2 / 2); MUFU.RSQ alone carries a
            // ~4e-7 mean relative bias that the 1e8-term pair sum does not average out
            const float2 corr = __ffma2_rn(__fmul2_rn(t, ff), half, three_half);
            e2 = __ffma2_rn(qf, corr, e2);
            const float2 g = __fmul2_rn(qf, ff);
            const float2 hh = __fmul2_rn(g, et);
            const float2 mw = __ffma2_rn(m025, hh, g);            // -2 k q_i q_j (1 - e/4) f^3
            const float2 yt = __fmul2_rn(hh, __ffma2_rn(p025, d2, bb));
            fi[m][0] = __ffma2_rn(dx, mw, fi[m][0]); fi[m][1] = __ffma2_rn(dy, mw, fi[m][1]); fi[m][2] = __ffma2_rn(dz, mw, fi[m][2]);
            fi[m][3] = __fadd2_rn(fi[m][3], yt);
            ax = __ffma2_rn(dx, mw, ax); ay = __ffma2_rn(dy, mw, ay); az = __ffma2_rn(dz, mw, az);
            aY = __fadd2_rn(aY, yt);
        }
        sj[n][0] = -(ax.x+ax.y); sj[n][1] = -(ay.x+ay.y); sj[n][2] = -(az.x+az.y); sj[n][3] = aY.x+aY.y;
    }
    // reduce-scatter the column-side sums over the 4 lanes sharing lj (lane bits 3,4): 8 + 4 shuffles, after which
    // lane (li,lj) owns column atom lj*4 + li
    float h2[2][4], h1[4];
    {
        const bool up = li & 2;
#pragma unroll
        for (int n = 0; n < 2; n++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float send = up ? sj[n][c] : sj[n+2][c];
                const float keep = up ? sj[n+2][c] : sj[n][c];
                h2[n][c] = keep + __shfl_xor_sync(FULL, send, 16);
            }
    }
    {
        const bool up = li & 1;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float send = up ? h2[0][c] : h2[1][c];
            const float keep = up ? h2[1][c] : h2[0][c];
            h1[c] = keep + __shfl_xor_sync(FULL, send, 8);
        }
    }
    atomicAdd(&A.gbacc[cb*TILE + lj*4 + li], make_float4(h1[0], h1[1], h1[2], h1[3]));
}

template <bool CUTOFF>
__global__ void __launch_bounds__(GB_THREADS, 3) k_gb(GBArgs A) {
    const int lane = threadIdx.x & 31;
    const int li = lane >> 3, lj = lane & 7;
    double e_acc = 0.0;
    unsigned long long npair = 0, ntile = 0;
    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(A.work_counter, 1);
        u = __shfl_sync(FULL, u, 0);
        if (u >= A.nunits) break;
        if (A.shard_count > 1 && (u % A.shard_count) != A.shard_rank) continue;
        const int2 un = A.units[u];
        const int ra = un.x;
        const int cend = min(un.y+GB_CHUNK, A.c.nb);
        // row atoms: 4 packed pairs (even atom in .x, odd atom in .y); positions negated, ib negated
        float2 nx[4], ny[4], nz[4], qi[4], bi[4], nib[4], fi[4][4];
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const float4* r = A.gbj + 3*(size_t) (ra*TILE + li*8 + 2*m);
            const float4 a0 = __ldg(r), a1 = __ldg(r+1), a2 = __ldg(r+2), b0 = __ldg(r+3), b1 = __ldg(r+4), b2 = __ldg(r+5);
            nx[m] = make_float2(-a0.x, -b0.x); ny[m] = make_float2(-a0.z, -b0.z); nz[m] = make_float2(-a1.x, -b1.x);
            qi[m] = make_float2(a1.z, b1.z); bi[m] = make_float2(a2.x, b2.x); nib[m] = make_float2(-a2.z, -b2.z);
#pragma unroll
            for (int c = 0; c < 4; c++) fi[m][c] = make_float2(0.f, 0.f);
        }
        const float4 ca = A.c.bbc[ra], ha = A.c.bbh[ra];
        float2 e2 = make_float2(0.f, 0.f);
        unsigned np32 = 0;
        for (int cb = un.y; cb < cend; cb++) {
            if (CUTOFF && box_box_dist2(ca, ha, A.c.bbc[cb], A.c.bbh[cb]) >= A.c.cut2) continue;
            ntile++;
            if (cb == ra) gb_tile<CUTOFF, true>(A, cb, li, lj, nx, ny, nz, qi, bi, nib, fi, e2, np32);
            else {
                gb_tile<CUTOFF, false>(A, cb, li, lj, nx, ny, nz, qi, bi, nib, fi, e2, np32);
                if (!CUTOFF) np32 += 32;
            }
            if (!CUTOFF && cb == ra && lane < 16) np32 += 31;      // 496 = 16*31 pairs in a diagonal tile
        }
        e_acc += (double) e2.x + (double) e2.y;
        npair += np32;
        // row side: reduce-scatter over the 8 lanes sharing li (lane bits 0..2): 16 + 8 + 4 shuffles; lane (li,lj) ends
        // with row atom li*8 + lj
        float v4[4][4], v2[2][4], v1[4];
        {
            const bool up = lj & 4;
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    // atom a (0..3) and atom a+4: pair index a>>1 / (a+4)>>1, half a&1
                    const float lo = (a & 1) ? fi[a >> 1][c].y : fi[a >> 1][c].x;
                    const float hi = (a & 1) ? fi[(a+4) >> 1][c].y : fi[(a+4) >> 1][c].x;
                    const float send = up ? lo : hi;
                    const float keep = up ? hi : lo;
                    v4[a][c] = keep + __shfl_xor_sync(FULL, send, 4);
                }
        }
        {
            const bool up = lj & 2;
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float send = up ? v4[a][c] : v4[a+2][c];
                    const float keep = up ? v4[a+2][c] : v4[a][c];
                    v2[a][c] = keep + __shfl_xor_sync(FULL, send, 2);
                }
        }
        {
            const bool up = lj & 1;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float send = up ? v2[0][c] : v2[1][c];
                const float keep = up ? v2[1][c] : v2[0][c];
                v1[c] = keep + __shfl_xor_sync(FULL, send, 1);
            }
        }
        atomicAdd(&A.gbacc[ra*TILE + li*8 + lj], make_float4(v1[0], v1[1], v1[2], v1[3]));
    }
    e_acc = warp_sum(e_acc);
    npair = (unsigned long long) warp_sum((double) npair);
    if (lane == 0) {
        atomicAdd(&A.scalars[SC_EGB], -e_acc);
        atomicAdd(&A.counters[CT_PGB], npair);
        atomicAdd(&A.counters[CT_TILES_GB], ntile);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_bw: bw_i = brw_i + bru_i, bru_i = -(k/4pi)(q_i^2 + Y_i B_i) fp_i   (ReferenceAGBNPKernels.cpp:537-542); folds the
// GB pair force of the atoms this shard reports into the fixed-point force accumulator
// ---------------------------------------------------------------------------------------------------------------
struct BwArgs {
    int np;
    const float4* posq;
    const float4* gbacc;
    const float *born, *bfp, *brw;
    float kdiel;
    float* bw;
    unsigned long long* force;
    int own_begin, own_end;     // sorted-index range whose GB force this shard adds (gbacc is all-reduced before)
};

__global__ void __launch_bounds__(256) k_bw(BwArgs A) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= A.np) return;
    const float q = A.posq[i].w;
    const float4 g = A.gbacc[i];
    const float y = g.w/(-2.f*A.kdiel);
    A.bw[i] = A.brw[i] - PIFAC*A.kdiel*(q*q + y*A.born[i])*A.bfp[i];
    if (i >= A.own_begin && i < A.own_end) {
        // the only writer of these entries at this point of the stream: plain read-modify-write
        A.force[i] += (unsigned long long) (long long) (g.x*4294967296.0f);
        A.force[(size_t) A.np+i] += (unsigned long long) (long long) (g.y*4294967296.0f);
        A.force[2*(size_t) A.np+i] += (unsigned long long) (long long) (g.z*4294967296.0f);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_deriv: Born-radius derivative pass as a row sum (ReferenceAGBNPKernels.cpp:555-586).  For atom a and partner b
// (d < 2.0, b != a), with D = r_b - r_a:
//   F_a  += D/d [ heavy(b) bw_a s_b Q'(d; ts_a, tj_b)  +  heavy(a) bw_b s_a Q'(d; ts_b, tj_a) ]
//   WU_a += heavy(a) bw_b Q(d; ts_b, tj_a)                      (W and U merged: both are linear in brw / bru)
// which is the reference's ordered-pair loop regrouped by the atom that receives the contribution.
// ---------------------------------------------------------------------------------------------------------------
struct DerivArgs {
    PairCommon c;
    const float* vsf;
    const float* bw;
    float* wu;                  // out [np]
    unsigned long long* force;
    unsigned long long* counters;
};

template <bool CUTOFF>
__global__ void __launch_bounds__(PAIR_THREADS) k_deriv(DerivArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* tab = (float4*) smem_raw;
    float4* s_pos = tab + A.c.ntables*I4_INTERVALS;                     // x,y,z,s_b
    float2* s_ex = (float2*) (s_pos + PAIR_WARPS*TILE);                 // bw_b, packed types
    float4* s_red = (float4*) (s_ex + PAIR_WARPS*TILE);                 // [PAIR_WARPS][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < A.c.ntables*I4_INTERVALS; i += blockDim.x) tab[i] = A.c.i4[i];
    __syncthreads();

    const int rowb = A.c.row_begin + blockIdx.x;
    const int a = rowb*TILE+lane;
    const float4 pa = A.c.posq[a];
    const int ts_a = A.c.ts[a];
    const int tj_a = A.c.tj[a];
    const bool hv_a = tj_a >= 0;
    const float s_a = A.vsf[a], bw_a = A.bw[a];
    const int base_a = ts_a*A.c.ntj*I4_INTERVALS;
    const float4 ca = A.c.bbc[rowb], ha = A.c.bbh[rowb];
    const float lim2 = CUTOFF ? fminf(A.c.range2, A.c.cut2) : A.c.range2;
    // heavy rows see every column block (they descreen hydrogens too); hydrogen rows only heavy columns
    const int ncol = rowb < A.c.nhb ? A.c.nb : A.c.nhb;
    float fx = 0.f, fy = 0.f, fz = 0.f, wu = 0.f;
    float4* my_pos = s_pos + warp*TILE;
    float2* my_ex = s_ex + warp*TILE;
    for (int cb = warp; cb < ncol; cb += PAIR_WARPS) {
        if (box_box_dist2(ca, ha, A.c.bbc[cb], A.c.bbh[cb]) >= lim2) continue;
        const int j = cb*TILE+lane;
        float4 pj = A.c.posq[j];
        pj.w = A.vsf[j];
        const int pk = (int) A.c.ts[j] | (((int) A.c.tj[j] & 0xff) << 8);
        __syncwarp();
        my_pos[lane] = pj;
        my_ex[lane] = make_float2(A.bw[j], __int_as_float(pk));
        __syncwarp();
#pragma unroll 2
        for (int jj = 0; jj < TILE; jj++) {
            const float4 q = my_pos[jj];
            const float2 ex = my_ex[jj];
            const float dx = q.x-pa.x, dy = q.y-pa.y, dz = q.z-pa.z;
            float d2;
            bool ok;
            if (CUTOFF) { d2 = dist2_exact(dx, dy, dz); ok = d2 < A.c.cut2 && d2 < A.c.range2; }
            else { d2 = dx*dx + dy*dy + dz*dz; ok = d2 < A.c.range2; }
            ok = ok && (cb*TILE+jj != a);
            if (ok) {
                const int pk2 = __float_as_int(ex.y);
                const int ts_b = pk2 & 0xff;
                const int tj_b = (int) (signed char) ((pk2 >> 8) & 0xff);
                const float inv_d = rsqrtf(fmaxf(d2, 1e-20f));
                const float d = d2*inv_d;
                const float t = d*A.c.inv_h;
                const int k = min((int) t, I4_INTERVALS-1);
                const float fr = t-(float) k;
                float w = 0.f;
                if (tj_b >= 0) {                                    // b descreens a
                    const float4 c1 = tab[base_a + tj_b*I4_INTERVALS + k];
                    w = bw_a*q.w*spline_deriv(c1, fr, A.c.inv_h);
                }
                if (hv_a) {                                         // a descreens b
                    const float4 c2 = tab[(ts_b*A.c.ntj + tj_a)*I4_INTERVALS + k];
                    w += ex.x*s_a*spline_deriv(c2, fr, A.c.inv_h);
                    wu += ex.x*spline_value(c2, fr);
                }
                w *= inv_d;
                fx = fmaf(dx, w, fx); fy = fmaf(dy, w, fy); fz = fmaf(dz, w, fz);
            }
        }
    }
    s_red[warp*TILE+lane] = make_float4(fx, fy, fz, wu);
    __syncthreads();
    if (warp == 0) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < PAIR_WARPS; w++) {
            const float4 v = s_red[w*TILE+lane];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        A.wu[a] = t.w;
        if (t.x != 0.f || t.y != 0.f || t.z != 0.f) {
            add_force_fixed(&A.force[a], t.x);
            add_force_fixed(&A.force[(size_t) A.c.np+a], t.y);
            add_force_fixed(&A.force[2*(size_t) A.c.np+a], t.z);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_finish: scatter the fixed-point forces (sorted order) into the caller's sink and fold the energy terms
// ---------------------------------------------------------------------------------------------------------------
struct FinishArgs {
    int np, n;
    const int* orig;
    const unsigned long long* force;
    float* out_f32;                     // layout 0: float[3n] interleaved, +=
    unsigned long long* out_fixed;      // layout 1: OpenMM fixed point [3][padded_n], atomic +=
    double* out_f64;                    // internal: double[3n] interleaved, = (host path)
    int padded_n;
    double* scalars;                    // SC_* terms
    const int* status;                  // capacity-overflow bits of this evaluation: nothing is delivered unless 0
    double* energy_accum;               // optional device accumulator (+=)
    double* energy_out;                 // optional device/pinned-mapped slot (=)
};

__global__ void __launch_bounds__(256) k_finish(FinishArgs A) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (A.status && *A.status != 0) return;
    if (k == 0) {
        const double e = A.scalars[SC_EVOL_L] + A.scalars[SC_EVOL_S] + A.scalars[SC_EGB] + A.scalars[SC_EVDW];
        A.scalars[SC_SPARE0] = e;
        if (A.energy_out) *A.energy_out = e;
        if (A.energy_accum) atomicAdd(A.energy_accum, e);
    }
    if (k >= A.np) return;
    const int o = A.orig[k];
    if (o < 0) return;
    const long long fx = (long long) A.force[k], fy = (long long) A.force[(size_t) A.np+k], fz = (long long) A.force[2*(size_t) A.np+k];
    if (A.out_f64) {
        A.out_f64[3*o+0] = (double) fx/FORCE_SCALE; A.out_f64[3*o+1] = (double) fy/FORCE_SCALE; A.out_f64[3*o+2] = (double) fz/FORCE_SCALE;
    }
    if (A.out_f32) {
        A.out_f32[3*o+0] += (float) ((double) fx/FORCE_SCALE); A.out_f32[3*o+1] += (float) ((double) fy/FORCE_SCALE);
        A.out_f32[3*o+2] += (float) ((double) fz/FORCE_SCALE);
    }
    if (A.out_fixed) {
        atomicAdd(&A.out_fixed[o], (unsigned long long) fx);
        atomicAdd(&A.out_fixed[(size_t) A.padded_n+o], (unsigned long long) fy);
        atomicAdd(&A.out_fixed[2*(size_t) A.padded_n+o], (unsigned long long) fz);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_list_pairs: diagnostic -- every (i<j, caller indices) with float r2 < cutoff2, the membership rule of the pair passes
// ---------------------------------------------------------------------------------------------------------------
struct ListArgs {
    PairCommon c;
    int2* pairs;
    long long cap;
    unsigned long long* count;
};

__global__ void __launch_bounds__(256) k_list_pairs(ListArgs A) {
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x*blockDim.x) >> 5;
    const long long ntiles = (long long) A.c.nb*A.c.nb;
    for (long long t = wid; t < ntiles; t += nw) {
        const int ra = (int) (t / A.c.nb), cb = (int) (t % A.c.nb);
        if (cb < ra) continue;
        if (box_box_dist2(A.c.bbc[ra], A.c.bbh[ra], A.c.bbc[cb], A.c.bbh[cb]) >= A.c.cut2) continue;
        const int i = ra*TILE+lane;
        const float4 pi = A.c.posq[i];
        const int oi = A.c.orig[i];
        for (int jj = 0; jj < TILE; jj++) {
            const int j = cb*TILE+jj;
            const float4 pj = A.c.posq[j];
            const int oj = A.c.orig[j];
            bool ok = oi >= 0 && oj >= 0 && (cb > ra || jj > lane);
            // same operand order as the oracle: pos[hi] - pos[lo] in caller indices; r2 is symmetric in sign anyway
            const float dx = pj.x-pi.x, dy = pj.y-pi.y, dz = pj.z-pi.z;
            ok = ok && dist2_exact(dx, dy, dz) < A.c.cut2;
            if (ok) {
                const unsigned long long p = atomicAdd(A.count, 1ull);
                if ((long long) p < A.cap) A.pairs[p] = make_int2(min(oi, oj), max(oi, oj));
            }
        }
    }
}

} // namespace agbnp_b200_impl
#endif
