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
constexpr int GB_CHUNK = 16;            // column tiles per GB work unit
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
// (ReferenceAGBNPKernels.cpp:476-498).  Charges are pre-scaled by sqrt(-2k) so that q_i q_j carries the GB prefactor.
// ---------------------------------------------------------------------------------------------------------------
struct GBArgs {
    PairCommon c;
    const float* born;
    const int2* units;          // (row block, first column block); NoCutoff: precomputed triangular cover
    int nunits;
    int shard_rank, shard_count;
    float qscale;               // sqrt(-2k)
    float* yq;                  // out: sum_j (-2k q_i q_j)(bb + d2/4) e f^3
    unsigned long long* force;
    double* scalars;
    unsigned long long* counters;
    int* work_counter;
};

struct GBAtom { float x, y, z, q, b, ib; };

__device__ __forceinline__ GBAtom gb_load(const float4* posq, const float* born, int idx, float qscale) {
    const float4 p = posq[idx];
    const float b = born[idx];
    GBAtom r;
    r.x = p.x; r.y = p.y; r.z = p.z; r.q = p.w*qscale; r.b = b;
    r.ib = 0.60056120439322491f/b;        // sqrt(0.25*log2(e))/B : ib_i*ib_j*d2 = d2/(4 B_i B_j) in base-2 exponent units
    return r;
}

template <bool CUTOFF>
__global__ void __launch_bounds__(GB_THREADS) k_gb(GBArgs A) {
    const int lane = threadIdx.x & 31;
    const int li = lane >> 2, lj = lane & 3;          // 8 x 4 lane grid: 4 i-atoms (li*4..) x 8 j-atoms (lj*8..) per lane
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
        GBAtom ai[4];
        float fi[4][4];                                // fx, fy, fz, Y per i-atom
#pragma unroll
        for (int m = 0; m < 4; m++) {
            ai[m] = gb_load(A.c.posq, A.born, ra*TILE + li*4+m, A.qscale);
            fi[m][0] = fi[m][1] = fi[m][2] = fi[m][3] = 0.f;
        }
        const float4 ca = A.c.bbc[ra], ha = A.c.bbh[ra];
        for (int cb = un.y; cb < cend; cb++) {
            if (CUTOFF && box_box_dist2(ca, ha, A.c.bbc[cb], A.c.bbh[cb]) >= A.c.cut2) continue;
            ntile++;
            const bool diag = cb == ra;
            float fj[8][4];
            float e_tile = 0.f;
#pragma unroll
            for (int n = 0; n < 8; n++) {
                const GBAtom aj = gb_load(A.c.posq, A.born, cb*TILE + lj*8+n, A.qscale);
                fj[n][0] = fj[n][1] = fj[n][2] = fj[n][3] = 0.f;
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const float dx = aj.x-ai[m].x, dy = aj.y-ai[m].y, dz = aj.z-ai[m].z;
                    float d2;
                    bool ok = true;
                    if (CUTOFF) { d2 = dist2_exact(dx, dy, dz); ok = d2 < A.c.cut2; }
                    else d2 = dx*dx + dy*dy + dz*dz;
                    if (diag) ok = ok && (li*4+m < lj*8+n);
                    const float bb = ai[m].b*aj.b;
                    const float et = exp2f(-d2*(ai[m].ib*aj.ib));
                    const float fgb = rsqrtf(fmaf(bb, et, d2));
                    float qq = ai[m].q*aj.q;                      // -2k q_i q_j
                    if (CUTOFF || diag) qq = ok ? qq : 0.f;
                    e_tile = fmaf(qq, fgb, e_tile);               // -(pair energy)
                    const float f3 = fgb*fgb*fgb;
                    const float qf3 = qq*f3;
                    const float mw = qf3*fmaf(-0.25f, et, 1.f);   // = -2 k q_i q_j (1 - e/4) f^3 with the sign folded
                    fi[m][0] = fmaf(dx, mw, fi[m][0]); fi[m][1] = fmaf(dy, mw, fi[m][1]); fi[m][2] = fmaf(dz, mw, fi[m][2]);
                    fj[n][0] = fmaf(-dx, mw, fj[n][0]); fj[n][1] = fmaf(-dy, mw, fj[n][1]); fj[n][2] = fmaf(-dz, mw, fj[n][2]);
                    const float yt = qf3*et*fmaf(0.25f, d2, bb);
                    fi[m][3] += yt; fj[n][3] += yt;
                    if (CUTOFF || diag) npair += ok ? 1 : 0;
                }
            }
            if (!CUTOFF && !diag) npair += 32;
            e_acc += (double) e_tile;
            // reduce-scatter the j-side partial sums over the 8 lanes sharing lj: 16 + 8 + 4 shuffles, after which
            // lane (li,lj) owns the 4 sums of j-atom lj*8+li
            float h4[4][4], h2[2][4], h1[4];
            {
                const bool up = li & 4;
#pragma unroll
                for (int n = 0; n < 4; n++)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const float send = up ? fj[n][c] : fj[n+4][c];
                        const float keep = up ? fj[n+4][c] : fj[n][c];
                        h4[n][c] = keep + __shfl_xor_sync(FULL, send, 16);
                    }
            }
            {
                const bool up = li & 2;
#pragma unroll
                for (int n = 0; n < 2; n++)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const float send = up ? h4[n][c] : h4[n+2][c];
                        const float keep = up ? h4[n+2][c] : h4[n][c];
                        h2[n][c] = keep + __shfl_xor_sync(FULL, send, 8);
                    }
            }
            {
                const bool up = li & 1;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float send = up ? h2[0][c] : h2[1][c];
                    const float keep = up ? h2[1][c] : h2[0][c];
                    h1[c] = keep + __shfl_xor_sync(FULL, send, 4);
                }
            }
            {
                const int j = cb*TILE + lj*8 + li;      // bits of li select 4/2/1 -> atom index li within the lj group
                if (h1[0] != 0.f || h1[1] != 0.f || h1[2] != 0.f) {
                    add_force_fixed(&A.force[j], h1[0]);
                    add_force_fixed(&A.force[(size_t) A.c.np+j], h1[1]);
                    add_force_fixed(&A.force[2*(size_t) A.c.np+j], h1[2]);
                }
                if (h1[3] != 0.f) atomicAdd(&A.yq[j], h1[3]);
            }
        }
        // i-side: reduce-scatter over the 4 lanes sharing li (8 + 4 shuffles); lane (li,lj) ends with i-atom li*4+lj
        float g2[2][4], g1[4];
        {
            const bool up = lj & 2;
#pragma unroll
            for (int m = 0; m < 2; m++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float send = up ? fi[m][c] : fi[m+2][c];
                    const float keep = up ? fi[m+2][c] : fi[m][c];
                    g2[m][c] = keep + __shfl_xor_sync(FULL, send, 2);
                }
        }
        {
            const bool up = lj & 1;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float send = up ? g2[0][c] : g2[1][c];
                const float keep = up ? g2[1][c] : g2[0][c];
                g1[c] = keep + __shfl_xor_sync(FULL, send, 1);
            }
        }
        {
            const int i = ra*TILE + li*4 + lj;
            if (g1[0] != 0.f || g1[1] != 0.f || g1[2] != 0.f) {
                add_force_fixed(&A.force[i], g1[0]);
                add_force_fixed(&A.force[(size_t) A.c.np+i], g1[1]);
                add_force_fixed(&A.force[2*(size_t) A.c.np+i], g1[2]);
            }
            if (g1[3] != 0.f) atomicAdd(&A.yq[i], g1[3]);
        }
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
// k_bw: bw_i = brw_i + bru_i, bru_i = -(k/4pi)(q_i^2 + Y_i B_i) fp_i   (ReferenceAGBNPKernels.cpp:537-542)
// ---------------------------------------------------------------------------------------------------------------
struct BwArgs {
    int np;
    const float4* posq;
    const float *yq, *born, *bfp, *brw;
    float kdiel;
    float* bw;
};

__global__ void __launch_bounds__(256) k_bw(BwArgs A) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= A.np) return;
    const float q = A.posq[i].w;
    const float y = A.yq[i]/(-2.f*A.kdiel);
    A.bw[i] = A.brw[i] - PIFAC*A.kdiel*(q*q + y*A.born[i])*A.bfp[i];
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
