// Pair-pass kernels of the AGBNP1 path for sm_100a: atom gather + block bounding boxes, inverse Born radii (S4-S5),
// GB pair energy/force + Y (S6), per-atom vdW / self terms (S7-S8), Born-radius derivative pass (S9), force scatter.
// Reference semantics: platforms/reference/src/ReferenceAGBNPKernels.cpp:421-586 (double loops over all pairs).
// Decomposition here: atoms in blocks of 32 (sorted order, heavy first); passes bounded by the 2.0 nm table range or
// by the cutoff cull block pairs on the fly with bounding boxes (no stored neighbor list); the Born and derivative
// passes test each 32x32 tile once, keep the in-range pairs as per-lane bit masks and walk only those (see "Range-limited
// pair passes" below); the GB pass, which has no range limit without a cutoff, uses symmetric 32x32 register tiles
// (8 i-atoms x 4 j-atoms per lane) in packed FP32x2 arithmetic, warp-shuffle reduce-scatter of the partial sums and one
// vector red.global per atom and tile.
#ifndef AGBNP_PAIR_CUH_
#define AGBNP_PAIR_CUH_

#include "agbnp_device.cuh"

namespace agbnp_b200_impl {

constexpr int GB_THREADS = 128;
#ifndef GB_MIN_BLOCKS
#define GB_MIN_BLOCKS 3
#endif
#ifndef GB_FAR_FACTOR
#define GB_FAR_FACTOR 64.f          // d^2 > 64 B_i B_j: exp(-d^2/4B_iB_j) < e^-16
#endif
#ifndef GB_CHUNK_TILES
#define GB_CHUNK_TILES 8
#endif
constexpr int GB_CHUNK = GB_CHUNK_TILES;    // column tiles per GB work unit (<= 32: one bit per tile)
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
    float4* slab;               // the per-evaluation accumulators and control words, zeroed here (slab_vec float4)
    int slab_vec;
    const float4* posq_ref;     // sorted positions at the time the pair masks were built (PairUnits::ctl)
    int* pq_ctl;                // [1] max over atoms of |x - x_ref|^2 (float bits; reset by k_finish)
    // layout of the caller's device buffers (agbnp_b200_set_device_layout): particle -> position in the buffers (the CUDA
    // platform reorders atoms), or null = identity; positions as double4 instead of float4 (its double-precision mode)
    const int* io;
    int posq_f64;
};

__global__ void __launch_bounds__(256) k_prep(PrepArgs A) {
    pdl_release();
    pdl_acquire();                  // the previous evaluation's k_finish still reads the slab this kernel zeroes
    const int lane = threadIdx.x & 31;
    const int gid = blockIdx.x*blockDim.x + threadIdx.x;
    for (int i = gid; i < A.slab_vec; i += gridDim.x*blockDim.x) A.slab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int blk = gid >> 5;
    if (blk*TILE >= A.np) return;
    const int k = blk*TILE+lane;
    const int o = A.orig[k];
    float4 p;
    float lo[3], hi[3];
    if (o >= 0) {
        const int src = A.io ? A.io[o] : o;
        if (A.posq_f64) { const double4 q = ((const double4*) A.posq_in)[src]; p = make_float4((float) q.x, (float) q.y, (float) q.z, 0.f); }
        else p = A.posq_in[src];
        p.w = A.charge[k];
        lo[0] = hi[0] = p.x; lo[1] = hi[1] = p.y; lo[2] = hi[2] = p.z;
    } else {
        // padding: far away, pairwise distinct, zero charge -- contributes exactly nothing anywhere
        p = make_float4(1.0e4f + 50.f*lane, 1.0e4f + 50.f*(blk % 1000), 1.0e4f + 50.f*(blk/1000), 0.f);
        lo[0] = lo[1] = lo[2] = 3.0e38f; hi[0] = hi[1] = hi[2] = -3.0e38f;
    }
    A.posq[k] = p;
    {
        // how far has any atom moved since the range-limited pair masks were built?  (k_born rebuilds them beyond skin/2)
        float d2 = 0.f;
        if (o >= 0) { const float4 q = A.posq_ref[k]; const float dx = p.x-q.x, dy = p.y-q.y, dz = p.z-q.z; d2 = dx*dx + dy*dy + dz*dz; }
        d2 = warp_max(d2);
        if (lane == 0 && d2 > 0.f) atomicMax(A.pq_ctl+1, __float_as_int(d2));     // non-negative floats order like their bit patterns
        if (gid == 0) { A.pq_ctl[2] = 0; A.pq_ctl[4] = 0; }        // "rebuilt in this evaluation" flags (k_born, k_tree)
    }
#pragma unroll
    for (int c = 0; c < 3; c++) { lo[c] = warp_min(lo[c]); hi[c] = warp_max(hi[c]); }
    if (lane == 0) {
        A.bbc[blk] = make_float4(0.5f*(lo[0]+hi[0]), 0.5f*(lo[1]+hi[1]), 0.5f*(lo[2]+hi[2]), 0.f);
        A.bbh[blk] = make_float4(0.5f*(hi[0]-lo[0]), 0.5f*(hi[1]-lo[1]), 0.5f*(hi[2]-lo[2]), 0.f);
    }
}

// I4 splines in power form, one float4 per (table, interval): with fr = fraction of the interval in [0,1),
//   Q(d)  = v.x + fr (v.y + fr (v.z + fr v.w)),   Q'(d) = (v.y + 2 fr v.z + 3 fr^2 v.w)/h   (spline_slope, k_deriv)
// -- the natural cubic spline of AGBNPI4LookupTable and its derivative (AGBNPUtils.cpp:102-130, AGBNPUtils.h:104-115)
// expanded around the left knot on the host in double (agbnp_b200.cu: upload_static).
__device__ __forceinline__ float spline_value(float4 v, float fr) { return fmaf(fr, fmaf(fr, fmaf(fr, v.w, v.z), v.y), v.x); }

// Interval of the uniform-knot spline tables without the two conversion instructions (F2I, I2F: quarter-rate XU pipe, which
// the pair passes share with MUFU.RSQ): adding 1.5 * 2^23 rounds t - 1/2 to the nearest integer k = floor(t) into the low
// mantissa bits (an exact tie, t an integer, may give k = t - 1 with fr = 1: the same point of the continuous spline).
// Returns k clamped to the last interval and sets fr = t - k (pairs beyond the table range are masked by their callers).
__device__ __forceinline__ int spline_interval(float t, float& fr) {
    const float tk = (t - 0.5f) + 12582912.f;
    fr = t - (tk - 12582912.f);
    return min(__float_as_int(tk) - 0x4B400000, I4_INTERVALS-1);
}
__device__ __forceinline__ float rsqrt_fast(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

struct PairCommon {
    int np, nhb, nb;            // padded atoms, heavy blocks, all blocks
    const float4* posq;
    const int* orig;
    const float4 *bbc, *bbh;
    const unsigned char* ts;    // screened radius type
    const signed char* tj;      // screener radius type, -1 hydrogens / padding
    const float4* i4v;          // value tables  [ntables*15] (the derivative comes from the same cubic)
    int ntj, ntables;
    int tab_smem;               // 1: tables are staged in shared memory; 0 (too many radius classes): read through L1
    float inv_h;
    float range2;               // (2.0 nm)^2 table range
    float cut2;                 // cutoff^2 (float product), used when CUTOFF
    int row_begin, row_end;     // row blocks whose per-atom terms this shard reports
};

// ---------------------------------------------------------------------------------------------------------------
// Range-limited pair passes (Born radii S5, Born-radius derivatives S9).
//
// The reference loops over ORDERED pairs (i any, j heavy, d < 2.0 nm; ReferenceAGBNPKernels.cpp:437-450,557-586).  Here a
// warp owns a work unit = (row block, chunk of PQ_CHUNK column blocks); for every column block whose bounding box is in
// range it
//   1. tests the 32x32 atom pairs ONCE: lane a scans the 32 column atoms (staged in shared memory) and keeps a 32-bit
//      row mask; the ballot of each test, kept by lane b, is column atom b's mask of row atoms (the transposed matrix);
//   2. primary role: lane a walks the set bits of its row mask.  The distance, the spline interval and BOTH table rows
//      of a pair -- (screened a, screener b) and (screened b, screener a) -- are evaluated once, here: what atom a
//      receives stays in registers, what column atom b receives from a goes into a per-warp 32x33 shared-memory matrix;
//   3. secondary role: lane b walks its column mask and only adds up its column of the matrix.
// Only pairs that are in range are visited (~18% of the tested ones for the 2.0 nm range and 32-atom blocks), and every
// sum lives in a register or in a matrix slot written by exactly one lane: no atomics inside a tile.  Row sums leave
// once per unit, column sums once per tile, as one red.global per atom.
// Both passes are bound by the shared-memory pipe (random-address table and partner loads: ncu r2p has k_deriv at 82 %
// of the l1tex wavefront peak), so the design minimises table lookups per pair: one 16-byte row gives the spline value
// AND its derivative (the derivative of the same cubic, AGBNPUtils.h:104-115), two rows serve both directions of a pair.
// Units: heavy rows x heavy columns cb >= ra (the diagonal tile is handled by the primary role alone: every lane visits
// all its partners) and hydrogen rows x heavy columns (hydrogens never descreen, ReferenceAGBNPKernels.cpp:442,563: one
// table row per pair).
// ---------------------------------------------------------------------------------------------------------------
constexpr int PQ_CHUNK = 8;             // most column blocks one unit may cover (far-apart block pairs are packed)
constexpr int PQ_MAX_THREADS = 1024;
constexpr int WMAT_STRIDE = 33;         // row stride of the per-warp pair matrices: lanes writing the same column hit different banks

// staged atom: table row offsets of both roles in one word -- bits 12..: ts*ntj*15 (row base as the SCREENED atom), bits 0..10:
// tj*15 (row offset as the SCREENER, 0 for hydrogens / padding), bit 11: not a screener
constexpr int PK_TJ = 0x7ff, PK_HYD = 0x800, PK_TS_SHIFT = 12;
__device__ __forceinline__ int pk_pack(int ts, int tj, int ntj) {
    return (ts*ntj*I4_INTERVALS << PK_TS_SHIFT) | (tj < 0 ? PK_HYD : tj*I4_INTERVALS);
}

struct PairUnits {
    const int2* units;          // (row block, first column block | number of column blocks << 20), heaviest first
    int nunits;
    // pair-test cache: k_born and k_deriv visit the same units and tiles, so k_born stores what it found and k_deriv
    // neither re-tests the bounding boxes nor the 32x32 atom pairs
    const int* tile_off;        // [nunits] first tile slot of the unit (static prefix of the units' column-block counts)
    unsigned* unit_hits;        // [nunits] bit t set: column block t of the unit is in range
    uint2* masks;               // [tile slots][32] (row mask, column mask) of every lane
    int* work_counter;
    int shard_rank, shard_count;    // units are dealt round-robin to shards (shard_count 1: all)
    // Verlet-style reuse of the pair-test cache ACROSS evaluations: the masks are built with the range enlarged by a skin
    // and stay valid while no atom has moved more than skin/2 since (k_prep measures it, k_born decides, k_finish moves
    // the reference positions); every walked pair is re-tested against the exact range, so the set of pairs that
    // contribute -- the membership -- is decided per evaluation exactly as before.
    int* ctl;                       // persistent: [0] masks valid for the current order, [1] max displacement^2 (float bits),
                                    // [2] this evaluation rebuilt the masks (written by k_born, read by k_finish)
    float list2;                    // (range + skin)^2: threshold of the box and atom tests when the masks are built
    float move2;                    // (0.49 skin)^2: rebuild when an atom has moved further than this since the last build
};

__device__ __forceinline__ float pq_dist2(bool exact, float dx, float dy, float dz) {
    return exact ? dist2_exact(dx, dy, dz) : fmaf(dz, dz, fmaf(dy, dy, dx*dx));
}

// step 1: masks of in-range pairs of one tile.  s_c: column atoms (x, y, z, -), (px,py,pz): this lane's row atom.
template <bool CUTOFF>
__device__ __forceinline__ void pq_masks(const float4* s_c, float px, float py, float pz,
                                         float lim2, bool diag, int lane, unsigned& rowmask, unsigned& colmask) {
    rowmask = 0; colmask = 0;
#pragma unroll 8
    for (int jj = 0; jj < TILE; jj++) {
        const float4 c = s_c[jj];                                   // one broadcast 16-byte load
        const float dx = c.x-px, dy = c.y-py, dz = c.z-pz;
        const float d2 = pq_dist2(CUTOFF, dx, dy, dz);
        const bool ok = d2 < lim2 && !(diag && jj == lane);
        const unsigned m = __ballot_sync(FULL, ok);
        if (lane == jj) colmask = m;
        rowmask |= (ok ? 1u : 0u) << jj;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_born: bsum_i = sum_{j heavy, j != i, d < 2.0} (s_j/4pi) Q(d; type_i, type_j)   (ReferenceAGBNPKernels.cpp:437-450)
// ---------------------------------------------------------------------------------------------------------------
struct BornArgs {
    PairCommon c;
    PairUnits u;
    const float4* accS;         // .w = self volumes (vdW radii), from k_tree
    const double* vS;           // atomic volumes (vdW radii)
    float* bsum;                // out [np], zeroed slab
    unsigned long long* counters;
};

struct BornSmem { float4 p[TILE]; int pk[TILE]; };      // p = (x, y, z, s_j/(4 pi)); .w = 0 for non-screeners; pk: see pk_pack
struct BornMe { float px, py, pz, s; int ts, tj; bool heavy; };        // ts = row base as the screened atom, tj = row offset as the screener
constexpr size_t BORN_WARP_SMEM = 2*sizeof(BornSmem) + WMAT_STRIDE*TILE*sizeof(float);

// partner jj of atom "me" (valid = the pair exists; invalid slots run the same instructions on a harmless atom): returns what
// me receives; BOTH: what the partner receives from me goes into vm[lane][jj] (off-diagonal tiles of heavy row blocks)
template <bool CUTOFF, bool BOTH>
__device__ __forceinline__ float born_term(const float4* tabv, const BornSmem& o, float* vm, int lane, int jj, bool valid, const BornMe& me,
                                           float inv_h, float lim2, unsigned& npair) {
    const int pk = o.pk[jj];
    const float4 c = o.p[jj];
    const float dx = c.x-me.px, dy = c.y-me.py, dz = c.z-me.pz;
    const float d2 = pq_dist2(CUTOFF, dx, dy, dz);
    const float d = d2*rsqrt_fast(fmaxf(d2, 1e-20f));
    float fr;
    const int k = spline_interval(d*inv_h, fr);
    const bool in = valid && d2 < lim2;                      // the masks carry a skin: the range itself is tested here
    const bool use = in && !(pk & PK_HYD);
    const float q = spline_value(tabv[me.ts + (pk & PK_TJ) + k], fr);
    npair += use;
    if (BOTH) {
        const bool rev = in && me.heavy;
        const float qr = spline_value(tabv[(pk >> PK_TS_SHIFT) + me.tj + k], fr);
        npair += rev;
        if (valid) vm[lane*WMAT_STRIDE + jj] = rev ? me.s*qr : 0.f;
    }
    return use ? c.w*q : 0.f;
}

// primary role: everything atom "me" receives from the partners in `mask`; two partners per trip for instruction-level parallelism
template <bool CUTOFF, bool BOTH>
__device__ __forceinline__ float born_role(const float4* tabv, const BornSmem& o, float* vm, int lane, unsigned mask, const BornMe& me,
                                           float inv_h, float lim2, unsigned& npair) {
    float s0 = 0.f, s1 = 0.f;
    while (mask) {
        const int j0 = __ffs(mask)-1;
        mask &= mask-1;
        const bool two = mask != 0;
        const int j1 = two ? __ffs(mask)-1 : j0;
        mask &= mask-1;
        s0 += born_term<CUTOFF, BOTH>(tabv, o, vm, lane, j0, true, me, inv_h, lim2, npair);
        s1 += born_term<CUTOFF, BOTH>(tabv, o, vm, lane, j1, two, me, inv_h, lim2, npair);
    }
    return s0+s1;
}

// DENSE tiles (at least PQ_DENSE of the 1024 pairs listed; most in-range pairs live in such tiles): the same terms without the
// mask walk -- every lane visits partner jj = 0..31 in step, so the partner's data is one broadcast load instead of a gather,
// the matrix is written and read without bank conflicts, and there is no bit-scan per term; a lane's unlisted pairs run as
// invalid slots (about a fifth of the slots of a dense tile, fewer than the mask walk leaves idle at its pace of the busiest lane)
#ifndef PQ_DENSE_MIN
#define PQ_DENSE_MIN 512
#endif
constexpr int PQ_DENSE = PQ_DENSE_MIN;
template <bool CUTOFF, bool BOTH>
__device__ __forceinline__ float born_role_dense(const float4* tabv, const BornSmem& o, float* vm, int lane, unsigned mask, const BornMe& me,
                                                 float inv_h, float lim2, unsigned& npair) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 2
    for (int jj = 0; jj < TILE; jj += 2) {
        const bool v0 = (mask >> jj) & 1u, v1 = (mask >> (jj+1)) & 1u;
        if (BOTH) { if (!v0) vm[lane*WMAT_STRIDE + jj] = 0.f; if (!v1) vm[lane*WMAT_STRIDE + jj+1] = 0.f; }
        s0 += born_term<CUTOFF, BOTH>(tabv, o, vm, lane, jj, v0, me, inv_h, lim2, npair);
        s1 += born_term<CUTOFF, BOTH>(tabv, o, vm, lane, jj+1, v1, me, inv_h, lim2, npair);
    }
    return s0+s1;
}
__device__ __forceinline__ float born_column_dense(const float* vm, int lane) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
    for (int a = 0; a < TILE; a += 2) { s0 += vm[a*WMAT_STRIDE + lane]; s1 += vm[(a+1)*WMAT_STRIDE + lane]; }
    return s0+s1;
}

// secondary role: column `lane` of the matrix over the rows in `mask`
__device__ __forceinline__ float born_column(const float* vm, int lane, unsigned mask) {
    float s0 = 0.f, s1 = 0.f;
    while (mask) {
        const int a0 = __ffs(mask)-1;
        mask &= mask-1;
        const bool two = mask != 0;
        const int a1 = two ? __ffs(mask)-1 : a0;
        mask &= mask-1;
        s0 += vm[a0*WMAT_STRIDE + lane];
        const float v1 = vm[a1*WMAT_STRIDE + lane];
        s1 += two ? v1 : 0.f;
    }
    return s0+s1;
}

__device__ __forceinline__ void born_load(const BornArgs& A, int blk, int lane, BornSmem& s) {
    const int j = blk*TILE+lane;
    const float4 p = A.c.posq[j];
    const double vj = A.vS[j];
    s.p[lane] = make_float4(p.x, p.y, p.z, vj > 0 ? PIFAC*(A.accS[j].w/(float) vj) : 0.f);
    s.pk[lane] = pk_pack((int) A.c.ts[j], (int) A.c.tj[j], A.c.ntj);
}

// TAB_SMEM: the spline tables are staged in shared memory (the normal case; a compile-time fact so that the lookups are
// LDS instead of generic loads); otherwise (too many radius classes) they are read from global memory through L1.
// Launch shape (warps per CTA, CTAs per SM): pq_shape() on the host; the grid is fully resident (first_unit).
template <bool CUTOFF, bool TAB_SMEM>
__global__ void __launch_bounds__(PQ_MAX_THREADS) k_born(BornArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ntab = TAB_SMEM ? A.c.ntables*I4_INTERVALS : 0;
    float4* s_tabv = (float4*) smem_raw;                                // [ntables*15]
    const int nwarp = blockDim.x >> 5;
    BornSmem* sm = (BornSmem*) (s_tabv + ntab);                         // [nwarp][2]: row block, column block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* vm = (float*) (sm + 2*nwarp) + warp*WMAT_STRIDE*TILE;        // [nwarp][32*33]: what column atoms receive
    pdl_release();
    for (int i = threadIdx.x; i < ntab; i += blockDim.x) s_tabv[i] = A.c.i4v[i];      // per-context constants: before the wait
    pdl_acquire();
    __syncthreads();
    const float4* tabv;
    if (TAB_SMEM) tabv = s_tabv; else tabv = A.c.i4v;
    BornSmem& R = sm[2*warp];
    BornSmem& Cc = sm[2*warp+1];
    const float lim2 = CUTOFF ? fminf(A.c.range2, A.c.cut2) : A.c.range2;
    // rebuild the pair masks (all units, this evaluation) or walk the stored ones?  Nothing below writes ctl[0] / ctl[1].
    const bool rebuild = A.u.ctl[0] == 0 || __int_as_float(A.u.ctl[1]) > A.u.move2;
    if (rebuild && blockIdx.x == 0 && threadIdx.x == 0) A.u.ctl[2] = 1;
    tail_begin(0);
    unsigned npair = 0;
    // units are dealt round-robin to the shards: claim c is this shard's c-th unit (a shard never touches, or pays a claim
    // for, the units of another)
    for (int c = first_unit(); ; c = next_unit(A.u.work_counter, lane)) {
        const int u = c*A.u.shard_count + A.u.shard_rank;
        if (u >= A.u.nunits) break;
        const int2 un = A.u.units[u];
        const int ra = un.x;
        const int cb0 = un.y & 0xfffff, nc = un.y >> 20;
        unsigned hits;
        if (rebuild) {
            bool hit = false;
            if (lane < nc) hit = box_box_dist2(A.c.bbc[ra], A.c.bbh[ra], A.c.bbc[cb0+lane], A.c.bbh[cb0+lane]) < A.u.list2;
            hits = __ballot_sync(FULL, hit);
            if (lane == 0) A.u.unit_hits[u] = hits;
        } else hits = A.u.unit_hits[u];
        if (!hits) continue;
        const int toff = A.u.tile_off[u];
        __syncwarp();
        born_load(A, ra, lane, R);
        __syncwarp();
        const bool row_heavy = ra < A.c.nhb;                 // hydrogen rows receive but never descreen
        BornMe me;
        { const float4 pa = R.p[lane]; const int pk = R.pk[lane];
          me.px = pa.x; me.py = pa.y; me.pz = pa.z; me.s = pa.w; me.ts = pk >> PK_TS_SHIFT; me.tj = pk & PK_TJ; me.heavy = !(pk & PK_HYD); }
        float rsum = 0.f;
        while (hits) {
            const int cb = cb0 + __ffs(hits)-1;
            hits &= hits-1;
            const bool diag = cb == ra;
            __syncwarp();                                    // the previous tile's column sums have been read
            born_load(A, cb, lane, Cc);
            __syncwarp();
            unsigned rowmask, colmask;
            if (rebuild) {
                pq_masks<CUTOFF>(Cc.p, me.px, me.py, me.pz, A.u.list2, diag, lane, rowmask, colmask);
                A.u.masks[(size_t) (toff + cb-cb0)*TILE + lane] = make_uint2(rowmask, colmask);
            } else {
                const uint2 mk = A.u.masks[(size_t) (toff + cb-cb0)*TILE + lane];
                rowmask = mk.x; colmask = mk.y;
            }
            const bool dense = __reduce_add_sync(FULL, __popc(rowmask)) >= PQ_DENSE;
            if (diag || !row_heavy) {
                if (dense) rsum += born_role_dense<CUTOFF, false>(tabv, Cc, vm, lane, rowmask, me, A.c.inv_h, lim2, npair);
                else rsum += born_role<CUTOFF, false>(tabv, Cc, vm, lane, rowmask, me, A.c.inv_h, lim2, npair);
            } else {
                float csum;
                if (dense) {
                    rsum += born_role_dense<CUTOFF, true>(tabv, Cc, vm, lane, rowmask, me, A.c.inv_h, lim2, npair);
                    __syncwarp();                            // the matrix is read by other lanes
                    csum = born_column_dense(vm, lane);
                } else {
                    rsum += born_role<CUTOFF, true>(tabv, Cc, vm, lane, rowmask, me, A.c.inv_h, lim2, npair);
                    __syncwarp();
                    csum = born_column(vm, lane, colmask);
                }
                if (csum != 0.f) atomicAdd(&A.bsum[cb*TILE+lane], csum);
            }
        }
        if (rsum != 0.f) atomicAdd(&A.bsum[ra*TILE+lane], rsum);
    }
    tail_end(0);
    unsigned long long np64 = (unsigned long long) warp_sum((double) npair);
    if (lane == 0 && np64) atomicAdd(&A.counters[CT_PQ], np64);
}

// ---------------------------------------------------------------------------------------------------------------
// k_born_finish: beta_i = 1/r_i - bsum_i; B_i = 1/swf(beta_i) (agbnp_swf_invbr, ReferenceAGBNPKernels.cpp:41-55,452-454),
// volume scaling factors (:421-430), GB self energy (:477), vdW energy and brw (:513-528), GB atom records for k_gb
// ---------------------------------------------------------------------------------------------------------------
struct BornFinishArgs {
    int np;
    const float4* posq;
    const int* orig;
    const float* bsum;
    const float4* accS;
    const double* vS;
    const float *radius, *alpha;
    float *vsf, *born, *bfp, *brw;
    float4* gbj;                // out [3*np]: GB atom records in broadcast form (see k_gb)
    float* bmax;                // out [nb]: largest Born radius of each 32-atom block (0 for a block of padding)
    float kdiel, hb_radius;
    double* scalars;
    int own_begin, own_end;     // sorted-index range whose per-atom energies this shard reports
};

__global__ void __launch_bounds__(256) k_born_finish(BornFinishArgs A) {
    pdl_release();
    pdl_acquire();
    const int a = blockIdx.x*blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    float evdw = 0.f, eself = 0.f, br_real = 0.f;
    if (a < A.np) {
        const float4 pa = A.posq[a];
        float br = 1.f;
        if (A.orig[a] >= 0) {
            const float beta = 1.f/A.radius[a] - A.bsum[a];
            const float ia = 0.5f, ia2 = 0.25f;          // 1/2.0, 1/2.0^2
            float t, fp;
            if (beta < 0.f) { t = ia; fp = 0.f; }
            else { t = sqrtf(ia2 + beta*beta); fp = beta/t; }
            br = 1.f/t;
            br_real = br;
            A.born[a] = br;
            A.bfp[a] = fp;
            const double va = A.vS[a];
            A.vsf[a] = va > 0 ? A.accS[a].w/(float) va : 0.f;
            const float q = pa.w, al = A.alpha[a];
            const float bh = br + A.hb_radius;
            const float bh3 = bh*bh*bh;
            if (a >= A.own_begin && a < A.own_end) { eself = A.kdiel*q*q/br; evdw = al/bh3; }
            A.brw[a] = -PIFAC*3.f*al*br*br*fp/(bh3*bh);
        } else {
            A.born[a] = 1.f; A.bfp[a] = 0.f; A.vsf[a] = 0.f; A.brw[a] = 0.f;
        }
        const float qs = pa.w, ib = 0.60056120439322491f/br;   // sqrt(log2(e)/4)/B
        float4* rec = A.gbj + 3*(size_t) a;
        rec[0] = make_float4(pa.x, pa.x, pa.y, pa.y);
        rec[1] = make_float4(pa.z, pa.z, qs, qs);
        rec[2] = make_float4(br, br, ib, ib);
    }
    const float bm = warp_max(br_real);
    if (lane == 0 && a < A.np) A.bmax[a >> 5] = bm;
    const double es = warp_sum((double) eself), ev = warp_sum((double) evdw);
    if (lane == 0 && (es != 0.0 || ev != 0.0)) { atomicAdd(&A.scalars[SC_EGB], es); atomicAdd(&A.scalars[SC_EVDW], ev); }
}

// ---------------------------------------------------------------------------------------------------------------
// k_gb: GB pair energy, direct force and Y accumulators over symmetric 32x32 tiles
// (ReferenceAGBNPKernels.cpp:476-498), in packed FP32x2 arithmetic (fma.rn.f32x2 -> FFMA2 on sm_100a).
//
// The pass is bound by the FP32 pipe: one pair costs 27 FP32 operations + 2 MUFU (ex2, rsqrt).  Scalar code spends an
// issue slot per operation; here every arithmetic instruction handles the pairs (i0,j) and (i1,j) of two row atoms at
// once, which halves the issue slots of the FP32 part and leaves the FMA pipe itself as the limit.
//   lane grid 4 (li) x 8 (lj): a lane owns 8 row atoms (4 packed pairs) for a whole work unit and 4 column atoms per
//   tile.  Column tiles are staged in shared memory, double buffered per warp with cp.async (one 48-byte record per
//   lane and tile, written by k_born_finish in broadcast form {x,x,y,y | z,z,q,q | B,B,ib,ib} so that 16-byte shared
//   loads land directly in aligned register pairs -- no MOVs to build the packed operands), so the next tile's
//   global-memory latency hides behind the current tile's 16 packed pair evaluations per lane;
//   column-side sums are reduce-scattered over the 4 lanes that share them (12 shuffles per tile) and leave as ONE
//   vector atomic (red.global.add.v4.f32) per lane and tile; row-side sums are flushed once per work unit.
// Charges enter UNSCALED: force fields use a few dozen distinct charge values, so the float rounding of a scaled charge
// q*sqrt(-2k) and of the products is the same for every pair of the same two atom types and does not average out
// (measured: +6.5e-7 .. +1.1e-6 relative bias of the pair energy on 2clr / RNase H; 5-8 times less unscaled).  The GB
// prefactor -2k is applied once per atom / per sum downstream (k_gb's energy reduction, k_bw, k_finish).
// ib = sqrt(log2(e)/4)/B, so that exp(-d2/(4 B_i B_j)) = ex2(-d2 ib_i ib_j).
// ---------------------------------------------------------------------------------------------------------------
struct GBArgs {
    PairCommon c;
    const float4* gbj;          // [3*np] GB atom records in broadcast form {x,x,y,y | z,z,q,q | B,B,ib,ib}
    const int2* units;          // (row block, first column block | column blocks << 20): triangular cover in chunks of up to `chunk` column tiles
    int nunits;
    int chunk;                  // column tiles per unit (<= GB_CHUNK), chosen on the host so that every warp gets several units
    int shard_rank, shard_count;
    float4* gbacc;              // out [np]: (fx, fy, fz)/(-2k) (GB pair force), Y_i
    double kdiel;               // k = 4.184*332/10*(-1/2)(1 - 1/80)  (ReferenceAGBNPKernels.cpp:465-468)
    const float* bmax;          // [nb] largest Born radius in each block (k_born_finish)
    double* scalars;
    unsigned long long* counters;
    int* work_counter;
};

__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((unsigned) __cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N)); }

// staged column tile: atom c of the tile lives at slot (c & 3)*8 + (c >> 2), so that the 8 lanes of a quarter warp (which
// read atoms lj*4 + n for lj = 0..7) hit 8 consecutive 16-byte slots: no bank conflicts
struct GBStage { float4 r0[TILE], r1[TILE], r2[TILE]; };

__device__ __forceinline__ void gb_prefetch(const GBArgs& A, int cb, int lane, GBStage& st) {
    const float4* rec = A.gbj + 3*(size_t) (cb*TILE + lane);
    const int slot = (lane & 3)*8 + (lane >> 2);
    cp_async16(&st.r0[slot], rec);
    cp_async16(&st.r1[slot], rec+1);
    cp_async16(&st.r2[slot], rec+2);
    cp_async_commit();
}

// one 32x32 tile: 4 column atoms x 4 packed row pairs per lane.
// FAR: every pair of the tile has d^2 > 64 B_i B_j, i.e. exp(-d^2/4B_iB_j) < e^-16 = 1.1e-7 and B_iB_j exp(..)/d^2 < 2e-9:
// below float resolution in f = 1/sqrt(d^2 + B_iB_j exp(..)), so the pair is plain Coulomb (f = 1/d, no Y term) and
// costs 17 instead of 27 FP32 operations and 1 instead of 2 MUFU.
template <bool CUTOFF, bool DIAG, bool FAR>
__device__ __forceinline__ void gb_tile(const GBArgs& A, const GBStage& st, int cb, int li, int lj,
                                        const float2 (&nx)[4], const float2 (&ny)[4], const float2 (&nz)[4],
                                        const float2 (&qi)[4], const float2 (&bi)[4], const float2 (&nib)[4],
                                        float2 (&fi)[4][4], float2& e2, unsigned& npair) {
    const float2 m025 = make_float2(-0.25f, -0.25f), p025 = make_float2(0.25f, 0.25f);
    const float2 cut2 = make_float2(A.c.cut2, A.c.cut2);
    float sj[4][4];
#pragma unroll
    for (int n = 0; n < 4; n++) {
        const int jl = lj*4 + n;                                  // column atom within the tile
        // broadcast-form records {x,x,y,y | z,z,q,q | B,B,ib,ib}: three 16-byte loads land directly in aligned register pairs
        const float4 r0 = st.r0[n*8 + lj], r1 = st.r1[n*8 + lj], r2 = st.r2[n*8 + lj];
        const float2 xj = make_float2(r0.x, r0.y), yj = make_float2(r0.z, r0.w), zj = make_float2(r1.x, r1.y);
        const float2 qj = make_float2(r1.z, r1.w), bj = make_float2(r2.x, r2.y), ibj = make_float2(r2.z, r2.w);
        float2 ax = make_float2(0.f, 0.f), ay = ax, az = ax, aY = ax;
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const float2 dx = __fadd2_rn(xj, nx[m]), dy = __fadd2_rn(yj, ny[m]), dz = __fadd2_rn(zj, nz[m]);
            float2 d2;
            if (CUTOFF) d2 = __fadd2_rn(__fadd2_rn(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy)), __fmul2_rn(dz, dz));   // membership rule: no contraction
            else d2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
            float2 bb, et, t;
            if (FAR) t = d2;
            else {
                bb = __fmul2_rn(bi[m], bj);
                const float2 arg = __fmul2_rn(d2, __fmul2_rn(nib[m], ibj));
                et = make_float2(fast_ex2(arg.x), fast_ex2(arg.y));
                t = __ffma2_rn(bb, et, d2);
            }
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
            // energy straight from MUFU.RSQ: on B200 rsqrt.approx.ftz has a mean relative error of -5e-9 over all mantissas
            // (rms 3.4e-8, max 1.25e-7; tools/rsqrt_bias.cu), which the pair sum does not see -- a Newton step on f (two
            // more packed operations per pair) bought nothing measurable in energy parity and cost 9 us (profiles/r2_experiments.md)
            e2 = __fadd2_rn(e2, qf);
            const float2 g = __fmul2_rn(qf, ff);
            float2 mw = g;
            if (!FAR) {
                const float2 hh = __fmul2_rn(g, et);
                mw = __ffma2_rn(m025, hh, g);                     // q_i q_j (1 - e/4) f^3
                const float2 yt = __fmul2_rn(hh, __ffma2_rn(p025, d2, bb));
                fi[m][3] = __fadd2_rn(fi[m][3], yt);
                aY = __fadd2_rn(aY, yt);
            }
            fi[m][0] = __ffma2_rn(dx, mw, fi[m][0]); fi[m][1] = __ffma2_rn(dy, mw, fi[m][1]); fi[m][2] = __ffma2_rn(dz, mw, fi[m][2]);
            ax = __ffma2_rn(dx, mw, ax); ay = __ffma2_rn(dy, mw, ay); az = __ffma2_rn(dz, mw, az);
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
__global__ void __launch_bounds__(GB_THREADS, GB_MIN_BLOCKS) k_gb(GBArgs A) {
    __shared__ GBStage s_stage[GB_THREADS/32][2];
    pdl_release();
    pdl_acquire();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int li = lane >> 3, lj = lane & 7;
    GBStage* stage = s_stage[warp];
    double e_acc = 0.0;
    unsigned long long npair = 0, ntile = 0;
    tail_begin(1);
    for (int c = claim_unit(A.work_counter, lane); ; c = claim_unit(A.work_counter, lane)) {
        const int u = c*A.shard_count + A.shard_rank;             // this shard's c-th unit (round-robin deal)
        if (u >= A.nunits) break;
        const int2 un = A.units[u];
        const int ra = un.x;
        const int ucol = un.y & 0xfffff, cend = ucol + (un.y >> 20);
        const float4 ca = A.c.bbc[ra], ha = A.c.bbh[ra];
        // column tiles of this unit that take part (cutoff: bounding boxes within range), as a bit mask
        unsigned tiles = 0, fartiles = 0;
        {
            bool hit = false, far = false;
            if (lane < cend-ucol) {
                const float bd2 = box_box_dist2(ca, ha, A.c.bbc[ucol+lane], A.c.bbh[ucol+lane]);
                hit = !CUTOFF || bd2 < A.c.cut2;
                far = bd2 > GB_FAR_FACTOR*A.bmax[ra]*A.bmax[ucol+lane];
            }
            tiles = __ballot_sync(FULL, hit);
            fartiles = __ballot_sync(FULL, far);
        }
        if (!tiles) continue;
        const int col0 = ucol;
        __syncwarp();                                             // the previous unit is done with the stages
        int cur = 0;
        gb_prefetch(A, col0 + __ffs(tiles)-1, lane, stage[0]);
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
        float2 e2 = make_float2(0.f, 0.f);
        unsigned np32 = 0;
        while (tiles) {
            const int tb = __ffs(tiles)-1;
            const int cb = col0 + tb;
            tiles &= tiles-1;
            if (tiles) { gb_prefetch(A, col0 + __ffs(tiles)-1, lane, stage[cur^1]); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncwarp();
            ntile++;
            if (cb == ra) {
                gb_tile<CUTOFF, true, false>(A, stage[cur], cb, li, lj, nx, ny, nz, qi, bi, nib, fi, e2, np32);
                if (!CUTOFF && lane < 16) np32 += 31;             // 496 = 16*31 pairs in a diagonal tile
            } else {
                if ((fartiles >> tb) & 1) gb_tile<CUTOFF, false, true>(A, stage[cur], cb, li, lj, nx, ny, nz, qi, bi, nib, fi, e2, np32);
                else gb_tile<CUTOFF, false, false>(A, stage[cur], cb, li, lj, nx, ny, nz, qi, bi, nib, fi, e2, np32);
                if (!CUTOFF) np32 += 32;
            }
            __syncwarp();                                         // everyone is done reading stage[cur] before it is refilled
            cur ^= 1;
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
    tail_end(1);
    e_acc = warp_sum(e_acc);
    npair = (unsigned long long) warp_sum((double) npair);
    if (lane == 0) {
        atomicAdd(&A.scalars[SC_EGB], 2.0*A.kdiel*e_acc);       // E_pair = 2k sum_{i<j} q_i q_j f   (ReferenceAGBNPKernels.cpp:484)
        atomicAdd(&A.counters[CT_PGB], npair);
        atomicAdd(&A.counters[CT_TILES_GB], ntile);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_deriv: Born-radius derivative pass (ReferenceAGBNPKernels.cpp:555-586), regrouped by the atom that RECEIVES each
// contribution.  For atom "me" and partner "o" (d < 2.0, o != me), with D = r_o - r_me:
//   F_me  += D/d [ heavy(o) bw_me s_o Q'(d; ts_me, tj_o)  +  heavy(me) bw_o s_me Q'(d; ts_o, tj_me) ]  =  D w(me,o)
//   WU_me += heavy(me) bw_o Q(d; ts_o, tj_me)                   (W and U merged: both are linear in brw / bru)
// Same unit / mask / role structure as k_born.  The force weight w is symmetric in (me, o) and the two table rows a pair
// needs -- (ts_me, tj_o) and (ts_o, tj_me) -- give Q and Q' of both directions, so the primary role (row atom) evaluates
// the pair completely and leaves (w, what the column atom's WU receives) in a 32x33 float2 matrix; the secondary role
// (column atom) adds up its column: F_o -= D w, WU_o += heavy(o) bw_me Q(d; ts_me, tj_o).
// ---------------------------------------------------------------------------------------------------------------
struct DerivArgs {
    PairCommon c;
    PairUnits u;
    const float* vsf;
    // bw_i = brw_i + bru_i, bru_i = -(k/4pi)(q_i^2 + Y_i B_i) fp_i   (ReferenceAGBNPKernels.cpp:537-542), formed when an
    // atom is staged: Y_i is complete once k_gb has finished
    const float4* gbacc;        // .w = Y_i
    const float *born, *bfp, *brw;
    float kdiel;
    float4* dacc;               // out [np]: fx, fy, fz, W+U (zeroed slab; float red.global)
};

struct DerivSmem { float4 p[TILE]; float bw[TILE]; int pk[TILE]; };     // p = (x, y, z, s); pk: see pk_pack

struct DerivMe { float px, py, pz, s, bw; int ts, tj; bool heavy; };    // ts = row base as the screened atom, tj = row offset as the screener

constexpr size_t DERIV_WARP_SMEM = 2*sizeof(DerivSmem) + WMAT_STRIDE*TILE*sizeof(float2);

// dQ/dfr of the power-form cubic (x + fr (y + fr (z + fr w)))' = y + 2 fr (z + 1.5 fr w); fr2 = 2 fr, fr15 = 1.5 fr
__device__ __forceinline__ float spline_slope(float4 v, float fr2, float fr15) { return fmaf(fr2, fmaf(fr15, v.w, v.z), v.y); }

// MODE 0: diagonal tile (both directions of a pair are visited by their own lanes: nothing is stored);
// MODE 1: off-diagonal tile of a heavy row block (both table rows, the column atom's share is stored);
// MODE 2: hydrogen row block: the row atoms never descreen, one table row, the column atom's share is stored;
// MODE 3: the roles of a hydrogen-row tile swapped ("me" is the heavy column atom, the partners are hydrogens): the other
//         table row only; what is stored for the partner is the force weight alone.
template <bool CUTOFF, int MODE>
__device__ __forceinline__ void deriv_term(const float4* tabv, const DerivSmem& o, float2* wm, int lane, int jj, bool valid,
                                           const DerivMe& me, float inv_h, float lim2, float& fx, float& fy, float& fz, float& wu) {
    const float4 c = o.p[jj];
    const float dx = c.x-me.px, dy = c.y-me.py, dz = c.z-me.pz;
    const float d2 = pq_dist2(CUTOFF, dx, dy, dz);
    const bool in = valid && d2 < lim2;                      // the masks carry a skin: the range itself is tested here
    const int pk = o.pk[jj];
    const float inv_d = rsqrt_fast(fmaxf(d2, 1e-20f));
    const float d = d2*inv_d;
    float fr;
    const int k = spline_interval(d*inv_h, fr);
    const float fr2 = fr+fr, fr15 = 1.5f*fr;
    float w = 0.f, vcol = 0.f;
    if (MODE != 3) {
        // row (ts_me, tj_o): o descreens me (needs heavy(o))
        const float4 tb = tabv[me.ts + (pk & PK_TJ) + k];
        const bool okb = in && !(pk & PK_HYD);
        w = okb ? me.bw*c.w*spline_slope(tb, fr2, fr15) : 0.f;
        if (MODE != 0) vcol = okb ? me.bw*spline_value(tb, fr) : 0.f;
    }
    if (MODE != 2) {
        // row (ts_o, tj_me): me descreens o (needs heavy(me))
        const float4 ta = tabv[(pk >> PK_TS_SHIFT) + me.tj + k];
        const float bwo = (in && me.heavy) ? o.bw[jj] : 0.f;
        w = fmaf(bwo*me.s, spline_slope(ta, fr2, fr15), w);
        wu = fmaf(bwo, spline_value(ta, fr), wu);
    }
    w *= inv_d*inv_h;
    if (MODE != 0 && valid) wm[lane*WMAT_STRIDE + jj] = make_float2(w, vcol);
    fx = fmaf(dx, w, fx); fy = fmaf(dy, w, fy); fz = fmaf(dz, w, fz);
}

template <bool CUTOFF, int MODE>
__device__ __forceinline__ float4 deriv_role(const float4* tabv, const DerivSmem& o, float2* wm, int lane, unsigned mask,
                                             const DerivMe& me, float inv_h, float lim2) {
    float fx0 = 0.f, fy0 = 0.f, fz0 = 0.f, wu0 = 0.f, fx1 = 0.f, fy1 = 0.f, fz1 = 0.f, wu1 = 0.f;
    while (mask) {
        const int j0 = __ffs(mask)-1;
        mask &= mask-1;
        const bool two = mask != 0;
        const int j1 = two ? __ffs(mask)-1 : j0;
        mask &= mask-1;
        deriv_term<CUTOFF, MODE>(tabv, o, wm, lane, j0, true, me, inv_h, lim2, fx0, fy0, fz0, wu0);
        deriv_term<CUTOFF, MODE>(tabv, o, wm, lane, j1, two, me, inv_h, lim2, fx1, fy1, fz1, wu1);
    }
    return make_float4(fx0+fx1, fy0+fy1, fz0+fz1, wu0+wu1);
}

// dense tiles (see born_role_dense): partner jj = 0..31 in step, unlisted pairs as invalid slots that store zeros
template <bool CUTOFF, int MODE>
__device__ __forceinline__ float4 deriv_role_dense(const float4* tabv, const DerivSmem& o, float2* wm, int lane, unsigned mask,
                                                   const DerivMe& me, float inv_h, float lim2) {
    float fx0 = 0.f, fy0 = 0.f, fz0 = 0.f, wu0 = 0.f, fx1 = 0.f, fy1 = 0.f, fz1 = 0.f, wu1 = 0.f;
#pragma unroll 2
    for (int jj = 0; jj < TILE; jj += 2) {
        const bool v0 = (mask >> jj) & 1u, v1 = (mask >> (jj+1)) & 1u;
        if (MODE != 0) {
            if (!v0) wm[lane*WMAT_STRIDE + jj] = make_float2(0.f, 0.f);
            if (!v1) wm[lane*WMAT_STRIDE + jj+1] = make_float2(0.f, 0.f);
        }
        deriv_term<CUTOFF, MODE>(tabv, o, wm, lane, jj, v0, me, inv_h, lim2, fx0, fy0, fz0, wu0);
        deriv_term<CUTOFF, MODE>(tabv, o, wm, lane, jj+1, v1, me, inv_h, lim2, fx1, fy1, fz1, wu1);
    }
    return make_float4(fx0+fx1, fy0+fy1, fz0+fz1, wu0+wu1);
}
__device__ __forceinline__ float4 deriv_column_dense(const DerivSmem& r, const float2* wm, int lane, float px, float py, float pz) {
    float fx0 = 0.f, fy0 = 0.f, fz0 = 0.f, wu0 = 0.f, fx1 = 0.f, fy1 = 0.f, fz1 = 0.f, wu1 = 0.f;
#pragma unroll 4
    for (int a = 0; a < TILE; a += 2) {
        const float4 r0 = r.p[a], r1 = r.p[a+1];
        const float2 m0 = wm[a*WMAT_STRIDE + lane], m1 = wm[(a+1)*WMAT_STRIDE + lane];
        fx0 = fmaf(r0.x-px, m0.x, fx0); fy0 = fmaf(r0.y-py, m0.x, fy0); fz0 = fmaf(r0.z-pz, m0.x, fz0); wu0 += m0.y;
        fx1 = fmaf(r1.x-px, m1.x, fx1); fy1 = fmaf(r1.y-py, m1.x, fy1); fz1 = fmaf(r1.z-pz, m1.x, fz1); wu1 += m1.y;
    }
    return make_float4(fx0+fx1, fy0+fy1, fz0+fz1, wu0+wu1);
}

// secondary role: column atom `lane` at (px, py, pz) adds up its column of the pair matrix over the rows in `mask`
__device__ __forceinline__ float4 deriv_column(const DerivSmem& r, const float2* wm, int lane, unsigned mask, float px, float py, float pz) {
    float fx0 = 0.f, fy0 = 0.f, fz0 = 0.f, wu0 = 0.f, fx1 = 0.f, fy1 = 0.f, fz1 = 0.f, wu1 = 0.f;
    while (mask) {
        const int a0 = __ffs(mask)-1;
        mask &= mask-1;
        const bool two = mask != 0;
        const int a1 = two ? __ffs(mask)-1 : a0;
        mask &= mask-1;
        const float4 r0 = r.p[a0], r1 = r.p[a1];
        const float2 m0 = wm[a0*WMAT_STRIDE + lane];
        float2 m1 = wm[a1*WMAT_STRIDE + lane];
        if (!two) m1 = make_float2(0.f, 0.f);
        fx0 = fmaf(r0.x-px, m0.x, fx0); fy0 = fmaf(r0.y-py, m0.x, fy0); fz0 = fmaf(r0.z-pz, m0.x, fz0); wu0 += m0.y;
        fx1 = fmaf(r1.x-px, m1.x, fx1); fy1 = fmaf(r1.y-py, m1.x, fy1); fz1 = fmaf(r1.z-pz, m1.x, fz1); wu1 += m1.y;
    }
    return make_float4(fx0+fx1, fy0+fy1, fz0+fz1, wu0+wu1);
}

__device__ __forceinline__ void deriv_load(const DerivArgs& A, int blk, int lane, DerivSmem& s) {
    const int j = blk*TILE+lane;
    const float4 p = A.c.posq[j];
    s.p[lane] = make_float4(p.x, p.y, p.z, A.vsf[j]);
    s.bw[lane] = A.brw[j] - PIFAC*A.kdiel*(p.w*p.w + A.gbacc[j].w*A.born[j])*A.bfp[j];
    s.pk[lane] = pk_pack((int) A.c.ts[j], (int) A.c.tj[j], A.c.ntj);
}

// Launch shape: pq_shape() on the host picks the warps per CTA that bring the most warps onto an SM (every warp needs
// DERIV_WARP_SMEM next to the CTA's copy of the spline table)
template <bool CUTOFF, bool TAB_SMEM>
__global__ void __launch_bounds__(PQ_MAX_THREADS) k_deriv(DerivArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ntab = TAB_SMEM ? A.c.ntables*I4_INTERVALS : 0;
    float4* s_tabv = (float4*) smem_raw;
    const int nwarp = blockDim.x >> 5;
    DerivSmem* sm = (DerivSmem*) (s_tabv + ntab);                       // [nwarp][2]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* wm = (float2*) (sm + 2*nwarp) + warp*WMAT_STRIDE*TILE;      // [nwarp][32*33] (force weight, column atom's WU share)
    pdl_release();
    for (int i = threadIdx.x; i < ntab; i += blockDim.x) s_tabv[i] = A.c.i4v[i];      // constants: before the wait
    pdl_acquire();
    __syncthreads();
    const float4* tabv;
    if (TAB_SMEM) tabv = s_tabv; else tabv = A.c.i4v;
    DerivSmem& R = sm[2*warp];
    DerivSmem& Cc = sm[2*warp+1];
    const float lim2 = CUTOFF ? fminf(A.c.range2, A.c.cut2) : A.c.range2;
    tail_begin(2);
    for (int c = first_unit(); ; c = next_unit(A.u.work_counter, lane)) {
        const int u = c*A.u.shard_count + A.u.shard_rank;         // this shard's c-th unit (round-robin deal, as in k_born)
        if (u >= A.u.nunits) break;
        const int2 un = A.u.units[u];
        const int ra = un.x;
        const int cb0 = un.y & 0xfffff;
        unsigned hits = A.u.unit_hits[u];                         // found by k_born
        if (!hits) continue;
        const int toff = A.u.tile_off[u];
        __syncwarp();
        deriv_load(A, ra, lane, R);
        __syncwarp();
        const int a = ra*TILE+lane;
        const bool row_heavy = ra < A.c.nhb;
        DerivMe me;
        { const float4 pa = R.p[lane]; const int pk = R.pk[lane];
          me.px = pa.x; me.py = pa.y; me.pz = pa.z; me.s = pa.w; me.bw = R.bw[lane];
          me.ts = pk >> PK_TS_SHIFT; me.tj = pk & PK_TJ; me.heavy = !(pk & PK_HYD); }
        float4 racc = make_float4(0.f, 0.f, 0.f, 0.f);
        while (hits) {
            const int cb = cb0 + __ffs(hits)-1;
            hits &= hits-1;
            const bool diag = cb == ra;
            __syncwarp();                                       // the previous tile's column sums have been read
            deriv_load(A, cb, lane, Cc);
            __syncwarp();
            const uint2 mk = A.u.masks[(size_t) (toff + cb-cb0)*TILE + lane];
            const unsigned rowmask = mk.x, colmask = mk.y;
            const bool dense = __reduce_add_sync(FULL, __popc(rowmask)) >= PQ_DENSE;
            // sparse tiles: a warp walks at the pace of its busiest lane, so the side whose busiest atom has fewer partners plays
            // the primary role (CPU count on 2clr: 22 % fewer primary trips on sparse tiles; measured k_deriv 76.6 -> 74.4 us)
            if (!dense && !diag && __reduce_max_sync(FULL, __popc(colmask)) < __reduce_max_sync(FULL, __popc(rowmask))) {
                DerivMe mc;
                { const float4 pc = Cc.p[lane]; const int pk = Cc.pk[lane];
                  mc.px = pc.x; mc.py = pc.y; mc.pz = pc.z; mc.s = pc.w; mc.bw = Cc.bw[lane];
                  mc.ts = pk >> PK_TS_SHIFT; mc.tj = pk & PK_TJ; mc.heavy = !(pk & PK_HYD); }
                const float4 cs = row_heavy ? deriv_role<CUTOFF, 1>(tabv, R, wm, lane, colmask, mc, A.c.inv_h, lim2)
                                            : deriv_role<CUTOFF, 3>(tabv, R, wm, lane, colmask, mc, A.c.inv_h, lim2);
                if (colmask) atomicAdd(&A.dacc[cb*TILE+lane], cs);
                __syncwarp();                                   // the matrix is read by other lanes
                const float4 rs = deriv_column(Cc, wm, lane, rowmask, me.px, me.py, me.pz);
                racc.x += rs.x; racc.y += rs.y; racc.z += rs.z; racc.w += rs.w;
                continue;
            }
            float4 r;
            if (dense) {
                if (diag) r = deriv_role_dense<CUTOFF, 0>(tabv, Cc, wm, lane, rowmask, me, A.c.inv_h, lim2);
                else if (row_heavy) r = deriv_role_dense<CUTOFF, 1>(tabv, Cc, wm, lane, rowmask, me, A.c.inv_h, lim2);
                else r = deriv_role_dense<CUTOFF, 2>(tabv, Cc, wm, lane, rowmask, me, A.c.inv_h, lim2);
            } else {
                if (diag) r = deriv_role<CUTOFF, 0>(tabv, Cc, wm, lane, rowmask, me, A.c.inv_h, lim2);
                else if (row_heavy) r = deriv_role<CUTOFF, 1>(tabv, Cc, wm, lane, rowmask, me, A.c.inv_h, lim2);
                else r = deriv_role<CUTOFF, 2>(tabv, Cc, wm, lane, rowmask, me, A.c.inv_h, lim2);
            }
            racc.x += r.x; racc.y += r.y; racc.z += r.z; racc.w += r.w;
            if (!diag) {
                __syncwarp();                                   // the matrix is read by other lanes
                const float4 pc = Cc.p[lane];
                const float4 cs = dense ? deriv_column_dense(R, wm, lane, pc.x, pc.y, pc.z) : deriv_column(R, wm, lane, colmask, pc.x, pc.y, pc.z);
                if (colmask) atomicAdd(&A.dacc[cb*TILE+lane], cs);
            }
        }
        atomicAdd(&A.dacc[a], racc);
    }
    tail_end(2);
}

// ---------------------------------------------------------------------------------------------------------------
// k_finish: scatter the fixed-point forces (sorted order) into the caller's sink and fold the energy terms
// ---------------------------------------------------------------------------------------------------------------
struct FinishArgs {
    int np, n;
    const int* orig;
    const float4 *accL, *accS;          // surface-tension gradients of sum coef*gamma*vol (enlarged / vdW radii), from k_tree
    const float4* gbacc;                // GB pair force (xyz) / gb_scale                   -- null for version 0
    double gb_scale;                    // -2k
    const float4* dacc;                 // Born-radius derivative pair force (xyz)          -- null for version 0
    const float4* gacc;                 // force of the W+U tree sweep (xyz)                -- null for version 0
    float inv_roffset;                  // nu = +gamma/roffset (enlarged radii), -gamma/roffset (vdW radii)
    float* out_f32;                     // layout 0: float[3n] interleaved, +=
    unsigned long long* out_fixed;      // layout 1: OpenMM fixed point [3][padded_n], atomic +=
    float* out_set;                     // internal: float[3n] interleaved, = (host path)
    int padded_n;
    double* scalars;                    // SC_* terms
    int* status;                        // capacity-overflow bits of this evaluation: nothing is delivered unless 0
    int sharded;                        // scalars[SC_FAULT] holds the number of shards whose status word is non-zero (k_status_fold + ENERGY exchange)
    const int* peer_fault;              // peer-memory exchange: non-zero once a wait for a peer has timed out (sticky), or null
    const float4* posq;                 // sorted positions of this evaluation
    float4* posq_ref;                   // reference positions of the pair masks: moved here when k_born rebuilt them
    int* pq_ctl;                        // PairUnits::ctl
    int* tree_ok_out;                   // build evaluations: 1 if the tree was built without overflow (read by k_tree_rescan)
    double* energy_accum;               // optional device accumulator (+=); a float accumulator if energy_f32
    double* energy_out;                 // optional device/pinned-mapped slot (=)
    const int* io;                      // particle -> position in the caller's force buffers, or null (see PrepArgs)
    int energy_f32;
    int* tail_out;                      // host path: pinned mirror of the slab's [scalars | counters | control words] span, written
    int tail_words;                     // by CTA 0 at the very end (one device->host copy less per call), or null
};

// sharded evaluations: this shard's status word, as 0/1, into the energy scalars, which the ENERGY exchange sums over the shards
struct StatusFoldArgs { const int* status; double* scalars; };
__global__ void k_status_fold(StatusFoldArgs A) {
    pdl_release();
    pdl_acquire();
    if (threadIdx.x == 0) A.scalars[SC_FAULT] = *A.status != 0 ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(256) k_finish(FinishArgs A) {
    pdl_release();
    pdl_acquire();
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (A.pq_ctl) {
        // Verlet lists (pair masks: k_born; level-2 candidate lists: k_tree) rebuilt in this evaluation are valid from now on,
        // relative to these positions; the reference positions are shared, so a list that was NOT rebuilt while they move
        // becomes void.  (Only thread 0 writes words 0, 1, 3 and nobody in this kernel reads them; 2 and 4 are stable here.)
        // (a search whose list capacity overflowed left truncated lists: they stay void until the host has grown them)
        const bool pq = A.pq_ctl[2] != 0, l2 = A.pq_ctl[4] != 0 && !(A.status && (*A.status & ST_NBR_OVERFLOW));
        if ((pq || l2) && k < A.np) A.posq_ref[k] = A.posq[k];
        if (k == 0) {
            if (pq || l2) { A.pq_ctl[0] = pq ? 1 : 0; A.pq_ctl[3] = l2 ? 1 : 0; }
            A.pq_ctl[1] = 0;
            A.pq_ctl[5] += pq ? 1 : 0; A.pq_ctl[6] += l2 ? 1 : 0; A.pq_ctl[7] += 1;     // statistics: rebuilds / evaluations
        }
    }
    // sharded: every shard withholds the delivery when ANY shard overflowed (the shards' status words were summed into
    // SC_FAULT by the ENERGY exchange), and the host of a shard that did not overflow itself learns it from ST_PEER_OVERFLOW
    const bool mine = A.status && *A.status != 0;
    const bool peers = A.sharded && A.scalars[SC_FAULT] != 0.0;
    const bool dead = A.peer_fault && *A.peer_fault != 0;
    const bool bad = mine || peers || dead;
    if (k == 0 && A.tree_ok_out) *A.tree_ok_out = bad ? 0 : 1;
    if (bad) {
        __syncthreads();                // every thread of CTA 0 has read the status word before thread 0 changes it
        if (k == 0 && !mine) atomicOr(A.status, dead ? ST_PEER_TIMEOUT : ST_PEER_OVERFLOW);
    } else {
        if (k == 0) {
            // E1 + E2 (ReferenceAGBNPKernels.cpp:188,233,266) + GB + vdW
            const double e = (A.scalars[SC_EVOL_L] - A.scalars[SC_EVOL_S])*(double) A.inv_roffset + A.scalars[SC_EGB] + A.scalars[SC_EVDW];
            A.scalars[SC_TOTAL] = e;
            if (A.energy_out) *A.energy_out = e;
            if (A.energy_accum) { if (A.energy_f32) atomicAdd((float*) A.energy_accum, (float) e); else atomicAdd(A.energy_accum, e); }
        }
        const int o = k < A.np ? A.orig[k] : -1;
        if (o >= 0) {
            // force = -gradient: -(grad_L - grad_S)/roffset
            const float4 l = A.accL[k], s = A.accS[k];
            double fx = ((double) s.x - (double) l.x)*(double) A.inv_roffset, fy = ((double) s.y - (double) l.y)*(double) A.inv_roffset,
                   fz = ((double) s.z - (double) l.z)*(double) A.inv_roffset;
            if (A.gbacc) {
                const float4 g = A.gbacc[k], d = A.dacc[k], t = A.gacc[k];
                fx += A.gb_scale*(double) g.x + (double) d.x + (double) t.x; fy += A.gb_scale*(double) g.y + (double) d.y + (double) t.y;
                fz += A.gb_scale*(double) g.z + (double) d.z + (double) t.z;
            }
            if (A.out_set) { A.out_set[3*o+0] = (float) fx; A.out_set[3*o+1] = (float) fy; A.out_set[3*o+2] = (float) fz; }
            const int dst = A.io ? A.io[o] : o;
            if (A.out_f32) { A.out_f32[3*dst+0] += (float) fx; A.out_f32[3*dst+1] += (float) fy; A.out_f32[3*dst+2] += (float) fz; }
            if (A.out_fixed) {
                atomicAdd(&A.out_fixed[dst], (unsigned long long) (long long) (fx*FORCE_SCALE));
                atomicAdd(&A.out_fixed[(size_t) A.padded_n+dst], (unsigned long long) (long long) (fy*FORCE_SCALE));
                atomicAdd(&A.out_fixed[2*(size_t) A.padded_n+dst], (unsigned long long) (long long) (fz*FORCE_SCALE));
            }
        }
    }
    if (A.tail_out && blockIdx.x == 0) {
        // host path: the evaluation's scalars, counters and control words (status, high-water marks) go straight to the pinned
        // mirror the host reads after its one synchronisation -- everything in them is final: the earlier kernels have
        // completed, and thread 0 of this CTA wrote the total energy / the peer bits of the status word above
        __syncthreads();
        const volatile int* src = (const volatile int*) A.scalars;
        for (int i = threadIdx.x; i < A.tail_words; i += blockDim.x) A.tail_out[i] = src[i];
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
