// GaussVol overlap-tree kernels for sm_100a.
//
// Semantics follow gaussvol/gaussvol.cpp of the reference (citations inline); the decomposition does not:
//   * the reference builds one global tree by depth-first recursion; here every heavy atom's subtree (all overlaps whose
//     lowest-index atom is that atom) is owned by ONE WARP, which builds it breadth-first, level by level
//     (candidate enumeration by warp prefix sums, acceptance compaction by ballots), so subtrees never communicate;
//   * the large-radius build (S1), the vdW-radius rescan (S3) and both up-sweeps (S2, second half of S3) are fused:
//     a node's vdW-radius Gaussian is computed when the node is created, and two bottom-up passes over the finished
//     subtree yield both sets of self-volumes and both surface-tension gradients (the energies need no tree
//     accumulation: E = sum over nodes of coef * gamma_1..n * volume);
//   * topology-deciding arithmetic (overlap volume, inclusion threshold, sibling sort key) is FP64 with the expression
//     structure of gaussvol.cpp:60-93; everything fed to energies/forces downstream is rounded to FP32.
// Memory plan per warp:
//   shared memory  -- everything on the dependent chain of the build: level-2 neighbor list, per-node parent/atom/child
//                     ranges (shorts), and for the level being expanded the candidate prefix sums, sort keys and sibling
//                     permutation.  (If a system needs capacities that
//                     do not fit, the same code runs with these arrays in a per-warp global scratch: TreeArgs::wk_global.)
//   global staging -- streamed, coalesced, one round trip per level: the nodes' Gaussians (read once by their children's
//                     candidates), the per-node sweep records and the children-to-parent sums of the sweeps.
// Only what the later "gamma" sweep (S10+S11 merged, linear in nu) needs is persisted compactly (TreeStore), by the
// bottom-up sweep itself.  The global staging is what bounds these kernels (ncu): every pass over a level
// waits for an L2 round trip with one chunk of 32 nodes in flight per warp -- passes over staged data are what to avoid.
// k_tree_rescan (opt-in tree reuse) re-evaluates a stored tree at new positions without any search.
#ifndef AGBNP_TREE_CUH_
#define AGBNP_TREE_CUH_

#include "agbnp_device.cuh"

namespace agbnp_b200_impl {

constexpr int MAX_LEVELS = 10;      // level index 1..8 used (MAX_ORDER 8)
constexpr int TREE_GROUP_MAX = 4;   // most roots one work item may carry
// work item (int2): x = index of its first root in item_roots[], y = number of roots | part << 8 | parts << 16
__host__ __device__ inline int item_nroots(int2 it) { return it.y & 0xff; }
__host__ __device__ inline int item_part(int2 it) { return (it.y >> 8) & 0xff; }
__host__ __device__ inline int item_parts(int2 it) { return it.y >> 16; }
#ifndef SCREEN_UNROLL
#define SCREEN_UNROLL 1             // candidates per lane and trip of the FP32 screen (measured: 1 is fastest on B200)
#endif

// persisted per-node records for the gamma sweep and for the topology dump: two float4 per node,
//   rec[2*o]   = (coefp*sfp, dvv1, a_i/a_1i, bits of the sorted index of the node's last atom)        (vdW radii)
//   rec[2*o+1] = (dv1 x, y, z, bits of: parent slot (low 16 bits, 0xffff for the root) | has-children flag << 16)
// The unit of tree work is an ITEM, one of two kinds (the host decides at sort time, agbnp_b200.cu: build_order):
//   * (root atom, part k of K): the subtrees below different level-2 nodes of a root never interact, so a large root is
//     split into K parts; part k builds the root and all of its level-2 nodes (cheap, and needed for the sibling lists) but
//     owns and expands only the level-2 nodes at sorted positions t with t mod K == k.  The root's own terms belong to
//     part 0.  Everything a sweep adds is linear in the owned nodes, so the parts just add up.
//   * a GROUP of up to TREE_GROUP_MAX small roots that are neighbors in the sorted order: subtrees of different roots never
//     interact either, so the group is built as one forest -- slots 0..G-1 are the roots (level 1), every later level holds
//     the nodes of all of them, parents first -- and the per-level passes run on chunks that are fuller than any of the
//     roots would fill alone.  Positions are relative to the first root of the item.
struct TreeStore {
    int cap;                 // node capacity
    int* cursor;             // bump allocator
    int* root_off;           // [items] first node of the item's subtree (slot 0 = the root atom itself)
    int* root_cnt;           // [items] nodes stored for the item including slot 0; 0 if not built
    short* root_lvs;         // [items*MAX_LEVELS] first slot of each level, root_lvs[i*MAX_LEVELS+l], l = 1..nlev+1
    float4* rec;             // [2*cap]
    short* rank;             // [cap] rank among siblings (topology dump only)
};

// a node's two Gaussians (enlarged / vdW radii) in the root's frame: positions are relative to the root atom
struct __align__(16) NodeGauss {
    double aL, vL, xL, yL, zL;   // the FP32 screen reads these five and rounds them itself: float copies would make the
    double aS, vS, xS, yS, zS;   // record 112 bytes, and the staged bytes per node are what the kernel pays for
    float gam;                   // gamma_1..n
    float pad[3];
};
static_assert(sizeof(NodeGauss) == 96, "NodeGauss layout");

constexpr int BLIST_MAX = 64;       // neighbor blocks listed per heavy block (more: the root scans all blocks)
#ifndef BLIST_MIN_BLOCKS
#define BLIST_MIN_BLOCKS 2048        // systems with fewer heavy blocks skip k_blocklist: their roots scan every block's box when they search
#endif

// ---------------------------------------------------------------------------------------------------------------
// k_blocklist: for every heavy block, the heavy blocks whose bounding box comes within the largest level-2 pair radius
// of its own box, in ascending order -- the roots of the block then test ~10 boxes instead of all of them
// ---------------------------------------------------------------------------------------------------------------
struct BlockListArgs {
    int nhb;
    const float4 *bbc, *bbh;
    float rc2;                  // largest conservative pair radius (+ list skin), squared
    int* bcount;                // [nhb] number of listed blocks, -1 = more than BLIST_MAX
    unsigned short* blist;      // [nhb*BLIST_MAX]
    const int* ctl;             // neighbor-list reuse control words (TreeArgs::ctl): the block lists are only needed by an
    float move2;                // evaluation that rebuilds the level-2 candidate lists
};

// control words shared by every Verlet-style list of an evaluation (pair masks: agbnp_pair.cuh; level-2 candidate lists: here)
enum ListCtl { LC_PQ_VALID = 0, LC_DISP2 = 1, LC_PQ_REBUILT = 2, LC_L2_VALID = 3, LC_L2_REBUILT = 4,
               LC_N_PQ = 5, LC_N_L2 = 6, LC_N_EVAL = 7 /* statistics since the lists were last voided */, LC_COUNT = 8 };

__global__ void __launch_bounds__(256) k_blocklist(BlockListArgs A) {
    pdl_release();
    pdl_acquire();
    if (A.ctl[LC_L2_VALID] != 0 && !(__int_as_float(A.ctl[LC_DISP2]) > A.move2)) return;     // k_tree will walk its stored lists
    const int lane = threadIdx.x & 31;
    const int rb = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
    if (rb >= A.nhb) return;
    const float4 ca = A.bbc[rb], ha = A.bbh[rb];
    int n = 0;
    for (int b0 = 0; b0 < A.nhb; b0 += 32) {
        const int b = b0+lane;
        const bool hit = b < A.nhb && box_box_dist2(ca, ha, A.bbc[b], A.bbh[b]) < A.rc2;
        const unsigned m = __ballot_sync(FULL, hit);
        if (hit) {
            const int p = n + __popc(m & lanemask_lt());
            if (p < BLIST_MAX) A.blist[rb*BLIST_MAX + p] = (unsigned short) b;
        }
        n += __popc(m);
    }
    if (lane == 0) A.bcount[rb] = n <= BLIST_MAX ? n : -1;
}

struct TreeArgs {
    int nh, nhb, np;
    const int2* items;                // [nitems] work items, most expensive first
    const int* item_roots;            // sorted indices of the items' roots
    int nitems;
    const int* bcount;                // k_blocklist output, or null: no block lists (every root scans all heavy blocks)
    const unsigned short* blist;
    const float4* posq;
    const int* orig;
    const int4* l2rec;                // (caller's index | radius bin << 24, bits of float aL, bits of float vL, -): one load serves
                                      // both level-2 tests and the screen's operands
    const unsigned char* rcbin;
    const double *aL, *vL, *aS, *vS;
    const float* gamma;
    const float4 *bbc, *bbh;
    const float* rc2;
    const float* rc2max;
    // Level-2 candidate lists kept between evaluations (Verlet style): built with the pair radii enlarged by a skin, valid
    // while no atom has moved more than skin/2 since (k_prep measures it), every listed candidate re-tested against the
    // exact radius in every evaluation -- the level-2 candidates of an evaluation, and their order, are exactly those of
    // a fresh search.
    const float* rc2s;                // [nbins*nbins] (pair radius + skin)^2
    const float* rc2maxs;             // [nbins]
    int* l2list;                      // [nhp*nbrmax] per root (sorted index): candidate atoms, in search order
    int* l2cnt;                       // [nhp]
    int* ctl;                         // ListCtl
    float move2;                      // (0.49 skin)^2, or < 0: search in every evaluation
    int nbins;
    double volmina, volminb, min_gvol, swd;
    float screen;                     // FP32 screen threshold: VOLMINA less the screen's error margin
    int max_order;
    int cap, wcap, nbrmax;            // capacities: nodes per root, nodes per level, level-2 neighbors per root
    unsigned char* stage;             // per-warp global staging (tree_stage_bytes(cap, wcap) each)
    size_t stage_stride;
    unsigned char* wk_global;         // per-warp work arrays in global memory, or nullptr = shared memory
    size_t wk_stride;
    float4 *accL, *accS;              // [np] out: (gradient x,y,z of sum coef*gamma*vol, self volume), enlarged / vdW radii
    double* scalars;
    unsigned long long* counters;
    TreeStore st;
    int* work_counter;
    int* status;
    int shard_rank, shard_count;      // roots are dealt to shards block-cyclically (blocks of 32 sorted heavy atoms)
    int *hw_nbr, *hw_nodes, *hw_width;   // high-water marks: level-2 neighbors / nodes / widest level of one root
};

__host__ __device__ inline size_t tree_work_bytes(int nbrmax, int cap, int wcap) {
    size_t b = (size_t) wcap*sizeof(float4)                   // sc4
             + (size_t) (wcap+2)*sizeof(double)               // key / pl (aliased: never live at the same time)
             + (size_t) wcap*sizeof(float)                    // scv
             + (size_t) wcap*sizeof(int)                      // cand
             + (size_t) nbrmax*6*sizeof(float)                // nbx, nby, nbz, nba, nbv, nbi
             + (size_t) (MAX_LEVELS+2)*sizeof(int)            // lvs
             + (size_t) 2*(TREE_GROUP_MAX+1)*sizeof(int)      // rt, nboff
             + (size_t) ((cap+31)/32)*sizeof(unsigned)        // hc
             + (size_t) (2*cap + 4*wcap)*sizeof(short);       // parent, nbr | cstart, ccount, perm, gend
    return (b + 15) & ~(size_t) 15;
}
__host__ __device__ inline size_t tree_stage_bytes(int cap, int wcap) {
    // k_tree: two windows of wcap Gaussians (reused for the hand-up sums) + two sweep records and a rank per node;
    // k_tree_rescan: Gaussians, two sweep records and hand-up sums per node
    const size_t build = (size_t) 2*wcap*sizeof(NodeGauss) + (size_t) cap*(4*sizeof(float4) + sizeof(short));
    const size_t rescan = (size_t) cap*(sizeof(NodeGauss) + 8*sizeof(float4));
    const size_t b = build > rescan ? build : rescan;
    return (b + 255) & ~(size_t) 255;
}

// polynomial switching function and derivative (gaussvol.cpp:18-41)
__device__ __forceinline__ void pol_switch(double gvol, double volmina, double volminb, double swd, double& s, double& sp) {
    if (gvol > volminb) { s = 1.0; sp = 0.0; }
    else if (gvol < volmina) { s = 0.0; sp = 0.0; }
    else {
        const double u = (gvol-volmina)*swd;
        const double u2 = u*u;
        s = u*u2*(10.0 - 15.0*u + 6.0*u2);
        sp = swd*30.0*u2*(1.0 - 2.0*u + u2);
    }
}

// Gaussian overlap volume V12 = V1 V2 (df/pi)^{3/2} exp(-df d^2), df = a1 a2/(a1+a2)   (gaussvol.cpp:60-93)
__device__ __forceinline__ double overlap_volume(double a1, double v1, double a2, double v2, double d2,
                                                 double& deltai, double& df) {
    const double a12 = a1+a2;
    deltai = 1.0/a12;
    df = a1*a2*deltai;
    const double ef = exp(-df*d2);
    const double u = df*0.31830988618379067154;   // df/pi ;  pow(pi/df,1.5)^-1 = u*sqrt(u)
    return (v1*v2)*(u*sqrt(u))*ef;
}

// per-warp work arrays (shared memory, or global scratch for oversize capacities)
struct TreeWork {
    float4* sc4;             // [wcap] FP32 copy (x, y, z, a) of the enlarged-radius Gaussian of every node of the level being
    float* scv;              //        expanded, and its volume: all the FP32 screen reads of a parent (written at node creation:
                             //        the screen of a level is complete before the exact phase creates the next one, in place)
    double* key;             // [wcap] sort key (switched volume) of the level being created (exact phase + sort)
    int2* pl;                // [wcap+1] ALIASES key (screen only): the expanding parents of the level, compacted:
                             //        (sorted position, first candidate index); pl[np] = (-, number of candidates)
    int* cand;               // [wcap] candidates that passed the FP32 screen: parent slot | neighbor index << 16
    float *nbx, *nby, *nbz;  // [nbrmax] level-2 candidate positions (absolute, float as given)
    float *nba, *nbv;        // [nbrmax] their enlarged-radius Gaussian exponent / volume rounded to float (screen only)
    int* nbi;                // [nbrmax] their sorted atom indices
    int* lvs;                // [MAX_LEVELS+2] first slot of each level
    int* rt;                 // [TREE_GROUP_MAX+1] sorted atom indices of the item's roots (= slots 0 .. G-1)
    int* nboff;              // [TREE_GROUP_MAX+1] first level-2 candidate of each root in nb*[]; nboff[G] = their number
    unsigned* hc;            // [cap/32] bit per slot: the node has children
    short *parent, *nbr;     // [cap] parent slot, level-2 neighbor index + 1 of the node's last atom
    short *cstart, *ccount;  // [wcap] first child slot / end of the child range of the nodes of the level being expanded (by
                             //        slot - ls); k_tree_rescan: ccount is [cap], 1 if the node has children
    short *perm, *gend;      // [wcap] sorted position -> slot, end of the sibling group (both relative to the level start)
    __device__ void bind(unsigned char* base, int nbrmax, int cap, int wcap) {
        sc4 = (float4*) base;
        key = (double*) (sc4+wcap);
        pl = (int2*) key;
        scv = (float*) (key+wcap+2);
        cand = (int*) (scv+wcap);
        nbx = (float*) (cand+wcap); nby = nbx+nbrmax; nbz = nby+nbrmax; nba = nbz+nbrmax; nbv = nba+nbrmax;
        nbi = (int*) (nbv+nbrmax);
        lvs = nbi+nbrmax;
        rt = lvs+MAX_LEVELS+2; nboff = rt+TREE_GROUP_MAX+1;
        hc = (unsigned*) (nboff+TREE_GROUP_MAX+1);
        parent = (short*) (hc+(cap+31)/32); nbr = parent+cap; cstart = nbr+cap; ccount = cstart+wcap;
        perm = ccount+wcap; gend = perm+wcap;
    }
};

// per-warp global staging (sweep records, children sums): written once and read once by the same warp within one item.
// TREE_STAGE_STREAM (off): st.global.cs / ld.global.cs (evict-first), so that the staging churn of the whole grid does not push
// the persisted TreeStore -- which k_tree_gamma reads 300 us later -- out of L2.  Measured (r2aa): k_tree_gamma 41.1 -> 39.4 us,
// but k_tree 151.5 -> 156.1 us (its own re-reads then miss more often): not enabled.
#ifdef TREE_STAGE_STREAM
#define STAGE_ST(p, v) __stcs((p), (v))
#define STAGE_LD(p) __ldcs((p))
#else
#define STAGE_ST(p, v) (*(p) = (v))
#define STAGE_LD(p) (*(p))
#endif

// position of a lane within its segment (lanes with equal key are contiguous) and the largest position in the warp:
// a segmented scan then needs only the steps d <= maxpos (sibling groups are short: usually 2-3 of the 5 steps), and a
// lane takes the value d lanes below iff pos >= d
__device__ __forceinline__ int seg_position(int key, int lane, int& maxpos) {
    const int kprev = __shfl_up_sync(FULL, key, 1);
    const unsigned heads = __ballot_sync(FULL, lane == 0 || kprev != key);
    const int pos = lane - (31 - __clz(heads & (0xffffffffu >> (31-lane))));
    maxpos = __reduce_max_sync(FULL, pos);
    return pos;
}

// segmented inclusive scan step: lanes with equal key are contiguous; adds the value `d` lanes below if it belongs to the
// same segment
__device__ __forceinline__ void seg_step(float (&v)[10], bool take, int d) {
#pragma unroll
    for (int c = 0; c < 10; c++) {
        const float t = __shfl_up_sync(FULL, v[c], d);
        if (take) v[c] += t;
    }
}

__device__ __forceinline__ void hu_store(float4* hu, int node, const float (&v)[10]) {
    STAGE_ST(&hu[4*node], make_float4(v[0], v[1], v[2], v[3]));
    STAGE_ST(&hu[4*node+1], make_float4(v[4], v[5], v[6], v[7]));
    STAGE_ST(&hu[4*node+2], make_float4(v[8], v[9], 0.f, 0.f));
}

// the bottom-up sweep over a finished subtree, both radius sets at once (gaussvol.cpp:400-487): self-volumes and the
// gradients of sum coef*gamma*vol w.r.t. every atom of the subtree, added to accL/accS[atom] as one vector red.global
// per node and radius set.   sw[2*sl] = (vol, sfp, dvv1, a_i/a_1i), sw[2*sl+1] = (dv1 x, y, z, gamma_1..n).
// A node hands (psi, dvv1 F, dv1 F + P a_1/a_1i) to its parent (gaussvol.cpp:476-484).  Siblings are contiguous, so the
// sums over a parent's children are a segmented warp scan over the child level; the last child of each parent stores
// the totals into hu[4*parent ..] (global staging, one plain store per parent, read back one level later).
// STORED: the node's last atom comes from `ja_arr` (k_tree_rescan) instead of the level-2 neighbor list.
// hu: STORED (k_tree_rescan): one entry per slot.  Build (k_tree): the sums of a level are only needed while its parents' level
// is swept, so they live in two alternating windows of `hu_w` entries indexed by the slot relative to its level start --
// the bytes a subtree keeps hot in L2 are what these kernels pay for.
template <bool STORED>
__device__ __forceinline__ void tree_sweep(const TreeWork& W, const float4* swL, const float4* swS, float4* hu, int nlev, int r, int lane,
                                           float4* accL, float4* accS, const int* ja_arr = nullptr,
                                           float4* rec_out = nullptr, short* rank_out = nullptr, const short* rk = nullptr, int hu_w = 0) {
    for (int lev = nlev; lev >= 1; lev--) {
        const int b = W.lvs[lev], e = W.lvs[lev+1];
        // window of this level's own sums (read) and of its parents' (written): slot -> hu index
        const int rd_off = STORED ? 0 : (lev & 1)*hu_w - b;
        const int wr_off = STORED ? 0 : (lev > 1 ? ((lev-1) & 1)*hu_w - W.lvs[lev-1] : 0);
        const float coefp = ((lev & 1) ? 1.f : -1.f)/(float) lev;
        int carry_key = -2;                         // parent whose children run across the chunk boundary
        float carry[10];
#pragma unroll
        for (int c = 0; c < 10; c++) carry[c] = 0.f;
        for (int s0 = b; s0 < e; s0 += 32) {
            const int sl = s0+lane;
            const bool valid = sl < e;
            int key = -3-lane;                      // distinct per lane: never merges
            float v[10];
#pragma unroll
            for (int c = 0; c < 10; c++) v[c] = 0.f;
            if (valid) {
                const float4 l0 = STAGE_LD(&swL[2*sl]), l1 = STAGE_LD(&swL[2*sl+1]), s0v = STAGE_LD(&swS[2*sl]), s1v = STAGE_LD(&swS[2*sl+1]);
                float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0, h2 = h0;
                const bool has_kids = STORED ? W.ccount[sl] > 0 : ((W.hc[sl >> 5] >> (sl & 31)) & 1u) != 0;
                if (has_kids) { const float4* hs = hu + 4*(sl+rd_off); h0 = STAGE_LD(hs); h1 = STAGE_LD(hs+1); h2 = STAGE_LD(hs+2); }
                int ja;
                if (STORED) ja = ja_arr[sl];
                else { const int ia = W.nbr[sl]; ja = ia == 0 ? W.rt[sl] : W.nbi[ia-1]; }     // ia == 0: a root, slot = its index
                key = W.parent[sl];
                if (!STORED && rec_out) {
                    // persist what the gamma sweep needs (TreeStore) while the records are in registers anyway
                    const int pk = (key & 0xffff) | (has_kids ? 0x10000 : 0);
                    rec_out[2*sl] = make_float4(coefp*s0v.y, s0v.z, s0v.w, __int_as_float(ja));
                    rec_out[2*sl+1] = make_float4(s1v.x, s1v.y, s1v.z, __int_as_float(pk));
                    rank_out[sl] = rk[sl];
                }
                {   // enlarged radii
                    const float ps = coefp*l0.x + h0.x, F = coefp*l0.y*l1.w + h0.y, px = h0.z, py = h0.w, pz = h1.x;
                    const float c2a = l0.w, c2b = 1.f-l0.w;
                    // gradient w.r.t. the node's last atom (gaussvol.cpp:467-474)
                    atomicAdd(&accL[ja], make_float4(fmaf(px, c2a, -l1.x*F), fmaf(py, c2a, -l1.y*F), fmaf(pz, c2a, -l1.z*F), ps));
                    v[0] = ps; v[1] = l0.z*F; v[2] = fmaf(px, c2b, l1.x*F); v[3] = fmaf(py, c2b, l1.y*F); v[4] = fmaf(pz, c2b, l1.z*F);
                }
                {   // vdW radii
                    const float ps = coefp*s0v.x + h1.y, F = coefp*s0v.y*s1v.w + h1.z, px = h1.w, py = h2.x, pz = h2.y;
                    const float c2a = s0v.w, c2b = 1.f-s0v.w;
                    atomicAdd(&accS[ja], make_float4(fmaf(px, c2a, -s1v.x*F), fmaf(py, c2a, -s1v.y*F), fmaf(pz, c2a, -s1v.z*F), ps));
                    v[5] = ps; v[6] = s0v.z*F; v[7] = fmaf(px, c2b, s1v.x*F); v[8] = fmaf(py, c2b, s1v.y*F); v[9] = fmaf(pz, c2b, s1v.z*F);
                }
            }
            if (lev == 1) break;                    // the root hands nothing up
            // children of the carried parent continue in this chunk?
            const int key0 = __shfl_sync(FULL, key, 0);
            if (carry_key >= 0) {
                if (key0 == carry_key) {
                    if (lane == 0) {
#pragma unroll
                        for (int c = 0; c < 10; c++) v[c] += carry[c];
                    }
                } else if (lane == 0) hu_store(hu, carry_key+wr_off, carry);
            }
            int maxpos;
            const int pos = seg_position(key, lane, maxpos);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                if (d > maxpos) break;
                seg_step(v, pos >= d, d);
            }
            const int kn = __shfl_down_sync(FULL, key, 1);
            const bool tail = valid && (lane == 31 || kn != key);
            const int last = min(31, e-1-s0);       // last valid lane of this chunk: its segment may continue
            if (tail && lane != last) hu_store(hu, key+wr_off, v);
            carry_key = __shfl_sync(FULL, key, last);
#pragma unroll
            for (int c = 0; c < 10; c++) carry[c] = __shfl_sync(FULL, v[c], last);
        }
        if (lev > 1 && carry_key >= 0 && lane == 0) hu_store(hu, carry_key+wr_off, carry);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_tree: build + vdW rescan + up-sweeps for every heavy root atom (reference S1-S3:
// gaussvol.cpp:103-250,254-327,389-519,589-606; ReferenceAGBNPKernels.cpp:290-380)
// ---------------------------------------------------------------------------------------------------------------
// SMEM_WORK: work arrays in shared memory, CTAs of 2 warps (the normal case; TREE_SMEM_CTAS per SM bounds the registers);
// otherwise work arrays in global scratch, CTAs of 8 warps
#ifndef TREE_WARPS
#define TREE_WARPS 2                // warps per CTA of the shared-memory instantiation (warps never cooperate)
#endif
#ifndef TREE_SMEM_CTAS
#define TREE_SMEM_CTAS (16/TREE_WARPS)
#endif
template <bool SMEM_WORK>
__global__ void __launch_bounds__(SMEM_WORK ? 32*TREE_WARPS : 256, SMEM_WORK ? TREE_SMEM_CTAS : 2) k_tree(TreeArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_release();
    pdl_acquire();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int nbrmax = A.nbrmax, cap = A.cap, wcap = A.wcap;
    const int gwarp = blockIdx.x*nwarp + warp;

    // the address space of the work arrays is a compile-time fact: in the shared-memory instantiation every access is an
    // LDS/STS with 32-bit addressing instead of a generic LD/ST
    TreeWork W;
    if (SMEM_WORK) W.bind(smem_raw + (size_t) warp*tree_work_bytes(nbrmax, cap, wcap), nbrmax, cap, wcap);
    else W.bind(A.wk_global + (size_t) gwarp*A.wk_stride, nbrmax, cap, wcap);
    // per-warp global staging.  The Gaussians of a level are read only by the candidates of the next one, so they live in two
    // alternating windows of wcap records (by level parity, indexed by slot - level start); the sweep's hand-up sums reuse
    // the same bytes afterwards (tree_sweep).  What stays per node until the sweep are the two sweep records and the rank.
    unsigned char* stage = A.stage + (size_t) gwarp*A.stage_stride;
    NodeGauss* G = (NodeGauss*) stage;              // [2*wcap]
    float4* hu = (float4*) stage;                   // [2*wcap*4] after the build
    float4* swL = (float4*) (G + 2*(size_t) wcap);
    float4* swS = swL + 2*(size_t) cap;
    short* rk = (short*) (swS + 2*(size_t) cap);

    double eL_tot = 0, eS_tot = 0, vsumL = 0, vsumS = 0;     // per-lane partial sums, reduced at the end
    unsigned long long c2_tot = 0, c3_tot = 0, m_tot = 0;
    int hw_nn = 0, hw_slots = 0, hw_w = 0;
    // search the level-2 candidates or walk the stored lists?  (nothing in this kernel writes LC_L2_VALID / LC_DISP2)
    const bool rebuild_l2 = A.ctl[LC_L2_VALID] == 0 || __int_as_float(A.ctl[LC_DISP2]) > A.move2;
    if (rebuild_l2 && blockIdx.x == 0 && threadIdx.x == 0) A.ctl[LC_L2_REBUILT] = 1;

    // raw claim -> item: with shards, items are dealt block-cyclically (blocks of 32 in longest-first order)
    const int raw_end = A.shard_count > 1 ? (A.nitems+TILE-1)/TILE*TILE : A.nitems;
    tail_begin(3);
    for (int raw = first_unit(); ; raw = next_unit(A.work_counter, lane)) {
        int item = raw;
        if (A.shard_count > 1) {
            item = ((raw >> 5)*A.shard_count + A.shard_rank)*TILE + (raw & 31);
            if (item >= raw_end) break;
            if (item >= A.nitems) continue;
        } else if (item >= A.nitems) break;
        const int2 itm = A.items[item];
        const int ng = item_nroots(itm);
#define part item_part(itm)
#define nparts item_parts(itm)
        const float4 pr = A.posq[A.item_roots[itm.x]];                // frame of the item: its first root
        __syncwarp();                                                 // the previous item is done with the work arrays
        if (lane < ng) W.rt[lane] = A.item_roots[itm.x+lane];

        // ---- level-2 candidate lists: per root, heavy atoms later in the caller's order within the conservative pair radius ----
        int nn = 0, nl_tot = 0;
        for (int g = 0; g < ng; g++) {
            const int r = A.item_roots[itm.x+g];
            const float4 prg = A.posq[r];
            const int orig_r = A.orig[r];
            const int rb = A.rcbin[r];
            if (lane == 0) W.nboff[g] = min(nn, nbrmax);
            if (rebuild_l2) {
                // search: block lists -> atoms; candidates within the enlarged radius go to the root's stored list, those
                // within the radius itself are this evaluation's candidates
                int nl = 0;
                const float rcmaxs = A.rc2maxs[rb];
                int* mylist = A.l2list + (size_t) r*nbrmax;
                const int nlist = A.bcount ? A.bcount[r >> 5] : -1;
                const int nscan = nlist >= 0 ? nlist : A.nhb;
                for (int b0 = 0; b0 < nscan; b0 += 32) {
                    int b = b0+lane;
                    bool hit = false;
                    if (b < nscan) {
                        if (nlist >= 0) b = A.blist[(r >> 5)*BLIST_MAX + b];
                        hit = point_box_dist2(prg.x, prg.y, prg.z, A.bbc[b], A.bbh[b]) < rcmaxs;
                    }
                    unsigned m = __ballot_sync(FULL, hit);
                    while (m) {                                           // four listed blocks per trip: their loads overlap
                        int jq[4];
                        bool act[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            act[q] = m != 0;
                            const int i = act[q] ? __ffs(m)-1 : 0;
                            m &= m-1;
                            jq[q] = __shfl_sync(FULL, b, i)*TILE+lane;
                        }
                        float4 pj[4];
                        int4 rc[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) { pj[q] = A.posq[jq[q]]; rc[q] = A.l2rec[jq[q]]; }
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int ob = rc[q].x;
                            const float dx = pj[q].x-prg.x, dy = pj[q].y-prg.y, dz = pj[q].z-prg.z;
                            const float d2 = dx*dx + dy*dy + dz*dz;
                            const int cls = rb*A.nbins + ((ob >> 24) & 0x7f);
                            const bool later = act[q] && ((ob < 0 ? -1 : (ob & 0xffffff)) > orig_r);
                            const bool listed = later && d2 < __ldg(A.rc2s + cls);
                            const bool ok = listed && d2 < __ldg(A.rc2 + cls);
                            const unsigned lm = __ballot_sync(FULL, listed);
                            const unsigned am = __ballot_sync(FULL, ok);
                            if (listed) {
                                const int p = nl + __popc(lm & lanemask_lt());
                                if (p < nbrmax) mylist[p] = jq[q];
                            }
                            if (ok) {
                                const int p = nn + __popc(am & lanemask_lt());
                                if (p < nbrmax) {
                                    W.nbi[p] = jq[q]; W.nbx[p] = pj[q].x; W.nby[p] = pj[q].y; W.nbz[p] = pj[q].z;
                                    W.nba[p] = __int_as_float(rc[q].y); W.nbv[p] = __int_as_float(rc[q].z);
                                }
                            }
                            nl += __popc(lm);
                            nn += __popc(am);
                        }
                    }
                }
                if (lane == 0) A.l2cnt[r] = min(nl, nbrmax);
                nl_tot += nl;
            } else {
                // walk the stored list: exact radius test, same order
                const int nl = A.l2cnt[r];
                const int* mylist = A.l2list + (size_t) r*nbrmax;
                nl_tot += nl;
                for (int e0 = 0; e0 < nl; e0 += 32) {
                    const int e = e0+lane;
                    const bool act = e < nl;
                    const int j = act ? mylist[e] : r;
                    const float4 pj = A.posq[j];
                    const int4 rc = A.l2rec[j];
                    const float dx = pj.x-prg.x, dy = pj.y-prg.y, dz = pj.z-prg.z;
                    const float d2 = dx*dx + dy*dy + dz*dz;
                    const bool ok = act && d2 < __ldg(A.rc2 + rb*A.nbins + ((rc.x >> 24) & 0x7f));
                    const unsigned am = __ballot_sync(FULL, ok);
                    if (ok) {
                        const int p = nn + __popc(am & lanemask_lt());
                        if (p < nbrmax) {
                            W.nbi[p] = j; W.nbx[p] = pj.x; W.nby[p] = pj.y; W.nbz[p] = pj.z;
                            W.nba[p] = __int_as_float(rc.y); W.nbv[p] = __int_as_float(rc.z);
                        }
                    }
                    nn += __popc(am);
                }
            }
        }
        // the lists of all roots of the item share the level-2 capacity (a stored list is at most as long as it)
        hw_nn = max(hw_nn, nl_tot);
        if (nl_tot > nbrmax) {
            if (lane == 0) atomicOr(A.status, ST_NBR_OVERFLOW);
            continue;
        }
        if (lane == 0) W.nboff[ng] = nn;

        // ---- slots 0 .. G-1: the root atoms (gaussvol.cpp:130-148), level 1 ----
        for (int w = lane; w < (cap+31)/32; w += 32) W.hc[w] = 0u;
        __syncwarp();
        int npar1 = 0;                                                // roots with candidates, compacted into pl[] below
        {
            const bool isroot = lane < ng;
            const int rg = isroot ? W.rt[lane] : W.rt[0];
            const int cnt_g = isroot ? W.nboff[lane+1] - W.nboff[lane] : 0;
            const unsigned hm = __ballot_sync(FULL, cnt_g > 0);
            npar1 = __popc(hm);
            if (isroot) {
                const float4 pg = A.posq[rg];
                const float gam_r = A.gamma[rg];
                NodeGauss g;
                g.aL = A.aL[rg]; g.vL = A.vL[rg];
                g.xL = (double) pg.x - (double) pr.x; g.yL = (double) pg.y - (double) pr.y; g.zL = (double) pg.z - (double) pr.z;
                g.aS = A.aS[rg]; g.vS = A.vS[rg]; g.xS = g.xL; g.yS = g.yL; g.zS = g.zL;
                g.gam = gam_r; g.pad[0] = g.pad[1] = g.pad[2] = 0.f;
                G[wcap+lane] = g;                                     // level 1: the odd window
                const float own = part == 0 ? 1.f : 0.f;              // the root's own terms belong to part 0
                swL[2*lane] = make_float4(own*(float) g.vL, 1.f, 1.f, 1.f); swL[2*lane+1] = make_float4(0.f, 0.f, 0.f, gam_r);
                swS[2*lane] = make_float4(own*(float) g.vS, 1.f, 1.f, 1.f); swS[2*lane+1] = make_float4(0.f, 0.f, 0.f, gam_r);
                rk[lane] = 0;
                W.parent[lane] = -1; W.nbr[lane] = 0; W.perm[lane] = (short) lane; W.gend[lane] = (short) (lane+1);
                W.sc4[lane] = make_float4((float) g.xL, (float) g.yL, (float) g.zL, (float) g.aL); W.scv[lane] = (float) g.vL;
                eL_tot += (double) (own*gam_r*(float) g.vL); eS_tot += (double) (own*gam_r*(float) g.vS);
                vsumL += (double) (own*(float) g.vL); vsumS += (double) (own*(float) g.vS);
                // level 1 -> 2 in the candidate enumeration's terms: "parent" g owns candidates nboff[g] .. nboff[g+1]-1
                if (cnt_g > 0) W.pl[__popc(hm & lanemask_lt())] = make_int2(lane, W.nboff[lane]);
            }
            if (lane == 0) { W.lvs[1] = 0; W.pl[npar1] = make_int2(0, nn); }
        }
        __syncwarp();

        // ---- breadth-first build: level -> level+1 ----
        int nslots = ng, ls = 0, le = ng, level = 1;
        bool failed = false;
        while (level < A.max_order) {           // a node at level >= MAX_ORDER gets no children (gaussvol.cpp:211)
            int T, npar = 0;
            const int width = le-ls;
            if (level == 1) {
                T = nn; npar = npar1;           // pl[] was filled with the roots' candidate ranges above
            } else {
                // candidates of the node at sorted position t: its younger siblings t+1 .. gend[t]-1 (gaussvol.cpp:221).
                // The parents that have any are compacted into pl[] with the index of their first candidate, parent-major.
                T = 0;
                for (int t0 = 0; t0 < width; t0 += 32) {
                    const int t = t0+lane;
                    int c = 0;
                    // level 2 -> 3: only the level-2 nodes this part owns are expanded
                    if (t < width && (level > 2 || t % nparts == part)) c = (int) W.gend[t] - t - 1;
                    const int inc = warp_incl_scan(c);
                    const unsigned hm = __ballot_sync(FULL, c > 0);
                    if (c > 0) W.pl[npar + __popc(hm & lanemask_lt())] = make_int2(t, T + inc - c);
                    npar += __popc(hm);
                    T += __shfl_sync(FULL, inc, 31);
                }
                if (lane == 0) W.pl[npar] = make_int2(0, T);
                __syncwarp();
            }
            if (level == 1) c2_tot += (lane == 0 && part == 0) ? (unsigned long long) T : 0ull;
            else c3_tot += (lane == 0) ? (unsigned long long) T : 0ull;

            const int new_start = nslots;
            const float cf = ((level+1) & 1) ? 1.f : -1.f;
            const float coefp = cf/(float) (level+1);
            // ---- phase A: FP32 screen of all T candidates.  A candidate whose FP32 overlap volume is below VOLMINA by
            // more than the screen's error margin is certainly rejected by the exact test (s = 0 below VOLMINA,
            // gaussvol.cpp:26-29,233); everything else goes, in candidate order, to the exact FP64 phase.
            // Everything the screen reads is in shared memory.  Candidate -> (parent, sibling): every expanding parent owns
            // at least one candidate, so the parents that START inside the 32 candidates of a trip are among the next 32 of
            // pl[]; one redux.or builds the mask of their start positions and a popc gives every lane its parent -- no
            // per-candidate search.
            int nmaybe = 0;
            int tp = 0;                                                   // pl index of the parent that owns candidate k0
            for (int k0 = 0; k0 < T; k0 += 32) {
                const int k = k0 + lane;
                const bool valid = k < T;
                int p, kn = valid ? k : 0;
                {
                    const int s = W.pl[min(tp+1+lane, npar)].y - k0 - 1;  // start of parent tp+1+lane, relative to k0+1
                    const unsigned heads = __reduce_or_sync(FULL, (unsigned) s < 32u ? (1u << s) : 0u);
                    const int2 me = W.pl[min(tp + __popc(heads & lanemask_lt()), npar-1)];
                    tp += __popc(heads);
                    p = W.perm[me.x];
                    if (level > 1) {            // the candidate atom is the sibling's; at level 1 it is the k-th listed neighbor
                        const int u = min(me.x + 1 + (k - me.y), width-1);
                        kn = valid ? (int) W.nbr[ls + W.perm[u]] - 1 : 0;
                    }
                }
                const float4 g1 = W.sc4[p];
                const float v1f = W.scv[p];
                p += ls;
                const float a2 = W.nba[kn], v2 = W.nbv[kn];
                const float dx = (W.nbx[kn]-pr.x)-g1.x, dy = (W.nby[kn]-pr.y)-g1.y, dz = (W.nbz[kn]-pr.z)-g1.z;
                const float d2 = dx*dx + dy*dy + dz*dz;
                const float df = __fdividef(g1.w*a2, g1.w+a2);
                const float u = df*0.318309886f;
                const float est = v1f*v2*(u*sqrtf(u))*__expf(-df*d2);
                const bool mb = valid && est > A.screen;
                const unsigned mm = __ballot_sync(FULL, mb);
                if (mb) {
                    const int q = nmaybe + __popc(mm & lanemask_lt());
                    if (q < wcap) W.cand[q] = p | (kn << 16);
                }
                nmaybe += __popc(mm);
            }
            if (nmaybe > wcap) { hw_w = max(hw_w, nmaybe); if (lane == 0) atomicOr(A.status, ST_LEVEL_OVERFLOW); failed = true; break; }
            __syncwarp();

            // ---- phase B: exact FP64 evaluation of the screened candidates; accepted ones become nodes ----
            int last_p = -1;                                              // parent of the newest node so far (warp-uniform)
            for (int k0 = 0; k0 < nmaybe; k0 += 32) {
                const int k = k0+lane;
                const bool valid = k < nmaybe;
                int p = 0, kn = 0;
                bool accept = false;
                double gvol = 0, df = 0, deltai = 0, s = 0, sp = 0, a1 = 0, v1 = 0, x1 = 0, y1 = 0, z1 = 0, a2 = 0, x2 = 0, y2 = 0, z2 = 0;
                int j = 0;
                double b1 = 0, w1 = 0, u1 = 0, q1 = 0, r1 = 0, b2 = 0, w2 = 0;
                float gam = 0.f;
                if (valid) {
                    const int pk = W.cand[k];
                    p = pk & 0xffff; kn = pk >> 16;
                    j = W.nbi[kn];
                    const NodeGauss* gp = G + ((level & 1)*wcap + p - ls);
                    a1 = gp->aL; v1 = gp->vL; x1 = gp->xL; y1 = gp->yL; z1 = gp->zL;
                    a2 = A.aL[j];
                    const double v2 = A.vL[j];
                    // the vdW-radius side is only needed for accepted candidates, but loading it here costs one memory
                    // round trip instead of two
                    b1 = gp->aS; w1 = gp->vS; u1 = gp->xS; q1 = gp->yS; r1 = gp->zS;
                    b2 = A.aS[j]; w2 = A.vS[j];
                    gam = gp->gam + A.gamma[j];                          // gaussvol.cpp:244
                    x2 = (double) W.nbx[kn] - (double) pr.x;              // exact in double
                    y2 = (double) W.nby[kn] - (double) pr.y;
                    z2 = (double) W.nbz[kn] - (double) pr.z;
                    const double dx = x2-x1, dy = y2-y1, dz = z2-z1;
                    const double d2 = dx*dx + dy*dy + dz*dz;
                    gvol = overlap_volume(a1, v1, a2, v2, d2, deltai, df);
                    pol_switch(gvol, A.volmina, A.volminb, A.swd, s, sp);
                    accept = (s*gvol > A.min_gvol);                          // gaussvol.cpp:233
                }
                const unsigned am = __ballot_sync(FULL, accept);
                {
                    // children ranges of the parents: candidates are enumerated parent-major, so the children of one parent
                    // are contiguous; the first child of a parent opens its range and closes the previous parent's
                    const unsigned before = am & lanemask_lt();
                    const int prev_p = __shfl_sync(FULL, p, before ? 31 - __clz(before) : 0);
                    const int pp = before ? prev_p : last_p;
                    const int slot = nslots + __popc(before);
                    if (accept && p != pp && slot-new_start < wcap) {
                        W.cstart[p-ls] = (short) slot;
                        if (pp >= 0) W.ccount[pp-ls] = (short) slot;         // ccount holds the END of the range
                        atomicOr(&W.hc[p >> 5], 1u << (p & 31));
                    }
                    if (am) last_p = __shfl_sync(FULL, p, 31 - __clz(am));
                }
                if (accept) {
                    const int slot = nslots + __popc(am & lanemask_lt());
                    if (slot < cap && slot-new_start < wcap) {
                        NodeGauss g;
                        // enlarged radii: topology + up-sweep data (gaussvol.cpp:234-245)
                        g.aL = a1+a2; g.vL = gvol;
                        g.xL = (x1*a1 + x2*a2)*deltai; g.yL = (y1*a1 + y2*a2)*deltai; g.zL = (z1*a1 + z2*a2)*deltai;
                        const double keyv = s*gvol;
                        W.key[slot-new_start] = keyv;
                        const double mL = 2.0*df*gvol;                        // -dVdr
                        const float vl = (float) keyv;
                        // dvv1 = V/V_parent feeds the sweeps only: a float division (the FP64 one is 30 instructions on the chain)
                        STAGE_ST(&swL[2*slot], make_float4(vl, (float) (sp*gvol + s), v1 > 0 ? (float) gvol/(float) v1 : 0.f, (float) a2/(float) g.aL));
                        STAGE_ST(&swL[2*slot+1], make_float4((float) ((x2-x1)*mL), (float) ((y2-y1)*mL), (float) ((z2-z1)*mL), gam));
                        // vdW radii on the same topology (rescan, gaussvol.cpp:261-279)
                        const double ex = x2-u1, ey = y2-q1, ez = z2-r1;
                        double dS, dfS, sS, spS;
                        const double gS = overlap_volume(b1, w1, b2, w2, ex*ex + ey*ey + ez*ez, dS, dfS);
                        pol_switch(gS, A.volmina, A.volminb, A.swd, sS, spS);
                        g.aS = b1+b2; g.vS = gS;
                        g.xS = (u1*b1 + x2*b2)*dS; g.yS = (q1*b1 + y2*b2)*dS; g.zS = (r1*b1 + z2*b2)*dS;
                        g.gam = gam; g.pad[0] = g.pad[1] = g.pad[2] = 0.f;
                        G[((level+1) & 1)*wcap + slot-new_start] = g;
                        W.sc4[slot-new_start] = make_float4((float) g.xL, (float) g.yL, (float) g.zL, (float) g.aL);
                        W.scv[slot-new_start] = (float) gvol;
                        const double mS = 2.0*dfS*gS;
                        const float vs = (float) (sS*gS);
                        STAGE_ST(&swS[2*slot], make_float4(vs, (float) (spS*gS + sS), w1 > 0 ? (float) gS/(float) w1 : 0.f, (float) b2/(float) g.aS));
                        STAGE_ST(&swS[2*slot+1], make_float4((float) (ex*mS), (float) (ey*mS), (float) (ez*mS), gam));
                        W.parent[slot] = (short) p; W.nbr[slot] = (short) (kn+1);
                        // energies and volumes need no tree accumulation (gaussvol.cpp:425-433 summed over the subtree)
                        if (level > 1 || nparts == 1) {                   // level-2 nodes of a split root: after the sort, when ownership is known
                            const float cg = coefp*gam;
                            eL_tot += (double) (cg*vl); eS_tot += (double) (cg*vs);
                            vsumL += (double) (cf*vl); vsumS += (double) (cf*vs);
                        }
                    }
                }
                nslots += __popc(am);
            }
            hw_slots = max(hw_slots, nslots); hw_w = max(hw_w, nslots-new_start);
            if (nslots > cap) { if (lane == 0) atomicOr(A.status, ST_NODE_OVERFLOW); failed = true; break; }
            if (nslots-new_start > wcap) { if (lane == 0) atomicOr(A.status, ST_LEVEL_OVERFLOW); failed = true; break; }
            if (nslots == new_start) break;
            if (lane == 0) W.ccount[last_p-ls] = (short) nslots;             // close the last parent's range
            __syncwarp();
            // siblings ordered by switched volume, larger first (gaussvol.cpp:97-100,171); rank sort within each group,
            // ties keep creation order
            for (int s0 = new_start; s0 < nslots; s0 += 32) {
                const int sl = s0+lane;
                if (sl < nslots) {
                    const int p = W.parent[sl] - ls;
                    const int cs = (int) W.cstart[p] - new_start, ce = (int) W.ccount[p] - new_start;
                    const int me = sl-new_start;
                    const double kv = W.key[me];
                    int rank = 0;
                    for (int y = cs; y < ce; y++) {
                        const double ky = W.key[y];
                        rank += (ky > kv) || (ky == kv && y < me);
                    }
                    W.perm[cs+rank] = (short) me;
                    W.gend[cs+rank] = (short) ce;
                    rk[sl] = (short) rank;
                    if (level == 1 && nparts > 1) {
                        // a level-2 node of a split root: its rank is its sorted position.  Owned: account its energy terms now;
                        // not owned: it stays as a sibling for the owned ones but contributes nothing (vol = sfp = 0)
                        if (rank % nparts == part) {
                            const float vl = swL[2*sl].x, vs = swS[2*sl].x, cg = coefp*swL[2*sl+1].w;
                            eL_tot += (double) (cg*vl); eS_tot += (double) (cg*vs);
                            vsumL += (double) (cf*vl); vsumS += (double) (cf*vs);
                            m_tot++;
                        } else {
                            float4 q = swL[2*sl]; q.x = 0.f; q.y = 0.f; swL[2*sl] = q;
                            q = swS[2*sl]; q.x = 0.f; q.y = 0.f; swS[2*sl] = q;
                        }
                    }
                }
            }
            __syncwarp();
            ls = new_start; le = nslots; level++;
            if (lane == 0) W.lvs[level] = new_start;
        }
        if (failed) continue;
        if (lane == 0) W.lvs[level+1] = nslots;
        __syncwarp();
        const int nlev = level;
        if (nparts == 1) m_tot += (lane == 0) ? (unsigned long long) (nslots-ng) : 0ull;
        else m_tot += (lane == 0 && nlev >= 2) ? (unsigned long long) (nslots-W.lvs[3]) : 0ull;   // + the owned level-2 nodes counted above

        // ---- bottom-up sweep, both radius sets (gaussvol.cpp:400-487); it also persists what the gamma sweep needs ----
        int off = 0;
        if (lane == 0) off = atomicAdd(A.st.cursor, nslots);
        off = __shfl_sync(FULL, off, 0);
        const bool fits = off+nslots <= A.st.cap;
        if (!fits) {
            if (lane == 0) atomicOr(A.status, ST_TREE_OVERFLOW);
        } else {
            if (lane == 0) { A.st.root_off[item] = off; A.st.root_cnt[item] = nslots; }
            if (lane <= nlev+1 && lane >= 1) A.st.root_lvs[item*MAX_LEVELS+lane] = (short) W.lvs[lane];
        }
        tree_sweep<false>(W, swL, swS, hu, nlev, 0, lane, A.accL, A.accS, nullptr,
                          fits ? A.st.rec + 2*(size_t) off : nullptr, fits ? A.st.rank + off : nullptr, rk, wcap);
        __syncwarp();
    }

#undef part
#undef nparts
    tail_end(3);
    eL_tot = warp_sum(eL_tot); eS_tot = warp_sum(eS_tot); vsumL = warp_sum(vsumL); vsumS = warp_sum(vsumS);
    m_tot = (unsigned long long) warp_sum((double) m_tot);
    if (lane == 0) {
        // raw sums of coef*gamma*vol; nu = +-gamma/roffset (ReferenceAGBNPKernels.cpp:297,364) is applied in k_finish
        atomicAdd(&A.scalars[SC_EVOL_L], eL_tot);
        atomicAdd(&A.scalars[SC_EVOL_S], eS_tot);
        atomicAdd(&A.scalars[SC_VOL_L], vsumL);
        atomicAdd(&A.scalars[SC_VOL_S], vsumS);
        atomicAdd(&A.counters[CT_C2], c2_tot);
        atomicAdd(&A.counters[CT_C3], c3_tot);
        atomicAdd(&A.counters[CT_M], m_tot);
        atomicMax(A.hw_nbr, hw_nn);
        atomicMax(A.hw_nodes, hw_slots);
        atomicMax(A.hw_width, hw_w);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_tree_rescan (opt-in, agbnp_b200_config::tree_reuse_interval > 1): the overlap tree of an EARLIER evaluation is kept
// -- same nodes, same parent / sibling structure -- and only re-evaluated at the new positions: every stored node's two
// Gaussians from its parent's and its last atom's (the arithmetic of k_tree's node creation, gaussvol.cpp:234-245 and
// :261-279, i.e. what the reference's rescan_tree_v does for the vdW radii, here for both radius sets), then the same
// bottom-up sweep.  No candidate search, no screen, no acceptance test, no sort.  An overlap that would newly pass the
// inclusion threshold is missing until the next build (its switched volume starts at 0, gaussvol.cpp:26-29), one that
// would be dropped stays with switched volume 0: between builds the energy is that of the frozen overlap set.  This is
// NOT the reference's semantics (it rebuilds every evaluation), hence opt-in; SURVEY 8f rank 3.
// ---------------------------------------------------------------------------------------------------------------
struct RescanArgs {
    const int2* items;
    int nitems;
    const float4* posq;
    const double *aL, *vL, *aS, *vS;
    const float* gamma;
    double volmina, volminb, swd;
    int cap;
    unsigned char* stage;             // k_tree's per-warp global staging
    size_t stage_stride;
    float4 *accL, *accS;
    double* scalars;
    unsigned long long* counters;
    TreeStore st;
    int* work_counter;
    int* status;
    const int* tree_ok;               // 1: the stored tree comes from a build evaluation that completed without overflow
};

__host__ __device__ inline size_t rescan_work_bytes(int cap) {
    return (((size_t) cap*(sizeof(int) + 2*sizeof(short)) + (MAX_LEVELS+2)*sizeof(int)) + 15) & ~(size_t) 15;
}

__global__ void __launch_bounds__(64, 8) k_tree_rescan(RescanArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_release();
    pdl_acquire();
    if (*A.tree_ok == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(A.status, ST_TREE_STALE);
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int cap = A.cap;
    const int gwarp = blockIdx.x*nwarp + warp;
    unsigned char* wk = smem_raw + (size_t) warp*rescan_work_bytes(cap);
    int* ja = (int*) wk;                            // [cap] sorted index of the node's last atom
    TreeWork W{};
    W.parent = (short*) (ja + cap);                 // [cap]
    W.ccount = W.parent + cap;                      // [cap] 1 if the node has children
    W.lvs = (int*) (W.ccount + cap);                // [MAX_LEVELS+2]
    unsigned char* stage = A.stage + (size_t) gwarp*A.stage_stride;
    NodeGauss* G = (NodeGauss*) stage;
    float4* swL = (float4*) (G+cap);
    float4* swS = swL + 2*(size_t) cap;
    float4* hu = swS + 2*(size_t) cap;

    double eL_tot = 0, eS_tot = 0, vsumL = 0, vsumS = 0;
    unsigned long long m_tot = 0;
    for (int item = first_unit(); item < A.nitems; item = next_unit(A.work_counter, lane)) {
        const int cnt = A.st.root_cnt[item];
        if (cnt <= 0) continue;                     // not this shard's item
        const int off = A.st.root_off[item];
        float4* rec = A.st.rec + 2*(size_t) off;
        const short* rank = A.st.rank + off;
        const short* slv = A.st.root_lvs + item*MAX_LEVELS;
        int nlev = 1;
        while (nlev+1 < MAX_LEVELS && slv[nlev+1] < cnt) nlev++;
        const int2 itm = A.items[item];
        const int ng = item_nroots(itm), part = item_part(itm), nparts = item_parts(itm);
        __syncwarp();                               // the previous item is done with the work arrays
        if (lane >= 1 && lane <= nlev) W.lvs[lane] = slv[lane];
        if (lane == 0) W.lvs[nlev+1] = cnt;
        for (int sl = lane; sl < cnt; sl += 32) {
            const int pk = __float_as_int(rec[2*sl+1].w);
            ja[sl] = __float_as_int(rec[2*sl].w);
            W.parent[sl] = (short) (pk & 0xffff);   // 0xffff (the root) becomes -1
            W.ccount[sl] = (short) ((pk >> 16) & 1);
        }
        __syncwarp();
        // slots 0 .. G-1: the root atoms (gaussvol.cpp:130-148); positions are relative to the first root, as in k_tree
        const float4 pr = A.posq[ja[0]];
        if (lane < ng) {
            const int rg = ja[lane];
            const float4 pg = A.posq[rg];
            const float gam_r = A.gamma[rg];
            NodeGauss g;
            g.aL = A.aL[rg]; g.vL = A.vL[rg];
            g.xL = (double) pg.x - (double) pr.x; g.yL = (double) pg.y - (double) pr.y; g.zL = (double) pg.z - (double) pr.z;
            g.aS = A.aS[rg]; g.vS = A.vS[rg]; g.xS = g.xL; g.yS = g.yL; g.zS = g.zL;
            g.gam = gam_r; g.pad[0] = g.pad[1] = g.pad[2] = 0.f;
            G[lane] = g;
            const float own = part == 0 ? 1.f : 0.f;
            swL[2*lane] = make_float4(own*(float) g.vL, 1.f, 1.f, 1.f); swL[2*lane+1] = make_float4(0.f, 0.f, 0.f, gam_r);
            swS[2*lane] = make_float4(own*(float) g.vS, 1.f, 1.f, 1.f); swS[2*lane+1] = make_float4(0.f, 0.f, 0.f, gam_r);
            eL_tot += (double) (own*gam_r*(float) g.vL); eS_tot += (double) (own*gam_r*(float) g.vS);
            vsumL += (double) (own*(float) g.vL); vsumS += (double) (own*(float) g.vS);
        }
        __syncwarp();
        // top-down: every stored node from its parent and its last atom.  (Measured without gain: two chunks of a level per
        // trip -- spills at the 128-register cap, 124 vs 114 us; narrow levels kept in registers with the parent fetched by
        // shuffle instead of through the staged G -- 113 vs 114 us.)
        for (int lev = 2; lev <= nlev; lev++) {
            const int b = W.lvs[lev], e = W.lvs[lev+1];
            const float cf = (lev & 1) ? 1.f : -1.f;
            const float coefp = cf/(float) lev;
            for (int sl = b+lane; sl < e; sl += 32) {
                const int p = W.parent[sl], j = ja[sl];
                const NodeGauss* gp = G+p;
                const double a1 = gp->aL, v1 = gp->vL, x1 = gp->xL, y1 = gp->yL, z1 = gp->zL;
                const double b1 = gp->aS, w1 = gp->vS, u1 = gp->xS, q1 = gp->yS, r1 = gp->zS;
                const double a2 = A.aL[j], v2 = A.vL[j], b2 = A.aS[j], w2 = A.vS[j];
                const float gam = gp->gam + A.gamma[j];                          // gaussvol.cpp:244
                const float4 pj = A.posq[j];
                const double x2 = (double) pj.x - (double) pr.x, y2 = (double) pj.y - (double) pr.y, z2 = (double) pj.z - (double) pr.z;
                const double dx = x2-x1, dy = y2-y1, dz = z2-z1;
                double deltai, df, s, sp;
                const double gvol = overlap_volume(a1, v1, a2, v2, dx*dx + dy*dy + dz*dz, deltai, df);
                pol_switch(gvol, A.volmina, A.volminb, A.swd, s, sp);
                NodeGauss g;
                g.aL = a1+a2; g.vL = gvol;
                g.xL = (x1*a1 + x2*a2)*deltai; g.yL = (y1*a1 + y2*a2)*deltai; g.zL = (z1*a1 + z2*a2)*deltai;
                const double mL = 2.0*df*gvol;
                float vl = (float) (s*gvol), sfl = (float) (sp*gvol + s);
                const double ex = x2-u1, ey = y2-q1, ez = z2-r1;
                double dS, dfS, sS, spS;
                const double gS = overlap_volume(b1, w1, b2, w2, ex*ex + ey*ey + ez*ez, dS, dfS);
                pol_switch(gS, A.volmina, A.volminb, A.swd, sS, spS);
                g.aS = b1+b2; g.vS = gS;
                g.xS = (u1*b1 + x2*b2)*dS; g.yS = (q1*b1 + y2*b2)*dS; g.zS = (r1*b1 + z2*b2)*dS;
                g.gam = gam; g.pad[0] = g.pad[1] = g.pad[2] = 0.f;
                G[sl] = g;
                const double mS = 2.0*dfS*gS;
                float vs = (float) (sS*gS), sfs = (float) (spS*gS + sS);
                // a level-2 node of a split root that this part does not own stays as a sibling: no terms of its own
                const bool owned = lev > 2 || nparts == 1 || ((int) rank[sl] % nparts) == part;
                if (!owned) { vl = 0.f; sfl = 0.f; vs = 0.f; sfs = 0.f; }
                swL[2*sl] = make_float4(vl, sfl, (float) (v1 > 0 ? gvol/v1 : 0.0), (float) a2/(float) g.aL);
                swL[2*sl+1] = make_float4((float) ((x2-x1)*mL), (float) ((y2-y1)*mL), (float) ((z2-z1)*mL), gam);
                const float4 qs0 = make_float4(vs, sfs, (float) (w1 > 0 ? gS/w1 : 0.0), (float) b2/(float) g.aS);
                const float4 qs1 = make_float4((float) (ex*mS), (float) (ey*mS), (float) (ez*mS), gam);
                swS[2*sl] = qs0; swS[2*sl+1] = qs1;
                // what the gamma sweep reads, refreshed in place (the topology words stay; the root's record is constant)
                rec[2*sl] = make_float4(coefp*qs0.y, qs0.z, qs0.w, __int_as_float(j));
                rec[2*sl+1] = make_float4(qs1.x, qs1.y, qs1.z, __int_as_float(((int) p & 0xffff) | ((int) W.ccount[sl] << 16)));
                if (owned) {
                    const float cg = coefp*gam;
                    eL_tot += (double) (cg*vl); eS_tot += (double) (cg*vs);
                    vsumL += (double) (cf*vl); vsumS += (double) (cf*vs);
                    m_tot++;
                }
            }
            __syncwarp();
        }
        if (lane == 0) atomicAdd(A.st.cursor, cnt);  // the control word reports the size of the tree, as after a build
        tree_sweep<true>(W, swL, swS, hu, nlev, 0, lane, A.accL, A.accS, ja);
    }
    eL_tot = warp_sum(eL_tot); eS_tot = warp_sum(eS_tot); vsumL = warp_sum(vsumL); vsumS = warp_sum(vsumS);
    m_tot = (unsigned long long) warp_sum((double) m_tot);
    if (lane == 0) {
        atomicAdd(&A.scalars[SC_EVOL_L], eL_tot);
        atomicAdd(&A.scalars[SC_EVOL_S], eS_tot);
        atomicAdd(&A.scalars[SC_VOL_L], vsumL);
        atomicAdd(&A.scalars[SC_VOL_S], vsumS);
        atomicAdd(&A.counters[CT_M], m_tot);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_tree_gamma: S10+S11 merged -- deposit nu_i = (W_i+U_i)/V_i on the stored tree (rescan_tree_g, gaussvol.cpp:330-372),
// run the energy-gradient up-sweep (gaussvol.cpp:400-487) and add force = -gradient
// (ReferenceAGBNPKernels.cpp:718-747; the merge of the W and U passes is exact because the sweep is linear in nu).
// Multi-GPU: each shard stores and sweeps only the subtrees of the roots it owns; the partial forces are all-reduced.
// ---------------------------------------------------------------------------------------------------------------
struct GammaArgs {
    int nitems, np;
    TreeStore st;
    const float4* dacc;         // [np] .w = W_i + U_i
    const float* inv_vS;        // [np] 1/V_i (vdW radii), 0 for hydrogens / padding
    float4* gacc;               // [np] out: force x,y,z (w unused)
    unsigned char* scratch;     // per-warp work arrays in global memory (oversize capacity), or nullptr = shared memory
    size_t scratch_stride;
    int cap;
    int* work_counter;
    int shard_rank, shard_count;    // items are dealt to shards as in k_tree (blocks of 32 in longest-first order)
};

__host__ __device__ inline size_t gamma_work_bytes(int cap) { return ((size_t) cap*(sizeof(float) + sizeof(float4) + sizeof(short)) + 15) & ~(size_t) 15; }

template <bool SMEM_WORK>
__global__ void __launch_bounds__(128) k_tree_gamma(GammaArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_release();
    pdl_acquire();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    unsigned char* wk;
    if (SMEM_WORK) wk = smem_raw + (size_t) warp*gamma_work_bytes(A.cap);
    else wk = A.scratch + (size_t) (blockIdx.x*nwarp+warp)*A.scratch_stride;
    float4* hu = (float4*) wk;                      // [cap] children sums (F', P'x, P'y, P'z) per parent slot
    float* gam = (float*) (hu + A.cap);             // [cap] gamma_1..n per slot
    short* par = (short*) (gam + A.cap);            // [cap] parent slot
    // r: item index (stored subtrees are per item; not-owned nodes carry zeros)
    const int raw_end = A.shard_count > 1 ? (A.nitems+TILE-1)/TILE*TILE : A.nitems;
    tail_begin(4);
    for (int raw = claim_unit(A.work_counter, lane); ; raw = claim_unit(A.work_counter, lane)) {
        int r = raw;
        if (A.shard_count > 1) {                    // this shard's raw-th item (the same deal as k_tree)
            r = ((raw >> 5)*A.shard_count + A.shard_rank)*TILE + (raw & 31);
            if (r >= raw_end) break;
            if (r >= A.nitems) continue;
        } else if (r >= A.nitems) break;
        const int cnt = A.st.root_cnt[r];
        if (cnt <= 1) continue;             // an atom without overlaps: dv1 = 0, no force (gaussvol.cpp:472)
        const float4* rec = A.st.rec + 2*(size_t) A.st.root_off[r];
        const short* lvs = A.st.root_lvs + r*MAX_LEVELS;
        int nlev = 1;
        while (nlev+1 < MAX_LEVELS && lvs[nlev+1] < cnt) nlev++;
        // nu of every slot's last atom first: the gathers rec -> atom -> (W+U, 1/V) of all slots are independent of each
        // other, so four of them per lane are in flight at once instead of one dependent chain per level
        for (int s0 = 0; s0 < cnt; s0 += 128) {
            int ja[4], pk[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int sl = min(s0 + 32*q + lane, cnt-1);
                ja[q] = __float_as_int(__ldg(&rec[2*sl].w));
                pk[q] = __float_as_int(__ldg(&rec[2*sl+1].w));
            }
            float nu[4];
#pragma unroll
            for (int q = 0; q < 4; q++) nu[q] = __ldg(&A.dacc[ja[q]].w)*__ldg(&A.inv_vS[ja[q]]);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int sl = s0 + 32*q + lane;
                if (sl < cnt) { gam[sl] = nu[q]; par[sl] = (short) (pk[q] & 0xffff); }
            }
        }
        __syncwarp();
        // top-down gamma1i (gaussvol.cpp:338-343): shared memory only
        for (int lev = 2; lev <= nlev; lev++) {
            const int b = lvs[lev], e = lev == nlev ? cnt : lvs[lev+1];
            for (int sl = b+lane; sl < e; sl += 32) gam[sl] += gam[par[sl]];
            __syncwarp();
        }
        // bottom-up energy-gradient sweep; children sums by segmented warp scans (see tree_sweep)
        for (int lev = nlev; lev >= 1; lev--) {
            const int b = lvs[lev], e = lev == nlev ? cnt : lvs[lev+1];
            int carry_key = -2;
            float cF = 0.f, cx = 0.f, cy = 0.f, cz = 0.f;
            for (int s0 = b; s0 < e; s0 += 32) {
                const int sl = s0+lane;
                const bool valid = sl < e;
                int key = -3-lane;
                float vF = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
                if (valid) {
                    const float4 q0 = rec[2*sl], q1 = rec[2*sl+1];
                    const int pk = __float_as_int(q1.w);
                    float f = q0.x*gam[sl], px = 0.f, py = 0.f, pz = 0.f;
                    if (pk & 0x10000) { const float4 h = hu[sl]; f += h.x; px = h.y; py = h.z; pz = h.w; }
                    const float c2a = q0.z, c2b = 1.f-q0.z;
                    const float gx = fmaf(px, c2a, -q1.x*f), gy = fmaf(py, c2a, -q1.y*f), gz = fmaf(pz, c2a, -q1.z*f);
                    if (gx != 0.f || gy != 0.f || gz != 0.f) atomicAdd(&A.gacc[__float_as_int(q0.w)], make_float4(-gx, -gy, -gz, 0.f));
                    key = pk & 0xffff;
                    vF = q0.y*f; vx = fmaf(px, c2b, q1.x*f); vy = fmaf(py, c2b, q1.y*f); vz = fmaf(pz, c2b, q1.z*f);
                }
                if (lev == 1) break;
                const int key0 = __shfl_sync(FULL, key, 0);
                if (carry_key >= 0) {
                    if (key0 == carry_key) { if (lane == 0) { vF += cF; vx += cx; vy += cy; vz += cz; } }
                    else if (lane == 0) hu[carry_key] = make_float4(cF, cx, cy, cz);
                }
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int ku = __shfl_up_sync(FULL, key, d);
                    const bool take = lane >= d && ku == key;
                    const float tF = __shfl_up_sync(FULL, vF, d), tx = __shfl_up_sync(FULL, vx, d);
                    const float ty = __shfl_up_sync(FULL, vy, d), tz = __shfl_up_sync(FULL, vz, d);
                    if (take) { vF += tF; vx += tx; vy += ty; vz += tz; }
                }
                const int kn = __shfl_down_sync(FULL, key, 1);
                const bool tail = valid && (lane == 31 || kn != key);
                const int last = min(31, e-1-s0);
                if (tail && lane != last) hu[key] = make_float4(vF, vx, vy, vz);
                carry_key = __shfl_sync(FULL, key, last);
                cF = __shfl_sync(FULL, vF, last); cx = __shfl_sync(FULL, vx, last);
                cy = __shfl_sync(FULL, vy, last); cz = __shfl_sync(FULL, vz, last);
            }
            if (lev > 1 && carry_key >= 0 && lane == 0) hu[carry_key] = make_float4(cF, cx, cy, cz);
            __syncwarp();
        }
    }
    tail_end(4);
}

} // namespace agbnp_b200_impl
#endif
