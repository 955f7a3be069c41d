// GaussVol overlap-tree kernels for sm_100a.
//
// Semantics follow gaussvol/gaussvol.cpp of the reference (citations inline); the decomposition does not:
//   * the reference builds one global tree by depth-first recursion; here every heavy atom's subtree (all overlaps whose
//     lowest-index atom is that atom) is owned by ONE WARP, which builds it breadth-first, level by level
//     (candidate enumeration by warp prefix sums, acceptance compaction by ballots), so subtrees never communicate;
//   * the large-radius build (S1), the vdW-radius rescan (S3) and both up-sweeps (S2, second half of S3) are fused:
//     a node's vdW-radius Gaussian is computed when the node is created, and one bottom-up pass yields both energies,
//     both sets of self-volumes and the combined surface-tension force;
//   * topology-deciding arithmetic (overlap volume, inclusion threshold, sibling sort key) is FP64 with the expression
//     structure of gaussvol.cpp:60-93; everything fed to energies/forces downstream is rounded to FP32.
// Only what the later "gamma" sweep (S10+S11 merged, linear in nu) needs is persisted to HBM (TreeStore).
#ifndef AGBNP_TREE_CUH_
#define AGBNP_TREE_CUH_

#include "agbnp_device.cuh"

namespace agbnp_b200_impl {

constexpr int TREE_THREADS = 256;
constexpr int TREE_WARPS = TREE_THREADS/32;
constexpr int MAX_LEVELS = 10;      // level index 1..8 used (MAX_ORDER 8)

// persisted per-node records (SoA) for the gamma sweep and for the topology dump
struct TreeStore {
    int cap;                 // node capacity
    int* cursor;             // bump allocator
    int* root_off;           // [nh] first node of the root's subtree (slot 0 = the root atom itself)
    int* root_cnt;           // [nh] nodes in the subtree including slot 0; 0 if not built
    short* root_lvs;         // [nh*MAX_LEVELS] first slot of each level, root_lvs[r*MAX_LEVELS+l], l = 1..nlev+1
    float *cs, *dvv, *dx, *dy, *dz, *c2a, *c2b;   // coefp*sfp, dvv1, dv1[3], a_i/a_1i, a_1/a_1i   (vdW radii)
    int* atom;               // sorted index of the node's last atom
    short *parent, *cstart, *ccount, *rank;       // slots relative to the subtree start; rank among siblings
};

struct TreeArgs {
    int nh, nhb, np;
    const float4* posq;
    const int* orig;
    const unsigned char* rcbin;
    const double *aL, *vL, *aS, *vS;
    const float* gamma;
    const float4 *bbc, *bbh;
    const float* rc2;
    const float* rc2max;
    int nbins;
    double volmina, volminb, min_gvol, swd;
    float inv_roffset;
    int max_order;
    unsigned char* scratch;
    size_t scratch_stride;
    int cap, nbrmax;
    double *svS, *svL;
    unsigned long long* force;        // [3][np] fixed point
    double* scalars;
    unsigned long long* counters;
    TreeStore st;
    int* work_counter;
    int* status;
    int shard_rank, shard_count;      // roots are dealt to shards block-cyclically (blocks of 32 sorted heavy atoms)
    int *hw_nbr, *hw_nodes;           // high-water marks: level-2 neighbors / nodes of one root
};

__host__ __device__ inline size_t tree_scratch_bytes(int cap) {
    size_t b = (size_t) cap*(11*sizeof(double) + 24*sizeof(float) + sizeof(int) + 6*sizeof(short) + 1) + 2*sizeof(int);
    return (b + 255) & ~(size_t) 255;
}
__host__ __device__ inline size_t tree_smem_per_warp(int nbrmax) {
    size_t b = (size_t) nbrmax*3*sizeof(double) + (size_t) 5*(nbrmax+1)*sizeof(float) + (size_t) nbrmax*sizeof(int)
             + (size_t) (MAX_LEVELS+2)*sizeof(int);
    return (b + 15) & ~(size_t) 15;
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

struct TreeScratch {
    double *gLa, *gLv, *gLx, *gLy, *gLz, *gSa, *gSv, *gSx, *gSy, *gSz, *key;
    float *sfpL, *dvvL, *dLx, *dLy, *dLz, *volS, *sfpS, *dvvS, *dSx, *dSy, *dSz, *gam;
    float *aEL, *afL, *apLx, *apLy, *apLz, *apsL, *aES, *afS, *apSx, *apSy, *apSz, *apsS;
    int* pref;               // [cap+2] exclusive prefix of candidate counts of the level being expanded
    short *parent, *nbr, *cstart, *ccount, *perm, *gend;
    unsigned char* lvl;
    __device__ void bind(unsigned char* base, int cap) {
        double* d = (double*) base;
        gLa = d; gLv = d+cap; gLx = d+2*cap; gLy = d+3*cap; gLz = d+4*cap;
        gSa = d+5*cap; gSv = d+6*cap; gSx = d+7*cap; gSy = d+8*cap; gSz = d+9*cap; key = d+10*cap;
        float* f = (float*) (d+11*(size_t) cap);
        sfpL = f; dvvL = f+cap; dLx = f+2*cap; dLy = f+3*cap; dLz = f+4*cap; volS = f+5*cap; sfpS = f+6*cap;
        dvvS = f+7*cap; dSx = f+8*cap; dSy = f+9*cap; dSz = f+10*cap; gam = f+11*cap;
        float* a = f+12*(size_t) cap;
        aEL = a; afL = a+cap; apLx = a+2*cap; apLy = a+3*cap; apLz = a+4*cap; apsL = a+5*cap;
        aES = a+6*cap; afS = a+7*cap; apSx = a+8*cap; apSy = a+9*cap; apSz = a+10*cap; apsS = a+11*cap;
        pref = (int*) (a+12*(size_t) cap);
        short* s = (short*) (pref + (cap+2 - (cap & 1)));      // keeps 4-byte alignment irrelevant for shorts; even count
        parent = s; nbr = s+cap; cstart = s+2*cap; ccount = s+3*cap; perm = s+4*cap; gend = s+5*cap;
        lvl = (unsigned char*) (s+6*(size_t) cap);
    }
};

// ---------------------------------------------------------------------------------------------------------------
// k_tree: build + vdW rescan + fused up-sweep for every heavy root atom (reference S1-S3:
// gaussvol.cpp:103-250,254-327,389-519,589-606; ReferenceAGBNPKernels.cpp:290-380)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TREE_THREADS, 2) k_tree(TreeArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nbrmax = A.nbrmax, cap = A.cap;

    unsigned char* sm = smem_raw + (size_t) warp*tree_smem_per_warp(nbrmax);
    double* nb_x = (double*) sm;
    double* nb_y = nb_x+nbrmax;
    double* nb_z = nb_y+nbrmax;
    float* acc = (float*) (nb_z+nbrmax);            // [5][nbrmax+1]: svS, svL, gx, gy, gz; index 0 = root, 1+k = neighbor k
    int* nb_idx = (int*) (acc+5*(nbrmax+1));
    int* lvs = nb_idx+nbrmax;                       // [MAX_LEVELS+2]
    const int accs = nbrmax+1;

    TreeScratch S;
    S.bind(A.scratch + (size_t) (blockIdx.x*TREE_WARPS+warp)*A.scratch_stride, cap);

    double eL_tot = 0, eS_tot = 0, vsumL = 0, vsumS = 0;     // per-lane partial sums, reduced at the end
    unsigned long long c2_tot = 0, c3_tot = 0, m_tot = 0;
    int hw_nn = 0, hw_slots = 0;

    for (;;) {
        int r = 0;
        if (lane == 0) r = atomicAdd(A.work_counter, 1);
        r = __shfl_sync(FULL, r, 0);
        if (A.shard_count > 1) {
            r = ((r >> 5)*A.shard_count + A.shard_rank)*TILE + (r & 31);
            if (r >= A.nhb*TILE) break;
            if (r >= A.nh) continue;
        } else if (r >= A.nh) break;

        const float4 pr = A.posq[r];
        const int orig_r = A.orig[r];
        const int rb = A.rcbin[r];
        const float rcmax = A.rc2max[rb];
        for (int i = lane; i < 5*accs; i += 32) acc[i] = 0.f;

        // ---- level-2 candidate list: heavy atoms later in the caller's order within the conservative pair radius ----
        int nn = 0;
        for (int b0 = 0; b0 < A.nhb; b0 += 32) {
            const int b = b0+lane;
            bool hit = false;
            if (b < A.nhb) hit = point_box_dist2(pr.x, pr.y, pr.z, A.bbc[b], A.bbh[b]) < rcmax;
            unsigned m = __ballot_sync(FULL, hit);
            while (m) {
                const int bb = b0+__ffs(m)-1;
                m &= m-1;
                const int j = bb*TILE+lane;
                const float4 pj = A.posq[j];
                const int oj = A.orig[j];
                const float dx = pj.x-pr.x, dy = pj.y-pr.y, dz = pj.z-pr.z;
                const float d2 = dx*dx + dy*dy + dz*dz;
                const bool ok = (oj > orig_r) && (d2 < A.rc2[rb*A.nbins + A.rcbin[j]]);
                const unsigned am = __ballot_sync(FULL, ok);
                if (ok) {
                    const int p = nn + __popc(am & lanemask_lt());
                    if (p < nbrmax) {
                        nb_idx[p] = j;
                        nb_x[p] = (double) pj.x - (double) pr.x;      // exact in double
                        nb_y[p] = (double) pj.y - (double) pr.y;
                        nb_z[p] = (double) pj.z - (double) pr.z;
                    }
                }
                nn += __popc(am);
            }
        }
        hw_nn = max(hw_nn, nn);
        if (nn > nbrmax) {
            if (lane == 0) atomicOr(A.status, ST_NBR_OVERFLOW);
            continue;
        }

        // ---- slot 0: the root atom (gaussvol.cpp:130-148) ----
        if (lane == 0) {
            S.gLa[0] = A.aL[r]; S.gLv[0] = A.vL[r]; S.gLx[0] = S.gLy[0] = S.gLz[0] = 0.0;
            S.gSa[0] = A.aS[r]; S.gSv[0] = A.vS[r]; S.gSx[0] = S.gSy[0] = S.gSz[0] = 0.0;
            S.key[0] = A.vL[r];
            S.sfpL[0] = 1.f; S.dvvL[0] = 1.f; S.dLx[0] = S.dLy[0] = S.dLz[0] = 0.f;
            S.volS[0] = (float) A.vS[r]; S.sfpS[0] = 1.f; S.dvvS[0] = 1.f; S.dSx[0] = S.dSy[0] = S.dSz[0] = 0.f;
            S.gam[0] = A.gamma[r];
            S.parent[0] = -1; S.nbr[0] = 0; S.cstart[0] = 0; S.ccount[0] = 0; S.perm[0] = 0; S.gend[0] = 1;
            S.lvl[0] = 1;
            lvs[1] = 0;
        }
        __syncwarp();

        // ---- breadth-first build: level -> level+1 ----
        int nslots = 1, ls = 0, le = 1, level = 1;
        bool failed = false;
        while (level < A.max_order) {           // a node at level >= MAX_ORDER gets no children (gaussvol.cpp:211)
            int T, width = le-ls;
            if (level == 1) {
                T = nn;
            } else {
                // candidates of node at sorted position t: its younger siblings t+1 .. gend[t]-1 (gaussvol.cpp:221)
                int carry = 0;
                for (int t0 = 0; t0 < width; t0 += 32) {
                    const int t = t0+lane;
                    int c = 0;
                    if (t < width) c = (int) S.gend[ls+t] - (ls+t) - 1;
                    const int inc = warp_incl_scan(c);
                    if (t < width) S.pref[t] = carry + inc - c;
                    carry += __shfl_sync(FULL, inc, 31);
                }
                if (lane == 0) S.pref[width] = carry;
                T = carry;
                __syncwarp();
            }
            if (level == 1) c2_tot += (lane == 0) ? (unsigned long long) T : 0ull;
            else c3_tot += (lane == 0) ? (unsigned long long) T : 0ull;

            const int new_start = nslots;
            for (int k0 = 0; k0 < T; k0 += 32) {
                const int k = k0+lane;
                const bool valid = k < T;
                int p = 0, kn = k;
                if (valid && level > 1) {
                    int lo = 0, hi = width-1;
                    while (lo < hi) {
                        const int mid = (lo+hi+1) >> 1;
                        if (S.pref[mid] <= k) lo = mid; else hi = mid-1;
                    }
                    const int tpos = ls+lo;
                    const int u = tpos + 1 + (k - S.pref[lo]);
                    p = S.perm[tpos];
                    kn = (int) S.nbr[S.perm[u]] - 1;
                }
                bool accept = false;
                double gvol = 0, df = 0, deltai = 0, s = 0, sp = 0, a1 = 0, v1 = 0, x1 = 0, y1 = 0, z1 = 0, a2 = 0, x2 = 0, y2 = 0, z2 = 0;
                int j = 0;
                if (valid) {
                    j = nb_idx[kn];
                    a1 = S.gLa[p]; v1 = S.gLv[p]; x1 = S.gLx[p]; y1 = S.gLy[p]; z1 = S.gLz[p];
                    a2 = A.aL[j];
                    const double v2 = A.vL[j];
                    x2 = nb_x[kn]; y2 = nb_y[kn]; z2 = nb_z[kn];
                    const double dx = x2-x1, dy = y2-y1, dz = z2-z1;
                    const double d2 = dx*dx + dy*dy + dz*dz;
                    gvol = overlap_volume(a1, v1, a2, v2, d2, deltai, df);
                    pol_switch(gvol, A.volmina, A.volminb, A.swd, s, sp);
                    accept = (s*gvol > A.min_gvol);                          // gaussvol.cpp:233
                }
                const unsigned am = __ballot_sync(FULL, accept);
                if (accept) {
                    const int slot = nslots + __popc(am & lanemask_lt());
                    if (slot < cap) {
                        // enlarged radii: topology + up-sweep data (gaussvol.cpp:234-245)
                        S.gLa[slot] = a1+a2; S.gLv[slot] = gvol;
                        S.gLx[slot] = (x1*a1 + x2*a2)*deltai;
                        S.gLy[slot] = (y1*a1 + y2*a2)*deltai;
                        S.gLz[slot] = (z1*a1 + z2*a2)*deltai;
                        S.key[slot] = s*gvol;
                        S.sfpL[slot] = (float) (sp*gvol + s);
                        S.dvvL[slot] = (float) (v1 > 0 ? gvol/v1 : 0.0);
                        const double mL = 2.0*df*gvol;                        // -dVdr
                        S.dLx[slot] = (float) ((x2-x1)*mL); S.dLy[slot] = (float) ((y2-y1)*mL); S.dLz[slot] = (float) ((z2-z1)*mL);
                        // vdW radii on the same topology (rescan, gaussvol.cpp:261-279)
                        const double b1 = S.gSa[p], w1 = S.gSv[p], u1 = S.gSx[p], q1 = S.gSy[p], r1 = S.gSz[p];
                        const double b2 = A.aS[j], w2 = A.vS[j];
                        const double ex = x2-u1, ey = y2-q1, ez = z2-r1;
                        double dS, dfS, sS, spS;
                        const double gS = overlap_volume(b1, w1, b2, w2, ex*ex + ey*ey + ez*ez, dS, dfS);
                        pol_switch(gS, A.volmina, A.volminb, A.swd, sS, spS);
                        S.gSa[slot] = b1+b2; S.gSv[slot] = gS;
                        S.gSx[slot] = (u1*b1 + x2*b2)*dS;
                        S.gSy[slot] = (q1*b1 + y2*b2)*dS;
                        S.gSz[slot] = (r1*b1 + z2*b2)*dS;
                        S.volS[slot] = (float) (sS*gS);
                        S.sfpS[slot] = (float) (spS*gS + sS);
                        S.dvvS[slot] = (float) (w1 > 0 ? gS/w1 : 0.0);
                        const double mS = 2.0*dfS*gS;
                        S.dSx[slot] = (float) (ex*mS); S.dSy[slot] = (float) (ey*mS); S.dSz[slot] = (float) (ez*mS);
                        S.gam[slot] = S.gam[p] + A.gamma[j];                  // gaussvol.cpp:244
                        S.parent[slot] = (short) p; S.nbr[slot] = (short) (kn+1);
                        S.cstart[slot] = 0; S.ccount[slot] = 0;
                        S.lvl[slot] = (unsigned char) (level+1);
                    }
                }
                nslots += __popc(am);
            }
            if (nslots > cap) { if (lane == 0) atomicOr(A.status, ST_NODE_OVERFLOW); failed = true; break; }
            __syncwarp();
            if (nslots == new_start) break;

            // children ranges of the parents (children of one parent are contiguous: candidates are enumerated parent-major)
            for (int s0 = new_start; s0 < nslots; s0 += 32) {
                const int sl = s0+lane;
                if (sl < nslots && (sl == new_start || S.parent[sl] != S.parent[sl-1])) S.cstart[S.parent[sl]] = (short) sl;
            }
            __syncwarp();
            for (int s0 = new_start; s0 < nslots; s0 += 32) {
                const int sl = s0+lane;
                if (sl < nslots && (sl == nslots-1 || S.parent[sl+1] != S.parent[sl])) {
                    const int p = S.parent[sl];
                    S.ccount[p] = (short) (sl+1 - S.cstart[p]);
                }
            }
            __syncwarp();
            // siblings ordered by switched volume, larger first (gaussvol.cpp:97-100,171); rank sort within each group,
            // ties keep creation order
            for (int s0 = new_start; s0 < nslots; s0 += 32) {
                const int sl = s0+lane;
                if (sl < nslots) {
                    const int p = S.parent[sl];
                    const int cs = S.cstart[p], ce = cs + S.ccount[p];
                    const double kv = S.key[sl];
                    int rank = 0;
                    for (int y = cs; y < ce; y++) {
                        const double ky = S.key[y];
                        rank += (ky > kv) || (ky == kv && y < sl);
                    }
                    S.perm[cs+rank] = (short) sl;
                    S.gend[cs+rank] = (short) ce;
                }
            }
            __syncwarp();
            ls = new_start; le = nslots; level++;
            if (lane == 0) lvs[level] = new_start;
        }
        if (failed) continue;
        if (lane == 0) lvs[level+1] = nslots;
        __syncwarp();
        const int nlev = level;
        hw_slots = max(hw_slots, nslots);
        m_tot += (lane == 0) ? (unsigned long long) (nslots-1) : 0ull;

        // ---- fused bottom-up sweep for both radius sets (gaussvol.cpp:400-487) ----
        float eL_root = 0.f, eS_root = 0.f;
        for (int lev = nlev; lev >= 1; lev--) {
            const int b = lvs[lev], e = lvs[lev+1];
            const float cf = (lev & 1) ? 1.f : -1.f;
            const float coefp = cf/(float) lev;
            for (int s0 = b; s0 < e; s0 += 32) {
                const int sl = s0+lane;
                if (sl < e) {
                    const float g = S.gam[sl];
                    const float vl = (float) S.key[sl], vs = S.volS[sl];
                    float EL = coefp*g*vl, fL = coefp*S.sfpL[sl]*g, pLx = 0.f, pLy = 0.f, pLz = 0.f, psL = coefp*vl;
                    float ES = coefp*g*vs, fS = coefp*S.sfpS[sl]*g, pSx = 0.f, pSy = 0.f, pSz = 0.f, psS = coefp*vs;
                    const int cs = S.cstart[sl], ce = cs + S.ccount[sl];
                    for (int c = cs; c < ce; c++) {
                        EL += S.aEL[c]; fL += S.afL[c]; pLx += S.apLx[c]; pLy += S.apLy[c]; pLz += S.apLz[c]; psL += S.apsL[c];
                        ES += S.aES[c]; fS += S.afS[c]; pSx += S.apSx[c]; pSy += S.apSy[c]; pSz += S.apSz[c]; psS += S.apsS[c];
                    }
                    const int ia = S.nbr[sl];
                    const int ja = ia == 0 ? r : nb_idx[ia-1];
                    const float a1iL = (float) S.gLa[sl], a1iS = (float) S.gSa[sl];
                    const float aiL = (float) A.aL[ja], aiS = (float) A.aS[ja];
                    const float c2aL = aiL/a1iL, c2bL = (a1iL-aiL)/a1iL;
                    const float c2aS = aiS/a1iS, c2bS = (a1iS-aiS)/a1iS;
                    const float dlx = S.dLx[sl], dly = S.dLy[sl], dlz = S.dLz[sl];
                    const float dsx = S.dSx[sl], dsy = S.dSy[sl], dsz = S.dSz[sl];
                    // gradient of (E_L - E_S) w.r.t. the node's last atom (gaussvol.cpp:467-474)
                    atomicAdd(&acc[0*accs+ia], psS);
                    atomicAdd(&acc[1*accs+ia], psL);
                    atomicAdd(&acc[2*accs+ia], (-dlx*fL + pLx*c2aL) - (-dsx*fS + pSx*c2aS));
                    atomicAdd(&acc[3*accs+ia], (-dly*fL + pLy*c2aL) - (-dsy*fS + pSy*c2aS));
                    atomicAdd(&acc[4*accs+ia], (-dlz*fL + pLz*c2aL) - (-dsz*fS + pSz*c2aS));
                    // hand the subtree sums to the parent (gaussvol.cpp:476-484)
                    S.aEL[sl] = EL; S.apsL[sl] = psL;
                    S.apLx[sl] = dlx*fL + pLx*c2bL; S.apLy[sl] = dly*fL + pLy*c2bL; S.apLz[sl] = dlz*fL + pLz*c2bL;
                    S.afL[sl] = S.dvvL[sl]*fL;
                    S.aES[sl] = ES; S.apsS[sl] = psS;
                    S.apSx[sl] = dsx*fS + pSx*c2bS; S.apSy[sl] = dsy*fS + pSy*c2bS; S.apSz[sl] = dsz*fS + pSz*c2bS;
                    S.afS[sl] = S.dvvS[sl]*fS;
                    vsumL += (double) (cf*vl); vsumS += (double) (cf*vs);
                    if (sl == 0) { eL_root = EL; eS_root = ES; }
                }
            }
            __syncwarp();
        }
        // E1 uses nu = +gamma/roffset, E2 uses nu = -gamma/roffset (ReferenceAGBNPKernels.cpp:297,364)
        eL_tot += (double) eL_root*(double) A.inv_roffset;
        eS_tot -= (double) eS_root*(double) A.inv_roffset;

        // ---- flush per-atom sums: self-volumes and the surface-tension force (force = -gradient) ----
        for (int i = lane; i <= nn; i += 32) {
            const int j = i == 0 ? r : nb_idx[i-1];
            const float s0 = acc[0*accs+i], s1 = acc[1*accs+i];
            if (s0 != 0.f) atomicAdd(&A.svS[j], (double) s0);
            if (s1 != 0.f) atomicAdd(&A.svL[j], (double) s1);
            const float gx = acc[2*accs+i], gy = acc[3*accs+i], gz = acc[4*accs+i];
            if (gx != 0.f || gy != 0.f || gz != 0.f) {
                add_force_fixed(&A.force[j], -gx*A.inv_roffset);
                add_force_fixed(&A.force[(size_t) A.np+j], -gy*A.inv_roffset);
                add_force_fixed(&A.force[2*(size_t) A.np+j], -gz*A.inv_roffset);
            }
        }

        // ---- persist what the gamma sweep needs ----
        int off = 0;
        if (lane == 0) off = atomicAdd(A.st.cursor, nslots);
        off = __shfl_sync(FULL, off, 0);
        if (off+nslots > A.st.cap) {
            if (lane == 0) atomicOr(A.status, ST_TREE_OVERFLOW);
        } else {
            if (lane == 0) { A.st.root_off[r] = off; A.st.root_cnt[r] = nslots; }
            if (lane <= nlev+1 && lane >= 1) A.st.root_lvs[r*MAX_LEVELS+lane] = (short) lvs[lane];
            for (int sl = lane; sl < nslots; sl += 32) {
                const int lev = S.lvl[sl];
                const float coefp = ((lev & 1) ? 1.f : -1.f)/(float) lev;
                const int ia = S.nbr[sl];
                const int ja = ia == 0 ? r : nb_idx[ia-1];
                const float a1iS = (float) S.gSa[sl], aiS = (float) A.aS[ja];
                const int o = off+sl;
                A.st.cs[o] = coefp*S.sfpS[sl];
                A.st.dvv[o] = S.dvvS[sl];
                A.st.dx[o] = S.dSx[sl]; A.st.dy[o] = S.dSy[sl]; A.st.dz[o] = S.dSz[sl];
                A.st.c2a[o] = aiS/a1iS; A.st.c2b[o] = (a1iS-aiS)/a1iS;
                A.st.atom[o] = ja;
                A.st.parent[o] = S.parent[sl]; A.st.cstart[o] = S.cstart[sl]; A.st.ccount[o] = S.ccount[sl];
                const int ps = S.perm[sl];                                  // node at sorted position sl
                if (sl > 0) A.st.rank[off+ps] = (short) (sl - S.cstart[S.parent[ps]]);
                else A.st.rank[off] = 0;
            }
        }
        __syncwarp();
    }

    eL_tot = warp_sum(eL_tot); eS_tot = warp_sum(eS_tot); vsumL = warp_sum(vsumL); vsumS = warp_sum(vsumS);
    if (lane == 0) {
        atomicAdd(&A.scalars[SC_EVOL_L], eL_tot);
        atomicAdd(&A.scalars[SC_EVOL_S], eS_tot);
        atomicAdd(&A.scalars[SC_VOL_L], vsumL);
        atomicAdd(&A.scalars[SC_VOL_S], vsumS);
        atomicAdd(&A.counters[CT_C2], c2_tot);
        atomicAdd(&A.counters[CT_C3], c3_tot);
        atomicAdd(&A.counters[CT_M], m_tot);
        atomicMax(A.hw_nbr, hw_nn);
        atomicMax(A.hw_nodes, hw_slots);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_tree_gamma: S10+S11 merged -- deposit nu_i = (W_i+U_i)/V_i on the stored tree (rescan_tree_g, gaussvol.cpp:330-372),
// run the energy-gradient up-sweep (gaussvol.cpp:400-487) and add force = -gradient
// (ReferenceAGBNPKernels.cpp:718-747; the merge of the W and U passes is exact because the sweep is linear in nu).
// Multi-GPU: each shard stores and sweeps only the subtrees of the roots it owns; the partial forces are all-reduced.
// ---------------------------------------------------------------------------------------------------------------
struct GammaArgs {
    int nh, np;
    TreeStore st;
    const float4* dacc;         // [np] .w = W_i + U_i
    const double* vS;           // atomic volumes, vdW radii
    unsigned long long* force;
    unsigned char* scratch;     // per-warp float[5*cap]: gam, f', p'x, p'y, p'z
    size_t scratch_stride;
    int cap;
    int* work_counter;
};

__global__ void __launch_bounds__(TREE_THREADS, 4) k_tree_gamma(GammaArgs A) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* gam = (float*) (A.scratch + (size_t) (blockIdx.x*TREE_WARPS+warp)*A.scratch_stride);
    float* af = gam+A.cap;
    float* apx = af+A.cap;
    float* apy = apx+A.cap;
    float* apz = apy+A.cap;
    for (;;) {
        int r = 0;
        if (lane == 0) r = atomicAdd(A.work_counter, 1);
        r = __shfl_sync(FULL, r, 0);
        if (r >= A.nh) break;
        const int cnt = A.st.root_cnt[r];
        if (cnt <= 1) continue;             // an atom without overlaps: dv1 = 0, no force (gaussvol.cpp:472)
        const int off = A.st.root_off[r];
        const short* lvs = A.st.root_lvs + r*MAX_LEVELS;
        int nlev = 1;
        while (nlev+1 < MAX_LEVELS && lvs[nlev+1] < cnt) nlev++;
        // top-down gamma1i (gaussvol.cpp:338-343)
        for (int lev = 1; lev <= nlev; lev++) {
            const int b = lvs[lev], e = lev == nlev ? cnt : lvs[lev+1];
            for (int sl = b+lane; sl < e; sl += 32) {
                const int ja = A.st.atom[off+sl];
                const float nu = (float) ((double) A.dacc[ja].w/A.vS[ja]);
                const int p = A.st.parent[off+sl];
                gam[sl] = (p >= 0 ? gam[p] : 0.f) + nu;
            }
            __syncwarp();
        }
        // bottom-up energy-gradient sweep
        for (int lev = nlev; lev >= 1; lev--) {
            const int b = lvs[lev], e = lev == nlev ? cnt : lvs[lev+1];
            for (int sl = b+lane; sl < e; sl += 32) {
                const int o = off+sl;
                float f = A.st.cs[o]*gam[sl], px = 0.f, py = 0.f, pz = 0.f;
                const int cs = A.st.cstart[o], ce = cs + A.st.ccount[o];
                for (int c = cs; c < ce; c++) { f += af[c]; px += apx[c]; py += apy[c]; pz += apz[c]; }
                const float dx = A.st.dx[o], dy = A.st.dy[o], dz = A.st.dz[o];
                const float c2a = A.st.c2a[o], c2b = A.st.c2b[o];
                const float gx = -dx*f + px*c2a, gy = -dy*f + py*c2a, gz = -dz*f + pz*c2a;
                const int ja = A.st.atom[o];
                if (gx != 0.f || gy != 0.f || gz != 0.f) {
                    add_force_fixed(&A.force[ja], -gx);
                    add_force_fixed(&A.force[(size_t) A.np+ja], -gy);
                    add_force_fixed(&A.force[2*(size_t) A.np+ja], -gz);
                }
                apx[sl] = dx*f + px*c2b; apy[sl] = dy*f + py*c2b; apz[sl] = dz*f + pz*c2b;
                af[sl] = A.st.dvv[o]*f;
            }
            __syncwarp();
        }
    }
}

} // namespace agbnp_b200_impl
#endif
