// libagbnp_b200.so -- C-ABI (include/agbnp_b200.h) and host driver of the sm_100a AGBNP1/GaussVol kernels.
// One handle = one CalcAGBNPForceKernel instance of the reference (openmmapi/include/AGBNPKernels.h:19-47).
// There is no CPU fallback anywhere in this file: every evaluation is the kernel sequence below.
#include "../../include/agbnp_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "agbnp_setup.h"
#include "agbnp_device.cuh"
#include "agbnp_tree.cuh"
#include "agbnp_pair.cuh"
#include "agbnp_peaks.cuh"

using namespace agbnp_b200_impl;

namespace agbnp_b200_impl {

// ---------------------------------------------------------------------------------------------------------------
// one-shot all-reduce over peer memory (see include/agbnp_b200.h, "peer-memory exchange")
// ---------------------------------------------------------------------------------------------------------------
constexpr int PEER_MAX = 8, PEER_KINDS = 7, PEER_KIND_POSITIONS = 6;     // kinds 0..5 = agbnp_b200_buffer, 6 = positions broadcast
struct PeerBox {
    int rank, count;
    unsigned char* mail[PEER_MAX];      // base of every shard's mailbox (mail[rank] is local)
    size_t kind_off[PEER_KINDS];        // offset of a buffer kind's slots; slot of source s at kind_off + s*kind_bytes
    size_t kind_bytes[PEER_KINDS];
    size_t flag_off;                    // int flags[PEER_KINDS][PEER_MAX] at the end of the mailbox
    int* counter;                       // local: [0] CTAs of the running kernel that have pushed, [1] that have finished
    int* epochs;                        // local: [PEER_KINDS] exchanges of each kind completed so far (device-side, so that the
                                        // launches carry no changing argument and a whole sharded evaluation is one CUDA graph)
    int* status;                        // local, sticky: ST_PEER_TIMEOUT once a wait has given up (k_finish then delivers nothing)
};

// Wait until a peer's flag reaches `epoch`.  Bounded: a peer that never arrives (its process died, or the shards were driven
// out of step) must not hang this GPU for ever -- after ~4 s of spinning the wait gives up, raises ST_PEER_TIMEOUT in the
// status word (the evaluation then delivers nothing and the host reports it) and the kernel runs to completion.
__device__ __forceinline__ void peer_wait(const volatile int* flag, int epoch, int* status, int kind = 0) {
    if (*flag >= epoch) return;
    const long long t0 = clock64();
    while (*flag < epoch) {
        if (clock64() - t0 > 8000000000ll) { atomicOr(status, ST_PEER_TIMEOUT | (1 << (8+kind))); break; }   // bits 8+: which exchange (diagnostics)
        __nanosleep(64);
    }
}

template <class T> __device__ __forceinline__ T peer_add(T a, T b) { return a+b; }
template <> __device__ __forceinline__ float4 peer_add<float4>(float4 a, float4 b) { return make_float4(a.x+b.x, a.y+b.y, a.z+b.z, a.w+b.w); }

// One kernel per exchange: every CTA pushes its slice of the partial buffer into the peers' mailboxes; the last CTA to
// finish raises this shard's flag at every peer; then every CTA waits for the peers' flags and adds their slices.
// The grid never exceeds the SM count, so all CTAs are resident and the spin cannot starve the CTA that raises the flag.
template <class T>
__global__ void __launch_bounds__(256) k_peer_allreduce(T* buf, size_t n, PeerBox pb, int kind) {
    const int epoch = pb.epochs[kind]+1;        // advanced by the last CTA to leave, i.e. after every CTA has read it
    for (size_t i = (size_t) blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x*blockDim.x) {
        const T v = buf[i];
        for (int p = 0; p < pb.count; p++)
            if (p != pb.rank) ((T*) (pb.mail[p] + pb.kind_off[kind] + (size_t) pb.rank*pb.kind_bytes[kind]))[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(pb.counter, 1);
        if (done == (int) gridDim.x-1) {                    // every CTA has fenced its stores: raise the flags
            *pb.counter = 0;
            __threadfence_system();
            for (int p = 0; p < pb.count; p++)
                if (p != pb.rank) *(volatile int*) (pb.mail[p] + pb.flag_off + sizeof(int)*(kind*PEER_MAX + pb.rank)) = epoch;
        }
        const volatile int* flags = (const volatile int*) (pb.mail[pb.rank] + pb.flag_off) + kind*PEER_MAX;
        for (int q = 0; q < pb.count; q++)
            if (q != pb.rank) peer_wait(flags+q, epoch, pb.status, kind);
        __threadfence_system();
    }
    __syncthreads();
    const unsigned char* base = pb.mail[pb.rank] + pb.kind_off[kind];
    for (size_t i = (size_t) blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x*blockDim.x) {
        T acc = buf[i];
        for (int q = 0; q < pb.count; q++)
            if (q != pb.rank) acc = peer_add(acc, __ldcg((const T*) (base + (size_t) q*pb.kind_bytes[kind]) + i));
        buf[i] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(pb.counter+1, 1) == (int) gridDim.x-1) { pb.counter[1] = 0; pb.epochs[kind] = epoch; }
}

// The last exchange of an evaluation, three in one: the partial forces of the gamma sweep (float4[n]), the energy scalars
// (SC_COUNT doubles) and this shard's status word, which travels as 0/1 in the SC_FAULT slot so that every shard's finish
// kernel knows whether ANY shard overflowed.  Same protocol as k_peer_allreduce, one flag (the FORCE kind's).
__global__ void __launch_bounds__(256) k_peer_allreduce_final(float4* buf, size_t n, double* scalars, const int* status, PeerBox pb) {
    const int kind = AGBNP_B200_BUF_FORCE, ekind = AGBNP_B200_BUF_ENERGY;
    const int epoch = pb.epochs[kind]+1;
    for (size_t i = (size_t) blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x*blockDim.x) {
        const float4 v = buf[i];
        for (int p = 0; p < pb.count; p++)
            if (p != pb.rank) ((float4*) (pb.mail[p] + pb.kind_off[kind] + (size_t) pb.rank*pb.kind_bytes[kind]))[i] = v;
    }
    double mine = 0.0;
    if (blockIdx.x == 0 && threadIdx.x < SC_COUNT) {
        mine = threadIdx.x == SC_FAULT ? (*status != 0 ? 1.0 : 0.0) : scalars[threadIdx.x];
        for (int p = 0; p < pb.count; p++)
            if (p != pb.rank) ((double*) (pb.mail[p] + pb.kind_off[ekind] + (size_t) pb.rank*pb.kind_bytes[ekind]))[threadIdx.x] = mine;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(pb.counter, 1);
        if (done == (int) gridDim.x-1) {
            *pb.counter = 0;
            __threadfence_system();
            for (int p = 0; p < pb.count; p++)
                if (p != pb.rank) *(volatile int*) (pb.mail[p] + pb.flag_off + sizeof(int)*(kind*PEER_MAX + pb.rank)) = epoch;
        }
        const volatile int* flags = (const volatile int*) (pb.mail[pb.rank] + pb.flag_off) + kind*PEER_MAX;
        for (int q = 0; q < pb.count; q++)
            if (q != pb.rank) peer_wait(flags+q, epoch, pb.status, kind);
        __threadfence_system();
    }
    __syncthreads();
    const unsigned char* base = pb.mail[pb.rank] + pb.kind_off[kind];
    for (size_t i = (size_t) blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x*blockDim.x) {
        float4 acc = buf[i];
        for (int q = 0; q < pb.count; q++)
            if (q != pb.rank) acc = peer_add(acc, __ldcg((const float4*) (base + (size_t) q*pb.kind_bytes[kind]) + i));
        buf[i] = acc;
    }
    if (blockIdx.x == 0 && threadIdx.x < SC_COUNT) {
        const unsigned char* ebase = pb.mail[pb.rank] + pb.kind_off[ekind];
        for (int q = 0; q < pb.count; q++)
            if (q != pb.rank) mine += __ldcg((const double*) (ebase + (size_t) q*pb.kind_bytes[ekind]) + threadIdx.x);
        scalars[threadIdx.x] = mine;
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(pb.counter+1, 1) == (int) gridDim.x-1) { pb.counter[1] = 0; pb.epochs[kind] = epoch; }
}

// broadcast of the positions from their owner: the owner stores them into every peer's mailbox and raises its flag there;
// the others wait for it and copy the slot into their own position buffer
__global__ void __launch_bounds__(256) k_peer_broadcast(float4* buf, size_t n, PeerBox pb, int owner) {
    const int kind = PEER_KIND_POSITIONS;
    const int epoch = pb.epochs[kind]+1;
    if (pb.rank == owner) {
        for (size_t i = (size_t) blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x*blockDim.x) {
            const float4 v = buf[i];
            for (int p = 0; p < pb.count; p++) if (p != owner) ((float4*) (pb.mail[p] + pb.kind_off[kind]))[i] = v;
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(pb.counter, 1) == (int) gridDim.x-1) {
            *pb.counter = 0;
            __threadfence_system();
            for (int p = 0; p < pb.count; p++)
                if (p != owner) *(volatile int*) (pb.mail[p] + pb.flag_off + sizeof(int)*(kind*PEER_MAX + owner)) = epoch;
        }
    } else {
        if (threadIdx.x == 0) {
            const volatile int* flags = (const volatile int*) (pb.mail[pb.rank] + pb.flag_off) + kind*PEER_MAX;
            peer_wait(flags+owner, epoch, pb.status, kind);
            __threadfence_system();
        }
        __syncthreads();
        const float4* src = (const float4*) (pb.mail[pb.rank] + pb.kind_off[kind]);
        for (size_t i = (size_t) blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x*blockDim.x) buf[i] = __ldcg(src+i);
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(pb.counter+1, 1) == (int) gridDim.x-1) { pb.counter[1] = 0; pb.epochs[kind] = epoch; }
}

} // namespace agbnp_b200_impl

namespace {

thread_local std::string g_create_error;

struct CudaFail { std::string msg; };
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    throw CudaFail{std::string(#call) + ": " + cudaGetErrorString(e_)}; } } while (0)

// Buffers replaced while a handle is in use (capacity growth, re-sort) are not cudaFree'd on the spot: cudaFree waits for ALL
// work on the device, including kernels of OTHER handles -- and a peer shard's exchange kernel may be spinning for this very
// handle's next launch (several shards on one GPU).  Inside a TrashScope they are parked in the handle's trash list instead
// and freed when the handle is destroyed (growth is geometric and rare: the parked memory is less than what is in use).
thread_local std::vector<void*>* g_trash = nullptr;
struct TrashScope {
    std::vector<void*>* prev;
    explicit TrashScope(std::vector<void*>* t) : prev(g_trash) { g_trash = t; }
    ~TrashScope() { g_trash = prev; }
};

template <class T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0, cap = 0;
    // (re)size; an allocation that is large enough is kept: cudaFree synchronises the device and cudaMalloc of tens of MB
    // takes a millisecond, and the periodic re-sort calls this for every order-dependent array
    void alloc(size_t count) {
        if (count <= cap && p) { n = count; return; }
        release();
        n = cap = count;
        if (count) CK(cudaMalloc((void**) &p, count*sizeof(T)));
    }
    void release() { if (p) { if (g_trash) g_trash->push_back(p); else cudaFree(p); } p = nullptr; n = cap = 0; }
    void upload(const std::vector<T>& h, cudaStream_t s) {
        if (h.size() > n) alloc(h.size());
        if (!h.empty()) CK(cudaMemcpyAsync(p, h.data(), h.size()*sizeof(T), cudaMemcpyHostToDevice, s));
    }
    ~DevBuf() { release(); }
};

enum KernelId { K_PREP = 0, K_TREE, K_BORN, K_BORNFIN, K_GB, K_DERIV, K_GAMMA, K_FINISH, K_BLIST, K_COUNT };
const char* const kKernelNames = "k_prep\nk_tree\nk_born\nk_born_finish\nk_gb\nk_deriv\nk_tree_gamma\nk_finish\nk_blocklist";

// control words inside the zeroed slab
enum Ctrl { CW_WORK_TREE = 0, CW_WORK_GB, CW_WORK_GAMMA, CW_STATUS, CW_TREE_CURSOR, CW_MAX_NBR, CW_MAX_NODES, CW_WORK_BORN, CW_WORK_DERIV, CW_MAX_WIDTH, CW_COUNT = 12 };

} // namespace

struct agbnp_b200 {
    agbnp_b200_config cfg;
    Constants k;
    SystemParams sp;
    std::string err;
    int n = 0, nh = 0, nhy = 0, nhp = 0, np = 0, nhb = 0, nb = 0;
    int num_sm = 148;
    cudaStream_t own_stream = nullptr;
    bool params_dirty = true, order_valid = false;
    long long evals_since_sort = 0, total_evals = 0;
    std::vector<int> orig;                  // sorted -> caller (size np, -1 padding)
    std::vector<float> box_lo, box_hi;      // block bounding boxes at sort time (host; unit ordering only)

    // static sorted arrays
    DevBuf<int> d_orig;
    DevBuf<int4> d_l2rec;
    DevBuf<float> d_charge, d_radius, d_alpha, d_gamma;
    DevBuf<double> d_aL, d_vL, d_aS, d_vS;
    DevBuf<unsigned char> d_rcbin, d_ts;
    DevBuf<signed char> d_tj;
    DevBuf<float> d_rc2, d_rc2max, d_rc2s, d_rc2maxs;     // level-2 pair radii squared; the same with the list skin
    DevBuf<int> d_l2list, d_l2cnt;          // per-root level-2 candidate lists kept between evaluations (agbnp_tree.cuh)
    DevBuf<float4> d_i4v;
    DevBuf<int2> d_units, d_pq_units;
    DevBuf<int> d_pq_toff;
    DevBuf<unsigned> d_pq_hits;
    DevBuf<uint2> d_pq_masks;
    DevBuf<float4> d_posq_ref;              // sorted positions when the pair masks were last built (agbnp_pair.cuh: PairUnits::ctl)
    DevBuf<int> d_pq_ctl;                   // [LC_COUNT] persistent control words of the list reuse (ListCtl)
    float pq_skin = 0.05f;                  // nm; AGBNP_B200_PAIR_SKIN overrides, 0 = rebuild the masks in every evaluation
    int nunits = 0, npq_units = 0;
    // per-evaluation arrays
    DevBuf<float4> d_posq, d_bbc, d_bbh, d_posq_in, d_gbj;
    DevBuf<float> d_vsf, d_born, d_bfp, d_brw, d_bmax;
    DevBuf<unsigned char> d_slab;           // zeroed every evaluation
    double* d_scalars = nullptr;
    unsigned long long* d_counters = nullptr;
    float4 *d_accL = nullptr, *d_accS = nullptr;   // tree: surface-tension gradient + self volume per atom (zeroed slab; accS follows accL)
    float4* d_gacc = nullptr;               // force of the W+U tree sweep (zeroed slab)
    float4* d_gbacc = nullptr;              // GB pair force + Y per atom (zeroed slab)
    float4* d_dacc = nullptr;               // derivative-pass force + (W+U) per atom (zeroed slab)
    float* d_bsum = nullptr;                // Born-radius pair sums (zeroed slab)
    int* d_root_cnt = nullptr;
    int* d_ctrl = nullptr;
    size_t slab_bytes = 0;
    DevBuf<float> d_force_out;              // float[3n] for the host path
    // layout of the caller's DEVICE buffers (agbnp_b200_set_device_layout); the host entry point ignores it
    DevBuf<int> d_io;                       // particle -> position, when the caller's buffers are in another atom order
    std::vector<int> io;                    // host copy
    bool io_set = false, posq_f64 = false, energy_f32 = false;
    bool cur_io = false;                    // the evaluation being enqueued came in through a device entry point
    // tree
    // tree capacities (grown on overflow): nodes per root, nodes per level, level-2 neighbors per root
    int tree_cap = 512, tree_wcap = 192, nbrmax = 64;
    int tree_grid = 0, tree_warps = 4, gamma_grid = 0, gb_grid = 0, gb_chunk = 8;
    int born_w = 8, born_c = 1, deriv_w = 8, deriv_c = 1, pq_shape_cutoff = -1;     // launch shapes of k_born / k_deriv (pq_shape)
    size_t pq_shape_tab = (size_t) -1;
    bool tree_work_global = false;          // work arrays too large for shared memory: per-warp global scratch instead
    DevBuf<unsigned char> d_tree_stage, d_tree_work, d_gamma_scratch;
    TreeStore st{};
    // opt-in tree reuse (cfg.tree_reuse_interval > 1): evaluations between builds re-evaluate the stored tree (k_tree_rescan)
    bool tree_built = false;                // a build evaluation has been enqueued for the current order / capacities
    int evals_since_build = 0;
    bool cur_eval_rescan = false;           // decision for the evaluation being enqueued (begin_eval)
    DevBuf<int> d_tree_ok;                  // device flag: the stored tree comes from a build evaluation without overflow
    size_t slab_keep_off = 0;               // slab offset of root_cnt: a rescan evaluation zeroes the slab only up to here
    DevBuf<int> d_root_off, d_bcount;
    DevBuf<int2> d_items;
    // tree work items, most expensive first (host copy): x = index of the item's first root in item_roots,
    // y = number of roots | part << 8 | parts << 16 (agbnp_tree.cuh: item_roots(), item_part(), item_parts())
    std::vector<int2> items;
    std::vector<int> item_roots;            // sorted indices of the items' roots
    DevBuf<int> d_item_roots;
    bool tree_group = true;                 // several small neighboring roots per item (AGBNP_B200_TREE_GROUP=0: one root per item)
    int max_items = 0;
    DevBuf<unsigned short> d_blist;
    float rc2_global = 0.f;
    DevBuf<short> d_root_lvs, d_st_rank;
    DevBuf<float4> d_st_rec;                // 2 float4 per node
    DevBuf<float> d_inv_vS;
    int gamma_warps = 4;
    bool gamma_work_global = false;
    // pinned host staging
    float4* h_posq = nullptr;
    float* h_force = nullptr;
    unsigned char* h_tail = nullptr;        // pinned mirror of the slab's [scalars | counters | control words] span: one copy per evaluation
    double* h_scal = nullptr;               // -> h_tail
    int* h_ctrl = nullptr;                  // -> h_tail + 512
    // diagnostics
    DevBuf<int2> d_pairs;
    cudaEvent_t ev[2] = {};
    bool have_events = false;
    cudaEvent_t tail_ev = nullptr;          // recorded behind everything this handle enqueues: "my own work is done"
    bool tail_valid = false;
    std::vector<void*> trash;               // device buffers replaced while in use, freed at destruction (TrashScope)
    long long launches = 0;                 // kernels launched by this handle (bench.py's gpu_launches)
    // per-kernel CUDA-event brackets (agbnp_b200_profile): events come from a pool so that a whole timed region can be
    // bracketed launch by launch and summed afterwards
    unsigned prof_mask = 0;
    std::vector<cudaEvent_t> prof_pool;
    size_t prof_used = 0;
    struct ProfRec { int id; size_t a, b; };
    std::vector<ProfRec> prof_recs;
    // deferred validation of asynchronous evaluations: status words of evaluation k land in slot k % ASYNC_DEPTH
    static constexpr int ASYNC_DEPTH = 4;
    int* h_async = nullptr;                 // pinned [ASYNC_DEPTH][CW_COUNT]
    cudaEvent_t async_ev[ASYNC_DEPTH] = {};
    bool async_pending[ASYNC_DEPTH] = {};
    long long async_issued = 0;
    bool async_fault = false;
    int deferred_rc = 0;                    // outcome of an earlier asynchronous evaluation, reported by the next call that returns one
    long long n_grow = 0, n_resort = 0, n_graph_inst = 0, n_async_fault = 0;    // AGBNP_B200_GET_STATS
    int ahead[16] = {};                     // capacities to grow before the next evaluation (grow_ahead)
    bool ahead_pending = false;
    // CUDA graphs of the whole kernel sequence, keyed by everything the launches bake in: `launch_gen` (bumped whenever
    // a buffer, capacity or launch shape changes) and the caller's pointers
    struct GraphEntry { long long gen; const void* posq; void* sink; int layout, padded_n; double* d_energy; bool sharded, rescan; cudaGraphExec_t exec; int kernels; long long last_use; };
    std::vector<GraphEntry> graphs;
    long long launch_gen = 0, graph_clock = 0;
    bool use_graph = true;
    bool use_pdl = true;             // programmatic dependent launch between the kernels of an evaluation (AGBNP_B200_NO_PDL=1: off)
    // peer-memory exchange
    PeerBox peer{};
    unsigned char* d_mailbox = nullptr;
    size_t mailbox_bytes = 0;
    int* d_peer_counter = nullptr;
    bool peer_ready = false;
    std::vector<void*> peer_opened;

    ~agbnp_b200() {
        if (h_posq) cudaFreeHost(h_posq);
        if (h_force) cudaFreeHost(h_force);
        if (h_tail) cudaFreeHost(h_tail);
        if (have_events) { for (auto& e : ev) cudaEventDestroy(e); for (auto& e : async_ev) cudaEventDestroy(e); cudaEventDestroy(tail_ev); }
        for (void* q : trash) cudaFree(q);
        for (auto& e : prof_pool) cudaEventDestroy(e);
        for (auto& g : graphs) cudaGraphExecDestroy(g.exec);
        for (void* p : peer_opened) cudaIpcCloseMemHandle(p);
        if (d_mailbox) cudaFree(d_mailbox);
        if (d_peer_counter) cudaFree(d_peer_counter);
        if (h_async) cudaFreeHost(h_async);
        if (own_stream) cudaStreamDestroy(own_stream);
    }
};

namespace {

// CUDA loads a kernel's code lazily, at its first launch, and the load waits for the device to drain: a shard whose first
// launch of some kernel comes while a PEER shard on the same device is already spinning in an exchange (waiting for this very
// shard) would stall until the peer's wait times out.  Every kernel of the library is therefore loaded when a handle is
// created, before anything can be waiting (cudaFuncGetAttributes forces the load).
template <class K> void preload(K kernel) { cudaFuncAttributes a; CK(cudaFuncGetAttributes(&a, kernel)); }
void preload_kernels() {
    static bool done[64] = {};
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (done[dev & 63]) return;
    done[dev & 63] = true;
    preload(k_prep); preload(k_blocklist); preload(k_tree<true>); preload(k_tree<false>); preload(k_tree_rescan);
    preload(k_born<false, true>); preload(k_born<true, true>); preload(k_born<false, false>); preload(k_born<true, false>);
    preload(k_born_finish); preload(k_gb<false>); preload(k_gb<true>);
    preload(k_deriv<false, true>); preload(k_deriv<true, true>); preload(k_deriv<false, false>); preload(k_deriv<true, false>);
    preload(k_tree_gamma<true>); preload(k_tree_gamma<false>); preload(k_status_fold); preload(k_finish); preload(k_list_pairs);
    preload(k_peer_allreduce<float4>); preload(k_peer_allreduce<float>); preload(k_peer_allreduce<double>);
    preload(k_peer_allreduce_final); preload(k_peer_broadcast);
}

// wait for the work THIS handle has enqueued (never cudaDeviceSynchronize: see TrashScope)
void wait_own_work(agbnp_b200* h) {
    if (h->tail_valid) CK(cudaEventSynchronize(h->tail_ev));
}
void mark_tail(agbnp_b200* h, cudaStream_t s) {
    CK(cudaEventRecord(h->tail_ev, s));
    h->tail_valid = true;
}

void alloc_store(agbnp_b200* h, int cap) {
    h->launch_gen++;
    h->tree_built = false;
    h->d_st_rec.alloc((size_t) 2*cap);
    h->d_st_rank.alloc(cap);
    TreeStore& s = h->st;
    s.cap = cap;
    s.rec = h->d_st_rec.p; s.rank = h->d_st_rank.p;
    s.root_off = h->d_root_off.p; s.root_cnt = h->d_root_cnt; s.root_lvs = h->d_root_lvs.p;
}

// per-root level-2 candidate lists (stride nbrmax); a new order or a new stride voids every stored list
void alloc_l2_lists(agbnp_b200* h, cudaStream_t s) {
    if (h->nhp <= 0) return;
    h->d_l2list.alloc((size_t) h->nhp*h->nbrmax);
    h->d_l2cnt.alloc(h->nhp);
    h->d_pq_ctl.alloc(LC_COUNT);
    CK(cudaMemsetAsync(h->d_pq_ctl.p, 0, LC_COUNT*sizeof(int), s));
}

// choose the launch shape of k_tree for the current capacities and (re)allocate its per-warp buffers
void alloc_tree_scratch(agbnp_b200* h) {
    h->launch_gen++;
    h->tree_built = false;
    alloc_l2_lists(h, h->own_stream);
    if (h->nhp > 0) CK(cudaStreamSynchronize(h->own_stream));
    const size_t per_warp = tree_work_bytes(h->nbrmax, h->tree_cap, h->tree_wcap);
    // shared memory the tree kernels may fill per SM, of 228 KB (every CTA also pays 1 KB).  Round 1 left 28 KB "for L1";
    // measured (r2ac): at the capacities a thermal state needs (576, 224, 96) the full budget fits 16 instead of 14 warps of
    // k_tree (180 -> 167.5 us, HIV-RT MD step 0.517 -> 0.491 ms), and the gamma sweep runs 16 instead of 12 warps at the default
    // capacities (35.2 -> 32.7 us); k_tree at the default capacities fits 16 warps either way.
    static const size_t smem_kb = std::getenv("AGBNP_B200_TREE_SMEM_KB") ? (size_t) std::atoi(std::getenv("AGBNP_B200_TREE_SMEM_KB")) : 226;
    const size_t smem_sm = smem_kb*1024;
    // CTAs of 2 warps (warps never cooperate); 128 registers/thread bound the residency at 16 warps per SM
    h->tree_warps = TREE_WARPS;
    size_t ctas = std::min<size_t>(TREE_SMEM_CTAS, smem_sm/(h->tree_warps*per_warp + 1024));
    h->tree_work_global = ctas < 2;                        // fewer than 4 warps per SM: shared memory no longer pays
    if (h->tree_work_global) { h->tree_warps = 8; ctas = 2; }
    h->tree_grid = h->num_sm*(int) ctas;
    const size_t nwarps = (size_t) h->tree_grid*h->tree_warps;
    h->d_tree_stage.alloc(nwarps*tree_stage_bytes(h->tree_cap, h->tree_wcap));
    if (h->tree_work_global) h->d_tree_work.alloc(nwarps*per_warp); else h->d_tree_work.release();
    const size_t smem = h->tree_work_global ? 0 : h->tree_warps*per_warp;
    // the attribute belongs to the function, not to the handle: several handles in one process (replicas, shards on one
    // device) may need different amounts, so it only ever grows
    static size_t tree_smem_max[64] = {};
    size_t& tmax = tree_smem_max[h->cfg.device & 63];
    if (smem > tmax) { tmax = smem; CK(cudaFuncSetAttribute(k_tree<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem, 1024))); }
    // k_tree_rescan (opt-in tree reuse) keeps parent / atom / flag per node in shared memory: 2 warps x (8 cap + 48) bytes
    // passes 48 KB at cap ~3070, while k_tree is still in shared-memory mode up to cap ~5800
    {
        static size_t rescan_smem_max[64] = {};
        size_t& rmax = rescan_smem_max[h->cfg.device & 63];
        const size_t rs = h->tree_work_global ? 0 : h->tree_warps*rescan_work_bytes(h->tree_cap);
        if (rs > rmax) { rmax = rs; CK(cudaFuncSetAttribute(k_tree_rescan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(rs, 1024))); }
    }
    // k_tree_gamma: per-warp gamma_1..n and children sums, in shared memory while they fit
    {
        const size_t pw = gamma_work_bytes(h->tree_cap);
        h->gamma_warps = 4;
        size_t gctas = std::min<size_t>(8, smem_sm/(h->gamma_warps*pw + 1024));
        h->gamma_work_global = gctas < 2;
        if (h->gamma_work_global) gctas = 4;
        h->gamma_grid = h->num_sm*(int) gctas;
        if (h->gamma_work_global) h->d_gamma_scratch.alloc((size_t) h->gamma_grid*h->gamma_warps*pw); else h->d_gamma_scratch.release();
        static size_t gamma_smem_max[64] = {};
        size_t& gmax = gamma_smem_max[h->cfg.device & 63];
        const size_t gs = h->gamma_work_global ? 0 : h->gamma_warps*pw;
        if (gs > gmax) { gmax = gs; CK(cudaFuncSetAttribute(k_tree_gamma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(gs, 1024))); }
    }
}

// (re)compute the internal atom order from caller-order positions and upload every order-dependent static array
void build_order(agbnp_b200* h, const float* xyz, int stride, cudaStream_t s) {
    const SystemParams& sp = h->sp;
    std::vector<int> hv, hy;
    morton_order(xyz, stride, sp.ishydrogen, hv, hy);
    h->nh = (int) hv.size(); h->nhy = (int) hy.size();
    h->nhp = (h->nh+TILE-1)/TILE*TILE;
    const int nhyp = (h->nhy+TILE-1)/TILE*TILE;
    h->np = h->nhp+nhyp;
    h->nhb = h->nhp/TILE; h->nb = h->np/TILE;
    h->orig.assign(h->np, -1);
    for (int i = 0; i < h->nh; i++) h->orig[i] = hv[i];
    for (int i = 0; i < h->nhy; i++) h->orig[h->nhp+i] = hy[i];
    h->order_valid = true;
    h->n_resort++;
    h->tree_built = false;                  // the stored tree is indexed by the old order
    h->params_dirty = true;
    h->evals_since_sort = 0;
    (void) s;
    // roots ordered by the number of heavy atoms within 0.5 nm (a proxy for the subtree size, which grows like its 2nd-3rd
    // power): the work-stealing loop of k_tree then starts the expensive subtrees first (longest-processing-time order)
    {
        const int nh = h->nh;
        const float cell = 0.5f;
        float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
        for (int k = 0; k < nh; k++) for (int c = 0; c < 3; c++) { const float v = xyz[(size_t) h->orig[k]*stride+c]; lo[c] = std::min(lo[c], v); hi[c] = std::max(hi[c], v); }
        int dim[3];
        for (int c = 0; c < 3; c++) dim[c] = nh ? std::min(256, std::max(1, (int) ((hi[c]-lo[c])/cell)+1)) : 1;
        auto cell_of = [&](int k, int* ic) {
            for (int c = 0; c < 3; c++) ic[c] = std::min(dim[c]-1, std::max(0, (int) ((xyz[(size_t) h->orig[k]*stride+c]-lo[c])/cell)));
        };
        std::vector<std::vector<int>> cells((size_t) dim[0]*dim[1]*dim[2]);
        for (int k = 0; k < nh; k++) { int ic[3]; cell_of(k, ic); cells[((size_t) ic[0]*dim[1]+ic[1])*dim[2]+ic[2]].push_back(k); }
        std::vector<int> count(nh, 0);
        for (int k = 0; k < nh; k++) {
            int ic[3]; cell_of(k, ic);
            const float* pk = xyz + (size_t) h->orig[k]*stride;
            for (int a = std::max(0, ic[0]-1); a <= std::min(dim[0]-1, ic[0]+1); a++)
            for (int b = std::max(0, ic[1]-1); b <= std::min(dim[1]-1, ic[1]+1); b++)
            for (int c = std::max(0, ic[2]-1); c <= std::min(dim[2]-1, ic[2]+1); c++)
                for (int j : cells[((size_t) a*dim[1]+b)*dim[2]+c]) {
                    if (h->orig[j] <= h->orig[k]) continue;           // level-2 children are later atoms only
                    const float* pj = xyz + (size_t) h->orig[j]*stride;
                    const float dx = pj[0]-pk[0], dy = pj[1]-pk[1], dz = pj[2]-pk[2];
                    if (dx*dx + dy*dy + dz*dz < cell*cell) count[k]++;
                }
        }
        // work items: a root whose neighbor count predicts a large subtree is split into parts (agbnp_tree.cuh), so that no
        // single warp carries a subtree that bounds the kernel's duration
        h->max_items = 2*h->nhp + 64;
        struct It { int first, nroots, part, parts; float cost; };
        std::vector<It> its;
        h->item_roots.clear();
        int budget = h->max_items - nh;
        // cost model: a subtree grows like the cube of its level-2 list.  A root is split when it alone would exceed twice
        // the balanced load of one resident warp (16 per SM, all shards together): large systems split only their extreme
        // tail (splitting repeats the level-2 work), small ones split everything that shortens the critical path
        double total = 0;
        for (int k = 0; k < nh; k++) total += 1.0 + (double) count[k]*count[k]*count[k];
        const double warps = (double) h->num_sm*16*h->cfg.shard_count;
        const double per_warp = total/warps;
        // Small roots that follow each other in the (Morton) order share one item: a warp builds their subtrees together,
        // level by level, so that the per-level passes run on fuller chunks and their latency is paid once (the average
        // level of a single root offers 25 nodes to 32 lanes).  Predictions from the later-neighbor count c (fit on 2clr):
        // nodes ~ 0.85 (c+1)^1.85, widest level ~ 0.4 nodes, listed level-2 candidates (with the skin) ~ 2.5 c.  A group is
        // closed when it would pass the node target -- sized so that every resident warp still gets about six items (measured on
        // HIV-RT: k_tree 161.7 / 152.1 / 150.4 / 151.1 us with 3 / 4 / 6 / 10 items per warp, 154.6 us ungrouped: larger groups
        // lose to load imbalance what they gain in lane use, so in effect only the smallest roots are merged) -- or
        // what the DEFAULT per-level / level-2 capacities hold comfortably (fixed numbers: every shard must form the same
        // items whatever its own capacities have grown to); the capacities grow as usual if a prediction was too low.
        double pred_total = 0;
        std::vector<float> pred(nh);
        for (int k = 0; k < nh; k++) { pred[k] = 0.85f*std::pow((float) count[k]+1.f, 1.85f); pred_total += pred[k]; }
        static const double items_per_warp = std::getenv("AGBNP_B200_TREE_GROUP_TARGET") ? std::atof(std::getenv("AGBNP_B200_TREE_GROUP_TARGET")) : 6.0;
        const float node_target = (float) std::min(300.0, std::max(40.0, pred_total/(items_per_warp*warps)));
        const float width_target = 80.f, nbr_target = 36.f;
        It open{0, 0, 0, 1, 0.f};
        float g_nodes = 0.f, g_nbr = 0.f;
        auto close = [&]() { if (open.nroots) { its.push_back(open); open.nroots = 0; } };
        for (int k = 0; k < nh; k++) {
            const double c = 1.0 + (double) count[k]*count[k]*count[k];
            int parts = (int) std::min(8.0, std::ceil(c/std::max(2.0*per_warp, 64.0)));
            parts = std::max(1, std::min(parts, count[k]/4));
            parts = std::min(parts, 1+std::max(0, budget));
            budget -= parts-1;
            const float nbr_k = 2.5f*(float) count[k];
            if (parts > 1 || !h->tree_group) {
                close();
                const int first = (int) h->item_roots.size();
                h->item_roots.push_back(k);
                for (int q = 0; q < parts; q++) its.push_back({first, 1, q, parts, (float) (c/parts)});
                continue;
            }
            if (open.nroots && (open.nroots >= TREE_GROUP_MAX || g_nodes + pred[k] > node_target || 0.4f*(g_nodes + pred[k]) > width_target ||
                                g_nbr + nbr_k > nbr_target)) close();
            if (!open.nroots) { open = It{(int) h->item_roots.size(), 0, 0, 1, 0.f}; g_nodes = 0.f; g_nbr = 0.f; }
            h->item_roots.push_back(k);
            open.nroots++; open.cost += (float) c; g_nodes += pred[k]; g_nbr += nbr_k;
        }
        close();
        // (ordering by predicted nodes instead of the cubic cost changes nothing measurable, r2ag)
        std::stable_sort(its.begin(), its.end(), [](const It& a, const It& b) { return a.cost > b.cost; });
        h->items.clear();
        for (const It& t : its) h->items.push_back(make_int2(t.first, t.nroots | (t.part << 8) | (t.parts << 16)));
    }
    // block bounding boxes at sort time: used only to ORDER and PACK the work units of the range-limited pair passes
    // (heaviest first, far-apart block pairs packed several per unit); membership is decided on the device every evaluation
    h->box_lo.assign((size_t) 3*h->nb, 3.0e38f); h->box_hi.assign((size_t) 3*h->nb, -3.0e38f);
    for (int k = 0; k < h->np; k++) {
        const int o = h->orig[k];
        if (o < 0) continue;
        for (int c = 0; c < 3; c++) {
            const float v = xyz[(size_t) o*stride+c];
            float& lo = h->box_lo[(size_t) 3*(k/TILE)+c]; float& hi = h->box_hi[(size_t) 3*(k/TILE)+c];
            lo = std::min(lo, v); hi = std::max(hi, v);
        }
    }
}

// work units of k_born / k_deriv: (row block, first column block | count << 20), see agbnp_pair.cuh
void build_pq_units(agbnp_b200* h, std::vector<int2>& out) {
    const bool cutoff = h->cfg.nonbonded_method == AGBNP_B200_CUTOFF_NONPERIODIC;
    const double lim = cutoff ? std::min(h->k.i4_maxa, h->cfg.cutoff) : h->k.i4_maxa;
    struct U { int ra, cb0, n; float cost; };
    std::vector<U> us;
    auto box_dist = [&](int a, int b) {
        double d2 = 0;
        for (int c = 0; c < 3; c++) {
            const double g = std::max(0.0, std::max((double) h->box_lo[3*a+c]-h->box_hi[3*b+c], (double) h->box_lo[3*b+c]-h->box_hi[3*a+c]));
            d2 += g*g;
        }
        return std::sqrt(d2);
    };
    auto tile_cost = [&](int ra, int cb) {
        const double d = box_dist(ra, cb);
        const double f = std::min(1.0, std::max(0.0, (lim-d)/1.0 + 0.3));
        float cost = 0.02f + (float) (f*f)*((ra < h->nhb && cb != ra) ? 2.f : 1.f);
        if (!(d < 1e30)) cost = 0.02f;                      // a block of padding only
        return cost;
    };
    // Unit size: a unit is closed at `unit_cost` (one near tile of heavy rows costs 2).  Every unit pays a claim and three or
    // four dependent loads before its first pair, so large systems want larger units -- measured on B200 (r2w): HIV-RT
    // (41.7 k cost units) k_born 56.2 / 55.6 / 53.9 / 52.1 / 64.7 us at 0.5 / 1 / 2 / 4 / 8, 2clr (12 k) 23.3 / 23.3 / 25.0 /
    // 33.2 / 43.4 us -- about two and a half units per resident warp of k_born (36 per SM), within [1, 4].
    float unit_cost = 1.0f;
    {
        double total = 0;
        for (int ra = 0; ra < h->nb; ra++)
            for (int cb = ra < h->nhb ? ra : 0; cb < h->nhb; cb++) total += tile_cost(ra, cb);
        unit_cost = (float) std::min(4.0, std::max(1.0, total/(2.5*36.0*h->num_sm*h->cfg.shard_count)));
        if (const char* e = std::getenv("AGBNP_B200_PQ_UNIT_COST")) unit_cost = (float) std::atof(e);
    }
    for (int ra = 0; ra < h->nb; ra++) {
        U cur{ra, 0, 0, 0.f};
        for (int cb = ra < h->nhb ? ra : 0; cb < h->nhb; cb++) {
            const float cost = tile_cost(ra, cb);
            if (cur.n == 0) cur.cb0 = cb;
            cur.n++; cur.cost += cost;
            if (cur.cost >= unit_cost || cur.n == PQ_CHUNK) { us.push_back(cur); cur.n = 0; cur.cost = 0.f; }
        }
        if (cur.n) us.push_back(cur);
    }
    if (std::getenv("AGBNP_B200_DEBUG_UNITS")) { double ct = 0; for (const U& u : us) ct += u.cost; std::fprintf(stderr, "pq units %zu total cost %.1f unit_cost %.2f nb %d nhb %d\n", us.size(), ct, unit_cost, h->nb, h->nhb); }
    std::stable_sort(us.begin(), us.end(), [](const U& a, const U& b) { return a.cost > b.cost; });
    // (Tried, r2ag: cutting the units that hold the last 40-60 % of the cost again, at a quarter of the size, because the average
    // k_born / k_deriv warp is done at 81 / 87 % of its kernel's span.  The warps then finish together -- 92 / 96 % -- but the
    // extra units cost what the shorter tail gains: k_born 52.6 -> 54.0 us, k_deriv 77.1 -> 76.3 us.)
    out.clear();
    for (const U& u : us) out.push_back(make_int2(u.ra, u.cb0 | (u.n << 20)));
}

void upload_static(agbnp_b200* h, cudaStream_t s) {
    h->launch_gen++;
    h->tree_built = false;
    const SystemParams& sp = h->sp;
    const int np = h->np;
    std::vector<float> charge(np, 0.f), radius(np, 0.15f), alpha(np, 0.f), gamma(np, 0.f);
    std::vector<double> aL(np, 1.0), vL(np, 0.0), aS(np, 1.0), vS(np, 0.0);
    std::vector<float> inv_vS(np, 0.f);
    std::vector<unsigned char> rcbin(np, 0), ts(np, 0);
    std::vector<signed char> tj(np, -1);
    for (int k = 0; k < np; k++) {
        const int o = h->orig[k];
        if (o < 0) continue;
        charge[k] = (float) sp.charge[o]; radius[k] = (float) sp.radius[o]; alpha[k] = (float) sp.alpha[o];
        gamma[k] = (float) sp.gamma[o];
        aL[k] = sp.aL[o]; vL[k] = sp.vL[o]; aS[k] = sp.aS[o]; vS[k] = sp.vS[o];
        inv_vS[k] = sp.vS[o] > 0 ? (float) (1.0/sp.vS[o]) : 0.f;
        rcbin[k] = (unsigned char) sp.rc_bin[o];
        ts[k] = (unsigned char) sp.i4.type_screened[o];
        tj[k] = (signed char) sp.i4.type_screener[o];
    }
    h->d_orig.upload(h->orig, s);
    {
        std::vector<int4> ob(np, make_int4(-1, 0, 0, 0));
        for (int k = 0; k < np; k++) {
            if (h->orig[k] < 0) continue;
            const float af = (float) aL[k], vf = (float) vL[k];
            int ai, vi;
            std::memcpy(&ai, &af, 4); std::memcpy(&vi, &vf, 4);
            ob[k] = make_int4(h->orig[k] | ((int) rcbin[k] << 24), ai, vi, 0);
        }
        h->d_l2rec.upload(ob, s);
    }
    h->d_charge.upload(charge, s); h->d_radius.upload(radius, s); h->d_alpha.upload(alpha, s); h->d_gamma.upload(gamma, s);
    h->d_aL.upload(aL, s); h->d_vL.upload(vL, s); h->d_aS.upload(aS, s); h->d_vS.upload(vS, s); h->d_inv_vS.upload(inv_vS, s);
    h->d_rcbin.upload(rcbin, s); h->d_ts.upload(ts, s); h->d_tj.upload(tj, s);
    h->d_rc2.upload(sp.rc2, s); h->d_rc2max.upload(sp.rc2max, s);
    {
        // the same radii with the list skin: (sqrt(rc2) + skin)^2, rounded up
        auto skinned = [&](float r2) { const float r = std::sqrt(r2) + h->pq_skin; return r*r*1.000001f; };
        std::vector<float> a(sp.rc2.size()), b(sp.rc2max.size());
        for (size_t i = 0; i < a.size(); i++) a[i] = skinned(sp.rc2[i]);
        for (size_t i = 0; i < b.size(); i++) b[i] = skinned(sp.rc2max[i]);
        h->d_rc2s.upload(a, s); h->d_rc2maxs.upload(b, s);
    }
    h->rc2_global = 0.f;
    for (float v : sp.rc2max) h->rc2_global = std::max(h->rc2_global, v);
    { const float r = std::sqrt(h->rc2_global) + h->pq_skin; h->rc2_global = r*r*1.000001f; }
    h->d_items.upload(h->items, s);
    h->d_item_roots.upload(h->item_roots, s);
    h->d_bcount.alloc(std::max(1, h->nhb)); h->d_blist.alloc((size_t) std::max(1, h->nhb)*BLIST_MAX);
    // I4 splines in power form around the left knot (see agbnp_pair.cuh): with zl = y2_k h^2/6, zu = y2_{k+1} h^2/6,
    //   y(fr) = yl + fr [(yu-yl) - 2 zl - zu] + fr^2 [3 zl] + fr^3 [zu - zl]
    {
        const I4Tables& t = sp.i4;
        const int ntab = t.ntypes_screened*t.ntypes_screener, ni = t.nodes-1;
        std::vector<float4> tv((size_t) ntab*ni);
        for (int tb = 0; tb < ntab; tb++) for (int k = 0; k < ni; k++) {
            const double yl = t.y[(size_t) tb*t.nodes+k], yu = t.y[(size_t) tb*t.nodes+k+1];
            const double zl = t.y2[(size_t) tb*t.nodes+k]*t.h*t.h/6.0, zu = t.y2[(size_t) tb*t.nodes+k+1]*t.h*t.h/6.0;
            const double v0 = yl, v1 = (yu-yl) - 2.0*zl - zu, v2 = 3.0*zl, v3 = zu-zl;
            tv[(size_t) tb*ni+k] = make_float4((float) v0, (float) v1, (float) v2, (float) v3);
        }
        h->d_i4v.upload(tv, s);
    }
    // GB work units: triangular cover of the block-pair matrix in chunks of gb_chunk column tiles -- up to GB_CHUNK, fewer for
    // small systems so that every resident warp still gets about four units (2clr: 17.6 k tiles over 1776 warps)
    {
        const long long tiles_total = (long long) h->nb*(h->nb+1)/2, warps = (long long) h->num_sm*GB_MIN_BLOCKS*(GB_THREADS/32);
        h->gb_chunk = (int) std::max<long long>(1, std::min<long long>(GB_CHUNK, tiles_total/(4*std::max<long long>(warps, 1))));
    }
    std::vector<int2> units;
    // (row block, first column block | number of column blocks << 20).  Units are claimed in this order, so the last ones set
    // the kernel's tail: the final fifth of the tiles goes out in half- and quarter-size units (AGBNP_B200_GB_TAIL=0: all equal)
    {
        const bool taper = !(std::getenv("AGBNP_B200_GB_TAIL") && std::atoi(std::getenv("AGBNP_B200_GB_TAIL")) == 0);
        const long long tiles_total = (long long) h->nb*(h->nb+1)/2;
        long long done = 0;
        for (int ra = 0; ra < h->nb; ra++)
            for (int c = ra; c < h->nb; ) {
                int n = h->gb_chunk;
                if (taper && done*100 >= tiles_total*92) n = std::max(1, h->gb_chunk/4);
                else if (taper && done*100 >= tiles_total*80) n = std::max(1, h->gb_chunk/2);
                n = std::min(n, h->nb-c);
                units.push_back(make_int2(ra, c | (n << 20)));
                c += n; done += n;
            }
    }
    h->nunits = (int) units.size();
    h->d_units.upload(units, s);
    // range-limited pair passes: heavy rows x heavy columns cb >= ra, then hydrogen rows x heavy columns (agbnp_pair.cuh)
    std::vector<int2> pq;
    build_pq_units(h, pq);
    h->npq_units = (int) pq.size();
    h->d_pq_units.upload(pq, s);
    {
        std::vector<int> toff(pq.size());
        int t = 0;
        for (size_t i = 0; i < pq.size(); i++) { toff[i] = t; t += pq[i].y >> 20; }
        h->d_pq_toff.upload(toff, s);
        h->d_pq_hits.alloc(std::max<size_t>(1, pq.size()));
        h->d_pq_masks.alloc((size_t) std::max(1, t)*TILE);
        // new order, new units: the stored masks and candidate lists are void
        h->d_posq_ref.alloc(np);
        alloc_l2_lists(h, s);
    }
    CK(cudaStreamSynchronize(s));      // the host vectors above go out of scope
    // per-evaluation arrays
    if (h->d_posq.n < (size_t) np) {
        h->d_posq.alloc(np); h->d_bbc.alloc(h->nb); h->d_bbh.alloc(h->nb); h->d_gbj.alloc((size_t) 3*np);
        h->d_vsf.alloc(np); h->d_born.alloc(np); h->d_bfp.alloc(np); h->d_brw.alloc(np); h->d_bmax.alloc(h->nb);
        size_t o = 0;
        auto take = [&](size_t bytes) { size_t r = o; o += (bytes+255)/256*256; return r; };
        const size_t o_accL = take(sizeof(float4)*np*2), o_gacc = take(sizeof(float4)*np);
        const size_t o_yq = take(sizeof(float4)*np), o_scal = take(sizeof(double)*SC_COUNT), o_cnt = take(sizeof(unsigned long long)*CT_COUNT);
        const size_t o_ctrl = take(sizeof(int)*CW_COUNT);
        const size_t o_dacc = take(sizeof(float4)*np), o_bsum = take(sizeof(float)*np), o_rcnt = take(sizeof(int)*h->max_items);
        h->slab_bytes = o;
        h->slab_keep_off = o_rcnt;
        h->d_slab.alloc(o);
        unsigned char* b = h->d_slab.p;
        h->d_accL = (float4*) (b+o_accL); h->d_accS = h->d_accL+np; h->d_gacc = (float4*) (b+o_gacc);
        h->d_gbacc = (float4*) (b+o_yq); h->d_scalars = (double*) (b+o_scal); h->d_counters = (unsigned long long*) (b+o_cnt);
        h->d_ctrl = (int*) (b+o_ctrl);
        h->d_dacc = (float4*) (b+o_dacc); h->d_bsum = (float*) (b+o_bsum); h->d_root_cnt = (int*) (b+o_rcnt);
        h->d_root_off.alloc(h->max_items); h->d_root_lvs.alloc((size_t) h->max_items*MAX_LEVELS);
        if (h->st.cap == 0) alloc_store(h, std::max(4096, 160*h->nh + 4096));
        else alloc_store(h, h->st.cap);
    }
    h->params_dirty = false;
}

// launch shape of the range-limited pair passes: the warps per CTA (w) and CTAs per SM (c) that put the most warps on an SM
// (one copy of the spline table per CTA, warp_bytes of shared memory per warp; registers and everything else through the
// occupancy calculator); ties go to the smaller CTAs.  The grid is num_sm*c CTAs, all resident (first_unit).
template <typename Args>
void pq_shape(void (*kern)(Args), size_t tab_bytes, size_t warp_bytes, const char* env, int& w_out, int& c_out) {
    const char* e = std::getenv(env);
    const int force_w = e ? std::atoi(e) : 0;
    int best = 0;
    w_out = 8; c_out = 1;
    for (int w = 4; w <= PQ_MAX_THREADS/32; w++) {
        if (force_w && w != force_w) continue;
        const size_t sm = tab_bytes + w*warp_bytes;
        if (sm > (size_t) 227*1024) continue;
        int c = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, kern, 32*w, sm) != cudaSuccess) { cudaGetLastError(); continue; }
        if (c < 1) continue;
        if (c*w > best || (c*w == best && c > c_out)) { best = c*w; w_out = w; c_out = c; }
    }
}

PairCommon pair_common(agbnp_b200* h) {
    PairCommon c;
    c.np = h->np; c.nhb = h->nhb; c.nb = h->nb;
    c.posq = h->d_posq.p; c.orig = h->d_orig.p; c.bbc = h->d_bbc.p; c.bbh = h->d_bbh.p;
    c.ts = h->d_ts.p; c.tj = h->d_tj.p; c.i4v = h->d_i4v.p;
    c.ntj = h->sp.i4.ntypes_screener;
    c.ntables = h->sp.i4.ntypes_screened*h->sp.i4.ntypes_screener;
    c.tab_smem = (size_t) c.ntables*I4_INTERVALS*sizeof(float4) <= 24*1024;     // both tables + atom staging stay under 64 KB
    c.inv_h = (float) (1.0/h->sp.i4.h);
    c.range2 = (float) (h->k.i4_maxa*h->k.i4_maxa);
    const float cut = (float) h->cfg.cutoff;
    c.cut2 = cut*cut;
    // shard: contiguous ranges of row blocks
    const int per = (h->nb + h->cfg.shard_count-1)/h->cfg.shard_count;
    c.row_begin = std::min(h->nb, per*h->cfg.shard_rank);
    c.row_end = std::min(h->nb, c.row_begin+per);
    return c;
}

enum Phase { PH_TREE = 1, PH_GB = 2, PH_DERIV = 4, PH_FINISH = 8, PH_GAMMA = 16, PH_BORN = 32, PH_BORNFIN = 64,
             PH_NOFOLD = 128 /* the caller's exchange carries the status word itself (k_peer_allreduce_final) */ };

struct ForceSink { void* ptr; int layout; int padded_n; double* d_energy; };

cudaEvent_t prof_take(agbnp_b200* h) {
    if (h->prof_used == h->prof_pool.size()) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        h->prof_pool.push_back(e);
    }
    return h->prof_pool[h->prof_used++];
}

// one kernel of an evaluation.  With programmatic dependent launch the kernel may be scheduled while its predecessor in the
// stream is still draining; it blocks in pdl_acquire() until that grid has completed (agbnp_device.cuh).
template <typename Args>
void launch(agbnp_b200* h, void (*kern)(Args), int grid, int block, size_t smem, cudaStream_t s, const Args& a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned) grid); cfg.blockDim = dim3((unsigned) block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = h->use_pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kern, a));
}

// decide, once per evaluation and before anything is enqueued, whether its tree phase builds or rescans
void begin_eval(agbnp_b200* h) {
    const bool want = h->cfg.tree_reuse_interval > 1 && !h->tree_work_global && h->tree_built &&
                      h->evals_since_build < h->cfg.tree_reuse_interval;
    h->cur_eval_rescan = want;
    if (want) h->evals_since_build++;
    else { h->tree_built = true; h->evals_since_build = 1; }
}

// enqueue one evaluation (phase_mask: Phase bits).
// Nothing here synchronises; k_finish delivers to the caller's sink only if no capacity overflowed (device-side check).
void enqueue(agbnp_b200* h, const float4* d_posq_in, cudaStream_t s, int phase_mask, const ForceSink* sink) {
    const bool cutoff = h->cfg.nonbonded_method == AGBNP_B200_CUTOFF_NONPERIODIC;
    const bool v1 = h->cfg.version == 1;
    // Verlet lists (pair masks, level-2 candidate lists): rebuild when an atom has moved more than 0.49 skin since the build
    const float pq_move2 = h->pq_skin > 0.f ? (0.49f*h->pq_skin)*(0.49f*h->pq_skin) : -1.f;
    size_t open_a = 0;
    auto begin = [&](int id) {
        if (h->prof_mask & (1u << id)) { cudaEvent_t e = prof_take(h); open_a = h->prof_used-1; CK(cudaEventRecord(e, s)); }
    };
    auto end = [&](int id, bool kernel = true) {
        if (kernel) h->launches++;
        if (h->prof_mask & (1u << id)) {
            cudaEvent_t e = prof_take(h);
            CK(cudaEventRecord(e, s));
            h->prof_recs.push_back({id, open_a, h->prof_used-1});
        }
    };
    PairCommon pc = pair_common(h);
    if (phase_mask & PH_TREE) {
        const bool rescan = h->cur_eval_rescan;
        // a rescan evaluation keeps root_cnt (the tail of the slab): it says which stored subtrees exist
        PrepArgs pa{h->np, d_posq_in, h->d_orig.p, h->d_charge.p, h->d_posq.p, h->d_bbc.p, h->d_bbh.p,
                    (float4*) h->d_slab.p, (int) ((rescan ? h->slab_keep_off : h->slab_bytes)/sizeof(float4)),
                    h->d_posq_ref.p, h->d_pq_ctl.p, (h->cur_io && h->io_set) ? h->d_io.p : nullptr, (h->cur_io && h->posq_f64) ? 1 : 0};
        begin(K_PREP);
        launch(h, k_prep, (h->nb+7)/8, 256, 0, s, pa);
        end(K_PREP);
        if (rescan) {
            RescanArgs ra{};
            ra.items = h->d_items.p; ra.nitems = (int) h->items.size(); ra.posq = h->d_posq.p;
            ra.aL = h->d_aL.p; ra.vL = h->d_vL.p; ra.aS = h->d_aS.p; ra.vS = h->d_vS.p; ra.gamma = h->d_gamma.p;
            ra.volmina = h->k.volmina; ra.volminb = h->k.volminb; ra.swd = 1.0/(h->k.volminb-h->k.volmina);
            ra.cap = h->tree_cap; ra.stage = h->d_tree_stage.p; ra.stage_stride = tree_stage_bytes(h->tree_cap, h->tree_wcap);
            ra.accL = h->d_accL; ra.accS = h->d_accS; ra.scalars = h->d_scalars; ra.counters = h->d_counters;
            ra.st = h->st; ra.st.cursor = h->d_ctrl+CW_TREE_CURSOR;
            ra.work_counter = h->d_ctrl+CW_WORK_TREE; ra.status = h->d_ctrl+CW_STATUS; ra.tree_ok = h->d_tree_ok.p;
            begin(K_TREE);
            launch(h, k_tree_rescan, h->tree_grid, 32*h->tree_warps, h->tree_warps*rescan_work_bytes(h->tree_cap), s, ra);
            end(K_TREE);
        } else {
        // Block lists shorten the level-2 SEARCH of a root from all heavy blocks to ~10.  Searches are rare (the candidate lists
        // are kept between evaluations), and below ~65 k heavy atoms a root scans every block's box in 64 trips or fewer, which
        // costs k_tree less (+3 % of a search evaluation on HIV-RT) than every evaluation pays for one more kernel in the chain
        // (4-6 us): the kernel only runs for larger systems.
        const bool use_blist = h->nhb > BLIST_MIN_BLOCKS;
        if (use_blist) {
            BlockListArgs bl{h->nhb, h->d_bbc.p, h->d_bbh.p, h->rc2_global, h->d_bcount.p, h->d_blist.p, h->d_pq_ctl.p, pq_move2};
            begin(K_BLIST);
            launch(h, k_blocklist, (h->nhb+7)/8, 256, 0, s, bl);
            end(K_BLIST);
        }
        TreeArgs ta{};
        ta.nh = h->nh; ta.nhb = h->nhb; ta.np = h->np;
        ta.items = h->d_items.p; ta.nitems = (int) h->items.size(); ta.bcount = use_blist ? h->d_bcount.p : nullptr; ta.blist = h->d_blist.p;
        ta.item_roots = h->d_item_roots.p;
        ta.posq = h->d_posq.p; ta.orig = h->d_orig.p; ta.l2rec = h->d_l2rec.p; ta.rcbin = h->d_rcbin.p;
        ta.aL = h->d_aL.p; ta.vL = h->d_vL.p; ta.aS = h->d_aS.p; ta.vS = h->d_vS.p; ta.gamma = h->d_gamma.p;
        ta.bbc = h->d_bbc.p; ta.bbh = h->d_bbh.p; ta.rc2 = h->d_rc2.p; ta.rc2max = h->d_rc2max.p; ta.nbins = h->sp.nbins;
        ta.rc2s = h->d_rc2s.p; ta.rc2maxs = h->d_rc2maxs.p; ta.l2list = h->d_l2list.p; ta.l2cnt = h->d_l2cnt.p;
        ta.ctl = h->d_pq_ctl.p; ta.move2 = pq_move2;
        ta.volmina = h->k.volmina; ta.volminb = h->k.volminb; ta.min_gvol = h->k.min_gvol;
        ta.swd = 1.0/(h->k.volminb-h->k.volmina);
        // FP32 screen: |relative error| of the float overlap volume is < 1e-4 (positions within a subtree are < 2 nm from the
        // root: 1e-7 nm rounding, exponent argument < 40, ex2.approx 2 ulp); 1e-3 leaves a factor 10
        ta.screen = (float) (h->k.volmina*(1.0 - 1.0e-3));
        ta.max_order = h->k.max_order;
        ta.cap = h->tree_cap; ta.wcap = h->tree_wcap; ta.nbrmax = h->nbrmax;
        ta.stage = h->d_tree_stage.p; ta.stage_stride = tree_stage_bytes(h->tree_cap, h->tree_wcap);
        ta.wk_stride = tree_work_bytes(h->nbrmax, h->tree_cap, h->tree_wcap);
        ta.wk_global = h->tree_work_global ? h->d_tree_work.p : nullptr;
        ta.accL = h->d_accL; ta.accS = h->d_accS; ta.scalars = h->d_scalars; ta.counters = h->d_counters;
        ta.st = h->st; ta.st.cursor = h->d_ctrl+CW_TREE_CURSOR;
        ta.work_counter = h->d_ctrl+CW_WORK_TREE; ta.status = h->d_ctrl+CW_STATUS;
        ta.shard_rank = h->cfg.shard_rank; ta.shard_count = h->cfg.shard_count;
        ta.hw_nbr = h->d_ctrl+CW_MAX_NBR; ta.hw_nodes = h->d_ctrl+CW_MAX_NODES; ta.hw_width = h->d_ctrl+CW_MAX_WIDTH;
        const size_t smem = h->tree_work_global ? 0 : h->tree_warps*ta.wk_stride;
        begin(K_TREE);
        if (h->tree_work_global) launch(h, k_tree<false>, h->tree_grid, 32*h->tree_warps, 0, s, ta);
        else launch(h, k_tree<true>, h->tree_grid, 32*h->tree_warps, smem, s, ta);
        end(K_TREE);
        }
    }
    const size_t tab_bytes = pc.tab_smem ? (size_t) pc.ntables*I4_INTERVALS*sizeof(float4) : 0;
    // pair-mask reuse: list range = pass range + skin
    const float pq_range = cutoff ? (float) std::min(h->k.i4_maxa, h->cfg.cutoff) : (float) h->k.i4_maxa;
    const float pq_list2 = (pq_range + h->pq_skin)*(pq_range + h->pq_skin);
    if (v1 && (h->pq_shape_tab != tab_bytes || h->pq_shape_cutoff != (int) cutoff)) {         // launch shapes: once per table size / kernel variant
        if (pc.tab_smem) {
            if (cutoff) { pq_shape(k_born<true, true>, tab_bytes, BORN_WARP_SMEM, "AGBNP_B200_BORN_WARPS", h->born_w, h->born_c);
                          pq_shape(k_deriv<true, true>, tab_bytes, DERIV_WARP_SMEM, "AGBNP_B200_DERIV_WARPS", h->deriv_w, h->deriv_c); }
            else { pq_shape(k_born<false, true>, tab_bytes, BORN_WARP_SMEM, "AGBNP_B200_BORN_WARPS", h->born_w, h->born_c);
                   pq_shape(k_deriv<false, true>, tab_bytes, DERIV_WARP_SMEM, "AGBNP_B200_DERIV_WARPS", h->deriv_w, h->deriv_c); }
        } else {
            if (cutoff) { pq_shape(k_born<true, false>, tab_bytes, BORN_WARP_SMEM, "AGBNP_B200_BORN_WARPS", h->born_w, h->born_c);
                          pq_shape(k_deriv<true, false>, tab_bytes, DERIV_WARP_SMEM, "AGBNP_B200_DERIV_WARPS", h->deriv_w, h->deriv_c); }
            else { pq_shape(k_born<false, false>, tab_bytes, BORN_WARP_SMEM, "AGBNP_B200_BORN_WARPS", h->born_w, h->born_c);
                   pq_shape(k_deriv<false, false>, tab_bytes, DERIV_WARP_SMEM, "AGBNP_B200_DERIV_WARPS", h->deriv_w, h->deriv_c); }
        }
        h->pq_shape_tab = tab_bytes; h->pq_shape_cutoff = (int) cutoff;
    }
    if (v1 && (phase_mask & PH_BORN)) {
        BornArgs ba{};
        ba.c = pc;
        ba.u = PairUnits{h->d_pq_units.p, h->npq_units, h->d_pq_toff.p, h->d_pq_hits.p, h->d_pq_masks.p,
                         h->d_ctrl+CW_WORK_BORN, h->cfg.shard_rank, h->cfg.shard_count, h->d_pq_ctl.p, pq_list2, pq_move2};
        ba.accS = h->d_accS; ba.vS = h->d_vS.p; ba.bsum = h->d_bsum; ba.counters = h->d_counters;
        const int bw = h->born_w, bc = h->born_c;
        const size_t sm = tab_bytes + bw*BORN_WARP_SMEM;
        const int bgrid = h->num_sm*bc;                                                 // resident CTAs only (first_unit)
        begin(K_BORN);
        if (pc.tab_smem) { if (cutoff) launch(h, k_born<true, true>, bgrid, 32*bw, sm, s, ba); else launch(h, k_born<false, true>, bgrid, 32*bw, sm, s, ba); }
        else { if (cutoff) launch(h, k_born<true, false>, bgrid, 32*bw, sm, s, ba); else launch(h, k_born<false, false>, bgrid, 32*bw, sm, s, ba); }
        end(K_BORN);
    }
    if (v1 && (phase_mask & PH_BORNFIN)) {
        BornFinishArgs bf{};
        bf.np = h->np; bf.posq = h->d_posq.p; bf.orig = h->d_orig.p; bf.bsum = h->d_bsum; bf.accS = h->d_accS; bf.vS = h->d_vS.p;
        bf.radius = h->d_radius.p; bf.alpha = h->d_alpha.p;
        bf.vsf = h->d_vsf.p; bf.born = h->d_born.p; bf.bfp = h->d_bfp.p; bf.brw = h->d_brw.p; bf.gbj = h->d_gbj.p; bf.bmax = h->d_bmax.p;
        bf.kdiel = (float) h->k.dielectric_factor; bf.hb_radius = (float) h->k.hb_radius;
        bf.scalars = h->d_scalars; bf.own_begin = pc.row_begin*TILE; bf.own_end = pc.row_end*TILE;
        begin(K_BORNFIN);
        launch(h, k_born_finish, (h->np+255)/256, 256, 0, s, bf);
        end(K_BORNFIN);
    }
    if (v1 && (phase_mask & PH_GB)) {
        GBArgs ga{};
        ga.c = pc; ga.gbj = h->d_gbj.p; ga.units = h->d_units.p; ga.nunits = h->nunits; ga.chunk = h->gb_chunk;
        ga.shard_rank = h->cfg.shard_rank; ga.shard_count = h->cfg.shard_count;
        ga.gbacc = h->d_gbacc; ga.scalars = h->d_scalars; ga.counters = h->d_counters; ga.kdiel = h->k.dielectric_factor; ga.bmax = h->d_bmax.p;
        ga.work_counter = h->d_ctrl+CW_WORK_GB;
        begin(K_GB);
        if (cutoff) launch(h, k_gb<true>, h->gb_grid, GB_THREADS, 0, s, ga);
        else launch(h, k_gb<false>, h->gb_grid, GB_THREADS, 0, s, ga);
        end(K_GB);
    }
    if (v1 && (phase_mask & PH_DERIV)) {
        DerivArgs da{};
        da.c = pc;
        da.u = PairUnits{h->d_pq_units.p, h->npq_units, h->d_pq_toff.p, h->d_pq_hits.p, h->d_pq_masks.p,
                         h->d_ctrl+CW_WORK_DERIV, h->cfg.shard_rank, h->cfg.shard_count, h->d_pq_ctl.p, pq_list2, pq_move2};
        da.vsf = h->d_vsf.p; da.gbacc = h->d_gbacc; da.born = h->d_born.p; da.bfp = h->d_bfp.p; da.brw = h->d_brw.p;
        da.kdiel = (float) h->k.dielectric_factor; da.dacc = h->d_dacc;
        const int dw = h->deriv_w, dc = h->deriv_c;          // set with k_born's shape above
        const size_t sm = tab_bytes + dw*DERIV_WARP_SMEM;
        const int dgrid = h->num_sm*dc;
        begin(K_DERIV);
        if (pc.tab_smem) { if (cutoff) launch(h, k_deriv<true, true>, dgrid, 32*dw, sm, s, da); else launch(h, k_deriv<false, true>, dgrid, 32*dw, sm, s, da); }
        else { if (cutoff) launch(h, k_deriv<true, false>, dgrid, 32*dw, sm, s, da); else launch(h, k_deriv<false, false>, dgrid, 32*dw, sm, s, da); }
        end(K_DERIV);
    }
    if (v1 && (phase_mask & PH_GAMMA)) {
        GammaArgs gm{};
        gm.nitems = (int) h->items.size(); gm.np = h->np; gm.st = h->st; gm.dacc = h->d_dacc; gm.inv_vS = h->d_inv_vS.p;
        gm.gacc = h->d_gacc; gm.scratch_stride = gamma_work_bytes(h->tree_cap);
        gm.scratch = h->gamma_work_global ? h->d_gamma_scratch.p : nullptr;
        gm.cap = h->tree_cap; gm.work_counter = h->d_ctrl+CW_WORK_GAMMA;
        gm.shard_rank = h->cfg.shard_rank; gm.shard_count = h->cfg.shard_count;
        begin(K_GAMMA);
        if (h->gamma_work_global) launch(h, k_tree_gamma<false>, h->gamma_grid, 32*h->gamma_warps, 0, s, gm);
        else launch(h, k_tree_gamma<true>, h->gamma_grid, 32*h->gamma_warps, h->gamma_warps*gm.scratch_stride, s, gm);
        end(K_GAMMA);
    }
    if ((phase_mask & PH_GAMMA) && !(phase_mask & PH_NOFOLD) && h->cfg.shard_count > 1) {
        // the status word joins the energy scalars, so that the ENERGY exchange tells every shard whether ANY shard overflowed
        StatusFoldArgs sf{h->d_ctrl+CW_STATUS, h->d_scalars};
        launch(h, k_status_fold, 1, 32, 0, s, sf);
        h->launches++;
    }
    if (phase_mask & PH_FINISH) {
        FinishArgs fa{};
        fa.np = h->np; fa.n = h->n; fa.orig = h->d_orig.p; fa.accL = h->d_accL; fa.accS = h->d_accS; fa.scalars = h->d_scalars;
        fa.inv_roffset = (float) (1.0/h->k.roffset);
        fa.status = h->d_ctrl+CW_STATUS;
        fa.sharded = h->cfg.shard_count > 1;
        fa.peer_fault = h->peer_ready ? h->peer.status : nullptr;
        fa.tree_ok_out = h->cur_eval_rescan ? nullptr : h->d_tree_ok.p;     // a build evaluation (in)validates the stored tree
        if (v1) { fa.gbacc = h->d_gbacc; fa.dacc = h->d_dacc; fa.gacc = h->d_gacc; fa.gb_scale = -2.0*h->k.dielectric_factor; }
        fa.posq = h->d_posq.p; fa.posq_ref = h->d_posq_ref.p; fa.pq_ctl = h->d_pq_ctl.p;
        fa.padded_n = sink ? sink->padded_n : 0;
        if (sink && sink->ptr) {
            if (sink->layout == 0) fa.out_f32 = (float*) sink->ptr;
            else if (sink->layout == 1) fa.out_fixed = (unsigned long long*) sink->ptr;
            else fa.out_set = (float*) sink->ptr;
        }
        fa.energy_accum = sink ? sink->d_energy : nullptr;
        fa.io = (h->cur_io && h->io_set) ? h->d_io.p : nullptr;
        fa.energy_f32 = (h->cur_io && h->energy_f32) ? 1 : 0;
        // the host-buffer call (sink layout 2) reads status and energy from a pinned mirror that k_finish fills itself
        if (sink && sink->layout == 2) { fa.tail_out = (int*) h->h_tail; fa.tail_words = (int) ((512 + sizeof(int)*CW_COUNT)/sizeof(int)); }
        begin(K_FINISH);
        launch(h, k_finish, (h->np+255)/256, 256, 0, s, fa);
        end(K_FINISH);
    }
    CK(cudaGetLastError());
}

constexpr int PH_ALL = PH_TREE|PH_BORN|PH_BORNFIN|PH_GB|PH_DERIV|PH_GAMMA|PH_FINISH;
// an evaluation whose caller wants no forces (includeForces == false, or no force sink) needs neither the Born-radius
// derivative pass nor the gamma sweep: they produce forces only (the Reference platform computes them regardless)
constexpr int PH_ENERGY_ONLY = PH_TREE|PH_BORN|PH_BORNFIN|PH_GB|PH_FINISH;
inline int phases_for(const ForceSink* sink) { return (sink && !sink->ptr) ? PH_ENERGY_ONLY : PH_ALL; }

// one whole evaluation on stream s: a cached CUDA graph of the kernel sequence (one launch instead of eight; the capture
// happens on the handle's own stream because the caller's may be the legacy default stream, which cannot be captured)
void peer_enqueue(agbnp_b200* h, int which, cudaStream_t s) {
    void* ptr = nullptr; size_t bytes = 0;
    if (agbnp_b200_shard_buffer(h, which, &ptr, &bytes) != AGBNP_B200_OK) throw CudaFail{h->err};
    if (which == AGBNP_B200_BUF_ENERGY) {
        k_peer_allreduce<double><<<1, 256, 0, s>>>((double*) ptr, bytes/sizeof(double), h->peer, which);
    } else {
        const size_t nvec = bytes/sizeof(float4);       // np is a multiple of 32: every buffer is whole float4s
        const int grid = (int) std::min<size_t>(64, (nvec+255)/256);
        k_peer_allreduce<float4><<<grid, 256, 0, s>>>((float4*) ptr, nvec, h->peer, which);
    }
    h->launches += 1;
    CK(cudaGetLastError());
}


// a whole sharded evaluation: the phases with the peer-memory exchange after each
void enqueue_sharded(agbnp_b200* h, const float4* d_posq_in, cudaStream_t s, const ForceSink* sink) {
    enqueue(h, d_posq_in, s, PH_TREE, nullptr);            peer_enqueue(h, AGBNP_B200_BUF_SELFVOL, s);
    enqueue(h, d_posq_in, s, PH_BORN, nullptr);            peer_enqueue(h, AGBNP_B200_BUF_BSUM, s);
    enqueue(h, d_posq_in, s, PH_BORNFIN|PH_GB, nullptr);   peer_enqueue(h, AGBNP_B200_BUF_YQ, s);
    enqueue(h, d_posq_in, s, PH_DERIV, nullptr);           peer_enqueue(h, AGBNP_B200_BUF_WU, s);
    static const bool split_final = std::getenv("AGBNP_B200_PEER_SPLIT_FINAL") != nullptr;     // diagnostics: the three exchanges one by one
    if (split_final) {
        enqueue(h, d_posq_in, s, PH_GAMMA, nullptr);       peer_enqueue(h, AGBNP_B200_BUF_FORCE, s);
        peer_enqueue(h, AGBNP_B200_BUF_ENERGY, s);
        enqueue(h, d_posq_in, s, PH_FINISH, sink);
        return;
    }
    enqueue(h, d_posq_in, s, PH_GAMMA|PH_NOFOLD, nullptr);
    {   // forces of the gamma sweep + energy scalars + status word in one exchange
        const size_t nvec = (size_t) h->np;
        const int grid = (int) std::min<size_t>(64, (nvec+255)/256);
        k_peer_allreduce_final<<<grid, 256, 0, s>>>(h->d_gacc, nvec, h->d_scalars, h->d_ctrl+CW_STATUS, h->peer);
        h->launches += 1;
        CK(cudaGetLastError());
    }
    enqueue(h, d_posq_in, s, PH_FINISH, sink);
}

void launch_all(agbnp_b200* h, const float4* d_posq_in, cudaStream_t s, const ForceSink* sink, bool sharded = false) {
    begin_eval(h);
    // a sharded evaluation is bounded by the waits of its exchange kernels, not by launch latency: measured on 2 and 8 B200 a
    // graph of it is no faster than the plain launches (490 vs 466 us, 372 vs 367 us), so it is launched directly
    if (!h->use_graph || h->prof_mask || sharded) {
        if (sharded) enqueue_sharded(h, d_posq_in, s, sink); else enqueue(h, d_posq_in, s, phases_for(sink), sink);
        mark_tail(h, s);
        return;
    }
    agbnp_b200::GraphEntry* hit = nullptr;
    for (auto& g : h->graphs)
        if (g.gen == h->launch_gen && g.posq == d_posq_in && g.sink == sink->ptr && g.layout == sink->layout &&
            g.padded_n == sink->padded_n && g.d_energy == sink->d_energy && g.sharded == sharded && g.rescan == h->cur_eval_rescan) { hit = &g; break; }
    if (!hit) {
        // drop graphs of an older configuration, and the least recently used one beyond 16
        for (size_t i = 0; i < h->graphs.size(); ) {
            if (h->graphs[i].gen != h->launch_gen) { cudaGraphExecDestroy(h->graphs[i].exec); h->graphs.erase(h->graphs.begin()+i); }
            else i++;
        }
        if (h->graphs.size() >= 16) {
            size_t lru = 0;
            for (size_t i = 1; i < h->graphs.size(); i++) if (h->graphs[i].last_use < h->graphs[lru].last_use) lru = i;
            cudaGraphExecDestroy(h->graphs[lru].exec);
            h->graphs.erase(h->graphs.begin()+lru);
        }
        const long long before = h->launches;
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(h->own_stream, cudaStreamCaptureModeThreadLocal));
        try { if (sharded) enqueue_sharded(h, d_posq_in, h->own_stream, sink); else enqueue(h, d_posq_in, h->own_stream, phases_for(sink), sink); }
        catch (...) { cudaStreamEndCapture(h->own_stream, &graph); if (graph) cudaGraphDestroy(graph); throw; }
        CK(cudaStreamEndCapture(h->own_stream, &graph));
        agbnp_b200::GraphEntry e{h->launch_gen, d_posq_in, sink->ptr, sink->layout, sink->padded_n, sink->d_energy, sharded, h->cur_eval_rescan, nullptr,
                                 (int) (h->launches-before), 0};
        h->launches = before;
        h->n_graph_inst++;
        const cudaError_t ce = cudaGraphInstantiate(&e.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) throw CudaFail{std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce)};
        h->graphs.push_back(e);
        hit = &h->graphs.back();
    }
    hit->last_use = ++h->graph_clock;
    CK(cudaGraphLaunch(hit->exec, s));
    h->launches += hit->kernels;
    mark_tail(h, s);
}

// read back status + scalars (synchronises the stream); returns the status bits
int fetch_status(agbnp_b200* h, cudaStream_t s) {
    // scalars, counters and control words are three consecutive 256-byte slots of the slab (upload_static)
    CK(cudaMemcpyAsync(h->h_tail, h->d_scalars, 512 + sizeof(int)*CW_COUNT, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return h->h_ctrl[CW_STATUS];
}

// grow whatever overflowed; returns false if a limit was hit.  ctrl = the evaluation's control words (host copy).
// A real overflow reports a high-water mark above the capacity it was enqueued with; if the capacity has been grown past
// that mark since (asynchronous evaluations that were already in flight when an earlier one faulted), nothing is grown
// again: one fault, one growth.
bool grow(agbnp_b200* h, const int* ctrl, bool ahead = false) {
    const int status = ctrl[CW_STATUS];
    if (status & ST_PEER_TIMEOUT) return false;        // not a capacity problem: a peer never arrived (sticky)
    const bool g_nbr = (status & ST_NBR_OVERFLOW) && (ahead || ctrl[CW_MAX_NBR] > h->nbrmax);
    const bool g_node = (status & ST_NODE_OVERFLOW) && (ahead || ctrl[CW_MAX_NODES] > h->tree_cap);
    const bool g_level = (status & ST_LEVEL_OVERFLOW) && (ahead || ctrl[CW_MAX_WIDTH] > h->tree_wcap);
    const bool g_store = (status & ST_TREE_OVERFLOW) && (ahead || ctrl[CW_TREE_CURSOR] > h->st.cap);
    if (!(g_nbr || g_node || g_level || g_store)) return true;
    if (g_nbr && h->nbrmax >= 1024) return false;
    if (g_node && h->tree_cap >= 16384) return false;
    if (g_level && h->tree_wcap >= 16384) return false;
    wait_own_work(h);                                  // buffers below may still be in use by queued evaluations
    TrashScope ts(&h->trash);
    h->n_grow++;
    // a real overflow doubles (its high-water mark is only a lower bound); growing ahead of need takes small steps so that
    // the work arrays keep fitting in shared memory
    auto bump = [ahead](int v, int seen) { return (std::max(ahead ? v + v/8 : 2*v, seen + seen/8) + 31)/32*32; };
    if (g_nbr) h->nbrmax = std::min(1024, bump(h->nbrmax, ctrl[CW_MAX_NBR]));
    if (g_node) h->tree_cap = std::min(16384, bump(h->tree_cap, ctrl[CW_MAX_NODES]));
    if (g_level) h->tree_wcap = std::min(16384, bump(h->tree_wcap, ctrl[CW_MAX_WIDTH]));
    h->tree_wcap = std::min(h->tree_wcap, h->tree_cap);
    if (g_nbr || g_node || g_level) alloc_tree_scratch(h);
    if (g_store) {
        const int need = ctrl[CW_TREE_CURSOR];
        alloc_store(h, std::max(need + need/4 + 4096, h->st.cap*2));
    }
    return true;
}

// keep headroom so that an asynchronous evaluation practically never overflows: grow when a high-water mark of a
// completed evaluation exceeds 3/4 of its capacity
void grow_ahead(agbnp_b200* h, const int* ctrl) {
    int fake[CW_COUNT] = {};
    fake[CW_MAX_NBR] = ctrl[CW_MAX_NBR]; fake[CW_MAX_NODES] = ctrl[CW_MAX_NODES]; fake[CW_MAX_WIDTH] = ctrl[CW_MAX_WIDTH];
    if (ctrl[CW_MAX_NBR]*10 > h->nbrmax*9 && h->nbrmax < 1024) fake[CW_STATUS] |= ST_NBR_OVERFLOW;
    if (ctrl[CW_MAX_NODES]*10 > h->tree_cap*9 && h->tree_cap < 16384) fake[CW_STATUS] |= ST_NODE_OVERFLOW;
    if (ctrl[CW_MAX_WIDTH]*10 > h->tree_wcap*9 && h->tree_wcap < h->tree_cap) fake[CW_STATUS] |= ST_LEVEL_OVERFLOW;
    if ((long long) ctrl[CW_TREE_CURSOR]*8 > (long long) h->st.cap*7) { fake[CW_STATUS] |= ST_TREE_OVERFLOW; fake[CW_TREE_CURSOR] = ctrl[CW_TREE_CURSOR]; }
    // applied at the start of the next evaluation (prepare): the buffers still hold this evaluation's by-products
    // (agbnp_b200_get reads the stored tree)
    if (fake[CW_STATUS]) { std::memcpy(h->ahead, fake, sizeof(fake)); h->ahead_pending = true; }
}

void prepare(agbnp_b200* h, const float* host_xyz, int stride, const void* d_posq_in, cudaStream_t s) {
    if (h->ahead_pending) { h->ahead_pending = false; grow(h, h->ahead, true); }
    static const bool timing = std::getenv("AGBNP_B200_HOST_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tp0 = timing ? now() : 0;
    double tp1 = tp0;
    const int interval = h->cfg.reorder_interval > 0 ? h->cfg.reorder_interval : 1000;
    if (!h->order_valid || h->evals_since_sort >= interval) {
        std::vector<float> tmp;
        if (!host_xyz) {
            // positions from the caller's device buffer, in its layout (agbnp_b200_set_device_layout), to particle order
            tmp.resize((size_t) 4*h->n);
            if (h->posq_f64) {
                std::vector<double> t64((size_t) 4*h->n);
                CK(cudaMemcpyAsync(t64.data(), d_posq_in, sizeof(double)*4*h->n, cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                for (int o = 0; o < h->n; o++) for (int c = 0; c < 4; c++) tmp[(size_t) 4*o+c] = (float) t64[(size_t) 4*(h->io_set ? h->io[o] : o)+c];
            } else {
                CK(cudaMemcpyAsync(tmp.data(), d_posq_in, sizeof(float4)*h->n, cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                if (h->io_set) {
                    std::vector<float> t2(tmp.size());
                    for (int o = 0; o < h->n; o++) for (int c = 0; c < 4; c++) t2[(size_t) 4*o+c] = tmp[(size_t) 4*h->io[o]+c];
                    tmp.swap(t2);
                }
            }
            host_xyz = tmp.data(); stride = 4;
        }
        wait_own_work(h);                              // queued evaluations still read the arrays about to be replaced
        build_order(h, host_xyz, stride, s);
        if (timing) tp1 = now();
    }
    if (h->params_dirty) {
        wait_own_work(h);
        TrashScope ts(&h->trash);
        upload_static(h, s);
        if (timing) std::fprintf(stderr, "prepare: build_order %.2f ms, upload_static %.2f ms\n", tp1-tp0, now()-tp1);
    }
}

// synchronous evaluation: everything is enqueued at once (the finish kernel checks the status word on the device), one
// synchronisation reads status + energy; on overflow the capacities are grown and the evaluation re-runs, so the
// caller's sink only ever receives forces of a complete evaluation.
int run_checked(agbnp_b200* h, const float4* d_posq_in, cudaStream_t s, const ForceSink* sink) {
    for (int attempt = 0; attempt < 8; attempt++) {
        launch_all(h, d_posq_in, s, sink);
        const int status = fetch_status(h, s);
        if (status == 0) {
            h->evals_since_sort++; h->total_evals++;
            grow_ahead(h, h->h_ctrl);
            return AGBNP_B200_OK;
        }
        h->tree_built = false;
        if (!grow(h, h->h_ctrl)) {
            h->err = (status & ST_PEER_TIMEOUT) ? "agbnp_b200: a peer-memory exchange timed out waiting for another shard; this handle no longer delivers results"
                                                : "agbnp_b200: internal capacity limit exceeded (status " + std::to_string(status) + ")";
            return AGBNP_B200_ERR_CAPACITY;
        }
    }
    h->err = "agbnp_b200: capacity growth did not converge";
    return AGBNP_B200_ERR_CAPACITY;
}

// retire the deferred status of asynchronous evaluation `k` (waits for it if still running)
int async_retire(agbnp_b200* h, long long k) {
    const int slot = (int) (k % agbnp_b200::ASYNC_DEPTH);
    if (!h->async_pending[slot]) return AGBNP_B200_OK;
    CK(cudaEventSynchronize(h->async_ev[slot]));
    h->async_pending[slot] = false;
    const int* ctrl = h->h_async + slot*CW_COUNT;
    if (ctrl[CW_STATUS] != 0) {
        h->tree_built = false;
        const bool ok = grow(h, ctrl);
        h->async_fault = true;
        h->n_async_fault++;
        const bool peer_only = (ctrl[CW_STATUS] & ~ST_PEER_OVERFLOW) == 0;
        h->err = std::string("agbnp_b200: asynchronous evaluation ") + std::to_string(k) + (peer_only ? " overflowed an internal capacity on another shard" :
               " overflowed an internal capacity") + " (status " + std::to_string(ctrl[CW_STATUS]) + "); its forces and energy were NOT delivered"
               + (h->cfg.shard_count > 1 ? " on any shard" : "") + (ok ? "; capacities grown, re-issue it" : "; a capacity limit was reached")
               + ".  Evaluations enqueued since (including the one issued by the call that returned this) ran with the old capacities and are reported separately";
        return AGBNP_B200_ERR_CAPACITY;
    }
    grow_ahead(h, ctrl);
    return AGBNP_B200_OK;
}

// Book-keeping of an asynchronous evaluation, called right AFTER it has been enqueued on `s`: its status words follow it to
// pinned memory (ring slot k % ASYNC_DEPTH), then the evaluation issued ASYNC_DEPTH-1 calls ago is retired and ITS outcome
// is what the call returns.  Enqueue-then-check matters for sharded evaluations: every shard has enqueued the full
// collective sequence of evaluation k before it looks at anything, so a fault on one shard can neither leave the peers'
// exchange kernels waiting for flags that never come nor let the shards' evaluation counters (epochs, re-sort schedule)
// drift apart; and since the status words are exchanged (SC_FAULT), all shards see the same fault at the same call.
int async_post(agbnp_b200* h, cudaStream_t s) {
    const long long k = h->async_issued;
    const int slot = (int) (k % agbnp_b200::ASYNC_DEPTH);
    if (h->async_pending[slot]) {                    // cannot happen in steady state (retired by the previous call)
        const int rc = async_retire(h, k-agbnp_b200::ASYNC_DEPTH);
        if (rc != AGBNP_B200_OK) h->deferred_rc = rc;
    }
    CK(cudaMemcpyAsync(h->h_async + slot*CW_COUNT, h->d_ctrl, sizeof(int)*CW_COUNT, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(h->async_ev[slot], s));
    h->async_pending[slot] = true;
    h->async_issued++;
    h->evals_since_sort++; h->total_evals++;
    int rc = AGBNP_B200_OK;
    if (k >= agbnp_b200::ASYNC_DEPTH-1) rc = async_retire(h, k-(agbnp_b200::ASYNC_DEPTH-1));
    if (rc == AGBNP_B200_OK && h->deferred_rc != AGBNP_B200_OK) { rc = h->deferred_rc; }
    h->deferred_rc = AGBNP_B200_OK;
    return rc;
}

// asynchronous evaluation: returns after enqueueing.  The status words travel to pinned memory behind the kernels and
// are checked ASYNC_DEPTH-1 evaluations later (or by agbnp_b200_synchronize).
int run_async(agbnp_b200* h, const float4* d_posq_in, cudaStream_t s, const ForceSink* sink) {
    launch_all(h, d_posq_in, s, sink);
    return async_post(h, s);
}

int async_drain(agbnp_b200* h) {
    int rc = AGBNP_B200_OK;
    for (long long k = std::max(0ll, h->async_issued-agbnp_b200::ASYNC_DEPTH); k < h->async_issued; k++) {
        const int r = async_retire(h, k);
        if (r != AGBNP_B200_OK) rc = r;
    }
    return rc;
}

} // namespace

extern "C" {

const char* agbnp_b200_version(void) { return "agbnp_b200 0.1 sm_100a"; }

void agbnp_b200_default_config(agbnp_b200_config* cfg) {
    cfg->version = 1; cfg->nonbonded_method = AGBNP_B200_NOCUTOFF; cfg->cutoff = 1.0; cfg->device = 0;
    cfg->shard_rank = 0; cfg->shard_count = 1; cfg->reorder_interval = 0; cfg->tree_reuse_interval = 0;
}

const char* agbnp_b200_last_error(const agbnp_b200* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int agbnp_b200_create(const agbnp_b200_config* cfg, int n, const double* radius, const double* gamma, const double* alpha,
                      const double* charge, const unsigned char* ishydrogen, agbnp_b200** out) {
    if (out) *out = nullptr;
    if (!cfg || !out || n <= 0 || !radius || !gamma || !alpha || !charge || !ishydrogen) {
        g_create_error = "agbnp_b200_create: null or empty argument"; return AGBNP_B200_ERR_ARG;
    }
    if (cfg->version < 0 || cfg->version > 2) { g_create_error = "AGBNPForce::setVersion(): illegal version number"; return AGBNP_B200_ERR_ARG; }
    if (cfg->version == 2) { g_create_error = "agbnp_b200: AGBNP2 (version 2) is outside this library's path"; return AGBNP_B200_ERR_ARG; }
    if (cfg->nonbonded_method == AGBNP_B200_CUTOFF_PERIODIC) {
        g_create_error = "agbnp_b200: CutoffPeriodic is not supported (the reference implements it on no platform)"; return AGBNP_B200_ERR_ARG;
    }
    if (cfg->nonbonded_method != AGBNP_B200_NOCUTOFF && cfg->nonbonded_method != AGBNP_B200_CUTOFF_NONPERIODIC) {
        g_create_error = "agbnp_b200: unknown nonbonded method"; return AGBNP_B200_ERR_ARG;
    }
    if (cfg->shard_count < 1 || cfg->shard_rank < 0 || cfg->shard_rank >= cfg->shard_count) {
        g_create_error = "agbnp_b200: bad shard rank/count"; return AGBNP_B200_ERR_ARG;
    }
    if (n >= (1 << 24)) { g_create_error = "agbnp_b200: more than 2^24 particles are not supported"; return AGBNP_B200_ERR_ARG; }
    if (cfg->nonbonded_method == AGBNP_B200_CUTOFF_NONPERIODIC && !(cfg->cutoff > 0)) {
        g_create_error = "agbnp_b200: cutoff must be positive"; return AGBNP_B200_ERR_ARG;
    }
    agbnp_b200* h = new agbnp_b200();
    h->cfg = *cfg;
    if (h->cfg.tree_reuse_interval <= 0) {      // opt-in without touching the caller's code (the C++ host layer, bench.py)
        const char* tr = std::getenv("AGBNP_B200_TREE_REUSE");
        if (tr) h->cfg.tree_reuse_interval = std::atoi(tr);
    }
    h->k = Constants::make();
    h->n = n;
    std::string e = h->sp.init(cfg->version, n, radius, gamma, alpha, charge, ishydrogen, h->k);
    if (!e.empty()) { g_create_error = e; delete h; return AGBNP_B200_ERR_ARG; }
    if (h->sp.i4.ntypes_screened > 255 || h->sp.i4.ntypes_screener > 127) {
        g_create_error = "agbnp_b200: too many distinct atomic radii (max 255 screened / 127 screener classes)"; delete h; return AGBNP_B200_ERR_ARG;
    }
    try {
        int ndev = 0;
        cudaError_t ce = cudaGetDeviceCount(&ndev);
        if (ce != cudaSuccess || ndev == 0) throw CudaFail{"no CUDA device available (this library has no CPU fallback)"};
        CK(cudaSetDevice(cfg->device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, cfg->device));
        h->num_sm = prop.multiProcessorCount;
        CK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
        preload_kernels();
        { const char* ng = std::getenv("AGBNP_B200_NO_GRAPH"); h->use_graph = !(ng && ng[0] == '1'); }
        { const char* ng = std::getenv("AGBNP_B200_NO_PDL"); h->use_pdl = !(ng && ng[0] == '1'); }
        if (const char* sk = std::getenv("AGBNP_B200_PAIR_SKIN")) h->pq_skin = std::max(0.f, (float) std::atof(sk));
        if (const char* tg = std::getenv("AGBNP_B200_TREE_GROUP")) h->tree_group = tg[0] != '0';
        for (auto& e2 : h->ev) CK(cudaEventCreate(&e2));
        for (auto& e2 : h->async_ev) CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->tail_ev, cudaEventDisableTiming));
        h->have_events = true;
        CK(cudaMallocHost((void**) &h->h_async, sizeof(int)*CW_COUNT*agbnp_b200::ASYNC_DEPTH));
        h->d_tree_ok.alloc(1);
        CK(cudaMemset(h->d_tree_ok.p, 0, sizeof(int)));
#ifndef GB_CTAS
#define GB_CTAS 4
#endif
        h->gb_grid = h->num_sm*GB_CTAS;
        const int pair_smem = (int) prop.sharedMemPerBlockOptin, deriv_smem = pair_smem;
        CK(cudaFuncSetAttribute(k_born<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair_smem));
        CK(cudaFuncSetAttribute(k_born<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair_smem));
        CK(cudaFuncSetAttribute(k_born<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair_smem));
        CK(cudaFuncSetAttribute(k_born<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair_smem));
        CK(cudaFuncSetAttribute(k_deriv<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, deriv_smem));
        CK(cudaFuncSetAttribute(k_deriv<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, deriv_smem));
        CK(cudaFuncSetAttribute(k_deriv<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, deriv_smem));
        CK(cudaFuncSetAttribute(k_deriv<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, deriv_smem));
        if (const char* ic = std::getenv("AGBNP_B200_INIT_CAPS")) {     // tests: start from small capacities to exercise the growth paths
            int a = 0, b = 0, c = 0;
            if (std::sscanf(ic, "%d,%d,%d", &a, &b, &c) == 3 && a >= 32 && b >= 32 && c >= 8) {
                h->tree_cap = a; h->tree_wcap = std::min(a, b); h->nbrmax = c;
            }
        }
        alloc_tree_scratch(h);
        CK(cudaMallocHost((void**) &h->h_posq, sizeof(float4)*n));
        CK(cudaMallocHost((void**) &h->h_force, sizeof(float)*3*n));
        CK(cudaMallocHost((void**) &h->h_tail, 512 + sizeof(int)*CW_COUNT));
        h->h_scal = (double*) h->h_tail; h->h_ctrl = (int*) (h->h_tail + 512);
        h->d_posq_in.alloc(n);
        h->d_force_out.alloc((size_t) 3*n);
    } catch (const CudaFail& f) {
        g_create_error = "agbnp_b200_create: " + f.msg;
        delete h;
        return AGBNP_B200_ERR_CUDA;
    }
    *out = h;
    return AGBNP_B200_OK;
}

void agbnp_b200_destroy(agbnp_b200* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    delete h;
}

int agbnp_b200_set_params(agbnp_b200* h, int n, const double* radius, const double* gamma, const double* alpha,
                          const double* charge, const unsigned char* ishydrogen) {
    if (!h) return AGBNP_B200_ERR_ARG;
    std::string e = h->sp.update(n, radius, gamma, alpha, charge, ishydrogen);
    if (!e.empty()) { h->err = e; return AGBNP_B200_ERR_PARAM_CHANGE; }
    h->params_dirty = true;
    return AGBNP_B200_OK;
}

int agbnp_b200_execute_host(agbnp_b200* h, const double* pos, int include_forces, int include_energy,
                            double* energy, double* forces) {
    if (!h || !pos) return AGBNP_B200_ERR_ARG;
    try {
        CK(cudaSetDevice(h->cfg.device));
        cudaStream_t s = h->own_stream;
        static const bool timing = std::getenv("AGBNP_B200_HOST_TIMING") != nullptr;
        static double acc[5] = {0, 0, 0, 0, 0};
        static long ncall = 0;
        auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double t0 = timing ? now() : 0;
        const int rc_deferred = async_drain(h);      // a fault of an earlier asynchronous evaluation is reported by this call
        std::string deferred_msg = h->err;
        h->cur_io = false;
        pack_positions(pos, (float*) h->h_posq, h->n);
        CK(cudaMemcpyAsync(h->d_posq_in.p, h->h_posq, sizeof(float4)*h->n, cudaMemcpyHostToDevice, s));
        const double t1 = timing ? now() : 0;
        prepare(h, (const float*) h->h_posq, 4, nullptr, s);
        const double t2 = timing ? now() : 0;
        double t3 = 0;
        const bool want_forces = include_forces && forces;
        ForceSink sink{want_forces ? h->d_force_out.p : nullptr, 2, h->n, nullptr};      // layout 2: float[3n], assigned; none: energy only
        for (int attempt = 0; ; attempt++) {
            launch_all(h, h->d_posq_in.p, s, &sink);
            if (include_forces && forces)
                CK(cudaMemcpyAsync(h->h_force, h->d_force_out.p, sizeof(float)*3*h->n, cudaMemcpyDeviceToHost, s));
            if (timing) t3 = now();
            CK(cudaStreamSynchronize(s));                   // the one synchronisation of the call; k_finish has filled h_tail
            const int status = h->h_ctrl[CW_STATUS];
            if (status == 0) break;
            h->tree_built = false;
            if (attempt >= 8 || !grow(h, h->h_ctrl)) {
                h->err = "agbnp_b200: internal capacity limit exceeded (status " + std::to_string(status) + ")";
                return AGBNP_B200_ERR_CAPACITY;
            }
        }
        h->evals_since_sort++; h->total_evals++;
        grow_ahead(h, h->h_ctrl);
        const double t4 = timing ? now() : 0;
        if (include_forces && forces) { if (include_forces == AGBNP_B200_FORCES_ASSIGN) set_forces(h->h_force, forces, 3*h->n); else add_forces(h->h_force, forces, 3*h->n); }
        if (energy) *energy = include_energy ? h->h_scal[SC_TOTAL] : 0.0;
        if (timing) {
            const double t5 = now();
            acc[0] += t1-t0; acc[1] += t2-t1; acc[2] += t3-t2; acc[3] += t4-t3; acc[4] += t5-t4;
            if (++ncall % 100 == 0) {
                std::fprintf(stderr, "execute_host us/call: pack+h2d enqueue %.1f  prepare %.1f  launch %.1f  wait %.1f  unpack %.1f\n",
                             acc[0]/100, acc[1]/100, acc[2]/100, acc[3]/100, acc[4]/100);
                for (double& a : acc) a = 0;
            }
        }
        if (rc_deferred != AGBNP_B200_OK) { h->err = deferred_msg; return rc_deferred; }
    } catch (const CudaFail& f) { h->err = f.msg; return AGBNP_B200_ERR_CUDA; }
    return AGBNP_B200_OK;
}

int agbnp_b200_execute_device(agbnp_b200* h, const void* d_posq, void* stream, void* d_force, int force_layout,
                              int padded_n, double* d_energy, double* h_energy) {
    if (!h || !d_posq) return AGBNP_B200_ERR_ARG;
    if (d_force && force_layout != 0 && force_layout != 1) { h->err = "agbnp_b200_execute_device: bad force layout"; return AGBNP_B200_ERR_ARG; }
    try {
        CK(cudaSetDevice(h->cfg.device));
        cudaStream_t s = (cudaStream_t) stream;
        h->cur_io = true;
        prepare(h, nullptr, 0, d_posq, s);
        ForceSink sink{d_force, force_layout, padded_n > 0 ? padded_n : h->n, d_energy};
        if (!h_energy) return run_async(h, (const float4*) d_posq, s, &sink);
        const int rc_deferred = async_drain(h);      // a fault of an earlier asynchronous evaluation is reported by this call
        const std::string deferred_msg = h->err;
        const int rc = run_checked(h, (const float4*) d_posq, s, &sink);
        if (rc != AGBNP_B200_OK) return rc;
        *h_energy = h->h_scal[SC_TOTAL];
        if (rc_deferred != AGBNP_B200_OK) { h->err = deferred_msg; return rc_deferred; }
    } catch (const CudaFail& f) { h->err = f.msg; return AGBNP_B200_ERR_CUDA; }
    return AGBNP_B200_OK;
}

int agbnp_b200_set_device_layout(agbnp_b200* h, const agbnp_b200_device_layout* lay) {
    if (!h) return AGBNP_B200_ERR_ARG;
    try {
        CK(cudaSetDevice(h->cfg.device));
        wait_own_work(h);                               // queued evaluations read the old index map
        TrashScope ts(&h->trash);
        h->launch_gen++;                                // cached graphs carry the old layout in their kernel arguments
        if (!lay || !lay->atom_index) { h->io_set = false; h->io.clear(); }
        else {
            std::vector<int> inv(h->n, -1);
            for (int p = 0; p < h->n; p++) {
                const int o = lay->atom_index[p];
                if (o < 0 || o >= h->n || inv[o] >= 0) { h->err = "agbnp_b200_set_device_layout: atom_index is not a permutation of 0..N-1"; return AGBNP_B200_ERR_ARG; }
                inv[o] = p;
            }
            h->io = inv;
            h->d_io.alloc(h->n);
            CK(cudaMemcpy(h->d_io.p, inv.data(), sizeof(int)*h->n, cudaMemcpyHostToDevice));
            h->io_set = true;
        }
        h->posq_f64 = lay && lay->posq_is_double;
        h->energy_f32 = lay && lay->energy_is_float;
    } catch (const CudaFail& f) { h->err = f.msg; return AGBNP_B200_ERR_CUDA; }
    return AGBNP_B200_OK;
}

int agbnp_b200_synchronize(agbnp_b200* h, void* stream) {
    if (!h) return AGBNP_B200_ERR_ARG;
    try {
        CK(cudaSetDevice(h->cfg.device));
        CK(cudaStreamSynchronize((cudaStream_t) stream));
        return async_drain(h);
    } catch (const CudaFail& f) { h->err = f.msg; return AGBNP_B200_ERR_CUDA; }
}

int agbnp_b200_profile(agbnp_b200* h, unsigned kernel_mask) {
    if (!h) return AGBNP_B200_ERR_ARG;
    h->prof_mask = kernel_mask;
    h->prof_used = 0;
    h->prof_recs.clear();
    return AGBNP_B200_OK;
}

int agbnp_b200_profile_read(agbnp_b200* h, double* ms_sum, int* launches, int max_kernels, const char** names) {
    if (!h || !ms_sum || !launches) return AGBNP_B200_ERR_ARG;
    if (names) *names = kKernelNames;
    try {
        CK(cudaSetDevice(h->cfg.device));
        CK(cudaDeviceSynchronize());
        for (int k = 0; k < max_kernels; k++) { ms_sum[k] = 0.0; launches[k] = 0; }
        for (const auto& r : h->prof_recs) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, h->prof_pool[r.a], h->prof_pool[r.b]));
            if (r.id < max_kernels) { ms_sum[r.id] += ms; launches[r.id]++; }
        }
        h->prof_used = 0;
        h->prof_recs.clear();
    } catch (const CudaFail& f) { h->err = f.msg; return AGBNP_B200_ERR_CUDA; }
    return K_COUNT;
}

int agbnp_b200_host_i4_tables(int n, const double* radius, const unsigned char* ishydrogen, int* type_screened,
                              int* type_screener, int* dims, double* y, double* y2, size_t cap) {
    if (n <= 0 || !radius || !ishydrogen || !dims) return AGBNP_B200_ERR_ARG;
    const Constants c = Constants::make();
    I4Tables t;
    t.build(std::vector<double>(radius, radius+n), std::vector<int>(ishydrogen, ishydrogen+n), c);
    dims[0] = t.ntypes_screened; dims[1] = t.ntypes_screener; dims[2] = t.nodes;
    for (int i = 0; i < n; i++) {
        if (type_screened) type_screened[i] = t.type_screened[i];
        if (type_screener) type_screener[i] = t.type_screener[i];
    }
    if (y || y2) {
        if (cap < t.y.size()) return AGBNP_B200_ERR_ARG;
        if (y) std::copy(t.y.begin(), t.y.end(), y);
        if (y2) std::copy(t.y2.begin(), t.y2.end(), y2);
    }
    return AGBNP_B200_OK;
}

int agbnp_b200_measure_peaks(int device, double* out, int n_out) {
    if (!out || n_out < 5) return AGBNP_B200_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return AGBNP_B200_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return AGBNP_B200_ERR_CUDA;
    cudaStream_t s; cudaEvent_t e0, e1; float* d_out = nullptr;
    if (cudaStreamCreate(&s) != cudaSuccess) return AGBNP_B200_ERR_CUDA;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaMalloc((void**) &d_out, 64);
    const int sm = prop.multiProcessorCount;
    out[0] = run_peak<0>(sm, s, e0, e1, d_out);     // FFMA lanes/s
    out[1] = run_peak<1>(sm, s, e0, e1, d_out);     // FFMA2 lanes/s
    out[2] = run_peak<2>(sm, s, e0, e1, d_out);     // MUFU.EX2 /s
    out[3] = run_peak<3>(sm, s, e0, e1, d_out);     // MUFU.RSQ /s
    out[4] = run_peak<4>(sm, s, e0, e1, d_out);     // mixed instr lanes/s
    const cudaError_t ce = cudaStreamSynchronize(s);
    cudaFree(d_out); cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(s);
    return ce == cudaSuccess ? AGBNP_B200_OK : AGBNP_B200_ERR_CUDA;
}

int agbnp_b200_get(agbnp_b200* h, int what, void* host_out, size_t bytes) {
    if (!h || !host_out) return AGBNP_B200_ERR_ARG;
    if (!h->order_valid) { h->err = "agbnp_b200_get: no evaluation has run yet"; return AGBNP_B200_ERR_ARG; }
    try {
        CK(cudaSetDevice(h->cfg.device));
        CK(cudaDeviceSynchronize());
        const int n = h->n, np = h->np;
        auto need = [&](size_t b) { if (bytes < b) throw CudaFail{"agbnp_b200_get: output buffer too small"}; };
        auto per_atom_w = [&](const float4* dptr, double* out) {
            std::vector<float4> t(np);
            CK(cudaMemcpy(t.data(), dptr, sizeof(float4)*np, cudaMemcpyDeviceToHost));
            for (int k = 0; k < np; k++) if (h->orig[k] >= 0) out[h->orig[k]] = t[k].w;
        };
        auto per_atom_f = [&](const float* dptr, double* out) {
            std::vector<float> t(np);
            CK(cudaMemcpy(t.data(), dptr, sizeof(float)*np, cudaMemcpyDeviceToHost));
            for (int k = 0; k < np; k++) if (h->orig[k] >= 0) out[h->orig[k]] = t[k];
        };
        double* od = (double*) host_out;
        switch (what) {
        case AGBNP_B200_GET_SELF_VOLUME_VDW: need(sizeof(double)*n); per_atom_w(h->d_accS, od); break;
        case AGBNP_B200_GET_SELF_VOLUME_LARGE: need(sizeof(double)*n); per_atom_w(h->d_accL, od); break;
        case AGBNP_B200_GET_SURFACE_AREA: {
            need(sizeof(double)*n);
            std::vector<double> a(n), b(n);
            per_atom_w(h->d_accL, a.data()); per_atom_w(h->d_accS, b.data());
            for (int i = 0; i < n; i++) od[i] = (a[i]-b[i])/h->k.roffset;
            break;
        }
        case AGBNP_B200_GET_BORN_RADIUS: need(sizeof(double)*n); per_atom_f(h->d_born.p, od); break;
        case AGBNP_B200_GET_VOLUME_SCALING: need(sizeof(double)*n); per_atom_f(h->d_vsf.p, od); break;
        case AGBNP_B200_GET_DERIV_Y: {
            need(sizeof(double)*n);
            std::vector<float4> t(np);
            CK(cudaMemcpy(t.data(), h->d_gbacc, sizeof(float4)*np, cudaMemcpyDeviceToHost));
            for (int k = 0; k < np; k++) if (h->orig[k] >= 0) od[h->orig[k]] = t[k].w;
            break;
        }
        case AGBNP_B200_GET_DERIV_WU: {
            need(sizeof(double)*n);
            std::vector<float4> t(np);
            CK(cudaMemcpy(t.data(), h->d_dacc, sizeof(float4)*np, cudaMemcpyDeviceToHost));
            for (int k = 0; k < np; k++) if (h->orig[k] >= 0) od[h->orig[k]] = t[k].w;
            break;
        }
        case AGBNP_B200_GET_SCALARS: {
            need(sizeof(double)*8);
            double sc[SC_COUNT];
            CK(cudaMemcpy(sc, h->d_scalars, sizeof(sc), cudaMemcpyDeviceToHost));
            unsigned long long cm = 0;
            CK(cudaMemcpy(&cm, h->d_counters+CT_M, sizeof(cm), cudaMemcpyDeviceToHost));
            const double ir = (double) (float) (1.0/h->k.roffset);          // the factor k_finish applies
            od[0] = sc[SC_EVOL_L]*ir; od[1] = -sc[SC_EVOL_S]*ir; od[2] = sc[SC_EGB]; od[3] = sc[SC_EVDW];
            od[4] = sc[SC_VOL_L]; od[5] = sc[SC_VOL_S]; od[6] = sc[SC_TOTAL]; od[7] = (double) cm;
            break;
        }
        case AGBNP_B200_GET_WORK_COUNTERS: {
            need(sizeof(double)*8);
            unsigned long long c[CT_COUNT];
            CK(cudaMemcpy(c, h->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
            for (int i = 0; i < 8; i++) od[i] = (double) c[i];
            break;
        }
        case AGBNP_B200_GET_STATS: {
            need(sizeof(double)*8);
            od[0] = (double) h->n_grow; od[1] = (double) h->n_resort; od[2] = (double) h->n_graph_inst; od[3] = (double) h->n_async_fault;
            od[4] = h->tree_cap; od[5] = h->tree_wcap; od[6] = h->nbrmax; od[7] = h->ahead_pending ? 1.0 : 0.0;
            break;
        }
        case AGBNP_B200_GET_PEER_STATE: {
            need(sizeof(double)*(PEER_KINDS + PEER_KINDS*PEER_MAX + 1));
            std::vector<int> ep(PEER_KINDS+1, 0), fl(PEER_KINDS*PEER_MAX, 0);
            if (h->d_peer_counter) CK(cudaMemcpy(ep.data(), h->d_peer_counter+2, sizeof(int)*(PEER_KINDS+1), cudaMemcpyDeviceToHost));
            if (h->d_mailbox) CK(cudaMemcpy(fl.data(), h->d_mailbox + h->peer.flag_off, sizeof(int)*PEER_KINDS*PEER_MAX, cudaMemcpyDeviceToHost));
            for (int k = 0; k < PEER_KINDS; k++) od[k] = ep[k];
            for (int k = 0; k < PEER_KINDS*PEER_MAX; k++) od[PEER_KINDS+k] = fl[k];
            od[PEER_KINDS + PEER_KINDS*PEER_MAX] = ep[PEER_KINDS];
            break;
        }
        case AGBNP_B200_GET_LIST_STATS: {
            need(sizeof(double)*4);
            int c[LC_COUNT] = {};
            if (h->d_pq_ctl.p) CK(cudaMemcpy(c, h->d_pq_ctl.p, sizeof(c), cudaMemcpyDeviceToHost));
            od[0] = c[LC_N_PQ]; od[1] = c[LC_N_L2]; od[2] = c[LC_N_EVAL]; od[3] = h->pq_skin;
            break;
        }
        case AGBNP_B200_GET_TREE_SIZE: {
            need(sizeof(long long));
            unsigned long long cm = 0;                  // nodes below the atom level, each counted by the part that owns it
            CK(cudaMemcpy(&cm, h->d_counters+CT_M, sizeof(cm), cudaMemcpyDeviceToHost));
            *(long long*) host_out = (long long) cm;
            break;
        }
        case AGBNP_B200_GET_TREE_TOPOLOGY: {
            int cur = 0;
            unsigned long long cm = 0;
            CK(cudaMemcpy(&cur, h->d_ctrl+CW_TREE_CURSOR, sizeof(int), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&cm, h->d_counters+CT_M, sizeof(cm), cudaMemcpyDeviceToHost));
            need(sizeof(int)*4*(size_t) cm);
            const int ni = (int) h->items.size();
            std::vector<int> off(ni), cnt(ni);
            std::vector<short> lvs((size_t) ni*MAX_LEVELS), rank(cur);
            std::vector<float4> rec((size_t) 2*cur);
            CK(cudaMemcpy(off.data(), h->st.root_off, sizeof(int)*ni, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(cnt.data(), h->st.root_cnt, sizeof(int)*ni, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(lvs.data(), h->st.root_lvs, sizeof(short)*ni*MAX_LEVELS, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(rec.data(), h->st.rec, sizeof(float4)*2*cur, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(rank.data(), h->st.rank, sizeof(short)*cur, cudaMemcpyDeviceToHost));
            // items ordered by the caller index of their first root (parts of a split root follow each other); a node is
            // reported by the part that owns it; a parent always precedes its children
            std::vector<int> order(ni);
            for (int i = 0; i < ni; i++) order[i] = i;
            auto first_root = [&](int i) { return h->orig[h->item_roots[h->items[i].x]]; };
            std::sort(order.begin(), order.end(), [&](int a, int b) {
                const int ra = first_root(a), rb = first_root(b);
                return ra != rb ? ra < rb : item_part(h->items[a]) < item_part(h->items[b]);
            });
            int* o = (int*) host_out;
            long long w = 0;
            std::vector<long long> dump_of;
            std::vector<int> root_of;
            for (int i : order) {
                const int ng = item_nroots(h->items[i]), part = item_part(h->items[i]), parts = item_parts(h->items[i]);
                const int n = cnt[i];
                if (n <= ng) continue;                                  // roots without overlaps
                const int l3 = lvs[(size_t) i*MAX_LEVELS+3];         // end of level 2 (level 2 exists: n > ng)
                dump_of.assign(n, -1);
                root_of.assign(n, -1);
                for (int g = 0; g < ng; g++) root_of[g] = h->orig[h->item_roots[h->items[i].x+g]];
                for (int sl = ng; sl < n; sl++) {
                    const size_t g = (size_t) off[i]+sl;
                    int a, pk;
                    std::memcpy(&a, &rec[2*g].w, 4); std::memcpy(&pk, &rec[2*g+1].w, 4);
                    const int par = pk & 0xffff;
                    if (par >= sl) throw CudaFail{"agbnp_b200_get: inconsistent tree store"};
                    root_of[sl] = root_of[par];
                    if (sl < l3 && rank[g] % parts != part) continue;   // a level-2 node of another part
                    if ((unsigned long long) w >= cm) throw CudaFail{"agbnp_b200_get: inconsistent tree store"};
                    dump_of[sl] = w;
                    o[4*w+0] = root_of[sl];
                    o[4*w+1] = par < ng ? -1 : (int) dump_of[par];
                    o[4*w+2] = h->orig[a];
                    o[4*w+3] = rank[g];
                    w++;
                }
            }
            break;
        }
        case AGBNP_B200_GET_NEIGHBOR_COUNT:
        case AGBNP_B200_GET_NEIGHBOR_PAIRS: {
            PairCommon pc = pair_common(h);
            DevBuf<unsigned long long> cnt;
            cnt.alloc(1);
            CK(cudaMemset(cnt.p, 0, sizeof(unsigned long long)));
            const long long cap = what == AGBNP_B200_GET_NEIGHBOR_PAIRS ? (long long) (bytes/(2*sizeof(int))) : 0;
            if (cap > 0) h->d_pairs.alloc((size_t) cap);
            ListArgs la{pc, h->d_pairs.p, cap, cnt.p};
            k_list_pairs<<<h->num_sm*4, 256>>>(la);
            CK(cudaDeviceSynchronize());
            unsigned long long c = 0;
            CK(cudaMemcpy(&c, cnt.p, sizeof(c), cudaMemcpyDeviceToHost));
            if (what == AGBNP_B200_GET_NEIGHBOR_COUNT) { need(sizeof(long long)); *(long long*) host_out = (long long) c; }
            else {
                if ((long long) c > cap) throw CudaFail{"agbnp_b200_get: output buffer too small for the pair list"};
                CK(cudaMemcpy(host_out, h->d_pairs.p, sizeof(int2)*c, cudaMemcpyDeviceToHost));
            }
            break;
        }
        default: h->err = "agbnp_b200_get: unknown selector"; return AGBNP_B200_ERR_ARG;
        }
    } catch (const CudaFail& f) { h->err = f.msg; return AGBNP_B200_ERR_CUDA; }
    return AGBNP_B200_OK;
}

int agbnp_b200_shard_phase(agbnp_b200* h, int phase, const void* d_posq, void* stream) {
    if (!h || phase < 0 || phase > 4) return AGBNP_B200_ERR_ARG;
    static const int masks[5] = {PH_TREE, PH_BORN, PH_BORNFIN|PH_GB, PH_DERIV, PH_GAMMA};
    try {
        CK(cudaSetDevice(h->cfg.device));
        cudaStream_t s = (cudaStream_t) stream;
        h->cur_io = true;
        if (phase == 0) {
            if (!d_posq) return AGBNP_B200_ERR_ARG;
            prepare(h, nullptr, 0, d_posq, s);
            begin_eval(h);
        }
        enqueue(h, (const float4*) d_posq, s, masks[phase], nullptr);
        mark_tail(h, s);
    } catch (const CudaFail& f) { h->err = f.msg; return AGBNP_B200_ERR_CUDA; }
    return AGBNP_B200_OK;
}

int agbnp_b200_shard_buffer(agbnp_b200* h, int which, void** d_ptr, size_t* bytes) {
    if (!h || !d_ptr || !bytes) return AGBNP_B200_ERR_ARG;
    if (!h->order_valid || h->params_dirty) { h->err = "agbnp_b200_shard_buffer: call shard_phase(0) first"; return AGBNP_B200_ERR_ARG; }
    switch (which) {
    case AGBNP_B200_BUF_SELFVOL: *d_ptr = h->d_accL; *bytes = sizeof(float4)*2*h->np; break;
    case AGBNP_B200_BUF_YQ: *d_ptr = h->d_gbacc; *bytes = sizeof(float4)*h->np; break;
    case AGBNP_B200_BUF_WU: *d_ptr = h->d_dacc; *bytes = sizeof(float4)*h->np; break;
    case AGBNP_B200_BUF_BSUM: *d_ptr = h->d_bsum; *bytes = sizeof(float)*h->np; break;
    case AGBNP_B200_BUF_FORCE: *d_ptr = h->d_gacc; *bytes = sizeof(float4)*h->np; break;
    case AGBNP_B200_BUF_ENERGY: *d_ptr = h->d_scalars; *bytes = sizeof(double)*SC_COUNT; break;
    default: return AGBNP_B200_ERR_ARG;
    }
    return AGBNP_B200_OK;
}

int agbnp_b200_shard_finish(agbnp_b200* h, void* stream, void* d_force, int force_layout, int padded_n,
                            double* d_energy, double* h_energy) {
    if (!h) return AGBNP_B200_ERR_ARG;
    try {
        CK(cudaSetDevice(h->cfg.device));
        cudaStream_t s = (cudaStream_t) stream;
        ForceSink sink{d_force, force_layout, padded_n > 0 ? padded_n : h->n, d_energy};
        h->cur_io = true;
        enqueue(h, nullptr, s, PH_FINISH, &sink);
        mark_tail(h, s);
        // asynchronous: the status words follow the evaluation through the same ring as agbnp_b200_execute_device's; the
        // call returns the outcome of the evaluation issued ASYNC_DEPTH-1 calls ago (identical on every shard: SC_FAULT)
        if (!h_energy) return async_post(h, s);
        h->evals_since_sort++; h->total_evals++;
        const int rc_deferred = async_drain(h);
        const std::string deferred_msg = h->err;
        const int status = fetch_status(h, s);
        if (status != 0) {
            h->tree_built = false;
            const bool ok = grow(h, h->h_ctrl);
            h->err = "agbnp_b200_shard_finish: capacity overflow (status " + std::to_string(status) + ")" + (ok ? "; capacities grown, re-run the evaluation" : "");
            return AGBNP_B200_ERR_CAPACITY;
        }
        grow_ahead(h, h->h_ctrl);
        *h_energy = h->h_scal[SC_TOTAL];
        if (rc_deferred != AGBNP_B200_OK) { h->err = deferred_msg; return rc_deferred; }
    } catch (const CudaFail& f) { h->err = f.msg; return AGBNP_B200_ERR_CUDA; }
    return AGBNP_B200_OK;
}

long long agbnp_b200_launch_count(const agbnp_b200* h) { return h ? h->launches : -1; }
#ifdef TAIL_DEBUG
// -DTAIL_DEBUG builds only (tools/tail_probe.py): when did the warps of the persistent kernels finish, relative to their kernel's span?
int agbnp_b200_debug_tail_reset(void) {
    unsigned long long z[16][8];
    for (auto& r : z) { for (auto& v : r) v = 0; r[0] = ~0ull; }
    cudaDeviceSynchronize();
    return cudaMemcpyToSymbol(g_tail, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
int agbnp_b200_debug_tail_read(unsigned long long* out) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, g_tail, sizeof(unsigned long long)*16*8) == cudaSuccess ? 0 : -1;
}
#endif

namespace {
// bytes of every exchange buffer; they depend only on the padded atom count, which is fixed at creation
void peer_layout(agbnp_b200* h) {
    const int nh = (int) std::count(h->sp.ishydrogen.begin(), h->sp.ishydrogen.end(), 0);
    const size_t np = (size_t) ((nh+TILE-1)/TILE*TILE + (h->n-nh+TILE-1)/TILE*TILE);
    size_t bytes[PEER_KINDS];
    bytes[AGBNP_B200_BUF_SELFVOL] = sizeof(float4)*2*np; bytes[AGBNP_B200_BUF_YQ] = sizeof(float4)*np;
    bytes[AGBNP_B200_BUF_FORCE] = sizeof(float4)*np; bytes[AGBNP_B200_BUF_ENERGY] = sizeof(double)*SC_COUNT;
    bytes[AGBNP_B200_BUF_WU] = sizeof(float4)*np; bytes[AGBNP_B200_BUF_BSUM] = (sizeof(float)*np + 15)/16*16;
    bytes[PEER_KIND_POSITIONS] = sizeof(float4)*(size_t) h->n;
    size_t o = 0;
    for (int k = 0; k < PEER_KINDS; k++) { h->peer.kind_off[k] = o; h->peer.kind_bytes[k] = bytes[k]; o += (bytes[k]*h->cfg.shard_count + 255)/256*256; }
    h->peer.flag_off = o;
    h->mailbox_bytes = o + sizeof(int)*PEER_KINDS*PEER_MAX;
}
}

namespace {
void ensure_mailbox(agbnp_b200* h) {
    CK(cudaSetDevice(h->cfg.device));
    if (h->d_mailbox) return;
    peer_layout(h);
    CK(cudaMalloc((void**) &h->d_mailbox, h->mailbox_bytes));
    CK(cudaMemset(h->d_mailbox, 0, h->mailbox_bytes));
    CK(cudaMalloc((void**) &h->d_peer_counter, sizeof(int)*(3+PEER_KINDS)));
    CK(cudaMemset(h->d_peer_counter, 0, sizeof(int)*(3+PEER_KINDS)));
    CK(cudaDeviceSynchronize());
}
void peer_bind(agbnp_b200* h, int shard_count) {
    h->peer.rank = h->cfg.shard_rank; h->peer.count = shard_count; h->peer.counter = h->d_peer_counter; h->peer.epochs = h->d_peer_counter+2;
    h->peer.status = h->d_peer_counter+2+PEER_KINDS;
}
}

int agbnp_b200_peer_import_local(agbnp_b200* h, agbnp_b200* const* shards, int shard_count) {
    if (!h || !shards || shard_count != h->cfg.shard_count || shard_count > PEER_MAX) return AGBNP_B200_ERR_ARG;
    try {
        for (int p = 0; p < shard_count; p++) {
            agbnp_b200* q = shards[p];
            if (!q || q->cfg.shard_count != shard_count || q->cfg.shard_rank != p || q->n != h->n) {
                h->err = "agbnp_b200_peer_import_local: shards[p] must be the handle of shard p of the same system"; return AGBNP_B200_ERR_ARG;
            }
            ensure_mailbox(q);
        }
        CK(cudaSetDevice(h->cfg.device));
        peer_bind(h, shard_count);
        for (int p = 0; p < shard_count; p++) {
            if (shards[p]->cfg.device != h->cfg.device) {
                const cudaError_t ce = cudaDeviceEnablePeerAccess(shards[p]->cfg.device, 0);
                if (ce == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (ce != cudaSuccess) throw CudaFail{std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(ce)};
            }
            h->peer.mail[p] = shards[p]->d_mailbox;
        }
        h->peer_ready = true;
    } catch (const CudaFail& f) { h->err = f.msg; return AGBNP_B200_ERR_CUDA; }
    return AGBNP_B200_OK;
}

int agbnp_b200_peer_export(agbnp_b200* h, void* ipc_handle) {
    if (!h || !ipc_handle) return AGBNP_B200_ERR_ARG;
    if (h->cfg.shard_count > PEER_MAX) { h->err = "agbnp_b200_peer_export: at most 8 shards"; return AGBNP_B200_ERR_ARG; }
    try {
        ensure_mailbox(h);
        cudaIpcMemHandle_t hd;
        CK(cudaIpcGetMemHandle(&hd, h->d_mailbox));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        std::memcpy(ipc_handle, &hd, 64);
    } catch (const CudaFail& f) { h->err = f.msg; return AGBNP_B200_ERR_CUDA; }
    return AGBNP_B200_OK;
}

int agbnp_b200_peer_import(agbnp_b200* h, const void* ipc_handles, int shard_count) {
    if (!h || !ipc_handles || shard_count != h->cfg.shard_count || !h->d_mailbox) return AGBNP_B200_ERR_ARG;
    try {
        CK(cudaSetDevice(h->cfg.device));
        peer_bind(h, shard_count);
        for (int p = 0; p < shard_count; p++) {
            if (p == h->cfg.shard_rank) { h->peer.mail[p] = h->d_mailbox; continue; }
            cudaIpcMemHandle_t hd;
            std::memcpy(&hd, (const unsigned char*) ipc_handles + 64*(size_t) p, 64);
            void* ptr = nullptr;
            CK(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
            h->peer.mail[p] = (unsigned char*) ptr;
            h->peer_opened.push_back(ptr);
        }
        h->peer_ready = true;
    } catch (const CudaFail& f) { h->err = f.msg; return AGBNP_B200_ERR_CUDA; }
    return AGBNP_B200_OK;
}

int agbnp_b200_peer_broadcast(agbnp_b200* h, void* d_posq, int owner, void* stream) {
    if (!h || !d_posq || owner < 0 || owner >= h->cfg.shard_count) return AGBNP_B200_ERR_ARG;
    if (!h->peer_ready) { h->err = "agbnp_b200_peer_broadcast: peer_export / peer_import first"; return AGBNP_B200_ERR_ARG; }
    try {
        CK(cudaSetDevice(h->cfg.device));
        k_peer_broadcast<<<(int) std::min<size_t>(64, ((size_t) h->n+255)/256), 256, 0, (cudaStream_t) stream>>>((float4*) d_posq, (size_t) h->n, h->peer, owner);
        h->launches += 1;
        CK(cudaGetLastError());
    } catch (const CudaFail& f) { h->err = f.msg; return AGBNP_B200_ERR_CUDA; }
    return AGBNP_B200_OK;
}

int agbnp_b200_peer_exchange(agbnp_b200* h, int which, void* stream) {
    if (!h || which < 0 || which >= PEER_KIND_POSITIONS) return AGBNP_B200_ERR_ARG;
    if (!h->peer_ready) { h->err = "agbnp_b200_peer_exchange: peer_export / peer_import first"; return AGBNP_B200_ERR_ARG; }
    try {
        CK(cudaSetDevice(h->cfg.device));
        peer_enqueue(h, which, (cudaStream_t) stream);
        mark_tail(h, (cudaStream_t) stream);
    } catch (const CudaFail& f) { h->err = f.msg; return AGBNP_B200_ERR_CUDA; }
    return AGBNP_B200_OK;
}

int agbnp_b200_shard_evaluate(agbnp_b200* h, void* d_posq, int owner, void* stream, void* d_force, int force_layout, int padded_n,
                              double* d_energy) {
    if (!h || !d_posq || owner < 0 || owner >= h->cfg.shard_count) return AGBNP_B200_ERR_ARG;
    if (d_force && force_layout != 0 && force_layout != 1) { h->err = "agbnp_b200_shard_evaluate: bad force layout"; return AGBNP_B200_ERR_ARG; }
    if (!h->peer_ready) { h->err = "agbnp_b200_shard_evaluate: peer_export / peer_import first"; return AGBNP_B200_ERR_ARG; }
    try {
        CK(cudaSetDevice(h->cfg.device));
        cudaStream_t s = (cudaStream_t) stream;
        k_peer_broadcast<<<(int) std::min<size_t>(64, ((size_t) h->n+255)/256), 256, 0, s>>>((float4*) d_posq, (size_t) h->n, h->peer, owner);
        h->launches += 1;
        h->cur_io = true;
        prepare(h, nullptr, 0, d_posq, s);                  // (re)sorting reads the broadcast positions; same evaluation count on every shard
        ForceSink sink{d_force, force_layout, padded_n > 0 ? padded_n : h->n, d_energy};
        launch_all(h, (const float4*) d_posq, s, &sink, true);
        // deferred validation, as in the asynchronous single-GPU path -- AFTER the whole collective sequence is enqueued
        return async_post(h, s);
    } catch (const CudaFail& f) { h->err = f.msg; return AGBNP_B200_ERR_CUDA; }
    return AGBNP_B200_OK;
}

} // extern "C"
