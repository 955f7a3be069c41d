/* agbnp_b200.h -- C-ABI of libagbnp_b200.so: the B200-native (sm_100a) AGBNP1 / GaussVol energy+force path.
 *
 * This is the drop-in boundary for ONE path of Gallicchio-Lab/openmm_agbnp_plugin: what the reference does inside
 *     CalcAGBNPForceKernel::initialize / execute / copyParametersToContext
 *         (openmmapi/include/AGBNPKernels.h:19-47; Reference platform: platforms/reference/src/ReferenceAGBNPKernels.cpp:58-149,
 *          1796-1815; OpenCL platform being replaced: platforms/opencl/src/OpenCLAGBNPKernels.cpp:393-556,5439-5468).
 * A platforms/cuda kernel object (see INTEGRATION.md and openmm_agbnp_plugin_b200/platforms/cuda/) holds one handle per
 * OpenMM Context and forwards those three calls here.  Plain C types only: no torch, no OpenMM, no C++ in the signatures.
 *
 * Conventions
 *   - every function returns 0 on success and a negative agbnp_b200_status on failure; nothing throws across the ABI;
 *     the message is available from agbnp_b200_last_error(handle) (or (NULL) for create failures).
 *   - units are OpenMM's: nm, kJ/mol, elementary charge (reference README.md:97-103).
 *   - a handle is bound to one CUDA device and one stream at a time and is not re-entrant; distinct handles are
 *     independent (replica mode: one handle per GPU).
 *   - there is no CPU fallback: if no CUDA device is usable, create fails with AGBNP_B200_ERR_CUDA.
 */
#ifndef AGBNP_B200_H_
#define AGBNP_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct agbnp_b200 agbnp_b200;

typedef enum {
    AGBNP_B200_OK = 0,
    AGBNP_B200_ERR_ARG = -1,          /* bad argument (incl. illegal version, CutoffPeriodic, multiple gamma values) */
    AGBNP_B200_ERR_CUDA = -2,         /* CUDA runtime failure / no device */
    AGBNP_B200_ERR_CAPACITY = -3,     /* internal capacity could not be grown */
    AGBNP_B200_ERR_PARAM_CHANGE = -4  /* copyParametersToContext rules violated (N, radius, heavy<->hydrogen) */
} agbnp_b200_status;

/* AGBNPForce::NonbondedMethod (openmmapi/include/AGBNPForce.h:44-59) */
#define AGBNP_B200_NOCUTOFF 0
#define AGBNP_B200_CUTOFF_NONPERIODIC 1
#define AGBNP_B200_CUTOFF_PERIODIC 2 /* accepted by the reference API, implemented nowhere: rejected here */

typedef struct {
    int version;           /* AGBNPForce::setVersion: 0 = GVolSA, 1 = AGBNP1 (default); 2 (AGBNP2) is out of scope -> ERR_ARG */
    int nonbonded_method;  /* AGBNP_B200_NOCUTOFF | AGBNP_B200_CUTOFF_NONPERIODIC */
    double cutoff;         /* nm; ignored for NoCutoff (AGBNPForce default 1.0, AGBNPForce.cpp:15) */
    int device;            /* CUDA device ordinal */
    /* sharding of one evaluation across GPUs (see "multi-GPU plumbing" below): this handle evaluates the
     * share of shard `shard_rank` of `shard_count`.  1 GPU: rank 0 of 1. */
    int shard_rank;
    int shard_count;
    int reorder_interval;  /* evaluations between spatial re-sorts of the internal atom order; <= 0: library default */
    /* OPT-IN, not the reference's semantics (it rebuilds the overlap tree in every evaluation): > 1 keeps the tree's
     * topology for this many evaluations and only re-evaluates the stored overlaps at the new positions in between
     * (for MD, where atoms move ~1e-3 nm per step).  <= 1: build every time (default).  The
     * environment variable AGBNP_B200_TREE_REUSE sets it when this field is <= 0. */
    int tree_reuse_interval;
} agbnp_b200_config;

/* fill with the reference defaults: version 1, NoCutoff, cutoff 1.0 nm, device 0, shard 0 of 1 */
void agbnp_b200_default_config(agbnp_b200_config* cfg);

/* == CalcAGBNPForceKernel::initialize.  Per-particle arrays are what AGBNPForce::addParticle(radius, gamma, vdw_alpha,
 * charge, ishydrogen) collected (AGBNPForce.h:75).  Hydrogens get gamma 0 and volume 0; all heavy-atom gammas must be
 * equal, otherwise ERR_ARG with the reference's message "initialize(): AGBNP does not support multiple gamma values."
 * (ReferenceAGBNPKernels.cpp:110-116). */
int agbnp_b200_create(const agbnp_b200_config* cfg, int num_particles, const double* radius, const double* gamma,
                      const double* vdw_alpha, const double* charge, const unsigned char* ishydrogen,
                      agbnp_b200** out);

void agbnp_b200_destroy(agbnp_b200* h);

const char* agbnp_b200_last_error(const agbnp_b200* h);

/* == CalcAGBNPForceKernel::copyParametersToContext (ReferenceAGBNPKernels.cpp:1796-1815): gamma, alpha and charge may
 * change; a different particle count, a radius changed by more than 1e-3 nm or a heavy atom turned hydrogen return
 * ERR_PARAM_CHANGE with the reference's messages. */
int agbnp_b200_set_params(agbnp_b200* h, int num_particles, const double* radius, const double* gamma,
                          const double* vdw_alpha, const double* charge, const unsigned char* ishydrogen);

/* == CalcAGBNPForceKernel::execute, host buffers (the Reference platform's calling convention: positions in, energy
 * returned, forces ADDED into the caller's array -- ReferenceAGBNPKernels.cpp:27-35,139-149).
 *   pos      [3*N] doubles, nm (rounded to float on upload: the device path computes in float / selective double)
 *   forces   [3*N] doubles, kJ/mol/nm, accumulated (+=); may be NULL when include_forces == 0
 *            include_forces == AGBNP_B200_FORCES_ASSIGN (2): assigned (=) instead -- for a caller that would zero the array
 *            right before the call anyway, as OpenMM's ContextImpl::calcForcesAndEnergy does before it runs the force kernels
 *            (saves that pass and the read of the old values: ~10 us of host time for 18 k atoms)
 *   energy   receives the potential energy (kJ/mol); may be NULL
 * The host<->device copies are part of the call (this is what bench.py's `e2e` times).  Synchronous. */
#define AGBNP_B200_FORCES_ASSIGN 2
int agbnp_b200_execute_host(agbnp_b200* h, const double* pos, int include_forces, int include_energy,
                            double* energy, double* forces);

/* == CalcAGBNPForceKernel::execute, device buffers (the CUDA-platform calling convention).
 *   d_posq        device float4[padded or N] (double4 after agbnp_b200_set_device_layout(posq_is_double)): x,y,z (nm) and a
 *                 charge slot that is ignored (charges come from create/set_params)
 *   stream        cudaStream_t cast to void* (NULL = default stream); all work is enqueued on it
 *   d_force       device force sink, or NULL:
 *                   layout 0: float[3*N] xyz-interleaved, forces are added (+=)
 *                   layout 1: OpenMM CUDA fixed point: unsigned long long[3*padded_n], component-major
 *                             (x[0..padded_n), y[..], z[..]), value*2^32, added with 64-bit atomics
 *   d_energy      device double (float after agbnp_b200_set_device_layout(energy_is_float)) accumulator to which the energy
 *                 is added, or NULL
 *   h_energy      host double receiving the energy (forces a stream synchronize), or NULL
 * Returns after enqueueing unless h_energy is given (see "Asynchronous use" below). */
int agbnp_b200_execute_device(agbnp_b200* h, const void* d_posq, void* stream, void* d_force, int force_layout,
                              int padded_n, double* d_energy, double* h_energy);

/* Layout of the caller's DEVICE buffers, for a caller that is OpenMM's CUDA platform (CudaContext):
 *   atom_index       host int[N]: atom_index[p] = the particle (numbered as in agbnp_b200_create) whose position is stored at
 *                    index p of d_posq and whose force belongs at index p of d_force -- CudaContext::getAtomIndex(), which
 *                    changes whenever the context reorders its atoms (call this again from a CudaContext::ReorderListener).
 *                    NULL = identity.  The array is copied.
 *   posq_is_double   d_posq is double4[] (CudaContext::getUseDoublePrecision()) instead of float4[]
 *   energy_is_float  d_energy points to a float accumulator (CudaContext's energy buffer in single-precision mode)
 * Takes effect from the next evaluation; the handle, its parameters and its internal (spatial) atom order are untouched, so
 * a reorder costs one N-int upload.  agbnp_b200_execute_host ignores the layout (its arrays are in particle order). */
typedef struct {
    const int* atom_index;
    int posq_is_double;
    int energy_is_float;
} agbnp_b200_device_layout;
int agbnp_b200_set_device_layout(agbnp_b200* h, const agbnp_b200_device_layout* layout /* NULL = defaults */);

/* Asynchronous use of agbnp_b200_execute_device (h_energy == NULL): the call returns after enqueueing; forces and energy
 * are delivered on the stream by the last kernel of the evaluation, and only if no internal capacity overflowed (the
 * check is on the device).  The status words follow the evaluation to pinned memory and are examined a few calls later
 * and by agbnp_b200_synchronize: capacities are grown ahead of need from the high-water marks every evaluation reports
 * (a capacity grows as soon as a high-water mark passes 90% of it, so an overflow needs a >11% jump in local packing
 * between two evaluations); if one happens anyway the call that
 * notices it returns ERR_CAPACITY naming the evaluation, which must be re-issued.  The call that reports a fault has
 * itself enqueued ITS evaluation as usual (an error code never means "this call did nothing"); evaluations that were
 * already in flight when the fault happened ran with the old capacities and are reported by the calls that retire them;
 * the capacities grow once per fault.  With h_energy != NULL (and in agbnp_b200_execute_host) the call is synchronous and
 * re-runs an overflowed evaluation itself; a pending fault of an earlier asynchronous evaluation is reported by it AFTER
 * its own evaluation has completed (energy and forces of the synchronous evaluation are valid).
 * agbnp_b200_synchronize waits for the stream and retires every pending status. */
int agbnp_b200_synchronize(agbnp_b200* h, void* stream);

/* Per-kernel CUDA-event brackets for bench.py's roofline: kernels whose bit is set in kernel_mask (bit i = i-th name of
 * agbnp_b200_profile_read's list) get an event before and after every launch, on the launching stream, until the mask
 * is cleared.  profile_read synchronises the device, sums the bracketed durations (ms) and launch counts per kernel,
 * and resets the record.  Returns the number of kernels in the list. */
int agbnp_b200_profile(agbnp_b200* h, unsigned kernel_mask);
int agbnp_b200_profile_read(agbnp_b200* h, double* ms_sum, int* launches, int max_kernels, const char** names);

/* kernels this handle has launched since creation (bench.py's gpu_launches) */
long long agbnp_b200_launch_count(const agbnp_b200* h);

/* Issue-rate microbenchmarks on `device` for the FP32/SFU roofline denominators (north_star: "fraction of the FP32/SFU
 * roofline"): out[0] scalar FFMA lanes/s, out[1] packed fma.rn.f32x2 lanes/s, out[2] MUFU.EX2 ops/s, out[3] MUFU.RSQ
 * ops/s, out[4] instruction lanes/s of a 14:1 FFMA:MUFU mix.  n_out >= 5. */
int agbnp_b200_measure_peaks(int device, double* out, int n_out);

/* Host-only diagnostic (no device needed): the radius typing and the natural-cubic-spline I4 tables that agbnp_b200_create
 * builds for these particles -- what AGBNPI4LookupTable / the Reference kernel's initialize produce
 * (openmmapi/src/AGBNPUtils.cpp:13-214, platforms/reference/src/ReferenceAGBNPKernels.cpp:96-137).
 * type_screened[N], type_screener[N] (-1 for hydrogens); dims[3] = {screened types, screener types, nodes per table};
 * y / y2 (each of capacity `cap` doubles, may be NULL to query dims): knot values and second derivatives,
 * table (ti, tj) at [(ti*dims[1] + tj)*dims[2] ...].  Knot k sits at k*2.0/(nodes-1) nm. */
int agbnp_b200_host_i4_tables(int n, const double* radius, const unsigned char* ishydrogen, int* type_screened,
                              int* type_screener, int* dims, double* y, double* y2, size_t cap);

/* diagnostics / by-products of the last evaluation, copied to host (atom order = caller's order).  `what`: */
typedef enum {
    AGBNP_B200_GET_SELF_VOLUME_VDW = 0,   /* double[N]  self-volumes, vdW radii (after S3) */
    AGBNP_B200_GET_SELF_VOLUME_LARGE = 1, /* double[N]  self-volumes, enlarged radii (after S2) */
    AGBNP_B200_GET_SURFACE_AREA = 2,      /* double[N]  (selfvol_large - selfvol_vdw)/roffset, nm^2 */
    AGBNP_B200_GET_BORN_RADIUS = 3,       /* double[N]  nm */
    AGBNP_B200_GET_VOLUME_SCALING = 4,    /* double[N]  s_i */
    AGBNP_B200_GET_SCALARS = 5,           /* double[8]  vol_energy1, vol_energy2, gb_energy(self+pair), vdw_energy,
                                                         volume1, volume2, total, n_tree_nodes */
    AGBNP_B200_GET_TREE_SIZE = 6,         /* long long[1] number of overlap-tree nodes below the atom level */
    AGBNP_B200_GET_TREE_TOPOLOGY = 7,     /* int[4*M]   per node: root atom, parent (index into this dump, -1 = root
                                                         atom), last atom, sibling rank; a parent precedes its children */
    AGBNP_B200_GET_DERIV_Y = 8,           /* double[N]  Y_i (GB derivative accumulator) */
    AGBNP_B200_GET_DERIV_WU = 9,          /* double[N]  (W_i + U_i) before division by the atomic volume */
    AGBNP_B200_GET_NEIGHBOR_PAIRS = 10,   /* int[2*P]   (i<j) with r2 < cutoff2 as used by the GB pass (cutoff mode) */
    AGBNP_B200_GET_NEIGHBOR_COUNT = 11,   /* long long[1] P */
    AGBNP_B200_GET_WORK_COUNTERS = 12,    /* double[8]  P_gb, P_q(directed, evaluated), C2, C3+, M, tiles_gb, tiles_q, - */
    AGBNP_B200_GET_STATS = 13             /* double[8]  since creation: capacity growths, spatial re-sorts, CUDA-graph instantiations,
                                                         asynchronous evaluations found overflowed; now: nodes-per-root capacity,
                                                         nodes-per-level capacity, level-2 neighbor capacity, 1 if a growth is pending */
    ,AGBNP_B200_GET_LIST_STATS = 14       /* double[4]  since the Verlet lists were last voided (re-sort, capacity growth): evaluations
                                                         that rebuilt the pair masks, that rebuilt the level-2 candidate lists,
                                                         evaluations in all; the list skin (nm) */
    ,AGBNP_B200_GET_PEER_STATE = 15       /* double[7 + 7*8 + 1] diagnostics of the peer-memory exchange: exchanges completed per buffer
                                                         kind (0..5 = agbnp_b200_buffer, 6 = positions), the flags the peers have raised in
                                                         this shard's mailbox [kind][source shard], the sticky fault word */
} agbnp_b200_get_what;

int agbnp_b200_get(agbnp_b200* h, int what, void* host_out, size_t bytes);

/* ---- multi-GPU plumbing: one process per GPU; the caller runs the phases and does the collectives between
 * them on the exported device buffers (NCCL through torch.distributed in this repo's host layer, sharding.py).
 * Work split: overlap-tree work items (roots, or parts of large roots) are dealt block-cyclically in most-expensive-first
 * order (each shard builds, sweeps and stores only its subtrees);
 * the work units of the Born-radius, GB and derivative pair passes are dealt round-robin.
 *   phase 0: gather/sort, tree build + rescan + sweeps for the owned roots   -> all-reduce SELFVOL (8*np floats)
 *   phase 1: Born-radius pair sums for the owned units                        -> all-reduce BSUM    (np floats)
 *   phase 2: Born radii (all atoms), GB pair pass for the owned tile units    -> all-reduce YQ      (4*np floats)
 *   phase 3: derivative pass for the owned units                              -> all-reduce WU      (4*np floats)
 *   phase 4: tree gamma sweep over the owned subtrees                         -> all-reduce FORCE   (4*np floats)
 *                                                                                 and ENERGY (8 doubles)
 *   finish : scatter forces to the caller's sink, total energy
 * With shard_count == 1 execute_* run all phases back to back. */
int agbnp_b200_shard_phase(agbnp_b200* h, int phase, const void* d_posq, void* stream);
typedef enum {
    AGBNP_B200_BUF_SELFVOL = 0,  /* float[8*np]   per atom (gradient xyz, self-volume): enlarged radii, then vdW radii; internal order */
    AGBNP_B200_BUF_YQ = 1,       /* float[4*np]   partial GB pair force (xyz) and GB derivative accumulator Y (w) */
    AGBNP_B200_BUF_FORCE = 2,    /* float[4*np]   partial forces of the W+U tree sweep (xyz, -) */
    AGBNP_B200_BUF_ENERGY = 3,   /* double[8]     partial energy scalars */
    AGBNP_B200_BUF_WU = 4,       /* float[4*np]   partial derivative-pass force (xyz) and W+U (w) */
    AGBNP_B200_BUF_BSUM = 5      /* float[np]     partial Born-radius pair sums */
} agbnp_b200_buffer;
int agbnp_b200_shard_buffer(agbnp_b200* h, int which, void** d_ptr, size_t* bytes);
/* h_energy == NULL: asynchronous, with the deferred validation of agbnp_b200_execute_device (the status words are part of
 * the ENERGY exchange, so a fault is seen by every shard, at the same call).  Otherwise synchronises and returns
 * ERR_CAPACITY if ANY shard overflowed: the last kernel of phase 4 folds the shard's status word into the ENERGY buffer, the
 * exchange sums it, and the finish kernel of every shard then withholds the delivery (sharding.py still all-reduces the
 * return code before re-running, which covers callers whose exchange is not the library's). */
int agbnp_b200_shard_finish(agbnp_b200* h, void* stream, void* d_force, int force_layout, int padded_n,
                            double* d_energy, double* h_energy);

/* ---- peer-memory exchange over NVLink (optional replacement of the all-reduces above; same node, one process per GPU).
 * Every shard owns a "mailbox" in its own HBM with one slot per (exchange buffer, source shard).  An exchange of buffer X is
 *   push: each shard stores its partial X into its slot of every PEER's mailbox (plain P2P stores over NVLink), fences,
 *         and raises a per-(buffer, source) epoch flag in the peer's mailbox;
 *   sum : each shard waits for the flags of all peers, then adds the peers' slots to its own partial -- a one-shot
 *         all-reduce whose latency is one NVLink store + one flag round trip instead of a ring/tree of NCCL steps.
 * Set-up: every shard exports a CUDA IPC handle of its mailbox (64 bytes), the caller gathers the handles of all shards
 * (any host-side all-gather) and hands them to every shard. */
int agbnp_b200_peer_export(agbnp_b200* h, void* ipc_handle_64_bytes);
int agbnp_b200_peer_import(agbnp_b200* h, const void* ipc_handles /* [shard_count][64] */, int shard_count);
/* The same wiring when all shards live in ONE process (one host thread driving several GPUs, or several shards on one GPU):
 * CUDA IPC handles cannot be opened by the process that exported them, so the mailboxes are linked directly (with
 * cudaDeviceEnablePeerAccess between different devices).  shards[p] = handle of shard p; call it on every handle.  With a
 * single host thread, issue the shards' calls of one evaluation on DIFFERENT streams and start with the position owner. */
int agbnp_b200_peer_import_local(agbnp_b200* h, agbnp_b200* const* shards, int shard_count);
/* The waits are bounded: a peer whose flag does not arrive within ~4 s (a dead process, shards driven out of step) makes the
 * exchange give up; from then on the handle delivers nothing and reports ERR_CAPACITY with status bit 128 (peer timeout)
 * instead of hanging the GPU. */
int agbnp_b200_peer_exchange(agbnp_b200* h, int which /* agbnp_b200_buffer */, void* stream);
/* positions (device float4[N], caller's order) from shard `owner` to every shard's own d_posq, same mechanism */
int agbnp_b200_peer_broadcast(agbnp_b200* h, void* d_posq, int owner, void* stream);
/* One whole sharded evaluation, asynchronous, in one call: position broadcast from `owner`, then the five phases with their
 * peer-memory exchanges and the finish kernel, enqueued back to back (the exchange kernels keep their epochs in device
 * memory, so no launch argument changes between evaluations).  Every shard must call it the same number of times.
 * A capacity overflow on ANY shard suppresses the delivery on EVERY shard (the status words travel with the ENERGY exchange)
 * and is reported, by every shard, by the call issued ASYNC_DEPTH-1 evaluations later -- after that call has enqueued its
 * own full collective sequence, so the shards never fall out of step. */
int agbnp_b200_shard_evaluate(agbnp_b200* h, void* d_posq, int owner, void* stream, void* d_force, int force_layout,
                              int padded_n, double* d_energy);

/* library / build identification, e.g. "agbnp_b200 0.1 sm_100a" */
const char* agbnp_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AGBNP_B200_H_ */
