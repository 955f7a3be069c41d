/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
 *
 * Plain-C restatement of the reference's Reference-platform AGBNP1 / GaussVol path
 * (gaussvol/gaussvol.cpp, openmmapi/src/AGBNPUtils.cpp, platforms/reference/src/ReferenceAGBNPKernels.cpp).
 * Parity status:
 *   NoCutoff        : PINNED -- bit-compared against oracle/_ref (the unmodified reference) and checked against the
 *                     golden files platforms/reference/tests/{v0,v1}.reference (tests/test_oracle.py).
 *   CutoffNonPeriodic: "parity unpinned" -- the Reference platform has no cutoff and no reference test covers it; the
 *                     rule restated here is the OpenCL back-end's (every pair pass keeps a pair iff r2 < cutoff2,
 *                     GVolOverlapTree.cl:293, AGBNPBornRadii.cl:430, AGBNPGBEnergy.cl:313), applied to the Reference
 *                     platform's double arithmetic.  r2 is evaluated in float from float-rounded coordinates so that
 *                     membership is bit-exact against the GPU.
 */
#ifndef AGBNP_ORACLE_H_
#define AGBNP_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct agbnp_oracle agbnp_oracle;

/* nonbonded_method: 0 NoCutoff, 1 CutoffNonPeriodic (2 CutoffPeriodic is rejected).  Returns NULL on error. */
agbnp_oracle* agbnp_oracle_create(int version, int nonbonded_method, double cutoff, int n,
                                  const double* radius, const double* gamma, const double* alpha,
                                  const double* charge, const int* ishydrogen);
void agbnp_oracle_destroy(agbnp_oracle* h);
const char* agbnp_oracle_last_error(void);

/* copyParametersToContext semantics; returns 0 or -1 */
int agbnp_oracle_set_params(agbnp_oracle* h, const double* radius, const double* gamma, const double* alpha,
                            const double* charge, const int* ishydrogen);

/* one evaluation: energy returned through *energy, forces[3n] overwritten */
int agbnp_oracle_execute(agbnp_oracle* h, const double* pos, double* energy, double* forces);

/* per-atom by-products of the last execute.  what: 0 self_volume (vdW radii), 1 self_volume (large radii),
 * 2 volume_scaling_factor, 3 born_radius, 4 inverse_born_radius_fp, 5 Y, 6 bru, 7 brw, 8 W, 9 U,
 * 10 radius_type_screened, 11 radius_type_screener, 12 free_volume (vdW), 13 free_volume (large) */
int agbnp_oracle_get(agbnp_oracle* h, int what, double* out);

/* scalars of the last execute: 0 vol_energy1, 1 vol_energy2, 2 gb_self, 3 gb_pair, 4 evdw, 5 volume1, 6 volume2 */
double agbnp_oracle_scalar(agbnp_oracle* h, int what);

/* overlap tree (topology is from the large-radius build; values are those of the last rescan) */
int agbnp_oracle_tree_size(agbnp_oracle* h);
int agbnp_oracle_tree_dump(agbnp_oracle* h, int* level, int* atom, int* parent, int* child_start, int* child_count,
                           double* volume, double* gvol);

/* I4 tables: dims and node values */
int agbnp_oracle_i4_dims(agbnp_oracle* h, int* ntypes_screened, int* ntypes_screener, int* nnodes);
int agbnp_oracle_i4_table(agbnp_oracle* h, int ti, int tj, double* x, double* y, double* y2);

/* pair-membership with the float r2 < cutoff2 rule: writes up to max_pairs (i<j) pairs, returns the total count */
long agbnp_oracle_neighbor_pairs(int n, const float* pos, float cutoff, int* pairs, long max_pairs);

/* counters of the last execute for the algorithmic-work formula:
 * 0 P_gb, 1 P_q (directed, evaluated), 2 C2 (level-2 ogauss evaluations), 3 C3+ (deeper ogauss evaluations), 4 M nodes */
double agbnp_oracle_counter(agbnp_oracle* h, int what);

#ifdef __cplusplus
}
#endif
#endif
