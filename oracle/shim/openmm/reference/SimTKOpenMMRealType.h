// Oracle shim: the Reference platform computes in double.
#ifndef ORACLE_SHIM_REALTYPE_H_
#define ORACLE_SHIM_REALTYPE_H_
typedef double RealOpenMM;
#define DOT3(u,v) ((u[0])*(v[0]) + (u[1])*(v[1]) + (u[2])*(v[2]))
#endif
