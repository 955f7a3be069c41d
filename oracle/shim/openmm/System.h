#ifndef ORACLE_SHIM_SYSTEM_H_
#define ORACLE_SHIM_SYSTEM_H_
namespace OpenMM { class System {}; }
#endif
