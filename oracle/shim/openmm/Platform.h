#ifndef ORACLE_SHIM_PLATFORM_H_
#define ORACLE_SHIM_PLATFORM_H_
#include <string>
namespace OpenMM {
class Platform { public: virtual ~Platform() {} };
}
#endif
