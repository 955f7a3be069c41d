// Oracle shim (test infrastructure only): minimal stand-in for openmm/OpenMMException.h so the
// reference's own sources under /root/reference compile without an OpenMM install.
#ifndef ORACLE_SHIM_OPENMMEXCEPTION_H_
#define ORACLE_SHIM_OPENMMEXCEPTION_H_
#include <exception>
#include <string>
namespace OpenMM {
class OpenMMException : public std::exception {
public:
    explicit OpenMMException(const std::string& m) : msg(m) {}
    ~OpenMMException() throw() {}
    const char* what() const throw() { return msg.c_str(); }
private:
    std::string msg;
};
}
#endif
