// Minimal stand-in for the handful of OpenMM types the AGBNP plugin touches, used ONLY when the plugin is built without an
// OpenMM installation (this image has none): enough of Platform / KernelFactory / KernelImpl / ContextImpl / Context for
// AGBNPForce -> AGBNPForceImpl -> CalcAGBNPForceKernel to run exactly as it does inside OpenMM's CUDA platform: the
// context owns a CudaContext (openmm/cuda/CudaContext.h stand-in) with device-resident posq / force / energy buffers in
// the platform's atom order and precision, and the kernel object reads and writes those.
// With -DAGBNP_B200_WITH_OPENMM the real headers are used and this file is not included.
#ifndef AGBNP_B200_OPENMM_MINI_H_
#define AGBNP_B200_OPENMM_MINI_H_

#include <cmath>
#include <exception>
#include <map>
#include <string>
#include <vector>

namespace OpenMM {

class OpenMMException : public std::exception {
public:
    explicit OpenMMException(const std::string& m) : msg(m) {}
    ~OpenMMException() throw() {}
    const char* what() const throw() { return msg.c_str(); }
private:
    std::string msg;
};

class Vec3 {
public:
    Vec3() { v[0] = v[1] = v[2] = 0.0; }
    Vec3(double x, double y, double z) { v[0] = x; v[1] = y; v[2] = z; }
    double operator[](int i) const { return v[i]; }
    double& operator[](int i) { return v[i]; }
    Vec3 operator*(double s) const { return Vec3(v[0]*s, v[1]*s, v[2]*s); }
    Vec3 operator+(const Vec3& o) const { return Vec3(v[0]+o.v[0], v[1]+o.v[1], v[2]+o.v[2]); }
    Vec3 operator-(const Vec3& o) const { return Vec3(v[0]-o.v[0], v[1]-o.v[1], v[2]-o.v[2]); }
    Vec3& operator+=(const Vec3& o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
    double dot(const Vec3& o) const { return v[0]*o.v[0] + v[1]*o.v[1] + v[2]*o.v[2]; }
private:
    double v[3];
};

class Force;
class ForceImpl;
class ContextImpl;
class Context;
class Platform;

class System {
public:
    ~System();
    int getNumParticles() const { return (int) masses.size(); }
    int addParticle(double mass) { masses.push_back(mass); return (int) masses.size()-1; }
    int addForce(Force* f) { forces.push_back(f); return (int) forces.size()-1; }      // takes ownership, as OpenMM does
    int getNumForces() const { return (int) forces.size(); }
    Force& getForce(int i) const { return *forces[i]; }
private:
    std::vector<double> masses;
    std::vector<Force*> forces;
};

class KernelImpl {
public:
    KernelImpl(std::string name, const Platform& platform) : name(name), platform(&platform) {}
    virtual ~KernelImpl() {}
    const std::string& getName() const { return name; }
    const Platform& getPlatform() const { return *platform; }
private:
    std::string name;
    const Platform* platform;
};

class Kernel {
public:
    Kernel() : impl(0) {}
    explicit Kernel(KernelImpl* impl) : impl(impl) {}
    template <class T> T& getAs() { return dynamic_cast<T&>(*impl); }
    KernelImpl* release() { KernelImpl* p = impl; impl = 0; return p; }
    KernelImpl* get() const { return impl; }
private:
    KernelImpl* impl;
};

class KernelFactory {
public:
    virtual ~KernelFactory() {}
    virtual KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const = 0;
};

class Platform {
public:
    explicit Platform(const std::string& name) : name(name) {}
    virtual ~Platform() {}
    const std::string& getName() const { return name; }
    void registerKernelFactory(const std::string& kernel, KernelFactory* f) { factories[kernel] = f; }
    Kernel createKernel(const std::string& kernel, ContextImpl& context) const {
        std::map<std::string, KernelFactory*>::const_iterator it = factories.find(kernel);
        if (it == factories.end()) throw OpenMMException("Called createKernel() on a Platform which does not support the requested kernel");
        return Kernel(it->second->createKernelImpl(kernel, *this, context));
    }
    static Platform& getPlatformByName(const std::string& name);
    static void registerPlatform(Platform* p);
private:
    std::string name;
    std::map<std::string, KernelFactory*> factories;
};

class Force {
public:
    Force() : group(0) {}
    virtual ~Force() {}
    int getForceGroup() const { return group; }
    virtual bool usesPeriodicBoundaryConditions() const { return false; }
protected:
    friend class ContextImpl;
    virtual ForceImpl* createImpl() const = 0;
    ForceImpl& getImplInContext(Context& context);
    ContextImpl& getContextImpl(Context& context);
private:
    int group;
};

class ForceImpl {
public:
    virtual ~ForceImpl() {}
    virtual void initialize(ContextImpl& context) = 0;
    virtual const Force& getOwner() const = 0;
    virtual void updateContextState(ContextImpl& context) {}
    virtual double calcForcesAndEnergy(ContextImpl& context, bool includeForces, bool includeEnergy, int groups) = 0;
    virtual std::map<std::string, double> getDefaultParameters() { return std::map<std::string, double>(); }
    virtual std::vector<std::string> getKernelNames() = 0;
};

class CudaContext;

class ContextImpl {
public:
    ContextImpl(Context& owner, const System& system, Platform& platform, int device, const std::string& precision);
    ~ContextImpl();
    const System& getSystem() const { return *system; }
    Platform& getPlatform() { return *platform; }
    void* getPlatformData() { return data; }                // CudaPlatform::PlatformData*
    CudaContext& getCudaContext();
    Context& getOwner() { return *owner; }
    double calcForcesAndEnergy(bool includeForces, bool includeEnergy);
    ForceImpl& getImpl(const Force* f);
    std::vector<Vec3> positions, forces;
private:
    Context* owner;
    const System* system;
    Platform* platform;
    void* data;
    std::vector<ForceImpl*> impls;
};

class Context {
public:
    // precision: the CUDA platform's "Precision" property ("single" | "mixed" | "double")
    Context(const System& system, Platform& platform, int device = 0, const std::string& precision = "single")
        : impl(new ContextImpl(*this, system, platform, device, precision)) {}
    ~Context() { delete impl; }
    void setPositions(const std::vector<Vec3>& p);
    double getPotentialEnergy() { return impl->calcForcesAndEnergy(true, true); }     // State::Energy | State::Forces
    const std::vector<Vec3>& getForces() const { return impl->forces; }
    ContextImpl& getImpl() { return *impl; }
private:
    ContextImpl* impl;
};

} // namespace OpenMM
#endif
