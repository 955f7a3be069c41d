// Oracle shim: 3-vector of doubles with the operators the reference path uses.
#ifndef ORACLE_SHIM_VEC3_H_
#define ORACLE_SHIM_VEC3_H_
#include <cmath>
namespace OpenMM {
class Vec3 {
public:
    Vec3() { d[0] = d[1] = d[2] = 0.0; }
    Vec3(double x, double y, double z) { d[0] = x; d[1] = y; d[2] = z; }
    double operator[](int i) const { return d[i]; }
    double& operator[](int i) { return d[i]; }
    Vec3 operator+() const { return *this; }
    Vec3 operator-() const { return Vec3(-d[0], -d[1], -d[2]); }
    Vec3 operator+(const Vec3& o) const { return Vec3(d[0]+o.d[0], d[1]+o.d[1], d[2]+o.d[2]); }
    Vec3 operator-(const Vec3& o) const { return Vec3(d[0]-o.d[0], d[1]-o.d[1], d[2]-o.d[2]); }
    Vec3& operator+=(const Vec3& o) { d[0]+=o.d[0]; d[1]+=o.d[1]; d[2]+=o.d[2]; return *this; }
    Vec3& operator-=(const Vec3& o) { d[0]-=o.d[0]; d[1]-=o.d[1]; d[2]-=o.d[2]; return *this; }
    Vec3 operator*(double s) const { return Vec3(d[0]*s, d[1]*s, d[2]*s); }
    Vec3& operator*=(double s) { d[0]*=s; d[1]*=s; d[2]*=s; return *this; }
    Vec3 operator/(double s) const { double inv = 1.0/s; return Vec3(d[0]*inv, d[1]*inv, d[2]*inv); }
    Vec3& operator/=(double s) { double inv = 1.0/s; d[0]*=inv; d[1]*=inv; d[2]*=inv; return *this; }
    double dot(const Vec3& o) const { return d[0]*o.d[0] + d[1]*o.d[1] + d[2]*o.d[2]; }
    Vec3 cross(const Vec3& o) const {
        return Vec3(d[1]*o.d[2]-d[2]*o.d[1], d[2]*o.d[0]-d[0]*o.d[2], d[0]*o.d[1]-d[1]*o.d[0]);
    }
private:
    double d[3];
};
static inline Vec3 operator*(double s, const Vec3& v) { return v*s; }
}
#endif
