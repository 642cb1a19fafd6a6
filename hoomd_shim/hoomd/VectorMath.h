// hoomd_shim/hoomd/VectorMath.h -- the vec3<Real> operations the plugin sources use
// (constructor from Scalar3/Scalar4, +, -, +=, scalar*vec3, dot).  See ShimCore.h.
#ifndef HOOMD_SHIM_VECTORMATH_H
#define HOOMD_SHIM_VECTORMATH_H
#include "ShimCore.h"

namespace hoomd
    {
template<class Real> struct vec3
    {
    vec3() : x(0), y(0), z(0) { }
    vec3(const Real& _x, const Real& _y, const Real& _z) : x(_x), y(_y), z(_z) { }
    explicit vec3(const Scalar3& a) : x(a.x), y(a.y), z(a.z) { }
    explicit vec3(const Scalar4& a) : x(a.x), y(a.y), z(a.z) { }
    Real x, y, z;
    };

template<class Real> inline vec3<Real> operator+(const vec3<Real>& a, const vec3<Real>& b)
    {
    return vec3<Real>(a.x + b.x, a.y + b.y, a.z + b.z);
    }
template<class Real> inline vec3<Real> operator-(const vec3<Real>& a, const vec3<Real>& b)
    {
    return vec3<Real>(a.x - b.x, a.y - b.y, a.z - b.z);
    }
template<class Real> inline vec3<Real> operator-(const vec3<Real>& a)
    {
    return vec3<Real>(-a.x, -a.y, -a.z);
    }
template<class Real> inline vec3<Real>& operator+=(vec3<Real>& a, const vec3<Real>& b)
    {
    a.x += b.x;
    a.y += b.y;
    a.z += b.z;
    return a;
    }
template<class Real> inline vec3<Real> operator*(const Real& s, const vec3<Real>& a)
    {
    return vec3<Real>(s * a.x, s * a.y, s * a.z);
    }
template<class Real> inline vec3<Real> operator*(const vec3<Real>& a, const Real& s)
    {
    return vec3<Real>(a.x * s, a.y * s, a.z * s);
    }
template<class Real> inline Real dot(const vec3<Real>& a, const vec3<Real>& b)
    {
    return a.x * b.x + a.y * b.y + a.z * b.z;
    }
    } // namespace hoomd
#endif
