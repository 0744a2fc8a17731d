// Triangle3D shim: the three accessors Intersections3D::intersect(Triangle3D&, AABB&) uses (SRC/Geometry/3D/Triangle3D.h:113-123).
#pragma once
#include "stdafx.h"
class Triangle3D {
public:
    Triangle3D(const vec3& a, const vec3& b, const vec3& c) : _a(a), _b(b), _c(c) {}
    vec3 getP1() const { return _a; }
    vec3 getP2() const { return _b; }
    vec3 getP3() const { return _c; }
private:
    vec3 _a, _b, _c;
};
