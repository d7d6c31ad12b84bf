// Host float math with the SAME operation order as the cyCodeBase types the reference uses,
// so that everything computed at load time (node transforms, camera frame, mesh normals,
// BVH boxes) is bit-identical to what xmlload.cpp leaves in the reference's globals.
// Compiled with -ffp-contract=off: no FMA contraction, like the x86-64 reference build.
//   V3  <-> cyPoint3f  (cyPoint.h:292-296, 346-349: Sum() = x+y+z, Dot = (x*x'+y*y')+z*z')
//   M3  <-> cyMatrix3f (cyMatrix.h:290-294 column-major; operator* :530-548; GetInverse :612-633)
#pragma once
#include <cmath>

namespace rtu {

struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    float &operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};

inline V3 operator+(const V3 &a, const V3 &b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(const V3 &a, const V3 &b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(const V3 &a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(const V3 &a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, const V3 &a) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(const V3 &a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
inline float dot(const V3 &a, const V3 &b)
{
    float px = a.x * b.x, py = a.y * b.y, pz = a.z * b.z;
    return px + py + pz; // (px+py)+pz, cyPoint.h:296
}
inline V3 cross(const V3 &a, const V3 &b)
{
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); // cyPoint.h:346
}
inline float length(const V3 &a) { return sqrtf(dot(a, a)); }
inline V3 normalized(const V3 &a) { return a / length(a); } // component-wise division, cyPoint.h:295

struct M3 {
    float d[9]; // | 0 3 6 |  | 1 4 7 |  | 2 5 8 |
    void identity()
    {
        for (int i = 0; i < 9; i++) d[i] = 0.f;
        d[0] = d[4] = d[8] = 1.f;
    }
    void zero()
    {
        for (int i = 0; i < 9; i++) d[i] = 0.f;
    }
};

// cyMatrix.h:530-541: per output column, a[j]+b[j]+c[j] with a from column 0 of the left matrix
inline M3 mul(const M3 &l, const M3 &r)
{
    M3 o;
    for (int i = 0; i < 9; i += 3)
        for (int j = 0; j < 3; j++) {
            float a = l.d[j] * r.d[i], b = l.d[3 + j] * r.d[i + 1], c = l.d[6 + j] * r.d[i + 2];
            o.d[i + j] = a + b + c;
        }
    return o;
}
// cyMatrix.h:542-547
inline V3 mul(const M3 &m, const V3 &p)
{
    return V3(p.x * m.d[0] + p.y * m.d[3] + p.z * m.d[6],
              p.x * m.d[1] + p.y * m.d[4] + p.z * m.d[7],
              p.x * m.d[2] + p.y * m.d[5] + p.z * m.d[8]);
}
// cyMatrix.h:612-633: adjugate, det from the first row of the source times first column of adj, then /=
inline M3 inverse(const M3 &m)
{
    const float *s = m.d;
    M3 o;
    o.d[0] = s[4] * s[8] - s[5] * s[7];
    o.d[1] = s[2] * s[7] - s[1] * s[8];
    o.d[2] = s[1] * s[5] - s[2] * s[4];
    o.d[3] = s[5] * s[6] - s[3] * s[8];
    o.d[4] = s[0] * s[8] - s[2] * s[6];
    o.d[5] = s[2] * s[3] - s[0] * s[5];
    o.d[6] = s[3] * s[7] - s[4] * s[6];
    o.d[7] = s[1] * s[6] - s[0] * s[7];
    o.d[8] = s[0] * s[4] - s[1] * s[3];
    float det = s[0] * o.d[0] + s[1] * o.d[3] + s[2] * o.d[6];
    for (int i = 0; i < 9; i++) o.d[i] /= det;
    return o;
}
// cyMatrix.h:412-430 SetRotation(axis, angle)
inline M3 rotation(const V3 &axis, float angle)
{
    float sinA = sinf(angle), cosA = cosf(angle);
    float t = 1.0f - cosA;
    float tx = t * axis.x, ty = t * axis.y, tz = t * axis.z;
    float txy = tx * axis.y, txz = tx * axis.z, tyz = ty * axis.z;
    float sx = sinA * axis.x, sy = sinA * axis.y, sz = sinA * axis.z;
    M3 o;
    o.d[0] = tx * axis.x + cosA; o.d[1] = txy + sz;           o.d[2] = txz - sy;
    o.d[3] = txy - sz;           o.d[4] = ty * axis.y + cosA; o.d[5] = tyz + sx;
    o.d[6] = txz + sy;           o.d[7] = tyz - sx;           o.d[8] = tz * axis.z + cosA;
    return o;
}

// Transformation (scene.h:223-261)
struct Xform {
    M3 tm, itm;
    V3 pos;
    Xform()
    {
        tm.identity();
        itm.identity();
    }
    void transform(const M3 &m) // scene.h:247
    {
        tm = mul(m, tm);
        pos = mul(m, pos);
        itm = inverse(tm);
    }
    void translate(const V3 &p) { pos = pos + p; }                 // scene.h:244
    void rotate(const V3 &axis, float degree)                       // scene.h:245
    {
        transform(rotation(axis, degree * (float)M_PI / 180.0f));
    }
    void scale(float sx, float sy, float sz)                        // scene.h:246
    {
        M3 m;
        m.zero();
        m.d[0] = sx; m.d[4] = sy; m.d[8] = sz;
        transform(m);
    }
};

} // namespace rtu
