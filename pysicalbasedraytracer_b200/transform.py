"""Host-side mirror of the reference's Transform / camera-matrix construction, in IEEE float32.

Follows Core/Transform.cpp (Inverse :46-112, Translate :130-137, Scale :138-143, LookAt :188-215,
Perspective :231-236, Orthographic :228-230, operator* :126-129) and the ProjectiveCamera ctor
(Camera/Camera.h:34-54) operation by operation so the resulting RasterToCamera / CameraToWorld
matrices are bit-identical to the reference's (checked in tests/test_camera_parity.py against the
compiled reference).  Pure host logic: tiny 4x4 matrices, executed once per render.
"""
import numpy as np

F = np.float32


def _mul(a, b):
    # Matrix4x4::Mul, Transform.h:38-46 (left-to-right float sums)
    r = np.zeros((4, 4), dtype=F)
    for i in range(4):
        for j in range(4):
            r[i, j] = F(F(F(a[i, 0] * b[0, j]) + F(a[i, 1] * b[1, j])) + F(a[i, 2] * b[2, j])) + F(a[i, 3] * b[3, j])
    return r


def _inverse(m):
    # Gauss-Jordan with full pivoting, Transform.cpp:46-112.  `pivinv = 1. / x` is a DOUBLE divide
    # narrowed to float (:88).
    indxc = [0] * 4
    indxr = [0] * 4
    ipiv = [0] * 4
    minv = np.array(m, dtype=F).copy()
    for i in range(4):
        irow = icol = 0
        big = F(0)
        for j in range(4):
            if ipiv[j] != 1:
                for k in range(4):
                    if ipiv[k] == 0:
                        if abs(minv[j, k]) >= big:
                            big = F(abs(minv[j, k]))
                            irow, icol = j, k
        ipiv[icol] += 1
        if irow != icol:
            minv[[irow, icol], :] = minv[[icol, irow], :]
        indxr[i], indxc[i] = irow, icol
        pivinv = F(np.float64(1.0) / np.float64(minv[icol, icol]))
        minv[icol, icol] = F(1)
        for j in range(4):
            minv[icol, j] = F(minv[icol, j] * pivinv)
        for j in range(4):
            if j != icol:
                save = minv[j, icol]
                minv[j, icol] = F(0)
                for k in range(4):
                    minv[j, k] = F(minv[j, k] - F(minv[icol, k] * save))
    for j in range(3, -1, -1):
        if indxr[j] != indxc[j]:
            minv[:, [indxr[j], indxc[j]]] = minv[:, [indxc[j], indxr[j]]]
    return minv


class Transform:
    def __init__(self, m=None, minv=None):
        self.m = np.eye(4, dtype=F) if m is None else np.array(m, dtype=F)
        self.minv = _inverse(self.m) if minv is None else np.array(minv, dtype=F)

    def __mul__(self, t2):  # Transform.cpp:126-129
        return Transform(_mul(self.m, t2.m), _mul(t2.minv, self.minv))

    def inverse(self):
        return Transform(self.minv, self.m)

    def point(self, p):  # Transform.h:138-150
        x, y, z = (F(v) for v in p)
        m = self.m
        r = [F(F(F(m[i, 0] * x) + F(m[i, 1] * y)) + F(m[i, 2] * z)) + m[i, 3] for i in range(4)]
        if r[3] == F(1):
            return np.array(r[:3], dtype=F)
        inv = F(F(1) / r[3])
        return np.array([F(inv * r[0]), F(inv * r[1]), F(inv * r[2])], dtype=F)


def translate(d):
    d = [F(v) for v in d]
    m = np.eye(4, dtype=F)
    mi = np.eye(4, dtype=F)
    for i in range(3):
        m[i, 3] = d[i]
        mi[i, 3] = -d[i]
    return Transform(m, mi)


def scale(x, y, z):
    x, y, z = F(x), F(y), F(z)
    m = np.diag([x, y, z, F(1)]).astype(F)
    mi = np.diag([F(F(1) / x), F(F(1) / y), F(F(1) / z), F(1)]).astype(F)
    return Transform(m, mi)


def _cross(a, b):  # Geometry.h:706-714: products and differences in double, narrowed to float
    a = [np.float64(v) for v in a]
    b = [np.float64(v) for v in b]
    return np.array([F(a[1] * b[2] - a[2] * b[1]), F(a[2] * b[0] - a[0] * b[2]), F(a[0] * b[1] - a[1] * b[0])], dtype=F)


def _length(v):
    return F(np.sqrt(F(F(F(v[0] * v[0]) + F(v[1] * v[1])) + F(v[2] * v[2]))))


def _normalize(v):  # v / Length(): multiplies by 1.f/len (Geometry.h:223-228,737-740)
    inv = F(F(1) / _length(v))
    return np.array([F(v[0] * inv), F(v[1] * inv), F(v[2] * inv)], dtype=F)


def look_at(pos, look, up):  # Transform.cpp:188-215 ; returns the WORLD->CAMERA transform
    pos = np.array(pos, dtype=F)
    look = np.array(look, dtype=F)
    up = np.array(up, dtype=F)
    c2w = np.eye(4, dtype=F)
    c2w[0, 3], c2w[1, 3], c2w[2, 3] = pos
    d = _normalize(np.array([F(look[i] - pos[i]) for i in range(3)], dtype=F))
    nup = _normalize(up)
    if _length(_cross(nup, d)) == 0:
        return Transform()
    right = _normalize(_cross(nup, d))
    new_up = _cross(d, right)
    c2w[0:3, 0] = right
    c2w[0:3, 1] = new_up
    c2w[0:3, 2] = d
    return Transform(_inverse(c2w), c2w)


def perspective(fov, n, f):  # Transform.cpp:231-236
    fov, n, f = F(fov), F(n), F(f)
    persp = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, F(f / F(f - n)), F(F(-f * n) / F(f - n))], [0, 0, 1, 0]], dtype=F)
    rad = F(F(F(np.float32(3.14159265358979323846)) / F(180)) * fov)
    # std::tan(float) -> tanf; evaluated in double and rounded once (glibc tanf is correctly rounded
    # for this argument; the parity test pins it).
    inv_tan = F(F(1) / F(np.tan(np.float64(F(rad / F(2))))))
    return scale(inv_tan, inv_tan, 1) * Transform(persp)


def orthographic(znear, zfar):  # Transform.cpp:228-230
    return scale(1, 1, F(F(1) / F(F(zfar) - F(znear)))) * translate([0, 0, -F(znear)])


def screen_window(width, height, ortho=False):
    # CreatePerspectiveCamera Perspective.cpp:84-104 ; CreateOrthographicCamera Orthographic.cpp:65-98
    frame = F(F(width) / F(height))
    if frame > F(1):
        w = [-frame, frame, F(-1), F(1)]
    else:
        w = [F(-1), F(1), F(F(-1) / frame), F(F(1) / frame)]
    if ortho:
        w = [F(v * F(2)) for v in w]
    return w  # pMin.x, pMax.x, pMin.y, pMax.y


def raster_to_camera(width, height, camera_to_screen, window):
    # ProjectiveCamera ctor, Camera/Camera.h:41-49
    pminx, pmaxx, pminy, pmaxy = window
    s2r = scale(width, height, 1) * scale(F(F(1) / F(pmaxx - pminx)), F(F(1) / F(pminy - pmaxy)), 1) * translate(
        [-pminx, -pmaxy, 0])
    r2s = s2r.inverse()
    return camera_to_screen.inverse() * r2s
