"""goref.py — a SECOND, independent restatement of the reference's CPU render path, in plain Python (TEST INFRASTRUCTURE ONLY).

Why it exists: the reference (MarkJulian19/path_trace_golang) cannot be run here (no Go toolchain) and ships no golden
vectors, so the C++ oracle (oracle/oracle.cpp) is "parity unpinned".  This module was written from the Go sources alone,
function by function, statement by statement, WITHOUT looking at oracle.cpp, with Python floats (IEEE binary64, no fused
multiply-add — what Go's gc compiler emits on amd64) and the expression order of the Go code.  tests/test_oracle_crosscheck.py
then demands that the two restatements agree BIT FOR BIT on converted worlds, cameras, primary hits, per-pixel radiance sums
and event counters.  Two independent readings of the same source that agree to the last bit is the strongest statement
about the oracle that can be made without executing the reference; any disagreement points at a line to re-read.

What it cannot settle (same caveat as the oracle): Go's pure-Go math.Tan / Sin / Cos / Exp / Pow may differ from the C
library's by an ulp; both restatements call the C library.  The random source is the shared counter hash
(DESIGN.md "RNG") standing in for randSource.Float64 (random.go:27-34), whose stream is time-seeded in the reference.

Only tests may import this module.  It is slow (pure-Python loops): use frames of a few hundred pixels.

Every function cites the reference lines it follows (internal/engine/*.go, internal/scene/scene.go).
"""
from __future__ import annotations

import math

MAXF = 1.7976931348623157e308           # math.MaxFloat64 (renderer.go:294, 326)

# materialType (materials.go:11-17)
LAMBERT, METAL, DIELECTRIC, EMISSIVE, MIRROR = 0, 1, 2, 3, 4
# world entry kinds (objects.go:31, 92, 136)
SPHERE, PLANE, BOX = 0, 1, 2


# ------------------------------------------------------------------ Go's math.Min / math.Max (special cases of the Go spec)
def go_min(x, y):
    if x == -math.inf or y == -math.inf:
        return -math.inf
    if x != x or y != y:
        return math.nan
    if x == 0 and x == y:
        return x if math.copysign(1.0, x) < 0 else y
    return x if x < y else y


def go_max(x, y):
    if x == math.inf or y == math.inf:
        return math.inf
    if x != x or y != y:
        return math.nan
    if x == 0 and x == y:
        return y if math.copysign(1.0, x) < 0 else x
    return x if x > y else y


def _fdiv(x, y):
    """Go's x / y on float64: a zero divisor gives +-Inf or NaN, never a panic (Python would raise ZeroDivisionError)."""
    if y == 0:
        if x == 0 or x != x:
            return math.nan
        return math.copysign(math.inf, x) * math.copysign(1.0, y)
    return x / y


def _recip(x):
    return _fdiv(1.0, x)


# ------------------------------------------------------------------ vec3 (math.go:5-37); a vec3 is a tuple (x, y, z)
def add(a, b): return (a[0] + b[0], a[1] + b[1], a[2] + b[2])                     # math.go:11
def sub(a, b): return (a[0] - b[0], a[1] - b[1], a[2] - b[2])                     # math.go:12
def mul(a, t): return (a[0] * t, a[1] * t, a[2] * t)                              # math.go:13


def div(a, t):                                                                    # math.go:14-17
    inv_t = 1.0 / t
    return (a[0] * inv_t, a[1] * inv_t, a[2] * inv_t)


def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]                     # math.go:19


def cross(a, b):                                                                  # math.go:21-27
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def length(a): return math.sqrt(dot(a, a))                                        # math.go:29


def unit(a):                                                                      # math.go:31-37
    l = length(a)
    if l == 0:
        return a
    return div(a, l)


def reflect_vec(v, n):                                                            # math.go:39-46
    d = dot(v, n)
    return (v[0] - n[0] * 2 * d, v[1] - n[1] * 2 * d, v[2] - n[2] * 2 * d)


def refract_vec(uv, n, etai_over_etat):                                           # math.go:48-64
    cos_theta = go_min(-uv[0] * n[0] - uv[1] * n[1] - uv[2] * n[2], 1.0)
    px = uv[0] + n[0] * cos_theta
    py = uv[1] + n[1] * cos_theta
    pz = uv[2] + n[2] * cos_theta
    px *= etai_over_etat
    py *= etai_over_etat
    pz *= etai_over_etat
    perp_len_sq = px * px + py * py + pz * pz
    par = -math.sqrt(abs(1.0 - perp_len_sq))
    return (px + n[0] * par, py + n[1] * par, pz + n[2] * par)


# ------------------------------------------------------------------ random source: the shared counter hash (DESIGN.md "RNG")
def _fmix(x):
    x ^= x >> 16
    x = (x * 0x21F0AAAD) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x735A2D97) & 0xFFFFFFFF
    x ^= x >> 15
    return x


class Rng:
    """Stands in for randSource (random.go:10-34): Float64() is uniform on [0, 1)."""
    __slots__ = ("key", "ctr")

    def __init__(self, seed, pixel, sample):
        k = _fmix((seed ^ 0x9E3779B9) & 0xFFFFFFFF)
        k = _fmix(k ^ (pixel & 0xFFFFFFFF))
        k = _fmix((k + sample * 0x9E3779B9) & 0xFFFFFFFF)
        self.key, self.ctr = k, 0

    def float64(self):
        x = _fmix((self.key + self.ctr * 0x9E3779B9) & 0xFFFFFFFF)
        self.ctr += 1
        return (x >> 8) * (1.0 / 16777216.0)


def random_in_unit_sphere(rng):                                                   # math.go:66-85
    while True:
        x = rng.float64() * 2 - 1
        y = rng.float64() * 2 - 1
        z = rng.float64() * 2 - 1
        len_sq = x * x + y * y + z * z
        if len_sq >= 1.0:
            continue
        return (x, y, z)


def random_cosine_direction(normal, rng):                                         # math.go:94-131
    r1 = rng.float64()
    r2 = rng.float64()
    phi = 2.0 * math.pi * r1
    cos_theta = math.sqrt(r2)
    sin_theta = math.sqrt(1.0 - r2)
    if abs(normal[0]) > 0.9:
        u = (0.0, 1.0, 0.0)
    else:
        u = (1.0, 0.0, 0.0)
    w = normal
    v_vec = unit(cross(w, u))
    u_vec = cross(v_vec, w)
    lx = sin_theta * math.cos(phi)
    ly = sin_theta * math.sin(phi)
    lz = cos_theta
    return (lx * u_vec[0] + ly * v_vec[0] + lz * w[0],
            lx * u_vec[1] + ly * v_vec[1] + lz * w[1],
            lx * u_vec[2] + ly * v_vec[2] + lz * w[2])


# ------------------------------------------------------------------ materials (materials.go)
class Material:
    __slots__ = ("typ", "albedo", "rough", "ior", "emit", "absorption")

    def __init__(self, typ=LAMBERT, albedo=(0.0, 0.0, 0.0), rough=0.0, ior=0.0, emit=(0.0, 0.0, 0.0), absorption=(0.0, 0.0, 0.0)):
        self.typ, self.albedo, self.rough, self.ior, self.emit, self.absorption = typ, albedo, rough, ior, emit, absorption


def clamp(x, lo, hi):                                                             # materials.go:57-65
    if x < lo:
        return lo
    if x > hi:
        return hi
    return x


def _num(d, k):
    """A JSON number field of a decoded object; a missing key or null leaves Go's zero value."""
    v = (d or {}).get(k, 0)
    return float(v or 0)


def _col(d): return (_num(d, "r"), _num(d, "g"), _num(d, "b"))
def _vec(d): return (_num(d, "x"), _num(d, "y"), _num(d, "z"))


def convert_material(m):                                                          # materials.go:28-55
    al = _col(m.get("albedo"))
    e, power = _col(m.get("emit")), _num(m, "power")
    em = (e[0] * power, e[1] * power, e[2] * power)
    ab = _col(m.get("absorption"))
    t = m.get("type")
    if t == "metal":                                                              # :34-40
        rough = _num(m, "rough")
        if _num(m, "smoothness") > 0:
            rough = 1.0 - clamp(_num(m, "smoothness"), 0, 1)
        return Material(METAL, albedo=al, rough=clamp(rough, 0, 1))
    if t == "dielectric":                                                         # :41-46
        ior = _num(m, "ior")
        if ior == 0:
            ior = 1.5
        return Material(DIELECTRIC, albedo=al, ior=ior, absorption=ab)
    if t == "emissive":                                                           # :47-48
        return Material(EMISSIVE, emit=em)
    if t == "mirror":                                                             # :49-50
        return Material(MIRROR, albedo=al)
    return Material(LAMBERT, albedo=al, rough=clamp(_num(m, "rough"), 0, 1))      # :51-54


def emitted(m):                                                                   # materials.go:67-72
    if m.typ == EMISSIVE:
        return m.emit
    return (0.0, 0.0, 0.0)


def reflectance(cosine, ref_idx):                                                 # materials.go:226-231
    r0 = (1 - ref_idx) / (1 + ref_idx)
    r0 = r0 * r0
    return r0 + (1 - r0) * math.pow(1 - cosine, 5)


class Hit:
    """hitRecord (objects.go:9-15) + the world index of the object that filled it (for the comparison with the oracle)."""
    __slots__ = ("p", "normal", "t", "front_face", "mat", "index")

    def __init__(self):
        self.p = (0.0, 0.0, 0.0)
        self.normal = (0.0, 0.0, 0.0)
        self.t = 0.0
        self.front_face = False
        self.mat = None
        self.index = -1


def scatter(m, rng, r_orig, r_dir, rec):                                          # materials.go:74-224
    """Returns (ok, attenuation, scattered origin, scattered direction)."""
    zero = (0.0, 0.0, 0.0)
    if m.typ == LAMBERT:                                                          # :76-97
        sd = random_cosine_direction(rec.normal, rng)
        if m.rough > 1e-6:
            off = random_in_unit_sphere(rng)
            sd = (sd[0] + off[0] * m.rough * 0.1, sd[1] + off[1] * m.rough * 0.1, sd[2] + off[2] * m.rough * 0.1)
            sd = unit(sd)
        return True, m.albedo, rec.p, sd
    if m.typ == METAL:                                                            # :99-160
        dir_len = math.sqrt(r_dir[0] * r_dir[0] + r_dir[1] * r_dir[1] + r_dir[2] * r_dir[2])
        if dir_len == 0:
            return False, zero, rec.p, r_dir
        inv_len = 1.0 / dir_len
        ud = (r_dir[0] * inv_len, r_dir[1] * inv_len, r_dir[2] * inv_len)
        reflected = reflect_vec(ud, rec.normal)
        if m.rough > 1e-6:
            sd = random_cosine_direction(reflected, rng)
            alpha = m.rough * m.rough
            sx = reflected[0] * (1.0 - alpha) + sd[0] * alpha
            sy = reflected[1] * (1.0 - alpha) + sd[1] * alpha
            sz = reflected[2] * (1.0 - alpha) + sd[2] * alpha
            len_sq = sx * sx + sy * sy + sz * sz
            if len_sq < 1e-8:
                sx, sy, sz = reflected
            else:
                inv = 1.0 / math.sqrt(len_sq)
                sx *= inv
                sy *= inv
                sz *= inv
            d = sx * rec.normal[0] + sy * rec.normal[1] + sz * rec.normal[2]
            if d <= 0:
                sx, sy, sz = reflected
            return True, m.albedo, rec.p, (sx, sy, sz)
        return True, m.albedo, rec.p, reflected
    if m.typ == DIELECTRIC:                                                       # :162-200
        attenuation = (1.0, 1.0, 1.0)
        ratio = 1.0 / m.ior if rec.front_face else m.ior
        dir_len = math.sqrt(r_dir[0] * r_dir[0] + r_dir[1] * r_dir[1] + r_dir[2] * r_dir[2])
        if dir_len == 0:
            return False, attenuation, rec.p, r_dir
        inv_len = 1.0 / dir_len
        ud = (r_dir[0] * inv_len, r_dir[1] * inv_len, r_dir[2] * inv_len)
        cos_theta = go_min(-(ud[0] * rec.normal[0] + ud[1] * rec.normal[1] + ud[2] * rec.normal[2]), 1.0)
        sin_theta = math.sqrt(1.0 - cos_theta * cos_theta)
        cannot_refract = ratio * sin_theta > 1.0
        reflect_prob = reflectance(cos_theta, ratio)
        if cannot_refract or reflect_prob > rng.float64():                        # `||` short-circuits: no draw when cannot_refract
            direction = reflect_vec(ud, rec.normal)
        else:
            direction = refract_vec(ud, rec.normal, ratio)
        return True, attenuation, rec.p, direction
    if m.typ == EMISSIVE:                                                         # :202-203
        return False, zero, zero, zero
    if m.typ == MIRROR:                                                           # :205-221
        dir_len = math.sqrt(r_dir[0] * r_dir[0] + r_dir[1] * r_dir[1] + r_dir[2] * r_dir[2])
        if dir_len == 0:
            return False, zero, rec.p, r_dir
        inv_len = 1.0 / dir_len
        ud = (r_dir[0] * inv_len, r_dir[1] * inv_len, r_dir[2] * inv_len)
        return True, m.albedo, rec.p, reflect_vec(ud, rec.normal)
    return False, zero, zero, zero


# ------------------------------------------------------------------ primitives (objects.go); a world entry is (kind, a, b, material)
def hit_sphere(center, radius, mat, o, d, t_min, t_max, rec):                     # objects.go:37-89
    ocx = o[0] - center[0]
    ocy = o[1] - center[1]
    ocz = o[2] - center[2]
    a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
    half_b = ocx * d[0] + ocy * d[1] + ocz * d[2]
    oc_len_sq = ocx * ocx + ocy * ocy + ocz * ocz
    radius_sq = radius * radius
    c = oc_len_sq - radius_sq
    disc = half_b * half_b - a * c
    if disc < 0:
        return False
    sqrt_d = math.sqrt(disc)
    root = _fdiv(-half_b - sqrt_d, a)
    if root < t_min or root > t_max:
        root = _fdiv(-half_b + sqrt_d, a)
        if root < t_min or root > t_max:
            return False
    rec.t = root
    px = o[0] + d[0] * root
    py = o[1] + d[1] * root
    pz = o[2] + d[2] * root
    rec.p = (px, py, pz)
    inv_radius = _recip(radius)
    nx = (px - center[0]) * inv_radius
    ny = (py - center[1]) * inv_radius
    nz = (pz - center[2]) * inv_radius
    dt = d[0] * nx + d[1] * ny + d[2] * nz
    rec.front_face = dt < 0
    rec.normal = (nx, ny, nz) if rec.front_face else (-nx, -ny, -nz)
    rec.mat = mat
    return True


def hit_plane(point, normal, mat, o, d, t_min, t_max, rec):                       # objects.go:98-133
    denom = normal[0] * d[0] + normal[1] * d[1] + normal[2] * d[2]
    if abs(denom) < 1e-6:
        return False
    mx = point[0] - o[0]
    my = point[1] - o[1]
    mz = point[2] - o[2]
    t = (mx * normal[0] + my * normal[1] + mz * normal[2]) / denom
    if t < t_min or t > t_max:
        return False
    rec.t = t
    rec.p = (o[0] + d[0] * t, o[1] + d[1] * t, o[2] + d[2] * t)
    rec.front_face = denom < 0
    rec.normal = (normal[0], normal[1], normal[2]) if rec.front_face else (-normal[0], -normal[1], -normal[2])
    rec.mat = mat
    return True




def hit_box(bmin, bmax, mat, o, d, t_min, t_max, rec):                            # objects.go:141-222
    t0 = t_min
    t1 = t_max
    for i in range(3):
        inv_d = _recip(d[i])
        orig = o[i]
        min_v = bmin[i]
        max_v = bmax[i]
        t_near = (min_v - orig) * inv_d
        t_far = (max_v - orig) * inv_d
        if inv_d < 0:
            t_near, t_far = t_far, t_near
        if t_near > t0:
            t0 = t_near
        if t_far < t1:
            t1 = t_far
        if t1 <= t0:
            return False
    rec.t = t0
    p = add(o, mul(d, t0))                                                        # r.at(t0), math.go:138-140
    rec.p = p
    dx_min = p[0] - bmin[0]
    dx_max = bmax[0] - p[0]
    dy_min = p[1] - bmin[1]
    dy_max = bmax[1] - p[1]
    dz_min = p[2] - bmin[2]
    dz_max = bmax[2] - p[2]
    min_dist = dx_min
    n = (-1.0, 0.0, 0.0)
    if dx_max < min_dist:
        min_dist = dx_max
        n = (1.0, 0.0, 0.0)
    if dy_min < min_dist:
        min_dist = dy_min
        n = (0.0, -1.0, 0.0)
    if dy_max < min_dist:
        min_dist = dy_max
        n = (0.0, 1.0, 0.0)
    if dz_min < min_dist:
        min_dist = dz_min
        n = (0.0, 0.0, -1.0)
    if dz_max < min_dist:
        n = (0.0, 0.0, 1.0)
    rec.front_face = dot(d, n) < 0                                                # setFaceNormal, objects.go:17-24
    rec.normal = n if rec.front_face else mul(n, -1)
    rec.mat = mat
    return True


def hit_object(ob, o, d, t_min, t_max, rec):
    kind, a, b, mat = ob
    if kind == SPHERE:
        return hit_sphere(a, b, mat, o, d, t_min, t_max, rec)
    if kind == PLANE:
        return hit_plane(a, b, mat, o, d, t_min, t_max, rec)
    return hit_box(a, b, mat, o, d, t_min, t_max, rec)


def scene_to_world(sc):                                                           # objects.go:225-269
    materials = {}
    for m in sc.get("materials") or []:
        materials[m.get("id") or ""] = convert_material(m)                        # a later duplicate id wins (map assignment)
    world = []
    for ob in sc.get("objects") or []:
        mat = materials.get(ob.get("material_id") or "", None) or Material()      # a missing id gives the zero material
        pos = _vec(ob.get("position"))
        size = _vec(ob.get("size"))
        t = ob.get("type")
        if t == "sphere" or t == "sphere_light":                                  # :238-250
            world.append((SPHERE, pos, size[0], mat))
        elif t == "plane":                                                        # :251-257
            world.append((PLANE, pos, (0.0, 1.0, 0.0), mat))
        elif t == "box":                                                          # :258-265
            world.append((BOX, sub(pos, mul(size, 0.5)), add(pos, mul(size, 0.5)), mat))
    return world


# ------------------------------------------------------------------ camera (camera.go)
class Camera:
    __slots__ = ("origin", "llc", "horizontal", "vertical", "u", "v", "w", "lens_radius")


def new_camera(cam, width, height):                                               # camera.go:19-58
    cam = cam or {}
    aspect = float(width) / float(height)
    if _num(cam, "aspect_ratio") != 0:
        aspect = _num(cam, "aspect_ratio")
    theta = _num(cam, "fov") * math.pi / 180
    h = math.tan(theta / 2)
    viewport_height = 2.0 * h
    viewport_width = aspect * viewport_height
    origin = _vec(cam.get("position"))
    target = _vec(cam.get("target"))
    up = _vec(cam.get("up"))
    w = unit(sub(origin, target))
    u = unit(cross(up, w))
    v_vec = cross(w, u)
    focus_dist = _num(cam, "focus_dist")
    if focus_dist == 0:
        focus_dist = length(sub(origin, target))
    horizontal = mul(u, viewport_width * focus_dist)
    vertical = mul(v_vec, viewport_height * focus_dist)
    c = Camera()
    c.origin = origin
    c.llc = sub(sub(sub(origin, div(horizontal, 2)), div(vertical, 2)), mul(w, focus_dist))
    c.horizontal, c.vertical, c.u, c.v, c.w = horizontal, vertical, u, v_vec, w
    c.lens_radius = _num(cam, "aperture") / 2
    return c


def get_ray(c, s, t, rng):                                                        # camera.go:60-74
    if c.lens_radius > 0 and rng is not None:
        rd = mul(random_in_unit_sphere(rng), c.lens_radius)
        offset = add(mul(c.u, rd[0]), mul(c.v, rd[1]))
        return add(c.origin, offset), sub(sub(add(add(c.llc, mul(c.horizontal, s)), mul(c.vertical, t)), c.origin), offset)
    return c.origin, sub(add(add(c.llc, mul(c.horizontal, s)), mul(c.vertical, t)), c.origin)


# ------------------------------------------------------------------ background (renderer.go:56-92)
def make_background(sc):
    sky = sc.get("sky")
    if sky is not None and sky.get("type") == "gradient":
        horizon, zenith = _col(sky.get("horizon")), _col(sky.get("zenith"))

        def bg(o, d):
            dir_len = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
            if dir_len == 0:
                return horizon
            t = (d[1] / dir_len + 1.0) * 0.5
            if t < 0:
                t = 0
            if t > 1:
                t = 1
            return (horizon[0] * (1 - t) + zenith[0] * t, horizon[1] * (1 - t) + zenith[1] * t, horizon[2] * (1 - t) + zenith[2] * t)
        return bg
    if sky is not None and sky.get("type") == "solid":
        color = _col(sky.get("color"))
    else:
        color = _col(sc.get("background"))
    return lambda o, d: color


# ------------------------------------------------------------------ integrator (renderer.go:286-404)
def new_stats():
    return dict(samples=0, segments=0, exit_scans=0, prim_tests=0, accepts=[0, 0, 0], scatters=0,
                end_sky=0, end_emissive=0, end_rr=0, end_depth=0, end_noscatter=0)


def ray_color(o, d, world, background, depth, rng, st, log=None):
    if depth <= 0:                                                                # :287-289
        st["end_depth"] += 1
        return (0.0, 0.0, 0.0)
    t_min = 0.001                                                                 # :292
    hit_anything = False
    closest = MAXF
    rec = Hit()
    st["segments"] += 1
    for i, ob in enumerate(world):                                                # :297-302
        st["prim_tests"] += 1
        if hit_object(ob, o, d, t_min, closest, rec):
            hit_anything = True
            closest = rec.t
            rec.index = i
            st["accepts"][ob[0]] += 1                                             # (every accepted hit() of the scan, not only the last)
    if log is not None:
        log.append((rec.index if hit_anything else -1, rec.t if hit_anything else 0.0, bool(rec.front_face) if hit_anything else False))
    if not hit_anything:                                                          # :304-306
        st["end_sky"] += 1
        return background(o, d)
    mat = rec.mat
    em = emitted(mat)                                                             # :308
    ok, attenuation, so, sd = scatter(mat, rng, o, d, rec)                        # :309
    if not ok:                                                                    # :310-312
        st["end_emissive" if mat.typ == EMISSIVE else "end_noscatter"] += 1
        return em
    st["scatters"] += 1
    if mat.typ == DIELECTRIC and rec.front_face:                                  # :316-371
        exit_t_min = 0.0001
        exit_rec = None
        hit_exit = False
        exit_t = MAXF
        st["exit_scans"] += 1
        for ob in world:                                                          # :329-349
            temp = Hit()
            st["prim_tests"] += 1
            if hit_object(ob, so, sd, exit_t_min, exit_t, temp):
                if temp.mat.typ == DIELECTRIC and not temp.front_face and temp.t < exit_t:
                    dx = temp.p[0] - rec.p[0]
                    dy = temp.p[1] - rec.p[1]
                    dz = temp.p[2] - rec.p[2]
                    dist_sq = dx * dx + dy * dy + dz * dz
                    if dist_sq > 1e-8 and dist_sq < 1000.0:
                        hit_exit = True
                        exit_t = temp.t
                        exit_rec = temp
        if hit_exit:                                                              # :352-369
            dx = exit_rec.p[0] - rec.p[0]
            dy = exit_rec.p[1] - rec.p[1]
            dz = exit_rec.p[2] - rec.p[2]
            distance = math.sqrt(dx * dx + dy * dy + dz * dz)
            ab = mat.absorption
            if ab[0] > 0 or ab[1] > 0 or ab[2] > 0:
                attenuation = (math.exp(-ab[0] * distance), math.exp(-ab[1] * distance), math.exp(-ab[2] * distance))
            so = exit_rec.p
    if depth <= 3:                                                                # Russian roulette, :374-393
        max_att = go_max(attenuation[0], go_max(attenuation[1], attenuation[2]))
        if max_att < 1e-6:
            st["end_rr"] += 1
            return em
        rr_prob = go_min(max_att, 0.95)
        if rng.float64() > rr_prob:
            st["end_rr"] += 1
            return em
        attenuation = (attenuation[0] / rr_prob, attenuation[1] / rr_prob, attenuation[2] / rr_prob)
    nxt = ray_color(so, sd, world, background, depth - 1, rng, st, log)           # :398
    return (em[0] + attenuation[0] * nxt[0], em[1] + attenuation[1] * nxt[1], em[2] + attenuation[2] * nxt[2])   # :399-403


# ------------------------------------------------------------------ the pixel loop (renderer.go:94-98, 171-221)
class Scene:
    def __init__(self, sc: dict):
        self.doc = sc
        self.world = scene_to_world(sc)
        self.background = make_background(sc)

    def camera22(self, width, height):
        c = new_camera(self.doc.get("camera"), width, height)
        return [*c.origin, *c.llc, *c.horizontal, *c.vertical, *c.u, *c.v, *c.w, c.lens_radius]

    def primary_hits(self, width, height, xi_u=0.5, xi_v=0.5):
        """World index (-1 = miss) and ray parameter of the closest hit of the lens-free camera ray (camera.go:70-73) through
        (x + xi_u, flipY + xi_v) of every pixel (renderer.go:174-183, 292-302)."""
        cam = new_camera(self.doc.get("camera"), width, height)
        inv_w, inv_h, hm1 = 1.0 / float(width - 1), 1.0 / float(height - 1), float(height - 1)
        ids, ts = [], []
        for y in range(height):
            flip_y = hm1 - float(y)
            for x in range(width):
                u = (float(x) + xi_u) * inv_w
                vv = (flip_y + xi_v) * inv_h
                o, d = get_ray(cam, u, vv, None)
                rec, closest, hit = Hit(), MAXF, -1
                for i, ob in enumerate(self.world):
                    if hit_object(ob, o, d, 0.001, closest, rec):
                        closest, hit = rec.t, i
                ids.append(hit)
                ts.append(closest if hit >= 0 else 0.0)
        return ids, ts

    def render_sum(self, width, height, spp, max_depth, seed=1, s_begin=0):
        """Un-normalised per-pixel radiance sums of samples [s_begin, s_begin + spp) (renderer.go:176-187) and the event
        counters; RNG keyed (seed, pixel = y * W + x, sample)."""
        cam = new_camera(self.doc.get("camera"), width, height)
        inv_w, inv_h, hm1 = 1.0 / float(width - 1), 1.0 / float(height - 1), float(height - 1)     # :95-98
        st = new_stats()
        out = []
        for y in range(height):
            flip_y = hm1 - float(y)                                                                # :174
            for x in range(width):
                col = (0.0, 0.0, 0.0)
                x_float = float(x)
                for s in range(s_begin, s_begin + spp):                                            # :181-187
                    rng = Rng(seed, y * width + x, s)
                    u = (x_float + rng.float64()) * inv_w
                    vv = (flip_y + rng.float64()) * inv_h
                    o, d = get_ray(cam, u, vv, rng)
                    st["samples"] += 1
                    col = add(col, ray_color(o, d, self.world, self.background, max_depth, rng, st))
                out.append(col)
        return out, st


def finalize_pixel(col, spp):                                                     # renderer.go:189-221
    inv_samples = 1.0 / float(spp)
    px = []
    for c in col:
        c = c * inv_samples
        c = math.sqrt(c) if c >= 0 else math.nan
        v = c * 255.999
        if v < 0:
            v = 0
        elif v > 255.999:
            v = 255.999
        px.append(int(v) if v == v else 0)                                        # uint8(NaN) is implementation-defined in Go: 0 on amd64
    return px + [255]
