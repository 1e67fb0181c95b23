"""Host-side mirror of the reference's scene classes (Python flavour, used by tests/bench).

Same names and argument meaning as raytracer/{Element,Obj,Bezier,Light,Scene,Camera}.h; the
C++ flavour of the same mirror lives in raytrace3_b200/include/.  Nothing here computes an
intersection: these classes only hold state and flatten it into the PODs of
include/sppm_b200.h (`Scene.flatten()`, `Camera.desc()`).
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import abi

M_EPS = 1e-4          # Vec3.h:6
PI = 3.1415926        # Vec3.h:8
DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _v(x):
    a = np.asarray(x, dtype=np.float64).reshape(3)
    return a.copy()


def _unit(v):
    """Vec3::GetUnit / Normalize (Vec3.h:48-56): untouched when |v| < 1e-4; multiplies by 1/|v|."""
    mo = math.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])
    if abs(mo) < M_EPS:
        return v.copy()
    inv = 1.0 / mo
    return np.array([v[0] * inv, v[1] * inv, v[2] * inv])


def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


class Material:
    """Element.h:7-19."""

    def __init__(self, d, s, r, rr=1.5, rl=1.0):
        self.diff, self.refl, self.refr, self.refrn, self.refln = _v(d), _v(s), _v(r), float(rr), float(rl)


class Texture:
    """Element.h:43-75, pre-decoded: a uint8 array [rows][cols][3] in BGR order.

    `Texture(path)` reads a JPEG/PNG with cv2 (the codec family the reference uses,
    Element.h:49); `Texture(array)` takes pixels directly."""

    def __init__(self, src):
        if isinstance(src, str):
            import cv2
            img = cv2.imread(src)
            if img is None:
                raise FileNotFoundError(f"texture {src!r} not readable (the reference would index out of range, Q18)")
            src = img
        self.map = np.ascontiguousarray(src, dtype=np.uint8)
        assert self.map.ndim == 3 and self.map.shape[2] == 3
        self.rows, self.cols = self.map.shape[:2]


class Obj:
    """Obj.h:5-54 (state only)."""
    kind = -1

    def __init__(self, mm: Material, cc):
        self.material, self.color, self.texture = mm, _v(cc), None

    def setTexture(self, texture):
        self.texture = texture


class PlaneObj(Obj):
    """Obj.h:55-101."""
    kind = abi.SPPM_OBJ_PLANE

    def __init__(self, pp, nn, cc, mm):
        super().__init__(mm, cc)
        self.P0, self.norm = _v(pp), _unit(_v(nn))
        self.texU, self.texV = _v((400, 0, 0)), _v((0, 0, 300))


class SphereObj(Obj):
    """Obj.h:102-154."""
    kind = abi.SPPM_OBJ_SPHERE

    def __init__(self, rradius, orgg, cc, mm):
        super().__init__(mm, cc)
        self.org, self.radius = _v(orgg), float(rradius)
        self.texU, self.texV = _unit(_v((0, 3, -3))), _unit(_v((1, 0, 0)))


def read_bpt(path):
    """The .bpt format read by the Bezier3Obj ctor (Bezier.h:204-218): `size`, then per patch
    `m n` and (m+1)(n+1) lines `x y z`; point (j,k) is stored at index j*(n+1)+k."""
    tok = open(path).read().split()
    it = iter(tok)
    size = int(next(it))
    out = np.zeros((size, 16, 3))
    for i in range(size):
        m, n = int(next(it)), int(next(it))
        assert (m + 1) * (n + 1) <= 16
        for j in range(m + 1):
            for k in range(n + 1):
                for l in range(3):
                    out[i, j * (n + 1) + k, l] = float(next(it))
    return out


def write_bpt(path, patches):
    """Write patches [n][16][3] in the .bpt format of read_bpt / Bezier.h:204-218."""
    with open(path, "w") as f:
        f.write(f"{len(patches)}\n")
        for P in patches:
            f.write("3 3\n")
            for q in range(16):
                f.write(" ".join(repr(float(x)) for x in P[q]) + "\n")


def write_bgr8(path, tex):
    """Write a pre-decoded texture for the C++ host mirror (include/raytracer/Element.h)."""
    m = tex.map if isinstance(tex, Texture) else np.ascontiguousarray(tex, dtype=np.uint8)
    with open(path, "wb") as f:
        f.write(f"BGR8 {m.shape[0]} {m.shape[1]}\n".encode())
        f.write(m.tobytes())


def teapot_patches():
    """The 32 bicubic patches of the Utah teapot (Newell's data set), raw file coordinates."""
    return np.load(os.path.join(DATA_DIR, "teapot32.npy"))


class Bezier3Obj(Obj):
    """Bezier.h:188-289: control points p -> Trans*(scale*p) + C (Bezier.h:207-222)."""
    kind = abi.SPPM_OBJ_BEZIER

    def __init__(self, Cc, scale, file, cc, mm, Trans=None):
        super().__init__(mm, cc)
        raw = read_bpt(file) if isinstance(file, str) else np.asarray(file, dtype=np.float64)
        T = np.eye(3) if Trans is None else np.asarray(Trans, dtype=np.float64)
        Cc = _v(Cc)
        P = np.zeros_like(raw)
        for i in range(raw.shape[0]):
            for q in range(16):
                p = raw[i, q] * scale
                for r in range(3):  # Matx * Vec: s = 0; s += m(r,k) * p(k)
                    s = 0.0
                    for k in range(3):
                        s += T[r, k] * p[k]
                    P[i, q, r] = s + Cc[r]
        self.P = P


class Light:
    """Light.h:5-17."""

    def __init__(self, p, c):
        self.pos, self.color = _v(p), _v(c)


class SpotLight(Light):
    """Light.h:19-26 (adds nothing)."""


def matmul33(a, b):
    """cv::Matx33d product with sequential accumulation."""
    out = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            s = 0.0
            for k in range(3):
                s += a[i, k] * b[k, j]
            out[i, j] = s
    return out


class Scene:
    """Scene.h:93-183: objvec / plvec containers; flatten() builds the SppmSceneDesc."""

    def __init__(self):
        self.objvec: list[Obj] = []
        self.plvec: list[Light] = []
        self._keep = None

    def flatten(self) -> abi.SppmSceneDesc:
        texs, tex_index = [], {}
        objs = (abi.SppmObject * max(1, len(self.objvec)))()
        patches = []
        for i, o in enumerate(self.objvec):
            d = objs[i]
            d.kind = o.kind
            d.tex_id = -1
            if o.texture is not None:
                if id(o.texture) not in tex_index:
                    tex_index[id(o.texture)] = len(texs)
                    texs.append(o.texture)
                d.tex_id = tex_index[id(o.texture)]
            m = o.material
            d.mat.diff[:], d.mat.refl[:], d.mat.refr[:], d.mat.refrn = m.diff, m.refl, m.refr, m.refrn
            d.color[:] = o.color
            if o.kind == abi.SPPM_OBJ_PLANE:
                d.p0[:], d.norm[:], d.texU[:], d.texV[:] = o.P0, o.norm, o.texU, o.texV
            elif o.kind == abi.SPPM_OBJ_SPHERE:
                d.p0[:], d.radius, d.texU[:], d.texV[:] = o.org, o.radius, o.texU, o.texV
            else:
                d.first_patch = sum(p.shape[0] for p in patches)
                d.n_patch = o.P.shape[0]
                patches.append(o.P)
        P = np.ascontiguousarray(np.concatenate(patches, axis=0) if patches else np.zeros((0, 16, 3)))
        lights = (abi.SppmLight * max(1, len(self.plvec)))()
        for i, l in enumerate(self.plvec):
            lights[i].pos[:], lights[i].color[:] = l.pos, l.color
        tarr = (abi.SppmTexture * max(1, len(texs)))()
        for i, t in enumerate(texs):
            tarr[i].rows, tarr[i].cols, tarr[i].bgr8 = t.rows, t.cols, t.map.ctypes.data
        desc = abi.SppmSceneDesc(
            n_objects=len(self.objvec), n_patches=P.shape[0], n_lights=len(self.plvec), n_textures=len(texs),
            objects=objs, patches=P.ctypes.data, lights=lights, textures=tarr)
        self._keep = (objs, P, lights, tarr, texs)  # keep the host arrays alive with the scene
        return desc

    def texture_bytes(self) -> int:
        seen = {id(o.texture): o.texture for o in self.objvec if o.texture is not None}
        return sum(t.map.nbytes for t in seen.values())


class Camera:
    """Camera.h:4-114: init / MoveTo / lookAt build the pinhole basis; `pic` is the film."""
    DEFAULT_HEIGHT = 1024
    DEFAULT_WIDTH = 1024

    def __init__(self):
        self.pos = np.zeros(3)
        self.look = np.zeros(3)
        self.w = self.h = 0
        self.dir = self.du = self.dv = np.zeros(3)
        self.pic = None

    def init(self, w_=DEFAULT_WIDTH, h_=DEFAULT_HEIGHT):
        self.h, self.w = int(h_), int(w_)

    def MoveTo(self, pos_):
        self.pos = _v(pos_)

    def lookAt(self, P, scale=1.0):  # Camera.h:32-54
        self.look = _v(P)
        up = np.array([0.0, 0.0, 1.0])
        self.w = int(self.w * scale)
        self.h = int(self.h * scale)
        d = _unit(self.look - self.pos)
        du = _unit(_cross(d, up))
        dv = _unit(-_cross(d, du))
        fov = 50. * PI / 180.
        self.dir = d * (.5 / math.tan(fov / 2.))
        self.du, self.dv = du, dv
        if self.pic is None:
            self.pic = np.zeros((self.h, self.w, 3))

    def copy(self):
        """Camera's copy constructor (Camera.h:55-74): same pose, a fresh zeroed film."""
        c = Camera()
        c.pos, c.look, c.w, c.h = self.pos.copy(), self.look.copy(), self.w, self.h
        c.dir, c.du, c.dv = self.dir.copy(), self.du.copy(), self.dv.copy()
        c.pic = np.zeros((self.h, self.w, 3))
        return c

    def clear(self):
        self.pic[:] = 0

    def desc(self) -> abi.SppmCameraDesc:
        d = abi.SppmCameraDesc(w=self.w, h=self.h)
        d.pos[:], d.dir[:], d.du[:], d.dv[:] = self.pos, self.dir, self.du, self.dv
        return d


# ---- the scenes BASELINE.json names ------------------------------------------------------
def synthetic_texture(rows, cols, seed):
    """Deterministic stand-in for the reference's JPEG textures (same sizes, BGR8): smooth
    low-frequency colour waves plus per-pixel noise, so bilinear taps differ everywhere."""
    rng = np.random.RandomState(seed)
    r = np.arange(rows, dtype=np.float32)[:, None]
    c = np.arange(cols, dtype=np.float32)[None, :]
    out = np.empty((rows, cols, 3), dtype=np.uint8)
    for k in range(3):
        f1, f2, ph = rng.uniform(2, 9), rng.uniform(2, 9), rng.uniform(0, 6.28)
        base = 150 + 60 * np.sin(f1 * r / rows * 6.28 + ph) * np.cos(f2 * c / cols * 6.28)
        out[:, :, k] = np.clip(base + rng.randint(-20, 21, size=(rows, cols)), 40, 250).astype(np.uint8)
    return out


def _materials():
    Z = (0, 0, 0)
    return dict(
        WHITE_DIFF=Material((.75, .75, .75), Z, Z, 0),
        MIRROR=Material(Z, (.999, .999, .999), Z, 0),
        MIRROR2=Material((.25, .25, .25), (.999, .999, .999), Z, 0),
        REFR0=Material(Z, Z, (.999, .999, .999), 1.5),
        REFR4=Material((.2, .2, .2), Z, (0, .5, .5), 1.5),
    )


def teapot_transform():
    """Scene.h:142-152: Trans2 * Trans with theta = (90/180)*PI (PI truncated)."""
    Trans = np.zeros((3, 3))
    Trans[0, 0] = 1
    Trans[1, 2] = 1
    Trans[2, 1] = 1
    Trans2 = np.zeros((3, 3))
    theta = (90. / 180.) * PI
    Trans2[0, 0] = math.cos(theta)
    Trans2[0, 2] = math.sin(theta)
    Trans2[1, 1] = 1
    Trans2[2, 0] = -math.sin(theta)
    Trans2[2, 2] = math.cos(theta)
    return matmul33(Trans2, Trans)


def cornell_objects():
    """Scene.h:117-125: five axis planes and three spheres (BASELINE config 2 / 4)."""
    m = _materials()
    W = m["WHITE_DIFF"]
    return [
        PlaneObj((1, 40.8, 81.6), (1, 0, 0), (.75, .25, .25), W),      # Left
        PlaneObj((99, 40.8, 81.6), (1, 0, 0), (.25, .25, .75), W),     # Right
        PlaneObj((50, 40.8, 0), (0, 0, 1), (.75, .75, .75), m["MIRROR"]),  # Back
        PlaneObj((50, 0, 81.6), (0, -1, 0), (.75, .75, .75), W),       # Bottom
        PlaneObj((50, 81.6, 81.6), (0, 1, 0), (.75, .75, .75), W),     # Top
        SphereObj(16.5, (27, 16.5, 47), (.999, .999, .999), m["MIRROR"]),
        SphereObj(16.5, (73, 16.5, 88), (.999, .999, .999), m["REFR0"]),
        SphereObj(8.5, (50, 8.5, 60), (.999, .999, .999), W),
    ]


def make_scene(name: str, textures="synthetic") -> Scene:
    """`cornell` = Scene.h:117-125 untextured; `reference` = Scene::init (Scene.h:98-157)
    with the teapot untextured (blue.jpg is absent upstream, Q18); `stress` = reference plus
    three more teapot instances (BASELINE config 5).  textures: "synthetic" | None | a dict
    {"wall","timg","planet"} of Texture objects."""
    s = Scene()
    s.objvec = cornell_objects()
    s.plvec = [SpotLight((50, 60, 85), (5000., 5000., 5000.))]  # Scene.h:157
    if name == "cornell":
        return s
    if name not in ("reference", "stress"):
        raise ValueError(name)
    if textures is not None:
        if textures == "synthetic":
            textures = dict(wall=Texture(synthetic_texture(2094, 2950, 1)),
                            timg=Texture(synthetic_texture(1675, 2360, 2)),
                            planet=Texture(synthetic_texture(399, 600, 3)))
        for i in (0, 1, 4):  # Scene.h:134-136
            s.objvec[i].setTexture(textures["wall"])
        s.objvec[7].setTexture(textures["planet"])  # Scene.h:139
        s.objvec[3].setTexture(textures["timg"])    # Scene.h:141
    W = _materials()["WHITE_DIFF"]
    T = teapot_transform()
    tp = teapot_patches()
    s.objvec.append(Bezier3Obj((20, 0, 120), 4, tp, (0, .999, .999), W, T))  # Scene.h:153
    if name == "stress":
        s.objvec.append(Bezier3Obj((75, 0, 130), 3, tp, (.999, .999, 0), W, T))
        s.objvec.append(Bezier3Obj((50, 0, 150), 2.5, tp, (.999, 0, .999), W, np.array([[1., 0, 0], [0, 0, 1], [0, 1, 0]])))
        s.objvec.append(Bezier3Obj((45, 30, 20), 3.5, tp, (.6, .999, .6), W, T))
    return s


def reference_camera(w, h) -> Camera:
    """main.cpp:22-27."""
    cam = Camera()
    org = np.array([50., 35., 230.])
    cam.init(w, h)
    cam.MoveTo(org)
    cam.lookAt(org + np.array([0, 0.042612, -1.]), 1 / 1.)
    return cam
