"""Scene description files and render jobs — the Python twin of
include/raytracer/SceneFile.h; the format is documented there.  Replaces the reference's
hard-coded Scene::init / main() (Scene.h:98-164, main.cpp:19-41)."""
from __future__ import annotations

import dataclasses
import os

import numpy as np

from . import abi
from .scene import Bezier3Obj, Camera, Material, PlaneObj, Scene, SphereObj, SpotLight, Texture, teapot_transform


@dataclasses.dataclass
class RenderJob:
    w: int = 320
    h: int = 240
    cam_pos: tuple = (50., 35., 230.)
    cam_look: tuple = (50., 35.042612, 229.)
    cam_scale: float = 1.0
    passes: int = 1
    replicas: int = 4
    rounds: int = 10
    photons: int = 100000
    seed: int = 0x5EED5EED
    shared: bool = False
    newton_mode: int = abi.SPPM_NEWTON_EXACT
    gather_mode: int = abi.SPPM_GATHER_FILTERED
    math_mode: int = abi.SPPM_MATH_EXACT
    jitter: float = 0.00015
    out_prefix: str = "aaanew"
    format: str = "ppm"
    quality: int = 95
    checkpoint: str = ""
    checkpoint_every: int = 1

    def camera(self) -> Camera:
        c = Camera()
        c.init(self.w, self.h)
        c.MoveTo(np.array(self.cam_pos))
        c.lookAt(np.array(self.cam_look), self.cam_scale)
        return c

    def params(self) -> abi.SppmParams:
        p = abi.default_params()
        p.seed, p.newton_mode, p.gather_mode, p.math_mode = self.seed, self.newton_mode, self.gather_mode, self.math_mode
        return p


class SceneFileError(ValueError):
    pass


def load_scene_file(path: str, texture_loader=None) -> tuple[Scene, RenderJob]:
    """Parse `path`.  texture_loader(file_path) -> Texture decodes textures (default: Texture(path),
    i.e. OpenCV; pass ctx-based nvJPEG decoding for the device path); a texture that cannot be read
    leaves its objects untextured, as Scene::init does for the absent blue.jpg (Q18)."""
    scene, job = Scene(), RenderJob()
    base = os.path.dirname(os.path.abspath(path))
    materials: dict[str, Material] = {}
    textures: dict[str, Texture | None] = {}
    load_tex = texture_loader or (lambda f: Texture(f))
    with open(path) as fh:
        for ln, raw in enumerate(fh, 1):
            tok = raw.split("#", 1)[0].split()
            if not tok:
                continue
            kw, it = tok[0], iter(tok[1:])

            def bad(why):
                raise SceneFileError(f"{path}:{ln}: {why}")

            def word():
                try:
                    return next(it)
                except StopIteration:
                    bad(f"word expected in '{kw}'")

            def num():
                try:
                    return float(word())
                except ValueError:
                    bad(f"number expected in '{kw}'")

            def vec():
                return (num(), num(), num())

            def material():
                n = word()
                if n not in materials:
                    bad(f"unknown material {n}")
                return materials[n]

            def options(allow_rot=False):
                tex, rot = None, None
                for opt in it:
                    if opt == "texture":
                        n = word()
                        if n not in textures:
                            bad(f"unknown texture {n}")
                        tex = textures[n]
                    elif opt == "rot" and allow_rot:
                        first = word()
                        if first == "teapot":
                            rot = teapot_transform()
                        else:
                            rot = np.array([float(first)] + [num() for _ in range(8)]).reshape(3, 3)
                    else:
                        bad(f"unknown option {opt}")
                return tex, rot

            if kw == "film":
                job.w, job.h = int(num()), int(num())
            elif kw == "camera":
                job.cam_pos, job.cam_look = vec(), vec()
                rest = list(it)
                if rest:
                    job.cam_scale = float(rest[0])
            elif kw == "material":
                n = word()
                f = dict(diff=(0., 0., 0.), refl=(0., 0., 0.), refr=(0., 0., 0.), refrn=1.5)
                for k in it:
                    if k in ("diff", "refl", "refr"):
                        f[k] = vec()
                    elif k == "refrn":
                        f[k] = num()
                    else:
                        bad(f"unknown material field {k}")
                materials[n] = Material(f["diff"], f["refl"], f["refr"], f["refrn"])
            elif kw == "texture":
                n, f = word(), word()
                try:
                    textures[n] = load_tex(os.path.join(base, f))
                except (OSError, FileNotFoundError) as e:
                    print(f"warning: {e} — objects using '{n}' stay untextured")
                    textures[n] = None
            elif kw == "plane":
                p0, nn, cc = vec(), vec(), vec()
                o = PlaneObj(p0, nn, cc, material())
                tex, _ = options()
                if tex is not None:
                    o.setTexture(tex)
                scene.objvec.append(o)
            elif kw == "sphere":
                r = num()
                c, cc = vec(), vec()
                o = SphereObj(r, c, cc, material())
                tex, _ = options()
                if tex is not None:
                    o.setTexture(tex)
                scene.objvec.append(o)
            elif kw == "bezier":
                f = word()
                c = vec()
                scale = num()
                cc = vec()
                m = material()
                tex, rot = options(allow_rot=True)
                o = Bezier3Obj(c, scale, os.path.join(base, f), cc, m, rot)
                if tex is not None:
                    o.setTexture(tex)
                scene.objvec.append(o)
            elif kw == "light":
                p, c = vec(), vec()
                scene.plvec.append(SpotLight(p, c))
            elif kw == "render":
                for k in it:
                    if k in ("passes", "replicas", "rounds", "photons"):
                        setattr(job, k, int(num()))
                    elif k == "seed":
                        job.seed = int(word(), 0)
                    elif k == "shared":
                        job.shared = num() != 0
                    elif k == "jitter":
                        job.jitter = num()
                    elif k == "newton":
                        job.newton_mode = abi.SPPM_NEWTON_FAST if word() == "fast" else abi.SPPM_NEWTON_EXACT
                    elif k == "gather":
                        job.gather_mode = abi.SPPM_GATHER_EXACT if word() == "exact" else abi.SPPM_GATHER_FILTERED
                    elif k == "math":
                        job.math_mode = abi.SPPM_MATH_FAST if word() == "fast" else abi.SPPM_MATH_EXACT
                    else:
                        bad(f"unknown render field {k}")
            elif kw == "output":
                for k in it:
                    if k == "prefix":
                        job.out_prefix = word()
                    elif k == "format":
                        job.format = word()
                    elif k == "quality":
                        job.quality = int(num())
                    elif k == "checkpoint":
                        job.checkpoint = word()
                    elif k == "every":
                        job.checkpoint_every = int(num())
                    else:
                        bad(f"unknown output field {k}")
            else:
                bad(f"unknown keyword {kw}")
    if not scene.objvec:
        raise SceneFileError(f"{path}: no objects")
    if not scene.plvec:
        raise SceneFileError(f"{path}: no lights")
    return scene, job


def render_job(ctx, scene: Scene, job: RenderJob):
    """Run a job on `ctx` through the host mirror (SPPMRayTracer); returns (film, bgr8 image)."""
    from .api import SPPMRayTracer
    ctx.set_params(job.params())
    sp = SPPMRayTracer(ctx)
    sp.sppmtime, sp.replicas, sp.cnt, sp.onetime = job.passes, job.replicas, job.rounds, job.photons
    sp.jitter, sp.shared_statistics = job.jitter, job.shared
    cam = job.camera()
    img = sp.render(scene, cam)
    return cam.pic, img
