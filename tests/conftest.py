import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context for the session; fails loudly if the library or device is missing."""
    from raytrace3_b200 import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture()
def params():
    from raytrace3_b200 import abi
    return abi.default_params()


def rays_at_teapot(n, seed, scene):
    """Rays from a shell around the first Bezier object aimed into its bounding box."""
    rng = np.random.RandomState(seed)
    bz = [o for o in scene.objvec if o.kind == 2][0]
    lo, hi = bz.P.reshape(-1, 3).min(0), bz.P.reshape(-1, 3).max(0)
    ctr = 0.5 * (lo + hi)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    org = ctr + d * rng.uniform(30, 70, size=(n, 1))
    tgt = lo + (hi - lo) * rng.uniform(size=(n, 3))
    dirv = tgt - org
    dirv /= np.linalg.norm(dirv, axis=1, keepdims=True)
    return np.ascontiguousarray(np.concatenate([org, dirv], axis=1))


def rays_in_room(n, seed):
    rng = np.random.RandomState(seed)
    org = np.stack([rng.uniform(2, 98, n), rng.uniform(1, 80, n), rng.uniform(1, 220, n)], axis=1)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.ascontiguousarray(np.concatenate([org, d], axis=1))


def rays_at_patches(patch_idx, seed, scene):
    """One ray per entry of patch_idx, aimed at a random point ON that patch (so most solve)."""
    rng = np.random.RandomState(seed)
    P = np.concatenate([o.P for o in scene.objvec if o.kind == 2], axis=0)
    n = len(patch_idx)
    u, v = rng.uniform(0.05, 0.95, n), rng.uniform(0.05, 0.95, n)
    B = lambda t: np.stack([(1 - t) ** 3, 3 * t * (1 - t) ** 2, 3 * t * t * (1 - t), t ** 3], axis=1)
    G = P[patch_idx].reshape(n, 4, 4, 3)
    tgt = np.einsum("ni,nk,nikc->nc", B(v), B(u), G)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    org = tgt + d * rng.uniform(20, 60, size=(n, 1))
    return np.ascontiguousarray(np.concatenate([org, -d], axis=1))


def custom_scene():
    """Materials the shipped scene never combines: diffuse+mirror+glass on one object (eye
    paths branch, Raytracer.h:312-335) and a textured sphere (Obj.h:140-153)."""
    from raytrace3_b200.scene import Material, SpotLight, Texture, make_scene, synthetic_texture
    s = make_scene("cornell", textures=None)
    s.objvec[5].material = Material((.3, .3, .3), (.4, .4, .4), (.3, .3, .3), 1.5)
    s.objvec[7].setTexture(Texture(synthetic_texture(40, 60, 9)))
    s.objvec[0].setTexture(Texture(synthetic_texture(50, 30, 8)))
    s.plvec.append(SpotLight((10, 40, 140), (0., 2000., 2000.)))  # Scene.h:158 (commented-out second light)
    return s


def quirks_scene():
    """Branches of the reference the shipped scene never reaches: a TEXTURED teapot (Bezier3Obj looks
    its colour up as GetColor(u, t) -- the ray parameter as second coordinate, Bezier.h:278, Q6) and
    materials with a zero diffuse channel (IsZero(Vec3) is true if ANY channel is < 1e-4, so such an
    object is not IsDiff() and gets no hit points / deposits, yet Obj::Roulette, which uses channel
    means, still sends photons down the DIFF branch, Vec3.h:72-79, Obj.h:12-15,30-45, Q8)."""
    from raytrace3_b200.scene import Material, Texture, make_scene, synthetic_texture
    s = make_scene("reference", textures=None)
    s.objvec[8].setTexture(Texture(synthetic_texture(37, 53, 11)))
    s.objvec[5].material = Material((.5, 0., .5), (.3, .3, .3), (0., 0., 0.), 1.5)
    s.objvec[2].material = Material((0., .7, .7), (0., 0., 0.), (0., 0., 0.), 1.5)
    return s


def branching_scene():
    """Eye paths that branch at most hits: diffuse + mirror + glass on the left and back walls and on two
    spheres (Raytracer.h:312-335 takes all three branches) -> about 8 hit points per pixel, more than
    the 4 per pixel the eye pass reserves up front (sppm_eye_pass_rows grows the buffer and traces again)."""
    from raytrace3_b200.scene import Material, make_scene
    s = make_scene("cornell", textures=None)
    for i in (0, 2, 5, 6):
        s.objvec[i].material = Material((.3, .3, .3), (.4, .4, .4), (.3, .3, .3), 1.5)
    return s
