"""Shared helpers for the parity tests: build the SAME synthetic scene on the GPU path (C-ABI) and
on the compiled reference (oracle/_ref) from byte-identical arrays."""
import numpy as np

from pysicalbasedraytracer_b200 import _abi as A

EYE, LOOK, UP = (0.0, 0.0, 5.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0)


class Pair:
    def __init__(self, api, R, data, W, H, spp, eye=EYE, look=LOOK, up=UP, cam_type=A.CAMERA_PERSPECTIVE, ctx=None,
                 sampler_type=A.SAMPLER_HALTON, lens_radius=0.0, focal_distance=0.0, camera_medium=None,
                 honour_camera_medium=False):
        self.W, self.H, self.spp = W, H, spp
        c2w = api.Inverse(api.LookAt(eye, look, up))
        self.cam = (api.CreatePerspectiveCamera if cam_type == A.CAMERA_PERSPECTIVE else api.CreateOrthographicCamera)(W, H, c2w, camera_medium)
        self.cam.lens_radius, self.cam.focal_distance = lens_radius, focal_distance
        if honour_camera_medium:  # library extension, checked against the harness loop that re-attaches the medium
            self.cam.medium = camera_medium
        self.sampler_type = sampler_type
        if sampler_type == A.SAMPLER_SOBOL:
            # the Sobol tables are the reference's data: fetched from the compiled reference
            self.smp = api.SobolSampler(spp, (0, 0, W, H), *R.sobol_tables((0, 0, W, H)))
        else:
            self.smp = api.HaltonSampler(spp, (0, 0, W, H))
        self.scene = api.Scene(data, ctx)
        self.scene.set_camera(self.cam)
        self.scene.set_sampler(self.smp)
        self.R = R
        if R is not None:
            self.rscene = R.RefScene(data)
            self.rcam = R.RefCamera(cam_type, W, H, eye, look, up, lens_radius, focal_distance)
            if camera_medium is not None:
                self.rcam.set_medium(self.rscene, camera_medium, honour=honour_camera_medium)
            self.rsmp = R.RefSampler(spp, (0, 0, W, H), sampler_type)


def rel_err(a, b, floor=1e-3):
    """per-pixel relative error of RGB images: max_c |a-b| / max(max_c |b|, floor)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    num = np.abs(a - b).max(axis=-1)
    den = np.maximum(np.abs(b).max(axis=-1), floor)
    return num / den


def rmse(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)))


def golden(name):
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))
