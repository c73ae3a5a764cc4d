"""``Renderer`` -- same surface as /root/reference/meshreg/neurender/renderer.py:12-295.

Constructor arguments, ``forward(vertices, faces, textures, mode, K, R, t, dist_coeffs, orig_size,
detach_renders)``, the ``render`` / ``render_rgb`` / ``render_silhouettes`` / ``render_depth`` /
``project`` methods and the returned dict are the reference's.  The camera / lighting helpers come
from ``nrfuncs`` (restating ``neural_renderer``), rasterization from ``rasterize`` (libhoc_b200.so).
"""
from __future__ import division

import math

import numpy
import torch
import torch.nn as nn

from . import nrfuncs as nr
from . import rasterize


def _cuda_float(a):
    return torch.as_tensor(a, dtype=torch.float32).cuda()


class Renderer(nn.Module):
    def __init__(self, image_size=256, anti_aliasing=True, background_color=[0, 0, 0], fill_back=True,
                 camera_mode="projection", K=None, R=None, t=None, dist_coeffs=None, orig_size=1024,
                 perspective=True, viewing_angle=30, camera_direction=[0, 0, 1], near=0.1, far=100,
                 light_intensity_ambient=0.5, light_intensity_directional=0.5, light_color_ambient=[1, 1, 1],
                 light_color_directional=[1, 1, 1], light_direction=[0, 1, 0], no_light=False,
                 return_face_inv_map=True, return_weight_map=True):
        """
        Wrapper on top of the rasterizer (renderer.py:13-83).  ``return_face_inv_map`` /
        ``return_weight_map`` are extensions (default: reference behaviour): a caller that never reads
        those maps can turn the 36 + 12 B/px writes off.
        """
        super(Renderer, self).__init__()
        # rendering
        self.image_size = image_size
        self.anti_aliasing = anti_aliasing
        self.background_color = background_color
        self.fill_back = fill_back
        self.no_light = no_light
        self.return_face_inv_map = return_face_inv_map
        self.return_weight_map = return_weight_map

        # camera
        self.camera_mode = camera_mode
        if self.camera_mode == "projection":
            self.K = K
            self.R = R
            self.t = t
            if isinstance(self.K, numpy.ndarray):
                self.K = _cuda_float(self.K)
            if isinstance(self.R, numpy.ndarray):
                self.R = _cuda_float(self.R)
            if isinstance(self.t, numpy.ndarray):
                self.t = _cuda_float(self.t)
            self.dist_coeffs = dist_coeffs
            self.orig_size = orig_size
        elif self.camera_mode in ["look", "look_at"]:
            self.perspective = perspective
            self.viewing_angle = viewing_angle
            self.eye = [0, 0, -(1.0 / math.tan(math.radians(self.viewing_angle)) + 1)]
            self.camera_direction = [0, 0, 1]
        else:
            raise ValueError("Camera mode has to be one of projection, look or look_at")

        self.near = near
        self.far = far

        # light
        self.light_intensity_ambient = light_intensity_ambient
        self.light_intensity_directional = light_intensity_directional
        self.light_color_ambient = light_color_ambient
        self.light_color_directional = light_color_directional
        self.light_direction = light_direction

        # rasterization
        self.rasterizer_eps = 1e-3

    def forward(self, vertices, faces, textures=None, mode=None, K=None, R=None, t=None, dist_coeffs=None,
                orig_size=None, detach_renders=False):
        """Dispatch on ``mode`` like renderer.py:85-112."""
        if mode is None:
            return self.render(vertices, faces, textures, K, R, t, dist_coeffs, orig_size,
                               detach_renders=detach_renders)
        elif mode == "rgb":
            return self.render_rgb(vertices, faces, textures, K, R, t, dist_coeffs, orig_size)
        elif mode == "silhouettes":
            return self.render_silhouettes(vertices, faces, K, R, t, dist_coeffs, orig_size)
        elif mode == "depth":
            return self.render_depth(vertices, faces, K, R, t, dist_coeffs, orig_size)
        else:
            raise ValueError("mode should be one of None, 'silhouettes' or 'depth'")

    # -- helpers ------------------------------------------------------------------------------
    def _fill_back(self, faces, textures=None, detach=True):
        faces = torch.cat((faces, faces.flip(-1)), dim=1)
        if detach:
            faces = faces.detach()
        if textures is not None:
            textures = torch.cat((textures, textures.permute((0, 1, 4, 3, 2, 5))), dim=1)
        return faces, textures

    def _light(self, vertices, faces, textures):
        faces_lighting = nr.vertices_to_faces(vertices, faces)
        return nr.lighting(faces_lighting, textures, self.light_intensity_ambient, self.light_intensity_directional,
                           self.light_color_ambient, self.light_color_directional, self.light_direction)

    def project(self, vertices, K=None, R=None, t=None, dist_coeffs=None, orig_size=None):
        """Viewpoint transformation (renderer.py:165-189)."""
        if self.camera_mode == "look_at":
            vertices = nr.look_at(vertices, self.eye)
            if self.perspective:
                vertices = nr.perspective(vertices, angle=self.viewing_angle)
        elif self.camera_mode == "look":
            vertices = nr.look(vertices, self.eye, self.camera_direction)
            if self.perspective:
                vertices = nr.perspective(vertices, angle=self.viewing_angle)
        elif self.camera_mode == "projection":
            if K is None:
                K = self.K
            if R is None:
                R = self.R
            if t is None:
                t = self.t
            if dist_coeffs is None:
                dist_coeffs = self.dist_coeffs
            if orig_size is None:
                orig_size = self.orig_size
            dev, dt = vertices.device, vertices.dtype
            if R is None:  # the reference's callers always give K only; identity extrinsics
                R = torch.eye(3, dtype=dt, device=dev)[None]
            if t is None:
                t = torch.zeros(1, 1, 3, dtype=dt, device=dev)
            if dist_coeffs is None:  # renderer.py:60-62
                dist_coeffs = torch.zeros(1, 5, dtype=dt, device=dev)
            vertices = nr.projection(vertices, K, R, t, dist_coeffs, orig_size)
        return vertices

    # -- render modes -------------------------------------------------------------------------
    def render_silhouettes(self, vertices, faces, K=None, R=None, t=None, dist_coeffs=None, orig_size=None):
        if self.fill_back:
            faces, _ = self._fill_back(faces, detach=False)
        vertices = self.project(vertices, K=K, R=R, t=t, dist_coeffs=dist_coeffs, orig_size=orig_size)
        faces = nr.vertices_to_faces(vertices, faces)
        return rasterize.rasterize_silhouettes(faces, self.image_size, self.anti_aliasing)

    def render_depth(self, vertices, faces, K=None, R=None, t=None, dist_coeffs=None, orig_size=None):
        if self.fill_back:
            faces, _ = self._fill_back(faces)
        vertices = self.project(vertices, K=K, R=R, t=t, dist_coeffs=dist_coeffs, orig_size=orig_size)
        faces = nr.vertices_to_faces(vertices, faces)
        return rasterize.rasterize_depth(faces, self.image_size, self.anti_aliasing)

    def render_rgb(self, vertices, faces, textures, K=None, R=None, t=None, dist_coeffs=None, orig_size=None):
        if self.fill_back:
            faces, textures = self._fill_back(faces, textures)
        if not self.no_light:
            textures = self._light(vertices, faces, textures)
        vertices = self.project(vertices, K=K, R=R, t=t, dist_coeffs=dist_coeffs, orig_size=orig_size)
        faces = nr.vertices_to_faces(vertices, faces)
        return rasterize.rasterize(faces, textures, self.image_size, self.anti_aliasing, self.near, self.far,
                                   self.rasterizer_eps, self.background_color)

    def render(self, vertices, faces, textures, K=None, R=None, t=None, dist_coeffs=None, orig_size=None,
               detach_renders=False):
        """rgb + alpha + depth + index maps (renderer.py:237-295)."""
        if self.fill_back:
            faces, textures = self._fill_back(faces, textures)
        if not self.no_light:
            textures = self._light(vertices, faces, textures)
        vertices = self.project(vertices, K=K, R=R, t=t, dist_coeffs=dist_coeffs, orig_size=orig_size)
        faces = nr.vertices_to_faces(vertices, faces)
        if detach_renders:
            faces = faces.detach()
        return rasterize.rasterize_rgbad(faces, textures, self.image_size, self.anti_aliasing, self.near, self.far,
                                         self.rasterizer_eps, self.background_color,
                                         return_face_inv_map=self.return_face_inv_map,
                                         return_weight_map=self.return_weight_map)
