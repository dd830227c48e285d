"""Learned camera distribution (reference src/training/networks_camera_adaptor.py:22-134): two small softplus MLPs map a camera sampled from
the prior -- (yaw, pitch, roll, radius) and (fov, look-at yaw / pitch / radius) -- to the posterior camera used for rendering, conditioned on
the class label (origin) and on (z, class) (look-at).  Inputs are normalised to the prior's ranges, outputs squashed back with a sigmoid.
Module / parameter names equal the reference's, so its snapshots load unchanged.  The gradient reaches these weights through
d(ray_o), d(ray_d) of the fused ray-march backward (csrc/raymarch_bwd*.cu)."""
import torch

from ..dnnlib import TensorGroup
from .layers import FullyConnectedLayer, normalize_2nd_moment
from .rendering_utils import sample_camera_params


class ParamsAdaptor(torch.nn.Module):
    """params [B, in] (+ z, + c embeddings) -> [B, out]   (:22-51)."""

    def __init__(self, cfg, in_channels, out_channels, use_z=True):
        super().__init__()
        self.cfg = cfg
        fc = lambda i, o, act: FullyConnectedLayer(i, o, activation=act, lr_multiplier=cfg.lr_multiplier)
        self.project_params = fc(in_channels, cfg.hid_dim, 'softplus')
        self.project_z = fc(cfg.z_dim, cfg.embed_dim, 'softplus') if use_z else None
        self.project_c = fc(cfg.c_dim, cfg.embed_dim, 'softplus') if cfg.c_dim > 0 else None
        width = cfg.hid_dim + (cfg.embed_dim if use_z else 0) + (cfg.embed_dim if cfg.c_dim > 0 else 0)
        self.main = torch.nn.Sequential(fc(width, cfg.hid_dim, 'softplus'), fc(cfg.hid_dim, out_channels, 'linear'))

    def forward(self, x, z=None, c=None):
        feats = [self.project_params(x)]
        if self.project_z is not None:
            feats.append(normalize_2nd_moment(self.project_z(z)))
        if self.project_c is not None:
            feats.append(normalize_2nd_moment(self.project_c(c)))
        return self.main(torch.cat(feats, dim=1))


_ORDER = ('yaw', 'pitch', 'roll', 'fov', 'radius', 'la_yaw', 'la_pitch', 'la_radius')      # column order of the unrolled [B, 8] form (:70-71)


class CameraAdaptor(torch.nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.num_origin_cam_params = 4       # yaw, pitch, roll, radius
        self.num_look_at_cam_params = 4      # fov, look-at yaw, pitch, radius
        self.num_cam_params = 8
        self.origin_adaptor = ParamsAdaptor(cfg, 4, 4, use_z=False)
        self.look_at_adaptor = ParamsAdaptor(cfg, 8, 4)

    def sample_from_prior(self, *args, **kwargs):
        return sample_camera_params(self.cfg.camera, *args, **kwargs)

    @staticmethod
    def unroll_camera_params(cp):
        return torch.cat([cp.angles, cp.fov.unsqueeze(1), cp.radius.unsqueeze(1), cp.look_at], dim=1)

    @staticmethod
    def roll_camera_params(t):
        return TensorGroup(angles=t[:, [0, 1, 2]], fov=t[:, 3], radius=t[:, 4], look_at=t[:, [5, 6, 7]])

    @staticmethod
    def _ranges(cam):
        """(min, max) of every squashed coordinate, as the reference reads them from the camera config (:77-82, :90-95)."""
        return dict(yaw=(cam.origin.angles.yaw.min, cam.origin.angles.yaw.max), pitch=(cam.origin.angles.pitch.min, cam.origin.angles.pitch.max),
                    fov=(cam.fov.min, cam.fov.max), la_yaw=(cam.look_at.angles.yaw.min, cam.look_at.angles.yaw.max),
                    la_pitch=(cam.look_at.angles.pitch.min, cam.look_at.angles.pitch.max), la_radius=(cam.look_at.radius.min, cam.look_at.radius.max))

    @staticmethod
    def normalize_camera_params(cam, cp, eps=1e-8):
        """prior camera -> [0, 1] per coordinate (roll and radius pass through)  (:74-84)."""
        cols = dict(zip(_ORDER, CameraAdaptor.unroll_camera_params(cp).split(1, dim=1)))
        for k, (lo, hi) in CameraAdaptor._ranges(cam).items():
            cols[k] = (cols[k] - lo) / (hi - lo + eps)
        return CameraAdaptor.roll_camera_params(torch.cat([cols[k] for k in _ORDER], dim=1))

    @staticmethod
    def denormalize_camera_params(cam, cp):
        """network output -> camera: sigmoid into the prior's range; pitch keeps 1e-5 off both poles, roll is forced to 0, the radius is
        not squashed; the look-at radius uses the reference's own expression (:95: its range mixes in the look-at pitch minimum)."""
        cols = dict(zip(_ORDER, CameraAdaptor.unroll_camera_params(cp).split(1, dim=1)))
        r = CameraAdaptor._ranges(cam)
        cols['yaw'] = cols['yaw'].sigmoid() * (r['yaw'][1] - r['yaw'][0]) + r['yaw'][0]
        cols['pitch'] = cols['pitch'].sigmoid() * (r['pitch'][1] - r['pitch'][0] - 2e-5) + r['pitch'][0] + 1e-5
        cols['roll'] = cols['roll'] * 0.0
        cols['fov'] = cols['fov'].sigmoid() * (r['fov'][1] - r['fov'][0]) + r['fov'][0]
        cols['la_yaw'] = cols['la_yaw'].sigmoid() * (r['la_yaw'][1] - r['la_yaw'][0]) + r['la_yaw'][0]
        cols['la_pitch'] = cols['la_pitch'].sigmoid() * (r['la_pitch'][1] - r['la_pitch'][0]) + r['la_pitch'][0]
        cols['la_radius'] = cols['la_radius'].sigmoid() * (r['la_radius'][1] - r['la_pitch'][0]) + r['la_pitch'][0]
        return CameraAdaptor.roll_camera_params(torch.cat([cols[k] for k in _ORDER], dim=1))

    def adjust_for_prior(self, old, new):
        """Coordinates the config does not let the adaptor move keep their prior value (:99-109)."""
        adj = self.cfg.adjust
        return TensorGroup(angles=new.angles if adj.angles else old.angles + 0.0 * new.angles,
                           fov=new.fov if adj.fov else old.fov + 0.0 * new.fov,
                           radius=new.radius if adj.radius else old.radius + 0.0 * new.radius,
                           look_at=new.look_at if adj.look_at else old.look_at + 0.0 * new.look_at)

    def compute_new_camera_params(self, old_norm, z, c):
        """origin first (class-conditional), then the look-at head sees the NEW origin next to the old fov / look-at  (:111-124)."""
        origin = self.origin_adaptor(torch.cat([old_norm.angles, old_norm.radius.unsqueeze(1)], dim=1), c=c)                      # [B, 4]
        la_in = torch.cat([origin[:, :3], old_norm.fov.unsqueeze(1), origin[:, [3]], old_norm.look_at], dim=1)                   # [B, 8]
        la = self.look_at_adaptor(la_in, z, c)                                                                                    # [B, 4]
        new = torch.cat([origin[:, :3], la[:, [0]], origin[:, [3]], la[:, [1, 2, 3]]], dim=1)
        if self.cfg.get('residual', False):
            new = new + self.unroll_camera_params(old_norm)
        return self.roll_camera_params(new)

    def forward(self, camera_params_old, z, c=None):
        old_norm = self.normalize_camera_params(self.cfg.camera, camera_params_old)
        new_norm = self.compute_new_camera_params(old_norm, z, c)
        new = self.denormalize_camera_params(self.cfg.camera, new_norm)
        return self.adjust_for_prior(camera_params_old, new)
