"""Small batched quaternion / vector ops for the env observation code (quaternions are (x, y, z, w)).
Role of the reference's ``utils/torch_utils.py:31-151``; plain PyTorch, written independently."""
import torch


def to_torch(x, dtype=torch.float, device="cuda:0", requires_grad=False):
    return torch.tensor(x, dtype=dtype, device=device, requires_grad=requires_grad)


def quat_mul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack((aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw,
                        aw * bw - ax * bx - ay * by - az * bz), dim=-1)


def quat_conjugate(q):
    return torch.cat((-q[..., :3], q[..., 3:]), dim=-1)


def quat_rotate(q, v):
    w = q[..., 3:]
    u = q[..., :3]
    return v * (2.0 * w * w - 1.0) + torch.cross(u, v, dim=-1) * w * 2.0 + u * (u * v).sum(-1, keepdim=True) * 2.0


def normalize(x, eps=1e-9):
    return x / x.norm(p=2, dim=-1).clamp(min=eps).unsqueeze(-1)


def quat_from_angle_axis(angle, axis):
    half = (angle / 2).unsqueeze(-1)
    return normalize(torch.cat((normalize(axis) * half.sin(), half.cos()), dim=-1))


def normalize_angle(x):
    return torch.atan2(torch.sin(x), torch.cos(x))
