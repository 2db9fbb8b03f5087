"""Stand-in for the ``urdfpy`` package (not in this image).

Only the slice of its object model that the reference's ``utils/load_utils.py:53-231``
touches is provided: ``URDF.load`` -> ``.links[].name/.collisions[].origin/.geometry``,
``.joints[].joint_type/.parent/.child/.axis/.origin/.limit/.dynamics``, ``.link_map`` and
``matrix_to_xyz_rpy``.  Used by the oracle only to let the reference CartPole env parse
``cartpole.urdf`` (test infrastructure; never imported by the product)."""
import math
import xml.etree.ElementTree as ET

import numpy as np


class _Record:
    def __init__(self, **fields):
        self.__dict__.update(fields)


def _floats(text, default):
    if text is None:
        return np.array(default, dtype=np.float64)
    return np.array([float(tok) for tok in text.split()], dtype=np.float64)


def _pose_matrix(node):
    """4x4 homogeneous matrix of an <origin xyz= rpy=> child (identity when absent)."""
    origin = None if node is None else node.find("origin")
    xyz = _floats(None if origin is None else origin.get("xyz"), (0.0, 0.0, 0.0))
    roll, pitch, yaw = _floats(None if origin is None else origin.get("rpy"), (0.0, 0.0, 0.0))
    cr, sr = math.cos(roll), math.sin(roll)
    cp, sp = math.cos(pitch), math.sin(pitch)
    cy, sy = math.cos(yaw), math.sin(yaw)
    out = np.eye(4)
    out[:3, :3] = [
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr],
    ]
    out[:3, 3] = xyz
    return out


def matrix_to_xyz_rpy(matrix):
    rot = matrix[:3, :3]
    pitch = -math.asin(max(-1.0, min(1.0, rot[2, 0])))
    roll = math.atan2(rot[2, 1], rot[2, 2])
    yaw = math.atan2(rot[1, 0], rot[0, 0])
    return np.array([matrix[0, 3], matrix[1, 3], matrix[2, 3], roll, pitch, yaw])


def _geometry(node):
    geo = _Record(box=None, sphere=None, cylinder=None, mesh=None)
    if node is None:
        return geo
    box, sphere, cyl = node.find("box"), node.find("sphere"), node.find("cylinder")
    if box is not None:
        geo.box = _Record(size=_floats(box.get("size"), (1.0, 1.0, 1.0)))
    if sphere is not None:
        geo.sphere = _Record(radius=float(sphere.get("radius")))
    if cyl is not None:
        geo.cylinder = _Record(radius=float(cyl.get("radius")), length=float(cyl.get("length")))
    return geo


class URDF:
    def __init__(self):
        self.links, self.joints, self.link_map = [], [], {}

    @staticmethod
    def load(filename):
        robot = URDF()
        root = ET.parse(filename).getroot()
        for link_node in root.findall("link"):
            collisions = [
                _Record(origin=_pose_matrix(c), geometry=_geometry(c.find("geometry")))
                for c in link_node.findall("collision")
            ]
            link = _Record(name=link_node.get("name"), collisions=collisions)
            robot.links.append(link)
            robot.link_map[link.name] = link
        for joint_node in root.findall("joint"):
            axis_node = joint_node.find("axis")
            limit_node = joint_node.find("limit")
            dyn_node = joint_node.find("dynamics")
            limit = None
            if limit_node is not None:
                lo, hi = limit_node.get("lower"), limit_node.get("upper")
                limit = _Record(lower=None if lo is None else float(lo), upper=None if hi is None else float(hi))
            dynamics = None
            if dyn_node is not None:
                dynamics = _Record(damping=float(dyn_node.get("damping", 0.0)))
            robot.joints.append(
                _Record(
                    name=joint_node.get("name"),
                    joint_type=joint_node.get("type"),
                    parent=joint_node.find("parent").get("link"),
                    child=joint_node.find("child").get("link"),
                    axis=_floats(None if axis_node is None else axis_node.get("xyz"), (1.0, 0.0, 0.0)),
                    origin=_pose_matrix(joint_node),
                    limit=limit,
                    dynamics=dynamics,
                )
            )
        return robot
