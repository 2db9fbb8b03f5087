"""``Model`` / ``State`` / ``ModelBuilder`` behind ``import dflex as df``.

Interface mirror of the reference ``dflex/dflex/model.py`` for the articulated-rigid-body path:
same class and method names, argument meaning and defaults, and the SAME finalized tensors (field
names, dtypes, env-concatenated layouts -- they are the data ABI of the simulation step and are
compared bit-for-bit with the reference's in tests/test_model_builder_parity.py).  Independently
written; particle / cloth / FEM builders (no DiffRL env uses them) are intentionally absent and
raise ``NotImplementedError``.

Reference anchors: State model.py:115, Model :180 (state() :338, collide() :424), ModelBuilder :521
(add_link :644, add_muscle :806, add_shape_* :837-968, _update_body_mass :1621, finalize :1646).
"""
import math

import numpy as np
import torch

from .util import (quat_identity, spatial_matrix_from_inertia, transform, transform_expand, transform_flatten_list,
                   transform_inertia, transform_point)

# shape geometry types (model.py:26-32)
GEO_SPHERE, GEO_BOX, GEO_CAPSULE, GEO_MESH, GEO_SDF, GEO_PLANE, GEO_NONE = range(7)
# joint types (model.py:35-39)
JOINT_PRISMATIC, JOINT_REVOLUTE, JOINT_BALL, JOINT_FIXED, JOINT_FREE = range(5)

_COORDS = {JOINT_PRISMATIC: 1, JOINT_REVOLUTE: 1, JOINT_BALL: 4, JOINT_FIXED: 0, JOINT_FREE: 7}
_DOFS = {JOINT_PRISMATIC: 1, JOINT_REVOLUTE: 1, JOINT_BALL: 3, JOINT_FIXED: 0, JOINT_FREE: 6}


class Mesh:
    """Triangle mesh with density-1 mass properties (used for mesh collision shapes)."""

    def __init__(self, vertices, indices):
        self.vertices, self.indices = vertices, indices
        verts = np.asarray(vertices, dtype=np.float64)
        com = np.mean(verts, 0)
        alpha = math.sqrt(5.0) / 5.0
        inertia, mass = np.zeros((3, 3)), 0.0
        for t in range(len(indices) // 3):
            p, q, r = (verts[indices[t * 3 + k]] for k in range(3))
            mid = (com + p + q + r) / 4.0
            volume = np.linalg.det(np.array((p - com, q - com, r - com)).T) / 6.0
            for corner in (p, q, r, com):   # order-2 quadrature on the tetrahedron (com, p, q, r)
                d = (mid + (corner - mid) * alpha) - com
                inertia += 0.25 * volume * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
                mass += 0.25 * volume
        self.I, self.mass, self.com = inertia, mass, com


class State:
    """Time-varying data of a model; made by :meth:`Model.state` and by the integrator."""

    def __init__(self):
        self.particle_count = 0
        self.link_count = 0

    def flatten(self):
        return [v for v in self.__dict__.values() if torch.is_tensor(v)]


class Model:
    """Static description of the simulation; tensors are flat and env-concatenated."""

    def __init__(self, adapter):
        self.adapter = adapter
        self.particle_count = self.joint_coord_count = self.joint_dof_count = 0
        self.link_count = self.shape_count = self.contact_count = 0
        self.tri_count = self.tet_count = self.edge_count = self.spring_count = 0
        self.articulation_count = self.muscle_count = 0
        self.gravity = torch.tensor((0.0, -9.8, 0.0), dtype=torch.float32, device=adapter)
        self.ground = True
        self.enable_tri_collisions = False
        self.contact_distance, self.contact_ke, self.contact_kd = 0.1, 1.0e3, 0.0
        self.contact_kf, self.contact_mu, self.particle_radius = 1.0e3, 0.5, 0.1
        self._engine = None
        self._engine_key = None

    # ------------------------------------------------------------------ state
    def state(self):
        """A State initialised with the model's initial configuration (model.py:338-392).  Derived
        fields (body_X_sc, ...) are allocated lazily by the integrator, not per call."""
        s = State()
        s.particle_count, s.link_count = self.particle_count, self.link_count
        if self.link_count:
            s.joint_q = torch.clone(self.joint_q)
            s.joint_qd = torch.clone(self.joint_qd)
            s.joint_act = torch.zeros_like(self.joint_qd)
            s.joint_q.requires_grad = True
            s.joint_qd.requires_grad = True
        return s

    def alloc_mass_matrix(self):
        """Kept for API compatibility: the joint-space inertia lives in shared memory inside the
        fused kernels (the reference allocates dense M, J, P, H, L here, model.py:394-405)."""
        return None

    def flatten(self):
        return [v for v in self.__dict__.values() if torch.is_tensor(v)]

    # ------------------------------------------------------------------ contacts
    def collide(self, state=None):
        """Static ground-contact point list (model.py:424-515): sphere -> centre, capsule -> the two cap
        centres on local x, box -> 8 corners in the reference's order, mesh -> every vertex."""
        body0, point, dist, mat = [], [], [], []
        geo_type = self.shape_geo_type.tolist()
        geo_scale = self.shape_geo_scale.tolist()
        shape_body = self.shape_body.tolist()
        shape_tf = self.shape_transform.tolist()

        def add(i, local, d):
            body0.append(shape_body[i])
            point.append(transform_point(transform_expand(shape_tf[i]), np.array(local)))
            dist.append(d)
            mat.append(i)

        for i in range(self.shape_count):
            t, sc = geo_type[i], geo_scale[i]
            if t == GEO_SPHERE:
                add(i, (0.0, 0.0, 0.0), sc[0])
            elif t == GEO_CAPSULE:
                add(i, (-sc[1], 0.0, 0.0), sc[0])
                add(i, (sc[1], 0.0, 0.0), sc[0])
            elif t == GEO_BOX:
                for sz in (-1.0, 1.0):
                    for sy in (-1.0, 1.0):
                        for sx in (-1.0, 1.0):
                            add(i, (sx * sc[0], sy * sc[1], sz * sc[2]), 0.0)
            elif t == GEO_MESH:
                for v in self.shape_geo_src[i].vertices:
                    add(i, (v[0] * sc[0], v[1] * sc[1], v[2] * sc[2]), 0.0)
        dev = self.adapter
        self.contact_body0 = torch.tensor(body0, dtype=torch.int32, device=dev)
        self.contact_body1 = torch.full((len(body0),), -1, dtype=torch.int32, device=dev)
        self.contact_point0 = torch.tensor(np.array(point, dtype=np.float64).reshape(-1, 3), dtype=torch.float32, device=dev)
        self.contact_dist = torch.tensor(dist, dtype=torch.float32, device=dev)
        self.contact_material = torch.tensor(mat, dtype=torch.int32, device=dev)
        self.contact_count = len(body0)
        self._engine = None   # contact list is part of the device-side pack


class ModelBuilder:
    """Incremental scene description in plain Python lists; :meth:`finalize` makes the tensors."""

    def __init__(self):
        self.particle_q, self.particle_qd, self.particle_mass = [], [], []
        self.shape_transform, self.shape_body, self.shape_geo_type = [], [], []
        self.shape_geo_scale, self.shape_geo_src, self.shape_materials = [], [], []
        self.geo_meshes, self.geo_sdfs = [], []
        self.muscle_start, self.muscle_params, self.muscle_activation = [], [], []
        self.muscle_links, self.muscle_points = [], []
        self.joint_parent, self.joint_child, self.joint_axis = [], [], []
        self.joint_X_pj, self.joint_X_cm = [], []
        self.joint_q_start, self.joint_qd_start, self.joint_type = [], [], []
        self.joint_armature, self.joint_target_ke, self.joint_target_kd, self.joint_target = [], [], [], []
        self.joint_limit_lower, self.joint_limit_upper, self.joint_limit_ke, self.joint_limit_kd = [], [], [], []
        self.joint_q, self.joint_qd, self.joint_qdd, self.joint_tau, self.joint_u = [], [], [], [], []
        self.body_mass, self.body_inertia, self.body_com = [], [], []
        self.articulation_start = []

    # ------------------------------------------------------------------ articulations
    def add_articulation(self):
        self.articulation_start.append(len(self.joint_type))
        return len(self.articulation_start) - 1

    def add_link(self, parent, X_pj, axis, type, armature=0.01, stiffness=0.0, damping=0.0,
                 limit_lower=-1.0e3, limit_upper=1.0e3, limit_ke=100.0, limit_kd=10.0,
                 com=np.zeros(3), I_m=np.zeros((3, 3)), m=0.0):
        """Add a rigid link connected to ``parent`` (-1 = world) by a joint of ``type`` located at
        ``X_pj`` in the parent frame.  Returns the link index.  (``com``/``I_m``/``m`` are accepted and,
        like in the reference, ignored: mass comes from the shapes.)"""
        self.joint_type.append(type)
        self.joint_axis.append(np.array(axis))
        self.joint_parent.append(parent)
        self.joint_X_pj.append(X_pj)
        self.joint_target_ke.append(stiffness)
        self.joint_target_kd.append(damping)
        self.joint_limit_ke.append(limit_ke)
        self.joint_limit_kd.append(limit_kd)
        self.joint_q_start.append(len(self.joint_q))
        self.joint_qd_start.append(len(self.joint_qd))
        if type in (JOINT_PRISMATIC, JOINT_REVOLUTE):
            self.joint_q.append(0.0)
            self.joint_qd.append(0.0)
            self.joint_target.append(0.0)
            self.joint_armature.append(armature)
            self.joint_limit_lower.append(limit_lower)
            self.joint_limit_upper.append(limit_upper)
        elif type == JOINT_BALL:
            self.joint_q += [0.0, 0.0, 0.0, 1.0]
            self.joint_qd += [0.0, 0.0, 0.0]
            self.joint_target += [0.0, 0.0, 0.0, 0.0]
            self.joint_armature += [armature] * 3
            self.joint_limit_lower += [limit_lower] * 3 + [0.0]
            self.joint_limit_upper += [limit_upper] * 3 + [0.0]
        elif type == JOINT_FREE:
            self.joint_q += [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]
            self.joint_qd += [0.0] * 6
            self.joint_armature += [0.0] * 6
            self.joint_target += [0.0] * 7
            self.joint_limit_lower += [0.0] * 7
            self.joint_limit_upper += [0.0] * 7
        elif type != JOINT_FIXED:
            raise ValueError("unknown joint type %r" % (type,))
        self.body_inertia.append(np.zeros((3, 3)))
        self.body_mass.append(0.0)
        self.body_com.append(np.zeros(3))
        return len(self.joint_type) - 1

    def add_muscle(self, links, positions, f0, lm, lt, lmax, pen):
        self.muscle_start.append(len(self.muscle_links))
        self.muscle_params.append((f0, lm, lt, lmax, pen))
        self.muscle_activation.append(0.0)
        for link, pos in zip(links, positions):
            self.muscle_links.append(link)
            self.muscle_points.append(pos)
        return len(self.muscle_start) - 1

    # ------------------------------------------------------------------ shapes
    def add_shape_plane(self, plane=(0.0, 1.0, 0.0, 0.0), ke=1.0e5, kd=1000.0, kf=1000.0, mu=0.5):
        self._add_shape(-1, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), GEO_PLANE, plane, None, 0.0, ke, kd, kf, mu)

    def add_shape_sphere(self, body, pos=(0.0, 0.0, 0.0), rot=(0.0, 0.0, 0.0, 1.0), radius=1.0, density=1000.0,
                         ke=1.0e5, kd=1000.0, kf=1000.0, mu=0.5):
        self._add_shape(body, pos, rot, GEO_SPHERE, (radius, 0.0, 0.0, 0.0), None, density, ke, kd, kf, mu)

    def add_shape_box(self, body, pos=(0.0, 0.0, 0.0), rot=(0.0, 0.0, 0.0, 1.0), hx=0.5, hy=0.5, hz=0.5,
                      density=1000.0, ke=1.0e5, kd=1000.0, kf=1000.0, mu=0.5):
        self._add_shape(body, pos, rot, GEO_BOX, (hx, hy, hz, 0.0), None, density, ke, kd, kf, mu)

    def add_shape_capsule(self, body, pos=(0.0, 0.0, 0.0), rot=(0.0, 0.0, 0.0, 1.0), radius=1.0, half_width=0.5,
                          density=1000.0, ke=1.0e5, kd=1000.0, kf=1000.0, mu=0.5):
        self._add_shape(body, pos, rot, GEO_CAPSULE, (radius, half_width, 0.0, 0.0), None, density, ke, kd, kf, mu)

    def add_shape_mesh(self, body, pos=(0.0, 0.0, 0.0), rot=(0.0, 0.0, 0.0, 1.0), mesh=None, scale=(1.0, 1.0, 1.0),
                       density=1000.0, ke=1.0e5, kd=1000.0, kf=1000.0, mu=0.5):
        self._add_shape(body, pos, rot, GEO_MESH, (scale[0], scale[1], scale[2], 0.0), mesh, density, ke, kd, kf, mu)

    def _add_shape(self, body, pos, rot, type, scale, src, density, ke, kd, kf, mu):
        self.shape_body.append(body)
        self.shape_transform.append(transform(pos, rot))
        self.shape_geo_type.append(type)
        self.shape_geo_scale.append((scale[0], scale[1], scale[2]))
        self.shape_geo_src.append(src)
        self.shape_materials.append((ke, kd, kf, mu))
        m, I = self._compute_shape_mass(type, scale, src, density)
        self._update_body_mass(body, m, I, np.array(pos), np.array(rot))

    # mass properties of primitive shapes about their own centre
    def compute_sphere_inertia(self, density, r):
        m = density * (4.0 / 3.0 * math.pi * r * r * r)
        Ia = 2.0 / 5.0 * m * r * r
        return m, np.array([[Ia, 0.0, 0.0], [0.0, Ia, 0.0], [0.0, 0.0, Ia]])

    def compute_capsule_inertia(self, density, r, l):
        ms = density * (4.0 / 3.0) * math.pi * r * r * r
        mc = density * math.pi * r * r * l
        Ia = mc * (0.25 * r * r + (1.0 / 12.0) * l * l) + ms * (0.4 * r * r + 0.375 * r * l + 0.25 * l * l)
        Ib = (mc * 0.5 + ms * 0.4) * r * r
        return ms + mc, np.array([[Ib, 0.0, 0.0], [0.0, Ia, 0.0], [0.0, 0.0, Ia]])

    def compute_box_inertia(self, density, w, h, d):
        m = density * (w * h * d)
        Ia = 1.0 / 12.0 * m * (h * h + d * d)
        Ib = 1.0 / 12.0 * m * (w * w + d * d)
        Ic = 1.0 / 12.0 * m * (w * w + h * h)
        return m, np.array([[Ia, 0.0, 0.0], [0.0, Ib, 0.0], [0.0, 0.0, Ic]])

    def _compute_shape_mass(self, type, scale, src, density):
        if density == 0:
            return 0, np.zeros((3, 3))
        if type == GEO_SPHERE:
            return self.compute_sphere_inertia(density, scale[0])
        if type == GEO_BOX:
            return self.compute_box_inertia(density, scale[0] * 2.0, scale[1] * 2.0, scale[2] * 2.0)
        if type == GEO_CAPSULE:
            return self.compute_capsule_inertia(density, scale[0], scale[1] * 2.0)
        if type == GEO_MESH:
            s = scale[0]
            return density * src.mass * s * s * s, density * src.I * s * s * s * s * s
        return 0, np.zeros((3, 3))

    def _update_body_mass(self, i, m, I, p, q):
        """Fold a shape's (m, I) at offset p / rotation q into link i's running mass, COM and inertia."""
        if i == -1:
            return
        new_mass = self.body_mass[i] + m
        if new_mass == 0.0:
            return
        new_com = (self.body_com[i] * self.body_mass[i] + p * m) / new_mass
        com_offset = new_com - self.body_com[i]
        shape_offset = new_com - p
        new_inertia = (transform_inertia(self.body_mass[i], self.body_inertia[i], com_offset, quat_identity()) +
                       transform_inertia(m, I, shape_offset, q))
        self.body_mass[i], self.body_inertia[i], self.body_com[i] = new_mass, new_inertia, new_com

    # ------------------------------------------------------------------ unsupported (no DiffRL env uses them)
    def _unsupported(self, *a, **k):
        raise NotImplementedError("particles / cloth / FEM are outside the articulated-rigid-body hot path")

    add_particle = add_spring = add_triangle = add_tetrahedron = add_edge = _unsupported
    add_cloth_grid = add_cloth_mesh = add_soft_grid = add_soft_mesh = _unsupported

    # ------------------------------------------------------------------ finalize
    def finalize(self, adapter):
        """Transfer the description to torch tensors on ``adapter`` (model.py:1646-1879)."""
        m = Model(adapter)
        f32 = dict(dtype=torch.float32, device=adapter)
        i32 = dict(dtype=torch.int32, device=adapter)

        def farr(x, shape=None):
            a = np.asarray(x, dtype=np.float64)
            if shape is not None:
                a = a.reshape(shape)
            return torch.tensor(a, **f32)

        empty_f = lambda: torch.zeros((0,), **f32)
        empty_i = lambda: torch.zeros((0,), **i32)
        # particles etc.: present (empty) so that generic code can read the fields
        m.particle_q = m.particle_qd = m.particle_mass = m.particle_inv_mass = empty_f()
        for name in ("spring_rest_length", "spring_stiffness", "spring_damping", "spring_control", "tri_poses",
                     "tri_activations", "edge_rest_angle", "tet_poses", "tet_activations", "tet_materials"):
            setattr(m, name, empty_f())
        for name in ("spring_indices", "tri_indices", "edge_indices", "tet_indices"):
            setattr(m, name, empty_i())

        m.shape_transform = farr(transform_flatten_list(self.shape_transform), (-1, 7))
        m.shape_body = torch.tensor(self.shape_body, **i32)
        m.shape_geo_type = torch.tensor(self.shape_geo_type, **i32)
        m.shape_geo_src = self.shape_geo_src
        m.shape_geo_scale = farr(self.shape_geo_scale, (-1, 3))
        m.shape_materials = farr(self.shape_materials, (-1, 4))

        muscle_count = len(self.muscle_start)
        muscle_start = list(self.muscle_start) + [len(self.muscle_links)]
        m.muscle_start = torch.tensor(muscle_start, **i32)
        m.muscle_params = farr(self.muscle_params, (-1, 5)) if muscle_count else empty_f()
        m.muscle_links = torch.tensor(self.muscle_links, **i32)
        m.muscle_points = farr(self.muscle_points, (-1, 3)) if muscle_count else empty_f()
        m.muscle_activation = torch.tensor(self.muscle_activation, **f32)

        link_count = len(self.joint_type)
        body_I_m = [spatial_matrix_from_inertia(self.body_inertia[i], self.body_mass[i]) for i in range(link_count)]
        body_X_cm = [transform(self.body_com[i], quat_identity()) for i in range(link_count)]
        m.body_I_m = farr(body_I_m, (-1, 6, 6))

        joint_q_start = list(self.joint_q_start) + [len(self.joint_q)]
        joint_qd_start = list(self.joint_qd_start) + [len(self.joint_qd)]
        art_start = list(self.articulation_start) + [link_count]
        articulation_count = len(self.articulation_start)
        J_start, M_start, H_start, M_rows, H_rows, J_rows, J_cols, dof_start, coord_start = ([] for _ in range(9))
        m.J_size = m.M_size = m.H_size = 0
        for a in range(articulation_count):
            first, last = art_start[a], art_start[a + 1]
            joints = last - first
            dofs = joint_qd_start[last] - joint_qd_start[first]
            J_start.append(m.J_size); M_start.append(m.M_size); H_start.append(m.H_size)
            dof_start.append(joint_qd_start[first]); coord_start.append(joint_q_start[first])
            M_rows.append(joints * 6); H_rows.append(dofs); J_rows.append(joints * 6); J_cols.append(dofs)
            m.J_size += 6 * joints * dofs
            m.M_size += 6 * joints * 6 * joints
            m.H_size += dofs * dofs
        m.articulation_joint_start = torch.tensor(art_start, **i32)
        m.articulation_J_start = torch.tensor(J_start, **i32)
        m.articulation_M_start = torch.tensor(M_start, **i32)
        m.articulation_H_start = torch.tensor(H_start, **i32)
        m.articulation_M_rows = torch.tensor(M_rows, **i32)
        m.articulation_H_rows = torch.tensor(H_rows, **i32)
        m.articulation_J_rows = torch.tensor(J_rows, **i32)
        m.articulation_J_cols = torch.tensor(J_cols, **i32)
        m.articulation_dof_start = torch.tensor(dof_start, **i32)
        m.articulation_coord_start = torch.tensor(coord_start, **i32)

        m.joint_q = farr(self.joint_q)
        m.joint_qd = farr(self.joint_qd)
        m.joint_type = torch.tensor(self.joint_type, **i32)
        m.joint_parent = torch.tensor(self.joint_parent, **i32)
        m.joint_X_pj = farr(transform_flatten_list(self.joint_X_pj), (-1, 7))
        m.joint_X_cm = farr(transform_flatten_list(body_X_cm), (-1, 7))
        m.joint_axis = farr(self.joint_axis, (-1, 3))
        m.joint_q_start = torch.tensor(joint_q_start, **i32)
        m.joint_qd_start = torch.tensor(joint_qd_start, **i32)
        m.joint_armature = farr(self.joint_armature)
        m.joint_target = farr(self.joint_target)
        m.joint_target_ke = farr(self.joint_target_ke)
        m.joint_target_kd = farr(self.joint_target_kd)
        m.joint_limit_lower = farr(self.joint_limit_lower)
        m.joint_limit_upper = farr(self.joint_limit_upper)
        m.joint_limit_ke = farr(self.joint_limit_ke)
        m.joint_limit_kd = farr(self.joint_limit_kd)

        m.particle_count = 0
        m.articulation_count = articulation_count
        m.joint_coord_count = len(self.joint_q)
        m.joint_dof_count = len(self.joint_qd)
        m.muscle_count = muscle_count
        m.link_count = link_count
        m.shape_count = len(self.shape_geo_type)
        m.tri_count = m.tet_count = m.edge_count = m.spring_count = 0
        m.contact_count = 0
        m.geo_meshes, m.geo_sdfs = self.geo_meshes, self.geo_sdfs
        m.ground = True
        m.enable_tri_collisions = False
        m.gravity = torch.tensor((0.0, -9.8, 0.0), **f32)
        return m


def model_from_articulation(arrays, num_envs, device, ground=True, gravity=(0.0, -9.81, 0.0)):
    """Build a finalized ``Model`` for ``num_envs`` copies of ONE articulation given as a dict of numpy
    arrays with the reference field names (what ``oracle/make_golden.py`` exports per env, i.e. the
    "build once and tile" start-up of SURVEY.md section 8f-3): O(1) Python work instead of the
    reference's per-env asset parsing loop.  Index fields are offset per environment exactly as
    ``ModelBuilder`` would produce them."""
    m = Model(device)
    n = int(num_envs)
    f32 = dict(dtype=torch.float32, device=device)
    i32 = dict(dtype=torch.int32, device=device)
    L = int(arrays["joint_type"].shape[0])
    Q = int(arrays["joint_q_start"][-1]); D = int(arrays["joint_qd_start"][-1])
    S = int(arrays["shape_body"].shape[0]); C = int(arrays["contact_body0"].shape[0])
    W = int(arrays["muscle_links"].shape[0]); M = int(arrays["muscle_start"].shape[0]) - 1

    def tile_f(name):
        a = np.asarray(arrays[name], dtype=np.float32)
        return torch.tensor(np.tile(a, (n,) + (1,) * (a.ndim - 1)), **f32)

    def tile_i(name, per_env_offset=0, sentinel=False):
        a = np.asarray(arrays[name], dtype=np.int64)
        body = a[:-1] if sentinel else a
        parts = [np.where(body >= 0, body + e * per_env_offset, body) for e in range(n)]
        out = np.concatenate(parts) if parts else body
        if sentinel:
            out = np.concatenate([out, [a[-1] + (n - 1) * per_env_offset]])
        return torch.tensor(out, **i32)

    for name in ("joint_X_pj", "joint_X_cm", "joint_axis", "body_I_m", "joint_target_ke", "joint_target_kd",
                 "joint_limit_ke", "joint_limit_kd", "joint_target", "joint_limit_lower", "joint_limit_upper",
                 "joint_armature", "joint_q", "joint_qd", "shape_transform", "shape_geo_scale", "shape_materials",
                 "contact_point0", "contact_dist", "muscle_points"):
        setattr(m, name, tile_f(name))
    m.joint_type = tile_i("joint_type")
    m.joint_parent = tile_i("joint_parent", L)
    m.joint_q_start = tile_i("joint_q_start", Q, sentinel=True)
    m.joint_qd_start = tile_i("joint_qd_start", D, sentinel=True)
    m.shape_body = tile_i("shape_body", L)
    m.shape_geo_type = tile_i("shape_geo_type")
    m.shape_geo_src = [None] * (S * n)
    m.contact_body0 = tile_i("contact_body0", L)
    m.contact_body1 = torch.full((C * n,), -1, **i32)
    m.contact_material = tile_i("contact_material", S)
    m.muscle_start = tile_i("muscle_start", W, sentinel=True)
    m.muscle_links = tile_i("muscle_links", L)
    m.muscle_params = tile_f("muscle_params") if "muscle_params" in arrays else torch.zeros((M * n, 5), **f32)
    m.muscle_activation = torch.zeros((M * n,), **f32)
    m.articulation_joint_start = torch.arange(0, (n + 1) * L, L, **i32)
    m.articulation_dof_start = torch.arange(0, n * D, D, **i32)
    m.articulation_coord_start = torch.arange(0, n * Q, Q, **i32)
    m.articulation_count, m.link_count = n, L * n
    m.joint_coord_count, m.joint_dof_count = Q * n, D * n
    m.shape_count, m.contact_count, m.muscle_count = S * n, C * n, M * n
    m.ground = bool(ground)
    m.gravity = torch.tensor(gravity, **f32)
    m._articulation_arrays = arrays   # lets the integrator skip the homogeneity scan
    return m
