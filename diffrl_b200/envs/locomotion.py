"""Concrete differentiable environments: Ant, Humanoid, SNU humanoid (muscle-actuated), CartPole
swing-up, Hopper, HalfCheetah.

Interface mirrors of the reference ``envs/{ant,humanoid,snu_humanoid,cartpole_swing_up,hopper,cheetah}.py``:
same class names, constructor signatures, observation layouts, rewards, termination rules, reset
distributions and constants (cited per class); obs / reward / gradient parity against the
reference's own rollouts is tested in tests/test_gpu_envs.py.  The simulation step itself is the
fused kernel behind ``df.sim.SemiImplicitIntegrator``.
"""
import ctypes
import math

import numpy as np
import torch

import diffrl_b200.dflex_api as df

from . import torch_ops as tu
from .base import DFlexEnv


class _FreeRootWalker(DFlexEnv):
    """Shared pieces of the three free-floating walkers (torso pose/velocity features, heading and
    up-vector projections, start-pose reset)."""

    nan_guard = False
    target_x = 10000.0

    def _init_common(self, start_height, start_rot):
        n, dev = self.num_envs, self.device
        self.start_rot = start_rot
        self.start_rotation = tu.to_torch(start_rot, device=dev)
        unit = lambda v: tu.to_torch(v, device=dev).repeat((n, 1))
        self.x_unit_tensor, self.y_unit_tensor, self.z_unit_tensor = unit([1, 0, 0]), unit([0, 1, 0]), unit([0, 0, 1])
        self.up_vec = self.y_unit_tensor.clone()
        self.heading_vec = self.x_unit_tensor.clone()
        self.inv_start_rot = tu.quat_conjugate(self.start_rotation).repeat((n, 1))
        self.basis_vec0 = self.heading_vec.clone()
        self.basis_vec1 = self.up_vec.clone()
        self.targets = tu.to_torch([self.target_x, 0.0, 0.0], device=dev).repeat((n, 1))
        self.env_dist = 0.0
        self.start_pos = tu.to_torch([[0.0, start_height, 0.0]] * n, device=dev)

    # ---- fused epilogue -------------------------------------------------------------------------
    def _walker_params(self):
        from ..env_ops import DfxWalkerParams
        p = DfxWalkerParams()
        p.num_q, p.num_qd, p.num_act, p.num_obs = self.num_joint_q, self.num_joint_qd, self.num_actions, self.num_observations
        p.obs_has_actions = int(self.obs_has_actions)
        p.height_mode, p.action_penalty_abs = int(self.height_mode), int(self.action_penalty_abs)
        p.early_termination = int(getattr(self, "early_termination", True))
        p.check_invalid, p.zero_reward_on_invalid = int(self.check_invalid), int(self.zero_reward_on_invalid)
        p.episode_length = int(self.episode_length)
        p.joint_vel_scale = float(self.joint_vel_obs_scaling)
        p.termination_height = float(self.termination_height)
        p.termination_tolerance = float(getattr(self, "termination_tolerance", 0.0))
        p.height_rew_scale = float(getattr(self, "height_rew_scale", 1.0))
        p.action_penalty = float(self.action_penalty)
        tgt = (self.targets[0] + self.start_pos[0]).tolist()
        p.target = (ctypes.c_float * 3)(*tgt)
        p.inv_start_rot = (ctypes.c_float * 4)(*self.inv_start_rot[0].tolist())
        p.basis_heading = (ctypes.c_float * 3)(*self.basis_vec0[0].tolist())
        p.basis_up = (ctypes.c_float * 3)(*self.basis_vec1[0].tolist())
        return p

    obs_has_actions, height_mode, action_penalty_abs = True, 0, False
    check_invalid, zero_reward_on_invalid = False, False
    fused_epilogue = True

    def _fused(self, want_reward):
        from ..env_ops import WalkerObsFunction
        if getattr(self, "_wparams", None) is None:
            self._wparams = self._walker_params()
        return WalkerObsFunction.apply(self._wparams, self.num_envs, want_reward, self.progress_buf,
                                       self.state.joint_q, self.state.joint_qd, self.actions)

    def _observe_and_reward(self):
        if self.fused_epilogue and torch.device(self.device).type == "cuda":
            self.obs_buf, self.rew_buf, self.reset_buf = self._fused(True)
        else:
            self.calculateObservations()
            self.calculateReward()

    def _start_state(self):
        """(start_q [n, Q], start_qd [n, D]): the state terminated environments restart from, for all n rows
        (same distributions as _reset_state / reference ant.py:192-228, sampled for every row and used by mask)."""
        n = self.num_envs
        if getattr(self, "_start_q_full", None) is None:
            self._start_q_full = torch.cat([self.start_pos, self.start_rotation.expand(n, 4),
                                            self.start_joint_q.expand(n, -1)], dim=-1).contiguous()
            self._zero_qd = torch.zeros((n, self.num_joint_qd), device=self.device)
            self._zero_act = torch.zeros((n, self.num_actions), device=self.device)
        start_q, start_qd = self._start_q_full, self._zero_qd
        if self.stochastic_init:
            dev = self.device
            start_q = start_q.clone()
            start_q[:, 0:3] = start_q[:, 0:3] + 0.1 * (torch.rand(size=(n, 3), device=dev) - 0.5) * 2.0
            angle = (torch.rand(n, device=dev) - 0.5) * np.pi / 12.0
            axis = torch.nn.functional.normalize(torch.rand((n, 3), device=dev) - 0.5)
            start_q[:, 3:7] = tu.quat_mul(start_q[:, 3:7], tu.quat_from_angle_axis(angle, axis))
            if self.randomize_joints:
                start_q[:, 7:] = start_q[:, 7:] + 0.2 * (torch.rand(size=(n, self.num_joint_q - 7), device=dev) - 0.5) * 2.0
            start_qd = 0.5 * (torch.rand(size=(n, self.num_joint_qd), device=dev) - 0.5)
        return start_q, start_qd

    # ---- fused step (DFlexEnv._fused_step): action map, simulation step and transition in ONE launch (dfx_env_step_forward) ----
    fused_transition = True

    def _action_map(self):
        """(width, offset, pre_scale, pre_bias, drive_scale, strength [A], is_muscle) of dfx_action_map_forward."""
        raise NotImplementedError

    def _transition_params(self):
        return self._walker_params()

    def step(self, actions):
        if not (self.fused_transition and self.fused_epilogue and self.sync_free_reset
                and torch.device(self.device).type == "cuda"):
            return super().step(actions)
        return self._fused_step(actions)

    def _torso_features(self):
        q = self.state.joint_q.view(self.num_envs, -1)
        qd = self.state.joint_qd.view(self.num_envs, -1)
        torso_pos, torso_rot = q[:, 0:3], q[:, 3:7]
        ang_vel = qd[:, 0:3]
        # spatial twist at the world origin -> velocity of the torso origin (ant.py:272-273)
        lin_vel = qd[:, 3:6] - torch.cross(torso_pos, ang_vel, dim=-1)
        to_target = self.targets + self.start_pos - torso_pos
        to_target[:, 1] = 0.0
        target_dirs = tu.normalize(to_target)
        torso_quat = tu.quat_mul(torso_rot, self.inv_start_rot)
        up_vec = tu.quat_rotate(torso_quat, self.basis_vec1)
        heading_vec = tu.quat_rotate(torso_quat, self.basis_vec0)
        heading_proj = (heading_vec * target_dirs).sum(dim=-1).unsqueeze(-1)
        return q, qd, torso_pos, torso_rot, lin_vel, ang_vel, up_vec[:, 1:2], heading_proj

    def _reset_state(self, env_ids):
        q = self.state.joint_q.view(self.num_envs, -1)
        qd = self.state.joint_qd.view(self.num_envs, -1)
        q[env_ids, 0:3] = self.start_pos[env_ids, :].clone()
        q[env_ids, 3:7] = self.start_rotation.clone()
        q[env_ids, 7:] = self.start_joint_q.clone()
        qd[env_ids, :] = 0.0
        if self.stochastic_init:
            k, dev = len(env_ids), self.device
            q[env_ids, 0:3] = q[env_ids, 0:3] + 0.1 * (torch.rand(size=(k, 3), device=dev) - 0.5) * 2.0
            angle = (torch.rand(k, device=dev) - 0.5) * np.pi / 12.0
            axis = torch.nn.functional.normalize(torch.rand((k, 3), device=dev) - 0.5)
            q[env_ids, 3:7] = tu.quat_mul(q[env_ids, 3:7], tu.quat_from_angle_axis(angle, axis))
            if self.randomize_joints:
                q[env_ids, 7:] = q[env_ids, 7:] + 0.2 * (torch.rand(size=(k, self.num_joint_q - 7), device=dev) - 0.5) * 2.0
            qd[env_ids, :] = 0.5 * (torch.rand(size=(k, self.num_joint_qd), device=dev) - 0.5)

    randomize_joints = True


class AntEnv(_FreeRootWalker):
    """reference envs/ant.py: 37 obs, 8 actions, 16 substeps, termination height 0.27."""

    def __init__(self, render=False, device="cuda:0", num_envs=4096, seed=0, episode_length=1000, no_grad=True,
                 stochastic_init=False, MM_caching_frequency=1, early_termination=True):
        super().__init__(num_envs, 37, 8, episode_length, MM_caching_frequency, seed, no_grad, render, device)
        self.stochastic_init, self.early_termination = stochastic_init, early_termination
        self.dt, self.sim_substeps = 1.0 / 60.0, 16
        self.sim_dt, self.ground = self.dt, True
        self.num_joint_q, self.num_joint_qd = 15, 14
        self._init_common(0.75, df.quat_from_axis_angle((1.0, 0.0, 0.0), -math.pi * 0.5))
        self.start_joint_q = tu.to_torch([0.0, 1.0, 0.0, -1.0, 0.0, -1.0, 0.0, 1.0], device=device)
        self.start_joint_target = self.start_joint_q.clone()
        self._build_model("AntEnv")
        self.termination_height, self.action_strength = 0.27, 200.0
        self.action_penalty, self.joint_vel_obs_scaling = 0.0, 0.1

    def _apply_actions(self, actions):
        self.state.joint_act.view(self.num_envs, -1)[:, 6:] = actions * self.action_strength

    def _action_map(self):
        strength = torch.full((self.num_actions,), float(self.action_strength), device=self.device)
        return self.num_joint_qd, 6, 1.0, 0.0, 1.0, strength, False

    def calculateObservations(self):
        if self.fused_epilogue and torch.device(self.device).type == "cuda":
            self.obs_buf = self._fused(False)
            return
        q, qd, pos, rot, lin_vel, ang_vel, up, heading = self._torso_features()
        self.obs_buf = torch.cat([pos[:, 1:2], rot, lin_vel, ang_vel, q[:, 7:], self.joint_vel_obs_scaling * qd[:, 6:],
                                  up, heading, self.actions.clone()], dim=-1)

    def calculateReward(self):
        o = self.obs_buf
        self.rew_buf = o[:, 5] + 0.1 * o[:, 27] + o[:, 28] + (o[:, 0] - self.termination_height) + \
            torch.sum(self.actions ** 2, dim=-1) * self.action_penalty
        if self.early_termination:
            self.reset_buf = torch.where(o[:, 0] < self.termination_height, torch.ones_like(self.reset_buf), self.reset_buf)
        self.reset_buf = torch.where(self.progress_buf > self.episode_length - 1, torch.ones_like(self.reset_buf), self.reset_buf)


class HumanoidEnv(_FreeRootWalker):
    """reference envs/humanoid.py: 76 obs, 21 actions, 48 substeps, termination height 0.74."""

    nan_guard = True
    target_x = 200.0
    height_mode, check_invalid, zero_reward_on_invalid = 1, True, True      # reference envs/humanoid.py:359-369
    MOTOR_STRENGTHS = [200, 200, 200, 200, 200, 600, 400, 100, 100, 200, 200, 600, 400, 100, 100, 100, 100, 200, 100, 100, 200]

    def __init__(self, render=False, device="cuda:0", num_envs=4096, seed=0, episode_length=1000, no_grad=True,
                 stochastic_init=False, MM_caching_frequency=1):
        super().__init__(num_envs, 76, 21, episode_length, MM_caching_frequency, seed, no_grad, render, device)
        self.stochastic_init = stochastic_init
        self.dt, self.sim_substeps = 1.0 / 60.0, 48
        self.sim_dt, self.ground = self.dt, True
        self.num_joint_q, self.num_joint_qd = 28, 27
        self._init_common(1.35, df.quat_from_axis_angle((1.0, 0.0, 0.0), -math.pi * 0.5))
        arrays = self._build_model("HumanoidEnv")
        self.start_joint_q = tu.to_torch(arrays["joint_q"][7:28], device=device)
        self.start_joint_target = self.start_joint_q.clone()
        self.termination_height, self.termination_tolerance = 0.74, 0.1
        self.motor_strengths = tu.to_torch(self.MOTOR_STRENGTHS, device=device).repeat((num_envs, 1))
        self.motor_scale, self.action_penalty = 0.35, -0.002
        self.joint_vel_obs_scaling, self.height_rew_scale = 0.1, 10.0

    def _apply_actions(self, actions):
        self.state.joint_act.view(self.num_envs, -1)[:, 6:] = actions * self.motor_scale * self.motor_strengths

    def _action_map(self):
        return self.num_joint_qd, 6, 1.0, 0.0, float(self.motor_scale), self.motor_strengths[0].contiguous(), False

    def calculateObservations(self):
        if self.fused_epilogue and torch.device(self.device).type == "cuda":
            self.obs_buf = self._fused(False)
            return
        q, qd, pos, rot, lin_vel, ang_vel, up, heading = self._torso_features()
        self.obs_buf = torch.cat([pos[:, 1:2], rot, lin_vel, ang_vel, q[:, 7:], self.joint_vel_obs_scaling * qd[:, 6:],
                                  up, heading, self.actions.clone()], dim=-1)

    def _height_reward(self, h):
        r = torch.clip(h - (self.termination_height + self.termination_tolerance), -1.0, self.termination_tolerance)
        r = torch.where(r < 0.0, -200.0 * r * r, r)
        return torch.where(r > 0.0, self.height_rew_scale * r, r)

    def calculateReward(self):
        o = self.obs_buf
        self.rew_buf = o[:, 5] + 0.1 * o[:, 53] + o[:, 54] + self._height_reward(o[:, 0]) + \
            torch.sum(self.actions ** 2, dim=-1) * self.action_penalty
        self.reset_buf = torch.where(o[:, 0] < self.termination_height, torch.ones_like(self.reset_buf), self.reset_buf)
        self.reset_buf = torch.where(self.progress_buf > self.episode_length - 1, torch.ones_like(self.reset_buf), self.reset_buf)
        invalid = self._invalid_state_mask()
        self.reset_buf = torch.where(invalid, torch.ones_like(self.reset_buf), self.reset_buf)
        self.rew_buf[invalid] = 0.0          # reference envs/humanoid.py:369: a NaN / Inf state must not poison the loss


class SNUHumanoidEnv(_FreeRootWalker):
    """reference envs/snu_humanoid.py (lower-body SNU skeleton, 152 muscle-tendon units): 53 obs, 152
    actions in [0, 1] after the affine map, 48 substeps, termination height 0.46."""

    nan_guard = True
    randomize_joints = False
    obs_has_actions, height_mode, action_penalty_abs = False, 2, True
    check_invalid, zero_reward_on_invalid = True, True

    def __init__(self, render=False, device="cuda:0", num_envs=4096, seed=0, episode_length=1000, no_grad=True,
                 stochastic_init=False, MM_caching_frequency=1):
        self.mtu_actuations = True
        self.num_joint_q, self.num_joint_qd = 29, 24
        self.num_dof, self.num_muscles, self.str_scale = 22, 152, 0.6
        super().__init__(num_envs, 53, self.num_muscles, episode_length, MM_caching_frequency, seed, no_grad, render, device)
        self.stochastic_init = stochastic_init
        self.inv_control_freq = 1
        self.dt, self.sim_substeps = 1.0 / 60.0, 48
        self.sim_dt, self.ground = self.dt, True
        self._init_common(1.0, df.quat_from_axis_angle((0.0, 1.0, 0.0), math.pi * 0.5))
        arrays = self._build_model("SNUHumanoidEnv")
        self.start_joint_q = tu.to_torch(arrays["joint_q"][7:29], device=device)
        self.start_joint_target = self.start_joint_q.clone()
        # str_scale is applied twice in the reference (snu_humanoid.py:176-180)
        f0 = arrays["muscle_params"][:, 0].astype(np.float64)
        self.muscle_strengths = tu.to_torch(self.str_scale * (self.str_scale * f0), device=device).repeat(num_envs)
        self.termination_height, self.termination_tolerance = 0.46, 0.05
        self.height_rew_scale, self.action_strength = 4.0, 100.0
        self.action_penalty, self.joint_vel_obs_scaling = -0.001, 0.1

    def _preprocess_actions(self, actions):
        return actions * 0.5 + 0.5

    def _apply_actions(self, actions):
        self.model.muscle_activation = actions.view(-1) * self.muscle_strengths

    def _action_map(self):
        return self.num_muscles, 0, 0.5, 0.5, 1.0, self.muscle_strengths[:self.num_muscles].contiguous(), True

    def calculateObservations(self):
        if self.fused_epilogue and torch.device(self.device).type == "cuda":
            self.obs_buf = self._fused(False)
            return
        q, qd, pos, rot, lin_vel, ang_vel, up, heading = self._torso_features()
        self.obs_buf = torch.cat([pos[:, 1:2], rot, lin_vel, ang_vel, q[:, 7:], self.joint_vel_obs_scaling * qd[:, 6:],
                                  up, heading], dim=-1)

    def calculateReward(self):
        o = self.obs_buf
        act_penalty = torch.sum(torch.abs(self.actions), dim=-1) * self.action_penalty
        self.rew_buf = o[:, 5] + 0.1 * o[:, 51] + o[:, 52] + act_penalty
        self.reset_buf = torch.where(o[:, 0] < self.termination_height, torch.ones_like(self.reset_buf), self.reset_buf)
        self.reset_buf = torch.where(self.progress_buf > self.episode_length - 1, torch.ones_like(self.reset_buf), self.reset_buf)
        invalid = self._invalid_state_mask()
        self.reset_buf = torch.where(invalid, torch.ones_like(self.reset_buf), self.reset_buf)
        self.rew_buf[invalid] = 0.0


class CartPoleSwingUpEnv(DFlexEnv):
    """reference envs/cartpole_swing_up.py: 5 obs, 1 action, 4 substeps, no ground."""

    clone_actions = False

    def __init__(self, render=False, device="cuda:0", num_envs=1024, seed=0, episode_length=240, no_grad=True,
                 stochastic_init=False, MM_caching_frequency=1, early_termination=False):
        super().__init__(num_envs, 5, 1, episode_length, MM_caching_frequency, seed, no_grad, render, device)
        self.stochastic_init, self.early_termination = stochastic_init, early_termination
        self.dt, self.sim_substeps = 1.0 / 60.0, 4
        self.sim_dt, self.ground = self.dt, False
        self.num_joint_q = self.num_joint_qd = 2
        self._build_model("CartPoleSwingUpEnv")
        # constants: detached (the reference clones the grad-requiring State tensors, cartpole_swing_up.py:74-75, which ties
        # every reset to the autograd leaf of the construction-time state -- on the default stream, so that a rollout
        # containing it cannot be captured into a CUDA graph)
        self.start_joint_q = self.state.joint_q.detach().clone()
        self.start_joint_qd = self.state.joint_qd.detach().clone()
        self.action_strength = 1000.0
        self.pole_angle_penalty, self.pole_velocity_penalty = 1.0, 0.1
        self.cart_position_penalty, self.cart_velocity_penalty, self.cart_action_penalty = 0.05, 0.1, 0.0

    def _apply_actions(self, actions):
        self.state.joint_act.view(self.num_envs, -1)[:, 0:1] = actions * self.action_strength

    def _reset_state(self, env_ids):
        q = self.state.joint_q.view(self.num_envs, -1)
        qd = self.state.joint_qd.view(self.num_envs, -1)
        q[env_ids, :] = self.start_joint_q.view(-1, self.num_joint_q)[env_ids, :].clone()
        qd[env_ids, :] = self.start_joint_qd.view(-1, self.num_joint_qd)[env_ids, :].clone()
        if self.stochastic_init:
            k, dev = len(env_ids), self.device
            q[env_ids, :] = q[env_ids, :] + np.pi * (torch.rand(size=(k, self.num_joint_q), device=dev) - 0.5)
            qd[env_ids, :] = qd[env_ids, :] + 0.5 * (torch.rand(size=(k, self.num_joint_qd), device=dev) - 0.5)

    def _start_state(self):
        n, dev = self.num_envs, self.device
        start_q = self.start_joint_q.view(n, self.num_joint_q)
        start_qd = self.start_joint_qd.view(n, self.num_joint_qd)
        if self.stochastic_init:
            start_q = start_q + np.pi * (torch.rand(size=(n, self.num_joint_q), device=dev) - 0.5)
            start_qd = start_qd + 0.5 * (torch.rand(size=(n, self.num_joint_qd), device=dev) - 0.5)
        return start_q, start_qd

    # ---- fused step (DFlexEnv._fused_step) ----
    fused_transition = True

    def _action_map(self):
        strength = torch.full((self.num_actions,), float(self.action_strength), device=self.device)
        return self.num_joint_qd, 0, 1.0, 0.0, 1.0, strength, False

    def _transition_params(self):
        from ..env_ops import DfxPlanarParams
        p = DfxPlanarParams()
        p.num_q, p.num_qd, p.num_act, p.num_obs = 2, 2, 1, 5
        p.kind, p.early_termination, p.zero_actions_on_reset = 2, 0, int(self.clone_actions)
        p.episode_length = int(self.episode_length)
        p.action_penalty = float(self.cart_action_penalty)
        p.pole_angle_penalty, p.pole_velocity_penalty = float(self.pole_angle_penalty), float(self.pole_velocity_penalty)
        p.cart_position_penalty, p.cart_velocity_penalty = float(self.cart_position_penalty), float(self.cart_velocity_penalty)
        return p

    def step(self, actions):
        if not (self.fused_transition and self.sync_free_reset and torch.device(self.device).type == "cuda"):
            return super().step(actions)
        return self._fused_step(actions)

    def clear_grad(self, checkpoint=None):
        with torch.no_grad():
            q, qd, act = self.state.joint_q.clone(), self.state.joint_qd.clone(), self.state.joint_act.clone()
            self.state = self.model.state()
            self.state.joint_q, self.state.joint_qd, self.state.joint_act = q, qd, act

    def calculateObservations(self):
        q = self.state.joint_q.view(self.num_envs, -1)
        qd = self.state.joint_qd.view(self.num_envs, -1)
        theta = q[:, 1:2]
        self.obs_buf = torch.cat([q[:, 0:1], qd[:, 0:1], torch.sin(theta), torch.cos(theta), qd[:, 1:2]], dim=-1)

    def calculateReward(self):
        q = self.state.joint_q.view(self.num_envs, -1)
        qd = self.state.joint_qd.view(self.num_envs, -1)
        theta = tu.normalize_angle(q[:, 1])
        self.rew_buf = -torch.pow(theta, 2.0) * self.pole_angle_penalty - torch.pow(qd[:, 1], 2.0) * self.pole_velocity_penalty \
            - torch.pow(q[:, 0], 2.0) * self.cart_position_penalty - torch.pow(qd[:, 0], 2.0) * self.cart_velocity_penalty \
            - torch.sum(self.actions ** 2, dim=-1) * self.cart_action_penalty
        self.reset_buf = torch.where(self.progress_buf > self.episode_length - 1, torch.ones_like(self.reset_buf), self.reset_buf)


class _PlanarHopper(DFlexEnv):
    """Planar (x, z, pitch) rooted chains: Hopper and HalfCheetah share reset and observation code."""

    init_noise = (0.05, 0.1, 0.05, 0.05 * 2.0)   # pos, pitch span, joints, velocity span

    def _init_planar(self, asset, start_height, n_joint):
        n, dev = self.num_envs, self.device
        self.dt, self.sim_substeps = 1.0 / 60.0, 16
        self.sim_dt, self.ground = self.dt, True
        self.num_joint_q = self.num_joint_qd = 3 + n_joint
        self.start_rotation = torch.tensor([0.0], device=dev)
        self.start_pos = tu.to_torch([[0.0, start_height]] * n, device=dev)
        self.start_joint_q = tu.to_torch([0.0] * n_joint, device=dev)
        self.start_joint_target = self.start_joint_q.clone()
        self._build_model(asset)

    def _apply_actions(self, actions):
        self.state.joint_act.view(self.num_envs, -1)[:, 3:] = actions * self.action_strength

    def _reset_state(self, env_ids):
        q = self.state.joint_q.view(self.num_envs, -1)
        qd = self.state.joint_qd.view(self.num_envs, -1)
        q[env_ids, 0:2] = self.start_pos[env_ids, :].clone()
        q[env_ids, 2] = self.start_rotation.clone()
        q[env_ids, 3:] = self.start_joint_q.clone()
        qd[env_ids, :] = 0.0
        if self.stochastic_init:
            k, dev = len(env_ids), self.device
            a_pos, a_rot, a_joint, a_vel = self.init_noise
            q[env_ids, 0:2] = q[env_ids, 0:2] + a_pos * (torch.rand(size=(k, 2), device=dev) - 0.5) * 2.0
            q[env_ids, 2] = (torch.rand(k, device=dev) - 0.5) * a_rot
            q[env_ids, 3:] = q[env_ids, 3:] + a_joint * (torch.rand(size=(k, self.num_joint_q - 3), device=dev) - 0.5) * 2.0
            qd[env_ids, :] = a_vel * (torch.rand(size=(k, self.num_joint_qd), device=dev) - 0.5)

    def _start_state(self):
        n, dev = self.num_envs, self.device
        if getattr(self, "_start_q_full", None) is None:
            self._start_q_full = torch.cat([self.start_pos, self.start_rotation.expand(n, 1),
                                            self.start_joint_q.expand(n, -1)], dim=-1).contiguous()
            self._zero_qd = torch.zeros((n, self.num_joint_qd), device=dev)
        start_q, start_qd = self._start_q_full, self._zero_qd
        if self.stochastic_init:
            a_pos, a_rot, a_joint, a_vel = self.init_noise
            start_q = start_q.clone()
            start_q[:, 0:2] = start_q[:, 0:2] + a_pos * (torch.rand(size=(n, 2), device=dev) - 0.5) * 2.0
            start_q[:, 2] = (torch.rand(n, device=dev) - 0.5) * a_rot
            start_q[:, 3:] = start_q[:, 3:] + a_joint * (torch.rand(size=(n, self.num_joint_q - 3), device=dev) - 0.5) * 2.0
            start_qd = a_vel * (torch.rand(size=(n, self.num_joint_qd), device=dev) - 0.5)
        return start_q, start_qd

    # ---- fused step (DFlexEnv._fused_step) ----
    fused_transition = True
    planar_kind = 0

    def _action_map(self):
        strength = torch.full((self.num_actions,), float(self.action_strength), device=self.device)
        return self.num_joint_qd, 3, 1.0, 0.0, 1.0, strength, False

    def _transition_params(self):
        from ..env_ops import DfxPlanarParams
        p = DfxPlanarParams()
        p.num_q, p.num_qd, p.num_act, p.num_obs = self.num_joint_q, self.num_joint_qd, self.num_actions, self.num_observations
        p.kind, p.early_termination = int(self.planar_kind), int(self.early_termination)
        p.zero_actions_on_reset, p.episode_length = int(self.clone_actions), int(self.episode_length)
        p.termination_height = float(getattr(self, "termination_height", 0.0))
        p.termination_height_tolerance = float(getattr(self, "termination_height_tolerance", 0.0))
        p.termination_angle = float(getattr(self, "termination_angle", 1.0))
        p.height_rew_scale = float(getattr(self, "height_rew_scale", 1.0))
        p.action_penalty = float(self.action_penalty)
        return p

    def step(self, actions):
        if not (self.fused_transition and self.sync_free_reset and torch.device(self.device).type == "cuda"):
            return super().step(actions)
        return self._fused_step(actions)

    def calculateObservations(self):
        self.obs_buf = torch.cat([self.state.joint_q.view(self.num_envs, -1)[:, 1:], self.state.joint_qd.view(self.num_envs, -1)], dim=-1)


class HopperEnv(_PlanarHopper):
    """reference envs/hopper.py: 11 obs, 3 actions."""

    def __init__(self, render=False, device="cuda:0", num_envs=4096, seed=0, episode_length=1000, no_grad=True,
                 stochastic_init=False, MM_caching_frequency=1, early_termination=True):
        super().__init__(num_envs, 11, 3, episode_length, MM_caching_frequency, seed, no_grad, render, device)
        self.stochastic_init, self.early_termination = stochastic_init, early_termination
        self._init_planar("HopperEnv", 0.0, 3)
        self.termination_height, self.termination_angle = -0.45, np.pi / 6.0
        self.termination_height_tolerance, self.termination_angle_tolerance = 0.15, 0.05
        self.height_rew_scale, self.action_strength, self.action_penalty = 1.0, 200.0, -1e-1

    def calculateReward(self):
        o = self.obs_buf
        h = torch.clip(o[:, 0] - (self.termination_height + self.termination_height_tolerance), -1.0, 0.3)
        h = torch.where(h < 0.0, -200.0 * h * h, h)
        h = torch.where(h > 0.0, self.height_rew_scale * h, h)
        angle_reward = 1.0 * (-o[:, 1] ** 2 / (self.termination_angle ** 2) + 1.0)
        self.rew_buf = o[:, 5] + h + angle_reward + torch.sum(self.actions ** 2, dim=-1) * self.action_penalty
        self.reset_buf = torch.where(self.progress_buf > self.episode_length - 1, torch.ones_like(self.reset_buf), self.reset_buf)
        if self.early_termination:
            self.reset_buf = torch.where(o[:, 0] < self.termination_height, torch.ones_like(self.reset_buf), self.reset_buf)


class CheetahEnv(_PlanarHopper):
    """reference envs/cheetah.py: 17 obs, 6 actions, no early termination."""

    init_noise = (0.1, 0.2, 0.1, 0.5)
    planar_kind = 1

    def __init__(self, render=False, device="cuda:0", num_envs=4096, seed=0, episode_length=1000, no_grad=True,
                 stochastic_init=False, MM_caching_frequency=1, early_termination=False):
        super().__init__(num_envs, 17, 6, episode_length, MM_caching_frequency, seed, no_grad, render, device)
        self.stochastic_init, self.early_termination = stochastic_init, early_termination
        self._init_planar("CheetahEnv", -0.2, 6)
        self.action_strength, self.action_penalty = 200.0, -0.1

    def calculateReward(self):
        self.rew_buf = self.obs_buf[:, 8] + torch.sum(self.actions ** 2, dim=-1) * self.action_penalty
        self.reset_buf = torch.where(self.progress_buf > self.episode_length - 1, torch.ones_like(self.reset_buf), self.reset_buf)
