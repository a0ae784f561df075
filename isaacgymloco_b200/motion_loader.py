"""AMPLoader: mocap clip tables on the GPU + the batch frame interpolation / expert-pair
kernels.  Reference: rsl_rl/rsl_rl/datasets/motion_loader.py (ML) -- same class name,
constructor signature, layout constants and method names.

Load time (JSON parse, leg reorder ML:134-164, per-frame quaternion normalise + standardise
ML:85-93) stays host-side numpy: it runs once.  Sampling of clip ids / times stays host numpy RNG
exactly as in the reference (ML:171-187), so a seeded `np.random` yields the reference's stream.
The per-sample work (`get_full_frame_at_time_batch` ML:231-255, `feed_forward_generator`
ML:315-343 preload branch) runs in libhimloco_b200.
"""
import glob
import json

import numpy as np
import torch

from . import _lib as L


class AMPLoader:
    POS_SIZE = 3
    ROT_SIZE = 4
    JOINT_POS_SIZE = 12
    TAR_TOE_POS_LOCAL_SIZE = 12
    LINEAR_VEL_SIZE = 3
    ANGULAR_VEL_SIZE = 3
    JOINT_VEL_SIZE = 12
    TAR_TOE_VEL_LOCAL_SIZE = 12

    ROOT_POS_START_IDX, ROOT_POS_END_IDX = 0, 3
    ROOT_ROT_START_IDX, ROOT_ROT_END_IDX = 3, 7
    JOINT_POSE_START_IDX, JOINT_POSE_END_IDX = 7, 19
    TAR_TOE_POS_LOCAL_START_IDX, TAR_TOE_POS_LOCAL_END_IDX = 19, 31
    LINEAR_VEL_START_IDX, LINEAR_VEL_END_IDX = 31, 34
    ANGULAR_VEL_START_IDX, ANGULAR_VEL_END_IDX = 34, 37
    JOINT_VEL_START_IDX, JOINT_VEL_END_IDX = 37, 49
    TAR_TOE_VEL_LOCAL_START_IDX, TAR_TOE_VEL_LOCAL_END_IDX = 49, 61

    def __init__(self, device, time_between_frames, data_dir="", preload_transitions=False,
                 num_preload_transitions=1000000, motion_files=None, clip_tables=None):
        """`clip_tables` (optional) = dict(frames=[(n_i,49) arrays], frame_durations=[...],
        weights=[...], names=[...]) loads pre-parsed clips instead of `motion_files`."""
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("AMPLoader (B200) runs on CUDA only (no CPU fallback)")
        self.time_between_frames = time_between_frames
        self.trajectories, self.trajectories_full = [], []
        self.trajectory_names, self.trajectory_idxs = [], []
        lens, weights, durations, nframes = [], [], [], []
        if clip_tables is None:
            if motion_files is None:
                motion_files = glob.glob("datasets/motion_files2/*")
            clip_tables = dict(frames=[], frame_durations=[], weights=[], names=[])
            for f in motion_files:
                data, dur, w = self._parse_motion_file(f)
                clip_tables["frames"].append(data)
                clip_tables["frame_durations"].append(dur)
                clip_tables["weights"].append(w)
                clip_tables["names"].append(f.split(".")[0])
        for i, frames in enumerate(clip_tables["frames"]):
            full = torch.as_tensor(np.asarray(frames)[:, :self.JOINT_VEL_END_IDX], dtype=torch.float32).to(self.device)
            self.trajectories_full.append(full)
            self.trajectories.append(full[:, self.ROOT_ROT_END_IDX:self.JOINT_VEL_END_IDX])
            self.trajectory_names.append(clip_tables.get("names", [str(k) for k in range(len(clip_tables["frames"]))])[i])
            self.trajectory_idxs.append(i)
            dur = float(clip_tables["frame_durations"][i])
            durations.append(dur)
            weights.append(float(clip_tables["weights"][i]))
            lens.append((full.shape[0] - 1) * dur)             # ML:109 (count-1) * dt
            nframes.append(float(full.shape[0]))               # ML:111 the frame COUNT
        self._raw_weights = np.array(weights, dtype=np.float64)
        self.trajectory_weights = np.array(weights) / np.sum(weights)
        self.trajectory_frame_durations = np.array(durations)
        self.trajectory_lens = np.array(lens)
        self.trajectory_num_frames = np.array(nframes)
        # one stacked table + per-clip row offsets for the kernels
        self.all_trajectories_full = torch.vstack(self.trajectories_full).contiguous()
        offs = np.concatenate([[0], np.cumsum([t.shape[0] for t in self.trajectories_full])[:-1]]).astype(np.int32)
        self._clip_offset = torch.from_numpy(offs).to(self.device)
        self._clip_len = torch.from_numpy(self.trajectory_lens.astype(np.float64)).to(self.device)
        self._clip_nf = torch.from_numpy(self.trajectory_num_frames.astype(np.float64)).to(self.device)
        self.preload_transitions = preload_transitions
        if self.preload_transitions:
            traj_idxs = self.weighted_traj_idx_sample_batch(num_preload_transitions)
            times = self.traj_time_sample_batch(traj_idxs)
            self.preloaded_s = self.get_full_frame_at_time_batch(traj_idxs, times)
            self.preloaded_s_next = self.get_full_frame_at_time_batch(traj_idxs, times + self.time_between_frames)

    # ------------------------------------------------------------------ binary clip cache
    # The reference re-parses the JSON clips and normalises every quaternion in a Python loop at each
    # start-up (ML:76-111).  The cache holds the already reordered / normalised / standardised float64
    # frame tables plus FrameDuration, MotionWeight and names: one .npz, loaded without parsing.
    CACHE_VERSION = 1

    @classmethod
    def build_cache(cls, motion_files, path):
        """Host-only: parse `motion_files` like the loader does and write the binary cache."""
        tabs = dict(frames=[], frame_durations=[], weights=[], names=[])
        for f in motion_files:
            data, dur, w = cls._parse_motion_file(f)
            tabs["frames"].append(data)
            tabs["frame_durations"].append(dur)
            tabs["weights"].append(w)
            tabs["names"].append(f.split(".")[0])
        cls._write_cache(path, tabs)
        return tabs

    @classmethod
    def _write_cache(cls, path, tabs):
        out = {"version": np.int64(cls.CACHE_VERSION), "n_clips": np.int64(len(tabs["frames"])),
               "frame_durations": np.asarray(tabs["frame_durations"], dtype=np.float64),
               "weights": np.asarray(tabs["weights"], dtype=np.float64), "names": np.asarray([str(x) for x in tabs["names"]])}
        for i, fr in enumerate(tabs["frames"]):
            out[f"clip{i}"] = np.asarray(fr)
        with open(path, "wb") as fh:
            np.savez(fh, **out)

    @classmethod
    def read_cache(cls, path):
        with np.load(path, allow_pickle=False) as z:
            if int(z["version"]) != cls.CACHE_VERSION:
                raise ValueError(f"mocap cache version {int(z['version'])} != {cls.CACHE_VERSION}")
            k = int(z["n_clips"])
            return dict(frames=[z[f"clip{i}"] for i in range(k)], frame_durations=list(z["frame_durations"]),
                        weights=list(z["weights"]), names=[str(x) for x in z["names"]])

    def save_cache(self, path):
        """Write this loader's clip tables (fp32 on the device -> exact round trip)."""
        self._write_cache(path, dict(frames=[t.cpu().numpy() for t in self.trajectories_full],
                                     frame_durations=self.trajectory_frame_durations, weights=self._raw_weights,
                                     names=self.trajectory_names))

    @classmethod
    def from_cache(cls, path, device, time_between_frames, **kw):
        return cls(device, time_between_frames, clip_tables=cls.read_cache(path), **kw)

    # ------------------------------------------------------------------ load time (host)
    @classmethod
    def _parse_motion_file(cls, path):
        with open(path, "r") as f:
            mj = json.load(f)
        data = cls.reorder_from_pybullet_to_isaac(np.array(mj["Frames"], dtype=np.float64))
        q = data[:, 3:7]
        q = q / np.linalg.norm(q, axis=1, keepdims=True)        # pose3d.QuaternionNormalize
        q = np.where(q[:, 3:4] < 0, -q, q)                      # motion_util.standardize_quaternion
        data[:, 3:7] = q
        return data, float(mj["FrameDuration"]), float(mj["MotionWeight"])

    @staticmethod
    def reorder_from_pybullet_to_isaac(motion_data):
        """ML:134-164: legs [FR, FL, RR, RL] -> [FL, FR, RL, RR] in every 12-wide per-leg block."""
        out = motion_data.copy()
        for start in (7, 19, 37, 49):
            blk = motion_data[:, start:start + 12].reshape(-1, 4, 3)
            out[:, start:start + 12] = blk[:, [1, 0, 3, 2], :].reshape(-1, 12)
        return out

    # ------------------------------------------------------------------ sampling (host numpy RNG, ML:166-187)
    def weighted_traj_idx_sample(self):
        return np.random.choice(self.trajectory_idxs, p=self.trajectory_weights)

    def weighted_traj_idx_sample_batch(self, size):
        return np.random.choice(self.trajectory_idxs, size=size, p=self.trajectory_weights, replace=True)

    def traj_time_sample(self, traj_idx):
        subst = self.time_between_frames + self.trajectory_frame_durations[traj_idx]
        return max(0, (self.trajectory_lens[traj_idx] * np.random.uniform() - subst))

    def traj_time_sample_batch(self, traj_idxs):
        subst = self.time_between_frames + self.trajectory_frame_durations[traj_idxs]
        t = self.trajectory_lens[traj_idxs] * np.random.uniform(size=len(traj_idxs)) - subst
        return np.maximum(np.zeros_like(t), t)

    def slerp(self, val0, val1, blend):
        return (1.0 - blend) * val0 + blend * val1

    def get_trajectory(self, traj_idx):
        return self.trajectories_full[traj_idx]

    # ------------------------------------------------------------------ kernels
    def get_full_frame_at_time_batch(self, traj_idxs, times, return_indices=False):
        """ML:231-255 -> (B,49) fp32 on the device.  traj_idxs: int array, times: float64 array."""
        ti = np.asarray(traj_idxs, dtype=np.int64)
        tm = np.asarray(times, dtype=np.float64)
        if ti.size:
            # same failure as the reference's frame lookup (ML:243-244): an index past the clip
            if ti.min() < 0 or ti.max() >= len(self.trajectories_full):
                raise IndexError("trajectory index out of range")
            pn = tm / self.trajectory_lens[ti] * self.trajectory_num_frames[ti]
            if np.floor(pn).min() < -self.trajectory_num_frames[ti].min() or (np.ceil(pn) > self.trajectory_num_frames[ti] - 1).any():
                raise IndexError("frame index out of bounds for the clip (time beyond its last frame)")
        idx_t = torch.as_tensor(ti).to(self.device, non_blocking=True)
        times_t = torch.as_tensor(tm).to(self.device, non_blocking=True)
        return self.get_full_frame_at_time_batch_device(idx_t, times_t, return_indices)

    def get_full_frame_at_time_batch_device(self, idx_t, times_t, return_indices=False):
        b = idx_t.numel()
        out = torch.empty(b, 49, device=self.device)
        lo = hi = None
        if return_indices:
            lo = torch.empty(b, dtype=torch.int32, device=self.device)
            hi = torch.empty(b, dtype=torch.int32, device=self.device)
        L.check(L.lib.hl_amp_frame_blend(L.ptr(self.all_trajectories_full), L.ptr(self._clip_offset), L.ptr(self._clip_len),
                                         L.ptr(self._clip_nf), len(self.trajectories_full), L.ptr(idx_t), L.ptr(times_t),
                                         L.ptr(out), L.ptr(lo), L.ptr(hi), b, L.stream()))
        return (out, lo, hi) if return_indices else out

    # ------------------------------------------------------------------ scalar API (ML:196-204,221-229,257-267,270-313)
    # One frame at a time: kept for API parity (the hot path is the batch form above).  A 1-sample launch of the
    # batch kernel gives the lerp columns; get_full_frame_at_time's rotation is the reference's scalar route,
    # `pybullet_utils.transformations.quaternion_slerp` + standardize_quaternion (ML:300-303) -- pybullet_utils is a
    # third-party module absent from the reference tree: its published algorithm (C. Gohlke's transformations.py) is
    # restated below on the host => parity unpinned at that one call.
    def get_frame_at_time(self, traj_idx, time):
        """ML:196-204 -> (42,): lerp of the AMP columns 7:49 of the two bracketing frames."""
        return self.get_full_frame_at_time_batch(np.array([traj_idx]), np.array([float(time)]))[0, self.ROOT_ROT_END_IDX:]

    @staticmethod
    def _quaternion_slerp_host(q0, q1, fraction):
        eps = np.finfo(float).eps * 4.0
        q0 = np.array(q0[:4], dtype=np.float64)
        q1 = np.array(q1[:4], dtype=np.float64)
        q0 /= np.linalg.norm(q0)
        q1 /= np.linalg.norm(q1)
        if fraction == 0.0:
            return q0
        if fraction == 1.0:
            return q1
        d = float(np.dot(q0, q1))
        if abs(abs(d) - 1.0) < eps:
            return q0
        if d < 0.0:
            d, q1 = -d, -q1
        angle = np.arccos(d)
        if abs(angle) < eps:
            return q0
        isin = 1.0 / np.sin(angle)
        return q0 * (np.sin((1.0 - fraction) * angle) * isin) + q1 * (np.sin(fraction * angle) * isin)

    def blend_frame_pose(self, frame0, frame1, blend):
        """ML:270-313 -> (49,)."""
        q = self._quaternion_slerp_host(frame0[3:7].cpu().numpy(), frame1[3:7].cpu().numpy(), blend)
        if q[3] < 0:                                     # motion_util.standardize_quaternion
            q = -q
        out = self.slerp(frame0, frame1, blend).clone()
        out[3:7] = torch.tensor(q, dtype=torch.float32, device=self.device)
        return out

    def get_full_frame_at_time(self, traj_idx, time):
        """ML:221-229 -> (49,)."""
        p = float(time) / self.trajectory_lens[traj_idx]
        n = self.trajectories_full[traj_idx].shape[0]
        lo, hi = int(np.floor(p * n)), int(np.ceil(p * n))
        return self.blend_frame_pose(self.trajectories_full[traj_idx][lo], self.trajectories_full[traj_idx][hi], p * n - lo)

    def get_frame(self):
        """ML:257-261: a random 42-wide AMP frame."""
        traj_idx = self.weighted_traj_idx_sample()
        return self.get_frame_at_time(traj_idx, self.traj_time_sample(traj_idx))

    def get_full_frame(self):
        """ML:263-267: a random full frame."""
        traj_idx = self.weighted_traj_idx_sample()
        return self.get_full_frame_at_time(traj_idx, self.traj_time_sample(traj_idx))

    def get_full_frame_batch(self, num_frames):
        if self.preload_transitions:
            idxs = np.random.choice(self.preloaded_s.shape[0], size=num_frames)
            return self.preloaded_s[torch.as_tensor(idxs, device=self.device)]
        traj_idxs = self.weighted_traj_idx_sample_batch(num_frames)
        return self.get_full_frame_at_time_batch(traj_idxs, self.traj_time_sample_batch(traj_idxs))

    def gather_pairs(self, idxs):
        """(s, s_next) for preloaded rows `idxs` (host array or device tensor)."""
        idx_t = idxs if isinstance(idxs, torch.Tensor) else torch.as_tensor(np.asarray(idxs, dtype=np.int64))
        idx_t = idx_t.to(self.device, torch.long, non_blocking=True).contiguous()
        b = idx_t.numel()
        s = torch.empty(b, 30, device=self.device)
        sn = torch.empty(b, 30, device=self.device)
        L.check(L.lib.hl_amp_gather_pairs(L.ptr(self.preloaded_s), L.ptr(self.preloaded_s_next), self.preloaded_s.shape[0],
                                          L.ptr(idx_t), L.ptr(s), L.ptr(sn), b, L.stream()))
        return s, sn

    def feed_forward_generator(self, num_mini_batch, mini_batch_size):
        """ML:315-343: yields (s, s_next), each (mini_batch_size, 30)."""
        for _ in range(num_mini_batch):
            if self.preload_transitions:
                idxs = np.random.choice(self.preloaded_s.shape[0], size=mini_batch_size)
                yield self.gather_pairs(idxs)
            else:
                # ML:331-343: a Python loop of get_frame_at_time + vstack there -- (mini_batch_size, 42) lerp frames
                # (the 42 AMP columns 7:49, not the 30 of the preload branch); here two batch launches
                traj_idxs = self.weighted_traj_idx_sample_batch(mini_batch_size)
                times = self.traj_time_sample_batch(traj_idxs)
                yield (self.get_full_frame_at_time_batch(traj_idxs, times)[:, self.ROOT_ROT_END_IDX:],
                       self.get_full_frame_at_time_batch(traj_idxs, times + self.time_between_frames)[:, self.ROOT_ROT_END_IDX:])

    @property
    def observation_dim(self):
        return self.trajectories[0].shape[1] - 12           # ML:346-349 -> 30

    @property
    def num_motions(self):
        return len(self.trajectory_names)

    # column accessors (ML:355-400)
    @staticmethod
    def get_root_pos_batch(p): return p[:, 0:3]
    @staticmethod
    def get_root_rot_batch(p): return p[:, 3:7]
    @staticmethod
    def get_joint_pose_batch(p): return p[:, 7:19]
    @staticmethod
    def get_tar_toe_pos_local_batch(p): return p[:, 19:31]
    @staticmethod
    def get_linear_vel_batch(p): return p[:, 31:34]
    @staticmethod
    def get_angular_vel_batch(p): return p[:, 34:37]
    @staticmethod
    def get_joint_vel_batch(p): return p[:, 37:49]
