"""AMP policy-transition ring buffer on the fused gather / ring-insert kernels.

Reference: rsl_rl/rsl_rl/storage/replay_buffer.py:35-74 (`ReplayBuffer`, used by
hybrid_ppo.py:75-76 / amp_ppo.py:74-75 as `amp_storage`).  Same constructor, attributes
(`states`, `next_states`, `step`, `num_samples`) and methods; `insert` is one launch for both
tensors (wrap-around included), `feed_forward_generator` draws its indices from the same host
`np.random.choice` stream and gathers both tensors with one `hl_minibatch_gather` launch.
"""
import numpy as np
import torch

from . import _lib as L


class ReplayBuffer:
    def __init__(self, obs_dim, buffer_size, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ReplayBuffer (B200) runs on CUDA only (no CPU fallback)")
        self.states = torch.zeros(buffer_size, obs_dim, device=self.device)
        self.next_states = torch.zeros(buffer_size, obs_dim, device=self.device)
        self.buffer_size = buffer_size
        self.step = 0
        self.num_samples = 0

    def insert(self, states, next_states):
        """replay_buffer.py:52-68: rows land at (step + r) mod buffer_size; the reference's two-slice
        wrap accepts at most 2*buffer_size - step rows (later rows overwrite earlier ones)."""
        num = int(states.shape[0])
        size, step = self.buffer_size, self.step
        end = step + num
        if end - size > size:
            raise RuntimeError(f"ReplayBuffer.insert: {num} rows do not fit the wrap-around of a {size}-row buffer at step {step}")
        s = states.detach().to(self.device, torch.float32).contiguous()
        sn = next_states.detach().to(self.device, torch.float32).contiguous()
        if num > 0:
            L.check(L.lib.hl_ring_insert(L.ptr(s), L.ptr(sn), L.ptr(self.states), L.ptr(self.next_states), num,
                                         int(self.states.shape[1]), size, step, L.stream()))
        self.num_samples = min(size, max(end, self.num_samples))
        self.step = (step + num) % size

    def gather(self, sample_idxs):
        idx = torch.as_tensor(np.asarray(sample_idxs), dtype=torch.int64).to(self.device)
        m, w = idx.numel(), int(self.states.shape[1])
        out_s = torch.empty(m, w, device=self.device)
        out_n = torch.empty(m, w, device=self.device)
        if m:
            g = L.HlGatherFields()
            g.struct_bytes = L.ctypes.sizeof(L.HlGatherFields)
            g.n_fields = 2
            g.src[0], g.dst[0], g.width[0] = L.ptr(self.states), L.ptr(out_s), w
            g.src[1], g.dst[1], g.width[1] = L.ptr(self.next_states), L.ptr(out_n), w
            L.check(L.lib.hl_minibatch_gather(L.ctypes.byref(g), L.ptr(idx), m, self.buffer_size, L.stream()))
        return out_s, out_n

    def feed_forward_generator(self, num_mini_batch, mini_batch_size):
        """replay_buffer.py:70-74 (host RNG draw kept: it is the reference's sampling stream)."""
        for _ in range(num_mini_batch):
            sample_idxs = np.random.choice(self.num_samples, size=mini_batch_size)
            yield self.gather(sample_idxs)
