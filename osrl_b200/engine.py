"""Python handle on one C-ABI engine (one process, one GPU).

Thin plumbing only: builds the ``osrl_config``, exposes the parameter arena as zero-copy
torch views (``__cuda_array_interface__``), forwards minibatches / noise by raw pointer on
torch's current CUDA stream.  All arithmetic happens in libosrl_b200.so.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, Iterable, Optional

import numpy as np
import torch

from . import _lib
from ._lib import ALGO, Batch, Config, DatasetView, Noise, ParamDesc, SeqBatch, SeqDatasetView, check

_BATCH_KEYS = ("observations", "next_observations", "actions", "rewards", "costs", "done")
_STEP_KEYS = _BATCH_KEYS + ("is_init",)   # COptiDICE batches carry a 7th tensor (coptidice.py:126-127)
_F_TYPES = {"chi2": 0, "softchi": 1, "kl": 2}


def make_config(algo: str, *, batch_size: int, seed: int = 0, world_size: int = 1, rank: int = 0, **kw) -> Config:
    """Fill an ``osrl_config`` from the reference's constructor/trainer keyword names."""
    cfg = Config()
    cfg.algo = ALGO[algo]
    cfg.batch_size, cfg.seed, cfg.world_size, cfg.rank = int(batch_size), int(seed), int(world_size), int(rank)
    ren = {"state_dim": "obs_dim", "action_dim": "act_dim", "vae_hidden_sizes": "vae_hidden"}
    # defaults of fields the reference defaults too
    cfg.max_action, cfg.gamma, cfg.tau, cfg.phi, cfg.lmbda, cfg.beta = 1.0, 0.99, 0.005, 0.05, 0.75, 0.5
    if algo == "cpq":
        cfg.beta = 1.5   # cpq.py:48 (BCQ-Lag / BEAR-Lag default to 0.5: bcql.py:56, bearl.py:56)
    cfg.pid_kp, cfg.pid_ki, cfg.pid_kd = 0.1, 0.003, 0.001
    cfg.num_q = cfg.num_qc = 1
    cfg.cost_limit, cfg.episode_len, cfg.sample_action_num = 10, 300, 10
    cfg.qc_scalar, cfg.mmd_sigma, cfg.target_mmd_thresh, cfg.num_samples_mmd_match = 1.5, 50.0, 0.05, 10
    cfg.start_update_policy_step = 20000
    cfg.actor_lr = cfg.critic_lr = cfg.vae_lr = 1e-4
    cfg.alpha_lr = 1e-4 if algo == "cpq" else 1e-3
    # CDT defaults (cdt.py:45-70, 291-310; the mode every reference task config uses)
    cfg.seq_len, cfg.embedding_dim, cfg.num_layers, cfg.num_heads = 10, 128, 4, 8
    cfg.use_rew = cfg.use_cost = cfg.cost_transform = cfg.stochastic = 1
    cfg.init_temperature, cfg.learning_rate, cfg.weight_decay = 0.1, 1e-4, 1e-4
    cfg.adam_beta1, cfg.adam_beta2, cfg.clip_grad, cfg.lr_warmup_steps = 0.9, 0.999, 0.25, 10000
    if algo == "cdt":
        cfg.episode_len = 1000
        cfg.target_entropy = -float(kw.get("action_dim", 0))
    # COptiDICE defaults (coptidice.py:68-85, 267-276)
    cfg.f_type, cfg.init_state_propotion, cfg.alpha, cfg.cost_ub_epsilon = 1, 1.0, 0.5, 0.01
    cfg.num_nu = cfg.num_chi = 1
    cfg.scalar_lr = 1e-3
    cfg._keep = []   # numpy arrays behind the host pointers of the struct
    for k, v in kw.items():
        k = ren.get(k, k)
        if k in ("a_hidden_sizes", "c_hidden_sizes"):
            v = list(v)
            if not 1 <= len(v) <= _lib.OSRL_MAX_HIDDEN:
                raise ValueError(f"{k}: 1..{_lib.OSRL_MAX_HIDDEN} hidden layers supported")
            p = k[0]
            setattr(cfg, f"n_{p}_hidden", len(v))
            arr = getattr(cfg, f"{p}_hidden")
            for i, x in enumerate(v):
                arr[i] = int(x)
        elif k == "PID":
            cfg.pid_kp, cfg.pid_ki, cfg.pid_kd = [float(x) for x in v]
        elif k == "kernel":
            cfg.mmd_kernel = {"gaussian": 0, "laplacian": 1}[v]
        elif k == "f_type":
            cfg.f_type = _F_TYPES[v] if isinstance(v, str) else int(v)
        elif k in ("observations_std", "actions_std"):
            arr = np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(-1))
            cfg._keep.append(arr)
            setattr(cfg, k, arr.ctypes.data)
        elif k == "betas":
            cfg.adam_beta1, cfg.adam_beta2 = float(v[0]), float(v[1])
        elif k in ("device", "episode_len_unused"):
            continue
        elif hasattr(cfg, k):
            setattr(cfg, k, v)
        else:
            raise TypeError(f"unknown config field {k!r}")
    return cfg


class _Cai:
    """Non-owning CUDA array view for torch.as_tensor."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2}


def _as_f32(x, device_index: int):
    """-> (tensor kept alive, pointer, on_host)"""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x, dtype=torch.float32)
    if x.dtype != torch.float32:
        x = x.float()
    x = x.contiguous()
    if x.is_cuda and x.device.index != device_index:
        raise ValueError("batch tensor lives on a different GPU than the engine")
    return x, x.data_ptr(), (0 if x.is_cuda else 1)


class Engine:
    def __init__(self, algo: str, *, batch_size: int, device: int = 0, seed: int = 0, world_size: int = 1,
                 rank: int = 0, **hyper):
        self.lib = _lib.load()
        self.algo = algo
        self.device = int(device)
        self.batch_size = int(batch_size)
        self.cfg = make_config(algo, batch_size=batch_size, seed=seed, world_size=world_size, rank=rank, **hyper)
        h = C.c_void_p()
        check(self.lib.osrl_engine_create(C.byref(self.cfg), self.device, C.byref(h)))
        self.h = h
        self._keep = []
        n = C.c_int()
        check(self.lib.osrl_param_table(self.h, None, 0, C.byref(n)))
        arr = (ParamDesc * n.value)()
        check(self.lib.osrl_param_table(self.h, arr, n.value, C.byref(n)))
        self.table = list(arr)
        names = (C.c_char_p * 16)()
        check(self.lib.osrl_stat_names(self.h, names, 16, C.byref(n)))
        self.stat_names = [names[i].decode() for i in range(n.value)]
        nn_ = (C.c_char_p * _lib.OSRL_MAX_NOISE)()
        cnt = (C.c_int64 * _lib.OSRL_MAX_NOISE)()
        check(self.lib.osrl_noise_layout(self.h, nn_, cnt, _lib.OSRL_MAX_NOISE, C.byref(n)))
        self.noise_layout = OrderedDict((nn_[i].decode(), int(cnt[i])) for i in range(n.value))

    # ------------------------------------------------------------------ life cycle
    def close(self):
        if getattr(self, "h", None):
            self.lib.osrl_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return int(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ parameters
    def param_views(self) -> "OrderedDict[str, torch.Tensor]":
        """state_dict-ordered zero-copy float32 views into the arena (trained + target sections)."""
        out = OrderedDict()
        for d in self.table:
            shape = (d.rows, d.cols) if d.cols else (d.rows,)
            t = torch.as_tensor(_Cai(d.ptr, shape), device=f"cuda:{self.device}")
            t._osrl_engine = self  # keep the arena alive as long as a view exists
            out[d.name.decode()] = t
        return out

    def load_params(self, params: Dict[str, torch.Tensor], sync_targets: bool = False) -> None:
        """Copy tensors (reference state_dict names) into the arena."""
        names = [d.name.decode() for d in self.table]
        for i, name in enumerate(names):
            if name not in params:
                continue
            t = params[name].detach().to(torch.float32).cpu().contiguous()
            check(self.lib.osrl_param_set(self.h, i, C.c_void_p(t.data_ptr()), t.numel()))
        if sync_targets:
            check(self.lib.osrl_sync_targets(self.h))

    def read_params(self) -> "OrderedDict[str, torch.Tensor]":
        out = OrderedDict()
        for i, d in enumerate(self.table):
            shape = (d.rows, d.cols) if d.cols else (d.rows,)
            t = torch.empty(shape, dtype=torch.float32)
            check(self.lib.osrl_param_get(self.h, i, C.c_void_p(t.data_ptr()), t.numel()))
            out[d.name.decode()] = t
        return out

    def debug_linear(self, impl: str, A, W, bias=None, act: int = 0) -> torch.Tensor:
        """act(A @ W.T + bias) through one GEMM kernel family ('ffma' | 'mma' | 'tc5'); unit-test hook."""
        A = A.detach().float().cpu().contiguous(); W = W.detach().float().cpu().contiguous()
        M, K = A.shape; N = W.shape[0]
        out = torch.empty(M, N)
        b = bias.detach().float().cpu().contiguous() if bias is not None else None
        check(self.lib.osrl_debug_linear(self.h, impl.encode(), M, N, K, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()),
                                         C.c_void_p(b.data_ptr()) if b is not None else None, int(act),
                                         C.c_void_p(out.data_ptr())))
        return out

    def debug_gemm(self, impl: str, A, B, a_kc: bool = True, b_kc: bool = True, colsum: bool = False):
        """op(A) @ op(B) in any operand layout of the step (forward 1,1; dgrad 1,0; wgrad 0,0); unit-test hook.
        A is [M,K] (a_kc) or [K,M]; B is [N,K] (b_kc) or [K,N].  Returns C [M,N] (and sum_k A(i,k) if colsum)."""
        A = A.detach().float().cpu().contiguous(); B = B.detach().float().cpu().contiguous()
        M, K = (A.shape if a_kc else A.shape[::-1])
        N = B.shape[0] if b_kc else B.shape[1]
        assert (B.shape[1] if b_kc else B.shape[0]) == K
        out = torch.empty(M, N)
        cs = torch.empty(M) if colsum else None
        check(self.lib.osrl_debug_gemm(self.h, impl.encode(), M, N, K, C.c_void_p(A.data_ptr()), int(a_kc),
                                       C.c_void_p(B.data_ptr()), int(b_kc), C.c_void_p(out.data_ptr()),
                                       C.c_void_p(cs.data_ptr()) if colsum else None))
        return (out, cs) if colsum else out

    def read_section(self, section: str) -> "OrderedDict[str, torch.Tensor]":
        """Trained-parameter-shaped tensors of another arena section: 'grad', 'adam_m', 'adam_v'."""
        sec = {"param": 0, "target": 1, "grad": 2, "adam_m": 3, "adam_v": 4}[section]
        out = OrderedDict()
        for d in self.table:
            if d.section != 0:
                continue
            shape = (d.rows, d.cols) if d.cols else (d.rows,)
            t = torch.empty(shape, dtype=torch.float32)
            check(self.lib.osrl_debug_read(self.h, sec, d.offset, t.numel(), C.c_void_p(t.data_ptr())))
            out[d.name.decode()] = t
        return out

    def sync_targets(self):
        check(self.lib.osrl_sync_targets(self.h))

    # ------------------------------------------------------------------ dataset
    def upload_dataset(self, data: dict, reward_scale: float = 1.0, cost_scale: float = 1.0) -> None:
        keep = {k: np.ascontiguousarray(data[k], dtype=np.float32)
                for k in ("observations", "next_observations", "actions", "rewards", "costs")}
        v = DatasetView()
        v.n = keep["observations"].shape[0]
        for k, a in keep.items():
            setattr(v, k, a.ctypes.data)
        if "done" in data:
            keep["done"] = np.ascontiguousarray(data["done"], dtype=np.float32)
            v.done = keep["done"].ctypes.data
        else:
            keep["terminals"] = np.ascontiguousarray(data["terminals"]).astype(np.uint8)
            keep["timeouts"] = np.ascontiguousarray(data["timeouts"]).astype(np.uint8)
            v.terminals, v.timeouts = keep["terminals"].ctypes.data, keep["timeouts"].ctypes.data
        v.reward_scale, v.cost_scale = float(reward_scale), float(cost_scale)
        if "is_init" in data:
            keep["is_init"] = np.ascontiguousarray(data["is_init"], dtype=np.float32)
            v.is_init = keep["is_init"].ctypes.data
        check(self.lib.osrl_buffer_upload(self.h, C.byref(v)))
        self.dataset_size = int(v.n)

    def upload_seq_dataset(self, trajs: dict, reward_scale: float = 1.0, cost_scale: float = 1.0) -> None:
        """trajs: flat per-transition arrays in trajectory order (observations, actions, returns, cost_returns,
        costs), `traj_offsets` [n_traj+1] and optional `sample_prob` [n_traj] -> osrl_seq_buffer_upload."""
        keep = {k: np.ascontiguousarray(trajs[k], dtype=np.float32)
                for k in ("observations", "actions", "returns", "cost_returns", "costs")}
        keep["traj_offsets"] = np.ascontiguousarray(trajs["traj_offsets"], dtype=np.int64)
        v = SeqDatasetView()
        v.n, v.n_traj = keep["observations"].shape[0], keep["traj_offsets"].shape[0] - 1
        for k, a in keep.items():
            setattr(v, k, a.ctypes.data)
        if trajs.get("sample_prob") is not None:
            keep["sample_prob"] = np.ascontiguousarray(trajs["sample_prob"], dtype=np.float64)
            v.sample_prob = keep["sample_prob"].ctypes.data
        v.reward_scale, v.cost_scale = float(reward_scale), float(cost_scale)
        check(self.lib.osrl_seq_buffer_upload(self.h, C.byref(v)))
        self.n_traj = int(v.n_traj)

    def preprocess_seq_dataset(self, data: dict, reward_scale: float = 1.0, cost_scale: float = 1.0,
                               cost_reverse: bool = False) -> Dict[str, np.ndarray]:
        """process_sequence_dataset (dataset.py:137-183) on the device -> the resident trajectory buffer
        (osrl_seq_preprocess).  `data` is the raw DSRL dictionary.  Returns the per-episode info the sampling options
        need: {"returns", "cost_returns"} = unscaled first return / cost return of each episode, "traj_offsets"."""
        keep = {k: np.ascontiguousarray(data[k], dtype=np.float32) for k in ("observations", "actions", "rewards", "costs")}
        keep["terminals"] = np.ascontiguousarray(data["terminals"]).astype(np.uint8)
        keep["timeouts"] = np.ascontiguousarray(data["timeouts"]).astype(np.uint8)
        v = DatasetView()
        v.n = keep["observations"].shape[0]
        for k, a in keep.items():
            setattr(v, k, a.ctypes.data)
        v.reward_scale, v.cost_scale = float(reward_scale), float(cost_scale)
        nt, nu = C.c_int64(), C.c_int64()
        check(self.lib.osrl_seq_preprocess(self.h, C.byref(v), int(bool(cost_reverse)), C.byref(nt), C.byref(nu)))
        self.n_traj = int(nt.value)
        r, c = np.empty(self.n_traj, np.float32), np.empty(self.n_traj, np.float32)
        off = np.empty(self.n_traj + 1, np.int64)
        check(self.lib.osrl_seq_episode_info(self.h, r.ctypes.data, c.ctypes.data, off.ctypes.data, self.n_traj))
        return {"returns": r, "cost_returns": c, "traj_offsets": off, "n_used": int(nu.value)}

    def set_seq_sample_prob(self, prob) -> None:
        """Categorical over the resident episodes (cost_sample / pf_sample, dataset.py:439-459); None = uniform."""
        if prob is None:
            check(self.lib.osrl_seq_set_sample_prob(self.h, None, 0))
            return
        p = np.ascontiguousarray(prob, dtype=np.float64)
        check(self.lib.osrl_seq_set_sample_prob(self.h, p.ctypes.data, p.shape[0]))

    def seq_gather(self, traj_idx, start_idx) -> Dict[str, torch.Tensor]:
        """[n, T, .] windows for explicit (trajectory, start) pairs -> CUDA tensors (bit-exact copies)."""
        ti = np.ascontiguousarray(traj_idx, dtype=np.int32)
        si = np.ascontiguousarray(start_idx, dtype=np.int32)
        n, T, o, a = ti.shape[0], self.cfg.seq_len, self.cfg.obs_dim, self.cfg.act_dim
        dev = f"cuda:{self.device}"
        out = {"states": torch.empty(n, T, o, device=dev), "actions": torch.empty(n, T, a, device=dev),
               "returns": torch.empty(n, T, device=dev), "costs_return": torch.empty(n, T, device=dev),
               "time_steps": torch.empty(n, T, dtype=torch.int64, device=dev), "mask": torch.empty(n, T, device=dev),
               "costs": torch.empty(n, T, device=dev)}
        b = SeqBatch()
        b.rows, b.seq_len, b.on_host = n, T, 0
        for k, t in out.items():
            setattr(b, k, t.data_ptr())
        check(self.lib.osrl_seq_gather(self.h, C.c_void_p(ti.ctypes.data), C.c_void_p(si.ctypes.data), n, C.byref(b),
                                       C.c_void_p(self._stream())))
        return out

    def alias_table(self):
        prob = np.empty(self.n_traj, dtype=np.float32)
        alias = np.empty(self.n_traj, dtype=np.int32)
        check(self.lib.osrl_seq_alias_table(self.h, C.c_void_p(prob.ctypes.data), C.c_void_p(alias.ctypes.data), self.n_traj))
        return prob, alias

    def last_sequences(self):
        t = np.empty(self.batch_size, dtype=np.int32)
        s = np.empty(self.batch_size, dtype=np.int32)
        check(self.lib.osrl_last_sequences(self.h, C.c_void_p(t.ctypes.data), C.c_void_p(s.ctypes.data), self.batch_size))
        return t, s

    def gather(self, idx) -> Dict[str, torch.Tensor]:
        """Bit-exact row gather by explicit indices -> six CUDA tensors."""
        idx = torch.as_tensor(idx, dtype=torch.int64).contiguous()
        n = idx.numel()
        o, a = self.cfg.obs_dim, self.cfg.act_dim
        dev = f"cuda:{self.device}"
        out = {"observations": torch.empty(n, o, device=dev), "next_observations": torch.empty(n, o, device=dev),
               "actions": torch.empty(n, a, device=dev), "rewards": torch.empty(n, device=dev),
               "costs": torch.empty(n, device=dev), "done": torch.empty(n, device=dev)}
        b = Batch()
        b.rows, b.on_host = n, 0
        for k in _BATCH_KEYS:
            setattr(b, k, out[k].data_ptr())
        check(self.lib.osrl_gather(self.h, C.c_void_p(idx.data_ptr()), n, 0 if idx.is_cuda else 1, C.byref(b),
                                   C.c_void_p(self._stream())))
        return out

    # ------------------------------------------------------------------ stepping
    def step(self, batch: dict, noise: Optional[dict] = None) -> None:
        """One train_one_step on an explicit minibatch (host or device tensors)."""
        keep = []
        b = Batch()
        b.rows = self.batch_size
        kinds = set()
        dims = {"observations": self.cfg.obs_dim, "next_observations": self.cfg.obs_dim, "actions": self.cfg.act_dim}
        for k in _STEP_KEYS:
            if k not in batch or batch[k] is None:
                continue
            t, p, on_host = _as_f32(batch[k], self.device)
            if t.numel() != self.batch_size * dims.get(k, 1):   # the C side reads batch_size * dim floats per tensor
                raise ValueError(f"{k}: expected {self.batch_size * dims.get(k, 1)} floats, got {t.numel()}")
            keep.append(t)
            kinds.add(on_host)
            setattr(b, k, p)
        if len(kinds) != 1:
            raise ValueError("all batch tensors must live on the same side (host or this GPU)")
        b.on_host = kinds.pop()
        nz_ptr = None
        if noise is not None:
            nz = Noise()
            nk = set()
            for i, name in enumerate(self.noise_layout):
                if name in noise and noise[name] is not None:
                    t, p, on_host = _as_f32(noise[name], self.device)
                    if t.numel() != self.noise_layout[name]:
                        raise ValueError(f"noise slot {name}: expected {self.noise_layout[name]} floats, got {t.numel()}")
                    keep.append(t)
                    nk.add(on_host)
                    nz.slot[i] = p
            if len(nk) > 1:
                raise ValueError("all noise tensors must live on the same side")
            nz.on_host = nk.pop() if nk else 0
            nz_ptr = C.byref(nz)
        check(self.lib.osrl_step(self.h, C.byref(b), nz_ptr, C.c_void_p(self._stream())))
        self._keep = keep  # pinned/device sources must outlive the async copies

    def step_seq(self, batch: dict, noise: Optional[dict] = None) -> None:
        """One CDT train_one_step on a collated SequenceDataset batch (host or device tensors).  `noise`: dropout
        multipliers by slot name (``noise_layout``) to replay; missing slots are drawn on the device."""
        keep, kinds = [], set()
        b = SeqBatch()
        b.rows, b.seq_len = self.batch_size, self.cfg.seq_len
        for k in ("states", "actions", "returns", "costs_return", "mask", "costs"):
            t, p, on_host = _as_f32(batch[k], self.device)
            keep.append(t); kinds.add(on_host)
            setattr(b, k, p)
        ts = batch["time_steps"]
        if isinstance(ts, np.ndarray):
            ts = torch.from_numpy(np.ascontiguousarray(ts))
        ts = ts.to(torch.int64).contiguous()
        if not ts.is_cuda and ts.numel():
            # the reference's nn.Embedding(episode_len + seq_len, E) raises on an index outside the table (cdt.py:75);
            # the kernels clamp, so bad host data is refused here (device tensors are not read back: that would
            # synchronise every step)
            lim = self.cfg.episode_len + self.cfg.seq_len
            if int(ts.min()) < 0 or int(ts.max()) >= lim:
                raise IndexError(f"time_steps outside the timestep embedding table [0, {lim})")
        keep.append(ts); kinds.add(0 if ts.is_cuda else 1)
        b.time_steps = ts.data_ptr()
        if len(kinds) != 1:
            raise ValueError("all batch tensors must live on the same side (host or this GPU)")
        b.on_host = kinds.pop()
        nz = None
        if noise is not None:
            nz = Noise()
            nkinds = set()
            for i, name in enumerate(self.noise_layout):
                if name in noise and noise[name] is not None:
                    t, p, on_host = _as_f32(noise[name], self.device)
                    if t.numel() != self.noise_layout[name]:
                        raise ValueError(f"noise slot {name}: expected {self.noise_layout[name]} floats, got {t.numel()}")
                    keep.append(t); nkinds.add(on_host)
                    nz.slot[i] = p
            if len(nkinds) > 1:
                raise ValueError("all noise tensors must live on the same side")
            nz.on_host = nkinds.pop() if nkinds else 1
        check(self.lib.osrl_step_seq(self.h, C.byref(b), C.byref(nz) if nz is not None else None,
                                     C.c_void_p(self._stream())))
        self._keep = keep

    def steps_host(self, batches, want_stats: bool = True):
        """k steps on k explicit host minibatches in ONE call (osrl_steps_host): `batches` is a list of batch dicts or
        one dict of stacked [k, B, ...] tensors.  Returns the k per-step stat dicts (or None with want_stats=False, in
        which case the call does not wait for the GPU)."""
        if isinstance(batches, (list, tuple)):
            keys = [k_ for k_ in _STEP_KEYS if k_ in batches[0] and batches[0][k_] is not None]
            batches = {k_: torch.stack([torch.as_tensor(b[k_], dtype=torch.float32).reshape(self.batch_size, -1)
                                        for b in batches]) for k_ in keys}
        k = int(batches["observations"].shape[0])
        dims = {"observations": self.cfg.obs_dim, "next_observations": self.cfg.obs_dim, "actions": self.cfg.act_dim}
        b = Batch()
        b.rows, b.on_host = self.batch_size, 1
        keep = []
        for name in _STEP_KEYS:
            if name not in batches or batches[name] is None:
                continue
            t, p, on_host = _as_f32(batches[name], self.device)
            if not on_host:
                raise ValueError("steps_host takes host tensors (device batches: step())")
            want = k * self.batch_size * dims.get(name, 1)
            if t.numel() != want:
                raise ValueError(f"{name}: expected {want} floats ([{k}, {self.batch_size}, ...]), got {t.numel()}")
            keep.append(t)
            setattr(b, name, p)
        out = (C.c_float * (k * len(self.stat_names)))() if want_stats else None
        check(self.lib.osrl_steps_host(self.h, C.byref(b), k, out, C.c_void_p(self._stream())))
        if out is None:
            return None
        n = len(self.stat_names)
        return [{self.stat_names[i]: float(out[j * n + i]) for i in range(n)} for j in range(k)]

    def steps(self, k: int) -> None:
        """k steps sampled on the device from the resident dataset."""
        check(self.lib.osrl_steps(self.h, int(k), C.c_void_p(self._stream())))

    def stats(self) -> Dict[str, float]:
        buf = (C.c_float * 16)()
        n = C.c_int()
        check(self.lib.osrl_stats(self.h, buf, 16, C.byref(n), C.c_void_p(self._stream())))
        return {self.stat_names[i]: float(buf[i]) for i in range(n.value)}

    def stats_lagged(self) -> Optional[Dict[str, float]]:
        """Stats of the PREVIOUS call's step, without synchronising the stream (None on the first call)."""
        buf = (C.c_float * 16)()
        n, valid = C.c_int(), C.c_int()
        check(self.lib.osrl_stats_lagged(self.h, buf, 16, C.byref(n), C.byref(valid), C.c_void_p(self._stream())))
        return {self.stat_names[i]: float(buf[i]) for i in range(n.value)} if valid.value else None

    # ------------------------------------------------------------------ resumable checkpoint
    def state_blob(self) -> torch.Tensor:
        """Everything a bit-exact resume needs (parameters, targets, Adam moments, PID / dual variables, step counters
        of the Philox streams) as one uint8 tensor; `load_state_blob` restores it into an engine of the same config."""
        nb = C.c_int64()
        check(self.lib.osrl_state_size(self.h, C.byref(nb)))
        t = torch.empty(nb.value, dtype=torch.uint8)
        check(self.lib.osrl_state_save(self.h, C.c_void_p(t.data_ptr()), nb.value))
        return t

    def load_state_blob(self, blob: torch.Tensor) -> None:
        t = blob.detach().to(torch.uint8).cpu().contiguous()
        check(self.lib.osrl_state_load(self.h, C.c_void_p(t.data_ptr()), t.numel()))

    def scalars(self) -> Dict[str, float]:
        names = (C.c_char_p * 16)()
        n = C.c_int()
        check(self.lib.osrl_scalar_names(self.h, names, 16, C.byref(n)))
        vals = (C.c_double * 16)()
        check(self.lib.osrl_scalars_get(self.h, vals, 16, C.byref(n)))
        return {names[i].decode(): float(vals[i]) for i in range(n.value)}

    def set_scalars(self, **kw) -> None:
        cur = self.scalars()
        cur.update(kw)
        vals = (C.c_double * len(cur))(*cur.values())
        check(self.lib.osrl_scalars_set(self.h, vals, len(cur)))

    def last_indices(self) -> np.ndarray:
        out = np.empty(self.batch_size, dtype=np.int64)
        check(self.lib.osrl_last_indices(self.h, C.c_void_p(out.ctypes.data), self.batch_size))
        return out

    def last_noise(self) -> Dict[str, np.ndarray]:
        out = {}
        for i, (name, cnt) in enumerate(self.noise_layout.items()):
            a = np.empty(cnt, dtype=np.float32)
            check(self.lib.osrl_last_noise(self.h, i, C.c_void_p(a.ctypes.data), cnt))
            out[name] = a
        return out

    def profile(self, reps: int = 20):
        """Per-launch (name, mean ms, algorithmic bytes, algorithmic flops) of one step, eager launches."""
        n = C.c_int()
        check(self.lib.osrl_profile(self.h, 1, C.byref(n), None, None, None, None, 0, C.c_void_p(self._stream())))
        k = n.value
        names, ms = (C.c_char_p * k)(), (C.c_double * k)()
        by, fl = (C.c_double * k)(), (C.c_double * k)()
        check(self.lib.osrl_profile(self.h, int(reps), C.byref(n), names, ms, by, fl, k, C.c_void_p(self._stream())))
        return [(names[i].decode(), float(ms[i]), float(by[i]), float(fl[i])) for i in range(k)]

    @property
    def launches(self) -> int:
        return int(self.lib.osrl_launch_count(self.h))

    @property
    def launches_per_step(self) -> int:
        return int(self.lib.osrl_launches_per_step(self.h))

    # ------------------------------------------------------------------ data parallel
    def init_comm(self, unique_id: bytes) -> None:
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        check(self.lib.osrl_comm_init(self.h, buf, self.cfg.world_size, self.cfg.rank))

    @property
    def dp_mode(self) -> str:
        """How the gradients travel between the ranks: "single", "nccl" (collectives in the step graph) or "peer"
        (summed out of NVLink peer memory inside the Adam kernel)."""
        return ("single", "nccl", "peer")[self.lib.osrl_dp_mode(self.h)]


def comm_unique_id() -> bytes:
    buf = (C.c_char * 128)()
    check(_lib.load().osrl_comm_unique_id(buf))
    return bytes(buf.raw)


def plan(algo: str, **hyper):
    """Parameter table (name, shape) without a GPU -- osrl_plan."""
    lib = _lib.load()
    cfg = make_config(algo, batch_size=hyper.pop("batch_size", 1), **hyper)
    n = C.c_int()
    check(lib.osrl_plan(C.byref(cfg), None, 0, C.byref(n)))
    arr = (ParamDesc * n.value)()
    check(lib.osrl_plan(C.byref(cfg), arr, n.value, C.byref(n)))
    return [(d.name.decode(), (d.rows, d.cols) if d.cols else (d.rows,), d.section, d.group) for d in arr]
