"""ctypes binding of libosrl_b200.so (the C ABI declared in include/osrl_b200.h).

The library is loaded from the package directory (built in-tree by ``osrl_b200.build``).
There is no fallback: if the shared library is missing or a call fails, a
``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("OSRL_B200_LIBNAME", "libosrl_b200.so"))

OSRL_MAX_HIDDEN = 4
OSRL_MAX_NOISE = 32
ALGO = {"bc": 0, "bcql": 1, "cpq": 2, "bearl": 3, "cdt": 4, "coptidice": 5}


class Config(C.Structure):
    _fields_ = [
        ("algo", C.c_int32), ("obs_dim", C.c_int32), ("act_dim", C.c_int32), ("max_action", C.c_float),
        ("n_a_hidden", C.c_int32), ("a_hidden", C.c_int32 * OSRL_MAX_HIDDEN),
        ("n_c_hidden", C.c_int32), ("c_hidden", C.c_int32 * OSRL_MAX_HIDDEN),
        ("vae_hidden", C.c_int32), ("sample_action_num", C.c_int32),
        ("gamma", C.c_float), ("tau", C.c_float), ("phi", C.c_float), ("lmbda", C.c_float), ("beta", C.c_float),
        ("pid_kp", C.c_float), ("pid_ki", C.c_float), ("pid_kd", C.c_float),
        ("num_q", C.c_int32), ("num_qc", C.c_int32), ("cost_limit", C.c_float), ("episode_len", C.c_int32),
        ("qc_scalar", C.c_float), ("mmd_sigma", C.c_float), ("target_mmd_thresh", C.c_float),
        ("num_samples_mmd_match", C.c_int32), ("mmd_kernel", C.c_int32), ("start_update_policy_step", C.c_int32),
        ("actor_lr", C.c_float), ("critic_lr", C.c_float), ("vae_lr", C.c_float), ("alpha_lr", C.c_float),
        ("seq_len", C.c_int32), ("embedding_dim", C.c_int32), ("num_layers", C.c_int32), ("num_heads", C.c_int32),
        ("attention_dropout", C.c_float), ("residual_dropout", C.c_float), ("embedding_dropout", C.c_float),
        ("use_rew", C.c_int32), ("use_cost", C.c_int32), ("cost_transform", C.c_int32), ("stochastic", C.c_int32),
        ("init_temperature", C.c_float), ("target_entropy", C.c_float),
        ("learning_rate", C.c_float), ("weight_decay", C.c_float), ("adam_beta1", C.c_float),
        ("adam_beta2", C.c_float), ("clip_grad", C.c_float), ("lr_warmup_steps", C.c_int32),
        ("loss_cost_weight", C.c_float), ("loss_state_weight", C.c_float),
        ("batch_size", C.c_int32), ("seed", C.c_uint64), ("world_size", C.c_int32), ("rank", C.c_int32),
        ("f_type", C.c_int32), ("init_state_propotion", C.c_float), ("alpha", C.c_float), ("cost_ub_epsilon", C.c_float),
        ("num_nu", C.c_int32), ("num_chi", C.c_int32), ("scalar_lr", C.c_float),
        ("observations_std", C.c_void_p), ("actions_std", C.c_void_p),
    ]


class ParamDesc(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("rows", C.c_int64), ("cols", C.c_int64), ("offset", C.c_int64),
                ("section", C.c_int32), ("group", C.c_int32), ("ptr", C.c_void_p)]


class DatasetView(C.Structure):
    _fields_ = [("n", C.c_int64), ("observations", C.c_void_p), ("next_observations", C.c_void_p),
                ("actions", C.c_void_p), ("rewards", C.c_void_p), ("costs", C.c_void_p), ("done", C.c_void_p),
                ("terminals", C.c_void_p), ("timeouts", C.c_void_p), ("reward_scale", C.c_float),
                ("cost_scale", C.c_float), ("is_init", C.c_void_p)]


class Batch(C.Structure):
    _fields_ = [("rows", C.c_int32), ("on_host", C.c_int32), ("observations", C.c_void_p),
                ("next_observations", C.c_void_p), ("actions", C.c_void_p), ("rewards", C.c_void_p),
                ("costs", C.c_void_p), ("done", C.c_void_p), ("is_init", C.c_void_p)]


class SeqBatch(C.Structure):
    _fields_ = [("rows", C.c_int32), ("seq_len", C.c_int32), ("on_host", C.c_int32), ("states", C.c_void_p),
                ("actions", C.c_void_p), ("returns", C.c_void_p), ("costs_return", C.c_void_p),
                ("time_steps", C.c_void_p), ("mask", C.c_void_p), ("episode_cost", C.c_void_p), ("costs", C.c_void_p)]


class SeqDatasetView(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_traj", C.c_int64), ("observations", C.c_void_p), ("actions", C.c_void_p),
                ("returns", C.c_void_p), ("cost_returns", C.c_void_p), ("costs", C.c_void_p),
                ("traj_offsets", C.c_void_p), ("sample_prob", C.c_void_p), ("reward_scale", C.c_float),
                ("cost_scale", C.c_float)]


class Noise(C.Structure):
    _fields_ = [("on_host", C.c_int32), ("slot", C.c_void_p * OSRL_MAX_NOISE)]


# every symbol include/osrl_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("osrl_abi_version", C.c_int, []),
    ("osrl_last_error", C.c_char_p, []),
    ("osrl_plan", C.c_int, [C.POINTER(Config), C.POINTER(ParamDesc), C.c_int, C.POINTER(C.c_int)]),
    ("osrl_engine_create", C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(C.c_void_p)]),
    ("osrl_engine_destroy", None, [C.c_void_p]),
    ("osrl_param_table", C.c_int, [C.c_void_p, C.POINTER(ParamDesc), C.c_int, C.POINTER(C.c_int)]),
    ("osrl_param_set", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
    ("osrl_param_get", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
    ("osrl_sync_targets", C.c_int, [C.c_void_p]),
    ("osrl_buffer_upload", C.c_int, [C.c_void_p, C.POINTER(DatasetView)]),
    ("osrl_gather", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(Batch), C.c_void_p]),
    ("osrl_step", C.c_int, [C.c_void_p, C.POINTER(Batch), C.POINTER(Noise), C.c_void_p]),
    ("osrl_step_seq", C.c_int, [C.c_void_p, C.POINTER(SeqBatch), C.POINTER(Noise), C.c_void_p]),
    ("osrl_seq_buffer_upload", C.c_int, [C.c_void_p, C.POINTER(SeqDatasetView)]),
    ("osrl_seq_gather", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(SeqBatch), C.c_void_p]),
    ("osrl_seq_alias_table", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    ("osrl_seq_preprocess", C.c_int, [C.c_void_p, C.POINTER(DatasetView), C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("osrl_seq_episode_info", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    ("osrl_seq_set_sample_prob", C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    ("osrl_last_sequences", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    ("osrl_steps", C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    ("osrl_steps_host", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_int, C.c_void_p, C.c_void_p]),
    ("osrl_stat_names", C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_int)]),
    ("osrl_stats", C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int), C.c_void_p]),
    ("osrl_scalar_names", C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_int)]),
    ("osrl_scalars_get", C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]),
    ("osrl_scalars_set", C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int]),
    ("osrl_noise_layout", C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.c_int,
                                    C.POINTER(C.c_int)]),
    ("osrl_last_indices", C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    ("osrl_last_noise", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
    ("osrl_debug_linear", C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_void_p]),
    ("osrl_debug_gemm", C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_void_p]),
    ("osrl_debug_read", C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p]),
    ("osrl_stats_lagged", C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]),
    ("osrl_state_size", C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    ("osrl_state_save", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    ("osrl_state_load", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    ("osrl_profile", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_char_p), C.POINTER(C.c_double),
                               C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_void_p]),
    ("osrl_profile_was_in_graph", C.c_int, [C.c_void_p]),
    ("osrl_launch_count", C.c_int64, [C.c_void_p]),
    ("osrl_launches_per_step", C.c_int, [C.c_void_p]),
    ("osrl_comm_unique_id", C.c_int, [C.c_char * 128]),
    ("osrl_comm_init", C.c_int, [C.c_void_p, C.c_char * 128, C.c_int, C.c_int]),
    ("osrl_dp_mode", C.c_int, [C.c_void_p]),
]

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and type every entry point.  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m osrl_b200.build` (needs nvcc). "
            "osrl_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().osrl_last_error()
        raise RuntimeError(f"osrl_b200 error {rc}: {msg.decode() if msg else '?'}")
