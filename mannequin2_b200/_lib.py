"""ctypes binding of libmq_b200.so (include/mq_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is present,
every numeric entry point of this package raises.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libmq_b200.so")

MQ_OK, MQ_ERR_INVALID, MQ_ERR_SHAPE, MQ_ERR_UNSUPPORTED, MQ_ERR_CUDA = 0, -1, -2, -3, -4
MQ_F32, MQ_F64 = 0, 1
ACT_NONE, ACT_TANH, ACT_LRELU = 0, 1, 2
HEAD_DY, HEAD_DIFF, HEAD_SOFTMAX_CE, HEAD_GAUSS_LOGP, HEAD_GAUSS_PPO = 0, 1, 2, 3, 4
HEAD_DISCRETE_LOGP, HEAD_DISCRETE_PPO = 5, 6
MAX_LAYERS = 8


class MqNet(ctypes.Structure):
    _fields_ = [("n_layers", c_int32), ("n_in", c_int32), ("n_out", c_int32),
                ("n_params", c_int32), ("logstd_off", c_int32),
                ("in_", c_int32 * MAX_LAYERS), ("out", c_int32 * MAX_LAYERS),
                ("act", c_int32 * MAX_LAYERS), ("leak", c_float * MAX_LAYERS),
                ("w_off", c_int32 * MAX_LAYERS), ("b_off", c_int32 * MAX_LAYERS)]


class MqHead(ctypes.Structure):
    _fields_ = [("kind", c_int32), ("p0", c_float), ("p1", c_float), ("ent", c_float), ("dy", c_void_p),
                ("target", c_void_p), ("adv", c_void_p), ("logp_old", c_void_p),
                # per-call options of the tensor-core kernels (all zero = automatic, exact): packed records, MQ_TC_* flags,
                # operand image
                ("records", c_void_p), ("rec_w", c_int32), ("tc", c_int32), ("tc_image", c_void_p)]


TC_FORCE, TC_OFF = 0x10, 0x20          # MQ_TC_* flags; the low two bits select the bf16 planes (0 = 3 = exact)
TC_FP16 = 0x80                         # MQ_TC_FP16: with two planes, fp16 planes with the low one scaled by 2^11
TC_KEEP_X = 0x100                      # MQ_TC_KEEP_X: layered MLP forward leaves the planes of x for dW_0 of the same step
TC_MODE_BITS = 0x83
TC_MODES = {"exact": 3, "bf16x3": 3, "fp16x2": 2 | TC_FP16, "bf16x2": 2, "bf16": 1}


class MqVae(ctypes.Structure):
    """mq_vae_t: the net of mnist/vae.py for the fused step (csrc/mq_vae.cu)."""
    _fields_ = [("batch", c_int64), ("d_img", c_int64), ("hidden", c_int64), ("latent", c_int64), ("n_hidden", c_int32),
                ("tc", c_int32), ("clip_lo", c_float), ("clip_hi", c_float)]


class MqEnv(ctypes.Structure):
    """mq_env_t: the synthetic on-device environment of the rollout kernels."""
    _fields_ = [("d", c_int32), ("a", c_int32), ("horizon", c_int32), ("dt", c_float), ("damp", c_float),
                ("sigma", c_float), ("ctrl", c_float), ("p_end", c_float), ("b", c_void_p), ("obs_scale", c_void_p),
                ("obs_shift", c_void_p)]


class MqPeer(ctypes.Structure):
    """mq_peer_t: peer-memory descriptor of the in-kernel gradient all-reduce."""
    _fields_ = [("rank", c_int32), ("world", c_int32), ("grad_bufs", c_void_p), ("flag_bufs", c_void_p),
                ("buf_stride", c_int64), ("step", ctypes.c_uint32), ("pull", c_int32)]


class MqStep(ctypes.Structure):
    """mq_step_t: one parameter vector's optimiser step (mq_reduce_step2)."""
    _fields_ = [("partials", c_void_p), ("n_parts", c_int32), ("n", c_int64), ("scale", c_float), ("value", c_void_p),
                ("m", c_void_p), ("v", c_void_p), ("theta_out", c_void_p), ("grad_out", c_void_p), ("c1", c_double),
                ("c2", c_double), ("lr", c_double), ("eps", c_double), ("peer", c_void_p), ("net", c_void_p), ("tc", c_int32),
                ("image", c_void_p)]


class MqTrain(ctypes.Structure):
    """mq_train_t: one net's side of mq_train_steps2."""
    _fields_ = [("net", c_void_p), ("theta", c_void_p), ("x", c_void_p), ("head", c_void_p), ("idx_table", c_void_p),
                ("idx_stride", c_int64), ("scale", c_float), ("value", c_void_p), ("m", c_void_p), ("v", c_void_p),
                ("coefs_host", c_void_p), ("lr", c_double), ("eps", c_double), ("peer", c_void_p), ("ws", c_void_p),
                ("ws_bytes", c_int64)]


P = c_void_p
NETP = POINTER(MqNet)
HEADP = POINTER(MqHead)
PEERP = POINTER(MqPeer)

SIGNATURES = {
    "mq_last_error": (c_char_p, []),
    "mq_version": (c_int, []),
    "mq_bench_ffma": (c_int, [c_int32, P, POINTER(c_double), P]),
    "mq_bench_tensor": (c_int, [c_int32, c_int32, POINTER(c_double), P]),
    "mq_debug_set_phase_buffer": (c_int, [P]),
    "mq_device_props": (c_int, [POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), POINTER(c_int64)]),
    "mq_adam_step": (c_int, [P, P, P, P, P, c_int64, c_int, c_int, c_double, c_double, c_double,
                             c_double, c_int, P]),
    "mq_adam_step_dev": (c_int, [P, P, P, P, P, c_int64, c_int, c_int, P, c_int, P]),
    "mq_norm_stats_ws_bytes": (c_int64, [c_int64, c_int64]),
    "mq_norm_stats": (c_int, [P, c_int, c_int64, c_int64, P, c_double, P, P, c_int64, P]),
    "mq_norm_update_from_stats": (c_int, [P, P, P, c_int64, c_double, c_double, c_double, P]),
    "mq_running_mean_update": (c_int, [P, P, c_int, c_int64, c_double, P, P]),
    "mq_norm_ws_bytes": (c_int64, [c_int64, c_int64]),
    "mq_col_mean": (c_int, [P, c_int, c_int64, c_int64, c_double, P, P, c_int64, P]),
    "mq_norm_cov": (c_int, [P, c_int, c_int64, c_int64, P, P, c_double, P, P, c_int64, P]),
    "mq_norm_apply": (c_int, [P, c_int, c_int64, c_int64, P, P, P, c_int, c_int, P]),
    "mq_gather_rows": (c_int, [P, P, c_int64, c_int64, c_int64, P, P]),
    "mq_scan_ws_bytes": (c_int64, [c_int64]),
    "mq_gae": (c_int, [P, P, P, c_int64, c_double, c_double, P, P, P, c_int64, P]),
    "mq_discounted": (c_int, [P, c_int64, c_double, P, P, c_int64, P]),
    "mq_mlp_acts_floats": (c_int64, [NETP, c_int64]),
    "mq_mlp_ws_bytes": (c_int64, [NETP, c_int64]),
    "mq_mlp_forward": (c_int, [NETP, P, P, P, c_int64, P, P]),
    "mq_mlp_forward_ws": (c_int, [NETP, P, P, P, c_int64, P, P, c_int64, P]),
    "mq_mlp_forward_tc": (c_int, [NETP, P, P, P, c_int64, P, c_int32, P, c_int64, P]),
    "mq_mlp_backward_tc": (c_int, [NETP, P, P, P, P, P, c_int64, c_float, P, P, c_int32, P, c_int64, P, P, P]),
    "mq_mlp_backward": (c_int, [NETP, P, P, P, P, P, c_int64, c_float, P, P, P, c_int64, P]),
    "mq_mlp_backward_overlap": (c_int, [NETP, P, P, P, P, P, c_int64, c_float, P, P, P, c_int64, P, P, P]),
    "mq_mlp_train_block_ws_bytes": (c_int64, [NETP, c_int64]),
    "mq_mlp_train_block": (c_int, [NETP, P, P, P, P, c_int, P, P, P, c_int64, c_int32, P, c_float, P, c_int64, P]),
    "mq_fused_tile_rows": (c_int, [NETP, c_int]),
    "mq_fused_ws_bytes": (c_int64, [NETP, c_int64]),
    "mq_fused_forward": (c_int, [NETP, P, P, P, c_int64, P, P, P, P]),
    "mq_fused_forward_tc": (c_int, [NETP, P, P, P, c_int64, P, P, P, c_int32, P]),
    "mq_fused_forward_rec": (c_int, [NETP, P, P, c_int32, P, P, c_int64, P, P, c_int32, P]),
    "mq_record_width": (c_int32, [NETP]),
    "mq_pack_records": (c_int, [NETP, P, P, P, P, c_int64, P, P]),
    "mq_tc_image_bytes": (c_int64, [NETP, c_int32]),
    "mq_tc_image_build": (c_int, [NETP, P, c_int32, P, P]),
    "mq_fused_grad": (c_int, [NETP, P, P, P, c_int64, HEADP, c_float, P, P, P, P, c_int64, P]),
    "mq_fused_grad_partials": (c_int, [NETP, P, P, P, c_int64, HEADP, P, P, P, c_int64, POINTER(c_int32), P]),
    "mq_step_blocks": (c_int64, [c_int64]),
    "mq_reduce_step": (c_int, [P, c_int32, c_int64, c_float, P, P, P, P, P, c_int, c_double, c_double, c_double,
                               c_double, c_int, PEERP, P]),
    "mq_reduce_step_image": (c_int, [P, c_int32, c_int64, c_float, P, P, P, P, P, c_int, c_double, c_double, c_double,
                                     c_double, c_int, PEERP, NETP, c_int32, P, P]),
    "mq_reduce_step2": (c_int, [POINTER(MqStep), POINTER(MqStep), c_int, c_int, P]),
    "mq_fused_grad_partials2": (c_int, [NETP, P, P, HEADP, P, c_int64, POINTER(c_int32), NETP, P, P, HEADP, P, c_int64,
                                        POINTER(c_int32), c_int64, P]),
    "mq_train_steps2": (c_int, [POINTER(MqTrain), POINTER(MqTrain), c_int32, c_int64, c_int, c_int, P]),
    "mq_train_steps": (c_int, [NETP, P, P, HEADP, P, c_int64, c_int32, c_int64, c_float, P, P, P, c_int, POINTER(c_double), c_double,
                               c_double, c_int, PEERP, P, c_int64, P]),
    "mq_fused_train": (c_int, [NETP, P, P, P, P, c_int, P, HEADP, P, c_int32, c_int32, c_double,
                               c_double, POINTER(c_double), c_double, c_double, P]),
    "mq_es_eval": (c_int, [NETP, P, P, c_int64, P, c_int64, P, P, P]),
    "mq_es_weighted_sum": (c_int, [P, P, c_int64, c_int64, c_double, P, P, c_int64, P]),
    "mq_es_ws_bytes": (c_int64, [c_int64, c_int64]),
    "mq_gemm": (c_int, [P, c_int64, c_int64, P, c_int64, c_int64, P, c_int64, c_int64, c_int64,
                        c_float, P]),
    "mq_gemm_ws": (c_int, [P, c_int64, c_int64, P, c_int64, c_int64, P, c_int64, c_int64, c_int64, c_float, P, c_int64, P]),
    "mq_gemm_ws_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "mq_gemm_ex": (c_int, [P, c_int64, c_int64, P, c_int64, c_int64, P, c_int64, c_int64, c_int64, c_float, c_int32, P, c_int64,
                           P]),
    "mq_add_bias": (c_int, [P, P, P, c_int64, c_int64, P]),
    "mq_binary": (c_int, [P, P, P, c_int64, c_int64, c_int, P]),
    "mq_col_sum": (c_int, [P, P, c_int64, c_int64, c_float, P, P]),
    "mq_act_forward": (c_int, [P, P, c_int64, c_int, c_float, P]),
    "mq_act_backward": (c_int, [P, P, P, c_int64, c_int, c_float, P]),
    "mq_clip_forward": (c_int, [P, P, c_int64, c_float, c_float, P]),
    "mq_clip_backward": (c_int, [P, P, P, c_int64, c_float, c_float, P]),
    "mq_gauss_logprob": (c_int, [P, P, c_int64, P, c_int64, c_int64, P, P]),
    "mq_obsnorm_step": (c_int, [P, c_int64, c_int32, ctypes.c_double, P, c_float, P, P]),
    "mq_randn_philox": (c_int, [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, P, c_int64, P, P]),
    "mq_env_reset": (c_int, [POINTER(MqEnv), c_int64, ctypes.c_uint64, P, P, P, P]),
    "mq_actor_env_step": (c_int, [POINTER(MqEnv), c_int64, c_int32, c_int32, ctypes.c_uint64, ctypes.c_uint32, P, P, P,
                                  P, P, P, P, P, P, P, P, P]),
    "mq_bootstrap_segments": (c_int, [P, P, P, c_int64, c_int64, c_float, P]),
    "mq_gauss_entropy": (c_int, [P, P, c_int64, c_int64, P, P, P]),
    "mq_gauss_logprob_bwd": (c_int, [P, P, c_int64, P, P, c_int64, c_int64, P, P, P]),
    "mq_discrete_logprob": (c_int, [P, P, P, c_int64, c_int64, P, P, P]),
    "mq_momentum_step": (c_int, [P, P, P, P, c_int64, c_int64, c_float, c_float, c_float, P]),
    "mq_onehot": (c_int, [P, c_int64, c_int64, P, P]),
    "mq_vae_n_params": (c_int64, [P]),
    "mq_vae_ws_bytes": (c_int64, [P]),
    "mq_vae_ws_init": (c_int, [P, P, c_int64, P]),
    "mq_vae_grad": (c_int, [P, P, P, P, P, P, P, P, c_int64, P]),
    "mq_vae_step": (c_int, [P, P, P, P, P, P, P, P, c_float, c_float, c_float, P, c_int64, P]),
    "mq_gauss_sample": (c_int, [P, P, c_int64, P, c_int64, c_int64, P, P, P]),
    "mq_dkl_forward": (c_int, [P, P, c_int64, c_int64, P, P]),
    "mq_dkl_backward": (c_int, [P, P, P, c_int64, c_int64, P, P, P]),
    "mq_softmax_ce_grad": (c_int, [P, P, P, c_int64, c_int64, c_float, P, P]),
    "mq_ppo_weight": (c_int, [P, P, P, P, c_int64, c_float, c_float, P, P]),
}


def bind(cdll):
    """Attach argtypes/restype for every symbol include/mq_b200.h declares."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(cdll, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return cdll


_lib = None


def lib():
    """The loaded CUDA library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "mannequin2_b200: %s is missing -- build it with "
                "`python -m mannequin2_b200.build` (there is no CPU fallback)" % LIB_PATH)
        _lib = bind(ctypes.CDLL(LIB_PATH))
    return _lib


class MqError(RuntimeError):
    pass


def check(rc, cdll=None):
    """Map MQ_ERR_* to the exception types the reference raises."""
    if rc == MQ_OK:
        return
    msg = (cdll or lib()).mq_last_error().decode("utf-8", "replace")
    if rc == MQ_ERR_INVALID:
        raise ValueError(msg)
    if rc == MQ_ERR_SHAPE:
        raise AssertionError(msg)
    if rc == MQ_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise MqError(msg)


def make_net(n_in, widths, acts, leak=0.1, logstd=False):
    """mq_net_t for Input(n_in) -> Affine(w)+act ... in the reference's flat layout."""
    if len(widths) > MAX_LAYERS:
        raise NotImplementedError("at most %d Affine layers" % MAX_LAYERS)
    net = MqNet()
    net.n_layers, net.n_in = len(widths), int(n_in)
    pos, k = 0, int(n_in)
    leaks = leak if isinstance(leak, (list, tuple)) else [leak] * len(widths)
    for l, (n, a) in enumerate(zip(widths, acts)):
        net.in_[l], net.out[l], net.act[l], net.leak[l] = k, int(n), int(a), float(leaks[l])
        net.w_off[l], net.b_off[l] = pos, pos + k * int(n)
        pos += k * int(n) + int(n)
        k = int(n)
    net.n_out = k
    net.logstd_off = pos if logstd else -1
    net.n_params = pos + (k if logstd else 0)
    return net
