"""ctypes binding of libxmeta.so (the C ABI declared in include/xmeta.h).

There is no fallback: if the shared library is missing ``load()`` raises, and every product path
that computes anything goes through ``load()``.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_double, c_float, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libxmeta.so')

XM_CONV_FWD, XM_CONV_DGRAD = 0, 1
XM_STAT_NONE, XM_STAT_SUM_SQ, XM_STAT_SUM_AUX = 0, 1, 2


class XmBlockGeom(Structure):
    _fields_ = [('tasks', c_int32), ('n', c_int32), ('cin', c_int32), ('cout', c_int32),
                ('hin', c_int32), ('win', c_int32), ('hz', c_int32), ('wz', c_int32),
                ('hp', c_int32), ('wp', c_int32), ('stride', c_int32), ('pool', c_int32)]


class XmConvArgs(Structure):
    _fields_ = [('g', XmBlockGeom), ('mode', c_int32), ('src_nchw', c_int32),
                ('row0', c_int32), ('row_step', c_int32), ('rows_per_task', c_int32),
                ('stat_mode', c_int32),
                ('src1', c_void_p), ('w1', c_void_p), ('w1_task_stride', c_int64),
                ('src2', c_void_p), ('w2', c_void_p), ('w2_task_stride', c_int64),
                ('out', c_void_p), ('aux', c_void_p), ('stats', c_void_p),
                ('workspace', c_void_p), ('workspace_bytes', c_int64)]


class XmWgradArgs(Structure):
    _fields_ = [('g', XmBlockGeom), ('src_nchw', c_int32), ('row0', c_int32), ('row_step', c_int32),
                ('rows_per_task', c_int32),
                ('x1', c_void_p), ('g1', c_void_p), ('x2', c_void_p), ('g2', c_void_p),
                ('out_w', c_void_p), ('out_b', c_void_p), ('out_task_stride', c_int64),
                ('base_w', c_void_p), ('base_b', c_void_p), ('base_task_stride', c_int64),
                ('scale', c_float), ('partial', c_void_p), ('partial_bytes', c_int64)]


class XmBnArgs(Structure):
    _fields_ = [('g', XmBlockGeom), ('eps', c_float),
                ('z', c_void_p), ('zdot', c_void_p), ('sums', c_void_p), ('dsums', c_void_p),
                ('gamma', c_void_p), ('beta', c_void_p), ('gb_task_stride', c_int64),
                ('gamma_dot', c_void_p), ('beta_dot', c_void_p), ('gbdot_task_stride', c_int64),
                ('mean_invstd', c_void_p), ('call_stats', c_void_p), ('bwd_red', c_void_p),
                ('dual_red', c_void_p),
                ('p', c_void_p), ('pdot', c_void_p), ('gp', c_void_p), ('gpdot', c_void_p),
                ('gz', c_void_p), ('gzdot', c_void_p),
                ('out_gamma', c_void_p), ('out_beta', c_void_p), ('out_task_stride', c_int64),
                ('base_gamma', c_void_p), ('base_beta', c_void_p), ('base_task_stride', c_int64),
                ('scale', c_float), ('scratch', c_void_p)]


class XmImgArgs(Structure):
    _fields_ = [('g', XmBlockGeom), ('row0', c_int32), ('row_step', c_int32), ('rows_per_task', c_int32),
                ('eps', c_float), ('scale', c_float),
                ('x', c_void_p), ('gram', c_void_p),
                ('w', c_void_p), ('w_task_stride', c_int64),
                ('w_dot', c_void_p), ('wdot_task_stride', c_int64),
                ('gamma', c_void_p), ('beta', c_void_p), ('gb_task_stride', c_int64),
                ('gamma_dot', c_void_p), ('beta_dot', c_void_p), ('gbdot_task_stride', c_int64),
                ('mean_invstd', c_void_p), ('call_stats', c_void_p), ('bwd_red', c_void_p),
                ('dual_red', c_void_p),
                ('p', c_void_p), ('zsel', c_void_p), ('sel', c_void_p),
                ('pdot', c_void_p), ('zdsel', c_void_p),
                ('gp', c_void_p), ('gpdot', c_void_p),
                ('ssum', c_void_p), ('scratch', c_void_p),
                ('out_w', c_void_p), ('out_b', c_void_p), ('out_gamma', c_void_p), ('out_beta', c_void_p),
                ('out_task_stride', c_int64),
                ('base_w', c_void_p), ('base_b', c_void_p), ('base_gamma', c_void_p), ('base_beta', c_void_p),
                ('base_task_stride', c_int64)]


class XmSampleArgs(Structure):
    _fields_ = [('tasks', c_int32), ('ways', c_int32), ('shots2', c_int32), ('channels', c_int32),
                ('height', c_int32), ('width', c_int32), ('num_classes', c_int32), ('rotate', c_int32),
                ('seed', c_uint64), ('first_task', c_int64),
                ('data', c_void_p), ('class_start', c_void_p),
                ('scale', c_float), ('offset', c_float),
                ('x', c_void_p), ('y', c_void_p), ('items', c_void_p), ('classes', c_void_p)]


class XmHeadArgs(Structure):
    _fields_ = [('tasks', c_int32), ('n', c_int32), ('ways', c_int32), ('c', c_int32), ('hw', c_int32),
                ('mode', c_int32), ('dual', c_int32),
                ('feat', c_void_p), ('feat_dot', c_void_p),
                ('labels', c_void_p), ('label_row0', c_int32), ('label_row_step', c_int32),
                ('labels_per_task', c_int32),
                ('w', c_void_p), ('b', c_void_p), ('wb_task_stride', c_int64),
                ('w_dot', c_void_p), ('b_dot', c_void_p), ('wbdot_task_stride', c_int64),
                ('loss', c_void_p), ('correct', c_void_p), ('logits', c_void_p),
                ('g_feat', c_void_p), ('g_feat_dot', c_void_p),
                ('out_w', c_void_p), ('out_b', c_void_p), ('out_task_stride', c_int64),
                ('base_w', c_void_p), ('base_b', c_void_p), ('base_task_stride', c_int64),
                ('scale', c_float)]


class XmAnilHeadArgs(Structure):
    _fields_ = [('tasks', c_int32), ('rows', c_int32), ('ways', c_int32), ('c', c_int32), ('hw', c_int32),
                ('mode', c_int32), ('steps', c_int32), ('first_order', c_int32), ('lr', c_float),
                ('feat', c_void_p), ('labels', c_void_p), ('w', c_void_p), ('b', c_void_p),
                ('loss', c_void_p), ('correct', c_void_p), ('g_feat', c_void_p),
                ('g_w', c_void_p), ('g_b', c_void_p), ('g_task_stride', c_int64),
                ('scratch', c_void_p), ('scratch_bytes', c_int64)]


class XmAdamArgs(Structure):
    _fields_ = [('theta', c_void_p), ('m', c_void_p), ('v', c_void_p), ('n_params', c_int64),
                ('local', c_void_p), ('reduced', c_void_p), ('n_total', c_int64),
                ('grad_scale', c_float), ('lr', c_float), ('beta1', c_float), ('beta2', c_float), ('eps', c_float),
                ('step', c_void_p)]


class XmRlAdvArgs(Structure):
    _fields_ = [('replays', c_int32), ('n', c_int32), ('state_dim', c_int32),
                ('gamma', c_double), ('tau', c_double), ('reg', c_double), ('coef_scale', c_double),
                ('states', c_void_p), ('next_states', c_void_p), ('rewards', c_void_p), ('dones', c_void_p),
                ('coef', c_void_p), ('returns', c_void_p), ('advantages', c_void_p)]


class XmRlSweepArgs(Structure):
    _fields_ = [('tasks', c_int32), ('n', c_int32), ('in_dim', c_int32), ('out_dim', c_int32), ('h1', c_int32),
                ('h2', c_int32), ('activation', c_int32), ('loss', c_int32), ('what', c_int32),
                ('states', c_void_p), ('actions', c_void_p), ('coef', c_void_p),
                ('mu_old', c_void_p), ('logstd_old', c_void_p), ('kl_scale', c_float),
                ('theta', c_void_p), ('theta_task_stride', c_int64),
                ('theta_dot', c_void_p), ('theta_dot_task_stride', c_int64),
                ('out', c_void_p), ('out_task_stride', c_int64),
                ('base', c_void_p), ('base_task_stride', c_int64), ('scale', c_float),
                ('task_loss', c_void_p), ('task_kl', c_void_p), ('mu_out', c_void_p),
                ('partial', c_void_p), ('partial_bytes', c_int64), ('clip', c_float), ('head_only', c_int32)]


XM_RL_A2C, XM_RL_SURROGATE, XM_RL_FISHER = 0, 1, 2
XM_RL_FORWARD, XM_RL_GRAD, XM_RL_HVP = 0, 1, 2
XM_ACT_RELU, XM_ACT_TANH = 0, 1
XM_COMM_MAX_WORLD, XM_IPC_HANDLE_BYTES = 8, 64

# name -> (restype, argtypes): every symbol include/xmeta.h declares.
SYMBOLS = {
    'xm_conv': (c_int32, [POINTER(XmConvArgs), c_void_p]),
    'xm_conv_workspace_bytes': (c_int64, [POINTER(XmBlockGeom)]),
    'xm_wgrad_scratch_bytes': (c_int64, [POINTER(XmBlockGeom)]),
    'xm_wgrad': (c_int32, [POINTER(XmWgradArgs), c_void_p]),
    'xm_bn_scratch_bytes': (c_int64, [POINTER(XmBlockGeom)]),
    'xm_bn_fwd': (c_int32, [POINTER(XmBnArgs), c_void_p]),
    'xm_bn_bwd': (c_int32, [POINTER(XmBnArgs), c_void_p]),
    'xm_bn_dual_fwd': (c_int32, [POINTER(XmBnArgs), c_void_p]),
    'xm_bn_dual_bwd': (c_int32, [POINTER(XmBnArgs), c_void_p]),
    'xm_img_supported': (c_int32, [POINTER(XmBlockGeom)]),
    'xm_img_gram_bytes': (c_int64, [POINTER(XmBlockGeom)]),
    'xm_img_scratch_bytes': (c_int64, [POINTER(XmBlockGeom)]),
    'xm_img_gram': (c_int32, [POINTER(XmImgArgs), c_void_p]),
    'xm_img_fwd': (c_int32, [POINTER(XmImgArgs), c_void_p]),
    'xm_img_bwd': (c_int32, [POINTER(XmImgArgs), c_void_p]),
    'xm_img_dual_fwd': (c_int32, [POINTER(XmImgArgs), c_void_p]),
    'xm_img_dual_bwd': (c_int32, [POINTER(XmImgArgs), c_void_p]),
    'xm_head': (c_int32, [POINTER(XmHeadArgs), c_void_p]),
    'xm_anil_head_scratch_bytes': (c_int64, [POINTER(XmAnilHeadArgs)]),
    'xm_anil_head': (c_int32, [POINTER(XmAnilHeadArgs), c_void_p]),
    'xm_sample_tasks': (c_int32, [POINTER(XmSampleArgs), c_void_p]),
    'xm_accumulate_tasks': (c_int32, [c_void_p, c_int64, c_int32, c_int64, c_void_p, c_int32, c_void_p]),
    'xm_adam_step': (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float,
                               c_float, c_float, c_float, c_int32, c_void_p]),
    'xm_comm_create': (c_int32, [c_int32, c_int32, c_int64, POINTER(c_void_p), c_void_p]),
    'xm_comm_connect': (c_int32, [c_void_p, c_void_p]),
    'xm_comm_error': (c_int32, [c_void_p]),
    'xm_comm_destroy': (c_int32, [c_void_p]),
    'xm_allreduce_adam': (c_int32, [c_void_p, POINTER(XmAdamArgs), c_void_p]),
    'xm_finish_shard': (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    'xm_rl_advantages': (c_int32, [POINTER(XmRlAdvArgs), c_void_p]),
    'xm_rl_sweep_scratch_bytes': (c_int64, [POINTER(XmRlSweepArgs)]),
    'xm_rl_sweep': (c_int32, [POINTER(XmRlSweepArgs), c_void_p]),
    'xm_bn_ema': (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32, c_int64, c_int32,
                            c_float, c_void_p]),
    'xm_set_precision': (c_int32, [c_int32]),
    'xm_set_tcgen05': (c_int32, [c_int32]),
    'xm_version': (c_int32, []),
    'xm_last_error': (ctypes.c_char_p, []),
    'xm_launch_count': (c_int64, []),
}

_lib = None


class XmetaError(RuntimeError):
    pass


def load():
    """Loads libxmeta.so (built by ``__graft_entry__.build()`` / ``python -m exploring_meta_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise XmetaError('libxmeta.so not found at %s -- build it with `python -m exploring_meta_b200.build` '
                         '(there is no CPU or PyTorch fallback)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if os.environ.get('XM_PRECISION'):            # A/B switch of the contraction precision (see xm_set_precision)
        lib.xm_set_precision(int(os.environ['XM_PRECISION']))
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().xm_last_error()
        raise XmetaError('%s failed (code %d): %s' % (what, code, msg.decode() if msg else ''))
