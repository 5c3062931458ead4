"""ctypes binding of libppde_b200.so (C-ABI declared in include/ppde_b200.h).

There is no CPU fallback: if the shared library is missing this module raises, and every
product entry point that needs a kernel fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libppde_b200.so")

MAX_NETS = 4
MAX_SUBSTEPS = 32

vp = C.c_void_p


class PottsT(C.Structure):
    _fields_ = [("L", C.c_int32), ("Lp", C.c_int32), ("win_lo", C.c_int32), ("D", C.c_int32),
                ("Jsym", vp), ("h", vp), ("wt", vp), ("wt_H", C.c_float), ("_pad", C.c_int32)]


class CnnNetT(C.Structure):
    _fields_ = [("T0", vp), ("b0", vp), ("W1", vp), ("W1T", vp), ("b1", vp), ("d", vp), ("W0r", vp), ("W1p", vp),
                ("c", C.c_float), ("w1_scale", C.c_float), ("r1_scale", C.c_float), ("w0_scale", C.c_float),
                ("adj_scale", C.c_float), ("_pad", C.c_int32)]


class CnnT(C.Structure):
    _fields_ = [("n_nets", C.c_int32), ("C", C.c_int32), ("L", C.c_int32), ("P", C.c_int32),
                ("net", CnnNetT * MAX_NETS)]


class ChainsT(C.Structure):
    _fields_ = [("n", C.c_int32), ("chain_offset", C.c_int32), ("L", C.c_int32), ("aa_stride", C.c_int32),
                ("aa", vp), ("aa_y", vp), ("row_cur", vp), ("G", vp), ("Gp", vp),
                ("E", vp), ("fit", vp), ("E_y", vp), ("fit_y", vp), ("Epotts_y", vp),
                ("n_fixed", C.c_int32), ("row_wt", C.c_int32),
                ("E_fixed", vp), ("fit_fixed", vp), ("aa_fixed", vp), ("anchor_fixed", vp),
                ("U", vp), ("idx", vp), ("old_aa", vp), ("lqf", vp), ("lqr", vp), ("log_acc", vp), ("accept", vp),
                ("E_hist", vp), ("fit_hist", vp), ("best_E", vp), ("best_fit", vp), ("best_aa", vp),
                ("traj_aa", vp), ("traj_chain", C.c_int32), ("_pad", C.c_int32)]


class TuneT(C.Structure):
    """ppde_tune_t: per-call tuning / measurement switches (NULL = production defaults)."""
    _fields_ = [("parts", C.c_int32), ("forward_ctas", C.c_int32), ("delta_layout", C.c_int32), ("dbg", C.c_int32),
                ("prof", vp)]


class PasParamsT(C.Structure):
    _fields_ = [("S", C.c_int32), ("nmut_threshold", C.c_int32), ("paper_results", C.c_int32), ("t", C.c_int32),
                ("min_pos", C.c_int32), ("max_pos", C.c_int32), ("seed", C.c_uint64), ("uniforms", vp), ("t_dev", vp),
                ("full_trace", C.c_int32), ("comb_nets", C.c_int32), ("comb_vals", vp), ("comb_wl", vp),
                ("comb_vcap", C.c_int32), ("comb_rec", C.c_int32), ("comb_scale", C.c_float), ("fuse_potts", C.c_int32)]


# name -> (restype, argtypes); every symbol declared in include/ppde_b200.h
SIGNATURES = {
    "ppde_version": (C.c_char_p, []),
    "ppde_last_launch_count": (C.c_int, []),
    "ppde_potts_symmetrize": (C.c_int, [vp, C.c_int32, vp, vp]),
    "ppde_potts_full": (C.c_int, [C.POINTER(PottsT), vp, C.c_int32, C.c_int32, vp, C.c_int64, vp, vp]),
    "ppde_potts_incremental": (C.c_int, [C.POINTER(PottsT), C.POINTER(ChainsT), C.POINTER(PasParamsT), vp]),
    "ppde_potts_dense_image_bytes": (C.c_int64, [C.c_int32]),
    "ppde_potts_dense_pack": (C.c_int, [C.POINTER(PottsT), C.c_float, vp, vp]),
    "ppde_potts_dense_full": (C.c_int, [C.POINTER(PottsT), vp, C.c_float, vp, C.c_int32, C.c_int32, vp, C.c_int64, vp, vp]),
    "ppde_cnn_forward": (C.c_int, [C.POINTER(CnnT), vp, C.c_int32, C.c_int32, vp, vp]),
    "ppde_cnn_forward_tc": (C.c_int, [C.POINTER(CnnT), vp, C.c_int32, C.c_int32, vp, vp, C.POINTER(TuneT), vp]),
    "ppde_cnn_backward_combine": (C.c_int, [C.POINTER(CnnT), C.POINTER(PottsT), vp, C.c_int32, C.c_int32, vp,
                                            C.c_float, vp, C.c_int64, vp, vp, vp, C.c_int64, vp, vp, vp, vp]),
    "ppde_cnn_backward_tc": (C.c_int, [C.POINTER(CnnT), C.POINTER(PottsT), vp, C.c_int32, C.c_int32, vp, C.c_float,
                                       vp, C.c_int64, vp, vp, C.c_int64, vp, vp, vp, C.POINTER(TuneT), vp]),
    "ppde_cnn_backward_scratch_floats": (C.c_int64, [C.POINTER(CnnT), C.c_int32]),
    "ppde_cnn_backward_delta_layout": (C.c_int, [C.POINTER(CnnT), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                                 C.POINTER(C.c_int64)]),
    "ppde_pas_reverse_fuse_max_len": (C.c_int32, []),
    "ppde_cnn_backward_tc_rows": (C.c_int, [C.POINTER(CnnT), C.POINTER(PottsT), vp, C.c_int32, C.c_int32, vp, C.c_float,
                                            vp, C.c_int64, vp, vp, C.c_int64, vp, vp, vp, C.c_int32, vp, vp,
                                            C.POINTER(TuneT), vp]),
    "ppde_cnn_dirty": (C.c_int, [C.POINTER(CnnT), vp, vp, C.c_int32, C.c_int32, vp, vp]),
    "ppde_cnn_forward_inc_ws_bytes": (C.c_int64, [C.c_int32]),
    "ppde_cnn_block_positions": (C.c_int32, []),
    "ppde_cnn_forward_inc": (C.c_int, [C.POINTER(CnnT), vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp, vp, C.c_int32, vp, vp,
                                       C.POINTER(TuneT), vp]),
    "ppde_cnn_backward_delta": (C.c_int, [C.POINTER(CnnT), C.POINTER(PottsT), vp, vp, C.c_int32, C.c_int32, vp, vp, C.c_float,
                                          vp, C.c_int64, vp, C.c_int64, vp, vp, vp, vp, vp, C.POINTER(TuneT), vp]),
    "ppde_potts_energy_rows": (C.c_int, [C.POINTER(PottsT), vp, C.c_int32, C.c_int32, vp, C.c_int64, vp, vp, vp]),
    "ppde_oracle_ridge": (C.c_int, [vp, C.c_float, C.c_float, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp]),
    "ppde_step_rows": (C.c_int, [C.POINTER(ChainsT), vp, vp]),
    "ppde_pas_propose": (C.c_int, [C.POINTER(PottsT), C.POINTER(ChainsT), C.POINTER(PasParamsT), vp]),
    "ppde_pas_reverse_accept": (C.c_int, [C.POINTER(PottsT), C.POINTER(ChainsT), C.POINTER(PasParamsT), vp]),
    "ppde_onehot_to_aa": (C.c_int, [vp, C.c_int32, C.c_int32, vp, C.c_int32, vp]),
    "ppde_aa_to_onehot": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, vp, vp]),
    "ppde_host_onehot_to_aa": (C.c_int, [vp, C.c_int64, C.c_int32, vp, C.c_int64, C.c_int32]),
    "ppde_host_aa_to_onehot": (C.c_int, [vp, C.c_int64, C.c_int64, C.c_int32, vp, C.c_int32]),
    "ppde_population_metrics": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp]),
    "ppde_counter_add": (C.c_int, [vp, C.c_int32, vp]),
    "ppde_quantiles": (C.c_int, [vp, C.c_int64, vp, C.c_int32, vp, vp]),
    "ppde_population_sums": (C.c_int, [vp, vp, C.c_int64, vp, vp]),
    "ppde_unique_count_table_entries": (C.c_int64, [C.c_int64]),
    "ppde_unique_count": (C.c_int, [vp, C.c_int64, C.c_int64, C.c_int32, vp, C.c_int64, vp, vp]),
    "ppde_topk": (C.c_int, [vp, C.c_int64, C.c_int32, C.c_int64, vp, vp, vp, vp]),
    "ppde_gather_rows": (C.c_int, [vp, C.c_int64, vp, C.c_int32, C.c_int64, vp, vp]),
    "ppde_pas_kat": (C.c_int, [vp, C.c_int32, vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp, vp]),
}

_lib = None


def load():
    """Load the shared library (once). Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C ppde_b200/csrc). ppde_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)           # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class CudaError(RuntimeError):
    pass


def check(code, what):
    if code != 0:
        raise CudaError(f"{what} failed with cudaError_t {code}")
