"""ctypes binding of libvireo_b200.so (see include/vireo_b200.h).

There is no CPU fallback: if the shared library is missing it is built with nvcc; if that fails, or
no CUDA device is visible when a compute entry point is called, an exception is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# VIREO_B200_LIB: another build of the same sources (the sanitizer variants of build.py), never a different backend
LIB_PATH = os.environ.get("VIREO_B200_LIB") or os.path.join(HERE, "libvireo_b200.so")

VB_I32, VB_I64, VB_F32, VB_F64 = 0, 1, 2, 3
PH_SNP, PH_THETA, PH_GT, PH_ID, PH_ELBO, PH_LOGLIK, PH_THETA_SUMS = 1, 2, 4, 8, 16, 32, 64
MAX_GT, MAX_DONOR = 8, 256
CTRL_N, SCAL_N = 8, 8

c_dp = C.c_void_p  # device pointer


class WsSizes(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("S", "W", "loglik", "ab", "part", "scal", "ctrl", "rpad", "heavy")]


class VireoArgs(C.Structure):
    _fields_ = [
        ("n_donor", C.c_int32), ("n_gt", C.c_int32), ("n_batch", C.c_int32),
        ("ase_mode", C.c_int32), ("learn_gt", C.c_int32), ("learn_theta", C.c_int32), ("fix_beta_sum", C.c_int32),
        ("id_prior_rows", C.c_int32), ("theta_prior_rows", C.c_int32),
        ("max_iter", C.c_int32), ("min_iter", C.c_int32), ("delay_fit_theta", C.c_int32),
        ("poll_every", C.c_int32), ("reserved", C.c_int32),
        ("epsilon_conv", C.c_double),
        ("id_prob", c_dp), ("gt_prob", c_dp), ("beta_mu", c_dp), ("beta_sum", c_dp),
        ("log_id_prior", c_dp), ("log_id_prior_kl", c_dp), ("log_gt_prior", c_dp), ("log_gt_prior_kl", c_dp),
        ("s1_prior", c_dp), ("s2_prior", c_dp),
        ("S1", c_dp), ("S2", c_dp), ("W", c_dp), ("loglik", c_dp), ("ab", c_dp), ("part", c_dp),
        ("scal", c_dp), ("ctrl", c_dp), ("elbo", c_dp), ("rpad", c_dp), ("heavy", c_dp),
        ("ws", WsSizes),
    ]


class DoubletWs(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("W", "heavy", "ab2")]


class BmmArgs(C.Structure):
    _fields_ = [
        ("n_donor", C.c_int32), ("n_batch", C.c_int32), ("fix_beta_sum", C.c_int32), ("id_prior_rows", C.c_int32),
        ("max_iter", C.c_int32), ("min_iter", C.c_int32), ("poll_every", C.c_int32), ("reserved", C.c_int32),
        ("epsilon_conv", C.c_double),
        ("id_prob", c_dp), ("beta_mu", c_dp), ("beta_sum", c_dp),
        ("log_id_prior", c_dp), ("log_id_prior_kl", c_dp), ("s1_prior", c_dp), ("s2_prior", c_dp),
        ("S1", c_dp), ("S2", c_dp), ("W", c_dp), ("loglik", c_dp), ("part", c_dp),
        ("scal", c_dp), ("ctrl", c_dp), ("elbo", c_dp), ("rpad", c_dp), ("heavy", c_dp),
        ("ws", WsSizes),
    ]


COMM_ID_BYTES = 128

# every symbol include/vireo_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "vb_counts_create": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                   C.c_void_p, C.POINTER(C.c_void_p)]),
    "vb_counts_destroy": (None, [C.c_void_p]),
    "vb_counts_slice": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p)]),
    "vb_counts_info": (C.c_int64, [C.c_void_p, C.c_int]),
    "vb_counts_note": (C.c_char_p, [C.c_void_p]),
    "vb_seg_verify": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64)]),
    "vb_binom_const": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    "vb_vireo_ws_sizes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(WsSizes)]),
    "vb_log_prior": (C.c_int, [c_dp, C.c_int64, C.c_int, c_dp, c_dp, C.c_void_p]),
    "vb_vireo_fit": (C.c_int, [C.c_void_p, C.POINTER(VireoArgs), C.c_void_p]),
    "vb_vireo_step": (C.c_int, [C.c_void_p, C.POINTER(VireoArgs), C.c_int, C.c_void_p]),
    "vb_bmm_ws_sizes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(WsSizes)]),
    "vb_bmm_fit": (C.c_int, [C.c_void_p, C.POINTER(BmmArgs), C.c_void_p]),
    "vb_bmm_step": (C.c_int, [C.c_void_p, C.POINTER(BmmArgs), C.c_int, C.c_void_p]),
    "vb_doublet_ws_sizes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(DoubletWs)]),
    "vb_vireo_doublet": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, C.c_int,
                                   c_dp, c_dp, c_dp, C.POINTER(DoubletWs), c_dp, c_dp, c_dp, C.c_void_p]),
    "vb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "vb_comm_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "vb_comm_destroy": (None, [C.c_void_p]),
    "vb_comm_allreduce": (C.c_int, [C.c_void_p, c_dp, C.c_int64, C.c_void_p]),
    "vb_comm_broadcast": (C.c_int, [C.c_void_p, c_dp, C.c_int64, C.c_int, C.c_void_p]),
    "vb_comm_allgather": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int64, C.c_void_p]),
    "vb_vireo_fit_sharded": (C.c_int, [C.c_void_p, C.POINTER(VireoArgs), C.c_void_p, c_dp, C.c_void_p]),
    "vb_vireo_gt_sharded": (C.c_int, [C.c_void_p, C.POINTER(VireoArgs), C.c_void_p, c_dp, C.c_void_p]),
    "vb_set_path": (None, [C.c_int]),
    "vb_set_graphs": (None, [C.c_int]),
    "vb_set_fuse": (None, [C.c_int]),
    "vb_launch_counts": (None, [C.POINTER(C.c_int64)]),
    "vb_profile_enable": (None, [C.c_int]),
    "vb_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "vb_last_error": (C.c_char_p, []),
    "vb_version": (C.c_char_p, []),
    # host-callable scalar math used by the CPU test-suite
    "vb_host_digamma": (C.c_double, [C.c_double]),
    "vb_host_beta_kl": (C.c_double, [C.c_double, C.c_double, C.c_double, C.c_double]),
    "vb_host_binom_term": (C.c_float, [C.c_uint32, C.c_uint32]),
    "vb_host_seg_count_code": (C.c_uint32, [C.c_uint32, C.c_int]),
}

_lib = None
PATHS = {"auto": 0, "rows": 1, "seg": 3, "seg32": 4}


class VireoB200Error(RuntimeError):
    pass


def load():
    """Load (building first if needed) libvireo_b200.so.  Raises if it cannot be produced."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    env = os.environ.get("VIREO_B200_PATH", "auto").lower()
    lib.vb_set_path(PATHS.get(env, 0))
    lib.vb_set_graphs(0 if os.environ.get("VIREO_B200_GRAPHS", "1") == "0" else 1)
    return lib


def set_path(mode):
    """'auto' | 'rows' | 'seg' | 'seg32': kernel family of the two sparse passes (see vb_set_path)."""
    load().vb_set_path(PATHS[mode])


KERNEL_CLASSES = ("k_snp", "k_theta", "k_gt", "k_cell", "k_elbo", "k_bmm_theta", "k_terms", "helpers")


def launch_counts():
    arr = (C.c_int64 * 8)()
    load().vb_launch_counts(arr)
    return dict(zip(KERNEL_CLASSES, [int(x) for x in arr]))


def profile_read():
    """{class: (total ms, launches)} of the launches recorded since vb_profile_enable(1)."""
    ms, n = (C.c_double * 8)(), (C.c_int64 * 8)()
    load().vb_profile_read(ms, n)
    return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(KERNEL_CLASSES)}


def check(rc):
    if rc != 0:
        msg = load().vb_last_error()
        raise VireoB200Error("libvireo_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
