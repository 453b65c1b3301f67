"""ctypes binding of include/spamm_b200.h.  There is NO CPU fallback: if the
CUDA library is missing this raises, and every product entry point goes
through it."""
import ctypes
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPAMM_B200_LIB", os.path.join(PKG, "libspamm_b200.so"))

ABI_VERSION = 4
MAX_COMPONENTS = 8
MAX_TEMPLATES = 16
SH95_NE, SH95_T, SH95_LINES = 7, 10, 48

COMP_NUCLEAR, COMP_FE, COMP_HOST, COMP_BALMER, COMP_EXTINCTION = 0, 1, 2, 3, 4
LINE_LORENTZIAN, LINE_GAUSSIAN = 0, 1
TEMPLATES_AUTO, TEMPLATES_BANK, TEMPLATES_DIRECT = 0, 1, 2

c_double_p = ctypes.POINTER(ctypes.c_double)


class SpammTemplateSet(ctypes.Structure):
    _fields_ = [
        ("n_templates", ctypes.c_int32),
        ("n_points", ctypes.c_int32 * MAX_TEMPLATES),
        ("log_wl", c_double_p),
        ("log_flux", c_double_p),
        ("norm_wl", ctypes.c_double * MAX_TEMPLATES),
        ("template_width", ctypes.c_double),
        ("sigma_denom", ctypes.c_double),
        ("sqrt_2pi", ctypes.c_double),
        ("kernel_size_sigma", ctypes.c_int32),
        ("clamp_edges", ctypes.c_int32),
        ("rebin_indptr", ctypes.POINTER(ctypes.c_int32)),
        ("rebin_index", ctypes.POINTER(ctypes.c_int32)),
        ("rebin_weight", c_double_p),
    ]


class SpammPlanDesc(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32),
        ("n_pix", ctypes.c_int32),
        ("wl", c_double_p),
        ("ln_wl", c_double_p),
        ("flux", c_double_p),
        ("flux_error", c_double_p),
        ("sum_ln_2pi_err2", ctypes.c_double),
        ("n_components", ctypes.c_int32),
        ("component_kind", ctypes.c_int32 * MAX_COMPONENTS),
        ("n_dim", ctypes.c_int32),
        ("prior_lo", c_double_p),
        ("prior_hi", c_double_p),
        ("nc_broken_pl", ctypes.c_int32),
        ("nc_norm_wl", ctypes.c_double),
        ("fe", SpammTemplateSet),
        ("host", SpammTemplateSet),
        ("template_mode", ctypes.c_int32),
        ("balmer_bc", ctypes.c_int32),
        ("balmer_bpc", ctypes.c_int32),
        ("balmer_line_type", ctypes.c_int32),
        ("balmer_line_min", ctypes.c_int32),
        ("balmer_n_lines", ctypes.c_int32),
        ("balmer_exact_line_sum", ctypes.c_int32),
        ("balmer_line_center", c_double_p),
        ("balmer_line_exparg", c_double_p),
        ("sh95_ne", ctypes.c_double * SH95_NE),
        ("sh95_T", ctypes.c_double * SH95_T),
        ("sh95_z", c_double_p),
        ("bc_ln_wl_new", c_double_p),
        ("bc_exp_ln_wl_new", c_double_p),
        ("c_kms", ctypes.c_double),
        ("c_cgs", ctypes.c_double),
        ("c_aa", ctypes.c_double),
        ("h_cgs", ctypes.c_double),
        ("kb_cgs", ctypes.c_double),
        ("balmer_edge", ctypes.c_double),
        ("max_walkers", ctypes.c_int32),
        ("ext_k", c_double_p),
        ("bc_rebin_indptr", ctypes.POINTER(ctypes.c_int32)),
        ("bc_rebin_index", ctypes.POINTER(ctypes.c_int32)),
        ("bc_rebin_weight", c_double_p),
    ]


# every symbol include/spamm_b200.h declares
EXPORTS = [
    "spamm_plan_create", "spamm_plan_destroy", "spamm_lnpost_batch", "spamm_lnpost_batch_host",
    "spamm_model_flux_batch", "spamm_likelihood_batch", "spamm_plan_n_dim", "spamm_plan_n_pix", "spamm_plan_template_mode",
    "spamm_plan_bank_rows", "spamm_plan_status", "spamm_plan_launch_count", "spamm_plan_set_walker_split", "spamm_plan_specialised",
    "spamm_stretch_propose", "spamm_stretch_accept", "spamm_chain_record", "spamm_stretch_run",
    "spamm_fp64_peak_probe", "spamm_broadening_probe", "spamm_last_error", "spamm_abi_version", "spamm_plan_check",
    "spamm_shard_bounds", "spamm_exchange_region_bytes", "spamm_exchange_create", "spamm_exchange_connect",
    "spamm_exchange_connect_ptrs", "spamm_exchange_region_ptr", "spamm_exchange_epoch", "spamm_exchange_destroy",
    "spamm_lnpost_batch_sharded", "spamm_lnpost_batch_sharded_host", "spamm_stretch_run_sharded",
]
IPC_HANDLE_BYTES = 64

_lib = None


class SpammError(RuntimeError):
    pass


def load():
    """Load libspamm_b200.so (building is the job of __graft_entry__.build /
    spamm_b200._build); raises if it is absent -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if "SPAMM_B200_LIB" not in os.environ and "SPAMM_B200_NO_AUTOBUILD" not in os.environ:
        # in-tree library missing or older than its sources: compile it (nvcc, sm_100a) -- still no
        # fallback of any other kind: if that fails, so does the import of the product path
        from . import _build
        try:
            if _build.needs_build():
                _build.build()
        except Exception as exc:
            if not os.path.exists(LIB_PATH):
                raise SpammError("CUDA library %s is missing and could not be built: %s" % (LIB_PATH, exc))
    if not os.path.exists(LIB_PATH):
        raise SpammError(
            "CUDA library %s is missing: run `python -m spamm_b200._build` (needs nvcc). "
            "spamm_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64, dbl = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32,
                                   ctypes.c_uint64, ctypes.c_double)
    lib.spamm_plan_create.argtypes = [ctypes.POINTER(SpammPlanDesc), ctypes.POINTER(vp)]
    lib.spamm_plan_create.restype = ctypes.c_int
    lib.spamm_plan_destroy.argtypes = [vp]
    lib.spamm_plan_destroy.restype = None
    lib.spamm_lnpost_batch.argtypes = [vp, vp, i32, vp, vp]
    lib.spamm_lnpost_batch.restype = ctypes.c_int
    lib.spamm_lnpost_batch_host.argtypes = [vp, vp, i32, vp]
    lib.spamm_lnpost_batch_host.restype = ctypes.c_int
    lib.spamm_model_flux_batch.argtypes = [vp, u32, vp, i32, vp, vp]
    lib.spamm_model_flux_batch.restype = ctypes.c_int
    lib.spamm_likelihood_batch.argtypes = [vp, vp, i32, vp, vp]
    lib.spamm_likelihood_batch.restype = ctypes.c_int
    for name in ("spamm_plan_n_dim", "spamm_plan_n_pix", "spamm_plan_template_mode",
                 "spamm_plan_bank_rows", "spamm_plan_specialised"):
        getattr(lib, name).argtypes = [vp]
        getattr(lib, name).restype = ctypes.c_int
    lib.spamm_plan_status.argtypes = [vp, vp]
    lib.spamm_plan_status.restype = ctypes.c_int
    lib.spamm_plan_launch_count.argtypes = [vp]
    lib.spamm_plan_launch_count.restype = i64
    lib.spamm_plan_set_walker_split.argtypes = [vp, i32]
    lib.spamm_plan_set_walker_split.restype = ctypes.c_int
    lib.spamm_stretch_propose.argtypes = [vp, i32, i32, i32, dbl, u64, u64, vp, vp, vp]
    lib.spamm_stretch_propose.restype = ctypes.c_int
    lib.spamm_stretch_accept.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, u64, u64, vp, vp]
    lib.spamm_stretch_accept.restype = ctypes.c_int
    lib.spamm_chain_record.argtypes = [vp, vp, i32, i32, i64, i64, vp, vp, vp]
    lib.spamm_chain_record.restype = ctypes.c_int
    lib.spamm_stretch_run.argtypes = [vp, vp, vp, i32, dbl, u64, u64, i32, vp, vp, vp, vp, vp, vp, i64, vp]
    lib.spamm_stretch_run.restype = ctypes.c_int
    lib.spamm_plan_check.argtypes = [vp, vp]
    lib.spamm_plan_check.restype = ctypes.c_int
    lib.spamm_shard_bounds.argtypes = [i64, i32, i32, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.spamm_shard_bounds.restype = None
    lib.spamm_exchange_region_bytes.argtypes = [i64]
    lib.spamm_exchange_region_bytes.restype = i64
    lib.spamm_exchange_create.argtypes = [i32, i32, i64, ctypes.POINTER(vp), vp]
    lib.spamm_exchange_create.restype = ctypes.c_int
    lib.spamm_exchange_connect.argtypes = [vp, vp]
    lib.spamm_exchange_connect.restype = ctypes.c_int
    lib.spamm_exchange_connect_ptrs.argtypes = [vp, ctypes.POINTER(u64)]
    lib.spamm_exchange_connect_ptrs.restype = ctypes.c_int
    lib.spamm_exchange_region_ptr.argtypes = [vp]
    lib.spamm_exchange_region_ptr.restype = u64
    lib.spamm_exchange_epoch.argtypes = [vp]
    lib.spamm_exchange_epoch.restype = u64
    lib.spamm_exchange_destroy.argtypes = [vp]
    lib.spamm_exchange_destroy.restype = None
    lib.spamm_lnpost_batch_sharded.argtypes = [vp, vp, vp, i32, vp, ctypes.POINTER(vp), vp]
    lib.spamm_lnpost_batch_sharded.restype = ctypes.c_int
    lib.spamm_lnpost_batch_sharded_host.argtypes = [vp, vp, vp, i32, vp]
    lib.spamm_lnpost_batch_sharded_host.restype = ctypes.c_int
    lib.spamm_stretch_run_sharded.argtypes = [vp, vp, vp, vp, i32, dbl, u64, u64, i32, vp, vp, vp, vp, vp, i64, vp]
    lib.spamm_stretch_run_sharded.restype = ctypes.c_int
    lib.spamm_broadening_probe.argtypes = [vp, i32, i32, i32, i32, ctypes.POINTER(dbl), ctypes.POINTER(dbl),
                                           ctypes.POINTER(dbl)]
    lib.spamm_broadening_probe.restype = ctypes.c_int
    lib.spamm_fp64_peak_probe.argtypes = [i32, i32, ctypes.POINTER(dbl)]
    lib.spamm_fp64_peak_probe.restype = ctypes.c_int
    lib.spamm_last_error.argtypes = []
    lib.spamm_last_error.restype = ctypes.c_char_p
    lib.spamm_abi_version.argtypes = []
    lib.spamm_abi_version.restype = ctypes.c_int
    if lib.spamm_abi_version() != ABI_VERSION:
        raise SpammError("libspamm_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().spamm_last_error()
        raise SpammError("spamm_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def as_c_i32(arr):
    """int32 C-contiguous copy + its int32_t* (keep the array alive!)."""
    a = np.ascontiguousarray(arr, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def as_c(arr):
    """float64 C-contiguous copy + its double* (keep the array alive!)."""
    a = np.ascontiguousarray(arr, dtype=np.float64)
    return a, a.ctypes.data_as(c_double_p)
