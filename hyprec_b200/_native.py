"""
ctypes binding of libhyprec_b200.so (include/hyprec_b200.h).  There is no CPU fallback: if the
library is missing, or no CUDA device is visible, creating a handle raises.
"""
import ctypes
import os

import numpy

from hyprec_b200 import build as _build

F64, F32 = 0, 1
SIDE_USER, SIDE_ITEM = 0, 1
T_GRAM, T_SOLVE, T_RMSE, T_SWEEP, T_TOPK, T_FLAGS, T_COUNT, T_PREP, T_COMM = range(9)

EXPORTS = [
    "hyprec_create", "hyprec_destroy", "hyprec_last_error", "hyprec_abi_version", "hyprec_launch_count",
    "hyprec_set_train_csr", "hyprec_set_shard", "hyprec_set_factors", "hyprec_get_factors",
    "hyprec_device_factors", "hyprec_set_item_prior", "hyprec_als_half_step", "hyprec_rmse", "hyprec_rmse_terms", "hyprec_train",
    "hyprec_predict_dense", "hyprec_score_dense_f32", "hyprec_set_eval_csr", "hyprec_set_candidates", "hyprec_set_hashed_candidates", "hyprec_eval_topk", "hyprec_eval_topk_u8", "hyprec_kernel_ms", "hyprec_reset_kernel_ms",
    "hyprec_flush_l2", "hyprec_probe_fma_tflops", "hyprec_solve_phase_cycles", "hyprec_host_alloc", "hyprec_host_free",
    "hyprec_comm_blob_bytes", "hyprec_comm_export", "hyprec_comm_attach", "hyprec_comm_detach", "hyprec_comm_stats",
    "hyprec_init_factors_uniform", "hyprec_get_factor_rows", "hyprec_mark", "hyprec_elapsed_ms",
    "hyprec_comm_arena", "hyprec_comm_attach_local", "hyprec_eval_stats", "hyprec_host_kfold_lists", "hyprec_host_count_nonzero", "hyprec_host_naive_picks", "hyprec_host_row_counts", "hyprec_host_dense_to_csr",
]

EPOCH_CB = ctypes.CFUNCTYPE(None, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_void_p)

_lib = None


class HyprecError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "hyprec_b200 error %d: %s" % (code, message))
        self.code = code


def load_library(rebuild=False):
    """Load (building first when stale and nvcc is present) the C-ABI library."""
    global _lib
    if _lib is not None and not rebuild:
        return _lib
    path = _build.LIB_PATH
    if rebuild or _build.is_stale():
        try:
            _build.build(force=rebuild)
        except Exception:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise RuntimeError("libhyprec_b200.so is missing (run `python -m hyprec_b200.build`); "
                           "hyprec_b200 has no CPU fallback")
    lib = ctypes.CDLL(path)
    c_i64, c_int, c_dbl, vp = ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_void_p
    lib.hyprec_create.argtypes = [ctypes.POINTER(vp), c_int, c_int]
    lib.hyprec_destroy.argtypes = [vp]
    lib.hyprec_last_error.restype = ctypes.c_char_p
    lib.hyprec_launch_count.argtypes = [vp]
    lib.hyprec_launch_count.restype = c_i64
    lib.hyprec_set_train_csr.argtypes = [vp, c_i64, c_i64, vp, vp, vp]
    lib.hyprec_set_shard.argtypes = [vp, c_i64, c_i64, c_i64, c_i64]
    lib.hyprec_set_factors.argtypes = [vp, vp, vp, c_int]
    lib.hyprec_get_factors.argtypes = [vp, vp, vp]
    lib.hyprec_device_factors.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(c_int)]
    lib.hyprec_set_item_prior.argtypes = [vp, vp]
    lib.hyprec_als_half_step.argtypes = [vp, c_int, c_dbl]
    lib.hyprec_rmse.argtypes = [vp, ctypes.POINTER(c_dbl)]
    lib.hyprec_rmse_terms.argtypes = [vp, ctypes.POINTER(c_dbl)]
    lib.hyprec_train.argtypes = [vp, c_int, c_dbl, c_dbl, EPOCH_CB, vp, ctypes.POINTER(c_int), ctypes.POINTER(c_dbl)]
    lib.hyprec_predict_dense.argtypes = [vp, vp]
    lib.hyprec_score_dense_f32.argtypes = [vp, vp]
    lib.hyprec_set_eval_csr.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.hyprec_set_candidates.argtypes = [vp, c_i64, c_i64, vp, vp]
    lib.hyprec_set_hashed_candidates.argtypes = [vp, c_int, c_int, ctypes.c_uint64]
    lib.hyprec_eval_topk.argtypes = [vp, c_i64, c_i64, vp, vp, c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.hyprec_eval_topk_u8.argtypes = [vp, c_i64, c_i64, vp, vp, c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.hyprec_kernel_ms.argtypes = [vp, c_int, ctypes.POINTER(c_dbl), ctypes.POINTER(c_i64), ctypes.POINTER(c_dbl)]
    lib.hyprec_reset_kernel_ms.argtypes = [vp]
    lib.hyprec_flush_l2.argtypes = [vp]
    lib.hyprec_host_naive_picks.argtypes = [vp, ctypes.POINTER(c_int), c_i64, vp, vp, vp]
    lib.hyprec_host_row_counts.argtypes = [vp, c_i64, c_i64, vp]
    lib.hyprec_host_dense_to_csr.argtypes = [vp, c_i64, c_i64, vp, vp, vp]
    lib.hyprec_host_count_nonzero.argtypes = [vp, c_i64]
    lib.hyprec_host_count_nonzero.restype = c_i64
    lib.hyprec_host_kfold_lists.argtypes = [vp, ctypes.POINTER(c_int), c_i64, c_i64, c_int, vp, vp, c_i64, vp, vp, vp, vp]
    lib.hyprec_host_alloc.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.hyprec_host_free.argtypes = [vp, vp]
    lib.hyprec_solve_phase_cycles.argtypes = [vp, vp, c_int]
    lib.hyprec_probe_fma_tflops.argtypes = [vp, c_int, ctypes.POINTER(c_dbl)]
    lib.hyprec_comm_export.argtypes = [vp, c_int, c_int, vp, ctypes.c_size_t]
    lib.hyprec_comm_attach.argtypes = [vp, vp, ctypes.c_size_t, vp, vp]
    lib.hyprec_comm_detach.argtypes = [vp]
    lib.hyprec_comm_arena.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_size_t)]
    lib.hyprec_comm_attach_local.argtypes = [vp, vp, vp, vp, vp]
    lib.hyprec_comm_stats.argtypes = [vp, vp]
    lib.hyprec_init_factors_uniform.argtypes = [vp, c_int, ctypes.c_uint64]
    lib.hyprec_get_factor_rows.argtypes = [vp, c_int, vp, c_i64, vp]
    lib.hyprec_eval_stats.argtypes = [vp, vp]
    lib.hyprec_mark.argtypes = [vp, c_int]
    lib.hyprec_elapsed_ms.argtypes = [vp, c_int, c_int, ctypes.POINTER(c_dbl)]
    _lib = lib
    return lib


def _ptr(array):
    return None if array is None else array.ctypes.data_as(ctypes.c_void_p)


def _as(array, dtype):
    if array is None:
        return None
    return numpy.ascontiguousarray(array, dtype=dtype)


def host_count_nonzero(block):
    """Non-zeros of a C-contiguous float64 array through the library (the call releases the GIL)."""
    return load_library().hyprec_host_count_nonzero(_ptr(block), int(block.size))


def host_dense_to_csr(matrix, pool):
    """(indptr int64, indices int32, values float64) of a C-contiguous float64 matrix: row blocks scanned by the
    library's host code on the threads of `pool` (the calls release the GIL)."""
    lib = load_library()
    rows, cols = matrix.shape
    step = max(1, -(-rows // 32))
    blocks = [(at, min(rows, at + step)) for at in range(0, rows, step)]
    counts = numpy.zeros(rows, dtype=numpy.int64)

    def count(block):
        return lib.hyprec_host_row_counts(_ptr(matrix[block[0]:block[1]]), block[1] - block[0], cols, _ptr(counts[block[0]:block[1]]))
    if any(code != 0 for code in pool.map(count, blocks)):
        raise HyprecError(-1, lib.hyprec_last_error().decode("utf-8", "replace"))
    indptr = numpy.zeros(rows + 1, dtype=numpy.int64)
    numpy.cumsum(counts, out=indptr[1:])
    indices = numpy.empty(int(indptr[-1]), dtype=numpy.int32)
    values = numpy.empty(int(indptr[-1]), dtype=numpy.float64)

    def fill(block):
        return lib.hyprec_host_dense_to_csr(_ptr(matrix[block[0]:block[1]]), block[1] - block[0], cols,
                                            _ptr(indptr[block[0]:block[1] + 1]), _ptr(indices), _ptr(values))
    if any(code != 0 for code in pool.map(fill, blocks)):
        raise HyprecError(-1, lib.hyprec_last_error().decode("utf-8", "replace"))
    return indptr, indices, values


def host_naive_picks(indptr, sizes):
    """CSR entry numbers Evaluator.naive_split_users samples (lib/evaluator.py:92-93: one numpy.random.choice with
    replacement per user), drawn from -- and advancing -- numpy's global legacy stream.  None when that stream is not
    MT19937."""
    lib = load_library()
    state = numpy.random.get_state()
    if state[0] != 'MT19937':
        return None
    key = numpy.ascontiguousarray(state[1], dtype=numpy.uint32).copy()
    pos = ctypes.c_int(int(state[2]))
    indptr = numpy.ascontiguousarray(indptr, dtype=numpy.int64)
    sizes = numpy.ascontiguousarray(sizes, dtype=numpy.int64)
    picks = numpy.empty(int(sizes.sum()), dtype=numpy.int64)
    code = lib.hyprec_host_naive_picks(_ptr(key), ctypes.byref(pos), len(indptr) - 1, _ptr(indptr), _ptr(sizes), _ptr(picks))
    if code != 0:
        raise HyprecError(code, lib.hyprec_last_error().decode("utf-8", "replace"))
    numpy.random.set_state((state[0], key, int(pos.value), state[3], state[4]))
    return picks


def host_kfold_lists(indptr, indices, n_users, n_items, k_folds):
    """Evaluator.get_kfold_indices' candidate lists (lib/evaluator.py:136-195) from the library's host code, drawn from
    -- and advancing -- numpy's GLOBAL legacy random stream exactly as the reference's per-user shuffles do.
    Returns (slot_ptr int64[slots + 1], slot_positives int64[slots], ids int32), or None when a user's lists need the
    reference's negative slice bounds (nothing was drawn then; the caller runs the reference's loop).  The ids are
    int32 (the reference's lists are int64 arrays; at 5 551 x 16 980 that is 754 MB of fresh pages for nothing)."""
    lib = load_library()
    state = numpy.random.get_state()
    if state[0] != 'MT19937':
        return None
    key = numpy.ascontiguousarray(state[1], dtype=numpy.uint32).copy()
    pos = ctypes.c_int(int(state[2]))
    indptr = numpy.ascontiguousarray(indptr, dtype=numpy.int64)
    indices = numpy.ascontiguousarray(indices, dtype=numpy.int32)
    slots = int(n_users) * int(k_folds)
    capacity = slots * (int(n_items) // int(k_folds))
    slot_ptr = numpy.empty(slots + 1, dtype=numpy.int64)
    slot_positives = numpy.empty(slots, dtype=numpy.int64)
    ids32 = numpy.empty(capacity, dtype=numpy.int32)
    code = lib.hyprec_host_kfold_lists(_ptr(key), ctypes.byref(pos), int(n_users), int(n_items), int(k_folds),
                                       _ptr(indptr), _ptr(indices), capacity, _ptr(slot_ptr), _ptr(slot_positives),
                                       None, _ptr(ids32))
    if code == -4:          # HYPREC_E_UNSUPPORTED
        return None
    if code != 0:
        raise HyprecError(code, lib.hyprec_last_error().decode("utf-8", "replace"))
    numpy.random.set_state((state[0], key, int(pos.value), state[3], state[4]))
    return slot_ptr, slot_positives, ids32


class Handle(object):
    """One GPU's worth of device state behind the C ABI."""

    def __init__(self, device=0, precision=F64):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        self.precision = precision
        self.m = self.n = self.k = 0
        self._pinned = []
        self._check(self._lib.hyprec_create(ctypes.byref(self._h), int(device), int(precision)))

    def _check(self, code):
        if code != 0:
            raise HyprecError(code, self._lib.hyprec_last_error().decode("utf-8", "replace"))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            for ptr in getattr(self, "_pinned", []):
                self._lib.hyprec_host_free(self._h, ptr)
            self._pinned = []
            self._lib.hyprec_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        return int(self._lib.hyprec_launch_count(self._h))

    # ---- data
    def set_train_csr(self, m, n, indptr, indices, values=None):
        indptr, indices, values = _as(indptr, numpy.int64), _as(indices, numpy.int32), _as(values, numpy.float64)
        self._check(self._lib.hyprec_set_train_csr(self._h, m, n, _ptr(indptr), _ptr(indices), _ptr(values)))
        self.m, self.n = int(m), int(n)

    def set_eval_csr(self, full, test):
        """full / test: (indptr, indices, values-or-None) triples."""
        arrs = []
        for indptr, indices, values in (full, test):
            arrs += [_as(indptr, numpy.int64), _as(indices, numpy.int32), _as(values, numpy.float64)]
        self._check(self._lib.hyprec_set_eval_csr(self._h, *[_ptr(a) for a in arrs]))

    def set_shard(self, user_lo, user_hi, item_lo, item_hi):
        self._check(self._lib.hyprec_set_shard(self._h, user_lo, user_hi, item_lo, item_hi))

    def set_factors(self, user_vecs, item_vecs):
        u, v = _as(user_vecs, numpy.float64), _as(item_vecs, numpy.float64)
        if u.shape[0] != self.m or v.shape[0] != self.n or u.shape[1] != v.shape[1]:
            raise ValueError("factor shapes %s / %s do not match a %d x %d problem" % (u.shape, v.shape, self.m, self.n))
        self.k = int(u.shape[1])
        self._check(self._lib.hyprec_set_factors(self._h, _ptr(u), _ptr(v), self.k))

    def init_factors_uniform(self, k, seed):
        """U, V ~ uniform [0, 1) drawn on the device (oracle.als.uniform_factors restates the generator)."""
        self.k = int(k)
        self._check(self._lib.hyprec_init_factors_uniform(self._h, self.k, int(seed) & (2 ** 64 - 1)))

    def get_factor_rows(self, side, rows):
        rows = _as(rows, numpy.int64)
        out = numpy.empty((rows.shape[0], self.k))
        self._check(self._lib.hyprec_get_factor_rows(self._h, int(side), _ptr(rows), rows.shape[0], _ptr(out)))
        return out

    # ---- multi-GPU exchange (one process per GPU; see include/hyprec_b200.h)
    def comm_export(self, rank, world):
        size = int(self._lib.hyprec_comm_blob_bytes())
        blob = ctypes.create_string_buffer(size)
        self._check(self._lib.hyprec_comm_export(self._h, int(rank), int(world), blob, size))
        return blob.raw

    def comm_attach(self, blobs, user_bounds, item_bounds):
        size = int(self._lib.hyprec_comm_blob_bytes())
        joined = b"".join(blobs)
        if len(joined) != size * len(blobs):
            raise ValueError("every rank publishes one blob of %d bytes" % size)
        ub, ib = _as(user_bounds, numpy.int64), _as(item_bounds, numpy.int64)
        self._check(self._lib.hyprec_comm_attach(self._h, joined, size, _ptr(ub), _ptr(ib)))

    def comm_arena(self):
        base, size = ctypes.c_void_p(), ctypes.c_size_t()
        self._check(self._lib.hyprec_comm_arena(self._h, ctypes.byref(base), ctypes.byref(size)))
        return base.value, size.value

    def comm_attach_local(self, arenas, devices, user_bounds, item_bounds):
        """Peers in this process: ``arenas`` = every handle's comm_arena()[0] in rank order."""
        ptrs = (ctypes.c_void_p * len(arenas))(*arenas)
        devs = numpy.ascontiguousarray(devices, dtype=numpy.int32)
        ub, ib = _as(user_bounds, numpy.int64), _as(item_bounds, numpy.int64)
        self._check(self._lib.hyprec_comm_attach_local(self._h, ptrs, _ptr(devs), _ptr(ub), _ptr(ib)))

    def comm_detach(self):
        self._check(self._lib.hyprec_comm_detach(self._h))

    def comm_stats(self):
        out = numpy.zeros(8)
        self._check(self._lib.hyprec_comm_stats(self._h, _ptr(out)))
        return dict(attached=bool(out[0]), bytes_pushed=float(out[1]), barriers=int(out[2]), arena_bytes=int(out[3]))

    def mark(self, slot):
        self._check(self._lib.hyprec_mark(self._h, int(slot)))

    def elapsed_ms(self, start, stop):
        out = ctypes.c_double()
        self._check(self._lib.hyprec_elapsed_ms(self._h, int(start), int(stop), ctypes.byref(out)))
        return out.value

    def get_factors(self, user_out=None, item_out=None):
        """Copy the factors into the given float64 C-contiguous arrays (allocated when None)."""
        if user_out is None:
            user_out = numpy.empty((self.m, self.k))
        if item_out is None:
            item_out = numpy.empty((self.n, self.k))
        for arr in (user_out, item_out):
            if arr is not False and not (arr.dtype == numpy.float64 and arr.flags.c_contiguous):
                raise ValueError("factor outputs must be C-contiguous float64")
        self._check(self._lib.hyprec_get_factors(self._h, _ptr(user_out) if user_out is not False else None,
                                                 _ptr(item_out) if item_out is not False else None))
        return user_out, item_out

    def device_factors(self):
        u, v, e = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int()
        self._check(self._lib.hyprec_device_factors(self._h, ctypes.byref(u), ctypes.byref(v), ctypes.byref(e)))
        return u.value, v.value, e.value

    def set_item_prior(self, theta):
        theta = _as(theta, numpy.float64)
        self._check(self._lib.hyprec_set_item_prior(self._h, _ptr(theta)))

    # ---- compute
    def als_half_step(self, side, lam):
        self._check(self._lib.hyprec_als_half_step(self._h, side, float(lam)))

    def rmse(self):
        out = ctypes.c_double()
        self._check(self._lib.hyprec_rmse(self._h, ctypes.byref(out)))
        return out.value

    def rmse_terms(self):
        """(<U^T U, V^T V>, sum_nz r u.v, sum_nz r^2); sparse sums over this handle's user shard."""
        out = (ctypes.c_double * 3)()
        self._check(self._lib.hyprec_rmse_terms(self._h, out))
        return out[0], out[1], out[2]

    def train(self, n_iter, lam, start_error=numpy.inf, callback=None):
        """Returns (epochs_run, last_rmse); callback(epoch, rmse, seconds)."""
        cb = EPOCH_CB(lambda e, r, s, _u: callback(e, r, s)) if callback else ctypes.cast(None, EPOCH_CB)
        epochs, last = ctypes.c_int(), ctypes.c_double()
        self._check(self._lib.hyprec_train(self._h, int(n_iter), float(lam), float(start_error), cb, None,
                                           ctypes.byref(epochs), ctypes.byref(last)))
        return epochs.value, last.value

    def predict_dense(self):
        out = numpy.empty((self.m, self.n))
        self._check(self._lib.hyprec_predict_dense(self._h, _ptr(out)))
        return out

    def set_candidates(self, cand_indptr, cand_ids, user_lo=0, user_hi=None):
        """Keep a fold's candidate streams on the device (None, None drops them)."""
        user_hi = self.m if user_hi is None else user_hi
        cand_indptr, cand_ids = _as(cand_indptr, numpy.int64), _as(cand_ids, numpy.int32)
        self._check(self._lib.hyprec_set_candidates(self._h, user_lo, user_hi, _ptr(cand_indptr), _ptr(cand_ids)))

    def set_hashed_candidates(self, fold, n_folds, seed=0):
        """Device-generated candidate lists: test positives + the hashed 1/n_folds share of the unrated items
        (include/hyprec_b200.h); n_folds <= 0 switches back to explicit lists / all items."""
        self._check(self._lib.hyprec_set_hashed_candidates(self._h, int(fold), int(n_folds), int(seed) & (2 ** 64 - 1)))

    def score_dense_f32(self):
        """Tensor-core (tcgen05, 3xTF32) approximation of U.V^T as float32."""
        out = numpy.empty((self.m, self.n), dtype=numpy.float32)
        self._check(self._lib.hyprec_score_dense_f32(self._h, _ptr(out)))
        return out

    def eval_topk(self, capacity, cand_indptr=None, cand_ids=None, user_lo=0, user_hi=None, want_counts=True,
                  want_topk=True, n_train_nz=None, n_test_nz=None, out=None, compact=False):
        """Returns dict(thresholds, ones_per_row, topk_idx, topk_val, topk_hit, train_flags, test_flags).
        `out`: a dict returned by an earlier identical call, to reuse its (possibly pinned) arrays.
        `compact`: hit values as uint8 and no score array (hyprec_eval_topk_u8; a tenth of the device -> host bytes);
        a reused `out` whose topk_hit is uint8 selects the same call."""
        user_hi = self.m if user_hi is None else user_hi
        nu = user_hi - user_lo
        cand_indptr, cand_ids = _as(cand_indptr, numpy.int64), _as(cand_ids, numpy.int32)
        if out is None:
            out = dict(thresholds=numpy.empty(nu))
            out["ones_per_row"] = numpy.zeros(nu, dtype=numpy.int64) if want_counts else None
            out["topk_idx"] = out["topk_val"] = out["topk_hit"] = None
            if want_topk:
                out["topk_idx"] = numpy.empty((nu, capacity), dtype=numpy.int32)
                if compact:
                    out["topk_hit"] = numpy.empty((nu, capacity), dtype=numpy.uint8)
                else:
                    out["topk_val"] = numpy.empty((nu, capacity))
                    out["topk_hit"] = numpy.empty((nu, capacity))
            out["train_flags"] = numpy.zeros(n_train_nz, dtype=numpy.uint8) if n_train_nz is not None else None
            out["test_flags"] = numpy.zeros(n_test_nz, dtype=numpy.uint8) if n_test_nz is not None else None
        hit = out["topk_hit"]
        call = self._lib.hyprec_eval_topk_u8 if hit is not None and hit.dtype == numpy.uint8 else self._lib.hyprec_eval_topk
        self._check(call(
            self._h, user_lo, user_hi, _ptr(cand_indptr), _ptr(cand_ids), int(capacity), _ptr(out["thresholds"]),
            _ptr(out["ones_per_row"]), _ptr(out["topk_idx"]), _ptr(out["topk_val"]), _ptr(hit),
            _ptr(out["train_flags"]), _ptr(out["test_flags"])))
        return out

    def eval_stats(self):
        out = numpy.zeros(8)
        self._check(self._lib.hyprec_eval_stats(self._h, _ptr(out)))
        return dict(fused=bool(out[0]), fallback_users=int(out[1]), rechecked_cells=int(out[2]),
                    survivor_buffer_bytes=int(out[3]), dense_score_bytes=int(out[4]))

    def kernel_ms(self, which):
        """(total ms, timed regions, ms of the latest API call) of one kernel class."""
        total, calls, last = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
        self._check(self._lib.hyprec_kernel_ms(self._h, which, ctypes.byref(total), ctypes.byref(calls),
                                               ctypes.byref(last)))
        return total.value, calls.value, last.value

    def reset_kernel_ms(self):
        self._check(self._lib.hyprec_reset_kernel_ms(self._h))

    def probe_fma_tflops(self, precision):
        out = ctypes.c_double()
        self._check(self._lib.hyprec_probe_fma_tflops(self._h, int(precision), ctypes.byref(out)))
        return out.value

    def solve_phase_cycles(self, reset=True):
        out = numpy.zeros(16, dtype=numpy.int64)
        self._check(self._lib.hyprec_solve_phase_cycles(self._h, _ptr(out), 1 if reset else 0))
        return out

    def pinned_empty(self, shape, dtype=numpy.float64):
        """numpy array backed by page-locked host memory (freed by close())."""
        dtype = numpy.dtype(dtype)
        count = int(numpy.prod(shape))
        ptr = ctypes.c_void_p()
        self._check(self._lib.hyprec_host_alloc(self._h, max(count * dtype.itemsize, 1), ctypes.byref(ptr)))
        self._pinned.append(ptr)
        buf = (ctypes.c_char * (count * dtype.itemsize)).from_address(ptr.value)
        return numpy.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def flush_l2(self):
        self._check(self._lib.hyprec_flush_l2(self._h))
