"""ctypes view of libqpde_b200.so (include/qpde_b200.h).

This is plumbing for tests/ and bench.py: it declares the C structs, loads the
in-tree shared library and turns error codes into exceptions.  There is no
fallback of any kind — if the library is missing or no B200 is present the
calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libqpde_b200.so")

ABI_VERSION = 6

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_NOT_CONVERGED, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5

SCHEME = {"rannacher": 0, "cn": 1, "bdf1": 2, "bdf2": 3, "bdf3": 4, "bdf4": 5, "bdf5": 6, "bdf6": 7}
PAYOFF = {"array": 0, "put": 1, "call": 2, "digital_put": 3, "digital_call": 4,
          "straddle": 5, "k_minus_s": 6, "none": -1}
DIVIDEND = {None: 0, "none": 0, "cash": 1, "proportional": 2}
VOL = {"const": 0, "inverse_s": 1, "nodes": 2}
KERNEL = {"auto": 0, "interleaved": 1, "resident": 2}
# event program op codes (QPDE_EV_*)
EV = {"end": 0, "S": 1, "const": 2, "param": 3, "add": 4, "sub": 5, "mul": 6, "div": 7, "max": 8, "min": 9,
      "neg": 10, "V": 11, "gt": 12, "select": 13}


def event_program(*tokens):
    """Assembles an event program from tokens: op names ("S", "add", "V", ...), ("const", i) and ("param", i)."""
    words = []
    for t in tokens:
        if isinstance(t, tuple):
            words += [EV[t[0]], int(t[1])]
        else:
            words.append(EV[t])
    if not words or words[-1] != EV["end"]:
        words.append(EV["end"])
    return np.array(words, dtype=np.int32)
KERNEL_NAME = {v: k for k, v in KERNEL.items()}

OUT_V, OUT_S, OUT_VALUE, OUT_STEPS, OUT_INNER_TOTAL, OUT_INNER, OUT_ACTIVE, OUT_STATUS = (
    1, 2, 4, 8, 16, 32, 64, 128)
OUT_COMPUTED = 256
OUT_ALL = 511

_dp = C.POINTER(C.c_double)


class Step(C.Structure):
    _fields_ = [("t", C.c_double), ("dt", C.c_double), ("same_dt", C.c_int32),
                ("event_after", C.c_int32), ("user_events", C.c_int32),
                ("reserved", C.c_int32)]


class Batch(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_contracts", C.c_int32),
        ("axis", _dp), ("axis_stride", C.c_int64), ("n_axis", C.c_int32), ("refine", C.c_int32),
        ("r", _dp), ("sigma", _dp), ("q", _dp),
        ("vol_model", C.c_int32), ("forward", C.c_int32),
        ("sigma_nodes", _dp), ("sigma_nodes_stride", C.c_int64),
        ("T", _dp), ("dt", _dp), ("scheme", C.c_int32), ("max_steps", C.c_int32),
        ("variable_target", _dp), ("variable_scale", C.c_double),
        ("n_exercise", C.c_int32), ("exercise_kind", C.c_int32),
        ("dividend_kind", C.c_int32), ("dividend_first", C.c_int32), ("dividend", _dp),
        ("strike", _dp), ("payoff_kind", C.c_int32), ("obstacle_kind", C.c_int32),
        ("payoff", _dp), ("payoff_stride", C.c_int64),
        ("obstacle", _dp), ("obstacle_stride", C.c_int64),
        ("american", C.c_int32), ("max_inner", C.c_int32),
        ("tolerance", C.c_double), ("scale", C.c_double), ("penalty_tolerance", C.c_double),
        ("spot", _dp),
        ("n_controls", C.c_int32), ("policy_max", C.c_int32),
        ("controls", _dp), ("controls_stride", C.c_int64),
        ("jump_lambda", _dp), ("jump_kappa", _dp), ("jump_x0", _dp), ("jump_dx", _dp),
        ("jump_weights", _dp), ("jump_weights_stride", C.c_int64),
        ("n_fft", C.c_int32), ("event_program_length", C.c_int32),
        ("event_program", C.POINTER(C.c_int32)), ("event_consts", _dp), ("event_params", _dp),
        ("event_params_stride", C.c_int64), ("n_event_consts", C.c_int32), ("n_event_params", C.c_int32),
        ("event_times", _dp), ("source_table", _dp), ("n_source", C.c_int32), ("reserved3", C.c_int32),
        ("kernel", C.c_int32), ("outputs", C.c_int32),
    ]


class Result(C.Structure):
    _fields_ = [
        ("V", _dp), ("V_stride", C.c_int64),
        ("S", _dp), ("S_stride", C.c_int64),
        ("value", _dp),
        ("n_steps", C.POINTER(C.c_int32)),
        ("inner_total", C.POINTER(C.c_int32)),
        ("inner", C.POINTER(C.c_uint8)), ("inner_stride", C.c_int64),
        ("active", C.POINTER(C.c_uint8)), ("active_stride", C.c_int64),
        ("status", C.POINTER(C.c_int32)),
        ("solves_computed", C.POINTER(C.c_int32)),
    ]


class Gmwb2d(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("refine", C.c_int32),
        ("w_axis", _dp), ("n_w_axis", C.c_int32), ("a_axis", _dp), ("n_a_axis", C.c_int32),
        ("controls", _dp), ("n_controls", C.c_int32),
        ("impulse_axis", _dp), ("n_impulse_axis", C.c_int32),
        ("T", C.c_double), ("timesteps", C.c_int32), ("max_inner", C.c_int32),
        ("r", C.c_double), ("eta", C.c_double), ("sigma", C.c_double), ("kappa", C.c_double), ("c", C.c_double),
        ("scaling_factor", C.c_double), ("tolerance", C.c_double),
        ("max_sweeps", C.c_int32), ("reserved", C.c_int32),
    ]


class Gmwb2dStats(C.Structure):
    _fields_ = [
        ("nodes_w", C.c_int32), ("nodes_a", C.c_int32), ("n_steps", C.c_int32), ("status", C.c_int32),
        ("mean_inner", C.c_double), ("inner", C.POINTER(C.c_int32)), ("inner_capacity", C.c_int32),
        ("reserved", C.c_int32),
    ]


class Hjbqvi1dBatch(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("model", C.c_int32), ("n_problems", C.c_int32), ("refine", C.c_int32),
        ("x_axis", _dp), ("n_x_axis", C.c_int32), ("reserved0", C.c_int32), ("x_axis_stride", C.c_int64),
        ("control_axis", _dp), ("n_control_axis", C.c_int32), ("reserved1", C.c_int32), ("control_axis_stride", C.c_int64),
        ("impulse_axis", _dp), ("n_impulse_axis", C.c_int32), ("reserved2", C.c_int32), ("impulse_axis_stride", C.c_int64),
        ("params", _dp), ("n_params", C.c_int32), ("reserved3", C.c_int32), ("params_stride", C.c_int64),
        ("T", C.c_double), ("timesteps", C.c_int32), ("max_inner", C.c_int32),
        ("scaling_factor", C.c_double), ("tolerance", C.c_double),
        ("max_sweeps", C.c_int32), ("reserved4", C.c_int32),
    ]


class Hjbqvi1dResult(C.Structure):
    _fields_ = [
        ("V", _dp), ("X", _dp), ("Q", _dp), ("Z", _dp), ("mask", C.POINTER(C.c_ubyte)),
        ("mean_inner", _dp), ("mean_sweeps", _dp), ("n_steps", C.POINTER(C.c_int32)), ("status", C.POINTER(C.c_int32)),
    ]


HJBQVI1D_EXCHANGE_RATE = 1


class Hjbqvi2d(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("model", C.c_int32), ("refine", C.c_int32), ("n_params", C.c_int32),
        ("w_axis", _dp), ("n_w_axis", C.c_int32), ("reserved0", C.c_int32),
        ("b_axis", _dp), ("n_b_axis", C.c_int32), ("reserved1", C.c_int32),
        ("control_axis", _dp), ("n_control_axis", C.c_int32), ("reserved2", C.c_int32),
        ("impulse_axis", _dp), ("n_impulse_axis", C.c_int32), ("reserved3", C.c_int32),
        ("params", _dp),
        ("T", C.c_double), ("timesteps", C.c_int32), ("max_inner", C.c_int32),
        ("scaling_factor", C.c_double), ("tolerance", C.c_double),
        ("max_sweeps", C.c_int32), ("reserved4", C.c_int32),
    ]


class Hjbqvi2dStats(C.Structure):
    _fields_ = [
        ("nodes_w", C.c_int32), ("nodes_b", C.c_int32), ("n_steps", C.c_int32), ("status", C.c_int32),
        ("mean_inner", C.c_double), ("mean_sweeps", C.c_double),
        ("inner", C.POINTER(C.c_int32)), ("inner_capacity", C.c_int32), ("reserved", C.c_int32),
    ]


HJBQVI2D_OPTIMAL_CONSUMPTION = 1

# every symbol include/qpde_b200.h declares
EXPORTS = [
    "qpde_abi_version", "qpde_create", "qpde_destroy", "qpde_last_error", "qpde_stream",
    "qpde_synchronize", "qpde_launch_count", "qpde_sm_count", "qpde_measure_fp64_peak", "qpde_host_alloc",
    "qpde_host_free", "qpde_make_schedule", "qpde_make_schedule_forward", "qpde_refined_size", "qpde_bs1d_sweep",
    "qpde_bs1d_plan_create", "qpde_bs1d_plan_run", "qpde_bs1d_plan_fetch",
    "qpde_bs1d_plan_destroy", "qpde_bs1d_plan_kernel", "qpde_bs1d_plan_nodes", "qpde_bs1d_plan_tma_staged",
    "qpde_bs1d_plan_max_steps", "qpde_jump_nfft", "qpde_jump_setup", "qpde_refine_axis",
    "qpde_gmwb2d_nodes", "qpde_gmwb2d_solve", "qpde_gmwb2d_solve_sharded", "qpde_comm_unique_id", "qpde_comm_init",
    "qpde_comm_destroy", "qpde_gmwb2d_shard_create", "qpde_gmwb2d_shard_destroy",
    "qpde_gmwb2d_shard_exit", "qpde_gmwb2d_shard_policy", "qpde_gmwb2d_shard_sweep", "qpde_gmwb2d_shard_relerr",
    "qpde_hjbqvi1d_nodes", "qpde_hjbqvi1d_solve", "qpde_hjbqvi2d_nodes", "qpde_hjbqvi2d_solve",
]

_lib = None


class QpdeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libqpde_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Loads quantpde_b200/lib/libqpde_b200.so.  Raises if it has not been built
    (run `python -c 'import __graft_entry__ as g; g.build()'` or make -C quantpde_b200/csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                "%s is missing: build it with __graft_entry__.build(); there is no fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.qpde_last_error.restype = C.c_char_p
        L.qpde_last_error.argtypes = [C.c_void_p]
        L.qpde_stream.restype = C.c_void_p
        L.qpde_stream.argtypes = [C.c_void_p]
        L.qpde_launch_count.restype = C.c_int64
        L.qpde_launch_count.argtypes = [C.c_void_p]
        L.qpde_sm_count.argtypes = [C.c_void_p]
        L.qpde_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.qpde_destroy.argtypes = [C.c_void_p]
        L.qpde_synchronize.argtypes = [C.c_void_p]
        L.qpde_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
        L.qpde_host_free.argtypes = [C.c_void_p]
        L.qpde_make_schedule.argtypes = [C.c_double, C.c_double, _dp, C.c_int,
                                         C.POINTER(Step), C.c_int]
        L.qpde_make_schedule_forward.argtypes = [C.c_double, C.c_double, _dp, C.c_int,
                                                 C.POINTER(Step), C.c_int]
        L.qpde_refined_size.argtypes = [C.c_int, C.c_int]
        L.qpde_bs1d_sweep.argtypes = [C.c_void_p, C.POINTER(Batch), C.POINTER(Result)]
        L.qpde_bs1d_plan_create.argtypes = [C.c_void_p, C.POINTER(Batch), C.POINTER(C.c_void_p)]
        L.qpde_bs1d_plan_run.argtypes = [C.c_void_p]
        L.qpde_bs1d_plan_fetch.argtypes = [C.c_void_p, C.POINTER(Result)]
        L.qpde_bs1d_plan_destroy.argtypes = [C.c_void_p]
        L.qpde_bs1d_plan_kernel.argtypes = [C.c_void_p]
        L.qpde_bs1d_plan_nodes.argtypes = [C.c_void_p]
        L.qpde_bs1d_plan_tma_staged.argtypes = [C.c_void_p]
        L.qpde_bs1d_plan_max_steps.argtypes = [C.c_void_p]
        L.qpde_jump_nfft.argtypes = [C.c_int]
        L.qpde_jump_setup.argtypes = [C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.qpde_refine_axis.argtypes = [_dp, C.c_int, C.c_int, _dp]
        L.qpde_gmwb2d_nodes.argtypes = [C.POINTER(Gmwb2d), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.qpde_gmwb2d_solve.argtypes = [C.c_void_p, C.POINTER(Gmwb2d), _dp, _dp, _dp, C.POINTER(Gmwb2dStats)]
        L.qpde_gmwb2d_solve_sharded.argtypes = [C.c_void_p, C.POINTER(Gmwb2d), _dp, _dp, _dp, C.POINTER(Gmwb2dStats)]
        L.qpde_comm_unique_id.argtypes = [C.c_char_p]
        L.qpde_comm_init.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        L.qpde_comm_destroy.argtypes = [C.c_void_p]
        L.qpde_gmwb2d_shard_create.argtypes = [C.c_void_p, C.POINTER(Gmwb2d), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.qpde_gmwb2d_shard_destroy.argtypes = [C.c_void_p]
        L.qpde_gmwb2d_shard_exit.argtypes = [C.c_void_p, C.c_void_p]
        L.qpde_gmwb2d_shard_policy.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
        L.qpde_gmwb2d_shard_sweep.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.qpde_gmwb2d_shard_relerr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.qpde_hjbqvi1d_nodes.argtypes = [C.POINTER(Hjbqvi1dBatch), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.qpde_hjbqvi1d_solve.argtypes = [C.c_void_p, C.POINTER(Hjbqvi1dBatch), C.POINTER(Hjbqvi1dResult)]
        L.qpde_hjbqvi2d_nodes.argtypes = [C.POINTER(Hjbqvi2d), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.qpde_hjbqvi2d_solve.argtypes = [C.c_void_p, C.POINTER(Hjbqvi2d), _dp, _dp, _dp, C.POINTER(C.c_ubyte), _dp, _dp,
                                          C.POINTER(Hjbqvi2dStats)]
        _lib = L
    return _lib


def check(code, handle=None):
    if code != OK:
        msg = lib().qpde_last_error(handle)
        raise QpdeError(code, msg.decode() if msg else "")
    return code


def make_schedule(T, dt, event_times=(), forward=False):
    """Host-only replay of the reference time loop (no GPU needed); forward: ForwardConstantStepper."""
    ev = np.ascontiguousarray(event_times, dtype=np.float64)
    evp = ev.ctypes.data_as(_dp) if len(ev) else None
    fn = lib().qpde_make_schedule_forward if forward else lib().qpde_make_schedule
    n = fn(T, dt, evp, len(ev), None, 0)
    if n < 0:
        raise QpdeError(n, "qpde_make_schedule")
    buf = (Step * n)()
    fn(T, dt, evp, len(ev), buf, n)
    return buf


def refine_axis(axis, refine):
    """RectilinearGrid1(axis).refined(refine) on the host (no GPU needed)."""
    axis = np.ascontiguousarray(axis, dtype=np.float64)
    out = np.zeros(refined_size(len(axis), refine))
    n = lib().qpde_refine_axis(axis.ctypes.data_as(_dp), len(axis), refine, out.ctypes.data_as(_dp))
    assert n == len(out)
    return out


def jump_setup(S, density="lognormal", params=(-0.1, 0.45)):
    """(x0, dx, kappa, weights): the host-side set-up of BlackScholesJumpDiffusion (no GPU needed)."""
    S = np.ascontiguousarray(S, dtype=np.float64)
    nfft = lib().qpde_jump_nfft(len(S))
    w = np.zeros(nfft)
    par = np.zeros(3); par[:len(params)] = params
    x0, dx, kappa = C.c_double(), C.c_double(), C.c_double()
    rc = lib().qpde_jump_setup(len(S), S.ctypes.data_as(_dp), 1 if density == "kou" else 0,
                               par.ctypes.data_as(_dp), C.byref(x0), C.byref(dx), C.byref(kappa),
                               w.ctypes.data_as(_dp))
    if rc != OK:
        raise QpdeError(rc, "qpde_jump_setup")
    return x0.value, dx.value, kappa.value, w


def refined_size(n_axis, refine):
    return lib().qpde_refined_size(n_axis, refine)


class Handle:
    """qpde_create / qpde_destroy."""

    def __init__(self, device=0):
        self.ptr = C.c_void_p()
        check(lib().qpde_create(device, C.byref(self.ptr)))

    def close(self):
        if self.ptr:
            lib().qpde_destroy(self.ptr)
            self.ptr = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def stream(self):
        return lib().qpde_stream(self.ptr)

    def synchronize(self):
        check(lib().qpde_synchronize(self.ptr), self.ptr)

    @property
    def launches(self):
        return lib().qpde_launch_count(self.ptr)

    @property
    def sm_count(self):
        return lib().qpde_sm_count(self.ptr)

    def comm_init(self, comm_id, rank, world):
        """qpde_comm_init: joins the NCCL communicator identified by the 128 bytes rank 0 made with
        comm_unique_id() (shipped by the caller).  Collective over the `world` ranks."""
        assert len(comm_id) == COMM_ID_BYTES
        check(lib().qpde_comm_init(self.ptr, bytes(comm_id), int(rank), int(world)), self.ptr)

    def comm_destroy(self):
        lib().qpde_comm_destroy(self.ptr)

    def fp64_peak_tflops(self):
        """Measured FP64 rate of the device (qpde_measure_fp64_peak)."""
        v = C.c_double()
        check(lib().qpde_measure_fp64_peak(self.ptr, C.byref(v)), self.ptr)
        return v.value


def _arr(x, n, dtype=np.float64):
    a = np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=dtype), (n,)))
    return a


class GmwbProblem:
    """Keeps the arrays a qpde_gmwb2d points into alive."""

    def __init__(self, w_axis, a_axis, controls, impulse_axis, refine, timesteps, T=10., r=.05, eta=0.,
                 sigma=.2, kappa=.1, c=0., scaling_factor=1e-2, tolerance=1e-6, max_inner=200, max_sweeps=500):
        self.keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (w_axis, a_axis, controls, impulse_axis)]
        keep = self.keep
        g = Gmwb2d()
        g.abi_version, g.refine = ABI_VERSION, int(refine)
        g.w_axis, g.n_w_axis = keep[0].ctypes.data_as(_dp), len(keep[0])
        g.a_axis, g.n_a_axis = keep[1].ctypes.data_as(_dp), len(keep[1])
        g.controls, g.n_controls = keep[2].ctypes.data_as(_dp), len(keep[2])
        g.impulse_axis, g.n_impulse_axis = keep[3].ctypes.data_as(_dp), len(keep[3])
        g.T, g.timesteps, g.max_inner = T, int(timesteps), int(max_inner)
        g.r, g.eta, g.sigma, g.kappa, g.c = r, eta, sigma, kappa, c
        g.scaling_factor, g.tolerance, g.max_sweeps = scaling_factor, tolerance, int(max_sweeps)
        self.struct = g
        nw, na = C.c_int(), C.c_int()
        check(lib().qpde_gmwb2d_nodes(C.byref(g), C.byref(nw), C.byref(na)))
        self.nw, self.na = nw.value, na.value
        self.W = refine_axis(keep[0], refine)
        self.A = refine_axis(keep[1], refine)


class GmwbShard:
    """qpde_gmwb2d_shard_*: one rank's A-lines [a_begin, a_end); the vectors are device
    pointers (torch CUDA tensors of the full grid)."""

    def __init__(self, handle, problem, a_begin, a_end):
        self.handle, self.problem = handle, problem
        self.ptr = C.c_void_p()
        check(lib().qpde_gmwb2d_shard_create(handle.ptr, C.byref(problem.struct), int(a_begin), int(a_end),
                                             C.byref(self.ptr)), handle.ptr)

    def exit_values(self, x):
        check(lib().qpde_gmwb2d_shard_exit(self.ptr, x.data_ptr()), self.handle.ptr)

    def policy(self, x, h0):
        check(lib().qpde_gmwb2d_shard_policy(self.ptr, x.data_ptr(), float(h0)), self.handle.ptr)

    def sweep(self, x, v0, status):
        check(lib().qpde_gmwb2d_shard_sweep(self.ptr, x.data_ptr(), v0.data_ptr(), status.data_ptr()), self.handle.ptr)

    def relerr(self, x, xprev, notdone):
        check(lib().qpde_gmwb2d_shard_relerr(self.ptr, x.data_ptr(), xprev.data_ptr(), notdone.data_ptr()),
              self.handle.ptr)

    def close(self):
        if self.ptr:
            lib().qpde_gmwb2d_shard_destroy(self.ptr)
            self.ptr = C.c_void_p()


COMM_ID_BYTES = 128


def comm_unique_id():
    """qpde_comm_unique_id: the 128 bytes that identify a new communicator (rank 0 makes them)."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    check(lib().qpde_comm_unique_id(buf))
    return buf.raw


def gmwb_solve(handle, w_axis, a_axis, controls, impulse_axis, refine, timesteps, T=10., r=.05, eta=0.,
               sigma=.2, kappa=.1, c=0., scaling_factor=1e-2, tolerance=1e-6, max_inner=200, max_sweeps=500,
               sharded=False):
    """qpde_gmwb2d_solve: the 2-D GMWB HJBQVI (host arrays in, host arrays out).  sharded=True:
    qpde_gmwb2d_solve_sharded, a collective call over the ranks of the handle's communicator."""
    keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (w_axis, a_axis, controls, impulse_axis)]
    g = Gmwb2d()
    g.abi_version, g.refine = ABI_VERSION, int(refine)
    g.w_axis, g.n_w_axis = keep[0].ctypes.data_as(_dp), len(keep[0])
    g.a_axis, g.n_a_axis = keep[1].ctypes.data_as(_dp), len(keep[1])
    g.controls, g.n_controls = keep[2].ctypes.data_as(_dp), len(keep[2])
    g.impulse_axis, g.n_impulse_axis = keep[3].ctypes.data_as(_dp), len(keep[3])
    g.T, g.timesteps, g.max_inner = T, int(timesteps), int(max_inner)
    g.r, g.eta, g.sigma, g.kappa, g.c = r, eta, sigma, kappa, c
    g.scaling_factor, g.tolerance, g.max_sweeps = scaling_factor, tolerance, int(max_sweeps)
    nw, na = C.c_int(), C.c_int()
    check(lib().qpde_gmwb2d_nodes(C.byref(g), C.byref(nw), C.byref(na)))
    V = np.zeros((na.value, nw.value)); W = np.zeros(nw.value); A = np.zeros(na.value)
    inner = np.zeros(int(timesteps) + 8, dtype=np.int32)
    st = Gmwb2dStats()
    st.inner = inner.ctypes.data_as(C.POINTER(C.c_int32)); st.inner_capacity = len(inner)
    fn = lib().qpde_gmwb2d_solve_sharded if sharded else lib().qpde_gmwb2d_solve
    check(fn(handle.ptr, C.byref(g), V.ctypes.data_as(_dp), W.ctypes.data_as(_dp), A.ctypes.data_as(_dp), C.byref(st)),
          handle.ptr)
    return dict(V=V, W=W, A=A, n_steps=st.n_steps, mean_inner=st.mean_inner, status=st.status,
                inner=inner[:st.n_steps].copy())


def exchange_rate_solve(handle, x_axis, q_axis, z_axis, params, refine, timesteps, T=10., scaling_factor=1e-2,
                        tolerance=1e-6, max_inner=200, max_sweeps=500):
    """qpde_hjbqvi1d_solve for the exchange-rate model (examples/hjbqvi/exchange_rate.cpp): a batch of problems.
    x_axis / q_axis / z_axis: coarse axes, one ([n]) shared by the batch or one per problem ([P][n]);
    params: [P][7] = rho, sigma, a, b, m, lambda, c.  Host arrays in, host arrays out."""
    params = np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)
    P = params.shape[0]
    axes = [np.ascontiguousarray(a, dtype=np.float64) for a in (x_axis, q_axis, z_axis)]
    g = Hjbqvi1dBatch()
    g.abi_version, g.model, g.n_problems, g.refine = ABI_VERSION, HJBQVI1D_EXCHANGE_RATE, P, int(refine)
    for name, a in zip(("x", "control", "impulse"), axes):
        if a.ndim == 2 and a.shape[0] != P:
            raise ValueError("%s axis: one row per problem" % name)
        setattr(g, name + "_axis", a.ctypes.data_as(_dp))
        setattr(g, "n_%s_axis" % name, a.shape[-1])
        setattr(g, name + "_axis_stride", a.shape[-1] if a.ndim == 2 else 0)
    g.params, g.n_params, g.params_stride = params.ctypes.data_as(_dp), params.shape[1], params.shape[1]
    g.T, g.timesteps, g.max_inner = T, int(timesteps), int(max_inner)
    g.scaling_factor, g.tolerance, g.max_sweeps = scaling_factor, tolerance, int(max_sweeps)
    n, nq, nz = C.c_int(), C.c_int(), C.c_int()
    check(lib().qpde_hjbqvi1d_nodes(C.byref(g), C.byref(n), C.byref(nq), C.byref(nz)))
    N = n.value
    V, X, Q, Z = (np.zeros((P, N)) for _ in range(4))
    mask = np.zeros((P, N), dtype=np.uint8)
    mean_inner = np.zeros(P); mean_sweeps = np.zeros(P); n_steps = np.zeros(P, dtype=np.int32); status = np.zeros(P, dtype=np.int32)
    r = Hjbqvi1dResult()
    r.V, r.X, r.Q, r.Z = (a.ctypes.data_as(_dp) for a in (V, X, Q, Z))
    r.mask = mask.ctypes.data_as(C.POINTER(C.c_ubyte))
    r.mean_inner = mean_inner.ctypes.data_as(_dp); r.mean_sweeps = mean_sweeps.ctypes.data_as(_dp)
    r.n_steps = n_steps.ctypes.data_as(C.POINTER(C.c_int32)); r.status = status.ctypes.data_as(C.POINTER(C.c_int32))
    check(lib().qpde_hjbqvi1d_solve(handle.ptr, C.byref(g), C.byref(r)), handle.ptr)
    return dict(V=V, X=X, Q=Q, Z=Z, mask=mask, mean_inner=mean_inner, mean_sweeps=mean_sweeps, n_steps=n_steps, status=status,
                n_controls=nq.value, n_impulses=nz.value)


def optimal_consumption_solve(handle, refine, timesteps=32, control_points=16, spatial_points=20, T=40., rho=.1, r=.07, mu=.11,
                              sigma=.3, gamma=.3, lam=.1, c=.05, q_max=100., boundary=200., scaling_factor=1e-2,
                              tolerance=1e-6, max_inner=200, max_sweeps=20000, spatial_points_b=None):
    """qpde_hjbqvi2d_solve for the model of examples/hjbqvi/optimal_consumption.cpp; `timesteps` is the count at
    refinement 0 (doubled per level, as HJBQVI::solve does).  Host arrays out."""
    def uniform(begin, end, points):
        dx = (end - begin) / (points - 1)                       # Axis::uniform, Core/Axis.hpp:219-226
        return np.array([begin + dx * i for i in range(points)])
    axis = uniform(0., boundary, spatial_points)
    axis_b = axis if spatial_points_b is None else uniform(0., boundary, spatial_points_b)
    qa, za = uniform(0., q_max, control_points), uniform(0., 1., control_points)
    par = np.array([rho, r, mu, sigma, gamma, lam, c, boundary])
    g = Hjbqvi2d()
    g.abi_version, g.model, g.refine, g.n_params = ABI_VERSION, HJBQVI2D_OPTIMAL_CONSUMPTION, int(refine), len(par)
    g.w_axis, g.n_w_axis, g.b_axis, g.n_b_axis = axis.ctypes.data_as(_dp), len(axis), axis_b.ctypes.data_as(_dp), len(axis_b)
    g.control_axis, g.n_control_axis = qa.ctypes.data_as(_dp), len(qa)
    g.impulse_axis, g.n_impulse_axis = za.ctypes.data_as(_dp), len(za)
    g.params = par.ctypes.data_as(_dp)
    g.T, g.timesteps, g.max_inner = T, int(timesteps) * 2 ** int(refine), int(max_inner)
    g.scaling_factor, g.tolerance, g.max_sweeps = scaling_factor, tolerance, int(max_sweeps)
    nw, nb = C.c_int(), C.c_int()
    check(lib().qpde_hjbqvi2d_nodes(C.byref(g), C.byref(nw), C.byref(nb)))
    shape = (nb.value, nw.value)
    V, Q, Z = np.zeros(shape), np.zeros(shape), np.zeros(shape)
    mask = np.zeros(shape, dtype=np.uint8)
    W, B = np.zeros(nw.value), np.zeros(nb.value)
    st = Hjbqvi2dStats()
    check(lib().qpde_hjbqvi2d_solve(handle.ptr, C.byref(g), V.ctypes.data_as(_dp), Q.ctypes.data_as(_dp), Z.ctypes.data_as(_dp),
                                    mask.ctypes.data_as(C.POINTER(C.c_ubyte)), W.ctypes.data_as(_dp), B.ctypes.data_as(_dp),
                                    C.byref(st)), handle.ptr)
    return dict(V=V, Q=Q, Z=Z, mask=mask, W=W, B=B, n_steps=st.n_steps, mean_inner=st.mean_inner, mean_sweeps=st.mean_sweeps,
                status=st.status)


class BatchSpec:
    """Keeps the numpy arrays a qpde_bs1d_batch points into alive."""

    def __init__(self, *, axis, refine, r, sigma, q, T, dt, strike, scheme="rannacher",
                 american=False, payoff="put", obstacle=None, payoff_array=None,
                 obstacle_array=None, vol_model="const", sigma_nodes=None,
                 n_exercise=0, exercise="k_minus_s", spot=None, tolerance=1e-6, scale=1.0,
                 penalty_tolerance=1e-6, max_inner=100, kernel="auto", outputs=OUT_ALL,
                 n_contracts=None, controls=None, policy_max=False, jump=None, dividend=None,
                 variable_target=None, variable_scale=1.0, max_steps=0, forward=False, event=None, event_times=None,
                 source=None):
        axis = np.ascontiguousarray(axis, dtype=np.float64)
        if n_contracts is None:
            n_contracts = axis.shape[0] if axis.ndim == 2 else len(np.atleast_1d(strike))
        Cn = int(n_contracts)
        self.C = Cn
        self.keep = []
        b = Batch()
        b.abi_version = ABI_VERSION
        b.n_contracts = Cn
        self.axis = axis
        b.axis = axis.ctypes.data_as(_dp)
        b.axis_stride = axis.shape[1] if axis.ndim == 2 else 0
        b.n_axis = axis.shape[-1]
        b.refine = int(refine)
        self.N = refined_size(b.n_axis, b.refine)

        def vec(x):
            a = _arr(x, Cn)
            self.keep.append(a)
            return a.ctypes.data_as(_dp)

        def rows(x):
            a = np.ascontiguousarray(x, dtype=np.float64)
            self.keep.append(a)
            stride = a.shape[1] if a.ndim == 2 else 0
            return a.ctypes.data_as(_dp), stride

        b.r, b.sigma, b.q = vec(r), vec(sigma), vec(q)
        b.vol_model = VOL[vol_model]
        b.forward = int(bool(forward))
        if sigma_nodes is not None:
            b.sigma_nodes, b.sigma_nodes_stride = rows(sigma_nodes)
        b.T, b.dt = vec(T), vec(dt)
        b.scheme = SCHEME[scheme]
        b.max_steps = int(max_steps)
        if variable_target is not None:
            b.variable_target = vec(variable_target)
            b.variable_scale = float(variable_scale)
        b.n_exercise = int(n_exercise)
        b.exercise_kind = PAYOFF[exercise]
        if event_times is not None:
            et = np.ascontiguousarray(event_times, dtype=np.float64)
            assert len(et) == int(n_exercise)
            self.keep.append(et)
            b.event_times = et.ctypes.data_as(_dp)
        if source is not None:
            src = np.ascontiguousarray(source, dtype=np.float64)
            self.keep.append(src)
            b.source_table, b.n_source = src.ctypes.data_as(_dp), len(src)
        if event is not None:
            # event = dict(program=int32 words, consts=[..], params=[n_exercise][n_params] or [C][n_exercise][n_params])
            prog = np.ascontiguousarray(event["program"], dtype=np.int32)
            consts = np.ascontiguousarray(event.get("consts", []), dtype=np.float64)
            params = np.ascontiguousarray(event.get("params", np.zeros((int(n_exercise), 0))), dtype=np.float64)
            self.keep += [prog, consts, params]
            b.event_program = prog.ctypes.data_as(C.POINTER(C.c_int32))
            b.event_program_length = len(prog)
            b.event_consts, b.n_event_consts = consts.ctypes.data_as(_dp), len(consts)
            b.event_params, b.n_event_params = params.ctypes.data_as(_dp), params.shape[-1]
            b.event_params_stride = params.shape[1] * params.shape[2] if params.ndim == 3 else 0
        if dividend is not None:
            # dividend = dict(kind="cash" | "proportional", amount=[C] or scalar, first=bool)
            b.dividend_kind = DIVIDEND[dividend["kind"]]
            b.dividend_first = int(bool(dividend.get("first", False)))
            b.dividend = vec(dividend["amount"])
        b.strike = vec(strike)
        b.payoff_kind = PAYOFF["array"] if payoff_array is not None else PAYOFF[payoff]
        if payoff_array is not None:
            b.payoff, b.payoff_stride = rows(payoff_array)
        if obstacle is None:
            obstacle = payoff
        b.obstacle_kind = PAYOFF["array"] if obstacle_array is not None else PAYOFF[obstacle]
        if obstacle_array is not None:
            b.obstacle, b.obstacle_stride = rows(obstacle_array)
        b.american = int(bool(american))
        b.max_inner = int(max_inner)
        b.tolerance, b.scale, b.penalty_tolerance = tolerance, scale, penalty_tolerance
        if spot is not None:
            b.spot = vec(spot)
        if jump is not None:
            # jump = dict(lam=[C], kappa=[C], x0=[C], dx=[C], weights=[C][n_fft] or [n_fft])
            b.jump_lambda, b.jump_kappa = vec(jump["lam"]), vec(jump["kappa"])
            b.jump_x0, b.jump_dx = vec(jump["x0"]), vec(jump["dx"])
            b.jump_weights, b.jump_weights_stride = rows(jump["weights"])
            b.n_fft = np.asarray(jump["weights"]).shape[-1]
        if controls is not None:
            ctl = np.ascontiguousarray(controls, dtype=np.float64)
            self.keep.append(ctl)
            b.controls = ctl.ctypes.data_as(_dp)
            b.controls_stride = ctl.shape[1] if ctl.ndim == 2 else 0
            b.n_controls = ctl.shape[-1]
            b.policy_max = int(bool(policy_max))
        b.kernel = KERNEL[kernel]
        b.outputs = int(outputs)
        self.struct = b


class ResultBuffers:
    """Host output buffers for qpde_bs1d_result."""

    def __init__(self, C_, N, max_steps, outputs=OUT_ALL, pinned=False):
        """pinned=True backs the large arrays with qpde_host_alloc (page-locked
        memory: full-rate, asynchronous D2H); call close() to release them."""
        r = Result()
        self.V = self.S = self.value = self.n_steps = self.inner_total = None
        self.inner = self.active = self.status = self.computed = None
        self._pinned = []

        def big(shape):
            if not pinned:
                return np.zeros(shape)
            n = int(np.prod(shape))
            ptr = C.c_void_p()
            check(lib().qpde_host_alloc(n * 8, C.byref(ptr)))
            self._pinned.append(ptr)
            arr = np.ctypeslib.as_array(C.cast(ptr, _dp), shape=(n,)).reshape(shape)
            arr[...] = 0.0
            return arr

        if outputs & OUT_V:
            self.V = big((C_, N)); r.V = self.V.ctypes.data_as(_dp); r.V_stride = N
        if outputs & OUT_S:
            self.S = big((C_, N)); r.S = self.S.ctypes.data_as(_dp); r.S_stride = N
        if outputs & OUT_VALUE:
            self.value = np.zeros(C_); r.value = self.value.ctypes.data_as(_dp)
        if outputs & OUT_STEPS:
            self.n_steps = np.zeros(C_, dtype=np.int32)
            r.n_steps = self.n_steps.ctypes.data_as(C.POINTER(C.c_int32))
        if outputs & OUT_INNER_TOTAL:
            self.inner_total = np.zeros(C_, dtype=np.int32)
            r.inner_total = self.inner_total.ctypes.data_as(C.POINTER(C.c_int32))
        if outputs & OUT_INNER:
            self.inner = np.zeros((C_, max_steps), dtype=np.uint8)
            r.inner = self.inner.ctypes.data_as(C.POINTER(C.c_uint8)); r.inner_stride = max_steps
        if outputs & OUT_ACTIVE:
            self.active = np.zeros((C_, N), dtype=np.uint8)
            r.active = self.active.ctypes.data_as(C.POINTER(C.c_uint8)); r.active_stride = N
        if outputs & OUT_STATUS:
            self.status = np.zeros(C_, dtype=np.int32)
            r.status = self.status.ctypes.data_as(C.POINTER(C.c_int32))
        if outputs & OUT_COMPUTED:
            self.computed = np.zeros(C_, dtype=np.int32)
            r.solves_computed = self.computed.ctypes.data_as(C.POINTER(C.c_int32))
        self.struct = r


    def close(self):
        self.V = self.S = None
        for ptr in self._pinned:
            lib().qpde_host_free(ptr)
        self._pinned = []


class Plan:
    """qpde_bs1d_plan_*: a batch resident in HBM."""

    def __init__(self, handle, spec):
        self.handle, self.spec = handle, spec
        self.ptr = C.c_void_p()
        check(lib().qpde_bs1d_plan_create(handle.ptr, C.byref(spec.struct), C.byref(self.ptr)), handle.ptr)
        self.N = lib().qpde_bs1d_plan_nodes(self.ptr)
        self.max_steps = lib().qpde_bs1d_plan_max_steps(self.ptr)
        self.kernel = KERNEL_NAME[lib().qpde_bs1d_plan_kernel(self.ptr)]
        self.tma_staged = bool(lib().qpde_bs1d_plan_tma_staged(self.ptr))

    def run(self):
        check(lib().qpde_bs1d_plan_run(self.ptr), self.handle.ptr)

    def fetch(self, outputs=None):
        outputs = self.spec.struct.outputs if outputs is None else outputs
        res = ResultBuffers(self.spec.C, self.N, self.max_steps, outputs)
        check(lib().qpde_bs1d_plan_fetch(self.ptr, C.byref(res.struct)), self.handle.ptr)
        return res

    def close(self):
        if self.ptr:
            lib().qpde_bs1d_plan_destroy(self.ptr)
            self.ptr = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def sweep_max_steps(spec):
    if spec.struct.max_steps > 0:
        return int(spec.struct.max_steps)
    Tmax = np.max(np.ctypeslib.as_array(spec.struct.T, (spec.C,)) /
                  np.ctypeslib.as_array(spec.struct.dt, (spec.C,)))
    return int(np.floor(Tmax)) + 3 + 2 * spec.struct.n_exercise


def sweep(handle, spec, outputs=OUT_ALL, result=None):
    """qpde_bs1d_sweep: host buffers in, host buffers out.  `result` reuses a
    ResultBuffers (e.g. page-locked) across calls."""
    res = result if result is not None else ResultBuffers(spec.C, spec.N, sweep_max_steps(spec), outputs)
    check(lib().qpde_bs1d_sweep(handle.ptr, C.byref(spec.struct), C.byref(res.struct)), handle.ptr)
    return res
