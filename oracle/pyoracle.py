"""TEST INFRASTRUCTURE ONLY: ctypes view of oracle/_ref/liboracle.so (the plain-C
restatement in oracle/qpde_oracle.c) plus a runner for oracle/_ref/qpde_ref (the
unmodified reference headers compiled against oracle/shim).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "liboracle.so")
REF_PATH = os.path.join(HERE, "_ref", "qpde_ref")

SCHEMES = {"rannacher": 0, "cn": 1, "bdf1": 2, "bdf2": 3, "bdf3": 4, "bdf4": 5, "bdf5": 6, "bdf6": 7}


class Step(C.Structure):
    _fields_ = [("t", C.c_double), ("dt", C.c_double), ("same_dt", C.c_int),
                ("event_after", C.c_int), ("user_events", C.c_int)]


class Problem(C.Structure):
    _fields_ = [("n", C.c_int), ("T", C.c_double),
                ("S", C.POINTER(C.c_double)),
                ("lo", C.POINTER(C.c_double)), ("di", C.POINTER(C.c_double)),
                ("up", C.POINTER(C.c_double)),
                ("v0", C.POINTER(C.c_double)),
                ("scheme", C.c_int), ("n_steps", C.c_int),
                ("steps", C.POINTER(Step)),
                ("american", C.c_int),
                ("obstacle", C.POINTER(C.c_double)),
                ("tolerance", C.c_double), ("scale", C.c_double),
                ("penalty_tol", C.c_double), ("max_inner", C.c_int),
                ("event_obstacle", C.POINTER(C.c_double)),
                ("n_controls", C.c_int), ("policy_max", C.c_int),
                ("plo", C.POINTER(C.c_double)), ("pdi", C.POINTER(C.c_double)),
                ("pup", C.POINTER(C.c_double)),
                ("jump_weights", C.POINTER(C.c_double)), ("jump_nfft", C.c_int),
                ("jump_lambda", C.c_double), ("jump_x0", C.c_double), ("jump_dx", C.c_double),
                ("solver", C.c_int),
                ("dividend_kind", C.c_int), ("dividend_first", C.c_int), ("dividend", C.c_double),
                ("variable_target", C.c_double), ("variable_scale", C.c_double), ("variable_dt0", C.c_double),
                ("steps_taken", C.POINTER(C.c_int)),
                ("event_program", C.POINTER(C.c_int)), ("event_consts", C.POINTER(C.c_double)),
                ("event_params", C.POINTER(C.c_double)), ("n_events", C.c_int), ("n_event_params", C.c_int),
                ("source_table", C.POINTER(C.c_double)), ("n_source", C.c_int),
                ("forward", C.c_int)]


DIVIDENDS = {None: 0, "none": 0, "cash": 1, "proportional": 2}


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.qpo_interp.restype = C.c_double
        _lib.qpo_interp.argtypes = [C.c_int, C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.c_double]
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def special_axis():
    out = np.zeros(33)
    lib().qpo_axis_special(_dp(out))
    return out


def axis_union(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    out = np.zeros(len(a) + len(b))
    n = lib().qpo_axis_union(_dp(a), len(a), _dp(b), len(b), _dp(out))
    return out[:n].copy()


def refine(axis, times):
    axis = np.ascontiguousarray(axis, dtype=np.float64)
    n = lib().qpo_axis_refine(_dp(axis), len(axis), int(times), None)
    out = np.zeros(n)
    lib().qpo_axis_refine(_dp(axis), len(axis), int(times), _dp(out))
    return out


TUTORIAL_AXIS = np.array([
    0., 10., 20., 30., 40., 50., 60., 70., 75., 80., 84., 88., 92., 94., 96.,
    98., 100., 102., 104., 106., 108., 110., 114., 118., 123., 130., 140.,
    150., 175., 225., 300., 750., 2000., 10000.])


def vanilla_grid(K, S0=None, refine_times=0):
    """(S_0 * Axis::special) + (K * Axis::special), refined
    (examples/vanilla_options.cpp:156, :48)."""
    S0 = K if S0 is None else S0
    sp = special_axis()
    return refine(axis_union(sp * S0, sp * K), refine_times)


def schedule(T, dt, event_times=(), forward=False):
    ev = np.ascontiguousarray(event_times, dtype=np.float64)
    cap = int(T / dt) + 8 + 2 * len(ev)
    buf = (Step * cap)()
    fn = lib().qpo_schedule_forward if forward else lib().qpo_schedule
    n = fn(C.c_double(T), C.c_double(dt), _dp(ev) if len(ev) else None, len(ev), buf, cap)
    assert n <= cap
    return buf, n


def bs_coeffs(S, r, v, q, lam=None, kappa=0.0):
    S = np.ascontiguousarray(S, dtype=np.float64)
    n = len(S)
    bc = lambda x: np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (n,)))
    r, v, q = bc(r), bc(v), bc(q)
    lamp = _dp(bc(lam)) if lam is not None else None
    lo, di, up = np.zeros(n), np.zeros(n), np.zeros(n)
    lib().qpo_bs_coeffs(n, _dp(S), _dp(r), _dp(v), _dp(q), lamp,
                        C.c_double(kappa), _dp(lo), _dp(di), _dp(up))
    return lo, di, up


def interp(x, v, c):
    x = np.ascontiguousarray(x, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    return lib().qpo_interp(len(x), _dp(x), _dp(v), C.c_double(c))


def sweep(S, lo, di, up, v0, T, steps, n_steps, scheme="rannacher",
          american=False, obstacle=None, event_obstacle=None,
          tolerance=1e-6, scale=1.0, penalty_tol=1e-6, max_inner=1000, solver=0,
          policy_rows=None, policy_max=False, jump=None, dividend=None, variable=None, forward=False, event=None,
          source=None):
    """Returns dict(V, inner, mask, status).  dividend = (kind, D, first); variable = (dt0, target,
    scale): ReverseVariableStepper, `steps` ignored, n_steps = capacity, out['n_steps'] = steps taken.  policy_rows = (lo, di, up), each
    [n_controls][n]: PolicyIteration over the controls."""
    n = len(S)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (S, lo, di, up, v0)]
    p = Problem()
    p.n = n
    p.T = T
    p.S, p.lo, p.di, p.up, p.v0 = [_dp(a) for a in arrs]
    p.scheme = SCHEMES[scheme]
    p.n_steps = n_steps
    if steps is not None:
        p.steps = C.cast(steps, C.POINTER(Step))
    taken = C.c_int(n_steps)
    if variable is not None:
        p.variable_dt0, p.variable_target, p.variable_scale = float(variable[0]), float(variable[1]), float(variable[2])
        p.steps_taken = C.pointer(taken)
    p.american = int(american)
    if obstacle is not None:
        obstacle = np.ascontiguousarray(obstacle, dtype=np.float64)
        p.obstacle = _dp(obstacle)
    if event_obstacle is not None:
        event_obstacle = np.ascontiguousarray(event_obstacle, dtype=np.float64)
        p.event_obstacle = _dp(event_obstacle)
    p.tolerance, p.scale, p.penalty_tol, p.max_inner = tolerance, scale, penalty_tol, max_inner
    p.solver = solver
    p.forward = int(forward)
    if source is not None:
        src = np.ascontiguousarray(source, dtype=np.float64)
        p.source_table, p.n_source = _dp(src), len(src)
    if event is not None:
        # event = dict(program=int words, consts=[..], params=[n_events][n_params])
        ev_prog = np.ascontiguousarray(event["program"], dtype=np.int32)
        ev_consts = np.ascontiguousarray(list(event.get("consts", [])) + [0.0], dtype=np.float64)
        ev_params = np.ascontiguousarray(event["params"], dtype=np.float64)
        p.event_program = ev_prog.ctypes.data_as(C.POINTER(C.c_int))
        p.event_consts, p.event_params = _dp(ev_consts), _dp(ev_params)
        p.n_events, p.n_event_params = ev_params.shape[0], ev_params.shape[1]
    if dividend is not None:
        p.dividend_kind, p.dividend, p.dividend_first = DIVIDENDS[dividend[0]], float(dividend[1]), int(dividend[2])
    if jump is not None:
        jw = np.ascontiguousarray(jump["weights"], dtype=np.float64)
        p.jump_weights = _dp(jw)
        p.jump_nfft = len(jw)
        p.jump_lambda, p.jump_x0, p.jump_dx = jump["lambda"], jump["x0"], jump["dx"]
    if policy_rows is not None:
        prs = [np.ascontiguousarray(a, dtype=np.float64) for a in policy_rows]
        p.n_controls = prs[0].shape[0]
        p.policy_max = int(policy_max)
        p.plo, p.pdi, p.pup = [_dp(a) for a in prs]
    V = np.zeros(n)
    inner = np.zeros(n_steps, dtype=np.int32)
    mask = np.zeros(n, dtype=np.uint8)
    status = lib().qpo_sweep_1d(C.byref(p), _dp(V),
                                inner.ctypes.data_as(C.POINTER(C.c_int)),
                                mask.ctypes.data_as(C.POINTER(C.c_ubyte)))
    if variable is not None:
        inner = inner[:taken.value]
    return dict(V=V, inner=inner, mask=mask, status=status, n_steps=taken.value)


def american_put(K, r, vol, T, n_steps, refine_times, q=0.0, S0=None, call=False,
                 american=True, scheme="rannacher", solver=0, variable_target=None, max_steps=100000,
                 E=0, dividend=None, dividend_kind="cash", dividend_first=False, exercise=False, event=None):
    """The vanilla_options.cpp wiring on the oracle.  variable_target: ReverseVariableStepper(0, T,
    T / n_steps, target) as vanilla_options.cpp:52-58 builds it.  E > 0: events at T e / E on top of the
    penalty method — the README's discrete dividends (README.md:357-375) and / or the exercise transform."""
    S = vanilla_grid(K, S0, refine_times)
    lo, di, up = bs_coeffs(S, r, vol, q)
    payoff = np.where(S < K, 0.0, S - K) if call else np.where(S < K, K - S, 0.0)
    if variable_target:
        out = sweep(S, lo, di, up, payoff, T, None, max_steps, scheme=scheme, american=american, obstacle=payoff,
                    solver=solver, variable=(T / n_steps, variable_target, 1.0))
        out["S"] = S
        return out
    steps, n = schedule(T, T / n_steps, [T / E * e for e in range(E)])
    out = sweep(S, lo, di, up, payoff, T, steps, n, scheme=scheme,
                american=american, obstacle=payoff, solver=solver,
                event_obstacle=(K - S) if (E > 0 and exercise) else None,
                dividend=None if (E == 0 or dividend is None) else (dividend_kind, dividend, dividend_first), event=event)
    out["S"] = S
    out["n_steps"] = n
    return out


def glwb_event(E, bonus=.06, GQ=5.):
    """The three-way choice of examples/experimental/glwb.cpp:84-121 (bonus / withdrawal / surrender) as an event
    program, with the per-event survival and penalty rates oracle/shim/ref_driver.cpp uses (event=glwb):
        best = (1 + bonus) V(S / (1 + bonus));  tmp = V(max(S - GQ, 0)) + R GQ;  if(S > GQ) tmp = R (GQ + (1 - pen)(S - GQ))
    consts: [1 + bonus, GQ, 0, 1, -inf]; params per event: [R_e, pen_e]."""
    OPS = {"end": 0, "S": 1, "const": 2, "param": 3, "add": 4, "sub": 5, "mul": 6, "div": 7, "max": 8, "min": 9,
           "neg": 10, "V": 11, "gt": 12, "select": 13}
    toks = [("const", 0), "S", ("const", 0), "div", "V", "mul",                                  # (1 + b) V(S / (1 + b))
            "S", ("const", 1), "sub", ("const", 2), "max", "V", ("param", 0), ("const", 1), "mul", "add", "max",
            "S", ("const", 1), "gt",                                                             # S > GQ ?
            ("param", 0), ("const", 1), ("const", 3), ("param", 1), "sub", "S", ("const", 1), "sub", "mul", "add", "mul",
            ("const", 4), "select", "max", "end"]
    words = []
    for t in toks:
        words += [OPS[t[0]], t[1]] if isinstance(t, tuple) else [OPS[t]]
    params = np.array([[1. - .01 * (e + 1), max(.03 - .01 * e, 0.)] for e in range(E)])
    return dict(program=np.array(words, dtype=np.int32), consts=[1. + bonus, GQ, 0., 1., -np.inf], params=params)


def glwb_tables(years):
    """The synthetic mortality / penalty tables of oracle/shim/ref_driver.cpp (mode glwb) and what follows from
    them: survival R[0 .. years + 1], the operator's source g[n] = -(R[n+1] - R[n]), the per-event parameters
    [R[n], penalty[n-1]] of the event at year n = 1 .. years."""
    mortality = np.array([1. if n == years - 1 else .01 * (n + 1) for n in range(years)])
    penalty = np.array([max(.03 - .01 * n, 0.) for n in range(years)])
    R = np.ones(years + 2)
    for n in range(1, years + 1):
        R[n] = R[n - 1] * (1. - mortality[n - 1])
    R[years + 1] = R[years]
    source = np.array([-(R[n + 1] - R[n]) for n in range(years + 1)])
    params = np.array([[R[n], penalty[n - 1]] for n in range(1, years + 1)])
    return R, source, params


def glwb(years, n_steps, refine_times, r=.04, vol=.2, S0=100., management_rate=.015, withdrawal_rate=.05,
         bonus_rate=.06, scheme="bdf2"):
    """examples/experimental/glwb.cpp run(k) on the oracle: GLWBOperator (source term), an event at every year
    1 .. years (the last AT the initial time: a zero-length first step), zero initial condition."""
    S = refine(special_axis() * S0, refine_times)
    lo, di, up = bs_coeffs(S, r, vol, management_rate)
    _, source, params = glwb_tables(years)
    ev = glwb_event(years, bonus=bonus_rate, GQ=withdrawal_rate * S0)
    ev["params"] = params
    T = float(years)
    steps, n = schedule(T, T / n_steps, [float(k) for k in range(1, years + 1)])
    out = sweep(S, lo, di, up, np.zeros(len(S)), T, steps, n, scheme=scheme, event=ev, source=source)
    out["S"] = S
    out["n_steps"] = n
    return out


def forward_option(K, r, vol, T, n_steps, refine_times, q=0.0, call=False, american=False, scheme="rannacher",
                   alpha=0.0, E=0, dividend=None, dividend_kind="cash", dividend_first=False, exercise=False):
    """The forward-time classes on the oracle (ForwardConstantStepper(0, T, T / n_steps) + ForwardRannacher /
    CrankNicolson / BDFk, IterativeMethod.hpp:1494-1495): marched from the payoff at t = 0 up to T; events at
    T e / E, e = 1 .. E.  alpha > 0: local volatility alpha / S instead of vol."""
    S = vanilla_grid(K, None, refine_times)
    with np.errstate(divide="ignore"):
        v = alpha / S if alpha > 0 else vol
    lo, di, up = bs_coeffs(S, r, v, q)
    payoff = np.where(S < K, 0.0, S - K) if call else np.where(S < K, K - S, 0.0)
    steps, n = schedule(T, T / n_steps, [T / E * e for e in range(1, E + 1)], forward=True)
    out = sweep(S, lo, di, up, payoff, T, steps, n, scheme=scheme, american=american, obstacle=payoff,
                event_obstacle=(K - S) if (E > 0 and exercise) else None,
                dividend=None if (E == 0 or dividend is None) else (dividend_kind, dividend, dividend_first),
                forward=True)
    out["S"] = S
    out["n_steps"] = n
    return out


def bermudan_put(K, r, alpha, T, n_steps, E, refine_times, q=0.0, scheme="bdf2",
                 dividend=None, dividend_kind="cash", dividend_first=False, exercise=True):
    """The tutorial.cpp wiring on the oracle (local volatility alpha / S); optionally with
    the README's discrete dividends at the exercise dates."""
    S = refine(TUTORIAL_AXIS * (K / 100.0), refine_times)
    with np.errstate(divide="ignore"):
        vol = alpha / S
    lo, di, up = bs_coeffs(S, r, vol, q)
    payoff = np.maximum(K - S, 0.0)
    ev = [T / E * e for e in range(E)]
    steps, n = schedule(T, T / n_steps, ev)
    out = sweep(S, lo, di, up, payoff, T, steps, n, scheme=scheme,
                american=False, obstacle=payoff, event_obstacle=(K - S) if exercise else None,
                dividend=None if dividend is None else (dividend_kind, dividend, dividend_first))
    out["S"] = S
    out["n_steps"] = n
    return out


class Gmwb(C.Structure):
    _fields_ = [("n_w", C.c_int), ("n_a", C.c_int), ("W", C.POINTER(C.c_double)), ("A", C.POINTER(C.c_double)),
                ("n_q", C.c_int), ("q", C.POINTER(C.c_double)), ("n_z", C.c_int), ("zfrac", C.POINTER(C.c_double)),
                ("T", C.c_double), ("timesteps", C.c_int), ("r", C.c_double), ("eta", C.c_double),
                ("sigma", C.c_double), ("kfee", C.c_double), ("c", C.c_double),
                ("scaling_factor", C.c_double), ("tolerance", C.c_double), ("max_inner", C.c_int)]


GMWB_W_AXIS = np.array([
    0., 5., 10., 15., 20., 25., 30., 35., 40., 45., 50., 55., 60., 65., 70., 72.5, 75., 77.5, 80., 82., 84.,
    86., 88., 90., 91., 92., 93., 94., 95., 96., 97., 98., 99., 100., 101., 102., 103., 104., 105., 106.,
    107., 108., 109., 110., 112., 114., 116., 118., 120., 123., 126., 130., 135., 140., 145., 150., 160.,
    175., 200., 225., 250., 300., 500., 750., 1000.])      # examples/hjbqvi/gmwb.cpp:92-105


def uniform_axis(begin, end, points):
    dx = (end - begin) / (points - 1)                          # Axis::uniform, Core/Axis.hpp:219-226
    return np.array([begin + dx * i for i in range(points)])


def gmwb(refine_times=0, a_points=51, timesteps=32, control_points=3, T=10., r=.05, eta=0., sigma=.2,
         G=10., kfee=.1, c=0., w0=100., max_inner=1000):
    """The examples/hjbqvi/gmwb.cpp wiring on the oracle; HJBQVI::solve(refinement) doubles the
    time steps per level and refines the spatial and impulse-control grids (HJBQVI.hpp:596-634)."""
    W = refine((w0 / 100.) * GMWB_W_AXIS, refine_times)
    A = refine(uniform_axis(0., w0, a_points), refine_times)
    z = refine(uniform_axis(0., 1., control_points), refine_times)
    q = np.array([0., G])
    p = Gmwb()
    p.n_w, p.n_a, p.W, p.A = len(W), len(A), _dp(W), _dp(A)
    p.n_q, p.q, p.n_z, p.zfrac = len(q), _dp(q), len(z), _dp(z)
    p.T, p.timesteps = T, timesteps * 2 ** refine_times
    p.r, p.eta, p.sigma, p.kfee, p.c = r, eta, sigma, kfee, c
    p.scaling_factor, p.tolerance, p.max_inner = 1e-2, 1e-6, max_inner
    V = np.zeros(len(W) * len(A))
    mean_inner, n_steps = C.c_double(), C.c_int()
    status = lib().qpo_gmwb_solve(C.byref(p), _dp(V), C.byref(mean_inner), C.byref(n_steps))
    return dict(W=W, A=A, V=V.reshape(len(A), len(W)), mean_inner=mean_inner.value,
                n_steps=n_steps.value, status=status)


class Hjbqvi1d(C.Structure):
    _fields_ = [("n_x", C.c_int), ("x", C.POINTER(C.c_double)),
                ("n_q", C.c_int), ("q", C.POINTER(C.c_double)),
                ("n_z", C.c_int), ("z", C.POINTER(C.c_double)),
                ("T", C.c_double), ("timesteps", C.c_int),
                ("rho", C.c_double), ("sigma", C.c_double), ("a", C.c_double), ("b", C.c_double),
                ("m", C.c_double), ("lam", C.c_double), ("c", C.c_double),
                ("scaling_factor", C.c_double), ("tolerance", C.c_double), ("max_inner", C.c_int)]


def exchange_rate(x_axis, q_axis, z_axis, refine_times=0, timesteps=16, T=10., rho=.02, sigma=.3, a=.25, b=3.,
                  m=0., lam=1., c=.1, max_inner=1000):
    """The examples/hjbqvi/exchange_rate.cpp wiring on the oracle.  The coarse axes come from the caller
    (Axis::cluster goes through libm's asinh / sinh: the fixtures carry the reference's own ticks);
    HJBQVI::solve(refinement) doubles the time steps per level and refines all three grids."""
    X, Q, Z = (refine(np.asarray(ax, dtype=np.float64), refine_times) for ax in (x_axis, q_axis, z_axis))
    p = Hjbqvi1d()
    p.n_x, p.x, p.n_q, p.q, p.n_z, p.z = len(X), _dp(X), len(Q), _dp(Q), len(Z), _dp(Z)
    p.T, p.timesteps = T, timesteps * 2 ** refine_times
    p.rho, p.sigma, p.a, p.b, p.m, p.lam, p.c = rho, sigma, a, b, m, lam, c
    p.scaling_factor, p.tolerance, p.max_inner = 1e-2, 1e-6, max_inner
    V, Qs, Zs = np.zeros(len(X)), np.zeros(len(X)), np.zeros(len(X))
    mask = np.zeros(len(X), dtype=np.uint8)
    mean_inner, n_steps = C.c_double(), C.c_int()
    status = lib().qpo_hjbqvi1d_solve(C.byref(p), _dp(V), _dp(Qs), _dp(Zs),
                                      mask.ctypes.data_as(C.POINTER(C.c_ubyte)), C.byref(mean_inner), C.byref(n_steps))
    return dict(X=X, QA=Q, ZA=Z, V=V, Q=Qs, Z=Zs, mask=mask, mean_inner=mean_inner.value,
                n_steps=n_steps.value, status=status)


class Hjbqvi2d(C.Structure):
    _fields_ = [("n_w", C.c_int), ("n_b", C.c_int), ("W", C.POINTER(C.c_double)), ("B", C.POINTER(C.c_double)),
                ("n_q", C.c_int), ("q", C.POINTER(C.c_double)), ("n_z", C.c_int), ("zfrac", C.POINTER(C.c_double)),
                ("T", C.c_double), ("timesteps", C.c_int),
                ("rho", C.c_double), ("r", C.c_double), ("mu", C.c_double), ("sigma", C.c_double), ("gamma", C.c_double),
                ("lam", C.c_double), ("c", C.c_double), ("boundary", C.c_double),
                ("scaling_factor", C.c_double), ("tolerance", C.c_double), ("max_inner", C.c_int), ("max_sweeps", C.c_int)]


def optimal_consumption(refine_times=0, timesteps=32, control_points=16, spatial_points=20, T=40., rho=.1, r=.07, mu=.11,
                        sigma=.3, gamma=.3, lam=.1, c=.05, q_max=100., boundary=200., max_inner=1000, max_sweeps=5000,
                        spatial_points_b=None):
    """The examples/hjbqvi/optimal_consumption.cpp wiring on the oracle."""
    W = refine(uniform_axis(0., boundary, spatial_points), refine_times)
    B = W.copy() if spatial_points_b is None else refine(uniform_axis(0., boundary, spatial_points_b), refine_times)
    q = refine(uniform_axis(0., q_max, control_points), refine_times)
    z = refine(uniform_axis(0., 1., control_points), refine_times)
    p = Hjbqvi2d()
    p.n_w, p.n_b, p.W, p.B = len(W), len(B), _dp(W), _dp(B)
    p.n_q, p.q, p.n_z, p.zfrac = len(q), _dp(q), len(z), _dp(z)
    p.T, p.timesteps = T, timesteps * 2 ** refine_times
    p.rho, p.r, p.mu, p.sigma, p.gamma, p.lam, p.c, p.boundary = rho, r, mu, sigma, gamma, lam, c, boundary
    p.scaling_factor, p.tolerance, p.max_inner, p.max_sweeps = 1e-2, 1e-6, max_inner, max_sweeps
    n = len(W) * len(B)
    V, Qs, Zs = np.zeros(n), np.zeros(n), np.zeros(n)
    mask = np.zeros(n, dtype=np.uint8)
    mean_inner, n_steps = C.c_double(), C.c_int()
    status = lib().qpo_hjbqvi2d_solve(C.byref(p), _dp(V), _dp(Qs), _dp(Zs), mask.ctypes.data_as(C.POINTER(C.c_ubyte)),
                                      C.byref(mean_inner), C.byref(n_steps))
    shape = (len(B), len(W))
    return dict(W=W, B=B, QA=q, ZA=z, V=V.reshape(shape), Q=Qs.reshape(shape), Z=Zs.reshape(shape), mask=mask.reshape(shape),
                mean_inner=mean_inner.value, n_steps=n_steps.value, status=status)


JUMP_AXIS_EXTRA = 1000.0   # examples/jump_diffusion.cpp:166: + S_0 * Axis{1000.}


def jump_setup(S, density="lognormal", params=(-0.1, 0.45, 0.0)):
    """(x0, dx, weights, kappa) as BlackScholesJumpDiffusion computes them at construction."""
    S = np.ascontiguousarray(S, dtype=np.float64)
    nfft = lib().qpo_jump_nfft(len(S))
    w = np.zeros(nfft)
    x0, dx, kappa = C.c_double(), C.c_double(), C.c_double()
    p = list(params) + [0.0] * (3 - len(params))
    lib().qpo_jump_setup(len(S), _dp(S), 1 if density == "kou" else 0, C.c_double(p[0]),
                         C.c_double(p[1]), C.c_double(p[2]), C.byref(x0), C.byref(dx), _dp(w),
                         C.byref(kappa))
    return x0.value, dx.value, w, kappa.value


def jump_call(K, r, vol, T, lam, n_steps, refine_times, q=0.0, density="lognormal",
              params=(-0.1, 0.45), scheme="bdf1", solver=0):
    """The jump_diffusion.cpp wiring on the oracle (European call)."""
    sp = special_axis()
    coarse = axis_union(axis_union(sp * K, sp * K), np.array([K * JUMP_AXIS_EXTRA]))
    S = refine(coarse, refine_times)
    x0, dx, w, kappa = jump_setup(S, density, params)
    lo, di, up = bs_coeffs(S, r, vol, q, lam=lam, kappa=kappa)
    payoff = np.where(S < K, 0.0, S - K)
    steps, n = schedule(T, T / n_steps)
    out = sweep(S, lo, di, up, payoff, T, steps, n, scheme=scheme,
                jump=dict(weights=w, x0=x0, dx=dx, **{"lambda": lam}), solver=solver)
    out.update(S=S, n_steps=n, x0=x0, dx=dx, weights=w, kappa=kappa)
    return out


def hjb_straddle(K, r_l, r_b, vol, T, n_steps, refine_times, q=0.0, long_position=False,
                 scheme="bdf2", solver=0):
    """The unequal_borrowing_lending_rates.cpp wiring on the oracle."""
    S = vanilla_grid(K, K, refine_times)
    rows = [bs_coeffs(S, r, vol, q) for r in (r_l, r_b)]
    plo, pdi, pup = [np.stack([rw[k] for rw in rows]) for k in range(3)]
    payoff = np.where(S < K, K - S, S - K)
    steps, n = schedule(T, T / n_steps)
    out = sweep(S, rows[0][0], rows[0][1], rows[0][2], payoff, T, steps, n, scheme=scheme,
                policy_rows=(plo, pdi, pup), policy_max=long_position, solver=solver)
    out["S"] = S
    out["n_steps"] = n
    return out


# --------------------------------------------------------------------------
# Runner for the compiled reference (oracle/_ref/qpde_ref)
# --------------------------------------------------------------------------

def have_ref():
    return os.path.exists(REF_PATH)


def run_ref(mode, **kw):
    """Runs oracle/_ref/qpde_ref and parses its dump into numpy arrays."""
    args = [REF_PATH, mode] + ["%s=%s" % (k, repr(v) if isinstance(v, float) else v)
                               for k, v in kw.items()]
    return parse_ref_dump(subprocess.check_output(args).decode())


def parse_ref_dump(txt):
    """The dump format of oracle/shim/ref_driver.cpp (hex floats) -> numpy arrays."""
    toks = txt.split()
    out = {}
    i = 0
    while i < len(toks):
        key = toks[i]
        if key in ("value", "mean_inner", "mean_solver"):
            out[key] = float.fromhex(toks[i + 1]); i += 2
        elif key in ("seconds",):
            out[key] = float(toks[i + 1]); i += 2
        elif key in ("steps",):
            out[key] = int(toks[i + 1]); i += 2
        else:
            n = int(toks[i + 1])
            vals = toks[i + 2:i + 2 + n]
            if key in ("inner", "mask"):
                out[key] = np.array([int(x) for x in vals], dtype=np.int64)
            else:
                out[key] = np.array([float.fromhex(x) for x in vals])
            i += 2 + n
    return out
