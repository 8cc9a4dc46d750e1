// qpde_interleaved.cu — kernel family "INTERLEAVED": one thread per contract,
// every per-node array stored as [C / 32][N][32] in HBM (node-major inside a
// block of 32 contracts): the 32 lanes of a warp touch 32 consecutive doubles
// at every node, and the rows a warp walks are one sequential stream per array
// (a [N][C] layout made every 256-byte row segment its own DRAM page visit).
// The whole sweep (all time steps, all penalty iterations) is one launch;
// nothing returns to the host in between.
//
// Replaces, per contract, the reference call stack of SURVEY.md §3.1/§3.2:
//   TimeIteration<false>::iterateUntilDone     Core/IterativeMethod.hpp:1163-1204,1259-1317
//   ToleranceIteration + relativeError         Core/IterativeMethod.hpp:1125-1137,1211-1254
//   PenaltyMethodDifference::onIterationStart  Core/PenaltyMethod.hpp:37-69,96-119
//   Rannacher / CrankNicolson / BDFOne / BDFTwo A(), b()
//   SparseLUSolver::initialize/solve           Bindings/Eigen.hpp:143-165
//     -> Thomas elimination (no pivoting: every system here is a strictly
//        diagonally dominant M-matrix)
//
// Algorithmic HBM bytes per inner solve and node: reads lo, di,
// up, lb, obstacle, x_old (48 B) + writes cp, dp (16 B) in the forward pass;
// reads cp, dp, x_old (24 B) + writes x_new (8 B) in the backward pass.
#include <math_constants.h>

#include "qpde_kernels.cuh"
#include "qpde_partition.cuh"   // rcp()

namespace qpde {

constexpr int ROWS = 8;   // rows whose loads are in flight together in the per-thread elimination loops

// ---------------------------------------------------------------------------
// Setup: refined grid, operator rows, initial condition, obstacles.
// One thread per contract walks its nodes; lanes stay coalesced on [N][C].
// ---------------------------------------------------------------------------
__global__ void k_interleaved_setup(DeviceBatch b, InterleavedWork w) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if(c >= b.C) return;
	const int N = b.N;
	constexpr size_t C = QPDE_WARP_ROW;                       // row stride inside a warp's block of contracts
	const size_t base = interleaved_base(c, N);               // [C / 32][N][32] layout
	const size_t NC = (size_t)N * (size_t)w.ld;               // one array
	const double *axis = b.axis + (size_t)c * b.axis_stride;
	const double K = b.strike[c];
	const double r = b.r[c], q = b.q[c], sig = b.sigma[c];

	double Sm = 0., Si = refined_node(axis, b.refine, 0);
	double Sp = N > 1 ? refined_node(axis, b.refine, 1) : Si;
	for(int i = 0; i < N; ++i) {
		const size_t at = (size_t)i * C + base;
		const double v_i = b.vol_model == QPDE_VOL_NODES
			? b.sigma_nodes[(size_t)c * b.sigma_nodes_stride + i]
			: vol_value(b.vol_model, sig, Si);
		w.S[at] = Si;
		if(b.n_controls > 0) {
			// one set of operator rows per control value of the interest rate
			for(int k = 0; k < b.n_controls; ++k) {
				const double rk = b.controls[(size_t)c * b.controls_stride + k];
				const Row row = bs_row(i, N, Sm, Si, Sp, rk, v_i, q);
				const size_t ak = (size_t)k * NC + at;
				w.lo[ak] = row.lo; w.di[ak] = row.di; w.up[ak] = row.up;
			}
		} else {
			const Row row = bs_row(i, N, Sm, Si, Sp, r, v_i, q);
			w.lo[at] = row.lo; w.di[at] = row.di; w.up[at] = row.up;
		}
		w.x[at] = b.payoff_kind == QPDE_PAYOFF_ARRAY
			? b.payoff[(size_t)c * b.payoff_stride + i]
			: payoff_value(b.payoff_kind, Si, K);
		if(w.obst) {
			w.obst[at] = b.obstacle_kind == QPDE_PAYOFF_ARRAY
				? b.obstacle[(size_t)c * b.obstacle_stride + i]
				: payoff_value(b.obstacle_kind, Si, K);
		}
		if(w.evobst) w.evobst[at] = payoff_value(b.exercise_kind, Si, K);   // allocated only with an exercise transform
		Sm = Si; Si = Sp;
		Sp = i + 2 < N ? refined_node(axis, b.refine, i + 2) : Sp;
	}
}

// ---------------------------------------------------------------------------
// BDF3 .. BDF6 (BDF.hpp:139-470), kept out of line so that the sweep's hot path keeps its
// registers: coefficients of one step, and the right-hand side
//   b = (c0 iterand(0) - c1 iterand(1) + c2 iterand(2) - ... + op.b) hh
// with iterand(m >= 1) in slot (head + m - 1) % slots of the history ring behind xp.
// ---------------------------------------------------------------------------
struct BdfHigh { double c[6]; double hh; };

__device__ __noinline__ BdfHigh bdf_high_setup(int k, double t, double time1, double time0,
		double th0, double th1, double th2, double th3, int forward) {
	BdfHigh f;
	double g[6] = { time1 - t, time0 - t, th0 - t, th1 - t, th2 - t, th3 - t };
	if(forward) for(int m = 0; m < 6; ++m) g[m] = 0. - g[m];      // t - time(m): exact negation
	for(int m = 0; m < 6; ++m) f.c[m] = 0.;
	bdf_coefficients(k, g, f.c, &f.hh);
	return f;
}

__device__ __noinline__ void bdf_high_rhs(int k, BdfHigh f, int N, size_t NC, const double *x, const double *xp,
		int head, int slots, double *lb) {
	constexpr size_t C = QPDE_WARP_ROW;
	for(int i = 0; i < N; ++i) {
		const size_t at = (size_t)i * C;
		double acc = f.c[0] * x[at];
		for(int m = 1; m < k; ++m) {
			const double term = f.c[m] * xp[(size_t)((head + m - 1) % slots) * NC + at];
			acc = (m & 1) ? acc - term : acc + term;
		}
		lb[at] = (acc + 0.) * f.hh;
	}
}

// ---------------------------------------------------------------------------
// The sweep.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_interleaved_sweep(DeviceBatch b, InterleavedWork w, DeviceResult out) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if(c >= b.C) return;
	const int N = b.N;
	constexpr size_t C = QPDE_WARP_ROW;                       // row stride inside a warp's block of contracts
	const size_t base = interleaved_base(c, N);               // [C / 32][N][32] layout
	const double tol = b.tolerance, scale = b.scale;
	const double pen = 1. / b.penalty_tolerance; // PenaltyMethod.hpp:85
	const double *lo = w.lo + base, *di = w.di + base, *up = w.up + base;
	double *x = w.x + base, *xp = w.xprev + base, *lb = w.lb + base;
	double *cp = w.cp + base, *dp = w.dp + base;
	const double *ob = w.obst ? w.obst + base : nullptr;
	const double *evob = w.evobst ? w.evobst + base : nullptr;

	const bool policy = b.n_controls > 0;
	const bool iterate = b.american || policy;   // a ToleranceIteration wraps every step
	const size_t NC = (size_t)N * (size_t)w.ld;               // one array (one set of operator rows)
	Replay rp; rp.start(b.T[c], b.dt[c], b.n_exercise, b.forward, b.event_times);
	NodeState ns; ns.start(b.scheme);
	double time1 = rp.initial(), time0 = rp.initial(); // times of iterand(0), iterand(1)
	Gamma gA; gA.kind = STEP_IMPLICIT; gA.h = 0.; // what the factorised A was built with
	int n_steps = 0, inner_total = 0, status = 0;
	bool more = true;
	// BDF3 .. BDF6: iterand(1) .. iterand(order - 1) live in a ring of order - 1 arrays behind
	// w.xprev; iterand(m) is slot (head + m - 1) % slots.  th[] = times of iterand(2) .. iterand(5).
	const int order = bdf_order(b.scheme);
	const int slots = order > 2 ? order - 1 : 1;
	int head = 0;
	double th0 = rp.initial(), th1 = rp.initial(), th2 = rp.initial(), th3 = rp.initial();

	// ReverseVariableStepper (Core/Stepper.hpp:60-126)
	const bool variable = b.variable_target != nullptr;
	const double vtarget = variable ? b.variable_target[c] : 0.;
	double vdt = b.dt[c];

	while(more) {
		qpde_step st;
		if(variable) {
			if(n_steps >= b.max_steps) { status = 3; break; }      // output capacity
			rp.dt_const = vdt;                                      // VariableStepper::timestep()
		}
		more = rp.next(st);
		const double t = st.t;
		const int kind = ns.kind();
		const double h0 = rp.lapse(time1, t);
		Gamma gb; gb.kind = kind; gb.h = h0;
		double c1 = 0., c0 = 0., den = 1.;
		if(kind == STEP_BDF2) {
			const double h1 = rp.lapse(time1, t), hh0 = rp.lapse(time0, t);     // BDF.hpp:68-73
			gb.h = div_step(hh0 * h1, hh0 + h1);
			c1 = div_step(hh0, h1); c0 = div_step(h1, hh0); den = hh0 - h1;
		}
		const double rden = div_step(1., den);
		if(kind >= STEP_BDF3) {
			// BDF.hpp:139-470: A = I + hh op.A; the right-hand side goes to lb here
			const BdfHigh f = bdf_high_setup(kind - STEP_BDF3 + 3, t, time1, time0, th0, th1, th2, th3, b.forward);
			gb.h = f.hh;
			bdf_high_rhs(kind - STEP_BDF3 + 3, f, N, NC, x, xp, head, slots, lb);
		}
		// IterativeMethod.hpp:906: re-assemble A unless the root says it is the same
		// (a penalty or policy node in the tree keeps LinearSystem's default: never)
		if(iterate || !ns.initialized || !ns.a_same(st.same_dt)) gA = gb;
		double *it1 = xp + (size_t)head * NC;       // iterand(1)
		const double hsA = gA.factor(), hsb = gb.factor();

		// ---- lb = root.b(t) of the discretisation node ------------------
		// x holds iterand(0) = V^n, xp holds iterand(1) for BDF2.
		if(kind < STEP_BDF3) {
			// b(t) of the operator: 0, or source_table[floor(t)] S (GLWBOperator::b, glwb.cpp:40-44)
			double gsrc = 0.;
			if(b.source_table) {
				int nidx = (int)floor(t);
				nidx = nidx < 0 ? 0 : (nidx > b.n_source - 1 ? b.n_source - 1 : nidx);
				gsrc = b.source_table[nidx];
			}
			// blocks of ROWS rows with all loads issued first (see the elimination loop)
			double vm = 0.;
			for(int ib = 0; ib < N; ib += ROWS) {
				double vx[ROWS + 1], vlo[ROWS], vdi[ROWS], vup[ROWS], vxp[ROWS];
#pragma unroll
				for(int u = 0; u < ROWS; ++u) {
					const size_t at = (size_t)(ib + u < N ? ib + u : N - 1) * C;
					vx[u] = x[at];
					vxp[u] = kind == STEP_BDF2 ? it1[at] : 0.;
					const bool cn = kind != STEP_IMPLICIT && kind != STEP_BDF2;
					vlo[u] = cn ? lo[at] : 0.; vdi[u] = cn ? di[at] : 0.; vup[u] = cn ? up[at] : 0.;
				}
				vx[ROWS] = ib + ROWS < N ? x[(size_t)(ib + ROWS) * C] : 0.;
#pragma unroll
				for(int u = 0; u < ROWS; ++u) {
					const int i = ib + u;
					if(i >= N) break;
					const double vi = vx[u], vp = i + 1 < N ? vx[u + 1] : 0.;
					double val;
					if(kind == STEP_IMPLICIT) {
						val = b.source_table ? vi + h0 * (gsrc * w.S[base + (size_t)i * C]) : vi + 0. * h0;
					} else if(kind == STEP_BDF2) {
						val = (div_by(c1 * vi - c0 * vxp[u], den, rden) + (b.source_table ? gsrc * w.S[base + (size_t)i * C] : 0.)) * gb.h;
					} else {
						// (I - A h/2) v0, row-major SpMV from zero in column order
						double acc = 0.;
						if(i > 0) acc += (0. - vlo[u] * hsb) * vm;
						acc += (1. - vdi[u] * hsb) * vi;
						if(i < N - 1) acc += (0. - vup[u] * hsb) * vp;
						val = acc + 0.;
					}
					lb[(size_t)i * C] = val;
					vm = vi;
				}
			}
		}
		// history push: iterand(1) <- iterand(0).  For BDF2 the previous
		// iterand must survive the solve, so copy it before x is overwritten.
		// Policy iteration under a theta scheme: Rannacher::_b2 / CrankNicolson::b form (I - (1 - theta) h A(t0)) V^n with
		// the policy node's A, i.e. under the controls of the CURRENT inner iteration (Rannacher.hpp:60-73,
		// CrankNicolson.hpp:64-78, PolicyIteration.hpp:107-115): the right-hand side is re-formed from V^n at every
		// control update, so V^n must survive the solves.
		const bool policy_theta = policy && (kind == STEP_CN_R || kind == STEP_CN_T);
		if(order >= 2 || variable || policy_theta) {   // the variable stepper needs V^n after the solve, too
			head = head == 0 ? slots - 1 : head - 1;    // the oldest slot becomes iterand(1)
			it1 = xp + (size_t)head * NC;
			for(int i = 0; i < N; ++i) it1[(size_t)i * C] = x[(size_t)i * C];
		}

		int its = 0;
		bool again;
		do {
			// ---- forward elimination of (lA + P) x = lb + P obstacle ------
			double cprev = 0., dprev = 0.;
			if(!policy) {
				// Every thread walks its own contract: the rows form a serial chain, and a load per
				// row would expose the full HBM latency 2 N times per solve.  The rows are taken in
				// blocks of ROWS: all loads of a block are issued first, then the chain runs on registers.
				for(int ib = 0; ib < N; ib += ROWS) {
					double vlo[ROWS], vdi[ROWS], vup[ROWS], vlb[ROWS], vob[ROWS], vx[ROWS];
#pragma unroll
					for(int u = 0; u < ROWS; ++u) {
						const size_t at = (size_t)(ib + u < N ? ib + u : N - 1) * C;
						vlo[u] = lo[at]; vdi[u] = di[at]; vup[u] = up[at]; vlb[u] = lb[at];
						vob[u] = b.american ? ob[at] : 0.; vx[u] = b.american ? x[at] : 0.;
					}
#pragma unroll
					for(int u = 0; u < ROWS; ++u) {
						if(ib + u >= N) break;
						const size_t at = (size_t)(ib + u) * C;
						double a_i = vlo[u] * hsA;            // Gamma::off(l) == l * factor(), without the switch per row
						double b_i = 1. + vdi[u] * hsA;
						double c_i = vup[u] * hsA;
						double rhs = vlb[u];
						if(b.american) {
							// PenaltyMethod.hpp:45,61-68: predicate on the previous iterand
							// (scale * (x - payoff)) < 0  <=>  x < payoff
							if(vx[u] < vob[u]) { b_i = b_i + pen * 1.; rhs = rhs + pen * vob[u]; }
						}
						const double denom = b_i - a_i * cprev;
						const double inv = rcp(denom);        // < 1 ulp; no division sequence on the serial chain
						cprev = c_i * inv;
						dprev = (rhs - a_i * dprev) * inv;
						cp[at] = cprev; dp[at] = dprev;
					}
				}
			} else {
				double xm = 0., xi = x[0], xq = N > 1 ? x[C] : 0.;   // previous iterate around node i
				for(int i = 0; i < N; ++i) {
					const size_t at = (size_t)i * C;
					size_t rows = at;
					if(policy) {
						// PolicyIteration::onIterationStart (PolicyIteration.hpp:54-105):
						// candidate = (A(q) x - b(q))_i per control value, strict order
						double best = b.policy_max ? -CUDART_INF : CUDART_INF;
						for(int k = 0; k < b.n_controls; ++k) {
							const size_t ak = (size_t)k * NC + at;
							double acc = 0.;
							if(i > 0 && i < N - 1) acc += lo[ak] * xm;
							acc += di[ak] * xi;
							if(i > 0 && i < N - 1) acc += up[ak] * xq;
							const double cand = acc - 0.;
							if(b.policy_max ? cand > best : cand < best) { best = cand; rows = ak; }
						}
						xm = xi; xi = xq;
						xq = i + 2 < N ? x[at + 2 * C] : 0.;
					}
					double a_i = gA.off(lo[rows]);
					double b_i = 1. + gA.off(di[rows]);
					double c_i = gA.off(up[rows]);
					double rhs = lb[at];
					if(policy_theta) {
						// row i of (I - (1 - theta) h A(q*)) V^n, SpMV from zero in column order; the row of A(q*)
						// carries the control chosen at its own node
						double acc = 0.;
						if(i > 0) acc += (0. - gb.off(lo[rows])) * it1[at - C];
						acc += (1. - gb.off(di[rows])) * it1[at];
						if(i < N - 1) acc += (0. - gb.off(up[rows])) * it1[at + C];
						rhs = acc + 0.;
					}
					if(b.american) {
						// PenaltyMethod.hpp:45,61-68: predicate on the previous iterand
						// (scale * (x - payoff)) < 0  <=>  x < payoff
						const double obi = ob[at];
						if(x[at] < obi) { b_i = b_i + pen * 1.; rhs = rhs + pen * obi; }
					}
					const double denom = b_i - a_i * cprev;
					const double inv = 1. / denom;
					cprev = c_i * inv;
					dprev = (rhs - a_i * dprev) * inv;
					cp[at] = cprev; dp[at] = dprev;
				}
			}
			// ---- back substitution + relativeError --------------------------
			double xn = 0.;
			bool notdone = false;
			for(int ib = N - 1; ib >= 0; ib -= ROWS) {
				double vcp[ROWS], vdp[ROWS], vxo[ROWS];
#pragma unroll
				for(int u = 0; u < ROWS; ++u) {
					const size_t at = (size_t)(ib - u >= 0 ? ib - u : 0) * C;
					vcp[u] = cp[at]; vdp[u] = dp[at]; vxo[u] = iterate ? x[at] : 0.;
				}
#pragma unroll
				for(int u = 0; u < ROWS; ++u) {
					if(ib - u < 0) break;
					const size_t at = (size_t)(ib - u) * C;
					xn = vdp[u] - vcp[u] * xn;
					if(iterate) {
						const double xo = vxo[u];
						const double fa = fabs(xn), fb = fabs(xo);
						double d = fa < fb ? fb : fa;
						d = scale < d ? d : scale;
						// IterativeMethod.hpp:1131-1135, decided in product form; the quotient is formed only
						// in the rounding band around the threshold (same decisions, no division per row)
						const double diff = fabs(xn - xo), bound = tol * d;
						if(diff > bound * (1. + 8e-15)) notdone = true;
						else if(diff >= bound * (1. - 8e-15) && diff / d > tol) notdone = true;
					}
					x[at] = xn;
				}
			}
			++its;
			again = iterate && notdone;
			if(again && its >= b.max_inner) { status = 1; again = false; }
		} while(again);

		if(out.inner && n_steps < b.max_steps) {
			out.inner[(size_t)c * out.inner_stride + n_steps] = (uint8_t)(its > 255 ? 255 : its);
		}
		inner_total += its;
		++n_steps;
		th3 = th2; th2 = th1; th1 = th0; th0 = time0;
		time0 = time1; time1 = t;
		ns.end_step();

		if(variable && more) {
			// dt *= target / (relativeError(iterand(0), iterand(1), scale) + epsilon), Stepper.hpp:68-82
			double err = 0.;
			for(int i = 0; i < N; ++i) {
				const double a = x[(size_t)i * C], v = it1[(size_t)i * C];
				const double fa = fabs(a), fb = fabs(v);
				double dn = fa < fb ? fb : fa;
				dn = b.variable_scale < dn ? dn : b.variable_scale;
				const double e = fabs(a - v) / dn;
				err = e > err ? e : err;
			}
			vdt *= vtarget / (err + DBL_EPSILON);
		}

		if(st.event_after) {
			// PenaltyMethod::constraintMask (PenaltyMethod.hpp:127-139) reads the tolerance iteration's last
			// iterand, which the events below do not touch (they transform the stepper's): with user events
			// the mask is taken here, before them; the last step of a sweep always ends in an event
			if(out.active && ob && b.n_exercise > 0) {
				for(int i = 0; i < N; ++i) {
					const double pred = (0. + 1. * x[(size_t)i * C]) - ob[(size_t)i * C];
					out.active[(size_t)c * out.active_stride + i] = pred < 0. ? 1 : 0;
				}
			}
			// Events at one time go in descending add() order (IterativeMethod.hpp:1340-1352)
			for(int ue = 0; ue < st.user_events; ++ue) {
				if(b.event_program) {
					// a general Event<1> transform (QPDE_EV_* program): V is the interpolant of the pre-event
					// solution, so that goes to the right-hand-side array first (free between steps)
					const int ev_index = b.forward ? rp.e - st.user_events + ue : rp.e + st.user_events - ue;
					const double *const ev_params = b.event_params + (size_t)c * b.event_params_stride + (size_t)ev_index * b.n_event_params;
					const double *S = w.S + base;
					double *const old = w.lb + base;
					for(int i = 0; i < N; ++i) old[(size_t)i * C] = x[(size_t)i * C];
					const double Sfirst = S[0], Slast = S[(size_t)(N - 1) * C];
					auto V_interp = [&] (double at) -> double {
						// linearInterpolationData + PiecewiseLinear::interpolate (Core/Interpolant.hpp:153-181,331-374)
						int k; double wgt;
						if(at <= Sfirst) { k = 0; wgt = 1.; }
						else if(at >= Slast) { k = N - 2; wgt = 0.; }
						else {
							int lo = 0, hi = N - 2;
							while(lo < hi) { const int mid = (lo + hi + 1) >> 1; if(S[(size_t)mid * C] <= at) lo = mid; else hi = mid - 1; }
							k = lo;
							wgt = (S[(size_t)(k + 1) * C] - at) / (S[(size_t)(k + 1) * C] - S[(size_t)k * C]);
						}
						double sum = 0.;
						if(wgt > DBL_EPSILON) sum += wgt * old[(size_t)k * C];
						if(1. - wgt > DBL_EPSILON) sum += (1. - wgt) * old[(size_t)(k + 1) * C];
						return sum;
					};
					for(int i = 0; i < N; ++i) x[(size_t)i * C] = event_eval(b.event_program, b.event_consts, ev_params, S[(size_t)i * C], V_interp);
					continue;
				}
				for(int op = 0; op < 2; ++op) {
					const bool is_dividend = (op == 0) == (b.dividend_first != 0);
					if(is_dividend && b.dividend_kind != QPDE_DIVIDEND_NONE) {
						// Event<1>::_doEvent with V(max(S - D, 0)) or V((1 - D) S) through the
						// piecewise-linear interpolant (Core/Event.hpp:100-112, Interpolant.hpp:153-181,
						// 331-374).  The shifted price never exceeds S_i, so the nodes are updated in
						// place from the top down, and the bracketing interval only moves down.
						const double D = b.dividend[c];
						const double *S = w.S + base;
						const double Sfirst = S[0], Slast = S[(size_t)(N - 1) * C];
						int idx = N - 2;
						for(int i = N - 1; i >= 0; --i) {
							const double Si = S[(size_t)i * C];
							const double at = b.dividend_kind == QPDE_DIVIDEND_CASH ? (Si - D > 0. ? Si - D : 0.) : (1. - D) * Si;
							int k; double wgt;
							if(at <= Sfirst) { k = 0; wgt = 1.; }
							else if(at >= Slast) { k = N - 2; wgt = 0.; }
							else {
								while(idx > 0 && S[(size_t)idx * C] > at) --idx;      // S[idx] <= at < S[idx + 1]
								k = idx;
								wgt = (S[(size_t)(k + 1) * C] - at) / (S[(size_t)(k + 1) * C] - S[(size_t)k * C]);
							}
							double sum = 0.;
							if(wgt > DBL_EPSILON) sum += wgt * x[(size_t)k * C];
							if(1. - wgt > DBL_EPSILON) sum += (1. - wgt) * x[(size_t)(k + 1) * C];
							x[(size_t)i * C] = sum;
						}
					}
					if(!is_dividend && evob) {
						for(int i = 0; i < N; ++i) {
							const size_t at = (size_t)i * C;
							const double ex = evob[at], v = x[at];
							x[at] = v > ex ? v : ex;       // QuantPDE::max, EarlyInclude.hpp:48
						}
					}
				}
			}
			ns.after_event();
		}
	}

	// ---- outputs (contract-major) ------------------------------------------
	if(out.V) for(int i = 0; i < N; ++i) out.V[(size_t)c * out.V_stride + i] = x[(size_t)i * C];
	if(out.S) for(int i = 0; i < N; ++i) out.S[(size_t)c * out.S_stride + i] = w.S[(size_t)i * C + base];
	if(out.active && !(ob && b.n_exercise > 0)) {
		for(int i = 0; i < N; ++i) {
			const double pred = ob ? (0. + 1. * x[(size_t)i * C]) - ob[(size_t)i * C] : 0.;
			out.active[(size_t)c * out.active_stride + i] = pred < 0. ? 1 : 0;
		}
	}
	if(out.value) {
		// PiecewiseLinear::interpolate (Core/Interpolant.hpp:153-181,331-374)
		const double *S = w.S + base;
		const double s0 = b.spot ? b.spot[c] : b.strike[c];
		int idx; double wgt;
		if(s0 <= S[0]) { idx = 0; wgt = 1.; }
		else if(s0 >= S[(size_t)(N - 1) * C]) { idx = N - 2; wgt = 0.; }
		else {
			int l = 0, hgh = N - 2, mid = 0; wgt = 0.;
			while(l <= hgh) {
				mid = (l + hgh) / 2;
				if(s0 < S[(size_t)mid * C]) hgh = mid - 1;
				else if(s0 >= S[(size_t)(mid + 1) * C]) l = mid + 1;
				else {
					wgt = (S[(size_t)(mid + 1) * C] - s0)
						/ (S[(size_t)(mid + 1) * C] - S[(size_t)mid * C]);
					break;
				}
			}
			idx = mid;
		}
		double sum = 0.;
		if(wgt > DBL_EPSILON) sum += wgt * x[(size_t)idx * C];
		if(1. - wgt > DBL_EPSILON) sum += (1. - wgt) * x[(size_t)(idx + 1) * C];
		out.value[c] = sum;
	}
	if(out.n_steps) out.n_steps[c] = n_steps;
	if(out.inner_total) out.inner_total[c] = inner_total;
	if(out.computed) out.computed[c] = inner_total;          // this sweep runs every inner iteration it counts
	if(out.status) out.status[c] = status;
}

void launch_interleaved_setup(const DeviceBatch &b, const InterleavedWork &w, cudaStream_t stream, long long *launches) {
	const int threads = b.C >= 148 * 512 ? 128 : (b.C >= 148 * 128 ? 64 : 32);
	const int blocks = (b.C + threads - 1) / threads;
	k_interleaved_setup<<<blocks, threads, 0, stream>>>(b, w);
	*launches += 1;
}

void launch_interleaved(const DeviceBatch &b, const InterleavedWork &w,
		const DeviceResult &out, cudaStream_t stream, long long *launches) {
	// small CTAs: a batch of a few thousand contracts still reaches every SM
	const int threads = b.C >= 148 * 512 ? 128 : (b.C >= 148 * 128 ? 64 : 32);
	const int blocks = (b.C + threads - 1) / threads;
	launch_interleaved_setup(b, w, stream, launches);
	k_interleaved_sweep<<<blocks, threads, 0, stream>>>(b, w, out);
	*launches += 1;
}

} // namespace qpde
