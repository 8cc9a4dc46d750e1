// qpde_resident.cu — kernel family "RESIDENT": one CTA per contract, the whole
// backward sweep (every time step, every penalty iteration) inside one launch
// with the contract's state held in registers (M grid rows per thread) and
// shared memory.  HBM is touched only to read the contract's parameters
// (~0.3 KB) and to write its results.
//
// The tridiagonal system of each (inner) iteration is solved by a three-level
// partition that follows the hardware hierarchy:
//   level 1  thread : M-1 interior rows eliminated serially (Thomas), leaving
//                     one separator unknown per thread (its first row) and two
//                     spike vectors;
//   level 2  warp   : the separator unknowns of lanes 1..31 by parallel cyclic
//                     reduction, 5 stages of warp shuffles; two more spikes;
//   level 3  CTA    : one separator per warp (lane 0), a W x W tridiagonal
//                     system (W <= 8) that every warp solves redundantly by a
//                     3-stage cyclic reduction across its lanes 0..W-1 from a
//                     shared-memory hand-over.
// Everything that depends only on the matrix (multipliers, reciprocals, spike
// vectors, reduction coefficients of levels 2 and 3) is cached in registers and
// recomputed only when the penalty active set or the step size changes; an
// ordinary inner iteration is right-hand-side work and two block barriers.
//
// Reference call stack replaced (per contract), SURVEY.md §3.1/§3.2:
//   TimeIteration<false>::iterateUntilDone   Core/IterativeMethod.hpp:1163-1204,1259-1317
//   ToleranceIteration / relativeError       Core/IterativeMethod.hpp:1125-1137,1211-1254
//   PenaltyMethodDifference                  Core/PenaltyMethod.hpp:37-69,96-119,127-139
//   Rannacher / CrankNicolson / BDFOne / BDFTwo   Core/Rannacher.hpp, CrankNicolson.hpp, BDF.hpp
//   BlackScholes<1,0>::A                     Modules/Operators/BlackScholes.hpp:136-226
//   SparseLUSolver::initialize / solve       Bindings/Eigen.hpp:143-165
//   Event<1>::_doEvent (exercise)            Core/Event.hpp:100-112
#include <math_constants.h>

#include <cstdlib>
#include <cstring>

#include "qpde_kernels.cuh"
#include <cooperative_groups.h>

#include "qpde_partition.cuh"

namespace cg = cooperative_groups;

namespace qpde {

namespace {

constexpr unsigned FULL = FULL_MASK;
constexpr int N_ROW_ARRAYS = 11; // S, lo, di, up, a, bd, c, na, me, nc, lb

// Fixed part of the shared memory of one CTA.  The per-row arrays ([M][PT]
// each, laid out [j][thread] so that a warp reading "row j of every thread"
// touches consecutive doubles) and xlast / xfirst ([PT + 1] each) follow it in
// dynamic shared memory.
struct Smem {
	PartitionSmem ps;   // hand-over between the warps of the partition solve
	// the scalar "tail" equation (row N-1 when N = P M + 1); touched only by
	// its owner, the last thread
	struct Tail { double S, di, x, pz, bd, me, lb, invb, pq, vprev, c, vn; } tail;
	// block-wide flags of one inner iteration (not converged / active set changed): one byte per warp
	// of the system, written with a plain store before the barrier and OR-ed by every thread after it
	// (an atomicOr here put its round trip in front of the barrier); two slots used alternately
	alignas(8) unsigned char wflags[2][PS_MAX_WARPS];
	double red[PS_MAX_WARPS];   // max-reduction of the variable stepper's relativeError
};

// `variable`: one more row array, V^n of the running step (ReverseVariableStepper)
inline size_t smem_bytes(int M, int P, bool variable = false) {
	return sizeof(Smem) + sizeof(double) * ((size_t)(N_ROW_ARRAYS + (variable ? 1 : 0)) * M * P + 2 * (P + 1));
}

// relativeError's per-element test  |a-b| / max(scale,|a|,|b|) > tol
// (Core/IterativeMethod.hpp:1125-1137), decided per thread from two integer
// maxima (the high words of max |a-b| and max |a| over the thread's rows):
//   every |a-b| <= tol scale (1 - eps)                        => converged
//   max |a-b| > tol (1 + eps) (max|a| + max|a-b| + scale)     => not converged
// and only a thread that falls between the two forms the reference's quotients
// row by row — so the decision is always the reference's.
struct ConvTest {
	__device__ __forceinline__ static void track(double xn, double xo, int &dmax_hi, int &xmax_hi) {
		dmax_hi = max(dmax_hi, __double2hiint(fabs(xn - xo)));
		xmax_hi = max(xmax_hi, __double2hiint(fabs(xn)));
	}
	// 0: converged, 1: not converged, 2: undecided (run exact())
	__device__ __forceinline__ static int screen(int dmax_hi, int xmax_hi, double tol, double scale) {
		const double D_up = __hiloint2double(dmax_hi + 1, 0);   // > every |a-b|
		if(D_up <= tol * scale * (1. - 8e-16)) return 0;
		const double D_dn = __hiloint2double(dmax_hi, 0);       // <= max |a-b|
		const double XN_up = __hiloint2double(xmax_hi + 1, 0);  // > every |a|
		// den <= max(scale, |a| + |a-b|)
		return D_dn > tol * (1. + 8e-16) * (XN_up + D_up + scale) ? 1 : 2;
	}
	__device__ __forceinline__ static bool exact(double xn, double xo, double tol, double scale) {
		const double fa = fabs(xn), fb = fabs(xo);
		double den = fa < fb ? fb : fa;
		den = scale < den ? den : scale;
		return fabs(xn - xo) / den > tol;
	}
};

} // namespace

enum { MODE_LINEAR = 0, MODE_PENALTY = 1, MODE_POLICY = 2 };

// MODE_LINEAR : root = the discretisation node (European / Bermudan)
// MODE_PENALTY: root = MinPenaltyMethodDifference1 under a ToleranceIteration
// MODE_POLICY : discretisation over Min/MaxPolicyIteration1_1 (two control
//               values of the interest rate) under a ToleranceIteration
// EXTRA: what the time loop carries besides the steps — user events, or the variable stepper
// (mutually exclusive at the boundary); compile-time so that the plain sweep pays for neither
enum { EXTRA_NONE = 0, EXTRA_EVENTS = 1, EXTRA_VARIABLE = 2 };

// CL: CTAs per contract.  CL = 2 is a thread-block cluster: the two CTAs hold the lower and the
// upper half of the rows (N up to 2 PT M + 1 = 4097), exchange the row at their boundary and the
// warps' separator data through distributed shared memory, and every barrier is the cluster's.
// GT ("general time"): forward-time sweeps, explicit event times, an operator source b(t), event programs.  The common
// sweep (GT = false, qpde_resident.cu) compiles them out of its time loop; qpde_resident_gt.cu holds the
// GT = true instantiations.
template <int M, int PT, int MODE, bool BDF2, int EXTRA, int CL, bool GT>
__global__ void __launch_bounds__(PT, 1)
k_resident_sweep(DeviceBatch b, DeviceResult out) {
	constexpr bool AMERICAN = MODE == MODE_PENALTY;
	constexpr bool POLICY = MODE == MODE_POLICY;
	constexpr bool ITERATE = MODE != MODE_LINEAR;
	constexpr bool EVENTS = EXTRA == EXTRA_EVENTS, VARIABLE = EXTRA == EXTRA_VARIABLE;
	constexpr int P = PT;                    // blockDim.x == PT
	extern __shared__ __align__(16) unsigned char smem_raw[];
	Smem &sm = *reinterpret_cast<Smem *>(smem_raw);

	const int cidx = b.c0 + (int)(blockIdx.x / CL);
	const int rank = CL == 1 ? 0 : (int)(blockIdx.x % CL);      // == %cluster_ctarank for a 1-D cluster
	const int tid = threadIdx.x;
	const int lane = tid & 31, warp = rank * (PT / 32) + (tid >> 5);   // warp: index in the whole system
	const int N = b.N;
	// P M rows are available.  A grid of exactly P M + 1 nodes (2049 = 256 * 8 + 1)
	// keeps its last row — diagonal-only in BlackScholes::A, BlackScholes.hpp:177-180 —
	// as a scalar "tail" equation folded into row N-2 (last row of the last thread);
	// smaller grids fit entirely and have no tail.
	const bool has_tail = N == CL * P * M + 1;
	const int n_main = has_tail ? N - 1 : N;
	const int i0 = (rank * P + tid) * M;       // first row of this thread
	const int row0 = rank * P * M;             // first row of this CTA
	constexpr int MP = M * P;

	double *const sm_S  = reinterpret_cast<double *>(smem_raw + sizeof(Smem));
	double *const sm_lo = sm_S + MP;   // operator rows of BlackScholes::A
	double *const sm_di = sm_lo + MP;
	double *const sm_up = sm_di + MP;
	double *const sm_a  = sm_up + MP;  // system rows for the current step: (a, bd = 1 + e, c)
	double *const sm_bd = sm_a + MP;
	double *const sm_c  = sm_bd + MP;
	double *const sm_na = sm_c + MP;   // right-hand-side rows of Crank-Nicolson: (0 - a, 1 - e, 0 - c)
	double *const sm_me = sm_na + MP;
	double *const sm_nc = sm_me + MP;
	double *const sm_lb = sm_nc + MP;  // right-hand side of the discretisation node (without penalty terms)
	double *const sm_xlast = sm_lb + MP;        // [P + 1]: x of each thread's last row, shifted by one
	double *const sm_xfirst = sm_xlast + P + 1; // [P + 1]: x of each thread's first row
	double *const sm_vn = sm_xfirst + P + 1;    // [MP], variable steps only: V^n of the running step

	const double K = b.strike[cidx];
	const double *const axis = b.axis + (size_t)cidx * b.axis_stride;

	// the other CTAs of the cluster: their Smem (solver hand-over, flags) and boundary slots
	PartitionTeam<CL> team; team.rank = rank; team.peer[0] = &sm.ps;
	Smem *peer_sm[CL]; peer_sm[0] = &sm;
	double *xlast_of_next = nullptr, *xfirst_of_prev = nullptr;
	if(CL > 1) {
		cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
		for(int q = 0; q < CL; ++q) { peer_sm[q] = cluster.map_shared_rank(&sm, q); team.peer[q] = &peer_sm[q]->ps; }
		if(rank + 1 < CL) xlast_of_next = cluster.map_shared_rank(sm_xlast, rank + 1);
		if(rank > 0) xfirst_of_prev = cluster.map_shared_rank(sm_xfirst, rank - 1);
	}
	// x of a thread's last / first row goes to the neighbour thread; across the CTA boundary
	// the slot lives in the neighbour CTA's shared memory
	auto publish_last = [&] (double v) {
		sm_xlast[tid + 1] = v;
		if(CL > 1 && tid == P - 1 && rank + 1 < CL) xlast_of_next[0] = v;
	};
	auto publish_first = [&] (double v) {
		sm_xfirst[tid] = v;
		if(CL > 1 && tid == 0 && rank > 0) xfirst_of_prev[P] = v;
	};

	// ------------------------------------------------------------------
	// Setup: refined grid -> shared memory, operator rows, payoff, obstacle
	// ------------------------------------------------------------------
	{
		for(int j = 0; j < M; ++j) {
			const int i = i0 + j;
			sm_S[j * P + tid] = i < n_main ? refined_node(axis, b.refine, i) : 0.;
		}
		if(tid == 0) sm.tail.S = has_tail ? refined_node(axis, b.refine, N - 1) : 0.;
	}
	team_sync<CL>();
	auto S_at = [&] (int i) -> double {       // node i of the refined grid, 0 <= i < N
		if(i >= n_main) return sm.tail.S;
		const int l = i - row0;               // rows of this CTA are in shared memory
		if(CL > 1 && (l < 0 || l >= MP)) return refined_node(axis, b.refine, i);
		return sm_S[(l % M) * P + (l / M)];
	};
	{
		const double r = b.r[cidx], q = b.q[cidx], sig = b.sigma[cidx];
		for(int j = 0; j <= M; ++j) {
			const bool tail = j == M;             // the tail row is handled by thread 0
			if(tail && (tid != 0 || !has_tail)) break;
			const int i = tail ? N - 1 : i0 + j;
			Row row; row.lo = 0.; row.di = 0.; row.up = 0.;
			if(tail || i < n_main) {
				const double Si = S_at(i);
				const double Sm = i > 0 ? S_at(i - 1) : 0.;
				const double Sp = i < N - 1 ? S_at(i + 1) : 0.;
				const double v_i = b.vol_model == QPDE_VOL_NODES
					? b.sigma_nodes[(size_t)cidx * b.sigma_nodes_stride + i]
					: vol_value(b.vol_model, sig, Si);
				if(POLICY) {
					// control value 0 -> (lo, di, up); control value 1 -> (na, me, nc)
					const double *ctl = b.controls + (size_t)cidx * b.controls_stride;
					row = bs_row(i, N, Sm, Si, Sp, ctl[0], v_i, q);
					const Row row1 = bs_row(i, N, Sm, Si, Sp, ctl[1], v_i, q);
					if(!tail) { sm_na[j * P + tid] = row1.lo; sm_me[j * P + tid] = row1.di; sm_nc[j * P + tid] = row1.up; }
				} else {
					row = bs_row(i, N, Sm, Si, Sp, r, v_i, q);
				}
			} else if(POLICY && !tail) {
				sm_na[j * P + tid] = 0.; sm_me[j * P + tid] = 0.; sm_nc[j * P + tid] = 0.;
			}
			if(tail) sm.tail.di = row.di;
			else { sm_lo[j * P + tid] = row.lo; sm_di[j * P + tid] = row.di; sm_up[j * P + tid] = row.up; }
		}
	}

	// ---- per-thread state (registers) ---------------------------------
	double x[M];        // current iterate / solution V^n
	double pz[M];       // obstacle (AMERICAN) or exercise value (EVENTS)
	double rr[M];       // right-hand side lb + penalty terms for the current active set
	double vprev[BDF2 ? M : 1]; // iterand(1) for BDF2
	double x_next;      // x of the next thread's first row
	PartitionSolver<M, PT, CL> ps;   // cached factorisation (registers)
	ps.init();
	unsigned curmask = 0, newmask = 0;  // bit j: row j penalised; bit M: tail row
	unsigned outmask = 0;               // penalty + events: the active set of the last penalty iterate
	// row N-2 (row M-1 of the last thread) couples to the tail row; that thread
	// owns the tail equation and folds the coupling into its rhs.
	const bool tail_owner = has_tail && rank == CL - 1 && tid == P - 1;

	auto initial_value = [&] (int i) -> double {
		return b.payoff_kind == QPDE_PAYOFF_ARRAY
			? b.payoff[(size_t)cidx * b.payoff_stride + i]
			: payoff_value(b.payoff_kind, S_at(i), K);
	};
	auto obstacle_value = [&] (int i) -> double {
		if(AMERICAN) {
			return b.obstacle_kind == QPDE_PAYOFF_ARRAY
				? b.obstacle[(size_t)cidx * b.obstacle_stride + i]
				: payoff_value(b.obstacle_kind, S_at(i), K);
		}
		return b.exercise_kind == QPDE_EXERCISE_NONE ? -CUDART_INF : payoff_value(b.exercise_kind, S_at(i), K);
	};
#pragma unroll
	for(int j = 0; j < M; ++j) {
		const int i = i0 + j;
		const bool real = i < n_main;
		x[j] = real ? initial_value(i) : 0.;
		pz[j] = real && (AMERICAN || EVENTS) ? obstacle_value(i) : -CUDART_INF;
		if(BDF2) vprev[j] = 0.;
		rr[j] = 0.;
	}
	if(tail_owner) {
		sm.tail.x = initial_value(N - 1);
		sm.tail.pz = AMERICAN || EVENTS ? obstacle_value(N - 1) : -CUDART_INF;
		sm.tail.vprev = 0.; sm.tail.c = 0.; sm.tail.bd = 1.; sm.tail.me = 1.; sm.tail.invb = 1.; sm.tail.pq = 0.;
	}
	x_next = i0 + M < n_main ? initial_value(i0 + M) : 0.;
	publish_last(x[M - 1]);
	if(tid == 0 && rank == 0) sm_xlast[0] = 0.;

	// PenaltyMethod.hpp:45,61-66: row penalised iff (scale * (x - payoff)) < 0,
	// i.e. iff x < payoff (scale > 0 and the difference of two distinct doubles
	// is never zero).
#define QPDE_MASK_OF(dst) \
	do { \
		unsigned mk_ = 0; \
		if(AMERICAN) { \
			_Pragma("unroll") \
			for(int j = 0; j < M; ++j) if(x[j] < pz[j]) mk_ |= 1u << j; \
			if(tail_owner && sm.tail.x < sm.tail.pz) mk_ |= 1u << M; \
		} \
		dst = mk_; \
	} while(0)
	// PolicyIteration::onIterationStart (Core/PolicyIteration.hpp:54-105): per row
	// candidate_k = (A(q_k) x - b(q_k))_i, SpMV from zero in column order, b = 0;
	// control 1 replaces control 0 only on a strict improvement.  Bit j = control
	// index of row j.  Reads sm_xlast: call after a barrier that follows its write.
#define QPDE_POLICY_OF(dst) \
	do { \
		unsigned mk_ = 0; \
		const double xm_ = sm_xlast[tid]; \
		_Pragma("unroll") \
		for(int j = 0; j < M; ++j) { \
			const int at_ = j * P + tid; \
			const double vm_ = j == 0 ? xm_ : x[j > 0 ? j - 1 : 0]; \
			const double vp_ = j == M - 1 ? (tail_owner ? sm.tail.x : x_next) : x[j + 1 < M ? j + 1 : j]; \
			double c0_ = 0., c1_ = 0.; \
			c0_ += sm_lo[at_] * vm_; c0_ += sm_di[at_] * x[j]; c0_ += sm_up[at_] * vp_; c0_ = c0_ - 0.; \
			c1_ += sm_na[at_] * vm_; c1_ += sm_me[at_] * x[j]; c1_ += sm_nc[at_] * vp_; c1_ = c1_ - 0.; \
			if(b.policy_max ? c1_ > c0_ : c1_ < c0_) mk_ |= 1u << j; \
		} \
		dst = mk_; \
	} while(0)
	team_sync<CL>();
	if(POLICY) QPDE_POLICY_OF(newmask); else QPDE_MASK_OF(newmask);
	team_sync<CL>();

	// ------------------------------------------------------------------
	// Time loop
	// ------------------------------------------------------------------
	ReplayT<GT> rp; rp.start(b.T[cidx], b.dt[cidx], b.n_exercise, b.forward, b.event_times);
	NodeState ns; ns.start(b.scheme);
	// ReverseVariableStepper (Core/Stepper.hpp:60-126): dt of the next step from the
	// relative change over the last one
	constexpr bool variable = VARIABLE;
	const double vtarget = variable ? b.variable_target[cidx] : 0.;
	double vdt = b.dt[cidx];
	double time1 = rp.initial(), time0 = rp.initial();
	int cur_kind = -1; double cur_h = 0.;
	int n_steps = 0, inner_total = 0, computed = 0, status = 0;
	bool more = true;
	bool refactor = true;        // block-uniform: the cached factorisation is stale
	unsigned flip = 0;           // block-uniform: which flag slot the next reduction uses

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
		Gamma g; g.kind = kind; g.h = h0;
		double c1 = 0., c0 = 0., den = 1.;
		if(BDF2 && kind == STEP_BDF2) {
			const double h1 = rp.lapse(time1, t), hh0 = rp.lapse(time0, t);      // BDF.hpp:68-73
			g.h = div_step(hh0 * h1, hh0 + h1);
			c1 = div_step(hh0, h1); c0 = div_step(h1, hh0); den = hh0 - h1;
		}
		const double rden = div_step(1., den);
		// (Re)build the system rows when the step kind or size changes.  h0 is a
		// difference of accumulated times and moves in its last bits from step
		// to step (by ulp(t) / dt, ~1e-13 relative); moves below 1e-11 relative
		// do not trigger a rebuild: the step is then taken with an h that is off
		// by that relative amount, which perturbs the solution at the 1e-13
		// level (the reference itself keeps a stale A for non-penalised roots,
		// IterativeMethod.hpp:906).
		if(kind != cur_kind || fabs(g.h - cur_h) > 1e-11 * cur_h) {
			cur_kind = kind; cur_h = g.h;
#pragma unroll
			for(int j = 0; j < M; ++j) {
				if(POLICY) break;     // rows are scaled at factorisation time, per selected control
				const int at = j * P + tid;
				const double aj = g.off(sm_lo[at]), ej = g.off(sm_di[at]);
				double cj = g.off(sm_up[at]);
				sm_na[at] = 0. - aj; sm_me[at] = 1. - ej; sm_nc[at] = 0. - cj;
				if(j == M - 1 && tail_owner) { sm.tail.c = cj; cj = 0.; }
				sm_a[at] = aj; sm_bd[at] = 1. + ej; sm_c[at] = cj;
			}
			if(tail_owner) {
				const double et = g.off(sm.tail.di);
				sm.tail.bd = 1. + et; sm.tail.me = 1. - et;
			}
			refactor = true;
			// rows are private to their thread: no barrier needed
		}

		// ---- lb = b(t) of the discretisation node (V^n = x), kept in shared
		// memory; rr = lb + penalty terms of the current active set ----------
		{
			double lb[M];
			// b(t) of the operator: 0, or source_table[floor(t)] S (GLWBOperator::b, glwb.cpp:40-44); enters the
			// BDF right-hand sides v + h b (BDF.hpp:41-46) and ((..) / (h0 - h1) + b) hh (BDF.hpp:96-108)
			double gsrc = 0.;
			if(GT && b.source_table) {
				int nidx = (int)floor(t);
				nidx = nidx < 0 ? 0 : (nidx > b.n_source - 1 ? b.n_source - 1 : nidx);
				gsrc = b.source_table[nidx];
			}
			if(kind == STEP_IMPLICIT) {
#pragma unroll
				for(int j = 0; j < M; ++j) lb[j] = GT && b.source_table ? x[j] + h0 * (gsrc * sm_S[j * P + tid]) : x[j] + 0. * h0;
				if(tail_owner) sm.tail.lb = GT && b.source_table ? sm.tail.x + h0 * (gsrc * sm.tail.S) : sm.tail.x + 0. * h0;
			} else if(BDF2 && kind == STEP_BDF2) {
#pragma unroll
				for(int j = 0; j < M; ++j) {
					lb[j] = (div_by(c1 * x[j] - c0 * vprev[j], den, rden) + (GT && b.source_table ? gsrc * sm_S[j * P + tid] : 0.)) * g.h;
				}
				if(tail_owner) sm.tail.lb = (div_by(c1 * sm.tail.x - c0 * sm.tail.vprev, den, rden) + (GT && b.source_table ? gsrc * sm.tail.S : 0.)) * g.h;
			} else {
				// (I - A h/2) V^n: row-major SpMV accumulated from zero in column order
				const double xm = sm_xlast[tid];       // x of row i0 - 1
#pragma unroll
				for(int j = 0; j < M; ++j) {
					const int at = j * P + tid;
					const double vm = j == 0 ? xm : x[j - 1];
					const double vp = j == M - 1 ? x_next : x[j + 1];
					double acc = 0.;
					acc += sm_na[at] * vm;
					acc += sm_me[at] * x[j];
					acc += sm_nc[at] * vp;                      // vp = 0 for the tail owner's last row
					if(j == M - 1 && tail_owner) acc += (0. - sm.tail.c) * sm.tail.x;
					lb[j] = acc + 0.;
				}
				if(tail_owner) sm.tail.lb = (0. + sm.tail.me * sm.tail.x) + 0.;
			}
#pragma unroll
			for(int j = 0; j < M; ++j) {
				sm_lb[j * P + tid] = lb[j];
				rr[j] = AMERICAN && (curmask >> j & 1u) ? lb[j] + b.pen * pz[j] : lb[j];
			}
		}
		if(BDF2) {
#pragma unroll
			for(int j = 0; j < M; ++j) vprev[j] = x[j];
			if(tail_owner) sm.tail.vprev = sm.tail.x;
		}
		if(variable) {
#pragma unroll
			for(int j = 0; j < M; ++j) sm_vn[j * P + tid] = x[j];
			if(tail_owner) sm.tail.vn = sm.tail.x;
		}

		int its = 0;
		bool again;
		do {
			// ==============================================================
			// Factorisation of (lA + P) for the current rows and active set
			// ==============================================================
			if(refactor) {
				refactor = false;
				curmask = newmask;
				{
					double aa[M], bb[M], cc[M];
#pragma unroll
					for(int j = 0; j < M; ++j) {
						const int at = j * P + tid;
						if(POLICY) {
							const bool sel = curmask >> j & 1u;
							Gamma gA; gA.kind = cur_kind; gA.h = cur_h;
							aa[j] = gA.off(sel ? sm_na[at] : sm_lo[at]);
							bb[j] = 1. + gA.off(sel ? sm_me[at] : sm_di[at]);
							cc[j] = gA.off(sel ? sm_nc[at] : sm_up[at]);
							if(j == M - 1 && tail_owner) { sm.tail.c = cc[j]; cc[j] = 0.; }
						} else {
							aa[j] = sm_a[at]; bb[j] = sm_bd[at]; cc[j] = sm_c[at];
						}
						const bool act = AMERICAN && (curmask >> j & 1u);
						if(act) bb[j] = bb[j] + b.pen * 1.;
						const double lbj = sm_lb[at];
						rr[j] = act ? lbj + b.pen * pz[j] : lbj;      // PenaltyMethod.hpp:96-119
					}
					if(tail_owner) {
						const bool act = AMERICAN && (curmask >> M & 1u);
						sm.tail.invb = rcp(act ? sm.tail.bd + b.pen * 1. : sm.tail.bd);
						sm.tail.pq = act ? b.pen * sm.tail.pz : 0.;
					}
					ps.factorize(aa, bb, cc, team, lane, warp);
				}
			}

			// ==============================================================
			// One solve with the cached factorisation
			// ==============================================================
			// right-hand side with penalty terms (PenaltyMethod.hpp:96-119); the
			// tail row is a scalar equation, folded into row N-2
			double xt_new = 0., r_last = rr[M - 1];
			if(tail_owner) {
				xt_new = (sm.tail.lb + sm.tail.pq) * sm.tail.invb;
				r_last = fma(-sm.tail.c, xt_new, r_last);
			}
			double G[M];
			{
				double r[M];
#pragma unroll
				for(int j = 0; j < M; ++j) r[j] = j == M - 1 ? r_last : rr[j];
				ps.solve(r, G, x_next, team, lane, warp);
			}
			bool nd = false;
			if(ITERATE) {
				int dmax_hi = 0, xmax_hi = 0;
#pragma unroll
				for(int j = 0; j < M; ++j) ConvTest::track(G[j], x[j], dmax_hi, xmax_hi);
				if(tail_owner) ConvTest::track(xt_new, sm.tail.x, dmax_hi, xmax_hi);
				const int verdict = ConvTest::screen(dmax_hi, xmax_hi, b.tolerance, b.scale);
				nd = verdict == 1;
				if(verdict == 2) {
#pragma unroll
					for(int j = 0; j < M; ++j) nd = nd | ConvTest::exact(G[j], x[j], b.tolerance, b.scale);
					if(tail_owner) nd = nd | ConvTest::exact(xt_new, sm.tail.x, b.tolerance, b.scale);
				}
			}
#pragma unroll
			for(int j = 0; j < M; ++j) x[j] = G[j];
			if(tail_owner) sm.tail.x = xt_new;
			publish_last(x[M - 1]);
			++its; ++computed;
			if(ITERATE) {
				if(POLICY) {
					team_sync<CL>();             // sm_xlast of the new iterate is visible
					QPDE_POLICY_OF(newmask);
				} else {
					QPDE_MASK_OF(newmask);
				}
				// one barrier carries both block-wide flags
				const unsigned bits = (nd ? 1u : 0u) | (newmask != curmask ? 2u : 0u);
				const unsigned wbits = __reduce_or_sync(FULL, bits);
				const unsigned slot = flip & 1u;   // two slots used alternately
				++flip;
				if(lane == 0) {
#pragma unroll
					for(int q = 0; q < CL; ++q) peer_sm[q]->wflags[slot][warp] = (unsigned char)wbits;   // every CTA sees every flag
				}
				team_sync<CL>();
				unsigned all = 0u;
				{
					constexpr int W = CL * (PT / 32);          // warps of the whole system
					if(W % 8 == 0) {
						const unsigned long long *words = reinterpret_cast<const unsigned long long *>(sm.wflags[slot]);
						unsigned long long acc = 0ull;
#pragma unroll
						for(int k = 0; k < W / 8; ++k) acc |= words[k];
						acc |= acc >> 32; acc |= acc >> 16; acc |= acc >> 8;
						all = (unsigned)acc & 3u;
					} else {
#pragma unroll
						for(int k = 0; k < W; ++k) all |= sm.wflags[slot][k];
					}
				}
				const bool notdone = (all & 1u) != 0, changed = (all & 2u) != 0;
				refactor = changed;
				again = notdone;
				if(notdone && !changed) {
					// Same active set => the next inner iteration would assemble the
					// same matrix and right-hand side and return this iterate again
					// (relativeError == 0): count it, do not redo it.
					if(its >= b.max_inner) status = 1; else ++its;
					again = false;
				} else if(again && its >= b.max_inner) {
					status = 1; again = false;
				}
			} else {
				team_sync<CL>();
				again = false;
			}
		} while(again);

		if(rank == 0 && tid == 0 && out.inner && n_steps < b.max_steps) {
			out.inner[(size_t)cidx * out.inner_stride + n_steps] = (unsigned char)(its > 255 ? 255 : its);
		}
		inner_total += its;
		++n_steps;
		time0 = time1; time1 = t;
		ns.end_step();

		if(variable && more) {
			// dt *= target / (relativeError(iterand(0), iterand(1), scale) + epsilon), Stepper.hpp:68-82;
			// relativeError = max_i |a - b| / max(scale, |a|, |b|), IterativeMethod.hpp:1125-1137
			double err = 0.;
#pragma unroll
			for(int j = 0; j < M; ++j) {
				const double vn = sm_vn[j * P + tid];
				const double fa = fabs(x[j]), fb = fabs(vn);
				double dn = fa < fb ? fb : fa;
				dn = b.variable_scale < dn ? dn : b.variable_scale;
				const double e = i0 + j < n_main ? fabs(x[j] - vn) / dn : 0.;
				err = e > err ? e : err;
			}
			{   // the tail row: formed by every thread (no branch in front of the shuffles), used by its owner
				const double fa = fabs(sm.tail.x), fb = fabs(sm.tail.vn);
				double dn = fa < fb ? fb : fa;
				dn = b.variable_scale < dn ? dn : b.variable_scale;
				const double e = fabs(sm.tail.x - sm.tail.vn) / dn;
				err = tail_owner && e > err ? e : err;
			}
#pragma unroll
			for(int dl = 16; dl >= 1; dl >>= 1) { const double o = shfl_xor(err, dl); err = o > err ? o : err; }
			if(lane == 0) sm.red[warp] = err;
			__syncthreads();
			err = 0.;
			for(int wv = 0; wv < P / 32; ++wv) err = sm.red[wv] > err ? sm.red[wv] : err;
			vdt *= vtarget / (err + DBL_EPSILON);
		}

		if(st.event_after) {
			// PenaltyMethod::constraintMask (PenaltyMethod.hpp:127-139) reads the tolerance iteration's last
			// iterand, which events do not touch (they transform the stepper's): the mask that goes out is
			// the one of the last penalty iterate (the last step of a sweep always ends in an event)
			if(AMERICAN && EVENTS) outmask = newmask;
			if(EVENTS && st.user_events > 0) {
				// Events at one time go in descending add() order (IterativeMethod.hpp:1340-1352)
				for(int ue = 0; ue < st.user_events; ++ue) {
					for(int op = 0; op < 2; ++op) {
						const bool is_dividend = (op == 0) == (b.dividend_first != 0);
						const bool program = GT && op == 0 && b.event_program != nullptr;            // a general Event<1> transform
						if(CL == 1 && ((is_dividend && b.dividend_kind != QPDE_DIVIDEND_NONE) || program)) {   // reads any node: one CTA only
							// Event<1>::_doEvent with V(max(S - D, 0)) or V((1 - D) S): the solution goes to
							// shared memory (sm_lb is free between steps), every node reads its two
							// neighbours of the shifted price (Core/Interpolant.hpp:153-181,331-374).
							// Fixed-trip bisection: no divergence in front of the solver's shuffles.
							const double D = program ? 0. : b.dividend[cidx];
							// the parameters of this event: user events are popped in descending (forward: ascending) index
							const int ev_index = GT && b.forward ? rp.e - st.user_events + ue : rp.e + st.user_events - ue;
							const double *const ev_params = program
								? b.event_params + (size_t)cidx * b.event_params_stride + (size_t)ev_index * b.n_event_params : nullptr;
#pragma unroll
							for(int j = 0; j < M; ++j) sm_lb[j * P + tid] = x[j];
							__syncthreads();
							auto V_at = [&] (int i) -> double { return i >= n_main ? sm.tail.x : sm_lb[(i % M) * P + (i / M)]; };
							const double Sfirst = S_at(0), Slast = S_at(N - 1);
							// PiecewiseLinear::interpolate of the pre-event solution at any price
							auto V_interp = [&] (double at) -> double {
								int lo = 0, hi = N - 2;
								for(int it = 0; it < 13; ++it) {          // 2^13 > any supported N
									const int mid = (lo + hi + 1) >> 1;
									const bool up = S_at(mid) <= at;
									lo = up ? mid : lo; hi = up ? hi : mid - 1;
								}
								double wgt = (S_at(lo + 1) - at) / (S_at(lo + 1) - S_at(lo));
								int idx = lo;
								if(at <= Sfirst) { idx = 0; wgt = 1.; }
								if(at >= Slast) { idx = N - 2; wgt = 0.; }
								double sum = 0.;
								if(wgt > DBL_EPSILON) sum += wgt * V_at(idx);
								if(1. - wgt > DBL_EPSILON) sum += (1. - wgt) * V_at(idx + 1);
								return sum;
							};
							auto shifted = [&] (double Si) -> double {
								if(program) return event_eval(b.event_program, b.event_consts, ev_params, Si, V_interp);
								return V_interp(b.dividend_kind == QPDE_DIVIDEND_CASH ? (Si - D > 0. ? Si - D : 0.) : (1. - D) * Si);
							};
							double xn[M];
#pragma unroll
							for(int j = 0; j < M; ++j) xn[j] = i0 + j < n_main ? shifted(sm_S[j * P + tid]) : 0.;
							const double xt = has_tail ? shifted(Slast) : 0.;   // uniform: every thread
							__syncthreads();
#pragma unroll
							for(int j = 0; j < M; ++j) x[j] = xn[j];
							if(tail_owner) sm.tail.x = xt;
						}
						if(!is_dividend && b.exercise_kind != QPDE_EXERCISE_NONE) {
							// Event<1>::_doEvent with max(V(S), K - S): Core/Event.hpp:100-112.  Under the penalty
							// iteration pz holds the obstacle, so the exercise value is formed from the grid.
#pragma unroll
							for(int j = 0; j < M; ++j) {
								const double ev = AMERICAN ? (i0 + j < n_main ? payoff_value(b.exercise_kind, sm_S[j * P + tid], K) : -CUDART_INF) : pz[j];
								x[j] = x[j] > ev ? x[j] : ev;
							}
							if(tail_owner) {
								const double ev = AMERICAN ? payoff_value(b.exercise_kind, sm.tail.S, K) : sm.tail.pz;
								sm.tail.x = sm.tail.x > ev ? sm.tail.x : ev;
							}
						}
					}
				}
				if(AMERICAN) {
					// the next step's first penalty iteration assembles P from the transformed solution
					// (PenaltyMethod.hpp:37-69 reads iterand(0)): new active set, stale factorisation
					QPDE_MASK_OF(newmask);
					refactor = true;
				}
				team_sync<CL>();
				publish_first(x[0]);
				publish_last(x[M - 1]);
				if(tid == 0 && rank == CL - 1) sm_xfirst[P] = 0.;
				team_sync<CL>();
				x_next = sm_xfirst[tid + 1];
			}
			ns.after_event();
		}
	}
#undef QPDE_MASK_OF
#undef QPDE_POLICY_OF

	// ------------------------------------------------------------------
	// Outputs
	// ------------------------------------------------------------------
#pragma unroll
	for(int j = 0; j < M; ++j) {
		const int i = i0 + j;
		if(i < n_main) {
			if(out.V) out.V[(size_t)cidx * out.V_stride + i] = x[j];
			if(out.S) out.S[(size_t)cidx * out.S_stride + i] = sm_S[j * P + tid];
			if(out.active) {
				const double pred = AMERICAN ? (0. + 1. * x[j]) - pz[j] : 0.;
				out.active[(size_t)cidx * out.active_stride + i] = (AMERICAN && EVENTS ? (outmask >> j & 1u) != 0u : pred < 0.) ? 1 : 0;
			}
		}
	}
	team_sync<CL>();   // the tail owner's last writes to sm.tail become visible; no CTA leaves while a peer may still write into it
	if(tail_owner) {
		if(out.V) out.V[(size_t)cidx * out.V_stride + N - 1] = sm.tail.x;
		if(out.S) out.S[(size_t)cidx * out.S_stride + N - 1] = sm.tail.S;
		if(out.active) {
			const double pred = AMERICAN ? (0. + 1. * sm.tail.x) - sm.tail.pz : 0.;
			out.active[(size_t)cidx * out.active_stride + N - 1] = (AMERICAN && EVENTS ? (outmask >> M & 1u) != 0u : pred < 0.) ? 1 : 0;
		}
	}
	if(rank == 0 && tid == 0) {
		if(out.n_steps) out.n_steps[cidx] = n_steps;
		if(out.inner_total) out.inner_total[cidx] = inner_total;
		if(out.computed) out.computed[cidx] = computed;
		if(out.status) out.status[cidx] = status;
	}
	if(out.value) {
		// PiecewiseLinear::interpolate (Core/Interpolant.hpp:153-181,331-374):
		// the thread owning the bracketing interval writes the value
		const double s0 = b.spot ? b.spot[cidx] : K;
		const double Sfirst = S_at(0), Slast = S_at(N - 1);
#pragma unroll
		for(int j = 0; j < M; ++j) {
			const int i = i0 + j;
			if(i >= N - 1) continue;              // intervals [S_i, S_{i+1}), i = 0 .. N-2
			const double Si = sm_S[j * P + tid];
			const double Sp = S_at(i + 1);
			const double vi = x[j];
			const double vp = i + 1 >= n_main ? sm.tail.x : (j == M - 1 ? x_next : x[j + 1 < M ? j + 1 : j]);
			bool mine = false; double wgt = 0.;
			if(s0 <= Sfirst) { mine = i == 0; wgt = 1.; }
			else if(s0 >= Slast) { mine = i == N - 2; wgt = 0.; }
			else if(!(s0 < Si) && !(s0 >= Sp)) { mine = true; wgt = (Sp - s0) / (Sp - Si); }
			if(mine) {
				double sum = 0.;
				if(wgt > DBL_EPSILON) sum += wgt * vi;
				if(1. - wgt > DBL_EPSILON) sum += (1. - wgt) * vp;
				out.value[cidx] = sum;
			}
		}
	}
}

// Geometry (rows per thread M, threads PT) for a grid of N nodes.  P M rows are
// available; a grid of P M + 1 nodes uses the tail equation.
//   N <=  257: M = 8, PT =  32 (one warp per contract: no idle second warp, eight contracts per SM;
//              2.1x over 8 x 64 at 33 .. 257 nodes.  A 4 x 32 variant for N <= 129 measured +11 % at
//              33 / 65 nodes and -1 % at 129: not kept)
//   N <=  513: M = 8, PT =  64      N <= 2049: M = 8, PT = 256 (or M = 4, PT = 512)
//   N <= 2305: M = 9, PT = 256
//   N <= 4097: M = 8, PT = 256, CL = 2 (a cluster of two CTAs per contract); N <= 8193: CL = 4
static int pick_geometry(int N, int *M, int *PT, int *CL = nullptr) {
	static const char *env = getenv("QPDE_RESIDENT_GEOMETRY");   // "4x512": A/B the wide variant
	if(CL) *CL = 1;
	if(N <= 8 * 32 + 1 && !(env && !strcmp(env, "no32"))) { *M = 8; *PT = 32; return 1; }
	if(N <= 8 * 64 + 1)       { *M = 8; *PT = 64;  return 1; }
	if(N <= 8 * 256 + 1) {
		if(env && !strcmp(env, "4x512")) { *M = 4; *PT = 512; return 1; }
		*M = 8; *PT = 256; return 1;
	}
	if(N <= 9 * 256 + 1)      { *M = 9; *PT = 256; return 1; }
	if(N <= 2 * 8 * 256 + 1)  { *M = 8; *PT = 256; if(CL) *CL = 2; return 1; }
	if(N <= 4 * 8 * 256 + 1)  { *M = 8; *PT = 256; if(CL) *CL = 4; return 1; }
	return 0;
}

#if !QPDE_RESIDENT_GT
bool resident_supported(const DeviceBatch &b) {
	int M, PT, CL;
	if(b.N < 10) return false;
	if(b.scheme > QPDE_SCHEME_BDF2) return false;   // BDF3 .. BDF6: INTERLEAVED family
	if(!pick_geometry(b.N, &M, &PT, &CL)) return false;
	if(CL > 1 && (b.variable_target || b.dividend_kind != QPDE_DIVIDEND_NONE || b.event_program)) return false; // one-CTA features
	if(b.american && b.n_exercise > 0 && (CL > 1 || (M == 4 && PT == 512))) return false; // penalty + events: one CTA per contract, the default geometries
	if(b.n_controls != 0 && b.n_controls != 2) return false; // more control values: INTERLEAVED family
	if(b.n_controls != 0 && bdf_order(b.scheme) == 0) return false; // policy iteration under a theta scheme (the right-hand side follows the controls): INTERLEAVED family
	if(b.variable_target && !((M == 8 && PT == 32) || (M == 8 && PT == 64) || (M == 8 && PT == 256))) return false; // no room for V^n: INTERLEAVED family
	return true;
}

#endif

template <int M, int PT, int MODE, bool BDF2, int EXTRA, int CL>
static int launch_variant(const DeviceBatch &b, const DeviceResult &out, cudaStream_t stream) {
	auto kern = k_resident_sweep<M, PT, MODE, BDF2, EXTRA, CL, QPDE_RESIDENT_GT != 0>;
	constexpr bool VARIABLE = EXTRA == EXTRA_VARIABLE;        // one more row array: V^n
	const size_t smem = smem_bytes(M, PT, VARIABLE);
	// per device / context, and a handle may live on any device: set on every launch (as the other families do)
	if(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
		return QPDE_ERR_CUDA;
	}
	if(CL == 1) {
		kern<<<b.C, PT, smem, stream>>>(b, out);
		return QPDE_OK;
	}
	// CL CTAs per contract as one thread-block cluster
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)b.C * CL); cfg.blockDim = dim3(PT); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kern, b, out) == cudaSuccess ? QPDE_OK : QPDE_ERR_CUDA;
}

template <int M, int PT, int EXTRA, int CL>
static int launch_extra(const DeviceBatch &b, const DeviceResult &out, cudaStream_t stream) {
	const bool bdf2 = b.scheme == QPDE_SCHEME_BDF2;
	if(b.n_controls > 0) return bdf2 ? launch_variant<M, PT, MODE_POLICY, true, EXTRA, CL>(b, out, stream)
	                                 : launch_variant<M, PT, MODE_POLICY, false, EXTRA, CL>(b, out, stream);
	if(b.american) return bdf2 ? launch_variant<M, PT, MODE_PENALTY, true, EXTRA, CL>(b, out, stream)
	                           : launch_variant<M, PT, MODE_PENALTY, false, EXTRA, CL>(b, out, stream);
	return bdf2 ? launch_variant<M, PT, MODE_LINEAR, true, EXTRA, CL>(b, out, stream)
	            : launch_variant<M, PT, MODE_LINEAR, false, EXTRA, CL>(b, out, stream);
}

// VARIABLE_OK: the geometry has room for the variable stepper's V^n array (and is worth the
// extra instantiations)
template <int M, int PT, bool VARIABLE_OK, int CL>
static int launch_rows(const DeviceBatch &b, const DeviceResult &out, cudaStream_t stream) {
	const bool bdf2 = b.scheme == QPDE_SCHEME_BDF2;
	if(b.variable_target) {
		if(VARIABLE_OK) return launch_extra<M, PT, VARIABLE_OK ? EXTRA_VARIABLE : EXTRA_NONE, CL>(b, out, stream);
		return QPDE_ERR_UNSUPPORTED;
	}
	if(b.n_exercise > 0) {      // events: linear sweeps, and penalised ones on one CTA (American with discrete dividends)
		constexpr bool PENALTY_OK = CL == 1 && !(M == 4 && PT == 512);
		if(b.american) {
			if(!PENALTY_OK) return QPDE_ERR_UNSUPPORTED;
			return bdf2 ? launch_variant<M, PT, PENALTY_OK ? MODE_PENALTY : MODE_LINEAR, true, EXTRA_EVENTS, CL>(b, out, stream)
			            : launch_variant<M, PT, PENALTY_OK ? MODE_PENALTY : MODE_LINEAR, false, EXTRA_EVENTS, CL>(b, out, stream);
		}
		return bdf2 ? launch_variant<M, PT, MODE_LINEAR, true, EXTRA_EVENTS, CL>(b, out, stream)
		            : launch_variant<M, PT, MODE_LINEAR, false, EXTRA_EVENTS, CL>(b, out, stream);
	}
	return launch_extra<M, PT, EXTRA_NONE, CL>(b, out, stream);
}

// The geometry has been picked (pick_geometry) and the batch accepted (resident_supported) by launch_resident.
int QPDE_RESIDENT_LAUNCH(const DeviceBatch &b, const DeviceResult &out, cudaStream_t stream, int M, int PT, int CL) {
	if(CL == 4)                  return launch_rows<8, 256, false, 4>(b, out, stream);
	else if(CL == 2)             return launch_rows<8, 256, false, 2>(b, out, stream);
	else if(M == 8 && PT == 32)  return launch_rows<8, 32, true, 1>(b, out, stream);
	else if(M == 8 && PT == 64)  return launch_rows<8, 64, true, 1>(b, out, stream);
	else if(M == 8 && PT == 256) return launch_rows<8, 256, true, 1>(b, out, stream);
	else if(M == 4 && PT == 512) return launch_rows<4, 512, false, 1>(b, out, stream);
	return launch_rows<9, 256, false, 1>(b, out, stream);
}

#if !QPDE_RESIDENT_GT
int launch_resident_general_time(const DeviceBatch &b, const DeviceResult &out, cudaStream_t stream, int M, int PT, int CL);   // qpde_resident_gt.cu

int launch_resident(const DeviceBatch &b, const DeviceResult &out, cudaStream_t stream,
		int sm_count, long long *launches) {
	(void)sm_count;
	int M = 0, PT = 0, CL = 1;
	if(!resident_supported(b) || !pick_geometry(b.N, &M, &PT, &CL)) return QPDE_ERR_UNSUPPORTED;
	if(lean_supported(b)) return launch_lean(b, out, stream, launches);      // two contracts per SM (qpde_lean.cu)
	const bool general_time = b.forward || b.event_times || b.source_table || b.event_program;   // (event programs ride along: their interpreter stays out of the common kernel's code)
	const int rc = general_time ? launch_resident_general_time(b, out, stream, M, PT, CL)
	                            : launch_resident_common_time(b, out, stream, M, PT, CL);
	if(rc == QPDE_OK) *launches += 1;
	return rc;
}
#endif

} // namespace qpde
