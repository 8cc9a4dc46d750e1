// qpde_resident.cu — kernel family "RESIDENT": one CTA per contract, the whole
// backward sweep (every time step, every penalty iteration) inside one launch
// with the contract's state held in registers (8 grid rows per thread) and
// shared memory.  HBM is touched only to read the contract's parameters
// (~0.3 KB) and to write its results.
//
// The tridiagonal system of each (inner) iteration is solved by a three-level
// partition that follows the hardware hierarchy:
//   level 1  thread : 7 interior rows eliminated serially (Thomas), leaving one
//                     separator unknown per thread (its first row);
//   level 2  warp   : the 31 separator unknowns of lanes 1..31 by parallel
//                     cyclic reduction, 5 stages of warp shuffles;
//   level 3  CTA    : one separator per warp (lane 0), a W x W tridiagonal
//                     system (W <= 9) solved by Thomas from shared memory.
// Everything that depends only on the matrix (multipliers, reciprocals, spike
// vectors, PCR coefficients) is cached in registers and recomputed only when
// the penalty active set or the step size changes; an ordinary inner iteration
// is right-hand-side work only: ~20 FMAs per row, two block barriers.
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

#include "qpde_kernels.cuh"

namespace qpde {

namespace {

constexpr int MAX_THREADS = 256; // 8 warps; M rows per thread (template) => N - 1 <= 256 M
constexpr int MAX_WARPS = MAX_THREADS / 32;
constexpr unsigned FULL = 0xffffffffu;

// Fixed part of the shared memory of one CTA.  The per-row arrays (S, lo, di,
// up: [M][P] each, laid out [j][thread] so that a warp reading "row j of every
// thread" touches consecutive doubles) and xlast / xfirst ([P + 1] each) follow
// it in dynamic shared memory, sized by the actual thread count P.
struct Smem {
	// level 1/2 -> level 3 hand-over, one slot per warp
	double f_Vs7[MAX_WARPS], f_Ws7[MAX_WARPS], f_V2l[MAX_WARPS], f_W2l[MAX_WARPS];
	double f_V2f[MAX_WARPS], f_W2f[MAX_WARPS], f_a0[MAX_WARPS], f_Bp[MAX_WARPS], f_Ct[MAX_WARPS];
	// level 3 factorisation
	double m3[MAX_WARPS], invd3[MAX_WARPS], C3[MAX_WARPS];
	double q1[MAX_WARPS], q2[MAX_WARPS], q3[MAX_WARPS];
	// right-hand-side hand-over
	double s_G7[MAX_WARPS], s_G2l[MAX_WARPS], s_P0[MAX_WARPS], s_G2f[MAX_WARPS];
	double S_tail, lo_tail, di_tail, up_tail;
	unsigned flags[2];
};

inline size_t smem_bytes(int M, int P) {
	return sizeof(Smem) + sizeof(double) * ((size_t)4 * M * P + 2 * (P + 1));
}

__device__ __forceinline__ double shfl_up(double v, int d) { return __shfl_up_sync(FULL, v, d); }
__device__ __forceinline__ double shfl_down(double v, int d) { return __shfl_down_sync(FULL, v, d); }

// 1 / d for the factorisation: hardware seed (MUFU.RCP64H, ~2^-22) plus two
// Newton steps => error below one ulp.  The factors need no more: any
// elimination order already moves the solution by a few ulp.
__device__ __forceinline__ double rcp(double d) {
	double y;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
	double e = fma(-d, y, 1.);
	y = fma(y, e, y);
	e = fma(-d, y, 1.);
	y = fma(y, e, y);
	return y;
}

// relativeError's per-element test |a-b| / max(scale,|a|,|b|) > tol
// (Core/IterativeMethod.hpp:1125-1137) without a division in the common case:
// the quotient is formed only when the product form is within rounding of the
// threshold, so the decision is the reference's.
__device__ __forceinline__ bool not_converged(double xn, double xo, double scale, double tol) {
	const double fa = fabs(xn), fb = fabs(xo);
	double den = fa < fb ? fb : fa;
	den = scale < den ? den : scale;
	const double d = fabs(xn - xo);
	const double lim = tol * den;
	if(d > lim * (1. + 1e-15)) return true;
	if(d < lim * (1. - 1e-15)) return false;
	return d / den > tol;
}

} // namespace

template <int M, bool AMERICAN, bool BDF2, bool EVENTS>
__global__ void __launch_bounds__(MAX_THREADS, 1)
k_resident_sweep(DeviceBatch b, DeviceResult out) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	Smem &sm = *reinterpret_cast<Smem *>(smem_raw);

	const int cidx = blockIdx.x;
	const int tid = threadIdx.x;
	const int P = blockDim.x;
	double *const sm_S = reinterpret_cast<double *>(smem_raw + sizeof(Smem));
	double *const sm_lo = sm_S + M * P;
	double *const sm_di = sm_lo + M * P;
	double *const sm_up = sm_di + M * P;
	double *const sm_xlast = sm_up + M * P;     // [P + 1]: x of each thread's last row, shifted by one
	double *const sm_xfirst = sm_xlast + P + 1; // [P + 1]: x of each thread's first row
	const int lane = tid & 31, warp = tid >> 5, W = P >> 5;
	const int N = b.N;
	const int n_main = N - 1;        // rows 0 .. N-2; row N-1 (decoupled, lo = 0) is the "tail"
	const int i0 = tid * M;

	const double tol = b.tolerance, scale = b.scale;
	const double pen = AMERICAN ? 1. / b.penalty_tolerance : 0.;   // PenaltyMethod.hpp:85
	const double K = b.strike[cidx];

	// ------------------------------------------------------------------
	// Setup: refined grid -> shared memory, operator rows, payoff, obstacle
	// ------------------------------------------------------------------
	{
		const double *axis = b.axis + (size_t)cidx * b.axis_stride;
		for(int j = 0; j < M; ++j) {
			const int i = i0 + j;
			sm_S[j * P + tid] = i < N ? refined_node(axis, b.refine, i) : 0.;
		}
		if(tid == 0) sm.S_tail = refined_node(axis, b.refine, N - 1);
	}
	__syncthreads();
	auto S_at = [&] (int i) -> double {       // node i of the refined grid, 0 <= i < N
		return i == N - 1 ? sm.S_tail : sm_S[(i % M) * P + (i / M)];
	};
	{
		const double r = b.r[cidx], q = b.q[cidx], sig = b.sigma[cidx];
		for(int j = 0; j <= M; ++j) {
			// j == M: the tail row, handled by thread 0
			const bool tail = j == M;
			if(tail && tid != 0) break;
			const int i = tail ? N - 1 : i0 + j;
			Row row; row.lo = 0.; row.di = 0.; row.up = 0.;
			if(i < N && (tail || i < n_main)) {
				const double Si = S_at(i);
				const double Sm = i > 0 ? S_at(i - 1) : 0.;
				const double Sp = i < N - 1 ? S_at(i + 1) : 0.;
				const double v_i = b.vol_model == QPDE_VOL_NODES
					? b.sigma_nodes[(size_t)cidx * b.sigma_nodes_stride + i]
					: vol_value(b.vol_model, sig, Si);
				row = bs_row(i, N, Sm, Si, Sp, r, v_i, q);
			}
			if(tail) { sm.lo_tail = row.lo; sm.di_tail = row.di; sm.up_tail = row.up; }
			else { sm_lo[j * P + tid] = row.lo; sm_di[j * P + tid] = row.di; sm_up[j * P + tid] = row.up; }
		}
	}
	__syncthreads();

	// ---- per-thread state (registers) ---------------------------------
	double x[M];        // current iterate / solution V^n
	double pz[M];       // obstacle (AMERICAN) or exercise value (EVENTS)
	double lb[M];       // right-hand side of the discretisation node
	double vprev[BDF2 ? M : 1]; // iterand(1) for BDF2
	double a[M], e[M], c[M];    // system rows: (a, 1 + e, c) for the current step kind
	double x_tail, pz_tail = 0., e_tail = 0., lb_tail = 0., vprev_tail = 0.;
	double ctail = 0.;  // coupling of row N-2 to the tail row (kept out of c[])
	double x_next;      // x of the next thread's first row
	// cached factorisation
	double f_m[M], f_invd[M], f_Vs[M], f_Ws[M];
	double f_al[5], f_ga[5], f_invB2 = 1., f_V2 = 0., f_W2 = 0.;
	unsigned curmask = 0, newmask = 0;  // bit j: row j penalised; bit M: tail row

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
		return payoff_value(b.exercise_kind, S_at(i), K);
	};
#pragma unroll
	for(int j = 0; j < M; ++j) {
		const int i = i0 + j;
		const bool real = i < n_main;
		x[j] = real ? initial_value(i) : 0.;
		pz[j] = real && (AMERICAN || EVENTS) ? obstacle_value(i) : -CUDART_INF;
		if(BDF2) vprev[j] = 0.;
		f_m[j] = 0.; f_invd[j] = 1.; f_Vs[j] = 0.; f_Ws[j] = 0.;
		lb[j] = 0.; a[j] = 0.; e[j] = 0.; c[j] = 0.;
	}
	x_tail = initial_value(N - 1);
	if(AMERICAN || EVENTS) pz_tail = obstacle_value(N - 1);
	x_next = i0 + M < n_main ? initial_value(i0 + M) : 0.;
	sm_xlast[tid + 1] = x[M - 1];
	if(tid == 0) { sm_xlast[0] = 0.; sm.flags[0] = 0u; sm.flags[1] = 0u; }
#pragma unroll
	for(int s = 0; s < 5; ++s) { f_al[s] = 0.; f_ga[s] = 0.; }

	auto mask_of = [&] () -> unsigned {
		// PenaltyMethod.hpp:45,61-66: (scale * (x - payoff)) < 0
		unsigned mk = 0;
		if(AMERICAN) {
#pragma unroll
			for(int j = 0; j < M; ++j) {
				const double pred = (0. + 1. * x[j]) - pz[j];
				if((pen * pred) < 0.) mk |= 1u << j;
			}
			const double pt = (0. + 1. * x_tail) - pz_tail;
			if((pen * pt) < 0.) mk |= 1u << M;
		}
		return mk;
	};
	newmask = mask_of();
	__syncthreads();

	// ------------------------------------------------------------------
	// Factorisation of (lA + P) for the current rows and active set
	// ------------------------------------------------------------------
	auto factorize = [&] () {
		double bb[M];
#pragma unroll
		for(int j = 0; j < M; ++j) {
			bb[j] = 1. + e[j];
			if(AMERICAN && (curmask >> j & 1u)) bb[j] = bb[j] + pen * 1.;
		}
		// level 1: interior rows 1 .. M-1
		double d = bb[1];
		f_invd[1] = rcp(d);
#pragma unroll
		for(int j = 2; j < M; ++j) {
			f_m[j] = a[j] * f_invd[j - 1];
			d = fma(-f_m[j], c[j - 1], bb[j]);
			f_invd[j] = rcp(d);
		}
		double y[M];
		y[1] = a[1];
#pragma unroll
		for(int j = 2; j < M; ++j) y[j] = -f_m[j] * y[j - 1];
		f_Vs[M - 1] = y[M - 1] * f_invd[M - 1];
		f_Ws[M - 1] = c[M - 1] * f_invd[M - 1];
#pragma unroll
		for(int j = M - 2; j >= 1; --j) {
			f_Vs[j] = (y[j] - c[j] * f_Vs[j + 1]) * f_invd[j];
			f_Ws[j] = (0. - c[j] * f_Ws[j + 1]) * f_invd[j];
		}
		// reduced row of this thread's separator (its row 0)
		const double Vs7p = shfl_up(f_Vs[M - 1], 1), Ws7p = shfl_up(f_Ws[M - 1], 1);
		const bool interior = lane >= 1;
		const double Bpart = bb[0] - c[0] * f_Vs[1];
		const double Ct = -c[0] * f_Ws[1];
		double A2 = interior ? -a[0] * Vs7p : 0.;
		double B2 = interior ? Bpart - a[0] * Ws7p : 1.;
		double C2 = interior ? Ct : 0.;
		// level 2: PCR over lanes 1 .. 31; the couplings of lane 1 to the warp
		// separator and of lane 31 to the next warp's separator become two
		// extra right-hand sides (spikes V2, W2)
		double V2 = lane == 1 ? A2 : 0.;
		double W2 = lane == 31 ? C2 : 0.;
		if(lane == 1) A2 = 0.;
		if(lane == 31) C2 = 0.;
#pragma unroll
		for(int s = 0; s < 5; ++s) {
			const int dl = 1 << s;
			const double rB = rcp(B2);
			const double rBm = shfl_up(rB, dl), rBq = shfl_down(rB, dl);
			const double Am = shfl_up(A2, dl), Aq = shfl_down(A2, dl);
			const double Cm = shfl_up(C2, dl), Cq = shfl_down(C2, dl);
			const double Vm = shfl_up(V2, dl), Vq = shfl_down(V2, dl);
			const double Wm = shfl_up(W2, dl), Wq = shfl_down(W2, dl);
			const bool okm = interior && lane - dl >= 1;
			const bool okq = interior && lane + dl <= 31;
			const double al = okm ? -A2 * rBm : 0.;
			const double ga = okq ? -C2 * rBq : 0.;
			f_al[s] = al; f_ga[s] = ga;
			B2 = B2 + (okm ? al * Cm : 0.) + (okq ? ga * Aq : 0.);
			V2 = V2 + (okm ? al * Vm : 0.) + (okq ? ga * Vq : 0.);
			W2 = W2 + (okm ? al * Wm : 0.) + (okq ? ga * Wq : 0.);
			A2 = okm ? al * Am : 0.;
			C2 = okq ? ga * Cq : 0.;
		}
		f_invB2 = rcp(B2);
		f_V2 = V2 * f_invB2;
		f_W2 = W2 * f_invB2;
		// hand-over to level 3
		if(lane == 31) { sm.f_Vs7[warp] = f_Vs[M - 1]; sm.f_Ws7[warp] = f_Ws[M - 1]; sm.f_V2l[warp] = f_V2; sm.f_W2l[warp] = f_W2; }
		if(lane == 1)  { sm.f_V2f[warp] = f_V2; sm.f_W2f[warp] = f_W2; }
		if(lane == 0)  { sm.f_a0[warp] = a[0]; sm.f_Bp[warp] = Bpart; sm.f_Ct[warp] = Ct; }
		__syncthreads();
		// level 3: W x W tridiagonal over the warp separators.  Lanes 0 .. W-1 of
		// warp 0 assemble one row each; lane 0 then runs the Thomas recurrence
		// (one reciprocal per row).
		if(warp == 0) {
			if(lane < W) {
				const int w = lane;
				const double a0 = sm.f_a0[w];
				const double Vs7q = w > 0 ? sm.f_Vs7[w - 1] : 0., Ws7q = w > 0 ? sm.f_Ws7[w - 1] : 0.;
				const double V2q = w > 0 ? sm.f_V2l[w - 1] : 0., W2q = w > 0 ? sm.f_W2l[w - 1] : 0.;
				const double At = -a0 * Vs7q;
				const double Bt = sm.f_Bp[w] - a0 * Ws7q;
				const double Cw = sm.f_Ct[w];
				sm.m3[w] = -At * V2q;                                // A3, turned into m3 below
				sm.invd3[w] = Bt - At * W2q - Cw * sm.f_V2f[w];      // B3, turned into 1 / d3 below
				sm.C3[w] = -Cw * sm.f_W2f[w];
				sm.q1[w] = -a0; sm.q2[w] = -At; sm.q3[w] = -Cw;
			}
			__syncwarp();
			if(lane == 0) {
				double cprev = 0., rprev = 0.;
				for(int w = 0; w < W; ++w) {
					const double m3 = sm.m3[w] * rprev;
					const double d3 = fma(-m3, cprev, sm.invd3[w]);
					rprev = rcp(d3);
					sm.m3[w] = m3; sm.invd3[w] = rprev;
					cprev = sm.C3[w];
				}
			}
		}
		__syncthreads();
	};

	// ------------------------------------------------------------------
	// One solve with the cached factorisation.  rhs r[] in, new iterate out.
	// Returns through xn[], xn_next; contains one block barrier.
	// ------------------------------------------------------------------
	auto solve = [&] (const double (&r)[M], double (&xn)[M], double &xn_next) {
		double G[M];
		G[1] = r[1];
#pragma unroll
		for(int j = 2; j < M; ++j) G[j] = fma(-f_m[j], G[j - 1], r[j]);
		G[M - 1] = G[M - 1] * f_invd[M - 1];
#pragma unroll
		for(int j = M - 2; j >= 1; --j) G[j] = fma(-c[j], G[j + 1], G[j]) * f_invd[j];
		const double G7p = shfl_up(G[M - 1], 1);
		const double P0 = fma(-c[0], G[1], r[0]);
		double R2 = lane >= 1 ? fma(-a[0], G7p, P0) : 0.;
#pragma unroll
		for(int s = 0; s < 5; ++s) {
			const int dl = 1 << s;
			const double Rm = shfl_up(R2, dl), Rq = shfl_down(R2, dl);
			R2 = fma(f_al[s], Rm, fma(f_ga[s], Rq, R2));   // al, ga are 0 where the partner is outside
		}
		const double G2 = R2 * f_invB2;
		if(lane == 31) { sm.s_G7[warp] = G[M - 1]; sm.s_G2l[warp] = G2; }
		if(lane == 1)  sm.s_G2f[warp] = G2;
		if(lane == 0)  sm.s_P0[warp] = P0;
		__syncthreads();
		// level 3: every thread runs the tiny Thomas recurrence itself
		double Y[MAX_WARPS + 1];
		{
			double z = 0.;
#pragma unroll
			for(int w = 0; w < MAX_WARPS; ++w) {
				if(w < W) {
					double R3 = sm.s_P0[w];
					if(w > 0) R3 = fma(sm.q1[w], sm.s_G7[w - 1], fma(sm.q2[w], sm.s_G2l[w - 1], R3));
					R3 = fma(sm.q3[w], sm.s_G2f[w], R3);
					z = fma(-sm.m3[w], z, R3);
					Y[w] = z;
				}
			}
			double yn = 0.;
#pragma unroll
			for(int w = MAX_WARPS - 1; w >= 0; --w) {
				if(w < W) {
					yn = fma(-sm.C3[w], yn, Y[w]) * sm.invd3[w];
					Y[w] = yn;
				}
			}
		}
		double Yw = 0., Yn = 0.;
#pragma unroll
		for(int w = 0; w < MAX_WARPS; ++w) {
			if(w == warp) Yw = Y[w];
			if(w == warp + 1 && w < W) Yn = Y[w];
		}
		const double ysep = lane == 0 ? Yw : fma(-Yn, f_W2, fma(-Yw, f_V2, G2));
		double ynext = shfl_down(ysep, 1);
		if(lane == 31) ynext = Yn;
		xn[0] = ysep;
#pragma unroll
		for(int j = 1; j < M; ++j) xn[j] = fma(-ynext, f_Ws[j], fma(-ysep, f_Vs[j], G[j]));
		xn_next = ynext;
	};

	// ------------------------------------------------------------------
	// Time loop
	// ------------------------------------------------------------------
	Replay rp; rp.start(b.T[cidx], b.dt[cidx], b.n_exercise);
	NodeState ns; ns.start(b.scheme);
	double time1 = rp.T, time0 = rp.T;
	int cur_kind = -1; double cur_h = 0.;
	int n_steps = 0, inner_total = 0, status = 0;
	bool more = true;
	bool refactor = true;        // block-uniform: the cached factorisation is stale
	unsigned flip = 0;           // block-uniform: which flag slot the next reduction uses
	const int jl_thread = (n_main - 1) / M, jl = (n_main - 1) % M; // owner of row N-2

	while(more) {
		qpde_step st;
		more = rp.next(st);
		const double t = st.t;
		const int kind = ns.kind();
		const double h0 = time1 - t;
		Gamma g; g.kind = kind; g.h = h0;
		double c1 = 0., c0 = 0., den = 1.;
		if(BDF2 && kind == STEP_BDF2) {
			const double h1 = time1 - t, hh0 = time0 - t;      // BDF.hpp:68-73
			g.h = (hh0 * h1) / (hh0 + h1);
			c1 = hh0 / h1; c0 = h1 / hh0; den = hh0 - h1;
		}
		// (Re)build the system rows when the step kind or size changes.  h0 is a
		// difference of accumulated times and moves in its last bits from step
		// to step (by ulp(t) / dt, ~1e-13 relative); moves below 1e-11 relative do not trigger a rebuild:
		// the step is then taken with an h that is off by that relative amount,
		// which perturbs the solution at the 1e-13 level (the reference itself
		// keeps a stale A for non-penalised roots, IterativeMethod.hpp:906).
		if(kind != cur_kind || fabs(g.h - cur_h) > 1e-11 * cur_h) {
			cur_kind = kind; cur_h = g.h;
#pragma unroll
			for(int j = 0; j < M; ++j) {
				a[j] = g.off(sm_lo[j * P + tid]);
				e[j] = g.off(sm_di[j * P + tid]);
				c[j] = g.off(sm_up[j * P + tid]);
			}
			e_tail = g.off(sm.di_tail);
			if(tid == jl_thread) {
#pragma unroll
				for(int j = 0; j < M; ++j) if(j == jl) { ctail = c[j]; c[j] = 0.; }
			}
			refactor = true;
		}

		// ---- lb = b(t) of the discretisation node (V^n = x) -------------
		{
			const double xm = sm_xlast[tid];       // x of row i0 - 1
#pragma unroll
			for(int j = 0; j < M; ++j) {
				const double vm = j == 0 ? xm : x[j - 1];
				const double vp = j == M - 1 ? x_next : x[j + 1];
				double val;
				if(kind == STEP_IMPLICIT) {
					val = x[j] + 0. * h0;
				} else if(BDF2 && kind == STEP_BDF2) {
					val = ((c1 * x[j] - c0 * vprev[j]) / den + 0.) * g.h;
				} else {
					double acc = 0.;
					acc += (0. - a[j]) * vm;
					acc += (1. - e[j]) * x[j];
					acc += (0. - c[j]) * vp;
					if(tid == jl_thread && j == jl) acc += (0. - ctail) * x_tail;
					val = acc + 0.;
				}
				lb[j] = val;
			}
			if(kind == STEP_IMPLICIT) lb_tail = x_tail + 0. * h0;
			else if(BDF2 && kind == STEP_BDF2) lb_tail = ((c1 * x_tail - c0 * vprev_tail) / den + 0.) * g.h;
			else lb_tail = (0. + (1. - e_tail) * x_tail) + 0.;
			if(BDF2) {
#pragma unroll
				for(int j = 0; j < M; ++j) vprev[j] = x[j];
				vprev_tail = x_tail;
			}
		}

		int its = 0;
		bool again;
		do {
			if(refactor) {
				curmask = newmask;
				factorize();
				refactor = false;
			}
			// ---- right-hand side with penalty terms, tail row, fold ----------
			double r[M], xn[M], xn_next;
#pragma unroll
			for(int j = 0; j < M; ++j) {
				r[j] = lb[j];
				if(AMERICAN && (curmask >> j & 1u)) r[j] = r[j] + pen * pz[j];
			}
			double bt = 1. + e_tail, rt = lb_tail;
			if(AMERICAN && (curmask >> M & 1u)) { bt = bt + pen * 1.; rt = rt + pen * pz_tail; }
			const double xt_new = rt / bt;
			if(tid == jl_thread) {
#pragma unroll
				for(int j = 0; j < M; ++j) if(j == jl) r[j] = fma(-ctail, xt_new, r[j]);
			}
			solve(r, xn, xn_next);
			// ---- relativeError + next active set -----------------------------
			bool nd = false;
			if(AMERICAN) {
#pragma unroll
				for(int j = 0; j < M; ++j) nd = nd || not_converged(xn[j], x[j], scale, tol);
				nd = nd || not_converged(xt_new, x_tail, scale, tol);
			}
#pragma unroll
			for(int j = 0; j < M; ++j) x[j] = xn[j];
			x_tail = xt_new;
			x_next = xn_next;
			sm_xlast[tid + 1] = x[M - 1];
			++its;
			if(AMERICAN) {
				newmask = mask_of();
				// one barrier carries both block-wide flags
				const unsigned bits = (nd ? 1u : 0u) | (newmask != curmask ? 2u : 0u);
				const unsigned wbits = __reduce_or_sync(FULL, bits);
				const unsigned slot = flip & 1u;   // two slots used alternately
				++flip;
				if(lane == 0 && wbits) atomicOr(&sm.flags[slot], wbits);
				__syncthreads();
				const unsigned all = *(volatile unsigned *)&sm.flags[slot];
				if(tid == 0) sm.flags[slot ^ 1u] = 0u;       // ordered by the barrier inside the next solve
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
				__syncthreads();
				again = false;
			}
		} while(again);

		if(tid == 0 && out.inner) {
			out.inner[(size_t)cidx * out.inner_stride + n_steps] = (unsigned char)(its > 255 ? 255 : its);
		}
		inner_total += its;
		++n_steps;
		time0 = time1; time1 = t;
		ns.end_step();

		if(st.event_after) {
			if(EVENTS && st.user_events > 0) {
				// Event<1>::_doEvent with max(V(S), K - S): Core/Event.hpp:100-112
#pragma unroll
				for(int j = 0; j < M; ++j) x[j] = x[j] > pz[j] ? x[j] : pz[j];
				x_tail = x_tail > pz_tail ? x_tail : pz_tail;
				sm_xfirst[tid] = x[0];
				sm_xlast[tid + 1] = x[M - 1];
				if(tid == 0) sm_xfirst[P] = 0.;
				__syncthreads();
				x_next = sm_xfirst[tid + 1];
			}
			ns.after_event();
		}
	}

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
				out.active[(size_t)cidx * out.active_stride + i] = pred < 0. ? 1 : 0;
			}
		}
	}
	if(tid == 0) {
		if(out.V) out.V[(size_t)cidx * out.V_stride + N - 1] = x_tail;
		if(out.S) out.S[(size_t)cidx * out.S_stride + N - 1] = sm.S_tail;
		if(out.active) {
			const double pred = AMERICAN ? (0. + 1. * x_tail) - pz_tail : 0.;
			out.active[(size_t)cidx * out.active_stride + N - 1] = pred < 0. ? 1 : 0;
		}
		if(out.n_steps) out.n_steps[cidx] = n_steps;
		if(out.inner_total) out.inner_total[cidx] = inner_total;
		if(out.status) out.status[cidx] = status;
	}
	if(out.value) {
		// PiecewiseLinear::interpolate (Core/Interpolant.hpp:153-181,331-374):
		// the thread owning the bracketing interval writes the value
		const double s0 = b.spot ? b.spot[cidx] : K;
		const double Sfirst = S_at(0), Slast = sm.S_tail;
#pragma unroll
		for(int j = 0; j < M; ++j) {
			const int i = i0 + j;
			if(i >= n_main) continue;
			const double Si = sm_S[j * P + tid];
			const double Sp = i + 1 == N - 1 ? sm.S_tail : S_at(i + 1);
			const double vi = x[j];
			const double vp = j == M - 1 ? (i + 1 == N - 1 ? x_tail : x_next) : (i + 1 == N - 1 ? x_tail : x[j + 1]);
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

// Rows per thread for a grid of N nodes: 8 up to 2049 nodes, 9 up to 2305.
static int rows_per_thread(int N) {
	const int n_main = N - 1;
	if(n_main <= 8 * MAX_THREADS) return 8;
	if(n_main <= 9 * MAX_THREADS) return 9;
	return 0;
}

bool resident_supported(const DeviceBatch &b) {
	if(b.N < 10) return false;
	if(rows_per_thread(b.N) == 0) return false;
	if(b.american && b.n_exercise > 0) return false; // penalty + events: INTERLEAVED family
	return true;
}

template <int M, bool AMERICAN, bool BDF2, bool EVENTS>
static int launch_variant(const DeviceBatch &b, const DeviceResult &out, cudaStream_t stream) {
	auto kern = k_resident_sweep<M, AMERICAN, BDF2, EVENTS>;
	const int n_main = b.N - 1;
	int threads = (n_main + M - 1) / M;
	threads = (threads + 31) / 32 * 32;
	static bool configured = false;
	if(!configured) {
		if(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
				(int)smem_bytes(M, MAX_THREADS)) != cudaSuccess) {
			return QPDE_ERR_CUDA;
		}
		configured = true;
	}
	kern<<<b.C, threads, smem_bytes(M, threads), stream>>>(b, out);
	return QPDE_OK;
}

template <int M>
static int launch_rows(const DeviceBatch &b, const DeviceResult &out, cudaStream_t stream) {
	const bool bdf2 = b.scheme == QPDE_SCHEME_BDF2;
	const bool events = b.n_exercise > 0;
	if(b.american) return bdf2 ? launch_variant<M, true, true, false>(b, out, stream)
	                           : launch_variant<M, true, false, false>(b, out, stream);
	if(events)     return bdf2 ? launch_variant<M, false, true, true>(b, out, stream)
	                           : launch_variant<M, false, false, true>(b, out, stream);
	return bdf2 ? launch_variant<M, false, true, false>(b, out, stream)
	            : launch_variant<M, false, false, false>(b, out, stream);
}

int launch_resident(const DeviceBatch &b, const DeviceResult &out, cudaStream_t stream,
		int sm_count, long long *launches) {
	(void)sm_count;
	if(!resident_supported(b)) return QPDE_ERR_UNSUPPORTED;
	const int rc = rows_per_thread(b.N) == 8 ? launch_rows<8>(b, out, stream)
	                                         : launch_rows<9>(b, out, stream);
	if(rc == QPDE_OK) *launches += 1;
	return rc;
}

} // namespace qpde
