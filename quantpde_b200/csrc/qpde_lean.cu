// qpde_lean.cu — the RESIDENT sweep at TWO contracts per SM.
//
// k_resident_sweep<8, 256> (qpde_resident.cu) keeps the cached factorisation of a contract in
// registers (252 per thread) and eleven row arrays in shared memory (188 KB): one CTA per SM,
// eight warps, and the kernel is bound by the latency of its own dependent instructions (ncu,
// round 1: issue slots 30 % busy, FP64 pipe 18 %).  This variant holds the same state in half
// the space so that two independent contracts share an SM and fill each other's stalls:
//
//   registers (<= 128)  : the iterate x, the reciprocal pivots and the back-substitution
//                         multipliers u of the thread's rows, three scalars of the warp-level
//                         reduction
//   shared memory       : system rows a, me, c (me = 1 - e is the Crank-Nicolson right-hand-side
//   (106 KB)              diagonal; the system diagonal 1 + e is re-formed as 2 - me, within one
//                         ulp, when a factorisation needs it; the off-diagonal rows of the
//                         right-hand side are 0 - a, 0 - c), the right-hand side lb, the obstacle,
//                         the coefficients of the warp-level cyclic reduction, the level-3 factors
//   recomputed          : the operator rows of BlackScholes::A at the (rare) re-scalings; the
//                         forward multipliers a_j / d_{j-1} on the way (one product per row)
//
// and the partition solve needs no spike vectors: once the separator unknowns are known
// (levels 2 and 3, as in qpde_partition.cuh) every thread solves its interior rows again with
// the two boundary values moved to the right-hand side — a second Thomas pass instead of two
// stored vectors per row.
//
// Same reference call stack as qpde_resident.cu (TimeIteration / ToleranceIteration /
// PenaltyMethodDifference / Rannacher, CrankNicolson, BDFOne / BlackScholes::A / SparseLU);
// MODE_PENALTY sweeps with constant steps and a generated obstacle, 513 < N <= 2049.
#include <math_constants.h>

#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "qpde_kernels.cuh"
#include "qpde_partition.cuh"

namespace qpde {

namespace {

constexpr int LP = 256;          // threads per contract
constexpr int LW = LP / 32;      // warps per contract
constexpr unsigned FULL = FULL_MASK;

struct LeanTail { double S, di, x, pz, bd, me, lb, invb, pq, c; };

template <int M>
struct LeanShared {
	// hand-over between the warps: right-hand side ...
	double s_G7[LW], s_G2l[LW], s_P0[LW], s_G2f[LW];
	// ... and factorisation
	double f_Vs7[LW], f_Ws7[LW], f_V2l[LW], f_W2l[LW], f_V2f[LW], f_W2f[LW], f_a0[LW], f_Bp[LW], f_Ct[LW];
	// level-3 factors (one entry per warp separator; written by warp 0, read by every warp)
	double q1[LW], q2[LW], q3[LW], al3[3][LW], ga3[3][LW], invB3[LW];
	LeanTail tail;
	int n_steps, inner_total, computed, status;   // kept by thread 0
	double xlast[LP + 1];         // x of each thread's last row, shifted by one
	// row arrays: rows (2p, 2p + 1) of thread t as one double2 at [p][t] — every access is a full
	// 16-byte one (half the load / store instructions of an [j][t] layout, no bank conflicts)
	double2 a2[M / 2 * LP], me2[M / 2 * LP], c2[M / 2 * LP];
	double2 lb2[M / 2 * LP];       // right-hand side lb of the discretisation node
	double2 pz2[M / 2 * LP];       // obstacle
	// warp-level cyclic reduction, [stage][thread]
	double al[5 * LP], ga[5 * LP];
};

// relativeError's per-element test, as in qpde_resident.cu (integer screen + exact fallback)
__device__ __forceinline__ void conv_track(double xn, double xo, int &dmax_hi, int &xmax_hi) {
	dmax_hi = max(dmax_hi, __double2hiint(fabs(xn - xo)));
	xmax_hi = max(xmax_hi, __double2hiint(fabs(xn)));
}
__device__ __forceinline__ int conv_screen(int dmax_hi, int xmax_hi, double tol, double scale) {
	const double D_up = __hiloint2double(dmax_hi + 1, 0);
	if(D_up <= tol * scale * (1. - 8e-16)) return 0;
	const double D_dn = __hiloint2double(dmax_hi, 0);
	const double XN_up = __hiloint2double(xmax_hi + 1, 0);
	return D_dn > tol * (1. + 8e-16) * (XN_up + D_up + scale) ? 1 : 2;
}
__device__ __forceinline__ bool conv_exact(double xn, double xo, double tol, double scale) {
	const double fa = fabs(xn), fb = fabs(xo);
	double den = fa < fb ? fb : fa;
	den = scale < den ? den : scale;
	return fabs(xn - xo) / den > tol;
}

} // namespace

// One CTA of 256 threads per contract, two CTAs per SM.  OBST: the obstacle generator
// (QPDE_PAYOFF_PUT or QPDE_PAYOFF_CALL), compile-time so that the per-row obstacle is three
// branch-free instructions.
template <int M, int OBST>
__global__ void __launch_bounds__(LP, 2)
k_lean_sweep(DeviceBatch b, DeviceResult out) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	LeanShared<M> &sm = *reinterpret_cast<LeanShared<M> *>(smem_raw);

	const int cidx = b.c0 + (int)blockIdx.x;
	const int tid = threadIdx.x;
	const int lane = tid & 31, warp = tid >> 5;
	const int N = b.N;
	const bool has_tail = N == LP * M + 1;
	const int n_main = has_tail ? N - 1 : N;
	const int i0 = tid * M;
	const bool tail_owner = has_tail && tid == LP - 1;

	const double K = b.strike[cidx];
	const double *const axis = b.axis + (size_t)cidx * b.axis_stride;
	const double pen = b.pen;

	// ---- system rows for one step kind / size: BlackScholes::A rows recomputed, then scaled
	// (Rannacher.hpp:32-58, CrankNicolson.hpp:59-62, BDF.hpp:32-46)
	const double r_c = b.r[cidx], q_c = b.q[cidx], sig_c = b.sigma[cidx];
	auto rescale = [&] (const Gamma &g) {
		for(int p2 = 0; p2 < M / 2; ++p2) {
			double ra[2], rme[2], rc[2];
#pragma unroll
			for(int k = 0; k < 2; ++k) {
				const int j = 2 * p2 + k;
				const int i = i0 + j;
				double aj = 0., ej = 0., cj = 0.;
				if(i < n_main) {
					const double Si = refined_node(axis, b.refine, i);
					const double Sm = i > 0 ? refined_node(axis, b.refine, i - 1) : 0.;
					const double Sp = i < N - 1 ? refined_node(axis, b.refine, i + 1) : 0.;
					const double v_i = b.vol_model == QPDE_VOL_NODES
						? b.sigma_nodes[(size_t)cidx * b.sigma_nodes_stride + i]
						: vol_value(b.vol_model, sig_c, Si);
					const Row row = bs_row(i, N, Sm, Si, Sp, r_c, v_i, q_c);
					aj = g.off(row.lo); ej = g.off(row.di); cj = g.off(row.up);
				}
				if(j == M - 1 && tail_owner) { sm.tail.c = cj; cj = 0.; }
				ra[k] = aj; rme[k] = 1. - ej; rc[k] = cj;
			}
			const int at = p2 * LP + tid;
			sm.a2[at] = make_double2(ra[0], ra[1]); sm.me2[at] = make_double2(rme[0], rme[1]); sm.c2[at] = make_double2(rc[0], rc[1]);
		}
		if(tail_owner) {
			const double et = g.off(sm.tail.di);
			sm.tail.bd = 1. + et; sm.tail.me = 1. - et;
		}
	};

	// ---- per-thread state
	double x[M];
	double f_invd[M], f_u[M];         // reciprocal pivots and back-substitution multipliers of rows 1 .. M-1 (index 0 unused)
	double f_invB2 = 1., f_V2 = 0., f_W2 = 0.;
	double x_next;
	unsigned curmask = 0, newmask = 0;
#pragma unroll
	for(int j = 0; j < M; ++j) { f_invd[j] = 1.; f_u[j] = 0.; }

	auto initial_value = [&] (int i) -> double {
		return b.payoff_kind == QPDE_PAYOFF_ARRAY
			? b.payoff[(size_t)cidx * b.payoff_stride + i]
			: payoff_value(b.payoff_kind, refined_node(axis, b.refine, i), K);
	};
#pragma unroll
	for(int j = 0; j < M; ++j) x[j] = i0 + j < n_main ? initial_value(i0 + j) : 0.;
	bool z0 = true;                   // every obstacle value of this thread is <= 0
#pragma unroll
	for(int p2 = 0; p2 < M / 2; ++p2) {
		// Modules/Lambdas/Payoffs.hpp:29-31,38-40; rows past the grid never bind
		double2 q;
		q.x = i0 + 2 * p2 < n_main ? payoff_value(OBST, refined_node(axis, b.refine, i0 + 2 * p2), K) : -CUDART_INF;
		q.y = i0 + 2 * p2 + 1 < n_main ? payoff_value(OBST, refined_node(axis, b.refine, i0 + 2 * p2 + 1), K) : -CUDART_INF;
		sm.pz2[p2 * LP + tid] = q;
		z0 = z0 && !(q.x > 0.) && !(q.y > 0.);
	}
	if(tail_owner) {
		sm.tail.S = refined_node(axis, b.refine, N - 1);
		sm.tail.di = q_c;                                   // row N-1 of BlackScholes::A: [q] (BlackScholes.hpp:177-180)
		sm.tail.x = initial_value(N - 1);
		sm.tail.pz = payoff_value(OBST, sm.tail.S, K);
		sm.tail.c = 0.; sm.tail.bd = 1.; sm.tail.me = 1.; sm.tail.invb = 1.; sm.tail.pq = 0.; sm.tail.lb = 0.;
	}
	x_next = i0 + M < n_main ? initial_value(i0 + M) : 0.;
	sm.xlast[tid + 1] = x[M - 1];
	if(tid == 0) sm.xlast[0] = 0.;

	// PenaltyMethod.hpp:45,61-66: row penalised iff x < payoff
#define QPDE_LEAN_MASK(dst) \
	do { \
		unsigned mk_ = 0; \
		_Pragma("unroll") \
		for(int p2 = 0; p2 < M / 2; ++p2) { \
			const double2 q_ = wpz0 ? make_double2(0., 0.) : sm.pz2[p2 * LP + tid]; \
			if(x[2 * p2] < q_.x) mk_ |= 1u << (2 * p2); \
			if(x[2 * p2 + 1] < q_.y) mk_ |= 1u << (2 * p2 + 1); \
		} \
		if(tail_owner && sm.tail.x < sm.tail.pz) mk_ |= 1u << M; \
		dst = mk_; \
	} while(0)
	// warps whose rows all have a zero obstacle (out of the money) test x < 0 and never load it
	if(tail_owner) z0 = z0 && !(sm.tail.pz > 0.);
	const bool wpz0 = __all_sync(FULL, z0);
	bool wpen = true;                 // some row of this warp is penalised (set at every factorisation)
	__syncthreads();
	QPDE_LEAN_MASK(newmask);

	// ---- time loop (TimeIteration<false>, Core/IterativeMethod.hpp:1259-1317)
	// Replay::next (qpde_common.cuh) without user events: the only event is the NullEvent at 0
	const double dt_c = b.dt[cidx];
	double time1 = b.T[cidx];
	int phase = 1;                    // steps since the start, saturating at 3 (NodeState)
	int cur_kind = -1; double cur_h = 0.;
	if(tid == 0) { sm.n_steps = 0; sm.inner_total = 0; sm.computed = 0; sm.status = 0; }
	bool more = true;
	bool refactor = true;             // block-wide: some warp's rows or active set changed
	bool wrefactor = true;            // this warp's rows or active set changed: levels 1 and 2 of the factorisation are
	                                  // functions of the warp's own rows, so the other warps keep theirs

	while(more) {
		double t = time1 + -1. * dt_c;
		if(fabs(t - 0.) < DBL_EPSILON) t = 0.;
		else if(t < 0.) t = 0.;
		more = 0. < t + -1. * DBL_EPSILON;
		const int kind = b.scheme == QPDE_SCHEME_RANNACHER ? (phase >= 3 ? STEP_CN_R : STEP_IMPLICIT)
			: (b.scheme == QPDE_SCHEME_CN ? STEP_CN_T : STEP_IMPLICIT);
		const double h0 = time1 - t;
		Gamma g; g.kind = kind; g.h = h0;
		// the rows are rebuilt when the step kind changes or the step size moves by more than
		// 1e-11 relative (qpde_resident.cu: the same rule and its justification)
		if(kind != cur_kind || fabs(g.h - cur_h) > 1e-11 * cur_h) {
			cur_kind = kind; cur_h = g.h;
			rescale(g);
			refactor = true; wrefactor = true;
		}

		// ---- lb = b(t) of the discretisation node (V^n = x)
		if(kind == STEP_IMPLICIT) {
#pragma unroll
			for(int p2 = 0; p2 < M / 2; ++p2) sm.lb2[p2 * LP + tid] = make_double2(x[2 * p2] + 0. * h0, x[2 * p2 + 1] + 0. * h0);
			if(tail_owner) sm.tail.lb = sm.tail.x + 0. * h0;
		} else {
			// (I - A h/2) V^n: row-major SpMV accumulated from zero in column order (Rannacher.hpp:60-73)
			const double xm = sm.xlast[tid];
#pragma unroll
			for(int p2 = 0; p2 < M / 2; ++p2) {
				const int at = p2 * LP + tid;
				const double2 a2 = sm.a2[at], me2 = sm.me2[at], c2 = sm.c2[at];
				double lbv[2];
#pragma unroll
				for(int k = 0; k < 2; ++k) {
					const int j = 2 * p2 + k;
					const double vm = j == 0 ? xm : x[j > 0 ? j - 1 : 0];
					const double vp = j == M - 1 ? x_next : x[j + 1 < M ? j + 1 : j];
					const double na = 0. - (k ? a2.y : a2.x), me = k ? me2.y : me2.x, nc = 0. - (k ? c2.y : c2.x);
					double acc = 0.;
					acc += na * vm;
					acc += me * x[j];
					acc += nc * vp;
					if(j == M - 1 && tail_owner) acc += (0. - sm.tail.c) * sm.tail.x;
					lbv[k] = acc + 0.;
				}
				sm.lb2[at] = make_double2(lbv[0], lbv[1]);
			}
			if(tail_owner) sm.tail.lb = (0. + sm.tail.me * sm.tail.x) + 0.;
		}

		int its = 0;
		bool again;
		do {
			// ==========================================================
			// factorisation of (lA + P) for the current rows and active set
			// ==========================================================
			if(refactor) {
				refactor = false;
				if(wrefactor) {
				wrefactor = false;
				curmask = newmask;
				wpen = __any_sync(FULL, curmask != 0u);
				double Vs1, Ws1, Vs7, Ws7, a0, c0, b0;
				{
					double aa[M], bb[M], cc[M], fm[M];
#pragma unroll
					for(int p2 = 0; p2 < M / 2; ++p2) {
						const int at = p2 * LP + tid;
						const double2 a2 = sm.a2[at], me2 = sm.me2[at], c2 = sm.c2[at];
						aa[2 * p2] = a2.x; aa[2 * p2 + 1] = a2.y; cc[2 * p2] = c2.x; cc[2 * p2 + 1] = c2.y;
						bb[2 * p2] = 2. - me2.x; bb[2 * p2 + 1] = 2. - me2.y;          // 1 + e within one ulp
					}
#pragma unroll
					for(int j = 0; j < M; ++j) {
						if(curmask >> j & 1u) bb[j] = bb[j] + pen * 1.;      // PenaltyMethod.hpp:96-119
					}
					a0 = aa[0]; c0 = cc[0]; b0 = bb[0];
					f_invd[1] = rcp(bb[1]);
#pragma unroll
					for(int j = 2; j < M; ++j) {
						fm[j] = aa[j] * f_invd[j - 1];
						f_invd[j] = rcp(fma(-fm[j], cc[j - 1], bb[j]));
					}
					double yv[M];
					yv[1] = aa[1];
#pragma unroll
					for(int j = 2; j < M; ++j) yv[j] = -fm[j] * yv[j - 1];
#pragma unroll
					for(int j = 1; j < M; ++j) f_u[j] = cc[j] * f_invd[j];
					// spikes of the interior block: only their first and last entries are needed
					double Vs = yv[M - 1] * f_invd[M - 1], Ws = f_u[M - 1];
					Vs7 = Vs; Ws7 = Ws;
#pragma unroll
					for(int j = M - 2; j >= 1; --j) {
						Vs = fma(-f_u[j], Vs, yv[j] * f_invd[j]);
						Ws = -f_u[j] * Ws;
					}
					Vs1 = Vs; Ws1 = Ws;
				}
				if(tail_owner) {
					const bool act = curmask >> M & 1u;
					sm.tail.invb = rcp(act ? sm.tail.bd + pen * 1. : sm.tail.bd);
					sm.tail.pq = act ? pen * sm.tail.pz : 0.;
				}
				// reduced row of this thread's separator (its row 0); levels 2 and 3 as in
				// qpde_partition.cuh, coefficients to shared memory
				const double Vs7p = shfl_up(Vs7, 1), Ws7p = shfl_up(Ws7, 1);
				const bool interior = lane >= 1;
				const double Bpart = fma(-c0, Vs1, b0);
				const double Ct = -c0 * Ws1;
				double A2 = interior ? -a0 * Vs7p : 0.;
				double B2 = interior ? fma(-a0, Ws7p, Bpart) : 1.;
				double C2 = interior ? Ct : 0.;
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
					sm.al[s * LP + tid] = al; sm.ga[s * LP + tid] = ga;
					B2 = fma(al, Cm, fma(ga, Aq, B2));
					V2 = fma(al, Vm, fma(ga, Vq, V2));
					W2 = fma(al, Wm, fma(ga, Wq, W2));
					A2 = al * Am;
					C2 = ga * Cq;
				}
				f_invB2 = rcp(B2);
				f_V2 = V2 * f_invB2;
				f_W2 = W2 * f_invB2;
				if(lane == 31) { sm.f_Vs7[warp] = Vs7; sm.f_Ws7[warp] = Ws7; sm.f_V2l[warp] = f_V2; sm.f_W2l[warp] = f_W2; }
				if(lane == 1)  { sm.f_V2f[warp] = f_V2; sm.f_W2f[warp] = f_W2; }
				if(lane == 0)  { sm.f_a0[warp] = a0; sm.f_Bp[warp] = Bpart; sm.f_Ct[warp] = Ct; }
				}
				__syncthreads();
				if(warp == 0) {
					// level 3: LW x LW tridiagonal over the warp separators, lanes 0 .. LW-1
					double A3 = 0., B3 = 1., C3 = 0.;
					double q1 = 0., q2 = 0., q3 = 0.;
					if(lane < LW) {
						const int w = lane;
						const double wa0 = sm.f_a0[w];
						const double Vs7q = w > 0 ? sm.f_Vs7[w - 1] : 0., Ws7q = w > 0 ? sm.f_Ws7[w - 1] : 0.;
						const double V2q = w > 0 ? sm.f_V2l[w - 1] : 0., W2q = w > 0 ? sm.f_W2l[w - 1] : 0.;
						const double At = -wa0 * Vs7q;
						const double Bt = fma(-wa0, Ws7q, sm.f_Bp[w]);
						const double Cw = sm.f_Ct[w];
						A3 = -At * V2q;
						B3 = fma(-Cw, sm.f_V2f[w], fma(-At, W2q, Bt));
						C3 = -Cw * sm.f_W2f[w];
						q1 = -wa0; q2 = -At; q3 = -Cw;
					}
#pragma unroll
					for(int s = 0; s < 3; ++s) {
						const int dl = 1 << s;
						const double rB = rcp(B3);
						const double rBm = shfl_up(rB, dl), rBq = shfl_down(rB, dl);
						const double Am = shfl_up(A3, dl), Aq = shfl_down(A3, dl);
						const double Cm = shfl_up(C3, dl), Cq = shfl_down(C3, dl);
						const bool okm = lane < LW && lane - dl >= 0;
						const bool okq = lane + dl < LW;
						const double al = okm ? -A3 * rBm : 0.;
						const double ga = okq ? -C3 * rBq : 0.;
						if(lane < LW) { sm.al3[s][lane] = al; sm.ga3[s][lane] = ga; }
						B3 = fma(al, Cm, fma(ga, Aq, B3));
						A3 = al * Am;
						C3 = ga * Cq;
					}
					if(lane < LW) { sm.q1[lane] = q1; sm.q2[lane] = q2; sm.q3[lane] = q3; sm.invB3[lane] = rcp(B3); }
				}
				// the level-3 factors are read after the barrier of the solve below
			}

			// ==========================================================
			// one solve with the cached factorisation
			// ==========================================================
			double xt_new = 0., tail_c = 0.;
			if(tail_owner) { xt_new = (sm.tail.lb + sm.tail.pq) * sm.tail.invb; tail_c = sm.tail.c; }
			// right-hand side with penalty terms (PenaltyMethod.hpp:96-119); the tail row is a scalar
			// equation folded into row N-2
			// (a warp without penalised rows takes lb as it is: `pen_tag` is a compile-time tag)
			// rows (2p, 2p + 1): one 16-byte load of lb, one of the obstacle
			auto rhs2 = [&] (auto pen_tag, int p2, double &re, double &ro) {
				const double2 l2 = sm.lb2[p2 * LP + tid];
				re = l2.x; ro = l2.y;
				if(decltype(pen_tag)::value) {
					const double2 q2 = sm.pz2[p2 * LP + tid];
					const double pe = l2.x + pen * q2.x, po = l2.y + pen * q2.y;   // formed for every row, selected below: no branch
					re = (curmask >> (2 * p2) & 1u) ? pe : l2.x;
					ro = (curmask >> (2 * p2 + 1) & 1u) ? po : l2.y;
				}
				if(p2 == M / 2 - 1) ro = fma(-tail_c, xt_new, ro);      // both 0 except for the tail owner
			};
			double G1, G7, r0, a0, a1;
			auto pass1 = [&] (auto pen_tag) {
				// pass 1: interior rows with zero boundary values; only the first and last are kept
				double z[M];
#pragma unroll
				for(int p2 = 0; p2 < M / 2; ++p2) {
					double re, ro;
					rhs2(pen_tag, p2, re, ro);
					const double2 a2 = sm.a2[p2 * LP + tid];
					if(p2 == 0) { r0 = re; z[1] = ro; a0 = a2.x; a1 = a2.y; }
					else {
						z[2 * p2] = fma(-(a2.x * f_invd[2 * p2 - 1]), z[2 * p2 - 1], re);
						z[2 * p2 + 1] = fma(-(a2.y * f_invd[2 * p2]), z[2 * p2], ro);
					}
				}
				double G = z[M - 1] * f_invd[M - 1];
				G7 = G;
#pragma unroll
				for(int j = M - 2; j >= 1; --j) G = fma(-f_u[j], G, z[j] * f_invd[j]);
				G1 = G;
			};
			if(wpen) pass1(std::true_type()); else pass1(std::false_type());
			const double c0 = sm.c2[tid].x;
			const double G7p = shfl_up(G7, 1);
			const double P0 = fma(-c0, G1, r0);
			double R2 = lane >= 1 ? fma(-a0, G7p, P0) : 0.;
#pragma unroll
			for(int s = 0; s < 5; ++s) {
				const int dl = 1 << s;
				const double Rm = shfl_up(R2, dl), Rq = shfl_down(R2, dl);
				R2 = fma(sm.al[s * LP + tid], Rm, fma(sm.ga[s * LP + tid], Rq, R2));
			}
			const double G2 = R2 * f_invB2;
			if(lane == 31) { sm.s_G7[warp] = G7; sm.s_G2l[warp] = G2; }
			if(lane == 1)  sm.s_G2f[warp] = G2;
			if(lane == 0)  sm.s_P0[warp] = P0;
			__syncthreads();
			double R3 = 0., invB3 = 0.;
			if(lane < LW) {
				R3 = sm.s_P0[lane];
				if(lane > 0) R3 = fma(sm.q1[lane], sm.s_G7[lane - 1], fma(sm.q2[lane], sm.s_G2l[lane - 1], R3));
				R3 = fma(sm.q3[lane], sm.s_G2f[lane], R3);
				invB3 = sm.invB3[lane];
			}
#pragma unroll
			for(int s = 0; s < 3; ++s) {
				const int dl = 1 << s;
				const double Rm = shfl_up(R3, dl), Rq = shfl_down(R3, dl);
				const double al = lane < LW ? sm.al3[s][lane] : 0., ga = lane < LW ? sm.ga3[s][lane] : 0.;
				R3 = fma(al, Rm, fma(ga, Rq, R3));
			}
			const double Y = R3 * invB3;                 // separator of warp `lane` (0 for lanes >= LW)
			const double Yw = shfl_idx(Y, warp);
			const double Yn = shfl_idx(Y, warp + 1);     // lane LW holds 0
			const double ysep = lane == 0 ? Yw : fma(-Yn, f_W2, fma(-Yw, f_V2, G2));
			double ynext = shfl_down(ysep, 1);
			if(lane == 31) ynext = Yn;
			// pass 2: the interior rows again, boundary values on the right-hand side
			double y[M];
			auto pass2 = [&] (auto pen_tag) {
				double z[M];
#pragma unroll
				for(int p2 = 0; p2 < M / 2; ++p2) {
					double re, ro;
					rhs2(pen_tag, p2, re, ro);
					if(p2 == 0) z[1] = fma(-a1, ysep, ro);
					else {
						const double2 a2 = sm.a2[p2 * LP + tid];
						z[2 * p2] = fma(-(a2.x * f_invd[2 * p2 - 1]), z[2 * p2 - 1], re);
						z[2 * p2 + 1] = fma(-(a2.y * f_invd[2 * p2]), z[2 * p2], ro);
					}
				}
				// c of the last row is 0 for the tail owner (the coupling went to tail.c) and ynext is 0 there
				y[M - 1] = fma(-sm.c2[(M / 2 - 1) * LP + tid].y, ynext, z[M - 1]) * f_invd[M - 1];
#pragma unroll
				for(int j = M - 2; j >= 1; --j) y[j] = fma(-f_u[j], y[j + 1], z[j] * f_invd[j]);
				y[0] = ysep;
			};
			if(wpen) pass2(std::true_type()); else pass2(std::false_type());
			x_next = ynext;

			bool nd = false;
			{
				int dmax_hi = 0, xmax_hi = 0;
#pragma unroll
				for(int j = 0; j < M; ++j) conv_track(y[j], x[j], dmax_hi, xmax_hi);
				if(tail_owner) conv_track(xt_new, sm.tail.x, dmax_hi, xmax_hi);
				const int verdict = conv_screen(dmax_hi, xmax_hi, b.tolerance, b.scale);
				nd = verdict == 1;
				if(verdict == 2) {
#pragma unroll
					for(int j = 0; j < M; ++j) nd = nd | conv_exact(y[j], x[j], b.tolerance, b.scale);
					if(tail_owner) nd = nd | conv_exact(xt_new, sm.tail.x, b.tolerance, b.scale);
				}
			}
#pragma unroll
			for(int j = 0; j < M; ++j) x[j] = y[j];
			if(tail_owner) sm.tail.x = xt_new;
			sm.xlast[tid + 1] = x[M - 1];
			++its;
			QPDE_LEAN_MASK(newmask);
			// one barrier carries both block-wide flags: lanes 0 .. 30 of a warp vote "active set changed",
			// lane 31 votes "not converged"; the barrier's population count is 31 x (warps changed) +
			// (warps not converged), and both are at most 8
			const unsigned bits = (nd ? 1u : 0u) | (newmask != curmask ? 2u : 0u);
			const unsigned wbits = __reduce_or_sync(FULL, bits);
			const int votes = __syncthreads_count(lane == 31 ? (wbits & 1u) : (wbits >> 1 & 1u));
			const bool notdone = votes % 31 != 0, changed = votes >= 31;
			refactor = changed;
			wrefactor = (wbits >> 1 & 1u) != 0u;
			again = notdone;
			if(tid == 0) ++sm.computed;
			if(notdone && !changed) {
				// same active set => the next inner iteration would return this iterate again
				// (relativeError == 0): counted, not recomputed
				if(its >= b.max_inner) { if(tid == 0) sm.status = 1; } else ++its;
				again = false;
			} else if(again && its >= b.max_inner) {
				if(tid == 0) sm.status = 1;
				again = false;
			}
		} while(again);

		if(tid == 0) {
			const int n = sm.n_steps;
			if(out.inner && n < b.max_steps) out.inner[(size_t)cidx * out.inner_stride + n] = (unsigned char)(its > 255 ? 255 : its);
			sm.inner_total += its;
			sm.n_steps = n + 1;
		}
		time1 = t;
		if(phase < 3) ++phase;
	}
#undef QPDE_LEAN_MASK

	// ---- outputs
#pragma unroll
	for(int j = 0; j < M; ++j) {
		const int i = i0 + j;
		if(i < n_main) {
			if(out.V) out.V[(size_t)cidx * out.V_stride + i] = x[j];
			if(out.S) out.S[(size_t)cidx * out.S_stride + i] = refined_node(axis, b.refine, i);
			if(out.active) {
				const double2 q2 = sm.pz2[(j >> 1) * LP + tid];
				const double pred = (0. + 1. * x[j]) - ((j & 1) ? q2.y : q2.x);
				out.active[(size_t)cidx * out.active_stride + i] = pred < 0. ? 1 : 0;
			}
		}
	}
	__syncthreads();
	if(tail_owner) {
		if(out.V) out.V[(size_t)cidx * out.V_stride + N - 1] = sm.tail.x;
		if(out.S) out.S[(size_t)cidx * out.S_stride + N - 1] = sm.tail.S;
		if(out.active) {
			const double pred = (0. + 1. * sm.tail.x) - sm.tail.pz;
			out.active[(size_t)cidx * out.active_stride + N - 1] = pred < 0. ? 1 : 0;
		}
	}
	if(tid == 0) {
		if(out.n_steps) out.n_steps[cidx] = sm.n_steps;
		if(out.inner_total) out.inner_total[cidx] = sm.inner_total;
		if(out.computed) out.computed[cidx] = sm.computed;
		if(out.status) out.status[cidx] = sm.status;
	}
	if(out.value) {
		// PiecewiseLinear::interpolate (Core/Interpolant.hpp:153-181,331-374): the thread owning
		// the bracketing interval writes the value
		const double s0 = b.spot ? b.spot[cidx] : K;
		const double Sfirst = refined_node(axis, b.refine, 0), Slast = refined_node(axis, b.refine, N - 1);
#pragma unroll
		for(int j = 0; j < M; ++j) {
			const int i = i0 + j;
			if(i >= N - 1) continue;
			const double Si = refined_node(axis, b.refine, i);
			const double Sp = refined_node(axis, b.refine, i + 1);
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

// American sweeps with constant steps, theta schemes or BDF1, a generated obstacle, on grids of
// 513 < N <= 2049 nodes.  Everything else: k_resident_sweep.
bool lean_supported(const DeviceBatch &b) {
	const char *env = getenv("QPDE_RESIDENT_LEAN");             // "0": A/B against k_resident_sweep (read per call)
	if(env && env[0] == '0') return false;
	if(!b.american || b.n_controls != 0 || b.n_exercise > 0 || b.variable_target || b.forward || b.source_table) return false;
	if(b.scheme != QPDE_SCHEME_RANNACHER && b.scheme != QPDE_SCHEME_CN && b.scheme != QPDE_SCHEME_BDF1) return false;
	if(b.obstacle_kind != QPDE_PAYOFF_PUT && b.obstacle_kind != QPDE_PAYOFF_CALL) return false;
	if(b.N <= 8 * 64 + 1 || b.N > 8 * LP + 1) return false;
	return true;
}

int launch_lean(const DeviceBatch &b, const DeviceResult &out, cudaStream_t stream, long long *launches) {
	if(!lean_supported(b)) return QPDE_ERR_UNSUPPORTED;
	auto kern = b.obstacle_kind == QPDE_PAYOFF_PUT ? k_lean_sweep<8, QPDE_PAYOFF_PUT> : k_lean_sweep<8, QPDE_PAYOFF_CALL>;
	size_t smem = sizeof(LeanShared<8>);
	const char *one = getenv("QPDE_LEAN_ONE_PER_SM");          // tuning: pad the request so that only one CTA fits an SM
	if(one && one[0] == '1') smem = 120 * 1024;
	if(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return QPDE_ERR_CUDA;
	if(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) return QPDE_ERR_CUDA;
	kern<<<b.C, LP, smem, stream>>>(b, out);
	*launches += 1;
	return QPDE_OK;
}

} // namespace qpde
