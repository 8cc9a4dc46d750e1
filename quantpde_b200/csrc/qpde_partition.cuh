// qpde_partition.cuh — the three-level partition solve of one tridiagonal
// system per CTA, shared by the RESIDENT kernels.
//
//   level 1  thread : rows 1 .. M-1 of each thread eliminated serially
//                     (Thomas), leaving one separator unknown per thread (its
//                     row 0) and two spike vectors Vs, Ws;
//   level 2  warp   : the separators of lanes 1 .. 31 by parallel cyclic
//                     reduction (5 stages of SHFL); the couplings to the warp's
//                     own separator (lane 0) and to the next warp's become two
//                     more spikes V2, W2;
//   level 3  CTA    : one separator per warp -> a W x W tridiagonal system that
//                     every warp reduces redundantly across its lanes 0 .. W-1
//                     (ceil(log2 W) stages) from a shared-memory hand-over.
//
// factorize() caches everything that depends only on the matrix in registers;
// solve() is then right-hand-side work: one __syncthreads each.
//
// This replaces SparseLUSolver::initialize / solve of the reference
// (QuantPDE/src/Bindings/Eigen.hpp:143-165): no pivoting, which is safe because
// every matrix on this path is a strictly row-diagonally-dominant M-matrix
// (I + h L [+ penalty], L from BlackScholes::A, BlackScholes.hpp:136-226).
#ifndef QPDE_PARTITION_CUH
#define QPDE_PARTITION_CUH

#include <cuda_runtime.h>

namespace qpde {

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr int PS_MAX_WARPS = 32;    // warps of one system: PT <= 512 threads per CTA, or a cluster of 4 x 256

__host__ __device__ constexpr int ilog2_ceil(int n) { return n <= 1 ? 0 : 1 + ilog2_ceil((n + 1) / 2); }

__device__ __forceinline__ double shfl_up(double v, int d) { return __shfl_up_sync(FULL_MASK, v, d); }
__device__ __forceinline__ double shfl_down(double v, int d) { return __shfl_down_sync(FULL_MASK, v, d); }
__device__ __forceinline__ double shfl_idx(double v, int l) { return __shfl_sync(FULL_MASK, v, l); }
__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(FULL_MASK, v, m); }

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

// Shared-memory hand-over between the warps (one slot per warp).
struct PartitionSmem {
	// factorisation: level 1/2 results that level 3 needs
	double f_Vs7[PS_MAX_WARPS], f_Ws7[PS_MAX_WARPS], f_V2l[PS_MAX_WARPS], f_W2l[PS_MAX_WARPS];
	double f_V2f[PS_MAX_WARPS], f_W2f[PS_MAX_WARPS], f_a0[PS_MAX_WARPS], f_Bp[PS_MAX_WARPS], f_Ct[PS_MAX_WARPS];
	// right-hand side
	double s_G7[PS_MAX_WARPS], s_G2l[PS_MAX_WARPS], s_P0[PS_MAX_WARPS], s_G2f[PS_MAX_WARPS];
};

// The CTAs whose warps together hold ONE tridiagonal system: a single CTA, or the CL CTAs of
// a thread-block cluster (each warp publishes its separator data in the shared memory of every
// CTA of the cluster — distributed shared memory — and the barrier is the cluster's).
template <int CL>
struct PartitionTeam {
	PartitionSmem *peer[CL];   // PartitionSmem of CTA q of the cluster, own included (generic pointers)
	int rank;                  // this CTA
};

template <int CL>
__device__ __forceinline__ void team_sync() {
	if(CL == 1) __syncthreads();
	else asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int M, int PT, int CL = 1>
struct PartitionSolver {
	static constexpr int W = CL * (PT / 32);     // warps of the whole system
	static constexpr int L3_STAGES = ilog2_ceil(W);
	static constexpr int L3_ALLOC = L3_STAGES > 0 ? L3_STAGES : 1;
	static_assert(W <= PS_MAX_WARPS && W <= 32, "too many warps");

	double f_m[M], f_invd[M], f_u[M], f_Vs[M], f_Ws[M];   // level 1 (index 0 unused)
	double f_a0, f_c0;
	double f_al[5], f_ga[5], f_invB2, f_V2, f_W2;          // level 2
	double f_q1, f_q2, f_q3, f_al3[L3_ALLOC], f_ga3[L3_ALLOC], f_invB3; // level 3 (lanes < W)

	__device__ __forceinline__ void init() {
#pragma unroll
		for(int j = 0; j < M; ++j) { f_m[j] = 0.; f_invd[j] = 1.; f_u[j] = 0.; f_Vs[j] = 0.; f_Ws[j] = 0.; }
		f_a0 = 0.; f_c0 = 0.; f_invB2 = 1.; f_V2 = 0.; f_W2 = 0.;
		f_q1 = 0.; f_q2 = 0.; f_q3 = 0.; f_invB3 = 1.;
#pragma unroll
		for(int s = 0; s < 5; ++s) { f_al[s] = 0.; f_ga[s] = 0.; }
#pragma unroll
		for(int s = 0; s < L3_ALLOC; ++s) { f_al3[s] = 0.; f_ga3[s] = 0.; }
	}

	// Rows (aa, bb, cc)[j] of this thread, j = 0 .. M-1; aa of the very first
	// row and cc of the very last row of the system must be 0.  Padding rows are
	// (0, 1, 0).  Contains one __syncthreads().
	__device__ __forceinline__ void factorize(const double (&aa)[M], const double (&bb)[M],
			const double (&cc)[M], PartitionSmem &sm, int lane, int warp) {
		PartitionTeam<1> team; team.peer[0] = &sm; team.rank = 0;
		factorize(aa, bb, cc, team, lane, warp);
	}
	// `warp` is the index of this warp in the whole system (rank * PT / 32 + warp of the CTA)
	__device__ __forceinline__ void factorize(const double (&aa)[M], const double (&bb)[M],
			const double (&cc)[M], const PartitionTeam<CL> &team, int lane, int warp) {
		PartitionSmem &sm = *team.peer[team.rank];
		f_a0 = aa[0]; f_c0 = cc[0];
		// level 1
		f_invd[1] = rcp(bb[1]);
#pragma unroll
		for(int j = 2; j < M; ++j) {
			f_m[j] = aa[j] * f_invd[j - 1];
			f_invd[j] = rcp(fma(-f_m[j], cc[j - 1], bb[j]));
		}
		double yv[M];
		yv[1] = aa[1];
#pragma unroll
		for(int j = 2; j < M; ++j) yv[j] = -f_m[j] * yv[j - 1];
#pragma unroll
		for(int j = 1; j < M; ++j) f_u[j] = cc[j] * f_invd[j];
		f_Vs[M - 1] = yv[M - 1] * f_invd[M - 1];
		f_Ws[M - 1] = f_u[M - 1];
#pragma unroll
		for(int j = M - 2; j >= 1; --j) {
			f_Vs[j] = fma(-f_u[j], f_Vs[j + 1], yv[j] * f_invd[j]);
			f_Ws[j] = -f_u[j] * f_Ws[j + 1];
		}
		// reduced row of this thread's separator (its row 0)
		const double Vs7p = shfl_up(f_Vs[M - 1], 1), Ws7p = shfl_up(f_Ws[M - 1], 1);
		const bool interior = lane >= 1;
		const double Bpart = fma(-f_c0, f_Vs[1], bb[0]);
		const double Ct = -f_c0 * f_Ws[1];
		double A2 = interior ? -f_a0 * Vs7p : 0.;
		double B2 = interior ? fma(-f_a0, Ws7p, Bpart) : 1.;
		double C2 = interior ? Ct : 0.;
		// level 2: cyclic reduction over lanes 1 .. 31; the couplings of lane 1 to
		// the warp separator and of lane 31 to the next warp's separator become
		// two extra right-hand sides (spikes V2, W2)
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
			B2 = fma(al, Cm, fma(ga, Aq, B2));
			V2 = fma(al, Vm, fma(ga, Vq, V2));
			W2 = fma(al, Wm, fma(ga, Wq, W2));
			A2 = al * Am;
			C2 = ga * Cq;
		}
		f_invB2 = rcp(B2);
		f_V2 = V2 * f_invB2;
		f_W2 = W2 * f_invB2;
		// hand-over to level 3
#pragma unroll
		for(int q = 0; q < CL; ++q) {
			PartitionSmem &to = *team.peer[q];
			if(lane == 31) { to.f_Vs7[warp] = f_Vs[M - 1]; to.f_Ws7[warp] = f_Ws[M - 1]; to.f_V2l[warp] = f_V2; to.f_W2l[warp] = f_W2; }
			if(lane == 1)  { to.f_V2f[warp] = f_V2; to.f_W2f[warp] = f_W2; }
			if(lane == 0)  { to.f_a0[warp] = f_a0; to.f_Bp[warp] = Bpart; to.f_Ct[warp] = Ct; }
		}
		team_sync<CL>();
		// level 3: W x W tridiagonal over the warp separators, reduced by every
		// warp redundantly across its lanes 0 .. W-1
		double A3 = 0., B3 = 1., C3 = 0.;
		f_q1 = 0.; f_q2 = 0.; f_q3 = 0.;
		if(lane < W) {
			const int w = lane;
			const double a0 = sm.f_a0[w];
			const double Vs7q = w > 0 ? sm.f_Vs7[w - 1] : 0., Ws7q = w > 0 ? sm.f_Ws7[w - 1] : 0.;
			const double V2q = w > 0 ? sm.f_V2l[w - 1] : 0., W2q = w > 0 ? sm.f_W2l[w - 1] : 0.;
			const double At = -a0 * Vs7q;
			const double Bt = fma(-a0, Ws7q, sm.f_Bp[w]);
			const double Cw = sm.f_Ct[w];
			A3 = -At * V2q;
			B3 = fma(-Cw, sm.f_V2f[w], fma(-At, W2q, Bt));
			C3 = -Cw * sm.f_W2f[w];
			f_q1 = -a0; f_q2 = -At; f_q3 = -Cw;
		}
#pragma unroll
		for(int s = 0; s < L3_STAGES; ++s) {
			const int dl = 1 << s;
			const double rB = rcp(B3);
			const double rBm = shfl_up(rB, dl), rBq = shfl_down(rB, dl);
			const double Am = shfl_up(A3, dl), Aq = shfl_down(A3, dl);
			const double Cm = shfl_up(C3, dl), Cq = shfl_down(C3, dl);
			const bool okm = lane < W && lane - dl >= 0;
			const bool okq = lane + dl < W;
			const double al = okm ? -A3 * rBm : 0.;
			const double ga = okq ? -C3 * rBq : 0.;
			f_al3[s] = al; f_ga3[s] = ga;
			B3 = fma(al, Cm, fma(ga, Aq, B3));
			A3 = al * Am;
			C3 = ga * Cq;
		}
		f_invB3 = rcp(B3);
	}

	// Solves with the cached factorisation.  r[] is this thread's right-hand
	// side; on return G[] holds the solution of its rows and x_next the solution
	// of the next thread's row 0 (0 for the last thread).  Contains one
	// __syncthreads().
	__device__ __forceinline__ void solve(const double (&r)[M], double (&G)[M], double &x_next,
			PartitionSmem &sm, int lane, int warp) const {
		PartitionTeam<1> team; team.peer[0] = &sm; team.rank = 0;
		solve(r, G, x_next, team, lane, warp);
	}
	__device__ __forceinline__ void solve(const double (&r)[M], double (&G)[M], double &x_next,
			const PartitionTeam<CL> &team, int lane, int warp) const {
		PartitionSmem &sm = *team.peer[team.rank];
		// level 1: G = T^-1 r on the interior rows
		G[1] = r[1];
#pragma unroll
		for(int j = 2; j < M; ++j) G[j] = fma(-f_m[j], G[j - 1], r[j]);
		G[M - 1] = G[M - 1] * f_invd[M - 1];
#pragma unroll
		for(int j = M - 2; j >= 1; --j) G[j] = fma(-f_u[j], G[j + 1], G[j] * f_invd[j]);
		const double G7p = shfl_up(G[M - 1], 1);
		const double P0 = fma(-f_c0, G[1], r[0]);
		// level 2
		double R2 = lane >= 1 ? fma(-f_a0, G7p, P0) : 0.;
#pragma unroll
		for(int s = 0; s < 5; ++s) {
			const int dl = 1 << s;
			const double Rm = shfl_up(R2, dl), Rq = shfl_down(R2, dl);
			R2 = fma(f_al[s], Rm, fma(f_ga[s], Rq, R2));   // al, ga are 0 where the partner is outside
		}
		const double G2 = R2 * f_invB2;
#pragma unroll
		for(int q = 0; q < CL; ++q) {
			PartitionSmem &to = *team.peer[q];
			if(lane == 31) { to.s_G7[warp] = G[M - 1]; to.s_G2l[warp] = G2; }
			if(lane == 1)  to.s_G2f[warp] = G2;
			if(lane == 0)  to.s_P0[warp] = P0;
		}
		team_sync<CL>();
		// level 3 (lanes 0 .. W-1 of every warp)
		double R3 = 0.;
		if(lane < W) {
			R3 = sm.s_P0[lane];
			if(lane > 0) R3 = fma(f_q1, sm.s_G7[lane - 1], fma(f_q2, sm.s_G2l[lane - 1], R3));
			R3 = fma(f_q3, sm.s_G2f[lane], R3);
		}
#pragma unroll
		for(int s = 0; s < L3_STAGES; ++s) {
			const int dl = 1 << s;
			const double Rm = shfl_up(R3, dl), Rq = shfl_down(R3, dl);
			R3 = fma(f_al3[s], Rm, fma(f_ga3[s], Rq, R3));
		}
		const double Y = R3 * f_invB3;            // separator of warp `lane` (0 for lanes >= W)
		const double Yw = shfl_idx(Y, warp);
		const double Yn = warp + 1 < 32 ? shfl_idx(Y, warp + 1) : 0.;   // lane W holds 0; W == 32 has no such lane
		// back substitution, levels 2 and 1
		const double ysep = lane == 0 ? Yw : fma(-Yn, f_W2, fma(-Yw, f_V2, G2));
		double ynext = shfl_down(ysep, 1);
		if(lane == 31) ynext = Yn;
		G[0] = ysep;
#pragma unroll
		for(int j = 1; j < M; ++j) G[j] = fma(-ynext, f_Ws[j], fma(-ysep, f_Vs[j], G[j]));
		x_next = ynext;
	}
};

} // namespace qpde

#endif
